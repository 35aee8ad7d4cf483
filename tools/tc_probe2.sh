#!/bin/bash
# runs every case of tools/tc_probe2 in its own process
P=./tools/tc_probe2
for c in "8 16 64 0" "40 176 200 0" "40 176 204 0" "48 48 300 4" "72 96 256 16"; do
  timeout 60 $P num $c 2>&1 | tail -1
done
for ts in 0 1; do for N in 16 32 48 64 96 128 176 192; do timeout 60 $P time $N 2 $ts 0; done; done
for ts in 0 1; do for N in 48 96 176; do for h in 2 4 8; do timeout 60 $P time $N 2 $ts $h; done; done; done
