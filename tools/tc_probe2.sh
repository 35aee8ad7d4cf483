#!/bin/bash
# runs every case of tools/tc_probe2 in its own process
P=./tools/tc_probe2
for c in "40 176 204 0"; do timeout 60 $P num $c 2>&1 | tail -1; done
for ts in 0 1; do for N in 48 176; do for h in 0 1 2 3 4 8 12 15; do timeout 60 $P time $N 1 $ts $h; done; done; done
