# final single-GPU record of a round: all GPU tests, smoke, the default bench line (with the CPU baseline), the reference arm,
# the cfg3 line, and one full ncu capture each of the cfg2 / cfg3 W-gradient kernels (traffic figures of profiles/ncu_traffic.json)
mkdir -p gpurun_out
T=${TAG:-r02_final}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
tail -c 300 gpurun_out/${T}_bench.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_ref.log
tail -c 200 gpurun_out/${T}_bench_ref.log
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.log 2>&1; echo "cfg3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gradw_ts_kernel" -s 2 -c 1 -o gpurun_out/prof_${T}_cfg2_gw -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/${T}_ncu_gw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gradw_ns_kernel" -s 4 -c 1 -o gpurun_out/prof_${T}_cfg3_gw -f python tools/prof_kernels.py cfg3 > gpurun_out/${T}_ncu_gwn.log 2>&1
tail -1 gpurun_out/${T}_ncu_gwn.log
