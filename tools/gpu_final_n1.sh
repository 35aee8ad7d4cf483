mkdir -p gpurun_out
T=r02_final
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
tail -c 600 gpurun_out/${T}_bench.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_ref.log
tail -c 400 gpurun_out/${T}_bench_ref.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg3"
ncu --set full --clock-control none --import-source on -k regex:"gradw_ts_kernel" -s 2 -c 1 -o gpurun_out/prof_r02_final_cfg2_gw -f $CMD > gpurun_out/${T}_ncu_gw.log 2>&1
tail -2 gpurun_out/${T}_ncu_gw.log
