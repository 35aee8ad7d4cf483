# full round check on one GPU: all GPU tests, smoke, default bench (with CPU baseline), reference arm
mkdir -p gpurun_out
T=${TAG:-r19}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
tail -c 1500 gpurun_out/${T}_bench.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_ref.log
tail -c 1200 gpurun_out/${T}_bench_ref.log
