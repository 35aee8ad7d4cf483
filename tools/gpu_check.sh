mkdir -p gpurun_out
rm -f gpurun_out/r15_*.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma_vs_oracle" > gpurun_out/r15_tma_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r15_tma_tests.log
tail -15 gpurun_out/r15_tma_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r15_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r15_bench.log
tail -3 gpurun_out/r15_bench.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r15_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r15_all_tests.log
tail -15 gpurun_out/r15_all_tests.log
