mkdir -p gpurun_out
rm -f gpurun_out/r18_*.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma_vs_oracle or trilinear" > gpurun_out/r18_tma_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r18_tma_tests.log
tail -4 gpurun_out/r18_tma_tests.log
summ() { python - "$1" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(sys.argv[1].split('/')[-1], 'ms/step %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in r['kernel_ms'].items()}, 'frac %.3f'%r['whole_step']['frac_of_fp32_peak'])
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r18_bench_default.log 2>&1; summ gpurun_out/r18_bench_default.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/r18_bench_cfg3.log 2>&1; summ gpurun_out/r18_bench_cfg3.log
for cfg in "1 8" "2 4" "5 1"; do
  set -- $cfg
  TNMF_TMA_WX=$1 TNMF_TMA_WY=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/r18_bench_cfg3_$1x$2.log 2>&1; summ gpurun_out/r18_bench_cfg3_$1x$2.log
done
