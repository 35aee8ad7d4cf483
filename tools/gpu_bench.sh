# quick bench of cfg2 / cfg3 (no CPU baseline); TAG names the logs
mkdir -p gpurun_out
T=${TAG:-rXX}
summ() { python - "$1" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print(sys.argv[1].split('/')[-1], 'ms/step %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in r['kernel_ms'].items()}, d['config']['kernel_path'], 'E', d['config']['final_energy'])
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/${T}_bench_cfg2.log 2>&1; summ gpurun_out/${T}_bench_cfg2.log || tail -5 gpurun_out/${T}_bench_cfg2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3 $EXTRA > gpurun_out/${T}_bench_cfg3.log 2>&1; summ gpurun_out/${T}_bench_cfg3.log || tail -5 gpurun_out/${T}_bench_cfg3.log
