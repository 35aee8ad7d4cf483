mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_reconstruct or random_shapes or graph_replay or nan_guards or energy_callback" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/s2k_bench.json 2> gpurun_out/s2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s2k_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('kernel_ms') or d.get('kernel_ms'), d['clocks'])
PY
