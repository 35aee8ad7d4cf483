mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow_atoms or random_shapes or nan_guards" > gpurun_out/s2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2i_pytest.log
tail -3 gpurun_out/s2i_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/s2i_bench_cfg3.json 2> gpurun_out/s2i_bench_cfg3.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s2i_bench_cfg3.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('kernel_ms') or d.get('kernel_ms'), d['config']['kernel_path'])
PY
