mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s2a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2a_pytest.log
tail -5 gpurun_out/s2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s2a_bench.json 2> gpurun_out/s2a_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s2a_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'] if 'kernel_ms' in d['roofline'] else d.get('kernel_ms'))
PY
