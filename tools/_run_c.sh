mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "narrow_atoms or allreduce_update_w or tc_reconstruct or random_shapes or nan_guards" > gpurun_out/s2c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2c_pytest.log
tail -25 gpurun_out/s2c_pytest.log
timeout -k 10 300 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "cfg3" > gpurun_out/s2c_cfg3.log 2>&1; echo "rc=$?" >> gpurun_out/s2c_cfg3.log
tail -8 gpurun_out/s2c_cfg3.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/s2c_bench_cfg3.json 2> gpurun_out/s2c_bench_cfg3.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s2c_bench_cfg3.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('kernel_ms') or d.get('kernel_ms'), d['config']['kernel_path'])
PY
tail -3 gpurun_out/s2c_bench_cfg3.err
