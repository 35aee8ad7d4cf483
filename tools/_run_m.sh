mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_hupdate or random_shapes or nan_guards or graph_replay or float32_100" > gpurun_out/s2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2m_pytest.log
tail -3 gpurun_out/s2m_pytest.log
for wl in cfg2 cfg3; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cfg3 --workload $wl > gpurun_out/s2m_bench_$wl.json 2> gpurun_out/s2m_bench_$wl.err; echo "bench rc=$?"
python - $wl <<'PY'
import json,sys
for l in open(f'gpurun_out/s2m_bench_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('kernel_ms') or d.get('kernel_ms'))
PY
done
