mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gradient_w or random_shapes or graph_replay" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cfg3 > gpurun_out/s2k_bench.json 2> gpurun_out/s2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s2k_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('kernel_ms') or d.get('kernel_ms'), d['clocks'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg3"
ncu --set full --clock-control none --import-source on -k regex:"gradw_ts_kernel" -s 2 -c 1 -o gpurun_out/prof_r02_final_cfg2_gw -f $CMD > gpurun_out/s2k_ncu_gw.log 2>&1
tail -1 gpurun_out/s2k_ncu_gw.log
