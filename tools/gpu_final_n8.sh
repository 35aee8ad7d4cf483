mkdir -p gpurun_out
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py > gpurun_out/r02_final_peer_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_peer_n8.log
grep -E "^\{|^rc=" gpurun_out/r02_final_peer_n8.log
( time timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 50 --warmup 3 ) > gpurun_out/r02_final_bench_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_bench_n8.log
grep -E "^rc=|^real" gpurun_out/r02_final_bench_n8.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_final_bench_n8.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism']); print(d['config'].get('cfg3'))
PY
