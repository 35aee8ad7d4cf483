# multi-GPU record of a round (run under `gpurun --gpus N`): the peer-exchange check and the bench line at N ranks
N=${N:-8}
mkdir -p gpurun_out
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py > gpurun_out/r02_final_peer_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_peer_n$N.log
grep -E "^\{|^rc=" gpurun_out/r02_final_peer_n$N.log
( time timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 3 ) > gpurun_out/r02_final_bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_bench_n$N.log
grep -E "^rc=|^real" gpurun_out/r02_final_bench_n$N.log
python - $N <<'PY'
import json,sys
for l in open(f'gpurun_out/r02_final_bench_n{sys.argv[1]}.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['parallelism']); print(d['config'].get('cfg3'))
PY
