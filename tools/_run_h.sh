mkdir -p gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/halo_check.py > gpurun_out/s2h_halo.log 2>&1; echo "rc=$?" >> gpurun_out/s2h_halo.log
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/s2h_halo.log | head -5
