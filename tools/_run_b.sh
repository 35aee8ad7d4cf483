mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "allreduce_update_w" > gpurun_out/s2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2b_pytest.log
tail -15 gpurun_out/s2b_pytest.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py > gpurun_out/s2b_peer.log 2>&1; echo "rc=$?" >> gpurun_out/s2b_peer.log
tail -25 gpurun_out/s2b_peer.log
