"""
Spatial (halo) sharding on real GPUs over NCCL (tnmf_b200.RowShardedNMF; the pytest version runs two gloo ranks on one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/halo_check.py

A few large samples (cfg2 geometry, 4 x 3x256x256, 16 atoms 11x11) are cut into N bands of activation rows; the result
after 10 iterations is compared with the same fit on rank 0 alone (tnmf_b200.TransformInvariantNMF, same seeded start),
and the iteration is timed on 8 x 3x1024x1024 samples against one GPU (wall clock between device synchronisations,
max over ranks).  One JSON line on rank 0.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnmf_b200 import RowShardedNMF, TransformInvariantNMF  # noqa: E402


def main():
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=device)
    rng = np.random.default_rng(2)
    V = rng.random((4, 3, 256, 256), dtype=np.float32)
    iters = 10
    np.random.seed(11)
    nmf = RowShardedNMF(16, (11, 11))
    nmf.fit(V, n_iterations=iters)
    e_sharded = nmf.energy()
    H, W = nmf.gather_H(), nmf.W
    out = {'world': world, 'energy_sharded': e_sharded, 'kernels': nmf.ops.be.kernel_names(), 'band': [nmf.plan['t0'], nmf.plan['t1']]}
    # timing on samples worth sharding: 8 x 3x1024x1024 (V 100 MB, H 550 MB), the sharded iteration launched eagerly
    Vbig = rng.random((8, 3, 1024, 1024), dtype=np.float32)
    big = RowShardedNMF(16, (11, 11))
    np.random.seed(12)
    big.initialize(Vbig)
    for _ in range(3):
        big.step()
    torch.cuda.synchronize(device)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        big.step()
    torch.cuda.synchronize(device)
    dt = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device=device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out['timing_problem'] = '8 x 3x1024x1024, 16 atoms 3x11x11'
    out['ms_per_step_sharded'] = 1e3 * float(dt.item())
    del big
    torch.cuda.empty_cache()
    ok = True
    if rank == 0:
        np.random.seed(11)
        one = TransformInvariantNMF(n_atoms=16, atom_shape=(11, 11), backend='b200', init='numpy', distributed=False)
        one.fit(V, n_iterations=iters)
        out['energy_single'] = one._energy_function()                                # pylint: disable=protected-access
        out['W_max_rel'] = float(np.abs(W - one.W).max() / np.abs(one.W).max())
        out['H_max_rel'] = float(np.abs(H - one.H).max() / np.abs(one.H).max())
        del one
        np.random.seed(12)
        one = TransformInvariantNMF(n_atoms=16, atom_shape=(11, 11), backend='b200', init='device', distributed=False)
        one._initialize_matrices(torch.from_numpy(Vbig).to(device), keep_W=False)   # pylint: disable=protected-access
        step = one._batch_step()                                                     # pylint: disable=protected-access
        for _ in range(3):
            step()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize(device)
        out['ms_per_step_single'] = 1e3 * (time.perf_counter() - t0) / 10
        ok = (abs(out['energy_sharded'] - out['energy_single']) <= 1e-4 * abs(out['energy_single'])
              and out['W_max_rel'] <= 1e-3 and out['H_max_rel'] <= 1e-3)
        out['ok'] = bool(ok)
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize(device)
    dist.barrier()
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
