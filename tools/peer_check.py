"""
Multi-GPU check of the fused all-reduce + W update over NVLink peer memory (tnmf_allreduce_update_w).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py

Every rank holds its own block of samples.  The sharded fit runs twice from the same seeded start - peer exchange on,
peer exchange off (NCCL all-reduce) - and the two must agree: the dictionary is BITWISE equal on all ranks in both runs,
and the peer run equals the NCCL run within float32 summation-order noise carried through 12 iterations (1e-4 of max|W|;
the two sum the ranks' gradients in different orders).  The same for a Cyclic_MU minibatch fit.  Then the cfg2 iteration
is timed with both exchanges (CUDA events, max over ranks).  Prints one JSON line on rank 0; exit code 1 on disagreement.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnmf_b200 import TransformInvariantNMF  # noqa: E402


def fit(V, peer, sharded, iters, algorithm=None):
    np.random.seed(7)
    nmf = TransformInvariantNMF(n_atoms=16, atom_shape=(11, 11), backend='b200', init='device', distributed=sharded,
                                input_is_local_shard=True, peer_exchange=peer)
    if algorithm is None:
        nmf.fit(V, n_iterations=iters)
    else:
        nmf.fit(V, algorithm=algorithm, batch_size=2, n_epochs=iters)
    return nmf


def main():
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=device)
    n_local = 4
    gen = torch.Generator(device=device)
    gen.manual_seed(100 + rank)
    V = torch.rand((n_local, 3, 96, 128), dtype=torch.float32, device=device, generator=gen)
    out = {'world': world}

    # H is drawn per rank with the device generator: make the draws reproducible per rank
    def run(peer, algorithm=None):
        torch.cuda.manual_seed(1000 + rank)
        return fit(V, peer, True, 12, algorithm)

    a, b = run(True), run(False)
    out['peer_active'] = a._peer is not None                                     # pylint: disable=protected-access
    Wa, Wb = a.W_device.clone(), b.W_device.clone()
    gathered = [torch.empty_like(Wa) for _ in range(world)]
    dist.all_gather(gathered, Wa)
    out['peer_bitwise_equal_across_ranks'] = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    dist.all_gather(gathered, Wb)
    out['nccl_bitwise_equal_across_ranks'] = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    out['peer_vs_nccl_max_rel'] = float((Wa - Wb).abs().max() / Wb.abs().max())
    out['H_peer_vs_nccl_max_rel'] = float((a.H_device - b.H_device).abs().max() / b.H_device.abs().max())
    ea, eb = a._energy_function(), b._energy_function()                          # pylint: disable=protected-access
    out['energy_peer'], out['energy_nccl'] = ea, eb

    # minibatch schedule through the same kernel (Cyclic_MU sums the batches, then one exchange per epoch)
    from tnmf_b200 import MiniBatchAlgorithm
    c, d = run(True, MiniBatchAlgorithm.Cyclic_MU), run(False, MiniBatchAlgorithm.Cyclic_MU)
    out['cyclic_peer_vs_nccl_max_rel'] = float((c.W_device - d.W_device).abs().max() / d.W_device.abs().max())

    # timing: the same iteration with both exchanges
    for name, peer in (('peer', True), ('nccl', False)):
        torch.cuda.manual_seed(1000 + rank)
        nmf = TransformInvariantNMF(n_atoms=16, atom_shape=(11, 11), backend='b200', init='device', distributed=True,
                                    input_is_local_shard=True, equal_shards=True, peer_exchange=peer)
        Vb = torch.rand((16, 3, 256, 256), dtype=torch.float32, device=device, generator=gen)
        nmf._initialize_matrices(Vb, keep_W=False)                               # pylint: disable=protected-access
        step = nmf._batch_step()                                                 # pylint: disable=protected-access
        for _ in range(5):
            step()
        torch.cuda.synchronize(device)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            step()
        e1.record()
        torch.cuda.synchronize(device)
        ms = torch.tensor([e0.elapsed_time(e1) / 50], dtype=torch.float64, device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[f'ms_per_step_{name}'] = float(ms.item())
        out[f'finite_{name}'] = bool(torch.isfinite(nmf.energy_device()).item())
        del nmf, step, Vb
    ok = (out['peer_active'] and out['peer_bitwise_equal_across_ranks'] and out['nccl_bitwise_equal_across_ranks']
          and out['peer_vs_nccl_max_rel'] <= 1e-4 and out['cyclic_peer_vs_nccl_max_rel'] <= 1e-4
          and abs(ea - eb) <= 1e-5 * abs(eb) and out['finite_peer'])
    out['ok'] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize(device)
    dist.barrier()
    os._exit(0 if ok else 1)      # no NCCL teardown under captured graphs (see bench.finish_ranks)


if __name__ == '__main__':
    main()
