"""
Host-side logic of the multi-GPU path on the CPU: two `gloo` ranks (world_size 2, 127.0.0.1).

The sharding layer (tnmf_b200/distributed.py) does not touch arithmetic, so it can be driven with the CPU oracle
as the per-rank arithmetic provider: every rank updates the activations of its block of samples, the stacked
W-gradient is summed with `SampleSharding.sum_gradient` and the W update runs redundantly on both ranks.  The
result must equal the single-process oracle - the reference's own Cyclic_MU == full-batch identity
(tnmf/TransformInvariantNMF.py:457-465, tnmf/tests/test_minibatch.py:19-20), which is what makes sample sharding
exact.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tnmf_oracle as orc
from tnmf_b200.distributed import SampleSharding, equal_batch_slices, shard_bounds


def test_shard_bounds_partition_the_samples():
    for n in (0, 1, 7, 8, 64, 8192):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_equal_batch_slices_pad_with_empty_batches():
    # same contiguous slices as tnmf/TransformInvariantNMF.py:29-37, padded so that all ranks walk equally many
    assert equal_batch_slices(7, 7, 3) == [slice(0, 3), slice(3, 6), slice(6, 7)]
    assert equal_batch_slices(4, 7, 3) == [slice(0, 3), slice(3, 4), slice(0, 0)]
    assert equal_batch_slices(5, 5, None) == [slice(None)]


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, V, W0, H0, iters, batch_size, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sh = SampleSharding()
        assert sh.is_sharded and sh.world_size == world and sh.rank == rank
        lo, hi = sh.bounds(V.shape[0])
        nmf = orc.OracleNMF(n_atoms=W0.shape[0], atom_shape=W0.shape[2:])
        nmf.V, nmf.W, nmf.H = V[lo:hi].copy(), W0.copy(), H0[lo:hi].copy()
        batches = equal_batch_slices(hi - lo, sh.max_local(V.shape[0]), batch_size)
        energies = []
        for _ in range(iters):
            acc = torch.zeros((2, *W0.shape), dtype=torch.float64)
            for b in batches:                                     # Cyclic_MU over the local minibatches
                if nmf.H[b].shape[0] == 0:
                    continue                                      # padding batch of the shorter shard
                nmf.update_H(b, sparsity=0.05)
                gneg, gpos = nmf.gradient_W(b)
                acc[0] += torch.from_numpy(gneg)
                acc[1] += torch.from_numpy(gpos)
            sh.sum_gradient(acc)                                  # the one collective of the iteration
            orc.multiplicative_update(nmf.W, acc[0].numpy(), acc[1].numpy(), 0.0, tuple(range(-(W0.ndim - 2), 0)))
            e = torch.tensor(float(nmf.energy()), dtype=torch.float64)
            energies.append(float(sh.sum_scalar(e)))
        # W must be bit-identical on all ranks (the update is redundant, the all-reduce result is shared)
        w = torch.from_numpy(nmf.W.copy())
        w0 = w.clone()
        sh.broadcast(w0, 0)
        assert torch.equal(w, w0)
        assert sh.collectives == 2 * iters + 1
        out[rank] = (nmf.W.copy(), nmf.H.copy(), energies)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('batch_size', [None, 2])
def test_two_gloo_ranks_equal_one_process(batch_size):
    rng = np.random.default_rng(5)
    N, C, M, D, A = 5, 2, 3, (12, 10), (4, 3)                    # odd N: shards of 3 and 2 samples
    V = rng.random((N, C) + D)
    W0 = rng.random((M, C) + A)
    orc.normalize(W0, (-2, -1))
    H0 = rng.random((N, M) + orc.transform_shape('valid', D, A))
    iters = 4
    ref = orc.OracleNMF(n_atoms=M, atom_shape=A)
    ref.V, ref.W, ref.H = V, W0.copy(), H0.copy()
    ref_e = []
    for _ in range(iters):
        ref.update_H(slice(None), sparsity=0.05)
        ref.update_W()
        ref_e.append(float(ref.energy()))

    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_main, args=(world, _free_port(), V, W0, H0, iters, batch_size, out), nprocs=world, join=True)
        results = [out[r] for r in range(world)]
    for rank, (W, H, energies) in enumerate(results):
        lo, hi = shard_bounds(N, world, rank)
        assert np.allclose(W, ref.W, rtol=1e-10, atol=1e-14)
        assert np.allclose(H, ref.H[lo:hi], rtol=1e-10, atol=1e-14)
        assert np.allclose(energies, ref_e, rtol=1e-10)
