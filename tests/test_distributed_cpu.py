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


# ---------------------------------------------------------------------------------------------------------
# spatial (halo) sharding: bands of activation rows per rank, halo exchange, masked W gradient (tnmf_b200/halo.py)
# ---------------------------------------------------------------------------------------------------------
class OracleOps:
    """The CPU oracle as the arithmetic provider of RowShardedNMF (what B200Ops is on the GPU box)."""

    def __init__(self, mode='valid'):
        self.mode = mode

    def setup(self, V_local, atom_shape, n_atoms, W0, H0):
        self.V = np.array(V_local, dtype=np.float64)
        return torch.from_numpy(np.array(W0, dtype=np.float64)), torch.from_numpy(np.array(H0, dtype=np.float64))

    def update_H(self, W, H, sparsity, eps):
        neg, pos = orc.reconstruction_gradient_H(self.V, W.numpy(), H.numpy(), self.mode)
        orc.multiplicative_update(H.numpy(), neg, pos, sparsity)

    def gradient_W(self, W, H, own):
        Hn, Wn = H.numpy(), W.numpy()
        R = orc.reconstruct(Wn, Hn, self.mode)
        H_own = np.zeros_like(Hn)
        H_own[:, :, own[0]:own[1]] = Hn[:, :, own[0]:own[1]]
        Hp = orc.pad_activations(H_own, Wn.shape[2:], self.mode)
        return torch.from_numpy(np.stack([orc._correlate_with_activations(self.V, Hp, Wn.shape[2:]),
                                          orc._correlate_with_activations(R, Hp, Wn.shape[2:])]))

    def apply_W(self, W, grad, eps):
        g = grad.numpy()
        orc.multiplicative_update(W.numpy(), g[0].copy(), g[1].copy(), normalization_axes=tuple(range(2, W.dim())))

    def energy_rows(self, W, H, rows):
        R = orc.reconstruct(W.numpy(), H.numpy(), self.mode)
        d = self.V[:, :, rows[0]:rows[1]] - R[:, :, rows[0]:rows[1]]
        return torch.tensor(0.5 * np.sum(d * d))


def _halo_rank_main(rank, world, port, V, atoms, atom_shape, iters, sparsity, seeds, out, mode='valid'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tnmf_b200.halo import RowShardedNMF
        np.random.seed(seeds[rank])
        nmf = RowShardedNMF(atoms, atom_shape, ops=OracleOps(mode), reconstruction_mode=mode)
        energies = []
        nmf.fit(V, n_iterations=iters, sparsity_H=sparsity,
                progress_callback=lambda m, i: energies.append(m.energy()) or True)
        H = nmf.gather_H()
        out[rank] = (nmf.W, H, energies, nmf.plan, nmf.sharding.exchanges)
    finally:
        dist.destroy_process_group()


def test_row_plan_bands_and_halos():
    from tnmf_b200.halo import row_plan
    for mode in ('valid', 'full'):
        for d, a, world in ((24, 5, 2), (24, 5, 3), (40, 1, 4), (17, 4, 2), (27, 9, 2)):
            p = a - 1
            t = d + p if mode == 'valid' else d - p
            plans = [row_plan(t, d, p, world, r, mode) for r in range(world)]
            assert plans[0]['t0'] == 0 and plans[-1]['t1'] == t
            assert all(x['t1'] == y['t0'] for x, y in zip(plans, plans[1:]))
            for r, pl in enumerate(plans):
                assert pl['lower'][1] - pl['lower'][0] == (p if r else 0)                   # p halo rows from below ...
                assert pl['upper'][1] - pl['upper'][0] == (p if r < world - 1 else 0)       # ... and from above
                # the local problem is an ordinary problem of the same mode: T_local = D_local +- p
                assert pl['upper'][1] == (pl['v1'] - pl['v0']) + (p if mode == 'valid' else -p)
                assert 0 <= pl['v0'] < pl['v1'] <= d
            # the energy rows partition the sample rows
            e = [(pl['e_rows'][0] + pl['v0'], pl['e_rows'][1] + pl['v0']) for pl in plans]
            assert e[0][0] == 0 and e[-1][1] == d and all(x[1] == y[0] for x, y in zip(e, e[1:]))
    with pytest.raises(ValueError):
        row_plan(12, 8, 4, 4, 0)            # bands of 3 rows, halo of 4
    with pytest.raises(NotImplementedError):
        row_plan(8, 8, 4, 2, 0, 'circular')


@pytest.mark.parametrize('mode', ['valid', 'full'])
@pytest.mark.parametrize('world,shape,atom_shape', [(2, (3, 2, 24, 10), (5, 4)), (3, (2, 1, 30, 7), (4, 3)),
                                                    (2, (4, 2, 40), (6,))])
def test_row_sharded_fit_equals_single_process(world, shape, atom_shape, mode):
    """Bands of activation rows + halo exchange + masked W gradient == the single-process fit (float64: identical up to
    summation order), from one seeded start; ranks 1.. seed differently on purpose (W is broadcast from rank 0, and every
    rank's band must come from rank 0's draw for the comparison - so H seeds are equal, W seeds are not needed)."""
    rng = np.random.default_rng(5)
    V = rng.random(shape)
    iters, sparsity, atoms = 8, 0.05, 3
    np.random.seed(31)
    ref = orc.OracleNMF(n_atoms=atoms, atom_shape=atom_shape, reconstruction_mode=mode)
    e_ref = []
    ref.fit_batch(V, n_iterations=iters, sparsity_H=sparsity,
                  progress_callback=lambda m, i: e_ref.append(float(m.energy())) or True)
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_halo_rank_main, args=(world, _free_port(), V, atoms, atom_shape, iters, sparsity, [31] * world, out, mode),
                 nprocs=world, join=True)
        results = [out[r] for r in range(world)]
    W0, H0, e0, _, exchanges = results[0]
    assert exchanges == 2 * iters + iters               # two per iteration + one per energy evaluation
    assert np.allclose(e0, e_ref, rtol=1e-10)
    assert np.allclose(W0, ref.W, rtol=1e-9, atol=1e-14)
    assert np.allclose(H0, ref.H, rtol=1e-9, atol=1e-14)
    for W, _, e, _, _ in results[1:]:
        assert np.array_equal(W, W0) and np.allclose(e, e0, rtol=1e-12)
