#!/usr/bin/env python
"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (emdgroup/tnmf,
mounted read-only at /root/reference) in the build container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed here as small .npz files and
the GPU/CPU parity tests compare against them.  `opt_einsum` and `more_itertools` are not installed in the
container; tests/golden/_shim/ maps the two opt_einsum entry points tnmf uses onto numpy.einsum and
provides `chunked` (see the docstrings there).  Nothing else of the reference is touched.

Every fixture stores its inputs (V, and the seed that yields W0/H0 through the reference's own
initialisation order: H first, then W, tnmf/backends/_Backend.py:92-96), and the reference's outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_shim'))
sys.path.insert(0, '/root/reference')

from tnmf.TransformInvariantNMF import TransformInvariantNMF, MiniBatchAlgorithm  # noqa: E402
from tnmf.backends.NumPy_FFT import NumPy_FFT_Backend                             # noqa: E402
from tnmf.backends.NumPy import NumPy_Backend                                     # noqa: E402
from tnmf.backends.PyTorch import PyTorch_Backend                                 # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB')


def energies_callback(store):
    def cb(nmf, it):
        store.append(float(nmf._energy_function()))  # pylint: disable=protected-access
        return True
    return cb


# ----------------------------------------------------------------------------------------------
# 1. the reference's own known-answer case, tnmf/tests/test_1d.py:32-53
# ----------------------------------------------------------------------------------------------
V1D = np.array([[1., 2., 3., 2., 1., 1., 2., 3., 2., 1., 1., 2., 3., 2., 1.],
                [1., 2., 2., 2., 1., 1., 2., 2., 2., 1., 1., 2., 2., 2., 1.],
                [0., 1., 2., 3., 4., 0., 1., 2., 3., 4., 0., 1., 2., 3., 4.]])[:, np.newaxis, :]


def golden_test_1d():
    out = {'V': V1D}
    for mode in ('valid', 'full', 'circular'):
        np.random.seed(42)
        nmf = TransformInvariantNMF(n_atoms=3, atom_shape=(5,), backend='numpy_fft', reconstruction_mode=mode)
        nmf.fit(V1D, inhibition_strength=0.1, n_iterations=10)
        out[f'W_{mode}'] = nmf.W
        out[f'H_{mode}'] = nmf.H
        out[f'R_{mode}'] = nmf.R
        out[f'E_{mode}'] = np.float64(nmf._energy_function())  # pylint: disable=protected-access
    save('ref_test_1d', **out)


# ----------------------------------------------------------------------------------------------
# 2. single operations, all modes, 1-D / 2-D / 3-D, from the numpy_fft backend (float64)
#    and the plain numpy backend (valid only); pytorch backend cross-checked on the fly
# ----------------------------------------------------------------------------------------------
def golden_ops():
    rng = np.random.default_rng(1234)
    cases = {
        'd1': dict(N=3, C=2, M=4, D=(37,), A=(6,)),
        'd2': dict(N=2, C=3, M=5, D=(19, 23), A=(4, 7)),
        'd3': dict(N=2, C=1, M=2, D=(7, 9, 8), A=(3, 2, 4)),
    }
    out = {}
    for cname, c in cases.items():
        for mode in ('valid', 'full', 'circular'):
            be = NumPy_FFT_Backend(reconstruction_mode=mode)
            V = rng.random((c['N'], c['C']) + c['D'])
            np.random.seed(7)
            W, H = be.initialize(V, c['A'], c['M'], None, tuple(range(-len(c['A']), 0)))
            W = W.copy()
            H = H.copy()
            R = be.reconstruct(W, H)
            negH, posH = be.reconstruction_gradient_H(V, W, H)
            negW, posW = be.reconstruction_gradient_W(V, W, H)
            E = be.reconstruction_energy(V, W, H)
            # cross-check with the torch backend (autograd definition of the adjoints)
            import torch
            bt = PyTorch_Backend(reconstruction_mode=mode)
            np.random.seed(7)
            Wt, Ht = bt.initialize(V, c['A'], c['M'], None, tuple(range(-len(c['A']), 0)))
            a, b = bt.reconstruction_gradient_H(V, Wt, Ht)
            assert np.allclose(a.numpy(), negH) and np.allclose(b.numpy(), posH), (cname, mode)
            a, b = bt.reconstruction_gradient_W(V, Wt, Ht)
            assert np.allclose(a.numpy(), negW) and np.allclose(b.numpy(), posW), (cname, mode)
            if mode == 'valid':
                bn = NumPy_Backend(reconstruction_mode=mode)
                np.random.seed(7)
                Wn, Hn = bn.initialize(V, c['A'], c['M'], None, tuple(range(-len(c['A']), 0)))
                assert np.allclose(bn.reconstruct(Wn, Hn), R)
                a, b = bn.reconstruction_gradient_W(V, Wn, Hn)
                assert np.allclose(a, negW) and np.allclose(b, posW)
            key = f'{cname}_{mode}_'
            out.update({key + 'V': V, key + 'W': W, key + 'H': H, key + 'R': R, key + 'negH': negH,
                        key + 'posH': posH, key + 'negW': negW, key + 'posW': posW, key + 'E': np.float64(E)})
    save('ref_ops', **out)


# ----------------------------------------------------------------------------------------------
# 3. full batch fits in 2-D: all modes (float64, numpy_fft) and float32 (numpy backend, valid),
#    with sparsity / inhibition / cross-atom inhibition variants
# ----------------------------------------------------------------------------------------------
def golden_fit_2d():
    rng = np.random.default_rng(99)
    V = rng.random((3, 2, 20, 17))
    out = {'V': V}
    variants = {
        'plain': dict(),
        'sparse': dict(sparsity_H=0.1),
        'inhib': dict(inhibition_strength=0.5),
        'cross': dict(cross_atom_inhibition_strength=0.3, sparsity_H=0.05),
        'all': dict(sparsity_H=0.1, inhibition_strength=0.2, cross_atom_inhibition_strength=0.4),
    }
    for mode in ('valid', 'full', 'circular'):
        for vname, kw in variants.items():
            np.random.seed(5)
            nmf = TransformInvariantNMF(n_atoms=4, atom_shape=(5, 3), backend='numpy_fft', reconstruction_mode=mode)
            traj = []
            nmf.fit(V, n_iterations=20, progress_callback=energies_callback(traj), **kw)
            key = f'{mode}_{vname}_'
            out.update({key + 'W': nmf.W, key + 'H': nmf.H, key + 'E': np.asarray(traj)})
    # custom inhibition range, tnmf/TransformInvariantNMF.py:154-160
    np.random.seed(5)
    nmf = TransformInvariantNMF(n_atoms=4, atom_shape=(5, 3), inhibition_range=(2, 1), backend='numpy_fft')
    traj = []
    nmf.fit(V, n_iterations=20, inhibition_strength=0.7, progress_callback=energies_callback(traj))
    out.update({'valid_range_W': nmf.W, 'valid_range_H': nmf.H, 'valid_range_E': np.asarray(traj)})
    # float32 through the plain numpy backend (the parity oracle named by the task)
    V32 = V.astype(np.float32)
    np.random.seed(5)
    nmf = TransformInvariantNMF(n_atoms=4, atom_shape=(5, 3), backend='numpy')
    traj = []
    nmf.fit(V32, n_iterations=100, sparsity_H=0.1, progress_callback=energies_callback(traj))
    assert nmf.W.dtype == np.float32
    out.update({'f32_W': nmf.W, 'f32_H': nmf.H, 'f32_E': np.asarray(traj)})
    save('ref_fit_2d', **out)


# ----------------------------------------------------------------------------------------------
# 4. minibatch schedules 4-8 and stream, tnmf/tests/test_minibatch.py / test_stream.py style
# ----------------------------------------------------------------------------------------------
def golden_minibatch():
    rng = np.random.default_rng(321)
    V = rng.random((8, 1, 12, 12))
    out = {'V': V}
    for alg in MiniBatchAlgorithm:
        np.random.seed(42)
        nmf = TransformInvariantNMF(n_atoms=3, atom_shape=(4, 4), backend='numpy_fft')
        nmf.fit_minibatches(V, sparsity_H=0.1, algorithm=alg, batch_size=3, n_epochs=5, sag_lambda=0.8,
                            progress_callback=lambda *_: True)
        out[f'{alg.name}_W'] = nmf.W
        out[f'{alg.name}_H'] = nmf.H
        out[f'{alg.name}_E'] = np.float64(nmf._energy_function())  # pylint: disable=protected-access
    np.random.seed(42)
    nmf = TransformInvariantNMF(n_atoms=3, atom_shape=(4, 4), backend='numpy_fft')
    nmf.fit_batch(V, sparsity_H=0.1, n_iterations=5, progress_callback=lambda *_: True)
    out['full_batch_W'] = nmf.W
    out['full_batch_E'] = np.float64(nmf._energy_function())  # pylint: disable=protected-access
    # stream: array and generator input, with and without max_subsamples (tnmf/tests/test_stream.py:47-108)
    np.random.seed(42)
    nmf = TransformInvariantNMF(n_atoms=3, atom_shape=(4, 4), backend='numpy_caching_fft')
    nmf.fit(V, subsample_size=3, batch_size=2, n_epochs=3, algorithm=MiniBatchAlgorithm.Cyclic_MU,
            progress_callback=lambda *_: True)
    out['stream_W'] = nmf.W
    out['stream_H'] = nmf.H
    out['stream_E'] = np.float64(nmf._energy_function())  # pylint: disable=protected-access
    np.random.seed(42)
    nmf = TransformInvariantNMF(n_atoms=3, atom_shape=(4, 4), backend='numpy_caching_fft')
    nmf.fit((v for v in V), subsample_size=3, max_subsamples=2, n_iterations=4, progress_callback=lambda *_: True)
    out['stream2_W'] = nmf.W
    out['stream2_E'] = np.float64(nmf._energy_function())  # pylint: disable=protected-access
    save('ref_minibatch', **out)


# ----------------------------------------------------------------------------------------------
# 5. BASELINE config 1 (the reference's own CPU-runnable case): seed 42 pulse trains, 100x1x1000,
#    5 atoms x 50, 100 iterations, numpy backend, float64
# ----------------------------------------------------------------------------------------------
def golden_cfg1():
    from tnmf.utils.signals import generate_pulse_train
    np.random.seed(42)
    V = np.stack([generate_pulse_train(symbols=['n', '-', '^', 'v', '_'], pulse_length=50, n_pulses=20)[0]
                  for _ in range(100)])
    assert V.shape == (100, 1, 1000)
    nmf = TransformInvariantNMF(n_atoms=5, atom_shape=(50,), backend='numpy')
    traj = []
    np.random.seed(42)   # re-seed so that W0/H0 are reproducible from the stored V alone
    nmf.fit(V, n_iterations=100, progress_callback=energies_callback(traj))
    print('cfg1 final energy', traj[-1])
    # H (100x5x1049 float64, 4 MB) is not stored: W, R-energy trajectory and a few H statistics pin the run
    H = nmf.H
    save('ref_cfg1', V=V, W=nmf.W, E=np.asarray(traj), H_sum=H.sum(axis=(0, 2)), H_max=H.max(axis=(0, 2)),
         H_head=H[:2, :, :64])


if __name__ == '__main__':
    which = sys.argv[1:] or ['test_1d', 'ops', 'fit_2d', 'minibatch', 'cfg1']
    for name in which:
        globals()['golden_' + name]()
