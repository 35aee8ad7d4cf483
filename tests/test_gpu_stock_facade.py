"""
The route a tnmf user takes without touching tnmf (INTEGRATION.md 3 i): the UNMODIFIED reference facade
(tnmf/TransformInvariantNMF.py, from the git-ignored install `baseline/_ref`) drives `B200_Backend` through the
abstract backend interface only - including the lateral-inhibition branch that mixes a backend tensor with an ndarray
(tnmf/TransformInvariantNMF.py:258) and the cross-atom branch (`.sum(axis=1, keepdims=True)`, :263).

Recipe and golden energies: tnmf/tests/test_1d.py:17-22,32-53.  Skipped where the reference install is absent.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
GOLDEN_E = {'valid': 2.34946, 'full': 1.87180, 'circular': 3.13228}        # tnmf/tests/test_1d.py:17-22


@pytest.fixture(scope='module')
def RefNMF():
    if not os.path.isdir(os.path.join(REF, 'tnmf')):
        pytest.skip('baseline/_ref (pip install --target of the reference) is not present')
    try:
        import opt_einsum  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden', '_shim'))
    sys.path.insert(0, REF)
    from tnmf.TransformInvariantNMF import TransformInvariantNMF
    return TransformInvariantNMF


V_1D = np.array([[1., 2., 3., 2., 1., 1., 2., 3., 2., 1., 1., 2., 3., 2., 1.],
                 [1., 2., 2., 2., 1., 1., 2., 2., 2., 1., 1., 2., 2., 2., 1.],
                 [0., 1., 2., 3., 4., 0., 1., 2., 3., 4., 0., 1., 2., 3., 4.]])[:, np.newaxis, :]


def _fit(RefNMF, V, mode, b200, n_atoms, atom_shape, **kw):
    from tnmf_b200 import B200_Backend
    np.random.seed(42)
    nmf = RefNMF(n_atoms=n_atoms, atom_shape=atom_shape, backend='numpy_fft', reconstruction_mode=mode)
    if b200:
        nmf._backend = B200_Backend(reconstruction_mode=mode)      # the only line a user adds
    nmf.fit(V, progress_callback=lambda *_: True, **kw)
    return nmf


@pytest.mark.parametrize('mode', list(GOLDEN_E))
def test_stock_facade_drives_b200_backend_1d(RefNMF, mode):
    kw = dict(inhibition_strength=0.1, n_iterations=10)
    ref = _fit(RefNMF, V_1D, mode, False, 3, (5,), **kw)
    got = _fit(RefNMF, V_1D, mode, True, 3, (5,), **kw)
    from tnmf_b200 import B200_Backend
    assert isinstance(got._backend, B200_Backend)
    assert np.isclose(got._energy_function(), GOLDEN_E[mode], rtol=1e-5)
    assert np.isclose(got._energy_function(), ref._energy_function(), rtol=1e-9)
    assert np.allclose(got.W, ref.W, rtol=1e-8, atol=1e-12) and np.allclose(got.H, ref.H, rtol=1e-7, atol=1e-12)
    assert np.allclose(got.R, ref.R, rtol=1e-8, atol=1e-12)
    assert np.allclose(got.R_partial(1), ref.R_partial(1), rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_stock_facade_2d_all_regularisers(RefNMF, dtype):
    rng = np.random.default_rng(9)
    V = rng.random((3, 2, 20, 17)).astype(dtype)
    kw = dict(n_iterations=15, sparsity_H=0.1, inhibition_strength=0.2, cross_atom_inhibition_strength=0.3)
    ref = _fit(RefNMF, V, 'valid', False, 4, (5, 3), **kw)
    got = _fit(RefNMF, V, 'valid', True, 4, (5, 3), **kw)
    tol = 1e-8 if dtype == np.float64 else 1e-3
    assert got.W.dtype == dtype
    assert np.isclose(got._energy_function(), ref._energy_function(), rtol=1e-9 if dtype == np.float64 else 1e-4)
    assert np.abs(got.W - ref.W).max() <= tol * np.abs(ref.W).max()
    assert np.abs(got.H - ref.H).max() <= tol * np.abs(ref.H).max()


def test_stock_facade_minibatch(RefNMF):
    """fit_minibatches of the stock facade (Cyclic_MU): H[s] views written in place, the W-gradient accumulator."""
    from tnmf.TransformInvariantNMF import MiniBatchAlgorithm
    rng = np.random.default_rng(10)
    V = rng.random((5, 1, 18, 16))
    kw = dict(algorithm=MiniBatchAlgorithm.Cyclic_MU, batch_size=2, n_epochs=6)
    ref = _fit(RefNMF, V, 'valid', False, 3, (4, 4), **kw)
    got = _fit(RefNMF, V, 'valid', True, 3, (4, 4), **kw)
    assert np.allclose(got.W, ref.W, rtol=1e-8, atol=1e-12) and np.allclose(got.H, ref.H, rtol=1e-7, atol=1e-12)
