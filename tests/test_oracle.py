"""
Pins the CPU oracle (oracle/tnmf_oracle.py) against
  * the reference's own known-answer values, tnmf/tests/test_1d.py:17-22 (hard-coded below), and
  * the fixtures produced from the unmodified reference by tests/golden/make_golden.py.
CPU only.
"""
import numpy as np
import pytest

from oracle import tnmf_oracle as orc

MODES = ('valid', 'full', 'circular')

# tnmf/tests/test_1d.py:17-22
REFERENCE_TEST_1D_ENERGIES = {'valid': 2.34946, 'full': 1.87180, 'circular': 3.13228}


@pytest.mark.parametrize('mode', MODES)
def test_known_answer_test_1d(mode, golden):
    g = golden('ref_test_1d')
    np.random.seed(42)
    nmf = orc.OracleNMF(n_atoms=3, atom_shape=(5,), reconstruction_mode=mode)
    nmf.fit(g['V'], inhibition_strength=0.1, n_iterations=10)
    assert np.isclose(nmf.energy(), REFERENCE_TEST_1D_ENERGIES[mode])          # the reference's own assertion
    assert np.isclose(nmf.energy(), g[f'E_{mode}'], rtol=1e-12)
    assert np.allclose(nmf.W, g[f'W_{mode}'], rtol=1e-10, atol=1e-13)
    assert np.allclose(nmf.H, g[f'H_{mode}'], rtol=1e-10, atol=1e-13)
    assert np.allclose(nmf.R, g[f'R_{mode}'], rtol=1e-10, atol=1e-13)
    assert np.allclose(nmf.W.sum(axis=-1), 1.0)                                 # tnmf/tests/test_1d.py:90-91


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('case', ['d1', 'd2', 'd3'])
def test_single_operations(case, mode, golden):
    g = golden('ref_ops')
    k = f'{case}_{mode}_'
    V, W, H = g[k + 'V'], g[k + 'W'], g[k + 'H']
    assert H.shape[2:] == orc.transform_shape(mode, V.shape[2:], W.shape[2:])
    assert np.allclose(orc.reconstruct(W, H, mode), g[k + 'R'], rtol=1e-11, atol=1e-12)
    neg, pos = orc.reconstruction_gradient_H(V, W, H, mode)
    assert np.allclose(neg, g[k + 'negH'], rtol=1e-11, atol=1e-12)
    assert np.allclose(pos, g[k + 'posH'], rtol=1e-11, atol=1e-12)
    neg, pos = orc.reconstruction_gradient_W(V, W, H, mode)
    assert np.allclose(neg, g[k + 'negW'], rtol=1e-11, atol=1e-11)
    assert np.allclose(pos, g[k + 'posW'], rtol=1e-11, atol=1e-11)
    assert np.isclose(orc.reconstruction_energy(V, W, H, mode), g[k + 'E'], rtol=1e-12)


VARIANTS = {
    'plain': dict(),
    'sparse': dict(sparsity_H=0.1),
    'inhib': dict(inhibition_strength=0.5),
    'cross': dict(cross_atom_inhibition_strength=0.3, sparsity_H=0.05),
    'all': dict(sparsity_H=0.1, inhibition_strength=0.2, cross_atom_inhibition_strength=0.4),
}


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('variant', list(VARIANTS))
def test_batch_fit_2d(mode, variant, golden):
    g = golden('ref_fit_2d')
    np.random.seed(5)
    nmf = orc.OracleNMF(n_atoms=4, atom_shape=(5, 3), reconstruction_mode=mode)
    traj = []
    nmf.fit(g['V'], n_iterations=20, progress_callback=lambda m, i: traj.append(m.energy()) or True, **VARIANTS[variant])
    k = f'{mode}_{variant}_'
    assert np.allclose(traj, g[k + 'E'], rtol=1e-9)
    assert np.allclose(nmf.W, g[k + 'W'], rtol=1e-8, atol=1e-12)
    assert np.allclose(nmf.H, g[k + 'H'], rtol=1e-8, atol=1e-12)


def test_batch_fit_custom_inhibition_range(golden):
    g = golden('ref_fit_2d')
    np.random.seed(5)
    nmf = orc.OracleNMF(n_atoms=4, atom_shape=(5, 3), inhibition_range=(2, 1))
    nmf.fit(g['V'], n_iterations=20, inhibition_strength=0.7)
    assert np.isclose(nmf.energy(), g['valid_range_E'][-1], rtol=1e-9)
    assert np.allclose(nmf.W, g['valid_range_W'], rtol=1e-8, atol=1e-12)


def test_batch_fit_float32_follows_dtype(golden):
    g = golden('ref_fit_2d')
    V32 = g['V'].astype(np.float32)
    np.random.seed(5)
    nmf = orc.OracleNMF(n_atoms=4, atom_shape=(5, 3))
    traj = []
    nmf.fit(V32, n_iterations=100, sparsity_H=0.1, progress_callback=lambda m, i: traj.append(m.energy()) or True)
    assert nmf.W.dtype == np.float32 and nmf.H.dtype == np.float32
    # two float32 implementations differ by summation order only
    assert np.allclose(traj, g['f32_E'], rtol=1e-4)
    assert np.abs(nmf.W - g['f32_W']).max() <= 1e-3 * np.abs(g['f32_W']).max()
    assert np.abs(nmf.H - g['f32_H']).max() <= 1e-3 * np.abs(g['f32_H']).max()


ALGS = {'Cyclic_MU': 'cyclic_mu', 'ASG_MU': 'asg_mu', 'GSG_MU': 'gsg_mu', 'ASAG_MU': 'asag_mu', 'GSAG_MU': 'gsag_mu'}


@pytest.mark.parametrize('ref_name', list(ALGS))
def test_minibatch_schedules(ref_name, golden):
    g = golden('ref_minibatch')
    np.random.seed(42)
    nmf = orc.OracleNMF(n_atoms=3, atom_shape=(4, 4))
    nmf.fit_minibatches(g['V'], sparsity_H=0.1, algorithm=ALGS[ref_name], batch_size=3, n_epochs=5, sag_lambda=0.8)
    assert np.isclose(nmf.energy(), g[f'{ref_name}_E'], rtol=1e-9)
    assert np.allclose(nmf.W, g[f'{ref_name}_W'], rtol=1e-8, atol=1e-12)
    assert np.allclose(nmf.H, g[f'{ref_name}_H'], rtol=1e-8, atol=1e-12)


def test_cyclic_equals_full_batch(golden):
    """tnmf/tests/test_minibatch.py:19-20: Cyclic_MU and the full batch reach the same energy."""
    g = golden('ref_minibatch')
    assert np.isclose(g['Cyclic_MU_E'], g['full_batch_E'], rtol=1e-10)
    np.random.seed(42)
    nmf = orc.OracleNMF(n_atoms=3, atom_shape=(4, 4))
    nmf.fit_batch(g['V'], sparsity_H=0.1, n_iterations=5)
    assert np.isclose(nmf.energy(), g['full_batch_E'], rtol=1e-9)
    assert np.allclose(nmf.W, g['Cyclic_MU_W'], rtol=1e-8)


def test_stream(golden):
    g = golden('ref_minibatch')
    np.random.seed(42)
    nmf = orc.OracleNMF(n_atoms=3, atom_shape=(4, 4))
    nmf.fit(g['V'], subsample_size=3, batch_size=2, n_epochs=3, algorithm='cyclic_mu')
    assert np.isclose(nmf.energy(), g['stream_E'], rtol=1e-9)
    assert np.allclose(nmf.W, g['stream_W'], rtol=1e-8)
    assert np.allclose(nmf.H, g['stream_H'], rtol=1e-8, atol=1e-12)
    np.random.seed(42)
    nmf = orc.OracleNMF(n_atoms=3, atom_shape=(4, 4))
    nmf.fit((v for v in g['V']), subsample_size=3, max_subsamples=2, n_iterations=4)
    assert np.isclose(nmf.energy(), g['stream2_E'], rtol=1e-9)
    assert np.allclose(nmf.W, g['stream2_W'], rtol=1e-8)


def test_cfg1_reference_run(golden):
    """BASELINE config 1: 100x1x1000 pulse trains, 5 atoms x 50, 100 iterations -> E = 3.808691."""
    g = golden('ref_cfg1')
    np.random.seed(42)
    nmf = orc.OracleNMF(n_atoms=5, atom_shape=(50,))
    traj = []
    nmf.fit(g['V'], n_iterations=100,
            progress_callback=lambda m, i: (traj.append(m.energy()) if i % 10 == 9 else None) or True)
    assert np.isclose(traj[-1], 3.808691, rtol=1e-6)
    assert np.allclose(traj, g['E'][9::10], rtol=1e-8)
    assert np.allclose(nmf.W, g['W'], rtol=1e-6, atol=1e-10)
    assert np.allclose(nmf.H[:2, :, :64], g['H_head'], rtol=1e-6, atol=1e-10)


def test_unknown_mode_raises():
    with pytest.raises(ValueError):
        orc.transform_shape('reflect', (8,), (3,))
    with pytest.raises(ValueError):
        orc.OracleNMF(2, (3,), reconstruction_mode='same')


# ---------------------------------------------------------------------------------------------------------
# the Fourier-domain restatement (CPU baseline of bench.py) against the coordinate-space oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [((3, 2, 4), (20,), (5,)), ((2, 3, 4), (24, 17), (5, 4)), ((2, 1, 2), (6, 7, 5), (2, 3, 2))])
def test_fft_restatement_matches_direct(shape):
    (N, C, M), D, A = shape
    rng = np.random.default_rng(11)
    V = rng.random((N, C) + D)
    W = rng.random((M, C) + A)
    H = rng.random((N, M) + orc.transform_shape('valid', D, A))
    assert np.allclose(orc.fft_reconstruct(W, H), orc.reconstruct(W, H), rtol=1e-10, atol=1e-12)
    for a, b in zip(orc.fft_gradient_H(V, W, H), orc.reconstruction_gradient_H(V, W, H)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-11)
    for a, b in zip(orc.fft_gradient_W(V, W, H), orc.reconstruction_gradient_W(V, W, H)):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-11)


def test_fft_restatement_fit_follows_direct_fit():
    rng = np.random.default_rng(12)
    V = rng.random((3, 2, 16, 12))
    fits = []
    for cls in (orc.OracleNMF, orc.OracleNMF_FFT, orc.OracleNMF_CachingFFT):
        np.random.seed(4)
        nmf = cls(3, (4, 3))
        nmf.fit_batch(V, n_iterations=8, sparsity_H=0.1)
        fits.append(nmf)
    for other in fits[1:]:          # the Fourier-domain forms (plain and with cached spectra) follow the direct one
        assert np.isclose(fits[0].energy(), other.energy(), rtol=1e-9)
        assert np.allclose(fits[0].W, other.W, rtol=1e-8, atol=1e-12)
        assert np.allclose(fits[0].H, other.H, rtol=1e-7, atol=1e-12)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm): one JSON line, the reference's own
    package when baseline/_ref is present (kind "reference"), else the caching-FFT port (kind "port")."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS='1')             # what torchrun exports: bench.py must undo it
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--workload', 'cfg1',
                          '--steps', '2', '--warmup', '1'], check=True, capture_output=True, text=True, env=env).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port')
    assert line['cpu_baseline']['kind'] == ('reference' if os.path.isdir(os.path.join(root, 'baseline', '_ref', 'tnmf')) else 'port')
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['threads']['OMP_NUM_THREADS'] == str(os.cpu_count())
