"""
CPU oracle for the shift-invariant NMF multiplicative-update (MU) iteration of emdgroup/tnmf.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package `tnmf_b200/` imports this module; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` do,
and there only as the checker / the timed CPU baseline.

This is a numpy *restatement* (not a copy) of the algorithm of the reference, written as explicit
shift-and-add loops over the atom offsets instead of the reference's im2col/einsum or FFT machinery, so
that it is an independent statement of the same mathematics.  Every function cites the reference
location (relative to /root/reference) whose behaviour it follows.

Parity status: PINNED.
  * against the reference's own known-answer test `tnmf/tests/test_1d.py:17-22,32-53`
    (energies 2.34946 / 1.87180 / 3.13228 for valid / full / circular) -> tests/test_oracle.py
  * against outputs of the unmodified reference run in the build container (numpy, numpy_fft and
    pytorch backends), committed as fixtures under tests/golden/ by tests/golden/make_golden.py.

Conventions (identical to the reference):
    V[n, c, *D]  samples,  W[m, c, *A]  dictionary,  H[n, m, *T]  activations,  R like V.
    p_i = A_i - 1;  T_i = D_i + A_i - 1 ('valid'), D_i - A_i + 1 ('full'), D_i ('circular').
"""
from itertools import count, islice, product
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

EPS = 1.0e-9  # tnmf/TransformInvariantNMF.py:166

MODES = ('valid', 'full', 'circular')


# ---------------------------------------------------------------------------------------------------
# shapes and padding
# ---------------------------------------------------------------------------------------------------
def transform_shape(mode: str, sample_shape: Sequence[int], atom_shape: Sequence[int]) -> Tuple[int, ...]:
    """Extent of H along the shift axes.  Follows tnmf/backends/_Backend.py:60-73."""
    if mode == 'valid':
        return tuple(int(d) + int(a) - 1 for d, a in zip(sample_shape, atom_shape))
    if mode == 'full':
        return tuple(int(d) - int(a) + 1 for d, a in zip(sample_shape, atom_shape))
    if mode == 'circular':
        return tuple(int(d) for d in sample_shape)
    raise ValueError(f'unsupported reconstruction mode {mode!r}')


def pad_activations(H: np.ndarray, atom_shape: Sequence[int], mode: str) -> np.ndarray:
    """Bring H to the extent D+A-1 on which the plain ('valid') correlation is evaluated.

    Follows tnmf/backends/_PyTorchBackend.py:42-52 (and _NumPyBackend.py:35-48):
    valid -> nothing, full -> zeros (p, p), circular -> wrap (p, 0).
    """
    lead = ((0, 0), (0, 0))
    if mode == 'valid':
        return H
    if mode == 'full':
        return np.pad(H, lead + tuple((a - 1, a - 1) for a in atom_shape), mode='constant')
    if mode == 'circular':
        return np.pad(H, lead + tuple((a - 1, 0) for a in atom_shape), mode='wrap')
    raise ValueError(f'unsupported reconstruction mode {mode!r}')


def unpad_gradient(G: np.ndarray, atom_shape: Sequence[int], mode: str, t_shape: Sequence[int]) -> np.ndarray:
    """Adjoint of `pad_activations`: maps a gradient on the padded extent back onto H's extent.

    valid -> identity; full -> crop; circular -> the first p entries of every axis are wrapped around and
    accumulated onto the last p (this is what autograd does in tnmf/backends/_PyTorchBackend.py:98-103 and
    what the slice tables in tnmf/backends/_NumPyFFTBackend.py:62-74 realise in Fourier space).
    """
    if mode == 'valid':
        return G
    k = len(atom_shape)
    if mode == 'full':
        sl = (slice(None), slice(None)) + tuple(slice(a - 1, a - 1 + t) for a, t in zip(atom_shape, t_shape))
        return G[sl].copy()
    if mode == 'circular':
        out = G
        for i, (a, t) in enumerate(zip(atom_shape, t_shape)):
            axis = out.ndim - k + i
            p = a - 1
            body = np.take(out, np.arange(p, p + t), axis=axis).copy()
            if p > 0:
                head = np.take(out, np.arange(0, p), axis=axis)
                idx = [slice(None)] * out.ndim
                idx[axis] = slice(t - p, t)
                body[tuple(idx)] += head
            out = body
        return out
    raise ValueError(f'unsupported reconstruction mode {mode!r}')


def _offsets(atom_shape: Sequence[int]) -> Iterable[Tuple[int, ...]]:
    return product(*(range(a) for a in atom_shape))


# ---------------------------------------------------------------------------------------------------
# the four hot-path operations
# ---------------------------------------------------------------------------------------------------
def reconstruct(W: np.ndarray, H: np.ndarray, mode: str = 'valid') -> np.ndarray:
    """R[n,c,d] = sum_m sum_a W[m,c,a] * Hpad[n,m,d+p-a].

    Follows tnmf/backends/_Backend.py:120-122 as implemented by NumPy.py:122-132 (valid) and
    PyTorch.py:26-43 (all modes: pad, then correlate with the flipped atoms).
    """
    atom_shape = W.shape[2:]
    Hp = pad_activations(H, atom_shape, mode)
    D = tuple(t - a + 1 for t, a in zip(Hp.shape[2:], atom_shape))
    R = np.zeros((H.shape[0], W.shape[1]) + D, dtype=np.result_type(W.dtype, H.dtype))
    for a in _offsets(atom_shape):
        # window of Hpad starting at p-a, length D, per axis
        sl = (slice(None), slice(None)) + tuple(slice(A - 1 - ai, A - 1 - ai + d) for ai, A, d in zip(a, atom_shape, D))
        w = W[(slice(None), slice(None)) + a]              # [m, c]
        R += np.einsum('mc,nm...->nc...', w, Hp[sl])
    return R


def _correlate_with_atoms(X: np.ndarray, W: np.ndarray) -> np.ndarray:
    """G[n,m,t'] = sum_c sum_a W[m,c,a] * X[n,c,t'-p+a]  (X == 0 outside), t' in [0, D+A-1).

    Follows tnmf/backends/NumPy.py:101-119 (zero-padded X, contraction with the unflipped atoms).
    """
    atom_shape = W.shape[2:]
    Xp = np.pad(X, ((0, 0), (0, 0)) + tuple((a - 1, a - 1) for a in atom_shape), mode='constant')
    Tp = tuple(d + a - 1 for d, a in zip(X.shape[2:], atom_shape))
    G = np.zeros((X.shape[0], W.shape[0]) + Tp, dtype=np.result_type(W.dtype, X.dtype))
    for a in _offsets(atom_shape):
        sl = (slice(None), slice(None)) + tuple(slice(ai, ai + t) for ai, t in zip(a, Tp))
        w = W[(slice(None), slice(None)) + a]              # [m, c]
        G += np.einsum('mc,nc...->nm...', w, Xp[sl])
    return G


def reconstruction_gradient_H(V: np.ndarray, W: np.ndarray, H: np.ndarray, mode: str = 'valid'
                              ) -> Tuple[np.ndarray, np.ndarray]:
    """(neg, pos) of the H gradient; both have H's shape.

    Follows tnmf/backends/_Backend.py:110-118 / NumPy.py:93-120: neg from V, pos from R = reconstruct(W, H).
    """
    atom_shape = W.shape[2:]
    R = reconstruct(W, H, mode)
    t_shape = H.shape[2:]
    neg = unpad_gradient(_correlate_with_atoms(V, W), atom_shape, mode, t_shape)
    pos = unpad_gradient(_correlate_with_atoms(R, W), atom_shape, mode, t_shape)
    return neg, pos


def _correlate_with_activations(X: np.ndarray, Hp: np.ndarray, atom_shape: Sequence[int]) -> np.ndarray:
    """G[m,c,a] = sum_n sum_d Hpad[n,m,d+p-a] * X[n,c,d].   Follows tnmf/backends/NumPy.py:69-91."""
    D = X.shape[2:]
    G = np.zeros((Hp.shape[1], X.shape[1]) + tuple(atom_shape), dtype=np.result_type(Hp.dtype, X.dtype))
    Xc = np.moveaxis(X, 1, 0).reshape(X.shape[1], -1)                  # [c, n*d]
    for a in _offsets(atom_shape):
        sl = (slice(None), slice(None)) + tuple(slice(A - 1 - ai, A - 1 - ai + d) for ai, A, d in zip(a, atom_shape, D))
        Hm = np.moveaxis(Hp[sl], 1, 0).reshape(Hp.shape[1], -1)        # [m, n*d]
        G[(slice(None), slice(None)) + a] = Hm @ Xc.T
    return G


def reconstruction_gradient_W(V: np.ndarray, W: np.ndarray, H: np.ndarray, mode: str = 'valid'
                              ) -> Tuple[np.ndarray, np.ndarray]:
    """(neg, pos) of the W gradient; both have W's shape.

    Follows tnmf/backends/_Backend.py:100-108 / NumPy.py:69-91: neg from V, pos from R = reconstruct(W, H).
    """
    atom_shape = W.shape[2:]
    R = reconstruct(W, H, mode)
    Hp = pad_activations(H, atom_shape, mode)
    return _correlate_with_activations(V, Hp, atom_shape), _correlate_with_activations(R, Hp, atom_shape)


def reconstruction_energy(V: np.ndarray, W: np.ndarray, H: np.ndarray, mode: str = 'valid'):
    """0.5 * ||V - R||^2.   Follows tnmf/backends/_Backend.py:127-130."""
    R = reconstruct(W, H, mode)
    return 0.5 * np.sum(np.square(V - R))


# ---------------------------------------------------------------------------------------------------
# update arithmetic of the facade
# ---------------------------------------------------------------------------------------------------
def inhibition_kernels(inhibition_range: Sequence[int]) -> Tuple[np.ndarray, ...]:
    """Parabolic 1-D kernels 1 - (j/(r+1))^2, j=-r..r.   Follows tnmf/TransformInvariantNMF.py:163."""
    return tuple(1.0 - (np.arange(-r, r + 1) / (r + 1)) ** 2 for r in inhibition_range)


def convolve_multi_1d(arr: np.ndarray, kernels: Sequence[np.ndarray], axes: Sequence[int]) -> np.ndarray:
    """Separable convolution with zero boundary, centred odd kernels.

    Follows tnmf/backends/_NumPyBackend.py:56-64 (scipy.ndimage.convolve1d, mode='constant', cval=0).
    Restated here with explicit shifts; the kernels used by the facade are symmetric.
    """
    out = np.asarray(arr)
    for axis, kern in zip(axes, kernels):
        kern = np.asarray(kern)
        r = (len(kern) - 1) // 2
        axis = axis % out.ndim
        L = out.shape[axis]
        acc = np.zeros(out.shape, dtype=np.result_type(out.dtype, kern.dtype))
        for j in range(-r, r + 1):
            # convolution: acc[x] += kern[r + j] * out[x - j]
            lo, hi = max(0, j), min(L, L + j)
            if lo >= hi:
                continue
            dst = [slice(None)] * out.ndim
            src = [slice(None)] * out.ndim
            dst[axis] = slice(lo, hi)
            src[axis] = slice(lo - j, hi - j)
            acc[tuple(dst)] += kern[r + j] * out[tuple(src)]
        out = acc.astype(arr.dtype, copy=False)
    return out


def normalize(arr: np.ndarray, axis) -> None:
    """In-place arr /= arr.sum(axis, keepdims).   Follows tnmf/backends/_Backend.py:75-77."""
    arr /= arr.sum(axis=axis, keepdims=True)


def multiplicative_update(arr: np.ndarray, neg: np.ndarray, pos: np.ndarray, sparsity: float = 0.0,
                          normalization_axes=None) -> None:
    """pos += eps (+ sparsity); arr *= neg; arr /= pos; optional normalisation.

    Follows tnmf/TransformInvariantNMF.py:217-238 including the order of the roundings.
    """
    reg = EPS
    if sparsity > 0:
        reg += sparsity
    pos += reg
    arr *= neg
    arr /= pos
    if normalization_axes is not None:
        normalize(arr, normalization_axes)


def minibatch_slices(length: int, batch_size: Optional[int]):
    """Contiguous blocks along n, the last one possibly short.  Follows tnmf/TransformInvariantNMF.py:29-37."""
    if batch_size is None:
        return [slice(None)]
    return [slice(s, min(length, s + batch_size)) for s in range(0, length, batch_size)]


# ---------------------------------------------------------------------------------------------------
# the iteration loops (facade restated)
# ---------------------------------------------------------------------------------------------------
class OracleNMF:
    """Restatement of the iteration logic of tnmf.TransformInvariantNMF on top of the functions above.

    Follows tnmf/TransformInvariantNMF.py:142-186 (constructor), :240-271 (updates), :282-348 (batch),
    :350-504 (minibatch schedules 4-8), :506-531 (stream, fit).  Random numbers come from the global legacy
    numpy RNG in exactly the reference's order (H first, then W: tnmf/backends/_Backend.py:92-96; batch order
    permutations for algorithms 5-8: tnmf/TransformInvariantNMF.py:40-44).
    """
    ALGORITHMS = ('cyclic_mu', 'asg_mu', 'gsg_mu', 'asag_mu', 'gsag_mu')   # reference enum values 4..8

    def __init__(self, n_atoms: int, atom_shape: Sequence[int], inhibition_range=None,
                 reconstruction_mode: str = 'valid'):
        if reconstruction_mode not in MODES:
            raise ValueError(f'unsupported reconstruction mode {reconstruction_mode!r}')
        self.n_atoms = int(n_atoms)
        self.atom_shape = tuple(int(a) for a in atom_shape)
        if inhibition_range is None:
            rng = tuple(a - 1 for a in self.atom_shape)
        elif isinstance(inhibition_range, int):
            rng = (inhibition_range,) * len(self.atom_shape)
        else:
            rng = tuple(inhibition_range)
        assert len(rng) == len(self.atom_shape)
        self.inhibition_kernels = inhibition_kernels(rng)
        self.mode = reconstruction_mode
        self.shift_axes = tuple(range(-len(self.atom_shape), 0))
        self.W = None
        self.H = None
        self.V = None

    # -- initialisation: tnmf/TransformInvariantNMF.py:273-280, tnmf/backends/_Backend.py:83-98
    def initialize(self, V: np.ndarray, keep_W: bool = False) -> None:
        self.V = V
        t_shape = transform_shape(self.mode, V.shape[2:], self.atom_shape)
        self.H = np.asarray(1 - np.random.rand(V.shape[0], self.n_atoms, *t_shape), dtype=V.dtype)
        if not (keep_W and self.W is not None):
            self.W = np.asarray(1 - np.random.rand(self.n_atoms, V.shape[1], *self.atom_shape), dtype=V.dtype)
            normalize(self.W, self.shift_axes)

    def energy(self):
        return reconstruction_energy(self.V, self.W, self.H, self.mode)

    @property
    def R(self):
        return reconstruct(self.W, self.H, self.mode)

    def R_partial(self, i_atom: int):
        """tnmf/backends/_Backend.py:124-125."""
        return reconstruct(self.W[i_atom:i_atom + 1], self.H[:, i_atom:i_atom + 1], self.mode)

    # -- updates: tnmf/TransformInvariantNMF.py:240-271
    def update_H(self, s=slice(None), sparsity=0.0, inhibition=0.0, cross_inhibition=0.0) -> None:
        Hs = self.H[s]
        neg, pos = reconstruction_gradient_H(self.V[s], self.W, Hs, self.mode)
        if inhibition > 0 or cross_inhibition > 0:
            G = convolve_multi_1d(Hs, self.inhibition_kernels, self.shift_axes)
            if inhibition > 0:
                tmp = G - Hs
                tmp *= inhibition
                pos += tmp
            if cross_inhibition > 0:
                tmp = G.sum(axis=1, keepdims=True)
                tmp = -G + tmp
                tmp *= cross_inhibition / (self.n_atoms - 1)
                pos += tmp
        multiplicative_update(Hs, neg, pos, sparsity=sparsity)

    def gradient_W(self, s=slice(None)):
        return reconstruction_gradient_W(self.V[s], self.W, self.H[s], self.mode)

    def update_W(self, s=slice(None)) -> None:
        neg, pos = self.gradient_W(s)
        multiplicative_update(self.W, neg, pos, normalization_axes=self.shift_axes)

    # -- batch: tnmf/TransformInvariantNMF.py:282-348
    def fit_batch(self, V, n_iterations=1000, update_H=True, update_W=True, keep_W=False, sparsity_H=0.0,
                  inhibition_strength=0.0, cross_atom_inhibition_strength=0.0, progress_callback=None):
        assert np.all(V >= 0)
        self.initialize(V, keep_W)
        for it in range(n_iterations):
            if update_H:
                self.update_H(slice(None), sparsity_H, inhibition_strength, cross_atom_inhibition_strength)
            if update_W:
                self.update_W()
            if progress_callback is not None and not progress_callback(self, it):
                break
        return self

    # -- minibatch: tnmf/TransformInvariantNMF.py:350-504
    def _accumulate(self, gneg, gpos, lam, s):
        neg, pos = self.gradient_W(s)
        if lam == 1:
            gneg = gneg + neg
            gpos = gpos + pos
        else:
            gneg = gneg * (1 - lam) + lam * neg
            gpos = gpos * (1 - lam) + lam * pos
        return gneg, gpos

    def fit_minibatches(self, V, algorithm='asg_mu', batch_size=3, n_epochs=1000, sag_lambda=0.2, keep_W=False,
                        sparsity_H=0.0, inhibition_strength=0.0, cross_atom_inhibition_strength=0.0,
                        progress_callback=None):
        assert algorithm in self.ALGORITHMS
        assert np.all(V >= 0)
        # NB: the reference never shuffles the samples (tnmf/TransformInvariantNMF.py:410 compares an Enum
        # with ints, which is always False), so neither do we.
        self.initialize(V, keep_W)
        batches = minibatch_slices(len(V), batch_size)
        kw = dict(sparsity=sparsity_H, inhibition=inhibition_strength, cross_inhibition=cross_atom_inhibition_strength)
        stat = None

        def shuffled():
            order = np.random.permutation(len(batches))
            return [batches[i] for i in order]

        for epoch in range(n_epochs):
            if algorithm == 'cyclic_mu':                                   # :457-465
                gneg, gpos = 0, 0
                for b in batches:
                    self.update_H(b, **kw)
                    gneg, gpos = self._accumulate(gneg, gpos, 1.0, b)
                multiplicative_update(self.W, gneg, gpos, normalization_axes=self.shift_axes)
            elif algorithm == 'asg_mu':                                    # :467-472
                for b in shuffled():
                    self.update_H(b, **kw)
                    self.update_W(b)
            elif algorithm == 'gsg_mu':                                    # :474-479
                b = slice(0, 0)
                for b in shuffled():
                    self.update_H(b, **kw)
                self.update_W(b)
            elif algorithm == 'asag_mu':                                   # :481-491
                if stat is None:
                    stat = (0, 0)
                for b in shuffled():
                    self.update_H(b, **kw)
                    stat = self._accumulate(*stat, sag_lambda, b)
                    # the state's `pos` array itself is handed over, so the `pos += eps` of the update
                    # leaks into the running average exactly as in the reference (:490)
                    multiplicative_update(self.W, stat[0], stat[1], normalization_axes=self.shift_axes)
            elif algorithm == 'gsag_mu':                                   # :493-504
                if stat is None:
                    stat = (0, 0)
                b = slice(0, 0)
                for b in shuffled():
                    self.update_H(b, **kw)
                stat = self._accumulate(*stat, sag_lambda, b)
                multiplicative_update(self.W, stat[0], stat[1], normalization_axes=self.shift_axes)
            if progress_callback is not None and not progress_callback(self, epoch):
                break
        return self

    # -- stream: tnmf/TransformInvariantNMF.py:506-523
    def fit_stream(self, V, subsample_size=3, max_subsamples=None, **kwargs):
        it = iter(V)
        for isub in count(0):
            chunk = list(islice(it, subsample_size))
            if not chunk:
                return self
            self.fit(np.asarray(chunk), keep_W=True, **kwargs)
            if max_subsamples is not None and isub == max_subsamples - 1:
                return self

    # -- kwarg router: tnmf/TransformInvariantNMF.py:525-531
    def fit(self, V, **kwargs):
        if 'subsample_size' in kwargs or 'max_subsamples' in kwargs:
            return self.fit_stream(V, **kwargs)
        if 'batch_size' in kwargs or 'algorithm' in kwargs:
            return self.fit_minibatches(V, **kwargs)
        return self.fit_batch(V, **kwargs)


# ---------------------------------------------------------------------------------------------------
# Fourier-domain restatement ('valid' mode): the algorithm of the reference's numpy_fft / numpy_caching_fft
# backends, used as the multi-threaded CPU baseline of bench.py (scipy.fft with workers=-1, like the reference)
# ---------------------------------------------------------------------------------------------------
def _fft_plan(sample_shape: Sequence[int], atom_shape: Sequence[int]):
    """Transform lengths: next_fast_len(D + T - 1) per shift axis.  Follows tnmf/backends/_NumPyFFTBackend.py:43."""
    from scipy.fft import next_fast_len
    t_shape = transform_shape('valid', sample_shape, atom_shape)
    return tuple(int(next_fast_len(int(d) + int(t) - 1)) for d, t in zip(sample_shape, t_shape))


def fft_reconstruct(W: np.ndarray, H: np.ndarray) -> np.ndarray:
    """R = sum_m H[:, m] (*) W[m] cropped to the sample ('valid' mode), evaluated as a product of spectra summed
    over the atoms.  Follows tnmf/backends/NumPy_FFT.py:16-40,90-93 and _NumPyFFTBackend.py:49-60."""
    from scipy.fft import irfftn, rfftn
    k = W.ndim - 2
    axes = tuple(range(-k, 0))
    atom_shape = W.shape[2:]
    sample_shape = tuple(t - a + 1 for t, a in zip(H.shape[2:], atom_shape))
    shape = _fft_plan(sample_shape, atom_shape)
    Hf = rfftn(H, s=shape, axes=axes, workers=-1)
    Wf = rfftn(W, s=shape, axes=axes, workers=-1)
    Rf = np.einsum('nm...,mc...->nc...', Hf, Wf, optimize=True)
    full = irfftn(Rf, s=shape, axes=axes, workers=-1)
    crop = (slice(None), slice(None)) + tuple(slice(a - 1, a - 1 + d) for a, d in zip(atom_shape, sample_shape))
    return np.ascontiguousarray(full[crop]).astype(H.dtype, copy=False)


def _fft_correlate(X: np.ndarray, Y: np.ndarray, contraction: str, out_shape: Sequence[int], shift: Sequence[int],
                   shape: Sequence[int]) -> np.ndarray:
    """c[k] = sum_d X[d] * Y[d + k] for k = shift_i .. shift_i + out_shape_i - 1 (circular indices), contracted over
    the labelled axes in Fourier space: conj(X^) * Y^.  Follows the 'correlate' branch of
    tnmf/backends/NumPy_FFT.py:28-29 (flip == conjugate spectrum up to a shift)."""
    from scipy.fft import irfftn, rfftn
    k = len(shape)
    axes = tuple(range(-k, 0))
    Xf = rfftn(X, s=shape, axes=axes, workers=-1)
    Yf = rfftn(Y, s=shape, axes=axes, workers=-1)
    Cf = np.einsum(contraction, np.conj(Xf), Yf, optimize=True)
    c = irfftn(Cf, s=shape, axes=axes, workers=-1)
    for ax, (s0, n, L) in enumerate(zip(shift, out_shape, shape)):
        idx = (np.arange(s0, s0 + n) % L)
        c = np.take(c, idx, axis=ax + 2)
    return c


def fft_gradient_H(V: np.ndarray, W: np.ndarray, H: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """neg/pos[n,m,t] = sum_c sum_a W[m,c,a] * X[n,c,t-p+a], X = V / R.  Follows tnmf/backends/NumPy_FFT.py:71-88."""
    atom_shape, sample_shape = W.shape[2:], V.shape[2:]
    shape = _fft_plan(sample_shape, atom_shape)
    R = fft_reconstruct(W, H)
    shift = tuple(-(a - 1) for a in atom_shape)
    out = []
    for X in (V, R):
        out.append(_fft_correlate(W, X, 'mc...,nc...->nm...', H.shape[2:], shift, shape).astype(H.dtype, copy=False))
    return out[0], out[1]


def fft_gradient_W(V: np.ndarray, W: np.ndarray, H: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """neg/pos[m,c,a] = sum_n sum_d H[n,m,d+p-a] * X[n,c,d].  Follows tnmf/backends/NumPy_FFT.py:52-69."""
    atom_shape, sample_shape = W.shape[2:], V.shape[2:]
    shape = _fft_plan(sample_shape, atom_shape)
    R = fft_reconstruct(W, H)
    out = []
    for X in (V, R):
        g = _fft_correlate(X, H, 'nc...,nm...->mc...', atom_shape, (0,) * len(atom_shape), shape)   # g[k], k = p - a
        out.append(np.ascontiguousarray(np.flip(g, axis=tuple(range(2, g.ndim)))).astype(W.dtype, copy=False))
    return out[0], out[1]


class OracleNMF_FFT(OracleNMF):
    """The iteration of OracleNMF with the three contractions evaluated in Fourier space ('valid' mode only)."""

    def __init__(self, n_atoms: int, atom_shape: Sequence[int], inhibition_range=None):
        super().__init__(n_atoms, atom_shape, inhibition_range, 'valid')

    def energy(self):
        return 0.5 * np.sum(np.square(self.V - fft_reconstruct(self.W, self.H)))

    def update_H(self, s=slice(None), sparsity=0.0, inhibition=0.0, cross_inhibition=0.0) -> None:
        assert inhibition == 0 and cross_inhibition == 0, 'the FFT restatement times the plain MU iteration only'
        Hs = self.H[s]
        neg, pos = fft_gradient_H(self.V[s], self.W, Hs)
        multiplicative_update(Hs, neg, pos, sparsity=sparsity)

    def gradient_W(self, s=slice(None)):
        return fft_gradient_W(self.V[s], self.W, self.H[s])


class OracleNMF_CachingFFT(OracleNMF_FFT):
    """OracleNMF_FFT with the spectra kept between uses, the way tnmf/backends/NumPy_CachingFFT.py:52-77,189-281 does:
    the spectrum of V is computed once per fit, that of W once per W update, that of H once per H update, and the
    spectrum of the reconstruction is taken from the (cropped) reconstruction once per update.  Same arithmetic as
    OracleNMF_FFT (tests/test_oracle.py pins it to the coordinate-space oracle); fewer transforms - it is the timed CPU
    baseline of bench.py wherever the reference package itself is not available."""

    def initialize(self, V: np.ndarray, keep_W: bool = False) -> None:
        super().initialize(V, keep_W)
        self._shape = _fft_plan(V.shape[2:], self.atom_shape)
        self._axes = tuple(range(-len(self.atom_shape), 0))
        self._Vf = self._rfft(V)
        self._Wf = None
        self._Hf = None

    def _rfft(self, X):
        from scipy.fft import rfftn
        return rfftn(X, s=self._shape, axes=self._axes, workers=-1)

    def _irfft(self, Xf):
        from scipy.fft import irfftn
        return irfftn(Xf, s=self._shape, axes=self._axes, workers=-1)

    def _spectra(self):
        if self._Wf is None:
            self._Wf = self._rfft(self.W)
        if self._Hf is None:
            self._Hf = self._rfft(self.H)
        return self._Wf, self._Hf

    def _reconstruct(self):
        Wf, Hf = self._spectra()
        full = self._irfft(np.einsum('nm...,mc...->nc...', Hf, Wf, optimize=True))
        crop = (slice(None), slice(None)) + tuple(slice(a - 1, a - 1 + d) for a, d in zip(self.atom_shape, self.V.shape[2:]))
        return np.ascontiguousarray(full[crop]).astype(self.H.dtype, copy=False)

    def _take(self, c, shift, out_shape):
        for ax, (s0, n, L) in enumerate(zip(shift, out_shape, self._shape)):
            c = np.take(c, np.arange(s0, s0 + n) % L, axis=ax + 2)
        return c

    def energy(self):
        return 0.5 * np.sum(np.square(self.V - self._reconstruct()))

    def update_H(self, s=slice(None), sparsity=0.0, inhibition=0.0, cross_inhibition=0.0) -> None:
        assert inhibition == 0 and cross_inhibition == 0 and s == slice(None), 'times the plain batch MU iteration only'
        Wf, _ = self._spectra()
        Rf = self._rfft(self._reconstruct())
        shift = tuple(-(a - 1) for a in self.atom_shape)
        out = []
        for Xf in (self._Vf, Rf):
            c = self._irfft(np.einsum('mc...,nc...->nm...', np.conj(Wf), Xf, optimize=True))
            out.append(self._take(c, shift, self.H.shape[2:]).astype(self.H.dtype, copy=False))
        multiplicative_update(self.H, out[0], out[1], sparsity=sparsity)
        self._Hf = None

    def gradient_W(self, s=slice(None)):
        assert s == slice(None)
        _, Hf = self._spectra()
        Rf = self._rfft(self._reconstruct())
        out = []
        for Xf in (self._Vf, Rf):
            c = self._irfft(np.einsum('nc...,nm...->mc...', np.conj(Xf), Hf, optimize=True))
            g = self._take(c, (0,) * len(self.atom_shape), self.atom_shape)
            out.append(np.ascontiguousarray(np.flip(g, axis=tuple(range(2, g.ndim)))).astype(self.W.dtype, copy=False))
        return out[0], out[1]

    def update_W(self, s=slice(None)) -> None:
        super().update_W(s)
        self._Wf = None
