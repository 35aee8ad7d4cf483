"""
The backend interface this package plugs into.

When emdgroup/tnmf is importable, `Backend` *is* the reference's abstract base class
(tnmf/backends/_Backend.py:13-130), so `B200_Backend` is a genuine subclass and can be registered in the
reference's `backend_map` (tnmf/TransformInvariantNMF.py:168-174).  When it is not (the
GPU box carries no copy of the reference) an interface-only mirror with the same method names, argument
meaning and defaults is used instead; it contains no arithmetic.
"""
from abc import ABC, abstractmethod

sliceNone = slice(None)

try:  # pragma: no cover - depends on the environment
    from tnmf.backends._Backend import Backend as Backend  # type: ignore  # noqa: F401
    HAVE_REFERENCE_PACKAGE = True
except Exception:  # noqa: BLE001 - any import problem means "not available"
    HAVE_REFERENCE_PACKAGE = False

    class Backend(ABC):  # type: ignore[no-redef]
        """Method-for-method mirror of tnmf.backends._Backend.Backend (interface only)."""

        def __init__(self, reconstruction_mode: str = 'valid'):
            self._reconstruction_mode = reconstruction_mode
            self.atom_shape = None
            self.n_samples = None
            self.n_channels = None
            self._sample_shape = None
            self._transform_shape = None
            self._n_shift_dimensions = None
            self._shift_dimensions = None

        @abstractmethod
        def initialize(self, V, atom_shape, n_atoms, W=None, axes_W_normalization=None):
            raise NotImplementedError

        @staticmethod
        @abstractmethod
        def to_ndarray(arr):
            raise NotImplementedError

        @staticmethod
        @abstractmethod
        def normalize(arr, axis=None):
            raise NotImplementedError

        @staticmethod
        @abstractmethod
        def convolve_multi_1d(arr, kernels, axes):
            raise NotImplementedError

        @abstractmethod
        def reconstruction_gradient_W(self, V, W, H, s=sliceNone):
            raise NotImplementedError

        @abstractmethod
        def reconstruction_gradient_H(self, V, W, H, s=sliceNone):
            raise NotImplementedError

        @abstractmethod
        def reconstruct(self, W, H):
            raise NotImplementedError

        def partial_reconstruct(self, W, H, i_atom: int):
            return self.reconstruct(W[i_atom:i_atom + 1], H[:, i_atom:i_atom + 1])

        @abstractmethod
        def reconstruction_energy(self, V, W, H) -> float:
            raise NotImplementedError
