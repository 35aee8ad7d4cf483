"""
tnmf_b200 - a B200 (sm_100a) native backend for the shift-invariant NMF multiplicative-update iteration of
emdgroup/tnmf.  See DESIGN.md for the scope and INTEGRATION.md for how it attaches to the reference.

    from tnmf_b200 import TransformInvariantNMF, MiniBatchAlgorithm, B200_Backend, RowShardedNMF
"""
from .backend import B200_Backend
from .halo import RowShardedNMF
from .nmf import MiniBatchAlgorithm, TransformInvariantNMF

__all__ = ['B200_Backend', 'MiniBatchAlgorithm', 'RowShardedNMF', 'TransformInvariantNMF']
__version__ = '0.1.0'
