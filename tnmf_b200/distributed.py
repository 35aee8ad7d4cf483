"""
Sample sharding across the GPUs of one box.

H, V and R are per-sample, so the iteration shards over samples exactly like the reference's own Cyclic_MU
minibatch schedule does (tnmf/TransformInvariantNMF.py:29-37,457-465; tnmf/tests/test_minibatch.py:19-20 pins
that schedule to the full-batch result): every rank updates the activations of its contiguous block of
samples, the W-gradient numerator and denominator (2*M*C*prod(A) numbers) are summed over ranks with one
all-reduce per W update, and the W update itself runs redundantly - and bit-identically - on every rank.

One process per GPU (torchrun); torch.distributed is the plumbing (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  Nothing here touches the arithmetic.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_samples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of samples owned by `rank`; blocks differ in size by at most one sample."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('invalid rank / world size')
    base, extra = divmod(int(n_samples), world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def equal_batch_slices(n_local: int, n_local_max: int, batch_size: Optional[int]) -> List[slice]:
    """Minibatch slices over the local samples (tnmf/TransformInvariantNMF.py:29-37), padded with empty slices so
    that every rank of a sharded run walks the same number of batches (= the same number of collectives)."""
    if batch_size is None:
        return [slice(None)]
    out = [slice(s, min(n_local, s + batch_size)) for s in range(0, n_local, batch_size)]
    n_batches = (int(n_local_max) + batch_size - 1) // batch_size
    out += [slice(0, 0)] * (n_batches - len(out))
    return out


class SampleSharding:
    """World description + the two collectives of the iteration (W-gradient sum, energy sum)."""

    def __init__(self, group=None, enabled: Optional[bool] = None):
        active = dist.is_available() and dist.is_initialized()
        if enabled is None:
            enabled = active
        if enabled and not active:
            raise RuntimeError('distributed=True needs an initialised torch.distributed process group')
        self.enabled = bool(enabled)
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world_size = dist.get_world_size(group) if self.enabled else 1
        self.collectives = 0

    @property
    def is_sharded(self) -> bool:
        return self.enabled and self.world_size > 1

    @property
    def capturable(self) -> bool:
        """Whether the collectives can be recorded into a CUDA graph (NCCL can, gloo cannot)."""
        return self.enabled and 'nccl' in str(dist.get_backend(self.group)).lower()

    def bounds(self, n_samples: int) -> Tuple[int, int]:
        return shard_bounds(n_samples, self.world_size, self.rank)

    def max_local(self, n_samples: int) -> int:
        return shard_bounds(n_samples, self.world_size, 0)[1]

    def sum_gradient(self, grad: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks of the stacked (neg, pos) W-gradient."""
        if self.is_sharded:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
        return grad

    def sum_scalar(self, value: torch.Tensor) -> torch.Tensor:
        if self.is_sharded:
            dist.all_reduce(value, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
        return value

    def broadcast(self, tensor: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.is_sharded:
            dist.broadcast(tensor, src=src, group=self.group)
            self.collectives += 1
        return tensor
