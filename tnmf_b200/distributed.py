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
import ctypes
import logging
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

_log = logging.getLogger(__name__)


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


class PeerExchange:
    """Symmetric NVLink peer buffers for `tnmf_allreduce_update_w` (include/tnmf_b200.h): the all-reduce of the W gradient
    fused with the W update in one kernel, instead of an NCCL all-reduce between two kernels.

    Every rank allocates one exchange buffer with torch's symmetric-memory allocator and the ranks map each other's
    buffers (`rendezvous`, a collective over the process group); the kernel then writes its gradient straight into the
    peers' memory.  torch provides the memory and the handle exchange - plumbing; the exchange itself is the library's
    kernel.  Construction is collective; `create` returns None (on every rank alike) where peer memory is not to be had -
    a CPU process group, ranks sharing a device, more than 16 ranks, a driver without peer mappings - and the caller
    keeps the NCCL all-reduce."""

    def __init__(self, lib, problem, group, device, world: int, rank: int):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        nbytes = int(lib.tnmf_peer_buffer_bytes(ctypes.byref(problem), world))
        if nbytes <= 0:
            raise RuntimeError('tnmf_peer_buffer_bytes refused the problem')
        self.buffer = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buffer.zero_()
        self.handle = symm.rendezvous(self.buffer, dist.group.WORLD if group is None else group)
        ptrs = [int(q) for q in self.handle.buffer_ptrs]
        if len(ptrs) != world or int(self.handle.rank) != rank:
            raise RuntimeError('symmetric-memory rendezvous disagrees with the process group')
        self.world = _lib.PeerWorld()
        self.world.world, self.world.rank = world, rank
        for r, q in enumerate(ptrs):
            self.world.buffers[r] = q
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        self.calls = 0
        torch.cuda.synchronize(device)

    @staticmethod
    def create(lib, problem, sharding: 'SampleSharding', device) -> Optional['PeerExchange']:
        from . import _lib
        sh = sharding
        ok = sh.is_sharded and sh.capturable and sh.world_size <= _lib.MAX_PEERS and device.type == 'cuda'
        if ok:
            # one device per rank (the kernel of one rank waits for the kernels of the others: they must run concurrently)
            mine = torch.tensor([device.index if device.index is not None else torch.cuda.current_device()],
                                dtype=torch.int64, device=device)
            every = [torch.empty_like(mine) for _ in range(sh.world_size)]
            dist.all_gather(every, mine, group=sh.group)
            ok = len({int(t.item()) for t in every}) == sh.world_size
        px = None
        if ok:
            try:
                px = PeerExchange(lib, problem, sh.group, device, sh.world_size, sh.rank)
            except Exception as exc:  # pylint: disable=broad-except
                _log.warning('NVLink peer exchange unavailable (%s): the W gradient goes through the NCCL all-reduce', exc)
        if sh.is_sharded and sh.capturable:
            # all or none: a rank that failed to map its peers must not leave the others waiting in the kernel
            flag = torch.tensor([1 if px is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=sh.group)
            if int(flag.item()) == 0:
                px = None
        return px
