"""Row partition and the (few) collectives of the path.

Mirrors ``torchdr/distributed/__init__.py:115-318`` (``DistributedContext``): rank r owns the
contiguous rows ``compute_chunk_bounds(n)``; the first ``n % W`` ranks own one extra row.
Collectives go through ``torch.distributed`` (NCCL on GPUs; the helpers are backend-agnostic so
the host logic is testable under gloo).  Differences from the reference, on purpose:

* edge exchange for the symmetrisation carries int64/int32 indices, not indices cast to fp32
  (``utils/sparse.py:286-293`` corrupts indices >= 2^24);
* the per-iteration exchange of the closed-form (UMAP) update is an all-gather of each rank's
  updated rows instead of an all-reduce of a zero-padded N x q gradient
  (``affinity_matcher.py:395-413``) — same result because plain SGD is row-local.
"""

from typing import Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if is_distributed() else 0


def get_world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


class DistributedContext:
    """``torchdr/distributed/__init__.py:115-318``."""

    def __init__(self, force_enable: bool = False):
        self.force_enable = force_enable
        if is_distributed():
            self.is_initialized = True
            self.rank = dist.get_rank()
            self.world_size = dist.get_world_size()
            import os

            self.local_rank = int(os.environ.get("LOCAL_RANK", self.rank))
        else:
            self.is_initialized = bool(force_enable)
            self.rank = 0
            self.world_size = 1
            self.local_rank = 0

    def compute_chunk_bounds(self, n_samples: int) -> Tuple[int, int]:
        """distributed/__init__.py:209-219."""
        return chunk_bounds(n_samples, self.rank, self.world_size)

    @staticmethod
    def get_rank_for_indices(indices: torch.Tensor, n_samples: int, world_size: int) -> torch.Tensor:
        """distributed/__init__.py:251-267."""
        base, extra = divmod(n_samples, world_size)
        cut = extra * (base + 1)
        late = extra + (indices - cut) // max(base, 1)
        ranks = torch.where(indices < cut, indices // (base + 1), late)
        return torch.clamp(ranks, 0, world_size - 1)


def chunk_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    base, extra = divmod(n, world)
    if rank < extra:
        start = rank * (base + 1)
        return start, start + base + 1
    start = rank * base + extra
    return start, start + base


def all_bounds(n: int, world: int):
    return [chunk_bounds(n, r, world) for r in range(world)]


def exchange_edges(counts: torch.Tensor, row: torch.Tensor, col: torch.Tensor, val: torch.Tensor, group=None):
    """All-to-all of (row, col, val) triples packed by destination rank (``utils/sparse.py:259-309``).

    ``counts[r]`` triples go to rank r (the slice order of the packed arrays).  Returns the
    concatenated triples received by this rank.  Works on any backend (NCCL / gloo).
    """
    world = dist.get_world_size(group)
    send_counts = [int(c) for c in counts.tolist()]
    recv_counts_t = torch.empty(world, dtype=torch.int64, device=row.device)
    dist.all_to_all_single(recv_counts_t, torch.tensor(send_counts, dtype=torch.int64, device=row.device), group=group)
    recv_counts = [int(c) for c in recv_counts_t.tolist()]
    out = []
    for t in (row, col, val):
        recv = torch.empty(sum(recv_counts), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(recv, t.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts,
                               group=group)
        out.append(recv)
    return tuple(out)


def all_gather_rows(Z_full: torch.Tensor, bounds, rank: int, group=None):
    """In place: every rank contributes ``Z_full[start:end]`` of its own chunk; afterwards all rows are current.

    Chunks differ by at most one row, so each is padded to the longest chunk for one
    ``all_gather_into_tensor`` (8 N bytes on the wire for q = 2).
    """
    world = len(bounds)
    longest = max(e - s for s, e in bounds)
    q = Z_full.shape[1]
    s, e = bounds[rank]
    if all(b - a == longest for a, b in bounds) and Z_full.is_contiguous():
        # equal chunks: in-place all-gather, each rank's slice of the output is its own input (no staging copies)
        dist.all_gather_into_tensor(Z_full, Z_full[s:e], group=group)
        return Z_full
    send = torch.zeros((longest, q), dtype=Z_full.dtype, device=Z_full.device)
    send[: e - s] = Z_full[s:e]
    recv = torch.empty((world * longest, q), dtype=Z_full.dtype, device=Z_full.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    for r, (rs, re) in enumerate(bounds):
        if r != rank:
            Z_full[rs:re] = recv[r * longest : r * longest + (re - rs)]
    return Z_full


class PeerEmbedding:
    """Double-buffered embedding ``Z[N, q]`` in symmetric memory (``torch.distributed._symmetric_memory``).

    Every rank holds the full embedding; the step kernel (``tdr_umap_step_p2p_f32``) stores each updated row
    into its own buffer AND, through the NVLink peer mappings exposed here, into every peer's buffer, so the
    per-iteration exchange is part of the compute kernel.  ``barrier`` is the symmetric-memory signal-pad
    barrier (a few microseconds on the stream), needed before the freshly written buffer is read.
    Raises if symmetric memory cannot be set up (the caller then uses the NCCL all-gather path).
    """

    def __init__(self, Z0: torch.Tensor, group=None):
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 9:
            raise RuntimeError("tdr_umap_step_p2p_f32 addresses at most 8 peers")
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm.empty(tuple(Z0.shape), dtype=Z0.dtype, device=Z0.device)
            self.handles.append(symm.rendezvous(t, group))
            self.bufs.append(t)
        # peer-mapped flag words for the native multi-step loop (tdr_umap_run_p2p_f32): one uint32 per rank
        self.flags = symm.empty((64,), dtype=torch.int32, device=Z0.device)
        self.flags.zero_()
        self.flag_handle = symm.rendezvous(self.flags, group)
        self.epoch = 0
        self.bufs[0].copy_(Z0)
        self.bufs[1].copy_(Z0)
        torch.cuda.synchronize(Z0.device)
        self.handles[0].barrier(channel=0)
        import ctypes

        def arr(ptrs):
            return (ctypes.c_uint64 * len(ptrs))(*ptrs)

        self.ptr_arrays = (arr(self.peer_ptrs(0)), arr(self.peer_ptrs(1)))
        self.flag_ptr_array = arr([int(p) for r, p in enumerate(self.flag_handle.buffer_ptrs) if r != self.rank])

    def peer_ptrs(self, i: int):
        return [int(p) for r, p in enumerate(self.handles[i].buffer_ptrs) if r != self.rank]

    def barrier(self, i: int):
        self.handles[i].barrier(channel=0)
