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

    ``counts[r]`` triples go to rank r (the slice order of the packed arrays; a CPU int64 tensor).  Returns the
    concatenated triples received by this rank.  Works on any backend (NCCL / gloo).  One host synchronisation (the
    received counts size the receive buffers); the three payloads travel as ONE byte all-to-all:
    row int64 | col int32 | val fp32 = 16 bytes per triple, split per destination.
    """
    world = dist.get_world_size(group)
    dev = row.device
    send_counts = [int(c) for c in counts.tolist()]
    recv_counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_counts_t, counts.to(device=dev, dtype=torch.int64), group=group)
    recv_counts = [int(c) for c in recv_counts_t.tolist()]
    n_send, n_recv = sum(send_counts), sum(recv_counts)
    # per-destination blocks [rows | cols | vals] as bytes
    send = torch.empty(16 * n_send, dtype=torch.uint8, device=dev)
    o = 0
    s0 = 0
    for c in send_counts:
        if c:
            send[o:o + 8 * c].view(torch.int64).copy_(row[s0:s0 + c])
            send[o + 8 * c:o + 12 * c].view(torch.int32).copy_(col[s0:s0 + c])
            send[o + 12 * c:o + 16 * c].view(torch.float32).copy_(val[s0:s0 + c])
        o += 16 * c
        s0 += c
    recv = torch.empty(16 * n_recv, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=[16 * c for c in recv_counts],
                           input_split_sizes=[16 * c for c in send_counts], group=group)
    out_row = torch.empty(n_recv, dtype=torch.int64, device=dev)
    out_col = torch.empty(n_recv, dtype=torch.int32, device=dev)
    out_val = torch.empty(n_recv, dtype=torch.float32, device=dev)
    o = 0
    r0 = 0
    for c in recv_counts:
        if c:
            out_row[r0:r0 + c].copy_(recv[o:o + 8 * c].view(torch.int64))
            out_col[r0:r0 + c].copy_(recv[o + 8 * c:o + 12 * c].view(torch.int32))
            out_val[r0:r0 + c].copy_(recv[o + 12 * c:o + 16 * c].view(torch.float32))
        o += 16 * c
        r0 += c
    return out_row, out_col, out_val


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


def upload_sharded(X_host: torch.Tensor, device, group=None):
    """Host -> device copy of the input for the row-sharded fit: every rank is handed the same full ``X`` on the host
    (the reference's contract, distance/base.py:184-186), but uploads only ITS row chunk over PCIe; the full matrix —
    the kNN database every rank searches — is then assembled by one all-gather over NVLink (N D 4 bytes at NVLink
    rate instead of W full copies through the host bridges)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = X_host.shape[0]
    bounds = all_bounds(n, world)
    s, e = bounds[rank]
    X = torch.empty(tuple(X_host.shape), dtype=torch.float32, device=device)
    X[s:e].copy_(X_host[s:e], non_blocking=True)  # casts if the host array is not fp32
    all_gather_rows(X, bounds, rank, group)
    return X


class PeerEmbedding:
    """Double-buffered embedding ``Z[N, q]`` in symmetric memory (``torch.distributed._symmetric_memory``).

    Every rank holds the full embedding; the persistent step kernel (``tdr_umap_run_p2p_f32``) stores each updated
    row into its own buffer AND, through the NVLink peer mappings exposed here, into every peer's buffer, and runs
    the per-iteration cross-GPU barrier on the peer-mapped ``flags`` words — the exchange is part of the compute
    kernel.  Allocation + rendezvous cost tens of milliseconds, so instances are cached per (shape, device, group)
    and reused by later fits (``PeerEmbedding.get``).
    Raises if symmetric memory cannot be set up (the caller then uses the NCCL all-gather path).
    """

    _cache = {}

    @classmethod
    def get(cls, Z0: torch.Tensor, group=None):
        """Instance for this shape, loaded with ``Z0`` (collective: every rank calls it).  The symmetric-memory
        buffers are kept per (device, group) with their capacity and re-used by every later fit that fits in them."""
        key = (Z0.dtype, Z0.device, id(group) if group is not None else None)
        peer = cls._cache.get(key)
        if peer is None or peer.capacity < Z0.numel():
            cls._cache.pop(key, None)
            peer = cls(Z0.numel(), Z0.dtype, Z0.device, group)
            cls._cache[key] = peer
        peer.load(Z0)
        return peer

    @classmethod
    def reserve(cls, n_points: int, device, q: int = 2, dtype=torch.float32, group=None):
        """Allocate (once per process) exchange buffers for embeddings of up to ``n_points`` rows, so that the first
        fit does not pay the symmetric-memory allocation + rendezvous (tens of milliseconds).  Collective."""
        key = (dtype, torch.device(device), id(group) if group is not None else None)
        peer = cls._cache.get(key)
        if peer is None or peer.capacity < n_points * q:
            cls._cache.pop(key, None)
            cls._cache[key] = cls(n_points * q, dtype, torch.device(device), group)

    def __init__(self, capacity: int, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm

        from . import ops

        group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 9:
            raise RuntimeError("tdr_umap_run_p2p_f32 addresses at most 8 peers")
        self.capacity = int(capacity)
        self._flat, self.handles = [], []
        for _ in range(2):
            t = symm.empty((self.capacity,), dtype=dtype, device=device)
            self.handles.append(symm.rendezvous(t, group))
            self._flat.append(t)
        self.bufs = list(self._flat)
        # peer-mapped flag words of the in-kernel exchange barrier: one uint32 per rank
        self.flags = symm.empty((64,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.flag_handle = symm.rendezvous(self.flags, group)
        self.sync = ops.RunSync(device)
        import ctypes

        def arr(ptrs):
            return (ctypes.c_uint64 * len(ptrs))(*ptrs)

        self.ptr_arrays = (arr(self.peer_ptrs(0)), arr(self.peer_ptrs(1)))
        self.flag_ptr_array = arr([int(p) for r, p in enumerate(self.flag_handle.buffer_ptrs) if r != self.rank])
        torch.cuda.synchronize(device)
        self.handles[0].barrier(channel=0)  # flags are zero everywhere before anyone announces an epoch

    def load(self, Z0: torch.Tensor):
        n = Z0.numel()
        self.bufs = [f[:n].view(Z0.shape) for f in self._flat]  # views at offset 0: the peers' base addresses apply
        self.bufs[0].copy_(Z0)
        self.bufs[1].copy_(Z0)
        # no rank may start storing into its peers before every peer has loaded its buffers
        self.handles[0].barrier(channel=0)

    def peer_ptrs(self, i: int):
        return [int(p) for r, p in enumerate(self.handles[i].buffer_ptrs) if r != self.rank]

    def barrier(self, i: int):
        self.handles[i].barrier(channel=0)
