"""Spatial re-ordering of the rows for inputs WITHOUT index locality.

The pruned sweep of ``csrc/knn_tc.cu`` skips database tiles whose bounding box is too far from the query
tile's box, and the step kernel's gathers of neighbour rows hit L1/L2 when neighbours are close in index — both
only pay when rows that are close in space are close in index.  Many pipelines deliver that (clusters or batches
appended one after the other, like the reference benchmark's generator); for any other order
``index_locality`` detects the lack of it and ``voronoi_tree_order`` creates it: the rows are sorted by the leaves
of an unbalanced Voronoi tree (every node hands its points to the nearest of <= ``branch`` centres sampled from the
node itself, one Lloyd iteration, recursing until a node has <= ``leaf`` rows).  The estimators then run the whole
fit in that order and undo the permutation on the embedding (``neighbor_embedding._fit_transform``); the affinity
classes, whose outputs are index-valued, search in that order and map rows and neighbour ids back
(``unpermute_knn_rows``).  Exactness is untouched: the pruning rule is exact for any order, and in the tree order the
sweep runs in its certified mode (TDR_KNN_PRUNE_CERTIFIED).  There is no counterpart in the reference
(``torchdr/distance/base.py`` hands unordered data to FAISS).  Design study: ``scripts/reorder_sim.py``; measured:
shuffled 1 M x 128 fit 1.33 s -> 0.26 s (``profiles/r2_first.log``).

Level-synchronous; the per-level nearest-of-16 search is a CUDA kernel (``csrc/reorder.cu``) on the device and plain
tensor ops on the CPU, where the host logic is unit-tested (``tests/test_host_logic.py``: same graph as the plain path,
bit for bit); sorts / segment bookkeeping are torch library calls (one-off plumbing around the fit).
"""

import torch

_BITS = 4  # 16 children per node -> 4 key bits per level
_MAX_DEPTH = 15


def _assign(X, rows, node_of_row, centres, valid, chunk=16384):
    """Nearest valid centre of every row's own node.  centres: [n_nodes, B, d], valid: [n_nodes, B]."""
    cn = (centres * centres).sum(-1)  # [n_nodes, B]
    cn = torch.where(valid, cn, torch.full_like(cn, float("inf")))
    if X.is_cuda and X.shape[1] <= 512:
        from . import ops

        return ops.tree_assign(X, rows, node_of_row, centres, cn)  # csrc/reorder.cu: one warp per row
    out = torch.empty(rows.numel(), dtype=torch.long, device=X.device)
    for a in range(0, rows.numel(), chunk):
        r = rows[a:a + chunk]
        nd = node_of_row[a:a + chunk]
        x = X[r]
        dots = torch.einsum("nd,nbd->nb", x, centres[nd])
        d2 = cn[nd] - 2.0 * dots  # + |x|^2, constant per row
        out[a:a + chunk] = d2.argmin(1)
    return out


def index_locality(X, tiles=256, tile_rows=128, sample=32768, perm=None):
    """Mean extent of a 128-row tile's bounding box relative to the data's extent (per dimension, averaged): ~0.05 when
    consecutive rows are neighbours in space (clusters stored one after the other), ~0.7-0.9 when the order carries
    no locality (shuffled, or data with no cluster structure at all).  A fixed-seed sample: O(tiles * d) work.
    ``perm``: measure the order ``X[perm]`` without materialising it."""
    n, d = X.shape
    if n < 2 * tile_rows:
        return 0.0
    g = torch.Generator(device=X.device).manual_seed(0x5EED)
    starts = torch.randint(0, n - tile_rows + 1, (min(tiles, n // tile_rows),), generator=g, device=X.device)
    rows = starts.unsqueeze(1) + torch.arange(tile_rows, device=X.device).unsqueeze(0)
    T = X[rows if perm is None else perm[rows]]  # [tiles, 128, d]
    ext_tile = (T.amax(1) - T.amin(1)).mean(0)
    S = X[torch.randint(0, n, (min(sample, n),), generator=g, device=X.device)]
    ext_all = (S.amax(0) - S.amin(0)).clamp_min(1e-30)
    return float((ext_tile / ext_all).mean())


LOCALITY_THRESHOLD = 0.6  # index_locality above this: the order carries no locality (tiles straddling 2-3 clusters: ~0.4)
LOCALITY_GAIN = 0.5       # the tree order is adopted only if it at least halves that figure
MIN_ROWS_FOR_REORDER = 64 * 128  # below 64 database tiles the kNN kernel does not prune at all


def choose_order(X, mode="auto", generator=None):
    """Permutation to run the fit in, or None to keep the input order.  mode: "input" (never), "tree" (always),
    "auto": only if the input order has no index locality AND the tree order has (data without cluster structure
    gains nothing from any order; it is searched as it is)."""
    if mode == "input" or X.shape[0] < MIN_ROWS_FOR_REORDER:
        return None
    before = index_locality(X) if mode == "auto" else 1.0
    if mode == "auto" and before <= LOCALITY_THRESHOLD:
        return None
    if generator is None:
        generator = torch.Generator(device=X.device).manual_seed(0x7EE)
    perm = voronoi_tree_order(X, generator=generator)
    if mode == "auto" and index_locality(X, perm=perm) > LOCALITY_GAIN * before:
        return None
    return perm


def _child_stats(X, rows, node_of_row, child, n_nodes, branch):
    """Per (node, child): sum of the member rows [n_nodes * branch, d] and member count [n_nodes * branch]."""
    if X.is_cuda and X.shape[1] <= 512:
        from . import ops

        return ops.tree_accumulate(X, rows, node_of_row, child, n_nodes, branch)  # csrc/reorder.cu: per-CTA partial sums
    flat = node_of_row * branch + child
    sums = torch.zeros((n_nodes * branch, X.shape[1]), dtype=torch.float32, device=X.device)
    sums.index_add_(0, flat, X[rows])
    cnt = torch.zeros(n_nodes * branch, dtype=torch.float32, device=X.device)
    cnt.index_add_(0, flat, torch.ones_like(flat, dtype=torch.float32))
    return sums, cnt


def voronoi_tree_order(X, branch=16, leaf=128, lloyd=1, generator=None):
    """Permutation (int64, on X's device) that sorts the rows of X[n, d] by the leaves of the tree.

    Level-synchronous.  ``perm`` holds the rows sorted by their node's key (the path from the root, 4 bits per level)
    at all times, so a node is a run of equal keys: no per-level ``unique`` / group-by, one stable sort per level after
    the keys of the split nodes have received their next 4 bits."""
    assert 2 <= branch <= (1 << _BITS)
    n = X.shape[0]
    dev = X.device
    X = X.float()
    perm = torch.arange(n, device=dev)
    key = torch.zeros(n, dtype=torch.long, device=dev)     # key[i] = key of row perm[i]
    closed = torch.zeros(n, dtype=torch.bool, device=dev)  # closed[i]: the node of perm[i] cannot be split (duplicates)
    for depth in range(_MAX_DEPTH):
        shift = 60 - _BITS * (depth + 1)
        # nodes = runs of equal keys in the sorted order
        first = torch.ones(n, dtype=torch.bool, device=dev)
        first[1:] = key[1:] != key[:-1]
        node_of_pos = torch.cumsum(first.long(), 0) - 1
        starts_all = torch.nonzero(first).squeeze(1)
        sizes_all = torch.diff(starts_all, append=torch.tensor([n], device=dev))
        big = (sizes_all > leaf) & ~closed[starts_all]
        if not bool(big.any()):
            break
        sel = big[node_of_pos]
        pos = torch.nonzero(sel).squeeze(1)               # positions of the rows of the nodes that are split, ascending
        rows = perm[pos]                                   # grouped by node: a node's rows are consecutive
        remap = torch.cumsum(big.long(), 0) - 1            # compact ids over the split nodes
        node_of_row = remap[node_of_pos[pos]]
        sizes = sizes_all[big]
        n_nodes = sizes.numel()
        starts = torch.cumsum(sizes, 0) - sizes
        n_centres = torch.clamp(sizes // leaf, min=2, max=branch)  # [n_nodes]
        u = torch.rand((n_nodes, branch), generator=generator, device=dev)
        pick = starts.unsqueeze(1) + torch.clamp((u * sizes.unsqueeze(1)).long(), max=(sizes - 1).unsqueeze(1))
        centres = X[rows[pick]]  # [n_nodes, branch, d]: member rows sampled as centres
        valid = torch.arange(branch, device=dev).unsqueeze(0) < n_centres.unsqueeze(1)
        child = _assign(X, rows, node_of_row, centres, valid)
        for _ in range(lloyd):
            sums, cnt = _child_stats(X, rows, node_of_row, child, n_nodes, branch)
            has = (cnt > 0).view(n_nodes, branch)
            centres = torch.where(has.unsqueeze(-1), (sums / cnt.clamp_min(1).unsqueeze(1)).view(n_nodes, branch, -1),
                                  centres)
            child = _assign(X, rows, node_of_row, centres, valid & has)
        # a node whose rows all chose the same child (duplicates) cannot be split: it becomes a leaf as it is
        cnt = torch.zeros(n_nodes * branch, dtype=torch.long, device=dev)
        cnt.index_add_(0, node_of_row * branch + child, torch.ones_like(child))
        stuck = (cnt.view(n_nodes, branch).max(1).values == sizes)[node_of_row]
        key[pos] |= torch.where(stuck, torch.zeros_like(child), child) << shift
        closed[pos] = stuck
        key, order = torch.sort(key, stable=True)
        perm, closed = perm[order], closed[order]
    return perm


def unpermute_knn_rows(perm, idx_p, dist_p, *aligned):
    """Map kNN rows computed on ``X[perm]`` back: neighbour ids become original ids, equal distances within a row
    are put in ascending original id (the tie order of every kernel of this engine), ``aligned`` tensors ([n, k],
    entry-aligned with dist_p, e.g. the affinity values) follow their entries, and row r of the result is the row of
    original point r.  Returns (idx, dist, *aligned) — idx in idx_p's dtype.

    The rows arrive sorted by distance (ties by permuted id), so only rows that contain EQUAL distances can need a
    different entry order after the ids are mapped: those (rare: duplicates) are re-sorted, the others only re-labelled."""
    idx_o = perm[idx_p.long()]
    tens = [idx_o, dist_p] + list(aligned)
    tied = (dist_p[:, 1:] == dist_p[:, :-1]).any(1) if dist_p.shape[1] > 1 else torch.zeros(0, dtype=torch.bool)
    if bool(tied.any()):
        r = torch.nonzero(tied).squeeze(1)
        o1 = torch.argsort(idx_o[r], dim=1, stable=True)
        o2 = torch.argsort(dist_p[r].gather(1, o1), dim=1, stable=True)
        order = o1.gather(1, o2)
        tens = [t.clone() if t is dist_p or any(t is a for a in aligned) else t for t in tens]
        for t in tens:
            t[r] = t[r].gather(1, order)
    out = []
    for t in tens:
        back = torch.empty_like(t)
        back[perm] = t
        out.append(back)
    out[0] = out[0].to(idx_p.dtype)
    return tuple(out)


def unpermute_rows(perm, *per_row):
    """Row r of every result = the entry computed for original point r (per-row scalars such as rho, sigma)."""
    out = []
    for tns in per_row:
        back = torch.empty_like(tns)
        back[perm] = tns
        out.append(back)
    return tuple(out)


def knn_in_any_order(X, k, knn_fn, perm=None, **tree_kwargs):
    """Run ``knn_fn(Xp) -> (dist[n, k], idx[n, k])`` on the re-ordered rows ``Xp = X[perm]`` and return the result in
    the original row order with original indices.  Equal distances within a row are put back in ascending original
    index (the order every kernel of this engine uses)."""
    if perm is None:
        perm = voronoi_tree_order(X, **tree_kwargs)
    dist_p, idx_p = knn_fn(X[perm].contiguous())
    idx, dist = unpermute_knn_rows(perm, idx_p, dist_p)
    return dist, idx, perm
