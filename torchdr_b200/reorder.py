"""EXPERIMENTAL (opt-in: ``TDR_KNN_REORDER=1`` in ``UMAPAffinity``, single GPU): spatial re-ordering for the pruned kNN sweep.

The pruned sweep of ``csrc/knn_tc.cu`` skips database tiles whose bounding box is too far from the query
tile's box, which only pays when rows that are close in space are close in index.  ``voronoi_tree_order``
creates that locality for an arbitrary input order: the rows are sorted by the leaves of an unbalanced
Voronoi tree (every node hands its points to the nearest of <= ``branch`` centres sampled from the node
itself, one Lloyd iteration, recursing until a node has <= ``leaf`` rows).  ``knn_in_any_order`` runs a kNN
callable on the re-ordered rows and maps the result back.  Design study and measurements:
``scripts/reorder_sim.py``, DESIGN.md section 8.  There is no counterpart in the reference
(``torchdr/distance/base.py`` hands unordered data to FAISS); exactness is untouched because the
pruning rule itself is exact for any order.

Level-synchronous and device-agnostic (plain tensor ops: sort, gather, index_add, batched dot products),
so the host logic is unit-tested on the CPU (``tests/test_host_logic.py``: same graph as the plain path, bit for
bit); the per-level nearest-of-16 search is the piece that becomes a CUDA kernel once the path has been measured.
"""

import torch

_BITS = 4  # 16 children per node -> 4 key bits per level
_MAX_DEPTH = 15


def _assign(X, rows, node_of_row, centres, valid, chunk=16384):
    """Nearest valid centre of every row's own node.  centres: [n_nodes, B, d], valid: [n_nodes, B]."""
    out = torch.empty(rows.numel(), dtype=torch.long, device=X.device)
    cn = (centres * centres).sum(-1)  # [n_nodes, B]
    cn = torch.where(valid, cn, torch.full_like(cn, float("inf")))
    for a in range(0, rows.numel(), chunk):
        r = rows[a:a + chunk]
        nd = node_of_row[a:a + chunk]
        x = X[r]
        dots = torch.einsum("nd,nbd->nb", x, centres[nd])
        d2 = cn[nd] - 2.0 * dots  # + |x|^2, constant per row
        out[a:a + chunk] = d2.argmin(1)
    return out


def voronoi_tree_order(X, branch=16, leaf=128, lloyd=1, generator=None):
    """Permutation (int64, on X's device) that sorts the rows of X[n, d] by the leaves of the tree."""
    assert 2 <= branch <= (1 << _BITS)
    n = X.shape[0]
    dev = X.device
    X = X.float()
    key = torch.zeros(n, dtype=torch.long, device=dev)  # path of the row's node, 4 bits per level, left-aligned
    open_rows = torch.arange(n, device=dev)  # rows whose node may still be split
    for depth in range(_MAX_DEPTH):
        if open_rows.numel() == 0:
            break
        shift = 60 - _BITS * (depth + 1)
        node_keys, node_of_row, sizes = torch.unique(key[open_rows], return_inverse=True, return_counts=True)
        big = sizes > leaf
        if not bool(big.any()):
            break
        keep = big[node_of_row]
        open_rows, node_of_row = open_rows[keep], node_of_row[keep]
        # compact node ids over the nodes that are split at this level
        remap = torch.cumsum(big.long(), 0) - 1
        node_of_row = remap[node_of_row]
        sizes = sizes[big]
        n_nodes = sizes.numel()
        # rows grouped by node (stable), to sample member rows as centres
        order = torch.argsort(node_of_row, stable=True)
        starts = torch.cumsum(sizes, 0) - sizes
        n_centres = torch.clamp(sizes // leaf, min=2, max=branch)  # [n_nodes]
        u = torch.rand((n_nodes, branch), generator=generator, device=dev)
        pick = starts.unsqueeze(1) + torch.clamp((u * sizes.unsqueeze(1)).long(), max=(sizes - 1).unsqueeze(1))
        centres = X[open_rows[order[pick]]]  # [n_nodes, branch, d]
        valid = torch.arange(branch, device=dev).unsqueeze(0) < n_centres.unsqueeze(1)
        child = _assign(X, open_rows, node_of_row, centres, valid)
        for _ in range(lloyd):
            flat = node_of_row * branch + child
            sums = torch.zeros((n_nodes * branch, X.shape[1]), dtype=torch.float32, device=dev)
            sums.index_add_(0, flat, X[open_rows])
            cnt = torch.zeros(n_nodes * branch, dtype=torch.float32, device=dev)
            cnt.index_add_(0, flat, torch.ones_like(flat, dtype=torch.float32))
            has = (cnt > 0).view(n_nodes, branch)
            centres = torch.where(has.unsqueeze(-1), (sums / cnt.clamp_min(1).unsqueeze(1)).view(n_nodes, branch, -1),
                                  centres)
            child = _assign(X, open_rows, node_of_row, centres, valid & has)
        # a node whose rows all chose the same child (duplicates) cannot be split: it becomes a leaf as it is
        flat = node_of_row * branch + child
        cnt = torch.zeros(n_nodes * branch, dtype=torch.long, device=dev)
        cnt.index_add_(0, flat, torch.ones_like(flat))
        stuck = (cnt.view(n_nodes, branch).max(1).values == sizes)[node_of_row]
        key[open_rows] |= torch.where(stuck, torch.zeros_like(child), child) << shift
        open_rows = open_rows[~stuck]
    return torch.argsort(key, stable=True)


def unpermute_knn_rows(perm, idx_p, dist_p, *aligned):
    """Map kNN rows computed on ``X[perm]`` back: neighbour ids become original ids, equal distances within a row
    are put in ascending original id (the tie order of every kernel of this engine), ``aligned`` tensors ([n, k],
    entry-aligned with dist_p, e.g. the affinity values) follow their entries, and row r of the result is the row of
    original point r.  Returns (idx, dist, *aligned) — idx in idx_p's dtype."""
    idx_o = perm[idx_p.long()]
    o1 = torch.argsort(idx_o, dim=1, stable=True)
    o2 = torch.argsort(dist_p.gather(1, o1), dim=1, stable=True)
    order = o1.gather(1, o2)
    out = []
    for tns in (idx_o, dist_p) + tuple(aligned):
        rows = tns.gather(1, order)
        back = torch.empty_like(rows)
        back[perm] = rows
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
