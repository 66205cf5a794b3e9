"""Oracle: row partition across ranks (test infrastructure).

Restates ``torchdr/distributed/__init__.py:209-219`` (chunk bounds: the first
``n % W`` ranks own one extra row) and ``:251-267`` (inverse map).
"""


def chunk_bounds(n, rank, world):
    base, extra = divmod(n, world)
    if rank < extra:
        start = rank * (base + 1)
        return start, start + base + 1
    start = rank * base + extra
    return start, start + base


def owner_of(i, n, world):
    base, extra = divmod(n, world)
    cut = extra * (base + 1)
    if i < cut:
        r = i // (base + 1)
    else:
        r = extra + (i - cut) // base
    return min(max(r, 0), world - 1)
