"""Test-side mirror of the (i,j,k) work deal used by the device code (afesp_b200/csrc/triples.cu: my_triples).

Unique triples i <= j <= k (orbit weights 6/3/1) are enumerated in lexicographic order and dealt round-robin:
work unit t belongs to rank t % nranks.  Every rank's share is independent; the six (T) sums add across ranks."""
from __future__ import annotations


def unique_triples(o: int, strict: bool = False):
    out = []
    for i in range(o):
        for j in range(i, o):
            for k in range(j, o):
                if strict and (i == j or j == k):
                    continue
                w = 1.0 if (i == j == k) else (3.0 if (i == j or j == k) else 6.0)
                out.append((i, j, k, w))
    return out


def my_triples(o: int, rank: int, nranks: int, strict: bool = False):
    return [t for n, t in enumerate(unique_triples(o, strict)) if n % nranks == rank]


def orbit(i: int, j: int, k: int):
    """Distinct ordered triples obtained by permuting (i,j,k)."""
    import itertools

    return sorted(set(itertools.permutations((i, j, k))))


def column_range(ncols: int, rank: int, nranks: int, gran: int = 64):
    """Mirror of Dist::col_range (afesp_b200/csrc/tensor.cuh): contiguous [lo, hi) in multiples of `gran` columns."""
    units = (ncols + gran - 1) // gran
    per = (units + nranks - 1) // nranks
    return min(ncols, per * rank * gran), min(ncols, per * (rank + 1) * gran)
