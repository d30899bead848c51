"""Synthetic benchmark snapshots in pure numpy / scipy (TEST INFRASTRUCTURE: the reference arm of bench.py and its CPU baseline).

Same random draws, in the same order, as ctgcn_b200/synth.py — `make_adj_list(kind, n, m, K, seed)` returns the torch sparse COO
list of exactly the graph the GPU arm builds for that seed — but nothing here imports the product package or its CUDA library:
`bench.py --impl reference` must not load libctgcn_b200.so (round-1 verdict).  What the list is: the reference's loader contract,
helper.py:51-82 — densest core first, +I on the first matrix, consecutive identical levels dropped — evaluated on exact k-core
numbers (preprocessing/structure_generation.py:32-56 uses networkx; here a vectorised peeling, checked against networkx in
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch


def core_numbers(n: int, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Exact core numbers by vectorised peeling: for k = 1, 2, …: repeatedly delete the nodes whose remaining degree is < k; a
    node deleted while k is being enforced has core number k − 1.  O((n + m) · rounds) numpy work."""
    u = np.asarray(u, dtype=np.int64)
    v = np.asarray(v, dtype=np.int64)
    deg = np.bincount(u, minlength=n) + np.bincount(v, minlength=n)
    core = np.zeros(n, dtype=np.int32)
    alive = np.ones(n, dtype=bool)
    eu, ev = u, v
    k = 1
    n_alive = n
    while n_alive > 0:
        low = alive & (deg < k)
        if not low.any():
            k = max(k + 1, int(deg[alive].min()) + 1) if n_alive else k + 1   # jump to the next level that removes something
            continue
        while low.any():
            core[low] = k - 1
            alive &= ~low
            n_alive -= int(low.sum())
            hit_u, hit_v = low[eu], low[ev]
            dead = hit_u | hit_v
            # an edge with exactly one endpoint deleted lowers the other endpoint's degree
            dec = np.bincount(ev[hit_u & ~hit_v], minlength=n) + np.bincount(eu[hit_v & ~hit_u], minlength=n)
            deg = deg - dec
            eu, ev = eu[~dead], ev[~dead]
            low = alive & (deg < k)
    return core


def _simple(n, u, v, m, rng):
    keep = u != v
    u, v = u[keep], v[keep]
    key = np.unique(np.minimum(u, v) * n + np.maximum(u, v))
    if key.shape[0] > m:
        drop = rng.choice(key.shape[0], size=key.shape[0] - m, replace=False)
        mask = np.ones(key.shape[0], dtype=bool)
        mask[drop] = False
        key = key[mask]
    return key // n, key % n


def er_edges(n: int, m: int, rng: np.random.Generator):
    u = rng.integers(0, n, size=int(m * 1.02) + 16, dtype=np.int64)
    v = rng.integers(0, n, size=u.shape[0], dtype=np.int64)
    return _simple(n, u, v, m, rng)


def powerlaw_edges(n: int, m: int, rng: np.random.Generator, exponent: float = 2.3):
    w = (np.arange(n, dtype=np.float64) + 1.0) ** (-1.0 / (exponent - 1.0))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    k = int(m * 1.15) + 16
    u = np.searchsorted(cdf, rng.random(k))
    v = np.searchsorted(cdf, rng.random(k))
    perm = rng.permutation(n)
    return _simple(n, perm[u], perm[v], m, rng)


def edge_levels(ce: np.ndarray, k: int, levels: str):
    """(list of kept core levels, densest first; per-edge index of the first list entry that holds the edge, 255 = in none)."""
    if levels == "top":                      # the K highest distinct core levels
        kept = np.unique(ce)[::-1][:k]
        lut = np.full(int(ce.max()) + 1, 255, dtype=np.uint8)
        lut[kept] = np.arange(kept.shape[0], dtype=np.uint8)
        return [int(x) for x in kept], lut[ce]
    if levels != "loader":
        raise ValueError(levels)
    kmax = int(ce.max())                     # helper.py:61-64 with max_core = K: files of cores 1..K, reversed
    top = min(k, kmax)
    present = np.zeros(top + 1, dtype=bool)
    present[np.minimum(ce, top)] = True
    kept = [top] + [lv for lv in range(top - 1, 0, -1) if present[lv]]   # helper.py:73-76: a level equal to the previous file is skipped
    lut = np.full(top + 1, 255, dtype=np.uint8)
    for idx, lv in enumerate(kept):
        lut[lv] = idx
    return kept, lut[np.minimum(ce, top)]


def make_adj_list(kind: str, n: int, m: int, k: int, seed: int, levels: str = "top"):
    """(adj_list as K torch sparse COO tensors, stats dict with nnz per matrix and Σ nnz = aggregated edges)."""
    rng = np.random.default_rng(seed)
    u, v = (er_edges if kind == "er" else powerlaw_edges)(n, m, rng)
    core = core_numbers(n, u, v)
    ce = np.minimum(core[u], core[v])
    sel = ce >= 1
    u, v, ce = u[sel], v[sel], ce[sel]
    kept, le = edge_levels(ce, k, levels)
    ok = le != 255
    u, v, le = u[ok], v[ok], le[ok]
    diag = np.arange(n, dtype=np.int64)
    adj, nnz = [], []
    for i in range(len(kept)):
        s = le <= i                                       # nested k-cores: entry i holds every edge of level index ≤ i
        rows = np.concatenate([u[s], v[s]] + ([diag] if i == 0 else []))
        cols = np.concatenate([v[s], u[s]] + ([diag] if i == 0 else []))
        idx = torch.from_numpy(np.vstack((rows, cols)))
        adj.append(torch.sparse_coo_tensor(idx, torch.ones(rows.shape[0], dtype=torch.float32), (n, n)))
        nnz.append(int(rows.shape[0]))
    return adj, dict(k=len(kept), core_levels=kept, nnz_per_core=nnz, edges_aggregated=int(sum(nnz)))


def features(n: int, d: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g, dtype=torch.float32)
