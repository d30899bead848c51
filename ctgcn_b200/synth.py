"""Synthetic snapshot generator for the benchmark configurations.

Produces, per snapshot, exactly what the reference's input pipeline would hand to the model
(preprocessing/structure_generation.py:32-56 → helper.py:51-82), but without materialising K matrices:
exact k-core numbers by bucket peeling (ctgcn_kcore_numbers, C++), then every undirected edge gets the
index of the first list entry that contains it.  The list is the K highest DISTINCT core levels, densest
level first, +I on the first entry, consecutive identical levels dropped.

``SnapshotGraph.coo_list()`` materialises the K torch sparse COO matrices (the reference contract) for
parity tests and the CPU baseline; ``SnapshotGraph.plan(device)`` uploads the union CSR directly
(ctgcn_plan_create_csr).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .plan import build_plan_csr


def _work_device():
    """Sorting / deduplication of the generators runs through torch on the GPU when there is one (data
    generation is setup work, not the measured path); numpy draws the random numbers so that a seed
    means the same graph everywhere."""
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def core_numbers(n: int, u, v) -> np.ndarray:
    """Exact core number of every node of the undirected simple graph with edge list (u, v)."""
    dev = _work_device()
    ut = torch.as_tensor(u, dtype=torch.int64, device=dev)
    vt = torch.as_tensor(v, dtype=torch.int64, device=dev)
    rows = torch.cat([ut, vt])
    cols = torch.cat([vt, ut])
    order = torch.argsort(rows)
    cols = cols[order].to(torch.int32).cpu().numpy()
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0).cpu().numpy()
    core = np.empty(n, dtype=np.int32)
    cols = np.ascontiguousarray(cols)
    _lib.check(_lib.lib.ctgcn_kcore_numbers(n, C.c_void_p(rowptr.ctypes.data), C.c_void_p(cols.ctypes.data),
                                            C.c_void_p(core.ctypes.data)), "ctgcn_kcore_numbers")
    return core


def er_edges(n: int, m: int, rng: np.random.Generator):
    """Erdős–Rényi G(n, m)-like simple undirected graph (self-loops and duplicate pairs removed)."""
    u = rng.integers(0, n, size=int(m * 1.02) + 16, dtype=np.int64)
    v = rng.integers(0, n, size=u.shape[0], dtype=np.int64)
    return _simple(n, u, v, m, rng)


def powerlaw_edges(n: int, m: int, rng: np.random.Generator, exponent: float = 2.3):
    """Chung–Lu graph with expected degrees ∝ (i+1)^(-1/(exponent-1)) (simple, undirected)."""
    w = (np.arange(n, dtype=np.float64) + 1.0) ** (-1.0 / (exponent - 1.0))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    k = int(m * 1.15) + 16
    # inverse-CDF sampling; torch.searchsorted (same 'left' convention as numpy's, same float64 inputs → same graph) runs
    # multi-threaded on the CPU and on the GPU when there is one: numpy needs 10 s per million-node snapshot here
    dev = _work_device()
    cdf_t = torch.as_tensor(cdf, device=dev)
    u = torch.searchsorted(cdf_t, torch.as_tensor(rng.random(k), device=dev)).cpu().numpy()
    v = torch.searchsorted(cdf_t, torch.as_tensor(rng.random(k), device=dev)).cpu().numpy()
    perm = rng.permutation(n)  # hubs are not the low node ids
    return _simple(n, perm[u], perm[v], m, rng)


def _simple(n, u, v, m, rng):
    dev = _work_device()
    ut, vt = torch.as_tensor(u, device=dev), torch.as_tensor(v, device=dev)
    keep = ut != vt
    ut, vt = ut[keep], vt[keep]
    key = torch.unique(torch.minimum(ut, vt) * n + torch.maximum(ut, vt))
    if key.shape[0] > m:  # drop a random surplus (the key order is by node id, so do not truncate)
        drop = torch.as_tensor(rng.choice(key.shape[0], size=key.shape[0] - m, replace=False), device=dev)
        mask = torch.ones(key.shape[0], dtype=torch.bool, device=dev)
        mask[drop] = False
        key = key[mask]
    key = key.cpu().numpy()
    return key // n, key % n


@dataclass
class SnapshotGraph:
    n: int
    k: int
    rowptr: np.ndarray   # int32 [n+1]
    col: np.ndarray      # int32 [entries]
    val: np.ndarray      # float32 [entries]
    level: np.ndarray    # uint8 [entries]  (bit 7 = one-shot: the +I diagonal)
    nnz_per_core: list   # stored nnz of A_0 … A_{K-1} (identity included in A_0)
    core_levels: list    # the k-core index of every list entry, densest first

    @property
    def entries(self) -> int:
        return int(self.rowptr[-1])

    @property
    def edges_aggregated(self) -> int:
        """Σ_i nnz(A_i): what K torch.sparse.mm calls of the reference consume for one CoreDiffusion layer."""
        return int(sum(self.nnz_per_core))

    def plan(self, device):
        return build_plan_csr(self.n, self.n, self.k, self.rowptr, self.col, self.val, self.level, self.edges_aggregated,
                              device)

    def coo_list(self, device="cpu"):
        """The K uncoalesced torch sparse COO matrices the reference's loader would build."""
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.rowptr))
        lev = self.level & 127
        one = (self.level & 128) != 0
        out = []
        for i in range(self.k):
            sel = np.where(one, lev == i, lev <= i)
            idx = torch.from_numpy(np.vstack((rows[sel], self.col[sel].astype(np.int64))))
            out.append(torch.sparse_coo_tensor(idx, torch.from_numpy(self.val[sel]), (self.n, self.n)).to(device))
        return out


def assemble_snapshot(n: int, u, v, w, le, levels) -> SnapshotGraph:
    """Level-tagged union CSR of one snapshot from its undirected edges: edge e belongs to list entries le[e] … K-1
    (nested k-cores), plus the +I diagonal that helper.py:72 adds to the FIRST entry only (one-shot at level 0)."""
    k_eff = int(len(levels))
    dev = _work_device()
    ut, vt = torch.as_tensor(u, dtype=torch.int64, device=dev), torch.as_tensor(v, dtype=torch.int64, device=dev)
    lt, wt = torch.as_tensor(le, dtype=torch.uint8, device=dev), torch.as_tensor(w, dtype=torch.float32, device=dev)
    diag = torch.arange(n, dtype=torch.int64, device=dev)
    rows = torch.cat([ut, vt, diag])
    cols = torch.cat([vt, ut, diag])
    lvl = torch.cat([lt, lt, torch.full((n,), 128, dtype=torch.uint8, device=dev)])  # diagonal: one-shot at level 0
    vals = torch.cat([wt, wt, torch.ones(n, dtype=torch.float32, device=dev)])
    bits = max(1, int(n - 1).bit_length())
    key = (((rows << 7) | (lvl & 127).to(torch.int64)) << bits) | cols               # (row, level, col)
    order = torch.argsort(key)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0).cpu().numpy()
    per_level = np.bincount(np.asarray(le, dtype=np.int64), minlength=k_eff) * 2
    nnz = np.cumsum(per_level).tolist()
    nnz[0] += n
    return SnapshotGraph(n, k_eff, rowptr.astype(np.int32), cols[order].to(torch.int32).cpu().numpy(),
                         vals[order].cpu().numpy(), lvl[order].cpu().numpy(), [int(x) for x in nnz],
                         [int(x) for x in levels])


def snapshot_from_edges(n: int, u: np.ndarray, v: np.ndarray, k: int, weights: np.ndarray = None) -> SnapshotGraph:
    core = core_numbers(n, u, v)
    ce = np.minimum(core[u], core[v])                      # the largest k-core that still contains the edge
    distinct = np.unique(ce)[::-1]                         # densest (highest) level first
    levels = distinct[:k]
    k_eff = int(levels.shape[0])
    if k_eff == 0:
        raise ValueError("graph has no edges")
    lut = np.full(int(distinct.max()) + 1, 255, dtype=np.uint8)
    lut[levels] = np.arange(k_eff, dtype=np.uint8)
    le = lut[ce]
    keep = le != 255
    u, v, le = u[keep], v[keep], le[keep]
    w = np.ones(u.shape[0], dtype=np.float32) if weights is None else weights[keep].astype(np.float32)
    return assemble_snapshot(n, u, v, w, le, levels)


def make_snapshot(kind: str, n: int, m: int, k: int, seed: int, levels: str = "top") -> SnapshotGraph:
    """levels='top': the K highest distinct core levels (for ER graphs, whose low cores coincide).
    levels='loader': what helper.py:51-82 yields with max_core = K — the files of cores 1..K, densest first, i.e. every edge of
    the graph (a power-law graph's top-K levels hold only its dense nucleus)."""
    rng = np.random.default_rng(seed)
    if kind == "er":
        u, v = er_edges(n, m, rng)
    elif kind == "powerlaw":
        u, v = powerlaw_edges(n, m, rng)
    else:
        raise ValueError(kind)
    if levels == "loader":
        from . import io
        return io.snapshot_from_graph(n, u, v, None, max_core=k)[0]
    if levels != "top":
        raise ValueError(levels)
    return snapshot_from_edges(n, u, v, k)


def features(n: int, d: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g, dtype=torch.float32)
