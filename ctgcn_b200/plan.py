"""Graph plans: the reference's per-snapshot ``adj_list`` (K torch sparse COO matrices, helper.py:51-82)
merged once into a level-tagged union CSR on the device (ctgcn_plan_create_coo).

The reference hands the SAME ``adj_list`` objects to ``model(x_list, adj_list)`` on every batch of every
epoch (embedding.py:346), so plans are cached per list identity and validated by the storage pointers of
their index tensors.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

_CACHE_SIZE = 256


class GraphPlan:
    """Owner of one opaque ``ctgcn_plan*``."""

    def __init__(self, handle: int, device: torch.device):
        self._h = C.c_void_p(handle)
        self.device = device
        st = (C.c_int64 * 8)()
        _lib.check(_lib.lib.ctgcn_plan_stats(self._h, st), "ctgcn_plan_stats")
        (self.n_rows, self.n_cols, self.k, self.entries, self.nnz_raw_sum, self.nnz_coalesced, self.n_oneshot,
         self.device_bytes) = [int(v) for v in st]

    @property
    def handle(self):
        if self._h is None:
            raise _lib.CtgcnError("plan already destroyed")
        return self._h

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib.ctgcn_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def transposed(self) -> "GraphPlan":
        """Plan of the transposed list [A_0ᵀ … A_{K-1}ᵀ] (built once, on first use: the backward pass gathers over it)."""
        if getattr(self, "_t", None) is None:
            rowptr, col, val, lvl = self.arrays()
            self._t = build_plan_csr(self.n_cols, self.n_rows, self.k, *transpose_csr(self.n_rows, self.n_cols, rowptr, col, val, lvl),
                                     self.nnz_raw_sum, self.device)
            self._t._t = self
        return self._t

    def arrays(self):
        """Copies of (rowptr, col, val, level) as torch tensors (tests / inspection only)."""
        out = [torch.empty(cnt, dtype=dt, device=self.device)
               for dt, cnt in zip((torch.int32, torch.int32, torch.float32, torch.uint8),
                                  (self.n_rows + 1, self.entries, self.entries, self.entries))]
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib.ctgcn_plan_arrays(self.handle, *[C.c_void_p(t.data_ptr()) for t in out],
                                                  C.c_void_p(stream)), "ctgcn_plan_arrays")
        return out


def transpose_csr(n_rows, n_cols, rowptr, col, val, lvl):
    """Level-tagged union CSR of the transposed matrices: entry (r, c, w, level) → (c, r, w, level), rows sorted by level.
    torch tensors in, torch tensors out (any device): a nested entry of A_i, i ≥ ℓ is a nested entry of A_iᵀ, i ≥ ℓ."""
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n_rows, device=rowptr.device, dtype=torch.int64), counts)
    c64 = col.to(torch.int64)
    key = c64 * 128 + (lvl & 127).to(torch.int64)
    order = torch.argsort(key, stable=True)
    t_rowptr = torch.zeros(n_cols + 1, dtype=torch.int64, device=rowptr.device)
    t_rowptr[1:] = torch.cumsum(torch.bincount(c64, minlength=n_cols), 0)
    return t_rowptr.to(torch.int32), rows[order].to(torch.int32), val[order], lvl[order]


def _coo_parts(m, device):
    """(rows int64, cols int64, vals fp32, shape) of one matrix as contiguous device tensors."""
    if isinstance(m, torch.Tensor):
        if m.layout != torch.sparse_coo:
            raise TypeError("adjacency / sparse feature matrices must be torch sparse COO tensors")
        idx, val = m._indices(), m._values()
        shape = tuple(m.shape)
    else:  # scipy sparse
        mc = m.tocoo()
        idx = torch.from_numpy(np.vstack((mc.row, mc.col)).astype(np.int64))
        val = torch.from_numpy(mc.data.astype(np.float32))
        shape = mc.shape
    idx = idx.to(device=device, dtype=torch.int64)
    val = val.to(device=device, dtype=torch.float32).contiguous()
    return idx[0].contiguous(), idx[1].contiguous(), val, shape


def build_plan_coo(mats, device) -> GraphPlan:
    """ctgcn_plan_create_coo over a list of K COO matrices of one common shape."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.CtgcnError("graph plans live on a CUDA device; got " + str(device))
    k = len(mats)
    if not 1 <= k <= _lib.MAX_CORES:
        raise _lib.CtgcnError(f"adj_list must hold 1..{_lib.MAX_CORES} matrices, got {k}")
    parts = [_coo_parts(m, device) for m in mats]
    shape = parts[0][3]
    for p in parts:
        if tuple(p[3]) != tuple(shape):
            raise _lib.CtgcnError("all matrices of one adj_list must share one shape")
    rows = (C.c_void_p * k)(*[p[0].data_ptr() for p in parts])
    cols = (C.c_void_p * k)(*[p[1].data_ptr() for p in parts])
    vals = (C.c_void_p * k)(*[p[2].data_ptr() for p in parts])
    nnz = (C.c_int64 * k)(*[p[2].numel() for p in parts])
    out = C.c_void_p()
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib.ctgcn_plan_create_coo(shape[0], shape[1], k, rows, cols, vals, nnz, 1, C.c_void_p(stream),
                                                  C.byref(out)), "ctgcn_plan_create_coo")
    return GraphPlan(out.value, device)


def build_plan_csr(n_rows, n_cols, k, rowptr, col, val, level, nnz_raw_sum, device) -> GraphPlan:
    """ctgcn_plan_create_csr from host (numpy) or device (torch) arrays of an already merged union CSR."""
    device = torch.device(device)
    on_device = isinstance(rowptr, torch.Tensor) and rowptr.is_cuda
    if on_device:
        arrs = [rowptr.to(torch.int32).contiguous(), col.to(torch.int32).contiguous(),
                val.to(torch.float32).contiguous(), level.to(torch.uint8).contiguous()]
        ptrs = [a.data_ptr() for a in arrs]
    else:
        arrs = [np.ascontiguousarray(rowptr, dtype=np.int32), np.ascontiguousarray(col, dtype=np.int32),
                np.ascontiguousarray(val, dtype=np.float32), np.ascontiguousarray(level, dtype=np.uint8)]
        ptrs = [a.ctypes.data for a in arrs]
    out = C.c_void_p()
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib.ctgcn_plan_create_csr(n_rows, n_cols, k, *[C.c_void_p(p) for p in ptrs], int(nnz_raw_sum),
                                                  1 if on_device else 0, C.c_void_p(stream), C.byref(out)),
                   "ctgcn_plan_create_csr")
    return GraphPlan(out.value, device)


_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
_cache_limits = {"entries": _CACHE_SIZE, "bytes": 16 << 30}


def set_cache_limits(entries: int = None, device_bytes: int = None):
    """Bounds of the plan cache: number of cached adj_lists and total cudaMalloc'd plan bytes (least recently used plans go
    first).  Plans live outside torch's caching allocator, so ``torch.cuda.empty_cache()`` never reclaims them."""
    if entries is not None:
        _cache_limits["entries"] = max(int(entries), 1)
    if device_bytes is not None:
        _cache_limits["bytes"] = max(int(device_bytes), 0)
    _trim()


def _plan_bytes(plan):
    t = getattr(plan, "_t", None)
    return plan.device_bytes + (t.device_bytes if t is not None else 0)


def _trim(keep=None):
    while len(_cache) > 1 and (len(_cache) > _cache_limits["entries"] or
                               sum(_plan_bytes(v[1]) for v in _cache.values()) > _cache_limits["bytes"]):
        key = next(iter(_cache))
        if key == keep:
            break
        _cache.pop(key)


def _signature(mats):
    sig = []
    for m in mats:
        if isinstance(m, torch.Tensor):
            sig.append((m._indices().data_ptr(), m._values().data_ptr(), m._nnz(), m._values()._version))
        else:
            sig.append((id(m), m.nnz))
    return tuple(sig)


def _evict(key):
    _cache.pop(key, None)


def plan_for(adj_list, device) -> GraphPlan:
    """Cached plan of one snapshot's adj_list (or of a single sparse feature matrix wrapped in a list).

    The cache never keeps the caller's matrices alive: it holds weak references only, and the entry (with its device plan) is
    dropped as soon as one of the matrices is garbage-collected — the reference trainer frees a time window's graphs with
    ``del adj_list, x_list`` (embedding.py:287, 365; window loop at train.py:269) and expects the memory back."""
    if isinstance(adj_list, GraphPlan):
        return adj_list
    key = (id(adj_list), str(device))
    if isinstance(adj_list, torch.Tensor) or hasattr(adj_list, "tocoo"):
        mats = [adj_list]  # a single sparse matrix (MLP input features)
    else:
        mats = list(adj_list)
    sig = _signature(mats)
    hit = _cache.get(key)
    if hit is not None and hit[0] == sig and all(r() is m for r, m in zip(hit[2], mats)):
        _cache.move_to_end(key)
        return hit[1]
    plan = build_plan_coo(mats, device)
    refs = []
    for m in mats:
        refs.append(weakref.ref(m))
        weakref.finalize(m, _evict, key)      # any matrix of the list dying invalidates the entry
    _cache[key] = (sig, plan, refs)
    _trim(keep=key)
    return plan


def clear_cache():
    """Destroy every cached plan now (window-based training: call between windows to return the device memory at once)."""
    _cache.clear()
