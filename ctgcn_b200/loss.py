"""Drop-in ``metrics.NegativeSamplingLoss`` with device-side sampling and a fused loss.

Reference: metrics.py:18-93.  Same constructor — ``NegativeSamplingLoss(node_pair_list, neg_freq_list, neg_num=20, Q=10)`` with
the loader's per-snapshot walk co-occurrence lists (helper.py:85-94: ``lil_matrix.rows``) and frequency-expanded negative lists
(helper.py:97-106) — and same call: ``loss_model([embeddings, batch_indices])`` where ``embeddings`` is the model output
([T,N,D] tensor, list of [N,D], or one [N,D]) and ``batch_indices`` an int64 tensor of node ids; returns a tensor of shape [1]
that supports ``.backward()`` (embedding.py:347-349).

What changes: the reference draws the samples in a Python loop over the batch with ``random.sample`` per node and builds index
tensors on the host every batch (metrics.py:68-93); here the lists live on the device as CSR arrays (uploaded once) and one
kernel draws for the whole batch (ctgcn_neg_sample); the loss and its gradient are gather kernels that never materialise the
[S, D] operands (ctgcn_neg_loss_fwd / _bwd).  The draws follow the same distribution (all neighbours when there are at most
``neg_num``, else a uniform ``neg_num``-subset; ``neg_num`` distinct positions of the frequency list) but not the same stream
of random numbers — the reference reseeds from the OS on every call, so its draws are not reproducible either.  Set
``loss_model.seed`` for reproducible draws (call k uses seed + k).  No CPU path: embeddings must be CUDA tensors.
"""
from __future__ import annotations

import itertools
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops


def _csr_from_rows(rows):
    """lil_matrix.rows (object array / list of per-node lists) or a scipy sparse matrix → (ptr int64 [N+1], idx int32)."""
    if hasattr(rows, "tocsr"):
        m = rows.tocsr()
        m.sort_indices()
        return np.asarray(m.indptr, dtype=np.int64), np.asarray(m.indices, dtype=np.int32)
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    idx = np.fromiter(itertools.chain.from_iterable(rows), dtype=np.int32, count=int(ptr[-1]))
    return ptr, idx


class _NegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, batch, pos, count, neg, q):
        loss, ws = ops.neg_loss_fwd(emb.detach(), batch, pos, count, neg, q)
        ctx.save_for_backward(emb, batch, pos, count, neg, ws)
        ctx.q = q
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        emb, batch, pos, count, neg, ws = ctx.saved_tensors
        return ops.neg_loss_bwd(emb.detach(), batch, pos, count, neg, ctx.q, grad_loss, ws), None, None, None, None, None


class NegativeSamplingLoss(nn.Module):
    """Unsupervised negative-sampling loss — reference metrics.py:18-93 (see the module docstring)."""

    def __init__(self, node_pair_list, neg_freq_list, neg_num=20, Q=10):
        super().__init__()
        if not 1 <= int(neg_num) <= _lib.MAX_NEG:
            raise ValueError(f"neg_num must be in [1, {_lib.MAX_NEG}]")
        self.node_pair_list = node_pair_list
        self.neg_freq_list = neg_freq_list
        self.neg_sample_num = int(neg_num)
        self.Q = Q
        self.seed = None            # int → reproducible draws (call k uses seed + k)
        self._calls = 0
        self._host = [None] * len(node_pair_list)     # per snapshot (ptr, idx, freq) as numpy, built on first use
        self._dev = {}                                # (snapshot, device) → CUDA tensors
        self.validate = True        # check index ranges before the kernels run (one device reduction + sync per call for the batch)

    def _host_arrays(self, i):
        if self._host[i] is None:
            ptr, idx = _csr_from_rows(self.node_pair_list[i])
            self._host[i] = (ptr, idx, np.asarray(self.neg_freq_list[i], dtype=np.int32))
        return self._host[i]

    def _check_ranges(self, i, n_rows_emb, batch):
        """The kernels index the embedding with co-occurrence, negative-list and batch node ids without bounds checks; the reference
        raises IndexError on any of them (metrics.py:56-57, 78-93).  Same here, before anything is launched."""
        ptr, idx, freq = self._host_arrays(i)
        hi = max(int(idx.max()) if idx.size else -1, int(freq.max()) if freq.size else -1)
        lo = min(int(idx.min()) if idx.size else 0, int(freq.min()) if freq.size else 0)
        if hi >= n_rows_emb or lo < 0:
            raise IndexError(f"snapshot {i}: node ids of the co-occurrence / negative lists span [{lo}, {hi}], the embedding has {n_rows_emb} rows")
        if self.validate and batch.numel():
            bmin, bmax = int(batch.min()), int(batch.max())
            if bmin < 0 or bmax >= n_rows_emb or bmax >= ptr.shape[0] - 1:
                raise IndexError(f"snapshot {i}: batch node ids span [{bmin}, {bmax}]; the embedding has {n_rows_emb} rows, "
                                 f"the co-occurrence list {ptr.shape[0] - 1}")

    def _arrays(self, i, device):
        key = (i, str(device))
        if key not in self._dev:
            self._dev[key] = tuple(torch.from_numpy(a).to(device) for a in self._host_arrays(i))
        return self._dev[key]

    def sample(self, i, batch_indices, seed):
        """The draw for snapshot i (metrics.py:68-93): (pos int32 [B, neg_num] padded with -1, count int32 [B], neg int32 [neg_num])."""
        ptr, idx, freq = self._arrays(i, batch_indices.device)
        return ops.neg_sample(ptr, idx, freq, batch_indices, self.neg_sample_num, seed)

    def forward(self, input_list):
        assert len(input_list) == 2
        node_embedding, batch_indices = input_list[0], input_list[1]
        if not isinstance(node_embedding, list) and node_embedding.dim() == 2:      # metrics.py:34
            node_embedding = [node_embedding]
        batch = batch_indices.to(torch.int64).contiguous()      # ops.neg_sample rejects CPU tensors: there is no CPU path
        base = (self.seed + self._calls) if self.seed is not None else random.getrandbits(62)
        self._calls += 1
        total = torch.zeros(1, dtype=torch.float32, device=batch.device)            # metrics.py:40
        for i in range(len(node_embedding)):
            emb = node_embedding[i]
            self._check_ranges(i, emb.shape[0], batch)
            if emb.dtype != torch.float32 or emb.stride(-1) != 1:
                emb = emb.float().contiguous()
            pos, count, neg = self.sample(i, batch, base * 1_000_003 + i)
            total = total + _NegLossFn.apply(emb, batch, pos, count, neg, float(self.Q))
        return total
