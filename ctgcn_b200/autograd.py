"""Backward of the hot path so that the reference's trainer (embedding.py:347-352:
``loss.backward()`` + Adam) can drive the drop-in modules.

Forward = the fused sm_100a kernels (nothing but the layer inputs is kept).  Backward recomputes the layer:

  * sparse parts — hand-written kernels over the graph plan: ``ctgcn_cumspmm_fwd`` (recompute of the per-core sums),
    ``ctgcn_cumspmm_bwd`` (one gathered row per stored entry of the transposed union CSR), ``ctgcn_cumspmm_fwd_ex`` (weight
    gradient of a sparse-input Linear);
  * dense contractions (dW, dX of the gate matmuls and the Linear layers) — plain library GEMMs (cuBLAS through
    ``torch.matmul``), the gate / LayerNorm / selu derivatives are torch elementwise ops.

The recurrence derivatives (``rnn_seq_bwd``) are device-agnostic torch code: tests check them on CPU against autograd of
``nn.GRU`` / ``nn.LSTM`` + ``LayerNorm``.  The Functions themselves need CUDA tensors like every other entry point.

Reference arithmetic differentiated here: layers.py:38-63 (CoreDiffusion), layers.py:95-106 (MLP), models.py:248-250 (temporal
GRU/LSTM + LayerNorm); PyTorch cell equations (gate packing GRU [r;z;n], LSTM [i;f;g;o]).
"""
from __future__ import annotations

import torch

from . import _lib, ops

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


# ----------------------------------------------------------------------------- device-agnostic derivative code
def layer_norm_bwd(dy, v, weight, eps):
    """Backward of LayerNorm over the last axis.  Returns (dv, dweight, dbias); v is the pre-norm input."""
    mu = v.mean(dim=-1, keepdim=True)
    var = ((v - mu) ** 2).mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    xhat = (v - mu) * rstd
    red = tuple(range(dy.dim() - 1))
    dw = (dy * xhat).sum(dim=red)
    db = dy.sum(dim=red)
    dxhat = dy * weight
    dv = rstd * (dxhat - dxhat.mean(dim=-1, keepdim=True) - xhat * (dxhat * xhat).mean(dim=-1, keepdim=True))
    return dv, dw, db


def rnn_recompute(seq, cell, w_ih, w_hh, b_ih, b_hh):
    """Forward of the single-layer batch_first GRU / LSTM with zero initial state, keeping what the backward needs.
    seq [N, L, D] → dict with h [N, L, H] and the per-step gate activations."""
    n, L, d = seq.shape
    H = w_hh.shape[1]
    G = 4 if cell == _lib.CELL_LSTM else 3
    gi = (seq.reshape(n * L, d) @ w_ih.t()).view(n, L, G * H)
    if b_ih is not None:
        gi = gi + b_ih
    h = seq.new_zeros(n, H)
    keep = {k: [] for k in (("i", "f", "g", "o", "c", "h") if cell == _lib.CELL_LSTM else ("r", "z", "n", "ghn", "h"))}
    c = seq.new_zeros(n, H)
    for s in range(L):
        gh = h @ w_hh.t()
        if b_hh is not None:
            gh = gh + b_hh
        if cell == _lib.CELL_LSTM:
            pre = gi[:, s] + gh
            i, f = torch.sigmoid(pre[:, :H]), torch.sigmoid(pre[:, H:2 * H])
            g, o = torch.tanh(pre[:, 2 * H:3 * H]), torch.sigmoid(pre[:, 3 * H:])
            c = f * c + i * g
            h = o * torch.tanh(c)
            for k, v in (("i", i), ("f", f), ("g", g), ("o", o), ("c", c), ("h", h)):
                keep[k].append(v)
        else:
            r = torch.sigmoid(gi[:, s, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, s, H:2 * H] + gh[:, H:2 * H])
            ghn = gh[:, 2 * H:]
            nn_ = torch.tanh(gi[:, s, 2 * H:] + r * ghn)
            h = (1.0 - z) * nn_ + z * h
            for k, v in (("r", r), ("z", z), ("n", nn_), ("ghn", ghn), ("h", h)):
                keep[k].append(v)
    return {k: torch.stack(v, dim=1) for k, v in keep.items()}


def rnn_seq_bwd(seq, cell, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, dy):
    """Backward of  y = LN(Σ_s h_s) (mode SUM_LN, dy [N, H])  or  y_s = LN(h_s) (mode EACH_LN, dy [N, L, H]).
    Returns (dseq, dw_ih, dw_hh, db_ih, db_hh, dln_w, dln_b); bias gradients are None when the cell has no bias."""
    n, L, d = seq.shape
    H = w_hh.shape[1]
    G = 4 if cell == _lib.CELL_LSTM else 3
    k = rnn_recompute(seq, cell, w_ih, w_hh, b_ih, b_hh)
    hs = k["h"]
    if mode == _lib.GRU_SUM_LN:
        d_o, dln_w, dln_b = layer_norm_bwd(dy, hs.sum(dim=1), ln_w, eps)
        dh_ext = d_o.unsqueeze(1).expand(n, L, H)
    else:
        dh_ext, dln_w, dln_b = layer_norm_bwd(dy, hs, ln_w, eps)
    dgi = seq.new_empty(n, L, G * H)
    dw_hh = torch.zeros_like(w_hh)
    db_hh = None if b_hh is None else torch.zeros_like(b_hh)
    dh = seq.new_zeros(n, H)
    dc = seq.new_zeros(n, H)
    zeros = seq.new_zeros(n, H)
    for s in range(L - 1, -1, -1):
        dh = dh + dh_ext[:, s]
        h_prev = hs[:, s - 1] if s > 0 else zeros
        if cell == _lib.CELL_LSTM:
            i, f, g, o, c = k["i"][:, s], k["f"][:, s], k["g"][:, s], k["o"][:, s], k["c"][:, s]
            c_prev = k["c"][:, s - 1] if s > 0 else zeros
            tc = torch.tanh(c)
            dc = dc + dh * o * (1.0 - tc * tc)
            dpre = torch.cat([dc * g * i * (1.0 - i), dc * c_prev * f * (1.0 - f), dc * i * (1.0 - g * g),
                              dh * tc * o * (1.0 - o)], dim=1)
            dc = dc * f
            dgh = dpre
            dh = dgh @ w_hh
        else:
            r, z, nn_, ghn = k["r"][:, s], k["z"][:, s], k["n"][:, s], k["ghn"][:, s]
            dpre_n = dh * (1.0 - z) * (1.0 - nn_ * nn_)
            dpre_z = dh * (h_prev - nn_) * z * (1.0 - z)
            dpre_r = dpre_n * ghn * r * (1.0 - r)
            dpre = torch.cat([dpre_r, dpre_z, dpre_n], dim=1)
            dgh = torch.cat([dpre_r, dpre_z, dpre_n * r], dim=1)
            dh = dh * z + dgh @ w_hh
        dgi[:, s] = dpre
        if s > 0:                                  # h_{-1} = 0: no contribution to dW_hh
            dw_hh += dgh.t() @ h_prev
        if db_hh is not None:
            db_hh += dgh.sum(dim=0)
    flat = dgi.reshape(n * L, G * H)
    dseq = (flat @ w_ih).view(n, L, d)
    dw_ih = flat.t() @ seq.reshape(n * L, d)
    db_ih = None if b_ih is None else flat.sum(dim=0)
    return dseq, dw_ih, dw_hh, db_ih, db_hh, dln_w, dln_b


# Bound on the recompute's intermediates (gate activations and their gradients of one row chunk), in bytes.  The forward honours
# ctgcn_set_workspace_cap; without a bound here the backward of a bench-size layer (1 M rows × K = 10 × 128) kept > 60 GB of
# [N, K, 3H] / [N, K, H] tensors alive (round-1 advice).  Rows are independent in the recurrence and in LayerNorm, so the backward
# runs chunk by chunk: dseq rows are written in place, weight / bias / LayerNorm gradients are summed over the chunks.
BWD_CHUNK_BYTES = 2 << 30


def set_backward_chunk_bytes(nbytes: int) -> None:
    global BWD_CHUNK_BYTES
    BWD_CHUNK_BYTES = max(int(nbytes), 1 << 20)


def _bwd_chunk_rows(L: int, d: int, H: int, G: int) -> int:
    per_row = 4 * L * (2 * G * H + 6 * H + d)          # gi, dgi, the kept gate tensors and h, dseq
    return max(256, BWD_CHUNK_BYTES // per_row)


def rnn_seq_bwd_chunked(seq, cell, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, dy, max_rows=None):
    """rnn_seq_bwd over row chunks of at most `max_rows` rows (default: from BWD_CHUNK_BYTES).  Same values as the unchunked call up
    to the summation order of the parameter gradients; one chunk = exactly the unchunked call."""
    n, L, d = seq.shape
    H = w_hh.shape[1]
    G = 4 if cell == _lib.CELL_LSTM else 3
    rows = _bwd_chunk_rows(L, d, H, G) if max_rows is None else int(max_rows)
    if n <= rows:
        return rnn_seq_bwd(seq, cell, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, dy)
    dseq = torch.empty_like(seq)
    acc = None
    for r0 in range(0, n, rows):
        sl = slice(r0, min(n, r0 + rows))
        out = rnn_seq_bwd(seq[sl], cell, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, dy[sl])
        dseq[sl] = out[0]
        if acc is None:
            acc = [None if g is None else g.clone() for g in out[1:]]
        else:
            for a, g in zip(acc, out[1:]):
                if a is not None:
                    a += g
    return (dseq, *acc)


def selu_bwd_from_output(dy, y):
    """d selu / d pre-activation expressed through the OUTPUT y = selu(v): scale for v > 0, y + scale·alpha otherwise."""
    return dy * torch.where(y > 0, torch.full_like(y, SELU_SCALE), y + SELU_SCALE * SELU_ALPHA)


# ----------------------------------------------------------------------------- autograd Functions (CUDA only)
class CoreDiffusionFn(torch.autograd.Function):
    """layers.CoreDiffusion.forward (layers.py:38-63) with a backward."""

    @staticmethod
    def forward(ctx, x, plan, cell, eps, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b):
        y = ops.core_diffusion(plan, x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, cell=cell)
        ctx.plan, ctx.cell, ctx.eps = plan, cell, eps
        ctx.save_for_backward(x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b = ctx.saved_tensors
        plan = ctx.plan
        with torch.no_grad():
            u = ops.cumspmm(plan, x)                                        # [N, K, D] = relu(S_i), recomputed
            du, dw_ih, dw_hh, db_ih, db_hh, dln_w, dln_b = rnn_seq_bwd_chunked(u, ctx.cell, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b,
                                                                               ctx.eps, _lib.GRU_SUM_LN, dy.contiguous())
            dx = None
            if ctx.needs_input_grad[0]:
                du.mul_(u > 0)                                              # relu mask (in place: du is ours)
                del u
                dx = ops.cumspmm_bwd(plan.transposed(), du)                 # Σ_j A_jᵀ Σ_{i≥j} dS_i
        return dx, None, None, None, dw_ih, dw_hh, db_ih, db_hh, dln_w, dln_b


class RnnSeqFn(torch.autograd.Function):
    """Temporal GRU / LSTM + per-step LayerNorm (models.py:249-250) with a backward.  seq [N, T, D] → [N, T, H]."""

    @staticmethod
    def forward(ctx, seq, cell, eps, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b):
        y = ops.rnn_seq(seq, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, _lib.GRU_EACH_LN, cell=cell)
        ctx.cell, ctx.eps = cell, eps
        ctx.save_for_backward(seq, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b)
        return y

    @staticmethod
    def backward(ctx, dy):
        seq, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b = ctx.saved_tensors
        with torch.no_grad():
            dseq, dw_ih, dw_hh, db_ih, db_hh, dln_w, dln_b = rnn_seq_bwd_chunked(seq, ctx.cell, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b,
                                                                                 ctx.eps, _lib.GRU_EACH_LN, dy.contiguous())
        return dseq, None, None, dw_ih, dw_hh, db_ih, db_hh, dln_w, dln_b


class LinearFn(torch.autograd.Function):
    """act(x Wᵀ + b) for dense x (layers.py:97-105) with a backward."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        y = ops.linear(x, w, b, act)
        ctx.act = act
        ctx.save_for_backward(x, w, y if act == _lib.ACT_SELU else None)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        with torch.no_grad():
            dpre = selu_bwd_from_output(dy, y) if ctx.act == _lib.ACT_SELU else dy
            dx = dpre @ w if ctx.needs_input_grad[0] else None
            dw = dpre.t() @ x.to(torch.float32)
            db = dpre.sum(dim=0) if ctx.has_bias else None
        return dx, dw, db, None


class SparseLinearFn(torch.autograd.Function):
    """act(X Wᵀ + b) for a sparse X given as a K = 1 plan (layers.py:97 with a COO input).  X is data: no dX."""

    @staticmethod
    def forward(ctx, w, b, plan, act):
        y = ops.spmm_linear(plan, w, b, act)
        ctx.plan, ctx.act, ctx.has_bias = plan, act, b is not None
        ctx.save_for_backward(y if act == _lib.ACT_SELU else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        with torch.no_grad():
            dpre = (selu_bwd_from_output(dy, y) if ctx.act == _lib.ACT_SELU else dy).contiguous()
            dwt = ops.cumspmm(ctx.plan.transposed(), dpre, relu=False)      # Xᵀ·dpre  → [d_in, 1, d_out]
            dw = dwt[:, 0, :].t()
            db = dpre.sum(dim=0) if ctx.has_bias else None
        return dw, db, None, None


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)
