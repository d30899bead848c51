"""Tensor-level wrappers over the C-ABI (torch tensors in, torch tensors out; torch only owns memory/streams)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .plan import GraphPlan

_vp = C.c_void_p


def _ptr(t):
    return _vp(0) if t is None else _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _f32_rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D fp32 CUDA tensor whose last dim is contiguous (row stride free)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.CtgcnError(f"{name} must be a CUDA tensor (ctgcn_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        raise _lib.CtgcnError(f"{name} must be 2-D, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _vec(t, name):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise _lib.CtgcnError(f"{name} must be an fp32 CUDA tensor")
    return t.detach().contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def cumspmm(plan: GraphPlan, x: torch.Tensor, relu: bool = True, out: torch.Tensor = None) -> torch.Tensor:
    """relu(cumsum_i A_i x) for all K cores → [N, K, D]  (layers.py:41-48); relu=False returns the sums themselves.
    `out`: optional contiguous fp32 buffer with at least N·K·D elements (its first N·K·D elements are used)."""
    x = _f32_rows(x, "x")
    if x.shape[0] != plan.n_cols:
        raise _lib.CtgcnError(f"x has {x.shape[0]} rows, the plan expects {plan.n_cols}")
    if out is None:
        u = torch.empty(plan.n_rows, plan.k, x.shape[1], dtype=torch.float32, device=x.device)
    else:
        need = plan.n_rows * plan.k * x.shape[1]
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < need or out.device != x.device:
            raise _lib.CtgcnError("out must be a contiguous fp32 buffer on x's device with at least N*K*D elements")
        u = out.view(-1)[:need].view(plan.n_rows, plan.k, x.shape[1])
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ctgcn_cumspmm_fwd_ex(plan.handle, _ptr(x), x.stride(0), x.shape[1], 1 if relu else 0, _ptr(u),
                                                 _stream()), "ctgcn_cumspmm_fwd_ex")
    return u


def cumspmm_bwd(plan_t: GraphPlan, g: torch.Tensor) -> torch.Tensor:
    """dL/dx of the cumulative SpMM: g [M, K, D] = dL/dS_i (relu mask applied) → [N, D], over the plan of the transposed list."""
    if g.dim() != 3 or not g.is_cuda or g.dtype != torch.float32:
        raise _lib.CtgcnError("g must be a 3-D fp32 CUDA tensor [rows, K, D] (ctgcn_b200 has no CPU path)")
    g = g.contiguous()
    if g.shape[0] != plan_t.n_cols or g.shape[1] != plan_t.k:
        raise _lib.CtgcnError(f"g has shape {tuple(g.shape)}, the transposed plan expects [{plan_t.n_cols}, {plan_t.k}, D]")
    d = g.shape[2]
    dx = torch.empty(plan_t.n_rows, d, dtype=torch.float32, device=g.device)
    ws_bytes = _lib.lib.ctgcn_cumspmm_bwd_workspace_bytes(plan_t.handle, d)
    ws = _workspace(ws_bytes, g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib.ctgcn_cumspmm_bwd(plan_t.handle, _ptr(g), d, _ptr(dx), dx.stride(0), _ptr(ws), ws_bytes, _stream()),
                   "ctgcn_cumspmm_bwd")
    return dx


def _rnn_weights(w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, cell, d_in):
    h = w_hh.shape[1]
    g = 4 if cell == _lib.CELL_LSTM else 3
    w_ih, w_hh, b_ih, b_hh, ln_w, ln_b = (_vec(t, nm) for t, nm in
                                           ((w_ih, "w_ih"), (w_hh, "w_hh"), (b_ih, "b_ih"), (b_hh, "b_hh"),
                                            (ln_w, "ln_w"), (ln_b, "ln_b")))
    if tuple(w_ih.shape) != (g * h, d_in) or tuple(w_hh.shape) != (g * h, h):
        raise _lib.CtgcnError(f"recurrent weight shapes {tuple(w_ih.shape)}, {tuple(w_hh.shape)} do not match "
                              f"d_in={d_in}, h={h}, gates={g}")
    return h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b


def rnn_seq(seq: torch.Tensor, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps: float, mode: int, out: torch.Tensor = None,
            cell: int = _lib.CELL_GRU):
    """GRU / LSTM over dim 1 of seq [N, L, D_in] (any row/step strides) + LayerNorm.

    mode GRU_SUM_LN → [N, H] = LN(Σ_s h_s); GRU_EACH_LN → [N, L, H] = LN(h_s).
    """
    if seq.dim() != 3 or not seq.is_cuda or seq.dtype != torch.float32 or seq.stride(2) != 1:
        raise _lib.CtgcnError("seq must be a 3-D fp32 CUDA tensor with a contiguous last dim")
    n, steps, d_in = seq.shape
    h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b = _rnn_weights(w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, cell, d_in)
    if out is None:
        out = torch.empty((n, h) if mode == _lib.GRU_SUM_LN else (n, steps, h), dtype=torch.float32, device=seq.device)
    yrs = out.stride(0)
    yss = out.stride(1) if mode == _lib.GRU_EACH_LN else 0
    if out.stride(-1) != 1:
        raise _lib.CtgcnError("out must have a contiguous last dim")
    ws_bytes = _lib.lib.ctgcn_rnn_workspace_bytes(cell, d_in, h)
    ws = _workspace(ws_bytes, seq.device)
    with torch.cuda.device(seq.device):
        _lib.check(_lib.lib.ctgcn_rnn_seq_fwd(cell, _ptr(seq), seq.stride(0), seq.stride(1), n, steps, d_in, h, _ptr(w_ih),
                                              _ptr(w_hh), _ptr(b_ih), _ptr(b_hh), _ptr(ln_w), _ptr(ln_b), float(eps), mode,
                                              _ptr(out), yrs, yss, _ptr(ws), ws_bytes, _stream()), "ctgcn_rnn_seq_fwd")
    return out


def gru_seq(seq, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps: float, mode: int, out: torch.Tensor = None):
    return rnn_seq(seq, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, out=out, cell=_lib.CELL_GRU)


def core_diffusion(plan: GraphPlan, x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps: float, out: torch.Tensor = None,
                   cell: int = _lib.CELL_GRU, scatter=None):
    """layers.CoreDiffusion.forward on one plan → [N, H] (written into `out`, any row stride, if given).

    scatter = (slice_ptrs int64 CUDA tensor of (peer-mapped) base pointers, slice_row_stride, slice_col_offset): the result
    rows are stored straight into the node slices' buffers instead (nothing is returned)."""
    x = _f32_rows(x, "x")
    d_in = x.shape[1]
    if x.shape[0] != plan.n_cols:
        raise _lib.CtgcnError(f"x has {x.shape[0]} rows, the plan expects {plan.n_cols}")
    h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b = _rnn_weights(w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, cell, d_in)
    ws_bytes = _lib.lib.ctgcn_core_diffusion_rnn_workspace_bytes(plan.handle, cell, d_in, h)
    ws = _workspace(ws_bytes, x.device)
    if scatter is not None:
        slice_ptrs, slice_row_stride, slice_col_offset = scatter
        if slice_ptrs.dtype != torch.int64 or not slice_ptrs.is_cuda:
            raise _lib.CtgcnError("slice_ptrs must be an int64 CUDA tensor of device pointers")
        y_ptr, ldy, sl = _vp(0), 0, (_ptr(slice_ptrs), slice_ptrs.numel(), int(slice_row_stride), int(slice_col_offset))
    else:
        if out is None:
            out = torch.empty(plan.n_rows, h, dtype=torch.float32, device=x.device)
        if out.stride(1) != 1 or tuple(out.shape) != (plan.n_rows, h):
            raise _lib.CtgcnError("out must be [N, H] with a contiguous last dim")
        y_ptr, ldy, sl = _ptr(out), out.stride(0), (_vp(0), 0, 0, 0)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ctgcn_core_diffusion_rnn_fwd(plan.handle, cell, _ptr(x), x.stride(0), d_in, h, _ptr(w_ih), _ptr(w_hh),
                                                         _ptr(b_ih), _ptr(b_hh), _ptr(ln_w), _ptr(ln_b), float(eps), y_ptr, ldy,
                                                         *sl, _ptr(ws), ws_bytes, _stream()),
                   "ctgcn_core_diffusion_rnn_fwd")
    return out if scatter is None else None


def core_diffusion_scatter(plan: GraphPlan, x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps: float, slice_ptrs: torch.Tensor,
                           slice_row_stride: int, slice_col_offset: int, cell: int = _lib.CELL_GRU):
    """CoreDiffusion.forward whose [N, H] result is scattered row-wise into the node slices' buffers."""
    return core_diffusion(plan, x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, cell=cell,
                          scatter=(slice_ptrs, slice_row_stride, slice_col_offset))


def linear(x, w, b, act: int):
    """act(x wᵀ + b) for dense x [N, d_in]."""
    x = _f32_rows(x, "x")
    w, b = _vec(w, "weight"), _vec(b, "bias")
    d_out, d_in = w.shape
    if x.shape[1] != d_in:
        raise _lib.CtgcnError(f"x width {x.shape[1]} != weight in_features {d_in}")
    y = torch.empty(x.shape[0], d_out, dtype=torch.float32, device=x.device)
    ws_bytes = _lib.lib.ctgcn_linear_workspace_bytes(d_in, d_out)
    ws = _workspace(ws_bytes, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib.ctgcn_linear_fwd(_ptr(x), x.stride(0), x.shape[0], d_in, _ptr(w), _ptr(b), d_out, act, _ptr(y),
                                             y.stride(0), _ptr(ws), ws_bytes, _stream()), "ctgcn_linear_fwd")
    return y


def spmm_linear(x_plan: GraphPlan, w, b, act: int):
    """act(x wᵀ + b) for a sparse x given as a K=1 plan of shape [N, d_in]."""
    w, b = _vec(w, "weight"), _vec(b, "bias")
    d_out, d_in = w.shape
    if x_plan.n_cols != d_in:
        raise _lib.CtgcnError(f"sparse x width {x_plan.n_cols} != weight in_features {d_in}")
    y = torch.empty(x_plan.n_rows, d_out, dtype=torch.float32, device=w.device)
    ws_bytes = _lib.lib.ctgcn_linear_workspace_bytes(d_in, d_out)
    ws = _workspace(ws_bytes, w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib.ctgcn_spmm_linear_fwd(x_plan.handle, _ptr(w), _ptr(b), d_out, act, _ptr(y), y.stride(0), _ptr(ws),
                                                  ws_bytes, _stream()), "ctgcn_spmm_linear_fwd")
    return y


# ---- negative-sampling loss (reference metrics.py:18-93)
def neg_sample(pair_ptr: torch.Tensor, pair_idx: torch.Tensor, freq: torch.Tensor, batch: torch.Tensor, neg_num: int, seed: int):
    """Device-side NegativeSamplingLoss.__get_node_indices (metrics.py:68-93) for one snapshot.
    pair_ptr int64 [N+1] / pair_idx int32: CSR of the walk co-occurrence lists; freq int32: frequency-expanded negative list;
    batch int64 node ids → (pos int32 [B, neg_num] (-1 padded), count int32 [B], neg int32 [neg_num])."""
    for t, dt, nm in ((pair_ptr, torch.int64, "pair_ptr"), (pair_idx, torch.int32, "pair_idx"), (freq, torch.int32, "freq"),
                      (batch, torch.int64, "batch")):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise _lib.CtgcnError(f"{nm} must be a contiguous {dt} CUDA tensor (ctgcn_b200 has no CPU path)")
    if freq.numel() < neg_num:
        raise ValueError("Sample larger than population or is negative")     # what random.sample raises at metrics.py:88
    nb = batch.numel()
    dev = batch.device
    pos = torch.empty(nb, neg_num, dtype=torch.int32, device=dev)
    count = torch.empty(nb, dtype=torch.int32, device=dev)
    neg = torch.empty(neg_num, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.ctgcn_neg_sample(_ptr(pair_ptr), _ptr(pair_idx), pair_ptr.numel() - 1, _ptr(freq), freq.numel(),
                                             _ptr(batch), nb, int(neg_num), int(seed) & (2 ** 64 - 1), _ptr(pos), _ptr(count),
                                             _ptr(neg), _stream()), "ctgcn_neg_sample")
    return pos, count, neg


def _loss_args(emb, batch, pos, count, neg):
    if emb.dim() != 2 or not emb.is_cuda or emb.dtype != torch.float32 or emb.stride(1) != 1:
        raise _lib.CtgcnError("embeddings must be a 2-D fp32 CUDA tensor with a contiguous last dim (ctgcn_b200 has no CPU path)")
    return (_ptr(emb), emb.stride(0), emb.shape[0], emb.shape[1], _ptr(batch), batch.numel(), _ptr(pos), _ptr(count), _ptr(neg),
            pos.shape[1])


def neg_loss_fwd(emb, batch, pos, count, neg, q: float):
    """(loss [1], workspace) for one snapshot — metrics.py:55-61 on the samples of neg_sample; the workspace goes to neg_loss_bwd."""
    args = _loss_args(emb, batch, pos, count, neg)
    ws_bytes = _lib.lib.ctgcn_neg_loss_workspace_bytes(batch.numel(), emb.shape[1])
    ws = _workspace(ws_bytes, emb.device)
    loss = torch.empty(1, dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.lib.ctgcn_neg_loss_fwd(*args, float(q), _ptr(loss), _ptr(ws), ws_bytes, _stream()), "ctgcn_neg_loss_fwd")
    return loss, ws


def neg_loss_bwd(emb, batch, pos, count, neg, q: float, grad_loss, ws):
    """grad_loss [1] · d loss / d emb → dense [N, D] (atomics; rows outside the samples stay zero)."""
    args = _loss_args(emb, batch, pos, count, neg)
    grad = torch.zeros(emb.shape[0], emb.shape[1], dtype=torch.float32, device=emb.device)
    gl = grad_loss.detach().reshape(-1)[:1].to(torch.float32).contiguous()
    with torch.cuda.device(emb.device):
        _lib.check(_lib.lib.ctgcn_neg_loss_bwd(*args, float(q), _ptr(gl), _ptr(grad), grad.stride(0), _ptr(ws), ws.numel(),
                                               _stream()), "ctgcn_neg_loss_bwd")
    return grad
