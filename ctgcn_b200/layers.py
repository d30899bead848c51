"""Drop-in ``layers`` module: ``CoreDiffusion`` and ``MLP`` with the reference's constructor / forward
signatures and ``state_dict`` keys (reference layers.py:9-63, 67-106), computing on sm_100a through
libctgcn_b200.so.  There is no torch fallback: tensors must live on a CUDA device.

Training: when gradients are required the layers go through the autograd Functions of
ctgcn_b200/autograd.py (fused forward kernels, recomputing backward); under ``torch.no_grad()`` they write
straight into caller-provided buffers.  ``rnn_type='LSTM'`` (layers.py:27-28) runs on the fp32 sequence kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from . import autograd as _ag
from .plan import GraphPlan, plan_for


class _ForwardOnly(torch.autograd.Function):
    """Marks `out` as depending on `params` so that .backward() fails loudly (no silent zero grads)."""

    @staticmethod
    def forward(ctx, out, *params):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("ctgcn_b200: backward through the snapshot-parallel (multi-GPU) forward is not implemented")


def _guard(out, module, *inputs):
    if torch.is_grad_enabled() and (any(p.requires_grad for p in module.parameters()) or
                                    any(isinstance(t, torch.Tensor) and t.requires_grad for t in inputs)):
        return _ForwardOnly.apply(out, *[p for p in module.parameters() if p.requires_grad])
    return out


class CoreDiffusion(nn.Module):
    """k-core diffusion layer — reference layers.py:9-63.

    forward(x, adj_list): S_i = S_{i-1} + A_i x (:41-47) → relu (:48) → GRU over the core axis (:59)
    → Σ over cores (:60) → LayerNorm (:62).  ``adj_list`` is one snapshot's list of K sparse COO
    matrices (or a prebuilt GraphPlan).  ``linear`` is kept because the reference registers it
    (:24) and checkpoints contain it, although forward never reads it.
    """

    def __init__(self, input_dim, output_dim, core_num=1, bias=True, rnn_type='GRU'):
        super().__init__()
        if rnn_type not in ('LSTM', 'GRU'):
            raise AssertionError("rnn_type must be 'LSTM' or 'GRU'")
        self.input_dim, self.output_dim = input_dim, output_dim
        self.bias, self.core_num, self.rnn_type = bias, core_num, rnn_type
        # same construction order as the reference → identical default initialisation under one seed
        self.linear = nn.Linear(input_dim, output_dim)
        rnn_cls = nn.LSTM if rnn_type == 'LSTM' else nn.GRU
        self.rnn = rnn_cls(input_size=input_dim, hidden_size=output_dim, num_layers=1, bias=bias, batch_first=True)
        self.norm = nn.LayerNorm(output_dim)
        self._cell = _lib.CELLS[rnn_type]

    def _gru_params(self):
        r = self.rnn
        return (r.weight_ih_l0, r.weight_hh_l0, getattr(r, "bias_ih_l0", None) if self.bias else None,
                getattr(r, "bias_hh_l0", None) if self.bias else None)

    def forward_into(self, x, adj_list, out=None, scatter=None):
        """forward() writing into `out` ([N, H] view, any row stride) or — `scatter` = (slice_ptrs, row_stride,
        col_offset) — straight into the node slices' (peer) buffers; then nothing is returned.  Both are
        inference-only fast paths: when gradients are required (grad mode on and an input / parameter requires grad)
        the result is a fresh tensor with a grad_fn and `out` is NOT written — callers check ``requires_grad``."""
        plan = plan_for(adj_list, x.device)
        w_ih, w_hh, b_ih, b_hh = self._gru_params()
        if scatter is not None:
            ops.core_diffusion_scatter(plan, x.detach(), w_ih, w_hh, b_ih, b_hh, self.norm.weight, self.norm.bias, self.norm.eps,
                                       *scatter, cell=self._cell)
            return None
        if _ag.needs_grad(x, w_ih, w_hh, b_ih, b_hh, self.norm.weight, self.norm.bias):
            return _ag.CoreDiffusionFn.apply(x, plan, self._cell, self.norm.eps, w_ih, w_hh, b_ih, b_hh, self.norm.weight,
                                             self.norm.bias)          # `out` is left untouched: the caller checks requires_grad
        return ops.core_diffusion(plan, x.detach(), w_ih, w_hh, b_ih, b_hh, self.norm.weight, self.norm.bias, self.norm.eps,
                                  out=out, cell=self._cell)

    def forward(self, x, adj_list):
        return self.forward_into(x, adj_list)


class MLP(nn.Module):
    """Multi-layer perceptron — reference layers.py:67-106.

    layer_num == 1 → attribute ``linear``; otherwise ``linears`` (ModuleList).  selu after EVERY layer
    (the last included) iff activate_type == 'N'.  x may be dense [N, d_in] or a sparse COO tensor
    (one-hot identity, degree features): then the first layer is a row gather of Wᵀ.
    """

    def __init__(self, input_dim, hidden_dim, output_dim, layer_num, bias=True, activate_type='N'):
        super().__init__()
        if activate_type not in ('L', 'N'):
            raise AssertionError("activate_type must be 'L' or 'N'")
        if layer_num <= 0:
            raise AssertionError("layer_num must be positive")
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.layer_num, self.bias, self.activate_type = layer_num, bias, activate_type
        if layer_num == 1:
            self.linear = nn.Linear(input_dim, output_dim, bias=bias)
        else:
            dims = [input_dim] + [hidden_dim] * (layer_num - 1) + [output_dim]
            self.linears = nn.ModuleList(nn.Linear(dims[j], dims[j + 1], bias=bias) for j in range(layer_num))

    def _layers(self):
        return [self.linear] if self.layer_num == 1 else list(self.linears)

    def forward(self, x):
        act = _lib.ACT_SELU if self.activate_type == 'N' else _lib.ACT_NONE
        h = x
        for j, lin in enumerate(self._layers()):
            if j == 0 and (isinstance(h, GraphPlan) or (isinstance(h, torch.Tensor) and h.layout == torch.sparse_coo)):
                plan = plan_for(h, lin.weight.device)
                if _ag.needs_grad(lin.weight, lin.bias):
                    h = _ag.SparseLinearFn.apply(lin.weight, lin.bias, plan, act)
                else:
                    h = ops.spmm_linear(plan, lin.weight, lin.bias, act)
            elif _ag.needs_grad(h, lin.weight, lin.bias):
                h = _ag.LinearFn.apply(h, lin.weight, lin.bias, act)
            else:
                h = ops.linear(h.detach(), lin.weight, lin.bias, act)
        return h
