"""Drop-in ``models`` module: ``CDN``, ``CGCN``, ``CTGCN`` (reference models.py:8-42, 129-187, 191-253) on
top of the sm_100a kernels, with the reference's constructor / forward signatures, attribute names
(``mlp_list``, ``duffision_list`` [sic], ``rnn``, ``norm``, ``method_name``) and ``state_dict`` keys.

Snapshot-parallel mode is an explicit opt-in (``model.snapshot_parallel = True``; the reference has no such
mode, so an initialised ``torch.distributed`` alone never changes behaviour — DDP-style use keeps the ordinary autograd path):
with world size G > 1 ``CTGCN.forward`` then computes only the snapshots t ≡ rank (mod G) — their MLP_t / CDN_t weights, features and
graph plans are disjoint (models.py:225-231) — exchanges the per-snapshot embeddings once, runs the
temporal GRU on this rank's node slice, and (optionally) gathers the result.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from . import autograd as _ag
from .layers import CoreDiffusion, MLP
from . import dist as _dist


_copy_streams = {}


class _HostFeatureStager:
    """Features handed over as (pinned) HOST tensors are uploaded on a side stream, `depth` snapshots ahead of the
    compute stream, so that the H2D copy of snapshot t+1 overlaps the kernels of snapshot t.  Device tensors and
    sparse inputs pass through untouched."""

    def __init__(self, x_list, order, dev, depth=2):
        self.x_list, self.order, self.dev, self.depth = x_list, list(order), dev, depth
        if str(dev) not in _copy_streams:            # one copy stream per device, created on first use
            _copy_streams[str(dev)] = torch.cuda.Stream(device=dev)
        self.stream = _copy_streams[str(dev)]
        self.pending = {}
        self.next = 0

    def _is_host_dense(self, x):
        return isinstance(x, torch.Tensor) and x.layout == torch.strided and not x.is_cuda

    def _issue(self):
        while self.next < len(self.order) and len(self.pending) < self.depth:
            t = self.order[self.next]
            self.next += 1
            x = self.x_list[t]
            if self._is_host_dense(x):
                with torch.cuda.stream(self.stream):
                    xd = x.to(self.dev, dtype=torch.float32, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                self.pending[t] = (xd, ev)

    def get(self, t):
        self._issue()
        x = self.x_list[t]
        if not self._is_host_dense(x):
            return x
        xd, ev = self.pending.pop(t)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        xd.record_stream(cur)
        self._issue()
        return xd


class CDN(nn.Module):
    """Stack of ``diffusion_num`` CoreDiffusion layers sharing one adj_list — reference models.py:8-42."""

    def __init__(self, input_dim, hidden_dim, output_dim, diffusion_num, bias=True, rnn_type='GRU'):
        super().__init__()
        if diffusion_num < 1:
            raise ValueError("number of layers should be positive!")
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.diffusion_num, self.bias, self.rnn_type = diffusion_num, bias, rnn_type
        widths = [input_dim, output_dim] if diffusion_num == 1 else \
            [input_dim] + [hidden_dim] * (diffusion_num - 1) + [output_dim]
        self.diffusion_list = nn.ModuleList(
            CoreDiffusion(widths[l], widths[l + 1], bias=bias, rnn_type=rnn_type) for l in range(diffusion_num))

    def forward_into(self, x, adj_list, out=None, scatter=None):
        last = self.diffusion_num - 1
        for l, layer in enumerate(self.diffusion_list):
            x = layer.forward_into(x, adj_list, out if l == last else None, scatter if l == last else None)
        return x

    def forward(self, x, adj_list):
        return self.forward_into(x, adj_list)


class CGCN(nn.Module):
    """Static k-core GCN: one shared MLP + CDN applied per snapshot — reference models.py:129-187."""

    def __init__(self, input_dim, hidden_dim, output_dim, trans_num, diffusion_num, bias=True, rnn_type='GRU',
                 model_type='C', trans_activate_type='L'):
        super().__init__()
        if model_type not in ('C', 'S'):
            raise AssertionError("model_type must be 'C' or 'S'")
        if trans_activate_type not in ('L', 'N'):
            raise AssertionError("trans_activate_type must be 'L' or 'N'")
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.trans_num, self.diffusion_num, self.bias, self.rnn_type = trans_num, diffusion_num, bias, rnn_type
        self.model_type, self.trans_activate_type = model_type, trans_activate_type
        self.method_name = 'CGCN' + '-' + model_type
        mid = hidden_dim if model_type == 'C' else output_dim
        self.mlp = MLP(input_dim, hidden_dim, mid, trans_num, bias=bias, activate_type=trans_activate_type)
        self.duffision = CDN(mid, output_dim, output_dim, diffusion_num, rnn_type=rnn_type)

    def cgcn(self, x, adj):
        trans = self.mlp(x)
        emb = self.duffision(trans, adj)
        return (emb, trans) if self.model_type == 'S' else emb

    def forward(self, x, adj):
        if not isinstance(x, list):
            return self.cgcn(x, adj)
        res = [self.cgcn(xi, ai) for xi, ai in zip(x, adj)]
        if self.model_type == 'C':
            return res
        return [r[0] for r in res], [r[1] for r in res]


class CTGCN(nn.Module):
    """k-core temporal GCN — reference models.py:191-253.

    forward(x_list, adj_list): per snapshot t, MLP_t then CDN_t with independent weights (:243-247); stack to
    [N,T,D] (:248); temporal GRU, h0 = 0 (:249); LayerNorm, returned as the transposed view [T,N,D] (:250).
    model_type 'S' additionally returns the list of MLP outputs (:251-253).

    The per-snapshot results are written straight into the [N,T,D] buffer the temporal GRU reads
    (no stack / transpose copies).
    """

    def __init__(self, input_dim, hidden_dim, output_dim, trans_num, diffusion_num, duration, bias=True, rnn_type='GRU',
                 model_type='C', trans_activate_type='L'):
        super().__init__()
        if model_type not in ('C', 'S'):
            raise AssertionError("model_type must be 'C' or 'S'")
        if trans_activate_type not in ('L', 'N'):
            raise AssertionError("trans_activate_type must be 'L' or 'N'")
        if rnn_type not in ('LSTM', 'GRU'):
            raise AssertionError("rnn_type must be 'LSTM' or 'GRU'")
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.rnn_type, self.model_type, self.trans_activate_type = rnn_type, model_type, trans_activate_type
        self.method_name = 'CTGCN' + '-' + model_type
        self.duration, self.trans_num, self.diffusion_num, self.bias = duration, trans_num, diffusion_num, bias
        mid = hidden_dim if model_type == 'C' else output_dim
        self.mlp_list = nn.ModuleList()
        self.duffision_list = nn.ModuleList()
        for _ in range(duration):  # interleaved like the reference → same default initialisation under one seed
            self.mlp_list.append(MLP(input_dim, hidden_dim, mid, trans_num, bias=bias, activate_type=trans_activate_type))
            self.duffision_list.append(CDN(mid, output_dim, output_dim, diffusion_num, rnn_type=rnn_type))
        rnn_cls = nn.LSTM if rnn_type == 'LSTM' else nn.GRU
        self.rnn = rnn_cls(output_dim, output_dim, num_layers=1, bias=bias, batch_first=True)
        self.norm = nn.LayerNorm(output_dim)
        self._cell = _lib.CELLS[rnn_type]
        # snapshot-parallel execution is opt-in: every rank must then pass the same x_list / adj_list structure, the forward is
        # inference-only, and for model_type 'S' trans_list holds None for snapshots owned by other ranks
        self.snapshot_parallel = False
        self.gather_output = True

    def _temporal(self, hx, out=None):
        r = self.rnn
        b_ih = r.bias_ih_l0 if self.bias else None
        b_hh = r.bias_hh_l0 if self.bias else None
        if _ag.needs_grad(hx, r.weight_ih_l0, r.weight_hh_l0, b_ih, b_hh, self.norm.weight, self.norm.bias):
            return _ag.RnnSeqFn.apply(hx, self._cell, self.norm.eps, r.weight_ih_l0, r.weight_hh_l0, b_ih, b_hh,
                                      self.norm.weight, self.norm.bias)
        return ops.rnn_seq(hx, r.weight_ih_l0, r.weight_hh_l0, b_ih, b_hh, self.norm.weight, self.norm.bias, self.norm.eps,
                           _lib.GRU_EACH_LN, out=out, cell=self._cell)

    def forward(self, x_list, adj_list):
        if self.snapshot_parallel and _dist.world_size() > 1:
            return _dist.ctgcn_forward_sharded(self, x_list, adj_list)
        T = len(x_list)
        dev = self.norm.weight.device
        hx, trans_list, emb_list = None, [], []
        stager = _HostFeatureStager(x_list, range(T), dev)
        for t in range(T):
            trans = self.mlp_list[t](stager.get(t))
            trans_list.append(trans)
            if hx is None:
                hx = torch.empty(trans.shape[0], T, self.output_dim, dtype=torch.float32, device=dev)
            emb_list.append(self.duffision_list[t].forward_into(trans, adj_list[t], out=hx[:, t, :]))
        if any(e.requires_grad for e in emb_list):
            hx = torch.stack(emb_list, dim=1)          # training: the autograd-visible [N, T, D] (models.py:248)
        out = self._temporal(hx).transpose(0, 1)
        return out if self.model_type == 'C' else (out, trans_list)
