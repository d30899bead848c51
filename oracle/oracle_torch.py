"""torch-CPU port of the CTGCN forward hot path (TEST INFRASTRUCTURE / CPU baseline).

Same library calls, in the same order, as the reference modules — so that timing
this port on the GPU box's host cores is an honest stand-in for "the reference's own
CPU path" (the reference is Python and lives only in the build container; it cannot
travel).  Functional style over a flat ``state_dict``; no nn.Module subclasses.

Reference lines restated (paths relative to /root/reference):
  layers.py:38-63  CoreDiffusion.forward      layers.py:95-106  MLP.forward
  models.py:39-42  CDN.forward                models.py:240-253 CTGCN.forward
  models.py:165-187 CGCN.forward              utils.py:89-95    scipy → torch COO
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def to_torch_coo(spmat):
    """utils.sparse_mx_to_torch_sparse_tensor (utils.py:89-95): uncoalesced COO, int64 idx, fp32 val."""
    m = spmat.tocoo()
    idx = torch.from_numpy(np.vstack((m.row, m.col))).long()
    val = torch.from_numpy(m.data).float()
    return torch.sparse_coo_tensor(idx, val, torch.Size(m.shape))


def _gru_all_outputs(seq, sd, prefix):
    """nn.GRU / nn.LSTM(num_layers=1, batch_first=True)(seq)[0] with zero initial state (layers.py:27-30,59;
    models.py:234-237,249); the cell is told by the stored weight shape."""
    w_ih, w_hh = sd[prefix + "rnn.weight_ih_l0"], sd[prefix + "rnn.weight_hh_l0"]
    has_bias = (prefix + "rnn.bias_ih_l0") in sd
    flat = [w_ih, w_hh] + ([sd[prefix + "rnn.bias_ih_l0"], sd[prefix + "rnn.bias_hh_l0"]] if has_bias else [])
    h0 = torch.zeros(1, seq.shape[0], w_hh.shape[1], dtype=seq.dtype, device=seq.device)
    if w_hh.shape[0] == 4 * w_hh.shape[1]:   # rnn_type='LSTM' (layers.py:27-28, models.py:234-235): [i;f;g;o] stacked
        return torch._VF.lstm(seq, (h0, torch.zeros_like(h0)), flat, has_bias, 1, 0.0, False, False, True)[0]
    out, _ = torch._VF.gru(seq, h0, flat, has_bias, 1, 0.0, False, False, True)
    return out


def core_diffusion(x, adj_list, sd, prefix=""):
    """layers.py:38-63: K × torch.sparse.mm with running sum, relu, stack→[N,K,D], GRU, Σ_K, LayerNorm."""
    partial = []
    for i, adj in enumerate(adj_list):
        prod = torch.sparse.mm(adj, x)
        partial.append(prod if i == 0 else partial[-1] + prod)
    hx = torch.stack([F.relu(p) for p in partial], dim=0).transpose(0, 1)
    summed = _gru_all_outputs(hx, sd, prefix).sum(dim=1)
    return F.layer_norm(summed, (summed.shape[-1],), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-5)


def mlp(x, sd, prefix, layer_num, activate_type):
    """layers.py:95-106 (x may be a sparse COO tensor: nn.Linear → sparse addmm)."""
    names = ["linear"] if layer_num == 1 else [f"linears.{j}" for j in range(layer_num)]
    h = x
    for nm in names:
        w, b = sd[prefix + nm + ".weight"], sd.get(prefix + nm + ".bias")
        h = F.linear(h, w, b)
        if activate_type == "N":
            h = F.selu(h)
    return h


def cdn(x, adj_list, sd, prefix, diffusion_num):
    """models.py:39-42."""
    for l in range(diffusion_num):
        x = core_diffusion(x, adj_list, sd, f"{prefix}diffusion_list.{l}.")
    return x


def cgcn(x, adj, sd, trans_num, diffusion_num, model_type="C", trans_activate_type="L"):
    """models.py:165-187."""
    def one(xi, ai):
        trans = mlp(xi, sd, "mlp.", trans_num, trans_activate_type)
        emb = cdn(trans, ai, sd, "duffision.", diffusion_num)
        return (emb, trans) if model_type == "S" else emb

    if isinstance(x, list):
        res = [one(xi, ai) for xi, ai in zip(x, adj)]
        if model_type == "C":
            return res
        return [r[0] for r in res], [r[1] for r in res]
    return one(x, adj)


def ctgcn(x_list, adj_list, sd, trans_num, diffusion_num, model_type="C", trans_activate_type="L"):
    """models.py:240-253."""
    hx, trans_list = [], []
    for t in range(len(x_list)):
        trans = mlp(x_list[t], sd, f"mlp_list.{t}.", trans_num, trans_activate_type)
        trans_list.append(trans)
        hx.append(cdn(trans, adj_list[t], sd, f"duffision_list.{t}.", diffusion_num))
    seq = torch.stack(hx).transpose(0, 1)
    out = _gru_all_outputs(seq, sd, "")
    out = F.layer_norm(out, (out.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-5).transpose(0, 1)
    return out if model_type == "C" else (out, trans_list)
