"""TEST INFRASTRUCTURE: a torch-CPU stand-in for ctgcn_b200.ops, built on the oracle, so that the host-side
autograd wiring of ctgcn_b200 (argument order, saved tensors, None gradients, stack / transpose handling) can be checked
without a GPU.  It is installed by monkeypatching inside tests only; the product has no CPU path."""
import numpy as np
import scipy.sparse as sp
import torch

from oracle import oracle_torch


class FakePlan:
    def __init__(self, mats):
        self.mats = [sp.csr_matrix(m).astype(np.float32) for m in mats]
        self.n_rows, self.n_cols = self.mats[0].shape
        self.k = len(self.mats)

    def transposed(self):
        return FakePlan([m.T for m in self.mats])

    def torch_list(self):
        return [oracle_torch.to_torch_coo(m) for m in self.mats]


def _to_scipy(m):
    if isinstance(m, torch.Tensor):
        idx, val = m._indices().numpy(), m._values().numpy()
        return sp.coo_matrix((val, (idx[0], idx[1])), shape=tuple(m.shape))
    return m


def plan_for(adj_list, device):
    if isinstance(adj_list, FakePlan):
        return adj_list
    if isinstance(adj_list, torch.Tensor) or hasattr(adj_list, "tocoo"):
        return FakePlan([_to_scipy(adj_list)])
    return FakePlan([_to_scipy(a) for a in adj_list])


def _sd(w_ih, w_hh, b_ih, b_hh, ln_w, ln_b):
    sd = {"rnn.weight_ih_l0": w_ih, "rnn.weight_hh_l0": w_hh, "norm.weight": ln_w, "norm.bias": ln_b}
    if b_ih is not None:
        sd.update({"rnn.bias_ih_l0": b_ih, "rnn.bias_hh_l0": b_hh})
    return {k: v.detach() for k, v in sd.items()}


def cumspmm(plan, x, relu=True):
    acc, outs = None, []
    for a in plan.torch_list():
        prod = torch.sparse.mm(a, x)
        acc = prod if acc is None else acc + prod
        outs.append(torch.relu(acc) if relu else acc)
    return torch.stack(outs, dim=1)


def cumspmm_bwd(plan_t, g):
    zo = g.flip(1).cumsum(1).flip(1)
    return sum(torch.sparse.mm(a, zo[:, j].contiguous()) for j, a in enumerate(plan_t.torch_list()))


def rnn_seq(seq, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, out=None, cell=0):
    hs = oracle_torch._gru_all_outputs(seq, _sd(w_ih, w_hh, b_ih, b_hh, ln_w, ln_b), "")
    pre = hs.sum(dim=1) if mode == 0 else hs
    y = torch.nn.functional.layer_norm(pre, (pre.shape[-1],), ln_w.detach(), ln_b.detach(), eps)
    if out is not None:
        out.copy_(y)
        return out
    return y


def core_diffusion(plan, x, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, out=None, cell=0, scatter=None):
    assert scatter is None
    return rnn_seq(cumspmm(plan, x.detach()), w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, 0, out=out, cell=cell)


def linear(x, w, b, act):
    y = torch.nn.functional.linear(x.detach(), w.detach(), None if b is None else b.detach())
    return torch.selu(y) if act == 1 else y


def spmm_linear(plan, w, b, act):
    y = torch.sparse.mm(plan.torch_list()[0], w.detach().t().contiguous())
    if b is not None:
        y = y + b.detach()
    return torch.selu(y) if act == 1 else y


# ---- negative-sampling loss (ctgcn_b200/loss.py): numpy stand-ins with the padded sample format of ctgcn_neg_sample
def neg_sample(pair_ptr, pair_idx, freq, batch, neg_num, seed):
    rng = np.random.default_rng(seed % (2 ** 63))
    ptr, idx, fr, bt = (t.numpy() for t in (pair_ptr, pair_idx, freq, batch))
    if len(fr) < neg_num:
        raise ValueError("Sample larger than population or is negative")
    pos = np.full((len(bt), neg_num), -1, dtype=np.int32)
    count = np.zeros(len(bt), dtype=np.int32)
    for b, node in enumerate(bt):
        nb = idx[ptr[node]:ptr[node + 1]]
        take = nb if len(nb) <= neg_num else rng.choice(nb, size=neg_num, replace=False)
        pos[b, :len(take)] = take
        count[b] = len(take)
    neg = fr[rng.choice(len(fr), size=neg_num, replace=False)].astype(np.int32)
    return torch.from_numpy(pos), torch.from_numpy(count), torch.from_numpy(neg)


def neg_loss_fwd(emb, batch, pos, count, neg, q):
    from oracle import oracle_loss
    ni, pi = oracle_loss.from_padded(batch.numpy(), pos.numpy(), count.numpy())
    loss, _ = oracle_loss.snapshot_loss(emb.detach().numpy(), ni, pi, neg.numpy(), q)
    return torch.tensor([loss], dtype=torch.float32), torch.zeros(1)


def neg_loss_bwd(emb, batch, pos, count, neg, q, grad_loss, ws):
    from oracle import oracle_loss
    ni, pi = oracle_loss.from_padded(batch.numpy(), pos.numpy(), count.numpy())
    _, grad = oracle_loss.snapshot_loss(emb.detach().numpy(), ni, pi, neg.numpy(), q)
    return torch.from_numpy((grad * float(grad_loss.reshape(-1)[0])).astype(np.float32))


class PassThroughStager:
    def __init__(self, x_list, order, dev, depth=2):
        self.x_list = x_list

    def get(self, t):
        return self.x_list[t]


def install(monkeypatch):
    import ctgcn_b200
    from ctgcn_b200 import layers, models, ops
    for name in ("cumspmm", "cumspmm_bwd", "rnn_seq", "core_diffusion", "linear", "spmm_linear", "neg_sample", "neg_loss_fwd",
                 "neg_loss_bwd"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(layers, "plan_for", plan_for)
    monkeypatch.setattr(models, "_HostFeatureStager", PassThroughStager)
    return ctgcn_b200
