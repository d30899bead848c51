"""Negative-sampling loss on the GPU: the device sampler bit for bit against its restatement and against the
reference's sampling contract; the fused loss / gradient kernels against goldens from the unmodified reference
metrics.NegativeSamplingLoss (its own draws, loss and autograd gradients); the drop-in module end to end.

First executed on a B200 in round 2 (call 1): 19 tests green at first run, un-gated since.

Tolerances: sampler exact (integer work); loss 1e-5 relative, gradients 1e-5 relL2 against the reference's fp32 autograd
(fp32 dot products in another order; float atomics)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle_loss

pytestmark = [pytest.mark.gpu]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["negloss_T1", "negloss_T3_128d", "negloss_small_neg"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return json.loads(bytes(z["meta"]).decode()), z


def dev_arrays(z, t, dev):
    return (torch.from_numpy(z[f"pair_ptr{t}"]).to(dev), torch.from_numpy(z[f"pair_idx{t}"]).to(dev),
            torch.from_numpy(z[f"freq{t}"]).to(dev))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("seed", [0, 12345, 2 ** 63 + 17])
def test_sampler_bit_exact_and_contract(name, seed, lib, cuda_device):
    from ctgcn_b200 import ops
    meta, z = load(name)
    k = meta["neg_num"]
    batch = torch.from_numpy(z["batch"]).to(cuda_device)
    for t in range(meta["T"]):
        ptr, idx, freq = dev_arrays(z, t, cuda_device)
        pos, count, neg = (a.cpu().numpy() for a in ops.neg_sample(ptr, idx, freq, batch, k, seed))
        oracle_loss.check_sample(z["batch"], z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], k, pos, count, neg)
        wpos, wcount, wneg = oracle_loss.device_sample(z["batch"], z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], k, seed)
        assert (pos == wpos).all() and (count == wcount).all() and (neg == wneg).all()


def test_sampler_edge_cases(lib, cuda_device):
    from ctgcn_b200 import ops
    meta, z = load("negloss_small_neg")
    k = meta["neg_num"]
    ptr, idx, freq = dev_arrays(z, 0, cuda_device)
    empty = torch.zeros(0, dtype=torch.int64, device=cuda_device)
    pos, count, neg = ops.neg_sample(ptr, idx, freq, empty, k, 1)
    assert tuple(pos.shape) == (0, k) and tuple(neg.shape) == (k,)
    outside = torch.tensor([-1, meta["n"], meta["n"] - 1], dtype=torch.int64, device=cuda_device)    # ids outside keep nothing
    pos, count, _ = ops.neg_sample(ptr, idx, freq, outside, k, 1)
    assert count.cpu().tolist() == [0, 0, 0] and bool((pos == -1).all())
    exact = freq[:k].contiguous()                                     # frequency list of exactly neg_num entries: all of them
    _, _, neg = ops.neg_sample(ptr, idx, exact, outside, k, 1)
    assert neg.cpu().tolist() == exact.cpu().tolist()
    with pytest.raises(ValueError):
        ops.neg_sample(ptr, idx, freq[: k - 1].contiguous(), outside, k, 1)
    with pytest.raises(lib.CtgcnError):
        ops.neg_sample(ptr.cpu(), idx, freq, outside, k, 1)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("view", [False, True])
def test_loss_and_gradient_against_reference(name, view, lib, cuda_device):
    """The reference's own draws through the fused kernels → the reference's loss and autograd gradients."""
    from ctgcn_b200 import ops
    meta, z = load(name)
    T, k, q = meta["T"], meta["neg_num"], float(meta["Q"])
    batch = torch.from_numpy(z["batch"]).to(cuda_device)
    embs = [torch.from_numpy(z[f"emb{t}"]).to(cuda_device) for t in range(T)]
    if view:                                                          # rows of a [N, T, D] buffer: row stride T·D
        buf = torch.stack(embs, dim=1).contiguous()
        embs = [buf[:, t, :] for t in range(T)]
    total = 0.0
    for t in range(T):
        pos, count = oracle_loss.to_padded(z["batch"], z[f"node_idx{t}"], z[f"pos_idx{t}"], k)
        pos, count = torch.from_numpy(pos).to(cuda_device), torch.from_numpy(count).to(cuda_device)
        neg = torch.from_numpy(z[f"neg_idx{t}"].astype(np.int32)).to(cuda_device)
        before = lib.launch_count()
        loss, ws = ops.neg_loss_fwd(embs[t], batch, pos, count, neg, q)
        assert lib.launch_count() - before == 3
        total += loss.item()
        want_l, want_g = oracle_loss.snapshot_loss(z[f"emb{t}"], z[f"node_idx{t}"], z[f"pos_idx{t}"], z[f"neg_idx{t}"], q)
        assert abs(loss.item() - want_l) <= 1e-5 * abs(want_l)
        for scale in (1.0, -2.5):                                     # twice on one workspace: the backward is re-entrant
            g = ops.neg_loss_bwd(embs[t], batch, pos, count, neg, q, torch.tensor([scale], device=cuda_device), ws).cpu().numpy()
            ref = z[f"grad{t}"] * scale
            assert np.linalg.norm(g - ref) <= 1e-5 * np.linalg.norm(ref)
            assert np.linalg.norm(g - want_g * scale) <= 1e-5 * np.linalg.norm(ref)
    assert abs(total - float(z["loss"][0])) <= 1e-5 * abs(float(z["loss"][0]))


@pytest.mark.parametrize("form", ["list", "tensor", "single"])
def test_module_end_to_end(form, lib, cuda_device):
    """ctgcn_b200.loss.NegativeSamplingLoss with its own device draws: value and gradients equal the oracle on those draws."""
    import scipy.sparse as sp
    import ctgcn_b200 as pkg
    meta, z = load("negloss_T3_128d")
    T, n, k, q = (1 if form == "single" else meta["T"]), meta["n"], meta["neg_num"], meta["Q"]
    mats = [sp.csr_matrix((np.ones(len(z[f"pair_idx{t}"])), z[f"pair_idx{t}"], z[f"pair_ptr{t}"]), shape=(n, n)) for t in range(T)]
    mod = pkg.loss.NegativeSamplingLoss([m.tolil().rows for m in mats], [z[f"freq{t}"].tolist() for t in range(T)], neg_num=k, Q=q)
    mod.seed = 11
    leaves = [torch.from_numpy(z[f"emb{t}"]).to(cuda_device).requires_grad_(True) for t in range(T)]
    emb = leaves[0] if form == "single" else (leaves if form == "list" else torch.stack(leaves, dim=1).transpose(0, 1))
    batch = torch.from_numpy(z["batch"]).to(cuda_device)
    loss = mod([emb, batch])
    assert tuple(loss.shape) == (1,)
    loss.backward()
    want = 0.0
    for t in range(T):
        pos, count, neg = (a.cpu().numpy() for a in mod.sample(t, batch, (11 + 0) * 1_000_003 + t))   # the draw of call 0
        ni, pi = oracle_loss.from_padded(z["batch"], pos, count)
        l, g = oracle_loss.snapshot_loss(z[f"emb{t}"], ni, pi, neg, q)
        want += l
        got = leaves[t].grad.cpu().numpy()
        assert np.linalg.norm(got - g) <= 1e-5 * np.linalg.norm(g)
    assert abs(loss.item() - want) <= 1e-5 * abs(want)
    second = mod([emb, batch]).item()                                  # call 1 draws afresh
    assert second != loss.item()
