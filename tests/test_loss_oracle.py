"""Negative-sampling loss without a GPU: the numpy oracle against goldens from the unmodified reference
metrics.NegativeSamplingLoss (its own draws, loss value and autograd gradients), and the host logic of
ctgcn_b200.loss.NegativeSamplingLoss (list / tensor conventions, CSR upload, autograd wiring) with the C-ABI entry points
replaced by oracle-backed stand-ins inside the test."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle_loss

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["negloss_T1", "negloss_T3_128d", "negloss_small_neg"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def samples_of(meta, z):
    return [(z[f"node_idx{t}"], z[f"pos_idx{t}"], z[f"neg_idx{t}"]) for t in range(meta["T"])]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_loss_and_gradients(name):
    meta, z = load(name)
    emb = [z[f"emb{t}"] for t in range(meta["T"])]
    loss, grads = oracle_loss.neg_sampling_loss(emb, samples_of(meta, z), meta["Q"])
    assert abs(loss - float(z["loss"][0])) <= 2e-6 * abs(float(z["loss"][0]))
    for t, g in enumerate(grads):
        ref = z[f"grad{t}"]
        assert np.linalg.norm(g - ref) <= 2e-6 * np.linalg.norm(ref)


@pytest.mark.parametrize("name", CASES)
def test_reference_draws_satisfy_the_sampling_contract(name):
    """check_sample (the checker the GPU sampler is held to) accepts the reference's own draws; corrupted draws are rejected."""
    meta, z = load(name)
    batch, k = z["batch"], meta["neg_num"]
    for t in range(meta["T"]):
        pos, count = oracle_loss.to_padded(batch, z[f"node_idx{t}"], z[f"pos_idx{t}"], k)
        ni, pi = oracle_loss.from_padded(batch, pos, count)
        assert (ni == z[f"node_idx{t}"]).all() and (pi == z[f"pos_idx{t}"]).all()
        args = (batch, z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], k)
        oracle_loss.check_sample(*args, pos, count, z[f"neg_idx{t}"])
        assert (count == np.minimum(np.diff(z[f"pair_ptr{t}"])[batch], k)).all()
        full = int(np.argmax(count == k))                    # a node whose positives were drawn: duplicate one of them
        if count[full] == k and np.diff(z[f"pair_ptr{t}"])[batch[full]] > k:
            bad = pos.copy()
            bad[full, 1] = bad[full, 0]
            with pytest.raises(AssertionError):
                oracle_loss.check_sample(*args, bad, count, z[f"neg_idx{t}"])


@pytest.mark.parametrize("name,form", [("negloss_T1", "single"), ("negloss_T3_128d", "list"), ("negloss_T3_128d", "tensor"),
                                       ("negloss_small_neg", "list")])
def test_module_host_logic_reproduces_the_reference(lib, monkeypatch, name, form):
    """ctgcn_b200.loss.NegativeSamplingLoss fed the reference's draws gives the reference's loss and gradients."""
    import scipy.sparse as sp
    import fake_backend
    pkg = fake_backend.install(monkeypatch)
    from ctgcn_b200 import ops
    meta, z = load(name)
    T, n, k = meta["T"], meta["n"], meta["neg_num"]
    mats = [sp.csr_matrix((np.ones(len(z[f"pair_idx{t}"])), z[f"pair_idx{t}"], z[f"pair_ptr{t}"]), shape=(n, n)) for t in range(T)]
    pair_list = [m.tolil().rows for m in mats]                       # what helper.py:91-92 hands to the loss
    freqs = [z[f"freq{t}"].tolist() for t in range(T)]
    mod = pkg.loss.NegativeSamplingLoss(pair_list, freqs, neg_num=k, Q=meta["Q"])
    assert (mod.neg_sample_num, mod.Q) == (k, meta["Q"]) and mod.node_pair_list is pair_list

    calls = []

    def golden_draw(pair_ptr, pair_idx, freq, batch, neg_num, seed):     # the reference's own draw for this snapshot
        t = len(calls) % T
        calls.append(seed)
        assert (pair_ptr.numpy() == z[f"pair_ptr{t}"]).all() and (pair_idx.numpy() == z[f"pair_idx{t}"]).all()
        assert (freq.numpy() == z[f"freq{t}"]).all() and pair_idx.dtype == torch.int32 and pair_ptr.dtype == torch.int64
        pos, count = oracle_loss.to_padded(batch.numpy(), z[f"node_idx{t}"], z[f"pos_idx{t}"], neg_num)
        return torch.from_numpy(pos), torch.from_numpy(count), torch.from_numpy(z[f"neg_idx{t}"].astype(np.int32))

    monkeypatch.setattr(ops, "neg_sample", golden_draw)
    leaves = [torch.from_numpy(z[f"emb{t}"]).requires_grad_(True) for t in range(T)]
    if form == "single":
        emb = leaves[0]
    elif form == "list":
        emb = leaves
    else:                                                           # CTGCN.forward's [T, N, D] transposed view of [N, T, D]
        emb = torch.stack(leaves, dim=1).transpose(0, 1)
    batch = torch.from_numpy(z["batch"])
    mod.seed = 7
    loss = mod([emb, batch])
    assert tuple(loss.shape) == (1,) and abs(loss.item() - float(z["loss"][0])) <= 1e-5 * abs(float(z["loss"][0]))
    loss.backward()
    for t in range(T):
        ref = z[f"grad{t}"]
        assert np.linalg.norm(leaves[t].grad.numpy() - ref) <= 1e-5 * np.linalg.norm(ref)
    assert len(set(calls)) == T                                      # one draw per snapshot, distinct seeds
    mod([emb, batch])
    assert len(set(calls)) == 2 * T                                  # a new call draws afresh


def test_module_sampler_contract_with_the_stand_in(lib, monkeypatch):
    """End to end on CPU with the stand-in sampler: contract holds, empty snapshots contribute nothing, short lists raise."""
    import fake_backend
    pkg = fake_backend.install(monkeypatch)
    meta, z = load("negloss_small_neg")
    T, n, k = meta["T"], meta["n"], meta["neg_num"]
    import scipy.sparse as sp
    mats = [sp.csr_matrix((np.ones(len(z[f"pair_idx{t}"])), z[f"pair_idx{t}"], z[f"pair_ptr{t}"]), shape=(n, n)) for t in range(T)]
    mod = pkg.loss.NegativeSamplingLoss(mats, [z[f"freq{t}"] for t in range(T)], neg_num=k, Q=meta["Q"])   # scipy matrices are accepted too
    batch = torch.from_numpy(z["batch"])
    for t in range(T):
        pos, count, neg = mod.sample(t, batch, seed=3)
        oracle_loss.check_sample(z["batch"], z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], k, pos.numpy(), count.numpy(), neg.numpy())
    lonely = torch.tensor([n - 1])                                   # the node without any walk pair
    emb = [torch.from_numpy(z[f"emb{t}"]) for t in range(T)]
    assert mod([emb, lonely]).item() == 0.0
    short = pkg.loss.NegativeSamplingLoss(mats, [[1, 2, 3]] * T, neg_num=k, Q=1)
    with pytest.raises(ValueError):
        short([emb, batch])
    with pytest.raises(ValueError):
        pkg.loss.NegativeSamplingLoss(mats, [[1]] * T, neg_num=1000)


def test_device_sampler_algorithm_is_uniform_and_meets_the_contract():
    """The bit-level restatement of the kernel's generator + Floyd sampling (what the GPU test compares against): every
    element of range(n) is included with probability m/n (chi-square over many seeds), subsets are distinct, and the full
    draw satisfies the reference's sampling contract."""
    n, m, trials = 37, 5, 6000
    hits = np.zeros(n)
    first = np.zeros(n)
    for seed in range(trials):
        s = oracle_loss.floyd(n, m, seed * 7919 + 1, stream=seed % 11)
        assert len(set(s)) == m and min(s) >= 0 and max(s) < n
        hits[s] += 1
        first[s[0]] += 1
    expect = trials * m / n
    chi2 = ((hits - expect) ** 2 / expect).sum()
    assert chi2 < 75.0, chi2                      # 36 degrees of freedom: P(chi2 > 75) ≈ 1.5e-4
    # pairs: inclusion of (0, 1) together should be m(m-1)/(n(n-1))
    both = sum(1 for seed in range(trials) if {0, 1} <= set(oracle_loss.floyd(n, m, seed * 104729 + 3, 0)))
    p = m * (m - 1) / (n * (n - 1))
    assert abs(both - trials * p) < 5 * np.sqrt(trials * p)
    meta, z = load("negloss_T3_128d")
    for t in range(meta["T"]):
        pos, count, neg = oracle_loss.device_sample(z["batch"], z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], meta["neg_num"], 99 + t)
        oracle_loss.check_sample(z["batch"], z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], meta["neg_num"], pos, count, neg)
        pos2, _, neg2 = oracle_loss.device_sample(z["batch"], z[f"pair_ptr{t}"], z[f"pair_idx{t}"], z[f"freq{t}"], meta["neg_num"], 100 + t)
        assert (pos != pos2).any() and (neg != neg2).any()           # another seed, another draw


def test_module_rejects_out_of_range_node_ids(lib):
    """The kernels index the embedding without bounds checks; like the reference (IndexError at metrics.py:56-57 / 78-93) the module
    refuses node ids outside the embedding — from the co-occurrence lists, the negative list or the batch — before anything is
    launched (no GPU needed for the refusal)."""
    import torch
    from ctgcn_b200.loss import NegativeSamplingLoss
    pairs = [[[1, 2], [0], [0, 3], [2]]]                       # 4 nodes
    emb = torch.zeros(4, 8)
    with pytest.raises(IndexError, match="negative lists"):
        NegativeSamplingLoss(pairs, [[0, 1, 2, 9]], neg_num=2)([emb, torch.tensor([0, 1])])
    with pytest.raises(IndexError, match="co-occurrence"):
        NegativeSamplingLoss([[[1, 7], [0], [0], [2]]], [[0, 1, 2, 3]], neg_num=2)([emb, torch.tensor([0, 1])])
    with pytest.raises(IndexError, match="batch node ids"):
        NegativeSamplingLoss(pairs, [[0, 1, 2, 3]], neg_num=2)([emb, torch.tensor([0, 4])])
    with pytest.raises(IndexError, match="embedding has 3 rows"):
        NegativeSamplingLoss(pairs, [[0, 1, 2, 3]], neg_num=2)([emb[:3], torch.tensor([0, 1])])
