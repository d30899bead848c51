#!/usr/bin/env python
"""Goldens for the negative-sampling loss from the UNMODIFIED reference `metrics.py` — TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):  python oracle/make_loss_golden.py
The reference reseeds `random` from the OS inside every call (metrics.py:72), so for the run that produces a golden
`random.seed` is pinned and every `random.sample` draw is recorded: the fixture stores the walk-pair lists, the frequency
list, the batch, the reference's own draws, its loss value and its autograd gradient w.r.t. every snapshot's embeddings.
`oracle/oracle_loss.py` is asserted against it here and re-checked by tests/test_loss_oracle.py.
"""
from __future__ import annotations

import json
import os
import random
import sys
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

from oracle import ref_compat  # noqa: E402
ref_compat.apply()
import metrics as ref_metrics  # noqa: E402  (reference)

from oracle import oracle_loss  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def walk_pairs(n, avg_deg, seed, heavy=()):
    """A symmetric co-occurrence matrix like preprocessing/random_walk.py:45-46 writes (0/1 entries, no diagonal) and the
    frequency-expanded negative list of :53-59.  `heavy`: nodes given many more neighbours than neg_num."""
    rng = np.random.default_rng(seed)
    m = int(n * avg_deg / 2)
    u, v = rng.integers(0, n, m), rng.integers(0, n, m)
    for h in heavy:
        extra = rng.choice(n, size=min(n - 1, 90), replace=False)
        u = np.concatenate([u, np.full(extra.shape, h)])
        v = np.concatenate([v, extra])
    keep = u != v
    a = sp.coo_matrix((np.ones(keep.sum()), (u[keep], v[keep])), shape=(n, n))
    a = ((a + a.T) > 0).astype(np.float64).tocsr()
    a[n - 1, :] = 0                                     # one node without any pair (isolated in the walks)
    a[:, n - 1] = 0
    a.eliminate_zeros()
    freq = np.asarray(a.sum(axis=1)).ravel() + 1.0
    z = 1e-3
    rep = ((freq / freq.sum()) ** 0.75 / z).astype(int)
    neg_list = np.repeat(np.arange(n), rep).tolist()
    return a.tocsr(), neg_list


def run_case(name, n, T, d, batch, neg_num, Q, seed):
    rng = np.random.default_rng(seed)
    mats, freqs = zip(*[walk_pairs(n, 6.0, seed * 10 + t, heavy=(3, 7 + t)) for t in range(T)])
    pair_list = [m.tolil().rows for m in mats]                       # helper.py:91-92
    emb = [torch.from_numpy((0.3 * rng.standard_normal((n, d))).astype(np.float32)).requires_grad_(True) for _ in range(T)]
    loss_mod = ref_metrics.NegativeSamplingLoss(pair_list, list(freqs), neg_num=neg_num, Q=Q)

    draws = []
    real_sample, real_seed = random.sample, random.seed

    def rec_sample(pop, k):
        out = real_sample(pop, k)
        draws.append(list(out))
        return out

    random.seed(seed)
    random.sample, random.seed = rec_sample, (lambda *a, **k: None)
    try:
        loss = loss_mod([emb, torch.from_numpy(np.asarray(batch, dtype=np.int64))])
    finally:
        random.sample, random.seed = real_sample, real_seed
    loss.backward()

    # rebuild the reference's index arrays from the recorded draws (same order as metrics.py:74-88)
    samples, k = [], 0
    for t in range(T):
        node_idx, pos_idx = [], []
        for b in batch:
            nb = list(pair_list[t][b])
            if len(nb) <= neg_num:
                take = nb
            else:
                take = draws[k]
                k += 1
            node_idx += [b] * len(take)
            pos_idx += take
        if node_idx:
            neg_idx = draws[k]
            k += 1
        else:
            neg_idx = []
        samples.append((node_idx, pos_idx, neg_idx))
    assert k == len(draws), (k, len(draws))

    want = float(loss.item())
    got, grads = oracle_loss.neg_sampling_loss([e.detach().numpy() for e in emb], samples, Q)
    gerr = max(np.linalg.norm(g - e.grad.numpy()) / max(np.linalg.norm(e.grad.numpy()), 1e-30) for g, e in zip(grads, emb))
    print(f"  {name}: reference loss {want:.6f}, oracle {got:.6f} (rel {abs(got - want) / abs(want):.1e}); worst grad relL2 {gerr:.1e}; "
          f"samples per snapshot {[len(s[0]) for s in samples]}")
    assert abs(got - want) / abs(want) < 2e-6 and gerr < 2e-6

    arrays = {}
    for t in range(T):
        m = mats[t]
        arrays[f"pair_ptr{t}"] = m.indptr.astype(np.int64)
        arrays[f"pair_idx{t}"] = m.indices.astype(np.int32)
        arrays[f"freq{t}"] = np.asarray(freqs[t], dtype=np.int32)
        arrays[f"emb{t}"] = emb[t].detach().numpy()
        arrays[f"grad{t}"] = emb[t].grad.numpy().astype(np.float32)
        for nm, a in zip(("node_idx", "pos_idx", "neg_idx"), samples[t]):
            arrays[f"{nm}{t}"] = np.asarray(a, dtype=np.int64)
    arrays["batch"] = np.asarray(batch, dtype=np.int64)
    arrays["loss"] = np.asarray([want], dtype=np.float64)
    meta = dict(name=name, kind="neg_sampling_loss", n=n, T=T, d=d, neg_num=neg_num, Q=Q, seed=seed)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"  wrote {name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    os.makedirs(GOLD, exist_ok=True)
    run_case("negloss_T1", n=200, T=1, d=32, batch=list(range(0, 200, 3)), neg_num=20, Q=10, seed=1)
    run_case("negloss_T3_128d", n=300, T=3, d=128, batch=[299, 3, 7, 8, 9, 150, 151, 152, 10, 11] + list(range(20, 120)),
             neg_num=20, Q=10, seed=2)
    run_case("negloss_small_neg", n=120, T=2, d=64, batch=list(range(120)), neg_num=5, Q=3, seed=3)


if __name__ == "__main__":
    main()
