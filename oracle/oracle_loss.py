"""numpy restatement of the reference's unsupervised negative-sampling loss (TEST INFRASTRUCTURE).

Follows /root/reference/metrics.py:
  * sampling  ``NegativeSamplingLoss.__get_node_indices`` (metrics.py:68-93): for every node of the batch, ALL of its walk
    co-occurrence neighbours when it has at most ``neg_num`` of them, else ``random.sample(neighbours, neg_num)`` (uniform,
    without replacement); the node is repeated once per kept neighbour; ``neg_num`` negatives are drawn ONCE per snapshot with
    ``random.sample(node_freqs, neg_num)`` — distinct POSITIONS of the frequency-expanded list (node ids may repeat);
  * loss      ``__negative_sampling_loss`` (metrics.py:38-66): per snapshot, with S = number of (node, positive) samples,
        pos_score_s = <e[node_s], e[pos_s]>,   neg_score_s = Σ_j <e[node_s], e[neg_j]> = <e[node_s], Σ_j e[neg_j]>,
        loss_t = mean_s softplus(−pos_score_s) + Q · mean_s softplus(neg_score_s)      (BCEWithLogits, mean reduction),
    summed over the snapshots; a snapshot with S = 0 contributes nothing.
The reference reseeds Python's RNG from the OS on every call (metrics.py:72), so its draws cannot be reproduced: parity is
pinned on the LOSS AND ITS GRADIENT FOR GIVEN INDICES (goldens record the reference's own draws) and on the sampling contract
(``check_sample``).
"""
from __future__ import annotations

import numpy as np


def softplus(v):
    return np.maximum(v, 0.0) + np.log1p(np.exp(-np.abs(v)))


def sigmoid(v):
    return 0.5 * (1.0 + np.tanh(0.5 * v))


def snapshot_loss(emb, node_idx, pos_idx, neg_idx, Q, dtype=np.float64):
    """loss_t and d loss_t / d emb for one snapshot (metrics.py:55-61)."""
    e = np.asarray(emb, dtype=dtype)
    grad = np.zeros_like(e)
    s = len(node_idx)
    if s == 0:
        return 0.0, grad
    node_idx, pos_idx, neg_idx = (np.asarray(a, dtype=np.int64) for a in (node_idx, pos_idx, neg_idx))
    negsum = e[neg_idx].sum(axis=0)
    en, ep = e[node_idx], e[pos_idx]
    pos = (en * ep).sum(axis=1)
    neg = en @ negsum
    loss = softplus(-pos).mean() + Q * softplus(neg).mean()
    gp = -sigmoid(-pos) / s                   # d loss / d pos_score
    gn = Q * sigmoid(neg) / s                 # d loss / d neg_score
    np.add.at(grad, node_idx, gp[:, None] * ep + gn[:, None] * negsum[None, :])
    np.add.at(grad, pos_idx, gp[:, None] * en)
    np.add.at(grad, neg_idx, np.broadcast_to((gn[:, None] * en).sum(axis=0), (len(neg_idx), e.shape[1])))
    return float(loss), grad


def neg_sampling_loss(emb_list, samples, Q, dtype=np.float64):
    """Σ_t loss_t (metrics.py:44-64).  samples[t] = (node_idx, pos_idx, neg_idx).  Returns (loss, [grad_t])."""
    total, grads = 0.0, []
    for emb, (ni, pi, gi) in zip(emb_list, samples):
        l, g = snapshot_loss(emb, ni, pi, gi, Q, dtype)
        total += l
        grads.append(g)
    return total, grads


def check_sample(batch, pair_ptr, pair_idx, freq, neg_num, pos, count, neg):
    """The sampling contract of metrics.py:74-88 on one draw in the padded form of ctgcn_neg_sample: pos [B, neg_num]
    (-1 after count[b] entries), count [B], neg [neg_num] node ids.  Every batch node with deg ≤ neg_num keeps all its
    neighbours in stored order, every other node neg_num DISTINCT neighbours; the negatives are ids of the frequency list."""
    batch, pos, count, neg = (np.asarray(a, dtype=np.int64) for a in (batch, pos, count, neg))
    freq = np.asarray(freq, dtype=np.int64)
    assert pos.shape == (len(batch), neg_num) and count.shape == (len(batch),)
    for b, node in enumerate(batch):
        nb = np.asarray(pair_idx[pair_ptr[node]:pair_ptr[node + 1]], dtype=np.int64)
        take = min(len(nb), neg_num)
        assert count[b] == take, ("count", int(node), int(count[b]), take)
        got = pos[b, :take]
        assert (pos[b, take:] == -1).all(), ("padding", int(node))
        if len(nb) <= neg_num:
            assert (got == nb).all(), ("all neighbours expected", int(node))
        else:
            assert len(np.unique(got)) == take, ("distinct positives expected", int(node))
            assert np.isin(got, nb).all(), ("positives must be neighbours", int(node))
    assert neg.shape == (neg_num,) and np.isin(neg, freq).all(), "negatives: ids of the frequency list"


def to_padded(batch, node_idx, pos_idx, neg_num):
    """The reference's flat (node_indices, pos_indices) of metrics.py:90-92 → padded (pos [B, neg_num], count [B])."""
    batch = np.asarray(batch, dtype=np.int64)
    pos = np.full((len(batch), neg_num), -1, dtype=np.int32)
    count = np.zeros(len(batch), dtype=np.int32)
    off = 0
    node_idx = np.asarray(node_idx, dtype=np.int64)
    for b, node in enumerate(batch):
        c = 0
        while off + c < len(node_idx) and node_idx[off + c] == node and c < neg_num:
            c += 1
        pos[b, :c] = np.asarray(pos_idx[off:off + c], dtype=np.int32)
        count[b] = c
        off += c
    assert off == len(node_idx), "node_idx is not grouped in batch order"
    return pos, count


def from_padded(batch, pos, count):
    """Inverse of to_padded: flat (node_idx, pos_idx)."""
    batch, pos, count = np.asarray(batch, dtype=np.int64), np.asarray(pos, dtype=np.int64), np.asarray(count, dtype=np.int64)
    node_idx = np.repeat(batch, count)
    pos_idx = np.concatenate([pos[b, :c] for b, c in enumerate(count)]) if len(batch) else np.zeros(0, dtype=np.int64)
    return node_idx, pos_idx


# ---- bit-level restatement of the device sampler (ctgcn_b200/csrc/loss.cu): the GPU draw must equal it exactly
_M64 = (1 << 64) - 1


def _mix64(z):
    z &= _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def rnd_below(seed, stream, draw, bound):
    r = _mix64((_mix64((seed ^ ((stream * 0x9E3779B97F4A7C15) & _M64)) & _M64) + 0xD1B54A32D192ED03 * (draw + 1)) & _M64)
    return (r * bound) >> 64


def floyd(n, m, seed, stream):
    """Floyd's subset sampling as the kernel runs it: a uniformly random m-subset of range(n), in draw order."""
    chosen = []
    for c, j in enumerate(range(n - m, n)):
        t = rnd_below(seed, stream, c, j + 1)
        if t in chosen:
            t = j
        chosen.append(t)
    return chosen


def device_sample(batch, pair_ptr, pair_idx, freq, neg_num, seed):
    """What ctgcn_neg_sample returns for these inputs: (pos [B, neg_num], count [B], neg [neg_num])."""
    seed &= _M64
    batch = np.asarray(batch, dtype=np.int64)
    pos = np.full((len(batch), neg_num), -1, dtype=np.int32)
    count = np.zeros(len(batch), dtype=np.int32)
    n_nodes = len(pair_ptr) - 1
    for b, node in enumerate(batch):
        if not 0 <= node < n_nodes:
            continue
        start, deg = int(pair_ptr[node]), int(pair_ptr[node + 1] - pair_ptr[node])
        if deg <= neg_num:
            pos[b, :deg] = pair_idx[start:start + deg]
            count[b] = deg
        else:
            pos[b] = [pair_idx[start + t] for t in floyd(deg, neg_num, seed, b)]
            count[b] = neg_num
    freq = np.asarray(freq)
    where = list(range(neg_num)) if len(freq) == neg_num else floyd(len(freq), neg_num, seed, _M64)
    return pos, count, freq[where].astype(np.int32)
