"""numpy restatement of the CTGCN forward hot path (TEST INFRASTRUCTURE).

Every function cites the reference lines it restates (paths relative to
/root/reference).  Arithmetic is plain numpy in a caller-chosen dtype
(float64 by default: the tie-breaker when two fp32 paths disagree at 1e-6).

Weights are passed as a flat ``dict`` with the reference's ``state_dict`` key
names (``rnn.weight_ih_l0`` …) so that goldens, the reference modules and the
CUDA modules all share one parameter container.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def selu(v):
    """F.selu as used at layers.py:98-99,104-105."""
    return SELU_SCALE * np.where(v > 0, v, SELU_ALPHA * np.expm1(np.minimum(v, 0)))


def layer_norm(v, weight, bias, eps=1e-5):
    """nn.LayerNorm(H) over the last axis (biased variance) — layers.py:31,62; models.py:238,250."""
    mu = v.mean(axis=-1, keepdims=True)
    var = ((v - mu) ** 2).mean(axis=-1, keepdims=True)
    return (v - mu) / np.sqrt(var + eps) * weight + bias


def gru_sequence(seq, w_ih, w_hh, b_ih, b_hh):
    """Single-layer batch_first nn.GRU with h0 = 0, returning every step's output.

    seq: [B, L, D_in] → [B, L, H].  PyTorch gate packing [r; z; n]
    (layers.py:30,59 and models.py:237,249 both instantiate nn.GRU(num_layers=1,
    batch_first=True) and call it without an initial state).
    """
    B, L, _ = seq.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H), dtype=seq.dtype)
    out = np.empty((B, L, H), dtype=seq.dtype)
    if b_ih is None:
        b_ih = np.zeros(3 * H, dtype=seq.dtype)
        b_hh = np.zeros(3 * H, dtype=seq.dtype)
    for s in range(L):
        gi = seq[:, s, :] @ w_ih.T + b_ih
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, :H] + gh[:, :H])
        z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        out[:, s, :] = h
    return out


def lstm_sequence(seq, w_ih, w_hh, b_ih, b_hh):
    """Single-layer batch_first nn.LSTM with h0 = c0 = 0, returning every step's output h_s (layers.py:27-28,59;
    models.py:234-235,249 with rnn_type='LSTM').  PyTorch gate packing [i; f; g; o]:
    c' = f ⊙ c + i ⊙ g,  h' = o ⊙ tanh(c')."""
    B, L, _ = seq.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H), dtype=seq.dtype)
    c = np.zeros((B, H), dtype=seq.dtype)
    out = np.empty((B, L, H), dtype=seq.dtype)
    if b_ih is None:
        b_ih = np.zeros(4 * H, dtype=seq.dtype)
        b_hh = np.zeros(4 * H, dtype=seq.dtype)
    for s in range(L):
        pre = seq[:, s, :] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f = _sigmoid(pre[:, :H]), _sigmoid(pre[:, H:2 * H])
        g, o = np.tanh(pre[:, 2 * H:3 * H]), _sigmoid(pre[:, 3 * H:])
        c = f * c + i * g
        h = o * np.tanh(c)
        out[:, s, :] = h
    return out


def rnn_sequence(seq, sd, prefix, dtype):
    """Dispatch on the stored weight shape: [3H, ·] → GRU, [4H, ·] → LSTM (the reference's rnn_type)."""
    w_ih, w_hh = _p(sd, prefix, "rnn.weight_ih_l0", dtype), _p(sd, prefix, "rnn.weight_hh_l0", dtype)
    b_ih, b_hh = _p(sd, prefix, "rnn.bias_ih_l0", dtype), _p(sd, prefix, "rnn.bias_hh_l0", dtype)
    fn = lstm_sequence if w_hh.shape[0] == 4 * w_hh.shape[1] else gru_sequence
    return fn(seq, w_ih, w_hh, b_ih, b_hh)


def _p(sd, prefix, name, dtype):
    key = prefix + name
    return None if key not in sd else np.asarray(sd[key], dtype=dtype)


def cumulative_core_sums(x, adj_list):
    """S_i = S_{i-1} + A_i·x, then relu — layers.py:41-48.  Returns [K, N, D]."""
    outs = []
    acc = None
    for a in adj_list:
        a = sp.csr_matrix(a).astype(x.dtype)  # duplicates are summed, like torch.sparse.mm
        prod = a @ x
        acc = prod if acc is None else acc + prod
        outs.append(acc)
    return np.maximum(np.stack(outs, axis=0), 0)


def core_diffusion(x, adj_list, sd, prefix="", dtype=np.float64, eps=1e-5):
    """layers.CoreDiffusion.forward (layers.py:38-63); GRU or LSTM by the stored weight shapes.

    cumulative SpMM (:41-47) → relu (:48) → stack/transposed view [N,K,D] (:58)
    → GRU over the core axis (:59) → Σ over cores (:60) → LayerNorm (:62).
    ``linear.*`` parameters exist in the reference but are never read (:24,:46).
    """
    x = np.asarray(x, dtype=dtype)
    u = cumulative_core_sums(x, adj_list).transpose(1, 0, 2)
    hs = rnn_sequence(u, sd, prefix, dtype)
    o = hs.sum(axis=1)
    return layer_norm(o, _p(sd, prefix, "norm.weight", dtype), _p(sd, prefix, "norm.bias", dtype), eps)


def mlp(x, sd, prefix, layer_num, activate_type, dtype=np.float64):
    """layers.MLP.forward (layers.py:95-106).  x dense ndarray or scipy sparse.

    layer_num == 1 reads ``linear.*`` (:96-100); otherwise ``linears.{j}.*`` with
    selu after EVERY layer, the last included, iff activate_type == 'N' (:102-105).
    """
    def affine(h, wkey, bkey):
        w = _p(sd, prefix, wkey, dtype)
        b = _p(sd, prefix, bkey, dtype)
        h = (h @ w.T) if not sp.issparse(h) else np.asarray((h.astype(dtype) @ w.T))
        return h if b is None else h + b

    if not sp.issparse(x):
        x = np.asarray(x, dtype=dtype)
    if layer_num == 1:
        h = affine(x, "linear.weight", "linear.bias")
        return selu(h) if activate_type == "N" else h
    h = x
    for j in range(layer_num):
        h = affine(h, f"linears.{j}.weight", f"linears.{j}.bias")
        if activate_type == "N":
            h = selu(h)
    return h


def cdn(x, adj_list, sd, prefix, diffusion_num, dtype=np.float64):
    """models.CDN.forward (models.py:39-42): the same adj_list feeds every layer."""
    for l in range(diffusion_num):
        x = core_diffusion(x, adj_list, sd, f"{prefix}diffusion_list.{l}.", dtype)
    return x


def cgcn(x, adj, sd, trans_num, diffusion_num, model_type="C", trans_activate_type="L", dtype=np.float64):
    """models.CGCN.forward / .cgcn (models.py:165-187): one shared MLP + CDN per snapshot."""
    def one(xi, ai):
        trans = mlp(xi, sd, "mlp.", trans_num, trans_activate_type, dtype)
        emb = cdn(trans, ai, sd, "duffision.", diffusion_num, dtype)
        return (emb, trans) if model_type == "S" else emb

    if isinstance(x, list):
        res = [one(xi, ai) for xi, ai in zip(x, adj)]
        if model_type == "C":
            return res
        return [r[0] for r in res], [r[1] for r in res]
    return one(x, adj)


def ctgcn(x_list, adj_list, sd, trans_num, diffusion_num, model_type="C", trans_activate_type="L",
          dtype=np.float64, eps=1e-5):
    """models.CTGCN.forward (models.py:240-253).

    Per snapshot t: MLP_t (:244) → CDN_t (:246) with independent weights
    (``mlp_list.{t}.``, ``duffision_list.{t}.`` — :225-231); stack to [N,T,D]
    (:248); temporal GRU h0=0 (:249); LayerNorm then transpose → [T,N,D] (:250).
    Returns out (C) or (out, trans_list) (S) (:251-253).
    """
    hx, trans_list = [], []
    for t, (x, adj) in enumerate(zip(x_list, adj_list)):
        trans = mlp(x, sd, f"mlp_list.{t}.", trans_num, trans_activate_type, dtype)
        trans_list.append(trans)
        hx.append(cdn(trans, adj, sd, f"duffision_list.{t}.", diffusion_num, dtype))
    seq = np.stack(hx, axis=0).transpose(1, 0, 2)
    out = rnn_sequence(seq, sd, "", dtype)
    out = layer_norm(out, _p(sd, "", "norm.weight", dtype), _p(sd, "", "norm.bias", dtype), eps)
    out = out.transpose(1, 0, 2)
    return out if model_type == "C" else (out, trans_list)


def build_core_adj_list(core_mats, max_core=-1):
    """Input contract of helper.DataLoader.get_core_adj_list for ONE snapshot (helper.py:58-80).

    core_mats: the k-core adjacency matrices in file order (1-core first … k_max-core last),
    i.e. what ``sorted(os.listdir(date_dir_path))`` enumerates (:58).  Returns the list handed to
    the model: densest-index first (:63-64), ``+I`` on the first entry (:71-72), an entry is dropped
    when it equals the previously LOADED matrix in sum (:73-76), no normalisation (:77).
    Returns (adj_list, max_core_used) because ``max_core == -1`` sticks after the first snapshot (:61-62).
    """
    if max_core == -1:
        max_core = len(core_mats)
    mats = list(core_mats)[:max_core][::-1]
    out, prev = [], None
    for j, m in enumerate(mats):
        m = sp.csr_matrix(m)
        if j == 0:
            out.append(sp.csr_matrix(m + sp.eye(m.shape[0])))
        else:
            if (m - prev).sum() != 0:
                out.append(m)
        prev = m
    return out, max_core


def core_diffusion_rows(x, adj_list, sd, rows, prefix="", dtype=np.float64, eps=1e-5):
    """core_diffusion restricted to a subset of output rows (full-size spot checks): same arithmetic as
    layers.py:38-63, evaluated only for `rows` (A_i[rows] · x needs all of x but only |rows| sums)."""
    x = np.asarray(x, dtype=dtype)
    outs, acc = [], None
    for a in adj_list:
        prod = sp.csr_matrix(a).astype(dtype)[rows] @ x
        acc = prod if acc is None else acc + prod
        outs.append(acc)
    u = np.maximum(np.stack(outs, axis=1), 0)
    hs = rnn_sequence(u, sd, prefix, dtype)
    return layer_norm(hs.sum(axis=1), _p(sd, prefix, "norm.weight", dtype), _p(sd, prefix, "norm.bias", dtype), eps)
