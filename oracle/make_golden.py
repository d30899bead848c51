#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
For every case it
  1. builds inputs deterministically (oracle/cases.py),
  2. loads the weights into the reference ``layers`` / ``models`` classes with
     ``load_state_dict(strict=True)`` (this also pins the state_dict key names/shapes),
  3. runs the reference forward on CPU fp32,
  4. asserts both restatements (oracle_np fp64, oracle_torch fp32) agree with it,
  5. stores graph + metadata + reference outputs.
The input-contract restatement ``oracle_np.build_core_adj_list`` is pinned against the reference's
``helper.DataLoader.get_core_adj_list`` run on k-core .npz files written to a temp dir.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import networkx as nx  # noqa: E402
import torch  # noqa: E402

np.int = int      # compat shim for the reference plumbing (utils.py:161…) — hot path untouched
np.float = float
import layers as ref_layers  # noqa: E402  (reference)
import models as ref_models  # noqa: E402  (reference)
import helper as ref_helper  # noqa: E402  (reference)

from oracle import cases, oracle_np, oracle_torch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(2)


def tsd(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def core_mats_from_edges(n, edges, weights=None):
    """k-core adjacency matrices 1..k_max like preprocessing/structure_generation.py:32-56."""
    g = nx.Graph()
    g.add_nodes_from(range(n))
    if weights is None:
        g.add_edges_from(edges)
    else:
        g.add_weighted_edges_from([(a, b, w) for (a, b), w in zip(edges, weights)])
    g.remove_edges_from(nx.selfloop_edges(g))
    core = nx.core_number(g)
    kmax = max(core.values())
    mats = []
    for k in range(1, kmax + 1):
        sub = nx.k_core(g, k=k, core_number=core)
        sub.add_nodes_from(range(n))
        mats.append(sp.csr_matrix(nx.to_scipy_sparse_array(sub, nodelist=range(n), dtype=np.float64)))
    return mats, core


def gnp_edges(n, p, seed):
    rng = np.random.default_rng(seed)
    iu = np.triu_indices(n, 1)
    keep = rng.random(iu[0].shape[0]) < p
    return list(zip(iu[0][keep].tolist(), iu[1][keep].tolist()))


def chung_lu_edges(n, m, seed, exponent=2.3):
    """Power-law (Chung–Lu) simple undirected edge list: endpoints drawn ∝ (i+1)^(-1/(exponent-1))."""
    rng = np.random.default_rng(seed)
    w = (np.arange(n) + 1.0) ** (-1.0 / (exponent - 1.0))
    cdf = np.cumsum(w) / w.sum()
    u = np.searchsorted(cdf, rng.random(3 * m))
    v = np.searchsorted(cdf, rng.random(3 * m))
    perm = rng.permutation(n)
    seen, out = set(), []
    for a, b in zip(perm[u].tolist(), perm[v].tolist()):
        if a != b and (min(a, b), max(a, b)) not in seen:
            seen.add((min(a, b), max(a, b)))
            out.append((a, b))
            if len(out) == m:
                break
    return out


def powerlaw_list(n, m, seed, k_keep):
    mats, _ = core_mats_from_edges(n, chung_lu_edges(n, m, seed))
    mats = mats[max(0, len(mats) - k_keep):]
    adj, _ = oracle_np.build_core_adj_list(mats)
    return adj


def nested_list(n, p, seed, k_keep, weighted=False):
    edges = gnp_edges(n, p, seed)
    w = None
    if weighted:
        w = np.random.default_rng(seed + 7).uniform(0.5, 2.0, len(edges)).round(3).tolist()
    mats, _ = core_mats_from_edges(n, edges, w)
    mats = mats[max(0, len(mats) - k_keep):]
    adj, _ = oracle_np.build_core_adj_list(mats)
    return adj


def uci_snapshots():
    nodes = [l.strip() for l in open(os.path.join(REF, "data/uci/nodes_set/nodes.csv")).read().split("\n") if l.strip()]
    idx = {v: i for i, v in enumerate(nodes)}
    snaps = []
    base = os.path.join(REF, "data/uci/1.format")
    for f in sorted(os.listdir(base)):
        edges = []
        for line in open(os.path.join(base, f)).read().split("\n")[1:]:
            if not line.strip():
                continue
            a, b = line.split("\t")[:2]
            if idx[a] != idx[b]:
                edges.append((idx[a], idx[b]))
        snaps.append(core_mats_from_edges(len(nodes), edges)[0])
    return len(nodes), snaps


def check_input_contract(n, snaps):
    """Pin oracle_np.build_core_adj_list against helper.DataLoader.get_core_adj_list (helper.py:51-82)."""
    with tempfile.TemporaryDirectory() as tmp:
        for t, mats in enumerate(snaps):
            d = os.path.join(tmp, f"s{t:02d}")
            os.makedirs(d)
            w = len(str(len(mats)))
            for k, m in enumerate(mats, start=1):
                sp.save_npz(os.path.join(d, f"{k:0>{w}}.npz"), m)
        loader = ref_helper.DataLoader(list(range(n)), len(snaps))
        ref = loader.get_core_adj_list(tmp, 0, len(snaps), max_core=-1)
    mine, mc = [], -1
    for mats in snaps:
        adj, mc = oracle_np.build_core_adj_list(mats, mc)
        mine.append(adj)
    assert [len(a) for a in ref] == [len(a) for a in mine], ([len(a) for a in ref], [len(a) for a in mine])
    for a_ref, a_me in zip(ref, mine):
        for r, m in zip(a_ref, a_me):
            assert abs(r.to_dense().numpy() - m.toarray()).max() == 0
    print("input contract pinned: K_t =", [len(a) for a in mine])
    return mine


def coo_list(adj):
    return [oracle_torch.to_torch_coo(a) if sp.issparse(a) else a for a in adj]


def save(name, meta, arrays):
    meta = dict(meta, name=name)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"  wrote {name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def agree(tag, ref, np64, t32):
    e1, e2 = cases.relerr(np64, ref), cases.relerr(t32, ref)
    print(f"  {tag}: relL2 oracle_np(fp64) vs ref {e1:.2e} | oracle_torch vs ref {e2:.2e}")
    assert e1 < 5e-6 and e2 < 5e-6, (tag, e1, e2)


COT_SEED = 900


def ref_grads(mod, outs, extra_inputs=()):
    """Reference autograd: loss = Σ_j Σ (outs[j] ⊙ cotangent(COT_SEED + j)); returns {'grad::<param>': array} for every
    parameter that received a gradient (CoreDiffusion.linear never does, layers.py:46) and 'grad::x' for a dense input."""
    loss = 0.0
    for j, o in enumerate(outs):
        loss = loss + (o * torch.from_numpy(cases.cotangent(COT_SEED + j, tuple(o.shape)))).sum()
    mod.zero_grad()
    loss.backward()
    out = {}
    for name, prm in mod.named_parameters():
        if prm.grad is not None:
            out["grad::" + name] = prm.grad.numpy().astype(np.float32)
    for name, t in extra_inputs:
        out["grad::" + name] = t.grad.numpy().astype(np.float32)
    return out


# ----------------------------------------------------------------------------- case runners
def run_core_diffusion(name, adj, d_in, d_out, x_seed, w_seed, bias=True, raw_coo=None, rnn_type="GRU", grad=False):
    n = adj[0].shape[0]
    x = cases.features(x_seed, n, d_in)
    sd = cases.core_diffusion_params(np.random.default_rng(w_seed), "", d_in, d_out, bias, rnn_type)
    mod = ref_layers.CoreDiffusion(d_in, d_out, bias=bias, rnn_type=rnn_type)
    mod.load_state_dict(tsd(sd), strict=True)
    tadj = raw_coo if raw_coo is not None else coo_list(adj)
    with torch.no_grad():
        y = mod(torch.from_numpy(x), tadj).numpy()
    grads = {}
    if grad:
        xt = torch.from_numpy(x).requires_grad_(True)
        grads = ref_grads(mod, [mod(xt, tadj)], [("x", xt)])
    mats = adj
    agree(name, y, oracle_np.core_diffusion(x, mats, sd), oracle_torch.core_diffusion(torch.from_numpy(x), tadj, tsd(sd)).numpy())
    # [K,N,D] fp64 from the fp32-rounded edge weights the model actually sees (utils.py:93) — checks the SpMM kernel alone
    u = oracle_np.cumulative_core_sums(x.astype(np.float64), [sp.coo_matrix(m).astype(np.float32) for m in mats])
    arrays = cases.pack_graph(mats)
    arrays.update(y=y.astype(np.float32), u_sum=u.sum(axis=2).astype(np.float64))
    arrays.update(grads)
    save(name, dict(kind="core_diffusion", d_in=d_in, d_out=d_out, x_seed=x_seed, w_seed=w_seed, bias=bias,
                    rnn_type=rnn_type, cot_seed=COT_SEED), arrays)


def run_mlp(name, n, d_in, hid, d_out, layer_num, act, x_kind, x_seed, w_seed, bias=True, grad=False):
    sd = cases.mlp_params(np.random.default_rng(w_seed), "", d_in, hid, d_out, layer_num, bias)
    if x_kind == "dense":
        xs = cases.features(x_seed, n, d_in)
        xt = torch.from_numpy(xs)
        arrays = {}
    elif x_kind == "eye":
        xs = sp.eye(n, format="coo", dtype=np.float32)
        xt = oracle_torch.to_torch_coo(xs)
        arrays = {}
    else:  # random sparse with duplicates
        rng = np.random.default_rng(x_seed)
        nnz = 6 * n
        xs = sp.coo_matrix((rng.standard_normal(nnz).astype(np.float32),
                            (rng.integers(0, n, nnz), rng.integers(0, d_in, nnz))), shape=(n, d_in))
        xt = oracle_torch.to_torch_coo(xs)
        arrays = {"xr": xs.row.astype(np.int32), "xc": xs.col.astype(np.int32), "xv": xs.data}
    mod = ref_layers.MLP(d_in, hid, d_out, layer_num, bias=bias, activate_type=act)
    mod.load_state_dict(tsd(sd), strict=True)
    with torch.no_grad():
        y = mod(xt).numpy()
    agree(name, y, oracle_np.mlp(xs, sd, "", layer_num, act), oracle_torch.mlp(xt, tsd(sd), "", layer_num, act).numpy())
    arrays["y"] = y.astype(np.float32)
    if grad:
        if x_kind == "dense":
            xg = torch.from_numpy(xs).requires_grad_(True)
            arrays.update(ref_grads(mod, [mod(xg)], [("x", xg)]))
        else:
            arrays.update(ref_grads(mod, [mod(xt)]))
    save(name, dict(kind="mlp", cot_seed=COT_SEED, n=n, d_in=d_in, hid=hid, d_out=d_out, layer_num=layer_num, act=act, x_kind=x_kind,
                    x_seed=x_seed, w_seed=w_seed, bias=bias), arrays)


def run_cdn(name, adj, d_in, hid, d_out, diffusion_num, x_seed, w_seed, rnn_type="GRU", grad=False):
    n = adj[0].shape[0]
    x = cases.features(x_seed, n, d_in)
    sd = cases.cdn_params(np.random.default_rng(w_seed), "", d_in, hid, d_out, diffusion_num, rnn_type=rnn_type)
    mod = ref_models.CDN(d_in, hid, d_out, diffusion_num, rnn_type=rnn_type)
    mod.load_state_dict(tsd(sd), strict=True)
    tadj = coo_list(adj)
    with torch.no_grad():
        y = mod(torch.from_numpy(x), tadj).numpy()
    agree(name, y, oracle_np.cdn(x, adj, sd, "", diffusion_num), oracle_torch.cdn(torch.from_numpy(x), tadj, tsd(sd), "", diffusion_num).numpy())
    arrays = cases.pack_graph(adj)
    arrays["y"] = y.astype(np.float32)
    if grad:
        xg = torch.from_numpy(x).requires_grad_(True)
        arrays.update(ref_grads(mod, [mod(xg, tadj)], [("x", xg)]))
    save(name, dict(kind="cdn", d_in=d_in, hid=hid, d_out=d_out, diffusion_num=diffusion_num, x_seed=x_seed, w_seed=w_seed,
                    rnn_type=rnn_type, cot_seed=COT_SEED), arrays)


def model_inputs(n, T, x_kind, d_in, x_seed):
    if x_kind == "eye":
        xs = [sp.eye(n, format="coo", dtype=np.float32) for _ in range(T)]
        return xs, [oracle_torch.to_torch_coo(x) for x in xs]
    xs = [cases.features(x_seed + t, n, d_in) for t in range(T)]
    return xs, [torch.from_numpy(x) for x in xs]


def run_model(name, cls, adj_lists, d_in, hid, d_out, trans_num, diffusion_num, model_type, act, x_kind, x_seed, w_seed,
              row_stride=1, single=False, rnn_type="GRU", grad=False, xs_override=None):
    T = len(adj_lists)
    n = adj_lists[0][0].shape[0]
    xs, xt = model_inputs(n, T, x_kind, d_in, x_seed)
    if xs_override is not None:
        xs, xt = xs_override, [torch.from_numpy(x) for x in xs_override]
    rng = np.random.default_rng(w_seed)
    tadj = [coo_list(a) for a in adj_lists]
    if cls == "ctgcn":
        sd = cases.ctgcn_params(rng, d_in, hid, d_out, trans_num, diffusion_num, T, model_type, rnn_type=rnn_type)
        mod = ref_models.CTGCN(d_in, hid, d_out, trans_num, diffusion_num, T, rnn_type=rnn_type, model_type=model_type,
                               trans_activate_type=act)
        mod.load_state_dict(tsd(sd), strict=True)
        with torch.no_grad():
            res = mod(xt, tadj)
        o_np = oracle_np.ctgcn(xs, adj_lists, sd, trans_num, diffusion_num, model_type, act)
        o_t = oracle_torch.ctgcn(xt, tadj, tsd(sd), trans_num, diffusion_num, model_type, act)
    else:
        sd = cases.cgcn_params(rng, d_in, hid, d_out, trans_num, diffusion_num, model_type, rnn_type=rnn_type)
        mod = ref_models.CGCN(d_in, hid, d_out, trans_num, diffusion_num, rnn_type=rnn_type, model_type=model_type,
                              trans_activate_type=act)
        mod.load_state_dict(tsd(sd), strict=True)
        args = (xt[0], tadj[0]) if single else (xt, tadj)
        with torch.no_grad():
            res = mod(*args)
        o_np = oracle_np.cgcn(xs[0] if single else xs, adj_lists[0] if single else adj_lists, sd, trans_num, diffusion_num, model_type, act)
        o_t = oracle_torch.cgcn(*args, tsd(sd), trans_num, diffusion_num, model_type, act)

    def split(r):
        if model_type == "S":
            out, trans = r
        else:
            out, trans = r, None
        if isinstance(out, (list, tuple)):
            out = np.stack([np.asarray(o) for o in out])
        out = np.asarray(out)
        if out.ndim == 2:
            out = out[None]
        if trans is not None:
            trans = np.stack([np.asarray(t) for t in trans]) if isinstance(trans, (list, tuple)) else np.asarray(trans)[None]
        return out, trans

    y, tr = split(res)
    y_np, tr_np = split(o_np)
    y_t, tr_t = split(o_t)
    agree(name, y, y_np, y_t)
    if tr is not None:
        agree(name + ".trans", tr, tr_np, tr_t)
    arrays = {}
    for t, a in enumerate(adj_lists):
        arrays.update({f"g{t}_{k}": v for k, v in cases.pack_graph(a).items()})
    arrays["y"] = np.ascontiguousarray(y[:, ::row_stride]).astype(np.float32)
    if tr is not None:
        arrays["trans"] = np.ascontiguousarray(tr[:, ::row_stride]).astype(np.float32)
    if grad:   # loss over the full (un-strided) outputs: cotangent COT_SEED on out [T,N,D], COT_SEED+1 on the stacked MLP outputs
        res_g = mod(*((xt[0], tadj[0]) if (cls == "cgcn" and single) else (xt, tadj)))
        out_g, tr_g = (res_g if model_type == "S" else (res_g, None))
        stk = lambda v: torch.stack(list(v)) if isinstance(v, (list, tuple)) else (v if v.dim() == 3 else v[None])
        arrays.update(ref_grads(mod, [stk(out_g)] + ([stk(tr_g)] if tr_g is not None else [])))
    save(name, dict(kind=cls, rnn_type=rnn_type, cot_seed=COT_SEED, n=n, T=T, d_in=d_in, hid=hid, d_out=d_out, trans_num=trans_num, diffusion_num=diffusion_num,
                    model_type=model_type, act=act, x_kind=x_kind, x_seed=x_seed, w_seed=w_seed, row_stride=row_stride,
                    single=single, state_dict_keys=sorted(sd.keys())), arrays)


def general_list(n=97):
    """Non-nested list with per-core weights, duplicate COO entries, an empty A_i and isolated nodes."""
    rng = np.random.default_rng(5)
    mats = []
    for i in range(5):
        if i == 2:
            mats.append(sp.coo_matrix((n, n), dtype=np.float32))
            continue
        nnz = 300 + 40 * i
        r = rng.integers(0, n - 7, nnz)        # last 7 nodes never appear as rows
        c = rng.integers(0, n, nnz)
        v = rng.uniform(-1, 1, nnz).round(2).astype(np.float32)
        r = np.concatenate([r, r[:25]])        # explicit duplicates (summed by torch.sparse.mm)
        c = np.concatenate([c, c[:25]])
        v = np.concatenate([v, v[:25]])
        mats.append(sp.coo_matrix((v, (r, c)), shape=(n, n)))
    return mats


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)

    print("== UCI input contract")
    n_uci, snaps = uci_snapshots()
    uci_lists = check_input_contract(n_uci, snaps)            # sticky max_core: K_t = 8,8,6,5,4,3,2
    uci_fresh_05, _ = oracle_np.build_core_adj_list(snaps[1])  # 2004-05 with a fresh max_core → K = 16

    print("== CoreDiffusion")
    for k in (1, 2, 5, 16):
        adj = nested_list(257, 0.14, 10 + k, k)
        run_core_diffusion(f"cd_nested_k{k}", adj, 128, 128, 100 + k, 200 + k)
    run_core_diffusion("cd_nested_weighted", nested_list(150, 0.1, 3, 6, weighted=True), 48, 40, 31, 32)
    run_core_diffusion("cd_nobias", nested_list(120, 0.1, 4, 4), 32, 64, 33, 34, bias=False)
    run_core_diffusion("cd_uci_0404_500_128", uci_lists[0], 500, 128, 41, 42)
    run_core_diffusion("cd_uci_0405_k16", uci_fresh_05, 64, 128, 43, 44)
    gen = general_list()
    raw = [torch.sparse_coo_tensor(torch.from_numpy(np.vstack((m.row, m.col))).long(), torch.from_numpy(m.data).float(),
                                   torch.Size(m.shape)) for m in gen]
    run_core_diffusion("cd_general", gen, 20, 24, 45, 46, raw_coo=raw)

    print("== MLP")
    run_mlp("mlp_1L_eye", 300, 300, 64, 96, 1, "L", "eye", 51, 52)
    run_mlp("mlp_3N_dense", 211, 70, 500, 128, 3, "N", "dense", 53, 54)
    run_mlp("mlp_3N_sparse", 211, 90, 64, 32, 3, "N", "sparse", 55, 56)
    run_mlp("mlp_1N_dense_nobias", 130, 33, 8, 17, 1, "N", "dense", 57, 58, bias=False)
    run_mlp("mlp_2L_dense", 64, 128, 128, 128, 2, "L", "dense", 59, 60)

    print("== CDN")
    adj5 = nested_list(257, 0.14, 15, 5)
    run_cdn("cdn_1layer", adj5, 64, 32, 48, 1, 61, 62)
    run_cdn("cdn_2layer", adj5, 96, 128, 64, 2, 63, 64)
    run_cdn("cdn_3layer", nested_list(100, 0.12, 16, 3), 16, 24, 8, 3, 65, 66)

    print("== CGCN / CTGCN")
    syn = [nested_list(180, 0.12, 70 + t, 4) for t in range(3)]
    run_model("cgcn_C_list", "cgcn", syn, 180, 64, 32, 1, 2, "C", "L", "eye", 71, 72)
    run_model("cgcn_S_single", "cgcn", syn[:1], 24, 64, 32, 3, 1, "S", "N", "dense", 73, 74, single=True)
    run_model("cgcn_S_list", "cgcn", syn[:2], 24, 64, 32, 3, 1, "S", "N", "dense", 75, 76)
    run_model("ctgcn_C_T1", "ctgcn", syn[:1], 180, 64, 32, 1, 2, "C", "L", "eye", 77, 78)
    run_model("ctgcn_C_T3", "ctgcn", syn, 180, 64, 32, 1, 2, "C", "L", "eye", 79, 80)
    run_model("ctgcn_S_T3", "ctgcn", syn, 40, 64, 32, 3, 1, "S", "N", "dense", 81, 82)
    run_model("ctgcn_128d_T2", "ctgcn", [nested_list(300, 0.1, 90 + t, 5) for t in range(2)], 128, 128, 128, 1, 1, "C", "L", "dense", 83, 84)
    run_model("ctgcn_C_uci_T7", "ctgcn", uci_lists, n_uci, 64, 32, 1, 2, "C", "L", "eye", 85, 86, row_stride=4)
    run_model("ctgcn_C_uci_T2_500_128", "ctgcn", uci_lists[:2], n_uci, 500, 128, 1, 2, "C", "L", "eye", 87, 88, row_stride=6)
    run_model("ctgcn_S_uci_T7", "ctgcn", uci_lists, 50, 64, 32, 3, 1, "S", "N", "dense", 89, 90, row_stride=4)
    print("== rnn_type='LSTM' (layers.py:27-28, models.py:234-235)")
    run_core_diffusion("cd_lstm_k5", nested_list(257, 0.14, 15, 5), 128, 128, 105, 205, rnn_type="LSTM", grad=True)
    run_core_diffusion("cd_lstm_nobias", nested_list(120, 0.1, 4, 4), 32, 64, 33, 34, bias=False, rnn_type="LSTM")
    run_core_diffusion("cd_lstm_uci_500_128", uci_lists[0], 500, 128, 41, 42, rnn_type="LSTM")
    run_model("ctgcn_C_T3_lstm", "ctgcn", syn, 180, 64, 32, 1, 2, "C", "L", "eye", 79, 80, rnn_type="LSTM", grad=True)
    run_model("ctgcn_S_T3_lstm", "ctgcn", syn, 40, 64, 32, 3, 1, "S", "N", "dense", 81, 82, rnn_type="LSTM")
    run_model("cgcn_C_list_lstm", "cgcn", syn, 180, 64, 32, 1, 2, "C", "L", "eye", 71, 72, rnn_type="LSTM")

    print("== reference gradients (autograd through the reference modules)")
    run_core_diffusion("cd_grad_k5", nested_list(200, 0.12, 21, 5), 48, 64, 111, 211, grad=True)
    run_core_diffusion("cd_grad_weighted", nested_list(150, 0.1, 3, 6, weighted=True), 48, 40, 31, 32, grad=True)
    run_core_diffusion("cd_grad_nobias", nested_list(120, 0.1, 4, 4), 32, 64, 33, 34, bias=False, grad=True)
    run_core_diffusion("cd_grad_general", gen, 20, 24, 45, 46, raw_coo=raw, grad=True)     # non-symmetric list
    run_core_diffusion("cd_grad_128", nested_list(257, 0.14, 15, 5), 128, 128, 105, 205, grad=True)
    run_mlp("mlp_grad_3N_dense", 211, 70, 96, 48, 3, "N", "dense", 53, 54, grad=True)
    run_mlp("mlp_grad_3N_sparse", 211, 90, 64, 32, 3, "N", "sparse", 55, 56, grad=True)
    run_mlp("mlp_grad_1L_eye", 150, 150, 64, 48, 1, "L", "eye", 51, 52, grad=True)
    run_mlp("mlp_grad_1N_nobias", 130, 33, 8, 17, 1, "N", "dense", 57, 58, bias=False, grad=True)
    run_cdn("cdn_grad_2layer", adj5, 48, 64, 32, 2, 63, 64, grad=True)
    run_model("ctgcn_grad_C_T3", "ctgcn", syn, 180, 64, 32, 1, 2, "C", "L", "eye", 79, 80, grad=True)
    run_model("ctgcn_grad_S_T3", "ctgcn", syn, 40, 64, 32, 3, 1, "S", "N", "dense", 81, 82, grad=True)
    run_model("cgcn_grad_S_list", "cgcn", syn[:2], 24, 64, 32, 3, 1, "S", "N", "dense", 75, 76, grad=True)
    run_model("ctgcn_grad_128d_T2", "ctgcn", [nested_list(300, 0.1, 90 + t, 5) for t in range(2)], 128, 128, 128, 1, 1, "C", "L",
              "dense", 83, 84, grad=True, row_stride=3)

    print("== reduced stand-ins of BASELINE.json configs[2] (Facebook-like CTGCN-S, T=12) and configs[4] (power-law, 256-d)")
    fb = [powerlaw_list(500, 2500, 300 + t, 9) for t in range(12)]
    deg_w = int(max(sp.csr_matrix(a[-1]).sum(axis=1).max() for a in fb)) + 1       # max degree + 1 (helper.py:128-135)
    fb_x = [cases.degree_gaussian_features(sp.csr_matrix(a[-1]), deg_w, 1e-4, 310 + t) for t, a in enumerate(fb)]
    run_model("ctgcn_S_fb_T12", "ctgcn", fb, deg_w, 500, 128, 3, 1, "S", "N", "degree_gaussian", 310, 311, row_stride=5,
              xs_override=fb_x)
    pl = powerlaw_list(420, 9000, 320, 20)
    run_core_diffusion("cd_powerlaw_256", pl, 256, 256, 321, 322)
    run_model("ctgcn_256d_T2_powerlaw", "ctgcn", [pl, powerlaw_list(420, 9000, 323, 20)], 256, 256, 256, 1, 1, "C", "L", "dense",
              324, 325, row_stride=4)
    print("done")


if __name__ == "__main__":
    main()
