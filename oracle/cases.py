"""Deterministic parity cases shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE).

A case = graph(s) + seeds + model hyper-parameters.  Graph structure is stored inside the golden
fixture (small integer arrays); features and weights are regenerated from numpy PCG64 seeds, which
are stable across numpy versions and machines, so the fixtures stay small.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------- parameters
def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def gru_params(rng, prefix, d_in, h, bias=True, rnn_type="GRU"):
    """nn.GRU / nn.LSTM(num_layers=1) parameters: gate-stacked [3H | 4H, ·] matrices."""
    b = 1.0 / np.sqrt(h)
    g = 4 if rnn_type == "LSTM" else 3
    sd = {prefix + "weight_ih_l0": _uniform(rng, (g * h, d_in), b),
          prefix + "weight_hh_l0": _uniform(rng, (g * h, h), b)}
    if bias:
        sd[prefix + "bias_ih_l0"] = _uniform(rng, (g * h,), b)
        sd[prefix + "bias_hh_l0"] = _uniform(rng, (g * h,), b)
    return sd


def linear_params(rng, prefix, d_in, d_out, bias=True):
    b = 1.0 / np.sqrt(d_in)
    sd = {prefix + "weight": _uniform(rng, (d_out, d_in), b)}
    if bias:
        sd[prefix + "bias"] = _uniform(rng, (d_out,), b)
    return sd


def norm_params(rng, prefix, h):
    return {prefix + "weight": (1.0 + 0.1 * rng.standard_normal(h)).astype(np.float32),
            prefix + "bias": (0.1 * rng.standard_normal(h)).astype(np.float32)}


def core_diffusion_params(rng, prefix, d_in, d_out, bias=True, rnn_type="GRU"):
    """state_dict of layers.CoreDiffusion (layers.py:24-31): linear (unused), rnn, norm."""
    sd = {}
    sd.update(linear_params(rng, prefix + "linear.", d_in, d_out))
    sd.update(gru_params(rng, prefix + "rnn.", d_in, d_out, bias, rnn_type))
    sd.update(norm_params(rng, prefix + "norm.", d_out))
    return sd


def mlp_params(rng, prefix, d_in, hid, d_out, layer_num, bias=True):
    """state_dict of layers.MLP (layers.py:86-93)."""
    if layer_num == 1:
        return linear_params(rng, prefix + "linear.", d_in, d_out, bias)
    sd = {}
    dims = [d_in] + [hid] * (layer_num - 1) + [d_out]
    for j in range(layer_num):
        sd.update(linear_params(rng, f"{prefix}linears.{j}.", dims[j], dims[j + 1], bias))
    return sd


def cdn_params(rng, prefix, d_in, hid, d_out, diffusion_num, bias=True, rnn_type="GRU"):
    """state_dict of models.CDN (models.py:25-33)."""
    sd = {}
    if diffusion_num == 1:
        dims = [d_in, d_out]
    else:
        dims = [d_in] + [hid] * (diffusion_num - 1) + [d_out]
    for l in range(diffusion_num):
        sd.update(core_diffusion_params(rng, f"{prefix}diffusion_list.{l}.", dims[l], dims[l + 1], bias, rnn_type))
    return sd


def cgcn_params(rng, d_in, hid, d_out, trans_num, diffusion_num, model_type, bias=True, rnn_type="GRU"):
    """state_dict of models.CGCN (models.py:157-163)."""
    sd = {}
    if model_type == "C":
        sd.update(mlp_params(rng, "mlp.", d_in, hid, hid, trans_num, bias))
        sd.update(cdn_params(rng, "duffision.", hid, d_out, d_out, diffusion_num, rnn_type=rnn_type))
    else:
        sd.update(mlp_params(rng, "mlp.", d_in, hid, d_out, trans_num, bias))
        sd.update(cdn_params(rng, "duffision.", d_out, d_out, d_out, diffusion_num, rnn_type=rnn_type))
    return sd


def ctgcn_params(rng, d_in, hid, d_out, trans_num, diffusion_num, duration, model_type, bias=True, rnn_type="GRU"):
    """state_dict of models.CTGCN (models.py:222-238)."""
    sd = {}
    for t in range(duration):
        if model_type == "C":
            sd.update(mlp_params(rng, f"mlp_list.{t}.", d_in, hid, hid, trans_num, bias))
            sd.update(cdn_params(rng, f"duffision_list.{t}.", hid, d_out, d_out, diffusion_num, rnn_type=rnn_type))
        else:
            sd.update(mlp_params(rng, f"mlp_list.{t}.", d_in, hid, d_out, trans_num, bias))
            sd.update(cdn_params(rng, f"duffision_list.{t}.", d_out, d_out, d_out, diffusion_num, rnn_type=rnn_type))
    sd.update(gru_params(rng, "rnn.", d_out, d_out, bias, rnn_type))
    sd.update(norm_params(rng, "norm.", d_out))
    return sd


# ----------------------------------------------------------------------------- graphs
def pack_graph(mats):
    """list of scipy matrices → dict of small arrays (COO order preserved, duplicates kept)."""
    out = {"n": np.int64(mats[0].shape[0]), "k": np.int64(len(mats))}
    for i, m in enumerate(mats):
        m = m.tocoo() if not isinstance(m, sp.coo_matrix) else m
        out[f"r{i}"] = m.row.astype(np.int32)
        out[f"c{i}"] = m.col.astype(np.int32)
        out[f"v{i}"] = m.data.astype(np.float32)
    return out


def unpack_graph(d, tag=""):
    n, k = int(d[tag + "n"]), int(d[tag + "k"])
    return [sp.coo_matrix((d[f"{tag}v{i}"], (d[f"{tag}r{i}"], d[f"{tag}c{i}"])), shape=(n, n)) for i in range(k)]


def features(seed, n, d):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def degree_gaussian_features(adj, width, std, seed):
    """'gaussian' degree features of helper.DataLoader.get_degree_feature_list (helper.py:128-135): row v is `width` draws of
    N(degree(v), std).  Regenerated from a PCG64 seed (the reference draws from numpy's global legacy stream)."""
    deg = np.asarray(adj.sum(axis=1)).reshape(-1, 1).astype(np.float64)
    noise = np.random.default_rng(seed).standard_normal((deg.shape[0], width))
    return (deg + std * noise).astype(np.float32)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ----------------------------------------------------------------------------- golden fixtures
import json as _json
import os as _os

GOLDEN_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "golden")


def golden_names(kind=None, rnn_type="GRU", grads=None):
    """Fixture names, filtered by case kind, recurrent cell (None = any) and presence of stored reference gradients."""
    names = sorted(f[:-4] for f in _os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
    out = []
    for n in names:
        meta = load_meta(n)
        if kind is not None and meta["kind"] != kind:
            continue
        if rnn_type is not None and meta["rnn_type"] != rnn_type:
            continue
        if grads is not None and bool(meta.get("has_grads")) != grads:
            continue
        out.append(n)
    return out


def load_meta(name):
    with np.load(_os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        meta = _json.loads(bytes(z["meta"]).decode())
        meta["has_grads"] = any(k.startswith("grad::") for k in z.files)
    meta.setdefault("rnn_type", "GRU")
    return meta


def load_case(name):
    """Rebuild the inputs of a golden case and return them with the stored reference outputs.

    keys: meta, sd (numpy state_dict), expected (dict of arrays) and, by kind,
      core_diffusion / cdn: x [N,D] float32, adj (list of scipy COO)
      mlp: x (ndarray or scipy COO)
      cgcn / ctgcn: x_list, adj_lists
    """
    z = dict(np.load(_os.path.join(GOLDEN_DIR, name + ".npz")))
    meta = _json.loads(bytes(z.pop("meta")).decode())
    meta.setdefault("rnn_type", "GRU")
    kind = meta["kind"]
    rt = meta["rnn_type"]
    rng = np.random.default_rng(meta["w_seed"])
    case = {"meta": meta, "expected": {}}
    if kind == "core_diffusion":
        adj = unpack_graph(z)
        n = adj[0].shape[0]
        case.update(adj=adj, x=features(meta["x_seed"], n, meta["d_in"]),
                    sd=core_diffusion_params(rng, "", meta["d_in"], meta["d_out"], meta["bias"], rt))
        case["expected"] = {"y": z["y"], "u_sum": z["u_sum"]}
    elif kind == "cdn":
        adj = unpack_graph(z)
        n = adj[0].shape[0]
        case.update(adj=adj, x=features(meta["x_seed"], n, meta["d_in"]),
                    sd=cdn_params(rng, "", meta["d_in"], meta["hid"], meta["d_out"], meta["diffusion_num"], rnn_type=rt))
        case["expected"] = {"y": z["y"]}
    elif kind == "mlp":
        n, d_in = meta["n"], meta["d_in"]
        if meta["x_kind"] == "dense":
            x = features(meta["x_seed"], n, d_in)
        elif meta["x_kind"] == "eye":
            x = sp.eye(n, format="coo", dtype=np.float32)
        else:
            x = sp.coo_matrix((z["xv"], (z["xr"], z["xc"])), shape=(n, d_in))
        case.update(x=x, sd=mlp_params(rng, "", d_in, meta["hid"], meta["d_out"], meta["layer_num"], meta["bias"]))
        case["expected"] = {"y": z["y"]}
    elif kind in ("cgcn", "ctgcn"):
        T, n = meta["T"], meta["n"]
        adj_lists = [unpack_graph(z, f"g{t}_") for t in range(T)]
        if meta["x_kind"] == "eye":
            x_list = [sp.eye(n, format="coo", dtype=np.float32) for _ in range(T)]
        elif meta["x_kind"] == "degree_gaussian":   # degrees of the sparsest (last) list entry = the snapshot graph
            x_list = [degree_gaussian_features(sp.csr_matrix(adj_lists[t][-1]), meta["d_in"], 1e-4, meta["x_seed"] + t)
                      for t in range(T)]
        else:
            x_list = [features(meta["x_seed"] + t, n, meta["d_in"]) for t in range(T)]
        if kind == "ctgcn":
            sd = ctgcn_params(rng, meta["d_in"], meta["hid"], meta["d_out"], meta["trans_num"], meta["diffusion_num"], T,
                              meta["model_type"], rnn_type=rt)
        else:
            sd = cgcn_params(rng, meta["d_in"], meta["hid"], meta["d_out"], meta["trans_num"], meta["diffusion_num"],
                             meta["model_type"], rnn_type=rt)
        case.update(x_list=x_list, adj_lists=adj_lists, sd=sd)
        case["expected"] = {"y": z["y"]}
        if "trans" in z:
            case["expected"]["trans"] = z["trans"]
    else:
        raise ValueError(kind)
    # reference autograd results (oracle/make_golden.py: loss = Σ out ⊙ cotangent(seed)), when the case stores them
    case["grads"] = {k[len("grad::"):]: v for k, v in z.items() if k.startswith("grad::")}
    return case


def cotangent(seed, shape):
    """The fixed upstream gradient dL/d(out) used for the stored reference gradients."""
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
