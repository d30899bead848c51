"""Data formats either side of the hot path: the reference's on-disk k-core files, the
``adj_list`` contract of its loader, and the embedding export — so that plans can be built straight from a dataset
folder (or from an edge list) without materialising K torch sparse COO tensors per snapshot.

Reference behaviour restated here (paths relative to the reference repo):
  utils.py:23-30      get_nx_graph: edge CSV ``from_id<sep>to_id[<sep>weight]`` with a header line, undirected, self-loops
                      dropped, a repeated pair keeps its LAST weight, weight 1.0 when the column is missing
  preprocessing/structure_generation.py:32-56
                      k-core files ``<core_folder>/<snapshot-stem>/{k:0>w}.npz`` for k = 1..k_max (w = digits of k_max,
                      utils.py:142-148), each the scipy CSR adjacency of the k-core over the FULL node list
  helper.py:51-82     get_core_adj_list: per snapshot ``sorted(files)[:max_core][::-1]`` (densest core first), ``+I`` on the
                      first entry, an entry equal to the previously loaded file is dropped, ``max_core == -1`` is replaced by
                      the first snapshot's file count and stays
  embedding.py:79-89  save_embedding: one ``<timestamp>.csv`` per snapshot, index = node names, header = column numbers

k-core numbers come from the C++ bucket peeling in libctgcn_b200.so (``ctgcn_kcore_numbers``) instead of
``networkx.core_number`` + one ``nx.k_core`` per level.  An edge (u, v) belongs to the k-core iff
min(core[u], core[v]) ≥ k, so all K nested matrices of a snapshot follow from one pass over its edges.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

from . import synth


# ----------------------------------------------------------------------------- edge lists
def read_node_list(path: str):
    """nodes.csv: one node name per line, no header (structure_generation.py:25-27)."""
    with open(path) as fh:
        return [ln.strip() for ln in fh.read().split("\n") if ln.strip()]


def read_edge_csv(path: str, node_index: dict, sep: str = "\t"):
    """One snapshot file → (u, v, w) with u < v, every undirected pair once (utils.py:23-30 semantics: header line, optional
    weight column (1.0 when missing), self-loops dropped, a repeated pair keeps its LAST weight).  Vectorised: the files of
    the larger datasets hold millions of lines."""
    import pandas as pd
    df = pd.read_csv(path, sep=sep, header=0, dtype=str, keep_default_na=False, skip_blank_lines=True)
    n = len(node_index)
    if df.shape[0] == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z.copy(), np.zeros(0, dtype=np.float32)
    a = df.iloc[:, 0].str.strip().map(node_index)
    b = df.iloc[:, 1].str.strip().map(node_index)
    if a.isna().any() or b.isna().any():
        bad = df.iloc[:, 0][a.isna()].tolist()[:1] + df.iloc[:, 1][b.isna()].tolist()[:1]
        raise KeyError(f"{path}: node {bad[0]!r} is not in the node list")
    a, b = a.to_numpy(dtype=np.int64), b.to_numpy(dtype=np.int64)
    if df.shape[1] > 2:
        wcol = df.iloc[:, 2].str.strip()
        w = np.where(wcol.to_numpy() == "", "1.0", wcol.to_numpy()).astype(np.float64)
    else:
        w = np.ones(a.shape[0], dtype=np.float64)
    keep = a != b
    lo, hi, w = np.minimum(a, b)[keep], np.maximum(a, b)[keep], w[keep]
    key = lo * n + hi
    _, first_in_reversed = np.unique(key[::-1], return_index=True)       # last occurrence of every pair
    last = key.shape[0] - 1 - first_in_reversed
    return lo[last], hi[last], w[last].astype(np.float32)


# ----------------------------------------------------------------------------- k-core structure
def edge_core_levels(n: int, u, v):
    """(core number per node, per-edge level = the largest k whose k-core contains the edge)."""
    core = synth.core_numbers(n, u, v) if len(u) else np.zeros(n, dtype=np.int32)
    return core, np.minimum(core[u], core[v]) if len(u) else np.zeros(0, dtype=np.int32)


def kcore_matrices(n: int, u, v, w):
    """The k-core adjacency matrices for k = 1..k_max, exactly the matrices structure_generation.py:47-56 saves."""
    _, ce = edge_core_levels(n, u, v)
    kmax = int(ce.max()) if ce.size else 0
    mats = []
    for k in range(1, kmax + 1):
        sel = ce >= k
        rows = np.concatenate([u[sel], v[sel]])
        cols = np.concatenate([v[sel], u[sel]])
        vals = np.concatenate([w[sel], w[sel]]).astype(np.float64)
        mats.append(sp.csr_matrix((vals, (rows, cols)), shape=(n, n)))
    return mats


def write_kcore_npz(output_dir: str, mats) -> list:
    """Save k-core matrices under the reference's file names ``{k:0>w}.npz`` (structure_generation.py:46,55-56)."""
    os.makedirs(output_dir, exist_ok=True)
    width = len(str(len(mats))) if mats else 1
    names = []
    for k, m in enumerate(mats, start=1):
        name = f"{k:0>{width}d}.npz"
        sp.save_npz(os.path.join(output_dir, name), sp.csr_matrix(m))
        names.append(name)
    return names


def preprocess_kcores(origin_dir: str, core_dir: str, node_file: str, sep: str = "\t") -> dict:
    """Drop-in for StructureInfoGenerator.get_kcore_graph_all_time (structure_generation.py:58-80): one k-core folder per
    snapshot file.  Returns {snapshot stem: k_max}."""
    nodes = read_node_list(node_file)
    index = {name: i for i, name in enumerate(nodes)}
    out = {}
    for f_name in sorted(os.listdir(origin_dir)):
        u, v, w = read_edge_csv(os.path.join(origin_dir, f_name), index, sep)
        mats = kcore_matrices(len(nodes), u, v, w)
        stem = f_name.split(".")[0]
        write_kcore_npz(os.path.join(core_dir, stem), mats)
        out[stem] = len(mats)
    return out


# ----------------------------------------------------------------------------- the loader contract (helper.py:51-82)
def select_core_list(core_mats, max_core: int = -1):
    """ONE snapshot's adj_list from its k-core matrices in file order (1-core first).  Returns (list, max_core used)."""
    if max_core == -1:
        max_core = len(core_mats)
    mats = list(core_mats)[:max_core][::-1]
    out, prev = [], None
    for j, m in enumerate(mats):
        m = sp.csr_matrix(m)
        if j == 0:
            out.append(sp.csr_matrix(m + sp.eye(m.shape[0])))
        elif (m - prev).sum() != 0:
            out.append(m)
        prev = m
    return out, max_core


def load_core_adj_list(core_base_path: str, start_idx: int, duration: int, max_core: int = -1, max_time_num: int = None):
    """helper.DataLoader.get_core_adj_list on scipy matrices: list (snapshots) of lists (cores, densest first)."""
    dirs = sorted(os.listdir(core_base_path))
    if start_idx >= len(dirs):
        raise AssertionError("start_idx beyond the number of snapshots")
    stop = min(start_idx + duration, len(dirs) if max_time_num is None else max_time_num)
    result = []
    for i in range(start_idx, stop):
        d = os.path.join(core_base_path, dirs[i])
        mats = [sp.load_npz(os.path.join(d, f)) for f in sorted(os.listdir(d))]
        adj, max_core = select_core_list(mats, max_core)
        result.append(adj)
    return result


def load_core_plans(core_base_path: str, start_idx: int, duration: int, device, max_core: int = -1, max_time_num: int = None):
    """The same list as graph plans on `device` (accepted by CTGCN.forward / CoreDiffusion.forward in place of the COO lists)."""
    from .plan import build_plan_coo
    return [build_plan_coo(adj, device) for adj in load_core_adj_list(core_base_path, start_idx, duration, max_core, max_time_num)]


def snapshot_from_graph(n: int, u, v, w=None, max_core: int = -1):
    """The loader contract evaluated directly on an edge list (no k-core files, no K matrices): returns
    (SnapshotGraph, max_core used).  Equivalent to kcore_matrices → select_core_list, in one pass over the edges."""
    u, v = np.asarray(u, dtype=np.int64), np.asarray(v, dtype=np.int64)
    w = np.ones(u.shape[0], dtype=np.float32) if w is None else np.asarray(w, dtype=np.float32)
    _, ce = edge_core_levels(n, u, v)
    kmax = int(ce.max()) if ce.size else 0
    if max_core == -1:
        max_core = kmax
    top = min(max_core, kmax)                              # files [:max_core] of 1..k_max, reversed: `top` comes first
    if top < 1:
        raise ValueError("snapshot has no k-core file to load (empty graph or max_core < 1)")
    present = np.zeros(top + 1, dtype=bool)
    present[np.minimum(ce, top)] = True                    # present[k]: some edge has level exactly k (levels > top clamp to top)
    # The first file (`top`) is always kept; a lower level k is kept iff its matrix differs from the (k+1)-core loaded just
    # before it (helper.py:73-76), i.e. iff some edge has level exactly k.  An edge of level k first appears in entry k.
    levels = [top] + [k for k in range(top - 1, 0, -1) if present[k]]
    lut = np.full(top + 1, 255, dtype=np.uint8)
    for idx, lev in enumerate(levels):
        lut[lev] = idx
    sel = ce >= 1
    le = lut[np.minimum(ce[sel], top)]
    ok = le != 255
    return synth.assemble_snapshot(n, u[sel][ok], v[sel][ok], w[sel][ok], le[ok], levels), max_core


# ----------------------------------------------------------------------------- embedding export (embedding.py:79-89)
def save_embedding(output_list, embedding_base_path: str, timestamp_list, full_node_list, start_idx: int = 0, sep: str = "\t"):
    """One ``<timestamp-stem>.csv`` per snapshot: header row = column numbers, index column = node names."""
    import torch
    if isinstance(output_list, torch.Tensor) and output_list.dim() == 2:
        output_list = [output_list]
    os.makedirs(embedding_base_path, exist_ok=True)
    paths = []
    for i in range(len(output_list)):
        emb = output_list[i]
        emb = emb.detach().cpu().numpy() if isinstance(emb, torch.Tensor) else np.asarray(emb)
        stem = str(timestamp_list[start_idx + i]).split(".")[0]
        path = os.path.join(embedding_base_path, stem + ".csv")
        with open(path, "w") as fh:
            fh.write(sep.join([""] + [str(j) for j in range(emb.shape[1])]) + "\n")
            for name, row in zip(full_node_list, emb):
                fh.write(sep.join([str(name)] + [str(x) for x in row.astype(np.float32)]) + "\n")
        paths.append(path)
    return paths
