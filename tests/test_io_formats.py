"""Input / output formats around the hot path, CPU only: the C++ k-core peeling and the one-pass k-core list
builder against networkx + the restated loader contract, the on-disk file naming, and the embedding export against pandas."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import oracle_np

nx = pytest.importorskip("networkx")


def _graph(n, m, seed, weighted=False):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, n, m)
    v = rng.integers(0, n, m)
    keep = u != v
    lo, hi = np.minimum(u, v)[keep], np.maximum(u, v)[keep]
    key = np.unique(lo * n + hi)
    u, v = key // n, key % n
    w = rng.uniform(0.5, 2.0, u.shape[0]).round(3).astype(np.float32) if weighted else np.ones(u.shape[0], dtype=np.float32)
    return u, v, w


def _nx_core_mats(n, u, v, w):
    """preprocessing/structure_generation.py:32-56 with networkx (what the reference writes to disk)."""
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_weighted_edges_from(zip(u.tolist(), v.tolist(), w.tolist()))
    core = nx.core_number(g)
    mats = []
    for k in range(1, max(core.values()) + 1):
        sub = nx.k_core(g, k=k, core_number=core)
        sub.add_nodes_from(range(n))
        mats.append(sp.csr_matrix(nx.to_scipy_sparse_array(sub, nodelist=range(n), dtype=np.float64)))
    return mats, core


@pytest.mark.parametrize("n,m,seed,weighted", [(60, 200, 0, False), (200, 1500, 1, True), (300, 700, 2, False), (40, 30, 3, False)])
def test_kcore_matrices_match_networkx(n, m, seed, weighted, lib):
    from ctgcn_b200 import io
    u, v, w = _graph(n, m, seed, weighted)
    mine = io.kcore_matrices(n, u, v, w)
    ref, core = _nx_core_mats(n, u, v, w)
    got_core, _ = io.edge_core_levels(n, u, v)
    assert got_core.tolist() == [core[i] for i in range(n)]
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert abs(a - b).max() < 1e-6


@pytest.mark.parametrize("max_core", [-1, 2, 4, 50])
@pytest.mark.parametrize("n,m,seed,weighted", [(60, 200, 0, False), (200, 1500, 1, True), (300, 700, 2, False)])
def test_snapshot_from_graph_equals_loader_contract(n, m, seed, weighted, max_core, lib):
    """One pass over the edges == k-core files → helper.get_core_adj_list (restated in oracle_np.build_core_adj_list, which is
    pinned against the reference loader in oracle/make_golden.py) — including dropped duplicate levels and sticky max_core."""
    from ctgcn_b200 import io
    u, v, w = _graph(n, m, seed, weighted)
    ref_mats, _ = _nx_core_mats(n, u, v, w)
    want, mc_want = oracle_np.build_core_adj_list(ref_mats, max_core)
    mine, mc_mine = io.select_core_list(io.kcore_matrices(n, u, v, w), max_core)
    snap, mc_snap = io.snapshot_from_graph(n, u, v, w, max_core)
    assert mc_want == mc_mine == mc_snap
    got = snap.coo_list()
    assert len(want) == len(mine) == len(got) == snap.k
    for a, b, c in zip(want, mine, got):
        dense = torch.sparse_coo_tensor(c._indices(), c._values(), c.shape).to_dense().numpy()
        assert abs(a.toarray() - b.toarray()).max() < 1e-6
        assert abs(a.toarray() - dense).max() < 1e-6
    assert snap.nnz_per_core == [int(a.nnz) for a in want]
    assert snap.edges_aggregated == sum(int(a.nnz) for a in want)


def test_kcore_files_and_loader_roundtrip(tmp_path, lib):
    """Edge CSVs → preprocess_kcores → load_core_adj_list: file names {k:0>w}.npz, last-duplicate-wins, self-loops dropped,
    max_core = -1 sticks to the first snapshot's file count (helper.py:61-62)."""
    from ctgcn_b200 import io
    nodes = [f"U{i}" for i in range(30)]
    (tmp_path / "nodes.csv").write_text("\n".join(nodes) + "\n")
    origin, core = tmp_path / "1.format", tmp_path / "2.core"
    origin.mkdir()
    rng = np.random.default_rng(0)
    specs = {"2004-04.csv": 60, "2004-05.csv": 200}
    for name, m in specs.items():
        lines = ["from_id\tto_id\tweight"]
        for _ in range(m):
            a, b = rng.integers(0, 30, 2)
            lines.append(f"U{a}\tU{b}\t{rng.integers(1, 4)}")
        lines.append(lines[1].rsplit("\t", 1)[0] + "\t9")            # repeated pair: the last weight wins
        (origin / name).write_text("\n".join(lines) + "\n")
    kmax = io.preprocess_kcores(str(origin), str(core), str(tmp_path / "nodes.csv"))
    assert sorted(kmax) == ["2004-04", "2004-05"] and kmax["2004-05"] > kmax["2004-04"]
    for stem, k in kmax.items():
        width = len(str(k))
        assert sorted(os.listdir(core / stem)) == [f"{i:0>{width}d}.npz" for i in range(1, k + 1)]
    index = {nm: i for i, nm in enumerate(nodes)}
    u, v, w = io.read_edge_csv(str(origin / "2004-04.csv"), index)
    assert (u < v).all() and len(set(zip(u.tolist(), v.tolist()))) == len(u)
    first = (origin / "2004-04.csv").read_text().split("\n")[1].split("\t")
    a, b = sorted((index[first[0]], index[first[1]]))
    if a != b:
        assert w[(u == a) & (v == b)][0] == 9.0
    adj = io.load_core_adj_list(str(core), 0, 2)
    assert len(adj) == 2
    assert (adj[0][0].diagonal() == 1).all() and adj[0][1].diagonal().sum() == 0      # +I on the first entry only
    # sticky max_core: the second snapshot uses only its first kmax["2004-04"] files
    mats2 = [sp.load_npz(str(core / "2004-05" / f)) for f in sorted(os.listdir(core / "2004-05"))]
    want2, _ = oracle_np.build_core_adj_list(mats2, kmax["2004-04"])
    assert len(adj[1]) == len(want2)
    for x, y in zip(adj[1], want2):
        assert abs(x - y).max() == 0


def test_save_embedding_matches_pandas(tmp_path, lib):
    pd = pytest.importorskip("pandas")
    from ctgcn_b200 import io
    nodes = [f"U{i}" for i in range(7)]
    out = torch.randn(2, 7, 5) * torch.tensor([1e-6, 1.0, 1e3, 1.0, 1.0])
    paths = io.save_embedding(out, str(tmp_path / "emb"), ["2004-04.csv", "2004-05.csv"], nodes)
    assert [os.path.basename(p) for p in paths] == ["2004-04.csv", "2004-05.csv"]
    for t, p in enumerate(paths):
        df = pd.read_csv(p, sep="\t", index_col=0)                   # how evaluation/link_prediction.py:228-233 reads it
        assert list(df.index) == nodes and list(df.columns) == [str(j) for j in range(5)]
        assert np.array_equal(df.values.astype(np.float32), out[t].numpy())
        ref = tmp_path / f"ref{t}.csv"
        pd.DataFrame(data=out[t].numpy(), index=nodes).to_csv(ref, sep="\t", header=True, index=True)   # embedding.py:86-88
        assert np.array_equal(pd.read_csv(ref, sep="\t", index_col=0).values.astype(np.float32), df.values.astype(np.float32))
