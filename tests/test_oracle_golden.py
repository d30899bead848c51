"""The oracle restatements against the golden vectors generated from the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cases, oracle_np, oracle_torch

TOL = 5e-6  # fp32 reference vs fp64 restatement / same-op torch port


def _tsd(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def _coo(mats):
    return [oracle_torch.to_torch_coo(m) for m in mats]


@pytest.mark.parametrize("name", cases.golden_names("core_diffusion", rnn_type=None))
def test_core_diffusion(name):
    c = cases.load_case(name)
    y64 = oracle_np.core_diffusion(c["x"], c["adj"], c["sd"])
    assert cases.relerr(y64, c["expected"]["y"]) < TOL
    if c["adj"][0].shape[0] <= 400:
        yt = oracle_torch.core_diffusion(torch.from_numpy(c["x"]), _coo(c["adj"]), _tsd(c["sd"])).numpy()
        assert cases.relerr(yt, c["expected"]["y"]) < TOL
    u = oracle_np.cumulative_core_sums(c["x"].astype(np.float64), c["adj"])
    np.testing.assert_allclose(u.sum(axis=2), c["expected"]["u_sum"], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("name", cases.golden_names("mlp", rnn_type=None))
def test_mlp(name):
    c = cases.load_case(name)
    m = c["meta"]
    y64 = oracle_np.mlp(c["x"], c["sd"], "", m["layer_num"], m["act"])
    assert cases.relerr(y64, c["expected"]["y"]) < TOL
    xt = torch.from_numpy(c["x"]) if isinstance(c["x"], np.ndarray) else oracle_torch.to_torch_coo(c["x"])
    yt = oracle_torch.mlp(xt, _tsd(c["sd"]), "", m["layer_num"], m["act"]).numpy()
    assert cases.relerr(yt, c["expected"]["y"]) < TOL


@pytest.mark.parametrize("name", cases.golden_names("cdn", rnn_type=None))
def test_cdn(name):
    c = cases.load_case(name)
    y64 = oracle_np.cdn(c["x"], c["adj"], c["sd"], "", c["meta"]["diffusion_num"])
    assert cases.relerr(y64, c["expected"]["y"]) < TOL


def _model_out(c, res):
    m = c["meta"]
    out, trans = res if m["model_type"] == "S" else (res, None)
    out = np.stack([np.asarray(o) for o in out]) if isinstance(out, (list, tuple)) else np.asarray(out)
    if out.ndim == 2:
        out = out[None]
    if trans is not None:
        trans = np.stack([np.asarray(t) for t in trans]) if isinstance(trans, (list, tuple)) else np.asarray(trans)[None]
    rs = m["row_stride"]
    return out[:, ::rs], None if trans is None else trans[:, ::rs]


@pytest.mark.parametrize("name", cases.golden_names("cgcn", rnn_type=None) + cases.golden_names("ctgcn", rnn_type=None))
def test_models(name):
    c = cases.load_case(name)
    m = c["meta"]
    if m["n"] > 500 and m["hid"] > 100:
        pytest.skip("large UCI case: covered by make_golden.py at generation time and by the GPU parity test")
    if m["kind"] == "ctgcn":
        res = oracle_np.ctgcn(c["x_list"], c["adj_lists"], c["sd"], m["trans_num"], m["diffusion_num"], m["model_type"], m["act"])
    else:
        xs, adj = (c["x_list"][0], c["adj_lists"][0]) if m["single"] else (c["x_list"], c["adj_lists"])
        res = oracle_np.cgcn(xs, adj, c["sd"], m["trans_num"], m["diffusion_num"], m["model_type"], m["act"])
    y, tr = _model_out(c, res)
    assert cases.relerr(y, c["expected"]["y"]) < TOL
    if tr is not None:
        assert cases.relerr(tr, c["expected"]["trans"]) < TOL


def test_build_core_adj_list_semantics():
    import scipy.sparse as sp
    a3 = sp.csr_matrix(np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0]], dtype=float))
    a2 = a3.copy()                       # identical to the previous level → dropped (helper.py:74-76)
    a1 = sp.csr_matrix(np.array([[0, 1, 1], [1, 0, 0], [1, 0, 0]], dtype=float))
    adj, mc = oracle_np.build_core_adj_list([a1, a2, a3])
    assert mc == 3 and len(adj) == 2
    assert (adj[0].toarray() == a3.toarray() + np.eye(3)).all()   # +I on the first (densest-index) entry only
    assert (adj[1].toarray() == a1.toarray()).all()
    adj2, mc2 = oracle_np.build_core_adj_list([a1, a2, a3], max_core=2)  # sticky max_core keeps files [:2]
    assert mc2 == 2 and len(adj2) == 2 and (adj2[0].toarray() == a2.toarray() + np.eye(3)).all()


def test_synth_np_matches_networkx_and_the_loader_contract():
    """oracle/synth_np.py (the reference arm's graph generator): exact core numbers (vs networkx) and the list of
    helper.py:51-82 (densest core first, +I on the first matrix, nested, identical levels dropped)."""
    import networkx as nx
    from oracle import synth_np
    for kind in ("er", "powerlaw"):
        n, m = 1500, 9000
        u, v = (synth_np.er_edges if kind == "er" else synth_np.powerlaw_edges)(n, m, np.random.default_rng(5))
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from(zip(u.tolist(), v.tolist()))
        cn = nx.core_number(g)
        assert (synth_np.core_numbers(n, u, v) == np.array([cn[i] for i in range(n)])).all()
        for levels, k in (("top", 3), ("loader", 5)):
            adj, st = synth_np.make_adj_list(kind, n, m, k, seed=5, levels=levels)
            dense = [a.to_dense().numpy() for a in adj]
            assert st["k"] == len(adj) <= k and st["edges_aggregated"] == sum(int(a._nnz()) for a in adj)
            first = dense[0] - np.eye(n, dtype=np.float32)
            assert (np.diag(first) == 0).all() and (first == first.T).all()
            prev = first
            for lvl, d in zip(st["core_levels"][1:], dense[1:]):
                assert (d - prev >= 0).all() and (d != prev).any()          # nested, and never a duplicate of the previous entry
                sub = nx.k_core(g, k=lvl, core_number=cn)
                a = np.zeros((n, n), dtype=np.float32)
                for x, y in sub.edges():
                    a[x, y] = a[y, x] = 1.0
                assert (a == d).all()
                prev = d
