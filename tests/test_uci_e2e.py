"""BASELINE.json configs[0]: config/uci.json CTGCN-C, T = 1 — the reference's own CPU-runnable case, end to end.

tests/golden_e2e/uci_e2e_ctgcn_C.npz was produced by running the UNMODIFIED reference pipeline in the build container
(oracle/make_uci_e2e_golden.py): preprocessing → train.gnn_embedding (2 epochs, shipped hyper-parameters) → exported TSV, with
the weights of the last forward captured.  Here the same snapshot goes through ctgcn_b200/io.py (k-core files, loader
contract) and the drop-in model; the result must equal the reference's exported embeddings."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import cases, oracle_np

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_e2e", "uci_e2e_ctgcn_C.npz")


@pytest.fixture(scope="module")
def uci(tmp_path_factory, lib):
    from ctgcn_b200 import io
    z = dict(np.load(FIX))
    meta = json.loads(bytes(z["meta"]).decode())
    base = tmp_path_factory.mktemp("uci")
    (base / "1.format").mkdir()
    (base / "1.format" / meta["snapshot"]).write_text(bytes(z["csv"]).decode())
    (base / "nodes.csv").write_text(bytes(z["nodes"]).decode() + "\n")
    kmax = io.preprocess_kcores(str(base / "1.format"), str(base / "cores"), str(base / "nodes.csv"))
    nodes = io.read_node_list(str(base / "nodes.csv"))
    sd = {k[4:]: v for k, v in z.items() if k.startswith("sd::")}
    return dict(meta=meta, base=base, kmax=kmax, nodes=nodes, sd=sd, emb=z["emb"])


def test_io_reproduces_the_reference_loader_and_oracle_the_export(uci):
    """CPU: k-core files + loader contract from ctgcn_b200/io.py give the list the reference's loader built (same K, same
    nnz per entry), and the fp64 oracle on that list reproduces the reference's exported TSV."""
    from ctgcn_b200 import io
    m = uci["meta"]
    assert uci["kmax"] == {m["snapshot"].split(".")[0]: 8}
    adj = io.load_core_adj_list(str(uci["base"] / "cores"), 0, 1)
    assert len(adj) == 1 and len(adj[0]) == m["k"]
    assert [int(a.nnz) for a in adj[0]] == m["nnz"]
    index = {nm: i for i, nm in enumerate(uci["nodes"])}
    u, v, w = io.read_edge_csv(str(uci["base"] / "1.format" / m["snapshot"]), index)
    snap, _ = io.snapshot_from_graph(m["n"], u, v, w)
    assert snap.nnz_per_core == m["nnz"]
    n = m["n"]
    y = oracle_np.ctgcn([sp.eye(n, format="coo", dtype=np.float32)], adj, uci["sd"], m["trans_num"], m["diffusion_num"], "C", "L")
    assert cases.relerr(y[0], uci["emb"]) < 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("graph_input", ["coo_list", "plans", "edge_list"])
def test_dropin_model_reproduces_the_exported_embeddings(uci, graph_input, lib, cuda_device, tmp_path):
    """GPU: the reference's checkpoint loads with strict=True; the forward on the one-hot sparse input (helper.py:169-172) over
    the list built by io equals the TSV the reference exported (1e-4); save_embedding writes what its evaluation reads."""
    pd = pytest.importorskip("pandas")
    import ctgcn_b200 as pkg
    from ctgcn_b200 import io
    from oracle import oracle_torch
    m = uci["meta"]
    n = m["n"]
    model = pkg.CTGCN(n, m["hid"], m["d_out"], m["trans_num"], m["diffusion_num"], 1, model_type="C", trans_activate_type="L").to(cuda_device)
    assert sorted(model.state_dict().keys()) == m["state_dict_keys"]
    model.load_state_dict({k: torch.from_numpy(v).to(cuda_device) for k, v in uci["sd"].items()}, strict=True)
    if graph_input == "coo_list":
        adj = [[oracle_torch.to_torch_coo(a).to(cuda_device) for a in al] for al in io.load_core_adj_list(str(uci["base"] / "cores"), 0, 1)]
    elif graph_input == "plans":
        adj = io.load_core_plans(str(uci["base"] / "cores"), 0, 1, cuda_device)
    else:
        index = {nm: i for i, nm in enumerate(uci["nodes"])}
        u, v, w = io.read_edge_csv(str(uci["base"] / "1.format" / m["snapshot"]), index)
        adj = [io.snapshot_from_graph(n, u, v, w)[0].plan(cuda_device)]
    x = [oracle_torch.to_torch_coo(sp.eye(n, format="coo", dtype=np.float32)).to(cuda_device)]
    model.eval()
    with torch.no_grad():
        out = model(x, adj)
    assert tuple(out.shape) == (1, n, m["d_out"])
    got = out[0].cpu().numpy()
    err = cases.relerr(got, uci["emb"])
    assert err <= 1e-4, err
    np.testing.assert_allclose(got, uci["emb"], rtol=1e-4, atol=1e-4 * np.abs(uci["emb"]).max())
    (path,) = io.save_embedding(out, str(tmp_path / "emb"), [m["snapshot"]], uci["nodes"])
    df = pd.read_csv(path, sep="\t", index_col=0)
    assert list(df.index) == uci["nodes"] and np.array_equal(df.values.astype(np.float32), got)
