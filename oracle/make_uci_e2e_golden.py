#!/usr/bin/env python
"""Golden fixture for BASELINE.json configs[0] (config/uci.json, CTGCN-C, T = 1, CPU) — TEST INFRASTRUCTURE.

Runs the UNMODIFIED reference end to end in the build container (needs /root/reference): its preprocessing
(k-core files, random walks), `train.gnn_embedding('CTGCN-C')` for 2 epochs on the first UCI snapshot with the
shipped hyper-parameters (hid 500, embed 128, 1 linear + 2 diffusion layers, U-neg loss, Adam), and its embedding export.
Stored in tests/golden/uci_e2e_ctgcn_C.npz:
  * the snapshot's edge file and the node list (inputs of ctgcn_b200/io.py),
  * the model's state_dict AT THE LAST FORWARD (the reference exports the output of the last training forward,
    embedding.py:346,361, i.e. before the final optimizer step) — captured with a forward pre-hook, reference files untouched,
  * the exported embedding TSV (parsed), and the list lengths / nnz of the adj_list its loader built.
The reference's plumbing needs a compatibility shim on this container's library versions; the hot path
(layers.py / models.py) runs as shipped.
"""
from __future__ import annotations

import json
import os
import random
import shutil
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

import pandas as pd  # noqa: E402
import torch  # noqa: E402

from oracle import ref_compat  # noqa: E402

ref_compat.apply()
from oracle import cases, oracle_np  # noqa: E402

SNAPSHOT = "2004-04.csv"
EPOCHS = 2


def main():
    torch.set_num_threads(2)
    cfg = json.load(open(os.path.join(REF, "config/uci.json")))
    tmp = tempfile.mkdtemp(prefix="uci_e2e_")
    try:
        base = os.path.join(tmp, "uci")
        os.makedirs(os.path.join(base, "1.format"))
        os.makedirs(os.path.join(base, "nodes_set"))
        shutil.copy(os.path.join(REF, "data/uci/1.format", SNAPSHOT), os.path.join(base, "1.format", SNAPSHOT))
        shutil.copy(os.path.join(REF, "data/uci/nodes_set/nodes.csv"), os.path.join(base, "nodes_set/nodes.csv"))

        random.seed(0)
        np.random.seed(0)
        torch.manual_seed(0)
        from preprocessing import preprocess
        pre = dict(cfg["preprocessing"]["CTGCN-C"], base_path=base, worker=-1)
        preprocess("CTGCN-C", pre)

        import train
        captured = {}
        orig_get_model = train.get_gnn_model

        def get_model_and_hook(method, time_length, args):
            model = orig_get_model(method, time_length, args)

            def pre_hook(mod, inputs):
                captured["sd"] = {k: v.detach().clone().numpy() for k, v in mod.state_dict().items()}
                captured["adj"] = inputs[1]
            model.register_forward_pre_hook(pre_hook)
            return model
        train.get_gnn_model = get_model_and_hook

        args = dict(cfg["embedding"]["CTGCN-C"], base_path=base, duration=1, start_idx=0, end_idx=-1, epoch=EPOCHS, use_cuda=False,
                    has_cuda=False, thread_num=2)
        train.gnn_embedding("CTGCN-C", args)

        emb_path = os.path.join(base, args["embed_folder"], SNAPSHOT)
        df = pd.read_csv(emb_path, sep="\t", index_col=0)
        nodes = [ln.strip() for ln in open(os.path.join(base, "nodes_set/nodes.csv")).read().split("\n") if ln.strip()]
        assert list(df.index) == nodes
        emb = df.values.astype(np.float32)
        sd = captured["sd"]
        adj_ref = captured["adj"][0]                          # the loader's list for the one snapshot
        n = len(nodes)
        print("exported", emb.shape, "K =", len(adj_ref), "nnz =", [int(a._nnz()) for a in adj_ref])

        # the restatement reproduces the exported embeddings from the captured weights and the loader's list
        mats = [sp.coo_matrix((a._values().numpy(), a._indices().numpy()), shape=(n, n)) for a in adj_ref]
        y = oracle_np.ctgcn([sp.eye(n, format="coo", dtype=np.float32)], [mats], sd, 1, 2, "C", "L")
        err = cases.relerr(y[0], emb)
        print(f"oracle_np(fp64) on captured weights vs exported TSV: relL2 {err:.2e}")
        assert err < 5e-6, err

        csv_text = open(os.path.join(base, "1.format", SNAPSHOT)).read()
        meta = dict(kind="uci_e2e", name="uci_e2e_ctgcn_C", snapshot=SNAPSHOT, n=n, epochs=EPOCHS, hid=args["hid_dim"], d_out=args["embed_dim"],
                    trans_num=args["trans_layer_num"], diffusion_num=args["diffusion_layer_num"], model_type="C", act="L", rnn_type="GRU",
                    k=len(adj_ref), nnz=[int(a._nnz()) for a in adj_ref], state_dict_keys=sorted(sd.keys()))
        out = os.path.join(ROOT, "tests", "golden_e2e")
        os.makedirs(out, exist_ok=True)
        arrays = {"sd::" + k: v for k, v in sd.items()}
        np.savez_compressed(os.path.join(out, "uci_e2e_ctgcn_C.npz"), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                            csv=np.frombuffer(csv_text.encode(), dtype=np.uint8), nodes=np.frombuffer("\n".join(nodes).encode(), dtype=np.uint8),
                            emb=emb, **arrays)
        print("wrote", os.path.join(out, "uci_e2e_ctgcn_C.npz"), os.path.getsize(os.path.join(out, "uci_e2e_ctgcn_C.npz")) // 1024, "KiB")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
