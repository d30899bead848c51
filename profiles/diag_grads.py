"""Diagnostic (GPU): per-tensor gradient error of every *grad* golden case under both implementations (SIMT fp32 /
tcgen05 split-bf16), printed worst-first.  python profiles/diag_grads.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import __graft_entry__
__graft_entry__.build()
import ctgcn_b200 as pkg
from ctgcn_b200 import _lib
import grad_checks
from oracle import cases

dev = torch.device("cuda:0")
names = sys.argv[1:] or (cases.golden_names("ctgcn", rnn_type=None, grads=True) + cases.golden_names("cgcn", rnn_type=None, grads=True))
for name in names:
    c = cases.load_case(name)
    m = c["meta"]
    for impl, code in (("simt", _lib.IMPL_SIMT), ("auto", _lib.IMPL_AUTO)):
        _lib.set_gru_impl(code)
        runs = []
        for rep in range(2):
            mod = grad_checks.build_model(pkg, m, dev)
            mod.load_state_dict(grad_checks.tsd(c["sd"], dev), strict=True)
            xs, adj = grad_checks.model_inputs(c, dev)
            res = mod(xs[0], adj[0]) if m.get("single") else mod(xs, adj)
            out, trans = res if m["model_type"] == "S" else (res, None)
            grad_checks.loss_of([grad_checks.stack3(out)] + ([grad_checks.stack3(trans)] if trans is not None else []),
                                m["cot_seed"]).backward()
            runs.append({k: p.grad.detach().cpu().numpy() for k, p in mod.named_parameters() if p.grad is not None})
        errs = sorted(((cases.relerr(g, c["grads"][k]), k) for k, g in runs[0].items()), reverse=True)
        rep_diff = max(float(np.abs(runs[0][k] - runs[1][k]).max()) for k in runs[0])
        fwd = cases.relerr(grad_checks.stack3(out).detach().cpu().numpy()[:, ::m["row_stride"]], c["expected"]["y"])
        print(f"{name} [{impl}] fwd {fwd:.1e} repeat-diff {rep_diff:.1e} worst:", ", ".join(f"{k} {e:.1e}" for e, k in errs[:6]), flush=True)
_lib.set_gru_impl(_lib.IMPL_AUTO)
