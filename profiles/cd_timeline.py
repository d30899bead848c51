#!/usr/bin/env python
"""Timeline of the fused CoreDiffusion kernel (block 0): gather-warp tile events next to the GRU pipeline's step events.
    python profiles/cd_timeline.py [--config cfg2]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4")
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    import bench
    from ctgcn_b200 import _lib, ops, synth
    from oracle import cases
    cfg = bench.CONFIGS[args.config]
    dev = torch.device("cuda:0")
    n, d = cfg["n"], cfg["D"]
    snap = synth.make_snapshot(cfg["kind"], n, cfg["m"], cfg["K"], seed=0)
    plan = snap.plan(dev)
    x = synth.features(n, d, 1000).to(dev)
    rng = np.random.default_rng(0)
    sd = cases.gru_params(rng, "rnn.", d, d)
    sd.update(cases.norm_params(rng, "norm.", d))
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd.items()}
    w = (sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"], sd["norm.weight"], sd["norm.bias"], 1e-5)
    y = torch.empty(n, d, device=dev)
    _lib.set_fusion(True)
    ops.core_diffusion(plan, x, *w, out=y)
    torch.cuda.synchronize()
    buf = torch.zeros(32, 64, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib.ctgcn_debug_gru_trace(C.c_void_p(buf.data_ptr())), "trace on")
    ops.core_diffusion(plan, x, *w, out=y)
    torch.cuda.synchronize()
    _lib.lib.ctgcn_debug_gru_trace(None)
    t = buf.cpu().numpy()
    t0 = t[24, 0]
    K = snap.k
    print(f"K = {K} steps per tile; cycles relative to the gather's first tile start (block 0)")
    names = {24: "gather: tile begin", 25: "gather: slot free", 26: "gather: tile done"}
    for e in (24, 25, 26):
        print(f"{names[e]:24s}", " ".join(f"{int(t[e, s] - t0):8d}" if t[e, s] else "       -" for s in range(8)))
    for e, nm in ((27, "gather w0: cp.async wait"), (28, "gather w0: flush (emit)"), (29, "gather w0: issue cp.async"), (30, "gather w0: metadata"), (31, "gather w0: L2 prefetch")):
        print(f"{nm:26s}", " ".join(f"{int(t[e, s]):8d}" for s in range(8)), " (cycles per tile)")
    for e, nm in ((0, "mma: step begin"), (3, "mma: h ready"), (7, "mma: last part issued"), (12, "wrk: h published"), (13, "ldr: U buffer free"), (14, "ldr: U staged"), (15, "mma: cycles waiting W")):
        print(f"{nm:24s}", " ".join(f"{int(t[e, s] - (0 if e == 15 else t0)):8d}" if t[e, s] else "       -" for s in range(0, 3 * K, 1))[:400])


if __name__ == "__main__":
    main()
