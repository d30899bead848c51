#!/usr/bin/env python
"""A/B of ONE CoreDiffusion call (layers.py:38-63) at bench size: the one-launch fused kernel against the two-kernel path
(cumulative SpMM → U in HBM → GRU).  python profiles/cd_ab.py [--config cfg4|cfg2] [--iters 10]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    import bench
    from ctgcn_b200 import _lib, ops, synth
    from oracle import cases
    cfg = bench.CONFIGS[args.config]
    dev = torch.device("cuda:0")
    n, d = cfg["n"], cfg["D"]
    snap = synth.make_snapshot(cfg["kind"], n, cfg["m"], cfg["K"], seed=0, levels=cfg.get("levels", "top"))
    plan = snap.plan(dev)
    x = synth.features(n, d, 1000).to(dev)
    rng = np.random.default_rng(0)
    sd = cases.gru_params(rng, "rnn.", d, d)
    sd.update(cases.norm_params(rng, "norm.", d))
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd.items()}
    w = (sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"], sd["norm.weight"], sd["norm.bias"], 1e-5)
    outs = {}
    for name, on in (("two-kernel", False), ("fused", True), ("two-kernel", False), ("fused", True)):
        _lib.set_fusion(on)
        y = torch.empty(n, d, device=dev)
        for _ in range(2):
            ops.core_diffusion(plan, x, *w, out=y)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            ops.core_diffusion(plan, x, *w, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        launches = (_lib.launch_count() - l0) / args.iters
        same = ""
        if name in outs:
            same = f"  repeatable: {torch.equal(outs[name], y)}"
        elif outs:
            first = next(iter(outs.values()))
            same = f"  bit-identical to two-kernel: {torch.equal(first, y)}  max|diff| {float((first - y).abs().max()):.2e}"
        outs.setdefault(name, y.clone())
        print(f"{args.config} CoreDiffusion {name:10s} {ms:7.3f} ms per call, {launches:.0f} launches per call, "
              f"{snap.edges_aggregated / ms / 1e6:8.1f} G edges-aggregated/s{same}", flush=True)
    _lib.set_fusion(True)


if __name__ == "__main__":
    main()
