#!/usr/bin/env python
"""Minimal driver for ncu captures of the hot kernels at benchmark size (no bench machinery around them).

    ncu --set full --clock-control none --import-source on -k regex:"gru_tc_kernel|cumspmm_vec|linear_kernel" \
        -o gpurun_out/prof python profiles/profile_kernels.py --config cfg4
Every kernel is launched twice (first launch = warm-up); use `-s`/`-c` or the regex to pick launches.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg2"])
    ap.add_argument("--gru-impl", default="auto")
    ap.add_argument("--fused", action="store_true", help="profile the one-launch CoreDiffusion build instead")
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    import bench
    from ctgcn_b200 import _lib, ops, synth
    from oracle import cases

    cfg = bench.CONFIGS[args.config]
    _lib.set_gru_impl({"auto": 0, "simt": 1, "tcgen05": 2}[args.gru_impl])
    dev = torch.device("cuda:0")
    n, d, T = cfg["n"], cfg["D"], cfg["T"]
    snap = synth.make_snapshot(cfg["kind"], n, cfg["m"], cfg["K"], seed=0)
    plan = snap.plan(dev)
    sd = {k: torch.from_numpy(v).to(dev) for k, v in cases.ctgcn_params(np.random.default_rng(0), d, d, d, 1, 1, 1, "C").items()}
    x = synth.features(n, d, 1000).to(dev)
    cd = "duffision_list.0.diffusion_list.0."
    if args.fused:
        _lib.set_fusion(True)
        for it in range(2):
            y = ops.core_diffusion(plan, x, sd[cd + "rnn.weight_ih_l0"], sd[cd + "rnn.weight_hh_l0"], sd[cd + "rnn.bias_ih_l0"],
                                   sd[cd + "rnn.bias_hh_l0"], sd[cd + "norm.weight"], sd[cd + "norm.bias"], 1e-5)
            torch.cuda.synchronize()
        print("ok fused", tuple(y.shape))
        return
    for it in range(2):
        h = ops.linear(x, sd["mlp_list.0.linear.weight"], sd["mlp_list.0.linear.bias"], _lib.ACT_NONE)
        u = ops.cumspmm(plan, h)
        y = ops.gru_seq(u, sd[cd + "rnn.weight_ih_l0"], sd[cd + "rnn.weight_hh_l0"], sd[cd + "rnn.bias_ih_l0"],
                        sd[cd + "rnn.bias_hh_l0"], sd[cd + "norm.weight"], sd[cd + "norm.bias"], 1e-5, _lib.GRU_SUM_LN)
        del u
        seq = torch.randn(n, T, d, device=dev)
        out = ops.gru_seq(seq, sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"],
                          sd["norm.weight"], sd["norm.bias"], 1e-5, _lib.GRU_EACH_LN)
        torch.cuda.synchronize()
    print("ok", tuple(y.shape), tuple(out.shape), f"entries={plan.entries} E_agg={snap.edges_aggregated}")


if __name__ == "__main__":
    main()
