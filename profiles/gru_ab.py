#!/usr/bin/env python
"""A/B timing of the tensor-core GRU builds on the sequence kernel alone (bench-size rows; CUDA events, L2-exceeding inputs).
    python profiles/gru_ab.py [--n 1000000] [--impls one_cta_r1,unpaired,auto]
Prints per build: ms per launch for the core GRU (K = 10 steps, SUM_LN) and the temporal GRU (T = 8 steps, EACH_LN), algorithmic
TFLOP/s, cycles per tile-step at the sampled SM clock, relL2 against the fp32 SIMT kernel on a 20 K-row slice."""
import argparse
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d-in", type=int, default=128)
    ap.add_argument("--h", type=int, default=128)
    ap.add_argument("--core-steps", type=int, default=10)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--impls", default="one_cta_r1,unpaired,auto")
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    from ctgcn_b200 import _lib, ops
    from oracle import cases
    dev = torch.device("cuda:0")
    n, d, h = args.n, args.d_in, args.h
    rng = np.random.default_rng(0)
    sd = cases.gru_params(rng, "rnn.", d, h)
    sd.update(cases.norm_params(rng, "norm.", h))
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd.items()}
    w = (sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"], sd["norm.weight"], sd["norm.bias"], 1e-5)
    codes = {"simt": _lib.IMPL_SIMT, "auto": _lib.IMPL_AUTO, "unpaired": _lib.IMPL_TC_UNPAIRED, "one_cta_r1": _lib.IMPL_TC_ONE_CTA_R1,
             "wide": _lib.IMPL_TC_WIDE}

    def select(name):
        _lib.set_gru_impl(codes[name])

    for steps, mode, label in ((args.core_steps, _lib.GRU_SUM_LN, f"core GRU  K={args.core_steps} SUM_LN "), (8, _lib.GRU_EACH_LN, "temporal  T=8  EACH_LN")):
        seq = torch.randn(n, steps, d, device=dev).abs_()
        small = seq[:20_000].contiguous()
        _lib.set_gru_impl(_lib.IMPL_SIMT)
        ref = ops.gru_seq(small, *w, mode)
        flops = n * steps * 6 * h * (d + h)
        first_out = None
        for name in args.impls.split(","):
            select(name)
            got = ops.gru_seq(small, *w, mode)
            err = ((got - ref).norm() / ref.norm()).item()
            out = torch.empty((n, h) if mode == _lib.GRU_SUM_LN else (n, steps, h), device=dev)
            for _ in range(3):
                ops.gru_seq(seq, *w, mode, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                ops.gru_seq(seq, *w, mode, out=out)
            e1.record()
            torch.cuda.synchronize()
            clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                                 capture_output=True, text=True).stdout.strip()
            ms = e0.elapsed_time(e1) / args.iters
            same = ""
            if name not in ("one_cta_r1", "wide", "simt"):             # the gru_tc2 builds do the same arithmetic in the same order: any difference is a race
                if first_out is None:
                    first_out = out.clone()
                else:
                    same = f"  bit-identical to first gru_tc2 build: {torch.equal(out, first_out)}"
            tile_steps = -(-n // 128) * steps / 148
            cyc = ms * 1e-3 * float(clk or 0) * 1e6 / tile_steps
            print(f"{label} {name:11s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s algorithmic ({3 * flops / ms / 1e9:7.1f} issued)  "
                  f"~{cyc:6.0f} cycles/tile-step @ {clk} MHz (idle-sampled)  relL2 vs fp32 kernel {err:.2e}{same}", flush=True)
        del seq
    _lib.set_gru_impl(_lib.IMPL_AUTO)


if __name__ == "__main__":
    main()
