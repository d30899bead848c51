"""EXPERIMENT (never run on a GPU yet): ONE experimental core-GRU build (ctgcn_set_coop_mode M) against the default kernel, in a
process of its own — a trap in one build must not take the measurements of the others down (try_coop.py runs them in a row).

    python profiles/try_mode.py --mode 2|3|5|6 [--config cfg2|cfg4] [--iters 10]

Prints the core-GRU launch time (default vs mode M), the relative L2 difference of the outputs (modes 2: bit-identical to the
default; 3, 5, 6: ≈ 1e-6 and bit-identical among themselves) and the whole CTGCN.forward time with the mode on."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__
__graft_entry__.build()
import bench
import ctgcn_b200 as pkg
from ctgcn_b200 import _lib, ops, synth
from try_coop import timed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", type=int, required=True, choices=[1, 2, 3, 5, 6])
    ap.add_argument("--config", default="cfg2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.config]
    dev = torch.device("cuda:0")
    T, n, d, K = cfg["T"], cfg["n"], cfg["D"], cfg["K"]
    plans = [synth.make_snapshot(cfg["kind"], n, cfg["m"], K, seed=t, levels=cfg.get("levels", "top")).plan(dev) for t in range(T)]
    xs = [synth.features(n, d, 1000 + t).to(dev) for t in range(T)]
    torch.manual_seed(0)
    model = pkg.CTGCN(d, d, d, 1, 1, T).to(dev).eval()
    lay = model.duffision_list[0].diffusion_list[0]
    u = ops.cumspmm(plans[0], xs[0])
    gru = lambda: ops.rnn_seq(u, *lay._gru_params(), lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    fwd = lambda: model(xs, plans)
    with torch.no_grad():
        ref, t_def = gru().clone(), timed(gru, args.iters)
        out_ref, t_fwd = fwd().clone(), timed(fwd, args.iters)
        _lib.set_coop_mode(args.mode)
        got = gru().clone()
        torch.cuda.synchronize()
        rel = ((got - ref).norm() / ref.norm()).item()
        print(f"mode {args.mode} {args.config}: core GRU launch default {t_def:.3f} ms → {timed(gru, args.iters):.3f} ms; "
              f"relL2 vs default {rel:.2e}, bit-identical {torch.equal(got, ref)}, checksum {got.double().sum().item():.10e}")
        out = fwd().clone()
        rel = ((out - out_ref).norm() / out_ref.norm()).item()
        print(f"mode {args.mode} {args.config}: CTGCN.forward default {t_fwd:.2f} ms → {timed(fwd, args.iters):.2f} ms; relL2 {rel:.2e}, "
              f"checksum {out.double().sum().item():.10e}")
    _lib.set_coop_mode(0)


if __name__ == "__main__":
    main()
