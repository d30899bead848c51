"""EXPERIMENT (round-2 groundwork, NOT run yet on a GPU): hub-row splitting for power-law graphs (BASELINE.json configs[4]).

    python profiles/try_hubsplit.py [--n 1000000] [--m 10000000] [--k 20] [--d 256] [--threshold 4096]

Chung-Lu snapshot with the loader's core levels (cores K..1 = every edge): the largest row has ~10^5 entries and the
warp-per-row SpMM walks it serially.  Compares ops.cumspmm with ops.cumspmm_hubsplit (values within fp32 summation order,
time per launch)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__
__graft_entry__.build()
from ctgcn_b200 import ops, synth


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=10_000_000)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--threshold", type=int, default=4096)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    snap = synth.make_snapshot("powerlaw", a.n, a.m, a.k, seed=0, levels="loader")
    deg = np.diff(snap.rowptr)
    print(f"entries {snap.entries}, K {snap.k}, longest rows {np.sort(deg)[-5:].tolist()}, rows > {a.threshold}: {(deg > a.threshold).sum()}")
    plan = snap.plan(dev)
    x = synth.features(a.n, a.d, 1).to(dev)
    ref = ops.cumspmm(plan, x)
    got = ops.cumspmm_hubsplit(plan, x, a.threshold)
    err = float((ref - got).abs().max() / ref.abs().max())
    print(f"max |diff| / max |ref| = {err:.2e}")
    t0 = timed(lambda: ops.cumspmm(plan, x))
    t1 = timed(lambda: ops.cumspmm_hubsplit(plan, x, a.threshold))
    print(f"cumspmm {t0:.2f} ms, hub-split {t1:.2f} ms ({t0 / t1:.2f}x)")


if __name__ == "__main__":
    main()
