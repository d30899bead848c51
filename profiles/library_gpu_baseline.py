"""The "library GPU" baseline SURVEY.md §8(d) mentions as optional: the reference's own arithmetic (torch.sparse.mm × K, nn.GRU through
cuDNN, LayerNorm, nn.Linear — oracle/oracle_torch.py issues the same library calls as layers.py / models.py) with every tensor on one
B200, next to this repo's forward on the same inputs.  Not a bench.py arm (the driver's reference arm is the CPU path); a context
number for the write-up: how much of the speed-up over the CPU comes from the GPU at all and how much from these kernels.

    python profiles/library_gpu_baseline.py [--config cfg2|cfg4|tiny] [--iters 5] [--device cuda|cpu]

`--device cpu --config tiny` runs the library leg alone (what the build container can check)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__
__graft_entry__.build()
import bench
from ctgcn_b200 import synth
from oracle import cases, oracle_torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.config]
    dev = torch.device(args.device)
    T, n, d, K = cfg["T"], cfg["n"], cfg["D"], cfg["K"]
    snaps = [synth.make_snapshot(cfg["kind"], n, cfg["m"], K, seed=t, levels=cfg.get("levels", "top")) for t in range(T)]
    e_agg = sum(s.edges_aggregated for s in snaps)
    xs = [synth.features(n, d, 1000 + t).to(dev) for t in range(T)]
    adj = [s.coo_list(dev) for s in snaps]                  # K uncoalesced COO matrices per snapshot, as helper.py hands them over
    sd_np = cases.ctgcn_params(np.random.default_rng(0), d, d, d, 1, 1, T, "C")
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd_np.items()}

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    def library():
        with torch.no_grad():
            return oracle_torch.ctgcn(xs, adj, sd, 1, 1, "C", "L")

    def timed(fn):
        fn()
        sync()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            out = fn()
        sync()
        return (time.perf_counter() - t0) / args.iters, out

    t_lib, out_lib = timed(library)
    print(f"{cfg['name']}: E_agg = {e_agg}")
    print(f"library path on {dev}: {t_lib * 1e3:.2f} ms per forward = {e_agg / t_lib:.3e} edges-aggregated/s")
    if dev.type != "cuda":
        return
    import ctgcn_b200 as pkg
    model = pkg.CTGCN(d, d, d, 1, 1, T).to(dev).eval()
    model.load_state_dict(sd)
    plans = [s.plan(dev) for s in snaps]

    def ours():
        with torch.no_grad():
            return model(xs, plans)

    t_own, out_own = timed(ours)
    rel = ((out_own - out_lib).norm() / out_lib.norm()).item()
    print(f"ctgcn_b200:           {t_own * 1e3:.2f} ms per forward = {e_agg / t_own:.3e} edges-aggregated/s  "
          f"({t_lib / t_own:.1f}x the library path); relL2 between the two outputs {rel:.1e}")


if __name__ == "__main__":
    main()
