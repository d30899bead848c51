#!/usr/bin/env python
"""Pipeline timeline of the step-per-launch GRU kernel for wide hidden states (gru_wide_tc.cu), block 0 of the LAST launch of a call.
    python profiles/gru_wide_timeline.py [--h 256] [--d-in 256] [--steps 4] [--mode 0|1]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

NAMES = {0: "mma: accumulators free (unit begin)", 1: "mma: first slice ready", 2: "mma: last MMA issued", 3: "epi: h prefetched, waiting",
         4: "epi: accumulators full", 5: "epi: done", 6: "ldr: slice 0 in registers", 7: "ldr: last slice stored",
         8: "mma: cycles waiting for A slices", 9: "mma: cycles waiting for weights"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=256)
    ap.add_argument("--d-in", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--n", type=int, default=400_000)
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    from ctgcn_b200 import _lib, ops
    from oracle import cases
    _lib.set_gru_impl(_lib.IMPL_TC_WIDE)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    sd = cases.gru_params(rng, "rnn.", args.d_in, args.h)
    sd.update(cases.norm_params(rng, "norm.", args.h))
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd.items()}
    # the last chunk of the call is a full one: its last launch is what the trace shows
    units = 4                                                   # UNITS_PER_CTA of gru_wide_tc.cu
    chunk = 148 * units * 128 // (args.h // 128)
    n = max(1, args.n // chunk) * chunk
    seq = torch.randn(n, args.steps, args.d_in, device=dev).abs()
    buf = torch.zeros(32, 64, dtype=torch.int64, device=dev)
    run = lambda: ops.gru_seq(seq, sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"],
                              sd["norm.weight"], sd["norm.bias"], 1e-5, args.mode)
    run()
    torch.cuda.synchronize()
    _lib.check(_lib.lib.ctgcn_debug_gru_trace(C.c_void_p(buf.data_ptr())), "trace on")
    run()
    torch.cuda.synchronize()
    _lib.lib.ctgcn_debug_gru_trace(None)
    t = buf.cpu().numpy()
    t0 = t[10, 0]
    print(f"n={n} rows, chunk={chunk}, {units} units per CTA and launch; cycles relative to the kernel's first instruction; one column per unit of block 0")
    for e in sorted(NAMES):
        row = [(int(t[e, s] - (0 if e in (8, 9) else t0)) if t[e, s] else None) for s in range(units)]
        print(f"{NAMES[e]:38s}", " ".join(f"{v:7d}" if v is not None else "      -" for v in row))
    print("kernel end:", int(t[11, 0] - t0))
    _lib.set_gru_impl(_lib.IMPL_AUTO)


if __name__ == "__main__":
    main()
