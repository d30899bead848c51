#!/usr/bin/env python
"""Pipeline timeline of the tcgen05 GRU kernel (block 0): clock64 stamps of MMA-issuer / worker / loader events.
    python profiles/gru_timeline.py [--steps 10] [--mode 0|1]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

NAMES = {0: "mma: step begin", 1: "mma: U ready & acc0 free", 2: "mma: X half0 issued", 3: "mma: h ready", 4: "mma: H half0 issued",
         5: "mma: acc1 free", 6: "mma: X half1 issued", 7: "mma: H half1 issued", 8: "wrk: wait acc0", 9: "wrk: acc0 full",
         10: "wrk: half0 math done / wait acc1", 11: "wrk: acc1 full", 12: "wrk: h published", 13: "ldr: U buffer free",
         14: "ldr: U staged", 15: "mma: cycles waiting for weights", 16: "wrk: half0 pass0 done", 17: "wrk: half0 pass1 done", 18: "wrk: half0 pass2 done", 19: "wrk: half0 pass3 done",
         20: "wrk: half1 pass0 done", 21: "wrk: half1 pass1 done", 22: "wrk: half1 pass2 done", 23: "wrk: half1 pass3 done"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--impl", default="auto", choices=["auto", "unpaired", "one_cta_r1"])
    args = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    from ctgcn_b200 import _lib, ops
    from oracle import cases
    _lib.set_gru_impl({"unpaired": _lib.IMPL_TC_UNPAIRED, "one_cta_r1": _lib.IMPL_TC_ONE_CTA_R1}.get(args.impl, _lib.IMPL_AUTO))
    dev = torch.device("cuda:0")
    n, d = 148 * 128 * 4, 128
    rng = np.random.default_rng(0)
    sd = cases.gru_params(rng, "rnn.", d, d)
    sd.update(cases.norm_params(rng, "norm.", d))
    sd = {k: torch.from_numpy(v).to(dev) for k, v in sd.items()}
    seq = torch.randn(n, args.steps, d, device=dev).abs()
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
    t0 = t[0, 0]
    nsteps = min(64, 4 * args.steps)
    print("cycles relative to the first step's begin; one column per global step of block 0 (4 tiles x steps)")
    for e in sorted(NAMES):
        row = [(int(t[e, s] - (0 if e == 15 else t0)) if t[e, s] else None) for s in range(nsteps)]
        print(f"{NAMES[e]:34s}", " ".join(f"{v:7d}" if v is not None else "      -" for v in row[: 2 * args.steps + 2]))
    per = np.diff(t[0, : nsteps].astype(np.int64))
    print("step period (cycles):", per[: 2 * args.steps + 2].tolist())


if __name__ == "__main__":
    main()
