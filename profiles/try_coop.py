"""EXPERIMENT (round-2 groundwork, NOT run yet on a GPU): SpMM / GRU co-residency.

    python profiles/try_coop.py [--config cfg2|cfg4] [--iters 10]

1. numerics: the reduced-register kernel variants (ctgcn_set_coop_mode(1): gru_tc_coop_kernel with Σh in an L2 scratch, the
   64-register SpMM) must give bit-identical results to the default kernels — first on one core-GRU launch, then on the whole
   CTGCN.forward, (a) with the chunk-pipelined CoreDiffusion inside the C-ABI call (SpMM of row chunk c+1 under the GRU of
   chunk c) and (b) with model.coop = True (SpMM(t+1) on the current stream under GRU(t) on a high-priority stream);
2. timing: ms per forward, default vs co-resident, and the per-kernel-class times (ctgcn_prof_*).
If the GRU launch regresses alone or the overlap does not materialise (block scheduler keeps the kernels apart), check with
`nsys`-less timeline: CUDA events around each launch on both streams."""
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


def timed(fn, iters):
    for _ in range(3):
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

    # ---- 1a. one core-GRU launch: default vs co-resident variant
    lay = model.duffision_list[0].diffusion_list[0]
    u = ops.cumspmm(plans[0], xs[0])
    w = lay._gru_params()
    ref = ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    t_def = timed(lambda: ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN), args.iters)
    t_spmm_def = timed(lambda: ops.cumspmm(plans[0], xs[0], out=u), args.iters)
    _lib.set_coop_mode(True)
    got = ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    t_coop = timed(lambda: ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN), args.iters)
    u2 = ops.cumspmm(plans[0], xs[0])
    t_spmm_coop = timed(lambda: ops.cumspmm(plans[0], xs[0], out=u2), args.iters)
    print(f"core GRU launch: default {t_def:.3f} ms, coop variant {t_coop:.3f} ms, bit-identical: {torch.equal(ref, got)}")
    _lib.set_coop_mode(2)
    got16 = ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    t_w16 = timed(lambda: ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN), args.iters)
    print(f"core GRU launch: 16 gate warps {t_w16:.3f} ms, bit-identical: {torch.equal(ref, got16)}")
    _lib.set_coop_mode(3)     # + input-side biases added by the tensor core (bf16 hi+lo of the bias: not bit-identical)
    got16f = ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    t_w16f = timed(lambda: ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN), args.iters)
    rel = ((got16f - ref).norm() / ref.norm()).item()
    print(f"core GRU launch: 16 gate warps + folded biases {t_w16f:.3f} ms, relL2 vs default {rel:.2e} (expect ~1e-6; bar 1e-4)")
    _lib.set_coop_mode(5)     # mode 3 with FADD2 / FMUL2 / FFMA2 gate math: the same arithmetic, so bit-identical to mode 3
    got16p = ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    t_w16p = timed(lambda: ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN), args.iters)
    print(f"core GRU launch: … + packed fp32x2 gate math {t_w16p:.3f} ms, bit-identical to mode 3: {torch.equal(got16f, got16p)}")
    _lib.set_coop_mode(6)     # + both input parts issued before h is awaited, accumulator sets released after their last TMEM load
    got16r = ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN)
    t_w16r = timed(lambda: ops.rnn_seq(u, *w, lay.norm.weight, lay.norm.bias, lay.norm.eps, _lib.GRU_SUM_LN), args.iters)
    print(f"core GRU launch: … + reordered MMA schedule {t_w16r:.3f} ms, bit-identical to mode 3: {torch.equal(got16f, got16r)}")
    _lib.set_coop_mode(True)
    print(f"SpMM launch:     default {t_spmm_def:.3f} ms, 64-reg variant {t_spmm_coop:.3f} ms, bit-identical: {torch.equal(u, u2)}")
    _lib.set_coop_mode(False)
    # ---- 1b / 2. whole forward
    with torch.no_grad():
        out_ref = model(xs, plans).clone()
        t_fwd = timed(lambda: model(xs, plans), args.iters)
        _lib.set_coop_mode(True)
        out_pipe = model(xs, plans).clone()          # C-ABI level: SpMM(chunk c+1) under GRU(chunk c) inside every CoreDiffusion call
        t_fwd_pipe = timed(lambda: model(xs, plans), args.iters)
        print(f"CTGCN.forward {args.config}: default {t_fwd:.2f} ms, chunk-pipelined CoreDiffusion {t_fwd_pipe:.2f} ms "
              f"({t_fwd / t_fwd_pipe:.2f}x), bit-identical: {torch.equal(out_ref, out_pipe)}")
        model.coop = True
        out_coop = model(xs, plans).clone()
        t_fwd_coop = timed(lambda: model(xs, plans), args.iters)
        _lib.prof_collect(reset=True)
        _lib.prof_enable(True)
        model(xs, plans)
        print("per-class ms of one co-resident forward:", {k: round(v["ms"], 3) for k, v in _lib.prof_collect().items()})
        _lib.prof_enable(False)
    print(f"CTGCN.forward {args.config}: default {t_fwd:.2f} ms, co-resident {t_fwd_coop:.2f} ms "
          f"({t_fwd / t_fwd_coop:.2f}x), bit-identical: {torch.equal(out_ref, out_coop)}")
    _lib.set_coop_mode(2)
    model.coop = False
    with torch.no_grad():
        out_w16 = model(xs, plans).clone()
        t_fwd_w16 = timed(lambda: model(xs, plans), args.iters)
    print(f"CTGCN.forward {args.config}: 16 gate warps in the core GRU {t_fwd_w16:.2f} ms ({t_fwd / t_fwd_w16:.2f}x), "
          f"bit-identical: {torch.equal(out_ref, out_w16)}")
    _lib.set_coop_mode(3)
    with torch.no_grad():
        out_w16f = model(xs, plans).clone()
        t_fwd_w16f = timed(lambda: model(xs, plans), args.iters)
    rel = ((out_w16f - out_ref).norm() / out_ref.norm()).item()
    print(f"CTGCN.forward {args.config}: 16 gate warps + folded biases {t_fwd_w16f:.2f} ms ({t_fwd / t_fwd_w16f:.2f}x), "
          f"relL2 vs default {rel:.2e}")
    _lib.set_coop_mode(6)
    with torch.no_grad():
        out_w16r = model(xs, plans).clone()
        t_fwd_w16r = timed(lambda: model(xs, plans), args.iters)
    print(f"CTGCN.forward {args.config}: mode 6 (16 gate warps, folded biases, packed math, reordered MMAs) {t_fwd_w16r:.2f} ms "
          f"({t_fwd / t_fwd_w16r:.2f}x), bit-identical to mode 3: {torch.equal(out_w16f, out_w16r)}")
    _lib.set_coop_mode(False)

    # ---- 3. LAST, because a trap here would poison the context: the pre-split-U path (profiles/r02_gru_design.md step 3)
    try:
        # pre-split U (design note step 3): does the 16-byte plane store pattern cost the SpMM anything?
        t_packed = timed(lambda: ops.cumspmm_packed(plans[0], xs[0]), args.iters)      # includes a memset of the 5 GB buffer at cfg4:
        t_zero = timed(lambda: torch.zeros(u.numel() * 4, dtype=torch.uint8, device=dev), args.iters)   # … measured and subtracted
        print(f"SpMM launch:     pre-split output {t_packed - t_zero:.3f} ms (default {t_spmm_def:.3f} ms)")
        # bulk-copy-fed GRU (16 gate warps, Σh in registers) behind the pre-split SpMM: one CoreDiffusion call, default vs packed
        cd_args = (*w, lay.norm.weight, lay.norm.bias, lay.norm.eps)
        y_def = ops.core_diffusion(plans[0], xs[0], *cd_args)
        t_cd_def = timed(lambda: ops.core_diffusion(plans[0], xs[0], *cd_args), args.iters)
        y_pk = ops.core_diffusion_packed(plans[0], xs[0], *cd_args)
        t_cd_pk = timed(lambda: ops.core_diffusion_packed(plans[0], xs[0], *cd_args), args.iters)
        rel = ((y_pk - y_def).norm() / y_def.norm()).item()
        print(f"CoreDiffusion call: default {t_cd_def:.3f} ms, pre-split U + bulk-copy-fed GRU {t_cd_pk:.3f} ms, relL2 {rel:.1e}")
    except Exception as exc:  # noqa: BLE001 - report and keep the measurements above
        print("pre-split-U path failed:", repr(exc))


if __name__ == "__main__":
    main()
