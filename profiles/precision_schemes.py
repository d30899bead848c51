"""CPU emulation of tensor-core operand precision schemes for the core GRU (why three MMAs per product).

    python profiles/precision_schemes.py [--n 4096] [--k 10] [--d 128]

One CoreDiffusion layer on a synthetic ER snapshot: U = relu(cumulative k-core sums) from the numpy oracle, PyTorch-default GRU
and LayerNorm weights, reference = fp64.  Every scheme replaces the two products of a GRU step (x·W_ihᵀ, h·W_hhᵀ) by what the
tensor core would compute from rounded operands (fp32 accumulation emulated in fp64: accumulation error is not the question
here), everything else stays fp64.  Printed: relative L2 error of LN(Σ_s h_s) against the fp64 result and the MMA cost in
bf16-MMA equivalents per product.  Needs no GPU and does not touch the product library.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np  # noqa: E402  (test infrastructure: this script is a numerics study, not the product)


def rnd(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dtype).to(torch.float64).numpy()


def tf32(a):
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    b = (b + 0x1000) & 0xFFFFE000          # round to 10 explicit mantissa bits
    return b.view(np.float32).astype(np.float64)


def split(a, dtype):
    hi = rnd(a, dtype)
    lo = rnd(a - hi, dtype)
    return hi, lo


def scheme_products(name):
    """Returns f(x, w) ≈ x @ w.T for one scheme."""
    bf, fp = torch.bfloat16, torch.float16
    if name == "fp64":
        return lambda x, w: x @ w.T
    if name == "bf16 x1":
        return lambda x, w: rnd(x, bf) @ rnd(w, bf).T
    if name == "fp16 x1":
        return lambda x, w: rnd(x, fp) @ rnd(w, fp).T
    if name == "tf32 x1":
        return lambda x, w: tf32(x) @ tf32(w).T
    if name == "bf16 x3 (shipped)":
        def f(x, w):
            xh, xl = split(x, bf)
            wh, wl = split(w, bf)
            return xh @ wh.T + xl @ wh.T + xh @ wl.T
        return f
    if name == "bf16 x2 (W unsplit)":
        def f(x, w):
            xh, xl = split(x, bf)
            return (xh + xl) @ rnd(w, bf).T
        return f
    if name == "bf16 x2 (x unsplit)":
        def f(x, w):
            wh, wl = split(w, bf)
            return rnd(x, bf) @ (wh + wl).T
        return f
    if name == "fp16 x2 (W unsplit)":
        def f(x, w):
            xh, xl = split(x, fp)
            return (xh + xl) @ rnd(w, fp).T
        return f
    if name == "fp16 x3":
        def f(x, w):
            xh, xl = split(x, fp)
            wh, wl = split(w, fp)
            return xh @ wh.T + xl @ wh.T + xh @ wl.T
        return f
    raise ValueError(name)


COST = {"fp64": "-", "bf16 x1": 1, "fp16 x1": 1, "tf32 x1": 2, "bf16 x3 (shipped)": 3, "bf16 x2 (W unsplit)": 2,
        "bf16 x2 (x unsplit)": 2, "fp16 x2 (W unsplit)": 2, "fp16 x3": 3}


def gru_sum_ln(u, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, mm):
    n, k, _ = u.shape
    hdim = w_hh.shape[1]
    h = np.zeros((n, hdim))
    acc = np.zeros((n, hdim))
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    for s in range(k):
        gi = mm(u[:, s, :], w_ih) + b_ih
        gh = mm(h, w_hh) + b_hh
        r = sig(gi[:, :hdim] + gh[:, :hdim])
        z = sig(gi[:, hdim:2 * hdim] + gh[:, hdim:2 * hdim])
        nn_ = np.tanh(gi[:, 2 * hdim:] + r * gh[:, 2 * hdim:])
        h = (1.0 - z) * nn_ + z * h
        acc += h
    return oracle_np.layer_norm(acc, ln_w, ln_b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--d", type=int, default=128)
    args = ap.parse_args()
    from ctgcn_b200 import synth
    snap = synth.make_snapshot("er", args.n, 10 * args.n, args.k, seed=0)
    import scipy.sparse as sp
    mats = [sp.coo_matrix((a._values().numpy(), a._indices().numpy()), shape=(args.n, args.n)) for a in snap.coo_list()]
    x = synth.features(args.n, args.d, 1000).numpy().astype(np.float64)
    u = oracle_np.cumulative_core_sums(x, mats).transpose(1, 0, 2)
    torch.manual_seed(0)
    gru = torch.nn.GRU(args.d, args.d, batch_first=True)
    p = {k: v.detach().numpy().astype(np.float64) for k, v in gru.named_parameters()}
    ln_w, ln_b = np.ones(args.d), np.zeros(args.d)
    ref = gru_sum_ln(u, p["weight_ih_l0"], p["weight_hh_l0"], p["bias_ih_l0"], p["bias_hh_l0"], ln_w, ln_b, scheme_products("fp64"))
    print(f"ER N={args.n}, K={snap.k} core steps, {args.d}->{args.d}; |U| max {u.max():.1f}; bar: 1e-4")
    print(f"{'scheme':24s} {'bf16-MMA equivalents':>22s} {'relL2 of LN(sum h)':>20s}")
    for name in COST:
        if name == "fp64":
            continue
        got = gru_sum_ln(u, p["weight_ih_l0"], p["weight_hh_l0"], p["bias_ih_l0"], p["bias_hh_l0"], ln_w, ln_b, scheme_products(name))
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f"{name:24s} {str(COST[name]):>22s} {err:20.2e}")


if __name__ == "__main__":
    main()
