#!/usr/bin/env python
"""Timing of the dense-layer kernels alone (CUDA events, inputs larger than the L2, 20 launches after 3 warm-ups).
    python profiles/linear_ab.py
Prints per shape: ms per launch, GB/s of algorithmic traffic 4·N·(d_in + d_out), TFLOP/s, relL2 against an fp64 product."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    import __graft_entry__
    __graft_entry__.build()
    from ctgcn_b200 import _lib, ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    for n, d_in, d_out, act in ((1_000_000, 128, 128, 0), (1_000_000, 128, 128, 1), (1_000_000, 64, 128, 0), (500_000, 256, 256, 0),
                                (60_730, 204, 500, 1), (60_730, 500, 500, 1), (60_730, 500, 128, 1), (400_000, 512, 512, 0)):
        x = torch.randn(n, d_in, generator=g).to(dev)
        w = (torch.randn(d_out, d_in, generator=g) / d_in ** 0.5).to(dev)
        b = torch.randn(d_out, generator=g).to(dev)
        y = ops.linear(x, w, b, act)
        ref = x[:4096].double() @ w.double().t() + b.double()
        if act:
            ref = torch.nn.functional.selu(ref)
        err = ((y[:4096].double() - ref).norm() / ref.norm()).item()
        for _ in range(3):
            ops.linear(x, w, b, act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.linear(x, w, b, act)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"linear {n:8d} x {d_in:4d} -> {d_out:4d} act={act}: {ms:7.3f} ms  {4 * n * (d_in + d_out) / ms / 1e6:7.0f} GB/s  "
              f"{2 * n * d_in * d_out / ms / 1e9:6.1f} TFLOP/s  relL2 {err:.1e}  (incl. weight pack + output allocation)", flush=True)


if __name__ == "__main__":
    main()
