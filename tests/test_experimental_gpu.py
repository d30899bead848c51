"""Opt-in GPU checks of experimental building blocks that were written without a GPU (round 1, budget spent):
    CTGCN_UNVERIFIED_GPU_TESTS=1 python -m pytest tests/test_experimental_gpu.py -m gpu
* ctgcn_selftest_umma_pair — one GRU half-step through tcgen05.mma.cta_group::2 on a CTA pair (csrc/umma2_selftest.cu,
  profiles/r02_gru_design.md step 2).  A trap or a wrong block tells which mechanism is off: columns [0,64) only the N = 192
  input stream, [64,192) input + the N = 128 recurrent stream, [192,256) the N = 64 stream; rows 128.. are the follower CTA.
* ops.cumspmm_packed — the cumulative SpMM storing U pre-split in the operand layout (csrc/spmm_packed.cu, design note step 3):
  bit for bit the bf16 hi / lo planes of the fp32 kernel's output.
* ops.core_diffusion_packed — that SpMM feeding gru_tc_packed_kernel (U fetched by bulk copies, 16 gate warps, Σh in registers):
  the default CoreDiffusion result up to the summation order inside the LayerNorm."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CTGCN_UNVERIFIED_GPU_TESTS") != "1",
                                 reason="written without a GPU: opt in with CTGCN_UNVERIFIED_GPU_TESTS=1")]


def test_umma_pair_selftest(lib, cuda_device):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((256, 64)).astype(np.float32)
    h = rng.uniform(-1, 1, (256, 128)).astype(np.float32)
    w_ih = rng.uniform(-0.1, 0.1, (384, 64)).astype(np.float32)
    w_hh = rng.uniform(-0.1, 0.1, (384, 128)).astype(np.float32)
    t = [torch.from_numpy(a).to(cuda_device) for a in (x, h, w_ih, w_hh)]
    out = torch.zeros(256, 256, device=cuda_device)
    ws = torch.zeros(512 * 1024, dtype=torch.uint8, device=cuda_device)
    rc = lib.lib.ctgcn_selftest_umma_pair(*[C.c_void_p(a.data_ptr()) for a in t], C.c_void_p(out.data_ptr()),
                                          C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    lib.check(rc, "ctgcn_selftest_umma_pair")
    torch.cuda.synchronize()
    x64, h64, wi, wh = (a.astype(np.float64) for a in (x, h, w_ih, w_hh))
    ref = np.concatenate([x64 @ wi[256:320].T,
                          x64 @ wi[0:64].T + h64 @ wh[0:64].T,
                          x64 @ wi[128:192].T + h64 @ wh[128:192].T,
                          h64 @ wh[256:320].T], axis=1)
    got = out.cpu().numpy()
    for cta in range(2):
        rows = slice(128 * cta, 128 * cta + 128)
        for blk, name in enumerate(("W_in x", "r", "z", "W_hn h")):
            err = cases.relerr(got[rows, 64 * blk:64 * blk + 64], ref[rows, 64 * blk:64 * blk + 64])
            assert err < 3e-5, (f"CTA {cta}", name, err)


@pytest.mark.parametrize("n,m,k", [(1000, 6000, 4), (128, 900, 3), (4099, 30000, 7)])
def test_cumspmm_packed_matches_the_split_of_the_fp32_kernel(n, m, k, lib, cuda_device):
    from ctgcn_b200 import ops, synth
    snap = synth.make_snapshot("er", n, m, k, seed=n)
    plan = snap.plan(cuda_device)
    x = synth.features(n, 128, 7).to(cuda_device)
    u = ops.cumspmm(plan, x)                                            # [N, K, 128] fp32
    packed = ops.cumspmm_packed(plan, x)                                # uint8 [tiles, K, 2, 16, 128, 16]
    tiles = packed.shape[0]
    bits = packed.view(torch.int16).view(tiles, plan.k, 2, 16, 128, 8)   # [tile, level, plane, kb, row, k%8]
    rows = bits.permute(2, 0, 4, 1, 3, 5).reshape(2, tiles * 128, plan.k, 128)[:, :n]   # [plane, row, level, feature]
    hi = u.to(torch.bfloat16)
    lo = (u - hi.float()).to(torch.bfloat16)
    assert torch.equal(rows[0], hi.view(torch.int16)) and torch.equal(rows[1], lo.view(torch.int16))
    if tiles * 128 > n:                                                 # rows beyond N are left untouched
        assert int(bits.permute(2, 0, 4, 1, 3, 5).reshape(2, tiles * 128, plan.k, 128)[:, n:].abs().sum()) == 0


@pytest.mark.parametrize("n,m,k,bias", [(1000, 6000, 4, True), (300, 2000, 1, True), (19000, 150000, 6, False)])
def test_core_diffusion_packed_matches_default(n, m, k, bias, lib, cuda_device):
    from ctgcn_b200 import ops, synth
    snap = synth.make_snapshot("er", n, m, k, seed=n + 1)
    plan = snap.plan(cuda_device)
    x = synth.features(n, 128, 3).to(cuda_device)
    sd = cases.core_diffusion_params(np.random.default_rng(k), "", 128, 128, bias)
    d = {key: torch.from_numpy(v).to(cuda_device) for key, v in sd.items()}
    args = (d["rnn.weight_ih_l0"], d["rnn.weight_hh_l0"], d.get("rnn.bias_ih_l0"), d.get("rnn.bias_hh_l0"), d["norm.weight"], d["norm.bias"], 1e-5)
    want = ops.core_diffusion(plan, x, *args)
    got = ops.core_diffusion_packed(plan, x, *args)
    err = ((got - want).norm() / want.norm()).item()
    assert err < 1e-6, err
