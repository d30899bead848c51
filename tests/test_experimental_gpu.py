"""Building-block self test of the 2-CTA tensor-core path (green on a B200 since round 2, call 1):
* ctgcn_selftest_umma_pair — one GRU half-step through tcgen05.mma.cta_group::2 on a CTA pair (csrc/umma2_selftest.cu,
  profiles/r02_gru_design.md step 2).  A trap or a wrong block tells which mechanism is off: columns [0,64) only the N = 192
  input stream, [64,192) input + the N = 128 recurrent stream, [192,256) the N = 64 stream; rows 128.. are the follower CTA."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = [pytest.mark.gpu]


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
