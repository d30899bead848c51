"""Row-chunked CoreDiffusion (ctgcn_set_workspace_cap): when the per-core sums [N, K, D] of a layer would exceed the cap
(BASELINE.json configs[4] at full size: 102 GB) the layer runs chunk by chunk.  Results must not depend on the chunking."""
import numpy as np
import pytest
import torch

from oracle import cases
from test_parity_gpu import coo, tsd

pytestmark = pytest.mark.gpu


@pytest.fixture
def cap(lib):
    yield lib.set_workspace_cap
    lib.set_workspace_cap(0)


@pytest.mark.parametrize("impl_code", ["simt", "auto"])
@pytest.mark.parametrize("name,cap_bytes", [("cd_nested_k5", 200 << 10), ("cd_uci_0404_500_128", 4 << 20), ("cd_nested_weighted", 1)])
def test_chunked_equals_unchunked_golden(name, cap_bytes, impl_code, cap, lib, cuda_device):
    import ctgcn_b200 as pkg
    from ctgcn_b200 import dist, plan as P
    lib.set_gru_impl(lib.IMPL_SIMT if impl_code == "simt" else lib.IMPL_AUTO)
    try:
        c = cases.load_case(name)
        m = c["meta"]
        mod = pkg.CoreDiffusion(m["d_in"], m["d_out"], bias=m["bias"]).to(cuda_device)
        mod.load_state_dict(tsd(c["sd"], cuda_device), strict=True)
        plan = P.build_plan_coo(coo(c["adj"], cuda_device), cuda_device)
        x = torch.from_numpy(c["x"]).to(cuda_device)
        n, h = plan.n_rows, m["d_out"]
        full_ws = lib.lib.ctgcn_core_diffusion_rnn_workspace_bytes(plan.handle, lib.CELL_GRU, m["d_in"], h)
        with torch.no_grad():
            ref = mod(x, plan)
        cap(cap_bytes)
        small_ws = lib.lib.ctgcn_core_diffusion_rnn_workspace_bytes(plan.handle, lib.CELL_GRU, m["d_in"], h)
        assert small_ws < full_ws
        before = lib.launch_count()
        with torch.no_grad():
            got = mod(x, plan)
            slices = dist.node_slices(n, 3)
            bufs = [torch.full((e - s + 1, 2, h), -5.0, device=cuda_device) for s, e in slices]
            ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=cuda_device)
            mod.forward_into(x, plan, scatter=(ptrs, 2 * h, h))
        assert lib.launch_count() - before >= 2 * 2 * 2            # at least two chunks × (SpMM + sequence kernel), twice
        assert torch.equal(got, ref), name
        for (s, e), b in zip(slices, bufs):
            assert torch.equal(b[: e - s, 1], ref[s:e]) and (b[: e - s, 0] == -5.0).all() and (b[e - s:] == -5.0).all()
    finally:
        lib.set_gru_impl(lib.IMPL_AUTO)


def test_chunked_whole_waves_midsize(cap, lib, cuda_device):
    """40 K nodes, K = 6, 128-d: a 60 MB cap gives chunks of one full wave (148 × 128 rows) of the persistent GRU kernel."""
    import ctgcn_b200 as pkg
    from ctgcn_b200 import synth
    n, d = 40_000, 128
    snap = synth.make_snapshot("er", n, 300_000, 6, seed=3)
    plan = snap.plan(cuda_device)
    sd = cases.core_diffusion_params(np.random.default_rng(0), "", d, d)
    mod = pkg.CoreDiffusion(d, d).to(cuda_device)
    mod.load_state_dict(tsd(sd, cuda_device))
    x = synth.features(n, d, 7).to(cuda_device)
    with torch.no_grad():
        ref = mod(x, plan)
    cap(60 << 20)
    ws = lib.lib.ctgcn_core_diffusion_rnn_workspace_bytes(plan.handle, lib.CELL_GRU, d, d)
    assert ws < 148 * 128 * snap.k * d * 4 + (4 << 20)
    before = lib.launch_count()
    with torch.no_grad():
        got = mod(x, plan)
    assert lib.launch_count() - before >= 3 * 2
    assert torch.equal(got, ref)
