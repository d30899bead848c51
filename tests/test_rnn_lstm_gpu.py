"""rnn_type='LSTM' (reference layers.py:27-28, models.py:234-235) on the GPU: the fp32 LSTM sequence kernel against the
fp64 oracle and against goldens generated from the unmodified reference.  Same tolerance as tests/test_parity_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import cases, oracle_np
from test_parity_gpu import close, coo, tsd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,steps,d_in,h,bias", [(300, 5, 128, 128, True), (77, 1, 128, 128, True), (130, 12, 128, 128, False),
                                                 (65, 3, 500, 128, True), (40, 4, 20, 24, True), (257, 7, 64, 32, True),
                                                 (129, 2, 256, 256, True), (64, 3, 32, 200, True)])
@pytest.mark.parametrize("mode", [0, 1])
def test_lstm_seq_kernel(n, steps, d_in, h, bias, mode, lib, cuda_device):
    from ctgcn_b200 import ops
    rng = np.random.default_rng(n + steps)
    sd = cases.gru_params(rng, "rnn.", d_in, h, bias, "LSTM")
    sd.update(cases.norm_params(rng, "norm.", h))
    seq = np.maximum(rng.standard_normal((n, steps, d_in)) * 3, 0).astype(np.float32)
    f64 = lambda k: None if k not in sd else sd[k].astype(np.float64)
    hs = oracle_np.lstm_sequence(seq.astype(np.float64), f64("rnn.weight_ih_l0"), f64("rnn.weight_hh_l0"), f64("rnn.bias_ih_l0"),
                                 f64("rnn.bias_hh_l0"))
    ref = oracle_np.layer_norm(hs.sum(axis=1) if mode == 0 else hs, f64("norm.weight"), f64("norm.bias"))
    d = tsd(sd, cuda_device)
    buf = torch.zeros(n, steps + 1, d_in + 4, device=cuda_device)       # strided input / output views
    buf[:, :steps, :d_in] = torch.from_numpy(seq).to(cuda_device)
    out = torch.full((n, steps + 2, h) if mode else (n, h + 3), 7.0, device=cuda_device)
    view = out[:, 1:steps + 1, :] if mode else out[:, :h]
    before = lib.launch_count()
    ops.rnn_seq(buf[:, :steps, :d_in], d["rnn.weight_ih_l0"], d["rnn.weight_hh_l0"], d.get("rnn.bias_ih_l0"),
                d.get("rnn.bias_hh_l0"), d["norm.weight"], d["norm.bias"], 1e-5, mode, out=view, cell=lib.CELL_LSTM)
    assert lib.launch_count() > before
    close(view.cpu().numpy(), ref, f"lstm n={n} L={steps} {d_in}->{h} mode={mode}")
    assert (out[:, 0] == 7.0).all() if mode else (out[:, h:] == 7.0).all()


def test_lstm_weight_shape_is_checked(lib, cuda_device):
    from ctgcn_b200 import ops
    seq = torch.zeros(4, 2, 8, device=cuda_device)
    w3 = torch.zeros(3 * 8, 8, device=cuda_device)
    ln = torch.ones(8, device=cuda_device)
    with pytest.raises(lib.CtgcnError):
        ops.rnn_seq(seq, w3, w3, None, None, ln, ln, 1e-5, 0, cell=lib.CELL_LSTM)   # GRU-shaped weights for an LSTM


@pytest.mark.parametrize("name", cases.golden_names("core_diffusion", rnn_type="LSTM"))
def test_core_diffusion_lstm_golden(name, lib, cuda_device):
    import ctgcn_b200 as pkg
    c = cases.load_case(name)
    m = c["meta"]
    mod = pkg.CoreDiffusion(m["d_in"], m["d_out"], bias=m["bias"], rnn_type="LSTM").to(cuda_device)
    mod.load_state_dict(tsd(c["sd"], cuda_device), strict=True)
    with torch.no_grad():
        y = mod(torch.from_numpy(c["x"]).to(cuda_device), coo(c["adj"], cuda_device))
    close(y.cpu().numpy(), c["expected"]["y"], name)


@pytest.mark.parametrize("name", cases.golden_names("cgcn", rnn_type="LSTM") + cases.golden_names("ctgcn", rnn_type="LSTM"))
def test_model_lstm_golden(name, lib, cuda_device):
    import grad_checks
    import ctgcn_b200 as pkg
    c = cases.load_case(name)
    m = c["meta"]
    mod = grad_checks.build_model(pkg, m, cuda_device)
    mod.load_state_dict(tsd(c["sd"], cuda_device), strict=True)
    xs, adj = grad_checks.model_inputs(c, cuda_device)
    with torch.no_grad():
        res = mod(xs, adj)
    out, trans = res if m["model_type"] == "S" else (res, None)
    rs = m["row_stride"]
    close(grad_checks.stack3(out).cpu().numpy()[:, ::rs], c["expected"]["y"], name)
    if trans is not None:
        close(grad_checks.stack3(trans).cpu().numpy()[:, ::rs], c["expected"]["trans"], name + ".trans")


def test_core_diffusion_lstm_scatter(lib, cuda_device):
    """The fused exchange epilogue (ctgcn_core_diffusion_rnn_fwd, y = NULL) with the LSTM cell."""
    import ctgcn_b200 as pkg
    from ctgcn_b200 import dist, plan as P
    c = cases.load_case("cd_lstm_k5")
    m = c["meta"]
    mod = pkg.CoreDiffusion(m["d_in"], m["d_out"], bias=m["bias"], rnn_type="LSTM").to(cuda_device)
    mod.load_state_dict(tsd(c["sd"], cuda_device), strict=True)
    plan = P.build_plan_coo(coo(c["adj"], cuda_device), cuda_device)
    x = torch.from_numpy(c["x"]).to(cuda_device)
    n, h, T, t = plan.n_rows, m["d_out"], 2, 1
    slices = dist.node_slices(n, 3)
    bufs = [torch.full((e - s + 1, T, h), -5.0, device=cuda_device) for s, e in slices]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=cuda_device)
    with torch.no_grad():
        ref = mod(x, plan)
        assert mod.forward_into(x, plan, scatter=(ptrs, T * h, t * h)) is None
    for (s, e), b in zip(slices, bufs):
        assert torch.equal(b[: e - s, t], ref[s:e])
        assert (b[: e - s, 0] == -5.0).all() and (b[e - s:] == -5.0).all()
