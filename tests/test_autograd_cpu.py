"""Host-side logic of the backward pass, CPU only.

1. ``autograd.rnn_seq_bwd`` / ``layer_norm_bwd`` (device-agnostic torch code) against autograd of nn.GRU / nn.LSTM + LayerNorm.
2. ``plan.transpose_csr`` on small level-tagged CSRs.
3. The autograd Functions + module wiring against the gradients the UNMODIFIED reference produced (tests/golden/*grad*),
   with the CUDA entry points replaced by the oracle-backed stand-in of tests/fake_backend.py (the kernels themselves are
   checked on the GPU in tests/test_train_gpu.py)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import cases

GRAD_TOL = 2e-5


@pytest.mark.parametrize("cell_name", ["GRU", "LSTM"])
@pytest.mark.parametrize("bias", [True, False])
@pytest.mark.parametrize("mode", [0, 1])
def test_rnn_seq_bwd_matches_torch_autograd(cell_name, bias, mode, lib):
    from ctgcn_b200 import autograd as ag
    torch.manual_seed(3)
    n, L, d, H = 37, 5, 12, 16
    rnn = (nn.GRU if cell_name == "GRU" else nn.LSTM)(d, H, 1, bias=bias, batch_first=True).double()
    ln = nn.LayerNorm(H).double()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.2, 0.2)
    seq = torch.randn(n, L, d, dtype=torch.double, requires_grad=True)
    out, _ = rnn(seq)
    y = ln(out.sum(1)) if mode == 0 else ln(out)
    dy = torch.randn_like(y)
    params = [rnn.weight_ih_l0, rnn.weight_hh_l0] + ([rnn.bias_ih_l0, rnn.bias_hh_l0] if bias else []) + [ln.weight, ln.bias]
    want = torch.autograd.grad(y, [seq] + params, dy)
    with torch.no_grad():
        got = ag.rnn_seq_bwd(seq.detach(), lib.CELLS[cell_name], rnn.weight_ih_l0, rnn.weight_hh_l0,
                             rnn.bias_ih_l0 if bias else None, rnn.bias_hh_l0 if bias else None, ln.weight, ln.bias, ln.eps,
                             mode, dy)
    assert (got[3] is None) == (not bias) and (got[4] is None) == (not bias)
    got = [g for g in got if g is not None]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert float((a - b).norm() / b.norm()) < 1e-10


@pytest.mark.parametrize("cell_name", ["GRU", "LSTM"])
@pytest.mark.parametrize("bias", [True, False])
@pytest.mark.parametrize("mode", [0, 1])
def test_rnn_seq_bwd_in_row_chunks(cell_name, bias, mode, lib):
    """The backward runs over row chunks so that its intermediates stay bounded (round-1 advice): same dseq rows, parameter
    gradients equal up to the order of the sums; a single chunk is exactly the unchunked call; the default bound gives whole-tensor
    calls at test sizes and ≈ 32 K-row chunks at the bench shape."""
    from ctgcn_b200 import autograd as ag
    torch.manual_seed(7)
    n, L, d, H = 53, 4, 10, 8
    G = 4 if cell_name == "LSTM" else 3
    w_ih, w_hh = torch.randn(G * H, d, dtype=torch.double) * 0.3, torch.randn(G * H, H, dtype=torch.double) * 0.3
    b_ih, b_hh = (torch.randn(G * H, dtype=torch.double) * 0.1, torch.randn(G * H, dtype=torch.double) * 0.1) if bias else (None, None)
    ln_w, ln_b = torch.rand(H, dtype=torch.double) + 0.5, torch.randn(H, dtype=torch.double) * 0.1
    seq = torch.randn(n, L, d, dtype=torch.double)
    dy = torch.randn(n, H, dtype=torch.double) if mode == 0 else torch.randn(n, L, H, dtype=torch.double)
    args = (seq, lib.CELLS[cell_name], w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, 1e-5, mode, dy)
    whole = ag.rnn_seq_bwd(*args)
    for max_rows in (7, 16, 52, 53, 1000):
        part = ag.rnn_seq_bwd_chunked(*args, max_rows=max_rows)
        assert len(part) == len(whole)
        for a, b in zip(part, whole):
            assert (a is None) == (b is None)
            if a is not None:
                if max_rows >= n:
                    assert torch.equal(a, b)
                else:
                    assert float((a - b).norm() / b.norm()) < 1e-12
    assert torch.equal(ag.rnn_seq_bwd_chunked(*args)[0], whole[0])            # default bound: one chunk at this size
    assert 20_000 < ag._bwd_chunk_rows(10, 128, 128, 3) < 60_000              # cfg4's core GRU: chunks of a few 10^4 rows
    ag.set_backward_chunk_bytes(1 << 20)
    try:
        small = ag.rnn_seq_bwd_chunked(*args)                                 # 1 MiB bound → the 256-row floor → still one chunk here
        assert torch.equal(small[0], whole[0])
        assert ag._bwd_chunk_rows(10, 128, 128, 3) == 256
    finally:
        ag.set_backward_chunk_bytes(2 << 30)


def test_selu_bwd_from_output(lib):
    from ctgcn_b200 import autograd as ag
    v = torch.linspace(-4, 4, 101, dtype=torch.double, requires_grad=True)
    y = torch.selu(v)
    dy = torch.randn_like(y)
    (want,) = torch.autograd.grad(y, v, dy)
    assert torch.allclose(ag.selu_bwd_from_output(dy, y.detach()), want, rtol=1e-12, atol=1e-12)


def test_transpose_csr(lib):
    from ctgcn_b200.plan import transpose_csr
    rng = np.random.default_rng(0)
    n, m, k = 7, 5, 3
    dense = rng.random((n, m)) < 0.5
    rows, cols = np.nonzero(dense)
    lvl = rng.integers(0, k, rows.shape[0]).astype(np.uint8)
    lvl[::4] |= 128
    order = np.lexsort((cols, lvl & 127, rows))           # rows sorted by level
    rows, cols, lvl = rows[order], cols[order], lvl[order]
    val = rng.standard_normal(rows.shape[0]).astype(np.float32)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(np.bincount(rows, minlength=n))
    t_rowptr, t_col, t_val, t_lvl = [t.numpy() for t in transpose_csr(
        n, m, torch.from_numpy(rowptr), torch.from_numpy(cols.astype(np.int32)), torch.from_numpy(val), torch.from_numpy(lvl))]
    assert t_rowptr.shape == (m + 1,) and t_rowptr[-1] == rows.shape[0]
    t_rows = np.repeat(np.arange(m), np.diff(t_rowptr))
    a = {(r, c, l): v for r, c, l, v in zip(rows, cols, lvl, val)}
    b = {(c, r, l): v for r, c, l, v in zip(t_rows, t_col, t_lvl, t_val)}
    assert a == b
    for r in range(m):                                     # level order inside every transposed row
        assert (np.diff((t_lvl[t_rowptr[r]:t_rowptr[r + 1]] & 127).astype(int)) >= 0).all()


@pytest.mark.parametrize("name", cases.golden_names("core_diffusion", rnn_type=None, grads=True))
def test_core_diffusion_backward_wiring(name, lib, monkeypatch):
    import fake_backend
    import grad_checks
    grad_checks.core_diffusion_grad(fake_backend.install(monkeypatch), cases.load_case(name), "cpu", GRAD_TOL)


@pytest.mark.parametrize("name", cases.golden_names("mlp", grads=True))
def test_mlp_backward_wiring(name, lib, monkeypatch):
    import fake_backend
    import grad_checks
    grad_checks.mlp_grad(fake_backend.install(monkeypatch), cases.load_case(name), "cpu", GRAD_TOL)


@pytest.mark.parametrize("name", cases.golden_names("cdn", rnn_type=None, grads=True))
def test_cdn_backward_wiring(name, lib, monkeypatch):
    import fake_backend
    import grad_checks
    grad_checks.cdn_grad(fake_backend.install(monkeypatch), cases.load_case(name), "cpu", GRAD_TOL)


@pytest.mark.parametrize("name", cases.golden_names("ctgcn", rnn_type=None, grads=True) + cases.golden_names("cgcn", rnn_type=None, grads=True))
def test_model_backward_wiring(name, lib, monkeypatch):
    import fake_backend
    import grad_checks
    grad_checks.model_grad(fake_backend.install(monkeypatch), cases.load_case(name), "cpu", GRAD_TOL)
