"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/ctgcn_b200.h declares;
argument validation and the host-only helper work without a GPU."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ctgcn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctgcn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib.lib, n), f"{n} declared in include/ctgcn_b200.h but not exported"
    assert set(names) == set(lib.SIGNATURES), set(names) ^ set(lib.SIGNATURES)
    assert lib.lib.ctgcn_version() == 100


def test_argument_validation_without_gpu(lib):
    L = lib.lib
    out = C.c_void_p()
    nnz = (C.c_int64 * 1)(0)
    assert L.ctgcn_plan_create_coo(0, 0, 1, None, None, None, nnz, 0, None, C.byref(out)) == lib.EINVAL
    assert "empty shape" in lib.last_error()
    assert L.ctgcn_plan_create_coo(4, 4, 65, None, None, None, nnz, 0, None, C.byref(out)) == lib.EINVAL
    assert L.ctgcn_set_gru_impl(7) == lib.EINVAL
    assert L.ctgcn_set_gru_impl(lib.IMPL_AUTO) == 0
    assert L.ctgcn_gru_workspace_bytes(128, 128) > 0
    # the selector codes of include/ctgcn_b200.h and of the binding agree, and the wide-state kernel's chunk buffers are only
    # charged to layers that use it (H > 128, or selected explicitly)
    hdr = open(os.path.join(ROOT, "include", "ctgcn_b200.h")).read()
    for name in ("AUTO", "SIMT", "TCGEN05", "TC_ONE_CTA_R1", "TC_UNPAIRED", "TC_WIDE"):
        assert int(re.search(rf"#define CTGCN_IMPL_{name} (\d+)", hdr).group(1)) == getattr(lib, f"IMPL_{name}")
    small, wide = L.ctgcn_gru_workspace_bytes(128, 128), L.ctgcn_gru_workspace_bytes(256, 256)
    assert small < (8 << 20) and (100 << 20) < wide < (160 << 20)
    assert L.ctgcn_set_gru_impl(lib.IMPL_TC_WIDE) == 0
    assert L.ctgcn_gru_workspace_bytes(128, 128) > (100 << 20)
    assert L.ctgcn_set_gru_impl(lib.IMPL_AUTO) == 0
    assert L.ctgcn_linear_workspace_bytes(10, 20) >= 800
    assert L.ctgcn_cumspmm_fwd(None, None, 0, 4, None, None) == lib.EINVAL
    assert L.ctgcn_plan_destroy(None) == 0
    # entry points added for rnn_type / autograd / chunking: argument checks happen before any CUDA call
    assert L.ctgcn_cumspmm_fwd_ex(None, None, 0, 4, 0, None, None) == lib.EINVAL
    assert L.ctgcn_cumspmm_bwd(None, None, 4, None, 4, None, 0, None) == lib.EINVAL
    assert L.ctgcn_cumspmm_bwd_workspace_bytes(None, 4) == 0
    assert L.ctgcn_rnn_workspace_bytes(lib.CELL_LSTM, 128, 128) == 4 * 128 * 256 * 4         # k-major fp32 copies of both matrices
    assert L.ctgcn_rnn_workspace_bytes(lib.CELL_GRU, 128, 128) == L.ctgcn_gru_workspace_bytes(128, 128)
    assert L.ctgcn_rnn_workspace_bytes(7, 128, 128) == 0
    assert L.ctgcn_rnn_seq_fwd(7, None, 0, 0, 0, 1, 1, 1, None, None, None, None, None, None, 1e-5, 0, None, 0, 0, None, 0, None) == lib.EINVAL
    assert "unknown cell" in lib.last_error()
    assert L.ctgcn_core_diffusion_rnn_fwd(None, 0, None, 0, 1, 1, None, None, None, None, None, None, 1e-5, None, 0, None, 0, 0, 0,
                                          None, 0, None) == lib.EINVAL
    assert L.ctgcn_set_workspace_cap(1 << 20) == 0 and L.ctgcn_set_workspace_cap(0) == 0
    # negative-sampling loss entry points
    one = C.c_void_p(8)
    assert L.ctgcn_neg_sample(None, None, 10, None, 100, None, 4, 20, 1, None, None, None, None) == lib.EINVAL
    assert L.ctgcn_neg_sample(one, one, 10, one, 100, one, 4, lib.MAX_NEG + 1, 1, one, one, one, None) == lib.EINVAL
    assert L.ctgcn_neg_sample(one, one, 10, one, 5, one, 4, 20, 1, one, one, one, None) == lib.EINVAL
    assert "fewer than neg_num" in lib.last_error()
    assert L.ctgcn_neg_loss_workspace_bytes(4, 128) == 2 * 512 + 256 + 4 * 8 and L.ctgcn_neg_loss_workspace_bytes(-1, 128) == 0
    assert L.ctgcn_neg_loss_fwd(one, 128, 10, 128, one, 4, one, one, one, 20, 1.0, one, None, 0, None) == lib.ENOMEM
    assert L.ctgcn_neg_loss_fwd(one, 64, 10, 128, one, 4, one, one, one, 20, 1.0, one, one, 1 << 20, None) == lib.EINVAL   # ld < d
    assert L.ctgcn_neg_loss_bwd(one, 128, 10, 128, one, 4, one, one, one, 20, 1.0, None, None, 128, one, 1 << 20, None) == lib.EINVAL


def test_kcore_numbers_match_networkx(lib):
    import networkx as nx
    from ctgcn_b200 import synth
    rng = np.random.default_rng(3)
    n = 400
    u, v = synth.er_edges(n, 3000, rng)
    core = synth.core_numbers(n, u, v)
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(zip(u.tolist(), v.tolist()))
    ref = nx.core_number(g)
    assert core.tolist() == [ref[i] for i in range(n)]
    # star + isolated nodes
    core2 = synth.core_numbers(6, np.array([0, 0, 0]), np.array([1, 2, 3]))
    assert core2.tolist() == [1, 1, 1, 1, 0, 0]
