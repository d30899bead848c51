"""ctypes binding of libctgcn_b200.so (the C-ABI declared in include/ctgcn_b200.h).

There is no Python/CPU fallback: if the shared library is missing, importing this module raises,
and every compute entry point raises ``CtgcnError`` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libctgcn_b200.so")

OK, EINVAL, ECUDA, ENOMEM, ENODEV = 0, -1, -2, -3, -4
ACT_NONE, ACT_SELU = 0, 1
GRU_SUM_LN, GRU_EACH_LN = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, IMPL_TC_ONE_CTA_R1, IMPL_TC_UNPAIRED, IMPL_TC_WIDE = 0, 1, 2, 3, 4, 5
CELL_GRU, CELL_LSTM = 0, 1
CELLS = {"GRU": CELL_GRU, "LSTM": CELL_LSTM}
MAX_CORES = 64
MAX_NEG = 64


class CtgcnError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first (python -m ctgcn_b200.build, or "
        "__graft_entry__.build()).  ctgcn_b200 has no CPU/PyTorch fallback path.")

lib = C.CDLL(LIB_PATH)

_p, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

SIGNATURES = {
    "ctgcn_version": (C.c_int, []),
    "ctgcn_last_error": (C.c_char_p, []),
    "ctgcn_launch_count": (_i64, []),
    "ctgcn_device_check": (C.c_int, []),
    "ctgcn_prof_enable": (C.c_int, [_i32]),
    "ctgcn_prof_collect": (C.c_int, [_p, _p, _i32]),
    "ctgcn_plan_create_coo": (C.c_int, [_i64, _i64, _i32, _p, _p, _p, _p, _i32, _p, _p]),
    "ctgcn_plan_create_csr": (C.c_int, [_i64, _i64, _i32, _p, _p, _p, _p, _i64, _i32, _p, _p]),
    "ctgcn_plan_destroy": (C.c_int, [_p]),
    "ctgcn_plan_stats": (C.c_int, [_p, _p]),
    "ctgcn_plan_arrays": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "ctgcn_cumspmm_fwd": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "ctgcn_cumspmm_fwd_ex": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p]),
    "ctgcn_cumspmm_bwd_workspace_bytes": (_sz, [_p, _i32]),
    "ctgcn_cumspmm_bwd": (C.c_int, [_p, _p, _i32, _p, _i64, _p, _sz, _p]),
    "ctgcn_gru_workspace_bytes": (_sz, [_i32, _i32]),
    "ctgcn_gru_seq_fwd": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _f32, _i32, _p, _i64,
                                    _i64, _p, _sz, _p]),
    "ctgcn_rnn_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "ctgcn_rnn_seq_fwd": (C.c_int, [_i32, _p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _f32, _i32, _p, _i64,
                                    _i64, _p, _sz, _p]),
    "ctgcn_set_gru_impl": (C.c_int, [_i32]),
    "ctgcn_set_fusion": (C.c_int, [_i32]),
    "ctgcn_debug_gru_trace": (C.c_int, [_p]),
    "ctgcn_selftest_umma": (C.c_int, [_p, _p, _p, _p, _p, _p, _sz, _p]),
    "ctgcn_selftest_umma_pair": (C.c_int, [_p, _p, _p, _p, _p, _p, _sz, _p]),
    "ctgcn_core_diffusion_workspace_bytes": (_sz, [_p, _i32, _i32]),
    "ctgcn_core_diffusion_fwd": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _f32, _p, _i64, _p, _sz, _p]),
    "ctgcn_core_diffusion_fwd_scatter": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _f32, _p, _i32, _i64, _i64,
                                                   _p, _sz, _p]),
    "ctgcn_core_diffusion_rnn_workspace_bytes": (_sz, [_p, _i32, _i32, _i32]),
    "ctgcn_core_diffusion_rnn_fwd": (C.c_int, [_p, _i32, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _f32, _p, _i64, _p, _i32,
                                               _i64, _i64, _p, _sz, _p]),
    "ctgcn_set_workspace_cap": (C.c_int, [_sz]),
    "ctgcn_linear_workspace_bytes": (_sz, [_i64, _i64]),
    "ctgcn_linear_fwd": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _i32, _p, _i64, _p, _sz, _p]),
    "ctgcn_spmm_linear_fwd": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _i64, _p, _sz, _p]),
    "ctgcn_kcore_numbers": (C.c_int, [_i64, _p, _p, _p]),
    "ctgcn_neg_sample": (C.c_int, [_p, _p, _i64, _p, _i64, _p, _i64, _i32, C.c_uint64, _p, _p, _p, _p]),
    "ctgcn_neg_loss_workspace_bytes": (_sz, [_i64, _i32]),
    "ctgcn_neg_loss_fwd": (C.c_int, [_p, _i64, _i64, _i32, _p, _i64, _p, _p, _p, _i32, _f32, _p, _p, _sz, _p]),
    "ctgcn_neg_loss_bwd": (C.c_int, [_p, _i64, _i64, _i32, _p, _i64, _p, _p, _p, _i32, _f32, _p, _p, _i64, _p, _sz, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.ctgcn_last_error() or b"").decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise CtgcnError(f"{what} failed (code {rc}): {last_error()}")


def launch_count() -> int:
    return int(lib.ctgcn_launch_count())


PROF_CLASSES = ("spmm", "gru", "linear", "pack", "spmm_linear")


def prof_enable(on: bool) -> None:
    check(lib.ctgcn_prof_enable(1 if on else 0), "ctgcn_prof_enable")


def prof_collect(reset: bool = True) -> dict:
    ms = (C.c_double * len(PROF_CLASSES))()
    cnt = (C.c_int64 * len(PROF_CLASSES))()
    check(lib.ctgcn_prof_collect(ms, cnt, 1 if reset else 0), "ctgcn_prof_collect")
    return {n: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, n in enumerate(PROF_CLASSES)}


def set_workspace_cap(nbytes: int) -> None:
    """Bound on the per-core-sums buffer of one CoreDiffusion call (0 = default 8 GiB); larger layers run in row chunks."""
    check(lib.ctgcn_set_workspace_cap(int(nbytes)), "ctgcn_set_workspace_cap")



def set_fusion(on: bool) -> None:
    """CoreDiffusion as one launch (default) or as the two-kernel path SpMM → U → GRU."""
    check(lib.ctgcn_set_fusion(1 if on else 0), "ctgcn_set_fusion")


def set_gru_impl(impl: int) -> None:
    check(lib.ctgcn_set_gru_impl(impl), "ctgcn_set_gru_impl")
