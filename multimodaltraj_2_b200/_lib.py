"""ctypes binding of ``lib/libmmt.so`` (the C-ABI declared in ``include/mmt.h``).

There is NO CPU fallback: if the shared library is missing it is built with nvcc; if it cannot be
built or loaded, importing the ops raises.  Every entry point returns an int status which is
turned into ``RuntimeError`` with ``mmt_last_error()``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

c_f = C.POINTER(C.c_float)
c_u8 = C.POINTER(C.c_uint8)
c_i32 = C.POINTER(C.c_int32)
vp = C.c_void_p


class CellWeights(C.Structure):
    _fields_ = [("W_e", vp), ("b_e", vp), ("W", vp), ("b", vp), ("w_If", vp), ("w_It", vp), ("w_Of", vp),
                ("w_Ot", vp), ("W_h", vp), ("b_h", vp), ("W_packed_bf16", vp), ("E", C.c_int), ("U", C.c_int),
                ("W_packed_bf16x3", vp), ("W_packed_f16", vp)]


class McrWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("W_i", "W_ii", "W_v", "b_v", "W_r", "W_c", "W_o")]


class EdgeWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("W1", "b1", "W2", "b2", "w_out", "b_out")] + [("He", C.c_int)]


class Shape(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("S", "N", "T", "P", "K", "U", "E", "He", "D", "relational", "prec", "img_h")]


class ForecastCfg(C.Structure):
    _fields_ = [("S", C.c_int), ("N", C.c_int), ("T", C.c_int), ("P", C.c_int), ("K", C.c_int),
                ("r2", C.c_float), ("inv_2sigma2", C.c_float), ("relational", C.c_int), ("prec", C.c_int),
                ("seed", C.c_uint64), ("agent_offset", C.c_uint64)]


# name -> (restype, argtypes); mirrors include/mmt.h one to one
SIGNATURES = {
    "mmt_version": (C.c_int, []),
    "mmt_last_error": (C.c_char_p, []),
    "mmt_launch_count": (C.c_uint64, []),
    "mmt_num_sms": (C.c_int, []),
    "mmt_last_trap": (C.c_int, [C.POINTER(C.c_uint32)]),
    "mmt_pairwise_adj_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, vp, vp]),
    "mmt_neighbor_index_i32": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "mmt_aggregate_f32": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "mmt_aggregate_transpose_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "mmt_edge_mlp_f32": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                   C.c_size_t, vp]),
    "mmt_edge_weights_packed_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mmt_pack_edge_weights_bf16": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "mmt_edge_mlp_bf16": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_size_t,
                                    vp]),
    "mmt_train_frame_inputs_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "mmt_train_gate_input_f32": (C.c_int, [vp] * 5 + [C.c_int] * 3 + [vp, vp]),
    "mmt_gsk_gates_packed_f32": (C.c_int, [vp] * 9 + [C.c_int] * 2 + [vp] * 4),
    "mmt_gsk_cell_backward_packed_f32": (C.c_int, [vp] * 12 + [C.c_int] * 2 + [vp] * 6),
    "mmt_train_backward_split_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "mmt_train_backward_merge_f32": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "mmt_attention_score_grad_f32": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "mmt_edge_mlp_backward_f32": (C.c_int, [vp] * 9 + [C.c_int] * 4 + [vp] * 7 + [C.c_size_t, vp]),
    "mmt_edge_mlp_backward_bf16": (C.c_int, [vp] * 8 + [C.c_int] * 4 + [vp] * 6 + [C.c_size_t, vp]),
    "mmt_gsk_cell": (C.c_int, [vp, vp, vp, vp, vp, vp, C.POINTER(CellWeights), C.c_int, C.c_int, vp, vp, vp, vp, vp,
                               C.c_int, vp, vp]),
    "mmt_gate_weights_packed_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mmt_pack_gate_weights_bf16": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "mmt_gate_weights_packed_x3_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mmt_pack_gate_weights_bf16x3": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "mmt_pack_gate_weights_f16": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "mmt_gridlstm_step_f32": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                        C.c_int, vp, vp, vp]),
    "mmt_mcr_step_f32": (C.c_int, [vp, vp, vp, vp, vp, C.POINTER(McrWeights), C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_float, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "mmt_mcr_forward_f32": (C.c_int, [vp, vp, vp, C.POINTER(McrWeights), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_int, vp, vp, vp, vp]),
    "mmt_mean_error_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "mmt_train_val_scores_f32": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "mmt_ade_fde_world_f32": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, C.c_float, C.c_float, vp, vp, vp, vp]),
    "mmt_static_context_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mmt_static_context_f32": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_float, vp, vp, vp,
                                         C.c_size_t, vp]),
    "mmt_sigmoid_f32": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "mmt_rowsoftmax_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
    "mmt_decode_score_f32": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                       C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "mmt_decode_score_dump_eps_f32": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                                C.c_int, vp, vp, vp, vp]),
    "mmt_scene_batch_f32": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                      vp, vp, vp]),
    "mmt_forecast_workspace_bytes": (C.c_size_t, [C.POINTER(ForecastCfg), C.c_int, C.c_int]),
    "mmt_forecast_f32": (C.c_int, [vp, vp, vp, C.POINTER(CellWeights), C.POINTER(EdgeWeights),
                                   C.POINTER(ForecastCfg), vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    "mmt_head_nll_f32": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, vp, vp, vp]),
    "mmt_gsk_gates_f32": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "mmt_gsk_cell_backward_f32": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp,
                                            vp, vp]),
    "mmt_gemm_tf32": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_float, C.c_int, vp]),
    "mmt_allreduce_f32": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "mmt_allreduce_max_f32": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "mmt_workspace_bytes": (C.c_int, [C.c_int, C.POINTER(Shape), C.POINTER(C.c_size_t)]),
    "mmt_rollout_bf16": (C.c_int, [vp, vp, vp, C.POINTER(CellWeights), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_float, vp, vp, vp]),
    "mmt_rollout_f16": (C.c_int, [vp, vp, vp, C.POINTER(CellWeights), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_float, vp, vp, vp]),
}

_lib = None


def lib_path() -> Path:
    import os
    alt = os.environ.get("MMT_LIB")     # another build of the same library (kernel experiments under scratch/)
    return Path(alt) if alt else _build.LIB


def load(build_if_missing: bool = True):
    """Load libmmt.so (building it first if the .so is absent).  Raises on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if path == _build.LIB and build_if_missing:
        _build.build()                   # no-op when the library matches the sources (content hash); else nvcc runs
    elif not path.exists():
        raise RuntimeError(f"{path} is missing; run `python -m multimodaltraj_2_b200.build`")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mmt_version() != 100:
        raise RuntimeError(f"libmmt version mismatch: {lib.mmt_version()}")
    _lib = lib
    return lib


def last_trap():
    """Record a trapping tcgen05 kernel left in host-mapped memory (mmt_last_trap), or None."""
    if _lib is None:
        return None
    out = (C.c_uint32 * 8)()
    if not _lib.mmt_last_trap(out):
        return None
    return {"site": hex(out[0]), "cta": out[1], "thread": out[2], "barrier_smem": hex(out[3]), "parity": out[4]}


def check(rc: int, what: str = "libmmt"):
    if rc != 0:
        msg = load().mmt_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
