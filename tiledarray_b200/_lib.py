"""ctypes binding of libtadev.so (the C ABI declared in include/tadev.h).

The library is built in-tree by ``__graft_entry__.build()`` (``tiledarray_b200/csrc/Makefile``)
and must be present: there is no Python or CPU fallback for any compute entry. Import fails
loudly when the shared object is missing.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtadev.so")

OK, EINVAL, ECUDA, ENODEVICE, ENCCL, ENOMEM = 0, 1, 2, 3, 4, 5
OP_N, OP_T = 0, 1


class TadevError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tadev error {code}: {msg}")
        self.code = code


class GemmTask(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("k", C.c_int32), ("reserved", C.c_int32)]


class GemmGroup(C.Structure):
    _fields_ = [("C", C.c_void_p), ("m", C.c_int32), ("n", C.c_int32), ("task_begin", C.c_int32),
                ("task_end", C.c_int32), ("accumulate", C.c_int32), ("raster", C.c_int32)]


class ProcGridC(C.Structure):
    _fields_ = [("proc_rows", C.c_int32), ("proc_cols", C.c_int32), ("proc_size", C.c_int32),
                ("rank_row", C.c_int32), ("rank_col", C.c_int32),
                ("local_rows", C.c_int64), ("local_cols", C.c_int64), ("local_size", C.c_int64)]


class ContractionPlanC(C.Structure):
    _fields_ = [("left_rank", C.c_int32), ("right_rank", C.c_int32), ("result_rank", C.c_int32),
                ("inner_rank", C.c_int32), ("opA", C.c_int32), ("opB", C.c_int32),
                ("left_permtype", C.c_int32), ("right_permtype", C.c_int32),
                ("perm_left", C.c_int32 * 16), ("perm_right", C.c_int32 * 16), ("perm_result", C.c_int32 * 16),
                ("left_target", C.c_char * 256), ("right_target", C.c_char * 256), ("result_gemm", C.c_char * 256)]


class SummaPlanC(C.Structure):
    _fields_ = [("Mt", C.c_int32), ("Nt", C.c_int32), ("Kt", C.c_int32),
                ("m_ext", C.POINTER(C.c_int64)), ("n_ext", C.POINTER(C.c_int64)), ("k_ext", C.POINTER(C.c_int64)),
                ("opA", C.c_int32), ("opB", C.c_int32), ("alpha", C.c_double),
                ("a_norms", C.POINTER(C.c_float)), ("b_norms", C.POINTER(C.c_float)), ("c_norms", C.POINTER(C.c_float)),
                ("threshold", C.c_float),
                ("a_tiles", C.POINTER(C.c_void_p)), ("b_tiles", C.POINTER(C.c_void_p)), ("c_tiles", C.POINTER(C.c_void_p)),
                ("accumulate", C.c_int32), ("depth", C.c_int32), ("steps_per_launch", C.c_int32),
                ("flags", C.c_int32), ("row_blocks", C.c_int32), ("reserved", C.c_int32),
                ("a_provider", C.c_void_p), ("a_user", C.c_void_p), ("b_provider", C.c_void_p), ("b_user", C.c_void_p)]


SUMMA_A_ON_HOST, SUMMA_B_ON_HOST, SUMMA_C_ON_HOST, SUMMA_A_LAZY, SUMMA_B_LAZY = 1, 2, 4, 8, 16


MEM_DEVICE, MEM_HOST, MEM_LAZY = 0, 1, 2
EW_AXPBY, EW_MULT = 0, 1


class ArrayDescC(C.Structure):
    _fields_ = [("rank", C.c_int32), ("memory", C.c_int32), ("bounds", C.POINTER(C.c_int64)), ("ntiles", C.POINTER(C.c_int32)),
                ("norms", C.POINTER(C.c_float)), ("tiles", C.POINTER(C.c_void_p)), ("lazy_seed", C.c_uint64),
                ("owners", C.POINTER(C.c_int32))]


class ContractOptionsC(C.Structure):
    _fields_ = [("exchange_operands", C.c_int32), ("stream_permutes", C.c_int32), ("stream_permute_bytes", C.c_int64),
                ("depth", C.c_int32), ("steps_per_launch", C.c_int32), ("row_blocks", C.c_int32), ("threshold", C.c_float),
                ("mask_norms", C.POINTER(C.c_float)), ("mask_threshold", C.c_float)]


class LayoutInfoC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("swapped", "Pr", "Pc", "Mt", "Nt", "Kt", "opA", "opB", "left_role", "right_role")]


class ContractionInfoC(C.Structure):
    _fields_ = [("rank", C.c_int32), ("swapped", C.c_int32), ("bounds", C.POINTER(C.c_int64)), ("ntiles", C.POINTER(C.c_int32)),
                ("norms", C.POINTER(C.c_float)), ("nzero", C.c_uint64), ("nlocal", C.c_int64),
                ("ordinals", C.POINTER(C.c_int64)), ("elems", C.POINTER(C.c_int64)), ("offsets", C.POINTER(C.c_int64)),
                ("arena_elems", C.c_int64), ("Pr", C.c_int32), ("Pc", C.c_int32), ("Mt", C.c_int32), ("Nt", C.c_int32),
                ("Kt", C.c_int32), ("opA", C.c_int32), ("opB", C.c_int32), ("needs_result_permute", C.c_int32)]


class UniformSourceC(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("seed", C.c_uint64)]


class PermuteSourceC(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("rank", C.c_int32), ("perm", C.c_int32 * 16),
                ("extents", C.POINTER(C.c_int64)), ("src", C.POINTER(C.c_void_p)),
                ("src_memory", C.c_int32), ("reserved", C.c_int32), ("lazy_seed", C.c_uint64),
                ("ordinals", C.POINTER(C.c_int64))]


class SummaStatsC(C.Structure):
    _fields_ = [("nsteps", C.c_int64), ("nsteps_skipped", C.c_int64), ("npairs", C.c_int64), ("nlaunches", C.c_int64),
                ("flops", C.c_double), ("bcast_bytes", C.c_int64), ("device_ms", C.c_float), ("row_blocks", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("lazy_tiles", C.c_int64),
                ("gemm_ms", C.c_float), ("list_ms", C.c_float)]


class ContractStatsC(C.Structure):
    _fields_ = [("summa", SummaStatsC), ("permute_ms", C.c_float)]


# every exported symbol of include/tadev.h with its prototype (restype, argtypes)
_vp, _i, _i64, _u64, _sz, _d, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_size_t, C.c_double, C.c_float
_P = C.POINTER
PROTOTYPES = {
    "tadev_last_error": (C.c_char_p, []),
    "tadev_version": (C.c_char_p, []),
    "tadev_init": (_i, [_i, _sz, _P(_vp)]),
    "tadev_finalize": (_i, [_vp]),
    "tadev_device_count": (_i, [_P(_i)]),
    "tadev_num_streams": (_i, [_vp, _P(_i)]),
    "tadev_get_stream": (_i, [_vp, _i, _P(_vp)]),
    "tadev_stream_for": (_i, [_vp, _u64, _P(_vp)]),
    "tadev_stream_sync": (_i, [_vp, _vp]),
    "tadev_alloc": (_i, [_vp, _sz, _P(_vp), _vp]),
    "tadev_free": (_i, [_vp, _vp, _vp]),
    "tadev_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "tadev_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "tadev_memset": (_i, [_vp, _vp, _i, _sz, _vp]),
    "tadev_event_create": (_i, [_vp, _P(_vp)]),
    "tadev_event_record": (_i, [_vp, _vp, _vp]),
    "tadev_event_elapsed_ms": (_i, [_vp, _vp, _vp, _P(_f)]),
    "tadev_event_destroy": (_i, [_vp, _vp]),
    "tadev_sync_event_create": (_i, [_vp, _P(_vp)]),
    "tadev_stream_wait_event": (_i, [_vp, _vp, _vp]),
    "tadev_event_query": (_i, [_vp, _vp, _P(_i)]),
    "tadev_event_sync": (_i, [_vp, _vp]),
    "tadev_stream_add_callback": (_i, [_vp, _vp, _vp, _vp]),
    "tadev_memcpy_d2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "tadev_host_alloc": (_i, [_sz, _P(_vp)]),
    "tadev_host_free": (_i, [_vp]),
    "tadev_gemm_grouped_f64": (_i, [_vp, _vp, _i, _i, _d, _P(GemmGroup), _i, _P(GemmTask), _i]),
    "tadev_gemm_grouped_f64_dev": (_i, [_vp, _vp, _i, _i, _d, _vp, _i, _vp, _vp, _i]),
    "tadev_gemm_f64": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _d, _vp]),
    "tadev_permute": (_i, [_vp, _vp, _i, _P(_i64), _P(C.c_int32), _i, _vp, _vp]),
    "tadev_permute_batched": (_i, [_vp, _vp, _i, _P(_i64), _P(C.c_int32), _i, _i, _P(_vp), _P(_vp)]),
    "tadev_add_to_f64": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "tadev_scale_f64": (_i, [_vp, _vp, _sz, _vp, _d]),
    "tadev_tile_sqnorms_f64": (_i, [_vp, _vp, _i, _vp, _vp, _i64, _vp]),
    "tadev_sqnorm_f64": (_i, [_vp, _vp, _sz, _vp, _P(_d)]),
    "tadev_tiles_binary_f64": (_i, [_vp, _vp, _i, _i, _P(_vp), _P(_vp), _P(_vp), _P(_i64), _d, _d]),
    "tadev_fill_uniform_f64": (_i, [_vp, _vp, _vp, _sz, _u64, _u64]),
    "tadev_shape_scale_f32": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _f, _vp]),
    "tadev_shape_gemm_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp]),
    "tadev_shape_mask_f32": (_i, [_vp, _vp, _i64, _vp, _vp, _f, _f, _vp]),
    "tadev_build_pairlist": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "tadev_build_tile_lists": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _vp, _vp, _i64, _vp]),
    "tadev_proc_grid_make": (_i, [_i, _i, _i64, _i64, _i64, _i64, _P(ProcGridC)]),
    "tadev_cyclic_owner": (_i, [_i64, _i64, _i, _i, _P(_i)]),
    "tadev_plan_contraction": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, _P(ContractionPlanC)]),
    "tadev_plan_contraction_opt": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, _P(ContractionPlanC), _P(C.c_int32)]),
    "tadev_contract_options_default": (_i, [_P(ContractOptionsC)]),
    "tadev_contraction_create": (_i, [_vp, C.c_char_p, C.c_char_p, C.c_char_p, _P(ArrayDescC), _P(ArrayDescC), _d,
                                      _P(ContractOptionsC), _P(_vp)]),
    "tadev_contraction_info_get": (_i, [_vp, _P(ContractionInfoC)]),
    "tadev_contraction_owner": (_i, [_vp, _i64, _P(_i)]),
    "tadev_contraction_eval": (_i, [_vp, _vp, _i, _i, _P(ContractStatsC)]),
    "tadev_contraction_eval_tiles": (_i, [_vp, _P(_vp), _i, _i, _P(ContractStatsC)]),
    "tadev_contraction_layout": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, _P(ArrayDescC), _P(ArrayDescC), _P(ContractOptionsC), _i,
                                      _P(LayoutInfoC), _vp, _vp, _vp, _vp]),
    "tadev_contraction_destroy": (_i, [_vp]),
    "tadev_elementwise_create": (_i, [_vp, _i, C.c_char_p, _d, C.c_char_p, _P(ArrayDescC), _d, C.c_char_p, _P(ArrayDescC), _f,
                                      _P(_vp)]),
    "tadev_elementwise_info_get": (_i, [_vp, _P(ContractionInfoC)]),
    "tadev_elementwise_eval": (_i, [_vp, _vp, _P(_f)]),
    "tadev_elementwise_destroy": (_i, [_vp]),
    "tadev_plan_general_product": (_i, [C.c_char_p, C.c_char_p, C.c_char_p, _P(ContractionPlanC), _P(C.c_int32)]),
    "tadev_comm_unique_id": (_i, [_vp]),
    "tadev_comm_init": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "tadev_comm_destroy": (_i, [_vp]),
    "tadev_shape_allreduce_max_f32": (_i, [_vp, _vp, _vp, _i64]),
    "tadev_bcast_panel": (_i, [_vp, _vp, _i, _i, _vp, _sz]),
    "tadev_exchange_tiles": (_i, [_vp, _vp, _i, _P(_vp), _P(_sz), _P(C.c_int32), _i, _P(_vp), _P(_sz), _P(C.c_int32)]),
    "tadev_summa_f64": (_i, [_vp, _P(SummaPlanC), _P(SummaStatsC)]),
    "tadev_provider_uniform": (_i, [_vp, _vp, _i, _P(_u64), _P(_vp), _P(_sz)]),
    "tadev_provider_permute": (_i, [_vp, _vp, _i, _P(_u64), _P(_vp), _P(_sz)]),
    "tadev_summa_schedule": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _P(C.c_int32), _vp, _vp,
                                  _i64, _P(_i64)]),
    "tadev_summa_steps": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                               _P(C.c_int32)]),
    "tadev_summa_comm_trace": (_i, [_i, _i, _i, _i, _P(SummaPlanC), _vp, _vp, _vp, _vp, _vp, _i64, _P(_i64)]),
    "tadev_probe_fp64_peak": (_i, [_vp, _i, _i, _P(_d), _P(_f)]),
    "tadev_probe_copy_gbs": (_i, [_vp, _sz, _i, _P(_d)]),
    "tadev_probe_pcie_gbs": (_i, [_vp, _sz, _P(_d), _P(_d), _P(_d), _P(_d)]),
    "tadev_launch_count": (_i, [_vp, _P(_i64)]),
}

_lib = None


def _preload_nccl() -> None:
    """libtadev.so needs libnccl.so.2. PyTorch ships a newer NCCL than the system one and both
    have the same SONAME, so whichever is mapped first serves the whole process: map the
    PyTorch-bundled copy first (when present) so that a later ``import torch`` still finds every
    symbol it was built against."""
    import sys
    for base in sys.path:
        cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
            except OSError:
                pass


def load() -> C.CDLL:
    """Load libtadev.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(tiledarray_b200 has no CPU or pure-Python compute path)")
    _preload_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise TadevError(rc, load().tadev_last_error().decode())
