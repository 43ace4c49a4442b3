"""Loader of the CUDA product library + ctypes signatures of every entry point of include/gsfm_ra.h.

The product has NO CPU fallback: `lib()` raises if libgsfm_ra.so is missing, and
every compute entry point of the library fails with GSFM_RA_ERR_NO_DEVICE when no
CUDA device is visible.  The struct / enum mirror lives in _abi.py (loader-free).
"""
import ctypes as C
import os

import numpy as np  # noqa: F401

from ._abi import *  # noqa: F401,F403
from ._abi import _dp, _u32p, _u8p  # noqa: F401
from . import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
# GSFM_RA_LIB: another build of the SAME library (kernel-tuning experiments, profiles/kernel_times.py)
LIB_PATH = os.environ.get("GSFM_RA_LIB") or os.path.join(HERE, "csrc", "libgsfm_ra.so")
ABI_VERSION = _abi.ABI_VERSION


def declare(lib):
    """Attach argtypes/restypes of every entry point of include/gsfm_ra.h."""
    pp, lp, op, sp = C.POINTER(Problem), C.POINTER(Loss), C.POINTER(Options), C.POINTER(Summary)
    vp = C.c_void_p
    lib.gsfm_ra_abi_version.restype = C.c_int
    lib.gsfm_ra_last_error.restype = C.c_char_p
    lib.gsfm_ra_device_count.restype = C.c_int
    lib.gsfm_ra_residual_dim.argtypes = [C.c_int32]
    lib.gsfm_ra_residual_dim.restype = C.c_int
    lib.gsfm_ra_default_options.argtypes = [op]
    lib.gsfm_ra_default_options.restype = None
    lib.gsfm_ra_solve.argtypes = [pp, op, _dp, sp]
    lib.gsfm_ra_solve_sigma_consensus.argtypes = [pp, op, C.c_int32, C.c_double, _dp, sp]
    lib.gsfm_ra_solver_create.argtypes = [pp, op, C.POINTER(vp)]
    lib.gsfm_ra_solver_create_sharded.argtypes = [pp, op, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.gsfm_ra_solver_destroy.argtypes = [vp]
    lib.gsfm_ra_solver_destroy.restype = None
    lib.gsfm_ra_solver_set_rotations.argtypes = [vp, _dp]
    lib.gsfm_ra_solver_get_rotations.argtypes = [vp, _dp]
    lib.gsfm_ra_solver_reset.argtypes = [vp]
    lib.gsfm_ra_solver_iterate.argtypes = [vp, C.c_int32, sp]
    lib.gsfm_ra_comm_unique_id.argtypes = [_u8p]
    lib.gsfm_ra_solver_comm_init.argtypes = [vp, _u8p]
    lib.gsfm_ra_solver_ipc_export.argtypes = [vp, _u8p]
    lib.gsfm_ra_solver_ipc_import.argtypes = [vp, _u8p]
    lib.gsfm_ra_solver_edge_range.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.gsfm_ra_solver_cuda_stream.argtypes = [vp]
    lib.gsfm_ra_solver_cuda_stream.restype = C.c_void_p
    lib.gsfm_ra_solver_time_kernels.argtypes = [vp, C.c_int32, _dp]
    lib.gsfm_ra_solver_info.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.gsfm_ra_measure_stream.argtypes = [C.c_uint64, C.c_int32, C.c_int32, _dp]
    lib.gsfm_ra_eval_edges.argtypes = [pp, lp, _dp, _dp, _dp, _dp, _dp, C.c_int32]
    lib.gsfm_ra_whiten.argtypes = [pp, _dp, C.c_int32]
    lib.gsfm_ra_assemble.argtypes = [pp, lp, _dp, _dp, _dp, _dp, _u32p, _u32p, _dp, C.c_int32]
    lib.gsfm_ra_cost.argtypes = [pp, lp, _dp, _dp, C.c_int32]
    lib.gsfm_ra_spmv.argtypes = [pp, lp, _dp, _dp, _dp, _dp, C.c_int32]
    lib.gsfm_ra_pcg.argtypes = [pp, lp, _dp, _dp, _dp, C.c_double, C.c_int32, _dp, C.POINTER(C.c_int32), _dp, C.c_int32]
    lib.gsfm_ra_eval_loss.argtypes = [lp, _dp, C.c_uint64, _dp, C.c_int32]
    lib.gsfm_ra_filter_view_pairs.argtypes = [pp, _dp, C.c_double, _u8p, _dp, C.c_int32]
    _i32p = C.POINTER(C.c_int32)
    lib.gsfm_ra_filter_initial_view_graph.argtypes = [C.c_uint32, C.c_uint64, _u32p, _u32p, _i32p, C.c_int32, _u8p, _u8p, C.c_int32]
    lib.gsfm_ra_init_orientations_mst.argtypes = [C.c_uint32, C.c_uint64, _u32p, _u32p, _dp, _i32p, C.c_int64, _dp, _u8p, _i32p, C.c_int32]
    PP = C.POINTER
    lib.gsfm_ra_free.argtypes = [C.c_void_p]
    lib.gsfm_ra_free.restype = None
    lib.gsfm_ra_read_covariance_rot.argtypes = [C.c_char_p, PP(C.c_uint64), PP(_u32p), PP(_u32p), PP(_dp), PP(_dp)]
    lib.gsfm_ra_write_covariance_rot.argtypes = [C.c_char_p, C.c_uint64, _u32p, _u32p, _dp, _dp]
    lib.gsfm_ra_read_1dsfm.argtypes = [C.c_char_p, PP(C.c_uint32), PP(C.c_uint64), PP(_u32p), PP(_dp), PP(C.c_uint64), PP(_u32p), PP(_u32p),
                                       PP(_dp), PP(_dp), PP(_i32p)]
    # include/gsfm_pa.h (translation averaging)
    qp = C.POINTER(PositionProblem)
    lib.gsfm_pa_default_options.argtypes = [op]
    lib.gsfm_pa_default_options.restype = None
    lib.gsfm_pa_as_ra_problem.argtypes = [qp, pp]
    lib.gsfm_pa_solve.argtypes = [qp, op, _dp, sp]
    lib.gsfm_pa_eval_edges.argtypes = [qp, lp, _dp, _dp, _dp, _dp, _dp, C.c_int32]
    lib.gsfm_pa_cost.argtypes = [qp, lp, _dp, _dp, C.c_int32]
    return lib


# every symbol include/gsfm_ra.h and include/gsfm_pa.h declare (checked by the CPU test-suite)
EXPORTED_SYMBOLS = [
    "gsfm_ra_abi_version", "gsfm_ra_last_error", "gsfm_ra_device_count", "gsfm_ra_default_options",
    "gsfm_ra_solve", "gsfm_ra_solve_sigma_consensus", "gsfm_ra_solver_create", "gsfm_ra_solver_create_sharded", "gsfm_ra_solver_destroy",
    "gsfm_ra_solver_set_rotations", "gsfm_ra_solver_get_rotations", "gsfm_ra_solver_reset",
    "gsfm_ra_solver_iterate", "gsfm_ra_comm_unique_id", "gsfm_ra_solver_comm_init",
    "gsfm_ra_solver_ipc_export", "gsfm_ra_solver_ipc_import", "gsfm_ra_solver_edge_range", "gsfm_ra_solver_cuda_stream", "gsfm_ra_solver_time_kernels", "gsfm_ra_solver_info", "gsfm_ra_measure_stream", "gsfm_ra_eval_edges", "gsfm_ra_whiten", "gsfm_ra_assemble", "gsfm_ra_cost",
    "gsfm_ra_spmv", "gsfm_ra_pcg", "gsfm_ra_eval_loss", "gsfm_ra_filter_view_pairs", "gsfm_ra_residual_dim",
    "gsfm_ra_filter_initial_view_graph", "gsfm_ra_init_orientations_mst",
    "gsfm_ra_free", "gsfm_ra_read_covariance_rot", "gsfm_ra_write_covariance_rot", "gsfm_ra_read_1dsfm",
    "gsfm_pa_default_options", "gsfm_pa_as_ra_problem", "gsfm_pa_solve", "gsfm_pa_eval_edges", "gsfm_pa_cost",
]

_lib = None


def lib():
    """Load csrc/libgsfm_ra.so (built by __graft_entry__.build()); raise loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        _lib = declare(C.CDLL(LIB_PATH))
        if _lib.gsfm_ra_abi_version() != ABI_VERSION:
            raise RuntimeError("libgsfm_ra.so ABI version mismatch")
    return _lib


class GsfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gsfm_ra error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != 0:
        raise GsfmError(rc, (lib().gsfm_ra_last_error() or b"").decode())
