"""ctypes mirror of include/gsfm_ra.h and loader of the CUDA product library.

The product has NO CPU fallback: `lib()` raises if libgsfm_ra.so is missing, and
every compute entry point of the library fails with GSFM_RA_ERR_NO_DEVICE when no
CUDA device is visible.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# GSFM_RA_LIB: another build of the SAME library (kernel-tuning experiments, profiles/kernel_times.py)
LIB_PATH = os.environ.get("GSFM_RA_LIB") or os.path.join(HERE, "csrc", "libgsfm_ra.so")

ABI_VERSION = 1
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 128

# gsfm_ra_status
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NUMERIC = 0, -1, -2, -3, -4, -5

# gsfm_ra_error_type == theia::RotationErrorType (include/pairwise_rotation_error_quat.hpp:50-61)
QUATERNION_NORM, ROTATION_MAT_FNORM, QUATERNION_COSINE = 0, 1, 2
ANGLE_AXIS_COVARIANCE, ANGLE_AXIS, ANGLE_AXIS_INLIERS = 3, 4, 5
ANGLE_AXIS_COV_INLIERS, ANGLE_AXIS_COVTRACE, ANGLE_AXIS_COVNORM = 6, 7, 8

# gsfm_ra_loss_kind
(LOSS_TRIVIAL, LOSS_HUBER, LOSS_SOFTLONE, LOSS_CAUCHY, LOSS_ARCTAN, LOSS_TOLERANT, LOSS_TUKEY,
 LOSS_LONEHALF, LOSS_LTWO, LOSS_GEMANMCCLURE, LOSS_MAGSAC3, LOSS_MAGSAC4, LOSS_MAGSAC9) = range(13)
LOSS_FLAG_INVERSE = 1

SOLVER_PCG, SOLVER_DENSE_CHOLESKY = 0, 1

TERMINATION = {0: "NONE", 1: "FUNCTION_TOLERANCE", 2: "GRADIENT_TOLERANCE", 3: "PARAMETER_TOLERANCE",
               4: "MAX_ITERATIONS", 5: "MIN_RADIUS", 6: "INVALID_STEPS", 7: "FAILURE"}

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


class Loss(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32), ("p", C.c_double * 4), ("scale", C.c_double)]

    @classmethod
    def make(cls, kind, *params, inverse=False, scale=1.0):
        l = cls()
        l.kind = int(kind)
        l.flags = LOSS_FLAG_INVERSE if inverse else 0
        for k, v in enumerate(params):
            l.p[k] = float(v)
        l.scale = float(scale)
        return l


class Problem(C.Structure):
    _fields_ = [("num_views", C.c_uint32), ("num_edges", C.c_uint64), ("edge_i", _u32p), ("edge_j", _u32p),
                ("omega_ij", _dp), ("cov6", _dp), ("edge_weight", _dp), ("error_type", C.c_int32),
                ("reserved", C.c_int32)]


class Options(C.Structure):
    _fields_ = [("loss", Loss), ("max_num_iterations", C.c_int32), ("jacobi_scaling", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("linear_solver", C.c_int32), ("pcg_max_iterations", C.c_int32),
                ("pcg_rtol", C.c_double), ("num_threads", C.c_int32), ("device", C.c_int32),
                ("verbose", C.c_int32), ("reserved", C.c_int32)]


class Iteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_successful", C.c_int32), ("step_is_valid", C.c_int32),
                ("linear_iterations", C.c_int32), ("cost", C.c_double), ("candidate_cost", C.c_double),
                ("cost_change", C.c_double), ("model_cost_change", C.c_double), ("relative_decrease", C.c_double),
                ("gradient_max_norm", C.c_double), ("step_norm", C.c_double), ("trust_region_radius", C.c_double),
                ("linear_residual", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("total_linear_iterations", C.c_int64),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("ms_setup", C.c_double),
                ("ms_assemble", C.c_double), ("ms_linear", C.c_double), ("ms_cost", C.c_double),
                ("ms_total", C.c_double), ("kernel_launches", C.c_int64), ("trace", C.POINTER(Iteration)),
                ("trace_capacity", C.c_int32), ("trace_size", C.c_int32), ("outer_iterations", C.c_int32), ("reserved", C.c_int32),
                ("last_weight_change", C.c_double)]


def residual_dim(error_type):
    """Residual dimension of a RotationErrorType (include/pairwise_rotation_error_quat.hpp: 4 for QuatFNorm, 9 for RotFNorm)."""
    return {QUATERNION_NORM: 4, ROTATION_MAT_FNORM: 9}.get(int(error_type), 3)


def default_options_py():
    """The Ceres 1.14 defaults the reference runs with (SURVEY Appendix B.3,
    src/GSfM_nonlinear_rotation_estimator.cpp:299-303); the same numbers
    gsfm_ra_default_options() fills in on the C side."""
    o = Options()
    o.loss = Loss.make(LOSS_TRIVIAL)
    o.max_num_iterations = 200
    o.jacobi_scaling = 1
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.linear_solver = SOLVER_PCG
    o.pcg_max_iterations = 500
    o.pcg_rtol = 1e-10
    o.num_threads = 0
    o.device = -1
    o.verbose = 0
    return o


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def ptr(a, ctype=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


class ProblemArrays:
    """Owns contiguous numpy arrays and the gsfm_ra_problem that points at them."""

    def __init__(self, num_views, edge_i, edge_j, omega_ij, cov6=None, edge_weight=None,
                 error_type=ANGLE_AXIS):
        self.edge_i = np.ascontiguousarray(edge_i, dtype=np.uint32)
        self.edge_j = np.ascontiguousarray(edge_j, dtype=np.uint32)
        E = len(self.edge_i)
        self.omega_ij = as_f64(omega_ij, (E, 3))
        self.cov6 = None if cov6 is None else as_f64(cov6, (E, 6))
        self.edge_weight = None if edge_weight is None else as_f64(edge_weight, (E,))
        self.num_views = int(num_views)
        self.num_edges = E
        self.error_type = int(error_type)
        p = Problem()
        p.num_views = self.num_views
        p.num_edges = E
        p.edge_i = ptr(self.edge_i, C.c_uint32)
        p.edge_j = ptr(self.edge_j, C.c_uint32)
        p.omega_ij = ptr(self.omega_ij)
        p.cov6 = ptr(self.cov6)
        p.edge_weight = ptr(self.edge_weight)
        p.error_type = self.error_type
        self.c = p


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
    return lib


# every symbol include/gsfm_ra.h declares (checked by the CPU test-suite)
EXPORTED_SYMBOLS = [
    "gsfm_ra_abi_version", "gsfm_ra_last_error", "gsfm_ra_device_count", "gsfm_ra_default_options",
    "gsfm_ra_solve", "gsfm_ra_solve_sigma_consensus", "gsfm_ra_solver_create", "gsfm_ra_solver_create_sharded", "gsfm_ra_solver_destroy",
    "gsfm_ra_solver_set_rotations", "gsfm_ra_solver_get_rotations", "gsfm_ra_solver_reset",
    "gsfm_ra_solver_iterate", "gsfm_ra_comm_unique_id", "gsfm_ra_solver_comm_init",
    "gsfm_ra_solver_ipc_export", "gsfm_ra_solver_ipc_import", "gsfm_ra_solver_edge_range", "gsfm_ra_solver_cuda_stream", "gsfm_ra_solver_time_kernels", "gsfm_ra_eval_edges", "gsfm_ra_whiten", "gsfm_ra_assemble", "gsfm_ra_cost",
    "gsfm_ra_spmv", "gsfm_ra_pcg", "gsfm_ra_eval_loss", "gsfm_ra_filter_view_pairs", "gsfm_ra_residual_dim",
    "gsfm_ra_filter_initial_view_graph", "gsfm_ra_init_orientations_mst",
    "gsfm_ra_free", "gsfm_ra_read_covariance_rot", "gsfm_ra_write_covariance_rot", "gsfm_ra_read_1dsfm",
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
