"""ctypes mirror of the STRUCTS and ENUMS of include/gsfm_ra.h -- no loader here: importing this module maps no
shared library (the CPU oracle's binding, oracle/ra_oracle.py, shares the struct layouts and must not pull in the
CUDA product library)."""
import ctypes as C
import numpy as np

ABI_VERSION = 3
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 128

# gsfm_ra_status
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NUMERIC = 0, -1, -2, -3, -4, -5

# gsfm_ra_error_type == theia::RotationErrorType (include/pairwise_rotation_error_quat.hpp:50-61)
QUATERNION_NORM, ROTATION_MAT_FNORM, QUATERNION_COSINE = 0, 1, 2
ANGLE_AXIS_COVARIANCE, ANGLE_AXIS, ANGLE_AXIS_INLIERS = 3, 4, 5
ANGLE_AXIS_COV_INLIERS, ANGLE_AXIS_COVTRACE, ANGLE_AXIS_COVNORM = 6, 7, 8
# translation averaging through the same solver (include/gsfm_pa.h): theia::PositionErrorType::BASELINE
POSITION_BASELINE = 16

# gsfm_ra_loss_kind
(LOSS_TRIVIAL, LOSS_HUBER, LOSS_SOFTLONE, LOSS_CAUCHY, LOSS_ARCTAN, LOSS_TOLERANT, LOSS_TUKEY,
 LOSS_LONEHALF, LOSS_LTWO, LOSS_GEMANMCCLURE, LOSS_MAGSAC3, LOSS_MAGSAC4, LOSS_MAGSAC9, LOSS_TABULATED) = range(14)
LOSS_FLAG_INVERSE = 1

SOLVER_PCG, SOLVER_DENSE_CHOLESKY, SOLVER_AUTO = 0, 1, 2
AUTO_DENSE_MAX_VIEWS = 1024
MIN_EDGES_PER_GPU = 500000

TERMINATION_FAILURE = 7
TERMINATION = {0: "NONE", 1: "FUNCTION_TOLERANCE", 2: "GRADIENT_TOLERANCE", 3: "PARAMETER_TOLERANCE",
               4: "MAX_ITERATIONS", 5: "MIN_RADIUS", 6: "INVALID_STEPS", 7: "FAILURE"}

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


class Loss(C.Structure):
    """gsfm_ra_loss: s -> scale * f(g(s)); f = kind/flags/p (or the table), g = inner_* (trivial by default)."""
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32), ("p", C.c_double * 4), ("scale", C.c_double),
                ("inner_kind", C.c_int32), ("inner_flags", C.c_uint32), ("inner_p", C.c_double * 4), ("inner_scale", C.c_double),
                ("table", C.POINTER(C.c_double)), ("table_min_exp", C.c_int32), ("table_octaves", C.c_int32),
                ("table_per_octave", C.c_int32), ("table_reserved", C.c_int32)]

    @classmethod
    def make(cls, kind, *params, inverse=False, scale=1.0):
        l = cls()
        l.kind = int(kind)
        l.flags = LOSS_FLAG_INVERSE if inverse else 0
        for k, v in enumerate(params):
            l.p[k] = float(v)
        l.scale = float(scale)
        l.inner_kind = LOSS_TRIVIAL
        l.inner_scale = 1.0
        return l

    @classmethod
    def compose(cls, outer, inner, scale=1.0):
        """ComposedLoss(f, g) (scripts/loss_functions.py:250-265): rho(s) = scale * f(g(s)).  `outer` must not itself
        be composed or scaled-inside; its scale multiplies the whole."""
        if outer.inner_kind != LOSS_TRIVIAL or inner.inner_kind != LOSS_TRIVIAL or inner.kind == LOSS_TABULATED or outer.kind == LOSS_TABULATED:
            raise ValueError("only one level of composition of closed-form losses has a native form")
        l = cls()
        C.memmove(C.byref(l), C.byref(outer), C.sizeof(cls))
        l.scale = (outer.scale or 1.0) * float(scale)
        l.inner_kind, l.inner_flags, l.inner_scale = inner.kind, inner.flags, (inner.scale or 1.0)
        for k in range(4):
            l.inner_p[k] = inner.p[k]
        return l


class Problem(C.Structure):
    _fields_ = [("num_views", C.c_uint32), ("num_edges", C.c_uint64), ("edge_i", _u32p), ("edge_j", _u32p),
                ("omega_ij", _dp), ("cov6", _dp), ("edge_weight", _dp), ("error_type", C.c_int32),
                ("total_pair_count", C.c_int32), ("orientation", _dp), ("fixed_view", C.c_int64)]


class PositionProblem(C.Structure):
    """gsfm_pa_problem (include/gsfm_pa.h)."""
    _fields_ = [("num_views", C.c_uint32), ("num_edges", C.c_uint64), ("edge_i", _u32p), ("edge_j", _u32p),
                ("position_2", _dp), ("orientation", _dp), ("edge_weight", _dp), ("fixed_view", C.c_int64),
                ("error_type", C.c_int32), ("reserved", C.c_int32)]


# gsfm_pa_error_type == theia::PositionErrorType (include/pairwise_translation_error_covariance.hpp:47-51)
PA_BASELINE, PA_COVARIANCE = 0, 1


class Options(C.Structure):
    _fields_ = [("loss", Loss), ("max_num_iterations", C.c_int32), ("jacobi_scaling", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("linear_solver", C.c_int32), ("pcg_max_iterations", C.c_int32),
                ("pcg_rtol", C.c_double), ("num_threads", C.c_int32), ("device", C.c_int32),
                ("verbose", C.c_int32), ("n_gpus", C.c_int32)]


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
                ("trace_capacity", C.c_int32), ("trace_size", C.c_int32), ("outer_iterations", C.c_int32), ("num_linear_unconverged", C.c_int32),
                ("last_weight_change", C.c_double), ("n_gpus_used", C.c_int32), ("reserved", C.c_int32)]


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
    o.linear_solver = SOLVER_AUTO
    o.pcg_max_iterations = 500
    o.pcg_rtol = 1e-10
    o.num_threads = 0
    o.device = -1
    o.verbose = 0
    o.n_gpus = 0
    return o


def clone(struct):
    """A bitwise copy of a ctypes struct (copy.copy refuses structs that hold pointers); Python-side attributes that keep
    pointed-to buffers alive (a tabulated loss's host table) are carried along."""
    out = type(struct)()
    C.memmove(C.byref(out), C.byref(struct), C.sizeof(type(struct)))
    for k, v in getattr(struct, "__dict__", {}).items():
        setattr(out, k, v)
    return out


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
                 error_type=ANGLE_AXIS, orientation=None, fixed_view=-1):
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
        # translation averaging (POSITION_BASELINE): omega_ij holds position_2, plus the global orientations and the fixed view
        self.orientation = None if orientation is None else as_f64(orientation, (self.num_views, 3))
        self.fixed_view = int(fixed_view)
        p.orientation = ptr(self.orientation)
        p.fixed_view = self.fixed_view
        self.c = p


