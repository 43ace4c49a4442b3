"""Host API of the translation-averaging path over the C ABI of include/gsfm_pa.h.

Mirrors GSfMNonlinearPositionEstimator::EstimatePositions (reference src/GSfM_nonlinear_position_estimator.cpp:87-234):
camera-to-camera constraints only (flags_1dsfm.yaml: position_estimation_min_num_tracks_per_view = 0), every position
starts at the origin, one view is held constant.  Everything here calls libgsfm_ra.so; nothing computes on the CPU and
nothing here imports the oracle."""
import ctypes as C

import numpy as np

from . import _capi as capi


class PositionProblemArrays:
    """Owns contiguous numpy arrays, the gsfm_pa_problem pointing at them, and the equivalent gsfm_ra_problem
    (error type POSITION_BASELINE) for the kernel-level entry points of include/gsfm_ra.h."""

    def __init__(self, num_views, edge_i, edge_j, position_2, orientation, edge_weight=None, fixed_view=0,
                 error_type=capi.PA_BASELINE):
        self.num_views = int(num_views)
        self.edge_i = np.ascontiguousarray(edge_i, dtype=np.uint32)
        self.edge_j = np.ascontiguousarray(edge_j, dtype=np.uint32)
        E = len(self.edge_i)
        self.num_edges = E
        self.position_2 = capi.as_f64(position_2, (E, 3))
        self.orientation = capi.as_f64(orientation, (self.num_views, 3))
        self.edge_weight = None if edge_weight is None else capi.as_f64(edge_weight, (E,))
        self.fixed_view = int(fixed_view)
        p = capi.PositionProblem()
        p.num_views, p.num_edges = self.num_views, E
        p.edge_i, p.edge_j = capi.ptr(self.edge_i, C.c_uint32), capi.ptr(self.edge_j, C.c_uint32)
        p.position_2, p.orientation, p.edge_weight = capi.ptr(self.position_2), capi.ptr(self.orientation), capi.ptr(self.edge_weight)
        p.fixed_view = self.fixed_view
        p.error_type = int(error_type)
        self.c = p

    def as_rotation_solver_problem(self):
        """capi.ProblemArrays with error type POSITION_BASELINE over the same data (what gsfm_pa_as_ra_problem builds):
        solver.eval_edges / assemble / cost / spmv / pcg / Solver take it, their `omega` arguments carry positions."""
        return capi.ProblemArrays(self.num_views, self.edge_i, self.edge_j, self.position_2, edge_weight=self.edge_weight,
                                  error_type=capi.POSITION_BASELINE, orientation=self.orientation, fixed_view=self.fixed_view)


def default_options():
    """gsfm_pa_default_options: the Ceres defaults, 400 iterations, HuberLoss(0.1)."""
    o = capi.Options()
    capi.lib().gsfm_pa_default_options(C.byref(o))
    return o


def solve(prob, options, positions0=None, trace_capacity=0):
    """gsfm_pa_solve.  positions0 None: every camera at the origin, as the reference starts.  Returns (positions, summary, trace)."""
    x = np.zeros((prob.num_views, 3)) if positions0 is None else capi.as_f64(np.array(positions0, dtype=np.float64, copy=True), (prob.num_views, 3))
    s = capi.Summary()
    trace = (capi.Iteration * max(1, trace_capacity))()
    if trace_capacity:
        s.trace = trace
        s.trace_capacity = trace_capacity
    capi.check(capi.lib().gsfm_pa_solve(C.byref(prob.c), C.byref(options), capi.ptr(x), C.byref(s)))
    return x, s, [trace[k] for k in range(s.trace_size)]


def eval_edges(prob, loss, positions, device=-1):
    E = prob.num_edges
    x = capi.as_f64(positions, (prob.num_views, 3))
    r, Ji, Jj, rho = np.zeros((E, 3)), np.zeros((E, 3, 3)), np.zeros((E, 3, 3)), np.zeros((E, 3))
    capi.check(capi.lib().gsfm_pa_eval_edges(C.byref(prob.c), C.byref(loss), capi.ptr(x), capi.ptr(r), capi.ptr(Ji), capi.ptr(Jj),
                                             capi.ptr(rho), device))
    return r, Ji, Jj, rho


def cost(prob, loss, positions, device=-1):
    x = capi.as_f64(positions, (prob.num_views, 3))
    c = C.c_double()
    capi.check(capi.lib().gsfm_pa_cost(C.byref(prob.c), C.byref(loss), capi.ptr(x), C.byref(c), device))
    return c.value


def synthetic_position_graph(num_views, num_edges, seed=56, direction_noise_deg=1.0, outlier_fraction=0.1, scene_scale=10.0):
    """A synthetic translation-averaging problem on the generator of viewgraph.synthetic_pose_graph (same edge set: a
    spanning chain + random distinct pairs): ground-truth positions uniform in a cube, ground-truth orientations, per pair the
    unit direction R_i (c_j - c_i)/|c_j - c_i| (camera-1 frame, TwoViewInfo::position_2 convention: R_i^T position_2 points
    from c_i to c_j in the world), perturbed by `direction_noise_deg` and replaced by a uniform random direction for a
    fraction of the pairs.  Returns (PositionProblemArrays-ready dict)."""
    from . import viewgraph as vg
    g = vg.synthetic_pose_graph(num_views, num_edges, seed=seed, noise_deg=0.0, outlier_fraction=0.0, init="gt_perturbed")
    rng = np.random.default_rng(seed + 1)
    c = scene_scale * rng.uniform(-1.0, 1.0, size=(num_views, 3))
    R = vg.so3_exp(g.omega_gt)                                    # [N,3,3] world -> camera
    d = c[g.edge_j] - c[g.edge_i]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    n = np.radians(direction_noise_deg) * rng.normal(size=d.shape)
    d = d + n
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = rng.uniform(size=len(d)) < outlier_fraction
    rnd = rng.normal(size=d.shape)
    rnd /= np.linalg.norm(rnd, axis=1, keepdims=True)
    d[out] = rnd[out]
    p2 = np.einsum("eab,eb->ea", R[g.edge_i], d)                 # camera-1 frame: position_2 = R_i * world direction
    return dict(num_views=num_views, edge_i=g.edge_i, edge_j=g.edge_j, position_2=p2, orientation=g.omega_gt, positions_gt=c,
                is_outlier=out)
