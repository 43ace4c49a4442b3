"""ctypes binding of the CPU oracle (oracle/libra_oracle.so).  TEST INFRASTRUCTURE.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (globalsfmpy_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from globalsfmpy_b200 import _abi as capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libra_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def _declare(lib):
    """argtypes / restypes of oracle/ra_oracle.h (the structs are the C ABI's own: include/gsfm_ra.h)."""
    pp, lp, op, sp = C.POINTER(capi.Problem), C.POINTER(capi.Loss), C.POINTER(capi.Options), C.POINTER(capi.Summary)
    dp, u32p, u8p = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    cb = C.CFUNCTYPE(None, C.c_double, dp, C.c_void_p)
    lib.ra_oracle_loss.argtypes = [lp, C.c_double, dp]
    lib.ra_oracle_loss.restype = None
    lib.ra_oracle_gamma_table.argtypes = [C.c_int, C.c_int]
    lib.ra_oracle_gamma_table.restype = C.c_double
    lib.ra_oracle_angle_axis_to_matrix.argtypes = [dp, dp]
    lib.ra_oracle_matrix_to_angle_axis.argtypes = [dp, dp]
    lib.ra_oracle_whiten.argtypes = [C.c_int, dp, C.c_double, dp]
    lib.ra_oracle_edge.argtypes = [dp] * 7
    lib.ra_oracle_eval_edges.argtypes = [pp, lp, dp, dp, dp, dp, dp, C.c_int]
    lib.ra_oracle_assemble.argtypes = [pp, lp, dp, dp, dp, dp, u32p, u32p, dp, C.c_int]
    lib.ra_oracle_cost.argtypes = [pp, lp, dp, dp, C.c_int]
    lib.ra_oracle_solve.argtypes = [pp, op, dp, sp, cb, C.c_void_p]
    lib.ra_oracle_solve_sigma_consensus.argtypes = [pp, op, C.c_int32, C.c_double, dp, sp, dp]
    lib.ra_oracle_filter_view_pairs.argtypes = [pp, dp, C.c_double, u8p, dp]
    lib.loss_cb_type = cb
    return lib


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def _p(a):
    return capi.ptr(a)


def loss(loss_struct, s):
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))
    out = np.zeros((len(s), 3))
    L = lib()
    for k in range(len(s)):
        L.ra_oracle_loss(C.byref(loss_struct), float(s[k]), _p(out[k]))
    return out


def gamma_table(nu, idx):
    L = lib()
    return np.array([L.ra_oracle_gamma_table(int(nu), int(i)) for i in np.atleast_1d(idx)])


def angle_axis_to_matrix(w):
    w = capi.as_f64(w, (3,))
    R = np.zeros((3, 3))
    lib().ra_oracle_angle_axis_to_matrix(_p(w), _p(R))
    return R


def matrix_to_angle_axis(R):
    R = capi.as_f64(R, (3, 3))
    w = np.zeros(3)
    lib().ra_oracle_matrix_to_angle_axis(_p(R), _p(w))
    return w


def whiten(error_type, cov6, edge_weight=1.0):
    c = capi.as_f64(cov6 if cov6 is not None else np.zeros(6), (6,))
    U = np.zeros((3, 3))
    lib().ra_oracle_whiten(int(error_type), _p(c), float(edge_weight), _p(U))
    return U


def edge(wi, wj, wij, U):
    wi, wj, wij, U = (capi.as_f64(a) for a in (wi, wj, wij, U))
    r, Ji, Jj = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 3))
    lib().ra_oracle_edge(_p(wi), _p(wj), _p(wij), _p(U), _p(r), _p(Ji), _p(Jj))
    return r, Ji, Jj


def eval_edges(prob, loss_struct, omega, num_threads=0):
    E = prob.num_edges
    omega = capi.as_f64(omega, (prob.num_views, 3))
    d = capi.residual_dim(prob.c.error_type)
    r, Ji, Jj, rho = np.zeros((E, d)), np.zeros((E, d, 3)), np.zeros((E, d, 3)), np.zeros((E, 3))
    rc = lib().ra_oracle_eval_edges(C.byref(prob.c), C.byref(loss_struct), _p(omega), _p(r), _p(Ji), _p(Jj), _p(rho),
                                    num_threads)
    assert rc == 0, rc
    return r, Ji, Jj, rho


def assemble(prob, loss_struct, omega, num_threads=0):
    N, E = prob.num_views, prob.num_edges
    omega = capi.as_f64(omega, (N, 3))
    cost = C.c_double()
    g, hd = np.zeros((N, 3)), np.zeros((N, 3, 3))
    rowptr, col, val = np.zeros(N + 1, np.uint32), np.zeros(2 * E, np.uint32), np.zeros((2 * E, 3, 3))
    rc = lib().ra_oracle_assemble(C.byref(prob.c), C.byref(loss_struct), _p(omega), C.byref(cost), _p(g), _p(hd),
                                  capi.ptr(rowptr, C.c_uint32), capi.ptr(col, C.c_uint32), _p(val), num_threads)
    assert rc == 0, rc
    return cost.value, g, hd, rowptr, col, val


def cost(prob, loss_struct, omega, num_threads=0):
    omega = capi.as_f64(omega, (prob.num_views, 3))
    c = C.c_double()
    rc = lib().ra_oracle_cost(C.byref(prob.c), C.byref(loss_struct), _p(omega), C.byref(c), num_threads)
    assert rc == 0, rc
    return c.value


def solve(prob, options, omega0, trace_capacity=0, loss_callback=None):
    """Returns (omega, summary, trace list).  loss_callback(s) -> (rho, rho', rho'') is called
    once per edge per evaluation through a C callback, as the reference calls its Python loss."""
    omega = capi.as_f64(np.array(omega0, dtype=np.float64, copy=True), (prob.num_views, 3))
    s = capi.Summary()
    trace = (capi.Iteration * max(1, trace_capacity))()
    if trace_capacity:
        s.trace = trace
        s.trace_capacity = trace_capacity
    L = lib()
    if loss_callback is None:
        cb = C.cast(None, L.loss_cb_type)
    else:
        def _cb(sq, out, _ctx):
            r = loss_callback(sq)
            out[0], out[1], out[2] = r[0], r[1], r[2]
        cb = L.loss_cb_type(_cb)
    rc = L.ra_oracle_solve(C.byref(prob.c), C.byref(options), _p(omega), C.byref(s), cb, None)
    assert rc == 0, rc
    return omega, s, [trace[k] for k in range(s.trace_size)]


def solve_sigma_consensus(prob, options, omega0, iters_num, sigma_max):
    omega = capi.as_f64(np.array(omega0, dtype=np.float64, copy=True), (prob.num_views, 3))
    s = capi.Summary()
    w = np.zeros(prob.num_edges)
    rc = lib().ra_oracle_solve_sigma_consensus(C.byref(prob.c), C.byref(options), int(iters_num), float(sigma_max), _p(omega),
                                               C.byref(s), _p(w))
    assert rc == 0, rc
    return omega, s, w


def filter_view_pairs(prob, omega, max_degrees):
    omega = capi.as_f64(omega, (prob.num_views, 3))
    keep = np.zeros(prob.num_edges, np.uint8)
    ang = np.zeros(prob.num_edges)
    rc = lib().ra_oracle_filter_view_pairs(C.byref(prob.c), _p(omega), float(max_degrees),
                                           capi.ptr(keep, C.c_uint8), _p(ang))
    assert rc == 0, rc
    return keep.astype(bool), ang


def position_problem(pp):
    """The gsfm_ra_problem (error type POSITION_BASELINE) of a globalsfmpy_b200.positions.PositionProblemArrays: every oracle
    entry point above takes it, with positions in place of omega (translation averaging)."""
    return pp.as_rotation_solver_problem()
