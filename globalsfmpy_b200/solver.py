"""Python host API over the C ABI (include/gsfm_ra.h).  Everything here calls libgsfm_ra.so;
nothing computes on the CPU, and nothing here imports the oracle."""
import ctypes as C
import os

import numpy as np

from . import _capi as capi


def make_problem(graph, error_type, edge_weight=None):
    """graph: viewgraph.PoseGraph (dense-renumbered views)."""
    cov = graph.cov6 if error_type in (capi.ANGLE_AXIS_COVARIANCE, capi.ANGLE_AXIS_COV_INLIERS,
                                       capi.ANGLE_AXIS_COVTRACE, capi.ANGLE_AXIS_COVNORM) else None
    w = edge_weight if edge_weight is not None else graph.edge_weight
    return capi.ProblemArrays(graph.num_views, graph.edge_i, graph.edge_j, graph.omega_ij, cov6=cov,
                              edge_weight=w, error_type=error_type)


def _omega(prob, omega):
    return capi.as_f64(omega, (prob.num_views, 3))


def eval_loss(loss, s, device=-1):
    s = capi.as_f64(np.atleast_1d(s))
    out = np.zeros((len(s), 3))
    capi.check(capi.lib().gsfm_ra_eval_loss(C.byref(loss), capi.ptr(s), len(s), capi.ptr(out), device))
    return out


def whiten(prob, device=-1):
    U = np.zeros((prob.num_edges, 3, 3))
    capi.check(capi.lib().gsfm_ra_whiten(C.byref(prob.c), capi.ptr(U), device))
    return U


def eval_edges(prob, loss, omega, device=-1):
    E = prob.num_edges
    omega = _omega(prob, omega)
    d = capi.residual_dim(prob.c.error_type)
    r, Ji, Jj, rho = np.zeros((E, d)), np.zeros((E, d, 3)), np.zeros((E, d, 3)), np.zeros((E, 3))
    capi.check(capi.lib().gsfm_ra_eval_edges(C.byref(prob.c), C.byref(loss), capi.ptr(omega), capi.ptr(r), capi.ptr(Ji),
                                             capi.ptr(Jj), capi.ptr(rho), device))
    return r, Ji, Jj, rho


def assemble(prob, loss, omega, device=-1):
    N, E = prob.num_views, prob.num_edges
    omega = _omega(prob, omega)
    cost = C.c_double()
    g, hd = np.zeros((N, 3)), np.zeros((N, 3, 3))
    rowptr, col, val = np.zeros(N + 1, np.uint32), np.zeros(2 * E, np.uint32), np.zeros((2 * E, 3, 3))
    capi.check(capi.lib().gsfm_ra_assemble(C.byref(prob.c), C.byref(loss), capi.ptr(omega), C.byref(cost), capi.ptr(g),
                                           capi.ptr(hd), capi.ptr(rowptr, C.c_uint32), capi.ptr(col, C.c_uint32),
                                           capi.ptr(val), device))
    return cost.value, g, hd, rowptr, col, val


def cost(prob, loss, omega, device=-1):
    omega = _omega(prob, omega)
    c = C.c_double()
    capi.check(capi.lib().gsfm_ra_cost(C.byref(prob.c), C.byref(loss), capi.ptr(omega), C.byref(c), device))
    return c.value


def spmv(prob, loss, omega, x, damping=None, device=-1):
    omega = _omega(prob, omega)
    x = capi.as_f64(x, (prob.num_views, 3))
    d = None if damping is None else capi.as_f64(damping, (prob.num_views, 3))
    y = np.zeros_like(x)
    capi.check(capi.lib().gsfm_ra_spmv(C.byref(prob.c), C.byref(loss), capi.ptr(omega), capi.ptr(d), capi.ptr(x),
                                       capi.ptr(y), device))
    return y


def pcg(prob, loss, omega, b, damping=None, rtol=1e-10, max_iterations=500, device=-1):
    omega = _omega(prob, omega)
    b = capi.as_f64(b, (prob.num_views, 3))
    d = None if damping is None else capi.as_f64(damping, (prob.num_views, 3))
    x = np.zeros_like(b)
    it, res = C.c_int32(), C.c_double()
    capi.check(capi.lib().gsfm_ra_pcg(C.byref(prob.c), C.byref(loss), capi.ptr(omega), capi.ptr(d), capi.ptr(b), rtol,
                                      max_iterations, capi.ptr(x), C.byref(it), C.byref(res), device))
    return x, it.value, res.value


def filter_view_pairs(prob, omega, max_degrees, device=-1):
    omega = _omega(prob, omega)
    keep = np.zeros(prob.num_edges, np.uint8)
    ang = np.zeros(prob.num_edges)
    capi.check(capi.lib().gsfm_ra_filter_view_pairs(C.byref(prob.c), capi.ptr(omega), float(max_degrees),
                                                    capi.ptr(keep, C.c_uint8), capi.ptr(ang), device))
    return keep.astype(bool), ang


def filter_initial_view_graph(num_views, edge_i, edge_j, num_verified_matches, min_num_two_view_inliers=30, device=-1):
    """FilterInitialViewGraph + largest connected component on the device (dense view indices).
    Reference: src/GSfM_global_reconstruction_estimator.cpp:369-390.  Returns (edge_keep, view_keep) boolean masks."""
    ei = np.ascontiguousarray(edge_i, dtype=np.uint32)
    ej = np.ascontiguousarray(edge_j, dtype=np.uint32)
    m = np.ascontiguousarray(num_verified_matches, dtype=np.int32)
    ek = np.zeros(len(ei), np.uint8)
    vk = np.zeros(int(num_views), np.uint8)
    capi.check(capi.lib().gsfm_ra_filter_initial_view_graph(int(num_views), len(ei), capi.ptr(ei, C.c_uint32), capi.ptr(ej, C.c_uint32),
                                                            capi.ptr(m, C.c_int32), int(min_num_two_view_inliers),
                                                            capi.ptr(ek, C.c_uint8), capi.ptr(vk, C.c_uint8), device))
    return ek.astype(bool), vk.astype(bool)


def init_orientations_mst(num_views, edge_i, edge_j, omega_ij, weights, root=None, device=-1):
    """OrientationsFromMaximumSpanningTree on the device (Boruvka + level-synchronous propagation).
    Reference: T/sfm/view_graph/orientations_from_maximum_spanning_tree.cc:109-178.
    Returns (omega [N,3] with NaN for unreachable views, tree-edge mask, Boruvka rounds)."""
    ei = np.ascontiguousarray(edge_i, dtype=np.uint32)
    ej = np.ascontiguousarray(edge_j, dtype=np.uint32)
    w = np.ascontiguousarray(weights, dtype=np.int32)
    wij = np.ascontiguousarray(omega_ij, dtype=np.float64).reshape(-1, 3)
    om = np.zeros((int(num_views), 3))
    tree = np.zeros(len(ei), np.uint8)
    rounds = C.c_int32(0)
    capi.check(capi.lib().gsfm_ra_init_orientations_mst(int(num_views), len(ei), capi.ptr(ei, C.c_uint32), capi.ptr(ej, C.c_uint32),
                                                        capi.ptr(wij), capi.ptr(w, C.c_int32), -1 if root is None else int(root),
                                                        capi.ptr(om), capi.ptr(tree, C.c_uint8), C.byref(rounds), device))
    return om, tree.astype(bool), rounds.value


def _take(ptr_obj, n, dtype):
    """Copy n elements out of a malloc'ed array returned by the library and release it."""
    if n == 0 or not ptr_obj:
        out = np.zeros(0, dtype=dtype)
    else:
        out = np.ctypeslib.as_array(ptr_obj, shape=(n,)).astype(dtype, copy=True)
    if ptr_obj:
        capi.lib().gsfm_ra_free(C.cast(ptr_obj, C.c_void_p))
    return out


def read_covariance_rot(path):
    """covariance_rot.txt through the native reader (reference src/uncertainty.cpp:200-229).
    Returns ids [C,2] int64, cov6 [C,6] (C00 C11 C22 C01 C02 C12), rot [C,3]."""
    n = C.c_uint64(0)
    a, b = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    c6, r3 = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    capi.check(capi.lib().gsfm_ra_read_covariance_rot(os.fsencode(path), C.byref(n), C.byref(a), C.byref(b), C.byref(c6), C.byref(r3)))
    k = n.value
    ids = np.stack([_take(a, k, np.int64), _take(b, k, np.int64)], axis=1) if k else np.zeros((0, 2), np.int64)
    return ids, _take(c6, 6 * k, np.float64).reshape(k, 6), _take(r3, 3 * k, np.float64).reshape(k, 3)


def write_covariance_rot(path, ids, cov6, rot):
    """store_covariance_rot's file layout (reference src/uncertainty.cpp:164-198) through the native writer."""
    ids = np.asarray(ids)
    a = np.ascontiguousarray(ids[:, 0], dtype=np.uint32)
    b = np.ascontiguousarray(ids[:, 1], dtype=np.uint32)
    c6 = capi.as_f64(cov6, (len(a), 6))
    r3 = capi.as_f64(rot, (len(a), 3))
    capi.check(capi.lib().gsfm_ra_write_covariance_rot(os.fsencode(path), len(a), capi.ptr(a, C.c_uint32), capi.ptr(b, C.c_uint32),
                                                       capi.ptr(c6), capi.ptr(r3)))


def read_1dsfm(dataset_directory, with_matches=True):
    """A 1DSfM dataset directory through the native reader (T/io/read_1dsfm.cc:93-412).  Returns a dict:
    num_listed_views, view_ids [V], focal_length_priors [V], pairs [P,2] (as listed in EGs.txt), rotation_2 [P,3],
    position_2 [P,3], num_verified_matches [P] (None unless with_matches)."""
    listed, nv, npairs = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    ids, a, b = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    focal, rot, pos = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    m = C.POINTER(C.c_int32)()
    capi.check(capi.lib().gsfm_ra_read_1dsfm(os.fsencode(dataset_directory), C.byref(listed), C.byref(nv), C.byref(ids), C.byref(focal),
                                             C.byref(npairs), C.byref(a), C.byref(b), C.byref(rot), C.byref(pos),
                                             C.byref(m) if with_matches else None))
    V, P = nv.value, npairs.value
    pairs = np.stack([_take(a, P, np.int64), _take(b, P, np.int64)], axis=1) if P else np.zeros((0, 2), np.int64)
    return dict(num_listed_views=listed.value, view_ids=_take(ids, V, np.int64), focal_length_priors=_take(focal, V, np.float64), pairs=pairs,
                rotation_2=_take(rot, 3 * P, np.float64).reshape(P, 3), position_2=_take(pos, 3 * P, np.float64).reshape(P, 3),
                num_verified_matches=_take(m, P, np.int64) if with_matches else None)


def measure_stream(nbytes, repeats=20, device=-1):
    """GB/s of the K2 record stream alone over `nbytes` of device memory (gsfm_ra_measure_stream)."""
    out = C.c_double(0.0)
    capi.check(capi.lib().gsfm_ra_measure_stream(int(nbytes), int(repeats), int(device), C.byref(out)))
    return out.value


def _summary(trace_capacity):
    s = capi.Summary()
    trace = (capi.Iteration * max(1, trace_capacity))()
    if trace_capacity:
        s.trace = trace
        s.trace_capacity = trace_capacity
    return s, trace


def solve(prob, options, omega0, trace_capacity=0):
    """One-shot gsfm_ra_solve: host buffers in, host buffers out. Returns (omega, summary, trace)."""
    omega = capi.as_f64(np.array(omega0, dtype=np.float64, copy=True), (prob.num_views, 3))
    s, trace = _summary(trace_capacity)
    capi.check(capi.lib().gsfm_ra_solve(C.byref(prob.c), C.byref(options), capi.ptr(omega), C.byref(s)))
    return omega, s, [trace[k] for k in range(s.trace_size)]


def solve_sigma_consensus(prob, options, omega0, iters_num, sigma_max):
    """gsfm_ra_solve_sigma_consensus (EstimateRotationsWithSigmaConsensus). Returns (omega, summary)."""
    omega = capi.as_f64(np.array(omega0, dtype=np.float64, copy=True), (prob.num_views, 3))
    s, _ = _summary(0)
    capi.check(capi.lib().gsfm_ra_solve_sigma_consensus(C.byref(prob.c), C.byref(options), int(iters_num), float(sigma_max),
                                                        capi.ptr(omega), C.byref(s)))
    return omega, s


class Solver:
    """Resident solver handle (problem stays in HBM across iterations)."""

    def __init__(self, prob, options, rank=0, world_size=1):
        self.prob = prob
        self.options = options
        self._h = C.c_void_p()
        if world_size == 1:
            capi.check(capi.lib().gsfm_ra_solver_create(C.byref(prob.c), C.byref(options), C.byref(self._h)))
        else:
            capi.check(capi.lib().gsfm_ra_solver_create_sharded(C.byref(prob.c), C.byref(options), rank, world_size,
                                                                C.byref(self._h)))

    def connect(self, dist, fused=None):
        """Join the ranks of a sharded solver: rank 0 makes the NCCL id, torch.distributed (`dist`, already
        initialised by the host framework) broadcasts it, every rank joins.  Collective."""
        import torch
        ident = np.zeros(capi.COMM_ID_BYTES, np.uint8)
        if dist.get_rank() == 0:
            capi.check(capi.lib().gsfm_ra_comm_unique_id(capi.ptr(ident, C.c_uint8)))
        t = torch.from_numpy(ident)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        ident = np.ascontiguousarray(t.cpu().numpy())
        capi.check(capi.lib().gsfm_ra_solver_comm_init(self._h, capi.ptr(ident, C.c_uint8)))
        if fused is None:
            import os
            fused = os.environ.get("GSFM_RA_EXCHANGE", "fused") != "nccl"
        if fused:
            # fused exchange: all-gather the CUDA IPC handles of the per-rank exchange blocks
            world = dist.get_world_size()
            mine = np.zeros(capi.IPC_HANDLE_BYTES, np.uint8)
            capi.check(capi.lib().gsfm_ra_solver_ipc_export(self._h, capi.ptr(mine, C.c_uint8)))
            tm = torch.from_numpy(mine)
            if dist.get_backend() == "nccl":
                tm = tm.cuda()
            parts = [torch.empty_like(tm) for _ in range(world)]
            dist.all_gather(parts, tm)
            allh = np.ascontiguousarray(torch.stack(parts).cpu().numpy())
            capi.check(capi.lib().gsfm_ra_solver_ipc_import(self._h, capi.ptr(allh, C.c_uint8)))
            dist.barrier()

    def edge_range(self):
        a, b = C.c_uint64(), C.c_uint64()
        capi.check(capi.lib().gsfm_ra_solver_edge_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if self._h:
            capi.lib().gsfm_ra_solver_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_rotations(self, omega):
        omega = _omega(self.prob, omega)
        capi.check(capi.lib().gsfm_ra_solver_set_rotations(self._h, capi.ptr(omega)))

    def get_rotations(self):
        omega = np.zeros((self.prob.num_views, 3))
        capi.check(capi.lib().gsfm_ra_solver_get_rotations(self._h, capi.ptr(omega)))
        return omega

    def reset(self):
        capi.check(capi.lib().gsfm_ra_solver_reset(self._h))

    @property
    def cuda_stream(self):
        """Raw cudaStream_t of the solver (wrap with torch.cuda.ExternalStream to record events on it)."""
        return capi.lib().gsfm_ra_solver_cuda_stream(self._h)

    def time_kernels(self, repeats=20):
        """Average ms per launch: dict(k1, k1c, spmv, pcg_iteration)."""
        out = np.zeros(4)
        capi.check(capi.lib().gsfm_ra_solver_time_kernels(self._h, int(repeats), capi.ptr(out)))
        return dict(k1=out[0], k1c=out[1], spmv=out[2], pcg_iteration=out[3])

    def info(self):
        """What the solver resolved at build time (gsfm_ra_solver_info)."""
        out = (C.c_int32 * 8)()
        capi.check(capi.lib().gsfm_ra_solver_info(self._h, out))
        return dict(linear_solver=out[0], stored_bytes_per_half_edge=out[1], pcg_grid=out[2], pcg_block=out[3], l2_keep8=out[4],
                    world=out[5], cuda_graphs=bool(out[6]))

    def iterate(self, num_iterations, trace_capacity=0):
        s, trace = _summary(trace_capacity)
        capi.check(capi.lib().gsfm_ra_solver_iterate(self._h, int(num_iterations), C.byref(s)))
        return s, [trace[k] for k in range(s.trace_size)]
