"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and through size-independent properties.
Tolerances (fp64 end to end): per-edge quantities 1e-9 relative to the row's magnitude (the oracle
differentiates with jets through matrix->quaternion->angle-axis, the kernel uses closed-form SO(3)
Jacobians; both are ~1e-13 accurate away from the theta=pi singularity), converged rotations
<= 1e-6 rad mean for smooth losses (north_star bar: 1e-4 rad)."""
import os

import numpy as np
import pytest

from globalsfmpy_b200 import _capi as capi, solver, viewgraph as vg
from oracle import ra_oracle as orc
from common import GOLDEN_LOSSES, assert_close, rel_err_rows

pytestmark = pytest.mark.gpu

CAUCHY = capi.Loss.make(capi.LOSS_CAUCHY, 0.05)
MAGSAC = capi.Loss.make(capi.LOSS_MAGSAC3, 0.02)


def _dense(N, hd, rowptr, col, val):
    H = np.zeros((3 * N, 3 * N))
    for a in range(N):
        H[3 * a:3 * a + 3, 3 * a:3 * a + 3] = hd[a]
        for s in range(rowptr[a], rowptr[a + 1]):
            H[3 * a:3 * a + 3, 3 * col[s]:3 * col[s] + 3] = val[s]
    return H


def test_device_present():
    assert capi.lib().gsfm_ra_device_count() >= 1


def test_losses_match_reference_and_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    s = z["s"]
    for name, L in GOLDEN_LOSSES.items():
        got = solver.eval_loss(L, s)
        ref = z[name]
        tol = (1e-9 if "inv" in name else 1e-12) * np.abs(ref) + 1e-11
        bad = np.abs(got - ref) > tol
        assert not bad.any(), (name, s[bad.any(axis=1)][:5], got[bad][:5], ref[bad][:5])
        o = orc.loss(L, s)
        assert not (np.abs(got - o) > tol).any(), name


def test_whitening(madrid):
    prob = solver.make_problem(madrid, capi.ANGLE_AXIS_COVARIANCE)
    U = solver.whiten(prob)
    for k in range(0, madrid.num_edges, 211):
        ref = orc.whiten(capi.ANGLE_AXIS_COVARIANCE, madrid.cov6[k])
        assert np.abs(U[k] - ref).max() <= 1e-12 * np.abs(ref).max()
    for t in (capi.ANGLE_AXIS_COVTRACE, capi.ANGLE_AXIS_COVNORM, capi.ANGLE_AXIS_COV_INLIERS):
        w = np.linspace(0.3, 2.0, madrid.num_edges)
        p2 = solver.make_problem(madrid, t, edge_weight=w)
        U = solver.whiten(p2)
        for k in range(0, madrid.num_edges, 997):
            ref = orc.whiten(t, madrid.cov6[k], w[k])
            assert np.array_equal(U[k], ref)


def test_known_answer_residuals():
    """pairwise_rotation_error_test.cc:87-139 through the CUDA kernel."""
    from scipy.spatial.transform import Rotation as R
    rz = lambda d: R.from_euler("z", d, degrees=True)
    xyz = lambda a, b, c: R.from_euler("x", a, degrees=True) * R.from_euler("y", b, degrees=True) * R.from_euler("z", c, degrees=True)
    I = R.identity()
    cases = [(rz(1), 1.0, I, rz(2)), (xyz(5.9, 1.8, 7.6), 1.0, I, xyz(5.3, 1.2, 8.1)), (rz(-179), 1.0, I, rz(179)),
             (rz(-179), 1.0, rz(179), I), (xyz(5.9, 1.8, 7.6), 2.0, I, xyz(5.3, 1.2, 8.1))]
    for R12, w, R1, R2 in cases:
        gt = w * (R2 * R1.inv() * R12.inv()).as_rotvec()
        prob = capi.ProblemArrays(2, [0], [1], [R12.as_rotvec()], edge_weight=[w])
        r, _, _, _ = solver.eval_edges(prob, capi.Loss.make(capi.LOSS_TRIVIAL), np.stack([R1.as_rotvec(), R2.as_rotvec()]))
        assert np.abs(r[0] - gt).max() < 1e-12


@pytest.mark.parametrize("etype", [capi.ANGLE_AXIS, capi.ANGLE_AXIS_COVARIANCE, capi.ANGLE_AXIS_COVTRACE])
def test_eval_edges_random(etype):
    g = vg.synthetic_pose_graph(200, 3000, seed=11, noise_deg=2.0, outlier_fraction=0.2, covariance=True)
    rng = np.random.default_rng(0)
    omega = g.omega_init + 0.05 * rng.normal(size=g.omega_init.shape)
    omega[3] *= 1e-9  # near-identity view
    omega[4] = 0.0
    omega[5] *= 4.0   # |omega| beyond pi is legal (no manifold wrap in the reference)
    prob = solver.make_problem(g, etype)
    r, Ji, Jj, rho = solver.eval_edges(prob, MAGSAC, omega)
    r0, Ji0, Jj0, rho0 = orc.eval_edges(prob, MAGSAC, omega)
    ang = np.linalg.norm(np.linalg.solve(orc.whiten(etype, g.cov6[0]), np.eye(3)), axis=0)  # noqa: F841 (sanity only)
    assert rel_err_rows(r, r0) < 1e-9
    # autodiff through atan2/sqrt is ill-conditioned next to theta = pi; exclude residual angles within 1e-3 of pi
    ok = np.abs(np.linalg.norm(vg.so3_log(vg.so3_exp(omega[g.edge_j]) @ np.transpose(vg.so3_exp(omega[g.edge_i]), (0, 2, 1))
                                          @ np.transpose(vg.so3_exp(g.omega_ij), (0, 2, 1))), axis=1) - np.pi) > 1e-3
    assert ok.mean() > 0.99
    assert rel_err_rows(Ji[ok], Ji0[ok]) < 1e-9
    assert rel_err_rows(Jj[ok], Jj0[ok]) < 1e-9
    assert np.allclose(rho, rho0, rtol=1e-9, atol=1e-9)


def test_eval_edges_madrid(madrid):
    prob = solver.make_problem(madrid, capi.ANGLE_AXIS_COVARIANCE)
    r, Ji, Jj, rho = solver.eval_edges(prob, MAGSAC, madrid.omega_init)
    r0, Ji0, Jj0, rho0 = orc.eval_edges(prob, MAGSAC, madrid.omega_init)
    assert rel_err_rows(r, r0) < 1e-9
    ok = np.abs(np.linalg.norm(np.linalg.solve(solver.whiten(prob), r0[..., None])[..., 0], axis=1) - np.pi) > 1e-3
    assert rel_err_rows(Ji[ok], Ji0[ok]) < 1e-8 and rel_err_rows(Jj[ok], Jj0[ok]) < 1e-8
    # LUT quantisation: an edge whose s sits within rounding of a .5 tie may land one step away
    close = np.isclose(rho, rho0, rtol=1e-9, atol=1e-9).all(axis=1)
    assert close.mean() > 0.9999


@pytest.mark.parametrize("loss_name", ["cauchy_0.05", "magsac3_0.02", "softlone_0.1", "huber_0.1", "magsac9_0.3", "trivial", "tukey_0.4"])
@pytest.mark.parametrize("etype", [capi.ANGLE_AXIS, capi.ANGLE_AXIS_COVARIANCE])
def test_assemble_matches_oracle(loss_name, etype):
    L = GOLDEN_LOSSES[loss_name]
    g = vg.synthetic_pose_graph(300, 5000, seed=3, noise_deg=1.5, outlier_fraction=0.1, covariance=True)
    prob = solver.make_problem(g, etype)
    c, gr, hd, rp, col, val = solver.assemble(prob, L, g.omega_init)
    c0, gr0, hd0, rp0, col0, val0 = orc.assemble(prob, L, g.omega_init)
    assert np.array_equal(rp, rp0) and np.array_equal(col, col0)
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert_close(gr, gr0, 1e-10, "gradient")
    assert_close(hd, hd0, 1e-10, "diagonal blocks")
    assert_close(val, val0, 1e-10, "off-diagonal blocks")


def test_assemble_madrid(madrid):
    prob = solver.make_problem(madrid, capi.ANGLE_AXIS_COVARIANCE)
    c, gr, hd, rp, col, val = solver.assemble(prob, MAGSAC, madrid.omega_init)
    c0, gr0, hd0, rp0, col0, val0 = orc.assemble(prob, MAGSAC, madrid.omega_init)
    assert np.array_equal(rp, rp0) and np.array_equal(col, col0)
    assert abs(c - 307496.5626035597) < 1e-6  # golden: oracle value, reproduced independently in SURVEY Appendix E
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert_close(gr, gr0, 1e-9, "gradient")
    assert_close(hd, hd0, 1e-9, "diagonal blocks")
    assert_close(val, val0, 1e-9, "off-diagonal blocks")
    assert abs(solver.cost(prob, MAGSAC, madrid.omega_init) - c0) <= 1e-12 * abs(c0)


def test_edge_direction_and_order_invariance():
    """Edges may come in any order and with i > j (R_j = R_ij R_i is directional)."""
    g = vg.synthetic_pose_graph(40, 200, seed=9, covariance=True)
    rng = np.random.default_rng(1)
    perm = rng.permutation(g.num_edges)
    flip = rng.uniform(size=g.num_edges) < 0.5
    # reversing an edge: R_i = R_ij^T R_j ; whitened residual is NOT invariant, so build the flipped
    # problem for the oracle too and compare like with like
    ei = np.where(flip, g.edge_j, g.edge_i)[perm]
    ej = np.where(flip, g.edge_i, g.edge_j)[perm]
    wij = np.where(flip[:, None], -g.omega_ij, g.omega_ij)[perm]
    prob = capi.ProblemArrays(40, ei, ej, wij, cov6=g.cov6[perm], error_type=capi.ANGLE_AXIS_COVARIANCE)
    c, gr, hd, rp, col, val = solver.assemble(prob, CAUCHY, g.omega_init)
    c0, gr0, hd0, rp0, col0, val0 = orc.assemble(prob, CAUCHY, g.omega_init)
    assert np.array_equal(col, col0)
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert_close(gr, gr0, 1e-10, "gradient")
    assert_close(val, val0, 1e-10, "blocks")
    # unweighted residual norm IS direction invariant: cost equal under flipping with ANGLE_AXIS
    p1 = capi.ProblemArrays(40, g.edge_i, g.edge_j, g.omega_ij)
    p2 = capi.ProblemArrays(40, ei, ej, wij)
    assert abs(solver.cost(p1, CAUCHY, g.omega_init) - solver.cost(p2, CAUCHY, g.omega_init)) < 1e-12 * c0


def test_spmv_and_pcg():
    g = vg.synthetic_pose_graph(150, 2000, seed=21, noise_deg=1.0, outlier_fraction=0.1)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    _, _, hd, rp, col, val = orc.assemble(prob, CAUCHY, g.omega_init)
    H = _dense(150, hd, rp, col, val)
    rng = np.random.default_rng(2)
    x = rng.normal(size=(150, 3))
    damp = np.abs(rng.normal(size=(150, 3))) + 0.5
    y = solver.spmv(prob, CAUCHY, g.omega_init, x, damping=damp)
    ref = (H @ x.ravel() + damp.ravel() * x.ravel()).reshape(150, 3)
    assert_close(y, ref, 1e-11, "SpMV")
    # linearity + symmetry (size-independent properties)
    x2 = rng.normal(size=(150, 3))
    y2 = solver.spmv(prob, CAUCHY, g.omega_init, x2, damping=damp)
    y12 = solver.spmv(prob, CAUCHY, g.omega_init, 2 * x - 3 * x2, damping=damp)
    assert_close(y12, 2 * y - 3 * y2, 1e-11, "linearity")
    assert abs((x2 * y).sum() - (x * y2).sum()) < 1e-10 * abs((x * y2).sum())
    b = rng.normal(size=(150, 3))
    sol, it, res = solver.pcg(prob, CAUCHY, g.omega_init, b, damping=damp, rtol=1e-12, max_iterations=500)
    ref = np.linalg.solve(H + np.diag(damp.ravel()), b.ravel()).reshape(150, 3)
    assert res <= 1e-12 and 0 < it < 500
    assert_close(sol, ref, 1e-9, "PCG solution")


def test_solve_noise_free_recovers_ground_truth():
    g = vg.synthetic_pose_graph(4, 6, seed=56, noise_deg=0.0, outlier_fraction=0.0, init="gt_perturbed")
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    o = capi.default_options_py()
    o.loss = capi.Loss.make(capi.LOSS_SOFTLONE, 0.1)
    om, s, _ = solver.solve(prob, o, g.omega_init)
    _, mx = vg.mean_angular_error(g.omega_gt, om)
    assert np.degrees(mx) < 1e-6


def _tight(loss, **kw):
    o = capi.default_options_py()
    o.loss = loss
    o.function_tolerance = 1e-14
    o.gradient_tolerance = 1e-12
    o.parameter_tolerance = 1e-12
    o.max_num_iterations = 400
    o.pcg_rtol = 1e-12
    o.pcg_max_iterations = 2000
    o.linear_solver = capi.SOLVER_PCG   # the default (AUTO) would pick the dense factorisation for the small test graphs
    for k, v in kw.items():
        setattr(o, k, v)
    return o


@pytest.mark.parametrize("cfg", ["terrace", "medium", "covariance"])
def test_converged_rotations_match_oracle(cfg):
    """north_star: converged rotations within 1e-4 rad mean of the reference solver. Smooth loss, both
    sides converged tightly; the oracle uses an exact (dense Cholesky) linear solve like the reference."""
    if cfg == "terrace":  # BASELINE configs[1] stand-in: ~23 cams, MAGSAC-free smooth variant
        g = vg.synthetic_pose_graph(23, 200, seed=23, noise_deg=0.5, outlier_fraction=0.1, covariance=True)
        prob, L = solver.make_problem(g, capi.ANGLE_AXIS_COVARIANCE), capi.Loss.make(capi.LOSS_CAUCHY, 1.0)
    elif cfg == "medium":
        g = vg.synthetic_pose_graph(300, 6000, seed=56, noise_deg=1.0, outlier_fraction=0.1)
        prob, L = solver.make_problem(g, capi.ANGLE_AXIS), CAUCHY
    else:
        g = vg.synthetic_pose_graph(200, 3000, seed=8, outlier_fraction=0.05, covariance=True)
        prob, L = solver.make_problem(g, capi.ANGLE_AXIS_COVARIANCE), capi.Loss.make(capi.LOSS_SOFTLONE, 1.0)
    og, sg, _ = solver.solve(prob, _tight(L), g.omega_init)
    oo, so, _ = orc.solve(prob, _tight(L, linear_solver=capi.SOLVER_DENSE_CHOLESKY), g.omega_init)
    mean, mx = vg.mean_angular_error(oo, og)
    assert mean <= 1e-6, (mean, mx, sg.final_cost, so.final_cost)
    assert sg.final_cost <= so.final_cost * (1 + 1e-9)


def test_reference_stopping_rule_trajectory():
    """Ceres-default tolerances: the GPU trust-region loop must walk the oracle's trajectory step for step
    (same accept/reject decisions, same costs) while the problem is well conditioned."""
    g = vg.synthetic_pose_graph(120, 1500, seed=31, noise_deg=1.0, outlier_fraction=0.1)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    o = capi.default_options_py()
    o.loss = CAUCHY
    o.pcg_rtol = 1e-13
    o.pcg_max_iterations = 2000
    o.linear_solver = capi.SOLVER_PCG
    og, sg, tg = solver.solve(prob, o, g.omega_init, trace_capacity=256)
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    oo, so, to = orc.solve(prob, o, g.omega_init, trace_capacity=256)
    assert sg.termination == so.termination and sg.num_iterations == so.num_iterations
    for a, b in zip(tg, to):
        assert a.step_is_successful == b.step_is_successful
        assert abs(a.cost - b.cost) <= 1e-7 * abs(b.cost)
    mean, _ = vg.mean_angular_error(oo, og)
    assert mean < 1e-6


def test_madrid_magsac_reference_path(madrid):
    """BASELINE configs[0]: the shipped dataset with the shipped pipeline's settings
    (ANGLE_AXIS_COVARIANCE + MAGSACWeightBasedLoss(0.02), Ceres defaults).  The quantised MAGSAC loss makes the
    trajectory chaotic (1e-10 step noise moves the answer by ~1e-3 rad), so the bar here is
    the first iterations step for step, and a final cost / solution within the oracle's own reproducibility."""
    prob = solver.make_problem(madrid, capi.ANGLE_AXIS_COVARIANCE)
    o = capi.default_options_py()
    o.loss = MAGSAC
    o.pcg_rtol = 1e-13
    o.pcg_max_iterations = 5000
    o.linear_solver = capi.SOLVER_PCG
    og, sg, tg = solver.solve(prob, o, madrid.omega_init, trace_capacity=256)
    assert sg.num_linear_unconverged == 0
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    oo, so, to = orc.solve(prob, o, madrid.omega_init, trace_capacity=256)
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-12 * so.initial_cost
    for a, b in list(zip(tg, to))[:10]:
        assert a.step_is_successful == b.step_is_successful
        assert abs(a.cost - b.cost) <= 1e-6 * abs(b.cost), (a.iteration, a.cost, b.cost)
    assert abs(sg.final_cost - so.final_cost) <= 2e-4 * so.final_cost
    mean, _ = vg.mean_angular_error(oo, og)
    assert mean < 5e-3, mean


def test_dense_cholesky_path_tracks_the_oracle_on_madrid(madrid):
    """GSFM_RA_SOLVER_DENSE_CHOLESKY: exact factorisation on the device, like the reference's SPARSE_NORMAL_CHOLESKY.
    With exact solves on both sides the GPU loop follows the oracle's trajectory on the shipped dataset with the shipped
    settings far longer than PCG can (MAGSAC's staircase loss amplifies 1e-10 step differences)."""
    prob = solver.make_problem(madrid, capi.ANGLE_AXIS_COVARIANCE)
    o = capi.default_options_py()
    o.loss = MAGSAC
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    og, sg, tg = solver.solve(prob, o, madrid.omega_init, trace_capacity=256)
    oo, so, to = orc.solve(prob, o, madrid.omega_init, trace_capacity=256)
    # step for step: same accept / reject decisions and costs to 1e-7 for the first 25 iterations; after that the staircase loss
    # lets 1e-12 differences grow (6e-5 measured) but the two runs stay on the same path: costs to 1e-3 at EVERY common
    # iteration, iteration counts within 3, final cost to 1e-6, solutions within the north_star bar
    rel = [abs(a.cost - b.cost) / abs(b.cost) for a, b in zip(tg, to)]
    together = next((k for k, (a, b) in enumerate(zip(tg, to)) if a.step_is_successful != b.step_is_successful or rel[k] > 1e-7), len(rel))
    print("madrid dense: iterations", len(tg), len(to), "together to 1e-7 for", together, "max rel cost diff", max(rel))
    assert together >= 25, (together, len(tg), len(to))
    assert max(rel) <= 1e-3 and abs(len(tg) - len(to)) <= 3, (max(rel), len(tg), len(to))
    assert abs(sg.final_cost - so.final_cost) <= 1e-6 * so.final_cost
    mean, _ = vg.mean_angular_error(oo, og)
    assert mean <= 1e-4, mean      # the north_star bar, on the shipped dataset with the shipped settings
    # the DEFAULT options take this path by themselves (linear_solver AUTO, 379 views): bit-identical result
    d = capi.default_options_py()
    d.loss = MAGSAC
    od, sd, _ = solver.solve(prob, d, madrid.omega_init)
    assert np.array_equal(od, og) and sd.num_iterations == sg.num_iterations
    # smooth loss, tight tolerances: the north_star bar with margin
    og, sg, _ = solver.solve(prob, _tight(CAUCHY, linear_solver=capi.SOLVER_DENSE_CHOLESKY), madrid.omega_init)
    oo, so, _ = orc.solve(prob, _tight(CAUCHY, linear_solver=capi.SOLVER_DENSE_CHOLESKY), madrid.omega_init)
    mean, mx = vg.mean_angular_error(oo, og)
    assert mean <= 1e-6, (mean, mx, sg.final_cost, so.final_cost)


def test_dense_cholesky_small_and_padded():
    for n_views, n_edges in ((5, 8), (33, 200), (150, 2500)):
        g = vg.synthetic_pose_graph(n_views, n_edges, seed=n_views, noise_deg=1.0, outlier_fraction=0.1)
        prob = solver.make_problem(g, capi.ANGLE_AXIS)
        a, sa, _ = solver.solve(prob, _tight(CAUCHY, linear_solver=capi.SOLVER_DENSE_CHOLESKY), g.omega_init)
        b, sb, _ = orc.solve(prob, _tight(CAUCHY, linear_solver=capi.SOLVER_DENSE_CHOLESKY), g.omega_init)
        mean, _ = vg.mean_angular_error(b, a)
        assert mean < 1e-7 and sa.num_iterations == sb.num_iterations, (n_views, mean, sa.num_iterations, sb.num_iterations)


def test_madrid_cauchy_tight(madrid):
    prob = solver.make_problem(madrid, capi.ANGLE_AXIS_COVARIANCE)
    og, sg, _ = solver.solve(prob, _tight(CAUCHY, pcg_max_iterations=5000), madrid.omega_init)
    oo, so, _ = orc.solve(prob, _tight(CAUCHY, linear_solver=capi.SOLVER_DENSE_CHOLESKY), madrid.omega_init)
    mean, mx = vg.mean_angular_error(oo, og)
    assert mean <= 1e-5, (mean, mx, sg.final_cost, so.final_cost)
    assert sg.final_cost <= so.final_cost * (1 + 1e-9)


def test_quaternion_cosine_type():
    """RotationErrorType.QUATERNION_COSINE, the default of EstimateGlobalRotations (rotation_estimator.cpp:82-198):
    residual 2 w vec(q_ij (q_j q_i^-1)^-1) on quaternion parameter blocks with EigenQuaternionParameterization.  The oracle
    differentiates the functor with 8-wide jets and projects with the parameterisation's Jacobian; the kernel uses the
    closed form in the left tangent frame."""
    g = vg.synthetic_pose_graph(200, 3000, seed=12, noise_deg=2.0, outlier_fraction=0.2)
    prob = capi.ProblemArrays(200, g.edge_i, g.edge_j, g.omega_ij, error_type=capi.QUATERNION_COSINE)
    rng = np.random.default_rng(5)
    omega = g.omega_init + 0.05 * rng.normal(size=g.omega_init.shape)
    L = capi.Loss.make(capi.LOSS_HUBER, 0.05)
    r, Ji, Jj, rho = solver.eval_edges(prob, L, omega)
    r0, Ji0, Jj0, rho0 = orc.eval_edges(prob, L, omega)
    # the sign of a quaternion is a representation choice: the residual is defined up to one global sign per edge
    sgn = np.sign(np.einsum("ek,ek->e", r, r0))
    sgn[sgn == 0] = 1
    assert rel_err_rows(r * sgn[:, None], r0) < 1e-10
    assert rel_err_rows(Ji * sgn[:, None, None], Ji0) < 1e-10 and rel_err_rows(Jj * sgn[:, None, None], Jj0) < 1e-10
    assert np.allclose(rho, rho0, rtol=1e-10, atol=1e-12)
    c, gr, hd, rp, col, val = solver.assemble(prob, L, omega)
    c0, gr0, hd0, rp0, col0, val0 = orc.assemble(prob, L, omega)
    assert abs(c - c0) <= 1e-12 * abs(c0)
    assert_close(gr, gr0, 1e-10, "gradient (local coordinates)")
    assert_close(hd, hd0, 1e-10, "diagonal blocks")
    assert_close(val, val0, 1e-10, "off-diagonal blocks")
    # default Ceres tolerances: same trajectory; tight: same minimiser
    o = capi.default_options_py()
    o.loss = L
    o.pcg_rtol = 1e-13
    o.pcg_max_iterations = 2000
    og, sg, tg = solver.solve(prob, o, g.omega_init, trace_capacity=256)
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    oo, so, to = orc.solve(prob, o, g.omega_init, trace_capacity=256)
    assert sg.termination == so.termination and sg.num_iterations == so.num_iterations
    for a, b in zip(tg, to):
        assert a.step_is_successful == b.step_is_successful and abs(a.cost - b.cost) <= 1e-8 * abs(b.cost)
        assert abs(a.step_norm - b.step_norm) <= 1e-6 * max(b.step_norm, 1e-12)
    assert vg.mean_angular_error(oo, og)[0] < 1e-7
    og, sg, _ = solver.solve(prob, _tight(L), g.omega_init)
    oo, so, _ = orc.solve(prob, _tight(L, linear_solver=capi.SOLVER_DENSE_CHOLESKY), g.omega_init)
    assert vg.mean_angular_error(oo, og)[0] < 1e-7
    assert np.degrees(vg.mean_angular_error(g.omega_gt, og)[1]) < 3.0


@pytest.mark.parametrize("etype", [capi.QUATERNION_NORM, capi.ROTATION_MAT_FNORM])
def test_general_two_block_types(etype):
    """RotationErrorType.QUATERNION_NORM (4 residuals, include/pairwise_rotation_error_quat.hpp:125-150) and ROTATION_MAT_FNORM
    (9 residuals, :169-196) of EstimateRotationsWithCustomizedLoss (rotation_estimator.cpp:82-198).  These residuals are not
    functions of the error rotation alone, so the two Jacobian blocks are unrelated: the kernel assembles a general block pair
    (9-double records).  The oracle differentiates the functors with 8-wide jets and EigenQuaternionParameterization."""
    g = vg.synthetic_pose_graph(150, 2000, seed=21, noise_deg=2.0, outlier_fraction=0.15)
    prob = capi.ProblemArrays(150, g.edge_i, g.edge_j, g.omega_ij, error_type=etype)
    rng = np.random.default_rng(8)
    omega = g.omega_init + 0.05 * rng.normal(size=g.omega_init.shape)
    for L in (capi.Loss.make(capi.LOSS_HUBER, 0.05), capi.Loss.make(capi.LOSS_MAGSAC9, 0.05)):  # MAGSAC9: rho'' > 0 -> Triggs term
        r, Ji, Jj, rho = solver.eval_edges(prob, L, omega)
        r0, Ji0, Jj0, rho0 = orc.eval_edges(prob, L, omega)
        assert r.shape == (2000, capi.residual_dim(etype))
        assert rel_err_rows(r, r0) < 1e-10
        assert rel_err_rows(Ji, Ji0) < 1e-10 and rel_err_rows(Jj, Jj0) < 1e-10
        assert np.allclose(rho, rho0, rtol=1e-9, atol=1e-12)
        c, gr, hd, rp, col, val = solver.assemble(prob, L, omega)
        c0, gr0, hd0, rp0, col0, val0 = orc.assemble(prob, L, omega)
        assert abs(c - c0) <= 1e-12 * abs(c0)
        assert_close(gr, gr0, 1e-9, "gradient (local coordinates)")
        assert_close(hd, hd0, 1e-9, "diagonal blocks")
        assert_close(val, val0, 1e-9, "off-diagonal blocks")
    L = capi.Loss.make(capi.LOSS_HUBER, 0.05)
    o = capi.default_options_py()
    o.loss = L
    o.pcg_rtol = 1e-13
    o.pcg_max_iterations = 2000
    og, sg, tg = solver.solve(prob, o, g.omega_init, trace_capacity=256)
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    oo, so, to = orc.solve(prob, o, g.omega_init, trace_capacity=256)
    assert sg.termination == so.termination and sg.num_iterations == so.num_iterations
    for a, b in zip(tg, to):
        assert a.step_is_successful == b.step_is_successful and abs(a.cost - b.cost) <= 1e-8 * abs(b.cost)
    assert vg.mean_angular_error(oo, og)[0] < 1e-6
    # the dense Cholesky path reads the 9-double records too
    og2, _, _ = solver.solve(prob, o, g.omega_init)
    assert vg.mean_angular_error(og2, og)[0] < 1e-7
    assert np.degrees(vg.mean_angular_error(g.omega_gt, og)[1]) < 3.0


def test_sigma_consensus_matches_oracle():
    """EstimateRotationsWithSigmaConsensus (rotation_estimator.cpp:314-457): outer re-weighting loop."""
    g = vg.synthetic_pose_graph(120, 1500, seed=77, noise_deg=1.0, outlier_fraction=0.15)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    o = capi.default_options_py()
    o.loss = capi.Loss.make(capi.LOSS_HUBER, 0.1)
    o.pcg_rtol = 1e-13
    o.pcg_max_iterations = 2000
    og, sg = solver.solve_sigma_consensus(prob, o, g.omega_init, 6, 0.05)
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    oo, so, w = orc.solve_sigma_consensus(prob, o, g.omega_init, 6, 0.05)
    assert sg.outer_iterations == so.outer_iterations == 6
    assert (w < 0).any() and w.max() > 15.0   # beyond k*sigma the table weight goes slightly negative; inliers ~ 2 C3 / sigma = 16.1
    assert abs(sg.last_weight_change - so.last_weight_change) <= 1e-6 * max(so.last_weight_change, 1e-12) + 1e-12
    assert abs(sg.final_cost - so.final_cost) <= 1e-6 * so.final_cost
    mean, _ = vg.mean_angular_error(oo, og)
    assert mean < 1e-6, mean
    _, mx = vg.mean_angular_error(g.omega_gt, og)
    assert np.degrees(mx) < 3.0


def test_resident_solver_stepwise_equals_one_shot():
    g = vg.synthetic_pose_graph(100, 1200, seed=13, noise_deg=1.0, outlier_fraction=0.1)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    o = capi.default_options_py()
    o.loss = CAUCHY
    o.linear_solver = capi.SOLVER_PCG
    one, s1, _ = solver.solve(prob, o, g.omega_init)
    S = solver.Solver(prob, o)
    S.set_rotations(g.omega_init)
    total = 0
    while True:
        s, _ = S.iterate(3)
        total = s.num_iterations
        if s.termination != 0:
            break
    assert total == s1.num_iterations and s.termination == s1.termination
    assert np.array_equal(S.get_rotations(), one)  # deterministic: bit identical
    S.close()


@pytest.mark.parametrize("cfg", ["angle_axis", "covariance", "quaternion"])
def test_batch_execution_paths_agree(cfg):
    """One trust-region iteration can run as a CUDA graph replay of (fused PCG kernel, K1, node finalize), as direct
    launches of the same kernels, or as the un-fused sequence (k_prepare_solve, PCG, k_apply_step, k_node_prep, K1,
    k_node_finalize).  Same arithmetic per view and per edge; only the grouping of a few grid-wide sums differs, so the
    solves must take the same number of iterations and agree to rounding."""
    if cfg == "covariance":
        g = vg.synthetic_pose_graph(150, 2500, seed=23, noise_deg=1.0, outlier_fraction=0.05, covariance=True)
        et, loss = capi.ANGLE_AXIS_COVARIANCE, capi.Loss.make(capi.LOSS_SOFTLONE, 0.5)
    else:
        g = vg.synthetic_pose_graph(150, 2500, seed=23, noise_deg=1.0, outlier_fraction=0.05)
        et, loss = (capi.ANGLE_AXIS if cfg == "angle_axis" else capi.QUATERNION_COSINE), CAUCHY
    prob = solver.make_problem(g, et)
    o = capi.default_options_py()
    o.loss = loss
    o.pcg_rtol = 1e-10
    o.linear_solver = capi.SOLVER_PCG
    res = {}
    for name, env in (("graph", {}), ("direct", {"GSFM_RA_NO_GRAPH": "1"}), ("unfused", {"GSFM_RA_NO_GRAPH": "1", "GSFM_RA_NO_FUSE": "1"})):
        for k in ("GSFM_RA_NO_GRAPH", "GSFM_RA_NO_FUSE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            res[name] = solver.solve(prob, o, g.omega_init)
        finally:
            for k in env:
                os.environ.pop(k, None)
    om0, s0, _ = res["graph"]
    assert capi.TERMINATION[s0.termination] in ("FUNCTION_TOLERANCE", "PARAMETER_TOLERANCE", "GRADIENT_TOLERANCE")
    # graph replay and direct launches run the very same kernels: bit identical
    assert np.array_equal(om0, res["direct"][0]) and s0.num_iterations == res["direct"][1].num_iterations
    om2, s2, _ = res["unfused"]
    assert s2.num_iterations == s0.num_iterations and s2.termination == s0.termination
    assert abs(s2.final_cost - s0.final_cost) <= 1e-11 * abs(s0.final_cost)
    assert vg.mean_angular_error(om0, om2)[0] < 1e-9
    # and the un-fused run launches more kernels per iteration
    assert s2.kernel_launches > s0.kernel_launches


def test_composed_and_tabulated_losses():
    """ComposedLoss (scripts/loss_functions.py:250-265) runs as the native composition f(g(s)); any other LossFunction object as a
    table of its own Evaluate.  Both against the oracle, which for the table calls the Python object back once per edge per
    evaluation exactly as the reference does (bind_src/GlobalSfMpy.cpp:36-59)."""
    import math
    from globalsfmpy_b200.losses import loss_to_struct, tabulate_loss
    g = vg.synthetic_pose_graph(80, 900, seed=41, noise_deg=1.0, outlier_fraction=0.1)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    # (1) native composition: Cauchy(0.1) o 0.5 * Huber(0.05), scaled by 3
    L = capi.Loss.compose(capi.Loss.make(capi.LOSS_CAUCHY, 0.1, scale=3.0), capi.Loss.make(capi.LOSS_HUBER, 0.05, scale=0.5))
    sq = np.concatenate([[0.0], np.exp(np.linspace(np.log(1e-9), np.log(50.0), 400))])
    assert_close(solver.eval_loss(L, sq), orc.loss(L, sq), 1e-13, "composed loss table")
    og, sg, _ = solver.solve(prob, _tight(L), g.omega_init)
    oo, so, _ = orc.solve(prob, _tight(L, linear_solver=capi.SOLVER_DENSE_CHOLESKY), g.omega_init)
    assert vg.mean_angular_error(oo, og)[0] <= 1e-6 and abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost

    # (2) a hand-written subclass (smooth): tabulated on the host, interpolated on the device
    class LogCosh:                       # rho(s) = 2 a^2 log cosh(sqrt(s) / a), a = 0.05 -- written with care for small arguments
        def Evaluate(self, s, out):
            a2 = 0.0025
            u = s / a2                   # r^2
            if u < 1e-4:                 # series: no cancellation
                out[0] = 2 * a2 * (u / 2 - u * u / 12 + u ** 3 / 45)
                out[1] = 1 - u / 3 + 2 * u * u / 15
                out[2] = (-1 / 3 + 4 * u / 15) / a2
                return
            r = math.sqrt(u)
            out[0] = 2 * a2 * (r + math.log1p(math.exp(-2 * r)) - math.log(2.0))
            t = math.tanh(r)
            out[1] = t / r
            out[2] = ((1 - t * t) * r - t) / (2 * r ** 3) / a2
    obj = LogCosh()
    T = tabulate_loss(obj)          # verifies every cell midpoint against the object on the device
    assert T._table_error is not None and T._table_error < 1e-7
    ref = np.array([[0.0] * 3] * len(sq))
    for k, v in enumerate(sq):
        obj.Evaluate(float(v), ref[k])
    dev = solver.eval_loss(T, sq)
    assert_close(dev[:, :2], ref[:, :2], 1e-8, "tabulated loss: rho, rho'")
    assert_close(dev[:, 2], ref[:, 2], 1e-5, "tabulated loss: rho''")
    og, sg, _ = solver.solve(prob, _tight(T), g.omega_init)

    def cb(v):
        out = [0.0, 0.0, 0.0]
        obj.Evaluate(v, out)
        return out
    oo, so, _ = orc.solve(prob, _tight(capi.Loss.make(capi.LOSS_TRIVIAL), linear_solver=capi.SOLVER_DENSE_CHOLESKY), g.omega_init, loss_callback=cb)
    assert vg.mean_angular_error(oo, og)[0] <= 1e-6, vg.mean_angular_error(oo, og)
    assert abs(sg.final_cost - so.final_cost) <= 1e-8 * so.final_cost
    # through the object mapping used by the GlobalSfMpy-compatible module
    assert loss_to_struct(obj).kind == capi.LOSS_TABULATED


def test_filter_view_pairs():
    g = vg.synthetic_pose_graph(500, 20000, seed=4, noise_deg=1.0, outlier_fraction=0.2)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    keep, ang = solver.filter_view_pairs(prob, g.omega_gt, 15.0)
    keep0, ang0 = orc.filter_view_pairs(prob, g.omega_gt, 15.0)
    assert np.abs(ang - ang0).max() < 1e-12
    assert np.array_equal(keep, keep0)


def test_edge_cases():
    # a single edge, an isolated view, a duplicate edge, an out-of-range id
    prob = capi.ProblemArrays(3, [0], [1], [[0.1, 0.2, 0.3]])
    om, s, _ = solver.solve(prob, capi.default_options_py(), np.zeros((3, 3)))
    assert s.final_cost < 1e-12 and np.all(om[2] == 0)
    for ei, ej in (([0, 1], [1, 0]), ([0, 0], [1, 1]), ([0], [7]), ([2], [2])):
        p = capi.ProblemArrays(3, ei, ej, np.zeros((len(ei), 3)))
        with pytest.raises(capi.GsfmError) as e:
            solver.solve(p, capi.default_options_py(), np.zeros((3, 3)))
        assert e.value.code == capi.ERR_INVALID


def test_large_graph_properties():
    """BASELINE-size check without the oracle: 10k views / 1M edges. Properties: (1) the cost of the fused kernel
    equals the sum of per-edge rho from the diagnostic kernel; (2) gradient = sum rho' J^T r from the per-edge
    Jacobians; (3) SpMV symmetry; (4) a solve decreases the cost and lands near ground truth."""
    g = vg.synthetic_pose_graph(10000, 1000000, seed=56, noise_deg=1.0, outlier_fraction=0.1)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    c, gr, hd, rp, col, val = solver.assemble(prob, CAUCHY, g.omega_init)
    r, Ji, Jj, rho = solver.eval_edges(prob, CAUCHY, g.omega_init)
    assert abs(c - 0.5 * rho[:, 0].sum()) <= 1e-11 * c
    gi = np.zeros((10000, 3))
    np.add.at(gi, g.edge_i, np.einsum("e,eki,ek->ei", rho[:, 1], Ji, r))
    np.add.at(gi, g.edge_j, np.einsum("e,eki,ek->ei", rho[:, 1], Jj, r))
    assert_close(gr, gi, 1e-10, "gradient")
    rng = np.random.default_rng(0)
    x, x2 = rng.normal(size=(2, 10000, 3))
    y, y2 = solver.spmv(prob, CAUCHY, g.omega_init, x), solver.spmv(prob, CAUCHY, g.omega_init, x2)
    assert abs((x2 * y).sum() - (x * y2).sum()) < 1e-9 * abs((x * y2).sum())
    o = capi.default_options_py()
    o.loss = CAUCHY
    o.pcg_rtol = 1e-6
    om, s, _ = solver.solve(prob, o, g.omega_init)
    assert s.final_cost < s.initial_cost
    mean, _ = vg.mean_angular_error(g.omega_gt, om)
    assert np.degrees(mean) < 0.2
    # ... and WITH the oracle at this size: the whole assembled system at the initial point and the cost at the solution
    co, go, ho, rpo, colo, valo = orc.assemble(prob, CAUCHY, g.omega_init, num_threads=os.cpu_count())
    assert abs(c - co) <= 1e-12 * abs(co)
    assert_close(gr, go, 1e-10, "gradient vs oracle, 1M edges")
    assert_close(hd, ho, 1e-10, "diagonal blocks vs oracle, 1M edges")
    assert np.array_equal(rp, rpo) and np.array_equal(col, colo)
    assert_close(val, valo, 1e-10, "off-diagonal blocks vs oracle, 1M edges")
    assert abs(solver.cost(prob, CAUCHY, om) - orc.cost(prob, CAUCHY, om, num_threads=os.cpu_count())) <= 1e-12 * s.final_cost


def test_baseline_sizes_against_the_oracle():
    """The other BASELINE-size graphs against the oracle: the Piccadilly stand-in (2.3k views / 300k edges, Cauchy) assembled and
    SOLVED on both sides; covariance-weighted MAGSAC (the 20M configuration's kernel path) assembled at 1M edges."""
    ncpu = os.cpu_count()
    g = vg.synthetic_pose_graph(2300, 300000, seed=56, noise_deg=1.0, outlier_fraction=0.1, init="bfs")
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    c, gr, hd, rp, col, val = solver.assemble(prob, CAUCHY, g.omega_init)
    co, go, ho, rpo, colo, valo = orc.assemble(prob, CAUCHY, g.omega_init, num_threads=ncpu)
    assert abs(c - co) <= 1e-12 * abs(co)
    assert_close(gr, go, 1e-10, "gradient, 300k"); assert_close(hd, ho, 1e-10, "diag, 300k"); assert_close(val, valo, 1e-10, "blocks, 300k")
    # converged solve, smooth loss, tight on both sides (the oracle with PCG 1e-12: the 6900 x 6900 dense factorisation is not
    # what limits it, 300k Jets per evaluation are)
    t = _tight(CAUCHY)
    t.num_threads = ncpu
    og, sg, _ = solver.solve(prob, t, g.omega_init)
    oo, so, _ = orc.solve(prob, t, g.omega_init)
    mean, mx = vg.mean_angular_error(oo, og)
    assert mean <= 1e-6, (mean, mx, sg.final_cost, so.final_cost)
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    # covariance + MAGSAC at 1M edges / 20k views (the per-edge path of BASELINE configs[4])
    g = vg.synthetic_pose_graph(20000, 1000000, seed=57, outlier_fraction=0.1, covariance=True, init="bfs")
    prob = solver.make_problem(g, capi.ANGLE_AXIS_COVARIANCE)
    L = capi.Loss.make(capi.LOSS_MAGSAC3, 1.0)
    c, gr, hd, rp, col, val = solver.assemble(prob, L, g.omega_init)
    co, go, ho, rpo, colo, valo = orc.assemble(prob, L, g.omega_init, num_threads=ncpu)
    assert abs(c - co) <= 1e-11 * abs(co)
    assert_close(gr, go, 1e-9, "gradient, cov 1M"); assert_close(hd, ho, 1e-9, "diag, cov 1M"); assert_close(val, valo, 1e-9, "blocks, cov 1M")
