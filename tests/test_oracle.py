"""CPU tests: pin the ORACLE against every piece of reference material that exists for this path
(SURVEY.md section 8c), and check the host-side shaping code.  No GPU needed."""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from globalsfmpy_b200 import _capi as capi, viewgraph as vg
from oracle import ra_oracle as orc
from common import GOLDEN_LOSSES, assert_close

I3 = np.eye(3)


def _rz(deg):
    return Rotation.from_euler("z", deg, degrees=True).as_matrix()


def _xyz(ax, ay, az):
    # Eigen: AngleAxis(x) * AngleAxis(y) * AngleAxis(z)
    return (Rotation.from_euler("x", ax, degrees=True) * Rotation.from_euler("y", ay, degrees=True)
            * Rotation.from_euler("z", az, degrees=True)).as_matrix()


# The known-answer cases of T/sfm/global_pose_estimation/pairwise_rotation_error_test.cc:87-139
KNOWN = [
    ("small", _rz(1.0), 1.0, I3, _rz(2.0)),
    ("nontrivial", _xyz(5.9, 1.8, 7.6), 1.0, I3, _xyz(5.3, 1.2, 8.1)),
    ("180a", _rz(-179.0), 1.0, I3, _rz(179.0)),
    ("180b", _rz(-179.0), 1.0, _rz(179.0), I3),
    ("weight", _xyz(5.9, 1.8, 7.6), 2.0, I3, _xyz(5.3, 1.2, 8.1)),
]


@pytest.mark.parametrize("name,R12,w,R1,R2", KNOWN, ids=[k[0] for k in KNOWN])
def test_residual_known_answers(name, R12, w, R1, R2):
    """pairwise_rotation_error_test.cc: residual == w * Log(R2 R1^T R12^T), built independently (scipy); 1e-12."""
    gt = w * Rotation.from_matrix(R2 @ R1.T @ R12.T).as_rotvec()
    to_aa = lambda R: Rotation.from_matrix(R).as_rotvec()
    r, _, _ = orc.edge(to_aa(R1), to_aa(R2), to_aa(R12), w * I3)
    assert np.abs(r - gt).max() < 1e-12


def test_rotation_conversions_roundtrip():
    rng = np.random.default_rng(0)
    for w in np.concatenate([vg.random_rotation_vectors(rng, 50), 1e-9 * rng.normal(size=(5, 3)), np.zeros((1, 3))]):
        R = orc.angle_axis_to_matrix(w)
        assert np.abs(R - Rotation.from_rotvec(w).as_matrix()).max() < 1e-14
        assert np.abs(orc.matrix_to_angle_axis(R) - w).max() < 1e-12
        assert np.abs(vg.so3_exp(w) - R).max() < 1e-15
        assert np.abs(vg.so3_log(R) - orc.matrix_to_angle_axis(R)).max() < 1e-15


def test_jacobian_finite_differences():
    """The jets must reproduce d r / d omega (central differences of the oracle's own residual)."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        wi, wj, wij = vg.random_rotation_vectors(rng, 3) * rng.uniform(0.1, 1.0)
        U = np.triu(rng.normal(size=(3, 3))) + 2 * I3
        r, Ji, Jj = orc.edge(wi, wj, wij, U)
        h = 1e-6
        for k in range(3):
            d = np.zeros(3); d[k] = h
            fi = (orc.edge(wi + d, wj, wij, U)[0] - orc.edge(wi - d, wj, wij, U)[0]) / (2 * h)
            fj = (orc.edge(wi, wj + d, wij, U)[0] - orc.edge(wi, wj - d, wij, U)[0]) / (2 * h)
            assert np.abs(fi - Ji[:, k]).max() < 1e-7 * max(1, np.abs(Ji).max())
            assert np.abs(fj - Jj[:, k]).max() < 1e-7 * max(1, np.abs(Jj).max())


def test_losses_match_reference_python(golden_dir):
    """rho, rho', rho'' of every loss vs the UNMODIFIED scripts/loss_functions.py (fixture made by
    tests/golden/make_loss_golden.py).  rho of the MAGSAC losses goes through the closed-form gamma
    table (<= 1e-14 abs from the shipped table, times one_over_sigma), hence the absolute term."""
    z = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    s = z["s"]
    for name, L in GOLDEN_LOSSES.items():
        ref = z[name]
        got = orc.loss(L, s)
        # the inverse variants return 1/weight with weight -> 0 at the truncation point, which
        # amplifies the table's 1e-14 absolute difference: allow 1e-9 relative there
        tol = (1e-9 if "inv" in name else 1e-12) * np.abs(ref) + 1e-11
        bad = np.abs(got - ref) > tol
        assert not bad.any(), (name, s[bad.any(axis=1)][:5], got[bad][:5], ref[bad][:5])


def test_gamma_tables_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    for nu in (3, 4, 9):
        got = orc.gamma_table(nu, z[f"gamma{nu}_idx"])
        assert np.abs(got - z[f"gamma{nu}_val"]).max() < 2e-14


def test_whitening_matches_numpy(madrid):
    for k in range(0, madrid.num_edges, 997):
        c = madrid.cov6[k]
        S = 1e8 * np.array([[c[0], c[3], c[4]], [c[3], c[1], c[5]], [c[4], c[5], c[2]]])
        U = orc.whiten(capi.ANGLE_AXIS_COVARIANCE, c)
        P = np.linalg.inv(S)
        assert np.allclose(U.T @ U, P, rtol=1e-9, atol=1e-12 * np.abs(P).max())
        assert np.allclose(U, np.linalg.cholesky(P).T, rtol=1e-8, atol=1e-12 * np.abs(U).max())
        assert np.isclose(orc.whiten(capi.ANGLE_AXIS_COVTRACE, c)[0, 0], np.sqrt(1 / np.trace(S)))
        assert np.isclose(orc.whiten(capi.ANGLE_AXIS_COVNORM, c)[1, 1], np.sqrt(1 / np.linalg.norm(S)))


def test_madrid_fixture_shape(madrid):
    """379 views / 18 811 edges after the >=30-match filter + largest CC."""
    assert madrid.num_views == 379 and madrid.num_edges == 18811
    assert np.all(madrid.edge_i < madrid.edge_j)
    assert not np.isnan(madrid.omega_init).any()
    assert np.all(madrid.omega_init[0] == 0)  # root = smallest view index


def test_covariance_text_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    ids = np.array([[1, 2], [3, 9]])
    cov6, rot = rng.normal(size=(2, 6)) * 1e-9, rng.normal(size=(2, 3))
    p = tmp_path / "covariance_rot.txt"
    vg.write_covariance_text(p, ids, cov6, rot)
    i2, c2, r2 = vg.parse_covariance_text(p)
    assert np.array_equal(i2, ids) and np.array_equal(c2, cov6) and np.array_equal(r2, rot)  # bit exact


def _opts(loss, **kw):
    o = capi.default_options_py()
    o.loss = loss
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def test_solve_noise_free_small():
    """robust_rotation_estimator_test.cc fixture pattern: 4 views / 6 edges, no noise -> exact."""
    g = vg.synthetic_pose_graph(4, 6, seed=56, noise_deg=0.0, outlier_fraction=0.0, init="gt_perturbed")
    prob = capi.ProblemArrays(4, g.edge_i, g.edge_j, g.omega_ij)
    om, s, _ = orc.solve(prob, _opts(capi.Loss.make(capi.LOSS_SOFTLONE, 0.1)), g.omega_init)
    err, mx = vg.mean_angular_error(g.omega_gt, om)
    assert np.degrees(mx) < 1e-6, (err, mx)


def test_solve_noisy_medium():
    """100 views / 800 edges, 2 degree noise -> every view within 5 degrees of ground truth."""
    g = vg.synthetic_pose_graph(100, 800, seed=56, noise_deg=2.0, outlier_fraction=0.0)
    prob = capi.ProblemArrays(100, g.edge_i, g.edge_j, g.omega_ij)
    om, s, _ = orc.solve(prob, _opts(capi.Loss.make(capi.LOSS_SOFTLONE, 0.1)), g.omega_init)
    _, mx = vg.mean_angular_error(g.omega_gt, om)
    assert np.degrees(mx) < 5.0
    assert s.final_cost < s.initial_cost


def test_oracle_pcg_matches_dense():
    g = vg.synthetic_pose_graph(80, 600, seed=5, noise_deg=1.0, outlier_fraction=0.1)
    prob = capi.ProblemArrays(80, g.edge_i, g.edge_j, g.omega_ij)
    L = capi.Loss.make(capi.LOSS_CAUCHY, 0.05)
    a, sa, _ = orc.solve(prob, _opts(L, function_tolerance=1e-14, max_num_iterations=400), g.omega_init)
    b, sb, _ = orc.solve(prob, _opts(L, function_tolerance=1e-14, max_num_iterations=400, linear_solver=capi.SOLVER_PCG,
                                     pcg_rtol=1e-13, pcg_max_iterations=2000), g.omega_init)
    err, _ = vg.mean_angular_error(a, b)
    assert err < 1e-7 and abs(sa.final_cost - sb.final_cost) < 1e-9 * sa.final_cost


def test_assembled_system_is_consistent():
    """H x from the oracle's block structure equals J~^T J~ x formed from its own edge Jacobians."""
    g = vg.synthetic_pose_graph(30, 120, seed=2, covariance=True)
    prob = capi.ProblemArrays(30, g.edge_i, g.edge_j, g.omega_ij, cov6=g.cov6, error_type=capi.ANGLE_AXIS_COVARIANCE)
    L = capi.Loss.make(capi.LOSS_MAGSAC9, 0.5)  # exercises the rho'' > 0 (Triggs) branch
    cost, grad, hd, rowptr, col, val = orc.assemble(prob, L, g.omega_init)
    r, Ji, Jj, rho = orc.eval_edges(prob, L, g.omega_init)
    assert (rho[:, 2] > 0).any()
    assert np.isclose(cost, 0.5 * rho[:, 0].sum(), rtol=1e-13)
    assert np.allclose(np.einsum("e,eki,ek->ei", rho[:, 1], Ji, r)[np.argsort(g.edge_i, kind="stable")].sum(), 
                       np.einsum("e,eki,ek->ei", rho[:, 1], Ji, r).sum())
    gi = np.zeros((30, 3))
    np.add.at(gi, g.edge_i, np.einsum("e,eki,ek->ei", rho[:, 1], Ji, r))
    np.add.at(gi, g.edge_j, np.einsum("e,eki,ek->ei", rho[:, 1], Jj, r))
    assert_close(grad, gi, 1e-12, "gradient = sum rho' J^T r")
    H = np.zeros((90, 90))
    for a in range(30):
        H[3 * a:3 * a + 3, 3 * a:3 * a + 3] = hd[a]
        for s in range(rowptr[a], rowptr[a + 1]):
            H[3 * a:3 * a + 3, 3 * col[s]:3 * col[s] + 3] = val[s]
    assert np.abs(H - H.T).max() <= 1e-12 * np.abs(H).max()
    assert np.linalg.eigvalsh(H).min() > -1e-9 * np.abs(H).max()


def test_filter_view_pairs():
    g = vg.synthetic_pose_graph(50, 300, seed=4, noise_deg=1.0, outlier_fraction=0.2)
    prob = capi.ProblemArrays(50, g.edge_i, g.edge_j, g.omega_ij)
    keep, ang = orc.filter_view_pairs(prob, g.omega_gt, 15.0)
    E = vg.so3_exp(g.omega_gt[g.edge_j]) @ np.transpose(vg.so3_exp(g.omega_gt[g.edge_i]), (0, 2, 1)) @ \
        np.transpose(vg.so3_exp(g.omega_ij), (0, 2, 1))
    ref = np.linalg.norm(vg.so3_log(E), axis=1)
    assert np.abs(ang - ref).max() < 1e-12
    assert np.array_equal(keep, ref <= np.radians(15.0))
    assert keep[~g.is_outlier].all() and (~keep[g.is_outlier]).mean() > 0.9


def _quat_xyzw(w):
    """ceres AngleAxisToQuaternion, Eigen coefficient order (x, y, z, w)."""
    t = np.linalg.norm(w)
    if t == 0:
        return np.array([0.0, 0.0, 0.0, 1.0])
    return np.concatenate([np.sin(t / 2) * w / t, [np.cos(t / 2)]])


def _qmul_xyzw(a, b):
    av, aw, bv, bw = a[:3], a[3], b[:3], b[3]
    return np.concatenate([aw * bv + bw * av + np.cross(av, bv), [aw * bw - av @ bv]])


def _eigen_rot(q):
    """Eigen::Quaternion::toRotationMatrix, coefficients (x, y, z, w)."""
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _general_residual(etype, qa, qb, qrel):
    """include/pairwise_rotation_error_quat.hpp:82-106 (Quat), :125-150 (QuatFNorm) and :169-196 (RotFNorm), restated with numpy."""
    if etype == capi.QUATERNION_COSINE:
        # hpp:82-106: delta_q = q_rel * conj(q_b * conj(q_a)); residual = weight * 2 * delta_q.vec()
        conj = lambda q: np.concatenate([-q[:3], q[3:]])
        return 2.0 * _qmul_xyzw(qrel, conj(_qmul_xyzw(qb, conj(qa))))[:3]
    if etype == capi.QUATERNION_NORM:
        est = _qmul_xyzw(qrel, qa)
        b = -qb if qb[1] < 0 else qb
        e = -est if est[1] < 0 else est
        return b - e
    return (_eigen_rot(qrel) @ _eigen_rot(qa) - _eigen_rot(qb)).reshape(-1, order="F")  # Eigen linear index = column-major


@pytest.mark.parametrize("etype", [capi.QUATERNION_COSINE, capi.QUATERNION_NORM, capi.ROTATION_MAT_FNORM])
def test_general_residual_types(etype):
    """The oracle's jet evaluation of the three quaternion-parameter functors (3-, 4- and 9-dimensional residuals; hpp:82-106,
    :125-150, :169-196) against a numpy restatement of the reference formulas,
    and its local-coordinate Jacobians against finite differences through EigenQuaternionParameterization::Plus
    (x (+) d = [sin|d| d/|d|, cos|d|] (x) x)."""
    g = vg.synthetic_pose_graph(12, 40, seed=3, noise_deg=5.0, outlier_fraction=0.2)
    prob = capi.ProblemArrays(12, g.edge_i, g.edge_j, g.omega_ij, error_type=etype)
    rng = np.random.default_rng(0)
    omega = g.omega_init + 0.1 * rng.normal(size=g.omega_init.shape)
    L = capi.Loss.make(capi.LOSS_TRIVIAL)
    r, Ji, Jj, _ = orc.eval_edges(prob, L, omega)
    d = capi.residual_dim(etype)
    assert r.shape == (40, d) and Ji.shape == (40, d, 3)

    def plus(q, dl):
        n = np.linalg.norm(dl)
        return _qmul_xyzw(np.concatenate([np.sin(n) * dl / n, [np.cos(n)]]), q) if n > 0 else q

    h = 1e-6
    for k in range(40):
        qa, qb, qrel = _quat_xyzw(omega[g.edge_i[k]]), _quat_xyzw(omega[g.edge_j[k]]), _quat_xyzw(g.omega_ij[k])
        assert np.allclose(r[k], _general_residual(etype, qa, qb, qrel), atol=1e-14)
        for c in range(3):
            e = np.zeros(3); e[c] = h
            fa = (_general_residual(etype, plus(qa, e), qb, qrel) - _general_residual(etype, plus(qa, -e), qb, qrel)) / (2 * h)
            fb = (_general_residual(etype, qa, plus(qb, e), qrel) - _general_residual(etype, qa, plus(qb, -e), qrel)) / (2 * h)
            assert np.allclose(Ji[k][:, c], fa, atol=1e-8) and np.allclose(Jj[k][:, c], fb, atol=1e-8)
    # the solve runs and reduces the cost
    o = _opts(capi.Loss.make(capi.LOSS_HUBER, 0.1), linear_solver=capi.SOLVER_DENSE_CHOLESKY)
    om, s, _ = orc.solve(prob, o, g.omega_init)
    assert s.final_cost < s.initial_cost and np.isfinite(om).all()
