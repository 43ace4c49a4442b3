"""Translation averaging (reference src/GSfM_nonlinear_position_estimator.cpp): the oracle pinned on Theia's
known-answer residuals and on numpy restatements (CPU), and the CUDA path through include/gsfm_pa.h against the oracle (GPU).

Tolerances (fp64 end to end): per-pair residuals / Jacobians / assembled blocks 1e-12 relative (closed form vs jets, a handful
of flops each); converged positions 1e-7 of the scene scale where both sides factor exactly, 1e-5 where both run PCG."""
import ctypes as C

import numpy as np
import pytest

from globalsfmpy_b200 import _capi as capi, positions as P, solver, viewgraph as vg
from oracle import ra_oracle as orc
from common import assert_close

HUBER = capi.Loss.make(capi.LOSS_HUBER, 0.1)


def _problem(n, e, seed=3, fixed=0, **kw):
    d = P.synthetic_position_graph(n, e, seed=seed, **kw)
    pp = P.PositionProblemArrays(d["num_views"], d["edge_i"], d["edge_j"], d["position_2"], d["orientation"], fixed_view=fixed)
    return d, pp, pp.as_rotation_solver_problem()


def _two_view(p1, p2, direction, weight):
    pp = P.PositionProblemArrays(2, [0], [1], [direction], np.zeros((2, 3)), edge_weight=[weight], fixed_view=-1)
    return pp, np.array([p1, p2], dtype=np.float64)


def _align_similarity(x, ref):
    """Least-squares similarity transform (Umeyama) of x onto ref: the gauge of a translation-averaging solution
    (src/compare_reconstructions.cpp aligns positions the same way before it reports position errors)."""
    a0, b0 = x - x.mean(0), ref - ref.mean(0)
    U, S, Vt = np.linalg.svd(b0.T @ a0)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    R = U @ D @ Vt
    s = (S * np.diag(D)).sum() / (a0 ** 2).sum()
    return s * (a0 @ R.T) + ref.mean(0)


# ------------------------------------------------------------------------------------------------ CPU: the oracle
# thirdparty/TheiaSfM/src/theia/sfm/global_pose_estimation/pairwise_translation_error_test.cc:87-127
_NOISY = (np.array([1.0, 0.0, 0.0]) + 0.01) / np.linalg.norm(np.array([1.0, 0.0, 0.0]) + 0.01)
KNOWN = [
    ("TranslationNoNoise", [0, 0, 0], [1, 0, 0], [1.0, 0.0, 0.0], 1.0),
    ("TranslationWithNoise", [0, 0, 0], [1, 0, 0], _NOISY, 1.0),
    ("NontrivialWeight", [0, 0, 0], [1, 0, 0], _NOISY, 1.1),
    ("NoTranslation", [1, 0, 0], [1, 0, 0], [0.0, 0.0, 0.0], 1.0),   # :69-80: coincident positions, the norm-tolerance branch
]


@pytest.mark.parametrize("name,p1,p2,direction,weight", KNOWN)
def test_residual_known_answers(name, p1, p2, direction, weight):
    """The expected value is the test's own: weight * (normalize(p2 - p1) - direction), with no normalisation below 1e-8."""
    pp, x = _two_view(p1, p2, direction, weight)
    r, _, _, _ = orc.eval_edges(pp.as_rotation_solver_problem(), HUBER, x)
    t = np.asarray(p2, float) - np.asarray(p1, float)
    if np.linalg.norm(t) > 1e-8:
        t = t / np.linalg.norm(t)
    expected = weight * (t - np.asarray(direction, float))
    assert np.abs(r[0] - expected).max() <= 4e-16, (name, r[0], expected)   # EXPECT_DOUBLE_EQ: 4 ulp


def test_rotated_direction_and_jacobians_match_numpy():
    d, pp, rp = _problem(25, 120)
    rng = np.random.default_rng(1)
    x = d["positions_gt"] + 0.3 * rng.normal(size=(25, 3))
    w = rng.uniform(0.5, 2.0, size=120)
    pp = P.PositionProblemArrays(25, d["edge_i"], d["edge_j"], d["position_2"], d["orientation"], edge_weight=w)
    r, Ji, Jj, rho = orc.eval_edges(pp.as_rotation_solver_problem(), HUBER, x)
    # GetRotatedTranslation (position_estimator.cpp:36-44): R(orientation_1)^T position_2
    R = vg.so3_exp(d["orientation"])
    t = np.einsum("eba,eb->ea", R[d["edge_i"]], d["position_2"])
    dd = x[d["edge_j"]] - x[d["edge_i"]]
    n = np.linalg.norm(dd, axis=1, keepdims=True)
    u = dd / n
    assert np.abs(r - w[:, None] * (u - t)).max() <= 1e-15
    B = (w / n[:, 0])[:, None, None] * (np.eye(3) - u[:, :, None] * u[:, None, :])
    assert np.abs(Jj - B).max() <= 1e-14 and np.abs(Ji + B).max() <= 1e-14
    s = (r * r).sum(1)
    assert np.allclose(rho, orc.loss(HUBER, s), rtol=0, atol=0)
    # central differences of the residual
    h = 1e-6
    for c in range(3):
        e = np.zeros((25, 3)); e[:, c] = h
        for k in (0, 17, 63):
            xp, xm = x.copy(), x.copy()
            xp[d["edge_j"][k], c] += h; xm[d["edge_j"][k], c] -= h
            rp_, rm_ = orc.eval_edges(pp.as_rotation_solver_problem(), HUBER, xp)[0][k], orc.eval_edges(pp.as_rotation_solver_problem(), HUBER, xm)[0][k]
            assert np.abs((rp_ - rm_) / (2 * h) - Jj[k][:, c]).max() <= 1e-8


def test_coincident_positions_use_the_constant_norm():
    """All cameras at the origin -- where the reference STARTS (position_estimator.cpp:236-257 zeroes its random draw): the norm
    is replaced by the constant 1, so r = -w t and the Jacobians are -+ w I (no NaN from the dual part of sqrt(0))."""
    d, pp, rp = _problem(12, 40)
    r, Ji, Jj, _ = orc.eval_edges(rp, HUBER, np.zeros((12, 3)))
    R = vg.so3_exp(d["orientation"])
    t = np.einsum("eba,eb->ea", R[d["edge_i"]], d["position_2"])
    assert np.abs(r + t).max() <= 4e-16
    assert np.array_equal(Jj, np.broadcast_to(np.eye(3), Jj.shape)) and np.array_equal(Ji, -Jj)
    c = orc.cost(rp, HUBER, np.zeros((12, 3)))
    assert abs(c - 40 * 0.5 * (2 * 0.1 * 1.0 - 0.01)) <= 1e-12     # |r| = 1 for every pair: Huber's linear branch


def test_fixed_view_is_removed_from_the_system():
    d, pp, rp = _problem(15, 60, fixed=4)
    x = d["positions_gt"] + 0.05
    cost, g, hd, rowptr, col, val = orc.assemble(rp, HUBER, x)
    assert np.array_equal(g[4], np.zeros(3))
    for a in range(15):
        for s_ in range(rowptr[a], rowptr[a + 1]):
            if a == 4 or col[s_] == 4:
                assert not val[s_].any()
            else:
                assert val[s_].any()
    # the neighbours keep the pair's contribution on their diagonal: same as the unconstrained assembly
    _, pp2, rp2 = _problem(15, 60, fixed=-1)
    _, g2, hd2, _, _, val2 = orc.assemble(rp2, HUBER, x)
    assert np.array_equal(hd, hd2) and np.array_equal(np.delete(g, 4, 0), np.delete(g2, 4, 0))


def test_oracle_recovers_noise_free_positions():
    d, pp, rp = _problem(40, 300, direction_noise_deg=0.0, outlier_fraction=0.0)
    o = capi.default_options_py()
    o.loss = HUBER
    o.max_num_iterations = 400
    o.function_tolerance = 1e-14
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    x, s, _ = orc.solve(rp, o, np.zeros((40, 3)))
    assert s.final_cost <= 1e-12 * s.initial_cost
    assert np.array_equal(x[0], np.zeros(3))                      # the constant view stays where it started
    err = np.linalg.norm(_align_similarity(x, d["positions_gt"]) - d["positions_gt"], axis=1).max()
    assert err <= 1e-5 * 10.0, err


def test_default_options_of_the_position_estimator():
    o = capi.Options()
    capi.lib().gsfm_pa_default_options(C.byref(o))
    assert o.max_num_iterations == 400 and o.loss.kind == capi.LOSS_HUBER and o.loss.p[0] == 0.1
    assert (o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance) == (1e-6, 1e-10, 1e-8)


def test_position_problem_is_checked_before_the_device():
    d, pp, rp = _problem(8, 20)
    bad = capi.ProblemArrays(8, d["edge_i"], d["edge_j"], d["position_2"], error_type=capi.POSITION_BASELINE)   # no orientations
    with pytest.raises(capi.GsfmError) as e:
        solver.cost(bad, HUBER, np.zeros((8, 3)))
    assert e.value.code == capi.ERR_INVALID
    q = capi.Problem()
    assert capi.lib().gsfm_pa_as_ra_problem(C.byref(pp.c), C.byref(q)) == 0
    assert q.error_type == capi.POSITION_BASELINE and q.fixed_view == 0 and q.num_edges == 20
    pp.c.error_type = 7
    assert capi.lib().gsfm_pa_as_ra_problem(C.byref(pp.c), C.byref(q)) == capi.ERR_INVALID


def test_no_cpu_fallback_for_positions():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d, pp, rp = _problem(8, 20)
    with pytest.raises(capi.GsfmError) as e:
        P.solve(pp, P.default_options())
    assert e.value.code == capi.ERR_NO_DEVICE
    with pytest.raises(capi.GsfmError) as e:
        P.eval_edges(pp, HUBER, np.zeros((8, 3)))
    assert e.value.code == capi.ERR_NO_DEVICE


# ------------------------------------------------------------------------------------------------ GPU: the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name,p1,p2,direction,weight", KNOWN)
def test_gpu_residual_known_answers(name, p1, p2, direction, weight):
    pp, x = _two_view(p1, p2, direction, weight)
    r, Ji, Jj, _ = P.eval_edges(pp, HUBER, x)
    t = np.asarray(p2, float) - np.asarray(p1, float)
    if np.linalg.norm(t) > 1e-8:
        t = t / np.linalg.norm(t)
    assert np.abs(r[0] - weight * (t - np.asarray(direction, float))).max() <= 4e-16
    ro, Jio, Jjo, _ = orc.eval_edges(pp.as_rotation_solver_problem(), HUBER, x)
    assert np.abs(Jj - Jjo).max() <= 1e-15 and np.abs(Ji - Jio).max() <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("at_origin", [False, True])
def test_gpu_eval_edges_match_oracle(at_origin):
    d, pp, rp = _problem(300, 4000, seed=11)
    w = np.random.default_rng(2).uniform(0.5, 2.0, size=4000)
    pp = P.PositionProblemArrays(300, d["edge_i"], d["edge_j"], d["position_2"], d["orientation"], edge_weight=w)
    x = np.zeros((300, 3)) if at_origin else d["positions_gt"] + 0.5 * np.random.default_rng(5).normal(size=(300, 3))
    for L in (HUBER, capi.Loss.make(capi.LOSS_CAUCHY, 0.3), capi.Loss.make(capi.LOSS_MAGSAC9, 0.3)):
        r, Ji, Jj, rho = P.eval_edges(pp, L, x)
        ro, Jio, Jjo, rhoo = orc.eval_edges(pp.as_rotation_solver_problem(), L, x)
        assert_close(r, ro, 1e-13, "residual")
        assert_close(Jj, Jjo, 1e-12, "d r / d c_j")
        assert_close(Ji, Jio, 1e-12, "d r / d c_i")
        assert np.abs(rho - rhoo).max() <= 1e-12 * np.abs(rhoo).max() + 1e-9 * (L.kind >= capi.LOSS_MAGSAC3)
        assert abs(P.cost(pp, L, x) - orc.cost(pp.as_rotation_solver_problem(), L, x)) <= 1e-12 * (1 + 0.5 * rhoo[:, 0].sum())


@pytest.mark.gpu
@pytest.mark.parametrize("fixed", [0, 17, -1])
@pytest.mark.parametrize("loss", [HUBER, capi.Loss.make(capi.LOSS_MAGSAC9, 0.3), capi.Loss.make(capi.LOSS_TUKEY, 0.8)])
def test_gpu_assemble_matches_oracle(fixed, loss):
    """Normal equations incl. the Triggs correction (MAGSAC nu = 9 and Tukey have rho'' > 0 regions) and the constant view."""
    d, pp, rp = _problem(200, 3000, seed=7, fixed=fixed)
    x = d["positions_gt"] + 0.4 * np.random.default_rng(9).normal(size=(200, 3))
    cg, gg, hg, rpg, colg, vg_ = solver.assemble(rp, loss, x)
    co, go, ho, rpo, colo, vo = orc.assemble(rp, loss, x)
    assert np.array_equal(rpg, rpo) and np.array_equal(colg, colo)
    assert abs(cg - co) <= 1e-12 * abs(co)
    assert_close(gg, go, 1e-11, "gradient")
    assert_close(hg, ho, 1e-11, "diagonal blocks")
    assert_close(vg_, vo, 1e-11, "off-diagonal blocks")
    if fixed >= 0:
        assert not gg[fixed].any() and not vg_[rpg[fixed]:rpg[fixed + 1]].any() and not vg_[colg == fixed].any()


@pytest.mark.gpu
def test_gpu_solve_tracks_the_oracle_from_the_origin():
    """The reference's own start (every camera at the origin), exact factorisation on both sides: same trajectory."""
    d, pp, rp = _problem(120, 1500, seed=5, fixed=3)
    o = P.default_options()
    o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
    x, s, tr = P.solve(pp, o, trace_capacity=401)
    xo, so, tro = orc.solve(rp, o, np.zeros((120, 3)), trace_capacity=401)
    assert s.num_iterations == so.num_iterations and s.termination == so.termination
    assert abs(s.initial_cost - 1500 * 0.5 * 0.19) <= 1e-9
    for a, b in zip(tr, tro):
        assert abs(a.cost - b.cost) <= 1e-9 * abs(b.cost), (a.iteration, a.cost, b.cost)
    assert np.array_equal(x[3], np.zeros(3))
    assert np.abs(x - xo).max() <= 1e-7 * np.abs(xo).max()
    err = np.linalg.norm(_align_similarity(x, d["positions_gt"]) - d["positions_gt"], axis=1)
    assert np.median(err) <= 0.5, np.median(err)                  # 1 degree of direction noise, 10 % outliers, scene of +-10


@pytest.mark.gpu
def test_gpu_pcg_path_converges_to_the_oracle_solution():
    """More than GSFM_RA_AUTO_DENSE_MAX_VIEWS cameras: AUTO resolves to the persistent PCG kernel (the reference switches to CGNR +
    JACOBI above 1000 cameras, position_estimator.cpp:129-140).  Both sides at a tight linear tolerance and a tight function
    tolerance; compared after removing the similarity gauge the one constant view does not fix (scale)."""
    d, pp, rp = _problem(2000, 30000, seed=13, outlier_fraction=0.05)
    o = P.default_options()
    o.function_tolerance = 1e-12
    o.pcg_rtol = 1e-12
    o.pcg_max_iterations = 2000
    x, s, _ = P.solve(pp, o)
    # (the cost is invariant to the scale of the scene, so once the damping has relaxed the normal equations are singular along
    # that direction and a solve or two may stop at the iteration cap -- Ceres' CGNR caps at 500 the same way)
    assert s.n_gpus_used == 1 and s.total_linear_iterations > 0
    o.linear_solver = capi.SOLVER_PCG
    xo, so, _ = orc.solve(rp, o, np.zeros((2000, 3)))
    print(f"PCG path: {s.num_iterations} iterations / {s.total_linear_iterations} CG steps ({s.num_linear_unconverged} capped), cost {s.final_cost:.12g}; "
          f"oracle {so.num_iterations} iterations, cost {so.final_cost:.12g}")
    assert abs(s.final_cost - so.final_cost) <= 1e-6 * so.final_cost, (s.final_cost, so.final_cost)
    err = np.linalg.norm(_align_similarity(x, xo) - xo, axis=1).mean() / np.abs(xo).max()
    assert err <= 1e-4, err
    # cost at the GPU's solution, evaluated by the oracle: the same number
    assert abs(orc.cost(rp, o.loss, x) - s.final_cost) <= 1e-11 * s.final_cost


@pytest.mark.gpu
def test_gpu_assemble_at_baseline_size():
    """10k cameras / 1M pairs (the synthetic graph of BASELINE config 4) through K1's translation instantiation."""
    d, pp, rp = _problem(10000, 1000000, seed=56)
    x = d["positions_gt"] + 0.2 * np.random.default_rng(3).normal(size=(10000, 3))
    cg, gg, hg, _, _, _ = solver.assemble(rp, HUBER, x)
    co, go, ho, _, _, _ = orc.assemble(rp, HUBER, x)
    assert abs(cg - co) <= 1e-11 * co
    assert_close(gg, go, 1e-10, "gradient")
    assert_close(hg, ho, 1e-10, "diagonal blocks")


@pytest.mark.gpu
def test_gpu_sharded_solve_matches_single_device():
    if capi.lib().gsfm_ra_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d, pp, rp = _problem(3000, 60000, seed=21)
    o = P.default_options()
    o.linear_solver = capi.SOLVER_PCG
    o.pcg_rtol = 1e-12
    x1, s1, _ = P.solve(pp, o)
    o.n_gpus = 2
    import os
    os.environ["GSFM_RA_MIN_EDGES_PER_GPU"] = "1000"
    try:
        x2, s2, _ = P.solve(pp, o)
    finally:
        del os.environ["GSFM_RA_MIN_EDGES_PER_GPU"]
    assert s2.n_gpus_used == 2
    assert abs(s1.final_cost - s2.final_cost) <= 1e-9 * s1.final_cost
    assert np.linalg.norm(_align_similarity(x2, x1) - x1, axis=1).mean() <= 1e-6 * np.abs(x1).max()
