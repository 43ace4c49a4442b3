"""Host-side input shaping for the rotation-averaging path (numpy, one-time O(E log E)).

Restates what produces the exact (edges, initial orientations) the reference solver sees
(SURVEY.md section 8 a8):
  * 1DSfM EGs convention           T/io/read_1dsfm.cc:299-373   (T/ = thirdparty/TheiaSfM/src/theia/)
  * covariance_rot.txt reader      src/uncertainty.cpp:200-229
  * FilterInitialViewGraph         src/GSfM_global_reconstruction_estimator.cpp:369-390
  * OrientationsFromMaximumSpanningTree   T/sfm/view_graph/orientations_from_maximum_spanning_tree.cc:109-178
  * the synthetic pose-graph fixture of T/sfm/global_pose_estimation/robust_rotation_estimator_test.cc:150-243
  * gauge alignment + AngularDifference, the parity metric  src/compare_reconstructions.cpp:7-16,149-177
"""
import numpy as np

DBL_EPSILON = np.finfo(np.float64).eps


# --------------------------------------------------------------------------- SO(3), Ceres conventions
def so3_exp(w):
    """Angle-axis [...,3] -> rotation matrices [...,3,3] (ceres AngleAxisToRotationMatrix)."""
    w = np.asarray(w, dtype=np.float64)
    theta2 = np.sum(w * w, axis=-1)
    small = theta2 <= DBL_EPSILON
    theta = np.sqrt(np.where(small, 1.0, theta2))
    a = w / theta[..., None]
    c, s = np.cos(theta), np.sin(theta)
    c = np.where(small, 1.0, c)
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -a[..., 2], a[..., 1]
    K[..., 1, 0], K[..., 1, 2] = a[..., 2], -a[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -a[..., 1], a[..., 0]
    R = (c[..., None, None] * np.eye(3) + s[..., None, None] * K
         + (1.0 - c)[..., None, None] * a[..., :, None] * a[..., None, :])
    if np.any(small):
        Ks = np.zeros_like(K)
        Ks[..., 0, 1], Ks[..., 0, 2] = -w[..., 2], w[..., 1]
        Ks[..., 1, 0], Ks[..., 1, 2] = w[..., 2], -w[..., 0]
        Ks[..., 2, 0], Ks[..., 2, 1] = -w[..., 1], w[..., 0]
        R = np.where(small[..., None, None], np.eye(3) + Ks, R)
    return R


def so3_log(R):
    """Rotation matrices [...,3,3] -> angle-axis [...,3] (ceres RotationMatrixToQuaternion +
    QuaternionToAngleAxis; angle in [0, pi])."""
    R = np.asarray(R, dtype=np.float64)
    flat = R.reshape(-1, 3, 3)
    n = flat.shape[0]
    q = np.zeros((n, 4))
    tr = flat[:, 0, 0] + flat[:, 1, 1] + flat[:, 2, 2]
    pos = tr >= 0
    if np.any(pos):
        M = flat[pos]
        t = np.sqrt(tr[pos] + 1.0)
        q[pos, 0] = 0.5 * t
        t = 0.5 / t
        q[pos, 1] = (M[:, 2, 1] - M[:, 1, 2]) * t
        q[pos, 2] = (M[:, 0, 2] - M[:, 2, 0]) * t
        q[pos, 3] = (M[:, 1, 0] - M[:, 0, 1]) * t
    neg = np.nonzero(~pos)[0]
    if len(neg):
        M = flat[neg]
        d = np.stack([M[:, 0, 0], M[:, 1, 1], M[:, 2, 2]], axis=1)
        i = np.zeros(len(neg), dtype=np.int64)
        i = np.where(d[:, 1] > d[:, 0], 1, i)
        i = np.where(d[:, 2] > d[np.arange(len(neg)), i], 2, i)
        j = (i + 1) % 3
        k = (j + 1) % 3
        ar = np.arange(len(neg))
        t = np.sqrt(M[ar, i, i] - M[ar, j, j] - M[ar, k, k] + 1.0)
        q[neg, i + 1] = 0.5 * t
        t = 0.5 / t
        q[neg, 0] = (M[ar, k, j] - M[ar, j, k]) * t
        q[neg, j + 1] = (M[ar, j, i] + M[ar, i, j]) * t
        q[neg, k + 1] = (M[ar, k, i] + M[ar, i, k]) * t
    sin2 = np.sum(q[:, 1:] ** 2, axis=1)
    s = np.sqrt(sin2)
    c = q[:, 0]
    two_theta = 2.0 * np.where(c < 0, np.arctan2(-s, -c), np.arctan2(s, c))
    kk = np.where(sin2 > 0, two_theta / np.where(sin2 > 0, s, 1.0), 2.0)
    return (q[:, 1:] * kk[:, None]).reshape(R.shape[:-2] + (3,))


def random_rotation_vectors(rng, n):
    """Uniform on SO(3) as angle-axis (via normalised Gaussian quaternions)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    s = np.linalg.norm(q[:, 1:], axis=1)
    ang = 2 * np.arctan2(s, q[:, 0])
    return q[:, 1:] * (ang / np.where(s > 0, s, 1.0))[:, None]


# --------------------------------------------------------------------------- graph plumbing
def largest_connected_component(num_nodes, ei, ej):
    """Boolean mask over nodes of the largest connected component (ties: the one holding the
    smallest node id)."""
    parent = np.arange(num_nodes)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for a, b in zip(ei.tolist(), ej.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    roots = np.array([find(a) for a in range(num_nodes)])
    touched = np.zeros(num_nodes, dtype=bool)
    touched[ei] = True
    touched[ej] = True
    cnt = np.bincount(roots[touched], minlength=num_nodes)
    best = int(np.argmax(cnt))
    return (roots == best) & touched


def filter_initial_view_graph(view_ids, edge_ij, num_verified_matches, min_num_two_view_inliers=30):
    """src/GSfM_global_reconstruction_estimator.cpp:369-390: drop edges with fewer verified
    matches than the threshold, keep the largest connected component.  Returns the boolean
    edge mask and the sorted view ids that remain."""
    edge_ij = np.asarray(edge_ij)
    keep = np.asarray(num_verified_matches) >= min_num_two_view_inliers
    ids = np.unique(np.concatenate([np.asarray(view_ids).ravel(), edge_ij.ravel()]))
    idx = np.searchsorted(ids, edge_ij)
    cc = largest_connected_component(len(ids), idx[keep, 0], idx[keep, 1])
    keep &= cc[idx[:, 0]] & cc[idx[:, 1]]
    return keep, ids[cc]


def max_spanning_tree_orientations(num_views, ei, ej, omega_ij, weights, root=None):
    """OrientationsFromMaximumSpanningTree (orientations_from_maximum_spanning_tree.cc:109-178).
    Kruskal over edges sorted by (-weight, i, j) (math/graph/minimum_spanning_tree.h:70-98), then
    chain R_neighbor = (src < nbr ? R_rel : R_rel^T) * R_src (:60-83) from the root, which the
    reference takes from an unordered_set (implementation-defined); here: the smallest view index.
    Views outside the root's component keep NaN."""
    ei, ej = np.asarray(ei, dtype=np.int64), np.asarray(ej, dtype=np.int64)
    order = np.lexsort((ej, ei, -np.asarray(weights, dtype=np.int64)))
    parent = np.arange(num_views)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    adj = [[] for _ in range(num_views)]
    n_tree = 0
    for k in order.tolist():
        a, b = int(ei[k]), int(ej[k])
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            adj[a].append((b, k))
            adj[b].append((a, k))
            n_tree += 1
            if n_tree == num_views - 1:
                break
    Rrel = so3_exp(omega_ij)
    omega = np.full((num_views, 3), np.nan)
    if root is None:
        root = int(min(ei.min(), ej.min()))
    omega[root] = 0.0
    stack = [root]
    while stack:
        a = stack.pop()
        Ra = so3_exp(omega[a])
        for b, k in adj[a]:
            if not np.isnan(omega[b, 0]):
                continue
            lo = min(a, b)
            # edge k stores R_hi = R_rel * R_lo  (TwoViewInfo::rotation_2, twoview_info.h:123-126)
            Rb = Rrel[k] @ Ra if a == lo else Rrel[k].T @ Ra
            omega[b] = so3_log(Rb)
            stack.append(b)
    return omega


def spanning_tree_orientations_bfs(num_views, ei, ej, Rrel, root=0):
    """Breadth-first spanning-tree initialisation (level-synchronous, vectorised): the role
    OrientationsFromMaximumSpanningTree plays in the pipeline, for graphs too large for a Python
    Kruskal.  R_child = R_rel R_parent along (parent=i, child=j) edges, R_rel^T R_parent otherwise."""
    ei, ej = np.asarray(ei, dtype=np.int64), np.asarray(ej, dtype=np.int64)
    R = np.zeros((num_views, 3, 3))
    R[root] = np.eye(3)
    done = np.zeros(num_views, dtype=bool)
    done[root] = True
    while not done.all():
        fwd = done[ei] & ~done[ej]
        bwd = done[ej] & ~done[ei]
        if not (fwd.any() or bwd.any()):
            break
        child = np.concatenate([ej[fwd], ei[bwd]])
        k = np.concatenate([np.nonzero(fwd)[0], np.nonzero(bwd)[0]])
        is_fwd = np.concatenate([np.ones(int(fwd.sum()), bool), np.zeros(int(bwd.sum()), bool)])
        _, first = np.unique(child, return_index=True)  # first discovering edge wins
        child, k, is_fwd = child[first], k[first], is_fwd[first]
        parent = np.where(is_fwd, ei[k], ej[k])
        Rk = np.where(is_fwd[:, None, None], Rrel[k], np.transpose(Rrel[k], (0, 2, 1)))
        R[child] = Rk @ R[parent]
        done[child] = True
    omega = so3_log(R)
    omega[~done] = np.nan
    return omega


# --------------------------------------------------------------------------- 1DSfM / covariance formats
def egs_to_rotation_2(edge_R_rowmajor):
    """T/io/read_1dsfm.cc:309-325: rotation = S * R^T * S, S = diag(1,-1,-1); angle-axis."""
    R = np.asarray(edge_R_rowmajor, dtype=np.float64).reshape(-1, 3, 3)
    S = np.diag([1.0, -1.0, -1.0])
    return so3_log(S @ np.transpose(R, (0, 2, 1)) @ S)


def parse_covariance_text(path):
    """src/uncertainty.cpp:200-229: two header lines, then `id1 id2` + 9 doubles bit-cast to
    uint64 decimal text (C00 C11 C22 C01 C02 C12 R0 R1 R2).  Returns ids [C,2], cov6 [C,6], rot [C,3]."""
    rows = [l.split() for l in open(path).read().splitlines()[2:] if l.strip()]
    ids = np.array([[int(r[0]), int(r[1])] for r in rows], dtype=np.int64)
    bits = np.array([[int(x) for x in r[2:11]] for r in rows], dtype=np.uint64)
    vals = bits.view(np.float64)
    return ids, np.ascontiguousarray(vals[:, :6]), np.ascontiguousarray(vals[:, 6:9])


def write_covariance_text(path, ids, cov6, rot):
    """src/uncertainty.cpp:164-198 (store_covariance_rot) file layout."""
    vals = np.concatenate([np.asarray(cov6, dtype=np.float64), np.asarray(rot, dtype=np.float64)], axis=1)
    bits = np.ascontiguousarray(vals).view(np.uint64)
    with open(path, "w") as f:
        f.write(f"{len(ids)}\n")
        f.write("view_id1 view_id2 C00 C11 C22 C01 C02 C12 R0 R1 R2\n")
        for (a, b), row in zip(np.asarray(ids).tolist(), bits.tolist()):
            f.write(f"{a} {b} " + " ".join(str(x) for x in row) + "\n")


class PoseGraph:
    """A dense-renumbered rotation-averaging problem: what the C-ABI consumes."""

    def __init__(self, view_ids, edge_i, edge_j, omega_ij, cov6=None, edge_weight=None, omega_init=None,
                 omega_gt=None, num_verified_matches=None, name=""):
        self.view_ids = np.asarray(view_ids)
        self.edge_i = np.ascontiguousarray(edge_i, dtype=np.uint32)
        self.edge_j = np.ascontiguousarray(edge_j, dtype=np.uint32)
        self.omega_ij = np.ascontiguousarray(omega_ij, dtype=np.float64)
        self.cov6 = None if cov6 is None else np.ascontiguousarray(cov6, dtype=np.float64)
        self.edge_weight = None if edge_weight is None else np.ascontiguousarray(edge_weight, dtype=np.float64)
        self.omega_init = omega_init
        self.omega_gt = omega_gt
        self.num_verified_matches = num_verified_matches
        self.name = name

    @property
    def num_views(self):
        return len(self.view_ids)

    @property
    def num_edges(self):
        return len(self.edge_i)


def load_madrid_fixture(npz_path, min_num_two_view_inliers=30):
    """The shipped 1DSfM Madrid_Metropolis dataset exactly as scripts/sfm_pipeline.py hands it to
    the solver: Read1DSFM conventions, FilterInitialViewGraph, covariance lookup (edges without a
    covariance are skipped, rotation_estimator.cpp:239-247), MST initialisation.
    Expected: 379 views / 18 811 edges."""
    z = np.load(npz_path)
    edge_ij = z["edge_ij"].astype(np.int64)
    nvm = z["num_verified_matches"]
    keep, ids = filter_initial_view_graph(z["cc"], edge_ij, nvm, min_num_two_view_inliers)
    cov_key = {(int(a), int(b)): k for k, (a, b) in enumerate(z["cov_ij"].tolist())}
    has_cov = np.array([(int(a), int(b)) in cov_key for a, b in edge_ij.tolist()])
    sel = np.nonzero(keep & has_cov)[0]
    cov_idx = np.array([cov_key[(int(a), int(b))] for a, b in edge_ij[sel].tolist()])
    dense = np.searchsorted(ids, edge_ij[sel])
    assert np.all(dense[:, 0] < dense[:, 1])
    omega_ij = egs_to_rotation_2(z["edge_R"][sel])
    # the MST initialisation runs on the filtered view graph (all kept edges, with or without covariance)
    all_sel = np.nonzero(keep)[0]
    dense_all = np.searchsorted(ids, edge_ij[all_sel])
    omega0 = max_spanning_tree_orientations(len(ids), dense_all[:, 0], dense_all[:, 1],
                                            egs_to_rotation_2(z["edge_R"][all_sel]), nvm[all_sel])
    return PoseGraph(ids, dense[:, 0], dense[:, 1], omega_ij, cov6=z["cov6"][cov_idx], omega_init=omega0,
                     num_verified_matches=nvm[sel], name="madrid_metropolis")


# --------------------------------------------------------------------------- synthetic pose graphs
def madrid_like_covariances(rng, E):
    """Per-edge covariance Sigma = Q diag(lambda) Q^T * 1e-8 with sorted log10(lambda) fitted to the
    Madrid_Metropolis percentiles of eig(1e8 Sigma)."""
    mu = np.array([-0.10, 1.47, 3.00])
    sd = np.array([0.73, 1.16, 1.12])
    lam = 10.0 ** np.clip(np.sort(rng.normal(mu, sd, size=(E, 3)), axis=1), -3, 9)
    Q = so3_exp(random_rotation_vectors(rng, E))
    S = np.einsum("eij,ej,ekj->eik", Q, lam * 1e-8, Q)
    # The reference whitens with a cofactor inverse + LLT and no PD check (rotation_estimator.cpp:252-255); at condition
    # numbers near 1e12 roughly one sample in a few million cancels to a negative pivot (NaN for the reference too).
    # Such samples are redrawn with the spread capped at 1e6 so a synthetic graph never carries an input the reference
    # itself could not digest.
    bad = ~whitening_is_finite(S)
    if bad.any():
        lam_b = lam[bad]
        lam_b = np.maximum(lam_b, lam_b[:, 2:3] * 1e-6)
        S[bad] = np.einsum("eij,ej,ekj->eik", Q[bad], lam_b * 1e-8, Q[bad])
    cov6 = np.stack([S[:, 0, 0], S[:, 1, 1], S[:, 2, 2], S[:, 0, 1], S[:, 0, 2], S[:, 1, 2]], axis=1)
    return cov6, S


def whitening_is_finite(S):
    """The whitening formula of rotation_estimator.cpp:252-255 (cofactor inverse of 1e8 Sigma, then Cholesky), vectorised,
    only to tell whether it stays finite for each covariance [E,3,3]."""
    with np.errstate(all="ignore"):
        a, d, f = S[:, 0, 0] * 1e8, S[:, 1, 1] * 1e8, S[:, 2, 2] * 1e8
        b, c, e = S[:, 0, 1] * 1e8, S[:, 0, 2] * 1e8, S[:, 1, 2] * 1e8
        c00, c01, c02 = d * f - e * e, c * e - b * f, b * e - c * d
        c11, c12, c22 = a * f - c * c, b * c - a * e, a * d - b * b
        idet = 1.0 / (a * c00 + b * c01 + c * c02)
        P00, P10, P20, P11, P21, P22 = c00 * idet, c01 * idet, c02 * idet, c11 * idet, c12 * idet, c22 * idet
        l00 = np.sqrt(P00); l10 = P10 / l00; l20 = P20 / l00
        l11 = np.sqrt(P11 - l10 * l10); l21 = (P21 - l20 * l10) / l11
        l22 = np.sqrt(P22 - l20 * l20 - l21 * l21)
        ok = np.isfinite(l00) & np.isfinite(l11) & np.isfinite(l22) & np.isfinite(l10) & np.isfinite(l20) & np.isfinite(l21)
        # keep a safety margin: pivots that survive only by a few ulps are treated as failures too
        ok &= (P11 - l10 * l10 > 1e-9 * np.abs(P11)) & (P22 - l20 * l20 - l21 * l21 > 1e-9 * np.abs(P22))
    return ok


def synthetic_pose_graph(num_views, num_edges, seed=56, noise_deg=1.0, outlier_fraction=0.1,
                         rotation_scale=0.2, covariance=False, init="chain", name="synthetic"):
    """The fixture pattern of robust_rotation_estimator_test.cc:150-243 scaled up (SURVEY section 8d
    configs 4/5): ground truth omega ~ rotation_scale*U(-1,1)^3; spanning chain (i-1,i) plus uniformly
    random distinct pairs i<j; R_ij = noise * R_j R_i^T (noise: random axis, N(0, noise_deg) angle, or
    drawn from N(0, Sigma) when covariance=True); a fraction of the NON-chain edges is replaced by a
    uniformly random rotation; initialisation by chaining the (noisy) spanning chain."""
    rng = np.random.default_rng(seed)
    N, E = int(num_views), int(num_edges)
    assert E >= N - 1
    gt = rotation_scale * rng.uniform(-1, 1, size=(N, 3))
    chain = np.stack([np.arange(N - 1), np.arange(1, N)], axis=1)
    extra = np.zeros((0, 2), dtype=np.int64)
    need = E - (N - 1)
    max_extra = N * (N - 1) // 2 - (N - 1)
    assert need <= max_extra, "more edges than distinct pairs"
    seen = None
    while len(extra) < need:
        m = int((need - len(extra)) * 1.3) + 16
        a = rng.integers(0, N, size=m)
        b = rng.integers(0, N, size=m)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        ok = (hi - lo) >= 2  # not a self loop, not a chain edge
        key = lo[ok].astype(np.int64) * N + hi[ok]
        key = key[np.sort(np.unique(key, return_index=True)[1])]
        if seen is not None:
            key = key[~np.isin(key, seen)]
        seen = key if seen is None else np.concatenate([seen, key])
        extra = np.stack([seen // N, seen % N], axis=1)
    extra = extra[:need]
    ij = np.concatenate([chain, extra], axis=0)
    order = np.lexsort((ij[:, 1], ij[:, 0]))
    is_chain = np.concatenate([np.ones(N - 1, bool), np.zeros(need, bool)])[order]
    ij = ij[order]
    Rgt = so3_exp(gt)
    Rrel = Rgt[ij[:, 1]] @ np.transpose(Rgt[ij[:, 0]], (0, 2, 1))
    cov6 = None
    if covariance:
        cov6, S = madrid_like_covariances(rng, E)
        L = np.linalg.cholesky(S)
        noise = np.einsum("eij,ej->ei", L, rng.normal(size=(E, 3)))
    else:
        axis = rng.normal(size=(E, 3))
        axis /= np.linalg.norm(axis, axis=1, keepdims=True)
        noise = axis * (np.deg2rad(noise_deg) * rng.normal(size=(E, 1)))
    Rrel = so3_exp(noise) @ Rrel
    outlier = (rng.uniform(size=E) < outlier_fraction) & ~is_chain
    n_out = int(outlier.sum())
    if n_out:
        Rrel[outlier] = so3_exp(random_rotation_vectors(rng, n_out))
    omega_ij = so3_log(Rrel)
    if init == "chain":
        omega0 = np.zeros((N, 3))
        ck = np.nonzero(is_chain)[0]
        ck = ck[np.argsort(ij[ck, 0])]
        R = np.eye(3)
        for k in ck.tolist():  # edge (a, a+1): R_{a+1} = R_rel R_a
            R = Rrel[k] @ R
            omega0[ij[k, 1]] = so3_log(R)
    elif init == "bfs":
        omega0 = spanning_tree_orientations_bfs(N, ij[:, 0], ij[:, 1], Rrel)
    elif init == "gt_perturbed":
        omega0 = so3_log(so3_exp(np.deg2rad(5.0) * rng.normal(size=(N, 3))) @ Rgt)
    else:
        raise ValueError(init)
    g = PoseGraph(np.arange(N), ij[:, 0], ij[:, 1], omega_ij, cov6=cov6, omega_init=omega0, omega_gt=gt, name=name)
    g.is_outlier = outlier
    return g


# --------------------------------------------------------------------------- parity metric
def align_rotations(omega_ref, omega):
    """Gauge alignment: the rotation G minimising sum ||R_ref_i - R_i G||_F (closed-form chordal
    mean; src/compare_reconstructions.cpp:149-177 does the same job with a robust Ceres fit).
    Returns omega' with Exp(omega'_i) = Exp(omega_i) G."""
    Rr, R = so3_exp(omega_ref), so3_exp(omega)
    M = np.einsum("nji,njk->ik", R, Rr)  # sum R_i^T Rref_i
    U, _, Vt = np.linalg.svd(M)
    G = U @ np.diag([1, 1, np.linalg.det(U @ Vt)]) @ Vt
    return so3_log(R @ G)


def align_rotations_robust(omega_ref, omega, loss_width=0.1, max_iterations=500):
    """AlignRotations of the reference (src/compare_reconstructions.cpp:149-177): G = argmin sum rho(|ref_i - Log(R_i Exp(g))|^2),
    rho = ceres::CauchyLoss(0.1), g an angle-axis vector started at 0, Levenberg-Marquardt with IRLS weights (three
    unknowns; forward-difference Jacobian).  Returns omega' with Exp(omega'_i) = Exp(omega_i) Exp(g).  Unlike the chordal
    alignment above this one is NOT invariant to the branch of the angle-axis vectors, exactly like the reference's."""
    ref = np.asarray(omega_ref, dtype=np.float64).reshape(-1, 3)
    R = so3_exp(np.asarray(omega, dtype=np.float64).reshape(-1, 3))
    b = loss_width * loss_width

    def resid(g):
        return (ref - so3_log(R @ so3_exp(g))).ravel()

    def cost(r):
        s = (r.reshape(-1, 3) ** 2).sum(axis=1)
        return 0.5 * float(np.sum(b * np.log1p(s / b)))
    g = np.zeros(3)
    r = resid(g)
    c = cost(r)
    lam = 1e-4
    for _ in range(max_iterations):
        J = np.empty((len(r), 3))
        for k in range(3):
            d = np.zeros(3); d[k] = 1e-7
            J[:, k] = (resid(g + d) - r) / 1e-7
        w = np.repeat(1.0 / (1.0 + (r.reshape(-1, 3) ** 2).sum(axis=1) / b), 3)     # rho'
        H = J.T @ (w[:, None] * J)
        grad = J.T @ (w * r)
        if np.abs(grad).max() < 1e-12:
            break
        step = np.linalg.solve(H + lam * np.diag(np.maximum(np.diag(H), 1e-12)), -grad)
        r2 = resid(g + step)
        c2 = cost(r2)
        if c2 < c:
            g, r, lam = g + step, r2, max(lam / 3.0, 1e-12)
            if c - c2 <= 1e-16 * c or np.linalg.norm(step) < 1e-14:
                c = c2
                break
            c = c2
        else:
            lam *= 4.0
            if lam > 1e12:
                break
    return so3_log(R @ so3_exp(g))


def angular_difference(omega_a, omega_b):
    """AngularDifference (src/compare_reconstructions.cpp:7-16): angle of R_a^T R_b, radians."""
    Ra, Rb = so3_exp(omega_a), so3_exp(omega_b)
    tr = np.einsum("nij,nij->n", Ra, Rb)
    return np.arccos(np.clip((tr - 1.0) / 2.0, -1.0, 1.0))


def mean_angular_error(omega_ref, omega):
    """Mean per-view angular difference (rad) after gauge alignment; small angles are taken from the
    rotation-vector norm (arccos loses half the digits below 1e-8)."""
    al = align_rotations(omega_ref, omega)
    d = so3_log(np.transpose(so3_exp(omega_ref), (0, 2, 1)) @ so3_exp(al))
    ang = np.linalg.norm(d, axis=1)
    return float(ang.mean()), float(ang.max())
