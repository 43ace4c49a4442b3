"""GPU parity of the view-graph shaping kernels (the steps immediately before the solve) against the
host restatements in globalsfmpy_b200/viewgraph.py of
  FilterInitialViewGraph                 reference src/GSfM_global_reconstruction_estimator.cpp:369-390
  OrientationsFromMaximumSpanningTree    T/sfm/view_graph/orientations_from_maximum_spanning_tree.cc:109-178
Integer results (masks, the tree) must be identical; orientations agree to rounding (quaternion chain vs matrix chain)."""
import numpy as np
import pytest

from globalsfmpy_b200 import _capi as capi, solver, viewgraph as vg

pytestmark = pytest.mark.gpu


def _random_graph(rng, n, e, components=1):
    """Distinct undirected pairs i<j inside `components` disjoint index blocks + a few isolated views."""
    bounds = np.linspace(0, n - 3, components + 1).astype(int)   # the last 3 views stay isolated
    pairs = set()
    for c in range(components):
        lo, hi = bounds[c], bounds[c + 1]
        for v in range(lo + 1, hi):                               # a chain keeps the block connected
            pairs.add((v - 1, v))
        want = len(pairs) + e // components
        while len(pairs) < want and hi - lo > 2:
            a, b = rng.integers(lo, hi, 2)
            if a != b:
                pairs.add((min(a, b), max(a, b)))
    p = np.array(sorted(pairs), dtype=np.uint32)
    p = p[rng.permutation(len(p))]
    flip = rng.random(len(p)) < 0.3                                # some edges stored as (larger, smaller)
    return np.where(flip, p[:, 1], p[:, 0]), np.where(flip, p[:, 0], p[:, 1])


@pytest.mark.parametrize("n,e,comps,seed", [(40, 100, 1, 0), (300, 1500, 3, 1), (2000, 12000, 5, 2), (5000, 60000, 1, 3)])
def test_filter_initial_view_graph_matches_host(n, e, comps, seed):
    rng = np.random.default_rng(seed)
    ei, ej = _random_graph(rng, n, e, comps)
    matches = rng.integers(5, 120, len(ei))
    keep_d, views_d = solver.filter_initial_view_graph(n, ei, ej, matches, 30)
    keep_h, ids_h = vg.filter_initial_view_graph(np.arange(n), np.stack([ei, ej], 1), matches, 30)
    assert np.array_equal(keep_d, keep_h)
    assert np.array_equal(np.nonzero(views_d)[0], ids_h)
    assert keep_d.sum() > 0


def test_filter_ties_and_degenerate_inputs():
    # two components of equal size: the one holding the smallest view index wins; all edges below the threshold: nothing kept
    ei = np.array([4, 5, 0, 1], np.uint32)
    ej = np.array([5, 6, 1, 2], np.uint32)
    keep, views = solver.filter_initial_view_graph(8, ei, ej, [50, 50, 50, 50], 30)
    assert keep.tolist() == [False, False, True, True] and np.nonzero(views)[0].tolist() == [0, 1, 2]
    keep, views = solver.filter_initial_view_graph(8, ei, ej, [1, 2, 3, 4], 30)
    assert not keep.any() and not views.any()
    with pytest.raises(capi.GsfmError):
        solver.filter_initial_view_graph(3, [0], [7], [50], 30)


@pytest.mark.parametrize("n,e,seed,wmax", [(30, 80, 0, 4), (400, 3000, 1, 10), (3000, 40000, 2, 200)])
def test_mst_orientations_match_host(n, e, seed, wmax):
    """Small integer weights => many ties: the (weight desc, i, j) order makes the tree unique, and the device must
    build exactly the host Kruskal's tree."""
    rng = np.random.default_rng(seed)
    ei, ej = _random_graph(rng, n, e, 1)
    E = len(ei)
    w = rng.integers(1, wmax + 1, E)
    gt = vg.random_rotation_vectors(rng, n)
    R = vg.so3_exp(gt)
    lo, hi = np.minimum(ei, ej), np.maximum(ei, ej)
    # rotation_2 of a pair is stored for the (smaller, larger) orientation whatever the order the ids are listed in
    wij = vg.so3_log(R[hi] @ np.transpose(R[lo], (0, 2, 1)))
    om_d, tree_d, rounds = solver.init_orientations_mst(n, ei, ej, wij, w)
    om_h = vg.max_spanning_tree_orientations(n, ei, ej, wij, w)
    reach = ~np.isnan(om_h[:, 0])
    assert np.array_equal(reach, ~np.isnan(om_d[:, 0]))
    assert tree_d.sum() == reach.sum() - 1 and 1 <= rounds <= int(np.ceil(np.log2(n))) + 1
    # same tree: total weight and, edge by edge, the host Kruskal (re-run here to expose its edge set)
    order = np.lexsort((ej.astype(np.int64), ei.astype(np.int64), -w.astype(np.int64)))
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    tree_h = np.zeros(E, bool)
    for k in order.tolist():
        ra, rb = find(int(ei[k])), find(int(ej[k]))
        if ra != rb:
            parent[ra] = rb
            tree_h[k] = True
    assert np.array_equal(tree_d, tree_h)
    # noise-free measurements: both initialisations reproduce the ground truth up to the gauge (root = identity)
    assert np.abs(vg.so3_exp(om_d[reach]) - vg.so3_exp(om_h[reach])).max() < 1e-9   # quaternion chain vs matrix chain
    root = int(min(ei.min(), ej.min()))
    assert np.allclose(om_d[root], 0.0)
    assert vg.mean_angular_error(gt[reach], om_d[reach])[0] < 1e-9


def test_mst_unreachable_views_and_explicit_root():
    ei = np.array([0, 1, 4], np.uint32)
    ej = np.array([1, 2, 5], np.uint32)
    wij = np.array([[0.1, 0.0, 0.0], [0.0, 0.2, 0.0], [0.0, 0.0, 0.3]])
    om, tree, _ = solver.init_orientations_mst(7, ei, ej, wij, [3, 3, 3])
    assert tree.tolist() == [True, True, True]                      # a spanning FOREST: every component gets its tree
    assert not np.isnan(om[:3]).any() and np.isnan(om[3:]).all()    # only the root's component is oriented
    assert np.allclose(om[1], [0.1, 0.0, 0.0])
    om5, _, _ = solver.init_orientations_mst(7, ei, ej, wij, [3, 3, 3], root=5)
    assert np.isnan(om5[:4]).all() and np.allclose(om5[5], 0.0) and np.allclose(om5[4], [0.0, 0.0, -0.3]) and np.isnan(om5[6]).all()


def test_mst_init_feeds_the_solver():
    """End to end on the device: filter -> spanning-tree initialisation -> robust solve."""
    g = vg.synthetic_pose_graph(300, 6000, seed=9, noise_deg=1.0, outlier_fraction=0.1)
    rng = np.random.default_rng(9)
    matches = rng.integers(10, 200, g.num_edges)
    keep, views = solver.filter_initial_view_graph(g.num_views, g.edge_i, g.edge_j, matches, 30)
    assert views.all()
    ei, ej, wij = g.edge_i[keep], g.edge_j[keep], g.omega_ij[keep]
    om0, tree, _ = solver.init_orientations_mst(g.num_views, ei, ej, wij, matches[keep])
    assert not np.isnan(om0).any() and tree.sum() == g.num_views - 1
    prob = capi.ProblemArrays(g.num_views, ei, ej, wij, error_type=capi.ANGLE_AXIS)
    o = capi.default_options_py()
    o.loss = capi.Loss.make(capi.LOSS_CAUCHY, 0.05)
    om, s, _ = solver.solve(prob, o, om0)
    assert s.final_cost < s.initial_cost
    assert np.degrees(vg.mean_angular_error(g.omega_gt, om)[0]) < 1.0
