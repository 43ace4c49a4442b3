"""`GlobalSfMpy`-compatible module: the rotation-averaging slice of the reference's pybind11 surface
(bind_src/GlobalSfMpy.cpp), backed by the B200 solver through the C ABI.

    sys.path.append('<repo>/globalsfmpy_b200/compat')     # where the reference scripts append '../build'
    import GlobalSfMpy as sfm

Everything `scripts/sfm_pipeline.py` touches through step 3 (global rotations) is here with the reference's names
and argument meaning: options + `load_1DSFM_config`, `Reconstruction`, `ViewGraph`,
`MapEdgesCovariance`, `Read1DSFM`, `ReadCovariance`, `ReconstructionBuilder.CheckView/get_view_graph/get_reconstruction`,
`GlobalReconstructionEstimator.FilterInitialViewGraphAndCalibrateCameras / EstimateGlobalRotationsUncertainty /
EstimateGlobalRotations / OrientationsFromMaximumSpanningTree / FilterRotations / orientations`, `SetOrientations`,
`LossFunction`, `RotationEstimator`, `NonlinearRotationEstimator`, `RotationErrorType`, the MAGSAC gamma constants.
Step 7 (`EstimatePosition`: robust translation averaging, src/GSfM_nonlinear_position_estimator.cpp) runs on the same device
solver through include/gsfm_pa.h and fills `.positions`.  The remaining steps (pairwise-translation refinement, triangulation,
bundle adjustment, file writers) are the rest of TheiaSfM and are out of scope: they raise NotImplementedError naming the
reference function.
"""
import enum
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from globalsfmpy_b200 import _capi as _capi  # noqa: E402
from globalsfmpy_b200 import solver as _solver  # noqa: E402
from globalsfmpy_b200 import viewgraph as _vg  # noqa: E402
from globalsfmpy_b200.loss_functions import LossFunction  # noqa: E402,F401
from globalsfmpy_b200.losses import UnsupportedLoss, loss_to_struct  # noqa: E402,F401


# ------------------------------------------------------------------------------- enums, containers
class RotationErrorType(enum.IntEnum):          # include/pairwise_rotation_error_quat.hpp:50-61, bind:431-439
    QUATERNION_NORM = 0
    ROTATION_MAT_FNORM = 1
    QUATERNION_COSINE = 2
    ANGLE_AXIS_COVARIANCE = 3
    ANGLE_AXIS = 4
    ANGLE_AXIS_INLIERS = 5
    ANGLE_AXIS_COV_INLIERS = 6
    ANGLE_AXIS_COVTRACE = 7
    ANGLE_AXIS_COVNORM = 8


class PositionErrorType(enum.IntEnum):           # include/pairwise_translation_error_covariance.hpp:47-51; bind:441-443 exports BASELINE only
    BASELINE = 0


class MapViewIdVector3d(dict):
    """std::unordered_map<ViewId, Eigen::Vector3d> (bind:153)."""


class MapEdges(dict):
    """std::unordered_map<ViewIdPair, TwoViewInfo> (bind:154)."""


class MapEdgesCovariance(dict):
    """std::unordered_map<ViewIdPair, pair<Matrix3d, Vector3d>> (bind:160, include/uncertainty.hpp:13)."""


class TwoViewInfo:                               # T/sfm/twoview_info.h:54-98
    def __init__(self):
        self.focal_length_1 = 0.0
        self.focal_length_2 = 0.0
        self.position_2 = np.zeros(3)
        self.rotation_2 = np.zeros(3)
        self.num_verified_matches = 0
        self.num_homography_inliers = 0
        self.visibility_score = 1


class ViewGraph:                                 # T/sfm/view_graph/view_graph.{h,cc}
    def __init__(self):
        self._adj = {}
        self._edges = MapEdges()

    def NumViews(self):
        return len(self._adj)

    def NumEdges(self):
        return len(self._edges)

    def HasView(self, view_id):
        return view_id in self._adj

    def HasEdge(self, a, b):
        return ((a, b) if a < b else (b, a)) in self._edges

    def ViewIds(self):
        return set(self._adj)

    def GetAllEdges(self):
        return self._edges

    def GetEdge(self, a, b):
        return self._edges.get((a, b) if a < b else (b, a))

    def GetNeighborIdsForView(self, view_id):
        return self._adj.get(view_id)

    def AddEdge(self, a, b, info):               # key normalised to (min,max); the info is stored as given (view_graph.cc:133-151)
        if a == b:
            return
        self._adj.setdefault(a, set()).add(b)
        self._adj.setdefault(b, set()).add(a)
        self._edges[(a, b) if a < b else (b, a)] = info

    def RemoveEdge(self, a, b):
        key = (a, b) if a < b else (b, a)
        if key not in self._edges:
            return False
        del self._edges[key]
        self._adj[a].discard(b)
        self._adj[b].discard(a)
        return True

    def RemoveView(self, view_id):
        if view_id not in self._adj:
            return False
        for n in list(self._adj[view_id]):
            self.RemoveEdge(view_id, n)
        del self._adj[view_id]
        return True


class _View:
    def __init__(self, name):
        self.name = name
        self.estimated = False
        self.orientation = np.zeros(3)
        self.focal_length_prior = None

    def Name(self):
        return self.name

    def IsEstimated(self):
        return self.estimated

    def GetOrientationAsAngleAxis(self):
        return self.orientation


class Reconstruction:
    """The slice of theia::Reconstruction the rotation stage touches: views by id, their names, estimated flags."""

    def __init__(self):
        self._views = {}
        self._next_id = 0
        self._view_tracks = {}   # view id -> set of track ids (common-track counts, read_1dsfm.cc:354-360)
        self._num_tracks = 0

    def AddView(self, name):
        vid = self._next_id
        self._next_id += 1
        self._views[vid] = _View(name)
        return vid

    def RemoveView(self, view_id):
        return self._views.pop(view_id, None) is not None

    def NumViews(self):
        return len(self._views)

    def NumTracks(self):
        return self._num_tracks

    def ViewIds(self):
        return list(self._views)

    def View(self, view_id):
        return self._views.get(view_id)

    def MutableView(self, view_id):
        return self._views.get(view_id)


class ReconstructionEstimatorOptions:            # T/sfm/reconstruction_estimator_options.h (fields the path reads)
    def __init__(self):
        self.num_threads = 1
        self.min_num_two_view_inliers = 30
        self.rotation_filtering_max_difference_degrees = 5.0
        self.global_rotation_estimator_type = "ROBUST_L1L2"
        self.global_position_estimator_type = "NONLINEAR"
        self.reconstruction_estimator_type = "GLOBAL"
        self.num_retriangulation_iterations = 1
        self.refine_camera_positions_and_points_after_position_estimation = True
        self.extra = {}


class ReconstructionBuilderOptions:
    def __init__(self):
        self.num_threads = 1
        self.min_track_length = 2
        self.max_track_length = 50
        self.reconstruct_largest_connected_component = False
        self.only_calibrated_views = False
        self.reconstruction_estimator_options = ReconstructionEstimatorOptions()

    def print(self):
        print({k: v for k, v in vars(self).items() if k != "reconstruction_estimator_options"},
              vars(self.reconstruction_estimator_options))


def load_1DSFM_config(flagfile, options):
    """bind_src/GlobalSfMpy.cpp:185-271: yaml -> options (the keys the rotation stage reads are typed; the
    rest is kept verbatim in reconstruction_estimator_options.extra)."""
    import yaml
    cfg = yaml.safe_load(open(flagfile))
    options.num_threads = int(cfg["num_threads"])
    options.min_track_length = int(cfg["min_track_length"])
    options.max_track_length = int(cfg["max_track_length"])
    options.reconstruct_largest_connected_component = bool(cfg["reconstruct_largest_connected_component"])
    options.only_calibrated_views = bool(cfg["only_calibrated_views"])
    r = options.reconstruction_estimator_options
    r.min_num_two_view_inliers = int(cfg["min_num_inliers_for_valid_match"])
    r.num_threads = int(cfg["num_threads"])
    r.reconstruction_estimator_type = str(cfg["reconstruction_estimator"])
    r.global_rotation_estimator_type = str(cfg["global_rotation_estimator"])
    r.global_position_estimator_type = str(cfg["global_position_estimator"])
    r.rotation_filtering_max_difference_degrees = float(cfg["post_rotation_filtering_degrees"])
    r.num_retriangulation_iterations = int(cfg["num_retriangulation_iterations"])
    r.refine_camera_positions_and_points_after_position_estimation = bool(
        cfg["refine_camera_positions_and_points_after_position_estimation"])
    r.extra = dict(cfg)


# ------------------------------------------------------------------------------- dataset readers
def ReadCovariance(dataset_directory, rot_covariances):
    """src/uncertainty.cpp:200-229 -> {(id1,id2): (3x3 covariance, refined rotation)}."""
    # the native reader of libgsfm_ra.so (csrc/gsfm_io.cpp); viewgraph.parse_covariance_text is its Python twin
    ids, cov6, rot = _solver.read_covariance_rot(os.path.join(dataset_directory, "covariance_rot.txt"))
    for (a, b), c, r in zip(ids.tolist(), cov6, rot):
        S = np.array([[c[0], c[3], c[4]], [c[3], c[1], c[5]], [c[4], c[5], c[2]]])
        rot_covariances[(int(a), int(b))] = (S, np.array(r))
    return True


def Read1DSFM(dataset_directory, reconstruction, view_graph, rot_covariances=None):
    """T/io/read_1dsfm.cc:375-412 (cc -> list -> tracks -> EGs) + read_covariance (bind:611-617)."""
    d = dataset_directory
    cc = set(int(t) for t in open(os.path.join(d, "cc.txt")).read().split())
    for line in open(os.path.join(d, "list.txt")):
        tok = line.split()
        if not tok:
            continue
        vid = reconstruction.AddView(os.path.splitext(os.path.basename(tok[0]))[0])
        if vid not in cc:
            reconstruction.RemoveView(vid)        # ids stay in sync with the line index (read_1dsfm.cc:139-146)
            continue
        if len(tok) >= 3 and float(tok[2]) != 0:
            reconstruction.View(vid).focal_length_prior = float(tok[2])
    tok = open(os.path.join(d, "tracks.txt")).read().split()
    pos, n_tracks = 1, int(tok[0])
    for t in range(n_tracks):
        n = int(tok[pos]); pos += 1
        views = [int(v) for v in tok[pos:pos + 2 * n:2]]
        pos += 2 * n
        if len(set(views)) != len(views):
            continue                               # Reconstruction::AddTrack rejects a track seeing one view twice
        for v in views:
            reconstruction._view_tracks.setdefault(v, set()).add(t)
        reconstruction._num_tracks += 1
    eg = np.loadtxt(os.path.join(d, "EGs.txt"), dtype=np.float64, ndmin=2)
    rot2 = _vg.egs_to_rotation_2(eg[:, 2:11])
    S = np.array([1.0, -1.0, -1.0])
    empty = set()
    for k in range(len(eg)):
        a, b = int(eg[k, 0]), int(eg[k, 1])
        if reconstruction.View(a) is None or reconstruction.View(b) is None:
            continue
        info = TwoViewInfo()
        info.rotation_2 = rot2[k].copy()
        info.position_2 = S * eg[k, 11:14]
        common = len(reconstruction._view_tracks.get(a, empty) & reconstruction._view_tracks.get(b, empty))
        info.num_verified_matches = common
        info.visibility_score = common
        view_graph.AddEdge(a, b, info)
    if rot_covariances is not None:
        ReadCovariance(d, rot_covariances)
    return True


def CalcCovariance(dataset_path):
    raise NotImplementedError("CalcCovariance (src/uncertainty.cpp:82-198) needs Ceres' covariance estimation over the matched "
                              "features: an offline pre-step outside the rotation-averaging path; use the shipped covariance_rot.txt")


# ------------------------------------------------------------------------------- solver front-ends
def _copy_options(options):
    """A private copy of a _capi.Options (or the defaults): the caller's object is never written to."""
    o = _capi.Options()
    src = options if options is not None else _capi.default_options_py()
    import ctypes
    ctypes.memmove(ctypes.byref(o), ctypes.byref(src), ctypes.sizeof(_capi.Options))
    if options is None:
        o.n_gpus = -1     # behind the module API a large view graph shards over the box's GPUs by itself
    return o


def _matched_features(reconstruction, a, b):
    """get_matched_features(view_id_pair, reconstruction, features).first.size() (rotation_estimator.cpp:260-274): the
    features of the two views that belong to a common track."""
    ta, tb = reconstruction._view_tracks.get(a), reconstruction._view_tracks.get(b)
    return len(ta & tb) if ta and tb else 0


def _flatten(view_pairs, orientations, covariances, error_type, reconstruction=None):
    """What rotation_estimator.cpp:228-293 does with the hash maps: skip edges whose endpoints have no initial
    orientation or (covariance types) no covariance; dense-renumber the views; per-edge weight #matched/100 for the
    *_INLIERS types (:260-274)."""
    ids = np.array(sorted(orientations), dtype=np.int64)
    dense = {int(v): k for k, v in enumerate(ids.tolist())}
    needs_cov = error_type in (3, 6, 7, 8)
    needs_matches = error_type in (5, 6)
    if needs_matches and reconstruction is None:
        raise ValueError("ANGLE_AXIS_INLIERS / ANGLE_AXIS_COV_INLIERS weight every edge by its matched features: pass the reconstruction")
    ei, ej, wij, cov6, weight = [], [], [], [], []
    for (a, b), info in view_pairs.items():
        if a not in dense or b not in dense:
            continue
        if needs_cov:
            c = covariances.get((a, b)) if covariances is not None else None
            if c is None:
                continue
            S = c[0]
            cov6.append([S[0][0], S[1][1], S[2][2], S[0][1], S[0][2], S[1][2]])
        ei.append(dense[a]); ej.append(dense[b]); wij.append(np.asarray(info.rotation_2, dtype=np.float64))
        if needs_matches:
            weight.append(_matched_features(reconstruction, a, b) / 100.0)
    omega = np.array([np.asarray(orientations[int(v)], dtype=np.float64) for v in ids.tolist()]).reshape(len(ids), 3)
    return ids, np.array(ei, np.uint32), np.array(ej, np.uint32), np.array(wij).reshape(len(ei), 3), \
        (np.array(cov6).reshape(len(ei), 6) if needs_cov else None), omega, (np.array(weight) if needs_matches else None)


def _solve(view_pairs, orientations, loss, error_type, covariances=None, num_threads=1, options=None, reconstruction=None):
    error_type = int(error_type)
    if len(orientations) == 0 or len(view_pairs) == 0:
        return False                                  # rotation_estimator.cpp:209-220
    ids, ei, ej, wij, cov6, omega, weight = _flatten(view_pairs, orientations, covariances, error_type, reconstruction)
    if len(ei) == 0:
        return True
    prob = _capi.ProblemArrays(len(ids), ei, ej, wij, cov6=cov6, edge_weight=weight, error_type=error_type)
    o = _copy_options(options)
    L = loss_to_struct(loss)                          # keeps a tabulated loss's host table alive until the call returns
    o.loss = L
    o.num_threads = int(num_threads)
    omega, summary, _ = _solver.solve(prob, o, omega)
    for k, v in enumerate(ids.tolist()):
        orientations[int(v)] = omega[k].copy()
    _solve.last_summary = summary
    return True                                       # the reference returns true whatever Ceres reports (:308)


class RotationEstimator:
    """theia::RotationEstimator (T/sfm/global_pose_estimation/rotation_estimator.h:50-66); subclassable."""

    def EstimateRotations(self, view_pairs, rotations):
        raise NotImplementedError("pure virtual")


class NonlinearRotationEstimator(RotationEstimator):
    """GSfMNonlinearRotationEstimator (include/GSfM_nonlinear_rotation_estimator.hpp:22-59)."""

    def __init__(self, robust_loss_width=0.1):
        self.robust_loss_width_ = robust_loss_width

    def EstimateRotations(self, view_pairs, rotations):
        """rotation_estimator.cpp:24-80: SoftLOneLoss(robust_loss_width), PairwiseRotationError weight 1."""
        return _solve(view_pairs, rotations, _capi.Loss.make(_capi.LOSS_SOFTLONE, self.robust_loss_width_),
                      RotationErrorType.ANGLE_AXIS)

    def EstimateRotationsWithCustomizedLoss(self, view_pairs, rotations, loss_function, thread_num=1,
                                            rotation_error_type=RotationErrorType.QUATERNION_COSINE):
        return _solve(view_pairs, rotations, loss_function, rotation_error_type, num_threads=thread_num)

    def EstimateRotationsWithCustomizedLossAndCovariance(self, view_pairs, rotations, loss_function, thread_num, covariances,
                                                         rotation_error_type, reconstruction=None):
        return _solve(view_pairs, rotations, loss_function, rotation_error_type, covariances, thread_num, reconstruction=reconstruction)


class ReconstructionBuilder:
    def __init__(self, options, reconstruction, view_graph=None):
        if view_graph is None:
            raise NotImplementedError("ReconstructionBuilder(options, database): the RocksDB feature/match database path is out of scope")
        self.options_, self.reconstruction_, self.view_graph_ = options, reconstruction, view_graph

    def CheckView(self):
        assert self.view_graph_.NumViews() >= 2, "At least 2 images must be provided in order to create a reconstruction."

    def get_view_graph(self):
        return self.view_graph_

    def get_reconstruction(self):
        return self.reconstruction_


def _on_device():
    """True when a CUDA device is present: the view-graph shaping then runs there as well (the solve always does)."""
    try:
        return _capi.lib().gsfm_ra_device_count() > 0
    except Exception:
        return False


class GlobalReconstructionEstimator:
    """The steppable estimator (src/GSfM_global_reconstruction_estimator.cpp), steps 1-3 and the rotation filter."""

    def __init__(self, options):
        self.options = options
        self.orientations = MapViewIdVector3d()
        self.positions = {}
        self.view_graph_ = None
        self.reconstruction_ = None
        self.solver_options = None       # optional _capi.Options override (PCG tolerances, device, ...)

    def get_view_graph(self):
        return self.view_graph_

    def get_reconstruction(self):
        return self.reconstruction_

    def FilterInitialViewGraphAndCalibrateCameras(self, view_graph, reconstruction):
        """Estimate_BeforeStep3 -> FilterInitialViewGraph (:369-390): drop edges with too few verified matches,
        keep the largest connected component; camera calibration from priors is a no-op for rotations."""
        self.view_graph_, self.reconstruction_ = view_graph, reconstruction
        self.orientations.clear()
        edges = view_graph.GetAllEdges()
        keys = list(edges)
        if not keys:
            return False
        ij = np.array(keys, dtype=np.int64)
        nvm = np.array([edges[k].num_verified_matches for k in keys])
        all_ids = np.unique(np.concatenate([np.array(sorted(view_graph.ViewIds()), dtype=np.int64), ij.ravel()]))
        if _on_device():   # edge filter + largest connected component on the device (gsfm_ra_filter_initial_view_graph)
            dense = np.searchsorted(all_ids, ij)
            keep, vkeep = _solver.filter_initial_view_graph(len(all_ids), dense[:, 0], dense[:, 1], nvm, self.options.min_num_two_view_inliers)
            ids = all_ids[vkeep]
        else:              # no CUDA device (CPU test-suite): the host restatement
            keep, ids = _vg.filter_initial_view_graph(all_ids, ij, nvm, self.options.min_num_two_view_inliers)
        for k, kp in zip(keys, keep.tolist()):
            if not kp:
                view_graph.RemoveEdge(*k)
        alive = set(int(v) for v in ids.tolist())
        for v in list(view_graph.ViewIds()):
            if v not in alive:
                view_graph.RemoveView(v)
        return view_graph.NumEdges() >= 1

    def OrientationsFromMaximumSpanningTree(self):
        edges = self.view_graph_.GetAllEdges()
        keys = sorted(edges)
        ids = np.array(sorted(self.view_graph_.ViewIds()), dtype=np.int64)
        ij = np.searchsorted(ids, np.array(keys, dtype=np.int64))
        w = np.array([edges[k].num_verified_matches for k in keys])
        rot = np.array([edges[k].rotation_2 for k in keys])
        if _on_device():   # Boruvka + tree propagation on the device (gsfm_ra_init_orientations_mst)
            om = _solver.init_orientations_mst(len(ids), ij[:, 0], ij[:, 1], rot, w)[0]
        else:
            om = _vg.max_spanning_tree_orientations(len(ids), ij[:, 0], ij[:, 1], rot, w)
        for k, v in enumerate(ids.tolist()):
            if not np.isnan(om[k, 0]):
                self.orientations[int(v)] = om[k].copy()
        return True

    def EstimateGlobalRotationsUncertainty(self, loss_func, covariances, rotation_error_type):
        """:463-485: MST initialisation, then EstimateRotationsWithCustomizedLossAndCovariance."""
        self.OrientationsFromMaximumSpanningTree()
        return _solve(self.view_graph_.GetAllEdges(), self.orientations, loss_func, rotation_error_type, covariances,
                      self.options.num_threads, self.solver_options, self.reconstruction_)

    def EstimateGlobalRotations(self, loss_func=None, rotation_error_type=RotationErrorType.QUATERNION_COSINE):
        """:440-461 (EstimateGlobalRotationsNonLinear)."""
        self.OrientationsFromMaximumSpanningTree()
        return _solve(self.view_graph_.GetAllEdges(), self.orientations, loss_func, rotation_error_type, None,
                      self.options.num_threads, self.solver_options)

    def EstimateGlobalRotationsWithSigmaConsensus(self, loss_func, iters_num, sigma_max):
        """EstimateGlobalRotationsSigmaConsensus -> EstimateRotationsWithSigmaConsensus (rotation_estimator.cpp:314-457)."""
        self.OrientationsFromMaximumSpanningTree()
        edges = self.view_graph_.GetAllEdges()
        if len(self.orientations) == 0 or len(edges) == 0:
            return False
        ids, ei, ej, wij, _, omega, _ = _flatten(edges, self.orientations, None, 4)
        prob = _capi.ProblemArrays(len(ids), ei, ej, wij, error_type=4)
        prob.c.total_pair_count = len(edges)       # the reference's stop test averages over view_pairs.size(), skipped pairs included
        o = _copy_options(self.solver_options)
        o.n_gpus = 0                               # the re-weighting loop runs on one device
        L = loss_to_struct(loss_func)
        o.loss = L
        omega, summary = _solver.solve_sigma_consensus(prob, o, omega, int(iters_num), float(sigma_max))
        for k, v in enumerate(ids.tolist()):
            self.orientations[int(v)] = omega[k].copy()
        _solve.last_summary = summary
        return True

    def FilterRotations(self):
        """src/GSfM_global_reconstruction_estimator.cpp:509-524: FilterViewPairsFromOrientation with
        options.rotation_filtering_max_difference_degrees (on the device), then RemoveDisconnectedViewPairs -- only the largest
        connected component survives -- and the orientations of the removed views are erased."""
        edges = self.view_graph_.GetAllEdges()
        keys = [k for k in edges if k[0] in self.orientations and k[1] in self.orientations]
        keyset = set(keys)
        missing = [k for k in edges if k not in keyset]
        ids, ei, ej, wij, _, omega, _ = _flatten({k: edges[k] for k in keys}, self.orientations, None, 4)
        prob = _capi.ProblemArrays(len(ids), ei, ej, wij)
        keep, _ = _solver.filter_view_pairs(prob, omega, self.options.rotation_filtering_max_difference_degrees)
        for k, kp in zip(keys, keep.tolist()):
            if not kp:
                self.view_graph_.RemoveEdge(*k)
        for k in missing:
            self.view_graph_.RemoveEdge(*k)
        # RemoveDisconnectedViewPairs (T/sfm/view_graph/remove_disconnected_view_pairs.cc:48-): keep the largest component
        left = list(self.view_graph_.GetAllEdges())
        all_ids = np.array(sorted(self.view_graph_.ViewIds()), dtype=np.int64)
        if left:
            ij = np.searchsorted(all_ids, np.array(left, dtype=np.int64))
            if _on_device():
                _, vkeep = _solver.filter_initial_view_graph(len(all_ids), ij[:, 0], ij[:, 1], np.ones(len(left), np.int32), 0)
            else:
                _, kept_ids = _vg.filter_initial_view_graph(all_ids, np.array(left, dtype=np.int64), np.ones(len(left), np.int64), 0)
                vkeep = np.isin(all_ids, kept_ids)
            alive = set(int(v) for v in all_ids[vkeep].tolist())
        else:
            alive = set()
        for v in all_ids.tolist():
            if int(v) not in alive:
                self.view_graph_.RemoveView(int(v))
                self.orientations.pop(int(v), None)
        for v in [v for v in self.orientations if v not in alive]:
            del self.orientations[v]
        return True

    def EstimatePosition(self, loss_func, error_type=PositionErrorType.BASELINE):
        """bind:564-579 -> EstimatePositionNonLinear (src/GSfM_global_reconstruction_estimator.cpp:605-617) ->
        GSfMNonlinearPositionEstimator::EstimatePositions(view_pairs, orientations_, &positions_, error_type, loss_func)
        (src/GSfM_nonlinear_position_estimator.cpp:151-234) on the device through gsfm_pa_solve: camera-to-camera constraints
        only (position_estimation_min_num_tracks_per_view = 0 in flags_1dsfm.yaml), every camera starts at the origin, one view
        is held constant (the reference: positions->begin() of a hash map; here the smallest estimated view id -- the choice
        moves the solution by a global translation only)."""
        edges = self.view_graph_.GetAllEdges()
        if len(edges) == 0 or len(self.orientations) == 0:
            return False                                  # position_estimator.cpp:159-164
        constrained = set()
        for a, b in edges:
            constrained.add(a); constrained.add(b)
        ids = np.array(sorted(v for v in self.orientations if v in constrained), dtype=np.int64)   # InitializeRandomPositions :236-257
        if len(ids) == 0:
            return False
        dense = {int(v): k for k, v in enumerate(ids.tolist())}
        ei, ej, p2 = [], [], []
        for (a, b), info in edges.items():
            if a in dense and b in dense:                 # :307-312
                ei.append(dense[a]); ej.append(dense[b]); p2.append(np.asarray(info.position_2, dtype=np.float64))
        orient = np.array([np.asarray(self.orientations[int(v)], dtype=np.float64) for v in ids.tolist()]).reshape(len(ids), 3)
        x = np.zeros((len(ids), 3))
        if ei:
            from globalsfmpy_b200 import positions as _pos
            prob = _pos.PositionProblemArrays(len(ids), np.array(ei, np.uint32), np.array(ej, np.uint32), np.array(p2).reshape(len(ei), 3), orient,
                                              fixed_view=0, error_type=int(error_type))
            o = _pos.default_options()                    # Ceres defaults, max_num_iterations 400
            if self.solver_options is not None:
                o = _copy_options(self.solver_options)
            else:
                o.n_gpus = -1
            L = loss_to_struct(loss_func)
            o.loss = L
            o.num_threads = int(self.options.num_threads)
            x, summary, _ = _pos.solve(prob, o)
            _solve.last_summary = summary
            if summary.termination == _capi.TERMINATION_FAILURE:
                return False                              # summary.IsSolutionUsable()
        self.positions = {int(v): x[k].copy() for k, v in enumerate(ids.tolist())}
        return True

    def _todo(name):  # noqa: N805
        def f(self, *a, **k):
            raise NotImplementedError(f"{name}: this pipeline step is TheiaSfM's (out of scope here: SURVEY section 2)")
        return f

    OptimizePairwiseTranslations = _todo("OptimizePairwiseTranslations")
    FilterRelativeTranslation = _todo("FilterRelativeTranslation")
    EstimateStructure = _todo("EstimateStructure")
    BundleAdjustCameraPositionsAndPoints = _todo("BundleAdjustCameraPositionsAndPoints")
    BundleAdjustmentAndRemoveOutlierPoints = _todo("BundleAdjustmentAndRemoveOutlierPoints")


def SetOrientations(orientations, reconstruction):
    """bind_src/GlobalSfMpy.cpp:80-98."""
    for vid in reconstruction.ViewIds():
        reconstruction.MutableView(vid).estimated = False
    for vid, w in orientations.items():
        v = reconstruction.MutableView(vid)
        if v is None:
            continue
        v.orientation = np.asarray(w, dtype=np.float64).copy()
        v.estimated = True


class CompareInfo:                               # include/compare_reconstructions.hpp:24-47, bind:387-393
    def __init__(self):
        self.rotation_diff_when_align = []
        self.position_errors = []
        self.num_3d_points = 0
        self.common_camera = 0
        self.num_reconstructed_view = 0


def AngularDifference(rotation1, rotation2):
    """src/compare_reconstructions.cpp:7-16: angle (rad) of R1^T R2."""
    return float(_vg.angular_difference(np.asarray(rotation1, dtype=np.float64)[None], np.asarray(rotation2, dtype=np.float64)[None])[0])


def AlignRotations(gt_rotation, rotation):
    """src/compare_reconstructions.cpp:149-177: the rotation G (angle-axis, started at 0) minimising
    sum rho(|gt_i - Log(R_i G)|^2) with rho = CauchyLoss(0.1), applied to `rotation` in place (a list / array of angle-axis
    vectors).  Host-side metric code: three unknowns, N residual blocks."""
    out = _vg.align_rotations_robust(np.asarray(gt_rotation, dtype=np.float64), np.asarray(rotation, dtype=np.float64))
    for k in range(len(rotation)):
        rotation[k] = out[k]
    return rotation


def FindCommonEstimatedViewsByName(reconstruction1, reconstruction2):
    """src/compare_reconstructions.cpp:180-195."""
    by_name = {reconstruction2.View(v).Name(): v for v in reconstruction2.ViewIds()}
    names = []
    for v in reconstruction1.ViewIds():
        name = reconstruction1.View(v).Name()
        w = by_name.get(name)
        if w is not None and reconstruction2.View(w).IsEstimated():
            names.append(name)
    return names


def compare_orientations(common_view_names, reference_reconstruction, reconstruction_to_align, robust_alignment_threshold=0.0):
    """src/compare_reconstructions.cpp:228-262 (bind:652): gather the rotations of the common views, AlignRotations
    (robust), per-view AngularDifference."""
    ref_ids = {reference_reconstruction.View(v).Name(): v for v in reference_reconstruction.ViewIds()}
    our_ids = {reconstruction_to_align.View(v).Name(): v for v in reconstruction_to_align.ViewIds()}
    r1 = [np.asarray(reference_reconstruction.View(ref_ids[n]).GetOrientationAsAngleAxis(), dtype=np.float64) for n in common_view_names]
    r2 = [np.asarray(reconstruction_to_align.View(our_ids[n]).GetOrientationAsAngleAxis(), dtype=np.float64) for n in common_view_names]
    result = CompareInfo()
    if r1:
        AlignRotations(r1, r2)
        result.rotation_diff_when_align = [AngularDifference(a, b) for a, b in zip(r1, r2)]
    result.common_camera = len(common_view_names)
    return result


def test_loss_with_input_x(loss_func, x):       # bind:179-183
    out = [0.0, 0.0, 0.0]
    loss_func.Evaluate(x, out)
    print(f"[{out[0]:g}, {out[1]:g}, {out[2]:g}]")


def InitGlog(log_level=0, logtostderr=True, log_dir="./log"):
    return None


def StopGlog():
    return None


tgamma = math.gamma                              # bind:31,667

# MAGSAC constants (include/gamma_values.cpp:6-11,384-389,780-785) and tables (closed forms)
nu3, C3, sigma_quantile3, upper_incomplete_gamma_of_k3 = 3.0, 4.029720004054876e-01, 3.368214175218727, 3.439485560754856e-03
nu4, C4, sigma_quantile4, upper_incomplete_gamma_of_k4 = 4.0, 2.525252525252525e-01, 3.643721193503644e+00, 3.611260617758625e-03
nu9, C9, sigma_quantile9, upper_incomplete_gamma_of_k9 = 9.0, 3.837828575290349e-03, 4.654674460524809e+00, 3.344206155099048e-02
stored_gamma_number3, stored_gamma_number4, stored_gamma_number9 = 36843, 38683, 48553
precision_of_stored_gamma3 = precision_of_stored_gamma4 = precision_of_stored_gamma9 = 1000.0


def _table(nu, n):
    x = np.arange(n) / 1000.0
    if nu == 3:
        return np.exp(-x).tolist()
    if nu == 4:
        return (0.5 * math.sqrt(math.pi) * np.array([math.erfc(math.sqrt(v)) for v in x]) + np.sqrt(x) * np.exp(-x)).tolist()
    return (np.exp(-x) * (((x + 3.0) * x + 6.0) * x + 6.0)).tolist()


def __getattr__(name):                           # the three 36k..48k-entry lists are built on first use
    if name in ("stored_gamma_values3", "stored_gamma_values4", "stored_gamma_values9"):
        nu = int(name[-1])
        val = _table(nu, {3: stored_gamma_number3, 4: stored_gamma_number4, 9: stored_gamma_number9}[nu])
        globals()[name] = val
        return val
    raise AttributeError(name)
