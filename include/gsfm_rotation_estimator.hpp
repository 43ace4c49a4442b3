// gsfm_rotation_estimator.hpp -- the reference's C++ estimator interface over the C ABI of gsfm_ra.h.
//
// Header-only host-side mirror of
//   theia::RotationEstimator                 T/sfm/global_pose_estimation/rotation_estimator.h:50-66
//   theia::GSfMNonlinearRotationEstimator    include/GSfM_nonlinear_rotation_estimator.hpp:22-59
//                                            src/GSfM_nonlinear_rotation_estimator.cpp:24-80, 82-198, 201-309, 314-457
// Same method names, argument order and meaning, same in/out contract for the orientation map, same skipping rules for
// edges whose endpoints have no initial orientation or (covariance types) no covariance, same return values
// (false only for empty inputs ... and, unlike Ceres, when the device call itself fails: see last_error()).
//
// The methods are templates over the container types, so the header compiles unchanged against Theia / Eigen
//   (std::unordered_map<theia::ViewIdPair, theia::TwoViewInfo>, std::unordered_map<theia::ViewId, Eigen::Vector3d>,
//    theia::CovarianceMap = unordered_map<ViewIdPair, pair<Eigen::Matrix3d, Eigen::Vector3d>>)
// and against any stand-in with the same shape (tests/cpp/test_shim.cc):
//   view_pairs          iterable of pair<pair<Id, Id>, Info>, Info::rotation_2 indexable [0..2]   (T/sfm/twoview_info.h:54-98)
//   global_orientations map Id -> V, V indexable [0..2]; find / end / size
//   covariances         map pair<Id, Id> -> pair<M, *>, M callable (row, col)                      (src/uncertainty.cpp:200-229)
// The only deliberate difference: the loss is a gsfm_ra_loss value (kind + parameters, include/gsfm_ra.h) instead of a
// borrowed ceres::LossFunction*; the factories below carry the names of scripts/loss_functions.py.  The maintainer-side
// adapter from the pybind11 trampoline (bind_src/GlobalSfMpy.cpp:33-65) to gsfm_ra_loss is shown in INTEGRATION.md.
#ifndef GSFM_ROTATION_ESTIMATOR_HPP_
#define GSFM_ROTATION_ESTIMATOR_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "gsfm_ra.h"

namespace gsfm_b200 {

// ---- losses, named as in scripts/loss_functions.py ------------------------------------------------------------
inline gsfm_ra_loss MakeLoss(int kind, double p0 = 0.0, double p1 = 0.0, unsigned flags = 0, double scale = 1.0) {
  gsfm_ra_loss l{};
  l.kind = kind; l.flags = flags; l.p[0] = p0; l.p[1] = p1; l.scale = scale;
  return l;
}
inline gsfm_ra_loss TrivialLoss() { return MakeLoss(GSFM_RA_LOSS_TRIVIAL); }
inline gsfm_ra_loss HuberLoss(double a) { return MakeLoss(GSFM_RA_LOSS_HUBER, a); }
inline gsfm_ra_loss SoftLOneLoss(double a) { return MakeLoss(GSFM_RA_LOSS_SOFTLONE, a); }
inline gsfm_ra_loss CauchyLoss(double a) { return MakeLoss(GSFM_RA_LOSS_CAUCHY, a); }
inline gsfm_ra_loss ArctanLoss(double a) { return MakeLoss(GSFM_RA_LOSS_ARCTAN, a); }
inline gsfm_ra_loss TolerantLoss(double a, double b) { return MakeLoss(GSFM_RA_LOSS_TOLERANT, a, b); }
inline gsfm_ra_loss TukeyLoss(double a) { return MakeLoss(GSFM_RA_LOSS_TUKEY, a); }
inline gsfm_ra_loss GemanMcClureLoss(double a, double sigma) { return MakeLoss(GSFM_RA_LOSS_GEMANMCCLURE, a, sigma); }
// MAGSACWeightBasedLoss(sigma, use_weight_inverse) scripts/loss_functions.py:285-341 (nu = 3); nu = 4 / 9: :344-459
inline gsfm_ra_loss MAGSACWeightBasedLoss(double sigma, bool use_weight_inverse = false, int nu = 3) {
  const int kind = nu == 3 ? GSFM_RA_LOSS_MAGSAC3 : (nu == 4 ? GSFM_RA_LOSS_MAGSAC4 : GSFM_RA_LOSS_MAGSAC9);
  return MakeLoss(kind, sigma, 0.0, use_weight_inverse ? 1u : 0u);
}
inline gsfm_ra_loss ScaledLoss(gsfm_ra_loss rho, double a) { rho.scale = (rho.scale == 0.0 ? 1.0 : rho.scale) * a; return rho; }
// ComposedLoss(f, g): rho(s) = f(g(s))  scripts/loss_functions.py:250-265.  f, g: closed-form losses (g below MAGSAC), neither
// composed itself; anything deeper goes through TabulatedLoss.
inline gsfm_ra_loss ComposedLoss(const gsfm_ra_loss& f, const gsfm_ra_loss& g) {
  gsfm_ra_loss l = f;
  l.inner_kind = g.kind; l.inner_flags = g.flags; l.inner_scale = (g.scale == 0.0 ? 1.0 : g.scale);
  for (int k = 0; k < 4; ++k) l.inner_p[k] = g.p[k];
  return l;
}

// ANY loss object with the ceres::LossFunction interface -- `void Evaluate(double s, double out[3]) const` -- as a device
// table (GSFM_RA_LOSS_TABULATED): this is the adapter for the borrowed `ceres::LossFunction*` the reference's estimator takes
// (include/GSfM_nonlinear_rotation_estimator.hpp:41-49) and for the pybind11 trampoline pyLossFunction
// (bind_src/GlobalSfMpy.cpp:33-65).  The object is called once per knot (2 + 144 * 32 times) instead of once per edge per
// evaluation; the table lives in this adapter, which must outlive the solve.
class TabulatedLoss {
 public:
  static constexpr int kMinExp = -80, kOctaves = 144, kPerOctave = 32;
  template <class LossLike>
  explicit TabulatedLoss(const LossLike& f) : table_(3 * (size_t)(2 + kOctaves * kPerOctave)) {
    size_t row = 0;
    auto put = [&](double s) { f.Evaluate(s, &table_[3 * row]); ++row; };
    put(0.0);
    for (int o = 0; o < kOctaves; ++o)
      for (int m = 0; m < kPerOctave; ++m) put(std::ldexp(1.0 + (double)m / kPerOctave, kMinExp + o));
    put(std::ldexp(1.0, kMinExp + kOctaves));
    loss_ = MakeLoss(GSFM_RA_LOSS_TABULATED);
    loss_.table = table_.data();
    loss_.table_min_exp = kMinExp; loss_.table_octaves = kOctaves; loss_.table_per_octave = kPerOctave;
  }
  TabulatedLoss(const TabulatedLoss&) = delete;
  TabulatedLoss& operator=(const TabulatedLoss&) = delete;
  const gsfm_ra_loss& get() const { return loss_; }
  operator const gsfm_ra_loss&() const { return loss_; }

 private:
  std::vector<double> table_;
  gsfm_ra_loss loss_;
};

// ---- theia::RotationEstimator ------------------------------------------------------------------------------------
template <class ViewPairs, class Orientations>
class RotationEstimator {
 public:
  virtual ~RotationEstimator() {}
  // Input: the view pairs (relative rotations) and an initial guess for every view to solve; output: the map is
  // overwritten in place.  Returns true on success.
  virtual bool EstimateRotations(const ViewPairs& view_pairs, Orientations* rotations) = 0;
};

namespace detail {

inline bool NeedsCovariance(int t) {
  return t == GSFM_RA_ANGLE_AXIS_COVARIANCE || t == GSFM_RA_ANGLE_AXIS_COV_INLIERS || t == GSFM_RA_ANGLE_AXIS_COVTRACE ||
         t == GSFM_RA_ANGLE_AXIS_COVNORM;
}
inline bool NeedsMatches(int t) { return t == GSFM_RA_ANGLE_AXIS_INLIERS || t == GSFM_RA_ANGLE_AXIS_COV_INLIERS; }

struct NoCovariances {};
struct NoMatches {
  template <class Pair>
  double operator()(const Pair&) const { return 100.0; }
};

// Flatten the hash maps into the arrays of gsfm_ra_problem: dense view numbering in ascending id order, edges in the
// map's iteration order with the reference's skipping rules (rotation_estimator.cpp:57-60, 231-247).
template <class Id>
struct Flat {
  std::vector<Id> ids;
  std::vector<uint32_t> ei, ej;
  std::vector<double> wij, cov6, weight, omega;
};

template <class ViewPairs, class Orientations, class Covariances, class Matches>
Flat<typename Orientations::key_type> Flatten(const ViewPairs& view_pairs, const Orientations& orientations, const Covariances* covariances,
                                               const Matches& matches, int error_type) {
  using Id = typename Orientations::key_type;
  Flat<Id> f;
  f.ids.reserve(orientations.size());
  for (const auto& kv : orientations) f.ids.push_back(kv.first);
  std::sort(f.ids.begin(), f.ids.end());
  std::unordered_map<Id, uint32_t> dense;
  dense.reserve(f.ids.size());
  for (uint32_t k = 0; k < f.ids.size(); ++k) dense[f.ids[k]] = k;
  f.omega.resize(3 * f.ids.size());
  for (uint32_t k = 0; k < f.ids.size(); ++k) {
    const auto& v = orientations.find(f.ids[k])->second;
    for (int t = 0; t < 3; ++t) f.omega[3 * k + t] = v[t];
  }
  const bool cov = NeedsCovariance(error_type), inl = NeedsMatches(error_type);
  for (const auto& vp : view_pairs) {
    const auto a = dense.find(vp.first.first), b = dense.find(vp.first.second);
    if (a == dense.end() || b == dense.end()) continue;
    if constexpr (!std::is_same<Covariances, NoCovariances>::value) {
      if (cov) {
        const auto c = covariances->find(vp.first);
        if (c == covariances->end()) continue;
        const auto& S = c->second.first;  // covariance_rot.txt order: C00 C11 C22 C01 C02 C12
        for (double x : {S(0, 0), S(1, 1), S(2, 2), S(0, 1), S(0, 2), S(1, 2)}) f.cov6.push_back(x);
      }
    }
    f.ei.push_back(a->second);
    f.ej.push_back(b->second);
    for (int t = 0; t < 3; ++t) f.wij.push_back(vp.second.rotation_2[t]);
    if (inl) f.weight.push_back(matches(vp.first) / 100.0);  // features.first.size() / 100.0, :261-263
  }
  return f;
}

}  // namespace detail

// ---- theia::GSfMNonlinearRotationEstimator ------------------------------------------------------------------------
template <class ViewPairs, class Orientations>
class GSfMNonlinearRotationEstimator : public RotationEstimator<ViewPairs, Orientations> {
 public:
  GSfMNonlinearRotationEstimator() : robust_loss_width_(0.1) {}
  explicit GSfMNonlinearRotationEstimator(const double robust_loss_width) : robust_loss_width_(robust_loss_width) {}

  // rotation_estimator.cpp:24-80: SoftLOneLoss(robust_loss_width), PairwiseRotationError with weight 1.
  bool EstimateRotations(const ViewPairs& view_pairs, Orientations* global_orientations) override {
    return Run(view_pairs, global_orientations, SoftLOneLoss(robust_loss_width_), 1, GSFM_RA_ANGLE_AXIS, static_cast<const detail::NoCovariances*>(nullptr),
               detail::NoMatches(), 0, 0.0);
  }

  // :82-198: quaternion parameter blocks with EigenQuaternionParameterization; QUATERNION_COSINE / QUATERNION_NORM /
  // ROTATION_MAT_FNORM.
  bool EstimateRotationsWithCustomizedLoss(const ViewPairs& view_pairs, Orientations* global_orientations, const gsfm_ra_loss& loss_function,
                                           int thread_num, int rotation_error_type = GSFM_RA_QUATERNION_COSINE) {
    if (rotation_error_type > GSFM_RA_QUATERNION_COSINE) { error_ = "EstimateRotationsWithCustomizedLoss takes the quaternion error types 0..2"; return false; }
    return Run(view_pairs, global_orientations, loss_function, thread_num, rotation_error_type, static_cast<const detail::NoCovariances*>(nullptr),
               detail::NoMatches(), 0, 0.0);
  }

  // :201-309: the angle-axis error types 3..8.  `covariances` is only read (the reference copies it by value);
  // `num_matched_features(view_id_pair)` replaces get_matched_features(view_id_pair, *reconstruction, features).first.size()
  // for the *_INLIERS types (:260-274).
  template <class Covariances, class Matches = detail::NoMatches>
  bool EstimateRotationsWithCustomizedLossAndCovariance(const ViewPairs& view_pairs, Orientations* global_orientations,
                                                        const gsfm_ra_loss& loss_function, int thread_num, const Covariances& covariances,
                                                        int rotation_error_type, const Matches& num_matched_features = Matches()) {
    if (rotation_error_type < GSFM_RA_ANGLE_AXIS_COVARIANCE) { error_ = "EstimateRotationsWithCustomizedLossAndCovariance takes the angle-axis error types 3..8"; return false; }
    return Run(view_pairs, global_orientations, loss_function, thread_num, rotation_error_type, &covariances, num_matched_features, 0, 0.0);
  }

  // :314-457: outer re-weighting loop (weights from the nu = 3 gamma table at sigma_max), PairwiseRotationError with the weight.
  bool EstimateRotationsWithSigmaConsensus(const ViewPairs& view_pairs, Orientations* global_orientations, const gsfm_ra_loss& loss_function,
                                           int thread_num, int iters_num, double sigma_max) {
    return Run(view_pairs, global_orientations, loss_function, thread_num, GSFM_RA_ANGLE_AXIS, static_cast<const detail::NoCovariances*>(nullptr),
               detail::NoMatches(), iters_num, sigma_max);
  }

  // Options of the last / next solve (Ceres defaults of :299-303 unless changed) and what the last solve reported.
  gsfm_ra_options& options() { EnsureOptions(); return options_; }
  const gsfm_ra_summary& summary() const { return summary_; }
  const std::string& last_error() const { return error_; }

 private:
  void EnsureOptions() {
    if (!options_ready_) {
      gsfm_ra_default_options(&options_);
      options_.n_gpus = -1;  // behind the plugin API the view graph shards over the box's GPUs by itself once it is large enough
      options_ready_ = true;
    }
  }

  template <class Covariances, class Matches>
  bool Run(const ViewPairs& view_pairs, Orientations* global_orientations, const gsfm_ra_loss& loss, int thread_num, int error_type,
           const Covariances* covariances, const Matches& matches, int sigma_iters, double sigma_max) {
    error_.clear();
    if (global_orientations == nullptr) { error_ = "global_orientations is null"; return false; }  // the reference CHECK-aborts (:208)
    if (global_orientations->size() == 0 || view_pairs.size() == 0) return false;                  // :209-220
    auto f = detail::Flatten(view_pairs, *global_orientations, covariances, matches, error_type);
    if (f.ei.empty()) return true;  // nothing to constrain: Ceres solves an empty problem and the reference returns true
    gsfm_ra_problem p{};
    p.num_views = static_cast<uint32_t>(f.ids.size());
    p.num_edges = f.ei.size();
    p.edge_i = f.ei.data(); p.edge_j = f.ej.data(); p.omega_ij = f.wij.data();
    p.cov6 = f.cov6.empty() ? nullptr : f.cov6.data();
    p.edge_weight = f.weight.empty() ? nullptr : f.weight.data();
    p.error_type = error_type;
    p.total_pair_count = sigma_iters > 0 && view_pairs.size() < 0x7fffffffu ? static_cast<int32_t>(view_pairs.size()) : 0;
    EnsureOptions();
    gsfm_ra_options o = options_;
    o.loss = loss;
    o.num_threads = thread_num;
    summary_ = gsfm_ra_summary{};
    const int rc = sigma_iters > 0 ? gsfm_ra_solve_sigma_consensus(&p, &o, sigma_iters, sigma_max, f.omega.data(), &summary_)
                                   : gsfm_ra_solve(&p, &o, f.omega.data(), &summary_);
    if (rc != 0 && rc != GSFM_RA_ERR_NUMERIC) { error_ = gsfm_ra_last_error(); return false; }
    for (uint32_t k = 0; k < f.ids.size(); ++k) {
      auto& v = global_orientations->find(f.ids[k])->second;
      for (int t = 0; t < 3; ++t) v[t] = f.omega[3 * k + t];
    }
    return true;  // the reference returns true whatever Ceres reports (:79, :197, :308, :456)
  }

  const double robust_loss_width_;
  gsfm_ra_options options_{};
  bool options_ready_ = false;
  gsfm_ra_summary summary_{};
  std::string error_;
};

// ---- the step before the solve -------------------------------------------------------------------------------------
// theia::OrientationsFromMaximumSpanningTree(const ViewGraph&, unordered_map<ViewId, Vector3d>*)
//   T/sfm/view_graph/orientations_from_maximum_spanning_tree.cc:109-178 -- here over the same view-pair map the
//   estimators take (Info::num_verified_matches is the tree weight, Info::rotation_2 the relative rotation of the pair
//   (smaller id, larger id)), on the device through gsfm_ra_init_orientations_mst.  Fills `orientations` with the views
//   the root's component reaches (root: the smallest view id that has a pair); returns false on empty input or a failed
//   device call (message in *error when given).
template <class ViewPairs, class Orientations>
bool OrientationsFromMaximumSpanningTree(const ViewPairs& view_pairs, Orientations* orientations, std::string* error = nullptr) {
  using Id = typename Orientations::key_type;
  if (!orientations || view_pairs.size() == 0) return false;
  std::vector<Id> ids;
  for (const auto& vp : view_pairs) { ids.push_back(vp.first.first); ids.push_back(vp.first.second); }
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  auto dense = [&](const Id& v) { return (uint32_t)(std::lower_bound(ids.begin(), ids.end(), v) - ids.begin()); };
  std::vector<uint32_t> ei, ej;
  std::vector<int32_t> w;
  std::vector<double> wij;
  for (const auto& vp : view_pairs) {
    ei.push_back(dense(vp.first.first)); ej.push_back(dense(vp.first.second));
    w.push_back((int32_t)vp.second.num_verified_matches);
    for (int t = 0; t < 3; ++t) wij.push_back(vp.second.rotation_2[t]);
  }
  std::vector<double> omega(3 * ids.size());
  const int rc = gsfm_ra_init_orientations_mst((uint32_t)ids.size(), ei.size(), ei.data(), ej.data(), wij.data(), w.data(), -1, omega.data(),
                                               nullptr, nullptr, -1);
  if (rc != 0) { if (error) *error = gsfm_ra_last_error(); return false; }
  for (uint32_t k = 0; k < ids.size(); ++k) {
    if (omega[3 * k] != omega[3 * k]) continue;  // NaN: not reachable from the root
    auto& v = (*orientations)[ids[k]];
    for (int t = 0; t < 3; ++t) v[t] = omega[3 * k + t];
  }
  return true;
}

// GSfMGlobalReconstructionEstimator::FilterInitialViewGraph  src/GSfM_global_reconstruction_estimator.cpp:369-390:
// erases the view pairs with fewer than min_num_two_view_inliers verified matches and everything outside the largest
// connected component (gsfm_ra_filter_initial_view_graph).  Returns false when no pair survives or the device call fails.
template <class ViewPairs>
bool FilterInitialViewGraph(ViewPairs* view_pairs, int min_num_two_view_inliers, std::string* error = nullptr) {
  if (!view_pairs || view_pairs->size() == 0) return false;
  using Id = typename std::decay<decltype(view_pairs->begin()->first.first)>::type;
  std::vector<Id> ids;
  for (const auto& vp : *view_pairs) { ids.push_back(vp.first.first); ids.push_back(vp.first.second); }
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  auto dense = [&](const Id& v) { return (uint32_t)(std::lower_bound(ids.begin(), ids.end(), v) - ids.begin()); };
  std::vector<uint32_t> ei, ej;
  std::vector<int32_t> m;
  std::vector<typename ViewPairs::key_type> keys;
  for (const auto& vp : *view_pairs) {
    keys.push_back(vp.first);
    ei.push_back(dense(vp.first.first)); ej.push_back(dense(vp.first.second));
    m.push_back((int32_t)vp.second.num_verified_matches);
  }
  std::vector<uint8_t> ekeep(ei.size()), vkeep(ids.size());
  const int rc = gsfm_ra_filter_initial_view_graph((uint32_t)ids.size(), ei.size(), ei.data(), ej.data(), m.data(), min_num_two_view_inliers,
                                                   ekeep.data(), vkeep.data(), -1);
  if (rc != 0) { if (error) *error = gsfm_ra_last_error(); return false; }
  for (size_t k = 0; k < keys.size(); ++k)
    if (!ekeep[k]) view_pairs->erase(keys[k]);
  return view_pairs->size() >= 1;
}

// ---- covariance_rot.txt ---------------------------------------------------------------------------------------------
// read_covariance / store_covariance_rot of the reference (src/uncertainty.cpp:200-229, :164-198) over the native reader
// and writer of the library (host-only code: no CUDA device needed).  `Covariances` is the reference's CovarianceMap shape:
// map pair<Id, Id> -> pair<M, V>, M with a mutable (row, col) accessor, V indexable.
template <class Covariances>
bool ReadCovariance(const std::string& dataset_directory, Covariances* covariances, std::string* error = nullptr) {
  if (!covariances) return false;
  uint64_t n = 0;
  uint32_t *a = nullptr, *b = nullptr;
  double *c6 = nullptr, *r3 = nullptr;
  const std::string path = dataset_directory + "/covariance_rot.txt";
  if (gsfm_ra_read_covariance_rot(path.c_str(), &n, &a, &b, &c6, &r3) != 0) { if (error) *error = gsfm_ra_last_error(); return false; }
  for (uint64_t e = 0; e < n; ++e) {
    auto& entry = (*covariances)[{a[e], b[e]}];
    const double* c = c6 + 6 * e;  // C00 C11 C22 C01 C02 C12
    auto& S = entry.first;
    S(0, 0) = c[0]; S(1, 1) = c[1]; S(2, 2) = c[2];
    S(0, 1) = S(1, 0) = c[3]; S(0, 2) = S(2, 0) = c[4]; S(1, 2) = S(2, 1) = c[5];
    for (int t = 0; t < 3; ++t) entry.second[t] = r3[3 * e + t];
  }
  gsfm_ra_free(a); gsfm_ra_free(b); gsfm_ra_free(c6); gsfm_ra_free(r3);
  return true;
}

template <class Covariances>
bool StoreCovarianceRot(const std::string& dataset_directory, const Covariances& covariances, std::string* error = nullptr) {
  std::vector<uint32_t> a, b;
  std::vector<double> c6, r3;
  for (const auto& kv : covariances) {
    a.push_back((uint32_t)kv.first.first); b.push_back((uint32_t)kv.first.second);
    const auto& S = kv.second.first;
    for (double x : {S(0, 0), S(1, 1), S(2, 2), S(0, 1), S(0, 2), S(1, 2)}) c6.push_back(x);
    for (int t = 0; t < 3; ++t) r3.push_back(kv.second.second[t]);
  }
  const std::string path = dataset_directory + "/covariance_rot.txt";
  if (gsfm_ra_write_covariance_rot(path.c_str(), a.size(), a.data(), b.data(), c6.data(), r3.data()) != 0) {
    if (error) *error = gsfm_ra_last_error();
    return false;
  }
  return true;
}

}  // namespace gsfm_b200

#endif  // GSFM_ROTATION_ESTIMATOR_HPP_
