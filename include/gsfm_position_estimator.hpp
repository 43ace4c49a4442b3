// gsfm_position_estimator.hpp -- the reference's C++ position-estimator interface over the C ABI of gsfm_pa.h.
//
// Header-only host-side mirror of
//   theia::PositionEstimator                 T/sfm/global_pose_estimation/position_estimator.h
//   theia::GSfMNonlinearPositionEstimator    include/GSfM_nonlinear_position_estimator.hpp:57-151
//                                            src/GSfM_nonlinear_position_estimator.cpp:66-234 (both EstimatePositions overloads)
// Same method names, argument order and meaning; same contract for the maps: `orientation` is read, `positions` is OUTPUT -- the
// estimator creates an entry for every view that has an orientation AND appears in a view pair (InitializeRandomPositions,
// :236-257; the reference overwrites its random draw with the origin, so every camera starts at (0,0,0)), holds one of them
// constant at the origin (:121-122) and returns the rest.  View pairs with an endpoint outside `positions` are skipped
// (:307-312).  Returns false for empty inputs (:93-98) and when the solve is not usable (Ceres' IsSolutionUsable: here a
// non-finite cost, or a failed device call -- see last_error()).
//
// Deliberate differences, all stated:
//   * the loss is a gsfm_ra_loss value instead of a borrowed ceres::LossFunction* (include/gsfm_rotation_estimator.hpp shows the
//     adapters, TabulatedLoss takes any object with Evaluate(double, double[3]));
//   * the constant view: the reference fixes positions->begin() of an std::unordered_map -- whatever the hash order puts first.
//     Here it is the SMALLEST view id among the estimated views unless options.fixed_view_id names one; the choice only moves the
//     solution by a global translation (the cost is translation invariant);
//   * Options::min_num_points_per_view > 0 (point-to-camera constraints from tracks, :345-466) is not on this path
//     (flags_1dsfm.yaml sets it to 0): EstimatePositions returns false with an explanatory last_error();
//   * Options::rng is not needed: the random initial positions are never used by the reference either.
#ifndef GSFM_POSITION_ESTIMATOR_HPP_
#define GSFM_POSITION_ESTIMATOR_HPP_

#include <algorithm>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "gsfm_pa.h"
#include "gsfm_rotation_estimator.hpp"  // the loss factories

namespace gsfm_b200 {

// theia::PositionErrorType  include/pairwise_translation_error_covariance.hpp:47-51
enum class PositionErrorType { BASELINE = 0, COVARIANCE = 1 };

// ---- theia::PositionEstimator --------------------------------------------------------------------------------------
template <class ViewPairs, class Vec3Map>
class PositionEstimator {
 public:
  virtual ~PositionEstimator() {}
  // Input: the view pairs (relative translation directions) and the global orientations; output: the positions.
  virtual bool EstimatePositions(const ViewPairs& view_pairs, const Vec3Map& orientation, Vec3Map* positions) = 0;
};

// ---- theia::GSfMNonlinearPositionEstimator -----------------------------------------------------------------------------
template <class ViewPairs, class Vec3Map>
class GSfMNonlinearPositionEstimator : public PositionEstimator<ViewPairs, Vec3Map> {
 public:
  typedef typename Vec3Map::key_type Id;
  // theia::NonlinearPositionEstimator::Options  T/sfm/global_pose_estimation/nonlinear_position_estimator.h:62-83
  struct Options {
    int num_threads = 1;
    int max_num_iterations = 400;
    double robust_loss_width = 0.1;
    int min_num_points_per_view = 0;
    double point_to_camera_weight = 0.5;
    // not in the reference (see the header comment): the view held constant; negative = the smallest estimated view id
    long long fixed_view_id = -1;
    // devices to shard the view pairs over (gsfm_ra_options::n_gpus; -1 = as many as keep >= 500k pairs each)
    int n_gpus = -1;
  };

  explicit GSfMNonlinearPositionEstimator(const Options& options = Options()) : options_(options) {}

  // :87-149: HuberLoss(robust_loss_width), PairwiseTranslationError with weight 1.
  bool EstimatePositions(const ViewPairs& view_pairs, const Vec3Map& orientation, Vec3Map* positions) override {
    return Run(view_pairs, orientation, positions, HuberLoss(options_.robust_loss_width), PositionErrorType::BASELINE);
  }
  // :151-234: the caller's loss.  error_type is accepted as the reference accepts it (both values build the BASELINE residual,
  // because the overload calls the two-argument AddCameraToCameraConstraints at :183).
  bool EstimatePositions(const ViewPairs& view_pairs, const Vec3Map& orientation, Vec3Map* positions, PositionErrorType error_type,
                         const gsfm_ra_loss& loss_func) {
    return Run(view_pairs, orientation, positions, loss_func, error_type);
  }

  const gsfm_ra_summary& summary() const { return summary_; }
  const std::string& last_error() const { return error_; }
  long long fixed_view() const { return fixed_; }

 private:
  bool Run(const ViewPairs& view_pairs, const Vec3Map& orientation, Vec3Map* positions, const gsfm_ra_loss& loss, PositionErrorType error_type) {
    error_.clear();
    if (positions == nullptr) { error_ = "positions is null"; return false; }               // the reference CHECK-aborts (:91)
    if (view_pairs.size() == 0 || orientation.size() == 0) return false;                    // :93-98
    if (options_.min_num_points_per_view > 0) { error_ = "point-to-camera constraints (min_num_points_per_view > 0) are not on this path"; return false; }
    // InitializeRandomPositions (:236-257): a position for every oriented view that some view pair mentions, at the origin
    std::unordered_set<Id> constrained;
    for (const auto& vp : view_pairs) { constrained.insert(vp.first.first); constrained.insert(vp.first.second); }
    std::vector<Id> ids;
    for (const auto& kv : orientation) if (constrained.count(kv.first)) ids.push_back(kv.first);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    if (ids.empty()) return false;
    std::unordered_map<Id, uint32_t> dense;
    for (uint32_t k = 0; k < ids.size(); ++k) dense[ids[k]] = k;
    std::vector<double> orient(3 * ids.size(), 0.0), pos(3 * ids.size(), 0.0);
    for (uint32_t k = 0; k < ids.size(); ++k) {
      const auto it = orientation.find(ids[k]);
      if (it != orientation.end()) for (int t = 0; t < 3; ++t) orient[3 * k + t] = it->second[t];
    }
    std::vector<uint32_t> ei, ej;
    std::vector<double> p2;
    for (const auto& vp : view_pairs) {
      const auto a = dense.find(vp.first.first), b = dense.find(vp.first.second);
      if (a == dense.end() || b == dense.end()) continue;                                   // :309-312
      ei.push_back(a->second); ej.push_back(b->second);
      for (int t = 0; t < 3; ++t) p2.push_back(vp.second.position_2[t]);
    }
    fixed_ = options_.fixed_view_id >= 0 && dense.count((Id)options_.fixed_view_id) ? options_.fixed_view_id : (long long)ids.front();
    summary_ = gsfm_ra_summary{};
    if (!ei.empty()) {
      gsfm_pa_problem p{};
      p.num_views = (uint32_t)ids.size(); p.num_edges = ei.size();
      p.edge_i = ei.data(); p.edge_j = ej.data(); p.position_2 = p2.data(); p.orientation = orient.data();
      p.fixed_view = (int64_t)dense[(Id)fixed_];
      p.error_type = (int32_t)error_type;
      gsfm_ra_options o;
      gsfm_pa_default_options(&o);
      o.loss = loss;
      o.max_num_iterations = options_.max_num_iterations;
      o.num_threads = options_.num_threads;
      o.n_gpus = options_.n_gpus;
      const int rc = gsfm_pa_solve(&p, &o, pos.data(), &summary_);
      if (rc != 0) { error_ = gsfm_ra_last_error(); return false; }                          // incl. GSFM_RA_ERR_NUMERIC: not usable
    }
    for (uint32_t k = 0; k < ids.size(); ++k) {
      auto& v = (*positions)[ids[k]];
      for (int t = 0; t < 3; ++t) v[t] = pos[3 * k + t];
    }
    // Ceres: IsSolutionUsable() = CONVERGENCE, NO_CONVERGENCE or USER_SUCCESS -- everything but a failure
    return summary_.termination != GSFM_RA_TERM_FAILURE;
  }

  const Options options_;
  gsfm_ra_summary summary_{};
  std::string error_;
  long long fixed_ = -1;
};

}  // namespace gsfm_b200

#endif  // GSFM_POSITION_ESTIMATOR_HPP_
