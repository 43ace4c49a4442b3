/*
 * gsfm_pa.h -- C ABI of the B200-native robust TRANSLATION averaging (camera positions), the step after rotation
 * averaging on the same edge-parallel machinery.
 *
 * Drop-in boundary for
 *   src/GSfM_nonlinear_position_estimator.cpp:87-149  EstimatePositions(view_pairs, orientations, positions)
 *   src/GSfM_nonlinear_position_estimator.cpp:151-234 EstimatePositions(..., PositionErrorType, ceres::LossFunction*)
 * reached through theia::PositionEstimator
 *   thirdparty/TheiaSfM/src/theia/sfm/global_pose_estimation/position_estimator.h
 * and, from Python, through GSfMGlobalReconstructionEstimator::EstimatePositionNonLinear (bind_src/GlobalSfMpy.cpp:564-579,
 * src/GSfM_global_reconstruction_estimator.cpp:605-617).
 *
 * What the reference does on this path (flags_1dsfm.yaml: position_estimation_min_num_tracks_per_view = 0, so there are
 * no point-to-camera constraints):
 *   - every position starts at the ORIGIN (InitializeRandomPositions overwrites its random draw with zero, :236-257);
 *   - one residual block per view pair whose two views have a position (:298-343):
 *       theia::PairwiseTranslationError(translation_direction, 1.0)
 *       (thirdparty/TheiaSfM/src/theia/sfm/global_pose_estimation/pairwise_translation_error.h:62-88)
 *       r = w * ((c_j - c_i) / |c_j - c_i| - t_ij),   |.| := 1 below 1e-12,
 *       t_ij = R(orientation_i)^T * TwoViewInfo::position_2   (GetRotatedTranslation, :36-44)
 *     under the caller's loss (HuberLoss(robust_loss_width = 0.1) for the plain overload);
 *     both overloads call the two-argument AddCameraToCameraConstraints (:110, :183), so PositionErrorType::COVARIANCE
 *     runs the same BASELINE residual -- the covariance variant at :259-296 is never reached;
 *   - positions->begin() is set to zero and held constant (:121-122);
 *   - Ceres trust region with max_num_iterations = 400 (nonlinear_position_estimator.h:70), SPARSE_NORMAL_CHOLESKY up
 *     to 1000 cameras, CGNR + JACOBI above (:129-147).
 *
 * The same solver as include/gsfm_ra.h runs it (error type GSFM_RA_POSITION_BASELINE): K1 evaluates the residual, its
 * closed-form Jacobian +-B, B = (w/|d|)(I - u u^T), and the robust loss per half-edge and assembles the block-3x3 Laplacian
 * stencil by warp-segmented reduction; the normal equations are solved by the persistent block-Jacobi PCG kernel (or the
 * on-device dense Cholesky for small graphs); the fixed view is removed by zeroing its gradient and the off-diagonal blocks
 * that touch it.  Plain C: host pointers and sizes; no CPU fallback.
 */
#ifndef GSFM_PA_H_
#define GSFM_PA_H_

#include "gsfm_ra.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as theia::PositionErrorType (include/pairwise_translation_error_covariance.hpp:47-51). */
typedef enum {
  GSFM_PA_BASELINE = 0,
  GSFM_PA_COVARIANCE = 1 /* accepted; runs the BASELINE residual exactly as the reference does (see above) */
} gsfm_pa_error_type;

/* One translation-averaging problem, views densely renumbered 0..num_views-1.  HOST memory, read only. */
typedef struct {
  uint32_t num_views;         /* N                                                                           */
  uint64_t num_edges;         /* E unique unordered pairs, edge_i[k] != edge_j[k]                            */
  const uint32_t* edge_i;     /* [E] view_id1 of the pair: its orientation rotates position_2                */
  const uint32_t* edge_j;     /* [E] view_id2                                                                */
  const double* position_2;   /* [E][3] TwoViewInfo::position_2 (T/sfm/twoview_info.h)                       */
  const double* orientation;  /* [N][3] global orientations, angle-axis (the rotation-averaging output)      */
  const double* edge_weight;  /* [E] or NULL (= 1.0, the weight the reference passes at :320)                */
  int64_t fixed_view;         /* held at its initial position (the reference: positions->begin()); < 0 none  */
  int32_t error_type;         /* gsfm_pa_error_type                                                          */
  int32_t reserved;
} gsfm_pa_problem;

/* gsfm_ra_default_options + what the position estimator changes: max_num_iterations = 400
 * (nonlinear_position_estimator.h:70) and HuberLoss(0.1) (robust_loss_width, :71; position_estimator.cpp:330).      */
void gsfm_pa_default_options(gsfm_ra_options* options);

/* The gsfm_ra_problem the solver of include/gsfm_ra.h takes for this problem (no copies: it points at the same arrays), so
 * every kernel-level entry point of gsfm_ra.h (gsfm_ra_eval_edges, gsfm_ra_assemble, gsfm_ra_cost, gsfm_ra_spmv, gsfm_ra_pcg,
 * the resident solver) works on a translation problem; the "omega" arguments then carry positions [N][3].            */
int gsfm_pa_as_ra_problem(const gsfm_pa_problem* problem, gsfm_ra_problem* out);

/* Replaces ceres::Solve at position_estimator.cpp:148 / :232.  positions_inout [N][3]: initial positions on entry (the
 * reference starts every camera at the origin), estimates on exit; the fixed view keeps its entry value.  options->n_gpus
 * shards the view pairs over devices exactly as gsfm_ra_solve does.                                                 */
int gsfm_pa_solve(const gsfm_pa_problem* problem, const gsfm_ra_options* options, double* positions_inout,
                  gsfm_ra_summary* summary);

/* Per view pair at positions [N][3]: r [E][3], d r/d c_i and d r/d c_j [E][3][3] row-major, rho [E][3] = loss at |r|^2.
 * Any output may be NULL.  Replaces one AutoDiffCostFunction<PairwiseTranslationError,3,3,3>::Evaluate +
 * LossFunction::Evaluate per pair.                                                                                   */
int gsfm_pa_eval_edges(const gsfm_pa_problem* problem, const gsfm_ra_loss* loss, const double* positions,
                       double* r, double* jac_i, double* jac_j, double* rho, int32_t device);

/* cost = sum over pairs of rho(|r|^2) / 2 at positions. */
int gsfm_pa_cost(const gsfm_pa_problem* problem, const gsfm_ra_loss* loss, const double* positions, double* cost,
                 int32_t device);

#ifdef __cplusplus
}
#endif
#endif /* GSFM_PA_H_ */
