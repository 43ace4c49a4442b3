/*
 * ra_oracle.h -- CPU ORACLE for the rotation-averaging hot path.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libgsfm_ra.so) never links or calls it.
 *
 * PARITY STATUS: the reference (GlobalSfMpy + TheiaSfM + Ceres 1.14 + Eigen +
 * SuiteSparse) cannot be compiled in this environment, and it ships no golden output
 * for the converged solve.  What IS pinned against reference material:
 *   - the residual functor against the four known-answer cases of
 *     theia/sfm/global_pose_estimation/pairwise_rotation_error_test.cc:87-139,
 *   - every robust loss against the unmodified scripts/loss_functions.py executed in
 *     the build container (tests/golden/make_loss_golden.py -> loss_golden.npz),
 *   - the MAGSAC gamma tables against include/gamma_values.cpp (sampled golden),
 *   - the translation residual (GSFM_RA_POSITION_BASELINE problems: theia::PairwiseTranslationError under
 *     src/GSfM_nonlinear_position_estimator.cpp:87-343) against the four known-answer cases of
 *     theia/sfm/global_pose_estimation/pairwise_translation_error_test.cc:69-127 and a numpy / finite-difference
 *     restatement of its Jacobian (tests/test_positions.py).
 * The trust-region loop restates Ceres Solver 1.14.0 (README.md:21, not vendored) from
 * its published algorithm: converged-solution parity vs. a Ceres binary is UNPINNED.
 */
#ifndef RA_ORACLE_H_
#define RA_ORACLE_H_

#include "../include/gsfm_ra.h"

#ifdef __cplusplus
extern "C" {
#endif

/* user loss callback: the reference evaluates a Python LossFunction once per edge per
 * evaluation (bind_src/GlobalSfMpy.cpp:36-59); pass one here to reproduce that cost. */
typedef void (*ra_oracle_loss_cb)(double s, double* rho3, void* ctx);

/* rho[3] at s; formulas of scripts/loss_functions.py. */
void ra_oracle_loss(const gsfm_ra_loss* loss, double s, double* rho3);
/* closed-form value of stored_gamma_values{nu}[index] (include/gamma_values.cpp). */
double ra_oracle_gamma_table(int nu, int index);

/* ceres/rotation.h restatements (double). R is row-major 3x3. */
void ra_oracle_angle_axis_to_matrix(const double* w, double* R);
void ra_oracle_matrix_to_angle_axis(const double* R, double* w);

/* Whitening U (row-major 3x3) for one edge (rotation_estimator.cpp:251-288). */
void ra_oracle_whiten(int error_type, const double* cov6, double edge_weight, double* U);

/* One edge through forward-mode jets exactly as ceres::AutoDiffCostFunction would run
 * PairwiseRotationErrorAngleAxis / PairwiseRotationError: r[3], Ji[9], Jj[9] row-major. */
void ra_oracle_edge(const double* wi, const double* wj, const double* wij, const double* U,
                    double* r, double* Ji, double* Jj);

/* All edges: r [E][d], Ji/Jj [E][d][3], rho [E][3], d = the residual dimension of the error type (3; 4 for QUATERNION_NORM,
 * 9 for ROTATION_MAT_FNORM); any output may be NULL.
 * Every entry point below also takes a TRANSLATION problem (error_type GSFM_RA_POSITION_BASELINE, include/gsfm_pa.h):
 * `omega` then holds camera positions [N][3], problem->omega_ij the pairs' position_2, problem->orientation the global
 * orientations, and problem->fixed_view the parameter block held constant (its gradient and the off-diagonal blocks that
 * touch it are zero in the assembled system -- the reduced program Ceres solves). */
int ra_oracle_eval_edges(const gsfm_ra_problem* p, const gsfm_ra_loss* loss, const double* omega,
                         double* r, double* Ji, double* Jj, double* rho, int num_threads);

/* Robustified normal equations (Ceres Corrector applied), same layout as gsfm_ra_assemble. */
int ra_oracle_assemble(const gsfm_ra_problem* p, const gsfm_ra_loss* loss, const double* omega,
                       double* cost, double* gradient, double* hdiag,
                       uint32_t* rowptr, uint32_t* col, double* val, int num_threads);
int ra_oracle_cost(const gsfm_ra_problem* p, const gsfm_ra_loss* loss, const double* omega,
                   double* cost, int num_threads);

/* Ceres-1.14 trust-region LM on the problem.  options->linear_solver:
 *   GSFM_RA_SOLVER_DENSE_CHOLESKY  exact solve (stands in for SPARSE_NORMAL_CHOLESKY)
 *   GSFM_RA_SOLVER_PCG             block-Jacobi PCG at options->pcg_rtol
 * loss_cb != NULL overrides options->loss (reference-faithful Python-loss timing). */
int ra_oracle_solve(const gsfm_ra_problem* p, const gsfm_ra_options* options, double* omega_inout,
                    gsfm_ra_summary* summary, ra_oracle_loss_cb loss_cb, void* cb_ctx);

/* EstimateRotationsWithSigmaConsensus (rotation_estimator.cpp:314-457); weights_out [E] = the last weights, may be NULL. */
int ra_oracle_solve_sigma_consensus(const gsfm_ra_problem* p, const gsfm_ra_options* options, int32_t iters_num, double sigma_max,
                                    double* omega_inout, gsfm_ra_summary* summary, double* weights_out);

/* FilterViewPairsFromOrientation restatement. */
int ra_oracle_filter_view_pairs(const gsfm_ra_problem* p, const double* omega, double max_degrees,
                                uint8_t* keep, double* angle_rad);

#ifdef __cplusplus
}
#endif
#endif
