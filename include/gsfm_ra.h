/*
 * gsfm_ra.h -- C ABI of the B200-native robust rotation-averaging solver.
 *
 * This is the drop-in boundary for ONE path of zhangganlin/GlobalSfMpy: the
 * Ceres-based robust rotation averaging in
 *   src/GSfM_nonlinear_rotation_estimator.cpp            (reference, all four entry points)
 * reached through Theia's plugin interface
 *   thirdparty/TheiaSfM/src/theia/sfm/global_pose_estimation/rotation_estimator.h:50-66
 * and, from Python, through bind_src/GlobalSfMpy.cpp:534-548.
 *
 * Plain C: pointers and sizes only, no torch / Eigen / STL types.  Host buffers
 * in, host buffers out (the library owns all device memory); a handle API keeps
 * a problem resident in HBM across iterations.  Every function returns 0 on
 * success and a negative gsfm_ra_status on failure; nothing aborts the process
 * (the reference CHECK-aborts on null pointers, rotation_estimator.cpp:28,88,208).
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with GSFM_RA_ERR_NO_DEVICE.
 *
 * Conventions: view i has world->camera rotation
 * R_i = Exp(omega_i); edge (i,j) carries omega_ij with R_j ~= R_ij * R_i
 * (thirdparty/TheiaSfM/src/theia/sfm/twoview_info.h:123-126).  The residual of an
 * edge is r = U * Log(R_j * R_i^T * R_ij^T) with U the 3x3 whitening / weight
 * (include/pairwise_rotation_error_quat.hpp:215-247 and
 *  theia/sfm/global_pose_estimation/pairwise_rotation_error.h:66-95).
 */
#ifndef GSFM_RA_H_
#define GSFM_RA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSFM_RA_ABI_VERSION 3

typedef enum {
  GSFM_RA_OK = 0,
  GSFM_RA_ERR_INVALID = -1,    /* bad argument / malformed problem                         */
  GSFM_RA_ERR_NO_DEVICE = -2,  /* no CUDA device: the product path has no CPU fallback     */
  GSFM_RA_ERR_CUDA = -3,       /* a CUDA runtime call or kernel failed                     */
  GSFM_RA_ERR_UNSUPPORTED = -4,/* valid request this build does not implement              */
  GSFM_RA_ERR_NUMERIC = -5     /* non-finite cost / linear solver breakdown                */
} gsfm_ra_status;

/* Same numeric values as theia::RotationErrorType,
 * include/pairwise_rotation_error_quat.hpp:50-61. */
typedef enum {
  GSFM_RA_QUATERNION_NORM = 0,
  GSFM_RA_ROTATION_MAT_FNORM = 1,
  GSFM_RA_QUATERNION_COSINE = 2,
  GSFM_RA_ANGLE_AXIS_COVARIANCE = 3,  /* U = chol((1e8*Sigma)^-1)^T   rotation_estimator.cpp:251-256 */
  GSFM_RA_ANGLE_AXIS = 4,             /* U = I                         :257-259 */
  GSFM_RA_ANGLE_AXIS_INLIERS = 5,     /* U = edge_weight * I           :260-264 (weight = #matches/100 from the host) */
  GSFM_RA_ANGLE_AXIS_COV_INLIERS = 6, /* U = Lt * edge_weight          :265-274 */
  GSFM_RA_ANGLE_AXIS_COVTRACE = 7,    /* U = sqrt(1/trace(1e8 Sigma)) I :275-282 */
  GSFM_RA_ANGLE_AXIS_COVNORM = 8,     /* U = sqrt(1/||1e8 Sigma||_F) I  :283-288 */
  /* Translation averaging (the boundary for it is include/gsfm_pa.h): the "views" are camera POSITIONS,
   * residual theia::PairwiseTranslationError, PositionErrorType::BASELINE
   * (src/GSfM_nonlinear_position_estimator.cpp:252-296).  The same solver, layouts and kernels; gsfm_ra_problem carries
   * position_2 in omega_ij plus the two extra fields at its end.                                                      */
  GSFM_RA_POSITION_BASELINE = 16
} gsfm_ra_error_type;

/* Robust losses of scripts/loss_functions.py (formulas at the cited lines). */
typedef enum {
  GSFM_RA_LOSS_TRIVIAL = 0,       /* :47   p: -                                */
  GSFM_RA_LOSS_HUBER = 1,         /* :56   p[0]=a                              */
  GSFM_RA_LOSS_SOFTLONE = 2,      /* :74   p[0]=a                              */
  GSFM_RA_LOSS_CAUCHY = 3,        /* :88   p[0]=a                              */
  GSFM_RA_LOSS_ARCTAN = 4,        /* :101  p[0]=a                              */
  GSFM_RA_LOSS_TOLERANT = 5,      /* :114  p[0]=a p[1]=b                       */
  GSFM_RA_LOSS_TUKEY = 6,         /* :167  p[0]=a                              */
  GSFM_RA_LOSS_LONEHALF = 7,      /* :187  p[0]=a                              */
  GSFM_RA_LOSS_LTWO = 8,          /* :216  p[0]=a                              */
  GSFM_RA_LOSS_GEMANMCCLURE = 9,  /* :239  p[0]=a p[1]=sigma2                  */
  GSFM_RA_LOSS_MAGSAC3 = 10,      /* :285  p[0]=sigma, flags bit0 = inverse    */
  GSFM_RA_LOSS_MAGSAC4 = 11,      /* :344                                      */
  GSFM_RA_LOSS_MAGSAC9 = 12,      /* :402                                      */
  GSFM_RA_LOSS_TABULATED = 13     /* ANY LossFunction object (bind_src/GlobalSfMpy.cpp:33-65 accepts every Python subclass):
                                     its Evaluate sampled by the host into `table`, see below                          */
} gsfm_ra_loss_kind;

#define GSFM_RA_LOSS_FLAG_INVERSE 1u

/* A loss is  s -> scale * f(g(s))  (ComposedLoss, loss_functions.py:250-265, and ScaledLoss, :267-281):
 *   kind/flags/p        the OUTER function f
 *   inner_*             the INNER function g; a zero-initialised struct has inner_kind = GSFM_RA_LOSS_TRIVIAL and
 *                       inner_scale 0 (= 1), i.e. g(s) = s: no composition
 *   scale               ScaledLoss factor applied to the whole (0 = 1)
 * GSFM_RA_LOSS_TABULATED (as f; g must be trivial): rho, rho', rho'' of an arbitrary loss object sampled by the host at the knots
 *   s_0 = 0,   s_{1 + o * per_octave + m} = 2^(min_exp + o) * (1 + m / per_octave),   o < octaves, m < per_octave,
 *   plus the closing knot 2^(min_exp + octaves)  =>  table holds 2 + octaves * per_octave rows of 3 doubles.
 * The device interpolates rho with the quintic Hermite polynomial of (rho, rho', rho'') at the two enclosing knots, rho' with
 * the cubic Hermite polynomial of (rho', rho'') and rho'' as the derivative of that cubic; beyond the last knot rho is
 * continued linearly.
 * per_octave must be a power of two <= 4096.                                                                              */
typedef struct {
  int32_t kind;     /* gsfm_ra_loss_kind                                            */
  uint32_t flags;   /* GSFM_RA_LOSS_FLAG_*                                          */
  double p[4];      /* parameters, see gsfm_ra_loss_kind                            */
  double scale;     /* ScaledLoss factor 'a' (loss_functions.py:267-281); 1 = none  */
  int32_t inner_kind;
  uint32_t inner_flags;
  double inner_p[4];
  double inner_scale;
  const double* table;        /* HOST memory, [2 + table_octaves * table_per_octave][3]; copied by the library */
  int32_t table_min_exp;
  int32_t table_octaves;
  int32_t table_per_octave;
  int32_t table_reserved;
} gsfm_ra_loss;

/* One rotation-averaging problem, views densely renumbered 0..num_views-1 by the
 * caller (the C++/Python shims do this from the reference's hash maps). All
 * arrays are HOST memory and are only read.                                    */
typedef struct {
  uint32_t num_views;        /* N                                                   */
  uint64_t num_edges;        /* E; (edge_i[k], edge_j[k]) unique unordered pairs,
                                edge_i[k] != edge_j[k]                               */
  const uint32_t* edge_i;    /* [E]                                                 */
  const uint32_t* edge_j;    /* [E]   R_j ~= R_ij R_i                               */
  const double* omega_ij;    /* [E][3] TwoViewInfo::rotation_2                       */
  const double* cov6;        /* [E][6] C00 C11 C22 C01 C02 C12 (covariance_rot.txt
                                order, src/uncertainty.cpp:200-229) or NULL          */
  const double* edge_weight; /* [E] scalar weight or NULL (=1)                       */
  int32_t error_type;        /* gsfm_ra_error_type                                  */
  int32_t total_pair_count;  /* gsfm_ra_solve_sigma_consensus only: size of the caller's view-pair map INCLUDING the pairs it
                                skipped (the reference's stop test averages |w - w_prev| over that count,
                                rotation_estimator.cpp:419-424); 0 = num_edges                                      */
  /* GSFM_RA_POSITION_BASELINE only (ignored otherwise; ABI v3): */
  const double* orientation; /* [N][3] global orientations (angle-axis): an edge's direction is R(orientation[edge_i])^T * omega_ij[k],
                                omega_ij[k] = TwoViewInfo::position_2 (position_estimator.cpp:36-44, 271-272)           */
  int64_t fixed_view;        /* the view whose position is held constant (position_estimator.cpp:121-122), < 0 = none  */
} gsfm_ra_problem;

typedef enum {
  GSFM_RA_SOLVER_PCG = 0,            /* block-Jacobi PCG on the block-3x3 CSR system   */
  GSFM_RA_SOLVER_DENSE_CHOLESKY = 1, /* on-device dense LL^T, small problems only      */
  GSFM_RA_SOLVER_AUTO = 2            /* (default) exact dense LL^T -- the role of the reference's SPARSE_NORMAL_CHOLESKY,
                                        rotation_estimator.cpp:300 -- when num_views <= GSFM_RA_AUTO_DENSE_MAX_VIEWS on one
                                        cooperative-launch device, PCG otherwise                                          */
} gsfm_ra_linear_solver;
#define GSFM_RA_AUTO_DENSE_MAX_VIEWS 1024

/* Trust-region options: the Ceres 1.14 defaults the reference runs with
 * (rotation_estimator.cpp:299-303 overrides only linear solver, 200 iterations,
 * num_threads).  gsfm_ra_default_options() fills these.                        */
typedef struct {
  gsfm_ra_loss loss;
  int32_t max_num_iterations;          /* 200                                       */
  int32_t jacobi_scaling;              /* 1                                         */
  double function_tolerance;           /* 1e-6                                      */
  double gradient_tolerance;           /* 1e-10                                     */
  double parameter_tolerance;          /* 1e-8                                      */
  double initial_trust_region_radius;  /* 1e4                                       */
  double max_trust_region_radius;      /* 1e16                                      */
  double min_trust_region_radius;      /* 1e-32                                     */
  double min_relative_decrease;        /* 1e-3                                      */
  double min_lm_diagonal;              /* 1e-6                                      */
  double max_lm_diagonal;              /* 1e32                                      */
  int32_t linear_solver;               /* gsfm_ra_linear_solver                     */
  int32_t pcg_max_iterations;          /* cap per linear solve                      */
  double pcg_rtol;                     /* stop when ||b - Ax|| <= rtol * ||b||      */
  int32_t num_threads;                 /* accepted for API parity (thread_num); unused */
  int32_t device;                      /* CUDA ordinal of the (first) device; -1 = current device */
  int32_t verbose;                     /* 0 silent, 1 one line per iteration        */
  int32_t n_gpus;                      /* gsfm_ra_solve only: 0 / 1 = one device; W > 1 = shard the edges over devices
                                          device .. device+W-1 of this process (peer access over NVLink, one host thread per
                                          device); -1 = every visible device when the problem is large enough
                                          (>= GSFM_RA_MIN_EDGES_PER_GPU edges per device), else one                        */
} gsfm_ra_options;
#define GSFM_RA_MIN_EDGES_PER_GPU 500000

typedef enum {
  GSFM_RA_TERM_NONE = 0,
  GSFM_RA_TERM_FUNCTION_TOLERANCE = 1,
  GSFM_RA_TERM_GRADIENT_TOLERANCE = 2,
  GSFM_RA_TERM_PARAMETER_TOLERANCE = 3,
  GSFM_RA_TERM_MAX_ITERATIONS = 4,
  GSFM_RA_TERM_MIN_RADIUS = 5,
  GSFM_RA_TERM_INVALID_STEPS = 6,
  GSFM_RA_TERM_FAILURE = 7
} gsfm_ra_termination;

/* One row per trust-region iteration (iteration 0 = the initial evaluation). */
typedef struct {
  int32_t iteration;
  int32_t step_is_successful;
  int32_t step_is_valid;
  int32_t linear_iterations;   /* PCG iterations of this step's solve            */
  double cost;                 /* cost at the accepted point after this iteration */
  double candidate_cost;
  double cost_change;
  double model_cost_change;
  double relative_decrease;
  double gradient_max_norm;
  double step_norm;
  double trust_region_radius;  /* radius after the update                        */
  double linear_residual;      /* ||b-Ax||/||b|| reached by the solve            */
} gsfm_ra_iteration;

typedef struct {
  int32_t termination;            /* gsfm_ra_termination                           */
  int32_t num_iterations;         /* trust-region iterations executed (excl. it.0) */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int64_t total_linear_iterations;
  double initial_cost;
  double final_cost;
  /* device time, CUDA events on the solver stream, milliseconds */
  double ms_setup;     /* upload + CSR structure + whitening                        */
  double ms_assemble;  /* K1: residual+Jacobian+loss+normal equations (all its.)    */
  double ms_linear;    /* K2/K3: PCG incl. SpMV (all iterations)                    */
  double ms_cost;      /* K1c: trial-point cost                                     */
  double ms_total;     /* whole call, host wall clock                               */
  int64_t kernel_launches; /* kernels launched by this call                         */
  /* optional trace: caller-provided array of trace_capacity rows, may be NULL     */
  gsfm_ra_iteration* trace;
  int32_t trace_capacity;
  int32_t trace_size;
  /* sigma-consensus only: outer re-weighting iterations executed and the last mean |w - w_prev| */
  int32_t outer_iterations;
  int32_t num_linear_unconverged; /* PCG solves that stopped at pcg_max_iterations above pcg_rtol (their steps were still used) */
  double last_weight_change;
  int32_t n_gpus_used;            /* devices the solve ran on (1 unless options.n_gpus asked for more and the problem qualified) */
  int32_t reserved;
} gsfm_ra_summary;

typedef struct gsfm_ra_solver gsfm_ra_solver; /* opaque, device-resident problem */

/* ---- library ---------------------------------------------------------------*/
int gsfm_ra_abi_version(void);
/* Human-readable message of the last failure on the calling thread. */
const char* gsfm_ra_last_error(void);
/* Number of visible CUDA devices (0 if none / driver missing). */
int gsfm_ra_device_count(void);
void gsfm_ra_default_options(gsfm_ra_options* options);

/* ---- one-shot solve: replaces ceres::Solve at rotation_estimator.cpp:76-77,
 *      181-182, 304-305 for the angle-axis entry points.  omega_inout [N][3] is the
 *      initial guess on entry and the estimate on exit (the reference updates the
 *      caller's unordered_map<ViewId,Vector3d> in place).                      */
int gsfm_ra_solve(const gsfm_ra_problem* problem, const gsfm_ra_options* options,
                  double* omega_inout, gsfm_ra_summary* summary);

/* ---- sigma-consensus outer loop: replaces EstimateRotationsWithSigmaConsensus
 *      (rotation_estimator.cpp:314-457).  Up to iters_num times: per edge w = (C3*2/sigma_max) *
 *      (Gamma_table[round(1000 r^2 / (2 sigma_max^2))] - Gamma_k) from the angular residual r at the current
 *      rotations (weight_zero below DBL_EPSILON), a full trust-region solve of PairwiseRotationError(omega_ij, w)
 *      under options->loss, stop once mean |w - w_prev| <= 1e-7 (checked after the solve, as the reference does).  The
 *      caller's trace buffer receives the rows of every inner solve one after the other.
 *      problem->error_type must be GSFM_RA_ANGLE_AXIS; problem->edge_weight is ignored.                        */
int gsfm_ra_solve_sigma_consensus(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int32_t iters_num,
                                  double sigma_max, double* omega_inout, gsfm_ra_summary* summary);

/* ---- resident solver ---------------------------------------------------------*/
int gsfm_ra_solver_create(const gsfm_ra_problem* problem, const gsfm_ra_options* options,
                          gsfm_ra_solver** out);
/* Edge-sharded variant (one process per GPU): every rank passes the SAME full problem and
 * keeps the contiguous slice [E*rank/world, E*(rank+1)/world) of its edge list resident.  All ranks hold
 * all N rotations and every PCG vector; they exchange one all-reduce of the 9N+2 per-view sums per outer
 * iteration and one all-reduce of the 3N partial matvec per CG step.  world_size==1 is the plain solver.
 * A sharded solver must be connected with gsfm_ra_solver_comm_init before it iterates.             */
int gsfm_ra_solver_create_sharded(const gsfm_ra_problem* problem, const gsfm_ra_options* options,
                                  int32_t rank, int32_t world_size, gsfm_ra_solver** out);
void gsfm_ra_solver_destroy(gsfm_ra_solver* solver);
int gsfm_ra_solver_set_rotations(gsfm_ra_solver* solver, const double* omega /*[N][3] host*/);
int gsfm_ra_solver_get_rotations(gsfm_ra_solver* solver, double* omega /*[N][3] host*/);
/* Restart the trust region (radius, Jacobi scaling, iteration counter). */
int gsfm_ra_solver_reset(gsfm_ra_solver* solver);
/* Run at most num_iterations further trust-region iterations from the current
 * state (stops earlier on convergence unless tolerances are <= 0).            */
int gsfm_ra_solver_iterate(gsfm_ra_solver* solver, int32_t num_iterations, gsfm_ra_summary* summary);
/* Exchange set-up for the sharded solver (NCCL over NVLink; libnccl.so.2 is bound at run time, so
 * the single-GPU path has no dependency on it).  Rank 0 creates a 128-byte id, the host framework
 * broadcasts it (e.g. torch.distributed), every rank then joins -- collective, all ranks must call. */
#define GSFM_RA_COMM_ID_BYTES 128
int gsfm_ra_comm_unique_id(uint8_t* id /*[GSFM_RA_COMM_ID_BYTES]*/);
int gsfm_ra_solver_comm_init(gsfm_ra_solver* solver, const uint8_t* id /*[GSFM_RA_COMM_ID_BYTES]*/);
/* Fused exchange (optional, after comm_init): the per-CG-step reduction then runs INSIDE the persistent PCG
 * kernel over peer memory (CUDA IPC + NVLink loads, system-scope flags) instead of as an NCCL call between
 * kernels.  Each rank exports one opaque handle; the host framework all-gathers them; every rank imports all.
 * Without it the sharded solver uses ncclAllReduce per CG step.                                           */
#define GSFM_RA_IPC_HANDLE_BYTES 128
int gsfm_ra_solver_ipc_export(gsfm_ra_solver* solver, uint8_t* handle /*[GSFM_RA_IPC_HANDLE_BYTES]*/);
int gsfm_ra_solver_ipc_import(gsfm_ra_solver* solver, const uint8_t* handles /*[world][GSFM_RA_IPC_HANDLE_BYTES]*/);
/* Edge range of the caller's list owned by this rank. */
int gsfm_ra_solver_edge_range(const gsfm_ra_solver* solver, uint64_t* edge_begin, uint64_t* edge_end);

/* The CUDA stream (a cudaStream_t) every kernel of this solver is launched on, so a host
 * framework can bracket calls with its own events.                               */
void* gsfm_ra_solver_cuda_stream(gsfm_ra_solver* solver);
/* Measurement aid: average device time in ms (CUDA events on the solver stream) of `repeats`
 * back-to-back launches of the production kernels on the solver's current linearisation.
 * out_ms[0] K1 fused edge kernel, [1] K1c cost-only edge kernel, [2] K2 block SpMV,
 * [3] one whole PCG iteration (K2 + the three vector kernels).  Does not change solver state. */
int gsfm_ra_solver_time_kernels(gsfm_ra_solver* solver, int32_t repeats, double* out_ms /*[4]*/);

/* What the solver resolved at build time: out[0] linear solver actually used (gsfm_ra_linear_solver, AUTO resolved),
 * [1] bytes stored per half-edge of the block matrix (36 compact scalar stencil, 52 symmetric, 76 general),
 * [2] persistent-kernel grid size, [3] threads per block of the K2-class kernels, [4] L2 keep fraction in eighths,
 * [5] world size, [6] 1 if trust-region batches replay as CUDA graphs, [7] reserved. */
int gsfm_ra_solver_info(const gsfm_ra_solver* solver, int32_t* out /*[8]*/);
/* Measurement aid (roofline denominators): GB/s reached by the K2 record stream alone -- the same per-warp TMA rings pulling
 * `bytes` of device memory `repeats` times with no arithmetic.  bytes below the L2 size measures the L2 -> SM stream rate,
 * far above it the HBM stream rate of this access pattern. */
int gsfm_ra_measure_stream(uint64_t bytes, int32_t repeats, int32_t device, double* gb_per_s);

/* ---- kernel-level entry points (parity tests, profiling) ---------------------*/
/* Residual dimension of an error type: 4 for QUATERNION_NORM (include/pairwise_rotation_error_quat.hpp:125-150),
 * 9 for ROTATION_MAT_FNORM (:169-196), 3 otherwise. */
int gsfm_ra_residual_dim(int32_t error_type);

/* K1 per edge, at omega [N][3]: raw residual r [E][d], raw Jacobians w.r.t. the parameters Ceres optimises
 * (the angle-axis vector for types 3..8; the 3 local coordinates of EigenQuaternionParameterization for types 0..2)
 * d r/d x_i, d r/d x_j [E][d][3] row-major, d = gsfm_ra_residual_dim(error_type), and rho[E][3] = loss at |r|^2.
 * Any output pointer may be NULL.  Replaces one AutoDiffCostFunction::Evaluate +
 * LossFunction::Evaluate per edge (src/pairwise_rotation_error.cpp:75-85,
 * bind_src/GlobalSfMpy.cpp:36-59).                                             */
int gsfm_ra_eval_edges(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega,
                       double* r, double* jac_i, double* jac_j, double* rho, int32_t device);
/* Whitening only: U [E][9] row-major upper-triangular (rotation_estimator.cpp:252-255). */
int gsfm_ra_whiten(const gsfm_ra_problem* problem, double* U, int32_t device);
/* K1 fused: robustified normal equations at omega. cost = sum 1/2 rho; gradient
 * g [N][3] = J~^T r~; hdiag [N][9] diagonal blocks of J~^T J~ (row-major); the
 * off-diagonal blocks in the solver's block-CSR order: rowptr [N+1], col [nnzb],
 * val [nnzb][9] row-major (nnzb = 2E).  Any output may be NULL.  No Jacobi
 * scaling, no damping.                                                          */
int gsfm_ra_assemble(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega,
                     double* cost, double* gradient, double* hdiag,
                     uint32_t* rowptr, uint32_t* col, double* val, int32_t device);
/* K1c: cost only. */
int gsfm_ra_cost(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega,
                 double* cost, int32_t device);
/* K2 on the matrix assembled at omega: y = (H + diag(damping)) x, x,y,damping [N][3];
 * damping may be NULL.                                                           */
int gsfm_ra_spmv(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega,
                 const double* damping, const double* x, double* y, int32_t device);
/* K2+K3: solve (H + diag(damping)) x = b with the production PCG; returns the
 * iteration count in *iterations and the relative residual in *rel_residual.   */
int gsfm_ra_pcg(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega,
                const double* damping, const double* b, double rtol, int32_t max_iterations,
                double* x, int32_t* iterations, double* rel_residual, int32_t device);
/* Loss table on the device: out [n][3] = (rho, rho', rho'') at s[i]. Replaces
 * LossFunction::Evaluate of scripts/loss_functions.py.                          */
int gsfm_ra_eval_loss(const gsfm_ra_loss* loss, const double* s, uint64_t n, double* out, int32_t device);

/* ---- the step after the path: FilterViewPairsFromOrientation
 *      (thirdparty/TheiaSfM/src/theia/sfm/filter_view_pairs_from_orientation.cc:55-118).
 *      keep[k] = 1 iff |Log(R_ij^T R_j R_i^T)| <= max_degrees.                 */
int gsfm_ra_filter_view_pairs(const gsfm_ra_problem* problem, const double* omega,
                              double max_relative_rotation_difference_degrees,
                              uint8_t* keep, double* angle_rad /*may be NULL*/, int32_t device);

/* ---- view-graph shaping on the device: the steps immediately before the solve ----------------- */

/* FilterInitialViewGraph (reference src/GSfM_global_reconstruction_estimator.cpp:369-390): drop view pairs with
 * num_verified_matches < min_num_two_view_inliers, then keep the largest connected component
 * (RemoveDisconnectedViewPairs, thirdparty/TheiaSfM/src/theia/sfm/view_graph/remove_disconnected_view_pairs.cc:48-;
 * ties between equally large components: the one holding the smallest view index).  Views are dense indices
 * 0..num_views-1.  edge_keep[E] / view_keep[num_views] receive 0/1. */
int gsfm_ra_filter_initial_view_graph(uint32_t num_views, uint64_t num_edges, const uint32_t* edge_i, const uint32_t* edge_j,
                                      const int32_t* num_verified_matches, int32_t min_num_two_view_inliers, uint8_t* edge_keep,
                                      uint8_t* view_keep, int32_t device);

/* OrientationsFromMaximumSpanningTree
 * (thirdparty/TheiaSfM/src/theia/sfm/view_graph/orientations_from_maximum_spanning_tree.cc:109-178): maximum spanning
 * tree over edge_weight (the reference: num_verified_matches; Kruskal, theia/math/graph/minimum_spanning_tree.h:70-98;
 * ties broken by (edge_i, edge_j) so that the tree is unique), then R_neighbor = (src < nbr ? R_rel : R_rel^T) R_src
 * from the root (:60-83).  root < 0: the smallest view index that has an edge.  omega_out[3 * num_views]: angle-axis,
 * NaN for views the root cannot reach.  edge_in_tree (E, optional) flags the tree edges, rounds_out (optional) the
 * number of Boruvka rounds. */
int gsfm_ra_init_orientations_mst(uint32_t num_views, uint64_t num_edges, const uint32_t* edge_i, const uint32_t* edge_j,
                                  const double* omega_ij, const int32_t* edge_weight, int64_t root, double* omega_out,
                                  uint8_t* edge_in_tree, int32_t* rounds_out, int32_t device);

/* ---- on-disk formats either side of the path; host-only code, no CUDA device needed ---------------- */

/* Arrays returned by the readers below live in malloc'ed memory: release each with gsfm_ra_free. */
void gsfm_ra_free(void* p);

/* covariance_rot.txt  (reference src/uncertainty.cpp:200-229 read_covariance; :164-198 store_covariance_rot): two header
 * lines, then per view pair `id1 id2` and 9 doubles written as the decimal of their uint64 bit pattern:
 * C00 C11 C22 C01 C02 C12 (cov6, the layout gsfm_ra_problem.cov6 takes) and R0 R1 R2 (the pair's rotation). */
int gsfm_ra_read_covariance_rot(const char* path, uint64_t* count, uint32_t** view_id1, uint32_t** view_id2, double** cov6,
                                double** rot);
int gsfm_ra_write_covariance_rot(const char* path, uint64_t count, const uint32_t* view_id1, const uint32_t* view_id2,
                                 const double* cov6, const double* rot);

/* A 1DSfM dataset directory (thirdparty/TheiaSfM/src/theia/io/read_1dsfm.cc:93-412: cc.txt, list.txt, tracks.txt, EGs.txt).
 * Views: the ids of cc.txt that list.txt defines (line index = id) with their focal-length priors (0 = none).  Pairs: the
 * EGs.txt entries between those views, rotation_2 = angle-axis of S R^T S and position_2 = S t with S = diag(1,-1,-1)
 * (:309-333), num_verified_matches = number of tracks that see both views.  num_listed_views, focal_length_priors,
 * position_2 and num_verified_matches may be NULL (tracks.txt is then not read). */
int gsfm_ra_read_1dsfm(const char* dataset_directory, uint32_t* num_listed_views, uint64_t* num_views, uint32_t** view_ids,
                       double** focal_length_priors, uint64_t* num_pairs, uint32_t** view_id1, uint32_t** view_id2,
                       double** rotation_2, double** position_2, int32_t** num_verified_matches);

#ifdef __cplusplus
}
#endif
#endif /* GSFM_RA_H_ */
