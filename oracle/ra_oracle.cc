/*
 * ra_oracle.cc -- CPU ORACLE for the rotation-averaging hot path.  TEST INFRASTRUCTURE ONLY.
 * See ra_oracle.h for who may load this library and for the parity status ("UNPINNED" for
 * the converged solve: no Ceres binary can be built here, the trust-region loop restates
 * Ceres Solver 1.14.0 from its published algorithm).
 *
 * Every function cites the reference lines it restates.  Paths are relative to the
 * reference checkout; T/ = thirdparty/TheiaSfM/src/theia/.
 *
 * The residual/Jacobian evaluation deliberately mirrors HOW the reference computes them:
 * forward-mode dual numbers ("jets", 6 infinitesimals = the two 3-vector parameter
 * blocks) pushed through Rodrigues -> matrix products -> matrix->quaternion -> angle-axis,
 * exactly the chain ceres::AutoDiffCostFunction<PairwiseRotationErrorAngleAxis,3,3,3>
 * evaluates (src/pairwise_rotation_error.cpp:75-85).  The CUDA product uses closed-form
 * SO(3) Jacobians instead; agreement of the two is what the parity tests check.
 */
#include "ra_oracle.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ------------------------------------------------------------------------------------ */
/* forward-mode dual numbers, the role ceres::Jet<double,6> plays in the reference       */
/* ------------------------------------------------------------------------------------ */
constexpr int kN = 8; /* 3+3 angle-axis blocks use the first six; the quaternion functor needs 4+4 */
struct Jet {
  double a;
  double v[kN];
  Jet() : a(0) { for (int k = 0; k < kN; ++k) v[k] = 0; }
  Jet(double x) : a(x) { for (int k = 0; k < kN; ++k) v[k] = 0; }  // NOLINT
};
inline Jet operator+(const Jet& x, const Jet& y) { Jet r; r.a = x.a + y.a; for (int k = 0; k < kN; ++k) r.v[k] = x.v[k] + y.v[k]; return r; }
inline Jet operator-(const Jet& x, const Jet& y) { Jet r; r.a = x.a - y.a; for (int k = 0; k < kN; ++k) r.v[k] = x.v[k] - y.v[k]; return r; }
inline Jet operator-(const Jet& x) { Jet r; r.a = -x.a; for (int k = 0; k < kN; ++k) r.v[k] = -x.v[k]; return r; }
inline Jet operator*(const Jet& x, const Jet& y) { Jet r; r.a = x.a * y.a; for (int k = 0; k < kN; ++k) r.v[k] = x.a * y.v[k] + x.v[k] * y.a; return r; }
inline Jet operator/(const Jet& x, const Jet& y) {
  Jet r; const double inv = 1.0 / y.a; r.a = x.a * inv;
  for (int k = 0; k < kN; ++k) r.v[k] = (x.v[k] - r.a * y.v[k]) * inv;
  return r;
}
inline Jet sqrt(const Jet& x) { Jet r; r.a = std::sqrt(x.a); const double d = 1.0 / (2.0 * r.a); for (int k = 0; k < kN; ++k) r.v[k] = x.v[k] * d; return r; }
inline Jet sin(const Jet& x) { Jet r; r.a = std::sin(x.a); const double c = std::cos(x.a); for (int k = 0; k < kN; ++k) r.v[k] = c * x.v[k]; return r; }
inline Jet cos(const Jet& x) { Jet r; r.a = std::cos(x.a); const double s = -std::sin(x.a); for (int k = 0; k < kN; ++k) r.v[k] = s * x.v[k]; return r; }
inline Jet atan2(const Jet& y, const Jet& x) {
  Jet r; r.a = std::atan2(y.a, x.a); const double d = 1.0 / (x.a * x.a + y.a * y.a);
  for (int k = 0; k < kN; ++k) r.v[k] = (x.a * y.v[k] - y.a * x.v[k]) * d;
  return r;
}
inline double val(const Jet& x) { return x.a; }
inline double val(double x) { return x; }
using std::atan2;
using std::cos;
using std::sin;
using std::sqrt;

/* 3x3 stored M[r][c] */
template <typename T>
struct Mat3 { T m[3][3]; };

/* ceres/rotation.h AngleAxisToRotationMatrix (Ceres 1.14):
 * Rodrigues with a unit axis when theta^2 > DBL_EPSILON, first-order I+[w]x otherwise. */
template <typename T>
void AngleAxisToMatrix(const T* w, Mat3<T>* R) {
  const T theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (val(theta2) > DBL_EPSILON) {
    const T theta = sqrt(theta2);
    const T wx = w[0] / theta, wy = w[1] / theta, wz = w[2] / theta;
    const T c = cos(theta), s = sin(theta);
    const T one_c = T(1.0) - c;
    R->m[0][0] = c + wx * wx * one_c;
    R->m[1][0] = wz * s + wx * wy * one_c;
    R->m[2][0] = -wy * s + wx * wz * one_c;
    R->m[0][1] = wx * wy * one_c - wz * s;
    R->m[1][1] = c + wy * wy * one_c;
    R->m[2][1] = wx * s + wy * wz * one_c;
    R->m[0][2] = wy * s + wx * wz * one_c;
    R->m[1][2] = -wx * s + wy * wz * one_c;
    R->m[2][2] = c + wz * wz * one_c;
  } else {
    R->m[0][0] = T(1.0); R->m[1][0] = w[2];   R->m[2][0] = -w[1];
    R->m[0][1] = -w[2];  R->m[1][1] = T(1.0); R->m[2][1] = w[0];
    R->m[0][2] = w[1];   R->m[1][2] = -w[0];  R->m[2][2] = T(1.0);
  }
}

/* ceres/rotation.h RotationMatrixToQuaternion + QuaternionToAngleAxis, which is what
 * RotationMatrixToAngleAxis is in Ceres 1.14. */
template <typename T>
void MatrixToAngleAxis(const Mat3<T>& R, T* w) {
  T q[4];
  const T trace = R.m[0][0] + R.m[1][1] + R.m[2][2];
  if (val(trace) >= 0.0) {
    T t = sqrt(trace + T(1.0));
    q[0] = T(0.5) * t;
    t = T(0.5) / t;
    q[1] = (R.m[2][1] - R.m[1][2]) * t;
    q[2] = (R.m[0][2] - R.m[2][0]) * t;
    q[3] = (R.m[1][0] - R.m[0][1]) * t;
  } else {
    int i = 0;
    if (val(R.m[1][1]) > val(R.m[0][0])) i = 1;
    if (val(R.m[2][2]) > val(R.m[i][i])) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    T t = sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + T(1.0));
    q[i + 1] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (R.m[k][j] - R.m[j][k]) * t;
    q[j + 1] = (R.m[j][i] + R.m[i][j]) * t;
    q[k + 1] = (R.m[k][i] + R.m[i][k]) * t;
  }
  const T sin2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (val(sin2) > 0.0) {
    const T s = sqrt(sin2);
    const T& c = q[0];
    const T two_theta = T(2.0) * ((val(c) < 0.0) ? atan2(-s, -c) : atan2(s, c));
    const T k = two_theta / s;
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  } else {
    const T k(2.0);
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  }
}

/* The functor body of include/pairwise_rotation_error_quat.hpp:215-247 (U = Lt) and of
 * T/sfm/global_pose_estimation/pairwise_rotation_error.h:66-95 (U = weight * I):
 *   loop = R2 * R1^T ; err = loop * R12^T ; e = Log(err) ; residual = U * e            */
template <typename T>
void EdgeResidual(const T* w1, const T* w2, const double* w12, const double* U, T* res) {
  Mat3<T> R1, R2;
  Mat3<double> R12;
  AngleAxisToMatrix(w1, &R1);
  AngleAxisToMatrix(w2, &R2);
  AngleAxisToMatrix(w12, &R12);
  Mat3<T> loop, err;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      loop.m[r][c] = R2.m[r][0] * R1.m[c][0] + R2.m[r][1] * R1.m[c][1] + R2.m[r][2] * R1.m[c][2];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      err.m[r][c] = loop.m[r][0] * T(R12.m[c][0]) + loop.m[r][1] * T(R12.m[c][1]) + loop.m[r][2] * T(R12.m[c][2]);
  T e[3];
  MatrixToAngleAxis(err, e);
  for (int r = 0; r < 3; ++r) res[r] = T(U[3 * r + 0]) * e[0] + T(U[3 * r + 1]) * e[1] + T(U[3 * r + 2]) * e[2];
}

/* include/pairwise_rotation_error_quat.hpp:82-106 (PairwiseRotationErrorQuat): parameter blocks are Eigen
 * quaternions in coefficient order (x,y,z,w); delta_q = q_rel * conj(q_b * conj(q_a)); residual = 2 w delta_q.vec(). */
template <typename T>
void QMul(const T* a, const T* b, T* o) { /* Hamilton product, (x,y,z,w) storage */
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
template <typename T>
void QuatCosineResidual(const T* qa, const T* qb, const double* qrel, double weight, T* res) {
  const T qa_inv[4] = {-qa[0], -qa[1], -qa[2], qa[3]};
  T est[4];
  QMul(qb, qa_inv, est);
  const T est_conj[4] = {-est[0], -est[1], -est[2], est[3]};
  const T rel[4] = {T(qrel[0]), T(qrel[1]), T(qrel[2]), T(qrel[3])};
  T dq[4];
  QMul(rel, est_conj, dq);
  for (int k = 0; k < 3; ++k) res[k] = T(weight) * T(2.0) * dq[k];
}
/* include/pairwise_rotation_error_quat.hpp:125-150 (PairwiseRotationErrorQuatFNorm): q_b_estimated = q_rel * q_a; both q_b and
 * the estimate are negated when their coeffs()[1] (the y component in Eigen's x,y,z,w order) is negative; 4 residuals. */
template <typename T>
void QuatFNormResidual(const T* qa, const T* qb, const double* qrel, double weight, T* res) {
  const T rel[4] = {T(qrel[0]), T(qrel[1]), T(qrel[2]), T(qrel[3])};
  T est[4];
  QMul(rel, qa, est);
  const bool flip_b = val(qb[1]) < 0.0, flip_e = val(est[1]) < 0.0;
  for (int k = 0; k < 4; ++k) {
    const T b = flip_b ? -qb[k] : qb[k];
    const T e = flip_e ? -est[k] : est[k];
    res[k] = T(weight) * (b - e);
  }
}
/* Eigen::QuaternionBase::toRotationMatrix (no normalisation), coefficients (x,y,z,w) */
template <typename T>
void EigenQuatToMatrix(const T* q, Mat3<T>* R) {
  const T tx = T(2.0) * q[0], ty = T(2.0) * q[1], tz = T(2.0) * q[2];
  const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R->m[0][0] = T(1.0) - (tyy + tzz); R->m[0][1] = txy - twz;          R->m[0][2] = txz + twy;
  R->m[1][0] = txy + twz;          R->m[1][1] = T(1.0) - (txx + tzz); R->m[1][2] = tyz - twx;
  R->m[2][0] = txz - twy;          R->m[2][1] = tyz + twx;          R->m[2][2] = T(1.0) - (txx + tyy);
}
/* include/pairwise_rotation_error_quat.hpp:169-196 (PairwiseRotationErrorRotFNorm): residual(k) = w (R_rel R_a - R_b)(k) with
 * Eigen's linear (column-major) index k = 0..8. */
template <typename T>
void RotFNormResidual(const T* qa, const T* qb, const double* qrel, double weight, T* res) {
  Mat3<T> Ra, Rb;
  Mat3<double> Rr;
  EigenQuatToMatrix(qa, &Ra);
  EigenQuatToMatrix(qb, &Rb);
  EigenQuatToMatrix(qrel, &Rr);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      T est = T(Rr.m[r][0]) * Ra.m[0][c] + T(Rr.m[r][1]) * Ra.m[1][c] + T(Rr.m[r][2]) * Ra.m[2][c];
      res[3 * c + r] = T(weight) * (est - Rb.m[r][c]);
    }
}
/* ceres AngleAxisToQuaternion, output in Eigen coefficient order (x,y,z,w) as rotation_estimator.cpp:127-136 builds it */
void AngleAxisToQuatXYZW(const double* w, double* q) {
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (t2 > 0.0) {
    const double t = std::sqrt(t2), half = t * 0.5, k = std::sin(half) / t;
    q[3] = std::cos(half); q[0] = w[0] * k; q[1] = w[1] * k; q[2] = w[2] * k;
  } else {
    q[3] = 1.0; q[0] = w[0] * 0.5; q[1] = w[1] * 0.5; q[2] = w[2] * 0.5;
  }
}
/* ceres QuaternionToAngleAxis on (w,x,y,z) = (q[3],q[0],q[1],q[2]) (rotation_estimator.cpp:185-194) */
void QuatXYZWToAngleAxis(const double* q, double* w) {
  const double sin2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  if (sin2 > 0.0) {
    const double s = std::sqrt(sin2), c = q[3];
    const double two_theta = 2.0 * ((c < 0.0) ? std::atan2(-s, -c) : std::atan2(s, c));
    const double k = two_theta / s;
    w[0] = q[0] * k; w[1] = q[1] * k; w[2] = q[2] * k;
  } else {
    w[0] = q[0] * 2.0; w[1] = q[1] * 2.0; w[2] = q[2] * 2.0;
  }
}
/* ceres EigenQuaternionParameterization: Plus and its 4x3 Jacobian at delta = 0 */
void QuatPlus(const double* x, const double* d, double* out) {
  const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double k = std::sin(n) / n;
    const double qd[4] = {k * d[0], k * d[1], k * d[2], std::cos(n)};
    QMul(qd, x, out);
  } else {
    for (int k = 0; k < 4; ++k) out[k] = x[k];
  }
}
void QuatPlusJacobian(const double* x, double* J /*4x3 row-major*/) {
  J[0] = x[3];  J[1] = x[2];   J[2] = -x[1];
  J[3] = -x[2]; J[4] = x[3];   J[5] = x[0];
  J[6] = x[1];  J[7] = -x[0];  J[8] = x[3];
  J[9] = -x[0]; J[10] = -x[1]; J[11] = -x[2];
}
/* one quaternion edge through jets; Jacobians returned in the 3-dim LOCAL (tangent) coordinates Ceres optimises in:
 * J_local = J_ambient(dim x 4) * PlusJacobian(4x3).  type selects the functor (0: QuatFNorm, 1: RotFNorm, 2: Quat). */
/* theia::PairwiseTranslationError::operator()
 * (thirdparty/TheiaSfM/src/theia/sfm/global_pose_estimation/pairwise_translation_error.h:62-88), as AutoDiffCostFunction
 * <PairwiseTranslationError, 3, 3, 3> runs it: below kNormTolerance the norm is REPLACED by the constant T(1.0), so the dual
 * part of the norm (0 * inf = NaN at coincident positions) is discarded with it.                                        */
template <typename T>
void TranslationResidual(const T* position1, const T* position2, const double* translation_direction, double weight, T* residuals) {
  const double kNormTolerance = 1e-12;
  T translation[3];
  translation[0] = position2[0] - position1[0];
  translation[1] = position2[1] - position1[1];
  translation[2] = position2[2] - position1[2];
  T norm = sqrt(translation[0] * translation[0] + translation[1] * translation[1] + translation[2] * translation[2]);
  if (val(norm) < kNormTolerance) norm = T(1.0);
  residuals[0] = T(weight) * (translation[0] / norm - T(translation_direction[0]));
  residuals[1] = T(weight) * (translation[1] / norm - T(translation_direction[1]));
  residuals[2] = T(weight) * (translation[2] / norm - T(translation_direction[2]));
}

/* GetRotatedTranslation, src/GSfM_nonlinear_position_estimator.cpp:36-44: rotation.transpose() * translation */
void RotatedTranslation(const double* orientation, const double* translation, double* out) {
  Mat3<double> R;
  AngleAxisToMatrix(orientation, &R);
  for (int c = 0; c < 3; ++c) out[c] = R.m[0][c] * translation[0] + R.m[1][c] * translation[1] + R.m[2][c] * translation[2];
}

int ResidualDim(int type) { return type == GSFM_RA_QUATERNION_NORM ? 4 : (type == GSFM_RA_ROTATION_MAT_FNORM ? 9 : 3); }
void QuatEdge(int type, const double* qa, const double* qb, const double* qrel, double weight, double* r, double* Ji, double* Jj) {
  Jet a[4], b[4], res[9];
  for (int k = 0; k < 4; ++k) { a[k] = Jet(qa[k]); a[k].v[k] = 1.0; b[k] = Jet(qb[k]); b[k].v[4 + k] = 1.0; }
  if (type == GSFM_RA_QUATERNION_NORM) QuatFNormResidual<Jet>(a, b, qrel, weight, res);
  else if (type == GSFM_RA_ROTATION_MAT_FNORM) RotFNormResidual<Jet>(a, b, qrel, weight, res);
  else QuatCosineResidual<Jet>(a, b, qrel, weight, res);
  const int dim = ResidualDim(type);
  double Pa[12], Pb[12];
  QuatPlusJacobian(qa, Pa);
  QuatPlusJacobian(qb, Pb);
  for (int q = 0; q < dim; ++q) {
    if (r) r[q] = res[q].a;
    for (int c = 0; c < 3; ++c) {
      double si = 0, sj = 0;
      for (int m = 0; m < 4; ++m) { si += res[q].v[m] * Pa[3 * m + c]; sj += res[q].v[4 + m] * Pb[3 * m + c]; }
      if (Ji) Ji[3 * q + c] = si;
      if (Jj) Jj[3 * q + c] = sj;
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* MAGSAC constants, include/gamma_values.cpp:6-11, 384-389, 780-785 (numbers, not code)  */
/* ------------------------------------------------------------------------------------ */
struct MagsacConst { double nu, C, quantile, gamma_k; int table_size; };
const MagsacConst kMagsac3 = {3.0, 4.029720004054876e-01, 3.368214175218727, 3.439485560754856e-03, 36843};
const MagsacConst kMagsac4 = {4.0, 2.525252525252525e-01, 3.643721193503644e+00, 3.611260617758625e-03, 38683};
const MagsacConst kMagsac9 = {9.0, 3.837828575290349e-03, 4.654674460524809e+00, 3.344206155099048e-02, 48553};
constexpr double kGammaPrecision = 1000.0; /* precision_of_stored_gamma{3,4,9} */

/* stored_gamma_values{nu}[i] = Gamma((nu-1)/2, i/1000) (upper incomplete); closed forms
 * (verified against the table to <= 1e-14 abs in the golden test). */
double GammaTable(int nu, int index) {
  const double x = index / kGammaPrecision;
  switch (nu) {
    case 3: return std::exp(-x);
    case 4: return 0.5 * std::sqrt(M_PI) * std::erfc(std::sqrt(x)) + std::sqrt(x) * std::exp(-x);
    case 9: return std::exp(-x) * (((x + 3.0) * x + 6.0) * x + 6.0);
  }
  return NAN;
}

/* Python's round(): half to even (scripts/loss_functions.py:309 uses the builtin). */
inline double RoundHalfEven(double x) { return std::nearbyint(x); /* default FE_TONEAREST */ }

/* scripts/loss_functions.py:285-341 (nu=3), 344-400 (nu=4), 402-459 (nu=9). */
void MagsacLoss(const MagsacConst& K, double sigma, bool inverse, double s_in, double* rho) {
  const double squared_sigma = sigma * sigma;
  const double squared_sigma_max_2 = 2.0 * squared_sigma;
  const double cubed_sigma_max = squared_sigma * sigma;
  const double dof_minus_one_per_two = (K.nu - 1.0) / 2.0;
  const double C_times_two_ad_dof = K.C * std::pow(2.0, dof_minus_one_per_two);
  const double one_over_sigma = C_times_two_ad_dof / sigma;
  const double gamma_value = std::tgamma(dof_minus_one_per_two);
  const double gamma_difference = gamma_value - K.gamma_k;
  const double weight_zero = one_over_sigma * gamma_difference;

  double squared_residual = s_in;
  bool zero_derivative = false;
  if (squared_residual > K.quantile * K.quantile * squared_sigma) {
    squared_residual = K.quantile * K.quantile * squared_sigma;
    zero_derivative = true;
  }
  double xr = RoundHalfEven(kGammaPrecision * squared_residual / squared_sigma_max_2);
  if (K.table_size < xr) xr = K.table_size;
  const int x = (int)xr;
  double s = x * squared_sigma_max_2 / kGammaPrecision;
  const double weight = one_over_sigma * (GammaTable((int)K.nu, x) - K.gamma_k);
  const double expo = K.nu / 2 - 1.5;
  /* Python: 0.0 ** 0.0 == 1.0, same as C pow */
  const double weight_derivative =
      -C_times_two_ad_dof * std::pow(s / squared_sigma_max_2, expo) * std::exp(-s / squared_sigma_max_2) / (2 * cubed_sigma_max);
  if (s < 1e-7) s = 1e-7;
  const double weight_second_derivative = 2.0 * C_times_two_ad_dof * std::pow(s / squared_sigma_max_2, expo) *
                                          (1.0 / squared_sigma - (K.nu - 3) / s) * std::exp(-s / squared_sigma_max_2) /
                                          (8 * cubed_sigma_max);
  if (inverse) {
    rho[0] = 1.0 / weight;
    rho[1] = -1.0 / (weight * weight) * weight_derivative;
    rho[2] = 2.0 / (weight * weight * weight) * weight_derivative * weight_derivative - weight_second_derivative / (weight * weight);
    if (zero_derivative) { rho[1] = 0.00001; rho[2] = 0.0; }
  } else {
    rho[0] = weight_zero - weight;
    rho[1] = -weight_derivative;
    rho[2] = -weight_second_derivative;
    if (rho[1] == 0) rho[1] = 0.00001;
    if (zero_derivative) { rho[1] = 0.00001; rho[2] = 0.0; }
  }
}

/* scripts/loss_functions.py, unscaled losses; lines per gsfm_ra_loss_kind in gsfm_ra.h. */
void BaseLoss(const gsfm_ra_loss* L, double s, double* out) {
  const double* p = L->p;
  switch (L->kind) {
    case GSFM_RA_LOSS_TRIVIAL: out[0] = s; out[1] = 1.0; out[2] = 0.0; return;                   /* :47-54 */
    case GSFM_RA_LOSS_HUBER: {                                                                   /* :56-72 */
      const double a = p[0], b = a * a;
      if (s > b) { const double r = std::sqrt(s); out[0] = 2 * a * r - b; out[1] = std::max(a / r, DBL_MIN); out[2] = -out[1] / (2.0 * s); }
      else { out[0] = s; out[1] = 1.0; out[2] = 0.0; }
      return;
    }
    case GSFM_RA_LOSS_SOFTLONE: {                                                                /* :74-86 */
      const double b = p[0] * p[0], c = 1.0 / b;
      const double sum = 1.0 + s * c, tmp = std::sqrt(sum);
      out[0] = 2.0 * b * (tmp - 1.0); out[1] = std::max(1.0 / tmp, DBL_MIN); out[2] = -(c * out[1]) / (2.0 * sum);
      return;
    }
    case GSFM_RA_LOSS_CAUCHY: {                                                                  /* :88-99 */
      const double b = p[0] * p[0], c = 1.0 / b;
      const double sum = 1.0 + s * c, inv = 1.0 / sum;
      out[0] = b * std::log(sum); out[1] = std::max(inv, DBL_MIN); out[2] = -c * (inv * inv);
      return;
    }
    case GSFM_RA_LOSS_ARCTAN: {                                                                  /* :101-112 */
      const double a = p[0], b = 1 / (a * a);
      const double sum = 1 + s * s * b, inv = 1.0 / sum;
      out[0] = a * std::atan2(s, a); out[1] = std::max(inv, DBL_MIN); out[2] = -2.0 * s * b * (inv * inv);
      return;
    }
    case GSFM_RA_LOSS_TOLERANT: {                                                                /* :114-165 */
      const double a = p[0], b = p[1], c = b * std::log(1 + std::exp(-a / b));
      const double x = (s - a) / b;
      if (x > 36.7) { out[0] = s - a - c; out[1] = 1.0; out[2] = 0.0; }
      else { const double e_x = std::exp(x); out[0] = b * std::log(1.0 + e_x) - c; out[1] = std::max(e_x / (1.0 + e_x), DBL_MIN); out[2] = 0.5 / (b * (1.0 + std::cosh(x))); }
      return;
    }
    case GSFM_RA_LOSS_TUKEY: {                                                                   /* :167-185 */
      const double a2 = p[0] * p[0];
      if (s <= a2) { const double v = 1.0 - s / a2, v2 = v * v; out[0] = a2 / 6.0 * (1.0 - v2 * v); out[1] = 0.5 * v2; out[2] = -1.0 / a2 * v; }
      else { out[0] = a2 / 6.0; out[1] = 0.0; out[2] = 0.0; }
      return;
    }
    case GSFM_RA_LOSS_LONEHALF: {                                                                /* :187-215 */
      const double a = p[0], sa = std::sqrt(a);
      out[0] = 2.0 * a * sa * std::pow(s, 0.25);
      if (s < 0.01) s = 0.01;
      out[1] = 0.5 * std::pow(a, -1.5) * std::pow(s, -0.75);
      out[2] = -0.375 * a * sa * std::pow(s, -1.75);
      return;
    }
    case GSFM_RA_LOSS_LTWO: {                                                                    /* :216-237 */
      const double a2 = p[0] * p[0];
      out[0] = s * s / (a2 * 2.0); out[1] = s / a2; out[2] = 1 / a2;
      return;
    }
    case GSFM_RA_LOSS_GEMANMCCLURE: {                                                            /* :239-248 */
      const double a2 = p[0] * p[0], sg = p[1];
      out[0] = a2 * sg * s / (2.0 * (s + a2 * sg));
      const double d = s / a2 + sg;
      out[1] = (sg * sg) / (2.0 * d * d);
      out[2] = -(sg * sg) / (a2 * d * d * d);
      return;
    }
    case GSFM_RA_LOSS_MAGSAC3: MagsacLoss(kMagsac3, p[0], (L->flags & GSFM_RA_LOSS_FLAG_INVERSE) != 0, s, out); return;
    case GSFM_RA_LOSS_MAGSAC4: MagsacLoss(kMagsac4, p[0], (L->flags & GSFM_RA_LOSS_FLAG_INVERSE) != 0, s, out); return;
    case GSFM_RA_LOSS_MAGSAC9: MagsacLoss(kMagsac9, p[0], (L->flags & GSFM_RA_LOSS_FLAG_INVERSE) != 0, s, out); return;
  }
  out[0] = out[1] = out[2] = NAN;
}

/* src/GSfM_nonlinear_rotation_estimator.cpp:251-288: the per-edge weight matrix U.
 * Eigen's 3x3 inverse() is the cofactor/determinant closed form; llt() reads the lower
 * triangle; U = L^T.  edge_weight multiplies U when the caller provides it (types 5/6
 * pass #matches/100; the reference passes nothing else).                               */
void Whiten(int type, const double* c6, double w, double* U) {
  for (int k = 0; k < 9; ++k) U[k] = 0.0;
  const bool use_cov = (type == GSFM_RA_ANGLE_AXIS_COVARIANCE || type == GSFM_RA_ANGLE_AXIS_COV_INLIERS);
  if (use_cov) {
    const double a = c6[0] * 1e8, d = c6[1] * 1e8, f = c6[2] * 1e8, b = c6[3] * 1e8, c = c6[4] * 1e8, e = c6[5] * 1e8;
    /* cov = [a b c; b d e; c e f]; inverse by cofactors */
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
    const double det = a * c00 + b * c01 + c * c02;
    const double id = 1.0 / det;
    const double P00 = c00 * id, P10 = c01 * id, P20 = c02 * id, P11 = c11 * id, P21 = c12 * id, P22 = c22 * id;
    const double l00 = std::sqrt(P00), l10 = P10 / l00, l20 = P20 / l00;
    const double l11 = std::sqrt(P11 - l10 * l10), l21 = (P21 - l20 * l10) / l11;
    const double l22 = std::sqrt(P22 - l20 * l20 - l21 * l21);
    U[0] = l00 * w; U[1] = l10 * w; U[2] = l20 * w;
    U[4] = l11 * w; U[5] = l21 * w;
    U[8] = l22 * w;
    return;
  }
  double s = w;
  if (type == GSFM_RA_ANGLE_AXIS_COVTRACE) s = w * std::sqrt(1.0 / ((c6[0] + c6[1] + c6[2]) * 1e8));
  if (type == GSFM_RA_ANGLE_AXIS_COVNORM) {
    double n2 = 0;
    for (int k = 0; k < 3; ++k) n2 += (c6[k] * 1e8) * (c6[k] * 1e8);
    for (int k = 3; k < 6; ++k) n2 += 2 * (c6[k] * 1e8) * (c6[k] * 1e8);
    s = w * std::sqrt(1.0 / std::sqrt(n2));
  }
  U[0] = U[4] = U[8] = s;
}

bool TypeNeedsCov(int type) {
  return type == GSFM_RA_ANGLE_AXIS_COVARIANCE || type == GSFM_RA_ANGLE_AXIS_COV_INLIERS ||
         type == GSFM_RA_ANGLE_AXIS_COVTRACE || type == GSFM_RA_ANGLE_AXIS_COVNORM;
}
bool IsPosition(const gsfm_ra_problem* p) { return p->error_type == GSFM_RA_POSITION_BASELINE; }
bool TypeSupported(int type) { return (type >= GSFM_RA_QUATERNION_NORM && type <= GSFM_RA_ANGLE_AXIS_COVNORM) || type == GSFM_RA_POSITION_BASELINE; }
bool IsQuat(const gsfm_ra_problem* p) { return p->error_type <= GSFM_RA_QUATERNION_COSINE; }

void EdgeU(const gsfm_ra_problem* p, uint64_t k, double* U) {
  Whiten(p->error_type, p->cov6 ? p->cov6 + 6 * k : nullptr, p->edge_weight ? p->edge_weight[k] : 1.0, U);
}

struct LossEval {
  const gsfm_ra_loss* loss;
  ra_oracle_loss_cb cb;
  void* ctx;
  void operator()(double s, double* rho) const {
    if (cb) { cb(s, rho, ctx); return; }
    ra_oracle_loss(loss, s, rho);
  }
};

/* One robustified residual block: ceres ResidualBlock::Evaluate + Corrector
 * (SURVEY Appendix B.2).  Jacobian is corrected first, from the uncorrected residual. */
struct EdgeEval {
  int dim;  /* residual dimension: 3, or 4 / 9 for QUATERNION_NORM / ROTATION_MAT_FNORM */
  double r[9], Ji[27], Jj[27], rho[3];
  double sq_norm() const { double s = 0; for (int q = 0; q < dim; ++q) s += r[q] * r[q]; return s; }
};

/* `omega` is the state: [N][3] angle-axis, or [N][4] quaternions (x,y,z,w) for QUATERNION_COSINE */
void EvalEdgeRaw(const gsfm_ra_problem* p, uint64_t k, const double* omega, EdgeEval* out, bool jac) {
  out->dim = ResidualDim(p->error_type);
  if (IsQuat(p)) {
    double qrel[4];
    AngleAxisToQuatXYZW(p->omega_ij + 3 * k, qrel);
    const double w = p->edge_weight ? p->edge_weight[k] : 1.0;  /* cost_weight = 1.0, rotation_estimator.cpp:125 */
    QuatEdge(p->error_type, omega + 4 * (size_t)p->edge_i[k], omega + 4 * (size_t)p->edge_j[k], qrel, w, out->r, jac ? out->Ji : nullptr,
             jac ? out->Jj : nullptr);
    return;
  }
  if (IsPosition(p)) {
    /* translation averaging: `omega` holds camera positions, omega_ij holds TwoViewInfo::position_2
     * (src/GSfM_nonlinear_position_estimator.cpp:298-343, weight 1.0 at :320) */
    double dir[3];
    RotatedTranslation(p->orientation + 3 * (size_t)p->edge_i[k], p->omega_ij + 3 * k, dir);
    const double w = p->edge_weight ? p->edge_weight[k] : 1.0;
    const double* c1 = omega + 3 * (size_t)p->edge_i[k];
    const double* c2 = omega + 3 * (size_t)p->edge_j[k];
    if (!jac) { TranslationResidual<double>(c1, c2, dir, w, out->r); return; }
    Jet a[3], b[3], res[3];
    for (int q = 0; q < 3; ++q) { a[q] = Jet(c1[q]); a[q].v[q] = 1.0; b[q] = Jet(c2[q]); b[q].v[3 + q] = 1.0; }
    TranslationResidual<Jet>(a, b, dir, w, res);
    for (int q = 0; q < 3; ++q) {
      out->r[q] = res[q].a;
      for (int c = 0; c < 3; ++c) { out->Ji[3 * q + c] = res[q].v[c]; out->Jj[3 * q + c] = res[q].v[3 + c]; }
    }
    return;
  }
  double U[9];
  EdgeU(p, k, U);
  const double* wi = omega + 3 * (size_t)p->edge_i[k];
  const double* wj = omega + 3 * (size_t)p->edge_j[k];
  if (jac) {
    ra_oracle_edge(wi, wj, p->omega_ij + 3 * k, U, out->r, out->Ji, out->Jj);
  } else {
    EdgeResidual<double>(wi, wj, p->omega_ij + 3 * k, U, out->r);
  }
}

void Robustify(EdgeEval* e, bool jac) {
  const int dim = e->dim;
  const double s = e->sq_norm();
  const double sqrt_rho1 = std::sqrt(e->rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (s == 0.0 || e->rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1;
    alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * s * e->rho[2] / e->rho[1];
    const double alpha = 1.0 - ((D > 0.0) ? std::sqrt(D) : 0.0);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / s;
  }
  if (jac) {
    for (double* J : {e->Ji, e->Jj}) {
      if (alpha_sq_norm == 0.0) {
        for (int k = 0; k < 3 * dim; ++k) J[k] *= sqrt_rho1;
      } else {
        for (int c = 0; c < 3; ++c) {
          double rtj = 0;
          for (int r = 0; r < dim; ++r) rtj += e->r[r] * J[3 * r + c];
          for (int r = 0; r < dim; ++r) J[3 * r + c] = sqrt_rho1 * (J[3 * r + c] - alpha_sq_norm * e->r[r] * rtj);
        }
      }
    }
  }
  for (int k = 0; k < dim; ++k) e->r[k] *= residual_scaling;
}

int Threads(int n) {
#ifdef _OPENMP
  if (n <= 0) n = omp_get_max_threads();
  return std::max(1, n);
#else
  (void)n;
  return 1;
#endif
}

/* Full-storage off-diagonal block-CSR structure of J^T J: row a lists every neighbour b
 * (sorted); slot_ij[k] / slot_ji[k] locate the two blocks of edge k.                    */
struct Structure {
  std::vector<uint32_t> rowptr, col;
  std::vector<uint64_t> slot_ij, slot_ji;
};
void BuildStructure(const gsfm_ra_problem* p, Structure* S) {
  const uint32_t N = p->num_views;
  const uint64_t E = p->num_edges;
  S->rowptr.assign(N + 1, 0);
  for (uint64_t k = 0; k < E; ++k) { S->rowptr[p->edge_i[k] + 1]++; S->rowptr[p->edge_j[k] + 1]++; }
  for (uint32_t a = 0; a < N; ++a) S->rowptr[a + 1] += S->rowptr[a];
  std::vector<std::pair<uint32_t, uint64_t>> ent(2 * E); /* (col, edge*2+side) */
  std::vector<uint32_t> fill(S->rowptr.begin(), S->rowptr.end() - 1);
  for (uint64_t k = 0; k < E; ++k) {
    ent[fill[p->edge_i[k]]++] = {p->edge_j[k], 2 * k};
    ent[fill[p->edge_j[k]]++] = {p->edge_i[k], 2 * k + 1};
  }
  S->col.resize(2 * E);
  S->slot_ij.resize(E);
  S->slot_ji.resize(E);
  for (uint32_t a = 0; a < N; ++a) {
    std::sort(ent.begin() + S->rowptr[a], ent.begin() + S->rowptr[a + 1]);
    for (uint32_t s = S->rowptr[a]; s < S->rowptr[a + 1]; ++s) {
      S->col[s] = ent[s].first;
      const uint64_t k = ent[s].second >> 1;
      if (ent[s].second & 1) S->slot_ji[k] = s; else S->slot_ij[k] = s;
    }
  }
}

/* Everything one evaluation produces.  scale == nullptr: unscaled.                       */
struct Linearization {
  double cost = 0;
  std::vector<double> g;     /* [3N]   J~^T r~                                          */
  std::vector<double> hdiag; /* [9N]   diagonal blocks, row-major                        */
  std::vector<double> hoff;  /* [9*2E] off-diagonal blocks in Structure order            */
};

int Linearize(const gsfm_ra_problem* p, const LossEval& loss, const double* omega, const Structure& S,
              Linearization* L, bool jac, int num_threads) {
  const uint32_t N = p->num_views;
  const uint64_t E = p->num_edges;
  if (jac) {
    L->g.assign(3 * (size_t)N, 0.0);
    L->hdiag.assign(9 * (size_t)N, 0.0);
    L->hoff.assign(9 * 2 * (size_t)E, 0.0);
  }
  std::vector<EdgeEval> ev;
  std::vector<double> half_rho(E);
  if (jac) ev.resize(E);
  const int nt = Threads(num_threads);
  /* Pass 1 (threaded, as Ceres' ProgramEvaluator): residual + Jacobian per block. The loss
   * callback (Python in the reference, under the GIL) is called serially in pass 2.     */
  std::vector<EdgeEval> tmp_noj;
  if (!jac) tmp_noj.resize(E);
  EdgeEval* arr = jac ? ev.data() : tmp_noj.data();
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t k = 0; k < (int64_t)E; ++k) EvalEdgeRaw(p, k, omega, &arr[k], jac);
  if (loss.cb) {
    for (uint64_t k = 0; k < E; ++k) {
      loss(arr[k].sq_norm(), arr[k].rho);
    }
  } else {
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t k = 0; k < (int64_t)E; ++k) {
      loss(arr[k].sq_norm(), arr[k].rho);
    }
  }
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t k = 0; k < (int64_t)E; ++k) {
    half_rho[k] = 0.5 * arr[k].rho[0];
    if (jac) Robustify(&arr[k], true);
  }
  /* cost: summed in edge order (Ceres sums per-thread scratch; order is not specified) */
  double cost = 0;
  for (uint64_t k = 0; k < E; ++k) cost += half_rho[k];
  L->cost = cost;
  if (!jac) return 0;
  /* serial accumulation in edge order: deterministic */
  for (uint64_t k = 0; k < E; ++k) {
    const EdgeEval& e = ev[k];
    const size_t i = p->edge_i[k], j = p->edge_j[k];
    double* gi = &L->g[3 * i];
    double* gj = &L->g[3 * j];
    double* Hii = &L->hdiag[9 * i];
    double* Hjj = &L->hdiag[9 * j];
    double* Hij = &L->hoff[9 * S.slot_ij[k]];
    double* Hji = &L->hoff[9 * S.slot_ji[k]];
    for (int a = 0; a < 3; ++a) {
      double ga = 0, gb = 0;
      for (int q = 0; q < e.dim; ++q) { ga += e.Ji[3 * q + a] * e.r[q]; gb += e.Jj[3 * q + a] * e.r[q]; }
      gi[a] += ga;
      gj[a] += gb;
      for (int b = 0; b < 3; ++b) {
        double hii = 0, hjj = 0, hij = 0;
        for (int q = 0; q < e.dim; ++q) {
          hii += e.Ji[3 * q + a] * e.Ji[3 * q + b];
          hjj += e.Jj[3 * q + a] * e.Jj[3 * q + b];
          hij += e.Ji[3 * q + a] * e.Jj[3 * q + b];
        }
        Hii[3 * a + b] += hii;
        Hjj[3 * a + b] += hjj;
        Hij[3 * a + b] = hij;
        Hji[3 * b + a] = hij;
      }
    }
  }
  if (IsPosition(p) && p->fixed_view >= 0) {
    /* problem_->SetParameterBlockConstant (position_estimator.cpp:121-122): Ceres drops the block from the program.  On the full
     * structure the same reduced system results from a zero gradient for that view and zero off-diagonal blocks in its row
     * and column -- its step is then exactly zero and nobody else sees it (its own diagonal block only keeps the system
     * non-singular); the diagonal blocks of its neighbours keep the edges' contributions, as in the reduced program.      */
    const size_t f = (size_t)p->fixed_view;
    for (int a = 0; a < 3; ++a) L->g[3 * f + a] = 0.0;
    for (uint64_t k = 0; k < E; ++k)
      if (p->edge_i[k] == f || p->edge_j[k] == f)
        for (int q = 0; q < 9; ++q) { L->hoff[9 * S.slot_ij[k] + q] = 0.0; L->hoff[9 * S.slot_ji[k] + q] = 0.0; }
  }
  return 0;
}

/* y = (H + diag(d)) x on the block structure. */
void Apply(const Structure& S, const Linearization& L, const double* d, const double* x, double* y, uint32_t N, int nt) {
#pragma omp parallel for num_threads(nt) schedule(dynamic, 64)
  for (int64_t a = 0; a < (int64_t)N; ++a) {
    const double* D = &L.hdiag[9 * a];
    const double* xa = x + 3 * a;
    double acc[3];
    for (int r = 0; r < 3; ++r) acc[r] = D[3 * r] * xa[0] + D[3 * r + 1] * xa[1] + D[3 * r + 2] * xa[2] + (d ? d[3 * a + r] * xa[r] : 0.0);
    for (uint32_t s = S.rowptr[a]; s < S.rowptr[a + 1]; ++s) {
      const double* B = &L.hoff[9 * (size_t)s];
      const double* xb = x + 3 * (size_t)S.col[s];
      for (int r = 0; r < 3; ++r) acc[r] += B[3 * r] * xb[0] + B[3 * r + 1] * xb[1] + B[3 * r + 2] * xb[2];
    }
    y[3 * a] = acc[0]; y[3 * a + 1] = acc[1]; y[3 * a + 2] = acc[2];
  }
}

bool Inv3Sym(const double* A, double* inv) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  if (!(std::fabs(det) > 0)) return false;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = (A[2] * A[7] - A[1] * A[8]) * id; inv[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  inv[3] = c01 * id; inv[4] = (A[0] * A[8] - A[2] * A[6]) * id; inv[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  inv[6] = c02 * id; inv[7] = (A[1] * A[6] - A[0] * A[7]) * id; inv[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}

/* Block-Jacobi preconditioned CG on (H + diag(d)) x = b. Returns iterations. */
int Pcg(const Structure& S, const Linearization& L, const double* d, const double* b, uint32_t N, double rtol, int max_it,
        double* x, double* rel_res, int nt) {
  const size_t n = 3 * (size_t)N;
  std::vector<double> Minv(9 * (size_t)N), r(b, b + n), z(n), pvec(n), Ap(n);
  for (size_t a = 0; a < N; ++a) {
    double A[9];
    for (int k = 0; k < 9; ++k) A[k] = L.hdiag[9 * a + k];
    if (d) { A[0] += d[3 * a]; A[4] += d[3 * a + 1]; A[8] += d[3 * a + 2]; }
    if (!Inv3Sym(A, &Minv[9 * a])) { for (int k = 0; k < 9; ++k) Minv[9 * a + k] = (k % 4 == 0) ? 1.0 : 0.0; }
  }
  auto precond = [&](const double* in, double* out) {
    for (size_t a = 0; a < N; ++a)
      for (int q = 0; q < 3; ++q) out[3 * a + q] = Minv[9 * a + 3 * q] * in[3 * a] + Minv[9 * a + 3 * q + 1] * in[3 * a + 1] + Minv[9 * a + 3 * q + 2] * in[3 * a + 2];
  };
  auto dot = [&](const double* u, const double* v) { double s = 0; for (size_t k = 0; k < n; ++k) s += u[k] * v[k]; return s; };
  std::fill(x, x + n, 0.0);
  const double bnorm = std::sqrt(dot(b, b));
  if (bnorm == 0) { *rel_res = 0; return 0; }
  precond(r.data(), z.data());
  pvec = z;
  double rz = dot(r.data(), z.data());
  int it = 0;
  double rn = bnorm;
  for (; it < max_it; ++it) {
    if (rn <= rtol * bnorm) break;
    Apply(S, L, d, pvec.data(), Ap.data(), N, nt);
    const double pAp = dot(pvec.data(), Ap.data());
    if (!(pAp > 0)) break;
    const double alpha = rz / pAp;
    for (size_t k = 0; k < n; ++k) { x[k] += alpha * pvec[k]; r[k] -= alpha * Ap[k]; }
    rn = std::sqrt(dot(r.data(), r.data()));
    precond(r.data(), z.data());
    const double rz_new = dot(r.data(), z.data());
    const double beta = rz_new / rz;
    rz = rz_new;
    for (size_t k = 0; k < n; ++k) pvec[k] = z[k] + beta * pvec[k];
  }
  *rel_res = rn / bnorm;
  return it;
}

/* Dense LL^T solve of (H + diag(d)) x = b; stands in for SPARSE_NORMAL_CHOLESKY
 * (src/GSfM_nonlinear_rotation_estimator.cpp:300): both are exact factorisations.      */
bool DenseSolve(const Structure& S, const Linearization& L, const double* d, const double* b, uint32_t N, double* x, int nt) {
  const size_t n = 3 * (size_t)N;
  std::vector<double> A(n * n, 0.0);
  for (size_t a = 0; a < N; ++a) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) A[(3 * a + r) * n + 3 * a + c] = L.hdiag[9 * a + 3 * r + c];
    for (int r = 0; r < 3; ++r) A[(3 * a + r) * n + 3 * a + r] += d ? d[3 * a + r] : 0.0;
    for (uint32_t s = S.rowptr[a]; s < S.rowptr[a + 1]; ++s) {
      const size_t bcol = S.col[s];
      if (bcol > a) continue; /* lower triangle only */
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) A[(3 * a + r) * n + 3 * bcol + c] = L.hoff[9 * (size_t)s + 3 * r + c];
    }
  }
  /* row-oriented Cholesky, contiguous dot products */
  for (size_t j = 0; j < n; ++j) {
    double* Lj = &A[j * n];
    double djj = Lj[j];
    for (size_t k = 0; k < j; ++k) djj -= Lj[k] * Lj[k];
    if (!(djj > 0)) return false;
    const double ljj = std::sqrt(djj);
    Lj[j] = ljj;
    const double inv = 1.0 / ljj;
#pragma omp parallel for num_threads(nt) schedule(static) if (n - j > 256)
    for (int64_t i = (int64_t)j + 1; i < (int64_t)n; ++i) {
      double* Li = &A[i * n];
      double v = Li[j];
      for (size_t k = 0; k < j; ++k) v -= Li[k] * Lj[k];
      Li[j] = v * inv;
    }
  }
  std::vector<double> y(n);
  for (size_t i = 0; i < n; ++i) {
    double v = b[i];
    for (size_t k = 0; k < i; ++k) v -= A[i * n + k] * y[k];
    y[i] = v / A[i * n + i];
  }
  for (size_t ii = n; ii-- > 0;) {
    double v = y[ii];
    for (size_t k = ii + 1; k < n; ++k) v -= A[k * n + ii] * x[k];
    x[ii] = v / A[ii * n + ii];
  }
  return true;
}

double NowMs() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int CheckProblem(const gsfm_ra_problem* p) {
  if (!p || !p->edge_i || !p->edge_j || !p->omega_ij) return GSFM_RA_ERR_INVALID;
  if (!TypeSupported(p->error_type)) return GSFM_RA_ERR_UNSUPPORTED;
  if (IsPosition(p) && (!p->orientation || p->fixed_view >= (int64_t)p->num_views)) return GSFM_RA_ERR_INVALID;
  if (TypeNeedsCov(p->error_type) && !p->cov6) return GSFM_RA_ERR_INVALID;
  for (uint64_t k = 0; k < p->num_edges; ++k)
    if (p->edge_i[k] >= p->num_views || p->edge_j[k] >= p->num_views || p->edge_i[k] == p->edge_j[k]) return GSFM_RA_ERR_INVALID;
  return 0;
}

}  // namespace

/* ------------------------------------------------------------------------------------ */
extern "C" {

void ra_oracle_loss(const gsfm_ra_loss* loss, double s, double* rho3) {
  if (loss->inner_kind != GSFM_RA_LOSS_TRIVIAL || (loss->inner_scale != 0.0 && loss->inner_scale != 1.0)) {
    /* ComposedLoss, scripts/loss_functions.py:250-265: rho(s) = f(g(s)); g may itself be a ScaledLoss */
    gsfm_ra_loss g;
    std::memset(&g, 0, sizeof(g));
    g.kind = loss->inner_kind; g.flags = loss->inner_flags; g.scale = 1.0;
    for (int k = 0; k < 4; ++k) g.p[k] = loss->inner_p[k];
    double og[3], of[3];
    BaseLoss(&g, s, og);
    if (loss->inner_scale != 0.0 && loss->inner_scale != 1.0) { og[0] *= loss->inner_scale; og[1] *= loss->inner_scale; og[2] *= loss->inner_scale; }
    BaseLoss(loss, og[0], of);
    rho3[0] = of[0];
    rho3[1] = of[1] * og[1];
    rho3[2] = of[2] * og[1] * og[1] + of[1] * og[2];
  } else {
    BaseLoss(loss, s, rho3);
  }
  /* ScaledLoss, scripts/loss_functions.py:267-281 */
  if (loss->scale != 1.0 && loss->scale != 0.0) { rho3[0] *= loss->scale; rho3[1] *= loss->scale; rho3[2] *= loss->scale; }
}

double ra_oracle_gamma_table(int nu, int index) { return GammaTable(nu, index); }

void ra_oracle_angle_axis_to_matrix(const double* w, double* R) {
  Mat3<double> M;
  AngleAxisToMatrix(w, &M);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[3 * r + c] = M.m[r][c];
}
void ra_oracle_matrix_to_angle_axis(const double* R, double* w) {
  Mat3<double> M;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M.m[r][c] = R[3 * r + c];
  MatrixToAngleAxis(M, w);
}

void ra_oracle_whiten(int error_type, const double* cov6, double edge_weight, double* U) { Whiten(error_type, cov6, edge_weight, U); }

void ra_oracle_edge(const double* wi, const double* wj, const double* wij, const double* U, double* r, double* Ji, double* Jj) {
  Jet a[3], b[3], res[3];
  for (int k = 0; k < 3; ++k) { a[k] = Jet(wi[k]); a[k].v[k] = 1.0; b[k] = Jet(wj[k]); b[k].v[3 + k] = 1.0; }
  EdgeResidual<Jet>(a, b, wij, U, res);
  for (int q = 0; q < 3; ++q) {
    if (r) r[q] = res[q].a;
    for (int c = 0; c < 3; ++c) {
      if (Ji) Ji[3 * q + c] = res[q].v[c];
      if (Jj) Jj[3 * q + c] = res[q].v[3 + c];
    }
  }
}

static std::vector<double> StateFromOmega(const gsfm_ra_problem* p, const double* omega) {
  std::vector<double> st;
  if (!IsQuat(p)) return st;
  st.resize(4 * (size_t)p->num_views);
  for (size_t a = 0; a < p->num_views; ++a) AngleAxisToQuatXYZW(omega + 3 * a, &st[4 * a]);
  return st;
}

int ra_oracle_eval_edges(const gsfm_ra_problem* p, const gsfm_ra_loss* loss, const double* omega, double* r, double* Ji,
                         double* Jj, double* rho, int num_threads) {
  if (int rc = CheckProblem(p)) return rc;
  const std::vector<double> qstate = StateFromOmega(p, omega);
  if (IsQuat(p)) omega = qstate.data();
  const int nt = Threads(num_threads);
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t k = 0; k < (int64_t)p->num_edges; ++k) {
    EdgeEval e;
    EvalEdgeRaw(p, k, omega, &e, true);
    const int d = e.dim;
    if (r) for (int q = 0; q < d; ++q) r[d * k + q] = e.r[q];
    if (Ji) for (int q = 0; q < 3 * d; ++q) Ji[3 * d * k + q] = e.Ji[q];
    if (Jj) for (int q = 0; q < 3 * d; ++q) Jj[3 * d * k + q] = e.Jj[q];
    if (rho && loss) ra_oracle_loss(loss, e.sq_norm(), rho + 3 * k);
  }
  return 0;
}

int ra_oracle_assemble(const gsfm_ra_problem* p, const gsfm_ra_loss* loss, const double* omega, double* cost, double* gradient,
                       double* hdiag, uint32_t* rowptr, uint32_t* col, double* val, int num_threads) {
  if (int rc = CheckProblem(p)) return rc;
  Structure S;
  BuildStructure(p, &S);
  Linearization L;
  LossEval le{loss, nullptr, nullptr};
  const std::vector<double> qstate = StateFromOmega(p, omega);
  if (IsQuat(p)) omega = qstate.data();
  Linearize(p, le, omega, S, &L, true, num_threads);
  if (cost) *cost = L.cost;
  if (gradient) std::memcpy(gradient, L.g.data(), L.g.size() * sizeof(double));
  if (hdiag) std::memcpy(hdiag, L.hdiag.data(), L.hdiag.size() * sizeof(double));
  if (rowptr) std::memcpy(rowptr, S.rowptr.data(), S.rowptr.size() * sizeof(uint32_t));
  if (col) std::memcpy(col, S.col.data(), S.col.size() * sizeof(uint32_t));
  if (val) std::memcpy(val, L.hoff.data(), L.hoff.size() * sizeof(double));
  return 0;
}

int ra_oracle_cost(const gsfm_ra_problem* p, const gsfm_ra_loss* loss, const double* omega, double* cost, int num_threads) {
  if (int rc = CheckProblem(p)) return rc;
  Structure S; /* unused for cost-only */
  Linearization L;
  LossEval le{loss, nullptr, nullptr};
  const std::vector<double> qstate = StateFromOmega(p, omega);
  if (IsQuat(p)) omega = qstate.data();
  Linearize(p, le, omega, S, &L, false, num_threads);
  *cost = L.cost;
  return 0;
}

/* Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy, with
 * the options of src/GSfM_nonlinear_rotation_estimator.cpp:299-303.  Order of the checks in
 * one iteration, as in trust_region_minimizer.cc of 1.14:
 *   compute step (invalid if model_cost_change <= 0) -> candidate cost -> parameter
 *   tolerance -> function tolerance (both BEFORE the step is accepted: on those two exits
 *   the parameters stay at the last accepted point) -> accept / reject -> (loop head)
 *   max iterations, gradient tolerance (after successful steps), min radius.          */
int ra_oracle_solve(const gsfm_ra_problem* p, const gsfm_ra_options* o, double* omega, gsfm_ra_summary* sum,
                    ra_oracle_loss_cb loss_cb, void* cb_ctx) {
  if (int rc = CheckProblem(p)) return rc;
  if (!o || !omega) return GSFM_RA_ERR_INVALID;
  const double t0 = NowMs();
  const uint32_t N = p->num_views;
  const size_t n = 3 * (size_t)N;
  const int nt = Threads(o->num_threads);
  Structure S;
  BuildStructure(p, &S);
  LossEval le{&o->loss, loss_cb, cb_ctx};
  Linearization L, Ltrial;
  const bool quat = IsQuat(p);
  /* QUATERNION_COSINE (rotation_estimator.cpp:82-198): the parameters are unit quaternions (x,y,z,w) with
   * EigenQuaternionParameterization; the trust region works in the 3-dim local coordinates, Plus() maps back. */
  std::vector<double> x = quat ? StateFromOmega(p, omega) : std::vector<double>(omega, omega + n);
  std::vector<double> cand(x.size()), scale(n, 1.0), damp(n), delta(n), negg(n), Hd(n), diag(n), tmp4(x.size());
  auto plus = [&](const std::vector<double>& from, const double* d, std::vector<double>* to) {
    if (!quat) { for (size_t c = 0; c < n; ++c) (*to)[c] = from[c] + d[c]; return; }
    for (size_t a = 0; a < N; ++a) QuatPlus(&from[4 * a], d + 3 * a, &(*to)[4 * a]);
  };
  gsfm_ra_summary local;
  std::memset(&local, 0, sizeof(local));
  if (sum) { local.trace = sum->trace; local.trace_capacity = sum->trace_capacity; }
  auto push = [&](const gsfm_ra_iteration& it) {
    if (local.trace && local.trace_size < local.trace_capacity) local.trace[local.trace_size++] = it;
  };
  double t_asm = 0, t_lin = 0, t_cost = 0;

  double t = NowMs();
  Linearize(p, le, x.data(), S, &L, true, nt);
  t_asm += NowMs() - t;
  double x_cost = L.cost;
  local.initial_cost = x_cost;
  if (!std::isfinite(x_cost)) { if (sum) *sum = local; return GSFM_RA_ERR_NUMERIC; }
  auto diag_of = [&](const Linearization& Lz, std::vector<double>* out) {
    for (size_t a = 0; a < N; ++a) for (int q = 0; q < 3; ++q) (*out)[3 * a + q] = Lz.hdiag[9 * a + 4 * q];
  };
  /* Jacobi scaling estimated once, at the initial point */
  if (o->jacobi_scaling) {
    diag_of(L, &diag);
    for (size_t c = 0; c < n; ++c) scale[c] = 1.0 / (1.0 + std::sqrt(diag[c]));
  }
  /* gradient_max_norm = |x - Plus(x, -g)|_inf (ambient coordinates) */
  auto gmax = [&](const Linearization& Lz) {
    double m = 0;
    if (!quat) { for (double v : Lz.g) m = std::max(m, std::fabs(v)); return m; }
    std::vector<double> ng(n);
    for (size_t c = 0; c < n; ++c) ng[c] = -Lz.g[c];
    plus(x, ng.data(), &tmp4);
    for (size_t c = 0; c < x.size(); ++c) m = std::max(m, std::fabs(x[c] - tmp4[c]));
    return m;
  };
  auto norm = [&](const std::vector<double>& v) { double s = 0; for (double u : v) s += u * u; return std::sqrt(s); };
  double x_norm = norm(x);
  double radius = o->initial_trust_region_radius;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid_steps = 0;
  gsfm_ra_iteration it0;
  std::memset(&it0, 0, sizeof(it0));
  it0.cost = x_cost; it0.gradient_max_norm = gmax(L); it0.trust_region_radius = radius;
  push(it0);
  int iteration = 0;
  int term = GSFM_RA_TERM_NONE;
  bool last_successful = false;
  double last_gmax = it0.gradient_max_norm;
  /* gradient tolerance can already hold at the start */
  if (last_gmax <= o->gradient_tolerance) term = GSFM_RA_TERM_GRADIENT_TOLERANCE;
  while (term == GSFM_RA_TERM_NONE) {
    /* loop-head checks (FinalizeIterationAndCheckIfMinimizerCanContinue) */
    if (iteration >= o->max_num_iterations) { term = GSFM_RA_TERM_MAX_ITERATIONS; break; }
    if (last_successful && last_gmax <= o->gradient_tolerance) { term = GSFM_RA_TERM_GRADIENT_TOLERANCE; break; }
    if (radius <= o->min_trust_region_radius) { term = GSFM_RA_TERM_MIN_RADIUS; break; }
    ++iteration;
    gsfm_ra_iteration it;
    std::memset(&it, 0, sizeof(it));
    it.iteration = iteration;
    /* LM diagonal in scaled coordinates: clamp(colnorm^2(J s), min, max) / radius.
     * Expressed on the unscaled system: damping_c = that / scale_c^2.                  */
    if (!reuse_diagonal) {
      diag_of(L, &diag);
      for (size_t c = 0; c < n; ++c) diag[c] = std::min(std::max(diag[c] * scale[c] * scale[c], o->min_lm_diagonal), o->max_lm_diagonal);
    }
    for (size_t c = 0; c < n; ++c) damp[c] = diag[c] / radius / (scale[c] * scale[c]);
    for (size_t c = 0; c < n; ++c) negg[c] = -L.g[c];
    t = NowMs();
    bool solved = true;
    /* GSFM_RA_SOLVER_AUTO: the exact factorisation for small graphs (the role of SPARSE_NORMAL_CHOLESKY), PCG above */
    if (o->linear_solver == GSFM_RA_SOLVER_DENSE_CHOLESKY || (o->linear_solver == GSFM_RA_SOLVER_AUTO && N <= GSFM_RA_AUTO_DENSE_MAX_VIEWS)) {
      solved = DenseSolve(S, L, damp.data(), negg.data(), N, delta.data(), nt);
      it.linear_iterations = 1;
    } else {
      it.linear_iterations = Pcg(S, L, damp.data(), negg.data(), N, o->pcg_rtol, o->pcg_max_iterations, delta.data(), &it.linear_residual, nt);
      local.total_linear_iterations += it.linear_iterations;
    }
    t_lin += NowMs() - t;
    reuse_diagonal = true;
    bool valid = solved;
    for (size_t c = 0; c < n && valid; ++c) valid = std::isfinite(delta[c]);
    double model_change = 0;
    if (valid) {
      /* model_cost_change = -(J d)^T (r + J d / 2) = -d^T g - d^T H d / 2 */
      Apply(S, L, nullptr, delta.data(), Hd.data(), N, nt);
      double dg = 0, dHd = 0;
      for (size_t c = 0; c < n; ++c) { dg += delta[c] * L.g[c]; dHd += delta[c] * Hd[c]; }
      model_change = -dg - 0.5 * dHd;
      if (!(model_change > 0.0)) valid = false;
    }
    it.model_cost_change = model_change;
    it.step_is_valid = valid;
    if (!valid) {
      if (++invalid_steps >= 5) { term = GSFM_RA_TERM_INVALID_STEPS; it.cost = x_cost; it.trust_region_radius = radius; push(it); break; }
      radius *= 0.5; /* StepIsInvalid */
      reuse_diagonal = true;
      it.cost = x_cost; it.trust_region_radius = radius; it.gradient_max_norm = last_gmax;
      last_successful = false;
      local.num_unsuccessful_steps++;
      push(it);
      continue;
    }
    invalid_steps = 0;
    plus(x, delta.data(), &cand);
    t = NowMs();
    Linearize(p, le, cand.data(), S, &Ltrial, false, nt);
    t_cost += NowMs() - t;
    double cand_cost = Ltrial.cost;
    if (!std::isfinite(cand_cost)) cand_cost = DBL_MAX;
    it.candidate_cost = cand_cost;
    if (!quat) it.step_norm = norm(delta);
    else { double s2 = 0; for (size_t c = 0; c < x.size(); ++c) s2 += (x[c] - cand[c]) * (x[c] - cand[c]); it.step_norm = std::sqrt(s2); }
    it.cost_change = x_cost - cand_cost;
    it.relative_decrease = it.cost_change / model_change;
    it.gradient_max_norm = last_gmax;
    if (it.step_norm <= o->parameter_tolerance * (x_norm + o->parameter_tolerance)) {
      term = GSFM_RA_TERM_PARAMETER_TOLERANCE; it.cost = x_cost; it.trust_region_radius = radius; push(it); break;
    }
    if (std::fabs(it.cost_change) <= o->function_tolerance * x_cost) {
      term = GSFM_RA_TERM_FUNCTION_TOLERANCE; it.cost = x_cost; it.trust_region_radius = radius; push(it); break;
    }
    if (it.relative_decrease > o->min_relative_decrease) {
      x = cand;
      x_norm = norm(x);
      t = NowMs();
      Linearize(p, le, x.data(), S, &L, true, nt);
      t_asm += NowMs() - t;
      x_cost = L.cost;
      last_gmax = gmax(L);
      it.gradient_max_norm = last_gmax;
      it.step_is_successful = 1;
      const double q = it.relative_decrease;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * q - 1.0, 3));
      radius = std::min(o->max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      last_successful = true;
      local.num_successful_steps++;
      it.cost = x_cost;
    } else {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      last_successful = false;
      local.num_unsuccessful_steps++;
      it.cost = cand_cost; /* Ceres reports the candidate cost on a rejected step */
    }
    it.trust_region_radius = radius;
    push(it);
    if (o->verbose)
      std::fprintf(stderr, "[oracle] it %3d cost %.12e dcost %+.3e |g| %.3e |step| %.3e q %.3e radius %.3e lin_it %d %s\n", iteration, x_cost,
                   it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, radius, it.linear_iterations,
                   it.step_is_successful ? "ok" : "rejected");
  }
  if (!quat) std::memcpy(omega, x.data(), n * sizeof(double));
  else {
    /* only the views that own a parameter block (touched by an edge) are converted back (:185-194) */
    std::vector<char> touched(N, 0);
    for (uint64_t k = 0; k < p->num_edges; ++k) { touched[p->edge_i[k]] = 1; touched[p->edge_j[k]] = 1; }
    for (size_t a = 0; a < N; ++a) if (touched[a]) QuatXYZWToAngleAxis(&x[4 * a], omega + 3 * a);
  }
  local.termination = term;
  local.num_iterations = iteration;
  local.final_cost = x_cost;
  local.ms_assemble = t_asm; local.ms_linear = t_lin; local.ms_cost = t_cost;
  local.ms_total = NowMs() - t0;
  if (sum) *sum = local;
  return 0;
}

/* src/GSfM_nonlinear_rotation_estimator.cpp:314-457.  Per outer iteration: weights from the MAGSAC nu=3 table at the
 * current angular residuals (C++ round(): half away from zero; the reference's clamp `if (N < x) x = N` then reads
 * table[N], one past the end -- the closed form continues the table there), a FULL Ceres solve with
 * PairwiseRotationError(omega_ij, weight) under the caller's loss, stop when mean |w - w_prev| <= 1e-7 (tested
 * after the solve).                                                                                             */
int ra_oracle_solve_sigma_consensus(const gsfm_ra_problem* p, const gsfm_ra_options* o, int32_t iters_num, double sigma_max,
                                    double* omega, gsfm_ra_summary* sum, double* weights_out) {
  if (!p || !o || !omega) return GSFM_RA_ERR_INVALID;
  if (p->error_type != GSFM_RA_ANGLE_AXIS) return GSFM_RA_ERR_INVALID;
  const uint64_t E = p->num_edges;
  const double squared_sigma_max_2 = sigma_max * sigma_max * 2.0;
  const double dof_minus_one_per_two = (kMagsac3.nu - 1.0) / 2.0;
  const double C_times_two_ad_dof = kMagsac3.C * std::pow(2.0, dof_minus_one_per_two);
  const double one_over_sigma = C_times_two_ad_dof / sigma_max;
  const double weight_zero = one_over_sigma * (std::tgamma(dof_minus_one_per_two) - kMagsac3.gamma_k);
  std::vector<double> last(E, 0.0), w(E, 0.0);
  gsfm_ra_problem q = *p;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  gsfm_ra_summary total;
  std::memset(&total, 0, sizeof(total));
  for (int it = 0; it < iters_num; ++it) {
    for (uint64_t k = 0; k < E; ++k) {
      double e[3];
      EdgeResidual<double>(omega + 3 * (size_t)p->edge_i[k], omega + 3 * (size_t)p->edge_j[k], p->omega_ij + 3 * k, I, e);
      const double residual = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
      if (residual < DBL_EPSILON) w[k] = weight_zero;
      else {
        double x = std::round(kGammaPrecision * (residual * residual) / squared_sigma_max_2);
        if ((double)kMagsac3.table_size < x) x = kMagsac3.table_size;
        w[k] = one_over_sigma * (std::exp(-x / kGammaPrecision) - kMagsac3.gamma_k);
      }
    }
    double diff = 0;
    for (uint64_t k = 0; k < E; ++k) diff += std::fabs(w[k] - last[k]);
    diff /= (double)E;
    std::swap(w, last);  /* `last` now holds the weights of this iteration */
    q.edge_weight = last.data();
    gsfm_ra_summary s1;
    std::memset(&s1, 0, sizeof(s1));
    const int rc = ra_oracle_solve(&q, o, omega, &s1, nullptr, nullptr);
    if (rc != 0) return rc;
    if (it == 0) total.initial_cost = s1.initial_cost;
    total.final_cost = s1.final_cost; total.termination = s1.termination;
    total.num_iterations += s1.num_iterations; total.num_successful_steps += s1.num_successful_steps;
    total.num_unsuccessful_steps += s1.num_unsuccessful_steps; total.total_linear_iterations += s1.total_linear_iterations;
    total.ms_total += s1.ms_total;
    total.outer_iterations = it + 1;
    total.last_weight_change = diff;
    if (diff <= 1e-7) break;
  }
  if (weights_out) std::memcpy(weights_out, last.data(), E * sizeof(double));
  if (sum) { total.trace = sum->trace; total.trace_capacity = sum->trace_capacity; *sum = total; }
  return 0;
}

/* T/sfm/filter_view_pairs_from_orientation.cc:55-118: an edge is kept iff the angle of
 * R_ij^T * (R_j R_i^T) is <= max_degrees.                                               */
int ra_oracle_filter_view_pairs(const gsfm_ra_problem* p, const double* omega, double max_degrees, uint8_t* keep, double* angle_rad) {
  if (!p || !p->edge_i || !p->edge_j || !p->omega_ij) return GSFM_RA_ERR_INVALID;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double thr = max_degrees * M_PI / 180.0;
  for (uint64_t k = 0; k < p->num_edges; ++k) {
    double e[3];
    EdgeResidual<double>(omega + 3 * (size_t)p->edge_i[k], omega + 3 * (size_t)p->edge_j[k], p->omega_ij + 3 * k, I, e);
    const double a = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    if (angle_rad) angle_rad[k] = a;
    if (keep) keep[k] = (a * a <= thr * thr) ? 1 : 0;
  }
  return 0;
}

} /* extern "C" */
