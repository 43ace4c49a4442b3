// so3_device.cuh -- device-side SO(3) algebra and robust losses for the rotation-averaging kernels.
//
// What it replaces (reference, CPU): one ceres::AutoDiffCostFunction evaluation of
//   PairwiseRotationErrorAngleAxis   include/pairwise_rotation_error_quat.hpp:215-247
//   theia::PairwiseRotationError     T/sfm/global_pose_estimation/pairwise_rotation_error.h:66-95
// plus one Python LossFunction.Evaluate (scripts/loss_functions.py) per edge per evaluation.
//
// Design (not a translation): rotations travel as unit quaternions (w,x,y,z), the residual
// e = Log(R_j R_i^T R_ij^T) comes straight from a quaternion product, and the Jacobians are the
// closed-form SO(3) ones in the BODY (right) TANGENT frame, R <- R Exp(beta):
//     de/d(beta_j) = Jl^-1(e) R_j =: B        de/d(beta_i) = -B
// so every edge is a symmetric 3x3 Laplacian stencil (see edge_terms).  The reference differentiates
// w.r.t. the angle-axis vector omega itself; d(beta) = Jr(omega) d(omega), a per-VIEW 3x3 factor that
// the solver applies once per view (node kernels) instead of once per edge.  All fp64.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace gsfm {

struct Q4 { double w, x, y, z; };

// Reciprocal / reciprocal square root for NORMAL, finite, non-zero arguments (every call site guarantees it): the hardware
// seed (MUFU.RCP64H / RSQ64H, ~20 bits) and two Newton steps -- 5 / 7 straight-line instructions with a result within one ulp,
// instead of the IEEE division / square root sequences with their special-case branches (~25 instructions and a slow-path
// call each; K1 is bound by instruction issue, profiles/r02a_k_edges_20M_dynamic_mix.txt).  Host builds use the exact operations.
__host__ __device__ __forceinline__ double fast_rcp(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}
__host__ __device__ __forceinline__ double fast_rsqrt(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
#else
  return 1.0 / sqrt(x);
#endif
}

__host__ __device__ inline Q4 qmul(const Q4& a, const Q4& b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__host__ __device__ inline Q4 qconj(const Q4& a) { return Q4{a.w, -a.x, -a.y, -a.z}; }

// Angle-axis -> unit quaternion (ceres AngleAxisToQuaternion: half-angle form, first-order below).
__host__ __device__ inline Q4 aa_to_quat(double wx, double wy, double wz) {
  const double t2 = wx * wx + wy * wy + wz * wz;
  if (t2 > 0.0) {
    const double t = sqrt(t2);
    double sh, ch;
    sincos(0.5 * t, &sh, &ch);
    const double k = sh / t;
    return Q4{ch, wx * k, wy * k, wz * k};
  }
  return Q4{1.0, 0.5 * wx, 0.5 * wy, 0.5 * wz};
}

// Unit quaternion -> rotation matrix, row-major.
__host__ __device__ inline void quat_to_mat(const Q4& q, double* R) {
  const double xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  const double xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
  const double wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
  R[0] = 1.0 - 2.0 * (yy + zz); R[1] = 2.0 * (xy - wz);       R[2] = 2.0 * (xz + wy);
  R[3] = 2.0 * (xy + wz);       R[4] = 1.0 - 2.0 * (xx + zz); R[5] = 2.0 * (yz - wx);
  R[6] = 2.0 * (xz - wy);       R[7] = 2.0 * (yz + wx);       R[8] = 1.0 - 2.0 * (xx + yy);
}

// GetRotatedTranslation (src/GSfM_nonlinear_position_estimator.cpp:36-44): R(omega)^T t, the relative translation of a view
// pair rotated into the global frame by the first view's orientation.  Returned in the x,y,z slots of a Q4 (w = 0).
__host__ __device__ inline Q4 rotated_translation(const double* omega, const double* t) {
  double R[9];
  quat_to_mat(aa_to_quat(omega[0], omega[1], omega[2]), R);
  return Q4{0.0, R[0] * t[0] + R[3] * t[1] + R[6] * t[2], R[1] * t[0] + R[4] * t[1] + R[7] * t[2], R[2] * t[0] + R[5] * t[1] + R[8] * t[2]};
}

// Log of a unit quaternion as a rotation vector with angle in [0, pi] (ceres QuaternionToAngleAxis),
// also returning c(theta) of Jl^-1(e) = I - [e]x/2 + c [e]x^2,
//   c = 1/theta^2 - cot(theta/2)/(2 theta),  cot(theta/2) = w/|v| (no trigonometry needed).
__host__ __device__ inline void quat_log(Q4 q, double* e, double* theta2_out, double* c_out) {
  if (q.w < 0.0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double s2 = q.x * q.x + q.y * q.y + q.z * q.z;
  double k = 2.0, theta2 = 0.0, c = 1.0 / 12.0;
  if (s2 > 1e-280) {
    const double s = s2 * fast_rsqrt(s2);
    const double theta = 2.0 * atan2(s, q.w);
    // one reciprocal serves theta/s, 1/theta^2 and w/(2 theta s): 1/theta = s inv, 1/s = theta inv
    const double inv = fast_rcp(theta * s);
    const double it = s * inv, is = theta * inv;
    k = theta * is;
    theta2 = theta * theta;
    if (theta2 < 1e-2) {
      c = 1.0 / 12.0 + theta2 * (1.0 / 720.0 + theta2 * (1.0 / 30240.0 + theta2 * (1.0 / 1209600.0 + theta2 * (1.0 / 47900160.0))));
    } else {
      c = it * (it - 0.5 * q.w * is);
    }
  }
  e[0] = q.x * k; e[1] = q.y * k; e[2] = q.z * k;
  *theta2_out = theta2;
  *c_out = c;
}

// Left Jacobian of SO(3) at omega, row-major: Jl = I + a [w]x + b [w]x^2,
// a = (1-cos t)/t^2 = 2 sin^2(t/2)/t^2, b = (t - sin t)/t^3.
__host__ __device__ inline void so3_left_jacobian(double wx, double wy, double wz, double* J) {
  const double t2 = wx * wx + wy * wy + wz * wz;
  double a, b;
  if (t2 < 0.04) {
    a = 0.5 + t2 * (-1.0 / 24.0 + t2 * (1.0 / 720.0 + t2 * (-1.0 / 40320.0 + t2 * (1.0 / 3628800.0 + t2 * (-1.0 / 479001600.0)))));
    b = 1.0 / 6.0 + t2 * (-1.0 / 120.0 + t2 * (1.0 / 5040.0 + t2 * (-1.0 / 362880.0 + t2 * (1.0 / 39916800.0 + t2 * (-1.0 / 6227020800.0)))));
  } else {
    const double t = sqrt(t2);
    double sh, ch;
    sincos(0.5 * t, &sh, &ch);
    a = 2.0 * sh * sh / t2;
    b = (t - 2.0 * sh * ch) / (t2 * t);
  }
  // [w]x^2 = w w^T - t2 I
  J[0] = 1.0 + b * (wx * wx - t2); J[1] = -a * wz + b * wx * wy;    J[2] = a * wy + b * wx * wz;
  J[3] = a * wz + b * wx * wy;     J[4] = 1.0 + b * (wy * wy - t2); J[5] = -a * wx + b * wy * wz;
  J[6] = -a * wy + b * wx * wz;    J[7] = a * wx + b * wy * wz;     J[8] = 1.0 + b * (wz * wz - t2);
}

__host__ __device__ inline bool inv3(const double* A, double* inv) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = (A[2] * A[7] - A[1] * A[8]) * id; inv[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  inv[3] = c01 * id; inv[4] = (A[0] * A[8] - A[2] * A[6]) * id; inv[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  inv[6] = c02 * id; inv[7] = (A[1] * A[6] - A[0] * A[7]) * id; inv[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return det != 0.0 && isfinite(id);
}

// symmetric 3x3 packed as (00,01,02,11,12,22)
__host__ __device__ inline void sym_inv(const double* S, double* inv) {
  const double c00 = S[3] * S[5] - S[4] * S[4], c01 = S[2] * S[4] - S[1] * S[5], c02 = S[1] * S[4] - S[2] * S[3];
  const double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = (S[0] * S[5] - S[2] * S[2]) * id; inv[4] = (S[1] * S[2] - S[0] * S[4]) * id;
  inv[5] = (S[0] * S[3] - S[1] * S[1]) * id;
}
__host__ __device__ inline void sym_mul_vec(const double* S, const double* x, double* y) {
  y[0] = S[0] * x[0] + S[1] * x[1] + S[2] * x[2];
  y[1] = S[1] * x[0] + S[3] * x[1] + S[4] * x[2];
  y[2] = S[2] * x[0] + S[4] * x[1] + S[5] * x[2];
}
// B^T S B for general B (row-major) and packed symmetric S; packed symmetric result.
__host__ __device__ inline void congruence(const double* B, const double* S, double* out) {
  double T[9];  // T = S B
  for (int c = 0; c < 3; ++c) {
    T[c] = S[0] * B[c] + S[1] * B[3 + c] + S[2] * B[6 + c];
    T[3 + c] = S[1] * B[c] + S[3] * B[3 + c] + S[4] * B[6 + c];
    T[6 + c] = S[2] * B[c] + S[4] * B[3 + c] + S[5] * B[6 + c];
  }
  out[0] = B[0] * T[0] + B[3] * T[3] + B[6] * T[6];
  out[1] = B[0] * T[1] + B[3] * T[4] + B[6] * T[7];
  out[2] = B[0] * T[2] + B[3] * T[5] + B[6] * T[8];
  out[3] = B[1] * T[1] + B[4] * T[4] + B[7] * T[7];
  out[4] = B[1] * T[2] + B[4] * T[5] + B[7] * T[8];
  out[5] = B[2] * T[2] + B[5] * T[5] + B[8] * T[8];
}

// ------------------------------------------------------------------------------------------
// Robust losses, scripts/loss_functions.py (line numbers per kind in include/gsfm_ra.h).
// ------------------------------------------------------------------------------------------
struct DevLoss {
  int kind;
  unsigned flags;
  double p0, p1, scale;
  double sq0, inv_sq0;        // p0^2 and 1 / p0^2, rounded exactly as the formulas below would compute them
  // MAGSAC (scripts/loss_functions.py:285-459), constants of include/gamma_values.cpp
  int nu;
  int table_size;
  double clamp_s;             // sigma_quantile^2 * sigma^2
  double sq_sigma;            // sigma^2
  double sq_sigma_max_2;      // 2 sigma^2
  double cubed_sigma;         // sigma^3
  double Ctd;                 // C * 2^((nu-1)/2)
  double one_over_sigma;      // Ctd / sigma
  double gamma_k;             // upper_incomplete_gamma_of_k
  double weight_zero;         // one_over_sigma * (tgamma((nu-1)/2) - gamma_k)
  double expo;                // nu/2 - 1.5
  // ComposedLoss (scripts/loss_functions.py:250-265): rho(s) = f(g(s)); the fields above describe f, these g
  // (a closed-form kind below MAGSAC; composed == 0: g(s) = s)
  int composed;
  int inner_kind;
  double ip0, ip1, iscale, isq0, iinv_sq0;
  // kLossTabulated: (rho, rho', rho'') of an arbitrary loss object at the knots 0, 2^(min_exp+o) (1 + m / 2^log2_per_octave), ...
  const double* table;
  int tab_min_exp, tab_octaves, tab_log2_per_octave, tab_rows;
};

enum { kLossTrivial = 0, kLossHuber, kLossSoftLOne, kLossCauchy, kLossArctan, kLossTolerant, kLossTukey,
       kLossLOneHalf, kLossLTwo, kLossGemanMcClure, kLossMagsac3, kLossMagsac4, kLossMagsac9, kLossTabulated };

// stored_gamma_values{nu}[index] = Gamma((nu-1)/2, index/1000), closed forms
__host__ __device__ inline double gamma_table(int nu, double x) {
  if (nu == 3) return exp(-x);
  if (nu == 4) return 0.88622692545275801365 * erfc(sqrt(x)) + sqrt(x) * exp(-x);
  return exp(-x) * (((x + 3.0) * x + 6.0) * x + 6.0);
}

__host__ __device__ inline double pow_expo(double b, int nu) {
  // (s/2sigma^2)^(nu/2-1.5): nu=3 -> b^0 = 1 (also at b = 0), nu=4 -> sqrt(b), nu=9 -> b^3
  if (nu == 3) return 1.0;
  if (nu == 4) return sqrt(b);
  return b * b * b;
}

// kNu = 3 / 4 / 9 fixes the degrees of freedom at compile time (the table and the power fold away); 0 reads L.nu
template <int kNu = 0>
__host__ __device__ inline void magsac_loss(const DevLoss& L, double s_in, double* rho) {
  const int nu = kNu ? kNu : L.nu;
  double sr = s_in;
  bool zero_derivative = false;
  if (sr > L.clamp_s) { sr = L.clamp_s; zero_derivative = true; }
  double xr = rint(1000.0 * sr / L.sq_sigma_max_2);  // Python round(): half to even
  if ((double)L.table_size < xr) xr = (double)L.table_size;
  double s = xr * L.sq_sigma_max_2 / 1000.0;
  // s / (2 sigma^2) == xr / 1000 up to one rounding: ONE exponential serves the table value (nu = 3: exp(-x)), rho' and rho''
  // (the reference evaluates three; the arguments differ by an ulp, the values by 1e-16 relative); only when s is floored at
  // 1e-7 for rho'' does that term need its own
  const double ex = exp(-s / L.sq_sigma_max_2);
  const double weight = L.one_over_sigma * ((nu == 3 ? ex : gamma_table(nu, xr / 1000.0)) - L.gamma_k);
  const double wd = -L.Ctd * pow_expo(s / L.sq_sigma_max_2, nu) * ex / (2.0 * L.cubed_sigma);
  double ex2 = ex;
  if (s < 1e-7) { s = 1e-7; ex2 = exp(-s / L.sq_sigma_max_2); }
  const double wdd = 2.0 * L.Ctd * pow_expo(s / L.sq_sigma_max_2, nu) * (1.0 / L.sq_sigma - ((double)nu - 3.0) / s) * ex2 / (8.0 * L.cubed_sigma);
  if (L.flags & 1u) {
    rho[0] = 1.0 / weight;
    rho[1] = -1.0 / (weight * weight) * wd;
    rho[2] = 2.0 / (weight * weight * weight) * wd * wd - wdd / (weight * weight);
    if (zero_derivative) { rho[1] = 0.00001; rho[2] = 0.0; }
  } else {
    rho[0] = L.weight_zero - weight;
    rho[1] = -wd;
    rho[2] = -wdd;
    if (rho[1] == 0.0) rho[1] = 0.00001;
    if (zero_derivative) { rho[1] = 0.00001; rho[2] = 0.0; }
  }
}

// Hermite interpolation of a tabulated loss (include/gsfm_ra.h, GSFM_RA_LOSS_TABULATED): rho from the quintic polynomial
// matching (rho, rho', rho'') at the two enclosing knots, rho' from the cubic matching (rho', rho''), rho'' as the derivative
// of that cubic.  The knot index comes straight from the exponent and the leading mantissa bits of s.
__host__ __device__ inline void tabulated_loss(const DevLoss& L, double s, double* out) {
  const double* T = L.table;
  if (!(s > 0.0)) { out[0] = T[0]; out[1] = T[1]; out[2] = T[2]; return; }
#ifdef __CUDA_ARCH__
  const long long bits = __double_as_longlong(s);
#else
  long long bits; memcpy(&bits, &s, 8);
#endif
  const int e = (int)((bits >> 52) & 0x7ff) - 1023;
  const int P = L.tab_log2_per_octave, M = 1 << P;
  int idx;
  double s0, h;
  if (e < L.tab_min_exp) { idx = 0; s0 = 0.0; h = ldexp(1.0, L.tab_min_exp); }
  else if (e >= L.tab_min_exp + L.tab_octaves) {  // beyond the last knot: linear continuation
    const double* last = T + 3 * (size_t)(L.tab_rows - 1);
    const double sl = ldexp(1.0, L.tab_min_exp + L.tab_octaves);
    out[0] = last[0] + last[1] * (s - sl); out[1] = last[1]; out[2] = 0.0;
    return;
  } else {
    const int m = (int)((bits >> (52 - P)) & (long long)(M - 1));
    idx = 1 + (e - L.tab_min_exp) * M + m;
    h = ldexp(1.0, e - P);
    s0 = ldexp(1.0, e) + m * h;
  }
  const double* A = T + 3 * (size_t)idx;
  const double f0 = A[0], d0 = A[1] * h, c0 = A[2] * h * h, f1 = A[3], d1 = A[4] * h, c1 = A[5] * h * h;
  const double t = (s - s0) / h, t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
  const double H0 = 1.0 - 10.0 * t3 + 15.0 * t4 - 6.0 * t5, H1 = t - 6.0 * t3 + 8.0 * t4 - 3.0 * t5, H2 = 0.5 * t2 - 1.5 * t3 + 1.5 * t4 - 0.5 * t5;
  const double H3 = 10.0 * t3 - 15.0 * t4 + 6.0 * t5, H4 = -4.0 * t3 + 7.0 * t4 - 3.0 * t5, H5 = 0.5 * t3 - t4 + 0.5 * t5;
  out[0] = f0 * H0 + d0 * H1 + c0 * H2 + f1 * H3 + d1 * H4 + c1 * H5;
  // rho' from the cubic Hermite polynomial of (rho', rho'') alone and rho'' as its derivative: the step then never depends on
  // rounding noise in the object's rho values (differences of rho between neighbouring knots can be below its own precision)
  const double g0 = A[1], g1 = A[4], e0 = A[2] * h, e1 = A[5] * h;
  out[1] = g0 * (2.0 * t3 - 3.0 * t2 + 1.0) + e0 * (t3 - 2.0 * t2 + t) + g1 * (-2.0 * t3 + 3.0 * t2) + e1 * (t3 - t2);
  out[2] = (g0 * (6.0 * t2 - 6.0 * t) + e0 * (3.0 * t2 - 4.0 * t + 1.0) + g1 * (-6.0 * t2 + 6.0 * t) + e1 * (3.0 * t2 - 2.0 * t)) / h;
}

// The closed-form losses below MAGSAC, parameters passed explicitly (outer and inner function of a composition share it).
// sq0 = p0^2 and inv_sq0 = 1 / p0^2, rounded exactly as the reference's formulas would compute them.
__host__ __device__ inline void eval_simple_loss(int kind, double p0, double p1, double sq0, double inv_sq0, double s, double* out) {
  switch (kind) {
    case kLossTrivial: out[0] = s; out[1] = 1.0; out[2] = 0.0; break;
    case kLossHuber: {
      const double a = p0, b = sq0;
      if (s > b) { const double r = sqrt(s); out[0] = 2.0 * a * r - b; out[1] = fmax(a / r, DBL_MIN); out[2] = -out[1] / (2.0 * s); }
      else { out[0] = s; out[1] = 1.0; out[2] = 0.0; }
      break;
    }
    case kLossSoftLOne: {
      const double b = sq0, c = inv_sq0;
      const double sum = 1.0 + s * c, rt = fast_rsqrt(sum), tmp = sum * rt;   // sum >= 1
      out[0] = 2.0 * b * (tmp - 1.0); out[1] = fmax(rt, DBL_MIN); out[2] = -(c * out[1]) * (0.5 * rt * rt);
      break;
    }
    case kLossCauchy: {
      const double b = sq0, c = inv_sq0;
      const double sum = 1.0 + s * c, inv = fast_rcp(sum);   // sum >= 1
      out[0] = b * log(sum); out[1] = fmax(inv, DBL_MIN); out[2] = -c * (inv * inv);
      break;
    }
    case kLossArctan: {
      const double a = p0, b = 1.0 / (a * a);
      const double sum = 1.0 + s * s * b, inv = 1.0 / sum;
      out[0] = a * atan2(s, a); out[1] = fmax(inv, DBL_MIN); out[2] = -2.0 * s * b * (inv * inv);
      break;
    }
    case kLossTolerant: {
      const double a = p0, b = p1, c = b * log(1.0 + exp(-a / b));
      const double x = (s - a) / b;
      if (x > 36.7) { out[0] = s - a - c; out[1] = 1.0; out[2] = 0.0; }
      else { const double e_x = exp(x); out[0] = b * log(1.0 + e_x) - c; out[1] = fmax(e_x / (1.0 + e_x), DBL_MIN); out[2] = 0.5 / (b * (1.0 + cosh(x))); }
      break;
    }
    case kLossTukey: {
      const double a2 = p0 * p0;
      if (s <= a2) { const double v = 1.0 - s / a2, v2 = v * v; out[0] = a2 / 6.0 * (1.0 - v2 * v); out[1] = 0.5 * v2; out[2] = -1.0 / a2 * v; }
      else { out[0] = a2 / 6.0; out[1] = 0.0; out[2] = 0.0; }
      break;
    }
    case kLossLOneHalf: {
      const double a = p0, sa = sqrt(a);
      out[0] = 2.0 * a * sa * pow(s, 0.25);
      if (s < 0.01) s = 0.01;
      out[1] = 0.5 * pow(a, -1.5) * pow(s, -0.75);
      out[2] = -0.375 * a * sa * pow(s, -1.75);
      break;
    }
    case kLossLTwo: {
      const double a2 = p0 * p0;
      out[0] = s * s / (a2 * 2.0); out[1] = s / a2; out[2] = 1.0 / a2;
      break;
    }
    case kLossGemanMcClure: {
      const double a2 = p0 * p0, sg = p1;
      const double d = s / a2 + sg;
      out[0] = a2 * sg * s / (2.0 * (s + a2 * sg));
      out[1] = (sg * sg) / (2.0 * d * d);
      out[2] = -(sg * sg) / (a2 * d * d * d);
      break;
    }
    default: out[0] = out[1] = out[2] = NAN;
  }
}

// kKind >= 0 fixes the loss at compile time (the switch folds away; plain, unscaled, uncomposed losses only -- the host picks
// such an instantiation only then); kKind < 0 dispatches on L.kind and handles composition.
template <int kKind = -1>
__host__ __device__ inline void eval_loss(const DevLoss& L, double s, double* out) {
  double og[3] = {s, 1.0, 0.0};
  const bool comp = kKind < 0 && L.composed != 0;
  if (comp) {
    eval_simple_loss(L.inner_kind, L.ip0, L.ip1, L.isq0, L.iinv_sq0, s, og);
    if (L.iscale != 1.0) { og[0] *= L.iscale; og[1] *= L.iscale; og[2] *= L.iscale; }
    s = og[0];
  }
  const int kind = kKind < 0 ? L.kind : kKind;
  if (kKind == kLossMagsac3) magsac_loss<3>(L, s, out);
  else if (kind >= kLossMagsac3 && kind <= kLossMagsac9) magsac_loss<0>(L, s, out);
  else if (kind == kLossTabulated) tabulated_loss(L, s, out);
  else eval_simple_loss(kind, L.p0, L.p1, L.sq0, L.inv_sq0, s, out);
  if (comp) {  // f(g(s)): f' g',  f'' g'^2 + f' g''
    const double f1 = out[1], f2 = out[2];
    out[1] = f1 * og[1];
    out[2] = f2 * og[1] * og[1] + f1 * og[2];
  }
  if (L.scale != 1.0 && L.scale != 0.0) { out[0] *= L.scale; out[1] *= L.scale; out[2] *= L.scale; }
}

// ------------------------------------------------------------------------------------------
// One relative-rotation constraint in the BODY (right) tangent frame, R <- R Exp(beta).
// For any residual that is a function of the error rotation E = R_j R_i^T R_ij^T,
//     d r / d beta_j = +B,   d r / d beta_i = -B,   B = A R_j,  A = d r / d(left perturbation of E),
// so one edge contributes a matrix-weighted graph-Laplacian stencil to the normal equations:
//     S = rho' (B^T B - kappa u u^T), u = B^T r :   H_ii += S, H_jj += S, H_ij = H_ji = -S   (S symmetric)
//     v = rho' u :                                   g_j += v, g_i -= v
// (Ceres Corrector, SURVEY Appendix B.2: J~ = sqrt(rho')(I - alpha/s r r^T) J, r~ = sqrt(rho')/(1-alpha) r
//  =>  J~^T J~ = rho' J^T (I - kappa r r^T) J with kappa = (2 alpha - alpha^2)/s, and J~^T r~ = rho' J^T r.)
// The reference differentiates w.r.t. the angle-axis vector omega; d(beta) = Jr(omega) d(omega) is a
// per-VIEW factor applied by the node kernels.  Inputs: endpoint quaternions, the measured q_ij, the
// upper triangular weight U (packed u00 u01 u02 u11 u12 u22; kScalarU: U = u00 I).
// ------------------------------------------------------------------------------------------
struct EdgeTerms {
  double r[3];
  double B[9];
  double S[6];
  double v[3];
  double rho[3];
  // scalar-weight stencil only (kStencilOnly && kScalarU && kResidual == 0): S = |ca| I + sign(ca) h h^T, the 4-double
  // form the compact block records store (ca carries the sign of the rank-one term; |ca| > 0 whenever rho' > 0 because
  // the eigenvalue of Jl^-T Jl^-1 orthogonal to e is (theta/2)^2 / sin^2(theta/2) >= 1)
  double ca, h[3];
};

// Ceres Corrector's kappa (zero for every loss with rho'' <= 0).
__host__ __device__ inline double triggs_kappa(double s, const double* rho) {
  if (s != 0.0 && rho[2] > 0.0) {
    const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
    const double alpha = 1.0 - ((D > 0.0) ? sqrt(D) : 0.0);
    return (2.0 * alpha - alpha * alpha) / s;
  }
  return 0.0;
}

// kResidual = 0: r = U Log(E)                    (angle-axis types 3..8);  A = U Jl^-1(e)
// kResidual = 1: r = -2 w vec(q_E) = 2 w vec(q_ij (q_j q_i^-1)^-1)   (QUATERNION_COSINE,
//                include/pairwise_rotation_error_quat.hpp:82-106; w = U[0]);  A = w ([v_E]x - w_E I).
//                No logarithm, smooth through theta = pi.
// kStencilOnly (K1): only S, v, rho are produced, B is not formed.  With a scalar weight w the stencil has a closed form:
//   B = w M R_j, M = Jl^-1(e) = I - K/2 + c K^2, K = [e]x  =>  M^T M = I + g K^2 = (1 - g theta^2) I + g e e^T,
//   g = 2c - 1/4 - c^2 theta^2, and M^T e = e, so with f = R_j^T e (the error in view j's body frame)
//   B^T B = w^2 ((1 - g theta^2) I + g f f^T),   u = B^T r = w^2 f,
//   S = rho' w^2 ((1 - g theta^2) I + (g - kappa w^2) f f^T),   v = rho' w^2 f
// -- no rotation matrix, no 3x3 products.
// kResidual = 2: TRANSLATION averaging (src/GSfM_nonlinear_position_estimator.cpp:252-296 with
//                theia::PairwiseTranslationError, T/sfm/global_pose_estimation/pairwise_translation_error.h:62-88).  The "views"
//                are camera positions c (Euclidean parameters, carried in the x,y,z slots of the Q4 arguments), the measurement
//                is the unit direction t_ij (in the x,y,z slots of qij) and
//                    r = w (d / |d| - t_ij),  d = c_j - c_i,  |d| := 1 when |d| < 1e-12 (the reference's kNormTolerance branch:
//                    the norm becomes the CONSTANT 1, so the Jacobian is w I there)
//                    d r / d c_j = +B,  d r / d c_i = -B,  B = (w / |d|) (I - u u^T),  u = d / |d|    (B symmetric)
//                -- again a Laplacian stencil: S = rho' (B^T B - kappa ub ub^T), ub = B^T r, v = rho' ub.
template <bool kNeedJacobian, int kLoss, bool kStencilOnly>
__host__ __device__ inline void translation_terms(const Q4& ci, const Q4& cj, const Q4& tij, double w, const DevLoss& L, EdgeTerms& o) {
  const double d0 = cj.x - ci.x, d1 = cj.y - ci.y, d2 = cj.z - ci.z;
  double n = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  const bool tiny = n < 1e-12;
  if (tiny) n = 1.0;
  const double u0 = d0 / n, u1 = d1 / n, u2 = d2 / n;
  o.r[0] = w * (u0 - tij.x); o.r[1] = w * (u1 - tij.y); o.r[2] = w * (u2 - tij.z);
  const double s = o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2];
  eval_loss<kLoss>(L, s, o.rho);
  if (!kNeedJacobian) return;
  const double a = w / n, p = tiny ? 0.0 : a;   // B = a I - p u u^T
  double B[6];                                  // symmetric, packed 00 01 02 11 12 22
  B[0] = a - p * u0 * u0; B[1] = -p * u0 * u1; B[2] = -p * u0 * u2;
  B[3] = a - p * u1 * u1; B[4] = -p * u1 * u2; B[5] = a - p * u2 * u2;
  if (!kStencilOnly) {
    o.B[0] = B[0]; o.B[1] = B[1]; o.B[2] = B[2]; o.B[3] = B[1]; o.B[4] = B[3]; o.B[5] = B[4]; o.B[6] = B[2]; o.B[7] = B[4]; o.B[8] = B[5];
  }
  const double ub0 = B[0] * o.r[0] + B[1] * o.r[1] + B[2] * o.r[2];
  const double ub1 = B[1] * o.r[0] + B[3] * o.r[1] + B[4] * o.r[2];
  const double ub2 = B[2] * o.r[0] + B[4] * o.r[1] + B[5] * o.r[2];
  const double kappa = triggs_kappa(s, o.rho);
  const double rho1 = o.rho[1];
  // B^T B = B B (symmetric)
  o.S[0] = rho1 * (B[0] * B[0] + B[1] * B[1] + B[2] * B[2] - kappa * ub0 * ub0);
  o.S[1] = rho1 * (B[0] * B[1] + B[1] * B[3] + B[2] * B[4] - kappa * ub0 * ub1);
  o.S[2] = rho1 * (B[0] * B[2] + B[1] * B[4] + B[2] * B[5] - kappa * ub0 * ub2);
  o.S[3] = rho1 * (B[1] * B[1] + B[3] * B[3] + B[4] * B[4] - kappa * ub1 * ub1);
  o.S[4] = rho1 * (B[1] * B[2] + B[3] * B[4] + B[4] * B[5] - kappa * ub1 * ub2);
  o.S[5] = rho1 * (B[2] * B[2] + B[4] * B[4] + B[5] * B[5] - kappa * ub2 * ub2);
  o.v[0] = rho1 * ub0; o.v[1] = rho1 * ub1; o.v[2] = rho1 * ub2;
}

template <bool kNeedJacobian, int kResidual = 0, bool kScalarU = false, int kLoss = -1, bool kStencilOnly = false>
__host__ __device__ inline void edge_terms(const Q4& qi, const Q4& qj, const Q4& qij, const double* U, const DevLoss& L, EdgeTerms& o) {
  if (kResidual == 2) { translation_terms<kNeedJacobian, kLoss, kStencilOnly>(qi, qj, qij, U[0], L, o); return; }
  const Q4 qE = qmul(qmul(qj, qconj(qi)), qconj(qij));  // error rotation R_j R_i^T R_ij^T
  if (kNeedJacobian && kStencilOnly && kScalarU && kResidual == 0) {
    double e[3], theta2, c;
    quat_log(qE, e, &theta2, &c);
    const double w = U[0], w2 = w * w;
    o.r[0] = w * e[0]; o.r[1] = w * e[1]; o.r[2] = w * e[2];
    const double s = o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2];  // exactly as the general path (quantised losses)
    eval_loss<kLoss>(L, s, o.rho);
    // f = R_j^T e = e - w_j t + v_j x t,  t = 2 v_j x e
    const double tx = 2.0 * (qj.y * e[2] - qj.z * e[1]), ty = 2.0 * (qj.z * e[0] - qj.x * e[2]), tz = 2.0 * (qj.x * e[1] - qj.y * e[0]);
    const double f0 = e[0] - qj.w * tx + (qj.y * tz - qj.z * ty);
    const double f1 = e[1] - qj.w * ty + (qj.z * tx - qj.x * tz);
    const double f2 = e[2] - qj.w * tz + (qj.x * ty - qj.y * tx);
    const double g = 2.0 * c - 0.25 - c * c * theta2;
    const double kappa = triggs_kappa(s, o.rho);
    const double rw = o.rho[1] * w2;
    const double a = rw * (1.0 - g * theta2), b = rw * (g - kappa * w2);
    const double bf0 = b * f0, bf1 = b * f1, bf2 = b * f2;
    o.S[0] = a + bf0 * f0; o.S[1] = bf0 * f1; o.S[2] = bf0 * f2;
    o.S[3] = a + bf1 * f1; o.S[4] = bf1 * f2; o.S[5] = a + bf2 * f2;
    o.v[0] = rw * f0; o.v[1] = rw * f1; o.v[2] = rw * f2;
    const double ab = fabs(b);
    const double hs = ab > 1e-290 ? ab * fast_rsqrt(ab) : 0.0;
    o.ca = copysign(a, b);
    o.h[0] = hs * f0; o.h[1] = hs * f1; o.h[2] = hs * f2;
    return;
  }
  double M[9];                                           // d r / d(left perturbation of E) before the weight
  if (kResidual == 1) {
    const double w = U[0];
    o.r[0] = -2.0 * w * qE.x; o.r[1] = -2.0 * w * qE.y; o.r[2] = -2.0 * w * qE.z;
  } else {
    double e[3], theta2, c;
    quat_log(qE, e, &theta2, &c);
    if (kScalarU) {
      o.r[0] = U[0] * e[0]; o.r[1] = U[0] * e[1]; o.r[2] = U[0] * e[2];
    } else {
      o.r[0] = U[0] * e[0] + U[1] * e[1] + U[2] * e[2];
      o.r[1] = U[3] * e[1] + U[4] * e[2];
      o.r[2] = U[5] * e[2];
    }
    if (kNeedJacobian) {
      // Jl^-1(e) = I - [e]x/2 + c (e e^T - theta2 I)
      const double d = 1.0 - c * theta2;
      M[0] = d + c * e[0] * e[0];           M[1] = 0.5 * e[2] + c * e[0] * e[1];  M[2] = -0.5 * e[1] + c * e[0] * e[2];
      M[3] = -0.5 * e[2] + c * e[0] * e[1]; M[4] = d + c * e[1] * e[1];           M[5] = 0.5 * e[0] + c * e[1] * e[2];
      M[6] = 0.5 * e[1] + c * e[0] * e[2];  M[7] = -0.5 * e[0] + c * e[1] * e[2]; M[8] = d + c * e[2] * e[2];
    }
  }
  const double s = o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2];
  eval_loss<kLoss>(L, s, o.rho);
  if (!kNeedJacobian) return;
  if (kResidual == 1) {
    M[0] = -qE.w; M[1] = -qE.z; M[2] = qE.y;
    M[3] = qE.z;  M[4] = -qE.w; M[5] = -qE.x;
    M[6] = -qE.y; M[7] = qE.x;  M[8] = -qE.w;
  }
  double Rj[9], T[9];
  quat_to_mat(qj, Rj);
#pragma unroll
  for (int rr = 0; rr < 3; ++rr)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) T[3 * rr + cc] = M[3 * rr] * Rj[cc] + M[3 * rr + 1] * Rj[3 + cc] + M[3 * rr + 2] * Rj[6 + cc];
  double* B = o.B;
  if (kScalarU || kResidual == 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k) B[k] = U[0] * T[k];
  } else {
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      B[cc] = U[0] * T[cc] + U[1] * T[3 + cc] + U[2] * T[6 + cc];
      B[3 + cc] = U[3] * T[3 + cc] + U[4] * T[6 + cc];
      B[6 + cc] = U[5] * T[6 + cc];
    }
  }
  double u[3];
#pragma unroll
  for (int cc = 0; cc < 3; ++cc) u[cc] = B[cc] * o.r[0] + B[3 + cc] * o.r[1] + B[6 + cc] * o.r[2];
  const double kappa = triggs_kappa(s, o.rho);
  const double rho1 = o.rho[1];
  o.S[0] = rho1 * (B[0] * B[0] + B[3] * B[3] + B[6] * B[6] - kappa * u[0] * u[0]);
  o.S[1] = rho1 * (B[0] * B[1] + B[3] * B[4] + B[6] * B[7] - kappa * u[0] * u[1]);
  o.S[2] = rho1 * (B[0] * B[2] + B[3] * B[5] + B[6] * B[8] - kappa * u[0] * u[2]);
  o.S[3] = rho1 * (B[1] * B[1] + B[4] * B[4] + B[7] * B[7] - kappa * u[1] * u[1]);
  o.S[4] = rho1 * (B[1] * B[2] + B[4] * B[5] + B[7] * B[8] - kappa * u[1] * u[2]);
  o.S[5] = rho1 * (B[2] * B[2] + B[5] * B[5] + B[8] * B[8] - kappa * u[2] * u[2]);
  o.v[0] = rho1 * u[0]; o.v[1] = rho1 * u[1]; o.v[2] = rho1 * u[2];
}

// ------------------------------------------------------------------------------------------
// The two residuals that are NOT functions of the error rotation alone (general two-block edges):
//   kType 0  QUATERNION_NORM      include/pairwise_rotation_error_quat.hpp:125-150, 4 residuals
//            r = w (s_b q_b - s_e q_e), q_e = q_ij q_a, s = -1 where the quaternion's y coefficient is negative
//   kType 1  ROTATION_MAT_FNORM   include/pairwise_rotation_error_quat.hpp:169-196, 9 residuals
//            r = w vec(R_ij R_a - R_b)
// Body-frame Jacobians (q <- q (x) [beta/2, 1], R <- R Exp(beta)); with M(q) = d(q (x) [beta, 0])/d beta (4x3, rows x,y,z,w):
//   type 0:  J_b = (w s_b / 2) M(q_b),   J_a = -(w s_e / 2) M(q_e)
//   type 1:  column k of J_b = -w vec(R_b [e_k]x),   of J_a = +w vec(R_e [e_k]x),  R_e = R_ij R_a
// The caller names the view whose ROW is being assembled: outputs are that row's diagonal contribution
// D = rho' (Jr^T Jr - kappa ur ur^T), off-diagonal block G = rho' (Jr^T Jc - kappa ur uc^T) (row-major, row view x column
// view), gradient g = rho' ur, with ur = Jr^T r, uc = Jc^T r.
// ------------------------------------------------------------------------------------------
struct GeneralTerms {
  double D[6];
  double G[9];
  double g[3];
  double rho[3];
};

__host__ __device__ inline void quat_right_jac(const Q4& q, double k, double* M /*4x3*/) {
  M[0] = k * q.w;  M[1] = -k * q.z; M[2] = k * q.y;
  M[3] = k * q.z;  M[4] = k * q.w;  M[5] = -k * q.x;
  M[6] = -k * q.y; M[7] = k * q.x;  M[8] = k * q.w;
  M[9] = -k * q.x; M[10] = -k * q.y; M[11] = -k * q.z;
}
// columns of k * R [e_k]x stacked as a 9x3 Jacobian; residual index 3*c + r (Eigen's column-major linear index)
__host__ __device__ inline void rot_right_jac(const double* R, double k, double* J /*9x3*/) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double r0 = k * R[3 * r], r1 = k * R[3 * r + 1], r2 = k * R[3 * r + 2];
    // d/d beta_x: (0, R[:,2], -R[:,1]);  d/d beta_y: (-R[:,2], 0, R[:,0]);  d/d beta_z: (R[:,1], -R[:,0], 0)
    J[3 * (0 + r) + 0] = 0.0;  J[3 * (0 + r) + 1] = -r2;  J[3 * (0 + r) + 2] = r1;   // column c = 0 of the matrix
    J[3 * (3 + r) + 0] = r2;   J[3 * (3 + r) + 1] = 0.0;  J[3 * (3 + r) + 2] = -r0;  // c = 1
    J[3 * (6 + r) + 0] = -r1;  J[3 * (6 + r) + 1] = r0;   J[3 * (6 + r) + 2] = 0.0;  // c = 2
  }
}

template <bool kNeedJacobian, int kType>
__host__ __device__ inline void general_edge_terms(const Q4& qa, const Q4& qb, const Q4& qij, double w, bool row_is_b, const DevLoss& L,
                                                   GeneralTerms& o) {
  constexpr int kDim = (kType == 0) ? 4 : 9;
  double r[kDim], Ja[3 * kDim], Jb[3 * kDim];
  if (kType == 0) {
    const Q4 qe = qmul(qij, qa);
    const double sb = (qb.y < 0.0) ? -1.0 : 1.0, se = (qe.y < 0.0) ? -1.0 : 1.0;
    r[0] = w * (sb * qb.x - se * qe.x); r[1] = w * (sb * qb.y - se * qe.y);
    r[2] = w * (sb * qb.z - se * qe.z); r[3] = w * (sb * qb.w - se * qe.w);
    if (kNeedJacobian) { quat_right_jac(qb, 0.5 * w * sb, Jb); quat_right_jac(qe, -0.5 * w * se, Ja); }
  } else {
    double Ra[9], Rb[9], Rr[9], Re[9];
    quat_to_mat(qa, Ra); quat_to_mat(qb, Rb); quat_to_mat(qij, Rr);
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int c = 0; c < 3; ++c) Re[3 * rr + c] = Rr[3 * rr] * Ra[c] + Rr[3 * rr + 1] * Ra[3 + c] + Rr[3 * rr + 2] * Ra[6 + c];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) r[3 * c + rr] = w * (Re[3 * rr + c] - Rb[3 * rr + c]);
    if (kNeedJacobian) { rot_right_jac(Re, w, Ja); rot_right_jac(Rb, -w, Jb); }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < kDim; ++q) s += r[q] * r[q];
  eval_loss(L, s, o.rho);
  if (!kNeedJacobian) return;
  const double* Jr = row_is_b ? Jb : Ja;
  const double* Jc = row_is_b ? Ja : Jb;
  double ur[3] = {0, 0, 0}, uc[3] = {0, 0, 0};
#pragma unroll
  for (int q = 0; q < kDim; ++q)
#pragma unroll
    for (int c = 0; c < 3; ++c) { ur[c] += Jr[3 * q + c] * r[q]; uc[c] += Jc[3 * q + c] * r[q]; }
  const double kappa = triggs_kappa(s, o.rho), rho1 = o.rho[1];
  int t = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double jj = 0.0, jc = 0.0;
#pragma unroll
      for (int q = 0; q < kDim; ++q) { jj += Jr[3 * q + a] * Jr[3 * q + b]; jc += Jr[3 * q + a] * Jc[3 * q + b]; }
      o.G[3 * a + b] = rho1 * (jc - kappa * ur[a] * uc[b]);
      if (b >= a) o.D[t++] = rho1 * (jj - kappa * ur[a] * ur[b]);
    }
  o.g[0] = rho1 * ur[0]; o.g[1] = rho1 * ur[1]; o.g[2] = rho1 * ur[2];
}

// Whitening, src/GSfM_nonlinear_rotation_estimator.cpp:251-288:
// U = chol((1e8 Sigma)^-1)^T for the covariance types, scalar otherwise; times edge_weight.
// The cofactor inverse of a covariance with condition number up to 1e12 loses
// cond * eps digits, so the RESULT depends on the exact operation sequence.  The sequence below is
// evaluated with individually rounded multiplies/adds (no FMA contraction): bit-identical to the
// CPU restatement (oracle/ra_oracle.cc Whiten) by construction; sqrt and division are IEEE on both.
#ifdef __CUDA_ARCH__
#define GSFM_MUL(a, b) __dmul_rn((a), (b))
#define GSFM_ADD(a, b) __dadd_rn((a), (b))
#define GSFM_SUB(a, b) __dsub_rn((a), (b))
#else
#define GSFM_MUL(a, b) ((a) * (b))
#define GSFM_ADD(a, b) ((a) + (b))
#define GSFM_SUB(a, b) ((a) - (b))
#endif
__host__ __device__ inline void whiten(int type, const double* c6, double w, double* U) {
  U[0] = U[1] = U[2] = U[3] = U[4] = U[5] = 0.0;
  if (type == 3 || type == 6) {
    const double a = GSFM_MUL(c6[0], 1e8), d = GSFM_MUL(c6[1], 1e8), f = GSFM_MUL(c6[2], 1e8);
    const double b = GSFM_MUL(c6[3], 1e8), c = GSFM_MUL(c6[4], 1e8), e = GSFM_MUL(c6[5], 1e8);
    const double c00 = GSFM_SUB(GSFM_MUL(d, f), GSFM_MUL(e, e)), c01 = GSFM_SUB(GSFM_MUL(c, e), GSFM_MUL(b, f));
    const double c02 = GSFM_SUB(GSFM_MUL(b, e), GSFM_MUL(c, d)), c11 = GSFM_SUB(GSFM_MUL(a, f), GSFM_MUL(c, c));
    const double c12 = GSFM_SUB(GSFM_MUL(b, c), GSFM_MUL(a, e)), c22 = GSFM_SUB(GSFM_MUL(a, d), GSFM_MUL(b, b));
    const double det = GSFM_ADD(GSFM_ADD(GSFM_MUL(a, c00), GSFM_MUL(b, c01)), GSFM_MUL(c, c02));
    const double id = 1.0 / det;
    const double P00 = GSFM_MUL(c00, id), P10 = GSFM_MUL(c01, id), P20 = GSFM_MUL(c02, id);
    const double P11 = GSFM_MUL(c11, id), P21 = GSFM_MUL(c12, id), P22 = GSFM_MUL(c22, id);
    const double l00 = sqrt(P00), l10 = P10 / l00, l20 = P20 / l00;
    const double l11 = sqrt(GSFM_SUB(P11, GSFM_MUL(l10, l10))), l21 = GSFM_SUB(P21, GSFM_MUL(l20, l10)) / l11;
    const double l22 = sqrt(GSFM_SUB(GSFM_SUB(P22, GSFM_MUL(l20, l20)), GSFM_MUL(l21, l21)));
    U[0] = GSFM_MUL(l00, w); U[1] = GSFM_MUL(l10, w); U[2] = GSFM_MUL(l20, w);
    U[3] = GSFM_MUL(l11, w); U[4] = GSFM_MUL(l21, w); U[5] = GSFM_MUL(l22, w);
    return;
  }
  double s = w;
  if (type == 7) s = GSFM_MUL(w, sqrt(1.0 / GSFM_MUL(GSFM_ADD(GSFM_ADD(c6[0], c6[1]), c6[2]), 1e8)));
  if (type == 8) {
    double n2 = 0.0;
    for (int k = 0; k < 3; ++k) { const double t = GSFM_MUL(c6[k], 1e8); n2 = GSFM_ADD(n2, GSFM_MUL(t, t)); }
    for (int k = 3; k < 6; ++k) { const double t = GSFM_MUL(c6[k], 1e8); n2 = GSFM_ADD(n2, GSFM_MUL(GSFM_MUL(2.0, t), t)); }
    s = GSFM_MUL(w, sqrt(1.0 / sqrt(n2)));
  }
  U[0] = U[3] = U[5] = s;
}

}  // namespace gsfm
