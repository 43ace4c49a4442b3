// ra_api_kernels.cuh -- API-only kernels: per-edge evaluation, whitening, loss table, exports, sigma-consensus weights, rotation filter
// Part of libgsfm_ra (one translation unit, see gsfm_ra.cu); reference citations sit next to each kernel.
#pragma once
#include "ra_common.cuh"
#include "ra_dense.cuh"
namespace {

// ---- API-only kernels (parity tests / diagnostics; not on the solve path) ------------------
// Raw per-edge outputs in EDGE order and Euclidean (angle-axis) coordinates, i.e. exactly what
// AutoDiffCostFunction::Evaluate + LossFunction::Evaluate return in the reference.
__global__ void k_eval_edges(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const double* __restrict__ omega_ij,
                             const double* __restrict__ cov6, const double* __restrict__ weight, int error_type, const double* __restrict__ node_q,
                             const double* __restrict__ node_JL, DevLoss loss, double* r, double* Ji, double* Jj, double* rho,
                             const double* __restrict__ orientation = nullptr) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  const uint32_t i = ei[k], j = ej[k];
  const double4 a = reinterpret_cast<const double4*>(node_q)[i], b = reinterpret_cast<const double4*>(node_q)[j];
  const Q4 qi{a.x, a.y, a.z, a.w}, qj{b.x, b.y, b.z, b.w};
  const Q4 qm = orientation ? rotated_translation(orientation + 3 * (size_t)i, omega_ij + 3 * k)
                            : aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
  double c6[6] = {0, 0, 0, 0, 0, 0}, u[6];
  if (cov6) for (int t = 0; t < 6; ++t) c6[t] = cov6[6 * k + t];
  whiten(error_type, c6, weight ? weight[k] : 1.0, u);
  EdgeTerms et;
  if (error_type == 2) edge_terms<true, 1>(qi, qj, qm, u, loss, et);
  else if (orientation) edge_terms<true, 2>(qi, qj, qm, u, loss, et);
  else edge_terms<true, 0>(qi, qj, qm, u, loss, et);
  if (r) for (int t = 0; t < 3; ++t) r[3 * k + t] = et.r[t];
  if (rho) for (int t = 0; t < 3; ++t) rho[3 * k + t] = et.rho[t];
  // d r / d(parameters of view j) = B D_j,  d r / d(parameters of view i) = -B D_i
  for (int side = 0; side < 2; ++side) {
    double* out = side ? Jj : Ji;
    if (!out) continue;
    const double* D = node_JL + 9 * (size_t)(side ? j : i);
    const double sgn = side ? 1.0 : -1.0;
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 3; ++c) out[9 * k + 3 * rr + c] = sgn * (et.B[3 * rr] * D[c] + et.B[3 * rr + 1] * D[3 + c] + et.B[3 * rr + 2] * D[6 + c]);
  }
}

// Same for the general two-block residuals: r [E][d], Ji/Jj [E][d][3], d = 4 (QUATERNION_NORM) or 9 (ROTATION_MAT_FNORM).
// The row-view interface of general_edge_terms does not expose the raw Jacobians, so they are rebuilt here from the same
// helpers (API-only path).
template <int kType>
__global__ void k_eval_edges_general(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const double* __restrict__ omega_ij,
                                     const double* __restrict__ weight, const double* __restrict__ node_q, const double* __restrict__ node_JL,
                                     DevLoss loss, double* r_out, double* Ji, double* Jj, double* rho) {
  constexpr int kDim = (kType == 0) ? 4 : 9;
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  const uint32_t i = ei[k], j = ej[k];
  const double4 a = reinterpret_cast<const double4*>(node_q)[i], b = reinterpret_cast<const double4*>(node_q)[j];
  const Q4 qa{a.x, a.y, a.z, a.w}, qb{b.x, b.y, b.z, b.w};
  const Q4 qm = aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
  const double w = weight ? weight[k] : 1.0;
  double r[kDim], Ja[3 * kDim], Jb[3 * kDim];
  if (kType == 0) {
    const Q4 qe = qmul(qm, qa);
    const double sb = (qb.y < 0.0) ? -1.0 : 1.0, se = (qe.y < 0.0) ? -1.0 : 1.0;
    r[0] = w * (sb * qb.x - se * qe.x); r[1] = w * (sb * qb.y - se * qe.y);
    r[2] = w * (sb * qb.z - se * qe.z); r[3] = w * (sb * qb.w - se * qe.w);
    quat_right_jac(qb, 0.5 * w * sb, Jb); quat_right_jac(qe, -0.5 * w * se, Ja);
  } else {
    double Ra[9], Rb[9], Rr[9], Re[9];
    quat_to_mat(qa, Ra); quat_to_mat(qb, Rb); quat_to_mat(qm, Rr);
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 3; ++c) Re[3 * rr + c] = Rr[3 * rr] * Ra[c] + Rr[3 * rr + 1] * Ra[3 + c] + Rr[3 * rr + 2] * Ra[6 + c];
    for (int c = 0; c < 3; ++c)
      for (int rr = 0; rr < 3; ++rr) r[3 * c + rr] = w * (Re[3 * rr + c] - Rb[3 * rr + c]);
    rot_right_jac(Re, w, Ja); rot_right_jac(Rb, -w, Jb);
  }
  double s = 0.0;
  for (int q = 0; q < kDim; ++q) { s += r[q] * r[q]; if (r_out) r_out[kDim * k + q] = r[q]; }
  if (rho) eval_loss(loss, s, rho + 3 * k);
  for (int side = 0; side < 2; ++side) {
    double* out = side ? Jj : Ji;
    if (!out) continue;
    const double* D = node_JL + 9 * (size_t)(side ? j : i);
    const double* J = side ? Jb : Ja;
    for (int q = 0; q < kDim; ++q)
      for (int c = 0; c < 3; ++c) out[3 * kDim * k + 3 * q + c] = J[3 * q] * D[c] + J[3 * q + 1] * D[3 + c] + J[3 * q + 2] * D[6 + c];
  }
}

__global__ void k_whiten_edges(uint64_t E, const double* __restrict__ cov6, const double* __restrict__ weight, int error_type, double* U9) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  double c6[6] = {0, 0, 0, 0, 0, 0}, u[6];
  if (cov6) for (int t = 0; t < 6; ++t) c6[t] = cov6[6 * k + t];
  whiten(error_type, c6, weight ? weight[k] : 1.0, u);
  double* o = U9 + 9 * k;
  o[0] = u[0]; o[1] = u[1]; o[2] = u[2]; o[3] = 0.0; o[4] = u[3]; o[5] = u[4]; o[6] = 0.0; o[7] = 0.0; o[8] = u[5];
}

__global__ void k_eval_loss(uint64_t n, const double* __restrict__ s, DevLoss loss, double* out) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double rho[3];
  eval_loss(loss, s[k], rho);
  out[3 * k] = rho[0]; out[3 * k + 1] = rho[1]; out[3 * k + 2] = rho[2];
}

// Tangent -> Euclidean export of the assembled system (API gsfm_ra_assemble).
__global__ void k_export_blocks(uint64_t H, int blk, const uint32_t* __restrict__ he_row, const uint32_t* __restrict__ he_col, const double* __restrict__ val,
                                const double* __restrict__ node_JL, const uint32_t* __restrict__ order, double* out_val, uint32_t* out_col) {
  // output position pos holds half-edge order[pos] (the (row, col)-sorted view of a column-blocked layout), or pos itself
  const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= H) return;
  const uint64_t h = order ? order[pos] : pos;
  const uint32_t row = he_row[h], col = he_col[h] & ~kSideBit;
  double B[9], T[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) B[3 * r + c] = blk_entry(val, h, blk, r, c);
  const double* Jr = node_JL + 9 * (size_t)row;
  const double* Jc = node_JL + 9 * (size_t)col;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = B[3 * r] * Jc[c] + B[3 * r + 1] * Jc[3 + c] + B[3 * r + 2] * Jc[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out_val[9 * pos + 3 * r + c] = Jr[r] * T[c] + Jr[3 + r] * T[3 + c] + Jr[6 + r] * T[6 + c];
  if (out_col) out_col[pos] = col;
}
__global__ void k_rowmajor_keys(uint64_t H, uint32_t N, const uint32_t* __restrict__ he_row, const uint32_t* __restrict__ he_col, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ idx) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  keys[h] = (uint64_t)he_row[h] * N + (he_col[h] & ~kSideBit);
  idx[h] = (uint32_t)h;
}
__global__ void k_export_nodes(uint32_t N, const double* __restrict__ Hd, const double* __restrict__ gt, const double* __restrict__ node_JL,
                               double* hdiag9, double* grad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* J = node_JL + 9 * (size_t)i;
  double He[6];
  congruence(J, Hd + 6 * (size_t)i, He);
  if (hdiag9) {
    double* o = hdiag9 + 9 * (size_t)i;
    o[0] = He[0]; o[1] = He[1]; o[2] = He[2]; o[3] = He[1]; o[4] = He[3]; o[5] = He[4]; o[6] = He[2]; o[7] = He[4]; o[8] = He[5];
  }
  if (grad) for (int c = 0; c < 3; ++c) grad[3 * (size_t)i + c] = J[c] * gt[3 * (size_t)i] + J[3 + c] * gt[3 * (size_t)i + 1] + J[6 + c] * gt[3 * (size_t)i + 2];
}
// v_out = Jl v (mode 0), Jl^T v (mode 1) [+ damp .* x2]
__global__ void k_node_apply(uint32_t N, const double* __restrict__ node_JL, const double* __restrict__ v, int mode, const double* __restrict__ damp,
                             const double* __restrict__ x2, double* out, int out_stride) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* J = node_JL + 9 * (size_t)i;
  const double a = v[3 * (size_t)i], b = v[3 * (size_t)i + 1], c = v[3 * (size_t)i + 2];
  for (int q = 0; q < 3; ++q) {
    double o = mode == 0 ? J[3 * q] * a + J[3 * q + 1] * b + J[3 * q + 2] * c : J[q] * a + J[3 + q] * b + J[6 + q] * c;
    if (damp) o += damp[3 * (size_t)i + q] * x2[3 * (size_t)i + q];
    out[(size_t)out_stride * i + q] = o;
  }
  if (out_stride == 4) out[4 * (size_t)i + 3] = 0.0;
}

// Sigma-consensus weights (rotation_estimator.cpp:378-418): w = (C3*2/sigma)(Gamma_tab[round(1000 r^2/(2 sigma^2))] - Gamma_k)
// from the angular residual at the current rotations; C++ round() = half away from zero; the table is exp(-x/1000).
// Also reduces sum |w - w_prev| (deterministic grid sum) into sc->dg.
__global__ void k_sigma_weights(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const double* __restrict__ omega_ij,
                                const double* __restrict__ node_q, double one_over_sigma, double sq_sigma_max_2, double gamma_k,
                                double weight_zero, double table_size, double* __restrict__ w, double* slots, unsigned* counter,
                                DevScalars* sc) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v[1] = {0.0};
  if (k < E) {
    const double4 a = reinterpret_cast<const double4*>(node_q)[ei[k]], b = reinterpret_cast<const double4*>(node_q)[ej[k]];
    const Q4 qi{a.x, a.y, a.z, a.w}, qj{b.x, b.y, b.z, b.w};
    const Q4 qm = aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
    const Q4 qE = qmul(qmul(qj, qconj(qi)), qconj(qm));
    double e[3], t2, c;
    quat_log(qE, e, &t2, &c);
    const double residual = sqrt(t2);
    double wk;
    if (residual < DBL_EPSILON) wk = weight_zero;
    else {
      double x = round(1000.0 * (residual * residual) / sq_sigma_max_2);
      if (table_size < x) x = table_size;
      wk = one_over_sigma * (exp(-x / 1000.0) - gamma_k);
    }
    v[0] = fabs(wk - w[k]);
    w[k] = wk;
  }
  double tot[1];
  if (grid_sum<1>(v, slots, counter, tot) && threadIdx.x == 0) sc->dg = tot[0];
}

// The step after the path: FilterViewPairsFromOrientation (T/sfm/filter_view_pairs_from_orientation.cc:55-118).
__global__ void k_filter_pairs(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const double* __restrict__ omega_ij,
                               const double* __restrict__ node_q, double sq_threshold, uint8_t* keep, double* angle) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  const double4 a = reinterpret_cast<const double4*>(node_q)[ei[k]], b = reinterpret_cast<const double4*>(node_q)[ej[k]];
  const Q4 qi{a.x, a.y, a.z, a.w}, qj{b.x, b.y, b.z, b.w};
  const Q4 qm = aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
  const Q4 qE = qmul(qmul(qj, qconj(qi)), qconj(qm));
  double e[3], t2, c;
  quat_log(qE, e, &t2, &c);
  if (angle) angle[k] = sqrt(t2);
  if (keep) keep[k] = (t2 <= sq_threshold) ? 1 : 0;
}

}  // namespace