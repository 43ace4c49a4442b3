// ra_dense.cuh -- on-device dense Cholesky for small view graphs
// Part of libgsfm_ra (one translation unit, see gsfm_ra.cu); reference citations sit next to each kernel.
#pragma once
#include "ra_common.cuh"
namespace {

// ------------------------------------------------------------------------------------------
// Small problems: exact dense LL^T of (Ht + Lam) on the device (GSFM_RA_SOLVER_DENSE_CHOLESKY), the
// role SPARSE_NORMAL_CHOLESKY plays in the reference (rotation_estimator.cpp:300).  For a view graph
// like Madrid_Metropolis (379 views, 26 % dense, covariance weights spanning 12 decades) PCG needs
// hundreds of steps per solve; the 1137 x 1137 factorisation does not care.
// One cooperative kernel, right-looking blocked factorisation with 32 x 32 tiles:
//   per panel: every CTA factors the diagonal tile in shared memory (redundantly: no barrier for it),
//   the tiles below are solved against it, barrier, the trailing tiles are updated, barrier;
//   then forward/backward substitution by CTA 0.
// A is column-major, lower triangle, n padded to a multiple of 32 with a unit diagonal.
// ------------------------------------------------------------------------------------------
constexpr int kNB = 32;

// The right-hand side rides along as an EXTRA ROW of the matrix (row index n, inside the padding): factoring
// [A b; b^T beta] = [L 0; y^T *][L^T y; 0 *] leaves y = L^-1 b in that row, so the forward substitution costs nothing.
// Entry (r, c) of the stored off-diagonal block of half-edge h (blk = 6: packed symmetric; 9: row-major).
__device__ __forceinline__ double blk_entry(const double* recs, uint64_t h, int blk, int r, int c) {
  int k;
  if (blk == 4) {  // compact scalar-weight stencil: -( |c0| delta_rc + sign(c0) h_r h_c )
    const int rd = blk * 32 + 16;
    const double c0 = recs[blk_index(h, 0, rd)];
    const double hh = recs[blk_index(h, 1 + r, rd)] * recs[blk_index(h, 1 + c, rd)];
    return -((r == c ? fabs(c0) : 0.0) + (signbit(c0) ? -hh : hh));
  }
  if (blk == 6) { const int a = r < c ? r : c, b2 = r < c ? c : r; k = a * 3 - a * (a - 1) / 2 + (b2 - a); }
  else k = 3 * r + c;
  return recs[blk_index(h, k, blk * 32 + 16)];
}

__global__ void k_dense_assemble(uint64_t H, uint32_t N, uint32_t np, int blk, const uint32_t* __restrict__ he_row, const uint32_t* __restrict__ he_col,
                                 const double* __restrict__ recs, const double* __restrict__ Dblk, const double* __restrict__ rhs,
                                 double* __restrict__ A) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 3 * N;
  if (t < H) {
    const uint32_t row = he_row[t], col = he_col[t] & ~kSideBit;
    if (row > col) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) A[(size_t)(3 * col + c) * np + 3 * row + r] = blk_entry(recs, t, blk, r, c);
    }
  }
  if (t < N) {
    const double* D = Dblk + 6 * (size_t)t;
    const size_t o = 3 * (size_t)t;
    A[o * np + o] = D[0]; A[o * np + o + 1] = D[1]; A[o * np + o + 2] = D[2];
    A[(o + 1) * np + o + 1] = D[3]; A[(o + 1) * np + o + 2] = D[4];
    A[(o + 2) * np + o + 2] = D[5];
  }
  if (t < n) A[(size_t)t * np + n] = rhs[t];
  if (t == n) A[(size_t)t * np + t] = 1e200;
  if (t > n && t < np) A[(size_t)t * np + t] = 1.0;
}

// 32 x 32 lower-triangular tile in shared memory: factor it (LL^T) and invert the factor, one warp, everything
// in registers with compile-time indices.  Lt <- L, Wt <- L^-1.
__device__ __forceinline__ void tile_potrf_inv(double (*Lt)[kNB + 1], double (*Wt)[kNB + 1], int* fail) {
  const int lane = threadIdx.x & 31;
  double row[kNB];
#pragma unroll
  for (int c = 0; c < kNB; ++c) row[c] = Lt[lane][c];
#pragma unroll
  for (int j = 0; j < kNB; ++j) {
    const double d = __shfl_sync(0xffffffffu, row[j], j);
    if (!(d > 0.0) && lane == 0) *fail = 1;
    const double inv = rsqrt(d > 0.0 ? d : 1.0);
    if (lane == j) row[j] = d * inv;
    else if (lane > j) row[j] *= inv;
    const double lj = row[j];
#pragma unroll
    for (int c = 0; c < kNB; ++c) {
      if (c > j) {
        const double lc = __shfl_sync(0xffffffffu, lj, c);
        if (lane >= c) row[c] -= lj * lc;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kNB; ++c) Lt[lane][c] = (c <= lane) ? row[c] : 0.0;
  __syncwarp();
  // column `lane` of X = L^-1 by forward substitution in saxpy form: once x[m] is known every remaining row takes its
  // update independently (no serial dot products), L is read from shared memory as broadcasts, and the 32 reciprocals of
  // the diagonal are formed in parallel (lane m holds L[m][m]) instead of one division per step.
  double dg = 1.0;  // L[lane][lane] (select chain: a dynamic index would push row[] to local memory)
#pragma unroll
  for (int c = 0; c < kNB; ++c)
    if (c == lane) dg = row[c];
  const double rdiag = 1.0 / dg;
  double x[kNB];
#pragma unroll
  for (int i = 0; i < kNB; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < kNB; ++m) {
    const double xm = x[m] * __shfl_sync(0xffffffffu, rdiag, m);
    x[m] = xm;
#pragma unroll
    for (int i = 0; i < kNB; ++i)
      if (i > m) x[i] -= Lt[i][m] * xm;
  }
#pragma unroll
  for (int i = 0; i < kNB; ++i) Wt[i][lane] = x[i];
}

__global__ void __launch_bounds__(kBlock) k_dense_cholesky_solve(uint32_t n, uint32_t np, double* __restrict__ A, double* __restrict__ x,
                                                                  double* __restrict__ winv, double* __restrict__ work, DevScalars* sc) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double L11[kNB][kNB + 1];
  __shared__ double W11[kNB][kNB + 1];
  __shared__ double T1[kNB][kNB + 1];
  __shared__ double T2[kNB][kNB + 1];
  __shared__ int s_fail;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;  // 32 x 8
  const uint32_t nblk = np / kNB;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (uint32_t kb = 0; kb < nblk; ++kb) {
    // (1) diagonal tile -> shared; factor + invert (every CTA, redundantly: no barrier needed for it)
    for (int c = ty; c < kNB; c += 8) L11[tx][c] = A[(size_t)(kb * kNB + c) * np + kb * kNB + tx];
    __syncthreads();
    if (tid < 32) tile_potrf_inv(L11, W11, &s_fail);
    __syncthreads();
    // (2) panel: L[ib][kb] = A[ib][kb] W^T   (W = L11^-1, lower triangular)
    for (uint32_t ib = kb + 1 + blockIdx.x; ib < nblk; ib += gridDim.x) {
      for (int c = ty; c < kNB; c += 8) T1[tx][c] = A[(size_t)(kb * kNB + c) * np + ib * kNB + tx];
      __syncthreads();
      {  // k outer: one T1 read (2 wavefronts) serves the thread's four outputs, the W11 reads are broadcasts
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int k = 0; k < kNB; ++k) {
          const double a = T1[tx][k];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] += a * W11[ty + 8 * q][k];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) A[(size_t)(kb * kNB + ty + 8 * q) * np + ib * kNB + tx] = acc[q];
      }
      __syncthreads();
    }
    grid.sync();
    if (blockIdx.x == 0) {  // factored tile + its inverse back to global (after the barrier: others were still loading it)
      for (int c = ty; c < kNB; c += 8) {
        if (tx >= c) A[(size_t)(kb * kNB + c) * np + kb * kNB + tx] = L11[tx][c];
        winv[(size_t)kb * kNB * kNB + c * kNB + tx] = W11[tx][c];  // winv[kb][col c][row tx]
      }
    }
    // (3) trailing update: A[i][j] -= L[i][kb] L[j][kb]^T for kb < j <= i
    const uint32_t m = nblk - kb - 1;
    const uint32_t ntiles = m * (m + 1) / 2;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      uint32_t i = (uint32_t)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((uint64_t)(i + 1) * (i + 2) / 2 <= t) ++i;
      while ((uint64_t)i * (i + 1) / 2 > t) --i;
      const uint32_t j = t - i * (i + 1) / 2;
      const uint32_t ib = kb + 1 + i, jb = kb + 1 + j;
      double cur[4];  // the tile being updated: loaded with the operands, not after the products
#pragma unroll
      for (int q = 0; q < 4; ++q) cur[q] = A[(size_t)(jb * kNB + ty + 8 * q) * np + ib * kNB + tx];
      for (int c = ty; c < kNB; c += 8) {
        T1[tx][c] = A[(size_t)(kb * kNB + c) * np + ib * kNB + tx];
        T2[tx][c] = A[(size_t)(kb * kNB + c) * np + jb * kNB + tx];
      }
      __syncthreads();
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int k = 0; k < kNB; ++k) {
          const double a = T1[tx][k];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] += a * T2[ty + 8 * q][k];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) A[(size_t)(jb * kNB + ty + 8 * q) * np + ib * kNB + tx] = cur[q] - acc[q];
      }
      __syncthreads();
    }
    grid.sync();
  }
  // (4) back substitution L^T x = y by CTA 0, y = row n of the factor; inverted diagonal tiles make each step a
  //     column dot-product sweep + a 32 x 32 mat-vec, no serial recurrence
  if (blockIdx.x == 0) {
    for (uint32_t i = tid; i < np; i += kBlock) work[i] = (i < n) ? A[(size_t)i * np + n] : 0.0;
    __syncthreads();
    for (int kbi = (int)nblk - 1; kbi >= 0; --kbi) {
      const uint32_t kb = (uint32_t)kbi;
      for (int c = ty; c < kNB; c += 8) W11[tx][c] = winv[(size_t)kb * kNB * kNB + c * kNB + tx];
      {
        double acc = 0.0;  // column kb*32+tx: sum over rows below the tile (rows >= n carry no unknowns)
        for (uint32_t r = (kb + 1) * kNB + ty; r < n; r += 8) acc += A[(size_t)(kb * kNB + tx) * np + r] * work[r];
        T2[ty][tx] = acc;
      }
      __syncthreads();
      if (tid < 32) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += T2[q][tid];
        T1[0][tid] = (kb * kNB + tid < n) ? work[kb * kNB + tid] - s : 0.0;
      }
      __syncthreads();
      if (tid < 32) {  // x_tile = W^T rhs
        double xv = 0.0;
#pragma unroll 8
        for (int k = 0; k < kNB; ++k) xv += W11[k][tid] * T1[0][k];
        work[kb * kNB + tid] = (kb * kNB + tid < n) ? xv : 0.0;
      }
      __syncthreads();
    }
    for (uint32_t i = tid; i < n; i += kBlock) x[i] = work[i];
    if (tid == 0) { sc->pcg_iter = 1; sc->pcg_done = 1; sc->pcg_breakdown = s_fail; sc->rr = 0.0; }
  }
}

}  // namespace