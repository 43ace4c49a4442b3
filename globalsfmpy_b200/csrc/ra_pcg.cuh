// ra_pcg.cuh -- K2 block SpMV and the block-Jacobi PCG (persistent cooperative kernel, multi-GPU exchange)
// Part of libgsfm_ra (one translation unit, see gsfm_ra.cu); reference citations sit next to each kernel.
#pragma once
#include "ra_common.cuh"
#include "ra_structure.cuh"
#include "ra_edges.cuh"
namespace {

// Per view, before a linear solve at trust-region radius mu:
//   LM diagonal in scaled coordinates  d_c = clamp(ediag_c s_c^2, lo, hi) / mu           (LevenbergMarquardtStrategy)
//   as damping of the unscaled Euclidean system  lam_c = d_c / s_c^2
//   moved to the tangent frame  Lam = Jl^-T diag(lam) Jl^-1 ;  Dblk = Hd + Lam ; Minv = Dblk^-1.
// Also initialises PCG: x = 0, r = b = -gt, z = Minv r, p = z, q = 0, and reduces rz, bb.  z and p are the vectors
// the SpMV gathers: stored with stride 4 (double4), everything else with stride 3.
struct PrepareArgs {
  double mu, lo, hi;
  const double *ediag, *scale, *node_JL, *Hd, *gt, *user_damp, *user_b;
  double *Dblk, *Minv, *x, *r, *z, *p, *q, *bvec;
};
// per-view body of k_prepare_solve; adds the view's (b.z, b.b) to v
__device__ __forceinline__ void prepare_view(uint32_t i, const PrepareArgs& A, double (&v)[2]) {
  // every load first (the arrays may alias as far as the compiler knows: a store between two loads serialises the
  // L2 round trips)
  double lam[3], J[9], Hd[6], b[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (A.user_damp) lam[c] = A.user_damp[3 * (size_t)i + c];
    else {
      const double s2 = A.scale[3 * (size_t)i + c] * A.scale[3 * (size_t)i + c];
      lam[c] = fmin(fmax(A.ediag[3 * (size_t)i + c] * s2, A.lo), A.hi) / A.mu / s2;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) J[k] = A.node_JL[9 * (size_t)i + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) Hd[k] = A.Hd[6 * (size_t)i + k];
#pragma unroll
  for (int c = 0; c < 3; ++c) b[c] = A.user_b ? A.user_b[3 * (size_t)i + c] : -A.gt[3 * (size_t)i + c];
  double Ji[9];
  inv3(J, Ji);
  double lamS[6] = {lam[0], 0.0, 0.0, lam[1], 0.0, lam[2]};
  double Lam[6];
  congruence(Ji, lamS, Lam);
  double D[6], M[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) D[k] = Hd[k] + Lam[k];
  sym_inv(D, M);
  if (A.user_b) {  // b given in Euclidean coordinates: bt = Jl^-T b
    const double u0 = b[0], u1 = b[1], u2 = b[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) b[c] = Ji[c] * u0 + Ji[3 + c] * u1 + Ji[6 + c] * u2;
  }
  double zz[3];
  sym_mul_vec(M, b, zz);
#pragma unroll
  for (int k = 0; k < 6; ++k) { A.Dblk[6 * (size_t)i + k] = D[k]; A.Minv[6 * (size_t)i + k] = M[k]; }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    A.x[3 * (size_t)i + c] = 0.0; A.r[3 * (size_t)i + c] = b[c]; A.bvec[3 * (size_t)i + c] = b[c]; A.q[3 * (size_t)i + c] = 0.0;
    v[0] += b[c] * zz[c];
    v[1] += b[c] * b[c];
  }
  // the gathered vectors are padded to one aligned 32 B sector per view
  reinterpret_cast<double4*>(A.z)[i] = make_double4(zz[0], zz[1], zz[2], 0.0);
  reinterpret_cast<double4*>(A.p)[i] = make_double4(zz[0], zz[1], zz[2], 0.0);
}

__global__ void k_prepare_solve(uint32_t N, PrepareArgs A, double* slots, unsigned* counter, DevScalars* sc, const IterParams* ip) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip) A.mu = ip->mu;
  if (i == 0) sc->t_begin = gtimer_ns();
  double v[2] = {0.0, 0.0};
  if (i < N) prepare_view(i, A, v);
  double tot[2];
  if (grid_sum<2>(v, slots, counter, tot) && threadIdx.x == 0) {
    sc->rz = tot[0]; sc->bb = tot[1]; sc->rr = tot[1];
    sc->pcg_iter = 0; sc->pcg_breakdown = 0;
    sc->pcg_done = (tot[1] == 0.0 || !isfinite(tot[1])) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// K2: block-3x3 CSR SpMV, off-diagonal part, TMA-staged.
// Every warp owns one contiguous range of chunk records.  Lane 0 keeps kStages bulk async copies
// (cp.async.bulk, one 2432 B record each, completion on a warp-private mbarrier) in flight; the
// warp consumes a record from shared memory (conflict-free: lane l reads word l of each of the 9
// component rows), gathers x[col] (24 B, L2) and accumulates.  Bytes in flight are set by the ring
// depth, not by registers or occupancy.  Row boundaries inside a range are handled by visiting
// the range segment by segment (segment = range ^ row); a record shared by two segments is read
// twice from shared memory, never twice from HBM.
// ------------------------------------------------------------------------------------------
// Stream this warp's record range (nrec records from half-edge lo) and call finish(t, y0, y1, y2)
// (all lanes, totals valid in every lane) for each segment t in [t0, t1).  x4 is the gathered vector, one
// aligned double4 (32 B sector) per view.  Record-major loop: the x gather of record c+1 (its columns are
// already in shared memory) is issued before record c is consumed, so the L2 gather latency overlaps the
// arithmetic and the next wait; each lane's 3-vector contribution is formed once per record and added to
// the running segment, segments that end inside the record are reduced and handed to finish().
template <int kBlk, typename Finish>
__device__ __forceinline__ void spmv_stream(WarpPipe& wp, const double* __restrict__ recs, uint64_t lo, uint64_t hi, uint32_t t0, uint32_t t1,
                                            const uint32_t* __restrict__ seg_begin, const uint32_t* __restrict__ seg_len, const double* x4,
                                            const double* xs /*shared-memory copy of x[xbase ...], stride 3, or null*/, uint32_t xbase, bool nogather,
                                            uint32_t keep8, Finish&& finish, bool wrap = false, bool stop_after_priming = false) {
  // CHUNK-major loop: the kCR records of a bulk copy are consumed together with compile-time indices -- their shared-memory
  // reads, gathers and products are independent instruction streams (ILP = kCR), the loop / ring / address overhead is paid
  // once per chunk instead of once per record, and the x gather of chunk k+1 (its columns are in shared memory as soon as
  // its copy has landed) is issued before chunk k is consumed.  Positions are 32-bit offsets from the start of the range.
  constexpr int kRD = Rec<kBlk>::kDoubles, kCR = Chunk<kBlk>::kRecs, kCD = Chunk<kBlk>::kDoubles;
  const int lane = threadIdx.x & 31;
  const uint32_t n = (uint32_t)(hi - lo);           // half-edges of this range
  const uint32_t nrec = (n + 31u) >> 5;
  const uint32_t nchunk = (nrec + kCR - 1) / kCR;   // bulk copies of this range: kCR consecutive records each (the last may be short)
  const uint32_t base = wp.pos;
  const double* src = recs + (size_t)(lo >> 5) * kRD;
  uint64_t pol_keep = 0, pol_stream = 0;
  if (keep8) { pol_keep = l2_policy_evict_last(); pol_stream = l2_policy_evict_first(); }
  const uint32_t chunk0 = (uint32_t)(lo >> 5) / kCR;
  // WRAP-AROUND PREFETCH (persistent PCG kernel): every pass streams the same range, and the matrix does not depend on the
  // vector being multiplied -- so when the last chunks of a pass have been consumed their ring stages are refilled with the
  // FIRST chunks of the next pass.  Those copies land while the block sits in the two grid barriers and the vector phase of the
  // CG step (the ring of an SM holds ~166 KB: a third of the whole matrix at 1M edges), and the next pass starts on data that
  // is already in shared memory.  Chunk index k >= nchunk means chunk k - nchunk of the next pass; ring positions keep counting.
  const bool wrap_ok = wrap && nchunk >= (uint32_t)kStages2;
  auto issue = [&](uint32_t k) {
    if (lane == 0) {
      const uint32_t st = (base + k) % kStages2;
      const uint32_t kk = k < nchunk ? k : k - nchunk;
      const uint32_t bytes = min((uint32_t)kCR, nrec - kk * kCR) * (uint32_t)Rec<kBlk>::kBytes;
      mbar_expect_tx(&wp.bars[st], bytes);
      if (keep8)
        tma_load_bulk_hint(wp.ring + (size_t)st * kCD, src + (size_t)kk * kCD, bytes, &wp.bars[st], ((chunk0 + kk) & 7u) < keep8 ? pol_keep : pol_stream);
      else
        tma_load_bulk(wp.ring + (size_t)st * kCD, src + (size_t)kk * kCD, bytes, &wp.bars[st]);
    }
  };
  // first record of chunk k, once its bulk copy has landed
  auto wait_chunk = [&](uint32_t k) -> const double* {
    const uint32_t p = base + k, st = p % kStages2;
    mbar_wait(&wp.bars[st], (p / kStages2) & 1u);
    return wp.ring + (size_t)st * kCD;
  };
  // x[col] of every record of chunk k (records past the end of the range and padding lanes read view 0: never accumulated)
  auto gather_chunk = [&](const double* ch, uint32_t k, double (&g)[kCR][3]) {
#pragma unroll
    for (int j = 0; j < kCR; ++j) {
      const uint32_t c = k * kCR + j;
      uint32_t col = xbase;  // padding: any view that is certainly inside the staged slice
      if (c < nrec) {
        col = reinterpret_cast<const uint32_t*>(ch + j * kRD + Rec<kBlk>::kColOffset)[lane] & ~kSideBit;
        if ((c << 5) + (uint32_t)lane >= n) col = xbase;
      }
      if (nogather) { g[j][0] = col; g[j][1] = 1.0; g[j][2] = 2.0; }  // measurement aid: stream-only ceiling
      else if (xs) { const double* p = xs + 3 * (size_t)(col - xbase); g[j][0] = p[0]; g[j][1] = p[1]; g[j][2] = p[2]; }
      else { const double4 xv = reinterpret_cast<const double4*>(x4)[col]; g[j][0] = xv.x; g[j][1] = xv.y; g[j][2] = xv.z; }
    }
  };
  for (uint32_t k = wp.primed; k < nchunk && k < (uint32_t)kStages2; ++k) issue(k);  // wp.primed of them went out at the end of the previous pass
  wp.primed = 0;
  if (stop_after_priming) { wp.primed = min(nchunk, (uint32_t)kStages2); return; }
  if (nrec == 0 || t0 == t1) { wp.pos = base + nchunk; return; }
  uint32_t t = t0;
  uint32_t sb = (uint32_t)(seg_begin[t] - lo), se = sb + seg_len[t];
  double y0 = 0.0, y1 = 0.0, y2 = 0.0;
  if (kBlk == 4 && kRecordMajor4) {
    // RECORD-major loop for the compact records: one record per iteration, the gather of record c+1 issued before record c is
    // consumed.  Measured (profiles/r02_col_blocks.txt): on the L2-resident 1M-edge matrix this form is ~2 us per pass faster
    // than the chunk-major one below (fewer live registers in the persistent kernel, gathers spread more evenly), which in turn
    // is 10 % faster on the HBM-resident 20M-edge matrix (6-double records).
    auto gather1 = [&](const double* rec, uint32_t c, double& x0, double& x1, double& x2) {
      uint32_t col = reinterpret_cast<const uint32_t*>(rec + Rec<kBlk>::kColOffset)[lane] & ~kSideBit;
      if ((c << 5) + (uint32_t)lane >= n) col = xbase;  // padding lanes of the last record
      if (nogather) { x0 = col; x1 = 1.0; x2 = 2.0; }
      else if (xs) { const double* p = xs + 3 * (size_t)(col - xbase); x0 = p[0]; x1 = p[1]; x2 = p[2]; }
      else { const double4 xv = reinterpret_cast<const double4*>(x4)[col]; x0 = xv.x; x1 = xv.y; x2 = xv.z; }
    };
    const double* rec = wait_chunk(0);
    double x0, x1, x2;
    gather1(rec, 0, x0, x1, x2);
    for (uint32_t c = 0; c < nrec; ++c) {
      const uint32_t cb = c << 5, ce = cb + 32u, h = cb + (uint32_t)lane;
      const double* rec_n = rec;
      double n0 = 0.0, n1 = 0.0, n2 = 0.0;
      if (c + 1 < nrec) {
        rec_n = ((c + 1) % kCR == 0) ? wait_chunk((c + 1) / kCR) : rec + kRD;
        gather1(rec_n, c + 1, n0, n1, n2);
      }
      const double c0 = rec[lane], h0 = rec[32 + lane], h1 = rec[64 + lane], h2 = rec[96 + lane];
      const double tt = h0 * x0 + h1 * x1 + h2 * x2;
      const double a = fabs(c0);
      const double st = __longlong_as_double(__double_as_longlong(tt) ^ (__double_as_longlong(c0) & (long long)0x8000000000000000ull));
      const double v0 = -(a * x0 + st * h0), v1 = -(a * x1 + st * h1), v2 = -(a * x2 + st * h2);
      if ((c + 1) % kCR == 0 || c + 1 == nrec) {  // a chunk's stage can be refilled as soon as every lane has read its last record
        __syncwarp();
        if (c / kCR + kStages2 < nchunk || wrap_ok) issue(c / kCR + kStages2);
      }
      while (true) {
        if (h >= sb && h < se) { y0 += v0; y1 += v1; y2 += v2; }
        if (se > ce) break;  // the segment continues in the next record
        finish(t, warp_sum_multi3(y0, y1, y2));
        y0 = y1 = y2 = 0.0;
        if (++t == t1) break;
        sb = se; se = sb + seg_len[t];
        if (sb >= ce) break;
      }
      rec = rec_n; x0 = n0; x1 = n1; x2 = n2;
      if (t == t1) break;
    }
    wp.pos = base + nchunk;
    wp.primed = wrap_ok ? (uint32_t)kStages2 : 0u;
    return;
  }
  const double* ch = wait_chunk(0);
  double g[kCR][3];
  gather_chunk(ch, 0, g);
  for (uint32_t k = 0; k < nchunk; ++k) {
    // the next chunk's gather goes out first and completes under this chunk's products and bookkeeping (measured, r02j: issuing
    // it after the products is neutral at 1M edges / L2 resident and 6 % slower at 20M edges / HBM resident)
    const double* ch_n = ch;
    double gn[kCR][3];
    if (k + 1 < nchunk) { ch_n = wait_chunk(k + 1); gather_chunk(ch_n, k + 1, gn); }
    double v[kCR][3];
#pragma unroll
    for (int j = 0; j < kCR; ++j) {
      const double* rec = ch + j * kRD;
      const double x0 = g[j][0], x1 = g[j][1], x2 = g[j][2];
      if (kBlk == 4) {
        // -S x with S = |c0| I + sign(c0) h h^T
        const double c0 = rec[lane], h0 = rec[32 + lane], h1 = rec[64 + lane], h2 = rec[96 + lane];
        const double tt = h0 * x0 + h1 * x1 + h2 * x2;
        const double a = fabs(c0);
        const double st = __longlong_as_double(__double_as_longlong(tt) ^ (__double_as_longlong(c0) & (long long)0x8000000000000000ull));
        v[j][0] = -(a * x0 + st * h0);
        v[j][1] = -(a * x1 + st * h1);
        v[j][2] = -(a * x2 + st * h2);
      } else if (kBlk == 6) {
        const double b0 = rec[lane], b1 = rec[32 + lane], b2 = rec[64 + lane], b3 = rec[96 + lane], b4 = rec[128 + lane], b5 = rec[160 + lane];
        v[j][0] = b0 * x0 + b1 * x1 + b2 * x2;
        v[j][1] = b1 * x0 + b3 * x1 + b4 * x2;
        v[j][2] = b2 * x0 + b4 * x1 + b5 * x2;
      } else {
        v[j][0] = rec[lane] * x0 + rec[32 + lane] * x1 + rec[64 + lane] * x2;
        v[j][1] = rec[96 + lane] * x0 + rec[128 + lane] * x1 + rec[160 + lane] * x2;
        v[j][2] = rec[192 + lane] * x0 + rec[224 + lane] * x1 + rec[256 + lane] * x2;
      }
    }
    // the chunk's stage can be refilled as soon as every lane has read it
    __syncwarp();
    if (k + kStages2 < nchunk || wrap_ok) issue(k + kStages2);
#pragma unroll
    for (int j = 0; j < kCR; ++j) {
      const uint32_t cb = (k * kCR + j) << 5, ce = cb + 32u, h = cb + (uint32_t)lane;
      if (t != t1 && cb < n) {
        while (true) {
          if (h >= sb && h < se) { y0 += v[j][0]; y1 += v[j][1]; y2 += v[j][2]; }
          if (se > ce) break;  // the segment continues in the next record
          // lanes 0 / 8 / 16 receive the segment's three sums (component multi_index4(lane))
          finish(t, warp_sum_multi3(y0, y1, y2));
          y0 = y1 = y2 = 0.0;
          if (++t == t1) break;
          sb = se; se = sb + seg_len[t];
          if (sb >= ce) break;
        }
      }
    }
    ch = ch_n;
#pragma unroll
    for (int j = 0; j < kCR; ++j) { g[j][0] = gn[j][0]; g[j][1] = gn[j][1]; g[j][2] = gn[j][2]; }
  }
  wp.pos = base + nchunk;
  wp.primed = wrap_ok ? (uint32_t)kStages2 : 0u;
}

// Wait for the bulk copies a wrap-around prefetch left in flight (a block must not exit with copies into its shared memory pending).
__device__ __forceinline__ void spmv_drain(WarpPipe& wp) {
  for (uint32_t k = 0; k < wp.primed; ++k) {
    const uint32_t p = wp.pos + k;
    mbar_wait(&wp.bars[p % kStages2], (p / kStages2) & 1u);
  }
  wp.pos += wp.primed;
  wp.primed = 0;
}

// Measurement aid (gsfm_ra_measure_stream): the K2 ring alone -- every warp pulls its share of `nchunks` chunks of kChunkBytes
// through the same kStages2-deep bulk-copy pipeline and adds one word per lane and 256 bytes, no gather, no reduction.
template <int kChunkBytes>
__global__ void __launch_bounds__(kPcgBlock, kPcgBlocksPerSM) k_stream_probe(const unsigned char* __restrict__ src, uint32_t nchunks, double* sink) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WarpPipe wp;
  pipe_init_bytes<kChunkBytes, kStages2>(wp, smem_raw);
  const int lane = threadIdx.x & 31;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t per = (nchunks + nw - 1) / nw, lo = gw * per, n = lo >= nchunks ? 0u : min(per, nchunks - lo);
  auto issue = [&](uint32_t k) {
    if (lane == 0 && k < n) {
      mbar_expect_tx(&wp.bars[k % kStages2], kChunkBytes);
      tma_load_bulk(reinterpret_cast<unsigned char*>(wp.ring) + (size_t)(k % kStages2) * kChunkBytes, src + (size_t)(lo + k) * kChunkBytes, kChunkBytes,
                    &wp.bars[k % kStages2]);
    }
  };
  for (uint32_t k = 0; k < (uint32_t)kStages2; ++k) issue(k);
  double acc = 0.0;
  for (uint32_t k = 0; k < n; ++k) {
    mbar_wait(&wp.bars[k % kStages2], (k / kStages2) & 1u);
    const double* r = wp.ring + (size_t)(k % kStages2) * (kChunkBytes / 8);
#pragma unroll
    for (int j = 0; j < kChunkBytes / 256; ++j) acc += r[j * 32 + lane];
    __syncwarp();
    issue(k + kStages2);
  }
  if (acc == 123.456) sink[0] = acc;
}

template <int kBlk>
__global__ void __launch_bounds__(kPcgBlock, kPcgBlocksPerSM)
k_spmv(uint32_t num_warps, uint64_t H, uint32_t warp_span, const uint32_t* __restrict__ warp_seg_ptr, const uint32_t* __restrict__ task_begin,
       const uint32_t* __restrict__ task_len, const double* __restrict__ recs, const double* __restrict__ x4, double* __restrict__ ypart,
       const DevScalars* sc, int check_done, uint32_t keep8) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (check_done == 1 && sc->pcg_done) return;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= num_warps) return;
  WarpPipe wp;
  pipe_init<kBlk>(wp, smem_raw);
  const uint64_t lo = (uint64_t)gw * warp_span, hi = min(H, lo + warp_span);
  spmv_stream<kBlk>(wp, recs, lo, hi, warp_seg_ptr[gw], warp_seg_ptr[gw + 1], task_begin, task_len, x4, nullptr, 0u, check_done == 2, keep8,
                    [&](uint32_t t, double ysum) {
                      if ((lane & 7) == 0 && lane < 24) ypart[3 * (size_t)t + multi_index4(lane)] = ysum;
                    });
}

// y_i = Dblk_i x_i + sum of the row's task partials (+ shard-local only: the diagonal part is
// added after the cross-GPU reduction).  mode 0: write y, reduce p.y -> alpha (PCG step 1).
// mode 1: y only.
__global__ void k_spmv_finish(uint32_t N, uint32_t ncb, const uint32_t* __restrict__ node_task_ptr, const double* __restrict__ ypart,
                              const double* __restrict__ Dblk, const double* __restrict__ x, double* y, const double* ysum,
                              int mode, double* slots, unsigned* counter, DevScalars* sc) {
  // ysum != null: the off-diagonal part was already summed (and all-reduced across GPUs) into ysum
  if (mode == 0 && sc->pcg_done) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[1] = {0.0};
  if (i < N) {
    double xi[3] = {x[4 * (size_t)i], x[4 * (size_t)i + 1], x[4 * (size_t)i + 2]};  // x is a gathered vector: stride 4
    double yi[3] = {0.0, 0.0, 0.0};
    if (Dblk) sym_mul_vec(Dblk + 6 * (size_t)i, xi, yi);
    if (ysum) { yi[0] += ysum[3 * (size_t)i]; yi[1] += ysum[3 * (size_t)i + 1]; yi[2] += ysum[3 * (size_t)i + 2]; }
    else
      for (uint32_t cb = 0; cb < ncb; ++cb)
        for (uint32_t t = node_task_ptr[cb * N + i]; t < node_task_ptr[cb * N + i + 1]; ++t) {
          yi[0] += ypart[3 * (size_t)t]; yi[1] += ypart[3 * (size_t)t + 1]; yi[2] += ypart[3 * (size_t)t + 2];
        }
    y[3 * (size_t)i] = yi[0]; y[3 * (size_t)i + 1] = yi[1]; y[3 * (size_t)i + 2] = yi[2];
    v[0] = xi[0] * yi[0] + xi[1] * yi[1] + xi[2] * yi[2];
  }
  if (mode != 0) return;
  double tot[1];
  if (grid_sum<1>(v, slots, counter, tot) && threadIdx.x == 0) {
    sc->pAp = tot[0];
    if (!(tot[0] > 0.0) || !isfinite(tot[0])) { sc->pcg_done = 1; sc->pcg_breakdown = 1; sc->alpha = 0.0; }
    else sc->alpha = sc->rz / tot[0];
  }
}

// PCG step 2: x += alpha p ; r -= alpha y ; z = Minv r ; reduce r.z, r.r -> beta, convergence.
__global__ void k_pcg_update(uint32_t N, const double* __restrict__ Minv, const double* __restrict__ p, const double* __restrict__ y,
                             double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double rtol2, int max_iter,
                             double* slots, unsigned* counter, DevScalars* sc) {
  if (sc->pcg_done) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  if (i < N) {
    const double alpha = sc->alpha;
    double ri[3], zi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      x[3 * (size_t)i + c] += alpha * p[4 * (size_t)i + c];
      ri[c] = r[3 * (size_t)i + c] - alpha * y[3 * (size_t)i + c];
      r[3 * (size_t)i + c] = ri[c];
    }
    sym_mul_vec(Minv + 6 * (size_t)i, ri, zi);
#pragma unroll
    for (int c = 0; c < 3; ++c) { z[4 * (size_t)i + c] = zi[c]; v[0] += ri[c] * zi[c]; v[1] += ri[c] * ri[c]; }
  }
  double tot[2];
  if (grid_sum<2>(v, slots, counter, tot) && threadIdx.x == 0) {
    sc->beta = tot[0] / sc->rz;
    sc->rz = tot[0];
    sc->rr = tot[1];
    sc->pcg_iter += 1;
    if (tot[1] <= rtol2 * sc->bb || sc->pcg_iter >= max_iter || !isfinite(tot[1])) sc->pcg_done = 1;
  }
}

// PCG step 3: p = z + beta p (both stride 4; the pad element stays 0).
__global__ void k_pcg_direction(uint32_t n3, const double* __restrict__ z, double* __restrict__ p, const DevScalars* sc) {
  if (sc->pcg_done) return;
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n3) p[c] = z[c] + sc->beta * p[c];
}

// ------------------------------------------------------------------------------------------
// The whole block-Jacobi PCG solve as ONE persistent cooperative kernel (one launch per linear
// solve, convergence decided on the device, two grid barriers per CG step).
//
// CG in the Chronopoulos-Gear arrangement: the matrix is applied to the preconditioned residual z (ONE gathered
// vector, one aligned 32 B sector per half-edge), the search direction and its image follow by recurrence
// (p = z + beta p, q = s + beta q, q = A p), and both inner products of a step come out of the same pass:
//   phase A  every warp streams its range of records: spart[seg] = sum blk z[col].  The warp that owns a row
//            (static owner = the warp holding the row's first segment) finishes it in fixed segment order:
//            s_i = D_i z_i + sum parts, stores s_i and accumulates gamma += r_i.z_i, delta += z_i.s_i.   -> barrier 1
//   phase C  gamma, delta from the block slots (every block adds them in the same order: bitwise equal
//            everywhere);  beta = gamma/gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old);
//            p = z + beta p, q = s + beta q, x += alpha p, r -= alpha q, z = Minv r, slots of r.r          -> barrier 2
// No epilogue pass: the model decrease needs x.H x = x.(b - r) - x.Lam x, all per-view quantities (k_apply_step).
// Work distribution: the half-edge array is cut into num_warps equal contiguous ranges (one per resident warp,
// grid = SMs x occupancy), each range into segments (range ^ row).
// Vectors written inside the kernel are never accessed through __restrict__/read-only paths.
// ------------------------------------------------------------------------------------------
// After the solve (xt = tangent step): parameter step delta = D^-1 xt, candidate = omega + delta (Ceres updates the
// angle-axis vector additively) or, on the manifold, R <- R Exp(xt); reduces delta.g (= xt.gt),
// delta.H.delta = xt.Ht.xt and |delta|^2.  The quadratic form needs no matrix pass: (Ht + Lam) xt = b - r with the
// solver's residual r, so xt.Ht.xt = xt.(b - r) - xt.Lam.xt, Lam_i = Dblk_i - Hd_i -- all per-view quantities.
struct ApplyArgs {
  const double *node_JL, *xt, *bvec, *res, *Dblk, *Hd, *gt, *omega;
  double *cand, *delta_out;
  int manifold;
};
// per-view body of k_apply_step: v += (delta.g, delta.H.delta, |delta|^2, non-finite flag); the candidate is also returned in w3
__device__ __forceinline__ void apply_view(uint32_t i, const ApplyArgs& A, double (&v)[4], double* w3) {
  // every load first (see prepare_view)
  double J[9], om[3] = {0.0, 0.0, 0.0}, g3[3] = {0.0, 0.0, 0.0}, lam[6] = {0, 0, 0, 0, 0, 0}, br[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 9; ++k) J[k] = A.node_JL[9 * (size_t)i + k];
  const double t0 = A.xt[3 * (size_t)i], t1 = A.xt[3 * (size_t)i + 1], t2 = A.xt[3 * (size_t)i + 2];
  if (A.omega) { om[0] = A.omega[3 * (size_t)i]; om[1] = A.omega[3 * (size_t)i + 1]; om[2] = A.omega[3 * (size_t)i + 2]; }
  if (A.gt) { g3[0] = A.gt[3 * (size_t)i]; g3[1] = A.gt[3 * (size_t)i + 1]; g3[2] = A.gt[3 * (size_t)i + 2]; }
  if (A.bvec) {
#pragma unroll
    for (int k = 0; k < 6; ++k) lam[k] = A.Dblk[6 * (size_t)i + k] - A.Hd[6 * (size_t)i + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) br[c] = A.bvec[3 * (size_t)i + c] - (A.res ? A.res[3 * (size_t)i + c] : 0.0);
  }
  double Ji[9];
  inv3(J, Ji);
  double d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c] = Ji[3 * c] * t0 + Ji[3 * c + 1] * t1 + Ji[3 * c + 2] * t2;
  w3[0] = w3[1] = w3[2] = 0.0;
  if (A.manifold == 1) {
    // x (+) delta = [sin|d| d/|d|, cos|d|] (x) x, a left rotation by 2 delta = R xt, i.e. R <- R Exp(xt); the state stays
    // an angle-axis vector (principal branch of the product quaternion).  |step| in the ambient quaternion space =
    // 2 sin(|delta| / 2) per view, |delta| = |xt| / 2.
    const Q4 qd = aa_to_quat(t0, t1, t2);
    const Q4 qo = aa_to_quat(om[0], om[1], om[2]);
    double th2, cc;
    quat_log(qmul(qo, qd), w3, &th2, &cc);
    const double dn = 0.5 * sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    const double sh = sin(0.5 * dn);
    v[2] += 4.0 * sh * sh;
    if (!isfinite(dn)) v[3] = 1.0;
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      w3[c] = om[c] + d[c];
      v[2] += d[c] * d[c];
      if (!isfinite(d[c])) v[3] = 1.0;
    }
  }
  if (A.gt) v[0] += t0 * g3[0] + t1 * g3[1] + t2 * g3[2];
  if (A.bvec) {
    const double t[3] = {t0, t1, t2};
    double lx[3];
    sym_mul_vec(lam, t, lx);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[1] += t[c] * (br[c] - lx[c]);
  }
  if (A.delta_out) { A.delta_out[3 * (size_t)i] = d[0]; A.delta_out[3 * (size_t)i + 1] = d[1]; A.delta_out[3 * (size_t)i + 2] = d[2]; }
  if (A.cand) { A.cand[3 * (size_t)i] = w3[0]; A.cand[3 * (size_t)i + 1] = w3[1]; A.cand[3 * (size_t)i + 2] = w3[2]; }
}

__global__ void k_apply_step(uint32_t N, ApplyArgs A, double* slots, unsigned* counter, DevScalars* sc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  double w3[3];
  if (i < N) apply_view(i, A, v, w3);
  double tot[4];
  if (grid_sum<4>(v, slots, counter, tot) && threadIdx.x == 0) {
    sc->dg = tot[0]; sc->dHd = tot[1]; sc->step2 = tot[2];
    sc->t_linear_end = gtimer_ns();
    (void)tot[3];  // a non-finite step makes step2 non-finite: the host treats it as an invalid step
  }
}

struct PcgParams {
  uint32_t N, num_warps, n_iso, warp_span;
  uint32_t keep8;  // records of every 8 loaded with the evict_last L2 policy (0: no cache hints)
  ColBlocks cbk;         // column blocks of the half-edge order; cbk.ncb = 0: no shared-memory staging of z
  int max_iter;
  uint64_t H;
  double rtol2;
  const uint32_t *warp_seg_ptr, *seg_row, *seg_begin, *seg_len, *node_seg_ptr, *iso;
  const double *val, *Dblk, *Minv;
  double *x, *r, *z, *p, *q, *s, *ypart;
  unsigned* row_cnt;
  unsigned long long* bar_slots;  // [2][grid][4]: grid barrier + reduction (grid_bar_sum2)
  double* slots;                  // grid_sum scratch (epilogue)
  unsigned* counter;
  DevScalars* sc;
  unsigned long long* prof;  // optional [8] phase timers in ns, accumulated by block 0 (measurement aid)
  // fused trust-region step: prologue = k_prepare_solve's per-view work, epilogue = k_apply_step + k_node_prep of the candidate
  int fused;
  const IterParams* ip;
  PrepareArgs prep;
  ApplyArgs apply;
  double *cand_q, *cand_JL;
  // edge-sharded multi-GPU (world > 1): every rank's exchange block (ra_common.cuh, LLCell), mapped into this process over
  // NVLink (CUDA IPC between processes, peer access inside one).  peer[rank] is this rank's own block.
  int world, rank;
  int owner_mode;  // exchange mode (ra_common.cuh): 0 direct, 1 owner-reduce
  LLCell* peer[kMaxPeers];
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Grid barrier + deterministic reduction of two doubles (the persistent PCG kernel): block sums go to per-block slots,
// cooperative groups' grid.sync, then every block adds the slots in the same order.  Two slot sets alternate by barrier
// parity (a block can reach barrier n+2 only after every block arrived at n+1, i.e. finished reading n).
// (Measured alternative, profiles/r01_h: publishing {data | sequence} words and polling all blocks' slots instead of
// grid.sync is slower -- 296 pollers x 296 slots of dependent L2 round trips.)
__device__ __forceinline__ void grid_bar_sum2(cg::grid_group& grid, unsigned long long* bar_slots, unsigned& seq, double v0, double v1,
                                              double* sm_red /*[2*kMaxWarpsPerBlock + 2]*/, double& out0, double& out1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v0 = warp_sum(v0); v1 = warp_sum(v1);
  if (lane == 0) { sm_red[2 * warp] = v0; sm_red[2 * warp + 1] = v1; }
  __syncthreads();
  ++seq;
  double* set = reinterpret_cast<double*>(bar_slots) + (size_t)(seq & 1u) * 2 * gridDim.x;
  if (threadIdx.x == 0) {
    double b0 = 0.0, b1 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { b0 += sm_red[2 * w]; b1 += sm_red[2 * w + 1]; }
    __stcg(set + 2 * (size_t)blockIdx.x, b0); __stcg(set + 2 * (size_t)blockIdx.x + 1, b1);
  }
  grid.sync();
  if (warp == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) { s0 += __ldcg(set + 2 * (size_t)b); s1 += __ldcg(set + 2 * (size_t)b + 1); }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if (lane == 0) { sm_red[2 * kMaxWarpsPerBlock] = s0; sm_red[2 * kMaxWarpsPerBlock + 1] = s1; }
  }
  __syncthreads();
  out0 = sm_red[2 * kMaxWarpsPerBlock]; out1 = sm_red[2 * kMaxWarpsPerBlock + 1];
}

// Row i once its off-diagonal sum (y0,y1,y2) is complete: s_i = D_i z_i + y, store, inner products.
__device__ __forceinline__ void finish_row(const PcgParams& P, uint32_t row, double y0, double y1, double y2, double& gamma, double& delta) {
  const double4 zv = reinterpret_cast<const double4*>(P.z)[row];
  const double zi[3] = {zv.x, zv.y, zv.z};
  double d[3];
  sym_mul_vec(P.Dblk + 6 * (size_t)row, zi, d);
  y0 += d[0]; y1 += d[1]; y2 += d[2];
  P.s[3 * (size_t)row] = y0; P.s[3 * (size_t)row + 1] = y1; P.s[3 * (size_t)row + 2] = y2;
  gamma += P.r[3 * (size_t)row] * zi[0] + P.r[3 * (size_t)row + 1] * zi[1] + P.r[3 * (size_t)row + 2] * zi[2];
  delta += zi[0] * y0 + zi[1] * y1 + zi[2] * y2;
}

// Put the first chunks of this warp's range in flight (before the prologue's barrier: they land under it).
template <int kBlk>
__device__ __forceinline__ void spmv_prime(const PcgParams& P, WarpPipe& wp) {
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (!kWrapPrefetch || gwarp >= P.num_warps) return;
  const uint64_t lo = (uint64_t)gwarp * P.warp_span, hi = min(P.H, lo + P.warp_span);
  spmv_stream<kBlk>(wp, P.val, lo, hi, 0u, 0u, P.seg_begin, P.seg_len, P.z, nullptr, 0u, false, P.keep8, [](uint32_t, double) {}, false, true);
}

// One SpMV pass over this warp's range, s = (Ht + Lam) z.  Accumulates (per lane) gamma = r.z and delta = z.s over
// the rows this lane finished.
template <int kBlk>
__device__ __forceinline__ void spmv_pass(const PcgParams& P, WarpPipe& wp, double& gamma, double& delta, const double* zs, uint32_t zbase,
                                          unsigned xseq = 0u) {
  // xseq != 0 (multi-GPU, exchange step xseq): only the shard-local off-diagonal row sums are produced; the row owner PUSHES
  // them, tagged with xseq, into the exchange block of every rank (its own included) while the pass is still running;
  // diagonal and inner products follow once the W contributions of a row have arrived (exchange_finish).
  const bool push = xseq != 0u;
  const int lane = threadIdx.x & 31;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gwarp < P.num_warps) {
    const uint64_t lo = (uint64_t)gwarp * P.warp_span, hi = min(P.H, lo + P.warp_span);
    // Rows are finished by a STATIC owner -- the warp holding the row's first segment (first non-empty column block, first
    // range) -- so every sum has a fixed order and a fixed place (bit-reproducible).  Every segment's sum goes to ypart[t];
    // a segment that is not its row's first one also counts itself in on the row's counter.  Only when its whole range is
    // streamed does a warp turn to the rows it owns: it waits for the row's other segments, adds all of them in segment
    // order and finishes the row.  Nothing is waited for before everything a warp owes others is published, and the
    // cooperative launch keeps every block resident, so the waits cannot deadlock.
    const uint32_t t0 = P.warp_seg_ptr[gwarp], t1 = P.warp_seg_ptr[gwarp + 1];
    const uint32_t ncb = P.cbk.ncb ? P.cbk.ncb : 1u;
    spmv_stream<kBlk>(wp, P.val, lo, hi, t0, t1, P.seg_begin, P.seg_len, P.z, zs, zbase, false, P.keep8, [&](uint32_t t, double ysum) {
      if ((lane & 7) == 0 && lane < 24) __stcg(P.ypart + 3 * (size_t)t + multi_index4(lane), ysum);
      if (ncb == 1u) {
        // one column block: the only segment of a range that is not its row's first is the range's FIRST segment (the row began
        // in an earlier range) -- publish it at once, its owner is about to need it
        const uint32_t rowf = P.seg_row[t];
        if (rowf & kSideBit) {
          __threadfence();
          __syncwarp();
          if (lane == 0) { __threadfence(); atomicAdd(P.row_cnt + (rowf & ~kSideBit), 1u); }
        }
      }
    }, kWrapPrefetch);
    if (ncb > 1u) {
      // several column blocks: a range holds many such segments; ONE fence for the whole range, then every one of them counts
      // itself in on its row's counter (lanes in parallel) -- a fence per segment would stall the stream every ~50 half-edges
      __threadfence();
      __syncwarp();
      __threadfence();
      for (uint32_t t = t0 + (uint32_t)lane; t < t1; t += 32u) {
        const uint32_t rowf = P.seg_row[t];
        if (rowf & kSideBit) atomicAdd(P.row_cnt + (rowf & ~kSideBit), 1u);
      }
    }
    __syncwarp();
    for (uint32_t t = t0 + (uint32_t)lane; t < t1; t += 32u) {
      const uint32_t rowf = P.seg_row[t];
      if (rowf & kSideBit) continue;
      const uint32_t row = rowf;
      uint32_t nseg = 0;
      for (uint32_t cb = 0; cb < ncb; ++cb) nseg += P.node_seg_ptr[cb * P.N + row + 1] - P.node_seg_ptr[cb * P.N + row];
      if (nseg > 1) {
        volatile unsigned* cnt = P.row_cnt + row;
        while (*cnt != nseg - 1) { }
        __threadfence();
        *cnt = 0u;
      }
      double my0 = 0.0, my1 = 0.0, my2 = 0.0;
      for (uint32_t cb = 0; cb < ncb; ++cb)
        for (uint32_t k = P.node_seg_ptr[cb * P.N + row]; k < P.node_seg_ptr[cb * P.N + row + 1]; ++k) {
          my0 += __ldcg(P.ypart + 3 * (size_t)k); my1 += __ldcg(P.ypart + 3 * (size_t)k + 1); my2 += __ldcg(P.ypart + 3 * (size_t)k + 2);
        }
      if (push) {
        const int r0 = P.owner_mode ? (int)(row % (uint32_t)P.world) : 0, r1 = P.owner_mode ? r0 + 1 : P.world;
        for (int r = r0; r < r1; ++r) {
          LLCell* dst = P.peer[r] + ll_cg_offset(P.N, P.world, xseq, P.rank) + 3 * (size_t)row;
          ll_store(dst, my0, xseq); ll_store(dst + 1, my1, xseq); ll_store(dst + 2, my2, xseq);
        }
      } else {
        finish_row(P, row, my0, my1, my2, gamma, delta);
      }
    }
  }
  // views without any half-edge (in this shard): s_i = D_i z_i; multi-GPU: a tagged zero contribution
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n_iso; k += gridDim.x * blockDim.x) {
    if (push) {
      const uint32_t row = P.iso[k];
      const int r0 = P.owner_mode ? (int)(row % (uint32_t)P.world) : 0, r1 = P.owner_mode ? r0 + 1 : P.world;
      for (int r = r0; r < r1; ++r) {
        LLCell* dst = P.peer[r] + ll_cg_offset(P.N, P.world, xseq, P.rank) + 3 * (size_t)row;
        ll_store(dst, 0.0, xseq); ll_store(dst + 1, 0.0, xseq); ll_store(dst + 2, 0.0, xseq);
      }
    } else {
      finish_row(P, P.iso[k], 0.0, 0.0, 0.0, gamma, delta);
    }
  }
}

// Fused cross-GPU reduction of the partial matvec, inside the persistent kernel: no NCCL call, no kernel boundary, no grid
// barrier and no fence.  Every rank pushed its row sums as tagged cells into every rank's exchange block during the pass
// (spmv_pass); here the views are dealt to groups of 4 lanes, lane q of a group waits for the cells of ranks q, q+4, ... in its
// OWN block (local memory: the peers' stores arrive over NVLink), the group adds them in a fixed order -- bitwise identical on
// every rank -- and finishes the row: s_i = D_i z_i + sum, inner products.  Two buffers (step parity) suffice: a rank reaches
// step seq+2 only after it finished step seq+1, which needs every peer's seq+1 cells, which a peer sends only after it finished
// reading step seq.
__device__ __forceinline__ void exchange_finish(const PcgParams& P, unsigned seq, double& gamma, double& delta) {
  constexpr int G = 4;
  const int lane = threadIdx.x & 31, sub = lane & (G - 1);
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const LLCell* mine = P.peer[P.rank];
  int bad = 0;
  const uint32_t W = (uint32_t)P.world;
  // owner mode: only the views of this rank (i % W == rank) are reduced here, then pushed to everybody; all views follow below
  const uint32_t nloop = P.owner_mode ? (P.N + W - 1) / W : P.N;
  for (uint32_t vb = gw * (32 / G); vb < nloop; vb += nwarps * (32 / G)) {
    const uint32_t k = vb + (uint32_t)(lane / G);
    const uint32_t i = P.owner_mode ? (uint32_t)P.rank + W * k : k;
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    if (i < P.N) {
      for (int r = sub; r < P.world; r += G) {
        const LLCell* c = mine + ll_cg_offset(P.N, P.world, seq, r) + 3 * (size_t)i;
        double a0, a1, a2;
        long long t0 = 0;
        while (true) {  // the three loads of a poll are issued together
          const bool ok0 = ll_try_load(c, seq, a0), ok1 = ll_try_load(c + 1, seq, a1), ok2 = ll_try_load(c + 2, seq, a2);
          if (ok0 && ok1 && ok2) break;
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) { bad = 2; a0 = a1 = a2 = 0.0; break; }  // ~4 s: a peer died; do not hang the GPU
        }
        y0 += a0; y1 += a1; y2 += a2;
      }
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      y0 += __shfl_xor_sync(0xffffffffu, y0, o); y1 += __shfl_xor_sync(0xffffffffu, y1, o); y2 += __shfl_xor_sync(0xffffffffu, y2, o);
    }
    if (P.owner_mode) {
      if (i < P.N)
        for (int r = sub; r < P.world; r += G) {  // the group's lanes share the W destinations
          LLCell* dst = P.peer[r] + ll_cgred_offset(P.N, P.world, seq) + 3 * (size_t)i;
          ll_store(dst, y0, seq); ll_store(dst + 1, y1, seq); ll_store(dst + 2, y2, seq);
        }
    } else if (i < P.N && sub == 0) {
      finish_row(P, i, y0, y1, y2, gamma, delta);
    }
  }
  if (P.owner_mode) {
    // every view: its reduced sums arrive from its owner
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += gridDim.x * blockDim.x) {
      double a[3];
      ll_wait_n<3>(mine + ll_cgred_offset(P.N, P.world, seq) + 3 * (size_t)i, seq, a, &bad);
      finish_row(P, i, a[0], a[1], a[2], gamma, delta);
    }
  }
  if (bad) P.sc->bad = bad;
}

template <int kBlk>
__global__ void __launch_bounds__(kPcgBlock, kPcgBlocksPerSM) k_pcg_persistent(PcgParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::grid_group grid = cg::this_grid();
  WarpPipe wp;
  pipe_init<kBlk>(wp, smem_raw);
  __shared__ double sm_red[2 * kMaxWarpsPerBlock + 2];
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  unsigned bseq = (unsigned)P.sc->bar_seq;  // barrier sequence number, continues across launches
  double bb;
  bool done;
  spmv_prime<kBlk>(P, wp);
  if (P.fused) {
    // ---- prologue (k_prepare_solve): LM damping in the tangent frame, block-Jacobi inverse, x = 0, r = b, z = p = M^-1 b
    if (gtid == 0) { P.sc->t_begin = gtimer_ns(); P.sc->bad = 0; }
    PrepareArgs A = P.prep;
    if (P.ip) A.mu = P.ip->mu;
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x + gridDim.x * threadIdx.x; i < P.N; i += gthreads) prepare_view(i, A, v);
    double rz0;
    grid_bar_sum2(grid, P.bar_slots, bseq, v[0], v[1], sm_red, rz0, bb);
    done = (bb == 0.0 || !isfinite(bb));
    if (gtid == 0) { P.sc->bb = bb; P.sc->rz = rz0; }
  } else {
    bb = P.sc->bb;
    done = P.sc->pcg_done != 0;
  }
  double rr = bb, beta = 0.0, alpha = 0.0, gamma_old = 0.0;
  int iter = 0, breakdown = 0;
  const bool multi = P.world > 1;
  unsigned seq = multi ? (unsigned)P.sc->xseq : 0u;  // exchange sequence number, continues across launches
  while (!done) {
    // ---- phase A: s = (Ht + Lam) z, gamma = r.z, delta = z.s ---------------------------------
    const bool prof = P.prof != nullptr && gtid == 0;
    unsigned long long tA = 0, tB = 0, tC = 0, tE = 0, tF = 0;
    if (prof) tA = gtimer();
    double g_part = 0.0, d_part = 0.0;
    const double* zs = nullptr;
    uint32_t zbase = 0;
    if (P.cbk.ncb) {
      // stage the slice of z this CTA's half-edges gather from (z is final since the barrier that closed the previous
      // phase): possible when the CTA's whole range lies inside ONE column block -- all but at most ncb - 1 CTAs
      const uint64_t wpb = blockDim.x >> 5;
      const uint64_t lo_cta = (uint64_t)blockIdx.x * wpb * P.warp_span, hi_cta = min(P.H, lo_cta + wpb * P.warp_span);
      int cb = -1;
      if (lo_cta < P.H)
        for (uint32_t c = 0; c < P.cbk.ncb; ++c)
          if ((uint64_t)P.cbk.begin[c] <= lo_cta && hi_cta <= (uint64_t)P.cbk.begin[c + 1]) cb = (int)c;
      if (cb >= 0) {
        constexpr int kSliceOffset = kPcgWarps * kStages2 * Chunk<kBlk>::kBytes + kPcgWarps * kStages2 * 8;  // == spmv_smem_bytes(kBlk)
        double* sl = reinterpret_cast<double*>(smem_raw + kSliceOffset);
        zbase = (uint32_t)cb * P.cbk.cbsize;
        const uint32_t nv = min(P.cbk.cbsize, P.N - zbase);
        for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) {
          const double4 zv = reinterpret_cast<const double4*>(P.z)[zbase + i];
          sl[3 * i] = zv.x; sl[3 * i + 1] = zv.y; sl[3 * i + 2] = zv.z;
        }
        __syncthreads();
        zs = sl;
      }
    }
    if (!multi) spmv_pass<kBlk>(P, wp, g_part, d_part, zs, zbase);
    else {
      if (++seq == 0u) ++seq;  // 0 means "no exchange" in spmv_pass
      spmv_pass<kBlk>(P, wp, g_part, d_part, zs, zbase, seq);
      exchange_finish(P, seq, g_part, d_part);
    }
    if (prof) tB = gtimer();
    double gamma, delta;
    grid_bar_sum2(grid, P.bar_slots, bseq, g_part, d_part, sm_red, gamma, delta);
    if (prof) tC = gtimer();
    // p = z + beta p  =>  p.Ap = delta - beta^2 (p_old.A p_old) = delta - beta gamma / alpha_old
    beta = (iter == 0) ? 0.0 : gamma / gamma_old;
    const double pAp = (iter == 0) ? delta : delta - beta * gamma / alpha;
    if (!(pAp > 0.0) || !isfinite(pAp)) { breakdown = 1; break; }
    alpha = gamma / pAp;
    gamma_old = gamma;
    // ---- phase C: p, q, x, r, z ; r.r ----------------------------------------------------------
    // views are dealt round-robin to the blocks (view i -> block i % grid) so every SM carries a few; all loads of a
    // view are issued before its first store (the vectors may alias as far as the compiler knows: a store between two
    // loads would serialise the L2 round trips)
    double v1 = 0.0;
    for (uint32_t i = blockIdx.x + gridDim.x * threadIdx.x; i < P.N; i += gthreads) {
      const double4 zv = reinterpret_cast<const double4*>(P.z)[i];
      const double4 pv = reinterpret_cast<const double4*>(P.p)[i];
      double sv[3], qv[3], xv[3], rv[3], Mi[6];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sv[c] = P.s[3 * (size_t)i + c]; qv[c] = P.q[3 * (size_t)i + c]; xv[c] = P.x[3 * (size_t)i + c]; rv[c] = P.r[3 * (size_t)i + c];
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) Mi[c] = P.Minv[6 * (size_t)i + c];
      const double zi[3] = {zv.x, zv.y, zv.z};
      const double po[3] = {pv.x, pv.y, pv.z};
      double ri[3], zn[3], pn[3], qn[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        pn[c] = zi[c] + beta * po[c];
        qn[c] = sv[c] + beta * qv[c];
        xv[c] += alpha * pn[c];
        ri[c] = rv[c] - alpha * qn[c];
        v1 += ri[c] * ri[c];
      }
      sym_mul_vec(Mi, ri, zn);
#pragma unroll
      for (int c = 0; c < 3; ++c) { P.q[3 * (size_t)i + c] = qn[c]; P.x[3 * (size_t)i + c] = xv[c]; P.r[3 * (size_t)i + c] = ri[c]; }
      reinterpret_cast<double4*>(P.p)[i] = make_double4(pn[0], pn[1], pn[2], 0.0);
      reinterpret_cast<double4*>(P.z)[i] = make_double4(zn[0], zn[1], zn[2], 0.0);
    }
    if (prof) tE = gtimer();
    double unused;
    grid_bar_sum2(grid, P.bar_slots, bseq, v1, 0.0, sm_red, rr, unused);
    if (prof) {
      tF = gtimer();
      P.prof[0] += tB - tA; P.prof[1] += tC - tB; P.prof[3] += tE - tC; P.prof[4] += tF - tE;
      P.prof[6] += 1;
    }
    ++iter;
    if (rr <= P.rtol2 * bb || iter >= P.max_iter || !isfinite(rr)) done = true;
  }
  spmv_drain(wp);
  if (gtid == 0) {
    P.sc->xseq = (int)seq;
    P.sc->bar_seq = (int)bseq;
    P.sc->rz = gamma_old; P.sc->rr = rr; P.sc->beta = beta; P.sc->alpha = alpha;
    P.sc->pcg_iter = iter; P.sc->pcg_done = 1; P.sc->pcg_breakdown = breakdown;
  }
  if (P.fused) {
    // ---- epilogue (k_apply_step + k_node_prep of the candidate): x is complete and visible (the loop ends on a barrier;
    // a breakdown leaves the previous, barrier-covered x) ------------------------------------------
    double v[4] = {0.0, 0.0, 0.0, 0.0}, xn = 0.0;
    for (uint32_t i = blockIdx.x + gridDim.x * threadIdx.x; i < P.N; i += gthreads) {
      double w3[3];
      apply_view(i, P.apply, v, w3);
      xn += node_prep_view(i, w3, P.cand_q, P.cand_JL, P.apply.manifold);
    }
    double v5[5] = {v[0], v[1], v[2], v[3], xn}, tot[5];
    if (grid_sum<5>(v5, P.slots, P.counter, tot) && threadIdx.x == 0) {
      P.sc->dg = tot[0]; P.sc->dHd = tot[1]; P.sc->step2 = tot[2]; P.sc->xnorm2 = tot[4];
      P.sc->gmax = 0.0;  // accumulated by the k_node_finalize of the candidate's evaluation
      P.sc->t_linear_end = gtimer_ns();
    }
  }
}

}  // namespace