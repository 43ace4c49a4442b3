// ra_edges.cuh -- K1: fused residual / SO(3) Jacobian / whitening / robust loss / normal-equation assembly; per-view kernels
// Part of libgsfm_ra (one translation unit, see gsfm_ra.cu); reference citations sit next to each kernel.
#pragma once
#include "ra_common.cuh"
namespace {


// Per view: quaternion + the factor D of d(beta) = D d(parameters) at the current estimate; also |x|^2.
//   angle-axis parameters:  D = Jr(omega) = Jl(-omega)   (R(omega + d omega) = R(omega) Exp(Jr d omega))
//   quaternion parameters with EigenQuaternionParameterization (x (+) delta = [sin|d| d/|d|, cos|d|] (x) x, i.e. a LEFT
//   perturbation phi = 2 delta):  beta = R^T phi  ->  D = 2 R^T;  |x|^2 = 1 per unit quaternion
// The array keeps its historical name node_JL.
// per-view body of k_node_prep; returns the view's contribution to |x|^2
__device__ __forceinline__ double node_prep_view(uint32_t i, const double* w3, double* __restrict__ node_q, double* __restrict__ node_JL, int manifold) {
  const double wx = w3[0], wy = w3[1], wz = w3[2];
  if (manifold == 2) {
    // translation averaging: the parameters are camera positions (Euclidean, D = I); the position rides in the x,y,z slots
    // of the 32 B per-view sector the edge kernel gathers
    reinterpret_cast<double4*>(node_q)[i] = make_double4(0.0, wx, wy, wz);
#pragma unroll
    for (int t = 0; t < 9; ++t) node_JL[9 * (size_t)i + t] = (t == 0 || t == 4 || t == 8) ? 1.0 : 0.0;
    return wx * wx + wy * wy + wz * wz;
  }
  const Q4 q = aa_to_quat(wx, wy, wz);
  reinterpret_cast<double4*>(node_q)[i] = make_double4(q.w, q.x, q.y, q.z);
  double J[9];
  double xn;
  if (manifold == 1) {
    double R[9];
    quat_to_mat(q, R);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) J[3 * r + c] = 2.0 * R[3 * c + r];
    xn = 1.0;
  } else {
    so3_left_jacobian(-wx, -wy, -wz, J);
    xn = wx * wx + wy * wy + wz * wz;
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) node_JL[9 * (size_t)i + t] = J[t];
  return xn;
}

__global__ void k_node_prep(uint32_t N, const double* __restrict__ omega, double* __restrict__ node_q, double* __restrict__ node_JL,
                            double* slots, unsigned* counter, DevScalars* sc, int manifold) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[1] = {0.0};
  if (i == 0) sc->gmax = 0.0;  // max |g| of this evaluation is accumulated (atomicMax) by the k_node_finalize that follows
  if (i < N) {
    const double w3[3] = {omega[3 * (size_t)i], omega[3 * (size_t)i + 1], omega[3 * (size_t)i + 2]};
    v[0] = node_prep_view(i, w3, node_q, node_JL, manifold);
  }
  double tot[1];
  if (grid_sum<1>(v, slots, counter, tot) && threadIdx.x == 0) sc->xnorm2 = tot[0];
}

// ------------------------------------------------------------------------------------------
// K1: fused residual + SO(3) Jacobian + whitening + robust loss + normal-equation assembly.
// One warp per balanced range of half-edges.  The per-half-edge constants arrive as INPUT RECORDS through the TMA
// ring (see k_setup_halfedges): lane l owns half-edge l of the record, reads its q_ij / U / col / row from shared
// memory, gathers the two endpoint quaternions (one aligned 32 B sector each, L2; issued one record ahead so the
// latency overlaps the arithmetic of the current record), evaluates the edge and writes the off-diagonal block -S
// (6 doubles, planar in the OUTPUT record, coalesced).  Per segment (range ^ row) the warp reduces the diagonal block /
// gradient / cost partial.  kWriteBlocks=false is K1c: cost only (trial point).
// ------------------------------------------------------------------------------------------
struct K1Args {
  uint32_t num_warps, warp_span;
  uint64_t H;
  const uint32_t *warp_seg_ptr, *seg_begin, *seg_len;
  const double *inrec, *node_q;
  double *val, *part;
  DevLoss loss;
  uint32_t fixed;  // translation averaging: the view held constant (kNoFixedView: none)
};
constexpr uint32_t kNoFixedView = 0xffffffffu;

template <bool kWriteBlocks, int kResidual, bool kScalarU, int kLoss>
// two blocks (16 warps) per SM: three (<= 80 registers) spill and measure 10 % slower (profiles/r01_g_microbench.txt item 5)
__global__ void __launch_bounds__(kBlock, 2) k_edges(const K1Args A) {
  constexpr int kU = (kScalarU || kResidual == 1) ? 1 : 6;
  constexpr bool kCompact = kCompactScalarStencil && kScalarU && kResidual == 0;  // 4-double block records (see Rec<4>)
  constexpr int kRD = (4 + kU) * 32 + 32;  // doubles per input record
  constexpr int kRB = kRD * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= A.num_warps) return;
  WarpPipe wp;
  pipe_init_bytes<kRB, kStages>(wp, smem_raw);
  const uint64_t lo = (uint64_t)gw * A.warp_span, hi = min(A.H, lo + A.warp_span);
  const uint32_t t0 = A.warp_seg_ptr[gw], t1 = A.warp_seg_ptr[gw + 1];
  const uint32_t nrec = (uint32_t)((hi - lo + 31) >> 5);
  const double* src = A.inrec + (size_t)(lo >> 5) * kRD;
  // The input records are read ONCE per evaluation (96 MB at 1M edges): loaded with the evict_first L2 policy they do not push
  // out the block records this kernel writes (72 MB), which the persistent PCG kernel streams right afterwards.
  // (-DGSFM_RA_K1_NO_HINT: plain loads, A/B builds.)
#ifndef GSFM_RA_K1_NO_HINT
  const uint64_t pol_once = l2_policy_evict_first();
#endif
  auto issue = [&](uint32_t c) {
    if (lane == 0) {
      const uint32_t st = c % kStages;
      mbar_expect_tx(&wp.bars[st], kRB);
#ifndef GSFM_RA_K1_NO_HINT
      tma_load_bulk_hint(wp.ring + (size_t)st * kRD, src + (size_t)c * kRD, kRB, &wp.bars[st], pol_once);
#else
      tma_load_bulk(wp.ring + (size_t)st * kRD, src + (size_t)c * kRD, kRB, &wp.bars[st]);
#endif
    }
  };
  auto wait_rec = [&](uint32_t c) -> const double* {
    const uint32_t st = c % kStages;
    mbar_wait(&wp.bars[st], (c / kStages) & 1u);
    return wp.ring + (size_t)st * kRD;
  };
  // the two endpoint sectors of a half-edge, already in EDGE order (view i, view j): the side is resolved on the two indices,
  // not by selecting between eight doubles afterwards
  auto gather = [&](const double* rec, uint64_t h, uint32_t& cf, Q4& qi, Q4& qj) {
    const uint32_t* idx = reinterpret_cast<const uint32_t*>(rec + (4 + kU) * 32);
    cf = idx[lane];
    uint32_t row = idx[32 + lane];
    if (h >= hi) { cf = 0; row = 0; }  // padding lanes of the last record
    const uint32_t col = cf & ~kSideBit;
    const bool rj = (cf & kSideBit) != 0;
    const double4 a = reinterpret_cast<const double4*>(A.node_q)[rj ? col : row];
    const double4 b = reinterpret_cast<const double4*>(A.node_q)[rj ? row : col];
    qi = Q4{a.x, a.y, a.z, a.w};
    qj = Q4{b.x, b.y, b.z, b.w};
  };
  for (uint32_t c = 0; c < nrec && c < (uint32_t)kStages; ++c) issue(c);
  if (nrec == 0 || t0 == t1) return;
  uint32_t t = t0;
  uint64_t sb = A.seg_begin[t], se = sb + A.seg_len[t];
  constexpr int kAcc = kWriteBlocks ? kPartStride : 1;
  double acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
  const double* rec = wait_rec(0);
  uint32_t cf;
  Q4 qi, qj;
  gather(rec, lo + lane, cf, qi, qj);
  for (uint32_t c = 0; c < nrec; ++c) {
    const uint64_t cb = lo + ((uint64_t)c << 5), ce = cb + 32, h = cb + lane;
    const double* rec_n = nullptr;
    uint32_t cf_n = 0;
    Q4 qi_n{1, 0, 0, 0}, qj_n{1, 0, 0, 0};
    if (c + 1 < nrec) { rec_n = wait_rec(c + 1); gather(rec_n, ce + lane, cf_n, qi_n, qj_n); }
    const bool row_is_j = (cf & kSideBit) != 0;
    const Q4 qm{rec[lane], rec[32 + lane], rec[64 + lane], rec[96 + lane]};
    // translation averaging holds ONE view constant (position_estimator.cpp:121-122 SetParameterBlockConstant): its gradient is
    // dropped and every off-diagonal block that touches it is stored as zero, so the step of that view is exactly zero and
    // the other views see the reduced system Ceres solves (their diagonal blocks keep the edge's contribution)
    bool fix_row = false, fix_any = false;
    if (kResidual == 2 && A.fixed != kNoFixedView) {
      const uint32_t rowv = reinterpret_cast<const uint32_t*>(rec + (4 + kU) * 32)[32 + lane];
      fix_row = rowv == A.fixed;
      fix_any = fix_row || (cf & ~kSideBit) == A.fixed;
    }
    double u[6];
#pragma unroll
    for (int k = 0; k < kU; ++k) u[k] = rec[(4 + k) * 32 + lane];
    // this record's slot can be refilled as soon as every lane has read it
    __syncwarp();
    if (c + kStages < nrec) issue(c + kStages);
    EdgeTerms et;
    edge_terms<kWriteBlocks, kResidual, kScalarU, kLoss, true>(qi, qj, qm, u, A.loss, et);  // (q_i, q_j): see gather
    double cur[kAcc];
    if (kWriteBlocks) {
      // both rows of the edge: diag += S, block(row, col) = -S; gradient: +v in row j, -v in row i; cost once, in row i
      const double sgn = row_is_j ? 1.0 : -1.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) cur[k] = et.S[k];
      cur[6] = sgn * et.v[0]; cur[7] = sgn * et.v[1]; cur[8] = sgn * et.v[2];
      cur[kAcc - 1] = row_is_j ? 0.0 : 0.5 * et.rho[0];
      if (kResidual == 2) {
        if (fix_row) cur[6] = cur[7] = cur[8] = 0.0;
        if (fix_any) {
#pragma unroll
          for (int k = 0; k < 6; ++k) et.S[k] = 0.0;   // cur[] already holds the diagonal contribution
        }
      }
      if (h < hi) {
        // lo is a multiple of 32: record (lo >> 5) + c, word `lane` of each component row
        if (kCompact) {
          double* out = A.val + ((size_t)(lo >> 5) + c) * Rec<4>::kDoubles + lane;
          out[0] = et.ca;
#pragma unroll
          for (int k = 0; k < 3; ++k) out[(1 + k) * 32] = et.h[k];
        } else {
          double* out = A.val + ((size_t)(lo >> 5) + c) * Rec<6>::kDoubles + lane;
#pragma unroll
          for (int k = 0; k < 6; ++k) out[k * 32] = -et.S[k];
        }
      }
    } else {
      cur[0] = row_is_j ? 0.0 : 0.5 * et.rho[0];
    }
    while (true) {
      if (h >= sb && h < se) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) acc[k] += cur[k];
      }
      if (se > ce) break;  // the segment continues in the next record
      if (kWriteBlocks) {  // the ten sums of the segment in one reduce-scatter butterfly; lane 2m ends up with value multi_index16(2m)
        const double tot = warp_sum_multi16<kAcc>(acc);
        const int idx = multi_index16(lane);
        if (!(lane & 1) && idx < kPartStride) A.part[(size_t)t * kPartStride + idx] = tot;
      } else {
        const double c = warp_sum(acc[0]);
        if (lane == 0) A.part[(size_t)t * kPartStride + 9] = c;
      }
#pragma unroll
      for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
      if (++t == t1) break;
      sb = se; se = sb + A.seg_len[t];
      if (sb >= ce) break;
    }
    rec = rec_n; cf = cf_n; qi = qi_n; qj = qj_n;
    if (t == t1) break;
  }
}

// K1 for the general two-block residuals (QUATERNION_NORM, ROTATION_MAT_FNORM): same work distribution, the stored
// off-diagonal block is a full row-major 3x3 (9-double records), the diagonal contribution depends on the side.
template <bool kWriteBlocks, int kType>
__global__ void __launch_bounds__(kBlock, 1)
k_edges_general(uint32_t num_warps, uint64_t H, const uint32_t* __restrict__ warp_seg_ptr, const uint32_t* __restrict__ task_row,
                const uint32_t* __restrict__ task_begin, const uint32_t* __restrict__ task_len, const uint32_t* __restrict__ he_col,
                const double* __restrict__ inrec, const double* __restrict__ node_q, DevLoss loss,
                double* __restrict__ val, double* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_global >= num_warps) return;
  for (uint32_t t = warp_seg_ptr[warp_global]; t < warp_seg_ptr[warp_global + 1]; ++t) {
    const uint32_t row = task_row[t] & ~kSideBit;
    const uint64_t begin = task_begin[t];
    const uint32_t len = task_len[t];
    const double4 qa4 = reinterpret_cast<const double4*>(node_q)[row];
    const Q4 qrow{qa4.x, qa4.y, qa4.z, qa4.w};
    double acc[kPartStride];
#pragma unroll
    for (int k = 0; k < kPartStride; ++k) acc[k] = 0.0;
    for (uint32_t off = lane; off < len; off += 32) {
      const uint64_t h = begin + off;
      const uint32_t cf = he_col[h];
      const uint32_t col = cf & ~kSideBit;
      const bool row_is_j = (cf & kSideBit) != 0;
      const double4 qb4 = reinterpret_cast<const double4*>(node_q)[col];
      const Q4 qcol{qb4.x, qb4.y, qb4.z, qb4.w};
      const double* rec = inrec + (size_t)(h >> 5) * in_rec_doubles(1) + (h & 31);  // scalar-weight input records
      const Q4 qm{rec[0], rec[32], rec[64], rec[96]};
      GeneralTerms gt;
      general_edge_terms<kWriteBlocks, kType>(row_is_j ? qcol : qrow, row_is_j ? qrow : qcol, qm, rec[128], row_is_j, loss, gt);
      if (!row_is_j) acc[9] += 0.5 * gt.rho[0];
      if (kWriteBlocks) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] += gt.D[k];
        acc[6] += gt.g[0]; acc[7] += gt.g[1]; acc[8] += gt.g[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) val[blk_index(h, k, Rec<9>::kDoubles)] = gt.G[k];
      }
    }
    if (kWriteBlocks) {
#pragma unroll
      for (int k = 0; k < kPartStride; ++k) acc[k] = warp_sum(acc[k]);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kPartStride; ++k) part[(size_t)t * kPartStride + k] = acc[k];
      }
    } else {
      const double c = warp_sum(acc[9]);
      if (lane == 0) part[(size_t)t * kPartStride + 9] = c;
    }
  }
}

// Per view: add the task partials in task order -> tangent diagonal block Hd (packed sym 6),
// tangent gradient gt; Euclidean gradient g = Jl^T gt (for the gradient tolerance), the Euclidean
// diagonal diag(Jl^T Hd Jl) (for Jacobi scaling and the LM diagonal), total cost.
__global__ void k_node_finalize(uint32_t N, uint32_t ncb, const uint32_t* __restrict__ node_task_ptr, const double* __restrict__ part,
                                const double* __restrict__ node_JL, double* __restrict__ Hd, double* __restrict__ gt,
                                double* __restrict__ ediag, int cost_only, int stage, double* tail, double* slots, unsigned* counter,
                                DevScalars* sc, HostMailbox* mailbox, unsigned mailbox_seq, const IterParams* ip) {
  // stage 0: single GPU, everything.  Edge-sharded: stage 1 = local sums (Hd, gt, tail = {cost, bad}) which
  // the host all-reduces, stage 2 = the per-view post-processing on the reduced sums.
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[2] = {0.0, 0.0};
  double gm = 0.0;
  if (i < N) {
    double a[kPartStride];
#pragma unroll
    for (int k = 0; k < kPartStride; ++k) a[k] = 0.0;
    if (stage != 2) {
      // the row's segments: piece by piece (column block by column block), in segment order -- a fixed summation order
      for (uint32_t cb = 0; cb < ncb; ++cb)
        for (uint32_t t = node_task_ptr[cb * N + i]; t < node_task_ptr[cb * N + i + 1]; ++t) {
          if (cost_only) a[9] += part[(size_t)t * kPartStride + 9];
          else {
#pragma unroll
            for (int k = 0; k < kPartStride; ++k) a[k] += part[(size_t)t * kPartStride + k];
          }
        }
      v[0] = a[9];
      if (!cost_only) {
#pragma unroll
        for (int k = 0; k < 6; ++k) Hd[6 * (size_t)i + k] = a[k];
        gt[3 * (size_t)i] = a[6]; gt[3 * (size_t)i + 1] = a[7]; gt[3 * (size_t)i + 2] = a[8];
      }
    } else if (!cost_only) {
#pragma unroll
      for (int k = 0; k < 6; ++k) a[k] = Hd[6 * (size_t)i + k];
      a[6] = gt[3 * (size_t)i]; a[7] = gt[3 * (size_t)i + 1]; a[8] = gt[3 * (size_t)i + 2];
    }
    if (!cost_only && stage != 1) {
      double J[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) J[k] = node_JL[9 * (size_t)i + k];
      double He[6];
      congruence(J, a, He);
      ediag[3 * (size_t)i] = He[0]; ediag[3 * (size_t)i + 1] = He[3]; ediag[3 * (size_t)i + 2] = He[5];
#pragma unroll
      for (int c = 0; c < 3; ++c) gm = fmax(gm, fabs(J[c] * a[6] + J[3 + c] * a[7] + J[6 + c] * a[8]));
      if (!(isfinite(a[0]) && isfinite(a[3]) && isfinite(a[5]) && isfinite(a[6]) && isfinite(a[7]) && isfinite(a[8]))) v[1] = 1.0;
    }
  }
  // max |g|: block max -> atomicMax on the bit pattern (non-negative doubles order like uint64)
  if (!cost_only && stage != 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    if ((threadIdx.x & 31) == 0 && gm > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(&sc->gmax), (unsigned long long)__double_as_longlong(gm));
  }
  double tot[2];
  if (grid_sum<2>(v, slots, counter, tot) && threadIdx.x == 0) {
    if (stage == 1) { tail[0] = tot[0]; tail[1] = tot[1]; return; }
    if (stage == 2) { tot[0] = tail[0]; tot[1] += tail[1]; }
    sc->cost = tot[0];
    if (tot[1] != 0.0 || !isfinite(tot[0])) sc->bad = 1;
    if (mailbox) {
      if (ip) mailbox_seq = ip->seq;
      sc->t_end = gtimer_ns();
      // every scalar of this batch is final: the kernels that wrote them precede this one in the stream, this block is
      // the last one of this kernel (grid_sum) and has fenced.  Copy, fence to the system, publish.
      __threadfence();
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(sc);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(&mailbox->sc);
#pragma unroll
      for (int k = 0; k < (int)(sizeof(DevScalars) / 8); ++k) dst[k] = __ldcg(src + k);
      __threadfence_system();
      mailbox->seq = mailbox_seq;
    }
  }
}

// Edge-sharded form of k_node_finalize with the cross-GPU reduction INSIDE the kernel (no NCCL call: the trust-region batch
// stays one CUDA graph).  Per view: the shard-local sums [Hd 6 | gt 3] go out as tagged cells into the exchange block of every
// rank (LLCell, ra_common.cuh; sequence number = evaluation counter sc->eseq + 1), then the thread waits for the W contributions
// of its view in its own block, adds them in rank order (bitwise identical everywhere) and post-processes as k_node_finalize
// does.  The cost and the bad flag travel the same way through the tail cells once this rank's grid-wide sum is complete.
// Spinning on peers is safe in a plain launch: what a rank waits for is pushed by its peers before THEY wait for anything.
__global__ void k_node_finalize_ll(uint32_t N, uint32_t ncb, const uint32_t* __restrict__ node_task_ptr, const double* __restrict__ part,
                                   const double* __restrict__ node_JL, double* __restrict__ Hd, double* __restrict__ gt,
                                   double* __restrict__ ediag, double* slots, unsigned* counter, DevScalars* sc, HostMailbox* mailbox,
                                   unsigned mailbox_seq, const IterParams* ip, PeerPtrs peers, int world, int rank, int owner_mode) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned eseq = (unsigned)sc->eseq + 1u;
  int bad = 0;
  double v[2] = {0.0, 0.0};
  if (i < N) {
    double a[kPartStride];
#pragma unroll
    for (int k = 0; k < kPartStride; ++k) a[k] = 0.0;
    for (uint32_t cb = 0; cb < ncb; ++cb)
      for (uint32_t t = node_task_ptr[cb * N + i]; t < node_task_ptr[cb * N + i + 1]; ++t) {
#pragma unroll
        for (int k = 0; k < kPartStride; ++k) a[k] += part[(size_t)t * kPartStride + k];
      }
    v[0] = a[9];
    const int r0 = owner_mode ? (int)(i % (uint32_t)world) : 0, r1 = owner_mode ? r0 + 1 : world;
    for (int r = r0; r < r1; ++r) {
      LLCell* dst = peers.p[r] + ll_lin_offset(N, world, eseq, rank) + 9 * (size_t)i;
#pragma unroll
      for (int k = 0; k < 9; ++k) ll_store(dst + k, a[k], eseq);
    }
    if (owner_mode && r0 == rank) {  // this rank owns view i: add the W contributions in rank order, push the sums to everybody
      double t[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) t[k] = 0.0;
      for (int r = 0; r < world; ++r) {
        double b9[9];
        ll_wait_n<9>(peers.p[rank] + ll_lin_offset(N, world, eseq, r) + 9 * (size_t)i, eseq, b9, &bad);
#pragma unroll
        for (int k = 0; k < 9; ++k) t[k] += b9[k];
      }
      for (int r = 0; r < world; ++r) {
        LLCell* dst = peers.p[r] + ll_linred_offset(N, world, eseq) + 9 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 9; ++k) ll_store(dst + k, t[k], eseq);
      }
    }
  }
  // this rank's cost: deterministic grid sum; its last block sends it to everybody
  double tot[2];
  if (grid_sum<2>(v, slots, counter, tot) && threadIdx.x == 0) {
    for (int r = 0; r < world; ++r) {
      LLCell* dst = peers.p[r] + ll_tail_offset(N, world, eseq, rank);
      ll_store(dst, tot[0], eseq);
    }
  }
  double gm = 0.0;
  double w[2] = {0.0, 0.0};
  if (i < N) {
    double a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = 0.0;
    const LLCell* mine = peers.p[rank];
    if (owner_mode) {
      ll_wait_n<9>(mine + ll_linred_offset(N, world, eseq) + 9 * (size_t)i, eseq, a, &bad);
    } else {
      for (int r = 0; r < world; ++r) {
        double b9[9];
        ll_wait_n<9>(mine + ll_lin_offset(N, world, eseq, r) + 9 * (size_t)i, eseq, b9, &bad);
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] += b9[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) Hd[6 * (size_t)i + k] = a[k];
    gt[3 * (size_t)i] = a[6]; gt[3 * (size_t)i + 1] = a[7]; gt[3 * (size_t)i + 2] = a[8];
    double J[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) J[k] = node_JL[9 * (size_t)i + k];
    double He[6];
    congruence(J, a, He);
    ediag[3 * (size_t)i] = He[0]; ediag[3 * (size_t)i + 1] = He[3]; ediag[3 * (size_t)i + 2] = He[5];
#pragma unroll
    for (int c = 0; c < 3; ++c) gm = fmax(gm, fabs(J[c] * a[6] + J[3 + c] * a[7] + J[6 + c] * a[8]));
    if (!(isfinite(a[0]) && isfinite(a[3]) && isfinite(a[5]) && isfinite(a[6]) && isfinite(a[7]) && isfinite(a[8]))) w[0] = 1.0;
  }
  if (bad) w[1] = 1.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
  if ((threadIdx.x & 31) == 0 && gm > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(&sc->gmax), (unsigned long long)__double_as_longlong(gm));
  // second grid-wide sum (its own slots and counter): non-finite flag, lost-peer flag; its last block closes the evaluation
  double tot2[2];
  if (grid_sum<2>(w, slots + 2 * (size_t)gridDim.x + 8, counter + 1, tot2) && threadIdx.x == 0) {
    const LLCell* mine = peers.p[rank];
    double cost = 0.0;
    for (int r = 0; r < world; ++r) cost += ll_wait(mine + ll_tail_offset(N, world, eseq, r), eseq, &bad);
    sc->cost = cost;
    sc->eseq = (int)eseq;
    if (tot2[0] != 0.0 || !isfinite(cost)) sc->bad = 1;
    if (tot2[1] != 0.0 || bad) sc->bad = 2;
    if (mailbox) {
      if (ip) mailbox_seq = ip->seq;
      sc->t_end = gtimer_ns();
      __threadfence();
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(sc);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(&mailbox->sc);
#pragma unroll
      for (int k = 0; k < (int)(sizeof(DevScalars) / 8); ++k) dst[k] = __ldcg(src + k);
      __threadfence_system();
      mailbox->seq = mailbox_seq;
    }
  }
}

// Jacobi scaling, estimated once at the initial point (Ceres: scale_c = 1/(1 + |J_col c|)).
__global__ void k_jacobi_scale(uint32_t n3, const double* __restrict__ ediag, double* __restrict__ scale, int enabled) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n3) scale[c] = enabled ? 1.0 / (1.0 + sqrt(ediag[c])) : 1.0;
}

}  // namespace