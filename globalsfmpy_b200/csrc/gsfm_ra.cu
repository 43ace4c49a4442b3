// gsfm_ra.cu -- B200 (sm_100a) robust rotation averaging behind the C ABI of include/gsfm_ra.h.
//
// Replaces the Ceres solve inside GSfMNonlinearRotationEstimator
// (reference src/GSfM_nonlinear_rotation_estimator.cpp:24-80, 201-309): per-edge residuals,
// SO(3) Jacobians, covariance whitening and robust reweighting (K1), the block-3x3 normal
// equations, and a block-Jacobi PCG whose SpMV is K2 -- all resident in HBM, fp64 throughout.
//
// Data layout.  The view graph is stored as HALF-EDGES: every edge (i,j)
// appears once in row i and once in row j, sorted by (row, col); a row's half-edges are
// contiguous, so per-view sums (diagonal block, gradient) are segmented reductions over a
// contiguous range and need no atomics.  Rows are cut into TASKS of <= kTaskLen half-edges; one
// warp owns one task, lanes stride the task with fully coalesced planar (SoA) loads.  The
// normal-equation matrix lives in the LEFT TANGENT frame (see so3_device.cuh): H = D^T Ht D with
// D = blockdiag(Jl(omega_i)), so the per-edge kernel never touches the per-view Jl factors; the
// Euclidean (angle-axis) Levenberg-Marquardt of Ceres is reproduced exactly by transforming the
// LM diagonal per view.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gsfm_ra.h"
#include "so3_device.cuh"

using namespace gsfm;
namespace cg = cooperative_groups;

namespace {

constexpr int kBlock = 256;        // threads per block of every kernel
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr uint32_t kSideBit = 0x80000000u;  // he_col bit 31: the ROW view is the j (second) view of the edge
constexpr int kPartStride = 10;    // per-task partial: diag(6) grad(3) cost(1)
// The block matrix is stored as CHUNK RECORDS of 32 consecutive half-edges:
//   { double blk[kBlk][32]; uint32_t col[32]; }   128 B aligned,
// so 32 lanes still read/write 256 contiguous bytes per block component (coalesced), and K2 can
// pull a whole record into shared memory with ONE bulk async copy (TMA, cp.async.bulk).
//   kBlk = 6: symmetric off-diagonal block -S packed (00,01,02,11,12,22) -- every residual that is a function of the
//             error rotation (types 2..8) is a Laplacian stencil in the body frame;  1664 B per record (52 B / half-edge)
//   kBlk = 9: general row-major block (QUATERNION_NORM, ROTATION_MAT_FNORM);         2432 B per record (76 B / half-edge)
template <int kBlk>
struct Rec {
  static constexpr int kDoubles = kBlk * 32 + 16;
  static constexpr int kBytes = kDoubles * 8;
  static constexpr int kColOffset = kBlk * 32;  // doubles: col[] starts here
};
constexpr int kStages = 4;                       // TMA ring depth per warp
constexpr int spmv_smem_bytes(int blk) { return kWarpsPerBlock * kStages * (blk * 32 + 16) * 8 + kWarpsPerBlock * kStages * 8; }

__device__ __host__ __forceinline__ size_t blk_index(uint64_t h, int k, int rec_doubles) { return (size_t)(h >> 5) * rec_doubles + (size_t)k * 32 + (h & 31); }

thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

}  // namespace
namespace gsfm_io {
// bridge for the host-only translation unit gsfm_io.cpp: same thread-local last-error slot
void set_io_error(const std::string& msg) { g_last_error = msg; }
}  // namespace gsfm_io
namespace {

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t err__ = (expr);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
      return GSFM_RA_ERR_CUDA;                                                                \
    }                                                                                         \
  } while (0)

#define RA_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != 0) return rc__; \
  } while (0)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Scalars living on the device for the whole solve (one cache line group).
struct DevScalars {
  // evaluation
  double cost;          // sum 1/2 rho at the last evaluated point
  double gmax;          // max |g| (Euclidean gradient) at the last evaluated point
  double xnorm2;        // |omega|^2 of the last evaluated point
  // PCG
  double rz, pAp, alpha, beta, rr, bb;
  int pcg_iter, pcg_done, pcg_breakdown, pad0;
  // step
  double dg, dHd, step2;  // delta.g, delta.H.delta, |delta|^2 (Euclidean step)
  int bad;                // non-finite detected (1) / a peer GPU did not show up (2)
  int xseq;               // cross-GPU exchange sequence number (multi-GPU persistent PCG)
  int bar_seq, pad1;      // grid barrier sequence number of the persistent PCG kernel
  // %globaltimer stamps of the trust-region batch: start of k_prepare_solve, end of k_apply_step, end of k_node_finalize
  unsigned long long t_begin, t_linear_end, t_end;
};
// What changes from one trust-region batch to the next when the batch is replayed as a CUDA graph: read by the kernels
// from device memory, refreshed by the graph's first node (a 16-byte H2D copy from pinned host memory).
struct IterParams {
  double mu;
  unsigned seq, pad;
};
static_assert(sizeof(DevScalars) % 8 == 0, "DevScalars is copied to the host mailbox in 8-byte words");

// Host mailbox (pinned, mapped into the device): the last kernel of a trust-region batch copies the device scalars here and
// then publishes the batch's sequence number, so the host learns the outcome by polling its own memory -- no D2H copy
// operation, no stream synchronisation on the critical path of an iteration.
struct HostMailbox {
  DevScalars sc;
  volatile unsigned seq;
};

__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Deterministic grid-wide sum of NV values: block tree -> per-block slot -> the LAST block to
// arrive adds the slots in a fixed order.  Returns true in the last block (all threads), with the
// totals in `tot` (valid in thread 0 only).
template <int NV>
__device__ bool grid_sum(double (&v)[NV], double* slots, unsigned* counter, double (&tot)[NV]) {
  __shared__ double sm[NV][kWarpsPerBlock];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int w = 0; w < kWarpsPerBlock; ++w) s += sm[k][w];
      slots[(size_t)blockIdx.x * NV + k] = s;
    }
    __threadfence();
    const unsigned ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  // the last block: warp 0 adds the slots, lane l the blocks l, l+32, ... in order, then a butterfly -- a fixed order
  // whatever block happens to be last, and all loads of a lane are independent (no serial chain of L2 round trips)
  if (warp == 0) {
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32)
#pragma unroll
      for (int k = 0; k < NV; ++k) acc[k] += __ldcg(slots + (size_t)b * NV + k);
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = warp_sum(acc[k]);
    if (lane == 0) *counter = 0u;
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// K0 setup: per half-edge, gather the edge's measurement and weight into K1's INPUT RECORDS, one record per 32
// consecutive half-edges:
//   { double qij[4][32]; double U[kU][32]; uint32_t col[32]; uint32_t row[32]; }      128 B aligned
//   qij = unit quaternion of omega_ij, U = whitening (rotation_estimator.cpp:251-288), kU = 6 (upper triangle) or 1
//   (scalar weight); col carries the side bit.  1536 B (kU = 1) / 2816 B (kU = 6) per record: K1 pulls a record with
//   ONE bulk async copy.
// ------------------------------------------------------------------------------------------
__device__ __host__ __forceinline__ int in_rec_doubles(int ku) { return (4 + ku) * 32 + 32; }

__global__ void k_setup_halfedges(uint64_t H, int ku, const uint32_t* __restrict__ he_edge, const uint32_t* __restrict__ he_row,
                                  const uint32_t* __restrict__ he_col, const double* __restrict__ omega_ij, const double* __restrict__ cov6,
                                  const double* __restrict__ weight, int error_type, double* __restrict__ inrec) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const uint64_t k = he_edge[h];
  const Q4 q = aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
  double* rec = inrec + (size_t)(h >> 5) * in_rec_doubles(ku);
  const int lane = (int)(h & 31);
  rec[lane] = q.w; rec[32 + lane] = q.x; rec[64 + lane] = q.y; rec[96 + lane] = q.z;
  double c6[6] = {0, 0, 0, 0, 0, 0};
  if (cov6) for (int t = 0; t < 6; ++t) c6[t] = cov6[6 * k + t];
  double u[6];
  whiten(error_type, c6, weight ? weight[k] : 1.0, u);
  for (int t = 0; t < ku; ++t) rec[(4 + t) * 32 + lane] = u[t];
  uint32_t* idx = reinterpret_cast<uint32_t*>(rec + (4 + ku) * 32);
  idx[lane] = he_col[h];
  idx[32 + lane] = he_row[h];
}

// ------------------------------------------------------------------------------------------
// Structure build on the device (one-time per problem): half-edge keys -> radix sort -> rows,
// row pointers, duplicate / range checks, balanced warp partitions and their segments.
// ------------------------------------------------------------------------------------------
__global__ void k_build_keys(uint64_t E, uint32_t N, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, uint64_t* __restrict__ keys,
                             uint32_t* __restrict__ vals, int* err) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  uint32_t i = ei[k], j = ej[k];
  if (i >= N || j >= N || i == j) { atomicMax(err, 1); i = 0; j = (N > 1) ? 1 : 0; }
  keys[2 * k] = (uint64_t)i * N + j;     vals[2 * k] = (uint32_t)k;
  keys[2 * k + 1] = (uint64_t)j * N + i; vals[2 * k + 1] = (uint32_t)k | kSideBit;
}
__global__ void k_unpack_keys(uint64_t H, uint32_t N, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t* __restrict__ he_row,
                              uint32_t* __restrict__ he_col, uint32_t* __restrict__ he_edge, int* err) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const uint64_t key = keys[h];
  const uint32_t v = vals[h];
  he_row[h] = (uint32_t)(key / N);
  he_col[h] = (uint32_t)(key % N) | (v & kSideBit);
  he_edge[h] = v & ~kSideBit;
  if (h > 0 && keys[h - 1] == key) atomicMax(err, 2);  // the same view pair twice
}
__global__ void k_rowptr(uint32_t N, uint64_t H, const uint64_t* __restrict__ keys, uint32_t* __restrict__ rowptr) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > N) return;
  const uint64_t target = (uint64_t)r * N;
  uint64_t lo = 0, hi = H;
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
  rowptr[r] = (uint32_t)lo;
}
__global__ void k_row_flags(uint32_t N, const uint32_t* __restrict__ rowptr, uint32_t* __restrict__ nonempty, uint32_t* __restrict__ isoflag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > N) return;
  const uint32_t ne = (r < N && rowptr[r + 1] > rowptr[r]) ? 1u : 0u;
  nonempty[r] = ne;
  isoflag[r] = (r < N) ? 1u - ne : 0u;
}
__global__ void k_iso_fill(uint32_t N, const uint32_t* __restrict__ isoflag, const uint32_t* __restrict__ iso_rank, uint32_t* __restrict__ iso) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < N && isoflag[r]) iso[iso_rank[r]] = r;
}
__global__ void k_part_count(uint32_t nw, uint32_t per, uint64_t H, const uint32_t* __restrict__ he_row, const uint32_t* __restrict__ nz_rank,
                             uint32_t* __restrict__ nseg) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nw) return;
  if (w == nw) { nseg[w] = 0; return; }
  const uint64_t lo = (uint64_t)w * per, hi = min(H, lo + per);
  nseg[w] = nz_rank[he_row[hi - 1]] - nz_rank[he_row[lo]] + 1;
}
__global__ void k_part_fill(uint32_t nw, uint32_t per, uint64_t H, const uint32_t* __restrict__ he_row, const uint32_t* __restrict__ rowptr,
                            const uint32_t* __restrict__ warp_seg_ptr, uint32_t* __restrict__ seg_row, uint32_t* __restrict__ seg_begin,
                            uint32_t* __restrict__ seg_len) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const uint64_t lo = (uint64_t)w * per, hi = min(H, lo + per);
  uint32_t t = warp_seg_ptr[w];
  uint64_t h = lo;
  uint32_t r = he_row[lo];
  while (h < hi) {
    while (rowptr[r + 1] <= h) ++r;
    const uint64_t end = min(hi, (uint64_t)rowptr[r + 1]);
    seg_row[t] = r | ((h > rowptr[r]) ? kSideBit : 0u);  // bit 31: continuation of a row begun in an earlier range
    seg_begin[t] = (uint32_t)h;
    seg_len[t] = (uint32_t)(end - h);
    ++t;
    h = end;
  }
}
__global__ void k_node_seg_count(uint32_t N, uint32_t per, const uint32_t* __restrict__ rowptr, uint32_t* __restrict__ cnt) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > N) return;
  uint32_t c = 0;
  if (r < N && rowptr[r + 1] > rowptr[r]) c = (rowptr[r + 1] - 1) / per - rowptr[r] / per + 1;
  cnt[r] = c;
}

// Column indices live inside the chunk records of both block buffers (written once).
__global__ void k_embed_cols(uint64_t H, int blk, const uint32_t* __restrict__ he_col, double* rec0, double* rec1) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const size_t w = (size_t)(h >> 5) * (blk * 32 + 16) + blk * 32;
  reinterpret_cast<uint32_t*>(rec0 + w)[h & 31] = he_col[h];
  reinterpret_cast<uint32_t*>(rec1 + w)[h & 31] = he_col[h];
}

// Per view: quaternion + the factor D of d(beta) = D d(parameters) at the current estimate; also |x|^2.
//   angle-axis parameters:  D = Jr(omega) = Jl(-omega)   (R(omega + d omega) = R(omega) Exp(Jr d omega))
//   quaternion parameters with EigenQuaternionParameterization (x (+) delta = [sin|d| d/|d|, cos|d|] (x) x, i.e. a LEFT
//   perturbation phi = 2 delta):  beta = R^T phi  ->  D = 2 R^T;  |x|^2 = 1 per unit quaternion
// The array keeps its historical name node_JL.
// per-view body of k_node_prep; returns the view's contribution to |x|^2
__device__ __forceinline__ double node_prep_view(uint32_t i, const double* w3, double* __restrict__ node_q, double* __restrict__ node_JL, int manifold) {
  const double wx = w3[0], wy = w3[1], wz = w3[2];
  const Q4 q = aa_to_quat(wx, wy, wz);
  reinterpret_cast<double4*>(node_q)[i] = make_double4(q.w, q.x, q.y, q.z);
  double J[9];
  double xn;
  if (manifold) {
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
// TMA record pipeline shared by K1 and K2: every warp keeps kStages bulk async copies (cp.async.bulk, one record
// each, completion on a warp-private mbarrier) in flight.  Bytes in flight are set by the ring depth, not by
// registers or occupancy.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_bulk_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// L2 residency of the matrix stream (K2): a PCG solve reads the same records once per CG step.  When the stream is somewhat
// larger than the L2 (104 MB vs 126 MB of L2 shared with everything else at 1M edges) plain LRU keeps almost nothing of a
// cyclic sweep; loading `keep8` of every 8 records with an evict_last policy and the rest with evict_first pins a fixed
// fraction of the matrix from one pass to the next and only the remainder streams from HBM (measured at 1M edges,
// profiles/r01_j_l2_keep.txt: K2 alone 25.9 -> 22.6 us, CG step 33.3 -> 31.9 us with 6 of 8).  keep8 = 0 (streams far larger
// than the L2, or small enough for LRU to hold them): no hints.  The host picks it from the device's L2 size.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

struct WarpPipe {
  double* ring;    // this warp's kStages records in shared memory
  uint64_t* bars;  // this warp's kStages mbarriers
  uint32_t pos;    // records consumed since init: ring slot = pos % kStages, phase = (pos / kStages) & 1
};

template <int kRecBytes>
__device__ __forceinline__ void pipe_init_bytes(WarpPipe& wp, unsigned char* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  wp.ring = reinterpret_cast<double*>(smem + (size_t)warp * kStages * kRecBytes);
  wp.bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarpsPerBlock * kStages * kRecBytes) + warp * kStages;
  wp.pos = 0;
  if (lane == 0) {
    for (int st = 0; st < kStages; ++st) mbar_init(&wp.bars[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}
template <int kBlk>
__device__ __forceinline__ void pipe_init(WarpPipe& wp, unsigned char* smem) { pipe_init_bytes<Rec<kBlk>::kBytes>(wp, smem); }

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
};

template <bool kWriteBlocks, int kResidual, bool kScalarU, int kLoss>
// two blocks (16 warps) per SM: three (<= 80 registers) spill and measure 10 % slower (profiles/r01_g_microbench.txt item 5)
__global__ void __launch_bounds__(kBlock, 2) k_edges(const K1Args A) {
  constexpr int kU = (kScalarU || kResidual == 1) ? 1 : 6;
  constexpr int kRD = (4 + kU) * 32 + 32;  // doubles per input record
  constexpr int kRB = kRD * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= A.num_warps) return;
  WarpPipe wp;
  pipe_init_bytes<kRB>(wp, smem_raw);
  const uint64_t lo = (uint64_t)gw * A.warp_span, hi = min(A.H, lo + A.warp_span);
  const uint32_t t0 = A.warp_seg_ptr[gw], t1 = A.warp_seg_ptr[gw + 1];
  const uint32_t nrec = (uint32_t)((hi - lo + 31) >> 5);
  const double* src = A.inrec + (size_t)(lo >> 5) * kRD;
  auto issue = [&](uint32_t c) {
    if (lane == 0) {
      const uint32_t st = c % kStages;
      mbar_expect_tx(&wp.bars[st], kRB);
      tma_load_bulk(wp.ring + (size_t)st * kRD, src + (size_t)c * kRD, kRB, &wp.bars[st]);
    }
  };
  auto wait_rec = [&](uint32_t c) -> const double* {
    const uint32_t st = c % kStages;
    mbar_wait(&wp.bars[st], (c / kStages) & 1u);
    return wp.ring + (size_t)st * kRD;
  };
  auto gather = [&](const double* rec, uint64_t h, uint32_t& cf, Q4& qrow, Q4& qcol) {
    const uint32_t* idx = reinterpret_cast<const uint32_t*>(rec + (4 + kU) * 32);
    cf = idx[lane];
    uint32_t row = idx[32 + lane];
    if (h >= hi) { cf = 0; row = 0; }  // padding lanes of the last record
    const double4 a = reinterpret_cast<const double4*>(A.node_q)[row];
    const double4 b = reinterpret_cast<const double4*>(A.node_q)[cf & ~kSideBit];
    qrow = Q4{a.x, a.y, a.z, a.w};
    qcol = Q4{b.x, b.y, b.z, b.w};
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
  Q4 qrow, qcol;
  gather(rec, lo + lane, cf, qrow, qcol);
  for (uint32_t c = 0; c < nrec; ++c) {
    const uint64_t cb = lo + ((uint64_t)c << 5), ce = cb + 32, h = cb + lane;
    const double* rec_n = nullptr;
    uint32_t cf_n = 0;
    Q4 qrow_n{1, 0, 0, 0}, qcol_n{1, 0, 0, 0};
    if (c + 1 < nrec) { rec_n = wait_rec(c + 1); gather(rec_n, ce + lane, cf_n, qrow_n, qcol_n); }
    const bool row_is_j = (cf & kSideBit) != 0;
    const Q4 qm{rec[lane], rec[32 + lane], rec[64 + lane], rec[96 + lane]};
    double u[6];
#pragma unroll
    for (int k = 0; k < kU; ++k) u[k] = rec[(4 + k) * 32 + lane];
    // this record's slot can be refilled as soon as every lane has read it
    __syncwarp();
    if (c + kStages < nrec) issue(c + kStages);
    EdgeTerms et;
    edge_terms<kWriteBlocks, kResidual, kScalarU, kLoss, true>(row_is_j ? qcol : qrow, row_is_j ? qrow : qcol, qm, u, A.loss, et);
    double cur[kAcc];
    if (kWriteBlocks) {
      // both rows of the edge: diag += S, block(row, col) = -S; gradient: +v in row j, -v in row i; cost once, in row i
      const double sgn = row_is_j ? 1.0 : -1.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) cur[k] = et.S[k];
      cur[6] = sgn * et.v[0]; cur[7] = sgn * et.v[1]; cur[8] = sgn * et.v[2];
      cur[kAcc - 1] = row_is_j ? 0.0 : 0.5 * et.rho[0];
      if (h < hi) {
#pragma unroll
        for (int k = 0; k < 6; ++k) A.val[blk_index(h, k, Rec<6>::kDoubles)] = -et.S[k];
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
#pragma unroll
      for (int k = 0; k < kAcc; ++k) acc[k] = warp_sum(acc[k]);
      if (lane == 0) {
        if (kWriteBlocks) {
#pragma unroll
          for (int k = 0; k < kPartStride; ++k) A.part[(size_t)t * kPartStride + k] = acc[k];
        } else {
          A.part[(size_t)t * kPartStride + 9] = acc[0];
        }
      }
#pragma unroll
      for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
      if (++t == t1) break;
      sb = se; se = sb + A.seg_len[t];
      if (sb >= ce) break;
    }
    rec = rec_n; cf = cf_n; qrow = qrow_n; qcol = qcol_n;
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
__global__ void k_node_finalize(uint32_t N, const uint32_t* __restrict__ node_task_ptr, const double* __restrict__ part,
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
      for (uint32_t t = node_task_ptr[i]; t < node_task_ptr[i + 1]; ++t) {
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

// Jacobi scaling, estimated once at the initial point (Ceres: scale_c = 1/(1 + |J_col c|)).
__global__ void k_jacobi_scale(uint32_t n3, const double* __restrict__ ediag, double* __restrict__ scale, int enabled) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n3) scale[c] = enabled ? 1.0 / (1.0 + sqrt(ediag[c])) : 1.0;
}

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
                                            bool nogather, uint32_t keep8, Finish&& finish) {
  constexpr int kRD = Rec<kBlk>::kDoubles;
  const int lane = threadIdx.x & 31;
  const uint32_t nrec = (uint32_t)((hi - lo + 31) >> 5);
  const uint32_t base = wp.pos;
  const double* src = recs + (size_t)(lo >> 5) * kRD;
  uint64_t pol_keep = 0, pol_stream = 0;
  if (keep8) { pol_keep = l2_policy_evict_last(); pol_stream = l2_policy_evict_first(); }
  const uint32_t rec0 = (uint32_t)(lo >> 5);
  auto issue = [&](uint32_t c) {
    if (lane == 0) {
      const uint32_t st = (base + c) % kStages;
      mbar_expect_tx(&wp.bars[st], Rec<kBlk>::kBytes);
      if (keep8)
        tma_load_bulk_hint(wp.ring + (size_t)st * kRD, src + (size_t)c * kRD, Rec<kBlk>::kBytes, &wp.bars[st],
                           ((rec0 + c) & 7u) < keep8 ? pol_keep : pol_stream);
      else
        tma_load_bulk(wp.ring + (size_t)st * kRD, src + (size_t)c * kRD, Rec<kBlk>::kBytes, &wp.bars[st]);
    }
  };
  auto wait_rec = [&](uint32_t c) -> const double* {
    const uint32_t p = base + c, st = p % kStages;
    mbar_wait(&wp.bars[st], (p / kStages) & 1u);
    return wp.ring + (size_t)st * kRD;
  };
  auto gather = [&](const double* rec, uint64_t h, double& x0, double& x1, double& x2) {
    uint32_t col = reinterpret_cast<const uint32_t*>(rec + Rec<kBlk>::kColOffset)[lane] & ~kSideBit;
    if (h >= hi) col = 0;  // padding lanes of the last record
    if (nogather) { x0 = col; x1 = 1.0; x2 = 2.0; return; }  // measurement aid: stream-only ceiling
    const double4 xv = reinterpret_cast<const double4*>(x4)[col];
    x0 = xv.x; x1 = xv.y; x2 = xv.z;
  };
  for (uint32_t c = 0; c < nrec && c < (uint32_t)kStages; ++c) issue(c);
  if (nrec == 0 || t0 == t1) { wp.pos = base + nrec; return; }
  uint32_t t = t0;
  uint64_t sb = seg_begin[t], se = sb + seg_len[t];
  double y0 = 0.0, y1 = 0.0, y2 = 0.0;
  const double* rec = wait_rec(0);
  double x0, x1, x2;
  gather(rec, lo + lane, x0, x1, x2);
  for (uint32_t c = 0; c < nrec; ++c) {
    const uint64_t cb = lo + ((uint64_t)c << 5), ce = cb + 32, h = cb + lane;
    // prefetch the next record's gather
    const double* rec_n = nullptr;
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;
    if (c + 1 < nrec) { rec_n = wait_rec(c + 1); gather(rec_n, ce + lane, n0, n1, n2); }
    double v0, v1, v2;
    if (kBlk == 6) {
      const double b0 = rec[lane], b1 = rec[32 + lane], b2 = rec[64 + lane], b3 = rec[96 + lane], b4 = rec[128 + lane], b5 = rec[160 + lane];
      v0 = b0 * x0 + b1 * x1 + b2 * x2;
      v1 = b1 * x0 + b3 * x1 + b4 * x2;
      v2 = b2 * x0 + b4 * x1 + b5 * x2;
    } else {
      v0 = rec[lane] * x0 + rec[32 + lane] * x1 + rec[64 + lane] * x2;
      v1 = rec[96 + lane] * x0 + rec[128 + lane] * x1 + rec[160 + lane] * x2;
      v2 = rec[192 + lane] * x0 + rec[224 + lane] * x1 + rec[256 + lane] * x2;
    }
    // this record's slot can be refilled as soon as every lane has read it
    __syncwarp();
    if (c + kStages < nrec) issue(c + kStages);
    while (true) {
      if (h >= sb && h < se) { y0 += v0; y1 += v1; y2 += v2; }
      if (se > ce) break;  // the segment continues in the next record
      y0 = warp_sum(y0); y1 = warp_sum(y1); y2 = warp_sum(y2);
      finish(t, y0, y1, y2);
      y0 = y1 = y2 = 0.0;
      if (++t == t1) break;
      sb = se; se = sb + seg_len[t];
      if (sb >= ce) break;
    }
    rec = rec_n; x0 = n0; x1 = n1; x2 = n2;
    if (t == t1) break;
  }
  wp.pos = base + nrec;
}

template <int kBlk>
__global__ void __launch_bounds__(kBlock)
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
  spmv_stream<kBlk>(wp, recs, lo, hi, warp_seg_ptr[gw], warp_seg_ptr[gw + 1], task_begin, task_len, x4, check_done == 2, keep8,
                    [&](uint32_t t, double y0, double y1, double y2) {
                      if (lane == 0) { ypart[3 * (size_t)t] = y0; ypart[3 * (size_t)t + 1] = y1; ypart[3 * (size_t)t + 2] = y2; }
                    });
}

// y_i = Dblk_i x_i + sum of the row's task partials (+ shard-local only: the diagonal part is
// added after the cross-GPU reduction).  mode 0: write y, reduce p.y -> alpha (PCG step 1).
// mode 1: y only.
__global__ void k_spmv_finish(uint32_t N, const uint32_t* __restrict__ node_task_ptr, const double* __restrict__ ypart,
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
      for (uint32_t t = node_task_ptr[i]; t < node_task_ptr[i + 1]; ++t) {
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
  if (A.manifold) {
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

constexpr int kMaxPeers = 16;

struct PcgParams {
  uint32_t N, num_warps, n_iso, warp_span;
  uint32_t keep8;  // records of every 8 loaded with the evict_last L2 policy (0: no cache hints)
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
  // edge-sharded multi-GPU (world > 1): every rank's exchange block, mapped into this process over NVLink
  // (CUDA IPC).  Layout of one block: double y[2][3N] (partial matvec, double buffered by step parity) followed
  // by the rank's sequence flag.  peer_y[rank] / peer_flag[rank] are this rank's own block.
  int world, rank;
  double* peer_y[kMaxPeers];
  unsigned* peer_flag[kMaxPeers];
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
                                              double* sm_red /*[2*kWarpsPerBlock + 2]*/, double& out0, double& out1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v0 = warp_sum(v0); v1 = warp_sum(v1);
  if (lane == 0) { sm_red[2 * warp] = v0; sm_red[2 * warp + 1] = v1; }
  __syncthreads();
  ++seq;
  double* set = reinterpret_cast<double*>(bar_slots) + (size_t)(seq & 1u) * 2 * gridDim.x;
  if (threadIdx.x == 0) {
    double b0 = 0.0, b1 = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) { b0 += sm_red[2 * w]; b1 += sm_red[2 * w + 1]; }
    __stcg(set + 2 * (size_t)blockIdx.x, b0); __stcg(set + 2 * (size_t)blockIdx.x + 1, b1);
  }
  grid.sync();
  if (warp == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) { s0 += __ldcg(set + 2 * (size_t)b); s1 += __ldcg(set + 2 * (size_t)b + 1); }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if (lane == 0) { sm_red[2 * kWarpsPerBlock] = s0; sm_red[2 * kWarpsPerBlock + 1] = s1; }
  }
  __syncthreads();
  out0 = sm_red[2 * kWarpsPerBlock]; out1 = sm_red[2 * kWarpsPerBlock + 1];
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

// One SpMV pass over this warp's range, s = (Ht + Lam) z.  Accumulates (per lane) gamma = r.z and delta = z.s over
// the rows this lane finished.
template <int kBlk>
__device__ __forceinline__ void spmv_pass(const PcgParams& P, WarpPipe& wp, double& gamma, double& delta, double* ylocal_out = nullptr) {
  // ylocal_out != null (multi-GPU): only the shard-local off-diagonal row sums are produced, into the exchange
  // buffer; diagonal and inner products follow after the cross-GPU reduction (exchange_finish).
  const int lane = threadIdx.x & 31;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gwarp < P.num_warps) {
    const uint64_t lo = (uint64_t)gwarp * P.warp_span, hi = min(P.H, lo + P.warp_span);
    // Rows are finished by a STATIC owner -- the warp holding the row's first segment -- so every
    // sum has a fixed order and a fixed place (bit-reproducible).  A continuation segment (always the
    // first segment of a warp's range) is published at once by lane 0: parts + fence + counter.
    // Owned segments are parked one per lane and completed as a batch, so the wait / load latencies
    // of up to 32 rows overlap instead of serialising while the record stream drains.  The owner
    // spins on the row counter; the warps it waits for publish first thing in their pass and the
    // cooperative launch keeps every block resident, so the wait cannot deadlock.
    const uint32_t t0 = P.warp_seg_ptr[gwarp], t1 = P.warp_seg_ptr[gwarp + 1];
    double my0 = 0.0, my1 = 0.0, my2 = 0.0;
    uint32_t my_t = 0, nbatch = 0;
    auto flush = [&]() {
      if ((uint32_t)lane < nbatch) {
        const uint32_t row = P.seg_row[my_t] & ~kSideBit;
        const uint32_t s0 = P.node_seg_ptr[row], s1 = P.node_seg_ptr[row + 1];
        if (s1 - s0 > 1) {
          volatile unsigned* cnt = P.row_cnt + row;
          while (*cnt != s1 - s0 - 1) { }
          __threadfence();
          *cnt = 0u;
          for (uint32_t k = s0 + 1; k < s1; ++k) {
            my0 += __ldcg(P.ypart + 3 * (size_t)k); my1 += __ldcg(P.ypart + 3 * (size_t)k + 1); my2 += __ldcg(P.ypart + 3 * (size_t)k + 2);
          }
        }
        if (ylocal_out) {
          ylocal_out[3 * (size_t)row] = my0; ylocal_out[3 * (size_t)row + 1] = my1; ylocal_out[3 * (size_t)row + 2] = my2;
        } else {
          finish_row(P, row, my0, my1, my2, gamma, delta);
        }
      }
      nbatch = 0;
      __syncwarp();
    };
    spmv_stream<kBlk>(wp, P.val, lo, hi, t0, t1, P.seg_begin, P.seg_len, P.z, false, P.keep8, [&](uint32_t t, double y0, double y1, double y2) {
      const uint32_t rowf = P.seg_row[t];
      if (rowf & kSideBit) {  // continuation of a row owned by an earlier warp
        if (lane == 0) {
          __stcg(P.ypart + 3 * (size_t)t, y0); __stcg(P.ypart + 3 * (size_t)t + 1, y1); __stcg(P.ypart + 3 * (size_t)t + 2, y2);
          __threadfence();
          atomicAdd(P.row_cnt + (rowf & ~kSideBit), 1u);
        }
        return;
      }
      if ((uint32_t)lane == nbatch) { my0 = y0; my1 = y1; my2 = y2; my_t = t; }
      if (++nbatch == 32) flush();
    });
    if (nbatch) flush();
  }
  // views without any half-edge: s_i = D_i z_i  (multi-GPU: their exchange slots stay zero, nothing to do)
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; !ylocal_out && k < P.n_iso; k += gridDim.x * blockDim.x)
    finish_row(P, P.iso[k], 0.0, 0.0, 0.0, gamma, delta);
}

// Fused cross-GPU reduction of the partial matvec, inside the persistent kernel (no NCCL call, no kernel
// boundary): publish "my partial sums for step `seq` are complete" with a system-scope release, wait for every
// peer's flag, then every rank adds the partial vectors of ALL ranks in rank order straight out of peer memory
// over NVLink (bitwise identical result everywhere) and finishes the row: s_i = D_i z_i + sum, inner products.
// Buffer reuse is safe with two buffers: a rank can only reach step seq+2 after every peer published seq+1,
// i.e. after every peer finished reading step seq.
__device__ __forceinline__ void exchange_finish(const PcgParams& P, cg::grid_group& grid, unsigned seq, double& gamma, double& delta) {
  grid.sync();  // all local row sums of this step are in my exchange buffer
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    *((volatile unsigned*)P.peer_flag[P.rank]) = seq;
  }
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < P.world; ++r) {
      if (r == P.rank) continue;
      volatile unsigned* f = (volatile unsigned*)P.peer_flag[r];
      while ((int)(*f - seq) < 0) {
        if (clock64() - t0 > 8000000000ll) { P.sc->bad = 2; break; }  // ~4 s: a peer died; do not hang the GPU
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const size_t off = (size_t)(seq & 1u) * 3 * P.N;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += gridDim.x * blockDim.x) {
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    for (int r = 0; r < P.world; ++r) {
      const double* src = P.peer_y[r] + off + 3 * (size_t)i;
      y0 += __ldcv(src); y1 += __ldcv(src + 1); y2 += __ldcv(src + 2);
    }
    finish_row(P, i, y0, y1, y2, gamma, delta);
  }
}

template <int kBlk>
__global__ void __launch_bounds__(kBlock) k_pcg_persistent(PcgParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::grid_group grid = cg::this_grid();
  WarpPipe wp;
  pipe_init<kBlk>(wp, smem_raw);
  __shared__ double sm_red[2 * kWarpsPerBlock + 2];
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  unsigned bseq = (unsigned)P.sc->bar_seq;  // barrier sequence number, continues across launches
  double bb;
  bool done;
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
  double* my_y = multi ? P.peer_y[P.rank] : nullptr;
  while (!done) {
    // ---- phase A: s = (Ht + Lam) z, gamma = r.z, delta = z.s ---------------------------------
    const bool prof = P.prof != nullptr && gtid == 0;
    unsigned long long tA = 0, tB = 0, tC = 0, tE = 0, tF = 0;
    if (prof) tA = gtimer();
    double g_part = 0.0, d_part = 0.0;
    if (!multi) spmv_pass<kBlk>(P, wp, g_part, d_part);
    else {
      ++seq;
      spmv_pass<kBlk>(P, wp, g_part, d_part, my_y + (size_t)(seq & 1u) * 3 * P.N);
      exchange_finish(P, grid, seq, g_part, d_part);
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

// ---- API-only kernels (parity tests / diagnostics; not on the solve path) ------------------
// Raw per-edge outputs in EDGE order and Euclidean (angle-axis) coordinates, i.e. exactly what
// AutoDiffCostFunction::Evaluate + LossFunction::Evaluate return in the reference.
__global__ void k_eval_edges(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const double* __restrict__ omega_ij,
                             const double* __restrict__ cov6, const double* __restrict__ weight, int error_type, const double* __restrict__ node_q,
                             const double* __restrict__ node_JL, DevLoss loss, double* r, double* Ji, double* Jj, double* rho) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  const uint32_t i = ei[k], j = ej[k];
  const double4 a = reinterpret_cast<const double4*>(node_q)[i], b = reinterpret_cast<const double4*>(node_q)[j];
  const Q4 qi{a.x, a.y, a.z, a.w}, qj{b.x, b.y, b.z, b.w};
  const Q4 qm = aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
  double c6[6] = {0, 0, 0, 0, 0, 0}, u[6];
  if (cov6) for (int t = 0; t < 6; ++t) c6[t] = cov6[6 * k + t];
  whiten(error_type, c6, weight ? weight[k] : 1.0, u);
  EdgeTerms et;
  if (error_type == 2) edge_terms<true, 1>(qi, qj, qm, u, loss, et);
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
                                const double* __restrict__ node_JL, double* out_val, uint32_t* out_col) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const uint32_t row = he_row[h], col = he_col[h] & ~kSideBit;
  double B[9], T[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) B[3 * r + c] = blk_entry(val, h, blk, r, c);
  const double* Jr = node_JL + 9 * (size_t)row;
  const double* Jc = node_JL + 9 * (size_t)col;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = B[3 * r] * Jc[c] + B[3 * r + 1] * Jc[3 + c] + B[3 * r + 2] * Jc[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out_val[9 * h + 3 * r + c] = Jr[r] * T[c] + Jr[3 + r] * T[3 + c] + Jr[6 + r] * T[6 + c];
  if (out_col) out_col[h] = col;
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int make_dev_loss(const gsfm_ra_loss* in, DevLoss* L) {
  std::memset(L, 0, sizeof(*L));
  if (!in) { L->kind = kLossTrivial; L->scale = 1.0; return 0; }
  if (in->kind < 0 || in->kind > GSFM_RA_LOSS_MAGSAC9) { set_error("unknown loss kind %d", in->kind); return GSFM_RA_ERR_INVALID; }
  L->kind = in->kind; L->flags = in->flags; L->p0 = in->p[0]; L->p1 = in->p[1];
  L->scale = (in->scale == 0.0) ? 1.0 : in->scale;
  L->sq0 = in->p[0] * in->p[0];
  L->inv_sq0 = 1.0 / L->sq0;
  const bool needs_p0 = in->kind != GSFM_RA_LOSS_TRIVIAL;
  if (needs_p0 && !(in->p[0] > 0.0)) { set_error("loss parameter p[0] must be > 0"); return GSFM_RA_ERR_INVALID; }
  if ((in->kind == GSFM_RA_LOSS_TOLERANT || in->kind == GSFM_RA_LOSS_GEMANMCCLURE) && !(in->p[1] > 0.0)) {
    set_error("loss parameter p[1] must be > 0"); return GSFM_RA_ERR_INVALID;
  }
  if (in->kind >= GSFM_RA_LOSS_MAGSAC3) {
    // include/gamma_values.cpp:6-11, 384-389, 780-785
    double C, quant, gk; int nu, n;
    if (in->kind == GSFM_RA_LOSS_MAGSAC3) { nu = 3; C = 4.029720004054876e-01; quant = 3.368214175218727; gk = 3.439485560754856e-03; n = 36843; }
    else if (in->kind == GSFM_RA_LOSS_MAGSAC4) { nu = 4; C = 2.525252525252525e-01; quant = 3.643721193503644e+00; gk = 3.611260617758625e-03; n = 38683; }
    else { nu = 9; C = 3.837828575290349e-03; quant = 4.654674460524809e+00; gk = 3.344206155099048e-02; n = 48553; }
    const double sigma = in->p[0];
    L->nu = nu; L->table_size = n;
    L->sq_sigma = sigma * sigma;
    L->sq_sigma_max_2 = 2.0 * L->sq_sigma;
    L->cubed_sigma = L->sq_sigma * sigma;
    L->clamp_s = quant * quant * L->sq_sigma;
    const double dof = (nu - 1.0) / 2.0;
    L->Ctd = C * std::pow(2.0, dof);
    L->one_over_sigma = L->Ctd / sigma;
    L->gamma_k = gk;
    L->weight_zero = L->one_over_sigma * (std::tgamma(dof) - gk);
    L->expo = nu / 2.0 - 1.5;
  }
  return 0;
}

bool type_needs_cov(int t) { return t == 3 || t == 6 || t == 7 || t == 8; }

int check_problem(const gsfm_ra_problem* p) {
  if (!p) { set_error("problem is NULL"); return GSFM_RA_ERR_INVALID; }
  if (p->num_views == 0 || p->num_edges == 0) { set_error("empty problem (views=%u edges=%llu)", p->num_views, (unsigned long long)p->num_edges); return GSFM_RA_ERR_INVALID; }
  if (!p->edge_i || !p->edge_j || !p->omega_ij) { set_error("edge arrays are NULL"); return GSFM_RA_ERR_INVALID; }
  if (p->num_views >= kSideBit) { set_error("too many views"); return GSFM_RA_ERR_INVALID; }
  if (p->num_edges >= (1ull << 31)) { set_error("too many edges for 32-bit half-edge offsets"); return GSFM_RA_ERR_UNSUPPORTED; }
  if (p->error_type < GSFM_RA_QUATERNION_NORM || p->error_type > GSFM_RA_ANGLE_AXIS_COVNORM) { set_error("unknown error_type %d", p->error_type); return GSFM_RA_ERR_INVALID; }
  if (type_needs_cov(p->error_type) && !p->cov6) { set_error("error_type %d needs cov6", p->error_type); return GSFM_RA_ERR_INVALID; }
  return 0;
}

int select_device(int device, int* out) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible: libgsfm_ra has no CPU fallback");
    return GSFM_RA_ERR_NO_DEVICE;
  }
  if (device < 0) { CUDA_TRY(cudaGetDevice(&device)); }
  if (device >= n) { set_error("device %d out of range (%d visible)", device, n); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(device));
  *out = device;
  return 0;
}

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) on the stream that
// is "current" for this thread while a solver is being built; the pool's release threshold is raised so a
// second gsfm_ra_solve() call reuses the memory of the first instead of paying cudaMalloc/cudaFree again.
thread_local cudaStream_t g_alloc_stream = nullptr;
thread_local bool g_alloc_async = false;

struct AllocScope {
  cudaStream_t prev_s;
  bool prev_a;
  explicit AllocScope(cudaStream_t st) : prev_s(g_alloc_stream), prev_a(g_alloc_async) { g_alloc_stream = st; g_alloc_async = true; }
  ~AllocScope() { g_alloc_stream = prev_s; g_alloc_async = prev_a; }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool async = false;
  cudaStream_t st = nullptr;
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return 0;
    if (g_alloc_async) {
      async = true; st = g_alloc_stream;
      CUDA_TRY(cudaMallocAsync(&p, count * sizeof(T), st));
    } else {
      async = false;
      CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
    }
    return 0;
  }
  void release() {
    if (p) { if (async) cudaFreeAsync(p, st); else cudaFree(p); }
    p = nullptr; n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Owns the solver's stream; declared FIRST in the solver so it is destroyed LAST, after every DevBuf has
// queued its cudaFreeAsync on it.
struct StreamHolder {
  cudaStream_t s = nullptr;
  ~StreamHolder() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
};

// Per-device facts that are expensive to query: cached for the life of the process.
struct DeviceInfo {
  bool ready = false;
  int sm_count = 0, coop = 0, occ_k2[2] = {1, 1};  // occ_k2[0]: 6-double records, [1]: 9-double records
  int l2_bytes = 0;
};
DeviceInfo g_device_info[64];

inline unsigned grid_for(uint64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace

// ------------------------------------------------------------------------------------------
// the resident solver
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so the single-GPU library has no link dependency on it.  When the
// host process already loaded a libnccl.so.2 (e.g. PyTorch's bundled one) the same copy is reused.
// ------------------------------------------------------------------------------------------
namespace ncclx {
constexpr int kUniqueIdBytes = 128;
struct UniqueId { char internal[kUniqueIdBytes]; };
typedef void* Comm;
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(Comm*, int, UniqueId, int);
typedef int (*AllReduce_t)(const void*, void*, size_t, int /*dtype*/, int /*op*/, Comm, cudaStream_t);
typedef int (*CommDestroy_t)(Comm);
typedef const char* (*GetErrorString_t)(int);
constexpr int kFloat64 = 8;  // ncclDouble
constexpr int kSum = 0;      // ncclSum
struct Api {
  void* handle = nullptr;
  GetUniqueId_t GetUniqueId = nullptr;
  CommInitRank_t CommInitRank = nullptr;
  AllReduce_t AllReduce = nullptr;
  CommDestroy_t CommDestroy = nullptr;
  GetErrorString_t GetErrorString = nullptr;
};
Api* api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = std::getenv("GSFM_RA_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (a.handle) {
      a.GetUniqueId = (GetUniqueId_t)dlsym(a.handle, "ncclGetUniqueId");
      a.CommInitRank = (CommInitRank_t)dlsym(a.handle, "ncclCommInitRank");
      a.AllReduce = (AllReduce_t)dlsym(a.handle, "ncclAllReduce");
      a.CommDestroy = (CommDestroy_t)dlsym(a.handle, "ncclCommDestroy");
      a.GetErrorString = (GetErrorString_t)dlsym(a.handle, "ncclGetErrorString");
      if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) a.handle = nullptr;
    }
  }
  return a.handle ? &a : nullptr;
}
}  // namespace ncclx

#define NCCL_TRY(expr)                                                                                     \
  do {                                                                                                     \
    int r__ = (expr);                                                                                      \
    if (r__ != 0) {                                                                                        \
      set_error("%s failed: %s", #expr, ncclx::api()->GetErrorString ? ncclx::api()->GetErrorString(r__) : "?"); \
      return GSFM_RA_ERR_CUDA;                                                                             \
    }                                                                                                      \
  } while (0)

// A balanced work partition of the half-edge array for one kernel class: num_warps equal
// contiguous ranges (a multiple of 32 half-edges each), every range cut at row boundaries into
// segments.  num_warps = SMs x resident blocks x warps per block of THAT kernel, so the kernel runs
// as exactly one full wave.
struct Partition {
  uint32_t num_warps = 0, num_segs = 0, grid = 0, span = 0;  // span = half-edges per warp (multiple of 32)
  DevBuf<uint32_t> warp_seg_ptr, seg_row, seg_begin, seg_len, node_seg_ptr;
};

struct gsfm_ra_solver {
  StreamHolder stream_holder;  // first member: destroyed last
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t N = 0;
  uint64_t E = 0;       // edges of this shard
  uint64_t edge_begin = 0;
  uint64_t H = 0;       // half-edges of this shard (2E)
  int rank = 0, world = 1;
  int error_type = 4;
  int blk = 6;          // doubles per stored off-diagonal block (6: symmetric Laplacian stencil, 9: general)
  bool scalar_u = false;  // the whitening factor is a scalar multiple of the identity (types 2, 4, 5, 7, 8)
  gsfm_ra_options opt;
  DevLoss loss;
  int64_t launches = 0;
  bool cooperative = true;  // persistent PCG kernel available
  ncclx::Comm comm = nullptr;  // edge-sharded exchange (world > 1)
  // fused exchange: one cudaMalloc'ed block per rank {double y[2][3N]; unsigned flag;}, every peer's block mapped here
  double* xchg = nullptr;
  void* peer_base[kMaxPeers] = {};
  bool peers_connected = false;

  // structure
  DevBuf<uint32_t> he_col, he_row, he_edge, iso;
  uint32_t n_iso = 0;
  Partition pk1, pk2;  // K1 (edge kernel) and K2 (SpMV / PCG) partitions
  // per half-edge constants, planar
  DevBuf<double> inrec;  // K1's input records (k_setup_halfedges)
  int ku = 1;            // whitening entries per half-edge in the input records: 1 (scalar) or 6 (upper triangle)
  // edge-order copies for the API kernels
  DevBuf<uint32_t> d_ei, d_ej;
  DevBuf<double> d_omega_ij, d_cov6, d_weight;
  // linearisation, double buffered: [cur] is the accepted point, [cur^1] the candidate
  DevBuf<double> omega[2], node_q[2], node_JL[2], val[2], ediag[2];
  // lin[b] = [Hd 6N | gt 3N | cost, bad]: everything one evaluation sums over edges, contiguous so the
  // edge-sharded solver reduces it across GPUs with ONE all-reduce
  DevBuf<double> lin[2];
  double* Hd_p[2] = {nullptr, nullptr};
  double* gt_p[2] = {nullptr, nullptr};
  DevBuf<double> part;
  int cur = 0;
  // PCG
  DevBuf<double> scale, Dblk, Minv, x, r, z, p, q, sv, bvec, y, ypart, delta, dense_A, dense_work, dense_winv;  // z, p: stride 4
  DevBuf<double> slots;
  DevBuf<unsigned long long> bar_slots;  // grid barrier + reduction slots of the persistent PCG kernel (see grid_bar_sum2)
  bool fused_step = false;               // set while a trust-region batch is enqueued: PCG kernel runs prologue + epilogue
  uint32_t keep8 = 0;                    // L2 residency of the matrix stream, see l2_policy_evict_last
  DevBuf<unsigned> counter, row_cnt;
  DevBuf<DevScalars> sc;
  DevScalars* h_sc = nullptr;  // pinned
  HostMailbox* mailbox = nullptr;      // pinned + mapped; device alias below
  HostMailbox* mailbox_dev = nullptr;
  unsigned mailbox_seq = 0;
  // One trust-region batch as a CUDA graph (one per linearisation buffer): replayed with a single launch, its kernels
  // run back to back instead of waiting for the host to enqueue them one by one.  What varies between replays lives in
  // it_params (refreshed by the graph's first node from h_it_params).
  DevBuf<IterParams> it_params;
  IterParams* h_it_params = nullptr;   // pinned
  cudaGraphExec_t batch_graph[2] = {nullptr, nullptr};
  int64_t batch_launches = 0;
  bool graph_params = false;           // true while a batch is being captured: kernels read mu / seq from it_params
  int graph_state = 0;                 // 0 untried, 1 in use, -1 unavailable (direct launches)
  unsigned long long* prof_buf = nullptr;  // device, set only by gsfm_ra_solver_time_kernels

  // trust-region state (host)
  bool linearized = false;
  bool scale_ready = false;
  double radius = 1e4, decrease_factor = 2.0, x_cost = 0.0, x_norm = 0.0, gmax = 0.0;
  int iteration = 0, invalid_steps = 0;
  bool last_successful = false;
  int termination = GSFM_RA_TERM_NONE;
  double initial_cost = 0.0;
  // timing accumulators (ms, CUDA events)
  double ms_setup = 0, ms_assemble = 0, ms_linear = 0, ms_cost = 0;

  ~gsfm_ra_solver() {
    if (comm && ncclx::api()) ncclx::api()->CommDestroy(comm);
    for (int r = 0; r < kMaxPeers; ++r) if (peer_base[r] && r != rank) cudaIpcCloseMemHandle(peer_base[r]);
    if (xchg) cudaFree(xchg);
    if (h_sc) cudaFreeHost(h_sc);
    if (mailbox) cudaFreeHost(mailbox);
    if (h_it_params) cudaFreeHost(h_it_params);
    for (auto& g : batch_graph) if (g) cudaGraphExecDestroy(g);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    // the stream itself is destroyed by stream_holder after the buffers have been returned to the pool
  }

  bool sharded() const { return world > 1; }
  // QUATERNION_COSINE: parameters live on the manifold (left-multiplicative update, local coordinates delta = phi/2)
  bool manifold() const { return error_type <= GSFM_RA_QUATERNION_COSINE; }
  bool general() const { return error_type < GSFM_RA_QUATERNION_COSINE; }  // two-block residuals, 9-double records
  int allreduce(double* buf, size_t count) {
    if (!comm) { set_error("sharded solver used before gsfm_ra_solver_comm_init"); return GSFM_RA_ERR_INVALID; }
    NCCL_TRY(ncclx::api()->AllReduce(buf, buf, count, ncclx::kFloat64, ncclx::kSum, comm, stream));
    return 0;
  }

  int fetch_scalars() {
    CUDA_TRY(cudaMemcpyAsync(h_sc, sc.p, sizeof(DevScalars), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
  }

  // Wait for the mailbox of the batch published as `seq` (see HostMailbox) and take its scalars.  Polls host memory;
  // cudaStreamQuery every few thousand polls catches a failed launch, and a finished stream without the flag is an error.
  int fetch_mailbox(unsigned seq) {
    unsigned spins = 0;
    while (mailbox->seq != seq) {
      if ((++spins & 0xfffu) == 0) {
        const cudaError_t q = cudaStreamQuery(stream);
        if (q == cudaSuccess) {
          if (mailbox->seq == seq) break;
          set_error("trust-region batch finished without publishing its scalars");
          return GSFM_RA_ERR_CUDA;
        }
        if (q != cudaErrorNotReady) { set_error("CUDA failure while waiting for a trust-region batch: %s", cudaGetErrorString(q)); return GSFM_RA_ERR_CUDA; }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    std::memcpy(h_sc, const_cast<DevScalars*>(&mailbox->sc), sizeof(DevScalars));
    return 0;
  }
  // kernel times of the batch whose scalars were just fetched (%globaltimer stamps written by its kernels)
  void account_batch() {
    if (h_sc->t_linear_end >= h_sc->t_begin) ms_linear += 1e-6 * (double)(h_sc->t_linear_end - h_sc->t_begin);
    if (h_sc->t_end >= h_sc->t_linear_end) ms_assemble += 1e-6 * (double)(h_sc->t_end - h_sc->t_linear_end);
  }

  int rec_doubles() const { return blk * 32 + 16; }
  int smem_bytes() const { return spmv_smem_bytes(blk); }
  // K1 is specialised on (Jacobian?, residual kind, scalar weight?, loss): the common losses get their own instantiation
  // (no switch, fewer registers), everything else runs the generic one.
  typedef void (*K1Fn)(const K1Args);
  template <bool JAC, int RES, bool SCAL>
  K1Fn pick_k1_loss() const {
    const bool plain = loss.scale == 1.0;
    if (plain && loss.kind == kLossCauchy) return k_edges<JAC, RES, SCAL, kLossCauchy>;
    if (plain && loss.kind == kLossSoftLOne) return k_edges<JAC, RES, SCAL, kLossSoftLOne>;
    if (plain && loss.kind == kLossHuber) return k_edges<JAC, RES, SCAL, kLossHuber>;
    if (plain && loss.kind == kLossMagsac3) return k_edges<JAC, RES, SCAL, kLossMagsac3>;
    return k_edges<JAC, RES, SCAL, -1>;
  }
  K1Fn pick_k1(bool jacobian) const {
    if (manifold()) return jacobian ? pick_k1_loss<true, 1, true>() : pick_k1_loss<false, 1, true>();
    if (scalar_u) return jacobian ? pick_k1_loss<true, 0, true>() : pick_k1_loss<false, 0, true>();
    return jacobian ? pick_k1_loss<true, 0, false>() : pick_k1_loss<false, 0, false>();
  }
  int k1_smem_bytes() const { return kWarpsPerBlock * kStages * in_rec_doubles(ku) * 8 + kWarpsPerBlock * kStages * 8; }
  template <bool JAC, int TYPE>
  void launch_edges_g(int b, double* val_out) {
    k_edges_general<JAC, TYPE><<<pk1.grid, kBlock, 0, stream>>>(pk1.num_warps, H, pk1.warp_seg_ptr.p, pk1.seg_row.p, pk1.seg_begin.p, pk1.seg_len.p,
                                                                he_col.p, inrec.p, node_q[b].p, loss, val_out, part.p);
  }
  void launch_edges(int b, bool jacobian, double* val_out) {
    if (error_type == GSFM_RA_QUATERNION_NORM) { if (jacobian) launch_edges_g<true, 0>(b, val_out); else launch_edges_g<false, 0>(b, nullptr); return; }
    if (error_type == GSFM_RA_ROTATION_MAT_FNORM) { if (jacobian) launch_edges_g<true, 1>(b, val_out); else launch_edges_g<false, 1>(b, nullptr); return; }
    K1Args A;
    A.num_warps = pk1.num_warps; A.warp_span = pk1.span; A.H = H;
    A.warp_seg_ptr = pk1.warp_seg_ptr.p; A.seg_begin = pk1.seg_begin.p; A.seg_len = pk1.seg_len.p;
    A.inrec = inrec.p; A.node_q = node_q[b].p; A.val = val_out; A.part = part.p; A.loss = loss;
    pick_k1(jacobian)<<<pk1.grid, kBlock, k1_smem_bytes(), stream>>>(A);
  }
  void launch_spmv(int b, const double* x4, int check_done) {
    if (blk == 6)
      k_spmv<6><<<pk2.grid, kBlock, smem_bytes(), stream>>>(pk2.num_warps, H, pk2.span, pk2.warp_seg_ptr.p, pk2.seg_begin.p, pk2.seg_len.p, val[b].p,
                                                            x4, ypart.p, sc.p, check_done, keep8);
    else
      k_spmv<9><<<pk2.grid, kBlock, smem_bytes(), stream>>>(pk2.num_warps, H, pk2.span, pk2.warp_seg_ptr.p, pk2.seg_begin.p, pk2.seg_len.p, val[b].p,
                                                            x4, ypart.p, sc.p, check_done, keep8);
  }

  // ---- evaluation at omega[b]: node prep, K1 (or K1c), node finalize -------------------------
  int evaluate(int b, bool jacobian, bool publish = false, bool prepped = false) {
    HostMailbox* mb = publish ? mailbox_dev : nullptr;
    const IterParams* ip_dev = graph_params ? it_params.p : nullptr;  // graph replay: the sequence number comes from device memory
    const unsigned mseq = (publish && !graph_params) ? ++mailbox_seq : 0u;
    if (!prepped) {
      k_node_prep<<<grid_for(N), kBlock, 0, stream>>>(N, omega[b].p, node_q[b].p, node_JL[b].p, slots.p, counter.p, sc.p, manifold() ? 1 : 0);
      launches += 1;
    }
    launch_edges(b, jacobian, val[b].p);
    double* tail = lin[b].p + 9ull * N;
    const int co = jacobian ? 0 : 1;
    if (!sharded()) {
      k_node_finalize<<<grid_for(N), kBlock, 0, stream>>>(N, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, co, 0, tail,
                                                           slots.p, counter.p, sc.p, mb, mseq, publish ? ip_dev : nullptr);
    } else {
      // edge-sharded: local sums -> ONE all-reduce of [Hd | gt | cost, bad] -> per-view post-processing
      k_node_finalize<<<grid_for(N), kBlock, 0, stream>>>(N, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, co, 1, tail,
                                                           slots.p, counter.p, sc.p, nullptr, 0u, nullptr);
      if (jacobian) RA_TRY(allreduce(lin[b].p, 9ull * N + 2));
      else RA_TRY(allreduce(tail, 2));
      k_node_finalize<<<grid_for(N), kBlock, 0, stream>>>(N, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, co, 2, tail,
                                                           slots.p, counter.p, sc.p, mb, mseq, publish ? ip_dev : nullptr);
      launches += 2;
    }
    launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }

  // y = (Ht offdiag + diag_blocks) x on linearisation b (separate-kernel path); xin has stride 4, yout stride 3.
  int spmv(int b, const double* xin, double* yout, const double* diag_blocks) {
    launch_spmv(b, xin, 0);
    if (!sharded()) {
      k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, pk2.node_seg_ptr.p, ypart.p, diag_blocks, xin, yout, nullptr, 1, slots.p, counter.p, sc.p);
    } else {
      k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, pk2.node_seg_ptr.p, ypart.p, nullptr, xin, yout, nullptr, 1, slots.p, counter.p, sc.p);
      RA_TRY(allreduce(yout, 3ull * N));
      k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, pk2.node_seg_ptr.p, ypart.p, diag_blocks, xin, yout, yout, 1, slots.p, counter.p, sc.p);
      launches += 2;
    }
    launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }

  PrepareArgs prepare_args(int b, double mu, const double* user_damp, const double* user_b) {
    PrepareArgs A;
    A.mu = mu; A.lo = opt.min_lm_diagonal; A.hi = opt.max_lm_diagonal;
    A.ediag = ediag[b].p; A.scale = scale.p; A.node_JL = node_JL[b].p; A.Hd = Hd_p[b]; A.gt = gt_p[b]; A.user_damp = user_damp; A.user_b = user_b;
    A.Dblk = Dblk.p; A.Minv = Minv.p; A.x = x.p; A.r = r.p; A.z = z.p; A.p = p.p; A.q = q.p; A.bvec = bvec.p;
    return A;
  }
  ApplyArgs apply_args(int b, int c) {
    ApplyArgs A;
    A.node_JL = node_JL[b].p; A.xt = x.p; A.bvec = bvec.p; A.res = r.p; A.Dblk = Dblk.p; A.Hd = Hd_p[b]; A.gt = gt_p[b]; A.omega = omega[b].p;
    A.cand = omega[c].p; A.delta_out = delta.p; A.manifold = manifold() ? 1 : 0;
    return A;
  }
  PcgParams pcg_params(int b, double rtol, int max_iter) {
    PcgParams P;
    P.N = N; P.num_warps = pk2.num_warps; P.n_iso = n_iso; P.max_iter = max_iter; P.H = H; P.rtol2 = rtol * rtol;
    P.keep8 = keep8;
    P.warp_seg_ptr = pk2.warp_seg_ptr.p; P.seg_row = pk2.seg_row.p; P.seg_begin = pk2.seg_begin.p; P.seg_len = pk2.seg_len.p;
    P.node_seg_ptr = pk2.node_seg_ptr.p; P.iso = iso.p; P.warp_span = pk2.span;
    P.val = val[b].p; P.Dblk = Dblk.p; P.Minv = Minv.p;
    P.x = x.p; P.r = r.p; P.z = z.p; P.p = p.p; P.q = q.p; P.s = sv.p; P.ypart = ypart.p;
    P.row_cnt = row_cnt.p; P.bar_slots = bar_slots.p; P.slots = slots.p; P.counter = counter.p; P.sc = sc.p; P.prof = prof_buf;
    P.fused = 0; P.ip = nullptr; P.cand_q = nullptr; P.cand_JL = nullptr;
    std::memset(&P.prep, 0, sizeof(P.prep)); std::memset(&P.apply, 0, sizeof(P.apply));
    P.world = peers_connected ? world : 1; P.rank = rank;
    for (int r = 0; r < kMaxPeers; ++r) {
      P.peer_y[r] = (double*)peer_base[r];
      P.peer_flag[r] = peer_base[r] ? (unsigned*)((double*)peer_base[r] + 6ull * N) : nullptr;
    }
    return P;
  }

  // Block-Jacobi PCG on (Ht + Lam) xt = bt for linearisation b; on return (stream order) x = xt, r = the
  // residual bt - (Ht + Lam) xt and bvec = bt.  One cooperative launch; no host synchronisation.
  int pcg_enqueue(int b, double mu, const double* user_damp, const double* user_b, double rtol, int max_iter) {
    const bool fuse = fused_step && !user_damp && !user_b;  // a trust-region batch: prologue and epilogue live in the PCG kernel
    if (!fuse) {
      k_prepare_solve<<<grid_for(N), kBlock, 0, stream>>>(N, prepare_args(b, mu, user_damp, user_b), slots.p, counter.p, sc.p,
                                                           graph_params ? it_params.p : nullptr);
      launches += 1;
    }
    if (opt.linear_solver == GSFM_RA_SOLVER_DENSE_CHOLESKY) {
      const uint32_t n = 3 * N, np = (n + 1 + kNB - 1) / kNB * kNB;  // room for the right-hand-side row
      if (dense_A.n < (size_t)np * np) {
        AllocScope scope(stream);
        RA_TRY(dense_A.alloc((size_t)np * np));
        RA_TRY(dense_work.alloc(np));
        RA_TRY(dense_winv.alloc((size_t)np * kNB));
      }
      CUDA_TRY(cudaMemsetAsync(dense_A.p, 0, (size_t)np * np * sizeof(double), stream));
      k_dense_assemble<<<grid_for(std::max<uint64_t>(H, np)), kBlock, 0, stream>>>(H, N, np, blk, he_row.p, he_col.p, val[b].p, Dblk.p, r.p, dense_A.p);
      uint32_t n_arg = n, np_arg = np;
      double* A_arg = dense_A.p; double* x_arg = x.p; double* wi_arg = dense_winv.p; double* w_arg = dense_work.p; DevScalars* sc_arg = sc.p;
      void* args[] = {&n_arg, &np_arg, &A_arg, &x_arg, &wi_arg, &w_arg, &sc_arg};
      // all SMs: the trailing update of panel kb has (nblk-kb)(nblk-kb-1)/2 tiles to spread
      int occ = 1;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dense_cholesky_solve, kBlock, 0));
      const unsigned grid = (unsigned)(sm_count * std::max(1, std::min(occ, 2)));
      CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_dense_cholesky_solve, dim3(grid), dim3(kBlock), args, 0, stream));
      CUDA_TRY(cudaMemsetAsync(r.p, 0, 3ull * N * sizeof(double), stream));  // exact solve: residual 0
      launches += 2;
      return 0;
    }
    if (cooperative && (!sharded() || peers_connected)) {
      PcgParams P = pcg_params(b, rtol, max_iter);
      P.fused = fuse ? 1 : 0;
      if (fuse) {
        P.prep = prepare_args(b, mu, nullptr, nullptr);
        P.apply = apply_args(b, b ^ 1);
        P.cand_q = node_q[b ^ 1].p; P.cand_JL = node_JL[b ^ 1].p;
        P.ip = graph_params ? it_params.p : nullptr;
      }
      void* args[] = {&P};
      const void* fn = (blk == 6) ? (const void*)k_pcg_persistent<6> : (const void*)k_pcg_persistent<9>;
      CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(pk2.grid), dim3(kBlock), args, smem_bytes(), stream));
      launches += 1;
      return 0;
    }
    // fallback (device without cooperative launch, or NCCL exchange): textbook PCG from separate kernels, polled
    const int poll = 8;
    const double rtol2 = rtol * rtol;
    int enq = 0;
    while (true) {
      for (int k = 0; k < poll && enq < max_iter; ++k, ++enq) {
        launch_spmv(b, p.p, 1);
        if (!sharded()) {
          k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, pk2.node_seg_ptr.p, ypart.p, Dblk.p, p.p, y.p, nullptr, 0, slots.p, counter.p, sc.p);
        } else {
          // the ONE collective of a CG step: all-reduce of the 3N partial matvec
          k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, pk2.node_seg_ptr.p, ypart.p, nullptr, p.p, y.p, nullptr, 1, slots.p, counter.p, sc.p);
          RA_TRY(allreduce(y.p, 3ull * N));
          k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, pk2.node_seg_ptr.p, ypart.p, Dblk.p, p.p, y.p, y.p, 0, slots.p, counter.p, sc.p);
          launches += 2;
        }
        k_pcg_update<<<grid_for(N), kBlock, 0, stream>>>(N, Minv.p, p.p, y.p, x.p, r.p, z.p, rtol2, max_iter, slots.p, counter.p, sc.p);
        k_pcg_direction<<<grid_for(4ull * N), kBlock, 0, stream>>>(4 * N, z.p, p.p, sc.p);
        launches += 4;
      }
      CUDA_TRY(cudaGetLastError());
      RA_TRY(fetch_scalars());
      if (h_sc->pcg_done || enq >= max_iter) break;
    }
    return 0;
  }

  double elapsed(cudaEvent_t a, cudaEvent_t b2) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b2);
    return ms;
  }
  double elapsed_since(cudaEvent_t a) {
    cudaEventRecord(ev[1], stream);
    cudaEventSynchronize(ev[1]);
    return elapsed(a, ev[1]);
  }
};

namespace {

int device_info(int device, DeviceInfo** out) {
  DeviceInfo& d = g_device_info[device & 63];
  if (!d.ready) {
    CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, device));
    CUDA_TRY(cudaDeviceGetAttribute(&d.l2_bytes, cudaDevAttrL2CacheSize, device));
    CUDA_TRY(cudaFuncSetAttribute(k_pcg_persistent<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(6)));
    CUDA_TRY(cudaFuncSetAttribute(k_pcg_persistent<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(9)));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(6)));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(9)));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_k2[0], k_pcg_persistent<6>, kBlock, spmv_smem_bytes(6)));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_k2[1], k_pcg_persistent<9>, kBlock, spmv_smem_bytes(9)));
    // keep freed blocks in the pool: the next solver reuses them
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    d.ready = true;
  }
  *out = &d;
  return 0;
}

template <typename T>
int exclusive_scan(const T* in, T* out, size_t n, cudaStream_t st) {
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
  DevBuf<unsigned char> tmp;
  RA_TRY(tmp.alloc(bytes + 16));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, st));
  return 0;
}

// Upload the (shard of the) problem and build every device structure; one host sync at the end.
int build_solver(const gsfm_ra_problem* prob, const gsfm_ra_options* options, int rank, int world, gsfm_ra_solver** out) {
  RA_TRY(check_problem(prob));
  if (!options || !out) { set_error("options/out is NULL"); return GSFM_RA_ERR_INVALID; }
  if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank/world"); return GSFM_RA_ERR_INVALID; }
  const uint32_t N = prob->num_views;
  const double tb0 = now_ms();
  double tb_prev = tb0;
  auto lap = [&](const char* what) {
    if (options->verbose >= 2) { const double t = now_ms(); std::fprintf(stderr, "[gsfm_ra] setup %-28s %8.2f ms\n", what, t - tb_prev); tb_prev = t; }
  };
  int device = 0;
  RA_TRY(select_device(options->device, &device));
  DeviceInfo* di = nullptr;
  RA_TRY(device_info(device, &di));
  std::unique_ptr<gsfm_ra_solver> s(new gsfm_ra_solver());
  s->device = device;
  s->opt = *options;
  s->rank = rank; s->world = world;
  s->error_type = prob->error_type;
  s->blk = s->general() ? 9 : 6;
  s->scalar_u = !(prob->error_type == GSFM_RA_ANGLE_AXIS_COVARIANCE || prob->error_type == GSFM_RA_ANGLE_AXIS_COV_INLIERS);
  s->ku = s->scalar_u ? 1 : 6;
  RA_TRY(make_dev_loss(&options->loss, &s->loss));
  s->sm_count = di->sm_count;
  CUDA_TRY(cudaStreamCreateWithFlags(&s->stream_holder.s, cudaStreamNonBlocking));
  s->stream = s->stream_holder.s;
  AllocScope alloc_scope(s->stream);
  for (auto& e : s->ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDefault));
  CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
  lap("context/stream");

  // shard: a contiguous range of the caller's edge list
  const uint64_t e0 = prob->num_edges * (uint64_t)rank / world, e1 = prob->num_edges * (uint64_t)(rank + 1) / world;
  const uint64_t E = e1 - e0, H = 2 * E;
  if (E == 0) { set_error("rank %d of %d owns no edges", rank, world); return GSFM_RA_ERR_INVALID; }
  s->N = N; s->E = E; s->H = H; s->edge_begin = e0;
  cudaStream_t st = s->stream;

  auto up32 = [&](DevBuf<uint32_t>& d, const uint32_t* src, size_t n) -> int {
    RA_TRY(d.alloc(n));
    if (n == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(d.p, src, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    return 0;
  };
  auto up64 = [&](DevBuf<double>& d, const double* src, size_t n) -> int {
    RA_TRY(d.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(d.p, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return 0;
  };
  RA_TRY(up32(s->d_ei, prob->edge_i + e0, E));
  RA_TRY(up32(s->d_ej, prob->edge_j + e0, E));
  RA_TRY(up64(s->d_omega_ij, prob->omega_ij + 3 * e0, 3 * E));
  if (prob->cov6) RA_TRY(up64(s->d_cov6, prob->cov6 + 6 * e0, 6 * E));
  if (prob->edge_weight) RA_TRY(up64(s->d_weight, prob->edge_weight + e0, E));
  lap("enqueue uploads");

  // ---- half-edges sorted by (row, col): keys -> radix sort -> unpack ----------------------------
  DevBuf<int> d_err;
  DevBuf<uint64_t> keys_a, keys_b;
  DevBuf<uint32_t> vals_a, vals_b, rowptr, flags_ne, flags_iso, nz_rank, iso_rank;
  RA_TRY(d_err.alloc(4));
  CUDA_TRY(cudaMemsetAsync(d_err.p, 0, 4 * sizeof(int), st));
  RA_TRY(keys_a.alloc(H)); RA_TRY(keys_b.alloc(H)); RA_TRY(vals_a.alloc(H)); RA_TRY(vals_b.alloc(H));
  k_build_keys<<<grid_for(E), kBlock, 0, st>>>(E, N, s->d_ei.p, s->d_ej.p, keys_a.p, vals_a.p, d_err.p);
  {
    int bits = 1;
    while (bits < 64 && ((uint64_t)N * N - 1) >> bits) ++bits;
    size_t bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_a.p, keys_b.p, vals_a.p, vals_b.p, (int)H, 0, bits, st));
    DevBuf<unsigned char> tmp;
    RA_TRY(tmp.alloc(bytes + 16));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_a.p, keys_b.p, vals_a.p, vals_b.p, (int)H, 0, bits, st));
  }
  RA_TRY(s->he_col.alloc(H)); RA_TRY(s->he_row.alloc(H)); RA_TRY(s->he_edge.alloc(H));
  k_unpack_keys<<<grid_for(H), kBlock, 0, st>>>(H, N, keys_b.p, vals_b.p, s->he_row.p, s->he_col.p, s->he_edge.p, d_err.p);
  RA_TRY(rowptr.alloc(N + 2)); RA_TRY(flags_ne.alloc(N + 2)); RA_TRY(flags_iso.alloc(N + 2)); RA_TRY(nz_rank.alloc(N + 2)); RA_TRY(iso_rank.alloc(N + 2));
  k_rowptr<<<grid_for(N + 1), kBlock, 0, st>>>(N, H, keys_b.p, rowptr.p);
  k_row_flags<<<grid_for(N + 1), kBlock, 0, st>>>(N, rowptr.p, flags_ne.p, flags_iso.p);
  RA_TRY(exclusive_scan(flags_ne.p, nz_rank.p, N + 1, st));
  RA_TRY(exclusive_scan(flags_iso.p, iso_rank.p, N + 1, st));
  RA_TRY(s->iso.alloc(N));
  k_iso_fill<<<grid_for(N), kBlock, 0, st>>>(N, flags_iso.p, iso_rank.p, s->iso.p);
  CUDA_TRY(cudaMemcpyAsync(d_err.p + 1, iso_rank.p + N, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));  // n_iso
  s->launches += 6;

  // ---- balanced partitions, one per kernel class, sized to exactly one resident wave of that kernel ----
  const int occ_k2 = di->occ_k2[s->blk == 6 ? 0 : 1];
  s->cooperative = di->coop != 0 && occ_k2 > 0;
  // Edge-sharded: the range length and the launch grid are sized from the LARGEST shard (ceil(E_total / world) edges), so
  // every rank launches the same grid even when the shards differ by an edge -- the replicated grid-wide sums then add in
  // the same order on every rank and the replicas stay bit-identical.  (A rank with fewer half-edges leaves its last
  // warps idle.)
  const uint64_t H_sizing = 2 * ((prob->num_edges + (uint64_t)world - 1) / (uint64_t)world);
  auto make = [&](Partition& P, int blocks_per_sm) -> int {
    const uint64_t max_warps = (uint64_t)s->sm_count * std::max(1, blocks_per_sm) * kWarpsPerBlock;
    uint64_t per = (H_sizing + max_warps - 1) / max_warps;
    per = std::max<uint64_t>(64, (per + 31) / 32 * 32);  // at least 64 half-edges per warp, whole records
    const uint32_t nw = (uint32_t)std::max<uint64_t>(1, (H + per - 1) / per);
    const uint32_t nw_sizing = (uint32_t)std::max<uint64_t>(1, (H_sizing + per - 1) / per);
    P.num_warps = nw; P.span = (uint32_t)per;
    P.num_segs = nw + N;  // upper bound: every range start + every row start opens one segment
    P.grid = (nw_sizing + kWarpsPerBlock - 1) / kWarpsPerBlock;
    DevBuf<uint32_t> nseg, cnt;
    RA_TRY(nseg.alloc(nw + 2)); RA_TRY(cnt.alloc(N + 2));
    RA_TRY(P.warp_seg_ptr.alloc(nw + 2)); RA_TRY(P.seg_row.alloc(P.num_segs)); RA_TRY(P.seg_begin.alloc(P.num_segs));
    RA_TRY(P.seg_len.alloc(P.num_segs)); RA_TRY(P.node_seg_ptr.alloc(N + 2));
    k_part_count<<<grid_for(nw + 1), kBlock, 0, st>>>(nw, (uint32_t)per, H, s->he_row.p, nz_rank.p, nseg.p);
    RA_TRY(exclusive_scan(nseg.p, P.warp_seg_ptr.p, nw + 1, st));
    k_part_fill<<<grid_for(nw), kBlock, 0, st>>>(nw, (uint32_t)per, H, s->he_row.p, rowptr.p, P.warp_seg_ptr.p, P.seg_row.p, P.seg_begin.p, P.seg_len.p);
    k_node_seg_count<<<grid_for(N + 1), kBlock, 0, st>>>(N, (uint32_t)per, rowptr.p, cnt.p);
    RA_TRY(exclusive_scan(cnt.p, P.node_seg_ptr.p, N + 1, st));
    s->launches += 5;
    return 0;
  };
  int occ_k1 = 1;
  if (!s->general()) {
    // the two K1 instantiations this solver will launch (with / without Jacobian): opt in to their shared memory, size the
    // partition for the occupancy of the heavier one
    for (int jac = 0; jac < 2; ++jac) CUDA_TRY(cudaFuncSetAttribute((const void*)s->pick_k1(jac != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, s->k1_smem_bytes()));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_k1, (const void*)s->pick_k1(true), kBlock, s->k1_smem_bytes()));
  }
  RA_TRY(make(s->pk1, occ_k1));
  RA_TRY(make(s->pk2, occ_k2));
  // the cooperative grid must be fully resident; node loops are grid-strided so any size works
  s->pk2.grid = std::min<uint32_t>(s->pk2.grid, (uint32_t)s->sm_count * std::max(1, occ_k2));
  {
    // L2 residency of the matrix stream: pin what 62 % of the L2 can hold when the stream is larger than that but still
    // comparable to the L2 (GSFM_RA_L2_KEEP=0..7 overrides, 0 disables)
    const double stream = (double)((H + 31) / 32) * s->rec_doubles() * 8.0, budget = 0.62 * (double)di->l2_bytes;
    uint32_t k = 0;
    if (stream > budget) k = (uint32_t)std::min(7.0, std::floor(8.0 * budget / stream));
    if (const char* e = std::getenv("GSFM_RA_L2_KEEP")) k = (uint32_t)std::min(7, std::max(0, std::atoi(e)));
    s->keep8 = k;
  }
  lap("enqueue structure build");

  // ---- per half-edge constants (K0) and the solver's working set ---------------------------------
  {
    const size_t nrec_in = (size_t)((H + 31) / 32), rd = (size_t)in_rec_doubles(s->ku);
    RA_TRY(s->inrec.alloc(nrec_in * rd));
    CUDA_TRY(cudaMemsetAsync(s->inrec.p + (nrec_in - 1) * rd, 0, rd * 8, st));  // lanes past H in the last record
  }
  k_setup_halfedges<<<grid_for(H), kBlock, 0, st>>>(H, s->ku, s->he_edge.p, s->he_row.p, s->he_col.p, s->d_omega_ij.p, s->d_cov6.p, s->d_weight.p,
                                                    prob->error_type, s->inrec.p);
  s->launches += 1;
  const size_t nrec = (size_t)((H + 31) / 32);
  for (int b = 0; b < 2; ++b) {
    RA_TRY(s->omega[b].alloc(3ull * N));
    RA_TRY(s->node_q[b].alloc(4ull * N));
    RA_TRY(s->node_JL[b].alloc(9ull * N));
    RA_TRY(s->val[b].alloc(nrec * s->rec_doubles()));
    // only the last (partial) record has lanes K1 never writes
    CUDA_TRY(cudaMemsetAsync(s->val[b].p + (nrec - 1) * s->rec_doubles(), 0, (size_t)s->rec_doubles() * 8, st));
    RA_TRY(s->lin[b].alloc(9ull * N + 2));
    s->Hd_p[b] = s->lin[b].p;
    s->gt_p[b] = s->lin[b].p + 6ull * N;
    RA_TRY(s->ediag[b].alloc(3ull * N));
    CUDA_TRY(cudaMemsetAsync(s->omega[b].p, 0, 3ull * N * sizeof(double), st));
  }
  k_embed_cols<<<grid_for(H), kBlock, 0, st>>>(H, s->blk, s->he_col.p, s->val[0].p, s->val[1].p);
  s->launches += 1;
  RA_TRY(s->part.alloc((size_t)s->pk1.num_segs * kPartStride));
  RA_TRY(s->ypart.alloc((size_t)s->pk2.num_segs * 3));
  for (DevBuf<double>* d : {&s->scale, &s->x, &s->r, &s->q, &s->sv, &s->bvec, &s->y, &s->delta}) RA_TRY(d->alloc(3ull * N));
  RA_TRY(s->z.alloc(4ull * N));
  RA_TRY(s->p.alloc(4ull * N));
  RA_TRY(s->bar_slots.alloc(2 * 4 * (size_t)s->pk2.grid + 8));
  CUDA_TRY(cudaMemsetAsync(s->bar_slots.p, 0, (2 * 4 * (size_t)s->pk2.grid + 8) * sizeof(unsigned long long), st));
  RA_TRY(s->row_cnt.alloc(N));
  CUDA_TRY(cudaMemsetAsync(s->row_cnt.p, 0, (size_t)N * sizeof(unsigned), st));
  RA_TRY(s->Dblk.alloc(6ull * N));
  RA_TRY(s->Minv.alloc(6ull * N));
  RA_TRY(s->slots.alloc((size_t)std::max<unsigned>(std::max(grid_for(3ull * N), grid_for(E)), 64) * 4 + 64));
  RA_TRY(s->counter.alloc(4));
  RA_TRY(s->sc.alloc(1));
  CUDA_TRY(cudaMemsetAsync(s->counter.p, 0, 4 * sizeof(unsigned), st));
  CUDA_TRY(cudaMemsetAsync(s->sc.p, 0, sizeof(DevScalars), st));
  CUDA_TRY(cudaMallocHost(&s->h_sc, sizeof(DevScalars)));
  CUDA_TRY(cudaHostAlloc(&s->mailbox, sizeof(HostMailbox), cudaHostAllocMapped));
  std::memset(s->mailbox, 0, sizeof(HostMailbox));
  CUDA_TRY(cudaHostGetDevicePointer(&s->mailbox_dev, s->mailbox, 0));
  CUDA_TRY(cudaMallocHost(&s->h_it_params, sizeof(IterParams)));
  RA_TRY(s->it_params.alloc(1));
  k_jacobi_scale<<<grid_for(3ull * N), kBlock, 0, st>>>(3 * N, s->ediag[0].p, s->scale.p, 0);
  s->launches += 1;
  CUDA_TRY(cudaGetLastError());
  lap("enqueue K0 + allocations");
  // the only synchronisation of the build: input checks and the isolated-view count
  int h_err[4] = {0, 0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(h_err, d_err.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  lap("wait for the device");
  if (h_err[0] == 1) { set_error("an edge is out of range or a self loop"); return GSFM_RA_ERR_INVALID; }
  if (h_err[0] == 2) { set_error("duplicate edge: the same view pair appears twice"); return GSFM_RA_ERR_INVALID; }
  s->n_iso = (uint32_t)h_err[1];
  if (options->verbose >= 2) std::fprintf(stderr, "[gsfm_ra] setup total %.2f ms\n", now_ms() - tb0);
  s->ms_setup = s->elapsed_since(s->ev[0]);
  s->radius = options->initial_trust_region_radius;
  *out = s.release();
  return 0;
}

void reset_trust_region(gsfm_ra_solver* s) {
  s->linearized = false;
  s->scale_ready = false;
  s->radius = s->opt.initial_trust_region_radius;
  s->decrease_factor = 2.0;
  s->iteration = 0;
  s->invalid_steps = 0;
  s->last_successful = false;
  s->termination = GSFM_RA_TERM_NONE;
}

void push_trace(gsfm_ra_summary* sum, const gsfm_ra_iteration& it) {
  if (sum && sum->trace && sum->trace_size < sum->trace_capacity) sum->trace[sum->trace_size++] = it;
}

// One trust-region batch at linearisation buffer b: damping + PCG initialisation, the whole PCG solve, the step and the
// candidate point, and the speculative linearisation of the candidate (node prep, K1, node finalize + host mailbox).
int enqueue_batch(gsfm_ra_solver* s, int b, double mu) {
  const int c = b ^ 1;
  const uint32_t N = s->N;
  // single GPU / fused exchange with the persistent PCG kernel: damping + PCG initialisation run as the kernel's prologue,
  // step + candidate + its per-view preparation as its epilogue -> the batch is PCG, K1, node finalize
  const bool fuse = s->opt.linear_solver == GSFM_RA_SOLVER_PCG && s->cooperative && (!s->sharded() || s->peers_connected) && !std::getenv("GSFM_RA_NO_FUSE");
  if (!fuse) CUDA_TRY(cudaMemsetAsync(&s->sc.p->bad, 0, sizeof(int), s->stream));
  s->fused_step = fuse;
  const int rc = s->pcg_enqueue(b, mu, nullptr, nullptr, s->opt.pcg_rtol, s->opt.pcg_max_iterations);
  s->fused_step = false;
  RA_TRY(rc);
  if (!fuse) {
    k_apply_step<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->apply_args(b, c), s->slots.p, s->counter.p, s->sc.p);
    s->launches += 1;
  }
  RA_TRY(s->evaluate(c, true, true, fuse));
  return 0;
}

// Run the batch: as a CUDA graph replay where possible (single GPU, persistent PCG kernel) -- one launch, the kernels run
// back to back -- else as direct launches.  GSFM_RA_NO_GRAPH=1 forces direct launches.
int run_batch(gsfm_ra_solver* s, int b) {
  if (s->graph_state == 0) {
    const bool eligible = s->opt.linear_solver == GSFM_RA_SOLVER_PCG && s->cooperative && !s->sharded() && !std::getenv("GSFM_RA_NO_GRAPH");
    s->graph_state = eligible ? 1 : -1;
  }
  if (s->graph_state == 1 && !s->batch_graph[b]) {
    // capture this buffer's batch once
    const int64_t l0 = s->launches;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      s->graph_params = true;
      int rc = (cudaMemcpyAsync(s->it_params.p, s->h_it_params, sizeof(IterParams), cudaMemcpyHostToDevice, s->stream) == cudaSuccess) ? 0 : 1;
      if (rc == 0) rc = enqueue_batch(s, b, 0.0);
      s->graph_params = false;
      const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
      ok = rc == 0 && e == cudaSuccess && graph != nullptr;
      if (ok) ok = cudaGraphInstantiate(&s->batch_graph[b], graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
    }
    s->batch_launches = s->launches - l0;
    s->launches = l0;
    if (!ok) {
      (void)cudaGetLastError();
      s->batch_graph[b] = nullptr;
      s->graph_state = -1;
      if (s->opt.verbose) std::fprintf(stderr, "[gsfm_ra] CUDA graph capture of the trust-region batch failed; using direct launches\n");
    }
  }
  if (s->graph_state == 1) {
    s->h_it_params->mu = s->radius;
    s->h_it_params->seq = ++s->mailbox_seq;
    CUDA_TRY(cudaGraphLaunch(s->batch_graph[b], s->stream));
    s->launches += s->batch_launches;
    return 0;
  }
  return enqueue_batch(s, b, s->radius);
}

// Ceres-1.14 trust-region loop, same order of checks as oracle/ra_oracle.cc.
// A candidate point is linearised speculatively (full K1 into the other buffer): an accepted step
// then needs no second evaluation, a rejected one simply keeps the current buffer.
int iterate(gsfm_ra_solver* s, int max_new_iterations, gsfm_ra_summary* sum) {
  const gsfm_ra_options& o = s->opt;
  const double t_start = now_ms();
  const int64_t launches0 = s->launches;
  const double asm0 = s->ms_assemble, lin0 = s->ms_linear, cost0 = s->ms_cost;
  const uint32_t N = s->N;
  int succ = 0, unsucc = 0;
  int64_t lin_total = 0;
  if (sum) sum->trace_size = 0;
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->linearized) {
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    CUDA_TRY(cudaMemsetAsync(&s->sc.p->bad, 0, sizeof(int), s->stream));
    RA_TRY(s->evaluate(s->cur, true));
    RA_TRY(s->fetch_scalars());
    s->ms_assemble += s->elapsed_since(s->ev[0]);
    if (s->h_sc->bad) { set_error("non-finite cost or Jacobian at the initial point"); s->termination = GSFM_RA_TERM_FAILURE; return GSFM_RA_ERR_NUMERIC; }
    s->x_cost = s->h_sc->cost; s->gmax = s->h_sc->gmax; s->x_norm = std::sqrt(s->h_sc->xnorm2);
    s->initial_cost = s->x_cost;
    s->linearized = true;
    if (!s->scale_ready) {
      k_jacobi_scale<<<grid_for(3ull * N), kBlock, 0, s->stream>>>(3 * N, s->ediag[s->cur].p, s->scale.p, o.jacobi_scaling);
      s->launches += 1;
      s->scale_ready = true;
    }
    gsfm_ra_iteration it0;
    std::memset(&it0, 0, sizeof(it0));
    it0.cost = s->x_cost; it0.gradient_max_norm = s->gmax; it0.trust_region_radius = s->radius;
    push_trace(sum, it0);
    if (s->gmax <= o.gradient_tolerance) s->termination = GSFM_RA_TERM_GRADIENT_TOLERANCE;
  }
  int done_here = 0;
  while (s->termination == GSFM_RA_TERM_NONE && done_here < max_new_iterations) {
    if (s->iteration >= o.max_num_iterations) { s->termination = GSFM_RA_TERM_MAX_ITERATIONS; break; }
    if (s->last_successful && s->gmax <= o.gradient_tolerance) { s->termination = GSFM_RA_TERM_GRADIENT_TOLERANCE; break; }
    if (s->radius <= o.min_trust_region_radius) { s->termination = GSFM_RA_TERM_MIN_RADIUS; break; }
    ++s->iteration;
    ++done_here;
    gsfm_ra_iteration it;
    std::memset(&it, 0, sizeof(it));
    it.iteration = s->iteration;
    const int b = s->cur, c = s->cur ^ 1;
    // ---- one trust-region iteration = one stream-ordered batch, ONE host synchronisation -------
    //   k_prepare_solve -> k_pcg_persistent (whole PCG + Ht x) -> k_apply_step (step, candidate)
    //   -> speculative linearisation of the candidate (k_node_prep, K1, k_node_finalize)
    RA_TRY(run_batch(s, b));
    RA_TRY(s->fetch_mailbox(s->mailbox_seq));
    s->account_batch();
    if (s->h_sc->bad == 2) { set_error("multi-GPU exchange timed out: a peer rank did not reach the same CG step"); s->termination = GSFM_RA_TERM_FAILURE; return GSFM_RA_ERR_CUDA; }
    const bool breakdown = s->h_sc->pcg_breakdown != 0;
    const int lin_it = s->h_sc->pcg_iter;
    const double lin_res = (s->h_sc->bb > 0.0) ? std::sqrt(s->h_sc->rr / s->h_sc->bb) : 0.0;
    it.linear_iterations = lin_it; it.linear_residual = lin_res;
    lin_total += lin_it;
    const double model_change = -s->h_sc->dg - 0.5 * s->h_sc->dHd;
    it.model_cost_change = model_change;
    // `bad` also covers a non-finite candidate evaluation; a non-finite STEP shows up in step2
    const bool step_finite = std::isfinite(s->h_sc->step2) && std::isfinite(model_change);
    bool valid = !breakdown && step_finite && model_change > 0.0;
    it.step_is_valid = valid;
    if (!valid) {
      it.cost = s->x_cost; it.gradient_max_norm = s->gmax;
      if (++s->invalid_steps >= 5) { s->termination = GSFM_RA_TERM_INVALID_STEPS; it.trust_region_radius = s->radius; push_trace(sum, it); break; }
      s->radius *= 0.5;
      it.trust_region_radius = s->radius;
      s->last_successful = false;
      ++unsucc;
      push_trace(sum, it);
      continue;
    }
    s->invalid_steps = 0;
    double cand_cost = s->h_sc->cost;
    const bool cand_bad = s->h_sc->bad != 0 || !std::isfinite(cand_cost);
    if (cand_bad) cand_cost = DBL_MAX;
    it.candidate_cost = cand_cost;
    it.step_norm = std::sqrt(s->h_sc->step2);
    it.cost_change = s->x_cost - cand_cost;
    it.relative_decrease = it.cost_change / model_change;
    it.gradient_max_norm = s->gmax;
    if (it.step_norm <= o.parameter_tolerance * (s->x_norm + o.parameter_tolerance)) {
      s->termination = GSFM_RA_TERM_PARAMETER_TOLERANCE; it.cost = s->x_cost; it.trust_region_radius = s->radius; push_trace(sum, it); break;
    }
    if (std::fabs(it.cost_change) <= o.function_tolerance * s->x_cost) {
      s->termination = GSFM_RA_TERM_FUNCTION_TOLERANCE; it.cost = s->x_cost; it.trust_region_radius = s->radius; push_trace(sum, it); break;
    }
    if (it.relative_decrease > o.min_relative_decrease && !cand_bad) {
      s->cur = c;
      s->x_cost = cand_cost;
      s->gmax = s->h_sc->gmax;
      s->x_norm = std::sqrt(s->h_sc->xnorm2);
      it.gradient_max_norm = s->gmax;
      it.step_is_successful = 1;
      const double q = it.relative_decrease;
      s->radius = s->radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * q - 1.0, 3));
      s->radius = std::min(o.max_trust_region_radius, s->radius);
      s->decrease_factor = 2.0;
      s->last_successful = true;
      ++succ;
      it.cost = s->x_cost;
    } else {
      s->radius = s->radius / s->decrease_factor;
      s->decrease_factor *= 2.0;
      s->last_successful = false;
      ++unsucc;
      it.cost = cand_cost;
    }
    it.trust_region_radius = s->radius;
    push_trace(sum, it);
    if (o.verbose)
      std::fprintf(stderr, "[gsfm_ra] it %3d cost %.12e dcost %+.3e |g| %.3e |step| %.3e q %.3e radius %.3e pcg %d (%.1e) %s\n", s->iteration, s->x_cost,
                   it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, s->radius, lin_it, lin_res, it.step_is_successful ? "ok" : "rejected");
  }
  if (sum) {
    sum->termination = s->termination;
    sum->num_iterations = s->iteration;
    sum->num_successful_steps = succ;
    sum->num_unsuccessful_steps = unsucc;
    sum->total_linear_iterations = lin_total;
    sum->initial_cost = s->initial_cost;
    sum->final_cost = s->x_cost;
    sum->ms_setup = s->ms_setup;
    sum->ms_assemble = s->ms_assemble - asm0;
    sum->ms_linear = s->ms_linear - lin0;
    sum->ms_cost = s->ms_cost - cost0;
    sum->ms_total = now_ms() - t_start;
    sum->kernel_launches = s->launches - launches0;
  }
  return 0;
}

// RAII temp solver for the one-shot API calls
struct TempSolver {
  gsfm_ra_solver* s = nullptr;
  ~TempSolver() { delete s; }
};

int make_temp(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, int device, TempSolver* t) {
  gsfm_ra_options o;
  gsfm_ra_default_options(&o);
  if (loss) o.loss = *loss;
  o.device = device;
  RA_TRY(build_solver(problem, &o, 0, 1, &t->s));
  if (omega) RA_TRY(gsfm_ra_solver_set_rotations(t->s, omega));
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int gsfm_ra_abi_version(void) { return GSFM_RA_ABI_VERSION; }
const char* gsfm_ra_last_error(void) { return g_last_error.c_str(); }
int gsfm_ra_residual_dim(int32_t error_type) {
  return error_type == GSFM_RA_QUATERNION_NORM ? 4 : (error_type == GSFM_RA_ROTATION_MAT_FNORM ? 9 : 3);
}
int gsfm_ra_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void gsfm_ra_default_options(gsfm_ra_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->loss.kind = GSFM_RA_LOSS_TRIVIAL;
  o->loss.scale = 1.0;
  o->max_num_iterations = 200;
  o->jacobi_scaling = 1;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->linear_solver = GSFM_RA_SOLVER_PCG;
  o->pcg_max_iterations = 500;
  o->pcg_rtol = 1e-10;
  o->num_threads = 0;
  o->device = -1;
  o->verbose = 0;
}

int gsfm_ra_solver_create(const gsfm_ra_problem* problem, const gsfm_ra_options* options, gsfm_ra_solver** out) {
  return build_solver(problem, options, 0, 1, out);
}
int gsfm_ra_solver_create_sharded(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int32_t rank, int32_t world_size,
                                  gsfm_ra_solver** out) {
  return build_solver(problem, options, rank, world_size, out);
}
void gsfm_ra_solver_destroy(gsfm_ra_solver* solver) {
  if (!solver) return;
  cudaSetDevice(solver->device);
  delete solver;
}
int gsfm_ra_solver_set_rotations(gsfm_ra_solver* s, const double* omega) {
  if (!s || !omega) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaMemcpyAsync(s->omega[s->cur].p, omega, 3ull * s->N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  reset_trust_region(s);
  return 0;
}
int gsfm_ra_solver_get_rotations(gsfm_ra_solver* s, double* omega) {
  if (!s || !omega) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaMemcpyAsync(omega, s->omega[s->cur].p, 3ull * s->N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}
int gsfm_ra_solver_reset(gsfm_ra_solver* s) {
  if (!s) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  reset_trust_region(s);
  return 0;
}
int gsfm_ra_solver_iterate(gsfm_ra_solver* s, int32_t num_iterations, gsfm_ra_summary* summary) {
  if (!s) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (s->opt.linear_solver == GSFM_RA_SOLVER_DENSE_CHOLESKY) {
    if (s->N > 2730) { set_error("dense Cholesky is limited to 2730 views (8190 unknowns); use PCG"); return GSFM_RA_ERR_UNSUPPORTED; }
    if (s->sharded() || !s->cooperative) { set_error("dense Cholesky needs a single cooperative-launch device"); return GSFM_RA_ERR_UNSUPPORTED; }
  } else if (s->opt.linear_solver != GSFM_RA_SOLVER_PCG) {
    set_error("unknown linear solver %d", s->opt.linear_solver);
    return GSFM_RA_ERR_INVALID;
  }
  return iterate(s, num_iterations, summary);
}
int gsfm_ra_comm_unique_id(uint8_t* id) {
  if (!id) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (!ncclx::api()) { set_error("libnccl.so.2 could not be loaded (set GSFM_RA_NCCL_LIB)"); return GSFM_RA_ERR_UNSUPPORTED; }
  ncclx::UniqueId u;
  NCCL_TRY(ncclx::api()->GetUniqueId(&u));
  std::memcpy(id, u.internal, ncclx::kUniqueIdBytes);
  return 0;
}
int gsfm_ra_solver_comm_init(gsfm_ra_solver* s, const uint8_t* id) {
  if (!s || !id) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (s->world == 1) return 0;
  if (!ncclx::api()) { set_error("libnccl.so.2 could not be loaded (set GSFM_RA_NCCL_LIB)"); return GSFM_RA_ERR_UNSUPPORTED; }
  CUDA_TRY(cudaSetDevice(s->device));
  ncclx::UniqueId u;
  std::memcpy(u.internal, id, ncclx::kUniqueIdBytes);
  NCCL_TRY(ncclx::api()->CommInitRank(&s->comm, s->world, u, s->rank));
  return 0;
}
int gsfm_ra_solver_ipc_export(gsfm_ra_solver* s, uint8_t* handle) {
  if (!s || !handle) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->xchg) {
    const size_t bytes = (6ull * s->N + 16) * sizeof(double);
    CUDA_TRY(cudaMalloc(&s->xchg, bytes));  // plain cudaMalloc: IPC handles cannot be taken from the async pool
    CUDA_TRY(cudaMemset(s->xchg, 0, bytes));
  }
  std::memset(handle, 0, GSFM_RA_IPC_HANDLE_BYTES);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, s->xchg));
  static_assert(sizeof(h) <= GSFM_RA_IPC_HANDLE_BYTES, "handle size");
  std::memcpy(handle, &h, sizeof(h));
  return 0;
}
int gsfm_ra_solver_ipc_import(gsfm_ra_solver* s, const uint8_t* handles) {
  if (!s || !handles) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (s->world > kMaxPeers) { set_error("at most %d ranks", kMaxPeers); return GSFM_RA_ERR_UNSUPPORTED; }
  if (!s->xchg) { set_error("call gsfm_ra_solver_ipc_export first"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) { s->peer_base[r] = s->xchg; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * GSFM_RA_IPC_HANDLE_BYTES, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(&s->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  s->peers_connected = true;
  return 0;
}
int gsfm_ra_solver_edge_range(const gsfm_ra_solver* s, uint64_t* e0, uint64_t* e1) {
  if (!s) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (e0) *e0 = s->edge_begin;
  if (e1) *e1 = s->edge_begin + s->E;
  return 0;
}

void* gsfm_ra_solver_cuda_stream(gsfm_ra_solver* s) { return s ? (void*)s->stream : nullptr; }

int gsfm_ra_solver_time_kernels(gsfm_ra_solver* s, int32_t repeats, double* out_ms) {
  if (!s || !out_ms || repeats < 1) { set_error("bad argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->linearized) { set_error("time_kernels needs a linearised solver (call iterate first)"); return GSFM_RA_ERR_INVALID; }
  const int b = s->cur, c = s->cur ^ 1;  // scratch output goes to the candidate buffer
  auto timed = [&](auto&& launch, double* ms) -> int {
    launch();  // warm
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    for (int k = 0; k < repeats; ++k) launch();
    CUDA_TRY(cudaGetLastError());
    *ms = s->elapsed_since(s->ev[0]) / repeats;
    return 0;
  };
  RA_TRY(timed([&] { s->launch_edges(b, true, s->val[c].p); }, &out_ms[0]));
  RA_TRY(timed([&] { s->launch_edges(b, false, nullptr); }, &out_ms[1]));
  CUDA_TRY(cudaMemsetAsync(s->z.p, 0, 4ull * s->N * sizeof(double), s->stream));
  RA_TRY(timed([&] { s->launch_spmv(b, s->z.p, std::getenv("GSFM_RA_DEBUG_NOGATHER") ? 2 : 0); }, &out_ms[2]));
  // one PCG iteration inside the persistent kernel: (time of R iterations) / R with rtol = 0
  {
    const int R = std::max(8, (int)repeats);
    RA_TRY(s->pcg_enqueue(b, s->radius, nullptr, nullptr, 0.0, R));  // warm
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    RA_TRY(s->pcg_enqueue(b, s->radius, nullptr, nullptr, 0.0, R));
    RA_TRY(s->fetch_scalars());
    const int done_it = std::max(1, s->h_sc->pcg_iter);
    out_ms[3] = s->elapsed_since(s->ev[0]) / done_it;
    if (std::getenv("GSFM_RA_PROFILE_PHASES")) {
      DevBuf<unsigned long long> pb;
      RA_TRY(pb.alloc(8));
      CUDA_TRY(cudaMemsetAsync(pb.p, 0, 64, s->stream));
      s->prof_buf = pb.p;
      const int rc = s->pcg_enqueue(b, s->radius, nullptr, nullptr, 0.0, R);
      s->prof_buf = nullptr;
      RA_TRY(rc);
      unsigned long long hp[8];
      CUDA_TRY(cudaMemcpyAsync(hp, pb.p, 64, cudaMemcpyDeviceToHost, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
      const double n = std::max<double>(1.0, (double)hp[6]);
      std::fprintf(stderr, "[gsfm_ra] PCG phases of block 0, us per iteration: spmv_pass %.2f | barrier+sum %.2f | vector update %.2f | barrier+sum %.2f  "
                   "(%d iterations)\n", hp[0] / n / 1e3, hp[1] / n / 1e3, hp[3] / n / 1e3, hp[4] / n / 1e3, (int)hp[6]);
    }
  }
  // restore the linearisation-dependent partials (K1 scratch wrote `part`): re-run the finalize inputs
  RA_TRY(s->evaluate(b, true));
  RA_TRY(s->fetch_scalars());
  return 0;
}

int gsfm_ra_solve(const gsfm_ra_problem* problem, const gsfm_ra_options* options, double* omega_inout, gsfm_ra_summary* summary) {
  if (!omega_inout) { set_error("omega_inout is NULL"); return GSFM_RA_ERR_INVALID; }
  const double t0 = now_ms();
  TempSolver t;
  RA_TRY(build_solver(problem, options, 0, 1, &t.s));
  RA_TRY(gsfm_ra_solver_set_rotations(t.s, omega_inout));
  const int rc = gsfm_ra_solver_iterate(t.s, options->max_num_iterations + 1, summary);
  if (rc != 0 && rc != GSFM_RA_ERR_NUMERIC) return rc;
  RA_TRY(gsfm_ra_solver_get_rotations(t.s, omega_inout));
  if (summary) summary->ms_total = now_ms() - t0;
  return rc;
}

int gsfm_ra_solve_sigma_consensus(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int32_t iters_num, double sigma_max,
                                  double* omega_inout, gsfm_ra_summary* summary) {
  if (!omega_inout || !options || !problem) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (problem->error_type != GSFM_RA_ANGLE_AXIS) { set_error("sigma consensus runs on GSFM_RA_ANGLE_AXIS (PairwiseRotationError with a scalar weight)"); return GSFM_RA_ERR_INVALID; }
  if (!(sigma_max > 0.0) || iters_num < 1) { set_error("sigma_max must be > 0 and iters_num >= 1"); return GSFM_RA_ERR_INVALID; }
  const double t0 = now_ms();
  gsfm_ra_problem q = *problem;
  q.edge_weight = nullptr;
  TempSolver t;
  RA_TRY(build_solver(&q, options, 0, 1, &t.s));
  gsfm_ra_solver* s = t.s;
  RA_TRY(gsfm_ra_solver_set_rotations(s, omega_inout));
  const uint64_t E = s->E, H = s->H;
  {
    AllocScope scope(s->stream);
    RA_TRY(s->d_weight.alloc(E));
  }
  CUDA_TRY(cudaMemsetAsync(s->d_weight.p, 0, E * sizeof(double), s->stream));  // last_weights start at 0
  // include/gamma_values.cpp:6-11 (nu = 3)
  const double C3 = 4.029720004054876e-01, gamma_k = 3.439485560754856e-03, table_size = 36843.0;
  const double sq2 = sigma_max * sigma_max * 2.0, one_over_sigma = C3 * 2.0 / sigma_max, weight_zero = one_over_sigma * (1.0 - gamma_k);
  gsfm_ra_summary total;
  std::memset(&total, 0, sizeof(total));
  if (summary) { total.trace = summary->trace; total.trace_capacity = summary->trace_capacity; }
  int rc = 0;
  for (int it = 0; it < iters_num; ++it) {
    const int b = s->cur;
    k_node_prep<<<grid_for(s->N), kBlock, 0, s->stream>>>(s->N, s->omega[b].p, s->node_q[b].p, s->node_JL[b].p, s->slots.p, s->counter.p, s->sc.p, 0);
    k_sigma_weights<<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->node_q[b].p, one_over_sigma, sq2, gamma_k,
                                                           weight_zero, table_size, s->d_weight.p, s->slots.p, s->counter.p, s->sc.p);
    k_setup_halfedges<<<grid_for(H), kBlock, 0, s->stream>>>(H, s->ku, s->he_edge.p, s->he_row.p, s->he_col.p, s->d_omega_ij.p, s->d_cov6.p, s->d_weight.p,
                                                             s->error_type, s->inrec.p);
    s->launches += 3;
    CUDA_TRY(cudaGetLastError());
    RA_TRY(s->fetch_scalars());
    const double diff = s->h_sc->dg / (double)E;
    reset_trust_region(s);
    gsfm_ra_summary s1;
    std::memset(&s1, 0, sizeof(s1));
    rc = gsfm_ra_solver_iterate(s, options->max_num_iterations + 1, &s1);
    if (rc != 0 && rc != GSFM_RA_ERR_NUMERIC) return rc;
    if (it == 0) total.initial_cost = s1.initial_cost;
    total.final_cost = s1.final_cost; total.termination = s1.termination;
    total.num_iterations += s1.num_iterations; total.num_successful_steps += s1.num_successful_steps;
    total.num_unsuccessful_steps += s1.num_unsuccessful_steps; total.total_linear_iterations += s1.total_linear_iterations;
    total.ms_assemble += s1.ms_assemble; total.ms_linear += s1.ms_linear; total.kernel_launches += s1.kernel_launches + 3;
    total.outer_iterations = it + 1;
    total.last_weight_change = diff;
    if (rc != 0 || diff <= 1e-7) break;
  }
  RA_TRY(gsfm_ra_solver_get_rotations(s, omega_inout));
  total.ms_setup = s->ms_setup;
  total.ms_total = now_ms() - t0;
  if (summary) *summary = total;
  return rc;
}

int gsfm_ra_eval_edges(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, double* r, double* jac_i, double* jac_j,
                       double* rho, int32_t device) {
  if (!omega) { set_error("omega is NULL"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint64_t E = s->E;
  DevBuf<double> dr, dji, djj, drho;
  const uint64_t d = (uint64_t)gsfm_ra_residual_dim(s->error_type);
  if (r) RA_TRY(dr.alloc(d * E));
  if (jac_i) RA_TRY(dji.alloc(3 * d * E));
  if (jac_j) RA_TRY(djj.alloc(3 * d * E));
  if (rho) RA_TRY(drho.alloc(3 * E));
  k_node_prep<<<grid_for(s->N), kBlock, 0, s->stream>>>(s->N, s->omega[0].p, s->node_q[0].p, s->node_JL[0].p, s->slots.p, s->counter.p, s->sc.p, s->manifold() ? 1 : 0);
  if (s->error_type == GSFM_RA_QUATERNION_NORM)
    k_eval_edges_general<0><<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->d_weight.p, s->node_q[0].p, s->node_JL[0].p,
                                                                   s->loss, dr.p, dji.p, djj.p, drho.p);
  else if (s->error_type == GSFM_RA_ROTATION_MAT_FNORM)
    k_eval_edges_general<1><<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->d_weight.p, s->node_q[0].p, s->node_JL[0].p,
                                                                   s->loss, dr.p, dji.p, djj.p, drho.p);
  else
    k_eval_edges<<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->d_cov6.p, s->d_weight.p, s->error_type,
                                                        s->node_q[0].p, s->node_JL[0].p, s->loss, dr.p, dji.p, djj.p, drho.p);
  CUDA_TRY(cudaGetLastError());
  if (r) CUDA_TRY(cudaMemcpyAsync(r, dr.p, d * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (jac_i) CUDA_TRY(cudaMemcpyAsync(jac_i, dji.p, 3 * d * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (jac_j) CUDA_TRY(cudaMemcpyAsync(jac_j, djj.p, 3 * d * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (rho) CUDA_TRY(cudaMemcpyAsync(rho, drho.p, 3 * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

int gsfm_ra_whiten(const gsfm_ra_problem* problem, double* U, int32_t device) {
  RA_TRY(check_problem(problem));
  if (!U) { set_error("U is NULL"); return GSFM_RA_ERR_INVALID; }
  int dev;
  RA_TRY(select_device(device, &dev));
  const uint64_t E = problem->num_edges;
  DevBuf<double> dc, dw, du;
  if (problem->cov6) { RA_TRY(dc.alloc(6 * E)); CUDA_TRY(cudaMemcpy(dc.p, problem->cov6, 6 * E * sizeof(double), cudaMemcpyHostToDevice)); }
  if (problem->edge_weight) { RA_TRY(dw.alloc(E)); CUDA_TRY(cudaMemcpy(dw.p, problem->edge_weight, E * sizeof(double), cudaMemcpyHostToDevice)); }
  RA_TRY(du.alloc(9 * E));
  k_whiten_edges<<<grid_for(E), kBlock>>>(E, dc.p, dw.p, problem->error_type, du.p);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(U, du.p, 9 * E * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int gsfm_ra_assemble(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, double* cost, double* gradient, double* hdiag,
                     uint32_t* rowptr, uint32_t* col, double* val, int32_t device) {
  if (!omega) { set_error("omega is NULL"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  RA_TRY(s->evaluate(0, true));
  RA_TRY(s->fetch_scalars());
  if (cost) *cost = s->h_sc->cost;
  const uint32_t N = s->N;
  const uint64_t H = s->H;
  DevBuf<double> dval, dh, dg;
  DevBuf<uint32_t> dcol;
  if (val || col) {
    RA_TRY(dval.alloc(9 * H));
    RA_TRY(dcol.alloc(H));
    k_export_blocks<<<grid_for(H), kBlock, 0, s->stream>>>(H, s->blk, s->he_row.p, s->he_col.p, s->val[0].p, s->node_JL[0].p, dval.p, dcol.p);
    CUDA_TRY(cudaGetLastError());
    if (val) CUDA_TRY(cudaMemcpyAsync(val, dval.p, 9 * H * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (col) CUDA_TRY(cudaMemcpyAsync(col, dcol.p, H * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  }
  if (hdiag || gradient) {
    RA_TRY(dh.alloc(9ull * N));
    RA_TRY(dg.alloc(3ull * N));
    k_export_nodes<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->Hd_p[0], s->gt_p[0], s->node_JL[0].p, dh.p, dg.p);
    CUDA_TRY(cudaGetLastError());
    if (hdiag) CUDA_TRY(cudaMemcpyAsync(hdiag, dh.p, 9ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (gradient) CUDA_TRY(cudaMemcpyAsync(gradient, dg.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (rowptr) {
    std::vector<uint32_t> he_row(H);
    CUDA_TRY(cudaMemcpy(he_row.data(), s->he_row.p, H * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::fill(rowptr, rowptr + N + 1, 0u);
    for (uint64_t h = 0; h < H; ++h) rowptr[he_row[h] + 1]++;
    for (uint32_t a = 0; a < N; ++a) rowptr[a + 1] += rowptr[a];
  }
  return 0;
}

int gsfm_ra_cost(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, double* cost, int32_t device) {
  if (!omega || !cost) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  RA_TRY(t.s->evaluate(0, false));
  RA_TRY(t.s->fetch_scalars());
  *cost = t.s->h_sc->cost;
  return 0;
}

int gsfm_ra_spmv(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, const double* damping, const double* x, double* y,
                 int32_t device) {
  if (!omega || !x || !y) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint32_t N = s->N;
  RA_TRY(s->evaluate(0, true));
  DevBuf<double> dx, dd;
  RA_TRY(dx.alloc(3ull * N));
  CUDA_TRY(cudaMemcpyAsync(dx.p, x, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (damping) { RA_TRY(dd.alloc(3ull * N)); CUDA_TRY(cudaMemcpyAsync(dd.p, damping, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream)); }
  // y = Jl^T Ht (Jl x) + damping .* x
  k_node_apply<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->node_JL[0].p, dx.p, 0, nullptr, nullptr, s->p.p, 4);
  RA_TRY(s->spmv(0, s->p.p, s->y.p, s->Hd_p[0]));
  k_node_apply<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->node_JL[0].p, s->y.p, 1, dd.p, dx.p, s->delta.p, 3);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(y, s->delta.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

int gsfm_ra_pcg(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, const double* damping, const double* b, double rtol,
                int32_t max_iterations, double* x, int32_t* iterations, double* rel_residual, int32_t device) {
  if (!omega || !b || !x) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint32_t N = s->N;
  RA_TRY(s->evaluate(0, true));
  DevBuf<double> db, dd;
  RA_TRY(db.alloc(3ull * N));
  RA_TRY(dd.alloc(3ull * N));
  CUDA_TRY(cudaMemcpyAsync(db.p, b, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (damping) CUDA_TRY(cudaMemcpyAsync(dd.p, damping, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  else CUDA_TRY(cudaMemsetAsync(dd.p, 0, 3ull * N * sizeof(double), s->stream));
  RA_TRY(s->pcg_enqueue(0, 1.0, dd.p, db.p, rtol, max_iterations));
  RA_TRY(s->fetch_scalars());
  const int it = s->h_sc->pcg_iter;
  const double res = (s->h_sc->bb > 0.0) ? std::sqrt(s->h_sc->rr / s->h_sc->bb) : 0.0;
  // x = Jl^-1 xt
  {
    ApplyArgs A;
    std::memset(&A, 0, sizeof(A));
    A.node_JL = s->node_JL[0].p; A.xt = s->x.p; A.delta_out = s->delta.p;
    k_apply_step<<<grid_for(N), kBlock, 0, s->stream>>>(N, A, s->slots.p, s->counter.p, s->sc.p);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(x, s->delta.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (iterations) *iterations = it;
  if (rel_residual) *rel_residual = res;
  return 0;
}

int gsfm_ra_eval_loss(const gsfm_ra_loss* loss, const double* s_in, uint64_t n, double* out, int32_t device) {
  if (!loss || !s_in || !out) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  int dev;
  RA_TRY(select_device(device, &dev));
  DevLoss L;
  RA_TRY(make_dev_loss(loss, &L));
  if (n == 0) return 0;
  DevBuf<double> ds, dout;
  RA_TRY(ds.alloc(n));
  RA_TRY(dout.alloc(3 * n));
  CUDA_TRY(cudaMemcpy(ds.p, s_in, n * sizeof(double), cudaMemcpyHostToDevice));
  k_eval_loss<<<grid_for(n), kBlock>>>(n, ds.p, L, dout.p);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out, dout.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int gsfm_ra_filter_view_pairs(const gsfm_ra_problem* problem, const double* omega, double max_degrees, uint8_t* keep, double* angle_rad,
                              int32_t device) {
  if (!omega || !problem) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (!(max_degrees >= 0.0)) { set_error("max_relative_rotation_difference_degrees must be >= 0"); return GSFM_RA_ERR_INVALID; }
  gsfm_ra_problem p2 = *problem;
  p2.error_type = GSFM_RA_ANGLE_AXIS;
  TempSolver t;
  RA_TRY(make_temp(&p2, nullptr, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint64_t E = s->E;
  DevBuf<uint8_t> dk;
  DevBuf<double> da;
  RA_TRY(dk.alloc(E));
  RA_TRY(da.alloc(E));
  const double thr = max_degrees * M_PI / 180.0;
  k_node_prep<<<grid_for(s->N), kBlock, 0, s->stream>>>(s->N, s->omega[0].p, s->node_q[0].p, s->node_JL[0].p, s->slots.p, s->counter.p, s->sc.p, s->manifold() ? 1 : 0);
  k_filter_pairs<<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->node_q[0].p, thr * thr, dk.p, da.p);
  CUDA_TRY(cudaGetLastError());
  if (keep) CUDA_TRY(cudaMemcpyAsync(keep, dk.p, E, cudaMemcpyDeviceToHost, s->stream));
  if (angle_rad) CUDA_TRY(cudaMemcpyAsync(angle_rad, da.p, E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

}  // extern "C"

#include "gsfm_graph.cuh"
