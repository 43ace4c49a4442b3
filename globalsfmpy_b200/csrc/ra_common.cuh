// ra_common.cuh -- constants, error plumbing, device-resident scalars, deterministic grid sum, TMA record pipeline
// Part of libgsfm_ra (one translation unit, see gsfm_ra.cu); reference citations sit next to each kernel.
#pragma once
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
#include <thread>
#include <vector>

#include "../../include/gsfm_ra.h"
#include "../../include/gsfm_pa.h"
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
//   kBlk = 4: COMPACT form of the scalar-weight stencil (angle-axis types 4, 5, 7, 8): S = a I + sigma h h^T is a scaled
//             identity plus a rank-one term, stored as (sigma a, h0, h1, h2) with a > 0 carrying the sign sigma -- the
//             off-diagonal block is -S;                                                  1152 B per record (36 B / half-edge)
//   kBlk = 6: symmetric off-diagonal block -S packed (00,01,02,11,12,22) -- every residual that is a function of the
//             error rotation (types 2..8) is a Laplacian stencil in the body frame;  1664 B per record (52 B / half-edge)
//   kBlk = 9: general row-major block (QUATERNION_NORM, ROTATION_MAT_FNORM);         2432 B per record (76 B / half-edge)
template <int kBlk>
struct Rec {
  static constexpr int kDoubles = kBlk * 32 + 16;
  static constexpr int kBytes = kDoubles * 8;
  static constexpr int kColOffset = kBlk * 32;  // doubles: col[] starts here
};
#ifdef GSFM_RA_NO_COMPACT
constexpr bool kCompactScalarStencil = false;  // A/B builds: keep 6-double records for the scalar-weight stencil
#else
constexpr bool kCompactScalarStencil = true;
#endif
constexpr int kStages = 4;                       // K1: TMA ring depth per warp (one input record per bulk copy)
// K2-class kernels (k_spmv, k_pcg_persistent).  Measured on B200 (profiles/r02b_*): the TMA unit of an SM retires ~25 bulk
// copies per microsecond whatever their size, so the matrix stream moves in CHUNKS of several consecutive records per bulk
// copy (3.3 - 3.5 KB), three chunks in flight per warp; and a grid barrier costs ~3 us with 148 participants against ~4 us
// with 296, so these kernels run ONE block of 16 warps per SM.  (-DGSFM_RA_PCG_WARPS=8: two blocks of 8 warps, A/B builds.)
#ifndef GSFM_RA_PCG_WARPS
#define GSFM_RA_PCG_WARPS 16
#endif
constexpr int kPcgWarps = GSFM_RA_PCG_WARPS;
constexpr int kPcgBlock = kPcgWarps * 32;
constexpr int kPcgBlocksPerSM = 16 / kPcgWarps;
constexpr int kMaxWarpsPerBlock = 16;
constexpr int kStages2 = 3;                      // K2: chunks in flight per warp
// K2 loop form for the compact 4-double records: record-major (one record per iteration) or chunk-major (the records of a bulk
// copy unrolled together); A/B builds with -DGSFM_RA_K2_CHUNK_MAJOR4.
#ifdef GSFM_RA_K2_CHUNK_MAJOR4
constexpr bool kRecordMajor4 = false;
#else
constexpr bool kRecordMajor4 = true;
#endif
#ifdef GSFM_RA_NO_WRAP_PREFETCH
constexpr bool kWrapPrefetch = false;   // A/B builds: no wrap-around prefetch of the matrix stream across the CG step's barriers
#else
constexpr bool kWrapPrefetch = true;
#endif
constexpr int chunk_recs(int blk) { return blk == 4 ? 3 : (blk == 6 ? 2 : 1); }   // records per bulk copy
template <int kBlk>
struct Chunk {
  static constexpr int kRecs = chunk_recs(kBlk);
  static constexpr int kDoubles = kRecs * Rec<kBlk>::kDoubles;
  static constexpr int kBytes = kDoubles * 8;
};
// Small graphs: the gathered vector z (24 B per view) is staged in shared memory once per pass and the per-half-edge gather
// reads it there -- the L1/L2 gather path of an SM retires about one 32 B sector per clock, which at 1M half-edges per pass is
// as expensive as the matrix stream itself.  kSliceMaxViews * 24 B must fit beside the rings (227 KB per block).
constexpr int kSliceMaxViews = 2560;
constexpr int spmv_smem_bytes(int blk) { return kPcgWarps * kStages2 * chunk_recs(blk) * (blk * 32 + 16) * 8 + kPcgWarps * kStages2 * 8; }

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

// Several values reduced across the warp at once (reduce-scatter butterfly): at every step a lane keeps one half of its values,
// sends the other half to its partner and adds what it receives, so 16 values cost 8 + 4 + 2 + 1 + 1 = 16 exchanges instead of
// 16 x 5.  On return lane L holds in `out` the warp total of value number multi_index16(L) (lanes L and L ^ 1 hold the same).
// Fixed order: bit-reproducible.
__device__ __forceinline__ int multi_index16(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
template <int K>
__device__ __forceinline__ double warp_sum_multi16(const double (&v)[K]) {
  static_assert(K <= 16, "at most 16 values");
  const int lane = threadIdx.x & 31;
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = i < K ? v[i] : 0.0;
  double b[8];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double keep = hi ? a[8 + i] : a[i], send = hi ? a[i] : a[8 + i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  double c[4];
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double keep = hi ? b[4 + i] : b[i], send = hi ? b[i] : b[4 + i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  double d[2];
  {
    const bool hi = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double keep = hi ? c[2 + i] : c[i], send = hi ? c[i] : c[2 + i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool hi = (lane & 2) != 0;
  double e = (hi ? d[1] : d[0]) + __shfl_xor_sync(0xffffffffu, hi ? d[0] : d[1], 2);
  e += __shfl_xor_sync(0xffffffffu, e, 1);
  return e;
}
// Three values (the SpMV's row sums): on return lanes 0, 8, 16 hold the totals of v0, v1, v2 (component = multi_index4(lane)
// on the lanes with (lane & 7) == 0).
__device__ __forceinline__ int multi_index4(int lane) { return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1); }
__device__ __forceinline__ double warp_sum_multi3(double v0, double v1, double v2) {
  const int lane = threadIdx.x & 31;
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  const double b0 = (hi16 ? v2 : v0) + __shfl_xor_sync(0xffffffffu, hi16 ? v0 : v2, 16);
  const double b1 = (hi16 ? 0.0 : v1) + __shfl_xor_sync(0xffffffffu, hi16 ? v1 : 0.0, 16);
  double e = (hi8 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, hi8 ? b0 : b1, 8);
  e += __shfl_xor_sync(0xffffffffu, e, 4);
  e += __shfl_xor_sync(0xffffffffu, e, 2);
  e += __shfl_xor_sync(0xffffffffu, e, 1);
  return e;
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
  int bar_seq;            // grid barrier sequence number of the persistent PCG kernel
  int eseq;               // cross-GPU exchange sequence number of the per-view sums (one per evaluation)
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

// ------------------------------------------------------------------------------------------
// Cross-GPU exchange cells (edge-sharded solver).  One double travels as a 16-byte cell {lo32, tag, hi32, tag} written with ONE
// vector store: each 8-byte half carries its own copy of the sequence tag, 8-byte stores are atomic over NVLink and in L2, so a
// reader that sees both tags equal to the sequence number it waits for has the value -- no fence, no separate flag, and the
// producer can push cells into peer memory while it is still computing (the low-latency protocol of NCCL, applied per double).
// The exchange block of a rank (cudaMalloc, zero-initialised, mapped by every peer over NVLink) holds, double buffered by
// sequence parity and indexed by SOURCE rank:   cg [2][W][3N]   lin [2][W][9N]   tail [2][W][4]   and the REDUCED values of the
// owner mode   cgred [2][3N]   linred [2][9N]     (cells).
// Two exchange modes.  DIRECT (W <= 2): every rank pushes its contribution for every view to every rank, every rank adds the W
// contributions itself -- one NVLink hop, W x the stores.  OWNER (W > 2): view i belongs to rank i % W; contributions go to the
// owner only, the owner adds them in rank order and pushes the sum to every rank -- two hops, but 2N instead of W N cell groups
// per rank, which is what counts once N W is large (100k views on 8 GPUs: 4x fewer NVLink writes).  Either way every rank ends
// up with bitwise the same sums.
// ------------------------------------------------------------------------------------------
struct alignas(16) LLCell { uint32_t lo, t0, hi, t1; };
constexpr int kMaxPeers = 16;
struct PeerPtrs { LLCell* p[kMaxPeers]; };
__host__ __device__ inline size_t ll_cells_total(uint32_t N, int world) { return 2ull * world * (12ull * N + 4) + 2ull * 12ull * N; }
__host__ __device__ inline size_t ll_cgred_offset(uint32_t N, int world, unsigned seq) { return 2ull * world * (12ull * N + 4) + (size_t)(seq & 1u) * 3ull * N; }
__host__ __device__ inline size_t ll_linred_offset(uint32_t N, int world, unsigned seq) { return 2ull * world * (12ull * N + 4) + 2ull * 3ull * N + (size_t)(seq & 1u) * 9ull * N; }
__host__ __device__ inline size_t ll_cg_offset(uint32_t N, int world, unsigned seq, int src) { return ((size_t)(seq & 1u) * world + src) * 3ull * N; }
__host__ __device__ inline size_t ll_lin_offset(uint32_t N, int world, unsigned seq, int src) { return 2ull * world * 3ull * N + ((size_t)(seq & 1u) * world + src) * 9ull * N; }
__host__ __device__ inline size_t ll_tail_offset(uint32_t N, int world, unsigned seq, int src) { return 2ull * world * 12ull * N + ((size_t)(seq & 1u) * world + src) * 4ull; }
__device__ __forceinline__ void ll_store(LLCell* p, double v, uint32_t tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)b), "r"(tag), "r"((uint32_t)(b >> 32)), "r"(tag) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const LLCell* p, uint32_t tag, double& v) {
  uint32_t a, b, c, d;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  v = __longlong_as_double((long long)(((unsigned long long)c << 32) | a));
  return b == tag && d == tag;
}
// Spin until the cell carries `tag` (a peer that never shows up: ~4 s, then *bad = 2 instead of a hung GPU).
__device__ __forceinline__ double ll_wait(const LLCell* p, uint32_t tag, int* bad) {
  double v;
  if (ll_try_load(p, tag, v)) return v;
  const long long t0 = clock64();
  while (!ll_try_load(p, tag, v)) {
    if (clock64() - t0 > 8000000000ll) { *bad = 2; return 0.0; }
  }
  return v;
}

// K cells at once: the loads of one poll are issued back to back (one L2 round trip), all tags are then checked together.
template <int K>
__device__ __forceinline__ void ll_wait_n(const LLCell* c, uint32_t tag, double (&a)[K], int* bad) {
  long long t0 = 0;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < K; ++k) ok = ll_try_load(c + k, tag, a[k]) && ok;
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 8000000000ll) {
      *bad = 2;
#pragma unroll
      for (int k = 0; k < K; ++k) a[k] = 0.0;
      return;
    }
  }
}

// Deterministic grid-wide sum of NV values: block tree -> per-block slot -> the LAST block to
// arrive adds the slots in a fixed order.  Returns true in the last block (all threads), with the
// totals in `tot` (valid in thread 0 only).
template <int NV>
__device__ bool grid_sum(double (&v)[NV], double* slots, unsigned* counter, double (&tot)[NV]) {
  __shared__ double sm[NV][kMaxWarpsPerBlock];
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
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[k][w];
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
  double* ring;    // this warp's ring of stages in shared memory
  uint64_t* bars;  // this warp's mbarriers, one per stage
  uint32_t pos;    // bulk copies consumed since init: ring slot = pos % stages, phase = (pos / stages) & 1
  uint32_t primed; // bulk copies of the NEXT pass already in flight (wrap-around prefetch of the persistent PCG kernel)
};

// Shared memory of a block: [warps][kNumStages][kStageBytes] rings, then [warps][kNumStages] mbarriers.
template <int kStageBytes, int kNumStages>
__device__ __forceinline__ void pipe_init_bytes(WarpPipe& wp, unsigned char* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  wp.ring = reinterpret_cast<double*>(smem + (size_t)warp * kNumStages * kStageBytes);
  wp.bars = reinterpret_cast<uint64_t*>(smem + (size_t)nwarps * kNumStages * kStageBytes) + warp * kNumStages;
  wp.pos = 0;
  wp.primed = 0;
  if (lane == 0) {
    for (int st = 0; st < kNumStages; ++st) mbar_init(&wp.bars[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}
template <int kBlk>
__device__ __forceinline__ void pipe_init(WarpPipe& wp, unsigned char* smem) { pipe_init_bytes<Chunk<kBlk>::kBytes, kStages2>(wp, smem); }

}  // namespace
