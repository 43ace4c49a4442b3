// tiled_k2_bench.cu -- round-2 design microbenchmark (NOT part of the product; compiled here, first run planned for round 2).
//
// Symmetric-half TILED SpMV for the Laplacian-stencil matrix (layout validated on the host by
// tiled_symmetric_prototype.py).  Every edge (i, j), i < j, carries ONE symmetric 3x3 block S:
//     y_i += S (x_i - x_j),   y_j -= S (x_i - x_j)
// Views are cut into groups of GS; an edge belongs to tile (group(i), group(j)); tiles are cut into CHUNKS of <= CE edges
// (whole 32-edge records), one CTA per chunk at a time.  A CTA
//   1. stages the two x slices of its tile in shared memory (coalesced),
//   2. streams its records {double S[6][32]; u16 row[32]; u16 col[32]} (1664 B, one bulk async copy each), forms
//      d_e = S_e (x_i - x_j) with shared-memory gathers and parks d_e in shared memory,
//   3. reduces d twice, deterministically: by row (edges of a chunk are sorted by row: a thread sums one row segment) and by
//      column (through the chunk's column permutation: a thread sums one column segment),
//   4. writes the segment sums to the chunk's partial slots.
// A second kernel adds every view's partial slots in a fixed order.  No float atomics anywhere.
// The program checks y against a host reference and prints the time per pass and the bytes streamed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tiled_k2_bench tiled_k2_bench.cu && ./tiled_k2_bench
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e__ = (x);                                                                      \
    if (e__ != cudaSuccess) { std::printf("%s failed: %s\n", #x, cudaGetErrorString(e__)); std::exit(1); } \
  } while (0)

#ifndef TILED_GS
#define TILED_GS 1250
#endif
#ifndef TILED_CE
#define TILED_CE 1664
#endif
constexpr int GS = TILED_GS;    // views per group: 2 slices x GS x 24 B of shared memory (1250 -> 60 KB)
constexpr int CE = TILED_CE;    // edges per chunk (a multiple of 32; 1664 = 52 records)
constexpr int kRecBytes = 6 * 32 * 8 + 2 * 32 * 2;  // 1664
constexpr int kThreads = 256, kWarps = 8, kStages = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Chunk {
  uint32_t rec_begin, num_recs;      // records of this chunk
  uint32_t num_edges;                // edges (the last record may be padded)
  uint32_t group_a, group_b;         // row group, column group
  uint32_t rowseg_begin, num_rowsegs;  // rowseg_ptr[rowseg_begin .. + num_rowsegs]: chunk-local edge offsets
  uint32_t colseg_begin, num_colsegs;
  uint32_t slot_begin;               // first partial slot (row segments first, then column segments)
  uint32_t perm_begin;               // into perm (chunk-local edge indices sorted by column)
};

struct Params {
  const unsigned char* recs;
  const Chunk* chunks;
  uint32_t num_chunks;
  const uint32_t* rowseg_ptr;   // per chunk: num_rowsegs + 1 chunk-local edge offsets
  const uint32_t* colseg_ptr;   // per chunk: num_colsegs + 1 offsets into the chunk's perm
  const uint16_t* perm;
  const double* x;              // [N][3]
  double* partial;              // [num_slots][3]
  uint32_t N;
};

// shared memory: xs_a[GS*3] xs_b[GS*3] d[CE*3] ring[kWarps][kStages][kRecBytes] bars[kWarps][kStages]
constexpr size_t kSmemBytes = (size_t)2 * GS * 3 * 8 + (size_t)CE * 3 * 8 + (size_t)kWarps * kStages * kRecBytes + kWarps * kStages * 8;

__global__ void __launch_bounds__(kThreads) k_tiled_pass(Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* xs_a = reinterpret_cast<double*>(smem);
  double* xs_b = xs_a + GS * 3;
  double* dbuf = xs_b + GS * 3;
  unsigned char* ring = reinterpret_cast<unsigned char*>(dbuf + CE * 3);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)kWarps * kStages * kRecBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* my_ring = ring + (size_t)warp * kStages * kRecBytes;
  uint64_t* my_bars = bars + warp * kStages;
  if (lane == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&my_bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t issued = 0, consumed = 0;  // this warp's record counters (ring slot / parity)
  uint32_t cur_a = 0xffffffffu, cur_b = 0xffffffffu;
  for (uint32_t c = blockIdx.x; c < P.num_chunks; c += gridDim.x) {
    const Chunk ch = P.chunks[c];
    // 1. x slices (skipped when the previous chunk of this CTA used the same group)
    if (ch.group_a != cur_a) {
      const uint32_t base = ch.group_a * GS, n = min((uint32_t)GS, P.N - base) * 3;
      for (uint32_t k = threadIdx.x; k < n; k += kThreads) xs_a[k] = P.x[(size_t)base * 3 + k];
      cur_a = ch.group_a;
    }
    if (ch.group_b != cur_b) {
      const uint32_t base = ch.group_b * GS, n = min((uint32_t)GS, P.N - base) * 3;
      for (uint32_t k = threadIdx.x; k < n; k += kThreads) xs_b[k] = P.x[(size_t)base * 3 + k];
      cur_b = ch.group_b;
    }
    __syncthreads();
    // 2. records warp, warp + 8, ... of the chunk
    const uint32_t my_n = ch.num_recs > (uint32_t)warp ? (ch.num_recs - warp + kWarps - 1) / kWarps : 0;
    auto issue = [&](uint32_t k) {  // k-th record of this warp in this chunk
      if (lane == 0) {
        const uint32_t st = issued % kStages;
        mbar_expect_tx(&my_bars[st], kRecBytes);
        tma_load_bulk(my_ring + (size_t)st * kRecBytes, P.recs + (size_t)(ch.rec_begin + warp + k * kWarps) * kRecBytes, kRecBytes, &my_bars[st]);
      }
      ++issued;
    };
    for (uint32_t k = 0; k < my_n && k < (uint32_t)kStages; ++k) issue(k);
    for (uint32_t k = 0; k < my_n; ++k) {
      const uint32_t st = consumed % kStages;
      mbar_wait(&my_bars[st], (consumed / kStages) & 1u);
      const unsigned char* rec = my_ring + (size_t)st * kRecBytes;
      const double* S = reinterpret_cast<const double*>(rec);
      const uint16_t* rc = reinterpret_cast<const uint16_t*>(rec + 6 * 32 * 8);
      const uint32_t row = rc[lane], col = rc[32 + lane];
      const double dx0 = xs_a[3 * row] - xs_b[3 * col], dx1 = xs_a[3 * row + 1] - xs_b[3 * col + 1], dx2 = xs_a[3 * row + 2] - xs_b[3 * col + 2];
      const double s0 = S[lane], s1 = S[32 + lane], s2 = S[64 + lane], s3 = S[96 + lane], s4 = S[128 + lane], s5 = S[160 + lane];
      const uint32_t e = (warp + k * kWarps) * 32 + lane;  // chunk-local edge index (padding edges carry S = 0)
      dbuf[3 * e] = s0 * dx0 + s1 * dx1 + s2 * dx2;
      dbuf[3 * e + 1] = s1 * dx0 + s3 * dx1 + s4 * dx2;
      dbuf[3 * e + 2] = s2 * dx0 + s4 * dx1 + s5 * dx2;
      ++consumed;
      __syncwarp();
      if (k + kStages < my_n) issue(k + kStages);
    }
    __syncthreads();
    // 3 + 4. row segments (contiguous edges), then column segments (through the permutation)
    const uint32_t* rp = P.rowseg_ptr + ch.rowseg_begin;
    for (uint32_t s = threadIdx.x; s < ch.num_rowsegs; s += kThreads) {
      double a0 = 0, a1 = 0, a2 = 0;
      for (uint32_t e = rp[s]; e < rp[s + 1]; ++e) { a0 += dbuf[3 * e]; a1 += dbuf[3 * e + 1]; a2 += dbuf[3 * e + 2]; }
      double* out = P.partial + 3 * (size_t)(ch.slot_begin + s);
      out[0] = a0; out[1] = a1; out[2] = a2;
    }
    const uint32_t* cp = P.colseg_ptr + ch.colseg_begin;
    const uint16_t* pm = P.perm + ch.perm_begin;
    for (uint32_t s = threadIdx.x; s < ch.num_colsegs; s += kThreads) {
      double a0 = 0, a1 = 0, a2 = 0;
      for (uint32_t k = cp[s]; k < cp[s + 1]; ++k) { const uint32_t e = pm[k]; a0 -= dbuf[3 * e]; a1 -= dbuf[3 * e + 1]; a2 -= dbuf[3 * e + 2]; }
      double* out = P.partial + 3 * (size_t)(ch.slot_begin + ch.num_rowsegs + s);
      out[0] = a0; out[1] = a1; out[2] = a2;
    }
    __syncthreads();  // dbuf and the x slices are reused by the next chunk
  }
}

// y_v = sum of the view's partial slots in slot order (view_slot_ptr / view_slots: CSR built on the host)
__global__ void k_gather_partials(uint32_t N, const uint32_t* __restrict__ view_slot_ptr, const uint32_t* __restrict__ view_slots,
                                  const double* __restrict__ partial, double* __restrict__ y) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  double a0 = 0, a1 = 0, a2 = 0;
  for (uint32_t k = view_slot_ptr[v]; k < view_slot_ptr[v + 1]; ++k) {
    const double* p = partial + 3 * (size_t)view_slots[k];
    a0 += p[0]; a1 += p[1]; a2 += p[2];
  }
  y[3 * (size_t)v] = a0; y[3 * (size_t)v + 1] = a1; y[3 * (size_t)v + 2] = a2;
}

int main(int argc, char** argv) {
  const uint32_t N = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 10000;
  const uint64_t E = argc > 2 ? (uint64_t)std::atoll(argv[2]) : 1000000;
  std::mt19937_64 rng(56);
  // edges i < j: a chain + random pairs (duplicates allowed: this is a bandwidth experiment)
  std::vector<uint32_t> ei(E), ej(E);
  for (uint64_t e = 0; e < E; ++e) {
    uint32_t a, b;
    if (e + 1 < N) { a = (uint32_t)e; b = (uint32_t)e + 1; }
    else { do { a = (uint32_t)(rng() % N); b = (uint32_t)(rng() % N); } while (a == b); }
    ei[e] = std::min(a, b); ej[e] = std::max(a, b);
  }
  std::vector<double> S(6 * E), x(3 * (size_t)N);
  std::normal_distribution<double> g(0.0, 1.0);
  for (auto& v : S) v = g(rng);
  for (auto& v : x) v = g(rng);
  // host reference
  std::vector<double> yref(3 * (size_t)N, 0.0);
  for (uint64_t e = 0; e < E; ++e) {
    const double* s = &S[6 * e];
    const double d0 = x[3 * ei[e]] - x[3 * ej[e]], d1 = x[3 * ei[e] + 1] - x[3 * ej[e] + 1], d2 = x[3 * ei[e] + 2] - x[3 * ej[e] + 2];
    const double r0 = s[0] * d0 + s[1] * d1 + s[2] * d2, r1 = s[1] * d0 + s[3] * d1 + s[4] * d2, r2 = s[2] * d0 + s[4] * d1 + s[5] * d2;
    yref[3 * ei[e]] += r0; yref[3 * ei[e] + 1] += r1; yref[3 * ei[e] + 2] += r2;
    yref[3 * ej[e]] -= r0; yref[3 * ej[e] + 1] -= r1; yref[3 * ej[e] + 2] -= r2;
  }
  // ---- layout: sort by (tile, row, col), cut tiles into chunks ------------------------------------------
  const uint32_t G = (N + GS - 1) / GS;
  std::vector<uint64_t> order(E);
  std::iota(order.begin(), order.end(), 0);
  auto tile_of = [&](uint64_t e) { return (uint64_t)(ei[e] / GS) * G + ej[e] / GS; };
  std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
    const uint64_t ta = tile_of(a), tb = tile_of(b);
    if (ta != tb) return ta < tb;
    if (ei[a] != ei[b]) return ei[a] < ei[b];
    return ej[a] < ej[b];
  });
  std::vector<Chunk> chunks;
  std::vector<unsigned char> recs;
  std::vector<uint32_t> rowseg_ptr, colseg_ptr, slot_view;  // slot_view: the view a partial slot belongs to
  std::vector<uint16_t> perm;
  uint64_t pos = 0;
  while (pos < E) {
    const uint64_t t = tile_of(order[pos]);
    uint64_t end = pos;
    while (end < E && end - pos < (uint64_t)CE && tile_of(order[end]) == t) ++end;
    Chunk ch{};
    ch.num_edges = (uint32_t)(end - pos);
    ch.num_recs = (ch.num_edges + 31) / 32;
    ch.rec_begin = (uint32_t)(recs.size() / kRecBytes);
    ch.group_a = (uint32_t)(t / G); ch.group_b = (uint32_t)(t % G);
    recs.resize(recs.size() + (size_t)ch.num_recs * kRecBytes, 0);
    for (uint32_t k = 0; k < ch.num_edges; ++k) {
      const uint64_t e = order[pos + k];
      unsigned char* rec = recs.data() + (size_t)(ch.rec_begin + k / 32) * kRecBytes;
      double* Sr = reinterpret_cast<double*>(rec);
      uint16_t* rc = reinterpret_cast<uint16_t*>(rec + 6 * 32 * 8);
      for (int q = 0; q < 6; ++q) Sr[q * 32 + k % 32] = S[6 * e + q];
      rc[k % 32] = (uint16_t)(ei[e] - ch.group_a * GS);
      rc[32 + k % 32] = (uint16_t)(ej[e] - ch.group_b * GS);
    }
    // row segments: runs of equal row (chunk-local edge offsets)
    ch.rowseg_begin = (uint32_t)rowseg_ptr.size();  // this chunk's num_rowsegs + 1 offsets start here
    ch.slot_begin = (uint32_t)slot_view.size();
    for (uint32_t k = 0; k < ch.num_edges; ++k)
      if (k == 0 || ei[order[pos + k]] != ei[order[pos + k - 1]]) { rowseg_ptr.push_back(k); slot_view.push_back(ei[order[pos + k]]); ++ch.num_rowsegs; }
    rowseg_ptr.push_back(ch.num_edges);
    // column permutation and segments
    std::vector<uint16_t> p(ch.num_edges);
    std::iota(p.begin(), p.end(), (uint16_t)0);
    std::stable_sort(p.begin(), p.end(), [&](uint16_t a, uint16_t b) { return ej[order[pos + a]] < ej[order[pos + b]]; });
    ch.perm_begin = (uint32_t)perm.size();
    ch.colseg_begin = (uint32_t)colseg_ptr.size();
    for (uint32_t k = 0; k < ch.num_edges; ++k) {
      if (k == 0 || ej[order[pos + p[k]]] != ej[order[pos + p[k - 1]]]) { colseg_ptr.push_back(k); slot_view.push_back(ej[order[pos + p[k]]]); ++ch.num_colsegs; }
      perm.push_back(p[k]);
    }
    colseg_ptr.push_back(ch.num_edges);
    chunks.push_back(ch);
    pos = end;
  }
  const uint32_t num_slots = (uint32_t)slot_view.size();
  // per view: its slots in slot order
  std::vector<uint32_t> view_slot_ptr(N + 1, 0), view_slots(num_slots);
  for (uint32_t s = 0; s < num_slots; ++s) ++view_slot_ptr[slot_view[s] + 1];
  for (uint32_t v = 0; v < N; ++v) view_slot_ptr[v + 1] += view_slot_ptr[v];
  {
    std::vector<uint32_t> fill(view_slot_ptr.begin(), view_slot_ptr.end() - 1);
    for (uint32_t s = 0; s < num_slots; ++s) view_slots[fill[slot_view[s]]++] = s;
  }
  std::printf("N %u E %llu: %u groups, %zu chunks, %zu records, %u partial slots; streamed per pass: records %.1f MB + perm %.1f MB + "
              "segment tables %.1f MB + partials %.1f MB (written) + %.1f MB (read)\n", N, (unsigned long long)E, G, chunks.size(),
              recs.size() / (size_t)kRecBytes, num_slots, recs.size() / 1e6, perm.size() * 2 / 1e6, (rowseg_ptr.size() + colseg_ptr.size()) * 4 / 1e6,
              num_slots * 24 / 1e6, num_slots * 28 / 1e6);
  // ---- device ------------------------------------------------------------------------------------------------
  auto up = [&](const void* h, size_t bytes) { void* d; CK(cudaMalloc(&d, std::max<size_t>(bytes, 16))); CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice)); return d; };
  Params P{};
  P.recs = (const unsigned char*)up(recs.data(), recs.size());
  P.chunks = (const Chunk*)up(chunks.data(), chunks.size() * sizeof(Chunk));
  P.num_chunks = (uint32_t)chunks.size();
  P.rowseg_ptr = (const uint32_t*)up(rowseg_ptr.data(), rowseg_ptr.size() * 4);
  P.colseg_ptr = (const uint32_t*)up(colseg_ptr.data(), colseg_ptr.size() * 4);
  P.perm = (const uint16_t*)up(perm.data(), perm.size() * 2);
  P.x = (const double*)up(x.data(), x.size() * 8);
  P.N = N;
  double *partial, *y;
  CK(cudaMalloc(&partial, (size_t)num_slots * 24)); CK(cudaMalloc(&y, (size_t)N * 24));
  P.partial = partial;
  const uint32_t* d_vsp = (const uint32_t*)up(view_slot_ptr.data(), view_slot_ptr.size() * 4);
  const uint32_t* d_vs = (const uint32_t*)up(view_slots.data(), view_slots.size() * 4);
  CK(cudaFuncSetAttribute(k_tiled_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  int occ = 0, sms = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tiled_pass, kThreads, kSmemBytes));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = std::min<int>((int)chunks.size(), sms * std::max(1, occ));
  std::printf("shared memory %zu B per CTA, %d CTA(s) per SM, grid %d\n", kSmemBytes, occ, grid);
  auto pass = [&]() {
    k_tiled_pass<<<grid, kThreads, kSmemBytes>>>(P);
    k_gather_partials<<<(N + 255) / 256, 256>>>(N, d_vsp, d_vs, partial, y);
  };
  pass();
  CK(cudaDeviceSynchronize());
  std::vector<double> yh(3 * (size_t)N);
  CK(cudaMemcpy(yh.data(), y, yh.size() * 8, cudaMemcpyDeviceToHost));
  double err = 0, ref = 0;
  for (size_t k = 0; k < yh.size(); ++k) { err = std::max(err, std::fabs(yh[k] - yref[k])); ref = std::max(ref, std::fabs(yref[k])); }
  std::printf("max |y - y_ref| / max |y_ref| = %.2e %s\n", err / ref, err / ref < 1e-12 ? "(ok)" : "(MISMATCH)");
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 500;  // CUDA event timestamps tick in ~41 us steps on these boxes: average many launches
  for (int w = 0; w < 5; ++w) pass();
  CK(cudaEventRecord(e0));
  for (int w = 0; w < reps; ++w) pass();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::printf("tiled symmetric pass (both kernels): %.2f us per pass; today's half-edge pass: 22.6 us (k_spmv, 104 MB stored)\n", 1e3 * ms / reps);
  return 0;
}
