// K2 v3 microbenchmark: blocks-only records (1536 B) streamed RPC records per bulk copy, S copies in flight per warp;
// column ids of the warp's whole range RESIDENT in shared memory (one bulk copy at kernel start); x[col] gathers issued
// G records ahead with cp.async (2 x 16 B) into a per-warp x ring.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* d, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int RB = 1536;
// GM: 0 no gather, 1 cp.async ring, 2 direct LDG.128+LDG.64 one record ahead
template <int RPC, int S, int G, int WPB, int MAXREC, int GM>
__global__ void __launch_bounds__(WPB * 32) k_stream(const unsigned char* __restrict__ recs, const uint32_t* __restrict__ cols, uint32_t nrec_total,
                                                     const double* __restrict__ x4, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpBytes = S * RPC * RB + MAXREC * 128 + (G + 1) * 1024;
  unsigned char* ring = smem + (size_t)warp * kWarpBytes;
  uint32_t* ccol = reinterpret_cast<uint32_t*>(ring + S * RPC * RB);
  unsigned char* xring = ring + S * RPC * RB + MAXREC * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WPB * kWarpBytes) + warp * (S + 1);
  if (lane == 0) { for (int s = 0; s < S + 1; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint32_t nwarps = gridDim.x * WPB, gw = blockIdx.x * WPB + warp;
  uint32_t per = (nrec_total + nwarps - 1) / nwarps;
  per = (per + RPC - 1) / RPC * RPC;
  const uint32_t lo = gw * per;
  const uint32_t n = lo >= nrec_total ? 0 : min(per, nrec_total - lo);
  if (n == 0) return;
  const uint32_t ncopy = (n + RPC - 1) / RPC;
  auto issue = [&](uint32_t c) {
    if (lane == 0 && c < ncopy) {
      const uint32_t nr = min((uint32_t)RPC, n - c * RPC);
      mbar_expect_tx(&bars[c % S], nr * RB);
      tma_load_bulk(ring + (size_t)(c % S) * RPC * RB, recs + ((size_t)lo + (size_t)c * RPC) * RB, nr * RB, &bars[c % S]);
    }
  };
  if (lane == 0) { mbar_expect_tx(&bars[S], n * 128); tma_load_bulk(ccol, cols + (size_t)lo * 32, n * 128, &bars[S]); }
  for (uint32_t c = 0; c < (uint32_t)S; ++c) issue(c);
  mbar_wait(&bars[S], 0);
  auto gather_async = [&](uint32_t k) {
    const uint32_t col = ccol[k * 32 + lane];
    unsigned char* d = xring + (size_t)(k % (G + 1)) * 1024 + 16 * lane;
    const unsigned char* g = reinterpret_cast<const unsigned char*>(x4) + 32 * (size_t)col;
    cp_async16(d, g); cp_async16(d + 512, g + 16);
  };
  auto gather_reg = [&](uint32_t k, double& x0, double& x1, double& x2) {
    const uint32_t col = ccol[k * 32 + lane];
    const double2 a = *reinterpret_cast<const double2*>(x4 + 4 * (size_t)col); x0 = a.x; x1 = a.y; x2 = x4[4 * (size_t)col + 2];
  };
  if (GM == 1) for (int g = 0; g < G; ++g) { if ((uint32_t)g < n) gather_async(g); cp_async_commit(); }
  double acc = 0.0, r0 = 1, r1 = 2, r2 = 3;
  if (GM == 2) gather_reg(0, r0, r1, r2);
  for (uint32_t c = 0; c < ncopy; ++c) {
    mbar_wait(&bars[c % S], (c / S) & 1u);
    const uint32_t nr = min((uint32_t)RPC, n - c * RPC);
#pragma unroll
    for (uint32_t q = 0; q < (uint32_t)RPC; ++q) {
      if (q >= nr) break;
      const uint32_t k = c * RPC + q;
      double x0 = 1, x1 = 2, x2 = 3, n0 = 1, n1 = 2, n2 = 3;
      if (GM == 1) {
        if (k + G < n) gather_async(k + G);
        cp_async_commit();
        cp_async_wait<G>();
        const unsigned char* xs = xring + (size_t)(k % (G + 1)) * 1024 + 16 * lane;
        const double2 a = *reinterpret_cast<const double2*>(xs); x0 = a.x; x1 = a.y; x2 = *reinterpret_cast<const double*>(xs + 512);
      }
      if (GM == 2) { x0 = r0; x1 = r1; x2 = r2; if (k + 1 < n) gather_reg(k + 1, n0, n1, n2); }
      const double* r = reinterpret_cast<const double*>(ring + (size_t)(c % S) * RPC * RB + (size_t)q * RB);
      const double b0 = r[lane], b1 = r[32 + lane], b2 = r[64 + lane], b3 = r[96 + lane], b4 = r[128 + lane], b5 = r[160 + lane];
      acc += b0 * x0 + b1 * x1 + b2 * x2 + b3 * x0 + b4 * x1 + b5 * x2;
      if (GM == 2) { r0 = n0; r1 = n1; r2 = n2; }
    }
    __syncwarp();
    issue(c + S);
  }
  if (acc == 123.456) out[0] = acc;
}
template <int RPC, int S, int G, int WPB, int MAXREC, int GM>
void run(const unsigned char* d, const uint32_t* cols, uint32_t nrec, const double* x4, double* out, int bps) {
  const int smem = WPB * (S * RPC * RB + MAXREC * 128 + (G + 1) * 1024) + WPB * (S + 1) * 8;
  auto fn = k_stream<RPC, S, G, WPB, MAXREC, GM>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, WPB * 32, smem);
  const int grid = 148 * bps;
  const uint32_t per = ((nrec + grid * WPB - 1) / (grid * WPB) + RPC - 1) / RPC * RPC;
  if (occ < bps || per > MAXREC) { printf("RPC %d S %d G %d WPB %d bps %d: occ %d per %u skip\n", RPC, S, G, WPB, bps, occ, per); return; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) fn<<<grid, WPB * 32, smem>>>(d, cols, nrec, x4, out);
  cudaEventRecord(e0);
  const int reps = 500;
  for (int w = 0; w < reps; ++w) fn<<<grid, WPB * 32, smem>>>(d, cols, nrec, x4, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("RPC %d S %d G %d WPB %2d bps %d GM %d smem %6d per %3u: %.2f us  %.0f GB/s (blocks+cols) %s\n", RPC, S, G, WPB, bps, GM, smem, per, 1e3 * ms / reps,
         (double)nrec * (RB + 128) / (ms / reps * 1e-3) / 1e9, err ? cudaGetErrorString(err) : "");
}
int main() {
  const uint32_t nrec = 62500, N = 10000;
  std::vector<uint32_t> hc((size_t)nrec * 32);
  srand(1);
  for (auto& c : hc) c = (uint32_t)(rand() % N);
  unsigned char* d; cudaMalloc(&d, (size_t)(nrec + 64) * RB); cudaMemset(d, 0, (size_t)(nrec + 64) * RB);
  uint32_t* cols; cudaMalloc(&cols, (hc.size() + 4096) * 4); cudaMemcpy(cols, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
  double* x4; cudaMalloc(&x4, N * 32); cudaMemset(x4, 0, N * 32);
  double* out; cudaMalloc(&out, 64);
  // 16 warps / SM
  run<1, 4, 3, 8, 28, 0>(d, cols, nrec, x4, out, 2);
  run<1, 4, 3, 8, 28, 1>(d, cols, nrec, x4, out, 2);
  run<1, 4, 3, 8, 28, 2>(d, cols, nrec, x4, out, 2);
  run<2, 2, 3, 8, 28, 0>(d, cols, nrec, x4, out, 2);
  run<2, 2, 3, 8, 28, 1>(d, cols, nrec, x4, out, 2);
  run<2, 2, 3, 8, 28, 2>(d, cols, nrec, x4, out, 2);
  run<2, 3, 3, 8, 28, 1>(d, cols, nrec, x4, out, 2);
  run<2, 2, 5, 8, 28, 1>(d, cols, nrec, x4, out, 2);
  run<2, 2, 1, 8, 28, 1>(d, cols, nrec, x4, out, 2);
  run<3, 2, 3, 8, 30, 1>(d, cols, nrec, x4, out, 2);
  run<4, 2, 3, 8, 28, 1>(d, cols, nrec, x4, out, 2);
  // 8 warps / SM
  run<4, 2, 4, 8, 56, 0>(d, cols, nrec, x4, out, 1);
  run<4, 2, 4, 8, 56, 1>(d, cols, nrec, x4, out, 1);
  run<4, 2, 8, 8, 56, 1>(d, cols, nrec, x4, out, 1);
  run<4, 3, 4, 8, 56, 1>(d, cols, nrec, x4, out, 1);
  run<4, 2, 4, 8, 56, 2>(d, cols, nrec, x4, out, 1);
  // 12 warps / SM
  run<3, 2, 4, 12, 36, 1>(d, cols, nrec, x4, out, 1);
  run<2, 3, 4, 12, 36, 1>(d, cols, nrec, x4, out, 1);
  // 16 warps in one block
  run<2, 2, 3, 16, 28, 1>(d, cols, nrec, x4, out, 1);
  return 0;
}
