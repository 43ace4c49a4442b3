// K2-shaped microbenchmark: per-warp TMA ring of 1664 B records (contiguous range per warp) + per-lane x[col] gather.
// GM 0: no gather; 1: LDG.128 + LDG.64 (one record ahead); 2: one LDG.256; 3: LDG.128 only (probe); 5: x block resident in
// shared memory (cols folded into a window of XW views), LDS gathers.
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
constexpr int RB = 1664, XW = 2500;
template <int GM, int S>
__global__ void __launch_bounds__(256) k_stream(const unsigned char* __restrict__ recs, uint32_t nrec_total, const double* __restrict__ x4, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* ring = smem + (size_t)warp * S * RB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)8 * S * RB) + warp * S;
  double* xs = reinterpret_cast<double*>(smem + 8 * S * RB + 8 * S * 8);
  if (lane == 0) { for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint32_t nwarps = gridDim.x * 8, gw = blockIdx.x * 8 + warp;
  const uint32_t per = (nrec_total + nwarps - 1) / nwarps, lo = gw * per;
  const uint32_t n = lo >= nrec_total ? 0 : min(per, nrec_total - lo);
  auto issue = [&](uint32_t k) { if (lane == 0 && k < n) { mbar_expect_tx(&bars[k % S], RB); tma_load_bulk(ring + (size_t)(k % S) * RB, recs + (size_t)(lo + k) * RB, RB, &bars[k % S]); } };
  for (uint32_t k = 0; k < (uint32_t)S; ++k) issue(k);
  if (GM == 5) {
    for (int i = threadIdx.x; i < XW * 3; i += 256) xs[i] = x4[(i / 3) * 4 + i % 3];
    __syncthreads();
  }
  auto gather = [&](const unsigned char* r, double& x0, double& x1, double& x2) {
    const uint32_t col = reinterpret_cast<const uint32_t*>(r + 1536)[lane];
    if (GM == 0) { x0 = col; x1 = 1; x2 = 2; }
    if (GM == 1) { const double2 a = *reinterpret_cast<const double2*>(x4 + 4 * (size_t)col); x0 = a.x; x1 = a.y; x2 = x4[4 * (size_t)col + 2]; }
    if (GM == 2) { double d; asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(d) : "l"(x4 + 4 * (size_t)col)); }
    if (GM == 3) { const double2 a = *reinterpret_cast<const double2*>(x4 + 4 * (size_t)col); x0 = a.x; x1 = a.y; x2 = 2; }
    if (GM == 5) { const uint32_t c = col % XW; x0 = xs[3 * c]; x1 = xs[3 * c + 1]; x2 = xs[3 * c + 2]; }
  };
  double acc = 0.0;
  if (n == 0) return;
  mbar_wait(&bars[0], 0);
  double x0, x1, x2;
  gather(ring, x0, x1, x2);
  for (uint32_t k = 0; k < n; ++k) {
    const double* r = reinterpret_cast<const double*>(ring + (size_t)(k % S) * RB);
    double n0 = 0, n1 = 0, n2 = 0;
    if (k + 1 < n) { mbar_wait(&bars[(k + 1) % S], ((k + 1) / S) & 1u); gather(ring + (size_t)((k + 1) % S) * RB, n0, n1, n2); }
    const double b0 = r[lane], b1 = r[32 + lane], b2 = r[64 + lane], b3 = r[96 + lane], b4 = r[128 + lane], b5 = r[160 + lane];
    acc += b0 * x0 + b1 * x1 + b2 * x2 + b3 * x0 + b4 * x1 + b5 * x2;
    __syncwarp();
    issue(k + S);
    x0 = n0; x1 = n1; x2 = n2;
  }
  if (acc == 123.456) out[0] = acc;
}
template <int GM, int S>
void run(const unsigned char* d, uint32_t nrec, const double* x4, double* out) {
  const int smem = 8 * S * RB + 8 * S * 8 + (GM == 5 ? XW * 24 : 0);
  cudaFuncSetAttribute(k_stream<GM, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<GM, S>, 256, smem);
  const int grid = 148 * 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k_stream<GM, S><<<grid, 256, smem>>>(d, nrec, x4, out);
  cudaEventRecord(e0);
  const int reps = 500;
  for (int w = 0; w < reps; ++w) k_stream<GM, S><<<grid, 256, smem>>>(d, nrec, x4, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("gather mode %d S %d occ %d: %.2f us  %.0f GB/s %s\n", GM, S, occ, 1e3 * ms / reps, (double)nrec * RB / (ms / reps * 1e-3) / 1e9, err ? cudaGetErrorString(err) : "");
}
int main() {
  const uint32_t nrec = 62500, N = 10000;
  std::vector<unsigned char> h((size_t)nrec * RB, 0);
  srand(1);
  for (uint32_t r = 0; r < nrec; ++r) { uint32_t* c = reinterpret_cast<uint32_t*>(h.data() + (size_t)r * RB + 1536); for (int l = 0; l < 32; ++l) c[l] = (uint32_t)(rand() % N); }
  unsigned char* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  double* x4; cudaMalloc(&x4, N * 32); cudaMemset(x4, 0, N * 32);
  double* out; cudaMalloc(&out, 64);
  run<0, 4>(d, nrec, x4, out);
  run<1, 4>(d, nrec, x4, out);
  run<2, 4>(d, nrec, x4, out);
  run<3, 4>(d, nrec, x4, out);
  run<5, 4>(d, nrec, x4, out);
  run<5, 3>(d, nrec, x4, out);
  run<0, 3>(d, nrec, x4, out);
  run<2, 3>(d, nrec, x4, out);
  return 0;
}
