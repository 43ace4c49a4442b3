// K2 v2 microbenchmark: block records (1536 B, 6 planes x 32 doubles) and column ids (128 B per record) as SEPARATE streams.
// Per warp: TMA ring of S block records, TMA ring of SC column chunks (CC records each), x[col] gathers issued G records
// ahead with cp.async into an x ring.  The block-record slot is consumed the moment it lands.
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
template <int S, int G, int CC, int SC, bool GATHER>
__global__ void __launch_bounds__(256) k_stream(const unsigned char* __restrict__ recs, const uint32_t* __restrict__ cols, uint32_t nrec_total,
                                                const double* __restrict__ x4, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpBytes = S * RB + SC * CC * 128 + (G + 1) * 1024;
  unsigned char* ring = smem + (size_t)warp * kWarpBytes;
  uint32_t* cring = reinterpret_cast<uint32_t*>(ring + S * RB);
  unsigned char* xring = ring + S * RB + SC * CC * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)8 * kWarpBytes) + warp * (S + SC);
  uint64_t* cbars = bars + S;
  if (lane == 0) { for (int s = 0; s < S + SC; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint32_t nwarps = gridDim.x * 8, gw = blockIdx.x * 8 + warp;
  const uint32_t per = (nrec_total + nwarps - 1) / nwarps, lo = gw * per;
  const uint32_t n = lo >= nrec_total ? 0 : min(per, nrec_total - lo);
  if (n == 0) return;
  const uint32_t nchunk = (n + CC - 1) / CC;
  auto issue = [&](uint32_t k) { if (lane == 0 && k < n) { mbar_expect_tx(&bars[k % S], RB); tma_load_bulk(ring + (size_t)(k % S) * RB, recs + (size_t)(lo + k) * RB, RB, &bars[k % S]); } };
  auto issue_cols = [&](uint32_t j) {
    if (lane == 0 && j < nchunk) {
      const uint32_t nr = min((uint32_t)CC, n - j * CC);
      mbar_expect_tx(&cbars[j % SC], nr * 128);
      tma_load_bulk(cring + (size_t)(j % SC) * CC * 32, cols + ((size_t)lo + (size_t)j * CC) * 32, nr * 128, &cbars[j % SC]);
    }
  };
  auto gather_async = [&](uint32_t k) {  // k < n
    const uint32_t j = k / CC;
    if (k % CC == 0) mbar_wait(&cbars[j % SC], (j / SC) & 1u);
    const uint32_t col = cring[(size_t)(j % SC) * CC * 32 + (k % CC) * 32 + lane];
    if (k % CC == CC - 1 || k == n - 1) { __syncwarp(); issue_cols(j + SC); }
    if (GATHER) {
      unsigned char* d = xring + (size_t)(k % (G + 1)) * 1024 + 16 * lane;
      const unsigned char* g = reinterpret_cast<const unsigned char*>(x4) + 32 * (size_t)col;
      cp_async16(d, g); cp_async16(d + 512, g + 16);
    }
  };
  for (uint32_t j = 0; j < (uint32_t)SC; ++j) issue_cols(j);
  for (uint32_t k = 0; k < (uint32_t)S; ++k) issue(k);
  for (int g = 0; g < G; ++g) { if ((uint32_t)g < n) gather_async(g); cp_async_commit(); }
  double acc = 0.0;
  for (uint32_t k = 0; k < n; ++k) {
    if (k + G < n) gather_async(k + G);
    cp_async_commit();
    cp_async_wait<G>();
    double x0 = 1, x1 = 2, x2 = 3;
    if (GATHER) {
      const unsigned char* xs = xring + (size_t)(k % (G + 1)) * 1024 + 16 * lane;
      const double2 a = *reinterpret_cast<const double2*>(xs); x0 = a.x; x1 = a.y; x2 = *reinterpret_cast<const double*>(xs + 512);
    }
    mbar_wait(&bars[k % S], (k / S) & 1u);
    const double* r = reinterpret_cast<const double*>(ring + (size_t)(k % S) * RB);
    const double b0 = r[lane], b1 = r[32 + lane], b2 = r[64 + lane], b3 = r[96 + lane], b4 = r[128 + lane], b5 = r[160 + lane];
    acc += b0 * x0 + b1 * x1 + b2 * x2 + b3 * x0 + b4 * x1 + b5 * x2;
    __syncwarp();
    issue(k + S);
  }
  if (acc == 123.456) out[0] = acc;
}
template <int S, int G, int CC, int SC, bool GATHER>
void run(const unsigned char* d, const uint32_t* cols, uint32_t nrec, const double* x4, double* out) {
  const int smem = 8 * (S * RB + SC * CC * 128 + (G + 1) * 1024) + 8 * (S + SC) * 8;
  cudaFuncSetAttribute(k_stream<S, G, CC, SC, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<S, G, CC, SC, GATHER>, 256, smem);
  const int grid = 148 * 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k_stream<S, G, CC, SC, GATHER><<<grid, 256, smem>>>(d, cols, nrec, x4, out);
  cudaEventRecord(e0);
  const int reps = 500;
  for (int w = 0; w < reps; ++w) k_stream<S, G, CC, SC, GATHER><<<grid, 256, smem>>>(d, cols, nrec, x4, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("S %d G %d CC %d SC %d gather %d smem %d occ %d: %.2f us  %.0f GB/s (blocks+cols) %s\n", S, G, CC, SC, (int)GATHER, smem, occ, 1e3 * ms / reps,
         (double)nrec * (RB + 128) / (ms / reps * 1e-3) / 1e9, err ? cudaGetErrorString(err) : "");
}
int main() {
  const uint32_t nrec = 62500, N = 10000;
  std::vector<uint32_t> hc((size_t)nrec * 32);
  srand(1);
  for (auto& c : hc) c = (uint32_t)(rand() % N);
  unsigned char* d; cudaMalloc(&d, (size_t)nrec * RB); cudaMemset(d, 0, (size_t)nrec * RB);
  uint32_t* cols; cudaMalloc(&cols, hc.size() * 4); cudaMemcpy(cols, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
  double* x4; cudaMalloc(&x4, N * 32); cudaMemset(x4, 0, N * 32);
  double* out; cudaMalloc(&out, 64);
  run<4, 2, 8, 2, false>(d, cols, nrec, x4, out);
  run<4, 1, 8, 2, true>(d, cols, nrec, x4, out);
  run<4, 2, 8, 2, true>(d, cols, nrec, x4, out);
  run<4, 3, 8, 2, true>(d, cols, nrec, x4, out);
  run<4, 4, 8, 3, true>(d, cols, nrec, x4, out);
  run<4, 6, 8, 3, true>(d, cols, nrec, x4, out);
  run<5, 3, 8, 2, true>(d, cols, nrec, x4, out);
  run<3, 3, 8, 2, true>(d, cols, nrec, x4, out);
  run<4, 3, 4, 3, true>(d, cols, nrec, x4, out);
  run<4, 3, 16, 2, true>(d, cols, nrec, x4, out);
  return 0;
}
