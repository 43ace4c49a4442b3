// Stream-pattern microbenchmark: how fast can 148 SMs pull R records of RB bytes with per-warp TMA rings, as a function of
// how the records are dealt to warps?  mode 0: contiguous range per warp (current K2); mode 1: tiles of (8 warps x T records)
// dealt round-robin to blocks; mode 2: plain coalesced LDG grid-stride (reference).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int RB, int S>
__global__ void __launch_bounds__(256) k_stream(const unsigned char* __restrict__ recs, uint32_t nrec_total, int mode, int T, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* ring = smem + (size_t)warp * S * RB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)8 * S * RB) + warp * S;
  if (lane == 0) { for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const uint32_t nwarps = gridDim.x * 8, gw = blockIdx.x * 8 + warp;
  // record index of this warp's k-th record
  uint32_t per = (nrec_total + nwarps - 1) / nwarps;
  auto rec_of = [&](uint32_t k) -> uint32_t {
    if (mode == 0) return gw * per + k;
    // tiles of 8*T records; tile index = round * gridDim.x + blockIdx.x; inside a tile warp w owns T consecutive records
    const uint32_t round = k / T, r = k % T;
    return (round * gridDim.x + blockIdx.x) * (8 * T) + warp * T + r;
  };
  uint32_t n = 0;
  if (mode == 0) { const uint32_t lo = gw * per; n = lo >= nrec_total ? 0 : min(per, nrec_total - lo); }
  else { per = (per + T - 1) / T * T; n = per; }
  auto valid = [&](uint32_t k) { return k < n && rec_of(k) < nrec_total; };
  auto issue = [&](uint32_t k) { if (lane == 0 && valid(k)) { mbar_expect_tx(&bars[k % S], RB); tma_load_bulk(ring + (size_t)(k % S) * RB, recs + (size_t)rec_of(k) * RB, RB, &bars[k % S]); } };
  for (uint32_t k = 0; k < (uint32_t)S; ++k) issue(k);
  double acc = 0.0;
  for (uint32_t k = 0; k < n; ++k) {
    if (!valid(k)) break;
    mbar_wait(&bars[k % S], (k / S) & 1u);
    const double* r = reinterpret_cast<const double*>(ring + (size_t)(k % S) * RB);
#pragma unroll
    for (int j = 0; j < RB / 256; ++j) acc += r[j * 32 + lane];
    __syncwarp();
    issue(k + S);
  }
  if (acc == 123.456) out[0] = acc;
}
__global__ void __launch_bounds__(256) k_ldg(const double4* __restrict__ p, size_t n4, double* out) {
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { const double4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
  if (acc == 123.456) out[0] = acc;
}
template <int RB, int S>
void run(const unsigned char* d, uint32_t nrec, int mode, int T, int blocks_per_sm, double* out, const char* tag) {
  const int smem = 8 * S * RB + 8 * S * 8;
  cudaFuncSetAttribute(k_stream<RB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<RB, S>, 256, smem);
  if (occ < blocks_per_sm) { printf("%s RB %d S %d bps %d: occupancy only %d\n", tag, RB, S, blocks_per_sm, occ); return; }
  const int grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k_stream<RB, S><<<grid, 256, smem>>>(d, nrec, mode, T, out);
  cudaEventRecord(e0);
  const int reps = 500;
  for (int w = 0; w < reps; ++w) k_stream<RB, S><<<grid, 256, smem>>>(d, nrec, mode, T, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("%s mode %d T %2d RB %4d S %d blocks/SM %d: %.2f us  %.0f GB/s %s\n", tag, mode, T, RB, S, blocks_per_sm, 1e3 * ms / reps, (double)nrec * RB / (ms / reps * 1e-3) / 1e9, err ? cudaGetErrorString(err) : "");
}
int main(int argc, char** argv) {
  const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 104) << 20;
  unsigned char* d; cudaMalloc(&d, bytes + (1 << 20)); cudaMemset(d, 0, bytes + (1 << 20));
  double* out; cudaMalloc(&out, 64);
  // reference: LDG
  { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int g : {148 * 2, 148 * 4, 148 * 8}) {
      for (int w = 0; w < 3; ++w) k_ldg<<<g, 256>>>((const double4*)d, bytes / 32, out);
      cudaEventRecord(e0); for (int w = 0; w < 500; ++w) k_ldg<<<g, 256>>>((const double4*)d, bytes / 32, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); printf("ldg grid %d: %.2f us %.0f GB/s\n", g, 1e3 * ms / 500, bytes / (ms / 500 * 1e-3) / 1e9); } }
  const uint32_t n1664 = bytes / 1664, n1536 = bytes / 1536, n3328 = bytes / 3328, n6656 = bytes / 6656;
  run<1664, 4>(d, n1664, 0, 1, 2, out, "contig");
  run<1664, 6>(d, n1664, 0, 1, 2, out, "contig");
  run<1664, 3>(d, n1664, 0, 1, 4, out, "contig");
  for (int T : {1, 2, 4, 8}) run<1664, 4>(d, n1664, 1, T, 2, out, "tiled ");
  for (int T : {1, 2, 4}) run<1664, 6>(d, n1664, 1, T, 2, out, "tiled ");
  for (int T : {1, 4}) run<1664, 3>(d, n1664, 1, T, 4, out, "tiled ");
  run<1536, 4>(d, n1536, 0, 1, 2, out, "contig");
  run<1536, 4>(d, n1536, 1, 4, 2, out, "tiled ");
  run<3328, 3>(d, n3328, 0, 1, 2, out, "contig");
  run<3328, 3>(d, n3328, 1, 2, 2, out, "tiled ");
  run<6656, 2>(d, n6656, 0, 1, 2, out, "contig");
  run<6656, 2>(d, n6656, 1, 1, 2, out, "tiled ");
  return 0;
}
