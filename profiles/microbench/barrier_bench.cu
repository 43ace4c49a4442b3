// Grid-barrier + deterministic 2-double reduction microbenchmark (round 2): what does one barrier of the persistent PCG
// kernel cost with nothing else in the way, and which implementation is cheapest?
//   mode 0  cooperative groups grid.sync() + per-block slots re-read by every block (round-1 grid_bar_sum2)
//   mode 1  atomic arrive counter; the LAST arriver adds the slots in fixed order and publishes {total, seq}; others poll
//   mode 2  no atomic: every block publishes {value, seq} slots, warp 0 of every block polls all slots (fused barrier+sum)
//   mode 3  atomic arrive + generation flag (hand-rolled grid.sync), slots re-read by every block
// Every variant gives bit-identical totals in every block (fixed summation order).  `work` = number of 8-byte stores per
// thread issued just before the barrier (the vector phase of a CG step leaves ~13 stores in flight).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
namespace cg = cooperative_groups;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_add_release(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
struct alignas(16) Tagged { double v; unsigned long long seq; };
__device__ __forceinline__ Tagged ld_tagged(const Tagged* p) {
  Tagged t;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<unsigned long long*>(&t.v)), "=l"(t.seq) : "l"(p) : "memory");
  return t;
}
__device__ __forceinline__ void st_tagged(Tagged* p, double v, unsigned long long seq) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(seq) : "memory");
}

struct Bufs {
  double* slots;        // [2][grid][2]
  Tagged* tslots;       // [2][grid][2]
  Tagged* totals;       // [2][2]
  unsigned* counter;    // arrive counter
  unsigned* gen;        // generation flag
  double* scratch;      // work stores
  double* out;
};

template <int MODE>
__global__ void __launch_bounds__(512) k_bar(Bufs B, int iters, int work) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[2 * 16 + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const unsigned G = gridDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  unsigned seq = 0;
  for (int it = 0; it < iters; ++it) {
    for (int w = 0; w < work; ++w) B.scratch[((size_t)w * G + blockIdx.x) * blockDim.x + threadIdx.x] = acc0 + w;
    double v0 = 1.0 + 1e-3 * threadIdx.x + acc0 * 1e-9, v1 = 0.5 + blockIdx.x;
    v0 = warp_sum(v0); v1 = warp_sum(v1);
    if (lane == 0) { sm[2 * warp] = v0; sm[2 * warp + 1] = v1; }
    __syncthreads();
    ++seq;
    double t0 = 0.0, t1 = 0.0;
    if (MODE == 0 || MODE == 3) {
      double* set = B.slots + (size_t)(seq & 1u) * 2 * G;
      if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0;
        for (int w = 0; w < nwarp; ++w) { b0 += sm[2 * w]; b1 += sm[2 * w + 1]; }
        __stcg(set + 2 * blockIdx.x, b0); __stcg(set + 2 * blockIdx.x + 1, b1);
      }
      if (MODE == 0) grid.sync();
      else {
        __syncthreads();
        if (threadIdx.x == 0) {
          const unsigned old = atom_add_release(B.counter, 1u);
          if (old == seq * G - 1) st_release(B.gen, seq);
          else while (ld_acquire(B.gen) != seq) { }
        }
        __syncthreads();
      }
      if (warp == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (unsigned b = lane; b < G; b += 32) { s0 += __ldcg(set + 2 * b); s1 += __ldcg(set + 2 * b + 1); }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sm[2 * 16] = s0; sm[2 * 16 + 1] = s1; }
      }
      __syncthreads();
      t0 = sm[2 * 16]; t1 = sm[2 * 16 + 1];
    } else if (MODE == 1) {
      double* set = B.slots + (size_t)(seq & 1u) * 2 * G;
      Tagged* tot = B.totals + (size_t)(seq & 1u) * 2;
      if (warp == 0) {
        unsigned last = 0;
        if (lane == 0) {
          double b0 = 0.0, b1 = 0.0;
          for (int w = 0; w < nwarp; ++w) { b0 += sm[2 * w]; b1 += sm[2 * w + 1]; }
          __stcg(set + 2 * blockIdx.x, b0); __stcg(set + 2 * blockIdx.x + 1, b1);
          const unsigned old = atom_add_release(B.counter, 1u);
          last = (old == seq * G - 1) ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          __threadfence();
          double s0 = 0.0, s1 = 0.0;
          for (unsigned b = lane; b < G; b += 32) { s0 += __ldcg(set + 2 * b); s1 += __ldcg(set + 2 * b + 1); }
          s0 = warp_sum(s0); s1 = warp_sum(s1);
          if (lane == 0) st_tagged(tot, s0, seq);
          if (lane == 1) st_tagged(tot + 1, s1, seq);
          t0 = s0; t1 = s1;
        } else {
          Tagged a;
          if (lane < 2) { do { a = ld_tagged(tot + lane); } while (a.seq != seq); }
          t0 = __shfl_sync(0xffffffffu, a.v, 0); t1 = __shfl_sync(0xffffffffu, a.v, 1);
        }
        if (lane == 0) { sm[2 * 16] = t0; sm[2 * 16 + 1] = t1; }
        __threadfence();  // acquire side: data written by other blocks before their arrive is visible after this
      }
      __syncthreads();
      t0 = sm[2 * 16]; t1 = sm[2 * 16 + 1];
    } else {  // MODE 2
      Tagged* set = B.tslots + (size_t)(seq & 1u) * 2 * G;
      if (warp == 0) {
        if (lane == 0) {
          double b0 = 0.0, b1 = 0.0;
          for (int w = 0; w < nwarp; ++w) { b0 += sm[2 * w]; b1 += sm[2 * w + 1]; }
          __threadfence();
          st_tagged(set + 2 * blockIdx.x, b0, seq); st_tagged(set + 2 * blockIdx.x + 1, b1, seq);
        }
        double s0 = 0.0, s1 = 0.0;
        for (unsigned b = lane; b < G; b += 32) {
          Tagged a0, a1;
          do { a0 = ld_tagged(set + 2 * b); } while (a0.seq != seq);
          do { a1 = ld_tagged(set + 2 * b + 1); } while (a1.seq != seq);
          s0 += a0.v; s1 += a1.v;
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sm[2 * 16] = s0; sm[2 * 16 + 1] = s1; }
        __threadfence();
      }
      __syncthreads();
      t0 = sm[2 * 16]; t1 = sm[2 * 16 + 1];
    }
    acc0 += t0; acc1 += t1;
  }
  if (threadIdx.x == 0) { B.out[2 * blockIdx.x] = acc0; B.out[2 * blockIdx.x + 1] = acc1; }
}

template <int MODE>
float run(Bufs B, int grid, int block, int iters, int work, double* check) {
  cudaMemset(B.counter, 0, 8); cudaMemset(B.gen, 0, 8);
  cudaMemset(B.tslots, 0, sizeof(Tagged) * 4 * grid); cudaMemset(B.totals, 0, sizeof(Tagged) * 4);
  void* args[] = {&B, &iters, &work};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int warm = 10;
  void* wargs[] = {&B, &warm, &work};
  cudaLaunchCooperativeKernel((void*)k_bar<MODE>, dim3(grid), dim3(block), wargs, 0, 0);
  cudaMemset(B.counter, 0, 8); cudaMemset(B.gen, 0, 8);
  cudaMemset(B.tslots, 0, sizeof(Tagged) * 4 * grid); cudaMemset(B.totals, 0, sizeof(Tagged) * 4);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  cudaLaunchCooperativeKernel((void*)k_bar<MODE>, dim3(grid), dim3(block), args, 0, 0);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("mode %d failed: %s\n", MODE, cudaGetErrorString(err)); return -1.f; }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double h[2];
  cudaMemcpy(h, B.out, 16, cudaMemcpyDeviceToHost);
  *check = h[0];
  return ms * 1e3f / iters;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  Bufs B;
  const int maxgrid = sms * 2;
  cudaMalloc(&B.slots, sizeof(double) * 4 * maxgrid);
  cudaMalloc(&B.tslots, sizeof(Tagged) * 4 * maxgrid);
  cudaMalloc(&B.totals, sizeof(Tagged) * 4);
  cudaMalloc(&B.counter, 8); cudaMalloc(&B.gen, 8);
  cudaMalloc(&B.scratch, sizeof(double) * 16 * (size_t)maxgrid * 512);
  cudaMalloc(&B.out, sizeof(double) * 2 * maxgrid);
  printf("SMs %d, %d barriers per launch; us per barrier (incl. a warp-sum + 2 __syncthreads of set-up)\n", sms, iters);
  const int cfg[3][2] = {{sms * 2, 256}, {sms, 512}, {sms, 256}};
  for (auto& c : cfg)
    for (int work : {0, 13}) {
      double c0, c1, c2, c3;
      const float t0 = run<0>(B, c[0], c[1], iters, work, &c0);
      const float t1 = run<1>(B, c[0], c[1], iters, work, &c1);
      const float t2 = run<2>(B, c[0], c[1], iters, work, &c2);
      const float t3 = run<3>(B, c[0], c[1], iters, work, &c3);
      printf("grid %3d x %3d, %2d stores: cg.sync+slots %.2f | last-arriver reduce %.2f | all-poll tagged slots %.2f | atomic+gen+slots %.2f   (totals equal: %d)\n",
             c[0], c[1], work, t0, t1, t2, t3, (int)(c0 == c1 && c1 == c2 && c2 == c3));
    }
  return 0;
}
