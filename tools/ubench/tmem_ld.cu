// tcgen05.ld latency / throughput probe under concurrent MMA load (B200).
//   MMA warp streams 128x128x16 SS MMAs into accumulator 1 (cols 256..383) forever (until consumers finish);
//   4 consumer warps read accumulator 0 (cols 0..127) of their lane quarter: MODE 0 = 4 x (ld.x32, wait); MODE 1 = 4 x ld.x32 then
//   one wait; MODE 2 = 8 x ld.x16 then one wait; MODE 3 = 1 x (ld.x32, wait).   Reports clk per 128-column read per warp.
#include <cstdio>
#include "common.cuh"
using namespace cb;
namespace cb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 1; } }

template <int MODE, int MMA_ON, int NWG>
__global__ void __launch_bounds__(32 + 128 * NWG, 1) probe(long long* out, float* sink, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_MMA = 4 * NWG;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == W_MMA) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == W_MMA) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, false, false);
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 1024, 3), b0 = umma_smem_desc(smem_u32(smem) + 32768, 16, 1024, 3);
    if (MMA_ON) {
      int n = 0;
      while (!stop && n < 100000) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tm + 256, umma_desc_add(a0, (k & 1) * 32), umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
        }
        __syncwarp();
        ++n;
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
    }
  } else {
    const int q = warp & 3;
    const uint32_t base = tm + (uint32_t(q * 32) << 16) + (warp >> 2) * 128 * 0;
    float acc = 0.f;
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0 || MODE == 3) {
#pragma unroll
        for (int c = 0; c < (MODE == 3 ? 1 : 4); ++c) {
          uint32_t r[32];
          tmem_ld32(base + c * 32, r);
          tmem_ld_wait();
          acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
        }
      } else if (MODE == 1) {
        uint32_t r[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(base + c * 32, r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += __uint_as_float(r[c][0]) + __uint_as_float(r[c][31]);
      } else {
        uint32_t r[8][16];
#pragma unroll
        for (int c = 0; c < 8; ++c) tmem_ld16(base + c * 16, r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += __uint_as_float(r[c][0]) + __uint_as_float(r[c][15]);
      }
    }
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (warp == 0 && lane == 0 && blockIdx.x == 0) *out = t1 - t0;
    asm volatile("bar.sync 1, %0;" ::"r"(128 * NWG));
    if (threadIdx.x == 0) stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tm, 512);
}

template <int MODE, int MMA_ON, int NWG>
void run(const char* name) {
  long long* out; float* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 200;
  auto k = probe<MODE, MMA_ON, NWG>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 32 + 128 * NWG, 100 * 1024>>>(out, sink, iters);
  k<<<148, 32 + 128 * NWG, 100 * 1024>>>(out, sink, iters);
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-46s mma=%d warpgroups=%d : %8.1f clk per %d-column row read per warp  %s\n", name, MMA_ON, NWG, (double)c / iters, MODE == 3 ? 32 : 128,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  run<3, 0, 1>("1 x (ld.x32, wait)");
  run<3, 1, 1>("1 x (ld.x32, wait)");
  run<0, 0, 1>("4 x (ld.x32, wait)");
  run<0, 1, 1>("4 x (ld.x32, wait)");
  run<1, 0, 1>("4 x ld.x32, one wait");
  run<1, 1, 1>("4 x ld.x32, one wait");
  run<2, 1, 1>("8 x ld.x16, one wait");
  run<0, 1, 2>("4 x (ld.x32, wait), 2 warpgroups");
  run<1, 1, 2>("4 x ld.x32, one wait, 2 warpgroups");
  return 0;
}
