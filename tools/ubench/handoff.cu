// Round-trip probe: MMA warp <-> consumer warpgroup handoff through tcgen05.commit / mbarrier, like the attention kernels.
//   chain c (c < NCHAIN): MMA warp waits p[c] -> issues G MMAs (128xNx16 SS) into accumulator c -> commits s[c];
//   consumer warpgroup c (128 threads) waits s[c] -> optional TMEM read -> arrives p[c] (128 arrivals).
// Reports cycles per round of all chains; ideal = NCHAIN * G * nominal.
#include <cstdio>
#include "common.cuh"
using namespace cb;
namespace cb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 1; } }

template <int N, int G, int NCHAIN, int READ, int WORK, int VAR = 0>
__global__ void __launch_bounds__(32 + 128 * NCHAIN, 1) probe(long long* out, float* sink, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t s_bar[4], p_bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_MMA = 4 * NCHAIN;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) { mbar_init(&s_bar[i], 1); mbar_init(&p_bar[i], 128); } fence_barrier_init(); }
  if (warp == W_MMA) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == W_MMA) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 1024, 3), b0 = umma_smem_desc(smem_u32(smem) + 32768, 16, 1024, 3);
    // prologue: first group of every chain
    for (int c = 0; c < NCHAIN; ++c) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < G; ++k) umma_ss(tm + c * 128, umma_desc_add(a0, (k & 1) * 32), umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
        tc_commit(&s_bar[c]);
      }
      __syncwarp();
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < NCHAIN; ++c) {
        if (VAR & 2) { if (lane == 0) mbar_wait(&p_bar[c], it & 1); __syncwarp(); } else mbar_wait(&p_bar[c], it & 1);
        if (!(VAR & 1)) tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < G; ++k) umma_ss(tm + c * 128, umma_desc_add(a0, (k & 1) * 32), umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
          tc_commit(&s_bar[c]);
        }
        __syncwarp();
      }
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) *out = t1 - t0;
  } else {
    const int c = warp >> 2, q = warp & 3;
    float acc = 0.f;
    for (int it = 0; it <= iters; ++it) {
      mbar_wait(&s_bar[c], it & 1);
      if (!(VAR & 4)) tc_fence_after();
      if (READ) {
        uint32_t r[32];
        tmem_ld32(tm + (uint32_t(q * 32) << 16) + c * 128, r);
        tmem_ld_wait();
        acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
      }
      if (WORK) {   // ~WORK cycles of dependent FMA work
        float x = acc;
        for (int k = 0; k < WORK / 4; ++k) x = fmaf(x, 1.0001f, 0.5f);
        acc = x;
      }
      if (!(VAR & 4)) tc_fence_before();
      if (VAR & 8) { __syncwarp(); if (lane == 0 && it < iters) for (int k = 0; k < 32; ++k) mbar_arrive(&p_bar[c]); }
      else if (it < iters) mbar_arrive(&p_bar[c]);
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tm, 512);
}

template <int N, int G, int NCHAIN, int READ, int WORK, int VAR = 0>
void run(const char* name) {
  long long* out; float* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 256;
  auto k = probe<N, G, NCHAIN, READ, WORK, VAR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 32 + 128 * NCHAIN, 100 * 1024>>>(out, sink, iters);
  k<<<148, 32 + 128 * NCHAIN, 100 * 1024>>>(out, sink, iters);
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double per = (double)c / iters, ideal = (double)NCHAIN * G * 128 * N / 256;
  printf("%-40s N=%3d G=%2d chains=%d read=%d work=%4d : %7.1f clk/round, MMA math %6.0f -> overhead %6.1f per handoff %s\n", name, N, G, NCHAIN, READ, WORK, per,
         ideal, (per - ideal) / NCHAIN, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  run<128, 4, 4, 1, 0, 0>("4 chains G=4 baseline");
  run<128, 4, 4, 1, 0, 1>("  no tc_fence_after in MMA warp");
  run<128, 4, 4, 1, 0, 2>("  lane-0 poll + syncwarp");
  run<128, 4, 4, 1, 0, 4>("  consumer without tcgen05 fences");
  run<128, 4, 4, 0, 0, 0>("  no TMEM read");
  run<128, 4, 4, 0, 0, 7>("  no read, no fences, lane-0 poll");
  run<128, 4, 4, 1, 0, 8>("  one lane arrives 32x");
  run<128, 2, 4, 1, 0, 0>("4 chains G=2");
  run<128, 8, 4, 1, 0, 0>("4 chains G=8");
  run<128, 16, 2, 1, 0, 0>("2 chains G=16");
  return 0;
}
