// Does consumer activity slow the tensor pipe?  MMA warp streams 128x128x16 SS MMAs (accumulator cols 256..383) for a fixed
// number of iterations and reports clk/MMA while NWG consumer warpgroups loop on: MODE 0 nothing (exit), 1 tcgen05.ld.x32 of
// other columns, 2 ld + st (like S -> P), 3 MUFU/FMA math only, 4 ld + math + st (softmax-like), 5 spin on an mbarrier.
#include <cstdio>
#include "common.cuh"
using namespace cb;
namespace cb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 1; } }
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE, int NWG, bool TS>
__global__ void __launch_bounds__(32 + 128 * NWG, 1) probe(long long* out, float* sink, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, never;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_MMA = 4 * NWG;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); fence_barrier_init(); stop = 0; }
  if (warp == W_MMA) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == W_MMA) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, false, false);
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 1024, 3), b0 = umma_smem_desc(smem_u32(smem) + 32768, 16, 1024, 3);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (TS) umma_ts(tm + 256, tm + 480 + (k & 3) * 8, umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
          else umma_ss(tm + 256, umma_desc_add(a0, (k & 1) * 32), umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) *out = t1 - t0;
    stop = 1;
  } else {
    const int q = warp & 3;
    const uint32_t base = tm + (uint32_t(q * 32) << 16) + ((warp >> 2) & 1) * 128;
    float acc = 0.f;
    if (MODE == 5) { while (!stop) { if (mbar_try_wait(&never, 0)) break; } }
    else if (MODE != 0) {
      while (!stop) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          if (MODE == 1 || MODE == 2 || MODE == 4) { tmem_ld32(base + c * 32, r); tmem_ld_wait(); }
          else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(acc + i);
          }
          if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(ex2f(fmaf(__uint_as_float(r[i]), 0.01f, -1.f)));
          }
          if (MODE == 2 || MODE == 4) {
            uint32_t p[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            tmem_st16(base + c * 16, p);
            tmem_st_wait();
          }
          acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
        }
      }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tm, 512);
}

template <int MODE, int NWG, bool TS>
void run(const char* name) {
  long long* out; float* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  auto k = probe<MODE, NWG, TS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 32 + 128 * NWG, 100 * 1024>>>(out, sink, iters);
  k<<<148, 32 + 128 * NWG, 100 * 1024>>>(out, sink, iters);
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s %s warpgroups=%d : %6.1f clk/MMA (nominal 64)  %s\n", name, TS ? "TS" : "SS", NWG, (double)c / (iters * 8.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  run<0, 2, false>("consumers idle (exited)");
  run<5, 2, false>("consumers spin on mbarrier try_wait");
  run<1, 2, false>("consumers: tcgen05.ld loop");
  run<2, 2, false>("consumers: ld + pack + st loop");
  run<3, 2, false>("consumers: MUFU + FMA math only");
  run<4, 2, false>("consumers: ld + exp + pack + st (softmax-like)");
  run<4, 2, true>("consumers: ld + exp + pack + st (softmax-like)");
  run<4, 1, false>("consumers: softmax-like");
  return 0;
}
