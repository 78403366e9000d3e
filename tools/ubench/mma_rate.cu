// tcgen05.mma throughput probe (B200): cycles per MMA for M=128, N in {32..256}, SS vs TS (A from TMEM), accumulating into
// ONE accumulator (dependent chain, like a K loop) or alternating between two.  Operands are whatever is in smem / TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../chadavit_b200/csrc -I ../../include -o mma_rate mma_rate.cu
#include <cstdio>
#include "common.cuh"
using namespace cb;
namespace cb { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 1; } }

template <int N, bool TS, int NACC, int SWZ, int CE = 0>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
    constexpr int ROWB = SWZ == 3 ? 128 : 64;   // bytes per smem row (one swizzle chunk)
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 8 * ROWB, SWZ), b0 = umma_smem_desc(smem_u32(smem) + 32768, 16, 8 * ROWB, SWZ);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t d = tm + ((NACC > 1 && (k & 1)) ? 256 : 0);
          if (TS) umma_ts(d, tm + 480 + (k & 3) * 8, umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
          else umma_ss(d, umma_desc_add(a0, (k & 1) * 32), umma_desc_add(b0, (k & 1) * 32), idesc, 1u);
          if (CE > 0 && (k % CE) == CE - 1) tc_commit(&bar2[k]);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// alternating groups like the attention forward: G1 x TS (N1, A in TMEM, accumulator 1) then G2 x SS/TS (N2, accumulator 0)
template <int N1, int G1, int N2, int G2, bool TS2, bool B1_MN>
__global__ void __launch_bounds__(128, 1) probe_alt(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    constexpr uint32_t id1 = umma_idesc_bf16(128, N1, false, B1_MN), id2 = umma_idesc_bf16(128, N2, false, false);
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 512, 2), b0 = umma_smem_desc(smem_u32(smem) + 32768, 16, 512, 2);
    const uint64_t bm = umma_smem_desc(smem_u32(smem) + 32768, 4096, 512, 2);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < G1; ++k) umma_ts(tm + 256, tm + 480 + (k & 3) * 8, umma_desc_add(B1_MN ? bm : b0, (k & 1) * (B1_MN ? 1024 : 32)), id1, 1u);
#pragma unroll
        for (int k = 0; k < G2; ++k) {
          if (TS2) umma_ts(tm, tm + 448 + (k & 3) * 8, umma_desc_add(b0, (k & 1) * 32), id2, 1u);
          else umma_ss(tm, umma_desc_add(a0, (k & 1) * 32), umma_desc_add(b0, (k & 1) * 32), id2, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
template <int N1, int G1, int N2, int G2, bool TS2, bool B1_MN>
void run_alt(const char* name) {
  long long* out; cudaMalloc(&out, 8);
  const int iters = 512;
  auto k = probe_alt<N1, G1, N2, G2, TS2, B1_MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 128, 100 * 1024>>>(out, iters);
  k<<<148, 128, 100 * 1024>>>(out, iters);
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-50s %7.1f clk/iteration (nominal %d)  %s\n", name, (double)c / iters, G1 * 128 * N1 / 256 + G2 * 128 * N2 / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

template <int N, bool TS, int NACC, int SWZ, int CE = 0>
void run(const char* name) {
  long long* out; cudaMalloc(&out, 8);
  const int iters = 512;
  cudaFuncSetAttribute(probe<N, TS, NACC, SWZ, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<N, TS, NACC, SWZ, CE><<<148, 128, 100 * 1024>>>(out, iters);
  probe<N, TS, NACC, SWZ, CE><<<148, 128, 100 * 1024>>>(out, iters);
  long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s N=%3d  %7.1f clk/MMA  (nominal %3d)  %s\n", name, N, (double)c / (iters * 8.0), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  run_alt<96, 4, 64, 6, false, false>("4 x TS N=96 | 6 x SS N=64 (K-major B)");
  run_alt<96, 4, 64, 6, false, true>("4 x TS N=96 (MN-major B) | 6 x SS N=64");
  run_alt<96, 8, 128, 6, false, true>("8 x TS N=96 (MN-major B) | 6 x SS N=128");
  run_alt<96, 3, 48, 6, true, true>("3 x TS N=96 (MN-major B) | 6 x TS N=48");
  run_alt<96, 8, 96, 8, true, true>("8 x TS N=96 (MN-major B) | 8 x TS N=96");
  run<96, true, 1, 2>("TS sw64 homogeneous");
  return 0;
}
