// Pipe-throughput probes for B200 (sm_100a): cycles per warp-instruction per SMSP for MUFU.EX2, F2FP (bf16x2 pack), FFMA,
// FMNMX and a mix; one or two warps per SMSP.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.01f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) x[i] = ex2f(x[i]);                                   // MUFU only
      if (MODE == 1) { acc ^= pack(x[i], x[(i + 1) & 7]); x[i] += 1.0f; } // F2FP + FADD
      if (MODE == 2) x[i] = fmaf(x[i], 1.0001f, 0.5f);                    // FFMA only
      if (MODE == 3) { x[i] = ex2f(fmaf(x[i], 0.5f, -1.f)); if (i & 1) acc ^= pack(x[i], x[i - 1]); }  // FFMA + MUFU + 0.5 F2FP
      if (MODE == 4) { float y = ex2f(fmaf(x[i], 0.5f, -1.f)); x[i] = y + x[(i + 1) & 7]; }            // FFMA + MUFU + FADD
      if (MODE == 5) { x[i] = x[i] + 1.0f; }                              // FADD only
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 0xff);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* cyc;
  const int threads = 128 * warps_per_smsp, iters = 4096;
  cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&cyc, 8);
  probe<MODE><<<148, threads>>>(out, cyc, iters);
  probe<MODE><<<148, threads>>>(out, cyc, iters);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps/SMSP=%d  cycles per loop-body instr-group per warp: %.2f  (per SMSP per warp-iteration of 8: %.1f)\n", name, warps_per_smsp,
         (double)c / (iters * 8.0), (double)c / iters);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("MUFU.EX2", w);
    run<1>("F2FP.bf16x2 + FADD", w);
    run<2>("FFMA", w);
    run<5>("FADD", w);
    run<3>("FFMA + MUFU + 0.5 F2FP", w);
    run<4>("FFMA + MUFU + FADD", w);
  }
  return 0;
}
