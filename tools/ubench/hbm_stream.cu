// HBM stream probe: pure write, pure read and copy with 256-bit accesses, at several grid sizes / block sizes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/hbm_stream tools/ubench/hbm_stream.cu && tools/ubench/hbm_stream
// Round 1 placed write-heavy kernels on a "3.9 TB/s" roofline taken from torch.fill_ and read-heavy ones on "4.3 TB/s" from a torch
// reduction; round 2 found the read figure to be an artefact (under-filled grids).  This probe measures what a plain stream reaches.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void k_write(uint4* __restrict__ p, size_t n32) {   // n32: number of 32-byte units
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 2 * i), "r"(0x3f803f80u) : "memory");
  }
}
__global__ void k_read(const uint4* __restrict__ p, size_t n32, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t r[8];
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + 2 * i));
    acc ^= r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
  }
  if (acc == 0x12345678u) *out = acc;
}
__global__ void k_copy(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n32) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t r[8];
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(s + 2 * i));
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d + 2 * i), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  }
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)2 << 30;     // 2 GiB per buffer (>> 126 MB L2)
  uint4 *a, *b; uint32_t* out;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 4);
  cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
  const size_t n32 = bytes / 32;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("# 2 GiB buffers, 256-bit accesses; TB/s counts bytes read + bytes written\n");
  const int grids[] = {sms, 2 * sms, 4 * sms, 8 * sms, 16 * sms, 32 * sms};
  const int blocks[] = {256, 512, 1024};
  for (int bs : blocks)
    for (int g : grids) {
      float w = time_ms([&] { k_write<<<g, bs>>>(a, n32); }, 10);
      float r = time_ms([&] { k_read<<<g, bs>>>(a, n32, out); }, 10);
      float c = time_ms([&] { k_copy<<<g, bs>>>(a, b, n32); }, 10);
      printf("block %4d grid %5d (%2d / SM): write %5.2f TB/s   read %5.2f TB/s   copy %5.2f TB/s\n", bs, g, g / sms, bytes / w / 1e9, bytes / r / 1e9, 2.0 * bytes / c / 1e9);
    }
  float ms = time_ms([&] { cudaMemsetAsync(a, 0, bytes); }, 10);
  printf("cudaMemsetAsync: %5.2f TB/s;  ", bytes / ms / 1e9);
  ms = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 10);
  printf("cudaMemcpyAsync D2D: %5.2f TB/s\n", 2.0 * bytes / ms / 1e9);
  return 0;
}
