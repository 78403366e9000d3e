// Feature preparation for the weighted k-NN evaluation (src/utils/knn.py:96-177).
//
// The similarity matrix test . train^T decides neighbour RANKS, so bf16 operands are not good enough; fp32 accuracy is
// recovered on the bf16 tensor cores by splitting every fp32 value x into hi = bf16(x), lo = bf16(x - hi) and running ONE
// tcgen05 GEMM over a three times longer K:   [a_hi | a_hi | a_lo] . [b_hi | b_lo | b_hi]^T = a_hi b_hi + a_hi b_lo + a_lo b_hi,
// i.e. a.b up to the dropped lo.lo term (2^-16 relative, fp32 accumulation).  This kernel writes that operand layout and,
// for the cosine distance, folds F.normalize (x / max(||x||, 1e-12), knn.py:114-116) into the same pass.  One warp per row;
// also returns ||x||^2 of the (normalised or raw) row for the euclidean distance.
#include "common.cuh"
#include "chadavit_b200.h"

namespace cb {

__global__ void split3_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, float* __restrict__ sqnorm, int rows,
                                   int D, int role_b, int normalize) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long)row * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = xr[c]; ss += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float scale = normalize ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
  __nv_bfloat16* o = out + (long)row * 3 * D;
  float s2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = xr[c] * scale;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    s2 += v * v;
    o[c] = hi;
    o[D + c] = role_b ? lo : hi;
    o[2 * D + c] = role_b ? hi : lo;
  }
  if (sqnorm) {
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o2);
    if (lane == 0) sqnorm[row] = s2;
  }
}

// sim[i, j] = 1 / (sqrt(max(|a_i|^2 + |b_j|^2 - 2 a_i.b_j, 0)) + eps)   (knn.py:141; in place on the dot products)
__global__ void inv_euclid_kernel(float* __restrict__ dots, const float* __restrict__ na, const float* __restrict__ nb, int M, int N, int ld,
                                  float eps) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)M * N) return;
  const int r = (int)(i / N), c = (int)(i % N);
  float* p = dots + (long)r * ld + c;
  const float d2 = fmaxf(na[r] + nb[c] - 2.f * *p, 0.f);
  *p = 1.f / (sqrtf(d2) + eps);
}

}  // namespace cb

using namespace cb;

extern "C" int cb_split_bf16x3(const float* x, void* out, float* sqnorm, int rows, int D, int role_b, int normalize, void* stream) {
  CB_CHECK(rows > 0 && D > 0, "split_bf16x3: rows=%d D=%d", rows, D);
  split3_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(out), sqnorm, rows,
                                                                                         D, role_b, normalize);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_inv_euclid(float* dots, const float* sqnorm_a, const float* sqnorm_b, int M, int N, int ld, float eps, void* stream) {
  CB_CHECK(M > 0 && N > 0 && ld >= N, "inv_euclid: M=%d N=%d ld=%d", M, N, ld);
  const long n = (long)M * N;
  inv_euclid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dots, sqnorm_a, sqnorm_b, M, N, ld, eps);
  CB_CUDA(cudaGetLastError());
  return 0;
}
