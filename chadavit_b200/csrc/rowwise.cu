// HBM-bound row-wise kernels of the encoder: im2col+cast, LayerNorm forward/backward (warp per token row,
// 16-byte vector accesses, fp32 statistics), column sums (bias gradients), casts.
#include "common.cuh"
#include "chadavit_b200.h"
#include <stdlib.h>

namespace cb {

// ------------------------------------------------------------------------------------------------ im2col
// x (G,1,H,W) fp32  ->  patches [G*hp*wp, P*P] bf16, patch p = py*wp+px, element r*P+c  (== Conv2d(k=s=P) unfold,
// chada_vit.py:128-133).  One thread moves 8 consecutive pixels of one image row: 32 B read, 16 B write.
__global__ void im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int G, int H, int W, int P) {
  const int hp = H / P, wp = W / P;
  const int chunks_per_row = (wp * P) / 8;
  const long total = (long)G * hp * P * chunks_per_row;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks_per_row);
    long t = i / chunks_per_row;
    const int y = (int)(t % (hp * P));
    const int gimg = (int)(t / (hp * P));
    const int xpix = ch * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + ((long)gimg * H + y) * W + xpix));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + ((long)gimg * H + y) * W + xpix + 4));
    const int py = y / P, r = y - py * P, px = xpix / P, c = xpix - px * P;
    __nv_bfloat16* dst = out + ((long)gimg * hp * wp + py * wp + px) * (P * P) + r * P + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// The residual stream / LayerNorm inputs are kept in fp32 (only GEMM operands are rounded to bf16), so the LN kernels
// read fp32 rows and can emit bf16 (next GEMM operand) and/or fp32 (next residual) outputs in one pass.
// LN_MAX_CHUNKS (template): 8-element chunks per lane; D <= 32 lanes * chunks * 8  (1 -> D<=256, 2 -> 512, 4 -> 1024)
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// y[r] = LN(x[i]) * gamma + beta;  i = in_idx ? in_idx[r] : r.   mean/rstd saved per output row r.
template <int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const int* __restrict__ in_idx,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const int nchunk = D / 8;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps_per_grid) {
    const long src = in_idx ? in_idx[r] : r;
    float v[LN_MAX_CHUNKS][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int c = lane + k * 32;
      if (c < nchunk) {
        ld8(x + src * D + c * 8, v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
      }
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k)
      if (lane + k * 32 < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; q += d * d; }
      }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) { if (mean_out) mean_out[r] = mean; if (rstd_out) rstd_out[r] = rstd; }
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int c = lane + k * 32;
      if (c < nchunk) {
        float gg[8], bb[8], o[8];
        ld8(gamma + c * 8, gg); ld8(beta + c * 8, bb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * gg[j] + bb[j];
        if (y) st8_bf16(y + (long)r * D + c * 8, o);
        if (y32) st8(y32 + (long)r * D + c * 8, o);
      }
    }
  }
}

// dx[i] = LNbwd(dy[r], x[i]) (+ dres[i]);  i = idx ? idx[r] : r.  dgamma/dbeta/dcolsum are ACCUMULATED (fp32 atomics,
// one per block per column).  dy, x, dres fp32; dx written as fp32 (dx32) and/or bf16 (dx16).
template <int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const int* __restrict__ idx, const float* __restrict__ gamma,
                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                            const float* __restrict__ dres, float* __restrict__ dx32,
                                                            __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows, int D) {
  extern __shared__ float red[];  // [3][D]
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const int nchunk = D / 8;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ag[LN_MAX_CHUNKS][8], ab[LN_MAX_CHUNKS][8], ac[LN_MAX_CHUNKS][8], gam[LN_MAX_CHUNKS][8];
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[k][j] = 0.f; ab[k][j] = 0.f; ac[k][j] = 0.f; gam[k][j] = 0.f; }
    if (lane + k * 32 < nchunk) ld8(gamma + (lane + k * 32) * 8, gam[k]);
  }
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps_per_grid) {
    const long xi = idx ? idx[r] : r;
    const float mean = mean_in[r], rstd = rstd_in[r];
    float xh[LN_MAX_CHUNKS][8], g[LN_MAX_CHUNKS][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int c = lane + k * 32;
      if (c < nchunk) {
        float d[8];
        ld8(dy + (long)r * D + c * 8, d);
        ld8(x + xi * D + c * 8, xh[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[k][j] = (xh[k][j] - mean) * rstd;
          ag[k][j] += d[j] * xh[k][j]; ab[k][j] += d[j];
          g[k][j] = d[j] * gam[k][j];
          s1 += g[k][j]; s2 += g[k][j] * xh[k][j];
        }
      }
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int c = lane + k * 32;
      if (c < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j] = rstd * (g[k][j] - s1 - xh[k][j] * s2); ac[k][j] += o[j]; }
        if (dres) {
          float e[8];
          ld8(dres + xi * D + c * 8, e);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += e[j];
        }
        if (dx32) st8(dx32 + xi * D + c * 8, o);
        if (dx16) st8_bf16(dx16 + xi * D + c * 8, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
    const int c = lane + k * 32;
    if (c < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[c * 8 + j], ag[k][j]);
        atomicAdd(&red[D + c * 8 + j], ab[k][j]);
        atomicAdd(&red[2 * D + c * 8 + j], ac[k][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[D + i]);
    if (dcolsum) atomicAdd(dcolsum + i, red[2 * D + i]);
  }
}

// ---- D in {64,128,192,256}: 8 lanes per row, 4 rows per warp (all 32 lanes busy at D=192, 4x the bytes in flight per warp)
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <int NJ>
__global__ void __launch_bounds__(256) layernorm_fwd_g8_kernel(const float* __restrict__ x, const int* __restrict__ in_idx,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                                                               float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, float eps) {
  constexpr int D = NJ * 64;
  const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwg = gridDim.x * (blockDim.x >> 5);
  float gg[NJ][8], bb[NJ][8];
#pragma unroll
  for (int j = 0; j < NJ; ++j) { ld8(gamma + j * 64 + gl * 8, gg[j]); ld8(beta + j * 64 + gl * 8, bb[j]); }
  for (int rb = wg * 4; rb < rows; rb += nwg * 4) {
    const int r = rb + grp;
    const bool ok = r < rows;
    const long src = ok ? (in_idx ? in_idx[r] : r) : 0;
    float v[NJ][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (ok) ld8(x + src * D + j * 64 + gl * 8, v[j]);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[j][k];
    }
    const float mean = group8_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[j][k] - mean; q += d * d; }
    const float rstd = rsqrtf(group8_sum(q) * (1.f / D) + eps);
    if (ok) {
      if (gl == 0) { if (mean_out) mean_out[r] = mean; if (rstd_out) rstd_out[r] = rstd; }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[j][k] - mean) * rstd * gg[j][k] + bb[j][k];
        if (y) st8_bf16(y + (long)r * D + j * 64 + gl * 8, o);
        if (y32) st8(y32 + (long)r * D + j * 64 + gl * 8, o);
      }
    }
  }
}

// Two LayerNorms back to back on a row kept in registers: y1 = LN_a(x) (fp32: the residual stream entering the next block,
// chada_vit.py:100 norm2) and y2 = LN_b(y1) (bf16: the next block's norm1 output, the QKV GEMM operand, chada_vit.py:96).
// Separate launches would write y1 and read it straight back (768 of 2688 bytes per token).
template <int NJ>
__global__ void __launch_bounds__(256) layernorm2_fwd_g8_kernel(const float* __restrict__ x, const float* __restrict__ gamma_a,
                                                                const float* __restrict__ beta_a, float eps_a, const float* __restrict__ gamma_b,
                                                                const float* __restrict__ beta_b, float eps_b, float* __restrict__ y1,
                                                                __nv_bfloat16* __restrict__ y2, float* __restrict__ mean_a,
                                                                float* __restrict__ rstd_a, float* __restrict__ mean_b, float* __restrict__ rstd_b,
                                                                int rows) {
  constexpr int D = NJ * 64;
  const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwg = gridDim.x * (blockDim.x >> 5);
  for (int rb = wg * 4; rb < rows; rb += nwg * 4) {
    const int r = rb + grp;
    const bool ok = r < rows;
    float v[NJ][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (ok) ld8(x + (long)r * D + j * 64 + gl * 8, v[j]);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[j][k];
    }
    const float m_a = group8_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[j][k] - m_a; q += d * d; }
    const float r_a = rsqrtf(group8_sum(q) * (1.f / D) + eps_a);
    s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float gg[8], bb[8];
      ld8(gamma_a + j * 64 + gl * 8, gg); ld8(beta_a + j * 64 + gl * 8, bb);
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[j][k] = (v[j][k] - m_a) * r_a * gg[k] + bb[k]; s += v[j][k]; }
      if (ok && y1) st8(y1 + (long)r * D + j * 64 + gl * 8, v[j]);
    }
    const float m_b = group8_sum(s) * (1.f / D);
    q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[j][k] - m_b; q += d * d; }
    const float r_b = rsqrtf(group8_sum(q) * (1.f / D) + eps_b);
    if (ok) {
      if (gl == 0) {
        if (mean_a) mean_a[r] = m_a;
        if (rstd_a) rstd_a[r] = r_a;
        if (mean_b) mean_b[r] = m_b;
        if (rstd_b) rstd_b[r] = r_b;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float gg[8], bb[8], o[8];
        ld8(gamma_b + j * 64 + gl * 8, gg); ld8(beta_b + j * 64 + gl * 8, bb);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[j][k] - m_b) * r_b * gg[k] + bb[k];
        st8_bf16(y2 + (long)r * D + j * 64 + gl * 8, o);
      }
    }
  }
}


// ---- 16 lanes per row (a lane owns 4 consecutive columns of every 64-column block; a warp = 2 rows per trip): half the per-lane
// state of the 8-lane kernels above, i.e. more resident blocks per SM for kernels that only stream memory (see layernorm_bwd_g16).
__device__ __forceinline__ float group16_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}
__device__ __forceinline__ float sum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float sq4(const float4& a, float m) {
  const float x = a.x - m, y = a.y - m, z = a.z - m, w = a.w - m;
  return (x * x + y * y) + (z * z + w * w);
}
__device__ __forceinline__ float4 affine4(const float4& a, float m, float r, const float4& g, const float4& b) {
  return make_float4((a.x - m) * r * g.x + b.x, (a.y - m) * r * g.y + b.y, (a.z - m) * r * g.z + b.z, (a.w - m) * r * g.w + b.w);
}

template <int NJ>
__global__ void __launch_bounds__(256, 4) layernorm_fwd_g16_kernel(const float* __restrict__ x, const int* __restrict__ in_idx,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                                                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, float eps) {
  constexpr int D = NJ * 64;
  const int lane = threadIdx.x & 31, grp = lane >> 4, gl = lane & 15;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwg = gridDim.x * (blockDim.x >> 5);
  float4 gg[NJ], bb[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    gg[j] = *reinterpret_cast<const float4*>(gamma + j * 64 + gl * 4);
    bb[j] = *reinterpret_cast<const float4*>(beta + j * 64 + gl * 4);
  }
  for (int rb = wg * 2; rb < rows; rb += nwg * 2) {
    const int r = rb + grp;
    const bool ok = r < rows;
    const long src = ok ? (in_idx ? in_idx[r] : r) : 0;
    float4 v[NJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) v[j] = ok ? *reinterpret_cast<const float4*>(x + src * D + j * 64 + gl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += sum4(v[j]);
    const float mean = group16_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) q += sq4(v[j], mean);
    const float rstd = rsqrtf(group16_sum(q) * (1.f / D) + eps);
    if (ok) {
      if (gl == 0) { if (mean_out) mean_out[r] = mean; if (rstd_out) rstd_out[r] = rstd; }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float4 o = affine4(v[j], mean, rstd, gg[j], bb[j]);
        if (y) *reinterpret_cast<uint2*>(y + (long)r * D + j * 64 + gl * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (y32) *reinterpret_cast<float4*>(y32 + (long)r * D + j * 64 + gl * 4) = o;
      }
    }
  }
}

template <int NJ>
__global__ void __launch_bounds__(256, 4) layernorm2_fwd_g16_kernel(const float* __restrict__ x, const float* __restrict__ gamma_a,
                                                                    const float* __restrict__ beta_a, float eps_a, const float* __restrict__ gamma_b,
                                                                    const float* __restrict__ beta_b, float eps_b, float* __restrict__ y1,
                                                                    __nv_bfloat16* __restrict__ y2, float* __restrict__ mean_a,
                                                                    float* __restrict__ rstd_a, float* __restrict__ mean_b, float* __restrict__ rstd_b,
                                                                    int rows) {
  constexpr int D = NJ * 64;
  const int lane = threadIdx.x & 31, grp = lane >> 4, gl = lane & 15;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwg = gridDim.x * (blockDim.x >> 5);
  for (int rb = wg * 2; rb < rows; rb += nwg * 2) {
    const int r = rb + grp;
    const bool ok = r < rows;
    float4 v[NJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) v[j] = ok ? *reinterpret_cast<const float4*>(x + (long)r * D + j * 64 + gl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += sum4(v[j]);
    const float m_a = group16_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) q += sq4(v[j], m_a);
    const float r_a = rsqrtf(group16_sum(q) * (1.f / D) + eps_a);
    s = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      v[j] = affine4(v[j], m_a, r_a, *reinterpret_cast<const float4*>(gamma_a + j * 64 + gl * 4), *reinterpret_cast<const float4*>(beta_a + j * 64 + gl * 4));
      s += sum4(v[j]);
      if (ok && y1) *reinterpret_cast<float4*>(y1 + (long)r * D + j * 64 + gl * 4) = v[j];
    }
    const float m_b = group16_sum(s) * (1.f / D);
    q = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) q += sq4(v[j], m_b);
    const float r_b = rsqrtf(group16_sum(q) * (1.f / D) + eps_b);
    if (ok) {
      if (gl == 0) {
        if (mean_a) mean_a[r] = m_a;
        if (rstd_a) rstd_a[r] = r_a;
        if (mean_b) mean_b[r] = m_b;
        if (rstd_b) rstd_b[r] = r_b;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float4 o = affine4(v[j], m_b, r_b, *reinterpret_cast<const float4*>(gamma_b + j * 64 + gl * 4), *reinterpret_cast<const float4*>(beta_b + j * 64 + gl * 4));
        *reinterpret_cast<uint2*>(y2 + (long)r * D + j * 64 + gl * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}

template <int NJ>
__global__ void __launch_bounds__(256) layernorm_bwd_g8_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                               const int* __restrict__ idx, const float* __restrict__ gamma,
                                                               const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                               const float* __restrict__ dres, float* __restrict__ dx32,
                                                               __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows) {
  constexpr int D = NJ * 64;
  __shared__ float red[3 * D];
  const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwg = gridDim.x * (blockDim.x >> 5);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ag[NJ][8], ab[NJ][8], ac[NJ][8], gam[NJ][8];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    ld8(gamma + j * 64 + gl * 8, gam[j]);
#pragma unroll
    for (int k = 0; k < 8; ++k) { ag[j][k] = 0.f; ab[j][k] = 0.f; ac[j][k] = 0.f; }
  }
  for (int rb = wg * 4; rb < rows; rb += nwg * 4) {
    const int r = rb + grp;
    const bool ok = r < rows;
    const long xi = ok ? (idx ? idx[r] : r) : 0;
    const float mean = ok ? mean_in[r] : 0.f, rstd = ok ? rstd_in[r] : 0.f;
    float xh[NJ][8], g[NJ][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float d[8];
      if (ok) { ld8(dy + (long)r * D + j * 64 + gl * 8, d); ld8(x + xi * D + j * 64 + gl * 8, xh[j]); }
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) { d[k] = 0.f; xh[j][k] = 0.f; }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[j][k] = (xh[j][k] - mean) * rstd;
        ag[j][k] += d[k] * xh[j][k]; ab[j][k] += d[k];
        g[j][k] = d[k] * gam[j][k];
        s1 += g[j][k]; s2 += g[j][k] * xh[j][k];
      }
    }
    s1 = group8_sum(s1) * (1.f / D); s2 = group8_sum(s2) * (1.f / D);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { o[k] = rstd * (g[j][k] - s1 - xh[j][k] * s2); ac[j][k] += o[k]; }
      if (ok) {
        if (dres) {
          float e[8];
          ld8(dres + xi * D + j * 64 + gl * 8, e);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += e[k];
        }
        if (dx32) st8(dx32 + xi * D + j * 64 + gl * 8, o);
        if (dx16) st8_bf16(dx16 + xi * D + j * 64 + gl * 8, o);
      }
    }
  }
  // fold the 4 row groups of the warp, then the warps of the block, then one atomic per column per block
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = ag[j][k], b = ab[j][k], c = ac[j][k];
      a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
      b += __shfl_xor_sync(0xffffffffu, b, 8); b += __shfl_xor_sync(0xffffffffu, b, 16);
      c += __shfl_xor_sync(0xffffffffu, c, 8); c += __shfl_xor_sync(0xffffffffu, c, 16);
      if (grp == 0) {
        const int col = j * 64 + gl * 8 + k;
        atomicAdd(&red[col], a); atomicAdd(&red[D + col], b); atomicAdd(&red[2 * D + col], c);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[D + i]);
    if (dcolsum) atomicAdd(dcolsum + i, red[2 * D + i]);
  }
}

// Same contract, 16 lanes per row (a lane owns 4 consecutive columns of every 64-column block; a warp = 2 rows per trip).
// The 8-lane kernel above keeps 3 x 24 column accumulators, gamma and two copies of its 24 row elements per lane: 202
// registers, i.e. ONE 8-warp block per SM for a kernel that only streams memory, and its `dres` loads are issued after the row
// reduction (two dependent memory phases per trip): 3.7 TB/s.  Half the columns per lane halve all of that (two blocks per SM),
// and dres is requested together with dy and x.
template <int NJ>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_g16_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                   const int* __restrict__ idx, const float* __restrict__ gamma,
                                                                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                   const float* __restrict__ dres, float* __restrict__ dx32,
                                                                   __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows) {
  constexpr int D = NJ * 64;
  __shared__ float red[3 * D];
  const int lane = threadIdx.x & 31, grp = lane >> 4, gl = lane & 15;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwg = gridDim.x * (blockDim.x >> 5);
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 ag[NJ], ab[NJ], ac[NJ], gam[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    gam[j] = *reinterpret_cast<const float4*>(gamma + j * 64 + gl * 4);
    ag[j] = ab[j] = ac[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int rb = wg * 2; rb < rows; rb += nwg * 2) {
    const int r = rb + grp;
    const bool ok = r < rows;
    const long xi = ok ? (idx ? idx[r] : r) : 0;
    const float mean = ok ? mean_in[r] : 0.f, rstd = ok ? rstd_in[r] : 0.f;
    float4 d[NJ], xh[NJ], e[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {   // every load of the trip is requested here, before the first use
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      d[j] = ok ? *reinterpret_cast<const float4*>(dy + (long)r * D + j * 64 + gl * 4) : z;
      xh[j] = ok ? *reinterpret_cast<const float4*>(x + xi * D + j * 64 + gl * 4) : z;
      e[j] = (ok && dres) ? *reinterpret_cast<const float4*>(dres + xi * D + j * 64 + gl * 4) : z;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (ok) { xh[j].x = (xh[j].x - mean) * rstd; xh[j].y = (xh[j].y - mean) * rstd; xh[j].z = (xh[j].z - mean) * rstd; xh[j].w = (xh[j].w - mean) * rstd; }
      ag[j].x += d[j].x * xh[j].x; ag[j].y += d[j].y * xh[j].y; ag[j].z += d[j].z * xh[j].z; ag[j].w += d[j].w * xh[j].w;
      ab[j].x += d[j].x; ab[j].y += d[j].y; ab[j].z += d[j].z; ab[j].w += d[j].w;
      d[j].x *= gam[j].x; d[j].y *= gam[j].y; d[j].z *= gam[j].z; d[j].w *= gam[j].w;          // d <- g = dy * gamma
      s1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
      s2 += (d[j].x * xh[j].x + d[j].y * xh[j].y) + (d[j].z * xh[j].z + d[j].w * xh[j].w);
    }
    s1 = group16_sum(s1) * (1.f / D); s2 = group16_sum(s2) * (1.f / D);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float4 o;
      o.x = rstd * (d[j].x - s1 - xh[j].x * s2); o.y = rstd * (d[j].y - s1 - xh[j].y * s2);
      o.z = rstd * (d[j].z - s1 - xh[j].z * s2); o.w = rstd * (d[j].w - s1 - xh[j].w * s2);
      ac[j].x += o.x; ac[j].y += o.y; ac[j].z += o.z; ac[j].w += o.w;
      if (ok) {
        o.x += e[j].x; o.y += e[j].y; o.z += e[j].z; o.w += e[j].w;
        if (dx32) *reinterpret_cast<float4*>(dx32 + xi * D + j * 64 + gl * 4) = o;
        if (dx16) *reinterpret_cast<uint2*>(dx16 + xi * D + j * 64 + gl * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
  // fold the 2 row groups of the warp, then the warps of the block, then one atomic per column per block
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float a4[4] = {ag[j].x, ag[j].y, ag[j].z, ag[j].w}, b4[4] = {ab[j].x, ab[j].y, ab[j].z, ab[j].w}, c4[4] = {ac[j].x, ac[j].y, ac[j].z, ac[j].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = a4[k] + __shfl_xor_sync(0xffffffffu, a4[k], 16), b = b4[k] + __shfl_xor_sync(0xffffffffu, b4[k], 16),
                  c = c4[k] + __shfl_xor_sync(0xffffffffu, c4[k], 16);
      if (grp == 0) {
        const int col = j * 64 + gl * 4 + k;
        atomicAdd(&red[col], a); atomicAdd(&red[D + col], b); atomicAdd(&red[2 * D + col], c);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[D + i]);
    if (dcolsum) atomicAdd(dcolsum + i, red[2 * D + i]);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[n] += sum_t x[t, n]   (bias gradients of in_proj / linear1).  grid (row chunks, column groups of 2048 columns).
// The 256 threads of a block are laid out as ncx column chunks (8 columns = one 16-byte load each) x nry rows, so narrow
// matrices (N = 576 -> 72 chunks x 3 rows) still use the whole block; partial sums meet in smem, then one fp32 atomic per column.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int T, int N, int ld,
                                                     int rows_per_block) {
  __shared__ float red[2048];
  const int ncols = min(N - blockIdx.y * 2048, 2048);
  const int ncx = ncols / 8, nry = 256 / ncx;
  for (int i = threadIdx.x; i < ncols; i += 256) red[i] = 0.f;
  __syncthreads();
  const int cx = threadIdx.x % ncx, ry = threadIdx.x / ncx;
  if (ry < nry) {
    const int c8 = blockIdx.y * 2048 + cx * 8;
    const int t0 = blockIdx.x * rows_per_block, t1 = min(T, t0 + rows_per_block);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
    for (int t = t0 + ry; t < t1; t += nry) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (long)t * ld + c8));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uu[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[cx * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncols; i += 256) atomicAdd(out + blockIdx.y * 2048 + i, red[i]);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long n) {
  const long i8 = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 8;
  if (i8 + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(in + i8), b = *reinterpret_cast<const float4*>(in + i8 + 4);
    *reinterpret_cast<uint4*>(out + i8) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  } else {
    for (long i = i8; i < n; ++i) out[i] = __float2bfloat16(in[i]);
  }
}

// rows of a bf16 matrix gathered to fp32 (CLS features -> head input) / scattered back
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out, int rows, int D) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < (long)rows * D) { const int r = (int)(i / D), c = (int)(i % D); out[i] = __bfloat162float(x[(long)idx[r] * D + c]); }
}

// C[M,N] (+)= A[M,K] B[K,N] (trans_a: A is stored [K,M]) in fp32 — only for the tiny positional-embedding resize
// (N' x 196 interpolation matrix times pos_embed), never on token-sized data.
__global__ void small_matmul_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N,
                                        int K, int trans_a, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += (trans_a ? A[(long)k * M + m] : A[(long)m * K + k]) * B[(long)k * N + n];
  if (accumulate) C[(long)m * N + n] += acc; else C[(long)m * N + n] = acc;
}

}  // namespace cb

using namespace cb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFM(p) reinterpret_cast<__nv_bfloat16*>(p)

extern "C" int cb_im2col_bf16(const float* x, void* patches, int G, int H, int W, int patch, void* stream) {
  CB_CHECK(G > 0 && patch > 0 && patch % 8 == 0 && H >= patch && W >= patch, "im2col: bad shape G=%d H=%d W=%d patch=%d", G, H, W, patch);
  CB_CHECK(W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "im2col: W must be a multiple of 4 and x 16B aligned");
  const long total = (long)G * (H / patch) * patch * ((W / patch) * patch / 8);
  const int threads = 256;
  long blocks = (total + threads - 1) / threads;
  if (blocks > (long)num_sms() * 32) blocks = (long)num_sms() * 32;
  im2col_kernel<<<(int)blocks, threads, 0, STREAM>>>(x, BFM(patches), G, H, W, patch);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_layernorm_fwd(const float* x, const int* in_idx, const float* gamma, const float* beta, void* y, float* y_f32,
                                float* mean, float* rstd, int rows, int D, float eps, void* stream) {
  CB_CHECK(rows > 0 && D % 8 == 0 && D <= 1024, "layernorm_fwd: rows=%d D=%d (D must be a multiple of 8, <= 1024)", rows, D);
  int blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
#define LNF(N) layernorm_fwd_kernel<N><<<blocks, 256, 0, STREAM>>>(x, in_idx, gamma, beta, BFM(y), y_f32, mean, rstd, rows, D, eps)
#define LNFG(N) layernorm_fwd_g8_kernel<N><<<gblocks, 256, 0, STREAM>>>(x, in_idx, gamma, beta, BFM(y), y_f32, mean, rstd, rows, eps)
  int gblocks = (rows + 31) / 32;
  if (gblocks > num_sms() * 8) gblocks = num_sms() * 8;
  static const int variant = [] { const char* e = getenv("CB_LN_FWD_LANES"); return e ? atoi(e) : 0; }();   // 8 forces the 8-lane kernels (A/B)
  int hblocks = (rows + 15) / 16;
  if (hblocks > num_sms() * 8) hblocks = num_sms() * 8;
#define LNFH(N) layernorm_fwd_g16_kernel<N><<<hblocks, 256, 0, STREAM>>>(x, in_idx, gamma, beta, BFM(y), y_f32, mean, rstd, rows, eps)
  if (variant != 8 && (D == 192 || D == 256)) { if (D == 192) LNFH(3); else LNFH(4); }
  else
#undef LNFH
  if (D == 64) LNFG(1); else if (D == 128) LNFG(2); else if (D == 192) LNFG(3); else if (D == 256) LNFG(4);
  else if (D <= 256) LNF(1); else if (D <= 512) LNF(2); else LNF(4);
#undef LNFG
#undef LNF
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_layernorm2_fwd(const float* x, const float* gamma_a, const float* beta_a, float eps_a, const float* gamma_b,
                                 const float* beta_b, float eps_b, float* y1_f32, void* y2_bf16, float* mean_a, float* rstd_a, float* mean_b,
                                 float* rstd_b, int rows, int D, void* stream) {
  CB_CHECK(rows > 0 && (D == 64 || D == 128 || D == 192 || D == 256) && y2_bf16, "layernorm2_fwd: rows=%d D=%d (D must be 64, 128, 192 or 256)", rows, D);
  int gblocks = (rows + 31) / 32;
  if (gblocks > num_sms() * 8) gblocks = num_sms() * 8;
#define LN2G(N) layernorm2_fwd_g8_kernel<N><<<gblocks, 256, 0, STREAM>>>(x, gamma_a, beta_a, eps_a, gamma_b, beta_b, eps_b, y1_f32, BFM(y2_bf16), mean_a, rstd_a, mean_b, rstd_b, rows)
  static const int variant = [] { const char* e = getenv("CB_LN_FWD_LANES"); return e ? atoi(e) : 0; }();
  int hblocks = (rows + 15) / 16;
  if (hblocks > num_sms() * 8) hblocks = num_sms() * 8;
#define LN2H(N) layernorm2_fwd_g16_kernel<N><<<hblocks, 256, 0, STREAM>>>(x, gamma_a, beta_a, eps_a, gamma_b, beta_b, eps_b, y1_f32, BFM(y2_bf16), mean_a, rstd_a, mean_b, rstd_b, rows)
  if (variant != 8 && (D == 192 || D == 256)) { if (D == 192) LN2H(3); else LN2H(4); }
  else
#undef LN2H
  if (D == 64) LN2G(1); else if (D == 128) LN2G(2); else if (D == 192) LN2G(3); else LN2G(4);
#undef LN2G
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_layernorm_bwd(const float* dy, const float* x, const int* idx, const float* gamma, const float* mean,
                                const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                                float* dcolsum, int rows, int D, void* stream) {
  CB_CHECK(rows > 0 && D % 8 == 0 && D <= 1024, "layernorm_bwd: rows=%d D=%d", rows, D);
  CB_CHECK(dx_f32 || dx_bf16, "layernorm_bwd: at least one of dx_f32 / dx_bf16");
  int blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
#define LNB(N) layernorm_bwd_kernel<N><<<blocks, 256, 3 * D * sizeof(float), STREAM>>>(dy, x, idx, gamma, mean, rstd, dres, dx_f32, BFM(dx_bf16), dgamma, dbeta, dcolsum, rows, D)
#define LNBG(N) layernorm_bwd_g8_kernel<N><<<gblocks, 256, 0, STREAM>>>(dy, x, idx, gamma, mean, rstd, dres, dx_f32, BFM(dx_bf16), dgamma, dbeta, dcolsum, rows)
  int gblocks = (rows + 31) / 32;
  if (gblocks > num_sms() * 4) gblocks = num_sms() * 4;
  // D = 192 / 256: 16 lanes per row (two resident blocks per SM); `variant` (A/B, tests): 8 forces the 8-lane kernel
  static const int variant = [] { const char* e = getenv("CB_LN_BWD_LANES"); return e ? atoi(e) : 0; }();
  int g16blocks = (rows + 15) / 16;
  if (g16blocks > num_sms() * 2) g16blocks = num_sms() * 2;
#define LNBH(N) layernorm_bwd_g16_kernel<N><<<g16blocks, 256, 0, STREAM>>>(dy, x, idx, gamma, mean, rstd, dres, dx_f32, BFM(dx_bf16), dgamma, dbeta, dcolsum, rows)
  if (variant != 8 && (D == 192 || D == 256)) { if (D == 192) LNBH(3); else LNBH(4); }
  else
#undef LNBH
  if (D == 64) LNBG(1); else if (D == 128) LNBG(2); else if (D == 192) LNBG(3); else if (D == 256) LNBG(4);
  else if (D <= 256) LNB(1); else if (D <= 512) LNB(2); else LNB(4);
#undef LNBG
#undef LNB
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_colsum_bf16(const void* x, int ld, float* out, int T, int N, void* stream) {
  CB_CHECK(T > 0 && N % 8 == 0 && ld % 8 == 0, "colsum: T=%d N=%d ld=%d", T, N, ld);
  const int col_groups = (N + 2047) / 2048;
  int row_blocks = (num_sms() * 8 + col_groups - 1) / col_groups;
  int rpb = (T + row_blocks - 1) / row_blocks;
  if (rpb < 32) rpb = 32;
  row_blocks = (T + rpb - 1) / rpb;
  colsum_kernel<<<dim3(row_blocks, col_groups), 256, 0, STREAM>>>(BF(x), out, T, N, ld, rpb);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_cast_f32_bf16(const float* in, void* out, long n, void* stream) {
  CB_CHECK(n > 0, "cast: n=%ld", n);
  const long blocks = (n + 2047) / 2048;
  cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>(in, BFM(out), n);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_gather_rows_f32(const void* x, const int* idx, float* out, int rows, int D, void* stream) {
  CB_CHECK(rows > 0 && D > 0, "gather_rows: rows=%d D=%d", rows, D);
  const long n = (long)rows * D;
  gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(BF(x), idx, out, rows, D);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_small_matmul_f32(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int accumulate,
                                   void* stream) {
  CB_CHECK(M > 0 && N > 0 && K > 0 && M <= 65535, "small_matmul: M=%d N=%d K=%d", M, N, K);
  small_matmul_f32_kernel<<<dim3((N + 127) / 128, M), 128, 0, STREAM>>>(A, B, C, M, N, K, trans_a, accumulate);
  CB_CUDA(cudaGetLastError());
  return 0;
}
