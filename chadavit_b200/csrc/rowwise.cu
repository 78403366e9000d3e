// HBM-bound row-wise kernels of the encoder: im2col+cast, LayerNorm forward/backward (warp per token row,
// 16-byte vector accesses, fp32 statistics), column sums (bias gradients), casts.
#include "common.cuh"
#include "chadavit_b200.h"

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
// LN_MAX_CHUNKS (template): 16-byte chunks per lane; D <= 32 lanes * chunks * 8  (1 -> D<=256, 2 -> 512, 4 -> 1024)

// y[o] = LN(x[i]) * gamma + beta;  i = in_idx ? in_idx[r] : r;  o = r.   mean/rstd saved per output row r.
template <int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ in_idx,
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
        const uint4 u = *reinterpret_cast<const uint4*>(x + src * D + c * 8);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uu[j]); v[k][2 * j] = f.x; v[k][2 * j + 1] = f.y; s += f.x + f.y; }
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
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * gg[j] + bb[j];
        if (y) *reinterpret_cast<uint4*>(y + (long)r * D + c * 8) =
            make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        if (y32) {
          *reinterpret_cast<float4*>(y32 + (long)r * D + c * 8) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(y32 + (long)r * D + c * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
}

// dx[i] = LNbwd(dy[r], x[i]) (+ dres[i]);  i = idx ? idx[r] : r.  dgamma/dbeta/dcolsum are ACCUMULATED (fp32 atomics,
// one per block per column).  dy is bf16 (dy) or fp32 (dy32).
template <int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ dy32,
                                                            const __nv_bfloat16* __restrict__ x, const int* __restrict__ idx,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
                                                            __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows, int D) {
  extern __shared__ float red[];  // [3][D]
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const int nchunk = D / 8;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ag[LN_MAX_CHUNKS][8], ab[LN_MAX_CHUNKS][8], ac[LN_MAX_CHUNKS][8];
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[k][j] = 0.f; ab[k][j] = 0.f; ac[k][j] = 0.f; }
  float gam[LN_MAX_CHUNKS][8];
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
    const int c = lane + k * 32;
    if (c < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) gam[k][j] = __ldg(gamma + c * 8 + j);
    }
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
        const uint4 ux = *reinterpret_cast<const uint4*>(x + xi * D + c * 8);
        const uint32_t uxx[4] = {ux.x, ux.y, ux.z, ux.w};
        float d[8];
        if (dy32) {
          const float4 d0 = *reinterpret_cast<const float4*>(dy32 + (long)r * D + c * 8), d1 = *reinterpret_cast<const float4*>(dy32 + (long)r * D + c * 8 + 4);
          d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
        } else {
          const uint4 ud = *reinterpret_cast<const uint4*>(dy + (long)r * D + c * 8);
          const uint32_t udd[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 fd = unpack_bf16(udd[j]); d[2 * j] = fd.x; d[2 * j + 1] = fd.y; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fx = unpack_bf16(uxx[j]);
          xh[k][2 * j] = (fx.x - mean) * rstd; xh[k][2 * j + 1] = (fx.y - mean) * rstd;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
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
          const uint4 u = *reinterpret_cast<const uint4*>(dres + xi * D + c * 8);
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uu[j]); o[2 * j] += f.x; o[2 * j + 1] += f.y; }
        }
        *reinterpret_cast<uint4*>(dx + xi * D + c * 8) =
            make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
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

// ------------------------------------------------------------------------------------------------ column sums
// out[n] += sum_t x[t, n]   (bias gradients of in_proj / linear1).  grid (row chunks, column groups of 2048).
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int T, int N, int ld,
                                                     int rows_per_block) {
  const int c8 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c8 >= N) return;
  const int t0 = blockIdx.x * rows_per_block, t1 = min(T, t0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
  for (int t = t0; t < t1; ++t) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (long)t * ld + c8));
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uu[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(out + c8 + j, acc[j]);
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

extern "C" int cb_layernorm_fwd(const void* x, const int* in_idx, const float* gamma, const float* beta, void* y, float* y_f32,
                                float* mean, float* rstd, int rows, int D, float eps, void* stream) {
  CB_CHECK(rows > 0 && D % 8 == 0 && D <= 1024, "layernorm_fwd: rows=%d D=%d (D must be a multiple of 8, <= 1024)", rows, D);
  int blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (D <= 256) layernorm_fwd_kernel<1><<<blocks, 256, 0, STREAM>>>(BF(x), in_idx, gamma, beta, BFM(y), y_f32, mean, rstd, rows, D, eps);
  else if (D <= 512) layernorm_fwd_kernel<2><<<blocks, 256, 0, STREAM>>>(BF(x), in_idx, gamma, beta, BFM(y), y_f32, mean, rstd, rows, D, eps);
  else layernorm_fwd_kernel<4><<<blocks, 256, 0, STREAM>>>(BF(x), in_idx, gamma, beta, BFM(y), y_f32, mean, rstd, rows, D, eps);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_layernorm_bwd(const void* dy, const float* dy_f32, const void* x, const int* idx, const float* gamma,
                                const float* mean, const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                float* dcolsum, int rows, int D, void* stream) {
  CB_CHECK(rows > 0 && D % 8 == 0 && D <= 1024, "layernorm_bwd: rows=%d D=%d", rows, D);
  CB_CHECK((dy != nullptr) != (dy_f32 != nullptr), "layernorm_bwd: exactly one of dy / dy_f32");
  int blocks = (rows + 7) / 8;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
#define LNB(N) layernorm_bwd_kernel<N><<<blocks, 256, 3 * D * sizeof(float), STREAM>>>(BF(dy), dy_f32, BF(x), idx, gamma, mean, rstd, BF(dres), BFM(dx), dgamma, dbeta, dcolsum, rows, D)
  if (D <= 256) LNB(1); else if (D <= 512) LNB(2); else LNB(4);
#undef LNB
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_colsum_bf16(const void* x, int ld, float* out, int T, int N, void* stream) {
  CB_CHECK(T > 0 && N % 8 == 0 && ld % 8 == 0, "colsum: T=%d N=%d ld=%d", T, N, ld);
  const int col_groups = (N + 2047) / 2048;
  int row_blocks = (num_sms() * 4 + col_groups - 1) / col_groups;
  int rpb = (T + row_blocks - 1) / row_blocks;
  if (rpb < 16) rpb = 16;
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
