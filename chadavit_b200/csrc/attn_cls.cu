// Self-attention of the LAST encoder block when the backbone returns only the CLS embedding (src/backbones/vit/chada_vit.py:289,
// `return x[:, 0]`; the SDPA is nn.MultiheadAttention at :105-111).  Behind the last block's attention everything is row-wise
// (out-projection, norm1, feed-forward, norm2, final norm), so only the CLS row of every sequence reaches the output, and only
// the CLS QUERY of the last attention is ever used: its keys and values are still all real tokens of the sequence.  Likewise
// in the backward pass d(attention output) is zero on every row but the CLS rows, so dS has one non-zero row per sequence.
//
//   forward :  o_b,h = softmax_j(q_cls . k_j / sqrt(d)) V                          -> out_cls [nseq, D] bf16, lse [nseq, H] (log2 domain)
//   backward:  p_j = exp2(s_j - lse),  dP_j = dO . v_j,  dS_j = p_j (dP_j - dO . o)
//              dV_j = p_j dO,  dK_j = dS_j q / sqrt(d),  dQ_cls = sum_j dS_j k_j / sqrt(d),  dQ_j = 0 (j > 0)
//              -> dqkv [T, 3D] bf16, every element written (the weight-gradient and input-gradient products read all of it)
//
// One query row against S keys is a matrix-vector product: 2 x S x d bf16 read once (forward), 3 x S x d written (backward):
// HBM-bound, no tensor-core shape.  One CTA per (sequence, head); a row of K (d bf16 = d / 8 sixteen-byte chunks) is spread over
// NCHP = pow2(d / 8) neighbouring lanes, so a warp reads 32 / NCHP whole rows per instruction (contiguous 2 d bytes each) and a
// lane keeps ITS chunk of q / dO / the accumulators in registers for the whole sequence; U rows per lane group are in flight
// at once.  The forward is a single pass (running maximum per lane group, groups merged through shared memory at the end).
#include "common.cuh"
#include "chadavit_b200.h"

namespace cb {

namespace acls {
constexpr int NT = 512;      // threads per CTA: 16 warps x U rows x 2 operands x 16 B per lane = 64 KB requested per trip
constexpr int U = 4;         // rows in flight per lane group
__host__ __device__ constexpr int pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
  return s;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {     // sum over the W neighbouring lanes of a lane group (W a power of two)
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace acls

template <int HD>
__global__ void __launch_bounds__(acls::NT, 2) attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ cu, int H,
                                                                float scale_log2, __nv_bfloat16* __restrict__ out_cls,
                                                                float* __restrict__ lse_cls) {
  using namespace acls;
  constexpr int NCH = HD / 8, NCHP = pow2ceil(NCH), RPW = 32 / NCHP, G = (NT / 32) * RPW;
  __shared__ float s_m[G], s_l[G];
  __shared__ float s_acc[G][HD];
  const int b = blockIdx.x, h = blockIdx.y;
  const int t0 = cu[b], S = cu[b + 1] - t0;
  const int D = H * HD;
  const long ld = 3L * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane / NCHP, part = lane % NCHP, g = warp * RPW + r;
  const bool active = part < NCH;
  const __nv_bfloat16* base = qkv + (long)t0 * ld + h * HD + (active ? part : 0) * 8;   // q chunk of the CLS row; k at +D, v at +2D
  float q[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base)), q);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = active ? q[i] * scale_log2 : 0.f;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j0 = 0; j0 < S; j0 += G * U) {           // warp-uniform trip count: the lane-group sums below need every lane
    uint4 kk[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * G + g;
      kk[u] = make_uint4(0u, 0u, 0u, 0u); vv[u] = kk[u];
      if (j < S && active) {
        const __nv_bfloat16* p = base + (long)j * ld;
        kk[u] = __ldg(reinterpret_cast<const uint4*>(p + D));
        vv[u] = __ldg(reinterpret_cast<const uint4*>(p + 2 * D));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kf[8], vf[8];
      unpack8(kk[u], kf);
      unpack8(vv[u], vf);
      const float s = group_sum<NCHP>(dot8(q, kf));
      if (j0 + u * G + g < S) {
        const float mn = fmaxf(m, s);
        const float f = ex2(m - mn), p = ex2(s - mn);     // first row: m = -inf -> f = 0
        l = fmaf(l, f, p);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], f, p * vf[i]);
        m = mn;
      }
    }
  }
  if (part == 0) { s_m[g] = m; s_l[g] = l; }
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[g][part * 8 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < HD) {                            // merge the lane groups: groups that saw no row carry m = -inf, l = 0
    const int c = threadIdx.x;
    float M = -INFINITY;
    for (int k = 0; k < G; ++k) M = fmaxf(M, s_m[k]);
    float L = 0.f, o = 0.f;
    for (int k = 0; k < G; ++k) {
      const float w = ex2(s_m[k] - M);
      L = fmaf(s_l[k], w, L);
      o = fmaf(s_acc[k][c], w, o);
    }
    out_cls[(long)b * D + h * HD + c] = __float2bfloat16(o / L);
    if (c == 0) lse_cls[(long)b * H + h] = M + log2f(L);
  }
}

template <int HD>
__global__ void __launch_bounds__(acls::NT) attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ dout_cls, const __nv_bfloat16* __restrict__ qkv,
                                                                const __nv_bfloat16* __restrict__ out_cls, const float* __restrict__ lse_cls,
                                                                const int* __restrict__ cu, int H, float scale, float scale_log2,
                                                                __nv_bfloat16* __restrict__ dqkv) {
  using namespace acls;
  constexpr int NCH = HD / 8, NCHP = pow2ceil(NCH), RPW = 32 / NCHP, G = (NT / 32) * RPW;
  __shared__ float s_acc[G][HD];
  const int b = blockIdx.x, h = blockIdx.y;
  const int t0 = cu[b], S = cu[b + 1] - t0;
  const int D = H * HD;
  const long ld = 3L * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane / NCHP, part = lane % NCHP, g = warp * RPW + r;
  const bool active = part < NCH;
  const int col = h * HD + (active ? part : 0) * 8;
  const __nv_bfloat16* base = qkv + (long)t0 * ld + col;
  __nv_bfloat16* dbase = dqkv + (long)t0 * ld + col;
  float q[8], dO[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base)), q);
  unpack8(__ldg(reinterpret_cast<const uint4*>(dout_cls + (long)b * D + col)), dO);
  unpack8(__ldg(reinterpret_cast<const uint4*>(out_cls + (long)b * D + col)), o);
  if (!active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = 0.f; dO[i] = 0.f; o[i] = 0.f; }
  }
  const float delta = group_sum<NCHP>(dot8(dO, o));
  const float lse = lse_cls[(long)b * H + h];
  float dq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i] = 0.f;
  for (int j0 = 0; j0 < S; j0 += G * U) {
    uint4 kk[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * G + g;
      kk[u] = make_uint4(0u, 0u, 0u, 0u); vv[u] = kk[u];
      if (j < S && active) {
        const __nv_bfloat16* p = base + (long)j * ld;
        kk[u] = __ldg(reinterpret_cast<const uint4*>(p + D));
        vv[u] = __ldg(reinterpret_cast<const uint4*>(p + 2 * D));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * G + g;
      float kf[8], vf[8];
      unpack8(kk[u], kf);
      unpack8(vv[u], vf);
      const float s = group_sum<NCHP>(dot8(q, kf));
      const float dp = group_sum<NCHP>(dot8(dO, vf));
      if (j < S && active) {
        const float p = ex2(fmaf(s, scale_log2, -lse));
        const float ds = p * (dp - delta);
        const float dsk = ds * scale;
        __nv_bfloat16* d = dbase + (long)j * ld;
        *reinterpret_cast<uint4*>(d + 2 * D) = make_uint4(pack_bf16(p * dO[0], p * dO[1]), pack_bf16(p * dO[2], p * dO[3]),
                                                          pack_bf16(p * dO[4], p * dO[5]), pack_bf16(p * dO[6], p * dO[7]));
        *reinterpret_cast<uint4*>(d + D) = make_uint4(pack_bf16(dsk * q[0], dsk * q[1]), pack_bf16(dsk * q[2], dsk * q[3]),
                                                      pack_bf16(dsk * q[4], dsk * q[5]), pack_bf16(dsk * q[6], dsk * q[7]));
        if (j > 0) *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);       // dQ of a non-CLS row
#pragma unroll
        for (int i = 0; i < 8; ++i) dq[i] = fmaf(ds, kf[i], dq[i]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[g][part * 8 + i] = dq[i];
  }
  __syncthreads();
  if (threadIdx.x < HD) {
    float a = 0.f;
    for (int k = 0; k < G; ++k) a += s_acc[k][threadIdx.x];
    dqkv[(long)t0 * ld + h * HD + threadIdx.x] = __float2bfloat16(a * scale);       // dQ of the CLS row
  }
}

template <int HD>
static int launch_cls_fwd(const __nv_bfloat16* qkv, const int* cu, int nseq, int H, float scale, __nv_bfloat16* out, float* lse, cudaStream_t st) {
  attn_cls_fwd_kernel<HD><<<dim3(nseq, H), acls::NT, 0, st>>>(qkv, cu, H, scale * 1.4426950408889634f, out, lse);
  CB_CUDA(cudaGetLastError());
  return 0;
}
template <int HD>
static int launch_cls_bwd(const __nv_bfloat16* dout, const __nv_bfloat16* qkv, const __nv_bfloat16* out, const float* lse, const int* cu, int nseq,
                          int H, float scale, __nv_bfloat16* dqkv, cudaStream_t st) {
  attn_cls_bwd_kernel<HD><<<dim3(nseq, H), acls::NT, 0, st>>>(dout, qkv, out, lse, cu, H, scale, scale * 1.4426950408889634f, dqkv);
  CB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cb

using namespace cb;

#define CB_CLS_DISPATCH(FN, ...)                                                                          \
  switch (head_dim) {                                                                                     \
    case 16: return FN<16>(__VA_ARGS__);                                                                  \
    case 32: return FN<32>(__VA_ARGS__);                                                                  \
    case 64: return FN<64>(__VA_ARGS__);                                                                  \
    case 96: return FN<96>(__VA_ARGS__);                                                                  \
    case 128: return FN<128>(__VA_ARGS__);                                                                \
    default: CB_CHECK(false, "attn_cls: unsupported head_dim %d (16/32/64/96/128)", head_dim);            \
  }                                                                                                       \
  return 1;

extern "C" int cb_attn_cls_fwd(const void* qkv, const int* cu_seqlens, int nseq, int num_heads, int head_dim, float scale, void* out_cls,
                               float* lse_cls, void* stream) {
  CB_CHECK(nseq > 0 && num_heads > 0 && num_heads <= 65535, "attn_cls_fwd: nseq=%d heads=%d", nseq, num_heads);
  CB_CLS_DISPATCH(launch_cls_fwd, reinterpret_cast<const __nv_bfloat16*>(qkv), cu_seqlens, nseq, num_heads, scale,
                  reinterpret_cast<__nv_bfloat16*>(out_cls), lse_cls, reinterpret_cast<cudaStream_t>(stream))
}

extern "C" int cb_attn_cls_bwd(const void* dout_cls, const void* qkv, const void* out_cls, const float* lse_cls, const int* cu_seqlens, int nseq,
                               int num_heads, int head_dim, float scale, void* dqkv, void* stream) {
  CB_CHECK(nseq > 0 && num_heads > 0 && num_heads <= 65535, "attn_cls_bwd: nseq=%d heads=%d", nseq, num_heads);
  CB_CLS_DISPATCH(launch_cls_bwd, reinterpret_cast<const __nv_bfloat16*>(dout_cls), reinterpret_cast<const __nv_bfloat16*>(qkv),
                  reinterpret_cast<const __nv_bfloat16*>(out_cls), lse_cls, cu_seqlens, nseq, num_heads, scale,
                  reinterpret_cast<__nv_bfloat16*>(dqkv), reinterpret_cast<cudaStream_t>(stream))
}
