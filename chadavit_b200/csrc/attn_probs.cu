// Attention probabilities of one encoder block, for ChAdaViT.get_last_selfattention (src/backbones/vit/chada_vit.py:313-320;
// consumer: main_attn.py:200-207).  out[b, h, i, j] = softmax_j(q_i . k_j / sqrt(d)) over the REAL tokens of
// sequence b (the packed layout holds no padded keys, which is what the reference's -inf key-padding mask leaves).
//
// An inference-time visualisation call, not a training hot op: the result (S^2 fp32 per sequence and head) is as large as
// the score matrix a flash kernel exists to avoid, so the kernel is bound by writing it.  One CTA = (sequence, head,
// 16 query rows): scores of the 16 rows against all keys go to shared memory (thread = one key, 16 dot products against
// broadcast q rows), then one warp per row does max / exp / sum and streams the normalised row out with coalesced stores.
#include "common.cuh"
#include "chadavit_b200.h"

namespace cb {

constexpr int PROB_ROWS = 16;

template <int HD>
__global__ void __launch_bounds__(256) attn_probs_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ cu, int H, int s_max,
                                                         float scale, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* q_s = sm;                       // [PROB_ROWS][HD]
  float* s_s = sm + PROB_ROWS * HD;      // [PROB_ROWS][s_max]
  const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * PROB_ROWS;
  const int t0 = cu[b], S = cu[b + 1] - t0;
  const int D = H * HD, ld = 3 * D;
  float* o = out + (((long)b * H + h) * s_max + r0) * s_max;
  const int nrow = min(PROB_ROWS, s_max - r0);
  if (r0 >= S) {                         // query rows beyond this sequence: zeros
    for (int i = threadIdx.x; i < nrow * s_max; i += 256) o[i] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < PROB_ROWS * HD; i += 256) {
    const int r = i / HD, c = i % HD;
    q_s[i] = (r0 + r < S) ? __bfloat162float(qkv[(long)(t0 + r0 + r) * ld + h * HD + c]) * scale : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < S; j += 256) {
    const uint4* kp = reinterpret_cast<const uint4*>(qkv + (long)(t0 + j) * ld + D + h * HD);   // 16 B aligned: D, HD multiples of 8
    float acc[PROB_ROWS];
#pragma unroll
    for (int r = 0; r < PROB_ROWS; ++r) acc[r] = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < HD / 8; ++c8) {
      const uint4 kv = __ldg(kp + c8);
      const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 k2 = unpack_bf16(w[e]);
#pragma unroll
        for (int r = 0; r < PROB_ROWS; ++r) {
          acc[r] = fmaf(q_s[r * HD + c8 * 8 + 2 * e], k2.x, acc[r]);
          acc[r] = fmaf(q_s[r * HD + c8 * 8 + 2 * e + 1], k2.y, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < PROB_ROWS; ++r) s_s[r * s_max + j] = acc[r];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nrow; r += 8) {
    float* orow = o + (long)r * s_max;
    if (r0 + r >= S) {
      for (int j = lane; j < s_max; j += 32) orow[j] = 0.f;
      continue;
    }
    const float* srow = s_s + r * s_max;
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) m = fmaxf(m, srow[j]);
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o2));
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) sum += __expf(srow[j] - m);
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o2);
    const float inv = 1.f / sum;
    for (int j = lane; j < s_max; j += 32) orow[j] = (j < S) ? __expf(srow[j] - m) * inv : 0.f;
  }
}

template <int HD>
static int launch_probs(const __nv_bfloat16* qkv, const int* cu, int nseq, int H, int s_max, float scale, float* out, cudaStream_t st) {
  const size_t smem = (size_t)(PROB_ROWS * HD + PROB_ROWS * s_max) * sizeof(float);
  CB_CHECK(smem <= 200 * 1024, "attn_probs: max sequence length %d needs %zu B of shared memory (limit 200 KB)", s_max, smem);
  CB_CUDA(cudaFuncSetAttribute(attn_probs_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((s_max + PROB_ROWS - 1) / PROB_ROWS, H, nseq);
  attn_probs_kernel<HD><<<grid, 256, smem, st>>>(qkv, cu, H, s_max, scale, out);
  CB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cb

using namespace cb;

extern "C" int cb_attn_probs(const void* qkv, const int* cu_seqlens, int nseq, int num_heads, int head_dim, int max_seqlen, float scale,
                             float* out, void* stream) {
  CB_CHECK(nseq > 0 && num_heads > 0 && max_seqlen > 0, "attn_probs: nseq=%d heads=%d max_seqlen=%d", nseq, num_heads, max_seqlen);
  CB_CHECK(nseq <= 65535 && num_heads <= 65535, "attn_probs: nseq / heads exceed the grid limits");
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 16: return launch_probs<16>(q, cu_seqlens, nseq, num_heads, max_seqlen, scale, out, st);
    case 32: return launch_probs<32>(q, cu_seqlens, nseq, num_heads, max_seqlen, scale, out, st);
    case 64: return launch_probs<64>(q, cu_seqlens, nseq, num_heads, max_seqlen, scale, out, st);
    case 96: return launch_probs<96>(q, cu_seqlens, nseq, num_heads, max_seqlen, scale, out, st);
    case 128: return launch_probs<128>(q, cu_seqlens, nseq, num_heads, max_seqlen, scale, out, st);
    default: CB_CHECK(false, "attn_probs: unsupported head_dim %d (16/32/64/96/128)", head_dim);
  }
  return 1;
}
