// Per-parameter norms, DINO gradient clipping and the LARS step over flat parameter arenas.
//
//   reference: src/utils/lars.py:113-167 (LARS.step), src/methods/dino.py:249-261 (dino_clip_gradients),
//              src/methods/base.py:416-440 (optimizer construction, weight-decay groups)
//
// All three are HBM streams.  Parameters sit in one fp32 arena at 64-element aligned offsets (arena.py), so every
// 64-element block belongs to exactly one parameter ("segment"); padding holds zeros in both the parameter and the
// gradient arena and contributes nothing to a norm.
//
//   param_norms  pass 1: one partial (sum p^2, sum g^2) per 64-element block  (8 B/param read, n/8 B written)
//                pass 2: one CTA per parameter adds its partials in a FIXED order -> ||p||, ||g||, clip coefficient.
//                No atomics: data-parallel replicas must derive bit-identical updates from the all-reduced gradient.
//   lars_step    p, g, momentum buffer (+ teacher EMA + bf16 shadows) in one pass: 28 B/param + 12 (EMA) + 4 (shadows).
#include "common.cuh"
#include "chadavit_b200.h"

namespace cb {

__global__ void sqnorm_partial_kernel(const float* __restrict__ p, const float* __restrict__ g, float2* __restrict__ partial, long n) {
  const long i0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  float sp = 0.f, sg = 0.f;
  if (i0 < n) {
    const float4 P = *reinterpret_cast<const float4*>(p + i0);
    const float4 G = *reinterpret_cast<const float4*>(g + i0);
    sp = P.x * P.x + P.y * P.y + P.z * P.z + P.w * P.w;
    sg = G.x * G.x + G.y * G.y + G.z * G.z + G.w * G.w;
  }
  // 16 lanes = one 64-element block; xor-shuffles below 16 stay inside the half warp (fixed tree -> deterministic)
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
  }
  if ((threadIdx.x & 15) == 0 && i0 < n) partial[i0 >> 6] = make_float2(sp, sg);
}

// norms[3*s] = ||p_s||, norms[3*s+1] = ||g_s * grad_scale * coef_s||, norms[3*s+2] = coef_s
// coef_s = min(1, clip / (||g_s * grad_scale|| + 1e-6)) where seg_clip[s] != 0 and clip > 0 (dino.py:256-261), else 1.
__global__ void sqnorm_final_kernel(const float2* __restrict__ partial, const int* __restrict__ seg_start_block,
                                    const uint8_t* __restrict__ seg_clip, float* __restrict__ norms, float grad_scale, float clip) {
  __shared__ float shp[256], shg[256];
  const int s = blockIdx.x;
  const int b0 = seg_start_block[s], b1 = seg_start_block[s + 1];
  float sp = 0.f, sg = 0.f;
  for (int b = b0 + threadIdx.x; b < b1; b += 256) {
    const float2 v = partial[b];
    sp += v.x;
    sg += v.y;
  }
  shp[threadIdx.x] = sp;
  shg[threadIdx.x] = sg;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      shp[threadIdx.x] += shp[threadIdx.x + o];
      shg[threadIdx.x] += shg[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float pn = sqrtf(shp[0]);
    float gn = sqrtf(shg[0]) * grad_scale;
    float coef = 1.f;
    if (clip > 0.f && seg_clip && seg_clip[s]) {
      const float c = clip / (gn + 1e-6f);
      if (c < 1.f) coef = c;
    }
    norms[3 * s] = pn;
    norms[3 * s + 1] = gn * coef;
    norms[3 * s + 2] = coef;
  }
}

// g *= coef of its parameter (used in front of optimizers that do not take the coefficient themselves)
__global__ void scale_grads_kernel(float* __restrict__ g, const int* __restrict__ seg_of_block, const float* __restrict__ norms, long n) {
  const long i0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  const float c = __ldg(norms + 3 * __ldg(seg_of_block + (i0 >> 6)) + 2);
  if (c == 1.f) return;
  float4 G = *reinterpret_cast<float4*>(g + i0);
  G.x *= c; G.y *= c; G.z *= c; G.w *= c;
  *reinterpret_cast<float4*>(g + i0) = G;
}

struct LarsArgs {
  float lr, momentum, dampening, wd, eta, eps, grad_scale, tau;
  int nesterov, clip_lr;
};
// flags[i]: bit0 weight decay applies (group weight_decay, else 0), bit1 frozen / no gradient (skipped, lars.py:128-129),
//           bit2 layer-wise adaptation applies (p.ndim != 1 or not exclude_bias_n_norm, lars.py:136),
//           bit3 first update of this parameter (momentum buffer := d_p, lars.py:151-152)
__global__ void lars_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, const uint8_t* __restrict__ flags,
                            const int* __restrict__ seg_of_block, const float* __restrict__ norms, __nv_bfloat16* __restrict__ p16,
                            float* __restrict__ teacher, __nv_bfloat16* __restrict__ teacher16, LarsArgs a,
                            const float* __restrict__ dev_hyper, long n) {
  const long i0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  if (dev_hyper) {   // CUDA-graph replay: {lr, -, -, tau} in device memory (same slots as the AdamW kernel)
    a.lr = __ldg(dev_hyper);
    a.tau = __ldg(dev_hyper + 3);
  }
  const int seg = __ldg(seg_of_block + (i0 >> 6));
  const float pn = __ldg(norms + 3 * seg), gn = __ldg(norms + 3 * seg + 1), coef = __ldg(norms + 3 * seg + 2);
  float4 P = *reinterpret_cast<float4*>(p + i0);
  const float4 G = *reinterpret_cast<const float4*>(g + i0);
  float4 Bf = *reinterpret_cast<float4*>(buf + i0);
  const uint32_t fl = *reinterpret_cast<const uint32_t*>(flags + i0);
  float* pp = &P.x; const float* gg = &G.x; float* bb = &Bf.x;
  const float gs = a.grad_scale * coef;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t f = (fl >> (8 * j)) & 0xFF;
    if (f & 2) continue;
    float d = gg[j] * gs;
    if ((f & 4) && pn != 0.f && gn != 0.f) {
      const float wd = (f & 1) ? a.wd : 0.f;
      float l = pn / (gn + pn * wd + a.eps) * a.eta;
      if (a.clip_lr) l = fminf(l / a.lr, 1.f);
      d = (d + wd * pp[j]) * l;
    }
    if (a.momentum != 0.f) {
      const float b = (f & 8) ? d : a.momentum * bb[j] + (1.f - a.dampening) * d;
      bb[j] = b;
      d = a.nesterov ? d + a.momentum * b : b;
    }
    pp[j] -= a.lr * d;
  }
  *reinterpret_cast<float4*>(p + i0) = P;
  *reinterpret_cast<float4*>(buf + i0) = Bf;
  if (p16) *reinterpret_cast<uint2*>(p16 + i0) = make_uint2(pack_bf16(P.x, P.y), pack_bf16(P.z, P.w));
  if (teacher) {
    float4 T = *reinterpret_cast<float4*>(teacher + i0);
    const float u = 1.f - a.tau;
    T.x = a.tau * T.x + u * P.x; T.y = a.tau * T.y + u * P.y; T.z = a.tau * T.z + u * P.z; T.w = a.tau * T.w + u * P.w;
    *reinterpret_cast<float4*>(teacher + i0) = T;
    if (teacher16) *reinterpret_cast<uint2*>(teacher16 + i0) = make_uint2(pack_bf16(T.x, T.y), pack_bf16(T.z, T.w));
  }
}

}  // namespace cb

using namespace cb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define BFM(p) reinterpret_cast<__nv_bfloat16*>(p)
static inline unsigned blocks4(long n) { return (unsigned)((n + 1023) / 1024); }

extern "C" int cb_param_norms(const float* p, const float* g, const int* seg_start_block, const unsigned char* seg_clip, float* partial,
                              float* norms, long n, int nseg, float grad_scale, float clip, void* stream) {
  CB_CHECK(n > 0 && n % 64 == 0 && nseg > 0, "param_norms: n=%ld (multiple of 64) nseg=%d", n, nseg);
  CB_CHECK(((uintptr_t)partial & 7) == 0, "param_norms: partial must be 8-byte aligned");
  sqnorm_partial_kernel<<<blocks4(n), 256, 0, STREAM>>>(p, g, reinterpret_cast<float2*>(partial), n);
  CB_CUDA(cudaGetLastError());
  sqnorm_final_kernel<<<nseg, 256, 0, STREAM>>>(reinterpret_cast<const float2*>(partial), seg_start_block, seg_clip, norms, grad_scale, clip);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_scale_grads(float* g, const int* seg_of_block, const float* norms, long n, void* stream) {
  CB_CHECK(n > 0 && n % 64 == 0, "scale_grads: n=%ld must be a positive multiple of 64", n);
  scale_grads_kernel<<<blocks4(n), 256, 0, STREAM>>>(g, seg_of_block, norms, n);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_lars_step(float* p, const float* g, float* buf, const unsigned char* flags, const int* seg_of_block, const float* norms,
                            void* p_bf16, float* teacher, void* teacher_bf16, long n, float lr, float momentum, float dampening,
                            int nesterov, float weight_decay, float eta, float eps, int clip_lr, float grad_scale, float tau,
                            const float* dev_hyper, void* stream) {
  CB_CHECK(n > 0 && n % 64 == 0, "lars_step: n=%ld must be a positive multiple of 64", n);
  CB_CHECK(flags && seg_of_block && norms, "lars_step: flags, seg_of_block and norms are required");
  CB_CHECK(!nesterov || (momentum > 0.f && dampening == 0.f), "lars_step: Nesterov momentum requires a momentum and zero dampening");
  LarsArgs a;
  a.lr = lr; a.momentum = momentum; a.dampening = dampening; a.wd = weight_decay; a.eta = eta; a.eps = eps;
  a.grad_scale = grad_scale; a.tau = tau; a.nesterov = nesterov; a.clip_lr = clip_lr;
  lars_kernel<<<blocks4(n), 256, 0, STREAM>>>(p, g, buf, flags, seg_of_block, norms, BFM(p_bf16), teacher, BFM(teacher_bf16), a, dev_hyper, n);
  CB_CUDA(cudaGetLastError());
  return 0;
}
