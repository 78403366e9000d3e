// DINO head / loss / center / EMA / optimizer kernels.  All HBM-bound,
// fp32 arithmetic, 16-byte vector accesses; the head's GEMMs themselves go through gemm.cu.
//   gelu fwd/bwd (exact erf, nn.GELU)                      src/methods/dino.py:65-75
//   L2 row normalise fwd/bwd (F.normalize, eps 1e-12)      src/methods/dino.py:109
//   weight-norm fwd/bwd (w = g * v / ||v||_row)            src/methods/dino.py:78-81
//   fused DINO loss fwd+bwd                                src/losses/dino.py:81-99
//   center EMA                                             src/losses/dino.py:111-118
//   teacher EMA over a flat arena                          src/utils/momentum.py:73-74
//   fused AdamW (+ optional teacher EMA, + bf16 shadows)   torch.optim.AdamW as configured in src/methods/base.py:416-440
#include "common.cuh"
#include "chadavit_b200.h"

namespace cb {

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// out_bf16 = gelu(pre)   (pre fp32, kept by the caller for the backward)
__global__ void gelu_fwd_kernel(const float* __restrict__ pre, __nv_bfloat16* __restrict__ out, long n) {
  const long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(pre + i);
  *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16(gelu_f(v.x), gelu_f(v.y)), pack_bf16(gelu_f(v.z), gelu_f(v.w)));
}
// dpre_bf16 = dact * gelu'(pre)
__global__ void gelu_bwd_kernel(const float* __restrict__ dact, const float* __restrict__ pre, __nv_bfloat16* __restrict__ dpre, long n) {
  const long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 d = *reinterpret_cast<const float4*>(dact + i), v = *reinterpret_cast<const float4*>(pre + i);
  *reinterpret_cast<uint2*>(dpre + i) = make_uint2(pack_bf16(d.x * gelu_grad_f(v.x), d.y * gelu_grad_f(v.y)),
                                                    pack_bf16(d.z * gelu_grad_f(v.z), d.w * gelu_grad_f(v.w)));
}

// nn.BatchNorm1d + nn.GELU of the projector (DINOHead(use_bn=True), src/methods/dino.py:66-73).  Block = 32 columns x 8 row
// groups; R (= views x batch rows) is a few hundred, so two passes over the L2-resident column strip are cheaper than anything
// clever.  Training mode: batch statistics (biased variance for the normalisation, unbiased for running_var, momentum as
// torch: running = (1 - m) running + m batch); eval mode: running statistics.
__device__ __forceinline__ float bn_block_sum(float v, float (*sh)[33]) {   // sum over the 8 row groups, result in every thread
  sh[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];
  __syncthreads();
  return s;
}
__global__ void __launch_bounds__(256) bn_gelu_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                                          float eps, int training, float* __restrict__ bn_out, __nv_bfloat16* __restrict__ act,
                                                          float* __restrict__ save_mean, float* __restrict__ save_invstd, int R, int C) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  float mean, invstd;
  if (training) {
    float s = 0.f;
    if (ok) for (int r = threadIdx.y; r < R; r += 8) s += pre[(long)r * C + c];
    mean = bn_block_sum(s, sh) / R;
    float q = 0.f;
    if (ok) for (int r = threadIdx.y; r < R; r += 8) { const float d = pre[(long)r * C + c] - mean; q += d * d; }
    const float var = bn_block_sum(q, sh) / R;
    invstd = rsqrtf(var + eps);
    if (ok && threadIdx.y == 0) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (R > 1 ? (float)R / (R - 1) : 1.f);
    }
  } else {
    mean = ok ? running_mean[c] : 0.f;
    invstd = ok ? rsqrtf(running_var[c] + eps) : 0.f;
  }
  if (!ok) return;
  if (threadIdx.y == 0 && save_mean) { save_mean[c] = mean; save_invstd[c] = invstd; }
  const float g = gamma[c], b = beta[c];
  for (int r = threadIdx.y; r < R; r += 8) {
    const float y = (pre[(long)r * C + c] - mean) * invstd * g + b;
    if (bn_out) bn_out[(long)r * C + c] = y;
    act[(long)r * C + c] = __float2bfloat16(gelu_f(y));
  }
}
// dpre (bf16) from dact = d loss / d gelu output; dgamma / dbeta accumulate (+=)
__global__ void __launch_bounds__(256) bn_gelu_bwd_kernel(const float* __restrict__ dact, const float* __restrict__ bn_out, const float* __restrict__ pre,
                                                          const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                          const float* __restrict__ save_invstd, int training, __nv_bfloat16* __restrict__ dpre,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int R, int C) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  const float mean = ok ? save_mean[c] : 0.f, invstd = ok ? save_invstd[c] : 0.f, g = ok ? gamma[c] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (ok)
    for (int r = threadIdx.y; r < R; r += 8) {
      const long i = (long)r * C + c;
      const float d = dact[i] * gelu_grad_f(bn_out[i]);
      s1 += d;
      s2 += d * (pre[i] - mean) * invstd;
    }
  s1 = bn_block_sum(s1, sh);
  s2 = bn_block_sum(s2, sh);
  if (!ok) return;
  if (threadIdx.y == 0) { dbeta[c] += s1; dgamma[c] += s2; }
  const float m1 = training ? s1 / R : 0.f, m2 = training ? s2 / R : 0.f;
  for (int r = threadIdx.y; r < R; r += 8) {
    const long i = (long)r * C + c;
    const float d = dact[i] * gelu_grad_f(bn_out[i]);
    const float xh = (pre[i] - mean) * invstd;
    dpre[i] = __float2bfloat16(g * invstd * (d - m1 - xh * m2));
  }
}

// One warp per row.  fwd: out = x / max(||x||, eps) (bf16), inv[r] = 1/max(||x||, eps).
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, float* __restrict__ inv, int rows, int C, float eps) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = x[(long)r * C + c]; s += v * v; }
  const float iv = 1.f / fmaxf(sqrtf(warp_sum(s)), eps);
  for (int c = lane; c < C; c += 32) out[(long)r * C + c] = __float2bfloat16(x[(long)r * C + c] * iv);
  if (lane == 0) inv[r] = iv;
}
// bwd: dx = inv * (dy - xn * (xn . dy)),  xn = x * inv          (dx bf16: it is the next GEMM operand)
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ inv,
                                  __nv_bfloat16* __restrict__ dx, int rows, int C) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float iv = inv[r];
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot += x[(long)r * C + c] * iv * dy[(long)r * C + c];
  dot = warp_sum(dot);
  for (int c = lane; c < C; c += 32) dx[(long)r * C + c] = __float2bfloat16(iv * (dy[(long)r * C + c] - x[(long)r * C + c] * iv * dot));
}

// weight norm (dim=0): w[k,:] = g[k] * v[k,:] / ||v[k,:]||  -> bf16 operand of the prototype GEMM
__global__ void weightnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, __nv_bfloat16* __restrict__ w,
                                      float* __restrict__ inv_norm, int K, int C) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= K) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float t = v[(long)r * C + c]; s += t * t; }
  const float iv = rsqrtf(warp_sum(s));
  const float sc = g[r] * iv;
  for (int c = lane; c < C; c += 32) w[(long)r * C + c] = __float2bfloat16(v[(long)r * C + c] * sc);
  if (lane == 0) inv_norm[r] = iv;
}
// dv += g*inv*(dw - vn (vn.dw)),  dg += vn.dw  (dg may be NULL when weight_g is frozen: norm_last_layer)
__global__ void weightnorm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                                      const float* __restrict__ inv_norm, float* __restrict__ dv, float* __restrict__ dg, int K, int C) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= K) return;
  const float iv = inv_norm[r];
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot += v[(long)r * C + c] * iv * dw[(long)r * C + c];
  dot = warp_sum(dot);
  const float sc = g[r] * iv;
  for (int c = lane; c < C; c += 32) dv[(long)r * C + c] += sc * (dw[(long)r * C + c] - v[(long)r * C + c] * iv * dot);
  if (dg && lane == 0) dg[r] += dot;
}

// ------------------------------------------------------------------------------------------------ DINO loss
__device__ __forceinline__ void online_merge(float& m, float& s, float x) {
  if (x > m) { s = s * __expf(m - x) + 1.f; m = x; } else { s += __expf(x - m); }
}
__device__ void block_reduce_ms(float& m, float& s, float* sh) {
  // combine (max, sumexp) pairs across the block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
    m = mm;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { sh[w] = m; sh[32 + w] = s; }
  __syncthreads();
  float mm = -INFINITY;
  for (int i = 0; i < nw; ++i) mm = fmaxf(mm, sh[i]);
  float ss = 0.f;
  for (int i = 0; i < nw; ++i) ss += (sh[i] == -INFINITY ? 0.f : sh[32 + i] * __expf(sh[i] - mm));
  m = mm; s = ss;
}
__device__ float block_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

// One CTA per batch row b.  student [V*B, K] (view v = rows v*B..), teacher [2*B, K], center [K].
//   q_i = softmax((t_i - c)/tt);  logp_v = log_softmax(s_v/ts)
//   loss += (1/(n_terms*B)) * sum_{i} sum_{v != i} ( lse_v - (1/ts) * sum_k q_i[k] s_v[k] )
//   ds_v[k] = ( n_q(v) * p_v[k] - sum_{i != v} q_i[k] ) / (ts * n_terms * B)
__global__ void __launch_bounds__(256) dino_loss_kernel(const float* __restrict__ student, const float* __restrict__ teacher,
                                                        const float* __restrict__ center, float* __restrict__ loss,
                                                        float* __restrict__ dstudent32, __nv_bfloat16* __restrict__ dstudent16,
                                                        int B, int K, int V, float inv_ts, float inv_tt) {
  __shared__ float sh[64];
  const int b = blockIdx.x;
  const int n_terms = 2 * V - 2;
  // teacher statistics
  float tm[2], tz[2];
  for (int i = 0; i < 2; ++i) {
    const float* t = teacher + ((long)i * B + b) * K;
    float m = -INFINITY, s = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) online_merge(m, s, (t[k] - center[k]) * inv_tt);
    block_reduce_ms(m, s, sh);
    tm[i] = m; tz[i] = 1.f / s;
  }
  const float* t0 = teacher + (long)b * K;
  const float* t1 = teacher + ((long)B + b) * K;
  float loss_acc = 0.f;
  const float gscale = inv_ts / (float)(n_terms * B);
  for (int v = 0; v < V; ++v) {
    const float* sv = student + ((long)v * B + b) * K;
    float m = -INFINITY, s = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) online_merge(m, s, sv[k] * inv_ts);
    block_reduce_ms(m, s, sh);
    const float lse = m + __logf(s), inv_z = 1.f / s;
    const bool use0 = v != 0, use1 = v != 1;
    const float nq = (use0 ? 1.f : 0.f) + (use1 ? 1.f : 0.f);
    float dot0 = 0.f, dot1 = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float sk = sv[k] * inv_ts, c = center[k];
      const float q0 = use0 ? __expf((t0[k] - c) * inv_tt - tm[0]) * tz[0] : 0.f;
      const float q1 = use1 ? __expf((t1[k] - c) * inv_tt - tm[1]) * tz[1] : 0.f;
      dot0 += q0 * sk; dot1 += q1 * sk;
      const float gk = (nq * __expf(sk - m) * inv_z - q0 - q1) * gscale;
      if (dstudent32) dstudent32[((long)v * B + b) * K + k] = gk;
      if (dstudent16) dstudent16[((long)v * B + b) * K + k] = __float2bfloat16(gk);
    }
    dot0 = block_reduce_sum(dot0, sh);
    dot1 = block_reduce_sum(dot1, sh);
    loss_acc += nq * lse - dot0 - dot1;
  }
  if (threadIdx.x == 0) atomicAdd(loss, loss_acc / (float)(n_terms * B));
}

// out[k] (+)= sum_r x[r, k]   (fp32 rows: teacher logits -> batch center)
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int R, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float a = 0.f;
  for (int r = 0; r < R; ++r) a += x[(long)r * K + k];
  out[k] = a;
}
// center = center * m + batch_sum * scale * (1 - m)
__global__ void center_ema_kernel(float* __restrict__ center, const float* __restrict__ batch_sum, float scale, float m, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) center[k] = center[k] * m + batch_sum[k] * scale * (1.f - m);
}

// ------------------------------------------------------------------------------------------------ EMA / AdamW over flat arenas
__global__ void ema_kernel(float* __restrict__ mom, const float* __restrict__ online, __nv_bfloat16* __restrict__ mom16, float tau, long n) {
  const long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 m = *reinterpret_cast<float4*>(mom + i);
  const float4 o = *reinterpret_cast<const float4*>(online + i);
  const float u = 1.f - tau;
  m.x = tau * m.x + u * o.x; m.y = tau * m.y + u * o.y; m.z = tau * m.z + u * o.z; m.w = tau * m.w + u * o.w;
  *reinterpret_cast<float4*>(mom + i) = m;
  if (mom16) *reinterpret_cast<uint2*>(mom16 + i) = make_uint2(pack_bf16(m.x, m.y), pack_bf16(m.z, m.w));
}

struct AdamArgs {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale, tau, bc1_late, bc2_sqrt_late;
};
// flags[i]: bit0 = apply weight decay, bit1 = frozen (no update), bit4 = the parameter started receiving gradients late (its
// own step count, torch.optim.AdamW's per-parameter state['step']: second bias-correction pair).  Optional fused teacher EMA
// and bf16 shadows.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             const uint8_t* __restrict__ flags, __nv_bfloat16* __restrict__ p16, float* __restrict__ teacher,
                             __nv_bfloat16* __restrict__ teacher16, AdamArgs a, const float* __restrict__ dev_hyper, long n) {
  const long i0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  if (dev_hyper) {   // CUDA-graph replay: per-step scalars live in device memory {lr, 1-beta1^t, sqrt(1-beta2^t), tau, late pair}
    a.lr = __ldg(dev_hyper); a.bc1 = __ldg(dev_hyper + 1); a.bc2_sqrt = __ldg(dev_hyper + 2); a.tau = __ldg(dev_hyper + 3);
    a.bc1_late = __ldg(dev_hyper + 4); a.bc2_sqrt_late = __ldg(dev_hyper + 5);
  }
  float4 P = *reinterpret_cast<float4*>(p + i0);
  const float4 G = *reinterpret_cast<const float4*>(g + i0);
  float4 M = *reinterpret_cast<float4*>(m + i0), Vv = *reinterpret_cast<float4*>(v + i0);
  uint32_t fl = flags ? *reinterpret_cast<const uint32_t*>(flags + i0) : 0x01010101u;
  float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &Vv.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t f = (fl >> (8 * j)) & 0xFF;
    if (f & 2) continue;
    const float gr = gg[j] * a.grad_scale;
    float x = pp[j];
    if (f & 1) x -= a.lr * a.wd * x;
    mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gr;
    vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gr * gr;
    const float bc1 = (f & 16) ? a.bc1_late : a.bc1, bc2s = (f & 16) ? a.bc2_sqrt_late : a.bc2_sqrt;
    x -= (a.lr / bc1) * mm[j] / (sqrtf(vv[j]) / bc2s + a.eps);
    pp[j] = x;
  }
  *reinterpret_cast<float4*>(p + i0) = P;
  *reinterpret_cast<float4*>(m + i0) = M;
  *reinterpret_cast<float4*>(v + i0) = Vv;
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

extern "C" int cb_gelu_fwd(const float* pre, void* out, long n, void* stream) {
  CB_CHECK(n > 0 && n % 4 == 0, "gelu_fwd: n=%ld must be a positive multiple of 4", n);
  gelu_fwd_kernel<<<blocks4(n), 256, 0, STREAM>>>(pre, BFM(out), n);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_gelu_bwd(const float* dact, const float* pre, void* dpre, long n, void* stream) {
  CB_CHECK(n > 0 && n % 4 == 0, "gelu_bwd: n=%ld must be a positive multiple of 4", n);
  gelu_bwd_kernel<<<blocks4(n), 256, 0, STREAM>>>(dact, pre, BFM(dpre), n);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_bn_gelu_fwd(const float* pre, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                              float eps, int training, float* bn_out, void* act, float* save_mean, float* save_invstd, int R, int C,
                              void* stream) {
  CB_CHECK(R > 0 && C > 0 && pre && gamma && beta && running_mean && running_var && act, "bn_gelu_fwd: R=%d C=%d or a NULL pointer", R, C);
  bn_gelu_fwd_kernel<<<(C + 31) / 32, dim3(32, 8), 0, STREAM>>>(pre, gamma, beta, running_mean, running_var, momentum, eps, training, bn_out,
                                                                 BFM(act), save_mean, save_invstd, R, C);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_bn_gelu_bwd(const float* dact, const float* bn_out, const float* pre, const float* gamma, const float* save_mean,
                              const float* save_invstd, int training, void* dpre, float* dgamma, float* dbeta, int R, int C, void* stream) {
  CB_CHECK(R > 0 && C > 0 && dact && bn_out && pre && gamma && save_mean && save_invstd && dpre && dgamma && dbeta,
           "bn_gelu_bwd: R=%d C=%d or a NULL pointer", R, C);
  bn_gelu_bwd_kernel<<<(C + 31) / 32, dim3(32, 8), 0, STREAM>>>(dact, bn_out, pre, gamma, save_mean, save_invstd, training, BFM(dpre), dgamma,
                                                                 dbeta, R, C);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_l2norm_fwd(const float* x, void* out, float* inv, int rows, int C, float eps, void* stream) {
  CB_CHECK(rows > 0 && C > 0, "l2norm_fwd: rows=%d C=%d", rows, C);
  l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, STREAM>>>(x, BFM(out), inv, rows, C, eps);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_l2norm_bwd(const float* dy, const float* x, const float* inv, void* dx, int rows, int C, void* stream) {
  CB_CHECK(rows > 0 && C > 0, "l2norm_bwd: rows=%d C=%d", rows, C);
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, STREAM>>>(dy, x, inv, BFM(dx), rows, C);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_weightnorm_fwd(const float* v, const float* g, void* w, float* inv_norm, int K, int C, void* stream) {
  CB_CHECK(K > 0 && C > 0, "weightnorm_fwd: K=%d C=%d", K, C);
  weightnorm_fwd_kernel<<<(K + 7) / 8, 256, 0, STREAM>>>(v, g, BFM(w), inv_norm, K, C);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg, int K,
                                 int C, void* stream) {
  CB_CHECK(K > 0 && C > 0, "weightnorm_bwd: K=%d C=%d", K, C);
  weightnorm_bwd_kernel<<<(K + 7) / 8, 256, 0, STREAM>>>(dw, v, g, inv_norm, dv, dg, K, C);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_dino_loss_fwd_bwd(const float* student, const float* teacher, const float* center, float* loss, float* dstudent_f32,
                                    void* dstudent_bf16, int B, int K, int V, float student_temp, float teacher_temp, void* stream) {
  CB_CHECK(B > 0 && K > 0 && V >= 2, "dino_loss: B=%d K=%d V=%d (needs V >= 2 views)", B, K, V);
  CB_CHECK(student_temp > 0.f && teacher_temp > 0.f, "dino_loss: temperatures must be positive");
  CB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), STREAM));
  dino_loss_kernel<<<B, 256, 0, STREAM>>>(student, teacher, center, loss, dstudent_f32, BFM(dstudent_bf16), B, K, V, 1.f / student_temp,
                                           1.f / teacher_temp);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_colsum_f32(const float* x, float* out, int R, int K, void* stream) {
  CB_CHECK(R > 0 && K > 0, "colsum_f32: R=%d K=%d", R, K);
  colsum_f32_kernel<<<(K + 127) / 128, 128, 0, STREAM>>>(x, out, R, K);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_dino_center_ema(float* center, const float* batch_sum, float scale, float momentum, int K, void* stream) {
  CB_CHECK(K > 0, "center_ema: K=%d", K);
  center_ema_kernel<<<(K + 255) / 256, 256, 0, STREAM>>>(center, batch_sum, scale, momentum, K);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_ema_update(float* momentum, const float* online, void* momentum_bf16, float tau, long n, void* stream) {
  CB_CHECK(n > 0 && n % 4 == 0, "ema_update: n=%ld must be a positive multiple of 4", n);
  ema_kernel<<<blocks4(n), 256, 0, STREAM>>>(momentum, online, BFM(momentum_bf16), tau, n);
  CB_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int cb_adamw_step(float* p, const float* g, float* m, float* v, const unsigned char* flags, void* p_bf16, float* teacher,
                             void* teacher_bf16, long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             int step_late, float grad_scale, float tau, const float* dev_hyper, void* stream) {
  CB_CHECK(n > 0 && n % 4 == 0 && step >= 1, "adamw_step: n=%ld step=%d", n, step);
  if (step_late < 1) step_late = 1;
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.grad_scale = grad_scale; a.tau = tau;
  a.bc1 = 1.f - powf(beta1, (float)step);
  a.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  a.bc1_late = 1.f - powf(beta1, (float)step_late);
  a.bc2_sqrt_late = sqrtf(1.f - powf(beta2, (float)step_late));
  adamw_kernel<<<blocks4(n), 256, 0, STREAM>>>(p, g, m, v, flags, BFM(p_bf16), teacher, BFM(teacher_bf16), a, dev_hyper, n);
  CB_CUDA(cudaGetLastError());
  return 0;
}
