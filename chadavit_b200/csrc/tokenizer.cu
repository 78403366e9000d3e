// Channel-adaptive tokenizer: replaces TokenLearner.forward + channel_aware_tokenization
// (src/backbones/vit/chada_vit.py:118-134, 219-270).
//   forward : fp32 pixels (ΣC,1,H,W) --im2col+cast--> bf16 patch rows laid out in PACKED TOKEN ORDER (CLS rows zero)
//             --tcgen05 GEMM [T,P²]x[P²,D] with fused epilogue (+bias +pos[p] +channel_token[c], CLS = cls+pos[0])-->
//             tokens (T,D) bf16.  Padded channels are never materialised (no FLOPs, no bytes).
//   backward: dW_pe via the split-K MN-major GEMM on the same packed patch rows, embedding gradients by segmented sums.
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"

namespace cb {

// One thread moves 8 consecutive pixels of one image row (32 B read, 16 B write) to row  g*N + p + chan_img[g] + 1.
__global__ void im2col_packed_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, const int* __restrict__ chan_img,
                                     const int* __restrict__ cu, int G, int B, int H, int W, int P) {
  const int hp = H / P, wp = W / P, PP = P * P;
  const int chunks_per_row = (wp * P) / 8;
  const long total = (long)G * hp * P * chunks_per_row;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = (int)(i % chunks_per_row);
    const long t = i / chunks_per_row;
    const int y = (int)(t % (hp * P));
    const int gimg = (int)(t / (hp * P));
    const int xpix = ch * 8;
    const float* src = x + ((long)gimg * H + y) * W + xpix;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
    const int py = y / P, r = y - py * P, px = xpix / P, c = xpix - px * P;
    const long row = (long)gimg * hp * wp + py * wp + px + (chan_img ? chan_img[gimg] + 1 : 0);
    *reinterpret_cast<uint4*>(out + row * PP + r * P + c) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  if (cu) {  // zero the CLS rows so they contribute nothing to the GEMM / to dW
    const long ztotal = (long)B * (PP / 8);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < ztotal; i += stride) {
      const int b = (int)(i / (PP / 8)), c8 = (int)(i % (PP / 8));
      *reinterpret_cast<uint4*>(out + (long)cu[b] * PP + c8 * 8) = make_uint4(0, 0, 0, 0);
    }
  }
}

// Embedding gradients from dTok (T,D) bf16.  grid = npatch + G + 1 blocks, blockDim = D/2 threads (2 columns each):
//   block p < npatch           : dpos_patch[p]  += sum_g dTok[row(g,p)]
//   block npatch + g           : dchan[c_g]     += sum_p dTok[row(g,p)],  dbias += same
//   block npatch + G           : dcls_tok, dpos0 += sum_b dTok[cu[b]]
__global__ void tokenizer_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dtok, const int* __restrict__ chan_img,
                                           const int* __restrict__ chan_idx, const int* __restrict__ cu, int G, int B, int npatch,
                                           int D, float* __restrict__ dpos_patch, float* __restrict__ dchan, float* __restrict__ dbias,
                                           float* __restrict__ dcls_tok, float* __restrict__ dpos0) {
  const int col = threadIdx.x * 2;
  if (col >= D) return;
  float a0 = 0.f, a1 = 0.f;
  const int blk = blockIdx.x;
  if (blk < npatch) {
    for (int g = 0; g < G; ++g) {
      const long row = (long)g * npatch + blk + chan_img[g] + 1;
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(dtok + row * D + col));
      a0 += f.x; a1 += f.y;
    }
    atomicAdd(dpos_patch + (long)blk * D + col, a0); atomicAdd(dpos_patch + (long)blk * D + col + 1, a1);
  } else if (blk < npatch + G) {
    const int g = blk - npatch;
    const long row0 = (long)g * npatch + chan_img[g] + 1;
    for (int p = 0; p < npatch; ++p) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(dtok + (row0 + p) * D + col));
      a0 += f.x; a1 += f.y;
    }
    if (dchan) { atomicAdd(dchan + (long)chan_idx[g] * D + col, a0); atomicAdd(dchan + (long)chan_idx[g] * D + col + 1, a1); }
    atomicAdd(dbias + col, a0); atomicAdd(dbias + col + 1, a1);
  } else {
    for (int b = 0; b < B; ++b) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(dtok + (long)cu[b] * D + col));
      a0 += f.x; a1 += f.y;
    }
    atomicAdd(dcls_tok + col, a0); atomicAdd(dcls_tok + col + 1, a1);
    atomicAdd(dpos0 + col, a0); atomicAdd(dpos0 + col + 1, a1);
  }
}

}  // namespace cb

using namespace cb;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int cb_tokenize_fwd(const float* x, int G, int H, int W, int patch, const int* cu_seqlens, const int* chan_img, int B,
                               const void* w_pe, const float* b_pe, const float* pos_patch, const float* pos0,
                               const float* cls_tok, const float* chan_tok, void* patches_ws, void* tokens, int T, int D, void* stream) {
  CB_CHECK(G > 0 && B > 0 && patch % 8 == 0 && H >= patch && W >= patch && W % 4 == 0, "tokenize_fwd: bad shape G=%d B=%d H=%d W=%d patch=%d", G, B, H, W, patch);
  const int npatch = (H / patch) * (W / patch), PP = patch * patch;
  CB_CHECK(T == B + G * npatch, "tokenize_fwd: T=%d != B + G*N = %d (index bookkeeping)", T, B + G * npatch);
  CB_CHECK(D % 8 == 0, "tokenize_fwd: D=%d must be a multiple of 8", D);
  const long total = (long)G * (H / patch) * patch * ((W / patch) * patch / 8);
  long blocks = (total + 255) / 256;
  if (blocks > (long)num_sms() * 32) blocks = (long)num_sms() * 32;
  im2col_packed_kernel<<<(int)blocks, 256, 0, STREAM>>>(x, reinterpret_cast<__nv_bfloat16*>(patches_ws), chan_img, cu_seqlens, G, B, H, W, patch);
  CB_CUDA(cudaGetLastError());
  GemmArgs g{};
  g.M = T; g.N = D; g.K = PP; g.k_splits = 1; g.C = tokens; g.ldc = D; g.bias = b_pe; g.flags = CB_EPI_TOKENIZE | CB_EPI_OUT_F32; g.alpha = 1.f;
  g.cu = cu_seqlens; g.nseq = B; g.pos = pos_patch; g.chan_tok = chan_tok; g.cls_tok = cls_tok; g.pos0 = pos0; g.npatch = npatch;
  return gemm_run(patches_ws, PP, 0, w_pe, PP, 0, g, STREAM);
}

extern "C" int cb_tokenize_bwd(const void* dtokens, const void* patches_ws, const int* cu_seqlens, const int* chan_img,
                               const int* chan_idx, int G, int B, int npatch, int patch_elems, int T, int D, float* dw_pe,
                               float* db_pe, float* dpos_patch, float* dpos0, float* dcls_tok, float* dchan_tok, int k_splits,
                               void* stream) {
  CB_CHECK(T == B + G * npatch && D % 32 == 0 && D <= 2048, "tokenize_bwd: bad shape T=%d B=%d G=%d N=%d D=%d", T, B, G, npatch, D);
  tokenizer_embed_bwd_kernel<<<npatch + G + 1, D / 2, 0, STREAM>>>(reinterpret_cast<const __nv_bfloat16*>(dtokens), chan_img, chan_idx,
                                                                     cu_seqlens, G, B, npatch, D, dpos_patch, dchan_tok, db_pe, dcls_tok, dpos0);
  CB_CUDA(cudaGetLastError());
  // dW_pe[D, P²] += dTok^T[D, T] · patches[T, P²]   (CLS rows of `patches` are zero)
  GemmArgs g{};
  g.M = D; g.N = patch_elems; g.K = T; g.k_splits = k_splits; g.C = dw_pe; g.ldc = patch_elems; g.flags = CB_EPI_ATOMIC; g.alpha = 1.f;
  return gemm_run(dtokens, D, 1, patches_ws, patch_elems, 1, g, STREAM);
}
