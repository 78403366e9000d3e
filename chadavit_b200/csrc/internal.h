// Declarations shared between translation units of libchadavit_b200 (not part of the public C ABI).
#pragma once
#include "common.cuh"

namespace cb {

struct GemmArgs {
  int M, N, K;
  int k_splits;
  void* C;
  int ldc;
  const float* bias;
  const __nv_bfloat16* aux;
  int ld_aux;
  int flags;
  float alpha;
  int direct;             // set by gemm_run: C / aux are 32-byte aligned with 32-byte row pitches -> row-direct epilogue allowed
  float* colsum;          // optional fp32, accumulated.  EM_BF16_MASK: [N] column sums of the stored C (bias gradient of linear1);
                          // EM_ATOMIC (dW = dY^T X): [M] sums of op(A) over K, formed on the tensor pipe (bias gradient next to dW)
  // LayerNorm fused into the fp32 row epilogue (cb_gemm_ln_fwd; N = 192 only): y = LN(C) with C = alpha A B^T + bias + aux.
  // C itself is stored only when g.C != nullptr (the backward needs it; the no-grad passes do not).
  const float* ln_gamma;  // [N]; non-null selects the fused epilogue
  const float* ln_beta;   // [N]
  float ln_eps;
  __nv_bfloat16* ln_y16;  // [M, N] bf16 (next GEMM operand)
  float* ln_y32;          // [M, N] fp32 or null (next residual)
  float* ln_mean;         // [M] or null
  float* ln_rstd;         // [M] or null
  // tokenizer epilogue (CB_EPI_TOKENIZE): A rows are already in packed token order (CLS rows hold zeros);
  // row t of sequence b (cu[b] <= t < cu[b+1]): off = t - cu[b]; off == 0 -> CLS row = cls_row, else
  // c = (off-1)/npatch, p = (off-1)%npatch: acc + bias + pos[p] + chan_tok[c]        (chada_vit.py:245-265)
  const int* cu;          // [nseq+1]
  int nseq;
  const float* pos;       // [npatch, N] fp32 (already interpolated if needed)
  const float* chan_tok;  // [max_ch, N] fp32 or null
  const float* cls_tok;   // [N] fp32 cls_token
  const float* pos0;      // [N] fp32 pos_embed[0,0,0]  (CLS row = cls_tok + pos0)
  int npatch;
};

// C = op(A) op(B)^T with the epilogue in g.flags; see gemm.cu
int gemm_run(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, GemmArgs g, cudaStream_t stream);

// fused feed-forward, cluster-of-two kernel (ffn_fwd3.cu): 128-unit hidden chunks, y tile in shared memory; same contract as
// cb_ffn_fwd with D = 192; needs F % 128 == 0
int ffn_fwd3_run(const void* y, const void* w1, const float* b1, const void* w2, const float* b2, const float* resid, float* z2, void* hid,
                 unsigned int* mask_bits, int ld_bits, int T, int F, cudaStream_t stream);

}  // namespace cb
