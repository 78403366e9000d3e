/*
 * chadavit_b200 — C ABI of the B200 (sm_100a) hot path for ChAda-ViT + DINO.
 *
 * Drop-in boundary.  The reference (nicoboou/chadavit) is pure Python/PyTorch; the "FFI" a
 * maintainer binds is therefore a ctypes/cffi stub inside the reference's own modules:
 *     src/backbones/vit/chada_vit.py   ChAdaViT.forward / channel_aware_tokenization / TransformerEncoderLayer
 *     src/methods/dino.py:32-111       DINOHead
 *     src/losses/dino.py:69-118        DINOLoss.forward / update_center
 *     src/utils/momentum.py:63-74      MomentumUpdater.update
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; cb_last_error() gives the message (thread local).
 *     Nothing throws, nothing calls exit().  There is NO CPU fallback: all pointers are device pointers.
 *   - every buffer is owned and allocated by the caller (torch); the library allocates nothing persistent.
 *   - `stream` is a cudaStream_t passed as void*; all entry points are asynchronous on that stream.
 *   - bf16 tensors are passed as `const void*` (raw __nv_bfloat16 storage), fp32 as `float*`, indices as int32.
 *   - packed varlen token layout (replaces the reference's pad-to-10-channels + key-padding mask,
 *     chada_vit.py:226-268):  tokens (T, D) row-major, cu_seqlens int32 (B+1), sequence b owns rows
 *     [cu[b], cu[b+1]) = 1 + C_b*N rows: row cu[b] is CLS, row cu[b]+1+c*N+p is patch p of channel c.
 */
#ifndef CHADAVIT_B200_H
#define CHADAVIT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CHADAVIT_B200_VERSION 1

/* ---- epilogue flags of cb_gemm_bf16 ---- */
#define CB_EPI_RELU 1       /* C = max(C, 0)                                  (linear1 + F.relu, chada_vit.py:115)   */
#define CB_EPI_RESIDUAL 2   /* reserved (bf16 residual; rejected: the residual stream is fp32, see CB_EPI_RESIDUAL_F32)  */
#define CB_EPI_RELU_MASK 4  /* C = aux > 0 ? C : 0                            (backward of F.relu)                    */
#define CB_EPI_OUT_F32 8    /* store fp32 instead of bf16                                                             */
#define CB_EPI_ATOMIC 16    /* fp32 atomic accumulate into C (split-K weight gradients)                               */
#define CB_EPI_TOKENIZE 32  /* internal: tokenizer scatter epilogue (use cb_tokenize_fwd)                             */
#define CB_EPI_RESIDUAL_F32 64 /* C += aux, aux fp32 [M,N], needs CB_EPI_OUT_F32      (x + attn / x + ff, chada_vit.py:99-100) */
#define CB_EPI_MASK_BITS 128 /* with CB_EPI_RELU_MASK: aux is a BIT mask, uint32 [N/32, ld_aux] (ld_aux >= M, in words): bit j of
                                aux[n/32][m] set <=> keep C[m, 32*(n/32) + j] — the layout cb_ffn_fwd writes (1 bit instead of 16
                                per hidden unit, and consecutive rows are consecutive words: coalesced)                          */

const char* cb_last_error(void);
int cb_version(void);
int cb_num_sms(void);
/* synchronise `stream` and report any asynchronous kernel failure */
int cb_sync_check(void* stream);
/*
 * HOST-side helper (no device work): the work list of cb_attn_varlen_fwd / _bwd for one ragged batch, derived from the host
 * copy of cu_seqlens (B+1 ints, i.e. from list_num_channels, src/data/channels_strategies.py:31-85 — no device sync).
 * Items are int32 quadruples {first row of the tile, seq_start, seq_end, head}, `tile` rows per item, longest sequences
 * first.  mode 0: the plain list.  mode 1 (forward cost model) / 2 (backward cost model): the list assigned
 * longest-processing-time-first to n_ctas persistent CTAs, written round-major (CTA c walks slots c, c + n_ctas, ...) and
 * padded with all-zero slots.  out: host int32 [cap, 4]; *n_out = slots written (or needed, when the call fails with cap
 * too small).
 */
int cb_attn_schedule(const int* cu_host, int B, int nheads, int tile, int mode, int n_ctas, int* out, int cap, int* n_out);

/*
 * C[M,N] (+)= alpha * op(A)[M,K] · op(B)[N,K]^T (+ bias[N]) with the epilogue in `flags`; bf16 operands, fp32
 * accumulation on the tcgen05 tensor cores.  a_mn = 0: A is [M,K] row-major (lda);  a_mn = 1: A is stored [K,M]
 * row-major (lda).  Same for B / b_mn with [N,K] vs [K,N].  Replaces every nn.Linear / F.linear on the path
 * (MHA in/out projection chada_vit.py:106, linear1/linear2 :115, DINOHead.mlp / last_layer dino.py:65-81) and
 * their autograd backward products.  N, lda, ldb multiples of 8; MN-major dims multiples of 32.
 * colsum (optional, accumulated): with CB_EPI_RELU_MASK fp32 [N] += column sums of the stored C; with CB_EPI_ATOMIC (the
 * weight-gradient product dW = dY^T X: a_mn = b_mn = 1, N a multiple of 192 but not of 256, M and N multiples of 64)
 * fp32 [M] += sum over K of op(A)[m,k], i.e. the bias gradient that belongs to dW (torch.autograd of F.linear:
 * grad_bias = grad_output.sum(0)), formed by the same tensor-core instructions as the tile (one extra block of ones behind
 * the B operand).
 */
int cb_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                 int K, const float* bias, const void* aux, int ld_aux, int flags, float alpha, int k_splits,
                 float* colsum, void* stream);

/*
 * Linear + residual + LayerNorm in one kernel — the encoder's  y = norm1(x + self_attn_out W_o^T + b_o)  (chada_vit.py:99 with
 * nn.MultiheadAttention's out_proj, :105-111):
 *   z = A W^T + bias + resid        A bf16 [M, K] (lda), W bf16 [N, K] (ldw), bias fp32 [N], resid fp32 [M, N] (ld_res) or NULL
 *   y = LayerNorm(z; ln_gamma, ln_beta, ln_eps)   two-pass fp32 statistics, as cb_layernorm_fwd
 * Outputs: y_bf16 [M, N] (required), y_f32 [M, N] or NULL, mean / rstd fp32 [M] or NULL (saved for cb_layernorm_bwd), z fp32 [M, N]
 * or NULL (kept only when the backward needs it).  N must be 192 (the row tile of the tensor-core kernel holds whole rows),
 * K <= 192, M >= 512; other shapes: cb_gemm_bf16 (CB_EPI_RESIDUAL_F32 | CB_EPI_OUT_F32) followed by cb_layernorm_fwd.
 */
int cb_gemm_ln_fwd(const void* A, int lda, const void* W, int ldw, const float* bias, const float* resid, int ld_res,
                   const float* ln_gamma, const float* ln_beta, float ln_eps, float* z, void* y_bf16, float* y_f32, float* mean,
                   float* rstd, int M, int N, int K, void* stream);

/*
 * TokenLearner + channel_aware_tokenization (chada_vit.py:118-134, 219-270) on the packed layout.
 *   x           fp32 (G,1,H,W): channel images in one_channel_collate_fn order (channels_strategies.py:31-85)
 *   cu_seqlens  int32 [B+1];  chan_img int32 [G] = image index b of channel image g
 *   w_pe bf16 [D, patch*patch] (token_learner.proj.weight), b_pe fp32 [D]
 *   pos_patch fp32 [N, D] (pos_embed[0,0,1:], bicubic-resized by the caller when N != 196, chada_vit.py:201-217)
 *   pos0 fp32 [D] = pos_embed[0,0,0], cls_tok fp32 [D] = cls_token (CLS row = cls_tok + pos0, chada_vit.py:259-262)
 *   chan_tok fp32 [max_ch, D] or NULL (chada_vit.py:248)
 *   patches_ws  bf16 [T, patch*patch] workspace (kept for cb_tokenize_bwd);  tokens fp32 [T, D] output
 */
int cb_tokenize_fwd(const float* x, int G, int H, int W, int patch, const int* cu_seqlens, const int* chan_img, int B,
                    const void* w_pe, const float* b_pe, const float* pos_patch, const float* pos0,
                    const float* cls_tok, const float* chan_tok, void* patches_ws, void* tokens, int T, int D, void* stream);
/* gradients of the tokenizer parameters from dtokens bf16 [T,D]; all outputs fp32 and ACCUMULATED (+=):
 * dw_pe [D,patch_elems], db_pe [D], dpos_patch [N,D], dpos0 [D], dcls_tok [D], dchan_tok [max_ch,D] or NULL */
int cb_tokenize_bwd(const void* dtokens, const void* patches_ws, const int* cu_seqlens, const int* chan_img,
                    const int* chan_idx, int G, int B, int npatch, int patch_elems, int T, int D, float* dw_pe,
                    float* db_pe, float* dpos_patch, float* dpos0, float* dcls_tok, float* dchan_tok, int k_splits,
                    void* stream);
/* plain unfold + cast (patch rows in channel-image order), used by tests */
int cb_im2col_bf16(const float* x, void* patches, int G, int H, int W, int patch, void* stream);

/*
 * nn.LayerNorm over the last dim (norm1 / norm2 / final norm, chada_vit.py:96,99,100,281).  x fp32 [*, D] (the
 * residual stream stays fp32; only GEMM operands are rounded to bf16);
 * row r of the output normalises input row in_idx[r] (or r when in_idx is NULL: the CLS gather of chada_vit.py:289
 * is in_idx = cu_seqlens).  Either output may be NULL: y bf16 [rows,D], y_f32 fp32 [rows,D].  mean/rstd fp32 [rows].
 */
int cb_layernorm_fwd(const float* x, const int* in_idx, const float* gamma, const float* beta, void* y, float* y_f32,
                     float* mean, float* rstd, int rows, int D, float eps, void* stream);
/*
 * Two LayerNorms back to back on rows kept in registers: y1 = LN_a(x) (fp32, may be NULL) and y2 = LN_b(y1) (bf16) —
 * norm2 of block i (chada_vit.py:100) followed by norm1 of block i+1 (chada_vit.py:96), without writing y1 and reading it
 * straight back.  D in {64, 128, 192, 256}.  mean/rstd of either stage fp32 [rows] or NULL.
 */
int cb_layernorm2_fwd(const float* x, const float* gamma_a, const float* beta_a, float eps_a, const float* gamma_b,
                      const float* beta_b, float eps_b, float* y1_f32, void* y2_bf16, float* mean_a, float* rstd_a, float* mean_b,
                      float* rstd_b, int rows, int D, void* stream);
/* dx[idx[r]] = dLN(dy[r]) (+ dres[idx[r]]), written as fp32 (dx_f32) and/or bf16 (dx_bf16); dy, x, dres fp32;
 * dgamma/dbeta/dcolsum (= column sum of dLN, i.e. the bias gradient of the preceding linear) fp32 [D], ACCUMULATED. */
int cb_layernorm_bwd(const float* dy, const float* x, const int* idx, const float* gamma, const float* mean,
                     const float* rstd, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                     float* dcolsum, int rows, int D, void* stream);
/* out[n] += sum_t x[t,n]  (bias gradients);  x bf16 [T,N] with row stride ld */
int cb_colsum_bf16(const void* x, int ld, float* out, int T, int N, void* stream);
int cb_cast_f32_bf16(const float* in, void* out, long n, void* stream);
int cb_gather_rows_f32(const void* x, const int* idx, float* out, int rows, int D, void* stream);
/* tiny fp32 C[M,N] (+)= op(A) B[K,N] for the bicubic pos-embed resize map (chada_vit.py:201-217); trans_a: A stored [K,M] */
int cb_small_matmul_f32(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int accumulate,
                        void* stream);

/*
 * Packed varlen multi-head self-attention forward (nn.MultiheadAttention inside _sa_block, chada_vit.py:105-111).
 *   qkv bf16 [T, 3D] (= in_proj output: q | k | v, head h in columns h*d..(h+1)*d of each third)
 *   work int32 [n_work, 4] = {first query row (global), seq_start, seq_end, head}: one entry per q_tile query rows
 *        (q_tile must be 256: two 128-row tiles per item, one softmax warpgroup each),
 *        built on the host from list_num_channels (no device sync).  The persistent kernels give CTA c the entries c, c + G,
 *        c + 2G, ... (G = min(n_work, SM count)); both kernels skip EMPTY entries (seq_end <=
 *        seq_start), so the host may pad the list to balance the CTAs (PackedLayout.attn_schedule: LPT assignment)
 *   out bf16 [T, D];  lse fp32 [H, T] (log-sum-exp of the scaled scores, natural log) or NULL
 */
int cb_attn_varlen_fwd(const void* qkv, const int* work, int n_work, int q_tile, void* out, float* lse, int T, int D, int H,
                       float softmax_scale, void* stream);

/*
 * Backward of cb_attn_varlen_fwd.  dout bf16 [T,D] (gradient of `out`), qkv/out/lse as saved by the forward, `work`
 * the same tile list.  Workspaces (caller-allocated): delta_ws fp32 [H,T], dq_acc_ws fp32 [T,D].  dqkv bf16 [T,3D].
 */
int cb_attn_varlen_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const int* work, int n_work,
                       float* delta_ws, float* dq_acc_ws, void* dqkv, int T, int D, int H, float softmax_scale, void* stream);

/*
 * Fused feed-forward block + residual of the encoder layer (chada_vit.py:113-116 _ff_block and the add of :100):
 *   z2 = resid + relu(y W1^T + b1) W2^T + b2
 * y bf16 [T, D] (norm1 output), w1 bf16 [F, D] (linear1.weight), w2 bf16 [D, F] (linear2.weight), b1 fp32 [F], b2 fp32 [D],
 * resid / z2 fp32 [T, D].  hid bf16 [T, F] receives relu(y W1^T + b1) when not NULL (saved for the backward pass); with
 * hid == NULL the hidden activations never leave the SM.  mask_bits (optional) receives the ReLU mask as bits, uint32
 * [F/32, ld_bits] (ld_bits >= T, a multiple of 32): bit j of mask_bits[w][t] <=> hid[t, 32 w + j] > 0 — what the backward
 * product d(hidden) = (dz2 W2) o (hidden > 0) reads through CB_EPI_MASK_BITS instead of the 16x larger hid.
 * D must be 192 (tensor-memory budget), F a multiple of 64; other shapes use two cb_gemm_bf16 calls.
 * kernel: 0 = the library's choice (cluster-of-two kernel without the hidden store, pair-of-tiles kernel with it); 1 / 3 force
 * the pair-of-tiles / the cluster kernel (A/B measurements and tests; 3 needs F % 128 == 0).
 */
int cb_ffn_fwd(const void* y, const void* w1, const float* b1, const void* w2, const float* b2, const float* resid, float* z2,
               void* hid, unsigned int* mask_bits, int ld_bits, int T, int D, int F, int kernel, void* stream);

/*
 * Backward of the same block through the hidden layer (torch.autograd of linear2(relu(linear1(y))) + the residual branch,
 * chada_vit.py:113-116 / :100), one kernel instead of two cb_gemm_bf16 calls:
 *   dh = (dz2 W2) o (hidden > 0)        bf16 [T, F], stored (the weight gradient dW1 = dh^T y reads it)
 *   dy = dh W1 + dz2                    fp32 [T, D]
 * dz2_bf16 bf16 [T, D] and dz2 fp32 [T, D] are the two forms of d(z2) that cb_layernorm_bwd returns; w2 bf16 [D, F]
 * (linear2.weight), w1 bf16 [F, D] (linear1.weight), both read in place (MN-major operands, no transposed copies);
 * mask_bits uint32 [F/32, ld_bits >= T] as written by cb_ffn_fwd.  The [T, F] d(hidden) is written once and not read back by
 * this kernel.  D must be 192, F a multiple of 64 (<= 2048); other shapes use cb_gemm_bf16 with CB_EPI_RELU_MASK and
 * CB_EPI_RESIDUAL_F32.
 */
int cb_ffn_bwd(const void* dz2_bf16, const void* w2, const void* w1, const unsigned int* mask_bits, int ld_bits, const float* dz2,
               float* dy, void* dh, int T, int D, int F, void* stream);

/* ---------------- DINOHead pieces (src/methods/dino.py:61-111); the Linear layers themselves are cb_gemm_bf16 ---------------- */
/* nn.GELU (exact erf): out bf16 = gelu(pre fp32);  backward: dpre bf16 = dact fp32 * gelu'(pre) */
int cb_gelu_fwd(const float* pre, void* out, long n, void* stream);
int cb_gelu_bwd(const float* dact, const float* pre, void* dpre, long n, void* stream);
/* nn.BatchNorm1d + nn.GELU of the projector when the head is built with use_bn=True (the class default, src/methods/dino.py:40,
 * 66-73): pre fp32 [R,C] (Linear output) -> bn_out fp32 [R,C] (normalised + affine, saved for the backward; may be NULL) and
 * act bf16 [R,C] = gelu(bn_out).  training != 0: batch statistics, running_mean / running_var updated in place with `momentum`
 * (unbiased variance, as torch), save_mean / save_invstd [C] written; training == 0: running statistics.  Backward: dpre bf16
 * from dact fp32 = d loss / d act; dgamma / dbeta accumulate (+=).  Per-process statistics (no SyncBatchNorm). */
int cb_bn_gelu_fwd(const float* pre, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                   float eps, int training, float* bn_out, void* act, float* save_mean, float* save_invstd, int R, int C, void* stream);
int cb_bn_gelu_bwd(const float* dact, const float* bn_out, const float* pre, const float* gamma, const float* save_mean,
                   const float* save_invstd, int training, void* dpre, float* dgamma, float* dbeta, int R, int C, void* stream);
/* F.normalize(x, dim=-1, eps): out bf16 [rows,C], inv fp32 [rows] = 1/max(||x||,eps);  backward -> dx bf16 */
int cb_l2norm_fwd(const float* x, void* out, float* inv, int rows, int C, float eps, void* stream);
int cb_l2norm_bwd(const float* dy, const float* x, const float* inv, void* dx, int rows, int C, void* stream);
/* nn.utils.weight_norm(Linear(C,K,bias=False)): w bf16 [K,C] = g[k] * v[k,:]/||v[k,:]||;  backward ACCUMULATES dv (and dg
 * unless NULL = weight_g frozen by norm_last_layer, dino.py:83-84) from dw fp32 [K,C] */
int cb_weightnorm_fwd(const float* v, const float* g, void* w, float* inv_norm, int K, int C, void* stream);
int cb_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg, int K,
                      int C, void* stream);

/*
 * DINOLoss.forward (src/losses/dino.py:81-99) fused with its gradient, one pass per batch row:
 *   student fp32 [V*B,K] (view v = rows v*B..v*B+B), teacher fp32 [2*B,K], center fp32 [K] (the OLD center, Q13)
 *   loss fp32 [1] (overwritten);  d(loss)/d(student) as fp32 [V*B,K] and/or bf16 [V*B,K] (either may be NULL)
 */
int cb_dino_loss_fwd_bwd(const float* student, const float* teacher, const float* center, float* loss, float* dstudent_f32,
                         void* dstudent_bf16, int B, int K, int V, float student_temp, float teacher_temp, void* stream);
/* update_center (src/losses/dino.py:103-118): out[k] = sum_r x[r,k]; then (after the caller's all-reduce)
 * center = center*momentum + batch_sum*scale*(1-momentum), scale = 1/(world_size * rows) */
int cb_colsum_f32(const float* x, float* out, int R, int K, void* stream);
int cb_dino_center_ema(float* center, const float* batch_sum, float scale, float momentum, int K, void* stream);

/* MomentumUpdater.update over a flat parameter arena (src/utils/momentum.py:73-74): mp = tau*mp + (1-tau)*op, one launch;
 * optionally refreshes the teacher's bf16 shadow in the same pass.  12 B/param of algorithmic HBM traffic. */
int cb_ema_update(float* momentum, const float* online, void* momentum_bf16, float tau, long n, void* stream);
/* torch.optim.AdamW step over a flat arena, optionally fused with the teacher EMA and the bf16 shadow
 * refresh of both networks.  flags[i]: bit0 = weight decay applies, bit1 = frozen (e.g. head.last_layer while
 * current_epoch < freeze_last_layer, dino.py:374-376), bit4 = the element belongs to a parameter whose own step count is
 * step_late (torch.optim.AdamW counts steps per parameter and skips parameters without a gradient, so head.last_layer starts
 * at 1 when it is unfrozen); NULL = decay everything.  dev_hyper (optional, device fp32[6] = {lr, 1-beta1^step,
 * sqrt(1-beta2^step), tau, 1-beta1^step_late, sqrt(1-beta2^step_late)}) overrides the per-step scalars so a captured CUDA
 * graph can be replayed. */
int cb_adamw_step(float* p, const float* g, float* m, float* v, const unsigned char* flags, void* p_bf16, float* teacher,
                  void* teacher_bf16, long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  int step_late, float grad_scale, float tau, const float* dev_hyper, void* stream);

/* Attention probabilities of one block for ChAdaViT.get_last_selfattention (src/backbones/vit/chada_vit.py:313-320, read by
 * main_attn.py:200-207): out[b, h, i, j] = softmax_j(q_i . k_j * scale) over the tokens of packed sequence b; qkv is the
 * packed bf16 [T, 3*H*d] projection (rows q | k | v, head h = columns h*d..), out is fp32 [nseq, H, max_seqlen, max_seqlen],
 * entries beyond a sequence's length are zero. */
int cb_attn_probs(const void* qkv, const int* cu_seqlens, int nseq, int num_heads, int head_dim, int max_seqlen, float scale,
                  float* out, void* stream);

/* Self-attention of the LAST encoder block when the backbone returns only the CLS embedding (`return x[:, 0]`,
 * src/backbones/vit/chada_vit.py:289; the SDPA of nn.MultiheadAttention, :105-111).  Everything behind that attention is row-wise,
 * so only the CLS query of every sequence is used: out_cls[b, h*d..] = softmax_j(q_cls(b) . k_j * scale) V over the tokens of packed
 * sequence b (bf16 [nseq, H*d]); lse_cls fp32 [nseq, H] (log2 domain, consumed by cb_attn_cls_bwd).  qkv as in cb_attn_varlen_fwd. */
int cb_attn_cls_fwd(const void* qkv, const int* cu_seqlens, int nseq, int num_heads, int head_dim, float scale, void* out_cls,
                    float* lse_cls, void* stream);

/* Backward of cb_attn_cls_fwd: dout_cls bf16 [nseq, H*d] = d(attention output) of the CLS rows (all other rows are zero by
 * construction).  Writes EVERY element of dqkv (bf16 [T, 3*H*d]): dK / dV of all tokens, dQ of the CLS rows, zeros elsewhere. */
int cb_attn_cls_bwd(const void* dout_cls, const void* qkv, const void* out_cls, const float* lse_cls, const int* cu_seqlens, int nseq,
                    int num_heads, int head_dim, float scale, void* dqkv, void* stream);

/* Weighted k-NN evaluation (src/utils/knn.py:96-177): operand preparation for an fp32-accurate similarity matrix on the bf16
 * tensor cores.  Row r of x (fp32 [rows, D]) is optionally L2-normalised (F.normalize, knn.py:114-116) and written as bf16
 * [hi | hi | lo] (role_b = 0) or [hi | lo | hi] (role_b = 1) of length 3*D, so that ONE cb_gemm_bf16 over K = 3*D yields
 * a.b up to a 2^-16 relative term.  sqnorm (optional) receives ||row||^2.  cb_inv_euclid turns such dot products in place into
 * 1 / (cdist + eps) (knn.py:141). */
int cb_split_bf16x3(const float* x, void* out, float* sqnorm, int rows, int D, int role_b, int normalize, void* stream);
int cb_inv_euclid(float* dots, const float* sqnorm_a, const float* sqnorm_b, int M, int N, int ld, float eps, void* stream);

/* Per-parameter L2 norms over a flat arena whose parameters start at 64-element aligned offsets.
 * seg_start_block[nseg+1]: first 64-element block of every parameter (int32, ascending; padding belongs to the parameter
 * in front of it and holds zeros).  partial: workspace of (n/64)*2 floats.  Output norms[3*s + {0,1,2}] =
 * { ||p_s||, ||g_s * grad_scale * coef_s||, coef_s } with coef_s = min(1, clip / (||g_s * grad_scale|| + 1e-6)) where
 * seg_clip[s] != 0 and clip > 0 — DINO.dino_clip_gradients (src/methods/dino.py:249-261), else 1.  Two launches, no atomics:
 * the result is bit-reproducible, so data-parallel replicas stay identical. */
int cb_param_norms(const float* p, const float* g, const int* seg_start_block, const unsigned char* seg_clip, float* partial,
                   float* norms, long n, int nseg, float grad_scale, float clip, void* stream);
/* g[i] *= coef of the parameter that owns element i (seg_of_block: int32 per 64-element block; norms from cb_param_norms):
 * the in-place form of dino_clip_gradients, for optimizers that do not take the coefficient themselves (cb_adamw_step). */
int cb_scale_grads(float* g, const int* seg_of_block, const float* norms, long n, void* stream);
/* LARS.step over a flat arena (src/utils/lars.py:113-167), fused with the teacher EMA and the bf16 shadow refresh like
 * cb_adamw_step.  flags[i]: bit0 = the group's weight decay applies (0 otherwise: base.py:426-427), bit1 = frozen / no
 * gradient, bit2 = layer-wise adaptation applies (p.ndim != 1 or not exclude_bias_n_norm), bit3 = first update of this
 * parameter (momentum buffer := d_p).  The gradient used is g * grad_scale * coef_s.  dev_hyper (optional, device fp32[>=4];
 * slots 0 = lr and 3 = tau are read) serves CUDA-graph replay. */
int cb_lars_step(float* p, const float* g, float* buf, const unsigned char* flags, const int* seg_of_block, const float* norms,
                 void* p_bf16, float* teacher, void* teacher_bf16, long n, float lr, float momentum, float dampening,
                 int nesterov, float weight_decay, float eta, float eps, int clip_lr, float grad_scale, float tau,
                 const float* dev_hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHADAVIT_B200_H */
