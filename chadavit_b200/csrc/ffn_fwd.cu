// Fused feed-forward block of the encoder layer (chada_vit.py:113-116 _ff_block, with the residual of :100):
//     z2 = resid + relu(y W1^T + b1) W2^T + b2            y bf16 [T, D], W1 [F, D], W2 [D, F], resid / z2 fp32 [T, D]
// in ONE kernel: the [T, F] hidden activations never go through HBM between the two GEMMs (separately, linear1 writes
// 2F bytes per token — it sits on the HBM write roofline — and linear2 reads them back).  Optionally the hidden
// activations are ALSO stored (bf16) for the backward pass of the student; the teacher / local-crop passes skip that.
//
// Structure = the attention forward with the softmax replaced by bias + ReLU.  A work item is a PAIR of 128-row tiles
// (A, B); the hidden dimension is walked in chunks of 64 units:
//     H_t(c)  = y_t · W1[c]^T          SS MMA, M = 128, N = 64, K = D      -> TMEM (64 fp32 columns)
//     P_t(c)  = bf16(relu(H_t(c) + b1[c]))   by 128 threads (1 thread = 1 row), written over H_t in TMEM
//     Z_t    += P_t(c) · W2[:, c]^T    TS MMA (A from TMEM), M = 128, N = D, K = 64     -> TMEM (D fp32 columns)
// TMEM (D = 192): H_A 64 | H_B 64 | Z_A 192 | Z_B 192 = 512 columns.  Each weight chunk (W1[c]: 64 x D, W2[:, c]: D x 64,
// 48 KB) is fetched once per PAIR of row tiles — with one tile per CTA the weight stream alone (96 KB per 1536 MMA clocks and
// SM) would exceed what L2 delivers.  Warps 0-3 / 4-7: bias + ReLU (+ hidden store) and final epilogue of tile A / B;
// warp 8: MMA issuer (convergent warp, elected lane) serving A, B, A, B ...; warp 9: TMA.  While the epilogue warps turn
// H_A(c) into P_A(c), the tensor pipe runs Z_B += P_B(c) W2 and H_B(c+1), and vice versa.
//
// BWD = true is the backward of the same block through the hidden layer (autograd of chada_vit.py:113-116), the same pipeline
// with the roles of the weights exchanged and the ReLU replaced by its mask:
//     DH_t(c) = dz2_t · W2[:, c]           SS MMA, B = W2 [D, F] read MN-major ([K, N] row-major), N = 64, K = D
//     P_t(c)  = bf16(DH_t(c)) where hidden > 0   (1 bit per unit, written by the forward), stored as d(hidden) for dW1 = dh^T y
//     DY_t   += P_t(c) · W1[c, :]          TS MMA, B = W1 [F, D] read MN-major, N = D, K = 64
//     dy      = DY + dz2 (fp32: the residual branch of chada_vit.py:100)
// i.e. cb_gemm_bf16(dz2, W2, relu-mask) + cb_gemm_bf16(dh, W1, +residual) without reading the [T, F] d(hidden) back.
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"
#include <stdlib.h>

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

constexpr int FF_D = 192;          // model width handled by this kernel (TMEM: 2 x 64 + 2 x D <= 512)
constexpr int FF_C = 64;           // hidden units per chunk
constexpr int FF_KB = FF_D / 64;   // 64-wide k-blocks of D
constexpr int FF_Y_BYTES = 128 * FF_D * 2;          // one row tile of y: FF_KB blocks of [128 x 64] (128B swizzle)
constexpr int FF_W1_BYTES = FF_C * FF_D * 2;        // W1 chunk: FF_KB blocks of [64 x 64]
constexpr int FF_W2_BYTES = FF_D * FF_C * 2;        // W2 chunk: [D x 64]
constexpr int FF_S1 = 3, FF_S2 = 2;                 // ring depths of the W1 / W2 chunk slots (separate rings: a W1 slot is free as soon
                                                    // as H(c) has retired, long before the W2 slot of the same chunk)
constexpr int FF_MAX_F = 2048;                      // b1 is staged in shared memory
constexpr int FF_SMEM_BYTES = 2 * FF_Y_BYTES + FF_S1 * FF_W1_BYTES + FF_S2 * FF_W2_BYTES + FF_MAX_F * 4 + 1024 /*align*/ + 512 /*barriers*/;
constexpr int FF_COL_Z = 128;      // H_t at 64 t, Z_t at 128 + 192 t

struct FfnArgs {
  const float* b1;     // [F]                                                                 (BWD: unused)
  const float* b2;     // [D]                                                                 (BWD: unused)
  const float* resid;  // [T, D] fp32 (norm1 output, the residual of chada_vit.py:100)        (BWD: dz2 fp32)
  float* z2;           // [T, D] fp32                                                         (BWD: dy)
  __nv_bfloat16* hid;  // [T, F] bf16 or null                                                 (BWD: d(hidden), always stored)
  uint32_t* mask_bits; // [F/32, ld_bits] ReLU mask as bits, or null                          (BWD: read)
  int ld_bits;
  int T, F;
};

__device__ __forceinline__ uint32_t ldg_u32_pinned(const uint32_t* p) {   // stays where it is written (see ldg_f32_pinned)
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// WMN: the two weight operands are read MN-major in place (the backward reads the stored weights: no transposed copies);
// false: K-major (forward).  Measured: the backward on pre-transposed K-major copies runs in the same time to the microsecond
// (156.7 vs 156.5 us per 68 k tokens, profiles/r02_microbench_ffn_bwd.txt), so only the in-place form is kept.
template <bool BWD, bool WMN = BWD>
__global__ void __launch_bounds__(320, 1)
ffn_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
           const FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sY = smem;                                   // [2] row tiles
  uint8_t* sW1 = sY + 2 * FF_Y_BYTES;                   // [S1] W1 chunks
  uint8_t* sW2 = sW1 + FF_S1 * FF_W1_BYTES;             // [S2] W2 chunks
  float* sB1 = reinterpret_cast<float*>(sW2 + FF_S2 * FF_W2_BYTES);   // [F]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + FF_MAX_F);
  uint64_t* y_full = bars + 0;              // [2]
  uint64_t* y_empty = bars + 2;             // [2] every H_t MMA of the item has retired
  uint64_t* w1_full = bars + 4;             // [S1]
  uint64_t* w1_empty = w1_full + FF_S1;     // [S1] 2 arrivals (both MMA warps)
  uint64_t* w2_full = w1_empty + FF_S1;     // [S2]
  uint64_t* w2_empty = w2_full + FF_S2;     // [S2] 2 arrivals
  uint64_t* h_full = w2_empty + FF_S2;      // [2] H_t(c) complete
  uint64_t* p_full = h_full + 2;            // [2] 256 arrivals: P_t(c) written
  uint64_t* z_full = p_full + 2;            // [2]
  uint64_t* z_empty = z_full + 2;           // [2] 256 arrivals: Z_t read out by the epilogue
  uint64_t* p_half = z_empty + 2;           // [2] 128 arrivals: column half 0 of H_t(c) has been read (half 1 may overwrite its P columns)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + 2);

  constexpr int W_MMA = 8, W_TMA = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.T + 127) / 128, n_items = (n_tiles + 1) / 2, n_chunks = a.F / FF_C;
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); mbar_init(&h_full[i], 1); mbar_init(&p_full[i], 256); mbar_init(&p_half[i], 128);
      mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 256);
    }
    for (int i = 0; i < FF_S1; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
    for (int i = 0; i < FF_S2; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    fence_barrier_init();
  }
  if (!BWD) for (int i = threadIdx.x; i < a.F; i += blockDim.x) sB1[i] = __ldg(a.b1 + i);
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s1 = 0, s2 = 0; uint32_t p1 = 0, p2 = 0, ni = 0, nib = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++ni) {
        const int nt = (2 * it + 1 < n_tiles) ? 2 : 1;
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&y_empty[t], ((t ? nib : ni) & 1) ^ 1);
          mbar_expect_tx(&y_full[t], FF_Y_BYTES);
#pragma unroll
          for (int kb = 0; kb < FF_KB; ++kb) tma_load_2d(sY + t * FF_Y_BYTES + kb * (128 * 128), &tmY, &y_full[t], kb * 64, (2 * it + t) * 128);
        }
        if (nt == 2) ++nib;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&w1_empty[s1], p1 ^ 1);
          mbar_expect_tx(&w1_full[s1], FF_W1_BYTES);
#pragma unroll
          for (int kb = 0; kb < FF_KB; ++kb) {
            if (!WMN) tma_load_2d(sW1 + s1 * FF_W1_BYTES + kb * (FF_C * 128), &tmW1, &w1_full[s1], kb * 64, c * FF_C);
            else tma_load_3d(sW1 + s1 * FF_W1_BYTES + kb * (FF_C * 128), &tmW1, &w1_full[s1], 0, kb * 64, c);   // W2 rows kb*64.., unit block c
          }
          if (++s1 == FF_S1) { s1 = 0; p1 ^= 1; }
          mbar_wait(&w2_empty[s2], p2 ^ 1);
          mbar_expect_tx(&w2_full[s2], FF_W2_BYTES);
          if (!WMN) tma_load_2d(sW2 + s2 * FF_W2_BYTES, &tmW2, &w2_full[s2], c * FF_C, 0);
          else tma_load_3d(sW2 + s2 * FF_W2_BYTES, &tmW2, &w2_full[s2], 0, c * FF_C, 0);   // W1 rows (units) c*64.., all FF_KB column blocks
          if (++s2 == FF_S2) { s2 = 0; p2 ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane)
    // ONE issuing warp serving both tiles in strict A, B, A, B order: the groups are large here (16 MMAs = 960 clk), so the
    // ~160 clk per group on the issuing warp is amortised, and exclusive back-to-back groups keep the ping-pong tight (with
    // one issuer per tile the two streams interleaved MMA by MMA and every group took twice as long to retire).
    constexpr uint32_t idesc_h = umma_idesc_bf16(128, FF_C, false, WMN);
    constexpr uint32_t idesc_z = umma_idesc_bf16(128, FF_D, false, WMN);
    const uint64_t y_desc0 = umma_smem_desc(smem_u32(sY), 16, 1024, 3);
    // forward: both weight chunks K-major (rows of 64 k-elements, 32 bytes per K = 16 step).  Backward: MN-major blocks of
    // [64 k-rows x 64 n-elements] (8 KB, LBO = distance between n-blocks), 2048 bytes per K = 16 step (gemm.cu, operand mode 1).
    constexpr uint32_t W_LBO = WMN ? 64 * 128 : 16, W_KSTEP = WMN ? 2048 : 32;
    const uint64_t w1_desc0 = umma_smem_desc(smem_u32(sW1), W_LBO, 1024, 3), w2_desc0 = umma_smem_desc(smem_u32(sW2), W_LBO, 1024, 3);
    int s1 = 0, s2 = 0; uint32_t p1 = 0, p2 = 0, ni = 0, nib = 0, np = 0, npb = 0;   // rings; items / p_full uses of tile A, tile B
    CB_TL_DECL(tl);
    auto mma_h = [&](int t, int slot) {                // H_t = y_t · W1[c]^T ; caller is the elected lane
      const uint64_t wd = umma_desc_add(w1_desc0, slot * FF_W1_BYTES), yd = umma_desc_add(y_desc0, t * FF_Y_BYTES);
#pragma unroll
      for (int kb = 0; kb < FF_KB; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + t * FF_C, umma_desc_add(yd, kb * (128 * 128) + k * 32), umma_desc_add(wd, kb * (FF_C * 128) + k * W_KSTEP), idesc_h, (kb > 0 || k > 0) ? 1u : 0u);
      tc_commit(&h_full[t]);
    };
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++ni) {
      const int nt = (2 * it + 1 < n_tiles) ? 2 : 1;
      mbar_wait(&y_full[0], ni & 1);
      if (nt == 2) mbar_wait(&y_full[1], nib & 1);
      mbar_wait(&w1_full[s1], p1);
      tc_fence_after();
      if (elect_one()) {
        mma_h(0, s1);
        if (nt == 2) mma_h(1, s1);
        tc_commit(&w1_empty[s1]);
      }
      __syncwarp();
      if (++s1 == FF_S1) { s1 = 0; p1 ^= 1; }
      for (int c = 0; c < n_chunks; ++c) {
        const bool more = c + 1 < n_chunks;
        CB_TL(0, tl, 1);
        mbar_wait(&w2_full[s2], p2);
        if (more) mbar_wait(&w1_full[s1], p1);
        CB_TL(0, tl, 2);
        for (int t = 0; t < nt; ++t) {
          const uint32_t npt = t ? npb : np, nit = t ? nib : ni;
          mbar_wait(&p_full[t], npt & 1);
          if (c == 0 && nit > 0) mbar_wait(&z_empty[t], (nit - 1) & 1);   // the previous item's Z_t has been read out (overwritten below)
          tc_fence_after();
          CB_TL(0, tl, 3 + t);
          if (elect_one()) {
            const uint64_t w2d = umma_desc_add(w2_desc0, s2 * FF_W2_BYTES);
#pragma unroll
            for (int kk = 0; kk < FF_C / 16; ++kk)     // Z_t += P_t(c) · W2[:, c]^T
              umma_ts(tmem_base + FF_COL_Z + t * FF_D, tmem_base + t * FF_C + kk * 8, umma_desc_add(w2d, kk * W_KSTEP), idesc_z, (c > 0 || kk > 0) ? 1u : 0u);
            if (more) mma_h(t, s1);                    // H_t(c+1) over P_t(c): the in-order pipe has retired its reader by then
            else { tc_commit(&y_empty[t]); tc_commit(&z_full[t]); }
          }
          __syncwarp();
          if (t) ++npb; else ++np;
        }
        if (elect_one()) { tc_commit(&w2_empty[s2]); if (more) tc_commit(&w1_empty[s1]); }
        __syncwarp();
        CB_TL(0, tl, 5);
        if (more && ++s1 == FF_S1) { s1 = 0; p1 ^= 1; }
        if (++s2 == FF_S2) { s2 = 0; p2 ^= 1; }
      }
      if (nt == 2) ++nib;
    }
  } else {
    // ------------------------------------------------------------------ bias + ReLU warps
    // ALL eight warps serve every tile: warp w owns TMEM lanes 32 (w & 3) .. (rows) and the column half hf = w >> 2 of the
    // 64-wide chunk (bias + ReLU is element-wise: no exchange between the halves).  That halves the H -> P latency on the
    // critical chain  H_t(c) -> P_t(c) -> Z_t += / H_t(c+1)  compared with one warpgroup per tile.
    const int hf = warp >> 2, q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    uint32_t nh0 = 0, nh1 = 0, ni0 = 0, ni1 = 0;   // h_full uses / items so far of tile A, tile B
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4) && lane == 0;
    // BWD: the ReLU mask word of this thread's 32 units of (item, chunk, tile) — a warp reads one 128-byte line.  It sits on the
    // critical chain H(c) -> P(c) -> Z += / H(c+1), so it is requested ONE CHUNK AHEAD (for the same tile, across item
    // boundaries too) and is in a register when the accumulator arrives; mwA / mwB hold the words of tile A / B.
    auto load_mw = [&](int it_, int c_, int t_) -> uint32_t {
      if (c_ == n_chunks) { c_ = 0; it_ += gridDim.x; }
      const long row_ = (long)(2 * it_ + t_) * 128 + r_in_tile;
      return (it_ < n_items && row_ < a.T) ? ldg_u32_pinned(a.mask_bits + (long)(2 * c_ + hf) * a.ld_bits + row_) : 0u;
    };
    uint32_t mwA = 0u, mwB = 0u;
    if (BWD && (int)blockIdx.x < n_items) { mwA = load_mw(blockIdx.x, 0, 0); mwB = load_mw(blockIdx.x, 0, 1); }
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int nt = (2 * it + 1 < n_tiles) ? 2 : 1;
      for (int c = 0; c < n_chunks; ++c) {
        // BWD: the two-tile loop is unrolled so that mwA / mwB are statically named registers — in the rolled loop the hand-over
        // `if (t) mwB = nx; else mwA = nx;` is a select that DEPENDS on the load just issued and blocks the warp for the whole
        // global-load latency (in-kernel timeline: 1550 clk in front of every accumulator wait, profiles/r02_timeline_ffn_bwd.txt).
#pragma unroll (BWD ? 2 : 1)   // forward: unrolling changes nothing (153.1 vs 152.7 us, same-box A/B), the rolled loop is half the code
        for (int t = 0; t < 2; ++t) {
          if (t >= nt) break;
          const long row = (long)(2 * it + t) * 128 + r_in_tile;
          const uint32_t h_addr = lane_addr + t * FF_C + hf * 32;
          if (tl_on) CB_TL(1 + hf, tl, 1 + 4 * t);
          const uint32_t mw = t ? mwB : mwA;
          if (BWD) { const uint32_t nx = load_mw(it, c + 1, t); if (t) mwB = nx; else mwA = nx; }
          mbar_wait(&h_full[t], (t ? nh1 : nh0) & 1);
          if (t) ++nh1; else ++nh0;
          tc_fence_after();
          if (tl_on) CB_TL(1 + hf, tl, 2 + 4 * t);
          uint32_t r0[32];
          tmem_ld32(h_addr, r0);
          tmem_ld_wait();
          uint32_t pk[16];
          if (!BWD) {
            const float* bp = sB1 + c * FF_C + hf * 32;    // bias of this half chunk: broadcast 16-byte smem loads
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 ba = *reinterpret_cast<const float4*>(bp + e);
              pk[e >> 1] = pack_bf16(fmaxf(__uint_as_float(r0[e]) + ba.x, 0.f), fmaxf(__uint_as_float(r0[e + 1]) + ba.y, 0.f));
              pk[(e >> 1) + 1] = pack_bf16(fmaxf(__uint_as_float(r0[e + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(r0[e + 3]) + ba.w, 0.f));
            }
          } else {   // d(hidden) = (dz2 . W2) where hidden > 0: bit 2e / 2e+1 of the mask word <=> low / high bf16 of pk[e] (relu_bits16)
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const uint32_t b2 = mw >> (2 * e);
              const uint32_t sel = ((b2 & 1u) ? 0xffffu : 0u) | ((b2 & 2u) ? 0xffff0000u : 0u);
              pk[e] = pack_bf16(__uint_as_float(r0[2 * e]), __uint_as_float(r0[2 * e + 1])) & sel;
            }
          }
          // P (bf16): hidden units 32 hf .. 32 hf + 31 of the chunk -> TMEM columns 16 hf .. 16 hf + 15 of H_t.  Half 1 writes
          // columns 16..31, which belong to half 0's fp32 input range: wait until half 0 has read its columns (it arrives on
          // p_half after its tcgen05.ld) — half 0 itself only overwrites columns it has already read.
          if (hf == 1) { mbar_wait(&p_half[t], (t ? nh1 : nh0) & 1 ^ 1); tc_fence_after(); }
          else { tc_fence_before(); mbar_arrive(&p_half[t]); }
          tmem_st16(lane_addr + t * FF_C + hf * 16, pk);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[t]);
          if (tl_on) CB_TL(1 + hf, tl, 3 + 4 * t);
          if (a.hid && row < a.T) {                      // hidden activations kept for the backward pass: 64 contiguous bytes per row
            __nv_bfloat16* dst = a.hid + row * a.F + c * FF_C + hf * 32;
            stg256(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            stg256(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
          }
          if (!BWD && a.mask_bits && row < a.T) a.mask_bits[(long)(2 * c + hf) * a.ld_bits + row] = relu_bits16(pk);   // a warp = 32 consecutive rows: one line
        }
      }
      // ---- final epilogue: z2 = Z + b2 + resid (fp32); this warp's half: 32-column slabs hf, hf + 2, hf + 4 (128 contiguous bytes each)
#pragma unroll 1
      for (int t = 0; t < nt; ++t) {
        const long row = (long)(2 * it + t) * 128 + r_in_tile;
        const bool row_ok = row < a.T;
        const uint32_t z_addr = lane_addr + FF_COL_Z + t * FF_D;
        mbar_wait(&z_full[t], (t ? ni1 : ni0) & 1);
        if (t) ++ni1; else ++ni0;
        tc_fence_after();
#pragma unroll 1
        for (int s = hf; s < FF_D / 32; s += 2) {
          uint32_t res[4][8];
          if (row_ok) {
            const float* rp = a.resid + row * FF_D + s * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) ldg256(rp + 8 * k, res[k]);
          }
          uint32_t x[32];
          tmem_ld32(z_addr + s * 32, x);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = a.z2 + row * FF_D + s * 32;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (!BWD) { b0 = __ldg(reinterpret_cast<const float4*>(a.b2 + s * 32 + 8 * k)); b1 = __ldg(reinterpret_cast<const float4*>(a.b2 + s * 32 + 8 * k + 4)); }
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(x[8 * k + e]) + bb[e] + __uint_as_float(res[k][e]));
              stg256(dst + 8 * k, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&z_empty[t]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

}  // namespace cb

#ifdef CB_TIMELINE
extern "C" int cb_debug_timeline_ffn(void* dst) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif

extern "C" int cb_ffn_fwd(const void* y, const void* w1, const float* b1, const void* w2, const float* b2, const float* resid, float* z2,
                          void* hid, unsigned int* mask_bits, int ld_bits, int T, int D, int F, int kernel, void* stream) {
  using namespace cb;
  CB_CHECK(T > 0 && D == FF_D && F % FF_C == 0 && F >= FF_C && F <= FF_MAX_F, "ffn_fwd: T=%d D=%d F=%d (this kernel handles D = %d, F a multiple of %d up to %d)", T, D, F, FF_D, FF_C, FF_MAX_F);
  CB_CHECK(!mask_bits || (ld_bits >= T && ld_bits % 32 == 0 && (reinterpret_cast<uintptr_t>(mask_bits) & 127) == 0),
           "ffn_fwd: mask_bits needs ld_bits >= T (a multiple of 32) and a 128-byte aligned buffer");
  CB_CHECK(((reinterpret_cast<uintptr_t>(resid) | reinterpret_cast<uintptr_t>(z2) | reinterpret_cast<uintptr_t>(hid) | reinterpret_cast<uintptr_t>(b1) |
             reinterpret_cast<uintptr_t>(b2)) & 31) == 0, "ffn_fwd: resid / z2 / hid / biases must be 32-byte aligned");
  // Two kernels (measured at T = 68664 / 137328, F = 2048: profiles/r01_microbench_kernels.txt):
  //   ffn_fwd3.cu  cluster of two, 128-unit chunks          122 / 230 us without the hidden store, 165 / 319 us with it
  //   this file    pair of tiles per CTA                    142 / 275 us                            155 / 299 us
  // The per-thread 32-byte stores of the hidden activations go through the same L1/shared-memory pipe as the SS operand
  // reads that bound the cluster kernel, so the student pass (hidden stored) stays on this one.  `kernel` = 1 | 3 forces one of
  // them (A/B measurements, tests); 0 = this choice.
  CB_CHECK(kernel == 0 || kernel == 1 || kernel == 3, "ffn_fwd: kernel must be 0 (auto), 1 or 3");
  const int use = kernel ? kernel : (hid == nullptr ? 3 : 1);
  if (use == 3 && F % 128 == 0) return ffn_fwd3_run(y, w1, b1, w2, b2, resid, z2, hid, mask_bits, ld_bits, T, F, reinterpret_cast<cudaStream_t>(stream));
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(ffn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ty, t1, t2;
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)T}; uint64_t strides[1] = {(uint64_t)D * 2}; uint32_t box[2] = {64, 128};
    if (make_tmap(&ty, y, 2, dims, strides, box, 3)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)F}; uint64_t strides[1] = {(uint64_t)D * 2}; uint32_t box[2] = {64, (uint32_t)FF_C};
    if (make_tmap(&t1, w1, 2, dims, strides, box, 3)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)F, (uint64_t)D}; uint64_t strides[1] = {(uint64_t)F * 2}; uint32_t box[2] = {64, (uint32_t)FF_D};
    if (make_tmap(&t2, w2, 2, dims, strides, box, 3)) return 1;
  }
  FfnArgs a{};
  a.b1 = b1; a.b2 = b2; a.resid = resid; a.z2 = z2; a.hid = reinterpret_cast<__nv_bfloat16*>(hid); a.mask_bits = mask_bits; a.ld_bits = ld_bits; a.T = T; a.F = F;
  const int n_items = ((T + 127) / 128 + 1) / 2;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  ffn_kernel<false><<<grid, 320, FF_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(ty, t1, t2, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cb_ffn_bwd(const void* dz2_bf16, const void* w2, const void* w1, const unsigned int* mask_bits, int ld_bits, const float* dz2,
                          float* dy, void* dh, int T, int D, int F, void* stream) {
  using namespace cb;
  CB_CHECK(T > 0 && D == FF_D && F % FF_C == 0 && F >= FF_C && F <= FF_MAX_F, "ffn_bwd: T=%d D=%d F=%d (this kernel handles D = %d, F a multiple of %d up to %d)", T, D, F, FF_D, FF_C, FF_MAX_F);
  CB_CHECK(mask_bits && dh && dz2 && dy && dz2_bf16 && w1 && w2, "ffn_bwd: null argument");
  CB_CHECK(ld_bits >= T && (reinterpret_cast<uintptr_t>(mask_bits) & 3) == 0, "ffn_bwd: mask_bits is uint32 [F/32, ld_bits >= T]");
  CB_CHECK(((reinterpret_cast<uintptr_t>(dz2) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dh)) & 31) == 0 &&
           ((reinterpret_cast<uintptr_t>(dz2_bf16) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2)) & 15) == 0,
           "ffn_bwd: dz2 / dy / dh must be 32-byte, the bf16 operands 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(ffn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ty, t1, t2;
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)T}; uint64_t strides[1] = {(uint64_t)D * 2}; uint32_t box[2] = {64, 128};
    if (make_tmap(&ty, dz2_bf16, 2, dims, strides, box, 3)) return 1;
  }
  FfnArgs a{};
  a.resid = dz2; a.z2 = dy; a.hid = reinterpret_cast<__nv_bfloat16*>(dh); a.mask_bits = const_cast<uint32_t*>(mask_bits); a.ld_bits = ld_bits; a.T = T; a.F = F;
  const int n_items = ((T + 127) / 128 + 1) / 2;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  {   // first product: B = W2 [D, F] row-major = [K, N]: (64 n, K, N / 64) boxes of one [64 k x 64 n] block
    uint64_t dims[3] = {64, (uint64_t)D, (uint64_t)(F / 64)}; uint64_t strides[2] = {(uint64_t)F * 2, 128}; uint32_t box[3] = {64, 64, 1};
    if (make_tmap(&t1, w2, 3, dims, strides, box, 3)) return 1;
  }
  {   // second product: B = W1 [F, D] row-major = [K, N]: boxes of [64 k (units) x D n] = FF_KB blocks
    uint64_t dims[3] = {64, (uint64_t)F, (uint64_t)(D / 64)}; uint64_t strides[2] = {(uint64_t)D * 2, 128}; uint32_t box[3] = {64, (uint32_t)FF_C, (uint32_t)FF_KB};
    if (make_tmap(&t2, w1, 3, dims, strides, box, 3)) return 1;
  }
  ffn_kernel<true, true><<<grid, 320, FF_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(ty, t1, t2, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}
