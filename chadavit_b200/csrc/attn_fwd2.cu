// Varlen multi-head self-attention forward (the SDPA inside nn.MultiheadAttention, chada_vit.py:105-111) for sm_100a.
//
// A work item is (sequence, head, PAIR of 128-query tiles A, B): every K/V tile fetched from L2 serves 256 query rows.
// Roles: warps 0-3 softmax of tile A, warps 4-7 softmax of tile B (1 thread = 1 query row, no cross-thread reductions),
// warps 8 / 9 MMA issuers of tile A / B (whole warp convergent, one elected lane issues: back-to-back UTCHMMA), warp 10 TMA.
//
// The KV dimension advances in 64-row SUB-TILES and every query tile owns TWO score buffers; Q and K/V come through TMA:
//     TMEM columns:  S_A0 64 | S_A1 64 | S_B0 64 | S_B1 64 | O_A 128 | O_B 128   = 512
//     shared memory: Q_A, Q_B (one [128 x d] tile each) + a K/V ring of [64 x d] sub-tiles (up to 8 stages)
//   * two score buffers: the MMA stream  S_A(0) S_B(0) S_A(1) S_B(1) | PV_A(0) S_A(2) | PV_B(0) S_B(2) | PV_A(1) S_A(3) ...  has
//     the next score tile of a warpgroup finished before that warpgroup is done with the current one.  The in-kernel timeline
//     (tools/timeline.py) of the earlier design (one 128-wide score buffer per query tile) showed each warpgroup idle for a
//     tensor-pipe round trip (PV + next QK) after every tile, and the two warpgroups then colliding on the MUFU unit
//     (4 ex2/clk/SMSP: 2048 clk per 128x256 score block against 1536 clk of MMA — the binding resource at d = 96).
//   * Measured and not kept: 48-row sub-tiles with Q in tensor memory (630 TFLOP/s against 704), 8 softmax
//     warps row-split on one tile, software prefetch of the next S.
//
// Softmax details: bf16 P overwrites the first half of its own S buffer and feeds O += P·V as a TS MMA (V read MN-major from
// its TMA tile).  S is read once per sub-tile: probabilities are computed speculatively against the running reference
// maximum and kept packed in registers; only when a row outgrows the reference by 2^8 (always checked, rarely true) is the
// sub-tile redone and O rescaled.  The scale is folded into one FFMA per element; only the ragged last sub-tile carries
// masking instructions; sub-tiles past the end of a sequence are never issued.
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

template <int HD>
struct Att2Cfg {
  static constexpr int KVT = 64;                                  // kv rows per sub-tile
  static constexpr int CHUNK = (HD % 64 == 0) ? 64 : (HD % 32 == 0 ? 32 : 16);   // d elements per swizzle chunk
  static constexpr int NCH = HD / CHUNK;
  static constexpr int SWZ = CHUNK == 64 ? 3 : (CHUNK == 32 ? 2 : 1);
  static constexpr int SBO = 8 * CHUNK * 2;                       // 8 rows of one chunk
  static constexpr int Q_CHUNK_BYTES = 128 * CHUNK * 2;
  static constexpr int Q_TILE_BYTES = NCH * Q_CHUNK_BYTES;        // one [128 x HD] bf16 query tile
  static constexpr int KV_CHUNK_BYTES = KVT * CHUNK * 2;
  static constexpr int KV_TILE_BYTES = NCH * KV_CHUNK_BYTES;      // one [KVT x HD] bf16 sub-tile
  static constexpr int AUX_BYTES = 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int NS_FIT = (227 * 1024 - 2 * Q_TILE_BYTES - AUX_BYTES) / (2 * KV_TILE_BYTES);
  static constexpr int KV_STAGES = NS_FIT > 8 ? 8 : NS_FIT;       // K/V ring depth (>= 3: QK runs two sub-tiles ahead of PV)
  static constexpr int SMEM_BYTES = 2 * Q_TILE_BYTES + 2 * KV_STAGES * KV_TILE_BYTES + AUX_BYTES;
  // TMEM columns: S buffer (t, b) at (2 t + b) KVT ; O_t at 256 + 128 t
  static constexpr int COL_O = 256;
  static_assert(KV_STAGES >= 3, "K/V ring too shallow");
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic minimax polynomial for 2^f
// (max relative error 1.0e-4, well below the bf16 rounding of P), n added into the exponent field.  Every fourth probability
// of the forward softmax takes this path: the MUFU unit (4 ex2/clk/SMSP) is the binding resource at d = 96, while the SMSP
// issue slots are only ~30 % used, so moving a quarter of the exponentials to ~8 FMA/ALU instructions shortens the critical
// resource by 25 % (the FA4 trick).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);                                  // -inf (masked) / underflow -> 2^-126 ~ 0
  const float xr = x + 12582912.f;                       // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (xr - 12582912.f);
  float p = fmaf(f, 0.05500893f, 0.24221095f);
  p = fmaf(p, f, 0.6932829f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

// ---- softmax helpers: a thread owns one query row of a sub-tile = one TMEM lane, KVT = 64 consecutive fp32 score columns.
__device__ __forceinline__ float mask_col(float x, int col, int kvv) { return col >= kvv ? -INFINITY : x; }

// Row maximum; RAGGED masks columns >= kvv (valid kv positions of this sub-tile).
template <bool RAGGED>
__device__ __forceinline__ float sub_row_max(uint32_t s_addr, int kvv) {
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  uint32_t r[2][32];
  tmem_ld32(s_addr, r[0]);
  tmem_ld32(s_addr + 32, r[1]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    float x0 = __uint_as_float(r[i >> 5][i & 31]), x1 = __uint_as_float(r[i >> 5][(i + 1) & 31]);
    float x2 = __uint_as_float(r[i >> 5][(i + 2) & 31]), x3 = __uint_as_float(r[i >> 5][(i + 3) & 31]);
    if (RAGGED) { x0 = mask_col(x0, i, kvv); x1 = mask_col(x1, i + 1, kvv); x2 = mask_col(x2, i + 2, kvv); x3 = mask_col(x3, i + 3, kvv); }
    mx0 = fmaxf(mx0, x0); mx1 = fmaxf(mx1, x1); mx2 = fmaxf(mx2, x2); mx3 = fmaxf(mx3, x3);
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// p = exp2(s * scale_log2 + neg_m) for the 64 scores, packed to bf16 in pk; returns their sum, mx_out = their maximum in the
// exp2 domain.  The TMEM read of the second 32 columns is in flight under the first 32 columns' MUFU work.
#ifndef CB_ATTN_POLY
#define CB_ATTN_POLY 0   // measured: no gain yet (688 vs 690 TFLOP/s) — the exp phase is latency-, not MUFU-bound; kept for the next round
#endif
template <bool RAGGED>
__device__ __forceinline__ float sub_exp(uint32_t s_addr, float scale_log2, float neg_m, int kvv, uint32_t (&pk)[32], float& mx_out) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  uint32_t r[2][32];
  tmem_ld32(s_addr, r[0]);
  tmem_ld_wait();
  tmem_ld32(s_addr + 32, r[1]);
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    if (i == 32) tmem_ld_wait();
    float x0 = __uint_as_float(r[i >> 5][i & 31]), x1 = __uint_as_float(r[i >> 5][(i + 1) & 31]);
    float x2 = __uint_as_float(r[i >> 5][(i + 2) & 31]), x3 = __uint_as_float(r[i >> 5][(i + 3) & 31]);
    if (RAGGED) { x0 = mask_col(x0, i, kvv); x1 = mask_col(x1, i + 1, kvv); x2 = mask_col(x2, i + 2, kvv); x3 = mask_col(x3, i + 3, kvv); }
    mx0 = fmaxf(mx0, x0); mx1 = fmaxf(mx1, x1); mx2 = fmaxf(mx2, x2); mx3 = fmaxf(mx3, x3);
    const float p0 = ex2f(fmaf(x0, scale_log2, neg_m)), p1 = ex2f(fmaf(x1, scale_log2, neg_m));
    const float p2 = ex2f(fmaf(x2, scale_log2, neg_m)), p3 = CB_ATTN_POLY ? ex2_poly(fmaf(x3, scale_log2, neg_m)) : ex2f(fmaf(x3, scale_log2, neg_m));
    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
    pk[i >> 1] = pack_bf16(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16(p2, p3);
  }
  mx_out = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
  return (s0 + s1) + (s2 + s3);
}

struct Attn2Args {
  const int4* work;  // {q_row0 (global row of tile A), seq_start, seq_end, head}; seq_end <= seq_start: empty slot
  int n_work;
  __nv_bfloat16* out;
  float* lse;
  int T, D;
  float scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(352, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Attn2Args a) {
  using Cfg = Att2Cfg<HD>;
  constexpr int NS = Cfg::KV_STAGES, KVT = Cfg::KVT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sQ = smem;                                   // [2] query tiles
  uint8_t* sK = sQ + 2 * Cfg::Q_TILE_BYTES;             // [NS] KVT-row sub-tiles
  uint8_t* sV = sK + NS * Cfg::KV_TILE_BYTES;           // [NS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * Cfg::KV_TILE_BYTES);
  uint64_t* q_full = bars + 0;              // [2] query tile t landed
  uint64_t* q_empty = bars + 2;             // [2] every Q_t·K^T of the item has retired
  uint64_t* kv_full = bars + 4;             // [NS] K and V sub-tile of a stage landed (one transaction barrier for both)
  uint64_t* kv_empty = kv_full + NS;        // [NS] 2 arrivals: both MMA warps are done with the stage
  uint64_t* s_full = kv_empty + NS;         // [2 tiles][2 buffers]
  uint64_t* p_full = s_full + 4;            // [2][2], 128 arrivals each
  uint64_t* pv_done = p_full + 4;           // [2]
  uint64_t* o_full = pv_done + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  // warps 0-3 / 4-7: softmax of tile A / B; warp 8 / 9: MMA issuer of tile A / B; warp 10: TMA producer
  constexpr int W_MMA = 8, W_TMA = 10;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&pv_done[i], 1); mbar_init(&o_full[i], 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int st = 0; uint32_t ph = 0, nqa = 0, nqb = 0;
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const int4 wk = a.work[w];
        if (wk.z <= wk.y) continue;
        const int head = wk.w;
        const int n_sub = (wk.z - wk.y + KVT - 1) / KVT;
        const int nt = (wk.x + 128 < wk.z) ? 2 : 1;
        for (int t = 0; t < nt; ++t) {
          uint32_t& nq = t ? nqb : nqa;
          mbar_wait(&q_empty[t], (nq & 1) ^ 1);
          ++nq;
          mbar_expect_tx(&q_full[t], Cfg::Q_TILE_BYTES);
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c)
            tma_load_2d(sQ + t * Cfg::Q_TILE_BYTES + c * Cfg::Q_CHUNK_BYTES, &tmQ, &q_full[t], head * HD + c * Cfg::CHUNK, wk.x + t * 128);
        }
        for (int i = 0; i < n_sub; ++i) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          const int row = wk.y + i * KVT;
          mbar_expect_tx(&kv_full[st], 2 * Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c)
            tma_load_2d(sK + st * Cfg::KV_TILE_BYTES + c * Cfg::KV_CHUNK_BYTES, &tmKV, &kv_full[st], a.D + head * HD + c * Cfg::CHUNK, row);
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c)
            tma_load_2d(sV + st * Cfg::KV_TILE_BYTES + c * Cfg::KV_CHUNK_BYTES, &tmKV, &kv_full[st], 2 * a.D + head * HD + c * Cfg::CHUNK, row);
          if (++st == NS) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA || warp == W_MMA + 1) {
    // ------------------------------------------------------------------ MMA issuers: warp 8 drives query tile A, warp 9 tile B
    // Each is a whole convergent warp with one elected issuing lane (descriptors stay in uniform registers, UTCHMMAs back to
    // back; under `if (lane == 0)` every tcgen05.mma cost ~12 instructions).  Two issuers because the per-group cost on the
    // issuing warp (mbarrier try_wait ~90 clk even when the phase is already complete, fence, elect, commit: ~160 clk per
    // group, tools/ubench/handoff.cu) exceeded the MMA time of the groups when one warp served both tiles; the in-order
    // tensor pipe interleaves the two independent streams.
    const int t = warp - W_MMA;
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, KVT, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, false, true);
    const uint64_t q_desc = umma_smem_desc(smem_u32(sQ + t * Cfg::Q_TILE_BYTES), 16, Cfg::SBO, Cfg::SWZ);
    const uint64_t k_desc0 = umma_smem_desc(smem_u32(sK), 16, Cfg::SBO, Cfg::SWZ);
    const uint64_t v_desc0 = umma_smem_desc(smem_u32(sV), Cfg::KV_CHUNK_BYTES, Cfg::SBO, Cfg::SWZ);   // MN-major: LBO = chunk stride
    const uint32_t s_col = tmem_base + t * 2 * KVT, o_col = tmem_base + Cfg::COL_O + t * 128;
    int kst = 0, vst = 0; uint32_t kph = 0;   // ring position of the next K sub-tile for QK (+ its phase) / of the V sub-tile for PV
    uint32_t pbits = 0, nq = 0;               // p_full[t][b] phase parities, items of this tile so far (q_full phase)
    CB_TL_DECL(tl);
    auto mma_qk = [&](int b, int stage) {     // S_t[b] = Q_t · K_stage^T ; caller is the elected lane
      const uint64_t kd = umma_desc_add(k_desc0, stage * Cfg::KV_TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const int c = (kk * 16) / Cfg::CHUNK, off = ((kk * 16) % Cfg::CHUNK) * 2;
        umma_ss(s_col + b * KVT, umma_desc_add(q_desc, c * Cfg::Q_CHUNK_BYTES + off), umma_desc_add(kd, c * Cfg::KV_CHUNK_BYTES + off), idesc_qk,
                kk > 0 ? 1u : 0u);
      }
      tc_commit(&s_full[t * 2 + b]);
    };
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      const int4 wk = a.work[w];
      if (wk.z <= wk.y) continue;
      const int n_sub = (wk.z - wk.y + KVT - 1) / KVT;
      const bool two = wk.x + 128 < wk.z;
      if (t == 1 && !two) {
        // No tile B in this item: walk the K/V ring anyway, releasing every stage on behalf of this warp.  The stages must be
        // WAITED for one by one — an mbarrier parity wait is only meaningful for the current phase, so a consumer may never
        // get a whole ring pass ahead of the fills (skipping by arithmetic made later waits return on a stale phase).
        for (int i = 0; i < n_sub; ++i) {
          mbar_wait(&kv_full[kst], kph);
          if (elect_one()) tc_commit(&kv_empty[kst]);
          __syncwarp();
          if (++kst == NS) { kst = 0; kph ^= 1; }
        }
        vst = kst;
        continue;
      }
      mbar_wait(&q_full[t], nq & 1);
      ++nq;
      // prologue: the first two score sub-tiles
      for (int i = 0; i < 2 && i < n_sub; ++i) {
        mbar_wait(&kv_full[kst], kph);
        tc_fence_after();
        if (elect_one()) { mma_qk(i, kst); if (i + 1 == n_sub) tc_commit(&q_empty[t]); }
        __syncwarp();
        if (++kst == NS) { kst = 0; kph ^= 1; }
      }
      for (int i = 0; i < n_sub; ++i) {
        const int b = i & 1;
        const bool more = i + 2 < n_sub;
        if (more) mbar_wait(&kv_full[kst], kph);         // K(i+2) landed (V(i) was covered when its K was waited for)
        CB_TL(t * 3, tl, 1);
        mbar_wait(&p_full[t * 2 + b], (pbits >> b) & 1);
        pbits ^= 1u << b;
        tc_fence_after();
        CB_TL(t * 3, tl, 2);
        if (elect_one()) {
          const uint64_t vd = umma_desc_add(v_desc0, vst * Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < KVT / 16; ++kk)       // O_t += P_t[b] · V_i
            umma_ts(o_col, s_col + b * KVT + kk * 8, umma_desc_add(vd, kk * 16 * Cfg::CHUNK * 2), idesc_pv, (i > 0 || kk > 0) ? 1u : 0u);
          tc_commit(&pv_done[t]);
          if (more) {
            mma_qk(b, kst);                            // S_t(i+2) into the buffer PV(i) has just consumed (in-order pipe)
            if (i + 3 == n_sub) tc_commit(&q_empty[t]);   // that was the last Q·K^T of the item: Q_t may be refilled
          }
          tc_commit(&kv_empty[vst]);                   // K_i / V_i: this warp is done once everything issued so far has retired
          if (i + 1 == n_sub) tc_commit(&o_full[t]);
        }
        __syncwarp();
        CB_TL(t * 3, tl, 3);
        if (more && ++kst == NS) { kst = 0; kph ^= 1; }
        if (++vst == NS) vst = 0;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (tile t = 0: warps 0-3, 1: warps 4-7)
    const int t = warp >> 2;
    const int q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t o_addr = lane_addr + Cfg::COL_O + t * 128;
    uint32_t sbits = 0, npv = 0, ow = 0;   // s_full[t][b] phase parities, PV commits of this tile so far, items of this tile so far
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4) && lane == 0;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      const int4 wk = a.work[w];
      const int q0 = wk.x + t * 128;
      if (wk.z <= wk.y || q0 >= wk.z) continue;         // empty slot / odd tail: this item has no second tile
      const int seq_len = wk.z - wk.y;
      const int n_sub = (seq_len + KVT - 1) / KVT;
      float m_used = -INFINITY, l = 0.f;
      for (int i = 0; i < n_sub; ++i, ++npv) {
        const int b = i & 1;
        const uint32_t s_addr = lane_addr + (t * 2 + b) * KVT;
        if (tl_on) CB_TL(1 + t, tl, 1);
        mbar_wait(&s_full[t * 2 + b], (sbits >> b) & 1);
        sbits ^= 1u << b;
        tc_fence_after();
        if (tl_on) CB_TL(1 + t, tl, 2);
        const int kvv = seq_len - i * KVT;
        const bool ragged = kvv < KVT;
        if (i == 0) m_used = (ragged ? sub_row_max<true>(s_addr, kvv) : sub_row_max<false>(s_addr, kvv)) * a.scale_log2;
        uint32_t pk[32];
        float tsum, alpha = 1.f;
        bool any_need = false;
#pragma unroll 1
        for (int pass = 0;; ++pass) {                    // one code copy of the exp pass; the second trip is rare (warp-uniform)
          float mx_seen;
          tsum = ragged ? sub_exp<true>(s_addr, a.scale_log2, -m_used, kvv, pk, mx_seen) : sub_exp<false>(s_addr, a.scale_log2, -m_used, kvv, pk, mx_seen);
          const bool need = pass == 0 && mx_seen > m_used + 8.f;   // lazy rescale: the reference moves only when outgrown by 2^8
          if (!__any_sync(0xffffffffu, need)) break;
          any_need = true;
          if (need) { alpha = ex2f(m_used - mx_seen); m_used = mx_seen; }
        }
        if (tl_on) CB_TL(1 + t, tl, any_need ? 13 : 3);
#pragma unroll
        for (int c = 0; c < 2; ++c) {                    // P (bf16) into the first 32 columns of this S buffer
          uint32_t t16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) t16[k] = pk[c * 16 + k];
          tmem_st16(s_addr + c * 16, t16);
        }
        l = l * alpha + tsum;
        if (i > 0 && any_need) {
          mbar_wait(&pv_done[t], (npv - 1) & 1);         // O complete (PV of the previous sub-tile retired) before the rescale
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD; c += 16) {
            uint32_t o[16];
            tmem_ld16(o_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st16(o_addr + c, o);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t * 2 + b]);
        if (tl_on) CB_TL(1 + t, tl, 4);
      }
      // ---- epilogue: O / l -> bf16, LSE
      mbar_wait(&o_full[t], ow & 1);
      ++ow;
      tc_fence_after();
      const int grow = q0 + r_in_tile;
      const bool ok = grow < wk.z;
      const float inv_l = 1.f / l;
      __nv_bfloat16* dst = a.out + (long)grow * a.D + wk.w * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 16) {
        uint32_t o[16];
        tmem_ld16(o_addr + c, o);
        tmem_ld_wait();
        if (ok) {   // one full 32-byte sector per store (see attn_bwd2.cu: 16-byte halves write every sector twice, partially)
          stg256(dst + c,
                 pack_bf16(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l), pack_bf16(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l),
                 pack_bf16(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l), pack_bf16(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l),
                 pack_bf16(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l), pack_bf16(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l),
                 pack_bf16(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l), pack_bf16(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l));
        }
      }
      if (ok && a.lse) a.lse[(long)wk.w * a.T + grow] = (m_used + log2f(l)) * 0.6931471805599453f;
      tc_fence_before();
      if (tl_on) CB_TL(1 + t, tl, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

template <int HD>
static int launch_fwd2(const void* qkv, const Attn2Args& a, cudaStream_t stream) {
  using Cfg = Att2Cfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tq, tkv;
  uint64_t dims[2] = {(uint64_t)(3 * a.D), (uint64_t)a.T};
  uint64_t strides[1] = {(uint64_t)(3 * a.D) * 2};
  uint32_t box_q[2] = {(uint32_t)Cfg::CHUNK, 128}, box_kv[2] = {(uint32_t)Cfg::CHUNK, (uint32_t)Cfg::KVT};
  if (make_tmap(&tq, qkv, 2, dims, strides, box_q, Cfg::SWZ)) return 1;
  if (make_tmap(&tkv, qkv, 2, dims, strides, box_kv, Cfg::SWZ)) return 1;
  const int grid = a.n_work < num_sms() ? a.n_work : num_sms();
  attn_fwd2_kernel<HD><<<grid, 352, Cfg::SMEM_BYTES, stream>>>(tq, tkv, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

static int attn_fwd2_run(const void* qkv, const int* work, int n_work, void* out, float* lse, int T, int D, int H, float softmax_scale,
                         cudaStream_t s) {
  Attn2Args a{};
  a.work = reinterpret_cast<const int4*>(work); a.n_work = n_work; a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.T = T; a.D = D; a.scale_log2 = softmax_scale * 1.4426950408889634f;
  switch (D / H) {
    case 16: return launch_fwd2<16>(qkv, a, s);
    case 32: return launch_fwd2<32>(qkv, a, s);
    case 64: return launch_fwd2<64>(qkv, a, s);
    case 96: return launch_fwd2<96>(qkv, a, s);
    case 128: return launch_fwd2<128>(qkv, a, s);
    default: set_error("attn_fwd: unsupported head_dim %d (supported: 16, 32, 64, 96, 128)", D / H); return 1;
  }
}

}  // namespace cb

extern "C" int cb_attn_varlen_fwd(const void* qkv, const int* work, int n_work, int q_tile, void* out, float* lse, int T, int D, int H,
                                  float softmax_scale, void* stream) {
  using namespace cb;
  CB_CHECK(T > 0 && H > 0 && D % H == 0 && n_work > 0, "attn_fwd: bad shape T=%d D=%d H=%d n_work=%d", T, D, H, n_work);
  CB_CHECK((3 * D) % 8 == 0, "attn_fwd: 3*D must be a multiple of 8");
  CB_CHECK(D % 16 == 0 && (D / H) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0,
           "attn_fwd: the output leaves as 32-byte sectors: D and head_dim must be multiples of 16 and out 32-byte aligned (D=%d H=%d)", D, H);
  CB_CHECK(q_tile == 256, "attn_fwd: q_tile must be 256 (work items are pairs of 128-row query tiles)");
  return attn_fwd2_run(qkv, work, n_work, out, lse, T, D, H, softmax_scale, reinterpret_cast<cudaStream_t>(stream));
}

#ifdef CB_TIMELINE
extern "C" int cb_debug_timeline_fwd(void* dst) {   // host buffer of CB_TL_ROLES * CB_TL_LEN u64; clears the device copy
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif
