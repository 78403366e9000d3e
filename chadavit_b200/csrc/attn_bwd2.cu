// Varlen multi-head self-attention backward, generation 2 (head_dim <= 96) — autograd of the SDPA inside
// nn.MultiheadAttention (chada_vit.py:105-111).  Same decomposition as generation 1 (attn_bwd.cu):
//   work item = (sequence, head, 128-row KV tile); the CTA loops over the sequence's 128-row Q tiles
//     S^T  = K Q^T,  dP^T = V dO^T                     (SS MMA, M = kv, N = q)              -> TMEM
//     P^T  = exp2(S^T c - LSE)                         ("E" phase)  -> bf16 in its OWN 64 TMEM columns
//     dS^T = P^T o (dP^T - delta) scale                ("D" phase)  -> bf16, swizzled smem
//     dV  += P^T dO (TS),  dQ_i = dS K,  dK += dS^T Q  (dQ leaves through TMA reduce-adds, dK/dV stay in TMEM)
// What changed is the schedule.  Generation 1 ran the tensor pipe and the 8 softmax warps strictly in turn (2040 clk of
// MMAs, then ~2060 clk of softmax per q tile: profiles/r01_timeline_attn.txt) because P^T overwrote S^T in place, so the
// next S^T could not be issued before dV had consumed P^T, and because every softmax warp did every phase.
//   TMEM  S^T 128 | dP^T 128 (dQ aliases it) | dK HD | dV HD | P^T 64  = 512 columns at HD = 96
// * P^T in its own columns frees the S^T buffer as soon as the E phase has READ it: S^T(i+1) is issued right then.
// * The softmax warps are specialised: warps 0-3 ("A") run only the E phase (one thread = one kv row, all 128 q columns:
//   MUFU-bound), warps 4-7 ("B") run the D phase (reading P^T back from TMEM) and the dQ drain (FMA / shared-memory bound);
//   dV epilogue on A, dK epilogue on B.  E(i+1) on A overlaps D(i) + drain(i) on B and the dV / dQ / dK products of tile i.
// What remains serial is the chain  D(i) -> dQ(i) -> drain -> dP^T(i+1) -> D(i+1)  (dQ aliases the dP^T columns, and the D phase
// of tile i+1 overwrites the dS^T tile dQ(i) / dK(i) read): DESIGN.md section 7 lists the variants that were measured against it.
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"
#include "attn_bwd.cuh"

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

// Work-item descriptor fetched ONE ITEM AHEAD (every role walks the same list): a plain load at the top of the item puts a
// dependent global-memory round trip (descriptor -> addresses -> LSE / delta / TMA coordinates) into every item boundary
// (in-kernel timeline, profiles/r02_timeline_attn_bwd_boundary.txt: ~4000 clk per role and boundary).
__device__ __forceinline__ int4 ldg_int4_pinned(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// L2 prefetch of one TMA box (no shared-memory destination, no barrier): the next item's K / V tiles and its first Q / dO tile
// are HBM misses when the item starts (4500 clk from issue to landing at the boundary, ~1000 from L2).
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}

template <int HD>
struct Bwd2Cfg : BwdCfg<HD> {
  static constexpr int COL_S = 0, COL_DP = 128, COL_DQ = 128 /* aliases dP^T */, COL_DK = 256, COL_DV = 256 + (HD < 32 ? 32 : HD), COL_PT = 448;
  static_assert(COL_DV + HD <= COL_PT, "generation 2 needs 64 free TMEM columns for P^T (head_dim <= 96)");
  // K, V, NS x (Q, dO), dS^T, dQ staging, 2 x (LSE, delta), barriers; the dynamic shared memory is declared 1024-byte aligned
  static constexpr int SMEM_BYTES = BwdCfg<HD>::TILE_BYTES * (2 + 2 * BwdCfg<HD>::QDO_STAGES) + BwdCfg<HD>::DS_BYTES + BwdCfg<HD>::DQ_STAGE_BYTES + 2048 + 256;
};

template <int HD>
__global__ void __launch_bounds__(352, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const AttnBwdArgs a) {
  using Cfg = Bwd2Cfg<HD>;
  constexpr int NS = Cfg::QDO_STAGES;
  extern __shared__ __align__(1024) uint8_t smem2_raw[];
  uint8_t* smem = smem2_raw;
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::TILE_BYTES;
  uint8_t* sQ = sV + Cfg::TILE_BYTES;                  // [NS]
  uint8_t* sDO = sQ + NS * Cfg::TILE_BYTES;            // [NS]
  uint8_t* sDS = sDO + NS * Cfg::TILE_BYTES;           // dS^T
  uint8_t* sDQ = sDS + Cfg::DS_BYTES;                  // dQ staging slabs
  float* sLSE = reinterpret_cast<float*>(sDQ + Cfg::DQ_STAGE_BYTES);   // [2][128]
  float* sDelta = sLSE + 256;                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qdo_full = bars + 2;           // [NS]
  uint64_t* qdo_empty = qdo_full + NS;     // [NS]
  uint64_t* s_full = qdo_empty + NS;
  uint64_t* dp_full = s_full + 1;
  uint64_t* s_free = dp_full + 1;          // 128 arrivals (A): the E phase has read S^T(i)
  uint64_t* pa_ready = s_free + 1;         // 128 arrivals (A): P^T(i) is in TMEM
  uint64_t* pt_free = pa_ready + 1;        // 1 commit (dV(i) has consumed P^T(i)) + 128 arrivals (B has read it)
  uint64_t* p_ready = pt_free + 1;         // 128 arrivals (B): dS^T(i) is in shared memory
  uint64_t* dq_full = p_ready + 1;
  uint64_t* dq_drained = dq_full + 1;      // 128 arrivals (B)
  uint64_t* dkv_full = dq_drained + 1;
  uint64_t* dq_staged = dkv_full + 1;      // 128 arrivals (B): dQ_i sits in the smem slabs, ready for the bulk reductions
  uint64_t* stage_free = dq_staged + 1;    // the reductions of dQ_i have finished reading the slabs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);

  // warps 0-3: E phase (A), warps 4-7: D phase + dQ staging (B), both: epilogue; warp 8: TMA producer, warp 9: MMA issuer, warp 10: dQ reduction issuer
  constexpr int W_TMA = 8, W_MMA = 9, W_RED = 10;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == W_TMA && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) { printf("chadavit_b200: attn_bwd2 shared memory base is not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(s_free, 128); mbar_init(pa_ready, 128); mbar_init(pt_free, 129);
    mbar_init(p_ready, 128); mbar_init(dq_full, 1); mbar_init(dq_drained, 128); mbar_init(dkv_full, 1); mbar_init(dq_staged, 128);
    mbar_init(stage_free, 1);
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
      uint32_t it = 0, wi = 0;
      CB_TL_DECL(tlp);
      int4 wk_next = ldg_int4_pinned(a.work + blockIdx.x);
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const int4 wk = wk_next;
        if (w + (int)gridDim.x < a.n_work) wk_next = ldg_int4_pinned(a.work + w + gridDim.x); else wk_next = make_int4(0, 0, 0, 0);
        if (wk.z <= wk.y) continue;               // empty slot of the balanced schedule
        const int head = wk.w;
        const int nq = (wk.z - wk.y + 127) / 128;
        CB_TL(3, tlp, 1);
        mbar_wait(kv_empty, (wi & 1) ^ 1);
        ++wi;
        CB_TL(3, tlp, 2);
        mbar_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::NCH; ++c) {
          tma_load_2d(sK + c * Cfg::CHUNK_BYTES, &tmQKV, kv_full, a.D + head * HD + c * Cfg::CHUNK, wk.x);
          tma_load_2d(sV + c * Cfg::CHUNK_BYTES, &tmQKV, kv_full, 2 * a.D + head * HD + c * Cfg::CHUNK, wk.x);
        }
        for (int i = 0; i < nq; ++i, ++it) {
          const int s = it % NS; const uint32_t ph = (it / NS) & 1;
          mbar_wait(&qdo_empty[s], ph ^ 1);
          mbar_expect_tx(&qdo_full[s], 2 * Cfg::TILE_BYTES);
          const int row = wk.y + i * 128;
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c) {
            tma_load_2d(sQ + s * Cfg::TILE_BYTES + c * Cfg::CHUNK_BYTES, &tmQKV, &qdo_full[s], head * HD + c * Cfg::CHUNK, row);
            tma_load_2d(sDO + s * Cfg::TILE_BYTES + c * Cfg::CHUNK_BYTES, &tmDO, &qdo_full[s], head * HD + c * Cfg::CHUNK, row);
          }
          // The next item's K, V and first Q / dO tile -> L2, issued BEHIND this item's own first loads (in front of them the
          // prefetches delayed the first S^T of the item by ~4000 clk) and, for items of more than two tiles, one tile later still.
          if (i == (nq > 2 ? 1 : 0) && wk_next.z > wk_next.y) {
#pragma unroll
            for (int c = 0; c < Cfg::NCH; ++c) {
              tma_prefetch_l2_2d(&tmQKV, a.D + wk_next.w * HD + c * Cfg::CHUNK, wk_next.x);
              tma_prefetch_l2_2d(&tmQKV, 2 * a.D + wk_next.w * HD + c * Cfg::CHUNK, wk_next.x);
              tma_prefetch_l2_2d(&tmQKV, wk_next.w * HD + c * Cfg::CHUNK, wk_next.y);
              tma_prefetch_l2_2d(&tmDO, wk_next.w * HD + c * Cfg::CHUNK, wk_next.y);
            }
          }
        }
      }
    }
  } else if (warp == W_RED) {
    // ------------------------------------------------------------------ dQ reduction issuer (as in generation 1): dQ_i leaves
    // through TMA reduce-adds (fp32) from the 64B-swizzled slabs the softmax warps filled; issuing the bulk reductions of a
    // tile blocks the issuing thread for ~1200 clk, so a warp of its own does it.
    if (lane == 0) {
      uint32_t it = 0;
      int4 wk_next = ldg_int4_pinned(a.work + blockIdx.x);
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const int4 wk = wk_next;
        if (w + (int)gridDim.x < a.n_work) wk_next = ldg_int4_pinned(a.work + w + gridDim.x);
        if (wk.z <= wk.y) continue;
        const int head = wk.w;
        const int nq = (wk.z - wk.y + 127) / 128;
        for (int i = 0; i < nq; ++i, ++it) {
          mbar_wait(dq_staged, it & 1);
          const int q0 = wk.y + i * 128;
#pragma unroll 1
          for (int ww = 0; ww < 8; ++ww) {
            const int q4 = ww & 3, half = ww >> 2;
#pragma unroll
            for (int k = 0; k < Cfg::DQ_SLABS; ++k)
              if (k * 32 + half * 16 < HD)
                tma_reduce_add_2d(&tmDQ, sDQ + ww * (Cfg::DQ_SLABS * 2048) + k * 2048, head * HD + k * 32 + half * 16, q0 + q4 * 32);
          }
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(stage_free);
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (whole warp convergent, one elected lane
    // issues: descriptors stay in uniform registers and the UTCHMMAs are emitted back to back)
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);   // S^T, dP^T : both K-major
    constexpr uint32_t idesc_dv = umma_idesc_bf16(128, HD, false, true);    // dV (TS) / dK (SS): B MN-major
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, HD, true, true);     // dQ: A (dS^T) MN-major, B (K) MN-major
    const uint64_t k_kd = umma_smem_desc(smem_u32(sK), 16, Cfg::SBO, Cfg::SWZ), v_kd = umma_smem_desc(smem_u32(sV), 16, Cfg::SBO, Cfg::SWZ);
    const uint64_t q_kd0 = umma_smem_desc(smem_u32(sQ), 16, Cfg::SBO, Cfg::SWZ), do_kd0 = umma_smem_desc(smem_u32(sDO), 16, Cfg::SBO, Cfg::SWZ);
    const uint64_t do_md0 = umma_smem_desc(smem_u32(sDO), Cfg::CHUNK_BYTES, Cfg::SBO, Cfg::SWZ);   // dO / Q / K tiles read MN-major
    const uint64_t q_md0 = umma_smem_desc(smem_u32(sQ), Cfg::CHUNK_BYTES, Cfg::SBO, Cfg::SWZ);
    const uint64_t k_md = umma_smem_desc(smem_u32(sK), Cfg::CHUNK_BYTES, Cfg::SBO, Cfg::SWZ);
    const uint64_t ds_kd = umma_smem_desc(smem_u32(sDS), 16, 1024, 3), ds_md = umma_smem_desc(smem_u32(sDS), 16384, 1024, 3);
    uint32_t it = 0, wi = 0;
    CB_TL_DECL(tl);
    int4 wk_next = ldg_int4_pinned(a.work + blockIdx.x);
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      const int4 wk = wk_next;
      if (w + (int)gridDim.x < a.n_work) wk_next = ldg_int4_pinned(a.work + w + gridDim.x);
      if (wk.z <= wk.y) continue;
      const int nq = (wk.z - wk.y + 127) / 128;
      CB_TL(0, tl, 7);
      mbar_wait(kv_full, wi & 1);
      ++wi;
      CB_TL(0, tl, 8);
      auto issue_s = [&](uint32_t itx) {   // S^T = K Q^T of iteration itx: its Q tile has landed and the E phase has read S^T(itx-1)
        const int sx = itx % NS;
        const uint64_t q_kd = umma_desc_add(q_kd0, sx * Cfg::TILE_BYTES);
        mbar_wait(&qdo_full[sx], (itx / NS) & 1);
        if (itx > 0) mbar_wait(s_free, (itx - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const int c = (kk * 16) / Cfg::CHUNK, off = ((kk * 16) % Cfg::CHUNK) * 2;
            umma_ss(tmem_base + Cfg::COL_S, umma_desc_add(k_kd, c * Cfg::CHUNK_BYTES + off), umma_desc_add(q_kd, c * Cfg::CHUNK_BYTES + off), idesc_s, kk > 0 ? 1u : 0u);
          }
          tc_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](uint32_t itx) {  // dP^T = V dO^T of iteration itx; its TMEM region held dQ of iteration itx-1
        const int sx = itx % NS;
        const uint64_t do_kd = umma_desc_add(do_kd0, sx * Cfg::TILE_BYTES);
        if (itx > 0) { mbar_wait(dq_drained, (itx - 1) & 1); tc_fence_after(); }
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const int c = (kk * 16) / Cfg::CHUNK, off = ((kk * 16) % Cfg::CHUNK) * 2;
            umma_ss(tmem_base + Cfg::COL_DP, umma_desc_add(v_kd, c * Cfg::CHUNK_BYTES + off), umma_desc_add(do_kd, c * Cfg::CHUNK_BYTES + off), idesc_s, kk > 0 ? 1u : 0u);
          }
          tc_commit(dp_full);
        }
        __syncwarp();
      };
      issue_s(it);
      CB_TL(0, tl, 9);
      issue_dp(it);
      for (int i = 0; i < nq; ++i, ++it) {
        const int s = it % NS;
        const uint64_t q_md = umma_desc_add(q_md0, s * Cfg::TILE_BYTES), do_md = umma_desc_add(do_md0, s * Cfg::TILE_BYTES);
        CB_TL(0, tl, 2);
        if (i + 1 < nq) issue_s(it + 1);   // as soon as the E phase of tile i has read S^T(i)
        CB_TL(0, tl, 3);
        mbar_wait(pa_ready, it & 1);
        tc_fence_after();
        CB_TL(0, tl, 4);
        if (elect_one()) {
          // dV += P^T dO (A = P^T in its own TMEM columns; B = dO tile read MN-major: N = HD, K = q)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(tmem_base + Cfg::COL_DV, tmem_base + Cfg::COL_PT + kk * 8, umma_desc_add(do_md, kk * 16 * Cfg::CHUNK * 2), idesc_dv,
                    (i > 0 || kk > 0) ? 1u : 0u);
          tc_commit(pt_free);
        }
        __syncwarp();
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        CB_TL(0, tl, 6);
        if (elect_one()) {
          // dQ_i = dS K: A = dS^T smem read MN-major (M = q: 2 blocks of 64, LBO 16 KB; K = kv), B = K tile MN-major
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tmem_base + Cfg::COL_DQ, umma_desc_add(ds_md, kk * 2048), umma_desc_add(k_md, kk * 16 * Cfg::CHUNK * 2), idesc_dq, kk > 0 ? 1u : 0u);
          tc_commit(dq_full);
          // dK += dS^T Q   (A = dS^T smem K-major: two [128x64] sub-tiles; B = Q tile MN-major)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tmem_base + Cfg::COL_DK, umma_desc_add(ds_kd, (kk >> 2) * 16384 + (kk & 3) * 32), umma_desc_add(q_md, kk * 16 * Cfg::CHUNK * 2), idesc_dv,
                    (i > 0 || kk > 0) ? 1u : 0u);
          tc_commit(&qdo_empty[s]);
        }
        __syncwarp();
        // dP^T(i+1) behind dQ(i) / dK(i) in the in-order pipe: the D phase it releases overwrites the dS^T tile they read
        if (i + 1 < nq) issue_dp(it + 1);  // needs dQ_i out of the dP^T columns (B drains it right after the D phase)
        CB_TL(0, tl, 5);
      }
      if (elect_one()) { tc_commit(dkv_full); tc_commit(kv_empty); }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax-backward warps.  One thread = one kv row of the
    // tile (TMEM lane); A (warps 0-3): p = exp2(s c - lse) for all 128 q columns -> P^T (bf16) in TMEM;  B (warps 4-7):
    // dS = p (dP scale - delta scale) -> dS^T (bf16, swizzled smem), then the dQ drain.  Epilogue: dV (A) / dK (B).
    const bool is_a = warp < 4;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;               // kv row of this thread inside the tile; also q row when draining dQ
    const int tid128 = q4 * 32 + lane;          // 0..127 within the role
    const uint32_t lane_addr = tmem_base + (uint32_t(q4 * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    uint32_t it = 0, wi = 0;
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4) && lane == 0;
    const int tl_role = warp == 0 ? 1 : 2;
    int4 wk_next = ldg_int4_pinned(a.work + blockIdx.x);
    float stage_pref = 0.f;      // LSE / delta of the NEXT item's first q tile, requested before this item's epilogue
    bool have_pref = false;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      const int4 wk = wk_next;
      if (w + (int)gridDim.x < a.n_work) wk_next = ldg_int4_pinned(a.work + w + gridDim.x); else wk_next = make_int4(0, 0, 0, 0);
      if (wk.z <= wk.y) continue;
      const int head = wk.w;
      const int nq = (wk.z - wk.y + 127) / 128;
      const bool kv_ok = wk.x + r < wk.z;
      // LSE (A) / delta (B) of a q tile: loaded into a register one iteration ahead, published to the role's double-buffered
      // smem row at the start of the iteration (buffer it & 1: its previous readers, tile it-2, are behind the role's named
      // barrier of tile it-1).
      // The RAW value is kept in the register and scaled only when it is published one tile later: an arithmetic instruction on
      // the loaded value here would stall this (in-order) warp for the whole global-memory latency (timeline: ~1000 clk/tile).
      const float* stage_src = (is_a ? a.lse : a.delta) + (long)head * a.T;
      const float stage_mul = is_a ? LOG2E : a.scale;    // LSE in the exp2 domain / delta pre-multiplied by the softmax scale
      const float stage_oob = is_a ? INFINITY : 0.f;     // +inf -> p = 0 for q rows past the sequence
      auto stage_load = [&](int i_) -> float {
        const int t = wk.y + i_ * 128 + tid128;
        return ldg_f32_pinned(stage_src + (t < wk.z ? t : wk.y));
      };
      auto stage_fix = [&](float v, int i_) -> float { return (wk.y + i_ * 128 + tid128 < wk.z) ? v * stage_mul : stage_oob; };
      float stage_val = have_pref ? stage_pref : stage_load(0);
      if (is_a) {
        for (int i = 0; i < nq; ++i, ++it) {
          float* lse_b = sLSE + (it & 1) * 128;
          if (tl_on) CB_TL(tl_role, tl, 7);
          lse_b[tid128] = stage_fix(stage_val, i);
          if (tl_on) CB_TL(tl_role, tl, 8);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (tl_on) CB_TL(tl_role, tl, 9);
          if (i + 1 < nq) stage_val = stage_load(i + 1);
          if (tl_on) CB_TL(tl_role, tl, 1);
          mbar_wait(s_full, it & 1);
          tc_fence_after();
          if (tl_on) CB_TL(tl_role, tl, 2);
          // ---- E phase: 128 q columns in eight 16-column chunks; the packed bf16 pairs of 64 columns go to TMEM with one store
          uint32_t sr[2][16];
          tmem_ld16(lane_addr + Cfg::COL_S, sr[0]);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c8 = hh * 4 + j;
              tmem_ld_wait();
              if (c8 < 7) tmem_ld16(lane_addr + Cfg::COL_S + (c8 + 1) * 16, sr[(c8 + 1) & 1]);
              else { tc_fence_before(); mbar_arrive(s_free); }   // every S^T(i) column has been read: S^T(i+1) may be issued
              const uint32_t (&sv)[16] = sr[c8 & 1];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(lse_b + c8 * 16 + e);
                const float p0 = fast_exp2_b(fmaf(__uint_as_float(sv[e]), a.scale_log2, -l4.x)), p1 = fast_exp2_b(fmaf(__uint_as_float(sv[e + 1]), a.scale_log2, -l4.y));
                const float p2 = fast_exp2_b(fmaf(__uint_as_float(sv[e + 2]), a.scale_log2, -l4.z)), p3 = fast_exp2_b(fmaf(__uint_as_float(sv[e + 3]), a.scale_log2, -l4.w));
                pk[j * 8 + (e >> 1)] = pack_bf16(p0, p1);
                pk[j * 8 + (e >> 1) + 1] = pack_bf16(p2, p3);
              }
            }
            if (hh == 0 && it > 0) { mbar_wait(pt_free, (it - 1) & 1); tc_fence_after(); }   // dV(i-1) and B have consumed P^T(i-1)
            tmem_st32(lane_addr + Cfg::COL_PT + hh * 32, pk);
          }
          if (tl_on) CB_TL(tl_role, tl, 3);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(pa_ready);
          if (tl_on) CB_TL(tl_role, tl, 4);
        }
      } else {
        const uint32_t keep = kv_ok ? 0xffffffffu : 0u;
        for (int i = 0; i < nq; ++i, ++it) {
          float* del_b = sDelta + (it & 1) * 128;
          del_b[tid128] = stage_fix(stage_val, i);
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (i + 1 < nq) stage_val = stage_load(i + 1);
          if (tl_on) CB_TL(tl_role, tl, 1);
          mbar_wait(dp_full, it & 1);
          mbar_wait(pa_ready, it & 1);
          tc_fence_after();
          if (tl_on) CB_TL(tl_role, tl, 2);
          // ---- D phase: dS = p (dP scale - delta scale), bf16, into the swizzled dS^T tile.  kv rows past the sequence end must
          // contribute nothing to dQ = dS K: their packed dS words are cleared (their P rows only feed dV / dK rows that are
          // never stored).  dP^T(i) being complete implies dK(i-1) / dQ(i-1), the readers of the dS^T tile, have retired.
          uint32_t dpr[2][16];
          tmem_ld16(lane_addr + Cfg::COL_DP, dpr[0]);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t pk[32];
            tmem_ld32(lane_addr + Cfg::COL_PT + hh * 32, pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c8 = hh * 4 + j;               // 16-column chunk index within the 128 q columns
              tmem_ld_wait();
              if (hh == 1 && j == 0) { tc_fence_before(); mbar_arrive(pt_free); }   // both halves of P^T(i) are in registers
              if (c8 < 7) tmem_ld16(lane_addr + Cfg::COL_DP + (c8 + 1) * 16, dpr[(c8 + 1) & 1]);
              const uint32_t (&dv)[16] = dpr[c8 & 1];
              uint32_t dsp[8];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 d4 = *reinterpret_cast<const float4*>(del_b + c8 * 16 + e);
                const uint32_t w0 = pk[j * 8 + (e >> 1)], w1 = pk[j * 8 + (e >> 1) + 1];
                const float e0 = __uint_as_float(w0 << 16) * fmaf(__uint_as_float(dv[e]), a.scale, -d4.x);
                const float e1 = __uint_as_float(w0 & 0xffff0000u) * fmaf(__uint_as_float(dv[e + 1]), a.scale, -d4.y);
                const float e2 = __uint_as_float(w1 << 16) * fmaf(__uint_as_float(dv[e + 2]), a.scale, -d4.z);
                const float e3 = __uint_as_float(w1 & 0xffff0000u) * fmaf(__uint_as_float(dv[e + 3]), a.scale, -d4.w);
                dsp[e >> 1] = pack_bf16(e0, e1) & keep;
                dsp[(e >> 1) + 1] = pack_bf16(e2, e3) & keep;
              }
              // dS^T row r, q columns [16 c8, 16 c8 + 16) -> sub-tile (c8 >> 2), 16B chunks 2 (c8 & 3), +1, 128B swizzle
              uint8_t* base = sDS + (c8 >> 2) * 16384 + r * 128;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int chunk = ((c8 & 3) * 2 + k) ^ (r & 7);
                *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(dsp[4 * k], dsp[4 * k + 1], dsp[4 * k + 2], dsp[4 * k + 3]);
              }
            }
          }
          fence_proxy_async();   // generic-proxy smem writes (dS^T) -> visible to the tensor-core (async) proxy
          mbar_arrive(p_ready);
          if (tl_on) CB_TL(tl_role, tl, 3);
          // ---- dQ_i (rows = q): TMEM -> 64B-swizzled smem slabs -> (warp 10) TMA reduce-add into the fp32 accumulator.  Rows past
          // the sequence end carry exact zeros (P = 0 there).  Slab set h*4 + q4 = q rows 32 q4.., columns k*32 + h*16..
          mbar_wait(dq_full, it & 1);
          tc_fence_after();
          if (tl_on) CB_TL(tl_role, tl, 5);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint8_t* my = sDQ + (h * 4 + q4) * (Cfg::DQ_SLABS * 2048);
            uint32_t o[Cfg::DQ_SLABS][16];
#pragma unroll
            for (int k = 0; k < Cfg::DQ_SLABS; ++k)
              if (k * 32 + h * 16 < HD) tmem_ld16(lane_addr + Cfg::COL_DQ + k * 32 + h * 16, o[k]);
            tmem_ld_wait();
            if (h == 1) { tc_fence_before(); mbar_arrive(dq_drained); }   // the dQ / dP^T columns may be overwritten from here on
            if (h == 0 && it > 0) mbar_wait(stage_free, (it - 1) & 1);    // the previous tile's reductions have read the slabs
#pragma unroll
            for (int k = 0; k < Cfg::DQ_SLABS; ++k)
              if (k * 32 + h * 16 < HD) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(my + k * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                      make_uint4(o[k][4 * j], o[k][4 * j + 1], o[k][4 * j + 2], o[k][4 * j + 3]);
              }
          }
          fence_proxy_async();
          mbar_arrive(dq_staged);                           // warp 10 issues the bulk reductions
          if (tl_on) CB_TL(tl_role, tl, 6);
        }
      }
      // ---- dV (A warps) / dK (B warps) of this kv tile -> bf16 into dqkv.  The roles are not arbitrary: the next item's first dV
      // product (which overwrites dV) is issued on pa_ready, i.e. after the A warps have left this epilogue, and its first dK
      // product on p_ready, after the B warps have.
      have_pref = wk_next.z > wk_next.y;
      if (have_pref) {   // first tile of the next item: in flight during this item's epilogue
        const int t = wk_next.y + tid128;
        stage_pref = ldg_f32_pinned((is_a ? a.lse : a.delta) + (long)wk_next.w * a.T + (t < wk_next.z ? t : wk_next.y));
      }
      if (tl_on) CB_TL(tl_role, tl, 10);
      mbar_wait(dkv_full, wi & 1);
      ++wi;
      tc_fence_after();
      if (tl_on) CB_TL(tl_role, tl, 11);
      {
        __nv_bfloat16* dst = a.dqkv + (long)(wk.x + r) * (3 * a.D) + (is_a ? 2 * a.D : a.D) + head * HD;
        const uint32_t col = is_a ? Cfg::COL_DV : Cfg::COL_DK;
#pragma unroll
        for (int c = 0; c < HD; c += 16) {
          uint32_t o[16];
          tmem_ld16(lane_addr + col + c, o);
          tmem_ld_wait();
          if (kv_ok) {
            // one FULL 32-byte sector per store (rows, the K / V column blocks and the head offsets are multiples of 32 bytes when
            // HD % 16 == 0): as two 16-byte halves every sector was written twice, partially — the 128 x HD tile took ~5000 clk to
            // leave the SM and held up the LSU for the next item's first loads (profiles/r02_timeline_attn_bwd_boundary.txt)
            stg256(dst + c, pack_bf16(__uint_as_float(o[0]), __uint_as_float(o[1])), pack_bf16(__uint_as_float(o[2]), __uint_as_float(o[3])),
                   pack_bf16(__uint_as_float(o[4]), __uint_as_float(o[5])), pack_bf16(__uint_as_float(o[6]), __uint_as_float(o[7])),
                   pack_bf16(__uint_as_float(o[8]), __uint_as_float(o[9])), pack_bf16(__uint_as_float(o[10]), __uint_as_float(o[11])),
                   pack_bf16(__uint_as_float(o[12]), __uint_as_float(o[13])), pack_bf16(__uint_as_float(o[14]), __uint_as_float(o[15])));
          }
        }
      }
      tc_fence_before();
      if (tl_on) CB_TL(tl_role, tl, 12);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

template <int HD>
static int launch_bwd2(const void* qkv, const void* dO, const AttnBwdArgs& a, cudaStream_t stream) {
  using Cfg = Bwd2Cfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tq, td, tdq;
  {
    uint64_t dims[2] = {(uint64_t)(3 * a.D), (uint64_t)a.T};
    uint64_t strides[1] = {(uint64_t)(3 * a.D) * 2};
    uint32_t box[2] = {(uint32_t)Cfg::CHUNK, 128};
    if (make_tmap(&tq, qkv, 2, dims, strides, box, Cfg::SWZ)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.D, (uint64_t)a.T};
    uint64_t strides[1] = {(uint64_t)a.D * 2};
    uint32_t box[2] = {(uint32_t)Cfg::CHUNK, 128};
    if (make_tmap(&td, dO, 2, dims, strides, box, Cfg::SWZ)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.D, (uint64_t)a.T};
    uint64_t strides[1] = {(uint64_t)a.D * 4};
    uint32_t box[2] = {16, 32};
    if (make_tmap(&tdq, a.dq_acc, 2, dims, strides, box, 2, 4)) return 1;
  }
  const int grid = a.n_work < num_sms() ? a.n_work : num_sms();
  attn_bwd2_kernel<HD><<<grid, 352, Cfg::SMEM_BYTES, stream>>>(tq, td, tdq, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

int attn_bwd2_launch(int hd, const void* qkv, const void* dO, const AttnBwdArgs& a, cudaStream_t stream) {
  switch (hd) {
    case 16: return launch_bwd2<16>(qkv, dO, a, stream);
    case 32: return launch_bwd2<32>(qkv, dO, a, stream);
    case 64: return launch_bwd2<64>(qkv, dO, a, stream);
    case 96: return launch_bwd2<96>(qkv, dO, a, stream);
    default: set_error("attn_bwd2: unsupported head_dim %d", hd); return 1;
  }
}

}  // namespace cb

#ifdef CB_TIMELINE
extern "C" int cb_debug_timeline_bwd2(void* dst) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif
