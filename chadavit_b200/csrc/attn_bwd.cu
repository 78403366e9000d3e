// Varlen multi-head self-attention backward for sm_100a (autograd of the SDPA inside nn.MultiheadAttention,
// chada_vit.py:105-111).  One pass, flash-style recompute:
//   work item = (sequence, head, 128-row KV tile); the CTA loops over the sequence's 128-row Q tiles
//     S^T  = K Q^T            (SS MMA, M = kv, N = q)            -> TMEM
//     dP^T = V dO^T           (SS MMA)                           -> TMEM
//     P^T  = exp2(S^T*c - LSE),  dS^T = scale * P^T o (dP^T - delta)      (128 threads, 1 thread = 1 kv row)
//     dV  += P^T dO           (TS MMA, P^T bf16 in TMEM, dO consumed MN-major from its TMA tile)
//     dK  += dS^T Q           (SS MMA, dS^T bf16 in swizzled smem, Q MN-major)
//     dQ_i = dS K             (SS MMA: the SAME dS^T smem tile read MN-major, K MN-major) -> TMEM -> fp32 atomics
//   dK/dV stay in TMEM for the whole work item and are written once as bf16 into dqkv[:, D:3D].
// delta = rowsum(dO o O) and the fp32->bf16 conversion of the dQ accumulator are small row-wise kernels below.
#include <stdlib.h>

#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"
#include "attn_bwd.cuh"

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

template <int HD>
__global__ void __launch_bounds__(352, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdArgs a) {
  using Cfg = BwdCfg<HD>;
  constexpr int NS = Cfg::QDO_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::TILE_BYTES;
  uint8_t* sQ = sV + Cfg::TILE_BYTES;                  // [NS]
  uint8_t* sDO = sQ + NS * Cfg::TILE_BYTES;            // [NS]
  uint8_t* sDS = sDO + NS * Cfg::TILE_BYTES;           // dS^T
  uint8_t* sDQ = sDS + Cfg::DS_BYTES;                  // dQ staging slabs
  float* sLSE = reinterpret_cast<float*>(sDQ + Cfg::DQ_STAGE_BYTES);
  float* sDelta = sLSE + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 128);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qdo_full = bars + 2;           // [NS]
  uint64_t* qdo_empty = qdo_full + NS;     // [NS]
  uint64_t* s_full = qdo_empty + NS;
  uint64_t* dp_full = s_full + 1;
  uint64_t* p_ready = dp_full + 1;         // 128 arrivals
  uint64_t* dq_full = p_ready + 1;
  uint64_t* dq_drained = dq_full + 1;      // 128 arrivals
  uint64_t* dkv_full = dq_drained + 1;
  uint64_t* dq_staged = dkv_full + 1;      // 256 arrivals: dQ_i sits in the smem slabs, ready for the bulk reductions
  uint64_t* stage_free = dq_staged + 1;    // the reductions of dQ_i have finished reading the slabs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);

  // warps 0-7: softmax-backward / dQ staging / epilogue, warp 8: TMA producer, warp 9: MMA issuer, warp 10: dQ reduction issuer
  constexpr int W_TMA = 8, W_MMA = 9, W_RED = 10;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(p_ready, 256); mbar_init(dq_full, 1); mbar_init(dq_drained, 256);
    mbar_init(dkv_full, 1); mbar_init(dq_staged, 256); mbar_init(stage_free, 1);
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
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const int4 wk = a.work[w];
        if (wk.z <= wk.y) continue;               // empty slot of the balanced schedule
        const int head = wk.w;
        const int nq = (wk.z - wk.y + 127) / 128;
        mbar_wait(kv_empty, (wi & 1) ^ 1);
        ++wi;
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
        }
      }
    }
  } else if (warp == W_RED) {
    // ------------------------------------------------------------------ dQ reduction issuer
    // dQ_i (rows = q) leaves through TMA reduce-add (fp32) from the 64B-swizzled slabs the softmax warps filled: whole
    // 64-byte row segments instead of per-thread REDs (which scatter 32 rows per instruction).  Issuing the 24 bulk reductions
    // of a tile takes ~1200 clk (the issuing thread blocks on the TMA queue), so a warp of its own does it: the softmax warps
    // and the tensor pipe move on to the next q tile meanwhile.
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const int4 wk = a.work[w];
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
    {
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
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
        const int4 wk = a.work[w];
        if (wk.z <= wk.y) continue;
        const int nq = (wk.z - wk.y + 127) / 128;
        mbar_wait(kv_full, wi & 1);
        ++wi;
        auto issue_s = [&](uint32_t itx) {   // S^T = K Q^T of iteration itx (its Q tile must have landed)
          const int sx = itx % NS;
          const uint64_t q_kd = umma_desc_add(q_kd0, sx * Cfg::TILE_BYTES);
          mbar_wait(&qdo_full[sx], (itx / NS) & 1);
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
        issue_dp(it);
        for (int i = 0; i < nq; ++i, ++it) {
          const int s = it % NS;
          const uint64_t q_md = umma_desc_add(q_md0, s * Cfg::TILE_BYTES), do_md = umma_desc_add(do_md0, s * Cfg::TILE_BYTES);
          CB_TL(0, tl, 3);
          mbar_wait(p_ready, it & 1);
          tc_fence_after();
          CB_TL(0, tl, 4);
          if (elect_one()) {
            // dQ_i = dS K first: its read-out by the softmax warps then runs under dV / dK / the next S^T, and the next dP^T
            // (same TMEM columns) can follow without a bubble.  A = dS^T smem read MN-major (M = q: 2 blocks of 64, LBO
            // 16 KB; K = kv), B = K tile MN-major.
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(tmem_base + Cfg::COL_DP, umma_desc_add(ds_md, kk * 2048), umma_desc_add(k_md, kk * 16 * Cfg::CHUNK * 2), idesc_dq, kk > 0 ? 1u : 0u);
            tc_commit(dq_full);
            // dV += P^T dO   (A = P^T in TMEM; B = dO tile read MN-major: N = HD, K = q)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts(tmem_base + Cfg::COL_DV, tmem_base + Cfg::COL_S + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8), umma_desc_add(do_md, kk * 16 * Cfg::CHUNK * 2), idesc_dv,
                      (i > 0 || kk > 0) ? 1u : 0u);
            // dK += dS^T Q   (A = dS^T smem K-major: two [128x64] sub-tiles; B = Q tile MN-major)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(tmem_base + Cfg::COL_DK, umma_desc_add(ds_kd, (kk >> 2) * 16384 + (kk & 3) * 32), umma_desc_add(q_md, kk * 16 * Cfg::CHUNK * 2), idesc_dv,
                      (i > 0 || kk > 0) ? 1u : 0u);
            tc_commit(&qdo_empty[s]);
          }
          __syncwarp();
          // next q tile: S^T right behind (the in-order pipe has retired dV, the last reader of P^T, by then), then dP^T as soon
          // as dQ_i has been read out of its columns
          if (i + 1 < nq) { issue_s(it + 1); issue_dp(it + 1); }
          CB_TL(0, tl, 5);
        }
        if (elect_one()) { tc_commit(dkv_full); tc_commit(kv_empty); }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax-backward / dQ drain / dK,dV epilogue
    // 8 warps: two per TMEM lane quarter; `half` selects which 64 of the 128 q columns (and which 16-column chunks of the
    // dQ / dK / dV accumulators) this warp handles — the backward softmax is element-wise, so no row reduction is split.
    const int q4 = warp & 3;
    const int half = warp >> 2;
    const int r = q4 * 32 + lane;               // kv row of this thread inside the tile; also q row when draining dQ
    const int tid256 = warp * 32 + lane;        // 0..255
    const uint32_t lane_addr = tmem_base + (uint32_t(q4 * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    uint32_t it = 0, wi = 0;
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4) && lane == 0;
    const int tl_role = warp == 0 ? 1 : 2;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x) {
      const int4 wk = a.work[w];
      if (wk.z <= wk.y) continue;
      const int head = wk.w;
      const int nq = (wk.z - wk.y + 127) / 128;
      const bool kv_ok = wk.x + r < wk.z;
      // LSE (threads 0-127) / delta (threads 128-255) of a q tile: loaded into a register one iteration ahead (the global
      // latency hides under the wait for the dV/dK/dQ MMAs), published to smem at the start of the iteration.
      const float* stage_src = (tid256 < 128 ? a.lse : a.delta) + (long)head * a.T;
      // the RAW value stays in the register and is scaled when it is published: arithmetic on it here would stall the warp for the
      // whole global-memory latency
      const float stage_mul = tid256 < 128 ? LOG2E : a.scale;   // LSE in the exp2 domain / delta pre-multiplied by the softmax scale
      const float stage_oob = tid256 < 128 ? INFINITY : 0.f;    // +inf -> p = 0 for q rows past the sequence
      auto stage_load = [&](int i_) -> float {
        const int t = wk.y + i_ * 128 + (tid256 & 127);
        return ldg_f32_pinned(stage_src + (t < wk.z ? t : wk.y));
      };
      auto stage_fix = [&](float v, int i_) -> float { return (wk.y + i_ * 128 + (tid256 & 127) < wk.z) ? v * stage_mul : stage_oob; };
      float stage_val = stage_load(0);
      for (int i = 0; i < nq; ++i, ++it) {
        const int q0 = wk.y + i * 128;
        if (tid256 < 128) sLSE[tid256] = stage_fix(stage_val, i); else sDelta[tid256 - 128] = stage_fix(stage_val, i);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tl_on) CB_TL(tl_role, tl, 1);
        mbar_wait(s_full, it & 1);
        if (tl_on) CB_TL(tl_role, tl, 2);
        mbar_wait(dp_full, it & 1);
        tc_fence_after();
        if (tl_on) CB_TL(tl_role, tl, 3);
        // This warp's 64 q columns in four 16-column chunks (ptxas re-uses one register set: load -> math per chunk).
        // Per element: p = exp2(s c - lse) (FFMA + MUFU), dS = p (dP scale - delta scale) (FFMA + FMUL), two bf16 packs per
        // pair; LSE / delta come as broadcast 16-byte smem loads (4 q columns each).  kv rows past the sequence end must
        // contribute nothing to dQ = dS K: their packed dS words are cleared with one AND per pair (their P rows only feed
        // dV / dK rows that are never stored).
        {
          const uint32_t keep = kv_ok ? 0xffffffffu : 0u;
          uint32_t sr[2][16], dpr[2][16];
          tmem_ld16(lane_addr + Cfg::COL_S + half * 64, sr[0]);
          tmem_ld16(lane_addr + Cfg::COL_DP + half * 64, dpr[0]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c8 = half * 4 + j;               // 16-column chunk index within the 128 q columns
            tmem_ld_wait();
            if (j < 3) {
              tmem_ld16(lane_addr + Cfg::COL_S + (c8 + 1) * 16, sr[(j + 1) & 1]);
              tmem_ld16(lane_addr + Cfg::COL_DP + (c8 + 1) * 16, dpr[(j + 1) & 1]);
            }
            const uint32_t (&sv)[16] = sr[j & 1];
            const uint32_t (&dv)[16] = dpr[j & 1];
            uint32_t pp[8], dsp[8];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(sLSE + c8 * 16 + e), d4 = *reinterpret_cast<const float4*>(sDelta + c8 * 16 + e);
              const float p0 = fast_exp2_b(fmaf(__uint_as_float(sv[e]), a.scale_log2, -l4.x)), p1 = fast_exp2_b(fmaf(__uint_as_float(sv[e + 1]), a.scale_log2, -l4.y));
              const float p2 = fast_exp2_b(fmaf(__uint_as_float(sv[e + 2]), a.scale_log2, -l4.z)), p3 = fast_exp2_b(fmaf(__uint_as_float(sv[e + 3]), a.scale_log2, -l4.w));
              const float e0 = p0 * fmaf(__uint_as_float(dv[e]), a.scale, -d4.x), e1 = p1 * fmaf(__uint_as_float(dv[e + 1]), a.scale, -d4.y);
              const float e2 = p2 * fmaf(__uint_as_float(dv[e + 2]), a.scale, -d4.z), e3 = p3 * fmaf(__uint_as_float(dv[e + 3]), a.scale, -d4.w);
              pp[e >> 1] = pack_bf16(p0, p1); pp[(e >> 1) + 1] = pack_bf16(p2, p3);
              dsp[e >> 1] = pack_bf16(e0, e1) & keep; dsp[(e >> 1) + 1] = pack_bf16(e2, e3) & keep;
            }
            // P^T (bf16 pairs) stays inside this warp's own half of the S region (columns already read): q cols 0-63 -> TMEM
            // cols 0..31, 64-127 -> 64..95
            tmem_st8(lane_addr + Cfg::COL_S + half * 64 + j * 8, pp);
            // dS^T row r, q columns [16 c8, 16 c8 + 16) -> sub-tile (c8 >> 2), 16B chunks 2 (c8 & 3), +1, 128B swizzle
            uint8_t* base = sDS + (c8 >> 2) * 16384 + r * 128;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int chunk = ((c8 & 3) * 2 + k) ^ (r & 7);
              *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(dsp[4 * k], dsp[4 * k + 1], dsp[4 * k + 2], dsp[4 * k + 3]);
            }
          }
        }
        if (tl_on) CB_TL(tl_role, tl, 31);
        tmem_st_wait();
        if (tl_on) CB_TL(tl_role, tl, 32);
        fence_proxy_async();   // generic-proxy smem writes (dS^T) -> visible to the tensor-core (async) proxy
        if (tl_on) CB_TL(tl_role, tl, 33);
        tc_fence_before();
        mbar_arrive(p_ready);
        if (tl_on) CB_TL(tl_role, tl, 4);
        // ---- drain dQ_i (rows = q): TMEM -> 64B-swizzled smem slab -> TMA reduce-add (fp32) into the dQ accumulator.
        // Per-thread REDs would scatter 32 rows per instruction (ncu: the kernel was bound by L2 atomic transactions);
        // the bulk reduction moves whole 64-byte row segments.  Rows past the sequence end carry exact zeros (P = 0 there).
        if (i + 1 < nq) stage_val = stage_load(i + 1);
        mbar_wait(dq_full, it & 1);
        tc_fence_after();
        if (tl_on) CB_TL(tl_role, tl, 5);
        {
          uint8_t* my = sDQ + warp * (Cfg::DQ_SLABS * 2048);
          uint32_t o[Cfg::DQ_SLABS][16];
#pragma unroll
          for (int k = 0; k < Cfg::DQ_SLABS; ++k)
            if (k * 32 + half * 16 < HD) tmem_ld16(lane_addr + Cfg::COL_DP + k * 32 + half * 16, o[k]);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(dq_drained);                          // the dQ / dP^T columns may be overwritten from here on
          if (tl_on) CB_TL(tl_role, tl, 21);
          if (it > 0) mbar_wait(stage_free, (it - 1) & 1);  // the previous tile's reductions have finished reading the slabs
#pragma unroll
          for (int k = 0; k < Cfg::DQ_SLABS; ++k)
            if (k * 32 + half * 16 < HD) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(my + k * 2048 + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_uint4(o[k][4 * j], o[k][4 * j + 1], o[k][4 * j + 2], o[k][4 * j + 3]);
            }
          fence_proxy_async();
          mbar_arrive(dq_staged);                           // warp 10 issues the bulk reductions (issuing them blocks ~400 clk each)
          if (tl_on) CB_TL(tl_role, tl, 6);
        }
      }
      // ---- dK, dV of this kv tile -> bf16 into dqkv
      mbar_wait(dkv_full, wi & 1);
      ++wi;
      tc_fence_after();
      {
        __nv_bfloat16* dk_dst = a.dqkv + (long)(wk.x + r) * (3 * a.D) + a.D + head * HD;
        __nv_bfloat16* dv_dst = dk_dst + a.D;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          __nv_bfloat16* dst = which ? dv_dst : dk_dst;
          const uint32_t col = which ? Cfg::COL_DV : Cfg::COL_DK;
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) {
            const int c = c0 + half * 16;
            if (c >= HD) continue;
            uint32_t o[16];
            tmem_ld16(lane_addr + col + c, o);
            tmem_ld_wait();
            if (kv_ok) {   // one full 32-byte sector per store (see attn_bwd2.cu)
              stg256(dst + c, pack_bf16(__uint_as_float(o[0]), __uint_as_float(o[1])), pack_bf16(__uint_as_float(o[2]), __uint_as_float(o[3])),
                     pack_bf16(__uint_as_float(o[4]), __uint_as_float(o[5])), pack_bf16(__uint_as_float(o[6]), __uint_as_float(o[7])),
                     pack_bf16(__uint_as_float(o[8]), __uint_as_float(o[9])), pack_bf16(__uint_as_float(o[10]), __uint_as_float(o[11])),
                     pack_bf16(__uint_as_float(o[12]), __uint_as_float(o[13])), pack_bf16(__uint_as_float(o[14]), __uint_as_float(o[15])));
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

// delta[h, t] = sum_c dO[t, h*d + c] * O[t, h*d + c].  Four lanes per (row, head): lane l takes the 16-byte chunks l, l + 4, ... of
// the head's d columns of both tensors (all loads of a thread requested before the first use), two shuffles fold the quad.  (One
// warp per row with 4-byte loads and a full warp reduction per head ran at 2.3 TB/s: 45 us for the two packed global crops.)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                                                         float* __restrict__ delta, int T, int D, int H) {
  const int d = D / H, nch = d / 8;                 // 16-byte chunks per head (d is a multiple of 8: checked by the caller)
  const int l = threadIdx.x & 3, lane = threadIdx.x & 31;
  const long npairs = (long)T * H;
  const long warp0 = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (long pw = warp0 * 8; pw < npairs; pw += nwarps * 8) {   // warp-uniform trip count: a warp = 8 (row, head) pairs per trip
    const long p = pw + (lane >> 2);
    const bool ok = p < npairs;
    const long t = ok ? p / H : 0; const int h = ok ? (int)(p - t * H) : 0;
    const __nv_bfloat16* a = dO + t * D + h * d; const __nv_bfloat16* b = O + t * D + h * d;
    uint4 xa[4], xb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = l + 4 * k;
      const bool in = ok && c < nch;
      xa[k] = in ? *reinterpret_cast<const uint4*>(a + c * 8) : make_uint4(0u, 0u, 0u, 0u);
      xb[k] = in ? *reinterpret_cast<const uint4*>(b + c * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t* ua = &xa[k].x; const uint32_t* ub = &xb[k].x;
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 x = unpack_bf16(ua[e]), y = unpack_bf16(ub[e]); acc += x.x * y.x + x.y * y.y; }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (ok && l == 0) delta[(long)h * T + t] = acc;
  }
}

// dqkv[t, 0:D] = bf16(dq_acc[t, :])
__global__ void attn_dq_convert_kernel(const float* __restrict__ dq, __nv_bfloat16* __restrict__ dqkv, long T, int D) {
  const long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 8;
  if (i >= T * D) return;
  const long t = i / D; const int c = (int)(i % D);
  const float4 x = *reinterpret_cast<const float4*>(dq + i), y = *reinterpret_cast<const float4*>(dq + i + 4);
  *reinterpret_cast<uint4*>(dqkv + t * 3 * D + c) = make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
}

template <int HD>
static int launch_bwd(const void* qkv, const void* dO, const AttnBwdArgs& a, cudaStream_t stream) {
  using Cfg = BwdCfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tq, td;
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
  CUtensorMap tdq;
  {
    uint64_t dims[2] = {(uint64_t)a.D, (uint64_t)a.T};
    uint64_t strides[1] = {(uint64_t)a.D * 4};
    uint32_t box[2] = {16, 32};
    if (make_tmap(&tdq, a.dq_acc, 2, dims, strides, box, 2, 4)) return 1;
  }
  const int grid = a.n_work < num_sms() ? a.n_work : num_sms();
  attn_bwd_kernel<HD><<<grid, 352, Cfg::SMEM_BYTES, stream>>>(tq, td, tdq, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cb

#ifdef CB_TIMELINE
extern "C" int cb_debug_timeline_bwd(void* dst) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif

extern "C" int cb_attn_varlen_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const int* work, int n_work,
                                  float* delta_ws, float* dq_acc_ws, void* dqkv, int T, int D, int H, float softmax_scale, void* stream) {
  using namespace cb;
  CB_CHECK(T > 0 && H > 0 && D % H == 0 && n_work > 0 && D % 8 == 0, "attn_bwd: bad shape T=%d D=%d H=%d n_work=%d", T, D, H, n_work);
  CB_CHECK(D % 16 == 0 && (D / H) % 16 == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 31) == 0,
           "attn_bwd: dK / dV leave as 32-byte sectors: D and head_dim must be multiples of 16 and dqkv 32-byte aligned (D=%d H=%d)", D, H);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CB_CUDA(cudaMemsetAsync(dq_acc_ws, 0, (size_t)T * D * sizeof(float), s));
  {
    CB_CHECK((D / H) % 8 == 0 && (D / H) <= 128 && ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "attn_bwd: head_dim must be a multiple of 8 (<= 128) and dout / out 16-byte aligned");
    const long quads = (long)T * H;
    long blocks = (quads * 4 + 255) / 256;
    if (blocks > (long)num_sms() * 16) blocks = (long)num_sms() * 16;
    attn_delta_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(out), delta_ws, T, D, H);
  }
  CB_CUDA(cudaGetLastError());
  AttnBwdArgs a{};
  a.work = reinterpret_cast<const int4*>(work); a.n_work = n_work; a.lse = lse; a.delta = delta_ws; a.dq_acc = dq_acc_ws;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); a.T = T; a.D = D; a.scale = softmax_scale; a.scale_log2 = softmax_scale * 1.4426950408889634f;
  // generation 2 (attn_bwd2.cu: E / D phases overlapped with the tensor pipe) needs 64 free TMEM columns: head_dim <= 96.
  // CB_ATTN_BWD_V=1 (read once) forces generation 1 everywhere — A/B measurements only.
  static const bool force_v1 = [] { const char* e = getenv("CB_ATTN_BWD_V"); return e && e[0] == '1'; }();
  int rc;
  const int hd = D / H;
  if (!force_v1 && (hd == 32 || hd == 64 || hd == 96)) {   // (head_dim 16: generation 1 measures 5 % faster)
    rc = attn_bwd2_launch(hd, qkv, dout, a, s);
  } else {
    switch (hd) {
      case 16: rc = launch_bwd<16>(qkv, dout, a, s); break;
      case 32: rc = launch_bwd<32>(qkv, dout, a, s); break;
      case 64: rc = launch_bwd<64>(qkv, dout, a, s); break;
      case 96: rc = launch_bwd<96>(qkv, dout, a, s); break;
      case 128: rc = launch_bwd<128>(qkv, dout, a, s); break;
      default: set_error("attn_bwd: unsupported head_dim %d (supported: 16, 32, 64, 96, 128)", hd); return 1;
    }
  }
  if (rc) return rc;
  const long n8 = ((long)T * D + 7) / 8;
  attn_dq_convert_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, s>>>(dq_acc_ws, reinterpret_cast<__nv_bfloat16*>(dqkv), T, D);
  CB_CUDA(cudaGetLastError());
  return 0;
}
