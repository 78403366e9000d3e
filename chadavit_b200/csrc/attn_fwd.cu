// Varlen (packed) multi-head self-attention forward for sm_100a — replaces nn.MultiheadAttention's
// softmax(QK^T/sqrt(d) + key_padding_mask)·V (chada_vit.py:105-111 -> torch F.multi_head_attention_forward).
// Key-padding semantics are reproduced by construction: sequence b only ever sees its own 1+C_b·N real tokens,
// which is exactly what masking the zero-padded channels to -inf does in the reference (SURVEY.md §8c probe).
//
//   * work item = (sequence, head, 128-query tile); persistent CTAs walk a host-built list sorted longest-first
//   * warp 0: TMA producer (Q once, K/V ring);  warp 1: single-thread tcgen05.mma issuer;
//     warps 2-5: online softmax (1 thread = 1 query row), lazy O rescale, epilogue
//   * S = Q·K^T (SS MMA) into double-buffered TMEM; P (bf16) overwrites S in TMEM and feeds O += P·V as a TS MMA
//     with V consumed MN-major straight from the TMA-swizzled tile (no transpose anywhere)
//   * fp32 softmax statistics, exp2 domain; LSE saved for the backward pass
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"

namespace cb {

constexpr int ATT_BM = 128;  // query rows per work item
constexpr int ATT_BN = 128;  // kv rows per inner tile
constexpr int ATT_KV_STAGES = 3;

template <int HD>
struct AttCfg {
  static constexpr int CHUNK = (HD % 64 == 0) ? 64 : (HD % 32 == 0 ? 32 : 16);
  static constexpr int NCH = HD / CHUNK;
  static constexpr int SWZ = CHUNK == 64 ? 3 : (CHUNK == 32 ? 2 : 1);
  static constexpr int CHUNK_BYTES = 128 * CHUNK * 2;       // one [128 x CHUNK] bf16 sub-tile
  static constexpr int TILE_BYTES = NCH * CHUNK_BYTES;      // one [128 x HD] tile
  static constexpr int SBO = 8 * CHUNK * 2;                 // 8 rows of a sub-tile
  static constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + 1024 + 256;
  static constexpr int O_COL = 256;
  static constexpr int TMEM_COLS = 512;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnFwdArgs {
  const int4* work;  // {q_row0 (global row), seq_start, seq_end, head}
  int n_work;
  __nv_bfloat16* out;  // [T, D]
  float* lse;          // [H, T]
  int T, D;
  float scale_log2;    // softmax scale * log2(e)
};

template <int HD>
__global__ void __launch_bounds__(192, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdArgs a) {
  using Cfg = AttCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::TILE_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_STAGES * Cfg::TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;                       // [KV_STAGES]
  uint64_t* v_full = k_full + ATT_KV_STAGES;         // [KV_STAGES]
  uint64_t* kv_empty = v_full + ATT_KV_STAGES;       // [KV_STAGES]
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;       // [2]
  uint64_t* p_full = s_full + 2;                     // [2]  (128 arrivals)
  uint64_t* pv_done = p_full + 2;                    // [2]
  uint64_t* o_full = pv_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < ATT_KV_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&pv_done[i], 1); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int st = 0; uint32_t ph = 0; uint32_t wi = 0;
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
        const int4 wk = a.work[w];
        const int head = wk.w;
        const int n_kv = (wk.z - wk.y + ATT_BN - 1) / ATT_BN;
        mbar_wait(q_empty, (wi & 1) ^ 1);
        mbar_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::NCH; ++c)
          tma_load_2d(sQ + c * Cfg::CHUNK_BYTES, &tmQKV, q_full, head * HD + c * Cfg::CHUNK, wk.x);
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          const int row = wk.y + j * ATT_BN;
          mbar_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c)
            tma_load_2d(sK + st * Cfg::TILE_BYTES + c * Cfg::CHUNK_BYTES, &tmQKV, &k_full[st], a.D + head * HD + c * Cfg::CHUNK, row);
          mbar_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c)
            tma_load_2d(sV + st * Cfg::TILE_BYTES + c * Cfg::CHUNK_BYTES, &tmQKV, &v_full[st], 2 * a.D + head * HD + c * Cfg::CHUNK, row);
          if (++st == ATT_KV_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, HD, false, true);
      int st = 0; uint32_t ph = 0;       // K/V ring position of the next QK
      int st_pv = 0; uint32_t ph_pv = 0; // K/V ring position of the next PV
      uint32_t it = 0;                   // global kv-tile counter (S/P double buffer)
      uint32_t wi = 0;
      const uint32_t q_addr = smem_u32(sQ);
      for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
        const int4 wk = a.work[w];
        const int n_kv = (wk.z - wk.y + ATT_BN - 1) / ATT_BN;
        mbar_wait(q_full, wi & 1);
        // (the previous work item's O is safe: its readers arrive on p_full of this item only after their epilogue)
        tc_fence_after();
        for (int j = 0; j <= n_kv; ++j) {
          if (j < n_kv) {
            // S[(it+j)&1] = Q K_j^T
            mbar_wait(&k_full[st], ph);
            tc_fence_after();
            const uint32_t k_addr = smem_u32(sK + st * Cfg::TILE_BYTES);
            const uint32_t s_tmem = tmem_base + ((it + j) & 1) * ATT_BN;
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) {
              const int c = (kk * 16) / Cfg::CHUNK, off = ((kk * 16) % Cfg::CHUNK) * 2;
              umma_ss(s_tmem, umma_smem_desc(q_addr + c * Cfg::CHUNK_BYTES + off, 16, Cfg::SBO, Cfg::SWZ),
                      umma_smem_desc(k_addr + c * Cfg::CHUNK_BYTES + off, 16, Cfg::SBO, Cfg::SWZ), idesc_qk, kk > 0 ? 1u : 0u);
            }
            tc_commit(&s_full[(it + j) & 1]);
            if (j == n_kv - 1) tc_commit(q_empty);
            if (++st == ATT_KV_STAGES) { st = 0; ph ^= 1; }
          }
          if (j >= 1) {
            // O += P_{j-1} V_{j-1}
            const uint32_t itp = it + j - 1;
            mbar_wait(&p_full[itp & 1], (itp >> 1) & 1);
            mbar_wait(&v_full[st_pv], ph_pv);
            tc_fence_after();
            const uint32_t v_addr = smem_u32(sV + st_pv * Cfg::TILE_BYTES);
            const uint32_t p_tmem = tmem_base + (itp & 1) * ATT_BN;
#pragma unroll
            for (int kk = 0; kk < ATT_BN / 16; ++kk)
              umma_ts(tmem_base + Cfg::O_COL, p_tmem + kk * 8,
                      umma_smem_desc(v_addr + kk * 16 * Cfg::CHUNK * 2, Cfg::CHUNK_BYTES, Cfg::SBO, Cfg::SWZ), idesc_pv,
                      (j > 1 || kk > 0) ? 1u : 0u);
            tc_commit(&kv_empty[st_pv]);
            tc_commit(&pv_done[itp & 1]);
            if (j == n_kv) tc_commit(o_full);
            if (++st_pv == ATT_KV_STAGES) { st_pv = 0; ph_pv ^= 1; }
          }
        }
        it += n_kv;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    uint32_t it = 0, wi = 0;
    for (int w = blockIdx.x; w < a.n_work; w += gridDim.x, ++wi) {
      const int4 wk = a.work[w];
      const int seq_len = wk.z - wk.y;
      const int n_kv = (seq_len + ATT_BN - 1) / ATT_BN;
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++it) {
        const uint32_t buf = it & 1, bph = (it >> 1) & 1;
        mbar_wait(&s_full[buf], bph);
        tc_fence_after();
        const uint32_t s_addr = lane_addr + buf * ATT_BN;
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + c * 32, sr[c]);
        tmem_ld_wait();
        const int kv_valid = seq_len - j * ATT_BN;  // columns >= kv_valid are past the end of the sequence
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(sr[c][i]) * a.scale_log2;
            if (c * 32 + i >= kv_valid) x = -INFINITY;
            sr[c][i] = __float_as_uint(x);
            mx = fmaxf(mx, x);
          }
        // lazy rescale: only move the reference max when it grows by more than 2^8 (warp-uniform decision)
        float alpha = 1.f;
        const bool need = mx > m_used + 8.f;
        const bool any_need = __any_sync(0xffffffffu, need);
        if (need) { alpha = fast_exp2(m_used - mx); m_used = mx; }  // first tile: m_used=-inf -> alpha = 0
        float sum = 0.f;
        uint32_t pk[64];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(__uint_as_float(sr[c][i]) - m_used);
            const float p1 = fast_exp2(__uint_as_float(sr[c][i + 1]) - m_used);
            sum += p0 + p1;
            pk[c * 16 + i / 2] = pack_bf16(p0, p1);
          }
        l = l * alpha + sum;
        // P -> TMEM (overwrites the S buffer just read)
        {
          uint32_t t0[32], t1[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { t0[i] = pk[i]; t1[i] = pk[32 + i]; }
          tmem_st32(s_addr, t0);
          tmem_st32(s_addr + 32, t1);
        }
        if (j > 0 && any_need) {
          // O must be complete (PV_{j-1} retired) before it is rescaled
          const uint32_t itp = it - 1;
          mbar_wait(&pv_done[itp & 1], (itp >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            if (HD - c >= 32) {
              uint32_t o[32];
              tmem_ld32(lane_addr + Cfg::O_COL + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(lane_addr + Cfg::O_COL + c, o);
            } else {
              uint32_t o[16];
              tmem_ld16(lane_addr + Cfg::O_COL + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(lane_addr + Cfg::O_COL + c, o);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[buf]);
      }
      // ---- epilogue: O / l -> bf16, LSE
      mbar_wait(o_full, wi & 1);
      tc_fence_after();
      const int grow = wk.x + r_in_tile;
      const bool ok = grow < wk.z;
      const float inv_l = 1.f / l;
      __nv_bfloat16* dst = a.out + (long)grow * a.D + wk.w * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        if (HD - c >= 32) {
          uint32_t o[32];
          tmem_ld32(lane_addr + Cfg::O_COL + c, o);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              *reinterpret_cast<uint4*>(dst + c + i) = make_uint4(
                  pack_bf16(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l),
                  pack_bf16(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l),
                  pack_bf16(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l),
                  pack_bf16(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l));
            }
          }
        } else {
          uint32_t o[16];
          tmem_ld16(lane_addr + Cfg::O_COL + c, o);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              *reinterpret_cast<uint4*>(dst + c + i) = make_uint4(
                  pack_bf16(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l),
                  pack_bf16(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l),
                  pack_bf16(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l),
                  pack_bf16(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l));
            }
          }
        }
      }
      if (ok && a.lse) a.lse[(long)wk.w * a.T + grow] = (m_used + log2f(l)) * 0.6931471805599453f;
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int HD>
static int launch_fwd(const void* qkv, const AttnFwdArgs& a, int T, int D, cudaStream_t stream) {
  using Cfg = AttCfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tm;
  uint64_t dims[2] = {(uint64_t)(3 * D), (uint64_t)T};
  uint64_t strides[1] = {(uint64_t)(3 * D) * 2};
  uint32_t box[2] = {(uint32_t)Cfg::CHUNK, 128};
  if (make_tmap(&tm, qkv, 2, dims, strides, box, Cfg::SWZ)) return 1;
  const int grid = a.n_work < num_sms() ? a.n_work : num_sms();
  attn_fwd_kernel<HD><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tm, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cb

namespace cb {
int attn_fwd2_run(const void* qkv, const int* work, int n_work, void* out, float* lse, int T, int D, int H, float softmax_scale,
                  cudaStream_t s);
}

extern "C" int cb_attn_varlen_fwd(const void* qkv, const int* work, int n_work, int q_tile, void* out, float* lse, int T, int D, int H,
                                  float softmax_scale, void* stream) {
  using namespace cb;
  CB_CHECK(T > 0 && H > 0 && D % H == 0 && n_work > 0, "attn_fwd: bad shape T=%d D=%d H=%d n_work=%d", T, D, H, n_work);
  CB_CHECK((3 * D) % 8 == 0, "attn_fwd: 3*D must be a multiple of 8");
  CB_CHECK(q_tile == 128 || q_tile == 256, "attn_fwd: q_tile must be 128 (one query tile per work item) or 256 (two)");
  if (q_tile == 256) return attn_fwd2_run(qkv, work, n_work, out, lse, T, D, H, softmax_scale, reinterpret_cast<cudaStream_t>(stream));
  AttnFwdArgs a{};
  a.work = reinterpret_cast<const int4*>(work); a.n_work = n_work; a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.T = T; a.D = D; a.scale_log2 = softmax_scale * 1.4426950408889634f;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (D / H) {
    case 16: return launch_fwd<16>(qkv, a, T, D, s);
    case 32: return launch_fwd<32>(qkv, a, T, D, s);
    case 64: return launch_fwd<64>(qkv, a, T, D, s);
    case 96: return launch_fwd<96>(qkv, a, T, D, s);
    case 128: return launch_fwd<128>(qkv, a, T, D, s);
    default: set_error("attn_fwd: unsupported head_dim %d (supported: 16, 32, 64, 96, 128)", D / H); return 1;
  }
}
