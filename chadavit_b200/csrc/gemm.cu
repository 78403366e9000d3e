// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] (+)= op(A)[M,K] * op(B)[N,K]^T
//   * operands bf16, accumulate fp32 in TMEM (double-buffered accumulators: epilogue of tile i overlaps MMA of tile i+1)
//   * TMA (cp.async.bulk.tensor) loads into 128B/64B-swizzled smem stages, mbarrier full/empty ring
//   * warps 0..7 = epilogue, warp 8 = TMA producer, warp 9 = MMA issuer (single thread; the LAST warp because the issue
//     arbiter of an SMSP prefers its highest warp id: below the busy epilogue warps the issuer starved)
//   * either operand may be K-major ([rows,K] row-major) or MN-major ([K,rows] row-major); the latter is what the
//     weight-gradient (dW = dY^T X) and input-gradient (dX = dY W) products of the encoder need without transposes
//   * epilogue (8 warps): TMEM -> registers (bias, ReLU, residual add, ReLU-mask, tokenizer embeddings) -> 128B-swizzled
//     per-warp smem slab (32 rows x 128 B) -> read back row-wise so every global store / fp32 atomic is a full coalesced
//     128-byte line (a thread owns a ROW of the accumulator in TMEM, so direct stores would scatter 32 rows per instruction)
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_SLAB_BYTES = 32 * 128;            // one warp's slab: 32 rows x 128 B
constexpr int EPI_BYTES = 8 * EPI_SLAB_BYTES;       // one staging slab per epilogue warp

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 128) ? 5 : 4;
  static constexpr int ACC_STRIDE = (BN <= 128) ? 128 : 256;  // TMEM columns between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// MODE: 0 = K-major (128B swizzle, 64-element K rows), 1 = MN-major 64-element blocks (128B swizzle),
//       2 = MN-major 32-element blocks (64B swizzle)
template <int MODE>
__device__ __forceinline__ uint64_t operand_desc(uint32_t saddr, int k16) {
  if (MODE == 0) return umma_smem_desc(saddr + k16 * 32, 16, 1024, 3);
  if (MODE == 1) return umma_smem_desc(saddr + k16 * 2048, 64 * 128, 1024, 3);
  return umma_smem_desc(saddr + k16 * 1024, 64 * 64, 512, 2);
}

template <int MODE>
__device__ __forceinline__ void operand_load(void* smem, const CUtensorMap* tm, uint64_t* bar, int k0, int r0) {
  if (MODE == 0) tma_load_2d(smem, tm, bar, k0, r0);
  else if (MODE == 1) tma_load_3d(smem, tm, bar, 0, k0, r0 / 64);
  else tma_load_3d(smem, tm, bar, 0, k0, r0 / 32);
}

enum EpiMode { EM_BF16 = 0, EM_BF16_MASK = 1, EM_F32 = 2, EM_ATOMIC = 3, EM_TOKENIZE = 4 };

constexpr int GEMM_THREADS = 320;  // warps 0..7 epilogue, warp 8 TMA, warp 9 MMA
constexpr int GEMM_W_TMA = 8, GEMM_W_MMA = 9;

// WRES ("weights resident", K <= 192, no split-K): a CTA keeps ONE column tile of B — the whole [BN x K] weight block — in
// shared memory for its lifetime and streams row tiles of A past it (full-K A tiles in a 2-deep ring, one barrier round trip
// and 12 back-to-back MMAs per output tile).  The small-K products of the encoder (qkv, proj, fc1 and the input-gradient
// GEMMs against W2 / Wo) are bound by shared-memory bandwidth, not by the tensor pipe: per 128 x 256 output tile the staged
// kernel moved 144 KB of operands INTO smem by TMA, 144 KB out again for the MMAs and 128 KB through the epilogue slabs
// (3250 clk at 128 B/clk against 1536 clk of MMA; ncu: L1TEX/smem pipe the busiest unit).  Re-loading the same 96 KB weight
// tile for every row tile was the avoidable third of that.
constexpr int WRES_A_STAGES = 2;

// TSTORE (weights-resident, bf16 output, BN = 192: the qkv projection and the out-projection's input gradient): the output tile
// leaves through shared memory and THREE TMA stores instead of per-thread 32-byte stores.  A "thread = row" store instruction
// writes one sector into each of 32 different lines, and that pattern is what bounds the row-direct epilogues (~3 clk per sector
// per SM: qkv stored at 2.6 TB/s where a contiguous stream writes 6.3, profiles/r02_hw_probe_hbm_stream.txt).  The tile is staged
// as three [128 rows x 64 columns] 128B-swizzled boxes (16-byte writes, conflict-free per quarter warp); one thread issues the
// stores, and the staging area is reused once the previous tile's stores have read it (cp.async.bulk.wait_group.read).
constexpr int TSTORE_BYTES = 3 * 128 * 128;        // 48 KB
template <int BN, int EM, bool WRES, bool CSUM, bool LNF>
__host__ __device__ constexpr bool use_tstore() {
#ifdef CB_NO_TSTORE      // A/B variant build (CB_VARIANT=prev CB_NVCC_EXTRA=-DCB_NO_TSTORE): the row-direct epilogue everywhere
  return false;
#else
  return BN == 192 && EM == 0 /* EM_BF16 */ && WRES && !CSUM && !LNF;
#endif
}

// CSUM (weight-gradient products dW = dY^T X, both operands MN-major, split-K, fp32 atomics): the sums of op(A) over K — the
// bias gradient that belongs to dW (rows of dY^T summed over the tokens) — are formed by the SAME tcgen05.mma that forms the
// tile: an 8 KB block of bf16 ones sits behind every B stage, exactly where a fourth 64-column block of an MN-major B tile
// would be, and the instruction runs with N = BN + 16 instead of BN (104 instead of 96 clk per k-step in a kernel that waits
// for HBM).  TMEM column BN of the accumulator then holds sum_k A[m, k]; the epilogue adds it to g.colsum[m] with one atomic
// per row.  This replaces the 31-shuffle transpose-reduce per slab in the d(hidden) epilogue (205 -> 168 us per 68 k tokens)
// and the separate column-sum pass over d(qkv).
constexpr int CSUM_ONES_BYTES = 64 * 128;          // one 64 (k) x 64 (n) bf16 block
constexpr int CSUM_EXTRA_N = 16;

__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

template <int BN, int AMODE, int BMODE, int EM, bool WRES, bool CSUM = false, bool LNF = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
            const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr bool TSTORE = use_tstore<BN, EM, WRES, CSUM, LNF>();
  constexpr int EPI_SZ = TSTORE ? TSTORE_BYTES + 1024 : EPI_BYTES;   // staging boxes + bias
  static_assert(!CSUM || (BN == 192 && BMODE == 1 && EM == EM_ATOMIC && !WRES), "CSUM: split-K weight-gradient product with BN = 192 and MN-major B only");
  static_assert(!LNF || (BN == 192 && EM == EM_F32 && WRES && AMODE == 0), "LNF: weights-resident 192-wide fp32 product only");
  constexpr int B_STRIDE = Cfg::B_BYTES + (CSUM ? CSUM_ONES_BYTES : 0);   // smem distance between two B stages
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  const int kb_total = (g.K + BK - 1) / BK;
  // staged: [STAGES x A k-block][STAGES x B k-block][epilogue slabs]; WRES: [kb_total x B k-block][2 x kb_total x A k-block][slabs]
  uint8_t* sB = WRES ? smem : smem + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* sA = WRES ? smem + kb_total * Cfg::B_BYTES : smem;
  uint8_t* sEpi = WRES ? sA + WRES_A_STAGES * kb_total * Cfg::A_BYTES : smem + Cfg::STAGES * (Cfg::A_BYTES + B_STRIDE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + EPI_SZ);
  uint64_t* full = bars;                         // WRES: full[0..1] = A stage landed, full[2] = weights landed
  uint64_t* empty = bars + Cfg::STAGES;          // WRES: empty[0..1]
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (g.M + BM - 1) / BM, num_n = (g.N + BN - 1) / BN;
  const int kb_per = (kb_total + g.k_splits - 1) / g.k_splits;
  const int num_tiles = num_m * num_n * g.k_splits;
  // tile enumeration: staged = tiles blockIdx.x, + gridDim.x, ... (n fastest); WRES = fixed column tile blockIdx.x % num_n, row
  // tiles blockIdx.x / num_n, + gridDim.x / num_n, ...  (the host launches a multiple of num_n CTAs)
  const int w_n0 = (blockIdx.x % num_n) * BN, w_m_first = blockIdx.x / num_n, w_m_step = gridDim.x / num_n;

  if (warp == GEMM_W_TMA && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == GEMM_W_MMA) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (CSUM) {   // the constant ones block behind every B stage (swizzle-invariant: every element is 1.0)
    for (int i = threadIdx.x; i < Cfg::STAGES * (CSUM_ONES_BYTES / 16); i += GEMM_THREADS)
      *reinterpret_cast<uint4*>(sB + (i / (CSUM_ONES_BYTES / 16)) * B_STRIDE + Cfg::B_BYTES + (i % (CSUM_ONES_BYTES / 16)) * 16) =
          make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor-core (async) proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == GEMM_W_TMA) {
    if (WRES) {
      if (lane == 0) {
        mbar_expect_tx(&full[2], kb_total * Cfg::B_BYTES);
        for (int kb = 0; kb < kb_total; ++kb) operand_load<BMODE>(sB + kb * Cfg::B_BYTES, &tmB, &full[2], kb * BK, w_n0);
        int s = 0; uint32_t ph = 0;
        for (int mi = w_m_first; mi < num_m; mi += w_m_step) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kb_total * Cfg::A_BYTES);
          for (int kb = 0; kb < kb_total; ++kb) operand_load<AMODE>(sA + (s * kb_total + kb) * Cfg::A_BYTES, &tmA, &full[s], kb * BK, mi * BM);
          if (++s == WRES_A_STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile / (num_m * num_n), mn = tile % (num_m * num_n);
        const int m0 = (mn / num_n) * BM, n0 = (mn % num_n) * BN;
        const int kb0 = split * kb_per, kb1 = min(kb0 + kb_per, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          operand_load<AMODE>(sA + s * Cfg::A_BYTES, &tmA, &full[s], kb * BK, m0);
          operand_load<BMODE>(sB + s * B_STRIDE, &tmB, &full[s], kb * BK, n0);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == GEMM_W_MMA) {
    // The whole warp runs the loop convergently, one elected lane issues: warp-uniform control flow lets the compiler keep
    // descriptors in uniform registers and emit back-to-back UTCHMMA (under `if (lane == 0)` each MMA cost ~12 instructions).
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN + (CSUM ? CSUM_EXTRA_N : 0), AMODE != 0, BMODE != 0);
    const uint64_t a_desc0 = operand_desc<AMODE>(smem_u32(sA), 0), b_desc0 = operand_desc<BMODE>(smem_u32(sB), 0);
    constexpr uint32_t a_kstep = AMODE == 0 ? 32 : (AMODE == 1 ? 2048 : 1024), b_kstep = BMODE == 0 ? 32 : (BMODE == 1 ? 2048 : 1024);
    int s = 0; uint32_t ph = 0; int it = 0;
    if (WRES) {
      mbar_wait(&full[2], 0);
      CB_TL_DECL(tl);
      for (int mi = w_m_first; mi < num_m; mi += w_m_step, ++it) {
        const int buf = it & 1; const uint32_t aph = (it >> 1) & 1;
        CB_TL(0, tl, 1);
        mbar_wait(&acc_empty[buf], aph ^ 1);
        CB_TL(0, tl, 2);
        mbar_wait(&full[s], ph);
        tc_fence_after();
        CB_TL(0, tl, 3);
        const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_STRIDE;
        const uint64_t ad = umma_desc_add(a_desc0, s * kb_total * Cfg::A_BYTES);
        if (elect_one()) {
          for (int kb = 0; kb < kb_total; ++kb) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(d_tmem, umma_desc_add(ad, kb * Cfg::A_BYTES + k * a_kstep), umma_desc_add(b_desc0, kb * Cfg::B_BYTES + k * b_kstep), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty[s]);
          tc_commit(&acc_full[buf]);
        }
        __syncwarp();
        CB_TL(0, tl, 4);
        if (++s == WRES_A_STAGES) { s = 0; ph ^= 1; }
      }
    } else
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int split = tile / (num_m * num_n);
      const int kb0 = split * kb_per, kb1 = min(kb0 + kb_per, kb_total);
      const int buf = it & 1; const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&acc_empty[buf], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_STRIDE;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t ad = umma_desc_add(a_desc0, s * Cfg::A_BYTES), bd = umma_desc_add(b_desc0, s * B_STRIDE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss(d_tmem, umma_desc_add(ad, k * a_kstep), umma_desc_add(bd, k * b_kstep), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          tc_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: 8 warps; warp w owns TMEM lanes 32*(w&3) .. +31 (rows) and every other 32-column slab.
    // Raw fp32 accumulators are staged in a swizzled smem slab (thread = row), then read back row-wise (8 lanes = one
    // 128-byte row) where ALL epilogue math happens, so bias / residual / mask reads and every store are coalesced and
    // the per-column vectors are loaded once per slab (the L1 is ~3 KB next to 225 KB of smem: scalar __ldg's would all
    // be serialised L2 round trips).
    const int q = warp & 3;
    const int ew = warp;                         // 0..7
    const int half = ew >> 2;                    // which of the two warps of this lane quarter
    constexpr int mode = EM;   // compile-time epilogue mode: dead branches and their register arrays disappear
    // ---------------- row-direct epilogue (bf16 / fp32 outputs of a CTA with a FIXED column tile: WRES or a single column tile)
    // A thread owns a ROW of the accumulator, i.e. 32 consecutive output columns per slab = 64 contiguous bytes of a bf16
    // row (128 of an fp32 row): it stores them itself as full 32-byte sectors (STG.256).  No smem staging, no warp
    // synchronisation, ~95 instead of ~170 instructions per slab; the bias of the column tile sits in smem (broadcast LDS).
    // In-kernel timeline of fc1 (profiles/r01_timeline_gemm.txt): the staged epilogue needed 1100 clk per slab, 4400 clk per
    // 128 x 256 tile against 1536 clk of MMA — the tensor pipe waited for the epilogue 60 % of the time.

    // ---------------- TSTORE: bf16 tile -> swizzled shared-memory boxes -> TMA stores (see above)
    if constexpr (TSTORE) {
      float* sBias = reinterpret_cast<float*>(sEpi + TSTORE_BYTES);
      const int n0 = w_n0;
      for (int i = threadIdx.x; i < BN; i += 256) sBias[i] = (g.bias != nullptr && n0 + i < g.N) ? __ldg(g.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float lo = (g.flags & CB_EPI_RELU) ? 0.f : -INFINITY;
      const int rr = q * 32 + lane;
      int it = 0;
      for (int tile = w_m_first; tile < num_m; tile += w_m_step, ++it) {
        const int m0 = tile * BM;
        const int buf = it & 1; const uint32_t aph = (it >> 1) & 1;
        const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + buf * Cfg::ACC_STRIDE;
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after();
        if (it > 0 && threadIdx.x == 0) tma_store_wait_read<0>();    // the previous tile's stores have read the staging boxes
        asm volatile("bar.sync 2, 256;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int sl = 3 * half + j, c = sl * 32;
          uint32_t x[32];
          tmem_ld32(t_addr + c, x);
          tmem_ld_wait();
          if (j == 2) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }   // accumulator fully read
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c + e);
            pk[e >> 1] = pack_bf16(fmaxf(fmaf(__uint_as_float(x[e]), g.alpha, b4.x), lo), fmaxf(fmaf(__uint_as_float(x[e + 1]), g.alpha, b4.y), lo));
            pk[(e >> 1) + 1] = pack_bf16(fmaxf(fmaf(__uint_as_float(x[e + 2]), g.alpha, b4.z), lo), fmaxf(fmaf(__uint_as_float(x[e + 3]), g.alpha, b4.w), lo));
          }
          // slab sl = columns [32 sl, 32 sl + 32) = half (sl & 1) of the 128-byte row rr of box sl >> 1: 16-byte chunks 4 (sl & 1) + k
          uint8_t* rowp = sEpi + (sl >> 1) * (128 * 128) + rr * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(rowp + (((4 * (sl & 1) + k) ^ (rr & 7)) << 4)) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        }
        fence_proxy_async();      // generic-proxy writes of the boxes -> visible to the TMA (async proxy)
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (threadIdx.x == 0) {
#pragma unroll
          for (int b = 0; b < 3; ++b) tma_store_2d(&tmC, sEpi + b * (128 * 128), n0 + 64 * b, m0);
          tma_store_commit();
        }
      }
      if (threadIdx.x == 0) tma_store_wait_all<0>();
    } else
    // ---------------- LayerNorm fused into the row epilogue (LNF; the encoder's  y = norm1(x + attn W_o^T + b_o),  chada_vit.py:99).
    // The 192-wide tile holds whole rows; a row is shared by the two warps of its lane quarter (columns 0..95 / 96..191), which
    // exchange their partial sums through shared memory: once for the mean, once for the centred sum of squares (the same
    // two-pass arithmetic as layernorm_fwd_g16_kernel).  A thread keeps its 96 values in registers: the residual is loaded INTO
    // that buffer before the accumulator wait and every TMEM slab is folded into it in place.  Replaces the separate LayerNorm
    // launch (36 per training step) and its read of z1; z1 itself is stored only when the caller keeps it for the backward.
    if constexpr (LNF) {
      float* sBias = reinterpret_cast<float*>(sEpi);
      float* sGam = sBias + 192;
      float* sBet = sGam + 192;
      float* sEx = sBet + 192;                     // [tile parity 2][pass 2][half 2][128 rows]
      for (int i = threadIdx.x; i < 192; i += 256) {
        sBias[i] = g.bias != nullptr ? __ldg(g.bias + i) : 0.f;
        sGam[i] = __ldg(g.ln_gamma + i);
        sBet[i] = __ldg(g.ln_beta + i);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool has_res = (g.flags & CB_EPI_RESIDUAL_F32) != 0;
      const int c0 = 96 * half;                    // this warp's columns
      const int rr = q * 32 + lane;                // row inside the tile
      int it = 0;
      for (int tile = w_m_first; tile < num_m; tile += w_m_step, ++it) {
        const int m0 = tile * BM;
        const int buf = it & 1; const uint32_t aph = (it >> 1) & 1;
        const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + buf * Cfg::ACC_STRIDE + c0;
        const long row = (long)m0 + rr;
        const bool row_ok = row < g.M;
        uint32_t v[3][32];
        if (row_ok && has_res) {
          const float* rp = reinterpret_cast<const float*>(g.aux) + row * g.ld_aux + c0;
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) ldg256(rp + j * 32 + 8 * k, *reinterpret_cast<uint32_t(*)[8]>(&v[j][8 * k]));
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 32; ++e) v[j][e] = 0u;
        }
        // The NEXT tile's residual rows -> L2 (no registers to prefetch them into: the row buffer is the register file): the
        // epilogues of a CTA's ~7 tiles run strictly one after the other, each starting with this 384-byte-per-thread load.
        if (has_res) {   // (same-box A/B: 61.5 -> 58.4 us with z kept, 50.0 -> 46.8 us without, per 68 k tokens)
          const long nrow = row + (long)w_m_step * BM;
          if (nrow < g.M) {
            const float* np = reinterpret_cast<const float*>(g.aux) + nrow * g.ld_aux + c0;
#pragma unroll
            for (int k = 0; k < 3; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + 32 * k));
          }
        }
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after();
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t x[32];
          tmem_ld32(t_addr + j * 32, x);
          tmem_ld_wait();
          if (j == 2) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }   // accumulator fully read
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + j * 32 + e);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float val = fmaf(__uint_as_float(x[e + t]), g.alpha, bb[t]) + __uint_as_float(v[j][e + t]);
              v[j][e + t] = __float_as_uint(val);
              s += val;
            }
          }
        }
        if (g.C != nullptr && row_ok) {
          float* dst = reinterpret_cast<float*>(g.C) + row * g.ldc + c0;
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              stg256(dst + j * 32 + 8 * k, v[j][8 * k], v[j][8 * k + 1], v[j][8 * k + 2], v[j][8 * k + 3], v[j][8 * k + 4], v[j][8 * k + 5], v[j][8 * k + 6], v[j][8 * k + 7]);
        }
        float* ex = sEx + (it & 1) * 512;
        ex[half * 128 + rr] = s;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        const float mean = (ex[rr] + ex[128 + rr]) * (1.f / 192.f);
        float qq = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int e = 0; e < 32; ++e) { const float d = __uint_as_float(v[j][e]) - mean; qq += d * d; }
        ex[256 + half * 128 + rr] = qq;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        const float rstd = rsqrtf((ex[256 + rr] + ex[256 + 128 + rr]) * (1.f / 192.f) + g.ln_eps);
        if (row_ok) {
          if (half == 0) { if (g.ln_mean) g.ln_mean[row] = mean; if (g.ln_rstd) g.ln_rstd[row] = rstd; }
          __nv_bfloat16* d16 = g.ln_y16 + row * 192 + c0;
          float* d32 = g.ln_y32 ? g.ln_y32 + row * 192 + c0 : nullptr;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            uint32_t o[32];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 g4 = *reinterpret_cast<const float4*>(sGam + c0 + j * 32 + e), b4 = *reinterpret_cast<const float4*>(sBet + c0 + j * 32 + e);
              o[e] = __float_as_uint((__uint_as_float(v[j][e]) - mean) * rstd * g4.x + b4.x);
              o[e + 1] = __float_as_uint((__uint_as_float(v[j][e + 1]) - mean) * rstd * g4.y + b4.y);
              o[e + 2] = __float_as_uint((__uint_as_float(v[j][e + 2]) - mean) * rstd * g4.z + b4.z);
              o[e + 3] = __float_as_uint((__uint_as_float(v[j][e + 3]) - mean) * rstd * g4.w + b4.w);
            }
            if (d32) {
#pragma unroll
              for (int k = 0; k < 4; ++k) stg256(d32 + j * 32 + 8 * k, o[8 * k], o[8 * k + 1], o[8 * k + 2], o[8 * k + 3], o[8 * k + 4], o[8 * k + 5], o[8 * k + 6], o[8 * k + 7]);
            }
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1]));
            stg256(d16 + j * 32, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            stg256(d16 + j * 32 + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
          }
        }
      }
    } else
    if ((mode == EM_BF16 || mode == EM_F32 || mode == EM_BF16_MASK) && g.direct && (WRES || num_n == 1)) {
      float* sBias = reinterpret_cast<float*>(sEpi);
      const int n0 = WRES ? w_n0 : 0;
      for (int i = threadIdx.x; i < BN; i += 256) sBias[i] = (g.bias != nullptr && n0 + i < g.N) ? __ldg(g.bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int n_lim = min(BN, g.N - n0);
      const int n_slabs = (n_lim + 31) / 32;
      const float lo = (g.flags & CB_EPI_RELU) ? 0.f : -INFINITY;
      const bool has_res = mode == EM_F32 && (g.flags & CB_EPI_RESIDUAL_F32);
      // EM_BF16_MASK: fused bias gradient of linear1 = column sums of the STORED (masked, bf16-rounded) values.  Lane c of this
      // warp accumulates column 32 sl + c of its slabs over all row tiles of the CTA (the column tile is fixed), one atomic each
      // at the end.  csum[j] belongs to this warp's j-th slab (s0 .. s3 below).
      float csum[4] = {0.f, 0.f, 0.f, 0.f};
      int it = 0;
      for (int tile = WRES ? w_m_first : blockIdx.x; tile < (WRES ? num_m : num_tiles); tile += (WRES ? w_m_step : gridDim.x), ++it) {
        const int m0 = (WRES ? tile : tile / num_n) * BM;     // num_n == 1 in the staged case (k_splits == 1: no atomic mode here)
        const int buf = it & 1; const uint32_t aph = (it >> 1) & 1;
        const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + buf * Cfg::ACC_STRIDE;
        const long row = (long)m0 + q * 32 + lane;
        const bool row_ok = row < g.M;
        // Accumulator-independent operand of a slab (fp32 residual row segment, 128 B, or ReLU mask = hidden activations, 64 B):
        // requested TWO slabs ahead — the first two before the accumulator wait — so that the DRAM latency (~1000 clk) of these
        // row-strided reads is covered by the processing of the slabs in between (8 warps x 2 loads in flight per thread).
        constexpr int NAUX = mode == EM_F32 ? 32 : 16;
        auto aux_load = [&](uint32_t (&m)[NAUX], int sl) {
          const int c = sl * 32;
#pragma unroll
          for (int e = 0; e < NAUX; ++e) m[e] = 0u;
          if (!row_ok) return;
          if (mode == EM_F32) {
            if (!has_res) return;
            const float* rp = reinterpret_cast<const float*>(g.aux) + row * g.ld_aux + n0 + c;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (c + 8 * k + 8 <= n_lim) ldg256(rp + 8 * k, *reinterpret_cast<uint32_t(*)[8]>(&m[8 * k]));
          } else if (mode == EM_BF16_MASK) {
            if (g.flags & CB_EPI_MASK_BITS) {   // one coalesced word per slab: bit j <=> hidden[row, n0 + c + j] > 0
              m[0] = __ldg(reinterpret_cast<const uint32_t*>(g.aux) + (long)((n0 + c) >> 5) * g.ld_aux + row);
              return;
            }
            const __nv_bfloat16* mp = g.aux + row * g.ld_aux + n0 + c;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (c + 16 * k + 16 <= n_lim) ldg256(mp + 16 * k, *reinterpret_cast<uint32_t(*)[8]>(&m[8 * k]));
              else if (c + 16 * k + 8 <= n_lim) { const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(mp + 16 * k)); m[8 * k] = m4.x; m[8 * k + 1] = m4.y; m[8 * k + 2] = m4.z; m[8 * k + 3] = m4.w; }
            }
          }
        };
        // one slab: accumulator wait -> next slab's TMEM read -> math -> stores.  `x` holds this slab, `nxt` receives the next
        // one (two statically indexed register sets, alternated by the unrolled sequence below); `m` = its residual / mask.
        auto slab_step = [&](uint32_t (&x)[32], uint32_t (&nxt)[32], const uint32_t (&m)[NAUX], int sl, int sl_next, float& cs) {
          const int c = sl * 32;
          tmem_ld_wait();
          if (sl_next < n_slabs) tmem_ld32(t_addr + sl_next * 32, nxt);         // next slab in flight under this slab's math
          else { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }   // accumulator fully read (the list is ascending: no later slab of this warp exists)
          if (mode == EM_BF16_MASK) {
            // d(hidden) = (dz2 . W2) where hidden > 0.  bf16 > 0  <=>  its bit pattern, read as a signed 16-bit integer, is > 0.
            uint32_t pk[16];
            float f[32];
            const bool bits = (g.flags & CB_EPI_MASK_BITS) != 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const uint32_t mm = m[e];
              const uint32_t b2 = m[0] >> (2 * e);
              const uint32_t sel = bits ? (((b2 & 1u) ? 0xffffu : 0u) | ((b2 & 2u) ? 0xffff0000u : 0u))
                                        : (((short)(mm & 0xffffu) > 0 ? 0xffffu : 0u) | ((int)mm >= 0x10000 ? 0xffff0000u : 0u));
              pk[e] = pack_bf16(__uint_as_float(x[2 * e]) * g.alpha, __uint_as_float(x[2 * e + 1]) * g.alpha) & sel;   // rows past M: mask = 0
              const float2 u = unpack_bf16(pk[e]);
              f[2 * e] = u.x; f[2 * e + 1] = u.y;
            }
            if (row_ok) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.C) + row * g.ldc + n0 + c;
              if (c + 32 <= n_lim) {
                stg256(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                stg256(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (c + 8 * k < n_lim) *reinterpret_cast<uint4*>(dst + 8 * k) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
              }
            }
            if (g.colsum) cs += warp_colsum32(f, lane);
            return;
          }
          if (!row_ok) return;
          if (mode == EM_BF16) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBias + c + e);
              pk[e >> 1] = pack_bf16(fmaxf(fmaf(__uint_as_float(x[e]), g.alpha, b4.x), lo), fmaxf(fmaf(__uint_as_float(x[e + 1]), g.alpha, b4.y), lo));
              pk[(e >> 1) + 1] = pack_bf16(fmaxf(fmaf(__uint_as_float(x[e + 2]), g.alpha, b4.z), lo), fmaxf(fmaf(__uint_as_float(x[e + 3]), g.alpha, b4.w), lo));
            }
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.C) + row * g.ldc + n0 + c;
            if (c + 32 <= n_lim) {
              stg256(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
              stg256(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (c + 8 * k < n_lim) *reinterpret_cast<uint4*>(dst + 8 * k) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
            }
          } else {
            float* dst = reinterpret_cast<float*>(g.C) + row * g.ldc + n0 + c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (c + 8 * k + 8 <= n_lim) {
                const float4 b0 = *reinterpret_cast<const float4*>(sBias + c + 8 * k), b1 = *reinterpret_cast<const float4*>(sBias + c + 8 * k + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint32_t o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(fmaxf(fmaf(__uint_as_float(x[8 * k + e]), g.alpha, bb[e]), lo) + __uint_as_float(m[(8 * k + e) % NAUX]));
                stg256(dst + 8 * k, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
              }
            }
          }
        };
        // this warp's slabs (BN <= 256: at most four): PAIRS of adjacent slabs, so that a thread writes 128 contiguous bytes of a
        // bf16 row (a whole line) back to back instead of leaving every line half-written until the partner warp gets to it.
        // BN = 192 has six slabs: dealt in pairs that is {0,1,4,5} / {2,3} — one warp of every lane quarter does twice the work
        // of its partner and the tile's epilogue takes four slab times — so there the warps take three consecutive slabs each.
        constexpr bool TRI = (BN == 192);
        const int s0 = TRI ? 3 * half : 2 * half, s1 = s0 + 1, s2 = TRI ? s0 + 2 : 2 * half + 4, s3 = TRI ? 1 << 20 : 2 * half + 5;
        uint32_t xa[32], xb[32], ma[NAUX], mb[NAUX];
        if (mode != EM_BF16) { if (s0 < n_slabs) aux_load(ma, s0); if (s1 < n_slabs) aux_load(mb, s1); }
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after();
        if (s0 < n_slabs) tmem_ld32(t_addr + s0 * 32, xa);
        else { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }   // no slab for this warp in a narrow tile
        if (s0 < n_slabs) { slab_step(xa, xb, ma, s0, s1, csum[0]); if (mode != EM_BF16 && s2 < n_slabs) aux_load(ma, s2); }
        if (s1 < n_slabs) { slab_step(xb, xa, mb, s1, s2, csum[1]); if (mode != EM_BF16 && s3 < n_slabs) aux_load(mb, s3); }
        if (s2 < n_slabs) slab_step(xa, xb, ma, s2, s3, csum[2]);
        if (s3 < n_slabs) slab_step(xb, xa, mb, s3, n_slabs, csum[3]);
      }
      if (mode == EM_BF16_MASK && g.colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int sj = (BN == 192) ? (j < 3 ? 3 * half + j : 1 << 20) : 2 * half + (j & 1) + 4 * (j >> 1);   // the slab csum[j] belongs to (s0 .. s3 above)
          const int col = n0 + sj * 32 + lane;
          if (sj < n_slabs && col < g.N) atomicAdd(g.colsum + col, csum[j]);
        }
      }
    } else {
    uint8_t* slab = sEpi + ew * EPI_SLAB_BYTES;
    uint8_t* srow = slab + lane * 128;
    const int rb_row = lane >> 3, rb_chunk = lane & 7;   // read-back mapping: 8 lanes cover one 128-byte row (4 columns each)
    int it = 0;
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4);
    for (int tile = WRES ? w_m_first : blockIdx.x; tile < (WRES ? num_m : num_tiles); tile += (WRES ? w_m_step : gridDim.x), ++it) {
      const int split = WRES ? 0 : tile / (num_m * num_n), mn = WRES ? 0 : tile % (num_m * num_n);
      const int m0 = WRES ? tile * BM : (mn / num_n) * BM, n0 = WRES ? w_n0 : (mn % num_n) * BN;
      const int buf = it & 1; const uint32_t aph = (it >> 1) & 1;
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + buf * Cfg::ACC_STRIDE;
      const int n_lim = min(BN, g.N - n0);  // valid columns of this tile (multiple of 8)
      const int n_slabs = (n_lim + 31) / 32;
      const int last_slab = ((n_slabs - 1 - half) >= 0) ? (n_slabs - 1 - ((n_slabs - 1 - half) & 1)) : -1;  // last slab of this warp
      int2 ri = make_int2(-1, -1);   // tokenizer: this lane's row -> {pos row offset, chan row offset}; shuffled in the read-back
      if (g.flags & CB_EPI_TOKENIZE) {
        // row t of sequence b: off = t - cu[b]; off == 0 -> CLS row (cls_token + pos_embed[0]), else patch p of channel c
        const int row = m0 + q * 32 + lane;
        if (row < g.M) {
          int lo = 0, hi = g.nseq;  // largest b with cu[b] <= row
          while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(g.cu + mid) <= row) lo = mid; else hi = mid; }
          const int off = row - __ldg(g.cu + lo);
          if (off == 0) ri = make_int2(-2, -2);
          else { const int c = (off - 1) / g.npatch; ri = make_int2(((off - 1) - c * g.npatch) * g.N, g.chan_tok ? c * g.N : -1); }
        }
      }
      // ncu (profiles/r01_ncu_gemm_fc1.txt): the two hottest stalls were the first use of the bias (an L2 round trip per
      // slab) and the tcgen05.ld wait.  So: the bias of a slab is requested one slab ahead (the first one before the
      // accumulator wait) and the TMEM read of slab s+1 is issued right after slab s was staged, under its math/stores.
      const bool has_bias = g.bias != nullptr && split == 0 && mode != EM_BF16_MASK;
      float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f);
      float biasl_next = 0.f;                         // EM_BF16: bias of column (slab start + lane)
      if (mode == EM_BF16) {
        if (has_bias && last_slab >= 0 && n0 + half * 32 + lane < g.N) biasl_next = __ldg(g.bias + n0 + half * 32 + lane);
      } else if (has_bias && last_slab >= 0 && n0 + half * 32 + rb_chunk * 4 < g.N) {
        bias_next = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + half * 32 + rb_chunk * 4));
      }
      if (tl_on) CB_TL(1 + half, tl, 1);
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      if (tl_on) CB_TL(1 + half, tl, 2);
      if (CSUM && half == 0) {   // column BN of the accumulator: sum over this split's K range of op(A)[row, k]
        uint32_t v;
        tmem_ld1(t_addr + BN, v);
        tmem_ld_wait();
        const long srow = (long)m0 + q * 32 + lane;
        if (srow < g.M) atomicAdd(g.colsum + srow, __uint_as_float(v) * g.alpha);
      }
      if (last_slab < 0) {   // nothing to do for this warp in this tile: still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      const long row0 = (long)m0 + q * 32 + rb_row;                               // this lane's rows: row0 + 4*i
      const int nrows = (int)max((long)0, min((long)8, (g.M - row0 + 3) / 4));    // valid i range (a prefix)
      uint32_t r[32];
      tmem_ld32(t_addr + half * 32, r);
#pragma unroll 1
      for (int sl = half; sl < n_slabs; sl += 2) {
        const int c = sl * 32;
        const int gcol = n0 + c + rb_chunk * 4;
        const bool col_ok = gcol < g.N;
        const float4 bias4 = bias_next;
        if (mode != EM_BF16 && has_bias && sl + 2 < n_slabs && gcol + 64 < g.N) bias_next = __ldg(reinterpret_cast<const float4*>(g.bias + gcol + 64));
        // accumulator-independent operands (ReLU mask / fp32 residual) are requested before the TMEM wait
        uint4 mk[4];            // EM_BF16_MASK: the ReLU mask (hidden activations) of this lane's 4 x 8 output elements
        float4 res[8];
        if (mode == EM_BF16_MASK) {
          const int gc8 = n0 + c + (lane & 3) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const long grow = (long)m0 + q * 32 + i * 8 + (lane >> 2);
            if (g.flags & CB_EPI_MASK_BITS) {   // expand this lane's 8 mask bits into the bf16 pattern the read-back tests (1.0 / 0)
              const uint32_t w = (grow < g.M && gc8 < g.N) ? __ldg(reinterpret_cast<const uint32_t*>(g.aux) + (long)(gc8 >> 5) * g.ld_aux + grow) >> (gc8 & 31) : 0u;
              uint32_t* mw = &mk[i].x;
#pragma unroll
              for (int k = 0; k < 4; ++k) mw[k] = ((w >> (2 * k)) & 1u ? 0x3f80u : 0u) | ((w >> (2 * k + 1)) & 1u ? 0x3f800000u : 0u);
            } else {
              mk[i] = (grow < g.M && gc8 < g.N) ? __ldg(reinterpret_cast<const uint4*>(g.aux + grow * g.ld_aux + gc8)) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        if (mode == EM_F32) {
          if (g.flags & CB_EPI_RESIDUAL_F32) {
            const float* rp = reinterpret_cast<const float*>(g.aux) + row0 * g.ld_aux + gcol;
#pragma unroll
            for (int i = 0; i < 8; ++i) res[i] = (col_ok && i < nrows) ? __ldg(reinterpret_cast<const float4*>(rp + (long)i * 4 * g.ld_aux)) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tmem_ld_wait();
        if (tl_on) CB_TL(1 + half, tl, 3);
        if (sl == last_slab) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf]); }
        if (mode == EM_BF16) {
          // bf16 outputs: bias (+ReLU) and the bf16 rounding happen in the thread=row layout (the bias of column c+lane
          // lives in lane `lane`, broadcast by shuffles), so only 64 B per row go through the smem slab (half the traffic
          // of staging fp32: ncu showed the L1TEX/smem pipe as the busiest unit of these write-heavy GEMMs).
          const float bl = biasl_next;
          if (has_bias && sl + 2 < n_slabs && n0 + c + 64 + lane < g.N) biasl_next = __ldg(g.bias + n0 + c + 64 + lane);
          const float lo = (g.flags & CB_EPI_RELU) ? 0.f : -INFINITY;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float b0 = __shfl_sync(0xffffffffu, bl, j), b1 = __shfl_sync(0xffffffffu, bl, j + 1);
            pk[j >> 1] = pack_bf16(fmaxf(fmaf(__uint_as_float(r[j]), g.alpha, b0), lo), fmaxf(fmaf(__uint_as_float(r[j + 1]), g.alpha, b1), lo));
          }
          uint8_t* srow64 = slab + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(srow64 + ((k ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          __syncwarp();
          if (sl + 2 < n_slabs) tmem_ld32(t_addr + c + 64, r);   // next slab's accumulator, in flight during the stores below
          const int ch = lane & 3;
          const int gcol8 = n0 + c + ch * 8;
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.C) + gcol8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + (lane >> 2);
            const long grow = (long)m0 + q * 32 + rl;
            const uint4 val = *reinterpret_cast<const uint4*>(slab + rl * 64 + ((ch ^ ((rl >> 1) & 3)) << 4));
            if (grow < g.M && gcol8 < g.N) *reinterpret_cast<uint4*>(dst + grow * g.ldc) = val;
          }
          __syncwarp();
          continue;
        }
        if (mode == EM_BF16_MASK) {
          // d(hidden) = (dz2 . W2) masked by hidden > 0: the product is rounded to bf16 in the row layout, staged as 64 B
          // rows, and masked in the coalesced read-back with 16-byte mask loads.
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) pk[j >> 1] = pack_bf16(__uint_as_float(r[j]) * g.alpha, __uint_as_float(r[j + 1]) * g.alpha);
          uint8_t* srow64 = slab + lane * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(srow64 + ((k ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          __syncwarp();
          if (sl + 2 < n_slabs) tmem_ld32(t_addr + c + 64, r);
          const int ch = lane & 3;
          const int gcol8 = n0 + c + ch * 8;
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.C) + gcol8;
          float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + (lane >> 2);
            const long grow = (long)m0 + q * 32 + rl;
            uint4 val = *reinterpret_cast<const uint4*>(slab + rl * 64 + ((ch ^ ((rl >> 1) & 3)) << 4));
            uint32_t* vv = &val.x;
            const uint32_t* mm = &mk[i].x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // bf16 > 0  <=>  its bit pattern, read as a signed 16-bit integer, is > 0
              const uint32_t sel = ((short)(mm[k] & 0xffffu) > 0 ? 0xffffu : 0u) | ((int)mm[k] >= 0x10000 ? 0xffff0000u : 0u);
              vv[k] &= sel;
              const float2 f = unpack_bf16(vv[k]);
              cs[2 * k] += f.x; cs[2 * k + 1] += f.y;
            }
            if (grow < g.M && gcol8 < g.N) *reinterpret_cast<uint4*>(dst + grow * g.ldc) = val;
          }
          if (g.colsum) {   // fused linear1 bias gradient: fold the 8 row-lanes that share a column chunk, one vector RED pair
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4); cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
              cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
            }
            if (lane < 4 && gcol8 < g.N) {
              atomicAdd(reinterpret_cast<float4*>(g.colsum + gcol8), make_float4(cs[0], cs[1], cs[2], cs[3]));
              atomicAdd(reinterpret_cast<float4*>(g.colsum + gcol8 + 4), make_float4(cs[4], cs[5], cs[6], cs[7]));
            }
          }
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<uint4*>(srow + ((k ^ (lane & 7)) << 4)) = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        __syncwarp();
        if (sl + 2 < n_slabs) tmem_ld32(t_addr + c + 64, r);   // next slab's accumulator, in flight during the math below
        // ---- read-back + epilogue math: this lane owns columns gcol..gcol+3 of rows rb_row, rb_row+4, ...
        // One lean, branch-free loop per epilogue mode (compile-time).
        float4 acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + rb_row;
          acc[i] = *reinterpret_cast<const float4*>(slab + rl * 128 + ((rb_chunk ^ (rl & 7)) << 4));
        }
        __syncwarp();   // slab may be overwritten by the next iteration from here on
        if (g.alpha != 1.f) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc[i].x *= g.alpha; acc[i].y *= g.alpha; acc[i].z *= g.alpha; acc[i].w *= g.alpha; }
        }
        if (!col_ok || nrows <= 0) continue;
        if (mode == EM_BF16) {
          const float lo = (g.flags & CB_EPI_RELU) ? 0.f : -INFINITY;
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.C) + row0 * g.ldc + gcol;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < nrows) {
              const float4 v = acc[i];
              *reinterpret_cast<uint2*>(dst + (long)i * 4 * g.ldc) =
                  make_uint2(pack_bf16(fmaxf(v.x + bias4.x, lo), fmaxf(v.y + bias4.y, lo)), pack_bf16(fmaxf(v.z + bias4.z, lo), fmaxf(v.w + bias4.w, lo)));
            }
          }
        } else if (mode == EM_F32) {
          float* dst = reinterpret_cast<float*>(g.C) + row0 * g.ldc + gcol;
          const float lo = (g.flags & CB_EPI_RELU) ? 0.f : -INFINITY;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < nrows) {
              const float4 v = acc[i];
              *reinterpret_cast<float4*>(dst + (long)i * 4 * g.ldc) =
                  make_float4(fmaxf(v.x + bias4.x, lo) + res[i].x, fmaxf(v.y + bias4.y, lo) + res[i].y, fmaxf(v.z + bias4.z, lo) + res[i].z,
                              fmaxf(v.w + bias4.w, lo) + res[i].w);
            }
          }
        } else if (mode == EM_ATOMIC) {
          float* dst = reinterpret_cast<float*>(g.C) + row0 * g.ldc + gcol;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < nrows) {
              const float4 v = acc[i];
              atomicAdd(reinterpret_cast<float4*>(dst + (long)i * 4 * g.ldc), make_float4(v.x + bias4.x, v.y + bias4.y, v.z + bias4.z, v.w + bias4.w));
            }
          }
        } else {  // EM_TOKENIZE: fp32 tokens = acc + conv bias + pos[p] + channel_token[c]; CLS rows = cls_token + pos_embed[0]
          float* dst = reinterpret_cast<float*>(g.C) + row0 * g.ldc + gcol;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int px = __shfl_sync(0xffffffffu, ri.x, i * 4 + rb_row), cx = __shfl_sync(0xffffffffu, ri.y, i * 4 + rb_row);
            if (i < nrows) {
              float4 v = acc[i];
              if (px == -2) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(g.cls_tok + gcol)), b4 = __ldg(reinterpret_cast<const float4*>(g.pos0 + gcol));
                v = make_float4(a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w);
              } else {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(g.pos + px + gcol));
                v.x += bias4.x + p4.x; v.y += bias4.y + p4.y; v.z += bias4.z + p4.z; v.w += bias4.w + p4.w;
                if (cx >= 0) {
                  const float4 c4 = __ldg(reinterpret_cast<const float4*>(g.chan_tok + cx + gcol));
                  v.x += c4.x; v.y += c4.y; v.z += c4.z; v.w += c4.w;
                }
              }
              *reinterpret_cast<float4*>(dst + (long)i * 4 * g.ldc) = v;
            }
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GEMM_W_MMA) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host
static int encode_operand(CUtensorMap* tm, const void* base, int rows, int K, int ld, int mode, int tile_rows) {
  if (mode == 0) {  // [rows, K] row-major, inner = K
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
    uint64_t strides[1] = {(uint64_t)ld * 2};
    uint32_t box[2] = {64, (uint32_t)tile_rows};
    return make_tmap(tm, base, 2, dims, strides, box, 3);
  }
  const int blk = mode == 1 ? 64 : 32;  // [K, rows] row-major viewed as (blk, K, rows/blk)
  uint64_t dims[3] = {(uint64_t)blk, (uint64_t)K, (uint64_t)(rows / blk)};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)blk * 2};
  uint32_t box[3] = {(uint32_t)blk, 64, (uint32_t)(tile_rows / blk)};
  return make_tmap(tm, base, 3, dims, strides, box, mode == 1 ? 3 : 2);
}

template <int BN, int AMODE, int BMODE, int EM, bool WRES, bool CSUM = false, bool LNF = false>
static int launch_k(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  const int kb_total = (g.K + BK - 1) / BK;
  constexpr int EPI_SZ = use_tstore<BN, EM, WRES, CSUM, LNF>() ? TSTORE_BYTES + 1024 : EPI_BYTES;
  CUtensorMap tmC = tmA;          // placeholder for the kernels that do not store through TMA
  if (use_tstore<BN, EM, WRES, CSUM, LNF>()) {   // bf16 C [M, N] (ldc): boxes of 64 columns x 128 rows, 128B swizzle
    uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
    uint64_t strides[1] = {(uint64_t)g.ldc * 2};
    uint32_t box[2] = {64, 128};
    if (make_tmap(&tmC, g.C, 2, dims, strides, box, 3)) return 1;
  }
  const int smem_bytes = WRES ? kb_total * Cfg::B_BYTES + WRES_A_STAGES * kb_total * Cfg::A_BYTES + EPI_SZ + 1024 + 256
                              : Cfg::SMEM_BYTES + (CSUM ? Cfg::STAGES * CSUM_ONES_BYTES : 0);
  static_assert(!CSUM || Cfg::SMEM_BYTES + Cfg::STAGES * CSUM_ONES_BYTES <= 227 * 1024, "CSUM: ones blocks do not fit");
  static int attr_set = 0;
  if (attr_set < smem_bytes) {
    CB_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, AMODE, BMODE, EM, WRES, CSUM, LNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = smem_bytes;
  }
  const int num_m = (g.M + BM - 1) / BM, num_n = (g.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n * g.k_splits;
  int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  if (WRES) {   // a multiple of the column-tile count: CTA c owns column tile c % num_n
    const int per_n = max(1, min(num_sms() / num_n, num_m));
    grid = per_n * num_n;
  }
  gemm_kernel<BN, AMODE, BMODE, EM, WRES, CSUM, LNF><<<grid, GEMM_THREADS, smem_bytes, stream>>>(tmA, tmB, tmC, g);
  CB_CUDA(cudaGetLastError());
  return 0;
}

// weights-resident variant: A K-major, K <= 192, no split-K, bf16 / masked-bf16 / fp32(+residual) epilogues
template <int BN, int AMODE, int BMODE, int EM>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  if constexpr (AMODE == 0 && (EM == EM_BF16 || EM == EM_BF16_MASK || EM == EM_F32)) {
    const int kb_total = (g.K + BK - 1) / BK;
    constexpr bool TS = use_tstore<BN, EM, true, false, false>();
    const bool fits = kb_total * Cfg::B_BYTES + WRES_A_STAGES * kb_total * Cfg::A_BYTES + (TS ? TSTORE_BYTES + 1024 : EPI_BYTES) + 1024 + 256 <= 227 * 1024;
    const int num_n = (g.N + BN - 1) / BN;
    // the TMA-store epilogue needs whole 64-column boxes and the aligned rows of the row-direct epilogue
    const bool ts_ok = !TS || (g.direct && g.N % 192 == 0 && g.ldc % 8 == 0);
    if (g.k_splits == 1 && kb_total <= 3 && fits && ts_ok && num_n <= num_sms() && g.M >= 4 * BM) return launch_k<BN, AMODE, BMODE, EM, true>(tmA, tmB, g, stream);
  }
  return launch_k<BN, AMODE, BMODE, EM, false>(tmA, tmB, g, stream);
}

// Instantiated (operand layout, epilogue) pairs: forward products are K-major x K-major; input gradients read the weight
// MN-major; weight gradients read both operands MN-major and accumulate with fp32 atomics (or store fp32 for the head).
template <int BN>
static int dispatch_modes(int am, int bm, int em, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t s) {
#define CB_CASE(A_, B_, E_) if (am == A_ && bm == B_ && em == E_) return launch<BN, A_, B_, E_>(tmA, tmB, g, s);
  CB_CASE(0, 0, EM_BF16) CB_CASE(0, 0, EM_F32) CB_CASE(0, 0, EM_TOKENIZE) CB_CASE(0, 0, EM_BF16_MASK) CB_CASE(0, 0, EM_ATOMIC)
  CB_CASE(0, 1, EM_BF16) CB_CASE(0, 1, EM_BF16_MASK) CB_CASE(0, 1, EM_F32)
  CB_CASE(0, 2, EM_BF16) CB_CASE(0, 2, EM_BF16_MASK) CB_CASE(0, 2, EM_F32)
  CB_CASE(1, 1, EM_ATOMIC) CB_CASE(1, 1, EM_F32) CB_CASE(1, 2, EM_ATOMIC) CB_CASE(1, 2, EM_F32)
  CB_CASE(2, 1, EM_ATOMIC) CB_CASE(2, 1, EM_F32) CB_CASE(2, 2, EM_ATOMIC) CB_CASE(2, 2, EM_F32)
#undef CB_CASE
  set_error("gemm: operand layout (a_mn=%d, b_mn=%d) with epilogue mode %d is not instantiated", am, bm, em);
  return 1;
}

int gemm_run(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, GemmArgs g, cudaStream_t stream) {
  CB_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  CB_CHECK(g.N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && g.ldc % 8 == 0, "gemm: N, lda, ldb, ldc must be multiples of 8 (N=%d lda=%d ldb=%d ldc=%d)", g.N, lda, ldb, g.ldc);
  CB_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0,
           "gemm: operands must be 16-byte aligned");
  CB_CHECK(g.C != nullptr || g.ln_gamma != nullptr, "gemm: output pointer is null");
  int am = 0, bm = 0;
  if (a_mn) { CB_CHECK(g.M % 32 == 0, "gemm: MN-major A needs M %% 32 == 0 (M=%d)", g.M); am = (g.M % 64 == 0) ? 1 : 2; }
  if (b_mn) { CB_CHECK(g.N % 32 == 0, "gemm: MN-major B needs N %% 32 == 0 (N=%d)", g.N); bm = (g.N % 64 == 0) ? 1 : 2; }
  const int BN = (g.N % 192 == 0 && g.N % 256 != 0) ? 192 : (g.N >= 256 ? 256 : 128);
  if (g.k_splits < 1) g.k_splits = 1;
  const int kb_total = (g.K + BK - 1) / BK;
  if (g.k_splits > kb_total) g.k_splits = kb_total;
  if (g.k_splits > 1) {
    // every split must own at least one k-block, otherwise its tile would add garbage
    const int kb_per = (kb_total + g.k_splits - 1) / g.k_splits;
    g.k_splits = (kb_total + kb_per - 1) / kb_per;
    CB_CHECK(g.flags & CB_EPI_ATOMIC, "gemm: split-K requires the atomic epilogue");
  }
  {
    const int esz = (g.flags & (CB_EPI_OUT_F32 | CB_EPI_ATOMIC)) ? 4 : 2;
    const bool c_ok = (reinterpret_cast<uintptr_t>(g.C) & 31) == 0 && ((long)g.ldc * esz) % 32 == 0 && g.N % 8 == 0;
    const int asz = (g.flags & CB_EPI_RESIDUAL_F32) ? 4 : 2;
    const bool aux_ok = !(g.flags & (CB_EPI_RESIDUAL_F32 | CB_EPI_RELU_MASK)) || (g.flags & CB_EPI_MASK_BITS) ||
                        ((reinterpret_cast<uintptr_t>(g.aux) & 31) == 0 && ((long)g.ld_aux * asz) % 32 == 0);
    g.direct = (c_ok && aux_ok) ? 1 : 0;
  }
  CUtensorMap tmA, tmB;
  if (encode_operand(&tmA, A, g.M, g.K, lda, am, BM)) return 1;
  if (encode_operand(&tmB, B, g.N, g.K, ldb, bm, BN)) return 1;
  const int em = (g.flags & CB_EPI_TOKENIZE) ? EM_TOKENIZE : (g.flags & CB_EPI_ATOMIC) ? EM_ATOMIC : (g.flags & CB_EPI_OUT_F32) ? EM_F32
                 : (g.flags & CB_EPI_RELU_MASK) ? EM_BF16_MASK : EM_BF16;
  if (g.ln_gamma != nullptr) {   // LayerNorm in the row epilogue (see LNF)
    const int kbt = (g.K + BK - 1) / BK;
    CB_CHECK(g.N == 192 && BN == 192 && am == 0 && bm == 0 && em == EM_F32 && g.k_splits == 1 && kbt <= 3 && g.M >= 4 * BM && g.direct,
             "gemm_ln: needs N = 192, K <= 192, M >= %d, K-major operands, 32-byte aligned rows (N=%d K=%d M=%d direct=%d)", 4 * BM, g.N, g.K, g.M, g.direct);
    return launch_k<192, 0, 0, EM_F32, true, false, true>(tmA, tmB, g, stream);
  }
  if (g.colsum != nullptr && em == EM_ATOMIC) {   // bias gradient on the tensor pipe (see CSUM above)
    CB_CHECK(BN == 192 && am == 1 && bm == 1, "gemm: colsum with the atomic epilogue needs N %% 192 == 0 (N %% 256 != 0) and both operands MN-major "
             "with M, N multiples of 64 (N=%d M=%d a_mn=%d b_mn=%d)", g.N, g.M, a_mn, b_mn);
    return launch_k<192, 1, 1, EM_ATOMIC, false, true>(tmA, tmB, g, stream);
  }
  if (BN == 192) return dispatch_modes<192>(am, bm, em, tmA, tmB, g, stream);
  if (BN == 256) return dispatch_modes<256>(am, bm, em, tmA, tmB, g, stream);
  return dispatch_modes<128>(am, bm, em, tmA, tmB, g, stream);
}

}  // namespace cb

#ifdef CB_TIMELINE
extern "C" int cb_debug_timeline_gemm(void* dst) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif

extern "C" int cb_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M,
                            int N, int K, const float* bias, const void* aux, int ld_aux, int flags, float alpha,
                            int k_splits, float* colsum, void* stream) {
  cb::GemmArgs g{};
  g.colsum = colsum;
  CB_CHECK(!colsum || (((flags & CB_EPI_RELU_MASK) != 0) != ((flags & CB_EPI_ATOMIC) != 0) && !(flags & CB_EPI_OUT_F32)),
           "cb_gemm_bf16: colsum goes with the ReLU-mask epilogue (column sums of C, [N]) or with the atomic epilogue (sums of op(A) over K, [M])");
  g.M = M; g.N = N; g.K = K; g.k_splits = k_splits; g.C = C; g.ldc = ldc; g.bias = bias;
  g.aux = reinterpret_cast<const __nv_bfloat16*>(aux); g.ld_aux = ld_aux; g.flags = flags; g.alpha = alpha;
  CB_CHECK(!(flags & CB_EPI_TOKENIZE), "cb_gemm_bf16: use cb_tokenize_fwd for the tokenizer epilogue");
  CB_CHECK(!(flags & CB_EPI_RESIDUAL), "cb_gemm_bf16: bf16 residual epilogue was removed (the residual stream is fp32: use CB_EPI_RESIDUAL_F32)");
  CB_CHECK(!(flags & CB_EPI_RESIDUAL_F32) || (flags & CB_EPI_OUT_F32), "cb_gemm_bf16: CB_EPI_RESIDUAL_F32 requires CB_EPI_OUT_F32");
  CB_CHECK(!(flags & (CB_EPI_RELU_MASK | CB_EPI_RESIDUAL_F32)) || aux, "cb_gemm_bf16: aux pointer required by flags");
  CB_CHECK(!(flags & CB_EPI_MASK_BITS) || ((flags & CB_EPI_RELU_MASK) && ld_aux >= M && (reinterpret_cast<uintptr_t>(aux) & 3) == 0),
           "cb_gemm_bf16: CB_EPI_MASK_BITS needs CB_EPI_RELU_MASK and a uint32 [N/32, ld_aux >= M] bit mask");
  return cb::gemm_run(A, lda, a_mn, B, ldb, b_mn, g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int cb_gemm_ln_fwd(const void* A, int lda, const void* W, int ldw, const float* bias, const float* resid, int ld_res,
                              const float* ln_gamma, const float* ln_beta, float ln_eps, float* z, void* y_bf16, float* y_f32, float* mean,
                              float* rstd, int M, int N, int K, void* stream) {
  CB_CHECK(A && W && ln_gamma && ln_beta && y_bf16, "cb_gemm_ln_fwd: null argument");
  CB_CHECK(N == 192 && K % 8 == 0 && K <= 192 && M >= 512, "cb_gemm_ln_fwd: N must be 192, K <= 192, M >= 512 (N=%d K=%d M=%d); use cb_gemm_bf16 + cb_layernorm_fwd otherwise", N, K, M);
  CB_CHECK(((reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(y_f32) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(resid)) & 31) == 0 &&
           (resid == nullptr || ld_res % 8 == 0), "cb_gemm_ln_fwd: z / y / resid must be 32-byte aligned (resid pitch a multiple of 8 floats)");
  cb::GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.k_splits = 1; g.C = z; g.ldc = N; g.bias = bias;
  g.aux = reinterpret_cast<const __nv_bfloat16*>(resid); g.ld_aux = ld_res; g.alpha = 1.f;
  g.flags = CB_EPI_OUT_F32 | (resid ? CB_EPI_RESIDUAL_F32 : 0);
  g.ln_gamma = ln_gamma; g.ln_beta = ln_beta; g.ln_eps = ln_eps; g.ln_y16 = reinterpret_cast<__nv_bfloat16*>(y_bf16); g.ln_y32 = y_f32;
  g.ln_mean = mean; g.ln_rstd = rstd;
  return cb::gemm_run(A, lda, 0, W, ldw, 0, g, reinterpret_cast<cudaStream_t>(stream));
}
