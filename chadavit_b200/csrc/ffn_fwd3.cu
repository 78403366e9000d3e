// Fused feed-forward block, third generation (same contract as ffn_fwd.cu: z2 = resid + relu(y W1^T + b1) W2^T + b2, D = 192).
//
// An earlier cluster kernel (two CTAs, one row tile per CTA, multicast weight halves, 64-unit chunks; removed) showed that a tcgen05.mma with N = 64 costs
// ~53 clocks whatever its operands (gpurun_out/tl_ffn5.txt: 12 H MMAs = 640 clk, while the 4 Z MMAs with N = 192 take their
// nominal 4 x 96) — the H products were paying a per-instruction floor.  Here the hidden dimension is walked in chunks of 128:
//     H(c)  = y_t · W1[c]^T        12 SS MMAs, M = 128, N = 128, K = 16 (nominal 64 clk each)       -> TMEM (128 fp32 columns)
//     P(c)  = bf16(relu(H(c) + b1[c]))   8 warps, written over H(c) (64 columns)
//     Z    += P(c) · W2[:, c]^T     8 TS MMAs, N = 192
// TMEM: H0 128 | H1 128 | Z 192 = 448 columns (y cannot also live in tensor memory: 96 more), so y_t comes from shared memory
// (one TMA tile per item) and the two hidden buffers alternate: the tensor pipe runs Z(c), H(c+2) while the epilogue warps turn
// H(c+1) into P(c+1).  A cluster of two CTAs shares every weight chunk (either CTA fetches half of
// W1[c] / W2[:, c] and multicasts it), so the L2 -> SM weight stream is one fetch per 256 rows.
// Shared memory: y 48 KB | W1 ring 2 x 48 KB | W2 ring 3 x 24 KB (64-unit k-blocks) | b1 8 KB.
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

namespace f3 {
constexpr int D = 192, C = 128, KB = D / 64;
constexpr int Y_BYTES = 128 * D * 2;     // y tile: KB blocks of [128 x 64] bf16, 128B swizzle (16 KB each)
constexpr int W1_BYTES = C * D * 2;      // W1 chunk: KB blocks of [128 x 64] (16 KB each)
constexpr int W2_BYTES = D * 64 * 2;     // W2 k-block: [192 x 64] (two per chunk)
constexpr int S1 = 2, S2 = 3;            // ring depths (W1 chunks / W2 k-blocks)
constexpr int NB = 2;                    // hidden-chunk buffers in TMEM
constexpr int MAX_F = 2048;
constexpr int SMEM_BYTES = Y_BYTES + S1 * W1_BYTES + S2 * W2_BYTES + MAX_F * 4 + 1024 /*align*/ + 512 /*barriers*/;
constexpr int COL_H = 0, COL_Z = 256;    // H_b at 128 b; Z: 192 columns
}  // namespace f3

struct Ffn3Args {
  const float* b1;         // [F]
  const float* b2;         // [D]
  const float* resid;      // [T, D] fp32
  float* z2;               // [T, D] fp32
  __nv_bfloat16* hid;      // [T, F] bf16 or null
  uint32_t* mask_bits;     // [F/32, ld_bits] ReLU mask as bits, or null
  int ld_bits;
  int T, F;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
ffn_fwd3_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                const Ffn3Args a) {
  using namespace f3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sY = smem;
  uint8_t* sW1 = sY + Y_BYTES;
  uint8_t* sW2 = sW1 + S1 * W1_BYTES;
  float* sB1 = reinterpret_cast<float*>(sW2 + S2 * W2_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + MAX_F);
  uint64_t* y_full = bars;                  // tx: the item's y tile
  uint64_t* y_empty = y_full + 1;           // every H product of the item has retired
  uint64_t* w1_full = y_empty + 1;          // [S1] tx: both halves of the chunk (own load + the peer's multicast)
  uint64_t* w1_empty = w1_full + S1;        // [S1] 2 arrivals: the MMA warps of both CTAs (multicast commit)
  uint64_t* w2_full = w1_empty + S1;        // [S2]
  uint64_t* w2_empty = w2_full + S2;        // [S2] 2 arrivals
  uint64_t* h_full = w2_empty + S2;         // [NB] H(c) complete
  uint64_t* p_full = h_full + NB;           // [NB] 256 arrivals: P(c) written
  uint64_t* p_half = p_full + NB;           // [NB] 128 arrivals: column half 0 has read its fp32 input
  uint64_t* z_full = p_half + NB;
  uint64_t* z_empty = z_full + 1;           // 256 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_empty + 1);

  constexpr int W_MMA = 8, W_TMA = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_tiles = (a.T + 127) / 128, n_items = (n_tiles + 1) / 2, n_chunks = a.F / C;
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(y_full, 1); mbar_init(y_empty, 1);
    for (int i = 0; i < S1; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 2); }
    for (int i = 0; i < S2; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 2); }
    for (int i = 0; i < NB; ++i) { mbar_init(&h_full[i], 1); mbar_init(&p_full[i], 256); mbar_init(&p_half[i], 128); }
    mbar_init(z_full, 1); mbar_init(z_empty, 256);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < a.F; i += blockDim.x) sB1[i] = __ldg(a.b1 + i);
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ------------------------------------------------------------------ TMA producer: own y tile; this CTA's half of every weight chunk, to both CTAs
    if (lane == 0) {
      int s1 = 0, s2 = 0; uint32_t p1 = 0, p2 = 0, ni = 0;
      auto load_w1 = [&](int c) {
        mbar_wait(&w1_empty[s1], p1 ^ 1);
        mbar_expect_tx(&w1_full[s1], W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)      // rows [64 rank, 64 rank + 64) of the 128-row chunk: 8 KB per k-block
          tma_load_2d_mc(sW1 + s1 * W1_BYTES + kb * (C * 128) + rank * 8192, &tmW1, &w1_full[s1], kb * 64, c * C + rank * 64, 0x3);
        if (++s1 == S1) { s1 = 0; p1 ^= 1; }
      };
      auto load_w2 = [&](int kblk) {         // k-block kblk = 64 hidden units: rows [96 rank, +96) of W2[:, 64 kblk ..]
        mbar_wait(&w2_empty[s2], p2 ^ 1);
        mbar_expect_tx(&w2_full[s2], W2_BYTES);
        tma_load_2d_mc(sW2 + s2 * W2_BYTES + rank * (96 * 128), &tmW2, &w2_full[s2], kblk * 64, rank * 96, 0x3);
        if (++s2 == S2) { s2 = 0; p2 ^= 1; }
      };
      for (int it = cluster_id; it < n_items; it += n_clusters, ++ni) {
        mbar_wait(y_empty, (ni & 1) ^ 1);
        mbar_expect_tx(y_full, Y_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sY + kb * (128 * 128), &tmY, y_full, kb * 64, (2 * it + (int)rank) * 128);
        for (int c = 0; c < NB && c < n_chunks; ++c) load_w1(c);
        for (int c = 0; c < n_chunks; ++c) {
          load_w2(2 * c);
          load_w2(2 * c + 1);
          if (c + NB < n_chunks) load_w1(c + NB);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane)
    constexpr uint32_t idesc_h = umma_idesc_bf16(128, C, false, false);
    constexpr uint32_t idesc_z = umma_idesc_bf16(128, D, false, false);
    const uint64_t y_desc0 = umma_smem_desc(smem_u32(sY), 16, 1024, 3);
    const uint64_t w1_desc0 = umma_smem_desc(smem_u32(sW1), 16, 1024, 3), w2_desc0 = umma_smem_desc(smem_u32(sW2), 16, 1024, 3);
    int s1 = 0, s2 = 0; uint32_t p1 = 0, p2 = 0, ni = 0, g = 0;    // g: chunks issued so far (buffer = g % NB, use = g / NB)
    CB_TL_DECL(tl);
    auto issue_h = [&](uint32_t gi, bool last_of_item) {   // H(chunk gi) = y · W1[c]^T into buffer gi % NB; whole warp
      const uint32_t b = gi % NB;
      mbar_wait(&w1_full[s1], p1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t wd = umma_desc_add(w1_desc0, s1 * W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + COL_H + b * C, umma_desc_add(y_desc0, kb * (128 * 128) + k * 32), umma_desc_add(wd, kb * (C * 128) + k * 32), idesc_h,
                    (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(&h_full[b]);
        tc_commit_mc(&w1_empty[s1], 0x3);
        if (last_of_item) tc_commit(y_empty);
      }
      __syncwarp();
      if (++s1 == S1) { s1 = 0; p1 ^= 1; }
    };
    for (int it = cluster_id; it < n_items; it += n_clusters, ++ni) {
      mbar_wait(y_full, ni & 1);
      tc_fence_after();
      for (int c = 0; c < NB && c < n_chunks; ++c) issue_h(g + c, c + 1 == n_chunks);
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t gi = g + c, b = gi % NB;
        CB_TL(0, tl, 1);
        const int za = s2; mbar_wait(&w2_full[s2], p2); if (++s2 == S2) { s2 = 0; p2 ^= 1; }
        const int zb = s2; mbar_wait(&w2_full[s2], p2); if (++s2 == S2) { s2 = 0; p2 ^= 1; }
        CB_TL(0, tl, 2);
        mbar_wait(&p_full[b], (gi / NB) & 1);
        CB_TL(0, tl, 3);
        if (c == 0 && ni > 0) mbar_wait(z_empty, (ni - 1) & 1);    // the previous item's Z has been read out
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wa = umma_desc_add(w2_desc0, za * W2_BYTES), wb = umma_desc_add(w2_desc0, zb * W2_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)           // Z += P(c)[:, 0:64] · W2[:, 128c .. +64]^T
            umma_ts(tmem_base + COL_Z, tmem_base + COL_H + b * C + kk * 8, umma_desc_add(wa, kk * 32), idesc_z, (c > 0 || kk > 0) ? 1u : 0u);
          tc_commit_mc(&w2_empty[za], 0x3);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)           // Z += P(c)[:, 64:128] · W2[:, 128c + 64 .. +64]^T
            umma_ts(tmem_base + COL_Z, tmem_base + COL_H + b * C + 32 + kk * 8, umma_desc_add(wb, kk * 32), idesc_z, 1u);
          tc_commit_mc(&w2_empty[zb], 0x3);
          if (c + 1 == n_chunks) tc_commit(z_full);
        }
        __syncwarp();
        CB_TL(0, tl, 4);
        if (c + NB < n_chunks) issue_h(gi + NB, c + NB + 1 == n_chunks);   // over P(c): the in-order pipe has retired its reader by then
        CB_TL(0, tl, 5);
      }
      g += n_chunks;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: lane quarter q, column half hf (64 of the 128 hidden units)
    const int hf = warp >> 2, q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    uint32_t ni = 0, g = 0;
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4) && lane == 0;
    for (int it = cluster_id; it < n_items; it += n_clusters, ++ni) {
      const long row = (long)(2 * it + rank) * 128 + r_in_tile;
      const bool row_ok = row < a.T;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t gi = g + c, b = gi % NB, ph = (gi / NB) & 1;
        if (tl_on) CB_TL(1 + hf, tl, 1);
        mbar_wait(&h_full[b], ph);
        tc_fence_after();
        if (tl_on) CB_TL(1 + hf, tl, 2);
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t r0[32];
          tmem_ld32(lane_addr + COL_H + b * C + hf * 64 + j * 32, r0);
          tmem_ld_wait();
          const float* bp = sB1 + c * C + hf * 64 + j * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 ba = *reinterpret_cast<const float4*>(bp + e);
            pk[j * 16 + (e >> 1)] = pack_bf16(fmaxf(__uint_as_float(r0[e]) + ba.x, 0.f), fmaxf(__uint_as_float(r0[e + 1]) + ba.y, 0.f));
            pk[j * 16 + (e >> 1) + 1] = pack_bf16(fmaxf(__uint_as_float(r0[e + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(r0[e + 3]) + ba.w, 0.f));
          }
        }
        // P (bf16) of hidden units 64 hf .. 64 hf + 63 -> columns 32 hf .. 32 hf + 31 of the buffer; half 1's target columns
        // (32..63) are part of half 0's fp32 input (0..63): it waits until half 0 has read them
        if (hf == 1) { mbar_wait(&p_half[b], ph); tc_fence_after(); }
        else { tc_fence_before(); mbar_arrive_relaxed(&p_half[b]); }
        {
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { lo[e] = pk[e]; hi[e] = pk[16 + e]; }
          tmem_st16(lane_addr + COL_H + b * C + hf * 32, lo);
          tmem_st16(lane_addr + COL_H + b * C + hf * 32 + 16, hi);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_relaxed(&p_full[b]);       // relaxed: must not wait for the hidden-activation stores of the previous chunk
        if (tl_on) CB_TL(1 + hf, tl, 3);
        if (a.hid && row_ok) {                       // 128 contiguous bytes per thread
          __nv_bfloat16* dst = a.hid + row * a.F + c * C + hf * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            stg256(dst + 16 * k, pk[8 * k], pk[8 * k + 1], pk[8 * k + 2], pk[8 * k + 3], pk[8 * k + 4], pk[8 * k + 5], pk[8 * k + 6], pk[8 * k + 7]);
        }
        if (a.mask_bits && row_ok) {
          a.mask_bits[(long)(4 * c + 2 * hf) * a.ld_bits + row] = relu_bits16(pk);
          a.mask_bits[(long)(4 * c + 2 * hf + 1) * a.ld_bits + row] = relu_bits16(pk + 16);
        }
      }
      g += n_chunks;
      // ---- final epilogue: z2 = Z + b2 + resid (fp32); this warp's half: 32-column slabs hf, hf + 2, hf + 4
      mbar_wait(z_full, ni & 1);
      tc_fence_after();
#pragma unroll 1
      for (int s = hf; s < D / 32; s += 2) {
        uint32_t res[4][8];
        if (row_ok) {
          const float* rp = a.resid + row * D + s * 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256(rp + 8 * k, res[k]);
        }
        uint32_t x[32];
        tmem_ld32(lane_addr + COL_Z + s * 32, x);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = a.z2 + row * D + s * 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.b2 + s * 32 + 8 * k)), b1 = __ldg(reinterpret_cast<const float4*>(a.b2 + s * 32 + 8 * k + 4));
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(x[8 * k + e]) + bb[e] + __uint_as_float(res[k][e]));
            stg256(dst + 8 * k, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(z_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

}  // namespace cb

#ifdef CB_TIMELINE
extern "C" int cb_debug_timeline_ffn3(void* dst) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif

namespace cb {

int ffn_fwd3_run(const void* y, const void* w1, const float* b1, const void* w2, const float* b2, const float* resid, float* z2, void* hid,
                 unsigned int* mask_bits, int ld_bits, int T, int F, cudaStream_t stream) {
  using namespace f3;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(ffn_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ty, t1, t2;
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)T}; uint64_t strides[1] = {(uint64_t)D * 2}; uint32_t box[2] = {64, 128};
    if (make_tmap(&ty, y, 2, dims, strides, box, 3)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)F}; uint64_t strides[1] = {(uint64_t)D * 2}; uint32_t box[2] = {64, 64};
    if (make_tmap(&t1, w1, 2, dims, strides, box, 3)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)F, (uint64_t)D}; uint64_t strides[1] = {(uint64_t)F * 2}; uint32_t box[2] = {64, 96};
    if (make_tmap(&t2, w2, 2, dims, strides, box, 3)) return 1;
  }
  Ffn3Args a{};
  a.b1 = b1; a.b2 = b2; a.resid = resid; a.z2 = z2; a.hid = reinterpret_cast<__nv_bfloat16*>(hid); a.mask_bits = mask_bits; a.ld_bits = ld_bits; a.T = T; a.F = F;
  const int n_items = ((T + 127) / 128 + 1) / 2;
  const int max_clusters = num_sms() / 2;
  const int clusters = n_items < max_clusters ? n_items : max_clusters;
  ffn_fwd3_kernel<<<2 * clusters, 320, SMEM_BYTES, stream>>>(ty, t1, t2, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cb
