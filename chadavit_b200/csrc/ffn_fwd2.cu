// Fused feed-forward block, second generation (same contract as ffn_fwd.cu: z2 = resid + relu(y W1^T + b1) W2^T + b2, D = 192).
//
// What the timeline of the first kernel showed (profiles/r01_timeline_ffn.txt): with a PAIR of row tiles per CTA all 512 TMEM
// columns are taken by two H buffers and two Z accumulators, so every tile has ONE hidden-chunk buffer and its chain
// H(c) -> bias/ReLU -> Z += P(c) W2 -> H(c+1) is strictly serial; the issuing thread blocks while a 16-MMA group drains, and
// the H products (SS, N = 64) pay the shared-memory port for the 4 KB A operand of every instruction.  Here:
//   * a CLUSTER of two CTAs shares each weight chunk: either CTA fetches half of W1[c] / W2[:, c] and the TMA multicasts it into
//     both shared memories, so the L2 -> SM weight stream stays at one fetch per 256 rows while every CTA owns ONE 128-row tile;
//   * the row tile y_t lives in TENSOR MEMORY (bf16, 96 columns, written once per item by the epilogue warps), so H = y W1[c]^T is
//     a TS MMA: no A operand through the shared-memory port (nominal N/2 clocks per instruction) and no y tile in smem;
//   * with one tile per CTA there is room for THREE hidden-chunk buffers (TMEM: y 96 | H 3 x 64 | Z 192), so the tensor pipe
//     always has Z(c) and H(c+3) queued while the epilogue warps turn H(c+1), H(c+2) into P — the chain is a pipeline.
// Warps 0-7: bias + ReLU (+ hidden store), y -> TMEM, final epilogue; warp 8: MMA issuer; warp 9: TMA.
#include "common.cuh"
#include "chadavit_b200.h"
#include "internal.h"

namespace cb {

#ifdef CB_TIMELINE
static __device__ unsigned long long g_cb_timeline[CB_TL_ROLES][CB_TL_LEN];
#endif

namespace f2 {
constexpr int D = 192, C = 64, KB = D / 64;
constexpr int W1_BYTES = C * D * 2;      // W1 chunk: KB blocks of [64 x 64] bf16, 128B swizzle (8 KB each)
constexpr int W2_BYTES = D * C * 2;      // W2 chunk: [192 x 64]
constexpr int S1 = 4, S2 = 4;            // ring depths
constexpr int NB = 3;                    // hidden-chunk buffers in TMEM
constexpr int MAX_F = 2048;
constexpr int SMEM_BYTES = S1 * W1_BYTES + S2 * W2_BYTES + MAX_F * 4 + 1024 /*align*/ + 512 /*barriers*/;
constexpr int COL_Y = 0, COL_H = 128, COL_Z = 320;   // y: 96 columns; H_b at 128 + 64 b; Z: 192 columns -> 512
}  // namespace f2

struct Ffn2Args {
  const __nv_bfloat16* y;  // [T, D] bf16
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
ffn_fwd2_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const Ffn2Args a) {
  using namespace f2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + S1 * W1_BYTES;
  float* sB1 = reinterpret_cast<float*>(sW2 + S2 * W2_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + MAX_F);
  uint64_t* w1_full = bars;                 // [S1] tx: both halves of the chunk (own load + the peer's multicast)
  uint64_t* w1_empty = w1_full + S1;        // [S1] 2 arrivals: the MMA warps of both CTAs (multicast commit)
  uint64_t* w2_full = w1_empty + S1;        // [S2]
  uint64_t* w2_empty = w2_full + S2;        // [S2] 2 arrivals
  uint64_t* h_full = w2_empty + S2;         // [NB] H(c) complete
  uint64_t* p_full = h_full + NB;           // [NB] 256 arrivals: P(c) written
  uint64_t* p_half = p_full + NB;           // [NB] 128 arrivals: column half 0 has read its fp32 input
  uint64_t* y_ready = p_half + NB;          // 256 arrivals: y tile of the item is in tensor memory
  uint64_t* z_full = y_ready + 1;
  uint64_t* z_empty = z_full + 1;           // 256 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_empty + 1);

  constexpr int W_MMA = 8, W_TMA = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_tiles = (a.T + 127) / 128, n_items = (n_tiles + 1) / 2, n_chunks = a.F / C;
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int i = 0; i < S1; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 2); }
    for (int i = 0; i < S2; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 2); }
    for (int i = 0; i < NB; ++i) { mbar_init(&h_full[i], 1); mbar_init(&p_full[i], 256); mbar_init(&p_half[i], 128); }
    mbar_init(y_ready, 256); mbar_init(z_full, 1); mbar_init(z_empty, 256);
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
    // ------------------------------------------------------------------ TMA producer: this CTA's half of every chunk, to both CTAs
    if (lane == 0) {
      int s1 = 0, s2 = 0; uint32_t p1 = 0, p2 = 0;
      auto load_w1 = [&](int c) {
        mbar_wait(&w1_empty[s1], p1 ^ 1);
        mbar_expect_tx(&w1_full[s1], W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)      // rows [32 rank, 32 rank + 32) of the 64-row chunk: 4 KB per k-block
          tma_load_2d_mc(sW1 + s1 * W1_BYTES + kb * (C * 128) + rank * 4096, &tmW1, &w1_full[s1], kb * 64, c * C + rank * 32, 0x3);
        if (++s1 == S1) { s1 = 0; p1 ^= 1; }
      };
      auto load_w2 = [&](int c) {
        mbar_wait(&w2_empty[s2], p2 ^ 1);
        mbar_expect_tx(&w2_full[s2], W2_BYTES);
        tma_load_2d_mc(sW2 + s2 * W2_BYTES + rank * (96 * 128), &tmW2, &w2_full[s2], c * C, rank * 96, 0x3);   // rows [96 rank, +96) of W2[:, c]
        if (++s2 == S2) { s2 = 0; p2 ^= 1; }
      };
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        for (int c = 0; c < NB && c < n_chunks; ++c) load_w1(c);
        for (int c = 0; c < n_chunks; ++c) {
          load_w2(c);
          if (c + NB < n_chunks) load_w1(c + NB);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane)
    constexpr uint32_t idesc_h = umma_idesc_bf16(128, C, false, false);
    constexpr uint32_t idesc_z = umma_idesc_bf16(128, D, false, false);
    const uint64_t w1_desc0 = umma_smem_desc(smem_u32(sW1), 16, 1024, 3), w2_desc0 = umma_smem_desc(smem_u32(sW2), 16, 1024, 3);
    int s1 = 0, s2 = 0; uint32_t p1 = 0, p2 = 0, ni = 0, g = 0;    // g: chunks issued so far (buffer = g % NB, use = g / NB)
    CB_TL_DECL(tl);
    // Measured (gpurun_out/tl_ffn5/6.txt): the issuing lane blocks while its MMAs drain (the tensor-core queue is shallow), so
    // the pipe idles during everything else this warp does.  Batching two chunks per elected block (all waits first, 32 MMAs in
    // one go) was tried and is SLOWER (140 us vs 133 us at T = 68664): the waits then become real waits for P(c+1).
    auto issue_h = [&](uint32_t gi) {        // H(chunk gi) = y · W1[c]^T into buffer gi % NB; whole warp
      const uint32_t b = gi % NB;
      mbar_wait(&w1_full[s1], p1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t wd = umma_desc_add(w1_desc0, s1 * W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts(tmem_base + COL_H + b * C, tmem_base + COL_Y + (kb * 4 + k) * 8, umma_desc_add(wd, kb * (C * 128) + k * 32), idesc_h, (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(&h_full[b]);
        tc_commit_mc(&w1_empty[s1], 0x3);
      }
      __syncwarp();
      if (++s1 == S1) { s1 = 0; p1 ^= 1; }
    };
    for (int it = cluster_id; it < n_items; it += n_clusters, ++ni) {
      mbar_wait(y_ready, ni & 1);
      tc_fence_after();
      for (int c = 0; c < NB && c < n_chunks; ++c) issue_h(g + c);
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t gi = g + c, b = gi % NB;
        CB_TL(0, tl, 1);
        mbar_wait(&w2_full[s2], p2);
        CB_TL(0, tl, 2);
        mbar_wait(&p_full[b], (gi / NB) & 1);
        CB_TL(0, tl, 3);
        if (c == 0 && ni > 0) mbar_wait(z_empty, (ni - 1) & 1);    // the previous item's Z has been read out
        tc_fence_after();
        if (elect_one()) {
          const uint64_t w2d = umma_desc_add(w2_desc0, s2 * W2_BYTES);
#pragma unroll
          for (int kk = 0; kk < C / 16; ++kk)      // Z += P(c) · W2[:, c]^T
            umma_ts(tmem_base + COL_Z, tmem_base + COL_H + b * C + kk * 8, umma_desc_add(w2d, kk * 32), idesc_z, (c > 0 || kk > 0) ? 1u : 0u);
          tc_commit_mc(&w2_empty[s2], 0x3);
          if (c + 1 == n_chunks) tc_commit(z_full);
        }
        __syncwarp();
        if (++s2 == S2) { s2 = 0; p2 ^= 1; }
        CB_TL(0, tl, 4);
        if (c + NB < n_chunks) issue_h(gi + NB);   // over P(c): the in-order pipe has retired its reader by then
        CB_TL(0, tl, 5);
      }
      g += n_chunks;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: lane quarter q, column half hf
    const int hf = warp >> 2, q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    uint32_t ni = 0, g = 0;
    CB_TL_DECL(tl);
    const bool tl_on = (warp == 0 || warp == 4) && lane == 0;
    auto put_y = [&](int item) {             // row of y (bf16) -> tensor memory: this warp's half = 48 of the 96 columns
      const long row = (long)(2 * item + rank) * 128 + r_in_tile;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = (hf * 3 + j) * 16;   // 16 columns = 32 bf16 = 64 bytes
        uint32_t v[16];
        if (row < a.T) {
          uint32_t lo[8], hi[8];
          ldg256(a.y + row * D + col * 2, lo);
          ldg256(a.y + row * D + col * 2 + 16, hi);
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e] = lo[e]; v[8 + e] = hi[e]; }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0u;
        }
        tmem_st16(lane_addr + COL_Y + col, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(y_ready);
    };
    if (cluster_id < n_items) put_y(cluster_id);
    for (int it = cluster_id; it < n_items; it += n_clusters, ++ni) {
      const long row = (long)(2 * it + rank) * 128 + r_in_tile;
      const bool row_ok = row < a.T;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t gi = g + c, b = gi % NB, ph = (gi / NB) & 1;
        if (tl_on) CB_TL(1 + hf, tl, 1);
        mbar_wait(&h_full[b], ph);
        tc_fence_after();
        if (tl_on) CB_TL(1 + hf, tl, 2);
        uint32_t r0[32];
        tmem_ld32(lane_addr + COL_H + b * C + hf * 32, r0);
        tmem_ld_wait();
        uint32_t pk[16];
        const float* bp = sB1 + c * C + hf * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 ba = *reinterpret_cast<const float4*>(bp + e);
          pk[e >> 1] = pack_bf16(fmaxf(__uint_as_float(r0[e]) + ba.x, 0.f), fmaxf(__uint_as_float(r0[e + 1]) + ba.y, 0.f));
          pk[(e >> 1) + 1] = pack_bf16(fmaxf(__uint_as_float(r0[e + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(r0[e + 3]) + ba.w, 0.f));
        }
        // P (bf16) of hidden units 32 hf .. 32 hf + 31 -> columns 16 hf .. 16 hf + 15 of the buffer; half 1's target columns
        // are part of half 0's fp32 input: it waits until half 0 has read them
        if (hf == 1) { mbar_wait(&p_half[b], ph); tc_fence_after(); }
        else { tc_fence_before(); mbar_arrive(&p_half[b]); }
        tmem_st16(lane_addr + COL_H + b * C + hf * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[b]);
        if (tl_on) CB_TL(1 + hf, tl, 3);
        if (a.hid && row_ok) {
          __nv_bfloat16* dst = a.hid + row * a.F + c * C + hf * 32;
          stg256(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
          stg256(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
        }
        if (a.mask_bits && row_ok) a.mask_bits[(long)(2 * c + hf) * a.ld_bits + row] = relu_bits16(pk);
      }
      g += n_chunks;
      // every H product of this item has completed (h_full of its last chunk was seen): the y columns may take the next tile,
      // and the tensor pipe starts on its H(0..2) while Z of this item is read out below
      if (it + n_clusters < n_items) put_y(it + n_clusters);
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
extern "C" int cb_debug_timeline_ffn2(void* dst) {
  CB_CUDA(cudaDeviceSynchronize());
  CB_CUDA(cudaMemcpyFromSymbol(dst, cb::g_cb_timeline, sizeof(cb::g_cb_timeline)));
  static unsigned long long zeros[CB_TL_ROLES][CB_TL_LEN];
  CB_CUDA(cudaMemcpyToSymbol(cb::g_cb_timeline, zeros, sizeof(zeros)));
  return 0;
}
#endif

namespace cb {

int ffn_fwd2_run(const void* y, const void* w1, const float* b1, const void* w2, const float* b2, const float* resid, float* z2, void* hid,
                 unsigned int* mask_bits, int ld_bits, int T, int F, cudaStream_t stream) {
  using namespace f2;
  static bool attr_set = false;
  if (!attr_set) {
    CB_CUDA(cudaFuncSetAttribute(ffn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap t1, t2;
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)F}; uint64_t strides[1] = {(uint64_t)D * 2}; uint32_t box[2] = {64, 32};
    if (make_tmap(&t1, w1, 2, dims, strides, box, 3)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)F, (uint64_t)D}; uint64_t strides[1] = {(uint64_t)F * 2}; uint32_t box[2] = {64, 96};
    if (make_tmap(&t2, w2, 2, dims, strides, box, 3)) return 1;
  }
  Ffn2Args a{};
  a.y = reinterpret_cast<const __nv_bfloat16*>(y); a.b1 = b1; a.b2 = b2; a.resid = resid; a.z2 = z2;
  a.hid = reinterpret_cast<__nv_bfloat16*>(hid); a.mask_bits = mask_bits; a.ld_bits = ld_bits; a.T = T; a.F = F;
  const int n_items = ((T + 127) / 128 + 1) / 2;
  const int max_clusters = num_sms() / 2;
  const int clusters = n_items < max_clusters ? n_items : max_clusters;
  ffn_fwd2_kernel<<<2 * clusters, 320, SMEM_BYTES, stream>>>(t1, t2, a);
  CB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cb
