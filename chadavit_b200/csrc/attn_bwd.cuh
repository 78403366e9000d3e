// Shared between the two generations of the varlen attention backward (attn_bwd.cu: generation 1, used for head_dim 128;
// attn_bwd2.cu: generation 2, head_dim <= 96).
#pragma once
#include "common.cuh"

namespace cb {

template <int HD>
struct BwdCfg {
  static constexpr int CHUNK = (HD % 64 == 0) ? 64 : (HD % 32 == 0 ? 32 : 16);
  static constexpr int NCH = HD / CHUNK;
  static constexpr int SWZ = CHUNK == 64 ? 3 : (CHUNK == 32 ? 2 : 1);
  static constexpr int CHUNK_BYTES = 128 * CHUNK * 2;
  static constexpr int TILE_BYTES = NCH * CHUNK_BYTES;
  static constexpr int SBO = 8 * CHUNK * 2;
  static constexpr int QDO_STAGES = HD <= 96 ? 2 : 1;
  static constexpr int DS_BYTES = 128 * 128 * 2;  // two [128 x 64] 128B-swizzled sub-tiles
  static constexpr int DQ_SLABS = (HD + 31) / 32;       // 16-column dQ chunks handled by one epilogue warp
  static constexpr int DQ_STAGE_BYTES = 8 * DQ_SLABS * 2048;   // dQ drain: per warp DQ_SLABS x (32 rows x 64 B) slabs for TMA reduce-add
  static constexpr int SMEM_BYTES = TILE_BYTES * (2 + 2 * QDO_STAGES) + DS_BYTES + DQ_STAGE_BYTES + 1024 /*lse,delta*/ + 1024 + 256;
  static constexpr int COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 384;
};

__device__ __forceinline__ float fast_exp2_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnBwdArgs {
  const int4* work;  // {kv_row0 (global row), seq_start, seq_end, head}
  int n_work;
  const float* lse;    // [H, T]
  const float* delta;  // [H, T]
  float* dq_acc;       // [T, D] fp32, zero-initialised
  __nv_bfloat16* dqkv; // [T, 3D]: dK -> cols [D,2D), dV -> cols [2D,3D)
  int T, D;
  float scale, scale_log2;
};

// generation 2 (attn_bwd2.cu); returns non-zero on launch failure.  HD in {16, 32, 64, 96}.
int attn_bwd2_launch(int hd, const void* qkv, const void* dO, const AttnBwdArgs& a, cudaStream_t stream);

}  // namespace cb
