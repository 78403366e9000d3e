// Host-side plumbing shared by every entry point: error strings, SM count, CUtensorMap encoding through the
// driver entry point (so the library loads on machines without libcuda, e.g. the CPU build/test container).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <queue>
#include <vector>

#include "common.cuh"
#include "chadavit_b200.h"

namespace cb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return 1000 + (int)e;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle, int elem_bytes) {
  EncodeTiledFn enc = get_encode();
  CB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu] box=[%u,%u,%u] stride0=%llu swz=%d base=%p",
              (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
              (unsigned long long)(rank > 2 ? d[2] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0,
              (unsigned long long)(rank > 1 ? s[0] : 0), swizzle, base);
    return 2;
  }
  return 0;
}

}  // namespace cb

extern "C" const char* cb_last_error(void) { return cb::g_err; }
extern "C" int cb_version(void) { return CHADAVIT_B200_VERSION; }
extern "C" int cb_num_sms(void) { return cb::num_sms(); }
extern "C" int cb_sync_check(void* stream) {
  CB_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  CB_CUDA(cudaGetLastError());
  return 0;
}

// Host-side work schedule of the persistent varlen attention kernels (no device work, no CUDA call).  Items
// {q_row0, seq_start, seq_end, head} in 128*k-row query/kv tiles, longest sequences first; for mode 1 / 2 they are then
// assigned longest-processing-time-first to the least loaded of n_ctas CTAs and written round-major (CTA c owns the slots
// c, c + n_ctas, ...), padded with all-zero slots the kernels skip.  Everything derives from the Python ints of
// list_num_channels (channels_strategies.py:31-85 contract), so a new ragged batch costs no device synchronisation.
extern "C" int cb_attn_schedule(const int* cu_host, int B, int nheads, int tile, int mode, int n_ctas, int* out, int cap, int* n_out) {
  using namespace cb;
  CB_CHECK(cu_host && out && n_out && B > 0 && nheads > 0 && tile > 0 && mode >= 0 && mode <= 2 && (mode == 0 || n_ctas > 0),
           "attn_schedule: bad arguments B=%d heads=%d tile=%d mode=%d n_ctas=%d", B, nheads, tile, mode, n_ctas);
  struct Item { int q0, s, e, h; long cost; };
  std::vector<int> order(B);
  for (int b = 0; b < B; ++b) order[b] = b;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cu_host[x + 1] - cu_host[x] > cu_host[y + 1] - cu_host[y]; });
  std::vector<Item> items;
  for (int b : order) {
    const int s = cu_host[b], e = cu_host[b + 1];
    for (int h = 0; h < nheads; ++h)
      for (int q0 = s; q0 < e; q0 += tile) {
        const long seq = e - s;
        long cost;
        if (mode == 1) cost = ((seq + 63) / 64) * std::min<long>((e - q0 + 127) / 128, tile / 128) + 3;   // kv sub-tiles x query tiles
        else cost = (seq + 127) / 128 + 1;                                                                 // query tiles per kv tile
        items.push_back({q0, s, e, h, cost});
      }
  }
  const int n = (int)items.size();
  auto put = [&](int slot, const Item& it) { out[4 * slot] = it.q0; out[4 * slot + 1] = it.s; out[4 * slot + 2] = it.e; out[4 * slot + 3] = it.h; };
  if (mode == 0 || n <= n_ctas) {
    *n_out = n;
    if (n > cap) { set_error("attn_schedule: need %d slots, capacity %d", n, cap); return 3; }
    for (int i = 0; i < n; ++i) put(i, items[i]);
    return 0;
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return items[x].cost > items[y].cost; });
  typedef std::pair<long, int> LC;   // (load, cta): least loaded first, ties to the lower CTA index
  std::priority_queue<LC, std::vector<LC>, std::greater<LC>> heap;
  for (int c = 0; c < n_ctas; ++c) heap.push({0, c});
  std::vector<int> cta_of(n), round_of(n), fill(n_ctas, 0);
  int rounds = 0;
  for (int i : idx) {
    LC t = heap.top(); heap.pop();
    cta_of[i] = t.second; round_of[i] = fill[t.second]++;
    rounds = std::max(rounds, fill[t.second]);
    heap.push({t.first + items[i].cost, t.second});
  }
  *n_out = rounds * n_ctas;
  if (*n_out > cap) { set_error("attn_schedule: need %d slots, capacity %d", *n_out, cap); return 3; }
  std::fill(out, out + 4 * (size_t)*n_out, 0);
  for (int i = 0; i < n; ++i) put(round_of[i] * n_ctas + cta_of[i], items[i]);
  return 0;
}
