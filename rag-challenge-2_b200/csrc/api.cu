// api.cu — library-level plumbing of the C ABI: thread-local error string, tunables, version.
#include <map>
#include <tuple>
#include <algorithm>
#include <mutex>
#include <string>

#include "b2r_common.cuh"

namespace b2r {

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int fail(int code, const std::string& msg) {
  last_error_ref() = msg;
  return code;
}

int& launch_counter() {
  static thread_local int n = 0;
  return n;
}

static std::mutex g_tun_mu;
static std::map<std::string, int>& tunables() {
  // 0 / negative == "let the library decide"
  static std::map<std::string, int> t = {
      {"flat.stages", 0},          // depth of the bulk-copy ring
      {"flat.tile_rows", 0},       // rows per tile (rounded to a multiple of the consumer warps)
      {"flat.ctas_per_sm", 0},     // resident CTAs per SM the plan aims for
      {"flat.consumer_warps", 0},  // 4 or 8
      {"flat.max_nq", 0},          // cap on queries sharing one pass (1, 2 or 4)
      {"flat.force_generic", 0},   // 1: use the non-TMA kernel even for templated widths
      {"flat.dynamic_tiles", 0},   // scan tile order: 0 default (dynamic, chunks of 2 tiles), -1 static (t mod grid), n chunk size
      {"flat.no_combine", 0},      // 1: concurrent host callers of one index are not combined into one batch
      {"flat.no_gemm", 0},         // 1: never use the tensor-core batch kernel
      {"flat.gemm_min_nq", 0},     // smallest batch that goes to the tensor-core kernel (default 5)
      {"flat.gemm_max_n", 0},      // queries per tensor-core pass (default 256)
      {"gemm.stages", 0},          // depth of the tensor-core kernel's TMA ring
      {"cosine.no_pair", 0},       // 1: N x N statistics on the single-CTA 3xTF32 kernel instead of the CTA-pair one
      {"gemm.no_bf16", 0},         // 1: ignore the bf16 copy of the rows (candidate pass stays tf32)
      {"gemm.no_pair", 0},         // 1: at most 128 queries per pass (no CTA-pair kernel)
      {"gemm.small_rescore", 0},   // 1: indexes under 16,384 rows also take the candidate + fp32 rescoring path
      {"gemm.no_wide", 0},         // 1: candidate pass keeps lists in shared memory (<= 64 queries / pass) instead of thresholds
      {"gemm.wide_sub", 0},        // 128-row sub-tiles per tile of the threshold candidate pass (1 or 2; default by width)
      {"gemm.tile256", 0},         // 1: candidate pass scores 256 rows per query slab (experimental)
      {"gemm.debug_skip", 0},      // perf experiments: 1 skip query-slab reloads, 2 skip MMAs, 4 skip select (wrong results)
      {"gemm.mode", 0},            // 0: tf32 candidates + exact fp32 rescoring (verified); 3: 3xTF32 only
      {"gemm.no_verify", 0},       // 1: skip the 3xTF32 redo pass of unproven queries
      {"select.buf_cap", 0},       // cap on the final-select scratch (keys); small values force the fallback
      {"bm25.block_docs", 0},      // documents per CTA in the doc-block BM25 kernel (bm25.v1)
      {"bm25.debug_skip", 0},      // perf experiments (wrong results): 1 no select, 2 no accumulate, 4 no merge
      {"bm25.no_skip", 0},         // 1: streaming kernel ignores the block index (binary searches in the producer)
      {"exchange.timeout_ms", 0},  // bound on the exchange kernel's wait for its peers (default 10000 ms)
      {"bm25.no_slots", 0},        // 1: streaming kernel without the per-query threshold slots (r02 A/B)
      {"bm25.pilot_blocks", 0},    // blocks per query in the pilot launch (0: about three waves in all, -1: no pilot launch)
      {"bm25.cfg", 0},             // ring geometry of the streaming kernel (A/B runs; see bm25_stream_launch)
      {"bm25.ctas_per_sm", 0},     // cap on resident streaming-kernel CTAs per SM
      {"hybrid.no_copy_stream", 0},// 1: the host step's H2D copy goes on the step's own stream (experiment)
      {"hybrid.host_chain", 0},    // 1: the host step (submit / collect) also chains consecutive steps by dependent launch (experiment)
      {"hybrid.no_fused_tail", 0}, // 1: the hybrid step keeps its fusion / exchange kernel (no fused tail in the BM25 kernel)
      {"bm25.v1", 0},              // 1: use the doc-block kernel (binary searches, per-token barriers) instead of the streaming one
  };
  return t;
}

struct PoolWs {
  unsigned char* base = nullptr;
  size_t bytes = 0;
  uint32_t seq = 0;
};
static std::mutex g_ws_mu;
static std::map<std::tuple<int, cudaStream_t, int>, PoolWs>& ws_pool() {
  static std::map<std::tuple<int, cudaStream_t, int>, PoolWs> m;
  return m;
}

int ws_pool_get(int device, cudaStream_t stream, int kind, size_t bytes, size_t zero_prefix, unsigned char** base,
                size_t* cap) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  PoolWs& ws = ws_pool()[std::make_tuple(device, stream, kind)];
  if (ws.bytes < bytes) {
    if (ws.base) {  // growing: earlier work on this stream may still be using the old buffer
      B2R_CUDA(cudaStreamSynchronize(stream));
      cudaFree(ws.base);
      ws.base = nullptr;
      ws.bytes = 0;
    }
    B2R_CUDA(cudaMalloc(&ws.base, bytes));
    ws.bytes = bytes;
    ws.seq = 0;  // the new buffer's counters start from zero
    if (zero_prefix) B2R_CUDA(cudaMemsetAsync(ws.base, 0, std::min(zero_prefix, bytes), stream));
  }
  *base = ws.base;
  if (cap) *cap = ws.bytes;
  return B2R_OK;
}

uint32_t ws_pool_advance(int device, cudaStream_t stream, int kind, uint32_t count) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  PoolWs& ws = ws_pool()[std::make_tuple(device, stream, kind)];
  const uint32_t before = ws.seq;
  ws.seq += count;
  return before;
}

std::mutex& ws_launch_mutex() {
  static std::mutex m;
  return m;
}

int tunable(const char* key) {
  std::lock_guard<std::mutex> lock(g_tun_mu);
  auto& t = tunables();
  auto it = t.find(key);
  return it == t.end() ? 0 : it->second;
}

}  // namespace b2r

using namespace b2r;

extern "C" {

const char* b2r_last_error(void) { return last_error_ref().c_str(); }

int b2r_version(void) { return 100; /* 0.1.0 */ }

int b2r_device_count(int* out) {
  if (!out) return fail(B2R_ERR_INVALID, "b2r_device_count: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    return fail(B2R_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *out = n;
  return B2R_OK;
}

int b2r_tunable_set(const char* key, int value) {
  if (!key) return fail(B2R_ERR_INVALID, "b2r_tunable_set: key is NULL");
  std::lock_guard<std::mutex> lock(g_tun_mu);
  auto& t = tunables();
  auto it = t.find(key);
  if (it == t.end()) return fail(B2R_ERR_INVALID, std::string("unknown tunable: ") + key);
  it->second = value;
  return B2R_OK;
}

int b2r_tunable_get(const char* key, int* value) {
  if (!key || !value) return fail(B2R_ERR_INVALID, "b2r_tunable_get: NULL argument");
  std::lock_guard<std::mutex> lock(g_tun_mu);
  auto& t = tunables();
  auto it = t.find(key);
  if (it == t.end()) return fail(B2R_ERR_INVALID, std::string("unknown tunable: ") + key);
  *value = it->second;
  return B2R_OK;
}

int b2r_last_launch_count(void) { return launch_counter(); }

}  // extern "C"
