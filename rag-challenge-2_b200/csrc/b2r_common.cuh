// b2r_common.cuh — shared host/device helpers for libb2r (sm_100a only).
//
// Contents: error plumbing for the C ABI, mbarrier / bulk-copy (TMA) PTX wrappers, the 64-bit
// ordered (score, id) key used by every select in the library, warp- and block-level bitonic
// sorts over shared memory, and block_select_lists(): the "many sorted lists -> one top-k"
// routine used by the last-CTA epilogue of the scan kernels and by the stand-alone merge kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "b2r.h"

namespace b2r {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned __int128 u128;  // (ordered float64 score << 64) | inverted id — BM25 keys

// ----------------------------------------------------------------------------------------------
// host: error plumbing
// ----------------------------------------------------------------------------------------------
std::string& last_error_ref();
int fail(int code, const std::string& msg);
int& launch_counter();  // thread-local: kernels enqueued by the last search on this thread

#define B2R_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      return ::b2r::fail(_e == cudaErrorMemoryAllocation ? B2R_ERR_NOMEM : B2R_ERR_CUDA,      \
                         std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
    }                                                                                         \
  } while (0)

int tunable(const char* key);  // current value of a tunable (see b2r_tunable_set)

// Kernel workspaces are shared by all index handles: one buffer per (device, stream, kind), grown on demand.  Work on
// one stream is ordered, so every handle searching on that stream can use the same scratch; HBM spent on scratch
// then scales with the number of streams in use (the callers' concurrency), not with the number of indexes — a
// corpus of thousands of per-report indexes costs no more than one.  kind: 0 = flat scan / tensor-core scratch,
// 1 = BM25 (separate because BM25 runs BESIDE the scan on the same stream).  The first `zero_prefix` bytes are
// zeroed when the buffer is (re)allocated; the kernels leave them zero.  The caller has the device selected.
int ws_pool_get(int device, cudaStream_t stream, int kind, size_t bytes, size_t zero_prefix, unsigned char** base,
                size_t* cap);
// A monotonically increasing 32-bit sequence per workspace (wraps): returns the value before adding `count`.  The scan
// kernels' dynamic tile counter is never reset — a launch subtracts the sequence value it was given — so the caller must
// take the value and enqueue its kernel while holding ws_launch_mutex() (launch order == sequence order).
uint32_t ws_pool_advance(int device, cudaStream_t stream, int kind, uint32_t count);
std::mutex& ws_launch_mutex();
}  // namespace b2r
// host-side range check of query term ids against an index (csrc/bm25.cu)
extern "C" int bm25_check_terms(const b2r_bm25* h, const int32_t* term_ids, int n);
namespace b2r {

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// ----------------------------------------------------------------------------------------------
// device: PTX wrappers (mbarrier + 1-D bulk async copy, the TMA engine without a tensor map —
// row tiles of a row-major matrix are contiguous, so a tile is one linear copy)
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ u32 smem_u32(const void* p) {
  return static_cast<u32>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
  // make the initialised barriers visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  u32 done;
  const u32 addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// L2 eviction policy for data that is streamed once per query.
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// global -> shared linear bulk copy, completion counted in bytes on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar,
                                         u64 policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ float4 lds128(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}

// ----------------------------------------------------------------------------------------------
// device: ordered (score, id) key
//   high 32 bits: the float score mapped so that unsigned order == float order
//   low  32 bits: 0xFFFFFFFF - local id, so that among equal scores the LOWER id is the larger key
//   key 0 is reserved for "empty" (it would decode to id 0xFFFFFFFF, which is never a valid row).
// Sorting keys descending therefore yields score descending, id ascending: the order faiss's heap
// reorder and Python's stable sort (src/retrieval.py:265) both produce.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 float_to_ordered(float s) {
  s = s + 0.0f;                    // -0.0 -> +0.0 so that equal scores have equal keys
  if (!(s == s)) s = -3.4028235e38f;  // NaN never outranks a number (faiss: comparisons are false)
  const u32 u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float ordered_to_float(u32 o) {
  const u32 u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}

__device__ __forceinline__ u64 make_key(float s, u32 local_id) {
  return (static_cast<u64>(float_to_ordered(s)) << 32) | static_cast<u64>(0xFFFFFFFFu - local_id);
}

__device__ __forceinline__ float key_score(u64 key) { return ordered_to_float(static_cast<u32>(key >> 32)); }
__device__ __forceinline__ u32 key_id(u64 key) { return 0xFFFFFFFFu - static_cast<u32>(key & 0xFFFFFFFFull); }

// ----------------------------------------------------------------------------------------------
// device: bitonic sorts (descending) over u64 keys in shared memory; n must be a power of two
// ----------------------------------------------------------------------------------------------
template <class K>
__device__ __forceinline__ void cmpxchg_desc(K* a, int i, int j, bool up) {
  // `up` == true: this pair belongs to an ascending run of the bitonic network
  const K x = a[i], y = a[j];
  const bool swap = up ? (x > y) : (x < y);
  if (swap) {
    a[i] = y;
    a[j] = x;
  }
}

// one warp sorts a[0..n) descending; all 32 lanes must call it.  n is a power of two, so all the
// index arithmetic of the bitonic network is shifts and masks (size = 1 << ls, stride = 1 << lj).
template <class K>
__device__ __forceinline__ void warp_sort_desc(K* a, int n, int lane) {
  for (int ls = 1; (1 << ls) <= n; ++ls) {
    for (int lj = ls - 1; lj >= 0; --lj) {
      const int stride = 1 << lj;
      for (int t = lane; t < (n >> 1); t += 32) {
        const int i = ((t >> lj) << (lj + 1)) | (t & (stride - 1));
        const bool up = ((i >> ls) & 1) != 0;
        cmpxchg_desc(a, i, i | stride, up);
      }
      __syncwarp();
    }
  }
}

// the whole CTA sorts a[0..n) descending; every thread of the block must call it
template <class K>
__device__ __forceinline__ void block_sort_desc(K* a, int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int ls = 1; (1 << ls) <= n; ++ls) {
    for (int lj = ls - 1; lj >= 0; --lj) {
      const int stride = 1 << lj;
      for (int t = tid; t < (n >> 1); t += nt) {
        const int i = ((t >> lj) << (lj + 1)) | (t & (stride - 1));
        const bool up = ((i >> ls) & 1) != 0;
        cmpxchg_desc(a, i, i | stride, up);
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ----------------------------------------------------------------------------------------------
// device: the kp largest of n UNSORTED keys in shared memory, by histogram (O(n) instead of a full
// bitonic sort).  Whole-CTA routine; leaves the result sorted descending in out[0..) (zero padded up
// to at least kp entries).
//   keys : shared, n keys (u64 ordered keys or u32 ordered scores); clobbered only by the fallback
//   hist : shared u32[kSelBins + 4]
//   out  : shared, out_cap keys (power of two, >= 2 kp)
// Keys are binned linearly over [min, max] of their score bits; the bins are summed from the top until
// kp keys are covered, the keys of those bins (a little over kp) are compacted into `out` and sorted.
// If they do not fit (heavy ties) the whole array is sorted instead.
// ----------------------------------------------------------------------------------------------
static const int kSelBins = 2048;

__device__ __forceinline__ u32 key_score_bits(u128 k) { return static_cast<u32>(k >> 96); }
__device__ __forceinline__ u32 key_score_bits(u64 k) { return static_cast<u32>(k >> 32); }
__device__ __forceinline__ u32 key_score_bits(u32 k) { return k; }

template <class K>
__device__ __forceinline__ void block_topk_hist(K* keys, int n, int n_pow2_cap, int kp, u32* hist, K* out, int out_cap) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (n <= out_cap) {  // small input: sort it directly
    int p = 2;
    while (p < n || p < kp) p <<= 1;
    for (int i = tid; i < p; i += nt) out[i] = i < n ? keys[i] : static_cast<K>(0);
    __syncthreads();
    block_sort_desc(out, p);
    return;
  }
  u32* s_lo = hist + kSelBins;      // min score bits
  u32* s_hi = hist + kSelBins + 1;  // max score bits
  u32* s_bin = hist + kSelBins + 2; // threshold bin
  u32* s_cnt = hist + kSelBins + 3; // survivors
  for (int i = tid; i < kSelBins; i += nt) hist[i] = 0;
  if (tid == 0) {
    *s_lo = 0xFFFFFFFFu;
    *s_hi = 0u;
    *s_cnt = 0u;
  }
  __syncthreads();
  {
    u32 lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = tid; i < n; i += nt) {
      const K v = keys[i];
      if (v == 0) continue;  // empty slot: never selected, and it must not stretch the bin range
      const u32 s = key_score_bits(v);
      lo = s < lo ? s : lo;
      hi = s > hi ? s : hi;
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) {
      atomicMin(s_lo, lo);
      atomicMax(s_hi, hi);
    }
  }
  __syncthreads();
  const u32 lo = *s_lo;
  const u32 range = *s_hi >= lo ? *s_hi - lo : 0u;  // no valid key at all: lo = 0xFFFFFFFF, hi = 0
  int sh = 0;
  while ((range >> sh) >= static_cast<u32>(kSelBins)) ++sh;
  for (int i = tid; i < n; i += nt) {
    const K v = keys[i];
    if (v != 0) atomicAdd(&hist[(key_score_bits(v) - lo) >> sh], 1u);
  }
  __syncthreads();
  if (tid < 32) {  // one warp walks the bins from the top: lane l owns bins [2048 - 64 (l + 1), 2048 - 64 l)
    constexpr int per = kSelBins / 32;
    const int top = kSelBins - per * tid;
    u32 sum = 0;
    for (int b = top - 1; b >= top - per; --b) sum += hist[b];
    u32 incl = sum;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const u32 v = __shfl_up_sync(0xffffffffu, incl, m);
      if (tid >= m) incl += v;
    }
    const u32 before = incl - sum;  // keys in the bins above this lane's
    if (before < static_cast<u32>(kp) && incl >= static_cast<u32>(kp)) {
      u32 acc = before;
      int b = top - 1;
      for (; b > top - per; --b) {
        acc += hist[b];
        if (acc >= static_cast<u32>(kp)) break;
      }
      *s_bin = static_cast<u32>(b);
    }
    if (tid == 31 && incl < static_cast<u32>(kp)) *s_bin = 0u;  // fewer than kp keys: keep everything
  }
  __syncthreads();
  const u32 bstar = *s_bin;
  for (int i = tid; i < n; i += nt) {
    const K v = keys[i];
    if (v != 0 && ((key_score_bits(v) - lo) >> sh) >= bstar) {
      const u32 pos = atomicAdd(s_cnt, 1u);
      if (pos < static_cast<u32>(out_cap)) out[pos] = v;
    }
  }
  __syncthreads();
  const u32 cnt = *s_cnt;
  if (cnt <= static_cast<u32>(out_cap)) {
    int p = 2;
    while (p < static_cast<int>(cnt) || p < kp) p <<= 1;
    for (int i = cnt + tid; i < p; i += nt) out[i] = 0;
    __syncthreads();
    block_sort_desc(out, p);
    return;
  }
  // too many keys share the threshold bin (ties): sort everything
  int p = 2;
  while (p < n) p <<= 1;
  if (p > n_pow2_cap) p = n_pow2_cap;
  for (int i = n + tid; i < p; i += nt) keys[i] = 0;
  __syncthreads();
  block_sort_desc(keys, p);
  for (int i = tid; i < out_cap; i += nt) out[i] = i < p ? keys[i] : static_cast<K>(0);
  __syncthreads();
}

// ----------------------------------------------------------------------------------------------
// device: select the k largest keys out of n_lists lists of `stride` keys each (every list sorted
// descending, empty slots == 0) and leave them sorted in buf[0..k).  Whole-CTA routine.
//   src      : global memory, n_lists * stride keys
//   buf      : shared scratch of buf_cap keys (power of two, >= 2k)
//   s_ctl    : shared u64[2] scratch
// Fast path: every key that can be in the answer is >= T = max over lists of the list's k-th key
// (that list alone already holds k keys >= T), so only the survivors of that filter are sorted —
// for randomly distributed data a little over k of them.  If the survivors do not fit (adversarial
// layouts) fall back to streaming rounds of "keep best k, append, sort".
// ----------------------------------------------------------------------------------------------
// the part of a key the prefilter compares (the score bits; for 128-bit keys ties on the score
// survive the filter and are resolved by the sort)
__device__ __forceinline__ u64 key_prune_part(u64 k) { return k; }
__device__ __forceinline__ u64 key_prune_part(u128 k) { return static_cast<u64>(k >> 64); }

template <class K>
__device__ __forceinline__ void block_select_lists(const K* __restrict__ src, int n_lists, int stride,
                                                   int k, K* buf, int buf_cap, u64* s_ctl, u32* hist = nullptr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long total = static_cast<long long>(n_lists) * stride;
  u32* s_cnt = reinterpret_cast<u32*>(s_ctl);  // [0] survivor count
  u64* s_thr = s_ctl + 1;                      // threshold key
  if (tid == 0) {
    *s_cnt = 0;
    *s_thr = 0;
  }
  __syncthreads();
  if (k <= stride) {
    u64 t = 0;
    for (int l = tid; l < n_lists; l += nt) {
      const u64 v = key_prune_part(src[static_cast<long long>(l) * stride + (k - 1)]);
      t = v > t ? v : t;
    }
    if (t) atomicMax(s_thr, t);
  }
  __syncthreads();
  const u64 thr = *s_thr;
  // Second, usually much tighter bound when there are many lists: the k-th largest of the lists'
  // best entries.  Those k lists each hold at least one key >= theta, so >= k keys survive, and
  // for data spread over many lists only a little over k do.  Compared as full keys, so runs of
  // equal scores (BM25's zero-score tail) do not all survive.
  K theta = 0;
  const int lp = next_pow2(n_lists < 2 ? 2 : n_lists);
  if (n_lists >= k && lp <= buf_cap) {
    for (int l = tid; l < lp; l += nt) buf[l] = l < n_lists ? src[static_cast<long long>(l) * stride] : static_cast<K>(0);
    __syncthreads();
    const int oc = next_pow2(2 * k);
    if (hist != nullptr && lp > oc && lp + oc <= buf_cap) {  // histogram select of the k-th largest list maximum
      block_topk_hist(buf, lp, lp, k, hist, buf + lp, oc);
      theta = buf[lp + k - 1];
    } else {
      block_sort_desc(buf, lp);
      theta = buf[k - 1];
    }
    __syncthreads();
  }
  for (long long i = tid; i < total; i += nt) {
    const K v = src[i];
    if (v != 0 && key_prune_part(v) >= thr && v >= theta) {
      const u32 pos = atomicAdd(s_cnt, 1u);
      if (pos < static_cast<u32>(buf_cap)) buf[pos] = v;
    }
  }
  __syncthreads();
  const u32 cnt = *s_cnt;
  if (cnt <= static_cast<u32>(buf_cap)) {
    const int need = static_cast<int>(cnt) > k ? static_cast<int>(cnt) : k;
    const int p = next_pow2(need < 2 ? 2 : need);
    for (int i = cnt + tid; i < p; i += nt) buf[i] = 0;
    __syncthreads();
    block_sort_desc(buf, p);
    return;
  }
  // fallback: streaming rounds; buf[0..k) holds the best so far
  for (int i = tid; i < buf_cap; i += nt) buf[i] = 0;
  __syncthreads();
  const int chunk = buf_cap - k;
  for (long long base = 0; base < total; base += chunk) {
    for (int i = tid; i < chunk; i += nt) {
      const long long g = base + i;
      buf[k + i] = g < total ? src[g] : static_cast<K>(0);
    }
    __syncthreads();
    block_sort_desc(buf, buf_cap);
  }
}

// write buf[0..k) (sorted keys) as (score, int64 id) with faiss padding for empty slots
__device__ __forceinline__ void block_emit_keys(const u64* buf, int k, float* D, long long* I,
                                                long long id_offset) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const u64 key = buf[i];
    if (key == 0) {
      D[i] = -3.4028235e38f;
      I[i] = -1;
    } else {
      D[i] = key_score(key);
      I[i] = static_cast<long long>(key_id(key)) + id_offset;
    }
  }
}

#endif  // __CUDACC__

}  // namespace b2r
