// topk_merge.cu — K3: merge per-shard / per-database top-k lists into one top-k per query.
//
// Replaces the Python `sorted(all_results, ...)[:top_n]` over per-report hits
// (src/retrieval.py:293-295, 1719, 1754) and is the consumer of the multi-GPU all-gather: the
// input layout [n_lists][nq][k_in] is exactly what an all-gather of per-rank [nq, k_in] results
// produces.  Latency-bound (a few hundred candidates per query): one CTA per query, candidates are
// packed into ordered 64-bit keys in shared memory and bitonic-sorted.
//
// Tie rule: score descending, then lower list index, then lower slot.  For row shards given in
// rank order (list l holds lower global ids than list l+1, each list already id-ascending among
// equal scores) this is the global "ties by ascending id" order of the unsharded search.
#include <algorithm>

#include "b2r_common.cuh"

namespace b2r {

// score type -> key type / packing.  float: 64-bit key; double (BM25): 128-bit key.
template <class T>
struct MergeKey;
template <>
struct MergeKey<float> {
  typedef u64 K;
  static __device__ __forceinline__ K pack(float s, u32 pos) { return make_key(s, pos); }
  static __device__ __forceinline__ u32 pos(K k) { return key_id(k); }
  static __device__ __forceinline__ float pad() { return -3.4028235e38f; }
};
template <>
struct MergeKey<double> {
  typedef u128 K;
  static __device__ __forceinline__ K pack(double s, u32 pos) {
    s = s + 0.0;
    const u64 u = static_cast<u64>(__double_as_longlong(s));
    const u64 o = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    return (static_cast<u128>(o) << 64) | static_cast<u128>(0xFFFFFFFFu - pos);
  }
  static __device__ __forceinline__ u32 pos(K k) { return 0xFFFFFFFFu - static_cast<u32>(k); }
  static __device__ __forceinline__ double pad() { return -1.7976931348623157e308; }
};

template <class T>
__global__ void __launch_bounds__(256) topk_merge_kernel(const T* __restrict__ D_lists,
                                                         const long long* __restrict__ I_lists, int n_lists, int nq,
                                                         int k_in, int k_out, T* __restrict__ D,
                                                         long long* __restrict__ I, int buf_cap) {
  typedef typename MergeKey<T>::K K;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* buf = reinterpret_cast<K*>(smem_raw);
  const int qi = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int total = n_lists * k_in;
  for (int i = tid; i < buf_cap; i += nt) buf[i] = 0;
  __syncthreads();
  const int chunk = buf_cap - k_out;
  for (int base = 0; base < total; base += chunk) {
    for (int i = tid; i < chunk; i += nt) {
      const int g = base + i;
      K key = 0;
      if (g < total) {
        const int l = g / k_in, s = g - l * k_in;
        const size_t idx = (static_cast<size_t>(l) * nq + qi) * k_in + s;
        if (I_lists[idx] >= 0) key = MergeKey<T>::pack(D_lists[idx], static_cast<u32>(g));
      }
      buf[k_out + i] = key;
    }
    __syncthreads();
    block_sort_desc(buf, buf_cap);
  }
  for (int i = tid; i < k_out; i += nt) {
    const K key = buf[i];
    const size_t o = static_cast<size_t>(qi) * k_out + i;
    if (key == 0) {
      D[o] = MergeKey<T>::pad();
      I[o] = -1;
    } else {
      const int g = static_cast<int>(MergeKey<T>::pos(key));
      const int l = g / k_in, s = g - l * k_in;
      const size_t idx = (static_cast<size_t>(l) * nq + qi) * k_in + s;
      D[o] = D_lists[idx];
      I[o] = I_lists[idx];
    }
  }
}

static int pow2_ge(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <class T>
static int merge_launch(const T* D_lists, const int64_t* I_lists, int n_lists, int nq, int k_in, int k_out,
                        T* D, int64_t* I, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_lists) * k_in;
  int buf_cap = pow2_ge(static_cast<int>(std::min<long long>(total, 4096)) + k_out);
  buf_cap = std::max(buf_cap, 2 * pow2_ge(k_out));
  buf_cap = std::min(buf_cap, 8192);
  const size_t smem = static_cast<size_t>(buf_cap) * sizeof(typename MergeKey<T>::K);
  B2R_CUDA(cudaFuncSetAttribute(topk_merge_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  topk_merge_kernel<T><<<nq, 256, smem, stream>>>(D_lists, reinterpret_cast<const long long*>(I_lists), n_lists, nq,
                                               k_in, k_out, D, reinterpret_cast<long long*>(I), buf_cap);
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

static int merge_check(const void* D_lists, const int64_t* I_lists, int n_lists, int nq, int k_in, int k_out,
                       const void* D, const int64_t* I) {
  if (n_lists < 1 || nq < 0 || k_in < 1 || k_out < 1) return fail(B2R_ERR_INVALID, "b2r_topk_merge: bad sizes");
  if (k_out > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_topk_merge: k_out > B2R_MAX_K (1024)");
  if (static_cast<long long>(n_lists) * k_in > (1LL << 30)) return fail(B2R_ERR_UNSUPPORTED, "b2r_topk_merge: too many candidates");
  if (nq > 0 && (!D_lists || !I_lists || !D || !I)) return fail(B2R_ERR_INVALID, "b2r_topk_merge: NULL buffer");
  return B2R_OK;
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_topk_merge_dev(const float* D_lists, const int64_t* I_lists, int n_lists, int nq, int k_in, int k_out,
                       float* D, int64_t* I, void* stream) {
  int rc = merge_check(D_lists, I_lists, n_lists, nq, k_in, k_out, D, I);
  if (rc) return rc;
  if (nq == 0) return B2R_OK;
  return merge_launch(D_lists, I_lists, n_lists, nq, k_in, k_out, D, I, static_cast<cudaStream_t>(stream));
}

int b2r_topk_merge_f64_dev(const double* S_lists, const int64_t* I_lists, int n_lists, int nq, int k_in, int k_out,
                           double* S, int64_t* I, void* stream) {
  int rc = merge_check(S_lists, I_lists, n_lists, nq, k_in, k_out, S, I);
  if (rc) return rc;
  if (nq == 0) return B2R_OK;
  return merge_launch(S_lists, I_lists, n_lists, nq, k_in, k_out, S, I, static_cast<cudaStream_t>(stream));
}

int b2r_topk_merge_host(const float* D_lists, const int64_t* I_lists, int n_lists, int nq, int k_in, int k_out,
                        float* D, int64_t* I, int device) {
  int rc = merge_check(D_lists, I_lists, n_lists, nq, k_in, k_out, D, I);
  if (rc) return rc;
  if (nq == 0) return B2R_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_topk_merge_host: cudaSetDevice failed");
  const size_t n_in = static_cast<size_t>(n_lists) * nq * k_in, n_out = static_cast<size_t>(nq) * k_out;
  float *dD = nullptr, *dDo = nullptr;
  int64_t *dI = nullptr, *dIo = nullptr;
  cudaError_t e = cudaMalloc(&dD, n_in * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&dI, n_in * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc(&dDo, n_out * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&dIo, n_out * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMemcpy(dD, D_lists, n_in * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dI, I_lists, n_in * sizeof(int64_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = merge_launch(dD, dI, n_lists, nq, k_in, k_out, dDo, dIo, nullptr);
    if (rc == B2R_OK) {
      e = cudaMemcpy(D, dDo, n_out * sizeof(float), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(I, dIo, n_out * sizeof(int64_t), cudaMemcpyDeviceToHost);
    }
  }
  cudaFree(dD);
  cudaFree(dI);
  cudaFree(dDo);
  cudaFree(dIo);
  if (rc == B2R_OK && e != cudaSuccess) rc = fail(B2R_ERR_CUDA, std::string("b2r_topk_merge_host: ") + cudaGetErrorString(e));
  return rc;
}

}  // extern "C"
