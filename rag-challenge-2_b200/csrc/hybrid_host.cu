// hybrid_host.cu — one hybrid query step behind ONE C call with host buffers.
//
// The call a user of the drop-in makes per question is "embedding + tokens in, fused top-k out".  Driving the
// three device entry points from Python costs ~90 us of interpreter / ctypes / torch time per step, which is
// 10 % of a step on a 1.25M-row shard; this entry point does the whole step from C: one pinned host->device
// copy of the packed inputs, dense scan, BM25 launched beside it (programmatic dependent launch), fusion (or
// the one-kernel NVLink exchange + merge + fusion when peer buffers are given) storing its 12 * nq * k result
// bytes straight into mapped pinned host memory, one stream synchronisation.  It replaces, for the single-index case, the Python loop over
// index.search / bm25.get_scores / sorted(...)[:top_n] of src/retrieval.py:262-295,1140-1157.
#include <string.h>

#include <algorithm>
#include <mutex>

#include "hybrid_fuse.cuh"

// Two slots of staging per context: a caller may SUBMIT the next question while the previous one is still on the GPU
// (b2r_hybrid_submit_host / b2r_hybrid_collect_host) — both are enqueued on the context's one stream, so the GPU goes
// from one step to the next without waiting for the host.  b2r_hybrid_search_host is submit + collect.
static const int kHybridSlots = 2;

struct HybridSlot {
  unsigned char* in_pin = nullptr;
  unsigned char* in_dev = nullptr;
  size_t in_cap = 0;
  unsigned char* out_pin = nullptr;      // mapped pinned host memory: the fusion stores F / O straight into it
  unsigned char* out_pin_dev = nullptr;  // device alias of out_pin
  unsigned char* out_dev = nullptr;      // intermediate lists [I int64][S double][J int64][D float]
  size_t out_cap = 0;                    // in (nq * k) units
  cudaEvent_t done = nullptr;
  cudaEvent_t in_ready = nullptr;        // the slot's inputs have arrived in in_dev (recorded on the copy stream)
  bool busy = false;
  size_t nk = 0, h2d = 0;
  int world = 1;
  uint32_t step = 0;
};

struct b2r_hybrid_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;    // H2D copies of the NEXT question overlap the step that is running
  HybridSlot slot[kHybridSlots];
  int next = 0;
  std::mutex mu;                     // submissions / collections of one context are serialised
  size_t last_h2d = 0, last_d2h = 0;
};

using namespace b2r;

// One hybrid step, everything resident, in TWO launches when it can: the dense scan, and BM25 launched beside it
// whose last CTA per query — after griddepcontrol.wait, i.e. once the scan has completed — fuses (or exchanges,
// merges and fuses) in place.  Falls back to scan, BM25, fusion / exchange kernel when the fused tail does not apply
// (more than 4 queries, k > 256, a batch that goes to the tensor cores, the streaming BM25 kernel forced).
// flags: B2R_LAUNCH_OVERLAP_PREVIOUS chains this step behind the previous one (see b2r_flat_search_dev_ex).
static int hybrid_step_dev(b2r_flat* flat, b2r_bm25* bm25, const float* q_dev, int nq, const int32_t* tids_dev,
                           const int32_t* offs_dev, int k, int64_t id_offset, void* const* peer_bufs, int rank, int world,
                           int nq_max, int k_max, uint32_t step, uint32_t* status, float* D_dev, int64_t* I_dev, double* S_dev,
                           int64_t* J_dev, float* F_dev, int64_t* O_dev, cudaStream_t stream, unsigned flags, int* launches) {
  int rc = b2r_flat_search_dev_ex(flat, q_dev, nq, k, D_dev, I_dev, id_offset, stream, flags);
  if (rc) return rc;
  int n_launch = b2r_last_launch_count();
  const bool fused_ok = nq <= 4 && k <= kTailMaxK && n_launch == 1 && !tunable("hybrid.no_fused_tail");
  if (fused_ok) {
    HybridTail tail;
    memset(&tail, 0, sizeof(tail));
    tail.mode = world > 1 ? 2 : 1;
    tail.trigger_next = (flags & B2R_LAUNCH_OVERLAP_PREVIOUS) ? 1 : 0;
    tail.D = D_dev;
    tail.I = reinterpret_cast<const long long*>(I_dev);
    tail.F = F_dev;
    tail.O = reinterpret_cast<long long*>(O_dev);
    if (world > 1) {
      for (int p = 0; p < 8; ++p) tail.x.peers[p] = p < world ? static_cast<unsigned char*>(peer_bufs[p]) : nullptr;
      tail.x.rank = rank;
      tail.x.world = world;
      tail.x.nq_max = nq_max;
      tail.x.k_max = k_max;
      tail.x.step = step;
      tail.x.status = status;
      const int tmo_ms = tunable("exchange.timeout_ms");
      tail.x.timeout_ns = static_cast<unsigned long long>(tmo_ms > 0 ? tmo_ms : 10000) * 1000000ull;
    }
    rc = bm25_topk_with_tail(bm25, tids_dev, offs_dev, nq, k, S_dev, J_dev, id_offset, stream, true, tail);
    if (rc == B2R_OK) {
      if (launches) *launches = n_launch + b2r_last_launch_count();
      return B2R_OK;
    }
    if (rc != B2R_ERR_UNSUPPORTED) return rc;  // UNSUPPORTED: nothing was launched, take the three-kernel path
  }
  rc = b2r_bm25_topk_dev_ex(bm25, tids_dev, offs_dev, nullptr, nq, k, S_dev, J_dev, id_offset, stream, B2R_LAUNCH_OVERLAP_PREVIOUS);
  if (rc) return rc;
  n_launch += b2r_last_launch_count();
  if (world > 1) {
    rc = b2r_hybrid_exchange_fuse_dev(D_dev, I_dev, S_dev, J_dev, nq, k, peer_bufs, rank, world, nq_max, k_max, step, F_dev, O_dev,
                                      status, stream);
  } else {
    rc = b2r_hybrid_fuse_dev(D_dev, I_dev, S_dev, J_dev, nq, k, k, k, F_dev, O_dev, nullptr, stream);
  }
  if (launches) *launches = n_launch + 1;
  return rc;
}

static int slot_reserve(HybridSlot* c, size_t in_bytes, size_t nk) {
  if (c->in_cap < in_bytes) {
    const size_t cap = std::max<size_t>(in_bytes, 1 << 16);
    if (c->in_pin) cudaFreeHost(c->in_pin);
    if (c->in_dev) cudaFree(c->in_dev);
    c->in_pin = c->in_dev = nullptr;
    c->in_cap = 0;
    B2R_CUDA(cudaMallocHost(&c->in_pin, cap));
    B2R_CUDA(cudaMalloc(&c->in_dev, cap));
    c->in_cap = cap;
  }
  if (c->out_cap < nk) {
    const size_t cap = std::max<size_t>(nk, 4096);
    if (c->out_pin) cudaFreeHost(c->out_pin);
    if (c->out_dev) cudaFree(c->out_dev);
    c->out_pin = c->out_dev = nullptr;
    c->out_cap = 0;
    B2R_CUDA(cudaHostAlloc(&c->out_pin, cap * 12 + 64, cudaHostAllocMapped));  // + the exchange kernel's status word
    B2R_CUDA(cudaHostGetDevicePointer(&c->out_pin_dev, c->out_pin, 0));
    B2R_CUDA(cudaMalloc(&c->out_dev, cap * 28));
    c->out_cap = cap;
  }
  if (!c->done) B2R_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  if (!c->in_ready) B2R_CUDA(cudaEventCreateWithFlags(&c->in_ready, cudaEventDisableTiming));
  return B2R_OK;
}

extern "C" {

int b2r_hybrid_ctx_create(int device, b2r_hybrid_ctx** out) {
  if (!out) return fail(B2R_ERR_INVALID, "b2r_hybrid_ctx_create: out is NULL");
  *out = nullptr;
  DeviceGuard g(device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_hybrid_ctx_create: cudaSetDevice failed");
  b2r_hybrid_ctx* c = new b2r_hybrid_ctx();
  c->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return fail(B2R_ERR_CUDA, std::string("b2r_hybrid_ctx_create: ") + cudaGetErrorString(e));
  }
  *out = c;
  return B2R_OK;
}

int b2r_hybrid_ctx_destroy(b2r_hybrid_ctx* c) {
  if (!c) return B2R_OK;
  DeviceGuard g(c->device);
  if (c->stream) {
    cudaStreamSynchronize(c->stream);
    cudaStreamDestroy(c->stream);
  }
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  for (HybridSlot& sl : c->slot) {
    if (sl.in_pin) cudaFreeHost(sl.in_pin);
    if (sl.in_dev) cudaFree(sl.in_dev);
    if (sl.out_pin) cudaFreeHost(sl.out_pin);
    if (sl.out_dev) cudaFree(sl.out_dev);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.in_ready) cudaEventDestroy(sl.in_ready);
  }
  delete c;
  return B2R_OK;
}

int b2r_hybrid_ctx_bytes(const b2r_hybrid_ctx* c, size_t* h2d, size_t* d2h) {
  if (!c) return fail(B2R_ERR_INVALID, "b2r_hybrid_ctx_bytes: NULL context");
  if (h2d) *h2d = c->last_h2d;
  if (d2h) *d2h = c->last_d2h;
  return B2R_OK;
}

int b2r_hybrid_submit_host(b2r_hybrid_ctx* c, b2r_flat* flat, b2r_bm25* bm25, const float* q, int nq, int d,
                           const int32_t* term_ids, const int32_t* term_offsets, int k, int64_t id_offset,
                           void* const* peer_bufs, int rank, int world, int nq_max, int k_max, uint32_t step, int* ticket) {
  if (!c || !flat || !bm25 || !ticket) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: NULL handle");
  if (nq < 1 || d < 1 || k < 1 || k > B2R_MAX_K) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: bad sizes");
  if (!q || !term_offsets) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: NULL buffer");
  int fd = 0;
  int rc = b2r_flat_dim(flat, &fd);
  if (rc) return rc;
  if (fd != d) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: query width does not match the index");
  const int n_terms = term_offsets[nq];
  if (term_offsets[0] != 0 || n_terms < 0 || (n_terms > 0 && !term_ids)) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: bad term offsets");
  // everything that can be rejected is rejected HERE, before anything is enqueued: in the multi-GPU step a rank that
  // fails after its peers launched leaves them waiting (for exchange.timeout_ms) on a kernel that never comes
  rc = bm25_check_terms(bm25, term_ids, n_terms);
  if (rc) return rc;
  int64_t n_rows = 0;
  rc = b2r_flat_ntotal(flat, &n_rows);
  if (rc) return rc;
  if (id_offset < 0 || id_offset + n_rows >= 0xFFFFFFF0LL) return fail(B2R_ERR_UNSUPPORTED, "b2r_hybrid_submit_host: global ids must stay below 2^32-16 (the fusion packs them in 32 bits)");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: bad rank / world");
  if (world > 1) {
    if (!peer_bufs) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: world > 1 needs the peer buffers");
    for (int p = 0; p < world; ++p)
      if (!peer_bufs[p]) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: NULL peer buffer");
    if (nq > nq_max || k > k_max || k > 256) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: nq / k exceed the exchange buffer geometry");
    if (step == 0) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: step must be >= 1");
  }
  DeviceGuard g(c->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_hybrid_submit_host: cudaSetDevice failed");
  std::lock_guard<std::mutex> lock(c->mu);
  HybridSlot* sl = &c->slot[c->next];
  if (sl->busy) return fail(B2R_ERR_INVALID, "b2r_hybrid_submit_host: both slots of this context are in flight: collect one first");
  const size_t qb = static_cast<size_t>(nq) * d * 4, ob = (static_cast<size_t>(nq) + 1) * 4;
  const size_t tb = static_cast<size_t>(std::max(n_terms, 1)) * 4;
  const size_t in_bytes = qb + ob + tb;
  const size_t nk = static_cast<size_t>(nq) * k;
  rc = slot_reserve(sl, in_bytes, nk);
  if (rc) return rc;
  memcpy(sl->in_pin, q, qb);
  memcpy(sl->in_pin + qb, term_offsets, ob);
  if (n_terms) memcpy(sl->in_pin + qb + ob, term_ids, static_cast<size_t>(n_terms) * 4);
  const size_t h2d = qb + ob + static_cast<size_t>(n_terms) * 4;
  // the copy runs on its own stream, so the inputs of this question travel while the previous question's step is still
  // executing; the step waits for the event, not for the copy engine's turn behind the previous kernels
  if (tunable("hybrid.no_copy_stream")) {
    B2R_CUDA(cudaMemcpyAsync(sl->in_dev, sl->in_pin, h2d, cudaMemcpyHostToDevice, c->stream));
  } else {
    B2R_CUDA(cudaMemcpyAsync(sl->in_dev, sl->in_pin, h2d, cudaMemcpyHostToDevice, c->copy_stream));
    B2R_CUDA(cudaEventRecord(sl->in_ready, c->copy_stream));
    B2R_CUDA(cudaStreamWaitEvent(c->stream, sl->in_ready, 0));
  }
  const float* q_dev = reinterpret_cast<const float*>(sl->in_dev);
  const int32_t* offs_dev = reinterpret_cast<const int32_t*>(sl->in_dev + qb);
  const int32_t* tids_dev = reinterpret_cast<const int32_t*>(sl->in_dev + qb + ob);
  // results: the fusion writes O[nq*k] int64 and F[nq*k] float directly into mapped pinned host memory
  // (12 * nq * k bytes of posted PCIe writes instead of two device->host copy operations)
  int64_t* O_dev = reinterpret_cast<int64_t*>(sl->out_pin_dev);
  float* F_dev = reinterpret_cast<float*>(sl->out_pin_dev + nk * 8);
  unsigned char* mid = sl->out_dev;
  int64_t* I_dev = reinterpret_cast<int64_t*>(mid);
  double* S_dev = reinterpret_cast<double*>(mid + sl->out_cap * 8);
  int64_t* J_dev = reinterpret_cast<int64_t*>(mid + sl->out_cap * 16);
  float* D_dev = reinterpret_cast<float*>(mid + sl->out_cap * 24);
  *reinterpret_cast<volatile uint32_t*>(sl->out_pin + sl->out_cap * 12) = 0;
  rc = hybrid_step_dev(flat, bm25, q_dev, nq, tids_dev, offs_dev, k, id_offset, peer_bufs, rank, world, nq_max, k_max, step,
                       reinterpret_cast<uint32_t*>(sl->out_pin_dev + sl->out_cap * 12), D_dev, I_dev, S_dev, J_dev, F_dev, O_dev,
                       c->stream, tunable("hybrid.host_chain") ? B2R_LAUNCH_OVERLAP_PREVIOUS : 0u, nullptr);
  if (rc) return rc;
  B2R_CUDA(cudaEventRecord(sl->done, c->stream));
  sl->busy = true;
  sl->nk = nk;
  sl->h2d = h2d;
  sl->world = world;
  sl->step = step;
  *ticket = c->next;
  c->next = (c->next + 1) % kHybridSlots;
  return B2R_OK;
}

int b2r_hybrid_collect_host(b2r_hybrid_ctx* c, int ticket, float* F, int64_t* O) {
  if (!c || !F || !O) return fail(B2R_ERR_INVALID, "b2r_hybrid_collect_host: NULL argument");
  if (ticket < 0 || ticket >= kHybridSlots) return fail(B2R_ERR_INVALID, "b2r_hybrid_collect_host: bad ticket");
  DeviceGuard g(c->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_hybrid_collect_host: cudaSetDevice failed");
  HybridSlot* sl = &c->slot[ticket];
  {
    std::lock_guard<std::mutex> lock(c->mu);
    if (!sl->busy) return fail(B2R_ERR_INVALID, "b2r_hybrid_collect_host: nothing was submitted under this ticket");
  }
  cudaError_t e = cudaEventSynchronize(sl->done);
  std::lock_guard<std::mutex> lock(c->mu);
  sl->busy = false;
  if (e != cudaSuccess) return fail(B2R_ERR_CUDA, std::string("b2r_hybrid_collect_host: ") + cudaGetErrorString(e));
  const uint32_t status = *reinterpret_cast<volatile uint32_t*>(sl->out_pin + sl->out_cap * 12);
  if (sl->world > 1 && status != 0) {
    return fail(B2R_ERR_CUDA, "b2r_hybrid_collect_host: the exchange timed out waiting for rank " + std::to_string(status - 1) +
                                  " (step " + std::to_string(sl->step) + "): a peer failed or the ranks are out of lockstep");
  }
  memcpy(O, sl->out_pin, sl->nk * 8);
  memcpy(F, sl->out_pin + sl->nk * 8, sl->nk * 4);
  c->last_h2d = sl->h2d;
  c->last_d2h = sl->nk * 12;
  return B2R_OK;
}

int b2r_hybrid_search_host(b2r_hybrid_ctx* c, b2r_flat* flat, b2r_bm25* bm25, const float* q, int nq, int d,
                           const int32_t* term_ids, const int32_t* term_offsets, int k, int64_t id_offset,
                           void* const* peer_bufs, int rank, int world, int nq_max, int k_max, uint32_t step, float* F,
                           int64_t* O) {
  if (!F || !O) return fail(B2R_ERR_INVALID, "b2r_hybrid_search_host: NULL buffer");
  int ticket = -1;
  int rc = b2r_hybrid_submit_host(c, flat, bm25, q, nq, d, term_ids, term_offsets, k, id_offset, peer_bufs, rank, world, nq_max,
                                  k_max, step, &ticket);
  if (rc) return rc;
  return b2r_hybrid_collect_host(c, ticket, F, O);
}

// The resident step (see hybrid_step_dev above): q_dev[nq,d], term ids / offsets, and all outputs live on the device;
// D/I (dense hits), S/J (BM25 hits) and F/O (fused) are [nq,k].  world == 1: peer_bufs may be NULL.  status: see
// b2r_hybrid_exchange_fuse_dev.  launches (optional) receives the number of kernels enqueued.
int b2r_hybrid_step_dev(b2r_flat* flat, b2r_bm25* bm25, const float* q_dev, int nq, const int32_t* term_ids_dev,
                        const int32_t* term_offsets_dev, int k, int64_t id_offset, void* const* peer_bufs, int rank, int world,
                        int nq_max, int k_max, uint32_t step, uint32_t* status, float* D_dev, int64_t* I_dev, double* S_dev,
                        int64_t* J_dev, float* F_dev, int64_t* O_dev, void* stream, unsigned flags, int* launches) {
  if (!flat || !bm25) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: NULL handle");
  if (nq < 1 || k < 1 || k > B2R_MAX_K) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: bad sizes");
  if (!q_dev || !term_offsets_dev || !D_dev || !I_dev || !S_dev || !J_dev || !F_dev || !O_dev) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: NULL buffer");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: bad rank / world");
  if (world > 1) {
    if (!peer_bufs) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: world > 1 needs the peer buffers");
    if (nq > nq_max || k > k_max || k > 256) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: nq / k exceed the exchange buffer geometry");
    if (step == 0) return fail(B2R_ERR_INVALID, "b2r_hybrid_step_dev: step must be >= 1");
  }
  int64_t n_rows = 0;
  int rc = b2r_flat_ntotal(flat, &n_rows);
  if (rc) return rc;
  if (id_offset < 0 || id_offset + n_rows >= 0xFFFFFFF0LL) return fail(B2R_ERR_UNSUPPORTED, "b2r_hybrid_step_dev: global ids must stay below 2^32-16");
  return hybrid_step_dev(flat, bm25, q_dev, nq, term_ids_dev, term_offsets_dev, k, id_offset, peer_bufs, rank, world, nq_max, k_max,
                         step, status, D_dev, I_dev, S_dev, J_dev, F_dev, O_dev, static_cast<cudaStream_t>(stream), flags, launches);
}

}  // extern "C"
