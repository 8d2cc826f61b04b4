// flat_ip.cu — exact inner-product top-k over a flat fp32 row store (the role faiss.IndexFlatIP
// plays in the reference: src/ingestion.py:79-84 builds it, src/retrieval.py:1140 searches it).
//
// Kernel K1 `flat_scan_tma_kernel` (HBM-bound, one pass over the rows per launch):
//   * persistent CTAs, grid = SMs x resident CTAs; a tile is tile_rows consecutive rows (one contiguous span of the
//     row-major matrix); tiles are handed out dynamically in chunks of 2 from a global counter (the first chunk of every
//     CTA is static), so the SMs that stream faster take more tiles and the kernel does not end with a straggler
//   * one producer lane per CTA streams tiles global->shared with 1-D bulk async copies
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) through a ring of `stages` buffers
//     guarded by full/empty mbarriers
//   * CW consumer warps: lanes split the row into 128-bit chunks (conflict-free LDS.128), the
//     query fragments live in registers (up to NQ queries share every byte loaded), fp32 FMA,
//     xor-butterfly warp reduce
//   * fused select: each warp keeps, per query, a threshold key and a small shared-memory
//     candidate list; a row is appended only if it beats the threshold, and a full list is
//     bitonic-sorted and cut back to k.  The score vector never exists in HBM.
//   * epilogue: CTA-level merge of the warp lists -> per-CTA top-k in a global scratch; the last
//     CTA to finish (atomic ticket) merges all per-CTA lists and writes D / I.
//
// Algorithmic bytes per launch: 4 * n * d  (+ 4*nq*d query, + 12*nq*k result) — see DESIGN.md.
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

#include <string.h>

#include "b2r_common.cuh"

namespace b2r {

struct ScanArgs {
  const float* xb;    // [n, d] row-major
  long long n;
  int d;
  const float* q;     // [nq, d]
  int nq;             // active queries in this launch (1..NQ)
  int k;
  int cap;            // per-(warp, query) candidate capacity, power of two >= 2k
  int stages;
  int tile_rows;
  long long n_tiles;
  u64* partial;       // [NQ][gridDim.x][k] per-CTA sorted keys
  u32* ticket;        // self-resetting arrival counter; ticket[1] is the dynamic tile counter (chunks handed out so far)
  int dyn_chunk;      // > 0: tiles are handed out dynamically in chunks of this many tiles (after a first static round)
  u32 dyn_base;       // value of the (never reset) tile counter at which this launch's hand-outs begin
  float* D;           // [nq, k]
  long long* I;       // [nq, k]
  long long id_offset;
  int sel_cap;        // final-select scratch capacity in keys (power of two >= 2k)
};

// append `key` to a warp-private candidate list; compact to the best k when it fills up.
// Warp-uniform: all 32 lanes call it with identical arguments.
__device__ __forceinline__ void warp_list_push(u64* list, int cap, int k, int& cnt, u64& tau, u64 key,
                                               int lane) {
  if (lane == 0) list[cnt] = key;
  ++cnt;
  if (cnt == cap) {
    __syncwarp();
    warp_sort_desc(list, cap, lane);
    cnt = k;
    tau = list[k - 1];
    __syncwarp();
  }
}

// CTA epilogue shared by both scan kernels: merge the warp lists of every query into the per-CTA
// partial lists, then let the last CTA produce the final answer.
//   cand layout: [NQ][n_lists_per_q][cap], zero padded by the caller, n_lists_per_q*cap = pow2
template <int NQ>
__device__ __forceinline__ void scan_epilogue(const ScanArgs& a, u64* cand, int lists_per_q, u64* sel_buf,
                                              u64* ctl) {
  const int tid = threadIdx.x;
  const int region = lists_per_q * a.cap;
  // Scratch (the idle stage ring): the per-CTA merge and the last CTA's threshold step select by histogram instead of
  // sorting everything when the scratch has room for it: [out / lists: sel_cap / 2 keys][histogram].
  const int oc = next_pow2(2 * a.k);
  constexpr int kHistKeys = (kSelBins + 4 + 1) / 2;
  const bool use_hist = region > oc && a.sel_cap / 2 >= oc && a.sel_cap / 2 >= kHistKeys;
  u32* hist = reinterpret_cast<u32*>(sel_buf + a.sel_cap / 2);
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    if (qi < a.nq) {
      u64* reg = cand + static_cast<size_t>(qi) * region;
      const u64* best = reg;
      int n_best = region;
      if (use_hist) {
        block_topk_hist(reg, region, region, a.k, hist, sel_buf, oc);
        best = sel_buf;
        n_best = oc;
      } else {
        block_sort_desc(reg, region);
      }
      u64* dst = a.partial + (static_cast<size_t>(qi) * gridDim.x + blockIdx.x) * a.k;
      for (int i = tid; i < a.k; i += blockDim.x) dst[i] = i < n_best ? best[i] : 0ull;
      __syncthreads();
    }
  }
  __threadfence();
  __syncthreads();
  u32* s_last = reinterpret_cast<u32*>(ctl + 2);
  if (tid == 0) {
    const u32 t = atomicAdd(a.ticket, 1u);
    *s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (*s_last == 0) return;
  __threadfence();
  for (int qi = 0; qi < a.nq; ++qi) {
    if (use_hist) {
      block_select_lists(a.partial + static_cast<size_t>(qi) * gridDim.x * a.k, gridDim.x, a.k, a.k, sel_buf, a.sel_cap / 2,
                         ctl, hist);
    } else {
      block_select_lists(a.partial + static_cast<size_t>(qi) * gridDim.x * a.k, gridDim.x, a.k, a.k, sel_buf, a.sel_cap, ctl);
    }
    block_emit_keys(sel_buf, a.k, a.D + static_cast<size_t>(qi) * a.k, a.I + static_cast<size_t>(qi) * a.k,
                    a.id_offset);
    __syncthreads();
  }
  if (tid == 0) *a.ticket = 0;  // ready for the next launch on this workspace (ticket[1], the tile counter, keeps counting)
}

// ----------------------------------------------------------------------------------------------
// K1: TMA-staged scan.  d == 128 * DV.  Threads: CW consumer warps + 1 producer warp.
// ----------------------------------------------------------------------------------------------
template <int DV, int NQ, int CW>
__global__ void __launch_bounds__((CW + 1) * 32) flat_scan_tma_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // Programmatic dependent launch: a kernel enqueued behind this one with B2R_LAUNCH_OVERLAP_PREVIOUS (the
  // hybrid path's BM25) may start once every CTA of this grid is running — i.e. beside the scan, but only
  // after all of its persistent CTAs hold their SM.  No effect on ordinary launches.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int d = DV * 128;
  const int tile_floats = a.tile_rows * d;
  const u32 tile_bytes = static_cast<u32>(tile_floats) * 4u;

  float* stage0 = reinterpret_cast<float*>(smem_raw);
  u64* full = reinterpret_cast<u64*>(smem_raw + static_cast<size_t>(a.stages) * tile_bytes);
  u64* empty = full + a.stages;
  u64* ctl = empty + a.stages;  // 4 x u64 control words
  long long* tile_of = reinterpret_cast<long long*>(ctl + 4);  // [stages] tile index of the stage (-1: no more tiles)
  u64* cand = reinterpret_cast<u64*>(tile_of + a.stages);      // [NQ][CW][cap]

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == CW) {
    // ---------------- producer: one lane drives the bulk-copy engine ----------------
    if (lane == 0) {
      const u64 policy = l2_policy_evict_first();
      int s = 0;
      u32 ph = 0;
      // Tile order.  Static: tile t -> CTA t mod grid.  Dynamic (dyn_chunk > 0): chunks of dyn_chunk consecutive tiles; every
      // CTA starts with chunk blockIdx.x and then takes the next unassigned chunk from a global counter, fetched one
      // chunk ahead so that the atomic's latency hides behind the copies of the current chunk.  The fastest SMs then
      // take more tiles and no CTA is left finishing alone (measured: -12 us on a 1.25M-row shard, -1.7 % at 10M rows).
      // The counter is never reset: exactly max(n_chunks, grid) hand-outs happen per launch (every CTA's last fetch is
      // the one that tells it to stop), so the host gives each launch the counter value it starts from (dyn_base) — a
      // chained next launch may begin while this one's last CTA is still selecting.
      const int G = a.dyn_chunk;
      const long long n_chunks = G > 0 ? (a.n_tiles + G - 1) / G : 0;
      long long chunk = blockIdx.x, next_chunk = 0;
      if (G > 0) next_chunk = static_cast<long long>(gridDim.x) + static_cast<u32>(atomicAdd(a.ticket + 1, 1u) - a.dyn_base);
      long long t = G > 0 ? chunk * G : static_cast<long long>(blockIdx.x);
      int in_chunk = 0;
      for (;;) {
        if (G > 0) {
          if (in_chunk == G || t >= a.n_tiles) {  // chunk exhausted: move to the prefetched one, prefetch another
            chunk = next_chunk;
            if (chunk >= n_chunks) break;
            next_chunk = static_cast<long long>(gridDim.x) + static_cast<u32>(atomicAdd(a.ticket + 1, 1u) - a.dyn_base);
            t = chunk * G;
            in_chunk = 0;
          }
        } else if (t >= a.n_tiles) {
          break;
        }
        mbar_wait(&empty[s], ph ^ 1u);
        const long long row0 = t * a.tile_rows;
        const long long rows = min(static_cast<long long>(a.tile_rows), a.n - row0);
        const u32 bytes = static_cast<u32>(rows) * static_cast<u32>(d) * 4u;
        tile_of[s] = t;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(stage0 + static_cast<size_t>(s) * tile_floats, a.xb + row0 * d, bytes, &full[s], policy);
        if (++s == a.stages) {
          s = 0;
          ph ^= 1u;
        }
        if (G > 0) {
          ++t;
          ++in_chunk;
        } else {
          t += gridDim.x;
        }
      }
      // end marker for the consumers
      mbar_wait(&empty[s], ph ^ 1u);
      tile_of[s] = -1;
      mbar_arrive(&full[s]);
    }
  } else {
    // ---------------- consumers ----------------
    float4 qf[NQ][DV];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      const float* qp = a.q + static_cast<size_t>(qi < a.nq ? qi : a.nq - 1) * d;
#pragma unroll
      for (int j = 0; j < DV; ++j) qf[qi][j] = *reinterpret_cast<const float4*>(qp + (lane + 32 * j) * 4);
    }
    int cnt[NQ];
    u64 tau[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      cnt[qi] = 0;
      tau[qi] = 0;
    }
    int s = 0;
    u32 ph = 0;
    // lane-constant bookkeeping of the multi-query path: after the transposed reduction lane l owns
    // the dot product number vi = bits 4,3,2 of l, i.e. (row slot vi / NQ, query vi % NQ)
    constexpr int R = NQ > 1 ? 8 / NQ : 1;  // rows processed together
    const int vi = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const int my_q = vi % NQ, my_r = vi / NQ;
    u64 tau_mine = 0;
    for (;;) {
      mbar_wait(&full[s], ph);
      const long long t = tile_of[s];
      if (t < 0) break;
      const float* tile = stage0 + static_cast<size_t>(s) * tile_floats;
      const long long row0 = t * a.tile_rows;
      const int rows = static_cast<int>(min(static_cast<long long>(a.tile_rows), a.n - row0));
      if constexpr (NQ == 1) {
        for (int r = warp; r < rows; r += CW) {
          const float* row = tile + r * d;
          float4 x[DV];
#pragma unroll
          for (int j = 0; j < DV; ++j) x[j] = lds128(row + (lane + 32 * j) * 4);
          const u32 rid = static_cast<u32>(row0 + r);
          // two accumulators (even / odd chunks) to shorten the dependent FMA chain
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int j = 0; j < DV; ++j) {
            float& acc = (j & 1) ? a1 : a0;
            acc = fmaf(x[j].x, qf[0][j].x, acc);
            acc = fmaf(x[j].y, qf[0][j].y, acc);
            acc = fmaf(x[j].z, qf[0][j].z, acc);
            acc = fmaf(x[j].w, qf[0][j].w, acc);
          }
          float sum = a0 + a1;
#pragma unroll
          for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
          const u64 key = make_key(sum, rid);
          if (key > tau[0]) {
            warp_list_push(cand + static_cast<size_t>(warp) * a.cap, a.cap, a.k, cnt[0], tau[0], key, lane);
          }
        }
      } else {
        // R rows x NQ queries = 8 dot products per pass: 8 independent FMA chains, then a
        // transposed reduction (4 + 2 + 1 + 1 + 1 shuffles instead of 8 x 5)
        for (int r0 = warp; r0 < rows; r0 += CW * R) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
          for (int j = 0; j < DV; ++j) {
            float4 x[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
              // rows past the end of a short tile re-read row r0 (their results are discarded)
              const int r = (r0 + rr * CW < rows) ? r0 + rr * CW : r0;
              x[rr] = lds128(tile + r * d + (lane + 32 * j) * 4);
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
#pragma unroll
              for (int qi = 0; qi < NQ; ++qi) {
                float& acc = v[rr * NQ + qi];
                acc = fmaf(x[rr].x, qf[qi][j].x, acc);
                acc = fmaf(x[rr].y, qf[qi][j].y, acc);
                acc = fmaf(x[rr].z, qf[qi][j].z, acc);
                acc = fmaf(x[rr].w, qf[qi][j].w, acc);
              }
            }
          }
          // transposed reduction: halve the number of live values at each of the first 3 steps
          const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
          float w4[4], w2[2], tot;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = up16 ? v[i] : v[i + 4];
            const float keep = up16 ? v[i + 4] : v[i];
            w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = up8 ? w4[i] : w4[i + 2];
            const float keep = up8 ? w4[i + 2] : w4[i];
            w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          {
            const float send = up4 ? w2[0] : w2[1];
            const float keep = up4 ? w2[1] : w2[0];
            tot = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          tot += __shfl_xor_sync(0xffffffffu, tot, 2);
          tot += __shfl_xor_sync(0xffffffffu, tot, 1);
          const int my_row = r0 + my_r * CW;
          const u64 key = make_key(tot, static_cast<u32>(row0 + my_row));
          const bool want = (my_row < rows) && (my_q < a.nq) && (key > tau_mine);
          unsigned m = __ballot_sync(0xffffffffu, want) & 0x11111111u;  // one lane per value
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const u64 kk = __shfl_sync(0xffffffffu, key, src);
            const int q2 = (((src >> 4) & 1) * 4 + ((src >> 3) & 1) * 2 + ((src >> 2) & 1)) % NQ;
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
              if (q2 == qi && kk > tau[qi]) {
                warp_list_push(cand + (static_cast<size_t>(qi) * CW + warp) * a.cap, a.cap, a.k, cnt[qi], tau[qi], kk,
                               lane);
                if (my_q == qi) tau_mine = tau[qi];
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == a.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    // zero-pad the unused tail of every list so the CTA-wide sort sees a dense power of two
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      u64* list = cand + (static_cast<size_t>(qi) * CW + warp) * a.cap;
      for (int i = cnt[qi] + lane; i < a.cap; i += 32) list[i] = 0;
    }
  }
  // launched beside the tail of the previous kernel (B2R_LAUNCH_OVERLAP_PREVIOUS): nothing has been written so far;
  // everything below writes workspace / results, so the previous kernel must have completed (no-op otherwise)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  // the stage ring is idle now (every issued copy was consumed): reuse it as select scratch
  scan_epilogue<NQ>(a, cand, CW, reinterpret_cast<u64*>(smem_raw), ctl);
}

// ----------------------------------------------------------------------------------------------
// K1g: generic scan for any d (no TMA staging: direct coalesced loads, query in shared memory).
// Same select machinery; used when d is not one of the templated widths.
// ----------------------------------------------------------------------------------------------
template <int NQ, int NW>
__global__ void __launch_bounds__(NW * 32) flat_scan_generic_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // see flat_scan_tma_kernel
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int d = a.d;
  // layout: [sel_buf: sel_cap keys][ctl: 4][cand: NQ*NW*cap][q: NQ*d floats]
  u64* sel_buf = reinterpret_cast<u64*>(smem_raw);
  u64* ctl = sel_buf + a.sel_cap;
  u64* cand = ctl + 4;
  float* qs = reinterpret_cast<float*>(cand + static_cast<size_t>(NQ) * NW * a.cap);
  for (int i = tid; i < NQ * d; i += blockDim.x) {
    const int qi = i / d;
    qs[i] = a.q[static_cast<size_t>(qi < a.nq ? qi : a.nq - 1) * d + (i - qi * d)];
  }
  __syncthreads();
  int cnt[NQ];
  u64 tau[NQ];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    cnt[qi] = 0;
    tau[qi] = 0;
  }
  const bool vec4 = (d & 3) == 0;
  const long long gw = static_cast<long long>(blockIdx.x) * NW + warp;
  const long long nw_total = static_cast<long long>(gridDim.x) * NW;
  for (long long r = gw; r < a.n; r += nw_total) {
    const float* row = a.xb + r * d;
    float acc[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) acc[qi] = 0.f;
    if (vec4) {
      for (int c = lane; c < (d >> 2); c += 32) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(row) + c);
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          const float4 qv = *reinterpret_cast<const float4*>(qs + qi * d + c * 4);
          acc[qi] = fmaf(x.x, qv.x, acc[qi]);
          acc[qi] = fmaf(x.y, qv.y, acc[qi]);
          acc[qi] = fmaf(x.z, qv.z, acc[qi]);
          acc[qi] = fmaf(x.w, qv.w, acc[qi]);
        }
      }
    } else {
      for (int c = lane; c < d; c += 32) {
        const float x = __ldg(row + c);
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) acc[qi] = fmaf(x, qs[qi * d + c], acc[qi]);
      }
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      float sum = acc[qi];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
      const u64 key = make_key(sum, static_cast<u32>(r));
      if (qi < a.nq && key > tau[qi]) {
        warp_list_push(cand + (static_cast<size_t>(qi) * NW + warp) * a.cap, a.cap, a.k, cnt[qi], tau[qi], key,
                       lane);
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    u64* list = cand + (static_cast<size_t>(qi) * NW + warp) * a.cap;
    for (int i = cnt[qi] + lane; i < a.cap; i += 32) list[i] = 0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // see flat_scan_tma_kernel: results are written after the previous kernel
  __syncthreads();
  scan_epilogue<NQ>(a, cand, NW, sel_buf, ctl);
}

// rows gather: out[i, :] = xb[ids[i], :]
__global__ void gather_rows_kernel(const float* __restrict__ xb, int d, const long long* __restrict__ ids, int m,
                                   float* __restrict__ out) {
  const int i = blockIdx.x;
  if (i >= m) return;
  const float* src = xb + ids[i] * static_cast<long long>(d);
  float* dst = out + static_cast<long long>(i) * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) dst[c] = src[c];
}

// fill D/I with the faiss "no result" padding (used when the index is empty)
__global__ void fill_empty_kernel(float* D, long long* I, long long count) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) {
    D[i] = -3.4028235e38f;
    I[i] = -1;
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static const int kMaxNQ = 4;           // queries sharing one pass over the rows
static const int kMaxGrid = 148 * 4;   // upper bound used to size workspaces
static const size_t kSmemLimit = 227 * 1024;

struct Workspace {
  unsigned char* base = nullptr;  // [256 B ticket area][partial keys]
  size_t bytes = 0;
};

// per-CTA partial lists of the scans; also the candidate lists of the tensor-core passes (148 CTAs x 256 queries x 128 keys)
static const size_t kPartialBytes = static_cast<size_t>(2) * kMaxNQ * kMaxGrid * B2R_MAX_K * sizeof(u64);

// csrc/flat_gemm.cu (tensor-core path; its scratch follows the partial lists in the workspace)
bool gemm_path_supported(int d, int k);
int gemm_max_batch(int k);
size_t gemm_scratch_bytes();
int gemm_max_row_norm2(const float* rows, long long n, int d, float* out_dev, cudaStream_t stream);
int gemm_search_chunk(const float* rows, const void* rows_bf16, long long n, int d, const float* q_dev, int nq, int k,
                      float* D_dev, long long* I_dev, long long id_offset, unsigned char* scratch, u64* partial,
                      size_t partial_bytes, const float* max_norm2_dev, int num_sms, cudaStream_t stream, int* launches);
// csrc/flat_gemm_wide.cu
int rows_to_bf16(const float* rows, long long n, int d, void* out, int num_sms, cudaStream_t stream);

static size_t workspace_bytes() { return 256 + kPartialBytes + gemm_scratch_bytes(); }

struct HostCtx {
  cudaStream_t stream = nullptr;
  float* q_dev = nullptr;
  float* q_pin = nullptr;
  size_t q_cap = 0;  // floats
  unsigned char* out_dev = nullptr;
  unsigned char* out_pin = nullptr;      // mapped pinned memory
  unsigned char* out_pin_dev = nullptr;  // its device alias: small results are stored straight into host memory
  size_t out_cap = 0;  // bytes
};

}  // namespace b2r

using namespace b2r;

struct b2r_flat {
  int d = 0;
  int device = 0;
  long long n = 0;
  long long capacity = 0;
  float* rows = nullptr;
  bool owns = true;
  int num_sms = 148;
  std::mutex mu;
  float* max_norm2 = nullptr;      // device scalar: max |row|^2 (error bound of the tf32 candidate pass)
  long long max_norm2_rows = -1;   // number of rows it was computed for
  // Concurrent host callers are COMBINED: a caller that finds a search of this index in flight queues its request; the
  // caller in charge ("leader") then runs everything queued as ONE batch — one pass over the rows for all of them
  // instead of one pass each (the reference fans (report, query) pairs out over a thread pool onto the same index
  // object, src/retrieval.py:1498-1504).  A lone caller never waits.
  struct HostReq {
    const float* q;
    int nq, k;
    float* D;
    int64_t* I;
    int rc = 0;
    bool done = false;
    std::string err;
  };
  std::mutex bq_mu;
  std::condition_variable bq_cv;
  std::deque<HostReq*> bq;
  bool bq_leader = false;
  long long combined_batches = 0, combined_requests = 0;  // statistics (b2r_flat_combine_stats)
  bool want_bf16 = false;          // b2r_flat_set_candidate_bf16: keep a bf16 copy of the rows for the candidate pass
  void* rows_bf16 = nullptr;
  long long bf16_rows = -1;        // number of rows the copy covers
};

namespace b2r {

static int get_workspace(b2r_flat* h, cudaStream_t stream, Workspace* out) {
  // shared by every index searching on this stream (ws_pool_get): [256 B ticket area][partial keys][gemm scratch]
  return ws_pool_get(h->device, stream, 0, workspace_bytes(), 256, &out->base, &out->bytes);
}

// max |row|^2 over the shard, computed once per index size on the searching stream
static int ensure_row_norm(b2r_flat* h, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->max_norm2) B2R_CUDA(cudaMalloc(&h->max_norm2, 256));
  if (h->max_norm2_rows != h->n) {
    int rc = gemm_max_row_norm2(h->rows, h->n, h->d, h->max_norm2, stream);
    if (rc) return rc;
    // later searches may run on other streams: make the value visible before anyone else uses it
    B2R_CUDA(cudaStreamSynchronize(stream));
    h->max_norm2_rows = h->n;
  }
  return B2R_OK;
}

// bf16 copy of the rows (optional), rebuilt when the index has grown
static int ensure_bf16(b2r_flat* h, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->want_bf16 || h->bf16_rows == h->n) return B2R_OK;
  if (h->rows_bf16) cudaFree(h->rows_bf16);
  h->rows_bf16 = nullptr;
  h->bf16_rows = -1;
  B2R_CUDA(cudaMalloc(&h->rows_bf16, static_cast<size_t>(h->n) * h->d * 2));
  int rc = rows_to_bf16(h->rows, h->n, h->d, h->rows_bf16, h->num_sms, stream);
  if (rc) return rc;
  B2R_CUDA(cudaStreamSynchronize(stream));
  h->bf16_rows = h->n;
  return B2R_OK;
}

struct ScanPlan {
  bool tma;
  int dv, nq, cw, stages, tile_rows, cap, sel_cap, grid, threads;
  size_t smem;
  bool pdl = false;  // launch with programmatic stream serialization (B2R_LAUNCH_OVERLAP_PREVIOUS)
};

// Launch a scan kernel, optionally BESIDE the tail of the previous kernel of the stream: the scan's main loop only
// reads the rows and the query; everything it writes (partial lists, ticket, D / I) is written by its epilogue, which
// starts with griddepcontrol.wait — i.e. after the previous kernel has completed.  A stream of hybrid steps then
// overlaps the fusion / exchange kernel of step s (including its wait for the slowest rank) and the launch latency
// with the first tiles of step s + 1.
template <class Kern>
static int launch_scan(Kern kern, const ScanPlan& p, const ScanArgs& a, cudaStream_t stream) {
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(p.threads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  B2R_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

template <int DV, int NQ>
static int launch_tma_cw(const ScanPlan& p, const ScanArgs& a, cudaStream_t stream) {
  if (p.cw == 8) return launch_scan(flat_scan_tma_kernel<DV, NQ, 8>, p, a, stream);
  return launch_scan(flat_scan_tma_kernel<DV, NQ, 4>, p, a, stream);
}

template <int DV>
static int launch_tma_nq(const ScanPlan& p, const ScanArgs& a, cudaStream_t stream) {
  switch (p.nq) {
    case 1: return launch_tma_cw<DV, 1>(p, a, stream);
    case 2: return launch_tma_cw<DV, 2>(p, a, stream);
    default: return launch_tma_cw<DV, 4>(p, a, stream);
  }
}

static int launch_tma(const ScanPlan& p, const ScanArgs& a, cudaStream_t stream) {
  switch (p.dv) {
    case 1: return launch_tma_nq<1>(p, a, stream);
    case 2: return launch_tma_nq<2>(p, a, stream);
    case 3: return launch_tma_nq<3>(p, a, stream);
    case 4: return launch_tma_nq<4>(p, a, stream);
    case 6: return launch_tma_nq<6>(p, a, stream);
    case 8: return launch_tma_nq<8>(p, a, stream);
  }
  return fail(B2R_ERR_INVALID, "no TMA scan kernel for this width");
}

template <int NQ>
static int launch_generic_nq(const ScanPlan& p, const ScanArgs& a, cudaStream_t stream) {
  return launch_scan(flat_scan_generic_kernel<NQ, 8>, p, a, stream);
}

static int launch_generic(const ScanPlan& p, const ScanArgs& a, cudaStream_t stream) {
  switch (p.nq) {
    case 1: return launch_generic_nq<1>(p, a, stream);
    case 2: return launch_generic_nq<2>(p, a, stream);
    default: return launch_generic_nq<4>(p, a, stream);
  }
}

static int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static bool has_tma_width(int d) {
  if (d % 128) return false;
  const int dv = d / 128;
  return dv == 1 || dv == 2 || dv == 3 || dv == 4 || dv == 6 || dv == 8;
}

// Decide kernel, tile shape, pipeline depth and grid for `nq_left` pending queries.
static int make_plan(const b2r_flat* h, int nq_left, int k, ScanPlan* p) {
  const int d = h->d;
  p->cap = pow2_at_least(std::max(2 * k, 64));
  p->tma = has_tma_width(d) && !tunable("flat.force_generic");
  int nq = nq_left >= 3 ? 4 : nq_left;  // template instances: 1, 2, 4
  const int nq_cap = tunable("flat.max_nq");
  if (nq_cap > 0) nq = std::min(nq, nq_cap == 3 ? 2 : nq_cap);
  if (p->tma) {
    p->dv = d / 128;
    // defaults from the B200 sweep (profiles/r01_flat_sweep.md): 8 consumer warps, 1 CTA / SM;
    // single query: 32 KiB tiles x 4 stages; several queries per pass: 64 KiB tiles x 3 stages
    p->cw = tunable("flat.consumer_warps") == 4 ? 4 : 8;
    const int row_bytes = d * 4;
    int tile_rows = tunable("flat.tile_rows");
    if (tile_rows <= 0) tile_rows = std::max(p->cw, ((nq == 1 ? 32 : 64) * 1024) / row_bytes);
    tile_rows = std::max(p->cw, (tile_rows / p->cw) * p->cw);
    int ctas_per_sm = tunable("flat.ctas_per_sm");
    if (ctas_per_sm <= 0) ctas_per_sm = 1;
    int stages = tunable("flat.stages");
    if (stages <= 0) stages = nq == 1 ? 4 : 3;
    for (;;) {
      const size_t cand = static_cast<size_t>(nq) * p->cw * p->cap * sizeof(u64);
      const size_t fixed = static_cast<size_t>(3 * stages + 4) * sizeof(u64) + cand + 128;  // barriers, ctl, tile_of
      const size_t ring = static_cast<size_t>(stages) * tile_rows * row_bytes;
      const size_t total = ring + fixed;
      const size_t budget = kSmemLimit / ctas_per_sm - 1024;  // 1 KiB per-CTA system reservation
      // the ring doubles as the final-select scratch: it must hold at least 2k keys
      const bool ring_ok = ring >= static_cast<size_t>(2) * pow2_at_least(k) * sizeof(u64);
      if (total <= budget && ring_ok) {
        p->smem = total;
        p->stages = stages;
        p->tile_rows = tile_rows;
        p->sel_cap = 1;
        while (static_cast<size_t>(p->sel_cap) * 2 * sizeof(u64) <= ring) p->sel_cap <<= 1;
        const int cap_limit = tunable("select.buf_cap");
        if (cap_limit > 0) p->sel_cap = std::max(2 * pow2_at_least(k), std::min(p->sel_cap, pow2_at_least(cap_limit)));
        break;
      }
      // shrink in this order: queries per pass, CTAs per SM, pipeline depth, tile size
      if (nq > 1 && cand > budget / 4) nq >>= 1;
      else if (ctas_per_sm > 1) --ctas_per_sm;
      else if (stages > 2) --stages;
      else if (nq > 1) nq >>= 1;
      else if (tile_rows > p->cw) tile_rows = std::max(p->cw, tile_rows / 2 / p->cw * p->cw);
      else return fail(B2R_ERR_UNSUPPORTED, "flat scan: no shared-memory configuration fits");
    }
    p->nq = nq;
    p->threads = (p->cw + 1) * 32;
    const long long n_tiles = (h->n + p->tile_rows - 1) / p->tile_rows;
    p->grid = static_cast<int>(std::max(1LL, std::min<long long>(n_tiles, static_cast<long long>(h->num_sms) * ctas_per_sm)));
    p->grid = std::min(p->grid, kMaxGrid);
    return B2R_OK;
  }
  // generic path
  p->dv = 0;
  p->cw = 8;
  p->threads = 8 * 32;
  p->stages = 0;
  p->tile_rows = 1;
  for (;;) {
    p->sel_cap = pow2_at_least(std::max(2 * k, 2048));
    const int cap_limit = tunable("select.buf_cap");
    if (cap_limit > 0) p->sel_cap = std::max(2 * pow2_at_least(k), std::min(p->sel_cap, pow2_at_least(cap_limit)));
    const size_t total = static_cast<size_t>(p->sel_cap) * 8 + 32 + static_cast<size_t>(nq) * 8 * p->cap * 8 +
                         static_cast<size_t>(nq) * d * 4 + 128;
    if (total <= kSmemLimit - 1024) {
      p->smem = total;
      break;
    }
    if (nq > 1) nq >>= 1;
    else return fail(B2R_ERR_UNSUPPORTED, "flat scan: d too large for the generic kernel");
  }
  p->nq = nq;
  const int ctas_per_sm = std::max<int>(1, std::min<size_t>(4, kSmemLimit / (p->smem + 1024)));
  const long long want = (h->n + 7) / 8;
  p->grid = static_cast<int>(std::max(1LL, std::min<long long>(want, static_cast<long long>(h->num_sms) * ctas_per_sm)));
  p->grid = std::min(p->grid, kMaxGrid);
  return B2R_OK;
}

static int flat_search_dev_impl(b2r_flat* h, const float* q_dev, int nq, int k, float* D_dev, int64_t* I_dev,
                                int64_t id_offset, cudaStream_t stream, unsigned flags = 0u) {
  launch_counter() = 0;
  if (nq == 0) return B2R_OK;
  if (h->n == 0) {
    const long long count = static_cast<long long>(nq) * k;
    fill_empty_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, stream>>>(
        D_dev, reinterpret_cast<long long*>(I_dev), count);
    B2R_CUDA(cudaGetLastError());
    launch_counter() = 1;
    return B2R_OK;
  }
  Workspace ws;
  int rc = get_workspace(h, stream, &ws);
  if (rc) return rc;
  int done = 0;
  // batches go to the tensor cores (K2): one pass over the rows per <= 256 queries
  // smallest batch that goes to the tensor cores: 3.  One tensor-core pass beats a 4-query SIMT pass at every index
  // size (profiles/r01_small_index_sweep.json: 10k rows, 4 queries SIMT 115 us vs 8 queries on the tensor path
  // 86 us; 200k rows 337 vs 233 us); two queries stay on the SIMT kernel (81 us at 10k rows).
  int gemm_min = tunable("flat.gemm_min_nq");
  if (gemm_min <= 0) gemm_min = 3;
  int gemm_max = tunable("flat.gemm_max_n");
  if (gemm_max <= 0 || gemm_max > 256) gemm_max = 256;
  const bool use_gemm = nq >= gemm_min && gemm_path_supported(h->d, k) && !tunable("flat.force_generic") &&
                        !tunable("flat.no_gemm") && h->n >= 128;
  if (use_gemm) {
    gemm_max = std::min(gemm_max, gemm_max_batch(k));
    rc = ensure_row_norm(h, stream);
    if (rc) return rc;
    rc = ensure_bf16(h, stream);
    if (rc) return rc;
  }
  while (use_gemm && nq - done >= gemm_min) {
    const int chunk = std::min(nq - done, gemm_max);
    int launches = 0;
    rc = gemm_search_chunk(h->rows, h->bf16_rows == h->n ? h->rows_bf16 : nullptr, h->n, h->d,
                           q_dev + static_cast<size_t>(done) * h->d, chunk, k,
                           D_dev + static_cast<size_t>(done) * k,
                           reinterpret_cast<long long*>(I_dev) + static_cast<size_t>(done) * k, id_offset,
                           ws.base + 256 + kPartialBytes, reinterpret_cast<u64*>(ws.base + 256), kPartialBytes, h->max_norm2,
                           h->num_sms, stream, &launches);
    if (rc) return rc;
    launch_counter() += launches;
    done += chunk;
  }
  while (done < nq) {
    ScanPlan p;
    rc = make_plan(h, nq - done, k, &p);
    if (rc) return rc;
    ScanArgs a;
    a.xb = h->rows;
    a.n = h->n;
    a.d = h->d;
    a.q = q_dev + static_cast<size_t>(done) * h->d;
    a.nq = std::min(p.nq, nq - done);
    a.k = k;
    a.cap = p.cap;
    a.stages = p.stages;
    a.tile_rows = p.tile_rows;
    a.n_tiles = (h->n + p.tile_rows - 1) / p.tile_rows;
    const int dyn = tunable("flat.dynamic_tiles");   // 0: default (on, chunks of 2 tiles), < 0: static order, > 0: chunk size
    a.dyn_chunk = (!p.tma || dyn < 0) ? 0 : (dyn == 0 ? 2 : dyn);
    a.dyn_base = 0;
    a.ticket = reinterpret_cast<u32*>(ws.base);
    a.partial = reinterpret_cast<u64*>(ws.base + 256);
    a.D = D_dev + static_cast<size_t>(done) * k;
    a.I = reinterpret_cast<long long*>(I_dev) + static_cast<size_t>(done) * k;
    a.id_offset = id_offset;
    a.sel_cap = p.sel_cap;
    // only the first kernel of a call may overlap the caller's previous kernel (the tensor-core path never does)
    p.pdl = (flags & B2R_LAUNCH_OVERLAP_PREVIOUS) != 0u && done == 0 && !use_gemm;
    if (a.dyn_chunk > 0) {
      // sequence value and launch under one lock: the kernels must run in the order their counter bases were taken
      std::lock_guard<std::mutex> lock(ws_launch_mutex());
      const long long n_chunks = (a.n_tiles + a.dyn_chunk - 1) / a.dyn_chunk;
      const uint32_t handouts = static_cast<uint32_t>(std::max<long long>(n_chunks, p.grid));
      a.dyn_base = ws_pool_advance(h->device, stream, 0, handouts);
      rc = launch_tma(p, a, stream);
      if (rc) ws_pool_advance(h->device, stream, 0, 0u - handouts);  // nothing was enqueued: the counter did not move
    } else {
      rc = p.tma ? launch_tma(p, a, stream) : launch_generic(p, a, stream);
    }
    if (rc) return rc;
    ++launch_counter();
    done += a.nq;
  }
  return B2R_OK;
}

// Host contexts (a stream + staging) are pooled per DEVICE, not per index: a host search borrows one for the call, so
// the number of streams — and of the workspaces keyed by them — follows the callers' concurrency.  A retriever over
// thousands of per-report indexes (src/retrieval.py builds one faiss index per report) then holds a handful of
// contexts, not one per report.
static std::mutex g_ctx_mu;
static std::map<int, std::vector<HostCtx*>> g_free_ctx;

static int acquire_ctx(b2r_flat* h, HostCtx** out) {
  {
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    auto& pool = g_free_ctx[h->device];
    if (!pool.empty()) {
      *out = pool.back();
      pool.pop_back();
      return B2R_OK;
    }
  }
  HostCtx* c = new HostCtx();
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return fail(B2R_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  *out = c;
  return B2R_OK;
}

static void release_ctx(b2r_flat* h, HostCtx* c) {
  std::lock_guard<std::mutex> lock(g_ctx_mu);
  g_free_ctx[h->device].push_back(c);
}

static int ensure_ctx_buffers(HostCtx* c, size_t q_floats, size_t out_bytes) {
  if (q_floats > c->q_cap) {
    if (c->q_dev) cudaFree(c->q_dev);
    if (c->q_pin) cudaFreeHost(c->q_pin);
    c->q_dev = nullptr;
    c->q_pin = nullptr;
    c->q_cap = 0;
    const size_t cap = std::max<size_t>(q_floats, 16 * 1024);
    B2R_CUDA(cudaMalloc(&c->q_dev, cap * sizeof(float)));
    B2R_CUDA(cudaMallocHost(&c->q_pin, cap * sizeof(float)));
    c->q_cap = cap;
  }
  if (out_bytes > c->out_cap) {
    if (c->out_dev) cudaFree(c->out_dev);
    if (c->out_pin) cudaFreeHost(c->out_pin);
    c->out_dev = nullptr;
    c->out_pin = nullptr;
    c->out_cap = 0;
    const size_t cap = std::max<size_t>(out_bytes, 64 * 1024);
    B2R_CUDA(cudaMalloc(&c->out_dev, cap));
    B2R_CUDA(cudaHostAlloc(&c->out_pin, cap, cudaHostAllocMapped));
    B2R_CUDA(cudaHostGetDevicePointer(&c->out_pin_dev, c->out_pin, 0));
    c->out_cap = cap;
  }
  return B2R_OK;
}

}  // namespace b2r

// ----------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------
extern "C" {

int b2r_flat_create(int d, int device, b2r_flat** out) {
  if (!out) return fail(B2R_ERR_INVALID, "b2r_flat_create: out is NULL");
  *out = nullptr;
  if (d <= 0 || d > 16384) return fail(B2R_ERR_INVALID, "b2r_flat_create: d must be in [1, 16384]");
  int ndev = 0;
  B2R_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2R_ERR_INVALID, "b2r_flat_create: no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_flat_create: cudaSetDevice failed");
  cudaDeviceProp prop;
  B2R_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return fail(B2R_ERR_UNSUPPORTED, "libb2r is built for sm_100a (B200) only; found sm_" +
                                         std::to_string(prop.major) + std::to_string(prop.minor));
  }
  b2r_flat* h = new b2r_flat();
  h->d = d;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  *out = h;
  return B2R_OK;
}

int b2r_flat_add(b2r_flat* h, const float* rows, int64_t n, int rows_on_device) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_add: NULL handle");
  if (n < 0 || (n > 0 && !rows)) return fail(B2R_ERR_INVALID, "b2r_flat_add: bad rows / n");
  if (!h->owns) return fail(B2R_ERR_INVALID, "b2r_flat_add: index borrows its rows (adopt); cannot grow");
  if (n == 0) return B2R_OK;
  if (h->n + n > 0xFFFFFFF0LL) return fail(B2R_ERR_UNSUPPORTED, "b2r_flat_add: more than 2^32-16 rows in one shard");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_flat_add: cudaSetDevice failed");
  if (h->n + n > h->capacity) {
    long long cap = std::max<long long>(h->n + n, h->capacity + h->capacity / 2);
    float* nr = nullptr;
    B2R_CUDA(cudaMalloc(&nr, static_cast<size_t>(cap) * h->d * sizeof(float)));
    if (h->n) {
      cudaError_t e = cudaMemcpy(nr, h->rows, static_cast<size_t>(h->n) * h->d * sizeof(float), cudaMemcpyDeviceToDevice);
      if (e != cudaSuccess) {
        cudaFree(nr);
        return fail(B2R_ERR_CUDA, std::string("b2r_flat_add: copy: ") + cudaGetErrorString(e));
      }
    }
    B2R_CUDA(cudaDeviceSynchronize());
    if (h->rows) cudaFree(h->rows);
    h->rows = nr;
    h->capacity = cap;
  }
  B2R_CUDA(cudaMemcpy(h->rows + static_cast<size_t>(h->n) * h->d, rows, static_cast<size_t>(n) * h->d * sizeof(float),
                      rows_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  h->n += n;
  return B2R_OK;
}

int b2r_flat_adopt(b2r_flat* h, const float* dev_rows, int64_t n) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_adopt: NULL handle");
  if (h->n != 0 || h->rows) return fail(B2R_ERR_INVALID, "b2r_flat_adopt: index is not empty");
  if (n < 0 || (n > 0 && !dev_rows)) return fail(B2R_ERR_INVALID, "b2r_flat_adopt: bad rows / n");
  if (reinterpret_cast<uintptr_t>(dev_rows) & 15) return fail(B2R_ERR_INVALID, "b2r_flat_adopt: rows must be 16-byte aligned");
  if (n > 0xFFFFFFF0LL) return fail(B2R_ERR_UNSUPPORTED, "b2r_flat_adopt: more than 2^32-16 rows in one shard");
  h->rows = const_cast<float*>(dev_rows);
  h->n = n;
  h->capacity = n;
  h->owns = false;
  return B2R_OK;
}

int b2r_flat_ntotal(const b2r_flat* h, int64_t* out) {
  if (!h || !out) return fail(B2R_ERR_INVALID, "b2r_flat_ntotal: NULL argument");
  *out = h->n;
  return B2R_OK;
}

int b2r_flat_dim(const b2r_flat* h, int* out) {
  if (!h || !out) return fail(B2R_ERR_INVALID, "b2r_flat_dim: NULL argument");
  *out = h->d;
  return B2R_OK;
}

int b2r_flat_rows_dev(const b2r_flat* h, const float** out) {
  if (!h || !out) return fail(B2R_ERR_INVALID, "b2r_flat_rows_dev: NULL argument");
  *out = h->rows;
  return B2R_OK;
}

int b2r_flat_search_dev(b2r_flat* h, const float* q_dev, int nq, int k, float* D_dev, int64_t* I_dev,
                        int64_t id_offset, void* stream) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_search: NULL handle");
  if (nq < 0 || k < 1) return fail(B2R_ERR_INVALID, "b2r_flat_search: need nq >= 0 and k >= 1");
  if (k > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_flat_search: k > B2R_MAX_K (1024)");
  if (nq > 0 && (!q_dev || !D_dev || !I_dev)) return fail(B2R_ERR_INVALID, "b2r_flat_search: NULL buffer");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_flat_search: cudaSetDevice failed");
  return flat_search_dev_impl(h, q_dev, nq, k, D_dev, I_dev, id_offset, static_cast<cudaStream_t>(stream));
}

int b2r_flat_search_dev_ex(b2r_flat* h, const float* q_dev, int nq, int k, float* D_dev, int64_t* I_dev,
                           int64_t id_offset, void* stream, unsigned flags) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_search: NULL handle");
  if (nq < 0 || k < 1) return fail(B2R_ERR_INVALID, "b2r_flat_search: need nq >= 0 and k >= 1");
  if (k > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_flat_search: k > B2R_MAX_K (1024)");
  if (nq > 0 && (!q_dev || !D_dev || !I_dev)) return fail(B2R_ERR_INVALID, "b2r_flat_search: NULL buffer");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_flat_search: cudaSetDevice failed");
  return flat_search_dev_impl(h, q_dev, nq, k, D_dev, I_dev, id_offset, static_cast<cudaStream_t>(stream), flags);
}

// one staged batch for several host requests with the same k: queries back to back in, results scattered out
static int flat_search_host_batch(b2r_flat* h, const std::vector<b2r_flat::HostReq*>& batch, int k) {
  HostCtx* c = nullptr;
  int rc = acquire_ctx(h, &c);
  if (rc) return rc;
  size_t nq = 0;
  for (const b2r_flat::HostReq* r : batch) nq += static_cast<size_t>(r->nq);
  const size_t q_floats = nq * h->d;
  const size_t nk = nq * k;
  const size_t d_bytes = (nk * sizeof(float) + 255) & ~static_cast<size_t>(255);
  const size_t out_bytes = d_bytes + nk * sizeof(int64_t);
  rc = ensure_ctx_buffers(c, q_floats, out_bytes);
  if (rc == B2R_OK) {
    size_t at = 0;
    for (const b2r_flat::HostReq* r : batch) {
      memcpy(c->q_pin + at * h->d, r->q, static_cast<size_t>(r->nq) * h->d * sizeof(float));
      at += static_cast<size_t>(r->nq);
    }
    // small results (the retriever's case: a few queries x top-n) are written by the kernels straight into mapped
    // pinned host memory — one copy operation (~10 us of a ~55 us call) less; large ones go through a device buffer
    const bool zero_copy = out_bytes <= 256 * 1024;
    unsigned char* out_base = zero_copy ? c->out_pin_dev : c->out_dev;
    float* D_dev = reinterpret_cast<float*>(out_base);
    int64_t* I_dev = reinterpret_cast<int64_t*>(out_base + d_bytes);
    cudaError_t e = cudaMemcpyAsync(c->q_dev, c->q_pin, q_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
      rc = flat_search_dev_impl(h, c->q_dev, static_cast<int>(nq), k, D_dev, I_dev, 0, c->stream);
      if (rc == B2R_OK) {
        if (!zero_copy) e = cudaMemcpyAsync(c->out_pin, c->out_dev, out_bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      }
    }
    if (rc == B2R_OK && e != cudaSuccess) rc = fail(B2R_ERR_CUDA, std::string("b2r_flat_search_host: ") + cudaGetErrorString(e));
    if (rc == B2R_OK) {
      at = 0;
      for (const b2r_flat::HostReq* r : batch) {
        memcpy(r->D, c->out_pin + at * k * sizeof(float), static_cast<size_t>(r->nq) * k * sizeof(float));
        memcpy(r->I, c->out_pin + d_bytes + at * k * sizeof(int64_t), static_cast<size_t>(r->nq) * k * sizeof(int64_t));
        at += static_cast<size_t>(r->nq);
      }
    }
  }
  release_ctx(h, c);
  return rc;
}

int b2r_flat_search_host(b2r_flat* h, const float* q, int nq, int k, float* D, int64_t* I) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_search: NULL handle");
  if (nq < 0 || k < 1) return fail(B2R_ERR_INVALID, "b2r_flat_search: need nq >= 0 and k >= 1");
  if (k > B2R_MAX_K) return fail(B2R_ERR_UNSUPPORTED, "b2r_flat_search: k > B2R_MAX_K (1024)");
  if (nq == 0) return B2R_OK;
  if (!q || !D || !I) return fail(B2R_ERR_INVALID, "b2r_flat_search: NULL buffer");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_flat_search: cudaSetDevice failed");
  b2r_flat::HostReq me;
  me.q = q;
  me.nq = nq;
  me.k = k;
  me.D = D;
  me.I = I;
  if (tunable("flat.no_combine")) {
    std::vector<b2r_flat::HostReq*> one(1, &me);
    return flat_search_host_batch(h, one, k);
  }
  std::unique_lock<std::mutex> lk(h->bq_mu);
  h->bq.push_back(&me);
  // somebody is searching: wait until my request has been served, or until the leadership is free
  h->bq_cv.wait(lk, [&] { return me.done || !h->bq_leader; });
  if (me.done) {
    lk.unlock();
    return me.rc == B2R_OK ? B2R_OK : fail(me.rc, me.err);
  }
  h->bq_leader = true;
  int my_rc = B2R_OK;
  while (!me.done) {
    // everything queued with the front request's k (at most 1024 queries) runs as one batch
    std::vector<b2r_flat::HostReq*> batch;
    const int kk = h->bq.front()->k;
    int total = 0;
    for (auto it = h->bq.begin(); it != h->bq.end();) {
      if ((*it)->k == kk && (batch.empty() || total + (*it)->nq <= 1024)) {
        total += (*it)->nq;
        batch.push_back(*it);
        it = h->bq.erase(it);
      } else {
        ++it;
      }
    }
    lk.unlock();
    const int rc = flat_search_host_batch(h, batch, kk);
    const std::string err = rc == B2R_OK ? std::string() : last_error_ref();
    lk.lock();
    ++h->combined_batches;
    h->combined_requests += static_cast<long long>(batch.size());
    for (b2r_flat::HostReq* r : batch) {
      r->rc = rc;
      r->err = err;
      r->done = true;  // (my own request included: that ends the loop)
    }
    if (me.done) my_rc = me.rc;
    h->bq_cv.notify_all();
  }
  h->bq_leader = false;  // whoever is still queued takes over
  h->bq_cv.notify_all();
  lk.unlock();
  return my_rc == B2R_OK ? B2R_OK : fail(my_rc, me.err);
}

int b2r_flat_combine_stats(b2r_flat* h, int64_t* batches, int64_t* requests) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_combine_stats: NULL handle");
  std::lock_guard<std::mutex> lk(h->bq_mu);
  if (batches) *batches = h->combined_batches;
  if (requests) *requests = h->combined_requests;
  return B2R_OK;
}

int b2r_flat_reconstruct(b2r_flat* h, const int64_t* ids, int m, float* out) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_reconstruct: NULL handle");
  if (m < 0 || (m > 0 && (!ids || !out))) return fail(B2R_ERR_INVALID, "b2r_flat_reconstruct: bad arguments");
  if (m == 0) return B2R_OK;
  for (int i = 0; i < m; ++i) {
    if (ids[i] < 0 || ids[i] >= h->n) {
      return fail(B2R_ERR_INVALID, "b2r_flat_reconstruct: id " + std::to_string(ids[i]) + " out of range [0, " +
                                       std::to_string(h->n) + ")");
    }
  }
  DeviceGuard g(h->device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_flat_reconstruct: cudaSetDevice failed");
  HostCtx* c = nullptr;
  int rc = acquire_ctx(h, &c);
  if (rc) return rc;
  const size_t id_bytes = (static_cast<size_t>(m) * sizeof(int64_t) + 255) & ~static_cast<size_t>(255);
  const size_t row_bytes = static_cast<size_t>(m) * h->d * sizeof(float);
  rc = ensure_ctx_buffers(c, 1, id_bytes + row_bytes);
  if (rc == B2R_OK) {
    memcpy(c->out_pin, ids, static_cast<size_t>(m) * sizeof(int64_t));
    cudaError_t e = cudaMemcpyAsync(c->out_dev, c->out_pin, static_cast<size_t>(m) * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
      gather_rows_kernel<<<m, 256, 0, c->stream>>>(h->rows, h->d, reinterpret_cast<const long long*>(c->out_dev), m,
                                                   reinterpret_cast<float*>(c->out_dev + id_bytes));
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->out_pin + id_bytes, c->out_dev + id_bytes, row_bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail(B2R_ERR_CUDA, std::string("b2r_flat_reconstruct: ") + cudaGetErrorString(e));
    else memcpy(out, c->out_pin + id_bytes, row_bytes);
  }
  release_ctx(h, c);
  return rc;
}

int b2r_flat_set_candidate_bf16(b2r_flat* h, int enable) {
  if (!h) return fail(B2R_ERR_INVALID, "b2r_flat_set_candidate_bf16: NULL handle");
  DeviceGuard g(h->device);
  std::lock_guard<std::mutex> lock(h->mu);
  h->want_bf16 = enable != 0;
  if (!h->want_bf16 && h->rows_bf16) {
    cudaDeviceSynchronize();
    cudaFree(h->rows_bf16);
    h->rows_bf16 = nullptr;
    h->bf16_rows = -1;
  }
  return B2R_OK;
}

int b2r_flat_destroy(b2r_flat* h) {
  if (!h) return B2R_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  if (h->rows_bf16) cudaFree(h->rows_bf16);
  // host contexts and kernel workspaces belong to the device-wide pools and outlive the handle
  if (h->owns && h->rows) cudaFree(h->rows);
  if (h->max_norm2) cudaFree(h->max_norm2);
  delete h;
  return B2R_OK;
}

}  // extern "C"
