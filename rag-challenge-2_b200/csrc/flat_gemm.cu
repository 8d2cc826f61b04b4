// flat_gemm.cu — K2: batched exact inner-product top-k on the 5th-gen tensor cores (tcgen05).
//
// Replaces index.search(x[nq, d], k) for query batches (the multi-query fan-out of
// src/retrieval.py:1475-1504 and 127-question batches): every row tile is read from HBM once and
// scored against up to N = 64 queries by the kernels of this file (list-keeping select; the 3xTF32 redo / k > 64 path
// and the small-index path) and against up to 256 by the threshold candidate passes of flat_gemm_wide.cu.
//
//   S[128 rows, N queries] = A[128, d] * B[N, d]^T          per row tile, accumulated over d in TMEM
//
// Exactness.  The kernel has two modes (template TERMS):
//   TERMS = 1  one tf32 MMA per k-step on the raw fp32 data (the tensor core uses the top 19 bits):
//              scores are approximate, |err| <= 2^-9 |q||x|.  It selects k' > k candidates per
//              query; `gemm_rescore_kernel` then recomputes their scores in fp32 with the SIMT
//              kernel's arithmetic, takes the best k and CHECKS that no row outside the candidate
//              set can reach the k-th exact score (approx k'-th + bound < exact k-th).  Queries
//              that fail the check are redone by the TERMS = 3 pass.
//   TERMS = 3  3xTF32: x = hi + lo with hi = rna_tf32(x), lo = x - hi; issues A_hi*B_hi +
//              A_lo*B_hi + A_hi*B_lo (fp32 accumulate, separate TMEM accumulators for the large
//              and the small terms, 2-way split-K), dropping only lo*lo: ~3e-6 relative, the same
//              order as a different fp32 summation order.  Shared-memory bound (7x the HBM stream:
//              ncu r01), hence only the fallback and the N x N cosine path.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over row tiles t = cta, cta + grid, ...):
//   warp 0      TMA producer: per K-slab (32 floats = 128 B, SWIZZLE_128B) one 2-D tensor copy of
//               the A slab [128 x 32] and two of the query slabs B_hi / B_lo [N x 32] (L2 resident)
//   warp 1      allocates TMEM; one lane issues tcgen05.mma (M = 128, N, K = 8) x 4 k-steps x 3 terms
//               per slab, tcgen05.commit releases the shared-memory slots / publishes the tile
//   warps 2-5   transform: split the landed A slab into hi (in place) and lo (second ring),
//               fence.proxy.async, signal the MMA warp
//   warps 6-9   epilogue: tcgen05.ld the 128 x N accumulator tile (thread = row), fused select with
//               per-query threshold + candidate lists in shared memory (retry on overflow)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// A second tiny kernel (one CTA per query) merges the per-CTA lists.
//
// Rooflines: HBM for N up to ~100 (4*n*d bytes per launch, independent of N); tensor beyond.
#include <algorithm>

#include "tc_ptx.cuh"

namespace b2r {

static const int kGemmThreads = 320;
static const int kGemmMaxBatch = 256;  // queries per gemm_search_chunk call (CTA-pair candidate pass)

// csrc/flat_gemm_wide.cu
bool wide_path_supported(long long n, int d, int kp);
int wide_candidates(const float* rows, const void* rows_bf16, long long n, int d, const float* q_dev, int nq, int kp,
                    float* B, float* tauf, u32* premax, u32* gcnt, u32* ovf, u64* glists, size_t glists_bytes, float* Dp,
                    long long* Ip, int num_sms, cudaStream_t stream, int* launches);
bool wide_uses_bf16(const void* rows_bf16, int d, int nq);
static const int kLoStages = 4;

struct GemmArgs {
  long long n;
  int d;
  int nq;       // active queries (<= N)
  int k, cap;   // cap: per-query candidate capacity (power of two >= 2k)
  int raw_stages;
  long long n_tiles;
  u64* partial;        // [N][gridDim.x][k]
  const u32* active;   // optional per-query flags (device); NULL = all queries active
  int debug_skip;      // perf experiments only (results become wrong): 1 = re-use the tile's first query slab,
                       // 2 = issue no MMAs, 4 = drain TMEM without selecting
  // STATS epilogue (N x N cosine analysis): the "queries" are rows [col0, col0 + N) of the same matrix
  long long col0;
  double* stats;       // [0] sum, [1] sum of squares, [2] count   (pairs i < j, each once)
  u32* minmax;         // [0] min, [1] max as ordered float keys (atomicMin / atomicMax)
  u64* hist;           // [n_bins] counts over [-1, 1]
  int n_bins;
};


// ---------------------------------------------------------------------------------------------
// prep: split the queries into tf32 hi / lo parts: Bs[0][q][:] = hi, Bs[1][q][:] = lo; rows nq..N-1 zero
// ---------------------------------------------------------------------------------------------
__global__ void gemm_split_queries_kernel(const float* __restrict__ q, int nq, int N, int d, float* __restrict__ Bs,
                                          int split, const u32* __restrict__ active) {
  if (active) {  // redo pass: nothing to prepare unless some query is flagged
    const int lane = threadIdx.x & 31;
    bool any = false;
    for (int c = lane; c < nq; c += 32) any |= active[c] != 0u;
    if (!__any_sync(0xffffffffu, any)) return;
  }
  const long long total = static_cast<long long>(N) * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / d);
    float hi = 0.f, lo = 0.f;
    if (row < nq) {
      const float x = q[i];
      hi = split ? rna_tf32(x) : x;  // split == 0: raw queries for the single-pass mode
      lo = x - hi;
    }
    Bs[i] = hi;
    Bs[total + i] = lo;
  }
}

// ---------------------------------------------------------------------------------------------
// the scan
// ---------------------------------------------------------------------------------------------
template <int N, int TERMS, int SUB, bool STATS>
__global__ void __launch_bounds__(kGemmThreads, 1)
    flat_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const GemmArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // SWIZZLE_128B atoms must start on 1024-byte boundaries: align by hand (the host reserves the slack)
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // TERMS == 1 scores 256 rows (two M = 128 MMAs) per query slab, halving the L2 traffic of the
  // re-fetched queries; TERMS == 3 keeps 128 (its shared memory holds the lo ring instead)
  constexpr int kSub = SUB;                               // 128-row sub-tiles per tile (2 only with TERMS == 1)
  constexpr int kAcc = TERMS == 3 ? 4 : 2;                // TMEM accumulators per sub-tile
  constexpr u32 kSubBytes = kTileM * kSlabK * 4;          // 16 KiB: one sub-tile slab
  constexpr u32 kABytes = kSub * kSubBytes;
  constexpr u32 kBBytes = N * kSlabK * 4;                 // N * 128 B
  constexpr u32 kRawStage = kABytes + (TERMS == 3 ? 2 : 1) * kBBytes;  // A (raw -> hi) | B[_hi] | (B_lo)
  constexpr int kLo = TERMS == 3 ? kLoStages : 0;
  // The tensor core truncates when it accumulates in fp32, so the error of a long MMA chain grows
  // with its length and with the size of the running sum.  Per tile we therefore keep FOUR
  // accumulators of N columns: hi*hi split 2-way over K (even / odd slabs) and the two small
  // cross terms likewise; the epilogue adds them in fp32 (round-to-nearest).  x2 tiles in flight.
  constexpr u32 kTileCols = kSub * kAcc * N;              // TMEM columns of one tile (= 4N in both modes)
  constexpr u32 kTmemCols = 2 * kTileCols;                // 128 / 256 / 512 columns for N = 16 / 32 / 64
  const int R = a.raw_stages;
  const int n_slabs = a.d / kSlabK;

  unsigned char* raw_ring = smem_raw;                                  // R stages
  unsigned char* lo_ring = raw_ring + static_cast<size_t>(R) * kRawStage;  // kLoStages x 16 KiB
  u64* bars = reinterpret_cast<u64*>(lo_ring + kLo * kSubBytes);
  u64* full = bars;                    // [R]   TMA landed
  u64* empty = full + R;               // [R]   MMAs that read the raw stage are done
  u64* xf_full = empty + R;            // [kLoStages] hi/lo written
  u64* xf_empty = xf_full + kLoStages; // [kLoStages] MMAs that read the lo stage are done
  u64* tm_full = xf_empty + kLoStages; // [2] accumulator tile complete
  u64* tm_empty = tm_full + 2;         // [2] accumulator tile drained
  u64* tau_s = tm_empty + 2;           // [N]
  u32* cnt_s = reinterpret_cast<u32*>(tau_s + N);  // [N]
  u32* flags = cnt_s + N;              // [2] retry flags, [2] tmem base slot
  u32* tmem_slot = flags + 2;
  u64* lists = reinterpret_cast<u64*>(tmem_slot + 2);  // [N][cap]

  // queries this launch has to answer; none (all redo flags clear) -> leave before touching TMEM
  u64 active_mask = (a.nq >= 64) ? ~0ull : ((1ull << a.nq) - 1ull);
  if (a.active) {
    const u32 f0 = lane < a.nq ? a.active[lane] : 0u;
    const u32 f1 = lane + 32 < a.nq ? a.active[lane + 32] : 0u;
    const u64 m = static_cast<u64>(__ballot_sync(0xffffffffu, f0 != 0u)) |
                  (static_cast<u64>(__ballot_sync(0xffffffffu, f1 != 0u)) << 32);
    active_mask &= m;
    if (!active_mask) return;  // uniform over the CTA
  }
  if (tid == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kLoStages; ++s) {
      mbar_init(&xf_full[s], 4);   // one arrive per transform warp
      mbar_init(&xf_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tm_full[b], 1);
      mbar_init(&tm_empty[b], 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  for (int i = tid; i < N; i += blockDim.x) {
    tau_s[i] = 0;
    cnt_s[i] = 0;
  }
  if (tid < 2) flags[tid] = 0;
  if constexpr (STATS) {
    u32* hist_s = reinterpret_cast<u32*>(lists);  // the candidate-list area holds the CTA's histogram
    for (int i = tid; i < a.n_bins; i += blockDim.x) hist_s[i] = 0;
  }
  if (warp == 1) tc_alloc(tmem_slot, kTmemCols);
  const long long n_tiles = a.n_tiles;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      const u64 pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
      int s = 0;
      u32 ph = 0;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int row0 = static_cast<int>(t * (kTileM * kSub));
        for (int sl = 0; sl < n_slabs; ++sl) {
          mbar_wait_tag(&empty[s], ph ^ 1u, 100 + s);
          unsigned char* st = raw_ring + static_cast<size_t>(s) * kRawStage;
          const bool skip_b = (a.debug_skip & 1) && sl > 0;
          mbar_arrive_expect_tx(&full[s], skip_b ? kABytes : kRawStage);
#pragma unroll
          for (int sub = 0; sub < kSub; ++sub) {
            tma_load_2d(st + sub * kSubBytes, &map_a, sl * kSlabK, row0 + sub * kTileM, &full[s], pol_a);
          }
          if (!skip_b) tma_load_2d(st + kABytes, &map_b, sl * kSlabK, 0, &full[s], pol_b);
          if constexpr (TERMS == 3) if (!skip_b) tma_load_2d(st + kABytes + kBBytes, &map_b, sl * kSlabK, N, &full[s], pol_b);
          if (++s == R) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr u32 idesc = make_idesc_tf32(N);
      int s = 0, l = 0, buf = 0;
      u32 ph_s = 0, ph_l = 0, ph_buf[2] = {0, 0};
      (void)l;
      (void)ph_l;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait_tag(&tm_empty[buf], ph_buf[buf] ^ 1u, 200 + buf);  // epilogue drained this accumulator tile
        tc_fence_after();
        const u32 tile_d = tmem_base + static_cast<u32>(buf) * kTileCols;
        for (int sl = 0; sl < n_slabs; ++sl) {
          const u32 d_main = tile_d + static_cast<u32>((sl & 1) * N);
          const u32 d_small = tile_d + static_cast<u32>((2 + (sl & 1)) * N);  // TERMS == 3 only
          (void)d_small;
          const bool first = sl < 2;  // first slab of this K-split: overwrite instead of accumulate
          if constexpr (TERMS == 3) {
            mbar_wait_tag(&xf_full[l], ph_l, 300 + l);
          } else {
            mbar_wait_tag(&full[s], ph_s, 400 + s);
          }
          tc_fence_after();
          const u32 st = smem_u32(raw_ring + static_cast<size_t>(s) * kRawStage);
          const u64 d_ahi = make_kmajor_sw128_desc(st);
          const u64 d_bhi = make_kmajor_sw128_desc(st + kABytes);
          if constexpr (TERMS == 3) {
            const u64 d_blo = make_kmajor_sw128_desc(st + kABytes + kBBytes);
            const u64 d_alo = make_kmajor_sw128_desc(smem_u32(lo_ring + static_cast<size_t>(l) * kSubBytes));
#pragma unroll
            for (int kk = 0; kk < kSlabK / 8; ++kk) {
              const u64 adv = static_cast<u64>(kk * 8 * 4 >> 4);  // 32 bytes per k-step inside the swizzle row
              const u32 acc = (first && kk == 0) ? 0u : 1u;
              tc_mma_tf32(d_main, d_ahi + adv, d_bhi + adv, idesc, acc);
              tc_mma_tf32(d_small, d_alo + adv, d_bhi + adv, idesc, acc);
              tc_mma_tf32(d_small, d_ahi + adv, d_blo + adv, idesc, 1u);
            }
            tc_commit(&empty[s]);      // raw stage reusable once these MMAs have read it
            tc_commit(&xf_empty[l]);   // lo stage likewise
            if (++l == kLoStages) {
              l = 0;
              ph_l ^= 1u;
            }
          } else {
#pragma unroll
            for (int sub = 0; sub < kSub; ++sub) {
              const u64 d_a = make_kmajor_sw128_desc(st + sub * kSubBytes);
              const u32 d_acc = d_main + static_cast<u32>(sub * kAcc * N);
#pragma unroll
              for (int kk = 0; kk < kSlabK / 8; ++kk) {
                const u64 adv = static_cast<u64>(kk * 8 * 4 >> 4);
                if (!(a.debug_skip & 2)) tc_mma_tf32(d_acc, d_a + adv, d_bhi + adv, idesc, (first && kk == 0) ? 0u : 1u);
              }
            }
            tc_commit(&empty[s]);
          }
          if (++s == R) {
            s = 0;
            ph_s ^= 1u;
          }
        }
        tc_commit(&tm_full[buf]);    // accumulator tile complete
        ph_buf[buf] ^= 1u;
        buf ^= 1;
      }
    }
  } else if (warp < 6) {
    // ======================= transform warps: A -> hi (in place), lo =======================
    if constexpr (TERMS == 3) {
    const int xt = tid - 64;  // 0..127
    int s = 0, l = 0;
    u32 ph_s = 0, ph_l = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int sl = 0; sl < n_slabs; ++sl) {
        mbar_wait_tag(&full[s], ph_s, 500 + s);
        mbar_wait_tag(&xf_empty[l], ph_l ^ 1u, 600 + l);
        float4* hi = reinterpret_cast<float4*>(raw_ring + static_cast<size_t>(s) * kRawStage);
        float4* lo = reinterpret_cast<float4*>(lo_ring + static_cast<size_t>(l) * kSubBytes);
#pragma unroll
        for (int i = 0; i < static_cast<int>(kSubBytes / 16 / 128); ++i) {
          const int idx = xt + i * 128;
          const float4 x = hi[idx];
          float4 h, r;
          h.x = rna_tf32(x.x);
          h.y = rna_tf32(x.y);
          h.z = rna_tf32(x.z);
          h.w = rna_tf32(x.w);
          // the remainder is exact in fp32 (<= 13 significant bits); the tensor core drops at most
          // its 2 lowest bits (2^-23 relative to x), so it is stored un-rounded
          r.x = x.x - h.x;
          r.y = x.y - h.y;
          r.z = x.z - h.z;
          r.w = x.w - h.w;
          hi[idx] = h;
          lo[idx] = r;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_full[l]);
        if (++s == R) {
          s = 0;
          ph_s ^= 1u;
        }
        if (++l == kLoStages) {
          l = 0;
          ph_l ^= 1u;
        }
      }
    }
    }
  } else {
    // ======================= epilogue warps: TMEM -> registers -> fused select =======================
    const int ew = warp - 6;       // 0..3
    const int quad = warp & 3;     // TMEM lane quadrant this warp may access
    const int cap = a.cap, k = a.k;
    int buf = 0, it = 0;
    u32 ph_buf[2] = {0, 0};
    double st_sum = 0.0, st_sq = 0.0;
    unsigned long long st_cnt = 0;
    float st_min = 3.4028235e38f, st_max = -3.4028235e38f;
    (void)st_sum; (void)st_sq; (void)st_cnt; (void)st_min; (void)st_max;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait_tag(&tm_full[buf], ph_buf[buf], 700 + buf);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < kSub; ++sub) {
      float sc[N];
      const u32 taddr = tmem_base + (static_cast<u32>(quad * 32) << 16) + static_cast<u32>(buf) * kTileCols +
                        static_cast<u32>(sub * kAcc * N);
#pragma unroll
      for (int c = 0; c < N; c += 16) {
        float m0[16], m1[16];
        tc_ld16(taddr + c, m0);
        if (n_slabs > 1) tc_ld16(taddr + N + c, m1);
        if constexpr (TERMS == 3) {
          float s0[16], s1[16];
          tc_ld16(taddr + 2 * N + c, s0);
          if (n_slabs > 1) tc_ld16(taddr + 3 * N + c, s1);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            sc[c + i] = (n_slabs > 1) ? (m0[i] + m1[i]) + (s0[i] + s1[i]) : m0[i] + s0[i];
          }
        } else {
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) sc[c + i] = (n_slabs > 1) ? m0[i] + m1[i] : m0[i];
        }
      }
      if (sub == kSub - 1) {  // last sub-tile is in registers: hand the accumulator tile back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tm_empty[buf]);
      }

      const long long row = t * (kTileM * kSub) + sub * kTileM + quad * 32 + lane;
      if constexpr (STATS) {
        // cosine statistics over the pairs (row, col) with row < col; S = clip(Xn Xn^T, -1, 1)
        // (generate_similarity_matrix.py:90-98); the diagonal is excluded (:162-163)
        u32* hist_s = reinterpret_cast<u32*>(lists);
        const float scale = 0.5f * static_cast<float>(a.n_bins);
#pragma unroll
        for (int c = 0; c < N; ++c) {
          const long long col = a.col0 + c;
          if (row < a.n && col < a.n && row < col) {
            const float v = fminf(1.0f, fmaxf(-1.0f, sc[c]));
            st_sum += static_cast<double>(v);
            st_sq += static_cast<double>(v) * static_cast<double>(v);
            st_min = fminf(st_min, v);
            st_max = fmaxf(st_max, v);
            ++st_cnt;
            int bin = static_cast<int>((v + 1.0f) * scale);
            bin = bin >= a.n_bins ? a.n_bins - 1 : bin;
            atomicAdd(&hist_s[bin], 1u);
          }
        }
        continue;
      }
      u64 pend = (row < a.n && !(a.debug_skip & 4)) ? active_mask : 0ull;
      for (;;) {
        u32* flag = flags + (it & 1);
#pragma unroll
        for (int c = 0; c < N; ++c) {
          if ((pend >> c) & 1ull) {
            const u64 key = make_key(sc[c], static_cast<u32>(row));
            if (key > tau_s[c]) {
              const u32 pos = atomicAdd(&cnt_s[c], 1u);
              if (pos < static_cast<u32>(cap)) {
                lists[static_cast<size_t>(c) * cap + pos] = key;
                pend &= ~(1ull << c);
              }
            } else {
              pend &= ~(1ull << c);
            }
          }
        }
        if (pend) *flag = 1u;
        named_bar_sync(1, 128);
        // compaction: a list filled past the midpoint between k and its capacity (or overflowed) is
        // cut back to its best k, which also raises the query's threshold
        for (int c = ew; c < a.nq; c += 4) {
          const u32 nc = cnt_s[c];
          if (nc >= static_cast<u32>((k + cap + 1) / 2)) {
            u64* list = lists + static_cast<size_t>(c) * cap;
            const int valid = nc < static_cast<u32>(cap) ? static_cast<int>(nc) : cap;
            for (int i = valid + lane; i < cap; i += 32) list[i] = 0;
            __syncwarp();
            warp_sort_desc(list, cap, lane);
            if (lane == 0) {
              const int keep = valid < k ? valid : k;
              cnt_s[c] = static_cast<u32>(keep);
              if (valid >= k) tau_s[c] = list[k - 1];
            }
            __syncwarp();
          }
        }
        const u32 again = *flag;
        if (tid == 6 * 32) flags[(it + 1) & 1] = 0;
        named_bar_sync(1, 128);
        ++it;
        if (!again) break;
      }
      }  // sub-tiles
      ph_buf[buf] ^= 1u;
      buf ^= 1;
    }
    // per-CTA result: best k of every query's list -> global scratch
    named_bar_sync(1, 128);
    if constexpr (STATS) {
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        st_sum += __shfl_xor_sync(0xffffffffu, st_sum, m);
        st_sq += __shfl_xor_sync(0xffffffffu, st_sq, m);
        st_cnt += __shfl_xor_sync(0xffffffffu, st_cnt, m);
        st_min = fminf(st_min, __shfl_xor_sync(0xffffffffu, st_min, m));
        st_max = fmaxf(st_max, __shfl_xor_sync(0xffffffffu, st_max, m));
      }
      if (lane == 0 && st_cnt) {
        atomicAdd(a.stats + 0, st_sum);
        atomicAdd(a.stats + 1, st_sq);
        atomicAdd(a.stats + 2, static_cast<double>(st_cnt));
        atomicMin(a.minmax + 0, float_to_ordered(st_min));
        atomicMax(a.minmax + 1, float_to_ordered(st_max));
      }
      const u32* hist_s = reinterpret_cast<const u32*>(lists);
      for (int b = tid - 6 * 32; b < a.n_bins; b += 128) {
        const u32 h = hist_s[b];
        if (h) atomicAdd(a.hist + b, static_cast<u64>(h));
      }
    }
    for (int c = ew; !STATS && c < a.nq; c += 4) {
      if (!((active_mask >> c) & 1ull)) continue;  // not this launch's query: its lists are never read
      u64* list = lists + static_cast<size_t>(c) * cap;
      const u32 nc = cnt_s[c];
      const int valid = nc < static_cast<u32>(cap) ? static_cast<int>(nc) : cap;
      for (int i = valid + lane; i < cap; i += 32) list[i] = 0;
      __syncwarp();
      warp_sort_desc(list, cap, lane);
      u64* dst = a.partial + (static_cast<size_t>(c) * gridDim.x + blockIdx.x) * k;
      for (int i = lane; i < k; i += 32) dst[i] = i < cap ? list[i] : 0ull;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, kTmemCols);
  }
}

// one CTA per query: merge the per-CTA lists of the scan into that query's top-k
__global__ void __launch_bounds__(256) select_partials_kernel(const u64* __restrict__ partial, int n_lists, int k,
                                                              float* __restrict__ D, long long* __restrict__ I,
                                                              long long id_offset, const u32* __restrict__ active) {
  __shared__ u64 buf[4096];
  __shared__ u64 ctl[4];
  const int q = blockIdx.x;
  if (active && !active[q]) return;  // answered by an earlier pass
  block_select_lists(partial + static_cast<size_t>(q) * n_lists * k, n_lists, k, k, buf, 4096, ctl);
  block_emit_keys(buf, k, D + static_cast<size_t>(q) * k, I + static_cast<size_t>(q) * k, id_offset);
}

// max over rows of |x|^2 (one pass over the shard, cached per index): the error bound of the
// single-pass tf32 scores is 2^-9 * |q| * max|x|
__global__ void max_row_norm2_kernel(const float* __restrict__ xb, long long n, int d, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  float best = 0.f;
  for (long long r = gw; r < n; r += nw) {
    const float* row = xb + r * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(row[c], row[c], acc);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    best = fmaxf(best, acc);
  }
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(best));  // non-negative floats order as ints
}

// Exact fp32 rescoring of the k' candidates of every query, final top-k, and the exactness check.
// One CTA per query; a warp scores one candidate at a time with the SIMT scan's arithmetic
// (lanes over 128-bit chunks, even / odd accumulators, xor butterfly).
__global__ void __launch_bounds__(512) gemm_rescore_kernel(const float* __restrict__ xb, int d,
                                                           const float* __restrict__ q, const float* __restrict__ Dp,
                                                           const long long* __restrict__ Ip, int kp, int k,
                                                           const float* __restrict__ max_norm2, float* __restrict__ D,
                                                           long long* __restrict__ I, long long id_offset,
                                                           u32* __restrict__ redo, const u32* __restrict__ ovf,
                                                           float operand_err) {
  __shared__ u64 keys[256];
  __shared__ float s_qn2;
  const int qi = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* qv = q + static_cast<size_t>(qi) * d;
  const long long* cand = Ip + static_cast<size_t>(qi) * kp;
  if (tid < 256) keys[tid] = 0;
  if (warp == 0) {
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(qv[c], qv[c], acc);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (lane == 0) s_qn2 = acc;
  }
  __syncthreads();
  const bool vec4 = (d & 3) == 0;
  const int n_warps = blockDim.x >> 5;
  for (int c = warp; c < kp; c += n_warps) {
    const long long id = cand[c];
    if (id < 0) continue;  // warp-uniform
    const float* row = xb + id * d;
    float a0 = 0.f, a1 = 0.f;
    if (vec4) {
      // chunk j = ch / 32 goes to accumulator j & 1 (the SIMT scan's order); four loads in flight per lane
      const int n4 = d >> 2;
      const float4* row4 = reinterpret_cast<const float4*>(row);
      const float4* q4 = reinterpret_cast<const float4*>(qv);
      for (int base = 0; base < n4; base += 128) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ch = base + u * 32 + lane;
          x[u] = ch < n4 ? __ldg(row4 + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ch = base + u * 32 + lane;
          if (ch < n4) {
            const float4 y = q4[ch];
            float& acc = (u & 1) ? a1 : a0;
            acc = fmaf(x[u].x, y.x, acc);
            acc = fmaf(x[u].y, y.y, acc);
            acc = fmaf(x[u].z, y.z, acc);
            acc = fmaf(x[u].w, y.w, acc);
          }
        }
      }
    } else {
      for (int ch = lane; ch < d; ch += 32) a0 = fmaf(__ldg(row + ch), qv[ch], a0);
    }
    float sum = a0 + a1;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    if (lane == 0) keys[c] = make_key(sum, static_cast<u32>(id));
  }
  __syncthreads();
  block_sort_desc(keys, 256);
  block_emit_keys(keys, k, D + static_cast<size_t>(qi) * k, I + static_cast<size_t>(qi) * k, id_offset);
  if (tid == 0) {
    // Rows outside the candidate set have approx score <= the k'-th approx score, hence exact
    // score <= that + bound.  If this cannot reach the exact k-th score the answer is proven exact.
    u32 need_redo = 0;
    const bool list_full = cand[kp - 1] >= 0;  // fewer than k' rows in the shard: every row was rescored
    if (list_full) {
      // operand error (tf32 truncation 2^-9, bf16 rounding 2^-8; +6 %) plus the truncating fp32 accumulation of a
      // length-d chain, d * 2^-24
      const float rel = operand_err * 1.0625f + static_cast<float>(d) * 5.9604645e-8f;
      const float bound = rel * sqrtf(s_qn2) * sqrtf(*max_norm2) + 1e-6f;
      const u64 kth = keys[k - 1];
      const float exact_kth = kth ? key_score(kth) : -3.4028235e38f;
      if (!(Dp[static_cast<size_t>(qi) * kp + kp - 1] + bound < exact_kth)) need_redo = 1;
    }
    if (ovf && ovf[qi]) need_redo = 1;  // a candidate list overflowed: the set may be incomplete
    redo[qi] = need_redo;
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  }
  return fn;
}

int make_map_2d(CUtensorMap* map, const float* base, unsigned long long inner, unsigned long long outer,
                       unsigned box_inner, unsigned box_outer) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  if (!enc) return fail(B2R_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2R_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return B2R_OK;
}

int make_map_2d_bf16(CUtensorMap* map, const void* base, unsigned long long inner, unsigned long long outer,
                     unsigned box_inner, unsigned box_outer) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  if (!enc) return fail(B2R_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2R_ERR_CUDA, "cuTensorMapEncodeTiled (bf16) failed with code " + std::to_string(r));
  return B2R_OK;
}

template <int N, int TERMS, int SUB, bool STATS = false>
static int gemm_launch_nt(const CUtensorMap& ma, const CUtensorMap& mb, const GemmArgs& a, int grid, size_t smem,
                          cudaStream_t stream) {
  auto kern = flat_gemm_tc_kernel<N, TERMS, SUB, STATS>;
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<grid, kGemmThreads, smem, stream>>>(ma, mb, a);
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

static int pow2_ge_g(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// one scan pass: `terms` = 1 (approximate, raw tf32) or 3 (3xTF32); selects `ksel` per query
static int gemm_scan(const float* rows, long long n, int d, const float* q_dev, int nq, int ksel, int terms,
                     const u32* active, float* D_out, long long* I_out, long long id_offset, float* bsplit,
                     u64* partial, int num_sms, cudaStream_t stream) {
  const int cap = pow2_ge_g(std::max(2 * ksel, 64));
  // candidate lists (N x cap keys) get at most 64 KiB of shared memory
  int N = nq <= 16 ? 16 : (nq <= 32 ? 32 : 64);
  if (static_cast<size_t>(N) * cap * 8 > 64 * 1024) return fail(B2R_ERR_INVALID, "gemm_scan: batch too wide for this k");
  const int sub = (terms == 1 && tunable("gemm.tile256") && n >= 256) ? 2 : 1;
  const int tile_rows = sub * kTileM;
  const long long n_tiles = (n + tile_rows - 1) / tile_rows;
  const int grid = static_cast<int>(std::min<long long>(n_tiles, num_sms));
  gemm_split_queries_kernel<<<64, 256, 0, stream>>>(q_dev, nq, N, d, bsplit, terms == 3 ? 1 : 0, active);
  B2R_CUDA(cudaGetLastError());
  CUtensorMap ma, mb;
  int rc = make_map_2d(&ma, rows, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kTileM);
  if (rc) return rc;
  rc = make_map_2d(&mb, bsplit, static_cast<unsigned long long>(d), static_cast<unsigned long long>(2 * N), kSlabK, N);
  if (rc) return rc;
  const size_t raw_stage = static_cast<size_t>(tile_rows) * kSlabK * 4 + static_cast<size_t>(terms == 3 ? 2 : 1) * N * kSlabK * 4;
  const size_t fixed = (terms == 3 ? kLoStages : 0) * static_cast<size_t>(kTileM) * kSlabK * 4 + 64 * 8 /*barriers*/ +
                       static_cast<size_t>(N) * 12 + 32 + static_cast<size_t>(N) * cap * 8 + 1024 /*alignment slack*/;
  int stages = tunable("gemm.stages");
  if (stages <= 0) stages = 8;
  const size_t limit = 227 * 1024 - 1024;
  while (stages > 2 && stages * raw_stage + fixed > limit) --stages;
  if (stages * raw_stage + fixed > limit) return fail(B2R_ERR_UNSUPPORTED, "tensor-core scan: k too large for shared memory");
  GemmArgs a;
  a.n = n;
  a.d = d;
  a.nq = nq;
  a.k = ksel;
  a.cap = cap;
  a.raw_stages = stages;
  a.n_tiles = n_tiles;
  a.partial = partial;
  a.active = active;
  a.debug_skip = tunable("gemm.debug_skip");
  a.col0 = 0;
  a.stats = nullptr;
  a.minmax = nullptr;
  a.hist = nullptr;
  a.n_bins = 0;
  const size_t smem = stages * raw_stage + fixed;
  if (terms == 3) {
    switch (N) {
      case 16: rc = gemm_launch_nt<16, 3, 1>(ma, mb, a, grid, smem, stream); break;
      case 32: rc = gemm_launch_nt<32, 3, 1>(ma, mb, a, grid, smem, stream); break;
      default: rc = gemm_launch_nt<64, 3, 1>(ma, mb, a, grid, smem, stream); break;
    }
  } else if (sub == 2) {
    switch (N) {
      case 16: rc = gemm_launch_nt<16, 1, 2>(ma, mb, a, grid, smem, stream); break;
      case 32: rc = gemm_launch_nt<32, 1, 2>(ma, mb, a, grid, smem, stream); break;
      default: rc = gemm_launch_nt<64, 1, 2>(ma, mb, a, grid, smem, stream); break;
    }
  } else {
    switch (N) {
      case 16: rc = gemm_launch_nt<16, 1, 1>(ma, mb, a, grid, smem, stream); break;
      case 32: rc = gemm_launch_nt<32, 1, 1>(ma, mb, a, grid, smem, stream); break;
      default: rc = gemm_launch_nt<64, 1, 1>(ma, mb, a, grid, smem, stream); break;
    }
  }
  if (rc) return rc;
  select_partials_kernel<<<nq, 256, 0, stream>>>(partial, grid, ksel, D_out, I_out, id_offset, active);
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

bool gemm_path_supported(int d, int k) { return d % kSlabK == 0 && d >= kSlabK && d <= 16384 && k <= 128; }

// largest batch one pass of the list-keeping scan can take for this k (candidate lists are capped at
// 64 KiB of shared memory)
static int list_scan_max_batch(int ksel) {
  const int cap = pow2_ge_g(std::max(2 * ksel, 64));
  int n = 64;
  while (n > 16 && static_cast<size_t>(n) * cap * 8 > 64 * 1024) n >>= 1;
  return n;
}

// largest batch gemm_search_chunk accepts
int gemm_max_batch(int k) {
  (void)k;
  return tunable("gemm.no_pair") ? 128 : kGemmMaxBatch;
}

int gemm_max_row_norm2(const float* rows, long long n, int d, float* out_dev, cudaStream_t stream) {
  B2R_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(float), stream));
  max_row_norm2_kernel<<<148 * 8, 256, 0, stream>>>(rows, n, d, out_dev);
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

// device scratch of one search stream, carved from one allocation
static const size_t kScrB = static_cast<size_t>(kGemmMaxBatch) * 16384 * sizeof(float);  // query operand(s): 8 MiB
static const size_t kScrDp = kScrB;                                                   // float [128][128]
static const size_t kScrIp = kScrDp + static_cast<size_t>(kGemmMaxBatch) * 128 * 4;   // int64 [128][128]
static const size_t kScrFlags = kScrIp + static_cast<size_t>(kGemmMaxBatch) * 128 * 8;  // redo, ovf, tauf [128] each
static const size_t kScrGcnt = kScrFlags + 4096;                                      // u32 [<= 256 CTAs][128]
static const size_t kScrPremax = kScrGcnt + static_cast<size_t>(256) * kGemmMaxBatch * 4;  // u32 [128][8192]
static const size_t kScrEnd = kScrPremax + static_cast<size_t>(kGemmMaxBatch) * 8192 * 4;

size_t gemm_scratch_bytes() { return kScrEnd + 1024; }

// Scores `nq` (<= kGemmMaxBatch) queries against the whole shard.  scratch: gemm_scratch_bytes() of device
// memory; partial: list scratch of partial_bytes (>= 64 * grid * 128 keys).
int gemm_search_chunk(const float* rows, const void* rows_bf16, long long n, int d, const float* q_dev, int nq, int k,
                      float* D_dev, long long* I_dev, long long id_offset, unsigned char* scratch, u64* partial,
                      size_t partial_bytes, const float* max_norm2_dev, int num_sms, cudaStream_t stream, int* launches) {
  float* bsplit = reinterpret_cast<float*>(scratch);
  float* Dp = reinterpret_cast<float*>(scratch + kScrDp);
  long long* Ip = reinterpret_cast<long long*>(scratch + kScrIp);
  u32* redo = reinterpret_cast<u32*>(scratch + kScrFlags);
  u32* ovf = redo + kGemmMaxBatch;
  float* tauf = reinterpret_cast<float*>(ovf + kGemmMaxBatch);
  u32* gcnt = reinterpret_cast<u32*>(scratch + kScrGcnt);
  u32* premax = reinterpret_cast<u32*>(scratch + kScrPremax);
  const int mode = tunable("gemm.mode");  // 0/1: approximate pass + exact rescoring (+ verified fallback); 3: 3xTF32 only
  int rc;
  // Report-sized indexes (below the threshold pass's 16,384 rows) are latency-bound: the 3xTF32 scan alone is three
  // launches, the candidate + rescoring + redo chain nine, and the 3xTF32 kernel's shared-memory appetite does not
  // matter on a few MB of rows (profiles/r01_small_index_sweep.json).
  const bool small_index = n < 16384 && !tunable("gemm.small_rescore");
  if (mode == 3 || k > 64 || small_index) {
    const int step = list_scan_max_batch(k);
    for (int off = 0; off < nq; off += step) {
      const int c = std::min(step, nq - off);
      rc = gemm_scan(rows, n, d, q_dev + static_cast<size_t>(off) * d, c, k, 3, nullptr, D_dev + static_cast<size_t>(off) * k,
                     I_dev + static_cast<size_t>(off) * k, id_offset, bsplit, partial, num_sms, stream);
      if (rc) return rc;
      *launches += 3;
    }
    return B2R_OK;
  }
  int kp = std::min(128, std::max(64, 2 * k));
  const bool wide = wide_path_supported(n, d, kp);
  // bf16 candidate pass (opt-in copy of the rows, batches of more than 128 queries): twice the operand error, so twice
  // the candidates for the completeness proof
  const bool bf16 = wide && wide_uses_bf16(rows_bf16, d, nq);
  if (bf16) kp = 128;
  const float operand_err = bf16 ? 0.00390625f : 0.001953125f;
  if (wide) {
    rc = wide_candidates(rows, bf16 ? rows_bf16 : nullptr, n, d, q_dev, nq, kp, bsplit, tauf, premax, gcnt, ovf, partial,
                         partial_bytes, Dp, Ip, num_sms, stream, launches);
    if (rc) return rc;
  } else {
    const int step = list_scan_max_batch(kp);
    for (int off = 0; off < nq; off += step) {
      const int c = std::min(step, nq - off);
      rc = gemm_scan(rows, n, d, q_dev + static_cast<size_t>(off) * d, c, kp, 1, nullptr, Dp + static_cast<size_t>(off) * kp,
                     Ip + static_cast<size_t>(off) * kp, 0, bsplit, partial, num_sms, stream);
      if (rc) return rc;
      *launches += 3;
    }
  }
  gemm_rescore_kernel<<<nq, 512, 0, stream>>>(rows, d, q_dev, Dp, Ip, kp, k, max_norm2_dev, D_dev, I_dev, id_offset, redo,
                                              wide ? ovf : nullptr, operand_err);
  B2R_CUDA(cudaGetLastError());
  *launches += 1;
  if (!tunable("gemm.no_verify")) {
    // queries whose candidate set could not be proven complete (or whose lists overflowed) are redone with
    // 3xTF32 accuracy by the list-keeping scan; its kernels exit at once when no flag is set
    const int step = list_scan_max_batch(k);
    for (int off = 0; off < nq; off += step) {
      const int c = std::min(step, nq - off);
      rc = gemm_scan(rows, n, d, q_dev + static_cast<size_t>(off) * d, c, k, 3, redo + off, D_dev + static_cast<size_t>(off) * k,
                     I_dev + static_cast<size_t>(off) * k, id_offset, bsplit, partial, num_sms, stream);
      if (rc) return rc;
      *launches += 3;
    }
  }
  return B2R_OK;
}

// ---------------------------------------------------------------------------------------------
// K5: N x N cosine analysis (semantic_similarity_analysis/generate_similarity_matrix.py:84-98,162-171)
// ---------------------------------------------------------------------------------------------
// Xn = X / (|X| + 1e-8), one warp per row
__global__ void cosine_normalize_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ xn) {
  const int lane = threadIdx.x & 31;
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = gw; r < n; r += nw) {
    const float* row = x + r * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(row[c], row[c], acc);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    const float inv = 1.0f / (sqrtf(acc) + 1e-8f);
    for (int c = lane; c < d; c += 32) xn[r * d + c] = row[c] * inv;
  }
}

__global__ void cosine_finalize_kernel(const double* acc, const u32* minmax, double* stats_out) {
  // pairs were visited once (i < j); the matrix is symmetric, so sums and counts double
  stats_out[0] = 2.0 * acc[0];
  stats_out[1] = 2.0 * acc[1];
  stats_out[2] = acc[2] > 0 ? static_cast<double>(ordered_to_float(minmax[0])) : 0.0;
  stats_out[3] = acc[2] > 0 ? static_cast<double>(ordered_to_float(minmax[1])) : 0.0;
  stats_out[4] = 2.0 * acc[2];
  stats_out[5] = 0.0;
}

__global__ void cosine_double_hist_kernel(u64* hist, int n_bins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bins) hist[i] *= 2ull;
}

// csrc/cosine_pair.cu
bool cosine_pair_supported(long long n, int d, int n_bins, int num_sms);
int cosine_stats_run_pair(const float* rows_dev, long long n, int d, long long col_begin, long long col_end,
                          double* stats_dev, u64* hist_dev, int n_bins, int num_sms, cudaStream_t stream);

int cosine_stats_run(const float* rows_dev, long long n, int d, long long col_begin, long long col_end,
                     double* stats_dev, u64* hist_dev, int n_bins, int num_sms, cudaStream_t stream) {
  if (cosine_pair_supported(n, d, n_bins, num_sms)) {
    return cosine_stats_run_pair(rows_dev, n, d, col_begin, col_end, stats_dev, hist_dev, n_bins, num_sms, stream);
  }
  constexpr int N = 64;
  float* xn = nullptr;
  unsigned char* scratch = nullptr;
  const size_t xn_bytes = static_cast<size_t>(n) * d * sizeof(float);
  const size_t bs_bytes = static_cast<size_t>(2) * N * d * sizeof(float);
  B2R_CUDA(cudaMalloc(&xn, xn_bytes));
  cudaError_t e = cudaMalloc(&scratch, bs_bytes + 256);
  if (e != cudaSuccess) {
    cudaFree(xn);
    return fail(B2R_ERR_NOMEM, "cosine: scratch allocation failed");
  }
  float* bsplit = reinterpret_cast<float*>(scratch);
  double* acc = reinterpret_cast<double*>(scratch + bs_bytes);       // sum, sumsq, count
  u32* minmax = reinterpret_cast<u32*>(scratch + bs_bytes + 64);
  int rc = B2R_OK;
  do {
    const u32 mm_init[2] = {0xFFFFFFFFu, 0u};
    if (cudaMemsetAsync(acc, 0, 64, stream) != cudaSuccess ||
        cudaMemcpyAsync(minmax, mm_init, sizeof(mm_init), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaMemsetAsync(hist_dev, 0, static_cast<size_t>(n_bins) * sizeof(u64), stream) != cudaSuccess) {
      rc = fail(B2R_ERR_CUDA, "cosine: memset failed");
      break;
    }
    cosine_normalize_kernel<<<num_sms * 8, 256, 0, stream>>>(rows_dev, n, d, xn);
    CUtensorMap ma, mb;
    rc = make_map_2d(&ma, xn, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kTileM);
    if (rc) break;
    rc = make_map_2d(&mb, bsplit, static_cast<unsigned long long>(d), static_cast<unsigned long long>(2 * N), kSlabK, N);
    if (rc) break;
    const int cap = 64;  // N * cap * 8 = 32 KiB: room for 8192 histogram bins
    const size_t raw_stage = static_cast<size_t>(kTileM) * kSlabK * 4 + static_cast<size_t>(2) * N * kSlabK * 4;
    const size_t fixed = kLoStages * static_cast<size_t>(kTileM) * kSlabK * 4 + 64 * 8 + static_cast<size_t>(N) * 12 + 32 +
                         static_cast<size_t>(N) * cap * 8 + 1024;
    int stages = 8;
    while (stages > 2 && stages * raw_stage + fixed > 227 * 1024 - 1024) --stages;
    const size_t smem = stages * raw_stage + fixed;
    for (long long c0 = (col_begin / N) * N; c0 < col_end && rc == B2R_OK; c0 += N) {
      const int ncols = static_cast<int>(std::min<long long>(N, n - c0));
      gemm_split_queries_kernel<<<64, 256, 0, stream>>>(xn + c0 * d, ncols, N, d, bsplit, 1, nullptr);
      GemmArgs a;
      a.n = n;
      a.d = d;
      a.nq = N;
      a.k = 1;
      a.cap = cap;
      a.raw_stages = stages;
      // only row tiles that contain some row i < max column of this block
      const long long last_col = std::min<long long>(n, c0 + N) - 1;
      a.n_tiles = last_col <= 0 ? 0 : (last_col - 1) / kTileM + 1;
      a.partial = nullptr;
      a.active = nullptr;
      a.debug_skip = 0;
      a.col0 = c0;
      a.stats = acc;
      a.minmax = minmax;
      a.hist = hist_dev;
      a.n_bins = n_bins;
      if (a.n_tiles == 0) continue;
      const int grid = static_cast<int>(std::min<long long>(a.n_tiles, num_sms));
      rc = gemm_launch_nt<N, 3, 1, true>(ma, mb, a, grid, smem, stream);
    }
    if (rc) break;
    cosine_finalize_kernel<<<1, 1, 0, stream>>>(acc, minmax, stats_dev);
    cosine_double_hist_kernel<<<(n_bins + 255) / 256, 256, 0, stream>>>(hist_dev, n_bins);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2R_ERR_CUDA, "cosine: launch failed");
  } while (0);
  cudaStreamSynchronize(stream);  // scratch is freed below
  cudaFree(xn);
  cudaFree(scratch);
  return rc;
}

}  // namespace b2r

extern "C" int b2r_cosine_stats_dev(const float* rows_dev, int64_t n, int d, int64_t row_begin, int64_t row_end,
                                    double* stats_dev, unsigned long long* hist_dev, int n_bins, int k, float* topk_S_dev,
                                    int64_t* topk_I_dev, int device, void* stream) {
  using namespace b2r;
  (void)topk_S_dev;
  (void)topk_I_dev;
  if (!rows_dev || !stats_dev || !hist_dev) return fail(B2R_ERR_INVALID, "b2r_cosine_stats: NULL buffer");
  if (n < 2 || d < 32 || d % 32) return fail(B2R_ERR_UNSUPPORTED, "b2r_cosine_stats: need n >= 2 and d a multiple of 32");
  if (n > 0x7FFFFFF0LL) return fail(B2R_ERR_UNSUPPORTED, "b2r_cosine_stats: n too large");
  if (n_bins < 2 || n_bins > 8192) return fail(B2R_ERR_INVALID, "b2r_cosine_stats: n_bins must be in [2, 8192]");
  if (row_begin < 0 || row_end > n || row_begin >= row_end) return fail(B2R_ERR_INVALID, "b2r_cosine_stats: bad column range");
  if (k != 0) {
    return fail(B2R_ERR_UNSUPPORTED,
                "b2r_cosine_stats: per-row top-k is served by b2r_flat_search_dev on the normalised rows (k must be 0 here)");
  }
  DeviceGuard g(device);
  if (!g.ok) return fail(B2R_ERR_CUDA, "b2r_cosine_stats: cudaSetDevice failed");
  cudaDeviceProp prop;
  B2R_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(B2R_ERR_UNSUPPORTED, "libb2r is built for sm_100a (B200) only");
  return cosine_stats_run(rows_dev, n, d, row_begin, row_end, stats_dev, reinterpret_cast<u64*>(hist_dev), n_bins,
                          prop.multiProcessorCount, static_cast<cudaStream_t>(stream));
}
