// flat_gemm_wide.cu — K2 candidate pass, threshold form: up to 128 queries per pass over the rows on one CTA per SM,
// up to 256 on CTA pairs (flat_gemm_pair_kernel, below).
//
// Same job as the TERMS = 1 pass of flat_gemm.cu (approximate tf32 scores select k' candidates per
// query, exact fp32 rescoring and the completeness proof follow), but the select epilogue no longer
// maintains sorted per-query lists while it streams.  It compares against a per-query threshold that
// is known BEFORE the scan:
//
//   1. pre-pass   the scan kernel runs over the first few tiles of every CTA (about 1.4 % of the
//                 rows) and writes, per query, the maximum score of every 32-row group.
//   2. threshold  tau_q = the k'-th largest group maximum (minus 4 ulps).  Those are scores of k'
//                 distinct rows, so at least k' rows score >= tau_q: a valid lower bound of the k'-th
//                 best score, and about as tight as the k'-th best of the sample.
//   3. main pass  every row whose score is >= tau_q is appended to the CTA's list of query q in
//                 global memory (expected ~25 appends per list, capacity 128; an overflowing list flags the
//                 query and the robust 3xTF32 pass of flat_gemm.cu redoes it).
//   4. select     one CTA per query gathers the lists and sorts them: the k' best approximate scores.
//
// Pre-pass and main pass are the same kernel instance, tile alignment and MMA order, so a row gets
// the same score in both.  The epilogue is then one compare per score and a rare append, which
// lets one pass serve N = 128 queries at the HBM rate (flat_gemm.cu's list epilogue stalled at 64).
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles of SUB x 128 rows):
//   warp 0     TMA producer: per K-slab SUB copies of A [128 x 32] + the query slab B [N x 32]
//   warp 1     TMEM allocation; one lane issues tcgen05.mma kind::tf32 (M = 128, N, K = 8)
//   warps 2-5  epilogue: tcgen05.ld 32 columns at a time (thread = row), threshold compare / group max
// Accumulators: one per sub-tile (no split-K; the accumulation error is part of the rescoring bound),
// two tiles in flight: 2 * SUB * N TMEM columns (512 at N = 128, SUB = 2).
//
// L2 traffic per byte of A: N / (SUB * 128) bytes of B — 0.5 at N = 128 with 256-row tiles.
#include <cuda_bf16.h>

#include <algorithm>

#include "tc_ptx.cuh"

namespace b2r {

static const int kWideThreads = 192;
static const int kListCap = 128;       // entries per (CTA, query) list
static const int kPremaxMax = 8192;    // group maxima per query the threshold kernel can sort
static const int kGatherCap = 8192;    // keys the final select sorts in shared memory

struct WideArgs {
  long long n;
  int d;
  int nq;                // active queries (<= N)
  int stages;
  long long n_tiles;     // tiles of SUB * 128 rows
  int tile_limit;        // tiles per CTA (pre-pass: a few; main pass: INT_MAX)
  // pre-pass output
  u32* premax;           // [N][pm_stride] ordered-float group maxima; NULL in the main pass
  int pm_stride;
  // main pass
  const float* tauf;     // [N] thresholds (+inf for padding queries)
  u64* glists;           // [grid][N][kListCap]
  u32* gcnt;             // [grid][N]
  u32* ovf;              // [N] set when a list overflowed
};

template <int N, int SUB>
__global__ void __launch_bounds__(kWideThreads, 1)
    flat_gemm_thr_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const WideArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr u32 kSubBytes = kTileM * kSlabK * 4;   // 16 KiB
  constexpr u32 kABytes = SUB * kSubBytes;
  constexpr u32 kBBytes = N * kSlabK * 4;
  constexpr u32 kStage = kABytes + kBBytes;
  constexpr u32 kTileCols = SUB * N;
  constexpr u32 kTmemCols = 2 * kTileCols < 32 ? 32 : 2 * kTileCols;
  const int R = a.stages;
  const int n_slabs = a.d / kSlabK;
  const bool pre = a.premax != nullptr;

  unsigned char* ring = smem_raw;
  u64* bars = reinterpret_cast<u64*>(ring + static_cast<size_t>(R) * kStage);
  u64* full = bars;             // [R]
  u64* empty = full + R;        // [R]
  u64* tm_full = empty + R;     // [2]
  u64* tm_empty = tm_full + 2;  // [2]
  float* tauf_s = reinterpret_cast<float*>(tm_empty + 2);  // [N]
  u32* cnt_s = reinterpret_cast<u32*>(tauf_s + N);         // [N]
  u32* tmem_slot = cnt_s + N;

  if (tid == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tm_full[b], 1);
      mbar_init(&tm_empty[b], 4);
    }
    mbar_fence_init();
  }
  for (int i = tid; i < N; i += blockDim.x) {
    tauf_s[i] = (!pre && i < a.nq) ? a.tauf[i] : __int_as_float(0x7f800000);
    cnt_s[i] = 0;
  }
  if (warp == 1) tc_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;
  const long long n_tiles = a.n_tiles;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      const u64 pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
      int s = 0, j = 0;
      u32 ph = 0;
      for (long long t = blockIdx.x; t < n_tiles && j < a.tile_limit; t += gridDim.x, ++j) {
        const int row0 = static_cast<int>(t * (kTileM * SUB));
        for (int sl = 0; sl < n_slabs; ++sl) {
          mbar_wait_tag(&empty[s], ph ^ 1u, 100 + s);
          unsigned char* st = ring + static_cast<size_t>(s) * kStage;
          mbar_arrive_expect_tx(&full[s], kStage);
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            tma_load_2d(st + sub * kSubBytes, &map_a, sl * kSlabK, row0 + sub * kTileM, &full[s], pol_a);
          }
          tma_load_2d(st + kABytes, &map_b, sl * kSlabK, 0, &full[s], pol_b);
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
      int s = 0, buf = 0, j = 0;
      u32 ph_s = 0, ph_buf[2] = {0, 0};
      for (long long t = blockIdx.x; t < n_tiles && j < a.tile_limit; t += gridDim.x, ++j) {
        mbar_wait_tag(&tm_empty[buf], ph_buf[buf] ^ 1u, 200 + buf);
        tc_fence_after();
        const u32 tile_d = tmem_base + static_cast<u32>(buf) * kTileCols;
        for (int sl = 0; sl < n_slabs; ++sl) {
          mbar_wait_tag(&full[s], ph_s, 400 + s);
          tc_fence_after();
          const u32 st = smem_u32(ring + static_cast<size_t>(s) * kStage);
          const u64 d_b = make_kmajor_sw128_desc(st + kABytes);
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            const u64 d_a = make_kmajor_sw128_desc(st + sub * kSubBytes);
            const u32 d_acc = tile_d + static_cast<u32>(sub * N);
#pragma unroll
            for (int kk = 0; kk < kSlabK / 8; ++kk) {
              const u64 adv = static_cast<u64>(kk * 8 * 4 >> 4);
              tc_mma_tf32(d_acc, d_a + adv, d_b + adv, idesc, (sl == 0 && kk == 0) ? 0u : 1u);
            }
          }
          tc_commit(&empty[s]);
          if (++s == R) {
            s = 0;
            ph_s ^= 1u;
          }
        }
        tc_commit(&tm_full[buf]);
        ph_buf[buf] ^= 1u;
        buf ^= 1;
      }
    }
  } else {
    // ======================= epilogue warps =======================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int buf = 0, j = 0;
    u32 ph_buf[2] = {0, 0};
    u64* my_lists = a.glists + static_cast<size_t>(blockIdx.x) * N * kListCap;
    for (long long t = blockIdx.x; t < n_tiles && j < a.tile_limit; t += gridDim.x, ++j) {
      mbar_wait_tag(&tm_full[buf], ph_buf[buf], 700 + buf);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < SUB; ++sub) {
        const long long row = t * (kTileM * SUB) + sub * kTileM + quad * 32 + lane;
        const bool row_ok = row < a.n;
        const u32 taddr = tmem_base + (static_cast<u32>(quad * 32) << 16) + static_cast<u32>(buf) * kTileCols +
                          static_cast<u32>(sub * N);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tc_ld16(taddr + c0, v);
          if (N > 16) tc_ld16(taddr + c0 + 16, v + 16);
          tc_wait_ld();
          if (sub == SUB - 1 && c0 + 32 >= N) {  // the whole tile has been read: hand the accumulators back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tm_empty[buf]);
          }
          constexpr int kCols = N > 16 ? 32 : 16;
          if (pre) {
            u32 keep = 0;
#pragma unroll
            for (int i = 0; i < kCols; ++i) {
              const u32 o = row_ok ? float_to_ordered(v[i]) : 0u;
              const u32 m = __reduce_max_sync(0xffffffffu, o);
              if (lane == i) keep = m;
            }
            if (lane < kCols && c0 + lane < a.nq) {
              const int slot = ((blockIdx.x * a.tile_limit + j) * SUB + sub) * 4 + (warp - 2);
              a.premax[static_cast<size_t>(c0 + lane) * a.pm_stride + slot] = keep;
            }
          } else {
#pragma unroll
            for (int i = 0; i < kCols; i += 4) {
              const float4 tq = *reinterpret_cast<const float4*>(tauf_s + c0 + i);
              const float th[4] = {tq.x, tq.y, tq.z, tq.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (!(v[i + u] < th[u])) {  // rare; NaN scores take the slow path and sort last
                  const int c = c0 + i + u;
                  if (row_ok && c < a.nq) {
                    const u32 pos = atomicAdd(&cnt_s[c], 1u);
                    if (pos < static_cast<u32>(kListCap)) {
                      my_lists[static_cast<size_t>(c) * kListCap + pos] = make_key(v[i + u], static_cast<u32>(row));
                    } else {
                      a.ovf[c] = 1u;
                    }
                  }
                }
              }
            }
          }
        }
      }
      ph_buf[buf] ^= 1u;
      buf ^= 1;
    }
    if (!pre) {
      named_bar_sync(1, 128);
      for (int c = tid - 64; c < N; c += 128) {
        const u32 nc = cnt_s[c];
        a.gcnt[static_cast<size_t>(blockIdx.x) * N + c] = nc < static_cast<u32>(kListCap) ? nc : kListCap;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: 256 queries per pass.  Two CTAs of a cluster (the two SMs of a TPC) execute ONE
// tcgen05.mma.cta_group::2 of M = 256 rows x N = 256 queries per k-step: each CTA streams its own 128-row
// sub-tiles of A and HALF of the query slab (128 queries); the tensor cores of both SMs read both halves, so the
// query operand costs each SM the same L2 traffic and shared-memory space as in the 128-query kernel while twice
// as many queries share every byte read from HBM.  Roles per CTA as above; only the leader CTA (cluster rank 0)
// issues MMAs.  Barriers: TMA completion bytes of both CTAs are counted on the leader's `full` barrier, the MMA
// commits are multicast to both CTAs' `empty` / `tm_full` barriers, and the epilogue warps of both CTAs arrive
// on the leader's `tm_empty`.  TMEM: SUB x 256 columns per tile x BUFS tiles = 512 columns.
// ---------------------------------------------------------------------------------------------
template <int SUB, int BUFS, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWideThreads, 1)
    flat_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const WideArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr int N = 256, NH = 128;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr u32 kSubBytes = kTileM * kSlabK * 4;   // 16 KiB
  constexpr u32 kABytes = SUB * kSubBytes;
  constexpr u32 kBBytes = NH * kSlabK * 4;         // this CTA's half of the query slab
  constexpr u32 kStage = kABytes + kBBytes;
  constexpr u32 kTileCols = SUB * N;
  constexpr u32 kTmemCols = BUFS * kTileCols;
  static_assert(kTmemCols == 512, "the pair kernel uses all of TMEM");
  // a K-slab is one 128-byte swizzle row: 32 floats (tf32 MMAs on the fp32 rows) or 64 bf16 (the optional bf16 copy)
  constexpr int kSlabElems = BF16 ? 64 : kSlabK;
  const int R = a.stages;
  const int n_slabs = a.d / kSlabElems;
  const bool pre = a.premax != nullptr;
  const u32 rank = cluster_ctarank();
  const bool leader = rank == 0;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  constexpr int kPairRows = 2 * SUB * kTileM;

  unsigned char* ring = smem_raw;
  u64* bars = reinterpret_cast<u64*>(ring + static_cast<size_t>(R) * kStage);
  u64* full = bars;             // [R]  (the leader's copy is the one in use)
  u64* empty = full + R;        // [R]
  u64* tm_full = empty + R;     // [2]
  u64* tm_empty = tm_full + 2;  // [2]  (leader's copy)
  float* tauf_s = reinterpret_cast<float*>(tm_empty + 2);  // [N]
  u32* cnt_s = reinterpret_cast<u32*>(tauf_s + N);         // [N]
  u32* tmem_slot = cnt_s + N;

  if (tid == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tm_full[b], 1);
      mbar_init(&tm_empty[b], 8);  // four epilogue warps of each CTA
    }
    mbar_fence_init();
  }
  for (int i = tid; i < N; i += blockDim.x) {
    tauf_s[i] = (!pre && i < a.nq) ? a.tauf[i] : __int_as_float(0x7f800000);
    cnt_s[i] = 0;
  }
  if (warp == 1) tc_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive or TMA completion
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;
  const long long n_tiles = a.n_tiles;  // pair tiles of kPairRows rows

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      const u64 pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
      int s = 0, j = 0;
      u32 ph = 0;
      for (long long t = pair; t < n_tiles && j < a.tile_limit; t += n_pairs, ++j) {
        const int row0 = static_cast<int>(t * kPairRows + static_cast<long long>(rank) * (SUB * kTileM));
        for (int sl = 0; sl < n_slabs; ++sl) {
          mbar_wait_cluster_tag(&empty[s], ph ^ 1u, 100 + s);
          unsigned char* st = ring + static_cast<size_t>(s) * kStage;
          if (leader) mbar_arrive_expect_tx(&full[s], 2u * kStage);  // bytes of both CTAs land on this barrier
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            tma_load_2d_pair(st + sub * kSubBytes, &map_a, sl * kSlabElems, row0 + sub * kTileM, &full[s], pol_a);
          }
          tma_load_2d_pair(st + kABytes, &map_b, sl * kSlabElems, static_cast<int>(rank) * NH, &full[s], pol_b);
          if (++s == R) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) =======================
    if (lane == 0 && leader) {
      constexpr u32 idesc = BF16 ? make_idesc_bf16_mn(2 * kTileM, N) : make_idesc_tf32_mn(2 * kTileM, N);
      int s = 0, buf = 0, j = 0;
      u32 ph_s = 0, ph_buf[2] = {0, 0};
      for (long long t = pair; t < n_tiles && j < a.tile_limit; t += n_pairs, ++j) {
        mbar_wait_cluster_tag(&tm_empty[buf], ph_buf[buf] ^ 1u, 200 + buf);
        tc_fence_after();
        const u32 tile_d = tmem_base + static_cast<u32>(buf) * kTileCols;
        for (int sl = 0; sl < n_slabs; ++sl) {
          mbar_wait_cluster_tag(&full[s], ph_s, 400 + s);
          tc_fence_after();
          const u32 st = smem_u32(ring + static_cast<size_t>(s) * kStage);
          const u64 d_b = make_kmajor_sw128_desc(st + kABytes);
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            const u64 d_a = make_kmajor_sw128_desc(st + sub * kSubBytes);
            const u32 d_acc = tile_d + static_cast<u32>(sub * N);
#pragma unroll
            for (int kk = 0; kk < kSlabK / 8; ++kk) {
              const u64 adv = static_cast<u64>(kk * 8 * 4 >> 4);
              if (BF16) tc_mma_bf16_pair(d_acc, d_a + adv, d_b + adv, idesc, (sl == 0 && kk == 0) ? 0u : 1u);
              else tc_mma_tf32_pair(d_acc, d_a + adv, d_b + adv, idesc, (sl == 0 && kk == 0) ? 0u : 1u);
            }
          }
          tc_commit_pair(&empty[s]);
          if (++s == R) {
            s = 0;
            ph_s ^= 1u;
          }
        }
        tc_commit_pair(&tm_full[buf]);
        ph_buf[buf] ^= 1u;
        if (BUFS == 2) buf ^= 1;
      }
    }
  } else {
    // ======================= epilogue warps (both CTAs) =======================
    const int quad = warp & 3;
    int buf = 0, j = 0;
    u32 ph_buf[2] = {0, 0};
    u64* my_lists = a.glists + static_cast<size_t>(blockIdx.x) * N * kListCap;
    for (long long t = pair; t < n_tiles && j < a.tile_limit; t += n_pairs, ++j) {
      mbar_wait_cluster_tag(&tm_full[buf], ph_buf[buf], 700 + buf);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < SUB; ++sub) {
        const long long row = t * kPairRows + static_cast<long long>(rank) * (SUB * kTileM) + sub * kTileM + quad * 32 + lane;
        const bool row_ok = row < a.n;
        const u32 taddr = tmem_base + (static_cast<u32>(quad * 32) << 16) + static_cast<u32>(buf) * kTileCols +
                          static_cast<u32>(sub * N);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tc_ld16(taddr + c0, v);
          tc_ld16(taddr + c0 + 16, v + 16);
          tc_wait_ld();
          if (sub == SUB - 1 && c0 + 32 >= N) {  // the whole tile has been read: hand the accumulators back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tm_empty[buf], 0);
          }
          if (pre) {
            u32 keep = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const u32 o = row_ok ? float_to_ordered(v[i]) : 0u;
              const u32 m = __reduce_max_sync(0xffffffffu, o);
              if (lane == i) keep = m;
            }
            if (c0 + lane < a.nq) {
              const int slot = ((blockIdx.x * a.tile_limit + j) * SUB + sub) * 4 + (warp - 2);
              a.premax[static_cast<size_t>(c0 + lane) * a.pm_stride + slot] = keep;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 tq = *reinterpret_cast<const float4*>(tauf_s + c0 + i);
              const float th[4] = {tq.x, tq.y, tq.z, tq.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (!(v[i + u] < th[u])) {
                  const int c = c0 + i + u;
                  if (row_ok && c < a.nq) {
                    const u32 pos = atomicAdd(&cnt_s[c], 1u);
                    if (pos < static_cast<u32>(kListCap)) {
                      my_lists[static_cast<size_t>(c) * kListCap + pos] = make_key(v[i + u], static_cast<u32>(row));
                    } else {
                      a.ovf[c] = 1u;
                    }
                  }
                }
              }
            }
          }
        }
      }
      ph_buf[buf] ^= 1u;
      if (BUFS == 2) buf ^= 1;
    }
    if (!pre) {
      named_bar_sync(1, 128);
      for (int c = tid - 64; c < N; c += 128) {
        const u32 nc = cnt_s[c];
        a.gcnt[static_cast<size_t>(blockIdx.x) * N + c] = nc < static_cast<u32>(kListCap) ? nc : kListCap;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc_pair(tmem_base, kTmemCols);
  }
}

// raw queries -> B operand [N][d] (rows nq..N-1 zero); also clears the overflow flags
__global__ void wide_prep_kernel(const float* __restrict__ q, int nq, int N, int d, float* __restrict__ B,
                                 u32* __restrict__ ovf) {
  const long long total = static_cast<long long>(N) * d;
  const long long live = static_cast<long long>(nq) * d;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i0 < N) ovf[i0] = 0;
  for (long long i = i0; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) B[i] = i < live ? q[i] : 0.f;
}

// bf16 variant of the query operand (round to nearest even, as the rows' copy)
__global__ void wide_prep_bf16_kernel(const float* __restrict__ q, int nq, int N, int d, unsigned short* __restrict__ B,
                                      u32* __restrict__ ovf) {
  const long long total = static_cast<long long>(N) * d;
  const long long live = static_cast<long long>(nq) * d;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i0 < N) ovf[i0] = 0;
  for (long long i = i0; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    B[i] = i < live ? __bfloat16_as_ushort(__float2bfloat16_rn(q[i])) : static_cast<unsigned short>(0);
  }
}

// rows -> bf16 copy (the optional candidate-pass operand), n * d elements
__global__ void rows_to_bf16_kernel(const float* __restrict__ x, long long total, unsigned short* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    out[i] = __bfloat16_as_ushort(__float2bfloat16_rn(x[i]));
  }
}

int rows_to_bf16(const float* rows, long long n, int d, void* out, int num_sms, cudaStream_t stream) {
  rows_to_bf16_kernel<<<num_sms * 16, 256, 0, stream>>>(rows, n * d, static_cast<unsigned short*>(out));
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

bool wide_uses_bf16(const void* rows_bf16, int d, int nq) {
  return rows_bf16 != nullptr && nq > 128 && d % 64 == 0 && !tunable("gemm.no_pair") && !tunable("gemm.no_bf16");
}

// one CTA per query: tau = k'-th largest group maximum, 4 ulps lower (insurance; costs nothing)
__global__ void __launch_bounds__(1024) wide_tau_kernel(const u32* __restrict__ premax, int pm_stride, int m, int kp,
                                                        float* __restrict__ tauf) {
  __shared__ u32 buf[kPremaxMax];
  __shared__ u32 hist[kSelBins + 4];
  __shared__ u32 top[512];
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < m; i += blockDim.x) buf[i] = premax[static_cast<size_t>(q) * pm_stride + i];
  __syncthreads();
  block_topk_hist(buf, m, kPremaxMax, kp, hist, top, 512);
  if (threadIdx.x == 0) {
    const u32 o = top[kp - 1];
    tauf[q] = o > 4u ? ordered_to_float(o - 4u) : __int_as_float(0xff800000);  // -inf: keep every row
  }
}

// one CTA per query: gather the per-CTA lists and select the k' best (approximate score, local id)
__global__ void __launch_bounds__(1024) wide_select_kernel(const u64* __restrict__ glists, const u32* __restrict__ gcnt,
                                                           int n_ctas, int N, int kp, float* __restrict__ Dp,
                                                           long long* __restrict__ Ip, u32* __restrict__ ovf) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  u64* buf = reinterpret_cast<u64*>(sel_smem);
  __shared__ u64 top[512];
  __shared__ u32 hist[kSelBins + 4];
  __shared__ u32 s_cnt;
  const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int c = warp; c < n_ctas; c += 32) {
    const u32 nc = gcnt[static_cast<size_t>(c) * N + q];
    const u64* list = glists + (static_cast<size_t>(c) * N + q) * kListCap;
    for (u32 i = lane; i < nc; i += 32) {
      const u64 key = list[i];
      const u32 pos = atomicAdd(&s_cnt, 1u);
      if (pos < static_cast<u32>(kGatherCap)) buf[pos] = key;
    }
  }
  __syncthreads();
  u32 total = s_cnt;
  if (total > static_cast<u32>(kGatherCap)) {  // cannot happen with a sane threshold; the robust pass redoes the query
    if (tid == 0) ovf[q] = 1u;
    total = kGatherCap;
  }
  block_topk_hist(buf, static_cast<int>(total), kGatherCap, kp, hist, top, 512);
  block_emit_keys(top, kp, Dp + static_cast<size_t>(q) * kp, Ip + static_cast<size_t>(q) * kp, 0);
}

template <int N, int SUB>
static int wide_launch(const CUtensorMap& ma, const CUtensorMap& mb, const WideArgs& a, int grid, size_t smem,
                       cudaStream_t stream) {
  auto kern = flat_gemm_thr_kernel<N, SUB>;
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<grid, kWideThreads, smem, stream>>>(ma, mb, a);
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

static int wide_dispatch(int N, int sub, const CUtensorMap& ma, const CUtensorMap& mb, const WideArgs& a, int grid,
                         size_t smem, cudaStream_t stream) {
  if (sub == 1) {
    switch (N) {
      case 16: return wide_launch<16, 1>(ma, mb, a, grid, smem, stream);
      case 32: return wide_launch<32, 1>(ma, mb, a, grid, smem, stream);
      case 64: return wide_launch<64, 1>(ma, mb, a, grid, smem, stream);
      default: return wide_launch<128, 1>(ma, mb, a, grid, smem, stream);
    }
  }
  switch (N) {
    case 16: return wide_launch<16, 2>(ma, mb, a, grid, smem, stream);
    case 32: return wide_launch<32, 2>(ma, mb, a, grid, smem, stream);
    case 64: return wide_launch<64, 2>(ma, mb, a, grid, smem, stream);
    default: return wide_launch<128, 2>(ma, mb, a, grid, smem, stream);
  }
}

bool wide_path_supported(long long n, int d, int kp) {
  return d % kSlabK == 0 && d >= kSlabK && kp <= kListCap && n >= 16384 && !tunable("gemm.no_wide");
}

template <int SUB, int BUFS, bool BF16>
static int pair_launch(const CUtensorMap& ma, const CUtensorMap& mb, const WideArgs& a, int grid, size_t smem,
                       cudaStream_t stream) {
  auto kern = flat_gemm_pair_kernel<SUB, BUFS, BF16>;
  B2R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<grid, kWideThreads, smem, stream>>>(ma, mb, a);  // cluster shape (2, 1, 1) is part of the kernel
  B2R_CUDA(cudaGetLastError());
  return B2R_OK;
}

// Candidate pass for nq <= 256 queries: fills Dp / Ip [nq][kp] (approximate scores, local ids) and ovf[nq].
// B: scratch for the query operand (256 * d floats); glists: partial-list scratch of glists_bytes.
// rows_bf16 != NULL (pair path only): the candidate pass reads the bf16 copy of the rows instead.
int wide_candidates(const float* rows, const void* rows_bf16, long long n, int d, const float* q_dev, int nq, int kp,
                    float* B, float* tauf, u32* premax, u32* gcnt, u32* ovf, u64* glists, size_t glists_bytes, float* Dp,
                    long long* Ip, int num_sms, cudaStream_t stream, int* launches) {
  const bool pair = nq > 128;
  const bool bf16 = pair && rows_bf16 != nullptr;
  const int N = pair ? 256 : (nq <= 16 ? 16 : (nq <= 32 ? 32 : (nq <= 64 ? 64 : 128)));
  // 128-row sub-tiles per CTA and tile: 2 for the 64 / 128-query kernels (halves the L2 re-reads of the query operand);
  // the pair kernel runs faster with one sub-tile and two tiles in flight (3.64 vs 4.31 ms per 256 queries at 4M rows)
  int sub = (!pair && N >= 64) ? 2 : 1;
  if (tunable("gemm.wide_sub") == 1 || tunable("gemm.wide_sub") == 2) sub = tunable("gemm.wide_sub");
  const int tile_rows = (pair ? 2 : 1) * sub * kTileM;  // rows per tile (pair: per pair of CTAs)
  const long long n_tiles = (n + tile_rows - 1) / tile_rows;
  const int units = pair ? num_sms / 2 : num_sms;        // CTAs, or CTA pairs
  int n_units = static_cast<int>(std::min<long long>(n_tiles, units));
  const size_t per_cta = static_cast<size_t>(N) * kListCap * sizeof(u64);
  n_units = static_cast<int>(std::min<size_t>(n_units, glists_bytes / (per_cta * (pair ? 2 : 1))));
  if (n_units < 1) return fail(B2R_ERR_INVALID, "wide_candidates: list scratch too small");
  const int grid = pair ? 2 * n_units : n_units;
  // sample size: aim at ~28 appends per (CTA, query) list in the main pass
  const long long tiles_per_unit = (n_tiles + n_units - 1) / n_units;
  const long long rows_per_cta = tiles_per_unit * (sub * kTileM);
  const long long sample_per_tile = static_cast<long long>(grid) * sub * kTileM;  // rows one pre-pass tile adds
  long long pre_tiles = (rows_per_cta * kp + 28LL * sample_per_tile - 1) / (28LL * sample_per_tile);
  const int pre_cap = kPremaxMax / (grid * sub * 4);
  pre_tiles = std::max<long long>(1, std::min<long long>(pre_tiles, std::min<long long>(std::max(pre_cap, 1), tiles_per_unit)));
  const int m = grid * static_cast<int>(pre_tiles) * sub * 4;
  if (m > kPremaxMax) return fail(B2R_ERR_UNSUPPORTED, "wide_candidates: too many CTAs for the threshold sample");

  if (bf16) wide_prep_bf16_kernel<<<128, 256, 0, stream>>>(q_dev, nq, N, d, reinterpret_cast<unsigned short*>(B), ovf);
  else wide_prep_kernel<<<128, 256, 0, stream>>>(q_dev, nq, N, d, B, ovf);
  B2R_CUDA(cudaGetLastError());
  B2R_CUDA(cudaMemsetAsync(premax, 0, static_cast<size_t>(nq) * m * sizeof(u32), stream));
  CUtensorMap ma, mb;
  int rc;
  if (bf16) {
    rc = make_map_2d_bf16(&ma, rows_bf16, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), 64, kTileM);
    if (rc) return rc;
    rc = make_map_2d_bf16(&mb, B, static_cast<unsigned long long>(d), static_cast<unsigned long long>(N), 64, 128);
  } else {
    rc = make_map_2d(&ma, rows, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kTileM);
    if (rc) return rc;
    rc = make_map_2d(&mb, B, static_cast<unsigned long long>(d), static_cast<unsigned long long>(N), kSlabK, pair ? 128 : N);
  }
  if (rc) return rc;
  const size_t stage = static_cast<size_t>(sub * kTileM) * kSlabK * 4 + static_cast<size_t>(pair ? 128 : N) * kSlabK * 4;
  const size_t fixed = 32 * 8 + static_cast<size_t>(N) * 8 + 64 + 1024;
  int stages = tunable("gemm.stages");
  if (stages <= 0) stages = 8;
  const size_t limit = 227 * 1024 - 1024;
  while (stages > 2 && stages * stage + fixed > limit) --stages;
  const size_t smem = stages * stage + fixed;
  WideArgs a;
  a.n = n;
  a.d = d;
  a.nq = nq;
  a.stages = stages;
  a.n_tiles = n_tiles;
  a.tile_limit = static_cast<int>(pre_tiles);
  a.premax = premax;
  a.pm_stride = m;
  a.tauf = nullptr;
  a.glists = glists;
  a.gcnt = gcnt;
  a.ovf = ovf;
  const bool single_buf = sub == 2;  // pair kernel: 2 sub-tiles x 256 columns fill TMEM, one tile in flight
  auto scan = [&](const WideArgs& wa) {
    if (!pair) return wide_dispatch(N, sub, ma, mb, wa, grid, smem, stream);
    if (bf16) {
      return single_buf ? pair_launch<2, 1, true>(ma, mb, wa, grid, smem, stream)
                        : pair_launch<1, 2, true>(ma, mb, wa, grid, smem, stream);
    }
    return single_buf ? pair_launch<2, 1, false>(ma, mb, wa, grid, smem, stream)
                      : pair_launch<1, 2, false>(ma, mb, wa, grid, smem, stream);
  };
  rc = scan(a);
  if (rc) return rc;
  wide_tau_kernel<<<nq, 1024, 0, stream>>>(premax, m, m, kp, tauf);
  B2R_CUDA(cudaGetLastError());
  a.tile_limit = 0x7fffffff;
  a.premax = nullptr;
  a.tauf = tauf;
  rc = scan(a);
  if (rc) return rc;
  auto sel = wide_select_kernel;
  const size_t sel_smem = static_cast<size_t>(kGatherCap) * sizeof(u64);
  B2R_CUDA(cudaFuncSetAttribute(sel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sel_smem)));
  sel<<<nq, 1024, sel_smem, stream>>>(glists, gcnt, grid, N, kp, Dp, Ip, ovf);
  B2R_CUDA(cudaGetLastError());
  *launches += 5;
  return B2R_OK;
}

}  // namespace b2r
