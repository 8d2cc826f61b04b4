// cosine_pair.cu — K5 on CTA pairs: N x N cosine statistics with 3xTF32 on tcgen05.mma.cta_group::2.
//
// Same arithmetic as the STATS epilogue of flat_gemm.cu (semantic_similarity_analysis/
// generate_similarity_matrix.py:84-98,162-171: S = clip(Xn Xn^T, -1, 1), off-diagonal mean / std / min / max and a
// histogram for the median; pairs i < j visited once and counted twice), reorganised so that the tensor pipe is
// the limiter instead of shared memory:
//   * the rows are normalised AND split once (hi = rna_tf32(xn), lo = xn - hi) into two arrays, so the scan kernel
//     has no transform warps: every operand reaches shared memory by TMA;
//   * tile = 256 rows (128 per CTA of the pair) x 256 columns (128 per CTA), M = 256 x N = 256 MMAs: the pair shares
//     the column operand.  Operand traffic decides this kernel: a tf32 MMA of K = 8 reads 4 KiB of rows and N x 32 B of
//     columns per SM; at N = 128 (r01) that is 8 KiB per 67 tensor cycles = 122 B/clk against the 128 B/clk shared-memory
//     port — the tensor pipe measured 53 % active —, at N = 256 it is 12 KiB per 134 cycles = 92 B/clk;
//   * two TMEM accumulators per tile, 2 x 256 columns = all of TMEM: hi*hi, and the cross terms lo*hi + hi*lo kept apart
//     from it (the tensor core truncates when it accumulates, so the small terms must not be added to the large running
//     sum).  r01's 2-way split of each over K (four accumulators) does not fit at N = 256; the chains are 128 MMAs long
//     instead of 64, still inside the 1e-5 bar (tests/test_cosine_gpu.py, N = 200k against float64);
//   * tiles are visited in L2-sized super-tiles (2048 x 2048) so that both operands of a tile come from L2.
#include <algorithm>
#include <mutex>
#include <vector>

#include "tc_ptx.cuh"

namespace b2r {

static const int kCosThreads = 320;   // producer, MMA, 8 epilogue warps (two per TMEM lane quadrant)
static const int kCosN = 256;        // columns per tile
static const int kCosSuper = 2048;   // rows / columns of an L2 super-tile
static const int kCosStages = 3;   // a 4th stage (all 227 KiB) was measured: no gain (r02) — the ring is not what limits this kernel

struct CosArgs {
  long long n;
  int d;
  long long n_tiles;
  const int2* tiles;     // (row pair-tile, column block) in visiting order
  long long col_lo, col_hi;
  double* stats;         // [0] sum, [1] sum of squares, [2] count
  u32* minmax;           // ordered float keys
  u64* hist;
  int n_bins;
};

// Xn = X / (|X| + 1e-8), split into tf32 hi / lo parts; one warp per row
__global__ void cosine_normalize_split_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ hi,
                                              float* __restrict__ lo) {
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
    for (int c = lane; c < d; c += 32) {
      const float v = row[c] * inv;
      const float h = rna_tf32(v);
      hi[r * d + c] = h;
      lo[r * d + c] = v - h;
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCosThreads, 1)
    cosine_pair_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                       const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                       const CosArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr int N = kCosN, NH = kCosN / 2, R = kCosStages;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr u32 kSubBytes = kTileM * kSlabK * 4;   // 16 KiB: 128 rows of one K-slab
  constexpr u32 kBBytes = NH * kSlabK * 4;         // 16 KiB: this CTA's 128 columns
  constexpr u32 kStage = 2 * kSubBytes + 2 * kBBytes;  // A_hi | A_lo | B_hi | B_lo
  constexpr u32 kTmemCols = 2 * N;                 // main (hi*hi), small (lo*hi + hi*lo)
  static_assert(kTmemCols == 512, "the accumulators fill TMEM");
  const int n_slabs = a.d / kSlabK;
  const u32 rank = cluster_ctarank();
  const bool leader = rank == 0;
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  unsigned char* ring = smem_raw;
  u32* hist_s = reinterpret_cast<u32*>(ring + static_cast<size_t>(R) * kStage);  // [n_bins]
  u64* bars = reinterpret_cast<u64*>(hist_s + ((a.n_bins + 1) & ~1));
  u64* full = bars;             // [R]
  u64* empty = full + R;        // [R]
  u64* tm_full = empty + R;     // [1]
  u64* tm_empty = tm_full + 1;  // [1]
  u32* tmem_slot = reinterpret_cast<u32*>(tm_empty + 1);

  if (tid == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tm_full, 1);
    mbar_init(tm_empty, 16);  // eight epilogue warps of each CTA
    mbar_fence_init();
  }
  for (int i = tid; i < a.n_bins; i += blockDim.x) hist_s[i] = 0;
  if (warp == 1) tc_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      u64 pol;  // default L2 policy: the operands are reused by the other tiles of the super-tile, then replaced
      asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
      int s = 0;
      u32 ph = 0;
      for (long long t = pair; t < a.n_tiles; t += n_pairs) {
        const int2 rc = a.tiles[t];
        const int row0 = rc.x * (2 * kTileM) + static_cast<int>(rank) * kTileM;
        const int col0 = rc.y * N + static_cast<int>(rank) * NH;
        for (int sl = 0; sl < n_slabs; ++sl) {
          mbar_wait_cluster_tag(&empty[s], ph ^ 1u, 100 + s);
          unsigned char* st = ring + static_cast<size_t>(s) * kStage;
          if (leader) mbar_arrive_expect_tx(&full[s], 2u * kStage);
          tma_load_2d_pair(st, &map_ahi, sl * kSlabK, row0, &full[s], pol);
          tma_load_2d_pair(st + kSubBytes, &map_alo, sl * kSlabK, row0, &full[s], pol);
          tma_load_2d_pair(st + 2 * kSubBytes, &map_bhi, sl * kSlabK, col0, &full[s], pol);
          tma_load_2d_pair(st + 2 * kSubBytes + kBBytes, &map_blo, sl * kSlabK, col0, &full[s], pol);
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
      constexpr u32 idesc = make_idesc_tf32_mn(2 * kTileM, N);
      int s = 0;
      u32 ph_s = 0, ph_t = 0;
      for (long long t = pair; t < a.n_tiles; t += n_pairs) {
        mbar_wait_cluster_tag(tm_empty, ph_t ^ 1u, 200);
        tc_fence_after();
        for (int sl = 0; sl < n_slabs; ++sl) {
          const u32 d_main = tmem_base;
          const u32 d_small = tmem_base + static_cast<u32>(N);
          const bool first = sl == 0;
          mbar_wait_cluster_tag(&full[s], ph_s, 400 + s);
          tc_fence_after();
          const u32 st = smem_u32(ring + static_cast<size_t>(s) * kStage);
          const u64 d_ahi = make_kmajor_sw128_desc(st);
          const u64 d_alo = make_kmajor_sw128_desc(st + kSubBytes);
          const u64 d_bhi = make_kmajor_sw128_desc(st + 2 * kSubBytes);
          const u64 d_blo = make_kmajor_sw128_desc(st + 2 * kSubBytes + kBBytes);
#pragma unroll
          for (int kk = 0; kk < kSlabK / 8; ++kk) {
            const u64 adv = static_cast<u64>(kk * 8 * 4 >> 4);
            const u32 acc = (first && kk == 0) ? 0u : 1u;
            tc_mma_tf32_pair(d_main, d_ahi + adv, d_bhi + adv, idesc, acc);
            tc_mma_tf32_pair(d_small, d_alo + adv, d_bhi + adv, idesc, acc);
            tc_mma_tf32_pair(d_small, d_ahi + adv, d_blo + adv, idesc, 1u);
          }
          tc_commit_pair(&empty[s]);
          if (++s == R) {
            s = 0;
            ph_s ^= 1u;
          }
        }
        tc_commit_pair(tm_full);
        ph_t ^= 1u;
      }
    }
  } else {
    // ======================= epilogue warps (both CTAs): statistics of the tile =======================
    // Warps 2-9: warp w reads TMEM lane quadrant w % 4 (a hardware rule) and one half of the 128 columns, so two
    // warps share a quadrant.  Per 16-column chunk the sum and the sum of squares are formed in fp32 (same order as
    // before the drain / statistics split: results are unchanged) and added to the float64 totals once; interior tiles
    // (every row below every column, all inside the matrix and the column range) skip the per-element masks.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // 0: columns 0-127, 1: columns 128-255
    u32 ph_t = 0;
    double st_sum = 0.0, st_sq = 0.0;
    unsigned long long st_cnt = 0;
    float st_min = 3.4028235e38f, st_max = -3.4028235e38f;
    const float scale = 0.5f * static_cast<float>(a.n_bins);
    const int last_bin = a.n_bins - 1;
    for (long long t = pair; t < a.n_tiles; t += n_pairs) {
      const int2 rc = a.tiles[t];
      const long long row_base = static_cast<long long>(rc.x) * (2 * kTileM);
      const long long row = row_base + static_cast<long long>(rank) * kTileM + quad * 32 + lane;
      const long long col_base = static_cast<long long>(rc.y) * N;
      const bool interior = row_base + 2 * kTileM - 1 < col_base && col_base + N <= a.n && col_base >= a.col_lo &&
                            col_base + N <= a.col_hi;
      mbar_wait_cluster_tag(tm_full, ph_t, 700);
      tc_fence_after();
      const u32 taddr = tmem_base + (static_cast<u32>(quad * 32) << 16) + static_cast<u32>(half * (N / 2));
      // (1) drain: this warp's 32 lanes x 128 columns of the two accumulators are combined into 128 registers per thread
      // and the accumulators are handed back at once — the MMAs of the NEXT tile run while (2) does the statistics.
      // (The accumulators fill all 512 TMEM columns, so they cannot be double-buffered in TMEM; registers can.)
      float v[N / 2];
#pragma unroll
      for (int c0 = 0; c0 < N / 2; c0 += 16) {
        float m0[16], s0[16];
        tc_ld16(taddr + c0, m0);
        tc_ld16(taddr + N + c0, s0);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[c0 + i] = m0[i] + s0[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tm_empty, 0);
      // (2) statistics from registers
#pragma unroll
      for (int c0 = 0; c0 < N / 2; c0 += 16) {
        float csum = 0.f, csq = 0.f;
        int ccnt = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const long long col = col_base + half * (N / 2) + c0 + i;
          if (interior || (row < a.n && col < a.n && row < col && col >= a.col_lo && col < a.col_hi)) {
            const float vv = fminf(1.0f, fmaxf(-1.0f, v[c0 + i]));
            csum += vv;
            csq = fmaf(vv, vv, csq);
            st_min = fminf(st_min, vv);
            st_max = fmaxf(st_max, vv);
            ++ccnt;
            const int bin = min(static_cast<int>((vv + 1.0f) * scale), last_bin);
            atomicAdd(&hist_s[bin], 1u);
          }
        }
        st_sum += static_cast<double>(csum);
        st_sq += static_cast<double>(csq);
        st_cnt += static_cast<unsigned long long>(ccnt);
      }
      ph_t ^= 1u;
    }
    named_bar_sync(1, 256);
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
    for (int b = tid - 64; b < a.n_bins; b += 256) {
      const u32 h = hist_s[b];
      if (h) atomicAdd(a.hist + b, static_cast<u64>(h));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc_pair(tmem_base, kTmemCols);
  }
}

// pairs were visited once (i < j); the matrix is symmetric, so sums and counts double
__global__ void cosine_pair_finalize_kernel(const double* acc, const u32* minmax, double* stats_out, u64* hist, int n_bins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bins) hist[i] *= 2ull;
  if (i == 0) {
    stats_out[0] = 2.0 * acc[0];
    stats_out[1] = 2.0 * acc[1];
    stats_out[2] = acc[2] > 0 ? static_cast<double>(ordered_to_float(minmax[0])) : 0.0;
    stats_out[3] = acc[2] > 0 ? static_cast<double>(ordered_to_float(minmax[1])) : 0.0;
    stats_out[4] = 2.0 * acc[2];
    stats_out[5] = 0.0;
  }
}

// Split operands (2 x n x d floats) and the tile list are kept between calls, per device: allocating and freeing
// 1.6 GB per call costs more than the kernel (cudaFree of a large block synchronises and unmaps).
struct CosScratch {
  float* hi = nullptr;
  float* lo = nullptr;
  size_t x_cap = 0;
  int2* tiles = nullptr;
  size_t tiles_cap = 0;
  unsigned char* small = nullptr;
};
static std::mutex g_cos_mu;
static CosScratch g_cos_scratch[16];

bool cosine_pair_supported(long long n, int d, int n_bins, int num_sms) {
  // d <= 1024: one accumulation chain per product is d / 8 MMAs long and the tensor core truncates at every step —
  // measured 5.5e-6 on a similarity of 1.0 at d = 1024, inside the 1e-5 bar; wider rows go to the single-CTA kernel,
  // which splits every chain in two
  return n >= 4096 && d % kSlabK == 0 && d <= 1024 && n_bins <= 8192 && num_sms >= 2 && !tunable("cosine.no_pair");
}

int cosine_stats_run_pair(const float* rows_dev, long long n, int d, long long col_begin, long long col_end,
                          double* stats_dev, u64* hist_dev, int n_bins, int num_sms, cudaStream_t stream) {
  float *xhi = nullptr, *xlo = nullptr;
  unsigned char* scratch = nullptr;
  int2* tiles_dev = nullptr;
  const size_t x_bytes = static_cast<size_t>(n) * d * sizeof(float);
  // tiles in L2 super-tile order: column super-tile, row super-tile (upper triangle), column block, row pair-tile
  std::vector<int2> tiles;
  const long long n_rt = (n + 2 * kTileM - 1) / (2 * kTileM), n_cb = (n + kCosN - 1) / kCosN;
  const long long rt_per = kCosSuper / (2 * kTileM), cb_per = kCosSuper / kCosN;
  const long long cb_lo = col_begin / kCosN, cb_hi = (col_end + kCosN - 1) / kCosN;
  for (long long sj = cb_lo / cb_per; sj * cb_per < cb_hi; ++sj) {
    for (long long si = 0; si * rt_per < n_rt; ++si) {
      if (si * kCosSuper > (sj + 1) * kCosSuper) break;  // rows entirely past the columns: nothing with i < j
      for (long long cb = std::max(cb_lo, sj * cb_per); cb < std::min({cb_hi, n_cb, (sj + 1) * cb_per}); ++cb) {
        const long long last_col = std::min<long long>(n, (cb + 1) * kCosN) - 1;
        for (long long rt = si * rt_per; rt < std::min(n_rt, (si + 1) * rt_per); ++rt) {
          if (rt * (2 * kTileM) < last_col) tiles.push_back(make_int2(static_cast<int>(rt), static_cast<int>(cb)));
        }
      }
    }
  }
  int rc = B2R_OK;
  int device = 0;
  cudaGetDevice(&device);
  std::lock_guard<std::mutex> lock(g_cos_mu);  // one analysis per process at a time: it uses the whole GPU anyway
  CosScratch& cs = g_cos_scratch[device & 15];
  cudaError_t e = cudaSuccess;
  if (cs.x_cap < x_bytes) {
    cudaFree(cs.hi);
    cudaFree(cs.lo);
    cs.hi = cs.lo = nullptr;
    cs.x_cap = 0;
    e = cudaMalloc(&cs.hi, x_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&cs.lo, x_bytes);
    if (e == cudaSuccess) cs.x_cap = x_bytes;
  }
  if (e == cudaSuccess && cs.tiles_cap < tiles.size()) {
    cudaFree(cs.tiles);
    cs.tiles = nullptr;
    cs.tiles_cap = 0;
    e = cudaMalloc(&cs.tiles, tiles.size() * sizeof(int2));
    if (e == cudaSuccess) cs.tiles_cap = tiles.size();
  }
  if (e == cudaSuccess && !cs.small) e = cudaMalloc(&cs.small, 256);
  if (e != cudaSuccess) {
    cudaFree(cs.hi);
    cudaFree(cs.lo);
    cs.hi = cs.lo = nullptr;
    cs.x_cap = 0;
    return fail(B2R_ERR_NOMEM, std::string("cosine: allocation failed: ") + cudaGetErrorString(e));
  }
  xhi = cs.hi;
  xlo = cs.lo;
  tiles_dev = cs.tiles;
  scratch = cs.small;
  double* acc = reinterpret_cast<double*>(scratch);
  u32* minmax = reinterpret_cast<u32*>(scratch + 64);
  do {
    const u32 mm_init[2] = {0xFFFFFFFFu, 0u};
    if (cudaMemsetAsync(acc, 0, 64, stream) != cudaSuccess ||
        cudaMemcpyAsync(minmax, mm_init, sizeof(mm_init), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaMemsetAsync(hist_dev, 0, static_cast<size_t>(n_bins) * sizeof(u64), stream) != cudaSuccess ||
        (!tiles.empty() && cudaMemcpyAsync(tiles_dev, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice,
                                           stream) != cudaSuccess)) {
      rc = fail(B2R_ERR_CUDA, "cosine: setup copies failed");
      break;
    }
    cosine_normalize_split_kernel<<<num_sms * 8, 256, 0, stream>>>(rows_dev, n, d, xhi, xlo);
    CUtensorMap mahi, malo, mbhi, mblo;
    rc = make_map_2d(&mahi, xhi, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kTileM);
    if (!rc) rc = make_map_2d(&malo, xlo, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kTileM);
    if (!rc) rc = make_map_2d(&mbhi, xhi, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kCosN / 2);
    if (!rc) rc = make_map_2d(&mblo, xlo, static_cast<unsigned long long>(d), static_cast<unsigned long long>(n), kSlabK, kCosN / 2);
    if (rc) break;
    if (!tiles.empty()) {
      CosArgs a;
      a.n = n;
      a.d = d;
      a.n_tiles = static_cast<long long>(tiles.size());
      a.tiles = tiles_dev;
      a.col_lo = col_begin;
      a.col_hi = col_end;
      a.stats = acc;
      a.minmax = minmax;
      a.hist = hist_dev;
      a.n_bins = n_bins;
      const size_t stage = 2 * static_cast<size_t>(kTileM) * kSlabK * 4 + 2 * static_cast<size_t>(kCosN / 2) * kSlabK * 4;
      const size_t smem = kCosStages * stage + static_cast<size_t>(n_bins + 1) * 4 + 16 * 8 + 64 + 1024;
      const int n_pairs = static_cast<int>(std::min<long long>(a.n_tiles, num_sms / 2));
      auto kern = cosine_pair_kernel;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
        rc = fail(B2R_ERR_CUDA, "cosine: cannot reserve shared memory");
        break;
      }
      kern<<<2 * n_pairs, kCosThreads, smem, stream>>>(mahi, malo, mbhi, mblo, a);
    }
    cosine_pair_finalize_kernel<<<(n_bins + 255) / 256, 256, 0, stream>>>(acc, minmax, stats_dev, hist_dev, n_bins);
    if (cudaGetLastError() != cudaSuccess) rc = fail(B2R_ERR_CUDA, "cosine: launch failed");
  } while (0);
  if (cudaStreamSynchronize(stream) != cudaSuccess && rc == B2R_OK) rc = fail(B2R_ERR_CUDA, "cosine: kernel failed");
  return rc;
}

}  // namespace b2r
