// tc_ptx.cuh — PTX wrappers shared by the tensor-core kernels (tcgen05 / TMEM / 2-D TMA), sm_100a only.
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdio.h>

#include "b2r_common.cuh"

namespace b2r {

static const int kSlabK = 32;          // floats per K-slab = one 128-byte swizzle row
static const int kTileM = 128;

#ifdef __CUDACC__
// ---------------- PTX wrappers specific to this kernel ----------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, u64* bar,
                                            u64 policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

__device__ __forceinline__ u64 l2_policy_evict_last() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_alloc(u32* slot, u32 ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(u32 taddr, u32 ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void tc_mma_tf32(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void tc_commit(u64* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tc_ld16(u32 taddr, float* v) {
  u32 r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest (ties away) to tf32 precision with two integer ops (cvt.rna.tf32.f32 runs on
// the slow conversion pipe): add half an ulp of the 10-bit mantissa to the magnitude, clear 13 bits
__device__ __forceinline__ float rna_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// mbarrier wait with a watchdog: a pipeline bug shows up as a trap with the barrier's tag, not a hang
__device__ __forceinline__ void mbar_wait_tag(u64* bar, u32 parity, int tag) {
  const u32 addr = smem_u32(bar);
  for (u32 spins = 0;; ++spins) {
    u32 done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins == (1u << 26)) {
      printf("b2r flat_gemm: barrier wait stuck: tag %d block %d thread %d parity %u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups
// 1024 B apart (the layout a {32 x rows} fp32 TMA box with SWIZZLE_128B produces)
__device__ __forceinline__ u64 make_kmajor_sw128_desc(u32 smem_addr) {
  u64 desc = 0;
  desc |= static_cast<u64>((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  desc |= static_cast<u64>(1) << 16;                      // leading byte offset (unused for swizzled K-major)
  desc |= static_cast<u64>(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
  desc |= static_cast<u64>(1) << 46;                      // descriptor version (Blackwell)
  desc |= static_cast<u64>(2) << 61;                      // SWIZZLE_128B
  return desc;
}

// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N
__host__ __device__ constexpr u32 make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<u32>(n >> 3) << 17) | (static_cast<u32>(kTileM >> 4) << 24);
}

// ---------------- CTA pair (cta_group::2): two SMs of a TPC execute one M = 256 MMA ----------------
__device__ __forceinline__ u32 cluster_ctarank() {
  u32 r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D tensor copy into THIS CTA's shared memory whose completion bytes are counted on the LEADER CTA's mbarrier
// (bit 24 of a shared::cluster address selects the odd CTA of the pair: clearing it addresses CTA 0's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, u64* bar,
                                                 u64 policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tc_alloc_pair(u32* slot, u32 ncols) {  // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc_pair(u32 taddr, u32 ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows in each CTA] * B[smem, N/2 rows in each CTA]^T; leader CTA, one thread
__device__ __forceinline__ void tc_mma_tf32_pair(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with bf16 operands (kind::f16, K = 16 per instruction)
__device__ __forceinline__ void tc_mma_bf16_pair(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in both CTAs once the previously issued MMAs have completed
__device__ __forceinline__ void tc_commit_pair(u64* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<unsigned short>(3))
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(u64* bar, u32 rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait with cluster-scope acquire (the arrivals come from the peer CTA) and a watchdog
__device__ __forceinline__ void mbar_wait_cluster_tag(u64* bar, u32 parity, int tag) {
  const u32 addr = smem_u32(bar);
  for (u32 spins = 0;; ++spins) {
    u32 done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins == (1u << 26)) {
      printf("b2r flat_gemm pair: barrier wait stuck: tag %d block %d thread %d parity %u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
// instruction descriptor: D fp32, A/B bf16, both K-major
__host__ __device__ constexpr u32 make_idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<u32>(n >> 3) << 17) | (static_cast<u32>(m >> 4) << 24);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M (128 or 256), N
__host__ __device__ constexpr u32 make_idesc_tf32_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<u32>(n >> 3) << 17) | (static_cast<u32>(m >> 4) << 24);
}

#endif  // __CUDACC__

// host: 2-D tensor map over a row-major float32 matrix [outer, inner], {box_inner x box_outer} boxes, SWIZZLE_128B
int make_map_2d(CUtensorMap* map, const float* base, unsigned long long inner, unsigned long long outer,
                unsigned box_inner, unsigned box_outer);
// same over a bfloat16 matrix (box_inner = 64 elements = one 128-byte swizzle row)
int make_map_2d_bf16(CUtensorMap* map, const void* base, unsigned long long inner, unsigned long long outer,
                     unsigned box_inner, unsigned box_outer);

}  // namespace b2r
