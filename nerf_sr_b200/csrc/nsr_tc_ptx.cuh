// nsr_tc_ptx.cuh -- inline-PTX wrappers for the sm_100a primitives shared by the tcgen05 kernels
// (nsr_tc.cu: fused render pass; nsr_train.cu: backward GEMMs): mbarrier, cp.async.bulk, tcgen05
// mma / ld / st / commit / fences, UMMA descriptors, and the 16-bit hi/lo split.
#pragma once
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace nsr {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
#ifndef NSR_TC_WATCHDOG
#define NSR_TC_WATCHDOG 1
#endif
// Blocks until the phase with the given parity has completed.  With the watchdog enabled a
// protocol deadlock becomes a trapped launch failure with a diagnostic instead of a hung GPU.
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("[nsr_tc] mbarrier wait timed out: block %d thread %d barrier-offset %u parity %u\n", (int)blockIdx.x,
         (int)threadIdx.x, bar, parity);
  __trap();
}
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
#if NSR_TC_WATCHDOG
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) mbar_timeout(bar, parity);
  }
#else
  while (!mbar_try_wait(bar, parity)) {}
#endif
}
// One inline probe (the common case in the issue loops: already complete), slow path out of line
// so the single-lane MMA / producer loops stay short.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// The same copy delivered to the same CTA-relative offsets (data and mbarrier) of every CTA in `cta_mask` of the cluster.
__device__ __forceinline__ void bulk_copy_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// commit -> one arrival on the mbarrier at the same CTA-relative offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// One lane of a converged warp (the pattern ptxas recognises for single-thread tcgen05/TMA issue:
// operands stay in uniform registers instead of a per-lane R2UR waterfall loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row
// groups 1024 B apart): start>>4 | LBO(unused)=1 | SBO=1024>>4 | version=1 | layout=2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A/B = fmt (0 f16, 1 bf16), K-major both, N, M
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TMEM_LD32(taddr, r)                                                                           \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                               \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                               \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"              \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),    \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
                 "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),          \
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),          \
                 "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])           \
               : "r"(taddr) : "memory")

#define TMEM_ST16(taddr, r)                                                                           \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                         \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                             \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),          \
                 "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),        \
                 "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

// 256-bit global store (sm_100: STG.256).  p must be 32-byte aligned.
__device__ __forceinline__ void stg256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// Store the 16-byte chunks (2*pair, 2*pair+1) of one 128-byte row of a SWIZZLE_128B tile image with ONE
// 32-byte store: chunk j lives at position j ^ (row & 7), so an even/odd pair always shares an aligned
// 32-byte sector (in swapped order when the row index is odd).  row_base: start of the row (128-B aligned).
__device__ __forceinline__ void store_chunk_pair(uint8_t* row_base, int pair, int r7, uint4 c_even, uint4 c_odd) {
  const int pos = (((2 * pair) ^ r7) & ~1) << 4;
  const bool swp = (r7 & 1) != 0;
  stg256(row_base + pos, swp ? c_odd : c_even, swp ? c_even : c_odd);
}

// ---- "tile image": how activations / gradients that the backward GEMMs consume live in HBM (nsr_train.cu) ----
// Per 128-point tile and 64-feature chunk: a hi plane (16 KB) then a lo plane (16 KB) of 16-bit values.  Inside a plane
// the unit is the 16-byte piece (8 consecutive features j of one point), ordered  [64-point half][j = 0..7][point & 63]:
//   offset(row, j) = (row >> 6) * 8192 + j * 1024 + (row & 63) * 16.
// Why: (1) a warp whose lanes are 32 consecutive points stores one piece per lane = 512 contiguous bytes = four full 128-byte
// lines per instruction (round 1's row-major swizzled image took 32 half-used sectors per instruction and made the stash /
// dZ stores, not the tensor pipe, the limiter of the forward-with-stash and of the dX chain); (2) a 64-point half of a plane
// is 8 KB contiguous, so the dW GEMMs fetch their stages with the same bulk copies as before; (3) it is a canonical
// no-swizzle UMMA operand both ways: MN-major (dW: points = K; 8-point groups 128 B apart = LBO, 8-feature groups 1 KB
// apart = SBO) and, with the two halves of a j-block placed side by side in shared memory, K-major (dX: points = M).
__host__ __device__ inline size_t img2_off(int row, int j) {
  return (size_t)(row >> 6) * 8192 + (size_t)j * 1024 + (size_t)(row & 63) * 16;
}
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- hi/lo split of two fp32 values into packed 16-bit pairs (even k in the low half) ----
template <int FMT> struct Split;
template <> struct Split<1> {   // bf16
  static __device__ __forceinline__ void apply(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - h1), "f"(v0 - h0));
  }
  static __host__ __device__ __forceinline__ void apply1(float v, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = *reinterpret_cast<const uint16_t*>(&h);
    lo = *reinterpret_cast<const uint16_t*>(&l);
  }
};
template <> struct Split<0> {   // fp16
  static __device__ __forceinline__ void apply(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - h.y), "f"(v0 - h.x));
  }
  static __host__ __device__ __forceinline__ void apply1(float v, uint16_t& hi, uint16_t& lo) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = *reinterpret_cast<const uint16_t*>(&h);
    lo = *reinterpret_cast<const uint16_t*>(&l);
  }
};

}  // namespace nsr
