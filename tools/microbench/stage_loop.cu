// stage_loop.cu -- what does each per-stage operation of the render kernel's MMA-issue loop cost,
// given the tensor pipe's shallow instruction queue?  One warp issues `n_stage` stages of 12
// tcgen05.mma (N=128, TS) and optionally: (1) a try_wait on an already-complete mbarrier in the
// middle of the stage, (2) tcgen05.fence::after_thread_sync after it, (4) elect/__syncwarp splits,
// (8) a tcgen05.commit per stage, (16) ~N cycles of dummy ALU work mid-stage.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(448, 1) k_stage(int n_stage, int flags, int dummy, int noise, long long* out, int* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar_done, bar_ready, bar_commit;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5;
  const uint32_t smb = smem_u32(sm);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_ready), 1); mbar_init(smem_u32(&bar_commit), 1u << 20);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive(smem_u32(&bar_ready));     // phase 0 of bar_ready is complete: wait(parity 0) returns at once
    stop_flag = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = umma_idesc(1, 128, 128);
  if (warp == 13) {     // highest warp id, like the render kernel
    int acc = 0;
    const long long t0 = clock64();
    for (int s = 0; s < n_stage; ++s) {
      const uint32_t wb = smb + (uint32_t)(s & 3) * 32768u;
      for (int kk = 0; kk < 2; ++kk) {
        if (elect_one()) {
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int k = 2 * kk + k2;
            const uint64_t bh = umma_desc(wb + 32 * k), bl = umma_desc(wb + 16384 + 32 * k);
            const uint32_t ah = 256u + 8u * k;
            mma_ts(128u * (s & 1), ah, bh, idesc, 1u);
            mma_ts(128u * (s & 1), ah + 128u, bh, idesc, 1u);
            mma_ts(128u * (s & 1), ah, bl, idesc, 1u);
          }
          if (kk == 1 && (flags & 8)) tc_commit(smem_u32(&bar_commit));
        }
        if (flags & 4) __syncwarp();
        if (kk == 0) {
          if (flags & 1) mbar_wait(smem_u32(&bar_ready), 0);
          if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (flags & 16) { for (int i = 0; i < dummy; ++i) acc = acc * 3 + i; }
        }
      }
    }
    if (elect_one()) tc_commit(smem_u32(&bar_done));
    __syncwarp();
    mbar_wait(smem_u32(&bar_done), 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x] = t1 - t0; sink[blockIdx.x] = acc; }
    stop_flag = 1;
  } else if (warp < 12 && noise) {
    // noise warps: 1 = FMA chains, 2 = TMEM ld/st on accumulator columns 256.. (not used by the MMAs' D),
    // 3 = shared-memory loads, 4 = mbarrier try_wait spinning on a never-completing barrier
    float x = threadIdx.x * 1e-3f, y = 1.0001f;
    uint32_t r[32];
    const uint32_t tl = ((uint32_t)(32 * (warp & 3)) << 16) + 384u;
    int guard = 0;
    while (!stop_flag && guard < (1 << 22)) {
      ++guard;
      if (noise == 1) {
#pragma unroll
        for (int i = 0; i < 64; ++i) x = fmaf(x, y, 0.5f);
      } else if (noise == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(tl) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(tl + 64u), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else if (noise == 3) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x += reinterpret_cast<volatile float*>(sm)[(threadIdx.x * 4 + i * 512) & 8191];
      } else if (noise == 4) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar_commit)), "r"(0) : "memory");
        x += ok;
      }
    }
    if (x == 123.456f) sink[1000] = (int)x + r[0];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0), "r"(512));
}

int main() {
  long long* d_out; int* d_sink;
  cudaMalloc(&d_out, 256 * sizeof(long long)); cudaMalloc(&d_sink, 2048 * sizeof(int));
  const int smem = 131072 + 1024, n_stage = 720;
  cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct { int flags, dummy; const char* name; int noise; } cases[] = {
      {0, 0, "12 MMA / stage only"}, {4, 0, "+ syncwarp"}, {8, 0, "+ commit per stage"}, {1, 0, "+ try_wait(ready) mid-stage"},
      {3, 0, "+ try_wait + fence::after"}, {15, 0, "all (render kernel's loop)"},
      {16, 16, "+ 16 dummy iters mid-stage"}, {16, 32, "+ 32 dummy iters"}, {16, 64, "+ 64 dummy iters"}, {16, 128, "+ 128 dummy iters"},
      {16, 256, "+ 256 dummy iters"},
      {15, 0, "render loop + 12 FMA-chain warps", 1}, {15, 0, "render loop + 12 TMEM ld/st warps", 2},
      {15, 0, "render loop + 12 LDS warps", 3}, {15, 0, "render loop + 12 try_wait spinners", 4}};
  for (auto& c : cases) {
    for (int rep = 0; rep < 2; ++rep) k_stage<<<148, 448, smem>>>(n_stage, c.flags, c.dummy, c.noise, d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    printf("%-34s cycles/stage %.1f (floor 768)\n", c.name, (double)mx / n_stage);
  }
  return 0;
}
