// mma_rate.cu -- cycles per tcgen05.mma (kind::f16, M=128, cta_group::1) for several shapes /
// operand sources on sm_100a.  One elected thread per CTA issues `iters` chains of MMAs into
// TMEM accumulators; data are garbage (we only time).  Usage: mma_rate [n_ctas]
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// mode: 0 TS, 1 SS.   pattern: 0 = same A and B every MMA; 1 = the render kernel's 3-pass pattern
// (A_hi,B_hi),(A_lo,B_hi),(A_hi,B_lo) walking k16 steps within a 64-wide stage and over 4 stages.
template <int N, int MODE, int PATTERN>
__global__ void __launch_bounds__(128, 1) k_rate(int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t smb = smem_u32(sm);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = umma_idesc(1, 128, N);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n_mma; i += 12) {
        const uint32_t stage = (uint32_t)((i / 12) & 3) * 32768u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t koff = PATTERN ? 32u * k : 0u;
          const uint64_t bh = umma_desc(smb + (PATTERN ? stage : 0u) + koff);
          const uint64_t bl = umma_desc(smb + (PATTERN ? stage : 0u) + 16384u + koff);
          const uint32_t ah = 256u + (PATTERN ? 8u * k : 0u), al = ah + 128u;
          const uint64_t eh = umma_desc(smb + 131072u + koff), el = umma_desc(smb + 131072u + 16384u + koff);
          if (MODE == 0) {
            mma_ts(0, ah, bh, idesc, 1u);
            mma_ts(0, PATTERN ? al : ah, bh, idesc, 1u);
            mma_ts(0, ah, PATTERN ? bl : bh, idesc, 1u);
          } else {
            mma_ss(0, eh, bh, idesc, 1u);
            mma_ss(0, PATTERN ? el : eh, bh, idesc, 1u);
            mma_ss(0, eh, PATTERN ? bl : bh, idesc, 1u);
          }
        }
      }
      tc_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0), "r"(512));
}

template <int N, int MODE, int PATTERN>
void run(const char* name, int n_ctas, long long* d_out) {
  const int n_mma = 12 * 400;
  const int smem = 131072 + 32768 + 1024;
  cudaFuncSetAttribute(k_rate<N, MODE, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) k_rate<N, MODE, PATTERN><<<n_ctas, 128, smem>>>(n_mma, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[256];
  cudaMemcpy(h, d_out, sizeof(long long) * n_ctas, cudaMemcpyDeviceToHost);
  double mx = 0, mn = 1e30;
  for (int i = 0; i < n_ctas; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("%-34s ctas=%3d cycles/MMA min %.1f max %.1f  (floor %d)  MAC/clk/SM %.0f\n", name, n_ctas, mn / n_mma, mx / n_mma,
         N / 2, 128.0 * N * 16 / (mx / n_mma));
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 148;
  long long* d_out;
  cudaMalloc(&d_out, 256 * sizeof(long long));
  for (int ctas : {1, n}) {
    run<128, 0, 1>("TS N=128 kernel pattern", ctas, d_out);
    run<128, 0, 0>("TS N=128 same operands", ctas, d_out);
    run<256, 0, 0>("TS N=256 same operands", ctas, d_out);
    run<256, 0, 1>("TS N=256 kernel pattern(B 256 rows)", ctas, d_out);
    run<64, 0, 0>("TS N=64 same operands", ctas, d_out);
    run<128, 1, 1>("SS N=128 kernel pattern", ctas, d_out);
    run<128, 1, 0>("SS N=128 same operands", ctas, d_out);
    run<256, 1, 0>("SS N=256 same operands", ctas, d_out);
  }
  return 0;
}
