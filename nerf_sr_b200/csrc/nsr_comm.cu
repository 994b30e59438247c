// nsr_comm.cu -- the data-parallel gradient all-reduce of the training iteration as ONE kernel over peer-mapped
// memory (NVLink 5 / NVSwitch P2P loads and stores), sm_100a.
//
// Replaces (reference): the bucketed NCCL all-reduce DistributedDataParallel issues from its autograd hooks while
// loss_tot.backward() runs (models/networks.py:72-86; the reference wraps netCoarse / netFine in DDP and the DDP
// loader gives every rank batch_size / n_gpus rays, data/__init__.py:94-99), including DDP's division by the world size.
//
// Design.  The gradient of both nets is one flat fp32 bucket (2 x 595 844 floats = 4.77 MB) that lives in a
// "symmetric" buffer: every rank cudaMalloc's the same layout, exports it with CUDA IPC, and maps all peers' buffers
// (nsr_comm_export / nsr_comm_connect_ipc; torch.distributed only carries the 64-byte handles).  nsr_backward's
// reduction kernel writes the bucket in place, so the all-reduce is a single launch with no staging copy:
//
//   barrier A   CTA b of rank r tells CTA b of every peer "my kernel runs, hence my bucket is complete" (one
//               st.release.sys per peer into the peer's flag array) and waits for the same word from each peer
//   two-shot    rank r owns slice r of the bucket: it LOADS that slice from every rank over NVLink (peer loads hit the
//               owner's L2), sums in fixed rank order 0..N-1, scales by 1/N (DDP's mean) and STORES the result into
//               slice r of every rank's bucket -- reduce-scatter and all-gather fused, in place
//   barrier B   after __threadfence_system(), CTA b signals every peer "my slice has landed in your bucket" and waits
//               for the peers' signals; when the kernel ends, the local bucket holds the mean gradient
//
// Every element is reduced by exactly one rank, in one fixed order, so all ranks end with BIT-IDENTICAL gradients
// (stronger than NCCL's guarantee) and the replicas cannot drift.  Per rank and call: (N-1)/N x 4.77 MB in and out over
// NVLink; two flag round trips (~2 us each).  Flags are monotonically increasing epochs (never reset), compared with >=.
//
// The kernel spins on peers, so all of its CTAs must be co-resident: the grid is kCommCtas (<< 148 SMs).
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "nsr_internal.h"

namespace nsr {

constexpr int kCommCtas = 32;        // x 512 threads, <= 64 registers: 2 CTAs per SM can be resident (8 single-GPU test ranks fit)
constexpr int kCommThreads = 512;
constexpr int kMaxWorld = 16;

struct CommKernelArgs {
  float4* peer[kMaxWorld];       // every rank's bucket (peer[rank] is the local one)
  uint32_t* flags[kMaxWorld];    // every rank's flag array: [2 phases][kCommCtas][kMaxWorld]
  int rank, world;
  long long n_vec;               // float4 elements in the bucket (zero padded)
  uint32_t epoch;
  float inv_world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {     // never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// phase 0 = barrier A, 1 = barrier B.  Thread p < world talks to peer p.
__device__ __forceinline__ void cross_gpu_barrier(const CommKernelArgs& a, int phase) {
  const int p = threadIdx.x;
  if (p < a.world) {
    const size_t slot = ((size_t)phase * kCommCtas + blockIdx.x) * kMaxWorld;
    st_release_sys(a.flags[p] + slot + a.rank, a.epoch);                  // into peer p's array, my column
    const uint32_t* mine = a.flags[a.rank] + slot + p;                    // my array, peer p's column
    unsigned long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) {
      if (++spins > (1ull << 26)) {                                       // ~a minute: a peer died; fail the launch, do not hang
        printf("[nsr_comm] rank %d CTA %d: peer %d never reached epoch %u (phase %d)\n", a.rank, (int)blockIdx.x, p, a.epoch, phase);
        __trap();
      }
    }
  }
  __syncthreads();
}

// W = compile-time world size (0: runtime, any world <= kMaxWorld), U = float4 elements per thread and trip: W x U loads
// are in flight before the first add (peer loads cost ~2 us of NVLink latency; bandwidth needs many of them outstanding).
template <int W, int U>
__global__ void __launch_bounds__(kCommThreads, 2) k_allreduce_mean(const CommKernelArgs a) {
  cross_gpu_barrier(a, 0);
  const int world = W ? W : a.world;
  const long long lo = a.n_vec * a.rank / world, hi = a.n_vec * (a.rank + 1) / world;
  const long long stride = (long long)kCommCtas * kCommThreads;
  for (long long i0 = lo + (long long)blockIdx.x * kCommThreads + threadIdx.x; i0 < hi; i0 += stride * U) {
    float4 s[U];
    if (W) {
      float4 v[U][W ? W : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
#pragma unroll
        for (int w = 0; w < (W ? W : 1); ++w)
          if (i < hi) v[u][w] = ld_peer(a.peer[w] + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s[u] = v[u][0];
#pragma unroll
        for (int w = 1; w < (W ? W : 1); ++w) {                           // fixed rank order 0, 1, ..., W-1
          s[u].x = __fadd_rn(s[u].x, v[u][w].x); s[u].y = __fadd_rn(s[u].y, v[u][w].y);
          s[u].z = __fadd_rn(s[u].z, v[u][w].z); s[u].w = __fadd_rn(s[u].w, v[u][w].w);
        }
      }
    } else {
      s[0] = ld_peer(a.peer[0] + i0);
      for (int w0 = 1; w0 < world; w0 += 8) {                             // batches of 8 loads, same order
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (w0 + k < world) v[k] = ld_peer(a.peer[w0 + k] + i0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (w0 + k < world) {
            s[0].x = __fadd_rn(s[0].x, v[k].x); s[0].y = __fadd_rn(s[0].y, v[k].y);
            s[0].z = __fadd_rn(s[0].z, v[k].z); s[0].w = __fadd_rn(s[0].w, v[k].w);
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi) continue;
      float4 r = s[u];
      r.x = __fmul_rn(r.x, a.inv_world); r.y = __fmul_rn(r.y, a.inv_world); r.z = __fmul_rn(r.z, a.inv_world); r.w = __fmul_rn(r.w, a.inv_world);
      for (int w = 0; w < world; ++w) st_peer(a.peer[w] + i, r);
    }
  }
  __threadfence_system();
  __syncthreads();
  cross_gpu_barrier(a, 1);
}

}  // namespace nsr

using namespace nsr;

struct NsrComm_ {
  NsrHandle_* h = nullptr;
  int rank = 0, world = 1;
  size_t n_floats = 0;
  long long n_vec = 0;
  size_t data_bytes = 0, total_bytes = 0;
  char* base = nullptr;                  // local allocation: [bucket | flags]
  char* peer[kMaxWorld] = {};
  bool ipc_opened[kMaxWorld] = {};
  bool connected = false;
  uint32_t epoch = 0;
};

static int cfail(NsrHandle_* h, int code, const std::string& msg) { if (h) h->err = msg; return code; }
#define NSR_CCUDA(h, expr)                                                                    \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return cfail(h, NSR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
  } while (0)

extern "C" int nsr_comm_create(NsrHandle* h, int rank, int world, int64_t n_floats, NsrComm** out) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n_floats <= 0)
    return cfail(h, NSR_ERR_INVALID_ARG, "nsr_comm_create: bad argument (world must be in [1,16])");
  *out = nullptr;
  NSR_CCUDA(h, cudaSetDevice(h->cfg.device));
  NsrComm_* c = new (std::nothrow) NsrComm_();
  if (!c) return cfail(h, NSR_ERR_CUDA, "out of host memory");
  c->h = h; c->rank = rank; c->world = world; c->n_floats = (size_t)n_floats;
  c->n_vec = (long long)((n_floats + 4095) / 4096 * 1024);                 // zero-padded to a multiple of 4096 floats
  c->data_bytes = (size_t)c->n_vec * sizeof(float4);
  const size_t flag_bytes = (size_t)2 * kCommCtas * kMaxWorld * sizeof(uint32_t);
  c->total_bytes = c->data_bytes + flag_bytes;
  cudaError_t e = cudaMalloc(&c->base, c->total_bytes);                    // plain cudaMalloc: exportable with CUDA IPC
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, c->total_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(c->base); delete c; return cfail(h, NSR_ERR_CUDA, std::string("nsr_comm_create: ") + cudaGetErrorString(e)); }
  c->peer[rank] = c->base;
  c->connected = (world == 1);
  *out = c;
  return NSR_OK;
}

extern "C" int nsr_comm_destroy(NsrComm* c) {
  if (!c) return NSR_OK;
  cudaSetDevice(c->h->cfg.device);
  cudaDeviceSynchronize();
  for (int w = 0; w < c->world; ++w)
    if (c->ipc_opened[w]) cudaIpcCloseMemHandle(c->peer[w]);
  cudaFree(c->base);
  delete c;
  return NSR_OK;
}

extern "C" float* nsr_comm_buffer(NsrComm* c) { return c ? reinterpret_cast<float*>(c->base) : nullptr; }
extern "C" int64_t nsr_comm_buffer_floats(const NsrComm* c) { return c ? (int64_t)c->n_vec * 4 : 0; }

extern "C" int nsr_comm_export(NsrComm* c, void* handle_out64) {
  if (!c || !handle_out64) return NSR_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  NSR_CCUDA(c->h, cudaSetDevice(c->h->cfg.device));
  cudaIpcMemHandle_t hd;
  NSR_CCUDA(c->h, cudaIpcGetMemHandle(&hd, c->base));
  memcpy(handle_out64, &hd, 64);
  return NSR_OK;
}

extern "C" int nsr_comm_connect_ipc(NsrComm* c, const void* handles64, int n_handles) {
  if (!c || !handles64) return NSR_ERR_INVALID_ARG;
  if (n_handles != c->world) return cfail(c->h, NSR_ERR_INVALID_ARG, "nsr_comm_connect_ipc: need one handle per rank");
  NSR_CCUDA(c->h, cudaSetDevice(c->h->cfg.device));
  for (int w = 0; w < c->world; ++w) {
    if (w == c->rank || c->ipc_opened[w]) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)handles64 + 64 * (size_t)w, 64);
    void* p = nullptr;
    NSR_CCUDA(c->h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c->peer[w] = (char*)p;
    c->ipc_opened[w] = true;
  }
  c->connected = true;
  return NSR_OK;
}

extern "C" int nsr_comm_connect_ptrs(NsrComm* c, void* const* peer_buffers, int n) {
  if (!c || !peer_buffers) return NSR_ERR_INVALID_ARG;
  if (n != c->world) return cfail(c->h, NSR_ERR_INVALID_ARG, "nsr_comm_connect_ptrs: need one pointer per rank");
  for (int w = 0; w < c->world; ++w) {
    if (w == c->rank) continue;
    if (!peer_buffers[w]) return cfail(c->h, NSR_ERR_INVALID_ARG, "nsr_comm_connect_ptrs: null peer buffer");
    c->peer[w] = (char*)peer_buffers[w];
  }
  c->connected = true;
  return NSR_OK;
}

extern "C" int nsr_comm_allreduce_mean(NsrComm* c, NsrStream stream) {
  if (!c) return NSR_ERR_INVALID_ARG;
  if (!c->connected) return cfail(c->h, NSR_ERR_INVALID_ARG, "nsr_comm_allreduce_mean: peers not connected");
  if (c->world == 1) return NSR_OK;
  NSR_CCUDA(c->h, cudaSetDevice(c->h->cfg.device));
  CommKernelArgs a{};
  for (int w = 0; w < c->world; ++w) {
    a.peer[w] = reinterpret_cast<float4*>(c->peer[w]);
    a.flags[w] = reinterpret_cast<uint32_t*>(c->peer[w] + c->data_bytes);
  }
  a.rank = c->rank; a.world = c->world; a.n_vec = c->n_vec;
  a.epoch = ++c->epoch;
  a.inv_world = 1.0f / (float)c->world;
  cudaStream_t st = (cudaStream_t)stream;
  switch (c->world) {
    case 2: k_allreduce_mean<2, 4><<<kCommCtas, kCommThreads, 0, st>>>(a); break;
    case 4: k_allreduce_mean<4, 2><<<kCommCtas, kCommThreads, 0, st>>>(a); break;
    case 8: k_allreduce_mean<8, 1><<<kCommCtas, kCommThreads, 0, st>>>(a); break;
    default: k_allreduce_mean<0, 1><<<kCommCtas, kCommThreads, 0, st>>>(a); break;
  }
  c->h->launches += 1;
  NSR_CCUDA(c->h, cudaGetLastError());
  return NSR_OK;
}
