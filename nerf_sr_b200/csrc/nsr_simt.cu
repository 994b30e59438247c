// nsr_simt.cu -- fp32 CUDA-core path of the MLP (NSR_PREC_FP32_SIMT).
//
// Generic over D / W / skips / no_dir / deg_* (W in {64,128,256}); it is the
// "any option combination" path and the on-GPU fp32 cross-check for the tcgen05
// kernel.  One CTA owns a tile of 64 points: positional encodings and every
// activation stay in shared memory (k-major [feature][point]), weights stream
// from L2 through a cp.async double buffer, each thread owns an 8-point x JN-
// output register tile.  Only (rgb, sigma) per point is written to HBM.
//
// Replaces (reference): PositionalEncoding.__call__ (models/embedding.py:44-63),
// the [P,90] concat of render_rays (models/nerf_downX_model.py:264-267) and
// VanillaMLP.forward (models/networks.py:199-224).
#include "nsr_internal.h"

namespace nsr {

constexpr int kTilePts = 64;
constexpr int kLd = 68;              // smem row stride (floats): 64 points + 4 pad -> conflict-free STS.128
constexpr int kSlabK = 16;
constexpr int kThreads = 256;

// ---- weight blob ----------------------------------------------------------
// tiled layer: W^T, rows padded to a multiple of 16, columns permuted so that
// thread tx's JN outputs (n = tx + 32 j) are contiguous: col = tx*JN + j.
static inline int64_t pad16(int64_t k) { return (k + 15) / 16 * 16; }

static void build_program(NsrHandle_* h) {
  const NsrConfig& c = h->cfg;
  SimtProgram& P = h->prog;
  const RenderParams& rp = h->rp;
  P.W = c.W; P.ch_pos = rp.ch_pos; P.ch_dir = rp.ch_dir;
  int64_t off = 0;
  int cur = 0;        // buffer holding the current activation (0 = enc_xyz for layer 0)
  int nl = 0;
  for (int i = 0; i < c.D; ++i) {
    SimtLayer L{};
    const bool skip = (c.skips_mask >> i) & 1u;
    if (i == 0) { L.K0 = rp.ch_pos; L.K1 = 0; L.src0 = 0; L.src1 = 0; }
    else if (skip) { L.K0 = rp.ch_pos; L.K1 = c.W; L.src0 = 0; L.src1 = cur; }   // cat([input_xyz, h]) networks.py:204
    else { L.K0 = c.W; L.K1 = 0; L.src0 = cur; L.src1 = cur; }
    L.dst = (cur == 2) ? 3 : 2;
    L.N = c.W; L.relu = 1;
    L.w_off = off; off += pad16(L.K0 + L.K1) * L.N;
    L.b_off = off; off += L.N;
    P.layers[nl++] = L;
    cur = L.dst;
  }
  P.sigma_src = cur;
  {  // xyz_encoding_final: Linear, no activation (networks.py:158,211)
    SimtLayer L{};
    L.K0 = c.W; L.K1 = 0; L.src0 = cur; L.src1 = cur; L.dst = (cur == 2) ? 3 : 2;
    L.N = c.W; L.relu = 0;
    L.w_off = off; off += pad16(L.K0) * L.N; L.b_off = off; off += L.N;
    P.layers[nl++] = L;
    // NOTE: sigma head reads h_D from `cur`, which the final layer does not overwrite
    // (it writes the other buffer), but the dir layer below will: heads run in order.
    cur = L.dst;
  }
  {  // dir_encoding: cat([feat, enc_dir]) -> W/2, ReLU (networks.py:213-221)
    SimtLayer L{};
    L.K0 = c.W; L.K1 = c.no_dir ? 0 : rp.ch_dir; L.src0 = cur; L.src1 = 1;
    L.dst = (cur == 2) ? 3 : 2;
    L.N = c.W / 2; L.relu = 1;
    L.w_off = off; off += pad16(L.K0 + L.K1) * L.N; L.b_off = off; off += L.N;
    P.layers[nl++] = L;
    cur = L.dst;
  }
  P.rgb_src = cur;
  P.n_layers = nl;
  P.w_sigma = off; off += c.W;
  P.b_sigma = off; off += 1;
  P.w_rgb = off; off += 3 * (c.W / 2);
  P.b_rgb = off; off += 3;
  h->net[0].simt_floats = h->net[1].simt_floats = (size_t)off;
}

size_t simt_blob_floats(const NsrHandle_* h) {
  if (h->net[0].simt_floats == 0) build_program(const_cast<NsrHandle_*>(h));
  return h->net[0].simt_floats;
}

__global__ void k_pack_tiled(const float* __restrict__ W, int N, int K, int Kpad, float* __restrict__ dst) {
  // W: [N][K] row-major (nn.Linear.weight) -> dst[k][tx*JN + j], n = tx + 32 j
  const int JN = N / 32;
  const int64_t total = (int64_t)Kpad * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / N), col = (int)(i % N);
    const int tx = col / JN, j = col % JN;
    const int n = tx + 32 * j;
    dst[i] = (k < K) ? W[(int64_t)n * K + k] : 0.f;
  }
}

__global__ void k_copy(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

cudaError_t simt_pack(NsrHandle_* h, int which, const float* const* params, cudaStream_t st) {
  const SimtProgram& P = h->prog;
  NetImages& net = h->net[which];
  float* blob = net.simt_blob;
  // state_dict order: trunk (w,b) x D, final (w,b), dir (w,b), sigma (w,b), rgb (w,b)
  for (int l = 0; l < P.n_layers; ++l) {
    const SimtLayer& L = P.layers[l];
    const int K = L.K0 + L.K1;
    const int Kp = (int)pad16(K);
    const int64_t total = (int64_t)Kp * L.N;
    k_pack_tiled<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(params[2 * l], L.N, K, Kp, blob + L.w_off);
    k_copy<<<1, 256, 0, st>>>(params[2 * l + 1], blob + L.b_off, L.N);
    h->launches += 2;
  }
  const int base = 2 * P.n_layers;
  k_copy<<<1, 256, 0, st>>>(params[base + 0], blob + P.w_sigma, P.W);
  k_copy<<<1, 32, 0, st>>>(params[base + 1], blob + P.b_sigma, 1);
  k_copy<<<1, 256, 0, st>>>(params[base + 2], blob + P.w_rgb, 3 * (P.W / 2));
  k_copy<<<1, 32, 0, st>>>(params[base + 3], blob + P.b_rgb, 3);
  h->launches += 4;
  return cudaGetLastError();
}

// ---- the kernel -------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SimtSmem {
  float enc[64 * kLd];
  float dir[32 * kLd];
  float buf[2][256 * kLd];
  float slab[2][kSlabK * 256];
  float red[4 * 4 * 64];
};

template <int JN>
__device__ __forceinline__ void gemm_layer(const SimtLayer& L, SimtSmem& sm, const float* __restrict__ blob) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  constexpr int N = 32 * JN;
  const int Ktot = L.K0 + L.K1;
  const int nslab = (Ktot + kSlabK - 1) / kSlabK;
  const float* wsrc = blob + L.w_off;
  auto bufptr = [&](int id) -> const float* {
    return id == 0 ? sm.enc : (id == 1 ? sm.dir : sm.buf[id - 2]);
  };
  const float* s0 = bufptr(L.src0);
  const float* s1 = bufptr(L.src1);
  float acc[8][JN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < JN; ++j) acc[i][j] = 0.f;

  auto prefetch = [&](int s) {
    const float* g = wsrc + (int64_t)s * kSlabK * N;
    float* d = sm.slab[s & 1];
    for (int i = tid; i < kSlabK * N / 4; i += kThreads) cp_async16(d + 4 * i, g + 4 * i);
    cp_async_commit();
  };
  prefetch(0);
  for (int s = 0; s < nslab; ++s) {
    if (s + 1 < nslab) { prefetch(s + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* wb = sm.slab[s & 1] + tx * JN;
    const int k0 = s * kSlabK;
    const int kmax = min(kSlabK, Ktot - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      const int k = k0 + kk;
      const float* row = (k < L.K0) ? (s0 + k * kLd) : (s1 + (k - L.K0) * kLd);
      const float4 a0 = *reinterpret_cast<const float4*>(row + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(row + ty * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[JN];
      if constexpr (JN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < JN; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(wb + kk * N + j);
          b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < JN; ++j) b[j] = wb[kk * N + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < JN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = sm.buf[L.dst - 2];
  const float* bias = blob + L.b_off;
#pragma unroll
  for (int j = 0; j < JN; ++j) {
    const int n = tx + 32 * j;
    const float bv = __ldg(bias + n);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = acc[i][j] + bv;
      if (L.relu) v[i] = fmaxf(v[i], 0.f);
    }
    *reinterpret_cast<float4*>(dst + n * kLd + ty * 8) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + n * kLd + ty * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
}

__device__ __forceinline__ void run_layer(const SimtLayer& L, SimtSmem& sm, const float* blob) {
  switch (L.N / 32) {
    case 8: gemm_layer<8>(L, sm, blob); break;
    case 4: gemm_layer<4>(L, sm, blob); break;
    case 2: gemm_layer<2>(L, sm, blob); break;
    default: gemm_layer<1>(L, sm, blob); break;
  }
}

// dot of one point's column of `src` ([K][kLd]) with w[K], split over 4 thread groups
__device__ __forceinline__ float head_partial(const float* src, const float* __restrict__ w, int K, int p, int part) {
  const int k0 = part * K / 4, k1 = (part + 1) * K / 4;
  float s = 0.f;
  for (int k = k0; k < k1; ++k) s = fmaf(src[k * kLd + p], __ldg(w + k), s);
  return s;
}

__global__ void __launch_bounds__(kThreads, 1)
k_simt_mlp(SimtProgram P, RenderParams rp, const SampleTables* __restrict__ tabs,
           const float* __restrict__ blob, const float* __restrict__ rays, int64_t n_rays,
           int ray_stride, const float* __restrict__ z, int S, float* __restrict__ raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SimtSmem& sm = *reinterpret_cast<SimtSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int64_t n_pts = n_rays * S;
  const int64_t n_tiles = (n_pts + kTilePts - 1) / kTilePts;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- sample point + positional encodings (k-major) ----
    {
      const int p = tid & 63, part = tid >> 6;
      const int64_t gp = tile * kTilePts + p;
      float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, vx = 0, vy = 0, vz = 0, zz = 0;
      if (gp < n_pts) {
        const int64_t r = gp / S;
        const float* rr = rays + r * ray_stride;
        ox = rr[0]; oy = rr[1]; oz = rr[2]; dx = rr[3]; dy = rr[4]; dz = rr[5];
        vx = rr[rp.viewdir_offset]; vy = rr[rp.viewdir_offset + 1]; vz = rr[rp.viewdir_offset + 2];
        zz = z[gp];
      }
      const float px = cast_point(ox, dx, zz), py = cast_point(oy, dy, zz), pz = cast_point(oz, dz, zz);
      // each of the 4 thread groups takes a share of the frequency bands
      if (part == 0 && !rp.no_xyz) {
        sm.enc[0 * kLd + p] = px; sm.enc[1 * kLd + p] = py; sm.enc[2 * kLd + p] = pz;
        sm.dir[0 * kLd + p] = vx; sm.dir[1 * kLd + p] = vy; sm.dir[2 * kLd + p] = vz;
      }
      const int base = rp.no_xyz ? 0 : 3;
      for (int k = part; k < rp.deg_pos; k += 4) {
        const float f = tabs->freq_pos[k];
        const float ax = __fmul_rn(f, px), ay = __fmul_rn(f, py), az = __fmul_rn(f, pz);
        const int c = base + 6 * k;
        sm.enc[(c + 0) * kLd + p] = sinf(ax); sm.enc[(c + 1) * kLd + p] = sinf(ay); sm.enc[(c + 2) * kLd + p] = sinf(az);
        sm.enc[(c + 3) * kLd + p] = cosf(ax); sm.enc[(c + 4) * kLd + p] = cosf(ay); sm.enc[(c + 5) * kLd + p] = cosf(az);
      }
      for (int k = part; k < rp.deg_dir; k += 4) {
        const float f = tabs->freq_dir[k];
        const float ax = __fmul_rn(f, vx), ay = __fmul_rn(f, vy), az = __fmul_rn(f, vz);
        const int c = base + 6 * k;
        sm.dir[(c + 0) * kLd + p] = sinf(ax); sm.dir[(c + 1) * kLd + p] = sinf(ay); sm.dir[(c + 2) * kLd + p] = sinf(az);
        sm.dir[(c + 3) * kLd + p] = cosf(ax); sm.dir[(c + 4) * kLd + p] = cosf(ay); sm.dir[(c + 5) * kLd + p] = cosf(az);
      }
    }
    __syncthreads();
    // ---- trunk, final, dir layers; sigma head right after the trunk ----
    const int p = tid & 63, part = tid >> 6;
    for (int l = 0; l < P.n_layers; ++l) {
      if (l == P.n_layers - 2) {   // h_D is complete: sigma head (networks.py:207)
        sm.red[(0 * 4 + part) * 64 + p] = head_partial(sm.buf[P.sigma_src - 2], blob + P.w_sigma, P.W, p, part);
        // visibility: the final layer's barriers order this before the read below
      }
      run_layer(P.layers[l], sm, blob);
    }
    {  // rgb head (networks.py:222)
      const float* src = sm.buf[P.rgb_src - 2];
      const int K = P.W / 2;
      for (int c = 0; c < 3; ++c)
        sm.red[((1 + c) * 4 + part) * 64 + p] = head_partial(src, blob + P.w_rgb + c * K, K, p, part);
    }
    __syncthreads();
    if (tid < 64) {
      const int64_t gp = tile * kTilePts + tid;
      if (gp < n_pts) {
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* r4 = sm.red + c * 4 * 64 + tid;
          o[c] = ((r4[0] + r4[64]) + r4[128]) + r4[192];
        }
        const float sigma = o[0] + __ldg(blob + P.b_sigma);
        float col[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = o[1 + c] + __ldg(blob + P.b_rgb + c);
          if (!rp.color_none) v = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));   // torch.sigmoid
          col[c] = v;
        }
        reinterpret_cast<float4*>(raw)[gp] = make_float4(col[0], col[1], col[2], sigma);
      }
    }
    __syncthreads();
  }
}

cudaError_t simt_mlp(NsrHandle_* h, int which, const float* rays, int64_t n_rays, int ray_stride,
                     const float* z, int S, float* raw, cudaStream_t st) {
  static_assert(sizeof(SimtSmem) <= 227 * 1024, "smem budget");
  const size_t smem = sizeof(SimtSmem);
  cudaError_t e = cudaFuncSetAttribute(k_simt_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t n_tiles = (n_rays * S + kTilePts - 1) / kTilePts;
  if (n_tiles == 0) return cudaSuccess;
  const int grid = (int)(n_tiles < h->sm_count ? n_tiles : h->sm_count);
  k_simt_mlp<<<grid, kThreads, smem, st>>>(h->prog, h->rp, h->d_tables, h->net[which].simt_blob, rays,
                                           n_rays, ray_stride, z, S, raw);
  h->launches += 1;
  return cudaGetLastError();
}

}  // namespace nsr
