// nsr_train.cu -- backward of the fused render pass + fused optimiser (scope row f-1), sm_100a only.
//
// Replaces (reference): loss_tot.backward() through forward_rays (models/nerf_downX_model.py:280-313,
// 390-396: autograd over VolumetricRenderer.forward, VanillaMLP.forward and the colour/sigma
// activations), nn.utils.clip_grad_norm_/clip_grad_value_ (:403-407) and torch.optim.Adam.step (:408,
// :201-204).  Sample positions carry no gradient (weights and rays only; coarse_weights.detach(), :302).
//
// Data flow of one net's backward (all device-resident, stream-ordered, no host sync):
//   the STASH variant of k_tc_pass (nsr_tc.cu) has saved every tile's MLP inputs and activations as
//   "tile images": per 128-point tile and 64-feature chunk, a bf16/fp16 hi plane (16 KB) and lo plane
//   (16 KB) of 128-B rows with XOR-swizzled 16-B chunks -- the exact SWIZZLE_128B shared-memory operand
//   layout, so every GEMM operand below is fetched with plain cp.async.bulk copies and no conversion.
//   The same image is a K-major operand (rows = points = M, K = features) for the dX GEMMs and an
//   MN-major operand (rows = points = K, features = M/N) for the dW GEMMs.
//
//   k_render_bwd   one warp per ray: alpha-compositing backward (reverse scan), colour / sigma
//                  activation backward, rgb-head backward -> images dHead (d sigma, d rgb_pre),
//                  dZ_dir (masked by the stashed dir activations), encdir (per-point copy of the ray's
//                  view-direction encoding), and the fp32 d sigma vector.
//   k_tg_dx        dH_{l-1} = dZ_l . W_l  (M = 128 points, N = 256, K = 64 per chunk), tcgen05.mma
//                  kind::f16 with hi/lo split operands (3 MMAs per product), fp32 accumulate in a
//                  double-buffered TMEM accumulator; epilogue: (+ d sigma x w_sigma) . ReLU mask from
//                  the stashed activation -> hi/lo split -> next dZ image.
//   k_tg_dw        dW_l = dZ_l^T . X_l over all points (K = points, split across CTAs), both operands
//                  MN-major views of the tile images; the bias gradient is one more N=16 MMA against a
//                  tile of ones; fp32 partials per CTA, summed in a fixed order by k_grad_reduce
//                  (deterministic).
//   k_adam         torch.optim.Adam single-tensor semantics on the caller's parameter tensors, with
//                  the clip coefficient / clip value folded in.
#include <cmath>
#include <string>

#include "nsr_internal.h"
#include "nsr_tc_ptx.cuh"
#include "nsr_tc_mma.cuh"

namespace nsr {

constexpr int kT = 128;               // points per tile
constexpr int kChunk = 32768;         // one 64-feature chunk of a tile image: hi plane | lo plane
constexpr int kPlane = 16384;
constexpr int kNumDxLayers = 9;       // dir, final, L8..L2
constexpr int kWtStages = 4 + 8 * 8;  // transposed-weight ring stages per tile (32 KB each), see build_wt_table
constexpr int kMaxSplit = 148;
constexpr int kDwLaunchesPerNet = 14;   // 13 used: rgb, dir, dir-enc, final+sigma, L8..L2 (7), L5-enc, L1

// K-major SWIZZLE_128B rows (the WEIGHT stages: 128-byte rows, 16-byte chunks XOR-swizzled with the row index)
__host__ __device__ inline size_t sw128_off(int row, int j) { return (size_t)row * 128 + (size_t)((j ^ (row & 7)) << 4); }
// (activation / gradient TILE IMAGES use img2_off, nsr_tc_ptx.cuh)

// ---------------------------------------------------------------------------
// transposed weight images for the dX GEMMs (B operand: rows = input feature n, K = output feature), laid out as the
// 32 KB ring STAGES of the fused chain in consumption order (nsr_tc_mma.cuh): stage (layer, half h, chunk c) = rows
// [128 h, 128 h + 128) x the 64 k-values of chunk c, hi plane 16 KB | lo plane 16 KB, K-major SWIZZLE_128B.
// Per layer the order is h0: c0 .. c(nkc-1), then h1: c0 .. c(nkc-1); layer 0 (dir) has nkc = 2, the others 4.
// ---------------------------------------------------------------------------
struct WtLayer { int param, n_out, ld, col0, stage0; };
struct WtTable { WtLayer l[kNumDxLayers]; };

static WtTable build_wt_table(int ch_dir, bool no_dir) {
  WtTable T{};
  int c = 0;
  T.l[0] = WtLayer{18, 128, no_dir ? 256 : 256 + ch_dir, 0, c}; c += 4;     // dir_encoding.0 (feat part): 2 halves x 2 chunks
  T.l[1] = WtLayer{16, 256, 256, 0, c}; c += 8;                              // xyz_encoding_final
  for (int i = 0; i < 7; ++i) {                                              // L8 .. L2
    const int L = 8 - i;
    const bool skip = (L == 5);
    T.l[2 + i] = WtLayer{2 * (L - 1), 256, skip ? 319 : 256, skip ? 63 : 0, c}; c += 8;
  }
  return T;
}

template <int FMT>
__global__ void k_pack_wt(WtTable T, const float* const* __restrict__ params, uint8_t* __restrict__ image) {
  // one thread per (stage, row r, 16-byte chunk j)
  const int total = kWtStages * 128 * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int st = idx / 1024, r = (idx / 8) % 128, j = idx % 8;
    int li = 0;
#pragma unroll
    for (int t = 1; t < kNumDxLayers; ++t) if (st >= T.l[t].stage0) li = t;
    const WtLayer L = T.l[li];
    const int nkc = L.n_out / 64;
    const int s_in = st - L.stage0, h = s_in / nkc, c = s_in % nkc;
    const int n = 128 * h + r;                                     // input feature of the forward layer = output column here
    const float* W = params[L.param];
    __align__(16) uint16_t hi[8];
    __align__(16) uint16_t lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int o = 64 * c + 8 * j + e;
      const float v = W[(int64_t)o * L.ld + L.col0 + n];
      Split<FMT>::apply1(v, hi[e], lo[e]);
    }
    uint8_t* dst = image + (size_t)st * kChunk + sw128_off(r, j);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + kPlane) = *reinterpret_cast<const uint4*>(lo);
  }
}

size_t train_wt_bytes() { return (size_t)kWtStages * kChunk; }

cudaError_t train_pack_wt(NsrHandle_* h, int which, const float* const* params_dev, cudaStream_t st) {
  const WtTable T = build_wt_table(h->rp.ch_dir, h->cfg.no_dir != 0);
  if (h->cfg.precision == NSR_PREC_FP16X3_TC) k_pack_wt<0><<<136, 256, 0, st>>>(T, params_dev, h->net[which].wt_image);
  else k_pack_wt<1><<<136, 256, 0, st>>>(T, params_dev, h->net[which].wt_image);
  h->launches += 1;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// generic image pack / unpack (test seams and small producers)
// ---------------------------------------------------------------------------
template <int FMT>
__global__ void k_pack_image(const float* __restrict__ src, long long n_rows, int n_cols, int ld, uint8_t* __restrict__ image) {
  // image [tiles][n_cols/64][32 KB]; rows past n_rows and columns past ld are zero
  const int cpt = n_cols / 64;
  const long long tiles = (n_rows + kT - 1) / kT;
  const long long total = tiles * cpt * kT * 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % 8), row = (int)((idx / 8) % kT);
    const long long tc = idx / (8 * kT);
    const int c = (int)(tc % cpt);
    const long long tile = tc / cpt, p = tile * kT + row;
    __align__(16) uint16_t hi[8];
    __align__(16) uint16_t lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = 64 * c + 8 * j + e;
      const float v = (p < n_rows && k < ld) ? src[p * ld + k] : 0.f;
      Split<FMT>::apply1(v, hi[e], lo[e]);
    }
    uint8_t* dst = image + (size_t)tc * kChunk + img2_off(row, j);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + kPlane) = *reinterpret_cast<const uint4*>(lo);
  }
}

template <int FMT> __device__ __forceinline__ float half_to_float(uint16_t x);
template <> __device__ __forceinline__ float half_to_float<1>(uint16_t x) { return __uint_as_float((uint32_t)x << 16); }
template <> __device__ __forceinline__ float half_to_float<0>(uint16_t x) { return __half2float(*reinterpret_cast<const __half*>(&x)); }

template <int FMT>
__global__ void k_unpack_image(const uint8_t* __restrict__ image, long long n_rows, int n_cols, int ld, float* __restrict__ dst) {
  const int cpt = n_cols / 64;
  const long long total = n_rows * ld;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long p = idx / ld;
    const int k = (int)(idx % ld);
    const long long tile = p / kT;
    const int row = (int)(p % kT), c = k / 64, j = (k % 64) / 8, e = k % 8;
    const uint8_t* s = image + (size_t)(tile * cpt + c) * kChunk + img2_off(row, j) + 2 * e;
    const uint16_t hi = *reinterpret_cast<const uint16_t*>(s), lo = *reinterpret_cast<const uint16_t*>(s + kPlane);
    dst[idx] = half_to_float<FMT>(hi) + half_to_float<FMT>(lo);
  }
}

// ---------------------------------------------------------------------------
// k_render_bwd: compositing + activation + rgb-head backward, one warp per ray
// ---------------------------------------------------------------------------
struct RenderBwdArgs {
  RenderParams rp;
  const SampleTables* tabs;
  const float* rays; long long n_rays; int ray_stride;
  const float* z;        // [N,S]
  const float* raw;      // [N,S,4]  rgb after the colour activation, raw sigma (VanillaMLP output)
  const float* noise;    // [N,S] or null
  int S;
  const float* g_rgb;    // [N,3] dL/d comp_rgb, or null
  const float* g_depth;  // [N] or null
  const float* g_opacity;// [N] or null
  const float* w_rgb;    // [3][128] fp32
  const uint8_t* stash_dir;   // [tiles][2][32 KB] (ReLU mask of the dir layer)
  uint8_t* dhead;        // [tiles][32 KB]    columns: 0 d sigma, 1..3 d rgb_pre
  uint8_t* dzdir;        // [tiles][2][32 KB]
  uint8_t* encdir;       // [tiles][32 KB]    27 valid columns
  float* dsig;           // [tiles*128]
  long long n_tiles;
};

template <int FMT>
__global__ void __launch_bounds__(128)
k_render_bwd(const RenderBwdArgs a) {
  extern __shared__ float bsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S, RPT = kT / S;
  const RenderParams& rp = a.rp;
  float* swr = bsm;                               // [384] rgb head weights
  float* base = bsm + 384 + warp * (10 * S + 32);
  float* sz = base;                               // z
  float* ssg = sz + S;                            // sigma (+ noise)
  float* sal = ssg + S;                           // alpha
  float* sT = sal + S;                            // 1 - alpha + eps  ->  transmittance (exclusive product)
  float* sG = sT + S;                             // dL/dw
  float* sR = sG + S;                             // G*w -> suffix sums
  float* sex = sR + S;                            // exp(-delta * act)
  float* srgb = sex + S;                          // [3S] colours entering the compositing sum
  float* senc = srgb + 3 * S;                     // [32] view-direction encoding of the ray
  for (int i = threadIdx.x; i < 384; i += blockDim.x) swr[i] = a.w_rgb[i];
  __syncthreads();
  const float eps = 1e-10f;
  const long long ray_slots = a.n_tiles * RPT;
  for (long long ray = blockIdx.x * 4ll + warp; ray < ray_slots; ray += (long long)gridDim.x * 4) {
    const bool valid = ray < a.n_rays;
    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, go = 0.f;
    if (valid) {
      if (a.g_rgb) { gr = a.g_rgb[ray * 3]; gg = a.g_rgb[ray * 3 + 1]; gb = a.g_rgb[ray * 3 + 2]; }
      if (a.g_depth) gd = a.g_depth[ray];
      if (a.g_opacity) go = a.g_opacity[ray];
      const float4* r4 = reinterpret_cast<const float4*>(a.raw) + ray * S;
      for (int i = lane; i < S; i += 32) {
        const float4 v = r4[i];
        float cr = v.x, cg = v.y, cb = v.z;
        if (rp.gamma_correct) { cr = powf(cr, 1.f / 2.2f); cg = powf(cg, 1.f / 2.2f); cb = powf(cb, 1.f / 2.2f); }
        srgb[3 * i] = cr; srgb[3 * i + 1] = cg; srgb[3 * i + 2] = cb;
        float s = v.w;
        if (a.noise) s = __fadd_rn(s, __fmul_rn(a.noise[ray * S + i], rp.noise_std));
        ssg[i] = s;
        sz[i] = a.z[ray * S + i];
      }
      // view-direction encoding (same channel order as the forward's per-ray dir bias)
      {
        const float* vd = a.rays + ray * a.ray_stride + rp.viewdir_offset;
        float val = 0.f;
        if (lane < 3) val = vd[lane];
        else if (lane < 27) {
          const int idx = lane - 3, k = idx / 6, r = idx % 6, comp = r % 3;
          const float arg = __fmul_rn(a.tabs->freq_dir[k], vd[comp]);
          val = (r >= 3) ? cosf(arg) : sinf(arg);
        }
        senc[lane] = val;
      }
      __syncwarp();
      // forward recompute (models/rendering.py:89-103), same roundings as composite_ray_warp
      for (int i = lane; i < S; i += 32) {
        const float delta = (i + 1 < S) ? __fsub_rn(sz[i + 1], sz[i]) : 1e10f;
        const float e = expf(__fmul_rn(-delta, sigma_act(ssg[i], rp.sigma_softplus)));
        const float al = __fsub_rn(1.f, e);
        sex[i] = e; sal[i] = al;
        sT[i] = __fadd_rn(__fsub_rn(1.f, al), eps);
      }
      __syncwarp();
      if (lane == 0) seq_scan_one_lane<true>(sT, S, 1.f);
      __syncwarp();
      const float white = rp.white_bkgd ? 1.f : 0.f;
      for (int i = lane; i < S; i += 32) {
        const float w = __fmul_rn(sal[i], sT[i]);
        // comp = sum w*c (+ 1 - sum w), depth = sum w*z, opacity = sum w
        const float G = gr * (srgb[3 * i] - white) + gg * (srgb[3 * i + 1] - white) + gb * (srgb[3 * i + 2] - white) +
                        gd * sz[i] + go;
        sG[i] = G;
        sR[i] = G * w;
      }
      __syncwarp();
      if (lane == 0) {          // R_i = sum_{k>i} G_k w_k  (reverse exclusive scan)
        float acc = 0.f;
        for (int i = S - 1; i >= 0; --i) { const float t = sR[i]; sR[i] = acc; acc += t; }
      }
      __syncwarp();
    }
    // per sample: gradients of the MLP outputs, then the three image rows of this point
    for (int i = lane; i < S; i += 32) {
      float dsg = 0.f, drgb[3] = {0.f, 0.f, 0.f};
      const long long p = ray * S + i;
      if (valid) {
        const float delta = (i + 1 < S) ? __fsub_rn(sz[i + 1], sz[i]) : 1e10f;
        const float x = __fadd_rn(__fsub_rn(1.f, sal[i]), eps);
        const float w = __fmul_rn(sal[i], sT[i]);
        const float dalpha = sG[i] * sT[i] - sR[i] / x;           // d w_i/d alpha_i = T_i ; d T_k/d alpha_i = -T_k / x_i
        const float dact = dalpha * delta * sex[i];               // alpha = 1 - exp(-delta * act)
        const float s = ssg[i];
        if (rp.sigma_softplus) {
          const float t = expf(__fsub_rn(s, 1.f));
          dsg = dact * (t / (1.f + t));
        } else {
          dsg = (s > 0.f) ? dact : 0.f;                           // relu backward: select, never 0 * inf
        }
        const float4 v = reinterpret_cast<const float4*>(a.raw)[p];
        const float c[3] = {v.x, v.y, v.z};
        const float g3[3] = {gr, gg, gb};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float d = g3[k] * w;
          if (rp.gamma_correct) d *= (1.f / 2.2f) * powf(c[k], 1.f / 2.2f - 1.f);
          if (!rp.color_none) d *= (1.f - c[k]) * c[k];           // sigmoid backward
          drgb[k] = d;
        }
      }
      const long long tile = p >> 7;
      const int row = (int)(p & 127);
      a.dsig[p] = dsg;
      // the three image rows of this point: 16-byte pieces j = 0..7 of each 64-feature chunk, lanes = consecutive points
      // (one 512-byte contiguous warp store per piece and plane)
      {   // dHead: (d sigma, d r, d g, d b, 0...) in piece 0
        uint32_t h0, l0, h1, l1;
        Split<FMT>::apply(dsg, drgb[0], h0, l0);
        Split<FMT>::apply(drgb[1], drgb[2], h1, l1);
        uint8_t* g = a.dhead + (size_t)tile * kChunk + img2_off(row, 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          stg128(g + 1024 * j, j == 0 ? h0 : 0u, j == 0 ? h1 : 0u, 0u, 0u);
          stg128(g + kPlane + 1024 * j, j == 0 ? l0 : 0u, j == 0 ? l1 : 0u, 0u, 0u);
        }
      }
      {   // encdir: the ray's 27 encoded view-direction channels (+ zero pad to 64)
        uint8_t* g = a.encdir + (size_t)tile * kChunk + img2_off(row, 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
          if (j < 4 && valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) Split<FMT>::apply(senc[8 * j + 2 * q], senc[8 * j + 2 * q + 1], hi[q], lo[q]);
          }
          stg128(g + 1024 * j, hi[0], hi[1], hi[2], hi[3]);
          stg128(g + kPlane + 1024 * j, lo[0], lo[1], lo[2], lo[3]);
        }
      }
      // dZ_dir = (d rgb_pre . W_rgb) masked by the dir layer's ReLU (networks.py:221-222)
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint8_t* g = a.dzdir + (size_t)(tile * 2 + c) * kChunk + img2_off(row, 0);
        const uint8_t* m = a.stash_dir + (size_t)(tile * 2 + c) * kChunk + img2_off(row, 0);
#pragma unroll 2
        for (int j = 0; j < 8; ++j) {
          uint32_t hi[4], lo[4];
          const uint4 mk = *reinterpret_cast<const uint4*>(m + 1024 * j);        // hi plane of the stashed activation: != 0 <=> active
          const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = 64 * c + 8 * j + 2 * q;
            float d0 = drgb[0] * swr[n] + drgb[1] * swr[128 + n] + drgb[2] * swr[256 + n];
            float d1 = drgb[0] * swr[n + 1] + drgb[1] * swr[128 + n + 1] + drgb[2] * swr[256 + n + 1];
            if ((mw[q] & 0x7fffu) == 0u) d0 = 0.f;
            if ((mw[q] & 0x7fff0000u) == 0u) d1 = 0.f;
            Split<FMT>::apply(d0, d1, hi[q], lo[q]);
          }
          stg128(g + 1024 * j, hi[0], hi[1], hi[2], hi[3]);
          stg128(g + kPlane + 1024 * j, lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// shared pieces of the two GEMM kernels
// ---------------------------------------------------------------------------
constexpr int kGemmThreads = 192;     // warp 0 producer, warp 1 MMA issue, warps 2-5 epilogue
constexpr int kSlotBytes = 98304;
constexpr int kSmSlots = 0;
constexpr int kSmAux = 2 * kSlotBytes;            // 196608: w_sigma (dx) / tile of ones (dw), 2 KB
constexpr int kSmGBar = kSmAux + 2048;
constexpr int kSmGTmem = kSmGBar + 16 * 8;
constexpr int kSmemGemmBytes = kSmGTmem + 16;
enum { G_FULL = 0, G_EMPTY = 2, G_ACCFULL = 4, G_ACCEMPTY = 6 };

// K-major SWIZZLE_128B descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SWIZZLE_128B descriptor over the same rows: a row is one k (point), its 128 B hold 64
// consecutive M/N indices (features); 8-k groups are 1024 B apart (SBO), 64-feature blocks `lbo` bytes
// apart (LBO).  Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}
// The same two views over a TILE IMAGE held in shared memory (img2 layout, no swizzle; canonical forms per CUTLASS
// make_umma_desc, in 16-byte units):
//   MN-major ((1,n),(8,k)):((X,SBO),(1,LBO)): 8 points of one 8-feature group are 128 contiguous bytes; the next 8 points
//            follow at LBO = 128 B, the next 8-feature group at SBO = 1024 B (a 64-point half: [j][64 points][16 B]; the
//            64-feature blocks of an operand are placed 8 KB apart, so the group stride is uniform across blocks)
//   K-major  ((8,n),2):((1,SBO),LBO): 8 points x 16 B contiguous; next 8 points at SBO = 128 B, the second 8-k piece of a
//            K = 16 step at LBO = 2048 B (a full tile: [j][128 points][16 B], the two halves of a j-block side by side)
__device__ __forceinline__ uint64_t img_mn_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t img_k_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
// Copy one chunk PLANE (16 KB in HBM: [half][j][64 points]) into the K-major arrangement [j][128 points] (16 x 1 KB pieces).
__device__ __forceinline__ void copy_plane_kmajor(uint32_t dst, const uint8_t* src, uint32_t bar) {
#pragma unroll 1
  for (int i = 0; i < 16; ++i) {
    const int half = i >> 3, j = i & 7;
    bulk_copy_g2s(dst + j * 2048 + half * 1024, src + half * 8192 + j * 1024, 1024, bar);
  }
}
__host__ __device__ constexpr uint32_t gemm_idesc(int fmt, int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t gemm_prologue(uint8_t* sm, uint32_t sm_base, int accempty_count) {
  if (sm_base & 1023u) { if (threadIdx.x == 0) printf("[nsr_train] dynamic smem base %u not 1024-aligned\n", sm_base); __trap(); }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const uint32_t bar = sm_base + kSmGBar;
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar + 8 * (G_FULL + i), 1); mbar_init(bar + 8 * (G_EMPTY + i), 1);
      mbar_init(bar + 8 * (G_ACCFULL + i), 1); mbar_init(bar + 8 * (G_ACCEMPTY + i), accempty_count);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_base + kSmGTmem), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(sm + kSmGTmem);
}
__device__ __forceinline__ void gemm_epilogue_free(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------
// k_tg_dx: out = relu_mask . (A . Wt^T [+ dsig x wsig])   per 128-point tile
//
// Epilogue data path: TMEM -> registers -> (mask, hi/lo split) -> a 32 KB shared-memory staging chunk in
// tile-image layout -> ONE cp.async.bulk store per 64-feature chunk.  (The first version stored 16-byte
// pieces straight to global memory: 32 lanes x 128-B stride = 32 half-used sectors per instruction, and
// read the mask from the stashed hi plane the same way; the load/store unit, not the tensor pipe or HBM,
// bounded the kernel -- profiles/r01_train_launches.md.)  The ReLU mask is the forward's 1-bit-per-
// activation stash: 32 B per row, read once per tile.
// ---------------------------------------------------------------------------
struct DxArgs {
  const uint8_t* a_img; int nkc;        // [n_tiles][nkc][32 KB]   dZ of this layer (K = its output features)
  const uint8_t* wt_img;                // [2][nkc][32 KB]         this layer's stages of the transposed-weight image (half, chunk)
  uint8_t* out_img;                     // [n_tiles][4][32 KB]     dZ of the previous layer
  const uint32_t* mask_bits;            // [n_tiles][128][8]       ReLU mask gating the output, or null
  const float* dsig; const float* wsig; // rank-1 term of the sigma head (final layer only), or null
  long long n_tiles;
};
constexpr int kSmDxStage = 2 * kSlotBytes;            // 196608: 32 KB output staging chunk
constexpr int kSmDxAux = kSmDxStage + kChunk;         // 229376: w_sigma (1 KB)
constexpr int kSmDxBar = kSmDxAux + 1024;
constexpr int kSmDxTmem = kSmDxBar + 16 * 8;
constexpr int kSmemDxBytes = kSmDxTmem + 16;
static_assert(kSmemDxBytes <= 227 * 1024, "dx smem budget");

__device__ __forceinline__ void bulk_store_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(kGemmThreads, 1) k_tg_dx(const DxArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  const uint32_t sm_base = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long my_tiles = (a.n_tiles > blockIdx.x) ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (a.wsig) for (int i = threadIdx.x; i < 256; i += kGemmThreads) reinterpret_cast<float*>(sm + kSmDxAux)[i] = a.wsig[i];
  if (sm_base & 1023u) { if (threadIdx.x == 0) printf("[nsr_train] dynamic smem base %u not 1024-aligned\n", sm_base); __trap(); }
  const uint32_t bar = sm_base + kSmDxBar;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar + 8 * (G_FULL + i), 1); mbar_init(bar + 8 * (G_EMPTY + i), 1);
      mbar_init(bar + 8 * (G_ACCFULL + i), 1); mbar_init(bar + 8 * (G_ACCEMPTY + i), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_base + kSmDxTmem), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kSmDxTmem);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t n = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * (long long)gridDim.x;
        for (int c = 0; c < a.nkc; ++c, ++n) {
          const uint32_t slot = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar + 8 * (G_EMPTY + slot), par ^ 1u);
          const uint32_t full = bar + 8 * (G_FULL + slot);
          const uint32_t dst = sm_base + kSmSlots + slot * kSlotBytes;
          mbar_expect_tx(full, kSlotBytes);
          const uint8_t* asrc = a.a_img + ((size_t)tile * a.nkc + c) * kChunk;
          copy_plane_kmajor(dst, asrc, full);                     // hi plane -> [j][128 points]
          copy_plane_kmajor(dst + kPlane, asrc + kPlane, full);   // lo plane
          for (int hf = 0; hf < 2; ++hf) {      // 256-row hi plane at +32 KB, lo plane at +64 KB: rows [128 hf, +128) from stage (hf, c)
            const uint8_t* stg = a.wt_img + ((size_t)hf * a.nkc + c) * kChunk;
            bulk_copy_g2s(dst + 32768 + hf * kPlane, stg, kPlane, full);
            bulk_copy_g2s(dst + 65536 + hf * kPlane, stg + kPlane, kPlane, full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = gemm_idesc(FMT, 128, 256, 0, 0);
      uint32_t n = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        const uint32_t buf = (uint32_t)(it & 1);
        mbar_wait(bar + 8 * (G_ACCEMPTY + buf), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem + 256u * buf;
        for (int c = 0; c < a.nkc; ++c, ++n) {
          const uint32_t slot = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar + 8 * (G_FULL + slot), par);
          tc_fence_after();
          const uint32_t sa = sm_base + kSmSlots + slot * kSlotBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ah = img_k_desc(sa + 4096 * k), al = img_k_desc(sa + kPlane + 4096 * k);
            const uint64_t bh = kmajor_desc(sa + 32768 + 32 * k), bl = kmajor_desc(sa + 65536 + 32 * k);
            mma_ss(d, ah, bh, idesc, (c | k) ? 1u : 0u);
            mma_ss(d, al, bh, idesc, 1u);
            mma_ss(d, ah, bl, idesc, 1u);
          }
          tc_commit(bar + 8 * (G_EMPTY + slot));
        }
        tc_commit(bar + 8 * (G_ACCFULL + buf));
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = 32 * q + lane, r7 = lane & 7;
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    const float* wsig = reinterpret_cast<const float*>(sm + kSmDxAux);
    const uint32_t stage_row = sm_base + kSmDxStage + (uint32_t)img2_off(row, 0);
    const bool issuer = (warp == 2 && lane == 0);
    for (long long it = 0; it < my_tiles; ++it) {
      const uint32_t buf = (uint32_t)(it & 1);
      const long long tile = blockIdx.x + it * (long long)gridDim.x;
      uint32_t mbits[8];
      if (a.mask_bits) {
        const uint4* mp = reinterpret_cast<const uint4*>(a.mask_bits + ((size_t)tile * kT + row) * 8);
        const uint4 m0 = mp[0], m1 = mp[1];
        mbits[0] = m0.x; mbits[1] = m0.y; mbits[2] = m0.z; mbits[3] = m0.w; mbits[4] = m1.x; mbits[5] = m1.y; mbits[6] = m1.z; mbits[7] = m1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) mbits[i] = 0xffffffffu;
      }
      const float ds = a.dsig ? a.dsig[tile * kT + row] : 0.f;
      mbar_wait(bar + 8 * (G_ACCFULL + buf), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        if ((cc & 1) == 0) {      // a new 64-feature chunk: the staging buffer must have been read out
          if (issuer) bulk_store_wait_read();
          named_bar_sync(3, 128);
        }
        uint32_t r[32];
        TMEM_LD32(tlane + 256u * buf + 32u * (uint32_t)cc, r);
        tc_wait_ld();
        const uint32_t mb = mbits[cc];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = 8 * jj + 2 * e;
            float v0 = __uint_as_float(r[col]), v1 = __uint_as_float(r[col + 1]);
            if (a.wsig) { v0 = fmaf(ds, wsig[32 * cc + col], v0); v1 = fmaf(ds, wsig[32 * cc + col + 1], v1); }
            if (!((mb >> col) & 1u)) v0 = 0.f;
            if (!((mb >> (col + 1)) & 1u)) v1 = 0.f;
            Split<FMT>::apply(v0, v1, hi[e], lo[e]);
          }
          const uint32_t pj = (uint32_t)(((cc & 1) * 4 + jj) * 1024);     // piece j of the chunk
          sts128(stage_row + pj, hi[0], hi[1], hi[2], hi[3]);
          sts128(stage_row + kPlane + pj, lo[0], lo[1], lo[2], lo[3]);
        }
        if (cc & 1) {             // chunk complete: hand it to the bulk-copy engine
          fence_proxy_async();
          named_bar_sync(3, 128);
          if (issuer) bulk_store_s2g(a.out_img + ((size_t)tile * 4 + (size_t)(cc >> 1)) * kChunk, sm_base + kSmDxStage, kChunk);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + 8 * (G_ACCEMPTY + buf));
    }
    if (issuer) bulk_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------
// k_tg_dxchain: the whole dX chain of one net, fused per 128-point tile (dir -> final -> L8 .. L2).
//
// Same on-chip dataflow AND the same issue schedule as the forward pass (nsr_tc.cu, shared code in nsr_tc_mma.cuh): the
// gradient of a layer's input never leaves the SM between layers -- the epilogue writes it back into TMEM as the next
// layer's A operand (hi/lo planes, TS-form MMAs) -- and the transposed weights stream through a 4-slot ring of 32 KB
// stages from L2.  HBM sees only what the dW GEMMs need afterwards: each layer's dZ image written once, the 1-bit ReLU
// masks read once.
//
// Schedule (round 2; round 1 issued N = 256 over the whole accumulator and only then ran the epilogue: tensor pipe 32 %
// active, 13.1 k cycles per tile-layer of which 6.1 k were MMAs).  The 256-column accumulator is two N = 128 halves; per
// layer the MMA lane issues half n0 over k-chunks c0..c3, then n1; the epilogue handles 64-column quarters q0..q3:
//   M(l+1, n0, c) needs quarter c of layer l's epilogue (A_READY[c]: accumulator quarter drained, A chunk c rewritten);
//   E(l, q) needs its accumulator half (ACC_FULL) and A chunk q no longer read by layer l (A_FREE);
// so every quarter-epilogue has >= 1536 cycles of queued MMAs to hide behind, and the dZ image stores (the reason this
// kernel exists for HBM) are issued AFTER the A_READY arrive, off the MMA lane's critical path.
//   warps 0-7  epilogue  (warp & 3 = TMEM lane quarter, warp >> 2 = 32-column half of the 64-column quarter)
//   warp  8    producer  (per tile: the dZ_dir tile, 64 KB, into its own buffer; then 68 weight stages through the ring)
//   warp  9    MMA issue (one elected lane; layer 0 SS-form from the dZ_dir buffer, layers 1..8 TS-form)
// A tile has 68 = 8 * 8 + 4 stages and 9 layers, so ring parities alternate between even and odd tiles: two instances of
// the tile body (stage phase P0 = 0 / 4); per-layer barrier parities are runtime (layer counter g).
// TMEM: [0,256) fp32 accumulator, [256,384) A hi plane, [384,512) A lo plane.
// ---------------------------------------------------------------------------
struct ChainArgs {
  const uint8_t* dzdir;     // [n_tiles][2][32 KB]
  const uint8_t* wt;        // kWtStages x 32 KB in consumption order (build_wt_table)
  uint8_t* dz;              // [9][n_tiles][4][32 KB]: 0 = d feat, 1 = dZ_8, ..., 8 = dZ_1
  const uint32_t* mask;     // [8][n_tiles][128][8]: ReLU bits of h_1..h_8
  const float* dsig; const float* wsig;
  long long n_tiles;
  int debug_flags;          // timing experiments (WRONG results): 8 = skip the dZ image stores, 16 = skip the mask loads,
                            // 32 = the weight producer re-arms the ring without copying (stale weights)
  long long* clk;           // SM-clock probe of CTA 0 (see TcKernelArgs::clk), or null
};
constexpr int kChEpiWarps = 8;
constexpr int kChProducerWarp = kChEpiWarps;          // warp 8
constexpr int kChMmaWarp = kChEpiWarps + 1;           // warp 9 (highest id: the arbiter favours it)
constexpr int kChainThreads = 32 * (kChEpiWarps + 2);
constexpr int kChRing = 0;                            // 4 x 32 KB weight stages
constexpr int kChDz = kRing * kStageBytes;            // 131072: the tile's dZ_dir image, 2 chunks x 32 KB
constexpr int kChStage = kChDz + 2 * kChunk;          // 196608: 32 KB staging chunk for the dZ image bulk stores
constexpr int kChAux = kChStage + kChunk;             // 229376: w_sigma (1 KB)
constexpr int kChBar = kChAux + 1024;
constexpr int kChTmem = kChBar + 32 * 8;
constexpr int kSmemChainBytes = kChTmem + 16;
static_assert(kSmemChainBytes <= 227 * 1024, "chain smem budget");
constexpr int C_DZFULL = 16, C_DZEMPTY = 17;          // (indices 0..15: B_WFULL .. B_AREADY of nsr_tc_mma.cuh)
static_assert(kStageBytes == kChunk && kPlaneBytes == kPlane, "ring stage == tile-image chunk");

// Layer-0 stage of the chain: A = one 64-k chunk of the tile's dZ_dir image in shared memory (K-major view of the tile-image
// layout, img_k_desc), B = a weight stage of the ring (K-major SWIZZLE_128B).  Mirrors mma_stage_ss (nsr_tc_mma.cuh).
template <int N8>
__device__ __forceinline__ void chain_stage_ss(const MmaCtx& c, uint32_t a_chunk, int half, bool first) {
  constexpr uint64_t HI = (64ull << 32) | (1ull << 46) | (2ull << 61);
  constexpr int slot = N8 & 3;
  constexpr int nslot = (N8 + 1) & 3, npar = ((N8 + 1) >> 2) & 1;
  const uint32_t wlo = ((c.ring + slot * kStageBytes) >> 4) | (1u << 16);
  const uint32_t d = 128u * half;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t bh = HI | (uint64_t)(wlo + 2u * k), bl = HI | (uint64_t)(wlo + 1024u + 2u * k);
    const uint64_t ah = img_k_desc(a_chunk + 4096u * k), al = img_k_desc(a_chunk + kPlane + 4096u * k);
    mma_ss(d, ah, bh, c.idesc, (k == 0 && first) ? 0u : 1u);
    mma_ss(d, al, bh, c.idesc, 1u);
    mma_ss(d, ah, bl, c.idesc, 1u);
    if (k == 1) { mbar_wait(c.bar + 8 * (B_WFULL + nslot), npar); tc_fence_after(); }     // next stage's weights
  }
  ring_release(c, slot);
}

// One tile of the MMA lane.  P0 = stage index of the tile's first stage mod 8 (0 for even tiles, 4 for odd ones).
template <int P0>
__device__ __forceinline__ void chain_tile_mma(const MmaCtx& c, uint32_t dz, uint32_t dz_par, uint32_t g0, bool last_tile) {
  // ---- layer 0 (dir layer, K = 128 = 2 chunks), A = the dZ_dir tile in shared memory
  mbar_wait(c.bar + 8 * C_DZFULL, dz_par);
  tc_fence_after();
  const uint32_t gp = (g0 - 1u) & 1u;                 // A_READY parity of the previous tile's last layer
  if (g0 > 0) { mma_wait_ready(c, 0, gp); mma_wait_ready(c, 1, gp); tc_fence_after(); }      // accumulator half 0 drained
  chain_stage_ss<(P0 + 0) & 7>(c, dz, 0, true);
  chain_stage_ss<(P0 + 1) & 7>(c, dz + kChunk, 0, false);
  tc_commit(c.bar + 8 * (B_ACCFULL + 0));
  if (g0 > 0) { mma_wait_ready(c, 2, gp); mma_wait_ready(c, 3, gp); tc_fence_after(); }      // half 1 drained
  chain_stage_ss<(P0 + 2) & 7>(c, dz, 1, true);
  tc_commit(c.bar + 8 * (B_AFREE + 0));
  chain_stage_ss<(P0 + 3) & 7>(c, dz + kChunk, 1, false);
  tc_commit(c.bar + 8 * (B_AFREE + 1));
  tc_commit(c.bar + 8 * (B_ACCFULL + 1));
  tc_commit(c.bar + 8 * C_DZEMPTY);                   // the dZ_dir buffer may be refilled for the next tile
  // ---- layers 1..8 (final, L8 .. L2): K = 256, A = the previous layer's masked gradient in TMEM
#pragma unroll 1
  for (uint32_t l = 1; l <= 8; ++l)
    mma_layer<3, (P0 + 4) & 7, false>(c, 0u, (g0 + l - 1u) & 1u, !(last_tile && l == 8));
}

template <int FMT>
__global__ void __launch_bounds__(kChainThreads, 1) k_tg_dxchain(const ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  const uint32_t sm_base = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long my_tiles = (a.n_tiles > blockIdx.x) ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (a.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    a.clk[0] = clock64(); a.clk[1] = (long long)ns;
  }
  for (int i = threadIdx.x; i < 256; i += kChainThreads) reinterpret_cast<float*>(sm + kChAux)[i] = a.wsig[i];
  if (sm_base & 1023u) { if (threadIdx.x == 0) printf("[nsr_train] dynamic smem base %u not 1024-aligned\n", sm_base); __trap(); }
  const uint32_t bar = sm_base + kChBar;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(bar + 8 * (B_WFULL + i), 1); mbar_init(bar + 8 * (B_WEMPTY + i), 1); }
    mbar_init(bar + 8 * (B_ACCFULL + 0), 1); mbar_init(bar + 8 * (B_ACCFULL + 1), 1);
    mbar_init(bar + 8 * (B_AFREE + 0), 1); mbar_init(bar + 8 * (B_AFREE + 1), 1);
    for (int q4 = 0; q4 < 4; ++q4) mbar_init(bar + 8 * (B_AREADY + q4), kChEpiWarps);
    mbar_init(bar + 8 * C_DZFULL, 1); mbar_init(bar + 8 * C_DZEMPTY, 1);
    fence_barrier_init();
  }
  if (warp == kChMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_base + kChTmem), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 columns are ours (one CTA per SM): the base is column 0, which the shared issue code assumes (checked)
  if (*reinterpret_cast<volatile uint32_t*>(sm + kChTmem) != 0u) {
    if (threadIdx.x == 0) printf("[nsr_train] unexpected TMEM base %u\n", *reinterpret_cast<volatile uint32_t*>(sm + kChTmem));
    __trap();
  }
  const uint32_t tmem = 0u;

  if (warp == kChProducerWarp) {
    if (elect_one()) {
      uint32_t slot = 0, par = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        const long long tile = blockIdx.x + it * (long long)gridDim.x;
        mbar_wait(bar + 8 * C_DZEMPTY, (uint32_t)((it & 1) ^ 1));          // layer 0 of the previous tile is done with the buffer
        mbar_expect_tx(bar + 8 * C_DZFULL, 2 * kChunk);
        for (int cp = 0; cp < 4; ++cp)                                     // (chunk, plane): 16 KB each -> [j][128 points]
          copy_plane_kmajor(sm_base + kChDz + cp * kPlane, a.dzdir + (size_t)tile * 2 * kChunk + (size_t)cp * kPlane, bar + 8 * C_DZFULL);
#pragma unroll 1
        for (int s = 0; s < kWtStages; ++s) {
          mbar_wait(bar + 8 * (B_WEMPTY + slot), par ^ 1u);
          const uint32_t full = bar + 8 * (B_WFULL + slot);
          if ((a.debug_flags & 32) && (it > 0 || s >= kRing)) { mbar_arrive(full); }
          else {
            mbar_expect_tx(full, kStageBytes);
            bulk_copy_g2s(sm_base + kChRing + slot * kStageBytes, a.wt + (size_t)s * kStageBytes, kStageBytes, full);
          }
          if (++slot == kRing) { slot = 0; par ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kChMmaWarp) {
    if (elect_one()) {
      const MmaCtx c{sm_base + kChRing, bar, umma_idesc(FMT, 128, 128), 0u};
      if (my_tiles > 0) { mbar_wait(bar + 8 * (B_WFULL + 0), 0); tc_fence_after(); }      // the first stage's weights
#pragma unroll 1
      for (long long it = 0; it < my_tiles; ++it) {
        const uint32_t g0 = 9u * (uint32_t)it;
        const bool last = it + 1 == my_tiles;
        if (it & 1) chain_tile_mma<4>(c, sm_base + kChDz, 1u, g0, last);
        else chain_tile_mma<0>(c, sm_base + kChDz, 0u, g0, last);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, hh = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    const float* wsig = reinterpret_cast<const float*>(sm + kChAux);
    const uint32_t stage_row = sm_base + kChStage + (uint32_t)img2_off(row, 0);
    const bool issuer = (warp == 0 && lane == 0);
    uint32_t g = 0;
#pragma unroll 1
    for (long long it = 0; it < my_tiles; ++it) {
      const long long tile = blockIdx.x + it * (long long)gridDim.x;
      const float ds = a.dsig[tile * kT + row];
#pragma unroll 1
      for (int l = 0; l < 9; ++l, ++g) {
        // output of step l >= 1 is the gradient w.r.t. h_{9-l}: gate it with that layer's ReLU bits (this thread's row,
        // mask word 2 q4 + hh of the 8 covers columns [64 q4 + 32 hh, +32))
        uint32_t mq[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (l >= 1 && !(a.debug_flags & 16)) {
          const uint4* mp = reinterpret_cast<const uint4*>(a.mask + (((size_t)(8 - l) * (size_t)a.n_tiles + (size_t)tile) * kT + row) * 8);
          const uint4 m0 = mp[0], m1 = mp[1];
          mq[0] = hh ? m0.y : m0.x; mq[1] = hh ? m0.w : m0.z; mq[2] = hh ? m1.y : m1.x; mq[3] = hh ? m1.w : m1.z;
        }
        uint8_t* ochunks = a.dz + (((size_t)l * (size_t)a.n_tiles + (size_t)tile) * 4) * kChunk;
#pragma unroll 1
        for (int q4 = 0; q4 < 4; ++q4) {
          // every barrier is waited for in every layer (one phase per layer: nobody can fall two phases behind)
          if (q4 == 0) { mbar_wait(bar + 8 * (B_ACCFULL + 0), g & 1u); mbar_wait(bar + 8 * (B_AFREE + 0), g & 1u); }
          else if (q4 == 1) mbar_wait(bar + 8 * (B_AFREE + 1), g & 1u);
          else if (q4 == 2) mbar_wait(bar + 8 * (B_ACCFULL + 1), g & 1u);
          tc_fence_after();
          const int col0 = 64 * q4 + 32 * hh;
          uint32_t r[32];
          TMEM_LD32(tlane + (uint32_t)col0, r);
          tc_wait_ld();
          uint32_t whi[16], wlo[16];
          const uint32_t mb = mq[q4];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float v0 = __uint_as_float(r[2 * e]), v1 = __uint_as_float(r[2 * e + 1]);
            if (l == 1) { v0 = fmaf(ds, wsig[col0 + 2 * e], v0); v1 = fmaf(ds, wsig[col0 + 2 * e + 1], v1); }
            if (!((mb >> (2 * e)) & 1u)) v0 = 0.f;
            if (!((mb >> (2 * e + 1)) & 1u)) v1 = 0.f;
            Split<FMT>::apply(v0, v1, whi[e], wlo[e]);
          }
          if (l < 8) {       // next layer's A operand (hi and lo planes), two k-values per TMEM column
            TMEM_ST16(tlane + 256u + (uint32_t)(col0 / 2), whi);
            TMEM_ST16(tlane + 384u + (uint32_t)(col0 / 2), wlo);
            tc_wait_st();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar + 8 * (B_AREADY + q4));
          // the layer's dZ image (what the dW GEMMs read) -- after the arrive: the MMA lane has moved on.  The 64-column
          // quarter of the tile is one 32 KB image chunk: the eight warps lay it out in shared memory (512 contiguous bytes per
          // warp store) and ONE thread hands it to the bulk-copy engine, which drains it to HBM asynchronously.  (Storing
          // from registers -- even fully coalesced -- kept every epilogue warp blocked on its own store issue for longer than
          // the MMA window it has: the chain ran at 347 us per 1024 tiles against 258 us without the stores.)
          if (a.debug_flags & 8) continue;
          if (issuer) bulk_store_wait_read();                 // the previous chunk has left shared memory
          named_bar_sync(3, 32 * kChEpiWarps);
          const uint32_t sp = stage_row + 4096u * (uint32_t)hh;
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            sts128(sp + 1024u * t2, whi[4 * t2], whi[4 * t2 + 1], whi[4 * t2 + 2], whi[4 * t2 + 3]);
            sts128(sp + kPlane + 1024u * t2, wlo[4 * t2], wlo[4 * t2 + 1], wlo[4 * t2 + 2], wlo[4 * t2 + 3]);
          }
          fence_proxy_async();
          named_bar_sync(3, 32 * kChEpiWarps);
          if (issuer) bulk_store_s2g(ochunks + (size_t)q4 * kChunk, sm_base + kChStage, kChunk);
        }
      }
    }
    if (issuer) bulk_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kChMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  if (a.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    a.clk[2] = clock64(); a.clk[3] = (long long)ns;
  }
}

// ---------------------------------------------------------------------------
// k_tg_dw: ALL dW GEMMs of one net in ONE persistent launch (round 2; round 1 launched each of the 13 separately and paid a
// ramp, a TMEM allocation and a tail per launch: ~5 us x 26 per iteration, first-order at the 256-rays-per-GPU DDP shape).
// Per GEMM g:  partial[split][n][m] = sum over this CTA's tiles of A[p][m] * B[p][n];  bias partial[split][m] = sum A[p][m].
// One CTA per SM.  The work items of the launch are (GEMM g, job, split), n_jobs_g x n_split_g per GEMM, numbered GEMM after
// GEMM; CTA b takes items b, b + grid, ... (at frame-sized batches every GEMM has as many items as there are CTAs, i.e. the
// decomposition of the former per-GEMM grids; at small batches, where a GEMM has fewer items than SMs, different GEMMs run
// side by side).  Items are independent (they read the stash / dZ images and write disjoint partial regions), so the CTA
// simply walks its list: the operand ring (2 x 96 KB), its barrier parities and the
// TMEM allocation carry over from one GEMM to the next, the producer prefetches GEMM g+1 while the epilogue of g drains the
// accumulator (ACC_FULL / ACC_EMPTY hand it back and forth).  Partials are stored [n][m] so that the 32 lanes of an
// epilogue warp (= 32 consecutive output rows m) write 128 contiguous bytes per column and k_grad_reduce reads them the
// same way.
// ---------------------------------------------------------------------------
struct DwJob {
  const uint8_t* a_img; int a_cpt;      // A image and its chunks per tile
  int a_blk0, a_blk1;                   // chunk indices giving output rows 0..63 and 64..127
  float* part;                          // [n_split][N][128]
  float* part_bias;                     // [n_split][128]
};
constexpr int kMaxDwJobs = 3;
constexpr int kMaxDwGemms = 14;
struct DwGemm {
  DwJob job[kMaxDwJobs]; int n_jobs;
  const uint8_t* b_img; int b_cpt; int b_chunk0; int nB;    // N = 64 * nB
  int n_split;
  int item0;                // index of this GEMM's first work item in the launch-wide list
};
struct DwAllArgs { DwGemm g[kMaxDwGemms]; int n_gemms; int n_items; long long n_tiles; };
static_assert(sizeof(DwAllArgs) <= 4000, "kernel parameter space");

template <int FMT>
__global__ void __launch_bounds__(kGemmThreads, 1) k_tg_dw(const DwAllArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  const uint32_t sm_base = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {   // a 16-row K-major tile of ones for the bias-gradient MMA
    const uint32_t one2 = (FMT == 1) ? 0x3f803f80u : 0x3c003c00u;
    for (int i = threadIdx.x; i < 512; i += kGemmThreads) reinterpret_cast<uint32_t*>(sm + kSmAux)[i] = one2;
    fence_proxy_async();
  }
  const uint32_t tmem = gemm_prologue(sm, sm_base, 4);
  const uint32_t bar = sm_base + kSmGBar;
  // work item -> (GEMM, job, split) and its number of tiles (the same arithmetic in every role)
  auto my_work = [&](int item, int& gi, int& job, int& split) -> long long {
    gi = 0;
    while (gi + 1 < a.n_gemms && item >= a.g[gi + 1].item0) ++gi;
    const DwGemm& G = a.g[gi];
    const int local = item - G.item0;
    job = local / G.n_split; split = local % G.n_split;
    return (a.n_tiles > split) ? (a.n_tiles - split + G.n_split - 1) / G.n_split : 0;
  };

  if (warp == 0) {
    if (elect_one()) {
      uint32_t n = 0;
#pragma unroll 1
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int gi, job, split;
        const long long my_tiles = my_work(item, gi, job, split);
        if (my_tiles <= 0) continue;
        const DwGemm& G = a.g[gi];
        const DwJob& J = G.job[job];
        const uint32_t bytes = (uint32_t)(4 + 2 * G.nB) * 8192u;
        for (long long it = 0; it < my_tiles; ++it) {
          const long long tile = split + it * (long long)G.n_split;
          for (int half = 0; half < 2; ++half, ++n) {
            const uint32_t slot = n & 1u, par = (n >> 1) & 1u;
            mbar_wait(bar + 8 * (G_EMPTY + slot), par ^ 1u);
            const uint32_t full = bar + 8 * (G_FULL + slot);
            const uint32_t dst = sm_base + kSmSlots + slot * kSlotBytes;
            mbar_expect_tx(full, bytes);
            for (int b = 0; b < 2; ++b) {
              const uint8_t* src = J.a_img + ((size_t)tile * J.a_cpt + (b ? J.a_blk1 : J.a_blk0)) * kChunk + (size_t)half * 8192;
              bulk_copy_g2s(dst + b * 8192, src, 8192, full);
              bulk_copy_g2s(dst + 16384 + b * 8192, src + kPlane, 8192, full);
            }
            for (int b = 0; b < G.nB; ++b) {
              const uint8_t* src = G.b_img + ((size_t)tile * G.b_cpt + G.b_chunk0 + b) * kChunk + (size_t)half * 8192;
              bulk_copy_g2s(dst + 32768 + b * 8192, src, 8192, full);
              bulk_copy_g2s(dst + 65536 + b * 8192, src + kPlane, 8192, full);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_b = gemm_idesc(FMT, 128, 16, 1, 0);
      const uint64_t ones = kmajor_desc(sm_base + kSmAux);
      uint32_t n = 0, uses = 0;
#pragma unroll 1
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int gi, job, split;
        const long long my_tiles = my_work(item, gi, job, split);
        if (my_tiles <= 0) continue;
        const DwGemm& G = a.g[gi];
        const uint32_t idesc = gemm_idesc(FMT, 128, 64 * G.nB, 1, 1);
        if (uses > 0) { mbar_wait(bar + 8 * (G_ACCEMPTY + 0), (uses - 1u) & 1u); tc_fence_after(); }   // previous GEMM drained
        for (long long it = 0; it < my_tiles; ++it) {
          for (int half = 0; half < 2; ++half, ++n) {
            const uint32_t slot = n & 1u, par = (n >> 1) & 1u;
            mbar_wait(bar + 8 * (G_FULL + slot), par);
            tc_fence_after();
            const uint32_t sa = sm_base + kSmSlots + slot * kSlotBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t first = (it | (long long)half | (long long)k) ? 1u : 0u;
              const uint64_t ah = img_mn_desc(sa + 256 * k), al = img_mn_desc(sa + 16384 + 256 * k);
              const uint64_t bh = img_mn_desc(sa + 32768 + 256 * k), bl = img_mn_desc(sa + 65536 + 256 * k);
              mma_ss(tmem, ah, bh, idesc, first);
              mma_ss(tmem, al, bh, idesc, 1u);
              mma_ss(tmem, ah, bl, idesc, 1u);
              mma_ss(tmem + 256u, ah, ones, idesc_b, first);
              mma_ss(tmem + 256u, al, ones, idesc_b, 1u);
            }
            tc_commit(bar + 8 * (G_EMPTY + slot));
          }
        }
        tc_commit(bar + 8 * (G_ACCFULL + 0));
        ++uses;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    uint32_t uses = 0;
#pragma unroll 1
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      int gi, job, split;
      const long long my_tiles = my_work(item, gi, job, split);
      const DwGemm& G = a.g[gi];
      const DwJob& J = G.job[job];
      const int N = 64 * G.nB;
      float* dst = J.part + (size_t)split * N * 128 + row;              // [n][m]: lanes = consecutive rows m
      if (my_tiles > 0) {
        mbar_wait(bar + 8 * (G_ACCFULL + 0), uses & 1u);
        tc_fence_after();
      }
#pragma unroll 1
      for (int cc = 0; cc < 2 * G.nB; ++cc) {
        uint32_t r[32];
        if (my_tiles > 0) { TMEM_LD32(tlane + 32u * (uint32_t)cc, r); tc_wait_ld(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[(size_t)(32 * cc + i) * 128] = __uint_as_float(r[i]);
      }
      {
        uint32_t r[32];
        float b = 0.f;
        if (my_tiles > 0) { TMEM_LD32(tlane + 256u, r); tc_wait_ld(); b = __uint_as_float(r[0]); }
        J.part_bias[(size_t)split * 128 + row] = b;
      }
      if (my_tiles > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + 8 * (G_ACCEMPTY + 0));         // the accumulator may be overwritten
        ++uses;
      }
    }
  }
  gemm_epilogue_free(tmem);
}

// ---------------------------------------------------------------------------
// k_grad_reduce: fixed-order sum of the per-CTA partials into the flat gradient
// ---------------------------------------------------------------------------
struct RedJob {
  const float* part; const float* part_bias;
  int n_split, N, row0, rows, cols, ld, col0;
  long long dst, bias_dst;             // offsets into the flat gradient; bias_dst < 0: no bias
};
constexpr int kMaxRedJobs = 28;
struct RedArgs { RedJob job[kMaxRedJobs]; int n_jobs; float* grad; };

__global__ void __launch_bounds__(256) k_grad_reduce(const RedArgs a) {
  const RedJob& J = a.job[blockIdx.y];
  const int total = J.rows * J.cols;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx % J.rows, c = idx / J.rows;                  // partials are [split][n][m]: lanes walk the rows m
    const float* p = J.part + (size_t)c * 128 + J.row0 + r;
    const size_t stride = (size_t)128 * J.N;
    // eight independent chains (fixed order: deterministic); the eight loads of a trip are issued before the first add, which
    // is what keeps enough bytes in flight for a kernel that only streams 0.2 GB of partials
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 8 <= J.n_split; k += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (size_t)(k + u) * stride);      // read once: streaming
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += v[u];
    }
    for (; k < J.n_split; ++k) s[k & 7] += __ldcs(p + (size_t)k * stride);
    a.grad[J.dst + (long long)r * J.ld + J.col0 + c] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  }
  if (J.bias_dst >= 0) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < J.rows; r += gridDim.x * blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < J.n_split; ++k) s += J.part_bias[(size_t)k * 128 + J.row0 + r];
      a.grad[J.bias_dst + r] = s;
    }
  }
}

// ---------------------------------------------------------------------------
// loss gradient, clip coefficient, Adam
// ---------------------------------------------------------------------------
// box average + squared error + dL/d(hr) for lambda * mean((lr - target)^2)
__global__ void __launch_bounds__(256)
k_lr_loss_grad(const float* __restrict__ hr, const float* __restrict__ target, int64_t n_lr, int ss, float scale,
               float* __restrict__ lr_out, float* __restrict__ g_hr, double* __restrict__ partials) {
  const int64_t total = n_lr * 3;
  double acc = 0.0;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / 3;
    const int c = (int)(idx % 3);
    float sum = 0.f;
    for (int k = 0; k < ss; ++k) sum = __fadd_rn(sum, hr[(p * ss + k) * 3 + c]);
    const float lr = __fdiv_rn(sum, (float)ss);
    if (lr_out) lr_out[idx] = lr;
    const float d = __fsub_rn(lr, target[idx]);
    acc += (double)__fmul_rn(d, d);
    if (g_hr) {
      const float g = scale * d;
      for (int k = 0; k < ss; ++k) g_hr[(p * ss + k) * 3 + c] = g;
    }
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

// All of calculate_losses' terms for one net's outputs (models/nerf_downX_model.py:326-378) and their gradients
// down to the HR composite colour / depth: box average + lambda * MSE + PSNR, the s x s sub-pixel variance sums
// of colour and of depth / far (torch.var: unbiased, two-pass), and the MSE of the HR colours against the SISR
// image.  One thread per LR pixel (its s*s rows are contiguous: 16 s^2 bytes in, 16 s^2 out); four fp64 partial
// sums per block, fixed-order final reduction (deterministic, no atomics).
struct LossTerms {
  int ss;
  float g_mse;        // 2 lambda_mse / (3 n_lr) / s^2
  float g_var;        // 2 lambda_var / (s^2 - 1)
  float g_dvar;       // 2 lambda_depth_var / (s^2 - 1) / far
  float g_sr;         // 2 lambda_hr / (3 n_lr s^2)
  float far_plane;
  int want_var, want_dvar;
};

__global__ void __launch_bounds__(256)
k_loss_epilogue(const float* __restrict__ rgb, const float* __restrict__ depth, const float* __restrict__ target,
                const float* __restrict__ target_hr, int64_t n_lr, LossTerms t, float* __restrict__ lr_rgb,
                float* __restrict__ lr_depth, float* __restrict__ g_rgb, float* __restrict__ g_depth,
                double* __restrict__ partials) {
  double a_mse = 0.0, a_var = 0.0, a_dvar = 0.0, a_sr = 0.0;
  const int ss = t.ss;
  const float inv_nm1 = ss > 1 ? 1.f / (float)(ss - 1) : 0.f;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_lr; p += (int64_t)gridDim.x * blockDim.x) {
    const float* x = rgb + p * ss * 3;
    float mean[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sum = 0.f;
      for (int k = 0; k < ss; ++k) sum = __fadd_rn(sum, x[k * 3 + c]);
      mean[c] = __fdiv_rn(sum, (float)ss);
      if (lr_rgb) lr_rgb[p * 3 + c] = mean[c];
      const float d = target ? __fsub_rn(mean[c], target[p * 3 + c]) : 0.f;
      a_mse += (double)__fmul_rn(d, d);
      float m2 = 0.f;
      const float gm = t.g_mse * d;
      for (int k = 0; k < ss; ++k) {
        const float v = x[k * 3 + c];
        const float dev = v - mean[c];
        m2 = fmaf(dev, dev, m2);
        float g = gm + t.g_var * dev;
        if (target_hr) {
          const float e = v - target_hr[(p * ss + k) * 3 + c];
          a_sr += (double)(e * e);
          g = fmaf(t.g_sr, e, g);
        }
        if (g_rgb) g_rgb[(p * ss + k) * 3 + c] = g;
      }
      if (t.want_var) a_var += (double)(m2 * inv_nm1);
    }
    if (depth) {
      const float* dp = depth + p * ss;
      float sum = 0.f;
      for (int k = 0; k < ss; ++k) sum = __fadd_rn(sum, dp[k]);
      if (lr_depth) lr_depth[p] = __fdiv_rn(sum, (float)ss);
      if (t.want_dvar) {
        // torch.var(depth / far): the division comes first (:351)
        float sf = 0.f;
        for (int k = 0; k < ss; ++k) sf += __fdiv_rn(dp[k], t.far_plane);
        const float mf = __fdiv_rn(sf, (float)ss);
        float m2 = 0.f;
        for (int k = 0; k < ss; ++k) {
          const float dev = __fdiv_rn(dp[k], t.far_plane) - mf;
          m2 = fmaf(dev, dev, m2);
          if (g_depth) g_depth[p * ss + k] = t.g_dvar * dev;
        }
        a_dvar += (double)(m2 * inv_nm1);
      } else if (g_depth) {
        for (int k = 0; k < ss; ++k) g_depth[p * ss + k] = 0.f;
      }
    }
  }
  __shared__ double sh[4][256];
  sh[0][threadIdx.x] = a_mse; sh[1][threadIdx.x] = a_var; sh[2][threadIdx.x] = a_dvar; sh[3][threadIdx.x] = a_sr;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) partials[blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x][0];
}

// metrics[8] = {lambda_mse * mse, psnr, var_sum, depth_var_sum, lambda_hr * mse_hr, total, 0, 0}
__global__ void k_loss_epilogue_final(const double* __restrict__ partials, int n_blocks, int64_t n_lr, int ss, float lambda_mse,
                                      float lambda_var, float lambda_dvar, int has_lr, int has_sr, float lambda_hr,
                                      float* __restrict__ metrics) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < n_blocks; ++i)
      for (int j = 0; j < 4; ++j) s[j] += partials[i * 4 + j];
    const float mse = has_lr ? (float)(s[0] / (double)(n_lr * 3)) : 0.f;
    const float var = (float)s[1], dvar = (float)s[2];
    const float sr = has_sr ? lambda_hr * (float)(s[3] / (double)(n_lr * ss * 3)) : 0.f;
    metrics[0] = mse * lambda_mse;
    metrics[1] = has_lr ? -10.f * log10f(mse) : 0.f;
    metrics[2] = var;
    metrics[3] = dvar;
    metrics[4] = sr;
    metrics[5] = mse * lambda_mse + sr + lambda_var * var + lambda_dvar * dvar;
    metrics[6] = 0.f;
    metrics[7] = 0.f;
  }
}

__global__ void k_loss_final(const double* __restrict__ partials, int n_blocks, int64_t n_elems, float lambda,
                             float* __restrict__ metrics) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += partials[i];
    const float mse = (float)(s / (double)n_elems);
    metrics[0] = mse * lambda;                 // loss term (criterions.py:7-15 times lambda_*_mse)
    metrics[1] = -10.f * log10f(mse);          // criterions.py:36
  }
}

__global__ void __launch_bounds__(256)
k_sqnorm_partial(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ partials) {
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = a[i];
    acc += (double)x * (double)x;
    if (b) { const float y = b[i]; acc += (double)y * (double)y; }
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

__global__ void k_clip_final(const double* __restrict__ partials, int n_blocks, float max_norm, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += partials[i];
    const float total = (float)sqrt(s);
    out[0] = fminf(max_norm / (total + 1e-6f), 1.f);   // nn.utils.clip_grad_norm_
    out[1] = total;
  }
}

constexpr int kMaxParams = 40;
struct AdamArgs {
  float* p[kMaxParams];
  long long off[kMaxParams + 1];
  int n;
  const float* grad; float* m; float* v;
  float step_size, bc2_sqrt, beta1, beta2, eps;
  const float* clip_coef;   // device scalar or null
  float clip_value;         // > 0: clamp gradients to [-v, v]
};

__global__ void __launch_bounds__(256) k_adam(const AdamArgs a) {
  const long long total = a.off[a.n];
  const float coef = a.clip_coef ? *a.clip_coef : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t = 0;
    while (i >= a.off[t + 1]) ++t;
    float g = a.grad[i];
    if (a.clip_coef) g = __fmul_rn(g, coef);
    if (a.clip_value > 0.f) g = fminf(fmaxf(g, -a.clip_value), a.clip_value);
    float m = a.m[i], v = a.v[i];
    m = __fadd_rn(m, __fmul_rn(1.f - a.beta1, __fsub_rn(g, m)));                 // exp_avg.lerp_(grad, 1 - beta1)
    v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(1.f - a.beta2, g), g));   // mul_(beta2).addcmul_(g, g, 1 - beta2)
    a.m[i] = m; a.v[i] = v;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), a.bc2_sqrt), a.eps);
    float* p = a.p[t] + (i - a.off[t]);
    *p = __fadd_rn(*p, __fmul_rn(-a.step_size, __fdiv_rn(m, denom)));            // addcdiv_(m, denom, -step_size)
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int tfail(NsrHandle_* h, int code, const std::string& msg) { h->err = msg; return code; }
#define NSR_TCUDA(h, expr)                                                                    \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return tfail(h, NSR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
  } while (0)

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }
static int fmt_of(const NsrHandle_* h) { return h->cfg.precision == NSR_PREC_FP16X3_TC ? 0 : 1; }

struct TrainWs {
  // per pass (0 coarse, 1 fine)
  size_t enc[2], hh[2], dir[2], mask[2], raw[2], z[2];
  // shared by the two backward passes
  size_t dhead, encdir, dzdir, dz, dsig, part;     // dz: [9][T][4 chunks]: d feat, dZ_8 .. dZ_1
  size_t part_region;       // bytes of one dW launch's partial region
  size_t total;
  long long tiles[2];
  int S[2];
};

static TrainWs train_layout(const NsrHandle_* h, int64_t n) {
  TrainWs L{};
  L.S[0] = h->cfg.n_coarse; L.S[1] = h->cfg.n_coarse + h->cfg.n_importance;
  size_t off = 0;
  for (int w = 0; w < 2; ++w) {
    const int rpt = kT / L.S[w];
    L.tiles[w] = (n + rpt - 1) / rpt;
    const size_t t = (size_t)L.tiles[w];
    L.enc[w] = off; off += t * kChunk;
    L.hh[w] = off; off += 9 * t * 4 * kChunk;
    L.dir[w] = off; off += t * 2 * kChunk;
    L.mask[w] = off; off += 8 * t * kT * 8 * sizeof(uint32_t);
    L.raw[w] = off; off += al256((size_t)n * L.S[w] * 4 * sizeof(float));
    L.z[w] = off; off += al256((size_t)n * L.S[w] * sizeof(float));
  }
  const size_t T = (size_t)(L.tiles[0] > L.tiles[1] ? L.tiles[0] : L.tiles[1]);
  L.dhead = off; off += T * kChunk;
  L.encdir = off; off += T * kChunk;
  L.dzdir = off; off += T * 2 * kChunk;
  L.dz = off; off += 9 * T * 4 * kChunk;
  L.dsig = off; off += al256(T * kT * sizeof(float));
  L.part_region = al256((size_t)kMaxSplit * 128 * 257 * sizeof(float));
  L.part = off; off += (size_t)kDwLaunchesPerNet * L.part_region;
  L.total = off + 1024;
  return L;
}

static cudaError_t launch_dx(NsrHandle_* h, const DxArgs& a, cudaStream_t st) {
  if (a.n_tiles == 0) return cudaSuccess;
  const int grid = (int)(a.n_tiles < h->sm_count ? a.n_tiles : h->sm_count);
  cudaError_t e;
  if (fmt_of(h) == 1) {
    e = cudaFuncSetAttribute(k_tg_dx<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDxBytes);
    if (e != cudaSuccess) return e;
    k_tg_dx<1><<<grid, kGemmThreads, kSmemDxBytes, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(k_tg_dx<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDxBytes);
    if (e != cudaSuccess) return e;
    k_tg_dx<0><<<grid, kGemmThreads, kSmemDxBytes, st>>>(a);
  }
  h->launches += 1;
  return cudaGetLastError();
}

static cudaError_t launch_dw(NsrHandle_* h, const DwAllArgs& a, cudaStream_t st) {
  if (a.n_tiles == 0 || a.n_gemms == 0) return cudaSuccess;
  const int grid = h->sm_count < kMaxSplit ? h->sm_count : kMaxSplit;      // persistent: one CTA per SM
  cudaError_t e;
  if (fmt_of(h) == 1) {
    e = cudaFuncSetAttribute(k_tg_dw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemGemmBytes);
    if (e != cudaSuccess) return e;
    k_tg_dw<1><<<grid, kGemmThreads, kSmemGemmBytes, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(k_tg_dw<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemGemmBytes);
    if (e != cudaSuccess) return e;
    k_tg_dw<0><<<grid, kGemmThreads, kSmemGemmBytes, st>>>(a);
  }
  h->launches += 1;
  return cudaGetLastError();
}

// flat-gradient offsets (state_dict order)
static std::vector<long long> grad_offsets(const NsrHandle_* h) {
  std::vector<long long> off(h->param_numel.size() + 1, 0);
  for (size_t i = 0; i < h->param_numel.size(); ++i) off[i + 1] = off[i] + h->param_numel[i];
  return off;
}

// Builder for one net's dW GEMMs: collects them (and the matching reduce jobs); flush() issues the ONE persistent launch.
struct DwPlanner {
  NsrHandle_* h; cudaStream_t st; long long n_tiles; char* part_base; size_t region; int launch_idx = 0;
  RedArgs red{};
  DwAllArgs all{};
  cudaError_t err = cudaSuccess;
  struct JobSpec { const uint8_t* a_img; int a_cpt, blk0, blk1; int row0, rows; long long dst; int ld, col0; long long bias_dst; };
  void launch(const uint8_t* b_img, int b_cpt, int b_chunk0, int nB, int cols, const JobSpec* js, int n_jobs) {
    if (err != cudaSuccess) return;
    if ((region != 0 && launch_idx >= kDwLaunchesPerNet) || all.n_gemms >= kMaxDwGemms || n_jobs > kMaxDwJobs) {
      err = cudaErrorInvalidValue; return;                     // partial regions / argument table exhausted
    }
    DwGemm& a = all.g[all.n_gemms++];
    a.b_img = b_img; a.b_cpt = b_cpt; a.b_chunk0 = b_chunk0; a.nB = nB; a.n_jobs = n_jobs;
    all.n_tiles = n_tiles;
    const int ctas = h->sm_count < kMaxSplit ? h->sm_count : kMaxSplit;
    int n_split = ctas / n_jobs;
    // every split costs one [N][128] fp32 partial (written here, read back by k_grad_reduce) whatever the batch: keep at
    // least four tiles per split, so small batches (the 256-rays-per-GPU DDP shape) do not drown in partial traffic
    if ((long long)n_split * 4 > n_tiles) n_split = (int)((n_tiles + 3) / 4);
    if (n_split < 1) n_split = 1;
    a.n_split = n_split;
    a.item0 = all.n_items;
    all.n_items += n_jobs * n_split;
    const int N = 64 * nB;
    char* reg = part_base + (size_t)launch_idx * region;
    size_t off = 0;
    for (int j = 0; j < n_jobs; ++j) {
      a.job[j].a_img = js[j].a_img; a.job[j].a_cpt = js[j].a_cpt; a.job[j].a_blk0 = js[j].blk0; a.job[j].a_blk1 = js[j].blk1;
      a.job[j].part = reinterpret_cast<float*>(reg + off); off += (size_t)n_split * 128 * N * sizeof(float);
      a.job[j].part_bias = reinterpret_cast<float*>(reg + off); off += (size_t)n_split * 128 * sizeof(float);
      RedJob& R = red.job[red.n_jobs++];
      R.part = a.job[j].part; R.part_bias = a.job[j].part_bias; R.n_split = n_split; R.N = N;
      R.row0 = js[j].row0; R.rows = js[j].rows; R.cols = cols; R.ld = js[j].ld; R.col0 = js[j].col0;
      R.dst = js[j].dst; R.bias_dst = js[j].bias_dst;
    }
    ++launch_idx;
  }
  void flush() {
    if (err == cudaSuccess) err = launch_dw(h, all, st);
  }
};

// Backward of ONE net over the stash of one pass.  g_* are dL/d(outputs) of that pass.
static int backward_net(NsrHandle_* h, int which, const float* rays, int64_t n, int stride, const float* noise,
                        const float* g_rgb, const float* g_depth, const float* g_opa, float* grad_flat,
                        char* ws, const TrainWs& L, cudaStream_t st) {
  const int S = L.S[which];
  const long long tiles = L.tiles[which];
  const NetImages& net = h->net[which];
  const std::vector<long long> off = grad_offsets(h);
  uint8_t* enc = (uint8_t*)(ws + L.enc[which]);
  uint8_t* hh = (uint8_t*)(ws + L.hh[which]);
  uint8_t* dir = (uint8_t*)(ws + L.dir[which]);
  uint8_t* dhead = (uint8_t*)(ws + L.dhead);
  uint8_t* encdir = (uint8_t*)(ws + L.encdir);
  uint8_t* dzdir = (uint8_t*)(ws + L.dzdir);
  float* dsig = (float*)(ws + L.dsig);
  auto h_layer = [&](int l) { return hh + (size_t)(l - 1) * (size_t)tiles * 4 * kChunk; };   // l = 1..8 activations, 9 = feat
  auto m_layer = [&](int l) { return (const uint32_t*)(ws + L.mask[which]) + (size_t)(l - 1) * (size_t)tiles * kT * 8; };   // l = 1..8

  // ---- compositing / activation / rgb-head backward ----
  {
    RenderBwdArgs a{};
    a.rp = h->rp; a.tabs = h->d_tables; a.rays = rays; a.n_rays = n; a.ray_stride = stride;
    a.z = (const float*)(ws + L.z[which]); a.raw = (const float*)(ws + L.raw[which]);
    a.noise = (h->cfg.noise_std > 0.f) ? noise : nullptr; a.S = S;
    a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_opacity = g_opa;
    a.w_rgb = net.tc_consts + kcWrgb; a.stash_dir = dir;
    a.dhead = dhead; a.dzdir = dzdir; a.encdir = encdir; a.dsig = dsig; a.n_tiles = tiles;
    const long long slots = tiles * (kT / S);
    int grid = (int)((slots + 3) / 4);
    if (grid > h->sm_count * 8) grid = h->sm_count * 8;
    const size_t smem = (size_t)(384 + 4 * (10 * S + 32)) * sizeof(float);
    if (fmt_of(h) == 1) k_render_bwd<1><<<grid, 128, smem, st>>>(a);
    else k_render_bwd<0><<<grid, 128, smem, st>>>(a);
    h->launches += 1;
    NSR_TCUDA(h, cudaGetLastError());
  }

  const WtTable WT = build_wt_table(h->rp.ch_dir, h->cfg.no_dir != 0);
  auto wt = [&](int idx) { return net.wt_image + (size_t)WT.l[idx].stage0 * kChunk; };
  // dz(i): i = 0 d feat, 1 dZ_8, ..., 8 dZ_1 (each [tiles][4 chunks])
  auto dzi = [&](int i) { return (uint8_t*)(ws + L.dz) + (size_t)i * (size_t)tiles * 4 * kChunk; };

  // ---- the dX chain: dir -> final -> L8 .. L2 ----
  if (!(h->debug_flags & 2)) {
    ChainArgs a{};
    a.dzdir = dzdir; a.wt = net.wt_image; a.dz = dzi(0); a.mask = (const uint32_t*)(ws + L.mask[which]);
    a.dsig = dsig; a.wsig = net.tc_consts + kcWsig; a.n_tiles = tiles; a.debug_flags = h->debug_flags; a.clk = h->d_clk;
    const int grid = (int)(tiles < h->sm_count ? tiles : h->sm_count);
    if (fmt_of(h) == 1) {
      NSR_TCUDA(h, cudaFuncSetAttribute(k_tg_dxchain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemChainBytes));
      k_tg_dxchain<1><<<grid, kChainThreads, kSmemChainBytes, st>>>(a);
    } else {
      NSR_TCUDA(h, cudaFuncSetAttribute(k_tg_dxchain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemChainBytes));
      k_tg_dxchain<0><<<grid, kChainThreads, kSmemChainBytes, st>>>(a);
    }
    h->launches += 1;
    NSR_TCUDA(h, cudaGetLastError());
  } else {       // debug flag 2: the layer-by-layer kernels (k_tg_dx), kept as the cross-check of the fused chain
    for (int i = 0; i < 9; ++i) {
      DxArgs a{};
      a.a_img = (i == 0) ? dzdir : dzi(i - 1); a.nkc = (i == 0) ? 2 : 4; a.wt_img = wt(i); a.out_img = dzi(i);
      a.mask_bits = (i == 0) ? nullptr : m_layer(9 - i);
      if (i == 1) { a.dsig = dsig; a.wsig = net.tc_consts + kcWsig; }
      a.n_tiles = tiles;
      NSR_TCUDA(h, launch_dx(h, a, st));
    }
  }

  // ---- dW GEMMs ----
  DwPlanner P{h, st, tiles, ws + L.part, L.part_region};
  P.red.grad = grad_flat;
  using JS = DwPlanner::JobSpec;
  const int ld_dir = h->cfg.no_dir ? 256 : 256 + h->rp.ch_dir;
  {   // rgb.0: dW = dHead[:,1:4]^T . dir_act, db
    const JS js[1] = {{dhead, 1, 0, 0, 1, 3, off[22], 128, 0, off[23]}};
    P.launch(dir, 2, 0, 2, 128, js, 1);
  }
  {   // dir_encoding.0 (feat part, bias) and its view-direction columns
    const JS js[1] = {{dzdir, 2, 0, 1, 0, 128, off[18], ld_dir, 0, off[19]}};
    P.launch(h_layer(9), 4, 0, 4, 256, js, 1);
    if (!h->cfg.no_dir) {
      const JS jd[1] = {{dzdir, 2, 0, 1, 0, 128, off[18], ld_dir, 256, -1}};
      P.launch(encdir, 1, 0, 1, h->rp.ch_dir, jd, 1);
    }
  }
  {   // xyz_encoding_final: dW = d feat^T . h8 ; sigma head: dW = d sigma^T . h8, db
    const uint8_t* df = dzi(0);
    const JS js[3] = {{df, 4, 0, 1, 0, 128, off[16], 256, 0, off[17]},
                      {df, 4, 2, 3, 0, 128, off[16] + 128 * 256, 256, 0, off[17] + 128},
                      {dhead, 1, 0, 0, 0, 1, off[20], 256, 0, off[21]}};
    P.launch(h_layer(8), 4, 0, 4, 256, js, 3);
  }
  for (int Lyr = 8; Lyr >= 1; --Lyr) {     // trunk
    const int pw = 2 * (Lyr - 1), pb = pw + 1;
    const uint8_t* dz = dzi(9 - Lyr);
    if (Lyr == 1) {          // input = enc (63 columns)
      const JS js[2] = {{dz, 4, 0, 1, 0, 128, off[pw], 63, 0, off[pb]}, {dz, 4, 2, 3, 0, 128, off[pw] + 128 * 63, 63, 0, off[pb] + 128}};
      P.launch(enc, 1, 0, 1, 63, js, 2);
    } else if (Lyr == 5) {   // input = cat(enc, h4)
      const JS js[2] = {{dz, 4, 0, 1, 0, 128, off[pw], 319, 63, off[pb]}, {dz, 4, 2, 3, 0, 128, off[pw] + 128 * 319, 319, 63, off[pb] + 128}};
      P.launch(h_layer(4), 4, 0, 4, 256, js, 2);
      const JS je[2] = {{dz, 4, 0, 1, 0, 128, off[pw], 319, 0, -1}, {dz, 4, 2, 3, 0, 128, off[pw] + 128 * 319, 319, 0, -1}};
      P.launch(enc, 1, 0, 1, 63, je, 2);
    } else {
      const JS js[2] = {{dz, 4, 0, 1, 0, 128, off[pw], 256, 0, off[pb]}, {dz, 4, 2, 3, 0, 128, off[pw] + 128 * 256, 256, 0, off[pb] + 128}};
      P.launch(h_layer(Lyr - 1), 4, 0, 4, 256, js, 2);
    }
  }
  P.flush();
  if (P.err != cudaSuccess) return tfail(h, NSR_ERR_CUDA, std::string("dW launch: ") + cudaGetErrorString(P.err));
  {
    const dim3 grid(128, (unsigned)P.red.n_jobs);     // one element per thread for the 128 x 256 jobs
    k_grad_reduce<<<grid, 256, 0, st>>>(P.red);
    h->launches += 1;
    NSR_TCUDA(h, cudaGetLastError());
  }
  return NSR_OK;
}

}  // namespace nsr

using namespace nsr;

// ---------------------------------------------------------------------------
// C ABI (include/nsr.h, training section)
// ---------------------------------------------------------------------------
static int train_supported(NsrHandle_* h) {
  // bf16 keeps fp32's exponent range, so gradients (1e-3 .. 1e-9 here) need no loss scaling; the fp16 split
  // would (its lo plane underflows below ~6e-8), hence training is offered on the bf16x3 path only.
  if (h->cfg.precision != NSR_PREC_BF16X3_TC)
    return tfail(h, NSR_ERR_UNSUPPORTED, "training needs the bf16 split-precision tensor-core path (precision bf16x3)");
  if (h->cfg.n_importance <= 0) return tfail(h, NSR_ERR_UNSUPPORTED, "training needs N_importance > 0 (coarse + fine)");
  if (h->cfg.n_coarse != 64 || h->cfg.n_importance != 64)
    return tfail(h, NSR_ERR_UNSUPPORTED, "training is built for N_coarse 64 + N_importance 64 (whole rays per 128-point tile in both passes)");
  if (h->cfg.no_dir) return tfail(h, NSR_ERR_UNSUPPORTED, "training with --no_dir is not supported");
  if (h->cfg.W != 256) return tfail(h, NSR_ERR_UNSUPPORTED, "training is built for the 256-wide net (narrower nets render zero-padded, inference only)");
  return NSR_OK;
}

extern "C" int64_t nsr_grad_numel(const NsrHandle* h) {
  if (!h) return 0;
  int64_t s = 0;
  for (int64_t v : h->param_numel) s += v;
  return s;
}

extern "C" size_t nsr_train_workspace_bytes(const NsrHandle* h, int64_t n_rays) {
  if (!h || n_rays < 0) return 0;
  return train_layout(h, n_rays).total;
}

extern "C" int nsr_render_train(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const NsrRng* rng,
                                const NsrOutputs* out, void* train_ws, size_t train_ws_bytes, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  int rc = train_supported(h);
  if (rc) return rc;
  if (!rays || !out || n_rays <= 0 || ray_stride < 8 || ray_stride < h->cfg.viewdir_offset + 3)
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_render_train: bad argument");
  if (!h->net[0].packed || !h->net[1].packed) return tfail(h, NSR_ERR_NOT_PACKED, "weights not packed");
  const TrainWs L = train_layout(h, n_rays);
  if (!train_ws || train_ws_bytes < L.total) return tfail(h, NSR_ERR_WORKSPACE, "train workspace too small: need " + std::to_string(L.total));
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)(((uintptr_t)train_ws + 255) & ~(uintptr_t)255);
  const NsrRng none{};
  const NsrRng& R = rng ? *rng : none;
  const bool noisy = h->cfg.noise_std > 0.f;
  for (int w = 0; w < 2; ++w) {
    TcPassArgs a{};
    a.rays = rays; a.n_rays = n_rays; a.ray_stride = ray_stride; a.S = L.S[w];
    a.z_in = w ? (const float*)(ws + L.z[1]) : nullptr;
    a.u_jitter = w ? nullptr : R.u_coarse;
    a.noise = noisy ? (w ? R.noise_fine : R.noise_coarse) : nullptr;
    a.u_resample = w ? nullptr : R.u_fine;
    a.do_resample = w ? 0 : 1;
    a.comp_rgb = w ? out->fine_comp_rgbs : out->coarse_comp_rgbs;
    a.depth = w ? out->fine_depth : out->coarse_depth;
    a.opacity = w ? out->fine_opacity : out->coarse_opacity;
    a.weights = w ? out->fine_weights : out->coarse_weights;
    a.raw = (float*)(ws + L.raw[w]);
    a.z_next = w ? nullptr : (float*)(ws + L.z[1]);
    a.z_out = w ? nullptr : (float*)(ws + L.z[0]);
    a.stash_enc = (uint8_t*)(ws + L.enc[w]); a.stash_h = (uint8_t*)(ws + L.hh[w]); a.stash_dir = (uint8_t*)(ws + L.dir[w]);
    a.stash_mask = (uint32_t*)(ws + L.mask[w]);
    a.trace = nullptr; a.debug_flags = 0;
    NSR_TCUDA(h, tc_pass(h, w, a, st));
  }
  // remember (host side, no sync) which workspaces hold a stash and for how many rays, so nsr_backward can refuse
  // a buffer that was filled for another batch -- the layout depends on n_rays
  {
    auto& v = h->train_stash;
    for (size_t i = 0; i < v.size(); ++i) if (v[i].first == train_ws) { v.erase(v.begin() + i); break; }
    if (v.size() >= 16) v.erase(v.begin());
    v.emplace_back(train_ws, n_rays);
  }
  if (out->z_fine)
    NSR_TCUDA(h, cudaMemcpyAsync(out->z_fine, ws + L.z[1], (size_t)n_rays * L.S[1] * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return NSR_OK;
}

extern "C" int nsr_backward(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const NsrRng* rng,
                            const NsrOutGrads* g, float* grad_coarse, float* grad_fine, void* train_ws,
                            size_t train_ws_bytes, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  int rc = train_supported(h);
  if (rc) return rc;
  if (!rays || !g || !grad_coarse || !grad_fine || n_rays <= 0) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_backward: bad argument");
  const TrainWs L = train_layout(h, n_rays);
  if (!train_ws || train_ws_bytes < L.total) return tfail(h, NSR_ERR_WORKSPACE, "train workspace too small: need " + std::to_string(L.total));
  {
    int64_t stashed = -1;
    for (const auto& e : h->train_stash) if (e.first == train_ws) stashed = e.second;
    if (stashed != n_rays)
      return tfail(h, NSR_ERR_INVALID_ARG, stashed < 0 ? "nsr_backward: train_ws was not filled by nsr_render_train"
                                                       : "nsr_backward: train_ws holds a stash for " + std::to_string(stashed) + " rays");
  }
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)(((uintptr_t)train_ws + 255) & ~(uintptr_t)255);
  const NsrRng none{};
  const NsrRng& R = rng ? *rng : none;
  const int64_t numel = nsr_grad_numel(h);
  // a net takes part as soon as ANY of its outputs carries a gradient (depth / opacity regularisers without a colour term)
  if (g->coarse_comp_rgbs || g->coarse_depth || g->coarse_opacity) {
    rc = backward_net(h, 0, rays, n_rays, ray_stride, R.noise_coarse, g->coarse_comp_rgbs, g->coarse_depth, g->coarse_opacity,
                      grad_coarse, ws, L, st);
    if (rc) return rc;
  } else {
    NSR_TCUDA(h, cudaMemsetAsync(grad_coarse, 0, (size_t)numel * sizeof(float), st));
  }
  if (g->fine_comp_rgbs || g->fine_depth || g->fine_opacity) {
    rc = backward_net(h, 1, rays, n_rays, ray_stride, R.noise_fine, g->fine_comp_rgbs, g->fine_depth, g->fine_opacity,
                      grad_fine, ws, L, st);
    if (rc) return rc;
  } else {
    NSR_TCUDA(h, cudaMemsetAsync(grad_fine, 0, (size_t)numel * sizeof(float), st));
  }
  return NSR_OK;
}

extern "C" int nsr_lr_loss_grad(NsrHandle* h, const float* hr_rgb, const float* target_lr, int64_t n_lr, int s, float lambda,
                                float* lr_rgb_out, float* metrics_out, float* g_hr_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!hr_rgb || !target_lr || !metrics_out || n_lr <= 0 || s < 1) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_lr_loss_grad: bad argument");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  int64_t blocks64 = (n_lr * 3 + 255) / 256;
  const int blocks = (int)(blocks64 > 1024 ? 1024 : blocks64);
  // d/d hr of lambda * mean((lr - t)^2): 2 (lr - t) lambda / (3 n_lr) / s^2
  const float scale = (float)(2.0 * (double)lambda / ((double)n_lr * 3.0) / (double)(s * s));
  k_lr_loss_grad<<<blocks, 256, 0, (cudaStream_t)stream>>>(hr_rgb, target_lr, n_lr, s * s, scale, lr_rgb_out, g_hr_out, h->d_partials);
  k_loss_final<<<1, 32, 0, (cudaStream_t)stream>>>(h->d_partials, blocks, n_lr * 3, lambda, metrics_out);
  h->launches += 2;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_loss_epilogue(NsrHandle* h, const float* hr_rgb, const float* hr_depth, const float* target_lr,
                                 const float* target_hr, int64_t n_lr, const NsrLossTerms* terms, float* lr_rgb_out,
                                 float* lr_depth_out, float* metrics_out, float* g_rgb_out, float* g_depth_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!hr_rgb || !metrics_out || !terms || n_lr <= 0) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_loss_epilogue: bad argument");
  if (!target_lr && !target_hr) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_loss_epilogue: needs target_lr and / or target_hr");
  if (terms->struct_size != sizeof(NsrLossTerms)) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_loss_epilogue: NsrLossTerms.struct_size mismatch");
  const int s = terms->s;
  if (s < 1 || s > 16) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_loss_epilogue: downscale must be in [1,16]");
  const bool want_var = terms->lambda_var != 0.f, want_dvar = terms->lambda_depth_var != 0.f;
  if ((want_var || want_dvar) && s < 2)
    return tfail(h, NSR_ERR_UNSUPPORTED, "nsr_loss_epilogue: the sub-pixel variance terms need downscale >= 2 (unbiased variance of one sample is NaN)");
  if (want_dvar && (!hr_depth || !(terms->far_plane != 0.f)))
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_loss_epilogue: the depth-variance term needs hr_depth and a non-zero far plane");
  if ((lr_depth_out || g_depth_out) && !hr_depth) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_loss_epilogue: depth outputs need hr_depth");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  const int ss = s * s;
  LossTerms t;
  t.ss = ss;
  t.g_mse = target_lr ? (float)(2.0 * (double)terms->lambda_mse / ((double)n_lr * 3.0) / (double)ss) : 0.f;
  t.g_var = want_var ? (float)(2.0 * (double)terms->lambda_var / (double)(ss - 1)) : 0.f;
  t.g_dvar = want_dvar ? (float)(2.0 * (double)terms->lambda_depth_var / (double)(ss - 1) / (double)terms->far_plane) : 0.f;
  t.g_sr = target_hr ? (float)(2.0 * (double)terms->lambda_hr / ((double)n_lr * (double)ss * 3.0)) : 0.f;
  t.far_plane = terms->far_plane;
  t.want_var = want_var;
  t.want_dvar = want_dvar;
  int64_t blocks64 = (n_lr + 255) / 256;
  const int blocks = (int)(blocks64 > 256 ? 256 : blocks64);      // 4 partials per block in the 1024-entry buffer
  k_loss_epilogue<<<blocks, 256, 0, (cudaStream_t)stream>>>(hr_rgb, hr_depth, target_lr, target_hr, n_lr, t, lr_rgb_out, lr_depth_out,
                                                            g_rgb_out, g_depth_out, h->d_partials);
  k_loss_epilogue_final<<<1, 32, 0, (cudaStream_t)stream>>>(h->d_partials, blocks, n_lr, ss, terms->lambda_mse, terms->lambda_var,
                                                            terms->lambda_depth_var, target_lr != nullptr, target_hr != nullptr,
                                                            terms->lambda_hr, metrics_out);
  h->launches += 2;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_clip_coef(NsrHandle* h, const float* grad_a, const float* grad_b, int64_t numel, float max_norm,
                             float* coef_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!grad_a || !coef_out || numel <= 0 || !(max_norm > 0.f)) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_clip_coef: bad argument");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  const int blocks = 592;
  k_sqnorm_partial<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad_a, grad_b, numel, h->d_partials);
  k_clip_final<<<1, 32, 0, (cudaStream_t)stream>>>(h->d_partials, blocks, max_norm, coef_out);
  h->launches += 2;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_adam_step(NsrHandle* h, float* const* param_ptrs, int n_params, const float* grad_flat, float* exp_avg,
                             float* exp_avg_sq, int64_t step, float lr, float beta1, float beta2, float eps,
                             const float* clip_coef_dev, float clip_value, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!param_ptrs || !grad_flat || !exp_avg || !exp_avg_sq || step < 1) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_adam_step: bad argument");
  if (n_params != (int)h->param_numel.size() || n_params > kMaxParams)
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_adam_step: expected " + std::to_string(h->param_numel.size()) + " parameter tensors");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  AdamArgs a{};
  a.n = n_params;
  a.off[0] = 0;
  for (int i = 0; i < n_params; ++i) {
    if (!param_ptrs[i]) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_adam_step: null parameter pointer");
    a.p[i] = param_ptrs[i];
    a.off[i + 1] = a.off[i] + h->param_numel[i];
  }
  a.grad = grad_flat; a.m = exp_avg; a.v = exp_avg_sq;
  // torch.optim.Adam (_single_tensor_adam): python-float scalars, rounded to fp32 at the tensor ops
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.bc2_sqrt = (float)std::sqrt(bc2);
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.clip_coef = clip_coef_dev; a.clip_value = clip_value;
  const long long total = a.off[n_params];
  int blocks = (int)((total + 255) / 256);
  if (blocks > h->sm_count * 8) blocks = h->sm_count * 8;
  k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  h->launches += 1;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

// ---- test seams ----------------------------------------------------------------
extern "C" int nsr_debug_train_layout(const NsrHandle* h, int64_t n_rays, int64_t* out16) {
  if (!h || !out16 || n_rays <= 0) return NSR_ERR_INVALID_ARG;
  const TrainWs L = train_layout(h, n_rays);
  const int64_t v[18] = {(int64_t)L.enc[0], (int64_t)L.hh[0], (int64_t)L.dir[0], (int64_t)L.raw[0], (int64_t)L.z[0], L.tiles[0],
                         (int64_t)L.enc[1], (int64_t)L.hh[1], (int64_t)L.dir[1], (int64_t)L.raw[1], (int64_t)L.z[1], L.tiles[1],
                         (int64_t)L.dhead, (int64_t)L.dzdir, (int64_t)L.dz, (int64_t)(L.dz + 4 * (size_t)kChunk * (size_t)(L.tiles[0] > L.tiles[1] ? L.tiles[0] : L.tiles[1])), (int64_t)L.mask[0], (int64_t)L.mask[1]};
  for (int i = 0; i < 18; ++i) out16[i] = v[i];
  return NSR_OK;
}

extern "C" int nsr_debug_pack_image(NsrHandle* h, const float* src, int64_t n_rows, int n_cols, int ld, void* image, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!src || !image || n_rows <= 0 || n_cols <= 0 || n_cols % 64 || ld <= 0 || ld > n_cols)
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_debug_pack_image: bad argument");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  if (fmt_of(h) == 1) k_pack_image<1><<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(src, n_rows, n_cols, ld, (uint8_t*)image);
  else k_pack_image<0><<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(src, n_rows, n_cols, ld, (uint8_t*)image);
  h->launches += 1;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_debug_unpack_image(NsrHandle* h, const void* image, int64_t n_rows, int n_cols, int ld, float* dst, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!dst || !image || n_rows <= 0 || n_cols <= 0 || n_cols % 64 || ld <= 0 || ld > n_cols)
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_debug_unpack_image: bad argument");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  if (fmt_of(h) == 1) k_unpack_image<1><<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)image, n_rows, n_cols, ld, dst);
  else k_unpack_image<0><<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)image, n_rows, n_cols, ld, dst);
  h->launches += 1;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

// out_img[rows,256] = mask . (a_img[rows, k_cols] . W[k_cols, ld][:, col0:col0+256] + dsig x wsig)
// with W given as the transposed-weight image of net `which`, dX layer `layer_idx` (0 dir, 1 final, 2.. = L8..L2).
__global__ void k_relu_bits(const float* __restrict__ x, long long n_rows, uint32_t* __restrict__ bits) {
  // bits[(row)*8 + w] bit j = x[row][32 w + j] > 0   (rows padded to a multiple of 128 with zeros)
  const long long rows_pad = (n_rows + 127) / 128 * 128;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows_pad * 8; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / 8;
    const int w = (int)(i % 8);
    uint32_t b = 0u;
    if (row < n_rows) for (int j = 0; j < 32; ++j) b |= (x[row * 256 + 32 * w + j] > 0.f ? 1u : 0u) << j;
    bits[i] = b;
  }
}

extern "C" int nsr_debug_relu_bits(NsrHandle* h, const float* x, int64_t n_rows, void* bits_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!x || !bits_out || n_rows <= 0) return tfail(h, NSR_ERR_INVALID_ARG, "nsr_debug_relu_bits: bad argument");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  k_relu_bits<<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(x, n_rows, (uint32_t*)bits_out);
  h->launches += 1;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_debug_dx(NsrHandle* h, int which, int layer_idx, const void* a_img, void* out_img, const void* mask_bits,
                            const float* dsig, const float* wsig, int64_t n_rows, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  int rc = train_supported(h);
  if (rc) return rc;
  if (which < 0 || which > 1 || layer_idx < 0 || layer_idx >= kNumDxLayers || !a_img || !out_img || n_rows <= 0)
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_debug_dx: bad argument");
  if (!h->net[which].packed) return tfail(h, NSR_ERR_NOT_PACKED, "weights not packed");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  const WtTable WT = build_wt_table(h->rp.ch_dir, h->cfg.no_dir != 0);
  DxArgs a{};
  a.a_img = (const uint8_t*)a_img; a.nkc = WT.l[layer_idx].n_out / 64;
  a.wt_img = h->net[which].wt_image + (size_t)WT.l[layer_idx].stage0 * kChunk;
  a.out_img = (uint8_t*)out_img; a.mask_bits = (const uint32_t*)mask_bits; a.dsig = dsig; a.wsig = wsig;
  a.n_tiles = (n_rows + kT - 1) / kT;
  NSR_TCUDA(h, launch_dx(h, a, (cudaStream_t)stream));
  return NSR_OK;
}

// out[128, b_cols] = a_img[rows, blocks blk0|blk1]^T . b_img[rows, b_cols];  bias_out[128] = column sums of A.
// scratch: >= 148*128*(b_cols+1)*4 bytes.
extern "C" int nsr_debug_dw(NsrHandle* h, const void* a_img, int a_cols, int blk0, int blk1, const void* b_img, int b_cols,
                            float* out, float* bias_out, int64_t n_rows, void* scratch, size_t scratch_bytes, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  int rc = train_supported(h);
  if (rc) return rc;
  if (!a_img || !b_img || !out || !bias_out || !scratch || a_cols % 64 || b_cols % 64 || b_cols < 64 || b_cols > 256 || n_rows <= 0 ||
      blk0 < 0 || blk1 < 0 || blk0 >= a_cols / 64 || blk1 >= a_cols / 64)
    return tfail(h, NSR_ERR_INVALID_ARG, "nsr_debug_dw: bad argument");
  if (scratch_bytes < (size_t)kMaxSplit * 128 * (b_cols + 1) * sizeof(float)) return tfail(h, NSR_ERR_WORKSPACE, "nsr_debug_dw: scratch too small");
  NSR_TCUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  DwPlanner P{h, st, (n_rows + kT - 1) / kT, (char*)scratch, 0};
  // flat "gradient": out rows then bias
  P.red.grad = out;
  const DwPlanner::JobSpec js[1] = {{(const uint8_t*)a_img, a_cols / 64, blk0, blk1, 0, 128, 0, b_cols, 0, -1}};
  P.launch((const uint8_t*)b_img, b_cols / 64, 0, b_cols / 64, b_cols, js, 1);
  P.flush();
  if (P.err != cudaSuccess) return tfail(h, NSR_ERR_CUDA, std::string("dW launch: ") + cudaGetErrorString(P.err));
  k_grad_reduce<<<dim3(32, 1), 256, 0, st>>>(P.red);
  RedArgs rb = P.red;           // second pass: the bias partials into bias_out
  rb.grad = bias_out;
  rb.job[0].rows = 128; rb.job[0].cols = 0; rb.job[0].bias_dst = 0;
  k_grad_reduce<<<dim3(1, 1), 256, 0, st>>>(rb);
  h->launches += 2;
  NSR_TCUDA(h, cudaGetLastError());
  return NSR_OK;
}
