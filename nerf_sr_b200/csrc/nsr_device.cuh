// nsr_device.cuh -- device functions shared by the SIMT and the tcgen05 paths:
// coarse sampling, point casting, positional encoding, alpha compositing and
// inverse-CDF resampling.  Scalar fp32 math follows the reference's operation
// ORDER (separate multiply / add roundings where PyTorch issues separate
// elementwise kernels), because the positional encoding amplifies 1-ulp
// differences in z by up to 2^9 (hard part 5).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsr {

constexpr int kMaxSamples = 256;     // per-ray samples the fused paths support
constexpr int kMaxFreqs = 16;

// ----------------------------------------------------------------------------
// Small per-handle tables passed by value in kernel parameters (no global state)
// ----------------------------------------------------------------------------
struct SampleTables {
  float t_coarse[kMaxSamples];       // torch.linspace(0,1,n_coarse)    (models/utils.py:31)
  float one_minus_t[kMaxSamples];    // 1 - t, rounded once like the reference's (1 - t_vals)
  float u_fine[kMaxSamples];         // torch.linspace(0,1,n_importance) (models/utils.py:75)
  float freq_pos[kMaxFreqs];         // frequency bands (models/embedding.py:39-42)
  float freq_dir[kMaxFreqs];
};

struct RenderParams {
  int n_coarse, n_importance;
  int deg_pos, deg_dir, no_xyz;
  int ch_pos, ch_dir;
  int lindisp, white_bkgd, sigma_softplus, color_none, gamma_correct, no_dir;
  int viewdir_offset;
  float noise_std;
};

// A posed pinhole camera and its supersampling raster (nsr_generate_rays / nsr_render_pose_host)
struct RayGenParams {
  float m[12];                       // c2w, 3 x 4 row-major
  int H, W;                          // HR raster
  float focal;
  int s, ndc;
  float near_plane, far_plane;
  float pixel_center;                // 0.5 (use_pixel_centers) or 0
  int unified_dir;
};

// ----------------------------------------------------------------------------
// a6: coarse z for sample i of a ray (models/utils.py:31-35)
// ----------------------------------------------------------------------------
__device__ __forceinline__ float coarse_z(float near, float far, float t, float omt, int lindisp) {
  if (lindisp) {
    // 1. / (1. / near * (1 - t) + 1. / far * t)
    float a = __fmul_rn(__fdiv_rn(1.f, near), omt);
    float b = __fmul_rn(__fdiv_rn(1.f, far), t);
    return __fdiv_rn(1.f, __fadd_rn(a, b));
  }
  // near * (1 - t) + far * t
  return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
}

// stratified jitter (models/utils.py:37-41): z = lower + u * (upper - lower)
__device__ __forceinline__ float jitter_z(float z_prev, float z_cur, float z_next, bool first,
                                          bool last, float u) {
  float lower = first ? z_cur : __fmul_rn(0.5f, __fadd_rn(z_prev, z_cur));
  float upper = last ? z_cur : __fmul_rn(0.5f, __fadd_rn(z_cur, z_next));
  return __fadd_rn(lower, __fmul_rn(u, __fsub_rn(upper, lower)));
}

// cast_rays (models/utils.py:14): o + z * d, multiply then add (no FMA)
__device__ __forceinline__ float cast_point(float o, float d, float z) {
  return __fadd_rn(o, __fmul_rn(z, d));
}

// ----------------------------------------------------------------------------
// a5: positional encoding of one 3-vector (models/embedding.py:57-63).
// Calls emit(channel, value) for channels in the reference's concatenation
// order: [x] + [sin(f0 x), cos(f0 x), sin(f1 x), ...], each a 3-vector.
// ----------------------------------------------------------------------------
template <typename Emit>
__device__ __forceinline__ void posenc3(float x, float y, float z, int n_freqs, const float* freqs,
                                        int no_xyz, Emit emit) {
  int c = 0;
  if (!no_xyz) {
    emit(0, x); emit(1, y); emit(2, z);
    c = 3;
  }
  for (int k = 0; k < n_freqs; ++k) {
    const float f = freqs[k];
    const float ax = __fmul_rn(f, x), ay = __fmul_rn(f, y), az = __fmul_rn(f, z);
    float sx, cx, sy, cy, sz, cz;     // one range reduction per argument (same results as sinf/cosf)
    sincosf(ax, &sx, &cx); sincosf(ay, &sy, &cy); sincosf(az, &sz, &cz);
    emit(c + 0, sx); emit(c + 1, sy); emit(c + 2, sz);
    emit(c + 3, cx); emit(c + 4, cy); emit(c + 5, cz);
    c += 6;
  }
}

// ----------------------------------------------------------------------------
// warp helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigma_act(float s, int softplus) {
  // rendering.py:70-73: relu, or log(1 + exp(x - 1))
  return softplus ? logf(__fadd_rn(1.f, expf(__fsub_rn(s, 1.f)))) : fmaxf(s, 0.f);
}

// ----------------------------------------------------------------------------
// Sequential scan of x[0..n) in shared memory by ONE lane, in index order (the
// rounding order of torch.cumsum / torch.cumprod on CPU).  Loads are batched 8
// ahead of the dependent chain, so the cost is ~4 cycles per element.
//   MUL = true : x[i] <- init * prod_{j<i} x[j]   (exclusive product)
//   MUL = false: x[i] <- init + sum_{j<=i} x[j]   (inclusive sum)
// ----------------------------------------------------------------------------
template <bool MUL>
__device__ __forceinline__ void seq_scan_one_lane(float* x, int n, float init) {
  float acc = init;
  int base = 0;
  for (; base + 8 <= n; base += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = x[base + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MUL) { const float t = acc; acc = __fmul_rn(acc, v[j]); v[j] = t; }
      else { acc = __fadd_rn(acc, v[j]); v[j] = acc; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) x[base + j] = v[j];
  }
  for (; base < n; ++base) {
    const float v = x[base];
    if (MUL) { x[base] = acc; acc = __fmul_rn(acc, v); }
    else { acc = __fadd_rn(acc, v); x[base] = acc; }
  }
}

// ----------------------------------------------------------------------------
// a10: alpha compositing of ONE ray by ONE warp (models/rendering.py:89-111).
//   z[S], sigma[S] (raw, noise already added), rgb[S*3] : shared memory
//   w_out[S] : shared memory (weights kept for resampling); tmp[S] : scratch
// Results (comp rgb, depth, opacity) are returned in every lane.
// alpha and (1 - alpha + 1e-10) are evaluated in parallel; the transmittance
// product itself is accumulated in sample order like torch.cumprod on CPU.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void composite_ray_warp(const float* z, const float* sigma, const float* rgb,
                                                   int S, int white_bkgd, int softplus, float* w_out,
                                                   float* tmp, float& r, float& g, float& b, float& depth,
                                                   float& opacity) {
  const int lane = threadIdx.x & 31;
  // alpha_i = 1 - exp(-delta_i * act(sigma_i)); delta_last = 1e10
  for (int i = lane; i < S; i += 32) {
    const float delta = (i + 1 < S) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
    const float a = __fsub_rn(1.f, expf(__fmul_rn(-delta, sigma_act(sigma[i], softplus))));
    w_out[i] = a;
    tmp[i] = __fadd_rn(__fsub_rn(1.f, a), 1e-10f);     // 1 - alpha + eps   (rendering.py:101)
  }
  __syncwarp();
  if (lane == 0) seq_scan_one_lane<true>(tmp, S, 1.f);   // T_i = prod_{j<i}(1 - alpha_j + eps), T_0 = 1
  __syncwarp();
  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, so = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float w = __fmul_rn(w_out[i], tmp[i]);
    w_out[i] = w;
    sr += w * rgb[3 * i + 0];
    sg += w * rgb[3 * i + 1];
    sb += w * rgb[3 * i + 2];
    sd += w * z[i];
    so += w;
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); so = warp_sum(so);
  if (white_bkgd) {
    const float bg = __fsub_rn(1.f, so);
    sr = __fadd_rn(sr, bg); sg = __fadd_rn(sg, bg); sb = __fadd_rn(sb, bg);
  }
  r = sr; g = sg; b = sb; depth = sd; opacity = so;
  __syncwarp();
}

// ----------------------------------------------------------------------------
// a11: inverse-CDF resampling + sort-merge of ONE ray by ONE warp
// (models/utils.py:61-93).
//   z[Sc], w[Sc]      : shared memory (coarse z-values and weights)
//   u                 : pointer to this ray's n_imp uniforms (global) or null
//   u_table           : linspace(0,1,n_imp) when u == null
//   scratch           : shared memory, >= 2*Sc + n_imp floats
//   z_out[Sc + n_imp] : shared or global memory; sorted ascending
// ----------------------------------------------------------------------------
__device__ __forceinline__ void resample_ray_warp(const float* z, const float* w, int Sc, int n_imp,
                                                  const float* u, const float* u_table, float* scratch,
                                                  float* z_out) {
  const int lane = threadIdx.x & 31;
  const float eps = 1e-5f;
  const int nb = Sc - 1;         // bins (mid-points) and cdf entries
  const int nw = Sc - 2;         // interior weights
  float* bins = scratch;         // [Sc-1]
  float* cdf = scratch + Sc;     // [Sc-1]
  float* znew = scratch + 2 * Sc;  // [n_imp]
  float part = 0.f;
  for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(z[i], z[i + 1]));
  for (int i = lane; i < nw; i += 32) part += __fadd_rn(w[i + 1], eps);
  const float total = warp_sum(part);
  // pdf in parallel, then cdf = cat(0, cumsum(pdf)) accumulated in order like torch.cumsum on CPU
  for (int i = lane; i < nw; i += 32) cdf[i + 1] = __fdiv_rn(__fadd_rn(w[i + 1], eps), total);
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  if (lane == 0) seq_scan_one_lane<false>(cdf + 1, nw, 0.f);
  __syncwarp();
  for (int j = lane; j < n_imp; j += 32) {
    const float uj = u ? u[j] : u_table[j];
    int inds = 0;                                   // searchsorted(cdf, u, right=True)
    for (int i = 0; i < nb; ++i) inds += (cdf[i] <= uj) ? 1 : 0;
    const int below = max(inds - 1, 0);
    const int above = min(inds, nw);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = bins[below], b1 = bins[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < eps) denom = 1.f;
    znew[j] = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(uj, c0), denom), __fsub_rn(b1, b0)));
  }
  __syncwarp();
  // sort(cat(z, z_new)) by rank: stable w.r.t. concatenation order, handles the
  // unsorted u of train mode as well as the two pre-sorted lists of eval mode.
  const int n_all = Sc + n_imp;
  for (int e = lane; e < n_all; e += 32) {
    const float v = (e < Sc) ? z[e] : znew[e - Sc];
    int rank = 0;
    for (int i = 0; i < Sc; ++i) {
      const float o = z[i];
      rank += (o < v || (o == v && i < e)) ? 1 : 0;
    }
    for (int i = 0; i < n_imp; ++i) {
      const float o = znew[i];
      rank += (o < v || (o == v && (Sc + i) < e)) ? 1 : 0;
    }
    z_out[rank] = v;
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------
// a0: one ray of a posed pinhole camera: get_ray_directions + get_rays (+ get_ndc_rays) in the dataset's
// '(h s1) (w s2) c -> (h w) (s1 s2) c' row order (models/utils.py:98-196; data/blender_downX_dataset.py:207-215).
// `idx` is the OUTPUT row: LR pixel (h, w), sub-pixel (s1, s2).  out = (o[3], d[3], near, far).
// Shared by k_generate_rays (nsr_api.cu) and the fused frame kernel's front-end (nsr_tc.cu): same arithmetic, same bits.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void generate_ray(const RayGenParams& g, int64_t idx, float* out) {
  const int s = g.s, W = g.W, H = g.H;
  const int w_lr = W / s;
  const int sub = (int)(idx % (s * s));
  const int64_t lr = idx / (s * s);
  const int hh = (int)(lr / w_lr), ww = (int)(lr % w_lr);
  const int s1 = sub / s, s2 = sub % s;
  const int row = hh * s + s1, col = ww * s + s2;
  // --unified_dir (data/llff_downX_dataset.py:273-277): one camera-space direction per LR pixel, computed on the
  // (H/s, W/s) raster with focal // s and repeated over its s x s sub-pixels; NDC below still uses H, W, focal
  const int dcol = g.unified_dir ? ww : col, drow = g.unified_dir ? hh : row;
  const float dW = g.unified_dir ? (float)(W / s) : (float)W, dH = g.unified_dir ? (float)(H / s) : (float)H;
  const float dfocal = g.unified_dir ? floorf(__fdiv_rn(g.focal, (float)s)) : g.focal;
  const float i = (float)dcol + g.pixel_center, j = (float)drow + g.pixel_center;
  const float cx = __fdiv_rn(__fsub_rn(i, dW / 2.f), dfocal);
  const float cy = -__fdiv_rn(__fsub_rn(j, dH / 2.f), dfocal);
  const float cz = -1.f;
  // rays_d = directions @ c2w[:, :3].T
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    d[k] = __fadd_rn(__fadd_rn(__fmul_rn(cx, g.m[4 * k + 0]), __fmul_rn(cy, g.m[4 * k + 1])), __fmul_rn(cz, g.m[4 * k + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  d[0] = __fdiv_rn(d[0], nrm); d[1] = __fdiv_rn(d[1], nrm); d[2] = __fdiv_rn(d[2], nrm);
  float o[3] = {g.m[3], g.m[7], g.m[11]};
  float nearv = g.near_plane, farv = g.far_plane;
  if (g.ndc) {   // get_ndc_rays at near = 1.0 (data/llff_downX_dataset.py:476-481)
    const float nr = 1.0f;
    const float t = __fdiv_rn(-__fadd_rn(nr, o[2]), d[2]);
    o[0] = __fadd_rn(o[0], __fmul_rn(t, d[0]));
    o[1] = __fadd_rn(o[1], __fmul_rn(t, d[1]));
    o[2] = __fadd_rn(o[2], __fmul_rn(t, d[2]));
    const float ox_oz = __fdiv_rn(o[0], o[2]), oy_oz = __fdiv_rn(o[1], o[2]);
    const float kx = (float)(-1.0 / ((double)W / (2.0 * (double)g.focal)));
    const float ky = (float)(-1.0 / ((double)H / (2.0 * (double)g.focal)));
    const float o0 = __fmul_rn(kx, ox_oz), o1 = __fmul_rn(ky, oy_oz);
    const float o2 = __fadd_rn(1.f, __fdiv_rn(__fmul_rn(2.f, nr), o[2]));
    const float d0 = __fmul_rn(kx, __fsub_rn(__fdiv_rn(d[0], d[2]), ox_oz));
    const float d1 = __fmul_rn(ky, __fsub_rn(__fdiv_rn(d[1], d[2]), oy_oz));
    const float d2 = __fsub_rn(1.f, o2);
    o[0] = o0; o[1] = o1; o[2] = o2; d[0] = d0; d[1] = d1; d[2] = d2;
    nearv = 0.f; farv = 1.f;
  }
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = d[0]; out[4] = d[1]; out[5] = d[2];
  out[6] = nearv; out[7] = farv;
}

}  // namespace nsr
