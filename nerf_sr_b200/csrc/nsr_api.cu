// nsr_api.cu -- C ABI of libnsr_b200 (see include/nsr.h) and the small kernels
// either side of the MLP: coarse sampling, compositing + inverse-CDF
// resampling (one warp per ray), positional encoding, box average, on-device
// ray generation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "nsr_internal.h"
#include "nsr_jet_lut.h"

using namespace nsr;

// ----------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------
static thread_local std::string g_create_err = "";

static int fail(NsrHandle_* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}
#define NSR_CUDA(h, expr)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(h, NSR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
  } while (0)

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// torch.linspace for fp32 (ATen RangeFactories: step = (end-start)/(steps-1); first half counts
// up from start, second half counts down from end; the multiply-add is FUSED in ATen's CPU and
// CUDA kernels, which changes the last bit of 23 of the 64 coarse t-values -- verified bit-exact
// against torch.linspace in tests/test_gpu_parity.py::test_sampling_seams).
static void host_linspace(float start, float end, int steps, float* out) {
  if (steps == 1) { out[0] = start; return; }
  const float step = (end - start) / (float)(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i) {
    if (i < half) out[i] = fmaf(step, (float)i, start);
    else out[i] = fmaf(-step, (float)(steps - i - 1), end);
  }
}

// ----------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------
__global__ void k_sample_coarse(RenderParams rp, const SampleTables* __restrict__ tabs,
                                const float* __restrict__ rays, int64_t n_rays, int ray_stride,
                                const float* __restrict__ u, float* __restrict__ z_out) {
  const int S = rp.n_coarse;
  const int64_t total = n_rays * S;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / S;
    const int i = (int)(idx % S);
    const float near = rays[r * ray_stride + 6], far = rays[r * ray_stride + 7];
    const float zc = coarse_z(near, far, tabs->t_coarse[i], tabs->one_minus_t[i], rp.lindisp);
    float zv = zc;
    if (u) {
      const float zp = i > 0 ? coarse_z(near, far, tabs->t_coarse[i - 1], tabs->one_minus_t[i - 1], rp.lindisp) : zc;
      const float zn = i + 1 < S ? coarse_z(near, far, tabs->t_coarse[i + 1], tabs->one_minus_t[i + 1], rp.lindisp) : zc;
      zv = jitter_z(zp, zc, zn, i == 0, i + 1 == S, u[idx]);
    }
    z_out[idx] = zv;
  }
}

// One warp per ray: (raw rgb/sigma, z) -> composite (+ resample).
__global__ void __launch_bounds__(128)
k_composite(RenderParams rp, const SampleTables* __restrict__ tabs, const float* __restrict__ raw,
            const float* __restrict__ z, const float* __restrict__ noise, int64_t n_rays, int S,
            int do_resample, const float* __restrict__ u_resample, float* __restrict__ comp_rgb,
            float* __restrict__ depth, float* __restrict__ opacity, float* __restrict__ weights,
            float* __restrict__ z_next, int per_warp_floats) {
  extern __shared__ float csm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* base = csm + warp * per_warp_floats;
  float* sz = base;            // [S]
  float* ssig = sz + S;        // [S]
  float* srgb = ssig + S;      // [3S]
  float* sw = srgb + 3 * S;    // [S]
  float* stmp = sw + S;        // [S]
  float* scratch = stmp + S;   // resample scratch [2S + n_imp] then z_out [S + n_imp]
  const int n_imp = rp.n_importance;
  for (int64_t ray = blockIdx.x * 4 + warp; ray < n_rays; ray += (int64_t)gridDim.x * 4) {
    const float4* r4 = reinterpret_cast<const float4*>(raw) + ray * S;
    for (int i = lane; i < S; i += 32) {
      const float4 v = r4[i];
      float cr = v.x, cg = v.y, cb = v.z;
      if (rp.gamma_correct) { cr = powf(cr, 1.f / 2.2f); cg = powf(cg, 1.f / 2.2f); cb = powf(cb, 1.f / 2.2f); }
      srgb[3 * i] = cr; srgb[3 * i + 1] = cg; srgb[3 * i + 2] = cb;
      float s = v.w;
      if (noise) s = __fadd_rn(s, __fmul_rn(noise[ray * S + i], rp.noise_std));   // utils.py:210
      ssig[i] = s;
      sz[i] = z[ray * S + i];
    }
    __syncwarp();
    float r, g, b, d, o;
    composite_ray_warp(sz, ssig, srgb, S, rp.white_bkgd, rp.sigma_softplus, sw, stmp, r, g, b, d, o);
    if (lane == 0) {
      if (comp_rgb) { comp_rgb[ray * 3] = r; comp_rgb[ray * 3 + 1] = g; comp_rgb[ray * 3 + 2] = b; }
      if (depth) depth[ray] = d;
      if (opacity) opacity[ray] = o;
    }
    if (weights) for (int i = lane; i < S; i += 32) weights[ray * S + i] = sw[i];
    if (do_resample) {
      float* zo = scratch + 2 * S + n_imp;
      resample_ray_warp(sz, sw, S, n_imp, u_resample ? u_resample + ray * n_imp : nullptr, tabs->u_fine,
                        scratch, zo);
      for (int i = lane; i < S + n_imp; i += 32) z_next[ray * (S + n_imp) + i] = zo[i];
    }
    __syncwarp();
  }
}

// One warp per ray: resampling only (nsr_resample seam).
__global__ void __launch_bounds__(128)
k_resample(RenderParams rp, const SampleTables* __restrict__ tabs, const float* __restrict__ z,
           const float* __restrict__ w, int64_t n_rays, const float* __restrict__ u,
           float* __restrict__ z_out, int per_warp_floats) {
  extern __shared__ float csm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = rp.n_coarse, n_imp = rp.n_importance;
  float* sz = csm + warp * per_warp_floats;
  float* sw = sz + S;
  float* scratch = sw + S;
  float* zo = scratch + 2 * S + n_imp;
  for (int64_t ray = blockIdx.x * 4 + warp; ray < n_rays; ray += (int64_t)gridDim.x * 4) {
    for (int i = lane; i < S; i += 32) { sz[i] = z[ray * S + i]; sw[i] = w[ray * S + i]; }
    __syncwarp();
    resample_ray_warp(sz, sw, S, n_imp, u ? u + ray * n_imp : nullptr, tabs->u_fine, scratch, zo);
    for (int i = lane; i < S + n_imp; i += 32) z_out[ray * (S + n_imp) + i] = zo[i];
    __syncwarp();
  }
}

__global__ void k_posenc(const float* __restrict__ x, int64_t n, int deg, int no_xyz,
                         const float* __restrict__ freqs, float* __restrict__ out) {
  const int ch = 6 * deg + (no_xyz ? 0 : 3);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float* o = out + i * ch;
    posenc3(x[3 * i], x[3 * i + 1], x[3 * i + 2], deg, freqs, no_xyz, [&](int c, float v) { o[c] = v; });
  }
}

__global__ void k_box_average(const float* __restrict__ in, int64_t n_lr, int ss, int channels,
                              float* __restrict__ out) {
  // torch.mean(reshape(x, (n_lr, s*s, -1)), dim=1)   (nerf_downX_model.py:337-348)
  const int64_t total = n_lr * channels;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / channels;
    const int c = (int)(idx % channels);
    float s = 0.f;
    for (int k = 0; k < ss; ++k) s = __fadd_rn(s, in[(p * ss + k) * channels + c]);
    out[idx] = __fdiv_rn(s, (float)ss);
  }
}

// box average + squared error against the LR target; one partial sum (double) per block
__global__ void __launch_bounds__(256)
k_lr_metrics_partial(const float* __restrict__ hr, const float* __restrict__ target, int64_t n_lr, int ss,
                     float* __restrict__ lr_out, double* __restrict__ partials) {
  const int64_t total = n_lr * 3;
  double acc = 0.0;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / 3;
    const int c = (int)(idx % 3);
    float sum = 0.f;
    for (int k = 0; k < ss; ++k) sum = __fadd_rn(sum, hr[(p * ss + k) * 3 + c]);
    const float lr = __fdiv_rn(sum, (float)ss);
    if (lr_out) lr_out[idx] = lr;
    const float d = __fsub_rn(lr, target[idx]);
    acc += (double)__fmul_rn(d, d);
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

__global__ void k_lr_metrics_final(const double* __restrict__ partials, int n_blocks, int64_t n_elems,
                                   float* __restrict__ metrics) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += partials[i];
    const float mse = (float)(s / (double)n_elems);
    metrics[0] = mse;
    metrics[1] = -10.f * log10f(mse);        // criterions.py:36
  }
}

__global__ void k_generate_rays(RayGenParams g, float* __restrict__ rays) {
  // get_ray_directions + get_rays (+ get_ndc_rays) + '(h s1) (w s2) c -> (h w) (s1 s2) c'
  // (models/utils.py:98-196; data/blender_downX_dataset.py:207-215): generate_ray(), nsr_device.cuh
  const int64_t total = (int64_t)g.H * g.W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float r[8];
    generate_ray(g, idx, r);
    float* out = rays + idx * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = r[k];
  }
}

// ----------------------------------------------------------------------------
// launch helpers
// ----------------------------------------------------------------------------
static int grid_for(int64_t work, int block, int cap) {
  int64_t g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

namespace nsr {
cudaError_t launch_composite(NsrHandle_* h, const float* raw, const float* z, const float* noise,
                             int64_t n_rays, int S, int do_resample, const float* u_resample,
                             float* comp_rgb, float* depth, float* opacity, float* weights,
                             float* z_next, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  const int n_imp = h->rp.n_importance;
  const int per_warp = 7 * S + (do_resample ? (3 * S + 2 * n_imp) : 0);
  const size_t smem = (size_t)per_warp * 4 * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_composite, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = grid_for(n_rays, 4, h->sm_count * 8);
  k_composite<<<grid, 128, smem, st>>>(h->rp, h->d_tables, raw, z, noise, n_rays, S, do_resample, u_resample,
                                       comp_rgb, depth, opacity, weights, z_next, per_warp);
  h->launches += 1;
  return cudaGetLastError();
}
}  // namespace nsr

// ----------------------------------------------------------------------------
// lifecycle
// ----------------------------------------------------------------------------
extern "C" int nsr_abi_version(void) { return NSR_ABI_VERSION; }

extern "C" const char* nsr_last_error(const NsrHandle* h) {
  return h ? h->err.c_str() : g_create_err.c_str();
}

static int popcount_below(uint32_t mask, int n) {
  int c = 0;
  for (int i = 0; i < n; ++i) c += (mask >> i) & 1u;
  return c;
}

extern "C" int nsr_create(const NsrConfig* cfg, NsrHandle** out_handle) {
  if (!cfg || !out_handle) return fail(nullptr, NSR_ERR_INVALID_ARG, "nsr_create: null argument");
  *out_handle = nullptr;
  if (cfg->struct_size != sizeof(NsrConfig))
    return fail(nullptr, NSR_ERR_INVALID_ARG, "nsr_create: NsrConfig.struct_size mismatch (ABI skew)");
  const NsrConfig& c = *cfg;
  if (c.D < 1 || c.D > kMaxTrunk) return fail(nullptr, NSR_ERR_UNSUPPORTED, "D must be in [1,16]");
  if (c.W != 64 && c.W != 128 && c.W != 256) return fail(nullptr, NSR_ERR_UNSUPPORTED, "W must be 64, 128 or 256");
  if (c.skips_mask & 1u) return fail(nullptr, NSR_ERR_UNSUPPORTED, "skip at layer 0 is ill-formed in the reference (networks.py:150-155)");
  if (c.skips_mask >> c.D) return fail(nullptr, NSR_ERR_INVALID_ARG, "skips_mask has bits >= D");
  if (c.deg_pos < 0 || c.deg_pos > 10 || c.deg_dir < 0 || c.deg_dir > 4)
    return fail(nullptr, NSR_ERR_UNSUPPORTED, "deg_pos must be <= 10 and deg_dir <= 4");
  if (c.no_xyz && (c.deg_pos == 0 || (c.deg_dir == 0 && !c.no_dir)))
    return fail(nullptr, NSR_ERR_INVALID_ARG, "no_xyz with zero frequencies gives an empty encoding");
  if (c.n_coarse < 3 || c.n_coarse > kMaxSamples) return fail(nullptr, NSR_ERR_UNSUPPORTED, "n_coarse must be in [3,256]");
  if (c.n_importance < 0 || c.n_coarse + c.n_importance > kMaxSamples)
    return fail(nullptr, NSR_ERR_UNSUPPORTED, "n_coarse + n_importance must be <= 256");
  if (c.precision < 0 || c.precision > 3) return fail(nullptr, NSR_ERR_INVALID_ARG, "unknown precision");
  if (c.viewdir_offset != 3 && c.viewdir_offset != 8) return fail(nullptr, NSR_ERR_INVALID_ARG, "viewdir_offset must be 3 or 8");
  if (c.sigma_activation < 0 || c.sigma_activation > 1 || c.color_activation < 0 || c.color_activation > 1)
    return fail(nullptr, NSR_ERR_INVALID_ARG, "unknown activation");
  if (c.precision != NSR_PREC_FP32_SIMT) {
    std::string why;
    if (!tc_supported(c, &why)) return fail(nullptr, NSR_ERR_UNSUPPORTED, "tensor-core path: " + why);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(nullptr, NSR_ERR_NO_DEVICE, "no CUDA device visible (libnsr_b200 has no CPU fallback)");
  }
  if (c.device < 0 || c.device >= ndev) return fail(nullptr, NSR_ERR_INVALID_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, NSR_ERR_CUDA, "cudaGetDeviceProperties failed"); }
  if (prop.major != 10) return fail(nullptr, NSR_ERR_NO_DEVICE, "device is not sm_100 (this library is built for sm_100a only)");

  NsrHandle_* h = new (std::nothrow) NsrHandle_();
  if (!h) return fail(nullptr, NSR_ERR_CUDA, "out of host memory");
  h->cfg = c;
  h->sm_count = prop.multiProcessorCount;
  if (const char* ev = getenv("NSR_TC_CLUSTER")) h->tc_cluster = atoi(ev) == 1 ? 1 : 2;
  if (const char* ev = getenv("NSR_TC_FUSED")) h->tc_fused = atoi(ev) != 0;
  if (h->sm_count % 2) h->tc_cluster = 1;
  RenderParams& rp = h->rp;
  rp.n_coarse = c.n_coarse; rp.n_importance = c.n_importance;
  rp.deg_pos = c.deg_pos; rp.deg_dir = c.deg_dir; rp.no_xyz = c.no_xyz;
  rp.ch_pos = 6 * c.deg_pos + (c.no_xyz ? 0 : 3);
  rp.ch_dir = 6 * c.deg_dir + (c.no_xyz ? 0 : 3);
  rp.lindisp = c.lindisp; rp.white_bkgd = c.white_bkgd; rp.sigma_softplus = c.sigma_activation;
  rp.color_none = c.color_activation; rp.gamma_correct = c.gamma_correct; rp.no_dir = c.no_dir;
  rp.viewdir_offset = c.viewdir_offset; rp.noise_std = c.noise_std;

  SampleTables& T = h->h_tables;
  memset(&T, 0, sizeof(T));
  host_linspace(0.f, 1.f, c.n_coarse, T.t_coarse);
  for (int i = 0; i < c.n_coarse; ++i) { volatile float v = 1.f - T.t_coarse[i]; T.one_minus_t[i] = v; }
  if (c.n_importance > 0) host_linspace(0.f, 1.f, c.n_importance, T.u_fine);
  auto bands = [&](int deg, float* out) {
    if (deg <= 0) return;
    if (c.no_logscale) host_linspace(1.f, (float)(1 << (deg - 1)), deg, out);   // embedding.py:42
    else for (int k = 0; k < deg; ++k) out[k] = (float)(1 << k);                 // embedding.py:40
  };
  bands(c.deg_pos, T.freq_pos);
  bands(c.deg_dir, T.freq_dir);

  // parameter table (state_dict order)
  {
    const int ch_pos = rp.ch_pos, ch_dir = rp.ch_dir;
    for (int i = 0; i < c.D; ++i) {
      const bool skip = (c.skips_mask >> i) & 1u;
      const int k = (i == 0) ? ch_pos : (skip ? c.W + ch_pos : c.W);
      h->param_numel.push_back((int64_t)c.W * k);
      h->param_numel.push_back(c.W);
    }
    h->param_numel.push_back((int64_t)c.W * c.W); h->param_numel.push_back(c.W);
    const int kd = c.no_dir ? c.W : c.W + ch_dir;
    h->param_numel.push_back((int64_t)(c.W / 2) * kd); h->param_numel.push_back(c.W / 2);
    h->param_numel.push_back(c.W); h->param_numel.push_back(1);
    h->param_numel.push_back(3 * (c.W / 2)); h->param_numel.push_back(3);
  }
  (void)popcount_below;

  cudaError_t e = cudaSetDevice(c.device);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_tables, sizeof(SampleTables));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_partials, 1024 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_tables, &T, sizeof(T), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_clk, 4 * sizeof(long long));
  if (e == cudaSuccess) e = cudaMemset(h->d_clk, 0, 4 * sizeof(long long));
  if (e == cudaSuccess && c.precision != NSR_PREC_FP32_SIMT) e = tc_init(h);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_jet, sizeof(kJetLut));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_jet, kJetLut, sizeof(kJetLut), cudaMemcpyHostToDevice);
  const size_t blob = simt_blob_floats(h);
  for (int w = 0; w < 2 && e == cudaSuccess; ++w) {
    e = cudaMalloc(&h->net[w].simt_blob, blob * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(h->net[w].simt_blob, 0, blob * sizeof(float));
    if (e == cudaSuccess && c.precision != NSR_PREC_FP32_SIMT) {
      h->net[w].tc_bytes = tc_image_bytes(h);
      e = cudaMalloc(&h->net[w].tc_image, h->net[w].tc_bytes);
      if (e == cudaSuccess) e = cudaMalloc(&h->net[w].tc_consts, 16384 * sizeof(float));
      if (e == cudaSuccess) e = cudaMalloc(&h->net[w].wt_image, train_wt_bytes());
    }
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("nsr_create: ") + cudaGetErrorString(e);
    nsr_destroy(h);
    cudaGetLastError();
    return fail(nullptr, NSR_ERR_CUDA, msg);
  }
  *out_handle = h;
  return NSR_OK;
}

extern "C" int nsr_destroy(NsrHandle* h) {
  if (!h) return NSR_OK;
  cudaSetDevice(h->cfg.device);
  for (int w = 0; w < 2; ++w) {
    cudaFree(h->net[w].simt_blob); cudaFree(h->net[w].tc_image); cudaFree(h->net[w].tc_consts); cudaFree(h->net[w].wt_image);
  }
  cudaFree(h->d_tables);
  cudaFree(h->d_partials);
  cudaFree(h->d_jet);
  cudaFree(h->d_clk);
  cudaFree(h->frame_rays);
  if (h->frame_ev) cudaEventDestroy(h->frame_ev);
  for (int w = 0; w < 2; ++w) if (h->pack_ev[w]) cudaEventDestroy(h->pack_ev[w]);
  for (int k = 0; k < 2; ++k) {
    if (h->dw_st[k]) cudaStreamDestroy(h->dw_st[k]);
    if (h->dw_done[k]) cudaEventDestroy(h->dw_done[k]);
  }
  if (h->dw_fork) cudaEventDestroy(h->dw_fork);
  for (int i = 0; i < 2; ++i) {
    if (h->pin_in[i]) cudaFreeHost(h->pin_in[i]);
    if (h->pin_out[i]) cudaFreeHost(h->pin_out[i]);
    cudaFree(h->dev_in[i]); cudaFree(h->dev_out[i]); cudaFree(h->dev_ws[i]);
    if (h->hev[i]) cudaEventDestroy(h->hev[i]);
    if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
  }
  delete h;
  return NSR_OK;
}

extern "C" int nsr_param_count(const NsrHandle* h) { return h ? (int)h->param_numel.size() : 0; }
extern "C" int64_t nsr_param_numel(const NsrHandle* h, int index) {
  if (!h || index < 0 || index >= (int)h->param_numel.size()) return -1;
  return h->param_numel[index];
}
extern "C" int nsr_debug_set_trace(NsrHandle* h, long long* device_buffer) {
  if (!h) return NSR_ERR_INVALID_ARG;
  h->trace_buf = device_buffer;
  return NSR_OK;
}
extern "C" int nsr_debug_set_flags(NsrHandle* h, int flags) {
  if (!h) return NSR_ERR_INVALID_ARG;
  h->debug_flags = flags;
  return NSR_OK;
}
extern "C" int64_t nsr_launch_count(const NsrHandle* h) { return h ? h->launches.load() : 0; }
extern "C" int nsr_debug_frame_schedule(int64_t pairs, int unit_rays, int first_unit, int unit_stride, int64_t* out) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long here");
  return frame_schedule((long long)pairs, unit_rays, first_unit, unit_stride, reinterpret_cast<long long*>(out)) ? NSR_ERR_INVALID_ARG : NSR_OK;
}

extern "C" int nsr_debug_kernel_clock(NsrHandle* h, int64_t* out4_host, NsrStream stream) {
  if (!h || !out4_host) return NSR_ERR_INVALID_ARG;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  NSR_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  long long v[4];
  NSR_CUDA(h, cudaMemcpy(v, h->d_clk, sizeof(v), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out4_host[i] = (int64_t)v[i];
  return NSR_OK;
}

extern "C" int nsr_pack_weights(NsrHandle* h, int which, const float* const* param_ptrs, int n_params,
                                NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (which < 0 || which > 1 || !param_ptrs) return fail(h, NSR_ERR_INVALID_ARG, "nsr_pack_weights: bad argument");
  if (n_params != (int)h->param_numel.size())
    return fail(h, NSR_ERR_INVALID_ARG, "nsr_pack_weights: expected " + std::to_string(h->param_numel.size()) +
                                            " state_dict tensors, got " + std::to_string(n_params));
  for (int i = 0; i < n_params; ++i)
    if (!param_ptrs[i]) return fail(h, NSR_ERR_INVALID_ARG, "nsr_pack_weights: null parameter pointer");
  cudaStream_t st = (cudaStream_t)stream;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  // each precision packs only the images its kernels read (training re-packs every step)
  if (h->cfg.precision == NSR_PREC_FP32_SIMT) NSR_CUDA(h, simt_pack(h, which, param_ptrs, st));
  if (h->cfg.precision != NSR_PREC_FP32_SIMT) {
    NSR_CUDA(h, tc_pack(h, which, param_ptrs, st));
    // the device-side pointer table tc_pack staged behind the consts blob also feeds the backward's images
    if (h->cfg.W == 256)        // (training is built for the 256-wide net)
      NSR_CUDA(h, train_pack_wt(h, which, reinterpret_cast<const float* const*>(h->net[which].tc_consts + 8192), st));
  }
  h->net[which].packed = true;
  // The host-buffer pipeline (nsr_render_host / nsr_render_pose_host) runs on library-owned non-blocking streams: order it
  // behind this pack (and behind whatever the caller enqueued on `stream` before it, e.g. the upload of the tensors)
  if (!h->pack_ev[which]) NSR_CUDA(h, cudaEventCreateWithFlags(&h->pack_ev[which], cudaEventDisableTiming));
  NSR_CUDA(h, cudaEventRecord(h->pack_ev[which], st));
  h->pack_pending[which] = true;
  return NSR_OK;
}

// ----------------------------------------------------------------------------
// hot path
// ----------------------------------------------------------------------------
struct WsLayout { size_t z_c, z_f, raw, total; };

static WsLayout ws_layout(const NsrHandle_* h, int64_t n) {
  const int Sc = h->cfg.n_coarse, Sf = Sc + h->cfg.n_importance;
  WsLayout L{};
  size_t off = 0;
  L.z_f = off; off += align_up((size_t)n * Sf * sizeof(float));
  if (h->cfg.precision == NSR_PREC_FP32_SIMT) {
    L.z_c = off; off += align_up((size_t)n * Sc * sizeof(float));
    L.raw = off; off += align_up((size_t)n * Sf * 4 * sizeof(float));
  } else if (Sf > 128) {      // tensor-core fine pass in MLP-only mode (192 / 256 samples): raw (rgb, sigma) for k_composite
    L.raw = off; off += align_up((size_t)n * Sf * 4 * sizeof(float));
  }
  L.total = off + 256;
  return L;
}

extern "C" size_t nsr_workspace_bytes(const NsrHandle* h, int64_t n_rays) {
  if (!h || n_rays < 0) return 0;
  return ws_layout(h, n_rays).total;
}

static int check_render_args(NsrHandle_* h, const float* rays, int64_t n_rays, int ray_stride) {
  if (!rays && n_rays > 0) return fail(h, NSR_ERR_INVALID_ARG, "rays is null");
  if (n_rays < 0) return fail(h, NSR_ERR_INVALID_ARG, "n_rays < 0");
  if (ray_stride < 8 || ray_stride < h->cfg.viewdir_offset + 3)
    return fail(h, NSR_ERR_INVALID_ARG, "ray_stride too small for (o,d,near,far[,viewdir])");
  return NSR_OK;
}

static int run_pass(NsrHandle_* h, int which, const float* rays, int64_t n, int stride, const float* z_in,
                    int S, const float* u_jitter, const float* noise, int do_resample, const float* u_res,
                    float* comp, float* depth, float* opa, float* wts, float* raw_out, float* z_next,
                    float* ws_z, float* ws_raw, cudaStream_t st) {
  if (!h->net[which].packed) return fail(h, NSR_ERR_NOT_PACKED, "weights of net " + std::to_string(which) + " not packed");
  if (h->cfg.precision == NSR_PREC_FP32_SIMT) {
    const float* z = z_in;
    if (!z) {   // coarse pass: sample z on device
      const int64_t total = n * S;
      k_sample_coarse<<<grid_for(total, 256, h->sm_count * 16), 256, 0, st>>>(h->rp, h->d_tables, rays, n, stride, u_jitter, ws_z);
      h->launches += 1;
      z = ws_z;
    }
    float* raw = raw_out ? raw_out : ws_raw;
    NSR_CUDA(h, simt_mlp(h, which, rays, n, stride, z, S, raw, st));
    NSR_CUDA(h, launch_composite(h, raw, z, noise, n, S, do_resample, u_res, comp, depth, opa, wts, z_next, st));
    return NSR_OK;
  }
  TcPassArgs a{};
  a.rays = rays; a.n_rays = n; a.ray_stride = stride; a.z_in = z_in; a.S = S;
  a.u_jitter = u_jitter; a.noise = noise; a.u_resample = u_res; a.do_resample = do_resample;
  a.trace = h->trace_buf;
  a.debug_flags = h->debug_flags;
  a.comp_rgb = comp; a.depth = depth; a.opacity = opa; a.weights = wts; a.raw = raw_out; a.z_next = z_next;
  if (S > 128) {              // MLP-only mode: tiles cut across rays, compositing afterwards with one warp per ray
    if (!z_in) return fail(h, NSR_ERR_UNSUPPORTED, "a pass with more than 128 samples per ray needs caller-supplied z-values");
    float* raw = raw_out ? raw_out : ws_raw;
    a.raw = raw; a.noise = nullptr; a.comp_rgb = nullptr; a.depth = nullptr; a.opacity = nullptr; a.weights = nullptr;
    NSR_CUDA(h, tc_pass(h, which, a, st));
    NSR_CUDA(h, launch_composite(h, raw, z_in, noise, n, S, 0, nullptr, comp, depth, opa, wts, nullptr, st));
    return NSR_OK;
  }
  NSR_CUDA(h, tc_pass(h, which, a, st));
  return NSR_OK;
}

// LR (box-averaged) outputs of a frame render; any pointer may be null.
struct LrOut { float* coarse_rgb = nullptr; float* coarse_depth = nullptr; float* fine_rgb = nullptr; float* fine_depth = nullptr; };

static bool fused_frame_ok(const NsrHandle_* h, int s) {
  return !(h->debug_flags & 64) && h->cfg.n_importance > 0 && tc_frame_supported(h, s) && h->net[0].packed && h->net[1].packed;
}

// Whether the box average belongs in the frame kernel's epilogue.  There a CTA's work quantum is a whole LR pixel's rays
// (s*s rays = 1.5 s*s tiles) instead of a ray pair (3 tiles), so the busiest CTA of a SMALL batch can end up with one
// quantum more than it would otherwise (2048 rays at s = 2: 24 tiles instead of 21, +14 %); from a few 10^4 rays on the
// difference vanishes (a 400 x 400 frame: 1626 vs 1623 tiles).  In-kernel when it costs < 3 %; else the frame kernel writes
// the HR composite and k_box_average follows (microseconds).
static bool lr_in_kernel_pays(const NsrHandle_* h, int64_t n, int s) {
  if (s <= 1 || !fused_frame_ok(h, s)) return false;
  auto tiles_of_busiest_cta = [&](int unit) {
    const int64_t units = (n + unit - 1) / unit;
    int64_t grid = units < h->sm_count ? units : h->sm_count;
    if (h->tc_cluster == 2 && units >= 2) grid = (grid + 1) & ~(int64_t)1;
    return (units + grid - 1) / grid * (unit / 2) * 3;
  };
  return tiles_of_busiest_cta(s * s) * 100 <= tiles_of_busiest_cta(2) * 103;
}

extern "C" int nsr_debug_frame_lr_in_kernel(const NsrHandle* h, int64_t n_rays, int s) {
  return (h && lr_in_kernel_pays(h, n_rays, s)) ? 1 : 0;
}

// forward_rays over a ray batch (+ the s x s box average when `lr` is given).  One launch (k_tc_pass<.., FUSED>) where the
// option set allows, else coarse pass, fine pass and one k_box_average per LR output.  `rg` != null: rays are generated from
// the pose inside the fused kernel (callers check fused_frame_ok first).
static int render_core(NsrHandle_* h, const float* rays, const RayGenParams* rg, int64_t n_rays, int ray_stride, int s,
                       const NsrRng* rng, const NsrOutputs* out, const LrOut* lr, char* ws, const WsLayout& L, cudaStream_t st,
                       int64_t rg_first = 0) {
  const int Sc = h->cfg.n_coarse, Ni = h->cfg.n_importance;
  float* z_f = out->z_fine ? out->z_fine : (float*)(ws + L.z_f);
  float* ws_zc = (float*)(ws + L.z_c);
  float* ws_raw = (float*)(ws + L.raw);
  const NsrRng none{};
  const NsrRng& R = rng ? *rng : none;
  const float* noise_c = (h->cfg.noise_std > 0.f) ? R.noise_coarse : nullptr;
  const float* noise_f = (h->cfg.noise_std > 0.f) ? R.noise_fine : nullptr;
  const bool want_lr = lr && s > 1;
  // box average in the kernel's epilogue: when it pays, or when the caller did not ask for the HR composite it would
  // otherwise be computed from
  bool lr_k = false;
  if (want_lr && fused_frame_ok(h, s)) {
    const bool hr_there = (!lr->coarse_rgb || out->coarse_comp_rgbs) && (!lr->coarse_depth || out->coarse_depth) &&
                          (!lr->fine_rgb || out->fine_comp_rgbs) && (!lr->fine_depth || out->fine_depth);
    lr_k = !hr_there || lr_in_kernel_pays(h, n_rays, s);
  }
  if (fused_frame_ok(h, lr_k ? s : 1)) {
    TcFrameArgs f{};
    f.rays = rays; f.rg = rg; f.rg_first = rg_first; f.n_rays = n_rays; f.ray_stride = ray_stride; f.s = lr_k ? s : 1;
    f.u_jitter = R.u_coarse; f.noise_c = noise_c; f.noise_f = noise_f; f.u_resample = R.u_fine;
    f.z_fine = z_f;
    f.c_rgb = out->coarse_comp_rgbs; f.c_depth = out->coarse_depth; f.c_opacity = out->coarse_opacity; f.c_weights = out->coarse_weights;
    f.f_rgb = out->fine_comp_rgbs; f.f_depth = out->fine_depth; f.f_opacity = out->fine_opacity; f.f_weights = out->fine_weights;
    if (lr_k) { f.lr_rgb_c = lr->coarse_rgb; f.lr_depth_c = lr->coarse_depth; f.lr_rgb_f = lr->fine_rgb; f.lr_depth_f = lr->fine_depth; }
    f.trace = h->trace_buf; f.debug_flags = h->debug_flags;
    NSR_CUDA(h, tc_frame(h, f, st));
    if (!want_lr || lr_k) return NSR_OK;
  } else {
    if (!rays) return fail(h, NSR_ERR_UNSUPPORTED, "in-kernel ray generation needs the fused frame kernel");
    int rc = run_pass(h, 0, rays, n_rays, ray_stride, nullptr, Sc, R.u_coarse, noise_c, Ni > 0, R.u_fine,
                      out->coarse_comp_rgbs, out->coarse_depth, out->coarse_opacity, out->coarse_weights, nullptr,
                      Ni > 0 ? z_f : nullptr, ws_zc, ws_raw, st);
    if (rc) return rc;
    if (Ni > 0) {
      rc = run_pass(h, 1, rays, n_rays, ray_stride, z_f, Sc + Ni, nullptr, noise_f, 0, nullptr, out->fine_comp_rgbs,
                    out->fine_depth, out->fine_opacity, out->fine_weights, nullptr, nullptr, ws_zc, ws_raw, st);
      if (rc) return rc;
    }
  }
  if (want_lr) {
    int rc = NSR_OK;
    const int64_t n_lr = n_rays / ((int64_t)s * s);
    auto box = [&](const float* in, int ch, float* o) -> int {
      if (!o) return NSR_OK;
      if (!in) return fail(h, NSR_ERR_INVALID_ARG, "an LR output needs the corresponding HR output buffer on this path");
      return nsr_box_average(h, in, n_lr, s, ch, o, st);
    };
    if ((rc = box(out->coarse_comp_rgbs, 3, lr->coarse_rgb))) return rc;
    if ((rc = box(out->coarse_depth, 1, lr->coarse_depth))) return rc;
    if (Ni > 0) {
      if ((rc = box(out->fine_comp_rgbs, 3, lr->fine_rgb))) return rc;
      if ((rc = box(out->fine_depth, 1, lr->fine_depth))) return rc;
    }
  }
  return NSR_OK;
}

extern "C" int nsr_render(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const NsrRng* rng,
                          const NsrOutputs* out, void* workspace, size_t workspace_bytes, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!out) return fail(h, NSR_ERR_INVALID_ARG, "nsr_render: out is null");
  int rc = check_render_args(h, rays, n_rays, ray_stride);
  if (rc) return rc;
  if (n_rays == 0) return NSR_OK;
  const WsLayout L = ws_layout(h, n_rays);
  if (!workspace || workspace_bytes < L.total) return fail(h, NSR_ERR_WORKSPACE, "workspace too small: need " + std::to_string(L.total));
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  return render_core(h, rays, nullptr, n_rays, ray_stride, 1, rng, out, nullptr, ws, L, (cudaStream_t)stream);
}

static int fill_raygen(NsrHandle_* h, const float* c2w_host, const NsrRayGen* spec, RayGenParams* p) {
  if (!c2w_host || !spec || spec->struct_size != sizeof(NsrRayGen))
    return fail(h, NSR_ERR_INVALID_ARG, "ray generation: null pose or NsrRayGen.struct_size mismatch");
  if (spec->H <= 0 || spec->W <= 0 || spec->s < 1 || !(spec->focal > 0.f) || spec->H % spec->s || spec->W % spec->s)
    return fail(h, NSR_ERR_INVALID_ARG, "ray generation: bad raster (H, W multiples of s; focal > 0)");
  if (spec->unified_dir && !(floorf(spec->focal / (float)spec->s) > 0.f))
    return fail(h, NSR_ERR_INVALID_ARG, "ray generation: unified_dir needs focal // s > 0");
  memcpy(p->m, c2w_host, sizeof(p->m));
  p->H = spec->H; p->W = spec->W; p->focal = spec->focal; p->s = spec->s; p->ndc = spec->ndc;
  p->near_plane = spec->near_plane; p->far_plane = spec->far_plane;
  p->pixel_center = spec->use_pixel_centers ? 0.5f : 0.f; p->unified_dir = spec->unified_dir;
  return NSR_OK;
}

extern "C" size_t nsr_frame_workspace_bytes(const NsrHandle* h, int64_t n_rays, int from_pose) {
  if (!h || n_rays < 0) return 0;
  // + room for what an option set on the multi-launch path needs: the generated rays (from a pose) and HR composites /
  // depths the caller did not ask for but the box average is computed from (8 floats per ray: rgb + depth of both nets)
  return ws_layout(h, n_rays).total + (from_pose ? align_up((size_t)n_rays * 8 * sizeof(float)) : 0) +
         align_up((size_t)n_rays * 8 * sizeof(float));
}

extern "C" int nsr_render_frame(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const float* c2w_host,
                                const NsrRayGen* spec, int s, const NsrRng* rng, const NsrOutputs* out,
                                const NsrLrOutputs* lr, void* workspace, size_t workspace_bytes, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!out || !lr) return fail(h, NSR_ERR_INVALID_ARG, "nsr_render_frame: out / lr is null");
  if (s < 1) return fail(h, NSR_ERR_INVALID_ARG, "nsr_render_frame: s < 1");
  RayGenParams rg{};
  const bool from_pose = rays == nullptr;
  int rc;
  if (from_pose) {
    if ((rc = fill_raygen(h, c2w_host, spec, &rg))) return rc;
    if (spec->s != s) return fail(h, NSR_ERR_INVALID_ARG, "nsr_render_frame: spec->s differs from s");
    if (h->cfg.viewdir_offset != 3) return fail(h, NSR_ERR_UNSUPPORTED, "pose rendering produces 8-column rays (NeRFDownXModel layout)");
    n_rays = (int64_t)spec->H * spec->W; ray_stride = 8;
  } else if ((rc = check_render_args(h, rays, n_rays, ray_stride))) return rc;
  if (n_rays % ((int64_t)s * s)) return fail(h, NSR_ERR_INVALID_ARG, "n_rays must be a multiple of s*s");
  if (n_rays == 0) return NSR_OK;
  const WsLayout L = ws_layout(h, n_rays);
  const size_t need = nsr_frame_workspace_bytes(h, n_rays, from_pose);
  if (!workspace || workspace_bytes < need) return fail(h, NSR_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need));
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  cudaStream_t st = (cudaStream_t)stream;
  LrOut l{lr->coarse_rgb, lr->coarse_depth, lr->fine_rgb, lr->fine_depth};
  const bool want_lr = s > 1 && (l.coarse_rgb || l.coarse_depth || l.fine_rgb || l.fine_depth);
  char* extra = ws + (L.total - 256);               // (L.total = 256-aligned pieces + 256 of alignment slack)
  if (from_pose && !fused_frame_ok(h, 1)) {      // other option sets: rays through HBM, then the ordinary passes
    float* gen = (float*)extra;
    k_generate_rays<<<grid_for(n_rays, 256, h->sm_count * 16), 256, 0, st>>>(rg, gen);
    h->launches += 1;
    NSR_CUDA(h, cudaGetLastError());
    rays = gen;
  }
  if (from_pose) extra += align_up((size_t)n_rays * 8 * sizeof(float));
  NsrOutputs o = *out;
  if (want_lr && !fused_frame_ok(h, s)) {
    // the box average runs as its own launches here and reads the HR composites: lend workspace to those the caller left out
    float* tmp = (float*)extra;
    if (l.coarse_rgb && !o.coarse_comp_rgbs) o.coarse_comp_rgbs = tmp;
    if (l.coarse_depth && !o.coarse_depth) o.coarse_depth = tmp + 3 * n_rays;
    if (l.fine_rgb && !o.fine_comp_rgbs) o.fine_comp_rgbs = tmp + 4 * n_rays;
    if (l.fine_depth && !o.fine_depth) o.fine_depth = tmp + 7 * n_rays;
  }
  return render_core(h, rays, rays ? nullptr : &rg, n_rays, ray_stride, s, rng, &o, want_lr ? &l : nullptr, ws, L, st);
}

extern "C" int nsr_render_pass(NsrHandle* h, int which, const float* rays, int64_t n_rays, int ray_stride,
                               const float* z_vals, int n_samples, const float* noise,
                               const NsrPassOutputs* out, void* workspace, size_t workspace_bytes,
                               NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!out || !z_vals || which < 0 || which > 1) return fail(h, NSR_ERR_INVALID_ARG, "nsr_render_pass: bad argument");
  int rc = check_render_args(h, rays, n_rays, ray_stride);
  if (rc) return rc;
  if (n_samples != h->cfg.n_coarse && n_samples != h->cfg.n_coarse + h->cfg.n_importance)
    return fail(h, NSR_ERR_UNSUPPORTED, "n_samples must be n_coarse or n_coarse+n_importance");
  if (n_rays == 0) return NSR_OK;
  const WsLayout L = ws_layout(h, n_rays);
  if (!workspace || workspace_bytes < L.total) return fail(h, NSR_ERR_WORKSPACE, "workspace too small: need " + std::to_string(L.total));
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  return run_pass(h, which, rays, n_rays, ray_stride, z_vals, n_samples, nullptr,
                  (h->cfg.noise_std > 0.f) ? noise : nullptr, 0, nullptr, out->comp_rgbs, out->depth,
                  out->opacity, out->weights, out->raw, nullptr, (float*)(ws + L.z_c), (float*)(ws + L.raw),
                  (cudaStream_t)stream);
}

extern "C" int nsr_sample_coarse(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride,
                                 const float* u, float* z_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  int rc = check_render_args(h, rays, n_rays, ray_stride);
  if (rc) return rc;
  if (!z_out) return fail(h, NSR_ERR_INVALID_ARG, "z_out is null");
  if (n_rays == 0) return NSR_OK;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  const int64_t total = n_rays * h->cfg.n_coarse;
  k_sample_coarse<<<grid_for(total, 256, h->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(
      h->rp, h->d_tables, rays, n_rays, ray_stride, u, z_out);
  h->launches += 1;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_resample(NsrHandle* h, const float* z_in, const float* weights, int64_t n_rays,
                            const float* u, float* z_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!z_in || !weights || !z_out || n_rays < 0) return fail(h, NSR_ERR_INVALID_ARG, "nsr_resample: bad argument");
  if (h->cfg.n_importance <= 0) return fail(h, NSR_ERR_UNSUPPORTED, "n_importance == 0");
  if (n_rays == 0) return NSR_OK;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  const int S = h->cfg.n_coarse, n_imp = h->cfg.n_importance;
  const int per_warp = 2 * S + (3 * S + 2 * n_imp);
  const size_t smem = (size_t)per_warp * 4 * sizeof(float);
  k_resample<<<grid_for(n_rays, 4, h->sm_count * 8), 128, smem, (cudaStream_t)stream>>>(
      h->rp, h->d_tables, z_in, weights, n_rays, u, z_out, per_warp);
  h->launches += 1;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_posenc(NsrHandle* h, const float* x, int64_t n, int deg, float* out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!x || !out || n < 0) return fail(h, NSR_ERR_INVALID_ARG, "nsr_posenc: bad argument");
  const float* freqs = nullptr;
  if (deg == h->cfg.deg_pos) freqs = h->d_tables->freq_pos;
  else if (deg == h->cfg.deg_dir) freqs = h->d_tables->freq_dir;
  else return fail(h, NSR_ERR_UNSUPPORTED, "nsr_posenc: deg must be the handle's deg_pos or deg_dir");
  if (n == 0) return NSR_OK;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  k_posenc<<<grid_for(n, 256, h->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(x, n, deg, h->cfg.no_xyz, freqs, out);
  h->launches += 1;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_box_average(NsrHandle* h, const float* in, int64_t n_lr, int s, int channels, float* out,
                               NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!in || !out || n_lr < 0 || s < 1 || channels < 1) return fail(h, NSR_ERR_INVALID_ARG, "nsr_box_average: bad argument");
  if (n_lr == 0) return NSR_OK;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  k_box_average<<<grid_for(n_lr * channels, 256, h->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(
      in, n_lr, s * s, channels, out);
  h->launches += 1;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

static int generate_rays_impl(NsrHandle* h, const float* c2w_host, int H, int W, float focal, int s, int ndc,
                              float near_plane, float far_plane, int use_pixel_centers, int unified_dir, float* rays_out,
                              NsrStream stream) {
  if (!c2w_host || !rays_out || H <= 0 || W <= 0 || s < 1 || !(focal > 0.f))
    return fail(h, NSR_ERR_INVALID_ARG, "nsr_generate_rays: bad argument");
  if (H % s || W % s) return fail(h, NSR_ERR_INVALID_ARG, "H and W must be multiples of the supersampling factor");
  if (unified_dir && !(floorf(focal / (float)s) > 0.f))
    return fail(h, NSR_ERR_INVALID_ARG, "nsr_generate_rays: unified_dir needs focal // s > 0");
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  RayGenParams p{};
  memcpy(p.m, c2w_host, sizeof(p.m));
  p.H = H; p.W = W; p.focal = focal; p.s = s; p.ndc = ndc; p.near_plane = near_plane; p.far_plane = far_plane;
  p.pixel_center = use_pixel_centers ? 0.5f : 0.f; p.unified_dir = unified_dir;
  k_generate_rays<<<grid_for((int64_t)H * W, 256, h->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(p, rays_out);
  h->launches += 1;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_generate_rays(NsrHandle* h, const float* c2w_host, int H, int W, float focal, int s, int ndc,
                                 float near_plane, float far_plane, float* rays_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  return generate_rays_impl(h, c2w_host, H, W, focal, s, ndc, near_plane, far_plane, 1, 0, rays_out, stream);
}

extern "C" int nsr_generate_rays_ex(NsrHandle* h, const float* c2w_host, const NsrRayGen* spec, float* rays_out,
                                    NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!spec || spec->struct_size != sizeof(NsrRayGen))
    return fail(h, NSR_ERR_INVALID_ARG, "nsr_generate_rays_ex: NsrRayGen.struct_size mismatch");
  return generate_rays_impl(h, c2w_host, spec->H, spec->W, spec->focal, spec->s, spec->ndc, spec->near_plane, spec->far_plane,
                            spec->use_pixel_centers, spec->unified_dir, rays_out, stream);
}

extern "C" int nsr_lr_metrics(NsrHandle* h, const float* hr_rgb, const float* target_lr, int64_t n_lr, int s,
                              float* lr_rgb_out, float* metrics_out, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!hr_rgb || !target_lr || !metrics_out || n_lr <= 0 || s < 1)
    return fail(h, NSR_ERR_INVALID_ARG, "nsr_lr_metrics: bad argument");
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  const int blocks = grid_for(n_lr * 3, 256, 1024);
  k_lr_metrics_partial<<<blocks, 256, 0, (cudaStream_t)stream>>>(hr_rgb, target_lr, n_lr, s * s, lr_rgb_out, h->d_partials);
  k_lr_metrics_final<<<1, 32, 0, (cudaStream_t)stream>>>(h->d_partials, blocks, n_lr * 3, metrics_out);
  h->launches += 2;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

// ----------------------------------------------------------------------------
// frame assembly (scope row f-3)
// ----------------------------------------------------------------------------
// numpy's float32 -> uint8 astype on x86-64: truncate through int32 (out of range / NaN -> INT32_MIN), low byte
__device__ __forceinline__ uint32_t cast_u8(float v) {
  if (!(fabsf(v) < 2147483648.f)) return 0u;
  return (uint32_t)__float2int_rz(v) & 0xFFu;
}

__global__ void __launch_bounds__(256)
k_assemble_frame(const float* __restrict__ rgb, const float* __restrict__ depth, const float* __restrict__ gt, int H, int W,
                 int s, float near, float far, const uint32_t* __restrict__ jet, uint8_t* __restrict__ out,
                 float* __restrict__ depth_mat) {
  const int panels = gt ? 3 : 2;
  const int w1 = W / s;
  const float denom = fmaxf(__fsub_rn(far, near), 1e-8f);
  const int64_t total = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i % W);
    // inverse of the dataset's '(h s1) (w s2) c -> (h w) (s1 s2) c' grouping (nerf_downX_model.py:410-416)
    const int64_t src = ((int64_t)(y / s) * w1 + (x / s)) * (s * s) + (y % s) * s + (x % s);
    uint8_t* o = out + ((int64_t)y * W * panels + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)cast_u8(__fmul_rn(rgb[src * 3 + c], 255.f));      // visualizer.py:54
    if (gt) {
      uint8_t* og = o + (int64_t)W * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) og[c] = (uint8_t)cast_u8(__fmul_rn(gt[src * 3 + c], 255.f));
    }
    float d = depth[src];
    if (d != d) d = 0.f;                                     // np.nan_to_num
    else if (isinf(d)) d = d > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    if (depth_mat) depth_mat[i] = d;
    const float xn = __fdiv_rn(__fsub_rn(d, near), denom);    // visualizer.py:170
    const uint32_t e = jet[cast_u8(__fmul_rn(255.f, xn))];     // :171-172, then (lut / 255.) * 255. -> uint8 is the identity
    uint8_t* od = o + (int64_t)W * 3 * (panels - 1);
    od[0] = (uint8_t)(e & 0xFFu); od[1] = (uint8_t)((e >> 8) & 0xFFu); od[2] = (uint8_t)((e >> 16) & 0xFFu);
  }
}

extern "C" int nsr_assemble_frame(NsrHandle* h, const float* rgb, const float* depth, const float* gt, int H, int W, int s,
                                  float near_plane, float far_plane, uint8_t* out_rgb8, float* depth_mat, NsrStream stream) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!rgb || !depth || !out_rgb8 || H <= 0 || W <= 0 || s < 1) return fail(h, NSR_ERR_INVALID_ARG, "nsr_assemble_frame: bad argument");
  if (H % s || W % s) return fail(h, NSR_ERR_INVALID_ARG, "H and W must be multiples of the supersampling factor");
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  k_assemble_frame<<<grid_for((int64_t)H * W, 256, h->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(
      rgb, depth, gt, H, W, s, near_plane, far_plane, h->d_jet, out_rgb8, depth_mat);
  h->launches += 1;
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

// ----------------------------------------------------------------------------
// host-buffer pipeline
// ----------------------------------------------------------------------------
static int ensure_host_state(NsrHandle_* h, int64_t chunk, int stride) {
  const size_t ws_bytes = ws_layout(h, chunk).total;
  if (h->host_chunk >= (size_t)chunk * stride && h->host_ws_bytes >= ws_bytes && h->hs[0]) return NSR_OK;
  for (int i = 0; i < 2; ++i) {
    if (h->pin_in[i]) cudaFreeHost(h->pin_in[i]);
    if (h->pin_out[i]) cudaFreeHost(h->pin_out[i]);
    cudaFree(h->dev_in[i]); cudaFree(h->dev_out[i]); cudaFree(h->dev_ws[i]);
    h->pin_in[i] = h->pin_out[i] = h->dev_in[i] = h->dev_out[i] = nullptr; h->dev_ws[i] = nullptr;
    if (!h->hs[i]) NSR_CUDA(h, cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
    if (!h->hev[i]) NSR_CUDA(h, cudaEventCreateWithFlags(&h->hev[i], cudaEventDisableTiming));
    NSR_CUDA(h, cudaMallocHost(&h->pin_in[i], (size_t)chunk * stride * sizeof(float)));
    NSR_CUDA(h, cudaMallocHost(&h->pin_out[i], (size_t)chunk * 4 * sizeof(float)));
    NSR_CUDA(h, cudaMalloc(&h->dev_in[i], (size_t)chunk * stride * sizeof(float)));
    NSR_CUDA(h, cudaMalloc(&h->dev_out[i], (size_t)chunk * 8 * sizeof(float)));   // rgb[3]+depth[1] HR, then LR copies
    NSR_CUDA(h, cudaMalloc(&h->dev_ws[i], ws_bytes));
  }
  h->host_chunk = (size_t)chunk * stride;
  h->host_ws_bytes = ws_bytes;
  return NSR_OK;
}

// nsr_pack_weights is asynchronous on the CALLER's stream; the frame pipeline below runs on hs[0..1].  Make both wait for the
// most recent pack of each net (stream-ordered, no host sync).
static int wait_for_packs(NsrHandle_* h) {
  for (int w = 0; w < 2; ++w) {
    if (!h->pack_pending[w]) continue;
    for (int i = 0; i < 2; ++i) NSR_CUDA(h, cudaStreamWaitEvent(h->hs[i], h->pack_ev[w], 0));
    h->pack_pending[w] = false;
  }
  return NSR_OK;
}

// Rays per chunk of the host-buffer pipeline.  The render needs 74 MB/s of rays at 2.3 M rays/s -- three orders of
// magnitude under PCIe -- so copy/compute overlap buys nothing, while every chunk boundary costs a partly filled last wave
// and two launches: chunks are as large as the staging buffers reasonably allow (a 160 000-ray frame is ONE chunk; its 5 MB
// upload takes 0.2 ms of 70), and only frames beyond that are pipelined over the two slots.
static constexpr int64_t kHostChunkRays = 262144;

// Chunked, double-buffered frame render.  Rays come either from host memory (staged through pinned
// buffers, H2D inside the pipeline) or from a device buffer the caller filled on stream hs[0]
// (rays_dev != null: nsr_render_pose_host generates them on the device).
static int render_frame_pipeline(NsrHandle* h, const float* rays_host, const float* rays_dev, const RayGenParams* rg,
                                 int64_t n_rays, int ray_stride, int s, float* rgb_host, float* depth_host) {
  int rc = NSR_OK;
  const int ss = s * s;
  int64_t chunk = kHostChunkRays;
  chunk -= chunk % ss;
  if (chunk > n_rays) chunk = n_rays;
  rc = ensure_host_state(h, chunk, ray_stride);
  if (rc) return rc;
  rc = wait_for_packs(h);
  if (rc) return rc;
  const bool fine = h->cfg.n_importance > 0;
  const int64_t n_chunks = (n_rays + chunk - 1) / chunk;
  // software pipeline over two slots: stage(k) | H2D+render+D2H(k) async | drain(k-1)
  auto drain = [&](int64_t k) {
    const int slot = (int)(k & 1);
    cudaEventSynchronize(h->hev[slot]);
    const int64_t r0 = k * chunk, nr = (r0 + chunk <= n_rays) ? chunk : n_rays - r0;
    const int64_t o0 = r0 / ss, no = nr / ss;
    if (rgb_host) memcpy(rgb_host + o0 * 3, h->pin_out[slot], (size_t)no * 3 * sizeof(float));
    if (depth_host) memcpy(depth_host + o0, h->pin_out[slot] + (size_t)chunk * 3, (size_t)no * sizeof(float));
  };
  for (int64_t k = 0; k < n_chunks; ++k) {
    const int slot = (int)(k & 1);
    if (k >= 2) drain(k - 2);
    const int64_t r0 = k * chunk, nr = (r0 + chunk <= n_rays) ? chunk : n_rays - r0;
    cudaStream_t st = h->hs[slot];
    const float* d_rays = h->dev_in[slot];
    if (rg) {
      d_rays = nullptr;                 // generated in the fused kernel's front-end from (pose, r0 + row)
    } else if (rays_dev) {
      d_rays = rays_dev + r0 * ray_stride;
    } else {
      memcpy(h->pin_in[slot], rays_host + r0 * ray_stride, (size_t)nr * ray_stride * sizeof(float));
      NSR_CUDA(h, cudaMemcpyAsync(h->dev_in[slot], h->pin_in[slot], (size_t)nr * ray_stride * sizeof(float),
                                  cudaMemcpyHostToDevice, st));
    }
    float* d_rgb = h->dev_out[slot];
    float* d_dep = d_rgb + (size_t)chunk * 3;
    float* d_rgb_lr = d_dep + (size_t)chunk;
    float* d_dep_lr = d_rgb_lr + (size_t)chunk * 3;
    // One launch per chunk on the fused path: the LR image comes straight out of the compositing epilogue and the HR
    // composite is not written at all; other option sets: coarse pass, fine pass, two box averages.
    const bool lr_k = lr_in_kernel_pays(h, nr, s);
    NsrOutputs o{};
    LrOut l{};
    if (!lr_k) {
      if (fine) { o.fine_comp_rgbs = d_rgb; o.fine_depth = d_dep; }
      else { o.coarse_comp_rgbs = d_rgb; o.coarse_depth = d_dep; }
    }
    if (s > 1) {
      if (fine) { l.fine_rgb = d_rgb_lr; l.fine_depth = d_dep_lr; }
      else { l.coarse_rgb = d_rgb_lr; l.coarse_depth = d_dep_lr; }
    }
    const WsLayout L = ws_layout(h, nr);
    char* ws = (char*)(((uintptr_t)h->dev_ws[slot] + 255) & ~(uintptr_t)255);
    rc = render_core(h, d_rays, rg, nr, ray_stride, s, nullptr, &o, s > 1 ? &l : nullptr, ws, L, st, r0);
    if (rc) return rc;
    const float* src_rgb = s > 1 ? d_rgb_lr : d_rgb;
    const float* src_dep = s > 1 ? d_dep_lr : d_dep;
    const int64_t no = nr / ss;
    if (rgb_host) NSR_CUDA(h, cudaMemcpyAsync(h->pin_out[slot], src_rgb, (size_t)no * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (depth_host) NSR_CUDA(h, cudaMemcpyAsync(h->pin_out[slot] + (size_t)chunk * 3, src_dep, (size_t)no * sizeof(float), cudaMemcpyDeviceToHost, st));
    NSR_CUDA(h, cudaEventRecord(h->hev[slot], st));
  }
  for (int64_t k = (n_chunks >= 2 ? n_chunks - 2 : 0); k < n_chunks; ++k) drain(k);
  NSR_CUDA(h, cudaGetLastError());
  return NSR_OK;
}

extern "C" int nsr_render_host(NsrHandle* h, const float* rays_host, int64_t n_rays, int ray_stride, int s,
                               float* rgb_host, float* depth_host) {
  if (!h) return NSR_ERR_INVALID_ARG;
  int rc = check_render_args(h, rays_host, n_rays, ray_stride);
  if (rc) return rc;
  if (s < 1 || n_rays % ((int64_t)s * s)) return fail(h, NSR_ERR_INVALID_ARG, "n_rays must be a multiple of s*s");
  if (n_rays == 0) return NSR_OK;
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  return render_frame_pipeline(h, rays_host, nullptr, nullptr, n_rays, ray_stride, s, rgb_host, depth_host);
}

extern "C" int nsr_render_pose_host(NsrHandle* h, const float* c2w_host, int H, int W, float focal, int s, int ndc,
                                    float near_plane, float far_plane, float* rgb_host, float* depth_host) {
  if (!h) return NSR_ERR_INVALID_ARG;
  if (!c2w_host || H <= 0 || W <= 0 || s < 1 || !(focal > 0.f)) return fail(h, NSR_ERR_INVALID_ARG, "nsr_render_pose_host: bad argument");
  if (H % s || W % s) return fail(h, NSR_ERR_INVALID_ARG, "H and W must be multiples of the supersampling factor");
  if (h->cfg.viewdir_offset != 3) return fail(h, NSR_ERR_UNSUPPORTED, "pose rendering produces 8-column rays (NeRFDownXModel layout)");
  NSR_CUDA(h, cudaSetDevice(h->cfg.device));
  const int64_t n_rays = (int64_t)H * W;
  int64_t chunk = kHostChunkRays;
  chunk -= chunk % (s * s);
  if (chunk > n_rays) chunk = n_rays;
  int rc = ensure_host_state(h, chunk, 8);      // creates the streams / events / staging
  if (rc) return rc;
  if (fused_frame_ok(h, 1)) {       // rays never exist in HBM: the fused kernel's front-end generates them from the pose
    NsrRayGen spec{};
    spec.struct_size = sizeof(NsrRayGen); spec.H = H; spec.W = W; spec.s = s; spec.focal = focal; spec.ndc = ndc;
    spec.near_plane = near_plane; spec.far_plane = far_plane; spec.use_pixel_centers = 1;
    RayGenParams rg{};
    if ((rc = fill_raygen(h, c2w_host, &spec, &rg))) return rc;
    if ((rc = wait_for_packs(h))) return rc;
    return render_frame_pipeline(h, nullptr, nullptr, &rg, n_rays, 8, s, rgb_host, depth_host);
  }
  if (h->frame_rays_cap < (size_t)n_rays * 8) {
    cudaFree(h->frame_rays);
    h->frame_rays = nullptr; h->frame_rays_cap = 0;
    NSR_CUDA(h, cudaMalloc(&h->frame_rays, (size_t)n_rays * 8 * sizeof(float)));
    h->frame_rays_cap = (size_t)n_rays * 8;
  }
  rc = wait_for_packs(h);
  if (rc) return rc;
  rc = nsr_generate_rays(h, c2w_host, H, W, focal, s, ndc, near_plane, far_plane, h->frame_rays, h->hs[0]);
  if (rc) return rc;
  if (!h->frame_ev) NSR_CUDA(h, cudaEventCreateWithFlags(&h->frame_ev, cudaEventDisableTiming));
  NSR_CUDA(h, cudaEventRecord(h->frame_ev, h->hs[0]));
  NSR_CUDA(h, cudaStreamWaitEvent(h->hs[1], h->frame_ev, 0));
  return render_frame_pipeline(h, nullptr, h->frame_rays, nullptr, n_rays, 8, s, rgb_host, depth_host);
}
