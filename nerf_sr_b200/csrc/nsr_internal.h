// nsr_internal.h -- host-side handle and cross-file declarations (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nsr.h"
#include "nsr_device.cuh"

namespace nsr {

constexpr int kMaxTrunk = 16;

// One Linear layer of VanillaMLP as the SIMT path executes it.
struct SimtLayer {
  int K0, K1;          // input segments: K0 rows from src0 then K1 rows from src1 (cat order)
  int src0, src1;      // buffer ids: 0 enc_xyz, 1 enc_dir, 2 bufA, 3 bufB
  int dst;             // output buffer id (2|3), or -1 for a head
  int N;               // outputs
  int relu;            // 1: ReLU
  int64_t w_off;       // offset (floats) of the permuted W^T image in the SIMT weight blob
  int64_t b_off;       // offset (floats) of the bias
};

struct SimtProgram {
  int n_layers;        // trunk + final + dir (tiled GEMM layers)
  SimtLayer layers[kMaxTrunk + 2];
  int sigma_src;       // buffer holding h_D (input of the sigma head)
  int rgb_src;         // buffer holding the dir layer's output
  int64_t w_sigma, b_sigma, w_rgb, b_rgb;   // offsets into the blob (plain [N][K] row-major)
  int W, ch_pos, ch_dir;
};

struct NetImages {
  bool packed = false;
  float* simt_blob = nullptr;      // device: permuted fp32 weights for the SIMT path
  size_t simt_floats = 0;
  uint8_t* tc_image = nullptr;     // device: swizzled hi/lo stage images for the tcgen05 path
  size_t tc_bytes = 0;
  float* tc_consts = nullptr;      // device: biases, head weights, dir-part weights (fp32)
  uint8_t* wt_image = nullptr;     // device: transposed hi/lo weight images for the backward dX GEMMs
};

}  // namespace nsr

struct NsrHandle_ {
  NsrConfig cfg;
  nsr::RenderParams rp;
  nsr::SimtProgram prog;
  nsr::SampleTables* d_tables = nullptr;
  nsr::SampleTables h_tables;
  double* d_partials = nullptr;     // [1024] block partial sums for nsr_lr_metrics
  uint32_t* d_jet = nullptr;        // [256] packed COLORMAP_JET entries (nsr_assemble_frame)
  long long* d_clk = nullptr;       // [4] (clock64, globaltimer) at entry / exit of CTA 0 of the last k_tc_pass launch
  nsr::NetImages net[2];
  std::vector<int64_t> param_numel;
  int sm_count = 0;
  std::atomic<int64_t> launches{0};
  long long* trace_buf = nullptr;
  int debug_flags = 0;
  int tc_cluster = 2;               // CTAs per cluster of k_tc_pass (2: the pair shares one multicast weight stream;
                                    // NSR_TC_CLUSTER=1 in the environment switches it off for A/B runs)
  int tc_fused = 1;                 // coarse + fine pass in one launch where the option set allows (64 + 64 samples);
                                    // NSR_TC_FUSED=0 keeps the two-launch path (A/B runs, bit-equality tests)
  std::string err;
  // nsr_render_host state (library-owned staging)
  cudaStream_t hs[2] = {nullptr, nullptr};
  cudaEvent_t hev[2] = {nullptr, nullptr};
  float* pin_in[2] = {nullptr, nullptr};
  float* pin_out[2] = {nullptr, nullptr};
  float* dev_in[2] = {nullptr, nullptr};
  float* dev_out[2] = {nullptr, nullptr};
  void* dev_ws[2] = {nullptr, nullptr};
  size_t host_chunk = 0, host_ws_bytes = 0;
  float* frame_rays = nullptr;      // nsr_render_pose_host: device rays of one frame
  size_t frame_rays_cap = 0;
  cudaEvent_t frame_ev = nullptr;
  // nsr_pack_weights records an event on the caller's stream; the host-buffer pipeline's own streams wait for it
  cudaEvent_t pack_ev[2] = {nullptr, nullptr};
  bool pack_pending[2] = {false, false};
  std::vector<std::pair<void*, int64_t>> train_stash;   // workspaces filled by nsr_render_train -> n_rays
  // nsr_backward: the 13 dW GEMMs of a net are independent of each other; they are spread over the caller's stream and
  // these two library-owned ones (event fork / join around them) so that one launch's tail overlaps the next one's ramp
  cudaStream_t dw_st[2] = {nullptr, nullptr};
  cudaEvent_t dw_fork = nullptr;
  cudaEvent_t dw_done[2] = {nullptr, nullptr};
};

namespace nsr {

// ---- SIMT path (nsr_simt.cu) ----
size_t simt_blob_floats(const NsrHandle_* h);
cudaError_t simt_pack(NsrHandle_* h, int which, const float* const* params, cudaStream_t st);
cudaError_t simt_mlp(NsrHandle_* h, int which, const float* rays, int64_t n_rays, int ray_stride,
                     const float* z, int S, float* raw, cudaStream_t st);

// ---- tcgen05 path (nsr_tc.cu) ----
// layout of NetImages::tc_consts (floats)
constexpr int kcBias = 0;        // 8 x 256 trunk, then final 256, then dir 128
constexpr int kcBiasFinal = 2048;
constexpr int kcBiasDir = 2304;
constexpr int kcWsig = 2432;     // 256
constexpr int kcWrgb = 2688;     // 3 x 128
constexpr int kcMisc = 3072;     // b_sigma, b_rgb[3]
constexpr int kcSmemFloats = 3080;
constexpr int kcWdd = 3080;      // dir-part of dir_encoding weight: [128][28]
constexpr int kcTotal = 3080 + 128 * 28;
bool tc_supported(const NsrConfig& cfg, std::string* why);
size_t tc_image_bytes(const NsrHandle_* h);
cudaError_t tc_pack(NsrHandle_* h, int which, const float* const* params, cudaStream_t st);
cudaError_t tc_init(NsrHandle_* h);     // per-function attributes (dynamic shared memory opt-in), once per handle
// Fused pass: sampling/encoding -> MLP -> compositing (-> resampling).
//   z_in   : [N,S] or null (coarse pass computes z itself; u_jitter optional)
//   z_next : [N, S+n_imp] or null; written when do_resample
struct TcPassArgs {
  const float* rays; int64_t n_rays; int ray_stride;
  const float* z_in; int S;
  const float* u_jitter; const float* noise; const float* u_resample;
  int do_resample;
  float* comp_rgb; float* depth; float* opacity; float* weights; float* raw; float* z_next;
  long long* trace;   // debug timeline buffer (NSR_TC_TRACE builds), else null
  int debug_flags;
  // training stash (all three set together, or all null): see TcKernelArgs in nsr_tc.cu
  uint8_t* stash_enc = nullptr; uint8_t* stash_h = nullptr; uint8_t* stash_dir = nullptr; uint32_t* stash_mask = nullptr;
  float* z_out = nullptr;
};
cudaError_t tc_pass(NsrHandle_* h, int which, const TcPassArgs& a, cudaStream_t st);

// Fused frame: coarse pass + resampling + fine pass of a ray batch in ONE launch (k_tc_pass<.., FUSED>), optionally with the
// rays generated in the front-end from a pose and with the s x s box average in the compositing epilogue.
//   rays     : [N, ray_stride] device rays, or null -> generate_ray(*rg, idx)
//   z_fine   : [N, 128] (required: the fine tiles read the merged z-values the coarse tiles write)
//   c_* / f_*: HR outputs of the coarse / fine pass, any of them null
//   lr_*     : box-averaged outputs [N / s^2], any of them null (N must then be a multiple of s^2)
struct TcFrameArgs {
  const float* rays = nullptr; int64_t n_rays = 0; int ray_stride = 0;
  const RayGenParams* rg = nullptr; int64_t rg_first = 0;   // row of ray 0 within the frame's raster order
  int s = 1;
  const float* u_jitter = nullptr; const float* noise_c = nullptr; const float* noise_f = nullptr; const float* u_resample = nullptr;
  float* z_fine = nullptr;
  float* c_rgb = nullptr; float* c_depth = nullptr; float* c_opacity = nullptr; float* c_weights = nullptr;
  float* f_rgb = nullptr; float* f_depth = nullptr; float* f_opacity = nullptr; float* f_weights = nullptr;
  float* lr_rgb_c = nullptr; float* lr_depth_c = nullptr; float* lr_rgb_f = nullptr; float* lr_depth_f = nullptr;
  long long* trace = nullptr; int debug_flags = 0;
};
bool tc_frame_supported(const NsrHandle_* h, int s);
cudaError_t tc_frame(NsrHandle_* h, const TcFrameArgs& a, cudaStream_t st);
int frame_schedule(long long pairs, int unit_rays, int first, int stride, long long* out);

// ---- training path (nsr_train.cu) ----
size_t train_wt_bytes();
cudaError_t train_pack_wt(NsrHandle_* h, int which, const float* const* params_dev, cudaStream_t st);

// ---- shared small kernels (nsr_api.cu) ----
cudaError_t launch_composite(NsrHandle_* h, const float* raw, const float* z, const float* noise,
                             int64_t n_rays, int S, int do_resample, const float* u_resample,
                             float* comp_rgb, float* depth, float* opacity, float* weights,
                             float* z_next, cudaStream_t st);

}  // namespace nsr
