/*
 * nsr.h -- C ABI of libnsr_b200: the B200-native volumetric-render hot path of
 * NeRF-SR (ray generation -> stratified / inverse-CDF sampling -> positional
 * encoding -> coarse/fine 8+1-layer MLP on tcgen05 tensor cores -> alpha
 * compositing -> s x s box average).
 *
 * The reference (cwchenwang/NeRF-SR) has no FFI layer: its boundary for this
 * path is the Python method NeRFDownXModel.forward_rays(rays[N,8]) -> dict
 * (models/nerf_downX_model.py:280-313, called from :318/:322/:580/:604 through
 * utils/utils.py:130-152 chunk_batch) and the finer seams below it.  Every entry
 * point here names the reference interface it replaces.  The ctypes binding a
 * reference maintainer would add is shown in INTEGRATION.md and shipped as
 * nerf_sr_b200/_lib.py.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types; no exceptions.
 *  - All tensor pointers are DEVICE pointers (fp32, row-major, contiguous)
 *    unless the name says host.  The caller owns every buffer, including outputs
 *    and workspace; the library borrows them for the stream-ordered call and owns
 *    only its packed-weight images.
 *  - Every call is asynchronous on the given stream (no host sync) except
 *    nsr_render_host() and nsr_create()/nsr_destroy().
 *  - Return value: 0 = NSR_OK, otherwise an NsrStatus; the message is available
 *    from nsr_last_error().  The library never exits or aborts.
 *  - No global mutable state: distinct handles are fully independent (one per thread / per GPU).  A
 *    single handle may be used by one thread at a time (it carries the last-error string and the
 *    nsr_render_host staging buffers); calls on it are ordered by the streams the caller passes.
 */
#ifndef NSR_B200_H_
#define NSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSR_ABI_VERSION 1

#if defined(__GNUC__)
#define NSR_API __attribute__((visibility("default")))
#else
#define NSR_API
#endif

typedef struct NsrHandle_ NsrHandle;   /* opaque */
typedef void* NsrStream;               /* cudaStream_t */

typedef enum NsrStatus {
  NSR_OK = 0,
  NSR_ERR_INVALID_ARG = 1,     /* null pointer, bad size, bad enum              */
  NSR_ERR_UNSUPPORTED = 2,     /* option combination this build cannot run      */
  NSR_ERR_NOT_PACKED = 3,      /* render called before nsr_pack_weights         */
  NSR_ERR_WORKSPACE = 4,       /* workspace too small                           */
  NSR_ERR_CUDA = 5,            /* a CUDA runtime call failed                    */
  NSR_ERR_NO_DEVICE = 6        /* no sm_100 device / kernel image not loadable  */
} NsrStatus;

/* Arithmetic of the MLP GEMMs (parity needs ~fp32 accuracy). */
typedef enum NsrPrecision {
  NSR_PREC_FP32_SIMT = 0,   /* fp32 FFMA on CUDA cores; generic D/W/skips      */
  NSR_PREC_BF16X3_TC = 1,   /* tcgen05 kind::f16, bf16 hi/lo split, 3 MMAs     */
  NSR_PREC_FP16X3_TC = 2,   /* tcgen05 kind::f16, fp16 hi/lo split, 3 MMAs     */
  NSR_PREC_BF16_TC   = 3    /* single bf16 pass: fast, NOT parity grade        */
} NsrPrecision;

/*
 * Option surface the path reads from the reference's `opt`
 * (models/nerf_model.py:46-72, models/networks.py:124-128,
 *  models/embedding.py:17-18, models/nerf_downX_model.py:125).
 */
typedef struct NsrConfig {
  uint32_t struct_size;        /* = sizeof(NsrConfig); ABI guard               */
  int32_t  device;             /* CUDA device ordinal                          */
  int32_t  precision;          /* NsrPrecision                                 */
  /* VanillaMLP (models/networks.py:121-180) */
  int32_t  D;                  /* --D, trunk depth (8)                         */
  int32_t  W;                  /* --W, trunk width (256)                       */
  uint32_t skips_mask;         /* bit i set <=> i in --skips (default 1<<4)    */
  int32_t  no_dir;             /* --no_dir                                     */
  int32_t  color_activation;   /* 0 sigmoid, 1 none                            */
  /* PositionalEncoding (models/embedding.py:28-42) */
  int32_t  deg_pos;            /* --deg_pos (10)                               */
  int32_t  deg_dir;            /* --deg_dir (4)                                */
  int32_t  no_xyz;             /* --no_xyz                                     */
  int32_t  no_logscale;        /* --no_logscale                                */
  /* sampling / rendering */
  int32_t  n_coarse;           /* --N_coarse (64)                              */
  int32_t  n_importance;       /* --N_importance (64); 0 => coarse only        */
  int32_t  lindisp;            /* --lindisp                                    */
  int32_t  white_bkgd;         /* --white_bkgd                                 */
  int32_t  sigma_activation;   /* 0 relu, 1 softplus(x-1) (rendering.py:70-73) */
  int32_t  gamma_correct;      /* --gamma_correct (nerf_downX_model.py:271)    */
  float    noise_std;          /* --noise_std                                  */
  int32_t  viewdir_offset;     /* column of the view direction in a ray row:
                                  3 = NeRFDownXModel (rays[:,3:6], :286),
                                  8 = NeRFModel (rays[:,8:11], nerf_model.py:213) */
  int32_t  reserved[8];
} NsrConfig;

/*
 * Explicit random inputs for train mode, in the reference's draw order
 * (models/utils.py:41, :210, :73, :210).  Any pointer may be null:
 *   u_coarse null  => no stratified jitter (eval, randomized=False)
 *   u_fine   null  => u = linspace(0,1,n_importance) (eval)
 *   noise_*  null  => no sigma noise
 */
typedef struct NsrRng {
  const float* u_coarse;       /* [N, n_coarse]   U[0,1)                       */
  const float* noise_coarse;   /* [N, n_coarse]   N(0,1)                       */
  const float* u_fine;         /* [N, n_importance]                            */
  const float* noise_fine;     /* [N, n_coarse+n_importance]                   */
} NsrRng;

/*
 * Outputs of forward_rays (models/nerf_downX_model.py:293-311), same shapes and
 * dtypes as the reference's dict.  Null pointers are skipped (e.g. drop the
 * per-sample weight maps for inference).  fine_* must be null or are ignored
 * when n_importance == 0.  z_fine is an extra: the merged fine z-values
 * (models/utils.py:93), useful for depth debugging and teacher-forced tests.
 */
typedef struct NsrOutputs {
  float* coarse_comp_rgbs;     /* [N,3]                                        */
  float* coarse_depth;         /* [N]                                          */
  float* coarse_opacity;       /* [N]                                          */
  float* coarse_weights;       /* [N, n_coarse]                                */
  float* fine_comp_rgbs;       /* [N,3]                                        */
  float* fine_depth;           /* [N]                                          */
  float* fine_opacity;         /* [N]                                          */
  float* fine_weights;         /* [N, n_coarse+n_importance]                   */
  float* z_fine;               /* [N, n_coarse+n_importance] (optional)        */
} NsrOutputs;

/* One rendered pass (VolumetricRenderer.forward outputs, models/rendering.py:75-111). */
typedef struct NsrPassOutputs {
  float* comp_rgbs;            /* [N,3]                                        */
  float* depth;                /* [N]                                          */
  float* opacity;              /* [N]                                          */
  float* weights;              /* [N,S]                                        */
  float* raw;                  /* [N,S,4] rgb(after colour act), raw sigma
                                  = VanillaMLP.forward output (networks.py:224);
                                  optional                                     */
} NsrPassOutputs;

/* ---- lifecycle ----------------------------------------------------------- */

/* ABI version of the loaded library (compare with NSR_ABI_VERSION). */
NSR_API int nsr_abi_version(void);

/* Replaces: NeRFDownXModel.__init__'s construction of netCoarse/netFine,
 * embeddings and renderer from `opt` (models/nerf_downX_model.py:178-197).
 * Validates the option combination; NSR_ERR_UNSUPPORTED for combinations this
 * build cannot run bit-faithfully (the caller then keeps the reference path --
 * never a silent difference). */
NSR_API int nsr_create(const NsrConfig* cfg, NsrHandle** out_handle);
NSR_API int nsr_destroy(NsrHandle* h);

/* Last error message for this handle (or for a failed nsr_create when h==NULL;
 * thread-local).  Never null. */
NSR_API const char* nsr_last_error(const NsrHandle* h);

/* Number of parameter tensors per MLP and their element counts, in the
 * reference's state_dict order (models/networks.py:149-180):
 * xyz_encoding_{1..D}.0.{weight,bias}, xyz_encoding_final.{weight,bias},
 * dir_encoding.0.{weight,bias}, sigma.{weight,bias}, rgb.0.{weight,bias}. */
NSR_API int nsr_param_count(const NsrHandle* h);
NSR_API int64_t nsr_param_numel(const NsrHandle* h, int index);

/* Replaces: load_networks / the implicit use of net.state_dict()
 * (models/base_model.py:198-219).  `which`: 0 = netCoarse, 1 = netFine.
 * param_ptrs[i] = device pointer of the i-th state_dict tensor (fp32,
 * contiguous, [out,in] row-major for weights).  Repacks into the kernel's
 * swizzled hi/lo images on `stream`; call again whenever parameters change
 * (every optimiser step in training). */
NSR_API int nsr_pack_weights(NsrHandle* h, int which, const float* const* param_ptrs,
                     int n_params, NsrStream stream);

/* ---- the hot path -------------------------------------------------------- */

/* Bytes of caller-provided scratch nsr_render needs for n_rays rays. */
NSR_API size_t nsr_workspace_bytes(const NsrHandle* h, int64_t n_rays);

/* Replaces: NeRFDownXModel.forward_rays (models/nerf_downX_model.py:280-313)
 * and NeRFModel.forward_rays (models/nerf_model.py:207-236), including the
 * chunk_batch loops around it (utils/utils.py:130-152): n_rays is unbounded,
 * no ray_chunk / point_chunk is needed because no [P,*] intermediate reaches HBM.
 * rays: [n_rays, ray_stride] rows (o3, d3, near, far[, viewdir3]); ray_stride
 * is 8 or 11.  rng may be null (eval mode). */
NSR_API int nsr_render(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride,
               const NsrRng* rng, const NsrOutputs* out,
               void* workspace, size_t workspace_bytes, NsrStream stream);

/* Replaces: render_rays(model, xyz, dir_embedded) + add_gaussian_noise +
 * self.renderer(rgb, sigma, z_vals, white_bkgd) for ONE network with caller-
 * supplied z-values (models/nerf_downX_model.py:260-278,289-291; rendering.py:
 * 75-111).  This is the teacher-forced seam of the parity protocol.
 * z_vals: [n_rays, n_samples]; noise: [n_rays, n_samples] or null. */
NSR_API int nsr_render_pass(NsrHandle* h, int which, const float* rays, int64_t n_rays,
                    int ray_stride, const float* z_vals, int n_samples,
                    const float* noise, const NsrPassOutputs* out,
                    void* workspace, size_t workspace_bytes, NsrStream stream);

/* Replaces: sample_along_rays z-values (models/utils.py:17-44).
 * z_out: [n_rays, n_coarse]; u may be null. */
NSR_API int nsr_sample_coarse(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride,
                      const float* u, float* z_out, NsrStream stream);

/* Replaces: resample_along_rays z-values (models/utils.py:47-95).
 * z_in/weights: [n_rays, n_coarse]; u: [n_rays, n_importance] or null;
 * z_out: [n_rays, n_coarse+n_importance] sorted. */
NSR_API int nsr_resample(NsrHandle* h, const float* z_in, const float* weights, int64_t n_rays,
                 const float* u, float* z_out, NsrStream stream);

/* Replaces: PositionalEncoding.__call__ (models/embedding.py:44-63).
 * x: [n, 3] -> out: [n, 3*2*deg + (no_xyz?0:3)]. */
NSR_API int nsr_posenc(NsrHandle* h, const float* x, int64_t n, int deg, float* out, NsrStream stream);

/* Replaces: comp_low_res_output's box average (models/nerf_downX_model.py:337-348):
 * in [n_lr*s*s, channels] (sub-pixels contiguous) -> out [n_lr, channels]. */
NSR_API int nsr_box_average(NsrHandle* h, const float* in, int64_t n_lr, int s, int channels,
                    float* out, NsrStream stream);

/* Replaces: get_ray_directions + get_rays (+ get_ndc_rays) + the HR->(LR,s*s)
 * grouping for one pose (models/utils.py:98-196; data/blender_downX_dataset.py:
 * 207-215; data/llff_downX_dataset.py:473-490).  c2w: HOST pointer to 12 floats
 * (3x4 row-major).  ndc != 0: project to NDC at near plane 1.0, then near=0,
 * far=1.  rays_out: [H*W, 8] in LR-pixel-major / sub-pixel-minor order. */
NSR_API int nsr_generate_rays(NsrHandle* h, const float* c2w_host, int H, int W, float focal,
                      int s, int ndc, float near_plane, float far_plane,
                      float* rays_out, NsrStream stream);

/* The same with the remaining dataset options (scope row f-4): --use_pixel_centers (options/base_options.py:59;
 * models/utils.py:114) and --unified_dir (data/llff_downX_dataset.py:273-277: one direction per LR pixel from the
 * (H/s, W/s) raster with focal // s, repeated over its sub-pixels). */
typedef struct NsrRayGen {
  uint32_t struct_size;        /* = sizeof(NsrRayGen)                           */
  int32_t  H, W;               /* HR raster (opt.img_wh reversed)               */
  int32_t  s;                  /* --downscale                                   */
  float    focal;
  int32_t  ndc;                /* forward-facing scene: NDC, near 0, far 1      */
  float    near_plane, far_plane;
  int32_t  use_pixel_centers;  /* --use_pixel_centers (reference default: 1)    */
  int32_t  unified_dir;        /* --unified_dir                                 */
  int32_t  reserved[6];
} NsrRayGen;
NSR_API int nsr_generate_rays_ex(NsrHandle* h, const float* c2w_host, const NsrRayGen* spec, float* rays_out,
                                 NsrStream stream);

/* Replaces: one iteration of the test sweep up to the images -- the dataset's ray generation for a pose
 * (data/{blender,llff}_downX_dataset.py test branch; models/utils.py:98-196), forward() = chunk_batch(forward_rays)
 * (models/nerf_downX_model.py:316-321) and comp_low_res_output's s x s box average (:337-348) -- as ONE kernel launch
 * where the option set allows (tensor-core precision, 64 + 64 samples, s in {1, 2, 4}): the fine tiles consume the z-values
 * the same CTA's coarse tiles produced, rays are generated in the kernel's front-end when `rays` is null, and the LR image
 * leaves the compositing epilogue directly.  Other option sets run the same computation as separate launches; results are
 * bit-identical either way (tests/test_gpu_fused_frame.py).
 *   rays      : [n_rays, ray_stride] device rays, or null -> rays of pose `c2w_host` (3x4 row-major, host) on the raster
 *               `spec` (n_rays / ray_stride are then ignored: H*W rays of 8 columns)
 *   out       : HR outputs, any pointer null (with the fused kernel a null HR output is simply not written)
 *   lr        : box-averaged outputs [n_rays / s^2], any pointer null; ignored when s == 1
 *   workspace : nsr_frame_workspace_bytes(h, n_rays, rays == null) bytes (fine z-values; on option sets that take
 *               separate launches also the generated rays and any HR composite the box average needs but `out` omits) */
typedef struct NsrLrOutputs {
  float* coarse_rgb;           /* [N/s^2,3]                                    */
  float* coarse_depth;         /* [N/s^2]                                      */
  float* fine_rgb;             /* [N/s^2,3]                                    */
  float* fine_depth;           /* [N/s^2]                                      */
} NsrLrOutputs;
NSR_API size_t nsr_frame_workspace_bytes(const NsrHandle* h, int64_t n_rays, int from_pose);
NSR_API int nsr_render_frame(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const float* c2w_host,
                             const NsrRayGen* spec, int s, const NsrRng* rng, const NsrOutputs* out,
                             const NsrLrOutputs* lr, void* workspace, size_t workspace_bytes, NsrStream stream);

/* Replaces (scope row f-2, the LR-image + loss epilogue): comp_low_res_output's box average followed by
 * ColorMSELoss and PSNR against the LR targets (models/nerf_downX_model.py:337-340,357,380;
 * models/criterions.py:7-15,27-36).  hr_rgb: [n_lr*s*s, 3] composite colours (sub-pixels contiguous);
 * target_lr: [n_lr, 3]; lr_rgb_out: [n_lr, 3] or null; metrics_out: device float[2] = {mse, psnr} with
 * mse = mean((lr - target)^2) over n_lr*3 elements, psnr = -10*log10(mse).  No host sync. */
NSR_API int nsr_lr_metrics(NsrHandle* h, const float* hr_rgb, const float* target_lr, int64_t n_lr, int s,
                           float* lr_rgb_out, float* metrics_out, NsrStream stream);

/* Replaces (scope row f-3, frame assembly): unflatten_reshape + depth2im + the [pred | gt | depth] concat of
 * calculate_vis + _save_image's float -> uint8 conversion (models/nerf_downX_model.py:410-450;
 * utils/visualizer.py:40-60,164-176), i.e. everything between the render outputs and the bytes handed to the
 * PNG / GIF encoder.  rgb [H*W,3], depth [H*W], gt [H*W,3] or null: rows in the renderer's LR-pixel-major /
 * sub-pixel-minor order (s = 1: plain raster).  out_rgb8: [H][W*(gt?3:2)][3] uint8 in the reference's channel
 * order before its final RGB->BGR swap (the depth panel is OpenCV COLORMAP_JET of
 * uint8(255 (d - near) / max(far - near, 1e-8)), numpy cast semantics incl. wrap-around).  depth_mat (optional):
 * [H][W] fp32 = np.nan_to_num of the raster depth, the matrix the reference saves as *-depth*.npz. */
NSR_API int nsr_assemble_frame(NsrHandle* h, const float* rgb, const float* depth, const float* gt, int H, int W, int s,
                               float near_plane, float far_plane, uint8_t* out_rgb8, float* depth_mat, NsrStream stream);

/* ---- host-buffer convenience (the end-to-end call) ------------------------ */

/* forward over a whole frame / batch with HOST buffers: stages rays through
 * pinned memory in chunks, overlaps H2D / render / D2H on internal streams and
 * (optionally) box-averages on the device so only [n_rays/s^2] rows return.
 * rays_host: [n_rays, ray_stride].  Any output pointer may be null.
 *   rgb_host:   [n_out, 3]   fine (or coarse if n_importance==0) composite
 *   depth_host: [n_out]
 * with n_out = n_rays/(s*s) if s > 1 else n_rays.  Synchronous. */
NSR_API int nsr_render_host(NsrHandle* h, const float* rays_host, int64_t n_rays, int ray_stride,
                    int s, float* rgb_host, float* depth_host);

/* Same pipeline for one camera pose (scope row f-4: dataset-side ray construction on the device, so a
 * test sweep uploads 48 bytes per frame instead of 32 bytes per ray): nsr_generate_rays for the whole
 * H x W raster (data/blender_downX_dataset.py:207-215, data/llff_downX_dataset.py:473-490), render, box
 * average, D2H.  rgb_host: [H*W/s^2, 3], depth_host: [H*W/s^2] (either may be null).  Synchronous. */
NSR_API int nsr_render_pose_host(NsrHandle* h, const float* c2w_host, int H, int W, float focal, int s, int ndc,
                                 float near_plane, float far_plane, float* rgb_host, float* depth_host);

/* ---- training (scope row f-1: backward + fused optimiser) ------------------ */

/* dL/d(outputs of forward_rays).  Null = that output does not enter the loss.  Per-sample weight maps
 * are not differentiable through this interface (the reference's losses never use them). */
typedef struct NsrOutGrads {
  const float* coarse_comp_rgbs;   /* [N,3] */
  const float* coarse_depth;       /* [N]   */
  const float* coarse_opacity;     /* [N]   */
  const float* fine_comp_rgbs;     /* [N,3] */
  const float* fine_depth;         /* [N]   */
  const float* fine_opacity;       /* [N]   */
} NsrOutGrads;

/* Elements of one net's flat gradient (= sum of nsr_param_numel): tensors in state_dict order. */
NSR_API int64_t nsr_grad_numel(const NsrHandle* h);

/* Bytes of caller-provided scratch that nsr_render_train fills (activation stash) and nsr_backward
 * consumes; the SAME buffer must be passed to both. */
NSR_API size_t nsr_train_workspace_bytes(const NsrHandle* h, int64_t n_rays);

/* Replaces: forward_rays in train mode with autograd recording (models/nerf_downX_model.py:280-313 under
 * torch.enable_grad()): same outputs as nsr_render, and every tile's MLP inputs / activations are kept in
 * train_ws for the backward.  rng as in nsr_render (train mode draws). */
NSR_API int nsr_render_train(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const NsrRng* rng,
                             const NsrOutputs* out, void* train_ws, size_t train_ws_bytes, NsrStream stream);

/* Replaces: loss_tot.backward() through forward_rays (models/nerf_downX_model.py:390-396): given
 * dL/d(outputs), writes dL/d(parameters) of netCoarse / netFine as flat fp32 buffers (state_dict order,
 * nsr_grad_numel elements each; overwritten, not accumulated).  rays / rng: the ones passed to
 * nsr_render_train (only the sigma noise is read).  Stream-ordered on `stream` like every other call; internally the
 * independent weight-gradient GEMMs of a net are spread over `stream` and two library-owned streams, forked and joined
 * by events inside the call (nothing runs past the point where work queued on `stream` after this call may start). */
NSR_API int nsr_backward(NsrHandle* h, const float* rays, int64_t n_rays, int ray_stride, const NsrRng* rng,
                         const NsrOutGrads* g, float* grad_coarse, float* grad_fine, void* train_ws,
                         size_t train_ws_bytes, NsrStream stream);

/* Replaces: comp_low_res_output + ColorMSELoss * lambda + PSNR + the loss's backward down to the HR composite
 * colours (models/nerf_downX_model.py:337-340,357-359,380-382).  metrics_out: device float[2] =
 * {lambda * mse, psnr}; g_hr_out: [n_lr*s*s, 3] = d(lambda * mse)/d(hr_rgb) (or null). */
NSR_API int nsr_lr_loss_grad(NsrHandle* h, const float* hr_rgb, const float* target_lr, int64_t n_lr, int s, float lambda,
                             float* lr_rgb_out, float* metrics_out, float* g_hr_out, NsrStream stream);

/* Replaces (scope row f-2, completed): every term of calculate_losses for ONE net's outputs and its backward down to
 * the HR composite colour / depth (models/nerf_downX_model.py:326-378):
 *   comp_low_res_output's box averages (:337-348), lambda_*_mse * ColorMSELoss + PSNR (:357-359,380-382),
 *   --use_var_loss: sum over LR pixels and channels of torch.var over the s*s sub-pixel colours (:331-335,374-375),
 *   --use_depth_var_loss: the same for depth / self.far (:349-353,376-378),
 *   --sisr_path: ColorMSELoss of the HR colours against data_rgbs_sr (:364-367), lambda_hr = 1;
 *   --with_ref: the same HR ColorMSELoss for the reference-view batch (out_ref_* vs data_ref_rgbs, :369-372):
 *               target_lr null, lambda_hr = 1 / s^2.
 * A zero lambda_var / lambda_depth_var switches that term off; target_hr null switches the HR term off;
 * target_lr null switches the LR term (box-average MSE, PSNR) off. */
typedef struct NsrLossTerms {
  uint32_t struct_size;        /* = sizeof(NsrLossTerms)                        */
  int32_t  s;                  /* --downscale                                   */
  float    lambda_mse;         /* --lambda_coarse_mse | --lambda_fine_mse       */
  float    lambda_var;         /* --lambda_*_var when --use_var_loss, else 0    */
  float    lambda_depth_var;   /* --lambda_*_depth_var when --use_depth_var_loss */
  float    far_plane;          /* self.far = rays[0,7] (:284)                   */
  float    lambda_hr;          /* weight of the HR-target MSE (1 | 1/s^2)       */
  int32_t  reserved[5];
} NsrLossTerms;

/* hr_rgb [n_lr*s*s,3], hr_depth [n_lr*s*s] (null unless a depth output / term is wanted), target_lr [n_lr,3] or null,
 * target_hr [n_lr*s*s,3] or null (at least one target).  Outputs (each may be null except metrics_out): lr_rgb_out [n_lr,3],
 * lr_depth_out [n_lr], g_rgb_out [n_lr*s*s,3] = d total / d hr_rgb, g_depth_out [n_lr*s*s] = d total / d hr_depth,
 * metrics_out: device float[8] = {lambda_mse * mse, psnr, var_sum, depth_var_sum, lambda_hr * mse_hr, total, 0, 0}
 * with total = lambda_mse * mse + lambda_hr * mse_hr + lambda_var * var_sum + lambda_depth_var * depth_var_sum.
 * No host sync. */
NSR_API int nsr_loss_epilogue(NsrHandle* h, const float* hr_rgb, const float* hr_depth, const float* target_lr,
                              const float* target_hr, int64_t n_lr, const NsrLossTerms* terms, float* lr_rgb_out,
                              float* lr_depth_out, float* metrics_out, float* g_rgb_out, float* g_depth_out, NsrStream stream);

/* Replaces: nn.utils.clip_grad_norm_ over chain(netCoarse, netFine) (models/nerf_downX_model.py:404-405).
 * coef_out: device float[2] = {min(1, max_norm / (total_norm + 1e-6)), total_norm}; grad_b may be null. */
NSR_API int nsr_clip_coef(NsrHandle* h, const float* grad_a, const float* grad_b, int64_t numel, float max_norm,
                          float* coef_out, NsrStream stream);

/* Replaces: torch.optim.Adam.step for one net (models/nerf_downX_model.py:201-204,408), in place on the
 * caller's parameter tensors (param_ptrs as in nsr_pack_weights).  exp_avg / exp_avg_sq: flat state
 * buffers (caller-owned, zero-initialised); step: 1-based count after the increment.  clip_coef_dev (or
 * null) multiplies the gradient (norm clipping); clip_value > 0 clamps it (clip_grad_value_).  Call
 * nsr_pack_weights afterwards. */
NSR_API int nsr_adam_step(NsrHandle* h, float* const* param_ptrs, int n_params, const float* grad_flat, float* exp_avg,
                          float* exp_avg_sq, int64_t step, float lr, float beta1, float beta2, float eps,
                          const float* clip_coef_dev, float clip_value, NsrStream stream);

/* Test seams of the backward GEMMs ("tile image" = [ceil(rows/128)][cols/64][hi 16 KB | lo 16 KB], 128-B rows,
 * XOR-swizzled 16-B chunks; see nsr_train.cu). */
/* Byte offsets (from the 256-aligned base of train_ws) of the stash regions, per pass p = 0 coarse / 1 fine:
 * out18 = {enc_p, h_p, dir_p, raw_p, z_p, n_tiles_p} x 2, then dhead, dzdir, g0, g1, mask_0, mask_1.
 * h_p = [9][n_tiles][4 chunks]; mask_p = [8][n_tiles][128][8] uint32 (1 bit per activation of h_1..h_8). */
NSR_API int nsr_debug_train_layout(const NsrHandle* h, int64_t n_rays, int64_t* out18);
/* bits_out[ceil(rows/128)*128][8] uint32: bit j of word w of a row = x[row][32 w + j] > 0  (x: [rows,256]). */
NSR_API int nsr_debug_relu_bits(NsrHandle* h, const float* x, int64_t n_rows, void* bits_out, NsrStream stream);
NSR_API int nsr_debug_pack_image(NsrHandle* h, const float* src, int64_t n_rows, int n_cols, int ld, void* image, NsrStream stream);
NSR_API int nsr_debug_unpack_image(NsrHandle* h, const void* image, int64_t n_rows, int n_cols, int ld, float* dst, NsrStream stream);
NSR_API int nsr_debug_dx(NsrHandle* h, int which, int layer_idx, const void* a_img, void* out_img, const void* mask_bits,
                         const float* dsig, const float* wsig, int64_t n_rows, NsrStream stream);
NSR_API int nsr_debug_dw(NsrHandle* h, const void* a_img, int a_cols, int blk0, int blk1, const void* b_img, int b_cols,
                         float* out, float* bias_out, int64_t n_rows, void* scratch, size_t scratch_bytes, NsrStream stream);

/* Debug: device buffer (>= 16*512 int64) that trace builds (-DNSR_TC_TRACE=1) fill with
 * (tag, clock64) pairs for one tile of CTA 0; ignored by normal builds.  tools/tc_trace.py. */
NSR_API int nsr_debug_set_trace(NsrHandle* h, long long* device_buffer);
/* Debug: timing-experiment switches (bit 0: weight producer skips its bulk copies -> WRONG results). */
NSR_API int nsr_debug_set_flags(NsrHandle* h, int flags);

/*
 * Data-parallel gradient all-reduce (the only collective of the path).  Replaces the bucketed NCCL
 * all-reduce + division by the world size that DistributedDataParallel performs during loss_tot.backward()
 * (models/networks.py:72-86; per-rank batch = batch_size / n_gpus, data/__init__.py:94-99).
 * One process per GPU.  Every rank creates a comm of the same size; its buffer (`nsr_comm_buffer`, device memory owned
 * by the comm) is where nsr_backward should write the flat gradients (grad_coarse = buffer, grad_fine = buffer + numel).
 * Ranks exchange the 64-byte handles of nsr_comm_export by any means (torch.distributed.all_gather_object in
 * nerf_sr_b200/parallel.py) and map each other's buffers with nsr_comm_connect_ipc (CUDA IPC, NVLink peer access).
 * nsr_comm_allreduce_mean is then ONE kernel on `stream`: in place, mean over ranks, summed in fixed rank order, so the
 * result is bit-identical on every rank.  Every rank must call it the same number of times (it is a collective); it
 * synchronises with the peers on the device only, never with the host.
 * nsr_comm_connect_ptrs wires comms that live in ONE process (several handles on one or more devices: tests).
 */
typedef struct NsrComm_ NsrComm;   /* opaque */
NSR_API int nsr_comm_create(NsrHandle* h, int rank, int world, int64_t n_floats, NsrComm** out);
NSR_API int nsr_comm_destroy(NsrComm* c);
NSR_API float* nsr_comm_buffer(NsrComm* c);                 /* device pointer, >= n_floats floats, zero padded */
NSR_API int64_t nsr_comm_buffer_floats(const NsrComm* c);   /* padded length */
NSR_API int nsr_comm_export(NsrComm* c, void* handle_out64);
NSR_API int nsr_comm_connect_ipc(NsrComm* c, const void* handles64, int n_handles);   /* world x 64 bytes, rank order */
NSR_API int nsr_comm_connect_ptrs(NsrComm* c, void* const* peer_buffers, int n);      /* same-process peers */
NSR_API int nsr_comm_allreduce_mean(NsrComm* c, NsrStream stream);

/* Debug / test seam (no GPU needed): the order in which one CTA of the one-launch frame kernel visits its tiles, computed on
 * the host by the function the kernel itself uses.  The CTA owns `pairs` ray pairs (a multiple of unit_rays / 2) in work units
 * first_unit, first_unit + unit_stride, ... of unit_rays (2, 4 or 16) consecutive rays.  out: int64[3 * pairs][3] =
 * (pass: 0 coarse / 1 fine, tile index within that pass: coarse tile t = rays 2t, 2t+1; fine tile t = ray t,
 *  1 if the trip depends on the one right before it). */
NSR_API int nsr_debug_frame_schedule(int64_t pairs, int unit_rays, int first_unit, int unit_stride, int64_t* out);
/* Debug: 1 if nsr_render_frame / nsr_render_host would run the s x s box average of an n_rays batch inside the frame kernel
 * (it does when that costs the busiest CTA < 3 % more tiles than ray-pair granularity), else 0. */
NSR_API int nsr_debug_frame_lr_in_kernel(const NsrHandle* h, int64_t n_rays, int s);

/* Debug: (clock64, globaltimer ns) stamped by CTA 0 at entry and at exit of the most recent fused-pass kernel
 * (k_tc_pass) of this handle -> out4_host = {clk0, ns0, clk1, ns1}; (clk1-clk0)/(ns1-ns0) GHz is the SM clock the
 * kernel actually ran at.  Synchronises `stream` (a measurement aid, not part of the render path). */
NSR_API int nsr_debug_kernel_clock(NsrHandle* h, int64_t* out4_host, NsrStream stream);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
NSR_API int64_t nsr_launch_count(const NsrHandle* h);

#ifdef __cplusplus
}
#endif
#endif  /* NSR_B200_H_ */
