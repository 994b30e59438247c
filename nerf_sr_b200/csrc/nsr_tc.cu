// nsr_tc.cu -- the tcgen05 fused render pass (NSR_PREC_*_TC), sm_100a only.
//
// One persistent CTA per SM; a tile is 128 sampled points (2 coarse rays x 64
// samples, or 1 fine ray x 128).  Per tile, entirely on-chip:
//   front-end warps : z sampling / jitter, o + z*d, sin/cos positional encoding,
//                     hi/lo 16-bit split, written as a K-major SWIZZLE_128B UMMA
//                     operand in shared memory; per-ray view-direction bias.
//   producer thread : streams the net's pre-swizzled weight image (72 stages x
//                     32 KB: 128 rows x 64 k, hi | lo) L2 -> smem ring with
//                     cp.async.bulk + mbarrier complete_tx.
//   MMA thread      : tcgen05.mma kind::f16, M=128 N=128 K=16, fp32 accumulate in
//                     TMEM.  Activations live in TMEM as the A operand (hi and lo
//                     planes), weights are the smem B operand; each product is
//                     evaluated as hi*hi + lo*hi + hi*lo (split precision, ~fp32
//                     accuracy -- SURVEY.md section 0.6).
//   epilogue warps  : tcgen05.ld accumulator -> +bias, ReLU, hi/lo split ->
//                     tcgen05.st next layer's A operand; sigma / rgb heads on CUDA
//                     cores; then alpha compositing (one warp per ray), inverse-CDF
//                     resampling + sort-merge for the fine pass, output stores.
// Encoded features, activations and per-sample (rgb, sigma) never reach HBM.
//
// Two launch shapes share this code (template parameter FUSED):
//   one pass  : the CTA walks tiles blockIdx.x, + gridDim.x, ... of ONE net (coarse or fine; the training forward with its
//               activation stash, the teacher-forced seam nsr_render_pass, option sets outside the frame variant);
//   one frame : coarse tiles, resampling, fine tiles -- plus ray generation in the front-end and the s x s box average in
//               the compositing epilogue -- of a whole ray batch in ONE launch; the fine tiles read the z-values the same
//               CTA's coarse tiles wrote two trips earlier (trip_of / frame_trip below).  Bit-identical to the passes.
//
// TMEM map (512 columns x 128 lanes x 32 bit):
//   [  0,256) fp32 accumulator, two N-halves of 128 columns
//   [256,384) A operand, hi plane (256 k-values, 2 per column)
//   [384,512) A operand, lo plane
// The two accumulator halves let the epilogue of layer L overlap the MMAs of
// layers L and L+1 (see the schedule in mma_role()).
//
// Replaces (reference): sample_along_rays + cast_rays (models/utils.py:5-44),
// PositionalEncoding (models/embedding.py:44-63), render_rays' [P,90] concat
// (models/nerf_downX_model.py:260-278), VanillaMLP.forward (models/networks.py:
// 199-224), add_gaussian_noise (models/utils.py:199-212), VolumetricRenderer.forward
// (models/rendering.py:89-111) and resample_along_rays (models/utils.py:47-95).
#include "nsr_internal.h"
#include "nsr_tc_ptx.cuh"
#include "nsr_tc_mma.cuh"

namespace nsr {

// ---------------------------------------------------------------------------
// compile-time geometry
// ---------------------------------------------------------------------------
constexpr int kTile = 128;                  // points per tile (UMMA M)
// (kStageBytes = 32768, kPlaneBytes = 16384, kRing = 4: nsr_tc_mma.cuh)
constexpr int kStagesPerTile = 72;
constexpr int kWarpsFront = 4, kWarpsEpi = 8;
constexpr int kThreadsTc = 32 * (2 + kWarpsFront + kWarpsEpi);   // 448
// Warp roles by id.  The SM's warp arbiter favours the highest warp id among eligible warps
// (B300_MICROARCH.md), so the latency-critical single-thread roles get the top ids and the
// throughput-tolerant front-end (which works one tile ahead) the bottom ones.
constexpr int kFrontWarp0 = 0;                                   // warps 0-3
constexpr int kEpiWarp0 = kWarpsFront;                           // warps 4-11 (warp % 4 = TMEM lane quarter)
constexpr int kProducerWarp = kWarpsFront + kWarpsEpi;           // warp 12
constexpr int kMmaWarp = kProducerWarp + 1;                      // warp 13

// (layout of the fp32 consts blob: kc* in nsr_internal.h)

// shared memory map (bytes, relative to a 1024-aligned base)
constexpr int kSmRing = 0;
constexpr int kSmEnc = kSmRing + kRing * kStageBytes;            // 131072
constexpr int kSmConst = kSmEnc + 2 * kStageBytes;               // 196608
constexpr int kConstBytes = kcSmemFloats * 4;                    // 12320: one net's biases + head weights
constexpr int kSmDirBias = kSmConst + 2 * kConstBytes;           // two nets (the fused frame kernel keeps both resident)
constexpr int kSmZ = kSmDirBias + 2 * 2 * 128 * 4;
constexpr int kSmDenc = kSmZ + 2 * 128 * 4;
constexpr int kSmSig = kSmDenc + 2 * 32 * 4;
constexpr int kSmRgb = kSmSig + 128 * 4;
constexpr int kSmW = kSmRgb + 384 * 4;
constexpr int kSmTmp = kSmW + 128 * 4;
constexpr int kSmXch = kSmTmp + 128 * 4;
constexpr int kSmScratch = kSmXch + 4 * 128 * 4;
// fused frame kernel: per-ray (r, g, b, depth) of the current unit's rays, [2 passes][16 rays][4], box-averaged when an LR
// pixel's last ray is composited.  It aliases the tail of the resampler scratch: the fused kernel runs 64 + 64 samples only,
// whose scratch need is 192 floats per ray at offsets 0 and 320 -- floats [512, 640) are free.
constexpr int kSmLrAcc = kSmScratch + 512 * 4;
constexpr int kSmBar = kSmScratch + 2 * 320 * 4;
constexpr int kSmTmemPtr = kSmBar + B_COUNT * 8;
constexpr int kSmemTcBytes = kSmTmemPtr + 16;
static_assert(kSmemTcBytes <= 232448, "k_tc_pass: over the 227 KB of shared memory a CTA can opt in to");
static_assert(kConstBytes % 16 == 0 && kSmBar % 8 == 0 && kSmDirBias % 16 == 0, "alignment of the shared-memory map");

// (barrier indices B_*: nsr_tc_mma.cuh)

// ---------------------------------------------------------------------------
// weight image: stage table shared by the packer and (by construction) the MMA
// issue order in mma_role().
// ---------------------------------------------------------------------------
struct StageDesc {
  int16_t param;     // state_dict index of the weight tensor
  int16_t half;      // output rows [128*half, 128*half+128)
  int16_t col0;      // first input column of this 64-wide k chunk
  int16_t kvalid;    // valid columns (63 for the encoding chunk, else 64)
  int16_t ld;        // row length (in_features) of the weight tensor
  int16_t rows;      // valid output rows of this 128-row half (128 at W = 256; fewer / none for narrower nets: zero-padded)
};
struct StageTable { StageDesc s[kStagesPerTile]; };

// Nets narrower than the kernel's 256 features (--W 128 ..., models/networks.py:125) run on the same schedule with their weights, biases and
// head weights ZERO-PADDED to 256 / 128: a padded activation is relu(0 + 0) = 0 and contributes exact zeros to every fp32
// accumulation, so the result is what a 128-wide evaluation with the same split arithmetic would give (at the MMA cost of
// the 256-wide net -- still several times the fp32 CUDA-core path's rate).
static StageTable build_stage_table(int ch_dir, bool no_dir, int W) {
  StageTable T{};
  int n = 0;
  auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
  auto rows_of = [&](int n_out, int half) { return clampi(n_out - 128 * half, 0, 128); };
  auto push = [&](int param, int half, int col0, int kvalid, int ld, int rows) {
    T.s[n++] = StageDesc{(int16_t)param, (int16_t)half, (int16_t)col0, (int16_t)kvalid, (int16_t)ld, (int16_t)rows};
  };
  // L1 (xyz_encoding_1, K=63): half 1 first, then half 0 (lets the next tile start while the
  // previous tile's last epilogue still owns accumulator half 0)
  push(0, 1, 0, 63, 63, rows_of(W, 1));
  push(0, 0, 0, 63, 63, rows_of(W, 0));
  for (int L = 2; L <= 9; ++L) {
    const int param = 2 * (L - 1);                 // L9 = xyz_encoding_final (state_dict index 16)
    const bool skip = (L == 5);
    const int ld = skip ? 63 + W : W;
    const int off = skip ? 63 : 0;                 // cat([input_xyz(63), h(W)])  networks.py:204
    for (int h = 0; h < 2; ++h) {
      if (skip) push(param, h, 0, 63, ld, rows_of(W, h));
      for (int c = 0; c < 4; ++c) push(param, h, off + 64 * c, clampi(W - 64 * c, 0, 64), ld, rows_of(W, h));
    }
  }
  const int ld_dir = no_dir ? W : W + ch_dir;       // cat([feat(W), enc_dir]) networks.py:214
  for (int c = 0; c < 4; ++c) push(18, 0, 64 * c, clampi(W - 64 * c, 0, 64), ld_dir, rows_of(W / 2, 0));
  return T;
}

template <int FMT>
__global__ void k_tc_pack_image(StageTable T, const float* const* __restrict__ params_dev, uint8_t* __restrict__ image) {
  // one thread per (stage, row, 16-byte chunk)
  const int total = kStagesPerTile * 128 * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int s = idx / 1024, r = (idx / 8) % 128, j = idx % 8;
    const StageDesc d = T.s[s];
    const float* W = params_dev[d.param];
    const int n = 128 * d.half + r;
    __align__(16) uint16_t hi[8];
    __align__(16) uint16_t lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kl = 8 * j + e;
      const float v = (kl < d.kvalid && r < d.rows) ? W[(int64_t)n * d.ld + d.col0 + kl] : 0.f;
      Split<FMT>::apply1(v, hi[e], lo[e]);
    }
    const size_t off = (size_t)s * kStageBytes + (size_t)r * 128 + (size_t)((j ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(image + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(image + off + kPlaneBytes) = *reinterpret_cast<const uint4*>(lo);
  }
}

__global__ void k_tc_pack_consts(const float* const* __restrict__ p, float* __restrict__ c, int ch_dir, int no_dir, int W) {
  // (zero padding from the net's width W / W/2 to the kernel's 256 / 128: see build_stage_table)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nt = gridDim.x * blockDim.x;
  const int Wd = W / 2;
  for (int i = t; i < 8 * 256; i += nt) c[kcBias + i] = (i % 256 < W) ? p[2 * (i / 256) + 1][i % 256] : 0.f;
  for (int i = t; i < 256; i += nt) c[kcBiasFinal + i] = (i < W) ? p[17][i] : 0.f;
  for (int i = t; i < 128; i += nt) c[kcBiasDir + i] = (i < Wd) ? p[19][i] : 0.f;
  for (int i = t; i < 256; i += nt) c[kcWsig + i] = (i < W) ? p[20][i] : 0.f;
  for (int i = t; i < 384; i += nt) c[kcWrgb + i] = (i % 128 < Wd) ? p[22][(i / 128) * Wd + i % 128] : 0.f;
  if (t == 0) { c[kcMisc] = p[21][0]; c[kcMisc + 1] = p[23][0]; c[kcMisc + 2] = p[23][1]; c[kcMisc + 3] = p[23][2]; }
  const int ld = no_dir ? W : W + ch_dir;
  for (int i = t; i < 128 * 28; i += nt) {
    const int j = i / 28, k = i % 28;
    c[kcWdd + i] = (!no_dir && k < ch_dir && j < Wd) ? p[18][(int64_t)j * ld + W + k] : 0.f;
  }
}

bool tc_supported(const NsrConfig& c, std::string* why) {
  auto no = [&](const char* m) { if (why) *why = m; return false; };
  if (c.D != 8 || c.skips_mask != (1u << 4)) return no("needs D=8, skips=[4]");
  if (c.W != 256 && c.W != 128 && c.W != 64) return no("needs a net width W of 64, 128 or 256 (narrower nets run zero-padded)");
  if (c.deg_pos != 10 || c.deg_dir != 4 || c.no_xyz) return no("needs deg_pos=10, deg_dir=4, xyz included");
  // coarse pass: whole rays per 128-point tile (64 or 128 samples); fine pass: the same, or 192 / 256 samples in the
  // MLP-only mode (tiles cut across rays, compositing in k_composite); the in-kernel resampler's scratch holds
  // 2 N_coarse + N_importance <= 320 floats per ray (640 when a tile is one ray)
  const int sc = c.n_coarse, sf = c.n_coarse + c.n_importance;
  const bool coarse_ok = (sc == 64 || sc == 128);
  const bool fine_ok = (sf == sc) || sf == 128 || sf == 192 || sf == 256;
  const bool scratch_ok = c.n_importance == 0 || 2 * sc + c.n_importance <= (sc == 128 ? 640 : 320);
  if (!(coarse_ok && fine_ok && scratch_ok))
    return no("needs N_coarse in {64, 128} and N_coarse + N_importance in {N_coarse, 128, 192, 256}");
  return true;
}

size_t tc_image_bytes(const NsrHandle_*) { return (size_t)kStagesPerTile * kStageBytes; }

cudaError_t tc_pack(NsrHandle_* h, int which, const float* const* params, cudaStream_t st) {
  NetImages& net = h->net[which];
  // the pointer table itself has to live on the device: stage it in the consts blob tail
  const int np = (int)h->param_numel.size();
  const float** d_ptrs = reinterpret_cast<const float**>(net.tc_consts + 8192);
  cudaError_t e = cudaMemcpyAsync(d_ptrs, params, np * sizeof(float*), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  const StageTable T = build_stage_table(h->rp.ch_dir, h->cfg.no_dir != 0, h->cfg.W);
  const int fmt = (h->cfg.precision == NSR_PREC_FP16X3_TC) ? 0 : 1;
  if (fmt == 1) k_tc_pack_image<1><<<144, 256, 0, st>>>(T, d_ptrs, net.tc_image);
  else k_tc_pack_image<0><<<144, 256, 0, st>>>(T, d_ptrs, net.tc_image);
  k_tc_pack_consts<<<8, 256, 0, st>>>(d_ptrs, net.tc_consts, h->rp.ch_dir, h->cfg.no_dir, h->cfg.W);
  h->launches += 2;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// optional timeline trace (tools/tc_trace.py): -DNSR_TC_TRACE=1 stamps clock64() at protocol
// points of CTA 0, tile iteration 3 into a global buffer: 512 slots per warp.
// ---------------------------------------------------------------------------
#ifndef NSR_TC_TRACE
#define NSR_TC_TRACE 0
#endif
#if NSR_TC_TRACE
#define TR_DECL(args, it) long long* tr_ = ((args).trace && blockIdx.x == 0 && (it) == 3 && (threadIdx.x & 31) == 0) \
                                             ? (args).trace + (threadIdx.x >> 5) * 2048 : nullptr; int trn_ = 0; (void)trn_
#define TR(tag) do { if (tr_ && trn_ < 1000) { tr_[2 * trn_] = (tag); tr_[2 * trn_ + 1] = clock64(); ++trn_; } } while (0)
#define TR_PARAMS , long long* tr_, int& trn_
#define TR_ARGS , tr_, trn_
#else
#define TR_DECL(args, it)
#define TR(tag)
#define TR_PARAMS
#define TR_ARGS
#endif

// ---------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------
struct TcKernelArgs {
  const uint8_t* image;
  const float* consts;
  const SampleTables* tabs;
  RenderParams rp;
  const float* rays; long long n_rays; int ray_stride;
  const float* z_in; int S;
  const float* u_jitter; const float* noise; const float* u_resample;
  int do_resample;
  float* comp_rgb; float* depth; float* opacity; float* weights; float* raw; float* z_next;
  long long n_tiles;
  // loose = 1: S does not divide the 128-point tile (192 or 256 samples per ray: --N_importance 128 and friends).  Tiles
  // then cut across rays (point p = 128 tile + row belongs to ray p / S), the kernel is the MLP only -- it writes `raw` --
  // and compositing runs afterwards in k_composite (one warp per ray), like on the fp32 path.
  int loose;
  long long* trace;
  int debug_flags;   // bit0: producer skips the bulk copies (timing experiment: stale weights)
  // training stash (STASH kernels only; scope row f-1): every tile's MLP inputs and activations in the
  // exact shared-memory operand format (tile image: hi plane 16 KB | lo plane 16 KB per 64-feature chunk,
  // 128-B rows, 16-B chunks XOR-swizzled), so the backward GEMMs load them with plain bulk copies.
  uint8_t* stash_enc;   // [n_tiles][32 KB]            encoded xyz (63 + zero pad)
  uint8_t* stash_h;     // [9][n_tiles][4][32 KB]      h_1..h_8 (post-ReLU) and feat = xyz_encoding_final(h_8)
  uint8_t* stash_dir;   // [n_tiles][2][32 KB]         dir layer output (post-ReLU, 128 wide)
  uint32_t* stash_mask; // [8][n_tiles][128][8]        ReLU masks of h_1..h_8, one bit per activation
  float* z_out;         // [N,S] z-values actually used (coarse pass computes them on the fly), or null
  // SM-clock probe: CTA 0 stamps (clock64, globaltimer ns) at kernel entry and exit -> clk[0..3]; the ratio of the
  // deltas is the SM clock the kernel actually ran at (nsr_debug_kernel_clock; bench.py's `clocks.sm_mhz_in_kernel`)
  long long* clk;
  // ---- fused frame (k_tc_pass<.., FUSED = true>): coarse AND fine pass of a ray batch in ONE launch -------------------------
  // The fields above describe the coarse pass (S = 64, z_next = the merged 128 z-values per ray); these the fine pass.
  const uint8_t* image_f; const float* consts_f; const float* noise_f;
  float* comp_rgb_f; float* depth_f; float* opacity_f; float* weights_f;
  // box-averaged (LR) outputs, [n_rays / ss] rows, any of them null: the compositing warp that finishes the last of an LR
  // pixel's ss = s*s rays sums the ss staged values in ray order and divides (k_box_average's arithmetic, bit for bit)
  float* lr_rgb_c; float* lr_depth_c; float* lr_rgb_f; float* lr_depth_f;
  int ss;            // s * s
  int unit_rays;     // U = lcm(2, ss): rays of one work unit (whole coarse tiles AND whole LR pixels); U <= 16
  long long n_units; // ceil(n_rays / U); unit u belongs to CTA u % gridDim.x
  // in-kernel ray generation (rays == null): pose -> (o, d, near, far) of ray `idx`, the arithmetic of k_generate_rays
  RayGenParams rg; int use_rg; long long rg_first;
};

// ---------------------------------------------------------------------------
// Work order of one CTA.
//   two-launch kernels: trip `it` is tile blockIdx.x + it * gridDim.x of the launch's single pass.
//   fused frame kernel: the CTA owns units (U consecutive rays) u = blockIdx.x + k * gridDim.x, i.e. P = my_units * U / 2 ray
//   pairs; pair p is one coarse tile C_p (2 rays x 64 samples) and two fine tiles F_2p, F_2p+1 (1 ray x 128 samples, whose
//   z-values C_p's compositing warp wrote).  The front-end encodes one tile AHEAD of the MLP and composites one tile BEHIND,
//   so a fine tile must sit at least two trips after its coarse tile; the order is skewed by one pair:
//        C_0 | C_1 F_0 F_1 | C_2 F_2 F_3 | ... | C_P-1 F_2P-4 F_2P-3 | F_2P-2 F_2P-1          (3 P trips)
//   With a single pair (P = 1: C_0 F_0 F_1) the distance is one trip: F_0 `depends on its predecessor` -- the epilogue
//   finishes C_0 at once instead of behind the next tile's first layer and the front-end composites C_0 before it
//   encodes F_0 (one pipeline bubble; only launches of <= 2 rays per CTA take it).
// ---------------------------------------------------------------------------
struct Trip { long long tile; int pass; int S; };

// (host + device: nsr_debug_frame_schedule exposes the order to the CPU tests)
__host__ __device__ __forceinline__ Trip frame_trip(int unit_rays, long long it, long long P, int first, int stride) {
  int fine = 0; long long idx = 0;                 // coarse: CTA-local pair index; fine: CTA-local ray index
  if (it > 0) {
    const long long t = it - 1, q = t / 3;
    const int r = (int)(t - 3 * q);
    if (q == P - 1) { fine = 1; idx = 2 * q + r; }
    else if (r == 0) { idx = q + 1; }
    else { fine = 1; idx = 2 * q + r - 1; }
  }
  const long long pl = fine ? (idx >> 1) : idx;    // CTA-local pair
  const int hp = unit_rays >> 1;                   // pairs per unit
  const long long k = pl / hp;
  const long long ray0 = ((long long)first + k * stride) * unit_rays + 2 * (pl - k * hp);
  if (fine) return Trip{ray0 + (idx & 1), 1, 128};
  return Trip{ray0 >> 1, 0, 64};
}
template <bool FUSED>
__device__ __forceinline__ Trip trip_of(const TcKernelArgs& a, long long it, long long P, int first, int stride) {
  if (!FUSED) return Trip{first + it * (long long)stride, 0, a.S};
  return frame_trip(a.unit_rays, it, P, first, stride);
}
template <bool FUSED>
__host__ __device__ __forceinline__ bool trip_depends_on_previous(long long it, long long P) { return FUSED && P == 1 && it == 1; }

// ---------------------------------------------------------------------------
// roles
// ---------------------------------------------------------------------------
// (executed by the whole warp, warp-uniformly; one elected lane issues)
template <bool FUSED>
__device__ __forceinline__ void producer_role(const TcKernelArgs& a, uint32_t sm_base, long long my_tiles, long long P,
                                              uint32_t crank, uint32_t csize) {
  if (!elect_one()) return;       // one elected lane runs the whole loop (uniform registers, see mma_role)
  // Cluster of 2 (CTA pair): each CTA fetches HALF of every stage (rank 0 the hi plane, rank 1 the lo plane) and the copy is
  // multicast into both CTAs' rings -- the weight stream L2 -> shared memory is halved.  (Measured, tools/l2_weight_ab.py:
  // that stream costs no cycles but ~7 % of SM clock under the 1 kW cap.)  Each CTA still expects the full 32 KB on its own
  // stage barrier; the slot is reused only after BOTH CTAs' MMAs released it (B_WEMPTY counts csize arrivals).
  const uint16_t mask = (uint16_t)((1u << csize) - 1u);
  const uint32_t part = kStageBytes / csize, poff = crank * part;
  uint32_t slot = 0, par = 0;
#pragma unroll 1
  for (long long it = 0; it < my_tiles; ++it) {
    // fused frame: the tile's net alternates (the schedule depends on P only, so a CTA pair streams the same image)
    const uint8_t* image = (FUSED && trip_of<FUSED>(a, it, P, 0, 0).pass) ? a.image_f : a.image;
#pragma unroll 1
    for (int s = 0; s < kStagesPerTile; ++s) {
      mbar_wait(sm_base + kSmBar + 8 * (B_WEMPTY + slot), par ^ 1);
      const uint32_t full = sm_base + kSmBar + 8 * (B_WFULL + slot);
      if ((a.debug_flags & 1) && (it > 0 || s >= kRing)) { mbar_arrive(full); }
      else {
        mbar_expect_tx(full, kStageBytes);
        const uint32_t dst = sm_base + kSmRing + slot * kStageBytes;
        const uint8_t* src = image + (size_t)s * kStageBytes;
        if (csize == 1) bulk_copy_g2s(dst, src, kStageBytes, full);
        else bulk_copy_g2s_mc(dst + poff, src + poff, part, full, mask);
      }
      if (++slot == kRing) { slot = 0; par ^= 1; }
    }
  }
}

// ---------------------------------------------------------------------------
// MMA issue role.
//
// Measured facts that shape this code (tools/microbench/mma_rate.cu, stage_loop.cu, tools/tc_trace.py):
//  * tcgen05.mma M=128 N=128 K=16 runs at exactly its 64-cycle floor, TS or SS, but the pipe buffers
//    only ~1-2 instructions: every cycle the issuing lane spends between two MMAs beyond that is lost.
//  * mbarrier try_wait (already complete), tcgen05.commit and fences are cheap enough to hide; R2UR
//    chains, op decoding, local-memory state and I-cache misses are not.
// So the whole role runs on ONE elected lane (uniform registers), and every address, ring slot and
// barrier parity is a compile-time constant of the code position: ring slot and weight-barrier
// parity have period 8 in the stage index (4 slots x 2 phases), a tile has 72 = 9 x 8 stages, and a
// tile has an even number (10) of layers, so A_READY parities are static per layer too.  Layers
// L2-L4 and L6-L9 share one body each (stage-index phase 2 and 4 mod 8).
// The stage ORDER is build_stage_table()'s (the weight image is laid out in issue order).
// ---------------------------------------------------------------------------
// (MmaCtx, mma_stage_ts / mma_stage_ss, mma_layer: nsr_tc_mma.cuh -- shared with the backward dX chain, nsr_train.cu)

template <int PASSES>
__device__ __forceinline__ void mma_role(const TcKernelArgs& a, uint8_t* sm, uint32_t sm_base, uint32_t tmem,
                                         uint32_t idesc, long long my_tiles, uint32_t csize) {
  (void)sm; (void)tmem;            // TMEM base is 0 (checked at kernel start)
  if (!elect_one()) return;
  const MmaCtx c{sm_base + kSmRing, sm_base + kSmBar, idesc, csize > 1 ? (1u << csize) - 1u : 0u};
  if (my_tiles > 0) { mbar_wait(c.bar + 8 * (B_WFULL + 0), 0); tc_fence_after(); }   // first stage's weights
#pragma unroll 1
  for (long long it = 0; it < my_tiles; ++it) {
    const uint32_t buf = (uint32_t)(it & 1);
    const uint32_t enc = sm_base + kSmEnc + buf * kStageBytes;
    TR_DECL(a, it);
    TR(1000);
    mbar_wait(c.bar + 8 * (B_ENCFULL + buf), (uint32_t)((it >> 1) & 1));
    tc_fence_after();
    TR(1001);
    // ---- L1 (stages 0,1): A = encoding.  Half 1 first and WITHOUT a wait: L10 (N = 128) never
    // touches accumulator half 1 and quarters 2,3 of L9 were waited for by L10's chunks 2,3, so this
    // stage overlaps the previous tile's last epilogue.  (An early A_READY arrival instead would let
    // that barrier run two phases ahead of this lane -- parity aliasing.)
    mma_stage_ss<PASSES, 0>(c, enc, 1, true);
    tc_commit(c.bar + 8 * (B_ACCFULL + 1));
    if (it > 0) { mma_wait_ready(c, 0, 1u); mma_wait_ready(c, 1, 1u); mma_wait_ready(c, 2, 1u); mma_wait_ready(c, 3, 1u); tc_fence_after(); }
    mma_stage_ss<PASSES, 1>(c, enc, 0, true);
    tc_commit(c.bar + 8 * (B_ACCFULL + 0));
    tc_commit(c.bar + 8 * (B_AFREE + 0));
    tc_commit(c.bar + 8 * (B_AFREE + 1));
    TR(1101);
    // ---- L2..L4 (stages 2..25), L5 = skip layer (26..35), L6..L9 (36..67)
    // A_READY parity of the previous layer: g-1 = 10*it + L-2  ->  L & 1
#pragma unroll 1
    for (int L = 2; L <= 4; ++L) { mma_layer<PASSES, 2, false>(c, enc, (uint32_t)(L & 1)); TR(1100 + L); }
    mma_layer<PASSES, 2, true>(c, enc, 1u);
    TR(1105);
#pragma unroll 1
    for (int L = 6; L <= 9; ++L) { mma_layer<PASSES, 4, false>(c, enc, (uint32_t)(L & 1)); TR(1100 + L); }
    // ---- L10 (stages 68..71): dir_encoding feat part, N = 128: half 0 only
    mma_wait_ready(c, 0, 0u); mma_wait_ready(c, 1, 0u); tc_fence_after();
    mma_stage_ts<PASSES, 4>(c, 0, 0, true, true);
    mma_stage_ts<PASSES, 5>(c, 0, 1, false, true);
    mma_wait_ready(c, 2, 0u); tc_fence_after();
    mma_stage_ts<PASSES, 6>(c, 0, 2, false, true);
    mma_wait_ready(c, 3, 0u); tc_fence_after();
    mma_stage_ts<PASSES, 7>(c, 0, 3, false, it + 1 < my_tiles);     // no next stage after the last tile
    tc_commit(c.bar + 8 * (B_ACCFULL + 0));
    tc_commit(c.bar + 8 * (B_AFREE + 0));
    tc_commit(c.bar + 8 * (B_AFREE + 1));
    tc_commit(c.bar + 8 * (B_ACCFULL + 1));
    TR(1110);
  }
}

// The kernel hosts four concurrently running roles; its hot code has to stay near the 32 KB
// instruction cache (measured: a fully inlined build was 264 KB and every role stalled on
// instruction fetch while the front-end ran).  Hence: one out-of-line sincos, rolled loops.
__device__ __noinline__ void sincos_shared(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __noinline__ void generate_ray_shared(const RayGenParams& g, long long idx, float* out) { generate_ray(g, idx, out); }

// ---- front-end: sample, cast, encode, split, swizzled store; per-ray dir bias ----
template <int FMT, bool STASH, bool FUSED>
__device__ __forceinline__ void frontend_role(const TcKernelArgs& a, uint8_t* sm, uint32_t sm_base, long long my_tiles,
                                              long long P, int first_tile, int tile_stride) {
  const int t = threadIdx.x - 32 * kFrontWarp0;   // 0..127 = tile row
  const int lane = threadIdx.x & 31;
  const RenderParams& rp = a.rp;
  const int fw = t >> 5;   // front-end warp index
  // Compositing (+ resampling) of tile `j`, staged in shared memory by the epilogue warps.
  // It runs here, on the low-priority front-end warps that otherwise idle between encodes, so the
  // epilogue warps go straight on to the next tile's first layer.
  auto composite_tile = [&](long long j) {
    const uint32_t cbuf = (uint32_t)(j & 1);
    const Trip tr = trip_of<FUSED>(a, j, P, first_tile, tile_stride);
    const int S = tr.S, RPT = a.loose ? 2 : kTile / S;        // rays a tile can touch
    const bool fine = FUSED && tr.pass;
    float* const o_rgb = fine ? a.comp_rgb_f : a.comp_rgb;
    float* const o_depth = fine ? a.depth_f : a.depth;
    float* const o_opacity = fine ? a.opacity_f : a.opacity;
    float* const o_weights = fine ? a.weights_f : a.weights;
    float* const lr_acc = reinterpret_cast<float*>(sm + kSmLrAcc) + (fine ? 64 : 0);     // [16 rays][4]
    mbar_wait(sm_base + kSmBar + 8 * B_COMPREADY, (uint32_t)(j & 1));
    if (fw < RPT && !a.loose) {
      const long long tile_j = tr.tile;
      const long long ray = tile_j * RPT + fw;
      if (ray < a.n_rays) {
        const float* zt = reinterpret_cast<const float*>(sm + kSmZ) + cbuf * 128 + fw * S;
        const float* ssig = reinterpret_cast<const float*>(sm + kSmSig) + fw * S;
        const float* srgb = reinterpret_cast<const float*>(sm + kSmRgb) + 3 * fw * S;
        float* sw = reinterpret_cast<float*>(sm + kSmW) + fw * S;
        float* stmp = reinterpret_cast<float*>(sm + kSmTmp) + fw * S;
        float r, gg, b, d, o;
        composite_ray_warp(zt, ssig, srgb, S, rp.white_bkgd, rp.sigma_softplus, sw, stmp, r, gg, b, d, o);
        if (lane == 0) {
          if (o_rgb) { o_rgb[ray * 3] = r; o_rgb[ray * 3 + 1] = gg; o_rgb[ray * 3 + 2] = b; }
          if (o_depth) o_depth[ray] = d;
          if (o_opacity) o_opacity[ray] = o;
          if (FUSED) {       // staged for the box average below
            float* acc = lr_acc + 4 * (int)(ray % a.unit_rays);
            acc[0] = r; acc[1] = gg; acc[2] = b; acc[3] = d;
          }
        }
        if (o_weights) for (int i = lane; i < S; i += 32) o_weights[ray * S + i] = sw[i];
        if (FUSED ? !fine : a.do_resample) {
          const int n_imp = rp.n_importance;
          float* scratch = reinterpret_cast<float*>(sm + kSmScratch) + fw * 320;
          resample_ray_warp(zt, sw, S, n_imp, a.u_resample ? a.u_resample + ray * n_imp : nullptr,
                            a.tabs->u_fine, scratch, a.z_next + ray * (S + n_imp));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm_base + kSmBar + 8 * B_COMPDONE);
    // all front-end warps wait here: the next encode overwrites the z / dir-bias buffers that the
    // compositing warps above are still reading
    named_bar_sync(1, 32 * kWarpsFront);
    if (FUSED && a.ss > 1 && t < 4) {
      // Box average (nerf_downX_model.py:337-348): a tile ends an LR pixel when its last ray is the pixel's last sub-pixel
      // ray (U and ss are even, tiles hold 1 or 2 whole rays).  Sum in ray order, then divide -- k_box_average's arithmetic.
      const long long last = tr.tile * RPT + RPT - 1;
      float* const lr_rgb = fine ? a.lr_rgb_f : a.lr_rgb_c;
      float* const lr_depth = fine ? a.lr_depth_f : a.lr_depth_c;
      if ((last + 1) % a.ss == 0 && last < a.n_rays) {
        const int r0 = (int)((last + 1 - a.ss) % a.unit_rays);
        float sum = 0.f;
        for (int k = 0; k < a.ss; ++k) sum = __fadd_rn(sum, lr_acc[4 * (r0 + k) + t]);
        const float mean = __fdiv_rn(sum, (float)a.ss);
        const long long px = last / a.ss;
        if (t < 3) { if (lr_rgb) lr_rgb[px * 3 + t] = mean; }
        else if (lr_depth) lr_depth[px] = mean;
      }
    }
  };
  auto encode_tile = [&](long long it) {
    const uint32_t buf = (uint32_t)(it & 1);
    // buffers `buf` were last used by tile it-2, whose compositing this role finished in the
    // previous iteration (program order + the named barrier closing composite_tile).
    TR_DECL(a, it - 1);     // front-end works one tile ahead: trace the work done for iteration 3 during tile 2..3
    TR(5000);
    const Trip tr = trip_of<FUSED>(a, it, P, first_tile, tile_stride);
    const long long tile = tr.tile;
    const int S = tr.S, RPT = a.loose ? 2 : kTile / S;
    const bool fine = FUSED && tr.pass;
    const float* const consts = fine ? a.consts_f : a.consts;
    const long long p0 = tile * kTile;                       // first point of the tile; its first ray is p0 / S
    const long long ray = (p0 + t) / S;
    const int i = (int)((p0 + t) % S);
    const bool valid = ray < a.n_rays;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, near = 0, far = 0;
    if (valid) {
      if (FUSED && a.use_rg) {       // pose -> ray in the front-end (nsr_render_pose_host): no ray buffer in HBM
        float r8[8];
        generate_ray_shared(a.rg, a.rg_first + ray, r8);
        ox = r8[0]; oy = r8[1]; oz = r8[2]; dx = r8[3]; dy = r8[4]; dz = r8[5]; near = r8[6]; far = r8[7];
      } else {
        const float* rr = a.rays + ray * a.ray_stride;
        ox = rr[0]; oy = rr[1]; oz = rr[2]; dx = rr[3]; dy = rr[4]; dz = rr[5]; near = rr[6]; far = rr[7];
      }
    }
    float zv = 0.f;
    if (valid) {
      if (FUSED ? fine : (a.z_in != nullptr)) {
        // fused frame: the merged z-values this CTA's own compositing warp wrote >= 1 named barrier ago (L2 load: the
        // writer may have been another warp of this SM, so no L1 line may be trusted)
        zv = FUSED ? __ldcg(a.z_next + ray * S + i) : a.z_in[ray * S + i];
      } else {
        const float zc = coarse_z(near, far, a.tabs->t_coarse[i], a.tabs->one_minus_t[i], rp.lindisp);
        zv = zc;
        if (a.u_jitter) {
          const float zp = i > 0 ? coarse_z(near, far, a.tabs->t_coarse[i - 1], a.tabs->one_minus_t[i - 1], rp.lindisp) : zc;
          const float zn = i + 1 < S ? coarse_z(near, far, a.tabs->t_coarse[i + 1], a.tabs->one_minus_t[i + 1], rp.lindisp) : zc;
          zv = jitter_z(zp, zc, zn, i == 0, i + 1 == S, a.u_jitter[ray * S + i]);
        }
      }
    }
    reinterpret_cast<float*>(sm + kSmZ)[buf * 128 + t] = zv;
    if (STASH && a.z_out && valid) a.z_out[ray * S + i] = zv;
    const float px = cast_point(ox, dx, zv), py = cast_point(oy, dy, zv), pz = cast_point(oz, dz, zv);
    // 63 encoded channels (+1 zero pad), reference channel order (embedding.py:57-63)
    float e[64];
    e[0] = px; e[1] = py; e[2] = pz;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float f = a.tabs->freq_pos[k];
      const float ax = __fmul_rn(f, px), ay = __fmul_rn(f, py), az = __fmul_rn(f, pz);
      float sn, cs;
      sincos_shared(ax, &sn, &cs); e[3 + 6 * k + 0] = sn; e[3 + 6 * k + 3] = cs;
      sincos_shared(ay, &sn, &cs); e[3 + 6 * k + 1] = sn; e[3 + 6 * k + 4] = cs;
      sincos_shared(az, &sn, &cs); e[3 + 6 * k + 2] = sn; e[3 + 6 * k + 5] = cs;
    }
    e[63] = 0.f;
    TR(5001);
    uint8_t* row_hi = sm + kSmEnc + buf * kStageBytes + t * 128;
    uint8_t* row_lo = row_hi + kPlaneBytes;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) Split<FMT>::apply(e[8 * j + 2 * q], e[8 * j + 2 * q + 1], hi[q], lo[q]);
      const int sw = (j ^ (t & 7)) << 4;
      *reinterpret_cast<uint4*>(row_hi + sw) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(row_lo + sw) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      if (STASH && tile < a.n_tiles) {   // the same 16-byte pieces -> the tile image in HBM (lanes = consecutive points: 512 B per warp store)
        uint8_t* g = a.stash_enc + (size_t)tile * kStageBytes + img2_off(t, j);
        stg128(g, hi[0], hi[1], hi[2], hi[3]);
        stg128(g + kPlaneBytes, lo[0], lo[1], lo[2], lo[3]);
      }
    }
    TR(5002);
    // view-direction encoding of the tile's rays -> smem, then the per-ray bias of the dir layer:
    // dirbias[r][j] = b_dir[j] + sum_c Wdir[j][256+c] * enc_dir[r][c]   (networks.py:214-221)
    float* denc = reinterpret_cast<float*>(sm + kSmDenc);
    if (t < RPT) {
      const long long r2 = p0 / S + t;                       // slot t of the tile's rays
      float vx = 0, vy = 0, vz = 0;
      if (r2 < a.n_rays) {
        if (FUSED && a.use_rg) {     // 8-column rays: the view direction is the (normalised) ray direction
          float r8[8];
          generate_ray_shared(a.rg, a.rg_first + r2, r8);
          vx = r8[3]; vy = r8[4]; vz = r8[5];
        } else {
          const float* rr = a.rays + r2 * a.ray_stride + rp.viewdir_offset;
          vx = rr[0]; vy = rr[1]; vz = rr[2];
        }
      }
      float* o = denc + t * 32;
      o[0] = vx; o[1] = vy; o[2] = vz;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const float f = a.tabs->freq_dir[k];
        float sn, cs;
        sincos_shared(__fmul_rn(f, vx), &sn, &cs); o[3 + 6 * k + 0] = sn; o[3 + 6 * k + 3] = cs;
        sincos_shared(__fmul_rn(f, vy), &sn, &cs); o[3 + 6 * k + 1] = sn; o[3 + 6 * k + 4] = cs;
        sincos_shared(__fmul_rn(f, vz), &sn, &cs); o[3 + 6 * k + 2] = sn; o[3 + 6 * k + 5] = cs;
      }
    }
    named_bar_sync(1, 32 * kWarpsFront);
    float* dbias = reinterpret_cast<float*>(sm + kSmDirBias) + buf * 256;
    for (int idx = t; idx < RPT * 128; idx += 128) {
      const int r = idx >> 7, j = idx & 127;
      float s = consts[kcBiasDir + j];
      if (!rp.no_dir) {
        const float* w = consts + kcWdd + j * 28;
        const float* de = denc + r * 32;
#pragma unroll
        for (int c = 0; c < 27; ++c) s = fmaf(__ldg(w + c), de[c], s);
      }
      dbias[idx] = s;
    }
    TR(5003);
    fence_proxy_async();     // make the generic-proxy enc writes visible to the tensor core's async proxy
    __syncwarp();
    if (lane == 0) mbar_arrive(sm_base + kSmBar + 8 * (B_ENCFULL + buf));
    named_bar_sync(1, 32 * kWarpsFront);   // denc is reused next tile
  };
  // encode(it) runs one tile ahead of the MLP; composite(it-1) follows it (single call sites keep
  // the instruction footprint small); one extra trip composites the last tile.
  // Order: e0 e1 c0 e2 c1 ... (encode at most one tile ahead of the last composited one); a tile that depends on its
  // predecessor (fused frame, P = 1) is encoded only after the predecessor's compositing.
  long long e = 0, c = 0;
#pragma unroll 1
  while (c < my_tiles) {
    const bool enc = e < my_tiles && (e == c || (e == c + 1 && !trip_depends_on_previous<FUSED>(e, P)));
    if (enc) encode_tile(e++);
    else composite_tile(c++);
  }
}

// One trunk layer's epilogue for this warp: the four 64-column accumulator quarters in order,
// 32 columns of each:  acc + bias (+ReLU) (-> sigma-head partial) -> hi/lo split -> A operand planes.
template <int FMT, int PASSES, bool SIGMA, bool STASH>
__device__ __forceinline__ void epi_layer(int L, float relu_floor, uint32_t g, uint32_t bar, uint32_t tlane,
                                          uint32_t cst_addr, int hh, int lane, float& sig_p, int row,
                                          uint8_t* stash_chunks, uint32_t* mask_row TR_PARAMS) {
  const uint32_t bias_addr = cst_addr + 4u * (uint32_t)((L - 1) * 256);   // L9 -> kcBiasFinal
#pragma unroll 1
  for (int q4 = 0; q4 < 4; ++q4) {
    TR(100 * L + 10 * q4 + 0);
    if (q4 == 0) { mbar_wait(bar + 8 * (B_ACCFULL + 0), g & 1); mbar_wait(bar + 8 * (B_AFREE + 0), g & 1); }
    else if (q4 == 1) mbar_wait(bar + 8 * (B_AFREE + 1), g & 1);
    else if (q4 == 2) mbar_wait(bar + 8 * (B_ACCFULL + 1), g & 1);
    tc_fence_after();
    TR(100 * L + 10 * q4 + 1);
    const int col0 = 64 * q4 + 32 * hh;
    uint32_t r[32];
    TMEM_LD32(tlane + (uint32_t)col0, r);
    tc_wait_ld();
    TR(100 * L + 10 * q4 + 2);
    uint32_t whi[16], wlo[16];
    uint32_t relu_bits = 0u;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = lds128(bias_addr + 4u * (uint32_t)(col0 + j));
      float v0 = __uint_as_float(r[j]) + b4.x, v1 = __uint_as_float(r[j + 1]) + b4.y;
      float v2 = __uint_as_float(r[j + 2]) + b4.z, v3 = __uint_as_float(r[j + 3]) + b4.w;
      v0 = fmaxf(v0, relu_floor); v1 = fmaxf(v1, relu_floor); v2 = fmaxf(v2, relu_floor); v3 = fmaxf(v3, relu_floor);
      if (SIGMA) {   // sigma head on h_8 (networks.py:207)
        const float4 w4 = lds128(cst_addr + 4u * (uint32_t)(kcWsig + col0 + j));
        sig_p = fmaf(v0, w4.x, sig_p); sig_p = fmaf(v1, w4.y, sig_p);
        sig_p = fmaf(v2, w4.z, sig_p); sig_p = fmaf(v3, w4.w, sig_p);
      }
      Split<FMT>::apply(v0, v1, whi[j / 2], wlo[j / 2]);
      Split<FMT>::apply(v2, v3, whi[j / 2 + 1], wlo[j / 2 + 1]);
      if (STASH) relu_bits |= ((v0 > 0.f ? 1u : 0u) | (v1 > 0.f ? 2u : 0u) | (v2 > 0.f ? 4u : 0u) | (v3 > 0.f ? 8u : 0u)) << j;
    }
    TMEM_ST16(tlane + 256u + (uint32_t)(col0 / 2), whi);
    if (PASSES == 3) TMEM_ST16(tlane + 384u + (uint32_t)(col0 / 2), wlo);
    TR(100 * L + 10 * q4 + 3);
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar + 8 * (B_AREADY + q4));
    if (STASH && stash_chunks) {   // this thread's 32 activations of layer L -> the layer's tile image, AFTER the arrive (the MMA lane is
                   // already released); four 16-byte pieces per plane, each a 512-byte contiguous warp store
      uint8_t* gp = stash_chunks + (size_t)q4 * kStageBytes + img2_off(row, 4 * hh);      // k chunk = 64-column quarter
#pragma unroll
      for (int t2 = 0; t2 < 4; ++t2) {
        stg128(gp + 1024 * t2, whi[4 * t2], whi[4 * t2 + 1], whi[4 * t2 + 2], whi[4 * t2 + 3]);
        stg128(gp + kPlaneBytes + 1024 * t2, wlo[4 * t2], wlo[4 * t2 + 1], wlo[4 * t2 + 2], wlo[4 * t2 + 3]);
      }
      if (mask_row) mask_row[2 * q4 + hh] = relu_bits;          // columns [64 q4 + 32 hh, +32)
    }
    TR(100 * L + 10 * q4 + 4);
  }
}

// ---- epilogue + compositing ----
template <int FMT, int PASSES, bool STASH, bool FUSED>
__device__ __forceinline__ void epilogue_role(const TcKernelArgs& a, uint8_t* sm, uint32_t sm_base, uint32_t tmem,
                                              long long my_tiles, long long P, int first_tile, int tile_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp - kEpiWarp0;          // 0..7
  const int q = warp & 3;                   // TMEM lane quarter this warp may access
  const int hh = ew >> 2;                   // which 64-column half of an accumulator half
  const int row = 32 * q + lane;            // tile row == TMEM lane
  const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t bar = sm_base + kSmBar;
  const RenderParams& rp = a.rp;
  uint32_t g = 0;
  float* xch = reinterpret_cast<float*>(sm + kSmXch);
  float* ssig = reinterpret_cast<float*>(sm + kSmSig);
  float* srgb = reinterpret_cast<float*>(sm + kSmRgb);
  // Tail of tile j: combine the two column halves of the head partial sums, activations, stage the
  // per-point (rgb, sigma) for the compositing done by the front-end warps.  It is NOT on the MMA
  // critical path, so it runs after the NEXT tile's first-layer epilogue (see the loop below).
  auto tile_tail = [&](long long j, float sig_p, float r0, float r1, float r2) {
    const Trip tr = trip_of<FUSED>(a, j, P, first_tile, tile_stride);
    const long long tile = tr.tile;
    const int S = tr.S;
    const float* cst = reinterpret_cast<const float*>(sm + kSmConst + ((FUSED && tr.pass) ? kConstBytes : 0));
    const float* noise = (FUSED && tr.pass) ? a.noise_f : a.noise;
    if (hh == 1) { xch[row] = sig_p; xch[128 + row] = r0; xch[256 + row] = r1; xch[384 + row] = r2; }
    named_bar_sync(2, 32 * kWarpsEpi);
    if (hh == 0) {
      const long long ray = (tile * kTile + row) / S;
      const bool valid = ray < a.n_rays;
      const float sigma = (sig_p + xch[row]) + cst[kcMisc];
      float col[3] = {r0 + xch[128 + row], r1 + xch[256 + row], r2 + xch[384 + row]};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = col[c] + cst[kcMisc + 1 + c];
        if (!rp.color_none) v = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));    // torch.sigmoid
        col[c] = v;
      }
      const long long gp = tile * kTile + row;
      if (a.raw && valid) reinterpret_cast<float4*>(a.raw)[gp] = make_float4(col[0], col[1], col[2], sigma);
      if (rp.gamma_correct) {
#pragma unroll
        for (int c = 0; c < 3; ++c) col[c] = powf(col[c], 1.f / 2.2f);        // nerf_downX_model.py:271
      }
      float sg = sigma;
      if (noise && valid) sg = __fadd_rn(sg, __fmul_rn(noise[gp], rp.noise_std));   // utils.py:210
      // the staging buffers are single: wait until the previous tile has been composited
      if (j >= 1) mbar_wait(bar + 8 * B_COMPDONE, (uint32_t)((j - 1) & 1));
      ssig[row] = sg; srgb[3 * row] = col[0]; srgb[3 * row + 1] = col[1]; srgb[3 * row + 2] = col[2];
    }
    named_bar_sync(2, 32 * kWarpsEpi);      // xch reads done before the next tail's writes
    if (lane == 0) mbar_arrive(bar + 8 * B_COMPREADY);
  };
  float sig_prev = 0.f, rgb_prev[3] = {0.f, 0.f, 0.f};
  bool tail_pending = false;              // the previous tile's tail has not run yet
#pragma unroll 1
  for (long long it = 0; it < my_tiles; ++it) {
    const uint32_t buf = (uint32_t)(it & 1);
    float sig_p = 0.f;
    const Trip tr = trip_of<FUSED>(a, it, P, first_tile, tile_stride);
    const int S = tr.S;
    const uint32_t cst_addr = sm_base + kSmConst + ((FUSED && tr.pass) ? kConstBytes : 0);
    TR_DECL(a, it);
    // ---- layers 1..9: bias (+ReLU) -> hi/lo split -> next A operand ----
    // (two code variants only -- the I-cache is shared by four roles: the ReLU floor is a runtime
    //  value: 0 for the trunk, -inf for the activation-free xyz_encoding_final, networks.py:158)
#pragma unroll 1
    for (int L = 1; L <= 9; ++L, ++g) {
      // the previous tile's tail goes here, behind this tile's first layer: the MMA lane is already
      // busy with L2 while the heads' activations are finished and staged
      if (L == 2 && tail_pending) { tile_tail(it - 1, sig_prev, rgb_prev[0], rgb_prev[1], rgb_prev[2]); tail_pending = false; }
      uint8_t* stash_chunks = nullptr;
      uint32_t* mask_row = nullptr;
      if (STASH && tr.tile < a.n_tiles) {      // (a cluster's dummy tile stores nothing)
        const size_t lt = (size_t)(L - 1) * (size_t)a.n_tiles + (size_t)tr.tile;
        stash_chunks = a.stash_h + lt * (size_t)(4 * kStageBytes);
        if (L <= 8) mask_row = a.stash_mask + (lt * 128 + (size_t)row) * 8;
      }
      if (L == 8) epi_layer<FMT, PASSES, true, STASH>(L, 0.f, g, bar, tlane, cst_addr, hh, lane, sig_p, row, stash_chunks, mask_row TR_ARGS);   // + sigma head
      else epi_layer<FMT, PASSES, false, STASH>(L, L <= 8 ? 0.f : -INFINITY, g, bar, tlane, cst_addr, hh, lane, sig_p, row, stash_chunks, mask_row TR_ARGS);
    }
    // ---- layer 10: dir layer (N=128, accumulator half 0) + rgb head ----
    float rgb_p[3] = {0.f, 0.f, 0.f};
    {
      // Only ACC_FULL[0] carries information here.  ACC_FULL[1] / A_FREE[] of L10 are committed by
      // the MMA lane purely to keep every barrier at one phase per layer (static parities) and must
      // NOT be waited for: the next tile's un-gated L1 half-1 stage commits ACC_FULL[1] again right
      // away, so a waiter here could fall two phases behind (parity aliasing -> deadlock).
      mbar_wait(bar + 8 * (B_ACCFULL + 0), g & 1);
      tc_fence_after();
      const long long p0 = tr.tile * kTile;
      const int slot = (int)((p0 + row) / S - p0 / S);       // which of the tile's rays this row belongs to (0 or 1)
      const uint32_t dbias_addr = sm_base + kSmDirBias + 4u * (uint32_t)(buf * 256 + slot * 128);
      const uint32_t wrgb_addr = cst_addr + 4u * kcWrgb;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = 64 * hh + 32 * c;
        uint32_t r[32];
        uint32_t dhi[16], dlo[16];
        TMEM_LD32(tlane + (uint32_t)col0, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = lds128(dbias_addr + 4u * (uint32_t)(col0 + j));
          const float4 w0 = lds128(wrgb_addr + 4u * (uint32_t)(col0 + j));
          const float4 w1 = lds128(wrgb_addr + 4u * (uint32_t)(128 + col0 + j));
          const float4 w2 = lds128(wrgb_addr + 4u * (uint32_t)(256 + col0 + j));
          const float v0 = fmaxf(__uint_as_float(r[j]) + b4.x, 0.f), v1 = fmaxf(__uint_as_float(r[j + 1]) + b4.y, 0.f);
          const float v2 = fmaxf(__uint_as_float(r[j + 2]) + b4.z, 0.f), v3 = fmaxf(__uint_as_float(r[j + 3]) + b4.w, 0.f);
          rgb_p[0] = fmaf(v0, w0.x, rgb_p[0]); rgb_p[0] = fmaf(v1, w0.y, rgb_p[0]); rgb_p[0] = fmaf(v2, w0.z, rgb_p[0]); rgb_p[0] = fmaf(v3, w0.w, rgb_p[0]);
          rgb_p[1] = fmaf(v0, w1.x, rgb_p[1]); rgb_p[1] = fmaf(v1, w1.y, rgb_p[1]); rgb_p[1] = fmaf(v2, w1.z, rgb_p[1]); rgb_p[1] = fmaf(v3, w1.w, rgb_p[1]);
          rgb_p[2] = fmaf(v0, w2.x, rgb_p[2]); rgb_p[2] = fmaf(v1, w2.y, rgb_p[2]); rgb_p[2] = fmaf(v2, w2.z, rgb_p[2]); rgb_p[2] = fmaf(v3, w2.w, rgb_p[2]);
          if (STASH) {
            Split<FMT>::apply(v0, v1, dhi[j / 2], dlo[j / 2]);
            Split<FMT>::apply(v2, v3, dhi[j / 2 + 1], dlo[j / 2 + 1]);
          }
        }
        if (STASH && tr.tile < a.n_tiles) {   // dir layer activations (post-ReLU) -> tile image, chunk hh, pieces j = 4c .. 4c+3
          uint8_t* gp = a.stash_dir + ((size_t)tr.tile * 2 + (size_t)hh) * (size_t)kStageBytes + img2_off(row, 4 * c);
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            stg128(gp + 1024 * t2, dhi[4 * t2], dhi[4 * t2 + 1], dhi[4 * t2 + 2], dhi[4 * t2 + 3]);
            stg128(gp + kPlaneBytes + 1024 * t2, dlo[4 * t2], dlo[4 * t2 + 1], dlo[4 * t2 + 2], dlo[4 * t2 + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) mbar_arrive(bar + 8 * (B_AREADY + q4));
      }
      ++g;
    }
    sig_prev = sig_p; rgb_prev[0] = rgb_p[0]; rgb_prev[1] = rgb_p[1]; rgb_prev[2] = rgb_p[2];
    // The tail normally runs behind the NEXT tile's first layer; the last tile's, and that of a tile the next one depends
    // on (fused frame, P = 1), runs at once.
    tail_pending = true;
    if (it + 1 == my_tiles || trip_depends_on_previous<FUSED>(it + 1, P)) {
      tile_tail(it, sig_prev, rgb_prev[0], rgb_prev[1], rgb_prev[2]);
      tail_pending = false;
    }
  }
}
template <int FMT, int PASSES, bool STASH, bool FUSED = false>
__global__ void __launch_bounds__(kThreadsTc, 1) k_tc_pass(const __grid_constant__ TcKernelArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B operands need 1024-B alignment
  uint8_t* sm = smem_raw;                                 // (keeps the __shared__ address space: LDS/STS)
  const uint32_t sm_base = smem_u32(sm);
  if (sm_base & 1023u) { if (threadIdx.x == 0) printf("[nsr_tc] dynamic smem base %u not 1024-aligned\n", sm_base); __trap(); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // In a cluster the CTAs share the weight ring stage by stage, so all of them run the SAME number of tiles; a tile index
  // past the end is a dummy (no valid ray: nothing is read or written for it).
  const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
  // (fused frame: the same rule over work units; P ray pairs = 3 P trips, see trip_of)
  const long long n_work = FUSED ? a.n_units : a.n_tiles;
  const long long my_work = (csize > 1) ? (n_work + gridDim.x - 1) / gridDim.x
                                        : ((n_work > blockIdx.x) ? (n_work - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const long long P = FUSED ? my_work * (a.unit_rays >> 1) : 0;
  const long long my_tiles = FUSED ? 3 * P : my_work;
  if (a.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    a.clk[0] = clock64(); a.clk[1] = (long long)ns;
  }

  if (threadIdx.x == 0) {
    const uint32_t bar = sm_base + kSmBar;
    for (int i = 0; i < kRing; ++i) { mbar_init(bar + 8 * (B_WFULL + i), 1); mbar_init(bar + 8 * (B_WEMPTY + i), csize); }
    mbar_init(bar + 8 * (B_ACCFULL + 0), 1); mbar_init(bar + 8 * (B_ACCFULL + 1), 1);
    mbar_init(bar + 8 * (B_AFREE + 0), 1); mbar_init(bar + 8 * (B_AFREE + 1), 1);
    for (int q4 = 0; q4 < 4; ++q4) mbar_init(bar + 8 * (B_AREADY + q4), kWarpsEpi);
    mbar_init(bar + 8 * (B_ENCFULL + 0), kWarpsFront); mbar_init(bar + 8 * (B_ENCFULL + 1), kWarpsFront);
    mbar_init(bar + 8 * B_COMPREADY, kWarpsEpi); mbar_init(bar + 8 * B_COMPDONE, kWarpsFront);
    fence_barrier_init();
  }
  // fp32 constants (biases, head weights) -> smem
  {
    float* cst = reinterpret_cast<float*>(sm + kSmConst);
    for (int i = threadIdx.x; i < kcSmemFloats; i += kThreadsTc) cst[i] = a.consts[i];
    if (FUSED) for (int i = threadIdx.x; i < kcSmemFloats; i += kThreadsTc) cst[kcSmemFloats + i] = a.consts_f[i];
  }
  if (warp == kMmaWarp) {   // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_base + kSmTmemPtr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (csize > 1) cluster_sync_all();     // the peer's barriers are initialised before anything is multicast into them
  // All 512 columns are allocated, so the base is lane 0 / column 0; using the literal keeps every
  // TMEM address warp-uniform for the compiler (checked, not assumed).
  if (*reinterpret_cast<volatile uint32_t*>(sm + kSmTmemPtr) != 0u) {
    if (threadIdx.x == 0) printf("[nsr_tc] unexpected TMEM base %u\n", *reinterpret_cast<volatile uint32_t*>(sm + kSmTmemPtr));
    __trap();
  }
  const uint32_t tmem = 0u;

  if (warp == kProducerWarp) {
    producer_role<FUSED>(a, sm_base, my_tiles, P, crank, csize);
    __syncwarp();
  } else if (warp == kMmaWarp) {
    mma_role<PASSES>(a, sm, sm_base, tmem, umma_idesc(FMT, 128, 128), my_tiles, csize);
    __syncwarp();
  } else if (warp < kEpiWarp0) {
    frontend_role<FMT, STASH, FUSED>(a, sm, sm_base, my_tiles, P, blockIdx.x, gridDim.x);
  } else {
    epilogue_role<FMT, PASSES, STASH, FUSED>(a, sm, sm_base, tmem, my_tiles, P, blockIdx.x, gridDim.x);
  }

  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();     // no CTA leaves while a peer's multicast copies / commits may still target it
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
  if (a.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    a.clk[2] = clock64(); a.clk[3] = (long long)ns;
  }
}

// The opt-in to > 48 KB of dynamic shared memory is a per-function, per-device attribute: set once at nsr_create
// (not on every launch).
cudaError_t tc_init(NsrHandle_* h) {
  cudaError_t e = cudaSuccess;
  auto set = [&](void (*k)(const TcKernelArgs)) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTcBytes);
  };
  switch (h->cfg.precision) {
    case NSR_PREC_BF16X3_TC: set(k_tc_pass<1, 3, true>); set(k_tc_pass<1, 3, false>); set(k_tc_pass<1, 3, false, true>); break;
    case NSR_PREC_FP16X3_TC: set(k_tc_pass<0, 3, true>); set(k_tc_pass<0, 3, false>); set(k_tc_pass<0, 3, false, true>); break;
    case NSR_PREC_BF16_TC: set(k_tc_pass<1, 1, false>); set(k_tc_pass<1, 1, false, true>); break;
    default: break;
  }
  return e;
}

cudaError_t tc_pass(NsrHandle_* h, int which, const TcPassArgs& p, cudaStream_t st) {
  const bool loose = (kTile % p.S) != 0 || p.S > kTile;
  if (p.S != 64 && p.S != 128 && p.S != 192 && p.S != 256) return cudaErrorInvalidValue;
  // MLP-only mode: caller-provided z-values in, raw (rgb, sigma) out; no stash, no in-kernel compositing / resampling
  if (loose && (!p.z_in || !p.raw || p.stash_enc || p.do_resample)) return cudaErrorInvalidValue;
  TcKernelArgs a{};
  a.image = h->net[which].tc_image; a.consts = h->net[which].tc_consts; a.tabs = h->d_tables; a.rp = h->rp;
  a.rays = p.rays; a.n_rays = p.n_rays; a.ray_stride = p.ray_stride; a.z_in = p.z_in; a.S = p.S;
  a.u_jitter = p.u_jitter; a.noise = p.noise; a.u_resample = p.u_resample; a.do_resample = p.do_resample;
  a.comp_rgb = p.comp_rgb; a.depth = p.depth; a.opacity = p.opacity; a.weights = p.weights; a.raw = p.raw;
  a.z_next = p.z_next;
  a.trace = p.trace;
  a.debug_flags = p.debug_flags;
  a.loose = loose ? 1 : 0;
  if (loose) a.n_tiles = (p.n_rays * p.S + kTile - 1) / kTile;
  else { const int rpt = kTile / p.S; a.n_tiles = (p.n_rays + rpt - 1) / rpt; }
  if (a.n_tiles == 0) return cudaSuccess;
  const int grid = (int)(a.n_tiles < h->sm_count ? a.n_tiles : h->sm_count);
  void (*kern)(const TcKernelArgs) = nullptr;
  const bool stash = p.stash_enc != nullptr;
  a.stash_enc = p.stash_enc; a.stash_h = p.stash_h; a.stash_dir = p.stash_dir; a.stash_mask = p.stash_mask; a.z_out = p.z_out;
  switch (h->cfg.precision) {
    case NSR_PREC_BF16X3_TC: kern = stash ? k_tc_pass<1, 3, true> : k_tc_pass<1, 3, false>; break;
    case NSR_PREC_FP16X3_TC: kern = stash ? k_tc_pass<0, 3, true> : k_tc_pass<0, 3, false>; break;
    case NSR_PREC_BF16_TC: if (stash) return cudaErrorInvalidValue; kern = k_tc_pass<1, 1, false>; break;
    default: return cudaErrorInvalidValue;
  }
  a.clk = h->d_clk;
  const int csize = (h->tc_cluster == 2 && a.n_tiles >= 2) ? 2 : 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(csize == 2 ? (grid + 1) & ~1 : grid));
  cfg.blockDim = dim3(kThreadsTc);
  cfg.dynamicSmemBytes = kSmemTcBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) return e;
  h->launches += 1;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// fused frame: coarse + fine pass (+ ray generation, + box average) of a ray batch in one launch
// ---------------------------------------------------------------------------
bool tc_frame_supported(const NsrHandle_* h, int s) {
  if (!h->tc_fused) return false;
  const int p = h->cfg.precision;
  if (p != NSR_PREC_BF16X3_TC && p != NSR_PREC_FP16X3_TC && p != NSR_PREC_BF16_TC) return false;
  if (h->cfg.n_coarse != 64 || h->cfg.n_importance != 64) return false;      // coarse tile = 2 rays, fine tile = 1 ray
  return s == 1 || s == 2 || s == 4;                                         // unit = lcm(2, s*s) <= 16 rays
}

// Debug / test seam: the tile order of one CTA of the frame kernel, computed on the host by the function the kernel uses.
// out[3 * trip + {0, 1, 2}] = pass (0 coarse, 1 fine), tile index within that pass, 1 if the trip depends on its predecessor.
int frame_schedule(long long pairs, int unit_rays, int first, int stride, long long* out) {
  if (pairs < 0 || (unit_rays != 2 && unit_rays != 4 && unit_rays != 16) || (pairs % (unit_rays / 2)) || !out) return 1;
  for (long long it = 0; it < 3 * pairs; ++it) {
    const Trip t = frame_trip(unit_rays, it, pairs, first, stride);
    out[3 * it] = t.pass; out[3 * it + 1] = t.tile; out[3 * it + 2] = trip_depends_on_previous<true>(it, pairs) ? 1 : 0;
  }
  return 0;
}

cudaError_t tc_frame(NsrHandle_* h, const TcFrameArgs& p, cudaStream_t st) {
  if (!tc_frame_supported(h, p.s) || !p.z_fine || (!p.rays && !p.rg)) return cudaErrorInvalidValue;
  const int ss = p.s * p.s;
  const bool lr = p.lr_rgb_c || p.lr_depth_c || p.lr_rgb_f || p.lr_depth_f;
  if (lr && (p.n_rays % ss)) return cudaErrorInvalidValue;
  if (p.n_rays == 0) return cudaSuccess;
  TcKernelArgs a{};
  a.image = h->net[0].tc_image; a.consts = h->net[0].tc_consts; a.image_f = h->net[1].tc_image; a.consts_f = h->net[1].tc_consts;
  a.tabs = h->d_tables; a.rp = h->rp;
  a.rays = p.rays; a.n_rays = p.n_rays; a.ray_stride = p.ray_stride; a.S = 64;
  if (!p.rays) { a.rg = *p.rg; a.use_rg = 1; a.rg_first = p.rg_first; }
  a.u_jitter = p.u_jitter; a.noise = p.noise_c; a.noise_f = p.noise_f; a.u_resample = p.u_resample; a.do_resample = 1;
  a.comp_rgb = p.c_rgb; a.depth = p.c_depth; a.opacity = p.c_opacity; a.weights = p.c_weights;
  a.comp_rgb_f = p.f_rgb; a.depth_f = p.f_depth; a.opacity_f = p.f_opacity; a.weights_f = p.f_weights;
  a.z_next = p.z_fine;
  a.lr_rgb_c = p.lr_rgb_c; a.lr_depth_c = p.lr_depth_c; a.lr_rgb_f = p.lr_rgb_f; a.lr_depth_f = p.lr_depth_f;
  a.ss = lr ? ss : 1;
  a.unit_rays = (a.ss == 1) ? 2 : a.ss;            // lcm(2, ss) for ss in {1, 4, 16}
  a.n_units = (p.n_rays + a.unit_rays - 1) / a.unit_rays;
  a.n_tiles = 0;
  a.trace = p.trace; a.debug_flags = p.debug_flags; a.clk = h->d_clk;
  void (*kern)(const TcKernelArgs) = nullptr;
  switch (h->cfg.precision) {
    case NSR_PREC_BF16X3_TC: kern = k_tc_pass<1, 3, false, true>; break;
    case NSR_PREC_FP16X3_TC: kern = k_tc_pass<0, 3, false, true>; break;
    case NSR_PREC_BF16_TC: kern = k_tc_pass<1, 1, false, true>; break;
    default: return cudaErrorInvalidValue;
  }
  const int grid = (int)(a.n_units < h->sm_count ? a.n_units : h->sm_count);
  const int csize = (h->tc_cluster == 2 && a.n_units >= 2) ? 2 : 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(csize == 2 ? (grid + 1) & ~1 : grid));
  cfg.blockDim = dim3(kThreadsTc);
  cfg.dynamicSmemBytes = kSmemTcBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) return e;
  h->launches += 1;
  return cudaGetLastError();
}

}  // namespace nsr
