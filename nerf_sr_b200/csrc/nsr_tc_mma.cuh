// nsr_tc_mma.cuh -- the single-lane tcgen05 issue schedule shared by the fused forward pass (nsr_tc.cu: k_tc_pass) and the
// fused backward dX chain (nsr_train.cu: k_tg_dxchain): a 4-slot ring of 32 KB weight stages (128 rows x 64 k, hi plane |
// lo plane, K-major SWIZZLE_128B), activations / gradients as the A operand in TMEM, a 256-column fp32 accumulator handled as
// two N = 128 halves so that the epilogue of one half overlaps the MMAs of the other and of the next layer.
//
// TMEM map assumed here:  [0,256) accumulator (half n at column 128 n), [256,384) A hi plane, [384,512) A lo plane.
// Barrier array (8 bytes each, index = enum below) at MmaCtx::bar.
#pragma once
#include "nsr_tc_ptx.cuh"

namespace nsr {

constexpr int kStageBytes = 32768;          // 128 rows x 64 k x 2 B, hi then lo
constexpr int kPlaneBytes = 16384;
constexpr int kRing = 4;

// barrier indices
enum {
  B_WFULL = 0,            // [4] weights landed (tx)
  B_WEMPTY = 4,           // [4] stage consumed (tcgen05.commit)
  B_ACCFULL = 8,          // [2] accumulator half complete (commit)
  B_AFREE = 10,           // [2] A operand k chunk 0 / 1 no longer read by this layer (commit)
  B_AREADY = 12,          // [4] epilogue finished 64-column quarter q: acc quarter drained, A k chunk q written
  B_ENCFULL = 16,         // [2] front-end produced tile inputs
  B_COMPREADY = 18,       // epilogue staged the tile's per-sample (rgb, sigma) for compositing
  B_COMPDONE = 19,        // front-end finished compositing the staged tile
  B_COUNT = 20
};

struct MmaCtx {
  uint32_t ring, bar, idesc;      // shared-memory addresses of the weight ring and of the barrier array; instruction descriptor
  uint32_t wmask;                 // 0: the ring is private to this CTA; else the cluster's CTA mask -- the weight stages are
                                  // multicast to all of them, so a slot is free only when EVERY CTA's MMAs have consumed it:
                                  // the stage-consumed commit arrives on B_WEMPTY of every CTA (barrier count = cluster size)
};
__device__ __forceinline__ void ring_release(const MmaCtx& c, int slot) {
  if (c.wmask) tc_commit_mc(c.bar + 8 * (B_WEMPTY + slot), (uint16_t)c.wmask);
  else tc_commit(c.bar + 8 * (B_WEMPTY + slot));
}

// Order of the three split terms of one k-step (timing experiment, tools/gpu_mma_order_ab.sh; the sum is the same up to
// fp32 rounding order): 0 = hi.hi, lo.hi, hi.lo (B repeats); 1 = hi.hi, hi.lo, lo.hi (A repeats); 2 = all four k-steps of a
// term before the next term (each operand plane is walked contiguously)
#ifndef NSR_MMA_ORDER
#define NSR_MMA_ORDER 0
#endif

// 12 (or 4) MMAs of one weight stage with A from TMEM.  N8 = stage index mod 8 (compile time).
template <int PASSES, int N8>
__device__ __forceinline__ void mma_stage_ts(const MmaCtx& c, int half, int chunk, bool first, bool wait_next) {
  constexpr uint64_t HI = (64ull << 32) | (1ull << 46) | (2ull << 61);   // SBO=1024, version 1, SWIZZLE_128B
  constexpr int slot = N8 & 3;
  constexpr int nslot = (N8 + 1) & 3, npar = ((N8 + 1) >> 2) & 1;
  const uint32_t wlo = ((c.ring + slot * kStageBytes) >> 4) | (1u << 16);
  const uint32_t d = 128u * half;
  const uint32_t a_hi = 256u + 32u * chunk;
#if NSR_MMA_ORDER == 2
  if (PASSES == 3) {
#pragma unroll
    for (int k = 0; k < 4; ++k) mma_ts(d, a_hi + 8u * k, HI | (uint64_t)(wlo + 2u * k), c.idesc, (k == 0 && first) ? 0u : 1u);
    if (wait_next) { mbar_wait(c.bar + 8 * (B_WFULL + nslot), npar); tc_fence_after(); }
#pragma unroll
    for (int k = 0; k < 4; ++k) mma_ts(d, a_hi + 128u + 8u * k, HI | (uint64_t)(wlo + 2u * k), c.idesc, 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k) mma_ts(d, a_hi + 8u * k, HI | (uint64_t)(wlo + 1024u + 2u * k), c.idesc, 1u);
    ring_release(c, slot);
    return;
  }
#endif
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t bh = HI | (uint64_t)(wlo + 2u * k), bl = HI | (uint64_t)(wlo + 1024u + 2u * k);
    mma_ts(d, a_hi + 8u * k, bh, c.idesc, (k == 0 && first) ? 0u : 1u);
#if NSR_MMA_ORDER == 1
    if (PASSES == 3) { mma_ts(d, a_hi + 8u * k, bl, c.idesc, 1u); mma_ts(d, a_hi + 128u + 8u * k, bh, c.idesc, 1u); }
#else
    if (PASSES == 3) { mma_ts(d, a_hi + 128u + 8u * k, bh, c.idesc, 1u); mma_ts(d, a_hi + 8u * k, bl, c.idesc, 1u); }
#endif
    // the NEXT stage's weights are waited for here, hidden behind this stage's queued MMAs
    if (k == 1 && wait_next) { mbar_wait(c.bar + 8 * (B_WFULL + nslot), npar); tc_fence_after(); }
  }
  ring_release(c, slot);
}

// Same with A = the tile's encoded inputs in shared memory (first layer and skip layer).
template <int PASSES, int N8>
__device__ __forceinline__ void mma_stage_ss(const MmaCtx& c, uint32_t enc, int half, bool wait_next, bool first = true) {
  constexpr uint64_t HI = (64ull << 32) | (1ull << 46) | (2ull << 61);
  constexpr int slot = N8 & 3;
  constexpr int nslot = (N8 + 1) & 3, npar = ((N8 + 1) >> 2) & 1;
  const uint32_t wlo = ((c.ring + slot * kStageBytes) >> 4) | (1u << 16);
  const uint32_t elo = (enc >> 4) | (1u << 16);
  const uint32_t d = 128u * half;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t bh = HI | (uint64_t)(wlo + 2u * k), bl = HI | (uint64_t)(wlo + 1024u + 2u * k);
    const uint64_t eh = HI | (uint64_t)(elo + 2u * k), el = HI | (uint64_t)(elo + 1024u + 2u * k);
    mma_ss(d, eh, bh, c.idesc, (k == 0 && first) ? 0u : 1u);      // the forward's encoding chunk always opens its half
    if (PASSES == 3) { mma_ss(d, el, bh, c.idesc, 1u); mma_ss(d, eh, bl, c.idesc, 1u); }
    if (k == 1 && wait_next) { mbar_wait(c.bar + 8 * (B_WFULL + nslot), npar); tc_fence_after(); }
  }
  ring_release(c, slot);
}

__device__ __forceinline__ void mma_wait_ready(const MmaCtx& c, int q, uint32_t gpar) {
  mbar_wait(c.bar + 8 * (B_AREADY + q), gpar);
}

// A 256x256 trunk layer (optionally with the 64-wide encoding chunk in front: the skip layer).
// P = stage index of the layer's first stage mod 8.  gpar = parity of the previous layer's A_READY.
// Schedule (accumulator halves n0/n1 = column quarters {0,1}/{2,3}; k chunk c of the A operand =
// quarter c of the previous layer's accumulator):
//   n0: [enc] c0 (needs quarters 0,1 drained + A chunk 0), c1, c2 (A chunk 2), c3 (A chunk 3) -> ACC_FULL[0]
//   n1: [enc] c0 -> A_FREE[0], c1 -> A_FREE[1], c2, c3                                        -> ACC_FULL[1]
// so every quarter-epilogue has >= 1536 cycles of MMA work to hide behind.
// wait_last: whether the layer's final stage prefetch-waits for the NEXT stage's weights (false only when no stage follows).
template <int PASSES, int P, bool SKIP>
__device__ __forceinline__ void mma_layer(const MmaCtx& c, uint32_t enc, uint32_t gpar, bool wait_last = true) {
  constexpr int E = SKIP ? 1 : 0;
  mma_wait_ready(c, 0, gpar); mma_wait_ready(c, 1, gpar); tc_fence_after();
  if (SKIP) mma_stage_ss<PASSES, (P + 0) & 7>(c, enc, 0, true);
  mma_stage_ts<PASSES, (P + E + 0) & 7>(c, 0, 0, !SKIP, true);
  mma_stage_ts<PASSES, (P + E + 1) & 7>(c, 0, 1, false, true);
  mma_wait_ready(c, 2, gpar); tc_fence_after();
  mma_stage_ts<PASSES, (P + E + 2) & 7>(c, 0, 2, false, true);
  mma_wait_ready(c, 3, gpar); tc_fence_after();
  mma_stage_ts<PASSES, (P + E + 3) & 7>(c, 0, 3, false, true);
  tc_commit(c.bar + 8 * (B_ACCFULL + 0));
  if (SKIP) mma_stage_ss<PASSES, (P + E + 4) & 7>(c, enc, 1, true);
  mma_stage_ts<PASSES, (P + 2 * E + 4) & 7>(c, 1, 0, !SKIP, true);
  tc_commit(c.bar + 8 * (B_AFREE + 0));
  mma_stage_ts<PASSES, (P + 2 * E + 5) & 7>(c, 1, 1, false, true);
  tc_commit(c.bar + 8 * (B_AFREE + 1));
  mma_stage_ts<PASSES, (P + 2 * E + 6) & 7>(c, 1, 2, false, true);
  mma_stage_ts<PASSES, (P + 2 * E + 7) & 7>(c, 1, 3, false, wait_last);
  tc_commit(c.bar + 8 * (B_ACCFULL + 1));
}

}  // namespace nsr
