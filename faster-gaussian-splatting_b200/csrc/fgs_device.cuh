// fgs_device.cuh -- shared device-side definitions of the B200 Gaussian rasterizer.
//
// Semantics restated from the reference (cited per item); the code, data layout and
// kernel decomposition are new and sm_100a-only.
#pragma once
#ifndef FGS_CUDA_EMU
#include <cuda_runtime.h>
#endif
#include <cstdint>

#include "fgs_launch.h"

namespace fgs {

// ---- constants: rasterization_config.h:10-22,55-60 -------------------------------------------------
constexpr float kDilation = 0.3f;
constexpr float kDilationAA = 0.1f;
constexpr float kMinCov2dDet = 1e-6f;
constexpr float kMinAlpha = 1.0f / 255.0f;
constexpr float kMinAlphaRcp = 255.0f;
constexpr float kOneMinusAlphaEps = 1e-6f;
constexpr float kTransmittanceThreshold = 1e-4f;
constexpr int kTileW = 16;
constexpr int kTileH = 12;
constexpr int kSubW = 8;   // a warp owns an 8x4 sub-tile, lane -> (lane % 8, lane / 8)
constexpr int kSubH = 4;
constexpr int kTilePixels = kTileW * kTileH;   // 192
constexpr int kBlendWarps = kTilePixels / 32;  // 6

// ---- per-Gaussian 2D record (48 B, three float4) ---------------------------------------------------
// q0 = (mean2d.x, mean2d.y, A2, B2)
// q1 = (C2, log2(opacity), color.r, color.g)
// q2 = (color.b, bits(xmin | xmax << 16), bits(ymin | ymax << 16), bits(primitive index))
// (A2, B2, C2) is the conic pre-scaled into the base-2 exponent domain,
//      A2 = -0.5 log2(e) conic.x,  B2 = -log2(e) conic.y,  C2 = -0.5 log2(e) conic.z,
// so that the blend evaluates  log2 G = A2 dx^2 + B2 dx dy + C2 dy^2  and feeds MUFU.EX2 directly.
// Colours are stored clamped to >= 0 (what the blend consumes, kernels_forward.cuh:430); a colour
// that was negative before the clamp is stored as -0.0f so the backward still knows the clamp mask
// (kernels_backward.cuh:314-318) from the sign bit.
struct __align__(16) Rec {
    float4 q0, q1, q2;
};
static_assert(sizeof(Rec) == 48, "Rec must be 48 bytes");

// 2D gradient record accumulated by the blend backward (12 floats, three float4 so that one
// REDG.F32x4 covers a third of it):
// g0 = (dL/dmean2d.x, dL/dmean2d.y, dL/dconic.x, dL/dconic.y[per-entry, see kernels_backward.cuh:443-447])
// g1 = (dL/dconic.z, dL/dopacity[w.r.t. raw logit unless AA], dL/dcolor.r, dL/dcolor.g)
// g2 = (dL/dcolor.b, 0, 0, 0)
struct __align__(16) Grad2D {
    float4 g0, g1, g2;
};
static_assert(sizeof(Grad2D) == 48, "Grad2D must be 48 bytes");

// Camera / launch constants passed by value to kernels (host struct mirrored in include/fgs_b200.h).
struct Cam {
    const float* w2c;      // DEVICE pointer: row-major world->camera matrix, rows 0..2 are read (12 floats)
    const float* cam_pos;  // DEVICE pointer: 3 floats
    const float* bg;       // DEVICE pointer: 3 floats
    float width, height;
    float focal_x, focal_y, center_x, center_y;
    float near_plane, far_plane;
    int   grid_w, grid_h;
    int   img_w, img_h;
    int   active_sh_bases;
    int   total_sh_rest;      // sh_rest.size(1)
    int   proper_aa;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Element count of a device-sized array whose buffers were allocated for `capacity` elements.  The instance
// buffers may be sized speculatively (c_abi.cu); if the true count turns out larger, the producer has skipped
// every write that did not fit, so part of the array is uninitialised: consumers then treat it as EMPTY (never
// as "the first `capacity` elements") and the host re-runs the stage with the exact size.
__device__ __forceinline__ uint32_t bounded_count(const uint32_t* __restrict__ count_dev, uint32_t capacity) {
    const uint32_t n = *count_dev;
    return n <= capacity ? n : 0u;
}

// ---- exact tile test: kernel_utils.cuh:71-114 ("will_primitive_contribute") -------------------------
// mean is the 2D mean shifted by -0.5 (pixel-centre coordinates), conic = (A, B, C).
// Rectangle of pixel centres [rx0, rx0 + ext_x] x [ry0, ry0 + ext_y]; the reference uses it with
// 16x12 tiles (ext = 15, 11); the blend kernels reuse it on 8x4 sub-tiles (ext = 7, 3).
__device__ __forceinline__ bool rect_contributes(float mx, float my, float A, float B, float C,
                                                 float rx0, float ry0, float ext_x, float ext_y, float tau) {
    const float rx1 = rx0 + ext_x, ry1 = ry0 + ext_y;
    const float x_min_diff = rx0 - mx;
    const float x_left = x_min_diff >= 0.0f ? 1.0f : 0.0f;
    const float not_in_x = x_left + (mx > rx1 ? 1.0f : 0.0f);
    const float y_min_diff = ry0 - my;
    const float y_above = y_min_diff >= 0.0f ? 1.0f : 0.0f;
    const float not_in_y = y_above + (my > ry1 ? 1.0f : 0.0f);
    if (not_in_x + not_in_y == 0.0f) return true;
    const float cx = x_left != 0.0f ? rx0 : rx1;
    const float cy = y_above != 0.0f ? ry0 : ry1;
    const float diff_x = mx - cx, diff_y = my - cy;
    const float dx = copysignf(ext_x, x_min_diff);
    const float dy = copysignf(ext_y, y_min_diff);
    const float tx = not_in_y * __saturatef((dx * A * diff_x + dx * B * diff_y) / (dx * A * dx));
    const float ty = not_in_x * __saturatef((dy * B * diff_x + dy * C * diff_y) / (dy * C * dy));
    const float ex = mx - (cx + tx * dx), ey = my - (cy + ty * dy);
    const float power = 0.5f * (A * ex * ex + C * ey * ey) + B * ex * ey;
    return power <= tau;
}

__device__ __forceinline__ bool tile_contributes(float mx, float my, float A, float B, float C,
                                                 uint32_t tile_x, uint32_t tile_y, float tau) {
    return rect_contributes(mx, my, A, B, C, (float)(tile_x * kTileW), (float)(tile_y * kTileH),
                            (float)(kTileW - 1), (float)(kTileH - 1), tau);
}

// Same predicate on the pre-scaled conic (A2, B2, C2) and tau2 = log2(255 * opacity):
//   0.5 (A ex^2 + C ey^2) + B ex ey <= tau   <=>   -(A2 ex^2 + B2 ex ey + C2 ey^2) <= tau2,
// and the closest-point parameters only involve ratios of conic entries.
__device__ __forceinline__ bool rect_contributes2(float mx, float my, float A2, float B2, float C2, float rx0,
                                                  float ry0, float ext_x, float ext_y, float tau2) {
    const float rx1 = rx0 + ext_x, ry1 = ry0 + ext_y;
    const float x_min_diff = rx0 - mx;
    const bool x_left = x_min_diff >= 0.0f;
    const bool x_out = x_left || mx > rx1;
    const float y_min_diff = ry0 - my;
    const bool y_above = y_min_diff >= 0.0f;
    const bool y_out = y_above || my > ry1;
    if (!x_out && !y_out) return true;
    const float cx = x_left ? rx0 : rx1;
    const float cy = y_above ? ry0 : ry1;
    const float diff_x = mx - cx, diff_y = my - cy;
    const float dx = copysignf(ext_x, x_min_diff);
    const float dy = copysignf(ext_y, y_min_diff);
    const float tx = y_out ? __saturatef((2.0f * A2 * diff_x + B2 * diff_y) / (2.0f * A2 * dx)) : 0.0f;
    const float ty = x_out ? __saturatef((B2 * diff_x + 2.0f * C2 * diff_y) / (2.0f * C2 * dy)) : 0.0f;
    const float ex = mx - (cx + tx * dx), ey = my - (cy + ty * dy);
    const float neg_power2 = A2 * ex * ex + C2 * ey * ey + B2 * ex * ey;
    return -neg_power2 <= tau2;
}

__device__ __forceinline__ bool tile_contributes2(float mx, float my, float A2, float B2, float C2, uint32_t tile_x,
                                                  uint32_t tile_y, float tau2) {
    return rect_contributes2(mx, my, A2, B2, C2, (float)(tile_x * kTileW), (float)(tile_y * kTileH),
                             (float)(kTileW - 1), (float)(kTileH - 1), tau2);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLog2_255 = 7.994353436858858f;  // tau2 = log2(255 * opacity) = log2(opacity) + log2(255)

// clamp to >= 0, remembering a negative input in the sign bit (-0.0f)
__device__ __forceinline__ float pack_color(float c) { return c >= 0.0f ? c : -0.0f; }
__device__ __forceinline__ float color_mask(float packed) { return (__float_as_uint(packed) >> 31) ? 0.0f : 1.0f; }

// tile rectangle of a screen-space bounding box: kernel_utils.cuh:61-68
__device__ __forceinline__ void tile_bounds_of(uint32_t xmin, uint32_t xmax, uint32_t ymin, uint32_t ymax,
                                               uint32_t& tx0, uint32_t& tx1, uint32_t& ty0, uint32_t& ty1) {
    tx0 = xmin / kTileW;
    tx1 = (xmax + kTileW - 1) / kTileW;
    ty0 = ymin / kTileH;
    ty1 = (ymax + kTileH - 1) / kTileH;
}

// ---- per-Gaussian tile footprint (8 bytes), written by preprocess, carried into depth order by the offsets scan --
// x = tx0 | ty0 << 12 | (tw - 1) << 24 | big << 31,  y = big ? number of touched tiles : bit mask of touched tiles
// (bit c <-> candidate tile (tx0 + c % tw, ty0 + c / tw)); big = the candidate rectangle has more than 32 tiles, in
// which case create_instances re-runs the exact tile test from the Gaussian's record.
constexpr uint32_t kTileInfoBig = 1u << 31;
constexpr int kMaxGridDim = 4096;  // 12 bits per tile coordinate
__device__ __forceinline__ uint2 pack_tile_info(uint32_t tx0, uint32_t ty0, uint32_t tw, uint32_t cand, uint32_t mask,
                                                uint32_t n_touched) {
    const bool big = cand > 32u;
    return make_uint2(tx0 | (ty0 << 12) | (big ? kTileInfoBig : ((tw - 1u) << 24)), big ? n_touched : mask);
}
__device__ __forceinline__ uint32_t tile_info_count(uint2 info) {
    return (info.x & kTileInfoBig) ? info.y : (uint32_t)__popc(info.y);
}

#ifdef FGS_CUDA_EMU
// ---- the same primitives on the test-only CPU emulator (tests/cuda_emu/cuda_emu.h): copies complete at issue, the
// mbarrier is a small state machine behind a mutex.  Checks indexing and phase logic, not asynchrony. ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { cuda_emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { cuda_emu::mbar_arrive(bar, 0); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { cuda_emu::mbar_arrive(bar, bytes); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return cuda_emu::mbar_try_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) std::this_thread::yield();
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    cuda_emu::bulk_copy(smem_dst, gmem_src, bytes, bar);
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_commit() {}
__device__ __forceinline__ void cp_async_wait_all() {}
#else
// ---- mbarrier + bulk async copy (TMA 1-D bulk; SASS: UBLKCP / SYNCS) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; src/dst 16-B aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous global -> shared copy (Ampere-style cp.async, SASS LDGSTS): no register staging, the issuing
// thread only waits at cp_async_wait_all()
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// named barrier over a subset of the CTA, with an AND reduction of a predicate (BAR.RED.AND)
__device__ __forceinline__ bool named_bar_and(int id, int nthreads, bool pred) {
    int r;
    asm volatile(
        "{ .reg .pred q, pr; setp.ne.s32 q, %3, 0; barrier.cta.red.and.pred pr, %1, %2, q; selp.s32 %0, 1, 0, pr; }"
        : "=r"(r)
        : "r"(id), "r"(nthreads), "r"((int)pred)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#endif  // FGS_CUDA_EMU

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// log2 G = A2 dx^2 + B2 dx dy + C2 dy^2 (kernels_forward.cuh:461 in the base-2 domain).  The operation
// order is pinned with explicit round-to-nearest intrinsics so that the forward and the backward
// kernel evaluate bit-identical alphas (the backward replays the forward's skip/stop decisions).
__device__ __forceinline__ float blend_power2(float A2, float B2, float C2, float dx, float dy) {
    const float u = __fmaf_rn(B2, dy, __fmul_rn(A2, dx));
    return __fmaf_rn(__fmul_rn(C2, dy), dy, __fmul_rn(u, dx));
}

// log2(alpha) = log2(opacity) + A2 dx^2 + B2 dx dy + C2 dy^2: the opacity rides in the exponent (one FMA
// instead of FMUL + FMUL); alpha = exp2(min(log2 alpha, log2 opacity)) equals o * exp(min(q, 0)).
__device__ __forceinline__ float blend_alpha(float A2, float B2, float C2, float lop, float dx, float dy) {
    const float u = __fmaf_rn(B2, dy, __fmul_rn(A2, dx));
    const float la = __fmaf_rn(__fmul_rn(C2, dy), dy, __fmaf_rn(u, dx, lop));
    return exp2f(fminf(la, lop));
}

// Blend inner loops: how the warp walks the records of a batch that reach its sub-tile.  0: peel the ballot mask bit by bit
// (ffs = BREV + FLO, two XU-pipe instructions per Gaussian next to the MUFU.EX2 / MUFU.RCP).  1: compact the hits once per
// batch into a list of byte offsets in shared memory (one POPC per lane and batch) and walk the list with broadcast loads.
// Measured on B200 at C2 (profiles/r2_results.md, calls 17-18): backward 0.523 -> 0.498 ms (with 76 registers: 0.493);
// forward unchanged (0.269 -> 0.271), so only the backward uses the list.
#ifndef FGS_HITLIST
#define FGS_HITLIST 0
#endif
#ifndef FGS_HITLIST_FWD
#define FGS_HITLIST_FWD FGS_HITLIST
#endif
#ifndef FGS_HITLIST_BWD
#define FGS_HITLIST_BWD 1
#endif

// 16-bit shared-memory load, zero-extended into a 32-bit register (the C++ form makes nvcc re-pack the half word)
__device__ __forceinline__ uint32_t lds_u16(const uint16_t* p) {
#ifdef FGS_CUDA_EMU
    return (uint32_t)*p;
#else
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
#endif
}

// ---- optional instrumentation of the blend kernels (build with -DFGS_BLEND_STATS=1; never in the product build) -------
// Counters for the issue roofline (profiles/blend_stats.py): how many (Gaussian, sub-tile) pairs the inner loops execute
// and how many of their 32 lanes do useful work.
//  [0] fwd pairs   [1] fwd lanes with alpha >= threshold (blended)   [2] fwd batches consumed   [3] fwd batches without a hit
//  [4] bwd pairs   [5] bwd lanes that contribute                     [6] bwd flushes            [7] bwd slots flushed
//  [8] fwd lanes inside the image and not yet done when the pair was evaluated
#ifdef FGS_BLEND_STATS
#define FGS_STAT_DECL(name) unsigned int name = 0
#define FGS_STAT_ADD(name, v) name += (unsigned int)(v)
#define FGS_STAT_FLUSH(idx, name)                                                        \
    do {                                                                                 \
        const unsigned int tot_ = __reduce_add_sync(0xffffffffu, name);                 \
        if ((threadIdx.x & 31) == 0 && tot_ && a.stats) atomicAdd(&a.stats[idx], (unsigned long long)tot_); \
        name = 0;                                                                        \
    } while (0)
#else
#define FGS_STAT_DECL(name)
#define FGS_STAT_ADD(name, v)
#define FGS_STAT_FLUSH(idx, name)
#endif

// mbarrier wait with explicit back-off: a warp that finds the phase incomplete sleeps instead of
// re-issuing try_wait back to back (measured: the bare spin loop took ~15 % of all issue slots of the
// blend kernels away from the warps that had work)
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(kSleepNs);
}

}  // namespace fgs
