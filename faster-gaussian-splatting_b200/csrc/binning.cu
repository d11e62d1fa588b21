// binning.cu -- instance generation ("duplicateWithKeys"), per-tile ranges and the gather of the
// per-Gaussian records into tile-contiguous instance order.
//
// Semantics: /root/reference/.../rasterization/include/kernels_forward.cuh:224-348 (A3 in SURVEY.md
// §8a): one (tile id, Gaussian index) pair per tile that passes the exact tile test, generated in
// global depth order so that the stable tile sort leaves every tile's list front-to-back.
// New here: the instance records are gathered once into instance order so that the blend kernels
// stream them with 1-D bulk copies (TMA) instead of gathering through an index in the hot loop.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kInstBlock = 256;

template <typename KeyT>
__global__ void __launch_bounds__(kInstBlock) create_instances_kernel(
    const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ offsets, const Rec* __restrict__ recs,
    const uint32_t* __restrict__ tile_mask, KeyT* __restrict__ inst_keys, uint32_t* __restrict__ inst_vals, const uint32_t* __restrict__ n_vis_dev,
    const int grid_w) {
    const uint32_t n_vis = *n_vis_dev;
    const uint32_t i = blockIdx.x * kInstBlock + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    if ((i & ~31u) >= n_vis) return;  // whole warp out of range
    const bool active = i < n_vis;
    const uint32_t prim = __ldg(sorted_idx + (active ? i : n_vis - 1));
    const float4 q0 = __ldg(&recs[prim].q0);
    const float4 q1 = __ldg(&recs[prim].q1);
    const float4 q2 = __ldg(&recs[prim].q2);
    const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
    uint32_t tx0, tx1, ty0, ty1;
    tile_bounds_of(bx & 0xffffu, bx >> 16, by & 0xffffu, by >> 16, tx0, tx1, ty0, ty1);
    const uint32_t tw = tx1 - tx0;
    const uint32_t cand = active ? tw * (ty1 - ty0) : 0u;
    const float msx = q0.x - 0.5f, msy = q0.y - 0.5f;
    const float A = q0.z, B = q0.w, C = q1.x;
    const float tau = q1.y + kLog2_255;  // base-2 domain, same expression as preprocess
    uint32_t wr = active ? __ldg(offsets + i) : 0u;

    if (cand > 0 && cand <= 32u) {
        // rectangles of up to 32 tiles: preprocess kept the result of the exact tile test as a bit mask
        uint32_t m = __ldg(tile_mask + prim);
        while (m) {
            const uint32_t c = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            const uint32_t tx = tx0 + c % tw, ty = ty0 + c / tw;
            inst_keys[wr] = (KeyT)(ty * (uint32_t)grid_w + tx);
            inst_vals[wr] = prim;
            ++wr;
        }
    }
    uint32_t big = __ballot_sync(0xffffffffu, cand > 32u);
    while (big) {
        const int src = __ffs((int)big) - 1;
        big &= big - 1;
        const float bmx = __shfl_sync(0xffffffffu, msx, src), bmy = __shfl_sync(0xffffffffu, msy, src);
        const float bA = __shfl_sync(0xffffffffu, A, src), bB = __shfl_sync(0xffffffffu, B, src);
        const float bC = __shfl_sync(0xffffffffu, C, src), btau = __shfl_sync(0xffffffffu, tau, src);
        const uint32_t bx0 = __shfl_sync(0xffffffffu, tx0, src), by0 = __shfl_sync(0xffffffffu, ty0, src);
        const uint32_t btw = __shfl_sync(0xffffffffu, tw, src), bcand = __shfl_sync(0xffffffffu, cand, src);
        const uint32_t bprim = __shfl_sync(0xffffffffu, prim, src);
        uint32_t bwr = __shfl_sync(0xffffffffu, wr, src);
        for (uint32_t c0 = 0; c0 < bcand; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t tx = bx0 + c % btw, ty = by0 + c / btw;
            const bool ok = c < bcand && tile_contributes2(bmx, bmy, bA, bB, bC, tx, ty, btau);
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const uint32_t pos = bwr + __popc(m & ((1u << lane) - 1u));
                inst_keys[pos] = (KeyT)(ty * (uint32_t)grid_w + tx);
                inst_vals[pos] = bprim;
            }
            bwr += __popc(m);
        }
    }
}

// One thread per instance: gathers the Gaussian's 48-byte record into instance order, replaces the
// bounding box (only needed up to here) by the 6-bit mask of the tile's 8x4 sub-tiles the Gaussian
// actually reaches -- strict bounding-box overlap (kernels_forward.cuh:445-451) and the exact
// maximum-inside-the-rectangle test, evaluated once here instead of once per warp and batch in both
// blend kernels -- and emits the tile range boundaries (kernels_forward.cuh:330-348).
template <typename KeyT>
__global__ void __launch_bounds__(256) ranges_gather_kernel(const KeyT* __restrict__ keys,
                                                            const uint32_t* __restrict__ vals,
                                                            const Rec* __restrict__ recs, Rec* __restrict__ inst_recs,
                                                            uint2* __restrict__ ranges, const uint32_t n_inst,
                                                            const uint32_t grid_w) {
    const uint32_t inst = blockIdx.x * 256u + threadIdx.x;
    if (inst >= n_inst) return;
    const uint32_t prim = __ldg(vals + inst);
    const uint32_t key = (uint32_t)keys[inst];
    const float4 q0 = __ldg(&recs[prim].q0);
    const float4 q1 = __ldg(&recs[prim].q1);
    float4 q2 = __ldg(&recs[prim].q2);
    const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
    const uint32_t xmin = bx & 0xffffu, xmax = bx >> 16, ymin = by & 0xffffu, ymax = by >> 16;
    const uint32_t tx0 = (key % grid_w) * kTileW, ty0 = (key / grid_w) * kTileH;
    const float tau2 = q1.y + kLog2_255;
    uint32_t mask = 0;
#pragma unroll
    for (int sub = 0; sub < kBlendWarps; ++sub) {
        const uint32_t sx = tx0 + (uint32_t)(sub & 1) * kSubW, sy = ty0 + (uint32_t)(sub >> 1) * kSubH;
        bool hit = xmin < sx + kSubW && sx < xmax && ymin < sy + kSubH && sy < ymax;
        if (hit)
            hit = rect_contributes2(q0.x - 0.5f, q0.y - 0.5f, q0.z, q0.w, q1.x, (float)sx, (float)sy, (float)(kSubW - 1),
                                    (float)(kSubH - 1), tau2);
        mask |= (hit ? 1u : 0u) << sub;
    }
    q2.y = __uint_as_float(mask);
    q2.z = 0.0f;
    Rec* dst = inst_recs + inst;
    dst->q0 = q0;
    dst->q1 = q1;
    dst->q2 = q2;
    if (inst == 0) {
        ranges[key].x = 0;
    } else {
        const uint32_t prev = (uint32_t)keys[inst - 1];
        if (prev != key) {
            ranges[prev].y = inst;
            ranges[key].x = inst;
        }
    }
    if (inst == n_inst - 1) ranges[key].y = n_inst;
}

}  // namespace

template <typename KeyT>
cudaError_t launch_create_instances(const uint32_t* sorted_idx, const uint32_t* offsets, const Rec* recs,
                                    const uint32_t* tile_mask, KeyT* inst_keys, uint32_t* inst_vals, const uint32_t* n_vis_dev, int n_max,
                                    int grid_w, cudaStream_t stream) {
    if (n_max <= 0) return cudaSuccess;
    create_instances_kernel<KeyT><<<div_up(n_max, kInstBlock), kInstBlock, 0, stream>>>(
        sorted_idx, offsets, recs, tile_mask, inst_keys, inst_vals, n_vis_dev, grid_w);
    return cudaGetLastError();
}
template cudaError_t launch_create_instances<uint16_t>(const uint32_t*, const uint32_t*, const Rec*, const uint32_t*, uint16_t*,
                                                       uint32_t*, const uint32_t*, int, int, cudaStream_t);
template cudaError_t launch_create_instances<uint32_t>(const uint32_t*, const uint32_t*, const Rec*, const uint32_t*, uint32_t*,
                                                       uint32_t*, const uint32_t*, int, int, cudaStream_t);

template <typename KeyT>
cudaError_t launch_ranges_and_gather(const KeyT* sorted_keys, const uint32_t* sorted_vals, const Rec* recs,
                                     Rec* inst_recs, uint2* tile_ranges, int n_instances, int grid_w,
                                     cudaStream_t stream) {
    if (n_instances <= 0) return cudaSuccess;
    ranges_gather_kernel<KeyT><<<(unsigned)((n_instances + 255) / 256), 256, 0, stream>>>(
        sorted_keys, sorted_vals, recs, inst_recs, tile_ranges, (uint32_t)n_instances, (uint32_t)grid_w);
    return cudaGetLastError();
}
template cudaError_t launch_ranges_and_gather<uint16_t>(const uint16_t*, const uint32_t*, const Rec*, Rec*, uint2*, int,
                                                        int, cudaStream_t);
template cudaError_t launch_ranges_and_gather<uint32_t>(const uint32_t*, const uint32_t*, const Rec*, Rec*, uint2*, int,
                                                        int, cudaStream_t);

}  // namespace fgs
