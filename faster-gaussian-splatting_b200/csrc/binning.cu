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

// One thread per visible Gaussian, in depth order.  Everything the common case needs arrives with coalesced
// loads: the Gaussian's index, its write offset and its 8-byte tile footprint (rectangle origin/width + the bit
// mask of tiles that passed the exact test in preprocess), carried into depth order by the offsets scan.
// Gaussians whose candidate rectangle has more than 32 tiles (rare, but the nearest -- i.e. first -- ones, so
// one warp used to serialise hundreds of 32-tile rounds: measured, that warp WAS the kernel's duration) are
// only appended to a work list here; create_instances_big_kernel spreads them over whole blocks.
template <typename KeyT>
__global__ void __launch_bounds__(kInstBlock) create_instances_kernel(
    const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ offsets, const uint2* __restrict__ sorted_info,
    KeyT* __restrict__ inst_keys, uint32_t* __restrict__ inst_vals, const uint32_t* __restrict__ n_vis_dev,
    const int grid_w, uint32_t* __restrict__ key_hist, const uint32_t key_hist_mask, uint32_t* __restrict__ big_count,
    uint32_t* __restrict__ big_prims, uint32_t* __restrict__ big_offsets, const uint32_t capacity) {
    // digit histogram of the tile sort's first pass, accumulated while the keys are produced
    __shared__ uint32_t s_hist[kRadixMax];
    s_hist[threadIdx.x] = 0;
    const uint32_t n_vis = *n_vis_dev;
    __syncthreads();
    const uint32_t i = blockIdx.x * kInstBlock + threadIdx.x;
    if (blockIdx.x * kInstBlock >= n_vis) return;  // whole block out of range
    if (i < n_vis) {
        const uint32_t prim = __ldg(sorted_idx + i);
        const uint2 info = __ldg(sorted_info + i);
        uint32_t wr = __ldg(offsets + i);
        if (info.x & kTileInfoBig) {
            const uint32_t slot = atomicAdd(big_count, 1u);
            big_prims[slot] = prim;
            big_offsets[slot] = wr;
        } else {
            const uint32_t tx0 = info.x & 0xfffu, ty0 = (info.x >> 12) & 0xfffu, tw = ((info.x >> 24) & 0x1fu) + 1u;
            // `capacity`: the instance buffers may have been sized speculatively (c_abi.cu); never write past them
            uint32_t m = wr + (uint32_t)__popc(info.y) <= capacity ? info.y : 0u;
            while (m) {
                const uint32_t c = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                const uint32_t key = (ty0 + c / tw) * (uint32_t)grid_w + tx0 + c % tw;
                inst_keys[wr] = (KeyT)key;
                inst_vals[wr] = prim;
                atomicAdd(&s_hist[key & key_hist_mask], 1u);
                ++wr;
            }
        }
    }
    __syncthreads();
    const uint32_t c = s_hist[threadIdx.x];
    if (c) atomicAdd(key_hist + threadIdx.x, c);
}

// The Gaussians with large footprints: one BLOCK per work-list entry and round of 256 candidate tiles; the exact
// tile test (kernels_forward.cuh:224-328 runs it per warp and 32 tiles) is re-run from the Gaussian's record with
// the very expressions preprocess used, so the count matches the offsets that were scanned.
template <typename KeyT>
__global__ void __launch_bounds__(kInstBlock) create_instances_big_kernel(
    const Rec* __restrict__ recs, KeyT* __restrict__ inst_keys, uint32_t* __restrict__ inst_vals, const int grid_w,
    uint32_t* __restrict__ key_hist, const uint32_t key_hist_mask, const uint32_t* __restrict__ big_count,
    const uint32_t* __restrict__ big_prims, const uint32_t* __restrict__ big_offsets, const uint32_t capacity) {
    __shared__ uint32_t s_hist[kRadixMax];
    __shared__ uint32_t s_wcnt[kInstBlock / 32];
    s_hist[threadIdx.x] = 0;
    const uint32_t n_big = *big_count;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t item = blockIdx.x; item < n_big; item += gridDim.x) {  // block-uniform
        const uint32_t prim = __ldg(big_prims + item);
        uint32_t wr = __ldg(big_offsets + item);
        const float4 q0 = __ldg(&recs[prim].q0);
        const float4 q1 = __ldg(&recs[prim].q1);
        const float4 q2 = __ldg(&recs[prim].q2);
        const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
        uint32_t tx0, tx1, ty0, ty1;
        tile_bounds_of(bx & 0xffffu, bx >> 16, by & 0xffffu, by >> 16, tx0, tx1, ty0, ty1);
        const uint32_t tw = tx1 - tx0, cand = tw * (ty1 - ty0);
        const float msx = q0.x - 0.5f, msy = q0.y - 0.5f;
        const float tau = q1.y + kLog2_255;  // base-2 domain, same expression as preprocess
        for (uint32_t c0 = 0; c0 < cand; c0 += kInstBlock) {
            const uint32_t c = c0 + threadIdx.x;
            const uint32_t tx = tx0 + c % tw, ty = ty0 + c / tw;
            const bool ok = c < cand && tile_contributes2(msx, msy, q0.z, q0.w, q1.x, tx, ty, tau);
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s_wcnt[warp] = (uint32_t)__popc(m);
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kInstBlock / 32; ++w) {
                const uint32_t cw = s_wcnt[w];
                before += (uint32_t)w < warp ? cw : 0u;
                total += cw;
            }
            const uint32_t pos = wr + before + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (ok && pos < capacity) {
                const uint32_t key = ty * (uint32_t)grid_w + tx;
                inst_keys[pos] = (KeyT)key;
                inst_vals[pos] = prim;
                atomicAdd(&s_hist[key & key_hist_mask], 1u);
            }
            wr += total;
            __syncthreads();  // s_wcnt is rewritten by the next round
        }
    }
    __syncthreads();
    const uint32_t c = s_hist[threadIdx.x];
    if (c) atomicAdd(key_hist + threadIdx.x, c);
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
                                                            uint2* __restrict__ ranges, const uint32_t* __restrict__ n_inst_dev,
                                                            const uint32_t capacity, const uint32_t grid_w,
                                                            const float4* __restrict__ colors) {
    const uint32_t n_inst = bounded_count(n_inst_dev, capacity);
    const uint32_t inst = blockIdx.x * 256u + threadIdx.x;
    if (inst >= n_inst) return;
    const uint32_t prim = __ldg(vals + inst);
    const uint32_t key = (uint32_t)keys[inst];
    const uint32_t prev = inst > 0 ? (uint32_t)keys[inst - 1] : 0xffffffffu;  // up front: not a third dependent round trip
    const float4 q0 = __ldg(&recs[prim].q0);
    float4 q1 = __ldg(&recs[prim].q1);
    float4 q2 = __ldg(&recs[prim].q2);
    if (colors != nullptr) {  // deferred-colour forward: the record's colour fields are still empty
        const float4 c = __ldg(colors + prim);
        q1.z = c.x;
        q1.w = c.y;
        q2.x = c.z;
    }
    const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
    const uint32_t xmin = bx & 0xffffu, xmax = bx >> 16, ymin = by & 0xffffu, ymax = by >> 16;
    const uint32_t tx0 = (key % grid_w) * kTileW, ty0 = (key / grid_w) * kTileH;
    // the sub-tile cull must be a SUPERSET of the per-pixel test alpha >= 1/255 that the blend kernels apply (they evaluate
    // log2 alpha with a different operation order): a small slack on the threshold keeps a pair whose best pixel sits on
    // the boundary, as the reference does (it only culls sub-tiles by bounding box, kernels_forward.cuh:445-451)
    const float tau2 = q1.y + kLog2_255 + 2e-3f;
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
    } else if (prev != key) {
        ranges[prev].y = inst;
        ranges[key].x = inst;
    }
    if (inst == n_inst - 1) ranges[key].y = n_inst;
}

}  // namespace

template <typename KeyT>
cudaError_t launch_create_instances(const uint32_t* sorted_idx, const uint32_t* offsets, const Rec* recs,
                                    const uint2* sorted_info, KeyT* inst_keys, uint32_t* inst_vals, const uint32_t* n_vis_dev, int n_max,
                                    int grid_w, uint32_t* key_hist, int key_hist_bits, uint32_t* big_count,
                                    uint32_t* big_prims, uint32_t* big_offsets, int capacity, cudaStream_t stream) {
    if (n_max <= 0) return cudaSuccess;
    static_assert(kInstBlock == kRadixMax, "one thread per histogram bin");
    const uint32_t mask = (1u << key_hist_bits) - 1u;
    FGS_LAUNCH(create_instances_kernel<KeyT>, div_up(n_max, kInstBlock), kInstBlock, stream,
        sorted_idx, offsets, sorted_info, inst_keys, inst_vals, n_vis_dev, grid_w, key_hist, mask, big_count, big_prims,
        big_offsets, (uint32_t)capacity);
    static PerDevice<int> s_sms_dev;
    int& s_sms = s_sms_dev.here();
    if (s_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev);
        if (s_sms < 1) s_sms = 1;
    }
    const int big_grid = n_max < 2 * s_sms ? n_max : 2 * s_sms;
    FGS_LAUNCH(create_instances_big_kernel<KeyT>, big_grid, kInstBlock, stream, recs, inst_keys, inst_vals, grid_w, key_hist,
                                                                          mask, big_count, big_prims, big_offsets,
                                                                          (uint32_t)capacity);
    return cudaGetLastError();
}
template cudaError_t launch_create_instances<uint16_t>(const uint32_t*, const uint32_t*, const Rec*, const uint2*, uint16_t*,
                                                       uint32_t*, const uint32_t*, int, int, uint32_t*, int, uint32_t*,
                                                       uint32_t*, uint32_t*, int, cudaStream_t);
template cudaError_t launch_create_instances<uint32_t>(const uint32_t*, const uint32_t*, const Rec*, const uint2*, uint32_t*,
                                                       uint32_t*, const uint32_t*, int, int, uint32_t*, int, uint32_t*,
                                                       uint32_t*, uint32_t*, int, cudaStream_t);

template <typename KeyT>
cudaError_t launch_ranges_and_gather(const KeyT* sorted_keys, const uint32_t* sorted_vals, const Rec* recs,
                                     Rec* inst_recs, uint2* tile_ranges, const uint32_t* n_instances_dev, int capacity,
                                     int grid_w, cudaStream_t stream, const float4* colors) {
    if (capacity <= 0) return cudaSuccess;
    FGS_LAUNCH(ranges_gather_kernel<KeyT>, (unsigned)((capacity + 255) / 256), 256, stream,
        sorted_keys, sorted_vals, recs, inst_recs, tile_ranges, n_instances_dev, (uint32_t)capacity, (uint32_t)grid_w, colors);
    return cudaGetLastError();
}
template cudaError_t launch_ranges_and_gather<uint16_t>(const uint16_t*, const uint32_t*, const Rec*, Rec*, uint2*,
                                                        const uint32_t*, int, int, cudaStream_t, const float4*);
template cudaError_t launch_ranges_and_gather<uint32_t>(const uint32_t*, const uint32_t*, const Rec*, Rec*, uint2*,
                                                        const uint32_t*, int, int, cudaStream_t, const float4*);

}  // namespace fgs
