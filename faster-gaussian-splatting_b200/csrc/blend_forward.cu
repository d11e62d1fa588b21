// blend_forward.cu -- persistent, warp-specialised front-to-back alpha blend for sm_100a.
//
// Semantics: /root/reference/.../rasterization/include/kernels_forward.cuh:362-498 and
// kernels_inference.cuh:348-463 (B1-B8 in SURVEY.md §8.1): 16x12 tiles, each of six warps owns an
// 8x4 sub-tile, alpha = o * exp(min(q, 0)), skip alpha < 1/255, no 0.99 clamp, "blend, then test
// T < 1e-4", background added at the end, planar CHW float image.
//
// New design:
//  * persistent CTAs pull tiles from an atomic queue; a dedicated producer warp streams each tile's
//    tile-contiguous 48-byte records into a shared-memory ring with 1-D bulk async copies
//    (cp.async.bulk + mbarrier complete_tx, i.e. TMA; SASS UBLKCP), six consumer warps blend;
//  * the per-stage header (tile id, sequence number, offset, count) travels with the data, so the
//    producer runs ahead across tile boundaries, the consumers never touch global memory for
//    scheduling and never meet at a CTA-wide barrier: each warp free-runs through the ring;
//  * early termination is a 6-bit "sub-tile finished" mask per tile in shared memory (one atomicOr per
//    warp and tile) instead of a __syncthreads_and every 192 Gaussians;
//  * sub-tile culling is two-level: the reference's strict bounding-box overlap, then the exact
//    "maximum of the Gaussian inside the rectangle" test on the 8x4 sub-tile (the same predicate the
//    reference uses for 16x12 tiles); both only remove pairs whose alpha is below 1/255;
//  * the exponent is evaluated in the base-2 domain on a pre-scaled conic (no log2(e) multiply, no
//    -0.5 factor in the loop), colours arrive pre-clamped;
//  * no bucket checkpoints are written: the backward pass recomputes transmittance front-to-back.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kBatch = 64;                                // records per bulk copy (3 KB)
constexpr int kStages = 8;
#ifndef FGS_UNROLL_FWD
#define FGS_UNROLL_FWD 1
#endif
constexpr int kUnroll = FGS_UNROLL_FWD;                               // Gaussians evaluated per loop iteration
constexpr int kSeqRing = 16;                              // > kStages: tiles in flight per CTA
constexpr int kConsumerThreads = kBlendWarps * 32;        // 192
constexpr int kFwdThreads = kConsumerThreads + 32;        // + producer warp
constexpr uint32_t kAllWarps = (1u << kBlendWarps) - 1u;

struct StageHdr {
    int tile;      // -1: end of stream
    int seq;       // per-CTA tile sequence number
    int offset;    // index of the first record of this batch inside the tile's list
    int count;     // records in this batch
};

struct __align__(128) FwdSmem {
    Rec recs[kStages][kBatch];
    StageHdr hdr[kStages];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint32_t done_mask[kSeqRing];  // bit w set: warp w has finished all its pixels of that tile
};

template <bool kTraining>
__global__ void __launch_bounds__(kFwdThreads) blend_forward_kernel(const BlendFwdArgs a, const Cam cam) {
    __shared__ FwdSmem sm;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kBlendWarps);
        }
        mbar_fence_init();
    }
    if (tid < kSeqRing) sm.done_mask[tid] = 0;
    __syncthreads();

    if (warp == kBlendWarps) {
        // ================================ producer warp ================================
        if (lane != 0) return;
        int stage = 0;
        uint32_t phase = 0;
        int seq = 0;
        // the tile queue is software-pipelined two deep: the atomic for tile t+2 and the range load of
        // tile t+1 are in flight while tile t streams, so their latency never stalls the ring
        int t1 = (int)atomicAdd(a.queue, 1u);
        uint2 r1 = t1 < a.n_tiles ? __ldg(a.tile_ranges + t1) : make_uint2(0u, 0u);
        int t2 = (int)atomicAdd(a.queue, 1u);
        for (;;) {
            const int tile = t1;
            if (tile >= a.n_tiles) break;
            const uint2 range = r1;
            t1 = t2;
            r1 = t1 < a.n_tiles ? __ldg(a.tile_ranges + t1) : make_uint2(0u, 0u);
            t2 = t1 < a.n_tiles ? (int)atomicAdd(a.queue, 1u) : a.n_tiles;
            const int n = (int)(range.y - range.x);
            const int nb = n > 0 ? (n + kBatch - 1) / kBatch : 1;
            volatile uint32_t* dm = &sm.done_mask[seq & (kSeqRing - 1)];
            for (int b = 0; b < nb; ++b) {
                mbar_wait_sleep<256>(&sm.empty[stage], phase ^ 1u);
                // the slot of sequence number seq was last used kSeqRing tiles ago; every consumer has
                // released that tile's stages long before (the ring is only kStages deep)
                if (b == 0) *dm = 0;
                else if (*dm == kAllWarps) break;
                const int cnt = min(kBatch, n - b * kBatch);
                sm.hdr[stage].tile = tile;
                sm.hdr[stage].seq = seq;
                sm.hdr[stage].offset = b * kBatch;
                sm.hdr[stage].count = cnt > 0 ? cnt : 0;
                if (cnt > 0) {
                    mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)cnt * (uint32_t)sizeof(Rec));
                    bulk_g2s(&sm.recs[stage][0], a.inst_recs + range.x + (uint32_t)(b * kBatch),
                             (uint32_t)cnt * (uint32_t)sizeof(Rec), &sm.full[stage]);
                } else {
                    mbar_arrive(&sm.full[stage]);
                }
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            ++seq;
        }
        mbar_wait_sleep<256>(&sm.empty[stage], phase ^ 1u);
        sm.hdr[stage].tile = -1;
        sm.hdr[stage].seq = seq;
        sm.hdr[stage].offset = 0;
        sm.hdr[stage].count = 0;
        mbar_arrive(&sm.full[stage]);
        return;
    }

    // ================================== consumer warps ==================================
    const float3 bg = make_float3(__ldg(cam.bg), __ldg(cam.bg + 1), __ldg(cam.bg + 2));
    int cur_tile = -2;  // distinct from the end-of-stream marker (-1): a CTA may receive no tile at all
    bool warp_done = false;  // every pixel of this warp's sub-tile is finished
    // per-pixel state
    float cr = 0.f, cg = 0.f, cb = 0.f, T = 1.f;
    uint32_t n_last = 0;
    bool done = true, inside = false;
    float px = 0.f, py = 0.f;
    uint32_t pix_x = 0, pix_y = 0;
    uint32_t sub_x0 = 0, sub_y0 = 0;

    auto finalize = [&]() {
        if (inside) {
            cr = fmaf(T, bg.x, cr);
            cg = fmaf(T, bg.y, cg);
            cb = fmaf(T, bg.z, cb);
            const uint32_t pidx = (uint32_t)cam.img_w * pix_y + pix_x;
            if (kTraining) {
                const uint32_t np = (uint32_t)cam.img_w * (uint32_t)cam.img_h;
                a.image[pidx] = cr;
                a.image[np + pidx] = cg;
                a.image[2 * np + pidx] = cb;
                a.final_T[pidx] = T;
                a.n_processed[pidx] = n_last;
            } else {
                if (a.clamp_output) {
                    cr = __saturatef(cr);
                    cg = __saturatef(cg);
                    cb = __saturatef(cb);
                }
                if (a.to_chw) {
                    const uint32_t np = (uint32_t)cam.img_w * (uint32_t)cam.img_h;
                    a.image[pidx] = cr;
                    a.image[np + pidx] = cg;
                    a.image[2 * np + pidx] = cb;
                } else {
                    a.image[3 * pidx] = cr;
                    a.image[3 * pidx + 1] = cg;
                    a.image[3 * pidx + 2] = cb;
                }
            }
        }
        if (kTraining) {
            const uint32_t wmax = __reduce_max_sync(0xffffffffu, n_last);
            if (lane == 0 && wmax > 0) atomicMax(&a.tile_max_n[cur_tile], wmax);
        }
    };

    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait_sleep<32>(&sm.full[stage], phase);
        const int tile = sm.hdr[stage].tile;
        const int offset = sm.hdr[stage].offset;
        const int count = sm.hdr[stage].count;
        if (tile != cur_tile) {
            if (cur_tile >= 0) finalize();
            if (tile < 0) break;
            cur_tile = tile;
            const uint32_t tile_x = (uint32_t)tile % (uint32_t)cam.grid_w, tile_y = (uint32_t)tile / (uint32_t)cam.grid_w;
            sub_x0 = tile_x * kTileW + (uint32_t)(warp & 1) * kSubW;
            sub_y0 = tile_y * kTileH + (uint32_t)(warp >> 1) * kSubH;
            pix_x = sub_x0 + (uint32_t)(lane & 7);
            pix_y = sub_y0 + (uint32_t)(lane >> 3);
            inside = pix_x < (uint32_t)cam.img_w && pix_y < (uint32_t)cam.img_h;
            px = (float)pix_x + 0.5f;
            py = (float)pix_y + 0.5f;
            cr = cg = cb = 0.f;
            T = 1.f;
            n_last = 0;
            done = !inside;
            warp_done = __all_sync(0xffffffffu, done);
            if (warp_done && lane == 0) atomicOr(&sm.done_mask[sm.hdr[stage].seq & (kSeqRing - 1)], 1u << warp);
        }
        if (!warp_done && count > 0) {
            const Rec* recs = &sm.recs[stage][0];
            const float fx0 = (float)sub_x0, fy0 = (float)sub_y0;
            for (int g0 = 0; g0 < count; g0 += 32) {
                const int j = g0 + lane;
                bool hit = false;
                if (j < count) {
                    const float4 q2 = recs[j].q2;
                    const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
                    // strict bounding-box / sub-tile overlap (kernels_forward.cuh:445-451)
                    hit = (bx & 0xffffu) < sub_x0 + kSubW && sub_x0 < (bx >> 16) && (by & 0xffffu) < sub_y0 + kSubH &&
                          sub_y0 < (by >> 16);
                    if (hit) {
                        const float4 q0 = recs[j].q0;
                        const float4 q1 = recs[j].q1;
                        hit = rect_contributes2(q0.x - 0.5f, q0.y - 0.5f, q0.z, q0.w, q1.x, fx0, fy0, (float)(kSubW - 1),
                                                (float)(kSubH - 1), __log2f(q1.y * kMinAlphaRcp));
                    }
                }
                uint32_t pending = __ballot_sync(0xffffffffu, hit);
                // kUnroll hits per iteration, branch-free: the alpha evaluations are independent; T, the
                // accumulated colour and the done flag chain from one Gaussian to the next
                while (pending) {
                    int jx[kUnroll];
                    bool vx[kUnroll];
#pragma unroll
                    for (int k = 0; k < kUnroll; ++k) {
                        vx[k] = pending != 0u;
                        jx[k] = vx[k] ? g0 + __ffs((int)pending) - 1 : (k > 0 ? jx[k - 1] : g0);
                        pending &= pending - 1;
                    }
                    float alpha[kUnroll];
#pragma unroll
                    for (int k = 0; k < kUnroll; ++k) {
                        const float4 q0 = recs[jx[k]].q0;
                        const float c2 = recs[jx[k]].q1.x, op = recs[jx[k]].q1.y;
                        const float gauss = exp2f(fminf(blend_power2(q0.z, q0.w, c2, q0.x - px, q0.y - py), 0.0f));
                        alpha[k] = __fmul_rn(op, gauss);
                        if (!(vx[k] && alpha[k] >= kMinAlpha)) alpha[k] = 0.0f;
                    }
#pragma unroll
                    for (int k = 0; k < kUnroll; ++k) {
                        const bool c = alpha[k] > 0.0f && !done;
                        const float ae = c ? alpha[k] : 0.0f;  // alpha_eff = 0 leaves the pixel untouched
                        const float w = __fmul_rn(T, ae);
                        cr = fmaf(w, recs[jx[k]].q1.z, cr);
                        cg = fmaf(w, recs[jx[k]].q1.w, cg);
                        cb = fmaf(w, recs[jx[k]].q2.x, cb);
                        T = __fmul_rn(T, __fsub_rn(1.0f, ae));
                        if (c) n_last = (uint32_t)(offset + jx[k] + 1);
                        done = done || T < kTransmittanceThreshold;
                    }
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    if (lane == 0) atomicOr(&sm.done_mask[sm.hdr[stage].seq & (kSeqRing - 1)], 1u << warp);
                    break;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
        if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
        }
    }
}

}  // namespace

cudaError_t launch_blend_forward(const BlendFwdArgs& a, const Cam& cam, cudaStream_t stream) {
    if (a.n_tiles <= 0) return cudaSuccess;
    static int s_grid_train = 0, s_grid_inf = 0;
    int& grid = a.training ? s_grid_train : s_grid_inf;
    if (grid == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (a.training)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blend_forward_kernel<true>, kFwdThreads, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blend_forward_kernel<false>, kFwdThreads, 0);
        if (occ < 1) occ = 1;
        grid = sms * occ;
    }
    const int g = grid < a.n_tiles ? grid : a.n_tiles;
    if (a.training)
        blend_forward_kernel<true><<<g, kFwdThreads, 0, stream>>>(a, cam);
    else
        blend_forward_kernel<false><<<g, kFwdThreads, 0, stream>>>(a, cam);
    return cudaGetLastError();
}

}  // namespace fgs
