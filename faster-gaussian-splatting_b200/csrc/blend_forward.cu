// blend_forward.cu -- persistent, warp-autonomous front-to-back alpha blend for sm_100a.
//
// Semantics: /root/reference/.../rasterization/include/kernels_forward.cuh:362-498 and
// kernels_inference.cuh:348-463 (B1-B8 in SURVEY.md §8.1): 16x12 tiles cut into six 8x4 sub-tiles
// (lane -> pixel (lane % 8, lane / 8)), alpha = o * exp(min(q, 0)), skip alpha < 1/255, no 0.99
// clamp, "blend, then test T < 1e-4", background added at the end, planar CHW float image.
//
// New design (measured history in DESIGN.md):
//  * the unit of work is one (tile, sub-tile) item = 32 pixels; every warp of a persistent CTA pulls
//    items from an atomic queue by itself (two-deep software-pipelined fetch), so warps never wait
//    for each other: no block barrier, no producer/consumer hand-off, no intra-tile load imbalance;
//  * each warp streams its tile's tile-contiguous 48-byte records through a private 3-slot ring in
//    shared memory with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx = TMA, SASS UBLKCP)
//    issued by its own lane 0 two batches ahead of use; the six warps of a tile re-read the same
//    records from L2 (the tile's list stays L2-resident while its six items are in flight);
//  * early termination is per sub-tile: a warp stops streaming as soon as its 32 pixels are done;
//  * sub-tile culling is two-level: the reference's strict bounding-box overlap, then the exact
//    "maximum of the Gaussian inside the rectangle" test on the 8x4 sub-tile (the predicate the
//    reference applies to 16x12 tiles, evaluated once at gather time with a small slack on the threshold, binning.cu);
//    both only remove pairs whose alpha is below 1/255 on every pixel of the sub-tile;
//  * the exponent is evaluated in the base-2 domain on a pre-scaled conic, colours arrive pre-clamped,
//    the inner loop is branch-free;
//  * no bucket checkpoints are written: the backward pass recomputes transmittance front-to-back and
//    only needs the per-pixel (final T, last contributor) and the per-item last contributor.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kWarps = 8;                  // warps per CTA, all identical
constexpr int kThreads = kWarps * 32;
constexpr int kBatch = 32;                 // records per bulk copy (1.5 KB) = one cull group
constexpr int kRing = 3;                   // ring slots per warp

struct __align__(128) WarpRing {
    Rec recs[kRing][kBatch];
    uint64_t full[kRing];
#if FGS_HITLIST_FWD
    uint16_t hit[kBatch + 2];  // byte offsets, inside the ring slot, of the records that reach this warp's sub-tile (fits the padding)
#endif
};

template <bool kTraining>
#if FGS_HITLIST_FWD
__global__ void __launch_bounds__(kThreads, 6) blend_forward_kernel(const BlendFwdArgs a, const Cam cam) {
#else
__global__ void __launch_bounds__(kThreads) blend_forward_kernel(const BlendFwdArgs a, const Cam cam) {
#endif
    __shared__ WarpRing rings[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpRing& R = rings[warp];
    if (lane == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(&R.full[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const float3 bg = make_float3(__ldg(cam.bg), __ldg(cam.bg + 1), __ldg(cam.bg + 2));
    const int n_items = a.n_tiles * kBlendWarps;
    auto grab = [&]() -> int {
        int v = 0;
        if (lane == 0) v = (int)atomicAdd(a.queue, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    // two-deep software pipeline of the item queue: the atomic for item t+2 and the range load of item
    // t+1 are in flight while item t is blended
    int i1 = grab();
    uint2 r1 = i1 < n_items ? __ldg(a.tile_ranges + i1 / kBlendWarps) : make_uint2(0u, 0u);
    int i2 = grab();
    uint32_t k = 0;  // batches issued by this warp so far: slot = k % kRing, parity = (k / kRing) & 1
    FGS_STAT_DECL(st_pairs);
    FGS_STAT_DECL(st_useful);
    FGS_STAT_DECL(st_batches);
    FGS_STAT_DECL(st_empty);
    FGS_STAT_DECL(st_live);

    for (;;) {
        const int item = i1;
        if (item >= n_items) break;
        const uint2 range = r1;
        i1 = i2;
        r1 = i1 < n_items ? __ldg(a.tile_ranges + i1 / kBlendWarps) : make_uint2(0u, 0u);
        i2 = i1 < n_items ? grab() : n_items;

        const int tile = item / kBlendWarps, sub = item - tile * kBlendWarps;
        const uint32_t tile_x = (uint32_t)tile % (uint32_t)cam.grid_w, tile_y = (uint32_t)tile / (uint32_t)cam.grid_w;
        const uint32_t sub_x0 = tile_x * kTileW + (uint32_t)(sub & 1) * kSubW;
        const uint32_t sub_y0 = tile_y * kTileH + (uint32_t)(sub >> 1) * kSubH;
        const uint32_t pix_x = sub_x0 + (uint32_t)(lane & 7), pix_y = sub_y0 + (uint32_t)(lane >> 3);
        const bool inside = pix_x < (uint32_t)cam.img_w && pix_y < (uint32_t)cam.img_h;
        const float px = (float)pix_x + 0.5f, py = (float)pix_y + 0.5f;
        float cr = 0.f, cg = 0.f, cb = 0.f, T = 1.f;
        uint32_t n_last = 0;
        // the pixel's "done" flag rides in its alpha threshold: 1/255 while it is live, kDone (> any alpha,
        // alpha <= opacity <= 1) once it left the image or its transmittance fell below 1e-4 -- one compare
        // per Gaussian instead of compare + predicate logic
        constexpr float kDone = 2.0f;
        float thr = inside ? kMinAlpha : kDone;
        bool warp_done = __all_sync(0xffffffffu, !inside);

        const int n = (int)(range.y - range.x);
        const int nb = (n + kBatch - 1) / kBatch;
        const Rec* src = a.inst_recs + range.x;
        int issued = 0;
        auto issue = [&]() {  // batch `issued` of this item -> ring slot (k + issued) % kRing
            if (lane == 0) {
                const uint32_t slot = (k + (uint32_t)issued) % kRing;
                const uint32_t bytes = (uint32_t)min(kBatch, n - issued * kBatch) * (uint32_t)sizeof(Rec);
                mbar_arrive_expect_tx(&R.full[slot], bytes);
                bulk_g2s(&R.recs[slot][0], src + issued * kBatch, bytes, &R.full[slot]);
            }
            ++issued;
        };
        if (!warp_done)
            while (issued < nb && issued < kRing) issue();

        for (int b = 0; b < issued; ++b) {
            const uint32_t kb = k + (uint32_t)b;
            const uint32_t slot = kb % kRing;
            mbar_wait(&R.full[slot], (kb / kRing) & 1u);
            if (!warp_done) {
                const Rec* recs = &R.recs[slot][0];
                const int count = min(kBatch, n - b * kBatch);
                const int offset = b * kBatch;
                // sub-tile culling was done once at gather time (binning.cu): bit `sub` of the record's mask
                const bool hit = lane < count && ((__float_as_uint(recs[lane].q2.y) >> sub) & 1u) != 0u;
                uint32_t pending = __ballot_sync(0xffffffffu, hit);
                FGS_STAT_ADD(st_batches, lane == 0);
                FGS_STAT_ADD(st_empty, lane == 0 && pending == 0u);
                int last_jj = -1;  // this pixel's last contributor inside the batch
#if FGS_HITLIST_FWD
                const int n_hits = __popc(pending);
                if (hit) R.hit[__popc(pending & ((1u << lane) - 1u))] = (uint16_t)(lane * (int)sizeof(Rec));
                __syncwarp();
                const unsigned char* base = reinterpret_cast<const unsigned char*>(recs);
                const uint16_t* hp = R.hit;
                uint32_t off_next = lds_u16(hp);
                int last_off = -1;
#pragma unroll 1
                for (int i = 0; i < n_hits; ++i) {
                    const uint32_t off = off_next;
                    off_next = lds_u16(++hp);  // one past the end holds a stale offset: loaded, never used
                    const Rec* rp = reinterpret_cast<const Rec*>(base + off);
                    const float4 q0 = rp->q0;
                    const float4 q1 = rp->q1;
                    const float colb = rp->q2.x;
#else
                while (pending) {
                    const int jj = __ffs((int)pending) - 1;
                    pending &= pending - 1;
                    const float4 q0 = recs[jj].q0;
                    const float4 q1 = recs[jj].q1;
                    const float colb = recs[jj].q2.x;
#endif
                    const float alpha = blend_alpha(q0.z, q0.w, q1.x, q1.y, q0.x - px, q0.y - py);
                    const bool c = alpha >= thr;
                    FGS_STAT_ADD(st_pairs, lane == 0);
                    FGS_STAT_ADD(st_useful, c);
                    FGS_STAT_ADD(st_live, thr != kDone);
                    const float ae = c ? alpha : 0.0f;  // alpha_eff = 0 leaves the pixel untouched
                    const float w = __fmul_rn(T, ae);
                    cr = fmaf(w, q1.z, cr);
                    cg = fmaf(w, q1.w, cg);
                    cb = fmaf(w, colb, cb);
                    T = __fmul_rn(T, __fsub_rn(1.0f, ae));
#if FGS_HITLIST_FWD
                    if (c) last_off = (int)off;
#else
                    if (c) last_jj = jj;
#endif
                    thr = T < kTransmittanceThreshold ? kDone : thr;
                }
#if FGS_HITLIST_FWD
                if (last_off >= 0) last_jj = last_off / (int)sizeof(Rec);
#endif
                if (last_jj >= 0) n_last = (uint32_t)(offset + last_jj + 1);
                warp_done = __all_sync(0xffffffffu, thr == kDone);
            }
            __syncwarp();
            if (!warp_done && issued < nb) issue();  // refill the slot just consumed
        }
        k += (uint32_t)issued;

        if (inside) {
            cr = fmaf(T, bg.x, cr);
            cg = fmaf(T, bg.y, cg);
            cb = fmaf(T, bg.z, cb);
            const uint32_t pidx = (uint32_t)cam.img_w * pix_y + pix_x;
            const uint32_t np = (uint32_t)cam.img_w * (uint32_t)cam.img_h;
            if (kTraining) {
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
            if (lane == 0) a.tile_max_n[item] = wmax;  // per item: exactly one writer
        }
    }
    FGS_STAT_FLUSH(0, st_pairs);
    FGS_STAT_FLUSH(1, st_useful);
    FGS_STAT_FLUSH(2, st_batches);
    FGS_STAT_FLUSH(3, st_empty);
    FGS_STAT_FLUSH(8, st_live);
}

}  // namespace

cudaError_t launch_blend_forward(const BlendFwdArgs& a, const Cam& cam, cudaStream_t stream) {
    if (a.n_tiles <= 0) return cudaSuccess;
    static PerDevice<int> s_grid_train, s_grid_inf;
    int& grid = a.training ? s_grid_train.here() : s_grid_inf.here();
    if (grid == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (a.training)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blend_forward_kernel<true>, kThreads, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blend_forward_kernel<false>, kThreads, 0);
        if (occ < 1) occ = 1;
        grid = sms * occ;
    }
    const int items = a.n_tiles * kBlendWarps;
    const int g = grid < (items + kWarps - 1) / kWarps ? grid : (items + kWarps - 1) / kWarps;
    if (a.training)
        FGS_LAUNCH(blend_forward_kernel<true>, g, kThreads, stream, a, cam);
    else
        FGS_LAUNCH(blend_forward_kernel<false>, g, kThreads, stream, a, cam);
    return cudaGetLastError();
}

}  // namespace fgs
