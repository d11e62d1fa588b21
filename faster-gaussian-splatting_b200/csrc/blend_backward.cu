// blend_backward.cu -- gradient of the front-to-back blend w.r.t. the 2D Gaussian quantities.
//
// Semantics: /root/reference/.../rasterization/include/kernels_backward.cuh:260-471 (G1-G6 in
// SURVEY.md §8.1): a Gaussian takes part for a pixel iff its list index is below that pixel's
// n_processed and alpha >= 1/255; dL/dalpha = (T c - C_after/(1-alpha)) . g + (-T_final (bg . g))/(1-alpha)
// with 1/(1-alpha) evaluated as 1/max(1-alpha, 1e-6); colour gradient masked by colour >= 0; the
// min(q,0) clamp is ignored in the derivative; dL/dconic uses the reference's per-entry convention
// (0.5 * h * (dx^2, dx dy, dy^2)); without anti-aliasing the opacity gradient is already w.r.t. the
// raw logit.
//
// New design (no bucket checkpoints, no 223-step shuffle pipeline):
//  * warp-autonomous persistent kernel, same (tile, sub-tile) item queue and per-warp TMA ring as the
//    forward (blend_forward.cu); a warp only streams the prefix of its tile's list that its own 32
//    pixels actually consumed in the forward;
//  * phase A, pixel-parallel: every lane replays the forward for its pixel front-to-back (so the
//    transmittance comes from registers, not from stored checkpoints) and emits two scalars per
//    (Gaussian, pixel) pair into a per-warp shared-memory slot: u = G * dL/dalpha and w = T * alpha;
//  * phase B, Gaussian-parallel: once 16 surviving Gaussians are queued, lanes k and k+16 own slot k
//    and each sum its nine gradient components over 16 of the warp's 32 pixels straight out of shared
//    memory; one shuffle step joins the halves -- a transpose through shared memory replaces nine
//    warp reductions (or nine shuffles per pixel step) per Gaussian;
//  * the nine sums leave as two 16-byte vector reductions + one scalar (REDG.E.ADD.F32x4) into a
//    packed 48-byte per-Gaussian gradient record.
// The stored conic is pre-scaled to the base-2 exponent domain (fgs_device.cuh); the gradients are
// converted back to the reference's conic convention in phase B.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

#ifndef FGS_BWD_MAXREG
#define FGS_BWD_MAXREG 76  // 3 CTAs per SM like 80, but no spill with the hit list; 72 / 96 / 128 measured slower (r2_results.md)
#endif
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kBatch = 32;   // records per bulk copy = one cull group
constexpr int kRing = 2;     // ring slots per warp
constexpr int kSlots = 16;   // surviving Gaussians queued per warp before phase B runs

struct __align__(128) WarpSmem {
    Rec recs[kRing][kBatch];
    float u[kSlots][34];    // [slot][pixel] -> u = alpha * dL/dalpha (row stride 34: 8-byte aligned pixel pairs,
    float w[kSlots][34];    // [slot][pixel] -> w = T * alpha         conflict-free 64-bit reads in phase B)
    float g[3][32];         // per pixel: dL/d(colour), one plane per channel
    Rec par[kSlots];        // per slot: the Gaussian's record
    uint64_t full[kRing];
#if FGS_HITLIST_BWD
    uint16_t hit[kBatch + 2];  // byte offsets, inside the ring slot, of the records that reach this warp's sub-tile
#endif
};

// Phase B: lanes k and k+16 own slot k and each reduce it over 16 of the warp's 32 pixels straight out
// of shared memory, two horizontally adjacent pixels per step with packed FP32x2 arithmetic (FFMA2 /
// FMUL2 / FADD2: half the issue slots of the scalar form); one shuffle step joins the halves, lane k
// emits the nine gradient components.
__device__ __noinline__ void flush_slots(WarpSmem& ws, const int nslots, const int lane, const float fx0,
                                         const float fy0, const bool proper_aa, Grad2D* __restrict__ grad2d) {
    __syncwarp();
    const int slot = lane & (kSlots - 1);
    const int p0 = lane & kSlots;  // 0 or 16: first pixel of this half (two rows of eight pixels)
    const float4 q0 = ws.par[slot].q0;
    const float4 q1 = ws.par[slot].q1;
    const float op = exp2f(q1.y);  // the record carries log2(opacity)
    float2 dop = make_float2(0.f, 0.f), dcr = dop, dcg = dop, dcb = dop;
    float2 mxx = dop, mxy = dop, myy = dop, mx1 = dop, my1 = dop;
    if (slot < nslots) {
        const float2 step = make_float2(-2.0f, -2.0f);
        const float* urow = &ws.u[slot][p0];
        const float* wrow = &ws.w[slot][p0];
        float dy = q0.y - (fy0 + (float)(p0 >> 3));
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            float2 dx = make_float2(q0.x - fx0, q0.x - fx0 - 1.0f);
            const float2 dy2 = make_float2(dy, dy);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = row * 8 + c * 2;
                const float2 u2 = *reinterpret_cast<const float2*>(urow + i);
                const float2 w2 = *reinterpret_cast<const float2*>(wrow + i);
                const float2 gr2 = *reinterpret_cast<const float2*>(&ws.g[0][p0 + i]);
                const float2 gg2 = *reinterpret_cast<const float2*>(&ws.g[1][p0 + i]);
                const float2 gb2 = *reinterpret_cast<const float2*>(&ws.g[2][p0 + i]);
                const float2 h2 = make_float2(-u2.x, -u2.y);  // -alpha * dL/dalpha
                dop = __fadd2_rn(dop, u2);
                dcr = __ffma2_rn(w2, gr2, dcr);
                dcg = __ffma2_rn(w2, gg2, dcg);
                dcb = __ffma2_rn(w2, gb2, dcb);
                const float2 hx = __fmul2_rn(h2, dx), hy = __fmul2_rn(h2, dy2);
                mx1 = __fadd2_rn(mx1, hx);
                my1 = __fadd2_rn(my1, hy);
                mxx = __ffma2_rn(hx, dx, mxx);
                mxy = __ffma2_rn(hx, dy2, mxy);
                myy = __ffma2_rn(hy, dy2, myy);
                dx = __fadd2_rn(dx, step);
            }
            dy -= 1.0f;
        }
    }
    float s_dop = dop.x + dop.y, s_dcr = dcr.x + dcr.y, s_dcg = dcg.x + dcg.y, s_dcb = dcb.x + dcb.y;
    float s_mx1 = mx1.x + mx1.y, s_my1 = my1.x + my1.y, s_mxx = mxx.x + mxx.y, s_mxy = mxy.x + mxy.y,
          s_myy = myy.x + myy.y;
    s_dop += __shfl_xor_sync(0xffffffffu, s_dop, 16);
    s_dcr += __shfl_xor_sync(0xffffffffu, s_dcr, 16);
    s_dcg += __shfl_xor_sync(0xffffffffu, s_dcg, 16);
    s_dcb += __shfl_xor_sync(0xffffffffu, s_dcb, 16);
    s_mx1 += __shfl_xor_sync(0xffffffffu, s_mx1, 16);
    s_my1 += __shfl_xor_sync(0xffffffffu, s_my1, 16);
    s_mxx += __shfl_xor_sync(0xffffffffu, s_mxx, 16);
    s_mxy += __shfl_xor_sync(0xffffffffu, s_mxy, 16);
    s_myy += __shfl_xor_sync(0xffffffffu, s_myy, 16);
    if (lane < nslots) {
        const float4 q2 = ws.par[slot].q2;
        // back to the reference's conic: A = A2 / (-0.5 log2 e), B = B2 / (-log2 e), C = C2 / (-0.5 log2 e)
        const float cA = q0.z * (-2.0f / kLog2e), cB = q0.w * (-1.0f / kLog2e), cC = q1.x * (-2.0f / kLog2e);
        const float dmx = cA * s_mx1 + cB * s_my1;
        const float dmy = cB * s_mx1 + cC * s_my1;
        // sum of G dL/dalpha = s_dop / o; without anti-aliasing the gradient is w.r.t. the raw logit: x o (1 - o)
        const float dopacity = proper_aa ? __fdividef(s_dop, op) : (1.0f - op) * s_dop;
        Grad2D* dst = grad2d + __float_as_uint(q2.w);
        atomicAdd(&dst->g0, make_float4(dmx, dmy, 0.5f * s_mxx, 0.5f * s_mxy));
        atomicAdd(&dst->g1, make_float4(0.5f * s_myy, dopacity, s_dcr * color_mask(q1.z), s_dcg * color_mask(q1.w)));
        atomicAdd(&dst->g2.x, s_dcb * color_mask(q2.x));
    }
    __syncwarp();
}

// Pruning-score variant of phase B (kernels_pruning_scores.cuh:495-500): slot k holds (o dL/dalpha)^2 per pixel,
// the score of a Gaussian is the sum over its pixels.
__device__ __noinline__ void flush_scores(WarpSmem& ws, const int nslots, const int lane, float* __restrict__ scores) {
    __syncwarp();
    const int slot = lane & (kSlots - 1);
    const int p0 = lane & kSlots;
    float acc = 0.f;
    if (slot < nslots) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += ws.u[slot][p0 + i];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (lane < nslots) atomicAdd(scores + __float_as_uint(ws.par[slot].q2.w), acc);
    __syncwarp();
}

// kScores = false: gradients of the blend (training backward).  kScores = true: Speedy-Splat pruning scores
// with dL/d(pixel colour) == 1 (kernels_pruning_scores.cuh:440-506); same traversal, different phase A payload.
template <bool kScores>
__global__ void __maxnreg__(FGS_BWD_MAXREG) blend_backward_kernel(const BlendBwdArgs a, const Cam cam) {
    FGS_DYNAMIC_SMEM(unsigned char, smem_raw, 128);
    WarpSmem* all = reinterpret_cast<WarpSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem& ws = all[warp];
    if (lane == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(&ws.full[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const float3 bg = make_float3(__ldg(cam.bg), __ldg(cam.bg + 1), __ldg(cam.bg + 2));
    const uint32_t n_pixels = (uint32_t)cam.img_w * (uint32_t)cam.img_h;
    const int n_items = a.n_tiles * kBlendWarps;
    auto grab = [&]() -> int {
        int v = 0;
        if (lane == 0) v = (int)atomicAdd(a.queue, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    // two-deep software pipeline of the item queue (see blend_forward.cu)
    int i1 = grab();
    uint2 r1 = make_uint2(0u, 0u);
    uint32_t m1 = 0;
    if (i1 < n_items) {
        r1 = __ldg(a.tile_ranges + i1 / kBlendWarps);
        m1 = __ldg(a.tile_max_n + i1);
    }
    int i2 = grab();
    uint32_t k = 0;
    FGS_STAT_DECL(st_pairs);
    FGS_STAT_DECL(st_useful);
    FGS_STAT_DECL(st_flushes);
    FGS_STAT_DECL(st_slots);

    for (;;) {
        const int item = i1;
        if (item >= n_items) break;
        const uint2 range = r1;
        // nothing past this sub-tile's last contributor matters (kernels_backward.cuh:295)
        const int n = min((int)(range.y - range.x), (int)m1);
        i1 = i2;
        if (i1 < n_items) {
            r1 = __ldg(a.tile_ranges + i1 / kBlendWarps);
            m1 = __ldg(a.tile_max_n + i1);
            i2 = grab();
        } else {
            i2 = n_items;
        }
        if (n <= 0) continue;

        const int tile = item / kBlendWarps, sub = item - tile * kBlendWarps;
        const uint32_t tile_x = (uint32_t)tile % (uint32_t)cam.grid_w, tile_y = (uint32_t)tile / (uint32_t)cam.grid_w;
        const uint32_t sub_x0 = tile_x * kTileW + (uint32_t)(sub & 1) * kSubW;
        const uint32_t sub_y0 = tile_y * kTileH + (uint32_t)(sub >> 1) * kSubH;
        const uint32_t pix_x = sub_x0 + (uint32_t)(lane & 7), pix_y = sub_y0 + (uint32_t)(lane >> 3);
        const bool inside = pix_x < (uint32_t)cam.img_w && pix_y < (uint32_t)cam.img_h;
        const float px = (float)pix_x + 0.5f, py = (float)pix_y + 0.5f;
        const float fx0 = (float)sub_x0, fy0 = (float)sub_y0;

        const int nb = (n + kBatch - 1) / kBatch;
        const Rec* src = a.inst_recs + range.x;
        int issued = 0;
        auto issue = [&]() {
            if (lane == 0) {
                const uint32_t slot = (k + (uint32_t)issued) % kRing;
                const uint32_t bytes = (uint32_t)min(kBatch, n - issued * kBatch) * (uint32_t)sizeof(Rec);
                mbar_arrive_expect_tx(&ws.full[slot], bytes);
                bulk_g2s(&ws.recs[slot][0], src + issued * kBatch, bytes, &ws.full[slot]);
            }
            ++issued;
        };
        while (issued < nb && issued < kRing) issue();

        // per-pixel state: dL/d(pixel colour) g, T_final * -(bg . g), the running transmittance T, and
        // sigma = (colour still to come behind the current Gaussian) . g.  The reference carries the
        // "colour after" vector and dots it with g for every Gaussian (kernels_backward.cuh:429-436); the
        // dot product is linear, so the scalar is carried instead: sigma -= w * (c . g).
        float gr = 0.f, gg = 0.f, gb = 0.f, galpha = 0.f, sigma = 0.f, T = 1.f;
        int n_rel = 0;  // this pixel's last contributor, relative to the current batch
        if (inside) {
            const uint32_t pidx = (uint32_t)cam.img_w * pix_y + pix_x;
            n_rel = (int)a.n_processed[pidx];
            const float Tf = a.final_T[pidx];
            if (kScores) {
                gr = gg = gb = 1.0f;
            } else {
                gr = a.grad_image[pidx];
                gg = a.grad_image[n_pixels + pidx];
                gb = a.grad_image[2 * n_pixels + pidx];
            }
            const float gbg = gr * bg.x + gg * bg.y + gb * bg.z;
            sigma = a.image[pidx] * gr + a.image[n_pixels + pidx] * gg + a.image[2 * n_pixels + pidx] * gb - Tf * gbg;
            galpha = -Tf * gbg;
        }
        ws.g[0][lane] = gr;
        ws.g[1][lane] = gg;
        ws.g[2][lane] = gb;
        __syncwarp();
        int nslots = 0;

        for (int b = 0; b < nb; ++b) {
            const uint32_t kb = k + (uint32_t)b;
            const uint32_t slot = kb % kRing;
            mbar_wait(&ws.full[slot], (kb / kRing) & 1u);
            const Rec* recs = &ws.recs[slot][0];
            const int count = min(kBatch, n - b * kBatch);
            const int offset = b * kBatch;
            // sub-tile culling was done once at gather time (binning.cu): bit `sub` of the record's mask
            const bool hit = lane < count && ((__float_as_uint(recs[lane].q2.y) >> sub) & 1u) != 0u;
            uint32_t pending = __ballot_sync(0xffffffffu, hit);
            // branch-free body: alpha_eff = 0 leaves T and the colour-after terms untouched
#if FGS_HITLIST_BWD
            const int n_hits = __popc(pending);
            if (hit) ws.hit[__popc(pending & ((1u << lane) - 1u))] = (uint16_t)(lane * (int)sizeof(Rec));
            __syncwarp();
            const unsigned char* base = reinterpret_cast<const unsigned char*>(recs);
            const int off_rel = n_rel * (int)sizeof(Rec);  // this pixel's last contributor as a byte offset into the batch
            const uint16_t* hp = ws.hit;
            uint32_t off_next = lds_u16(hp);
#pragma unroll 1
            for (int i = 0; i < n_hits; ++i) {
                const uint32_t off = off_next;
                off_next = lds_u16(++hp);  // one past the end holds a stale offset: loaded, never used
                const Rec* rp = reinterpret_cast<const Rec*>(base + off);
                const float4 q0 = rp->q0;
                const float4 q1 = rp->q1;
                const float c_b = rp->q2.x;
                const float alpha = blend_alpha(q0.z, q0.w, q1.x, q1.y, q0.x - px, q0.y - py);
                const bool c = alpha >= kMinAlpha && (int)off < off_rel;
#else
            while (pending) {
                const int jj = __ffs((int)pending) - 1;
                pending &= pending - 1;
                const Rec* rp = &recs[jj];
                const float4 q0 = recs[jj].q0;
                const float4 q1 = recs[jj].q1;
                const float c_b = recs[jj].q2.x;
                const float alpha = blend_alpha(q0.z, q0.w, q1.x, q1.y, q0.x - px, q0.y - py);
                const bool c = alpha >= kMinAlpha && jj < n_rel;
#endif
                FGS_STAT_ADD(st_pairs, lane == 0);
                FGS_STAT_ADD(st_useful, c);
                const float ae = c ? alpha : 0.0f;
                const float w = __fmul_rn(T, ae);
                const float cg = fmaf(c_b, gb, fmaf(q1.w, gg, q1.z * gr));  // c . g
                sigma = fmaf(-w, cg, sigma);
                const float oma = __fsub_rn(1.0f, ae);
                const float rcp = __fdividef(1.0f, fmaxf(oma, kOneMinusAlphaEps));
                // dL/dalpha = (T c - C_after / (1 - alpha)) . g + galpha / (1 - alpha)
                const float dlda = fmaf(T, cg, rcp * (galpha - sigma));
                T = __fmul_rn(T, oma);
                float u = ae * dlda;  // alpha * dL/dalpha (0 when the pair does not contribute)
                if (kScores) {        // (opacity * dL/dalpha)^2 (kernels_pruning_scores.cuh:496-497)
                    const float od = exp2f(q1.y) * dlda;
                    u = c ? od * od : 0.0f;
                }
                // every Gaussian that reaches the sub-tile takes a slot (98 % of them contribute to at least one
                // pixel: voting on that and branching cost more than the few all-zero slots); lanes 0-2 copy its
                // record out of the ring
                {
                    ws.u[nslots][lane] = u;
                    ws.w[nslots][lane] = w;
                    if (lane < 3)
                        reinterpret_cast<float4*>(&ws.par[nslots])[lane] = reinterpret_cast<const float4*>(rp)[lane];
                    if (++nslots == kSlots) {
                        FGS_STAT_ADD(st_flushes, lane == 0);
                        FGS_STAT_ADD(st_slots, lane == 0 ? kSlots : 0);
                        if (kScores) flush_scores(ws, nslots, lane, a.scores);
                        else flush_slots(ws, nslots, lane, fx0 + 0.5f, fy0 + 0.5f, cam.proper_aa != 0, a.grad2d);
                        nslots = 0;
                    }
                }
            }
            __syncwarp();
            n_rel -= kBatch;
            if (issued < nb) issue();  // refill the slot just consumed
        }
        k += (uint32_t)nb;
        if (nslots > 0) {
            FGS_STAT_ADD(st_flushes, lane == 0);
            FGS_STAT_ADD(st_slots, lane == 0 ? nslots : 0);
            if (kScores) flush_scores(ws, nslots, lane, a.scores);
            else flush_slots(ws, nslots, lane, fx0 + 0.5f, fy0 + 0.5f, cam.proper_aa != 0, a.grad2d);
        }
    }
    FGS_STAT_FLUSH(4, st_pairs);
    FGS_STAT_FLUSH(5, st_useful);
    FGS_STAT_FLUSH(6, st_flushes);
    FGS_STAT_FLUSH(7, st_slots);
}

}  // namespace

template <bool kScores>
static cudaError_t launch_blend_backward_t(const BlendBwdArgs& a, const Cam& cam, cudaStream_t stream) {
    static PerDevice<int> s_grid_dev;
    int& s_grid = s_grid_dev.here();
    const size_t smem = sizeof(WarpSmem) * kWarps;
    if (s_grid == 0) {  // function attributes are per device too
        cudaError_t e = cudaFuncSetAttribute(blend_backward_kernel<kScores>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(blend_backward_kernel<kScores>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blend_backward_kernel<kScores>, kThreads, smem);
        if (occ < 1) occ = 1;
        s_grid = sms * occ;
    }
    const int items = a.n_tiles * kBlendWarps;
    const int g = s_grid < (items + kWarps - 1) / kWarps ? s_grid : (items + kWarps - 1) / kWarps;
    FGS_LAUNCH_SMEM(blend_backward_kernel<kScores>, g, kThreads, smem, stream, a, cam);
    return cudaGetLastError();
}

cudaError_t launch_blend_backward(const BlendBwdArgs& a, const Cam& cam, cudaStream_t stream) {
    if (a.n_tiles <= 0) return cudaSuccess;
    return a.scores ? launch_blend_backward_t<true>(a, cam, stream) : launch_blend_backward_t<false>(a, cam, stream);
}

}  // namespace fgs
