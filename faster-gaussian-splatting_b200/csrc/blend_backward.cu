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
//  * same persistent producer/consumer ring as the forward (tile-contiguous records via TMA bulk copies);
//  * phase A, pixel-parallel: every lane replays the forward for its pixel front-to-back (so the
//    transmittance comes from registers, not from stored checkpoints) and emits two scalars per
//    (Gaussian, pixel) pair into a per-warp shared-memory slot: u = G * dL/dalpha and w = T * alpha;
//  * phase B, Gaussian-parallel: once 32 surviving Gaussians are queued, lane k owns slot k and sums its
//    nine gradient components over the warp's 32 pixels straight out of shared memory -- a transpose
//    through shared memory replaces nine warp reductions (or nine shuffles per pixel step) per Gaussian;
//  * the nine sums leave as two 16-byte vector reductions + one scalar (REDG.E.ADD.F32x4) into a
//    packed 48-byte per-Gaussian gradient record.
// The stored conic is pre-scaled to the base-2 exponent domain (fgs_device.cuh); the gradients are
// converted back to the reference's conic convention in phase B.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kBatch = 32;
constexpr int kStages = 6;
constexpr int kConsumerThreads = kBlendWarps * 32;
constexpr int kBwdThreads = kConsumerThreads + 32;

struct StageHdr {
    int tile, offset, count, pad;
};

#ifndef FGS_UNROLL_BWD
#define FGS_UNROLL_BWD 1
#endif
#ifndef FGS_BWD_MAXREG
#define FGS_BWD_MAXREG 80
#endif
constexpr int kUnroll = FGS_UNROLL_BWD;   // Gaussians evaluated per phase-A iteration
constexpr int kSlots = 16;  // surviving Gaussians queued per warp before phase B runs

struct WarpScratch {
    float2 uw[kSlots][33];  // [slot][pixel] -> (u, w)
    float4 g[32];           // per pixel: dL/d(colour)
    Rec par[kSlots];        // per slot: the Gaussian's record
};

// Phase B: lanes k and k+16 own slot k and each reduce it over 16 of the warp's 32 pixels straight out
// of shared memory; one shuffle step joins the halves, lane k emits the nine gradient components.
__device__ __noinline__ void flush_slots(WarpScratch& ws, const int nslots, const int lane, const float fx0,
                                         const float fy0, const bool proper_aa, Grad2D* __restrict__ grad2d) {
    __syncwarp();
    const int slot = lane & (kSlots - 1);
    const int p0 = lane & kSlots;  // 0 or 16: first pixel of this half
    const float4 q0 = ws.par[slot].q0;
    const float4 q1 = ws.par[slot].q1;
    const float mx = q0.x, my = q0.y, op = q1.y;
    float dop = 0.f, dcr = 0.f, dcg = 0.f, dcb = 0.f;
    float mxx = 0.f, mxy = 0.f, myy = 0.f, mx1 = 0.f, my1 = 0.f;
    if (slot < nslots) {
#pragma unroll 8
        for (int i = 0; i < 16; ++i) {
            const int p = p0 + i;
            const float2 uw = ws.uw[slot][p];
            const float4 g4 = ws.g[p];
            const float dx = mx - (fx0 + (float)(p & 7));
            const float dy = my - (fy0 + (float)(p >> 3));
            const float h = -op * uw.x;  // -alpha * dL/dalpha
            dop += uw.x;
            dcr = fmaf(uw.y, g4.x, dcr);
            dcg = fmaf(uw.y, g4.y, dcg);
            dcb = fmaf(uw.y, g4.z, dcb);
            const float hx = h * dx, hy = h * dy;
            mx1 += hx;
            my1 += hy;
            mxx = fmaf(hx, dx, mxx);
            mxy = fmaf(hx, dy, mxy);
            myy = fmaf(hy, dy, myy);
        }
    }
    dop += __shfl_xor_sync(0xffffffffu, dop, 16);
    dcr += __shfl_xor_sync(0xffffffffu, dcr, 16);
    dcg += __shfl_xor_sync(0xffffffffu, dcg, 16);
    dcb += __shfl_xor_sync(0xffffffffu, dcb, 16);
    mx1 += __shfl_xor_sync(0xffffffffu, mx1, 16);
    my1 += __shfl_xor_sync(0xffffffffu, my1, 16);
    mxx += __shfl_xor_sync(0xffffffffu, mxx, 16);
    mxy += __shfl_xor_sync(0xffffffffu, mxy, 16);
    myy += __shfl_xor_sync(0xffffffffu, myy, 16);
    if (lane < nslots) {
        const float4 q2 = ws.par[slot].q2;
        // back to the reference's conic: A = A2 / (-0.5 log2 e), B = B2 / (-log2 e), C = C2 / (-0.5 log2 e)
        const float cA = q0.z * (-2.0f / kLog2e), cB = q0.w * (-1.0f / kLog2e), cC = q1.x * (-2.0f / kLog2e);
        const float dmx = cA * mx1 + cB * my1;
        const float dmy = cB * mx1 + cC * my1;
        const float dopacity = proper_aa ? dop : op * (1.0f - op) * dop;
        Grad2D* dst = grad2d + __float_as_uint(q2.w);
        atomicAdd(&dst->g0, make_float4(dmx, dmy, 0.5f * mxx, 0.5f * mxy));
        atomicAdd(&dst->g1, make_float4(0.5f * myy, dopacity, dcr * color_mask(q1.z), dcg * color_mask(q1.w)));
        atomicAdd(&dst->g2.x, dcb * color_mask(q2.x));
    }
    __syncwarp();
}

struct __align__(128) BwdSmem {
    Rec recs[kStages][kBatch];
    StageHdr hdr[kStages];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    WarpScratch ws[kBlendWarps];
};

__global__ void __maxnreg__(FGS_BWD_MAXREG) blend_backward_kernel(const BlendBwdArgs a, const Cam cam) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kBlendWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kBlendWarps) {
        // ================================ producer warp ================================
        if (lane != 0) return;
        int stage = 0;
        uint32_t phase = 0;
        // two-deep software pipeline of the tile queue (see blend_forward.cu)
        int t1 = (int)atomicAdd(a.queue, 1u);
        uint2 r1 = make_uint2(0u, 0u);
        uint32_t m1 = 0;
        if (t1 < a.n_tiles) {
            r1 = __ldg(a.tile_ranges + t1);
            m1 = __ldg(a.tile_max_n + t1);
        }
        int t2 = (int)atomicAdd(a.queue, 1u);
        for (;;) {
            const int tile = t1;
            if (tile >= a.n_tiles) break;
            const uint2 range = r1;
            // nothing past the tile's last contributor matters (kernels_backward.cuh:295)
            const int n = min((int)(range.y - range.x), (int)m1);
            t1 = t2;
            if (t1 < a.n_tiles) {
                r1 = __ldg(a.tile_ranges + t1);
                m1 = __ldg(a.tile_max_n + t1);
                t2 = (int)atomicAdd(a.queue, 1u);
            } else {
                t2 = a.n_tiles;
            }
            if (n <= 0) continue;
            const int nb = (n + kBatch - 1) / kBatch;
            for (int b = 0; b < nb; ++b) {
                mbar_wait_sleep<256>(&sm.empty[stage], phase ^ 1u);
                const int cnt = min(kBatch, n - b * kBatch);
                sm.hdr[stage].tile = tile;
                sm.hdr[stage].offset = b * kBatch;
                sm.hdr[stage].count = cnt;
                mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)cnt * (uint32_t)sizeof(Rec));
                bulk_g2s(&sm.recs[stage][0], a.inst_recs + range.x + (uint32_t)(b * kBatch),
                         (uint32_t)cnt * (uint32_t)sizeof(Rec), &sm.full[stage]);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        mbar_wait_sleep<256>(&sm.empty[stage], phase ^ 1u);
        sm.hdr[stage].tile = -1;
        sm.hdr[stage].offset = 0;
        sm.hdr[stage].count = 0;
        mbar_arrive(&sm.full[stage]);
        return;
    }

    // ================================== consumer warps ==================================
    WarpScratch& ws = sm.ws[warp];
    const float3 bg = make_float3(__ldg(cam.bg), __ldg(cam.bg + 1), __ldg(cam.bg + 2));
    const uint32_t n_pixels = (uint32_t)cam.img_w * (uint32_t)cam.img_h;
    int cur_tile = -2;  // distinct from the end-of-stream marker (-1): a CTA may receive no tile at all
    // per-pixel state (phase A)
    float sr = 0.f, sg = 0.f, sb = 0.f;  // colour still to come behind the current Gaussian ("colour after")
    float gr = 0.f, gg = 0.f, gb = 0.f;  // dL/dpixel colour
    float galpha = 0.f;                  // T_final * -(bg . g)
    float T = 1.f;
    uint32_t n_proc = 0, warp_max = 0;
    float px = 0.f, py = 0.f;
    uint32_t sub_x0 = 0, sub_y0 = 0;
    int nslots = 0;

    auto flush = [&]() {
        flush_slots(ws, nslots, lane, (float)sub_x0 + 0.5f, (float)sub_y0 + 0.5f, cam.proper_aa != 0, a.grad2d);
        nslots = 0;
    };

    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait_sleep<32>(&sm.full[stage], phase);
        const int tile = sm.hdr[stage].tile;
        const int offset = sm.hdr[stage].offset;
        const int count = sm.hdr[stage].count;
        if (tile != cur_tile) {
            if (nslots > 0) flush();
            if (tile < 0) break;
            cur_tile = tile;
            const uint32_t tile_x = (uint32_t)tile % (uint32_t)cam.grid_w, tile_y = (uint32_t)tile / (uint32_t)cam.grid_w;
            sub_x0 = tile_x * kTileW + (uint32_t)(warp & 1) * kSubW;
            sub_y0 = tile_y * kTileH + (uint32_t)(warp >> 1) * kSubH;
            const uint32_t pix_x = sub_x0 + (uint32_t)(lane & 7);
            const uint32_t pix_y = sub_y0 + (uint32_t)(lane >> 3);
            const bool inside = pix_x < (uint32_t)cam.img_w && pix_y < (uint32_t)cam.img_h;
            px = (float)pix_x + 0.5f;
            py = (float)pix_y + 0.5f;
            T = 1.f;
            n_proc = 0;
            gr = gg = gb = 0.f;
            sr = sg = sb = 0.f;
            galpha = 0.f;
            if (inside) {
                const uint32_t pidx = (uint32_t)cam.img_w * pix_y + pix_x;
                n_proc = a.n_processed[pidx];
                const float Tf = a.final_T[pidx];
                gr = a.grad_image[pidx];
                gg = a.grad_image[n_pixels + pidx];
                gb = a.grad_image[2 * n_pixels + pidx];
                sr = a.image[pidx] - Tf * bg.x;
                sg = a.image[n_pixels + pidx] - Tf * bg.y;
                sb = a.image[2 * n_pixels + pidx] - Tf * bg.z;
                galpha = Tf * -(gr * bg.x + gg * bg.y + gb * bg.z);
            }
            warp_max = __reduce_max_sync(0xffffffffu, n_proc);
            ws.g[lane] = make_float4(gr, gg, gb, 0.f);
            __syncwarp();
        }
        if ((uint32_t)offset < warp_max) {
            const Rec* recs = &sm.recs[stage][0];
            const float fx0 = (float)sub_x0, fy0 = (float)sub_y0;
            // one 32-record batch = one cull group
            bool hit = false;
            if (lane < count && (uint32_t)(offset + lane) < warp_max) {
                const float4 q2 = recs[lane].q2;
                const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
                hit = (bx & 0xffffu) < sub_x0 + kSubW && sub_x0 < (bx >> 16) && (by & 0xffffu) < sub_y0 + kSubH &&
                      sub_y0 < (by >> 16);
                if (hit) {
                    const float4 q0 = recs[lane].q0;
                    const float4 q1 = recs[lane].q1;
                    hit = rect_contributes2(q0.x - 0.5f, q0.y - 0.5f, q0.z, q0.w, q1.x, fx0, fy0, (float)(kSubW - 1),
                                            (float)(kSubH - 1), __log2f(q1.y * kMinAlphaRcp));
                }
            }
            uint32_t pending = __ballot_sync(0xffffffffu, hit);
            // kUnroll hits per iteration, branch-free: the alpha evaluations and the dL/dalpha arithmetic of
            // different Gaussians are independent; only T (one multiply) and the colour-after terms (one
            // FMA per channel) chain from one Gaussian to the next
            while (pending) {
                int jx[kUnroll];
                bool vx[kUnroll];
#pragma unroll
                for (int k = 0; k < kUnroll; ++k) {
                    vx[k] = pending != 0u;
                    jx[k] = vx[k] ? __ffs((int)pending) - 1 : (k > 0 ? jx[k - 1] : 0);
                    pending &= pending - 1;
                }
                float gauss[kUnroll], alpha[kUnroll], u[kUnroll], w[kUnroll];
                bool cx[kUnroll];
#pragma unroll
                for (int k = 0; k < kUnroll; ++k) {
                    const float4 q0 = recs[jx[k]].q0;
                    const float c2 = recs[jx[k]].q1.x, op = recs[jx[k]].q1.y;
                    gauss[k] = exp2f(fminf(blend_power2(q0.z, q0.w, c2, q0.x - px, q0.y - py), 0.0f));
                    alpha[k] = __fmul_rn(op, gauss[k]);
                    cx[k] = vx[k] && alpha[k] >= kMinAlpha && (uint32_t)(offset + jx[k]) < n_proc;
                }
#pragma unroll
                for (int k = 0; k < kUnroll; ++k) {
                    const float c_r = recs[jx[k]].q1.z, c_g = recs[jx[k]].q1.w, c_b = recs[jx[k]].q2.x;
                    const float ae = cx[k] ? alpha[k] : 0.0f;  // alpha_eff = 0 leaves T and the colour-after terms untouched
                    const float Tk = T;
                    w[k] = __fmul_rn(Tk, ae);
                    sr = fmaf(-w[k], c_r, sr);
                    sg = fmaf(-w[k], c_g, sg);
                    sb = fmaf(-w[k], c_b, sb);
                    const float oma = __fsub_rn(1.0f, ae);
                    T = __fmul_rn(Tk, oma);
                    const float rcp = __fdividef(1.0f, fmaxf(oma, kOneMinusAlphaEps));
                    const float dlda = fmaf(fmaf(-sr, rcp, Tk * c_r), gr,
                                            fmaf(fmaf(-sg, rcp, Tk * c_g), gg,
                                                 fmaf(fmaf(-sb, rcp, Tk * c_b), gb, galpha * rcp)));
                    u[k] = cx[k] ? gauss[k] * dlda : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < kUnroll; ++k) {
                    if (__any_sync(0xffffffffu, cx[k])) {
                        ws.uw[nslots][lane] = make_float2(u[k], w[k]);
                        if (lane < 3)
                            reinterpret_cast<float4*>(&ws.par[nslots])[lane] =
                                reinterpret_cast<const float4*>(&recs[jx[k]])[lane];
                        if (++nslots == kSlots) flush();
                    }
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

cudaError_t launch_blend_backward(const BlendBwdArgs& a, const Cam& cam, cudaStream_t stream) {
    if (a.n_tiles <= 0) return cudaSuccess;
    static int s_grid = 0;
    const size_t smem = sizeof(BwdSmem);
    if (s_grid == 0) {
        cudaError_t e = cudaFuncSetAttribute(blend_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(blend_backward_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, blend_backward_kernel, kBwdThreads, smem);
        if (occ < 1) occ = 1;
        s_grid = sms * occ;
    }
    const int g = s_grid < a.n_tiles ? s_grid : a.n_tiles;
    blend_backward_kernel<<<g, kBwdThreads, smem, stream>>>(a, cam);
    return cudaGetLastError();
}

}  // namespace fgs
