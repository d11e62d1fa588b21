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
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kBatch = 64;
constexpr int kStages = 4;
constexpr int kConsumerThreads = kBlendWarps * 32;
constexpr int kBwdThreads = kConsumerThreads + 32;

struct StageHdr {
    int tile, offset, count, pad;
};

struct WarpScratch {
    float u[32][33];
    float w[32][33];
    float4 g[32];
};

struct __align__(128) BwdSmem {
    Rec recs[kStages][kBatch];
    StageHdr hdr[kStages];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    WarpScratch ws[kBlendWarps];
};

__global__ void __launch_bounds__(kBwdThreads) blend_backward_kernel(const BlendBwdArgs a, const Cam cam) {
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
        for (;;) {
            const int tile = (int)atomicAdd(a.queue, 1u);
            if (tile >= a.n_tiles) break;
            const uint2 range = a.tile_ranges[tile];
            // nothing past the tile's last contributor matters (kernels_backward.cuh:295)
            const int n = min((int)(range.y - range.x), (int)a.tile_max_n[tile]);
            if (n <= 0) continue;
            const int nb = (n + kBatch - 1) / kBatch;
            for (int b = 0; b < nb; ++b) {
                mbar_wait(&sm.empty[stage], phase ^ 1u);
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
        mbar_wait(&sm.empty[stage], phase ^ 1u);
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
    // per-slot state (phase B): lane k owns slot k
    int nslots = 0;
    float s_mx = 0.f, s_my = 0.f, s_A = 0.f, s_B = 0.f, s_C = 0.f, s_o = 0.f;
    float s_mr = 0.f, s_mg = 0.f, s_mb = 0.f;  // colour >= 0 masks
    uint32_t s_prim = 0;

    auto flush = [&]() {
        __syncwarp();
        if (lane < nslots) {
            float dop = 0.f, dcr = 0.f, dcg = 0.f, dcb = 0.f;
            float mxx = 0.f, mxy = 0.f, myy = 0.f, mx1 = 0.f, my1 = 0.f;
            const float fx0 = (float)sub_x0 + 0.5f, fy0 = (float)sub_y0 + 0.5f;
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float u = ws.u[lane][p];
                const float w = ws.w[lane][p];
                const float4 g4 = ws.g[p];
                const float dx = s_mx - (fx0 + (float)(p & 7));
                const float dy = s_my - (fy0 + (float)(p >> 3));
                const float h = -s_o * u;  // -alpha * dL/dalpha
                dop += u;
                dcr = fmaf(w, g4.x, dcr);
                dcg = fmaf(w, g4.y, dcg);
                dcb = fmaf(w, g4.z, dcb);
                const float hx = h * dx, hy = h * dy;
                mx1 += hx;
                my1 += hy;
                mxx = fmaf(hx, dx, mxx);
                mxy = fmaf(hx, dy, mxy);
                myy = fmaf(hy, dy, myy);
            }
            const float dmx = s_A * mx1 + s_B * my1;
            const float dmy = s_B * mx1 + s_C * my1;
            const float dopacity = cam.proper_aa ? dop : s_o * (1.0f - s_o) * dop;
            Grad2D* dst = a.grad2d + s_prim;
            atomicAdd(&dst->g0, make_float4(dmx, dmy, 0.5f * mxx, 0.5f * mxy));
            atomicAdd(&dst->g1, make_float4(0.5f * myy, dopacity, dcr * s_mr, dcg * s_mg));
            atomicAdd(&dst->g2.x, dcb * s_mb);
        }
        __syncwarp();
        nslots = 0;
    };

    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&sm.full[stage], phase);
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
            __syncwarp();
            ws.g[lane] = make_float4(gr, gg, gb, 0.f);
            __syncwarp();
        }
        if ((uint32_t)offset < warp_max) {
            const Rec* recs = &sm.recs[stage][0];
            for (int g0 = 0; g0 < count && (uint32_t)(offset + g0) < warp_max; g0 += 32) {
                const int j = g0 + lane;
                bool hit = false;
                if (j < count) {
                    const float4 q2 = recs[j].q2;
                    const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
                    hit = (bx & 0xffffu) < sub_x0 + kSubW && sub_x0 < (bx >> 16) && (by & 0xffffu) < sub_y0 + kSubH &&
                          sub_y0 < (by >> 16);
                    if (hit) {
                        const float4 q0 = recs[j].q0;
                        const float4 q1 = recs[j].q1;
                        hit = rect_contributes(q0.x - 0.5f, q0.y - 0.5f, q0.z, q0.w, q1.x, (float)sub_x0, (float)sub_y0,
                                               (float)(kSubW - 1), (float)(kSubH - 1), __logf(q1.y * kMinAlphaRcp));
                    }
                }
                uint32_t pending = __ballot_sync(0xffffffffu, hit);
                while (pending) {
                    const int jj = g0 + __ffs((int)pending) - 1;
                    pending &= pending - 1;
                    const float4 q0 = recs[jj].q0;
                    const float4 q1 = recs[jj].q1;
                    const float4 q2 = recs[jj].q2;
                    float u = 0.f, w = 0.f;
                    bool contrib = false;
                    if ((uint32_t)(offset + jj) < n_proc) {
                        const float dx = q0.x - px, dy = q0.y - py;
                        const float power = blend_power(q0.z, q0.w, q1.x, dx, dy);
                        const float gauss = __expf(fminf(power, 0.0f));
                        const float alpha = __fmul_rn(q1.y, gauss);
                        if (alpha >= kMinAlpha) {
                            contrib = true;
                            w = __fmul_rn(T, alpha);
                            const float c_r = fmaxf(q1.z, 0.0f), c_g = fmaxf(q1.w, 0.0f), c_b = fmaxf(q2.x, 0.0f);
                            sr = fmaf(-w, c_r, sr);
                            sg = fmaf(-w, c_g, sg);
                            sb = fmaf(-w, c_b, sb);
                            const float oma = __fsub_rn(1.0f, alpha);
                            const float rcp = 1.0f / fmaxf(oma, kOneMinusAlphaEps);
                            const float dlda = (T * c_r - sr * rcp) * gr + (T * c_g - sg * rcp) * gg +
                                               (T * c_b - sb * rcp) * gb + galpha * rcp;
                            u = gauss * dlda;
                            T = __fmul_rn(T, oma);
                        }
                    }
                    if (__any_sync(0xffffffffu, contrib)) {
                        ws.u[nslots][lane] = u;
                        ws.w[nslots][lane] = w;
                        if (lane == nslots) {
                            s_mx = q0.x;
                            s_my = q0.y;
                            s_A = q0.z;
                            s_B = q0.w;
                            s_C = q1.x;
                            s_o = q1.y;
                            s_mr = q1.z >= 0.0f ? 1.0f : 0.0f;
                            s_mg = q1.w >= 0.0f ? 1.0f : 0.0f;
                            s_mb = q2.x >= 0.0f ? 1.0f : 0.0f;
                            s_prim = __float_as_uint(q2.w);
                        }
                        if (++nslots == 32) flush();
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
