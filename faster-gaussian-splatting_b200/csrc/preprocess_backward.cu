// preprocess_backward.cu -- chain rule from the 2D gradients (mean2d, conic, opacity, colour) to
// the six parameter tensors, optionally fused with the Adam update.
//
// Semantics: /root/reference/.../rasterization/include/kernels_backward.cuh:15-257 (G7-G11 in
// SURVEY.md §8.1) and sh_utils.cuh:71-155; Adam: FasterGSCudaBackend/adam/src/adam.cu:10-34 (G12).
// Rows of invisible Gaussians are exactly zero; the anti-aliasing factor is detached from cov2d;
// densification statistics are accumulated for visible Gaussians.
//
// New here: the derivative is written from the matrix identities (dL/dSigma2 = -Q G Q,
// dL/dSigma3 = (JW)^T D (JW), dL/dv_k = r_k^T D r_k, dL/dR = 2 D R diag(v), R = I + (2/|q|^2) K(q)),
// all 59 per-Gaussian outputs are staged per warp in shared memory and leave as 16-byte coalesced
// stores (the (N,3)/(N,15,3) layouts would otherwise produce 4-byte stores at 12/180-byte strides),
// every output row is written (no separate zero-fill pass), and in fused mode the staged gradients
// are consumed by Adam in the same kernel: parameters and both moments are read and written once,
// the gradient tensors never exist in HBM.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kBlock = 128;
constexpr int kWarps = kBlock / 32;

constexpr float C0 = 0.28209479177387814f;
constexpr float C1 = 0.48860251190291987f;
constexpr float C2a = 1.0925484305920792f, C2b = 0.94617469575755997f, C2c = 0.31539156525251999f,
                C2d = 0.54627421529603959f;
constexpr float C3a = 0.59004358992664352f, C3b = 1.7701307697799304f, C3c = 2.8906114426405538f,
                C3d = 0.45704579946446572f, C3e = 2.2852289973223288f, C3f = 1.865881662950577f,
                C3g = 1.1195289977703462f, C3h = 1.4453057213202769f;

// per-warp staging area (floats): rows of 32 Gaussians for each output tensor
constexpr int kOffMeans = 0;            // 32 x 3
constexpr int kOffScales = 96;          // 32 x 3
constexpr int kOffRot = 192;            // 32 x 4
constexpr int kOffOpac = 320;           // 32 x 1
constexpr int kOffSh0 = 352;            // 32 x 3
constexpr int kOffRest = 448;           // 32 x 45 (input coefficients staged here, gradients written in place)
constexpr int kWarpFloats = 448 + 32 * 45;  // 1888

__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, float step_size, float beta1,
                                            float beta2, float eps, float bias2_sqrt_rcp) {
    const float g2 = g * g;
    m = fmaf(beta1, m - g, g);
    v = fmaf(beta2, v - g2, g2);
    const float denom = sqrtf(v) * bias2_sqrt_rcp + eps;
    p -= step_size * m / denom;
}

// The warp writes (or, fused, applies) `rows` x K staged floats to the contiguous span starting at
// element `first_elem` of a parameter-shaped tensor.
__device__ __forceinline__ void emit_span(const float* __restrict__ stage, float* __restrict__ grad_out,
                                          float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                          size_t first_elem, int n_floats, bool fused, float step_size, float beta1,
                                          float beta2, float eps, float b2, int lane) {
    const bool vec = (n_floats & 3) == 0;
    if (!fused) {
        float* dst = grad_out + first_elem;
        if (vec && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            for (int i = lane; i < (n_floats >> 2); i += 32)
                reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(stage)[i];
        } else {
            for (int i = lane; i < n_floats; i += 32) dst[i] = stage[i];
        }
        return;
    }
    float* pp = p + first_elem;
    float* mm = m + first_elem;
    float* vv = v + first_elem;
    if (vec && ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(mm) | reinterpret_cast<uintptr_t>(vv)) & 15u) == 0) {
        // kU independent 16-byte loads of p, m and v per lane are issued before any of them is consumed:
        // the kernel is a pure HBM stream (24 B in, 12 B out per float), latency is hidden by bytes in flight
        constexpr int kU = 4;
        const int n4 = n_floats >> 2;
        for (int i0 = lane; i0 < n4; i0 += 32 * kU) {
            float4 p4[kU], m4[kU], v4[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + 32 * u;
                if (i < n4) {
                    p4[u] = reinterpret_cast<const float4*>(pp)[i];
                    m4[u] = reinterpret_cast<const float4*>(mm)[i];
                    v4[u] = reinterpret_cast<const float4*>(vv)[i];
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + 32 * u;
                if (i < n4) {
                    const float4 g4 = reinterpret_cast<const float4*>(stage)[i];
                    adam_update(g4.x, p4[u].x, m4[u].x, v4[u].x, step_size, beta1, beta2, eps, b2);
                    adam_update(g4.y, p4[u].y, m4[u].y, v4[u].y, step_size, beta1, beta2, eps, b2);
                    adam_update(g4.z, p4[u].z, m4[u].z, v4[u].z, step_size, beta1, beta2, eps, b2);
                    adam_update(g4.w, p4[u].w, m4[u].w, v4[u].w, step_size, beta1, beta2, eps, b2);
                    reinterpret_cast<float4*>(pp)[i] = p4[u];
                    reinterpret_cast<float4*>(mm)[i] = m4[u];
                    reinterpret_cast<float4*>(vv)[i] = v4[u];
                }
            }
        }
    } else {
        for (int i = lane; i < n_floats; i += 32) {
            float pe = pp[i], me = mm[i], ve = vv[i];
            adam_update(stage[i], pe, me, ve, step_size, beta1, beta2, eps, b2);
            pp[i] = pe;
            mm[i] = me;
            vv[i] = ve;
        }
    }
}

// Fused Adam restricted to the rows of `row_mask` (bit r = row r of the warp's 32 Gaussians; `row_floats` floats per
// row, 9 <= row_floats <= 45 so a float4 touches at most two rows).  float4s without such a row are neither read nor
// written; elements of other rows inside a shared float4 are written back unchanged.  `stage` == nullptr: zero
// gradient (the dense-Adam decay of a Gaussian that was not visible).  Used by the overlapped schedule: the rows of
// invisible Gaussians are updated by adam_rest_invisible_kernel while the blend backward runs, the visible ones here.
template <int kU>
__device__ __forceinline__ void emit_span_rows(const float* __restrict__ stage, float* __restrict__ p, float* __restrict__ m,
                                               float* __restrict__ v, size_t first_elem, int n_floats, int row_floats,
                                               uint32_t row_mask, float step_size, float beta1, float beta2, float eps,
                                               float b2, int lane) {
    float* pp = p + first_elem;
    float* mm = m + first_elem;
    float* vv = v + first_elem;
    const uint32_t magic = (65536u + (uint32_t)row_floats - 1u) / (uint32_t)row_floats;  // e / row_floats == (e * magic) >> 16, e < 32 * 45
    const bool vec = (n_floats & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(mm) | reinterpret_cast<uintptr_t>(vv)) & 15u) == 0;
    if (vec) {
        const int n4 = n_floats >> 2;
        for (int i0 = lane; i0 < n4; i0 += 32 * kU) {
            float4 p4[kU], m4[kU], v4[kU];
            uint32_t sel[kU];  // bit c: element c of the float4 belongs to a row of the mask
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + 32 * u;
                sel[u] = 0u;
                if (i < n4) {
                    const uint32_t e0 = 4u * (uint32_t)i;
                    const uint32_t r0 = (e0 * magic) >> 16, r1 = ((e0 + 3u) * magic) >> 16;
                    const uint32_t first_of_r1 = r1 * (uint32_t)row_floats;  // elements >= this index belong to row r1
#pragma unroll
                    for (uint32_t c = 0; c < 4u; ++c) sel[u] |= ((row_mask >> (e0 + c >= first_of_r1 ? r1 : r0)) & 1u) << c;
                    if (sel[u]) {
                        p4[u] = reinterpret_cast<const float4*>(pp)[i];
                        m4[u] = reinterpret_cast<const float4*>(mm)[i];
                        v4[u] = reinterpret_cast<const float4*>(vv)[i];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + 32 * u;
                if (sel[u]) {
                    const float4 g4 = stage ? reinterpret_cast<const float4*>(stage)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (sel[u] & 1u) adam_update(g4.x, p4[u].x, m4[u].x, v4[u].x, step_size, beta1, beta2, eps, b2);
                    if (sel[u] & 2u) adam_update(g4.y, p4[u].y, m4[u].y, v4[u].y, step_size, beta1, beta2, eps, b2);
                    if (sel[u] & 4u) adam_update(g4.z, p4[u].z, m4[u].z, v4[u].z, step_size, beta1, beta2, eps, b2);
                    if (sel[u] & 8u) adam_update(g4.w, p4[u].w, m4[u].w, v4[u].w, step_size, beta1, beta2, eps, b2);
                    reinterpret_cast<float4*>(pp)[i] = p4[u];
                    reinterpret_cast<float4*>(mm)[i] = m4[u];
                    reinterpret_cast<float4*>(vv)[i] = v4[u];
                }
            }
        }
    } else {
        for (int i = lane; i < n_floats; i += 32) {
            if ((row_mask >> (((uint32_t)i * magic) >> 16)) & 1u) {
                float pe = pp[i], me = mm[i], ve = vv[i];
                adam_update(stage ? stage[i] : 0.0f, pe, me, ve, step_size, beta1, beta2, eps, b2);
                pp[i] = pe;
                mm[i] = me;
                vv[i] = ve;
            }
        }
    }
}

// Overlapped schedule, part 1 (side stream, while the blend backward runs): dense Adam with zero gradient on the
// sh_rest rows of the Gaussians that were not visible in this view -- 76 % of the parameter bytes of 39 % of the
// Gaussians at C2.  At most one small CTA per SM (<= 32 registers, no shared memory) so that it fits next to the
// three resident CTAs of the blend backward.
__global__ void __launch_bounds__(128, 16) adam_rest_invisible_kernel(const uint32_t* __restrict__ n_touched, float* __restrict__ p,
                                                                      float* __restrict__ m, float* __restrict__ v, const int n,
                                                                      const int row_floats, const float step_size,
                                                                      const float beta1, const float beta2, const float eps,
                                                                      const float b2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = (n + 31) >> 5;
    for (int group = blockIdx.x * 4 + warp; group < groups; group += gridDim.x * 4) {
        const int wfirst = group << 5;
        const int rows = min(32, n - wfirst);
        const bool invisible = lane < rows && __ldg(n_touched + wfirst + lane) == 0u;
        const uint32_t mask = __ballot_sync(0xffffffffu, invisible);
        if (mask == 0u) continue;
        emit_span_rows<1>(nullptr, p, m, v, (size_t)wfirst * row_floats, rows * row_floats, row_floats, mask, step_size, beta1,
                          beta2, eps, b2, lane);
    }
}

// SH basis values Y_1..Y_15 of a unit direction (sh_utils.cuh:32-69; same expressions as FGS_SH in the backward kernel)
__device__ __forceinline__ void sh_basis_rest(float x, float y, float z, int active_bases, float* __restrict__ Y) {
#pragma unroll
    for (int k = 0; k < 15; ++k) Y[k] = 0.f;
    if (active_bases > 1) {
        Y[0] = -C1 * y;
        Y[1] = C1 * z;
        Y[2] = -C1 * x;
        if (active_bases > 4) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            Y[3] = C2a * xy;
            Y[4] = -C2a * yz;
            Y[5] = C2b * zz - C2c;
            Y[6] = -C2a * xz;
            Y[7] = C2d * (xx - yy);
            if (active_bases > 9) {
                Y[8] = y * (C3a * yy - C3b * xx);
                Y[9] = C3c * xy * z;
                Y[10] = y * (C3d - C3e * zz);
                Y[11] = z * (C3f * zz - C3g);
                Y[12] = x * (C3d - C3e * zz);
                Y[13] = C3h * z * (xx - yy);
                Y[14] = x * (C3b * yy - C3a * xx);
            }
        }
    }
}

// Deferred Adam update of sh_rest (fgs_backward_adam_ex, FGS_DEFER_SH_REST): launched on a side stream after
// preprocess_backward_kernel, so that its 1080 bytes per Gaussian of pure HBM traffic -- three quarters of the fused
// kernel's -- overlap the latency-bound front of the NEXT forward (geometry preprocess + sorts), which does not read SH
// coefficients when run with deferred colour.  Persistent and deliberately small (one or two 128-thread CTAs per SM, <= 40
// registers): it has to coexist with the forward's kernels, not to win the machine.  Gradient of a visible Gaussian =
// Y_k(dir) * dcol with dir and dcol left in its 2D gradient record by the backward kernel; invisible rows get the
// zero-gradient decay of dense Adam.  Same arithmetic as the fused kernel, same result.
__global__ void __launch_bounds__(128, 4) adam_rest_kernel(const uint32_t* __restrict__ n_touched, const Grad2D* __restrict__ grad2d,
                                                           float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                           const int n, const int row_floats, const int active_bases,
                                                           const float step_size, const float beta1, const float beta2,
                                                           const float eps, const float b2) {
    FGS_DYNAMIC_SMEM(float, s_rows, 16);  // 4 warps x 32 rows x 45 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = s_rows + warp * (32 * 45);
    const int groups = (n + 31) >> 5;
    for (int group = blockIdx.x * 4 + warp; group < groups; group += gridDim.x * 4) {
        const int wfirst = group << 5;
        const int rows = min(32, n - wfirst);
        const int idx = wfirst + lane;
        const bool visible = lane < rows && __ldg(n_touched + idx) > 0u;
        float* mine = stage + lane * row_floats;
        if (visible) {
            const float4 g0 = __ldg(&grad2d[idx].g0);  // (dir.xyz, -)
            const float4 g1 = __ldg(&grad2d[idx].g1);
            const float g2x = __ldg(&grad2d[idx].g2.x);
            float Y[15];
            sh_basis_rest(g0.x, g0.y, g0.z, active_bases, Y);
#pragma unroll
            for (int k = 0; k < 15; ++k) {
                if (3 * k < row_floats) {
                    mine[3 * k] = Y[k] * g1.z;
                    mine[3 * k + 1] = Y[k] * g1.w;
                    mine[3 * k + 2] = Y[k] * g2x;
                }
            }
        } else if (lane < rows) {
            for (int k = 0; k < row_floats; ++k) mine[k] = 0.f;
        }
        __syncwarp();
        emit_span(stage, nullptr, p, m, v, (size_t)wfirst * row_floats, rows * row_floats, true, step_size, beta1, beta2, eps,
                  b2, lane);
        __syncwarp();
    }
}

// Loads of the PARAMETER tensors.  In fused mode (fgs_backward_adam) means / scales / rotations / opacities / sh_rest are
// the very tensors this launch updates in place (emit_span), so they must not go through the read-only data cache
// (ld.global.nc is only defined for memory that stays constant for the whole launch): plain ld.global.
template <typename T>
__device__ __forceinline__ T ld_param(const T* p) {
    return *p;
}

// kSkipInvisibleRest: the sh_rest rows of invisible Gaussians have already been updated by adam_rest_invisible_kernel
// (fused mode only); everything else is identical.
// kCov3d: covariance-input mode (see preprocess.cu): a.scales holds (N,6) covariances, a.g_scales receives their (N,6)
// gradient, rotations are not touched (unfused only).
template <bool kSkipInvisibleRest, bool kCov3d>
__global__ void __launch_bounds__(kBlock) preprocess_backward_kernel(const PreBwdArgs a, const Cam cam) {
    FGS_DYNAMIC_SMEM(float, s_stage, 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* st = s_stage + warp * kWarpFloats;
    const int wfirst = blockIdx.x * kBlock + warp * 32;
    if (wfirst >= a.n) return;
    const int rows = min(32, a.n - wfirst);
    const int idx = wfirst + lane;
    const bool in_range = lane < rows;
    const int nrest = cam.total_sh_rest;
    const int rest_floats = nrest * 3;  // <= 45
    const int active_bases = min(cam.active_sh_bases, nrest + 1);

    const bool visible = in_range && __ldg(a.n_touched + idx) > 0u;
    const uint32_t vis_mask = __ballot_sync(0xffffffffu, visible);

    // stage the SH rest rows of this warp (coalesced) when any Gaussian needs them, else zero them
    float* rest = st + kOffRest;
    if (vis_mask != 0u && active_bases > 1) {
        const float* src = a.sh_rest + (size_t)wfirst * rest_floats;
        const int nf = rows * rest_floats;
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (nf & 3) == 0) {
            for (int i = lane; i < (nf >> 2); i += 32)
                reinterpret_cast<float4*>(rest)[i] = ld_param(reinterpret_cast<const float4*>(src) + i);
        } else {
            for (int i = lane; i < nf; i += 32) rest[i] = ld_param(src + i);
        }
    }
    __syncwarp();

    float gm[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gop = 0.f;
    float gsh0[3] = {0.f, 0.f, 0.f};
    float* my_rest = rest + lane * rest_floats;

    if (visible) {
        const Grad2D* gp = a.grad2d + idx;
        const float4 g0 = ld_param(&gp->g0);  // plain load: the deferred mode overwrites this quarter of the record below
        const float4 g1 = __ldg(&gp->g1);
        const float g2x = __ldg(&gp->g2.x);
        const float3 dcol = make_float3(g1.z, g1.w, g2x);
        const float3 mean = make_float3(ld_param(a.means + 3 * (size_t)idx), ld_param(a.means + 3 * (size_t)idx + 1),
                                        ld_param(a.means + 3 * (size_t)idx + 2));

        // ---- SH backward (sh_utils.cuh:71-155) ----
        gsh0[0] = C0 * dcol.x;
        gsh0[1] = C0 * dcol.y;
        gsh0[2] = C0 * dcol.z;
        float3 dmean_sh = make_float3(0.f, 0.f, 0.f);
        if (active_bases > 1) {
            const float rx = mean.x - __ldg(cam.cam_pos), ry = mean.y - __ldg(cam.cam_pos + 1), rz = mean.z - __ldg(cam.cam_pos + 2);
            const float len2 = rx * rx + ry * ry + rz * rz;
            const float inv_len = rsqrtf(len2);
            const float x = rx * inv_len, y = ry * inv_len, z = rz * inv_len;
            // deferred sh_rest update: adam_rest_kernel rebuilds this Gaussian's sh_rest gradient (Y_k(dir) * dcol) later,
            // from the view direction of THIS forward (the means are updated below, so it cannot be recomputed) --
            // the direction replaces the consumed (dmean2d, dconic.xy) quarter of the 2D gradient record
            if (a.defer_rest) const_cast<Grad2D*>(gp)->g0 = make_float4(x, y, z, 0.f);
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;  // dL/d(direction)
            // coefficient k: reads c_k . dcol for the direction gradient, then overwrites it with Y_k * dcol
#define FGS_SH(k, Y, DX, DY, DZ)                                                             \
    {                                                                                        \
        float* c = my_rest + 3 * (k);                                                        \
        const float cd = c[0] * dcol.x + c[1] * dcol.y + c[2] * dcol.z;                      \
        ddx += (DX) * cd;                                                                    \
        ddy += (DY) * cd;                                                                    \
        ddz += (DZ) * cd;                                                                    \
        const float yk = (Y);                                                                \
        c[0] = yk * dcol.x;                                                                  \
        c[1] = yk * dcol.y;                                                                  \
        c[2] = yk * dcol.z;                                                                  \
    }
            FGS_SH(0, -C1 * y, 0.f, -C1, 0.f)
            FGS_SH(1, C1 * z, 0.f, 0.f, C1)
            FGS_SH(2, -C1 * x, -C1, 0.f, 0.f)
            if (active_bases > 4) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
                FGS_SH(3, C2a * xy, C2a * y, C2a * x, 0.f)
                FGS_SH(4, -C2a * yz, 0.f, -C2a * z, -C2a * y)
                FGS_SH(5, C2b * zz - C2c, 0.f, 0.f, 2.0f * C2b * z)
                FGS_SH(6, -C2a * xz, -C2a * z, 0.f, -C2a * x)
                FGS_SH(7, C2d * (xx - yy), 2.0f * C2d * x, -2.0f * C2d * y, 0.f)
                if (active_bases > 9) {
                    FGS_SH(8, y * (C3a * yy - C3b * xx), -2.0f * C3b * xy, 3.0f * C3a * yy - C3b * xx, 0.f)
                    FGS_SH(9, C3c * xy * z, C3c * yz, C3c * xz, C3c * xy)
                    FGS_SH(10, y * (C3d - C3e * zz), 0.f, C3d - C3e * zz, -2.0f * C3e * yz)
                    FGS_SH(11, z * (C3f * zz - C3g), 0.f, 0.f, 3.0f * C3f * zz - C3g)
                    FGS_SH(12, x * (C3d - C3e * zz), C3d - C3e * zz, 0.f, -2.0f * C3e * xz)
                    FGS_SH(13, C3h * z * (xx - yy), 2.0f * C3h * xz, -2.0f * C3h * yz, C3h * (xx - yy))
                    FGS_SH(14, x * (C3b * yy - C3a * xx), C3b * yy - 3.0f * C3a * xx, 2.0f * C3b * xy, 0.f)
                }
            }
#undef FGS_SH
            // inactive bases get zero gradient
            for (int k = 3 * (active_bases - 1); k < rest_floats; ++k) my_rest[k] = 0.f;
            // through the normalisation: (I - d d^T) / |r| applied to dL/dd
            const float dot_dd = x * ddx + y * ddy + z * ddz;
            dmean_sh = make_float3((ddx - x * dot_dd) * inv_len, (ddy - y * dot_dd) * inv_len,
                                   (ddz - z * dot_dd) * inv_len);
        } else {
            for (int k = 0; k < rest_floats; ++k) my_rest[k] = 0.f;
        }

        // ---- recompute the projection (kernels_backward.cuh:57-117) ----
        float W[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) W[i] = __ldg(cam.w2c + i);
        const float depth = W[8] * mean.x + W[9] * mean.y + W[10] * mean.z + W[11];
        const float inv_z = 1.0f / depth;
        const float x = (W[0] * mean.x + W[1] * mean.y + W[2] * mean.z + W[3]) * inv_z;
        const float y = (W[4] * mean.x + W[5] * mean.y + W[6] * mean.z + W[7]) * inv_z;
        const float3 var = kCov3d ? make_float3(1.f, 1.f, 1.f)
                                  : make_float3(__expf(2.0f * ld_param(a.scales + 3 * (size_t)idx)),
                                                __expf(2.0f * ld_param(a.scales + 3 * (size_t)idx + 1)),
                                                __expf(2.0f * ld_param(a.scales + 3 * (size_t)idx + 2)));
        const float4 q = kCov3d ? make_float4(1.f, 0.f, 0.f, 0.f) : ld_param(reinterpret_cast<const float4*>(a.rotations) + idx);
        const float qxx = q.y * q.y, qyy = q.z * q.z, qzz = q.w * q.w;
        const float qxy = q.y * q.z, qxz = q.y * q.w, qyz = q.z * q.w;
        const float qrx = q.x * q.y, qry = q.x * q.z, qrz = q.x * q.w;
        const float s2 = 2.0f / (q.x * q.x + qxx + qyy + qzz);
        // K(q): R = I + s2 * K
        const float K11 = -(qyy + qzz), K12 = qxy - qrz, K13 = qxz + qry;
        const float K21 = qxy + qrz, K22 = -(qxx + qzz), K23 = qyz - qrx;
        const float K31 = qxz - qry, K32 = qyz + qrx, K33 = -(qxx + qyy);
        const float R11 = 1.0f + s2 * K11, R12 = s2 * K12, R13 = s2 * K13;
        const float R21 = s2 * K21, R22 = 1.0f + s2 * K22, R23 = s2 * K23;
        const float R31 = s2 * K31, R32 = s2 * K32, R33 = 1.0f + s2 * K33;
        // RV = R diag(var)
        const float V11 = R11 * var.x, V12 = R12 * var.y, V13 = R13 * var.z;
        const float V21 = R21 * var.x, V22 = R22 * var.y, V23 = R23 * var.z;
        const float V31 = R31 * var.x, V32 = R32 * var.y, V33 = R33 * var.z;
        float c11 = V11 * R11 + V12 * R12 + V13 * R13;
        float c12 = V11 * R21 + V12 * R22 + V13 * R23;
        float c13 = V11 * R31 + V12 * R32 + V13 * R33;
        float c22 = V21 * R21 + V22 * R22 + V23 * R23;
        float c23 = V21 * R31 + V22 * R32 + V23 * R33;
        float c33 = V31 * R31 + V32 * R32 + V33 * R33;
        if (kCov3d) {
            const float* cv = a.scales + 6 * (size_t)idx;
            c11 = __ldg(cv), c12 = __ldg(cv + 1), c13 = __ldg(cv + 2), c22 = __ldg(cv + 3), c23 = __ldg(cv + 4), c33 = __ldg(cv + 5);
        }

        const float clip_l = (-0.15f * cam.width - cam.center_x) / cam.focal_x;
        const float clip_r = (1.15f * cam.width - cam.center_x) / cam.focal_x;
        const float clip_t = (-0.15f * cam.height - cam.center_y) / cam.focal_y;
        const float clip_b = (1.15f * cam.height - cam.center_y) / cam.focal_y;
        const float xc = fminf(fmaxf(x, clip_l), clip_r);
        const float yc = fminf(fmaxf(y, clip_t), clip_b);
        const float j11 = cam.focal_x * inv_z, j13 = -j11 * xc;
        const float j22 = cam.focal_y * inv_z, j23 = -j22 * yc;
        const float a1 = j11 * W[0] + j13 * W[8], a2 = j11 * W[1] + j13 * W[9], a3 = j11 * W[2] + j13 * W[10];
        const float b1 = j22 * W[4] + j23 * W[8], b2 = j22 * W[5] + j23 * W[9], b3 = j22 * W[6] + j23 * W[10];
        // (JW) Sigma rows
        const float u1 = a1 * c11 + a2 * c12 + a3 * c13, u2 = a1 * c12 + a2 * c22 + a3 * c23,
                    u3 = a1 * c13 + a2 * c23 + a3 * c33;
        const float v1 = b1 * c11 + b2 * c12 + b3 * c13, v2 = b1 * c12 + b2 * c22 + b3 * c23,
                    v3 = b1 * c13 + b2 * c23 + b3 * c33;
        const float a_raw = u1 * a1 + u2 * a2 + u3 * a3;
        const float cb = u1 * b1 + u2 * b2 + u3 * b3;
        const float c_raw = v1 * b1 + v2 * b2 + v3 * b3;
        const float ksz = cam.proper_aa ? kDilationAA : kDilation;
        const float ca = a_raw + ksz, cc = c_raw + ksz;
        const float det = ca * cc - cb * cb;
        const float inv_det = 1.0f / det;
        const float Q11 = cc * inv_det, Q12 = -cb * inv_det, Q22 = ca * inv_det;

        // ---- dL/dSigma2 = -Q G Q, G in the reference's per-entry convention ----
        const float GA = g0.z, GB = g0.w, GC = g1.x;
        const float m11 = GA * Q11 + GB * Q12, m12 = GA * Q12 + GB * Q22;
        const float m21 = GB * Q11 + GC * Q12, m22 = GB * Q12 + GC * Q22;
        const float Dx = -(Q11 * m11 + Q12 * m21);
        const float Dy = -(Q11 * m12 + Q12 * m22);
        const float Dz = -(Q12 * m12 + Q22 * m22);

        // ---- opacity ----
        gop = g1.y;
        if (cam.proper_aa) {
            const float o = sigmoidf_(ld_param(a.opacities + idx));
            const float factor = sqrtf(fmaxf((a_raw * c_raw - cb * cb) * inv_det, 0.0f));
            gop = g1.y * factor * o * (1.0f - o);  // coverage factor detached from cov2d
        }

        // ---- dL/dSigma3 = (JW)^T D (JW) (symmetric, per-entry) ----
        const float S11 = a1 * a1 * Dx + 2.0f * a1 * b1 * Dy + b1 * b1 * Dz;
        const float S12 = a1 * a2 * Dx + (a1 * b2 + a2 * b1) * Dy + b1 * b2 * Dz;
        const float S13 = a1 * a3 * Dx + (a1 * b3 + a3 * b1) * Dy + b1 * b3 * Dz;
        const float S22 = a2 * a2 * Dx + 2.0f * a2 * b2 * Dy + b2 * b2 * Dz;
        const float S23 = a2 * a3 * Dx + (a2 * b3 + a3 * b2) * Dy + b2 * b3 * Dz;
        const float S33 = a3 * a3 * Dx + 2.0f * a3 * b3 * Dy + b3 * b3 * Dz;

        // ---- dL/d(JW rows) and the non-zero Jacobian entries ----
        const float da1 = 2.0f * (u1 * Dx + v1 * Dy), da2 = 2.0f * (u2 * Dx + v2 * Dy), da3 = 2.0f * (u3 * Dx + v3 * Dy);
        const float db1 = 2.0f * (u1 * Dy + v1 * Dz), db2 = 2.0f * (u2 * Dy + v2 * Dz), db3 = 2.0f * (u3 * Dy + v3 * Dz);
        const float dj11 = W[0] * da1 + W[1] * da2 + W[2] * da3;
        const float dj13 = W[8] * da1 + W[9] * da2 + W[10] * da3;
        const float dj22 = W[4] * db1 + W[5] * db2 + W[6] * db3;
        const float dj23 = W[8] * db1 + W[9] * db2 + W[10] * db3;

        // ---- densification statistics (kernels_backward.cuh:194-201) ----
        if (a.densification_info != nullptr) {
            a.densification_info[idx] += 1.0f;
            const float nx = 0.5f * g0.x * cam.width, ny = 0.5f * g0.y * cam.height;
            a.densification_info[a.n + idx] += sqrtf(nx * nx + ny * ny);
        }

        // ---- camera-space gradient of the mean, clip-aware (kernels_backward.cuh:203-220) ----
        const bool valid_x = x >= clip_l && x <= clip_r;
        const bool valid_y = y >= clip_t && y <= clip_b;
        float dcx = j11 * g0.x;
        float dcy = j22 * g0.y;
        float dcz = -j11 * x * g0.x - j22 * y * g0.y;
        if (valid_x) dcx -= j11 * dj13 * inv_z;
        if (valid_y) dcy -= j22 * dj23 * inv_z;
        const float fxm = valid_x ? 2.0f : 1.0f, fym = valid_y ? 2.0f : 1.0f;
        dcz += (j11 * (fxm * xc * dj13 - dj11) + j22 * (fym * yc * dj23 - dj22)) * inv_z;
        gm[0] = W[0] * dcx + W[4] * dcy + W[8] * dcz + dmean_sh.x;
        gm[1] = W[1] * dcx + W[5] * dcy + W[9] * dcz + dmean_sh.y;
        gm[2] = W[2] * dcx + W[6] * dcy + W[10] * dcz + dmean_sh.z;

        // ---- scales: dL/dv_k = r_k^T S r_k, d/d(raw) = 2 v_k ----
        const float dv1 = R11 * (S11 * R11 + S12 * R21 + S13 * R31) + R21 * (S12 * R11 + S22 * R21 + S23 * R31) +
                          R31 * (S13 * R11 + S23 * R21 + S33 * R31);
        const float dv2 = R12 * (S11 * R12 + S12 * R22 + S13 * R32) + R22 * (S12 * R12 + S22 * R22 + S23 * R32) +
                          R32 * (S13 * R12 + S23 * R22 + S33 * R32);
        const float dv3 = R13 * (S11 * R13 + S12 * R23 + S13 * R33) + R23 * (S12 * R13 + S22 * R23 + S23 * R33) +
                          R33 * (S13 * R13 + S23 * R23 + S33 * R33);
        gs[0] = 2.0f * var.x * dv1;
        gs[1] = 2.0f * var.y * dv2;
        gs[2] = 2.0f * var.z * dv3;

        // ---- rotation: dL/dR = 2 S (R diag(v)); R = I + s2 K(q) ----
        const float D11 = 2.0f * (S11 * V11 + S12 * V21 + S13 * V31), D12 = 2.0f * (S11 * V12 + S12 * V22 + S13 * V32),
                    D13 = 2.0f * (S11 * V13 + S12 * V23 + S13 * V33);
        const float D21 = 2.0f * (S12 * V11 + S22 * V21 + S23 * V31), D22 = 2.0f * (S12 * V12 + S22 * V22 + S23 * V32),
                    D23 = 2.0f * (S12 * V13 + S22 * V23 + S23 * V33);
        const float D31 = 2.0f * (S13 * V11 + S23 * V21 + S33 * V31), D32 = 2.0f * (S13 * V12 + S23 * V22 + S33 * V32),
                    D33 = 2.0f * (S13 * V13 + S23 * V23 + S33 * V33);
        const float tK = D11 * K11 + D12 * K12 + D13 * K13 + D21 * K21 + D22 * K22 + D23 * K23 + D31 * K31 +
                         D32 * K32 + D33 * K33;
        const float st_ = s2 * tK;  // contribution through s2 = 2 / |q|^2
        const float er = q.y * (D32 - D23) + q.z * (D13 - D31) + q.w * (D21 - D12);
        const float ex = -2.0f * q.y * (D22 + D33) + q.z * (D12 + D21) + q.w * (D13 + D31) + q.x * (D32 - D23);
        const float ey = -2.0f * q.z * (D11 + D33) + q.y * (D12 + D21) + q.w * (D23 + D32) + q.x * (D13 - D31);
        const float ez = -2.0f * q.w * (D11 + D22) + q.y * (D13 + D31) + q.z * (D23 + D32) + q.x * (D21 - D12);
        gq[0] = s2 * (er - st_ * q.x);
        gq[1] = s2 * (ex - st_ * q.y);
        gq[2] = s2 * (ey - st_ * q.z);
        gq[3] = s2 * (ez - st_ * q.w);
        if (kCov3d) {  // gradient of the six covariance entries: an off-diagonal entry stands for two matrix entries
            gs[0] = S11, gs[1] = 2.0f * S12, gs[2] = 2.0f * S13;
            gq[0] = S22, gq[1] = 2.0f * S23, gq[2] = S33;
        }
    } else if (in_range) {
        for (int k = 0; k < rest_floats; ++k) my_rest[k] = 0.f;
    }

    // ---- stage the small tensors, then emit everything coalesced ----
    if (in_range) {
        st[kOffMeans + 3 * lane] = gm[0];
        st[kOffMeans + 3 * lane + 1] = gm[1];
        st[kOffMeans + 3 * lane + 2] = gm[2];
        if (kCov3d) {  // 32 x 6 floats from kOffScales on (the scales and rotations areas are adjacent)
            st[kOffScales + 6 * lane] = gs[0];
            st[kOffScales + 6 * lane + 1] = gs[1];
            st[kOffScales + 6 * lane + 2] = gs[2];
            st[kOffScales + 6 * lane + 3] = gq[0];
            st[kOffScales + 6 * lane + 4] = gq[1];
            st[kOffScales + 6 * lane + 5] = gq[2];
        } else {
            st[kOffScales + 3 * lane] = gs[0];
            st[kOffScales + 3 * lane + 1] = gs[1];
            st[kOffScales + 3 * lane + 2] = gs[2];
            st[kOffRot + 4 * lane] = gq[0];
            st[kOffRot + 4 * lane + 1] = gq[1];
            st[kOffRot + 4 * lane + 2] = gq[2];
            st[kOffRot + 4 * lane + 3] = gq[3];
        }
        st[kOffOpac + lane] = gop;
        st[kOffSh0 + 3 * lane] = gsh0[0];
        st[kOffSh0 + 3 * lane + 1] = gsh0[1];
        st[kOffSh0 + 3 * lane + 2] = gsh0[2];
    }
    __syncwarp();
    const AdamHyper& h = a.hyper;
    const size_t w0 = (size_t)wfirst;
    // tensor order: 0 means, 1 scales, 2 rotations, 3 opacities, 4 sh0, 5 sh_rest
    emit_span(st + kOffMeans, a.g_means, a.p[0], a.m[0], a.v[0], w0 * 3, rows * 3, a.fused, h.step_size[0], h.beta1,
              h.beta2, h.eps, h.bias2_sqrt_rcp[0], lane);
    if (kCov3d) {
        emit_span(st + kOffScales, a.g_scales, nullptr, nullptr, nullptr, w0 * 6, rows * 6, false, 0.f, 0.f, 0.f, 0.f, 0.f, lane);
    } else {
        emit_span(st + kOffScales, a.g_scales, a.p[1], a.m[1], a.v[1], w0 * 3, rows * 3, a.fused, h.step_size[1], h.beta1,
                  h.beta2, h.eps, h.bias2_sqrt_rcp[1], lane);
        emit_span(st + kOffRot, a.g_rotations, a.p[2], a.m[2], a.v[2], w0 * 4, rows * 4, a.fused, h.step_size[2], h.beta1,
                  h.beta2, h.eps, h.bias2_sqrt_rcp[2], lane);
    }
    emit_span(st + kOffOpac, a.g_opacities, a.p[3], a.m[3], a.v[3], w0, rows, a.fused, h.step_size[3], h.beta1, h.beta2,
              h.eps, h.bias2_sqrt_rcp[3], lane);
    emit_span(st + kOffSh0, a.g_sh0, a.p[4], a.m[4], a.v[4], w0 * 3, rows * 3, a.fused, h.step_size[4], h.beta1, h.beta2,
              h.eps, h.bias2_sqrt_rcp[4], lane);
    if (rest_floats > 0 && !a.defer_rest && (a.fused || a.g_sh_rest != nullptr)) {
        if (kSkipInvisibleRest) {
            if (vis_mask != 0u)
                emit_span_rows<4>(rest, a.p[5], a.m[5], a.v[5], w0 * rest_floats, rows * rest_floats, rest_floats, vis_mask,
                                  h.step_size[5], h.beta1, h.beta2, h.eps, h.bias2_sqrt_rcp[5], lane);
        } else {
            emit_span(rest, a.g_sh_rest, a.p[5], a.m[5], a.v[5], w0 * rest_floats, rows * rest_floats, a.fused,
                      h.step_size[5], h.beta1, h.beta2, h.eps, h.bias2_sqrt_rcp[5], lane);
        }
    }
}

__global__ void __launch_bounds__(256) adam_kernel(const float* __restrict__ grad, float* __restrict__ param,
                                                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                   const size_t n, const size_t n4, const float step_size,
                                                   const float beta1, const float beta2, const float eps,
                                                   const float b2) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad) + i);
        float4 p4 = reinterpret_cast<float4*>(param)[i];
        float4 m4 = reinterpret_cast<float4*>(exp_avg)[i];
        float4 v4 = reinterpret_cast<float4*>(exp_avg_sq)[i];
        adam_update(g4.x, p4.x, m4.x, v4.x, step_size, beta1, beta2, eps, b2);
        adam_update(g4.y, p4.y, m4.y, v4.y, step_size, beta1, beta2, eps, b2);
        adam_update(g4.z, p4.z, m4.z, v4.z, step_size, beta1, beta2, eps, b2);
        adam_update(g4.w, p4.w, m4.w, v4.w, step_size, beta1, beta2, eps, b2);
        reinterpret_cast<float4*>(param)[i] = p4;
        reinterpret_cast<float4*>(exp_avg)[i] = m4;
        reinterpret_cast<float4*>(exp_avg_sq)[i] = v4;
    }
    // scalar tail (and the whole tensor when it is not 16-byte aligned: n4 == 0)
    const size_t t = n4 * 4 + i;
    if (i < n - n4 * 4 && t < n) {
        float p = param[t], m = exp_avg[t], v = exp_avg_sq[t];
        adam_update(grad[t], p, m, v, step_size, beta1, beta2, eps, b2);
        param[t] = p;
        exp_avg[t] = m;
        exp_avg_sq[t] = v;
    }
}

}  // namespace

cudaError_t launch_preprocess_backward(const PreBwdArgs& a, const Cam& cam, cudaStream_t stream, bool skip_invisible_rest) {
    if (a.n <= 0) return cudaSuccess;
    static PerDevice<bool> s_attr_dev;
    bool& s_attr = s_attr_dev.here();
    const size_t smem = (size_t)kWarps * kWarpFloats * sizeof(float);
    if (!s_attr) {
        auto k00 = preprocess_backward_kernel<false, false>;
        auto k10 = preprocess_backward_kernel<true, false>;
        auto k01 = preprocess_backward_kernel<false, true>;
        cudaError_t e = cudaFuncSetAttribute(k00, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k10, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k01, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        s_attr = true;
    }
    if (a.cov3d && a.fused) return cudaErrorInvalidValue;  // the covariance-input mode has no fused Adam
    auto kernel = a.cov3d ? preprocess_backward_kernel<false, true>
                          : (skip_invisible_rest && a.fused) ? preprocess_backward_kernel<true, false> : preprocess_backward_kernel<false, false>;
    FGS_LAUNCH_SMEM(kernel, div_up(a.n, kBlock), kBlock, smem, stream, a, cam);
    return cudaGetLastError();
}

bool preprocess_backward_can_defer_rest(const PreBwdArgs& a, const Cam& cam) {
    const int rest_floats = cam.total_sh_rest * 3;
    return a.fused && !a.cov3d && a.n > 0 && rest_floats >= 3 && rest_floats <= 45 && cam.active_sh_bases > 1;
}

cudaError_t launch_adam_rest(const PreBwdArgs& a, const Cam& cam, cudaStream_t stream, int ctas_per_sm) {
    if (!preprocess_backward_can_defer_rest(a, cam)) return cudaErrorInvalidValue;
    static PerDevice<int> s_sms_dev;
    int& s_sms = s_sms_dev.here();
    if (s_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev);
        if (s_sms < 1) s_sms = 1;
    }
    const int groups = (a.n + 31) / 32;
    const int want = (groups + 3) / 4;
    const int cap = s_sms * (ctas_per_sm > 0 ? ctas_per_sm : 2);
    const int grid = want < cap ? want : cap;
    const AdamHyper& h = a.hyper;
    const size_t smem = 4 * 32 * 45 * sizeof(float);
    const int rest_floats = cam.total_sh_rest * 3;
    const int active_bases = cam.active_sh_bases < cam.total_sh_rest + 1 ? cam.active_sh_bases : cam.total_sh_rest + 1;
    FGS_LAUNCH_SMEM(adam_rest_kernel, grid, 128, smem, stream, a.n_touched, a.grad2d, a.p[5], a.m[5], a.v[5], a.n, rest_floats,
                    active_bases, h.step_size[5], h.beta1, h.beta2, h.eps, h.bias2_sqrt_rcp[5]);
    return cudaGetLastError();
}

bool preprocess_backward_can_overlap(const PreBwdArgs& a, const Cam& cam) {
    const int rest_floats = cam.total_sh_rest * 3;
    return a.fused && a.n > 0 && rest_floats >= 9 && rest_floats <= 45;
}

cudaError_t launch_adam_rest_invisible(const PreBwdArgs& a, const Cam& cam, cudaStream_t stream) {
    if (!preprocess_backward_can_overlap(a, cam)) return cudaSuccess;
    static PerDevice<int> s_sms_dev;
    int& s_sms = s_sms_dev.here();
    if (s_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev);
        if (s_sms < 1) s_sms = 1;
    }
    const int groups = (a.n + 31) / 32;
    const int grid = (groups + 3) / 4 < s_sms ? (groups + 3) / 4 : s_sms;  // at most one CTA per SM
    const AdamHyper& h = a.hyper;
    FGS_LAUNCH(adam_rest_invisible_kernel, grid, 128, stream, a.n_touched, a.p[5], a.m[5], a.v[5], a.n, cam.total_sh_rest * 3,
               h.step_size[5], h.beta1, h.beta2, h.eps, h.bias2_sqrt_rcp[5]);
    return cudaGetLastError();
}

cudaError_t launch_adam(const float* grad, float* param, float* exp_avg, float* exp_avg_sq, size_t n, float step_size,
                        float beta1, float beta2, float eps, float bias2_sqrt_rcp, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const bool aligned = ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(param) |
                           reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0;
    const size_t n4 = aligned ? n / 4 : 0;
    const size_t tail = n - n4 * 4;
    const size_t threads = n4 > tail ? n4 : tail;
    FGS_LAUNCH_SERIAL(adam_kernel, (unsigned)((threads + 255) / 256), 256, stream, grad, param, exp_avg, exp_avg_sq, n, n4, step_size,
                                                                      beta1, beta2, eps, bias2_sqrt_rcp);
    return cudaGetLastError();
}

}  // namespace fgs
