// preprocess.cu -- per-Gaussian projection kernel (world -> screen, 3D -> 2D covariance / conic,
// opacity-aware bounds, exact tile count, SH -> RGB) for sm_100a.
//
// Semantics: /root/reference/.../rasterization/include/kernels_forward.cuh:14-209 (P1-P12 in
// SURVEY.md §8.1).  New here: AoS parameter rows are staged through shared memory with 16-byte
// coalesced loads (the (N,3)/(N,15,3) layouts are fixed by the API), one packed 48-byte record per
// Gaussian instead of five scratch arrays, warp-aggregated compaction, counts stay on the device.
#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

constexpr int kPreBlock = 128;
constexpr int kSeqTiles = 16;  // candidate tiles tested by the owning thread before going warp-cooperative

// sh_utils.cuh:9-30
constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.48860251190291987f;
constexpr float SH_C2a = 1.0925484305920792f, SH_C2b = 0.94617469575755997f, SH_C2c = 0.31539156525251999f,
                SH_C2d = 0.54627421529603959f;
constexpr float SH_C3a = 0.59004358992664352f, SH_C3b = 1.7701307697799304f, SH_C3c = 2.8906114426405538f,
                SH_C3d = 0.45704579946446572f, SH_C3e = 2.2852289973223288f, SH_C3f = 1.865881662950577f,
                SH_C3g = 1.1195289977703462f, SH_C3h = 1.4453057213202769f;

// Stage `rows` consecutive rows of K floats (a contiguous span) into shared memory with 16-byte
// loads when the span start is 16-byte aligned.
template <int K>
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int n_floats,
                                           int tid, int nthreads) {
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int n4 = n_floats >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n4; i += nthreads) d4[i] = __ldg(s4 + i);
        for (int i = (n4 << 2) + tid; i < n_floats; i += nthreads) dst[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < n_floats; i += nthreads) dst[i] = __ldg(src + i);
    }
}

// sh_utils.cuh:32-69; `c` points at this Gaussian's rest coefficients (stride 1 float, rgb interleaved).
__device__ __forceinline__ float3 eval_sh(const float3 sh0, const float* __restrict__ c, float3 dir, int active) {
    float3 r = make_float3(0.5f + SH_C0 * sh0.x, 0.5f + SH_C0 * sh0.y, 0.5f + SH_C0 * sh0.z);
    if (active > 1) {
        const float x = dir.x, y = dir.y, z = dir.z;
#define FGS_ACC(k, w)                 \
    {                                 \
        const float w_ = (w);         \
        r.x += w_ * c[3 * (k) + 0];   \
        r.y += w_ * c[3 * (k) + 1];   \
        r.z += w_ * c[3 * (k) + 2];   \
    }
        FGS_ACC(0, -SH_C1 * y)
        FGS_ACC(1, SH_C1 * z)
        FGS_ACC(2, -SH_C1 * x)
        if (active > 4) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            FGS_ACC(3, SH_C2a * xy)
            FGS_ACC(4, -SH_C2a * yz)
            FGS_ACC(5, SH_C2b * zz - SH_C2c)
            FGS_ACC(6, -SH_C2a * xz)
            FGS_ACC(7, SH_C2d * (xx - yy))
            if (active > 9) {
                FGS_ACC(8, y * (SH_C3a * yy - SH_C3b * xx))
                FGS_ACC(9, SH_C3c * xy * z)
                FGS_ACC(10, y * (SH_C3d - SH_C3e * zz))
                FGS_ACC(11, z * (SH_C3f * zz - SH_C3g))
                FGS_ACC(12, x * (SH_C3d - SH_C3e * zz))
                FGS_ACC(13, SH_C3h * z * (xx - yy))
                FGS_ACC(14, x * (SH_C3b * yy - SH_C3a * xx))
            }
        }
#undef FGS_ACC
    }
    return r;
}

// kCov3d: `scales` points at (N,6) rows (xx, xy, xz, yy, yz, zz) of the 3D covariance and `rotations` is not read -- the
// input mode of Gaussians whose covariance does not come from (scale, rotation), e.g. 4D Gaussians sliced at a time t
// (csrc/slice4d.cu).  Everything downstream of the covariance is identical.
// kDeferColor: geometry only.  sh0 / sh_rest are NOT read and the record's colour fields stay zero; color_kernel fills a
// separate colour array later (after the sorts), so that whatever still writes the SH coefficients -- the deferred Adam
// update of sh_rest on a side stream, or their broadcast from a peer GPU -- can overlap this kernel and the sorts.
template <bool kInference, bool kCov3d, bool kDeferColor>
__global__ void __launch_bounds__(kPreBlock) preprocess_kernel(
    const float* __restrict__ means, const float* __restrict__ scales, const float* __restrict__ rotations,
    const float* __restrict__ opacities, const float* __restrict__ sh0, const float* __restrict__ sh_rest,
    Rec* __restrict__ recs, uint32_t* __restrict__ n_touched_out, uint2* __restrict__ tile_info_out,
    uint32_t* __restrict__ depth_keys,
    uint32_t* __restrict__ vis_indices, uint32_t* __restrict__ counters, uint32_t* __restrict__ depth_hist, const int n,
    const Cam cam) {
    __shared__ __align__(16) float s_means[kPreBlock * 3];
    __shared__ __align__(16) float s_scales[kPreBlock * (kCov3d ? 6 : 3)];
    __shared__ __align__(16) float s_sh0[kPreBlock * 3];
    FGS_DYNAMIC_SMEM(float, s_rest, 16);  // 4 warps x 32 x 45 floats (only used when total_sh_rest == 15)
    __shared__ __align__(16) float s_w2c[12];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int first = blockIdx.x * kPreBlock;
    const int rows = min(kPreBlock, n - first);
    // camera rows: one 48-byte bulk async copy (TMA) when the matrix is 16-byte aligned
    const bool cam_bulk = (reinterpret_cast<uintptr_t>(cam.w2c) & 15u) == 0;
    if (cam_bulk) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&s_bar, 48u);
            bulk_g2s(s_w2c, cam.w2c, 48u, &s_bar);
        }
    } else if (tid < 12) {
        s_w2c[tid] = __ldg(cam.w2c + tid);
    }
    // this thread's opacity logit and quaternion: issued before the staging barrier so that their latency
    // overlaps the staged loads (measured: two exposed DRAM round trips after the barrier otherwise)
    const int idx = first + tid;
    const float raw_opacity = __ldg(opacities + (idx < n ? idx : first));
    const float4 q = kCov3d ? make_float4(1.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(rotations) + (idx < n ? idx : first));
    stage_rows<3>(s_means, means + (size_t)first * 3, rows * 3, tid, kPreBlock);
    if (kCov3d)
        stage_rows<6>(s_scales, scales + (size_t)first * 6, rows * 6, tid, kPreBlock);
    else
        stage_rows<3>(s_scales, scales + (size_t)first * 3, rows * 3, tid, kPreBlock);
    if (!kDeferColor) stage_rows<3>(s_sh0, sh0 + (size_t)first * 3, rows * 3, tid, kPreBlock);
    __syncthreads();
    if (cam_bulk) mbar_wait(&s_bar, 0u);
    float W[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) W[i] = s_w2c[i];

    bool active = idx < n;
    const int r = active ? tid : 0;  // inactive tail threads mirror row 0 and stay alive for warp collectives

    const float3 mean = make_float3(s_means[3 * r], s_means[3 * r + 1], s_means[3 * r + 2]);
    // P1: depth and near/far cull
    const float depth = W[8] * mean.x + W[9] * mean.y + W[10] * mean.z + W[11];
    if (depth < cam.near_plane || depth > cam.far_plane) active = false;

    // P2: activated opacity, cull below 1/255
    float opacity = sigmoidf_(raw_opacity);
    if (opacity < kMinAlpha) active = false;

    // P3: covariance from exp(2 s) and the unnormalised quaternion (r, x, y, z)
    const float3 var = kCov3d ? make_float3(1.f, 1.f, 1.f)
                              : make_float3(__expf(2.0f * s_scales[3 * r]), __expf(2.0f * s_scales[3 * r + 1]),
                                            __expf(2.0f * s_scales[3 * r + 2]));
    const float qxx = q.y * q.y, qyy = q.z * q.z, qzz = q.w * q.w;
    const float qxy = q.y * q.z, qxz = q.y * q.w, qyz = q.z * q.w;
    const float qrx = q.x * q.y, qry = q.x * q.z, qrz = q.x * q.w;
    const float norm_sq = q.x * q.x + qxx + qyy + qzz;
    if (norm_sq < 1e-8f) active = false;
    const float ns = 2.0f / norm_sq;
    const float R11 = 1.0f - (qyy + qzz) * ns, R12 = (qxy - qrz) * ns, R13 = (qxz + qry) * ns;
    const float R21 = (qxy + qrz) * ns, R22 = 1.0f - (qxx + qzz) * ns, R23 = (qyz - qrx) * ns;
    const float R31 = (qxz - qry) * ns, R32 = (qyz + qrx) * ns, R33 = 1.0f - (qxx + qyy) * ns;
    const float S11 = R11 * var.x, S12 = R12 * var.y, S13 = R13 * var.z;
    const float S21 = R21 * var.x, S22 = R22 * var.y, S23 = R23 * var.z;
    const float S31 = R31 * var.x, S32 = R32 * var.y, S33 = R33 * var.z;
    float c11 = S11 * R11 + S12 * R12 + S13 * R13;
    float c12 = S11 * R21 + S12 * R22 + S13 * R23;
    float c13 = S11 * R31 + S12 * R32 + S13 * R33;
    float c22 = S21 * R21 + S22 * R22 + S23 * R23;
    float c23 = S21 * R31 + S22 * R32 + S23 * R33;
    float c33 = S31 * R31 + S32 * R32 + S33 * R33;
    if (kCov3d) {
        c11 = s_scales[6 * r], c12 = s_scales[6 * r + 1], c13 = s_scales[6 * r + 2];
        c22 = s_scales[6 * r + 3], c23 = s_scales[6 * r + 4], c33 = s_scales[6 * r + 5];
    }

    // P4: normalised image coordinates, clipped copy for the Jacobian only
    const float x = (W[0] * mean.x + W[1] * mean.y + W[2] * mean.z + W[3]) / depth;
    const float y = (W[4] * mean.x + W[5] * mean.y + W[6] * mean.z + W[7]) / depth;
    const float clip_l = (-0.15f * cam.width - cam.center_x) / cam.focal_x;
    const float clip_r = (1.15f * cam.width - cam.center_x) / cam.focal_x;
    const float clip_t = (-0.15f * cam.height - cam.center_y) / cam.focal_y;
    const float clip_b = (1.15f * cam.height - cam.center_y) / cam.focal_y;
    const float xc = fminf(fmaxf(x, clip_l), clip_r);
    const float yc = fminf(fmaxf(y, clip_t), clip_b);
    const float j11 = cam.focal_x / depth, j13 = -j11 * xc;
    const float j22 = cam.focal_y / depth, j23 = -j22 * yc;
    const float a1 = j11 * W[0] + j13 * W[8];
    const float a2 = j11 * W[1] + j13 * W[9];
    const float a3 = j11 * W[2] + j13 * W[10];
    const float b1 = j22 * W[4] + j23 * W[8];
    const float b2 = j22 * W[5] + j23 * W[9];
    const float b3 = j22 * W[6] + j23 * W[10];
    const float u1 = a1 * c11 + a2 * c12 + a3 * c13;
    const float u2 = a1 * c12 + a2 * c22 + a3 * c23;
    const float u3 = a1 * c13 + a2 * c23 + a3 * c33;
    const float v1 = b1 * c11 + b2 * c12 + b3 * c13;
    const float v2 = b1 * c12 + b2 * c22 + b3 * c23;
    const float v3 = b1 * c13 + b2 * c23 + b3 * c33;
    // P5: 2D covariance, dilation, conic
    float cov_a = u1 * a1 + u2 * a2 + u3 * a3;
    const float cov_b = u1 * b1 + u2 * b2 + u3 * b3;
    float cov_c = v1 * b1 + v2 * b2 + v3 * b3;
    const float det_raw = cov_a * cov_c - cov_b * cov_b;
    const float ksz = cam.proper_aa ? kDilationAA : kDilation;
    cov_a += ksz;
    cov_c += ksz;
    const float det = cov_a * cov_c - cov_b * cov_b;
    if (det < kMinCov2dDet) active = false;
    const float conic_a = cov_c / det, conic_b = -cov_b / det, conic_c = cov_a / det;
    if (cam.proper_aa) {
        opacity *= sqrtf(fmaxf(det_raw / det, 0.0f));
        if (opacity < kMinAlpha) active = false;
    }
    const float m2x = x * cam.focal_x + cam.center_x;
    const float m2y = y * cam.focal_y + cam.center_y;

    // P6/P7: opacity-aware extent, bounds clamped to the padded tile grid
    const float lop = __log2f(opacity);
    const float tau2 = lop + kLog2_255;  // log2(255 * opacity) = tau * log2(e)
    const float tau = tau2 * 0.6931471805599453f;
    const float ext_x = fmaxf(sqrtf(cov_a * 2.0f * tau) - 0.5f, 0.0f);
    const float ext_y = fmaxf(sqrtf(cov_c * 2.0f * tau) - 0.5f, 0.0f);
    // conic in the base-2 exponent domain (see fgs_device.cuh): everything downstream uses these
    const float A2 = -0.5f * kLog2e * conic_a, B2 = -kLog2e * conic_b, C2 = -0.5f * kLog2e * conic_c;
    const int pw = cam.grid_w * kTileW, ph = cam.grid_h * kTileH;
    const uint32_t xmin = (uint32_t)min(pw, max(0, __float2int_rd(m2x - ext_x)));
    const uint32_t xmax = (uint32_t)min(pw, max(0, __float2int_ru(m2x + ext_x)));
    const uint32_t ymin = (uint32_t)min(ph, max(0, __float2int_rd(m2y - ext_y)));
    const uint32_t ymax = (uint32_t)min(ph, max(0, __float2int_ru(m2y + ext_y)));
    uint32_t tx0, tx1, ty0, ty1;
    tile_bounds_of(xmin, xmax, ymin, ymax, tx0, tx1, ty0, ty1);
    const uint32_t tw = tx1 - tx0;
    const uint32_t cand = active ? tw * (ty1 - ty0) : 0u;
    if (cand == 0) active = false;

    // P9 prefetch: the 180-byte SH rest rows of the warp (32 consecutive rows = one contiguous 5760-byte span) start
    // streaming into shared memory now, asynchronously (cp.async, SASS LDGSTS), and are waited for only when the
    // colour is evaluated, after the exact tile count below
    const int nrest = cam.total_sh_rest;
    const int active_bases = min(cam.active_sh_bases, nrest + 1);
    float* my_rest = s_rest + warp * (32 * 45);
    bool rest_staged = false;
    if (!kDeferColor) {
        const int wfirst = first + warp * 32;
        const int wrows = max(0, min(32, n - wfirst));
        const float* wsrc = sh_rest + (size_t)wfirst * 45;
        const uint32_t cand_mask = __ballot_sync(0xffffffffu, active);
        if (nrest == 15 && active_bases > 4 && __popc(cand_mask) >= 8 && (reinterpret_cast<uintptr_t>(wsrc) & 15u) == 0) {
            const int n_floats = wrows * 45, n4 = n_floats >> 2;
            for (int i = lane; i < n4; i += 32) cp_async_16(my_rest + 4 * i, wsrc + 4 * i);
            cp_async_commit();
            for (int i = (n4 << 2) + lane; i < n_floats; i += 32) my_rest[i] = __ldg(wsrc + i);
            rest_staged = true;
        }
    }

    // P8: exact number of touched tiles
    const float msx = m2x - 0.5f, msy = m2y - 0.5f;
    uint32_t n_touched = 0;
    uint32_t tile_mask = 0;  // bit c: candidate tile c of the rectangle is touched (kept for rectangles of <= 32 tiles)
    if (active && cand <= (uint32_t)kSeqTiles) {
        for (uint32_t c = 0; c < cand; ++c) {
            const uint32_t tx = tx0 + c % tw, ty = ty0 + c / tw;
            tile_mask |= (tile_contributes2(msx, msy, A2, B2, C2, tx, ty, tau2) ? 1u : 0u) << c;
        }
        n_touched = __popc(tile_mask);
    }
    uint32_t big = __ballot_sync(0xffffffffu, active && cand > (uint32_t)kSeqTiles);
    while (big) {
        const int src = __ffs((int)big) - 1;
        big &= big - 1;
        const float bmx = __shfl_sync(0xffffffffu, msx, src), bmy = __shfl_sync(0xffffffffu, msy, src);
        const float bA = __shfl_sync(0xffffffffu, A2, src), bB = __shfl_sync(0xffffffffu, B2, src);
        const float bC = __shfl_sync(0xffffffffu, C2, src), btau = __shfl_sync(0xffffffffu, tau2, src);
        const uint32_t bx0 = __shfl_sync(0xffffffffu, tx0, src), by0 = __shfl_sync(0xffffffffu, ty0, src);
        const uint32_t btw = __shfl_sync(0xffffffffu, tw, src), bcand = __shfl_sync(0xffffffffu, cand, src);
        uint32_t cnt = 0, first_chunk = 0;
        for (uint32_t c0 = 0; c0 < bcand; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool ok = c < bcand && tile_contributes2(bmx, bmy, bA, bB, bC, bx0 + c % btw, by0 + c / btw, btau);
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (c0 == 0) first_chunk = m;
            cnt += __popc(m);
        }
        if (lane == src) {
            n_touched = cnt;
            tile_mask = first_chunk;
        }
    }
    const bool visible = active && n_touched > 0;
    if (idx < n) n_touched_out[idx] = visible ? n_touched : 0u;
    if (visible) tile_info_out[idx] = pack_tile_info(tx0, ty0, tw, cand, tile_mask, n_touched);

    // P10/P11 (first half): reserve the warp's slots in the (depth key, index) list -- one atomic per warp, issued
    // here so that its round trip overlaps the colour evaluation; order inside the warp is preserved
    const uint32_t vis_mask = __ballot_sync(0xffffffffu, visible);
    uint32_t base = 0;
    if (vis_mask != 0) {
        const int leader = __ffs((int)vis_mask) - 1;
        const uint32_t wsum = __reduce_add_sync(0xffffffffu, visible ? n_touched : 0u);
        if (lane == leader) {
            base = atomicAdd(&counters[kCtrVisible], (uint32_t)__popc(vis_mask));
            atomicAdd(&counters[kCtrInstances], wsum);
        }
    }
    // P9: SH -> RGB for visible Gaussians
    if (rest_staged) {
        cp_async_wait_all();
        __syncwarp();
    }
    if (vis_mask == 0) return;
    const float* rest_ptr = nullptr;
    if (active_bases > 1)
        rest_ptr = rest_staged ? my_rest + lane * 45 : sh_rest + (size_t)(idx < n ? idx : first) * (size_t)(nrest * 3);
    base = __shfl_sync(0xffffffffu, base, __ffs((int)vis_mask) - 1);
    if (!visible) return;
    float3 color = make_float3(0.f, 0.f, 0.f);
    if (!kDeferColor) {
        float3 dir = make_float3(mean.x - __ldg(cam.cam_pos), mean.y - __ldg(cam.cam_pos + 1), mean.z - __ldg(cam.cam_pos + 2));
        const float inv_len = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        dir.x *= inv_len;
        dir.y *= inv_len;
        dir.z *= inv_len;
        color = eval_sh(make_float3(s_sh0[3 * r], s_sh0[3 * r + 1], s_sh0[3 * r + 2]), rest_ptr, dir, active_bases);
    }
    // colours are stored clamped (what every blend consumes, kernels_forward.cuh:430 /
    // kernels_inference.cuh:200); the clamp mask for the backward rides in the sign bit
    Rec rec;
    rec.q0 = make_float4(m2x, m2y, A2, B2);
    rec.q1 = make_float4(C2, lop, pack_color(color.x), pack_color(color.y));
    rec.q2 = make_float4(pack_color(color.z), __uint_as_float(xmin | (xmax << 16)), __uint_as_float(ymin | (ymax << 16)),
                         __uint_as_float((uint32_t)idx));
    recs[idx] = rec;

    // P10/P11 (second half): append (depth key, index)
    const uint32_t slot = base + __popc(vis_mask & ((1u << lane) - 1u));
    depth_keys[slot] = __float_as_uint(depth);
    vis_indices[slot] = (uint32_t)idx;
    // digit histogram of the depth sort's first pass (low 8 key bits are mantissa noise: no aggregation to
    // be had inside a block, so the fire-and-forget atomics are spread over kDepthHistReplicas copies)
    atomicAdd(depth_hist + (blockIdx.x % kDepthHistReplicas) * kRadixMax + (__float_as_uint(depth) & 0xffu), 1u);
}

// The colour of every visible Gaussian (n_touched > 0), evaluated after the fact for a forward that ran
// preprocess_kernel<.., kDeferColor = true>: same staging (cp.async of the warp's 32 SH rows), same expressions, same
// clamped / sign-tagged encoding as the record's colour fields; colors[idx] = (r, g, b, 0).
__global__ void __launch_bounds__(kPreBlock) color_kernel(const float* __restrict__ means, const float* __restrict__ sh0,
                                                          const float* __restrict__ sh_rest,
                                                          const uint32_t* __restrict__ n_touched, float4* __restrict__ colors,
                                                          const int n, const Cam cam) {
    FGS_DYNAMIC_SMEM(float, s_rest, 16);  // 4 warps x 32 x 45 floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wfirst = blockIdx.x * kPreBlock + warp * 32;
    if (wfirst >= n) return;
    const int idx = wfirst + lane;
    const bool visible = idx < n && __ldg(n_touched + idx) > 0u;
    const uint32_t vis_mask = __ballot_sync(0xffffffffu, visible);
    if (vis_mask == 0u) return;
    const int nrest = cam.total_sh_rest;
    const int active_bases = min(cam.active_sh_bases, nrest + 1);
    float* my_rest = s_rest + warp * (32 * 45);
    bool rest_staged = false;
    {
        const int wrows = min(32, n - wfirst);
        const float* wsrc = sh_rest + (size_t)wfirst * 45;
        if (nrest == 15 && active_bases > 4 && __popc(vis_mask) >= 8 && (reinterpret_cast<uintptr_t>(wsrc) & 15u) == 0) {
            const int n_floats = wrows * 45, n4 = n_floats >> 2;
            for (int i = lane; i < n4; i += 32) cp_async_16(my_rest + 4 * i, wsrc + 4 * i);
            cp_async_commit();
            for (int i = (n4 << 2) + lane; i < n_floats; i += 32) my_rest[i] = __ldg(wsrc + i);
            rest_staged = true;
        }
    }
    float3 mean = make_float3(0.f, 0.f, 0.f), c0 = mean;
    if (visible) {
        mean = make_float3(__ldg(means + 3 * (size_t)idx), __ldg(means + 3 * (size_t)idx + 1), __ldg(means + 3 * (size_t)idx + 2));
        c0 = make_float3(__ldg(sh0 + 3 * (size_t)idx), __ldg(sh0 + 3 * (size_t)idx + 1), __ldg(sh0 + 3 * (size_t)idx + 2));
    }
    if (rest_staged) {
        cp_async_wait_all();
        __syncwarp();
    }
    if (!visible) return;
    const float* rest_ptr = nullptr;
    if (active_bases > 1) rest_ptr = rest_staged ? my_rest + lane * 45 : sh_rest + (size_t)idx * (size_t)(nrest * 3);
    float3 dir = make_float3(mean.x - __ldg(cam.cam_pos), mean.y - __ldg(cam.cam_pos + 1), mean.z - __ldg(cam.cam_pos + 2));
    const float inv_len = rsqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x *= inv_len;
    dir.y *= inv_len;
    dir.z *= inv_len;
    const float3 color = eval_sh(c0, rest_ptr, dir, active_bases);
    colors[idx] = make_float4(pack_color(color.x), pack_color(color.y), pack_color(color.z), 0.f);
}

cudaError_t launch_color(const float* means, const float* sh0, const float* sh_rest, const uint32_t* n_touched,
                         float4* colors, int n, const Cam& cam, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int grid = (n + kPreBlock - 1) / kPreBlock;
    const size_t smem = (size_t)(kPreBlock / 32) * 32 * 45 * sizeof(float);
    FGS_LAUNCH_SMEM(color_kernel, grid, kPreBlock, smem, stream, means, sh0, sh_rest, n_touched, colors, n, cam);
    return cudaGetLastError();
}

cudaError_t launch_preprocess(const float* means, const float* scales, const float* rotations,
                              const float* opacities, const float* sh0, const float* sh_rest, Rec* recs,
                              uint32_t* n_touched, uint2* tile_info, uint32_t* depth_keys, uint32_t* vis_indices, uint32_t* counters,
                              uint32_t* depth_hist, int n, const Cam& cam, bool inference, cudaStream_t stream,
                              const float* cov3d, bool defer_color) {
    if (n <= 0) return cudaSuccess;
    const int grid = (n + kPreBlock - 1) / kPreBlock;
    const size_t smem = (size_t)(kPreBlock / 32) * 32 * 45 * sizeof(float);
#define FGS_PRE_LAUNCH(INF, COV, DEFER, SCALES)                                                                        \
    do {                                                                                                               \
        auto kernel = preprocess_kernel<INF, COV, DEFER>;                                                              \
        FGS_LAUNCH_SMEM(kernel, grid, kPreBlock, DEFER ? 0 : smem, stream, means, SCALES, rotations, opacities, sh0,    \
                        sh_rest, recs, n_touched, tile_info, depth_keys, vis_indices, counters, depth_hist, n, cam);   \
    } while (0)
    if (cov3d) {
        if (defer_color) return cudaErrorInvalidValue;  // the covariance-input mode has no deferred colour
        if (inference) FGS_PRE_LAUNCH(true, true, false, cov3d);
        else FGS_PRE_LAUNCH(false, true, false, cov3d);
    } else if (defer_color) {
        if (inference) FGS_PRE_LAUNCH(true, false, true, scales);
        else FGS_PRE_LAUNCH(false, false, true, scales);
    } else {
        if (inference) FGS_PRE_LAUNCH(true, false, false, scales);
        else FGS_PRE_LAUNCH(false, false, false, scales);
    }
#undef FGS_PRE_LAUNCH
    return cudaGetLastError();
}

}  // namespace fgs
