// slice4d.cu -- time slice of 4D Gaussians and its backward (BASELINE.json configs[4]).
//
// The reference's FasterGS4D branch (README.md:39) is not in the tree; it follows Yang et al., "Real-time Photorealistic
// Dynamic Scene Representation and Rendering with 4D Gaussian Splatting" (ICLR 2024).  Restated from the paper (parity
// unpinned, oracle/fourd_oracle.py): a 4D Gaussian = mean (x, y, z, t), four log-scales, two unnormalised quaternions;
//     R = L(q_l / |q_l|) Rr(q_r / |q_r|)      left- times right-isoclinic rotation
//     Sigma4 = M M^T, M = R diag(exp(s))
//     mean3 = mu_xyz + b (t - mu_t) / c,   cov3 = A - b b^T / c,   with A = Sigma4[:3,:3], b = Sigma4[:3,3], c = Sigma4[3,3]
//     opacity = sigmoid(o) exp(-(t - mu_t)^2 / (2 c))
// The slice is an ordinary 3D Gaussian given by its covariance: it is rendered by the rasterizer's covariance-input mode
// (fgs_forward with rotations == NULL), which takes opacity LOGITS, so the kernel emits logit(opacity).  The backward
// kernel is the hand-written reverse of the same arithmetic (no intermediate is stored: 25 floats in, 10 out per
// Gaussian forward; one thread per Gaussian, streaming).
//
// This file is also compiled by g++ against tests/cuda_emu/cuda_emu.h (test-only emulator) -- see fgs_launch.h.
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "../../include/fgs_b200.h"
#include "fgs_launch.h"

namespace fgs {
namespace {

constexpr int kSliceBlock = 128;
constexpr float kMaxOpacity = 0.99999994f;  // largest float below 1: keeps logit finite

struct Slice4 {
    float ql[4], qr[4];    // normalised quaternions
    float nl, nr;          // 1 / |q_l|, 1 / |q_r|
    float L[4][4], Rr[4][4], R[4][4], S[4], M[4][4];
    float A[3][3], b[3], c, dt;
};

__device__ __forceinline__ void build_slice(const float* __restrict__ means4, const float* __restrict__ scales4,
                                            const float* __restrict__ rot_l, const float* __restrict__ rot_r, const size_t i,
                                            const float t, Slice4& s) {
    const float4 l = __ldg(reinterpret_cast<const float4*>(rot_l) + i), r = __ldg(reinterpret_cast<const float4*>(rot_r) + i);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scales4) + i);
    s.nl = 1.0f / sqrtf(l.x * l.x + l.y * l.y + l.z * l.z + l.w * l.w);
    s.nr = 1.0f / sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    s.ql[0] = l.x * s.nl, s.ql[1] = l.y * s.nl, s.ql[2] = l.z * s.nl, s.ql[3] = l.w * s.nl;
    s.qr[0] = r.x * s.nr, s.qr[1] = r.y * s.nr, s.qr[2] = r.z * s.nr, s.qr[3] = r.w * s.nr;
    const float a = s.ql[0], b = s.ql[1], c = s.ql[2], d = s.ql[3];
    const float p = s.qr[0], q = s.qr[1], u = s.qr[2], v = s.qr[3];
    const float L[4][4] = {{a, -b, -c, -d}, {b, a, -d, c}, {c, d, a, -b}, {d, -c, b, a}};
    const float Rr[4][4] = {{p, q, u, v}, {-q, p, -v, u}, {-u, v, p, -q}, {-v, -u, q, p}};
    s.S[0] = expf(sc.x), s.S[1] = expf(sc.y), s.S[2] = expf(sc.z), s.S[3] = expf(sc.w);
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            s.L[m][n] = L[m][n];
            s.Rr[m][n] = Rr[m][n];
        }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fmaf(L[m][k], Rr[k][n], acc);
            s.R[m][n] = acc;
            s.M[m][n] = acc * s.S[n];
        }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int n = 0; n < 3; ++n) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fmaf(s.M[m][k], s.M[n][k], acc);
            s.A[m][n] = acc;
        }
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(s.M[m][k], s.M[3][k], acc);
        s.b[m] = acc;
    }
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fmaf(s.M[3][k], s.M[3][k], acc);
    s.c = acc;
    s.dt = t - __ldg(means4 + 4 * i + 3);
}

__global__ void __launch_bounds__(kSliceBlock) slice4d_forward_kernel(const float* __restrict__ means4, const float* __restrict__ scales4,
                                                                      const float* __restrict__ rot_l, const float* __restrict__ rot_r,
                                                                      const float* __restrict__ opacities, const float t, const int n,
                                                                      float* __restrict__ means3, float* __restrict__ cov3,
                                                                      float* __restrict__ logits) {
    const int i = blockIdx.x * kSliceBlock + threadIdx.x;
    if (i >= n) return;
    Slice4 s;
    build_slice(means4, scales4, rot_l, rot_r, (size_t)i, t, s);
    const float rc = 1.0f / s.c;
    const float k = s.dt * rc;
    means3[3 * (size_t)i] = __ldg(means4 + 4 * (size_t)i) + s.b[0] * k;
    means3[3 * (size_t)i + 1] = __ldg(means4 + 4 * (size_t)i + 1) + s.b[1] * k;
    means3[3 * (size_t)i + 2] = __ldg(means4 + 4 * (size_t)i + 2) + s.b[2] * k;
    float* cv = cov3 + 6 * (size_t)i;
    cv[0] = s.A[0][0] - s.b[0] * s.b[0] * rc;
    cv[1] = s.A[0][1] - s.b[0] * s.b[1] * rc;
    cv[2] = s.A[0][2] - s.b[0] * s.b[2] * rc;
    cv[3] = s.A[1][1] - s.b[1] * s.b[1] * rc;
    cv[4] = s.A[1][2] - s.b[1] * s.b[2] * rc;
    cv[5] = s.A[2][2] - s.b[2] * s.b[2] * rc;
    const float sig = 1.0f / (1.0f + expf(-__ldg(opacities + i)));
    const float o = fminf(sig * expf(-0.5f * s.dt * s.dt * rc), kMaxOpacity);
    logits[i] = logf(o) - log1pf(-o);  // o == 0: -inf, the rasterizer culls it (sigmoid(-inf) = 0 < 1/255)
}

__global__ void __launch_bounds__(kSliceBlock) slice4d_backward_kernel(
    const float* __restrict__ means4, const float* __restrict__ scales4, const float* __restrict__ rot_l,
    const float* __restrict__ rot_r, const float* __restrict__ opacities, const float t, const int n,
    const float* __restrict__ g_means3, const float* __restrict__ g_cov3, const float* __restrict__ g_logits,
    float* __restrict__ g_means4, float* __restrict__ g_scales4, float* __restrict__ g_rot_l, float* __restrict__ g_rot_r,
    float* __restrict__ g_opacities) {
    const int i = blockIdx.x * kSliceBlock + threadIdx.x;
    if (i >= n) return;
    Slice4 s;
    build_slice(means4, scales4, rot_l, rot_r, (size_t)i, t, s);
    const float rc = 1.0f / s.c, dt = s.dt;
    const float gm[3] = {__ldg(g_means3 + 3 * (size_t)i), __ldg(g_means3 + 3 * (size_t)i + 1), __ldg(g_means3 + 3 * (size_t)i + 2)};
    const float* g6 = g_cov3 + 6 * (size_t)i;
    // dL/dcov3 as a full symmetric matrix: an off-diagonal entry of the six stands for two matrix entries
    const float D[3][3] = {{__ldg(g6), 0.5f * __ldg(g6 + 1), 0.5f * __ldg(g6 + 2)},
                           {0.5f * __ldg(g6 + 1), __ldg(g6 + 3), 0.5f * __ldg(g6 + 4)},
                           {0.5f * __ldg(g6 + 2), 0.5f * __ldg(g6 + 4), __ldg(g6 + 5)}};
    // opacity = sigmoid(o) * marginal, logit = log(opacity / (1 - opacity))
    const float sig = 1.0f / (1.0f + expf(-__ldg(opacities + i)));
    const float marg = expf(-0.5f * dt * dt * rc);
    const float o = sig * marg;
    float g_o = 0.0f;  // dL/d(opacity); 0 where the forward clamped or underflowed (such Gaussians are culled or saturated)
    if (o > 1e-30f && o < kMaxOpacity) g_o = __ldg(g_logits + i) / (o * (1.0f - o));
    const float g_marg = g_o * sig;
    g_opacities[i] = g_o * marg * sig * (1.0f - sig);

    // cov3 = A - b b^T / c,  mean3 = mu + b dt / c,  marginal = exp(-dt^2 / (2 c))
    float Db[3], gb[3];
    float bDb = 0.0f, gmb = 0.0f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        Db[m] = D[m][0] * s.b[0] + D[m][1] * s.b[1] + D[m][2] * s.b[2];
        bDb = fmaf(s.b[m], Db[m], bDb);
        gmb = fmaf(gm[m], s.b[m], gmb);
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) gb[m] = (-2.0f * Db[m] + gm[m] * dt) * rc;
    const float gc = (bDb - gmb * dt + g_marg * marg * 0.5f * dt * dt) * rc * rc;
    const float gdt = (gmb - g_marg * marg * dt) * rc;
    g_means4[4 * (size_t)i] = gm[0];
    g_means4[4 * (size_t)i + 1] = gm[1];
    g_means4[4 * (size_t)i + 2] = gm[2];
    g_means4[4 * (size_t)i + 3] = -gdt;

    // Sigma4 = M M^T: dL/dM = 2 E M with E the full symmetric gradient of Sigma4
    float E[4][4];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int k = 0; k < 3; ++k) E[m][k] = D[m][k];
        E[m][3] = E[3][m] = 0.5f * gb[m];
    }
    E[3][3] = gc;
    float gM[4][4], gR[4][4], gS[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = fmaf(E[m][j], s.M[j][k], acc);
            gM[m][k] = 2.0f * acc;
        }
    // M = R diag(S)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            gS[k] = fmaf(gM[m][k], s.R[m][k], gS[k]);
            gR[m][k] = gM[m][k] * s.S[k];
        }
    }
    reinterpret_cast<float4*>(g_scales4)[i] = make_float4(gS[0] * s.S[0], gS[1] * s.S[1], gS[2] * s.S[2], gS[3] * s.S[3]);
    // R = L Rr: dL/dL = gR Rr^T, dL/dRr = L^T gR
    float gL[4][4], gRr[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float accl = 0.0f, accr = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                accl = fmaf(gR[m][j], s.Rr[k][j], accl);
                accr = fmaf(s.L[j][m], gR[j][k], accr);
            }
            gL[m][k] = accl;
            gRr[m][k] = accr;
        }
    // the isoclinic matrices are linear in the (normalised) quaternion components
    float gql[4], gqr[4];
    gql[0] = gL[0][0] + gL[1][1] + gL[2][2] + gL[3][3];
    gql[1] = -gL[0][1] + gL[1][0] - gL[2][3] + gL[3][2];
    gql[2] = -gL[0][2] + gL[1][3] + gL[2][0] - gL[3][1];
    gql[3] = -gL[0][3] - gL[1][2] + gL[2][1] + gL[3][0];
    gqr[0] = gRr[0][0] + gRr[1][1] + gRr[2][2] + gRr[3][3];
    gqr[1] = gRr[0][1] - gRr[1][0] - gRr[2][3] + gRr[3][2];
    gqr[2] = gRr[0][2] + gRr[1][3] - gRr[2][0] - gRr[3][1];
    gqr[3] = gRr[0][3] - gRr[1][2] + gRr[2][1] - gRr[3][0];
    // through q / |q|: (g - q_hat (q_hat . g)) / |q|
    const float dl = s.ql[0] * gql[0] + s.ql[1] * gql[1] + s.ql[2] * gql[2] + s.ql[3] * gql[3];
    const float dr = s.qr[0] * gqr[0] + s.qr[1] * gqr[1] + s.qr[2] * gqr[2] + s.qr[3] * gqr[3];
    reinterpret_cast<float4*>(g_rot_l)[i] = make_float4((gql[0] - s.ql[0] * dl) * s.nl, (gql[1] - s.ql[1] * dl) * s.nl,
                                                        (gql[2] - s.ql[2] * dl) * s.nl, (gql[3] - s.ql[3] * dl) * s.nl);
    reinterpret_cast<float4*>(g_rot_r)[i] = make_float4((gqr[0] - s.qr[0] * dr) * s.nr, (gqr[1] - s.qr[1] * dr) * s.nr,
                                                        (gqr[2] - s.qr[2] * dr) * s.nr, (gqr[3] - s.qr[3] * dr) * s.nr);
}

bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

}  // namespace
}  // namespace fgs

extern "C" {

int fgs_slice4d_forward(const float* means4, const float* scales4, const float* rot_l, const float* rot_r, const float* opacities,
                        float t, int32_t n_primitives, float* means3, float* cov3, float* opacity_logits, void* stream) {
    if (n_primitives <= 0) return 0;
    if (!means4 || !scales4 || !rot_l || !rot_r || !opacities || !means3 || !cov3 || !opacity_logits)
        return fgs::report_error(-1, "fgs_slice4d_forward", "NULL argument");
    if (fgs::misaligned16(scales4) || fgs::misaligned16(rot_l) || fgs::misaligned16(rot_r))
        return fgs::report_error(-1, "fgs_slice4d_forward", "scales4 / rot_l / rot_r must be 16-byte aligned");
    auto k = fgs::slice4d_forward_kernel;
    FGS_LAUNCH_SERIAL(k, dim3((n_primitives + fgs::kSliceBlock - 1) / fgs::kSliceBlock), dim3(fgs::kSliceBlock),
                      static_cast<cudaStream_t>(stream), means4, scales4, rot_l, rot_r, opacities, t, (int)n_primitives, means3, cov3,
                      opacity_logits);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fgs::report_error((int)e, "fgs_slice4d_forward", cudaGetErrorString(e));
}

int fgs_slice4d_backward(const float* means4, const float* scales4, const float* rot_l, const float* rot_r, const float* opacities,
                         float t, int32_t n_primitives, const float* grad_means3, const float* grad_cov3,
                         const float* grad_opacity_logits, float* grad_means4, float* grad_scales4, float* grad_rot_l,
                         float* grad_rot_r, float* grad_opacities, void* stream) {
    if (n_primitives <= 0) return 0;
    if (!means4 || !scales4 || !rot_l || !rot_r || !opacities || !grad_means3 || !grad_cov3 || !grad_opacity_logits || !grad_means4 ||
        !grad_scales4 || !grad_rot_l || !grad_rot_r || !grad_opacities)
        return fgs::report_error(-1, "fgs_slice4d_backward", "NULL argument");
    if (fgs::misaligned16(scales4) || fgs::misaligned16(rot_l) || fgs::misaligned16(rot_r) || fgs::misaligned16(grad_scales4) ||
        fgs::misaligned16(grad_rot_l) || fgs::misaligned16(grad_rot_r))
        return fgs::report_error(-1, "fgs_slice4d_backward", "(N,4) tensors must be 16-byte aligned");
    auto k = fgs::slice4d_backward_kernel;
    FGS_LAUNCH_SERIAL(k, dim3((n_primitives + fgs::kSliceBlock - 1) / fgs::kSliceBlock), dim3(fgs::kSliceBlock),
                      static_cast<cudaStream_t>(stream), means4, scales4, rot_l, rot_r, opacities, t, (int)n_primitives, grad_means3,
                      grad_cov3, grad_opacity_logits, grad_means4, grad_scales4, grad_rot_l, grad_rot_r, grad_opacities);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fgs::report_error((int)e, "fgs_slice4d_backward", cudaGetErrorString(e));
}

}  // extern "C"
