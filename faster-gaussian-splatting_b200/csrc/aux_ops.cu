// aux_ops.cu -- the three small per-Gaussian operators that complete the reference's `_C` surface
// around the rasterizer: Mip-Splatting 3D filter update and the two 3DGS-MCMC
// helpers.  All three are one-thread-per-Gaussian streaming kernels.
//
// Semantics:
//   update_3d_filter ....... /root/reference/.../filter3d/src/filter3d.cu:9-86
//   relocation_adjustment .. .../densification/include/kernels_mcmc.cuh:10-57 (Eq. 9 of 3DGS-MCMC)
//   add_noise .............. .../densification/include/kernels_mcmc.cuh:69-127
#include <cmath>

#include "fgs_device.cuh"
#include "fgs_host.h"

namespace fgs {

namespace {

constexpr int kMaxSamples = 50;  // densification_config.h:9
constexpr int kAuxBlock = 256;

// c[n][k] = (-1)^k C(n, k) / sqrt(k + 1)
__constant__ float c_reloc[kMaxSamples * kMaxSamples];

__global__ void __launch_bounds__(kAuxBlock) filter3d_kernel(const float* __restrict__ positions,
                                                             const float* __restrict__ w2c, float* __restrict__ filter_3d,
                                                             unsigned char* __restrict__ visibility, const int n,
                                                             const float left, const float right, const float top,
                                                             const float bottom, const float near_plane,
                                                             const float distance2filter) {
    const int i = blockIdx.x * kAuxBlock + threadIdx.x;
    if (i >= n) return;
    const float px = __ldg(positions + 3 * (size_t)i), py = __ldg(positions + 3 * (size_t)i + 1),
                pz = __ldg(positions + 3 * (size_t)i + 2);
    const float z = __ldg(w2c + 8) * px + __ldg(w2c + 9) * py + __ldg(w2c + 10) * pz + __ldg(w2c + 11);
    if (z < near_plane) return;
    const float xc = __ldg(w2c + 0) * px + __ldg(w2c + 1) * py + __ldg(w2c + 2) * pz + __ldg(w2c + 3);
    if (xc < left * z || xc > right * z) return;  // inside the dilated frustum, tested without dividing by z
    const float yc = __ldg(w2c + 4) * px + __ldg(w2c + 5) * py + __ldg(w2c + 6) * pz + __ldg(w2c + 7);
    if (yc < top * z || yc > bottom * z) return;
    const float candidate = distance2filter * z;
    if (filter_3d[i] < candidate) return;  // running minimum over the views
    filter_3d[i] = candidate;
    visibility[i] = 1;
}

__global__ void __launch_bounds__(kAuxBlock) relocation_kernel(const float* __restrict__ old_opacities,
                                                               const float* __restrict__ old_scales,
                                                               const long long* __restrict__ n_samples,
                                                               float* __restrict__ new_opacities,
                                                               float* __restrict__ new_scales, const int n) {
    const int i = blockIdx.x * kAuxBlock + threadIdx.x;
    if (i >= n) return;
    const float o = __ldg(old_opacities + i);
    const int ns = min(max((int)__ldg(n_samples + i), 1), kMaxSamples);
    // opacity that, composited ns times, reproduces the old one
    const float o_new = 1.0f - powf(1.0f - o, 1.0f / (float)ns);
    new_opacities[i] = o_new;
    // sum_{m < ns} sum_{k <= m} c[m][k] o_new^(k+1), summed in this order (alternating series)
    float denom = 0.0f;
    for (int m = 0; m < ns; ++m) {
        float pw = o_new;
        for (int k = 0; k <= m; ++k, pw *= o_new) denom += c_reloc[m * kMaxSamples + k] * pw;
    }
    const float f = o / denom;
    new_scales[3 * (size_t)i] = f * __ldg(old_scales + 3 * (size_t)i);
    new_scales[3 * (size_t)i + 1] = f * __ldg(old_scales + 3 * (size_t)i + 1);
    new_scales[3 * (size_t)i + 2] = f * __ldg(old_scales + 3 * (size_t)i + 2);
}

__global__ void __launch_bounds__(kAuxBlock) add_noise_kernel(const float* __restrict__ raw_scales,
                                                              const float* __restrict__ raw_rotations,
                                                              const float* __restrict__ raw_opacities,
                                                              const float* __restrict__ noise, float* __restrict__ means,
                                                              const int n, const float lr) {
    const int i = blockIdx.x * kAuxBlock + threadIdx.x;
    if (i >= n) return;
    const float4 q = __ldg(reinterpret_cast<const float4*>(raw_rotations) + i);
    const float xx = q.y * q.y, yy = q.z * q.z, zz = q.w * q.w;
    const float nsq = q.x * q.x + xx + yy + zz;
    if (nsq < 1e-8f) return;
    const float s2 = 2.0f / nsq;
    const float xy = q.y * q.z, xz = q.y * q.w, yz = q.z * q.w, rx = q.x * q.y, ry = q.x * q.z, rz = q.x * q.w;
    const float R11 = 1.0f - (yy + zz) * s2, R12 = (xy - rz) * s2, R13 = (xz + ry) * s2;
    const float R21 = (xy + rz) * s2, R22 = 1.0f - (xx + zz) * s2, R23 = (yz - rx) * s2;
    const float R31 = (xz - ry) * s2, R32 = (yz + rx) * s2, R33 = 1.0f - (xx + yy) * s2;
    const float v1 = __expf(2.0f * __ldg(raw_scales + 3 * (size_t)i)), v2 = __expf(2.0f * __ldg(raw_scales + 3 * (size_t)i + 1)),
                v3 = __expf(2.0f * __ldg(raw_scales + 3 * (size_t)i + 2));
    const float e1 = __ldg(noise + 3 * (size_t)i), e2 = __ldg(noise + 3 * (size_t)i + 1), e3 = __ldg(noise + 3 * (size_t)i + 2);
    // Sigma eps = R diag(v) R^T eps, evaluated right to left (no 3x3 covariance is formed)
    const float t1 = v1 * (R11 * e1 + R21 * e2 + R31 * e3);
    const float t2 = v2 * (R12 * e1 + R22 * e2 + R32 * e3);
    const float t3 = v3 * (R13 * e1 + R23 * e2 + R33 * e3);
    const float o = sigmoidf_(__ldg(raw_opacities + i));
    const float gate = 1.0f / (1.0f + __expf(100.0f * o - 0.5f));  // ~1 for transparent, ~0 for opaque Gaussians
    const float f = lr * gate;
    means[3 * (size_t)i] += f * (R11 * t1 + R12 * t2 + R13 * t3);
    means[3 * (size_t)i + 1] += f * (R21 * t1 + R22 * t2 + R23 * t3);
    means[3 * (size_t)i + 2] += f * (R31 * t1 + R32 * t2 + R33 * t3);
}

PerDevice<bool> g_reloc_ready;  // the __constant__ table is per device

}  // namespace

cudaError_t launch_update_3d_filter(const float* positions, const float* w2c, float* filter_3d, unsigned char* visibility,
                                    int n, int width, int height, float focal_x, float focal_y, float center_x,
                                    float center_y, float near_plane, float clipping_tolerance, float distance2filter,
                                    cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    // frustum dilated by `clipping_tolerance` of the image size on every side (filter3d.cu:56-66)
    const float bf = clipping_tolerance + 0.5f;
    const float wf = (float)width, hf = (float)height;
    const float ox = center_x - 0.5f * wf, oy = center_y - 0.5f * hf;
    const float left = (-bf * wf - ox) / focal_x, right = (bf * wf - ox) / focal_x;
    const float top = (-bf * hf - oy) / focal_y, bottom = (bf * hf - oy) / focal_y;
    FGS_LAUNCH_SERIAL(filter3d_kernel, div_up(n, kAuxBlock), kAuxBlock, stream, positions, w2c, filter_3d, visibility, n, left, right,
                                                                  top, bottom, near_plane, distance2filter);
    return cudaGetLastError();
}

cudaError_t launch_relocation(const float* old_opacities, const float* old_scales, const long long* n_samples,
                              float* new_opacities, float* new_scales, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (!g_reloc_ready.here()) {
        static float table[kMaxSamples * kMaxSamples];
        for (int m = 0; m < kMaxSamples; ++m) {
            double binom = 1.0, sign = 1.0;
            for (int k = 0; k <= m; ++k, sign = -sign) {
                table[m * kMaxSamples + k] = (float)(binom * sign / std::sqrt((double)(k + 1)));
                binom *= (double)(m - k) / (double)(k + 1);
            }
        }
        cudaError_t e = cudaMemcpyToSymbolAsync(c_reloc, table, sizeof(table), 0, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        g_reloc_ready.here() = true;
    }
    FGS_LAUNCH_SERIAL(relocation_kernel, div_up(n, kAuxBlock), kAuxBlock, stream, old_opacities, old_scales, n_samples, new_opacities,
                                                                    new_scales, n);
    return cudaGetLastError();
}

cudaError_t launch_add_noise(const float* raw_scales, const float* raw_rotations, const float* raw_opacities,
                             const float* noise, float* means, int n, float lr, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    FGS_LAUNCH_SERIAL(add_noise_kernel, div_up(n, kAuxBlock), kAuxBlock, stream, raw_scales, raw_rotations, raw_opacities, noise, means,
                                                                   n, lr);
    return cudaGetLastError();
}

}  // namespace fgs
