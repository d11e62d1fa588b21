// loss.cu -- photometric loss of the training step fused with its gradient.
//
// Reference semantics: /root/reference/Loss.py:15-16,26-33 -- loss = LAMBDA_L1 * l1_loss(image, target)
// + LAMBDA_DSSIM * fused_dssim(image, target) on the (3,H,W) image the rasterizer returns, called once per training
// iteration between the forward and loss.backward() (Trainer.py:191-196).  `fused_dssim` lives in NeRFICG
// (Optim.Losses.DSSIM, not in the reference tree, no version pinned): restated here from the published fused-ssim
// algorithm it wraps -- SSIM with an 11x11 Gaussian window (sigma 1.5, separable), zero padding, C1 = 0.01^2,
// C2 = 0.03^2, per channel, mean over all 3*H*W values; DSSIM = 1 - mean SSIM (3DGS convention).  Parity of the DSSIM
// term is therefore pinned by oracle/loss_oracle.py (torch conv2d + autograd), not by the reference: "parity unpinned".
//
// Kernels (HBM-bound by design; x = image, y = target, n = 3*H*W):
//   l1_kernel ............ lambda_dssim == 0: one streaming pass, float4 / uchar4 loads, sums |x-y| and (optionally)
//                          writes the gradient sign(x-y) * lambda_l1 / n.  Algorithmic bytes: 4n + (4|1)n in, 4n out.
//   ssim_stats_kernel .... per (32x32 tile, channel): stages the 42x42 halo tiles of x and y in shared memory,
//                          separable 11-tap convolution of (x, y, x^2, y^2, xy) with a register sliding window,
//                          SSIM and its three partial derivatives per pixel; writes the derivative maps, sums SSIM and
//                          |x-y|.  Algorithmic bytes: (4 + 4|1)n in, 12n out.
//   ssim_grad_kernel ..... same tiling over the three derivative maps: dL/dx = conv(dm_dmu1) + 2x conv(dm_dsigma1^2)
//                          + y conv(dm_dsigma12), fused with the L1 term and the incoming dL/dloss (read from the
//                          device: no host sync).  Algorithmic bytes: 12n + (4 + 4|1)n in, 4n out.
// The last block to finish (ticket) turns the two double-precision sums into the loss scalars: no extra launch.
// The target may be float32 or uint8 (value / 255, as image datasets store it): the uint8 -> float pass of the
// trainer disappears.
//
// This file is also compiled by g++ against tests/cuda_emu/cuda_emu.h (test-only emulator) -- see fgs_launch.h.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "../../include/fgs_b200.h"
#include "fgs_launch.h"

namespace fgs {
namespace {

constexpr int kLossTile = 32;                                // output pixels per CTA side
constexpr int kSsimRadius = 5;
constexpr int kSsimTaps = 2 * kSsimRadius + 1;               // 11
constexpr int kHalo = kLossTile + 2 * kSsimRadius;           // 42
constexpr int kHaloStride = 43;                              // odd: the horizontal pass reads conflict-free
constexpr int kRowStride = 33;                               // row stride of the horizontally filtered tiles
constexpr int kLossBlock = 256;
constexpr int kHorizOut = 8;                                 // outputs per thread in the horizontal pass
constexpr int kVertOut = 4;                                  // outputs per thread in the vertical pass
constexpr float kC1 = 0.01f * 0.01f;
constexpr float kC2 = 0.03f * 0.03f;
static_assert(kHalo * (kLossTile / kHorizOut) <= kLossBlock, "horizontal pass: one item per thread");
static_assert(kLossTile * (kLossTile / kVertOut) == kLossBlock, "vertical pass: one item per thread");

struct SsimWindow {
    float w[kSsimTaps];
};

struct LossHeader {  // first 128 bytes of the scratch buffer, zeroed before every forward
    double sum_abs;   // sum |x - y|
    double sum_ssim;  // sum of the SSIM map
    unsigned int ticket;
    unsigned int pad[27];
};
static_assert(sizeof(LossHeader) == 128, "LossHeader must be 128 bytes");

// uint8 targets: value * (1/255), which is what torch's CUDA `t.float().div(255.0)` evaluates (division by a host
// scalar is multiplication by its reciprocal there)
constexpr float kU8Scale = 1.0f / 255.0f;

template <typename TgtT>
__device__ __forceinline__ float load_target(const TgtT* __restrict__ p, size_t i);
template <>
__device__ __forceinline__ float load_target<float>(const float* __restrict__ p, size_t i) {
    return __ldg(p + i);
}
template <>
__device__ __forceinline__ float load_target<unsigned char>(const unsigned char* __restrict__ p, size_t i) {
    return (float)__ldg(p + i) * kU8Scale;
}

// four consecutive target values (index 4q .. 4q+3) with one 16-byte / 4-byte load
template <typename TgtT>
__device__ __forceinline__ float4 load_target4(const TgtT* __restrict__ p, size_t q);
template <>
__device__ __forceinline__ float4 load_target4<float>(const float* __restrict__ p, size_t q) {
    return __ldg(reinterpret_cast<const float4*>(p) + q);
}
template <>
__device__ __forceinline__ float4 load_target4<unsigned char>(const unsigned char* __restrict__ p, size_t q) {
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p) + q);
    return make_float4((float)v.x * kU8Scale, (float)v.y * kU8Scale, (float)v.z * kU8Scale, (float)v.w * kU8Scale);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sign_of(float d) { return (float)(d > 0.0f) - (float)(d < 0.0f); }

// Adds the block's two partial sums to the header; the last block to arrive writes the loss scalars:
// loss_out = (lambda_l1 * L1 + lambda_dssim * DSSIM, L1, DSSIM).  Called by thread 0 only.
__device__ __forceinline__ void publish_sums(LossHeader* __restrict__ hdr, float* __restrict__ loss_out, double sum_abs,
                                             double sum_ssim, unsigned n_blocks, double n_values, float lambda_l1,
                                             float lambda_dssim, bool has_ssim) {
    atomicAdd(&hdr->sum_abs, sum_abs);
    if (has_ssim) atomicAdd(&hdr->sum_ssim, sum_ssim);
    __threadfence();
    const unsigned t = atomicAdd(&hdr->ticket, 1u);
    if (t == n_blocks - 1u) {
        __threadfence();
        const double ta = atomicAdd(&hdr->sum_abs, 0.0), ts = atomicAdd(&hdr->sum_ssim, 0.0);
        const float l1 = (float)(ta / n_values);
        const float ds = has_ssim ? (float)(1.0 - ts / n_values) : 0.0f;
        loss_out[0] = lambda_l1 * l1 + lambda_dssim * ds;
        loss_out[1] = l1;
        loss_out[2] = ds;
    }
}

// ---- lambda_dssim == 0 -----------------------------------------------------------------------------------
template <typename TgtT, bool kSum, bool kGrad>
__global__ void __launch_bounds__(kLossBlock) l1_kernel(const float* __restrict__ image, const TgtT* __restrict__ target,
                                                        float* __restrict__ grad_image, LossHeader* __restrict__ hdr,
                                                        float* __restrict__ loss_out, const float* __restrict__ grad_loss,
                                                        const size_t n_values, const float lambda_l1) {
    __shared__ float s_red[kLossBlock / 32];
    const size_t n_quads = n_values / 4;
    const float gs = kGrad ? (grad_loss ? __ldg(grad_loss) : 1.0f) * lambda_l1 / (float)n_values : 0.0f;
    float local = 0.0f;
    for (size_t q = (size_t)blockIdx.x * kLossBlock + threadIdx.x; q < n_quads; q += (size_t)gridDim.x * kLossBlock) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(image) + q);
        const float4 y = load_target4<TgtT>(target, q);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        if (kSum) local += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
        if (kGrad)
            reinterpret_cast<float4*>(grad_image)[q] = make_float4(gs * sign_of(d0), gs * sign_of(d1), gs * sign_of(d2), gs * sign_of(d3));
    }
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n_values - 4 * n_quads)) {  // at most 3 trailing values
        const size_t i = 4 * n_quads + threadIdx.x;
        const float d = __ldg(image + i) - load_target<TgtT>(target, i);
        if (kSum) local += fabsf(d);
        if (kGrad) grad_image[i] = gs * sign_of(d);
    }
    if (kSum) {
        local = warp_sum(local);
        if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int w = 0; w < kLossBlock / 32; ++w) a += (double)s_red[w];
            publish_sums(hdr, loss_out, a, 0.0, gridDim.x, (double)n_values, lambda_l1, 0.0f, false);
        }
    }
}

// ---- separable 11-tap convolution of a staged halo tile ---------------------------------------------------
// Horizontal pass of one thread: row `r` of the halo tile, kHorizOut consecutive outputs starting at column
// kHorizOut * g; every staged value is loaded once and pushed into the outputs whose window covers it.
template <int kMaps, typename Expand>
__device__ __forceinline__ void conv_rows(const SsimWindow& win, float (*__restrict__ s_h)[kHalo][kRowStride], const int r,
                                          const int g, Expand expand) {
    float acc[kMaps][kHorizOut];
#pragma unroll
    for (int m = 0; m < kMaps; ++m)
#pragma unroll
        for (int o = 0; o < kHorizOut; ++o) acc[m][o] = 0.0f;
#pragma unroll
    for (int j = 0; j < kHorizOut + kSsimTaps - 1; ++j) {
        float v[kMaps];
        expand(r, kHorizOut * g + j, v);
#pragma unroll
        for (int o = 0; o < kHorizOut; ++o) {
            const int k = j - o;
            if (k >= 0 && k < kSsimTaps) {
#pragma unroll
                for (int m = 0; m < kMaps; ++m) acc[m][o] = fmaf(win.w[k], v[m], acc[m][o]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < kMaps; ++m)
#pragma unroll
        for (int o = 0; o < kHorizOut; ++o) s_h[m][r][kHorizOut * g + o] = acc[m][o];
}

// Vertical pass of one thread: column `c`, kVertOut consecutive output rows starting at kVertOut * rg.
template <int kMaps>
__device__ __forceinline__ void conv_cols(const SsimWindow& win, const float (*__restrict__ s_h)[kHalo][kRowStride], const int c,
                                          const int rg, float (&acc)[kMaps][kVertOut]) {
#pragma unroll
    for (int m = 0; m < kMaps; ++m)
#pragma unroll
        for (int o = 0; o < kVertOut; ++o) acc[m][o] = 0.0f;
#pragma unroll
    for (int j = 0; j < kVertOut + kSsimTaps - 1; ++j) {
        float v[kMaps];
#pragma unroll
        for (int m = 0; m < kMaps; ++m) v[m] = s_h[m][kVertOut * rg + j][c];
#pragma unroll
        for (int o = 0; o < kVertOut; ++o) {
            const int k = j - o;
            if (k >= 0 && k < kSsimTaps) {
#pragma unroll
                for (int m = 0; m < kMaps; ++m) acc[m][o] = fmaf(win.w[k], v[m], acc[m][o]);
            }
        }
    }
}

// ---- SSIM statistics, SSIM map sum, derivative maps ---------------------------------------------------------
template <typename TgtT>
__global__ void __launch_bounds__(kLossBlock) ssim_stats_kernel(const float* __restrict__ image, const TgtT* __restrict__ target,
                                                                float* __restrict__ maps, LossHeader* __restrict__ hdr,
                                                                float* __restrict__ loss_out, const int width, const int height,
                                                                const SsimWindow win, const float lambda_l1,
                                                                const float lambda_dssim) {
    __shared__ float s_x[kHalo * kHaloStride];
    __shared__ float s_y[kHalo * kHaloStride];
    __shared__ float s_h[5][kHalo][kRowStride];
    __shared__ float s_red[2][kLossBlock / 32];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kLossTile, y0 = blockIdx.y * kLossTile;
    const size_t plane_px = (size_t)width * height;
    const size_t plane = (size_t)blockIdx.z * plane_px;

    for (int i = tid; i < kHalo * kHalo; i += kLossBlock) {  // zero padding outside the image
        const int r = i / kHalo, c = i - r * kHalo;
        const int gy = y0 - kSsimRadius + r, gx = x0 - kSsimRadius + c;
        float xv = 0.0f, yv = 0.0f;
        if (gx >= 0 && gx < width && gy >= 0 && gy < height) {
            const size_t idx = plane + (size_t)gy * width + gx;
            xv = __ldg(image + idx);
            yv = load_target<TgtT>(target, idx);
        }
        s_x[r * kHaloStride + c] = xv;
        s_y[r * kHaloStride + c] = yv;
    }
    __syncthreads();

    if (tid < kHalo * (kLossTile / kHorizOut)) {
        const int r = tid / (kLossTile / kHorizOut), g = tid % (kLossTile / kHorizOut);
        conv_rows<5>(win, s_h, r, g, [&](int rr, int cc, float (&v)[5]) {
            const float xv = s_x[rr * kHaloStride + cc], yv = s_y[rr * kHaloStride + cc];
            v[0] = xv;
            v[1] = yv;
            v[2] = xv * xv;
            v[3] = yv * yv;
            v[4] = xv * yv;
        });
    }
    __syncthreads();

    const int c = tid & (kLossTile - 1), rg = tid / kLossTile;
    float acc[5][kVertOut];
    conv_cols<5>(win, s_h, c, rg, acc);

    float l_abs = 0.0f, l_ssim = 0.0f;
    const int gx = x0 + c;
#pragma unroll
    for (int o = 0; o < kVertOut; ++o) {
        const int row = kVertOut * rg + o, gy = y0 + row;
        if (gx < width && gy < height) {
            const float mu1 = acc[0][o], mu2 = acc[1][o];
            const float s11 = acc[2][o] - mu1 * mu1, s22 = acc[3][o] - mu2 * mu2, s12 = acc[4][o] - mu1 * mu2;
            const float A = mu1 * mu1 + mu2 * mu2 + kC1, B = s11 + s22 + kC2;
            const float C = 2.0f * mu1 * mu2 + kC1, D = 2.0f * s12 + kC2;
            const float rA = 1.0f / A, rB = 1.0f / B;
            const float rAB = rA * rB;
            const float m = C * D * rAB;
            // d m / d mu1 with sigma1^2 = E[x^2] - mu1^2 and sigma12 = E[xy] - mu1 mu2 substituted:
            //   2 mu2 (D - C) / (A B) + 2 mu1 m (1/B - 1/A)
            const float dm_dmu1 = 2.0f * (mu2 * (D - C) * rAB + mu1 * m * (rB - rA));
            const float dm_dsigma1_sq = -m * rB;
            const float dm_dsigma12 = 2.0f * C * rAB;
            const size_t idx = plane + (size_t)gy * width + gx;
            maps[idx] = dm_dmu1;
            maps[3 * plane_px + idx] = dm_dsigma1_sq;
            maps[6 * plane_px + idx] = dm_dsigma12;
            const int sidx = (row + kSsimRadius) * kHaloStride + c + kSsimRadius;
            l_abs += fabsf(s_x[sidx] - s_y[sidx]);
            l_ssim += m;
        }
    }
    l_abs = warp_sum(l_abs);
    l_ssim = warp_sum(l_ssim);
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = l_abs;
        s_red[1][tid >> 5] = l_ssim;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, s = 0.0;
        for (int w = 0; w < kLossBlock / 32; ++w) {
            a += (double)s_red[0][w];
            s += (double)s_red[1][w];
        }
        publish_sums(hdr, loss_out, a, s, gridDim.x * gridDim.y * gridDim.z, 3.0 * (double)plane_px, lambda_l1, lambda_dssim, true);
    }
}

// ---- gradient: convolve the derivative maps, add the L1 term --------------------------------------------------
template <typename TgtT>
__global__ void __launch_bounds__(kLossBlock) ssim_grad_kernel(const float* __restrict__ image, const TgtT* __restrict__ target,
                                                               const float* __restrict__ maps, const float* __restrict__ grad_loss,
                                                               float* __restrict__ grad_image, const int width, const int height,
                                                               const SsimWindow win, const float lambda_l1,
                                                               const float lambda_dssim) {
    __shared__ float s_m[3][kHalo * kHaloStride];
    __shared__ float s_h[3][kHalo][kRowStride];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kLossTile, y0 = blockIdx.y * kLossTile;
    const size_t plane_px = (size_t)width * height;
    const size_t plane = (size_t)blockIdx.z * plane_px;

    for (int i = tid; i < kHalo * kHalo; i += kLossBlock) {  // the maps only exist inside the image: zero outside
        const int r = i / kHalo, c = i - r * kHalo;
        const int gy = y0 - kSsimRadius + r, gx = x0 - kSsimRadius + c;
        float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
        if (gx >= 0 && gx < width && gy >= 0 && gy < height) {
            const size_t idx = plane + (size_t)gy * width + gx;
            v0 = __ldg(maps + idx);
            v1 = __ldg(maps + 3 * plane_px + idx);
            v2 = __ldg(maps + 6 * plane_px + idx);
        }
        s_m[0][r * kHaloStride + c] = v0;
        s_m[1][r * kHaloStride + c] = v1;
        s_m[2][r * kHaloStride + c] = v2;
    }
    __syncthreads();

    if (tid < kHalo * (kLossTile / kHorizOut)) {
        const int r = tid / (kLossTile / kHorizOut), g = tid % (kLossTile / kHorizOut);
        conv_rows<3>(win, s_h, r, g, [&](int rr, int cc, float (&v)[3]) {
            v[0] = s_m[0][rr * kHaloStride + cc];
            v[1] = s_m[1][rr * kHaloStride + cc];
            v[2] = s_m[2][rr * kHaloStride + cc];
        });
    }
    __syncthreads();

    const int c = tid & (kLossTile - 1), rg = tid / kLossTile;
    float acc[3][kVertOut];
    conv_cols<3>(win, s_h, c, rg, acc);

    const float n_rcp = 1.0f / (3.0f * (float)plane_px);
    const float gs = (grad_loss ? __ldg(grad_loss) : 1.0f) * n_rcp;
    const float g_l1 = gs * lambda_l1, g_ds = -gs * lambda_dssim;  // DSSIM = 1 - mean SSIM
    const int gx = x0 + c;
#pragma unroll
    for (int o = 0; o < kVertOut; ++o) {
        const int gy = y0 + kVertOut * rg + o;
        if (gx < width && gy < height) {
            const size_t idx = plane + (size_t)gy * width + gx;
            const float xv = __ldg(image + idx), yv = load_target<TgtT>(target, idx);
            const float dssim_dx = acc[0][o] + 2.0f * xv * acc[1][o] + yv * acc[2][o];
            grad_image[idx] = g_l1 * sign_of(xv - yv) + g_ds * dssim_dx;
        }
    }
}

SsimWindow make_window() {
    SsimWindow win;
    double g[kSsimTaps], sum = 0.0;
    for (int k = 0; k < kSsimTaps; ++k) {
        const double d = (double)(k - kSsimRadius);
        g[k] = std::exp(-d * d / (2.0 * 1.5 * 1.5));
        sum += g[k];
    }
    for (int k = 0; k < kSsimTaps; ++k) win.w[k] = (float)(g[k] / sum);
    return win;
}

int div_up_i(int a, int b) { return (a + b - 1) / b; }

struct LossScratch {
    LossHeader* hdr;
    float* maps;
    explicit LossScratch(void* p) : hdr(reinterpret_cast<LossHeader*>(p)), maps(reinterpret_cast<float*>(reinterpret_cast<char*>(p) + sizeof(LossHeader))) {}
};

bool misaligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

int check_args(const void* image, const void* target, bool target_is_u8, const void* grad_image, int width, int height,
               const void* scratch, const char* who) {
    if (!image || !target || !scratch) return report_error(-1, who, "NULL argument");
    if (width <= 0 || height <= 0) return report_error(-1, who, "empty image");
    // the streaming kernel uses 16-byte (float) / 4-byte (uint8) vector accesses
    if (misaligned(image, 16) || misaligned(grad_image, 16) || misaligned(target, target_is_u8 ? 4 : 16) || misaligned(scratch, 16))
        return report_error(-1, who, "image, grad_image, a float target and scratch must be 16-byte aligned, a uint8 target 4-byte");
    return 0;
}

template <typename TgtT>
int forward_t(const float* image, const TgtT* target, int width, int height, float lambda_l1, float lambda_dssim,
              void* scratch, float* loss_out, float* grad_image /* non-null: also the gradient for dL/dloss = 1 */,
              cudaStream_t stream) {
    LossScratch sc(scratch);
    cudaError_t e = cudaMemsetAsync(sc.hdr, 0, sizeof(LossHeader), stream);
    if (e != cudaSuccess) return report_error((int)e, "fgs_loss_forward: memset", cudaGetErrorString(e));
    const size_t n_values = (size_t)3 * width * height;
    if (lambda_dssim == 0.0f) {
        const unsigned blocks = (unsigned)std::min<size_t>((n_values / 4 + kLossBlock - 1) / kLossBlock + 1, 148 * 8);
        if (grad_image) {
            auto k = l1_kernel<TgtT, true, true>;
            FGS_LAUNCH(k, dim3(blocks), dim3(kLossBlock), stream, image, target, grad_image, sc.hdr, loss_out,
                       (const float*)nullptr, n_values, lambda_l1);
        } else {
            auto k = l1_kernel<TgtT, true, false>;
            FGS_LAUNCH(k, dim3(blocks), dim3(kLossBlock), stream, image, target, (float*)nullptr, sc.hdr, loss_out,
                       (const float*)nullptr, n_values, lambda_l1);
        }
    } else {
        const SsimWindow win = make_window();
        const dim3 grid(div_up_i(width, kLossTile), div_up_i(height, kLossTile), 3);
        auto k = ssim_stats_kernel<TgtT>;
        FGS_LAUNCH(k, grid, dim3(kLossBlock), stream, image, target, sc.maps, sc.hdr, loss_out, width, height, win, lambda_l1,
                   lambda_dssim);
        if (grad_image) {
            auto kg = ssim_grad_kernel<TgtT>;
            FGS_LAUNCH(kg, grid, dim3(kLossBlock), stream, image, target, (const float*)sc.maps, (const float*)nullptr,
                       grad_image, width, height, win, lambda_l1, lambda_dssim);
        }
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : report_error((int)e, "fgs_loss_forward: launch", cudaGetErrorString(e));
}

template <typename TgtT>
int backward_t(const float* image, const TgtT* target, int width, int height, float lambda_l1, float lambda_dssim,
               const void* scratch, const float* grad_loss, float* grad_image, cudaStream_t stream) {
    LossScratch sc(const_cast<void*>(scratch));
    const size_t n_values = (size_t)3 * width * height;
    if (lambda_dssim == 0.0f) {
        const unsigned blocks = (unsigned)std::min<size_t>((n_values / 4 + kLossBlock - 1) / kLossBlock + 1, 148 * 8);
        auto k = l1_kernel<TgtT, false, true>;
        FGS_LAUNCH(k, dim3(blocks), dim3(kLossBlock), stream, image, target, grad_image, (LossHeader*)nullptr, (float*)nullptr,
                   grad_loss, n_values, lambda_l1);
    } else {
        const SsimWindow win = make_window();
        const dim3 grid(div_up_i(width, kLossTile), div_up_i(height, kLossTile), 3);
        auto kg = ssim_grad_kernel<TgtT>;
        FGS_LAUNCH(kg, grid, dim3(kLossBlock), stream, image, target, (const float*)sc.maps, grad_loss, grad_image, width,
                   height, win, lambda_l1, lambda_dssim);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : report_error((int)e, "fgs_loss_backward: launch", cudaGetErrorString(e));
}

}  // namespace
}  // namespace fgs

extern "C" {

size_t fgs_loss_scratch_bytes(int32_t width, int32_t height) {
    if (width <= 0 || height <= 0) return sizeof(fgs::LossHeader);
    return sizeof(fgs::LossHeader) + (size_t)9 * width * height * sizeof(float);
}

int fgs_loss_forward(const float* image, const void* target, int32_t target_is_u8, int32_t width, int32_t height,
                     float lambda_l1, float lambda_dssim, void* scratch, float* loss_out, void* stream) {
    if (int rc = fgs::check_args(image, target, target_is_u8 != 0, nullptr, width, height, scratch, "fgs_loss_forward")) return rc;
    if (!loss_out) return fgs::report_error(-1, "fgs_loss_forward", "NULL loss_out");
    return target_is_u8 ? fgs::forward_t<unsigned char>(image, static_cast<const unsigned char*>(target), width, height, lambda_l1,
                                                        lambda_dssim, scratch, loss_out, nullptr, static_cast<cudaStream_t>(stream))
                        : fgs::forward_t<float>(image, static_cast<const float*>(target), width, height, lambda_l1, lambda_dssim,
                                                scratch, loss_out, nullptr, static_cast<cudaStream_t>(stream));
}

int fgs_loss_backward(const float* image, const void* target, int32_t target_is_u8, int32_t width, int32_t height,
                      float lambda_l1, float lambda_dssim, const void* scratch, const float* grad_loss, float* grad_image,
                      void* stream) {
    if (int rc = fgs::check_args(image, target, target_is_u8 != 0, grad_image, width, height, scratch, "fgs_loss_backward")) return rc;
    if (!grad_image) return fgs::report_error(-1, "fgs_loss_backward", "NULL grad_image");
    return target_is_u8 ? fgs::backward_t<unsigned char>(image, static_cast<const unsigned char*>(target), width, height,
                                                         lambda_l1, lambda_dssim, scratch, grad_loss, grad_image,
                                                         static_cast<cudaStream_t>(stream))
                        : fgs::backward_t<float>(image, static_cast<const float*>(target), width, height, lambda_l1, lambda_dssim,
                                                 scratch, grad_loss, grad_image, static_cast<cudaStream_t>(stream));
}

int fgs_loss_forward_backward(const float* image, const void* target, int32_t target_is_u8, int32_t width, int32_t height,
                              float lambda_l1, float lambda_dssim, void* scratch, float* loss_out, float* grad_image,
                              void* stream) {
    if (int rc = fgs::check_args(image, target, target_is_u8 != 0, grad_image, width, height, scratch, "fgs_loss_forward_backward")) return rc;
    if (!loss_out || !grad_image) return fgs::report_error(-1, "fgs_loss_forward_backward", "NULL output");
    return target_is_u8 ? fgs::forward_t<unsigned char>(image, static_cast<const unsigned char*>(target), width, height, lambda_l1,
                                                        lambda_dssim, scratch, loss_out, grad_image, static_cast<cudaStream_t>(stream))
                        : fgs::forward_t<float>(image, static_cast<const float*>(target), width, height, lambda_l1, lambda_dssim,
                                                scratch, loss_out, grad_image, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
