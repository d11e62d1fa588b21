// peer_adam.cu -- view-parallel training: gradient reduction + Adam + parameter broadcast in ONE kernel over
// NVLink / NVSwitch peer memory.
//
// No reference counterpart (the reference trains one view per iteration on one GPU, Trainer.py:180,
// Renderer.py:58-61).  Semantics = gradient accumulation over the batch followed by the
// reference's dense Adam (adam/src/adam.cu:24-33,54-56): p, m, v after the step equal what every rank would hold
// after "all-reduce(sum) the gradients, then FusedAdam.step" -- up to the order of the W-term gradient sum, which
// here is fixed (rank 0 first), so all replicas stay bit-identical.
//
// Every rank keeps the six gradient tensors and the six parameter tensors in two flat symmetric-memory buffers
// (same layout on every rank, every tensor region padded to a multiple of 4 floats) whose peer addresses are known
// to all ranks.  Rank r owns, of every tensor, the r-th of W contiguous slices (in float4 units).  One launch:
//     g  = sum over q of  peer_grads[q][i]          (W-1 of the loads cross NVLink)
//     Adam on (p, m, v)[i] with the rank's SHARD of the moments (Adam's HBM traffic / W, moments memory / W)
//     peer_params[q][i] = p  for every q             (W-1 of the stores cross NVLink)
// i.e. reduce-scatter + sharded Adam + all-gather with the same wire bytes as one all-reduce, without staging the
// reduced gradient in HBM.  The caller brackets the launch with two symmetric-memory barriers (all gradients
// written before; all parameter stores landed after).
//
// STATUS (round 2): run and measured on 2, 4 and 8 B200s (profiles/r2_results.md); the two-GPU test checks
// every variant against all-reduce + FusedAdam.  The default of the view-parallel path is the split compact step further
// down (fgs_peer_adam_small_step + fgs_peer_adam_rows_step); this one-kernel form serves several views per rank.
//
// This file is also compiled by g++ against tests/cuda_emu/cuda_emu.h (test-only emulator) -- see fgs_launch.h: the
// slicing / sharding logic is checked on the CPU by running W "ranks" one after the other on host buffers.
#include <cmath>
#include <cstdint>
#include <cstdio>

#include "../../include/fgs_b200.h"
#include "fgs_launch.h"

namespace fgs {

namespace {

constexpr int kPeerBlock = 256;

struct PeerAdamArgs {
    const float4* grads[FGS_MAX_PEERS];  // peer addresses of the flat gradient buffers, index = rank
    float4* params[FGS_MAX_PEERS];       // peer addresses of the flat parameter buffers
    float4* exp_avg;                     // this rank's shard of the first moments (float4 index = shard_off + j)
    float4* exp_avg_sq;
    uint32_t begin[6];      // first float4 of this rank's slice of tensor t inside the flat buffers
    uint32_t count[6];      // float4s in that slice
    uint32_t shard_off[6];  // first float4 of tensor t inside the moment shards
    float step_size[6];     // lr / (1 - beta1^t)
    float b2[6];            // 1 / sqrt(1 - beta2^t)
    float beta1, beta2, eps;
    int world, rank;
    uint32_t total;         // sum of count[]
    // NVLS variant (fgs_peer_adam_step_multicast): multicast addresses of the two flat buffers.  A load from grads_mc
    // with multimem.ld_reduce returns the sum over all W replicas, formed inside the NVSwitch; a multimem.st to
    // params_mc lands in all W replicas.  Per GPU and step: W-1 fewer gradient slices come in over NVLink, and the
    // new parameters leave once instead of W-1 times.
    const float4* grads_mc;
    float4* params_mc;
    // split compact step: the pre-update values of tensor 0 (means) of this rank's slice, float4 j of the slice -> means_old[j];
    // the sh_rest kernel that runs later on a side stream rebuilds the view directions from them
    float4* means_old;
};

__device__ __forceinline__ float4 ld_peer(const float4* p) {  // L2-only (the data is rewritten by its owner every step)
    return __ldcg(p);
}

__device__ __forceinline__ void adam1(float g, float& p, float& m, float& v, float step_size, float beta1, float beta2,
                                      float eps, float b2) {  // adam.cu:24-33
    const float g2 = g * g;
    m = fmaf(beta1, m - g, g);
    v = fmaf(beta2, v - g2, g2);
    p -= step_size * m / (sqrtf(v) * b2 + eps);
}

#ifndef FGS_CUDA_EMU
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Same slices, same Adam, reduction and broadcast done by the switch (one load + one store per float4 instead of W + W).
__global__ void __launch_bounds__(kPeerBlock) peer_adam_multicast_kernel(const PeerAdamArgs a) {
    for (uint32_t idx = blockIdx.x * kPeerBlock + threadIdx.x; idx < a.total; idx += gridDim.x * kPeerBlock) {
        uint32_t j = idx;
        int t = 0;
        while (t < 5 && j >= a.count[t]) {
            j -= a.count[t];
            ++t;
        }
        const size_t e = (size_t)a.begin[t] + j;
        const float4 s = multimem_ld_reduce_add(a.grads_mc + e);
        float4 p = a.params[a.rank][e];
        float4 m = a.exp_avg[a.shard_off[t] + j];
        float4 v = a.exp_avg_sq[a.shard_off[t] + j];
        if (t == 0 && a.means_old != nullptr) a.means_old[j] = p;
        const float ss = a.step_size[t], b2 = a.b2[t];
        adam1(s.x, p.x, m.x, v.x, ss, a.beta1, a.beta2, a.eps, b2);
        adam1(s.y, p.y, m.y, v.y, ss, a.beta1, a.beta2, a.eps, b2);
        adam1(s.z, p.z, m.z, v.z, ss, a.beta1, a.beta2, a.eps, b2);
        adam1(s.w, p.w, m.w, v.w, ss, a.beta1, a.beta2, a.eps, b2);
        a.exp_avg[a.shard_off[t] + j] = m;
        a.exp_avg_sq[a.shard_off[t] + j] = v;
        multimem_st(a.params_mc + e, p);
    }
}
#endif

#ifndef FGS_CUDA_EMU
// a contiguous range of this rank's buffer to the same range of every replica: one multimem.st per 16 bytes, replicated by
// the NVSwitch (the sh_rest slice after the row-sliced step, on a side stream under the next forward)
__global__ void __launch_bounds__(kPeerBlock) peer_broadcast_multicast_kernel(const float4* __restrict__ src, float4* mc_dst,
                                                                              const uint64_t n4) {
    constexpr int kU = 4;  // loads in flight per thread: the grid is small on purpose, the bytes in flight must not be
    const uint64_t stride = (uint64_t)gridDim.x * kPeerBlock;
    for (uint64_t i0 = (uint64_t)blockIdx.x * kPeerBlock + threadIdx.x; i0 < n4; i0 += stride * kU) {
        float4 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (i0 + stride * u < n4) v[u] = src[i0 + stride * u];
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (i0 + stride * u < n4) multimem_st(mc_dst + i0 + stride * u, v[u]);
    }
}
#endif

__global__ void __launch_bounds__(kPeerBlock) peer_adam_kernel(const PeerAdamArgs a) {
    for (uint32_t idx = blockIdx.x * kPeerBlock + threadIdx.x; idx < a.total; idx += gridDim.x * kPeerBlock) {
        uint32_t j = idx;
        int t = 0;
        while (t < 5 && j >= a.count[t]) {
            j -= a.count[t];
            ++t;
        }
        const size_t e = (size_t)a.begin[t] + j;
        // all W gradient loads in flight before the first one is consumed
        float4 g[FGS_MAX_PEERS];
#pragma unroll
        for (int q = 0; q < FGS_MAX_PEERS; ++q)
            if (q < a.world) g[q] = ld_peer(a.grads[q] + e);
        float4 p = a.params[a.rank][e];
        float4 m = a.exp_avg[a.shard_off[t] + j];
        float4 v = a.exp_avg_sq[a.shard_off[t] + j];
        if (t == 0 && a.means_old != nullptr) a.means_old[j] = p;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < FGS_MAX_PEERS; ++q) {  // fixed order: every replica receives the same bits
            if (q < a.world) {
                s.x += g[q].x;
                s.y += g[q].y;
                s.z += g[q].z;
                s.w += g[q].w;
            }
        }
        const float ss = a.step_size[t], b2 = a.b2[t];
        adam1(s.x, p.x, m.x, v.x, ss, a.beta1, a.beta2, a.eps, b2);
        adam1(s.y, p.y, m.y, v.y, ss, a.beta1, a.beta2, a.eps, b2);
        adam1(s.z, p.z, m.z, v.z, ss, a.beta1, a.beta2, a.eps, b2);
        adam1(s.w, p.w, m.w, v.w, ss, a.beta1, a.beta2, a.eps, b2);
        a.exp_avg[a.shard_off[t] + j] = m;
        a.exp_avg_sq[a.shard_off[t] + j] = v;
#pragma unroll
        for (int q = 0; q < FGS_MAX_PEERS; ++q)
            if (q < a.world) a.params[q][e] = p;
    }
}

// ---- row-sliced variant with a COMPACT gradient exchange (fgs_peer_adam_rows_step) ------------------------------------
// One view per rank and step.  The sh_rest gradient of a view is rank-1 per Gaussian: dL/dsh_rest[k] = Y_k(dir) * dL/dcolour
// (sh_utils.cuh:85-153), and dL/dsh0 = C0 * dL/dcolour is in the exchange anyway.  So the 45 sh_rest floats of every peer
// are never read: the owner of a Gaussian rebuilds each peer's contribution from that peer's sh0 gradient (3 floats), the
// peer's camera position (a 4-float slot at the end of its gradient buffer) and the Gaussian's mean -- 14 instead of 59
// floats per Gaussian and peer cross NVLink.  Ranks own ROWS (Gaussians), not float4 ranges, so that a warp sees whole
// Gaussians; the new sh_rest parameters are optionally kept local (`broadcast_rest` = 0) for the caller to broadcast on a
// side stream while the next forward's geometry pass runs (fgs_b200.parallel.PeerAdam(compact=True)).
constexpr int kRowsBlock = 128;
constexpr int kRowsWarps = kRowsBlock / 32;
constexpr int kRowsWarpFloats = 32 * 45 + FGS_MAX_PEERS * 96 + 96;  // staged sh_rest gradient, per-peer sh0 gradients, their sum

struct PeerRowsArgs {
    const float* grads[FGS_MAX_PEERS];
    float* params[FGS_MAX_PEERS];
    float* exp_avg;
    float* exp_avg_sq;
    uint32_t off[6];        // float offset of tensor t inside the flat buffers
    uint32_t kf[6];         // floats per row of tensor t (3, 3, 4, 1, 3, 3 * n_rest)
    uint32_t shard_off[6];  // float offset of tensor t inside the moment shards
    uint32_t cam_off;       // float offset of the camera-position slot inside every gradient buffer
    uint32_t row_begin, row_count;
    int active_bases;
    int broadcast_rest;
    float step_size[6];
    float b2[6];
    float beta1, beta2, eps;
    int world, rank;
    // NVLS (optional, both or none): multicast addresses of the flat gradient / parameter buffers.  The four plain sums
    // (means, scales, rotations, opacities) then come back from the switch already reduced (multimem.ld_reduce) and every
    // broadcast is ONE store that the switch replicates (multimem.st) instead of W peer loads / stores.
    const float* grads_mc;
    float* params_mc;
    // split compact step: only sh_rest is handled here (the five small tensors went through the float4-sliced kernel), and
    // the means come from the snapshot that kernel took before updating them (floats 3 * local_row ..)
    int rest_only;
    const float* means_old;
};

constexpr float kShC0 = 0.28209479177387814f;
constexpr float kShC1 = 0.48860251190291987f;
constexpr float kShC2a = 1.0925484305920792f, kShC2b = 0.94617469575755997f, kShC2c = 0.31539156525251999f,
                kShC2d = 0.54627421529603959f;
constexpr float kShC3a = 0.59004358992664352f, kShC3b = 1.7701307697799304f, kShC3c = 2.8906114426405538f,
                kShC3d = 0.45704579946446572f, kShC3e = 2.2852289973223288f, kShC3f = 1.865881662950577f,
                kShC3g = 1.1195289977703462f, kShC3h = 1.4453057213202769f;

// Y_1 .. Y_15 of a unit direction: the expressions of the backward kernel (preprocess_backward.cu, sh_utils.cuh:32-69)
__device__ __forceinline__ void peer_sh_basis(float x, float y, float z, int active_bases, float* __restrict__ Y) {
#pragma unroll
    for (int k = 0; k < 15; ++k) Y[k] = 0.f;
    if (active_bases > 1) {
        Y[0] = -kShC1 * y;
        Y[1] = kShC1 * z;
        Y[2] = -kShC1 * x;
        if (active_bases > 4) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            Y[3] = kShC2a * xy;
            Y[4] = -kShC2a * yz;
            Y[5] = kShC2b * zz - kShC2c;
            Y[6] = -kShC2a * xz;
            Y[7] = kShC2d * (xx - yy);
            if (active_bases > 9) {
                Y[8] = y * (kShC3a * yy - kShC3b * xx);
                Y[9] = kShC3c * xy * z;
                Y[10] = y * (kShC3d - kShC3e * zz);
                Y[11] = z * (kShC3f * zz - kShC3g);
                Y[12] = x * (kShC3d - kShC3e * zz);
                Y[13] = kShC3h * z * (xx - yy);
                Y[14] = x * (kShC3b * yy - kShC3a * xx);
            }
        }
    }
}

// One tensor's rows of the warp: gradient = sum over the peers' buffers (stage == nullptr) or the staged values; Adam with
// the rank's moment shard; new parameters to every peer (or only to this rank's replica).
__device__ __forceinline__ void rows_span(const PeerRowsArgs& a, int t, size_t elem0, size_t shard0, int n_floats,
                                          const float* __restrict__ stage, bool to_all, int lane) {
    float* m = a.exp_avg + shard0;
    float* v = a.exp_avg_sq + shard0;
    const float ss = a.step_size[t], b2 = a.b2[t];
    const bool vec = (n_floats & 3) == 0 && (elem0 & 3u) == 0 && (shard0 & 3u) == 0;
    if (vec) {
        for (int i = lane; i < (n_floats >> 2); i += 32) {
            float4 g;
            if (stage) {
                g = reinterpret_cast<const float4*>(stage)[i];
#ifndef FGS_CUDA_EMU
            } else if (a.grads_mc != nullptr) {
                g = multimem_ld_reduce_add(reinterpret_cast<const float4*>(a.grads_mc + elem0) + i);
#endif
            } else {
                float4 gq[FGS_MAX_PEERS];
#pragma unroll
                for (int q = 0; q < FGS_MAX_PEERS; ++q)
                    if (q < a.world) gq[q] = ld_peer(reinterpret_cast<const float4*>(a.grads[q] + elem0) + i);
                g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < FGS_MAX_PEERS; ++q) {
                    if (q < a.world) {
                        g.x += gq[q].x;
                        g.y += gq[q].y;
                        g.z += gq[q].z;
                        g.w += gq[q].w;
                    }
                }
            }
            float4 p4 = reinterpret_cast<const float4*>(a.params[a.rank] + elem0)[i];
            float4 m4 = reinterpret_cast<const float4*>(m)[i];
            float4 v4 = reinterpret_cast<const float4*>(v)[i];
            adam1(g.x, p4.x, m4.x, v4.x, ss, a.beta1, a.beta2, a.eps, b2);
            adam1(g.y, p4.y, m4.y, v4.y, ss, a.beta1, a.beta2, a.eps, b2);
            adam1(g.z, p4.z, m4.z, v4.z, ss, a.beta1, a.beta2, a.eps, b2);
            adam1(g.w, p4.w, m4.w, v4.w, ss, a.beta1, a.beta2, a.eps, b2);
            reinterpret_cast<float4*>(m)[i] = m4;
            reinterpret_cast<float4*>(v)[i] = v4;
            if (to_all) {
#ifndef FGS_CUDA_EMU
                if (a.params_mc != nullptr) {
                    multimem_st(reinterpret_cast<float4*>(a.params_mc + elem0) + i, p4);
                } else
#endif
                {
#pragma unroll
                    for (int q = 0; q < FGS_MAX_PEERS; ++q)
                        if (q < a.world) reinterpret_cast<float4*>(a.params[q] + elem0)[i] = p4;
                }
            } else {
                reinterpret_cast<float4*>(a.params[a.rank] + elem0)[i] = p4;
            }
        }
    } else {
        for (int i = lane; i < n_floats; i += 32) {
            float g = 0.f;
            if (stage) {
                g = stage[i];
            } else {
                for (int q = 0; q < a.world; ++q) g += __ldcg(a.grads[q] + elem0 + i);
            }
            float pe = a.params[a.rank][elem0 + i], me = m[i], ve = v[i];
            adam1(g, pe, me, ve, ss, a.beta1, a.beta2, a.eps, b2);
            m[i] = me;
            v[i] = ve;
            if (to_all) {
                for (int q = 0; q < a.world; ++q) a.params[q][elem0 + i] = pe;
            } else {
                a.params[a.rank][elem0 + i] = pe;
            }
        }
    }
}

// ---- the fast path of a full group (rows a multiple of 4: every tensor's span is whole float4s) -------------------------
// The kernel is a chain of memory round trips (remote loads cost microseconds), so every phase first issues ALL its loads
// and only then consumes them: the five small tensors are one merged list of float4 items (up to four per lane in flight),
// sh_rest goes four float4s per lane at a time, the peers' sh0 gradients arrive as one float4 per lane and peer.
struct RowsItem {
    size_t e;   // float offset inside the flat buffers
    size_t s;   // float offset inside the moment shards
    int t;      // tensor: 0 means, 1 scales, 2 rotations, 3 opacities, 4 sh0 (gradient staged), 5 sh_rest (gradient staged)
    int i;      // float4 index inside the tensor's span of this group
    bool valid;
};

__device__ __forceinline__ RowsItem rows_small_item(const PeerRowsArgs& a, int j, int rows, uint32_t r0, uint32_t local0) {
    const int c0 = (rows * 3) >> 2, c2 = rows, c3 = rows >> 2;  // float4s of means (= scales = sh0), rotations, opacities
    RowsItem it;
    it.valid = j < 3 * c0 + c2 + c3;
    if (j < c0) { it.t = 0; it.i = j; }
    else if (j < 2 * c0) { it.t = 1; it.i = j - c0; }
    else if (j < 2 * c0 + c2) { it.t = 2; it.i = j - 2 * c0; }
    else if (j < 2 * c0 + c2 + c3) { it.t = 3; it.i = j - 2 * c0 - c2; }
    else { it.t = 4; it.i = j - 2 * c0 - c2 - c3; }
    if (!it.valid) { it.t = 0; it.i = 0; }
    it.e = (size_t)a.off[it.t] + (size_t)r0 * a.kf[it.t] + 4u * (size_t)it.i;
    it.s = (size_t)a.shard_off[it.t] + (size_t)local0 * a.kf[it.t] + 4u * (size_t)it.i;
    return it;
}

template <bool kMC>
__device__ __forceinline__ float4 rows_peer_sum(const PeerRowsArgs& a, size_t e) {
#ifndef FGS_CUDA_EMU
    if (kMC) return multimem_ld_reduce_add(reinterpret_cast<const float4*>(a.grads_mc + e));
#endif
    float4 gq[FGS_MAX_PEERS];
#pragma unroll
    for (int q = 0; q < FGS_MAX_PEERS; ++q)
        if (q < a.world) gq[q] = ld_peer(reinterpret_cast<const float4*>(a.grads[q] + e));
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < FGS_MAX_PEERS; ++q) {
        if (q < a.world) {
            g.x += gq[q].x;
            g.y += gq[q].y;
            g.z += gq[q].z;
            g.w += gq[q].w;
        }
    }
    return g;
}

template <bool kMC>
__device__ __forceinline__ void rows_store_param(const PeerRowsArgs& a, size_t e, float4 p4, bool to_all) {
    if (!to_all) {
        *reinterpret_cast<float4*>(a.params[a.rank] + e) = p4;
        return;
    }
#ifndef FGS_CUDA_EMU
    if (kMC) {
        multimem_st(reinterpret_cast<float4*>(a.params_mc + e), p4);
        return;
    }
#endif
#pragma unroll
    for (int q = 0; q < FGS_MAX_PEERS; ++q)
        if (q < a.world) *reinterpret_cast<float4*>(a.params[q] + e) = p4;
}

__device__ __forceinline__ void rows_adam4(const PeerRowsArgs& a, int t, float4 g, float4& p4, float4& m4, float4& v4) {
    const float ss = a.step_size[t], b2 = a.b2[t];
    adam1(g.x, p4.x, m4.x, v4.x, ss, a.beta1, a.beta2, a.eps, b2);
    adam1(g.y, p4.y, m4.y, v4.y, ss, a.beta1, a.beta2, a.eps, b2);
    adam1(g.z, p4.z, m4.z, v4.z, ss, a.beta1, a.beta2, a.eps, b2);
    adam1(g.w, p4.w, m4.w, v4.w, ss, a.beta1, a.beta2, a.eps, b2);
}

// kMC: NVLS build (multimem.ld_reduce sums, multimem.st broadcasts); otherwise W peer loads / stores per float4.
template <bool kMC>
__global__ void __launch_bounds__(kRowsBlock, 3) peer_adam_rows_kernel(const PeerRowsArgs a) {
    FGS_DYNAMIC_SMEM(float, s_rows, 16);
    __shared__ float s_cam[FGS_MAX_PEERS * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage_rest = s_rows + warp * kRowsWarpFloats;
    float* s_dcol = stage_rest + 32 * 45;          // [peer][96]
    float* stage_sh0 = s_dcol + FGS_MAX_PEERS * 96;  // [96]
    const int rest_floats = (int)a.kf[5];
    const uint32_t groups = (a.row_count + 31u) >> 5;
    // every peer's camera position (remote, but only once per CTA)
    if ((int)threadIdx.x < a.world * 3) s_cam[threadIdx.x] = __ldcg(a.grads[threadIdx.x / 3] + a.cam_off + threadIdx.x % 3);
    __syncthreads();
    for (uint32_t group = blockIdx.x * kRowsWarps + warp; group < groups; group += gridDim.x * kRowsWarps) {
        const uint32_t r0 = a.row_begin + (group << 5);
        const int rows = (int)min(32u, a.row_begin + a.row_count - r0);
        const uint32_t local0 = group << 5;  // first row of the group inside this rank's slice
        const bool fast = (rows & 3) == 0;
        // the Gaussian's mean BEFORE this step's update: the view directions of the forward pass were computed from it
        float3 mean = make_float3(0.f, 0.f, 0.f);
        if (lane < rows) {
            const float* mp = a.rest_only ? a.means_old + (size_t)(local0 + lane) * 3
                                          : a.params[a.rank] + a.off[0] + (size_t)(r0 + lane) * 3;
            mean = make_float3(mp[0], mp[1], mp[2]);
        }
        // every peer's sh0 gradient of these rows (= C0 * dL/dcolour of that peer's view), and their sum in peer order
        if (fast) {
            const int n4 = (rows * 3) >> 2;
            if (lane < n4) {
                float4 gq[FGS_MAX_PEERS];
#pragma unroll
                for (int q = 0; q < FGS_MAX_PEERS; ++q)
                    if (q < a.world) gq[q] = ld_peer(reinterpret_cast<const float4*>(a.grads[q] + a.off[4] + (size_t)r0 * 3) + lane);
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < FGS_MAX_PEERS; ++q) {
                    if (q < a.world) {
                        reinterpret_cast<float4*>(s_dcol + q * 96)[lane] = gq[q];
                        sum.x += gq[q].x;
                        sum.y += gq[q].y;
                        sum.z += gq[q].z;
                        sum.w += gq[q].w;
                    }
                }
                reinterpret_cast<float4*>(stage_sh0)[lane] = sum;
            }
        } else {
            for (int i = lane; i < rows * 3; i += 32) {
                float sum = 0.f;
                for (int q = 0; q < a.world; ++q) {
                    const float g = __ldcg(a.grads[q] + a.off[4] + (size_t)r0 * 3 + i);
                    s_dcol[q * 96 + i] = g;
                    sum += g;
                }
                stage_sh0[i] = sum;
            }
        }
        __syncwarp();
        // rebuild the summed sh_rest gradient of this lane's Gaussian: sum over peers of Y_k(dir_q) * dL/dcolour_q
        if (rest_floats > 0 && lane < rows) {
            // one colour channel at a time (15 accumulators + 15 basis values live instead of 45 + 15: the basis is cheap to
            // recompute, registers are what limits the CTAs per SM here)
            float* mine = stage_rest + lane * rest_floats;
#pragma unroll 1
            for (int ch = 0; ch < 3; ++ch) {
                float acc[15];
#pragma unroll
                for (int k = 0; k < 15; ++k) acc[k] = 0.f;
                for (int q = 0; q < a.world; ++q) {
                    const float d = s_dcol[q * 96 + 3 * lane + ch] * (1.0f / kShC0);
                    if (d == 0.f) continue;  // not visible in that view (or no gradient in this channel)
                    const float rx = mean.x - s_cam[3 * q], ry = mean.y - s_cam[3 * q + 1], rz = mean.z - s_cam[3 * q + 2];
                    const float inv_len = rsqrtf(rx * rx + ry * ry + rz * rz);
                    float Y[15];
                    peer_sh_basis(rx * inv_len, ry * inv_len, rz * inv_len, a.active_bases, Y);
#pragma unroll
                    for (int k = 0; k < 15; ++k) acc[k] += Y[k] * d;
                }
#pragma unroll
                for (int k = 0; k < 15; ++k)
                    if (3 * k < rest_floats) mine[3 * k + ch] = acc[k];
            }
        }
        __syncwarp();
        if (!fast) {
            // ragged tail of the tensor (at most one group per rank): tensor after tensor, scalar where it has to be
            if (!a.rest_only) {
                for (int t = 0; t < 4; ++t)
                    rows_span(a, t, (size_t)a.off[t] + (size_t)r0 * a.kf[t], (size_t)a.shard_off[t] + (size_t)local0 * a.kf[t],
                              rows * (int)a.kf[t], nullptr, true, lane);
                rows_span(a, 4, (size_t)a.off[4] + (size_t)r0 * 3, (size_t)a.shard_off[4] + (size_t)local0 * 3, rows * 3, stage_sh0, true, lane);
            }
            if (rest_floats > 0)
                rows_span(a, 5, (size_t)a.off[5] + (size_t)r0 * rest_floats, (size_t)a.shard_off[5] + (size_t)local0 * rest_floats,
                          rows * rest_floats, stage_rest, a.broadcast_rest != 0, lane);
            __syncwarp();
            continue;
        }
        // means, scales, rotations, opacities (sums over the peers) and sh0 (the staged sum): one merged list of float4
        // items, up to four per lane, every load in flight before the first Adam update
        if (!a.rest_only) {
            constexpr int kU = 4;
            RowsItem it[kU];
            float4 g4[kU], p4[kU], m4[kU], v4[kU];
            if (!kMC) {  // W peer loads per item: one item at a time keeps the registers (and still overlaps its W + 3 loads)
#pragma unroll 1
                for (int u = 0; u < kU; ++u) {
                    const RowsItem x = rows_small_item(a, lane + 32 * u, rows, r0, local0);
                    if (!x.valid) continue;
                    float4 g = x.t == 4 ? reinterpret_cast<const float4*>(stage_sh0)[x.i] : rows_peer_sum<false>(a, x.e);
                    float4 pp = *reinterpret_cast<const float4*>(a.params[a.rank] + x.e);
                    float4 mm = *reinterpret_cast<const float4*>(a.exp_avg + x.s);
                    float4 vv = *reinterpret_cast<const float4*>(a.exp_avg_sq + x.s);
                    rows_adam4(a, x.t, g, pp, mm, vv);
                    *reinterpret_cast<float4*>(a.exp_avg + x.s) = mm;
                    *reinterpret_cast<float4*>(a.exp_avg_sq + x.s) = vv;
                    rows_store_param<false>(a, x.e, pp, true);
                }
            } else {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                it[u] = rows_small_item(a, lane + 32 * u, rows, r0, local0);
                if (it[u].valid) {
                    g4[u] = it[u].t == 4 ? reinterpret_cast<const float4*>(stage_sh0)[it[u].i] : rows_peer_sum<true>(a, it[u].e);
                    p4[u] = *reinterpret_cast<const float4*>(a.params[a.rank] + it[u].e);
                    m4[u] = *reinterpret_cast<const float4*>(a.exp_avg + it[u].s);
                    v4[u] = *reinterpret_cast<const float4*>(a.exp_avg_sq + it[u].s);
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (it[u].valid) {
                    rows_adam4(a, it[u].t, g4[u], p4[u], m4[u], v4[u]);
                    *reinterpret_cast<float4*>(a.exp_avg + it[u].s) = m4[u];
                    *reinterpret_cast<float4*>(a.exp_avg_sq + it[u].s) = v4[u];
                    rows_store_param<true>(a, it[u].e, p4[u], true);
                }
            }
            }
        }
        // sh_rest: the rebuilt gradient from shared memory, parameters and moments of this rank's rows, four float4s per lane
        // in flight; the new parameters go to every replica here or stay local for the caller's side-stream broadcast
        if (rest_floats > 0) {
            constexpr int kU = 4;
            const int n4 = (rows * rest_floats) >> 2;
            const size_t e0 = (size_t)a.off[5] + (size_t)r0 * rest_floats, s0 = (size_t)a.shard_off[5] + (size_t)local0 * rest_floats;
            const bool to_all = a.broadcast_rest != 0;
            for (int i0 = lane; i0 < n4; i0 += 32 * kU) {
                float4 p4[kU], m4[kU], v4[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int i = i0 + 32 * u;
                    if (i < n4) {
                        p4[u] = reinterpret_cast<const float4*>(a.params[a.rank] + e0)[i];
                        m4[u] = reinterpret_cast<const float4*>(a.exp_avg + s0)[i];
                        v4[u] = reinterpret_cast<const float4*>(a.exp_avg_sq + s0)[i];
                    }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int i = i0 + 32 * u;
                    if (i < n4) {
                        rows_adam4(a, 5, reinterpret_cast<const float4*>(stage_rest)[i], p4[u], m4[u], v4[u]);
                        reinterpret_cast<float4*>(a.exp_avg + s0)[i] = m4[u];
                        reinterpret_cast<float4*>(a.exp_avg_sq + s0)[i] = v4[u];
                        rows_store_param<kMC>(a, e0 + 4u * (size_t)i, p4[u], to_all);
                    }
                }
            }
        }
        __syncwarp();
    }
}

// slice of tensor t (count floats, padded to float4s) owned by `rank`: [begin4, begin4 + count4), chunk4 per rank
void slice_of(uint64_t count, int world, int rank, uint64_t* chunk4, uint64_t* begin4, uint64_t* count4) {
    const uint64_t n4 = (count + 3) / 4;
    const uint64_t c = (n4 + (uint64_t)world - 1) / (uint64_t)world;
    const uint64_t b = (uint64_t)rank * c < n4 ? (uint64_t)rank * c : n4;
    const uint64_t e = b + c < n4 ? b + c : n4;
    *chunk4 = c;
    *begin4 = b;
    *count4 = e - b;
}

thread_local char g_peer_err[256] = "";

}  // namespace

}  // namespace fgs

using namespace fgs;

extern "C" {

const char* fgs_peer_last_error(void) { return g_peer_err; }

size_t fgs_peer_adam_shard_floats(const uint64_t counts[6], int32_t world) {
    if (!counts || world < 1) return 0;
    uint64_t total4 = 0;
    for (int t = 0; t < 6; ++t) {
        uint64_t c, b, n;
        slice_of(counts[t], world, 0, &c, &b, &n);
        total4 += c;
    }
    return (size_t)(4 * total4);
}

int fgs_peer_adam_slice(const uint64_t counts[6], int32_t world, int32_t rank, uint64_t begin_floats[6],
                        uint64_t count_floats[6], uint64_t shard_offset_floats[6]) {
    if (!counts || world < 1 || rank < 0 || rank >= world) return -1;
    uint64_t off4 = 0;
    for (int t = 0; t < 6; ++t) {
        uint64_t c, b, n;
        slice_of(counts[t], world, rank, &c, &b, &n);
        if (begin_floats) begin_floats[t] = 4 * b;
        if (count_floats) count_floats[t] = 4 * n;
        if (shard_offset_floats) shard_offset_floats[t] = 4 * off4;
        off4 += c;
    }
    return 0;
}

static int peer_adam_step_impl(const float* const* peer_grads, float* const* peer_params, const float* grads_mc,
                               float* params_mc, bool multicast, int32_t world, int32_t rank, float* exp_avg_shard,
                               float* exp_avg_sq_shard, const uint64_t offsets[6], const uint64_t counts[6],
                               const int32_t step[6], const double lr[6], double beta1, double beta2, double eps,
                               void* stream) {
    auto fail = [](const char* msg) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "%s", msg);
        return -1;
    };
    if (!peer_grads || !peer_params || !exp_avg_shard || !exp_avg_sq_shard || !offsets || !counts || !step || !lr)
        return fail("fgs_peer_adam_step: NULL argument");
    if (world < 1 || world > FGS_MAX_PEERS || rank < 0 || rank >= world) return fail("fgs_peer_adam_step: bad world/rank");
    PeerAdamArgs a{};
    if (multicast) {
        if (!grads_mc || !params_mc) return fail("fgs_peer_adam_step_multicast: NULL multicast address (NVLS not available?)");
        if ((reinterpret_cast<uintptr_t>(grads_mc) | reinterpret_cast<uintptr_t>(params_mc)) & 15u)
            return fail("fgs_peer_adam_step_multicast: multicast addresses must be 16-byte aligned");
        a.grads_mc = reinterpret_cast<const float4*>(grads_mc);
        a.params_mc = reinterpret_cast<float4*>(params_mc);
    }
    for (int q = 0; q < world; ++q) {
        if (multicast && q != rank) continue;  // only the rank's own parameter replica is read directly
        if (!peer_grads[q] || !peer_params[q]) return fail("fgs_peer_adam_step: NULL peer pointer");
        if ((reinterpret_cast<uintptr_t>(peer_grads[q]) | reinterpret_cast<uintptr_t>(peer_params[q])) & 15u)
            return fail("fgs_peer_adam_step: peer buffers must be 16-byte aligned");
        a.grads[q] = reinterpret_cast<const float4*>(peer_grads[q]);
        a.params[q] = reinterpret_cast<float4*>(peer_params[q]);
    }
    if ((reinterpret_cast<uintptr_t>(exp_avg_shard) | reinterpret_cast<uintptr_t>(exp_avg_sq_shard)) & 15u)
        return fail("fgs_peer_adam_step: moment shards must be 16-byte aligned");
    a.exp_avg = reinterpret_cast<float4*>(exp_avg_shard);
    a.exp_avg_sq = reinterpret_cast<float4*>(exp_avg_sq_shard);
    uint64_t off4 = 0, total = 0;
    for (int t = 0; t < 6; ++t) {
        if (offsets[t] & 3u) return fail("fgs_peer_adam_step: tensor offsets must be multiples of 4 floats");
        uint64_t c, b, n;
        slice_of(counts[t], world, rank, &c, &b, &n);
        if (offsets[t] / 4 + b + n > 0xffffffffull || off4 + c > 0xffffffffull) return fail("fgs_peer_adam_step: buffers too large");
        a.begin[t] = (uint32_t)(offsets[t] / 4 + b);
        a.count[t] = (uint32_t)n;
        a.shard_off[t] = (uint32_t)off4;
        off4 += c;
        total += n;
        // bias corrections in double on the host, adam.cu:54-56
        a.step_size[t] = (float)(lr[t] / (1.0 - std::pow(beta1, (double)step[t])));
        a.b2[t] = (float)(1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step[t])));
    }
    a.beta1 = (float)beta1;
    a.beta2 = (float)beta2;
    a.eps = (float)eps;
    a.world = world;
    a.rank = rank;
    // the grid-stride index is 32-bit: leave room for one full stride past the end so that it cannot wrap
    if (total > 0xffffffffull - (1ull << 26)) return fail("fgs_peer_adam_step: more than 2^32 float4s per rank");
    a.total = (uint32_t)total;
    if (total == 0) return 0;
    static int s_grid_dev[64] = {};  // per device (fgs_host.h is not included here: this file stands alone for the emulator)
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0) cur_dev = 0;
    int& s_grid = s_grid_dev[cur_dev % 64];
    if (s_grid == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, peer_adam_kernel, kPeerBlock, 0);
        s_grid = (sms > 0 ? sms : 1) * (occ > 0 ? occ : 1);  // one persistent wave, grid-stride
    }
    const uint64_t want = (total + kPeerBlock - 1) / kPeerBlock;
    const int grid = want < (uint64_t)s_grid ? (int)want : s_grid;
#ifndef FGS_CUDA_EMU
    if (multicast) {
        auto km = peer_adam_multicast_kernel;
        FGS_LAUNCH_SERIAL(km, dim3((unsigned)grid), dim3(kPeerBlock), reinterpret_cast<cudaStream_t>(stream), a);
    } else
#endif
    {
        auto k = peer_adam_kernel;
        FGS_LAUNCH_SERIAL(k, dim3((unsigned)grid), dim3(kPeerBlock), reinterpret_cast<cudaStream_t>(stream), a);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_adam_step: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// The five small tensors of the split compact step through the float4-sliced kernel, with ROW-aligned slices (rank r owns
// the Gaussians [r * R, (r + 1) * R), like fgs_peer_adam_rows_step) and the moment layout of the row-sliced step.
int fgs_peer_adam_small_step(const float* const* peer_grads, float* const* peer_params, int32_t world, int32_t rank,
                             float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6], uint64_t n_rows,
                             int32_t n_rest, const int32_t step[6], const double lr[6], double beta1, double beta2,
                             double eps, float* means_old, const float* grads_multicast, float* params_multicast,
                             void* stream) {
    auto fail = [](const char* msg) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "%s", msg);
        return -1;
    };
    if (!peer_grads || !peer_params || !exp_avg_shard || !exp_avg_sq_shard || !offsets || !step || !lr)
        return fail("fgs_peer_adam_small_step: NULL argument");
    if (world < 1 || world > FGS_MAX_PEERS || rank < 0 || rank >= world) return fail("fgs_peer_adam_small_step: bad world/rank");
    if (n_rest < 0 || n_rest > 15) return fail("fgs_peer_adam_small_step: sh_rest must have <= 15 bases");
    PeerAdamArgs a{};
    for (int q = 0; q < world; ++q) {
        if (!peer_grads[q] || !peer_params[q]) return fail("fgs_peer_adam_small_step: NULL peer pointer");
        if ((reinterpret_cast<uintptr_t>(peer_grads[q]) | reinterpret_cast<uintptr_t>(peer_params[q])) & 15u)
            return fail("fgs_peer_adam_small_step: peer buffers must be 16-byte aligned");
        a.grads[q] = reinterpret_cast<const float4*>(peer_grads[q]);
        a.params[q] = reinterpret_cast<float4*>(peer_params[q]);
    }
    if ((reinterpret_cast<uintptr_t>(exp_avg_shard) | reinterpret_cast<uintptr_t>(exp_avg_sq_shard) |
         reinterpret_cast<uintptr_t>(means_old)) & 15u)
        return fail("fgs_peer_adam_small_step: moment shards and the means snapshot must be 16-byte aligned");
    a.exp_avg = reinterpret_cast<float4*>(exp_avg_shard);
    a.exp_avg_sq = reinterpret_cast<float4*>(exp_avg_sq_shard);
    a.means_old = reinterpret_cast<float4*>(means_old);
    if ((grads_multicast == nullptr) != (params_multicast == nullptr))
        return fail("fgs_peer_adam_small_step: pass both multicast addresses or none");
    if ((reinterpret_cast<uintptr_t>(grads_multicast) | reinterpret_cast<uintptr_t>(params_multicast)) & 15u)
        return fail("fgs_peer_adam_small_step: multicast addresses must be 16-byte aligned");
    a.grads_mc = reinterpret_cast<const float4*>(grads_multicast);
    a.params_mc = reinterpret_cast<float4*>(params_multicast);
    const uint32_t kf[6] = {3u, 3u, 4u, 1u, 3u, 3u * (uint32_t)n_rest};
    const uint64_t per = (((n_rows + (uint64_t)world - 1) / (uint64_t)world) + 31) / 32 * 32;
    const uint64_t begin = (uint64_t)rank * per < n_rows ? (uint64_t)rank * per : n_rows;
    const uint64_t end = begin + per < n_rows ? begin + per : n_rows;
    uint64_t shard = 0, total = 0;
    for (int t = 0; t < 6; ++t) {
        if (offsets[t] & 3u) return fail("fgs_peer_adam_small_step: tensor offsets must be multiples of 4 floats");
        // rows [begin, end) of tensor t: whole float4s (begin * kf is a multiple of 4; a ragged end takes the padding along)
        const uint64_t f0 = begin * kf[t], f1 = end * kf[t];
        const uint64_t n4 = t < 5 ? (f1 - f0 + 3) / 4 : 0;  // sh_rest is not handled here
        if (offsets[t] / 4 + f0 / 4 + n4 > 0xffffffffull) return fail("fgs_peer_adam_small_step: buffers too large");
        a.begin[t] = (uint32_t)(offsets[t] / 4 + f0 / 4);
        a.count[t] = (uint32_t)n4;
        a.shard_off[t] = (uint32_t)(shard / 4);
        shard += per * kf[t];
        total += n4;
        a.step_size[t] = (float)(lr[t] / (1.0 - std::pow(beta1, (double)step[t])));
        a.b2[t] = (float)(1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step[t])));
    }
    a.beta1 = (float)beta1;
    a.beta2 = (float)beta2;
    a.eps = (float)eps;
    a.world = world;
    a.rank = rank;
    a.total = (uint32_t)total;
    if (total == 0) return 0;
    int cur_dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0) cur_dev = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, peer_adam_kernel, kPeerBlock, 0);
    const uint64_t cap = (uint64_t)(sms > 0 ? sms : 1) * (uint64_t)(occ > 0 ? occ : 1);
    const uint64_t want = (total + kPeerBlock - 1) / kPeerBlock;
    const int grid = (int)(want < cap ? want : cap);
#ifndef FGS_CUDA_EMU
    if (a.grads_mc != nullptr) {  // NVLS: the sum comes back from the switch, one store reaches every replica
        auto km = peer_adam_multicast_kernel;
        FGS_LAUNCH_SERIAL(km, dim3((unsigned)grid), dim3(kPeerBlock), reinterpret_cast<cudaStream_t>(stream), a);
    } else
#endif
    {
        auto k = peer_adam_kernel;
        FGS_LAUNCH_SERIAL(k, dim3((unsigned)grid), dim3(kPeerBlock), reinterpret_cast<cudaStream_t>(stream), a);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_adam_small_step: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int fgs_peer_broadcast_multicast(const float* local_buffer, float* multicast_buffer, uint64_t first_float, uint64_t n_floats,
                                 int32_t max_ctas, void* stream) {
#ifdef FGS_CUDA_EMU
    (void)local_buffer; (void)multicast_buffer; (void)first_float; (void)n_floats; (void)max_ctas; (void)stream;
    std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_broadcast_multicast: not available on the emulator");
    return -1;
#else
    if (!local_buffer || !multicast_buffer) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_broadcast_multicast: NULL buffer (NVLS not available?)");
        return -1;
    }
    if ((first_float & 3u) || (n_floats & 3u) ||
        ((reinterpret_cast<uintptr_t>(local_buffer) | reinterpret_cast<uintptr_t>(multicast_buffer)) & 15u)) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_broadcast_multicast: range and buffers must be 16-byte aligned");
        return -1;
    }
    if (n_floats == 0) return 0;
    const uint64_t n4 = n_floats / 4;
    const uint64_t want = (n4 + kPeerBlock - 1) / kPeerBlock;
    const int cap = max_ctas > 0 ? max_ctas : 64;  // small on purpose: it runs next to the forward's kernels
    const int grid = want < (uint64_t)cap ? (int)want : cap;
    auto k = peer_broadcast_multicast_kernel;
    FGS_LAUNCH_SERIAL(k, dim3((unsigned)grid), dim3(kPeerBlock), reinterpret_cast<cudaStream_t>(stream),
                      reinterpret_cast<const float4*>(local_buffer + first_float),
                      reinterpret_cast<float4*>(multicast_buffer + first_float), n4);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_broadcast_multicast: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
#endif
}

uint64_t fgs_peer_adam_rows_per_rank(uint64_t n_rows, int32_t world) {
    if (world < 1) return 0;
    const uint64_t per = (n_rows + (uint64_t)world - 1) / (uint64_t)world;
    return (per + 31) / 32 * 32;  // multiples of 32 rows: every tensor's slice starts float4-aligned
}

int fgs_peer_adam_rows_step(const float* const* peer_grads, float* const* peer_params, int32_t world, int32_t rank,
                            float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6], uint64_t n_rows,
                            int32_t n_rest, int32_t active_sh_bases, uint64_t cam_offset, int32_t broadcast_rest,
                            const int32_t step[6], const double lr[6], double beta1, double beta2, double eps,
                            const float* grads_multicast, float* params_multicast, const float* means_old,
                            int32_t max_ctas_per_sm, void* stream) {
    auto fail = [](const char* msg) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "%s", msg);
        return -1;
    };
    if (!peer_grads || !peer_params || !exp_avg_shard || !exp_avg_sq_shard || !offsets || !step || !lr)
        return fail("fgs_peer_adam_rows_step: NULL argument");
    if (world < 1 || world > FGS_MAX_PEERS || rank < 0 || rank >= world) return fail("fgs_peer_adam_rows_step: bad world/rank");
    if (n_rest < 0 || n_rest > 15) return fail("fgs_peer_adam_rows_step: sh_rest must have <= 15 bases");
    PeerRowsArgs a{};
    for (int q = 0; q < world; ++q) {
        if (!peer_grads[q] || !peer_params[q]) return fail("fgs_peer_adam_rows_step: NULL peer pointer");
        if ((reinterpret_cast<uintptr_t>(peer_grads[q]) | reinterpret_cast<uintptr_t>(peer_params[q])) & 15u)
            return fail("fgs_peer_adam_rows_step: peer buffers must be 16-byte aligned");
        a.grads[q] = peer_grads[q];
        a.params[q] = peer_params[q];
    }
    if ((reinterpret_cast<uintptr_t>(exp_avg_shard) | reinterpret_cast<uintptr_t>(exp_avg_sq_shard)) & 15u)
        return fail("fgs_peer_adam_rows_step: moment shards must be 16-byte aligned");
    a.exp_avg = exp_avg_shard;
    a.exp_avg_sq = exp_avg_sq_shard;
    const uint32_t kf[6] = {3u, 3u, 4u, 1u, 3u, 3u * (uint32_t)n_rest};
    const uint64_t per = fgs_peer_adam_rows_per_rank(n_rows, world);
    const uint64_t begin = (uint64_t)rank * per < n_rows ? (uint64_t)rank * per : n_rows;
    const uint64_t end = begin + per < n_rows ? begin + per : n_rows;
    uint64_t shard = 0;
    for (int t = 0; t < 6; ++t) {
        if (offsets[t] & 3u) return fail("fgs_peer_adam_rows_step: tensor offsets must be multiples of 4 floats");
        if (offsets[t] + n_rows * kf[t] > 0xffffffffull) return fail("fgs_peer_adam_rows_step: buffers too large");
        a.off[t] = (uint32_t)offsets[t];
        a.kf[t] = kf[t];
        a.shard_off[t] = (uint32_t)shard;
        shard += per * kf[t];
        a.step_size[t] = (float)(lr[t] / (1.0 - std::pow(beta1, (double)step[t])));
        a.b2[t] = (float)(1.0 / std::sqrt(1.0 - std::pow(beta2, (double)step[t])));
    }
    if (cam_offset > 0xffffffffull) return fail("fgs_peer_adam_rows_step: buffers too large");
    a.cam_off = (uint32_t)cam_offset;
    a.row_begin = (uint32_t)begin;
    a.row_count = (uint32_t)(end - begin);
    a.active_bases = active_sh_bases < n_rest + 1 ? active_sh_bases : n_rest + 1;
    a.broadcast_rest = broadcast_rest;
    if ((grads_multicast == nullptr) != (params_multicast == nullptr))
        return fail("fgs_peer_adam_rows_step: pass both multicast addresses or none");
    if ((reinterpret_cast<uintptr_t>(grads_multicast) | reinterpret_cast<uintptr_t>(params_multicast)) & 15u)
        return fail("fgs_peer_adam_rows_step: multicast addresses must be 16-byte aligned");
    a.grads_mc = grads_multicast;
    a.params_mc = params_multicast;
    a.rest_only = means_old != nullptr;
    a.means_old = means_old;
    a.beta1 = (float)beta1;
    a.beta2 = (float)beta2;
    a.eps = (float)eps;
    a.world = world;
    a.rank = rank;
    if (a.row_count == 0) return 0;
    const size_t smem = (size_t)kRowsWarps * kRowsWarpFloats * sizeof(float);
    static int s_rows_grid[64] = {};
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0) cur_dev = 0;
    int& grid_cap = s_rows_grid[cur_dev % 64];
    const bool mc = a.grads_mc != nullptr;
    if (grid_cap == 0) {
        auto k0 = peer_adam_rows_kernel<false>;
        auto k1 = peer_adam_rows_kernel<true>;
        if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("fgs_peer_adam_rows_step: cannot reserve shared memory");
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
        grid_cap = (sms > 0 ? sms : 1) * 3;  // a persistent wave: three 128-thread CTAs per SM (168 registers)
    }
    const uint32_t groups = (a.row_count + 31u) / 32u;
    const uint32_t want = (groups + kRowsWarps - 1) / kRowsWarps;
    int cap = grid_cap;
    if (max_ctas_per_sm > 0 && max_ctas_per_sm < 3) cap = grid_cap / 3 * max_ctas_per_sm;  // next to another stream's kernels
    const int grid = want < (uint32_t)cap ? (int)want : cap;
    if (mc) {
        auto k = peer_adam_rows_kernel<true>;
        FGS_LAUNCH_SMEM(k, dim3((unsigned)grid), dim3(kRowsBlock), smem, reinterpret_cast<cudaStream_t>(stream), a);
    } else {
        auto k = peer_adam_rows_kernel<false>;
        FGS_LAUNCH_SMEM(k, dim3((unsigned)grid), dim3(kRowsBlock), smem, reinterpret_cast<cudaStream_t>(stream), a);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_adam_rows_step: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int fgs_peer_adam_step(const float* const* peer_grads, float* const* peer_params, int32_t world, int32_t rank,
                       float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6],
                       const uint64_t counts[6], const int32_t step[6], const double lr[6], double beta1, double beta2,
                       double eps, void* stream) {
    return peer_adam_step_impl(peer_grads, peer_params, nullptr, nullptr, false, world, rank, exp_avg_shard,
                               exp_avg_sq_shard, offsets, counts, step, lr, beta1, beta2, eps, stream);
}

int fgs_peer_adam_step_multicast(const float* grads_multicast, float* params_multicast, float* local_params,
                                 int32_t world, int32_t rank, float* exp_avg_shard, float* exp_avg_sq_shard,
                                 const uint64_t offsets[6], const uint64_t counts[6], const int32_t step[6],
                                 const double lr[6], double beta1, double beta2, double eps, void* stream) {
#ifdef FGS_CUDA_EMU
    (void)grads_multicast; (void)params_multicast; (void)local_params; (void)world; (void)rank; (void)exp_avg_shard;
    (void)exp_avg_sq_shard; (void)offsets; (void)counts; (void)step; (void)lr; (void)beta1; (void)beta2; (void)eps; (void)stream;
    std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_adam_step_multicast: not available on the emulator");
    return -1;
#else
    if (world < 1 || world > FGS_MAX_PEERS || rank < 0 || rank >= world) {
        std::snprintf(g_peer_err, sizeof(g_peer_err), "fgs_peer_adam_step_multicast: bad world/rank");
        return -1;
    }
    const float* g[FGS_MAX_PEERS] = {};
    float* p[FGS_MAX_PEERS] = {};
    g[rank] = reinterpret_cast<const float*>(local_params);  // placeholder (unused by the multicast kernel), non-NULL
    p[rank] = local_params;
    return peer_adam_step_impl(g, p, grads_multicast, params_multicast, true, world, rank, exp_avg_shard, exp_avg_sq_shard,
                               offsets, counts, step, lr, beta1, beta2, eps, stream);
#endif
}

}  // extern "C"
