// densify.cu -- densification / pruning / reordering of the Gaussians together with their Adam state
// (SURVEY.md §8f rank 2): the every-100-iterations step either side of the rasterizer hot loop.
//
// Reference semantics (all torch code in the reference, dozens of launches and several .item() syncs per call):
//   Model.py:312-366  adaptive_density_control  (duplicate small / split large Gaussians, then prune)
//   Model.py:276-292  prune(mask)               (NeRFICG Optim.adam_utils.prune_param_groups)
//   Model.py:294-310  sort(ordering)            (adam_utils.sort_param_groups)
//   Model.py:340-347  extend_param_groups       (append rows, zero Adam state for them)
//   Model.py:406,449  reset_state(indices)      (zero the Adam state of some rows)
//   Model.py:459-463  apply_morton_ordering     (CudaUtils.MortonEncoding.morton_encode + argsort + sort)
// NeRFICG's adam_utils and morton_encode are NOT in the reference tree: their behaviour is restated
// from the call sites (row selection / concatenation of every parameter tensor and of exp_avg / exp_avg_sq; new rows
// get zero state) and, for the Morton code, from the standard construction (21 bits per axis of the position
// normalised to the bounding box, x most significant) -- "parity unpinned", see oracle/densify_oracle.py.
//
// Design: every one of these operations is "build a row plan, then move rows".  A plan is an int64 per OUTPUT row:
//   bits  0..31  source row
//   bits 32..33  kind: 0 = existing Gaussian (parameters and Adam state move together)
//                      1 = clone (parameters copied, Adam state zero)
//                      2,3 = first / second child of a split (mean displaced, scale shrunk, Adam state zero)
//   bits 34..63  row of the split's normal sample
// so a plain torch int64 index tensor IS a plan of kind 0.  One kernel (`gather_rows_kernel`) then moves all rows of up
// to 24 tensors -- the 6 parameters, their 12 moments, densification_info, filter_3d -- in a single launch: HBM-bound,
// 708 B read + 708 B written per Gaussian, output fully coalesced.  Plans come from a flags byte per Gaussian through
// an order-preserving compaction (block counts -> one-block scan -> emit), so the result order is exactly the
// reference's cat/mask order.  The only host synchronisation is the read-back of the four totals (the output tensors
// must be allocated), where the reference has one .item() per mask.
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

constexpr int kPlanBlock = 1024;  // one Gaussian per thread in the compaction kernels
constexpr int kRowBlock = 256;
constexpr int kMaxRowTensors = FGS_MAX_ROW_TENSORS;
constexpr unsigned kFull = 0xffffffffu;

// flags byte per Gaussian
constexpr unsigned kKeepOrig = 1u, kKeepClone = 2u, kKeepSplit = 4u, kIsSplit = 8u;

// ---- adaptive density control: classify (Model.py:314-321,359-365) ------------------------------------------------
__global__ void __launch_bounds__(kRowBlock) adc_flags_kernel(const float* __restrict__ info, const float* __restrict__ scales,
                                                              const float* __restrict__ rotations,
                                                              const float* __restrict__ opacities, const int n,
                                                              const float grad_threshold, const float small_scale_log,
                                                              const float min_opacity_logit, const int prune_large,
                                                              const float large_scale_log, unsigned char* __restrict__ flags) {
    const int i = blockIdx.x * kRowBlock + threadIdx.x;
    if (i >= n) return;
    const float count = __ldg(info + i), accum = __ldg(info + (size_t)n + i);
    const bool densify = accum >= grad_threshold * fmaxf(count, 1.0f);
    const float s0 = __ldg(scales + 3 * (size_t)i), s1 = __ldg(scales + 3 * (size_t)i + 1), s2 = __ldg(scales + 3 * (size_t)i + 2);
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    const bool small = smax <= small_scale_log;
    const bool clone = densify && small, split = densify && !small;
    const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
    const bool bad = __ldg(opacities + i) < min_opacity_logit || (q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w) < 1e-8f;
    // the prune mask is evaluated on the EXTENDED set: clones carry the parent's scales, split children the shrunk ones
    const bool large_parent = prune_large && smax > large_scale_log;
    const bool large_child = prune_large && logf(0.625f * expf(smax)) > large_scale_log;
    unsigned f = 0;
    if (!split && !bad && !large_parent) f |= kKeepOrig;
    if (clone && !bad && !large_parent) f |= kKeepClone;
    if (split && !bad && !large_child) f |= kKeepSplit;
    if (split) f |= kIsSplit;
    flags[i] = (unsigned char)f;
}

// ---- order-preserving compaction of the four flag bits -----------------------------------------------------------
__global__ void __launch_bounds__(kPlanBlock) plan_count_kernel(const unsigned char* __restrict__ flags, const int n,
                                                                uint32_t* __restrict__ block_counts) {
    __shared__ uint32_t s_w[kPlanBlock / 32][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * kPlanBlock + tid;
    const unsigned f = i < n ? flags[i] : 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned m = __ballot_sync(kFull, (f >> b) & 1u);
        if (lane == 0) s_w[warp][b] = (uint32_t)__popc(m);
    }
    __syncthreads();
    if (tid < 4) {
        uint32_t sum = 0;
        for (int w = 0; w < kPlanBlock / 32; ++w) sum += s_w[w][tid];
        block_counts[4 * (size_t)blockIdx.x + tid] = sum;
    }
}

// one block: exclusive scan of the per-block counts (each thread owns a contiguous run of blocks) + the four totals
__global__ void __launch_bounds__(kPlanBlock) plan_scan_kernel(const uint32_t* __restrict__ block_counts, const int n_blocks,
                                                               uint32_t* __restrict__ block_offsets, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_warp[kPlanBlock / 32][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n_blocks + kPlanBlock - 1) / kPlanBlock;
    const int b0 = min(tid * per, n_blocks), b1 = min(b0 + per, n_blocks);
    uint32_t loc[4] = {0, 0, 0, 0}, incl[4];
    for (int b = b0; b < b1; ++b)
#pragma unroll
        for (int k = 0; k < 4; ++k) loc[k] += block_counts[4 * (size_t)b + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = loc[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(kFull, v, o);
            if (lane >= o) v += u;
        }
        incl[k] = v;
        if (lane == 31) s_warp[warp][k] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = s_warp[lane][k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(kFull, v, o);
                if (lane >= o) v += u;
            }
            __syncwarp();
            s_warp[lane][k] = v;  // inclusive over the warps
        }
    }
    __syncthreads();
    uint32_t run[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) run[k] = (warp > 0 ? s_warp[warp - 1][k] : 0u) + incl[k] - loc[k];
    for (int b = b0; b < b1; ++b)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            block_offsets[4 * (size_t)b + k] = run[k];
            run[k] += block_counts[4 * (size_t)b + k];
        }
    if (tid < 4) totals[tid] = s_warp[kPlanBlock / 32 - 1][tid];
}

__global__ void __launch_bounds__(kPlanBlock) plan_emit_kernel(const unsigned char* __restrict__ flags, const int n,
                                                               const uint32_t* __restrict__ block_offsets,
                                                               const uint32_t* __restrict__ totals, long long* __restrict__ plan) {
    __shared__ uint32_t s_w[kPlanBlock / 32][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * kPlanBlock + tid;
    const unsigned f = i < n ? flags[i] : 0u;
    uint32_t rank[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned m = __ballot_sync(kFull, (f >> b) & 1u);
        rank[b] = (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_w[warp][b] = (uint32_t)__popc(m);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t before = block_offsets[4 * (size_t)blockIdx.x + b];
        for (int w = 0; w < warp; ++w) before += s_w[w][b];
        rank[b] += before;
    }
    if (i >= n) return;
    const unsigned long long k_orig = totals[0], k_clone = totals[1], k_split = totals[2], n_split = totals[3];
    const unsigned long long src = (unsigned long long)(unsigned)i;
    if (f & kKeepOrig) plan[rank[0]] = (long long)src;
    if (f & kKeepClone) plan[k_orig + rank[1]] = (long long)(src | (1ull << 32));
    if (f & kKeepSplit) {
        // sample rows follow the reference's randn_like of the (2 * n_split, 3) tensor: copy c of the k-th split (k
        // counted over ALL splits, kept or not) uses row c * n_split + k
        plan[k_orig + k_clone + rank[2]] = (long long)(src | (2ull << 32) | ((unsigned long long)rank[3] << 34));
        plan[k_orig + k_clone + k_split + rank[2]] = (long long)(src | (3ull << 32) | ((n_split + rank[3]) << 34));
    }
}

// ---- move rows ----------------------------------------------------------------------------------------------------
struct RowTensor {
    const float* src;
    const float* appended;  // rows for j >= n_copy (NULL: zeros)
    float* dst;
    int row_floats;
    int role;  // FGS_ROW_*
};
struct RowArgs {
    RowTensor t[kMaxRowTensors];
    const long long* plan;  // NULL: identity
    long long n_out;
    long long n_copy;
    const float* split_scales;     // source (N,3) raw log-scales   } needed when a plan contains split children
    const float* split_rotations;  // source (N,4) raw quaternions  }
    const float* split_samples;    // (2 * n_split, 3) standard normal samples
};

// IdxT: unsigned when every flat element index fits 32 bits (a 32-bit division per element instead of a 64-bit one)
template <typename IdxT>
__global__ void __launch_bounds__(kRowBlock) gather_rows_kernel(const RowArgs a) {
    const RowTensor t = a.t[blockIdx.y];
    const IdxT r = (IdxT)t.row_floats;
    const IdxT total = (IdxT)a.n_out * r;
    for (IdxT e = (IdxT)blockIdx.x * kRowBlock + threadIdx.x; e < total; e += (IdxT)gridDim.x * kRowBlock) {
        const IdxT jq = e / r;
        const int c = (int)(e - jq * r);
        const long long j = (long long)jq;
        float v = 0.0f;
        if (j >= a.n_copy) {
            if (t.appended) v = __ldg(t.appended + (j - a.n_copy) * (long long)r + c);
        } else {
            const unsigned long long entry = a.plan ? (unsigned long long)__ldg(a.plan + j) : (unsigned long long)j;
            const long long src_row = (long long)(entry & 0xffffffffull);
            const unsigned kind = (unsigned)(entry >> 32) & 3u;
            if (t.role == FGS_ROW_STATE) {
                if (kind == 0u) v = __ldg(t.src + src_row * (long long)r + c);
            } else {
                v = __ldg(t.src + src_row * (long long)r + c);
                if (kind >= 2u && t.role == FGS_ROW_SCALES) {
                    v = logf(0.625f * expf(v));  // Model.py:334 (1 / 1.6)
                } else if (kind >= 2u && t.role == FGS_ROW_MEANS) {
                    // Model.py:330-333: mean + R(q / |q|) (exp(s) * z); row c of the product
                    const long long sample = (long long)(entry >> 34);
                    const float4 q = __ldg(reinterpret_cast<const float4*>(a.split_rotations) + src_row);
                    const float nsq = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
                    const float s2 = 2.0f / nsq;
                    const float e0 = expf(__ldg(a.split_scales + 3 * src_row)) * __ldg(a.split_samples + 3 * sample);
                    const float e1 = expf(__ldg(a.split_scales + 3 * src_row + 1)) * __ldg(a.split_samples + 3 * sample + 1);
                    const float e2 = expf(__ldg(a.split_scales + 3 * src_row + 2)) * __ldg(a.split_samples + 3 * sample + 2);
                    const float xx = q.y * q.y, yy = q.z * q.z, zz = q.w * q.w;
                    const float xy = q.y * q.z, xz = q.y * q.w, yz = q.z * q.w, rx = q.x * q.y, ry = q.x * q.z, rz = q.x * q.w;
                    float off;
                    if (c == 0)
                        off = (1.0f - (yy + zz) * s2) * e0 + (xy - rz) * s2 * e1 + (xz + ry) * s2 * e2;
                    else if (c == 1)
                        off = (xy + rz) * s2 * e0 + (1.0f - (xx + zz) * s2) * e1 + (yz - rx) * s2 * e2;
                    else
                        off = (xz - ry) * s2 * e0 + (yz + rx) * s2 * e1 + (1.0f - (xx + yy) * s2) * e2;
                    v += off;
                }
            }
        }
        t.dst[e] = v;
    }
}

// rows[index[k]] = 0 for every tensor (Adam state reset of relocated Gaussians, Model.py:406,449)
struct ZeroArgs {
    float* dst[kMaxRowTensors];
    int row_floats[kMaxRowTensors];
    const long long* index;
    long long n_index;
};
__global__ void __launch_bounds__(kRowBlock) zero_rows_kernel(const ZeroArgs a) {
    float* dst = a.dst[blockIdx.y];
    const long long r = a.row_floats[blockIdx.y];
    const long long total = a.n_index * r;
    for (long long e = (long long)blockIdx.x * kRowBlock + threadIdx.x; e < total; e += (long long)gridDim.x * kRowBlock) {
        const long long k = e / r;
        dst[__ldg(a.index + k) * r + (e - k * r)] = 0.0f;
    }
}

// ---- Morton order ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ordered_bits(float f) {  // monotone float -> uint map
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bbox[0..2] = min, bbox[3..5] = max as ordered bits (initialised to 0xffffffff / 0 by the launcher)
__global__ void __launch_bounds__(kRowBlock) bbox_kernel(const float* __restrict__ means, const int n, unsigned* __restrict__ bbox) {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * kRowBlock + threadIdx.x; i < n; i += gridDim.x * kRowBlock)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned u = ordered_bits(__ldg(means + 3 * (size_t)i + c));
            lo[c] = min(lo[c], u);
            hi[c] = max(hi[c], u);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = min(lo[c], __shfl_xor_sync(kFull, lo[c], o));
            hi[c] = max(hi[c], __shfl_xor_sync(kFull, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bbox + c, lo[c]);
            atomicMax(bbox + 3 + c, hi[c]);
        }
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long v) {  // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(kRowBlock) morton_kernel(const float* __restrict__ means, const int n,
                                                           const unsigned* __restrict__ bbox, long long* __restrict__ codes,
                                                           uint32_t* __restrict__ key_lo, uint32_t* __restrict__ key_hi,
                                                           uint32_t* __restrict__ iota) {
    const int i = blockIdx.x * kRowBlock + threadIdx.x;
    if (i >= n) return;
    unsigned long long code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lo = from_ordered_bits(bbox[c]), hi = from_ordered_bits(bbox[3 + c]);
        const float extent = hi - lo;
        // IEEE division / multiplication whatever the build's fast-math setting: the code is an integer result
        const float u = extent > 0.0f ? __fdiv_rn(__ldg(means + 3 * (size_t)i + c) - lo, extent) : 0.0f;  // [0, 1]
        const float scaled = fminf(fmaxf(__fmul_rn(u, 2097151.0f), 0.0f), 2097151.0f);                    // 2^21 - 1
        code |= spread3((unsigned long long)scaled) << (2 - c);                                    // x most significant
    }
    if (codes) codes[i] = (long long)code;
    if (key_lo) key_lo[i] = (uint32_t)(code & 0xffffffffull);
    if (key_hi) key_hi[i] = (uint32_t)(code >> 32);
    if (iota) iota[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kRowBlock) gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ index,
                                                               uint32_t* __restrict__ dst, long long* __restrict__ dst_i64, const int n) {
    const int i = blockIdx.x * kRowBlock + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = __ldg(index + i);
    if (dst) dst[i] = __ldg(src + k);
    if (dst_i64) dst_i64[i] = (long long)k;
}

// ---- host side --------------------------------------------------------------------------------------------------------
struct PlanScratch {  // layout of the caller's scratch buffer (fgs_densify_scratch_bytes)
    uint32_t* totals;         // 4 (+ 6 bbox words behind them)
    uint32_t* block_counts;   // 4 per block
    uint32_t* block_offsets;  // 4 per block
    PlanScratch(void* p, int n) {
        const size_t nb = (size_t)((n > 0 ? n : 1) + kPlanBlock - 1) / kPlanBlock;
        totals = reinterpret_cast<uint32_t*>(p);
        block_counts = totals + 32;
        block_offsets = block_counts + 4 * nb;
    }
    static size_t bytes(int n) {
        const size_t nb = (size_t)((n > 0 ? n : 1) + kPlanBlock - 1) / kPlanBlock;
        return (32 + 8 * nb) * sizeof(uint32_t);
    }
};

int launch_status(const char* who) {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : report_error((int)e, who, cudaGetErrorString(e));
}

unsigned grid_for(long long elements) {
    const long long blocks = (elements + kRowBlock - 1) / kRowBlock;
    return (unsigned)std::max<long long>(1, std::min<long long>(blocks, 148 * 16));
}

}  // namespace
}  // namespace fgs

extern "C" {

size_t fgs_densify_scratch_bytes(int32_t n_primitives) { return fgs::PlanScratch::bytes(n_primitives); }

int fgs_adc_flags(const float* densification_info, const float* scales, const float* rotations, const float* opacities,
                  int32_t n_primitives, float grad_threshold, float small_scale_log, float min_opacity_logit,
                  int32_t prune_large, float large_scale_log, uint8_t* flags, void* stream) {
    if (n_primitives <= 0) return 0;
    if (!densification_info || !scales || !rotations || !opacities || !flags) return fgs::report_error(-1, "fgs_adc_flags", "NULL argument");
    auto k = fgs::adc_flags_kernel;
    FGS_LAUNCH_SERIAL(k, dim3((n_primitives + fgs::kRowBlock - 1) / fgs::kRowBlock), dim3(fgs::kRowBlock), static_cast<cudaStream_t>(stream),
               densification_info, scales, rotations, opacities, (int)n_primitives, grad_threshold, small_scale_log,
               min_opacity_logit, (int)prune_large, large_scale_log, flags);
    return fgs::launch_status("fgs_adc_flags");
}

int fgs_plan_count(const uint8_t* flags, int32_t n_primitives, void* scratch, uint32_t totals_host[4], void* stream) {
    if (!totals_host) return fgs::report_error(-1, "fgs_plan_count", "NULL totals_host");
    totals_host[0] = totals_host[1] = totals_host[2] = totals_host[3] = 0;
    if (n_primitives <= 0) return 0;
    if (!flags || !scratch) return fgs::report_error(-1, "fgs_plan_count", "NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    fgs::PlanScratch sc(scratch, n_primitives);
    const int nb = (n_primitives + fgs::kPlanBlock - 1) / fgs::kPlanBlock;
    auto kc = fgs::plan_count_kernel;
    FGS_LAUNCH(kc, dim3(nb), dim3(fgs::kPlanBlock), s, flags, (int)n_primitives, sc.block_counts);
    auto ks = fgs::plan_scan_kernel;
    FGS_LAUNCH(ks, dim3(1), dim3(fgs::kPlanBlock), s, (const uint32_t*)sc.block_counts, nb, sc.block_offsets, sc.totals);
    if (int rc = fgs::launch_status("fgs_plan_count")) return rc;
    cudaError_t e = cudaMemcpyAsync(totals_host, sc.totals, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the caller allocates the outputs from these four numbers
    return e == cudaSuccess ? 0 : fgs::report_error((int)e, "fgs_plan_count: read-back", cudaGetErrorString(e));
}

int fgs_plan_emit(const uint8_t* flags, int32_t n_primitives, const void* scratch, int64_t* plan, void* stream) {
    if (n_primitives <= 0) return 0;
    if (!flags || !scratch || !plan) return fgs::report_error(-1, "fgs_plan_emit", "NULL argument");
    fgs::PlanScratch sc(const_cast<void*>(scratch), n_primitives);
    const int nb = (n_primitives + fgs::kPlanBlock - 1) / fgs::kPlanBlock;
    auto k = fgs::plan_emit_kernel;
    FGS_LAUNCH(k, dim3(nb), dim3(fgs::kPlanBlock), static_cast<cudaStream_t>(stream), flags, (int)n_primitives,
               (const uint32_t*)sc.block_offsets, (const uint32_t*)sc.totals, reinterpret_cast<long long*>(plan));
    return fgs::launch_status("fgs_plan_emit");
}

int fgs_gather_rows(const fgs_row_tensor* tensors, int32_t n_tensors, const int64_t* plan, int64_t n_out, int64_t n_copy,
                    const float* split_scales, const float* split_rotations, const float* split_samples, void* stream) {
    if (n_tensors < 0 || n_tensors > FGS_MAX_ROW_TENSORS) return fgs::report_error(-1, "fgs_gather_rows", "too many tensors");
    if (n_out <= 0 || n_tensors == 0) return 0;
    if (!tensors || n_copy < 0 || n_copy > n_out) return fgs::report_error(-1, "fgs_gather_rows", "bad argument");
    fgs::RowArgs a{};
    long long widest = 1;
    for (int t = 0; t < n_tensors; ++t) {
        if (!tensors[t].dst || tensors[t].row_floats <= 0 || (n_copy > 0 && !tensors[t].src))
            return fgs::report_error(-1, "fgs_gather_rows", "tensor without source / destination / row size");
        if ((tensors[t].role == FGS_ROW_MEANS || tensors[t].role == FGS_ROW_SCALES) && tensors[t].row_floats != 3)
            return fgs::report_error(-1, "fgs_gather_rows", "means / scales rows must have 3 floats");
        a.t[t].src = tensors[t].src;
        a.t[t].appended = tensors[t].appended;
        a.t[t].dst = tensors[t].dst;
        a.t[t].row_floats = tensors[t].row_floats;
        a.t[t].role = tensors[t].role;
        widest = std::max<long long>(widest, tensors[t].row_floats);
    }
    a.plan = reinterpret_cast<const long long*>(plan);
    a.n_out = n_out;
    a.n_copy = n_copy;
    a.split_scales = split_scales;
    a.split_rotations = split_rotations;
    a.split_samples = split_samples;
    const dim3 grid(fgs::grid_for(n_out * widest), (unsigned)n_tensors);
    // the grid-stride loop adds gridDim.x * 256 to the index once past the end: keep that below 2^32 too
    if (n_out * widest < (1ll << 32) - (long long)grid.x * fgs::kRowBlock) {
        auto k = fgs::gather_rows_kernel<unsigned>;
        FGS_LAUNCH_SERIAL(k, grid, dim3(fgs::kRowBlock), static_cast<cudaStream_t>(stream), a);
    } else {
        auto k = fgs::gather_rows_kernel<unsigned long long>;
        FGS_LAUNCH_SERIAL(k, grid, dim3(fgs::kRowBlock), static_cast<cudaStream_t>(stream), a);
    }
    return fgs::launch_status("fgs_gather_rows");
}

int fgs_zero_rows(float* const* tensors, const int32_t* row_floats, int32_t n_tensors, const int64_t* index, int64_t n_index,
                  void* stream) {
    if (n_tensors < 0 || n_tensors > FGS_MAX_ROW_TENSORS) return fgs::report_error(-1, "fgs_zero_rows", "too many tensors");
    if (n_index <= 0 || n_tensors == 0) return 0;
    if (!tensors || !row_floats || !index) return fgs::report_error(-1, "fgs_zero_rows", "NULL argument");
    fgs::ZeroArgs a{};
    long long widest = 1;
    for (int t = 0; t < n_tensors; ++t) {
        if (!tensors[t] || row_floats[t] <= 0) return fgs::report_error(-1, "fgs_zero_rows", "bad tensor");
        a.dst[t] = tensors[t];
        a.row_floats[t] = row_floats[t];
        widest = std::max<long long>(widest, row_floats[t]);
    }
    a.index = reinterpret_cast<const long long*>(index);
    a.n_index = n_index;
    auto k = fgs::zero_rows_kernel;
    FGS_LAUNCH_SERIAL(k, dim3(fgs::grid_for(n_index * widest), (unsigned)n_tensors), dim3(fgs::kRowBlock), static_cast<cudaStream_t>(stream), a);
    return fgs::launch_status("fgs_zero_rows");
}

int fgs_morton_keys(const float* means, int32_t n_primitives, void* scratch, int64_t* codes, uint32_t* key_lo, uint32_t* key_hi,
                    uint32_t* iota, void* stream) {
    if (n_primitives <= 0) return 0;
    if (!means || !scratch) return fgs::report_error(-1, "fgs_morton_keys", "NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned* bbox = reinterpret_cast<unsigned*>(scratch) + 8;  // behind the plan totals
    cudaError_t e = cudaMemsetAsync(bbox, 0xff, 3 * sizeof(unsigned), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(bbox + 3, 0, 3 * sizeof(unsigned), s);
    if (e != cudaSuccess) return fgs::report_error((int)e, "fgs_morton_keys: memset", cudaGetErrorString(e));
    auto kb = fgs::bbox_kernel;
    FGS_LAUNCH(kb, dim3(fgs::grid_for(n_primitives)), dim3(fgs::kRowBlock), s, means, (int)n_primitives, bbox);
    auto km = fgs::morton_kernel;
    FGS_LAUNCH_SERIAL(km, dim3((n_primitives + fgs::kRowBlock - 1) / fgs::kRowBlock), dim3(fgs::kRowBlock), s, means, (int)n_primitives,
               (const unsigned*)bbox, reinterpret_cast<long long*>(codes), key_lo, key_hi, iota);
    return fgs::launch_status("fgs_morton_keys");
}

int fgs_gather_u32(const uint32_t* src, const uint32_t* index, uint32_t* dst, int64_t* index_as_i64, int32_t n, void* stream) {
    if (n <= 0) return 0;
    if (!index || (dst && !src) || (!dst && !index_as_i64)) return fgs::report_error(-1, "fgs_gather_u32", "NULL argument");
    auto k = fgs::gather_u32_kernel;
    FGS_LAUNCH_SERIAL(k, dim3((n + fgs::kRowBlock - 1) / fgs::kRowBlock), dim3(fgs::kRowBlock), static_cast<cudaStream_t>(stream), src, index,
               dst, reinterpret_cast<long long*>(index_as_i64), (int)n);
    return fgs::launch_status("fgs_gather_u32");
}

}  // extern "C"
