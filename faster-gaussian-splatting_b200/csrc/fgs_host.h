// fgs_host.h -- host-side declarations shared by the kernels' launchers and the C-ABI drivers.
#pragma once
#ifndef FGS_CUDA_EMU
#include <cuda_runtime.h>
#endif
#include <cstddef>
#include <cstdint>

#include "fgs_device.cuh"

namespace fgs {

// device counter block (uint32 each); zeroed at the start of every forward
enum : int {
    kCtrVisible = 0,    // V: number of visible Gaussians
    kCtrInstances = 1,  // I: number of (Gaussian, tile) instances
    kCtrTileQueue = 2,  // persistent blend-forward tile queue
    kCtrTileQueueBwd = 3,
    kCtrTicket0 = 4,    // tickets for the chained-scan kernels (one per launch): 4..14
    kCtrBigList = 15,   // length of the work list of large-footprint Gaussians (binning.cu)
    kCtrCount = 16
};

constexpr int kSortBlock = 256;
#ifndef FGS_SORT_ITEMS
#define FGS_SORT_ITEMS 8
#endif
constexpr int kSortItems = FGS_SORT_ITEMS;
constexpr int kSortTile = kSortBlock * kSortItems;  // keys per block per pass
constexpr int kRadixMax = 256;

// Host-side caches (occupancy-derived grids, one-time function attributes, side streams, capacity hints) are kept PER
// DEVICE: a process may drive several GPUs (one rasterizer call at a time per device; calls on different devices may come
// from different host threads).  `here()` is the slot of the calling thread's current device.
constexpr int kMaxDevices = 64;
inline int device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    return d % kMaxDevices;
}
template <typename T>
struct PerDevice {
    T slot[kMaxDevices] = {};
    T& here() { return slot[device_slot()]; }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- launchers ---------------------------------------------------------------------------------------
cudaError_t launch_preprocess(const float* means, const float* scales, const float* rotations,
                              const float* opacities, const float* sh0, const float* sh_rest, Rec* recs,
                              uint32_t* n_touched, uint2* tile_info, uint32_t* depth_keys, uint32_t* vis_indices,
                              uint32_t* counters, uint32_t* depth_hist,
                              int n, const Cam& cam, bool inference, cudaStream_t stream,
                              const float* cov3d = nullptr /* (N,6): 3D covariances given directly, scales/rotations unused */,
                              bool defer_color = false /* geometry only; launch_color fills `colors` later */);
// colours (clamped, sign-tagged like the record's colour fields) of the visible Gaussians, for a deferred-colour forward
cudaError_t launch_color(const float* means, const float* sh0, const float* sh_rest, const uint32_t* n_touched,
                         float4* colors, int n, const Cam& cam, cudaStream_t stream);

// LSD radix sort of (key, value) pairs, onesweep style (one read + one write of the data per pass).
// `n_dev` points at the element count on the device, `n_max` bounds it on the host (grid sizing).
// Keys are KeyT in {uint16_t, uint32_t}; bits [0, end_bit) are sorted.  `hist` needs
// passes * 256 uint32, `lookback` needs passes * div_up(n_max, kSortTile) * 256 uint32; both are
// zeroed inside when `zero_scratch` (they must be adjacent: hist first) or by the caller.  A non-null
// `pass0_hist` says the producer of the keys has already accumulated the digit histogram of the first
// pass (bits [0, radix_sort_pass0_bits(end_bit))) there, spread over `pass0_hist_replicas` copies of 256
// counters (summed by the first pass).  Returns in *result_in_alt whether the sorted data ended in the
// *_alt buffers.
template <typename KeyT>
cudaError_t radix_sort_pairs(KeyT* keys, KeyT* keys_alt, uint32_t* vals, uint32_t* vals_alt, const uint32_t* n_dev,
                             int n_max, int end_bit, uint32_t* hist, uint32_t* lookback, uint32_t* tickets,
                             bool* result_in_alt, bool zero_scratch, const uint32_t* pass0_hist, int pass0_hist_replicas,
                             cudaStream_t stream);
constexpr int kDepthHistReplicas = 32;  // copies of the depth keys' first-digit histogram (spreads preprocess's atomics over L2)
size_t radix_sort_lookback_words(int n_max, int end_bit);
int radix_sort_passes(int end_bit);
int radix_sort_pass0_bits(int end_bit);

// offsets[i] = exclusive prefix sum over i of the touched-tile count of Gaussian sorted_idx[i], i < *n_dev;
// sorted_info[i] = tile_info[sorted_idx[i]] (the footprints in depth order, so create_instances reads them coalesced)
cudaError_t launch_offsets_scan(const uint32_t* sorted_idx, const uint2* tile_info, uint32_t* offsets, uint2* sorted_info,
                                const uint32_t* n_dev, int n_max, uint32_t* lookback, uint32_t* ticket,
                                bool zero_scratch, cudaStream_t stream);
size_t offsets_scan_lookback_words(int n_max);

template <typename KeyT>
cudaError_t launch_create_instances(const uint32_t* sorted_idx, const uint32_t* offsets, const Rec* recs,
                                    const uint2* sorted_info,
                                    KeyT* inst_keys, uint32_t* inst_vals, const uint32_t* n_vis_dev, int n_max,
                                    int grid_w, uint32_t* key_hist, int key_hist_bits, uint32_t* big_count,
                                    uint32_t* big_prims, uint32_t* big_offsets, int capacity, cudaStream_t stream);

// per-tile [start, end) ranges from the tile-sorted keys + gather of the 48-byte records into
// instance order (tile-contiguous, so the blend kernels can bulk-copy them).  The instance count is read
// from the device and clamped to `capacity` (the size the instance buffers were allocated for).
template <typename KeyT>
cudaError_t launch_ranges_and_gather(const KeyT* sorted_keys, const uint32_t* sorted_vals, const Rec* recs,
                                     Rec* inst_recs, uint2* tile_ranges, const uint32_t* n_instances_dev, int capacity,
                                     int grid_w, cudaStream_t stream,
                                     const float4* colors = nullptr /* deferred-colour forward: colours per Gaussian */);

struct BlendFwdArgs {
    const Rec* inst_recs;
    const uint2* tile_ranges;
    float* image;             // (3,H,W) or (H,W,3)
    float* final_T;           // per pixel, training only
    uint32_t* n_processed;    // per pixel, training only
    uint32_t* tile_max_n;     // per tile, training only
    uint32_t* queue;          // tile queue counter
    unsigned long long* stats;  // instrumentation build only (-DFGS_BLEND_STATS), else unused
    int n_tiles;
    bool training;
    bool to_chw;
    bool clamp_output;
};
cudaError_t launch_blend_forward(const BlendFwdArgs& a, const Cam& cam, cudaStream_t stream);

struct BlendBwdArgs {
    const Rec* inst_recs;
    const uint2* tile_ranges;
    const float* image;
    const float* grad_image;
    const float* final_T;
    const uint32_t* n_processed;
    const uint32_t* tile_max_n;
    Grad2D* grad2d;           // indexed by primitive index, zeroed by the caller
    float* scores;            // non-null: accumulate Speedy-Splat pruning scores instead of gradients
    uint32_t* queue;
    unsigned long long* stats;  // instrumentation build only (-DFGS_BLEND_STATS), else unused
    int n_tiles;
};
cudaError_t launch_blend_backward(const BlendBwdArgs& a, const Cam& cam, cudaStream_t stream);

struct AdamHyper {
    float step_size[6];            // lr / (1 - beta1^t) per tensor: means, scales, rotations, opacities, sh0, sh_rest
    float bias2_sqrt_rcp[6];       // 1 / sqrt(1 - beta2^t)
    float beta1, beta2, eps;
};
struct PreBwdArgs {
    const float* means;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* sh_rest;
    const uint32_t* n_touched;
    const Grad2D* grad2d;
    float* densification_info;     // (2,N) or nullptr
    // unfused: gradient outputs (every row written, zeros for invisible Gaussians)
    float* g_means;
    float* g_scales;
    float* g_rotations;
    float* g_opacities;
    float* g_sh0;
    float* g_sh_rest;
    // covariance-input mode (unfused only): `scales` holds the (N,6) covariances, g_scales receives their (N,6) gradient
    // (xx, xy, xz, yy, yz, zz; an off-diagonal entry counts once, i.e. twice the per-entry matrix gradient),
    // rotations / g_rotations are unused
    bool cov3d;
    // fused backward + Adam: parameters and moments updated in place, no gradient tensors
    float* p[6];
    float* m[6];
    float* v[6];
    AdamHyper hyper;
    bool fused;
    // fused only: leave sh_rest (parameters and moments) alone and store the view direction in the 2D gradient record;
    // launch_adam_rest applies that part of the step later (side stream)
    bool defer_rest;
    int n;
};
cudaError_t launch_preprocess_backward(const PreBwdArgs& a, const Cam& cam, cudaStream_t stream, bool skip_invisible_rest = false);
// Overlapped schedule of the fused backward + Adam (opt-in, fgs_debug_set_overlap_adam): the zero-gradient Adam update of
// the sh_rest rows of invisible Gaussians runs on a side stream while the blend backward runs;
// launch_preprocess_backward(..., skip_invisible_rest = true) then leaves those rows alone.  Results are bit-identical.
bool preprocess_backward_can_defer_rest(const PreBwdArgs& a, const Cam& cam);
cudaError_t launch_adam_rest(const PreBwdArgs& a, const Cam& cam, cudaStream_t stream, int ctas_per_sm);
bool preprocess_backward_can_overlap(const PreBwdArgs& a, const Cam& cam);
cudaError_t launch_adam_rest_invisible(const PreBwdArgs& a, const Cam& cam, cudaStream_t stream);

cudaError_t launch_adam(const float* grad, float* param, float* exp_avg, float* exp_avg_sq, size_t n, float step_size,
                        float beta1, float beta2, float eps, float bias2_sqrt_rcp, cudaStream_t stream);

// small per-Gaussian operators around the rasterizer (aux_ops.cu)
cudaError_t launch_update_3d_filter(const float* positions, const float* w2c, float* filter_3d, unsigned char* visibility,
                                    int n, int width, int height, float focal_x, float focal_y, float center_x,
                                    float center_y, float near_plane, float clipping_tolerance, float distance2filter,
                                    cudaStream_t stream);
cudaError_t launch_relocation(const float* old_opacities, const float* old_scales, const long long* n_samples,
                              float* new_opacities, float* new_scales, int n, cudaStream_t stream);
cudaError_t launch_add_noise(const float* raw_scales, const float* raw_rotations, const float* raw_opacities,
                             const float* noise, float* means, int n, float lr, cudaStream_t stream);

}  // namespace fgs
