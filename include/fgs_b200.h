/* fgs_b200.h -- C ABI of the B200-native Gaussian rasterizer (libfgs_b200.so).
 *
 * Drop-in boundary for the hot path of nerficg-project/faster-gaussian-splatting.  The reference
 * exposes this path as a pybind11 module over torch::Tensor (torch_bindings/bindings.cpp:12-21);
 * below it sits a torch-free layer taking raw device pointers and `std::function<char*(size_t)>`
 * resize callbacks (rasterization/include/forward.h:10-36, backward.h:8-44, inference.h:9-36,
 * adam/include/adam.h:7-16).  This header restates exactly that layer as plain C: pointers, sizes,
 * a CUDA stream and C callbacks -- no torch types.
 *
 * Conventions
 *  - every `const float*` / `float*` is a DEVICE pointer to contiguous float32 data, including
 *    w2c / cam_position / bg_color (the reference's kernels dereference them on the device too,
 *    rasterization_api.cu:68-70);
 *  - parameter layouts are the reference's: means (N,3), scales (N,3) raw log-scales, rotations (N,4)
 *    unnormalised (w,x,y,z), opacities (N,1) logits, sh0 (N,1,3), sh_rest (N,R,3), image (3,H,W);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it (the reference uses the
 *    legacy default stream);
 *  - scratch memory is owned by the caller and grown through `fgs_resize_fn` callbacks
 *    (utils/torch_utils.h:6-12): the callback must return a device pointer, aligned to >= 128 bytes,
 *    to at least `bytes` bytes that stay valid until the matching backward has run; the contents are
 *    opaque and only ever passed back to fgs_backward*;
 *  - return value: 0 on success, otherwise a cudaError_t value (or -1 for argument errors); the text
 *    is available from fgs_last_error().  The reference's release build has no error checking at all
 *    (utils/utils.h:8-15); the Python layer turns a non-zero status into RuntimeError like the
 *    reference's debug build does.
 *  - not thread-safe per process (same as the reference's function-local static stream,
 *    rasterization/src/forward.cu:47-53).
 */
#ifndef FGS_B200_H_
#define FGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGS_B200_ABI_VERSION 1

/* std::function<char*(size_t)> of forward.h:11-14 as a C callback */
typedef char* (*fgs_resize_fn)(void* user, size_t bytes);

/* Scalar camera / render settings: RasterizerSettings minus the three device tensors
 * (torch_bindings/rasterization.py:8-21). */
typedef struct fgs_settings {
    int32_t active_sh_bases;     /* 1, 4, 9 or 16 */
    int32_t total_sh_bases_rest; /* sh_rest.size(1) */
    int32_t width;
    int32_t height;
    float focal_x;
    float focal_y;
    float center_x;
    float center_y;
    float near_plane;
    float far_plane;
    int32_t proper_antialiasing;
} fgs_settings;

/* Opaque forward state handed to the backward (the reference's n_instances / n_buckets / selector
 * ints, rasterization_api.h:8-28). */
typedef struct fgs_forward_state {
    int32_t n_instances;
    int32_t n_buckets;   /* always 0: this implementation keeps no bucket checkpoints */
    int32_t selector;    /* always 0 */
    int32_t n_visible;
} fgs_forward_state;

int fgs_abi_version(void);
const char* fgs_last_error(void);

/* Covariance-input mode (no reference counterpart on `main`; what the FasterGS4D branch, README.md:39, needs from the
 * rasterizer): fgs_forward / fgs_inference / fgs_backward called with rotations == NULL read `scales` as (N,6) rows
 * (xx, xy, xz, yy, yz, zz) of the 3D covariance, and fgs_backward (grad_rotations == NULL) writes the (N,6) gradient of
 * those six entries to grad_scales (an off-diagonal entry counts once).  Everything else is unchanged.  Not available
 * for fgs_backward_adam. */

/* forward.h:10-36 -- differentiable forward.  Writes image (3,H,W). */
int fgs_forward(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance,
                const float* means, const float* scales, const float* rotations, const float* opacities,
                const float* sh_coefficients_0, const float* sh_coefficients_rest, const float* w2c,
                const float* cam_position, const float* bg_color, float* image, int32_t n_primitives,
                const fgs_settings* settings, void* stream, fgs_forward_state* state_out);

/* backward.h:8-44 -- writes the six gradient tensors (every row; zeros for invisible Gaussians).
 * `grad2d_scratch` is a caller-provided device buffer of fgs_grad2d_scratch_bytes(n) bytes (the
 * reference's grad_mean2d / grad_conic helpers, rasterization_api.cu:133-134); it is zeroed inside.
 * `densification_info` is (2,N) or NULL. */
size_t fgs_grad2d_scratch_bytes(int32_t n_primitives);
int fgs_backward(const float* grad_image, const float* image, const float* means, const float* scales,
                 const float* rotations, const float* opacities, const float* sh_coefficients_rest,
                 const float* w2c, const float* cam_position, const float* bg_color,
                 const char* primitive_buffers, const char* tile_buffers, const char* instance_buffers,
                 float* grad_means, float* grad_scales, float* grad_rotations, float* grad_opacities,
                 float* grad_sh_coefficients_0, float* grad_sh_coefficients_rest, void* grad2d_scratch,
                 float* densification_info, int32_t n_primitives, const fgs_settings* settings,
                 const fgs_forward_state* state, void* stream);

/* FasterGSFused semantics (README.md:37; branch not in the reference tree): backward followed by
 * FusedAdam.step on the six tensors, in one pass -- parameters / exp_avg / exp_avg_sq are updated
 * in place and no gradient tensor is materialised.  Tensor order everywhere: means, scales,
 * rotations, opacities, sh_coefficients_0, sh_coefficients_rest.  `step` is the 1-based Adam step
 * of each tensor AFTER this update (adam.py:29). */
int fgs_backward_adam(const float* grad_image, const float* image, float* const params[6],
                      float* const exp_avg[6], float* const exp_avg_sq[6], const int32_t step[6],
                      const double lr[6], double beta1, double beta2, double eps, const float* w2c,
                      const float* cam_position, const float* bg_color, const char* primitive_buffers,
                      const char* tile_buffers, const char* instance_buffers, void* grad2d_scratch,
                      float* densification_info, int32_t n_primitives, const fgs_settings* settings,
                      const fgs_forward_state* state, void* stream);

/* inference.h:9-36 -- forward without backward state; image is (3,H,W) if to_chw else (H,W,3). */
int fgs_inference(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance,
                  const float* means, const float* scales, const float* rotations, const float* opacities,
                  const float* sh_coefficients_0, const float* sh_coefficients_rest, const float* w2c,
                  const float* cam_position, const float* bg_color, float* image, int32_t n_primitives,
                  const fgs_settings* settings, int32_t to_chw, int32_t clamp_output, void* stream);

/* ---- deferred SH schedule (no reference counterpart; same results as fgs_forward / fgs_backward_adam) -----------------
 * The SH coefficients (sh0, sh_rest: 192 of the 236 parameter bytes of a Gaussian, with their Adam moments three quarters of
 * the optimizer's HBM traffic) are only needed late in the forward -- to colour the visible Gaussians -- while the front of
 * the forward (projection, both sorts, instance generation) is latency-bound and leaves HBM idle.  Two flags move the SH
 * work under that front:
 *   fgs_backward_adam_ex(..., FGS_DEFER_SH_REST): the fused backward + Adam updates means / scales / rotations / opacities /
 *     sh0 on `stream` as usual; the Adam update of sh_rest (gradient rebuilt from the view direction and dL/dcolour the
 *     backward leaves behind) runs on a low-priority side stream of the library and is "pending" afterwards.
 *     `grad2d_scratch` and `primitive_buffers` must stay allocated until the pending update has been joined.
 *   fgs_forward_ex(..., FGS_DEFER_COLOR, color_wait_event): geometry pass without SH, colours evaluated by a separate kernel
 *     after the tile sort; the pending update (and `color_wait_event`, an optional cudaEvent_t -- e.g. the arrival of the
 *     coefficients from a peer GPU) is waited for only there.
 * Every other entry point that reads or writes the coefficients joins a pending update first; code outside the library that
 * touches sh_rest or its moments calls fgs_join_deferred(stream) before (stream-ordered, no host wait). */
#define FGS_DEFER_COLOR 1u
#define FGS_DEFER_SH_REST 1u
int fgs_forward_ex(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance, const float* means,
                   const float* scales, const float* rotations, const float* opacities, const float* sh_coefficients_0,
                   const float* sh_coefficients_rest, const float* w2c, const float* cam_position, const float* bg_color,
                   float* image, int32_t n_primitives, const fgs_settings* settings, void* stream,
                   fgs_forward_state* state_out, uint32_t flags, void* color_wait_event);
int fgs_backward_adam_ex(const float* grad_image, const float* image, float* const params[6], float* const exp_avg[6],
                         float* const exp_avg_sq[6], const int32_t step[6], const double lr[6], double beta1, double beta2,
                         double eps, const float* w2c, const float* cam_position, const float* bg_color,
                         const char* primitive_buffers, const char* tile_buffers, const char* instance_buffers,
                         void* grad2d_scratch, float* densification_info, int32_t n_primitives,
                         const fgs_settings* settings, const fgs_forward_state* state, void* stream, uint32_t flags);
int fgs_join_deferred(void* stream);
/* CTAs per SM of the deferred update kernel (default 2): how much of the machine it may take from the forward it runs under */
void fgs_debug_set_deferred_ctas_per_sm(int32_t ctas);

/* adam.h:7-16 -- one dense Adam step on one tensor of n_elements floats. */
int fgs_adam_step(const float* param_grad, float* param, float* exp_avg, float* exp_avg_sq,
                  size_t n_elements, int32_t step_count, double learning_rate, double beta1, double beta2,
                  double epsilon, void* stream);

/* pruning_scores.h:9-33 -- Speedy-Splat pruning scores of one view, accumulated into scores[n_primitives]:
 * sum over pixels of (opacity * dL/dalpha)^2 with dL/d(pixel colour) == 1.  `image_scratch` is a caller
 * buffer of 3*H*W floats.  The resize callbacks must return the current buffer when called with bytes == 0. */
int fgs_pruning_scores(fgs_resize_fn resize, void* user_primitive, void* user_tile, void* user_instance,
                       const float* means, const float* scales, const float* rotations, const float* opacities,
                       const float* sh_coefficients_0, const float* sh_coefficients_rest, const float* w2c,
                       const float* cam_position, const float* bg_color, float* image_scratch, float* scores,
                       int32_t n_primitives, const fgs_settings* settings, void* stream);

/* filter3d/include/filter3d.h:7-20 -- Mip-Splatting 3D filter: running minimum over views of
 * distance2filter * depth for points inside the frustum dilated by `clipping_tolerance`;
 * visibility_mask is a byte (bool) tensor set to 1 for updated points. */
int fgs_update_3d_filter(const float* positions, const float* w2c, float* filter_3d, uint8_t* visibility_mask,
                         int32_t n_points, int32_t width, int32_t height, float focal_x, float focal_y,
                         float center_x, float center_y, float near_plane, float clipping_tolerance,
                         float distance2filter, void* stream);

/* densification/include/densification_api.h:8-21 -- 3DGS-MCMC helpers: opacity / scale of relocated
 * Gaussians (Eq. 9, at most 50 samples) and the covariance-shaped exploration noise on the means. */
int fgs_relocation_adjustment(const float* old_opacities, const float* old_scales,
                              const int64_t* n_samples_per_primitive, float* new_opacities, float* new_scales,
                              int32_t n_primitives, void* stream);
int fgs_add_noise(const float* raw_scales, const float* raw_rotations, const float* raw_opacities,
                  const float* random_samples, float* means, int32_t n_primitives, float current_lr, void* stream);

/* ---- photometric loss fused with its gradient (the step between forward and backward of a training iteration).
 * Replaces /root/reference/Loss.py:15-16,26-33 as called from Trainer.py:191-196:
 *     loss = lambda_l1 * mean|image - target| + lambda_dssim * (1 - mean SSIM(image, target))
 * with SSIM as in NeRFICG's fused_dssim / the published fused-ssim kernel (11x11 Gaussian window, sigma 1.5, zero
 * padding, C1 = 0.01^2, C2 = 0.03^2; not in the reference tree: parity pinned by oracle/loss_oracle.py only).
 * image and grad_image are (3,H,W) float32; target is (3,H,W) float32, or uint8 (value / 255) if target_is_u8.
 * `scratch` is a caller-owned device buffer of fgs_loss_scratch_bytes(width, height) bytes, 16-byte aligned (like image,
 * grad_image and a float target; a uint8 target 4-byte), that
 * must stay untouched between fgs_loss_forward and the matching fgs_loss_backward.  loss_out receives three
 * floats on the device: the weighted loss, the L1 term, the DSSIM term.  grad_loss is a DEVICE pointer to
 * dL/dloss (NULL: 1); nothing here synchronises with the host. */
size_t fgs_loss_scratch_bytes(int32_t width, int32_t height);
int fgs_loss_forward(const float* image, const void* target, int32_t target_is_u8, int32_t width, int32_t height,
                     float lambda_l1, float lambda_dssim, void* scratch, float* loss_out, void* stream);
int fgs_loss_backward(const float* image, const void* target, int32_t target_is_u8, int32_t width, int32_t height,
                      float lambda_l1, float lambda_dssim, const void* scratch, const float* grad_loss,
                      float* grad_image, void* stream);
/* forward and backward (with dL/dloss = 1) in one call: what a trainer that calls loss.backward() right away needs */
int fgs_loss_forward_backward(const float* image, const void* target, int32_t target_is_u8, int32_t width,
                              int32_t height, float lambda_l1, float lambda_dssim, void* scratch, float* loss_out,
                              float* grad_image, void* stream);

/* ---- densification / pruning / reordering of the Gaussians together with their Adam state (csrc/densify.cu).
 * Replaces the torch-op implementations in /root/reference/Model.py:276-310 (prune, sort), :312-366
 * (adaptive_density_control), :340-347,:436-443 (extend_param_groups), :406,:449 (reset_state), :459-463
 * (apply_morton_ordering); NeRFICG's Optim.adam_utils / CudaUtils.MortonEncoding behind them are not in the
 * reference tree (behaviour restated from the call sites; Morton code: 21 bits per axis of the position normalised
 * to the bounding box, x most significant).
 *
 * Everything is "build a row plan, then move rows".  A plan entry (int64, one per OUTPUT row) is
 *     bits 0..31 source row | bits 32..33 kind | bits 34..63 row of the split's normal sample
 * kind 0 = existing Gaussian (Adam state moves with it), 1 = clone (state zero), 2 / 3 = first / second child of a
 * split (mean displaced by R(q/|q|) (exp(s) * z), scale log(0.625 exp(s)), state zero).  A torch int64 index tensor
 * is therefore a valid plan (kind 0).  All pointers are device pointers unless named *_host. */
#define FGS_MAX_ROW_TENSORS 24
#define FGS_ROW_COPY 0   /* rows are copied whatever the kind (rotations, opacities, SH, auxiliary per-Gaussian data) */
#define FGS_ROW_STATE 1  /* Adam moment: copied for kind 0, zero otherwise */
#define FGS_ROW_MEANS 2  /* (N,3) means: split children are displaced */
#define FGS_ROW_SCALES 3 /* (N,3) raw log-scales: split children are shrunk by 1.6 */
typedef struct fgs_row_tensor {
    const float* src;      /* (N_in, row_floats) */
    const float* appended; /* rows for output rows >= n_copy, (n_out - n_copy, row_floats); NULL = zeros */
    float* dst;            /* (n_out, row_floats) */
    int32_t row_floats;
    int32_t role;          /* FGS_ROW_* */
} fgs_row_tensor;

/* dst[j] = row plan[j] of src for j < n_copy (plan == NULL: row j), appended[j - n_copy] (or zeros) behind; every
 * tensor in ONE launch.  split_* are only read if the plan contains kinds 2 / 3: the SOURCE scales (N_in,3),
 * rotations (N_in,4) and the (2 * n_split, 3) standard-normal samples (Model.py:332 draws them with randn_like). */
int fgs_gather_rows(const fgs_row_tensor* tensors_host, int32_t n_tensors, const int64_t* plan, int64_t n_out,
                    int64_t n_copy, const float* split_scales, const float* split_rotations,
                    const float* split_samples, void* stream);
/* tensors[t][index[k]] = 0 rows (Adam state reset, Model.py:406,449) */
int fgs_zero_rows(float* const* tensors_host, const int32_t* row_floats_host, int32_t n_tensors, const int64_t* index,
                  int64_t n_index, void* stream);

/* Order-preserving plans from one flags byte per Gaussian: bit 0 keep the Gaussian, bit 1 keep a clone of it,
 * bit 2 keep its two split children, bit 3 it is split (kept or not: fixes the sample rows).  Output order =
 * the reference's: kept originals, kept clones, first children, second children, each in index order.
 * fgs_plan_count returns (and waits for) the four totals {kept, clones, split pairs kept, splits} -- the one host
 * synchronisation, needed to allocate the outputs (n_out = kept + clones + 2 * pairs); fgs_plan_emit then writes the
 * n_out plan entries.  `scratch`: fgs_densify_scratch_bytes(n) bytes, untouched between the two calls.
 * A boolean mask (uint8 0/1) is a flags array with only bit 0: prune(mask) = count + emit + gather. */
size_t fgs_densify_scratch_bytes(int32_t n_primitives);
int fgs_plan_count(const uint8_t* flags, int32_t n_primitives, void* scratch, uint32_t totals_host[4], void* stream);
int fgs_plan_emit(const uint8_t* flags, int32_t n_primitives, const void* scratch, int64_t* plan, void* stream);
/* Model.py:314-321,359-365: flags of adaptive_density_control.  densification_info is (2,N); the thresholds are the
 * reference's host-side scalars: small_scale_log = log(percent_dense * extent), min_opacity_logit =
 * log(min_opacity / (1 - min_opacity)), large_scale_log = log(0.1 * extent) (only used if prune_large). */
int fgs_adc_flags(const float* densification_info, const float* scales, const float* rotations, const float* opacities,
                  int32_t n_primitives, float grad_threshold, float small_scale_log, float min_opacity_logit,
                  int32_t prune_large, float large_scale_log, uint8_t* flags, void* stream);

/* Morton codes of the means (any output may be NULL): the 63-bit code, its low / high 32-bit halves as radix-sort keys
 * and the identity permutation to sort along.  Two stable fgs_sort_pairs_u32 passes (low half, then -- after
 * fgs_gather_u32 of the high half into the new order -- high half) give the Morton order.
 * `scratch`: fgs_densify_scratch_bytes(n) bytes. */
int fgs_morton_keys(const float* means, int32_t n_primitives, void* scratch, int64_t* codes, uint32_t* key_lo,
                    uint32_t* key_hi, uint32_t* iota, void* stream);
/* dst[i] = src[index[i]] (dst may be NULL) and / or index_as_i64[i] = index[i] (may be NULL) */
int fgs_gather_u32(const uint32_t* src, const uint32_t* index, uint32_t* dst, int64_t* index_as_i64, int32_t n,
                   void* stream);

/* ---- time slice of 4D Gaussians (csrc/slice4d.cu; the FasterGS4D branch, README.md:39, is not in the reference tree:
 * restated from Yang et al., ICLR 2024 -- parity unpinned, oracle/fourd_oracle.py).  means4 (N,4) = (x, y, z, t), scales4
 * (N,4) log-scales, rot_l / rot_r (N,4) unnormalised quaternions of the left / right isoclinic rotation, opacities (N,1)
 * logits.  Forward: the 3D Gaussians at time t as means3 (N,3), cov3 (N,6) = (xx, xy, xz, yy, yz, zz) and opacity
 * logits (N,1) including the temporal marginal -- the inputs of fgs_forward in covariance-input mode.  Backward: the
 * gradients of the five 4D tensors from those of the three outputs. */
int fgs_slice4d_forward(const float* means4, const float* scales4, const float* rot_l, const float* rot_r,
                        const float* opacities, float t, int32_t n_primitives, float* means3, float* cov3,
                        float* opacity_logits, void* stream);
int fgs_slice4d_backward(const float* means4, const float* scales4, const float* rot_l, const float* rot_r,
                         const float* opacities, float t, int32_t n_primitives, const float* grad_means3,
                         const float* grad_cov3, const float* grad_opacity_logits, float* grad_means4,
                         float* grad_scales4, float* grad_rot_l, float* grad_rot_r, float* grad_opacities, void* stream);

/* ---- view-parallel training over NVLink peer memory (no reference counterpart: the reference is single-GPU,
 * Renderer.py:58-61).  Run and measured on 2 / 4 / 8 B200s (profiles/r2_results.md); semantics checked against
 * all-reduce + fgs_adam_step on two GPUs (tests/test_peer_adam.py).
 *
 * Every rank keeps its six gradient tensors in one flat buffer and its six parameter tensors in another, both in
 * symmetric memory with the same layout on every rank: tensor t occupies floats [offsets[t], offsets[t] + counts[t]),
 * offsets[t] a multiple of 4, the region padded with zeros to a multiple of 4 floats.  peer_grads[q] / peer_params[q]
 * are the addresses of rank q's buffers as mapped into THIS process.  One call on every rank, bracketed by
 * cross-rank barriers (all gradients written before, all parameter stores landed after), leaves on every rank the
 * parameters that "all-reduce(sum) the gradients, then fgs_adam_step on every tensor" would leave; the Adam moments
 * are sharded: rank r keeps them only for its slice of every tensor (fgs_peer_adam_slice), packed into two arrays of
 * fgs_peer_adam_shard_floats(counts, world) floats. */
#define FGS_MAX_PEERS 8
int fgs_peer_adam_step(const float* const* peer_grads, float* const* peer_params, int32_t world, int32_t rank,
                       float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6],
                       const uint64_t counts[6], const int32_t step[6], const double lr[6], double beta1,
                       double beta2, double eps, void* stream);
/* NVLS variant of fgs_peer_adam_step: `grads_multicast` / `params_multicast` are the MULTICAST addresses of the two flat
 * buffers (torch.distributed._symmetric_memory: handle.multicast_ptr; CUDA: cuMulticastCreate / cuMulticastBindMem).  The
 * gradient sum over all W replicas is formed inside the NVSwitch (multimem.ld_reduce.add.f32), the updated parameters are
 * stored once and replicated by the switch (multimem.st); `local_params` is this rank's own replica of the parameter
 * buffer (read directly).  Same slicing, same sharded moments, same barriers around the call. */
int fgs_peer_adam_step_multicast(const float* grads_multicast, float* params_multicast, float* local_params, int32_t world,
                                 int32_t rank, float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6],
                                 const uint64_t counts[6], const int32_t step[6], const double lr[6], double beta1,
                                 double beta2, double eps, void* stream);
/* Row-sliced variant with a compact gradient exchange (one view per rank and step): rank r owns the Gaussians
 * [r * R, (r + 1) * R), R = fgs_peer_adam_rows_per_rank(n_rows, world) (a multiple of 32); the peers' sh_rest gradients are
 * never read -- the owner rebuilds them from each peer's sh0 gradient (= C0 * dL/dcolour), the peer's camera position (3
 * floats at float offset `cam_offset` of its gradient buffer, written by the caller before the barrier) and the Gaussian's
 * mean, so 14 instead of 59 floats per Gaussian and peer cross NVLink.  `offsets[t]`: float offset of tensor t (means,
 * scales, rotations, opacities, sh0, sh_rest; rows of 3, 3, 4, 1, 3, 3 * n_rest floats) in the flat buffers; moment shards:
 * R * 59-ish floats, tensor after tensor.  `broadcast_rest` = 0 keeps the new sh_rest parameters in this rank's replica
 * only: the caller copies the slice to the peers itself (e.g. on a side stream, overlapped with the next forward's
 * geometry pass, whose colour kernel waits for it: fgs_forward_ex).  Same result as fgs_peer_adam_step up to the rounding
 * of dL/dcolour = (C0 * dL/dcolour) / C0. */
uint64_t fgs_peer_adam_rows_per_rank(uint64_t n_rows, int32_t world);
int fgs_peer_adam_rows_step(const float* const* peer_grads, float* const* peer_params, int32_t world, int32_t rank,
                            float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6], uint64_t n_rows,
                            int32_t n_rest, int32_t active_sh_bases, uint64_t cam_offset, int32_t broadcast_rest,
                            const int32_t step[6], const double lr[6], double beta1, double beta2, double eps,
                            const float* grads_multicast /* NVLS: both multicast addresses, or both NULL */,
                            float* params_multicast,
                            const float* means_old /* non-NULL: sh_rest only, means from this snapshot (split step, below) */,
                            int32_t max_ctas_per_sm /* 0: a full wave; 1 or 2: run next to another stream's kernels */,
                            void* stream);
/* The split compact step: fgs_peer_adam_small_step sums the five small gradients (means, scales, rotations, opacities, sh0)
 * over the peers, applies Adam to this rank's Gaussians and stores the new parameters to every peer with the float4 kernel
 * of fgs_peer_adam_step (row-aligned slices, moment layout of the row-sliced step), and leaves the PRE-update means of the
 * rank's Gaussians in `means_old` (3 * R floats).  fgs_peer_adam_rows_step(..., means_old, ...) then handles sh_rest alone --
 * rebuild the gradient, Adam, broadcast -- and can run on a side stream under the next forward's geometry pass, whose
 * colour kernel waits for it (fgs_forward_ex). */
int fgs_peer_adam_small_step(const float* const* peer_grads, float* const* peer_params, int32_t world, int32_t rank,
                             float* exp_avg_shard, float* exp_avg_sq_shard, const uint64_t offsets[6], uint64_t n_rows,
                             int32_t n_rest, const int32_t step[6], const double lr[6], double beta1, double beta2,
                             double eps, float* means_old,
                             const float* grads_multicast /* NVLS: both multicast addresses, or both NULL */,
                             float* params_multicast, void* stream);
/* floats [first_float, first_float + n_floats) of this rank's buffer to the same range of every replica through the
 * buffer's multicast address (one multimem.st per 16 bytes); `max_ctas` bounds the grid (default 64) */
int fgs_peer_broadcast_multicast(const float* local_buffer, float* multicast_buffer, uint64_t first_float, uint64_t n_floats,
                                 int32_t max_ctas, void* stream);
size_t fgs_peer_adam_shard_floats(const uint64_t counts[6], int32_t world);
/* rank's slice of every tensor: first float (relative to the tensor), number of floats (padding included),
 * position of the slice's moments inside the shard arrays; any output pointer may be NULL */
int fgs_peer_adam_slice(const uint64_t counts[6], int32_t world, int32_t rank, uint64_t begin_floats[6],
                        uint64_t count_floats[6], uint64_t shard_offset_floats[6]);
const char* fgs_peer_last_error(void);

/* Optional per-kernel timing for bench.py: while enabled every kernel (group) launched by the entry
 * points above is bracketed by CUDA events on the caller's stream.  fgs_profile_read waits for them
 * and writes one "name=milliseconds@start_offset_ms\n" line per launch, in launch order, then clears the list. */
void fgs_profile_enable(int32_t on);
int fgs_profile_read(char* out, size_t out_bytes);

/* fgs_forward / fgs_inference size the instance buffers with the capacity the previous call suggested and
 * check their own instance count afterwards (one re-run if it did not fit).  These two let the unit tests force
 * the re-run path (a tiny hint) or the cold path (0). */
void fgs_debug_set_capacity_hint(int32_t capacity);
int32_t fgs_debug_get_capacity_hint(void);
/* number of forwards on the current device whose speculative tail did not fit and was queued a second time */
int32_t fgs_debug_get_forward_reruns(void);
/* instrumentation builds (-DFGS_BLEND_STATS=1) only: the blend kernels' pair / useful-lane counters (layout in
 * csrc/fgs_device.cuh); returns -1 in the product build */
int fgs_debug_blend_stats(uint64_t out[16], int32_t reset);

/* EXPERIMENTAL schedule of fgs_backward_adam (default off; also FGS_OVERLAP_ADAM=1 in the environment): the
 * zero-gradient Adam update of the sh_rest rows of Gaussians that were not visible runs on a private side stream
 * concurrently with the blend backward.  Results are bit-identical to the default schedule. */
void fgs_debug_set_overlap_adam(int32_t on);
int32_t fgs_debug_get_overlap_adam(void);

/* Building blocks exported for the unit tests (each is one stage of fgs_forward). */
int fgs_sort_pairs_u32(uint32_t* keys, uint32_t* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                       const uint32_t* n_dev, int32_t n_max, int32_t end_bit, void* scratch,
                       size_t scratch_bytes, int32_t* result_in_alt, void* stream);
/* same with 16-bit keys (the instantiation used for the tile sort when the tile grid has <= 65536 tiles) */
int fgs_sort_pairs_u16(uint16_t* keys, uint16_t* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                       const uint32_t* n_dev, int32_t n_max, int32_t end_bit, void* scratch,
                       size_t scratch_bytes, int32_t* result_in_alt, void* stream);
size_t fgs_sort_scratch_bytes(int32_t n_max, int32_t end_bit);

#ifdef __cplusplus
}
#endif
#endif /* FGS_B200_H_ */
