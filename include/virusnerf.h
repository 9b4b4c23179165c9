/* virusnerf.h -- C ABI of libvirusnerf_sm100.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (nas-git-nas/VIRUS-NeRF) has no FFI layer: its hot path is Python modules
 * (modules/*.py) wrapping Taichi JIT kernels.  Each entry point below replaces one of those
 * kernels / torch op sequences; the reference interface it replaces is cited as file:line
 * (paths relative to the reference checkout).  The Python host side that mirrors the
 * reference's module API on top of this ABI lives in virus-nerf_b200/modules/.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter name starts with h_ (host);
 *    the caller (PyTorch) owns and sizes every buffer, kernels never allocate;
 *  - all tensors are dense row-major ("contiguous"), dtypes as in the signatures;
 *  - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL =
 *    the legacy default stream); no call synchronises the host;
 *  - return value: 0 on success, a negative VN_E* code otherwise; vn_last_error() returns a
 *    thread-local human-readable description of the last failure;
 *  - no C++ exceptions cross this boundary.
 */
#ifndef VIRUSNERF_H_
#define VIRUSNERF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VN_ABI_VERSION 1
#define VN_MAX_LEVELS 32
#define VN_MAX_SAMPLES 1024 /* modules/utils.py:12 */

#define VN_OK 0
#define VN_EINVAL -1  /* bad argument (null pointer, size, alignment, unsupported config) */
#define VN_ELAUNCH -2 /* CUDA launch / runtime failure; see vn_last_error() */

/* flags for the hash-encoder kernels (tuning switches, results identical within tolerance) */
#define VN_HASH_DEFAULT 0
#define VN_HASH_NO_WARP_AGG 1     /* bwd: plain per-corner atomics, no warp pre-reduction */
#define VN_HASH_LEVEL_GROUPS_1 16  /* one level per thread (grid.y = levels): level-major */
#define VN_HASH_LEVEL_GROUPS_4 32  /* four levels per thread */
#define VN_HASH_LEVEL_GROUPS_8 64  /* eight levels per thread */
#define VN_HASH_LEVEL_GROUPS_16 128 /* all (<= 16) levels per thread: every row touched once */
#define VN_HASH_LEVEL_GROUPS_2 256 /* two levels per thread */
#define VN_HASH_PLANAR 512 /* f32 only: out / dout are [levels/2][S] float4 planes (level pair
                              p = levels 2p, 2p+1 of every point) instead of [S, 2*levels] rows;
                              the layout vn_mlp_fwd/bwd read and write with enc_format = 2 */
#define VN_HASH_TIGHT_REGS 1024 /* planar bwd: 48-register variant, 5 CTAs per SM */
#define VN_HASH_SKIP_ZERO_GRADS 4096 /* f32 bwd: do not scatter (level, sample) pairs whose gradient is exactly 0 */
#define VN_HASH_PAIR_LOADS 2048 /* planar fwd: 16-byte loads for x / x+1 corner pairs that are neighbours */
#define VN_HASH_FUSED_SCATTER 16384 /* native step runner: MLP backward and hash backward run as one kernel
                                       (vn_mlp_bwd_scatter); needs VN_HASH_F16_CHUNKS */
#define VN_HASH_F16_CHUNKS 8192 /* fwd (f32 or f16 table): out is [levels/4][S] x 16 B "chunk planes": plane c holds
                                  levels 4c..4c+3 of every point as 8 fp16 values = one row of one column chunk of the
                                  fused MLP's tensor-core operand (vn_mlp_fwd enc_format 3, vn_mlp_bwd 3 / 4); bwd
                                  (vn_hash_encode_bwd_f16): dout is in the same layout */
/* default (no GROUPS flag): 4 (backward: 2 when the table exceeds the L2, > 96 MB) */

const char* vn_last_error(void);
int vn_abi_version(void);
/* number of kernels this library has launched so far in this process */
int64_t vn_launch_count(void);
/* in-stream kernel timing: while enabled, the launchers of the major kernels bracket their
 * launch with CUDA events on the launching stream; _get synchronises on record i and returns
 * (kernel id, problem size = samples / rays / params, milliseconds).  ids: 0 hash fwd, 1 hash
 * bwd, 2 MLP fwd, 3 MLP bwd, 4 march count, 5 march write, 6 composite fwd, 7 composite bwd,
 * 8 Adam. */
int vn_profile_enable(int on);
/* the same restricted to the kernel ids whose bit is set in `mask` (bit k = kernel id k) */
int vn_profile_enable_mask(unsigned mask);
/* programmatic dependent launch between the kernels of the train step (default on; VN_PDL=0 in the
 * environment or vn_set_pdl(0) launches them with plain stream order) */
int vn_set_pdl(int on);
int vn_profile_count(void);
int vn_profile_get(int i, int* h_kernel_id, int64_t* h_size, float* h_ms);
/* SM count / device name of the current device; proves the library talks to a GPU */
int vn_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* ---------------------------------------------------------------------------------------
 * a1. Hash-grid level geometry -- HashEncoder.__init__, modules/hash_encoder.py:149-235 with
 * helpers modules/utils.py:19-42.  Host-only.  `scales`/`res` are the per-level constants
 * the Taichi kernel derives in f32 (hash_encoder.py:73-80), precomputed once here.
 * ------------------------------------------------------------------------------------- */
typedef struct vn_hash_levels {
    int32_t levels;
    int32_t begin_fast_hash_level;   /* first level that uses the xor hash */
    int32_t offsets[VN_MAX_LEVELS];  /* entry offset of each level in the table */
    int32_t sizes[VN_MAX_LEVELS];    /* hash_map_sizes */
    float scales[VN_MAX_LEVELS];     /* base_res * exp(l * log_b) - 1 (f32) */
    uint32_t res[VN_MAX_LEVELS];     /* ceil(scale) + 1 */
    int64_t total_entries;           /* sum(sizes); hash_table has 2 * total_entries floats */
    double log_b;
} vn_hash_levels_t;

int vn_hash_levels_init(double base_res, double max_res, int levels, int64_t max_params,
                        vn_hash_levels_t* h_out);

/* a2. hash_encoder_kernel, modules/hash_encoder.py:89-143 (feature_per_level = 2).
 * xyz [S,3] f32 in [0,1]; table [2*total_entries] f32; out [S, 2*levels] f32. */
int vn_hash_encode_fwd_f32(const float* xyz, const float* table, float* out, int64_t S,
                           const vn_hash_levels_t* h_lv, int flags, void* stream);

/* a3. hash_encoder_kernel.grad via _module_function.backward, hash_encoder.py:264-277.
 * grad [2*total_entries] f32 is ACCUMULATED into (caller zeroes it). dout [S, 2*levels]. */
int vn_hash_encode_bwd_f32(const float* xyz, const float* dout, float* grad, int64_t S,
                           const vn_hash_levels_t* h_lv, int flags, void* stream);

/* the same restricted to levels [level_begin, level_end): lets a data-parallel caller start
 * the allreduce of a finished level slab while the remaining levels are still scattering */
int vn_hash_encode_bwd_f32_levels(const float* xyz, const float* dout, float* grad, int64_t S,
                                  const vn_hash_levels_t* h_lv, int flags, int level_begin,
                                  int level_end, void* stream);

/* a4. half encoder, modules/hash_encoder_half.py:112-161 (fwd) and :164-213 (bwd).
 * table_h [total_entries,2] fp16 (the per-call hash_table.to(float16) copy, :367);
 * out_h [S, levels, 2] fp16; dout_h same shape; grad [total_entries,2] f32 (hash_grad, :300). */
int vn_hash_encode_fwd_f16(const float* xyz, const void* table_h, void* out_h, int64_t S,
                           const vn_hash_levels_t* h_lv, int flags, void* stream);
int vn_hash_encode_bwd_f16(const float* xyz, const void* dout_h, float* grad, int64_t S,
                           const vn_hash_levels_t* h_lv, int flags, void* stream);
/* f32 master table -> fp16 copy (hash_encoder_half.py:367), n floats */
int vn_f32_to_f16(const float* src, void* dst_h, int64_t n, void* stream);

/* known-answer helper: the 8 level-local corner indices [S,levels,8] i32 and weights
 * [S,levels,8] f32 (weights may be NULL) -- hash_encoder.py:43-71, 114-133 */
int vn_hash_indices(const float* xyz, int64_t S, const vn_hash_levels_t* h_lv, int32_t* idx,
                    float* w, void* stream);

/* a5. ray_aabb_intersect, modules/intersection.py:8-37.  hits_t [N,2]. */
int vn_ray_aabb(const float* rays_o, const float* rays_d, float scale, int64_t N, float* hits_t,
                void* stream);

/* a6. raymarching_train_kernel, modules/ray_march.py:9-124, split at the atomic counter:
 *  _count: pass 1 (:29-82).  Writes counts [N] i32, then the deterministic equivalent of the
 *          atomic counter: rays_a [N,3] i32 = (r, exclusive scan of counts in ray order, count)
 *          and counter [2] i32 = (total samples, N).  scan_tmp: >= vn_march_scan_tmp_ints(N) i32.
 *  _write: pass 2 (:84-124).  Re-marches and writes xyzs, dirs [*,3], deltas, ts [*] rows
 *          rays_a[r,1] .. +count; rows >= capacity are dropped (capacity = rows allocated).
 *          xyzs_unit (may be NULL): the same positions mapped to the unit cube,
 *          (x - xyz_min) / (xyz_max - xyz_min) of NGP.density (networks.py:142). */
int64_t vn_march_scan_tmp_ints(int64_t N);
int vn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t,
                         const uint8_t* bitfield, const float* noise, int64_t N, int cascades,
                         int grid_size, float scale, float exp_step_factor, int max_samples,
                         int32_t* counts, int32_t* rays_a, int32_t* counter, int32_t* scan_tmp,
                         void* stream);
int vn_march_train_write(const float* rays_o, const float* rays_d, const float* hits_t,
                         const uint8_t* bitfield, const float* noise, int64_t N, int cascades,
                         int grid_size, float scale, float exp_step_factor, const int32_t* rays_a,
                         int64_t capacity, float* xyzs, float* dirs, float* deltas, float* ts,
                         float* xyzs_unit, void* stream);
/* single-pass variant of a6 used by the native step runner: _count_rows = _count that also
 * stores the t of emitted sample k of ray r in ts_rows[r * max_samples + k]; _expand = pass 2
 * without a second march: every output of :84-124 is a function of (ray, t) -- xyz = o + t d
 * (:45), dt = calc_dt(t) (:46) -- evaluated with the same IEEE operations, so the result is
 * bit-identical to _write. */
int vn_march_train_count_rows(const float* rays_o, const float* rays_d, const float* hits_t,
                              const uint8_t* bitfield, const float* noise, int64_t N, int cascades,
                              int grid_size, float scale, float exp_step_factor, int max_samples,
                              int32_t* counts, int32_t* rays_a, int32_t* counter,
                              int32_t* scan_tmp, float* ts_rows, void* stream);
int vn_march_train_expand(const float* rays_o, const float* rays_d, const int32_t* rays_a,
                          const float* ts_rows, int64_t N, int max_samples, int grid_size,
                          float scale, float exp_step_factor, int64_t capacity, float* xyzs,
                          float* dirs, float* deltas, float* ts, float* xyzs_unit, void* stream);
/* the same, and additionally the direction encoding SH16((d/|d|+1)/2) of every sample (a function of the ray,
 * evaluated once per ray: networks.py:160-161, spherical_harmonics.py:16-42) as two fp16 operand-chunk planes
 * sh_planes[0][s], sh_planes[1][s] (16 B each, plane stride sh_stride samples) -- planes 4 and 5 of the fused MLP's
 * enc_format 5 input */
int vn_march_train_expand_sh(const float* rays_o, const float* rays_d, const int32_t* rays_a,
                          const float* ts_rows, int64_t N, int max_samples, int grid_size,
                          float scale, float exp_step_factor, int64_t capacity, float* xyzs,
                          float* dirs, float* deltas, float* ts, float* xyzs_unit, void* sh_planes, int64_t sh_stride, void* stream);

/* a7. raymarching_test_kernel, modules/ray_march.py:198-269.  alive [A] i64; slot layout
 * n*max_samples+s; ray_indices i64, valid_mask u8 (caller zeroes), deltas/ts f32, all
 * [A*max_samples]; counter [A] i32; hits_t [N,2] is updated in place (:258). */
int vn_march_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive,
                  int64_t A, const uint8_t* bitfield, int cascades, int grid_size, float scale,
                  float exp_step_factor, int max_samples, int64_t* ray_indices,
                  uint8_t* valid_mask, float* deltas, float* ts, int32_t* counter, void* stream);
/* the wrapper's compaction (ray_march.py:328-335) without boolean-mask host syncs:
 * packed_info [A,2] i64 = (exclusive cumsum, count); compacted ray_indices/deltas/ts are
 * written densely in slot order; total [1] i64 receives the number of valid samples. */
int vn_march_test_compact(const int32_t* counter, int64_t A, int max_samples,
                          const int64_t* ray_indices, const float* deltas, const float* ts,
                          int64_t* packed_info, int64_t* ray_indices_out, float* deltas_out,
                          float* ts_out, int64_t* total, int32_t* scan_tmp, void* stream);

/* a8. volume_rendering_kernel, modules/volume_train.py:6-48.  rays_a [N,3] i32; sigmas,
 * deltas, ts, ws [S]; rgbs [S,3]; outputs indexed by ray id: total_samples [N] i32,
 * opacity, depth [N], rgb [N,3].  ws of skipped samples is written 0. */
int vn_composite_train_fwd(const float* sigmas, const float* rgbs, const float* deltas,
                           const float* ts, const int32_t* rays_a, int64_t N, int64_t S,
                           float T_threshold, int32_t* total_samples, float* opacity, float* depth,
                           float* rgb, float* ws, void* stream);
/* a9. volume_rendering_kernel.grad, modules/volume_train.py:130-175.  dL_dws may be NULL.
 * Writes dsigmas [S], drgbs [S,3] (every row, zeros where skipped). */
int vn_composite_train_bwd(const float* sigmas, const float* rgbs, const float* deltas,
                           const float* ts, const int32_t* rays_a, int64_t N, int64_t S,
                           float T_threshold, const float* dL_dopacity, const float* dL_ddepth,
                           const float* dL_drgb, const float* dL_dws, float* dsigmas, float* drgbs,
                           void* stream);
/* a10. composite_test, modules/volume_render_test.py:4-54.  pack_info [A,2] i64,
 * alive [A] i64 (entries set to -1 when the ray is finished); opacity/depth/rgb in place. */
int vn_composite_test(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                      const int64_t* pack_info, int64_t* alive, int64_t A, float T_threshold,
                      float* opacity, float* depth, float* rgb, void* stream);

/* a11. dir_encoder, modules/spherical_harmonics.py:7-42.  dirs [B,3] -> emb [B,16]. */
int vn_sh_encode(const float* dirs, int64_t B, float* emb, void* stream);

/* a15. morton3D / morton3D_invert / packbits, modules/utils.py:120-169 */
int vn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, void* stream);
int vn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, void* stream);
int vn_packbits(const float* grid, int64_t n_bytes, float threshold, uint8_t* bitfield,
                void* stream);

/* a14. OccupancyGrid, modules/occupancy_grid.py.
 *  _calc_pos_prob: _calcPos (:293-335, with helpers/geometric_fcts.py:151-171 and _c2idx
 *     :479-480) fused with _rayProb (:338-389) when meas != NULL.  noise [N,M,3] uniform
 *     [0,1) or NULL.  Outputs (any may be NULL): cell_dists [N,M], cell_pos [N*M,3],
 *     cell_idxs [N*M,3] i32, probs_occ / probs_emp [N,M].
 *  _nerf_prob: _nerfProb (:392-408) given densities; mean/threshold computed on device
 *     (scratch: >= 514 floats, 8-byte aligned; [512] = mean, [513] = h_thr on exit), no host
 *     sync (the reference does torch.mean(...).item(), :402).
 *  _bayes_update: _updateGrid (:411-430) with deterministic last-writer-wins for duplicate
 *     cells.  winner: int32 [G^3] scratch that must be all -1 on entry and is restored.
 *  _decay_pack: update() tail (:96-105) = optional in-place decay + cartesian2morton
 *     (grid.py:165-170) + packbits (grid.py:205-211) in one pass. */
int vn_occ_calc_pos_prob(const float* rays_o, const float* rays_d, const float* noise,
                         const float* meas, int64_t N, int M, int I, int grid_size, float scale,
                         float noise_every_m, float p_false, float std_every_m, float prob_min,
                         float* cell_dists, float* cell_pos, int32_t* cell_idxs, float* probs_occ,
                         float* probs_emp, void* stream);
/* _rayProb (:338-389) on explicit distances dists [N,M] */
int vn_occ_ray_prob(const float* meas, const float* dists, int64_t N, int M, int I, float p_false,
                    float std_every_m, float prob_min, float* probs_occ, float* probs_emp,
                    void* stream);
/* the same with return_probs=True (:387-388): terms [4][N,M] = P[meas=dist|emp], P[meas=dist|occ],
 * P[meas not< dist|emp], P[meas not< dist|occ] (the plotting / analysis path of the reference) */
int vn_occ_ray_prob_terms(const float* meas, const float* dists, int64_t N, int M, int I, float p_false,
                          float std_every_m, float prob_min, float* probs_occ, float* probs_emp,
                          float* terms, void* stream);
int vn_occ_nerf_prob(const float* density, int64_t n, double thr_max, float slope, float* scratch,
                     float* probs_occ, float* probs_emp, void* stream);
int vn_occ_bayes_update(float* grid, int grid_size, const int32_t* cell_idxs, int64_t n,
                        const float* probs_occ, const float* probs_emp, int32_t* winner,
                        float* new_probs_tmp, void* stream);
int vn_occ_decay_pack(float* grid, int grid_size, float decay, int apply_decay, float threshold,
                      uint8_t* bitfield, void* stream);
/* OccupancyGrid.update (:65-105) after its two batches have been sampled, as ONE host call: depth-sensor update of
 * N_ray rays (r_rays_o / r_rays_d [N_ray,3], r_meas [N_ray], no NaN), NeRF update of N_nerf rays (n_noise [N_nerf,M,3]
 * uniform [0,1): the torch.rand of :326) with NGP.density (networks.py:134-148) evaluated as (x - xyz_min) /
 * (xyz_max - xyz_min) -> hash forward (f32 table, or -- when table_h != NULL -- the half-precision encoder on a fresh
 * fp16 copy of the table, hash_encoder_half.py:367) -> density-only fused MLP (W1 [64,32], W2 [16,64]); then the warm-up
 * decay (apply_decay) and the bitfield repack.  The same kernels in the same order as vn_occ_calc_pos_prob /
 * vn_occ_bayes_update / vn_hash_encode_fwd_* / vn_mlp_fwd / vn_occ_nerf_prob / vn_occ_decay_pack called one by one:
 * the grid is bit-identical.  winner: i32 [grid_size^3] filled with -1 (restored on return); ws: caller-owned scratch of
 * at least vn_occ_update_ws_floats(N_ray, N_nerf, M) floats, 16-byte aligned. */
int64_t vn_occ_update_ws_floats(int64_t N_ray, int64_t N_nerf, int M);
int vn_occ_update(float* grid, int grid_size, uint8_t* bitfield, int32_t* winner, const float* r_rays_o,
                  const float* r_rays_d, const float* r_meas, int64_t N_ray, const float* n_rays_o,
                  const float* n_rays_d, const float* n_noise, int64_t N_nerf, int M, int I, float scale,
                  float noise_every_m, float p_false, float std_every_m, float prob_min, double nerf_thr_max,
                  float nerf_slope, float decay, int apply_decay, float threshold, const float* table,
                  void* table_h, const vn_hash_levels_t* lv, int hash_flags, const float* W1, const float* W2,
                  float xyz_min, float xyz_max, float* ws, int64_t ws_floats, void* stream);

/* f2 (caller side). training/loss.py:34-198 as two kernels around the (optional) allreduce of
 * the valid counts.  Per-ray inputs: rgb [N,3] (composite output, before background), opacity,
 * depth [N]; targets gt_rgb [N,3], uss / tof / rgbd [N] (NaN = no measurement; any may be NULL).
 *  _fwd: accumulates sums[4] / counts[4] (order: colour, USS, ToF, RGBD) -- caller zeroes both.
 *        pred = rgb + bg * (1 - opacity) (rendering.py:219-226); USS counts only rays with
 *        depth < uss - uss_tol (loss.py:186-194).
 *  _bwd: with the (global) counts, writes the loss-scaled gradient seeds dL/drgb [N,3],
 *        dL/ddepth [N], dL/dopacity [N] and loss_out[0] = sum_k weights[k] * sums[k]/counts[k]
 *        (this rank's share of the global loss).  scale_dev = GradScaler scale [1] (device). */
int vn_loss_fwd(const float* rgb, const float* opacity, const float* depth, const float* gt_rgb,
                const float* uss, const float* tof, const float* rgbd, int64_t N, float bg,
                float uss_tol, float* sums, float* counts, void* stream);
int vn_loss_bwd(const float* rgb, const float* opacity, const float* depth, const float* gt_rgb,
                const float* uss, const float* tof, const float* rgbd, int64_t N, float bg,
                float uss_tol, const float* sums, const float* counts, float w_color, float w_uss,
                float w_tof, float w_rgbd, const float* scale_dev, float* dL_drgb, float* dL_ddepth,
                float* dL_dopacity, float* loss_out, void* stream);
/* a8 + f2 / a9 + f2 fused (what the native step runner enqueues): vn_composite_train_fwd that also accumulates sums[4] /
 * counts[4] of its rays' loss terms (one block-level reduction, then atomics -- vn_loss_fwd without a launch of its own),
 * and vn_composite_train_bwd that forms the per-ray gradient seeds itself from rgb / opacity / depth [N,*] (the forward's
 * outputs), the targets and the (global) counts -- vn_loss_bwd's arithmetic -- and writes loss_out[0].  Same outputs as
 * the separate calls (the sums differ in the order of their floating-point additions). */
int vn_composite_loss_fwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                          const int32_t* rays_a, int64_t N, int64_t S, float T_threshold,
                          int32_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                          const float* gt_rgb, const float* uss, const float* tof, const float* rgbd, float bg,
                          float uss_tol, float* sums, float* counts, void* stream);
int vn_composite_loss_bwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                          const int32_t* rays_a, int64_t N, int64_t S, float T_threshold, const float* rgb,
                          const float* opacity, const float* depth, const float* gt_rgb, const float* uss,
                          const float* tof, const float* rgbd, float bg, float uss_tol, const float* sums,
                          const float* counts, float w_color, float w_uss, float w_tof, float w_rgbd,
                          const float* scale_dev, float* dsigmas, float* drgbs, float* loss_out, void* stream);

/* f1 (caller side). GradScaler.unscale_ + inf check + torch.optim.Adam(eps=1e-15) step,
 * training/trainer.py:49-57,138-141, as one pass.  found_inf [1] f32 (device): set to 1 by
 * vn_grad_check when a scaled gradient is non-finite; the step kernel skips the update when
 * it is non-zero (GradScaler.step semantics).  step is the 1-based Adam step count.  When
 * scale_dev (device, [1] f32) is non-NULL the unscale factor is 1 / scale_dev[0] (the
 * GradScaler's device-side scale) and inv_scale is ignored.
 * The hyper-parameters are doubles: torch derives 1 - beta1, 1 - beta2, lr / bias_correction1 and
 * sqrt(bias_correction2) in python doubles and rounds ONCE to f32 when the scalar meets the f32
 * tensor (1 - beta2 = 0.001f, not 1.0f - 0.999f); vn_adam_config (host only) returns those six f32
 * constants (beta2, 1-beta1, 1-beta2, eps, step_size, bias_correction2_sqrt) for inspection. */
int vn_grad_check(const float* g, int64_t n, float* found_inf, void* stream);
int vn_adam_config(double lr, double beta1, double beta2, double eps, int step, float* h_out6);
int vn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float inv_scale, double lr,
                 double beta1, double beta2, double eps, int step, const float* found_inf,
                 const float* scale_dev, void* stream);
/* GradScaler.update() on device (torch _amp_update_scale_): scale [1] f32, growth_tracker
 * [1] i32; found_inf is reset to 0 afterwards. */
int vn_scaler_update(float* scale, int32_t* growth_tracker, float* found_inf, float growth_factor,
                     float backoff_factor, int growth_interval, void* stream);
/* Optimiser step count on the device (torch's GradScaler.step skips optimizer.step() on an overflow, so Adam's
 * state['step'] only advances on applied steps -- training/trainer.py:140-141): opt_state = 4 floats {lr / (1 - beta1^t),
 * sqrt(1 - beta2^t) for the NEXT step t, bit pattern of the int32 count of applied steps, unused}.
 * vn_opt_state_init sets it for `applied_steps` steps already taken; vn_adam_step_dev takes its bias corrections from
 * it; vn_scaler_update_dev advances it when found_inf == 0 (double-precision arithmetic, one thread). */
int vn_opt_state_init(float* opt_state, int applied_steps, double lr, double beta1, double beta2, void* stream);
int vn_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1,
                     double beta2, double eps, const float* opt_state, const float* found_inf,
                     const float* scale_dev, void* stream);
int vn_scaler_update_dev(float* scale, int32_t* growth_tracker, float* found_inf, float growth_factor,
                         float backoff_factor, int growth_interval, float* opt_state, double lr,
                         double beta1, double beta2, void* stream);

/* a12 (+a11 fused). NGP.density / NGP.forward dense part, modules/networks.py:134-164 and
 * MLP.forward :271-282, with DirEncoder (spherical_harmonics.py:7-42) fused into the input
 * stage: ONE kernel on the tcgen05 tensor cores (fp16 operands, fp32 accumulation in TMEM --
 * the reference runs these layers as fp16 cuBLAS GEMMs under torch.autocast, trainer.py:104).
 *   enc [S,32] hash encoding, f32 rows (enc_format = 0), fp16 rows (1) or f32 level-pair planes
 *   [8][S] float4 (2, the VN_HASH_PLANAR layout; denc is then written in the same layout), or fp16 chunk
 *   planes [4][S] x 16 B (3, the VN_HASH_F16_CHUNKS layout: tiles are bulk-copied straight into the tensor-core
 *   operand), or the same followed by the two SH planes of vn_march_train_expand_sh ([6][S] x 16 B, enc_format 5:
 *   the whole network input arrives in operand layout, dirs is ignored).  vn_mlp_bwd writes denc as f32 planes for
 *   enc_format 3 / 5 and, with VN_MLP_DENC_F16 (8) added to enc_format, as fp16 chunk planes [4][S]; dirs [S,3] raw ray
 *   directions (normalised and mapped to (d+1)/2 inside, networks.py:160-161); W1 [64,32],
 *   W2 [16,64], W3 [64,32], W4 [64,64], W5 [3,64] f32 in torch Linear layout [out,in].
 * _fwd: sigmas [S] = exp(h0), rgbs [S,3] = sigmoid(...), optional h_out [S,16] (return_feat).
 *       density_only != 0 evaluates only W1, W2 (NGP.density; dirs/W3-5/rgbs may be NULL).
 * _bwd: recomputes the forward per tile; inputs dsigmas [S], drgbs [S,3]; writes denc [S,32]
 *       f32 (gradient w.r.t. the encoding) and ACCUMULATES dW1..dW5 (same shapes as W, f32;
 *       TruncExp backward clamps h0 to [-15,15], networks.py:28). */
#define VN_MLP_DENC_F16 8
int vn_mlp_fwd(const void* enc, int enc_format, const float* dirs, const float* W1, const float* W2,
               const float* W3, const float* W4, const float* W5, int64_t S, int density_only,
               float* sigmas, float* rgbs, float* h_out, void* stream);
int vn_mlp_bwd(const void* enc, int enc_format, const float* dirs, const float* W1, const float* W2,
               const float* W3, const float* W4, const float* W5, int64_t S, int density_only,
               const float* dsigmas, const float* drgbs, float* denc, float* dW1, float* dW2,
               float* dW3, float* dW4, float* dW5, void* stream);
/* a12 + a3 / a4 fused: vn_mlp_bwd followed by vn_hash_encode_bwd_* as ONE kernel (enc_format 3 or 5, 16 levels x 2
 * features).  The gradient w.r.t. the encoding never leaves the SM: every thread of the MLP epilogue reads its levels of
 * d(enc) back from tensor memory and runs the warp-aggregated scatter of hash_encoder.py:264-277 /
 * hash_encoder_half.py:164-213 (zero-skip rule, f32 red.global.add) into table_grad [total_entries,2], ACCUMULATING like
 * vn_hash_encode_bwd_f32.  xyz [S,3] = the unit-cube positions the forward encoded; round_f16 != 0 rounds d(enc) to
 * fp16 first (the half-precision encoder's gradient dtype).  Equivalent to vn_mlp_bwd(enc_format [+ VN_MLP_DENC_F16])
 * + vn_hash_encode_bwd_f32 / _f16 up to the summation order of the atomics.
 * found_inf (optional, device float[1]): set to 1.0 when any d(enc) value or weight-gradient entry of this launch is not
 * finite -- GradScaler's inf check (training/trainer.py:140, vn_grad_check) evaluated on the contributions instead of a
 * separate pass over the 45.7 MB gradient: |d(enc)| <= 64 * 65504^2 when finite, so a sum of fewer than 2^31 finite
 * contributions cannot overflow and "some gradient entry is non-finite" <=> "some contribution is non-finite".  Never
 * cleared here. */
int vn_mlp_bwd_scatter(const void* enc, int enc_format, const float* dirs, const float* W1, const float* W2,
                       const float* W3, const float* W4, const float* W5, int64_t S, const float* dsigmas,
                       const float* drgbs, const float* xyz, const vn_hash_levels_t* lv, int round_f16,
                       float* table_grad, float* dW1, float* dW2, float* dW3, float* dW4, float* dW5,
                       float* found_inf, void* stream);

/* ---------------------------------------------------------------------------------------
 * Native step runner (caller side): the body of Trainer.train()'s loop
 * (training/trainer.py:104-141) as TWO host calls that enqueue every kernel of a step on
 * `stream` with plain C++ launch overhead instead of one Python -> ctypes round trip per
 * kernel.  All buffers are caller-owned device memory; `S` is the sample total read back from
 * counter[0] after _prepare.
 *  _prepare: vn_ray_aabb + vn_march_train_count (no dependence on gradients in flight).
 *  _run    : cudaMemsetAsync(grad, loss accumulators) + march write + hash fwd + fused MLP fwd +
 *            composite fwd + loss fwd + loss bwd + composite bwd + fused MLP bwd + hash bwd,
 *            and, when do_optim != 0, grad check + Adam + scaler update.  A data-parallel
 *            caller passes do_optim = 0, allreduces `flat_g` (and, between two _run calls split
 *            at the loss, the counts) and then calls vn_train_step_optim. */
typedef struct vn_step {
    /* rays (N) */
    int64_t N;
    const float *rays_o, *rays_d, *noise, *gt_rgb, *uss, *tof, *rgbd; /* targets may be NULL */
    float *hits_t; int32_t *counts, *rays_a, *counter, *scan_tmp;
    const uint8_t* bitfield;
    int32_t cascades, grid_size, max_samples; float scale, exp_step_factor, T_threshold, bg, uss_tol;
    /* samples (capacity >= S rows) */
    float *xyzs, *dirs, *unit, *deltas, *ts, *enc, *sigmas, *rgbs, *ws, *d_sigmas, *d_rgbs, *d_enc;
    /* per ray outputs / seeds */
    int32_t* vr_samples; float *opacity, *depth, *rgb, *d_rgb, *d_depth, *d_opacity;
    /* model: flat parameter / gradient / Adam buffers and the slices inside them */
    float *flat_p, *flat_g, *flat_m, *flat_v; int64_t n_params;
    int64_t table_off, w_off[5];          /* offsets (floats) of the hash table and W1..W5 */
    vn_hash_levels_t levels; int32_t hash_flags;
    /* loss + optimiser state */
    float *loss_acc /* [8]: sums[4], counts[4] */, *loss_out /* [1] */;
    float w_color, w_uss, w_tof, w_rgbd;
    float *scale_dev, *found_inf; int32_t* growth_tracker;
    double lr, beta1, beta2, eps; int32_t adam_step;
    /* optional [N, max_samples] scratch: when set, _prepare records the sample positions along each
     * ray (vn_march_train_count_rows) and _run expands them (vn_march_train_expand) instead of
     * marching a second time */
    float* ts_rows;
    /* half-precision encoder inside the step (modules/hash_encoder_half.py:112-213, 364-368): when table_h != NULL
     * the step converts the fp32 table to fp16 (the reference's hash_table.to(float16) per forward), encodes with the
     * half kernel, and the fused MLP backward hands fp16 gradients (what autocast produces) to the half backward
     * kernel (zero-skip rule, fp32 accumulation straight into flat_g).  Needs VN_HASH_F16_CHUNKS | VN_HASH_PLANAR. */
    void* table_h;    /* [total_entries, 2] fp16 scratch */
    /* device-side optimiser step counter (torch Adam's state['step']): int32[1], advanced by the optimiser only when
     * the step is not skipped (found_inf == 0); NULL = use the host counter adam_step */
    float* step_dev;  /* the 4-float optimiser state of vn_opt_state_init */
} vn_step_t;

int vn_train_step_prepare(const vn_step_t* h_step, void* stream);
/* phase: 0 = whole forward+backward, 1 = up to and including loss fwd, 2 = from loss bwd on; + VN_STEP_SKIP_EXPAND when
 * the caller has already enqueued vn_train_step_expand for this step.
 * vn_train_step_expand: the sample-expansion stage alone (march write, or the expansion of the single-pass march's
 * ts_rows, incl. the per-ray SH planes): it depends on _prepare's outputs only -- not on the parameters -- so a caller
 * with double-buffered sample arrays can run it on another stream under the previous step's backward / optimiser. */
#define VN_STEP_SKIP_EXPAND 8
int vn_train_step_expand(const vn_step_t* h_step, int64_t S, void* stream);
int vn_train_step_run(const vn_step_t* h_step, int64_t S, int phase, int do_optim, void* stream);
int vn_train_step_optim(const vn_step_t* h_step, void* stream);

/* ---------------------------------------------------------------------------------------
 * (e) Data-parallel exchange over NVLink peer memory: two-shot sum-allreduce of the flat fp32
 * gradient buffer with plain peer loads (reduce-scatter + all-gather, flag barriers in peer
 * memory with time-outs).  Slice r is reduced by rank r in fixed rank order, so all replicas
 * receive bit-identical sums.
 *  vn_ipc_get_handle: 64-byte cudaIpcMemHandle of the allocation containing a device pointer
 *               and the pointer's byte offset inside it.
 *  vn_ipc_open: map a peer's allocation from its 64-byte cudaIpcMemHandle (+ byte offset).
 *  vn_p2p_init: h_bufs / h_flags = `world` device pointers (own buffer at [rank]); flags are
 *               int32 [world] arrays, zero-initialised; err_dev = local int32 set to 1 on a
 *               barrier time-out.
 *  vn_p2p_allreduce: in place over the first n floats (n % 4 == 0) of every rank's buffer;
 *               must be enqueued by all ranks. */
int vn_ipc_get_handle(const void* ptr, void* h_handle64_out, int64_t* h_offset_out);
int vn_ipc_open(const void* h_handle64, int64_t offset_bytes, void** h_ptr_out);
int vn_p2p_init(int rank, int world, void* const* h_bufs, void* const* h_flags, int* err_dev);
int vn_p2p_allreduce(int64_t n, void* stream);
/* Sharded optimiser fused with the exchange (ZeRO-1 style, replaces allreduce + vn_grad_check +
 * vn_adam_step + vn_scaler_update under data parallelism):
 *  vn_p2p_attach: h_pbufs = `world` device pointers to the ranks' flat PARAMETER buffers (may be
 *               NULL if only the small allreduce is used); h_mbox = `world` pointers to zeroed
 *               fp32 mailboxes of 2 * world * 8 floats.
 *  vn_p2p_allreduce_small: in-place sum (or max) of n <= 8 floats across the ranks through the
 *               mailboxes -- one kernel, replaces the NCCL allreduce of the loss normalisers.
 *  vn_p2p_reduce_adam: barrier -> rank r sums slice r of the gradient over all ranks (fixed order)
 *               and checks it for inf/nan -> barrier + OR of the inf flags -> Adam on slice r with
 *               rank-local m / v (only slice r of m, v is ever touched on rank r) and PUSH of the
 *               updated parameters into every replica -> barrier -> GradScaler update (growth 2,
 *               backoff 0.5, interval 2000).  Arithmetic identical to vn_adam_step. */
int vn_p2p_attach(void* const* h_pbufs, void* const* h_mbox);
int vn_p2p_allreduce_small(float* data, int n, int use_max, void* stream);
int vn_p2p_reduce_adam(int64_t n, float* m, float* v, double lr, double beta1, double beta2, double eps,
                       int step, float* found_inf, float* scale_dev, int32_t* growth_tracker,
                       void* stream);
/*  vn_p2p_step: the same update as ONE kernel with two cross-rank barriers instead of three.  found_inf is an INPUT:
 *               the rank's own inf check (vn_mlp_bwd_scatter's in-kernel check, or vn_grad_check) -- the start barrier
 *               carries it through the mailboxes and every rank applies max over ranks.  Then one pass over slice r:
 *               sum of the ranks' gradients in fixed rank order (peer loads) -> Adam (bias corrections from opt_state,
 *               see vn_opt_state_init) with rank-local m / v -> the updated parameters stored into every replica (peer
 *               stores; inbound and outbound NVLink traffic overlap) -> end barrier -> GradScaler update and, when the
 *               step was applied, opt_state advanced (vn_scaler_update_dev's arithmetic).  Parameters are bit-identical
 *               to vn_p2p_allreduce + vn_adam_step_dev on every rank.  The reduced gradient is not written back.
 *  vn_p2p_shutdown: releases the process-wide exchange context (vn_p2p_init refuses to replace a live one).
 *  Every cross-rank wait gives up after VN_P2P_TIMEOUT_MS (environment, default 20000): *err_dev = 1 and the step is
 *  skipped like an overflow; the host must treat a non-zero err_dev as fatal (TrainEngine raises). */
int vn_p2p_step(int64_t n, float* m, float* v, double lr, double beta1, double beta2, double eps,
                float* opt_state, float* found_inf, float* scale_dev, int32_t* growth_tracker, void* stream);
int vn_p2p_shutdown(void);
/*  vn_p2p_set_multicast: NVLink-multicast (NVLS) addresses of the gradient and the parameter buffers -- one address that
 *               stands for the same offset in every rank's buffer (cuMulticast*; torch's symmetric memory hands them out
 *               as multicast_ptr).  vn_p2p_step then sums slice r with one multimem.ld_reduce per 16 bytes (the NVSwitch
 *               adds the replicas) and distributes the new parameters with one multimem.st (the switch replicates): the
 *               bytes on a GPU's links drop from 2 (n-1)/n to about 2/n of the buffer per direction.  The order of the
 *               additions inside the switch is unspecified: replicas stay bit-identical to each other, the result agrees
 *               with the fixed-order path to rounding.  (NULL, NULL) switches back to peer loads / stores. */
int vn_p2p_set_multicast(void* mc_grad, void* mc_params);

/* ---------------------------------------------------------------------------------------
 * (f) row 2. Batch assembly on the device -- the gather half of DatasetBase.__call__
 * (datasets/dataset_base.py:23-76), _calcRayPoses (:194-243) and get_rays
 * (datasets/ray_utils.py:51-80): rays_o = c2w[:, 3], rays_d = directions[pix] @ c2w[:, :3]^T,
 * rgb / depth gathers.  img_idxs, pix_idxs [B] i32 (the Sampler's dtype, training/sampler.py:111);
 * poses [N_img,3,4]; cam_slot [N_img] = row of directions [n_cams, HW, 3] of the image's camera
 * (the sensor_ids == id masks of :218-231); rgbs [N_img, HW, rgb_stride >= 3]; depth0..3 up to four
 * [N_img, HW] sensor maps (NULL = absent, then out_depthK must be NULL).  An out-of-range index or
 * camera leaves NaN in the ray (:207-208) and sets err[0] = 1 (err may be NULL).
 * ------------------------------------------------------------------------------------- */
int vn_batch_assemble(const int32_t* img_idxs, const int32_t* pix_idxs, int64_t B, const float* poses,
                      const int32_t* cam_slot, int64_t n_imgs, const float* directions, int n_cams,
                      int64_t HW, const float* rgbs, int rgb_stride, const float* depth0,
                      const float* depth1, const float* depth2, const float* depth3,
                      const int32_t* sensor_ids, const float* times, float* rays_o, float* rays_d,
                      float* rgb, float* out_depth0, float* out_depth1, float* out_depth2,
                      float* out_depth3, int32_t* ids_out, float* time_out, int32_t* err, void* stream);

/* (f) row 2, resident ray pool: batch n = pool[sel[draw[n]]] (sel == NULL: pool[draw[n]]) for rays that are already
 * assembled on the device (the real-time mode's seen-so-far rays; the synthetic dataset of the benchmark): rays_o /
 * rays_d / rgb [pool,3] and up to three per-ray depth sensors [pool]; draw [B] i64 uniform integers (torch.randint),
 * sel [n_sel] i64 = the rays that carry a measurement of the sensor being sampled (datasets/dataset_base.py:216-222's
 * valid-depth filter, precomputed).  ONE launch for all outputs; an out-of-range index leaves NaN and sets err[0]. */
int vn_pool_gather(const int64_t* draw, const int64_t* sel, int64_t n_sel, int64_t pool, int64_t B,
                   const float* rays_o, const float* rays_d, const float* rgb, const float* depth0,
                   const float* depth1, const float* depth2, float* out_rays_o, float* out_rays_d,
                   float* out_rgb, float* out_depth0, float* out_depth1, float* out_depth2, int32_t* err,
                   void* stream);

/* ---------------------------------------------------------------------------------------
 * (f) row 3. NGPGrid, modules/ngp_grid.py:37-152 (the Instant-NGP baseline grid of the ablation).
 *  vn_ngp_sample_occupied: indices2 = nonzero(occ > thr)[randint(len, (M,))] (:53-60) without the
 *      host sync of nonzero: indices_out[i] = Morton index of the (rand_idx[i] mod count)-th
 *      occupied cell, -1 when none is occupied.  tmp: >= vn_ngp_select_tmp_ints(n_cells) i32.
 *  vn_ngp_cell_positions: xyzs_w = (coords/(G-1)*2-1)*span + (noise*2-1)*half_cell (:139-143), one
 *      IEEE operation per torch op; span = (float)(s - s/G), half_cell = (float)(s/G).
 *  vn_ngp_grid_update: tmp[indices] = sigmas (duplicates: largest position wins, the CPU
 *      index_put_ order; indices < 0 are skipped), then occ = occ < 0 ? occ : max(occ * decay, tmp)
 *      (:148-152) and tmp is re-zeroed.  winner: [n_cells] i32 filled with -1 (restored on exit);
 *      decay_cells: optional per-cell decay (the `erode` branch, :146-147).
 *  vn_ngp_threshold_pack: thr_out = (mean(occ[occ > 0]), min(mean, density_threshold)) (:155-156,
 *      deterministic double-precision mean) and bitfield = packbits(occ, threshold) (:159-163) over
 *      n_cells_total cells.  scratch: >= vn_ngp_threshold_tmp_bytes() bytes, 8-byte aligned.
 * ------------------------------------------------------------------------------------- */
int64_t vn_ngp_select_tmp_ints(int64_t n_cells);
int vn_ngp_sample_occupied(const float* occ, int64_t n_cells, float threshold, const int64_t* rand_idx,
                           int64_t M, int32_t* tmp, int64_t* indices_out, void* stream);
int vn_ngp_cell_positions(const int32_t* coords, const float* noise, int64_t M, int grid_size,
                          float span, float half_cell, float* xyzs_w, void* stream);
int vn_ngp_grid_update(float* occ, float* tmp, int32_t* winner, int64_t n_cells, const int64_t* indices,
                       const float* sigmas, int64_t M, float decay, const float* decay_cells,
                       void* stream);
int64_t vn_ngp_threshold_tmp_bytes(void);
int vn_ngp_threshold_pack(const float* occ, int64_t n_cells_total, float density_threshold,
                          void* scratch, float* thr_out, uint8_t* bitfield, void* stream);

/* tcgen05 self-test (development / CI): one 128 x N x K fp16 product through the tensor
 * cores in the three operand modes the fused MLP uses (0 forward A*B^T, 1 dgrad A*B,
 * 2 wgrad A^T*B); A, B fp16 row-major as stored, D [128,N] f32. */
int vn_umma_selftest(int mode, int N, int K, const void* A, const void* B, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIRUSNERF_H_ */
