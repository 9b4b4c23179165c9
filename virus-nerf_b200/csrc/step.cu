// step.cu -- native step runner: enqueues every kernel of one train step from C++ (see
// include/virusnerf.h "Native step runner").  Pure orchestration over the other entry points.
#include "common.cuh"

#define VN_TRY(call) do { int rc__ = (call); if (rc__ != VN_OK) return rc__; } while (0)

VN_API int vn_train_step_prepare(const vn_step_t* s, void* stream) {
    VN_REQUIRE(s != nullptr, "vn_train_step_prepare: null step");
    VN_TRY(vn_ray_aabb(s->rays_o, s->rays_d, s->scale, s->N, s->hits_t, stream));
    if (s->ts_rows)     // single-pass march: the count pass records the sample positions along each ray
        VN_TRY(vn_march_train_count_rows(s->rays_o, s->rays_d, s->hits_t, s->bitfield, s->noise, s->N, s->cascades,
                                         s->grid_size, s->scale, s->exp_step_factor, s->max_samples, s->counts, s->rays_a,
                                         s->counter, s->scan_tmp, s->ts_rows, stream));
    else
        VN_TRY(vn_march_train_count(s->rays_o, s->rays_d, s->hits_t, s->bitfield, s->noise, s->N, s->cascades, s->grid_size,
                                    s->scale, s->exp_step_factor, s->max_samples, s->counts, s->rays_a, s->counter,
                                    s->scan_tmp, stream));
    return VN_OK;
}

VN_API int vn_train_step_optim(const vn_step_t* s, void* stream) {
    VN_REQUIRE(s != nullptr, "vn_train_step_optim: null step");
    // VN_HASH_FUSED_SCATTER: the fused backward kernel has already evaluated the inf check on the contributions
    if (!(s->hash_flags & VN_HASH_FUSED_SCATTER)) VN_TRY(vn_grad_check(s->flat_g, s->n_params, s->found_inf, stream));
    if (s->step_dev) {       // step count on the device: skipped steps do not advance Adam's bias corrections
        VN_TRY(vn_adam_step_dev(s->flat_p, s->flat_g, s->flat_m, s->flat_v, s->n_params, s->lr, s->beta1, s->beta2, s->eps,
                                (const float*)s->step_dev, s->found_inf, s->scale_dev, stream));
        VN_TRY(vn_scaler_update_dev(s->scale_dev, s->growth_tracker, s->found_inf, 2.0f, 0.5f, 2000, (float*)s->step_dev, s->lr,
                                    s->beta1, s->beta2, stream));
        return VN_OK;
    }
    VN_TRY(vn_adam_step(s->flat_p, s->flat_g, s->flat_m, s->flat_v, s->n_params, 1.0f, s->lr, s->beta1, s->beta2, s->eps,
                        s->adam_step, s->found_inf, s->scale_dev, stream));
    VN_TRY(vn_scaler_update(s->scale_dev, s->growth_tracker, s->found_inf, 2.0f, 0.5f, 2000, stream));
    return VN_OK;
}

// the sample-expansion stage of a step (march write / expand): depends only on the front half, not on the parameters, so
// the caller may enqueue it early on another stream (into buffers the previous step does not use) and pass
// VN_STEP_SKIP_EXPAND to vn_train_step_run
VN_API int vn_train_step_expand(const vn_step_t* s, int64_t S, void* stream) {
    VN_REQUIRE(s != nullptr && S >= 0, "vn_train_step_expand: bad arguments");
    const bool chunks = (s->hash_flags & VN_HASH_F16_CHUNKS) != 0;
    if (s->ts_rows && chunks)      // enc = [4 hash planes | 2 SH planes][S] x 16 B
        VN_TRY(vn_march_train_expand_sh(s->rays_o, s->rays_d, s->rays_a, s->ts_rows, s->N, s->max_samples, s->grid_size,
                                        s->scale, s->exp_step_factor, S, s->xyzs, s->dirs, s->deltas, s->ts, s->unit,
                                        (char*)s->enc + (size_t)S * 64, S, stream));
    else if (s->ts_rows)
        VN_TRY(vn_march_train_expand(s->rays_o, s->rays_d, s->rays_a, s->ts_rows, s->N, s->max_samples, s->grid_size,
                                     s->scale, s->exp_step_factor, S, s->xyzs, s->dirs, s->deltas, s->ts, s->unit, stream));
    else
        VN_TRY(vn_march_train_write(s->rays_o, s->rays_d, s->hits_t, s->bitfield, s->noise, s->N, s->cascades,
                                    s->grid_size, s->scale, s->exp_step_factor, s->rays_a, S, s->xyzs, s->dirs, s->deltas,
                                    s->ts, s->unit, stream));
    return VN_OK;
}

VN_API int vn_train_step_run(const vn_step_t* s, int64_t S, int phase, int do_optim, void* stream) {
    const bool skip_expand = (phase & VN_STEP_SKIP_EXPAND) != 0;
    phase &= ~VN_STEP_SKIP_EXPAND;
    VN_REQUIRE(s != nullptr && S >= 0 && phase >= 0 && phase <= 2, "vn_train_step_run: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const float* W[5];
    float* dW[5];
    for (int i = 0; i < 5; ++i) { W[i] = s->flat_p + s->w_off[i]; dW[i] = s->flat_g + s->w_off[i]; }
    const float* table = s->flat_p + s->table_off;
    float* table_grad = s->flat_g + s->table_off;
    float* sums = s->loss_acc;
    float* cnts = s->loss_acc + 4;
    // enc / d_enc stay inside the step: use the level-pair-plane layout when the flags ask for it
    // VN_HASH_F16_CHUNKS: the hash forward emits fp16 operand chunks that the MLP kernels bulk-copy (enc_format 3,
    // d_enc stays f32 planes); else f32 planes (2) or rows (0)
    // (5 when the single-pass march also emits the per-ray direction encoding as two more operand planes)
    const bool chunks = (s->hash_flags & VN_HASH_F16_CHUNKS) != 0;
    const int enc_fmt = chunks ? (s->ts_rows ? 5 : 3) : ((s->hash_flags & VN_HASH_PLANAR) ? 2 : 0);
    const bool half_enc = s->table_h != nullptr;       // hash_encoder_half.py inside the step
    VN_REQUIRE(!half_enc || chunks, "vn_train_step_run: the half-precision encoder needs VN_HASH_F16_CHUNKS");
    const bool fused_scatter = (s->hash_flags & VN_HASH_FUSED_SCATTER) != 0;
    VN_REQUIRE(!fused_scatter || (chunks && s->levels.levels == 16),
               "vn_train_step_run: VN_HASH_FUSED_SCATTER needs VN_HASH_F16_CHUNKS and 16 levels");
    VN_REQUIRE(!(s->hash_flags & VN_HASH_F16_CHUNKS) || (s->hash_flags & VN_HASH_PLANAR),
               "vn_train_step_run: VN_HASH_F16_CHUNKS needs VN_HASH_PLANAR (d_enc planes)");
    if (phase == 0 || phase == 1) {
        VN_CUDA(cudaMemsetAsync(s->flat_g, 0, sizeof(float) * (size_t)s->n_params, st));
        VN_CUDA(cudaMemsetAsync(s->loss_acc, 0, sizeof(float) * 8, st));
        if (!skip_expand) VN_TRY(vn_train_step_expand(s, S, stream));
        if (half_enc) {
            VN_TRY(vn_f32_to_f16(table, s->table_h, 2 * s->levels.total_entries, stream));       // hash_encoder_half.py:367
            VN_TRY(vn_hash_encode_fwd_f16(s->unit, s->table_h, s->enc, S, &s->levels, s->hash_flags, stream));
        } else {
            VN_TRY(vn_hash_encode_fwd_f32(s->unit, table, s->enc, S, &s->levels, s->hash_flags, stream));
        }
        VN_TRY(vn_mlp_fwd(s->enc, enc_fmt, s->dirs, W[0], W[1], W[2], W[3], W[4], S, 0, s->sigmas, s->rgbs, nullptr, stream));
        // compositing forward + the loss terms of its rays in one kernel (vn_composite_train_fwd + vn_loss_fwd)
        VN_TRY(vn_composite_loss_fwd(s->sigmas, s->rgbs, s->deltas, s->ts, s->rays_a, s->N, S, s->T_threshold, s->vr_samples,
                                     s->opacity, s->depth, s->rgb, s->ws, s->gt_rgb, s->uss, s->tof, s->rgbd, s->bg, s->uss_tol,
                                     sums, cnts, stream));
    }
    if (phase == 0 || phase == 2) {
        // loss backward (gradient seeds from the global counts) + compositing backward in one kernel
        VN_TRY(vn_composite_loss_bwd(s->sigmas, s->rgbs, s->deltas, s->ts, s->rays_a, s->N, S, s->T_threshold, s->rgb, s->opacity,
                                     s->depth, s->gt_rgb, s->uss, s->tof, s->rgbd, s->bg, s->uss_tol, sums, cnts, s->w_color,
                                     s->w_uss, s->w_tof, s->w_rgbd, s->scale_dev, s->d_sigmas, s->d_rgbs, s->loss_out, stream));
        if (fused_scatter) {
            // one kernel: d(enc) goes from tensor memory straight into the table gradient
            VN_TRY(vn_mlp_bwd_scatter(s->enc, enc_fmt, s->dirs, W[0], W[1], W[2], W[3], W[4], S, s->d_sigmas, s->d_rgbs, s->unit,
                                      &s->levels, half_enc ? 1 : 0, table_grad, dW[0], dW[1], dW[2], dW[3], dW[4], s->found_inf,
                                      stream));
        } else {
            VN_TRY(vn_mlp_bwd(s->enc, half_enc ? (enc_fmt | VN_MLP_DENC_F16) : enc_fmt, s->dirs, W[0], W[1], W[2], W[3], W[4], S, 0,
                              s->d_sigmas, s->d_rgbs, s->d_enc, dW[0], dW[1], dW[2], dW[3], dW[4], stream));
            if (half_enc) VN_TRY(vn_hash_encode_bwd_f16(s->unit, s->d_enc, table_grad, S, &s->levels, s->hash_flags, stream));
            else          VN_TRY(vn_hash_encode_bwd_f32(s->unit, s->d_enc, table_grad, S, &s->levels, s->hash_flags, stream));
        }
        if (do_optim) VN_TRY(vn_train_step_optim(s, stream));
    }
    return VN_OK;
}
