/* voicemap_b200 -- C ABI of the B200-native voicemap speaker-embedding hot path.
 *
 * The reference (oscarknagg/voicemap @ dd79c69) is pure Python on Keras 2.2.2 / TensorFlow 1.10 and has no FFI of
 * its own; its boundary for this path is two Python builder functions (voicemap/models.py:6, :44) plus the Keras
 * model methods its scripts call.  The Python host side (voicemap_b200/models.py) mirrors that interface and
 * binds THIS library with ctypes; every entry point below names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every pointer except `stream` is a DEVICE pointer owned by the caller; nothing here allocates or
 *     synchronises, all work is enqueued on `stream`;
 *   - return value 0 (VM_OK) or a negative VM_ERR_* code; vm_last_error_string() describes the last failure on
 *     the calling thread.  Functions never throw and never fall back to another implementation;
 *   - layouts are Keras': activations (N, L, C) channels-last, conv kernel (K, Cin, Cout), dense (in, out);
 *   - "planes": an fp32 tensor carried as two fp16 tensors of the same shape, x = hi + lo (~22 significant
 *     bits).  This is the inter-block activation format (4 bytes/element, same HBM traffic as fp32) and lets the
 *     fp16 tensor cores produce fp32-grade results with three MMAs per K step (`precision` = 3).  `precision` = 1
 *     uses the hi plane only (throughput mode; lo pointers may be NULL).  `precision` = 2 keeps the fp16 hi plane
 *     and replaces the lo plane by a "Q" plane of the same shape: 16 bits per element holding two e5m2 numbers
 *     {e5m2(x * 2^-6), e5m2((x - hi) * 2^6)}; the two small correction products of the split scheme then run as ONE
 *     fp8 tensor-core product over those byte pairs (8 instead of 12 MMAs per K chunk) -- embeddings stay within
 *     ~4e-5 of the fp64 oracle (tolerance 1e-4), see DESIGN.md section 2.  A forward-pass (eval) mode.
 *   - requires an sm_100a device (B200).
 */
#ifndef VOICEMAP_B200_H
#define VOICEMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VM_OK 0
#define VM_ERR_SHAPE (-1)       /* bad shape / argument */
#define VM_ERR_UNSUPPORTED (-2) /* configuration not implemented (e.g. channel count not a multiple of 8) */
#define VM_ERR_CUDA (-3)        /* CUDA runtime / driver error, see vm_last_error_string() */
#define VM_ERR_ARCH (-4)        /* device is not sm_100 */

#define VM_METRIC_UNIFORM_EUCLIDEAN 0 /* voicemap/models.py:61-69 */
#define VM_METRIC_WEIGHTED_L1 1       /* voicemap/models.py:55-60 */
#define VM_LOSS_NONE 0
#define VM_LOSS_CONTRASTIVE 1 /* voicemap/utils.py:77-85 */
#define VM_LOSS_BCE 2         /* keras 'binary_crossentropy', experiments/train_siamese.py:57 */

int vm_version(void);
const char* vm_last_error_string(void);
/* 0 when the current device can run the kernels (compute capability 10.x), else VM_ERR_ARCH / VM_ERR_CUDA. */
int vm_check_device(void);

/* ---- sizes (host-side arithmetic only) ------------------------------------------------------------------- */
size_t vm_conv1_wpack_bytes(int cout);          /* packed block-1 weights */
size_t vm_conv3_wpack_bytes(int cin, int cout); /* packed block-2..4 weights: fp16 hi, fp16 lo and e5m2x2 Q planes */
size_t vm_epi_bytes(int cout);                  /* per-channel epilogue constants (padded to 128 channels) */
int vm_conv3_num_position_tiles(int L);         /* T of the gmax_partial tensor: 2 * ceil(L / 256) */
int vm_padded_channels(int cout);

/* ---- weight preparation ---------------------------------------------------------------------------------
 * Replaces the (implicit) Keras weight layout of Conv1D + BatchNormalization (voicemap/models.py:13-35):
 * kernel (K, Cin, Cout), bias, BN gamma/beta/moving_mean/moving_variance, eps (Keras default 1e-3) are folded
 * into fp16 (hi, lo) weight planes in the MMA operand layout and four per-channel epilogue constants (opaque to the
 * caller).  With s = gamma / sqrt(var + eps), t = beta - mean * s, sigma = sign(s): weights are stored sigma-scaled
 * so that max-pooling commutes with the BN affine when gamma < 0, and bias + ReLU + BN become one clamp and one
 * FMA on the pooled accumulator maximum. */
int vm_pack_conv1(const float* kernel /* (32, 1, Cout) */, const float* bias, const float* gamma,
                  const float* beta, const float* mean, const float* var, float eps, int cout, void* wpack,
                  float* epi, void* stream);
int vm_pack_conv3(const float* kernel /* (3, Cin, Cout) */, const float* bias, const float* gamma,
                  const float* beta, const float* mean, const float* var, float eps, int cin, int cout, void* wpack,
                  float* epi, void* stream);

/* ---- block 1: Conv1D(filters, 32, 'same', relu) -> BatchNormalization -> MaxPool1D(pool, pool) ------------
 * voicemap/models.py:13-19.  x (N, L) fp32 (Keras (N, L, 1)); out planes (N, L/pool, Cout).  pool = 4 is the
 * reference architecture (voicemap/models.py:19); pool = 2 is the older architecture of the checkpoint shipped
 * under models/n_seconds/ (its model_config has four MaxPooling1D(2)). */
int vm_conv1_relu_bn_pool_fwd(const float* x, int N, int L, int cout, int pool, const void* wpack, const float* epi,
                              uint16_t* out_hi, uint16_t* out_lo, int precision, void* stream);

/* ---- blocks 2-4: Conv1D(C, 3, 'same', relu) -> BatchNormalization -> MaxPool1D(2) -------------------------
 * voicemap/models.py:22-35.  in planes (N, L, Cin); out planes (N, L/2, Cout).
 * If gmax_partial != NULL the block is merged with GlobalMaxPool1D (voicemap/models.py:37): nothing is written
 * to out_*; instead gmax_partial (N, T, Cpad) receives per-position-tile maxima of the raw accumulators, to be
 * finished by vm_gmax_dense_fwd.  T = vm_conv3_num_position_tiles(L), Cpad = vm_padded_channels(Cout). */
int vm_conv3_relu_bn_pool2_fwd(const uint16_t* in_hi, const uint16_t* in_lo, int N, int L, int cin, int cout,
                               const void* wpack, const float* epi, uint16_t* out_hi, uint16_t* out_lo,
                               float* gmax_partial, int precision, void* stream);

/* ---- GlobalMaxPool1D finalisation + Dense(embedding_dimension) -------------------------------------------
 * voicemap/models.py:37-39.  gmax_out (N, C) optional; emb (N, E) = gmax . dense_w (C, E) + dense_b. */
int vm_gmax_dense_fwd(const float* gmax_partial, int N, int T, int C, const float* epi, const float* dense_w,
                      const float* dense_b, int E, float* gmax_out, float* emb, void* stream);

/* ---- siamese head + loss ----------------------------------------------------------------------------------
 * voicemap/models.py:55-69 (+ voicemap/utils.py:77-85 / keras binary_crossentropy).  e1, e2 (N, E);
 * metric VM_METRIC_*: uniform_euclidean: d = sqrt(sum (e1-e2)^2), p = sigmoid(head_w[0]*d + head_b[0]);
 * weighted_l1: p = sigmoid(sum_j head_w[j]*|e1-e2|_j + head_b[0]).  y_true (N) with 0 = same speaker
 * (voicemap/librispeech.py:194).  dist (N) / prob (N) / loss (1) may each be NULL. */
int vm_pair_head_loss_fwd(const float* e1, const float* e2, int N, int E, int metric, const float* head_w,
                          const float* head_b, const float* y_true, int loss_kind, float* dist, float* prob,
                          float* loss, void* stream);

/* ---- plane conversion (per-block fp32 views for callers and tests) ---------------------------------------- */
int vm_split_planes(const float* x, size_t n, uint16_t* hi, uint16_t* lo, void* stream);
int vm_merge_planes(const uint16_t* hi, const uint16_t* lo, size_t n, float* x, void* stream);
/* the same for precision 2: (fp16 hi, e5m2x2 Q) planes; merging decodes hi + residual byte */
int vm_split_planes_q(const float* x, size_t n, uint16_t* hi, uint16_t* q, void* stream);
int vm_merge_planes_q(const uint16_t* hi, const uint16_t* q, size_t n, float* x, void* stream);

/* ---- whole encoder (voicemap/models.py:6-41, eval mode) ---------------------------------------------------
 * x (N, L) fp32 -> emb (N, E).  `wpack[i]`, `epi[i]` (i = 0..3) from vm_pack_conv1 / vm_pack_conv3 for channel
 * widths filters*{1,2,3,4}.  first_pool = size of the first MaxPool1D (4, or 2 for the older architecture, see
 * vm_conv1_relu_bn_pool_fwd).  `workspace` must hold vm_encoder_workspace_bytes(N, L, filters, first_pool) bytes. */
size_t vm_encoder_workspace_bytes(int N, int L, int filters, int first_pool);
int vm_encoder_fwd(const float* x, int N, int L, int filters, int first_pool, const void* const* wpack,
                   const float* const* epi, const float* dense_w, const float* dense_b, int E, void* workspace,
                   float* emb, int precision, void* stream);

/* ---- fused preprocessing (voicemap/utils.py:22-34, 88-101) ------------------------------------------------
 * The reference decimates raw 16 kHz clips on the host (instances[:, ::downsampling, :]) and whitens them: per-clip
 * mean removal, then ONE scale rms / sqrt(mean(batch^2)) per whiten() call, taken over the un-centred decimated
 * batch.  vm_preprocess_stats computes mean (N) and scale (N; equal inside each of the G groups = whiten() calls)
 * in double precision; `scale` must hold N floats followed by 4*N + 4 floats of scratch.
 * vm_encoder_fwd_raw runs the encoder on raw audio x (N, T) fp32 with the decimation (strided read) and the
 * whitening affine fused into block 1's operand producer: L = ceil(T / downsampling) samples enter the network.
 * whiten_groups = 0 disables whitening.  `workspace` must hold vm_encoder_workspace_bytes(N, L, filters, first_pool)
 * + vm_preprocess_scratch_bytes(N) bytes. */
size_t vm_preprocess_scratch_bytes(int N);
int vm_preprocess_stats(const float* x, int N, int T, int downsampling, int G, float rms, float* mean, float* scale,
                        void* stream);
int vm_encoder_fwd_raw(const float* x, int N, int T, int downsampling, int whiten_groups, float rms, int filters,
                       int first_pool, const void* const* wpack, const float* const* epi, const float* dense_w,
                       const float* dense_b, int E, void* workspace, float* emb, int precision, void* stream);

/* =============================================================================================================
 * Training (SURVEY.md 8(a) a3 train mode, a4, a13): what Keras' fit_generator does implicitly for
 * experiments/train_siamese.py:56-65 / train_classifier.py:114-120.  "BN group" = one application of the shared
 * encoder (the siamese net applies it once per branch: voicemap/models.py:52-53); clips [g*N/G, (g+1)*N/G) are
 * group g.  Gradients carry the loss scale folded into vm_pair_head_loss_bwd; vm_adam_step divides it out.
 * ============================================================================================================= */

/* "raw" packing for train mode: identity BN, sigma = +1, epilogue y = relu(acc + bias).  bias may be NULL. */
int vm_pack_conv1_raw(const float* kernel, const float* bias, int cout, void* wpack, float* epi, void* stream);
int vm_pack_conv3_raw(const float* kernel, const float* bias, int cin, int cout, void* wpack, float* epi,
                      void* stream);
/* dgrad operand: tap-flipped, channel-transposed kernel (bf16 planes) in the conv3 layout with (cin' = Cout, cout' = Cin);
 * wpack holds vm_conv3_wpack_bytes(cout, cin) bytes, epi vm_epi_bytes(cin). */
int vm_pack_conv3_dgrad(const float* kernel /* (3, Cin, Cout) */, int cin, int cout, void* wpack, float* epi,
                        void* stream);

/* Train-mode conv forward: u = relu(conv(x) + bias), un-pooled fp32 (N, L, Cout), plus per-channel {sum, sumsq}
 * partial rows stat_partial (N * 2*ceil(L/256), Cpad) float2 (may be NULL). */
int vm_conv1_raw_fwd(const float* x, int N, int L, int cout, const void* wpack, const float* epi, float* u,
                     float* stat_partial, int precision, void* stream);
/* linear = 0: as above for blocks 2-4.  linear = 1: plain convolution output (dgrad: in = dU bf16 planes, wpack from
 * vm_pack_conv3_dgrad, out = dX fp32 (N, L, Cin)). */
int vm_conv3_raw_fwd(const uint16_t* in_hi, const uint16_t* in_lo, int N, int L, int cin, int cout,
                     const void* wpack, const float* epi, float* out, float* stat_partial, int linear, int precision,
                     void* stream);
int vm_stat_rows_per_clip(int L); /* 2 * ceil(L / 256) */
/* bytes of the `red_scratch` buffer the two-stage (deterministic, atomics-free) channel reductions need */
size_t vm_reduce_scratch_bytes(int G, int C);

/* Batch statistics -> bn_const (G, C) x {s, t, mean, rstd}; Keras moving-average update (momentum .99, sample
 * variance n/(n-(1+eps))) applied once per group, in order.  moving_* may be NULL. */
int vm_bn_stats_finalize(const float* stat_partial, int rows_per_clip, int N, int G, int L, int C,
                         const float* gamma, const float* beta, float eps, float momentum, float* moving_mean,
                         float* moving_var, float* bn_const, double* red_scratch, void* stream);
/* y = bn(u) * mask -> MaxPool1D(pool) -> fp16 planes (N, L/pool, C) for the next block's forward conv, and
 * (optional, both or neither) the same values as bf16 planes for vm_wgrad3 (the tensor core cannot mix fp16 with
 * bf16 operands).  mask (N, C) = SpatialDropout1D keep/(1-p), or NULL. */
int vm_bn_pool_fwd(const float* u, int N, int L, int C, int G, int pool, const float* bn_const, const float* mask,
                   uint16_t* out_hi, uint16_t* out_lo, uint16_t* bf_hi, uint16_t* bf_lo, void* stream);
/* block 4: bn -> MaxPool1D(2) -> GlobalMaxPool1D merged; gmax (N, C), argmax (N, C) un-pooled position. */
int vm_bn_gmax_fwd(const float* u, int N, int L, int C, int G, const float* bn_const, const float* mask, float* gmax,
                   int32_t* argmax, void* stream);
int vm_dense_fwd(const float* x, int N, int C, const float* w, const float* b, int E, float* y, void* stream);

/* Backward of the siamese head + loss (emb (2N, E): branch 1 rows then branch 2 rows). */
int vm_pair_head_loss_bwd(const float* emb, int N, int E, int metric, const float* head_w, const float* head_b,
                          const float* y_true, int loss_kind, float loss_scale, float* d_emb, float* d_head_w,
                          float* d_head_b, void* stream);
int vm_dense_bwd(const float* x, const float* dy, const float* w, int N, int C, int E, float* dw, float* db, float* dx,
                 void* stream);
/* BN + MaxPool + ReLU backward of one block.  Give dy_pooled (N, L/pool, C) (blocks 1-3) XOR d_gmax + argmax
 * (block 4).  Outputs: dgamma, dbeta, dbias (C), dU as bf16 (hi, lo) planes (N, L, C) -- gradients are carried in
 * bf16 pairs (fp32 exponent range, ~16 significant bits), so no loss scaling is required.  scratch_f2 / scratch_f hold
 * N * chunks * max(1, 512/C) rows of C float2 / float; bwd_const (G, C) float4. */
int vm_bn_bwd(const float* u, const float* dy_pooled, const float* d_gmax, const int32_t* argmax, int N, int L, int C,
              int G, int pool, const float* bn_const, const float* mask, float* scratch_f2, int chunks,
              float* bwd_const, float* dgamma, float* dbeta, uint16_t* du_hi, uint16_t* du_lo, float* scratch_f,
              float* dbias, double* red_scratch, void* stream);
/* Synchronised BatchNorm across data-parallel ranks (the reference's BN sees the whole batch on one
 * device).  The two calls above are split at the point where the per-(group, channel) sums exist, so that the caller
 * can all-reduce them (torch.distributed / NCCL) in between:
 *   forward : vm_bn_stats_sums -> all-reduce(sums) -> vm_bn_stats_from_sums(count = global clips per group * L)
 *   backward: vm_bn_bwd_sums   -> all-reduce(copy of sums) -> vm_bn_bwd_from_sums(local sums, global sums, count)
 * sums: (G, C) x {sum, sum of squares} resp. {sum dy, sum dy*xhat} as doubles.  dgamma / dbeta are formed from the
 * LOCAL sums (the gradient all-reduce adds the ranks), the batch means from the GLOBAL sums.  With one rank
 * (global = local) the results equal vm_bn_stats_finalize / vm_bn_bwd bit for bit. */
int vm_bn_stats_sums(const float* stat_partial, int rows_per_clip, int N, int G, int C, double* red_scratch,
                     double* sums, void* stream);
int vm_bn_stats_from_sums(const double* sums, double count, int G, int C, const float* gamma, const float* beta,
                          float eps, float momentum, float* moving_mean, float* moving_var, float* bn_const,
                          void* stream);
int vm_bn_bwd_sums(const float* u, const float* dy_pooled, const float* d_gmax, const int32_t* argmax, int N, int L,
                   int C, int G, int pool, const float* bn_const, const float* mask, float* scratch_f2, int chunks,
                   double* red_scratch, double* sums, void* stream);
int vm_bn_bwd_from_sums(const double* local_sums, const double* global_sums, double count, const float* u,
                        const float* dy_pooled, const float* d_gmax, const int32_t* argmax, int N, int L, int C, int G,
                        int pool, const float* bn_const, const float* mask, int chunks, float* bwd_const,
                        float* dgamma, float* dbeta, uint16_t* du_hi, uint16_t* du_lo, float* scratch_f, float* dbias,
                        double* red_scratch, void* stream);
/* dW (3, Cin, Cout) = sum_{n,p} X[n][p+tap-1][ci] * dU[n][p][co] on tensor cores; x_* and du_* are bf16 planes;
 * partial: scratch. */
int vm_wgrad3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* du_hi, const uint16_t* du_lo, int N, int L,
              int cin, int cout, int precision, float* partial, size_t partial_bytes, float* dw, void* stream);
/* dW1 (32, 1, Cout) = sum_{n,p} x[n][p+k-15] * dU1[n][p][co].  precision 3 / 1: tensor cores (bf16 Toeplitz operand
 * built in shared memory, 3 or 1 MMAs per K step); precision 0: fp32 CUDA-core reference kernel. */
int vm_wgrad1(const float* x, const uint16_t* du_hi, const uint16_t* du_lo, int N, int L, int cout, int precision,
              float* partial, size_t partial_bytes, float* dw, void* stream);
/* keras.optimizers.Adam update with global-norm clipping on a flat parameter buffer:
 * g' = g * inv_scale * min(1, clipnorm / ||g * inv_scale||) (clipnorm <= 0: off); m, v, p updated in place with
 * p -= lr_t * m / (sqrt(v) + eps), lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) computed by the caller. */
int vm_adam_step(float* p, const float* g, float* m, float* v, size_t n, double* scratch, float inv_scale,
                 float clipnorm, float lr_t, float beta1, float beta2, float eps, void* stream);

/* ---- testing / tuning knobs -------------------------------------------------------------------------------
 * key "max_ctas" (0 = all SMs; limits the persistent grid).  Returns the previous value or VM_ERR_SHAPE. */
int vm_set_option(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* VOICEMAP_B200_H */
