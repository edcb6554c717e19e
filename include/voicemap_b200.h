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

/* ---- k-way n-shot scoring on embeddings ---------------------------------------------------------------------
 * voicemap/utils.py:156-212: query (T, E); support (T, k*n, E) ordered [class 0]*n + ... + [class k-1]*n
 * (voicemap/librispeech.py:204-240).  Per task: mean embedding of each class, `distance` to the query
 * (VM_DISTANCE_EUCLIDEAN: |mean - q|; VM_DISTANCE_COSINE: 1 - cos(mean of unit vectors, q), scipy cdist 'cosine';
 * VM_DISTANCE_DOT: -(mean unit vector * mean norm) . q), then best[t] = arg-min class (first minimum): the task is
 * solved when best[t] == 0.  scores (T, k) optional.  Sums in double, like the reference's numpy / scipy. */
#define VM_DISTANCE_EUCLIDEAN 0
#define VM_DISTANCE_COSINE 1
#define VM_DISTANCE_DOT 2
int vm_nshot_score(const float* query, const float* support, int T, int k, int n, int E, int distance, float* scores,
                   int32_t* best, void* stream);

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
/* dgrad operand: tap-flipped, channel-transposed kernel (fp16 hi/lo planes) in the conv3 layout with
 * (cin' = Cout, cout' = Cin); wpack holds vm_conv3_wpack_bytes(cout, cin) bytes, epi vm_epi_bytes(cin). */
int vm_pack_conv3_dgrad(const float* kernel /* (3, Cin, Cout) */, int cin, int cout, void* wpack, float* epi,
                        void* stream);

/* The three calls above for all four blocks in ONE launch (the weights change once per training step): kernels[4],
 * biases[4] in Keras layout for blocks 1-4 with Cout = filters * (1, 2, 3, 4); wraw[4] / eraw[4] receive the "raw"
 * forward operands, wdg[4] / edg[4] (entries 1-3 used; entry 0 ignored) the dgrad operands. */
int vm_pack_train(const float* const* kernels, const float* const* biases, int filters, void* const* wraw,
                  float* const* eraw, void* const* wdg, float* const* edg, void* stream);

/* Train-mode conv forward of one block: u = relu(conv(x) + bias) at every position, of which the call keeps
 *   u16 (N, L, Cout)       16 bits per element: fp16(u) in bits 0-14, bit 15 = "arg-max of its MaxPool window" (first
 *                          winner on ties; the arg-MIN of u where gamma[c] < 0, i.e. the arg-max after BatchNorm);
 *   ext (N, L/pool, Cout)  fp32 extreme of u per window (max, or min where gamma[c] < 0): the forward pass continues
 *                          from these exact values, BatchNorm being monotone per channel;
 *   stat_partial           per-channel {sum, sumsq} rows (N * rows_per_clip, Cpad) float2 for the batch statistics
 *                          (rows_per_clip: vm_stat_rows_per_clip for block 1, vm_conv3_train_rows_per_clip for 2-4).
 * gamma (Cout) supplies the signs only (NULL = all maxima).  pool: 4 or 2 for block 1, always 2 for blocks 2-4.
 * precision 2 (blocks 2-4): in_lo is the e5m2x2 Q plane; 3: the fp16 residual plane. */
int vm_conv1_train_fwd(const float* x, int N, int L, int cout, int pool, const void* wpack, const float* epi,
                       const float* gamma, uint16_t* u16, float* ext, float* stat_partial, int precision, void* stream);
int vm_conv3_train_fwd(const uint16_t* in_hi, const uint16_t* in_lo, int N, int L, int cin, int cout,
                       const void* wpack, const float* epi, const float* gamma, uint16_t* u16, float* ext,
                       float* stat_partial, int precision, void* stream);
/* Data gradient of a block's convolution: dX (N, L, Cin) fp32 = conv3(dU, flipped/transposed kernel).  du_hi / du_lo
 * (N, L, Cout): the scaled fp16 gradient planes vm_bn_bwd wrote, grad_absmax the word it derived their power-of-two
 * scale from (the epilogue takes the scale out again).  precision 3: dUh*Wh + dUl*Wh + dUh*Wl; 2: one-plane gradient
 * (du_lo NULL), dUh*Wh + dUh*Wl; 1: dUh*Wh.
 * below_* (optional, below_partial NULL = off): the BatchNorm-backward reduction of the block BELOW, whose pooled
 * gradient this call produces, taken in the epilogue while the values are in registers: below_ext (N, L, Cin) its
 * window extremes, below_bn_const (G, Cin) x 4, below_mask (N, Cin) or NULL -> below_partial,
 * vm_conv3_train_rows_per_clip(L) rows of Cpad float2 {sum dy, sum dy * xhat} per clip, and *below_absmax (largest
 * |s * dy|).  The following vm_bn_bwd of that block is then called with presummed_rows_per_clip = that row count, scratch_f2 =
 * below_partial, grad_absmax = below_absmax, and skips its own pass over the pooled tensors. */
int vm_conv3_dgrad(const uint16_t* du_hi, const uint16_t* du_lo, int N, int L, int cout, int cin,
                   const void* wpack_dgrad, const float* epi_dgrad, const uint32_t* grad_absmax, float* dx,
                   int precision, const float* below_ext, const float* below_bn_const, const float* below_mask,
                   int below_groups, float* below_partial, uint32_t* below_absmax, void* stream);
int vm_stat_rows_per_clip(int L);        /* 2 * ceil(L / 256): partial rows per clip written by vm_conv1_train_fwd */
int vm_conv3_train_rows_per_clip(int L); /* 4 * ceil(L / 256): ... by vm_conv3_train_fwd (stat_partial) and by the fused
                                          * reduction of vm_conv3_dgrad (below_partial) */
/* bytes of the `red_scratch` buffer the two-stage (deterministic, atomics-free) channel reductions need */
size_t vm_reduce_scratch_bytes(int G, int C);

/* Batch statistics -> bn_const (G, C) x {s, t, mean, rstd}; Keras moving-average update (momentum .99, sample
 * variance n/(n-(1+eps))) applied once per group, in order.  moving_* may be NULL. */
int vm_bn_stats_finalize(const float* stat_partial, int rows_per_clip, int N, int G, int L, int C,
                         const float* gamma, const float* beta, float eps, float momentum, float* moving_mean,
                         float* moving_var, float* bn_const, double* red_scratch, void* stream);
/* y = (s * ext + t) * mask on the window extremes = MaxPool1D(pool)(SpatialDropout(BatchNorm(u))) -> planes
 * (N, Lout, C) of the next block: out_hi fp16 always; out_lo the fp16 residual plane (forward precision 3, and the
 * second activation plane of vm_wgrad3) and / or out_q the e5m2x2 Q plane (forward precision 2), each optional
 * (NULL).  mask (N, C) = SpatialDropout1D keep/(1-p), or NULL. */
int vm_bn_pool_fwd(const float* ext, int N, int Lout, int C, int G, const float* bn_const, const float* mask,
                   uint16_t* out_hi, uint16_t* out_lo, uint16_t* out_q, void* stream);
/* block 4: bn on the MaxPool1D(2) window extremes -> GlobalMaxPool1D; gmax (N, C), jstar (N, C) the winning window. */
int vm_bn_gmax_fwd(const float* ext, int N, int Lout, int C, int G, const float* bn_const, const float* mask,
                   float* gmax, int32_t* jstar, void* stream);
int vm_dense_fwd(const float* x, int N, int C, const float* w, const float* b, int E, float* y, void* stream);

/* Backward of the siamese head + loss (emb (2N, E): branch 1 rows then branch 2 rows).  accuracy (1 float, optional):
 * keras' 'accuracy' metric of the batch, mean(round(p) == y). */
int vm_pair_head_loss_bwd(const float* emb, int N, int E, int metric, const float* head_w, const float* head_b,
                          const float* y_true, int loss_kind, float loss_scale, float* d_emb, float* d_head_w,
                          float* d_head_b, float* accuracy, void* stream);
int vm_dense_bwd(const float* x, const float* dy, const float* w, int N, int C, int E, float* dw, float* db, float* dx,
                 void* stream);
/* The siamese training head in one call (two launches): for pair n of the 2N-clip batch (gmax (2N, C): branch 1 rows,
 * then branch 2 rows) Dense -> emb (2N, E), the distance layer and Dense(1, sigmoid) of voicemap/models.py:52-69 ->
 * prob (N, optional), the loss (1 contrastive, voicemap/utils.py:77-85; 2 binary cross-entropy) -> loss_acc {mean loss,
 * keras accuracy}, and the backward pass down to d_gmax (2N, C): d_emb (2N, E), the Dense gradients d_dense_w (C, E) /
 * d_dense_b (E) and the head gradients d_head_w (1 | E) / d_head_b (1), all multiplied by loss_scale.  Equivalent to
 * vm_dense_fwd + vm_pair_head_loss_fwd + vm_pair_head_loss_bwd + vm_dense_bwd.  pair_scratch: 4 * N floats. */
int vm_siamese_head_train(const float* gmax, int N, int C, int E, const float* dense_w, const float* dense_b, int metric,
                          const float* head_w, const float* head_b, const float* y_true, int loss_kind, float loss_scale,
                          float* emb, float* prob, float* d_emb, float* d_gmax, float* pair_scratch, float* d_dense_w,
                          float* d_dense_b, float* d_head_w, float* d_head_b, float* loss_acc, void* stream);
/* BN + MaxPool + ReLU backward of one block.  Give dy_pooled (N, L/pool, C) (blocks 1-3) XOR d_gmax (N, C) + jstar
 * (block 4: the gradient sits in window jstar[n][c]).  Outputs: dgamma, dbeta, dbias (C) and dU (N, L, C) as fp16
 * planes scaled by a power of two that is derived from the largest |s * dy| of the block (stored as float bits in
 * *grad_absmax; vm_wgrad* / vm_conv3_dgrad take the scale out again): du_hi always, du_lo = fp16 residual when not
 * NULL.  scratch_f2 / scratch_f: vm_bn_bwd_scratch_elems(N) float2 / float entries; bwd_const (G, C) float4. */
size_t vm_bn_bwd_scratch_elems(int N);
int vm_bn_bwd(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax, const int32_t* jstar,
              int N, int L, int C, int G, int pool, const float* bn_const, const float* mask, float* scratch_f2,
              float* bwd_const, float* dgamma, float* dbeta, uint32_t* grad_absmax, uint16_t* du_hi, uint16_t* du_lo,
              float* scratch_f, float* dbias, double* red_scratch, int presummed_rows_per_clip, void* stream);
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
int vm_bn_bwd_sums(const float* ext, const float* dy_pooled, const float* d_gmax, const int32_t* jstar, int N, int L,
                   int C, int G, int pool, const float* bn_const, const float* mask, float* scratch_f2,
                   uint32_t* grad_absmax, double* red_scratch, double* sums, int presummed_rows_per_clip, void* stream);
int vm_bn_bwd_from_sums(const double* local_sums, const double* global_sums, double count, const uint16_t* u16,
                        const float* dy_pooled, const float* d_gmax, const int32_t* jstar, int N, int L, int C, int G,
                        int pool, const float* bn_const, const float* mask, float* bwd_const, float* dgamma,
                        float* dbeta, const uint32_t* grad_absmax, uint16_t* du_hi, uint16_t* du_lo, float* scratch_f,
                        float* dbias, double* red_scratch, void* stream);
/* The same with the sum over the ranks INSIDE the reduction kernel, over NVLink peer memory instead of a collective
 * call: every rank of the node (at most 8) owns one exchange buffer (vm_p2p_alloc) that all ranks have mapped through
 * CUDA IPC (vm_p2p_export -> 64-byte handle -> vm_p2p_import on the peers).  The block that finishes a 32-channel
 * column of the reduction pushes that column's sums into every peer's buffer, raises a flag, waits for the peers' flags
 * on its own buffer, adds the W vectors in rank order (bit-identical totals on every rank) and forms the constants --
 * no extra launch, no host involvement.  peers: W buffer addresses as mapped in THIS process (peers[rank] = the local
 * one); seq: 1, 2, 3, ... identical on every rank and incremented per call (all ranks must issue the same calls in the
 * same order); total_sums receives the global sums (2*G*C doubles). */
size_t vm_p2p_buffer_bytes(void);
int vm_p2p_alloc(void** ptr);
int vm_p2p_free(void* ptr);
int vm_p2p_export(void* ptr, unsigned char* handle64);
int vm_p2p_import(const unsigned char* handle64, void** ptr);
int vm_p2p_unimport(void* ptr);
/* vm_bn_stats_finalize / vm_bn_bwd with that exchange (arguments as there, plus the peer arguments).  local_sums
 * (2*G*C doubles; optional for the statistics), total_sums (2*G*C doubles), G <= 4. */
int vm_bn_stats_finalize_peers(const float* stat_partial, int rows_per_clip, int N, int G, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                               float* bn_const, double* red_scratch, void* const* peers, int rank, int world,
                               uint32_t seq, double count, double* local_sums, double* total_sums, void* stream);
int vm_bn_bwd_peers(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax,
                    const int32_t* jstar, int N, int L, int C, int G, int pool, const float* bn_const, const float* mask,
                    float* scratch_f2, float* bwd_const, float* dgamma, float* dbeta, uint32_t* grad_absmax,
                    uint16_t* du_hi, uint16_t* du_lo, float* scratch_f, float* dbias, double* red_scratch,
                    int presummed_rows_per_clip, void* const* peers, int rank, int world, uint32_t seq, double count,
                    double* local_sums, double* total_sums, void* stream);
/* dW (3, Cin, Cout) = sum_{n,p} X[n][p+tap-1][ci] * dU[n][p][co] on tensor cores.  x_hi / x_lo: the fp16 planes the
 * forward conv of the block consumed (x_lo = fp16 residual plane: training keeps precision-3 planes for this);
 * du_*: the scaled gradient planes of vm_bn_bwd, grad_absmax their scale word.  precision 3: Xh*Uh + Xl*Uh + Xh*Ul;
 * 2: one-plane gradient (du_lo NULL), Xh*Uh + Xl*Uh; 1: Xh*Uh.  partial: scratch. */
int vm_wgrad3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* du_hi, const uint16_t* du_lo, int N, int L,
              int cin, int cout, int precision, const uint32_t* grad_absmax, float* partial, size_t partial_bytes,
              float* dw, void* stream);
/* dW1 (32, 1, Cout) = sum_{n,p} x[n][p+k-15] * dU1[n][p][co] on tensor cores (fp16 Toeplitz operand of the waveform
 * built in shared memory).  precision 3: Uh*Th + Ul*Th + Uh*Tl; 2: one-plane gradient, Uh*Th + Uh*Tl; 1: Uh*Th. */
int vm_wgrad1(const float* x, const uint16_t* du_hi, const uint16_t* du_lo, int N, int L, int cout, int precision,
              const uint32_t* grad_absmax, float* partial, size_t partial_bytes, float* dw, void* stream);
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
