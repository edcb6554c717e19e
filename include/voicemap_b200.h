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
 *     uses the hi plane only (throughput mode; lo pointers may be NULL).
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
size_t vm_conv3_wpack_bytes(int cin, int cout); /* packed block-2..4 weights */
size_t vm_epi_bytes(int cout);                  /* per-channel epilogue constants (padded to 128 channels) */
int vm_conv3_num_position_tiles(int L);         /* T of the gmax_partial tensor: 2 * ceil(L / 256) */
int vm_padded_channels(int cout);

/* ---- weight preparation ---------------------------------------------------------------------------------
 * Replaces the (implicit) Keras weight layout of Conv1D + BatchNormalization (voicemap/models.py:13-35):
 * kernel (K, Cin, Cout), bias, BN gamma/beta/moving_mean/moving_variance, eps (Keras default 1e-3) are folded
 * into fp16 (hi, lo) weight planes in the MMA operand layout and per-channel constants {sigma, bias, s, t} with
 * s = gamma / sqrt(var + eps), t = beta - mean * s, sigma = sign(s) (weights are stored sigma-scaled so that
 * max-pooling commutes with the BN affine when gamma < 0). */
int vm_pack_conv1(const float* kernel /* (32, 1, Cout) */, const float* bias, const float* gamma,
                  const float* beta, const float* mean, const float* var, float eps, int cout, void* wpack,
                  float* epi, void* stream);
int vm_pack_conv3(const float* kernel /* (3, Cin, Cout) */, const float* bias, const float* gamma,
                  const float* beta, const float* mean, const float* var, float eps, int cin, int cout, void* wpack,
                  float* epi, void* stream);

/* ---- block 1: Conv1D(filters, 32, 'same', relu) -> BatchNormalization -> MaxPool1D(4, 4) ------------------
 * voicemap/models.py:13-19.  x (N, L) fp32 (Keras (N, L, 1)); out planes (N, L/4, Cout). */
int vm_conv1_relu_bn_pool4_fwd(const float* x, int N, int L, int cout, const void* wpack, const float* epi,
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

/* ---- whole encoder (voicemap/models.py:6-41, eval mode) ---------------------------------------------------
 * x (N, L) fp32 -> emb (N, E).  `wpack[i]`, `epi[i]` (i = 0..3) from vm_pack_conv1 / vm_pack_conv3 for channel
 * widths filters*{1,2,3,4}.  `workspace` must hold vm_encoder_workspace_bytes(N, L, filters) bytes. */
size_t vm_encoder_workspace_bytes(int N, int L, int filters);
int vm_encoder_fwd(const float* x, int N, int L, int filters, const void* const* wpack, const float* const* epi,
                   const float* dense_w, const float* dense_b, int E, void* workspace, float* emb, int precision,
                   void* stream);

/* ---- testing / tuning knobs -------------------------------------------------------------------------------
 * key "max_ctas" (0 = all SMs; limits the persistent grid).  Returns the previous value or VM_ERR_SHAPE. */
int vm_set_option(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* VOICEMAP_B200_H */
