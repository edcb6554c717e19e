// Internal declarations shared by the kernel translation units and the C-ABI layer (vm_api.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/voicemap_b200.h"

namespace vm {

// ---- error reporting (thread-local last error string, C-ABI: vm_last_error_string) ----
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int num_sms();

// ---- TMA descriptor creation through the driver entry point (no link-time libcuda dependency) ----
enum { VM_SWIZZLE_NONE = 0, VM_SWIZZLE_32B = 1, VM_SWIZZLE_64B = 2, VM_SWIZZLE_128B = 3 };
int make_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box, int swizzle,
                    int f32 = 0);

// ---- block 1 ----
struct Conv1Params {
  const float* x;        // (N, L) fp32 (Keras (N, L, 1)), or raw audio when x_stride > 1
  int N, L;              // L = samples per clip entering the convolution (after decimation)
  int x_stride;          // decimation stride into x (voicemap/utils.py:29), 1 = none
  long long x_clip_stride;  // elements between consecutive clips in x
  const float* pre_mean; // fused whitening: sample = (x - pre_mean[n]) * pre_scale[n]; null = off
  const float* pre_scale;
  int cout, cout_pad, nslab;
  int lout;              // L / 4
  int nptile;            // ceil(L / 256)
  int products;          // 3: fp16x3 (fp32-grade), 2: same MMAs, output planes (fp16 hi, e5m2x2 Q), 1: fp16x1
  int nstages;           // Toeplitz ring depth (2..4, limited by shared memory when Cout > 128)
  const uint4* wpack;    // [slab][plane][8 KB smem image]
  const float4* epi;     // [cout_pad] {sigma, bias, s, t}
  __half* out_hi;        // (N, lout, cout)
  __half* out_lo;
  // train-mode forward (out_u16 != null): encoded un-pooled activations (N, L, cout) (encode_u, vm_common.cuh) and
  // the fp32 extreme of every pool window (N, lout, cout); sign_src[cout] < 0 selects the minimum, null = maxima
  uint16_t* out_u16;
  float* out_ext;
  const float* sign_src;
  float2* stat_partial;  // (N*nptile*2, cout_pad) {sum, sum of squares} over valid positions, or null
};
int launch_conv1(const float* x, int N, int L, int cout, const void* wpack, const float* epi, __half* out_hi,
                 __half* out_lo, uint16_t* out_u16, float* out_ext, const float* sign_src, float* stat_partial,
                 int products, int max_ctas, cudaStream_t stream, int x_stride = 1, long long x_clip_stride = 0,
                 const float* pre_mean = nullptr, const float* pre_scale = nullptr, int pool = 4);
int launch_preprocess_stats(const float* x, int N, int T, int stride, int G, float rms, float* mean, float* scale,
                            cudaStream_t stream);

// ---- blocks 2-4 ----
struct Conv3Reduce {
  const float* ext = nullptr;       // (N, L, cout) window extremes of the block below (its pooled positions = this L)
  const float4* bn_const = nullptr; // (G, cout) {s, t, mean, rstd} of the block below
  const float* mask = nullptr;      // (N, cout) dropout mask of the block below, or null
  int G = 1;
  float2* partial = nullptr;        // (N * 2 * nptile, cout_pad) {sum dy, sum dy * xhat} rows, one per (tile, column half)
  unsigned int* absmax = nullptr;   // raised with atomicMax (float bits of max |s * dy|)
};
struct Conv3Params {
  int N, L, cin, cout, cout_pad;
  int lout;              // L / 2
  int nptile;            // ceil(L / 256)
  int nslab;             // cout_pad / 128
  int nchunk;            // ceil(cin / 64)
  int products;
  const float4* epi;     // [cout_pad]
  __half* out_hi;        // (N, lout, cout) or null
  __half* out_lo;
  float* gmax_partial;   // (N, 2*nptile, cout_pad) raw accumulator maxima, or null
  float* out_f32;        // (N, L, cout) un-pooled fp32 output (train-mode forward / dgrad), or null
  float2* stat_partial;  // (N, 2*nptile, cout_pad) {sum, sum of squares} of out_f32 over valid positions, or null
  int linear;            // out_f32 mode: 1 = store the raw accumulator (dgrad), 0 = apply the epilogue constants
  int x_single;          // products 3 with a single input plane: Xh*Wh + Xh*Wl (dgrad of a one-plane gradient)
  int slabs_per_unit;    // tile schedule: 1, or nslab when the X tile stays in shared memory for all cout slabs
  // train-mode forward (out_u16 != null): encoded un-pooled activations (N, L, cout) (encode_u) + fp32 window
  // extremes (N, lout, cout); sign_src[cout] < 0 selects the minimum (negative BatchNorm scale), null = all maxima
  uint16_t* out_u16;
  float* out_ext;
  const float* sign_src;
  const uint32_t* grad_absmax;  // dgrad: bits of the largest |s * dy| of the block -> power-of-two unscale, or null
  // dgrad epilogue, optional: the BatchNorm-backward reduction of the block BELOW (whose pooled gradient this kernel
  // writes) -- per-channel partial sums of dy and dy * xhat and the largest |s * dy|, taken while dy is in registers
  Conv3Reduce red;
};
struct Conv3Extra {            // optional arguments of launch_conv3 beyond the eval-forward set
  uint16_t* out_u16 = nullptr;
  float* out_ext = nullptr;
  const float* sign_src = nullptr;
  const uint32_t* grad_absmax = nullptr;
  int x_single = 0;
  Conv3Reduce red = Conv3Reduce();
};
int launch_conv3(const __half* in_hi, const __half* in_lo, int N, int L, int cin, int cout, const __half* wpack,
                 const float* epi, __half* out_hi, __half* out_lo, float* gmax_partial, float* out_f32,
                 float* stat_partial, int linear, int products, int max_ctas, cudaStream_t stream,
                 const Conv3Extra& extra = Conv3Extra());

// ---- weight gradients (vm_wgrad.cu) ----
struct Wgrad3Params {
  int N, L, cin, cout, products;
  int nchunk;            // ceil(L / 64) position chunks per clip
  int nco_tiles, ncombo; // (ci slab, co tile) work items
  int nsplit;            // reduction splits over (clip, chunk) steps
  float* partial;        // [nsplit][3][cin][cout]
};
int launch_wgrad3(const __half* x_hi, const __half* x_lo, const __half* du_hi, const __half* du_lo, int N, int L,
                  int cin, int cout, int products, float* partial, size_t partial_bytes, float* dw,
                  const unsigned int* grad_absmax, cudaStream_t stream);
int launch_wgrad1(const float* x, const __half* du_hi, const __half* du_lo, int N, int L, int cout, float* partial,
                  size_t partial_bytes, float* dw, const unsigned int* grad_absmax, cudaStream_t stream,
                  int products = 3);
int launch_wgrad1_tc(const float* x, const __half* du_hi, const __half* du_lo, int N, int L, int cout, int products,
                     float* partial, size_t partial_bytes, int* nsplit_out, cudaStream_t stream);

// ---- training elementwise / reduction kernels (vm_train.cu) ----
int launch_bn_stats_finalize(const float* partial, int rows_per_clip, int c_pad, int N, int G, int L, int C,
                             const float* gamma, const float* beta, float eps, float momentum, float* moving_mean,
                             float* moving_var, float* bn_const, double* red_scratch, cudaStream_t st);
int launch_bn_pool_fwd(const float* ext, int N, int lout, int C, int G, const float* bn_const, const float* mask,
                       __half* out_hi, __half* out_lo, uint16_t* out_q, cudaStream_t st);
int launch_bn_gmax_fwd(const float* ext, int N, int lout, int C, int G, const float* bn_const, const float* mask,
                       float* gmax, int* jstar, cudaStream_t st);
int launch_dense_fwd(const float* x, int N, int C, const float* w, const float* b, int E, float* y, cudaStream_t st);
int launch_dense_bwd(const float* x, const float* dy, const float* w, int N, int C, int E, float* dw, float* db,
                     float* dx, cudaStream_t st);
int launch_pair_head_loss_bwd(const float* emb, int N, int E, int metric, const float* head_w, const float* head_b,
                              const float* y_true, int loss_kind, float loss_scale, float* d_emb, float* d_head_w,
                              float* d_head_b, float* accuracy, cudaStream_t st);
int launch_bn_stats_finalize_peers(const float* partial, int rows_per_clip, int c_pad, int N, int G, int C,
                                   double* red_scratch, void* const* peers, int rank, int world, unsigned int seq,
                                   double count, const float* gamma, const float* beta, float eps, float momentum,
                                   float* moving_mean, float* moving_var, float* bn_const, double* local_sums,
                                   double* total_sums, cudaStream_t st);
int launch_bn_bwd_peers(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax,
                        const int* jstar, int N, int L, int C, int G, int pool, const float* bn_const, const float* mask,
                        float* partial, float* bwd_const, float* dgamma, float* dbeta, unsigned int* absmax,
                        __half* du_hi, __half* du_lo, float* dbias_partial, float* dbias, double* red_scratch,
                        int presummed_rows, void* const* peers, int rank, int world, unsigned int seq, double count,
                        double* local_sums, double* total_sums, cudaStream_t st);
int launch_siamese_head_train(const float* gmax, int N, int C, int E, const float* dense_w, const float* dense_b,
                              int metric, const float* head_w, const float* head_b, const float* y_true, int loss_kind,
                              float loss_scale, float* emb, float* prob, float* d_emb, float* d_gmax, float* pair_scratch,
                              float* d_dense_w, float* d_dense_b, float* d_head_w, float* d_head_b, float* loss_acc,
                              cudaStream_t st);
size_t bn_bwd_scratch_elems(int N);
int launch_bn_bwd(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax, const int* jstar,
                  int N, int L, int C, int G, int pool, const float* bn_const, const float* mask, float* partial,
                  float* bwd_const, float* dgamma, float* dbeta, unsigned int* absmax, __half* du_hi, __half* du_lo,
                  float* dbias_partial, float* dbias, double* red_scratch, int presummed_rows, cudaStream_t st);
// split forms for synchronised BatchNorm (sums -> caller's all-reduce -> constants)
int launch_bn_stats_sums(const float* partial, int rows_per_clip, int c_pad, int N, int G, int C, double* red_scratch,
                         double* sums, cudaStream_t st);
int launch_bn_stats_from_sums(const double* sums, double count, int G, int C, const float* gamma, const float* beta,
                              float eps, float momentum, float* moving_mean, float* moving_var, float* bn_const,
                              cudaStream_t st);
int launch_bn_bwd_sums(const float* ext, const float* dy_pooled, const float* d_gmax, const int* jstar, int N, int L,
                       int C, int G, int pool, const float* bn_const, const float* mask, float* partial,
                       unsigned int* absmax, double* red_scratch, double* sums, int presummed_rows, cudaStream_t st);
int launch_bn_bwd_from_sums(const double* local_sums, const double* global_sums, double count, const uint16_t* u16,
                            const float* dy_pooled, const float* d_gmax, const int* jstar, int N, int L, int C, int G,
                            int pool, const float* bn_const, const float* mask, float* bwd_const, float* dgamma,
                            float* dbeta, const unsigned int* absmax, __half* du_hi, __half* du_lo,
                            float* dbias_partial, float* dbias, double* red_scratch, cudaStream_t st);
// cross-rank forms: the sum over the ranks happens inside the kernel, over NVLink peer memory (vm_p2p.cuh)
int launch_adam_step(float* p, const float* g, float* m, float* v, size_t n, double* sumsq_scratch, float inv_scale,
                     float clipnorm, float lr_t, float beta1, float beta2, float eps, cudaStream_t st);
int launch_pack_conv3_dgrad(const float* w, int cin, int cout, void* wpack, float* epi, cudaStream_t stream);

// ---- small kernels (vm_head.cu) ----
struct PackTrainTask {
  const float* w;
  const float* bias;
  void* wpack;
  float* epi;
  int cin, cout, pad, kind;   // kind 0: conv1 raw, 1: conv3 raw, 2: conv3 dgrad; pad = padded Cout (Cin for dgrad)
  unsigned block0;            // first block of the task's range
};
struct PackTrainArgs {
  PackTrainTask t[7];
  int ntasks;
};
int launch_pack_train(const float* const* kernels, const float* const* biases, int filters, void* const* wraw,
                      float* const* eraw, void* const* wdg, float* const* edg, cudaStream_t stream);
int launch_pack_conv1(const float* w, const float* bias, const float* gamma, const float* beta, const float* mean,
                      const float* var, float eps, int cout, void* wpack, float* epi, cudaStream_t stream);
int launch_pack_conv3(const float* w, const float* bias, const float* gamma, const float* beta, const float* mean,
                      const float* var, float eps, int cin, int cout, void* wpack, float* epi, cudaStream_t stream);
int launch_split_planes(const float* x, size_t n, __half* hi, __half* lo, cudaStream_t stream);
int launch_merge_planes(const __half* hi, const __half* lo, size_t n, float* x, cudaStream_t stream);
int launch_split_planes_q(const float* x, size_t n, __half* hi, uint16_t* q, cudaStream_t stream);
int launch_merge_planes_q(const __half* hi, const uint16_t* q, size_t n, float* x, cudaStream_t stream);
int launch_gmax_dense(const float* partial, int N, int T, int C, int c_pad, const float* epi, const float* dense_w,
                      const float* dense_b, int E, float* gmax_out, float* emb, cudaStream_t stream);
int launch_pair_head_loss(const float* e1, const float* e2, int N, int E, int metric, const float* head_w,
                          const float* head_b, const float* y_true, int loss_kind, float* dist, float* prob,
                          float* loss, cudaStream_t stream);
int launch_nshot_score(const float* query, const float* support, int T, int k, int n, int E, int distance,
                       float* scores, int* best, cudaStream_t stream);

}  // namespace vm
