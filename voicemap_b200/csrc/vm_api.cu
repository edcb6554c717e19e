// C-ABI layer of libvoicemap_b200.so: argument checks, TMA descriptor creation, error strings, and the
// whole-encoder launcher.  See include/voicemap_b200.h for the contract of each entry point.
#include <stdio.h>
#include <string.h>

#include "vm_kernels.h"
#include "vm_p2p.cuh"
#include <string.h>

namespace vm {

static thread_local char g_err[512] = "no error";
static int g_max_ctas = 0;

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return VM_ERR_CUDA;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle, int f32) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return set_error(VM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(VM_ERR_SHAPE, "TMA base pointer not 16B aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) return set_error(VM_ERR_SHAPE, "TMA stride not a multiple of 16 bytes");
    gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle == VM_SWIZZLE_32B) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle == VM_SWIZZLE_64B) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle == VM_SWIZZLE_128B) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return VM_ERR_CUDA;
  }
  return VM_OK;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct EncoderPlan {
  int l1, l2, l3, t4, c4_pad;
  size_t a1, a2, a3, part, total;  // plane sizes in bytes (single plane), offsets derived below
};
static EncoderPlan plan_encoder(int N, int L, int filters, int first_pool) {
  EncoderPlan pl{};
  pl.l1 = L / first_pool; pl.l2 = pl.l1 / 2; pl.l3 = pl.l2 / 2;
  pl.t4 = 2 * ((pl.l3 + 255) / 256);  // two partial rows (column halves) per 256-position tile
  pl.c4_pad = (4 * filters + 127) / 128 * 128;
  pl.a1 = align_up(size_t(N) * pl.l1 * filters * 2, 1024);
  pl.a2 = align_up(size_t(N) * pl.l2 * 2 * filters * 2, 1024);
  pl.a3 = align_up(size_t(N) * pl.l3 * 3 * filters * 2, 1024);
  pl.part = align_up(size_t(N) * pl.t4 * pl.c4_pad * 4, 1024);
  pl.total = 2 * pl.a1 + 2 * pl.a2 + 2 * pl.a3 + pl.part;
  return pl;
}

}  // namespace vm

using namespace vm;

extern "C" {

int vm_version(void) { return 100; }
const char* vm_last_error_string(void) { return g_err; }

int vm_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceGetAttribute");
  if (major != 10) return set_error(VM_ERR_ARCH, "voicemap_b200 requires an sm_100 (B200) device");
  return VM_OK;
}

int vm_padded_channels(int cout) { return (cout + 127) / 128 * 128; }
size_t vm_conv1_wpack_bytes(int cout) { return size_t(vm_padded_channels(cout)) / 128 * 16384; }
size_t vm_conv3_wpack_bytes(int cin, int cout) { return size_t(3) * 3 * vm_padded_channels(cout) * cin * 2; }  // hi, lo, q
size_t vm_epi_bytes(int cout) { return size_t(vm_padded_channels(cout)) * 16; }
int vm_conv3_num_position_tiles(int L) { return 2 * ((L + 255) / 256); }

int vm_set_option(const char* key, int value) {
  if (key == nullptr) return set_error(VM_ERR_SHAPE, "vm_set_option: null key");
  int* slot = nullptr;
  if (strcmp(key, "max_ctas") == 0) slot = &g_max_ctas;
  if (slot == nullptr) return set_error(VM_ERR_SHAPE, "vm_set_option: unknown key");
  int old = *slot;
  *slot = value;
  return old;
}

int vm_pack_conv1(const float* kernel, const float* bias, const float* gamma, const float* beta, const float* mean,
                  const float* var, float eps, int cout, void* wpack, float* epi, void* stream) {
  return launch_pack_conv1(kernel, bias, gamma, beta, mean, var, eps, cout, wpack, epi, (cudaStream_t)stream);
}
int vm_pack_conv3(const float* kernel, const float* bias, const float* gamma, const float* beta, const float* mean,
                  const float* var, float eps, int cin, int cout, void* wpack, float* epi, void* stream) {
  return launch_pack_conv3(kernel, bias, gamma, beta, mean, var, eps, cin, cout, wpack, epi, (cudaStream_t)stream);
}

int vm_conv1_relu_bn_pool_fwd(const float* x, int N, int L, int cout, int pool, const void* wpack, const float* epi,
                              uint16_t* out_hi, uint16_t* out_lo, int precision, void* stream) {
  if (x == nullptr || wpack == nullptr || epi == nullptr || out_hi == nullptr)
    return set_error(VM_ERR_SHAPE, "conv1: null pointer");
  if (precision >= 2 && out_lo == nullptr) return set_error(VM_ERR_SHAPE, "conv1: out_lo required for precision 2 / 3");
  return launch_conv1(x, N, L, cout, wpack, epi, reinterpret_cast<__half*>(out_hi),
                      reinterpret_cast<__half*>(out_lo), nullptr, nullptr, nullptr, nullptr, precision, g_max_ctas,
                      (cudaStream_t)stream, 1, 0, nullptr, nullptr, pool);
}

int vm_conv3_relu_bn_pool2_fwd(const uint16_t* in_hi, const uint16_t* in_lo, int N, int L, int cin, int cout,
                               const void* wpack, const float* epi, uint16_t* out_hi, uint16_t* out_lo,
                               float* gmax_partial, int precision, void* stream) {
  if (in_hi == nullptr || wpack == nullptr || epi == nullptr) return set_error(VM_ERR_SHAPE, "conv3: null pointer");
  if (gmax_partial == nullptr && precision >= 2 && out_lo == nullptr)
    return set_error(VM_ERR_SHAPE, "conv3: out_lo required for precision 2 / 3");
  return launch_conv3(reinterpret_cast<const __half*>(in_hi), reinterpret_cast<const __half*>(in_lo), N, L, cin, cout,
                      static_cast<const __half*>(wpack), epi, reinterpret_cast<__half*>(out_hi),
                      reinterpret_cast<__half*>(out_lo), gmax_partial, nullptr, nullptr, 0, precision, g_max_ctas,
                      (cudaStream_t)stream);
}

int vm_gmax_dense_fwd(const float* gmax_partial, int N, int T, int C, const float* epi, const float* dense_w,
                      const float* dense_b, int E, float* gmax_out, float* emb, void* stream) {
  if (gmax_partial == nullptr || epi == nullptr) return set_error(VM_ERR_SHAPE, "gmax_dense: null pointer");
  return launch_gmax_dense(gmax_partial, N, T, C, vm_padded_channels(C), epi, dense_w, dense_b, E, gmax_out, emb,
                           (cudaStream_t)stream);
}

int vm_pair_head_loss_fwd(const float* e1, const float* e2, int N, int E, int metric, const float* head_w,
                          const float* head_b, const float* y_true, int loss_kind, float* dist, float* prob,
                          float* loss, void* stream) {
  if (e1 == nullptr || e2 == nullptr || head_w == nullptr || head_b == nullptr)
    return set_error(VM_ERR_SHAPE, "pair_head_loss: null pointer");
  return launch_pair_head_loss(e1, e2, N, E, metric, head_w, head_b, y_true, loss_kind, dist, prob, loss,
                               (cudaStream_t)stream);
}

int vm_nshot_score(const float* query, const float* support, int T, int k, int n, int E, int distance, float* scores,
                   int32_t* best, void* stream) {
  return launch_nshot_score(query, support, T, k, n, E, distance, scores, best, (cudaStream_t)stream);
}

int vm_split_planes(const float* x, size_t n, uint16_t* hi, uint16_t* lo, void* stream) {
  if (x == nullptr || hi == nullptr) return set_error(VM_ERR_SHAPE, "split_planes: null pointer");
  return launch_split_planes(x, n, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo),
                             (cudaStream_t)stream);
}
int vm_merge_planes(const uint16_t* hi, const uint16_t* lo, size_t n, float* x, void* stream) {
  if (x == nullptr || hi == nullptr) return set_error(VM_ERR_SHAPE, "merge_planes: null pointer");
  return launch_merge_planes(reinterpret_cast<const __half*>(hi), reinterpret_cast<const __half*>(lo), n, x,
                             (cudaStream_t)stream);
}

int vm_split_planes_q(const float* x, size_t n, uint16_t* hi, uint16_t* q, void* stream) {
  if (x == nullptr || hi == nullptr || q == nullptr) return set_error(VM_ERR_SHAPE, "split_planes_q: null pointer");
  return launch_split_planes_q(x, n, reinterpret_cast<__half*>(hi), q, (cudaStream_t)stream);
}
int vm_merge_planes_q(const uint16_t* hi, const uint16_t* q, size_t n, float* x, void* stream) {
  if (x == nullptr || hi == nullptr || q == nullptr) return set_error(VM_ERR_SHAPE, "merge_planes_q: null pointer");
  return launch_merge_planes_q(reinterpret_cast<const __half*>(hi), q, n, x, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------
// training entry points
// ---------------------------------------------------------------------------------------------------------
#define ST ((cudaStream_t)stream)
#define H16(p) reinterpret_cast<__half*>(p)
#define CH16(p) reinterpret_cast<const __half*>(p)

int vm_pack_conv1_raw(const float* kernel, const float* bias, int cout, void* wpack, float* epi, void* stream) {
  return launch_pack_conv1(kernel, bias, nullptr, nullptr, nullptr, nullptr, 0.f, cout, wpack, epi, ST);
}
int vm_pack_conv3_raw(const float* kernel, const float* bias, int cin, int cout, void* wpack, float* epi,
                      void* stream) {
  return launch_pack_conv3(kernel, bias, nullptr, nullptr, nullptr, nullptr, 0.f, cin, cout, wpack, epi, ST);
}
int vm_pack_conv3_dgrad(const float* kernel, int cin, int cout, void* wpack, float* epi, void* stream) {
  return launch_pack_conv3_dgrad(kernel, cin, cout, wpack, epi, ST);
}
int vm_pack_train(const float* const* kernels, const float* const* biases, int filters, void* const* wraw,
                  float* const* eraw, void* const* wdg, float* const* edg, void* stream) {
  return launch_pack_train(kernels, biases, filters, wraw, eraw, wdg, edg, ST);
}
int vm_stat_rows_per_clip(int L) { return 2 * ((L + 255) / 256); }
int vm_conv3_train_rows_per_clip(int L) { return 4 * ((L + 255) / 256); }   // two epilogue sets x two column halves
size_t vm_reduce_scratch_bytes(int G, int C) { return size_t(G > 0 ? G : 1) * 32 * size_t(C) * 16; }

int vm_conv1_train_fwd(const float* x, int N, int L, int cout, int pool, const void* wpack, const float* epi,
                       const float* gamma, uint16_t* u16, float* ext, float* stat_partial, int precision, void* stream) {
  if (x == nullptr || wpack == nullptr || epi == nullptr || u16 == nullptr || ext == nullptr)
    return set_error(VM_ERR_SHAPE, "conv1_train: null pointer");
  return launch_conv1(x, N, L, cout, wpack, epi, nullptr, nullptr, u16, ext, gamma, stat_partial, precision,
                      g_max_ctas, ST, 1, 0, nullptr, nullptr, pool);
}
int vm_conv3_train_fwd(const uint16_t* in_hi, const uint16_t* in_lo, int N, int L, int cin, int cout,
                       const void* wpack, const float* epi, const float* gamma, uint16_t* u16, float* ext,
                       float* stat_partial, int precision, void* stream) {
  if (in_hi == nullptr || wpack == nullptr || epi == nullptr || u16 == nullptr || ext == nullptr)
    return set_error(VM_ERR_SHAPE, "conv3_train: null pointer");
  Conv3Extra ex;
  ex.out_u16 = u16; ex.out_ext = ext; ex.sign_src = gamma;
  return launch_conv3(CH16(in_hi), CH16(in_lo), N, L, cin, cout, static_cast<const __half*>(wpack), epi, nullptr,
                      nullptr, nullptr, nullptr, stat_partial, 0, precision, g_max_ctas, ST, ex);
}
int vm_conv3_dgrad(const uint16_t* du_hi, const uint16_t* du_lo, int N, int L, int cout, int cin,
                   const void* wpack_dgrad, const float* epi_dgrad, const uint32_t* grad_absmax, float* dx,
                   int precision, const float* below_ext, const float* below_bn_const, const float* below_mask,
                   int below_groups, float* below_partial, uint32_t* below_absmax, void* stream) {
  if (du_hi == nullptr || wpack_dgrad == nullptr || epi_dgrad == nullptr || dx == nullptr)
    return set_error(VM_ERR_SHAPE, "conv3_dgrad: null pointer");
  if (precision < 1 || precision > 3) return set_error(VM_ERR_SHAPE, "conv3_dgrad: precision must be 1, 2 or 3");
  if (precision == 3 && du_lo == nullptr) return set_error(VM_ERR_SHAPE, "conv3_dgrad: du_lo required for precision 3");
  Conv3Extra ex;
  ex.grad_absmax = grad_absmax;
  ex.x_single = (precision == 2) ? 1 : 0;
  if (below_partial != nullptr) {
    ex.red.ext = below_ext;
    ex.red.bn_const = reinterpret_cast<const float4*>(below_bn_const);
    ex.red.mask = below_mask;
    ex.red.G = below_groups;
    ex.red.partial = reinterpret_cast<float2*>(below_partial);
    ex.red.absmax = below_absmax;
  }
  // the gradient is the "input" operand of this convolution (channels = Cout of the forward conv), the output has Cin
  return launch_conv3(CH16(du_hi), CH16(du_lo), N, L, cout, cin, static_cast<const __half*>(wpack_dgrad), epi_dgrad,
                      nullptr, nullptr, nullptr, dx, nullptr, /*linear=*/1, precision == 1 ? 1 : 3, g_max_ctas, ST, ex);
}
int vm_bn_stats_finalize(const float* stat_partial, int rows_per_clip, int N, int G, int L, int C,
                         const float* gamma, const float* beta, float eps, float momentum, float* moving_mean,
                         float* moving_var, float* bn_const, double* red_scratch, void* stream) {
  return launch_bn_stats_finalize(stat_partial, rows_per_clip, vm_padded_channels(C), N, G, L, C, gamma, beta, eps,
                                  momentum, moving_mean, moving_var, bn_const, red_scratch, ST);
}
int vm_bn_pool_fwd(const float* ext, int N, int Lout, int C, int G, const float* bn_const, const float* mask,
                   uint16_t* out_hi, uint16_t* out_lo, uint16_t* out_q, void* stream) {
  if (ext == nullptr || bn_const == nullptr || out_hi == nullptr) return set_error(VM_ERR_SHAPE, "bn_pool_fwd: null pointer");
  return launch_bn_pool_fwd(ext, N, Lout, C, G, bn_const, mask, H16(out_hi), H16(out_lo), out_q, ST);
}
int vm_bn_gmax_fwd(const float* ext, int N, int Lout, int C, int G, const float* bn_const, const float* mask,
                   float* gmax, int32_t* jstar, void* stream) {
  return launch_bn_gmax_fwd(ext, N, Lout, C, G, bn_const, mask, gmax, jstar, ST);
}
int vm_dense_fwd(const float* x, int N, int C, const float* w, const float* b, int E, float* y, void* stream) {
  return launch_dense_fwd(x, N, C, w, b, E, y, ST);
}
int vm_pair_head_loss_bwd(const float* emb, int N, int E, int metric, const float* head_w, const float* head_b,
                          const float* y_true, int loss_kind, float loss_scale, float* d_emb, float* d_head_w,
                          float* d_head_b, float* accuracy, void* stream) {
  return launch_pair_head_loss_bwd(emb, N, E, metric, head_w, head_b, y_true, loss_kind, loss_scale, d_emb, d_head_w,
                                   d_head_b, accuracy, ST);
}
int vm_dense_bwd(const float* x, const float* dy, const float* w, int N, int C, int E, float* dw, float* db, float* dx,
                 void* stream) {
  return launch_dense_bwd(x, dy, w, N, C, E, dw, db, dx, ST);
}
int vm_siamese_head_train(const float* gmax, int N, int C, int E, const float* dense_w, const float* dense_b, int metric,
                          const float* head_w, const float* head_b, const float* y_true, int loss_kind, float loss_scale,
                          float* emb, float* prob, float* d_emb, float* d_gmax, float* pair_scratch, float* d_dense_w,
                          float* d_dense_b, float* d_head_w, float* d_head_b, float* loss_acc, void* stream) {
  return launch_siamese_head_train(gmax, N, C, E, dense_w, dense_b, metric, head_w, head_b, y_true, loss_kind, loss_scale,
                                   emb, prob, d_emb, d_gmax, pair_scratch, d_dense_w, d_dense_b, d_head_w, d_head_b,
                                   loss_acc, ST);
}
size_t vm_bn_bwd_scratch_elems(int N) { return bn_bwd_scratch_elems(N); }
int vm_bn_bwd(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax, const int32_t* jstar,
              int N, int L, int C, int G, int pool, const float* bn_const, const float* mask, float* scratch_f2,
              float* bwd_const, float* dgamma, float* dbeta, uint32_t* grad_absmax, uint16_t* du_hi, uint16_t* du_lo,
              float* scratch_f, float* dbias, double* red_scratch, int presummed_rows_per_clip, void* stream) {
  return launch_bn_bwd(u16, ext, dy_pooled, d_gmax, jstar, N, L, C, G, pool, bn_const, mask, scratch_f2, bwd_const,
                       dgamma, dbeta, grad_absmax, H16(du_hi), H16(du_lo), scratch_f, dbias, red_scratch,
                       presummed_rows_per_clip, ST);
}
int vm_bn_stats_sums(const float* stat_partial, int rows_per_clip, int N, int G, int C, double* red_scratch,
                     double* sums, void* stream) {
  return launch_bn_stats_sums(stat_partial, rows_per_clip, vm_padded_channels(C), N, G, C, red_scratch, sums, ST);
}
int vm_bn_stats_from_sums(const double* sums, double count, int G, int C, const float* gamma, const float* beta,
                          float eps, float momentum, float* moving_mean, float* moving_var, float* bn_const,
                          void* stream) {
  return launch_bn_stats_from_sums(sums, count, G, C, gamma, beta, eps, momentum, moving_mean, moving_var, bn_const, ST);
}
int vm_bn_bwd_sums(const float* ext, const float* dy_pooled, const float* d_gmax, const int32_t* jstar, int N, int L,
                   int C, int G, int pool, const float* bn_const, const float* mask, float* scratch_f2,
                   uint32_t* grad_absmax, double* red_scratch, double* sums, int presummed_rows_per_clip,
                   void* stream) {
  return launch_bn_bwd_sums(ext, dy_pooled, d_gmax, jstar, N, L, C, G, pool, bn_const, mask, scratch_f2, grad_absmax,
                            red_scratch, sums, presummed_rows_per_clip, ST);
}
int vm_bn_bwd_from_sums(const double* local_sums, const double* global_sums, double count, const uint16_t* u16,
                        const float* dy_pooled, const float* d_gmax, const int32_t* jstar, int N, int L, int C, int G,
                        int pool, const float* bn_const, const float* mask, float* bwd_const, float* dgamma,
                        float* dbeta, const uint32_t* grad_absmax, uint16_t* du_hi, uint16_t* du_lo, float* scratch_f,
                        float* dbias, double* red_scratch, void* stream) {
  return launch_bn_bwd_from_sums(local_sums, global_sums, count, u16, dy_pooled, d_gmax, jstar, N, L, C, G, pool,
                                 bn_const, mask, bwd_const, dgamma, dbeta, grad_absmax, H16(du_hi), H16(du_lo),
                                 scratch_f, dbias, red_scratch, ST);
}
int vm_wgrad3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* du_hi, const uint16_t* du_lo, int N, int L,
              int cin, int cout, int precision, const uint32_t* grad_absmax, float* partial, size_t partial_bytes,
              float* dw, void* stream) {
  return launch_wgrad3(CH16(x_hi), CH16(x_lo), CH16(du_hi), CH16(du_lo), N, L, cin, cout, precision, partial,
                       partial_bytes, dw, grad_absmax, ST);
}
int vm_wgrad1(const float* x, const uint16_t* du_hi, const uint16_t* du_lo, int N, int L, int cout, int precision,
              const uint32_t* grad_absmax, float* partial, size_t partial_bytes, float* dw, void* stream) {
  return launch_wgrad1(x, CH16(du_hi), CH16(du_lo), N, L, cout, partial, partial_bytes, dw, grad_absmax, ST, precision);
}
// ---- peer exchange buffers (CUDA IPC) and the BatchNorm calls that sum over the ranks inside the kernel -------------
size_t vm_p2p_buffer_bytes(void) { return kP2PBufferBytes; }
int vm_p2p_alloc(void** ptr) {
  if (ptr == nullptr) return set_error(VM_ERR_SHAPE, "p2p_alloc: null pointer");
  cudaError_t e = cudaMalloc(ptr, kP2PBufferBytes);        // own allocation: an IPC handle names a whole cudaMalloc block
  if (e != cudaSuccess) return set_cuda_error(e, "p2p_alloc: cudaMalloc");
  e = cudaMemset(*ptr, 0, kP2PBufferBytes);
  if (e != cudaSuccess) return set_cuda_error(e, "p2p_alloc: cudaMemset");
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return set_cuda_error(e, "p2p_alloc: sync");
  return VM_OK;
}
int vm_p2p_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? VM_OK : set_cuda_error(e, "p2p_free");
}
int vm_p2p_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (ptr == nullptr || handle64 == nullptr) return set_error(VM_ERR_SHAPE, "p2p_export: null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return set_cuda_error(e, "p2p_export: cudaIpcGetMemHandle");
  memcpy(handle64, &h, 64);
  return VM_OK;
}
int vm_p2p_import(const unsigned char* handle64, void** ptr) {
  if (ptr == nullptr || handle64 == nullptr) return set_error(VM_ERR_SHAPE, "p2p_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? VM_OK : set_cuda_error(e, "p2p_import: cudaIpcOpenMemHandle (peer access over NVLink needed)");
}
int vm_p2p_unimport(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? VM_OK : set_cuda_error(e, "p2p_unimport");
}
int vm_bn_stats_finalize_peers(const float* stat_partial, int rows_per_clip, int N, int G, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                               float* bn_const, double* red_scratch, void* const* peers, int rank, int world,
                               uint32_t seq, double count, double* local_sums, double* total_sums, void* stream) {
  return launch_bn_stats_finalize_peers(stat_partial, rows_per_clip, vm_padded_channels(C), N, G, C, red_scratch, peers,
                                        rank, world, seq, count, gamma, beta, eps, momentum, moving_mean, moving_var,
                                        bn_const, local_sums, total_sums, ST);
}
int vm_bn_bwd_peers(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax,
                    const int32_t* jstar, int N, int L, int C, int G, int pool, const float* bn_const, const float* mask,
                    float* scratch_f2, float* bwd_const, float* dgamma, float* dbeta, uint32_t* grad_absmax,
                    uint16_t* du_hi, uint16_t* du_lo, float* scratch_f, float* dbias, double* red_scratch,
                    int presummed_rows_per_clip, void* const* peers, int rank, int world, uint32_t seq, double count,
                    double* local_sums, double* total_sums, void* stream) {
  return launch_bn_bwd_peers(u16, ext, dy_pooled, d_gmax, jstar, N, L, C, G, pool, bn_const, mask, scratch_f2, bwd_const,
                             dgamma, dbeta, grad_absmax, H16(du_hi), H16(du_lo), scratch_f, dbias, red_scratch,
                             presummed_rows_per_clip, peers, rank, world, seq, count, local_sums, total_sums, ST);
}
int vm_adam_step(float* p, const float* g, float* m, float* v, size_t n, double* scratch, float inv_scale,
                 float clipnorm, float lr_t, float beta1, float beta2, float eps, void* stream) {
  return launch_adam_step(p, g, m, v, n, scratch, inv_scale, clipnorm, lr_t, beta1, beta2, eps, ST);
}
#undef ST
#undef H16
#undef CH16

size_t vm_encoder_workspace_bytes(int N, int L, int filters, int first_pool) {
  if (N <= 0 || filters <= 0 || (first_pool != 2 && first_pool != 4) || L < 8 * first_pool) return 0;
  return plan_encoder(N, L, filters, first_pool).total;
}

static int encoder_fwd_impl(const float* x, int N, int L, int filters, int first_pool, const void* const* wpack,
                            const float* const* epi, const float* dense_w, const float* dense_b, int E,
                            void* workspace, float* emb, int precision, cudaStream_t st, int x_stride,
                            long long x_clip_stride, const float* pre_mean, const float* pre_scale) {
  const EncoderPlan pl = plan_encoder(N, L, filters, first_pool);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* a1h = reinterpret_cast<__half*>(ws);
  __half* a1l = reinterpret_cast<__half*>(ws + pl.a1);
  __half* a2h = reinterpret_cast<__half*>(ws + 2 * pl.a1);
  __half* a2l = reinterpret_cast<__half*>(ws + 2 * pl.a1 + pl.a2);
  __half* a3h = reinterpret_cast<__half*>(ws + 2 * pl.a1 + 2 * pl.a2);
  __half* a3l = reinterpret_cast<__half*>(ws + 2 * pl.a1 + 2 * pl.a2 + pl.a3);
  float* part = reinterpret_cast<float*>(ws + 2 * pl.a1 + 2 * pl.a2 + 2 * pl.a3);
  const int f = filters;
  int rc;
  if ((rc = launch_conv1(x, N, L, f, wpack[0], epi[0], a1h, a1l, nullptr, nullptr, nullptr, nullptr, precision, g_max_ctas, st, x_stride,
                         x_clip_stride, pre_mean, pre_scale, first_pool)))
    return rc;
  if ((rc = launch_conv3(a1h, a1l, N, pl.l1, f, 2 * f, static_cast<const __half*>(wpack[1]), epi[1], a2h, a2l,
                         nullptr, nullptr, nullptr, 0, precision, g_max_ctas, st)))
    return rc;
  if ((rc = launch_conv3(a2h, a2l, N, pl.l2, 2 * f, 3 * f, static_cast<const __half*>(wpack[2]), epi[2], a3h, a3l,
                         nullptr, nullptr, nullptr, 0, precision, g_max_ctas, st)))
    return rc;
  if ((rc = launch_conv3(a3h, a3l, N, pl.l3, 3 * f, 4 * f, static_cast<const __half*>(wpack[3]), epi[3], nullptr,
                         nullptr, part, nullptr, nullptr, 0, precision, g_max_ctas, st)))
    return rc;
  // (finishing GlobalMaxPool1D + Dense inside block 4's kernel -- the CTA that completes a clip's last tile, found by a
  // ticket per clip -- was built and measured: bit-identical, but 0.033 ms SLOWER per 256 clips than this 0.017 ms
  // launch: the ticket and the L2-latency-bound Dense sit in the epilogue warps' path)
  return launch_gmax_dense(part, N, pl.t4, 4 * f, pl.c4_pad, epi[3], dense_w, dense_b, E, nullptr, emb, st);
}

static int encoder_args_ok(const void* x, const void* wpack, const void* epi, const void* workspace, const void* emb,
                           int N, int L, int filters, int first_pool) {
  if (x == nullptr || wpack == nullptr || epi == nullptr || workspace == nullptr || emb == nullptr)
    return set_error(VM_ERR_SHAPE, "encoder: null pointer");
  if (N <= 0 || filters <= 0) return set_error(VM_ERR_SHAPE, "encoder: bad shape");
  if (first_pool != 2 && first_pool != 4)
    return set_error(VM_ERR_UNSUPPORTED, "encoder: first MaxPool1D size must be 4 (voicemap/models.py:19) or 2");
  if (L < 8 * first_pool) return set_error(VM_ERR_SHAPE, "encoder: L must cover the four pooling stages (p*2*2*2)");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0)
    return set_error(VM_ERR_SHAPE, "encoder: workspace must be 1024-byte aligned");
  return VM_OK;
}

int vm_encoder_fwd(const float* x, int N, int L, int filters, int first_pool, const void* const* wpack,
                   const float* const* epi, const float* dense_w, const float* dense_b, int E, void* workspace,
                   float* emb, int precision, void* stream) {
  int rc = encoder_args_ok(x, wpack, epi, workspace, emb, N, L, filters, first_pool);
  if (rc) return rc;
  return encoder_fwd_impl(x, N, L, filters, first_pool, wpack, epi, dense_w, dense_b, E, workspace, emb, precision,
                          (cudaStream_t)stream, 1, 0, nullptr, nullptr);
}

size_t vm_preprocess_scratch_bytes(int N) { return N > 0 ? align_up(size_t(N) * 4 * 6 + 64, 1024) : 0; }

int vm_preprocess_stats(const float* x, int N, int T, int downsampling, int G, float rms, float* mean, float* scale,
                        void* stream) {
  if (x == nullptr || mean == nullptr || scale == nullptr) return set_error(VM_ERR_SHAPE, "preprocess: null pointer");
  return launch_preprocess_stats(x, N, T, downsampling, G, rms, mean, scale, (cudaStream_t)stream);
}

int vm_encoder_fwd_raw(const float* x, int N, int T, int downsampling, int whiten_groups, float rms, int filters,
                       int first_pool, const void* const* wpack, const float* const* epi, const float* dense_w,
                       const float* dense_b, int E, void* workspace, float* emb, int precision, void* stream) {
  if (downsampling <= 0 || T <= 0) return set_error(VM_ERR_SHAPE, "encoder_raw: bad downsampling / length");
  const int L = (T + downsampling - 1) / downsampling;  // numpy x[:, ::d] keeps ceil(T / d) samples
  int rc = encoder_args_ok(x, wpack, epi, workspace, emb, N, L, filters, first_pool);
  if (rc) return rc;
  const float* mean = nullptr;
  const float* scale = nullptr;
  if (whiten_groups > 0) {
    if (N % whiten_groups != 0) return set_error(VM_ERR_SHAPE, "encoder_raw: N must be a multiple of whiten_groups");
    float* m = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + plan_encoder(N, L, filters, first_pool).total);
    float* sc = m + N;
    if ((rc = launch_preprocess_stats(x, N, T, downsampling, whiten_groups, rms, m, sc, (cudaStream_t)stream)))
      return rc;
    mean = m;
    scale = sc;
  }
  return encoder_fwd_impl(x, N, L, filters, first_pool, wpack, epi, dense_w, dense_b, E, workspace, emb, precision,
                          (cudaStream_t)stream, downsampling, T, mean, scale);
}

}  // extern "C"
