// Training-mode kernels around the tensor-core convolutions (CUDA cores; all HBM-bound elementwise / reduction work).
//
// What a block keeps from its forward pass (written by the conv kernels' train-mode epilogues, vm_conv1/3.cu):
//   u16 (N, L, C)      the un-pooled activation u = relu(conv + bias) as fp16 + an arg-max flag in the sign bit
//                      (encode_u, vm_common.cuh) -- 2 bytes per element, the only full-resolution tensor of the block;
//   ext (N, L/p, C)    the fp32 extreme of u in every MaxPool window: the maximum, or the minimum where the BatchNorm
//                      scale is negative (sign(gamma) is known before the batch statistics are).  BatchNorm is
//                      monotone per channel, so the pooled output is s * ext + t exactly as if the affine had been
//                      applied before the pool (voicemap/models.py:17-19);
//   {sum, sumsq}       per-channel partials of u for the batch statistics.
// Forward (train):  conv -> bn_stats_finalize (batch moments per BN group, Keras moving-average update)
//                   -> bn_pool_fwd (y = s*ext + t, SpatialDropout mask -> fp16 planes (hi, lo | Q) of the next conv)
//                      / bn_gmax_fwd for block 4 (GlobalMaxPool1D over the windows, with the winning window)
//                   -> siamese_head_train (embedding Dense, pair head, loss and their backward: two launches)
//                      [classifier: dense_fwd, torch softmax head, dense_bwd]
// Backward:         per block (4..1):
//                   sum dy, sum dy*xhat (dy is non-zero at the window arg-max only, where u == ext, so these come from
//                   pooled tensors only -- taken in the dgrad epilogue of the block above, block 4: bn_bwd_reduce; also
//                   the largest |s*dy| -> gradient scale)
//                   -> column reduction + finishing step (+ dgamma, dbeta; data parallel: + the exchange with the peers)
//                   -> bn_relu_bwd (u16 + dy -> dU = relu'(u) * s * (dy - mean_dy - xhat*mean_dyxhat) as scaled fp16
//                      planes, + conv-bias gradient partials)
//                   -> wgrad (vm_wgrad.cu) and dgrad (conv3_kernel on flipped/transposed weights).
// Reference semantics: keras BatchNormalization / SpatialDropout1D / MaxPool1D / GlobalMaxPool1D / Dense in training
// mode as used by voicemap/models.py:6-81; losses voicemap/utils.py:77-85 and keras binary_crossentropy.
//
// A "BN group" is one application of the shared encoder (the siamese net applies it once per branch, so batch
// statistics are per branch: voicemap/models.py:52-53); clips [g*N/G, (g+1)*N/G) form group g.
// All gradients flowing through these kernels carry the loss scale the caller folded into the loss gradient.
#include "vm_common.cuh"
#include "vm_kernels.h"
#include "vm_p2p.cuh"

namespace vm {

static int check_launch_t(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return VM_OK;
}

// ---------------------------------------------------------------------------------------------
// Deterministic two-stage column reduction of a (rows, stride) fp32 matrix with NC interleaved components per entry
// (NC = 2: float2 {a, b}; NC = 1: float).  Stage 1: grid (ceil(C/32), G*kRB), block (32, 8): row block rb of group
// g -> tmp[(g*kRB + rb)][c] (double2).  Stage 2 (the finishing functors): sum the kRB partials per (g, c).
// ---------------------------------------------------------------------------------------------
constexpr int kRB = 32;

__device__ __forceinline__ double2 rowsum_stage2(const double2* __restrict__ tmp, int g, int C, int c) {
  // all kRB loads are issued before the first add (the sum order stays rb = 0, 1, ...: deterministic)
  double2 v[kRB];
#pragma unroll
  for (int rb = 0; rb < kRB; ++rb) v[rb] = tmp[(size_t(g) * kRB + rb) * C + c];
  double a = 0.0, b = 0.0;
#pragma unroll
  for (int rb = 0; rb < kRB; ++rb) {
    a += v[rb].x;
    b += v[rb].y;
  }
  return make_double2(a, b);
}

// Both stages in ONE launch: the stage-1 grid as above, and the block that finishes last among the G*kRB blocks of a
// channel column (a self-resetting ticket per column) runs the finishing functor `fin(tmp, c)` for its 32 channels.
// The stage-2 sums are taken in the fixed order rb = 0, 1, ... whichever block comes last, so the result does not depend
// on the schedule.  One training step runs at a time per device (the tickets are per device, not per stream).
__device__ unsigned int g_rowsum_tickets[64];   // columns of 32 channels: C <= 2048
template <int NC, class Fin>
__global__ void rowsum_fused_kernel(const float* __restrict__ m, size_t rows_per_group, int stride, int C,
                                    double2* __restrict__ tmp, const Fin fin) {
  __shared__ double2 sm[8][32];
  __shared__ bool last;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y / kRB, rb = blockIdx.y % kRB;
  const size_t per = (rows_per_group + kRB - 1) / kRB;
  const size_t r0 = size_t(g) * rows_per_group + size_t(rb) * per;
  const size_t r1 = min(size_t(g + 1) * rows_per_group, r0 + per);
  double a = 0.0, b = 0.0;
  if (c < C) {
    // four rows in flight per thread, added in row order (one load per iteration made this a chain of memory latencies)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 32) {
      float2 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t ri = r + 8 * i;
        v[i] = make_float2(0.f, 0.f);
        if (ri < r1) {
          if (NC == 2) v[i] = reinterpret_cast<const float2*>(m)[ri * stride + c];
          else v[i].x = m[ri * stride + c];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a += double(v[i].x);
        if (NC == 2) b += double(v[i].y);
      }
    }
  }
  sm[threadIdx.y][threadIdx.x] = make_double2(a, b);
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int k = 1; k < 8; ++k) { a += sm[k][threadIdx.x].x; b += sm[k][threadIdx.x].y; }
    tmp[size_t(blockIdx.y) * C + c] = make_double2(a, b);
  }
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();                                   // this block's partials before its ticket
    const unsigned int t = atomicAdd(&g_rowsum_tickets[blockIdx.x], 1u);
    last = (t == gridDim.y - 1);
    if (last) g_rowsum_tickets[blockIdx.x] = 0u;       // ready for the next launch
  }
  __syncthreads();
  if (last && threadIdx.y == 0) {
    __threadfence();                                   // the other blocks' partials after their tickets
    if constexpr (Fin::kWarp) fin(tmp, c);             // warp-collective finisher (all 32 lanes; it checks c < C itself)
    else if (c < C) fin(tmp, c);
  }
}

// ---------------------------------------------------------------------------------------------
// batch statistics -> BN constants {s, t, mean, rstd} per (group, channel); moving-average update
// ---------------------------------------------------------------------------------------------
struct BnStatsFin {
  static constexpr bool kWarp = false;
  int N, G, L, C;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* moving_mean;
  float* moving_var;
  float4* bn_const;
  __device__ void operator()(const double2* __restrict__ tmp, int c) const {
    const int clips = N / G;
    const double cnt = double(clips) * double(L);
    float mm = moving_mean ? moving_mean[c] : 0.f;
    float mv = moving_var ? moving_var[c] : 0.f;
    for (int g = 0; g < G; ++g) {
      const double2 sums = rowsum_stage2(tmp, g, C, c);
      const double s1 = sums.x, s2 = sums.y;
      const double mean = s1 / cnt;
      double var = s2 / cnt - mean * mean;  // biased batch variance (tf.nn.moments)
      if (var < 0.0) var = 0.0;
      const float rstd = float(1.0 / sqrt(var + double(eps)));
      const float s = gamma[c] * rstd;
      bn_const[size_t(g) * C + c] = make_float4(s, beta[c] - float(mean) * s, float(mean), rstd);
      // keras: sample variance var * n / (n - (1 + eps)); moving <- moving * momentum + stat * (1 - momentum)
      const float var_unbiased = float(var * (cnt / (cnt - (1.0 + double(eps)))));
      mm = mm * momentum + float(mean) * (1.f - momentum);
      mv = mv * momentum + var_unbiased * (1.f - momentum);
    }
    if (moving_mean) moving_mean[c] = mm;
    if (moving_var) moving_var[c] = mv;
  }
};

// ---------------------------------------------------------------------------------------------
// Split form of the statistics for synchronised BatchNorm across ranks: (1) per-(group, channel) double sums,
// (2) the caller all-reduces them, (3) constants from the global sums and the global count.
// ---------------------------------------------------------------------------------------------
struct SumsFin {
  static constexpr bool kWarp = false;
  int G, C;
  double2* sums;
  __device__ void operator()(const double2* __restrict__ tmp, int c) const {
    for (int g = 0; g < G; ++g) sums[size_t(g) * C + c] = rowsum_stage2(tmp, g, C, c);
  }
};
__device__ __forceinline__ void bn_stats_from_sums_channel(int c, const double2* __restrict__ sums, double cnt, int G,
                                                           int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float momentum,
                                                           float* __restrict__ moving_mean,
                                                           float* __restrict__ moving_var,
                                                           float4* __restrict__ bn_const) {
  float mm = moving_mean ? moving_mean[c] : 0.f;
  float mv = moving_var ? moving_var[c] : 0.f;
  for (int g = 0; g < G; ++g) {   // same arithmetic as bn_stats_finalize_kernel
    const double2 sm = sums[size_t(g) * C + c];
    const double mean = sm.x / cnt;
    double var = sm.y / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = float(1.0 / sqrt(var + double(eps)));
    const float s = gamma[c] * rstd;
    bn_const[size_t(g) * C + c] = make_float4(s, beta[c] - float(mean) * s, float(mean), rstd);
    const float var_unbiased = float(var * (cnt / (cnt - (1.0 + double(eps)))));
    mm = mm * momentum + float(mean) * (1.f - momentum);
    mv = mv * momentum + var_unbiased * (1.f - momentum);
  }
  if (moving_mean) moving_mean[c] = mm;
  if (moving_var) moving_var[c] = mv;
}
__global__ void bn_stats_from_sums_kernel(const double2* __restrict__ sums, double cnt, int G, int C,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                          float momentum, float* __restrict__ moving_mean,
                                          float* __restrict__ moving_var, float4* __restrict__ bn_const) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_stats_from_sums_channel(c, sums, cnt, G, C, gamma, beta, eps, momentum, moving_mean, moving_var, bn_const);
}
// backward: the batch means of dy and dy*xhat come from the GLOBAL sums, dgamma / dbeta from this rank's own sums
// (the gradient all-reduce adds the ranks' contributions)
__device__ __forceinline__ void bn_bwd_from_sums_channel(int c, const double2* __restrict__ local,
                                                         const double2* __restrict__ global, double cnt, int G, int C,
                                                         const float4* __restrict__ bn_const,
                                                         float4* __restrict__ bwd_const, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta) {
  double tg = 0.0, tb = 0.0;
  for (int g = 0; g < G; ++g) {
    const double2 gs = global[size_t(g) * C + c], ls = local[size_t(g) * C + c];
    bwd_const[size_t(g) * C + c] =
        make_float4(bn_const[size_t(g) * C + c].x, float(gs.x / cnt), float(gs.y / cnt), 0.f);
    tb += ls.x;
    tg += ls.y;
  }
  dgamma[c] = float(tg);
  dbeta[c] = float(tb);
}
// Column finishers with the cross-rank sum inside (p2p_allreduce_column): the reduction kernel's last block of every
// 32-channel column exchanges that column's sums and derives the constants -- reduction, exchange and constants are
// ONE launch (they were two: *_sums, then a one-block *_sync kernel).
constexpr int kPeersMaxGroups = 4;
struct BnStatsPeersFin {
  static constexpr bool kWarp = true;
  int G, C;
  double cnt;
  P2PPeers peers;
  unsigned int seq;
  double2* local;   // (G, C) this rank's sums (kept for inspection; may be null)
  double2* total;   // (G, C) sums over the ranks
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* moving_mean;
  float* moving_var;
  float4* bn_const;
  __device__ void operator()(const double2* __restrict__ tmp, int c) const {
    double2 mine[kPeersMaxGroups];
    for (int g = 0; g < G; ++g) {
      mine[g] = (c < C) ? rowsum_stage2(tmp, g, C, c) : make_double2(0.0, 0.0);
      if (local != nullptr && c < C) local[size_t(g) * C + c] = mine[g];
    }
    p2p_allreduce_column<kPeersMaxGroups>(mine, G, C, c, blockIdx.x, peers, seq, total);
    if (c < C)
      bn_stats_from_sums_channel(c, total, cnt, G, C, gamma, beta, eps, momentum, moving_mean, moving_var, bn_const);
  }
};
struct BnBwdPeersFin {
  static constexpr bool kWarp = true;
  int G, C;
  double cnt;
  P2PPeers peers;
  unsigned int seq;
  double2* local;   // (G, C) this rank's sums: dgamma / dbeta come from them
  double2* total;
  const float4* bn_const;
  float4* bwd_const;
  float* dgamma;
  float* dbeta;
  __device__ void operator()(const double2* __restrict__ tmp, int c) const {
    double2 mine[kPeersMaxGroups];
    for (int g = 0; g < G; ++g) {
      mine[g] = (c < C) ? rowsum_stage2(tmp, g, C, c) : make_double2(0.0, 0.0);
      if (c < C) local[size_t(g) * C + c] = mine[g];
    }
    p2p_allreduce_column<kPeersMaxGroups>(mine, G, C, c, blockIdx.x, peers, seq, total);
    if (c < C) bn_bwd_from_sums_channel(c, local, total, cnt, G, C, bn_const, bwd_const, dgamma, dbeta);
  }
};
__global__ void bn_bwd_from_sums_kernel(const double2* __restrict__ local, const double2* __restrict__ global,
                                        double cnt, int G, int C, const float4* __restrict__ bn_const,
                                        float4* __restrict__ bwd_const, float* __restrict__ dgamma,
                                        float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_bwd_from_sums_channel(c, local, global, cnt, G, C, bn_const, bwd_const, dgamma, dbeta);
}

// ---------------------------------------------------------------------------------------------
// Thread layout of the elementwise passes: grid (clip, position chunk); inside a block of 256 threads, thread
// (channel group cg, stream) owns kPer adjacent channels for the whole launch -- its BatchNorm constants live in
// registers -- and walks the positions stream, stream + streams, ... of the chunk.
// ---------------------------------------------------------------------------------------------
constexpr int kEwThreads = 256;
__host__ __device__ constexpr int ew_streams(int C, int per) { return (kEwThreads / (C / per)) > 0 ? kEwThreads / (C / per) : 1; }
static int ew_chunks(int N, int items, int streams) {
  // ~8 blocks per SM, but no more chunks than there are stream passes to hand out
  int chunks = (kNumSMs * 8 + N - 1) / N;
  const int most = (items + streams - 1) / streams;
  if (chunks > most) chunks = most;
  return chunks < 1 ? 1 : chunks;
}
__device__ __forceinline__ float f4get(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return uint32_t(__half_as_ushort(a)) | (uint32_t(__half_as_ushort(b)) << 16);
}

// BN affine (+ dropout mask) on the window extremes -> planes of the next conv: fp16 hi always, the fp16 residual
// plane (kLo: forward precision 3, and the weight gradient's second activation plane) and / or the e5m2x2 Q plane
// (kQ: forward precision 2).
template <bool kLo, bool kQ>
__global__ void __launch_bounds__(kEwThreads)
bn_pool_fwd_kernel(const float* __restrict__ ext, int N, int lout, int C, int G, const float4* __restrict__ bn_const,
                   const float* __restrict__ mask, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                   uint16_t* __restrict__ out_q) {
  const int n = blockIdx.x, chunk = blockIdx.y, chunks = gridDim.y;
  const int groups = C >> 3, streams = ew_streams(C, 8);
  if (int(threadIdx.x) >= groups * streams) return;
  const int cg = threadIdx.x % groups, stream = threadIdx.x / groups, c = 8 * cg;
  const int g = n / (N / G);
  float sv[8], tv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 bc = bn_const[size_t(g) * C + c + k];
    const float mk = mask ? mask[size_t(n) * C + c + k] : 1.f;
    sv[k] = bc.x * mk;
    tv[k] = bc.y * mk;
  }
  const int per = (lout + chunks - 1) / chunks;
  const int j0 = chunk * per, j1 = min(lout, j0 + per);
  auto finish = [&](int j, const float4& a, const float4& b) {
    __half h[8], l[8];
    uint16_t q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float y = fmaf(sv[k], k < 4 ? f4get(a, k) : f4get(b, k - 4), tv[k]);
      split_f32(y, h[k], l[k]);
      if (kQ) split_f16_q(y, h[k], q[k]);   // same hi
    }
    const size_t o = (size_t(n) * lout + j) * C + c;
    *reinterpret_cast<uint4*>(out_hi + o) =
        make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
    if (kLo)
      *reinterpret_cast<uint4*>(out_lo + o) =
          make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
    if (kQ)
      *reinterpret_cast<uint4*>(out_q + o) =
          make_uint4(uint32_t(q[0]) | (uint32_t(q[1]) << 16), uint32_t(q[2]) | (uint32_t(q[3]) << 16),
                     uint32_t(q[4]) | (uint32_t(q[5]) << 16), uint32_t(q[6]) | (uint32_t(q[7]) << 16));
  };
  // four positions in flight per thread (32 bytes each)
  for (int j = j0 + stream; j < j1; j += 4 * streams) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int jj = j + i * streams;
      if (jj < j1) {
        const float4* src = reinterpret_cast<const float4*>(ext + (size_t(n) * lout + jj) * C + c);
        a[i] = __ldcs(src);
        b[i] = __ldcs(src + 1);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (j + i * streams < j1) finish(j + i * streams, a[i], b[i]);
  }
}

// Block 4: BN affine (+ mask) on the MaxPool(2) window extremes -> GlobalMaxPool1D over the windows.
// grid (N, ceil(C/32)); block (32 channels, 8 window lanes).  Writes the max and the winning window (first on ties).
__global__ void bn_gmax_fwd_kernel(const float* __restrict__ ext, int N, int lout, int C, int G,
                                   const float4* __restrict__ bn_const, const float* __restrict__ mask,
                                   float* __restrict__ gmax, int* __restrict__ jstar) {
  __shared__ float smax[8][32];
  __shared__ int sidx[8][32];
  const int n = blockIdx.x;
  const int c = blockIdx.y * 32 + threadIdx.x;
  const int g = n / (N / G);
  float best = -INFINITY;
  int bi = 0;
  if (c < C) {
    const float4 bc = bn_const[size_t(g) * C + c];
    const float mk = mask ? mask[size_t(n) * C + c] : 1.f;
    const float s = bc.x * mk, t = bc.y * mk;
    const float* col = ext + size_t(n) * lout * C + c;
    int j = threadIdx.y;
    for (; j + 56 < lout; j += 64) {   // eight windows in flight per thread, compared in window order
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldcs(col + size_t(j + 8 * i) * C);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = fmaf(s, v[i], t);
        if (y > best) { best = y; bi = j + 8 * i; }
      }
    }
    for (; j < lout; j += 8) {
      const float y = fmaf(s, col[size_t(j) * C], t);
      if (y > best) { best = y; bi = j; }
    }
  }
  smax[threadIdx.y][threadIdx.x] = best;
  sidx[threadIdx.y][threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int k = 1; k < 8; ++k) {
      const float v = smax[k][threadIdx.x];
      const int i = sidx[k][threadIdx.x];
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    gmax[size_t(n) * C + c] = best;
    jstar[size_t(n) * C + c] = bi;
  }
}

// ---------------------------------------------------------------------------------------------
// Dense forward / backward (embedding layer, voicemap/models.py:39)
// ---------------------------------------------------------------------------------------------
// One block per clip, 256 threads = 4 channel quarters x 64 outputs: every thread accumulates a quarter of the dot
// product with 32 weight loads in flight (a single serial 512-term chain per output took 65 us for 32 clips), the
// quarters are summed in a fixed order.
__global__ void __launch_bounds__(256)
dense_fwd_kernel(const float* __restrict__ x, int N, int C, const float* __restrict__ w,
                 const float* __restrict__ b, int E, float* __restrict__ y) {
  extern __shared__ float xs[];   // [C] inputs, then [4][64] partial dot products
  float* red = xs + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) xs[c] = x[size_t(n) * C + c];
  __syncthreads();
  const int part = threadIdx.x >> 6, lane_e = threadIdx.x & 63;
  const int cq = (C + 3) / 4;
  const int c0 = part * cq, c1 = min(C, c0 + cq);
  for (int e0 = 0; e0 < E; e0 += 64) {
    const int e = e0 + lane_e;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (e < E) {
      int c = c0;
      for (; c + 32 <= c1; c += 32) {
        float wv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) wv[i] = __ldg(w + size_t(c + i) * E + e);
#pragma unroll
        for (int i = 0; i < 32; ++i) a4[i & 3] = fmaf(xs[c + i], wv[i], a4[i & 3]);
      }
      for (; c < c1; ++c) a4[0] = fmaf(xs[c], __ldg(w + size_t(c) * E + e), a4[0]);
    }
    red[part * 64 + lane_e] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    __syncthreads();
    if (part == 0 && e < E)
      y[size_t(n) * E + e] = ((red[lane_e] + red[64 + lane_e]) + (red[128 + lane_e] + red[192 + lane_e])) + b[e];
    __syncthreads();
  }
}
// dW[c][e] = sum_n x[n][c] * dy[n][e];  db[e] = sum_n dy[n][e] (block 0).  Block = 8 channels x E outputs (one thread
// per e, strided): a dy row is loaded once per clip for the 8 channels, four clips in flight (the first version ran one
// serial chain of N dependent loads per output and took 47 us for 128 clips).
__global__ void __launch_bounds__(64)
dense_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int C, int E,
                   float* __restrict__ dw, float* __restrict__ db) {
  const int c0 = blockIdx.x * 8;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accb = 0.f;
    int n = 0;
    for (; n + 4 <= N; n += 4) {
      float d[4], xv[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        d[i] = dy[size_t(n + i) * E + e];
#pragma unroll
        for (int k = 0; k < 8; ++k) xv[i][k] = (c0 + k < C) ? __ldg(x + size_t(n + i) * C + c0 + k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {   // clips in order: the sums do not depend on the unrolling
        accb += d[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv[i][k], d[i], acc[k]);
      }
    }
    for (; n < N; ++n) {
      const float d = dy[size_t(n) * E + e];
      accb += d;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (c0 + k < C) acc[k] = fmaf(x[size_t(n) * C + c0 + k], d, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c0 + k < C) dw[size_t(c0 + k) * E + e] = acc[k];
    if (blockIdx.x == 0) db[e] = accb;
  }
}
// dx[n][c] = sum_e dy[n][e] * w[c][e]   (grid N blocks, C threads strided)
__global__ void dense_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ w, int N, int C, int E,
                                   float* __restrict__ dx) {
  extern __shared__ float ds[];
  const int n = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) ds[e] = dy[size_t(n) * E + e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(ds[e], w[size_t(c) * E + e], acc);
    dx[size_t(n) * C + c] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Siamese head + loss backward.  emb (2N, E): rows [0,N) branch 1, [N,2N) branch 2.  One block.
// d_emb (2N, E), d_head_w (1 or E), d_head_b (1); everything multiplied by loss_scale.
// ---------------------------------------------------------------------------------------------
__global__ void pair_head_loss_bwd_kernel(const float* __restrict__ emb, int N, int E, int metric,
                                          const float* __restrict__ head_w, const float* __restrict__ head_b,
                                          const float* __restrict__ y_true, int loss_kind, float loss_scale,
                                          float* __restrict__ d_emb, float* __restrict__ d_head_w,
                                          float* __restrict__ d_head_b, float* __restrict__ accuracy) {
  extern __shared__ float dzs[];  // [N] dL/dz (scaled), then [N] distance
  float* dist = dzs + N;
  __shared__ int hits;
  if (threadIdx.x == 0) hits = 0;
  __syncthreads();
  int my_hits = 0;
  const float* e1 = emb;
  const float* e2 = emb + size_t(N) * E;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* a = e1 + size_t(n) * E;
    const float* b = e2 + size_t(n) * E;
    float z, d = 0.f;
    if (metric == 0) {
      float ss = 0.f;
      for (int j = 0; j < E; ++j) { const float t = a[j] - b[j]; ss = fmaf(t, t, ss); }
      d = sqrtf(fmaxf(ss, 0.f));
      z = fmaf(d, head_w[0], head_b[0]);
    } else {
      float acc = 0.f;
      for (int j = 0; j < E; ++j) acc = fmaf(fabsf(a[j] - b[j]), head_w[j], acc);
      z = acc + head_b[0];
    }
    const float p = 1.0f / (1.0f + expf(-z));
    const float y = y_true[n];
    my_hits += ((p > 0.5f ? 1.f : 0.f) == y) ? 1 : 0;     // keras 'accuracy' of a sigmoid output: round(p) == y
    float dp;
    if (loss_kind == 1) {  // contrastive: (1-y) p^2 + y max(1-p,0)^2
      dp = (1.f - y) * 2.f * p - y * 2.f * fmaxf(1.f - p, 0.f);
    } else {               // BCE on clip(p, 1e-7, 1-1e-7): zero gradient outside the clip range
      const float lo = 1e-7f, hi = 1.0f - 1e-7f;
      dp = (p < lo || p > hi) ? 0.f : (-y / p + (1.f - y) / (1.f - p));
    }
    dzs[n] = dp * p * (1.f - p) * (loss_scale / float(N));
    dist[n] = d;
  }
  if (accuracy != nullptr && my_hits) atomicAdd(&hits, my_hits);   // integer count: order does not matter
  __syncthreads();
  if (accuracy != nullptr && threadIdx.x == 0) accuracy[0] = float(hits) / float(N);
  // d_emb
  for (int idx = threadIdx.x; idx < N * E; idx += blockDim.x) {
    const int n = idx / E, j = idx % E;
    const float diff = e1[idx] - e2[idx];
    float g;
    if (metric == 0) {
      const float d = dist[n];
      g = (d > 0.f) ? dzs[n] * head_w[0] * diff / d : 0.f;  // sqrt'(0) guarded (the reference yields NaN there)
    } else {
      g = dzs[n] * head_w[j] * ((diff > 0.f) - (diff < 0.f));
    }
    d_emb[idx] = g;
    d_emb[size_t(N) * E + idx] = -g;
  }
  // head gradients
  if (metric == 0) {
    if (threadIdx.x == 0) {
      float gw = 0.f, gb = 0.f;
      for (int n = 0; n < N; ++n) { gw = fmaf(dzs[n], dist[n], gw); gb += dzs[n]; }
      d_head_w[0] = gw;
      d_head_b[0] = gb;
    }
  } else {
    for (int j = threadIdx.x; j < E; j += blockDim.x) {
      float gw = 0.f;
      for (int n = 0; n < N; ++n) gw = fmaf(dzs[n], fabsf(e1[size_t(n) * E + j] - e2[size_t(n) * E + j]), gw);
      d_head_w[j] = gw;
    }
    if (threadIdx.x == 0) {
      float gb = 0.f;
      for (int n = 0; n < N; ++n) gb += dzs[n];
      d_head_b[0] = gb;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The siamese training head in two launches (was five: dense_fwd, pair_head_loss, pair_head_loss_bwd, dense_bwd_w,
// dense_bwd_x -- 86 us of a 64-pair step, most of it launch and drain latency of one-block kernels).
// A pair's loss term depends on its own two clips only, so everything up to d loss / d gmax runs per pair:
//   kernel 1, one block per pair n: Dense (both clips, the arithmetic of dense_fwd_kernel) -> distance -> sigmoid ->
//     loss term, hit, dL/dz -> d_emb (branch 2 is the exact negative of branch 1) -> d_gmax = d_emb . W^T;
//   kernel 2: Dense weight gradient (blocks of 8 channels, the arithmetic of dense_bwd_w_kernel) and, in one extra
//     block, the batch reductions in a fixed order: loss mean, accuracy, head gradients.
// Semantics: voicemap/models.py:39,52-69 (Dense, the K-lambdas of the head, Dense(1, sigmoid)), voicemap/utils.py:77-85.
// ---------------------------------------------------------------------------------------------
struct SiameseHeadArgs {
  const float* gmax;      // (2N, C) GlobalMaxPool output: branch 1 rows, then branch 2 rows
  int N, C, E, metric, loss_kind;
  const float* dense_w;   // (C, E)
  const float* dense_b;   // (E)
  const float* head_w;    // (1) uniform_euclidean, (E) weighted_l1
  const float* head_b;    // (1)
  const float* y_true;    // (N)
  float loss_scale;
  float* emb;             // (2N, E)
  float* prob;            // (N) or null
  float* d_emb;           // (2N, E)
  float* d_gmax;          // (2N, C)
  float4* pair;           // (N) {loss term, hit, dL/dz (scaled), distance}
  float* d_dense_w;       // (C, E)
  float* d_dense_b;       // (E)
  float* d_head_w;
  float* d_head_b;
  float* loss_acc;        // [2] {mean loss, accuracy}
};

// y[e] = sum_c xs[c] * w[c][e] + b[e] for e < E: 256 threads = 4 channel quarters x 64 outputs (see dense_fwd_kernel)
__device__ __forceinline__ void dense_row_256(const float* __restrict__ xs, int C, const float* __restrict__ w,
                                              const float* __restrict__ b, int E, float* __restrict__ red,
                                              float* __restrict__ ys, float* __restrict__ yg) {
  const int part = threadIdx.x >> 6, lane_e = threadIdx.x & 63;
  const int cq = (C + 3) / 4;
  const int c0 = part * cq, c1 = min(C, c0 + cq);
  for (int e0 = 0; e0 < E; e0 += 64) {
    const int e = e0 + lane_e;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (e < E) {
      int c = c0;
      for (; c + 32 <= c1; c += 32) {
        float wv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) wv[i] = __ldg(w + size_t(c + i) * E + e);
#pragma unroll
        for (int i = 0; i < 32; ++i) a4[i & 3] = fmaf(xs[c + i], wv[i], a4[i & 3]);
      }
      for (; c < c1; ++c) a4[0] = fmaf(xs[c], __ldg(w + size_t(c) * E + e), a4[0]);
    }
    red[part * 64 + lane_e] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    __syncthreads();
    if (part == 0 && e < E) {
      const float y = ((red[lane_e] + red[64 + lane_e]) + (red[128 + lane_e] + red[192 + lane_e])) + b[e];
      ys[e] = y;
      yg[e] = y;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
siamese_head_pair_kernel(const SiameseHeadArgs a) {
  extern __shared__ float sh[];   // [2][C] gmax rows, [2][E] embeddings, [E] d_emb of branch 1, [256] partials
  const int n = blockIdx.x, N = a.N, C = a.C, E = a.E;
  float* xs = sh;
  float* es = xs + 2 * C;
  float* de = es + 2 * E;
  float* red = de + E;
  __shared__ float s_dz, s_dist;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    xs[c] = a.gmax[size_t(n) * C + c];
    xs[C + c] = a.gmax[size_t(N + n) * C + c];
  }
  __syncthreads();
  dense_row_256(xs, C, a.dense_w, a.dense_b, E, red, es, a.emb + size_t(n) * E);
  dense_row_256(xs + C, C, a.dense_w, a.dense_b, E, red, es + E, a.emb + size_t(N + n) * E);
  if (threadIdx.x == 0) {   // the pair's scalar chain, in the summation order of pair_head_loss_kernel
    float z, d = 0.f;
    if (a.metric == 0) {
      float ss = 0.f;
      for (int j = 0; j < E; ++j) { const float t = es[j] - es[E + j]; ss = fmaf(t, t, ss); }
      d = sqrtf(fmaxf(ss, 0.f));
      z = fmaf(d, a.head_w[0], a.head_b[0]);
    } else {
      float acc = 0.f;
      for (int j = 0; j < E; ++j) acc = fmaf(fabsf(es[j] - es[E + j]), a.head_w[j], acc);
      z = acc + a.head_b[0];
    }
    const float p = 1.0f / (1.0f + expf(-z));
    const float y = a.y_true[n];
    float term, dp;
    if (a.loss_kind == 1) {   // contrastive: (1-y) p^2 + y max(1-p,0)^2
      const float mg = fmaxf(1.0f - p, 0.f);
      term = (1.f - y) * p * p + y * mg * mg;
      dp = (1.f - y) * 2.f * p - y * 2.f * mg;
    } else {                  // BCE on clip(p, 1e-7, 1-1e-7): zero gradient outside the clip range
      const float lo = 1e-7f, hi = 1.0f - 1e-7f;
      const float pc = fminf(fmaxf(p, lo), hi);
      term = -y * logf(pc) - (1.f - y) * logf(1.0f - pc);
      dp = (p < lo || p > hi) ? 0.f : (-y / p + (1.f - y) / (1.f - p));
    }
    const float dz = dp * p * (1.f - p) * (a.loss_scale / float(N));
    const float hit = ((p > 0.5f ? 1.f : 0.f) == y) ? 1.f : 0.f;   // keras 'accuracy' of a sigmoid output
    if (a.prob != nullptr) a.prob[n] = p;
    a.pair[n] = make_float4(term, hit, dz, d);
    s_dz = dz;
    s_dist = d;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    const float diff = es[j] - es[E + j];
    float g;
    if (a.metric == 0)
      g = (s_dist > 0.f) ? s_dz * a.head_w[0] * diff / s_dist : 0.f;   // sqrt'(0) guarded (the reference yields NaN)
    else
      g = s_dz * a.head_w[j] * ((diff > 0.f) - (diff < 0.f));
    de[j] = g;
    a.d_emb[size_t(n) * E + j] = g;
    a.d_emb[size_t(N + n) * E + j] = -g;
  }
  __syncthreads();
  // d_gmax[c] = sum_e d_emb[e] * W[c][e]; four partial chains per channel, summed in a fixed order.  Branch 2 receives
  // the exact negative (d_emb2 = -d_emb1 and rounding is symmetric).
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float4* wr = reinterpret_cast<const float4*>(a.dense_w + size_t(c) * E);
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if ((E & 3) == 0) {
      for (int e = 0; e < E; e += 4) {
        const float4 w4 = __ldg(wr + (e >> 2));
        a4[0] = fmaf(de[e], w4.x, a4[0]);
        a4[1] = fmaf(de[e + 1], w4.y, a4[1]);
        a4[2] = fmaf(de[e + 2], w4.z, a4[2]);
        a4[3] = fmaf(de[e + 3], w4.w, a4[3]);
      }
    } else {
      for (int e = 0; e < E; ++e) a4[e & 3] = fmaf(de[e], __ldg(a.dense_w + size_t(c) * E + e), a4[e & 3]);
    }
    const float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    a.d_gmax[size_t(n) * C + c] = acc;
    a.d_gmax[size_t(N + n) * C + c] = -acc;
  }
}

// blocks [0, ceil(C/8)): dW rows of 8 channels over all 2N clips -- block (E-lane 64, clip quarter 4): a thread sums
// its quarter of the clips in order (four clips in flight), the quarters are added in a fixed order;
// the last block: loss, accuracy and head gradients from the per-pair records, fixed order.
__global__ void __launch_bounds__(256)
siamese_head_finish_kernel(const SiameseHeadArgs a) {
  const int N = a.N, C = a.C, E = a.E, NB = 2 * a.N;
  const int wblocks = (C + 7) / 8;
  const int lane_e = threadIdx.x, part = threadIdx.y;   // blockDim = (64, 4)
  if (int(blockIdx.x) < wblocks) {
    __shared__ float red[3][9][64];   // quarters 1-3: 8 channel sums + the bias sum per E-lane
    const int c0 = blockIdx.x * 8;
    const int nq = (NB + 3) / 4;
    const int n0 = part * nq, n1 = min(NB, n0 + nq);
    for (int e0 = 0; e0 < E; e0 += 64) {
      const int e = e0 + lane_e;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accb = 0.f;
      if (e < E) {
        int n = n0;
        for (; n + 4 <= n1; n += 4) {
          float d[4], xv[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            d[i] = a.d_emb[size_t(n + i) * E + e];
#pragma unroll
            for (int k = 0; k < 8; ++k) xv[i][k] = (c0 + k < C) ? __ldg(a.gmax + size_t(n + i) * C + c0 + k) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            accb += d[i];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv[i][k], d[i], acc[k]);
          }
        }
        for (; n < n1; ++n) {
          const float d = a.d_emb[size_t(n) * E + e];
          accb += d;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (c0 + k < C) acc[k] = fmaf(a.gmax[size_t(n) * C + c0 + k], d, acc[k]);
        }
      }
      if (part > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[part - 1][k][lane_e] = acc[k];
        red[part - 1][8][lane_e] = accb;
      }
      __syncthreads();
      if (part == 0 && e < E) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c0 + k < C)
            a.d_dense_w[size_t(c0 + k) * E + e] = ((acc[k] + red[0][k][lane_e]) + red[1][k][lane_e]) + red[2][k][lane_e];
        if (blockIdx.x == 0) a.d_dense_b[e] = ((accb + red[0][8][lane_e]) + red[1][8][lane_e]) + red[2][8][lane_e];
      }
      __syncthreads();
    }
    return;
  }
  const int t = part * 64 + lane_e;
  if (a.metric == 0) {
    if (t == 0) {
      float gw = 0.f, gb = 0.f;
      for (int n = 0; n < N; ++n) { const float4 r = a.pair[n]; gw = fmaf(r.z, r.w, gw); gb += r.z; }
      a.d_head_w[0] = gw;
      a.d_head_b[0] = gb;
    }
  } else {
    for (int j = t; j < E; j += 256) {
      float gw = 0.f;
      for (int n = 0; n < N; ++n)
        gw = fmaf(a.pair[n].z, fabsf(a.emb[size_t(n) * E + j] - a.emb[size_t(N + n) * E + j]), gw);
      a.d_head_w[j] = gw;
    }
    if (t == 255) {
      float gb = 0.f;
      for (int n = 0; n < N; ++n) gb += a.pair[n].z;
      a.d_head_b[0] = gb;
    }
  }
  if (t == 32) {   // another warp: loss mean (double, pairs in order) and accuracy
    double ls = 0.0;
    float hits = 0.f;
    for (int n = 0; n < N; ++n) { const float4 r = a.pair[n]; ls += double(r.x); hits += r.y; }
    a.loss_acc[0] = float(ls / double(N));
    a.loss_acc[1] = hits / float(N);
  }
}

// ---------------------------------------------------------------------------------------------
// BN backward, pass 1: per-channel sums of dy and dy * xhat over the batch, and the largest |s * dy|.
// The gradient of a pooled output lands on the window's arg-max, where u equals the stored extreme, so
//   sum dy        = sum over windows of dy,      sum dy * xhat = sum over windows of dy * (ext - mean) * rstd
// and this pass reads the two POOLED tensors only.
//   dense  (dy_pooled != null): gradient on the pooled tensor (N, lout, C);
//   sparse (dy_pooled == null): block 4, gradient d_gmax (N, C) sits in window jstar[n][c].
// grid (N, chunks); thread = 4 adjacent channels; partial rows [((n*chunks + chunk)*streams + stream)][C] float2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_reduce_kernel(const float* __restrict__ ext, const float* __restrict__ dy_pooled,
                     const float* __restrict__ d_gmax, const int* __restrict__ jstar, int N, int lout, int C, int G,
                     const float4* __restrict__ bn_const, const float* __restrict__ mask,
                     float2* __restrict__ partial, unsigned int* __restrict__ absmax) {
  const int n = blockIdx.x, chunk = blockIdx.y, chunks = gridDim.y;
  const int groups = C >> 2, streams = ew_streams(C, 4);
  float amax = 0.f;
  if (int(threadIdx.x) < groups * streams) {
    const int cg = threadIdx.x % groups, stream = threadIdx.x / groups, c = 4 * cg;
    const int g = n / (N / G);
    float sabs[4], mean[4], rstd[4], mk[4], s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 bc = bn_const[size_t(g) * C + c + k];
      mk[k] = mask ? mask[size_t(n) * C + c + k] : 1.f;
      sabs[k] = fabsf(bc.x); mean[k] = bc.z; rstd[k] = bc.w;
    }
    if (dy_pooled != nullptr) {
      const int per = (lout + chunks - 1) / chunks;
      const int j0 = chunk * per, j1 = min(lout, j0 + per);
      auto add = [&](const float4& e, const float4& d) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float dy = f4get(d, k) * mk[k];
          s1[k] += dy;
          s2[k] = fmaf(dy, (f4get(e, k) - mean[k]) * rstd[k], s2[k]);
          amax = fmaxf(amax, fabsf(dy) * sabs[k]);
        }
      };
      // four windows in flight per thread; the sums are still taken in window order
      for (int j = j0 + stream; j < j1; j += 4 * streams) {
        float4 e[4], d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int jj = j + i * streams;
          if (jj < j1) {
            const size_t o = (size_t(n) * lout + jj) * C + c;
            e[i] = *reinterpret_cast<const float4*>(ext + o);
            d[i] = __ldcs(reinterpret_cast<const float4*>(dy_pooled + o));
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j + i * streams < j1) add(e[i], d[i]);
      }
    } else if (chunk == 0 && stream == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = jstar[size_t(n) * C + c + k];
        const float dy = d_gmax[size_t(n) * C + c + k] * mk[k];
        s1[k] = dy;
        s2[k] = dy * (ext[(size_t(n) * lout + j) * C + c + k] - mean[k]) * rstd[k];
        amax = fmaxf(amax, fabsf(dy) * sabs[k]);
      }
    }
    float2* row = partial + ((size_t(n) * chunks + chunk) * streams + stream) * C + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) row[k] = make_float2(s1[k], s2[k]);
  }
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(absmax, __float_as_uint(amax));
}

// pass 2: per (group, channel) means -> bwd constants {s, mean_dy, mean_dyxhat, 0}; dgamma/dbeta summed over groups.
struct BnBwdFin {
  static constexpr bool kWarp = false;
  int N, G, L, C;
  const float4* bn_const;
  float4* bwd_const;
  float* dgamma;
  float* dbeta;
  __device__ void operator()(const double2* __restrict__ tmp, int c) const {
    const int clips = N / G;
    const double cnt = double(clips) * double(L);
    double tg = 0.0, tb = 0.0;
    for (int g = 0; g < G; ++g) {
      const double2 sums = rowsum_stage2(tmp, g, C, c);
      const double s1 = sums.x, s2 = sums.y;
      bwd_const[size_t(g) * C + c] = make_float4(bn_const[size_t(g) * C + c].x, float(s1 / cnt), float(s2 / cnt), 0.f);
      tb += s1;
      tg += s2;
    }
    dgamma[c] = float(tg);
    dbeta[c] = float(tb);
  }
};

// pass 3: dU = relu'(u) * s * (dy - mean_dy - xhat * mean_dyxhat) at every un-pooled position (incl. the 'valid'
// tail, which receives no dy but still the batch-statistics terms), written as fp16 planes scaled by the block's
// power-of-two gradient scale (hi, and lo = fp16(residual) when kPlanes == 2); conv-bias gradient partials
// sum_positions dU per channel (un-scaled).  Per channel the expression is affine in u and dy:
//   dU = A*dy + B + Cc*u,  A = s*mk,  B = s*(mean*rstd*mean_dyxhat - mean_dy),  Cc = -s*rstd*mean_dyxhat
// with u and the arg-max flag decoded from the 16-bit activation word.  grid (N, chunks) over pool windows.
template <bool kSparse, int kPlanes, int kPool, int kPer>
__global__ void __launch_bounds__(kEwThreads, kPer == 8 ? 2 : 3)
bn_relu_bwd_kernel(const uint16_t* __restrict__ u16, const float* __restrict__ dy_pooled,
                   const float* __restrict__ d_gmax, const int* __restrict__ jstar, int N, int L, int C, int G,
                   int /*pool == kPool*/, const float4* __restrict__ bn_const, const float4* __restrict__ bwd_const,
                   const float* __restrict__ mask, const unsigned int* __restrict__ absmax,
                   __half* __restrict__ du_hi, __half* __restrict__ du_lo, float* __restrict__ dbias_partial) {
  // kPer adjacent channels per thread: 8 (16-byte rows, 128 registers, 2 blocks per SM) or 4 (8-byte rows, 64
  // registers, 4 blocks per SM: the same bytes in flight per SM from twice the threads, and blocks small enough to
  // share an SM with a weight-gradient CTA of the side stream)
  constexpr int kWords = kPer / 2;
  const int n = blockIdx.x, chunk = blockIdx.y, chunks = gridDim.y;
  const int groups = C / kPer, streams = ew_streams(C, kPer);
  __shared__ float bias_red[kEwThreads * kPer];   // [stream][C] conv-bias partials of the block's threads
  if (int(threadIdx.x) >= groups * streams) return;   // (a barrier waits for the non-exited threads only)
  {
  const int cg = threadIdx.x % groups, stream = threadIdx.x / groups, c = kPer * cg;
  const int g = n / (N / G);
  constexpr int pool = kPool;
  const int lout = L / pool;
  const int wins = (L + pool - 1) / pool;
  const int per = (wins + chunks - 1) / chunks;
  const int w0 = chunk * per, w1 = min(wins, w0 + per);
  const float scale = grad_scale_from_absmax(__uint_as_float(*absmax));
  float A[kPer], B[kPer], Cc[kPer], sb[kPer];
  int js[kSparse ? kPer : 1];
  float dg[kSparse ? kPer : 1];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const float4 bc = bn_const[size_t(g) * C + c + k];
    const float4 bw = bwd_const[size_t(g) * C + c + k];
    const float mk = mask ? mask[size_t(n) * C + c + k] : 1.f;
    A[k] = bw.x * mk * scale;
    B[k] = bw.x * (bc.z * bc.w * bw.z - bw.y) * scale;
    Cc[k] = -bw.x * bc.w * bw.z * scale;
    sb[k] = 0.f;
    if (kSparse) {
      js[k] = jstar[size_t(n) * C + c + k];
      dg[k] = d_gmax[size_t(n) * C + c + k];
    }
  }
  struct Window {
    uint32_t ur[kPool][kWords];  // the window's rows of kPer encoded activations
    float d[kPer];               // its pooled gradient
  };
  auto load_window = [&](int w, Window& W) {
    const int l0 = w * pool, wl = min(pool, L - l0);
    const uint16_t* up = u16 + (size_t(n) * L + l0) * C + c;
#pragma unroll
    for (int i = 0; i < kPool; ++i) {
      if (i < wl) {
        if (kPer == 8) {
          const uint4 v = __ldcs(reinterpret_cast<const uint4*>(up + size_t(i) * C));
          W.ur[i][0] = v.x; W.ur[i][1] = v.y; W.ur[i][kWords - 2] = v.z; W.ur[i][kWords - 1] = v.w;
        } else {
          const uint2 v = __ldcs(reinterpret_cast<const uint2*>(up + size_t(i) * C));
          W.ur[i][0] = v.x; W.ur[i][1] = v.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kWords; ++q) W.ur[i][q] = 0u;
      }
    }
    if (!kSparse && w < lout) {
      const float4* dp = reinterpret_cast<const float4*>(dy_pooled + (size_t(n) * lout + w) * C + c);
#pragma unroll
      for (int h = 0; h < kPer / 4; ++h) {
        const float4 v = __ldcs(dp + h);
        W.d[4 * h] = v.x; W.d[4 * h + 1] = v.y; W.d[4 * h + 2] = v.z; W.d[4 * h + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kPer; ++k) W.d[k] = 0.f;
    }
  };
  auto apply_window = [&](int w, const Window& W) {
    const int l0 = w * pool, wl = min(pool, L - l0);
    float Bf[kPer];   // B + A * dy of this window: what the flagged element receives instead of B
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const float dy = kSparse ? ((w == js[k]) ? dg[k] : 0.f) : W.d[k];
      Bf[k] = fmaf(A[k], dy, B[k]);
    }
    __half* oh = du_hi + (size_t(n) * L + l0) * C + c;
    __half* ol = (kPlanes == 2) ? du_lo + (size_t(n) * L + l0) * C + c : nullptr;
#pragma unroll
    for (int i = 0; i < kPool; ++i) {
      if (i < wl) {
        uint32_t ph[kWords], pl[kWords];
#pragma unroll
        for (int q = 0; q < kWords; ++q) {   // two channels per 32-bit word: paired conversions both ways
          const uint32_t wd = W.ur[i][q];
          const uint32_t mag = wd & 0x7FFF7FFFu;
          const __half2 uh = *reinterpret_cast<const __half2*>(&mag);
          const float2 u = __half22float2(uh);
          const float b0 = (wd & 0x8000u) ? Bf[2 * q] : B[2 * q];
          const float b1 = (int(wd) < 0) ? Bf[2 * q + 1] : B[2 * q + 1];
          float d0 = fmaf(Cc[2 * q], u.x, b0), d1 = fmaf(Cc[2 * q + 1], u.y, b1);
          d0 = (u.x > 0.f) ? d0 : 0.f;      // relu'(u)
          d1 = (u.y > 0.f) ? d1 : 0.f;
          sb[2 * q] += d0;
          sb[2 * q + 1] += d1;
          const __half2 h = __floats2half2_rn(d0, d1);
          ph[q] = *reinterpret_cast<const uint32_t*>(&h);
          if (kPlanes == 2) {
            const float2 hf = __half22float2(h);
            const __half2 lo = __floats2half2_rn(d0 - hf.x, d1 - hf.y);
            pl[q] = *reinterpret_cast<const uint32_t*>(&lo);
          }
        }
        if (kPer == 8) {
          __stcs(reinterpret_cast<uint4*>(oh + size_t(i) * C), make_uint4(ph[0], ph[1], ph[kWords - 2], ph[kWords - 1]));
          if (kPlanes == 2)
            __stcs(reinterpret_cast<uint4*>(ol + size_t(i) * C), make_uint4(pl[0], pl[1], pl[kWords - 2], pl[kWords - 1]));
        } else {
          __stcs(reinterpret_cast<uint2*>(oh + size_t(i) * C), make_uint2(ph[0], ph[1]));
          if (kPlanes == 2) __stcs(reinterpret_cast<uint2*>(ol + size_t(i) * C), make_uint2(pl[0], pl[1]));
        }
      }
    }
  };
  // 8 / kPool windows in flight per thread (the same row count for either pool size): with two windows of MaxPool(2) a
  // thread had 64-96 bytes outstanding and blocks 2-4 ran at 2-4 TB/s
  constexpr int kFlight = 8 / kPool;
  // (requesting the first windows before the constants' loads was measured: no gain for MaxPool(4), and the MaxPool(2)
  // variants spill at the 128-register cap and lose 15 %)
  for (int w = w0 + stream; w < w1; w += kFlight * streams) {
    Window Wf[kFlight];
#pragma unroll
    for (int f = 0; f < kFlight; ++f)
      if (w + f * streams < w1) load_window(w + f * streams, Wf[f]);
#pragma unroll
    for (int f = 0; f < kFlight; ++f)
      if (w + f * streams < w1) apply_window(w + f * streams, Wf[f]);
  }
  const float inv = 1.0f / scale;
#pragma unroll
  for (int k = 0; k < kPer; ++k) bias_red[stream * C + c + k] = sb[k] * inv;
  }
  // one partial row per block: the streams' sums are added in stream order (deterministic)
  __syncthreads();
  float* row = dbias_partial + (size_t(n) * chunks + chunk) * C;
  for (int t = threadIdx.x; t < C; t += groups * streams) {
    float s = 0.f;
    for (int st = 0; st < streams; ++st) s += bias_red[st * C + t];
    row[t] = s;
  }
}

// second stage of a plain column sum -> out[C]
struct ColsumFin {
  static constexpr bool kWarp = false;
  int C;
  float* out;
  __device__ void operator()(const double2* __restrict__ tmp, int c) const { out[c] = float(rowsum_stage2(tmp, 0, C, c).x); }
};

// ---------------------------------------------------------------------------------------------
// Keras Adam with global-norm clipping (keras.optimizers.Adam(clipnorm=...))
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const double v = double(g[i]);
    s += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}
// p -= lr_t * m / (sqrt(v) + eps) with g' = g * inv_scale * min(1, clipnorm / ||g * inv_scale||)
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, size_t n, const double* __restrict__ sumsq, float inv_scale,
                                 float clipnorm, float lr_t, float beta1, float beta2, float eps) {
  float coef = inv_scale;
  if (clipnorm > 0.f) {
    const float norm = float(sqrt(*sumsq)) * inv_scale;
    if (norm > clipnorm) coef *= clipnorm / norm;
  }
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int launch_bn_stats_finalize(const float* partial, int rows_per_clip, int c_pad, int N, int G, int L, int C,
                             const float* gamma, const float* beta, float eps, float momentum, float* moving_mean,
                             float* moving_var, float* bn_const, double* red_scratch, cudaStream_t st) {
  if (N <= 0 || G <= 0 || N % G != 0 || C <= 0) return set_error(VM_ERR_SHAPE, "bn_stats_finalize: bad shape");
  if (red_scratch == nullptr) return set_error(VM_ERR_SHAPE, "bn_stats_finalize: reduction scratch missing");
  double2* tmp = reinterpret_cast<double2*>(red_scratch);
  if (C > 2048) return set_error(VM_ERR_UNSUPPORTED, "bn_stats_finalize: C > 2048");
  const BnStatsFin fin{N, G, L, C, gamma, beta, eps, momentum, moving_mean, moving_var, reinterpret_cast<float4*>(bn_const)};
  rowsum_fused_kernel<2><<<dim3((C + 31) / 32, G * kRB), dim3(32, 8), 0, st>>>(
      partial, size_t(N / G) * rows_per_clip, c_pad, C, tmp, fin);
  return check_launch_t("bn_stats_finalize");
}

int launch_bn_stats_sums(const float* partial, int rows_per_clip, int c_pad, int N, int G, int C, double* red_scratch,
                         double* sums, cudaStream_t st) {
  if (N <= 0 || G <= 0 || N % G != 0 || C <= 0) return set_error(VM_ERR_SHAPE, "bn_stats_sums: bad shape");
  if (red_scratch == nullptr || sums == nullptr) return set_error(VM_ERR_SHAPE, "bn_stats_sums: null buffer");
  double2* tmp = reinterpret_cast<double2*>(red_scratch);
  if (C > 2048) return set_error(VM_ERR_UNSUPPORTED, "bn_stats_sums: C > 2048");
  const SumsFin fin{G, C, reinterpret_cast<double2*>(sums)};
  rowsum_fused_kernel<2><<<dim3((C + 31) / 32, G * kRB), dim3(32, 8), 0, st>>>(
      partial, size_t(N / G) * rows_per_clip, c_pad, C, tmp, fin);
  return check_launch_t("bn_stats_sums");
}

int launch_bn_stats_from_sums(const double* sums, double count, int G, int C, const float* gamma, const float* beta,
                              float eps, float momentum, float* moving_mean, float* moving_var, float* bn_const,
                              cudaStream_t st) {
  if (G <= 0 || C <= 0 || !(count > 1.0)) return set_error(VM_ERR_SHAPE, "bn_stats_from_sums: bad shape");
  bn_stats_from_sums_kernel<<<(C + 63) / 64, 64, 0, st>>>(reinterpret_cast<const double2*>(sums), count, G, C, gamma,
                                                         beta, eps, momentum, moving_mean, moving_var,
                                                         reinterpret_cast<float4*>(bn_const));
  return check_launch_t("bn_stats_from_sums");
}

static int p2p_check(void* const* peers, int rank, int world, int n, P2PPeers* out) {
  if (peers == nullptr || world < 1 || world > kP2PMaxRanks || rank < 0 || rank >= world)
    return set_error(VM_ERR_SHAPE, "p2p: bad rank / world (at most 8 ranks of one node)");
  if (n > kP2PMaxDoubles) return set_error(VM_ERR_UNSUPPORTED, "p2p: vector too long for the exchange buffer");
  for (int r = 0; r < world; ++r) {
    if (peers[r] == nullptr) return set_error(VM_ERR_SHAPE, "p2p: null peer buffer");
    out->buf[r] = peers[r];
  }
  out->rank = rank;
  out->world = world;
  return VM_OK;
}


int launch_bn_stats_finalize_peers(const float* partial, int rows_per_clip, int c_pad, int N, int G, int C,
                                   double* red_scratch, void* const* peers, int rank, int world, unsigned int seq,
                                   double count, const float* gamma, const float* beta, float eps, float momentum,
                                   float* moving_mean, float* moving_var, float* bn_const, double* local_sums,
                                   double* total_sums, cudaStream_t st) {
  if (red_scratch == nullptr) return set_error(VM_ERR_SHAPE, "bn_stats_finalize_peers: reduction scratch missing");
  if (N <= 0 || G <= 0 || G > kPeersMaxGroups || N % G != 0 || C > 2048 || total_sums == nullptr || !(count > 0.0))
    return set_error(VM_ERR_SHAPE, "bn_stats_finalize_peers: bad arguments");
  P2PPeers pp{};
  int rc = p2p_check(peers, rank, world, 2 * G * C, &pp);
  if (rc) return rc;
  const BnStatsPeersFin fin{G, C, count, pp, seq, reinterpret_cast<double2*>(local_sums),
                            reinterpret_cast<double2*>(total_sums), gamma, beta, eps, momentum, moving_mean, moving_var,
                            reinterpret_cast<float4*>(bn_const)};
  rowsum_fused_kernel<2><<<dim3((C + 31) / 32, G * kRB), dim3(32, 8), 0, st>>>(
      partial, size_t(N / G) * rows_per_clip, c_pad, C, reinterpret_cast<double2*>(red_scratch), fin);
  return check_launch_t("bn_stats_finalize_peers");
}

int launch_bn_pool_fwd(const float* ext, int N, int lout, int C, int G, const float* bn_const, const float* mask,
                       __half* out_hi, __half* out_lo, uint16_t* out_q, cudaStream_t st) {
  if (C % 8 != 0 || N <= 0 || G <= 0 || N % G != 0 || lout <= 0) return set_error(VM_ERR_SHAPE, "bn_pool_fwd: bad shape");
  const dim3 grid(N, ew_chunks(N, lout, ew_streams(C, 8)));
  const float4* bc = reinterpret_cast<const float4*>(bn_const);
#define VM_POOL_FWD(LO, Q) \
  bn_pool_fwd_kernel<LO, Q><<<grid, kEwThreads, 0, st>>>(ext, N, lout, C, G, bc, mask, out_hi, out_lo, out_q)
  if (out_lo != nullptr && out_q != nullptr) VM_POOL_FWD(true, true);
  else if (out_lo != nullptr) VM_POOL_FWD(true, false);
  else if (out_q != nullptr) VM_POOL_FWD(false, true);
  else VM_POOL_FWD(false, false);
#undef VM_POOL_FWD
  return check_launch_t("bn_pool_fwd");
}

int launch_bn_gmax_fwd(const float* ext, int N, int lout, int C, int G, const float* bn_const, const float* mask,
                       float* gmax, int* jstar, cudaStream_t st) {
  if (N <= 0 || G <= 0 || N % G != 0 || lout < 1) return set_error(VM_ERR_SHAPE, "bn_gmax_fwd: bad shape");
  bn_gmax_fwd_kernel<<<dim3(N, (C + 31) / 32), dim3(32, 8), 0, st>>>(ext, N, lout, C, G,
                                                                    reinterpret_cast<const float4*>(bn_const), mask,
                                                                    gmax, jstar);
  return check_launch_t("bn_gmax_fwd");
}

int launch_dense_fwd(const float* x, int N, int C, const float* w, const float* b, int E, float* y, cudaStream_t st) {
  if (size_t(C + 256) * 4 > 48 * 1024) return set_error(VM_ERR_UNSUPPORTED, "dense_fwd: C too large");
  dense_fwd_kernel<<<N, 256, (C + 256) * sizeof(float), st>>>(x, N, C, w, b, E, y);
  return check_launch_t("dense_fwd");
}

int launch_dense_bwd(const float* x, const float* dy, const float* w, int N, int C, int E, float* dw, float* db,
                     float* dx, cudaStream_t st) {
  dense_bwd_w_kernel<<<(C + 7) / 8, 64, 0, st>>>(x, dy, N, C, E, dw, db);
  dense_bwd_x_kernel<<<N, 128, E * sizeof(float), st>>>(dy, w, N, C, E, dx);
  return check_launch_t("dense_bwd");
}

int launch_pair_head_loss_bwd(const float* emb, int N, int E, int metric, const float* head_w, const float* head_b,
                              const float* y_true, int loss_kind, float loss_scale, float* d_emb, float* d_head_w,
                              float* d_head_b, float* accuracy, cudaStream_t st) {
  if (N <= 0 || E <= 0 || (metric != 0 && metric != 1) || (loss_kind != 1 && loss_kind != 2))
    return set_error(VM_ERR_SHAPE, "pair_head_loss_bwd: bad arguments");
  if (size_t(N) * 8 > 48 * 1024) return set_error(VM_ERR_UNSUPPORTED, "pair_head_loss_bwd: N too large");
  pair_head_loss_bwd_kernel<<<1, 256, 2 * N * sizeof(float), st>>>(emb, N, E, metric, head_w, head_b, y_true,
                                                                   loss_kind, loss_scale, d_emb, d_head_w, d_head_b,
                                                                   accuracy);
  return check_launch_t("pair_head_loss_bwd");
}

int launch_siamese_head_train(const float* gmax, int N, int C, int E, const float* dense_w, const float* dense_b,
                              int metric, const float* head_w, const float* head_b, const float* y_true, int loss_kind,
                              float loss_scale, float* emb, float* prob, float* d_emb, float* d_gmax, float* pair_scratch,
                              float* d_dense_w, float* d_dense_b, float* d_head_w, float* d_head_b, float* loss_acc,
                              cudaStream_t st) {
  if (N <= 0 || C <= 0 || E <= 0) return set_error(VM_ERR_SHAPE, "siamese_head_train: bad shape");
  if (metric != 0 && metric != 1) return set_error(VM_ERR_UNSUPPORTED, "siamese_head_train: metric not implemented");
  if (loss_kind != 1 && loss_kind != 2) return set_error(VM_ERR_UNSUPPORTED, "siamese_head_train: loss not implemented");
  if (!gmax || !dense_w || !dense_b || !head_w || !head_b || !y_true || !emb || !d_emb || !d_gmax || !pair_scratch ||
      !d_dense_w || !d_dense_b || !d_head_w || !d_head_b || !loss_acc)
    return set_error(VM_ERR_SHAPE, "siamese_head_train: null pointer");
  const size_t smem = (size_t(2) * C + size_t(3) * E + 256) * sizeof(float);
  if (smem > 48 * 1024) return set_error(VM_ERR_UNSUPPORTED, "siamese_head_train: C / E too large");
  const SiameseHeadArgs a{gmax, N, C, E, metric, loss_kind, dense_w, dense_b, head_w, head_b, y_true, loss_scale, emb,
                          prob, d_emb, d_gmax, reinterpret_cast<float4*>(pair_scratch), d_dense_w, d_dense_b, d_head_w,
                          d_head_b, loss_acc};
  siamese_head_pair_kernel<<<N, 256, smem, st>>>(a);
  siamese_head_finish_kernel<<<(C + 7) / 8 + 1, dim3(64, 4), 0, st>>>(a);
  return check_launch_t("siamese_head_train");
}

static int bn_bwd_check(const double* red_scratch, int N, int G, int C, const void* dy_pooled, const void* d_gmax,
                        const void* absmax) {
  if (red_scratch == nullptr) return set_error(VM_ERR_SHAPE, "bn_bwd: reduction scratch missing");
  if (absmax == nullptr) return set_error(VM_ERR_SHAPE, "bn_bwd: gradient-scale word missing");
  if (N <= 0 || G <= 0 || N % G != 0) return set_error(VM_ERR_SHAPE, "bn_bwd: bad shape");
  if ((dy_pooled == nullptr) == (d_gmax == nullptr)) return set_error(VM_ERR_SHAPE, "bn_bwd: give dy_pooled xor d_gmax");
  if (C % 8 != 0) return set_error(VM_ERR_SHAPE, "bn_bwd: C must be a multiple of 8");
  return VM_OK;
}
// scratch contract of the two partial buffers: vm_bn_bwd_scratch_elems(N) entries each (float2 / float)
size_t bn_bwd_scratch_elems(int N) {
  if (N <= 0) return 0;
  return size_t(N) * size_t((kNumSMs * 8 + N - 1) / N) * 2048;
}

// passes 1-2: per-(group, channel) sums of dy and dy*xhat over this rank's clips -> tmp (and, if asked, `sums`)
template <class Fin>
static int bn_bwd_reduce(const float* ext, const float* dy_pooled, const float* d_gmax, const int* jstar, int N,
                         int lout, int C, int G, const float* bn_const, const float* mask, float* partial,
                         unsigned int* absmax, double* red_scratch, const Fin& fin, int presummed_rows,
                         cudaStream_t st) {
  if (C > 2048) return set_error(VM_ERR_UNSUPPORTED, "bn_bwd: C > 2048");
  double2* tmp = reinterpret_cast<double2*>(red_scratch);
  if (presummed_rows > 0) {
    // the dgrad kernel of the block above already left `presummed_rows` partial rows per clip (padded-channel stride)
    // and the gradient-scale word: only the deterministic column reduction + finishing step remain
    const int c_pad = (C + 127) / 128 * 128;
    rowsum_fused_kernel<2><<<dim3((C + 31) / 32, G * kRB), dim3(32, 8), 0, st>>>(
        partial, size_t(N / G) * presummed_rows, c_pad, C, tmp, fin);
    return VM_OK;
  }
  cudaError_t e = cudaMemsetAsync(absmax, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return set_cuda_error(e, "bn_bwd: memset");
  const int streams = ew_streams(C, 4);
  const int chunks = dy_pooled != nullptr ? ew_chunks(N, lout, streams) : 1;
  bn_bwd_reduce_kernel<<<dim3(N, chunks), kEwThreads, 0, st>>>(ext, dy_pooled, d_gmax, jstar, N, lout, C, G,
                                                               reinterpret_cast<const float4*>(bn_const), mask,
                                                               reinterpret_cast<float2*>(partial), absmax);
  rowsum_fused_kernel<2><<<dim3((C + 31) / 32, G * kRB), dim3(32, 8), 0, st>>>(
      partial, size_t(N / G) * chunks * streams, C, C, tmp, fin);
  return VM_OK;
}

// pass 3 + conv-bias gradient, given bwd_const
static int bn_bwd_apply(const uint16_t* u16, const float* dy_pooled, const float* d_gmax, const int* jstar, int N, int L,
                        int C, int G, int pool, const float* bn_const, const float* mask, const float* bwd_const,
                        const unsigned int* absmax, __half* du_hi, __half* du_lo, float* dbias_partial, float* dbias,
                        double* red_scratch, cudaStream_t st) {
  double2* tmp = reinterpret_cast<double2*>(red_scratch);
  // channels per thread: 4 when the block then still has a thread for every channel group (C <= 1024), else 8
  const int per = (C > 1024) ? 8 : 4;
  const int streams = ew_streams(C, per);
  const int chunks = ew_chunks(N, (L + pool - 1) / pool, streams);
  const dim3 grid(N, chunks);
  const float4* bc = reinterpret_cast<const float4*>(bn_const);
  const float4* bw = reinterpret_cast<const float4*>(bwd_const);
  if (pool != 2 && pool != 4) return set_error(VM_ERR_UNSUPPORTED, "bn_bwd: pool must be 2 or 4");
#define VM_RELU_BWD(SPARSE, PLANES, POOL, PER)                                                                          \
  bn_relu_bwd_kernel<SPARSE, PLANES, POOL, PER><<<grid, kEwThreads, 0, st>>>(u16, dy_pooled, d_gmax, jstar, N, L, C, G, \
                                                                             pool, bc, bw, mask, absmax, du_hi, du_lo,  \
                                                                             dbias_partial)
#define VM_RELU_BWD_Q(SPARSE, PLANES, POOL) \
  do { if (per == 8) VM_RELU_BWD(SPARSE, PLANES, POOL, 8); else VM_RELU_BWD(SPARSE, PLANES, POOL, 4); } while (0)
#define VM_RELU_BWD_P(SPARSE, PLANES) \
  do { if (pool == 2) VM_RELU_BWD_Q(SPARSE, PLANES, 2); else VM_RELU_BWD_Q(SPARSE, PLANES, 4); } while (0)
  if (dy_pooled == nullptr) { if (du_lo) VM_RELU_BWD_P(true, 2); else VM_RELU_BWD_P(true, 1); }
  else { if (du_lo) VM_RELU_BWD_P(false, 2); else VM_RELU_BWD_P(false, 1); }
#undef VM_RELU_BWD_P
#undef VM_RELU_BWD_Q
#undef VM_RELU_BWD
  const ColsumFin fin{C, dbias};
  rowsum_fused_kernel<1><<<dim3((C + 31) / 32, kRB), dim3(32, 8), 0, st>>>(dbias_partial, size_t(N) * chunks, C, C, tmp,
                                                                          fin);
  return VM_OK;
}

int launch_bn_bwd(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax, const int* jstar,
                  int N, int L, int C, int G, int pool, const float* bn_const, const float* mask, float* partial,
                  float* bwd_const, float* dgamma, float* dbeta, unsigned int* absmax, __half* du_hi, __half* du_lo,
                  float* dbias_partial, float* dbias, double* red_scratch, int presummed_rows, cudaStream_t st) {
  int rc = bn_bwd_check(red_scratch, N, G, C, dy_pooled, d_gmax, absmax);
  if (rc) return rc;
  const BnBwdFin fin{N, G, L, C, reinterpret_cast<const float4*>(bn_const), reinterpret_cast<float4*>(bwd_const),
                     dgamma, dbeta};
  if ((rc = bn_bwd_reduce(ext, dy_pooled, d_gmax, jstar, N, L / pool, C, G, bn_const, mask, partial, absmax,
                          red_scratch, fin, presummed_rows, st)))
    return rc;
  bn_bwd_apply(u16, dy_pooled, d_gmax, jstar, N, L, C, G, pool, bn_const, mask, bwd_const, absmax, du_hi, du_lo,
               dbias_partial, dbias, red_scratch, st);
  return check_launch_t("bn_bwd");
}

int launch_bn_bwd_sums(const float* ext, const float* dy_pooled, const float* d_gmax, const int* jstar, int N, int L,
                       int C, int G, int pool, const float* bn_const, const float* mask, float* partial,
                       unsigned int* absmax, double* red_scratch, double* sums, int presummed_rows, cudaStream_t st) {
  int rc = bn_bwd_check(red_scratch, N, G, C, dy_pooled, d_gmax, absmax);
  if (rc) return rc;
  if (sums == nullptr) return set_error(VM_ERR_SHAPE, "bn_bwd_sums: null sums");
  const SumsFin fin{G, C, reinterpret_cast<double2*>(sums)};
  if ((rc = bn_bwd_reduce(ext, dy_pooled, d_gmax, jstar, N, L / pool, C, G, bn_const, mask, partial, absmax,
                          red_scratch, fin, presummed_rows, st)))
    return rc;
  return check_launch_t("bn_bwd_sums");
}

int launch_bn_bwd_from_sums(const double* local_sums, const double* global_sums, double count, const uint16_t* u16,
                            const float* dy_pooled, const float* d_gmax, const int* jstar, int N, int L, int C, int G,
                            int pool, const float* bn_const, const float* mask, float* bwd_const, float* dgamma,
                            float* dbeta, const unsigned int* absmax, __half* du_hi, __half* du_lo,
                            float* dbias_partial, float* dbias, double* red_scratch, cudaStream_t st) {
  int rc = bn_bwd_check(red_scratch, N, G, C, dy_pooled, d_gmax, absmax);
  if (rc) return rc;
  if (local_sums == nullptr || global_sums == nullptr || !(count > 0.0))
    return set_error(VM_ERR_SHAPE, "bn_bwd_from_sums: bad sums / count");
  bn_bwd_from_sums_kernel<<<(C + 63) / 64, 64, 0, st>>>(reinterpret_cast<const double2*>(local_sums),
                                                       reinterpret_cast<const double2*>(global_sums), count, G, C,
                                                       reinterpret_cast<const float4*>(bn_const),
                                                       reinterpret_cast<float4*>(bwd_const), dgamma, dbeta);
  bn_bwd_apply(u16, dy_pooled, d_gmax, jstar, N, L, C, G, pool, bn_const, mask, bwd_const, absmax, du_hi, du_lo,
               dbias_partial, dbias, red_scratch, st);
  return check_launch_t("bn_bwd_from_sums");
}


int launch_bn_bwd_peers(const uint16_t* u16, const float* ext, const float* dy_pooled, const float* d_gmax,
                        const int* jstar, int N, int L, int C, int G, int pool, const float* bn_const, const float* mask,
                        float* partial, float* bwd_const, float* dgamma, float* dbeta, unsigned int* absmax,
                        __half* du_hi, __half* du_lo, float* dbias_partial, float* dbias, double* red_scratch,
                        int presummed_rows, void* const* peers, int rank, int world, unsigned int seq, double count,
                        double* local_sums, double* total_sums, cudaStream_t st) {
  int rc = bn_bwd_check(red_scratch, N, G, C, dy_pooled, d_gmax, absmax);
  if (rc) return rc;
  if (local_sums == nullptr || total_sums == nullptr || !(count > 0.0) || G > kPeersMaxGroups)
    return set_error(VM_ERR_SHAPE, "bn_bwd_peers: bad sums / count / groups");
  P2PPeers pp{};
  if ((rc = p2p_check(peers, rank, world, 2 * G * C, &pp))) return rc;
  const BnBwdPeersFin fin{G, C, count, pp, seq, reinterpret_cast<double2*>(local_sums),
                          reinterpret_cast<double2*>(total_sums), reinterpret_cast<const float4*>(bn_const),
                          reinterpret_cast<float4*>(bwd_const), dgamma, dbeta};
  if ((rc = bn_bwd_reduce(ext, dy_pooled, d_gmax, jstar, N, L / pool, C, G, bn_const, mask, partial, absmax,
                          red_scratch, fin, presummed_rows, st)))
    return rc;
  bn_bwd_apply(u16, dy_pooled, d_gmax, jstar, N, L, C, G, pool, bn_const, mask, bwd_const, absmax, du_hi, du_lo,
               dbias_partial, dbias, red_scratch, st);
  return check_launch_t("bn_bwd_peers");
}

int launch_adam_step(float* p, const float* g, float* m, float* v, size_t n, double* sumsq_scratch, float inv_scale,
                     float clipnorm, float lr_t, float beta1, float beta2, float eps, cudaStream_t st) {
  if (n == 0) return VM_OK;
  cudaError_t e = cudaMemsetAsync(sumsq_scratch, 0, sizeof(double), st);
  if (e != cudaSuccess) return set_cuda_error(e, "adam: memset");
  const unsigned blocks = unsigned(min(size_t(148 * 4), (n + 255) / 256));
  sumsq_kernel<<<blocks, 256, 0, st>>>(g, n, sumsq_scratch);
  adam_step_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, sumsq_scratch, inv_scale, clipnorm, lr_t, beta1, beta2, eps);
  return check_launch_t("adam_step");
}

}  // namespace vm
