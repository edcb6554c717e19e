// Small CUDA-core kernels around the two conv kernels:
//   * weight packing (Keras layout fp32 -> fp16 (hi, lo) planes in the layouts the MMA kernels consume, BN folded
//     into per-channel epilogue constants)                                   voicemap/models.py:13-35
//   * fp32 <-> (hi, lo) plane conversion for the per-block C-ABI and the tests
//   * GlobalMaxPool1D finalisation + Dense(embedding_dimension)               voicemap/models.py:37-39
//   * siamese head (pairwise L2 / |.|, Dense(1), sigmoid) + contrastive / BCE loss
//                                                   voicemap/models.py:55-69, voicemap/utils.py:77-85
#include "vm_common.cuh"
#include "vm_kernels.h"

namespace vm {

// Epilogue constants {a, c, t, s}.  With acc' = sigma * acc (weights are packed sigma-scaled, sigma = sign(s)) and
// M = max over the pool window of acc':   y = s * relu(sigma * M + bias) + t
//                                           = s >= 0 ? max(a * M + c, t) : min(a * M + c, t),  a = s * sigma, c = s * bias + t
// (one FFMA + one FMNMX per pooled value; s only selects min/max).
__device__ __forceinline__ float bn_scale(const float* gamma, const float* var, float eps, int c) {
  return gamma ? gamma[c] * (1.0f / sqrtf(var[c] + eps)) : 1.0f;  // null BN = identity (train-mode "raw" packing)
}
__device__ __forceinline__ float4 fold_bn(float bias, float gamma, float beta, float mean, float var, float eps) {
  const float s = gamma * (1.0f / sqrtf(var + eps));
  const float t = beta - mean * s;
  const float c = fmaf(s, bias, t);
  return (s < 0.f) ? make_float4(-s, c, -INFINITY, bias) : make_float4(s, c, -bias, INFINITY);  // see apply_epi
}

// y = relu(acc + bias) = max(acc, -bias) + bias (bit-identical: both branches round once)
__device__ __forceinline__ float4 epi_relu_only(float bias) { return make_float4(1.f, bias, -bias, INFINITY); }

// ---------------------------------------------------------------------------------------------
// conv3 weights: w (3, cin, cout) fp32 -> wpack [plane][tap][cout_pad][cin], 16 bits per entry; planes: fp16 hi, fp16 lo,
// e5m2x2 Q (split_w_q: the operand of the single-MMA correction product of precision 2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_conv3_body(size_t idx, const float* __restrict__ w, const float* __restrict__ bias,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ var,
                                                float eps, int cin, int cout, int cout_pad,
                                                __half* __restrict__ wpack, float4* __restrict__ epi) {
  const size_t total = size_t(3) * cout_pad * cin;
  if (idx < size_t(cout_pad)) {
    const int co = int(idx);
    if (co >= cout) epi[co] = make_float4(0.f, 0.f, 0.f, 0.f);
    else if (gamma == nullptr) epi[co] = epi_relu_only(bias ? bias[co] : 0.f);
    else epi[co] = fold_bn(bias[co], gamma[co], beta[co], mean[co], var[co], eps);
  }
  if (idx >= total) return;
  const int ci = int(idx % cin);
  const int co = int((idx / cin) % cout_pad);
  const int tap = int(idx / (size_t(cin) * cout_pad));
  float v = 0.f;
  if (co < cout) {
    const float s = bn_scale(gamma, var, eps, co);
    v = w[(size_t(tap) * cin + ci) * cout + co];
    if (s < 0.f) v = -v;
  }
  __half h, l;
  split_f32(v, h, l);
  wpack[idx] = h;
  wpack[total + idx] = l;
  uint16_t q;
  split_w_q(v, h, q);
  wpack[2 * total + idx] = __ushort_as_half(q);
}
__global__ void pack_conv3_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ var, float eps, int cin,
                                  int cout, int cout_pad, __half* __restrict__ wpack, float4* __restrict__ epi) {
  pack_conv3_body(size_t(blockIdx.x) * blockDim.x + threadIdx.x, w, bias, gamma, beta, mean, var, eps, cin, cout,
                  cout_pad, wpack, epi);
}

// ---------------------------------------------------------------------------------------------
// conv1 weights: w (32, 1, cout) fp32 -> [slab][plane] 8 KB smem images (no-swizzle K-major core matrices:
// byte offset of (row, tap) = (row/8)*512 + (tap/8)*128 + (row%8)*16 + (tap%8)*2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_conv1_body(int idx, const float* __restrict__ w, const float* __restrict__ bias,
                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ var,
                                                float eps, int cout, int cout_pad, __half* __restrict__ wpack,
                                                float4* __restrict__ epi) {
  if (idx < cout_pad) {
    if (idx >= cout) epi[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    else if (gamma == nullptr) epi[idx] = epi_relu_only(bias ? bias[idx] : 0.f);
    else epi[idx] = fold_bn(bias[idx], gamma[idx], beta[idx], mean[idx], var[idx], eps);
  }
  if (idx >= cout_pad * 32) return;
  const int tap = idx & 31;
  const int co = idx >> 5;
  float v = 0.f;
  if (co < cout) {
    const float s = bn_scale(gamma, var, eps, co);
    v = w[size_t(tap) * cout + co];
    if (s < 0.f) v = -v;
  }
  __half h, l;
  split_f32(v, h, l);
  const int slab = co >> 7, row = co & 127;
  const int off = (row >> 3) * 256 + (tap >> 3) * 64 + (row & 7) * 8 + (tap & 7);  // in halves
  __half* img = wpack + size_t(slab) * 8192;                                        // 2 planes x 4096 halves
  img[off] = h;
  img[4096 + off] = l;
}
__global__ void pack_conv1_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ var, float eps, int cout,
                                  int cout_pad, __half* __restrict__ wpack, float4* __restrict__ epi) {
  pack_conv1_body(blockIdx.x * blockDim.x + threadIdx.x, w, bias, gamma, beta, mean, var, eps, cout, cout_pad, wpack,
                  epi);
}

// ---------------------------------------------------------------------------------------------
// dgrad weights: dX[p][ci] = sum_t sum_co dU[p + t - 1][co] * W[2 - t][ci][co] is a k=3 'same' convolution of dU with
// the tap-flipped, channel-transposed kernel.  w (3, cin, cout) -> wpack [plane][tap][cin_pad][cout] fp16, i.e. the
// conv3 operand layout with the roles (cin' = cout, cout' = cin), as fp16 (hi, lo) planes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_conv3_dgrad_body(size_t idx, const float* __restrict__ w, int cin, int cout,
                                                      int cin_pad, __half* __restrict__ wpack,
                                                      float4* __restrict__ epi) {
  const size_t total = size_t(3) * cin_pad * cout;
  if (idx < size_t(cin_pad)) epi[idx] = make_float4(0.f, 0.f, 0.f, 0.f);  // unused (linear epilogue)
  if (idx >= total) return;
  const int co = int(idx % cout);
  const int ci = int((idx / cout) % cin_pad);
  const int tap = int(idx / (size_t(cout) * cin_pad));
  const float v = (ci < cin) ? w[(size_t(2 - tap) * cin + ci) * cout + co] : 0.f;
  __half h, l;    // fp16 (hi, lo) planes, like the forward weights; the other operand is the scaled fp16 gradient
  split_f32(v, h, l);
  wpack[idx] = h;
  wpack[total + idx] = l;
}
__global__ void pack_conv3_dgrad_kernel(const float* __restrict__ w, int cin, int cout, int cin_pad,
                                        __half* __restrict__ wpack, float4* __restrict__ epi) {
  pack_conv3_dgrad_body(size_t(blockIdx.x) * blockDim.x + threadIdx.x, w, cin, cout, cin_pad, wpack, epi);
}

// All operand packing of a training step in ONE launch (weights change once per step): block ranges select the task.
__global__ void pack_train_kernel(const PackTrainArgs a) {
  int t = 0;
#pragma unroll
  for (int i = 1; i < 7; ++i)
    if (i < a.ntasks && blockIdx.x >= a.t[i].block0) t = i;
  const PackTrainTask& k = a.t[t];
  const size_t idx = size_t(blockIdx.x - k.block0) * blockDim.x + threadIdx.x;
  if (k.kind == 0) {
    pack_conv1_body(int(idx), k.w, k.bias, nullptr, nullptr, nullptr, nullptr, 0.f, k.cout, k.pad,
                    static_cast<__half*>(k.wpack), reinterpret_cast<float4*>(k.epi));
  } else if (k.kind == 1) {
    pack_conv3_body(idx, k.w, k.bias, nullptr, nullptr, nullptr, nullptr, 0.f, k.cin, k.cout, k.pad,
                    static_cast<__half*>(k.wpack), reinterpret_cast<float4*>(k.epi));
  } else {
    pack_conv3_dgrad_body(idx, k.w, k.cin, k.cout, k.pad, static_cast<__half*>(k.wpack),
                          reinterpret_cast<float4*>(k.epi));
  }
}

// ---------------------------------------------------------------------------------------------
// plane conversion
// ---------------------------------------------------------------------------------------------
__global__ void split_planes_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ hi,
                                    __half* __restrict__ lo) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    __half h, l;
    split_f32(x[i], h, l);
    hi[i] = h;
    if (lo != nullptr) lo[i] = l;
  }
}
// precision-2 planes: (fp16 hi, e5m2x2 Q) -- see vm_common.cuh
__global__ void split_planes_q_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ hi,
                                      uint16_t* __restrict__ q) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    __half h;
    uint16_t qq;
    split_f16_q(x[i], h, qq);
    hi[i] = h;
    q[i] = qq;
  }
}
__global__ void merge_planes_q_kernel(const __half* __restrict__ hi, const uint16_t* __restrict__ q, size_t n,
                                      float* __restrict__ x) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    x[i] = __half2float(hi[i]) + e5m2_to_float(q[i] & 0xFFu) * kQDown;
}
__global__ void merge_planes_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, size_t n,
                                    float* __restrict__ x) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    x[i] = __half2float(hi[i]) + (lo != nullptr ? __half2float(lo[i]) : 0.f);
}

// ---------------------------------------------------------------------------------------------
// GlobalMaxPool finalisation + Dense.  partial: (N, T, c_pad) raw accumulator maxima per position tile.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gmax_dense_kernel(const float* __restrict__ partial, int T, int C, int c_pad, const float4* __restrict__ epi,
                  const float* __restrict__ dense_w, const float* __restrict__ dense_b, int E,
                  float* __restrict__ gmax_out, float* __restrict__ emb) {
  extern __shared__ float g[];      // [C] pooled activations, then [4][E] partial dot products
  float* red = g + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, partial[(size_t(n) * T + t) * c_pad + c]);
    const float y = apply_epi(epi[c], m);
    g[c] = y;
    if (gmax_out != nullptr) gmax_out[size_t(n) * C + c] = y;
  }
  __syncthreads();
  if (emb == nullptr) return;
  // 4 channel quarters x 64 outputs per pass: each thread accumulates a quarter of the dot product
  const int part = threadIdx.x >> 6, lane_e = threadIdx.x & 63;
  const int cq = (C + 3) / 4;
  const int c0 = part * cq, c1 = min(C, c0 + cq);
  for (int e0 = 0; e0 < E; e0 += 64) {
    const int e = e0 + lane_e;
    float acc = 0.f;
    if (e < E) {
      // latency bound on the L2-resident weight loads: keep 32 of them in flight (four accumulators break the FMA chain)
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      int c = c0;
      for (; c + 32 <= c1; c += 32) {
        float w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = __ldg(dense_w + size_t(c + i) * E + e);
#pragma unroll
        for (int i = 0; i < 32; ++i) a4[i & 3] = fmaf(g[c + i], w[i], a4[i & 3]);
      }
      for (; c < c1; ++c) a4[0] = fmaf(g[c], __ldg(dense_w + size_t(c) * E + e), a4[0]);
      acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    }
    red[part * 64 + lane_e] = acc;
    __syncthreads();
    if (part == 0 && e < E)
      emb[size_t(n) * E + e] = ((red[lane_e] + red[64 + lane_e]) + (red[128 + lane_e] + red[192 + lane_e])) + dense_b[e];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// siamese head + loss.  metric 0: uniform_euclidean (head_w, head_b scalars), 1: weighted_l1 (head_w[E]).
// loss_kind 0: none, 1: contrastive (margin 1), 2: binary cross-entropy (keras clip 1e-7).  One block.
// ---------------------------------------------------------------------------------------------
__global__ void pair_head_loss_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int N, int E,
                                      int metric, const float* __restrict__ head_w, const float* __restrict__ head_b,
                                      const float* __restrict__ y_true, int loss_kind, float* __restrict__ dist,
                                      float* __restrict__ prob, float* __restrict__ loss) {
  __shared__ float red[32];
  float local = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* a = e1 + size_t(n) * E;
    const float* b = e2 + size_t(n) * E;
    float z;
    if (metric == 0) {
      float ss = 0.f;
      for (int j = 0; j < E; ++j) {
        const float d = a[j] - b[j];
        ss = fmaf(d, d, ss);
      }
      const float d = sqrtf(fmaxf(ss, 0.f));
      if (dist != nullptr) dist[n] = d;
      z = fmaf(d, head_w[0], head_b[0]);
    } else {
      float acc = 0.f;
      for (int j = 0; j < E; ++j) acc = fmaf(fabsf(a[j] - b[j]), head_w[j], acc);
      z = acc + head_b[0];
    }
    const float pr = 1.0f / (1.0f + expf(-z));
    if (prob != nullptr) prob[n] = pr;
    if (loss_kind == 1) {
      const float y = y_true[n];
      const float mg = fmaxf(1.0f - pr, 0.f);
      local += (1.f - y) * pr * pr + y * mg * mg;
    } else if (loss_kind == 2) {
      const float y = y_true[n];
      const float pc = fminf(fmaxf(pr, 1e-7f), 1.0f - 1e-7f);
      local += -y * logf(pc) - (1.f - y) * logf(1.0f - pc);
    }
  }
  if (loss_kind == 0 || loss == nullptr) return;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) loss[0] = v / float(N);
  }
}

// ---------------------------------------------------------------------------------------------
// k-way n-shot scoring on embeddings (voicemap/utils.py:156-212): class means over the n shots of each of the k
// support classes, distance to the query (0 euclidean, 1 cosine, 2 dot product, as the reference defines them), and
// the arg-min class (first minimum, like np.argmin).  One block per task, one warp per class (round robin); sums are
// taken in double like scipy's cdist / the float64 numpy of the reference, so near-ties resolve the same way.
//   cosine:      centre = mean_i s_i / |s_i|;                 score = 1 - centre.q / (|centre| |q|)
//   dot product: centre = (mean_i s_i / |s_i|) * mean_i |s_i|; score = -centre.q
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void nshot_score_kernel(const float* __restrict__ query, const float* __restrict__ support, int T, int k,
                                   int n, int E, int distance, float* __restrict__ scores, int* __restrict__ best) {
  extern __shared__ double sc[];   // [k] scores of this task
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* q = query + size_t(t) * E;
  for (int c = warp; c < k; c += nwarps) {
    const float* s = support + (size_t(t) * k + c) * size_t(n) * E;
    double score;
    if (distance == 0) {
      double ss = 0.0;
      for (int j = lane; j < E; j += 32) {
        double m = 0.0;
        for (int i = 0; i < n; ++i) m += double(s[size_t(i) * E + j]);
        const double d = m / double(n) - double(q[j]);
        ss += d * d;
      }
      score = sqrt(warp_sum(ss));
    } else {
      // unit vectors of the shots: 1 / |s_i| per shot first (n is small: 1 or 5 in the reference's scripts)
      double cq = 0.0, cc = 0.0, qq = 0.0, mean_norm = 0.0;
      for (int j0 = 0; j0 < E; j0 += 32) {   // centre_j needs every shot's norm: recomputed per j block (n*E is tiny)
        const int j = j0 + lane;
        double cj = 0.0;
        for (int i = 0; i < n; ++i) {
          double nn = 0.0;
          for (int jj = lane; jj < E; jj += 32) { const double v = double(s[size_t(i) * E + jj]); nn += v * v; }
          const double norm = sqrt(warp_sum(nn));
          if (j < E) cj += double(s[size_t(i) * E + j]) / norm;
          if (j0 == 0) mean_norm += norm;
        }
        cj /= double(n);
        if (j < E) { const double qj = double(q[j]); cq += cj * qj; cc += cj * cj; qq += qj * qj; }
      }
      cq = warp_sum(cq); cc = warp_sum(cc); qq = warp_sum(qq);
      mean_norm /= double(n);
      score = (distance == 1) ? 1.0 - cq / (sqrt(cc) * sqrt(qq)) : -(cq * mean_norm);
    }
    if (lane == 0) {
      sc[c] = score;
      if (scores != nullptr) scores[size_t(t) * k + c] = float(score);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = 0;
    for (int c = 1; c < k; ++c)
      if (sc[c] < sc[b]) b = c;
    best[t] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// Preprocessing statistics (voicemap/utils.py:22-34, 88-101): decimate x[:, ::stride], then per-clip mean and ONE
// scale rms / sqrt(mean(batch^2)) per whiten() call (= per group of clips; the mean of squares is taken over the
// un-centred decimated batch).  Accumulates in double like the reference's float64 numpy.
// ---------------------------------------------------------------------------------------------
__global__ void preprocess_clip_sums_kernel(const float* __restrict__ x, int T, int stride, int L,
                                            double* __restrict__ sums /* (N, 2) */) {
  __shared__ double r1[32], r2[32];
  const int n = blockIdx.x;
  const float* xc = x + size_t(n) * T;
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const double v = double(xc[size_t(i) * stride]);
    s1 += v;
    s2 += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) { a += r1[w]; b += r2[w]; }
    sums[2 * n] = a;
    sums[2 * n + 1] = b;
  }
}
__global__ void preprocess_finalize_kernel(const double* __restrict__ sums, int N, int G, int L, float rms,
                                           float* __restrict__ mean, float* __restrict__ scale) {
  const int clips = N / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double sq = 0.0;
    for (int n = g * clips; n < (g + 1) * clips; ++n) sq += sums[2 * n + 1];
    const float sc = float(double(rms) / sqrt(sq / (double(clips) * double(L))));
    for (int n = g * clips; n < (g + 1) * clips; ++n) {
      mean[n] = float(sums[2 * n] / double(L));
      scale[n] = sc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return VM_OK;
}

int launch_pack_conv3(const float* w, const float* bias, const float* gamma, const float* beta, const float* mean,
                      const float* var, float eps, int cin, int cout, void* wpack, float* epi, cudaStream_t stream) {
  if (cin <= 0 || cout <= 0) return set_error(VM_ERR_SHAPE, "pack_conv3: bad shape");
  const int cout_pad = (cout + 127) / 128 * 128;
  const size_t total = size_t(3) * cout_pad * cin;
  const int threads = 256;
  const unsigned blocks = unsigned((total + threads - 1) / threads);
  pack_conv3_kernel<<<blocks, threads, 0, stream>>>(w, bias, gamma, beta, mean, var, eps, cin, cout, cout_pad,
                                                   static_cast<__half*>(wpack), reinterpret_cast<float4*>(epi));
  return check_launch("pack_conv3");
}

int launch_pack_conv3_dgrad(const float* w, int cin, int cout, void* wpack, float* epi, cudaStream_t stream) {
  if (cin <= 0 || cout <= 0) return set_error(VM_ERR_SHAPE, "pack_conv3_dgrad: bad shape");
  const int cin_pad = (cin + 127) / 128 * 128;
  const size_t total = size_t(3) * cin_pad * cout;
  pack_conv3_dgrad_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(w, cin, cout, cin_pad,
                                                                            static_cast<__half*>(wpack),
                                                                            reinterpret_cast<float4*>(epi));
  return check_launch("pack_conv3_dgrad");
}

int launch_pack_conv1(const float* w, const float* bias, const float* gamma, const float* beta, const float* mean,
                      const float* var, float eps, int cout, void* wpack, float* epi, cudaStream_t stream) {
  if (cout <= 0) return set_error(VM_ERR_SHAPE, "pack_conv1: bad shape");
  const int cout_pad = (cout + 127) / 128 * 128;
  const int total = cout_pad * 32;
  pack_conv1_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, bias, gamma, beta, mean, var, eps, cout, cout_pad,
                                                            static_cast<__half*>(wpack),
                                                            reinterpret_cast<float4*>(epi));
  return check_launch("pack_conv1");
}

int launch_preprocess_stats(const float* x, int N, int T, int stride, int G, float rms, float* mean, float* scale,
                            cudaStream_t stream) {
  if (N <= 0 || T <= 0 || stride <= 0 || G <= 0 || N % G != 0) return set_error(VM_ERR_SHAPE, "preprocess: bad shape");
  // scratch: the (N, 2) double sums live in the tail of the `scale` allocation contract: caller provides
  // mean (N floats) and scale (N floats + 4*N floats of scratch)
  double* sums = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(scale + N) + 7) & ~uintptr_t(7));
  const int L = (T + stride - 1) / stride;
  preprocess_clip_sums_kernel<<<N, 256, 0, stream>>>(x, T, stride, L, sums);
  preprocess_finalize_kernel<<<1, 128, 0, stream>>>(sums, N, G, L, rms, mean, scale);
  return check_launch("preprocess_stats");
}

int launch_split_planes(const float* x, size_t n, __half* hi, __half* lo, cudaStream_t stream) {
  if (n == 0) return VM_OK;
  const unsigned blocks = unsigned(min(size_t(148 * 16), (n + 255) / 256));
  split_planes_kernel<<<blocks, 256, 0, stream>>>(x, n, hi, lo);
  return check_launch("split_planes");
}

int launch_split_planes_q(const float* x, size_t n, __half* hi, uint16_t* q, cudaStream_t stream) {
  if (n == 0) return VM_OK;
  const unsigned blocks = unsigned(min(size_t(148 * 16), (n + 255) / 256));
  split_planes_q_kernel<<<blocks, 256, 0, stream>>>(x, n, hi, q);
  return check_launch("split_planes_q");
}
int launch_merge_planes_q(const __half* hi, const uint16_t* q, size_t n, float* x, cudaStream_t stream) {
  if (n == 0) return VM_OK;
  const unsigned blocks = unsigned(min(size_t(148 * 16), (n + 255) / 256));
  merge_planes_q_kernel<<<blocks, 256, 0, stream>>>(hi, q, n, x);
  return check_launch("merge_planes_q");
}
int launch_merge_planes(const __half* hi, const __half* lo, size_t n, float* x, cudaStream_t stream) {
  if (n == 0) return VM_OK;
  const unsigned blocks = unsigned(min(size_t(148 * 16), (n + 255) / 256));
  merge_planes_kernel<<<blocks, 256, 0, stream>>>(hi, lo, n, x);
  return check_launch("merge_planes");
}

int launch_gmax_dense(const float* partial, int N, int T, int C, int c_pad, const float* epi, const float* dense_w,
                      const float* dense_b, int E, float* gmax_out, float* emb, cudaStream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || c_pad < C) return set_error(VM_ERR_SHAPE, "gmax_dense: bad shape");
  if (emb != nullptr && (E <= 0 || dense_w == nullptr || dense_b == nullptr))
    return set_error(VM_ERR_SHAPE, "gmax_dense: dense weights missing");
  if (size_t(C + 256) * sizeof(float) > 48 * 1024) return set_error(VM_ERR_UNSUPPORTED, "gmax_dense: C too large");
  gmax_dense_kernel<<<N, 256, (C + 256) * sizeof(float), stream>>>(partial, T, C, c_pad,
                                                                  reinterpret_cast<const float4*>(epi), dense_w, dense_b,
                                                                  E, gmax_out, emb);
  return check_launch("gmax_dense");
}

int launch_pair_head_loss(const float* e1, const float* e2, int N, int E, int metric, const float* head_w,
                          const float* head_b, const float* y_true, int loss_kind, float* dist, float* prob,
                          float* loss, cudaStream_t stream) {
  if (N <= 0 || E <= 0) return set_error(VM_ERR_SHAPE, "pair_head_loss: bad shape");
  if (metric != 0 && metric != 1) return set_error(VM_ERR_UNSUPPORTED, "pair_head_loss: metric not implemented");
  if (loss_kind != 0 && y_true == nullptr) return set_error(VM_ERR_SHAPE, "pair_head_loss: y_true required");
  pair_head_loss_kernel<<<1, 256, 0, stream>>>(e1, e2, N, E, metric, head_w, head_b, y_true, loss_kind, dist, prob,
                                               loss);
  return check_launch("pair_head_loss");
}

int launch_nshot_score(const float* query, const float* support, int T, int k, int n, int E, int distance,
                       float* scores, int* best, cudaStream_t stream) {
  if (T <= 0 || k <= 0 || n <= 0 || E <= 0) return set_error(VM_ERR_SHAPE, "nshot_score: bad shape");
  if (distance < 0 || distance > 2) return set_error(VM_ERR_UNSUPPORTED, "nshot_score: distance must be 0, 1 or 2");
  if (size_t(k) * sizeof(double) > 48 * 1024) return set_error(VM_ERR_UNSUPPORTED, "nshot_score: k too large");
  if (query == nullptr || support == nullptr || best == nullptr) return set_error(VM_ERR_SHAPE, "nshot_score: null pointer");
  nshot_score_kernel<<<T, 128, k * sizeof(double), stream>>>(query, support, T, k, n, E, distance, scores, best);
  return check_launch("nshot_score");
}

int launch_pack_train(const float* const* kernels, const float* const* biases, int filters, void* const* wraw,
                      float* const* eraw, void* const* wdg, float* const* edg, cudaStream_t stream) {
  if (filters <= 0 || kernels == nullptr || biases == nullptr || wraw == nullptr || eraw == nullptr || wdg == nullptr ||
      edg == nullptr)
    return set_error(VM_ERR_SHAPE, "pack_train: bad arguments");
  PackTrainArgs a{};
  unsigned blocks = 0;
  int cin = 1;
  for (int b = 0; b < 4; ++b) {
    const int cout = filters * (b + 1);
    const int cout_pad = (cout + 127) / 128 * 128;
    PackTrainTask& f = a.t[a.ntasks++];
    f.w = kernels[b]; f.bias = biases[b]; f.wpack = wraw[b]; f.epi = eraw[b];
    f.cin = cin; f.cout = cout; f.pad = cout_pad; f.kind = (b == 0) ? 0 : 1; f.block0 = blocks;
    const size_t total = (b == 0) ? size_t(cout_pad) * 32 : size_t(3) * cout_pad * cin;
    blocks += unsigned((total + 255) / 256);
    if (b > 0) {
      PackTrainTask& d = a.t[a.ntasks++];
      const int cin_pad = (cin + 127) / 128 * 128;
      d.w = kernels[b]; d.bias = nullptr; d.wpack = wdg[b]; d.epi = edg[b];
      d.cin = cin; d.cout = cout; d.pad = cin_pad; d.kind = 2; d.block0 = blocks;
      blocks += unsigned((size_t(3) * cin_pad * cout + 255) / 256);
    }
    cin = cout;
  }
  pack_train_kernel<<<blocks, 256, 0, stream>>>(a);
  return check_launch("pack_train");
}

}  // namespace vm
