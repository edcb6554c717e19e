// Block 1 of the voicemap encoder: Conv1D(filters, 32, 'same') + bias + ReLU + BatchNorm(eval) + MaxPool1D(4,4)
// on the raw waveform (Cin = 1).  Reference: voicemap/models.py:13-19 (zero pad 15 left / 16 right).
//
// HBM-bound layer (AI ~62 F/B): the 32-tap filter bank is run on tcgen05 tensor cores so that arithmetic is
// free and the kernel streams at the output-write rate.  D[cout, position] = W[cout, tap] * T[position, tap]
// where T is the Toeplitz (im2col) matrix of the waveform: T[p, k] = x[p + k - 15].
//   * producer warps split the waveform into fp16 (hi, lo) and write the Toeplitz operand in the no-swizzle K-major
//     UMMA layout with OVERLAPPING core matrices: the 8x8 core matrix of (position group g, tap chunk c) holds
//     x[8(g+c) + i + j] (row i, tap j), i.e. it depends on g + c only, so with LBO = SBO one 128-byte block
//     (at a 144-byte pitch) per value of g + c serves every (g, c) pair: 36 blocks (4.5 KB) per plane and tile instead of the
//     32 x 4 blocks of a materialised im2col tile -- 8x duplication of the waveform instead of 32x;
//   * one thread issues 6 MMAs per tile (2 K-steps x {Th*Wh, Tl*Wh, Th*Wl}) into a double-buffered TMEM
//     accumulator (128 cout lanes x 256 positions);
//   * 16 epilogue warps (two groups of 8, one per accumulator buffer) pool the raw accumulators 4:1, apply
//     bias/ReLU/BN, split to fp16 (hi, lo) planes, stage them in shared memory and store channels-last with TMA.
//     As in vm_conv3.cu the packed weights carry sigma = sign(BN scale) so that the max-pool commutes with the
//     affine.
#include <type_traits>

#include "vm_common.cuh"
#include "vm_kernels.h"

namespace vm {

namespace c1 {
constexpr int kTileN = 256;
constexpr int kTileM = 128;
constexpr int kGroupStride = 528;                        // materialised tile (wgrad1): 4 k-chunks x 128 B + 16 B pad
constexpr int kPlaneBytes = (kTileN / 8) * kGroupStride;  // 16896 (wgrad1)
constexpr int kBlocks = kTileN / 8 + 4;                  // overlapped layout: g + c = 0 .. 35
constexpr int kBlockStride = 144;                        // 128 B + 16 B pad: the lanes' 16-byte stores hit distinct banks
constexpr int kOvPlaneBytes = kBlocks * kBlockStride;    // 5184
constexpr int kStageBytes = 2 * kOvPlaneBytes;           // hi + lo
constexpr int kMaxStages = 4;
constexpr int kOutBoxBytes = 64 * 128;                   // TMA store box: 64 pooled positions x 64 channels fp16
constexpr int kOutBufBytes = 4 * kOutBoxBytes;           // [plane][channel half][pos][64 ch] = 32 KB
constexpr int kWPlaneBytes = kTileM * 64;                // 128 cout x 32 taps fp16 = 8192
constexpr int kWSlabBytes = 2 * kWPlaneBytes;
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 16;                            // two groups of 8 (one per TMEM accumulator buffer)
constexpr int kProdWarps = 3;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;  // warps 0-15 epilogue, 16 MMA, 17-19 producers
constexpr int kMaxSlabs = 4;
__host__ __device__ constexpr int smem_bytes(int nslab, int nstages) {
  return nslab * kWSlabBytes + nstages * kStageBytes + 2 * kOutBufBytes + 1024 + 256;
}
}  // namespace c1

struct __align__(8) Conv1Barriers {
  uint64_t full[c1::kMaxStages], empty[c1::kMaxStages];
  uint64_t tfull[2], tempty[2];
  uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------------
// Toeplitz operand producer (one warp = one 256-position tile; lane = 8 consecutive positions).
// Loads the lane's 39-sample strip x[p0 - 15 + 8*lane + k] (strided + whitened when preprocessing is fused), splits it
// into 16-bit (hi, lo) planes -- fp16 for the forward conv, bf16 for the weight gradient -- and writes the lane's 8
// rows x 4 tap-chunks as 16-byte shared-memory stores into the no-swizzle UMMA core-matrix layout
// (row r, chunk j) -> (r/8)*kGroupStride + j*128 + (r%8)*16.
// ---------------------------------------------------------------------------------------------
template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (kBf16) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  if (kBf16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

__device__ __forceinline__ void toeplitz_load_strip(const float* __restrict__ xc, int xs, float pm, float ps, int e0,
                                                    int L, float (&xv)[39]) {
  if (e0 >= 0 && e0 + 39 <= L) {  // interior strip: no bounds predicates
#pragma unroll
    for (int k = 0; k < 39; ++k) xv[k] = (__ldg(xc + size_t(e0 + k) * xs) - pm) * ps;
  } else {
#pragma unroll
    for (int k = 0; k < 39; ++k) {
      const int e = e0 + k;
      xv[k] = (e >= 0 && e < L) ? (__ldg(xc + size_t(e) * xs) - pm) * ps : 0.f;
    }
  }
}

template <bool kBf16>
__device__ __forceinline__ void toeplitz_store_plane(const float (&v)[39], uint32_t base) {
  // pe[k] = (h[2k], h[2k+1]), po[k] = (h[2k+1], h[2k+2]): even- and odd-aligned pairs of the 16-bit plane
  uint32_t pe[19], po[19];
#pragma unroll
  for (int k = 0; k < 19; ++k) {
    pe[k] = pack2<kBf16>(v[2 * k], v[2 * k + 1]);
    po[k] = pack2<kBf16>(v[2 * k + 1], v[2 * k + 2]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = r + 8 * j;  // first element of this 8-tap chunk
      const uint32_t* src = (o & 1) ? (po + (o - 1) / 2) : (pe + o / 2);
      sts_v4(base + j * 128 + r * 16, src[0], src[1], src[2], src[3]);
    }
  }
}
template <bool kBf16>
__device__ __forceinline__ float round16(float x) {
  if (kBf16) return __bfloat162float(__float2bfloat16_rn(x));
  return __half2float(__float2half_rn(x));
}
// hi plane first, then the residuals in place for the lo plane (keeps the register peak at one plane's worth)
template <bool kBf16>
__device__ __forceinline__ void toeplitz_store(float (&xv)[39], uint32_t st, int nplanes, int plane_bytes) {
  toeplitz_store_plane<kBf16>(xv, st);
  if (nplanes == 2) {
#pragma unroll
    for (int k = 0; k < 39; ++k) xv[k] -= round16<kBf16>(xv[k]);
    toeplitz_store_plane<kBf16>(xv, st + plane_bytes);
  }
}

// ---------------------------------------------------------------------------------------------
// Overlapped Toeplitz producer for the forward conv (see the file header).  Block m of a tile holds the eight
// 16-byte rows  row i = h[8m + i .. 8m + i + 7],  h[q] = fp16 plane of the preprocessed sample x[p0 - 15 + q].
// One lane builds one block from 15 samples: an aligned 16-float window when the clip is contiguous and the window
// is interior, guarded scalar loads otherwise (strided / decimated input, clip borders -> 'same' zero padding).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void toeplitz_block_load(const float* __restrict__ xc, int xs, float pm, float ps, int e0,
                                                    int L, float (&v)[15]) {
  // needs x[e0 .. e0 + 14]; e0 = p0 - 15 + 8m, so e0 - 1 is a multiple of 8
  const float* w = xc + (e0 - 1);
  if (xs == 1 && e0 >= 1 && e0 + 15 <= L && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4 a = __ldg(w4), b = __ldg(w4 + 1), c = __ldg(w4 + 2), d = __ldg(w4 + 3);
    const float t[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
    for (int k = 0; k < 15; ++k) v[k] = (t[k + 1] - pm) * ps;
  } else {
#pragma unroll
    for (int k = 0; k < 15; ++k) {
      const int e = e0 + k;
      v[k] = (e >= 0 && e < L) ? (__ldg(xc + size_t(e) * xs) - pm) * ps : 0.f;
    }
  }
}
__device__ __forceinline__ void toeplitz_block_store_plane(const float (&v)[15], uint32_t dst) {
  uint32_t pe[7], po[7];  // pe[k] = (h[2k], h[2k+1]), po[k] = (h[2k+1], h[2k+2])
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    pe[k] = pack2<false>(v[2 * k], v[2 * k + 1]);
    po[k] = pack2<false>(v[2 * k + 1], v[2 * k + 2]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t* src = (i & 1) ? (po + (i - 1) / 2) : (pe + i / 2);
    sts_v4(dst + i * 16, src[0], src[1], src[2], src[3]);
  }
}
__device__ __forceinline__ void toeplitz_block_store(float (&v)[15], uint32_t dst, int nplanes) {
  toeplitz_block_store_plane(v, dst);
  if (nplanes == 2) {
#pragma unroll
    for (int k = 0; k < 15; ++k) v[k] -= round16<false>(v[k]);
    toeplitz_block_store_plane(v, dst + c1::kOvPlaneBytes);
  }
}

// 32 accumulator columns of one channel -> 32 / kPool pooled outputs (MaxPool kPool) -> bias/ReLU/BN -> fp16 (hi, lo)
// -> staging rows sh/sl + j * 128.  Both clamp forms are evaluated and one is selected: measured faster than
// branching on no_hi inside the pipelined loop (profiles/r01_conv1_epilogue_ab.log: 0.097 vs 0.111 ms per launch).
// kSecond: second output plane -- 0 none, 1 fp16 residual, 2 e5m2x2 Q pair (precision 2, vm_common.cuh).
template <int kPool, int kSecond>
__device__ __forceinline__ void pool_epilogue(const float4& ep, bool no_hi, const uint32_t (&r)[32], uint32_t sh,
                                              uint32_t sl) {
#pragma unroll
  for (int j = 0; j < 32 / kPool; ++j) {
    float y;
    if (kPool == 4) {
      const float v0 = __uint_as_float(r[4 * j]), v1 = __uint_as_float(r[4 * j + 1]);
      const float v2 = __uint_as_float(r[4 * j + 2]), v3 = __uint_as_float(r[4 * j + 3]);
      y = no_hi ? apply_epi_pool4<false>(ep, v0, v1, v2, v3) : apply_epi_pool4<true>(ep, v0, v1, v2, v3);
    } else {
      const float v0 = __uint_as_float(r[2 * j]), v1 = __uint_as_float(r[2 * j + 1]);
      y = no_hi ? apply_epi_pool2<false>(ep, v0, v1) : apply_epi_pool2<true>(ep, v0, v1);
    }
    __half h;
    if (kSecond == 2) {
      uint16_t q;
      split_f16_q(y, h, q);
      sts_b16(sl + j * 128, q);
    } else {
      __half l;
      split_f32(y, h, l);
      if (kSecond == 1) sts_u16(sl + j * 128, l);
    }
    sts_u16(sh + j * 128, h);
  }
}

template <int kPool>
__global__ void __launch_bounds__(c1::kThreads, 1)
conv1_kernel(const __grid_constant__ CUtensorMap tm_oh, const __grid_constant__ CUtensorMap tm_ol,
             const Conv1Params p) {
  using namespace c1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* stages = smem + p.nslab * kWSlabBytes;
  const int kStages = p.nstages;
  uint8_t* outbuf = stages + kStages * kStageBytes;
  Conv1Barriers* bars = reinterpret_cast<Conv1Barriers*>(outbuf + 2 * kOutBufBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = p.N * p.nptile;
  const int nplanes = (p.products >= 2) ? 2 : 1;       // block 1 itself always runs fp16 x 3 for products >= 2
  const int second = (p.products == 2) ? 2 : nplanes - 1;   // format of the second output plane

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 256); }
    fence_mbar_init();
  }
  // packed weights -> shared memory (already in the UMMA smem image layout)
  {
    const uint4* src = p.wpack;
    const uint32_t dst = smem_u32(wsm);
    const int n16 = p.nslab * kWSlabBytes / 16;
    for (int i = threadIdx.x; i < n16; i += kThreads) {
      const uint4 w = __ldg(src + i);
      sts_v4(dst + i * 16, w.x, w.y, w.z, w.w);
    }
    fence_proxy_async_smem();
  }
  if (warp == kEpiWarps) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp > kEpiWarps) {
    // ===================== Toeplitz producers: warp q builds local tiles q, q + P, ...; tile i uses stage i % S ===
    const int q = warp - kEpiWarps - 1;
    for (uint32_t i = q; int(blockIdx.x + i * gridDim.x) < ntiles; i += kProdWarps) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int sidx = i % kStages;
      const int n = tile / p.nptile;
      const int p0 = (tile % p.nptile) * kTileN;
      const float* xc = p.x + size_t(n) * size_t(p.x_clip_stride);
      const int xs = p.x_stride;
      // fused preprocessing (voicemap/utils.py:22-34,88-101): decimation = strided read, whitening = per-clip
      // affine; 'same' zero padding applies to the preprocessed signal
      const float pm = p.pre_mean ? __ldg(p.pre_mean + n) : 0.f;
      const float ps = p.pre_scale ? __ldg(p.pre_scale + n) : 1.f;
      // lane builds block m = lane (and lanes 0-3 block 32 + lane): samples x[p0 - 15 + 8m + (0..14)]
      float xv[15], xw[15];
      toeplitz_block_load(xc, xs, pm, ps, p0 - 15 + 8 * lane, p.L, xv);
      if (lane < kBlocks - 32) toeplitz_block_load(xc, xs, pm, ps, p0 - 15 + 8 * (32 + lane), p.L, xw);
      mbar_wait(&bars->empty[sidx], (((i / kStages) & 1) ^ 1));
      const uint32_t sb = smem_u32(stages + sidx * kStageBytes);
      toeplitz_block_store(xv, sb + lane * kBlockStride, nplanes);
      if (lane < kBlocks - 32) toeplitz_block_store(xw, sb + (32 + lane) * kBlockStride, nplanes);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->full[sidx]);
    }
  } else if (warp == kEpiWarps) {
    // ===================== MMA issuer: whole warp runs the loop, one elected lane issues (see elect_one_sync) ====
    {
      const bool elected = elect_one_sync();
      const uint32_t idesc = make_idesc_f16(kTileM, kTileN);
      uint32_t i = 0, ait = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int s = i % kStages;
        mbar_wait(&bars->full[s], (i / kStages) & 1);
        tc_fence_after_sync();
        const uint32_t th = smem_u32(stages + s * kStageBytes);
        const uint32_t tl = th + kOvPlaneBytes;
        for (int slab = 0; slab < p.nslab; ++slab, ++ait) {
          const int buf = ait & 1;
          mbar_wait(&bars->tempty[buf], ((ait >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + buf * kTileN;
          const uint32_t wh = smem_u32(wsm + slab * kWSlabBytes);
          const uint32_t wl = wh + kWPlaneBytes;
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (elected) umma_f16(d_tmem, make_smem_desc(wh + k * 256, 128, 512, kLayoutNone),
                     make_smem_desc(th + k * 2 * kBlockStride, kBlockStride, kBlockStride, kLayoutNone), idesc, k);
          if (nplanes == 2) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (elected) umma_f16(d_tmem, make_smem_desc(wh + k * 256, 128, 512, kLayoutNone),
                       make_smem_desc(tl + k * 2 * kBlockStride, kBlockStride, kBlockStride, kLayoutNone), idesc, 1);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (elected) umma_f16(d_tmem, make_smem_desc(wl + k * 256, 128, 512, kLayoutNone),
                       make_smem_desc(th + k * 2 * kBlockStride, kBlockStride, kBlockStride, kLayoutNone), idesc, 1);
          }
          if (elected) umma_commit(&bars->tfull[buf]);
        }
        if (elected) umma_commit(&bars->empty[s]);
      }
    }
  } else {
    // ===================== epilogue (warps 0-15) =====================
    // Two independent groups of 8 warps: group g owns TMEM buffer g (every second accumulator), one 32 KB staging
    // buffer, its own named barrier and its own TMA-store leader, so the epilogues of consecutive tiles overlap.
    // Inside a group: thread = cout channel (TMEM lane), the two warps of a lane quarter split the 256 columns.
    // Outputs are staged in shared memory and written with TMA bulk tensor stores (positions / channels out of
    // range are clipped by the TMA unit).
    const int grp = warp >> 3;
    const int wg = warp & 7;
    const int q = wg & 3;            // TMEM lane quarter
    const int chalf = wg >> 2;       // which half of the 256 position columns
    const bool leader = (wg == 0 && lane == 0);
    const uint32_t bar_id = 1 + grp;
    const int ch = q * 32 + lane;
    uint8_t* ob = outbuf + grp * kOutBufBytes;
    const int buf = grp;
    uint32_t t_gcount = 0;   // train-mode forward: granules staged so far by this group (selects the staging half)
    // this group's accumulators are ait = grp, grp + 2, ...; (tile, slab, n, p0) advance incrementally
    const int dn = int(gridDim.x) / p.nptile, dpt = int(gridDim.x) % p.nptile;
    int tile = blockIdx.x, n = tile / p.nptile, pt = tile % p.nptile, slab = grp;
    const auto next_tile = [&]() {
      tile += gridDim.x; n += dn; pt += dpt;
      if (pt >= p.nptile) { pt -= p.nptile; ++n; }
    };
    while (slab >= p.nslab && tile < ntiles) { slab -= p.nslab; next_tile(); }
    for (uint32_t it = 0; tile < ntiles; ++it) {
      const int p0 = pt * kTileN;
      {
        const int co = slab * kTileM + ch;
        const float4 ep = p.epi[co];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * kTileN;
        if (p.out_u16 != nullptr) {
          // train-mode forward: per position u = relu(acc + bias) -> {sum, sum of squares} partials for the batch
          // statistics, the encoded un-pooled activation (fp16 + arg-max flag, encode_u) and the fp32 extreme of every
          // MaxPool(kPool) window (max, or min where the BatchNorm scale is negative).  Granule = 32 positions x 128
          // channels; the group's staging buffer is used as two halves of 16 KB, each
          // [u16: 2 boxes of 32 pos x 64 ch][extremes: 4 boxes of 32/kPool windows x 32 ch fp32], so that the TMA
          // store of one granule drains while the next one is computed: ONE named barrier per granule (before it the
          // leader waits for the store issued a granule earlier, i.e. for the half that is written next).  The two
          // warps of a lane quarter take 16 columns each; TMEM loads are software-pipelined.
          if (wg == 0) mbar_wait(&bars->tfull[buf], it & 1);  // one polling warp per group; the rest block in bar.sync
          named_bar_sync(bar_id, 256);
          tc_fence_after_sync();
          constexpr int kWin = 16 / kPool;                     // windows per warp and granule
          constexpr int kExtBox = (32 / kPool) * 128;          // bytes of one extremes box
          const bool neg = (p.sign_src != nullptr && co < p.cout) ? (p.sign_src[co] < 0.f) : false;
          const uint32_t sgn = neg ? 0x80000000u : 0u;         // compare -u where the minimum is wanted (u >= 0)
          const int lvalid = p.lout * kPool;
          const bool interior = (p0 + kTileN <= p.L) && (p0 + kTileN <= lvalid);
          float s1 = 0.f, s2 = 0.f;
          const uint32_t off_u = (ch >> 6) * 4096 + (chalf * 16) * 128 + (ch & 63) * 2;
          const uint32_t off_m = 8192 + (ch >> 5) * kExtBox + (chalf * kWin) * 128 + (ch & 31) * 4;
          auto granule = [&](const uint32_t (&r)[16], int gr, auto interior_tag) {
            constexpr bool kInterior = decltype(interior_tag)::value;
            const uint32_t st = smem_u32(ob) + (t_gcount & 1) * 16384;
            const int pos0 = p0 + gr * 32 + chalf * 16;
#pragma unroll
            for (int w = 0; w < kWin; ++w) {
              float y[kPool];
#pragma unroll
              for (int i = 0; i < kPool; ++i) {
                y[i] = apply_epi(ep, __uint_as_float(r[kPool * w + i]));
                if (kInterior || pos0 + kPool * w + i < p.L) { s1 += y[i]; s2 = fmaf(y[i], y[i], s2); }
              }
              // winner of the window, first on ties: a chain of strict comparisons on the (sign-adjusted) values
              bool later[kPool];
              float best = __uint_as_float(__float_as_uint(y[0]) ^ sgn), ext = y[0];
#pragma unroll
              for (int i = 1; i < kPool; ++i) {
                const float key = __uint_as_float(__float_as_uint(y[i]) ^ sgn);
                later[i] = key > best;
                best = later[i] ? key : best;
                ext = later[i] ? y[i] : ext;
              }
              const bool win = kInterior || (pos0 + kPool * w + kPool - 1 < lvalid);   // 'valid' pooling drops the tail
#pragma unroll
              for (int i = 0; i < kPool; ++i) {
                bool flag = (i == 0) ? true : later[i];
#pragma unroll
                for (int k = i + 1; k < kPool; ++k) flag = flag && !later[k];
                sts_b16(st + off_u + (kPool * w + i) * 128, encode_u(y[i], flag && win));
              }
              sts_f32(st + off_m + w * 128, ext);
            }
            fence_proxy_async_smem();
            if (leader) tma_store_wait_read<0>();   // the store of the previous granule (other half) has drained
            named_bar_sync(bar_id, 256);
            if (leader) {
              const uint8_t* sb = ob + (t_gcount & 1) * 16384;
              const int pos = p0 + gr * 32;
              if (pos < p.L) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const int c0 = slab * kTileM + half * 64;
                  if (c0 < p.cout) tma_store_3d(&tm_oh, sb + half * 4096, c0, pos, n);
                }
                if (pos / kPool < p.lout) {
#pragma unroll
                  for (int b4 = 0; b4 < 4; ++b4) {
                    const int c0 = slab * kTileM + b4 * 32;
                    if (c0 < p.cout) tma_store_3d(&tm_ol, sb + 8192 + b4 * kExtBox, c0, pos / kPool, n);
                  }
                }
              }
              tma_store_commit();
            }
            ++t_gcount;
          };
          auto tile_loop = [&](auto interior_tag) {
            uint32_t ra[16], rb[16];
            const uint32_t tcol = taddr + chalf * 16;
            tmem_ld_32x16_issue(tcol, ra);
#pragma unroll 1
            for (int gr = 0; gr < kTileN / 32; gr += 2) {
              tmem_ld_wait(ra);
              tmem_ld_32x16_issue(tcol + (gr + 1) * 32, rb);
              granule(ra, gr, interior_tag);
              tmem_ld_wait(rb);
              if (gr + 2 < kTileN / 32) {
                tmem_ld_32x16_issue(tcol + (gr + 2) * 32, ra);
              } else {  // all TMEM reads of this accumulator are done
                tc_fence_before_sync();
                mbar_arrive(&bars->tempty[buf]);
              }
              granule(rb, gr + 1, interior_tag);
            }
          };
          if (interior) tile_loop(std::true_type{});
          else tile_loop(std::false_type{});
          if (p.stat_partial != nullptr)
            p.stat_partial[(size_t(tile) * 2 + chalf) * p.cout_pad + co] = make_float2(s1, s2);
        } else {
          const uint32_t st_h = smem_u32(ob) + (ch >> 6) * kOutBoxBytes + (ch & 63) * 2;
          const uint32_t st_l = st_h + 2 * kOutBoxBytes;
          const bool no_hi = epi_no_upper_clamp(ep);
          // The staging buffer holds 64 pooled positions: MaxPool 4 stages the whole tile in one pass, MaxPool 2
          // (first pool of the older reference architecture) takes two passes of 128 columns.
          constexpr int kPasses = 4 / kPool;
          constexpr int kLoads = (kTileN / 64) / kPasses;   // 32-column TMEM loads per warp and pass
          constexpr int kOutPerLoad = 32 / kPool;
          // the second plane's format is fixed per launch: instantiate the pipelined loop per format instead of
          // branching inside it (a branch in this loop costs ~10 %, profiles/r01_conv1_epilogue_ab.log)
          auto passes = [&](auto second_tag) {
            constexpr int kSecond = decltype(second_tag)::value;
#pragma unroll 1
            for (int pass = 0; pass < kPasses; ++pass) {
              // warp 0 of the group waits for (a) this group's previous TMA store to have drained the staging buffer
              // and (b) the accumulator; the other seven warps block in the named barrier instead of spinning
              if (wg == 0) {
                if (lane == 0) tma_store_wait_read<0>();
                if (pass == 0) mbar_wait(&bars->tfull[buf], it & 1);
              }
              named_bar_sync(bar_id, 256);
              tc_fence_after_sync();
              // software-pipelined TMEM reads: the next 32 columns are in flight while these are processed
              const int g0 = pass * 2 * kLoads + chalf * kLoads;   // first 32-column group of this warp
              const uint32_t row0 = chalf * kLoads * kOutPerLoad * 128;
              uint32_t ra[32], rb[32];
              tmem_ld_32x32_issue(taddr + g0 * 32, ra);
#pragma unroll
              for (int gg = 0; gg < kLoads; gg += 2) {
                tmem_ld_wait(ra);
                tmem_ld_32x32_issue(taddr + (g0 + gg + 1) * 32, rb);
                pool_epilogue<kPool, kSecond>(ep, no_hi, ra, st_h + row0 + gg * kOutPerLoad * 128,
                                              st_l + row0 + gg * kOutPerLoad * 128);
                tmem_ld_wait(rb);
                if (gg + 2 < kLoads) {
                  tmem_ld_32x32_issue(taddr + (g0 + gg + 2) * 32, ra);
                } else if (pass == kPasses - 1) {  // all TMEM reads of this accumulator are done
                  tc_fence_before_sync();
                  mbar_arrive(&bars->tempty[buf]);
                }
                pool_epilogue<kPool, kSecond>(ep, no_hi, rb, st_h + row0 + (gg + 1) * kOutPerLoad * 128,
                                              st_l + row0 + (gg + 1) * kOutPerLoad * 128);
              }
              fence_proxy_async_smem();
              named_bar_sync(bar_id, 256);
              if (leader) {
                const int pos = p0 / kPool + pass * 64;
                if (pos < p.lout) {
#pragma unroll
                  for (int half = 0; half < 2; ++half) {
                    const int c0 = slab * kTileM + half * 64;
                    if (c0 < p.cout) {
                      tma_store_3d(&tm_oh, ob + half * kOutBoxBytes, c0, pos, n);
                      if (nplanes == 2) tma_store_3d(&tm_ol, ob + (2 + half) * kOutBoxBytes, c0, pos, n);
                    }
                  }
                }
                tma_store_commit();
              }
            }
          };
          if (second == 2) passes(std::integral_constant<int, 2>{});
          else if (second == 1) passes(std::integral_constant<int, 1>{});
          else passes(std::integral_constant<int, 0>{});
        }
      }
      slab += 2;
      while (slab >= p.nslab && tile < ntiles) { slab -= p.nslab; next_tile(); }
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int launch_conv1(const float* x, int N, int L, int cout, const void* wpack, const float* epi, __half* out_hi,
                 __half* out_lo, uint16_t* out_u16, float* out_ext, const float* sign_src, float* stat_partial,
                 int products, int max_ctas, cudaStream_t stream, int x_stride, long long x_clip_stride,
                 const float* pre_mean, const float* pre_scale, int pool) {
  using namespace c1;
  if (pool != 2 && pool != 4) return set_error(VM_ERR_UNSUPPORTED, "conv1: first MaxPool1D size must be 2 or 4");
  if (N <= 0 || L < pool) return set_error(VM_ERR_SHAPE, "conv1: need N > 0 and L >= pool size");
  if (cout <= 0 || cout % 8 != 0) return set_error(VM_ERR_UNSUPPORTED, "conv1: Cout must be a positive multiple of 8");
  if (products < 1 || products > 3) return set_error(VM_ERR_SHAPE, "conv1: products must be 1, 2 or 3");
  if ((out_u16 != nullptr) != (out_ext != nullptr))
    return set_error(VM_ERR_SHAPE, "conv1: the train-mode forward needs both out_u16 and out_ext");
  if (products == 2 && out_u16 != nullptr) products = 3;   // block 1 itself always runs fp16 x 3 for products >= 2
  const int cout_pad = (cout + kTileM - 1) / kTileM * kTileM;
  const int nslab = cout_pad / kTileM;
  if (nslab > kMaxSlabs) return set_error(VM_ERR_UNSUPPORTED, "conv1: Cout > 512 not supported");
  Conv1Params p{};
  p.x = x; p.N = N; p.L = L; p.cout = cout;
  p.x_stride = x_stride > 0 ? x_stride : 1;
  p.x_clip_stride = x_clip_stride > 0 ? x_clip_stride : (long long)L * p.x_stride;
  p.pre_mean = pre_mean; p.pre_scale = pre_scale; p.cout_pad = cout_pad; p.nslab = nslab;
  p.lout = L / pool;
  p.nptile = (L + kTileN - 1) / kTileN;
  p.products = products;
  p.wpack = reinterpret_cast<const uint4*>(wpack);
  p.epi = reinterpret_cast<const float4*>(epi);
  p.out_hi = out_hi; p.out_lo = out_lo;
  p.nstages = kMaxStages;
  p.out_u16 = out_u16; p.out_ext = out_ext; p.sign_src = sign_src;
  p.stat_partial = reinterpret_cast<float2*>(stat_partial);
  CUtensorMap oh, ol;
  if (out_u16 != nullptr) {
    const uint64_t udims[3] = {uint64_t(cout), uint64_t(L), uint64_t(N)};
    const uint64_t ustr[2] = {uint64_t(cout) * 2, uint64_t(L) * cout * 2};
    const uint32_t ubox[3] = {64, 32, 1};
    const uint64_t mdims[3] = {uint64_t(cout), uint64_t(p.lout), uint64_t(N)};
    const uint64_t mstr[2] = {uint64_t(cout) * 4, uint64_t(p.lout) * cout * 4};
    const uint32_t mbox[3] = {32, uint32_t(32 / pool), 1};
    int rc;
    if ((rc = make_tensor_map(&oh, out_u16, 3, udims, ustr, ubox, VM_SWIZZLE_NONE))) return rc;
    if ((rc = make_tensor_map(&ol, out_ext, 3, mdims, mstr, mbox, VM_SWIZZLE_NONE, /*f32=*/1))) return rc;
  } else {
    if (out_hi == nullptr) return set_error(VM_ERR_SHAPE, "conv1: no output given");
    if (products >= 2 && out_lo == nullptr) return set_error(VM_ERR_SHAPE, "conv1: out_lo required for products>=2");
    const uint64_t odims[3] = {uint64_t(cout), uint64_t(p.lout), uint64_t(N)};
    const uint64_t ostr[2] = {uint64_t(cout) * 2, uint64_t(p.lout) * cout * 2};
    const uint32_t obox[3] = {64, 64, 1};
    int rc;
    if ((rc = make_tensor_map(&oh, out_hi, 3, odims, ostr, obox, VM_SWIZZLE_NONE))) return rc;
    if ((rc = make_tensor_map(&ol, products >= 2 ? out_lo : out_hi, 3, odims, ostr, obox, VM_SWIZZLE_NONE)))
      return rc;
  }
  const int smem = smem_bytes(nslab, p.nstages);
  const auto kern = (pool == 2) ? conv1_kernel<2> : conv1_kernel<4>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return set_cuda_error(e, "conv1: cudaFuncSetAttribute");
  const int ntiles = N * p.nptile;
  int grid = max_ctas > 0 ? max_ctas : num_sms();
  if (grid > ntiles) grid = ntiles;
  kern<<<grid, kThreads, smem, stream>>>(oh, ol, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "conv1: launch");
  return VM_OK;
}


// =============================================================================================================
// Block-1 weight gradient on tensor cores:  dW1[k][co] = sum_{n,p} x[n][p + k - 15] * dU1[n][p][co].
//   D[co (128 lanes), tap (32 columns)] += dU^T[co, p] * T[p, tap] with the POSITION axis as the MMA K dimension:
//   A = dU1 tile [64 positions][128 co] as TMA writes it (SWIZZLE_128B), read MN-major (two 64-channel atoms);
//   B = the same Toeplitz tile the forward producer builds (bf16 planes here), read MN-major: 8-tap chunks 128 B
//       apart (SBO), 8-position row groups kGroupStride apart (LBO).
// One CTA accumulates over all its tiles and writes a [32][cout] partial; wgrad_reduce sums the CTAs.
// HBM-bound: the 2 x 2-byte dU1 planes are read once (algorithmic bytes 4 * N * L * cout).
// =============================================================================================================
namespace w1 {
constexpr int kTileN = 256;
constexpr int kSub = 64;
constexpr int kGroupStride = c1::kGroupStride;
constexpr int kTPlaneBytes = c1::kPlaneBytes;      // 16896
constexpr int kTStageBytes = 2 * kTPlaneBytes;
constexpr int kTStages = 2;
constexpr int kUHalfBytes = kSub * 128;            // 64 positions x 64 channels
constexpr int kUPlaneBytes = 2 * kUHalfBytes;      // 128 channels
constexpr int kUStageBytes = 2 * kUPlaneBytes;     // hi + lo
constexpr int kUStages = 4;                        // two-plane stages; a one-plane gradient cuts the ring into 8
constexpr int kUMaxStages = 8;
constexpr int kThreads = 192;                      // warps 0-1 Toeplitz producers, 0-3 final epilogue, 4 TMA, 5 MMA
constexpr int kSmemBytes = kTStages * kTStageBytes + kUStages * kUStageBytes + 1024 + 256;
}  // namespace w1

struct __align__(8) Wgrad1Barriers {
  uint64_t tfull[w1::kTStages], tempty[w1::kTStages];
  uint64_t ufull[w1::kUMaxStages], uempty[w1::kUMaxStages];
  uint64_t done;
  uint32_t tmem_base;
};

struct Wgrad1Params {
  const float* x;
  int N, L, cout, nptile, products;
  float* partial;  // [gridDim.x][32][cout]
};

__global__ void __launch_bounds__(w1::kThreads, 1)
wgrad1_tc_kernel(const __grid_constant__ CUtensorMap tm_uh, const __grid_constant__ CUtensorMap tm_ul,
                 const Wgrad1Params p) {
  using namespace w1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* uring = smem;                                  // SWIZZLE_128B tiles first (1024-byte aligned)
  uint8_t* tring = smem + kUStages * kUStageBytes;
  Wgrad1Barriers* bars = reinterpret_cast<Wgrad1Barriers*>(tring + kTStages * kTStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // products 3: Uh*Th + Ul*Th + Uh*Tl;  2: Uh*Th + Uh*Tl (one-plane gradient);  1: Uh*Th
  const int tplanes = (p.products >= 2) ? 2 : 1;
  const int uplanes = (p.products == 3) ? 2 : 1;
  const int ntiles = p.N * p.nptile;
  const int co0 = blockIdx.y * 128;
  // gradient ring: stages hold only the planes in use (8 one-plane stages = twice the bytes in flight per SM)
  const int ustage_bytes = uplanes * kUPlaneBytes;
  const int ustages = (uplanes == 1) ? kUMaxStages : kUStages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTStages; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 1); }
    for (int i = 0; i < kUMaxStages; ++i) { mbar_init(&bars->ufull[i], 1); mbar_init(&bars->uempty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(&bars->tmem_base, 32);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < kTStages) {
    // Toeplitz producers (fp16 planes of the waveform): warp q fills stage q for local tiles q, q + 2, ...
    const int q = warp;
    uint32_t i = q;
    for (int tile = blockIdx.x + q * gridDim.x; tile < ntiles; tile += kTStages * gridDim.x, i += kTStages) {
      const int n = tile / p.nptile;
      const int p0 = (tile % p.nptile) * kTileN;
      float xv[39];
      toeplitz_load_strip(p.x + size_t(n) * p.L, 1, 0.f, 1.f, p0 - 15 + 8 * lane, p.L, xv);
      mbar_wait(&bars->tempty[q], ((i / kTStages) & 1) ^ 1);
      toeplitz_store<false>(xv, smem_u32(tring + q * kTStageBytes + lane * kGroupStride), tplanes, kTPlaneBytes);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tfull[q]);
    }
  } else if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_uh);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / p.nptile;
        const int p0 = (tile % p.nptile) * kTileN;
        for (int j = 0; j < kTileN / kSub; ++j, ++it) {
          const int s = it % ustages;
          mbar_wait(&bars->uempty[s], ((it / ustages) & 1) ^ 1);
          uint8_t* base = uring + s * ustage_bytes;
          mbar_arrive_expect_tx(&bars->ufull[s], uplanes * kUPlaneBytes);
          for (int pl = 0; pl < uplanes; ++pl) {
            const CUtensorMap* m = pl ? &tm_ul : &tm_uh;
            tma_load_3d(base + pl * kUPlaneBytes, m, &bars->ufull[s], co0, p0 + j * kSub, n);
            tma_load_3d(base + pl * kUPlaneBytes + kUHalfBytes, m, &bars->ufull[s], co0 + 64, p0 + j * kSub, n);
          }
        }
      }
    }
  } else if (warp == 5) {
    {
      const bool elected = elect_one_sync();   // whole warp runs the loop (uniform descriptor math), one lane issues
      // M = 128 (co), N = 32 (taps), both operands MN-major, both fp16
      const uint32_t idesc = make_idesc_f16(128, 32) | (1u << 15) | (1u << 16);
      uint32_t i = 0, uit = 0, first = 1;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int ts = i % kTStages;
        mbar_wait(&bars->tfull[ts], (i / kTStages) & 1);
        tc_fence_after_sync();
        const uint32_t th = smem_u32(tring + ts * kTStageBytes), tl = th + kTPlaneBytes;
        for (int j = 0; j < kTileN / kSub; ++j, ++uit) {
          const int us = uit % ustages;
          mbar_wait(&bars->ufull[us], (uit / ustages) & 1);
          tc_fence_after_sync();
          const uint32_t uh = smem_u32(uring + us * ustage_bytes), ul = uh + kUPlaneBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t arow = kk * 16 * 128;                       // 16 positions down the dU tile
            const uint32_t brow = (8 * j + 2 * kk) * kGroupStride;      // two 8-position row groups per K step
            if (elected) umma_f16(tmem_base, make_smem_desc(uh + arow, kUHalfBytes, 1024, kLayoutSW128),
                     make_smem_desc(th + brow, kGroupStride, 128, kLayoutNone), idesc, first ? 0u : 1u);
            first = 0;
            if (uplanes == 2 && elected)
              umma_f16(tmem_base, make_smem_desc(ul + arow, kUHalfBytes, 1024, kLayoutSW128),
                       make_smem_desc(th + brow, kGroupStride, 128, kLayoutNone), idesc, 1);
            if (tplanes == 2 && elected)
              umma_f16(tmem_base, make_smem_desc(uh + arow, kUHalfBytes, 1024, kLayoutSW128),
                       make_smem_desc(tl + brow, kGroupStride, 128, kLayoutNone), idesc, 1);
          }
          if (elected) umma_commit(&bars->uempty[us]);
        }
        if (elected) umma_commit(&bars->tempty[ts]);
      }
      if (elected) umma_commit(&bars->done);
    }
  }
  if (warp < 4) {
    // final epilogue: TMEM lane = co, column = tap
    mbar_wait(&bars->done, 0);
    tc_fence_after_sync();
    float v[32];
    tmem_ld_32x32(tmem_base + (uint32_t(warp * 32) << 16), v);
    const int co = co0 + warp * 32 + lane;
    const bool any = int(blockIdx.x) < ntiles;
    if (co < p.cout) {
      float* out = p.partial + size_t(blockIdx.x) * 32 * p.cout + co;
#pragma unroll
      for (int k = 0; k < 32; ++k) out[size_t(k) * p.cout] = any ? v[k] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 32);
  }
}

int launch_wgrad1_tc(const float* x, const __half* du_hi, const __half* du_lo, int N, int L, int cout, int products,
                     float* partial, size_t partial_bytes, int* nsplit_out, cudaStream_t stream) {
  using namespace w1;
  if (N <= 0 || L <= 0 || cout <= 0 || cout % 8 != 0) return set_error(VM_ERR_SHAPE, "wgrad1: bad shape");
  if (products == 3 && du_lo == nullptr) return set_error(VM_ERR_SHAPE, "wgrad1: lo plane required for products = 3");
  Wgrad1Params p{};
  p.x = x; p.N = N; p.L = L; p.cout = cout; p.products = products;
  p.nptile = (L + kTileN - 1) / kTileN;
  p.partial = partial;
  const int nco = (cout + 127) / 128;
  int gx = max(1, num_sms() / nco);
  gx = min(gx, N * p.nptile);
  if (size_t(gx) * 32 * cout * 4 > partial_bytes) return set_error(VM_ERR_SHAPE, "wgrad1: partial buffer too small");
  CUtensorMap uh, ul;
  const uint64_t dims[3] = {uint64_t(cout), uint64_t(L), uint64_t(N)};
  const uint64_t str[2] = {uint64_t(cout) * 2, uint64_t(L) * cout * 2};
  const uint32_t box[3] = {64, kSub, 1};
  int rc;
  if ((rc = make_tensor_map(&uh, du_hi, 3, dims, str, box, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&ul, products == 3 ? du_lo : du_hi, 3, dims, str, box, VM_SWIZZLE_128B))) return rc;
  cudaError_t e = cudaFuncSetAttribute(wgrad1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad1: cudaFuncSetAttribute");
  wgrad1_tc_kernel<<<dim3(gx, nco), kThreads, kSmemBytes, stream>>>(uh, ul, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad1: launch");
  *nsplit_out = gx;
  return VM_OK;
}

}  // namespace vm
