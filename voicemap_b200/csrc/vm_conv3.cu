// Blocks 2-4 of the voicemap encoder: Conv1D(k=3,'same') + bias + ReLU + BatchNorm(eval) + MaxPool1D(2),
// optionally merged with GlobalMaxPool1D (block 4).  Reference: voicemap/models.py:22-37.
//
// Implicit GEMM on tcgen05 tensor cores, transposed so that D[cout, position]:
//   A (M=128)  = packed weights  Wp[plane][tap][cout][cin]   (K-major, TMA 2D, SWIZZLE_128B)
//   B (N=256)  = activations     X[plane][n][l][cin]         (K-major = channels-last, TMA 3D, SWIZZLE_128B)
//   D          = 128 cout lanes x 256 position columns, fp32 in TMEM (double buffered: 2 x 256 columns)
// The three taps reuse ONE halo tile of 258 position rows: tap t reads the same shared-memory tile through a
// descriptor whose start address is shifted by t rows (t*128 B); 'same' zero padding is TMA out-of-bounds fill.
//
// fp32-grade accuracy from fp16 tensor cores: every fp32 operand is carried as two fp16 planes (hi, lo) with
// x = hi + lo to ~2^-22; D = Xh*Wh + Xl*Wh + Xh*Wl (three MMAs per K step, fp32 accumulate).  `products == 1`
// keeps only Xh*Wh (throughput mode, ~2^-11 operand rounding).  `products == 2` computes the two correction products
// as ONE kind::f8f6f4 MMA over e5m2 byte pairs (the "Q" planes of vm_common.cuh: K doubles, the instruction count per
// K chunk drops from 12 to 8), which is enough for them because they are 2^-12 of the main term.
//
// Epilogue (4 warps, thread = one cout channel, columns = consecutive positions): pooling is done on the raw
// accumulators first -- weights of channels with a negative BN scale are packed negated (sigma = -1) so that
// max-pooling commutes with the affine:  y = s * relu(sigma * max(acc) + bias) + t.
#include "vm_common.cuh"
#include "vm_kernels.h"

namespace vm {

namespace c3 {
constexpr int kTileN = 256;              // positions per tile (MMA N)
constexpr int kTileM = 128;              // cout per slab (MMA M)
constexpr int kKC = 64;                  // channels per K chunk (one 128-byte swizzle row of fp16)
constexpr int kXRows = 264;              // 258 halo rows, padded to a multiple of 8
constexpr int kXPlaneBytes = kXRows * 128;         // 33792
constexpr int kXMainBytes = 256 * 128;             // 32768 (rows 0..255)
constexpr int kXSlotBytes = 2 * kXPlaneBytes;      // hi + lo
constexpr int kXStages = 2;                          // two-plane slots; a one-plane launch cuts them into 4 half slots
constexpr int kXMaxStages = 4;
constexpr int kWTileBytes = kTileM * 128;          // 16384
constexpr int kWStages = 4;
constexpr int kStagePos = 16;                       // pooled positions per output granule
constexpr int kStageBoxBytes = kStagePos * 128;     // one TMA store box: 16 positions x 64 channels fp16
constexpr int kStageBufBytes = 4 * kStageBoxBytes;  // [plane][channel half][pos][64 ch] = 8 KB
constexpr int kStageBytes = 2 * kStageBufBytes;     // double-buffered
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 8;                         // per epilogue set: 2 per TMEM lane quarter (threads = (4 + 8 * sets) * 32)
constexpr int kSmemBytes =
    kXStages * kXSlotBytes + kWStages * kWTileBytes + kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
}  // namespace c3

struct __align__(8) Conv3Barriers {
  uint64_t xfull[c3::kXMaxStages], xempty[c3::kXMaxStages];
  uint64_t wfull[c3::kWStages], wempty[c3::kWStages];
  uint64_t tfull[2], tempty[2];
  uint64_t sfull[2], sempty[2];   // pooled-output staging buffers: epilogue warps -> store warp -> epilogue warps
  uint32_t tmem_base;
};

// Tile schedule.  tile = (position tile) * nslab + (cout slab).  A CTA takes "units" of `spu` consecutive tiles round
// robin: spu = 1 deals single tiles (the slabs of a position tile run on neighbouring CTAs at the same time and share
// the X tile through L2); spu = nslab keeps all slabs of a position tile on one CTA, which is used when the whole K
// extent of the X tile fits the shared-memory ring (Cin <= 128): X is then loaded once per position tile instead of
// once per slab, and the next position tile's X has nslab tiles of MMA time to arrive from HBM.
struct TileIter {
  int unit, s, spu, nunits, stride;
  __device__ TileIter(int first, int stride_, int ntiles, int spu_)
      : unit(first), s(0), spu(spu_), nunits(ntiles / spu_), stride(stride_) {}
  __device__ bool valid() const { return unit < nunits; }
  __device__ int tile() const { return unit * spu + s; }
  __device__ bool first() const { return s == 0; }
  __device__ bool last() const { return s == spu - 1; }
  __device__ void next() {
    if (++s == spu) { s = 0; unit += stride; }
  }
};

// 16 accumulator columns of one channel -> 8 pooled outputs -> bias/ReLU/BN clamp form -> staging rows.
// kSecond: 0 = hi plane only, 1 = fp16 residual plane, 2 = e5m2x2 Q plane.
template <bool kClampHi, int kSecond>
__device__ __forceinline__ void pooled_granule(const float4& ep, const uint32_t (&r)[16], uint32_t st_h, uint32_t st_l) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float y = apply_epi_pool2<kClampHi>(ep, __uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
    __half h;
    if (kSecond == 2) {
      uint16_t q;
      split_f16_q(y, h, q);
      sts_b16(st_l + j * 128, q);
    } else {
      __half l;
      split_f32(y, h, l);
      if (kSecond == 1) sts_u16(st_l + j * 128, l);
    }
    sts_u16(st_h + j * 128, h);
  }
}

// Pooled epilogue of one accumulator tile for one warp: 8 granules of 16 columns, TMEM loads software-pipelined (the
// next granule's columns are in flight while this one is pooled, converted and staged).  Granules are handed to the
// store warp through the sfull / sempty mbarriers of the two staging buffers.
template <bool kClampHi, int kSecond>
__device__ __forceinline__ void pooled_tile(const float4& ep, uint32_t taddr, uint8_t* stage, uint32_t thread_off,
                                            Conv3Barriers* bars, int buf, uint32_t& gcount) {
  using namespace c3;
  auto granule = [&](const uint32_t (&r)[16]) {
    const int sb = gcount & 1;
    const uint32_t st_h = smem_u32(stage + sb * kStageBufBytes) + thread_off;
    // the store of the granule that used this staging buffer last (two granules ago) has read it out
    if (gcount >= 2) mbar_wait(&bars->sempty[sb], ((gcount >> 1) - 1) & 1);
    pooled_granule<kClampHi, kSecond>(ep, r, st_h, st_h + 2 * kStageBoxBytes);
    fence_proxy_async_smem();          // this thread's staging writes -> visible to the TMA store
    mbar_arrive(&bars->sfull[sb]);     // 256 arrivals = granule complete; the store warp takes it from here
    ++gcount;
  };
  uint32_t ra[16], rb[16];
  tmem_ld_32x16_issue(taddr, ra);
#pragma unroll 1
  for (int gr = 0; gr < kTileN / 32; gr += 2) {
    tmem_ld_wait(ra);
    tmem_ld_32x16_issue(taddr + (gr + 1) * 32, rb);
    granule(ra);
    tmem_ld_wait(rb);
    if (gr + 2 < kTileN / 32) {
      tmem_ld_32x16_issue(taddr + (gr + 2) * 32, ra);
    } else {  // all TMEM reads of this accumulator are done
      tc_fence_before_sync();
      mbar_arrive(&bars->tempty[buf]);
    }
    granule(rb);
  }
}

// Train-mode epilogue of one accumulator tile for one warp (the two warps of a lane quarter take 8 of every 16
// columns): u = relu(acc + bias) per position -> per-channel {sum, sum of squares} for the batch statistics, the
// un-pooled activation as fp16 + arg-max flag (encode_u) and the fp32 extreme of every MaxPool(2) window (max, or
// min for channels whose BatchNorm scale is negative: the value the normalised maximum comes from).  Granule = 16
// positions: staging buffer = [u16: 2 boxes of 16 pos x 64 ch][extremes: 4 boxes of 8 windows x 32 ch fp32] = 8 KB,
// handed to the store warp like the pooled granules.
// kSets epilogue sets of 8 warps: with two, set s takes the granules g = s, s + 2, ... and owns staging buffer s
// (gcount counts this set's granules; the store warp walks all granules in order, buffer g & 1).
template <int kSets>
__device__ __forceinline__ void raw2_tile(const float4& ep, bool neg, uint32_t taddr, uint8_t* stage, int ch, int chalf,
                                          int set, Conv3Barriers* bars, int buf, uint32_t& gcount, int pos_base, int L,
                                          int lvalid, float& s1, float& s2) {
  using namespace c3;
  const uint32_t off_u = (ch >> 6) * 2048 + (chalf * 8) * 128 + (ch & 63) * 2;
  const uint32_t off_m = 4096 + (ch >> 5) * 1024 + (chalf * 4) * 128 + (ch & 31) * 4;
  auto granule = [&](const uint32_t (&r)[8], int gr) {
    const int sb = (kSets == 2) ? set : int(gcount & 1);
    const uint32_t use = (kSets == 2) ? gcount : (gcount >> 1);   // how often this buffer has been filled before
    const uint32_t st = smem_u32(stage + sb * kStageBufBytes);
    if (use >= 1) mbar_wait(&bars->sempty[sb], (use - 1) & 1);
    const int pos0 = pos_base + gr * 16 + chalf * 8;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float a = apply_epi(ep, __uint_as_float(r[2 * w]));
      const float b = apply_epi(ep, __uint_as_float(r[2 * w + 1]));
      if (pos0 + 2 * w < L) { s1 += a; s2 = fmaf(a, a, s2); }
      if (pos0 + 2 * w + 1 < L) { s1 += b; s2 = fmaf(b, b, s2); }
      const bool second = neg ? (b < a) : (b > a);          // first winner on ties
      const bool win = pos0 + 2 * w + 1 < lvalid;           // the window exists ('valid' pooling drops an odd tail)
      sts_b16(st + off_u + (2 * w) * 128, encode_u(a, win && !second));
      sts_b16(st + off_u + (2 * w + 1) * 128, encode_u(b, win && second));
      sts_f32(st + off_m + w * 128, second ? b : a);
    }
    fence_proxy_async_smem();
    mbar_arrive(&bars->sfull[sb]);
    ++gcount;
  };
  uint32_t ra[8], rb[8];
  const uint32_t tcol = taddr + chalf * 8;
  constexpr int kStep = kSets;   // granules between two of this set's
  tmem_ld_32x8_issue(tcol + set * 16, ra);
#pragma unroll 1
  for (int gr = set; gr < kTileN / 16; gr += 2 * kStep) {
    tmem_ld_wait(ra);
    tmem_ld_32x8_issue(tcol + (gr + kStep) * 16, rb);
    granule(ra, gr);
    tmem_ld_wait(rb);
    if (gr + 2 * kStep < kTileN / 16) {
      tmem_ld_32x8_issue(tcol + (gr + 2 * kStep) * 16, ra);
    } else {
      tc_fence_before_sync();
      mbar_arrive(&bars->tempty[buf]);
    }
    granule(rb, gr + kStep);
  }
}

// fp32 epilogue of one accumulator tile for one warp (dgrad): granule = 16 positions x 128 channels fp32 = 4 boxes of
// 32 channels, pipelined TMEM loads, staging buffers handed to the store warp (see raw2_tile).
// kReduce: also accumulate, per channel, the BatchNorm-backward sums of the block below from the values in registers
// (red_*: the thread's channel constants; erow: its column of the window extremes of this clip; lrem: valid positions
// from this tile's first one).
template <bool kReduce, int kSets>
__device__ __forceinline__ void linear_tile(float unscale, uint32_t taddr, uint8_t* stage, int ch, int chalf, int set,
                                            Conv3Barriers* bars, int buf, uint32_t& gcount, const float* erow,
                                            size_t estride, int lrem, float mk, float mean, float rstd, float sabs,
                                            float& s1, float& s2, float& amax) {
  using namespace c3;
  const uint32_t off = (ch >> 5) * 2048 + (chalf * 8) * 128 + (ch & 31) * 4;
  int gr_pos = set * 16 + chalf * 8;   // first position (relative to the tile) of the granule being processed
  auto load_ext = [&](float (&e)[8], int pos) {   // the thread's 8 window extremes of the granule starting at pos
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = (pos + j < lrem) ? __ldg(erow + size_t(pos + j) * estride) : mean;
  };
  // the extremes are fetched TWO granules ahead (registers ea / eb alternate like the TMEM loads): a global load has
  // two granules of epilogue work to land, the first version (loads issued inside the granule that used them) exposed
  // one DRAM latency per granule and doubled the kernel time.  (Also tried: prefetch.global.L2 of the next tile's
  // extremes at tile start -- 8 % slower, the epilogue is issue-bound once the latency is covered.)
  auto granule = [&](const uint32_t (&r)[8], float (&e)[8]) {
    const int sb = (kSets == 2) ? set : int(gcount & 1);
    const uint32_t use = (kSets == 2) ? gcount : (gcount >> 1);
    const uint32_t st = smem_u32(stage + sb * kStageBufBytes) + off;
    if (use >= 1) mbar_wait(&bars->sempty[sb], (use - 1) & 1);
    float dy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dy[j] = __uint_as_float(r[j]) * unscale;
      sts_f32(st + j * 128, dy[j]);
    }
    fence_proxy_async_smem();
    mbar_arrive(&bars->sfull[sb]);
    if (kReduce) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = (gr_pos + j < lrem) ? dy[j] * mk : 0.f;
        s1 += d;
        s2 = fmaf(d, (e[j] - mean) * rstd, s2);
        amax = fmaxf(amax, fabsf(d) * sabs);
      }
      load_ext(e, gr_pos + 32 * kSets);
    }
    gr_pos += 16 * kSets;
    ++gcount;
  };
  uint32_t ra[8], rb[8];
  float ea[8], eb[8];
  if (kReduce) {
    load_ext(ea, gr_pos);
    load_ext(eb, gr_pos + 16 * kSets);
  }
  const uint32_t tcol = taddr + chalf * 8;
  constexpr int kStep = kSets;
  tmem_ld_32x8_issue(tcol + set * 16, ra);
#pragma unroll 1
  for (int gr = set; gr < kTileN / 16; gr += 2 * kStep) {
    tmem_ld_wait(ra);
    tmem_ld_32x8_issue(tcol + (gr + kStep) * 16, rb);
    granule(ra, ea);
    tmem_ld_wait(rb);
    if (gr + 2 * kStep < kTileN / 16) {
      tmem_ld_32x8_issue(tcol + (gr + 2 * kStep) * 16, ra);
    } else {
      tc_fence_before_sync();
      mbar_arrive(&bars->tempty[buf]);
    }
    granule(rb, eb);
  }
}

// kSets: epilogue sets of 8 warps.  1: the pooled / global-max epilogues of the eval forward.  2: the train-mode
// forward and dgrad, whose epilogues do 2-3x the work per accumulator column (statistics, encoding, window extremes;
// fused BatchNorm-backward sums) and left the tensor pipe 42-64 % busy with one set.
// kXS: slots of the X ring (2 two-plane slots, or 4 one-plane slots in the same shared memory) -- a compile-time constant:
// as a kernel parameter the slot index cost two integer divisions per K chunk in the MMA-issuing thread and the eval
// forward lost 2-3 % (same-box A/B, profiles/r02_ab_forward_fused_head_and_xring.log).
template <int kSets, int kXS>
__global__ void __launch_bounds__((4 + 8 * kSets) * 32, 1)
conv3_kernel(const __grid_constant__ CUtensorMap tm_xh_main, const __grid_constant__ CUtensorMap tm_xh_halo,
             const __grid_constant__ CUtensorMap tm_xl_main, const __grid_constant__ CUtensorMap tm_xl_halo,
             const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
             const __grid_constant__ CUtensorMap tm_oh, const __grid_constant__ CUtensorMap tm_ol,
             const Conv3Params p) {
  using namespace c3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xring = smem;
  uint8_t* wring = smem + kXStages * kXSlotBytes;
  uint8_t* stage = wring + kWStages * kWTileBytes;
  Conv3Barriers* bars = reinterpret_cast<Conv3Barriers*>(stage + kStageBytes);

  constexpr int kXSlot = (kXS == kXStages) ? kXSlotBytes : kXPlaneBytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = p.N * p.nptile * p.nslab;
  const int spu = p.slabs_per_unit;
  const int wplanes = (p.products >= 2) ? 2 : 1;
  const bool mixed = (p.products == 2);   // second plane of X and W is the e5m2x2 Q plane

  if (threadIdx.x == 0) {
    for (int i = 0; i < kXMaxStages; ++i) { mbar_init(&bars->xfull[i], 1); mbar_init(&bars->xempty[i], 1); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&bars->wfull[i], 1); mbar_init(&bars->wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], 8 * kSets * 32); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->sfull[i], kEpiWarps * 32); mbar_init(&bars->sempty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== X producer: one halo tile (hi+lo) per (tile, K chunk) =====================
    if (lane == 0) {
      tma_prefetch_desc(&tm_xh_main); tma_prefetch_desc(&tm_xh_halo);
      tma_prefetch_desc(&tm_xl_main); tma_prefetch_desc(&tm_xl_halo);
      uint32_t it = 0;
      for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next()) {
        if (!ti.first()) continue;   // the unit's later slabs reuse the resident X tile
        const int pt_lin = ti.tile() / p.nslab;
        const int n = pt_lin / p.nptile;
        const int p0 = (pt_lin % p.nptile) * kTileN;
        for (int c = 0; c < p.nchunk; ++c, ++it) {
          const int s = it % kXS;
          mbar_wait(&bars->xempty[s], ((it / kXS) & 1) ^ 1);
          uint8_t* dst = xring + s * kXSlot;
          const int xplanes = p.x_single ? 1 : wplanes;
          mbar_arrive_expect_tx(&bars->xfull[s], xplanes * kXPlaneBytes);
          tma_load_3d(dst, &tm_xh_main, &bars->xfull[s], c * kKC, p0 - 1, n);
          tma_load_3d(dst + kXMainBytes, &tm_xh_halo, &bars->xfull[s], c * kKC, p0 + 255, n);
          if (xplanes == 2) {
            tma_load_3d(dst + kXPlaneBytes, &tm_xl_main, &bars->xfull[s], c * kKC, p0 - 1, n);
            tma_load_3d(dst + kXPlaneBytes + kXMainBytes, &tm_xl_halo, &bars->xfull[s], c * kKC, p0 + 255, n);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== W producer: one [128 cout x 64 cin] tile per (tile, chunk, tap, plane) ==========
    if (lane == 0) {
      tma_prefetch_desc(&tm_wh); tma_prefetch_desc(&tm_wl);
      uint32_t it = 0;
      for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next()) {
        const int slab = ti.tile() % p.nslab;
        for (int c = 0; c < p.nchunk; ++c) {
          for (int tap = 0; tap < 3; ++tap) {
            for (int pl = 0; pl < wplanes; ++pl, ++it) {
              const int s = it % kWStages;
              mbar_wait(&bars->wempty[s], ((it / kWStages) & 1) ^ 1);
              mbar_arrive_expect_tx(&bars->wfull[s], kWTileBytes);
              tma_load_2d(wring + s * kWTileBytes, pl == 0 ? &tm_wh : &tm_wl, &bars->wfull[s], c * kKC,
                          tap * p.cout_pad + slab * kTileM);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      const bool elected = elect_one_sync();
      uint32_t xit = 0, wit = 0, tit = 0;
      for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next(), ++tit) {
        const int tile = ti.tile();
        const int buf = tit & 1;
        // the last position tile of a clip is usually ragged: issue only as many MMA columns (multiple of 16) as
        // there are positions left -- the tensor pipe is the bound, so unused columns are pure waste
        const int p0 = ((tile / p.nslab) % p.nptile) * kTileN;
        const int ncols = min(kTileN, (p.L - p0 + 15) & ~15);
        const uint32_t idesc = make_idesc_f16(kTileM, ncols);
        const uint32_t idesc8 = make_idesc_f8(kTileM, ncols);
        mbar_wait(&bars->tempty[buf], ((tit >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * kTileN;
        uint32_t acc = 0;
        for (int c = 0; c < p.nchunk; ++c) {
          const uint32_t xi = xit + c;   // X chunks of this unit (loaded once, by its first tile)
          const int xs = xi % kXS;
          if (ti.first()) {
            mbar_wait(&bars->xfull[xs], (xi / kXS) & 1);
            tc_fence_after_sync();
          }
          const uint32_t xh = smem_u32(xring + xs * kXSlot);
          const uint32_t xl = xh + kXPlaneBytes;
          for (int tap = 0; tap < 3; ++tap) {
            const uint32_t bh = xh + tap * 128, bl = xl + tap * 128;
            {  // W hi: Xh*Wh (+ Xl*Wh)
              const int ws = wit % kWStages;
              mbar_wait(&bars->wfull[ws], (wit / kWStages) & 1);
              tc_fence_after_sync();
              const uint32_t wa = smem_u32(wring + ws * kWTileBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (elected)
                  umma_f16(d_tmem, make_smem_desc(wa + k * 32, 16, 1024, kLayoutSW128),
                           make_smem_desc(bh + k * 32, 16, 1024, kLayoutSW128), idesc, acc);
                acc = 1;
              }
              if (wplanes == 2 && !mixed && !p.x_single) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (elected)
                    umma_f16(d_tmem, make_smem_desc(wa + k * 32, 16, 1024, kLayoutSW128),
                             make_smem_desc(bl + k * 32, 16, 1024, kLayoutSW128), idesc, 1);
              }
              if (elected) umma_commit(&bars->wempty[ws]);
              ++wit;
            }
            if (wplanes == 2) {  // second W plane: Xh*Wl (fp16 lo planes), or Xq*Wq = Xl*Wh + Xh*Wl (e5m2 pairs)
              const int ws = wit % kWStages;
              mbar_wait(&bars->wfull[ws], (wit / kWStages) & 1);
              tc_fence_after_sync();
              const uint32_t wa = smem_u32(wring + ws * kWTileBytes);
              if (mixed) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (elected)
                    umma_f8(d_tmem, make_smem_desc(wa + k * 32, 16, 1024, kLayoutSW128),
                            make_smem_desc(bl + k * 32, 16, 1024, kLayoutSW128), idesc8, 1);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (elected)
                    umma_f16(d_tmem, make_smem_desc(wa + k * 32, 16, 1024, kLayoutSW128),
                             make_smem_desc(bh + k * 32, 16, 1024, kLayoutSW128), idesc, 1);
              }
              if (elected) umma_commit(&bars->wempty[ws]);
              ++wit;
            }
          }
          if (ti.last() && elected) umma_commit(&bars->xempty[xs]);
        }
        if (ti.last()) xit += p.nchunk;
        if (elected) umma_commit(&bars->tfull[buf]);
      }
    }
  } else if (warp == 2) {
    // ===================== store warp (pooled outputs): staging buffer -> TMA bulk tensor stores =====================
    // The epilogue warps hand over each 8 KB output granule through an mbarrier pair per staging buffer instead of
    // meeting in a CTA-wide named barrier: no epilogue warp waits for another one or for the store issue, they only
    // wait (rarely) for a staging buffer whose previous store has not been read out of shared memory yet.
    if (lane == 0 && p.out_u16 != nullptr) {
      // train-mode forward: 16 granules per tile, each 2 boxes of encoded activations + 4 boxes of window extremes
      uint32_t g = 0;
      for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next()) {
        const int tile = ti.tile();
        const int slab = tile % p.nslab;
        const int pt_lin = tile / p.nslab;
        const int n = pt_lin / p.nptile;
        const int p0 = (pt_lin % p.nptile) * kTileN;
        for (int gr = 0; gr < kTileN / 16; ++gr, ++g) {
          const int b = g & 1;
          mbar_wait(&bars->sfull[b], (g >> 1) & 1);
          const uint8_t* sbuf = stage + b * kStageBufBytes;
          const int pos = p0 + gr * 16;
          if (pos < p.L) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c0 = slab * kTileM + half * 64;
              if (c0 < p.cout) tma_store_3d(&tm_oh, sbuf + half * 2048, c0, pos, n);
            }
            if ((pos >> 1) < p.lout) {
#pragma unroll
              for (int b4 = 0; b4 < 4; ++b4) {
                const int c0 = slab * kTileM + b4 * 32;
                if (c0 < p.cout) tma_store_3d(&tm_ol, sbuf + 4096 + b4 * 1024, c0, pos >> 1, n);
              }
            }
          }
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&bars->sempty[b]);
        }
      }
      tma_store_wait_all<0>();
    } else if (lane == 0 && p.out_f32 != nullptr) {
      // dgrad: 16 granules per tile, each 4 boxes of 16 positions x 32 channels fp32
      uint32_t g = 0;
      for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next()) {
        const int tile = ti.tile();
        const int slab = tile % p.nslab;
        const int pt_lin = tile / p.nslab;
        const int n = pt_lin / p.nptile;
        const int p0 = (pt_lin % p.nptile) * kTileN;
        for (int gr = 0; gr < kTileN / 16; ++gr, ++g) {
          const int b = g & 1;
          mbar_wait(&bars->sfull[b], (g >> 1) & 1);
          const uint8_t* sbuf = stage + b * kStageBufBytes;
          const int pos = p0 + gr * 16;
          if (pos < p.L) {
#pragma unroll
            for (int b4 = 0; b4 < 4; ++b4) {
              const int c0 = slab * kTileM + b4 * 32;
              if (c0 < p.cout) tma_store_3d(&tm_oh, sbuf + b4 * 2048, c0, pos, n);
            }
          }
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&bars->sempty[b]);
        }
      }
      tma_store_wait_all<0>();
    } else if (lane == 0 && p.gmax_partial == nullptr && p.out_f32 == nullptr) {
      uint32_t g = 0;
      for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next()) {
        const int tile = ti.tile();
        const int slab = tile % p.nslab;
        const int pt_lin = tile / p.nslab;
        const int n = pt_lin / p.nptile;
        const int p0 = (pt_lin % p.nptile) * kTileN;
        for (int gr = 0; gr < kTileN / 32; ++gr, ++g) {
          const int b = g & 1;
          mbar_wait(&bars->sfull[b], (g >> 1) & 1);
          const uint8_t* sbuf = stage + b * kStageBufBytes;
          const int pos = (p0 >> 1) + gr * kStagePos;
          if (pos < p.lout) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c0 = slab * kTileM + half * 64;
              if (c0 < p.cout) {
                tma_store_3d(&tm_oh, sbuf + half * kStageBoxBytes, c0, pos, n);
                if (wplanes == 2) tma_store_3d(&tm_ol, sbuf + (2 + half) * kStageBoxBytes, c0, pos, n);
              }
            }
          }
          tma_store_commit();
          tma_store_wait_read<0>();          // this thread has nothing else to do: hand the buffer back as soon as
          mbar_arrive(&bars->sempty[b]);     // the TMA unit has read it out of shared memory
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = cout channel, columns = positions =====================
    // 8 warps: q = TMEM lane quarter, chalf = which half of every output granule's columns.  Output granules go
    // through a double-buffered 2 x 8 KB staging area and leave with TMA bulk tensor stores (positions / channels
    // out of range are clipped by the TMA unit).  Per granule there is one named barrier: before arriving, the
    // leader waits until the store issued one granule earlier has drained its buffer (that store had a whole
    // granule of compute to finish), so the buffer written next is known to be free by everyone who passes.
    const int q = warp & 3;
    const int set = (warp - 4) >> 3;          // 0 with one set
    const int chalf = ((warp - 4) & 7) >> 2;
    const int ch = q * 32 + lane;
    const int rows = 2 * kSets, row = chalf * kSets + set;   // partial-sum rows per tile / this thread's row
    const int lvalid = p.lout * 2;  // 'valid' pooling drops an odd tail position
    uint32_t tit = 0, gcount = 0;   // gcount: granules staged so far (selects the staging buffer)
    for (TileIter ti(blockIdx.x, gridDim.x, ntiles, spu); ti.valid(); ti.next(), ++tit) {
      const int tile = ti.tile();
      const int slab = tile % p.nslab;
      const int pt_lin = tile / p.nslab;
      const int n = pt_lin / p.nptile;
      const int pt = pt_lin % p.nptile;
      const int p0 = pt * kTileN;
      const int buf = tit & 1;
      const int co = slab * kTileM + ch;
      const float4 ep = p.epi[co];  // {a, c, lo, hi} (see apply_epi); padded channels hold zeros
      // one polling warp; the other seven block in bar.sync instead of spinning on the mbarrier
      if (((warp - 4) & 7) == 0) mbar_wait(&bars->tfull[buf], (tit >> 1) & 1);
      named_bar_sync(1 + set, 256);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * kTileN;
      if (p.out_u16 != nullptr) {
        const bool neg = (p.sign_src != nullptr && co < p.cout) ? (p.sign_src[co] < 0.f) : false;
        float s1 = 0.f, s2 = 0.f;
        raw2_tile<kSets>(ep, neg, taddr, stage, ch, chalf, set, bars, buf, gcount, p0, p.L, lvalid, s1, s2);
        if (p.stat_partial != nullptr)
          p.stat_partial[(size_t(n) * (rows * p.nptile) + rows * pt + row) * p.cout_pad + co] = make_float2(s1, s2);
      } else if (kSets != 1 && p.out_f32 == nullptr) {
        __trap();   // the pooled / global-max epilogues are instantiated with one set only (host-side dispatch)
      } else if (p.gmax_partial != nullptr) {
        float m = -INFINITY;
#pragma unroll 1
        for (int gg = 0; gg < kTileN / 64; ++gg) {
          const int g = 2 * gg + chalf;
          float v[32];
          tmem_ld_32x32(taddr + g * 32, v);
          const int pos = p0 + g * 32;
          if (pos + 32 <= lvalid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (pos + j < lvalid) m = fmaxf(m, v[j]);
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&bars->tempty[buf]);
        // the two column halves keep separate partial rows: (N, 2*nptile, cout_pad), raw accumulator maxima
        p.gmax_partial[(size_t(n) * (2 * p.nptile) + 2 * pt + chalf) * p.cout_pad + co] = m;
      } else if (p.out_f32 != nullptr) {
        // plain fp32 output of the accumulator (dgrad: dX = conv3(dU, flipped / transposed weights)); the incoming
        // gradient planes carry a power-of-two scale (vm_common.cuh), taken out here
        const float unscale = (p.grad_absmax != nullptr)
                                  ? 1.0f / grad_scale_from_absmax(__uint_as_float(*p.grad_absmax)) : 1.0f;
        float s1 = 0.f, s2 = 0.f, amax = 0.f;
        if (p.red.partial != nullptr) {
          // BatchNorm-backward sums of the block below, from the gradient values this thread is about to store
          const bool live = co < p.cout;
          const int g = n / (p.N / p.red.G);
          const float4 bc = live ? p.red.bn_const[size_t(g) * p.cout + co] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float mk = live ? (p.red.mask ? p.red.mask[size_t(n) * p.cout + co] : 1.f) : 0.f;
          const float* erow = p.red.ext + (size_t(n) * p.L + p0) * p.cout + (live ? co : 0);
          linear_tile<true, kSets>(unscale, taddr, stage, ch, chalf, set, bars, buf, gcount, erow, size_t(p.cout),
                                   live ? p.L - p0 : 0, mk, bc.z, bc.w, fabsf(bc.x), s1, s2, amax);
          p.red.partial[(size_t(n) * (rows * p.nptile) + rows * pt + row) * p.cout_pad + co] = make_float2(s1, s2);
          for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
          if (lane == 0 && amax > 0.f) atomicMax(p.red.absmax, __float_as_uint(amax));
        } else {
          linear_tile<false, kSets>(unscale, taddr, stage, ch, chalf, set, bars, buf, gcount, nullptr, 0, 0, 0.f, 0.f,
                                    0.f, 0.f, s1, s2, amax);
        }
      } else {
        // pooled outputs: granule = 16 pooled positions x 128 channels x 2 planes; the two warps of a lane quarter
        // take 16 raw columns (8 pooled positions) each.
        const bool no_hi = epi_no_upper_clamp(ep);
        const uint32_t toff = (ch >> 6) * kStageBoxBytes + (ch & 63) * 2 + chalf * 8 * 128;
        const uint32_t ta = taddr + chalf * 16;
        if (mixed) {
          if (no_hi) pooled_tile<false, 2>(ep, ta, stage, toff, bars, buf, gcount);
          else pooled_tile<true, 2>(ep, ta, stage, toff, bars, buf, gcount);
        } else if (wplanes == 2) {
          if (no_hi) pooled_tile<false, 1>(ep, ta, stage, toff, bars, buf, gcount);
          else pooled_tile<true, 1>(ep, ta, stage, toff, bars, buf, gcount);
        } else {
          if (no_hi) pooled_tile<false, 0>(ep, ta, stage, toff, bars, buf, gcount);
          else pooled_tile<true, 0>(ep, ta, stage, toff, bars, buf, gcount);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
int launch_conv3(const __half* in_hi, const __half* in_lo, int N, int L, int cin, int cout, const __half* wpack,
                 const float* epi, __half* out_hi, __half* out_lo, float* gmax_partial, float* out_f32,
                 float* stat_partial, int linear, int products, int max_ctas, cudaStream_t stream,
                 const Conv3Extra& extra) {
  using namespace c3;
  if (N <= 0 || L <= 0) return set_error(VM_ERR_SHAPE, "conv3: N and L must be positive");
  // K chunks are 64 channels wide; a ragged last chunk is zero-filled by TMA in both operands
  if (cin % 8 != 0) return set_error(VM_ERR_UNSUPPORTED, "conv3: Cin must be a multiple of 8");
  if (cout % 8 != 0) return set_error(VM_ERR_UNSUPPORTED, "conv3: Cout must be a multiple of 8");
  if (products < 1 || products > 3) return set_error(VM_ERR_SHAPE, "conv3: products must be 1, 2 or 3");
  if (products == 2 && (out_f32 != nullptr || extra.x_single))
    return set_error(VM_ERR_UNSUPPORTED, "conv3: products=2 (e5m2 correction planes) is a forward mode");
  if (extra.x_single && products != 3) return set_error(VM_ERR_SHAPE, "conv3: x_single goes with products=3");
  if ((extra.out_u16 != nullptr) != (extra.out_ext != nullptr))
    return set_error(VM_ERR_SHAPE, "conv3: the train-mode forward needs both out_u16 and out_ext");
  if (gmax_partial == nullptr && out_hi == nullptr && out_f32 == nullptr && extra.out_u16 == nullptr)
    return set_error(VM_ERR_SHAPE, "conv3: no output given");
  if (L / 2 <= 0) return set_error(VM_ERR_SHAPE, "conv3: L must be >= 2");
  const int cout_pad = (cout + kTileM - 1) / kTileM * kTileM;

  Conv3Params p{};
  p.N = N; p.L = L; p.cin = cin; p.cout = cout; p.cout_pad = cout_pad;
  p.lout = L / 2;
  p.nptile = (L + kTileN - 1) / kTileN;
  p.nslab = cout_pad / kTileM;
  p.nchunk = (cin + kKC - 1) / kKC;
  p.products = products;
  p.epi = reinterpret_cast<const float4*>(epi);
  p.out_hi = out_hi; p.out_lo = out_lo; p.gmax_partial = gmax_partial;
  p.out_f32 = out_f32; p.stat_partial = reinterpret_cast<float2*>(stat_partial); p.linear = linear;
  p.x_single = extra.x_single;
  p.out_u16 = extra.out_u16; p.out_ext = extra.out_ext; p.sign_src = extra.sign_src;
  p.grad_absmax = extra.grad_absmax;
  p.red = extra.red;
  if (p.red.partial != nullptr) {
    if (!linear || out_f32 == nullptr || p.red.ext == nullptr || p.red.bn_const == nullptr || p.red.absmax == nullptr ||
        p.red.G <= 0 || N % p.red.G != 0)
      return set_error(VM_ERR_SHAPE, "conv3: the fused BatchNorm-backward reduction goes with the dgrad output");
    cudaError_t me = cudaMemsetAsync(p.red.absmax, 0, sizeof(unsigned int), stream);
    if (me != cudaSuccess) return set_cuda_error(me, "conv3: memset");
  }

  CUtensorMap xh_main, xh_halo, xl_main, xl_halo, wh, wl;
  // X planes: (N, L, Cin) fp16, dims fastest-first {Cin, L, N}
  const uint64_t xdims[3] = {uint64_t(cin), uint64_t(L), uint64_t(N)};
  const uint64_t xstr[2] = {uint64_t(cin) * 2, uint64_t(L) * cin * 2};
  const uint32_t box_main[3] = {kKC, 256, 1}, box_halo[3] = {kKC, 8, 1};
  int rc;
  if ((rc = make_tensor_map(&xh_main, in_hi, 3, xdims, xstr, box_main, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&xh_halo, in_hi, 3, xdims, xstr, box_halo, VM_SWIZZLE_128B))) return rc;
  const bool need_lo = products >= 2 && !extra.x_single;
  const __half* lo_src = need_lo ? in_lo : in_hi;
  if (need_lo && in_lo == nullptr) return set_error(VM_ERR_SHAPE, "conv3: second plane required for products>=2");
  if ((rc = make_tensor_map(&xl_main, lo_src, 3, xdims, xstr, box_main, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&xl_halo, lo_src, 3, xdims, xstr, box_halo, VM_SWIZZLE_128B))) return rc;
  // W planes: [plane][tap*cout_pad + cout][cin] fp16
  const uint64_t wdims[2] = {uint64_t(cin), uint64_t(3) * cout_pad};
  const uint64_t wstr[1] = {uint64_t(cin) * 2};
  const uint32_t wbox[2] = {kKC, kTileM};
  const __half* w_hi = wpack;
  // weight planes: [hi][lo][q]; products == 2 pairs the hi plane with the e5m2x2 plane
  const __half* w_lo = wpack + size_t(products == 2 ? 2 : 1) * 3 * cout_pad * cin;
  if ((rc = make_tensor_map(&wh, w_hi, 2, wdims, wstr, wbox, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&wl, w_lo, 2, wdims, wstr, wbox, VM_SWIZZLE_128B))) return rc;

  // pooled output planes (N, lout, cout) fp16 -- TMA store boxes of 64 channels x 16 positions
  CUtensorMap oh, ol;
  if (extra.out_u16 != nullptr) {
    // encoded activations (N, L, cout) 16-bit: boxes of 64 channels x 16 positions; window extremes (N, lout, cout)
    // fp32: boxes of 32 channels x 8 windows
    const uint64_t udims[3] = {uint64_t(cout), uint64_t(L), uint64_t(N)};
    const uint64_t ustr[2] = {uint64_t(cout) * 2, uint64_t(L) * cout * 2};
    const uint32_t ubox[3] = {64, kStagePos, 1};
    if ((rc = make_tensor_map(&oh, extra.out_u16, 3, udims, ustr, ubox, VM_SWIZZLE_NONE))) return rc;
    const uint64_t mdims[3] = {uint64_t(cout), uint64_t(p.lout), uint64_t(N)};
    const uint64_t mstr[2] = {uint64_t(cout) * 4, uint64_t(p.lout) * cout * 4};
    const uint32_t mbox[3] = {32, 8, 1};
    if ((rc = make_tensor_map(&ol, extra.out_ext, 3, mdims, mstr, mbox, VM_SWIZZLE_NONE, /*f32=*/1))) return rc;
  } else if (out_f32 != nullptr) {
    // un-pooled fp32 output (N, L, cout): TMA store boxes of 32 channels x 16 positions
    const uint64_t odims[3] = {uint64_t(cout), uint64_t(L), uint64_t(N)};
    const uint64_t ostr[2] = {uint64_t(cout) * 4, uint64_t(L) * cout * 4};
    const uint32_t obox[3] = {32, 16, 1};
    if ((rc = make_tensor_map(&oh, out_f32, 3, odims, ostr, obox, VM_SWIZZLE_NONE, /*f32=*/1))) return rc;
    ol = oh;
  } else if (gmax_partial == nullptr) {
    if (products >= 2 && out_lo == nullptr) return set_error(VM_ERR_SHAPE, "conv3: out_lo required for products>=2");
    const uint64_t odims[3] = {uint64_t(cout), uint64_t(p.lout), uint64_t(N)};
    const uint64_t ostr[2] = {uint64_t(cout) * 2, uint64_t(p.lout) * cout * 2};
    const uint32_t obox[3] = {64, kStagePos, 1};
    if ((rc = make_tensor_map(&oh, out_hi, 3, odims, ostr, obox, VM_SWIZZLE_NONE))) return rc;
    if ((rc = make_tensor_map(&ol, products >= 2 ? out_lo : out_hi, 3, odims, ostr, obox, VM_SWIZZLE_NONE)))
      return rc;
  } else {
    oh = wh;  // unused by the kernel in gmax mode
    ol = wh;
  }
  const bool two_sets = (extra.out_u16 != nullptr || out_f32 != nullptr);
  // X ring: the same bytes hold 2 two-plane slots or 4 one-plane slots.  A one-plane chunk is only 1536 MMA cycles, and
  // two of them in flight did not cover the latency of the next TMA load (dgrad with one-plane gradients ran at half
  // its MMA rate)
  const bool one_x_plane = (products == 1) || extra.x_single;
  const int xstages = one_x_plane ? kXMaxStages : kXStages;
  const int ntiles = N * p.nptile * p.nslab;
  // all cout slabs of a position tile on one CTA when its X tile fits the ring (see TileIter)
  p.slabs_per_unit = (p.nchunk <= xstages && p.nslab > 1) ? p.nslab : 1;
  const int nunits = ntiles / p.slabs_per_unit;
  int grid = max_ctas > 0 ? max_ctas : num_sms();
  if (grid > nunits) grid = nunits;
#define VM_CONV3_LAUNCH(SETS, XS)                                                                                   \
  do {                                                                                                              \
    cudaError_t ea = cudaFuncSetAttribute(conv3_kernel<SETS, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                          kSmemBytes);                                                              \
    if (ea != cudaSuccess) return set_cuda_error(ea, "conv3: cudaFuncSetAttribute");                                \
    conv3_kernel<SETS, XS><<<grid, (4 + 8 * SETS) * 32, kSmemBytes, stream>>>(xh_main, xh_halo, xl_main, xl_halo, wh, \
                                                                              wl, oh, ol, p);                       \
  } while (0)
  if (two_sets) { if (one_x_plane) VM_CONV3_LAUNCH(2, kXMaxStages); else VM_CONV3_LAUNCH(2, kXStages); }
  else { if (one_x_plane) VM_CONV3_LAUNCH(1, kXMaxStages); else VM_CONV3_LAUNCH(1, kXStages); }
#undef VM_CONV3_LAUNCH
  cudaError_t e;
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "conv3: launch");
  return VM_OK;
}

}  // namespace vm
