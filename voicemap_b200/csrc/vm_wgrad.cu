// Weight gradients of the convolutions (training backward).
//
// wgrad3 (blocks 2-4):  dW[tap][ci][co] = sum_{n,p} X[n][p + tap - 1][ci] * dU[n][p][co]
//   A GEMM whose reduction dimension is the POSITION axis, so both operands are "MN-major" for the tensor core:
//   a shared-memory tile [position rows][64 channels = 128 B] as TMA writes it (SWIZZLE_128B) is read by tcgen05
//   with a_major = b_major = MN (rows are the K dimension, 8-row groups 1024 B apart, the two 64-channel atoms
//   of a 128-wide operand LBO apart).  The three taps are, again, the same X tile read through descriptors shifted
//   by `tap` rows.  D[tap] = 128 ci lanes x 128 co columns fp32 in TMEM (3 x 128 columns), accumulated over the
//   CTA's slice of (clip, 64-position chunk) steps, then written to a per-split partial buffer that
//   wgrad_reduce sums deterministically (no atomics).
//   Both operands are fp16 planes: X is read from the very planes the forward convolution consumed (no copy in
//   another format), dU arrives scaled by the block's power-of-two gradient scale (vm_common.cuh) which
//   wgrad_reduce takes out again.  products = 3: Xh*Uh + Xl*Uh + Xh*Ul (two-plane gradient, ~2^-21);
//   products = 2: Xh*Uh + Xl*Uh (one-plane gradient: dU rounded to fp16, 2^-12 relative per element, unbiased);
//   products = 1: Xh*Uh.
//
// wgrad1 (block 1):  dW1[k][co] = sum_{n,p} x[n][p + k - 15] * dU1[n][p][co]   (tensor cores too: vm_conv1.cu)
#include "vm_common.cuh"
#include "vm_kernels.h"

namespace vm {

namespace wg {
constexpr int kPosChunk = 64;                    // positions per pipeline stage (4 MMA K-steps of 16)
constexpr int kXRows = 72;                       // 66 halo rows (p0-1 .. p0+64) padded to a multiple of 8
constexpr int kXHalfBytes = kXRows * 128;        // one 64-channel atom of the X tile
constexpr int kUHalfBytes = kPosChunk * 128;     // one 64-channel atom of the dU tile
constexpr int kXPlaneBytes = 2 * kXHalfBytes;    // 128 ci
constexpr int kUPlaneBytes = 2 * kUHalfBytes;    // 128 co
constexpr int kStageBytes = 2 * kXPlaneBytes + 2 * kUPlaneBytes;  // hi + lo of both operands = 69632
constexpr int kStages = 3;
constexpr int kThreads = 192;                    // warp 0 TMA producer, warp 1 MMA, warps 2-5 epilogue
constexpr int kTmemCols = 512;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 128;
}  // namespace wg

struct __align__(8) WgradBarriers {
  uint64_t full[wg::kStages], empty[wg::kStages];
  uint64_t done;
  uint32_t tmem_base;
};

// instruction descriptor with both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int M, int N) {
  return make_idesc_f16(M, N) | (1u << 15) | (1u << 16);   // both operands fp16
}

__global__ void __launch_bounds__(wg::kThreads, 1)
wgrad3_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
              const __grid_constant__ CUtensorMap tm_uh, const __grid_constant__ CUtensorMap tm_ul,
              const Wgrad3Params p) {
  using namespace wg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WgradBarriers* bars = reinterpret_cast<WgradBarriers*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // products 3: Xh*Uh + Xl*Uh + Xh*Ul;  2: Xh*Uh + Xl*Uh (one-plane gradient);  1: Xh*Uh
  const int xplanes = (p.products >= 2) ? 2 : 1;
  const int uplanes = (p.products == 3) ? 2 : 1;

  // work item: (ci slab, co tile, split)
  const int combo = blockIdx.x % p.ncombo;
  const int split = blockIdx.x / p.ncombo;
  const int ci0 = (combo / p.nco_tiles) * 128;
  const int co0 = (combo % p.nco_tiles) * 128;
  const int steps_total = p.N * p.nchunk;                 // (clip, 64-position chunk) pairs
  const int per = (steps_total + p.nsplit - 1) / p.nsplit;
  const int s0 = split * per, s1 = min(steps_total, s0 + per);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_xh); tma_prefetch_desc(&tm_uh);
      uint32_t it = 0;
      for (int st = s0; st < s1; ++st, ++it) {
        const int n = st / p.nchunk, p0 = (st % p.nchunk) * kPosChunk;
        const int s = it % kStages;
        mbar_wait(&bars->empty[s], ((it / kStages) & 1) ^ 1);
        uint8_t* base = smem + s * kStageBytes;
        mbar_arrive_expect_tx(&bars->full[s], xplanes * kXPlaneBytes + uplanes * kUPlaneBytes);
        for (int pl = 0; pl < xplanes; ++pl) {
          uint8_t* xb = base + pl * kXPlaneBytes;
          const CUtensorMap* mx = pl ? &tm_xl : &tm_xh;
          tma_load_3d(xb, mx, &bars->full[s], ci0, p0 - 1, n);
          tma_load_3d(xb + kXHalfBytes, mx, &bars->full[s], ci0 + 64, p0 - 1, n);
        }
        for (int pl = 0; pl < uplanes; ++pl) {
          uint8_t* ub = base + 2 * kXPlaneBytes + pl * kUPlaneBytes;
          const CUtensorMap* mu = pl ? &tm_ul : &tm_uh;
          tma_load_3d(ub, mu, &bars->full[s], co0, p0, n);
          tma_load_3d(ub + kUHalfBytes, mu, &bars->full[s], co0 + 64, p0, n);
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool elected = elect_one_sync();   // whole warp runs the loop (uniform descriptor math), one lane issues
      const uint32_t idesc = make_idesc_f16_mn(128, 128);
      uint32_t it = 0;
      for (int st = s0; st < s1; ++st, ++it) {
        const int s = it % kStages;
        mbar_wait(&bars->full[s], (it / kStages) & 1);
        tc_fence_after_sync();
        const uint32_t base = smem_u32(smem + s * kStageBytes);
        const uint32_t xh = base, xl = base + kXPlaneBytes;
        const uint32_t uh = base + 2 * kXPlaneBytes, ul = uh + kUPlaneBytes;
        for (int tap = 0; tap < 3; ++tap) {
          const uint32_t d = tmem_base + tap * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t arow = (tap + 16 * kk) * 128, brow = (16 * kk) * 128;
            const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
            if (elected) umma_f16(d, make_smem_desc(xh + arow, kXHalfBytes, 1024, kLayoutSW128),
                     make_smem_desc(uh + brow, kUHalfBytes, 1024, kLayoutSW128), idesc, acc);
            if (xplanes == 2 && elected)
              umma_f16(d, make_smem_desc(xl + arow, kXHalfBytes, 1024, kLayoutSW128),
                       make_smem_desc(uh + brow, kUHalfBytes, 1024, kLayoutSW128), idesc, 1);
            if (uplanes == 2 && elected)
              umma_f16(d, make_smem_desc(xh + arow, kXHalfBytes, 1024, kLayoutSW128),
                       make_smem_desc(ul + brow, kUHalfBytes, 1024, kLayoutSW128), idesc, 1);
          }
        }
        if (elected) umma_commit(&bars->empty[s]);
      }
      if (elected) umma_commit(&bars->done);
    }
  } else {
    // epilogue: TMEM lane = ci row; write partial[split][tap][ci][co]
    const int q = warp & 3;
    mbar_wait(&bars->done, 0);
    tc_fence_after_sync();
    const int ci = ci0 + q * 32 + lane;
    const bool any = s1 > s0;
    for (int tap = 0; tap < 3; ++tap) {
      float* row = p.partial + ((size_t(split) * 3 + tap) * p.cin + ci) * p.cout + co0;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        float v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + tap * 128 + g * 32, v);
        if (ci < p.cin) {
          // a thread owns 32 consecutive floats of its row (one 128-byte line): 16-byte stores instead of 32 scalar
          // ones, which each touched a different sector per lane (cout % 8 == 0, so a float4 is all in or all out)
          float4* r4 = reinterpret_cast<float4*>(row + g * 32);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (co0 + g * 32 + 4 * k < p.cout)
              r4[k] = any ? make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3])
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// dW[i] = sum_s partial[s][i] / (the block's power-of-two gradient scale, vm_common.cuh)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, size_t n,
                                    const unsigned int* __restrict__ grad_absmax, float* __restrict__ out) {
  const float unscale = grad_absmax ? 1.0f / grad_scale_from_absmax(__uint_as_float(*grad_absmax)) : 1.0f;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[size_t(k) * n + i];
    out[i] = s * unscale;
  }
}

// The same for MANY splits of a small gradient (block 1: one partial per CTA, 148 x 4096 values): 32 outputs per block,
// eight threads per output take every eighth split with four loads in flight, fixed summation order.  (One thread per
// output walked 148 dependent-latency loads: 20 us for 2.4 MB.)
__global__ void __launch_bounds__(256)
wgrad_reduce_tall_kernel(const float* __restrict__ partial, int nsplit, size_t n,
                         const unsigned int* __restrict__ grad_absmax, float* __restrict__ out) {
  __shared__ float sm[8][32];
  const size_t i = size_t(blockIdx.x) * 32 + threadIdx.x;
  float s = 0.f;
  if (i < n) {
    for (int k = threadIdx.y; k < nsplit; k += 32) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (k + 8 * j < nsplit) ? partial[size_t(k + 8 * j) * n + i] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += v[j];
    }
  }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    for (int k = 1; k < 8; ++k) s += sm[k][threadIdx.x];
    const float unscale = grad_absmax ? 1.0f / grad_scale_from_absmax(__uint_as_float(*grad_absmax)) : 1.0f;
    out[i] = s * unscale;
  }
}

int launch_wgrad3(const __half* x_hi, const __half* x_lo, const __half* du_hi, const __half* du_lo, int N, int L,
                  int cin, int cout, int products, float* partial, size_t partial_bytes, float* dw,
                  const unsigned int* grad_absmax, cudaStream_t stream) {
  using namespace wg;
  if (N <= 0 || L <= 0 || cin % 8 != 0 || cout % 8 != 0) return set_error(VM_ERR_SHAPE, "wgrad3: bad shape");
  if (products < 1 || products > 3) return set_error(VM_ERR_SHAPE, "wgrad3: products must be 1, 2 or 3");
  if ((products >= 2 && x_lo == nullptr) || (products == 3 && du_lo == nullptr))
    return set_error(VM_ERR_SHAPE, "wgrad3: lo planes required");
  Wgrad3Params p{};
  p.N = N; p.L = L; p.cin = cin; p.cout = cout; p.products = products;
  p.nchunk = (L + kPosChunk - 1) / kPosChunk;
  p.nco_tiles = (cout + 127) / 128;
  p.ncombo = ((cin + 127) / 128) * p.nco_tiles;
  const size_t wsize = size_t(3) * cin * cout;
  // ONE wave of CTAs (a CTA takes an SM's whole shared memory).  Two waves (the first version) paid the prologue, the
  // 196 KB partial tile and its share of the reduce twice per SM, and as a side-stream kernel one wave shares the GPU
  // better with the main stream: 64-pair step 2.57 -> 2.44 ms, 16-pair 0.99 -> 0.93 ms; three waves 2.64 ms
  // (profiles/r02_wgrad_waves_ab.log)
  int nsplit = max(1, num_sms() / p.ncombo);
  nsplit = min(nsplit, N * p.nchunk);
  while (nsplit > 1 && size_t(nsplit) * wsize * 4 > partial_bytes) --nsplit;
  if (size_t(nsplit) * wsize * 4 > partial_bytes) return set_error(VM_ERR_SHAPE, "wgrad3: partial buffer too small");
  p.nsplit = nsplit;
  p.partial = partial;

  CUtensorMap xh, xl, uh, ul;
  const uint64_t xdims[3] = {uint64_t(cin), uint64_t(L), uint64_t(N)};
  const uint64_t xstr[2] = {uint64_t(cin) * 2, uint64_t(L) * cin * 2};
  const uint32_t xbox[3] = {64, kXRows, 1};
  const uint64_t udims[3] = {uint64_t(cout), uint64_t(L), uint64_t(N)};
  const uint64_t ustr[2] = {uint64_t(cout) * 2, uint64_t(L) * cout * 2};
  const uint32_t ubox[3] = {64, kPosChunk, 1};
  int rc;
  if ((rc = make_tensor_map(&xh, x_hi, 3, xdims, xstr, xbox, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&xl, products >= 2 ? x_lo : x_hi, 3, xdims, xstr, xbox, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&uh, du_hi, 3, udims, ustr, ubox, VM_SWIZZLE_128B))) return rc;
  if ((rc = make_tensor_map(&ul, products == 3 ? du_lo : du_hi, 3, udims, ustr, ubox, VM_SWIZZLE_128B))) return rc;

  cudaError_t e = cudaFuncSetAttribute(wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad3: cudaFuncSetAttribute");
  wgrad3_kernel<<<p.ncombo * nsplit, kThreads, kSmemBytes, stream>>>(xh, xl, uh, ul, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad3: launch");
  const unsigned blocks = unsigned(min(size_t(148 * 8), (wsize + 255) / 256));
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(partial, nsplit, wsize, grad_absmax, dw);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad3: reduce launch");
  return VM_OK;
}

int launch_wgrad1(const float* x, const __half* du_hi, const __half* du_lo, int N, int L, int cout, float* partial,
                  size_t partial_bytes, float* dw, const unsigned int* grad_absmax, cudaStream_t stream, int products) {
  if (N <= 0 || L <= 0 || cout <= 0) return set_error(VM_ERR_SHAPE, "wgrad1: bad shape");
  if (products < 1 || products > 3) return set_error(VM_ERR_SHAPE, "wgrad1: products must be 1, 2 or 3");
  int nsplit = 0;
  int rc = launch_wgrad1_tc(x, du_hi, du_lo, N, L, cout, products, partial, partial_bytes, &nsplit, stream);
  if (rc) return rc;
  const size_t wsz = size_t(32) * cout;
  if (nsplit >= 32)
    wgrad_reduce_tall_kernel<<<unsigned((wsz + 31) / 32), dim3(32, 8), 0, stream>>>(partial, nsplit, wsz, grad_absmax, dw);
  else
    wgrad_reduce_kernel<<<unsigned((wsz + 255) / 256), 256, 0, stream>>>(partial, nsplit, wsz, grad_absmax, dw);
  cudaError_t e2 = cudaGetLastError();
  if (e2 != cudaSuccess) return set_cuda_error(e2, "wgrad1: reduce launch");
  return VM_OK;
}

}  // namespace vm
