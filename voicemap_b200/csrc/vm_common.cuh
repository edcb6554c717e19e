// voicemap_b200 -- sm_100a device primitives (mbarrier, TMA, tcgen05/TMEM) as thin inline-PTX
// wrappers.  Everything here is Blackwell-only; there is no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vm {

static constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning through the issue slots the epilogue warps need.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (true) {
#pragma unroll 1
    for (int i = 0; i < 256; ++i)
      if (mbar_try_wait(addr, parity)) return;
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("vm: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  }
}

// explicit shared-state-space stores (32-bit addresses; generic stores cost 64-bit address arithmetic)
__device__ __forceinline__ void sts_u16(uint32_t addr, __half v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// proxies / fences
// ---------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// shared memory -> global (bulk tensor store, bulk-group completion).  Out-of-bounds parts of the box are dropped.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N committed bulk groups still have to READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// named barrier among a subset of warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ---------------------------------------------------------------------------------------------
// TMEM allocation (one warp, .sync.aligned)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05.mma  (kind::f16, fp32 accumulate in TMEM), operands from shared-memory descriptors
// ---------------------------------------------------------------------------------------------
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format[4,6)=1 (F32), a_format[7,10)=0 (F16),
// b_format[10,13)=0 (F16), a_major bit15=0 (K), b_major bit16=0 (K), n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
// a_bf16 / b_bf16 select bf16 instead of fp16 for that operand (kind::f16 takes either, independently).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_bf16 = 0, int b_bf16 = 0) {
  return (1u << 4) | (uint32_t(a_bf16 & 1) << 7) | (uint32_t(b_bf16 & 1) << 10) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major operand.
//   start_address [0,14) = addr>>4, LBO [16,30) = bytes>>4, SBO [32,46) = bytes>>4, version [46,48) = 1,
//   base_offset [49,52), layout_type [61,64): 0 none, 2 SW128, 4 SW64, 6 SW32.
enum : uint32_t { kLayoutNone = 0, kLayoutSW128 = 2, kLayoutSW64 = 4, kLayoutSW32 = 6 };
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFFu);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(base_offset & 7u) << 49;
  d |= uint64_t(layout & 7u) << 61;
  return d;
}

// One lane of the (converged) warp.  The MMA issue loops run on the whole warp with only tcgen05.mma / commit under this
// predicate, so that descriptor arithmetic stays warp-uniform (uniform registers) instead of being computed under a
// divergent `if (lane == 0)` and moved into uniform registers through ELECT / R2UR waterfall loops.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4: the same shared-memory descriptors, but the 32 operand bytes per row that one instruction consumes are
// 32 fp8 K elements instead of 16 fp16 ones -- twice the MACs per instruction at the same instruction time
// (measured: an f8f6f4 MMA of M=128, N=256 issues at the rate of the f16 one, profiles/r01_f8_rate_probe.log).
// Formats (cute::UMMA::F8F6F4Format): 0 = e4m3, 1 = e5m2.
__host__ __device__ constexpr uint32_t make_idesc_f8(int M, int N, int a_fmt = 1, int b_fmt = 1) {
  return (1u << 4) | (uint32_t(a_fmt & 7) << 7) | (uint32_t(b_fmt & 7) << 10) | (uint32_t(N >> 3) << 17) |
         (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All MMAs issued so far by this thread -> one arrive on `bar` when they have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (thread i of warp w reads lane 32*(w%4)+i)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form for software pipelining: issue the load, do other work, then wait.  The wait takes the destination
// registers as read-write operands so that no use of them can be scheduled ahead of it.
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Fused bias + ReLU + BatchNorm affine on a pooled accumulator maximum m, constants {a, c, lo, hi} from vm_pack_conv*:
//   BN scale s >= 0:  max_pool(s*relu(acc + b) + t) = s*max(M, -b) + (s*b + t)           -> {s, s*b + t, -b, +inf}
//   BN scale s <  0:  weights are packed negated (M = max_pool(-acc)) and the pool becomes a min:
//                     s*min_pool(relu(acc + b)) + t = (-s)*min(M, b) + (s*b + t)          -> {-s, s*b + t, -inf, b}
// i.e. y = a*clamp(M, lo, hi) + c.  kClampHi = false is the specialisation for "all scales >= 0" (hi = +inf).
template <bool kClampHi = true>
__device__ __forceinline__ float apply_epi(const float4& ep, float m) {
  float v = fmaxf(m, ep.z);
  if (kClampHi) v = fminf(v, ep.w);
  return fmaf(ep.x, v, ep.y);
}
// pooled variants: the lower clamp rides in the last 3-input max of the pool
template <bool kClampHi = true>
__device__ __forceinline__ float apply_epi_pool2(const float4& ep, float v0, float v1) {
  float v = max3(v0, v1, ep.z);
  if (kClampHi) v = fminf(v, ep.w);
  return fmaf(ep.x, v, ep.y);
}
template <bool kClampHi = true>
__device__ __forceinline__ float apply_epi_pool4(const float4& ep, float v0, float v1, float v2, float v3) {
  float v = max3(max3(v0, v1, v2), v3, ep.z);
  if (kClampHi) v = fminf(v, ep.w);
  return fmaf(ep.x, v, ep.y);
}
// true when no lane of the warp needs the upper clamp
__device__ __forceinline__ bool epi_no_upper_clamp(const float4& ep) {
  return __all_sync(0xffffffffu, ep.w == INFINITY);
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// split form of the 16-column load (see tmem_ld_32x32_issue / tmem_ld_wait)
__device__ __forceinline__ void tmem_ld_32x16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])
               :
               : "memory");
}
// split form of the 8-column load
__device__ __forceinline__ void tmem_ld_32x8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) fp16 split.  hi = rn16(x), lo = rn16(x - hi).  hi+lo carries ~22 significant
// bits (absolute floor 2^-25 from the fp16 subnormal grid); products of hi/lo pairs are exact in fp32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_f32(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// bf16 (hi, lo) split for gradients: fp32 exponent range (no loss scaling needed), ~16 significant bits.
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
// ---------------------------------------------------------------------------------------------
// "Q" planes (precision 2): the two correction products of the split scheme, Xl*Wh + Xh*Wl, only need a few bits, so
// they run as ONE kind::f8f6f4 product over a doubled K: every fp32 value carries, next to its fp16 `hi`, a 16-bit
// pair of e5m2 numbers {upper, lower} and the MMA contracts lower*lower + upper*upper over the byte pairs:
//   activations: lower = e5m2((x - hi) * 2^6)   upper = e5m2(x * 2^-6)
//   weights:     lower = e5m2(w * 2^-6)          upper = e5m2((w - hi) * 2^6)
// so that lower*lower ~ Xl*W and upper*upper ~ X*Wl land in the accumulator at scale 1 (e5m2 has fp16's exponent
// range, so the residuals 2^-12 below their values stay normal; e4m3 would need a second accumulator).  Error of the
// corrections ~2^-3.5 of a 2^-12 term: embeddings 1e-5 .. 4e-5 from the fp64 oracle (tools/sim_split_precision.py).
// ---------------------------------------------------------------------------------------------
static constexpr float kQUp = 64.f, kQDown = 1.f / 64.f;
__device__ __forceinline__ uint16_t pack_e5m2x2(float upper, float lower) {
  uint16_t d;
  asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(d) : "f"(upper), "f"(lower));
  return d;
}
__device__ __forceinline__ void split_f16_q(float x, __half& hi, uint16_t& q) {
  hi = __float2half_rn(x);
  q = pack_e5m2x2(x * kQDown, (x - __half2float(hi)) * kQUp);
}
__device__ __forceinline__ void split_w_q(float w, __half& hi, uint16_t& q) {
  hi = __float2half_rn(w);
  q = pack_e5m2x2((w - __half2float(hi)) * kQUp, w * kQDown);
}
// decode of an activation pair: hi + lower / 2^6 (tests / plane merging)
__device__ __forceinline__ float e5m2_to_float(uint32_t byte) { return __half2float(__ushort_as_half(uint16_t(byte << 8))); }
__device__ __forceinline__ void sts_b16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(uint32_t(b) << 16); }

// ---------------------------------------------------------------------------------------------
// Training: the un-pooled activation u = relu(conv + bias) is kept for the backward pass as ONE 16-bit word per
// element: fp16(u) in bits 0-14 (u >= 0, so the sign bit is free) and, in bit 15, "this element is the arg-max of its
// MaxPool window" (first winner on ties; for channels with a negative BatchNorm scale the arg-MIN of u, which is the
// arg-max of the normalised value).  The forward pass itself continues from the fp32 window extremes (see
// vm_train.cu), so the rounding to fp16 only enters the backward pass through xhat = (u - mean) * rstd.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t encode_u(float u, bool is_argmax) {
  return uint16_t(__half_as_ushort(__float2half_rn(u)) | (is_argmax ? 0x8000u : 0u));
}
__device__ __forceinline__ float decode_u(uint32_t bits) { return __half2float(__ushort_as_half(uint16_t(bits & 0x7FFFu))); }

// Gradients travel between the backward kernels as fp16 planes scaled by a power of two chosen per block from the
// largest |s * dy| the BatchNorm-backward reduction saw (a float's bits, kept in a device word and raised with
// atomicMax: non-negative floats order like unsigned integers).  Every kernel that writes or reads the scaled planes
// derives the scale from that word with this function, so producer and consumers agree exactly; powers of two make the
// scaling itself exact.  Target: the largest routed gradient lands in [32, 64) -- 2^10 of head room below fp16's
// maximum for the batch-statistics terms, 2^19 of normal range below.
__host__ __device__ __forceinline__ float grad_scale_from_absmax(float a) {
  if (!(a > 0.f) || !(a < 3.0e38f)) return 1.f;
  int e;
  frexpf(a, &e);                 // a = m * 2^e, m in [0.5, 1)
  e = 6 - e;
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return ldexpf(1.f, e);
}

}  // namespace vm
