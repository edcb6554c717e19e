// Cross-rank sum of a small vector (the per-(group, channel) BatchNorm sums, <= 64 KB) over NVLink peer memory, inside
// the reduction kernel that produces the local sums -- no NCCL call, no host involvement, no extra launch.
//
// Every rank owns one exchange buffer that all ranks of the node have mapped (CUDA IPC):
//     flags [2 slots][kP2PMaxColumns][kP2PMaxRanks] uint32   (offset kP2PColFlagOffset)
//     data  [2 slots][kP2PMaxRanks][kP2PMaxDoubles]          (offset kP2PDataOffset)
// The unit of exchange is a COLUMN of 32 channels (one warp: lane = channel): the block that finishes a column of the
// two-stage reduction exchanges that column while the other columns are still being summed.
// A call with sequence number seq (the same on every rank, starting at 1) uses slot seq & 1; per column:
//   1. PUSH: rank r stores the column's sums into data[slot][r] of EVERY rank's buffer (remote stores over NVLink),
//      fences (system scope) and then stores seq into flags[slot][column][r] of every rank's buffer (release);
//   2. WAIT: it spins on its OWN buffer until flags[slot][column][q] == seq for all q (acquire) -- local polling;
//   3. SUM:  it adds data[slot][0 .. W-1] of its own buffer in rank order: the same order on every rank, so all ranks
//      hold bit-identical sums (BatchNorm constants must agree exactly or the replicas drift apart).
// Two slots are enough: a rank can only be one call ahead of the slowest one (a column of call k+1 cannot complete
// before every rank has raised that column's flag for k+1, which it does after its kernel of call k has finished).
// The wait is bounded: a missing peer traps the kernel (an error) instead of hanging the GPU.
#pragma once
#include <stdint.h>

namespace vm {

constexpr int kP2PMaxRanks = 8;
constexpr int kP2PMaxDoubles = 8192;          // 2 sums x 2 groups x 2048 channels
constexpr int kP2PMaxColumns = 64;            // 32-channel columns, C <= 2048
constexpr size_t kP2PColFlagOffset = 256;     // flags [2 slots][kP2PMaxColumns][kP2PMaxRanks] uint32 = 4 KB
constexpr size_t kP2PDataOffset = kP2PColFlagOffset + size_t(2) * kP2PMaxColumns * kP2PMaxRanks * 4;
constexpr size_t kP2PBufferBytes = kP2PDataOffset + size_t(2) * kP2PMaxRanks * kP2PMaxDoubles * sizeof(double);

struct P2PPeers {
  void* buf[kP2PMaxRanks];   // every rank's exchange buffer as mapped into this process (buf[rank] is the local one)
  int rank, world;
};

__device__ __forceinline__ double* p2p_data(void* base, int slot, int r) {
  return reinterpret_cast<double*>(static_cast<char*>(base) + kP2PDataOffset) +
         (size_t(slot) * kP2PMaxRanks + r) * kP2PMaxDoubles;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ONE WARP exchanges the G double2 sums of its 32 channels (lane = channel c of column `col`).  Vector layout: double2
// index g*C + c.
__device__ __forceinline__ unsigned int* p2p_col_flag(void* base, int slot, int col, int r) {
  return reinterpret_cast<unsigned int*>(static_cast<char*>(base) + kP2PColFlagOffset) +
         (size_t(slot) * kP2PMaxColumns + col) * kP2PMaxRanks + r;
}
// mine[g] (this rank's sums of channel c, c < C only) -> returns with total[g*C + c] written (sum over ranks in rank
// order) for every valid lane.  All 32 lanes of the warp must call.
template <int kMaxG>
__device__ __forceinline__ void p2p_allreduce_column(const double2 (&mine)[kMaxG], int G, int C, int c, int col,
                                                     const P2PPeers& peers, unsigned int seq,
                                                     double2* __restrict__ total) {
  const int slot = seq & 1, lane = threadIdx.x & 31;
  if (c < C) {
    for (int q = 0; q < peers.world; ++q) {
      double2* dst = reinterpret_cast<double2*>(p2p_data(peers.buf[q], slot, peers.rank));
      for (int g = 0; g < G; ++g) dst[size_t(g) * C + c] = mine[g];
    }
  }
  __threadfence_system();
  __syncwarp();
  if (lane < peers.world) {
    st_release_sys(p2p_col_flag(peers.buf[lane], slot, col, peers.rank), seq);
    const unsigned int* flag = p2p_col_flag(peers.buf[peers.rank], slot, col, lane);
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != seq) {
      if (clock64() - t0 > 8000000000LL) {   // ~4 s: a peer never arrived
        printf("vm: p2p column exchange timeout rank %d waiting for rank %d seq %u column %d\n", peers.rank, lane, seq,
               col);
        __trap();
      }
    }
  }
  __syncwarp();
  if (c < C) {
    for (int g = 0; g < G; ++g) {
      double a = 0.0, b = 0.0;
      for (int r = 0; r < peers.world; ++r) {   // L2: written by the peers
        const double* src = p2p_data(peers.buf[peers.rank], slot, r) + 2 * (size_t(g) * C + c);
        a += __ldcg(src);
        b += __ldcg(src + 1);
      }
      total[size_t(g) * C + c] = make_double2(a, b);
    }
  }
}

}  // namespace vm
