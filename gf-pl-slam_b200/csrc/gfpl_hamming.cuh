// gfpl_hamming.cuh — 256-bit Hamming distance primitives and the TMA-bulk smem staging used by every
// matching kernel (sm_100a).
//
// Descriptor rows are 32 B = two 128-bit words.  A query row lives in 8 registers; train rows are
// staged in shared memory with 1-D bulk TMA (cp.async.bulk.shared::cluster.global + mbarrier
// complete_tx) and read back with two broadcast LDS.128 per row (all lanes of a warp read the same
// address -> one wavefront, no bank conflict).  The distance is 8 x LOP3 xor, a 3-CSA carry-save compression
// and 5 POPC: the POPC pipe (16 lanes/clk/SM) is the roofline of these kernels, not HBM and not the tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gfpl_c.h"

namespace gfpl {

struct Desc256 {
  uint32_t w[8];
};

__device__ __forceinline__ Desc256 load_desc_global(const uint8_t* row) {
  const uint4* p = reinterpret_cast<const uint4*>(row);
  uint4 a = __ldg(p), b = __ldg(p + 1);
  Desc256 d;
  d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
  d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
  return d;
}

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// NORM_HAMMING: popcount of the xor.  NORM_HAMMING2: number of non-zero aligned 2-bit groups.
template <int NORM>
__device__ __forceinline__ uint32_t hamming256(const Desc256& q, const uint4& a, const uint4& b) {
  uint32_t x0 = q.w[0] ^ a.x, x1 = q.w[1] ^ a.y, x2 = q.w[2] ^ a.z, x3 = q.w[3] ^ a.w;
  uint32_t x4 = q.w[4] ^ b.x, x5 = q.w[5] ^ b.y, x6 = q.w[6] ^ b.z, x7 = q.w[7] ^ b.w;
  if (NORM == GFPL_NORM_HAMMING2) {
    x0 = (x0 | (x0 >> 1)) & 0x55555555u; x1 = (x1 | (x1 >> 1)) & 0x55555555u;
    x2 = (x2 | (x2 >> 1)) & 0x55555555u; x3 = (x3 | (x3 >> 1)) & 0x55555555u;
    x4 = (x4 | (x4 >> 1)) & 0x55555555u; x5 = (x5 | (x5 >> 1)) & 0x55555555u;
    x6 = (x6 | (x6 >> 1)) & 0x55555555u; x7 = (x7 | (x7 >> 1)) & 0x55555555u;
  }
  if (NORM == GFPL_NORM_HAMMING2)  // already ALU-pipe bound (8 xor + 16 fold ops): plain 8 popcounts
    return (__popc(x0) + __popc(x1) + __popc(x2)) + (__popc(x3) + __popc(x4) + __popc(x5)) + (__popc(x6) + __popc(x7));
  // Carry-save compression (exact): popc(a) + popc(b) + popc(c) = popc(a^b^c) + 2*popc(maj(a,b,c)).  Three CSAs
  // (6 LOP3 on the ALU pipe, which has slack) turn the 8 popcounts into 5, so the 16-lane POPC pipe — the roofline
  // of the naive loop — does 5/8 of the work: ~40 POPC-pipe cycles and ~40 ALU-pipe cycles per warp-row, balanced.
  // (explicit lop3 so that the 8 xors stay materialised: 8 + 6 LOP3 per pair; left to itself the compiler
  //  re-associates the xors into the adders and ends up with ~21)
  const uint32_t s0 = lop3_xor3(x0, x1, x2), c0 = lop3_maj(x0, x1, x2);
  const uint32_t s1 = lop3_xor3(x3, x4, x5), c1 = lop3_maj(x3, x4, x5);
  const uint32_t s2 = lop3_xor3(s0, s1, x6), c2 = lop3_maj(s0, s1, x6);
  return (__popc(s2) + __popc(x7)) + 2u * (__popc(c0) + __popc(c1) + __popc(c2));
}

// (distance, trainIdx) packed so that unsigned order == (distance asc, trainIdx asc): the tie rule of
// cv::BFMatcher::knnMatch.  distance <= 256 needs 9 bits, index gets 20.
constexpr int KEY_IDX_BITS = 20;
constexpr uint32_t KEY_IDX_MASK = (1u << KEY_IDX_BITS) - 1;
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t make_key(uint32_t d, uint32_t idx) { return (d << KEY_IDX_BITS) | idx; }
__device__ __forceinline__ void top2_update(uint32_t& k0, uint32_t& k1, uint32_t key) {
  uint32_t m = max(k0, key);
  k0 = min(k0, key);
  k1 = min(k1, m);
}

// Problems of the "both sides <= 128 rows" class that the small tensor-core kernel leaves to the integer kernel: empty
// ones and those too small to be worth a product round trip.
constexpr int TC_TINY = 40;
__device__ __forceinline__ bool tc_small_skips(int nq, int nt) { return (nq <= TC_TINY && nt <= TC_TINY) || nq == 0 || nt == 0; }

// ---- mbarrier + 1-D bulk TMA ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy, completion counted in bytes on `bar`.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Two-stage ring of train tiles.  One elected thread issues; every thread of the CTA consumes.
template <int TILE_ROWS>
struct TrainPipe {
  static constexpr int TILE_BYTES = TILE_ROWS * GFPL_DESC_BYTES;
  uint8_t* buf;    // [2][TILE_BYTES], 128-byte aligned
  uint64_t* bars;  // [2]
  const uint8_t* src;
  int nrows, ntiles;

  __device__ __forceinline__ void init(uint8_t* smem_buf, uint64_t* smem_bars, const uint8_t* train, int nt) {
    buf = smem_buf; bars = smem_bars; src = train; nrows = nt;
    ntiles = (nt + TILE_ROWS - 1) / TILE_ROWS;
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (ntiles > 0) issue(0);
      if (ntiles > 1) issue(1);
    }
  }
  __device__ __forceinline__ int rows_in(int tile) const { return min(TILE_ROWS, nrows - tile * TILE_ROWS); }
  __device__ __forceinline__ void issue(int tile) {
    uint32_t bytes = (uint32_t)rows_in(tile) * GFPL_DESC_BYTES;
    mbar_expect_tx(&bars[tile & 1], bytes);
    tma_load_1d(buf + (tile & 1) * TILE_BYTES, src + (size_t)tile * TILE_BYTES, bytes, &bars[tile & 1]);
  }
  // wait until tile `tile` has landed; returns its smem base
  __device__ __forceinline__ const uint8_t* acquire(int tile) {
    mbar_wait(&bars[tile & 1], (tile >> 1) & 1);
    return buf + (tile & 1) * TILE_BYTES;
  }
  // every thread is done with tile `tile`: refill its slot with tile+2
  __device__ __forceinline__ void release(int tile) {
    __syncthreads();
    if (threadIdx.x == 0 && tile + 2 < ntiles) issue(tile + 2);
  }
};

}  // namespace gfpl
