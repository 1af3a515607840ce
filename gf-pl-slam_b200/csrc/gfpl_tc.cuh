// gfpl_tc.cuh — shared device code of the tensor-core Hamming kernels (k_hamming_tc.cu, k_f2f.cu): descriptor expansion
// to int8 +-1 in the UMMA shared-memory layout, tcgen05.mma issue, TMEM read-back and the thread-local top-2 selection.
#pragma once
#include "gfpl_hamming.cuh"

namespace gfpl {

constexpr int TC_M = 128;            // query rows per tile = TMEM lanes
constexpr int TC_THREADS = 256;      // warps w and w + 4 share the TMEM lanes 32 (w % 4) .. +31 and split the columns
constexpr int TC_NPASS = 128;        // train rows per accumulator buffer (TMEM columns)
constexpr int TC_TMEM_COLS = 256;    // two accumulator buffers; two CTAs share the SM's 512 columns
constexpr int TC_ROW_BYTES = 256;    // one expanded descriptor
constexpr int TC_LBO = 128;          // bytes between the 16-byte K chunks of a core-matrix row group
constexpr int TC_SBO = 16 * TC_LBO;  // bytes between 8-row groups (16 K chunks of 128 B each)

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48),
  // layout type SWIZZLE_NONE = 0 [61,64)
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ uint32_t umma_idesc_i8(int n) {
  // InstrDescriptor: c_format S32 = 2 [4,6), a/b format INT8 signed = 1 [7,10) [10,13), K-major A and B, N >> 3 [17,23),
  // M >> 4 [24,29)
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

// Descriptor bits -> int8 of magnitude 64.  Train side (n): bit set -> +64 (0x40), clear -> -64 (0xC0); query side (m): the
// negation.  A product of the two is then  -4096 * (256 - 2 * hamming) = 8192 * hamming - 2^20 : the accumulator already is
// the sort key's distance field (hamming << 13), and the selection only adds the column index.
constexpr int TC_KEY_SHIFT = 13;                 // accumulator = (hamming << 13) - 2^20; train rows per problem < 8192
constexpr int TC_KEY_BIAS = 1 << 20;
constexpr uint32_t TC_PAIR_TRAIN = 0x000040C0u;  // byte 0 = the int8 of a clear bit, byte 1 = of a set bit
constexpr uint32_t TC_PAIR_QUERY = 0x0000C040u;  // negated

// One descriptor word (32 bits) -> 32 int8 = two 16-byte K chunks, with byte permutes instead of table look-ups:
// (w >> s) & 0x11111111 holds bits s, s + 4, ..., s + 28 in its eight nibbles, and a nibble that is 0 or 1 is exactly a
// PRMT selector for byte 0 or 1 of `pair`.  Within a word the K positions come out permuted (chunk lo: bits
// 0,4,8,12 | 1,5,9,13 | 2,6,10,14 | 3,7,11,15; chunk hi: the same + 16) — the same permutation for every row of both
// operands, which a dot product does not see.  19 ALU instructions per word against 8 shared-memory look-ups + ~24.
__device__ __forceinline__ void expand_word(uint32_t w, uint32_t pair, uint4& lo, uint4& hi) {
  const uint32_t t0 = w & 0x11111111u, t1 = (w >> 1) & 0x11111111u, t2 = (w >> 2) & 0x11111111u, t3 = (w >> 3) & 0x11111111u;
  lo = make_uint4(__byte_perm(pair, 0u, t0), __byte_perm(pair, 0u, t1), __byte_perm(pair, 0u, t2), __byte_perm(pair, 0u, t3));
  hi = make_uint4(__byte_perm(pair, 0u, t0 >> 16), __byte_perm(pair, 0u, t1 >> 16), __byte_perm(pair, 0u, t2 >> 16),
                  __byte_perm(pair, 0u, t3 >> 16));
}

// one descriptor row (8 words) -> 16 chunks of 16 int8 in the UMMA layout.  Rows past the end of a side are not written
// at all: as query rows they only produce accumulator lanes nobody reads, as train rows columns the selection masks.
__device__ __forceinline__ void expand_row(uint8_t* tile, int r, const uint4& lo, const uint4& hi, uint32_t pair) {
  uint8_t* row = tile + (r >> 3) * TC_SBO + (r & 7) * 16;
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int c = 0; c < 8; c++) {
    uint4 a, b;
    expand_word(w[c], pair, a, b);
    *reinterpret_cast<uint4*>(row + (2 * c) * TC_LBO) = a;
    *reinterpret_cast<uint4*>(row + (2 * c + 1) * TC_LBO) = b;
  }
}
// half a row (16 of its 32 bytes -> 8 chunks): two threads share one row of the 128-row tile
__device__ __forceinline__ void expand_half_row(uint8_t* tile, int r, int h, const uint4& q, uint32_t pair) {
  uint8_t* row = tile + (r >> 3) * TC_SBO + (r & 7) * 16 + h * 8 * TC_LBO;
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int c = 0; c < 4; c++) {
    uint4 a, b;
    expand_word(w[c], pair, a, b);
    *reinterpret_cast<uint4*>(row + (2 * c) * TC_LBO) = a;
    *reinterpret_cast<uint4*>(row + (2 * c + 1) * TC_LBO) = b;
  }
}

#define TC_LD16(v, addr)                                                                                                  \
  asm volatile(                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n" \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),      \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                          \
      : "r"(addr))

// tcgen05.ld is asynchronous: its destination registers are valid only after tcgen05.wait::ld.  Tying them to the wait
// as read-write operands keeps the compiler from scheduling their uses above it.
#define TC_WAIT16(v)                                                                                                     \
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"                                                                        \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),          \
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])     \
               :                                                                                                          \
               : "memory")

// Relative signed selection keys (see tc_direction): key = accumulator + column - kbase.
constexpr int TC_KNONE = 0x3FFFFFFF;          // stays above every real key (< 2^22) however often kbase moves
struct TcKeys {
  int k0a, k1a, k0b, k1b, kbase;
  __device__ __forceinline__ void reset() { k0a = k1a = k0b = k1b = TC_KNONE; kbase = 0; }
};
__device__ __forceinline__ uint32_t tc_canonical(int k, int kbase) {   // -> distance << 20 | train row (KEY_NONE if unset)
  if (k >= (TC_KNONE >> 1)) return KEY_NONE;
  const uint32_t kabs = (uint32_t)(k + kbase + TC_KEY_BIAS);   // hamming << 13 | train row
  return ((kabs >> TC_KEY_SHIFT) << KEY_IDX_BITS) | (kabs & ((1u << TC_KEY_SHIFT) - 1u));
}

// 16 accumulator columns [c0, c0 + 16) of one row -> the running keys.  `col0` = index of column 0 of the product in the
// train side; FULL: all 16 columns are real rows.
template <bool TOP2, bool FULL>
__device__ __forceinline__ void tc_reduce16(TcKeys& K, const uint32_t* v, int col0, int c0, int nvalid) {
  const int shift = (col0 + c0) - K.kbase;
  K.kbase = col0 + c0;
  K.k0a -= shift; K.k0b -= shift;
  if (TOP2) { K.k1a -= shift; K.k1b -= shift; }
  if (FULL || c0 + 16 <= nvalid) {
#pragma unroll
    for (int jx = 0; jx < 16; jx += 4) {   // DPX: fused add-min / add-max and three-input minimum
      const int a0 = (int)v[jx], a1 = (int)v[jx + 1], a2 = (int)v[jx + 2], a3 = (int)v[jx + 3];
      if (TOP2) {
        const int ta0 = __viaddmax_s32(a0, jx, K.k0a);      K.k0a = __viaddmin_s32(a0, jx, K.k0a);
        const int tb0 = __viaddmax_s32(a1, jx + 1, K.k0b);  K.k0b = __viaddmin_s32(a1, jx + 1, K.k0b);
        const int ta1 = __viaddmax_s32(a2, jx + 2, K.k0a);  K.k0a = __viaddmin_s32(a2, jx + 2, K.k0a);
        const int tb1 = __viaddmax_s32(a3, jx + 3, K.k0b);  K.k0b = __viaddmin_s32(a3, jx + 3, K.k0b);
        K.k1a = __vimin3_s32(K.k1a, ta0, ta1);
        K.k1b = __vimin3_s32(K.k1b, tb0, tb1);
      } else {
        K.k0a = __viaddmin_s32(a0, jx, K.k0a);      K.k0b = __viaddmin_s32(a1, jx + 1, K.k0b);
        K.k0a = __viaddmin_s32(a2, jx + 2, K.k0a);  K.k0b = __viaddmin_s32(a3, jx + 3, K.k0b);
      }
    }
  } else {
#pragma unroll
    for (int jx = 0; jx < 16; jx++) {
      if (c0 + jx < nvalid) {
        const int key = (int)v[jx] + jx;
        if (TOP2) { K.k1a = min(K.k1a, max(key, K.k0a)); K.k0a = min(key, K.k0a); }
        else K.k0a = min(key, K.k0a);
      }
    }
  }
}

// this thread's share (column half `half`) of one product of N columns (nvalid of them real) in the accumulator buffer at
// `taddr` (TMEM lane = row already included): full products straight-line, short ones chunk by chunk
template <bool TOP2>
__device__ __forceinline__ void tc_reduce_product(TcKeys& K, uint32_t taddr, int half, int N, int nvalid, int col0) {
  uint32_t va[16], vb[16];
  if (nvalid == TC_NPASS) {
    const int cbeg = half * (TC_NPASS / 2);
    TC_LD16(va, taddr + (uint32_t)cbeg);
    TC_LD16(vb, taddr + (uint32_t)(cbeg + 16));
    TC_WAIT16(va);
    tc_reduce16<TOP2, true>(K, va, col0, cbeg, nvalid);
    TC_LD16(va, taddr + (uint32_t)(cbeg + 32));
    TC_WAIT16(vb);
    tc_reduce16<TOP2, true>(K, vb, col0, cbeg + 16, nvalid);
    TC_LD16(vb, taddr + (uint32_t)(cbeg + 48));
    TC_WAIT16(va);
    tc_reduce16<TOP2, true>(K, va, col0, cbeg + 32, nvalid);
    TC_WAIT16(vb);
    tc_reduce16<TOP2, true>(K, vb, col0, cbeg + 48, nvalid);
  } else {
    // (a short product is split evenly as well: the first ceil(chunks / 2) chunks to half 0, the others to half 1)
    const int csplit = N == TC_NPASS ? TC_NPASS / 2 : 16 * (((N >> 4) + 1) >> 1);
    const int cbeg = half ? csplit : 0, cend = half ? N : csplit;
    if (cbeg < cend) {
      TC_LD16(va, taddr + (uint32_t)cbeg);
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        TC_WAIT16(va);
        if (c0 + 16 < cend) TC_LD16(vb, taddr + (uint32_t)(c0 + 16));
        tc_reduce16<TOP2, false>(K, va, col0, c0, nvalid);
        if (c0 + 16 < cend) {
          TC_WAIT16(vb);
          if (c0 + 32 < cend) TC_LD16(va, taddr + (uint32_t)(c0 + 32));
          tc_reduce16<TOP2, false>(K, vb, col0, c0 + 16, nvalid);
        }
      }
    }
  }
}

// thread 0: one 128 x N x 256 product (8 K-steps) into the accumulator buffer `d_tmem`, then commit to `bar`
__device__ __forceinline__ void tc_issue_product(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, int N, uint64_t* bar) {
  const uint32_t idesc = umma_idesc_i8(N);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
  for (int ks = 0; ks < 8; ks++) {  // K = 256 in steps of 32 int8 = two 16-byte chunks
    const uint64_t da = umma_desc(a_addr + ks * 2 * TC_LBO);
    const uint64_t db = umma_desc(b_addr + ks * 2 * TC_LBO);
    const uint32_t acc = ks > 0 ? 1u : 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One match direction: rows `m` (nm of them) against rows `n` (nn).  TOP2: keep the two best train rows per m-row
// (forward direction, [i][0] and [i][1]); else only the best (reverse direction, [j][0]).  The n side is expanded in
// full into sN, the m side tile by tile into sM.  Accumulators are double-buffered in TMEM: the products of pass i+1
// are issued before pass i is reduced.  A thread owns one m-row (= TMEM lane) and half of the columns of every pass, so
// the selection is thread-local; the two halves of a row are merged through shared memory once per tile and handed to
// `emit(m_row, k0, k1)`.
template <bool TOP2, class Emit>
__device__ __forceinline__ void tc_direction(const uint8_t* __restrict__ m_base, int nm, const uint8_t* __restrict__ n_base,
                                             int nn, uint8_t* sM, uint8_t* sN, uint64_t* bars /*[2]*/,
                                             uint32_t* phases /*[2]*/, uint32_t tmem, uint2* s_merge /*[TC_M]*/, Emit emit) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (TC_M - 1);     // m-row inside the tile = TMEM lane
  const int half = tid >> 7;            // which half of a pass's columns this thread reduces
  const int nnp = (nn + 15) & ~15;
  // this thread's half descriptor row of the first query tile is requested before the train side is expanded, and that of
  // the NEXT tile before the current tile is reduced (global latency)
  uint4 next_row = __ldg(reinterpret_cast<const uint4*>(m_base + (size_t)min(row, nm - 1) * 32) + half);
  {  // both n-rows of this thread are fetched before either is expanded
    uint4 lo[2], hi[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = tid + u * TC_THREADS;
      const uint4* src = reinterpret_cast<const uint4*>(n_base + (size_t)r * 32);
      if (r < nn) { lo[u] = __ldg(src); hi[u] = __ldg(src + 1); }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = tid + u * TC_THREADS;
      if (r < nn) expand_row(sN, r, lo[u], hi[u], TC_PAIR_TRAIN);
    }
    for (int r = tid + 2 * TC_THREADS; r < nn; r += TC_THREADS) {
      const uint4* src = reinterpret_cast<const uint4*>(n_base + (size_t)r * 32);
      expand_row(sN, r, __ldg(src), __ldg(src + 1), TC_PAIR_TRAIN);
    }
  }
  const uint32_t sM_addr = smem_u32(sM), sN_addr = smem_u32(sN);
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int npass = (nnp + TC_NPASS - 1) / TC_NPASS;

  auto issue = [&](int pass) {  // thread 0: the 8 K-steps of one pass into buffer (pass & 1), then commit to its barrier
    const int n0 = pass * TC_NPASS;
    const int N = min(TC_NPASS, nnp - n0);
    const uint32_t idesc = umma_idesc_i8(N);
    const uint32_t d_tmem = tmem + (uint32_t)((pass & 1) * TC_NPASS);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {  // K = 256 in steps of 32 int8 = two 16-byte chunks
      const uint64_t da = umma_desc(sM_addr + ks * 2 * TC_LBO);
      const uint64_t db = umma_desc(sN_addr + (n0 >> 3) * TC_SBO + ks * 2 * TC_LBO);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
          "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[pass & 1]))
                 : "memory");
  };

  for (int m0 = 0; m0 < nm; m0 += TC_M) {
    {
      const int mi = m0 + row;
      const uint4 cur_row = next_row;
      if (m0 + TC_M < nm)
        next_row = __ldg(reinterpret_cast<const uint4*>(m_base + (size_t)min(mi + TC_M, nm - 1) * 32) + half);
      if (mi < nm) expand_half_row(sM, row, half, cur_row, TC_PAIR_QUERY);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA
    __syncthreads();
    if (tid == 0) issue(0);
    // Two independent selection chains (even / odd columns), merged after the last pass.  Keys are signed and RELATIVE:
    // key = accumulator + column - kbase, where kbase (the first column of the 16-column chunk being reduced) is
    // subtracted from the running keys instead of being added to every new one — a new key is then `accumulator + jx`
    // with jx a compile-time constant, which the add-min-max instruction (VIADDMNMX) fuses: 2.5 instructions per
    // distance for the two nearest, 1 for the nearest (tc_reduce16).
    TcKeys K;
    K.reset();
    for (int pass = 0; pass < npass; pass++) {
      const int n0 = pass * TC_NPASS;
      const int N = min(TC_NPASS, nnp - n0);
      if (tid == 0 && pass + 1 < npass) issue(pass + 1);  // its buffer was drained before the barrier that ended pass - 1
      mbar_wait(&bars[pass & 1], phases[pass & 1]);
      phases[pass & 1] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // this thread's share of the pass (tc_reduce_product): half of the buffer's columns, 16 at a time, next chunk in flight
      // (a warp whose 32 rows all lie past nm has nothing to read: TMEM read bandwidth is what bounds this kernel)
      if (m0 + (warp & 3) * 32 < nm)
        tc_reduce_product<TOP2>(K, tlane + (uint32_t)((pass & 1) * TC_NPASS), half, N, min(N, nn - n0), n0);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // this buffer is drained (pass + 2 may overwrite it); after the last pass sM is free as well
    }
    // relative signed keys -> the packed keys of the integer kernels: distance << 20 | train row
    uint32_t k0 = tc_canonical(K.k0a, K.kbase), k1 = TOP2 ? tc_canonical(K.k1a, K.kbase) : KEY_NONE;
    if (TOP2) { top2_update(k0, k1, tc_canonical(K.k0b, K.kbase)); top2_update(k0, k1, tc_canonical(K.k1b, K.kbase)); }
    else k0 = min(k0, tc_canonical(K.k0b, K.kbase));
    if (half == 1) s_merge[row] = make_uint2(k0, k1);
    __syncthreads();
    const int mi = m0 + row;
    if (half == 0 && mi < nm) {
      const uint2 o = s_merge[row];
      if (TOP2) { top2_update(k0, k1, o.x); top2_update(k0, k1, o.y); }
      else k0 = min(k0, o.x);
      emit(mi, k0, k1);  // packed keys (distance << 20 | n-row); k1 == KEY_NONE without a second row (or !TOP2)
    }
    // (the next tile's first __syncthreads orders the s_merge reads above before its writes)
  }
}

// CTA-wide set-up of the tensor path: TMEM allocation (warp 0), the two accumulator barriers, the nibble LUT
struct TcShared {
  uint64_t bar[2];
  uint32_t tmem;
  uint2 merge[TC_M];
};
__device__ __forceinline__ uint32_t tc_begin(TcShared& S) {
  const int tid = threadIdx.x;
  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return S.tmem;
}
__device__ __forceinline__ void tc_end(uint32_t tmem) {
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}
__host__ __device__ inline size_t tc_smem_bytes(int cap_n) { return (size_t)TC_M * TC_ROW_BYTES + (size_t)((cap_n + 15) & ~15) * TC_ROW_BYTES; }

}  // namespace gfpl
