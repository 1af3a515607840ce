// k_hamming_tc.cu — knn k=2 (forward) + best match (reverse) of 256-bit descriptors on the 5th-generation tensor cores.
//
// The INT-pipe kernel of k_hamming.cu runs at ~85 % of the ALU pipe: 5 POPC + ~25 ALU operations per 256-bit distance.
// The only higher ceiling on the chip for this arithmetic is the tensor pipe, and the distance matrix of one problem
// IS a matrix product: with every descriptor bit mapped to an int8 +64 / -64 (query side negated),
//     dot(a, b) = -4096 * (#equal bits - #different bits) = (hamming(a, b) << 13) - 2^20     (exact in the s32 accumulator)
// so D = A * B^T (M = queries, N = train rows, K = 256) gives all distances of the problem, already in the form of the
// distance field of a sort key.  One CTA per problem:
//   1. the 32-byte descriptors are expanded to int8 rows in shared memory (byte permutes, gfpl_tc.cuh), directly in the
//      K-major no-swizzle UMMA layout (8-row x 16-byte core matrices: byte(r, k) at (r/8)*SBO + (k/16)*LBO + (r%8)*16 + k%16),
//   2. one thread issues tcgen05.mma.kind::i8 (M = 128, N <= 128, K = 32 per instruction, 8 per product) into one of two
//      TMEM accumulator buffers and commits to that buffer's mbarrier; the next product is issued before this one is read,
//   3. the CTA's 256 threads read accumulator rows with tcgen05.ld (TMEM lane = row, two threads per row, half of the
//      columns each) and keep the top-2 (top-1) of the row with fused add-min/max instructions: 2.5 (1) per distance —
//      thread-local, no shuffles, no atomics.
// knn2_tc_kernel (128..~650 rows per side, NORM_HAMMING) expands each side once and takes both directions from it;
// knn2_tc_small_kernel (<= 128 rows per side, both norms) is persistent over problems.
// Results are bit-identical to k_hamming.cu: same keys (distance << 20 | index), same tie rule.
#include "gfpl_internal.h"
#include "gfpl_tc.cuh"

namespace gfpl {

// Both directions from ONE expansion of each side.  The train rows b are expanded once into sN; a query tile of a in sM
// then serves two families of products (both UMMA operands are K-major, so a tile can be either operand):
//   forward  D  = sM x chunk_p(sN)^T : rows = the tile's a-rows, columns = 128 b-rows  -> two nearest b-rows of every a-row
//   reverse  D' = chunk_i(sN) x sM^T : rows = 128 b-rows, columns = the tile's a-rows  -> nearest a-row of every b-row,
//                                      kept across the tiles in s_rev[b-row]
// i.e. 2 * ceil(nb / 128) products per tile in one uninterrupted pipeline over the two TMEM buffers, and half the
// expansion work of two separate directions.  sM lies BEHIND sN: the 128-row M operand of the last reverse chunk may start
// in sN's last rows and run into sM — inside the allocation, and those lanes are never read.
__host__ __device__ inline size_t knn2_tc_smem_bytes(int cap_b) {
  return (size_t)((cap_b + 15) & ~15) * TC_ROW_BYTES + (size_t)TC_M * TC_ROW_BYTES + 2 * (size_t)cap_b * sizeof(uint32_t);
}

// BLOCKED (train sides too large to sit expanded in shared memory, e.g. 2000 x 2000 rows): grid = (column blocks, problems).
// A CTA owns the b-rows [blockIdx.x * blk, + blk) of its problem — expanded once, resident — and walks ALL a-tiles against
// them.  The reverse direction is complete inside the CTA (every a-row meets the block's b-rows); the forward direction
// yields the two nearest b-rows WITHIN the block per a-row: packed keys (distance << 20 | global b-row) go to
// `part` [problem][block][cap_a], and knn2_tc_merge_kernel takes the two smallest keys over the blocks (the integer order of
// the keys is OpenCV's order, so the merged result is bit-identical to the unblocked one).  The a-tiles are re-expanded by
// every block of the problem: ~+15 % ALU work at 256-row blocks against a train side that no longer has to fit.
template <bool BLOCKED>
__global__ void __launch_bounds__(TC_THREADS)
knn2_tc_kernel(const uint8_t* __restrict__ a, const int32_t* __restrict__ na, int cap_a, const uint8_t* __restrict__ b,
               const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ab, Knn2Out out_ba, int blk, uint2* __restrict__ part) {
  const int cap_n = BLOCKED ? blk : cap_b;                                   // b-rows this CTA holds
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sN = smem;                                                        // [cap_n rounded to 16][256] rows of b
  uint8_t* sM = smem + (size_t)((cap_n + 15) & ~15) * TC_ROW_BYTES;          // [128][256] one tile of a
  uint32_t* s_rev = reinterpret_cast<uint32_t*>(sM + TC_M * TC_ROW_BYTES);   // [2][cap_n] best key of every b-row so far, per column half
  __shared__ __align__(8) TcShared S;

  const int p = BLOCKED ? blockIdx.y : blockIdx.x;
  const int j_first = BLOCKED ? blockIdx.x * blk : 0;                        // first b-row of this CTA
  const int nq = na[p];
  const int nt_all = nb[p];
  const int nt = BLOCKED ? max(0, min(blk, nt_all - j_first)) : nt_all;     // b-rows of this CTA
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (TC_M - 1);     // row inside a tile = TMEM lane
  const int half = tid >> 7;            // which half of a product's columns this thread reduces
  const float inf = __int_as_float(0x7f800000);
  const uint8_t* q_base = a + (size_t)p * cap_a * 32;
  const uint8_t* t_base = b + ((size_t)p * cap_b + j_first) * 32;
  int2* idx_ab = out_ab.idx + (size_t)p * cap_a;
  float2* dist_ab = out_ab.dist + (size_t)p * cap_a;
  int2* idx_ba = out_ba.idx + (size_t)p * cap_b + j_first;
  float2* dist_ba = out_ba.dist + (size_t)p * cap_b + j_first;
  uint2* part_p = BLOCKED ? part + ((size_t)p * gridDim.x + blockIdx.x) * cap_a : nullptr;
  if (nq == 0 || nt == 0) {  // nothing to multiply: "no match" rows on both sides (uniform exit, before any allocation)
    if (BLOCKED) { for (int i = tid; i < nq; i += TC_THREADS) part_p[i] = make_uint2(KEY_NONE, KEY_NONE); }
    else { for (int i = tid; i < nq; i += TC_THREADS) { idx_ab[i] = make_int2(-1, -1); dist_ab[i] = make_float2(inf, inf); } }
    for (int j = tid; j < nt; j += TC_THREADS) { idx_ba[j] = make_int2(-1, -1); dist_ba[j] = make_float2(inf, inf); }
    return;
  }
  const uint32_t tmem = tc_begin(S);
  uint2* s_merge = S.merge;             // [TC_M]: the column halves of an a-row meet here once per tile
  // this thread's half descriptor row of the first a-tile is requested before b is expanded (global latency)
  uint4 next_row = __ldg(reinterpret_cast<const uint4*>(q_base + (size_t)min(row, nq - 1) * 32) + half);
  for (int r = tid; r < nt; r += TC_THREADS) {
    const uint4* src = reinterpret_cast<const uint4*>(t_base + (size_t)r * 32);
    expand_row(sN, r, __ldg(src), __ldg(src + 1), TC_PAIR_TRAIN);
    s_rev[r] = KEY_NONE;
    s_rev[cap_n + r] = KEY_NONE;
  }
  const uint32_t sM_addr = smem_u32(sM), sN_addr = smem_u32(sN);
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int nnp = (nt + 15) & ~15;
  const int nchunk = (nnp + TC_NPASS - 1) / TC_NPASS;   // 128-row chunks of b: forward passes and reverse M tiles
  const int P = 2 * nchunk;                             // products per a-tile
  uint32_t phases[2] = {0u, 0u};
  uint32_t prod = 0u;                                   // products so far: product t uses accumulator buffer t & 1

  for (int m0 = 0; m0 < nq; m0 += TC_M) {
    {
      const int mi = m0 + row;
      const uint4 cur_row = next_row;
      if (m0 + TC_M < nq)
        next_row = __ldg(reinterpret_cast<const uint4*>(q_base + (size_t)min(mi + TC_M, nq - 1) * 32) + half);
      if (mi < nq) expand_half_row(sM, row, half, cur_row, TC_PAIR_QUERY);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA
    __syncthreads();
    const int nqt = min(TC_M, nq - m0);            // a-rows of this tile
    const int nqt16 = (nqt + 15) & ~15;
    auto issue = [&](int t, uint32_t buf) {        // thread 0: product t of this tile into accumulator buffer buf
      const uint32_t d_tmem = tmem + buf * TC_NPASS;
      if (t < nchunk) tc_issue_product(d_tmem, sM_addr, sN_addr + (uint32_t)t * (TC_NPASS / 8) * TC_SBO, min(TC_NPASS, nnp - t * TC_NPASS), &S.bar[buf]);
      else tc_issue_product(d_tmem, sN_addr + (uint32_t)(t - nchunk) * (TC_NPASS / 8) * TC_SBO, sM_addr, nqt16, &S.bar[buf]);
    };
    if (tid == 0) issue(0, prod & 1u);
    TcKeys F;   // the two nearest b-rows of this thread's a-row, over the forward products of the tile
    F.reset();
    for (int t = 0; t < P; t++, prod++) {
      const uint32_t buf = prod & 1u;
      if (tid == 0 && t + 1 < P) issue(t + 1, buf ^ 1u);   // that buffer was drained before the barrier that ended product t - 1
      mbar_wait(&S.bar[buf], phases[buf]);
      phases[buf] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tlane + buf * TC_NPASS;
      const bool fwd = t < nchunk;
      uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
      if (fwd) {
        const int n0 = t * TC_NPASS;
        const int N = min(TC_NPASS, nnp - n0);
        if (m0 + (warp & 3) * 32 < nq) tc_reduce_product<true>(F, taddr, half, N, min(N, nt - n0), n0);
        if (t == nchunk - 1) {   // last forward product of the tile: the row's two nearest are complete
          k0 = tc_canonical(F.k0a, F.kbase); k1 = tc_canonical(F.k1a, F.kbase);
          top2_update(k0, k1, tc_canonical(F.k0b, F.kbase));
          top2_update(k0, k1, tc_canonical(F.k1b, F.kbase));
          if (half == 1) s_merge[row] = make_uint2(k0, k1);
        }
      } else {
        const int j0 = (t - nchunk) * TC_NPASS;     // first b-row of this reverse chunk
        if (j0 + (warp & 3) * 32 < nt) {
          TcKeys R;
          R.reset();
          tc_reduce_product<false>(R, taddr, half, nqt16, nqt, m0);
          const int j = j0 + row;                   // this thread owns (b-row j, its column half) in every tile
          if (j < nt) s_rev[half * cap_n + j] = min(s_rev[half * cap_n + j], min(tc_canonical(R.k0a, R.kbase), tc_canonical(R.k0b, R.kbase)));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // this buffer is drained (the product after the next may overwrite it)
      if (fwd && t == nchunk - 1 && half == 0) {
        const int mi = m0 + row;
        if (mi < nq) {
          const uint2 o = s_merge[row];
          top2_update(k0, k1, o.x); top2_update(k0, k1, o.y);
          if (BLOCKED) {   // keys carry block-local b-rows: make them global (the index field is the low 20 bits)
            part_p[mi] = make_uint2(k0 + (uint32_t)j_first, k1 != KEY_NONE ? k1 + (uint32_t)j_first : KEY_NONE);
          } else {
            idx_ab[mi] = make_int2((int)(k0 & KEY_IDX_MASK), k1 != KEY_NONE ? (int)(k1 & KEY_IDX_MASK) : -1);
            dist_ab[mi] = make_float2((float)(k0 >> KEY_IDX_BITS), k1 != KEY_NONE ? (float)(k1 >> KEY_IDX_BITS) : inf);
          }
        }
      }
    }
    // (sM may be overwritten: every product of the tile has completed and was waited for by every thread)
  }
  tc_end(tmem);   // CTA barrier inside: s_rev is complete
  for (int j = tid; j < nt; j += TC_THREADS) {
    const uint32_t k0 = min(s_rev[j], s_rev[cap_n + j]);
    idx_ba[j] = make_int2((int)(k0 & KEY_IDX_MASK), -1);
    dist_ba[j] = make_float2((float)(k0 >> KEY_IDX_BITS), inf);
  }
}

// forward direction of the blocked form: the two smallest keys of every a-row over the column blocks
__global__ void __launch_bounds__(256)
knn2_tc_merge_kernel(const int32_t* __restrict__ na, int cap_a, int nblk, const uint2* __restrict__ part, Knn2Out out_ab) {
  const int p = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na[p]) return;
  const float inf = __int_as_float(0x7f800000);
  uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
  for (int c = 0; c < nblk; c++) {
    const uint2 k = part[((size_t)p * nblk + c) * cap_a + i];
    top2_update(k0, k1, k.x);
    top2_update(k0, k1, k.y);
  }
  out_ab.idx[(size_t)p * cap_a + i] = make_int2(k0 != KEY_NONE ? (int)(k0 & KEY_IDX_MASK) : -1, k1 != KEY_NONE ? (int)(k1 & KEY_IDX_MASK) : -1);
  out_ab.dist[(size_t)p * cap_a + i] = make_float2(k0 != KEY_NONE ? (float)(k0 >> KEY_IDX_BITS) : inf, k1 != KEY_NONE ? (float)(k1 >> KEY_IDX_BITS) : inf);
}

// ---- small problems (both sides <= 128 rows: line descriptors), both norms -----------------------------------------
// One query tile, one pass, and the SAME expanded rows serve both products (both operands of an UMMA are K-major here):
// D = A B^T gives the two nearest b-rows of every a-row (warps 0..3, one TMEM lane = one a-row), D^T = B A^T the nearest
// a-row of every b-row (warps 4..7) — nothing is expanded twice and no warp idles.  Both sides carry +64 / -64, so a
// product is 4096 * (#equal - #different) and the NEAREST row has the LARGEST accumulator: the selection keeps maxima of
// `accumulator - column` (ties -> lowest column, OpenCV's rule).
// NORM_HAMMING2 (number of differing 2-bit groups) is bilinear too: append to the 256 bit values the 128 products of the
// two bits of every group (in +-1 arithmetic the XOR of two bits is their product).  Per group the three terms add up to
// 3 when the groups are equal and to -1 otherwise (the vertices of a tetrahedron), i.e. K = 384 and
// accumulator = 4096 * (4 * #equal groups - 128) = 4096 * 384 - 16384 * distance.
template <int NORM> struct TcSmall {
  static constexpr int KB = NORM == GFPL_NORM_HAMMING2 ? 384 : 256;   // int8 per expanded row
  static constexpr int CH = KB / 16;                                  // 16-byte K chunks per row
  static constexpr int SBO = CH * TC_LBO;                             // bytes between 8-row groups
  static constexpr int KSTEPS = KB / 32;
  static constexpr int ACC0 = 4096 * KB;                              // accumulator at distance 0
  static constexpr int SHIFT = NORM == GFPL_NORM_HAMMING2 ? 14 : 13;  // accumulator = ACC0 - (distance << SHIFT)
  static constexpr size_t SMEM = 2 * (size_t)(TC_M / 8) * SBO;
};

__device__ __forceinline__ uint64_t umma_desc_sbo(uint32_t smem_addr, int sbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

template <int NORM>
__device__ __forceinline__ void expand_row_small(uint8_t* tile, int r, const uint8_t* __restrict__ src) {
  using T = TcSmall<NORM>;
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(src)), hi = __ldg(reinterpret_cast<const uint4*>(src) + 1);
  uint8_t* row = tile + (r >> 3) * T::SBO + (r & 7) * 16;
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int c = 0; c < 8; c++) {
    uint4 x, y;
    expand_word(w[c], TC_PAIR_TRAIN, x, y);
    *reinterpret_cast<uint4*>(row + (2 * c) * TC_LBO) = x;
    *reinterpret_cast<uint4*>(row + (2 * c + 1) * TC_LBO) = y;
  }
  if (NORM == GFPL_NORM_HAMMING2) {
#pragma unroll
    for (int c = 0; c < 4; c++) {   // XOR of the two bits of each group: 16 group bits per word at the even positions
      const uint32_t g0 = (w[2 * c] ^ (w[2 * c] >> 1)) & 0x55555555u, g1 = (w[2 * c + 1] ^ (w[2 * c + 1] >> 1)) & 0x55555555u;
      uint4 x, y;
      expand_word(g0 | (g1 << 1), TC_PAIR_TRAIN, x, y);   // the groups of two words interleaved: 32 group bits -> 32 int8
      *reinterpret_cast<uint4*>(row + (16 + 2 * c) * TC_LBO) = x;
      *reinterpret_cast<uint4*>(row + (16 + 2 * c + 1) * TC_LBO) = y;
    }
  }
}

template <int NORM>
__global__ void __launch_bounds__(TC_THREADS)
knn2_tc_small_kernel(int B, const uint8_t* __restrict__ a, const int32_t* __restrict__ na, int cap_a, const uint8_t* __restrict__ b,
                     const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ab, Knn2Out out_ba) {
  using T = TcSmall<NORM>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                               // rows of a, expanded
  uint8_t* sB = smem + (TC_M / 8) * T::SBO;         // rows of b
  __shared__ __align__(8) TcShared S;
  const int tid = threadIdx.x, warp = tid >> 5;
  const float inf = __int_as_float(0x7f800000);
  const int dir = tid >> 7;                         // 0: rows of a against b (two nearest), 1: rows of b against a (nearest)
  const int row = tid & (TC_M - 1);
  {  // a CTA whose problems are all left to the integer kernel (e.g. frame-to-frame lines) returns before any set-up
    __shared__ int s_any;
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (int p = blockIdx.x + tid * gridDim.x; p < B; p += gridDim.x * TC_THREADS)
      if (!tc_small_skips(na[p], nb[p])) s_any = 1;
    __syncthreads();
    if (!s_any) return;
  }
  // persistent over problems: TMEM, barriers and the LUT are set up once per CTA
  const uint32_t tmem = tc_begin(S);
  uint32_t phase = 0u;
  for (int p = blockIdx.x; p < B; p += gridDim.x) {
    const int nq = na[p], nt = nb[p];
    const uint8_t* a_base = a + (size_t)p * cap_a * 32;
    const uint8_t* b_base = b + (size_t)p * cap_b * 32;
    int2* idx_o = dir == 0 ? out_ab.idx + (size_t)p * cap_a : out_ba.idx + (size_t)p * cap_b;
    float2* dist_o = dir == 0 ? out_ab.dist + (size_t)p * cap_a : out_ba.dist + (size_t)p * cap_b;
    const int nrows = dir == 0 ? nq : nt;            // rows of this thread's direction
    const int ncols = dir == 0 ? nt : nq;            // the rows they are matched against
    const uint8_t* my_rows = dir == 0 ? a_base : b_base;
    if (tc_small_skips(nq, nt)) continue;   // CTA-uniform: served by the integer kernel launched beside this one
    // threads 0..127: row tid of a, threads 128..255: row tid - 128 of b.  Rows past the end are never read back.
    if (row < nrows) expand_row_small<NORM>(dir == 0 ? sA : sB, row, my_rows + (size_t)row * 32);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int nq16 = (nq + 15) & ~15, nt16 = (nt + 15) & ~15;
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int d = 0; d < 2; d++) {   // d 0: D = A B^T -> columns [0, 128); d 1: D^T = B A^T -> columns [128, 256)
        const uint32_t mA = smem_u32(d == 0 ? sA : sB), nA = smem_u32(d == 0 ? sB : sA);
        const uint32_t idesc = umma_idesc_i8(d == 0 ? nt16 : nq16);
        const uint32_t d_tmem = tmem + (uint32_t)(d * TC_NPASS);
#pragma unroll
        for (int ks = 0; ks < T::KSTEPS; ks++) {
          const uint64_t da = umma_desc_sbo(mA + ks * 2 * TC_LBO, T::SBO);
          const uint64_t db = umma_desc_sbo(nA + ks * 2 * TC_LBO, T::SBO);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
              "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.bar[d])) : "memory");
      }
    }
    // selection: warps 0..3 the rows of a (two nearest), warps 4..7 the rows of b (nearest); TMEM lane = row
    const int ncols16 = dir == 0 ? nt16 : nq16;
    constexpr int KNONE = -0x3FFFFFFF;               // stays below every real key however often kbase moves
    int k0a = KNONE, k1a = KNONE, k0b = KNONE, k1b = KNONE;
    int kbase = 0;                                   // keys are relative: accumulator - (column - kbase)
    mbar_wait(&S.bar[dir], phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if ((warp & 3) * 32 < nrows) {
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(dir * TC_NPASS);
      uint32_t va[16], vb[16];
      auto reduce16 = [&](const uint32_t* v, int c0) {
        const int shift = c0 - kbase;
        kbase = c0;
        k0a += shift; k0b += shift; k1a += shift; k1b += shift;
        if (c0 + 16 <= ncols) {
#pragma unroll
          for (int jx = 0; jx < 16; jx += 4) {
            const int a0 = (int)v[jx], a1 = (int)v[jx + 1], a2 = (int)v[jx + 2], a3 = (int)v[jx + 3];
            if (dir == 0) {
              const int ta0 = __viaddmin_s32(a0, -jx, k0a);        k0a = __viaddmax_s32(a0, -jx, k0a);
              const int tb0 = __viaddmin_s32(a1, -(jx + 1), k0b);  k0b = __viaddmax_s32(a1, -(jx + 1), k0b);
              const int ta1 = __viaddmin_s32(a2, -(jx + 2), k0a);  k0a = __viaddmax_s32(a2, -(jx + 2), k0a);
              const int tb1 = __viaddmin_s32(a3, -(jx + 3), k0b);  k0b = __viaddmax_s32(a3, -(jx + 3), k0b);
              k1a = __vimax3_s32(k1a, ta0, ta1);
              k1b = __vimax3_s32(k1b, tb0, tb1);
            } else {
              k0a = __viaddmax_s32(a0, -jx, k0a);        k0b = __viaddmax_s32(a1, -(jx + 1), k0b);
              k0a = __viaddmax_s32(a2, -(jx + 2), k0a);  k0b = __viaddmax_s32(a3, -(jx + 3), k0b);
            }
          }
        } else {
#pragma unroll
          for (int jx = 0; jx < 16; jx++) {
            if (c0 + jx < ncols) {
              const int key = (int)v[jx] - jx;
              const int t = min(k0a, key);
              k0a = max(k0a, key);
              k1a = max(k1a, t);
            }
          }
        }
      };
      TC_LD16(va, taddr);
      for (int c0 = 0; c0 < ncols16; c0 += 32) {
        TC_WAIT16(va);
        if (c0 + 16 < ncols16) TC_LD16(vb, taddr + (uint32_t)(c0 + 16));
        reduce16(va, c0);
        if (c0 + 16 < ncols16) {
          TC_WAIT16(vb);
          if (c0 + 32 < ncols16) TC_LD16(va, taddr + (uint32_t)(c0 + 32));
          reduce16(vb, c0 + 16);
        }
      }
      if (row < nrows) {
        // merge the two chains (largest first), then relative signed keys -> (distance, column)
        const int k0 = max(k0a, k0b);
        const int k1 = max(max(k1a, k1b), min(k0a, k0b));
        const bool has0 = k0 > (KNONE >> 1), has1 = dir == 0 && k1 > (KNONE >> 1);
        const uint32_t u0 = (uint32_t)(T::ACC0 - (k0 - kbase)), u1 = (uint32_t)(T::ACC0 - (k1 - kbase));   // distance << SHIFT | column
        const int i0 = has0 ? (int)(u0 & ((1u << T::SHIFT) - 1u)) : -1, i1 = has1 ? (int)(u1 & ((1u << T::SHIFT) - 1u)) : -1;
        const float d0 = has0 ? (float)(u0 >> T::SHIFT) : inf, d1 = has1 ? (float)(u1 >> T::SHIFT) : inf;
        idx_o[row] = make_int2(i0, i1);
        dist_o[row] = make_float2(d0, d1);
      }
    }
    // the accumulators are drained and every product has completed: the next problem may overwrite TMEM and the rows
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tc_end(tmem);
}

bool knn2_tc_small_fits(int cap_a, int cap_b) { return cap_a <= TC_M && cap_b <= TC_M && cap_a >= 32 && cap_b >= 32; }

int launch_knn2_tc_small(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b, const int32_t* nb,
                         int cap_b, int norm, Knn2Out out_ab, Knn2Out out_ba) {
  // persistent CTAs: two per SM (256 TMEM columns each), each walks the batch with the grid as stride
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const int grid = B < 2 * sms ? B : 2 * sms;
  if (norm == GFPL_NORM_HAMMING2) {
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(knn2_tc_small_kernel<GFPL_NORM_HAMMING2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)TcSmall<GFPL_NORM_HAMMING2>::SMEM));
    knn2_tc_small_kernel<GFPL_NORM_HAMMING2><<<grid, TC_THREADS, TcSmall<GFPL_NORM_HAMMING2>::SMEM, ctx->stream>>>(B, a, na, cap_a, b, nb, cap_b,
                                                                                                                    out_ab, out_ba);
  } else {
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(knn2_tc_small_kernel<GFPL_NORM_HAMMING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)TcSmall<GFPL_NORM_HAMMING>::SMEM));
    knn2_tc_small_kernel<GFPL_NORM_HAMMING><<<grid, TC_THREADS, TcSmall<GFPL_NORM_HAMMING>::SMEM, ctx->stream>>>(B, a, na, cap_a, b, nb, cap_b,
                                                                                                                  out_ab, out_ba);
  }
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// ---- roofline denominator: dense tcgen05.mma kind::i8 throughput, measured ----------------------------------------
// One CTA per SM, operands resident in shared memory (no loads in the loop), M = 128, N = 256, K = 256 per product
// (8 instructions of K = 32), products issued back to back into the two halves of the SM's 512 TMEM columns; a commit +
// wait every PEAK_BURST products keeps the issue queue bounded.  Operand bytes are pseudo-random (+-64 like the
// matching operands would toggle fewer datapath bits and flatter the clock).  ops = 2 * 128 * 256 * 256 per product.
constexpr int PEAK_N = 256, PEAK_BURST = 8;
constexpr size_t PEAK_SMEM = (size_t)(TC_M + PEAK_N) * TC_ROW_BYTES;

__global__ void __launch_bounds__(128) int8_peak_kernel(int products, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  uint32_t x = (blockIdx.x * 128u + tid) * 2654435761u + 12345u;
  for (size_t i = tid * 4; i < PEAK_SMEM; i += 128 * 4) {
    x = x * 1664525u + 1013904223u;
    *reinterpret_cast<uint32_t*>(smem + i) = x;
  }
  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + (size_t)TC_M * TC_ROW_BYTES);
    const uint32_t idesc = umma_idesc_i8(PEAK_N);
    uint32_t phase = 0u;
    for (int p = 0; p < products; p += PEAK_BURST) {
      for (int q = 0; q < PEAK_BURST; q++) {
        const uint32_t d_tmem = tmem + (uint32_t)((q & 1) * PEAK_N);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
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
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(&bar, phase);
      phase ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if ((tid >> 5) == 0) {   // read one accumulator word so that the products are observable
    uint32_t v[16];
    TC_LD16(v, tmem);
    TC_WAIT16(v);
    if (tid == 0) sink[blockIdx.x] = v[0];
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int launch_measure_int8_peak(gfpl_ctx* ctx, double* int8_tops) {
  const int blocks = ctx->sm_count, products = 2048;
  uint32_t* sink = nullptr;
  GFPL_CUDA_OK(ctx, cudaMalloc((void**)&sink, (size_t)blocks * sizeof(uint32_t)));
  GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PEAK_SMEM));
  cudaEvent_t e0, e1;
  GFPL_CUDA_OK(ctx, cudaEventCreate(&e0));
  GFPL_CUDA_OK(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    float ms;
    cudaEventRecord(e0, ctx->stream);
    int8_peak_kernel<<<blocks, 128, PEAK_SMEM, ctx->stream>>>(products, sink);
    cudaEventRecord(e1, ctx->stream);
    GFPL_LAUNCH_OK(ctx);
    GFPL_CUDA_OK(ctx, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *int8_tops = (double)blocks * products * 2.0 * TC_M * PEAK_N * 256.0 / (best * 1e-3) / 1e12;
  return GFPL_OK;
}

// a train side of up to TC_RESIDENT_ROWS rows sits expanded in shared memory with two CTAs per SM (304 rows: 112 KB)
constexpr int TC_RESIDENT_ROWS = 320;
#ifndef TC_BLOCK_ROWS_N
#define TC_BLOCK_ROWS_N 256   // 256: two full accumulator passes per block (2000 x 2000: 2.20 ms per 1024 problems); 288: 2.52, 224: 2.46
#endif
constexpr int TC_BLOCK_ROWS = TC_BLOCK_ROWS_N;     // column block of the blocked form, two CTAs per SM

bool knn2_tc_fits(int cap_a, int cap_b) {
  // small problems are dominated by the fixed cost of a CTA (TMEM allocation, barriers): they go to knn2_tc_small_kernel;
  // larger train sides than the resident form holds go through the column-blocked form
  return cap_a >= 128 && cap_b >= 128 && cap_a < (1 << TC_KEY_SHIFT) && cap_b < (1 << TC_KEY_SHIFT);
}

// number of column blocks and their size (a multiple of 16, as even as possible, at most TC_BLOCK_ROWS)
void tc_column_blocks(int cap_n, int* nblk, int* blk) {
  *nblk = (cap_n + TC_BLOCK_ROWS - 1) / TC_BLOCK_ROWS;
  *blk = (((cap_n + *nblk - 1) / *nblk) + 15) & ~15;
}

int launch_knn2_tc(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b, const int32_t* nb,
                   int cap_b, Knn2Out out_ab, Knn2Out out_ba) {
  if (cap_b <= TC_RESIDENT_ROWS) {
    const size_t smem = knn2_tc_smem_bytes(cap_b);
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(knn2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn2_tc_kernel<false><<<B, TC_THREADS, smem, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba, 0, nullptr);
    GFPL_LAUNCH_OK(ctx);
    return GFPL_OK;
  }
  int nblk, blk;
  tc_column_blocks(cap_b, &nblk, &blk);
  uint2* part = nullptr;   // forward partials [B][nblk][cap_a], from the context's stream-ordered pool
  const size_t part_bytes = (size_t)B * nblk * cap_a * sizeof(uint2);
  GFPL_CUDA_OK(ctx, ctx->pool ? cudaMallocFromPoolAsync((void**)&part, part_bytes, ctx->pool, ctx->stream)
                              : cudaMallocAsync((void**)&part, part_bytes, ctx->stream));
  int rc = GFPL_OK;
  const size_t smem = knn2_tc_smem_bytes(blk);
  if (cudaFuncSetAttribute(knn2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) rc = GFPL_E_CUDA;
  if (rc == GFPL_OK) {
    knn2_tc_kernel<true><<<dim3(nblk, B), TC_THREADS, smem, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba, blk, part);
    knn2_tc_merge_kernel<<<dim3((cap_a + 255) / 256, B), 256, 0, ctx->stream>>>(na, cap_a, nblk, part, out_ab);
    ctx->launches += 2;
    if (cudaGetLastError() != cudaSuccess) rc = GFPL_E_CUDA;
  }
  cudaFreeAsync(part, ctx->stream);   // stream-ordered: after the merge
  if (rc != GFPL_OK) set_last_error(ctx, "blocked tensor-core knn launch failed (%d blocks of %d rows)", nblk, blk);
  return rc;
}

}  // namespace gfpl
