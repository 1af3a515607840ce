// k_hamming_tc.cu — NORM_HAMMING knn k=2 (forward) + best match (reverse) on the 5th-generation tensor cores.
//
// The INT-pipe kernel of k_hamming.cu runs at ~85 % of the ALU pipe: 5 POPC + ~25 ALU operations per 256-bit distance.
// The only higher ceiling on the chip for this arithmetic is the tensor pipe, and the distance matrix of one problem
// IS a matrix product: with every descriptor bit mapped to an int8 +1 / -1,
//     dot(a, b) = #equal bits - #different bits = 256 - 2 * hamming(a, b)          (exact in the s32 accumulator)
// so D = A * B^T (M = queries, N = train rows, K = 256) gives all distances of the problem.  One CTA per problem:
//   1. the 32-byte descriptors are expanded to int8 rows in shared memory, directly in the K-major no-swizzle UMMA
//      layout (8-row x 16-byte core matrices: byte(r, k) at (r/8)*SBO + (k/16)*LBO + (r%8)*16 + k%16),
//   2. one thread issues tcgen05.mma.kind::i8 (M = 128, N <= 128, K = 32 per instruction, 8 per pass) into one of two
//      TMEM accumulator buffers and commits to that buffer's mbarrier; the next pass is issued before this one is read,
//   3. the CTA's 256 threads read accumulator rows with tcgen05.ld (TMEM lane = query row, two threads per row, half of
//      the columns each) and keep the packed-key top-2 of the row — thread-local, no shuffles, no atomics.
// The reverse direction (nearest query of every train row) is the same product with the roles swapped (tensor time is
// nearly free here) and a top-1 selection.  Per distance the ALU pipe sees ~5 (+3 reverse) operations instead of ~25.
// Results are bit-identical to k_hamming.cu: same keys (distance << 20 | index), same tie rule.
#include "gfpl_hamming.cuh"
#include "gfpl_internal.h"

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

// 4 descriptor bits -> 4 int8: bit set -> +1 (0x01), clear -> -1 (0xFF)
__device__ __forceinline__ uint32_t expand_nibble(uint32_t nib) {
  const uint32_t t = (nib * 0x00204081u) & 0x01010101u;  // bit b of the nibble -> LSB of byte b
  return t | ((t ^ 0x01010101u) * 255u);
}

// one descriptor row (8 words, or zeros past the end) -> 16 chunks of 16 int8 in the UMMA layout.  lut[nibble] = its 4
// int8: 16 words in 16 different banks, so the look-ups of a warp never conflict.
__device__ __forceinline__ void expand_row(uint8_t* tile, int r, const uint4& lo, const uint4& hi, bool zero,
                                           const uint32_t* __restrict__ lut) {
  uint8_t* row = tile + (r >> 3) * TC_SBO + (r & 7) * 16;
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const uint32_t bits = w[c >> 1] >> ((c & 1) * 16);
    uint4 v = make_uint4(lut[bits & 15u], lut[(bits >> 4) & 15u], lut[(bits >> 8) & 15u], lut[(bits >> 12) & 15u]);
    if (zero) v = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(row + c * TC_LBO) = v;
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

// One match direction: rows `m` (nm of them) against rows `n` (nn).  TOP2: keep the two best train rows per m-row
// (forward direction, [i][0] and [i][1]); else only the best (reverse direction, [j][0]).  The n side is expanded in
// full into sN, the m side tile by tile into sM.  Accumulators are double-buffered in TMEM: the products of pass i+1
// are issued before pass i is reduced.  A thread owns one m-row (= TMEM lane) and half of the columns of every pass, so
// the selection is thread-local; the two halves of a row are merged through shared memory once per tile.
template <bool TOP2>
__device__ __forceinline__ void tc_direction(const uint8_t* __restrict__ m_base, int nm, const uint8_t* __restrict__ n_base,
                                             int nn, int2* __restrict__ out_idx, float2* __restrict__ out_dist, uint8_t* sM,
                                             uint8_t* sN, const uint32_t* lut, uint64_t* bars /*[2]*/, uint32_t* phases /*[2]*/,
                                             uint32_t tmem, uint2* s_merge /*[TC_M]*/) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (TC_M - 1);     // m-row inside the tile = TMEM lane
  const int half = tid >> 7;            // which half of a pass's columns this thread reduces
  const float inf = __int_as_float(0x7f800000);
  const int nnp = (nn + 15) & ~15;
  for (int r = tid; r < nnp; r += TC_THREADS) {
    const uint4* src = reinterpret_cast<const uint4*>(n_base + (size_t)min(r, nn - 1) * 32);
    expand_row(sN, r, __ldg(src), __ldg(src + 1), r >= nn, lut);
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
    if (tid < TC_M) {
      const int mi = m0 + tid;
      const uint4* src = reinterpret_cast<const uint4*>(m_base + (size_t)min(mi, nm - 1) * 32);
      expand_row(sM, tid, __ldg(src), __ldg(src + 1), mi >= nm, lut);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA
    __syncthreads();
    if (tid == 0) issue(0);
    // two independent selection chains (even / odd columns), merged after the last pass
    uint32_t k0a = KEY_NONE, k1a = KEY_NONE, k0b = KEY_NONE, k1b = KEY_NONE;
    for (int pass = 0; pass < npass; pass++) {
      const int n0 = pass * TC_NPASS;
      const int N = min(TC_NPASS, nnp - n0);
      if (tid == 0 && pass + 1 < npass) issue(pass + 1);  // its buffer was drained before the barrier that ended pass - 1
      mbar_wait(&bars[pass & 1], phases[pass & 1]);
      phases[pass & 1] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      // this thread's share of the pass: columns [cbeg, cend) of the buffer, 16 at a time, next chunk in flight
      const int cbeg = half * (TC_NPASS / 2), cend = min(N, cbeg + TC_NPASS / 2);
      const int nvalid = min(N, nn - n0);  // columns of this pass that are real train rows
      const uint32_t taddr = tlane + (uint32_t)((pass & 1) * TC_NPASS);
      uint32_t va[16], vb[16];
      // key = (2 * hamming) << 19 | column = distance << 20 | column, with 2 * hamming = 256 - dot
      auto reduce16 = [&](const uint32_t* v, int c0) {
        const uint32_t colbase = (uint32_t)(n0 + c0);
        if (c0 + 16 <= nvalid) {
#pragma unroll
          for (int jx = 0; jx < 16; jx++) {
            const uint32_t key = ((256u - v[jx]) << (KEY_IDX_BITS - 1)) + (colbase + jx);
            if (TOP2) { if (jx & 1) top2_update(k0b, k1b, key); else top2_update(k0a, k1a, key); }
            else { if (jx & 1) k0b = min(k0b, key); else k0a = min(k0a, key); }
          }
        } else {
#pragma unroll
          for (int jx = 0; jx < 16; jx++) {
            if (c0 + jx < nvalid) {
              const uint32_t key = ((256u - v[jx]) << (KEY_IDX_BITS - 1)) + (colbase + jx);
              if (TOP2) top2_update(k0a, k1a, key); else k0a = min(k0a, key);
            }
          }
        }
      };
      if (cbeg < cend) {
        TC_LD16(va, taddr + (uint32_t)cbeg);
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          TC_WAIT16(va);
          if (c0 + 16 < cend) TC_LD16(vb, taddr + (uint32_t)(c0 + 16));
          reduce16(va, c0);
          if (c0 + 16 < cend) {
            TC_WAIT16(vb);
            if (c0 + 32 < cend) TC_LD16(va, taddr + (uint32_t)(c0 + 32));
            reduce16(vb, c0 + 16);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // this buffer is drained (pass + 2 may overwrite it); after the last pass sM is free as well
    }
    uint32_t k0 = k0a, k1 = k1a;
    if (TOP2) { top2_update(k0, k1, k0b); top2_update(k0, k1, k1b); }
    else k0 = min(k0a, k0b);
    if (half == 1) s_merge[row] = make_uint2(k0, k1);
    __syncthreads();
    const int mi = m0 + row;
    if (half == 0 && mi < nm) {
      const uint2 o = s_merge[row];
      if (TOP2) { top2_update(k0, k1, o.x); top2_update(k0, k1, o.y); }
      else k0 = min(k0, o.x);
      int2 idx;
      float2 dist;
      idx.x = (int)(k0 & KEY_IDX_MASK);
      dist.x = (float)(k0 >> KEY_IDX_BITS);
      idx.y = (TOP2 && k1 != KEY_NONE) ? (int)(k1 & KEY_IDX_MASK) : -1;
      dist.y = (TOP2 && k1 != KEY_NONE) ? (float)(k1 >> KEY_IDX_BITS) : inf;
      out_idx[mi] = idx;
      out_dist[mi] = dist;
    }
    // (the next tile's first __syncthreads orders the s_merge reads above before its writes)
  }
}

__global__ void __launch_bounds__(TC_THREADS)
knn2_tc_kernel(const uint8_t* __restrict__ a, const int32_t* __restrict__ na, int cap_a, const uint8_t* __restrict__ b,
               const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ab, Knn2Out out_ba) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sM = smem;                                   // [128][256] int8 tile of the "query" side, UMMA layout
  uint8_t* sN = smem + TC_M * TC_ROW_BYTES;             // [<= max(cap_a, cap_b) rounded to 16][256] "train" side
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_lut[16];
  __shared__ uint2 s_merge[TC_M];

  const int p = blockIdx.x;
  const int nq = na[p], nt = nb[p];
  const int tid = threadIdx.x, warp = tid >> 5;
  const float inf = __int_as_float(0x7f800000);
  const uint8_t* q_base = a + (size_t)p * cap_a * 32;
  const uint8_t* t_base = b + (size_t)p * cap_b * 32;
  if (nq == 0 || nt == 0) {  // nothing to multiply: "no match" rows on both sides (uniform exit, before any allocation)
    for (int i = tid; i < nq; i += TC_THREADS) {
      out_ab.idx[(size_t)p * cap_a + i] = make_int2(-1, -1);
      out_ab.dist[(size_t)p * cap_a + i] = make_float2(inf, inf);
    }
    for (int j = tid; j < nt; j += TC_THREADS) {
      out_ba.idx[(size_t)p * cap_b + j] = make_int2(-1, -1);
      out_ba.dist[(size_t)p * cap_b + j] = make_float2(inf, inf);
    }
    return;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  if (tid < 16) s_lut[tid] = expand_nibble(tid);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  uint32_t phases[2] = {0u, 0u};
  // forward: queries a, train b, two nearest; reverse: queries b, train a, nearest only (all the mutual checks read)
  tc_direction<true>(q_base, nq, t_base, nt, out_ab.idx + (size_t)p * cap_a, out_ab.dist + (size_t)p * cap_a, sM, sN, s_lut,
                     s_bar, phases, tmem, s_merge);
  __syncthreads();
  tc_direction<false>(t_base, nt, q_base, nq, out_ba.idx + (size_t)p * cap_b, out_ba.dist + (size_t)p * cap_b, sM, sN, s_lut,
                      s_bar, phases, tmem, s_merge);
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

static size_t knn2_tc_smem(int cap_a, int cap_b) {
  const int cap = cap_a > cap_b ? cap_a : cap_b;
  return (size_t)TC_M * TC_ROW_BYTES + (size_t)((cap + 15) & ~15) * TC_ROW_BYTES;
}

bool knn2_tc_fits(int cap_a, int cap_b) {
  // small problems are dominated by the fixed cost of a CTA (TMEM allocation, barriers): they stay on the integer pipes
  return cap_a >= 128 && cap_b >= 128 && cap_a < (1 << 19) && cap_b < (1 << 19) && knn2_tc_smem(cap_a, cap_b) <= 200 * 1024;
}

int launch_knn2_tc(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b, const int32_t* nb,
                   int cap_b, Knn2Out out_ab, Knn2Out out_ba) {
  const size_t smem = knn2_tc_smem(cap_a, cap_b);
  GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(knn2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn2_tc_kernel<<<B, TC_THREADS, smem, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
