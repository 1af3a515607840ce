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
#include "gfpl_internal.h"
#include "gfpl_tc.cuh"

namespace gfpl {

__global__ void __launch_bounds__(TC_THREADS)
knn2_tc_kernel(const uint8_t* __restrict__ a, const int32_t* __restrict__ na, int cap_a, const uint8_t* __restrict__ b,
               const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ab, Knn2Out out_ba) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sM = smem;                                   // [128][256] int8 tile of the "query" side, UMMA layout
  uint8_t* sN = smem + TC_M * TC_ROW_BYTES;             // [<= max(cap_a, cap_b) rounded to 16][256] "train" side
  __shared__ __align__(8) TcShared S;

  const int p = blockIdx.x;
  const int nq = na[p], nt = nb[p];
  const int tid = threadIdx.x;
  const float inf = __int_as_float(0x7f800000);
  const uint8_t* q_base = a + (size_t)p * cap_a * 32;
  const uint8_t* t_base = b + (size_t)p * cap_b * 32;
  int2* idx_ab = out_ab.idx + (size_t)p * cap_a;
  float2* dist_ab = out_ab.dist + (size_t)p * cap_a;
  int2* idx_ba = out_ba.idx + (size_t)p * cap_b;
  float2* dist_ba = out_ba.dist + (size_t)p * cap_b;
  if (nq == 0 || nt == 0) {  // nothing to multiply: "no match" rows on both sides (uniform exit, before any allocation)
    for (int i = tid; i < nq; i += TC_THREADS) { idx_ab[i] = make_int2(-1, -1); dist_ab[i] = make_float2(inf, inf); }
    for (int j = tid; j < nt; j += TC_THREADS) { idx_ba[j] = make_int2(-1, -1); dist_ba[j] = make_float2(inf, inf); }
    return;
  }
  const uint32_t tmem = tc_begin(S);
  uint32_t phases[2] = {0u, 0u};
  // forward: queries a, train b, two nearest; reverse: queries b, train a, nearest only (all the mutual checks read)
  tc_direction<true>(q_base, nq, t_base, nt, sM, sN, S.lut, S.bar, phases, tmem, S.merge,
                     [&](int i, uint32_t k0, uint32_t k1) {
                       idx_ab[i] = make_int2((int)(k0 & KEY_IDX_MASK), k1 != KEY_NONE ? (int)(k1 & KEY_IDX_MASK) : -1);
                       dist_ab[i] = make_float2((float)(k0 >> KEY_IDX_BITS), k1 != KEY_NONE ? (float)(k1 >> KEY_IDX_BITS) : inf);
                     });
  __syncthreads();
  tc_direction<false>(t_base, nt, q_base, nq, sM, sN, S.lut, S.bar, phases, tmem, S.merge,
                      [&](int j, uint32_t k0, uint32_t) {
                        idx_ba[j] = make_int2((int)(k0 & KEY_IDX_MASK), -1);
                        dist_ba[j] = make_float2((float)(k0 >> KEY_IDX_BITS), inf);
                      });
  tc_end(tmem);
}

bool knn2_tc_fits(int cap_a, int cap_b) {
  // small problems are dominated by the fixed cost of a CTA (TMEM allocation, barriers): they stay on the integer pipes
  return cap_a >= 128 && cap_b >= 128 && cap_a < (1 << 19) && cap_b < (1 << 19) &&
         tc_smem_bytes(cap_a > cap_b ? cap_a : cap_b) <= 200 * 1024;
}

int launch_knn2_tc(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b, const int32_t* nb,
                   int cap_b, Knn2Out out_ab, Knn2Out out_ba) {
  const size_t smem = tc_smem_bytes(cap_a > cap_b ? cap_a : cap_b);
  GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(knn2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn2_tc_kernel<<<B, TC_THREADS, smem, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
