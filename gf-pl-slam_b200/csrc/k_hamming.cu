// k_hamming.cu — K1 (knn k=2) and K2 (radius) brute-force 256-bit Hamming matching for sm_100a, plus the
// POPC / FP64 micro-benchmarks that give the roofline denominators.
//
// Replaces cv::BFMatcher::knnMatch(q,t,out,2) / radiusMatch behind StereoFrame::matchPointFeatures,
// matchLineFeatures, matchPointFeatures_radius, matchLineFeatures_radius (reference stereoFrame.cpp:1229-1257).
//
// Work decomposition (K1): grid = (query tiles, B problems, directions).  A CTA owns QPT*blockDim.x
// consecutive query rows of one problem/direction; each thread keeps QPT query rows in registers and
// streams the whole train set through a two-stage TMA-bulk ring in shared memory.  Result order follows
// cv::BFMatcher: ascending distance, ties -> ascending trainIdx (packed-key min, see gfpl_hamming.cuh).
// Algorithmic work per launch = sum over problems of nq*nt pair distances = 8 popc.b32 each.
#include <cstdlib>

#include "gfpl_hamming.cuh"
#include "gfpl_internal.h"

namespace gfpl {

constexpr int KNN_TILE_ROWS = 256;  // 8 KB per stage, 16 KB per CTA

template <int NORM, int QPT>
__global__ void __launch_bounds__(256) knn2_kernel(const uint8_t* __restrict__ a, const int32_t* __restrict__ na,
                                                   int cap_a, const uint8_t* __restrict__ b,
                                                   const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ab,
                                                   Knn2Out out_ba) {
  __shared__ __align__(128) uint8_t s_train[2 * KNN_TILE_ROWS * GFPL_DESC_BYTES];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int p = blockIdx.y;
  const bool rev = blockIdx.z != 0;
  const uint8_t* q_base = (rev ? b + (size_t)p * cap_b * 32 : a + (size_t)p * cap_a * 32);
  const uint8_t* t_base = (rev ? a + (size_t)p * cap_a * 32 : b + (size_t)p * cap_b * 32);
  const int nq = rev ? nb[p] : na[p];
  const int nt = rev ? na[p] : nb[p];
  const int cap_q = rev ? cap_b : cap_a;
  const Knn2Out out = rev ? out_ba : out_ab;

  const int q0 = blockIdx.x * (blockDim.x * QPT);
  if (q0 >= nq) return;  // whole CTA beyond this problem's queries (uniform)

  Desc256 q[QPT];
  int qi[QPT];
#pragma unroll
  for (int k = 0; k < QPT; k++) {
    qi[k] = q0 + k * blockDim.x + threadIdx.x;
    q[k] = load_desc_global(q_base + (size_t)min(qi[k], nq - 1) * 32);
  }
  uint32_t k0[QPT], k1[QPT];
#pragma unroll
  for (int k = 0; k < QPT; k++) k0[k] = k1[k] = KEY_NONE;

  TrainPipe<KNN_TILE_ROWS> pipe;
  pipe.init(s_train, s_bar, t_base, nt);
  for (int tile = 0; tile < pipe.ntiles; tile++) {
    const uint4* rows = reinterpret_cast<const uint4*>(pipe.acquire(tile));
    const int n = pipe.rows_in(tile);
    const uint32_t t0 = tile * KNN_TILE_ROWS;
#pragma unroll 4
    for (int r = 0; r < n; r++) {
      const uint4 ra = rows[2 * r], rb = rows[2 * r + 1];
#pragma unroll
      for (int k = 0; k < QPT; k++) top2_update(k0[k], k1[k], make_key(hamming256<NORM>(q[k], ra, rb), t0 + r));
    }
    pipe.release(tile);
  }

#pragma unroll
  for (int k = 0; k < QPT; k++) {
    if (qi[k] < nq) {
      int2 idx;
      float2 dist;
      idx.x = k0[k] == KEY_NONE ? -1 : (int)(k0[k] & KEY_IDX_MASK);
      idx.y = k1[k] == KEY_NONE ? -1 : (int)(k1[k] & KEY_IDX_MASK);
      dist.x = k0[k] == KEY_NONE ? __int_as_float(0x7f800000) : (float)(k0[k] >> KEY_IDX_BITS);
      dist.y = k1[k] == KEY_NONE ? __int_as_float(0x7f800000) : (float)(k1[k] >> KEY_IDX_BITS);
      out.idx[(size_t)p * cap_q + qi[k]] = idx;
      out.dist[(size_t)p * cap_q + qi[k]] = dist;
    }
  }
}

// Both directions from ONE pass over the distance matrix (d(a_i, b_j) serves a->b and b->a): the forward direction keeps
// the full k=2 result per query row as above; the reverse direction keeps only its best match — all that the mutual
// checks of the accept loops read (stereoFrame.cpp:836, stereoFrameHandler.cpp:735, mapHandler.cpp:253) — as a column
// minimum of the key (distance, query index): warp-wide redux.min, then one shared-memory atomicMin per warp and train
// row.  Ties resolve to the lowest index on both sides, as two independent knnMatch calls would.
// multi_tile: the query rows of a problem span several CTAs; column minima are merged with global atomicMin on
// out_ba.idx[].x (pre-set to 0xFFFFFFFF) and decoded by knn2_decode_kernel.
template <int NORM, int QPT>
__global__ void __launch_bounds__(256) knn2_sym_kernel(const uint8_t* __restrict__ a, const int32_t* __restrict__ na,
                                                       int cap_a, const uint8_t* __restrict__ b,
                                                       const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ab,
                                                       Knn2Out out_ba, int multi_tile) {
  extern __shared__ uint32_t s_col[];  // [cap_b]
  __shared__ __align__(128) uint8_t s_train[2 * KNN_TILE_ROWS * GFPL_DESC_BYTES];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int p = blockIdx.y;
  const uint8_t* q_base = a + (size_t)p * cap_a * 32;
  const uint8_t* t_base = b + (size_t)p * cap_b * 32;
  const int nq = na[p], nt = nb[p];
  // multi_tile bit 1: this launch only serves the problems the small tensor-core kernel leaves out (tc_small_skips)
  if ((multi_tile & 2) && !tc_small_skips(nq, nt)) return;
  multi_tile &= 1;
  const int q0 = blockIdx.x * (blockDim.x * QPT);
  const float inf = __int_as_float(0x7f800000);
  if (q0 >= nq) {  // uniform; a problem without query rows still owes the reverse direction its "no match" rows
    if (!multi_tile && blockIdx.x == 0)
      for (int j = threadIdx.x; j < nt; j += blockDim.x) {
        out_ba.idx[(size_t)p * cap_b + j] = make_int2(-1, -1);
        out_ba.dist[(size_t)p * cap_b + j] = make_float2(inf, inf);
      }
    return;
  }
  for (int j = threadIdx.x; j < nt; j += blockDim.x) s_col[j] = KEY_NONE;

  Desc256 q[QPT];
  int qi[QPT];
  uint32_t qk[QPT];  // reverse-key low part; rows past nq get a key above every valid one
#pragma unroll
  for (int k = 0; k < QPT; k++) {
    qi[k] = q0 + k * blockDim.x + threadIdx.x;
    q[k] = load_desc_global(q_base + (size_t)min(qi[k], nq - 1) * 32);
    qk[k] = qi[k] < nq ? (uint32_t)qi[k] : 0xE0000000u;
  }
  uint32_t k0[QPT], k1[QPT];
#pragma unroll
  for (int k = 0; k < QPT; k++) k0[k] = k1[k] = KEY_NONE;
  const bool leader = (threadIdx.x & 31) == 0;
  const bool warp_live = q0 + (int)(threadIdx.x & ~31u) < nq;  // slot 0 of this warp holds at least one real query

  TrainPipe<KNN_TILE_ROWS> pipe;
  pipe.init(s_train, s_bar, t_base, nt);  // (its __syncthreads also publishes s_col)
  for (int tile = 0; tile < pipe.ntiles; tile++) {
    const uint4* rows = reinterpret_cast<const uint4*>(pipe.acquire(tile));
    const int n = pipe.rows_in(tile);
    const uint32_t t0 = tile * KNN_TILE_ROWS;
    // ragged problems: a warp whose query rows all lie past nq skips the arithmetic (warp-uniform) and only keeps
    // pace at the tile barriers
    if (warp_live) {
#pragma unroll 4
      for (int r = 0; r < n; r++) {
        const uint4 ra = rows[2 * r], rb = rows[2 * r + 1];
        uint32_t rk = KEY_NONE;
#pragma unroll
        for (int k = 0; k < QPT; k++) {
          const uint32_t d = hamming256<NORM>(q[k], ra, rb);
          top2_update(k0[k], k1[k], make_key(d, t0 + r));
          rk = min(rk, (d << KEY_IDX_BITS) + qk[k]);
        }
        rk = __reduce_min_sync(0xffffffffu, rk);
        if (leader) atomicMin(&s_col[t0 + r], rk);
      }
    }
    pipe.release(tile);
  }

#pragma unroll
  for (int k = 0; k < QPT; k++) {
    if (qi[k] < nq) {
      int2 idx;
      float2 dist;
      idx.x = k0[k] == KEY_NONE ? -1 : (int)(k0[k] & KEY_IDX_MASK);
      idx.y = k1[k] == KEY_NONE ? -1 : (int)(k1[k] & KEY_IDX_MASK);
      dist.x = k0[k] == KEY_NONE ? inf : (float)(k0[k] >> KEY_IDX_BITS);
      dist.y = k1[k] == KEY_NONE ? inf : (float)(k1[k] >> KEY_IDX_BITS);
      out_ab.idx[(size_t)p * cap_a + qi[k]] = idx;
      out_ab.dist[(size_t)p * cap_a + qi[k]] = dist;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nt; j += blockDim.x) {
    const uint32_t key = s_col[j];
    if (multi_tile) {
      atomicMin(reinterpret_cast<unsigned int*>(&out_ba.idx[(size_t)p * cap_b + j].x), key);
    } else {
      out_ba.idx[(size_t)p * cap_b + j] = make_int2((int)(key & KEY_IDX_MASK), -1);
      out_ba.dist[(size_t)p * cap_b + j] = make_float2((float)(key >> KEY_IDX_BITS), inf);
    }
  }
}

// multi-tile epilogue: merged keys -> (index, distance) rows of the reverse direction
__global__ void knn2_decode_kernel(const int32_t* __restrict__ nb, int cap_b, Knn2Out out_ba) {
  const int p = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb[p]) return;
  const float inf = __int_as_float(0x7f800000);
  const uint32_t key = (uint32_t)out_ba.idx[(size_t)p * cap_b + j].x;
  out_ba.idx[(size_t)p * cap_b + j] = key == KEY_NONE ? make_int2(-1, -1) : make_int2((int)(key & KEY_IDX_MASK), -1);
  out_ba.dist[(size_t)p * cap_b + j] = key == KEY_NONE ? make_float2(inf, inf) : make_float2((float)(key >> KEY_IDX_BITS), inf);
}

// choose warps per CTA (2..8) so that tiles * warps * 32 * QPT covers `cap` with the least padding
static void pick_tiling(int cap, int qpt, int* block, int* tiles) {
  int best_w = 4, best_waste = 1 << 30;
  for (int w = 8; w >= 2; w--) {
    int slots = w * 32 * qpt;
    int t = (cap + slots - 1) / slots;
    int waste = t * slots - cap;
    if (waste < best_waste) { best_waste = waste; best_w = w; }
  }
  *block = best_w * 32;
  *tiles = (cap + best_w * 32 * qpt - 1) / (best_w * 32 * qpt);
}

template <int NORM, int QPT>
static int launch_knn2_sym(gfpl_ctx* ctx, dim3 grid, int block, size_t smem, const uint8_t* a, const int32_t* na, int cap_a,
                           const uint8_t* b, const int32_t* nb, int cap_b, Knn2Out out_ab, Knn2Out out_ba, int multi) {
  if (smem > 32 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(knn2_sym_kernel<NORM, QPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn2_sym_kernel<NORM, QPT><<<grid, block, smem, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba, multi);
  return GFPL_OK;
}

int launch_knn2(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b,
                const int32_t* nb, int cap_b, int norm, Knn2Out out_ab, Knn2Out out_ba, bool reverse_best_only) {
  if (B <= 0 || cap_a <= 0 || cap_b <= 0) return GFPL_OK;
  const bool both = out_ba.idx != nullptr;
  static const bool no_tc = getenv("GFPL_NO_TC") != nullptr;   // A/B switch for measurements
  if (both && reverse_best_only && norm == GFPL_NORM_HAMMING && !no_tc && knn2_tc_fits(cap_a, cap_b))
    return launch_knn2_tc(ctx, B, a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
  // both sides <= 128 rows (line descriptors): the persistent small tensor-core kernel takes the problems that are worth
  // a product round trip, the integer kernel below the tiny and the empty ones (a few dozen rows: counts are only known
  // on the device, so both are launched and each skips the other's problems)
  const bool tc_small = both && reverse_best_only && !no_tc && knn2_tc_small_fits(cap_a, cap_b);
  if (both && reverse_best_only && (size_t)cap_b * 4 <= 160 * 1024) {
    const int qpt = cap_a >= 256 ? 2 : 1;
    int block, tiles;
    pick_tiling(cap_a, qpt, &block, &tiles);
    const int multi = (tiles > 1 ? 1 : 0) | (tc_small ? 2 : 0);   // (tc_small implies one tile)
    if (multi & 1) GFPL_CUDA_OK(ctx, cudaMemsetAsync(out_ba.idx, 0xFF, (size_t)B * cap_b * sizeof(int2), ctx->stream));
    dim3 grid(tiles, B, 1);
    const size_t smem = (size_t)cap_b * 4;
    int rc;
    if (norm == GFPL_NORM_HAMMING2)
      rc = qpt == 2 ? launch_knn2_sym<GFPL_NORM_HAMMING2, 2>(ctx, grid, block, smem, a, na, cap_a, b, nb, cap_b, out_ab, out_ba, multi)
                    : launch_knn2_sym<GFPL_NORM_HAMMING2, 1>(ctx, grid, block, smem, a, na, cap_a, b, nb, cap_b, out_ab, out_ba, multi);
    else
      rc = qpt == 2 ? launch_knn2_sym<GFPL_NORM_HAMMING, 2>(ctx, grid, block, smem, a, na, cap_a, b, nb, cap_b, out_ab, out_ba, multi)
                    : launch_knn2_sym<GFPL_NORM_HAMMING, 1>(ctx, grid, block, smem, a, na, cap_a, b, nb, cap_b, out_ab, out_ba, multi);
    if (rc != GFPL_OK) return rc;
    GFPL_LAUNCH_OK(ctx);
    if (multi & 1) {
      knn2_decode_kernel<<<dim3((cap_b + 255) / 256, B), 256, 0, ctx->stream>>>(nb, cap_b, out_ba);
      GFPL_LAUNCH_OK(ctx);
    }
    if (tc_small) return launch_knn2_tc_small(ctx, B, a, na, cap_a, b, nb, cap_b, norm, out_ab, out_ba);
    return GFPL_OK;
  }
  const int cap_q = both ? (cap_a > cap_b ? cap_a : cap_b) : cap_a;
  const int qpt = cap_q >= 256 ? 2 : 1;
  int block, tiles;
  pick_tiling(cap_q, qpt, &block, &tiles);
  dim3 grid(tiles, B, both ? 2 : 1);
  if (norm == GFPL_NORM_HAMMING2) {
    if (qpt == 2) knn2_kernel<GFPL_NORM_HAMMING2, 2><<<grid, block, 0, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
    else knn2_kernel<GFPL_NORM_HAMMING2, 1><<<grid, block, 0, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
  } else {
    if (qpt == 2) knn2_kernel<GFPL_NORM_HAMMING, 2><<<grid, block, 0, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
    else knn2_kernel<GFPL_NORM_HAMMING, 1><<<grid, block, 0, ctx->stream>>>(a, na, cap_a, b, nb, cap_b, out_ab, out_ba);
  }
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// ---- K2: radius match, one problem, CSR output ------------------------------------------------------
// pass 1 counts, the host scans, pass 2 fills each list in ascending trainIdx and then orders it by
// (distance, trainIdx) — the canonical tie order defined in SURVEY.md §8a M3.
template <int NORM, bool FILL>
__global__ void __launch_bounds__(128) radius_kernel(const uint8_t* __restrict__ q, int nq,
                                                     const uint8_t* __restrict__ t, int nt, float radius,
                                                     int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                                                     int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
  __shared__ __align__(128) uint8_t s_train[2 * KNN_TILE_ROWS * GFPL_DESC_BYTES];
  __shared__ __align__(8) uint64_t s_bar[2];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const Desc256 qd = load_desc_global(q + (size_t)min(i, nq - 1) * 32);
  int n = 0;
  const int base = (FILL && i < nq) ? offsets[i] : 0;
  TrainPipe<KNN_TILE_ROWS> pipe;
  pipe.init(s_train, s_bar, t, nt);
  for (int tile = 0; tile < pipe.ntiles; tile++) {
    const uint4* rows = reinterpret_cast<const uint4*>(pipe.acquire(tile));
    const int nr = pipe.rows_in(tile);
    for (int r = 0; r < nr; r++) {
      uint32_t d = hamming256<NORM>(qd, rows[2 * r], rows[2 * r + 1]);
      if ((float)d <= radius && i < nq) {
        if (FILL) {
          // insertion keeps (distance, trainIdx) order: new trainIdx is the largest so far
          int pos = n;
          while (pos > 0 && out_dist[base + pos - 1] > (float)d) {
            out_dist[base + pos] = out_dist[base + pos - 1];
            out_idx[base + pos] = out_idx[base + pos - 1];
            pos--;
          }
          out_dist[base + pos] = (float)d;
          out_idx[base + pos] = tile * KNN_TILE_ROWS + r;
        }
        n++;
      }
    }
    pipe.release(tile);
  }
  if (!FILL && i < nq) counts[i] = n;
}

int launch_radius_count(gfpl_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int norm, float radius,
                        int32_t* counts) {
  if (nq <= 0) return GFPL_OK;
  dim3 grid((nq + 127) / 128);
  if (norm == GFPL_NORM_HAMMING2)
    radius_kernel<GFPL_NORM_HAMMING2, false><<<grid, 128, 0, ctx->stream>>>(q, nq, t, nt, radius, counts, nullptr, nullptr, nullptr);
  else
    radius_kernel<GFPL_NORM_HAMMING, false><<<grid, 128, 0, ctx->stream>>>(q, nq, t, nt, radius, counts, nullptr, nullptr, nullptr);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

int launch_radius_fill(gfpl_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int norm, float radius,
                       const int32_t* offsets, int32_t* out_idx, float* out_dist) {
  if (nq <= 0) return GFPL_OK;
  dim3 grid((nq + 127) / 128);
  if (norm == GFPL_NORM_HAMMING2)
    radius_kernel<GFPL_NORM_HAMMING2, true><<<grid, 128, 0, ctx->stream>>>(q, nq, t, nt, radius, nullptr, offsets, out_idx, out_dist);
  else
    radius_kernel<GFPL_NORM_HAMMING, true><<<grid, 128, 0, ctx->stream>>>(q, nq, t, nt, radius, nullptr, offsets, out_idx, out_dist);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// ---- micro-benchmarks: measured POPC and FP64 peaks (SURVEY.md §8d roofline denominators) -------------
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t* out, int iters) {
  uint32_t x0 = threadIdx.x * 2654435761u + blockIdx.x, x1 = x0 ^ 0x9e3779b9u, x2 = x0 * 3u, x3 = ~x0;
  uint32_t x4 = x0 + 77u, x5 = x1 + 1234567u, x6 = x2 ^ 0xdeadbeefu, x7 = x3 * 5u;
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  for (int i = 0; i < iters; i++) {
    // 8 independent popc per iteration; the xor with the accumulator keeps every popc live
    a0 += __popc(x0 ^ a0); a1 += __popc(x1 ^ a1); a2 += __popc(x2 ^ a2); a3 += __popc(x3 ^ a3);
    a4 += __popc(x4 ^ a4); a5 += __popc(x5 ^ a5); a6 += __popc(x6 ^ a6); a7 += __popc(x7 ^ a7);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) lop3_peak_kernel(uint32_t* out, int iters) {
  uint32_t a0 = threadIdx.x * 2654435761u + blockIdx.x, a1 = a0 ^ 0x9e3779b9u, a2 = a0 * 3u, a3 = ~a0;
  uint32_t a4 = a0 + 77u, a5 = a1 + 1234567u, a6 = a2 ^ 0xdeadbeefu, a7 = a3 * 5u;
  const uint32_t b = blockIdx.x * 747796405u + 1u, c = threadIdx.x * 2891336453u + 7u;
  for (int i = 0; i < iters; i++) {
    // 8 independent 3-input LOP3 per iteration (ALU pipe); the asm keeps them from being folded
    a0 = lop3_maj(a0, b, c); a1 = lop3_xor3(a1, b, c); a2 = lop3_maj(a2, c, b); a3 = lop3_xor3(a3, c, b);
    a4 = lop3_maj(a4, b, c); a5 = lop3_xor3(a5, b, c); a6 = lop3_maj(a6, c, b); a7 = lop3_xor3(a7, c, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int launch_measure_alu_peak(gfpl_ctx* ctx, double* lop3_tops) {
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  void* buf = nullptr;
  GFPL_CUDA_OK(ctx, cudaMalloc(&buf, (size_t)blocks * threads * sizeof(uint32_t)));
  cudaEvent_t e0, e1;
  GFPL_CUDA_OK(ctx, cudaEventCreate(&e0));
  GFPL_CUDA_OK(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    float ms;
    cudaEventRecord(e0, ctx->stream);
    lop3_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)buf, iters);
    cudaEventRecord(e1, ctx->stream);
    GFPL_LAUNCH_OK(ctx);
    GFPL_CUDA_OK(ctx, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *lop3_tops = (double)blocks * threads * iters * 8.0 / (best * 1e-3) / 1e12;
  return GFPL_OK;
}

int launch_measure_peaks(gfpl_ctx* ctx, double* popc_tops, double* fp64_tflops) {
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  void* buf = nullptr;
  GFPL_CUDA_OK(ctx, cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  GFPL_CUDA_OK(ctx, cudaEventCreate(&e0));
  GFPL_CUDA_OK(ctx, cudaEventCreate(&e1));
  float best_p = 1e30f, best_f = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    float ms;
    cudaEventRecord(e0, ctx->stream);
    popc_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)buf, iters);
    cudaEventRecord(e1, ctx->stream);
    GFPL_LAUNCH_OK(ctx);
    GFPL_CUDA_OK(ctx, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_p) best_p = ms;
    cudaEventRecord(e0, ctx->stream);
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)buf, iters);
    cudaEventRecord(e1, ctx->stream);
    GFPL_LAUNCH_OK(ctx);
    GFPL_CUDA_OK(ctx, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_f) best_f = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  const double n = (double)blocks * threads * iters * 8.0;
  *popc_tops = n / (best_p * 1e-3) / 1e12;
  *fp64_tflops = 2.0 * n / (best_f * 1e-3) / 1e12;
  return GFPL_OK;
}

}  // namespace gfpl
