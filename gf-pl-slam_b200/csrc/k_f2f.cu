// k_f2f.cu — frame-to-frame matching: the fused radius + projection-gate search of
// StereoFrameHandler::crossFrameMatching_Hybrid (reference stereoFrameHandler.cpp:451-603) and the accept
// loops of _Hybrid (lines, :607-686) and _Brute (:698-860).
//
// Hybrid points.  The reference radius-matches every previous point against the current frame, keeps
// every candidate whose projection (projectPrev3DPoint, stereoFrame.cpp:1550-1570) lies within 10 px,
// inserts them in a multimap keyed by the CURRENT index and resolves each key to the smallest distance
// (ties -> first inserted = lowest previous index).  That is a per-current-point arg-min over previous
// points under two gates, so it needs no variable-length lists: one thread per current point streams the
// previous descriptors (TMA-bulk staged) and keeps min(distance<<20 | prev_idx) over gated candidates.
// POPC-bound like K1; algorithmic work = n_prev*n_curr pair distances.
#include <cstdlib>

#include "gfpl_internal.h"
#include "gfpl_math.cuh"
#include "gfpl_tc.cuh"

namespace gfpl {

constexpr int F2F_TILE_ROWS = 256;
constexpr int F2F_THREADS = 160;  // 300 current points -> 2 CTAs x 160 (94 % lane use); 128 and 256 measured slower

__device__ __forceinline__ void project_prev_point(const gfpl_camera& cam, const double* prev_Tfw,
                                                   const double* curr_Tfw_inv, const double* P, double* uv) {
  const double v[4] = {P[0], P[1], P[2], 1.0};
  double w[4], c[4];
  matmul<4, 4, 1>(prev_Tfw, v, w);
  matmul<4, 4, 1>(curr_Tfw_inv, w, c);
  projection(cam, c, uv);
}

// out of line: its ~50 live doubles must not set the register budget of the POPC loops that share the kernel
__device__ __noinline__ void predicted_pose(const gfpl_frame_state* st, double* s_T /*[32]: prev Tfw | inverse of curr Tfw*/) {
  double Tc[16], Tci[16];
  matmul<4, 4, 4>(st->Tfw, st->DT, Tc);  // predictFramePose (stereoFrameHandler.cpp:157)
  inverse4(Tc, Tci);
  for (int i = 0; i < 16; i++) { s_T[i] = st->Tfw[i]; s_T[16 + i] = Tci[i]; }
}

// shared prologue: predicted pose, its inverse, and the projections of all previous points into smem
__device__ __forceinline__ void stage_projections(const DevStereoFrames& prev, const gfpl_frame_state* state, int p,
                                                  int n_prev, const gfpl_camera& cam, double* s_T /*[32]*/,
                                                  double2* s_proj) {
  if (threadIdx.x == 0) predicted_pose(state + p, s_T);
  __syncthreads();
  for (int q = threadIdx.x; q < n_prev; q += blockDim.x) {
    double uv[2];
    project_prev_point(cam, s_T, s_T + 16, prev.pt[(size_t)p * prev.cap_pt + q].P, uv);
    s_proj[q] = make_double2(uv[0], uv[1]);
  }
  __syncthreads();
}

// Per-problem pose pair of the projection gate, computed once by its own small kernel instead of serially by thread 0 of
// every search CTA: T32[p] = { prev Tfw (16) | inverse of the predicted curr Tfw = (Tfw * DT)^-1 (16) }.
__global__ void f2f_predict_kernel(const gfpl_frame_state* __restrict__ state, int B, double* __restrict__ T32) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < B) predicted_pose(state + p, T32 + (size_t)p * 32);
}

// path of the hybrid search for a previous descriptor within the radius: the 10 px projection gate of
// stereoFrameHandler.cpp:534-538 with projectPrev3DPoint (stereoFrame.cpp:1550-1570) evaluated for this one candidate.
// Out of line so that its fp64 registers are not live in the POPC loop.
__device__ __noinline__ uint32_t f2f_gate(uint32_t best, uint32_t key, const gfpl_camera& cam, const double* __restrict__ T32,
                                          const double* __restrict__ P, double plx, double ply, double window) {
  double uv[2];
  project_prev_point(cam, T32, T32 + 16, P, uv);
  const double dx = uv[0] - plx, dy = uv[1] - ply;
  return (!(sqrt(dx * dx + dy * dy) > window)) ? min(best, key) : best;
}

#ifndef F2F_MIN_BLOCKS
#define F2F_MIN_BLOCKS 6
#endif
constexpr int F2F_QPT = 2;  // current points per thread: halves the LDS traffic and the per-CTA projection prologue
                            // (measured: 1 per thread with 256-thread CTAs is 20 % slower)

__global__ void __launch_bounds__(F2F_THREADS, F2F_MIN_BLOCKS)
f2f_hybrid_points_kernel(DevStereoFrames prev, DevStereoFrames curr, const double* __restrict__ T32_all,
                         uint32_t* __restrict__ best_key, gfpl_camera cam, gfpl_config cfg) {
  __shared__ __align__(128) uint8_t s_train[2 * F2F_TILE_ROWS * 32];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int p = blockIdx.y;
  const int n_prev = prev.n_pt[p], n_curr = curr.n_pt[p];
  const int t0 = blockIdx.x * (blockDim.x * F2F_QPT);
  if (t0 >= n_curr) return;
  const bool active = cfg.has_points && n_prev > 0;

  uint32_t best[F2F_QPT];
  int tq[F2F_QPT];
#pragma unroll
  for (int k = 0; k < F2F_QPT; k++) { best[k] = KEY_NONE; tq[k] = t0 + k * blockDim.x + threadIdx.x; }
  if (active) {
    const double* T32 = T32_all + (size_t)p * 32;
    const gfpl_point_feature* prev_pt = prev.pt + (size_t)p * prev.cap_pt;
    Desc256 qd[F2F_QPT];
    double plx[F2F_QPT], ply[F2F_QPT];
#pragma unroll
    for (int k = 0; k < F2F_QPT; k++) {
      const int t = min(tq[k], n_curr - 1);
      qd[k] = load_desc_global(curr.pdesc + ((size_t)p * curr.cap_pt + t) * 32);
      const gfpl_point_feature& cf = curr.pt[(size_t)p * curr.cap_pt + t];
      plx[k] = cf.pl[0]; ply[k] = cf.pl[1];
    }
    // integer form of `(float)d <= radius` (radiusMatch is inclusive): d <= floor(radius)
    const float radius = (float)cfg.point_match_radius;
    const int rint = radius >= 0.f ? (int)floorf(fminf(radius, 1024.f)) : -1;
    const double window = cfg.f2f_point_window;
    // Hot loop is branch-free: count the previous rows within the radius and keep the smallest key among them.
    // Candidates are rare (non-matching descriptors sit at ~128 +- 8 bits, the radius is 50), so the projection
    // gate runs after the loop: one candidate -> gate it; several -> exact rescan of this query (rare path).
    uint32_t cnt[F2F_QPT];
#pragma unroll
    for (int k = 0; k < F2F_QPT; k++) cnt[k] = 0;
    const int warp_first = t0 + (threadIdx.x & ~31);   // first query of this warp's slot 0; slot 1 = + blockDim.x
    const int warp_slots = (F2F_QPT == 2 && warp_first + (int)blockDim.x < n_curr) ? 2 : (warp_first < n_curr ? 1 : 0);
    TrainPipe<F2F_TILE_ROWS> pipe;
    pipe.init(s_train, s_bar, prev.pdesc + (size_t)p * prev.cap_pt * 32, n_prev);
    for (int tile = 0; tile < pipe.ntiles; tile++) {
      const uint4* rows = reinterpret_cast<const uint4*>(pipe.acquire(tile));
      const int n = pipe.rows_in(tile);
      const int q0 = tile * F2F_TILE_ROWS;
      // the stereo survivors rarely fill the CTA's 2 x 160 slots: a warp whose second (or only) query slot lies past
      // n_curr skips that slot's arithmetic (warp-uniform), it only keeps pace at the tile barriers
      if (F2F_QPT == 2 && warp_slots == 2) {
#pragma unroll 4
        for (int r = 0; r < n; r++) {
          const uint4 ra = rows[2 * r], rb = rows[2 * r + 1];
#pragma unroll
          for (int k = 0; k < F2F_QPT; k++) {
            const uint32_t d = hamming256<GFPL_NORM_HAMMING>(qd[k], ra, rb);
            const bool hit = (int)d <= rint;
            cnt[k] += hit ? 1u : 0u;
            best[k] = min(best[k], hit ? make_key(d, q0 + r) : KEY_NONE);
          }
        }
      } else if (warp_slots == 1) {
#pragma unroll 4
        for (int r = 0; r < n; r++) {
          const uint32_t d = hamming256<GFPL_NORM_HAMMING>(qd[0], rows[2 * r], rows[2 * r + 1]);
          const bool hit = (int)d <= rint;
          cnt[0] += hit ? 1u : 0u;
          best[0] = min(best[0], hit ? make_key(d, q0 + r) : KEY_NONE);
        }
      }
      pipe.release(tile);
    }
#pragma unroll
    for (int k = 0; k < F2F_QPT; k++) {
      if (cnt[k] == 1) {
        best[k] = f2f_gate(KEY_NONE, best[k], cam, T32, prev_pt[best[k] & KEY_IDX_MASK].P, plx[k], ply[k], window);
      } else if (cnt[k] > 1) {
        uint32_t b = KEY_NONE;
        const uint8_t* rows_g = prev.pdesc + (size_t)p * prev.cap_pt * 32;
        for (int q = 0; q < n_prev; q++) {
          const uint4* rp = reinterpret_cast<const uint4*>(rows_g + (size_t)q * 32);
          const uint32_t d = hamming256<GFPL_NORM_HAMMING>(qd[k], __ldg(rp), __ldg(rp + 1));
          if ((int)d <= rint) b = f2f_gate(b, make_key(d, q), cam, T32, prev_pt[q].P, plx[k], ply[k], window);
        }
        best[k] = b;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < F2F_QPT; k++)
    if (tq[k] < n_curr) best_key[(size_t)p * curr.cap_pt + tq[k]] = best[k];
}

// Projection gate of one current point t, given the two smallest keys (k0 <= k1) of its row over ALL previous points:
// none / exactly one / several candidates within the radius (keys of candidates are < lim).
__device__ __forceinline__ uint32_t f2f_resolve_row(int t, uint32_t k0, uint32_t k1, uint32_t lim, int rint, int n_prev,
                                                    const uint8_t* __restrict__ prev_desc, const uint8_t* __restrict__ curr_desc,
                                                    const gfpl_point_feature* __restrict__ prev_pt,
                                                    const gfpl_point_feature* __restrict__ curr_pt, const double* __restrict__ T32,
                                                    const gfpl_camera& cam, double window) {
  uint32_t best = KEY_NONE;
  if (k0 < lim) {
    const double plx = curr_pt[t].pl[0], ply = curr_pt[t].pl[1];
    if (!(k1 < lim)) {
      best = f2f_gate(KEY_NONE, k0, cam, T32, prev_pt[k0 & KEY_IDX_MASK].P, plx, ply, window);
    } else {  // several previous points within the radius: every one of them is gated (:524-546)
      const Desc256 qd = load_desc_global(curr_desc + (size_t)t * 32);
      for (int q = 0; q < n_prev; q++) {
        const uint4* rp = reinterpret_cast<const uint4*>(prev_desc + (size_t)q * 32);
        const uint32_t d = hamming256<GFPL_NORM_HAMMING>(qd, __ldg(rp), __ldg(rp + 1));
        if ((int)d <= rint) best = f2f_gate(best, make_key(d, q), cam, T32, prev_pt[q].P, plx, ply, window);
      }
    }
  }
  return best;
}

// Tensor-core form of the search above (gfpl_tc.cuh): rows = current points, columns = previous points.  The radius
// test needs the number of candidates and the smallest key among them; the two smallest keys of a row decide both:
// none / exactly one / several candidates within the radius.  One candidate (the usual case) goes through the
// projection gate, several are rescanned exactly as in the integer kernel.
__global__ void __launch_bounds__(TC_THREADS)
f2f_hybrid_points_tc_kernel(DevStereoFrames prev, DevStereoFrames curr, const double* __restrict__ T32_all,
                            uint32_t* __restrict__ best_key, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sM = smem;
  uint8_t* sN = smem + TC_M * TC_ROW_BYTES;
  __shared__ __align__(8) TcShared S;
  const int p = blockIdx.x;
  const int n_prev = prev.n_pt[p], n_curr = curr.n_pt[p];
  uint32_t* out = best_key + (size_t)p * curr.cap_pt;
  if (!(cfg.has_points && n_prev > 0) || n_curr == 0) {
    for (int t = threadIdx.x; t < n_curr; t += TC_THREADS) out[t] = KEY_NONE;
    return;
  }
  const uint8_t* prev_desc = prev.pdesc + (size_t)p * prev.cap_pt * 32;
  const uint8_t* curr_desc = curr.pdesc + (size_t)p * curr.cap_pt * 32;
  const gfpl_point_feature* prev_pt = prev.pt + (size_t)p * prev.cap_pt;
  const gfpl_point_feature* curr_pt = curr.pt + (size_t)p * curr.cap_pt;
  const double* T32 = T32_all + (size_t)p * 32;
  const float radius = (float)cfg.point_match_radius;
  const int rint = radius >= 0.f ? (int)floorf(fminf(radius, 1024.f)) : -1;   // (float)d <= radius  <=>  d <= rint
  const uint32_t lim = (uint32_t)(rint + 1) << KEY_IDX_BITS;                  // keys of candidates are < lim
  const double window = cfg.f2f_point_window;
  // The two smallest keys of every current point are parked in shared memory; the projection gate (fp64, with dependent
  // global loads of the candidate's 3-D point) runs once at the end with all 256 threads, instead of after every query
  // tile with half of them while the other half waits at the next barrier.
  uint2* s_keys = reinterpret_cast<uint2*>(smem + tc_smem_bytes(prev.cap_pt));   // [curr.cap_pt]
  // What the gate will read is requested before the matching starts: the 3-D points of the previous frame are pulled into
  // L2 (which of them a thread needs is only known at the end: one dependent load, then an L2 hit instead of a DRAM round
  // trip with the CTA's 105 KB of shared memory waiting for it).
  for (int q = threadIdx.x; q < n_prev; q += TC_THREADS) {
    const char* a = reinterpret_cast<const char*>(&prev_pt[q]);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 16));   // P[0..2] = bytes 0..23 of the 72-byte record: may straddle a sector
  }
  const uint32_t tmem = tc_begin(S);
  uint32_t phases[2] = {0u, 0u};
  tc_direction<true>(curr_desc, n_curr, prev_desc, n_prev, sM, sN, S.bar, phases, tmem, S.merge,
                     [&](int t, uint32_t k0, uint32_t k1) { s_keys[t] = make_uint2(k0, k1); });
  tc_end(tmem);   // (CTA barrier inside: the keys are visible)
  for (int t = threadIdx.x; t < n_curr; t += TC_THREADS)
    out[t] = f2f_resolve_row(t, s_keys[t].x, s_keys[t].y, lim, rint, n_prev, prev_desc, curr_desc, prev_pt, curr_pt, T32, cam, window);
}

// Column-blocked form for previous frames too large to sit expanded in shared memory (dense-feature stress: ~2000 stereo
// survivors): grid = (blocks of previous points, problems); a CTA holds `blk` previous rows and walks all current tiles,
// leaving the two smallest keys of every current point WITHIN its block (global previous indices) in `part`
// [problem][block][curr.cap_pt]; f2f_resolve_kernel merges the blocks and applies the projection gate.
__global__ void __launch_bounds__(TC_THREADS)
f2f_hybrid_points_tc_blocked_kernel(DevStereoFrames prev, DevStereoFrames curr, int blk, uint2* __restrict__ part, gfpl_config cfg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sM = smem;
  uint8_t* sN = smem + TC_M * TC_ROW_BYTES;
  __shared__ __align__(8) TcShared S;
  const int p = blockIdx.y, j_first = blockIdx.x * blk;
  const int n_prev = cfg.has_points ? max(0, min(blk, prev.n_pt[p] - j_first)) : 0, n_curr = curr.n_pt[p];
  uint2* out = part + ((size_t)p * gridDim.x + blockIdx.x) * curr.cap_pt;
  if (n_prev == 0 || n_curr == 0) {
    for (int t = threadIdx.x; t < n_curr; t += TC_THREADS) out[t] = make_uint2(KEY_NONE, KEY_NONE);
    return;
  }
  const uint8_t* prev_desc = prev.pdesc + ((size_t)p * prev.cap_pt + j_first) * 32;
  const uint8_t* curr_desc = curr.pdesc + (size_t)p * curr.cap_pt * 32;
  const uint32_t tmem = tc_begin(S);
  uint32_t phases[2] = {0u, 0u};
  tc_direction<true>(curr_desc, n_curr, prev_desc, n_prev, sM, sN, S.bar, phases, tmem, S.merge,
                     [&](int t, uint32_t k0, uint32_t k1) {   // block-local previous rows -> global (index = low 20 bits)
                       out[t] = make_uint2(k0 + (uint32_t)j_first, k1 != KEY_NONE ? k1 + (uint32_t)j_first : KEY_NONE);
                     });
  tc_end(tmem);
}

__global__ void __launch_bounds__(128)
f2f_resolve_kernel(DevStereoFrames prev, DevStereoFrames curr, int nblk, const uint2* __restrict__ part,
                   const double* __restrict__ T32_all, uint32_t* __restrict__ best_key, gfpl_camera cam, gfpl_config cfg) {
  const int p = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= curr.n_pt[p]) return;
  const int n_prev = cfg.has_points ? prev.n_pt[p] : 0;
  uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
  for (int c = 0; c < nblk; c++) {
    const uint2 k = part[((size_t)p * nblk + c) * curr.cap_pt + t];
    top2_update(k0, k1, k.x);
    top2_update(k0, k1, k.y);
  }
  const float radius = (float)cfg.point_match_radius;
  const int rint = radius >= 0.f ? (int)floorf(fminf(radius, 1024.f)) : -1;
  const uint32_t lim = (uint32_t)(rint + 1) << KEY_IDX_BITS;
  best_key[(size_t)p * curr.cap_pt + t] =
      f2f_resolve_row(t, k0, k1, lim, rint, n_prev, prev.pdesc + (size_t)p * prev.cap_pt * 32, curr.pdesc + (size_t)p * curr.cap_pt * 32,
                      prev.pt + (size_t)p * prev.cap_pt, curr.pt + (size_t)p * curr.cap_pt, T32_all + (size_t)p * 32, cam,
                      cfg.f2f_point_window);
}

int launch_f2f_hybrid_points(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr,
                             const gfpl_frame_state* state, double* T32 /*[B][32] scratch*/, uint32_t* best_key) {
  if (prev.B <= 0 || curr.cap_pt <= 0) return GFPL_OK;
  f2f_predict_kernel<<<(prev.B + 63) / 64, 64, 0, ctx->stream>>>(state, prev.B, T32);
  GFPL_LAUNCH_OK(ctx);
  static const bool no_tc = getenv("GFPL_NO_TC") != nullptr;   // A/B switch for measurements
  const size_t tc_smem = tc_smem_bytes(prev.cap_pt) + (size_t)curr.cap_pt * sizeof(uint2);
  const bool tc_ok = !no_tc && prev.cap_pt >= 128 && curr.cap_pt >= 128 && prev.cap_pt < (1 << TC_KEY_SHIFT) && curr.cap_pt < (1 << TC_KEY_SHIFT);
  if (tc_ok && tc_smem <= 113 * 1024) {   // previous frame resident in shared memory, two CTAs per SM
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(f2f_hybrid_points_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    f2f_hybrid_points_tc_kernel<<<prev.B, TC_THREADS, tc_smem, ctx->stream>>>(prev, curr, T32, best_key, ctx->cam, ctx->cfg);
    GFPL_LAUNCH_OK(ctx);
    return GFPL_OK;
  }
  if (tc_ok) {
    int nblk, blk;
    tc_column_blocks(prev.cap_pt, &nblk, &blk);
    uint2* part = nullptr;
    const size_t part_bytes = (size_t)prev.B * nblk * curr.cap_pt * sizeof(uint2);
    GFPL_CUDA_OK(ctx, ctx->pool ? cudaMallocFromPoolAsync((void**)&part, part_bytes, ctx->pool, ctx->stream)
                                : cudaMallocAsync((void**)&part, part_bytes, ctx->stream));
    int rc = GFPL_OK;
    const size_t smem = tc_smem_bytes(blk);
    if (cudaFuncSetAttribute(f2f_hybrid_points_tc_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      rc = GFPL_E_CUDA;
    if (rc == GFPL_OK) {
      f2f_hybrid_points_tc_blocked_kernel<<<dim3(nblk, prev.B), TC_THREADS, smem, ctx->stream>>>(prev, curr, blk, part, ctx->cfg);
      f2f_resolve_kernel<<<dim3((curr.cap_pt + 127) / 128, prev.B), 128, 0, ctx->stream>>>(prev, curr, nblk, part, T32, best_key, ctx->cam,
                                                                                            ctx->cfg);
      ctx->launches += 2;
      if (cudaGetLastError() != cudaSuccess) rc = GFPL_E_CUDA;
    }
    cudaFreeAsync(part, ctx->stream);
    if (rc != GFPL_OK) set_last_error(ctx, "blocked tensor-core f2f search launch failed (%d blocks of %d rows)", nblk, blk);
    return rc;
  }
  const int per_cta = F2F_THREADS * F2F_QPT;
  dim3 grid((curr.cap_pt + per_cta - 1) / per_cta, prev.B);
  f2f_hybrid_points_kernel<<<grid, F2F_THREADS, 0, ctx->stream>>>(prev, curr, T32, best_key, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// knn-candidate variant (bestLRMatches && lrInParallel false, stereoFrameHandler.cpp:484-493): the two nearest
// current rows of every previous point are the candidates.
__global__ void __launch_bounds__(256)
f2f_hybrid_points_knn_kernel(DevStereoFrames prev, DevStereoFrames curr, const gfpl_frame_state* __restrict__ state,
                             Knn2Out pt_12, uint32_t* __restrict__ best_key, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  double2* s_proj = reinterpret_cast<double2*>(s_dyn);                               // [prev.cap_pt]
  uint32_t* s_best = reinterpret_cast<uint32_t*>(s_dyn + (size_t)prev.cap_pt * 16);  // [curr.cap_pt]
  __shared__ double s_T[32];
  const int p = blockIdx.x;
  const int n_prev = prev.n_pt[p], n_curr = curr.n_pt[p];
  for (int t = threadIdx.x; t < n_curr; t += blockDim.x) s_best[t] = KEY_NONE;
  if (cfg.has_points && n_prev > 0 && n_curr >= 2) {  // D3
    stage_projections(prev, state, p, n_prev, cam, s_T, s_proj);
    for (int q = threadIdx.x; q < n_prev; q += blockDim.x) {
      const int2 mi = pt_12.idx[(size_t)p * prev.cap_pt + q];
      const float2 md = pt_12.dist[(size_t)p * prev.cap_pt + q];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int t = k ? mi.y : mi.x;
        const float d = k ? md.y : md.x;
        const gfpl_point_feature& cf = curr.pt[(size_t)p * curr.cap_pt + t];
        const double dx = s_proj[q].x - cf.pl[0], dy = s_proj[q].y - cf.pl[1];
        if (!(sqrt(dx * dx + dy * dy) > cfg.f2f_point_window)) atomicMin(&s_best[t], make_key((uint32_t)d, q));
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_curr; t += blockDim.x) best_key[(size_t)p * curr.cap_pt + t] = s_best[t];
}

int launch_f2f_hybrid_points_knn(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr,
                                 const gfpl_frame_state* state, Knn2Out pt_12, uint32_t* best_key) {
  if (prev.B <= 0 || curr.cap_pt <= 0) return GFPL_OK;
  const size_t smem = (size_t)prev.cap_pt * 16 + (size_t)curr.cap_pt * 4;
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(f2f_hybrid_points_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  f2f_hybrid_points_knn_kernel<<<prev.B, 256, smem, ctx->stream>>>(prev, curr, state, pt_12, best_key, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// ---- accept loops ------------------------------------------------------------------------------------
#ifndef ACC_THREADS_N
#define ACC_THREADS_N 128   // 128: 0.113 ms per 8192 pairs (69 % of the HBM peak); 256: 0.143, 64: 0.114, 32: 0.125, 512: 0.231
#endif
constexpr int ACC_THREADS = ACC_THREADS_N;

__device__ __forceinline__ int block_compact_acc(bool pred, int& base, int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, pred);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < ACC_THREADS / 32; w++) {
    int c = s_warp[w];
    before += (w < warp) ? c : 0;
    total += c;
  }
  const int slot = base + before + __popc(bal & ((1u << lane) - 1));
  base += total;
  __syncthreads();
  return slot;
}

__device__ float block_select_acc(const float* v, int n, int rank, float* s_out) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float x = v[j];
    int r = 0;
    for (int k = 0; k < n; k++) {
      const float y = v[k];
      r += (y < x) || (y == x && k < j);
    }
    if (r == rank) *s_out = x;
  }
  __syncthreads();
  const float out = *s_out;
  __syncthreads();
  return out;
}

__device__ __forceinline__ void write_pt_obs(const DevStereoFrames& prev, const DevStereoFrames& curr, int p, int q,
                                             int t, gfpl_pt_obs* o) {
  const gfpl_point_feature& pf = prev.pt[(size_t)p * prev.cap_pt + q];
  const gfpl_point_feature& cf = curr.pt[(size_t)p * curr.cap_pt + t];
  o->P[0] = pf.P[0]; o->P[1] = pf.P[1]; o->P[2] = pf.P[2];
  o->pl_obs[0] = cf.pl[0]; o->pl_obs[1] = cf.pl[1];
  o->sigma2 = pf.sigma2;
  o->prev_idx = q; o->curr_idx = t;
}

__global__ void __launch_bounds__(ACC_THREADS)
f2f_accept_kernel(DevStereoFrames prev, DevStereoFrames curr, int mode, const uint32_t* __restrict__ best_key,
                  Knn2Out pt_12, Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21, DevTrackObs out, gfpl_config cfg) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  const int vcap = max(prev.cap_pt, prev.cap_ln);
  float* s_v = reinterpret_cast<float*>(s_dyn);                  // [max(prev.cap_pt, prev.cap_ln)]
  int* s_last = reinterpret_cast<int*>(s_dyn + (size_t)vcap * 4);  // [prev.cap_pt]
  __shared__ int s_warp[ACC_THREADS / 32 + 1];
  __shared__ float s_sel;
  const int p = blockIdx.x;
  const size_t op = (size_t)p * out.cap_pt, ol = (size_t)p * out.cap_ln;
  const int pt_cap = min(cfg.max_point_match_num, out.cap_pt);
  const int ln_cap = min(cfg.max_line_match_num, out.cap_ln);

  // blockIdx.y = 0: the points of pair p, 1: its lines — two independent CTAs
  // ---- points --------------------------------------------------------------------------------------
  if (blockIdx.y == 0) {
    const int n_prev = prev.n_pt[p], n_curr = curr.n_pt[p];
    int base = 0;
    if (mode == GFPL_F2F_HYBRID) {
      // resolver of stereoFrameHandler.cpp:551-599 with the semantics of SURVEY A.4 D2: ascending current index
      if (cfg.has_points && n_prev > 0 && n_curr > 0) {
        for (int q = threadIdx.x; q < n_prev; q += ACC_THREADS) s_last[q] = -1;
        __syncthreads();
        for (int t0 = 0; t0 < n_curr; t0 += ACC_THREADS) {
          const int t = t0 + threadIdx.x;
          const uint32_t key = t < n_curr ? best_key[(size_t)p * curr.cap_pt + t] : KEY_NONE;
          const bool keep = key != KEY_NONE;
          const int q = (int)(key & KEY_IDX_MASK);
          gfpl_pt_obs ob{};   // the record is read before the compaction's two CTA barriers, so its latency overlaps them
          if (keep) write_pt_obs(prev, curr, p, q, t, &ob);
          const int slot = block_compact_acc(keep, base, s_warp);
          if (keep && slot < pt_cap) {
            out.pt[op + slot] = ob;
            out.pt_inlier[op + slot] = 1;
            atomicMax(&s_last[q], t);
          }
        }
        base = min(base, pt_cap);
        __syncthreads();
        // D9: a previous point that won several keys is one PointFeature: pl_obs = its last assignment
        for (int k = threadIdx.x; k < base; k += ACC_THREADS) {
          gfpl_pt_obs& o = out.pt[op + k];
          const int tl = s_last[o.prev_idx];
          if (tl != o.curr_idx) {
            const gfpl_point_feature& cf = curr.pt[(size_t)p * curr.cap_pt + tl];
            o.pl_obs[0] = cf.pl[0]; o.pl_obs[1] = cf.pl[1];
          }
        }
      }
    } else {
      // crossFrameMatching_Brute (stereoFrameHandler.cpp:701-771)
      if (cfg.has_points && n_prev >= 2 && n_curr >= 2) {  // D3
        for (int j = threadIdx.x; j < n_prev; j += ACC_THREADS) s_v[j] = pt_12.dist[(size_t)p * prev.cap_pt + j].x;
        __syncthreads();
        const float budget = block_select_acc(s_v, n_prev, min(cfg.max_point_match_num, n_prev) - 1, &s_sel);
        const double budget_dist_th = (double)budget;
        for (int i0 = 0; i0 < n_prev; i0 += ACC_THREADS) {
          const int i = i0 + threadIdx.x;
          bool keep = false;
          int tdx = 0;
          if (i < n_prev) {
            const int2 mi = pt_12.idx[(size_t)p * prev.cap_pt + i];
            const float2 md = pt_12.dist[(size_t)p * prev.cap_pt + i];
            tdx = mi.x;
            const int rl = cfg.best_lr_matches ? pt_21.idx[(size_t)p * curr.cap_pt + tdx].x : i;
            if (!((double)md.x > cfg.budget_slack * budget_dist_th)) {
              const double dist_12 = (double)(md.x / md.y);
              keep = (i == rl) && dist_12 <= cfg.max_ratio_12_p;
            }
          }
          const int slot = block_compact_acc(keep, base, s_warp);
          if (keep && slot < pt_cap) {
            write_pt_obs(prev, curr, p, i, tdx, &out.pt[op + slot]);
            out.pt_inlier[op + slot] = 1;
          }
        }
        base = min(base, pt_cap);
      }
    }
    if (threadIdx.x == 0) out.n_pt[p] = base;
    return;
  }

  // ---- lines (stereoFrameHandler.cpp:607-686 == 775-851) -------------------------------------------------
  {
    const int n_prev = prev.n_ln[p], n_curr = curr.n_ln[p];
    int base = 0;
    if (cfg.has_lines && n_prev >= 2 && n_curr >= 2) {  // D3
      for (int j = threadIdx.x; j < n_prev; j += ACC_THREADS) {
        const float2 md = ln_12.dist[(size_t)p * prev.cap_ln + j];
        s_v[j] = fabsf((float)((double)(md.y - md.x) - cfg.mad_uninit_value));
      }
      __syncthreads();
      const float med = block_select_acc(s_v, n_prev, n_prev / 2, &s_sel);
      const double nn12_dist_th = (1.4826 * (double)med) * cfg.desc_th_l;
      for (int j = threadIdx.x; j < n_prev; j += ACC_THREADS) s_v[j] = ln_12.dist[(size_t)p * prev.cap_ln + j].x;
      __syncthreads();
      const double budget_dist_th =
          (double)block_select_acc(s_v, n_prev, min(cfg.max_line_match_num, n_prev) - 1, &s_sel);
      for (int i0 = 0; i0 < n_prev; i0 += ACC_THREADS) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        int tdx = 0;
        if (i < n_prev) {
          const int2 mi = ln_12.idx[(size_t)p * prev.cap_ln + i];
          const float2 md = ln_12.dist[(size_t)p * prev.cap_ln + i];
          tdx = mi.x;
          const int rl = cfg.best_lr_matches ? ln_21.idx[(size_t)p * curr.cap_ln + tdx].x : i;
          if (!((double)md.x > cfg.budget_slack * budget_dist_th)) {
            const double dist_12 = (double)(md.y - md.x);
            keep = (i == rl) && dist_12 > nn12_dist_th;
          }
        }
        const int slot = block_compact_acc(keep, base, s_warp);
        if (keep && slot < ln_cap) {
          const gfpl_line_feature& pf = prev.ln[(size_t)p * prev.cap_ln + i];
          const gfpl_line_feature& cf = curr.ln[(size_t)p * curr.cap_ln + tdx];
          gfpl_ln_obs o;
#pragma unroll
          for (int k = 0; k < 3; k++) { o.sP[k] = pf.sP[k]; o.eP[k] = pf.eP[k]; o.le_obs[k] = cf.le[k]; }
          o.sigma2 = pf.sigma2;
          o.prev_idx = i; o.curr_idx = tdx;
          out.ln[ol + slot] = o;
          out.ln_inlier[ol + slot] = 1;
        }
      }
      base = min(base, ln_cap);
    }
    if (threadIdx.x == 0) out.n_ln[p] = base;
  }
}

// ---- KF <-> KF accept loop of MapHandler::lookForCommonMatches (mapHandler.cpp:245-337, 377-489) ---------------------
// mutual check + ratio / MAD gap test + chi-square reprojection gate; accepted pairs in ascending kf0 index.
__global__ void __launch_bounds__(ACC_THREADS)
kf_accept_kernel(DevStereoFrames kf0, DevStereoFrames kf1, const double* __restrict__ DT_all, Knn2Out pt_12, Knn2Out pt_21,
                 Knn2Out ln_12, Knn2Out ln_21, DevTrackObs out, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  float* s_v = reinterpret_cast<float*>(s_dyn);  // [kf0.cap_ln]
  __shared__ int s_warp[ACC_THREADS / 32 + 1];
  __shared__ float s_sel;
  __shared__ double s_DT[16];
  const int p = blockIdx.x;
  const size_t op = (size_t)p * out.cap_pt, ol = (size_t)p * out.cap_ln;
  if (threadIdx.x < 16) s_DT[threadIdx.x] = DT_all[(size_t)p * 16 + threadIdx.x];
  __syncthreads();
  const double chi = sqrt(7.815);
  {
    const int n0 = kf0.n_pt[p], n1 = kf1.n_pt[p];
    int base = 0;
    if (cfg.has_points && n0 >= 2 && n1 >= 2) {  // D3
      for (int i0 = 0; i0 < n0; i0 += ACC_THREADS) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        int tdx = 0;
        if (i < n0) {
          const int2 mi = pt_12.idx[(size_t)p * kf0.cap_pt + i];
          const float2 md = pt_12.dist[(size_t)p * kf0.cap_pt + i];
          tdx = mi.x;
          const int rl = cfg.best_lr_matches ? pt_21.idx[(size_t)p * kf1.cap_pt + tdx].x : i;
          const double dist_12 = (double)(md.x / md.y);
          if (i == rl && dist_12 <= cfg.max_ratio_12_p) {
            const gfpl_point_feature& f0 = kf0.pt[(size_t)p * kf0.cap_pt + i];
            const gfpl_point_feature& f1 = kf1.pt[(size_t)p * kf1.cap_pt + tdx];
            double P_[3], uv[2];
            se3_apply(s_DT, f0.P, P_);
            projection(cam, P_, uv);
            const double dx = uv[0] - f1.pl[0], dy = uv[1] - f1.pl[1];
            keep = sqrt(dx * dx + dy * dy) * sqrt(f0.sigma2) < chi;
          }
        }
        const int slot = block_compact_acc(keep, base, s_warp);
        if (keep && slot < out.cap_pt) {
          write_pt_obs(kf0, kf1, p, i, tdx, &out.pt[op + slot]);
          out.pt_inlier[op + slot] = 1;
        }
      }
      base = min(base, out.cap_pt);
    }
    if (threadIdx.x == 0) out.n_pt[p] = base;
    __syncthreads();
  }
  {
    const int n0 = kf0.n_ln[p], n1 = kf1.n_ln[p];
    int base = 0;
    if (cfg.has_lines && n0 >= 2 && n1 >= 2) {  // D3
      for (int jx = threadIdx.x; jx < n0; jx += ACC_THREADS) {
        const float2 md = ln_12.dist[(size_t)p * kf0.cap_ln + jx];
        s_v[jx] = fabsf((float)((double)(md.y - md.x) - cfg.mad_uninit_value));
      }
      __syncthreads();
      const float med = block_select_acc(s_v, n0, n0 / 2, &s_sel);
      const double nn12_dist_th = (1.4826 * (double)med) * cfg.desc_th_l;
      for (int i0 = 0; i0 < n0; i0 += ACC_THREADS) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        int tdx = 0;
        if (i < n0) {
          const int2 mi = ln_12.idx[(size_t)p * kf0.cap_ln + i];
          const float2 md = ln_12.dist[(size_t)p * kf0.cap_ln + i];
          tdx = mi.x;
          const int rl = cfg.best_lr_matches ? ln_21.idx[(size_t)p * kf1.cap_ln + tdx].x : i;
          const double dist_12 = (double)(md.y - md.x);
          if (i == rl && dist_12 > nn12_dist_th) {
            const gfpl_line_feature& f0 = kf0.ln[(size_t)p * kf0.cap_ln + i];
            double sP_[3], eP_[3], sp[2], ep[2];
            se3_apply(s_DT, f0.sP, sP_);
            projection(cam, sP_, sp);
            se3_apply(s_DT, f0.eP, eP_);
            projection(cam, eP_, ep);
            const double e0 = f0.le[0] * sp[0] + f0.le[1] * sp[1] + f0.le[2];   // kf0's own line equation (:398)
            const double e1 = f0.le[0] * ep[0] + f0.le[1] * ep[1] + f0.le[2];
            keep = sqrt(e0 * e0 + e1 * e1) * sqrt(f0.sigma2) < chi;
          }
        }
        const int slot = block_compact_acc(keep, base, s_warp);
        if (keep && slot < out.cap_ln) {
          const gfpl_line_feature& f0 = kf0.ln[(size_t)p * kf0.cap_ln + i];
          const gfpl_line_feature& f1 = kf1.ln[(size_t)p * kf1.cap_ln + tdx];
          gfpl_ln_obs o;
#pragma unroll
          for (int k = 0; k < 3; k++) { o.sP[k] = f0.sP[k]; o.eP[k] = f0.eP[k]; o.le_obs[k] = f1.le[k]; }
          o.sigma2 = f0.sigma2;
          o.prev_idx = i; o.curr_idx = tdx;
          out.ln[ol + slot] = o;
          out.ln_inlier[ol + slot] = 1;
        }
      }
      base = min(base, out.cap_ln);
    }
    if (threadIdx.x == 0) out.n_ln[p] = base;
  }
}

int launch_kf_accept(gfpl_ctx* ctx, const DevStereoFrames& kf0, const DevStereoFrames& kf1, const double* DT, Knn2Out pt_12,
                     Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21, DevTrackObs out) {
  if (kf0.B <= 0) return GFPL_OK;
  const size_t smem = (size_t)(kf0.cap_ln > 0 ? kf0.cap_ln : 1) * 4;
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(kf_accept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kf_accept_kernel<<<kf0.B, ACC_THREADS, smem, ctx->stream>>>(kf0, kf1, DT, pt_12, pt_21, ln_12, ln_21, out, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

int launch_f2f_accept(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr, int mode,
                      const uint32_t* best_key, Knn2Out pt_12, Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21,
                      DevTrackObs out) {
  if (prev.B <= 0) return GFPL_OK;
  const int vcap = prev.cap_pt > prev.cap_ln ? prev.cap_pt : prev.cap_ln;
  const size_t smem = (size_t)(vcap > 0 ? vcap : 1) * 4 + (size_t)(prev.cap_pt > 0 ? prev.cap_pt : 1) * 4;
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(f2f_accept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  f2f_accept_kernel<<<dim3(prev.B, 2), ACC_THREADS, smem, ctx->stream>>>(prev, curr, mode, best_key, pt_12, pt_21, ln_12, ln_21,
                                                               out, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
