// k_stereo.cu — K3 (stereo half): accept loops that turn knn k=2 results into stereo_pt / stereo_ls, and
// StereoFrame::estimateStereoUncertainty.
//
// Replaces the loops of StereoFrame::extractInitialStereoFeatures (reference stereoFrame.cpp:206-245,
// 287-332) and extractStereoFeatures_Brute (:829-878, :919-1012; line half identical to _ORBSLAM :670-763),
// lineDescriptorMAD (:1287-1313), lineSegmentOverlapStereo (:1343-1371), estimateStereoUncertainty
// (:1448-1484) with getCovMat2D_3D / getJacob2D_3D (:1375-1392, 1434-1446).
//
// Two CTAs per stereo frame (points, lines).  Accepted features keep the reference's output order (ascending left index):
// a block-wide ordered compaction (ballot + warp totals) assigns the output slot.  HBM-/latency-bound,
// O(N) per frame: algorithmic bytes = knn results (16 B/query, two directions) + key points (16 B) +
// key lines (32 B) in, feature records (72 B / 152 B) + 32 B descriptor out.
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

#ifndef ACCEPT_THREADS_N
#define ACCEPT_THREADS_N 64    // 64: 0.189 ms per 8192 frames; 128: 0.192, 32: 0.191, 96: 0.212, 256: 0.285
#endif
constexpr int ACCEPT_THREADS = ACCEPT_THREADS_N;
#ifndef ACCEPT_UNROLL_N
#define ACCEPT_UNROLL_N 3
#endif
constexpr int SIG_LEVELS = 16;      // pyramid levels with a precomputed sigma2 per CTA
constexpr int ACCEPT_UNROLL = ACCEPT_UNROLL_N;    // stereo points: elements per thread with their loads in flight together

// ordered block compaction: returns the output slot of this thread's element (valid when pred), advances
// `base` (uniform across the CTA) by the number of accepted elements of this round.
__device__ __forceinline__ int block_compact(bool pred, int& base, int* s_warp /*[ACCEPT_THREADS/32 + 1]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, pred);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < ACCEPT_THREADS / 32; w++) {
    int c = s_warp[w];
    before += (w < warp) ? c : 0;
    total += c;
  }
  const int slot = base + before + __popc(bal & ((1u << lane) - 1));
  base += total;
  __syncthreads();
  return slot;
}

// element of rank `rank` (0-based) of v[0..n): what `sort(v); v[rank]` reads.  Every thread returns it.
__device__ float block_select(const float* v, int n, int rank, float* s_out) {
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

// PointFeature / LineFeature sigma2 (stereoFeatures.cpp:41-48, 93-101)
__device__ __forceinline__ double level_sigma2(int level, double scale) {
  double sigma2 = 1.0;
  for (int i = 0; i < level + 1; i++) sigma2 *= scale;
  return 1.0 / (sigma2 * sigma2);
}

// StereoFrame::lineSegmentOverlapStereo (stereoFrame.cpp:1343-1371)
__device__ __forceinline__ double line_segment_overlap_stereo(double spl_obs, double epl_obs, double spl_proj,
                                                              double epl_proj) {
  const double sln = fmin(spl_obs, epl_obs), eln = fmax(spl_obs, epl_obs);
  const double spn = fmin(spl_proj, epl_proj), epn = fmax(spl_proj, epl_proj);
  const double length = eln - spn;
  double overlap;
  if ((epn < sln) || (spn > eln)) overlap = 0.0;
  else {
    if ((epn > eln) && (spn < sln)) overlap = eln - sln;
    else overlap = fmin(eln, epn) - fmax(sln, spn);
  }
  if (length > (double)0.01f) overlap = overlap / length;
  else overlap = 0.0;
  return overlap;
}

// analytic endpoint covariance and its largest eigenvalue (stereoFrame.cpp:956-998)
__device__ __noinline__ double endpoint_cov_max_eig(const gfpl_camera& cam, double px, double py, double disp) {
  const double f = cam.fx;
  const double px_hat = px - cam.cx, py_hat = py - cam.cy;
  const double disp2 = disp * disp;
  double c[9];
  c[0] = disp2 + 2.0 * px_hat * px_hat;
  c[1] = 2.0 * px_hat * py_hat;
  c[2] = 2.0 * f * px_hat;
  c[4] = disp2 + 2.0 * py_hat * py_hat;
  c[5] = 2.0 * f * py_hat;
  c[8] = 2.0 * f * f;
  c[3] = c[1]; c[6] = c[2]; c[7] = c[5];
  const double d4 = disp2 * disp2;
  for (int i = 0; i < 9; i++) c[i] = c[i] * cam.b * cam.b / d4;
  double ev[3];
  sym_eigenvalues_unrolled<3>(c, ev);   // unsorted diagonal
  // ascending insertion sort of three values, written out (same comparisons as sym_eigenvalues, NaN included)
  if (ev[0] > ev[1]) { const double t = ev[0]; ev[0] = ev[1]; ev[1] = t; }
  if (ev[1] > ev[2]) {
    const double v = ev[2];
    ev[2] = ev[1];
    if (ev[0] > v) { ev[1] = ev[0]; ev[0] = v; } else ev[1] = v;
  }
  return ev[2];
}

// The endpoint covariance is  k (2 v v^T + diag(d^2, d^2, 0))  with v = (x - cx, y - cy, f), d the disparity and
// k = b^2 / d^4, so by Weyl's inequalities its largest eigenvalue lies in [2 k |v|^2, k (2 |v|^2 + d^2)] — an interval of
// relative width d^2 / (2 |v|^2) < 1e-4.  +1: surely below `th`, -1: surely not below, 0: inside the interval (or not
// finite) — only then is the Jacobi iteration of the reference run.  Margins of 1e-9 cover the rounding of both sides.
__device__ __forceinline__ int endpoint_cov_screen(const gfpl_camera& cam, double px, double py, double disp, double th) {
  const double x = px - cam.cx, y = py - cam.cy, d2 = disp * disp;
  const double k = cam.b * cam.b / (d2 * d2);
  const double v2 = x * x + y * y + cam.fx * cam.fx;
  const double lo = 2.0 * k * v2, hi = k * (2.0 * v2 + d2);
  if (hi < th * (1.0 - 1e-9)) return 1;
  if (lo > th * (1.0 + 1e-9) && lo < 1e300) return -1;
  return 0;
}

#ifndef ACCEPT_MIN_CTAS
#define ACCEPT_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(ACCEPT_THREADS, ACCEPT_MIN_CTAS)
stereo_accept_kernel(DevRawFrames raw, int variant, Knn2Out pt_lr, Knn2Out pt_rl, Knn2Out ln_lr, Knn2Out ln_rl,
                     DevStereoFrames out, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ float s_v[];  // [raw.cap_ln] MAD scratch
  __shared__ int s_warp[ACCEPT_THREADS / 32 + 1];
  __shared__ float s_sel;
  const int p = blockIdx.x;
  const size_t rp = (size_t)p * raw.cap_pt, rl_ = (size_t)p * raw.cap_ln;
  const size_t op = (size_t)p * out.cap_pt, ol = (size_t)p * out.cap_ln;
  // sigma2 of the pyramid levels, once per CTA by the same sequence of products (the loop was 12 % of the kernel's
  // instructions: up to eight dependent multiplications and a division per accepted feature)
  __shared__ double s_sigma2[SIG_LEVELS];
  if (threadIdx.x < SIG_LEVELS) s_sigma2[threadIdx.x] = level_sigma2(threadIdx.x, blockIdx.y == 0 ? cfg.orb_scale_factor : cfg.lsd_scale);
  __syncthreads();

  // blockIdx.y = 0: the points of frame p, 1: its lines — two independent CTAs (the halves share nothing, and a kernel
  // bound by the latency of dependent global loads wants many small CTAs in flight)
  // ---- points (stereoFrame.cpp:206-245 / 829-878) -------------------------------------------------
  if (blockIdx.y == 0) {
    const int n_l = raw.n_pt_l[p], n_r = raw.n_pt_r[p];
    int base = 0;
    if (cfg.has_points && n_l >= 2 && n_r >= 2) {  // D3
      // ACCEPT_UNROLL elements per thread and round: all their first-level loads (knn record, left key point) go out
      // together, then all second-level loads (reverse record, right key point) — the loads form two levels instead of
      // four per element, and ACCEPT_UNROLL elements share each level's latency (ncu: long-scoreboard 11 of 18 stall cycles
      // per issue with one element per round)
      for (int i0 = 0; i0 < n_l; i0 += ACCEPT_THREADS * ACCEPT_UNROLL) {
        int2 mi[ACCEPT_UNROLL];
        float2 md[ACCEPT_UNROLL];
        gfpl_keypoint kl[ACCEPT_UNROLL], kr[ACCEPT_UNROLL];
        int rl_tdx[ACCEPT_UNROLL];
        uint4 dlo[ACCEPT_UNROLL], dhi[ACCEPT_UNROLL];   // the left descriptor: requested with the first level, stored after the compaction
#pragma unroll
        for (int u = 0; u < ACCEPT_UNROLL; u++) {
          const int i = i0 + u * ACCEPT_THREADS + threadIdx.x;
          if (i < n_l) {
            mi[u] = pt_lr.idx[rp + i]; md[u] = pt_lr.dist[rp + i]; kl[u] = raw.pt_l[rp + i];
            const uint4* sd = reinterpret_cast<const uint4*>(raw.pdesc_l + (rp + i) * 32);
            dlo[u] = __ldg(sd); dhi[u] = __ldg(sd + 1);
          } else { mi[u] = make_int2(0, 0); md[u] = make_float2(0.f, 0.f); kl[u] = gfpl_keypoint{}; dlo[u] = dhi[u] = make_uint4(0, 0, 0, 0); }
        }
#pragma unroll
        for (int u = 0; u < ACCEPT_UNROLL; u++) {
          const int i = i0 + u * ACCEPT_THREADS + threadIdx.x;
          if (i < n_l) {
            rl_tdx[u] = cfg.best_lr_matches ? pt_rl.idx[rp + mi[u].x].x : i;
            kr[u] = raw.pt_r[rp + mi[u].x];
          } else { rl_tdx[u] = -1; kr[u] = gfpl_keypoint{}; }
        }
#pragma unroll
        for (int u = 0; u < ACCEPT_UNROLL; u++) {
          const int i = i0 + u * ACCEPT_THREADS + threadIdx.x;
          bool keep = false;
          float disp_sub = 0.f;
          const int tdx = mi[u].x;
          if (i < n_l) {
            const double dist_12 = (double)(md[u].x / md[u].y);  // float / float, widened (stereoFrame.cpp:219,842)
            if (i == rl_tdx[u] && dist_12 <= cfg.max_ratio_12_p) {
              if ((double)fabsf(kl[u].y - kr[u].y) <= cfg.max_dist_epip) {
                const double disp_ = (double)(kl[u].x - kr[u].x);
                if (disp_ >= cfg.min_disp) {
                  disp_sub = (float)disp_;
                  keep = disp_sub >= 0.f;
                }
              }
            }
          }
          if (i0 + u * ACCEPT_THREADS >= n_l) break;   // CTA-uniform: no element of this sub-round exists
          const int slot = block_compact(keep, base, s_warp);
          if (keep) {
            gfpl_point_feature f;
            f.pl[0] = kl[u].x; f.pl[1] = kl[u].y;
            back_projection(cam, f.pl[0], f.pl[1], (double)disp_sub, f.P);
            f.disp = disp_sub;
            f.level = kl[u].octave;
            f.sigma2 = ((unsigned)f.level < (unsigned)SIG_LEVELS) ? s_sigma2[f.level] : level_sigma2(f.level, cfg.orb_scale_factor);
            f.idx = slot; f.src_l = i; f.src_r = tdx;
            out.pt[op + slot] = f;
            uint4* d = reinterpret_cast<uint4*>(out.pdesc + (op + slot) * 32);
            d[0] = dlo[u]; d[1] = dhi[u];
          }
        }
      }
    }
    if (threadIdx.x == 0) out.n_pt[p] = base;
    return;
  }

  // ---- lines (stereoFrame.cpp:248-334 / 881-1014) ---------------------------------------------------
  {
    const int n_l = raw.n_ln_l[p], n_r = raw.n_ln_r[p];
    int base = 0;
    if (cfg.has_lines && n_l >= 2 && n_r >= 2) {  // D3
      // lineDescriptorMAD: 1.4826 * element[n/2] of ascending fabsf(d1 - d0 - U)  (U: SURVEY A.4 D1)
      for (int j = threadIdx.x; j < n_l; j += ACCEPT_THREADS) {
        const float2 md = ln_lr.dist[rl_ + j];
        s_v[j] = fabsf((float)((double)(md.y - md.x) - cfg.mad_uninit_value));
      }
      __syncthreads();
      const float med = block_select(s_v, n_l, n_l / 2, &s_sel);
      const double nn12_dist_th = (1.4826 * (double)med) * cfg.desc_th_l;
      const int n_matches = cfg.best_lr_matches ? min(n_l, n_r) : n_l;  // D4
      for (int i0 = 0; i0 < n_matches; i0 += ACCEPT_THREADS) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        gfpl_line_feature f{};
        uint4 dlo = make_uint4(0, 0, 0, 0), dhi = dlo;   // the left descriptor: requested now, stored after the compaction
        if (i < n_matches) {
          const uint4* sd = reinterpret_cast<const uint4*>(raw.ldesc_l + (rl_ + i) * 32);
          dlo = __ldg(sd); dhi = __ldg(sd + 1);
          const int tdx = ln_lr.idx[rl_ + i].x;
          const float2 md = ln_lr.dist[rl_ + i];
          const gfpl_keyline L = raw.ln_l[rl_ + i];      // (two load levels, as for the points)
          const int rl_tdx = cfg.best_lr_matches ? ln_rl.idx[rl_ + tdx].x : i;
          const gfpl_keyline R = raw.ln_r[rl_ + tdx];
          const double dist_12 = (double)(md.y - md.x);
          if (i == rl_tdx && dist_12 > nn12_dist_th) {
            const double sp_l[3] = {L.startPointX, L.startPointY, 1.0};
            const double ep_l[3] = {L.endPointX, L.endPointY, 1.0};
            double le_l[3] = {sp_l[1] * ep_l[2] - sp_l[2] * ep_l[1], sp_l[2] * ep_l[0] - sp_l[0] * ep_l[2],
                              sp_l[0] * ep_l[1] - sp_l[1] * ep_l[0]};
            const double nrm = sqrt(le_l[0] * le_l[0] + le_l[1] * le_l[1]);
            le_l[0] = le_l[0] / nrm; le_l[1] = le_l[1] / nrm; le_l[2] = le_l[2] / nrm;
            const double sp_r[3] = {R.startPointX, R.startPointY, 1.0};
            const double ep_r[3] = {R.endPointX, R.endPointY, 1.0};
            const double le_r[3] = {sp_r[1] * ep_r[2] - sp_r[2] * ep_r[1], sp_r[2] * ep_r[0] - sp_r[0] * ep_r[2],
                                    sp_r[0] * ep_r[1] - sp_r[1] * ep_r[0]};
            const double overlap = line_segment_overlap_stereo(sp_l[1], ep_l[1], sp_r[1], ep_r[1]);
            const double spr0 = -(le_r[2] + le_r[1] * (double)L.startPointY) / le_r[0];
            const double epr0 = -(le_r[2] + le_r[1] * (double)L.endPointY) / le_r[0];
            const double disp_s = (double)L.startPointX - spr0;
            const double disp_e = (double)L.endPointX - epr0;
            const double horiz = variant == GFPL_STEREO_INITIAL ? le_r[0] : le_l[0];  // :316 vs :949
            if (disp_s >= cfg.min_disp && disp_e >= cfg.min_disp && (double)fabsf((float)horiz) > cfg.line_horiz_th &&
                overlap > cfg.stereo_overlap_th) {
              keep = true;
              if (variant != GFPL_STEREO_INITIAL) {
                const int ss = endpoint_cov_screen(cam, sp_l[0], sp_l[1], disp_s, cfg.line_cov_th);
                const int se = endpoint_cov_screen(cam, ep_l[0], ep_l[1], disp_e, cfg.line_cov_th);
                if (ss < 0 || se < 0) keep = false;          // one end point alone puts the maximum at or above the threshold
                else if (ss > 0 && se > 0) keep = true;
                else {
                  const double max_eig = fmax(endpoint_cov_max_eig(cam, sp_l[0], sp_l[1], disp_s),
                                              endpoint_cov_max_eig(cam, ep_l[0], ep_l[1], disp_e));
                  keep = max_eig < cfg.line_cov_th;
                }
              }
              if (keep) {
                f.spl[0] = sp_l[0]; f.spl[1] = sp_l[1]; f.epl[0] = ep_l[0]; f.epl[1] = ep_l[1];
                f.sdisp = disp_s; f.edisp = disp_e;
                back_projection(cam, sp_l[0], sp_l[1], disp_s, f.sP);
                back_projection(cam, ep_l[0], ep_l[1], disp_e, f.eP);
                f.le[0] = le_l[0]; f.le[1] = le_l[1]; f.le[2] = le_l[2];
                f.angle = L.angle;
                f.level = L.octave;
                f.sigma2 = ((unsigned)f.level < (unsigned)SIG_LEVELS) ? s_sigma2[f.level] : level_sigma2(f.level, cfg.lsd_scale);
                f.src_l = i; f.src_r = tdx;
              }
            }
          }
        }
        const int slot = block_compact(keep, base, s_warp);
        if (keep) {
          f.idx = variant == GFPL_STEREO_INITIAL ? slot : -1;
          out.ln[ol + slot] = f;
          uint4* d = reinterpret_cast<uint4*>(out.ldesc + (ol + slot) * 32);
          d[0] = dlo; d[1] = dhi;
        }
      }
    }
    if (threadIdx.x == 0) out.n_ln[p] = base;
  }
}

int launch_stereo_accept(gfpl_ctx* ctx, const DevRawFrames& raw, int variant, Knn2Out pt_lr, Knn2Out pt_rl,
                         Knn2Out ln_lr, Knn2Out ln_rl, DevStereoFrames out, bool lines_only) {
  if (raw.B <= 0) return GFPL_OK;
  const size_t smem = (size_t)(raw.cap_ln > 0 ? raw.cap_ln : 1) * sizeof(float);
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(stereo_accept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gfpl_config cfg = ctx->cfg;
  if (lines_only) cfg.has_points = 0;  // the point part then only writes n_pt = 0 (overwritten by the caller's point kernel)
  stereo_accept_kernel<<<dim3(raw.B, 2), ACCEPT_THREADS, smem, ctx->stream>>>(raw, variant, pt_lr, pt_rl, ln_lr, ln_rl, out,
                                                                      ctx->cam, cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// ---- C1: estimateStereoUncertainty (stereoFrame.cpp:1448-1484): cov_mat_2d_3d / line_cov_from_feature live in gfpl_math.cuh
// (the line cut evaluates them for the matched lines itself in the tracking step) ----------------------------------------
__global__ void __launch_bounds__(128) stereo_uncertainty_kernel(DevStereoFrames fr, gfpl_line_cov* out,
                                                                 gfpl_camera cam, gfpl_config cfg) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= fr.n_ln[p]) return;
  gfpl_line_cov c;
  line_cov_from_feature(cam, cfg, fr.ln[(size_t)p * fr.cap_ln + i], c);
  out[(size_t)p * fr.cap_ln + i] = c;
}

// slim previous-frame records -> full feature records (fields the tracking step does not read are zeroed)
__global__ void __launch_bounds__(128) expand_prev_kernel(gfpl_prev_frames slim, DevStereoFrames fr) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < slim.n_pt[p]) {
    const gfpl_prev_point s = slim.pt[(size_t)p * slim.cap_pt + i];
    gfpl_point_feature f{};
    f.P[0] = s.P[0]; f.P[1] = s.P[1]; f.P[2] = s.P[2];
    f.sigma2 = s.sigma2;
    f.idx = i; f.src_l = f.src_r = -1;
    fr.pt[(size_t)p * fr.cap_pt + i] = f;
  }
  if (i < slim.n_ln[p]) {
    const gfpl_prev_line s = slim.ln[(size_t)p * slim.cap_ln + i];
    gfpl_line_feature f{};
#pragma unroll
    for (int k = 0; k < 3; k++) { f.sP[k] = s.sP[k]; f.eP[k] = s.eP[k]; }
    f.spl[0] = s.spl[0]; f.spl[1] = s.spl[1]; f.epl[0] = s.epl[0]; f.epl[1] = s.epl[1];
    f.sdisp = s.sdisp; f.edisp = s.edisp;
    f.le[0] = s.le0;
    f.sigma2 = s.sigma2;
    f.idx = -1; f.src_l = f.src_r = -1;
    fr.ln[(size_t)p * fr.cap_ln + i] = f;
  }
  if (i == 0) { fr.n_pt[p] = slim.n_pt[p]; fr.n_ln[p] = slim.n_ln[p]; }
}

int launch_expand_prev(gfpl_ctx* ctx, const gfpl_prev_frames& slim, DevStereoFrames fr) {
  if (slim.B <= 0) return GFPL_OK;
  const int cap = slim.cap_pt > slim.cap_ln ? slim.cap_pt : slim.cap_ln;
  dim3 grid((cap + 127) / 128, slim.B);
  expand_prev_kernel<<<grid, 128, 0, ctx->stream>>>(slim, fr);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// compact step input -> the records the kernels read.  blockIdx.y = pair, blockIdx.z selects the record type.
__global__ void __launch_bounds__(128) expand_compact_kernel(gfpl_compact_input c, DevRawFrames raw, DevStereoFrames prev,
                                                             gfpl_camera cam, gfpl_config cfg) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.z == 0) {  // key points, left and right
    if (i < c.cap_pt) {
      const size_t o = (size_t)p * c.cap_pt + i;
      const gfpl_keypoint8 a = c.pt_l[o], b = c.pt_r[o];
      const_cast<gfpl_keypoint*>(raw.pt_l)[o] = gfpl_keypoint{a.x, a.y, (int32_t)c.pt_oct_l[o], -1};
      const_cast<gfpl_keypoint*>(raw.pt_r)[o] = gfpl_keypoint{b.x, b.y, (int32_t)c.pt_oct_r[o], -1};
    }
  } else if (blockIdx.z == 1) {  // key lines
    if (i < c.cap_ln) {
      const size_t o = (size_t)p * c.cap_ln + i;
      const gfpl_keyline20 a = c.ln_l[o], b = c.ln_r[o];
      const_cast<gfpl_keyline*>(raw.ln_l)[o] = gfpl_keyline{a.startPointX, a.startPointY, a.endPointX, a.endPointY, a.angle, (int32_t)c.ln_oct_l[o], 0.f, 0.f};
      const_cast<gfpl_keyline*>(raw.ln_r)[o] = gfpl_keyline{b.startPointX, b.startPointY, b.endPointX, b.endPointY, b.angle, (int32_t)c.ln_oct_r[o], 0.f, 0.f};
    }
  } else if (blockIdx.z == 2) {  // previous stereo points: P and sigma2 by the expressions of the stereo stage
    if (i < prev.n_pt[p]) {
      const size_t o = (size_t)p * c.prev_cap_pt + i;
      const gfpl_prev_point12 s = c.prev_pt[o];
      const int level = c.prev_pt_level[o];
      gfpl_point_feature f{};
      f.pl[0] = s.x; f.pl[1] = s.y;
      f.disp = s.disp;
      back_projection(cam, f.pl[0], f.pl[1], (double)s.disp, f.P);
      f.level = level;
      f.sigma2 = level_sigma2(level, cfg.orb_scale_factor);
      f.idx = i; f.src_l = f.src_r = -1;
      prev.pt[(size_t)p * prev.cap_pt + i] = f;
    }
  } else {  // previous stereo lines: sP / eP, le[0] and sigma2 by the expressions of the stereo accept loop (:924-927, :1002)
    if (i < prev.n_ln[p]) {
      const size_t o = (size_t)p * c.prev_cap_ln + i;
      const gfpl_prev_line32 s = c.prev_ln[o];
      const int level = c.prev_ln_level[o];
      gfpl_line_feature f{};
      const double sp_l[3] = {s.spl[0], s.spl[1], 1.0};
      const double ep_l[3] = {s.epl[0], s.epl[1], 1.0};
      const double l0 = sp_l[1] * ep_l[2] - sp_l[2] * ep_l[1], l1 = sp_l[2] * ep_l[0] - sp_l[0] * ep_l[2];
      const double nrm = sqrt(l0 * l0 + l1 * l1);
      f.le[0] = l0 / nrm;
      f.spl[0] = sp_l[0]; f.spl[1] = sp_l[1]; f.epl[0] = ep_l[0]; f.epl[1] = ep_l[1];
      f.sdisp = s.sdisp; f.edisp = s.edisp;
      back_projection(cam, sp_l[0], sp_l[1], s.sdisp, f.sP);
      back_projection(cam, ep_l[0], ep_l[1], s.edisp, f.eP);
      f.level = level;
      f.sigma2 = level_sigma2(level, cfg.lsd_scale);
      f.idx = -1; f.src_l = f.src_r = -1;
      prev.ln[(size_t)p * prev.cap_ln + i] = f;
    }
  }
}

// with_prev = false: only the current frame's key points / key lines (sequence-chained replay: the previous frame is the
// last step's current frame, already on the device)
int launch_expand_compact(gfpl_ctx* ctx, const gfpl_compact_input& c, DevRawFrames raw, DevStereoFrames prev, bool with_prev) {
  if (c.B <= 0) return GFPL_OK;
  int cap = c.cap_pt > c.cap_ln ? c.cap_pt : c.cap_ln;
  if (with_prev) {
    cap = cap > c.prev_cap_pt ? cap : c.prev_cap_pt;
    cap = cap > c.prev_cap_ln ? cap : c.prev_cap_ln;
  }
  dim3 grid((cap + 127) / 128, c.B, with_prev ? 4 : 2);
  expand_compact_kernel<<<grid, 128, 0, ctx->stream>>>(c, raw, prev, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// StereoFrameHandler::updateFrame (stereoFrameHandler.cpp:908-912): the current frame's pose record becomes the previous
// frame's state.  stamps: time stamps of the NEXT current frames, or NULL: the same step again.
__global__ void advance_state_kernel(int B, const gfpl_pose_result* __restrict__ pose, const double* __restrict__ stamps,
                                     gfpl_frame_state* __restrict__ state) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  gfpl_frame_state& s = state[p];
  const gfpl_pose_result& r = pose[p];
  for (int i = 0; i < 16; i++) { s.Tfw[i] = r.Tfw[i]; s.DT[i] = r.DT[i]; }
  for (int i = 0; i < 36; i++) { s.DT_cov[i] = r.DT_cov[i]; s.Tfw_cov[i] = r.Tfw_cov[i]; }
  const double t0 = s.time_stamp, t1 = s.time_stamp_curr;
  s.time_stamp = t1;
  s.time_stamp_curr = stamps ? stamps[p] : t1 + (t1 - t0);
}

int launch_advance_state(gfpl_ctx* ctx, int B, const gfpl_pose_result* pose, const double* stamps, gfpl_frame_state* state) {
  if (B <= 0) return GFPL_OK;
  advance_state_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(B, pose, stamps, state);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

int launch_stereo_uncertainty(gfpl_ctx* ctx, const DevStereoFrames& fr, gfpl_line_cov* out) {
  if (fr.B <= 0 || fr.cap_ln <= 0) return GFPL_OK;
  dim3 grid((fr.cap_ln + 127) / 128, fr.B);
  stereo_uncertainty_kernel<<<grid, 128, 0, ctx->stream>>>(fr, out, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
