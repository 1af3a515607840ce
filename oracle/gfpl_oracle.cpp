// TEST INFRASTRUCTURE — CPU oracle of the GF-PL-SLAM tracking hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
// The product (gf-pl-slam_b200/) never includes, links or calls anything in oracle/.
//
// Restates, function by function and in the reference's operation order, the CPU algorithm of
// /root/reference (SimonsRoad/gf-pl-slam).  Third-party arithmetic that is NOT in the reference tree:
//   * cv::BFMatcher::{knnMatch,radiusMatch} — OpenCV 3.4.1 (reference build_dep.sh:24-28): restated from its
//     published semantics (brute force, ascending distance, ties -> ascending trainIdx) and pinned against
//     cv2.BFMatcher 4.13 by tests/golden (oracle/make_golden.py).
//   * Eigen3 (reference README.md:87 "tested with 3.2.92"): fixed-size products, LDLT, LLT, inverse,
//     SelfAdjointEigenSolver — restated in orc_math.h.
// PARITY: matching is pinned (cv2); the fp64 stages are "parity unpinned" (no reference fixtures exist).
// Build: g++ -O2 -std=c++17 -ffp-contract=off (see oracle/Makefile).
#include "gfpl_oracle.h"
#include "orc_math.h"

#include <climits>
#include <cstdlib>
#include <map>
#include <thread>
#include <vector>

using namespace orc;

// ------------------------------------------------------------------------------------------------
// configuration (reference src/config.cpp:26-154)
// ------------------------------------------------------------------------------------------------
extern "C" void orc_config_default(gfpl_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->has_points = 1; c->has_lines = 1; c->lr_in_parallel = 1; c->best_lr_matches = 1;
  c->use_line_conf_cut = 1; c->max_vol_line_cut = 1;
  c->max_line_match_num = 300; c->max_point_match_num = 500;
  c->min_features = 10; c->max_iters = 5; c->max_iters_ref = 10; c->max_iters_lba = 20;
  c->ratio_disp_std = 0.15; c->ratio_disp_std_hor = 0.9;
  c->max_dist_epip = 2.0; c->min_disp = 1.0; c->max_ratio_12_p = 0.9; c->point_match_radius = 50.0;
  c->stereo_overlap_th = 0.5; c->line_horiz_th = 0.1; c->desc_th_l = 0.1; c->line_cov_th = 10.0;
  c->line_match_radius = 80.0;
  c->homog_th = 0.0000001; c->min_error = 0.0000001; c->min_error_change = 0.0000001;
  c->inlier_k = 2.0; c->motion_step_th = 10; c->orb_scale_factor = 1.2; c->lsd_scale = 1;
  c->lambda_lba_lm = 0.001; c->lambda_lba_k = 10.0;
  c->f2f_point_window = 10.0; c->budget_slack = 1.2; c->cut_step = 0.05; c->cut_rng[0] = 0; c->cut_rng[1] = 1.0;
  c->mad_uninit_value = 0.0;
  c->max_kf_epip_p = 1.0; c->max_kf_epip_l = 1.0;  // config.cpp:42-43
  c->orb_nlevels = 4;                              // :136
}

// ------------------------------------------------------------------------------------------------
// M1-M3: cv::BFMatcher brute force on 32-byte rows
// ------------------------------------------------------------------------------------------------
// NORM_HAMMING: popcount of the XOR (same value as StereoFrame::descriptorDistance, stereoFrame.h:185-201).
static inline int hamming256(const uint8_t* a, const uint8_t* b) {
  uint64_t x[4], y[4];
  std::memcpy(x, a, 32);
  std::memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
         __builtin_popcountll(x[3] ^ y[3]);
}
// NORM_HAMMING2: number of non-zero aligned 2-bit groups of the XOR (OpenCV normHamming cellSize=2).
static inline int hamming2_256(const uint8_t* a, const uint8_t* b) {
  uint64_t x[4], y[4];
  std::memcpy(x, a, 32);
  std::memcpy(y, b, 32);
  int s = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t v = x[i] ^ y[i];
    s += __builtin_popcountll((v | (v >> 1)) & 0x5555555555555555ull);
  }
  return s;
}
static inline int desc_dist(const uint8_t* a, const uint8_t* b, int norm) {
  return norm == GFPL_NORM_HAMMING2 ? hamming2_256(a, b) : hamming256(a, b);
}

struct Knn2 { int i0, i1; float d0, d1; };

// knnMatch(k=2) for one query row: ascending distance, ties -> ascending trainIdx.  nt may be 1 (i1 = -1).
static inline Knn2 knn2_row(const uint8_t* q, const uint8_t* t, int nt, int norm) {
  int b0 = 1 << 30, b1 = 1 << 30, i0 = -1, i1 = -1;
  for (int j = 0; j < nt; j++) {
    int d = desc_dist(q, t + (size_t)j * 32, norm);
    if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
    else if (d < b1) { b1 = d; i1 = j; }
  }
  Knn2 r;
  r.i0 = i0; r.i1 = i1;
  r.d0 = i0 >= 0 ? (float)b0 : INFINITY;
  r.d1 = i1 >= 0 ? (float)b1 : INFINITY;
  return r;
}

static void knn2_set(const uint8_t* q, int nq, const uint8_t* t, int nt, int norm, std::vector<Knn2>& out) {
  out.resize(nq);
  for (int i = 0; i < nq; i++) out[i] = knn2_row(q + (size_t)i * 32, t, nt, norm);
}

extern "C" int orc_knn2_hamming256(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t norm,
                                   int32_t* out_idx, float* out_dist) {
  if (nq < 0 || nt < 2 || (nq > 0 && (!q || !out_idx || !out_dist)) || !t) return GFPL_E_INVALID;
  if (norm != GFPL_NORM_HAMMING && norm != GFPL_NORM_HAMMING2) return GFPL_E_INVALID;
  for (int i = 0; i < nq; i++) {
    Knn2 r = knn2_row(q + (size_t)i * 32, t, nt, norm);
    out_idx[2 * i] = r.i0; out_idx[2 * i + 1] = r.i1;
    out_dist[2 * i] = r.d0; out_dist[2 * i + 1] = r.d1;
  }
  return GFPL_OK;
}

// radiusMatch: every train row with distance <= radius; canonical order (distance, trainIdx).
static void radius_row(const uint8_t* q, const uint8_t* t, int nt, int norm, float radius,
                       std::vector<std::pair<int, int>>& out /* (dist, idx) */) {
  out.clear();
  for (int j = 0; j < nt; j++) {
    int d = desc_dist(q, t + (size_t)j * 32, norm);
    if ((float)d <= radius) out.emplace_back(d, j);
  }
  std::sort(out.begin(), out.end());
}

extern "C" int orc_radius_hamming256(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t norm,
                                     float radius, int32_t* offsets, int32_t* out_idx, float* out_dist, int32_t cap) {
  if (nq < 0 || nt < 0 || !offsets) return GFPL_E_INVALID;
  std::vector<std::pair<int, int>> row;
  int total = 0;
  offsets[0] = 0;
  for (int i = 0; i < nq; i++) {
    radius_row(q + (size_t)i * 32, t, nt, norm, radius, row);
    for (auto& e : row) {
      if (total < cap) { out_idx[total] = e.second; out_dist[total] = (float)e.first; }
      total++;
    }
    offsets[i + 1] = total;
  }
  return total > cap ? GFPL_E_CAPACITY : GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// M6: robust thresholds on knn results
// ------------------------------------------------------------------------------------------------
// StereoFrame::lineDescriptorMAD (stereoFrame.cpp:1287-1313), nn12_mad output only (nn_mad is never
// consumed).  d: [n][2] float distances.  D1: the uninitialised nn_dist_median is `U`.
extern "C" double orc_line_descriptor_mad(const float* d, int32_t n, double U) {
  if (n <= 0) return 0.0;
  // sort by (d1-d0) descending is irrelevant for the result: the values are overwritten before the
  // second sort and only element [n/2] of that second (ascending) sort is read.
  std::vector<float> v(n);
  for (int j = 0; j < n; j++) v[j] = fabsf((float)((double)(d[2 * j + 1] - d[2 * j]) - U));  // :1309 (float-float, -double, fabsf)
  std::sort(v.begin(), v.end());                                                              // :1310
  return 1.4826 * v[n / 2];                                                                   // :1311
}

// pointDescriptorBudgetThres / lineDescriptorBudgetThres (stereoFrame.cpp:1315-1341)
extern "C" double orc_budget_thres(const float* d, int32_t n, int32_t max_match_num) {
  if (n <= 0) return 0.0;
  std::vector<float> v(n);
  for (int j = 0; j < n; j++) v[j] = d[2 * j];
  std::sort(v.begin(), v.end());
  size_t idx = std::min((size_t)max_match_num, (size_t)n) - 1;
  return v[idx];
}

// vector_stdv_mad(vector<double>) (auxiliar.cpp:521-537) — note fabsf: float truncation of |x - median|
extern "C" double orc_vector_stdv_mad(const double* residues, int32_t n) {
  if (n == 0) return 0.0;
  std::vector<double> r(residues, residues + n);
  std::sort(r.begin(), r.end());
  double median = r[n / 2];
  for (int i = 0; i < n; i++) r[i] = fabsf(r[i] - median);
  std::sort(r.begin(), r.end());
  double MAD = r[n / 2];
  return 1.4826 * MAD;
}

// ------------------------------------------------------------------------------------------------
// camera (pinholeStereoCamera.cpp:133-170)
// ------------------------------------------------------------------------------------------------
static inline void back_projection(const gfpl_camera* c, double u, double v, double disp, double* P) {
  double bd = c->b / disp;
  P[0] = bd * (u - c->cx);
  P[1] = bd * (v - c->cy);
  P[2] = bd * c->fx;
}
static inline void projection(const gfpl_camera* c, const double* P, double* uv) {
  uv[0] = c->cx + c->fx * P[0] / P[2];
  uv[1] = c->cy + c->fy * P[1] / P[2];
}
static inline double get_disparity(const gfpl_camera* c, double pZ) { return c->fx * c->b / pZ; }

// PointFeature / LineFeature sigma2 (stereoFeatures.cpp:41-48, 93-101)
static inline double level_sigma2(int level, double scale) {
  double sigma2 = 1.0;
  for (int i = 0; i < level + 1; i++) sigma2 *= scale;
  return 1.f / (sigma2 * sigma2);
}

// ------------------------------------------------------------------------------------------------
// M4: stereo point accept loop (stereoFrame.cpp:206-245 Initial, 829-878 Brute)
// ------------------------------------------------------------------------------------------------
static int stereo_points(const gfpl_camera* cam, const gfpl_config* cfg, int n_l, const gfpl_keypoint* kl,
                         const uint8_t* dl, int n_r, const gfpl_keypoint* kr, const uint8_t* dr,
                         gfpl_point_feature* out, uint8_t* out_desc) {
  if (!cfg->has_points || n_l < 2 || n_r < 2) return 0;  // D3: [i][1] needs two train rows
  std::vector<Knn2> lr, rl;
  knn2_set(dl, n_l, dr, n_r, GFPL_NORM_HAMMING, lr);
  if (cfg->best_lr_matches) knn2_set(dr, n_r, dl, n_l, GFPL_NORM_HAMMING, rl);
  double nn12_dist_th = cfg->max_ratio_12_p;
  int pt_idx = 0;
  for (int i = 0; i < n_l; i++) {
    int lr_qdx = i, lr_tdx = lr[i].i0;
    int rl_tdx = cfg->best_lr_matches ? rl[lr_tdx].i0 : lr_qdx;
    double dist_12 = lr[i].d0 / lr[i].d1;  // float / float -> float, widened
    if (lr_qdx == rl_tdx && dist_12 <= nn12_dist_th) {
      if (fabsf(kl[lr_qdx].y - kr[lr_tdx].y) <= cfg->max_dist_epip) {
        double disp_ = kl[lr_qdx].x - kr[lr_tdx].x;  // float - float, assigned to double
        if (disp_ >= cfg->min_disp) {
          float disp_sub_ = disp_;  // Brute :858 (exact: disp_ is a widened float)
          if (disp_sub_ >= 0) {
            gfpl_point_feature& f = out[pt_idx];
            f.pl[0] = kl[lr_qdx].x; f.pl[1] = kl[lr_qdx].y;
            back_projection(cam, f.pl[0], f.pl[1], disp_sub_, f.P);
            f.disp = disp_sub_;
            f.level = kl[lr_qdx].octave;
            f.sigma2 = level_sigma2(f.level, cfg->orb_scale_factor);
            f.idx = pt_idx; f.src_l = lr_qdx; f.src_r = lr_tdx;
            std::memcpy(out_desc + (size_t)pt_idx * 32, dl + (size_t)lr_qdx * 32, 32);
            pt_idx++;
          }
        }
      }
    }
  }
  return pt_idx;
}

// StereoFrame::lineSegmentOverlapStereo (stereoFrame.cpp:1343-1371)
static double line_segment_overlap_stereo(double spl_obs, double epl_obs, double spl_proj, double epl_proj) {
  double sln = std::min(spl_obs, epl_obs);
  double eln = std::max(spl_obs, epl_obs);
  double spn = std::min(spl_proj, epl_proj);
  double epn = std::max(spl_proj, epl_proj);
  double length = eln - spn;
  double overlap;
  if ((epn < sln) || (spn > eln)) overlap = 0.f;
  else {
    if ((epn > eln) && (spn < sln)) overlap = eln - sln;
    else overlap = std::min(eln, epn) - std::max(sln, spn);
  }
  if (length > 0.01f) overlap = overlap / length;
  else overlap = 0.f;
  return overlap;
}

// analytic endpoint covariance + max eigenvalue (stereoFrame.cpp:956-998)
static double endpoint_cov_max_eig(const gfpl_camera* cam, double px, double py, double disp) {
  double f = cam->fx;
  double px_hat = px - cam->cx, py_hat = py - cam->cy;
  double disp2 = disp * disp;
  double c[9];
  c[0] = disp2 + 2.f * px_hat * px_hat;
  c[1] = 2.f * px_hat * py_hat;
  c[2] = 2.f * f * px_hat;
  c[4] = disp2 + 2.f * py_hat * py_hat;
  c[5] = 2.f * f * py_hat;
  c[8] = 2.f * f * f;
  c[3] = c[1]; c[6] = c[2]; c[7] = c[5];
  double d4 = disp2 * disp2;
  for (int i = 0; i < 9; i++) c[i] = c[i] * cam->b * cam->b / d4;
  double ev[3];
  sym_eigenvalues<3>(c, ev);
  return ev[2];
}

// M5: stereo line accept loop (stereoFrame.cpp:248-334 Initial; 881-1014 Brute == 633-767 ORBSLAM)
static int stereo_lines(const gfpl_camera* cam, const gfpl_config* cfg, int variant, int n_l, const gfpl_keyline* ll,
                        const uint8_t* dl, int n_r, const gfpl_keyline* lr_, const uint8_t* dr,
                        gfpl_line_feature* out, uint8_t* out_desc) {
  if (!cfg->has_lines || n_l < 2 || n_r < 2) return 0;  // D3
  int norm = variant == GFPL_STEREO_INITIAL ? GFPL_NORM_HAMMING : GFPL_NORM_HAMMING2;  // :252 vs :636,885
  std::vector<Knn2> lr, rl;
  knn2_set(dl, n_l, dr, n_r, norm, lr);
  if (cfg->best_lr_matches) knn2_set(dr, n_r, dl, n_l, norm, rl);
  std::vector<float> d(2 * (size_t)n_l);
  for (int i = 0; i < n_l; i++) { d[2 * i] = lr[i].d0; d[2 * i + 1] = lr[i].d1; }
  double nn12_dist_th = orc_line_descriptor_mad(d.data(), n_l, cfg->mad_uninit_value);
  nn12_dist_th = nn12_dist_th * cfg->desc_th_l;
  int n_matches = cfg->best_lr_matches ? std::min(n_l, n_r) : n_l;  // D4
  int ls_idx = 0;
  for (int i = 0; i < n_matches; i++) {
    int lr_qdx = i, lr_tdx = lr[i].i0;
    int rl_tdx = cfg->best_lr_matches ? rl[lr_tdx].i0 : lr_qdx;
    double dist_12 = lr[i].d1 - lr[i].d0;  // float - float
    if (lr_qdx == rl_tdx && dist_12 > nn12_dist_th) {
      const gfpl_keyline& L = ll[lr_qdx];
      const gfpl_keyline& R = lr_[lr_tdx];
      double sp_l[3] = {L.startPointX, L.startPointY, 1.0};
      double ep_l[3] = {L.endPointX, L.endPointY, 1.0};
      double le_l[3] = {sp_l[1] * ep_l[2] - sp_l[2] * ep_l[1], sp_l[2] * ep_l[0] - sp_l[0] * ep_l[2],
                        sp_l[0] * ep_l[1] - sp_l[1] * ep_l[0]};
      double nrm = std::sqrt(le_l[0] * le_l[0] + le_l[1] * le_l[1]);
      le_l[0] = le_l[0] / nrm; le_l[1] = le_l[1] / nrm; le_l[2] = le_l[2] / nrm;
      double sp_r[3] = {R.startPointX, R.startPointY, 1.0};
      double ep_r[3] = {R.endPointX, R.endPointY, 1.0};
      double le_r[3] = {sp_r[1] * ep_r[2] - sp_r[2] * ep_r[1], sp_r[2] * ep_r[0] - sp_r[0] * ep_r[2],
                        sp_r[0] * ep_r[1] - sp_r[1] * ep_r[0]};
      double overlap = line_segment_overlap_stereo(sp_l[1], ep_l[1], sp_r[1], ep_r[1]);
      sp_r[0] = -(le_r[2] + le_r[1] * L.startPointY) / le_r[0];
      ep_r[0] = -(le_r[2] + le_r[1] * L.endPointY) / le_r[0];
      double disp_s = L.startPointX - sp_r[0];
      double disp_e = L.endPointX - ep_r[0];
      double horiz = variant == GFPL_STEREO_INITIAL ? le_r[0] : le_l[0];  // :316 vs :949
      if (disp_s >= cfg->min_disp && disp_e >= cfg->min_disp && fabsf(horiz) > cfg->line_horiz_th &&
          overlap > cfg->stereo_overlap_th) {
        bool keep = true;
        if (variant != GFPL_STEREO_INITIAL) {
          double max_eig = std::max(endpoint_cov_max_eig(cam, sp_l[0], sp_l[1], disp_s),
                                    endpoint_cov_max_eig(cam, ep_l[0], ep_l[1], disp_e));
          keep = max_eig < cfg->line_cov_th;
        }
        if (keep) {
          gfpl_line_feature& f = out[ls_idx];
          f.spl[0] = sp_l[0]; f.spl[1] = sp_l[1]; f.epl[0] = ep_l[0]; f.epl[1] = ep_l[1];
          f.sdisp = disp_s; f.edisp = disp_e;
          back_projection(cam, sp_l[0], sp_l[1], disp_s, f.sP);
          back_projection(cam, ep_l[0], ep_l[1], disp_e, f.eP);
          f.le[0] = le_l[0]; f.le[1] = le_l[1]; f.le[2] = le_l[2];
          f.angle = L.angle;
          f.level = L.octave;
          f.sigma2 = level_sigma2(f.level, cfg->lsd_scale);
          f.idx = variant == GFPL_STEREO_INITIAL ? ls_idx : -1;
          f.src_l = lr_qdx; f.src_r = lr_tdx;
          std::memcpy(out_desc + (size_t)ls_idx * 32, dl + (size_t)lr_qdx * 32, 32);
          ls_idx++;
        }
      }
    }
  }
  return ls_idx;
}

static void stereo_match_one(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_raw_frames* raw, int variant,
                             gfpl_stereo_frames* out, int p) {
  size_t op = (size_t)p * raw->cap_pt, ol = (size_t)p * raw->cap_ln;
  size_t oop = (size_t)p * out->cap_pt, ool = (size_t)p * out->cap_ln;
  out->n_pt[p] = stereo_points(cam, cfg, raw->n_pt_l[p], raw->pt_l + op, raw->pdesc_l + op * 32, raw->n_pt_r[p],
                               raw->pt_r + op, raw->pdesc_r + op * 32, out->pt + oop, out->pdesc + oop * 32);
  out->n_ln[p] = stereo_lines(cam, cfg, variant, raw->n_ln_l[p], raw->ln_l + ol, raw->ldesc_l + ol * 32,
                              raw->n_ln_r[p], raw->ln_r + ol, raw->ldesc_r + ol * 32, out->ln + ool,
                              out->ldesc + ool * 32);
}

extern "C" int orc_stereo_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_raw_frames* raw,
                                int32_t variant, gfpl_stereo_frames* out) {
  if (!cam || !cfg || !raw || !out || out->cap_pt < raw->cap_pt || out->cap_ln < raw->cap_ln) return GFPL_E_INVALID;
  for (int p = 0; p < raw->B; p++) stereo_match_one(cam, cfg, raw, variant, out, p);
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// StereoFrame::extractStereoFeatures_ORBSLAM, point part (stereoFrame.cpp:453-630) + subPixelStereoRefine_ORBSLAM (:340-404)
// ------------------------------------------------------------------------------------------------
// ORBextractor's scale tables (ORBextractor.cc:415-431): float arithmetic, scaleFactor passed as float
static void orb_scale_tables(const gfpl_config* cfg, int n, float* sf, float* isf) {
  const float scaleFactor = (float)cfg->orb_scale_factor;
  sf[0] = 1.0f;
  for (int i = 1; i < n; i++) sf[i] = sf[i - 1] * scaleFactor;
  for (int i = 0; i < n; i++) isf[i] = 1.0f / sf[i];
}

static void subpixel_refine_orbslam(const gfpl_pyramid* pyr, const uint8_t* frame, const float* sf, const float* isf,
                                    const gfpl_keypoint& kpL, const gfpl_keypoint& kpR, float& disparity, float& bestuR) {
  disparity = -1;
  bestuR = kpR.x;
  if (kpL.octave < 0 || kpL.octave >= pyr->n_levels) return;  // precondition of the reference (vector index)
  const float uR0 = kpR.x;
  const float scaleFactor = isf[kpL.octave];
  const float scaleduL = roundf(kpL.x * scaleFactor);
  const float scaledvL = roundf(kpL.y * scaleFactor);
  const float scaleduR0 = roundf(uR0 * scaleFactor);
  const int w = 5, L = 5;
  const int W = pyr->width[kpL.octave], H = pyr->height[kpL.octave], step = pyr->row_stride[kpL.octave];
  const uint8_t* img = frame + pyr->offset[kpL.octave];   // both windows come from the RIGHT pyramid (:358, :376)
  const int r0 = (int)(scaledvL - w), r1 = (int)(scaledvL + w + 1);
  const int cl0 = (int)(scaleduL - w), cl1 = (int)(scaleduL + w + 1);
  if (r0 < 0 || r1 > H || cl0 < 0 || cl1 > W) return;       // defined: cv::Mat::rowRange/colRange would throw
  const float iniu = scaleduR0 + L - w;
  const float endu = scaleduR0 + L + w + 1;
  if (iniu < 0 || endu >= W) return;                          // :370-371
  if ((int)(scaleduR0 - L - w) < 0) return;                   // defined: first window's colRange would throw
  float IL[11][11];
  for (int i = 0; i < 11; i++)
    for (int j = 0; j < 11; j++) IL[i][j] = (float)img[(size_t)(r0 + i) * step + cl0 + j];
  const float cL = IL[w][w];
  for (int i = 0; i < 11; i++)
    for (int j = 0; j < 11; j++) IL[i][j] = IL[i][j] - cL;
  int bestDist = INT_MAX, bestincR = 0;
  float vDists[2 * 5 + 1];
  for (int incR = -L; incR <= +L; incR++) {
    const int c0 = (int)(scaleduR0 + incR - w);
    const float cR = (float)img[(size_t)(r0 + w) * step + c0 + w];
    double acc = 0.0;                                          // cv::norm(NORM_L1) on CV_32F accumulates in double
    for (int i = 0; i < 11; i++)
      for (int j = 0; j < 11; j++) {
        const float ir = (float)img[(size_t)(r0 + i) * step + c0 + j] - cR;
        acc += std::fabs(IL[i][j] - ir);
      }
    const float dist = (float)acc;
    if (dist < bestDist) { bestDist = dist; bestincR = incR; }
    vDists[L + incR] = dist;
  }
  if (bestincR == -L || bestincR == L) return;
  const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
  const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
  if (deltaR < -1 || deltaR > 1) return;
  bestuR = sf[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
  disparity = (kpL.x - bestuR);
}

static int stereo_points_orbslam(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_pyramid* pyr,
                                 const uint8_t* frame, int N, const gfpl_keypoint* kl, const uint8_t* dl, int Nr,
                                 const gfpl_keypoint* kr, const uint8_t* dr, gfpl_point_feature* out, uint8_t* out_desc) {
  if (!cfg->has_points || N == 0 || Nr == 0) return 0;  // :453
  float sf[GFPL_MAX_PYR_LEVELS], isf[GFPL_MAX_PYR_LEVELS];
  orb_scale_tables(cfg, pyr->n_levels, sf, isf);
  const int thOrbDist = 80;
  const int nRows = pyr->height[0];
  std::vector<std::vector<size_t>> vRowIndices(nRows);
  std::vector<float> mvuRight(N, -1.0f), mvDepth(N, -1.0f);
  std::vector<int> src_r(N, -1);
  for (int iR = 0; iR < Nr; iR++) {
    const float kpY = kr[iR].y;
    if (kr[iR].octave < 0 || kr[iR].octave >= pyr->n_levels) continue;   // precondition of the reference (vector index): skipped
    const float r = 2.0f * sf[kr[iR].octave];
    const int maxr = (int)ceilf(kpY + r);
    const int minr = (int)floorf(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);   // rows outside the table are UB in the reference
  }
  const float minD = 0;
  const float maxD = cam->fx;
  const float mbf = cam->fx * cam->b;
  std::vector<std::pair<int, int>> vDistIdx;
  for (int iL = 0; iL < N; iL++) {
    const gfpl_keypoint& kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    if (!(vL >= 0) || (size_t)vL >= (size_t)nRows) continue;      // vRowIndices[vL] out of range: UB in the reference
    const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
    if (vCandidates.empty()) continue;
    const float minU = uL - maxD;
    const float maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = 100;
    size_t bestIdxR = 0;
    for (size_t iC = 0; iC < vCandidates.size(); iC++) {
      const size_t iR = vCandidates[iC];
      const gfpl_keypoint& kpR = kr[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = hamming256(dl + (size_t)iL * 32, dr + iR * 32);   // descriptorDistance (stereoFrame.h:185-201)
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDist < thOrbDist) {
      float disparity, bestuR;
      subpixel_refine_orbslam(pyr, frame, sf, isf, kpL, kr[bestIdxR], disparity, bestuR);
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
        mvDepth[iL] = mbf / disparity;
        mvuRight[iL] = bestuR;
        src_r[iL] = (int)bestIdxR;
        vDistIdx.push_back(std::pair<int, int>(bestDist, iL));
      }
    }
  }
  if (vDistIdx.empty()) return 0;   // defined: the reference indexes vDistIdx[0] of an empty vector (:587)
  std::sort(vDistIdx.begin(), vDistIdx.end());
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int pt_idx = 0;
  for (size_t i = 0; i < vDistIdx.size(); i++) {
    if (vDistIdx[i].first >= thDist) break;
    const int iL = vDistIdx[i].second;
    const float disparity = mbf / mvDepth[iL];
    if (disparity < 0) continue;
    gfpl_point_feature& f = out[pt_idx];
    f.pl[0] = kl[iL].x; f.pl[1] = kl[iL].y;
    back_projection(cam, f.pl[0], f.pl[1], disparity, f.P);
    f.disp = disparity;
    f.level = kl[iL].octave;
    f.sigma2 = level_sigma2(f.level, cfg->orb_scale_factor);
    f.idx = pt_idx; f.src_l = iL; f.src_r = src_r[iL];
    std::memcpy(out_desc + (size_t)pt_idx * 32, dl + (size_t)iL * 32, 32);
    pt_idx++;
  }
  return pt_idx;
}

extern "C" int orc_stereo_match_orbslam(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_raw_frames* raw,
                                        const gfpl_pyramid* pyr, gfpl_stereo_frames* out) {
  if (!cam || !cfg || !raw || !pyr || !out || out->cap_pt < raw->cap_pt || out->cap_ln < raw->cap_ln) return GFPL_E_INVALID;
  if (pyr->n_levels != cfg->orb_nlevels || pyr->n_levels < 1 || pyr->n_levels > GFPL_MAX_PYR_LEVELS) return GFPL_E_INVALID;
  for (int p = 0; p < raw->B; p++) {
    size_t op = (size_t)p * raw->cap_pt, ol = (size_t)p * raw->cap_ln;
    size_t oop = (size_t)p * out->cap_pt, ool = (size_t)p * out->cap_ln;
    out->n_pt[p] = stereo_points_orbslam(cam, cfg, pyr, pyr->data + (size_t)p * pyr->frame_stride, raw->n_pt_l[p],
                                         raw->pt_l + op, raw->pdesc_l + op * 32, raw->n_pt_r[p], raw->pt_r + op,
                                         raw->pdesc_r + op * 32, out->pt + oop, out->pdesc + oop * 32);
    out->n_ln[p] = stereo_lines(cam, cfg, GFPL_STEREO_BRUTE, raw->n_ln_l[p], raw->ln_l + ol, raw->ldesc_l + ol * 32,
                                raw->n_ln_r[p], raw->ln_r + ol, raw->ldesc_r + ol * 32, out->ln + ool,
                                out->ldesc + ool * 32);   // :633-767 == :881-1014
  }
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// C1: StereoFrame::estimateStereoUncertainty (stereoFrame.cpp:1375-1392, 1434-1484)
// ------------------------------------------------------------------------------------------------
static void cov_mat_2d_3d(const gfpl_camera* cam, double u, double v, double u_std, double d, double d_std,
                          double* cov) {
  double cov2D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // D5: off-diagonals defined as 0
  cov2D[0] = u_std * u_std; cov2D[4] = u_std * u_std; cov2D[8] = d_std * d_std;  // pow(x,2)
  double b = cam->b, d_2 = d * d;
  double J[9];
  J[0] = b / d; J[3] = 0.0; J[6] = 0.0;
  J[1] = 0.0; J[4] = b / d; J[7] = 0.0;
  J[2] = -(u - cam->cx) * b / d_2;
  J[5] = -(v - cam->cy) * b / d_2;
  J[8] = -cam->fx * b / d_2;
  double Jt[9], tmp[9];
  transpose<3, 3>(J, Jt);
  matmul<3, 3, 3>(J, cov2D, tmp);
  matmul<3, 3, 3>(tmp, Jt, cov);
}

static void stereo_uncertainty_line(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_line_feature* l,
                                    gfpl_line_cov* c) {
  double spl_std = 1.0, epl_std = 1.0, sdisp_std, edisp_std;
  if (std::fabs(l->le[0]) > 0.15) {
    sdisp_std = cfg->ratio_disp_std * l->sdisp;
    edisp_std = cfg->ratio_disp_std * l->edisp;
  } else {
    sdisp_std = cfg->ratio_disp_std_hor * l->sdisp;
    edisp_std = cfg->ratio_disp_std_hor * l->edisp;
  }
  cov_mat_2d_3d(cam, l->spl[0], l->spl[1], spl_std, l->sdisp, sdisp_std, c->covS);
  cov_mat_2d_3d(cam, l->epl[0], l->epl[1], epl_std, l->edisp, edisp_std, c->covE);
}

extern "C" int orc_stereo_uncertainty(const gfpl_camera* cam, const gfpl_config* cfg,
                                      const gfpl_stereo_frames* fr, gfpl_line_cov* out) {
  if (!cam || !cfg || !fr || !out) return GFPL_E_INVALID;
  for (int p = 0; p < fr->B; p++)
    for (int i = 0; i < fr->n_ln[p]; i++)
      stereo_uncertainty_line(cam, cfg, fr->ln + (size_t)p * fr->cap_ln + i, out + (size_t)p * fr->cap_ln + i);
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// M7/M8: frame-to-frame matching
// ------------------------------------------------------------------------------------------------
static void predicted_pose(const gfpl_frame_state* st, double* Tfw_curr) {
  matmul<4, 4, 4>(st->Tfw, st->DT, Tfw_curr);  // predictFramePose, stereoFrameHandler.cpp:157
}

// StereoFrame::projectPrev3DPoint (stereoFrame.cpp:1550-1570), left projection only
static void project_prev_point(const gfpl_camera* cam, const double* prev_Tfw, const double* curr_Tfw_inv,
                               const double* P, double* uv) {
  double v[4] = {P[0], P[1], P[2], 1.0}, w[4], c[4];
  matmul<4, 4, 1>(prev_Tfw, v, w);
  matmul<4, 4, 1>(curr_Tfw_inv, w, c);
  projection(cam, c, uv);
}

static void f2f_lines(const gfpl_config* cfg, int n_prev, const gfpl_line_feature* lp, const uint8_t* dp, int n_curr,
                      const gfpl_line_feature* lc, const uint8_t* dc, gfpl_ln_obs* out, int cap, int* n_out) {
  *n_out = 0;
  if (!cfg->has_lines || n_prev < 2 || n_curr < 2) return;  // D3
  std::vector<Knn2> m12, m21;
  knn2_set(dp, n_prev, dc, n_curr, GFPL_NORM_HAMMING, m12);  // stereoFrameHandler.cpp:609,777
  if (cfg->best_lr_matches) knn2_set(dc, n_curr, dp, n_prev, GFPL_NORM_HAMMING, m21);
  std::vector<float> d(2 * (size_t)n_prev);
  for (int i = 0; i < n_prev; i++) { d[2 * i] = m12[i].d0; d[2 * i + 1] = m12[i].d1; }
  double nn12_dist_th = orc_line_descriptor_mad(d.data(), n_prev, cfg->mad_uninit_value) * cfg->desc_th_l;
  double budget_dist_th = orc_budget_thres(d.data(), n_prev, cfg->max_line_match_num);
  int n = 0;
  for (int i = 0; i < n_prev; i++) {
    int lr_qdx = i, lr_tdx = m12[i].i0;
    int rl_tdx = cfg->best_lr_matches ? m21[lr_tdx].i0 : lr_qdx;
    if (m12[i].d0 > cfg->budget_slack * budget_dist_th) continue;  // :659
    double dist_12 = m12[i].d1 - m12[i].d0;
    if (lr_qdx == rl_tdx && dist_12 > nn12_dist_th) {
      if (n < cap) {
        gfpl_ln_obs& o = out[n];
        for (int k = 0; k < 3; k++) { o.sP[k] = lp[lr_qdx].sP[k]; o.eP[k] = lp[lr_qdx].eP[k]; o.le_obs[k] = lc[lr_tdx].le[k]; }
        o.sigma2 = lp[lr_qdx].sigma2;
        o.prev_idx = lr_qdx; o.curr_idx = lr_tdx;
      }
      n++;
    }
    if (n >= cfg->max_line_match_num) break;  // :678
  }
  *n_out = std::min(n, cap);
}

static void fill_pt_obs(gfpl_pt_obs& o, const gfpl_point_feature* pp, const gfpl_point_feature* pc, int q, int t) {
  for (int k = 0; k < 3; k++) o.P[k] = pp[q].P[k];
  o.pl_obs[0] = pc[t].pl[0]; o.pl_obs[1] = pc[t].pl[1];
  o.sigma2 = pp[q].sigma2;
  o.prev_idx = q; o.curr_idx = t;
}

// crossFrameMatching_Hybrid, point half (stereoFrameHandler.cpp:454-603)
static void f2f_points_hybrid(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_frame_state* st, int n_prev,
                              const gfpl_point_feature* pp, const uint8_t* dp, int n_curr,
                              const gfpl_point_feature* pc, const uint8_t* dc, gfpl_pt_obs* out, int cap, int* n_out) {
  *n_out = 0;
  if (!cfg->has_points || n_prev == 0 || n_curr == 0) return;
  bool use_radius = cfg->best_lr_matches && cfg->lr_in_parallel;  // :471-483
  if (!use_radius && n_curr < 2) return;                          // D3
  double Tfw_curr[16], Tfw_curr_inv[16];
  predicted_pose(st, Tfw_curr);
  inverse4(Tfw_curr, Tfw_curr_inv);
  std::multimap<int, std::pair<int, float>> pair_frame;
  std::vector<std::pair<int, int>> cand;
  for (int i = 0; i < n_prev; i++) {
    const uint8_t* q = dp + (size_t)i * 32;
    if (use_radius) radius_row(q, dc, n_curr, GFPL_NORM_HAMMING, (float)cfg->point_match_radius, cand);
    else {
      Knn2 r = knn2_row(q, dc, n_curr, GFPL_NORM_HAMMING);
      cand.clear();
      cand.emplace_back((int)r.d0, r.i0);
      cand.emplace_back((int)r.d1, r.i1);
    }
    if (cand.empty()) continue;
    double proj[2];
    project_prev_point(cam, st->Tfw, Tfw_curr_inv, pp[i].P, proj);
    for (auto& c : cand) {
      int t = c.second;
      double dx = proj[0] - pc[t].pl[0], dy = proj[1] - pc[t].pl[1];
      if (std::sqrt(dx * dx + dy * dy) > cfg->f2f_point_window) continue;  // :536
      pair_frame.insert(std::make_pair(t, std::make_pair(i, (float)c.first)));
    }
  }
  // D2: visit the distinct keys in ascending order; per key min distance, ties -> first inserted
  int n = 0;
  auto it = pair_frame.begin();
  while (it != pair_frame.end()) {
    int key = it->first;
    float min_value = 1e+7;
    int matched_previous_pt = 0;
    auto it2 = pair_frame.upper_bound(key);
    for (; it != it2; ++it)
      if (it->second.second < min_value) { min_value = it->second.second; matched_previous_pt = it->second.first; }
    if (n < cap) fill_pt_obs(out[n], pp, pc, matched_previous_pt, key);
    n++;
    if (n >= cfg->max_point_match_num) break;  // :595
  }
  n = std::min(n, cap);
  // D9: a prev point that won several keys is ONE PointFeature object: pl_obs holds the last assignment
  std::map<int, int> last_t;
  for (int k = 0; k < n; k++) last_t[out[k].prev_idx] = out[k].curr_idx;
  for (int k = 0; k < n; k++) {
    int t = last_t[out[k].prev_idx];
    out[k].pl_obs[0] = pc[t].pl[0]; out[k].pl_obs[1] = pc[t].pl[1];
  }
  *n_out = n;
}

// crossFrameMatching_Brute, point half (stereoFrameHandler.cpp:701-771)
static void f2f_points_brute(const gfpl_config* cfg, int n_prev, const gfpl_point_feature* pp, const uint8_t* dp,
                             int n_curr, const gfpl_point_feature* pc, const uint8_t* dc, gfpl_pt_obs* out, int cap,
                             int* n_out) {
  *n_out = 0;
  if (!cfg->has_points || n_prev < 2 || n_curr < 2) return;  // D3
  std::vector<Knn2> m12, m21;
  knn2_set(dp, n_prev, dc, n_curr, GFPL_NORM_HAMMING, m12);
  if (cfg->best_lr_matches) knn2_set(dc, n_curr, dp, n_prev, GFPL_NORM_HAMMING, m21);
  std::vector<float> d(2 * (size_t)n_prev);
  for (int i = 0; i < n_prev; i++) { d[2 * i] = m12[i].d0; d[2 * i + 1] = m12[i].d1; }
  double nn12_dist_th = cfg->max_ratio_12_p;
  double budget_dist_th = orc_budget_thres(d.data(), n_prev, cfg->max_point_match_num);
  int n = 0;
  for (int i = 0; i < n_prev; i++) {
    int lr_qdx = i, lr_tdx = m12[i].i0;
    int rl_tdx = cfg->best_lr_matches ? m21[lr_tdx].i0 : lr_qdx;
    if (m12[i].d0 > cfg->budget_slack * budget_dist_th) continue;  // :748
    double dist_12 = m12[i].d0 / m12[i].d1;
    if (lr_qdx == rl_tdx && dist_12 <= nn12_dist_th) {
      if (n < cap) fill_pt_obs(out[n], pp, pc, lr_qdx, lr_tdx);
      n++;
    }
    if (n >= cfg->max_point_match_num) break;
  }
  *n_out = std::min(n, cap);
}

static void f2f_match_one(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* prev,
                          const gfpl_stereo_frames* curr, const gfpl_frame_state* state, int mode,
                          gfpl_track_obs* out, int p) {
  size_t pp = (size_t)p * prev->cap_pt, pl = (size_t)p * prev->cap_ln;
  size_t cp = (size_t)p * curr->cap_pt, cl = (size_t)p * curr->cap_ln;
  size_t op = (size_t)p * out->cap_pt, ol = (size_t)p * out->cap_ln;
  int npt = 0, nln = 0;
  if (mode == GFPL_F2F_BRUTE)
    f2f_points_brute(cfg, prev->n_pt[p], prev->pt + pp, prev->pdesc + pp * 32, curr->n_pt[p], curr->pt + cp,
                     curr->pdesc + cp * 32, out->pt + op, out->cap_pt, &npt);
  else
    f2f_points_hybrid(cam, cfg, state + p, prev->n_pt[p], prev->pt + pp, prev->pdesc + pp * 32, curr->n_pt[p],
                      curr->pt + cp, curr->pdesc + cp * 32, out->pt + op, out->cap_pt, &npt);
  f2f_lines(cfg, prev->n_ln[p], prev->ln + pl, prev->ldesc + pl * 32, curr->n_ln[p], curr->ln + cl,
            curr->ldesc + cl * 32, out->ln + ol, out->cap_ln, &nln);
  out->n_pt[p] = npt; out->n_ln[p] = nln;
  for (int k = 0; k < npt; k++) out->pt_inlier[op + k] = 1;
  for (int k = 0; k < nln; k++) out->ln_inlier[ol + k] = 1;
}

extern "C" int orc_f2f_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* prev,
                             const gfpl_stereo_frames* curr, const gfpl_frame_state* state, int32_t mode,
                             gfpl_track_obs* out) {
  if (!cam || !cfg || !prev || !curr || !state || !out) return GFPL_E_INVALID;
  for (int p = 0; p < prev->B; p++) f2f_match_one(cam, cfg, prev, curr, state, mode, out, p);
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// KF <-> KF matching of MapHandler::lookForCommonMatches (mapHandler.cpp:211-490), matching + gates only
// ------------------------------------------------------------------------------------------------
static void kf_match_one(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* kf0,
                         const gfpl_stereo_frames* kf1, const double* DT, gfpl_track_obs* out, int p) {
  const size_t p0 = (size_t)p * kf0->cap_pt, l0 = (size_t)p * kf0->cap_ln;
  const size_t p1 = (size_t)p * kf1->cap_pt, l1 = (size_t)p * kf1->cap_ln;
  const size_t op = (size_t)p * out->cap_pt, ol = (size_t)p * out->cap_ln;
  const double chi = std::sqrt(7.815);
  int npt = 0, nln = 0;
  {
    const int n0 = kf0->n_pt[p], n1 = kf1->n_pt[p];
    const gfpl_point_feature* f0 = kf0->pt + p0;
    const gfpl_point_feature* f1 = kf1->pt + p1;
    if (cfg->has_points && n0 >= 2 && n1 >= 2) {  // D3
      std::vector<Knn2> m12, m21;
      knn2_set(kf0->pdesc + p0 * 32, n0, kf1->pdesc + p1 * 32, n1, GFPL_NORM_HAMMING, m12);  // :213
      if (cfg->best_lr_matches) knn2_set(kf1->pdesc + p1 * 32, n1, kf0->pdesc + p0 * 32, n0, GFPL_NORM_HAMMING, m21);
      for (int i = 0; i < n0; i++) {
        const int lr_qdx = i, lr_tdx = m12[i].i0;
        const int rl_tdx = cfg->best_lr_matches ? m21[lr_tdx].i0 : lr_qdx;
        const double dist_12 = m12[i].d0 / m12[i].d1;  // float / float (:257)
        if (lr_qdx == rl_tdx && dist_12 <= cfg->max_ratio_12_p) {
          double P_[3], uv[2];
          se3_apply(DT, f0[lr_qdx].P, P_);
          projection(cam, P_, uv);
          const double dx = uv[0] - f1[lr_tdx].pl[0], dy = uv[1] - f1[lr_tdx].pl[1];
          const double error = std::sqrt(dx * dx + dy * dy) * std::sqrt(f0[lr_qdx].sigma2);  // :264
          if (error < chi) {
            if (npt < out->cap_pt) fill_pt_obs(out->pt[op + npt], f0, f1, lr_qdx, lr_tdx);
            npt++;
          }
        }
      }
    }
  }
  {
    const int n0 = kf0->n_ln[p], n1 = kf1->n_ln[p];
    const gfpl_line_feature* f0 = kf0->ln + l0;
    const gfpl_line_feature* f1 = kf1->ln + l1;
    if (cfg->has_lines && n0 >= 2 && n1 >= 2) {  // D3
      std::vector<Knn2> m12, m21;
      knn2_set(kf0->ldesc + l0 * 32, n0, kf1->ldesc + l1 * 32, n1, GFPL_NORM_HAMMING, m12);  // :345
      if (cfg->best_lr_matches) knn2_set(kf1->ldesc + l1 * 32, n1, kf0->ldesc + l0 * 32, n0, GFPL_NORM_HAMMING, m21);
      std::vector<float> d(2 * (size_t)n0);
      for (int i = 0; i < n0; i++) { d[2 * i] = m12[i].d0; d[2 * i + 1] = m12[i].d1; }
      const double nn12_dist_th = orc_line_descriptor_mad(d.data(), n0, cfg->mad_uninit_value) * cfg->desc_th_l;  // :370-371
      for (int i = 0; i < n0; i++) {
        const int lr_qdx = i, lr_tdx = m12[i].i0;
        const int rl_tdx = cfg->best_lr_matches ? m21[lr_tdx].i0 : lr_qdx;
        const double dist_12 = m12[i].d1 - m12[i].d0;
        if (lr_qdx == rl_tdx && dist_12 > nn12_dist_th) {
          double sP_[3], eP_[3], sp[2], ep[2];
          se3_apply(DT, f0[lr_qdx].sP, sP_);
          projection(cam, sP_, sp);
          se3_apply(DT, f0[lr_qdx].eP, eP_);
          projection(cam, eP_, ep);
          const double* l_obs = f0[lr_qdx].le;  // :398 — the reference measures against kf0's OWN line equation
          const double e0 = l_obs[0] * sp[0] + l_obs[1] * sp[1] + l_obs[2];
          const double e1 = l_obs[0] * ep[0] + l_obs[1] * ep[1] + l_obs[2];
          if (std::sqrt(e0 * e0 + e1 * e1) * std::sqrt(f0[lr_qdx].sigma2) < chi) {  // :403
            if (nln < out->cap_ln) {
              gfpl_ln_obs& o = out->ln[ol + nln];
              for (int k = 0; k < 3; k++) { o.sP[k] = f0[lr_qdx].sP[k]; o.eP[k] = f0[lr_qdx].eP[k]; o.le_obs[k] = f1[lr_tdx].le[k]; }
              o.sigma2 = f0[lr_qdx].sigma2;
              o.prev_idx = lr_qdx; o.curr_idx = lr_tdx;
            }
            nln++;
          }
        }
      }
    }
  }
  out->n_pt[p] = std::min(npt, out->cap_pt);
  out->n_ln[p] = std::min(nln, out->cap_ln);
  for (int k = 0; k < out->n_pt[p]; k++) out->pt_inlier[op + k] = 1;
  for (int k = 0; k < out->n_ln[p]; k++) out->ln_inlier[ol + k] = 1;
}

extern "C" int orc_kf_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* kf0,
                            const gfpl_stereo_frames* kf1, const double* DT, gfpl_track_obs* out) {
  if (!cam || !cfg || !kf0 || !kf1 || !DT || !out) return GFPL_E_INVALID;
  for (int p = 0; p < kf0->B; p++) kf_match_one(cam, cfg, kf0, kf1, DT + (size_t)p * 16, out, p);
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// local map <-> KF matching of MapHandler::lookForCommonMatches (mapHandler.cpp:509-770), matching + gates only
// ------------------------------------------------------------------------------------------------
static inline bool in_image(const gfpl_camera* cam, const double* uv, double z) {  // :526, :649, :651
  return uv[0] > 0 && uv[0] < cam->width && uv[1] > 0 && uv[1] < cam->height && z > 0.0;
}

static void map_match_one(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_local_map* map,
                          const gfpl_stereo_frames* kf1, const double* Twf, gfpl_map_matches* out, int p) {
  int npt = 0, nln = 0;
  {
    const double* X = map->pt3d + (size_t)p * map->cap_pt * 3;
    const uint8_t* D = map->pdesc + (size_t)p * map->cap_pt * 32;
    const gfpl_point_feature* f1 = kf1->pt + (size_t)p * kf1->cap_pt;
    const uint8_t* D1 = kf1->pdesc + (size_t)p * kf1->cap_pt * 32;
    std::vector<int> sel, un;                 // map_local_points (:513-533), unmatched_points (:536-547)
    std::vector<uint8_t> dsel, dun;
    for (int i = 0; i < map->n_pt[p]; i++) {
      double Pf[3], pf[2];
      se3_apply(Twf, X + 3 * i, Pf);
      projection(cam, Pf, pf);
      if (in_image(cam, pf, Pf[2])) { sel.push_back(i); dsel.insert(dsel.end(), D + 32 * i, D + 32 * i + 32); }
    }
    for (int i = 0; i < kf1->n_pt[p]; i++)
      if (f1[i].idx == -1) { un.push_back(i); dun.insert(dun.end(), D1 + 32 * i, D1 + 32 * i + 32); }
    const int n0 = (int)sel.size(), n1 = (int)un.size();
    if (cfg->has_points && n0 >= 2 && n1 >= 2) {  // D3
      std::vector<Knn2> m12, m21;
      knn2_set(dsel.data(), n0, dun.data(), n1, GFPL_NORM_HAMMING, m12);  // :551
      if (cfg->best_lr_matches) knn2_set(dun.data(), n1, dsel.data(), n0, GFPL_NORM_HAMMING, m21);
      for (int i = 0; i < n0; i++) {
        const int lr_qdx = i, lr_tdx = m12[i].i0;
        const int rl_tdx = cfg->best_lr_matches ? m21[lr_tdx].i0 : lr_qdx;
        const double dist_12 = m12[i].d0 / m12[i].d1;  // float / float (:588)
        if (lr_qdx == rl_tdx && dist_12 <= cfg->max_ratio_12_p) {
          double Pf_map[3], pf_map[2];
          se3_apply(Twf, X + 3 * sel[lr_qdx], Pf_map);
          if (Pf_map[2] > 0.0) {
            projection(cam, Pf_map, pf_map);
            const double dx = pf_map[0] - f1[un[lr_tdx]].pl[0], dy = pf_map[1] - f1[un[lr_tdx]].pl[1];
            if (std::sqrt(dx * dx + dy * dy) < cfg->max_kf_epip_p) {  // :606-607
              if (npt < out->cap_pt) out->pt[(size_t)p * out->cap_pt + npt] = gfpl_map_pair{sel[lr_qdx], un[lr_tdx]};
              npt++;
            }
          }
        }
      }
    }
  }
  {
    const double* X = map->ln3d + (size_t)p * map->cap_ln * 6;
    const uint8_t* D = map->ldesc + (size_t)p * map->cap_ln * 32;
    const gfpl_line_feature* f1 = kf1->ln + (size_t)p * kf1->cap_ln;
    const uint8_t* D1 = kf1->ldesc + (size_t)p * kf1->cap_ln * 32;
    std::vector<int> sel, un;                 // map_local_lines (:634-660), unmatched_lines (:663-674)
    std::vector<uint8_t> dsel, dun;
    for (int i = 0; i < map->n_ln[p]; i++) {
      double sPf[3], ePf[3], spf[2], epf[2];
      se3_apply(Twf, X + 6 * i, sPf);
      projection(cam, sPf, spf);
      se3_apply(Twf, X + 6 * i + 3, ePf);
      projection(cam, ePf, epf);
      if (in_image(cam, spf, sPf[2]) && in_image(cam, epf, ePf[2])) {
        sel.push_back(i); dsel.insert(dsel.end(), D + 32 * i, D + 32 * i + 32);
      }
    }
    for (int i = 0; i < kf1->n_ln[p]; i++)
      if (f1[i].idx == -1) { un.push_back(i); dun.insert(dun.end(), D1 + 32 * i, D1 + 32 * i + 32); }
    const int n0 = (int)sel.size(), n1 = (int)un.size();
    if (cfg->has_lines && n0 >= 2 && n1 >= 2) {  // D3
      std::vector<Knn2> m12, m21;
      knn2_set(dsel.data(), n0, dun.data(), n1, GFPL_NORM_HAMMING, m12);  // :679
      if (cfg->best_lr_matches) knn2_set(dun.data(), n1, dsel.data(), n0, GFPL_NORM_HAMMING, m21);
      std::vector<float> d(2 * (size_t)n0);
      for (int i = 0; i < n0; i++) { d[2 * i] = m12[i].d0; d[2 * i + 1] = m12[i].d1; }
      const double nn12_dist_th = orc_line_descriptor_mad(d.data(), n0, cfg->mad_uninit_value) * cfg->desc_th_l;  // :702-703
      for (int i = 0; i < n0; i++) {
        const int lr_qdx = i, lr_tdx = m12[i].i0;
        const int rl_tdx = cfg->best_lr_matches ? m21[lr_tdx].i0 : lr_qdx;
        const double dist_12 = m12[i].d1 - m12[i].d0;
        if (lr_qdx == rl_tdx && dist_12 > nn12_dist_th) {
          double sP_[3], eP_[3], sp[2], ep[2];
          se3_apply(Twf, X + 6 * sel[lr_qdx], sP_);
          projection(cam, sP_, sp);
          se3_apply(Twf, X + 6 * sel[lr_qdx] + 3, eP_);
          projection(cam, eP_, ep);
          const double* l_obs = f1[un[lr_tdx]].le;
          if (sP_[2] > 0.0 && eP_[2] > 0.0) {
            const double e0 = l_obs[0] * sp[0] + l_obs[1] * sp[1] + l_obs[2];  // signed, no abs (:741-743)
            const double e1 = l_obs[0] * ep[0] + l_obs[1] * ep[1] + l_obs[2];
            if (e0 < cfg->max_kf_epip_l && e1 < cfg->max_kf_epip_l) {
              if (nln < out->cap_ln) out->ln[(size_t)p * out->cap_ln + nln] = gfpl_map_pair{sel[lr_qdx], un[lr_tdx]};
              nln++;
            }
          }
        }
      }
    }
  }
  out->n_pt[p] = std::min(npt, out->cap_pt);
  out->n_ln[p] = std::min(nln, out->cap_ln);
}

extern "C" int orc_map_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_local_map* map,
                             const gfpl_stereo_frames* kf1, const double* Twf, gfpl_map_matches* out) {
  if (!cam || !cfg || !map || !kf1 || !Twf || !out) return GFPL_E_INVALID;
  for (int p = 0; p < map->B; p++) map_match_one(cam, cfg, map, kf1, Twf + (size_t)p * 16, out, p);
  return GFPL_OK;
}

// MapPoint / MapLine::updateAverageDescDir, descriptor part (mapFeatures.cpp:51-82, 128-160)
extern "C" int orc_medoid_desc(const uint8_t* desc, const int32_t* offsets, int32_t L, int32_t* out_idx) {
  for (int l = 0; l < L; l++) {
    const uint8_t* D = desc + (size_t)offsets[l] * 32;
    const int n = offsets[l + 1] - offsets[l];
    if (n < 2) { out_idx[l] = 0; continue; }   // only reached through add*Observation in the reference (n >= 2)
    std::vector<int> conf((size_t)n * n, 0);
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) {
        const int d = hamming256(D + 32 * i, D + 32 * j);
        conf[(size_t)i * n + j] = d; conf[(size_t)j * n + i] = d;
      }
    int max_dist = 99999, max_idx = 0;
    for (int i = 0; i < n; i++) {
      std::vector<int> dist_idx(conf.begin() + (size_t)i * n, conf.begin() + (size_t)(i + 1) * n);
      std::sort(dist_idx.begin(), dist_idx.end());
      const int idx_median = dist_idx[int(1 + 0.5 * (n - 1))];
      if (idx_median < max_dist) { max_dist = idx_median; max_idx = i; }
    }
    out_idx[l] = max_idx;
  }
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// C2-C5: good line cutting
// ------------------------------------------------------------------------------------------------
// the 6-vector both optimizeFunctions and getPoseInfoOnLine build (stereoFrameHandler.cpp:1383-1388, 2150-2155)
static inline void jac6(double fgz2, double gx, double gy, double gz, double lx, double ly, double* J) {
  J[0] = +fgz2 * lx * gz;
  J[1] = +fgz2 * ly * gz;
  J[2] = -fgz2 * (gx * lx + gy * ly);
  J[3] = -fgz2 * (gx * gy * lx + gy * gy * ly + gz * gz * ly);
  J[4] = +fgz2 * (gx * gx * lx + gz * gz * lx + gx * gy * ly);
  J[5] = +fgz2 * (gx * gz * ly - gy * gz * lx);
}

// StereoFrame::getJacob3D_2D rows 0-1 (stereoFrame.cpp:1394-1413)
static inline void jacob3d_2d(const gfpl_camera* cam, double px, double py, double pz, double* J23) {
  double f = cam->fx, pz_2 = pz * pz;
  J23[0] = f / pz; J23[1] = 0.0; J23[2] = -f * px / pz_2;
  J23[3] = 0.0; J23[4] = f / pz; J23[5] = -f * py / pz_2;
}

// J_loss^T * (Jp * R * cov * R^T * Jp^T) * J_loss   (stereoFrameHandler.cpp:1359-1362)
static inline double residual_var(const double* Jp23, const double* R, const double* cov, const double* Jl) {
  double A[6], B[6], Rt[9], C[6], Jpt[6], D[4];
  matmul<2, 3, 3>(Jp23, R, A);
  matmul<2, 3, 3>(A, cov, B);
  transpose<3, 3>(R, Rt);
  matmul<2, 3, 3>(B, Rt, C);
  transpose<2, 3>(Jp23, Jpt);
  matmul<2, 3, 2>(C, Jpt, D);
  double r0 = Jl[0] * D[0] + Jl[1] * D[2];
  double r1 = Jl[0] * D[1] + Jl[1] * D[3];
  return r0 * Jl[0] + r1 * Jl[1];
}

// StereoFrameHandler::getPoseInfoOnLine (stereoFrameHandler.cpp:1342-1411)
static void pose_info_on_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, const double* sP,
                              const double* eP, const double* le_obs, const gfpl_line_cov* cov, const double* cut,
                              double* info) {
  double J_loss[2] = {le_obs[0], le_obs[1]};
  double sP_tmp[3], eP_tmp[3], cov_sP[9], cov_eP[9], R[9];
  for (int k = 0; k < 3; k++) {
    sP_tmp[k] = (1 - cut[0]) * sP[k] + cut[0] * eP[k];
    eP_tmp[k] = (1 - cut[1]) * eP[k] + cut[1] * sP[k];
  }
  double a0 = (1 - cut[0]) * (1 - cut[0]), b0 = cut[0] * cut[0];
  double a1 = (1 - cut[1]) * (1 - cut[1]), b1 = cut[1] * cut[1];
  for (int k = 0; k < 9; k++) {
    cov_sP[k] = a0 * cov->covS[k] + b0 * cov->covE[k];
    cov_eP[k] = a1 * cov->covE[k] + b1 * cov->covS[k];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = DT_inv[i * 4 + j];
  double cur_s[3], cur_e[3], Jp[6];
  se3_apply(DT_inv, sP_tmp, cur_s);
  jacob3d_2d(cam, cur_s[0], cur_s[1], cur_s[2], Jp);
  double cov_sR = residual_var(Jp, R, cov_sP, J_loss);
  se3_apply(DT_inv, eP_tmp, cur_e);
  jacob3d_2d(cam, cur_e[0], cur_e[1], cur_e[2], Jp);
  double cov_eR = residual_var(Jp, R, cov_eP, J_loss);
  double Js[6], Je[6];
  double gz2 = cur_s[2] * cur_s[2];
  jac6(cam->fx / std::max(cfg->homog_th, gz2), cur_s[0], cur_s[1], cur_s[2], J_loss[0], J_loss[1], Js);
  gz2 = cur_e[2] * cur_e[2];
  jac6(cam->fx / std::max(cfg->homog_th, gz2), cur_e[0], cur_e[1], cur_e[2], J_loss[0], J_loss[1], Je);
  // cov_r.inverse() for the 2x2 diag(cov_sR, cov_eR): Eigen's 2x2 path = adjugate * (1/det)
  double det = cov_sR * cov_eR - 0.0 * 0.0;
  double invdet = 1.0 / det;
  double i00 = cov_eR * invdet, i11 = cov_sR * invdet;
  double i01 = -0.0 * invdet, i10 = -0.0 * invdet;
  double A[12];  // Jac_r2x (6x2) * inv (2x2)
  for (int i = 0; i < 6; i++) {
    A[i * 2 + 0] = Js[i] * i00 + Je[i] * i10;
    A[i * 2 + 1] = Js[i] * i01 + Je[i] * i11;
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) info[i * 6 + j] = A[i * 2 + 0] * Js[j] + A[i * 2 + 1] * Je[j];
}

// StereoFrameHandler::getPoseInfoPoint (stereoFrameHandler.cpp:1414-1447)
static void pose_info_point(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, const double* P,
                            const double* pl_obs, double* info) {
  double c[3], uv[2], J[6];
  se3_apply(DT_inv, P, c);
  projection(cam, c, uv);
  double dx = uv[0] - pl_obs[0], dy = uv[1] - pl_obs[1];
  double gz2 = c[2] * c[2];
  jac6(cam->fx / std::max(cfg->homog_th, gz2), c[0], c[1], c[2], dx, dy, J);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) info[i * 6 + j] = J[i] * J[j];
}

extern "C" void orc_pose_info_on_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv,
                                      const gfpl_ln_obs* ln, const gfpl_line_cov* cov, const double cut[2],
                                      double* info36) {
  pose_info_on_line(cam, cfg, DT_inv, ln->sP, ln->eP, ln->le_obs, cov, cut, info36);
}
extern "C" void orc_pose_info_point(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv,
                                    const gfpl_pt_obs* pt, double* info36) {
  pose_info_point(cam, cfg, DT_inv, pt->P, pt->pl_obs, info36);
}
extern "C" double orc_logdet6(const double* M) { return logdet6(M); }
// getMinEigenValue (linespec.cpp:53-66): smallest real part of EigenSolver's eigenvalues; the argument is
// symmetric, so the symmetric solver returns the same spectrum.
extern "C" double orc_min_eig6(const double* M) {
  double S[36], ev[6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) { S[i * 6 + j] = M[i * 6 + j]; S[j * 6 + i] = M[i * 6 + j]; }
  sym_eigenvalues<6>(S, ev);
  return ev[0];
}
extern "C" void orc_sym_eig6(const double* M, double* ev) { sym_eigenvalues<6>(M, ev); }
extern "C" void orc_sym_eig3(const double* M, double* ev) { sym_eigenvalues<3>(M, ev); }

static inline double cut_metric(const gfpl_config* cfg, const double* M) {
  return cfg->max_vol_line_cut ? logdet6(M) : orc_min_eig6(M);
}

// StereoFrameHandler::estimateProjUncertainty_submodular (stereoFrameHandler.cpp:1618-1764)
static void linecut_one(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, int npt,
                        const gfpl_pt_obs* pts, int nln, gfpl_ln_obs* lns, const gfpl_line_cov* prev_cov,
                        gfpl_line_cut* cut) {
  if (nln == 0) return;
  const double step = cfg->cut_step;
  const double* rng = cfg->cut_rng;
  const double nb[8][2] = {{step, 0}, {-step, 0}, {0, step}, {0, -step}, {step, step}, {step, -step}, {-step, step}, {-step, -step}};
  double sum[36];
  for (int i = 0; i < 36; i++) sum[i] = 0.0;
  for (int l = 0; l < nln; l++) {
    cut[l].cutRatio[0] = 0; cut[l].cutRatio[1] = 0;
    pose_info_on_line(cam, cfg, DT_inv, lns[l].sP, lns[l].eP, lns[l].le_obs, prev_cov + lns[l].prev_idx,
                      cut[l].cutRatio, cut[l].invCovPose);
    for (int i = 0; i < 36; i++) sum[i] = sum[i] + cut[l].invCovPose[i];
  }
  for (int k = 0; k < npt; k++) {
    double tmp[36];
    pose_info_point(cam, cfg, DT_inv, pts[k].P, pts[k].pl_obs, tmp);
    for (int i = 0; i < 36; i++) sum[i] = sum[i] + tmp[i];
  }
  for (int l = 0; l < nln; l++) {
    gfpl_line_cut& C = cut[l];
    const gfpl_line_cov* cov = prev_cov + lns[l].prev_idx;
    double metric_last = cut_metric(cfg, sum);  // metricArr.back()
    for (int i = 0; i < 36; i++) sum[i] = sum[i] - C.invCovPose[i];
    while (C.cutRatio[0] + C.cutRatio[1] <= 1.0) {
      bool hitLower = false;
      double cand[2] = {0, 0}, info_cand[36];
      double metric_init = metric_last;
      for (int j = 0; j < 8; j++) {
        double ct[2] = {C.cutRatio[0] + nb[j][0], C.cutRatio[1] + nb[j][1]};
        if (ct[0] + ct[1] > 1.0) continue;
        if (ct[0] < rng[0] || ct[0] > rng[1]) continue;
        if (ct[1] < rng[0] || ct[1] > rng[1]) continue;
        double info_tmp[36], M[36];
        pose_info_on_line(cam, cfg, DT_inv, lns[l].sP, lns[l].eP, lns[l].le_obs, cov, ct, info_tmp);
        for (int i = 0; i < 36; i++) M[i] = info_tmp[i] + sum[i];
        double metric_tmp = cut_metric(cfg, M);
        if (metric_tmp > metric_init) {
          metric_init = metric_tmp;
          cand[0] = ct[0]; cand[1] = ct[1];
          std::memcpy(info_cand, info_tmp, sizeof(info_cand));
          hitLower = true;
        }
      }
      if (hitLower) {
        C.cutRatio[0] = cand[0]; C.cutRatio[1] = cand[1];
        std::memcpy(C.invCovPose, info_cand, sizeof(info_cand));
        metric_last = metric_init;
      } else break;
    }
    // updateEndPointByRatio (stereoFrameHandler.cpp:1451-1470)
    C.spl[0] = C.spl[1] = C.epl[0] = C.epl[1] = C.sdisp = C.edisp = 0.0;
    if (!(std::fabs(C.cutRatio[0]) < 0.0001 && std::fabs(C.cutRatio[1]) < 0.0001)) {
      if (std::fabs(C.cutRatio[0]) > 0.0001) {
        double s[3];
        for (int k = 0; k < 3; k++) s[k] = (1 - C.cutRatio[0]) * lns[l].sP[k] + C.cutRatio[0] * lns[l].eP[k];
        for (int k = 0; k < 3; k++) lns[l].sP[k] = s[k];
        projection(cam, s, C.spl);
        C.sdisp = get_disparity(cam, s[2]);
      }
      if (std::fabs(C.cutRatio[1]) > 0.0001) {
        double e[3];
        for (int k = 0; k < 3; k++) e[k] = (1 - C.cutRatio[1]) * lns[l].eP[k] + C.cutRatio[1] * lns[l].sP[k];
        for (int k = 0; k < 3; k++) lns[l].eP[k] = e[k];
        projection(cam, e, C.epl);
        C.edisp = get_disparity(cam, e[2]);
      }
    }
    for (int i = 0; i < 36; i++) sum[i] = sum[i] + C.invCovPose[i];
  }
}

extern "C" int orc_linecut_submodular(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv,
                                      gfpl_track_obs* obs, const gfpl_line_cov* prev_cov, int32_t cov_cap,
                                      gfpl_line_cut* cut) {
  if (!cam || !cfg || !DT_inv || !obs || !prev_cov || !cut) return GFPL_E_INVALID;
  for (int p = 0; p < obs->B; p++)
    linecut_one(cam, cfg, DT_inv + (size_t)p * 16, obs->n_pt[p], obs->pt + (size_t)p * obs->cap_pt, obs->n_ln[p],
                obs->ln + (size_t)p * obs->cap_ln, prev_cov + (size_t)p * cov_cap, cut + (size_t)p * obs->cap_ln);
  return GFPL_OK;
}

// ---- C6: estimateProjUncertainty_descent (stereoFrameHandler.cpp:1476-1610) ---------------------------------------
// StereoFrameHandler::getPoseCovOnLine (stereoFrameHandler.cpp:1249-1338): the pose covariance a line contributes at a
// cut, cov_pose = Jinv^T * diag(cov_sR, cov_eR) * Jinv with Jinv = (J^T J)^-1 J^T, J = [Jac_sr2x Jac_er2x] (6x2).
static void pose_cov_on_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, const double* sP,
                             const double* eP, const double* le_obs, const gfpl_line_cov* cov, const double* cut,
                             double* cov_pose) {
  double J_loss[2] = {le_obs[0], le_obs[1]};
  double sP_tmp[3], eP_tmp[3], cov_sP[9], cov_eP[9], R[9];
  for (int k = 0; k < 3; k++) {
    sP_tmp[k] = (1 - cut[0]) * sP[k] + cut[0] * eP[k];
    eP_tmp[k] = (1 - cut[1]) * eP[k] + cut[1] * sP[k];
  }
  double a0 = (1 - cut[0]) * (1 - cut[0]), b0 = cut[0] * cut[0];
  double a1 = (1 - cut[1]) * (1 - cut[1]), b1 = cut[1] * cut[1];
  for (int k = 0; k < 9; k++) {
    cov_sP[k] = a0 * cov->covS[k] + b0 * cov->covE[k];
    cov_eP[k] = a1 * cov->covE[k] + b1 * cov->covS[k];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = DT_inv[i * 4 + j];
  double cur_s[3], cur_e[3], Jp[6];
  se3_apply(DT_inv, sP_tmp, cur_s);
  jacob3d_2d(cam, cur_s[0], cur_s[1], cur_s[2], Jp);
  double cov_sR = residual_var(Jp, R, cov_sP, J_loss);
  se3_apply(DT_inv, eP_tmp, cur_e);
  jacob3d_2d(cam, cur_e[0], cur_e[1], cur_e[2], Jp);
  double cov_eR = residual_var(Jp, R, cov_eP, J_loss);
  double Js[6], Je[6];
  double gz2 = cur_s[2] * cur_s[2];
  jac6(cam->fx / std::max(cfg->homog_th, gz2), cur_s[0], cur_s[1], cur_s[2], J_loss[0], J_loss[1], Js);
  gz2 = cur_e[2] * cur_e[2];
  jac6(cam->fx / std::max(cfg->homog_th, gz2), cur_e[0], cur_e[1], cur_e[2], J_loss[0], J_loss[1], Je);
  // G = J^T J (2x2, sequential dot products), G^-1 = adjugate * (1/det) (Eigen's 2x2 inverse)
  double g00 = Js[0] * Js[0], g01 = Js[0] * Je[0], g11 = Je[0] * Je[0];
  for (int i = 1; i < 6; i++) {
    g00 = g00 + Js[i] * Js[i];
    g01 = g01 + Js[i] * Je[i];
    g11 = g11 + Je[i] * Je[i];
  }
  double det = g00 * g11 - g01 * g01;
  double invdet = 1.0 / det;
  double i00 = g11 * invdet, i01 = -g01 * invdet, i10 = -g01 * invdet, i11 = g00 * invdet;
  double Jinv[12];  // 2x6: G^-1 * J^T
  for (int j = 0; j < 6; j++) {
    Jinv[j] = i00 * Js[j] + i01 * Je[j];
    Jinv[6 + j] = i10 * Js[j] + i11 * Je[j];
  }
  // (Jinv^T * cov_r) * Jinv with cov_r = [cov_sR 0; 0 cov_eR]
  double A[12];  // 6x2
  for (int i = 0; i < 6; i++) {
    A[i * 2 + 0] = Jinv[i] * cov_sR + Jinv[6 + i] * 0.0;
    A[i * 2 + 1] = Jinv[i] * 0.0 + Jinv[6 + i] * cov_eR;
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) cov_pose[i * 6 + j] = A[i * 2 + 0] * Jinv[j] + A[i * 2 + 1] * Jinv[6 + j];
}

// getEigenValue (linespec.cpp:37-51) sorts the real parts of EigenSolver's eigenvalues in DESCENDING order; the descent
// metric is ev[0] * ev[1].  cov_pose is symmetric (rank 2, positive semi-definite) up to rounding, so the symmetric solver
// applied to its lower triangle returns the same spectrum; rule chosen here: Jacobi eigenvalues of the mirrored lower
// triangle, sorted ascending, product of the two LARGEST (ev[5] * ev[4]).
static double descent_volume(const double* cov_pose) {
  double S[36], ev[6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) { S[i * 6 + j] = cov_pose[i * 6 + j]; S[j * 6 + i] = cov_pose[i * 6 + j]; }
  sym_eigenvalues<6>(S, ev);
  return ev[5] * ev[4];
}
extern "C" void orc_pose_cov_on_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv,
                                     const gfpl_ln_obs* ln, const gfpl_line_cov* cov, const double cut[2], double* cov36) {
  pose_cov_on_line(cam, cfg, DT_inv, ln->sP, ln->eP, ln->le_obs, cov, cut, cov36);
}
extern "C" double orc_descent_volume(const double* cov36) { return descent_volume(cov36); }

// full 3x3 getJacob3D_2D (stereoFrame.cpp:1394-1413), row-major
static inline void jacob3d_2d_full(const gfpl_camera* cam, double px, double py, double pz, double* J) {
  double f = cam->fx, b = cam->b, pz_2 = pz * pz;
  J[0] = f / pz; J[1] = 0.0; J[2] = -f * px / pz_2;
  J[3] = 0.0; J[4] = f / pz; J[5] = -f * py / pz_2;
  J[6] = 0.0; J[7] = 0.0; J[8] = -f * b / pz_2;
}

// one line of estimateProjUncertainty_descent (:1503-1607).  The lines are independent of each other here (no information
// sum): the cut moves as soon as a neighbour is better (inside the j loop, :1545-1550), so later neighbours of the same
// round are taken around the cut that has already moved.
static void linecut_descent_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, gfpl_ln_obs* ln,
                                 const gfpl_line_cov* cov, gfpl_line_cut_descent* C) {
  const double step = cfg->cut_step;
  const double* rng = cfg->cut_rng;
  const double nb[8][2] = {{step, 0}, {-step, 0}, {0, step}, {0, -step}, {step, step}, {step, -step}, {-step, step}, {-step, -step}};
  double cut[2] = {0, 0}, cov_x[36];
  pose_cov_on_line(cam, cfg, DT_inv, ln->sP, ln->eP, ln->le_obs, cov, cut, cov_x);
  double vol_last = descent_volume(cov_x);
  int n_vol = 1;
  while (cut[0] + cut[1] < 1.0) {
    bool hitLower = false;
    for (int j = 0; j < 8; j++) {
      double ct[2] = {cut[0] + nb[j][0], cut[1] + nb[j][1]};
      if (ct[0] + ct[1] >= 1.0) continue;
      if (ct[0] < rng[0] || ct[0] > rng[1]) continue;
      if (ct[1] < rng[0] || ct[1] > rng[1]) continue;
      pose_cov_on_line(cam, cfg, DT_inv, ln->sP, ln->eP, ln->le_obs, cov, ct, cov_x);
      double vol = descent_volume(cov_x);
      if (vol < vol_last) {
        vol_last = vol; n_vol++;
        cut[0] = ct[0]; cut[1] = ct[1];
        hitLower = true;
      }
    }
    if (!hitLower) break;
  }
  C->cutRatio[0] = cut[0]; C->cutRatio[1] = cut[1];
  C->volume = vol_last; C->n_volume = n_vol;
  C->fst_idx = 0; C->lst_idx = 3; C->pad = 0;
  // :1567-1606 — both end points move from the ORIGINAL end points (unlike updateEndPointByRatio)
  double sP_cut[3], eP_cut[3], cov_s[9], cov_e[9], R[9], Rt[9];
  for (int k = 0; k < 3; k++) {
    sP_cut[k] = (1 - cut[0]) * ln->sP[k] + cut[0] * ln->eP[k];
    eP_cut[k] = (1 - cut[1]) * ln->eP[k] + cut[1] * ln->sP[k];
  }
  double a0 = (1 - cut[0]) * (1 - cut[0]), b0 = cut[0] * cut[0];
  double a1 = (1 - cut[1]) * (1 - cut[1]), b1 = cut[1] * cut[1];
  for (int k = 0; k < 9; k++) {
    cov_s[k] = a0 * cov->covS[k] + b0 * cov->covE[k];
    cov_e[k] = a1 * cov->covE[k] + b1 * cov->covS[k];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = DT_inv[i * 4 + j];
  transpose<3, 3>(R, Rt);
  double cur_s[3], cur_e[3], cur_s0[3], cur_e0[3], Jp[9], Jpt[9], T1[9], T2[9], T3[9];
  se3_apply(DT_inv, sP_cut, cur_s);
  se3_apply(DT_inv, eP_cut, cur_e);
  se3_apply(DT_inv, ln->sP, cur_s0);
  se3_apply(DT_inv, ln->eP, cur_e0);
  // covlPt2D[1] / [2] = J_proj * J_dt * cov * J_dt^T * J_proj^T, left to right
  jacob3d_2d_full(cam, cur_s[0], cur_s[1], cur_s[2], Jp);
  transpose<3, 3>(Jp, Jpt);
  matmul<3, 3, 3>(Jp, R, T1); matmul<3, 3, 3>(T1, cov_s, T2); matmul<3, 3, 3>(T2, Rt, T3); matmul<3, 3, 3>(T3, Jpt, C->covlPt2D[0]);
  jacob3d_2d_full(cam, cur_e[0], cur_e[1], cur_e[2], Jp);
  transpose<3, 3>(Jp, Jpt);
  matmul<3, 3, 3>(Jp, R, T1); matmul<3, 3, 3>(T1, cov_e, T2); matmul<3, 3, 3>(T2, Rt, T3); matmul<3, 3, 3>(T3, Jpt, C->covlPt2D[1]);
  projection(cam, cur_s0, C->lPt2D[0]);
  projection(cam, cur_s, C->lPt2D[1]);
  projection(cam, cur_e, C->lPt2D[2]);
  projection(cam, cur_e0, C->lPt2D[3]);
  for (int k = 0; k < 3; k++) { ln->sP[k] = sP_cut[k]; ln->eP[k] = eP_cut[k]; }
  projection(cam, sP_cut, C->spl);
  projection(cam, eP_cut, C->epl);
  C->sdisp = get_disparity(cam, sP_cut[2]);
  C->edisp = get_disparity(cam, eP_cut[2]);
}

extern "C" int orc_linecut_descent(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, gfpl_track_obs* obs,
                                   const gfpl_line_cov* prev_cov, int32_t cov_cap, gfpl_line_cut_descent* cut) {
  if (!cam || !cfg || !DT_inv || !obs || !prev_cov || !cut) return GFPL_E_INVALID;
  for (int p = 0; p < obs->B; p++)
    for (int l = 0; l < obs->n_ln[p]; l++) {
      gfpl_ln_obs* ln = obs->ln + (size_t)p * obs->cap_ln + l;
      linecut_descent_line(cam, cfg, DT_inv + (size_t)p * 16, ln, prev_cov + (size_t)p * cov_cap + ln->prev_idx,
                           cut + (size_t)p * obs->cap_ln + l);
    }
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// P1-P4: pose optimisation
// ------------------------------------------------------------------------------------------------
// StereoFrameHandler::optimizeFunctions (stereoFrameHandler.cpp:2118-2245)
static void optimize_functions(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT, int npt,
                               const gfpl_pt_obs* pts, const uint8_t* pin, int nln, const gfpl_ln_obs* lns,
                               const uint8_t* lin, double* H, double* g, double* e) {
  double H_l[36], H_p[36], g_l[6], g_p[6], e_l = 0.0, e_p = 0.0;
  for (int i = 0; i < 36; i++) H_l[i] = H_p[i] = 0.0;
  for (int i = 0; i < 6; i++) g_l[i] = g_p[i] = 0.0;
  int N_p = 0;
  for (int k = 0; k < npt; k++) {
    if (!pin[k]) continue;
    double P_[3], uv[2], J[6];
    se3_apply(DT, pts[k].P, P_);
    projection(cam, P_, uv);
    double dx = uv[0] - pts[k].pl_obs[0], dy = uv[1] - pts[k].pl_obs[1];
    double err_i_norm = std::sqrt(dx * dx + dy * dy);
    double gz2 = P_[2] * P_[2];
    double fgz2 = cam->fx / std::max(cfg->homog_th, gz2);
    jac6(fgz2, P_[0], P_[1], P_[2], dx, dy, J);
    double dn = std::max(cfg->homog_th, err_i_norm);
    for (int i = 0; i < 6; i++) J[i] = J[i] / dn;
    double s2 = pts[k].sigma2;
    double w = 1.0 / (1.0 + err_i_norm * err_i_norm * s2);
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 6; j++) H_p[i * 6 + j] = H_p[i * 6 + j] + J[i] * J[j] * w;
      g_p[i] = g_p[i] + J[i] * err_i_norm * w;
    }
    e_p = e_p + err_i_norm * err_i_norm * w;
    N_p++;
  }
  int N_l = 0;
  for (int k = 0; k < nln; k++) {
    if (!lin[k]) continue;
    double sP_[3], eP_[3], sp[2], ep[2], Js[6], Je[6], J[6];
    se3_apply(DT, lns[k].sP, sP_);
    projection(cam, sP_, sp);
    se3_apply(DT, lns[k].eP, eP_);
    projection(cam, eP_, ep);
    const double* l = lns[k].le_obs;
    double ds = l[0] * sp[0] + l[1] * sp[1] + l[2];
    double de = l[0] * ep[0] + l[1] * ep[1] + l[2];
    double err_i_norm = std::sqrt(ds * ds + de * de);
    double gz2 = sP_[2] * sP_[2];
    jac6(cam->fx / std::max(cfg->homog_th, gz2), sP_[0], sP_[1], sP_[2], l[0], l[1], Js);
    gz2 = eP_[2] * eP_[2];
    jac6(cam->fx / std::max(cfg->homog_th, gz2), eP_[0], eP_[1], eP_[2], l[0], l[1], Je);
    double dn = std::max(cfg->homog_th, err_i_norm);
    for (int i = 0; i < 6; i++) J[i] = (Js[i] * ds + Je[i] * de) / dn;
    double s2 = lns[k].sigma2;
    double w = 1.0 / (1.0 + err_i_norm * err_i_norm * s2);
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 6; j++) H_l[i * 6 + j] = H_l[i * 6 + j] + J[i] * J[j] * w;
      g_l[i] = g_l[i] + J[i] * err_i_norm * w;
    }
    e_l = e_l + err_i_norm * err_i_norm * w;
    N_l++;
  }
  for (int i = 0; i < 36; i++) H[i] = H_p[i] + H_l[i];
  for (int i = 0; i < 6; i++) g[i] = g_p[i] + g_l[i];
  *e = e_p + e_l;
  *e = *e / (N_l + N_p);
}

extern "C" int orc_pose_hg(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT,
                           const gfpl_track_obs* obs, double* H, double* g, double* e) {
  if (!cam || !cfg || !DT || !obs || !H || !g || !e) return GFPL_E_INVALID;
  for (int p = 0; p < obs->B; p++)
    optimize_functions(cam, cfg, DT + (size_t)p * 16, obs->n_pt[p], obs->pt + (size_t)p * obs->cap_pt,
                       obs->pt_inlier + (size_t)p * obs->cap_pt, obs->n_ln[p], obs->ln + (size_t)p * obs->cap_ln,
                       obs->ln_inlier + (size_t)p * obs->cap_ln, H + (size_t)p * 36, g + (size_t)p * 6, e + p);
  return GFPL_OK;
}

struct ObsView {
  int npt, nln;
  gfpl_pt_obs* pts; uint8_t* pin;
  gfpl_ln_obs* lns; uint8_t* lin;
};

// StereoFrameHandler::gaussNewtonOptimization (stereoFrameHandler.cpp:2032-2056)
static int gauss_newton(const gfpl_camera* cam, const gfpl_config* cfg, const ObsView& o, double* DT, double* DT_cov,
                        double* err_, int max_iters) {
  double H[36], g[6], DT_inc[6];
  double err = 0.0, err_prev = 999999999.9;
  for (int i = 0; i < 36; i++) H[i] = 0.0;  // (uninitialised in the reference when max_iters == 0)
  int iters = 0;
  for (; iters < max_iters; iters++) {
    optimize_functions(cam, cfg, DT, o.npt, o.pts, o.pin, o.nln, o.lns, o.lin, H, g, &err);
    if ((std::fabs(err - err_prev) < cfg->min_error_change) || (err < cfg->min_error)) break;
    ldlt6_solve(H, g, DT_inc);
    double E[16], Ei[16];
    expmap_se3(DT_inc, E);
    inverse_se3(E, Ei);
    matmul<4, 4, 4>(DT, Ei, DT);
    double nrm = 0.0;
    for (int i = 0; i < 6; i++) nrm = nrm + DT_inc[i] * DT_inc[i];
    if (std::sqrt(nrm) < std::numeric_limits<double>::epsilon()) break;
    err_prev = err;
  }
  inverse6(H, DT_cov);
  *err_ = err;
  return iters;
}

// StereoFrameHandler::removeOutliers (stereoFrameHandler.cpp:2058-2116); returns the decrements
static void remove_outliers(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT, ObsView& o,
                            int* n_inliers, int* n_inliers_pt, int* n_inliers_ls) {
  std::vector<double> res_p, res_l;
  for (int k = 0; k < o.npt; k++) {
    double P_[3], uv[2];
    se3_apply(DT, o.pts[k].P, P_);
    projection(cam, P_, uv);
    double dx = uv[0] - o.pts[k].pl_obs[0], dy = uv[1] - o.pts[k].pl_obs[1];
    res_p.push_back(std::sqrt(dx * dx + dy * dy) * std::sqrt(o.pts[k].sigma2));
  }
  for (int k = 0; k < o.nln; k++) {
    double sP_[3], eP_[3], sp[2], ep[2];
    se3_apply(DT, o.lns[k].sP, sP_);
    se3_apply(DT, o.lns[k].eP, eP_);
    projection(cam, sP_, sp);
    projection(cam, eP_, ep);
    const double* l = o.lns[k].le_obs;
    double e0 = l[0] * sp[0] + l[1] * sp[1] + l[2];
    double e1 = l[0] * ep[0] + l[1] * ep[1] + l[2];
    res_l.push_back(std::sqrt(e0 * e0 + e1 * e1) * std::sqrt(o.lns[k].sigma2));
  }
  double inlier_th_p = cfg->inlier_k * orc_vector_stdv_mad(res_p.data(), (int)res_p.size());
  double inlier_th_l = cfg->inlier_k * orc_vector_stdv_mad(res_l.data(), (int)res_l.size());
  for (int k = 0; k < o.npt; k++)
    if (res_p[k] > inlier_th_p) { o.pin[k] = 0; (*n_inliers)--; (*n_inliers_pt)--; }
  for (int k = 0; k < o.nln; k++)
    if (res_l[k] > inlier_th_l) { o.lin[k] = 0; (*n_inliers)--; (*n_inliers_ls)--; }
}

// StereoFrameHandler::optimizePose() (:1864-1937) and optimizePose(Matrix4d) (:1939-2030)
static void optimize_pose(const gfpl_camera* cam, const gfpl_config* cfg, ObsView& o, const gfpl_frame_state* st,
                          const double* DT_ini, int overload, gfpl_pose_result* r) {
  double DT[16], DT_[16], DT_cov[36], err = 0.0;
  std::memcpy(DT, (overload == GFPL_POSE_SEED_GIVEN && DT_ini) ? DT_ini : st->DT, sizeof(DT));
  std::memcpy(DT_cov, st->DT_cov, sizeof(DT_cov));
  int n_inliers_pt = o.npt, n_inliers_ls = o.nln, n_inliers = o.npt + o.nln;  // :688-690
  int status = GFPL_POSE_OK, it1 = 0, it2 = 0;
  bool err_valid = false;
  if (n_inliers > cfg->min_features) {
    std::memcpy(DT_, DT, sizeof(DT));
    it1 = gauss_newton(cam, cfg, o, DT_, DT_cov, &err, cfg->max_iters);
    err_valid = true;
    if (is_finite(DT_, 16)) {
      remove_outliers(cam, cfg, DT_, o, &n_inliers, &n_inliers_pt, &n_inliers_ls);
      if (n_inliers > cfg->min_features)
        it2 = gauss_newton(cam, cfg, o, DT, DT_cov, &err, cfg->max_iters_ref);  // D8: restarts from the seed
      else {
        eye4(DT);
        for (int i = 0; i < 36; i++) DT_cov[i] = 0.0;
        status |= GFPL_POSE_FEW_INLIERS;
      }
    } else {
      eye4(DT);
      for (int i = 0; i < 36; i++) DT_cov[i] = 0.0;
      status |= GFPL_POSE_NONFINITE_1;
    }
  } else {
    eye4(DT);
    for (int i = 0; i < 36; i++) DT_cov[i] = 0.0;
    status |= GFPL_POSE_FEW_FEATURES;
  }
  // `err` is an uninitialised local in the reference when no GN stage ran; defined here as 0.0
  if (!err_valid) err = 0.0;
  bool accept = false;
  if (is_finite(DT, 16)) {
    inverse_se3(DT, r->DT);
    accept = true;
    if (overload == GFPL_POSE_SEED_GIVEN) {
      double tn = std::sqrt(r->DT[3] * r->DT[3] + r->DT[7] * r->DT[7] + r->DT[11] * r->DT[11]);
      if (!(tn < cfg->motion_step_th * (st->time_stamp_curr - st->time_stamp))) {
        accept = false;
        status |= GFPL_POSE_STEP_REJECTED;
      }
    }
  } else
    status |= GFPL_POSE_NONFINITE_2;
  std::memcpy(r->DT_cov, DT_cov, sizeof(DT_cov));
  sym_eigenvalues<6>(DT_cov, r->DT_cov_eig);
  if (accept) {
    matmul<4, 4, 4>(st->Tfw, r->DT, r->Tfw);
    unccomp_se3(st->Tfw, st->Tfw_cov, DT_cov, r->Tfw_cov);
    r->err_norm = err;
  } else {
    eye4(r->DT);
    std::memcpy(r->Tfw, st->Tfw, sizeof(r->Tfw));
    std::memcpy(r->Tfw_cov, st->Tfw_cov, sizeof(r->Tfw_cov));
    r->err_norm = -1.0;
  }
  r->status = status;
  r->n_inliers = n_inliers; r->n_inliers_pt = n_inliers_pt; r->n_inliers_ls = n_inliers_ls;
  r->iters_first = it1; r->iters_ref = it2;
}

static ObsView obs_view(gfpl_track_obs* obs, int p) {
  ObsView o;
  o.npt = obs->n_pt[p]; o.nln = obs->n_ln[p];
  o.pts = obs->pt + (size_t)p * obs->cap_pt; o.pin = obs->pt_inlier + (size_t)p * obs->cap_pt;
  o.lns = obs->ln + (size_t)p * obs->cap_ln; o.lin = obs->ln_inlier + (size_t)p * obs->cap_ln;
  return o;
}

extern "C" int orc_pose_optimize(const gfpl_camera* cam, const gfpl_config* cfg, gfpl_track_obs* obs,
                                 const gfpl_frame_state* state, const double* DT_ini, int32_t overload,
                                 gfpl_pose_result* out) {
  if (!cam || !cfg || !obs || !state || !out) return GFPL_E_INVALID;
  for (int p = 0; p < obs->B; p++) {
    ObsView o = obs_view(obs, p);
    optimize_pose(cam, cfg, o, state + p, DT_ini ? DT_ini + (size_t)p * 16 : nullptr, overload, out + p);
  }
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// whole step: insertStereoPair (brute stereo stage) + optimizePose  (stereoFrameHandler.cpp:83-151)
// ------------------------------------------------------------------------------------------------
static void track_one(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* prev,
                      const gfpl_frame_state* state, const gfpl_raw_frames* raw, int flags, gfpl_stereo_frames* curr,
                      gfpl_track_obs* obs, gfpl_line_cut* cut, gfpl_pose_result* pose, int p) {
  stereo_match_one(cam, cfg, raw, (flags & GFPL_TRACK_STEREO_INITIAL) ? GFPL_STEREO_INITIAL : GFPL_STEREO_BRUTE, curr, p);
  f2f_match_one(cam, cfg, prev, curr, state, (flags & GFPL_TRACK_F2F_BRUTE) ? GFPL_F2F_BRUTE : GFPL_F2F_HYBRID, obs, p);
  if ((flags & GFPL_TRACK_LINECUT) && cfg->use_line_conf_cut && cut) {
    int nl = prev->n_ln[p];
    std::vector<gfpl_line_cov> cov(std::max(nl, 1));
    for (int i = 0; i < nl; i++) stereo_uncertainty_line(cam, cfg, prev->ln + (size_t)p * prev->cap_ln + i, &cov[i]);
    double Tc[16], Tci[16], DT_inv[16];
    predicted_pose(state + p, Tc);
    inverse4(Tc, Tci);
    matmul<4, 4, 4>(Tci, state[p].Tfw, DT_inv);  // stereoFrameHandler.cpp:1635
    linecut_one(cam, cfg, DT_inv, obs->n_pt[p], obs->pt + (size_t)p * obs->cap_pt, obs->n_ln[p],
                obs->ln + (size_t)p * obs->cap_ln, cov.data(), cut + (size_t)p * obs->cap_ln);
  }
  ObsView o = obs_view(obs, p);
  optimize_pose(cam, cfg, o, state + p, state[p].DT,
                (flags & GFPL_TRACK_SEED_GIVEN) ? GFPL_POSE_SEED_GIVEN : GFPL_POSE_SEED_PREV_DT, pose + p);
}

extern "C" int orc_track(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* prev,
                         const gfpl_frame_state* state, const gfpl_raw_frames* raw, int32_t flags,
                         gfpl_stereo_frames* curr, gfpl_track_obs* obs, gfpl_line_cut* cut, gfpl_pose_result* pose,
                         int32_t n_threads) {
  if (!cam || !cfg || !prev || !state || !raw || !curr || !obs || !pose) return GFPL_E_INVALID;
  int B = raw->B;
  if (n_threads <= 1) {
    for (int p = 0; p < B; p++) track_one(cam, cfg, prev, state, raw, flags, curr, obs, cut, pose, p);
    return GFPL_OK;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++)
    th.emplace_back([=]() {
      for (int p = t; p < B; p += n_threads) track_one(cam, cfg, prev, state, raw, flags, curr, obs, cut, pose, p);
    });
  for (auto& x : th) x.join();
  return GFPL_OK;
}

// ------------------------------------------------------------------------------------------------
// L2: one assembly pass of MapHandler::levMarquardtOptimizationLBA (mapHandler.cpp:1243-1426 / 1466-1649)
// ------------------------------------------------------------------------------------------------
extern "C" int orc_lba_assemble(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_lba_problem* pb,
                                gfpl_lba_blocks* out) {
  if (!cam || !cfg || !pb || !out) return GFPL_E_INVALID;
  const int Nkf = pb->Nkf, Npt = pb->Npt, Nls = pb->Nls;
  const int N = 6 * Nkf + 3 * Npt + 6 * Nls;
  for (int i = 0; i < Nkf * 36; i++) out->Hkk[i] = 0.0;
  for (int i = 0; i < Npt * 9; i++) out->Hpp[i] = 0.0;
  for (int i = 0; i < Nls * 36; i++) out->Hll[i] = 0.0;
  for (int i = 0; i < pb->n_pt_obs * 18; i++) out->Hpk[i] = 0.0;
  for (int i = 0; i < pb->n_ls_obs * 36; i++) out->Hlk[i] = 0.0;
  for (int i = 0; i < N; i++) out->g[i] = 0.0;
  double err = 0.0;
  const double fx = cam->fx, fy = cam->fy;
  for (int o = 0; o < pb->n_pt_obs; o++) {
    int lm = pb->pt_obs_lm[o], kf = pb->pt_obs_kf[o];
    const double* Tkw = kf >= 0 ? pb->kf_Tkw + (size_t)kf * 16 : pb->fix_Tkw + (size_t)pb->pt_obs_fix[o] * 16;
    double Tiw[16], Xwi[3], prj[2];
    inverse_se3(Tkw, Tiw);
    se3_apply(Tiw, pb->X_pt + (size_t)lm * 3, Xwi);
    projection(cam, Xwi, prj);
    double dx = pb->pt_obs_uv[2 * o] - prj[0], dy = pb->pt_obs_uv[2 * o + 1] - prj[1];
    double nrm = std::sqrt(dx * dx + dy * dy);
    double gx = Xwi[0], gy = Xwi[1], gz = Xwi[2];
    double gz2 = 1.0 / std::max(cfg->homog_th, gz * gz);
    double fxdx = fx * dx, fydy = fy * dy;
    double JT[6], j3[3], JX[3];
    JT[0] = +gz2 * fxdx * gz;
    JT[1] = +gz2 * fydy * gz;
    JT[2] = -gz2 * (fxdx * gx + fydy * gy);
    JT[3] = -gz2 * (fxdx * gx * gy + fydy * gy * gy + fydy * gz * gz);
    JT[4] = +gz2 * (fxdx * gx * gx + fxdx * gz * gz + fydy * gx * gy);
    JT[5] = +gz2 * (fydy * gx * gz - fxdx * gy * gz);
    double dn = std::max(cfg->homog_th, nrm);
    for (int i = 0; i < 6; i++) JT[i] = JT[i] / dn;
    j3[0] = +gz2 * fxdx * gz; j3[1] = +gz2 * fydy * gz; j3[2] = -gz2 * (fxdx * gx + fydy * gy);
    for (int c = 0; c < 3; c++) {  // Jij_Xwj.transpose() * Tiw.block(0,0,3,3) / max(.)
      double s = j3[0] * Tiw[0 * 4 + c];
      s = s + j3[1] * Tiw[1 * 4 + c];
      s = s + j3[2] * Tiw[2 * 4 + c];
      JX[c] = s / dn;
    }
    double w = 1.0 / (1.0 + nrm * nrm * pb->pt_obs_s2[o]);
    double* gp = out->g + 6 * Nkf + 3 * lm;
    double* Hpp = out->Hpp + (size_t)lm * 9;
    for (int i = 0; i < 3; i++) gp[i] = gp[i] + JX[i] * nrm * w;
    err = err + nrm * nrm * w;
    if (kf >= 0) {
      double* gk = out->g + 6 * kf;
      double* Hkk = out->Hkk + (size_t)kf * 36;
      double* Hpk = out->Hpk + (size_t)o * 18;
      for (int i = 0; i < 6; i++) gk[i] = gk[i] + JT[i] * nrm * w;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 6; j++) Hpk[i * 6 + j] = JX[i] * JT[j] * w;
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) Hkk[i * 6 + j] = Hkk[i * 6 + j] + JT[i] * JT[j] * w;
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Hpp[i * 3 + j] = Hpp[i * 3 + j] + JX[i] * JX[j] * w;
  }
  for (int o = 0; o < pb->n_ls_obs; o++) {
    int lm = pb->ls_obs_lm[o], kf = pb->ls_obs_kf[o];
    const double* Tkw = kf >= 0 ? pb->kf_Tkw + (size_t)kf * 16 : pb->fix_Tkw + (size_t)pb->ls_obs_fix[o] * 16;
    double Tiw[16], Pwi[3], Qwi[3], pp[2], qp[2];
    inverse_se3(Tkw, Tiw);
    se3_apply(Tiw, pb->X_ls + (size_t)lm * 6, Pwi);
    se3_apply(Tiw, pb->X_ls + (size_t)lm * 6 + 3, Qwi);
    projection(cam, Pwi, pp);
    projection(cam, Qwi, qp);
    const double* l = pb->ls_obs_le + (size_t)o * 3;
    double e0 = l[0] * pp[0] + l[1] * pp[1] + l[2];
    double e1 = l[0] * qp[0] + l[1] * qp[1] + l[2];
    double nrm = std::sqrt(e0 * e0 + e1 * e1);
    double dn = std::max(cfg->homog_th, nrm);
    double fxlx = fx * e0, fyly = fy * e1;  // D11: the endpoint errors stand in for (lx, ly)
    double JP[6], JQ[6], jP[3], jQ[3], JT[6], JL[6];
    {
      double gx = Pwi[0], gy = Pwi[1], gz = Pwi[2];
      double gz2 = 1.0 / std::max(cfg->homog_th, gz * gz);
      JP[0] = +gz2 * fxlx * gz;
      JP[1] = +gz2 * fyly * gz;
      JP[2] = -gz2 * (fxlx * gx + fyly * gy);
      JP[3] = -gz2 * (fxlx * gx * gy + fyly * gy * gy + fyly * gz * gz);
      JP[4] = +gz2 * (fxlx * gx * gx + fxlx * gz * gz + fyly * gx * gy);
      JP[5] = +gz2 * (fyly * gx * gz - fxlx * gy * gz);
      jP[0] = JP[0]; jP[1] = JP[1]; jP[2] = JP[2];
    }
    {
      double gx = Qwi[0], gy = Qwi[1], gz = Qwi[2];
      double gz2 = 1.0 / std::max(cfg->homog_th, gz * gz);
      JQ[0] = +gz2 * fxlx * gz;
      JQ[1] = +gz2 * fyly * gz;
      JQ[2] = -gz2 * (fxlx * gx + fyly * gy);
      JQ[3] = -gz2 * (fxlx * gx * gy + fyly * gy * gy + fyly * gz * gz);
      JQ[4] = +gz2 * (fxlx * gx * gx + fxlx * gz * gz + fyly * gx * gy);
      JQ[5] = +gz2 * (fyly * gx * gz - fxlx * gy * gz);
      jQ[0] = JQ[0]; jQ[1] = JQ[1]; jQ[2] = JQ[2];
    }
    for (int c = 0; c < 3; c++) {  // (j^T * R) * l_err(k) / max(.)
      double s = jP[0] * Tiw[0 * 4 + c];
      s = s + jP[1] * Tiw[1 * 4 + c];
      s = s + jP[2] * Tiw[2 * 4 + c];
      JL[c] = s * e0 / dn;
      double u = jQ[0] * Tiw[0 * 4 + c];
      u = u + jQ[1] * Tiw[1 * 4 + c];
      u = u + jQ[2] * Tiw[2 * 4 + c];
      JL[3 + c] = u * e1 / dn;
    }
    for (int i = 0; i < 6; i++) JT[i] = (JP[i] * e0 + JQ[i] * e1) / dn;
    double w = 1.0 / (1.0 + nrm * nrm * pb->ls_obs_s2[o]);
    double* gl = out->g + 6 * Nkf + 3 * Npt + 6 * lm;
    double* Hll = out->Hll + (size_t)lm * 36;
    for (int i = 0; i < 6; i++) gl[i] = gl[i] + JL[i] * nrm * w;
    err = err + nrm * nrm * w;
    if (kf >= 0) {
      double* gk = out->g + 6 * kf;
      double* Hkk = out->Hkk + (size_t)kf * 36;
      double* Hlk = out->Hlk + (size_t)o * 36;
      for (int i = 0; i < 6; i++) gk[i] = gk[i] + JT[i] * nrm * w;
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
          Hlk[i * 6 + j] = JL[i] * JT[j] * w;
          Hkk[i * 6 + j] = Hkk[i * 6 + j] + JT[i] * JT[j] * w;
        }
    }
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) Hll[i * 6 + j] = Hll[i * 6 + j] + JL[i] * JL[j] * w;
  }
  out->err[0] = err;
  return GFPL_OK;
}

extern "C" void orc_expmap_se3(const double* x6, double* T16) { expmap_se3(x6, T16); }
extern "C" void orc_logmap_se3(const double* T16, double* x6) { logmap_se3(T16, x6); }
extern "C" void orc_inverse_se3(const double* T16, double* Tinv16) { inverse_se3(T16, Tinv16); }

// One LM pass of levMarquardtOptimizationLBA: assembly + damped solve on the reference's DENSE system
// (mapHandler.cpp:1428-1441 / 1652-1660).  The reference factorises with Eigen::SimplicialLDLT; this restatement solves the
// same dense system by Gaussian elimination with partial pivoting (a different, independent route to the same DX).
extern "C" int orc_lba_step(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_lba_problem* pb, double lambda,
                            int32_t scale_by_hmax, double* DX, double* err_sum, double* lambda_eff) {
  if (!cam || !cfg || !pb || !DX || !err_sum || !lambda_eff) return GFPL_E_INVALID;
  const int Nkf = pb->Nkf, Npt = pb->Npt, Nls = pb->Nls;
  const int N = 6 * Nkf + 3 * Npt + 6 * Nls;
  std::vector<double> Hkk((size_t)Nkf * 36 + 1), Hpp((size_t)Npt * 9 + 1), Hll((size_t)Nls * 36 + 1),
      Hpk((size_t)pb->n_pt_obs * 18 + 1), Hlk((size_t)pb->n_ls_obs * 36 + 1), g((size_t)N + 1);
  gfpl_lba_blocks b{Hkk.data(), Hpp.data(), Hll.data(), Hpk.data(), Hlk.data(), g.data(), err_sum};
  int rc = orc_lba_assemble(cam, cfg, pb, &b);
  if (rc != GFPL_OK) return rc;
  std::vector<double> H((size_t)N * N, 0.0);
  auto at = [&](int r, int c) -> double& { return H[(size_t)r * N + c]; };
  for (int k = 0; k < Nkf; k++)
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) at(6 * k + i, 6 * k + j) = Hkk[(size_t)k * 36 + i * 6 + j];
  for (int p = 0; p < Npt; p++)
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) at(6 * Nkf + 3 * p + i, 6 * Nkf + 3 * p + j) = Hpp[(size_t)p * 9 + i * 3 + j];
  for (int l = 0; l < Nls; l++)
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++)
      at(6 * Nkf + 3 * Npt + 6 * l + i, 6 * Nkf + 3 * Npt + 6 * l + j) = Hll[(size_t)l * 36 + i * 6 + j];
  for (int o = 0; o < pb->n_pt_obs; o++) {
    const int kf = pb->pt_obs_kf[o], jdx = 6 * Nkf + 3 * pb->pt_obs_lm[o];
    if (kf < 0) continue;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 6; j++) {
      at(jdx + i, 6 * kf + j) += Hpk[(size_t)o * 18 + i * 6 + j];   // H.block(jdx,idx,3,6) += ... (:1308)
      at(6 * kf + j, jdx + i) += Hpk[(size_t)o * 18 + i * 6 + j];   // and its transpose (:1309)
    }
  }
  for (int o = 0; o < pb->n_ls_obs; o++) {
    const int kf = pb->ls_obs_kf[o], jdx = 6 * Nkf + 3 * Npt + 6 * pb->ls_obs_lm[o];
    if (kf < 0) continue;
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
      at(jdx + i, 6 * kf + j) += Hlk[(size_t)o * 36 + i * 6 + j];
      at(6 * kf + j, jdx + i) += Hlk[(size_t)o * 36 + i * 6 + j];
    }
  }
  double lam = lambda;
  if (scale_by_hmax) {
    int Hmax = 0;  // `int Hmax` (:1429-1435)
    for (int i = 0; i < N; i++)
      if (at(i, i) > Hmax || at(i, i) < -Hmax) Hmax = (int)std::fabs(at(i, i));
    lam = lam * Hmax;
  }
  *lambda_eff = lam;
  for (int i = 0; i < N; i++) at(i, i) += lam * at(i, i);   // :1436-1437
  std::vector<double> x(g.begin(), g.begin() + N);
  for (int k = 0; k < N; k++) {
    int piv = k;
    for (int i = k + 1; i < N; i++) if (std::fabs(at(i, k)) > std::fabs(at(piv, k))) piv = i;
    if (piv != k) {
      for (int c = k; c < N; c++) std::swap(at(k, c), at(piv, c));
      std::swap(x[k], x[piv]);
    }
    const double d = at(k, k);
    for (int i = k + 1; i < N; i++) {
      const double f = at(i, k) / d;
      if (f == 0.0) continue;
      for (int c = k; c < N; c++) at(i, c) -= f * at(k, c);
      x[i] -= f * x[k];
    }
  }
  for (int k = N - 1; k >= 0; k--) {
    double s = x[k];
    for (int c = k + 1; c < N; c++) s -= at(k, c) * x[c];
    x[k] = s / at(k, k);
  }
  for (int i = 0; i < N; i++) DX[i] = x[i];
  return GFPL_OK;
}
