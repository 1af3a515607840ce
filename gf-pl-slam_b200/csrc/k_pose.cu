// k_pose.cu — K4: the Gauss-Newton pose optimiser of StVO::StereoFrameHandler for sm_100a.
//
// Replaces optimizeFunctions (reference stereoFrameHandler.cpp:2118-2245), gaussNewtonOptimization
// (:2032-2056), removeOutliers (:2058-2116) + vector_stdv_mad (auxiliar.cpp:521-537) and both
// optimizePose overloads (:1864-2030).
//
// One warp owns one frame pair for the whole optimisation (both GN stages and the outlier pass), so the
// kernel is persistent across iterations: the observation lists are read from L1/L2, the pose and the
// 28 normal-equation partials (21 lower-triangle H + 6 g + 1 e) live in registers, and each iteration
// ends in a fixed xor-butterfly of fp64 warp shuffles — every lane ends with bit-identical H, g, e and
// the result is run-to-run reproducible.  All lanes then run the 6x6 pivoted LDLT solve and the SE(3)
// update redundantly (no broadcast, no divergence).  FP64-ALU / latency bound:
// flops per iteration ~ 114 per point + 212 per line.
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

constexpr int POSE_WARPS = 4;
#ifndef POSE_MIN_BLOCKS
#define POSE_MIN_BLOCKS 3  // <= 168 registers (measured: 1.70 ms vs 1.98 ms at 4 blocks = 128 registers with spills in the GN loop)
#endif

// hand-off record between the GN kernel and the finalize kernel
struct PoseRaw {
  double DT[16];
  double H[36];
  double err;
  int status, cov_from_H, n_inliers, n_inliers_pt, n_inliers_ls, iters_first, iters_ref, pad;
};

struct Acc {
  double v[28];  // [0,21) lower triangle of H, [21,27) g, [27] e
};

__device__ __forceinline__ void acc_add(Acc& a, const double* J, double r, double w) {
  const double rw = r * w;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double Jw = J[i] * w;
#pragma unroll
    for (int j = 0; j <= i; j++) a.v[tri(i, j)] = fma(Jw, J[j], a.v[tri(i, j)]);
    a.v[21 + i] = fma(J[i], rw, a.v[21 + i]);
  }
  a.v[27] = fma(r, rw, a.v[27]);
}

// This translation unit is tolerance-compared (1e-9) and compiled with FMA contraction: divisions that the
// reference performs element-wise (J/max(eps,r), fx*X/Z, ...) are done as one reciprocal and multiplies.

// point term (stereoFrameHandler.cpp:2131-2167); returns the residual norm
__device__ __forceinline__ double point_residual(const double* DT, const gfpl_pt_obs& o, const gfpl_camera& cam,
                                                 double* P_, double* dxy) {
  se3_apply(DT, o.P, P_);
  const double iz = 1.0 / P_[2];
  dxy[0] = (cam.cx + cam.fx * P_[0] * iz) - o.pl_obs[0];
  dxy[1] = (cam.cy + cam.fy * P_[1] * iz) - o.pl_obs[1];
  return sqrt(dxy[0] * dxy[0] + dxy[1] * dxy[1]);
}

// line term (:2171-2235)
__device__ __forceinline__ double line_residual(const double* DT, const gfpl_ln_obs& o, const gfpl_camera& cam,
                                                double* sP_, double* eP_, double* dse) {
  se3_apply(DT, o.sP, sP_);
  se3_apply(DT, o.eP, eP_);
  const double izs = 1.0 / sP_[2], ize = 1.0 / eP_[2];
  const double su = cam.cx + cam.fx * sP_[0] * izs, sv = cam.cy + cam.fy * sP_[1] * izs;
  const double eu = cam.cx + cam.fx * eP_[0] * ize, ev = cam.cy + cam.fy * eP_[1] * ize;
  dse[0] = o.le_obs[0] * su + o.le_obs[1] * sv + o.le_obs[2];
  dse[1] = o.le_obs[0] * eu + o.le_obs[1] * ev + o.le_obs[2];
  return sqrt(dse[0] * dse[0] + dse[1] * dse[1]);
}


// ---- LDLT<Matrix6d>(H).solve(g) in registers ------------------------------------------------------------
// Same algorithm as Eigen's ldlt_inplace<Lower>::unblocked (largest |diagonal| pivot, left-looking update,
// D^-1 with a zero-pivot guard) but fully unrolled: the runtime pivot index is applied with predicated
// swaps, so nothing is indexed dynamically and nothing spills to local memory.
__device__ __forceinline__ void swap_rc(double* A, int* perm, int k, int p, bool doit) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double a = A[k * 6 + j], b = A[p * 6 + j];
    A[k * 6 + j] = doit ? b : a;
    A[p * 6 + j] = doit ? a : b;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double a = A[i * 6 + k], b = A[i * 6 + p];
    A[i * 6 + k] = doit ? b : a;
    A[i * 6 + p] = doit ? a : b;
  }
  const int pa = perm[k], pb = perm[p];
  perm[k] = doit ? pb : pa;
  perm[p] = doit ? pa : pb;
}

__device__ __forceinline__ void ldlt6_solve_reg(const double* Hin, const double* g, double* x) {
  double A[36];
  int perm[6];
#pragma unroll
  for (int i = 0; i < 36; i++) A[i] = Hin[i];
#pragma unroll
  for (int i = 0; i < 6; i++) perm[i] = i;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    int p = k;
    double best = fabs(A[k * 6 + k]);
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      const double v = fabs(A[i * 6 + i]);
      if (v > best) { best = v; p = i; }
    }
#pragma unroll
    for (int pp = k + 1; pp < 6; pp++) swap_rc(A, perm, k, pp, p == pp);
    double temp[6];
#pragma unroll
    for (int j = 0; j < k; j++) temp[j] = A[j * 6 + j] * A[k * 6 + j];
    if (k > 0) {
      double dot = A[k * 6 + 0] * temp[0];
#pragma unroll
      for (int j = 1; j < k; j++) dot = dot + A[k * 6 + j] * temp[j];
      A[k * 6 + k] = A[k * 6 + k] - dot;
    }
    const double d = A[k * 6 + k];
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      double sv = A[i * 6 + k];
      if (k > 0) {
        double dot = A[i * 6 + 0] * temp[0];
#pragma unroll
        for (int j = 1; j < k; j++) dot = dot + A[i * 6 + j] * temp[j];
        sv = sv - dot;
      }
      A[i * 6 + k] = (fabs(d) > 0.0) ? sv / d : sv;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < 6; j++) v = (perm[i] == j) ? g[j] : v;
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < i; j++) y[i] = y[i] - A[i * 6 + j] * y[j];
  const double tol = 1.0 / 1.7976931348623157e308;
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] = (fabs(A[i * 6 + i]) > tol) ? y[i] / A[i * 6 + i] : 0.0;
#pragma unroll
  for (int i = 5; i >= 0; i--)
#pragma unroll
    for (int j = i + 1; j < 6; j++) y[i] = y[i] - A[j * 6 + i] * y[j];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) v = (perm[i] == j) ? y[i] : v;
    x[j] = v;
  }
}

// H is symmetric positive definite whenever the pose is observable, and then the pivoted and the unpivoted LDL^T are the
// same solve up to rounding (both backward stable; this translation unit is compared at 1e-9).  The unpivoted
// factorisation needs no row / column exchanges — the predicated swaps above are ~4/5 of the instructions of the
// pivoted routine, executed by every lane in every GN iteration.  It reports failure when a pivot is not safely positive
// (relative to the largest diagonal entry); the caller then runs the pivoted routine, which is what Eigen would do.
__device__ __forceinline__ bool ldlt6_solve_nopivot(const double* H, const double* g, double* x) {
  double L[36], d[6], id[6];
  double dmax = H[0];
#pragma unroll
  for (int i = 1; i < 6; i++) dmax = fmax(dmax, H[i * 6 + i]);
  const double floor_ = 1e-11 * dmax;
  bool ok = dmax > 0.0 && dmax < 1e300;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double t[6];
#pragma unroll
    for (int j = 0; j < k; j++) t[j] = L[k * 6 + j] * d[j];
    double dk = H[k * 6 + k];
#pragma unroll
    for (int j = 0; j < k; j++) dk = fma(-L[k * 6 + j], t[j], dk);
    d[k] = dk;
    ok = ok && (dk > floor_);
    id[k] = 1.0 / dk;
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      double v = H[i * 6 + k];
#pragma unroll
      for (int j = 0; j < k; j++) v = fma(-L[i * 6 + j], t[j], v);
      L[i * 6 + k] = v * id[k];
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double v = g[i];
#pragma unroll
    for (int j = 0; j < i; j++) v = fma(-L[i * 6 + j], y[j], v);
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] = y[i] * id[i];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double v = y[i];
#pragma unroll
    for (int j = i + 1; j < 6; j++) v = fma(-L[j * 6 + i], x[j], v);
    x[i] = v;
  }
  return ok;
}

// pivoted fallback, out of line (gfpl_math.cuh ldlt6_solve, local-memory arrays): only for H that is not safely SPD
__device__ __noinline__ void ldlt6_solve_pivoted_slow(const double* H, const double* g, double* x) { ldlt6_solve(H, g, x); }

struct PairObs {
  int npt, nln;
  const gfpl_pt_obs* pts;
  const gfpl_ln_obs* lns;
  uint8_t* pin;
  uint8_t* lin;
};

// optimizeFunctions for one pair by one warp.  Returns H (full 6x6), g, e in every lane.
__device__ __forceinline__ void warp_optimize_functions(const double* DT, const PairObs& o, const gfpl_camera& cam,
                                        const gfpl_config& cfg, double* H, double* g, double* e) {
  const int lane = threadIdx.x & 31;
  Acc a;
#pragma unroll
  for (int i = 0; i < 28; i++) a.v[i] = 0.0;
  int N = 0;
  // software-pipelined: the observation of the next round is loaded before this round's arithmetic
  gfpl_pt_obs nxt;
  uint8_t nxt_in = 0;
  if (lane < o.npt) { nxt = o.pts[lane]; nxt_in = o.pin[lane]; }
  for (int k = lane; k < o.npt; k += 32) {
    const gfpl_pt_obs ob = nxt;
    const uint8_t in = nxt_in;
    if (k + 32 < o.npt) { nxt = o.pts[k + 32]; nxt_in = o.pin[k + 32]; }
    if (!in) continue;
    double P_[3], d[2], J[6];
    const double r = point_residual(DT, ob, cam, P_, d);
    const double fgz2 = cam.fx / fmax(cfg.homog_th, P_[2] * P_[2]);
    jac6(fgz2, P_[0], P_[1], P_[2], d[0], d[1], J);
    const double idn = 1.0 / fmax(cfg.homog_th, r);
#pragma unroll
    for (int i = 0; i < 6; i++) J[i] = J[i] * idn;
    const double w = 1.0 / (1.0 + r * r * ob.sigma2);
    acc_add(a, J, r, w);
    N++;
  }
  for (int k = lane; k < o.nln; k += 32) {
    if (!o.lin[k]) continue;
    double sP_[3], eP_[3], d[2], Js[6], Je[6], J[6];
    const gfpl_ln_obs ob = o.lns[k];
    const double r = line_residual(DT, ob, cam, sP_, eP_, d);
    jac6(cam.fx / fmax(cfg.homog_th, sP_[2] * sP_[2]), sP_[0], sP_[1], sP_[2], ob.le_obs[0], ob.le_obs[1], Js);
    jac6(cam.fx / fmax(cfg.homog_th, eP_[2] * eP_[2]), eP_[0], eP_[1], eP_[2], ob.le_obs[0], ob.le_obs[1], Je);
    const double idn = 1.0 / fmax(cfg.homog_th, r);
#pragma unroll
    for (int i = 0; i < 6; i++) J[i] = (Js[i] * d[0] + Je[i] * d[1]) * idn;
    const double w = 1.0 / (1.0 + r * r * ob.sigma2);
    acc_add(a, J, r, w);
    N++;
  }
  // fixed xor-butterfly: identical association in every lane, deterministic run to run
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < 28; i++) a.v[i] = a.v[i] + shfl_xor_f64(0xffffffffu, a.v[i], s);
    N += __shfl_xor_sync(0xffffffffu, N, s);
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) { H[i * 6 + j] = a.v[tri(i, j)]; H[j * 6 + i] = a.v[tri(i, j)]; }
    g[i] = a.v[21 + i];
  }
  *e = a.v[27] / (double)N;  // e /= (N_l + N_p), 0/0 -> NaN like the reference
}

// gaussNewtonOptimization (:2032-2056); H of the last evaluation is returned for DT_cov = H^-1
__device__ __forceinline__ int warp_gauss_newton(double* DT, const PairObs& o, const gfpl_camera& cam, const gfpl_config& cfg,
                                 int max_iters, double* H, double* err_) {
  double g[6], inc[6];
  double err = 0.0, err_prev = 999999999.9;
  for (int i = 0; i < 36; i++) H[i] = 0.0;
  int iters = 0;
  for (; iters < max_iters; iters++) {
    warp_optimize_functions(DT, o, cam, cfg, H, g, &err);
    if ((fabs(err - err_prev) < cfg.min_error_change) || (err < cfg.min_error)) break;
    if (!ldlt6_solve_nopivot(H, g, inc)) {
      double Hm[36], gm[6], xm[6];
      for (int i = 0; i < 36; i++) Hm[i] = H[i];
      for (int i = 0; i < 6; i++) gm[i] = g[i];
      ldlt6_solve_pivoted_slow(Hm, gm, xm);
      for (int i = 0; i < 6; i++) inc[i] = xm[i];
    }
    double E[16], Ei[16];
    expmap_se3(inc, E);
    inverse_se3(E, Ei);
    matmul<4, 4, 4>(DT, Ei, DT);
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) nrm = nrm + inc[i] * inc[i];
    if (sqrt(nrm) < 2.220446049250313e-16) break;
    err_prev = err;
  }
  *err_ = err;
  return iters;
}

// element [rank] of the ascending order of n unsigned keys, without sorting: MSB-first binary search on the key bits,
// one warp-wide count (REDUX.SUM) per probe.  key k lives at keys[k * STRIDE]; with RESTRICT only the elements whose
// companion word keys[k * STRIDE + 1] equals `hi` take part (ties of the high word of a double, see below).
template <int STRIDE, bool RESTRICT, int TOP_BIT>
__device__ __forceinline__ uint32_t warp_select_u32(const uint32_t* keys, int n, int rank, uint32_t hi) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0u;
  for (int bit = TOP_BIT; bit >= 0; bit--) {
    const uint32_t trial = prefix | (1u << bit);
    int cnt = 0;
    for (int k = lane; k < n; k += 32) {
      const bool in = !RESTRICT || keys[k * STRIDE + 1] == hi;
      cnt += (in && keys[k * STRIDE] < trial) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt <= rank) prefix = trial;
  }
  return prefix;
}

// 1.4826 * MAD of res[0..n) exactly as vector_stdv_mad (auxiliar.cpp:521-537) computes it:
// median = sorted[n/2]; a = (double)fabsf(res - median) (float truncation); MAD = sorted(a)[n/2].
// Order statistics of non-negative IEEE values are order statistics of their bit patterns, so both selections run on
// 32-bit integer keys: the median on the high words of the doubles (low words only among the elements that tie on the
// high word), the MAD on the float patterns of `a`, staged once in `scratch` (n words).
__device__ __forceinline__ double warp_stdv_mad(const double* res, int n, uint32_t* scratch) {
  if (n == 0) return 0.0;
  const int lane = threadIdx.x & 31;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(res);  // little endian: w[2k] low word, w[2k+1] high word
  const int rank = n / 2;
  const uint32_t hi = warp_select_u32<2, false, 30>(w + 1, n, rank, 0u);
  int below = 0, same = 0;
  for (int k = lane; k < n; k += 32) {
    below += w[2 * k + 1] < hi ? 1 : 0;
    same += w[2 * k + 1] == hi ? 1 : 0;
  }
  below = __reduce_add_sync(0xffffffffu, below);
  same = __reduce_add_sync(0xffffffffu, same);
  uint32_t lo;
  if (same == 1) {  // the usual case: the element with that high word is the median
    uint32_t mine = 0u;
    for (int k = lane; k < n; k += 32) mine = w[2 * k + 1] == hi ? w[2 * k] : mine;
    lo = __reduce_max_sync(0xffffffffu, mine);
  } else {
    lo = warp_select_u32<2, true, 31>(w, n, rank - below, hi);
  }
  const double median = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  for (int k = lane; k < n; k += 32) scratch[k] = __float_as_uint(fabsf((float)(res[k] - median)));
  __syncwarp();
  const uint32_t mad = warp_select_u32<1, false, 30>(scratch, n, rank, 0u);
  __syncwarp();
  return 1.4826 * (double)__uint_as_float(mad);
}

// removeOutliers (:2058-2116): returns the number of points / lines flagged
__device__ __forceinline__ void warp_remove_outliers(const double* DT, const PairObs& o, const gfpl_camera& cam,
                                     const gfpl_config& cfg, double* s_res, int* drop_pt, int* drop_ln) {
  const int lane = threadIdx.x & 31;
  double* res_p = s_res;
  double* res_l = s_res + o.npt;
  for (int k = lane; k < o.npt; k += 32) {
    double P_[3], d[2];
    const gfpl_pt_obs ob = o.pts[k];
    const double r = point_residual(DT, ob, cam, P_, d);
    res_p[k] = r * sqrt(ob.sigma2);
  }
  for (int k = lane; k < o.nln; k += 32) {
    double sP_[3], eP_[3], d[2];
    const gfpl_ln_obs ob = o.lns[k];
    const double r = line_residual(DT, ob, cam, sP_, eP_, d);
    res_l[k] = r * sqrt(ob.sigma2);
  }
  __syncwarp();
  uint32_t* scratch = reinterpret_cast<uint32_t*>(s_res + o.npt + o.nln);  // [max(npt, nln)] words behind the residuals
  const double th_p = cfg.inlier_k * warp_stdv_mad(res_p, o.npt, scratch);
  const double th_l = cfg.inlier_k * warp_stdv_mad(res_l, o.nln, scratch);
  int dp = 0, dl = 0;
  for (int k = lane; k < o.npt; k += 32)
    if (res_p[k] > th_p) { o.pin[k] = 0; dp++; }
  for (int k = lane; k < o.nln; k += 32)
    if (res_l[k] > th_l) { o.lin[k] = 0; dl++; }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    dp += __shfl_xor_sync(0xffffffffu, dp, s);
    dl += __shfl_xor_sync(0xffffffffu, dl, s);
  }
  *drop_pt = dp;
  *drop_ln = dl;
  __syncwarp();
}

// lane-0 epilogue helpers (local-memory loops; executed once per pair)
__device__ __noinline__ void unccomp_se3(const double* T1, const double* covT1, const double* covTinc, double* covT2) {
  double Ad[36], tmp[36];
  double sk[9], skR[9];
  const double t[3] = {T1[3], T1[7], T1[11]};
  skew(t, sk);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = sk[i * 3 + 0] * T1[0 * 4 + j];
      s = s + sk[i * 3 + 1] * T1[1 * 4 + j];
      s = s + sk[i * 3 + 2] * T1[2 * 4 + j];
      skR[i * 3 + j] = s;
    }
  for (int i = 0; i < 36; i++) Ad[i] = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Ad[i * 6 + j] = T1[i * 4 + j];
      Ad[i * 6 + 3 + j] = skR[i * 3 + j];
      Ad[(i + 3) * 6 + 3 + j] = T1[i * 4 + j];
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = Ad[i * 6 + 0] * covTinc[0 * 6 + j];
      for (int k = 1; k < 6; k++) s = s + Ad[i * 6 + k] * covTinc[k * 6 + j];
      tmp[i * 6 + j] = s;
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = tmp[i * 6 + 0] * Ad[j * 6 + 0];
      for (int k = 1; k < 6; k++) s = s + tmp[i * 6 + k] * Ad[j * 6 + k];
      covT2[i * 6 + j] = covT1[i * 6 + j] + s;
    }
}

__host__ __device__ inline int pose_smem_doubles(int cap_pt, int cap_ln) { return cap_pt + cap_ln + (cap_pt + cap_ln + 1) / 2; }

__global__ void __launch_bounds__(POSE_WARPS * 32, POSE_MIN_BLOCKS)
pose_optimize_kernel(DevTrackObs obs, const gfpl_frame_state* __restrict__ state, const double* __restrict__ DT_ini,
                     int overload, PoseRaw* __restrict__ raw, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ double s_res_all[];  // [POSE_WARPS][pose_smem_doubles(cap_pt, cap_ln)]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * POSE_WARPS + warp;
  if (p >= obs.B) return;
  // per warp: cap_pt + cap_ln residuals, then (cap_pt + cap_ln) / 2 doubles = cap_pt + cap_ln words of select scratch
  double* s_res = s_res_all + (size_t)warp * pose_smem_doubles(obs.cap_pt, obs.cap_ln);
  PairObs o;
  o.npt = obs.n_pt[p]; o.nln = obs.n_ln[p];
  o.pts = obs.pt + (size_t)p * obs.cap_pt; o.pin = obs.pt_inlier + (size_t)p * obs.cap_pt;
  o.lns = obs.ln + (size_t)p * obs.cap_ln; o.lin = obs.ln_inlier + (size_t)p * obs.cap_ln;
  const gfpl_frame_state& st = state[p];

  double DT[16], H[36], err = 0.0;
  const double* seed = (overload == GFPL_POSE_SEED_GIVEN && DT_ini) ? DT_ini + (size_t)p * 16 : st.DT;
#pragma unroll
  for (int i = 0; i < 16; i++) DT[i] = seed[i];
  int n_inliers_pt = o.npt, n_inliers_ls = o.nln, n_inliers = o.npt + o.nln;
  int status = GFPL_POSE_OK, it1 = 0, it2 = 0;
  bool cov_from_H = false;  // DT_cov ends as H^-1 of the refinement or as zero (:1889-1906)
  if (n_inliers > cfg.min_features) {
    // two GN stages share one inlined loop body: stage 0 = first estimate on a copy (5 iterations) followed by
    // the MAD outlier pass, stage 1 = refinement restarting from the seed (D8) (10 iterations)
    double DTw[16];
    for (int stage = 0; stage < 2; stage++) {
#pragma unroll
      for (int i = 0; i < 16; i++) DTw[i] = DT[i];
      const int iters = warp_gauss_newton(DTw, o, cam, cfg, stage == 0 ? cfg.max_iters : cfg.max_iters_ref, H, &err);
      if (stage == 1) {
        it2 = iters;
        cov_from_H = true;
#pragma unroll
        for (int i = 0; i < 16; i++) DT[i] = DTw[i];
        break;
      }
      it1 = iters;
      if (!is_finite16(DTw)) {
        eye4(DT);
        status |= GFPL_POSE_NONFINITE_1;
        break;
      }
      int dp, dl;
      warp_remove_outliers(DTw, o, cam, cfg, s_res, &dp, &dl);
      n_inliers_pt -= dp; n_inliers_ls -= dl; n_inliers -= dp + dl;
      if (!(n_inliers > cfg.min_features)) {
        eye4(DT);
        status |= GFPL_POSE_FEW_INLIERS;
        break;
      }
    }
  } else {
    eye4(DT);
    status |= GFPL_POSE_FEW_FEATURES;
  }

  if (lane != 0) return;
  PoseRaw& w = raw[p];
#pragma unroll
  for (int i = 0; i < 16; i++) w.DT[i] = DT[i];
#pragma unroll
  for (int i = 0; i < 36; i++) w.H[i] = H[i];
  w.err = err;
  w.status = status; w.cov_from_H = cov_from_H ? 1 : 0;
  w.n_inliers = n_inliers; w.n_inliers_pt = n_inliers_pt; w.n_inliers_ls = n_inliers_ls;
  w.iters_first = it1; w.iters_ref = it2;
}

// serial tail of optimizePose (:1984-2028): DT_cov = H^-1, inverse_se3, motion-step gate, Tfw, eigenvalues of
// DT_cov, unccomp_se3.  One THREAD per pair (the GN kernel spends a warp per pair; this part has no
// parallelism inside a pair, so it runs 32 pairs per warp here).
__global__ void __launch_bounds__(128)
pose_finalize_kernel(int B, const PoseRaw* __restrict__ raw, const gfpl_frame_state* __restrict__ state, int overload,
                     gfpl_pose_result* __restrict__ out, gfpl_config cfg) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  const PoseRaw& w = raw[p];
  const gfpl_frame_state& st = state[p];
  gfpl_pose_result& r = out[p];
  int status = w.status;
  double DT[16], DT_cov[36];
  for (int i = 0; i < 16; i++) DT[i] = w.DT[i];
  if (w.cov_from_H) inverse6(w.H, DT_cov);
  else
    for (int i = 0; i < 36; i++) DT_cov[i] = 0.0;
  bool accept = false;
  double DTinv[16];
  if (is_finite16(DT)) {
    inverse_se3(DT, DTinv);
    accept = true;
    if (overload == GFPL_POSE_SEED_GIVEN) {
      const double tn = sqrt(DTinv[3] * DTinv[3] + DTinv[7] * DTinv[7] + DTinv[11] * DTinv[11]);
      if (!(tn < cfg.motion_step_th * (st.time_stamp_curr - st.time_stamp))) {
        accept = false;
        status |= GFPL_POSE_STEP_REJECTED;
      }
    }
  } else
    status |= GFPL_POSE_NONFINITE_2;
  for (int i = 0; i < 36; i++) r.DT_cov[i] = DT_cov[i];
  sym_eigenvalues<6>(DT_cov, r.DT_cov_eig);
  if (accept) {
    double Tfw[16];
    matmul<4, 4, 4>(st.Tfw, DTinv, Tfw);
    for (int i = 0; i < 16; i++) { r.DT[i] = DTinv[i]; r.Tfw[i] = Tfw[i]; }
    unccomp_se3(st.Tfw, st.Tfw_cov, DT_cov, r.Tfw_cov);
    r.err_norm = w.err;
  } else {
    for (int i = 0; i < 16; i++) { r.DT[i] = (i % 5 == 0) ? 1.0 : 0.0; r.Tfw[i] = st.Tfw[i]; }
    for (int i = 0; i < 36; i++) r.Tfw_cov[i] = st.Tfw_cov[i];
    r.err_norm = -1.0;
  }
  r.status = status;
  r.n_inliers = w.n_inliers; r.n_inliers_pt = w.n_inliers_pt; r.n_inliers_ls = w.n_inliers_ls;
  r.iters_first = w.iters_first; r.iters_ref = w.iters_ref;
}

int launch_pose_optimize(gfpl_ctx* ctx, DevTrackObs obs, const gfpl_frame_state* state, const double* DT_ini,
                         int overload, gfpl_pose_result* out) {
  if (obs.B <= 0) return GFPL_OK;
  // the hand-off buffer lives in the (stream-ordered) default pool; it is tiny: 456 B per pair
  PoseRaw* raw = nullptr;
  GFPL_CUDA_OK(ctx, cudaMallocAsync((void**)&raw, (size_t)obs.B * sizeof(PoseRaw), ctx->stream));
  const size_t smem = (size_t)POSE_WARPS * pose_smem_doubles(obs.cap_pt, obs.cap_ln) * sizeof(double);
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(pose_optimize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (obs.B + POSE_WARPS - 1) / POSE_WARPS;
  pose_optimize_kernel<<<grid, POSE_WARPS * 32, smem, ctx->stream>>>(obs, state, DT_ini, overload, raw, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  pose_finalize_kernel<<<(obs.B + 127) / 128, 128, 0, ctx->stream>>>(obs.B, raw, state, overload, out, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  GFPL_CUDA_OK(ctx, cudaFreeAsync(raw, ctx->stream));
  return GFPL_OK;
}

// ---- P1 alone: H, g, e at given poses ------------------------------------------------------------------
__global__ void __launch_bounds__(POSE_WARPS * 32)
pose_hg_kernel(DevTrackObs obs, const double* __restrict__ DT_all, double* __restrict__ Hout,
               double* __restrict__ gout, double* __restrict__ eout, gfpl_camera cam, gfpl_config cfg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * POSE_WARPS + warp;
  if (p >= obs.B) return;
  PairObs o;
  o.npt = obs.n_pt[p]; o.nln = obs.n_ln[p];
  o.pts = obs.pt + (size_t)p * obs.cap_pt; o.pin = obs.pt_inlier + (size_t)p * obs.cap_pt;
  o.lns = obs.ln + (size_t)p * obs.cap_ln; o.lin = obs.ln_inlier + (size_t)p * obs.cap_ln;
  double DT[16], H[36], g[6], e;
#pragma unroll
  for (int i = 0; i < 16; i++) DT[i] = DT_all[(size_t)p * 16 + i];
  warp_optimize_functions(DT, o, cam, cfg, H, g, &e);
  if (lane == 0) {
    for (int i = 0; i < 36; i++) Hout[(size_t)p * 36 + i] = H[i];
    for (int i = 0; i < 6; i++) gout[(size_t)p * 6 + i] = g[i];
    eout[p] = e;
  }
}

int launch_pose_hg(gfpl_ctx* ctx, const double* DT, const DevTrackObs& obs, double* H, double* g, double* e) {
  if (obs.B <= 0) return GFPL_OK;
  const int grid = (obs.B + POSE_WARPS - 1) / POSE_WARPS;
  pose_hg_kernel<<<grid, POSE_WARPS * 32, 0, ctx->stream>>>(obs, DT, H, g, e, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
