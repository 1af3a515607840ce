// k_pose.cu — K4: the Gauss-Newton pose optimiser of StVO::StereoFrameHandler for sm_100a.
//
// Replaces optimizeFunctions (reference stereoFrameHandler.cpp:2118-2245), gaussNewtonOptimization
// (:2032-2056), removeOutliers (:2058-2116) + vector_stdv_mad (auxiliar.cpp:521-537) and both
// optimizePose overloads (:1864-2030).
//
// A group of POSE_GROUP lanes (16: two pairs per warp) owns one frame pair for the whole optimisation (both GN stages
// and the outlier pass), so the kernel is persistent across iterations: the observation lists are read from L1/L2, the
// pose and the 28 normal-equation partials (21 lower-triangle H + 6 g + 1 e) live in registers, and each iteration ends
// in a fixed xor-butterfly of fp64 shuffles inside the group — every lane ends with bit-identical H, g, e and the result
// is run-to-run reproducible.  All lanes then run the 6x6 LDL^T solve and the SE(3) update redundantly (no broadcast):
// that serial part costs a warp the same instructions whether it serves one pair or several, which is why a warp serves
// several.  The groups of a warp stay convergent (full-mask shuffles): a group that has stopped iterating is predicated
// off, the warp leaves a loop when all its groups have.  FP64-ALU / latency bound: flops per iteration ~ 114 per point
// + 212 per line.
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

#ifndef POSE_WARPS_N
#define POSE_WARPS_N 2
#endif
constexpr int POSE_WARPS = POSE_WARPS_N;
#ifndef POSE_GROUP
#define POSE_GROUP 16     // lanes per frame pair
#endif
#ifndef POSE_PF_DIST
#define POSE_PF_DIST 1     // rounds between the L1 prefetch of an observation record and its use
#endif
#ifndef POSE_CARVEOUT
#define POSE_CARVEOUT 100  // per cent of the SM's L1 / shared memory configured as shared memory
#endif
#ifndef POSE_MAX_REGS
#define POSE_MAX_REGS 168  // 6 CTAs = 12 warps per SM, nothing in local memory (144 = 14 warps: 0.92 ms against 0.89 ms)
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

// Branch-free reciprocal and reciprocal square root: the hardware seed (MUFU.RCP64H / MUFU.RSQ64H, > 20 good bits)
// and one third-order correction, i.e. a relative error far below 2^-53 before the final rounding.  `1.0 / x` and
// sqrt() compile to the same seed plus a longer refinement and a slow-path call behind a branch (five of those per point
// in the first version of this file).  Zero, infinite and NaN arguments end as NaN or Inf; the callers' guards
// (max(homog_th, .)) take those cases before the value is used.
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double rsqrt_nr(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double e = fma(-s, y * y, 1.0);
  return fma(y * e, fma(e, 0.375, 0.5), y);
}

// This translation unit is tolerance-compared (1e-9) and compiled with FMA contraction: divisions that the
// reference performs element-wise (J/max(eps,r), fx*X/Z, ...) are done as one reciprocal and multiplies.

// point term (stereoFrameHandler.cpp:2131-2167); returns the residual norm
__device__ __forceinline__ double point_residual(const double* DT, const gfpl_pt_obs& o, const gfpl_camera& cam,
                                                 double* P_, double* dxy) {
  se3_apply(DT, o.P, P_);
  const double iz = rcp_nr(P_[2]);
  dxy[0] = (cam.cx + cam.fx * P_[0] * iz) - o.pl_obs[0];
  dxy[1] = (cam.cy + cam.fy * P_[1] * iz) - o.pl_obs[1];
  const double s = dxy[0] * dxy[0] + dxy[1] * dxy[1];
  return (s == 0.0) ? 0.0 : s * rsqrt_nr(s);
}

// line term (:2171-2235)
__device__ __forceinline__ double line_residual(const double* DT, const gfpl_ln_obs& o, const gfpl_camera& cam,
                                                double* sP_, double* eP_, double* dse) {
  se3_apply(DT, o.sP, sP_);
  se3_apply(DT, o.eP, eP_);
  const double izs = rcp_nr(sP_[2]), ize = rcp_nr(eP_[2]);
  const double su = cam.cx + cam.fx * sP_[0] * izs, sv = cam.cy + cam.fy * sP_[1] * izs;
  const double eu = cam.cx + cam.fx * eP_[0] * ize, ev = cam.cy + cam.fy * eP_[1] * ize;
  dse[0] = o.le_obs[0] * su + o.le_obs[1] * sv + o.le_obs[2];
  dse[1] = o.le_obs[0] * eu + o.le_obs[1] * ev + o.le_obs[2];
  const double s = dse[0] * dse[0] + dse[1] * dse[1];
  return (s == 0.0) ? 0.0 : s * rsqrt_nr(s);
}


// H is symmetric positive definite whenever the pose is observable, and then the pivoted and the unpivoted LDL^T are the
// same solve up to rounding (both backward stable; this translation unit is compared at 1e-9).  The unpivoted
// factorisation needs no row / column exchanges — the predicated swaps above are ~4/5 of the instructions of the
// pivoted routine, executed by every lane in every GN iteration.  It reports failure when a pivot is not safely positive
// (relative to the largest diagonal entry); the caller then runs the pivoted routine, which is what Eigen would do.
__device__ __forceinline__ bool ldlt6_solve_nopivot(const double* Hp /*lower triangle, packed*/, const double* g, double* x) {
  double L[36], d[6], id[6];
  double dmax = Hp[0];
#pragma unroll
  for (int i = 1; i < 6; i++) dmax = fmax(dmax, Hp[tri(i, i)]);
  const double floor_ = 1e-11 * dmax;
  bool ok = dmax > 0.0 && dmax < 1e300;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double t[6];
#pragma unroll
    for (int j = 0; j < k; j++) t[j] = L[k * 6 + j] * d[j];
    double dk = Hp[tri(k, k)];
#pragma unroll
    for (int j = 0; j < k; j++) dk = fma(-L[k * 6 + j], t[j], dk);
    d[k] = dk;
    ok = ok && (dk > floor_);
    id[k] = rcp_nr(dk);
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      double v = Hp[tri(i, k)];
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

// expmap_se3 (auxiliar.cpp:165-182) for this translation unit: the reference divides 27 matrix entries by theta
// element-wise; here 1/theta is formed once (rcp_nr) and multiplied in — every fp64 division is ~25 instructions with a
// slow-path call, and all lanes of a group execute this serial part in every iteration (ncu: 13 % of the kernel's
// instructions before).  Tolerance-compared like the rest of the file.
__device__ __forceinline__ void expmap_se3_fast(const double* x, double* T) {
  const double w0 = x[3], w1 = x[4], w2 = x[5];
  double t0 = x[0], t1 = x[1], t2 = x[2];
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
  if (!(th2 < 0.000001 * 0.000001)) {   // theta >= 1e-6 (NaN takes this branch, as in the reference)
    const double ith = rsqrt_nr(th2), theta = th2 * ith;
    const double a0 = w0 * ith, a1 = w1 * ith, a2 = w2 * ith;          // unit axis: s = skew(a)
    double sn, cs;
    sincos(theta, &sn, &cs);
    const double omc = 1.0 - cs, tms = theta - sn;
    // s = skew(a), ss = s * s = a a^T - |a|^2 I
    const double s[9] = {0.0, -a2, a1, a2, 0.0, -a0, -a1, a0, 0.0};
    const double ss[9] = {-(a1 * a1 + a2 * a2), a0 * a1, a0 * a2, a1 * a0, -(a0 * a0 + a2 * a2), a1 * a2, a2 * a0, a2 * a1, -(a0 * a0 + a1 * a1)};
    const double c1 = omc * ith, c2 = tms * ith;
    double V[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = (I + s[i] * sn) + ss[i] * omc;
      V[i] = (I + s[i] * c1) + ss[i] * c2;
    }
    const double n0 = V[0] * t0 + V[1] * t1 + V[2] * t2;
    const double n1 = V[3] * t0 + V[4] * t1 + V[5] * t2;
    const double n2 = V[6] * t0 + V[7] * t1 + V[8] * t2;
    t0 = n0; t1 = n1; t2 = n2;
  }
  T[0] = R[0]; T[1] = R[1]; T[2] = R[2]; T[3] = t0;
  T[4] = R[3]; T[5] = R[4]; T[6] = R[5]; T[7] = t1;
  T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = t2;
  T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
}

// pivoted fallback, out of line (gfpl_math.cuh ldlt6_solve, local-memory arrays): only for H that is not safely SPD
__device__ __noinline__ void ldlt6_solve_pivoted_slow(const double* H, const double* g, double* x) { ldlt6_solve(H, g, x); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct PairObs {
  int npt, nln;
  const gfpl_pt_obs* pts;
  const gfpl_ln_obs* lns;
  uint8_t* pin;
  uint8_t* lin;
  const uint32_t* plist;   // indices of the inlier points, ascending (shared memory), or nullptr: every point, flags in pin
  int nplist;
};

// ---- reductions inside a group of G lanes (full-mask shuffles: the groups of a warp call them together) -------------
template <int G>
__device__ __forceinline__ int group_sum(int v) {
  if (G == 32) return __reduce_add_sync(0xffffffffu, v);
#pragma unroll
  for (int s = G / 2; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_max(uint32_t v) {
  if (G == 32) return __reduce_max_sync(0xffffffffu, v);
#pragma unroll
  for (int s = G / 2; s >= 1; s >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

struct PoseConst {  // loop invariants of the observation terms
  double hth, inv_hth, fx_over_hth;
};

// The pose under optimisation lives in the pair's shared-memory slot (12 doubles, rows of [R | t]).  The observation
// loop re-reads it per observation with 128-bit shared loads (same address in every lane of the group: a broadcast)
// instead of carrying it in 24 registers across the loop — with the 56 registers of partial sums that is what lets the
// kernel run 14 warps per SM without local-memory traffic.  Every lane of a group stores the same values.
__device__ __forceinline__ void pose_load(const double* DTs, double* T /*16*/) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(DTs);
#pragma unroll
  for (int i = 0; i < 6; i++)
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(T[2 * i]), "=d"(T[2 * i + 1]) : "r"(a + 16u * i));
  T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;   // poses are SE(3)
}
__device__ __forceinline__ void pose_store(double* DTs, const double* T) {
#pragma unroll
  for (int i = 0; i < 12; i++) DTs[i] = T[i];
}

// point term (stereoFrameHandler.cpp:2131-2167): J (already divided by max(homog_th, r)), r and the weight.  Branch-free,
// so that two observations evaluated back to back interleave in one basic block.  `on` false (outlier, or past the end
// of the list) gives an all-zero term.
// WEIGHT_ONLY: an `off` slot is known to hold a copy of a valid record (the padding slot of a round over the inlier
// list), so a zero weight alone removes it — every term it contributes is a finite value times w.
template <bool WEIGHT_ONLY>
__device__ __forceinline__ void point_front(const double* DTs, const gfpl_pt_obs& ob, bool on, const gfpl_camera& cam,
                                            const PoseConst& c, double* J, double& r, double& w) {
  double DT[16], P_[3];
  pose_load(DTs, DT);
  se3_apply(DT, ob.P, P_);
  const double iz = rcp_nr(P_[2]);
  const double dx = (cam.cx + cam.fx * P_[0] * iz) - ob.pl_obs[0];
  const double dy = (cam.cy + cam.fy * P_[1] * iz) - ob.pl_obs[1];
  const double s = dx * dx + dy * dy;
  const double fgz2 = (P_[2] * P_[2] > c.hth) ? cam.fx * iz * iz : c.fx_over_hth;   // fx / max(homog_th, Z^2)
  const double rs = rsqrt_nr(s);
  r = (s == 0.0) ? 0.0 : s * rs;                       // |err|
  const double idn = (r > c.hth) ? rs : c.inv_hth;     // 1 / max(homog_th, r)
  jac6(fgz2, P_[0], P_[1], P_[2], dx, dy, J);
#pragma unroll
  for (int i = 0; i < 6; i++) J[i] = (WEIGHT_ONLY || on) ? J[i] * idn : 0.0;
  w = on ? rcp_nr(1.0 + s * ob.sigma2) : 0.0;
  r = (WEIGHT_ONLY || on) ? r : 0.0;
}
// line term (:2171-2235)
__device__ __forceinline__ void line_term(const double* DTs, const gfpl_ln_obs& ob, const gfpl_camera& cam, const PoseConst& c,
                                          Acc& a) {
  double DT[16], sP_[3], eP_[3], Js[6], Je[6], J[6];
  pose_load(DTs, DT);
  se3_apply(DT, ob.sP, sP_);
  se3_apply(DT, ob.eP, eP_);
  const double izs = rcp_nr(sP_[2]), ize = rcp_nr(eP_[2]);
  const double su = cam.cx + cam.fx * sP_[0] * izs, sv = cam.cy + cam.fy * sP_[1] * izs;
  const double eu = cam.cx + cam.fx * eP_[0] * ize, ev = cam.cy + cam.fy * eP_[1] * ize;
  const double ds = ob.le_obs[0] * su + ob.le_obs[1] * sv + ob.le_obs[2];
  const double de = ob.le_obs[0] * eu + ob.le_obs[1] * ev + ob.le_obs[2];
  const double s = ds * ds + de * de;
  double ks = cam.fx * izs * izs, ke = cam.fx * ize * ize;
  if (!(sP_[2] * sP_[2] > c.hth)) ks = c.fx_over_hth;
  if (!(eP_[2] * eP_[2] > c.hth)) ke = c.fx_over_hth;
  const double rs = rsqrt_nr(s);
  const double r = (s == 0.0) ? 0.0 : s * rs;
  const double idn = (r > c.hth) ? rs : c.inv_hth;
  jac6(ks, sP_[0], sP_[1], sP_[2], ob.le_obs[0], ob.le_obs[1], Js);
  jac6(ke, eP_[0], eP_[1], eP_[2], ob.le_obs[0], ob.le_obs[1], Je);
#pragma unroll
  for (int i = 0; i < 6; i++) J[i] = (Js[i] * ds + Je[i] * de) * idn;
  const double w = rcp_nr(1.0 + s * ob.sigma2);
  acc_add(a, J, r, w);
}

// optimizeFunctions for one pair by one group of G lanes: a.v = [H lower triangle, packed (21) | g (6) | e] (nothing is
// evaluated for a group that is not `act`: its sums are zero).  Every lane of the group ends with the same values.
// LISTED: the points are walked through the compacted inlier list o.plist (the optimiser); else all points, flags in o.pin
template <int G, bool LISTED>
__device__ __forceinline__ void group_optimize_functions(const double* DTs, const PairObs& o, bool act, const gfpl_camera& cam,
                                                         const PoseConst& c, Acc& a) {
  const int lg = threadIdx.x & (G - 1);
  const int npt = act ? o.npt : 0, nln = act ? o.nln : 0;
#pragma unroll
  for (int i = 0; i < 28; i++) a.v[i] = 0.0;
  int N = 0;
  {
    // two observations per lane and round in one basic block: their dependency chains (transform, reciprocal, projection,
    // reciprocal square root, Jacobian) interleave, which is where a latency-bound warp gets its instruction-level
    // parallelism from; an observation that is an outlier or past the end contributes zeros
    // (the optimiser walks the compacted list of inliers: after the MAD pass a fifth of the points are outliers, and a
    // branch-free term costs the same whether it counts or not)
    const int nwalk = LISTED ? (act ? o.nplist : 0) : npt;
    for (int k = lg; k < nwalk; k += 2 * G) {
      const bool twoB = k + G < nwalk;
      const int k2 = twoB ? k + G : k;
      const int iA = LISTED ? (int)o.plist[k] : k, iB = LISTED ? (int)o.plist[k2] : k2;
      if (k + 2 * G * POSE_PF_DIST < nwalk) {   // the records of a later round: pulled into L1 now, no registers held
        const int k2n = k + 2 * G * POSE_PF_DIST;
        const int k3 = k2n + G < nwalk ? k2n + G : k2n;
        const char* nA = reinterpret_cast<const char*>(&o.pts[LISTED ? (int)o.plist[k2n] : k2n]);
        const char* nB = reinterpret_cast<const char*>(&o.pts[LISTED ? (int)o.plist[k3] : k3]);
        prefetch_l1(nA); prefetch_l1(nA + 48);
        prefetch_l1(nB); prefetch_l1(nB + 48);
      }
      const gfpl_pt_obs A = o.pts[iA], Bq = o.pts[iB];
      const bool onA = LISTED ? true : o.pin[iA] != 0, onB = twoB && (LISTED ? true : o.pin[iB] != 0);
      double JA[6], JB[6], rA, rB, wA, wB;
      point_front<LISTED>(DTs, A, onA, cam, c, JA, rA, wA);
      point_front<LISTED>(DTs, Bq, onB, cam, c, JB, rB, wB);
      acc_add(a, JA, rA, wA);
      acc_add(a, JB, rB, wB);
      N += (onA ? 1 : 0) + (onB ? 1 : 0);
    }
  }
  for (int k = lg; k < nln; k += G) {
    if (!o.lin[k]) continue;
    const gfpl_ln_obs ob = o.lns[k];
    line_term(DTs, ob, cam, c, a);
    N++;
  }
  // fixed xor-butterfly: identical association in every lane, deterministic run to run
#pragma unroll
  for (int s = G / 2; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < 28; i++) a.v[i] = a.v[i] + shfl_xor_f64(0xffffffffu, a.v[i], s);
    N += __shfl_xor_sync(0xffffffffu, N, s);
  }
  a.v[27] = a.v[27] / (double)N;  // e /= (N_l + N_p), 0/0 -> NaN like the reference
}

// Ascending list of the points whose inlier flag is set, for the groups of a warp together (ballot compaction per
// group).  npt = 0 for a group without a problem.  Returns the length.
template <int G>
__device__ __forceinline__ int group_inlier_list(const uint8_t* pin, int npt, uint32_t* list) {
  const int lane = threadIdx.x & 31, lg = lane & (G - 1), gfirst = lane & ~(G - 1);
  const int nmax = __reduce_max_sync(0xffffffffu, npt);
  int base = 0;
  for (int k0 = 0; k0 < nmax; k0 += G) {
    const int k = k0 + lg;
    const bool keep = k < npt && pin[k] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const unsigned gb = G == 32 ? bal : ((bal >> gfirst) & ((1u << (G & 31)) - 1u));
    if (keep) list[base + __popc(gb & ((1u << lg) - 1u))] = (uint32_t)k;
    base += __popc(gb);
  }
  __syncwarp();
  return base;
}

// gaussNewtonOptimization (:2032-2056) for the groups of a warp together; H of the last evaluation is returned for
// DT_cov = H^-1.  A group without a problem to solve (`alive` false) only takes part in the shuffles.
template <int G>
__device__ __forceinline__ int group_gauss_newton(double* DTs /*shared, 12*/, const PairObs& o, bool alive, const gfpl_camera& cam,
                                                  const gfpl_config& cfg, const PoseConst& pc, int max_iters, double* Hs /*shared, 21*/,
                                                  double* err_) {
  const int lg = threadIdx.x & (G - 1);
  double inc[6];
  double err = 0.0, err_prev = 999999999.9;
  if (alive && lg == 0) {
#pragma unroll
    for (int i = 0; i < 21; i++) Hs[i] = 0.0;
  }
  bool act = alive;
  int iters = 0;
  for (int it = 0; it < max_iters; it++) {
    if (!__any_sync(0xffffffffu, act)) break;
    Acc a;
    group_optimize_functions<G, true>(DTs, o, act, cam, pc, a);
    if (act) {  // H of the last evaluation is kept in the pair's shared-memory slot, not in 42 registers across the loop
      err = a.v[27];
      if (lg == 0) {
#pragma unroll
        for (int i = 0; i < 21; i++) Hs[i] = a.v[i];
      }
    }
    bool go = act && !((fabs(err - err_prev) < cfg.min_error_change) || (err < cfg.min_error));
    const bool solved = ldlt6_solve_nopivot(a.v, a.v + 21, inc);
    double DT[16];
    pose_load(DTs, DT);
    const bool upd = go;
    if (go) {
      if (!solved) {
        double Hm[36], gm[6], xm[6];
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) Hm[i * 6 + j] = a.v[i >= j ? tri(i, j) : tri(j, i)];
#pragma unroll
        for (int i = 0; i < 6; i++) gm[i] = a.v[21 + i];
        ldlt6_solve_pivoted_slow(Hm, gm, xm);
#pragma unroll
        for (int i = 0; i < 6; i++) inc[i] = xm[i];
      }
      double E[16], Ei[16];
      expmap_se3_fast(inc, E);
      inverse_se3(E, Ei);
      matmul<4, 4, 4>(DT, Ei, DT);
      double nrm = 0.0;
#pragma unroll
      for (int i = 0; i < 6; i++) nrm = nrm + inc[i] * inc[i];
      if (nrm < 2.220446049250313e-16 * 2.220446049250313e-16) go = false;   // |inc| < eps (:2051), without the square root
      else { err_prev = err; iters++; }
    }
    __syncwarp();               // every lane of the group has read the pose before any lane replaces it
    if (upd) pose_store(DTs, DT);
    __syncwarp();
    act = go;
  }
  if (alive) *err_ = err;
  return iters;
}

// element [rank] of the ascending order of n unsigned keys, without sorting: MSB-first binary search on the key bits,
// one group-wide count per probe.  key k lives at keys[k * STRIDE]; with RESTRICT only the elements whose companion word
// keys[k * STRIDE + 1] equals `hi` take part (ties of the high word of a double, see below).
template <int G, int STRIDE, bool RESTRICT, int TOP_BIT>
__device__ __forceinline__ uint32_t group_select_u32(const uint32_t* keys, int n, int rank, uint32_t hi) {
  const int lg = threadIdx.x & (G - 1);
  uint32_t prefix = 0u;
  for (int bit = TOP_BIT; bit >= 0; bit--) {
    const uint32_t trial = prefix | (1u << bit);
    int cnt = 0;
    for (int k = lg; k < n; k += G) {
      const bool in = !RESTRICT || keys[k * STRIDE + 1] == hi;
      cnt += (in && keys[k * STRIDE] < trial) ? 1 : 0;
    }
    cnt = group_sum<G>(cnt);
    if (cnt <= rank) prefix = trial;
  }
  return prefix;
}

// sum of small non-negative counts (< 65536 per group) inside a group: for two groups per warp ONE warp-wide integer
// reduction serves both (each group's count in its own 16-bit field) instead of four dependent shuffle + add steps
template <int G>
__device__ __forceinline__ int group_sum_small(int v) {
  if (G == 32) return __reduce_add_sync(0xffffffffu, v);
  if (G == 16) {
    const int sh = (threadIdx.x & 16);   // 0 for the warp's first group, 16 for the second
    const unsigned both = __reduce_add_sync(0xffffffffu, (unsigned)v << sh);
    return (int)((both >> sh) & 0xFFFFu);
  }
  return group_sum<G>(v);
}

// element [rank] of the ascending order of the n keys (< 2^31) of a COMPACT array in shared memory, 16-byte aligned and
// padded with 0xFFFFFFFF to a multiple of four; the array is destroyed.  Same MSB-first search as group_select_u32 in two
// phases: the upper SEL_HI_BITS bits against the whole array (128-bit loads, four keys per instruction), then — the keys
// that share those bits with the answer are usually a handful — the candidates are compacted (ballots), each lane takes
// one into a register, and the remaining bits cost a ballot + popc per probe.  More than G candidates (heavy ties): the
// generic loop over the compacted array.
constexpr int SEL_LO_BITS = 15;
template <int G>
__device__ __forceinline__ uint32_t group_select_compact(uint32_t* keys, int n, int rank) {
  const int lane = threadIdx.x & 31, lg = lane & (G - 1), gfirst = lane & ~(G - 1);
  const unsigned gbits = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  const uint4* k4 = reinterpret_cast<const uint4*>(keys);
  const int n4 = (n + 3) >> 2;
  uint32_t prefix = 0u;
  for (int bit = 30; bit >= SEL_LO_BITS; bit--) {
    const uint32_t trial = prefix | (1u << bit);
    int cnt = 0;
    for (int q = lg; q < n4; q += G) {
      const uint4 v = k4[q];
      cnt += (v.x < trial ? 1 : 0) + (v.y < trial ? 1 : 0) + (v.z < trial ? 1 : 0) + (v.w < trial ? 1 : 0);
    }
    cnt = group_sum_small<G>(cnt);
    if (cnt <= rank) prefix = trial;
  }
  // candidates: keys whose upper bits equal the prefix; `below` = keys in lower buckets
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  int m = 0, below = 0;
  for (int k0 = 0; k0 < nmax; k0 += G) {
    const int k = k0 + lg;
    const uint32_t key = k < n ? keys[k] : 0xFFFFFFFFu;
    const bool cand = k < n && ((key ^ prefix) >> SEL_LO_BITS) == 0u;
    const unsigned bc = (__ballot_sync(0xffffffffu, cand) >> gfirst) & gbits;
    const unsigned bb = (__ballot_sync(0xffffffffu, k < n && key < prefix) >> gfirst) & gbits;
    if (cand) keys[m + __popc(bc & ((1u << lg) - 1u))] = key;   // in place: never past the chunk just read
    m += __popc(bc); below += __popc(bb);
  }
  __syncwarp();
  const int r2 = rank - below;
  if (__all_sync(0xffffffffu, m <= G)) {
    const uint32_t ck = lg < m ? keys[lg] : 0xFFFFFFFFu;
#pragma unroll
    for (int bit = SEL_LO_BITS - 1; bit >= 0; bit--) {
      const uint32_t trial = prefix | (1u << bit);
      const int cnt = __popc((__ballot_sync(0xffffffffu, ck < trial) >> gfirst) & gbits);
      if (cnt <= r2) prefix = trial;
    }
  } else {
    for (int bit = SEL_LO_BITS - 1; bit >= 0; bit--) {
      const uint32_t trial = prefix | (1u << bit);
      int cnt = 0;
      for (int k = lg; k < m; k += G) cnt += keys[k] < trial ? 1 : 0;
      cnt = group_sum_small<G>(cnt);
      if (cnt <= r2) prefix = trial;
    }
  }
  __syncwarp();
  return prefix;
}

// 1.4826 * MAD of res[0..n) exactly as vector_stdv_mad (auxiliar.cpp:521-537) computes it:
// median = sorted[n/2]; a = (double)fabsf(res - median) (float truncation); MAD = sorted(a)[n/2].
// Order statistics of non-negative IEEE values are order statistics of their bit patterns, so both selections run on
// 32-bit integer keys: the median on the high words of the doubles (low words only among the elements that tie on the
// high word), the MAD on the float patterns of `a`; both key sets are staged in `scratch` (n words rounded up to a
// multiple of four, 16-byte aligned) for group_select_compact.  n == 0 gives 0.
template <int G>
__device__ __forceinline__ double group_stdv_mad(const double* res, int n, uint32_t* scratch) {
  const int lg = threadIdx.x & (G - 1);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(res);  // little endian: w[2k] low word, w[2k+1] high word
  const int rank = n / 2;
  const int npad = (n + 3) & ~3;
  for (int k = lg; k < npad; k += G) scratch[k] = k < n ? w[2 * k + 1] : 0xFFFFFFFFu;
  __syncwarp();
  const uint32_t hi = group_select_compact<G>(scratch, n, rank);
  int below = 0, same = 0;
  uint32_t mine = 0u;
  for (int k = lg; k < n; k += G) {
    below += w[2 * k + 1] < hi ? 1 : 0;
    same += w[2 * k + 1] == hi ? 1 : 0;
    mine = w[2 * k + 1] == hi ? w[2 * k] : mine;
  }
  below = group_sum_small<G>(below);
  same = group_sum_small<G>(same);
  uint32_t lo = group_max<G>(mine);   // the usual case: one element has that high word and it is the median
  const bool ties = n > 0 && same != 1;
  if (__any_sync(0xffffffffu, ties)) {
    const uint32_t lo2 = group_select_u32<G, 2, true, 31>(w, n, rank - below, hi);
    if (ties) lo = lo2;
  }
  const double median = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  for (int k = lg; k < npad; k += G) scratch[k] = k < n ? __float_as_uint(fabsf((float)(res[k] - median))) : 0xFFFFFFFFu;
  __syncwarp();
  const uint32_t mad = group_select_compact<G>(scratch, n, rank);
  return n == 0 ? 0.0 : 1.4826 * (double)__uint_as_float(mad);
}

// removeOutliers (:2058-2116): returns the number of points / lines flagged
template <int G>
__device__ __forceinline__ void group_remove_outliers(const double* DTs, const PairObs& o, bool alive, const gfpl_camera& cam,
                                                      const gfpl_config& cfg, double* s_res, int* drop_pt, int* drop_ln) {
  const int lg = threadIdx.x & (G - 1);
  const int npt = alive ? o.npt : 0, nln = alive ? o.nln : 0;
  double* res_p = s_res;
  double* res_l = s_res + npt;
  double DT[16];
  pose_load(DTs, DT);
  for (int k = lg; k < npt; k += G) {
    double P_[3], d[2];
    const gfpl_pt_obs ob = o.pts[k];
    const double r = point_residual(DT, ob, cam, P_, d);
    res_p[k] = r * (ob.sigma2 * rsqrt_nr(ob.sigma2));   // r * sqrt(sigma2), sigma2 > 0
  }
  for (int k = lg; k < nln; k += G) {
    double sP_[3], eP_[3], d[2];
    const gfpl_ln_obs ob = o.lns[k];
    const double r = line_residual(DT, ob, cam, sP_, eP_, d);
    res_l[k] = r * (ob.sigma2 * rsqrt_nr(ob.sigma2));
  }
  __syncwarp();
  uint32_t* scratch = reinterpret_cast<uint32_t*>(s_res + ((npt + nln + 1) & ~1));  // [max(npt, nln) + 3] words behind the residuals, 16-byte aligned
  const double th_p = cfg.inlier_k * group_stdv_mad<G>(res_p, npt, scratch);
  const double th_l = cfg.inlier_k * group_stdv_mad<G>(res_l, nln, scratch);
  int dp = 0, dl = 0;
  for (int k = lg; k < npt; k += G)
    if (res_p[k] > th_p) { o.pin[k] = 0; dp++; }
  for (int k = lg; k < nln; k += G)
    if (res_l[k] > th_l) { o.lin[k] = 0; dl++; }
  *drop_pt = group_sum<G>(dp);
  *drop_ln = group_sum<G>(dl);
  __syncwarp();
}

// an even number of doubles (16-byte aligned slots): residuals, select scratch / inlier list, padding of the scratch
__host__ __device__ inline int pose_smem_doubles(int cap_pt, int cap_ln) { return (cap_pt + cap_ln + (cap_pt + cap_ln + 1) / 2 + 4 + 1) & ~1; }

template <int G>
__global__ void __maxnreg__(POSE_MAX_REGS)
pose_optimize_kernel(DevTrackObs obs, const gfpl_frame_state* __restrict__ state, const double* __restrict__ DT_ini,
                     int overload, PoseRaw* __restrict__ raw, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ __align__(16) double s_res_all[];  // [POSE_WARPS * 32 / G][pose_smem_doubles(cap_pt, cap_ln)]
  constexpr int PER_WARP = 32 / G;
  __shared__ double s_H[POSE_WARPS * PER_WARP][21];   // H of the last evaluation, per pair
  __shared__ __align__(16) double s_DT[POSE_WARPS * PER_WARP][12];   // the pose under optimisation, per pair
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp * PER_WARP + lane / G;                 // pair slot inside the CTA
  if ((blockIdx.x * POSE_WARPS + warp) * PER_WARP >= obs.B) return;   // the whole warp is past the batch
  const int p_slot = blockIdx.x * (POSE_WARPS * PER_WARP) + slot;
  const bool exists = p_slot < obs.B;       // a group past the batch only takes part in the shuffles
  const int p = exists ? p_slot : obs.B - 1;
  // per pair: cap_pt + cap_ln residuals, then (cap_pt + cap_ln) / 2 doubles = cap_pt + cap_ln words of select scratch
  double* s_res = s_res_all + (size_t)slot * pose_smem_doubles(obs.cap_pt, obs.cap_ln);
  PairObs o;
  o.npt = obs.n_pt[p]; o.nln = obs.n_ln[p];
  o.pts = obs.pt + (size_t)p * obs.cap_pt; o.pin = obs.pt_inlier + (size_t)p * obs.cap_pt;
  o.lns = obs.ln + (size_t)p * obs.cap_ln; o.lin = obs.ln_inlier + (size_t)p * obs.cap_ln;
  o.plist = nullptr; o.nplist = 0;
  const gfpl_frame_state& st = state[p];

  double DT[16], err = 0.0;
  double* Hs = s_H[slot];
  double* DTs = s_DT[slot];
  const double* seed = (overload == GFPL_POSE_SEED_GIVEN && DT_ini) ? DT_ini + (size_t)p * 16 : st.DT;
  if ((lane & (G - 1)) == 0) {
#pragma unroll
    for (int i = 0; i < 21; i++) Hs[i] = 0.0;
  }
  int n_inliers_pt = o.npt, n_inliers_ls = o.nln, n_inliers = o.npt + o.nln;
  int status = GFPL_POSE_OK, it1 = 0, it2 = 0;
  bool cov_from_H = false;  // DT_cov ends as H^-1 of the refinement or as zero (:1889-1906)
  bool alive = exists && n_inliers > cfg.min_features;
  if (!alive) status |= GFPL_POSE_FEW_FEATURES;
  // stage 0: first estimate on a copy of the seed (max_iters iterations), then the MAD outlier pass; stage 1: refinement
  // restarting from the seed (D8) (max_iters_ref iterations).  The stages are iterations of ONE loop that is not
  // unrolled: two inlined copies of the optimiser doubled the code, and the warps of an SM are spread over all of it
  // (ncu: 12 % of the stall samples were instruction-cache misses).
  // (the select scratch behind the residuals doubles as the inlier list: the two are never live together)
  uint32_t* plist = reinterpret_cast<uint32_t*>(s_res + obs.cap_pt + obs.cap_ln);
  o.plist = plist;
  PoseConst pc;   // loop invariants of the observation terms, once per kernel (two divisions)
  pc.hth = cfg.homog_th; pc.inv_hth = 1.0 / cfg.homog_th; pc.fx_over_hth = cam.fx / cfg.homog_th;
#pragma unroll 1
  for (int stage = 0; stage < 2; stage++) {
    o.nplist = group_inlier_list<G>(o.pin, alive ? o.npt : 0, plist);
    pose_store(DTs, seed);
    const int it = group_gauss_newton<G>(DTs, o, alive, cam, cfg, pc, stage == 0 ? cfg.max_iters : cfg.max_iters_ref, Hs, &err);
    if (stage == 0) {
      it1 = it;
      pose_load(DTs, DT);
      if (alive && !is_finite16(DT)) { status |= GFPL_POSE_NONFINITE_1; alive = false; }
      int dp, dl;
      group_remove_outliers<G>(DTs, o, alive, cam, cfg, s_res, &dp, &dl);
      if (alive) {
        n_inliers_pt -= dp; n_inliers_ls -= dl; n_inliers -= dp + dl;
        if (!(n_inliers > cfg.min_features)) { status |= GFPL_POSE_FEW_INLIERS; alive = false; }
      }
    } else
      it2 = it;
  }
  if (alive) { cov_from_H = true; pose_load(DTs, DT); }
  else eye4(DT);

  if ((lane & (G - 1)) != 0 || !exists) return;
  PoseRaw& w = raw[p];
#pragma unroll
  for (int i = 0; i < 16; i++) w.DT[i] = DT[i];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) w.H[i * 6 + j] = Hs[i >= j ? tri(i, j) : tri(j, i)];   // written by this lane
  w.err = err;
  w.status = status; w.cov_from_H = cov_from_H ? 1 : 0;
  w.n_inliers = n_inliers; w.n_inliers_pt = n_inliers_pt; w.n_inliers_ls = n_inliers_ls;
  w.iters_first = it1; w.iters_ref = it2;
}

// serial tail of optimizePose (:1984-2028): DT_cov = H^-1, inverse_se3, motion-step gate, Tfw, eigenvalues of DT_cov,
// unccomp_se3.  No parallelism inside these pieces, but they are independent of each other once DT_cov exists: FOUR
// lanes per pair — every lane inverts H itself (same operations, same bits), then lane 0 runs the Jacobi iteration for
// the eigenvalues, lane 1 the pose / gate / counters, lanes 2 and 3 three rows each of the covariance propagation.
// One thread per pair took 96 us per 8192 pairs (a chain of ~190 k cycles on 64 SMs), 72 us with the Jacobi matrix in registers.
constexpr int FINALIZE_LANES = 4;
#ifndef FINALIZE_THREADS_N
#define FINALIZE_THREADS_N 64
#endif
constexpr int FINALIZE_THREADS = FINALIZE_THREADS_N;

// Eigenvalues of a symmetric 6x6 matrix by the cyclic Jacobi iteration of gfpl_math.cuh, matrix in registers, with the
// divisions and square roots of a rotation (three divisions and two square roots, each a ~150-cycle dependent sequence:
// the iteration was ~45 us of the finalize kernel's ~80 us chain) replaced by the reciprocal / reciprocal-square-root
// refinements above.  A rotation that is orthogonal to ~1e-16 instead of exactly rounded changes the eigenvalues by
// O(eps |A|); DT_cov_eig is compared at 1e-7 relative to the largest eigenvalue.  Unsorted diagonal -> ev.
__device__ __forceinline__ void sym_eigenvalues6_fast(const double* Ain, double* ev) {
  constexpr int N = 6;
  double A[N * N];
#pragma unroll
  for (int i = 0; i < N * N; i++) A[i] = Ain[i];
  for (int sweep = 0; sweep < 16; sweep++) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < N; p++)
#pragma unroll
      for (int q = p + 1; q < N; q++) off = off + A[p * N + q] * A[p * N + q];
    if (off == 0.0) break;
#pragma unroll
    for (int p = 0; p < N; p++)
#pragma unroll
      for (int q = p + 1; q < N; q++) {
        const double apq = A[p * N + q];
        // the off-diagonal entries shrink quadratically and pass through the subnormal range on their way to zero, where
        // the reciprocal seed flushes: below 1e-280 the rotation is the identity to every bit that matters (entries zeroed)
        if (fabs(apq) < 1e-280) { A[p * N + q] = 0.0; A[q * N + p] = 0.0; }
        else {   // (NaN takes this branch and propagates, as in the exact iteration)
          const double app = A[p * N + p], aqq = A[q * N + q];
          const double tau = (aqq - app) * 0.5 * rcp_nr(apq);
          const double at = fabs(tau);
          double t;
          if (at < 1e100) {
            const double s1 = fma(tau, tau, 1.0);
            t = rcp_nr(at + s1 * rsqrt_nr(s1));          // 1 / (|tau| + sqrt(1 + tau^2))
          } else
            t = 0.5 * rcp_nr(at);                         // (tau^2 would overflow; also Inf / NaN from a tiny apq: t -> 0)
          t = (tau >= 0.0) ? t : -t;
          if (at > 1e300) t = 0.0;                        // apq so small that the rotation is the identity (NaN stays NaN)
          const double c = rsqrt_nr(fma(t, t, 1.0)), sn = t * c;
#pragma unroll
          for (int k = 0; k < N; k++) {
            const double akp = A[k * N + p], akq = A[k * N + q];
            A[k * N + p] = c * akp - sn * akq;
            A[k * N + q] = sn * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < N; k++) {
            const double apk = A[p * N + k], aqk = A[q * N + k];
            A[p * N + k] = c * apk - sn * aqk;
            A[q * N + k] = sn * apk + c * aqk;
          }
          A[p * N + q] = 0.0;
          A[q * N + p] = 0.0;
        }
      }
  }
#pragma unroll
  for (int i = 0; i < N; i++) ev[i] = A[i * N + i];
}

// rows [r0, r1) of unccomp_se3 (auxiliar.cpp:221-238): covT2 = covT1 + Ad(T1) * covTinc * Ad(T1)^T
__device__ __noinline__ void unccomp_se3_rows(const double* T1, const double* covT1, const double* covTinc, int r0, int r1, double* covT2) {
  double Ad[36];
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
  for (int i = r0; i < r1; i++) {
    double tmp[6];
    for (int j = 0; j < 6; j++) {
      double s = Ad[i * 6 + 0] * covTinc[0 * 6 + j];
      for (int k = 1; k < 6; k++) s = s + Ad[i * 6 + k] * covTinc[k * 6 + j];
      tmp[j] = s;
    }
    for (int j = 0; j < 6; j++) {
      double s = tmp[0] * Ad[j * 6 + 0];
      for (int k = 1; k < 6; k++) s = s + tmp[k] * Ad[j * 6 + k];
      covT2[i * 6 + j] = covT1[i * 6 + j] + s;
    }
  }
}

__global__ void __launch_bounds__(FINALIZE_THREADS)
pose_finalize_kernel(int B, const PoseRaw* __restrict__ raw, const gfpl_frame_state* __restrict__ state, int overload,
                     gfpl_pose_result* __restrict__ out, gfpl_config cfg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = t / FINALIZE_LANES, sub = t % FINALIZE_LANES;
  if (p >= B) return;
  const PoseRaw& w = raw[p];
  const gfpl_frame_state& st = state[p];
  gfpl_pose_result& r = out[p];
  int status = w.status;
  double DT[16], DT_cov[36];
  for (int i = 0; i < 16; i++) DT[i] = w.DT[i];
  if (sub != 1) {   // (lane 1 does not read DT_cov)
    if (w.cov_from_H) inverse6(w.H, DT_cov);
    else
      for (int i = 0; i < 36; i++) DT_cov[i] = 0.0;
  }
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
  if (sub == 0) {
    for (int i = 0; i < 36; i++) r.DT_cov[i] = DT_cov[i];
    // eigenvalues of DT_cov: the cyclic Jacobi iteration with the matrix in registers (the out-of-line routine keeps it in
    // local memory: every rotation then waits for ~50 dependent local loads), ascending insertion sort as before
    double ev[6];
    sym_eigenvalues6_fast(DT_cov, ev);
#pragma unroll
    for (int i = 1; i < 6; i++) {
#pragma unroll
      for (int j = i; j > 0; j--) {
        const double a = ev[j - 1], b = ev[j];
        const bool sw = a > b;
        ev[j - 1] = sw ? b : a; ev[j] = sw ? a : b;
      }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) r.DT_cov_eig[i] = ev[i];
  } else if (sub == 1) {
    if (accept) {
      double Tfw[16];
      matmul<4, 4, 4>(st.Tfw, DTinv, Tfw);
      for (int i = 0; i < 16; i++) { r.DT[i] = DTinv[i]; r.Tfw[i] = Tfw[i]; }
      r.err_norm = w.err;
    } else {
      for (int i = 0; i < 16; i++) { r.DT[i] = (i % 5 == 0) ? 1.0 : 0.0; r.Tfw[i] = st.Tfw[i]; }
      r.err_norm = -1.0;
    }
    r.status = status;
    r.n_inliers = w.n_inliers; r.n_inliers_pt = w.n_inliers_pt; r.n_inliers_ls = w.n_inliers_ls;
    r.iters_first = w.iters_first; r.iters_ref = w.iters_ref;
  } else {
    const int r0 = (sub - 2) * 3, r1 = r0 + 3;
    if (accept) unccomp_se3_rows(st.Tfw, st.Tfw_cov, DT_cov, r0, r1, r.Tfw_cov);
    else
      for (int i = r0 * 6; i < r1 * 6; i++) r.Tfw_cov[i] = st.Tfw_cov[i];
  }
}

size_t pose_raw_bytes(int B) { return (size_t)(B > 0 ? B : 1) * sizeof(PoseRaw); }

static int launch_pose_optimize_on(gfpl_ctx* ctx, DevTrackObs obs, const gfpl_frame_state* state, const double* DT_ini,
                                   int overload, gfpl_pose_result* out, PoseRaw* raw) {
  constexpr int G = POSE_GROUP, PAIRS_PER_CTA = POSE_WARPS * 32 / G;
  const size_t smem = (size_t)PAIRS_PER_CTA * pose_smem_doubles(obs.cap_pt, obs.cap_ln) * sizeof(double);
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(pose_optimize_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // prefer shared memory over L1: with the default carve-out the scratch allowed 5 CTAs per SM, the registers allow 6
  static bool carveout_set = false;
  if (!carveout_set) {
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(pose_optimize_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, POSE_CARVEOUT));
    carveout_set = true;
  }
  const int grid = (obs.B + PAIRS_PER_CTA - 1) / PAIRS_PER_CTA;
  pose_optimize_kernel<G><<<grid, POSE_WARPS * 32, smem, ctx->stream>>>(obs, state, DT_ini, overload, raw, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  pose_finalize_kernel<<<(obs.B * FINALIZE_LANES + FINALIZE_THREADS - 1) / FINALIZE_THREADS, FINALIZE_THREADS, 0, ctx->stream>>>(obs.B, raw, state, overload, out, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

// raw_scratch: pose_raw_bytes(obs.B) bytes of device memory owned by the caller (the resident replay), or NULL: the
// hand-off buffer then comes from the stream-ordered pool and is returned to it on every exit path.
int launch_pose_optimize(gfpl_ctx* ctx, DevTrackObs obs, const gfpl_frame_state* state, const double* DT_ini,
                         int overload, gfpl_pose_result* out, void* raw_scratch) {
  if (obs.B <= 0) return GFPL_OK;
  if (raw_scratch) return launch_pose_optimize_on(ctx, obs, state, DT_ini, overload, out, (PoseRaw*)raw_scratch);
  PoseRaw* raw = nullptr;
  GFPL_CUDA_OK(ctx, cudaMallocAsync((void**)&raw, pose_raw_bytes(obs.B), ctx->stream));
  const int rc = launch_pose_optimize_on(ctx, obs, state, DT_ini, overload, out, raw);
  const cudaError_t fe = cudaFreeAsync(raw, ctx->stream);   // also on the error paths above
  if (rc != GFPL_OK) return rc;
  GFPL_CUDA_OK(ctx, fe);
  return GFPL_OK;
}

// ---- vector_stdv_mad alone (auxiliar.cpp:521-537): B arrays, one lane group each — the selection of removeOutliers --------
__global__ void __launch_bounds__(POSE_WARPS * 32)
stdv_mad_kernel(const double* __restrict__ res_all, const int32_t* __restrict__ n_all, int B, int cap, double* __restrict__ out) {
  extern __shared__ __align__(16) double s_mad[];   // per group: cap residuals (even count), then cap + 4 words of scratch
  constexpr int G = POSE_GROUP, PER_CTA = POSE_WARPS * 32 / G;
  const int slot = threadIdx.x / G, lg = threadIdx.x & (G - 1);
  const int p = blockIdx.x * PER_CTA + slot;
  const bool exists = p < B;                   // a group past the batch only takes part in the warp-wide operations
  const int n = exists ? n_all[p] : 0;
  const int cap2 = (cap + 1) & ~1;
  const int per = (cap2 + (cap + 4 + 1) / 2 + 1 + 1) & ~1;   // doubles per group, even: 16-byte aligned groups
  double* res = s_mad + (size_t)slot * per;
  uint32_t* scratch = reinterpret_cast<uint32_t*>(res + cap2);
  for (int k = lg; k < n; k += G) res[k] = res_all[(size_t)p * cap + k];
  __syncwarp();
  const double v = group_stdv_mad<G>(res, n, scratch);
  if (exists && lg == 0) out[p] = v;
}

int launch_stdv_mad(gfpl_ctx* ctx, const double* res, const int32_t* n, int B, int cap, double* out) {
  if (B <= 0) return GFPL_OK;
  constexpr int PER_CTA = POSE_WARPS * 32 / POSE_GROUP;
  const int cap2 = (cap + 1) & ~1;
  size_t per = (size_t)(cap2 + (cap + 4 + 1) / 2 + 1);
  per = (per + 1) & ~(size_t)1;                // 16-byte aligned groups
  const size_t smem = PER_CTA * per * sizeof(double);
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(stdv_mad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stdv_mad_kernel<<<(B + PER_CTA - 1) / PER_CTA, POSE_WARPS * 32, smem, ctx->stream>>>(res, n, B, cap, out);
  GFPL_LAUNCH_OK(ctx);
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
  o.plist = nullptr; o.nplist = 0;
  __shared__ __align__(16) double s_DT[POSE_WARPS][12];
  Acc a;
  pose_store(s_DT[warp], DT_all + (size_t)p * 16);
  PoseConst pc;
  pc.hth = cfg.homog_th; pc.inv_hth = 1.0 / cfg.homog_th; pc.fx_over_hth = cam.fx / cfg.homog_th;
  group_optimize_functions<32, false>(s_DT[warp], o, true, cam, pc, a);
  const double* Hp = a.v;
  const double* g = a.v + 21;
  const double e = a.v[27];
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < 6; j++) Hout[(size_t)p * 36 + i * 6 + j] = Hp[i >= j ? tri(i, j) : tri(j, i)];
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
