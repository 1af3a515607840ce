// k_linecut.cu — K5: good-line-cutting, the per-line cut-ratio search that maximises the log-det (or the
// smallest eigenvalue) of the pose information matrix.
//
// Replaces StereoFrameHandler::estimateProjUncertainty_submodular (reference stereoFrameHandler.cpp:1618-1764),
// getPoseInfoOnLine (:1342-1411), getPoseInfoPoint (:1414-1447), updateEndPointByRatio (:1451-1470),
// logdet (linespec.h:43-56) and getMinEigenValue (linespec.cpp:53-66).
//
// The search is sequential over the lines of one frame pair (every line sees the information of all the
// others through invCov_sum), so the parallelism is 8 neighbour candidates x B pairs: a group of 8 lanes
// owns one pair, lane j evaluates neighbour j (information matrix + 6x6 Cholesky + 6 logs, all in
// registers), and the reference's "first strictly better in j order" rule is replayed with width-8
// shuffles.  Four pairs share a warp.  invCov_sum is accumulated in the reference's order (lines, then
// points, sequentially), and -fmad=false keeps every candidate metric bit-comparable with an IEEE CPU
// evaluation, so the discrete cut decisions are reproducible.  A determinant-lemma evaluation screens the rounds
// (see `Screening path` below); ambiguous rounds are re-decided with the exact arithmetic.  FP64-ALU / latency bound.
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

#ifndef CUT_THREADS_N
#define CUT_THREADS_N 64
#endif
constexpr int CUT_THREADS = CUT_THREADS_N;    // 64: 8 pairs per CTA, 8192 pairs = 1024 CTAs = one wave of 8 CTAs / SM
constexpr int CUT_GROUP = 8;

// StereoFrame::getJacob3D_2D rows 0-1 (stereoFrame.cpp:1394-1413)
__device__ __forceinline__ void jacob3d_2d(const gfpl_camera& cam, double px, double py, double pz, double* J23) {
  const double f = cam.fx, pz_2 = pz * pz;
  J23[0] = f / pz; J23[1] = 0.0; J23[2] = -f * px / pz_2;
  J23[3] = 0.0; J23[4] = f / pz; J23[5] = -f * py / pz_2;
}

// J_loss^T * (Jp * R * cov * R^T * Jp^T) * J_loss   (stereoFrameHandler.cpp:1359-1362)
__device__ __forceinline__ double residual_var(const double* Jp23, const double* R, const double* cov,
                                               const double* Jl) {
  double A[6], Bm[6], Rt[9], Cm[6], Jpt[6], D[4];
  matmul<2, 3, 3>(Jp23, R, A);
  matmul<2, 3, 3>(A, cov, Bm);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Rt[j * 3 + i] = R[i * 3 + j];
  matmul<2, 3, 3>(Bm, Rt, Cm);
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Jpt[j * 2 + i] = Jp23[i * 3 + j];
  matmul<2, 3, 2>(Cm, Jpt, D);
  const double r0 = Jl[0] * D[0] + Jl[1] * D[2];
  const double r1 = Jl[0] * D[1] + Jl[1] * D[3];
  return r0 * Jl[0] + r1 * Jl[1];
}

struct LineInfo {  // what getPoseInfoOnLine needs to form info = [Js Je] * inv(diag(cov_sR, cov_eR)) * [Js Je]^T
  double Js[6], Je[6];
  double i00, i01, i10, i11;
};

// line data layout used by the evaluation: sP[3] eP[3] le[2] covS[9] covE[9] = 26 doubles
constexpr int LD_SP = 0, LD_EP = 3, LD_LE = 6, LD_CS = 8, LD_CE = 17, LD_SIZE = 26;

// One endpoint of getPoseInfoOnLine (:1347-1400): A / SA = the endpoint (and its covariance) that moves towards B / SB by
// the cut ratio c.  Writes the Jacobian column J[6] and returns the residual variance.  `DT` = [R (9) | t (3)].
__device__ __forceinline__ double pose_info_endpoint(const gfpl_camera& cam, const gfpl_config& cfg, const double* DT,
                                                     const double* A, const double* Bp, const double* SA, const double* SB,
                                                     const double* le, double c, double* J) {
  double R[9], T4[16];
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = DT[i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    T4[i * 4 + 0] = R[i * 3 + 0]; T4[i * 4 + 1] = R[i * 3 + 1]; T4[i * 4 + 2] = R[i * 3 + 2]; T4[i * 4 + 3] = DT[9 + i];
  }
  double P[3], cur[3], cv[9], Jp[6];
  const double a0 = (1 - c) * (1 - c), b0 = c * c;
#pragma unroll
  for (int k = 0; k < 3; k++) P[k] = (1 - c) * A[k] + c * Bp[k];
#pragma unroll
  for (int k = 0; k < 9; k++) cv[k] = a0 * SA[k] + b0 * SB[k];
  se3_apply(T4, P, cur);
  jacob3d_2d(cam, cur[0], cur[1], cur[2], Jp);
  const double var = residual_var(Jp, R, cv, le);
  jac6(cam.fx / fmax(cfg.homog_th, cur[2] * cur[2]), cur[0], cur[1], cur[2], le[0], le[1], J);
  return var;
}

// the 2x2 inverse of diag(cov_sR, cov_eR) as adjugate * (1/det), as the reference's Matrix2d::inverse() forms it
__device__ __forceinline__ void pose_info_weights(double cov_sR, double cov_eR, LineInfo& out) {
  const double det = cov_sR * cov_eR - 0.0 * 0.0;
  const double invdet = 1.0 / det;
  out.i00 = cov_eR * invdet;
  out.i11 = cov_sR * invdet;
  out.i01 = -0.0 * invdet;
  out.i10 = -0.0 * invdet;
}

// getPoseInfoOnLine up to the two Jacobian columns and the 2x2 inverse.  `L` = line data; DT and L may live in shared
// memory (pass 3: one copy per 8-lane group) or be read straight from global (pass 1).
__device__ __forceinline__ void pose_info_on_line(const gfpl_camera& cam, const gfpl_config& cfg, const double* DT,
                                                  const double* L, double c0, double c1, LineInfo& out) {
  const double cov_sR = pose_info_endpoint(cam, cfg, DT, &L[LD_SP], &L[LD_EP], &L[LD_CS], &L[LD_CE], &L[LD_LE], c0, out.Js);
  const double cov_eR = pose_info_endpoint(cam, cfg, DT, &L[LD_EP], &L[LD_SP], &L[LD_CE], &L[LD_CS], &L[LD_LE], c1, out.Je);
  pose_info_weights(cov_sR, cov_eR, out);
}

__device__ __forceinline__ double info_entry(const LineInfo& I, int i, int j) {
  const double Ai0 = I.Js[i] * I.i00 + I.Je[i] * I.i10;
  const double Ai1 = I.Js[i] * I.i01 + I.Je[i] * I.i11;
  return Ai0 * I.Js[j] + Ai1 * I.Je[j];
}

// metric on a lower-triangle-packed matrix
__device__ __forceinline__ double cut_metric(const gfpl_config& cfg, const double* M21) {
  if (cfg.max_vol_line_cut) return logdet6_packed(M21);
  double S[36], ev[6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) { S[i * 6 + j] = M21[tri(i, j)]; S[j * 6 + i] = M21[tri(i, j)]; }
  sym_eigenvalues<6>(S, ev);
  return ev[0];
}


// ---------------------------------------------------------------------------------------------------------------
// Screening path.  All candidates of one line share S = invCov_sum - info(line), and a candidate adds the rank-2
// term  a*u*u^T + b*v*v^T  (u = Js(c0), a = 1/cov_sR(c0), v = Je(c1), b = 1/cov_eR(c1)), so by the matrix
// determinant lemma   logdet(S + info(c0,c1)) - logdet(S) = log( (1 + a u'S^-1 u)(1 + b v'S^-1 v) - a b (u'S^-1 v)^2 ).
// That difference ("gain") orders the candidates exactly like the reference's metric does, costs one 6x6 mat-vec per
// half instead of a 6x6 Cholesky + 6 logs per candidate, and the u- and v-halves are shared by the three candidates
// that have the same c0 resp. c1.  It is NOT the reference's arithmetic, so it only SCREENS: a round is decided by
// the gains when every comparison the reference would make is clear by more than CUT_TAU (>= 100x the combined
// rounding error of both evaluations); otherwise the round is re-evaluated with the exact path below, which then
// decides exactly as before.  Outputs (cutRatio, invCovPose, moved endpoints, invCov_sum) always come from the
// exact arithmetic.  FMAs are written explicitly here (the file is compiled with -fmad=false).
// ---------------------------------------------------------------------------------------------------------------
constexpr double CUT_TAU = 1e-8;

// one cached half (8 doubles): vec[6], a = 1 / residual variance, g = u' S^-1 u.  vec is the Jacobian column u for the
// start-point axis and w = S^-1 v for the end-point axis: a candidate's cross term u' S^-1 v is then one 6-element dot
// product of its two halves.  A half depends on (line, axis, cut ratio) only — S^-1 is fixed while a line is searched —
// so halves are cached per cut index (ratio = index * step) in a ring of HC_RING slots per axis and computed four per
// axis at a time by the group's 8 lanes; a round then costs one dot product per lane instead of one half evaluation.
constexpr int HC_VEC = 0, HC_A = 6, HC_G = 7, HC_SIZE = 8, HC_RING = 8, HC_CHUNK = 4;

__device__ __forceinline__ double dot3f(const double* a, const double* b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }

// reciprocal to ~1e-15 relative: float seed + 2 Newton steps (screening path only)
__device__ __forceinline__ double rcp_fast(double x) {
  double r = (double)__frcp_rn((float)x);
  r = r * fma(-x, r, 2.0);
  r = r * fma(-x, r, 2.0);
  return r;
}

// one endpoint of getPoseInfoOnLine at cut ratio c: A/SA = the endpoint that shrinks towards B/SB.  Result -> out[HC_SIZE]
__device__ __forceinline__ void half_fast(const gfpl_camera& cam, const gfpl_config& cfg, const double* DT, const double* A,
                                          const double* Bp, const double* SA, const double* SB, const double* le,
                                          const double* Sinv /*6x6 row-major*/, double c, bool store_w, double* out) {
  const double omc = 1.0 - c;
  double P[3], cur[3], cv[9];
#pragma unroll
  for (int k = 0; k < 3; k++) P[k] = fma(c, Bp[k], omc * A[k]);
  const double a2 = omc * omc, b2 = c * c;
#pragma unroll
  for (int k = 0; k < 9; k++) cv[k] = fma(b2, SB[k], a2 * SA[k]);
#pragma unroll
  for (int i = 0; i < 3; i++) cur[i] = fma(DT[i * 3 + 2], P[2], fma(DT[i * 3 + 1], P[1], fma(DT[i * 3], P[0], DT[9 + i])));
  // q = R^T Jp^T l  (3-vector), variance = q' cov q
  const double f = cam.fx, iz = rcp_fast(cur[2]);
  const double fiz = f * iz, lx = le[0], ly = le[1];
  const double j0 = fiz * lx, j1 = fiz * ly, j2 = -(fiz * iz) * fma(cur[1], ly, cur[0] * lx);
  double q[3], t[3];
#pragma unroll
  for (int k = 0; k < 3; k++) q[k] = fma(DT[6 + k], j2, fma(DT[3 + k], j1, DT[k] * j0));
#pragma unroll
  for (int i = 0; i < 3; i++) t[i] = dot3f(&cv[i * 3], q);
  const double a = rcp_fast(dot3f(q, t));
  // jac6 with k = fx / max(homog_th, Z^2)
  const double gx = cur[0], gy = cur[1], gz = cur[2], gz2 = gz * gz;
  const double kf = gz2 > cfg.homog_th ? fiz * iz : f / cfg.homog_th;
  double u[6];
  const double xl = gx * lx, yl = gy * ly;
  u[0] = kf * lx * gz;
  u[1] = kf * ly * gz;
  u[2] = -kf * (xl + yl);
  u[3] = -kf * fma(gy, xl + yl, gz2 * ly);
  u[4] = kf * fma(gx, xl + yl, gz2 * lx);
  u[5] = kf * gz * fma(gx, ly, -(gy * lx));
  double g = 0.0, w[6];
  const double2* S2 = reinterpret_cast<const double2*>(Sinv);  // rows of 3 x 128 bit
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double2 r0 = S2[i * 3], r1 = S2[i * 3 + 1], r2 = S2[i * 3 + 2];
    double sacc = fma(r0.x, u[0], 0.0);
    sacc = fma(r0.y, u[1], sacc);
    sacc = fma(r1.x, u[2], sacc);
    sacc = fma(r1.y, u[3], sacc);
    sacc = fma(r2.x, u[4], sacc);
    sacc = fma(r2.y, u[5], sacc);
    w[i] = sacc;
    g = fma(u[i], sacc, g);
  }
  double2* o2 = reinterpret_cast<double2*>(out);  // [vec0..vec5 | a g]
#pragma unroll
  for (int i = 0; i < 3; i++)
    o2[HC_VEC / 2 + i] = store_w ? make_double2(w[2 * i], w[2 * i + 1]) : make_double2(u[2 * i], u[2 * i + 1]);
  o2[HC_A / 2] = make_double2(a, g);
}

// per-group (one frame pair) state kept in shared memory: every lane of the group reads the same addresses
// (broadcast), which frees ~120 registers per thread and lets 16 warps / SM hide the fp64 dependency chains.
struct __align__(16) GroupState {
  double Sinv[36];          // inverse of (invCov_sum - info(line)) for the screening path, row-major (16-byte aligned rows)
  double hc[2][HC_RING][HC_SIZE];  // cached halves: [axis][cut index & (HC_RING - 1)] (64-byte rows)
  double DT[12];            // R | t of DT_inv
  double L[LD_SIZE];        // current line
  double J[CUT_GROUP][6];   // point Jacobians of the current chunk of the invCov_sum pass
  double sum[21];           // invCov_sum, lower triangle
  double sum0[21];          // invCov_sum before the current line's information is removed (exact metric of cut {0,0})
};
static_assert(sizeof(GroupState) % 16 == 0 && HC_SIZE % 2 == 0 && HC_CHUNK * 2 == CUT_GROUP, "128-bit shared-memory accesses below");

#ifndef CUT_BALANCE
#define CUT_BALANCE 1     // pairs served in balanced order (descending line count, snake over the warps)
#endif
#ifndef CUT_MAX_REGS
#define CUT_MAX_REGS 128   // 8 CTAs (16 warps) per SM
#endif

// once per line, out of line: every lane factorises S (Cholesky, redundantly), lane j < 6 of the group then solves for
// column j of S^-1 only and publishes it
__device__ __noinline__ bool spd_inverse6_publish(GroupState* G, int j, unsigned gmask) {
  double L[21];
#pragma unroll
  for (int i = 0; i < 21; i++) L[i] = G->sum[i];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double x = L[tri(k, k)];
#pragma unroll
    for (int jx = 0; jx < k; jx++) x = fma(-L[tri(k, jx)], L[tri(k, jx)], x);
    ok = ok && (x > 0.0);
    const double r = rsqrt(x);
    L[tri(k, k)] = r;  // store 1/L_kk
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      double sv = L[tri(i, k)];
#pragma unroll
      for (int jx = 0; jx < k; jx++) sv = fma(-L[tri(i, jx)], L[tri(k, jx)], sv);
      L[tri(i, k)] = sv * r;
    }
  }
  const int c = j < 6 ? j : 5;  // lanes 6, 7 repeat column 5 (not stored)
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {  // L y = e_c
    double sv = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; k++) sv = fma(-L[tri(i, k)], y[k], sv);
    y[i] = (i < c) ? 0.0 : sv * L[tri(i, i)];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {  // L^T x = y
    double sv = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) sv = fma(-L[tri(k, i)], y[k], sv);
    y[i] = sv * L[tri(i, i)];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    ok = ok && (fabs(y[i]) < 1e300);
    if (j < 6) G->Sinv[i * 6 + c] = y[i];
  }
  return (__ballot_sync(gmask, ok) & gmask) == gmask;
}

// HC_CHUNK halves per axis from index b0 (start-point axis, lanes 0..3) / b1 (end-point axis, lanes 4..7) into the
// group's cache.  Out of line: it runs once per ~4 rounds and its ~80 live doubles stay out of the round loop.
__device__ __noinline__ void fill_halves(const gfpl_camera* cam, const gfpl_config* cfg, GroupState* G, int j, int b0, int b1,
                                         bool go0, bool go1, double step) {
  const bool isS = j < HC_CHUNK;
  const int k = (isS ? b0 : b1) + (j & (HC_CHUNK - 1));
  if (isS ? go0 : go1)
    half_fast(*cam, *cfg, G->DT, isS ? &G->L[LD_SP] : &G->L[LD_EP], isS ? &G->L[LD_EP] : &G->L[LD_SP],
              isS ? &G->L[LD_CS] : &G->L[LD_CE], isS ? &G->L[LD_CE] : &G->L[LD_CS], &G->L[LD_LE], G->Sinv,
              (double)k * step, !isS, G->hc[isS ? 0 : 1][k & (HC_RING - 1)]);
}

// covSpt3D / covEpt3D of one previous-frame stereo line, out of line
__device__ __noinline__ void line_cov_store(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_line_feature* l, gfpl_line_cov* out) {
  gfpl_line_cov c;
  line_cov_from_feature(*cam, *cfg, *l, c);
  *out = c;
}

// initial information of one line at cut {0,0} (:1638-1649), out of line
__device__ __noinline__ void exact_info_initial(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT,
                                                const gfpl_ln_obs* o, const gfpl_line_cov* c, gfpl_line_cut* C) {
  double Lr[LD_SIZE];
  LineInfo I;
#pragma unroll
  for (int k = 0; k < 3; k++) { Lr[LD_SP + k] = o->sP[k]; Lr[LD_EP + k] = o->eP[k]; }
  Lr[LD_LE] = o->le_obs[0]; Lr[LD_LE + 1] = o->le_obs[1];
#pragma unroll
  for (int k = 0; k < 9; k++) { Lr[LD_CS + k] = c->covS[k]; Lr[LD_CE + k] = c->covE[k]; }
  pose_info_on_line(*cam, *cfg, DT, Lr, 0.0, 0.0, I);
  C->cutRatio[0] = 0.0; C->cutRatio[1] = 0.0;
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) C->invCovPose[a * 6 + b] = info_entry(I, a, b);
}

// ---- exact (reference-arithmetic) evaluations, out of line -----------------------------------------------------------
// They run only in the rare rounds the screening cannot decide and once per cut line; keeping their ~100 live doubles
// out of the round loop is what lets the loop run from registers.  cam / cfg are the CTA's shared-memory copies.
__device__ __noinline__ double exact_metric_at(const gfpl_camera* cam, const gfpl_config* cfg, const GroupState* G,
                                               double c0, double c1) {
  LineInfo I;
  pose_info_on_line(*cam, *cfg, G->DT, G->L, c0, c1, I);
  double M[21];
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) M[tri(a, b)] = info_entry(I, a, b) + G->sum[tri(a, b)];
  return cut_metric(*cfg, M);
}
__device__ __noinline__ double exact_metric_start(const gfpl_config* cfg, const GroupState* G) { return cut_metric(*cfg, G->sum0); }
// Information of the accepted cut (once per cut line).  The two endpoints are independent until the 6x6 product: lanes
// 0..3 evaluate the start point, lanes 4..7 the end point (same operations as pose_info_on_line, so the same bits), the
// halves meet in the group's half cache (free at the end of a line's search), then every lane forms its entries.
__device__ __noinline__ void exact_info_store(const gfpl_camera* cam, const gfpl_config* cfg, GroupState* G, double c0,
                                              double c1, int j, unsigned gmask, double* invCovPose) {
  const bool isS = j < CUT_GROUP / 2;
  double J[6];
  const double var = pose_info_endpoint(*cam, *cfg, G->DT, isS ? &G->L[LD_SP] : &G->L[LD_EP], isS ? &G->L[LD_EP] : &G->L[LD_SP],
                                        isS ? &G->L[LD_CS] : &G->L[LD_CE], isS ? &G->L[LD_CE] : &G->L[LD_CS], &G->L[LD_LE],
                                        isS ? c0 : c1, J);
  double* xs = &G->hc[0][0][0];   // [Js (6) cov_sR . | Je (6) cov_eR .]
  if ((j & (CUT_GROUP / 2 - 1)) == 0) {
    double* o = xs + (isS ? 0 : 8);
#pragma unroll
    for (int i = 0; i < 6; i++) o[i] = J[i];
    o[6] = var;
  }
  __syncwarp(gmask);
  LineInfo I;
#pragma unroll
  for (int i = 0; i < 6; i++) { I.Js[i] = xs[i]; I.Je[i] = xs[8 + i]; }
  pose_info_weights(xs[6], xs[14], I);
  for (int e = j; e < 36; e += CUT_GROUP) invCovPose[e] = info_entry(I, e / 6, e % 6);
}

// The greedy search of one line (:1662-1763), out of line: its loop-carried state then has the register file to itself
// (inlined in the kernel, part of it lived in local memory).  Returns the accepted cut {c0, c1}.
__device__ __noinline__ double2 search_line(const gfpl_camera* cam, const gfpl_config* cfg, GroupState* G, int j, unsigned gmask,
                                            bool fast_ok) {
  const double step = cfg->cut_step, rng0 = cfg->cut_rng[0], rng1 = cfg->cut_rng[1];
  const int d0 = (j == 0 || j == 4 || j == 5) ? 1 : ((j == 1 || j == 6 || j == 7) ? -1 : 0);   // neighbour j, in steps
  const int d1 = (j == 2 || j == 4 || j == 6) ? 1 : ((j == 3 || j == 5 || j == 7) ? -1 : 0);
  const double nb0 = d0 > 0 ? step : (d0 < 0 ? -step : 0.0), nb1 = d1 > 0 ? step : (d1 < 0 ? -step : 0.0);
  double cur0 = 0.0, cur1 = 0.0;
  int i0 = 0, i1 = 0;            // the same cut as integer multiples of the step (indices of the cached halves)
  int lo0 = -1, hi0 = -1, lo1 = -1, hi1 = -1;   // cached index ranges [lo, hi) per axis, hi - lo <= HC_RING
  double metric_last = 0.0;      // the reference's metric of the current cut; evaluated lazily (exact rounds only)
  bool have_exact_last = false;
  bool at_start = true;          // current cut is still {0,0}: its exact metric is logdet(sum0)
  double dcur = 0.0;             // screening determinant factor of the current cut
  bool have_dcur = false;
  // The search alternates between a call-free loop of screened rounds (everything it carries stays in registers) and
  // the rare out-of-line work it stops for: a refill of the half cache or a round in the reference's arithmetic.
  enum { STOP, REFILL, EXACT };
  for (;;) {
    int why = STOP;
    if (!fast_ok) why = (cur0 + cur1 <= 1.0) ? EXACT : STOP;
    else
      while (cur0 + cur1 <= 1.0) {
        // a round reads the halves at index - 1, index, index + 1 of both axes
        if (i0 - 1 < lo0 || i0 + 1 >= hi0 || i1 - 1 < lo1 || i1 + 1 >= hi1) { why = REFILL; break; }
        const double ct0 = cur0 + nb0, ct1 = cur1 + nb1;
        bool valid = !(ct0 + ct1 > 1.0);
        valid = valid && !(ct0 < rng0 || ct0 > rng1);
        valid = valid && !(ct1 < rng0 || ct1 > rng1);
        // determinant factor of (si, ei): (1 + a u'S^-1u)(1 + b v'S^-1v) - a b (u'S^-1v)^2  (monotone in the metric)
        auto factor = [&](int si, int ei) {
          const double2* hs = reinterpret_cast<const double2*>(G->hc[0][si & (HC_RING - 1)]);
          const double2* he = reinterpret_cast<const double2*>(G->hc[1][ei & (HC_RING - 1)]);
          double guv = 0.0;
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const double2 uu = hs[HC_VEC / 2 + i], ww = he[HC_VEC / 2 + i];
            guv = fma(uu.x, ww.x, guv);
            guv = fma(uu.y, ww.y, guv);
          }
          const double2 ags = hs[HC_A / 2], age = he[HC_A / 2];   // (a, g) of each half
          const double a = ags.x, b = age.x;
          return fma(-(a * b) * guv, guv, fma(a, ags.y, 1.0) * fma(b, age.y, 1.0));
        };
        const double dj = factor(i0 + d0, i1 + d1);                  // this lane's candidate
        if (!have_dcur) { dcur = factor(i0, i1); have_dcur = true; }   // the current cut (same value in every lane)
        // The reference keeps the first candidate (in j order) that is strictly better than everything before it,
        // i.e. the arg-max with ties to the lowest j, or none if no candidate beats the current cut (:1722-1730).
        // The factors decide that when the largest one is clear of every other candidate and of the current cut by
        // more than CUT_TAU (relative), or when the current cut is clear of all candidates.  Usable factors are
        // positive, so their order is the order of their bit patterns: the maximum is two 32-bit butterflies.
        const bool usable = dj > 0.0 && dj < 1e300;
        const unsigned hj = valid ? (unsigned)__double2hiint(dj) : 0u, lj = valid ? (unsigned)__double2loint(dj) : 0u;
        unsigned mh = hj;
#pragma unroll
        for (int dlt = 1; dlt < CUT_GROUP; dlt <<= 1) mh = max(mh, __shfl_xor_sync(gmask, mh, dlt, CUT_GROUP));
        unsigned ml = (hj == mh) ? lj : 0u;
#pragma unroll
        for (int dlt = 1; dlt < CUT_GROUP; dlt <<= 1) ml = max(ml, __shfl_xor_sync(gmask, ml, dlt, CUT_GROUP));
        const double best = __hiloint2double((int)mh, (int)ml);   // largest candidate factor (0: no valid candidate)
        // flags of this lane: bit j = "I am the maximum", bit 8 = a factor that cannot be used, bit 9 = too close to it
        const bool is_max = valid && hj == mh && lj == ml;
        unsigned fl = is_max ? (1u << j) : 0u;
        if (valid && !usable) fl |= 1u << 8;
        if (valid && !is_max && !(best - dj > CUT_TAU * best)) fl |= 1u << 9;
#pragma unroll
        for (int dlt = 1; dlt < CUT_GROUP; dlt <<= 1) fl |= __shfl_xor_sync(gmask, fl, dlt, CUT_GROUP);
        const bool sane = dcur > 0.0 && dcur < 1e300 && !(fl & (1u << 8)) && !(best >= 1e300);
        const unsigned winners = fl & 0xFFu;
        int cand;
        if (sane && best - dcur > CUT_TAU * best && !(fl & (1u << 9)) && (winners & (winners - 1u)) == 0u)
          cand = __ffs(winners) - 1;                         // one candidate is clearly the best
        else if (sane && dcur - best > CUT_TAU * dcur) break;  // every candidate is clearly worse: the search stops (STOP)
        else { why = EXACT; break; }
        have_exact_last = false; dcur = best;
        // the winner's cut: the same additions lane `cand` made (cur + nb), so every lane can make them itself
        const int c0d = (int)((0xA52u >> (2 * cand)) & 3u) - 1, c1d = (int)((0x2225u >> (2 * cand)) & 3u) - 1;
        cur0 = cur0 + (c0d > 0 ? step : (c0d < 0 ? -step : 0.0));
        cur1 = cur1 + (c1d > 0 ? step : (c1d < 0 ? -step : 0.0));
        i0 += c0d; i1 += c1d;
        at_start = false;
      }
    if (why == STOP) break;
    if (why == REFILL) {
      // per axis: append HC_CHUNK indices at hi (if that does not evict index - 1), or restart the range at index - 1
      int b0 = hi0, b1 = hi1;
      bool go0 = true, go1 = true;
      if (i0 - 1 < lo0 || i0 - 1 > hi0) { lo0 = i0 - 1; b0 = lo0; }
      else if (hi0 + HC_CHUNK - HC_RING > i0 - 1) go0 = false;
      if (i1 - 1 < lo1 || i1 - 1 > hi1) { lo1 = i1 - 1; b1 = lo1; }
      else if (hi1 + HC_CHUNK - HC_RING > i1 - 1) go1 = false;
      fill_halves(cam, cfg, G, j, b0, b1, go0, go1, step);
      if (go0) { hi0 = b0 + HC_CHUNK; lo0 = max(lo0, hi0 - HC_RING); }
      if (go1) { hi1 = b1 + HC_CHUNK; lo1 = max(lo1, hi1 - HC_RING); }
      __syncwarp(gmask);
      continue;
    }
    // exact round: the reference's arithmetic (also the only path for the min-eigenvalue metric)
    const double ct0 = cur0 + nb0, ct1 = cur1 + nb1;
    bool valid = !(ct0 + ct1 > 1.0);
    valid = valid && !(ct0 < rng0 || ct0 > rng1);
    valid = valid && !(ct1 < rng0 || ct1 > rng1);
    if (!have_exact_last) {  // metric of the current cut, needed as the running best
      metric_last = at_start ? exact_metric_start(cfg, G) : exact_metric_at(cam, cfg, G, cur0, cur1);
      have_exact_last = true;
    }
    const double metric = valid ? exact_metric_at(cam, cfg, G, ct0, ct1) : 0.0;
    // "first strictly better in j order" (:1722-1730)
    int cand = -1;
    double best = metric_last;
#pragma unroll
    for (int jj = 0; jj < CUT_GROUP; jj++) {
      const double mj = shfl_f64(gmask, metric, jj, CUT_GROUP);
      const int vj = __shfl_sync(gmask, (int)valid, jj, CUT_GROUP);
      if (vj && mj > best) { best = mj; cand = jj; }
    }
    metric_last = best;
    have_dcur = false;   // the screening factor of the new current cut is re-evaluated next round
    if (cand < 0) break;  // early termination (:1741-1743)
    const int c0d = (int)((0xA52u >> (2 * cand)) & 3u) - 1, c1d = (int)((0x2225u >> (2 * cand)) & 3u) - 1;
    cur0 = cur0 + (c0d > 0 ? step : (c0d < 0 ? -step : 0.0));
    cur1 = cur1 + (c1d > 0 ? step : (c1d < 0 ? -step : 0.0));
    i0 += c0d; i1 += c1d;
    at_start = false;
  }
  return make_double2(cur0, cur1);
}

// The search is one dependent chain per pair, so every global round trip on it is exposed (ncu: two thirds of the stall
// samples of the kernel body were long-scoreboard).  What the next step will read is pulled into L1 ahead of time with
// prefetch.global.L1 — no registers, no dependence.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__global__ void __maxnreg__(CUT_MAX_REGS)
linecut_kernel(const double* __restrict__ DT_inv_all, const gfpl_frame_state* __restrict__ state, DevTrackObs obs,
               gfpl_line_cov* prev_cov, int cov_cap, const gfpl_line_feature* __restrict__ prev_ln, int prev_cap_ln,
               gfpl_line_cut* __restrict__ cut_all, const int32_t* __restrict__ perm, gfpl_camera cam, gfpl_config cfg) {
  __shared__ GroupState s_grp[CUT_THREADS / CUT_GROUP];
  __shared__ gfpl_camera s_cam;
  __shared__ gfpl_config s_cfg;
  if (threadIdx.x == 0) { s_cam = cam; s_cfg = cfg; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int j = lane & (CUT_GROUP - 1);             // neighbour index of this lane
  const int grp = lane / CUT_GROUP;                 // pair slot inside the warp
  const unsigned gmask = 0xFFu << (grp * CUT_GROUP);
  // Which pair this group serves.  The kernel is one wave and ends with its slowest warp; the pairs' chains differ (9..36
  // matched lines around a mean of 20 on the EuRoC-shaped data).  `perm` lists the pairs by descending number of matched
  // lines (linecut_order_kernel), and warp w takes ranks w, 2W-1-w, 2W+w, 4W-1-w (W warps): long and short chains share a
  // warp, the sums are balanced.  perm == nullptr: natural order.
  int p;
  if (perm) {
    const int w = blockIdx.x * (CUT_THREADS / 32) + (threadIdx.x >> 5);
    const int W = gridDim.x * (CUT_THREADS / 32);     // every warp of the grid takes part (the grid fills the free warp slots
    const int rank = grp * W + ((grp & 1) ? W - 1 - w : w);   // of one wave: the warps of the longest pairs then serve three pairs)
    if (rank >= obs.B) return;
    p = perm[rank];
  } else {
    p = (blockIdx.x * (CUT_THREADS / 32) + (threadIdx.x >> 5)) * (32 / CUT_GROUP) + grp;
    if (p >= obs.B) return;
  }
  const int nln = obs.n_ln[p], npt = obs.n_pt[p];
  if (nln == 0) return;  // :1621
  GroupState& G = s_grp[threadIdx.x / CUT_GROUP];
  gfpl_ln_obs* lns = obs.ln + (size_t)p * obs.cap_ln;
  const gfpl_pt_obs* pts = obs.pt + (size_t)p * obs.cap_pt;
  gfpl_line_cov* cov = prev_cov + (size_t)p * cov_cap;
  gfpl_line_cut* cut = cut_all + (size_t)p * obs.cap_ln;

  if (j == 0) {
    double DT_inv[16];
    if (DT_inv_all) {
#pragma unroll
      for (int i = 0; i < 16; i++) DT_inv[i] = DT_inv_all[(size_t)p * 16 + i];
    } else {
      double Tc[16], Tci[16];
      matmul<4, 4, 4>(state[p].Tfw, state[p].DT, Tc);  // predicted curr->Tfw (:157)
      inverse4(Tc, Tci);
      matmul<4, 4, 4>(Tci, state[p].Tfw, DT_inv);      // :1635
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      G.DT[i * 3 + 0] = DT_inv[i * 4 + 0]; G.DT[i * 3 + 1] = DT_inv[i * 4 + 1]; G.DT[i * 3 + 2] = DT_inv[i * 4 + 2];
      G.DT[9 + i] = DT_inv[i * 4 + 3];
    }
  }
  __syncwarp(gmask);

  // pass 1a: initial information of every line (cut = 0), lanes stride over the lines (:1638-1649)
  // prev_ln given (the tracking step): estimateStereoUncertainty (stereoFrame.cpp:1448-1484) for the MATCHED lines only,
  // evaluated here instead of by a kernel of its own over every stereo line of the previous frame; the records land in
  // prev_cov, where pass 3 reads them
  for (int l = j; l < nln; l += CUT_GROUP) {
    const int pi = lns[l].prev_idx;
    if (prev_ln) line_cov_store(&s_cam, &s_cfg, &prev_ln[(size_t)p * prev_cap_ln + pi], &cov[pi]);
    exact_info_initial(&s_cam, &s_cfg, G.DT, &lns[l], &cov[pi], &cut[l]);
  }
  __syncwarp(gmask);

  // pass 1b/2: invCov_sum in the reference's order — lines, then points, each element summed sequentially.
  // Lane j owns lower-triangle elements e = j, j+8, j+16.
  // lane j owns the lower-triangle elements e = j, j + 8, j + 16 of the 6x6 sums: (row, column) of each, once per kernel
  int ea[3], eb[3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const int e = min(j + 8 * q, 20);
    int a = 0;
#pragma unroll
    for (int t = 1; t < 6; t++) a += ((t * (t + 1)) / 2 <= e) ? 1 : 0;
    ea[q] = a; eb[q] = e - a * (a + 1) / 2;
  }
  {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int l = 0; l < nln; l++) {
      const double* info = cut[l].invCovPose;
#pragma unroll
      for (int q = 0; q < 3; q++) acc[q] = acc[q] + info[ea[q] * 6 + eb[q]];
    }
    double T4[16];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      T4[i * 4 + 0] = G.DT[i * 3 + 0]; T4[i * 4 + 1] = G.DT[i * 3 + 1]; T4[i * 4 + 2] = G.DT[i * 3 + 2]; T4[i * 4 + 3] = G.DT[9 + i];
    }
    // getPoseInfoPoint (:1414-1447): the group's 8 lanes evaluate 8 Jacobians at a time (one point each) and publish
    // them; every lane then adds its three products point by point, i.e. in the reference's order
    for (int k0 = 0; k0 < npt; k0 += CUT_GROUP) {
      if (k0 + 2 * CUT_GROUP + j < npt) {   // the point this lane evaluates two rounds from now
        const char* nx = reinterpret_cast<const char*>(&pts[k0 + 2 * CUT_GROUP + j]);
        prefetch_l1(nx); prefetch_l1(nx + sizeof(gfpl_pt_obs) - 8);
      }
      if (k0 + j < npt) {
        const gfpl_pt_obs& po = pts[k0 + j];
        double c[3], uv[2], J[6];
        se3_apply(T4, po.P, c);
        projection(cam, c, uv);
        const double dx = uv[0] - po.pl_obs[0], dy = uv[1] - po.pl_obs[1];
        jac6(cam.fx / fmax(cfg.homog_th, c[2] * c[2]), c[0], c[1], c[2], dx, dy, J);
#pragma unroll
        for (int i = 0; i < 6; i++) G.J[j][i] = J[i];
      }
      __syncwarp(gmask);
      const int nk = min(CUT_GROUP, npt - k0);
      for (int k = 0; k < nk; k++) {
#pragma unroll
        for (int q = 0; q < 3; q++) acc[q] = acc[q] + G.J[k][ea[q]] * G.J[k][eb[q]];
      }
      __syncwarp(gmask);
    }
#pragma unroll
    for (int q = 0; q < 3; q++)
      if (j + 8 * q < 21) G.sum[j + 8 * q] = acc[q];
  }
  __syncwarp(gmask);

  // pass 3: sequential greedy search per line (:1662-1763)
  int prev_idx_next = nln > 1 ? lns[1].prev_idx : 0;   // covariance record of line l + 1 (loaded one line ahead)
  for (int l = 0; l < nln; l++) {
    gfpl_line_cut& C = cut[l];
    int prev_idx_next2 = 0;
    if (l + 1 < nln) {   // line l + 1: observation (88 B), covariances (144 B), initial information (288 B of the cut record)
      const char* o1 = reinterpret_cast<const char*>(&lns[l + 1]);
      const char* c1 = reinterpret_cast<const char*>(&cov[prev_idx_next]);
      const char* i1 = reinterpret_cast<const char*>(&cut[l + 1]);
      if (j < 5) prefetch_l1(c1 + 32 * j); else prefetch_l1(o1 + 32 * (j - 5));
      prefetch_l1(i1 + 32 * j);
      if (j < 3) prefetch_l1(i1 + 256 + 32 * j);
      if (l + 2 < nln) prev_idx_next2 = lns[l + 2].prev_idx;
    }
    {  // stage the line into the group's shared memory
      const gfpl_ln_obs& o = lns[l];
      const gfpl_line_cov& c = cov[o.prev_idx];
      for (int k = j; k < LD_SIZE; k += CUT_GROUP) {
        double v;
        if (k < 3) v = o.sP[k];
        else if (k < 6) v = o.eP[k - 3];
        else if (k < 8) v = o.le_obs[k - 6];
        else if (k < 17) v = c.covS[k - 8];
        else v = c.covE[k - 17];
        G.L[k] = v;
      }
    }
    __syncwarp(gmask);
    {
      const double* info = C.invCovPose;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int e = j + 8 * q;
        if (e < 21) {
          G.sum0[e] = G.sum[e];
          G.sum[e] = G.sum[e] - info[ea[q] * 6 + eb[q]];
        }
      }
    }
    __syncwarp(gmask);
    // screening set-up: S^-1 once per line (every lane computes it redundantly, lanes publish it together)
    bool fast_ok = cfg.max_vol_line_cut != 0;
    if (fast_ok) fast_ok = spd_inverse6_publish(&G, j, gmask);
    __syncwarp(gmask);
    const double2 cr = search_line(&s_cam, &s_cfg, &G, j, gmask, fast_ok);
    const double cur0 = cr.x, cur1 = cr.y;
    // invCovPose of the accepted cut, in the reference's arithmetic (cut {0,0} is already stored by pass 1)
    if (cur0 != 0.0 || cur1 != 0.0) exact_info_store(&s_cam, &s_cfg, &G, cur0, cur1, j, gmask, C.invCovPose);
    __syncwarp(gmask);
    // updateEndPointByRatio (:1451-1470) — lane 0 of the group writes the results
    if (j == 0) {
      C.cutRatio[0] = cur0; C.cutRatio[1] = cur1;
      C.spl[0] = C.spl[1] = C.epl[0] = C.epl[1] = 0.0;
      C.sdisp = 0.0; C.edisp = 0.0;
      if (!(fabs(cur0) < 0.0001 && fabs(cur1) < 0.0001)) {
        double sP[3] = {G.L[LD_SP], G.L[LD_SP + 1], G.L[LD_SP + 2]};
        if (fabs(cur0) > 0.0001) {
#pragma unroll
          for (int k = 0; k < 3; k++) sP[k] = (1 - cur0) * G.L[LD_SP + k] + cur0 * G.L[LD_EP + k];
#pragma unroll
          for (int k = 0; k < 3; k++) lns[l].sP[k] = sP[k];
          projection(cam, sP, C.spl);
          C.sdisp = get_disparity(cam, sP[2]);
        }
        if (fabs(cur1) > 0.0001) {
          double eP[3];
#pragma unroll
          for (int k = 0; k < 3; k++) eP[k] = (1 - cur1) * G.L[LD_EP + k] + cur1 * sP[k];  // reads the MOVED start point
#pragma unroll
          for (int k = 0; k < 3; k++) lns[l].eP[k] = eP[k];
          projection(cam, eP, C.epl);
          C.edisp = get_disparity(cam, eP[2]);
        }
      }
    }
    {
      const double* info = C.invCovPose;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int e = j + 8 * q;
        if (e < 21) G.sum[e] = G.sum[e] + info[ea[q] * 6 + eb[q]];
      }
    }
    __syncwarp(gmask);
    prev_idx_next = prev_idx_next2;
  }
}

// perm[rank] = pair index, pairs ordered by DESCENDING number of matched lines (counting sort by one CTA; the order among
// equal counts is whatever the atomics give — it only decides which warp serves a pair, never a result)
constexpr int ORDER_THREADS = 1024, ORDER_MAX_KEY = 2047;
__global__ void __launch_bounds__(ORDER_THREADS) linecut_order_kernel(const int32_t* __restrict__ n_ln, int B, int32_t* __restrict__ perm) {
  __shared__ int s_cnt[ORDER_MAX_KEY + 2];
  for (int k = threadIdx.x; k < ORDER_MAX_KEY + 2; k += ORDER_THREADS) s_cnt[k] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < B; p += ORDER_THREADS) atomicAdd(&s_cnt[min(max(n_ln[p], 0), ORDER_MAX_KEY)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {   // exclusive offsets in descending key order (2 K additions: negligible next to the search)
    int run = 0;
    for (int k = ORDER_MAX_KEY; k >= 0; k--) { const int c = s_cnt[k]; s_cnt[k] = run; run += c; }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < B; p += ORDER_THREADS) perm[atomicAdd(&s_cnt[min(max(n_ln[p], 0), ORDER_MAX_KEY)], 1)] = p;
}

// prev_ln / prev_cap_ln: the previous frames' stereo lines, or nullptr: prev_cov already holds the covariances
int launch_linecut(gfpl_ctx* ctx, const double* DT_inv, const gfpl_frame_state* state, DevTrackObs obs,
                   gfpl_line_cov* prev_cov, int cov_cap, gfpl_line_cut* cut, const gfpl_line_feature* prev_ln, int prev_cap_ln) {
  if (obs.B <= 0 || obs.cap_ln <= 0) return GFPL_OK;
  const int pairs_per_cta = (CUT_THREADS / 32) * (32 / CUT_GROUP);
  const int grid = (obs.B + pairs_per_cta - 1) / pairs_per_cta;
  // batches of more than a warp's worth of pairs are served in balanced order (see the kernel)
  int32_t* perm = nullptr;
  int grid_bal = grid;
  const int one_wave = ctx->sm_count * (65536 / (CUT_MAX_REGS * CUT_THREADS));   // CTAs resident at once (register-limited)
  // (only batches that fill at least half of a wave: a quarter-filled GPU — 2048 dense pairs — is faster with four pairs
  // per warp on few warps than spread over every slot: 2.20 against 2.45 ms)
  if (CUT_BALANCE && 2 * grid >= one_wave) {
    if (grid < one_wave) grid_bal = one_wave;      // spread over every warp slot of the wave
    GFPL_CUDA_OK(ctx, ctx->pool ? cudaMallocFromPoolAsync((void**)&perm, (size_t)obs.B * sizeof(int32_t), ctx->pool, ctx->stream)
                                : cudaMallocAsync((void**)&perm, (size_t)obs.B * sizeof(int32_t), ctx->stream));
    linecut_order_kernel<<<1, ORDER_THREADS, 0, ctx->stream>>>(obs.n_ln, obs.B, perm);
    ctx->launches++;
  }
  linecut_kernel<<<perm ? grid_bal : grid, CUT_THREADS, 0, ctx->stream>>>(DT_inv, state, obs, prev_cov, cov_cap, prev_ln, prev_cap_ln, cut, perm, ctx->cam, ctx->cfg);
  const cudaError_t le = cudaGetLastError();
  if (perm) cudaFreeAsync(perm, ctx->stream);   // stream-ordered: after the kernel
  if (le != cudaSuccess) { set_last_error(ctx, "line-cut launch failed: %s", cudaGetErrorString(le)); return GFPL_E_CUDA; }
  ctx->launches++;
  return GFPL_OK;
}

// ---- C6: estimateProjUncertainty_descent (stereoFrameHandler.cpp:1476-1610) -------------------------------------------
// Unlike the submodular search the lines of a pair do not see each other here: one THREAD per matched line, the whole
// batch in one launch (B x n_ln independent chains).  A candidate = getPoseCovOnLine (:1249-1338) + the eigenvalues of
// the 6x6 covariance (getEigenValue, linespec.cpp:37-51: sorted, the two largest multiplied); the Jacobi iteration runs
// with the matrix in registers.  Same operations in the same order as the oracle, -fmad=false: the decisions replay
// exactly.

// full 3x3 getJacob3D_2D (stereoFrame.cpp:1394-1413), row-major
__device__ __forceinline__ void jacob3d_2d_full(const gfpl_camera& cam, double px, double py, double pz, double* J) {
  const double f = cam.fx, b = cam.b, pz_2 = pz * pz;
  J[0] = f / pz; J[1] = 0.0; J[2] = -f * px / pz_2;
  J[3] = 0.0; J[4] = f / pz; J[5] = -f * py / pz_2;
  J[6] = 0.0; J[7] = 0.0; J[8] = -f * b / pz_2;
}

// volume of the pose covariance of line data L at cut {c0, c1}
__device__ __noinline__ double descent_volume_at(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT, const double* L,
                                                 double c0, double c1) {
  double Js[6], Je[6];
  const double cov_sR = pose_info_endpoint(*cam, *cfg, DT, &L[LD_SP], &L[LD_EP], &L[LD_CS], &L[LD_CE], &L[LD_LE], c0, Js);
  const double cov_eR = pose_info_endpoint(*cam, *cfg, DT, &L[LD_EP], &L[LD_SP], &L[LD_CE], &L[LD_CS], &L[LD_LE], c1, Je);
  // G = J^T J, G^-1 = adjugate * (1 / det)
  double g00 = Js[0] * Js[0], g01 = Js[0] * Je[0], g11 = Je[0] * Je[0];
#pragma unroll
  for (int i = 1; i < 6; i++) {
    g00 = g00 + Js[i] * Js[i];
    g01 = g01 + Js[i] * Je[i];
    g11 = g11 + Je[i] * Je[i];
  }
  const double det = g00 * g11 - g01 * g01;
  const double invdet = 1.0 / det;
  const double i00 = g11 * invdet, i01 = -g01 * invdet, i10 = -g01 * invdet, i11 = g00 * invdet;
  double Jinv[12], A[12], S[36];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    Jinv[j] = i00 * Js[j] + i01 * Je[j];
    Jinv[6 + j] = i10 * Js[j] + i11 * Je[j];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    A[i * 2 + 0] = Jinv[i] * cov_sR + Jinv[6 + i] * 0.0;
    A[i * 2 + 1] = Jinv[i] * 0.0 + Jinv[6 + i] * cov_eR;
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {   // the solver reads the lower triangle, mirrored
      const double v = A[i * 2 + 0] * Jinv[j] + A[i * 2 + 1] * Jinv[6 + j];
      S[i * 6 + j] = v; S[j * 6 + i] = v;
    }
  double ev[6];
  sym_eigenvalues_unrolled<6>(S, ev);
#pragma unroll
  for (int i = 1; i < 6; i++) {   // ascending insertion sort (any correct sort gives the oracle's std::sort outcome)
#pragma unroll
    for (int j = i; j > 0; j--) {
      const double a = ev[j - 1], b = ev[j];
      const bool sw = a > b;
      ev[j - 1] = sw ? b : a; ev[j] = sw ? a : b;
    }
  }
  return ev[5] * ev[4];
}

__global__ void __launch_bounds__(64)
linecut_descent_kernel(const double* __restrict__ DT_inv_all, DevTrackObs obs, const gfpl_line_cov* __restrict__ prev_cov,
                       int cov_cap, gfpl_line_cut_descent* __restrict__ out_all, gfpl_camera cam, gfpl_config cfg) {
  __shared__ gfpl_camera s_cam;
  __shared__ gfpl_config s_cfg;
  if (threadIdx.x == 0) { s_cam = cam; s_cfg = cfg; }
  __syncthreads();
  const int p = blockIdx.y, l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= obs.n_ln[p]) return;
  gfpl_ln_obs& o = obs.ln[(size_t)p * obs.cap_ln + l];
  const gfpl_line_cov& c = prev_cov[(size_t)p * cov_cap + o.prev_idx];
  gfpl_line_cut_descent& C = out_all[(size_t)p * obs.cap_ln + l];
  double DT[12], T4[16], L[LD_SIZE];
#pragma unroll
  for (int i = 0; i < 16; i++) T4[i] = DT_inv_all[(size_t)p * 16 + i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    DT[i * 3 + 0] = T4[i * 4 + 0]; DT[i * 3 + 1] = T4[i * 4 + 1]; DT[i * 3 + 2] = T4[i * 4 + 2];
    DT[9 + i] = T4[i * 4 + 3];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { L[LD_SP + k] = o.sP[k]; L[LD_EP + k] = o.eP[k]; }
  L[LD_LE] = o.le_obs[0]; L[LD_LE + 1] = o.le_obs[1];
#pragma unroll
  for (int k = 0; k < 9; k++) { L[LD_CS + k] = c.covS[k]; L[LD_CE + k] = c.covE[k]; }

  const double step = cfg.cut_step, rng0 = cfg.cut_rng[0], rng1 = cfg.cut_rng[1];
  double cut0 = 0.0, cut1 = 0.0;
  double vol_last = descent_volume_at(&s_cam, &s_cfg, DT, L, cut0, cut1);
  int n_vol = 1;
  while (cut0 + cut1 < 1.0) {
    bool hitLower = false;
    for (int j = 0; j < 8; j++) {
      const int d0 = (j == 0 || j == 4 || j == 5) ? 1 : ((j == 1 || j == 6 || j == 7) ? -1 : 0);
      const int d1 = (j == 2 || j == 4 || j == 6) ? 1 : ((j == 3 || j == 5 || j == 7) ? -1 : 0);
      const double ct0 = cut0 + (d0 > 0 ? step : (d0 < 0 ? -step : 0.0)), ct1 = cut1 + (d1 > 0 ? step : (d1 < 0 ? -step : 0.0));
      if (ct0 + ct1 >= 1.0) continue;
      if (ct0 < rng0 || ct0 > rng1) continue;
      if (ct1 < rng0 || ct1 > rng1) continue;
      const double vol = descent_volume_at(&s_cam, &s_cfg, DT, L, ct0, ct1);
      if (vol < vol_last) {   // the cut moves at once: later neighbours of this round are taken around the new cut (:1545-1550)
        vol_last = vol; n_vol++;
        cut0 = ct0; cut1 = ct1;
        hitLower = true;
      }
    }
    if (!hitLower) break;
  }
  C.cutRatio[0] = cut0; C.cutRatio[1] = cut1;
  C.volume = vol_last; C.n_volume = n_vol;
  C.fst_idx = 0; C.lst_idx = 3; C.pad = 0;
  // :1567-1606 — both end points move from the ORIGINAL end points
  double sP_cut[3], eP_cut[3], cov_s[9], cov_e[9], R[9], Rt[9];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    sP_cut[k] = (1 - cut0) * L[LD_SP + k] + cut0 * L[LD_EP + k];
    eP_cut[k] = (1 - cut1) * L[LD_EP + k] + cut1 * L[LD_SP + k];
  }
  const double a0 = (1 - cut0) * (1 - cut0), b0 = cut0 * cut0, a1 = (1 - cut1) * (1 - cut1), b1 = cut1 * cut1;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    cov_s[k] = a0 * L[LD_CS + k] + b0 * L[LD_CE + k];
    cov_e[k] = a1 * L[LD_CE + k] + b1 * L[LD_CS + k];
    R[k] = DT[k];
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Rt[j * 3 + i] = R[i * 3 + j];
  double cur_s[3], cur_e[3], cur_s0[3], cur_e0[3], Jp[9], Jpt[9], T1[9], T2[9], T3[9], out9[9];
  se3_apply(T4, sP_cut, cur_s);
  se3_apply(T4, eP_cut, cur_e);
  se3_apply(T4, &L[LD_SP], cur_s0);
  se3_apply(T4, &L[LD_EP], cur_e0);
  for (int e = 0; e < 2; e++) {   // covlPt2D[1] / [2] = J_proj * J_dt * cov * J_dt^T * J_proj^T, left to right
    const double* cur = e == 0 ? cur_s : cur_e;
    jacob3d_2d_full(cam, cur[0], cur[1], cur[2], Jp);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Jpt[j * 3 + i] = Jp[i * 3 + j];
    matmul<3, 3, 3>(Jp, R, T1);
    matmul<3, 3, 3>(T1, e == 0 ? cov_s : cov_e, T2);
    matmul<3, 3, 3>(T2, Rt, T3);
    matmul<3, 3, 3>(T3, Jpt, out9);
    for (int k = 0; k < 9; k++) C.covlPt2D[e][k] = out9[k];
  }
  projection(cam, cur_s0, C.lPt2D[0]);
  projection(cam, cur_s, C.lPt2D[1]);
  projection(cam, cur_e, C.lPt2D[2]);
  projection(cam, cur_e0, C.lPt2D[3]);
#pragma unroll
  for (int k = 0; k < 3; k++) { o.sP[k] = sP_cut[k]; o.eP[k] = eP_cut[k]; }
  projection(cam, sP_cut, C.spl);
  projection(cam, eP_cut, C.epl);
  C.sdisp = get_disparity(cam, sP_cut[2]);
  C.edisp = get_disparity(cam, eP_cut[2]);
}

int launch_linecut_descent(gfpl_ctx* ctx, const double* DT_inv, DevTrackObs obs, const gfpl_line_cov* prev_cov, int cov_cap,
                           gfpl_line_cut_descent* out) {
  if (obs.B <= 0 || obs.cap_ln <= 0) return GFPL_OK;
  dim3 grid((obs.cap_ln + 63) / 64, obs.B);
  linecut_descent_kernel<<<grid, 64, 0, ctx->stream>>>(DT_inv, obs, prev_cov, cov_cap, out, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
