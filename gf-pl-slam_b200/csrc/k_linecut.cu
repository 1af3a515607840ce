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

constexpr int CUT_THREADS = 128;
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

// getPoseInfoOnLine up to the two Jacobian columns and the 2x2 inverse.  `DT` = [R (9) | t (3)], `L` = line data;
// both may live in shared memory (pass 3: one copy per 8-lane group) or be read straight from global (pass 1).
__device__ __forceinline__ void pose_info_on_line(const gfpl_camera& cam, const gfpl_config& cfg, const double* DT,
                                                  const double* L, double c0, double c1, LineInfo& out) {
  double R[9], T4[16];
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = DT[i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    T4[i * 4 + 0] = R[i * 3 + 0]; T4[i * 4 + 1] = R[i * 3 + 1]; T4[i * 4 + 2] = R[i * 3 + 2]; T4[i * 4 + 3] = DT[9 + i];
  }
  const double le[2] = {L[LD_LE], L[LD_LE + 1]};
  double cov_sR, cov_eR;
  {
    double P[3], cur[3], cv[9], Jp[6];
    const double a0 = (1 - c0) * (1 - c0), b0 = c0 * c0;
#pragma unroll
    for (int k = 0; k < 3; k++) P[k] = (1 - c0) * L[LD_SP + k] + c0 * L[LD_EP + k];
#pragma unroll
    for (int k = 0; k < 9; k++) cv[k] = a0 * L[LD_CS + k] + b0 * L[LD_CE + k];
    se3_apply(T4, P, cur);
    jacob3d_2d(cam, cur[0], cur[1], cur[2], Jp);
    cov_sR = residual_var(Jp, R, cv, le);
    jac6(cam.fx / fmax(cfg.homog_th, cur[2] * cur[2]), cur[0], cur[1], cur[2], le[0], le[1], out.Js);
  }
  {
    double P[3], cur[3], cv[9], Jp[6];
    const double a1 = (1 - c1) * (1 - c1), b1 = c1 * c1;
#pragma unroll
    for (int k = 0; k < 3; k++) P[k] = (1 - c1) * L[LD_EP + k] + c1 * L[LD_SP + k];
#pragma unroll
    for (int k = 0; k < 9; k++) cv[k] = a1 * L[LD_CE + k] + b1 * L[LD_CS + k];
    se3_apply(T4, P, cur);
    jacob3d_2d(cam, cur[0], cur[1], cur[2], Jp);
    cov_eR = residual_var(Jp, R, cv, le);
    jac6(cam.fx / fmax(cfg.homog_th, cur[2] * cur[2]), cur[0], cur[1], cur[2], le[0], le[1], out.Je);
  }
  // 2x2 inverse of diag(cov_sR, cov_eR) as adjugate * (1/det)
  const double det = cov_sR * cov_eR - 0.0 * 0.0;
  const double invdet = 1.0 / det;
  out.i00 = cov_eR * invdet;
  out.i11 = cov_sR * invdet;
  out.i01 = -0.0 * invdet;
  out.i10 = -0.0 * invdet;
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

// one published half: u[6] (Jacobian column), w[6] = S^-1 u, a = 1 / residual variance, g = u' S^-1 u
constexpr int HF_U = 0, HF_W = 6, HF_A = 12, HF_G = 13, HF_SIZE = 14;

__device__ __forceinline__ double dot3f(const double* a, const double* b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }

// reciprocal to ~1e-15 relative: float seed + 2 Newton steps (screening path only)
__device__ __forceinline__ double rcp_fast(double x) {
  double r = (double)__frcp_rn((float)x);
  r = r * fma(-x, r, 2.0);
  r = r * fma(-x, r, 2.0);
  return r;
}

// one endpoint of getPoseInfoOnLine at cut ratio c: A/SA = the endpoint that shrinks towards B/SB.  Result -> out[HF_SIZE]
__device__ __forceinline__ void half_fast(const gfpl_camera& cam, const gfpl_config& cfg, const double* DT, const double* A,
                                          const double* Bp, const double* SA, const double* SB, const double* le,
                                          const double* Sinv /*6x6 row-major*/, double c, double* out) {
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
  double2* o2 = reinterpret_cast<double2*>(out);  // [u0..u5 | w0..w5 | a g]
#pragma unroll
  for (int i = 0; i < 3; i++) {
    o2[HF_U / 2 + i] = make_double2(u[2 * i], u[2 * i + 1]);
    o2[HF_W / 2 + i] = make_double2(w[2 * i], w[2 * i + 1]);
  }
  o2[HF_A / 2] = make_double2(a, g);
}

// S^-1 (full 6x6 row-major) of the lower-packed SPD matrix S (21) by Cholesky; false when S is not safely positive definite
__device__ __forceinline__ bool spd_inverse6(const double* S, double* Sinv) {
  double L[21];
#pragma unroll
  for (int i = 0; i < 21; i++) L[i] = S[i];
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
  // columns of S^-1: solve L y = e_c, L^T x = y
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double sv = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; k++) sv = fma(-L[tri(i, k)], y[k], sv);
      y[i] = (i < c) ? 0.0 : sv * L[tri(i, i)];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      double sv = y[i];
#pragma unroll
      for (int k = i + 1; k < 6; k++) sv = fma(-L[tri(k, i)], y[k], sv);
      y[i] = sv * L[tri(i, i)];
    }
#pragma unroll
    for (int i = c; i < 6; i++) { Sinv[i * 6 + c] = y[i]; Sinv[c * 6 + i] = y[i]; }
  }
#pragma unroll
  for (int i = 0; i < 36; i++) ok = ok && (fabs(Sinv[i]) < 1e300);
  return ok;
}

// per-group (one frame pair) state kept in shared memory: every lane of the group reads the same addresses
// (broadcast), which frees ~120 registers per thread and lets 16 warps / SM hide the fp64 dependency chains.
struct __align__(16) GroupState {
  double Sinv[36];          // inverse of (invCov_sum - info(line)) for the screening path, row-major (16-byte aligned rows)
  double half[6][HF_SIZE];  // the six published halves of the current round (112-byte rows)
  double DT[12];            // R | t of DT_inv
  double L[LD_SIZE];        // current line
  double J[CUT_GROUP][6];   // point Jacobians of the current chunk of the invCov_sum pass
  double sum[21];           // invCov_sum, lower triangle
  double sum0[21];          // invCov_sum before the current line's information is removed (exact metric of cut {0,0})
};
static_assert(sizeof(GroupState) % 16 == 0 && HF_SIZE % 2 == 0, "128-bit shared-memory accesses below");

#ifndef CUT_MIN_BLOCKS
#define CUT_MIN_BLOCKS 4
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
__device__ __noinline__ void exact_info_store(const gfpl_camera* cam, const gfpl_config* cfg, const GroupState* G, double c0,
                                              double c1, int j, double* invCovPose) {
  LineInfo I;
  pose_info_on_line(*cam, *cfg, G->DT, G->L, c0, c1, I);
  for (int e = j; e < 36; e += CUT_GROUP) invCovPose[e] = info_entry(I, e / 6, e % 6);
}

__device__ __forceinline__ double shfl_up_f64(unsigned mask, double v, int delta, int width) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(mask, lo, delta, width);
  hi = __shfl_up_sync(mask, hi, delta, width);
  return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(CUT_THREADS, CUT_MIN_BLOCKS)
linecut_kernel(const double* __restrict__ DT_inv_all, const gfpl_frame_state* __restrict__ state, DevTrackObs obs,
               const gfpl_line_cov* __restrict__ prev_cov, int cov_cap, gfpl_line_cut* __restrict__ cut_all,
               gfpl_camera cam, gfpl_config cfg) {
  __shared__ GroupState s_grp[CUT_THREADS / CUT_GROUP];
  __shared__ gfpl_camera s_cam;
  __shared__ gfpl_config s_cfg;
  if (threadIdx.x == 0) { s_cam = cam; s_cfg = cfg; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int j = lane & (CUT_GROUP - 1);             // neighbour index of this lane
  const int grp = lane / CUT_GROUP;                 // pair slot inside the warp
  const unsigned gmask = 0xFFu << (grp * CUT_GROUP);
  const int p = (blockIdx.x * (CUT_THREADS / 32) + (threadIdx.x >> 5)) * (32 / CUT_GROUP) + grp;
  if (p >= obs.B) return;
  const int nln = obs.n_ln[p], npt = obs.n_pt[p];
  if (nln == 0) return;  // :1621
  GroupState& G = s_grp[threadIdx.x / CUT_GROUP];
  gfpl_ln_obs* lns = obs.ln + (size_t)p * obs.cap_ln;
  const gfpl_pt_obs* pts = obs.pt + (size_t)p * obs.cap_pt;
  const gfpl_line_cov* cov = prev_cov + (size_t)p * cov_cap;
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

  const double step = cfg.cut_step;
  const double nb0 = (j == 0 || j == 4 || j == 5) ? step : ((j == 1 || j == 6 || j == 7) ? -step : 0.0);
  const double nb1 = (j == 2 || j == 4 || j == 6) ? step : ((j == 3 || j == 5 || j == 7) ? -step : 0.0);

  // pass 1a: initial information of every line (cut = 0), lanes stride over the lines (:1638-1649)
  for (int l = j; l < nln; l += CUT_GROUP) exact_info_initial(&s_cam, &s_cfg, G.DT, &lns[l], &cov[lns[l].prev_idx], &cut[l]);
  __syncwarp(gmask);

  // pass 1b/2: invCov_sum in the reference's order — lines, then points, each element summed sequentially.
  // Lane j owns lower-triangle elements e = j, j+8, j+16.
  {
    int ea[3], eb[3];
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const int e = min(j + 8 * q, 20);
      int a = 0;
      while ((a + 1) * (a + 2) / 2 <= e) a++;
      ea[q] = a; eb[q] = e - a * (a + 1) / 2;
    }
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
  for (int l = 0; l < nln; l++) {
    gfpl_line_cut& C = cut[l];
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
      for (int e = j; e < 21; e += CUT_GROUP) {
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= e) a++;
        const int b = e - a * (a + 1) / 2;
        G.sum0[e] = G.sum[e];
        G.sum[e] = G.sum[e] - info[a * 6 + b];
      }
    }
    __syncwarp(gmask);
    // screening set-up: S^-1 once per line (every lane computes it redundantly, lanes publish it together)
    bool fast_ok = cfg.max_vol_line_cut != 0;
    if (fast_ok) fast_ok = spd_inverse6_publish(&G, j, gmask);
    __syncwarp(gmask);
    double cur0 = 0.0, cur1 = 0.0;
    double metric_last = 0.0;      // the reference's metric of the current cut; evaluated lazily (exact rounds only)
    bool have_exact_last = false;
    bool at_start = true;          // current cut is still {0,0}: its exact metric is logdet(sum0)
    double dcur = 0.0;             // screening determinant factor of the current cut
    bool have_dcur = false;
    while (cur0 + cur1 <= 1.0) {
      const double ct0 = cur0 + nb0, ct1 = cur1 + nb1;
      bool valid = !(ct0 + ct1 > 1.0);
      valid = valid && !(ct0 < cfg.cut_rng[0] || ct0 > cfg.cut_rng[1]);
      valid = valid && !(ct1 < cfg.cut_rng[0] || ct1 > cfg.cut_rng[1]);
      int cand = -1;
      bool decided = false;
      if (fast_ok) {
        // lanes 0..2: start-point halves at c0 - step, c0, c0 + step; lanes 3..5: end-point halves likewise
        if (j < 6) {
          const bool isS = j < 3;
          const int kk = (j % 3) - 1;
          const double ch = (isS ? cur0 : cur1) + (kk < 0 ? -step : (kk > 0 ? step : 0.0));
          half_fast(cam, cfg, G.DT, isS ? &G.L[LD_SP] : &G.L[LD_EP], isS ? &G.L[LD_EP] : &G.L[LD_SP],
                    isS ? &G.L[LD_CS] : &G.L[LD_CE], isS ? &G.L[LD_CE] : &G.L[LD_CS], &G.L[LD_LE], G.Sinv, ch, G.half[j]);
        }
        __syncwarp(gmask);
        // determinant factor of (si, ei): (1 + a u'S^-1u)(1 + b v'S^-1v) - a b (u'S^-1v)^2  (monotone in the metric)
        auto factor = [&](int si, int ei) {
          const double2* hs = reinterpret_cast<const double2*>(G.half[si]);
          const double2* he = reinterpret_cast<const double2*>(G.half[ei]);
          double guv = 0.0;
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const double2 uu = hs[HF_U / 2 + i], ww = he[HF_W / 2 + i];
            guv = fma(uu.x, ww.x, guv);
            guv = fma(uu.y, ww.y, guv);
          }
          const double2 ags = hs[HF_A / 2], age = he[HF_A / 2];   // (a, g) of each half
          const double a = ags.x, b = age.x;
          return fma(-(a * b) * guv, guv, fma(a, ags.y, 1.0) * fma(b, age.y, 1.0));
        };
        const int si = (nb0 < 0.0) ? 0 : ((nb0 > 0.0) ? 2 : 1);
        const int ei = 3 + ((nb1 < 0.0) ? 0 : ((nb1 > 0.0) ? 2 : 1));
        const double dj = factor(si, ei);          // this lane's candidate
        if (!have_dcur) { dcur = factor(1, 4); have_dcur = true; }   // the current cut (same value in every lane)
        __syncwarp(gmask);
        // replay the reference's selection ("first strictly better in j order", :1722-1730) on the factors, but only if
        // every comparison it makes is clear by > CUT_TAU (relative).  Candidate j is compared with the running best
        // before it = max(current cut, valid candidates i < j): an exclusive prefix maximum over the group's lanes.
        const double mv = valid ? dj : 0.0;   // usable factors are > 0: a masked lane never raises the maximum
        double pre = mv;
#pragma unroll
        for (int dlt = 1; dlt < CUT_GROUP; dlt <<= 1) {
          const double t = shfl_up_f64(gmask, pre, dlt, CUT_GROUP);
          pre = (j >= dlt) ? fmax(pre, t) : pre;
        }
        const double excl = shfl_up_f64(gmask, pre, 1, CUT_GROUP);
        const double before = (j == 0) ? dcur : fmax(dcur, excl);
        const bool lane_clear = !valid || ((dj > 0.0) && dj < 1e300 && fabs(dj - before) > CUT_TAU * before);
        const unsigned okmask = __ballot_sync(gmask, lane_clear) & gmask;
        const bool clear = (dcur > 0.0) && dcur < 1e300 && okmask == gmask;
        const double best = fmax(dcur, shfl_f64(gmask, pre, CUT_GROUP - 1, CUT_GROUP));
        const unsigned winners = __ballot_sync(gmask, valid && dj == best) & gmask;
        const int bj = (best > dcur && winners) ? (__ffs(winners) - 1) - grp * CUT_GROUP : -1;
        if (clear) { decided = true; cand = bj; have_exact_last = false; dcur = best; }
      }
      if (!decided) {
        // exact round: the reference's arithmetic (also the only path for the min-eigenvalue metric)
        if (!have_exact_last) {  // metric of the current cut, needed as the running best
          metric_last = at_start ? exact_metric_start(&s_cfg, &G) : exact_metric_at(&s_cam, &s_cfg, &G, cur0, cur1);
          have_exact_last = true;
        }
        const double metric = valid ? exact_metric_at(&s_cam, &s_cfg, &G, ct0, ct1) : 0.0;
        // "first strictly better in j order" (:1722-1730)
        double best = metric_last;
#pragma unroll
        for (int jj = 0; jj < CUT_GROUP; jj++) {
          const double mj = shfl_f64(gmask, metric, jj, CUT_GROUP);
          const int vj = __shfl_sync(gmask, (int)valid, jj, CUT_GROUP);
          if (vj && mj > best) { best = mj; cand = jj; }
        }
        metric_last = best;
        have_dcur = false;   // the screening factor of the new current cut is re-evaluated next round
      }
      if (cand < 0) break;  // early termination (:1741-1743)
      cur0 = shfl_f64(gmask, ct0, cand, CUT_GROUP);
      cur1 = shfl_f64(gmask, ct1, cand, CUT_GROUP);
      at_start = false;
    }
    // invCovPose of the accepted cut, in the reference's arithmetic (cut {0,0} is already stored by pass 1)
    if (cur0 != 0.0 || cur1 != 0.0) exact_info_store(&s_cam, &s_cfg, &G, cur0, cur1, j, C.invCovPose);
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
      for (int e = j; e < 21; e += CUT_GROUP) {
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= e) a++;
        const int b = e - a * (a + 1) / 2;
        G.sum[e] = G.sum[e] + info[a * 6 + b];
      }
    }
    __syncwarp(gmask);
  }
}

int launch_linecut(gfpl_ctx* ctx, const double* DT_inv, const gfpl_frame_state* state, DevTrackObs obs,
                   const gfpl_line_cov* prev_cov, int cov_cap, gfpl_line_cut* cut) {
  if (obs.B <= 0 || obs.cap_ln <= 0) return GFPL_OK;
  const int pairs_per_cta = (CUT_THREADS / 32) * (32 / CUT_GROUP);
  const int grid = (obs.B + pairs_per_cta - 1) / pairs_per_cta;
  linecut_kernel<<<grid, CUT_THREADS, 0, ctx->stream>>>(DT_inv, state, obs, prev_cov, cov_cap, cut, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
