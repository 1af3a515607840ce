// k_lba.cu — K6: residual / Jacobian / block-sparse normal-equation assembly of one pass of local bundle
// adjustment.
//
// Replaces the per-observation loops of PLSLAM::MapHandler::levMarquardtOptimizationLBA (reference
// mapHandler.cpp:1243-1317 + 1322-1426 first pass, 1466-1543 + 1545-1649 iteration pass).  The reference
// accumulates into a dense N x N MatrixXd (re-zeroed every iteration); here the same blocks are emitted
// block-sparse: per-KF 6x6, per-landmark 3x3 / 6x6, and one off-diagonal 3x6 / 6x6 block per observation
// (an (landmark, KF) pair is observed once, so those are plain stores).  One thread per observation;
// KF blocks are reduced in shared memory per CTA and flushed with fp64 atomics, landmark blocks go
// straight to global fp64 atomics (<= ~4 observations per landmark).  HBM-bound:
// bytes = #obs * (12 idx + 16..24 obs) + state reads + block writes.
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

constexpr int LBA_THREADS = 128;
constexpr int LBA_MAX_KF_SMEM = 64;

__device__ __forceinline__ void jac_lba(double gz2, double fxa, double fyb, double gx, double gy, double gz, double* J) {
  J[0] = +gz2 * fxa * gz;
  J[1] = +gz2 * fyb * gz;
  J[2] = -gz2 * (fxa * gx + fyb * gy);
  J[3] = -gz2 * (fxa * gx * gy + fyb * gy * gy + fyb * gz * gz);
  J[4] = +gz2 * (fxa * gx * gx + fxa * gz * gz + fyb * gx * gy);
  J[5] = +gz2 * (fyb * gx * gz - fxa * gy * gz);
}

// row-vector (j^T * R) of the inverted pose, R = Tiw.block(0,0,3,3)
__device__ __forceinline__ void jt_times_R(const double* j3, const double* Tiw, double* out) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double s = j3[0] * Tiw[0 * 4 + c];
    s = s + j3[1] * Tiw[1 * 4 + c];
    s = s + j3[2] * Tiw[2 * 4 + c];
    out[c] = s;
  }
}

// grid = (observation tiles of the largest window, windows): blockIdx.y selects the window (independent local-BA problems
// of a batched replay: one per key-frame insertion)
template <bool LINES>
__global__ void __launch_bounds__(LBA_THREADS) lba_assemble_kernel(const LbaWin* __restrict__ wins, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ double s_kf[];  // [Nkf][42]: 36 H + 6 g, when Nkf <= LBA_MAX_KF_SMEM
  __shared__ double s_err[LBA_THREADS / 32];
  const gfpl_lba_problem pb = wins[blockIdx.y].pb;
  const gfpl_lba_blocks out = wins[blockIdx.y].H;
  if ((size_t)blockIdx.x * LBA_THREADS >= (size_t)(LINES ? pb.n_ls_obs : pb.n_pt_obs)) return;   // a smaller window of the batch
  const bool use_smem = pb.Nkf <= LBA_MAX_KF_SMEM;
  if (use_smem) {
    for (int i = threadIdx.x; i < pb.Nkf * 42; i += LBA_THREADS) s_kf[i] = 0.0;
    __syncthreads();
  }
  const int n_obs = LINES ? pb.n_ls_obs : pb.n_pt_obs;
  const int o = blockIdx.x * LBA_THREADS + threadIdx.x;
  double err = 0.0;
  if (o < n_obs) {
    const int lm = LINES ? pb.ls_obs_lm[o] : pb.pt_obs_lm[o];
    const int kf = LINES ? pb.ls_obs_kf[o] : pb.pt_obs_kf[o];
    const int fix = LINES ? pb.ls_obs_fix[o] : pb.pt_obs_fix[o];
    const double* Tkw = kf >= 0 ? pb.kf_Tkw + (size_t)kf * 16 : pb.fix_Tkw + (size_t)fix * 16;
    double T[16], Tiw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) T[i] = Tkw[i];
    inverse_se3(T, Tiw);
    const double fx = cam.fx, fy = cam.fy;
    double JT[6], JL[6], nrm, w;
    if (!LINES) {
      double X[3], Xwi[3], prj[2];
#pragma unroll
      for (int i = 0; i < 3; i++) X[i] = pb.X_pt[(size_t)lm * 3 + i];
      se3_apply(Tiw, X, Xwi);
      projection(cam, Xwi, prj);
      const double dx = pb.pt_obs_uv[2 * (size_t)o] - prj[0], dy = pb.pt_obs_uv[2 * (size_t)o + 1] - prj[1];
      nrm = sqrt(dx * dx + dy * dy);
      const double gz2 = 1.0 / fmax(cfg.homog_th, Xwi[2] * Xwi[2]);
      jac_lba(gz2, fx * dx, fy * dy, Xwi[0], Xwi[1], Xwi[2], JT);
      const double dn = fmax(cfg.homog_th, nrm);
      double j3[3] = {JT[0], JT[1], JT[2]}, r3[3];
      jt_times_R(j3, Tiw, r3);
#pragma unroll
      for (int i = 0; i < 6; i++) JT[i] = JT[i] / dn;
#pragma unroll
      for (int i = 0; i < 3; i++) JL[i] = r3[i] / dn;
      w = 1.0 / (1.0 + nrm * nrm * pb.pt_obs_s2[o]);
    } else {
      double P[3], Q[3], Pwi[3], Qwi[3], pp[2], qp[2];
#pragma unroll
      for (int i = 0; i < 3; i++) { P[i] = pb.X_ls[(size_t)lm * 6 + i]; Q[i] = pb.X_ls[(size_t)lm * 6 + 3 + i]; }
      se3_apply(Tiw, P, Pwi);
      se3_apply(Tiw, Q, Qwi);
      projection(cam, Pwi, pp);
      projection(cam, Qwi, qp);
      const double* l = pb.ls_obs_le + (size_t)o * 3;
      const double e0 = l[0] * pp[0] + l[1] * pp[1] + l[2];
      const double e1 = l[0] * qp[0] + l[1] * qp[1] + l[2];
      nrm = sqrt(e0 * e0 + e1 * e1);
      const double dn = fmax(cfg.homog_th, nrm);
      const double fxlx = fx * e0, fyly = fy * e1;  // endpoint errors stand in for (lx, ly)
      double JP[6], JQ[6], rP[3], rQ[3];
      jac_lba(1.0 / fmax(cfg.homog_th, Pwi[2] * Pwi[2]), fxlx, fyly, Pwi[0], Pwi[1], Pwi[2], JP);
      jac_lba(1.0 / fmax(cfg.homog_th, Qwi[2] * Qwi[2]), fxlx, fyly, Qwi[0], Qwi[1], Qwi[2], JQ);
      jt_times_R(JP, Tiw, rP);
      jt_times_R(JQ, Tiw, rQ);
#pragma unroll
      for (int c = 0; c < 3; c++) { JL[c] = rP[c] * e0 / dn; JL[3 + c] = rQ[c] * e1 / dn; }
#pragma unroll
      for (int i = 0; i < 6; i++) JT[i] = (JP[i] * e0 + JQ[i] * e1) / dn;
      w = 1.0 / (1.0 + nrm * nrm * pb.ls_obs_s2[o]);
    }
    err = nrm * nrm * w;
    constexpr int LD = LINES ? 6 : 3;
    double* gl = out.g + 6 * pb.Nkf + (LINES ? 3 * pb.Npt + 6 * lm : 3 * lm);
    double* Hl = LINES ? out.Hll + (size_t)lm * 36 : out.Hpp + (size_t)lm * 9;
#pragma unroll
    for (int i = 0; i < LD; i++) {
      atomicAdd(&gl[i], JL[i] * nrm * w);
#pragma unroll
      for (int jx = 0; jx < LD; jx++) atomicAdd(&Hl[i * LD + jx], JL[i] * JL[jx] * w);
    }
    if (kf >= 0) {
      double* Hoff = LINES ? out.Hlk + (size_t)o * 36 : out.Hpk + (size_t)o * 18;
#pragma unroll
      for (int i = 0; i < LD; i++)
#pragma unroll
        for (int jx = 0; jx < 6; jx++) Hoff[i * 6 + jx] = JL[i] * JT[jx] * w;
      double* Hk = use_smem ? s_kf + kf * 42 : out.Hkk + (size_t)kf * 36;
      double* gk = use_smem ? s_kf + kf * 42 + 36 : out.g + 6 * kf;
#pragma unroll
      for (int i = 0; i < 6; i++) {
        atomicAdd(&gk[i], JT[i] * nrm * w);
#pragma unroll
        for (int jx = 0; jx < 6; jx++) atomicAdd(&Hk[i * 6 + jx], JT[i] * JT[jx] * w);
      }
    }
  }
  // error: warp shuffle -> smem -> one atomic per CTA
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) err = err + shfl_xor_f64(0xffffffffu, err, s);
  if ((threadIdx.x & 31) == 0) s_err[threadIdx.x >> 5] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int i = 0; i < LBA_THREADS / 32; i++) e = e + s_err[i];
    atomicAdd(out.err, e);
  }
  if (use_smem) {
    for (int i = threadIdx.x; i < pb.Nkf * 42; i += LBA_THREADS) {
      const double v = s_kf[i];
      if (v != 0.0) {
        const int kf = i / 42, r = i % 42;
        atomicAdd(r < 36 ? &out.Hkk[(size_t)kf * 36 + r] : &out.g[6 * kf + (r - 36)], v);
      }
    }
  }
}

// block storage of every window must be zero on entry (the callers clear the one device buffer that holds them all)
int launch_lba_assemble(gfpl_ctx* ctx, const LbaWin* wins, int W, int max_kf, int max_pt_obs, int max_ls_obs) {
  const size_t smem = max_kf <= LBA_MAX_KF_SMEM ? (size_t)(max_kf > 0 ? max_kf : 1) * 42 * 8 : 8;
  if (max_pt_obs > 0) {
    lba_assemble_kernel<false><<<dim3((max_pt_obs + LBA_THREADS - 1) / LBA_THREADS, W), LBA_THREADS, smem, ctx->stream>>>(wins, ctx->cam, ctx->cfg);
    GFPL_LAUNCH_OK(ctx);
  }
  if (max_ls_obs > 0) {
    lba_assemble_kernel<true><<<dim3((max_ls_obs + LBA_THREADS - 1) / LBA_THREADS, W), LBA_THREADS, smem, ctx->stream>>>(wins, ctx->cam, ctx->cfg);
    GFPL_LAUNCH_OK(ctx);
  }
  return GFPL_OK;
}

}  // namespace gfpl
