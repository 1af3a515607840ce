// k_lba_solve.cu — damped normal-equation solve of one Levenberg-Marquardt pass of local bundle adjustment on the
// block-sparse system that k_lba.cu assembles.
//
// The reference copies a dense N x N matrix into a sparse one and factorises it with Eigen::SimplicialLDLT
// (mapHandler.cpp:1436-1441, 1654-1660).  H has a block-diagonal landmark part (one 3x3 / 6x6 block per landmark), so
// the same solution  DX = (H + lambda diag(H))^-1 g  is obtained by eliminating the landmarks:
//   S   = Hkk' - sum_j  W_j^T D_j'^-1 W_j          (6 Nkf x 6 Nkf, dense, tiny)
//   rhs = g_k  - sum_j  W_j^T D_j'^-1 g_j
//   dx_k = S^-1 rhs ;  dx_j = D_j'^-1 (g_j - W_j dx_k)
// (' = diagonal scaled by 1 + lambda).  Kernels: diag-max / S init, one warp per landmark for the elimination, one
// CTA for the dense LDL^T, one thread per landmark row for the back substitution.
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

// lambda_eff = lambda * (int) max |H_ii|  (quirk iii: `int Hmax`, mapHandler.cpp:1429-1435), or lambda itself
__global__ void __launch_bounds__(256)
lba_lambda_kernel(const LbaWin* __restrict__ wins, double lambda, int scale_by_hmax) {
  __shared__ double s_max[8];
  const gfpl_lba_problem pb = wins[blockIdx.x].pb;
  const gfpl_lba_blocks H = wins[blockIdx.x].H;
  double* lambda_eff = wins[blockIdx.x].lambda_eff;
  double m = 0.0;
  if (scale_by_hmax) {
    const int nk = pb.Nkf * 6, np = pb.Npt * 3, nl = pb.Nls * 6;
    for (int i = threadIdx.x; i < nk + np + nl; i += 256) {
      double h;
      if (i < nk) h = H.Hkk[(size_t)(i / 6) * 36 + (i % 6) * 7];
      else if (i < nk + np) { const int k = i - nk; h = H.Hpp[(size_t)(k / 3) * 9 + (k % 3) * 4]; }
      else { const int k = i - nk - np; h = H.Hll[(size_t)(k / 6) * 36 + (k % 6) * 7]; }
      h = fabs(h);
      if (h > m) m = h;  // NaN never wins, like `h > Hmax || h < -Hmax`
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++) m = fmax(m, s_max[w]);
      *lambda_eff = lambda * (double)(int)m;
    }
  } else if (threadIdx.x == 0) {
    *lambda_eff = lambda;
  }
}

// S = damped block diagonal of Hkk, rhs = g[0 : 6 Nkf]
__global__ void lba_schur_init_kernel(const LbaWin* __restrict__ wins) {
  const gfpl_lba_problem pb = wins[blockIdx.y].pb;
  const gfpl_lba_blocks H = wins[blockIdx.y].H;
  const double* lambda_eff = wins[blockIdx.y].lambda_eff;
  double* S = wins[blockIdx.y].w.S;
  double* rhs = wins[blockIdx.y].DX;
  const int nk = pb.Nkf * 6;
  const double lam = *lambda_eff;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)nk * nk; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / nk), c = (int)(e % nk);
    double v = 0.0;
    if (r / 6 == c / 6) {
      v = H.Hkk[(size_t)(r / 6) * 36 + (r % 6) * 6 + (c % 6)];
      if (r == c) v = v + lam * v;  // H(i,i) += lambda * H(i,i)  (:1436-1437)
    }
    S[e] = v;
    if (c == 0) rhs[r] = H.g[r];
  }
}

// in-register inverse of a damped SPD block through its Cholesky factor (every lane computes it redundantly)
template <int DIM>
__device__ __forceinline__ void damped_inverse(const double* __restrict__ D, double lam, double* inv /*[DIM*DIM]*/) {
  double L[DIM * DIM];
#pragma unroll
  for (int i = 0; i < DIM; i++)
#pragma unroll
    for (int j = 0; j < DIM; j++) {
      double v = D[i * DIM + j];
      if (i == j) v = v + lam * v;
      L[i * DIM + j] = v;
    }
#pragma unroll
  for (int j = 0; j < DIM; j++) {
    double d = L[j * DIM + j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[j * DIM + k] * L[j * DIM + k];
    d = sqrt(d);
    L[j * DIM + j] = d;
    const double rd = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < DIM; i++) {
      double s = L[i * DIM + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i * DIM + k] * L[j * DIM + k];
      L[i * DIM + j] = s * rd;
    }
  }
  // Linv (lower) by forward substitution, then inv = Linv^T Linv
  double Li[DIM * DIM];
#pragma unroll
  for (int c = 0; c < DIM; c++) {
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      if (i < c) { Li[i * DIM + c] = 0.0; continue; }
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = c; k < i; k++) s -= L[i * DIM + k] * Li[k * DIM + c];
      Li[i * DIM + c] = s / L[i * DIM + i];
    }
  }
#pragma unroll
  for (int i = 0; i < DIM; i++)
#pragma unroll
    for (int j = 0; j < DIM; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = (i > j ? i : j); k < DIM; k++) s += Li[k * DIM + i] * Li[k * DIM + j];
      inv[i * DIM + j] = s;
    }
}

// One warp per landmark: Y_a = D'^-1 W_a for each of its observations from local KFs, z = D'^-1 g_j,
// rhs -= W_a^T z, S -= W_a^T Y_b.  obs_off: CSR offsets of the (landmark-sorted) observation list.
template <int DIM>
__global__ void __launch_bounds__(128)
lba_eliminate_kernel(const LbaWin* __restrict__ wins) {
  const LbaWin& wn = wins[blockIdx.y];
  const int n_lm = DIM == 3 ? wn.pb.Npt : wn.pb.Nls;
  const int32_t* __restrict__ obs_off = DIM == 3 ? wn.pt_off : wn.ls_off;
  const int32_t* __restrict__ obs_kf = DIM == 3 ? wn.pb.pt_obs_kf : wn.pb.ls_obs_kf;
  const double* __restrict__ Dall = DIM == 3 ? wn.H.Hpp : wn.H.Hll;
  const double* __restrict__ Wall = DIM == 3 ? wn.H.Hpk : wn.H.Hlk;
  const int nk = 6 * wn.pb.Nkf;
  const double* __restrict__ gj_all = wn.H.g + nk + (DIM == 3 ? 0 : 3 * (size_t)wn.pb.Npt);
  const double* __restrict__ lambda_eff = wn.lambda_eff;
  double* __restrict__ S = wn.w.S;
  double* __restrict__ rhs = wn.DX;
  double* __restrict__ Y = DIM == 3 ? wn.w.Ypt : wn.w.Yls;   // [n_obs][DIM*6]
  double* __restrict__ Z = DIM == 3 ? wn.w.Zpt : wn.w.Zls;   // [n_lm][DIM]
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n_lm) return;
  double inv[DIM * DIM], z[DIM];
  damped_inverse<DIM>(Dall + (size_t)j * DIM * DIM, *lambda_eff, inv);
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) s += inv[i * DIM + k] * gj_all[(size_t)j * DIM + k];
    z[i] = s;
  }
  if (lane < DIM) Z[(size_t)j * DIM + lane] = z[lane < DIM ? lane : 0];
  const int o0 = obs_off[j], o1 = obs_off[j + 1];
  for (int a = o0; a < o1; a++) {
    if (obs_kf[a] < 0) continue;
    const double* W = Wall + (size_t)a * DIM * 6;
    for (int e = lane; e < DIM * 6; e += 32) {  // Y_a[i][c]
      const int i = e / 6, c = e % 6;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; k++) s += inv[i * DIM + k] * W[k * 6 + c];
      Y[(size_t)a * DIM * 6 + e] = s;
    }
    if (lane < 6) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; i++) s += W[i * 6 + lane] * z[i];
      atomicAdd(&rhs[6 * obs_kf[a] + lane], -s);
    }
  }
  __syncwarp();
  for (int a = o0; a < o1; a++) {
    const int ka = obs_kf[a];
    if (ka < 0) continue;
    const double* W = Wall + (size_t)a * DIM * 6;
    for (int b = o0; b < o1; b++) {
      const int kb = obs_kf[b];
      if (kb < 0) continue;
      const double* Yb = Y + (size_t)b * DIM * 6;
      for (int e = lane; e < 36; e += 32) {
        const int r = e / 6, c = e % 6;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; i++) s += W[i * 6 + r] * Yb[i * 6 + c];
        atomicAdd(&S[(size_t)(6 * ka + r) * nk + 6 * kb + c], -s);
      }
    }
  }
}

// dense LDL^T (no pivoting: S is SPD after damping) + solve by one CTA.  n = 6 Nkf is a few tens: the matrix is staged in
// shared memory when it fits (n <= 96), else factorised in place in global memory.
__device__ __forceinline__ void ldlt_solve_cta(int n, double* __restrict__ S, double* __restrict__ x) {
  const int t = threadIdx.x;
  for (int k = 0; k < n; k++) {
    const double d = S[(size_t)k * n + k];
    __syncthreads();
    for (int i = k + 1 + t; i < n; i += 256) S[(size_t)i * n + k] = S[(size_t)i * n + k] / d;   // l_ik
    __syncthreads();
    const int m = n - k - 1;
    for (int e = t; e < m * m; e += 256) {
      const int i = k + 1 + e / m, jx = k + 1 + e % m;
      if (jx <= i) S[(size_t)i * n + jx] -= S[(size_t)i * n + k] * d * S[(size_t)jx * n + k];
    }
    __syncthreads();
  }
  // L y = b (unit lower), D z = y, L^T x = z — sequential over rows, parallel over the update of the remaining entries
  for (int k = 0; k < n; k++) {
    const double xk = x[k];
    __syncthreads();
    for (int i = k + 1 + t; i < n; i += 256) x[i] -= S[(size_t)i * n + k] * xk;
    __syncthreads();
  }
  for (int i = t; i < n; i += 256) x[i] = x[i] / S[(size_t)i * n + i];
  __syncthreads();
  for (int k = n - 1; k >= 0; k--) {
    const double xk = x[k];
    __syncthreads();
    for (int i = t; i < k; i += 256) x[i] -= S[(size_t)k * n + i] * xk;
    __syncthreads();
  }
}

constexpr int LBA_SOLVE_SMEM_N = 96;

__global__ void __launch_bounds__(256) lba_dense_solve_kernel(const LbaWin* __restrict__ wins, int smem_ok) {
  extern __shared__ double s_mat[];  // [n*n + n] when n <= LBA_SOLVE_SMEM_N
  const int n = 6 * wins[blockIdx.x].pb.Nkf;
  double* __restrict__ S = wins[blockIdx.x].w.S;
  double* __restrict__ x = wins[blockIdx.x].DX;
  if (n == 0) return;
  if (smem_ok) {   // the launch sized shared memory for the largest window: (n * n + n) doubles
    double* sx = s_mat + (size_t)n * n;
    for (int e = threadIdx.x; e < n * n; e += 256) s_mat[e] = S[e];
    for (int e = threadIdx.x; e < n; e += 256) sx[e] = x[e];
    __syncthreads();
    ldlt_solve_cta(n, s_mat, sx);
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += 256) x[e] = sx[e];
  } else {
    ldlt_solve_cta(n, S, x);
  }
}

// dx_j = z_j - sum_a Y_a dx_k(a)
template <int DIM>
__global__ void lba_backsub_kernel(const LbaWin* __restrict__ wins) {
  const LbaWin& wn = wins[blockIdx.y];
  const int n_lm = DIM == 3 ? wn.pb.Npt : wn.pb.Nls;
  const int32_t* __restrict__ obs_off = DIM == 3 ? wn.pt_off : wn.ls_off;
  const int32_t* __restrict__ obs_kf = DIM == 3 ? wn.pb.pt_obs_kf : wn.pb.ls_obs_kf;
  const double* __restrict__ Y = DIM == 3 ? wn.w.Ypt : wn.w.Yls;
  const double* __restrict__ Z = DIM == 3 ? wn.w.Zpt : wn.w.Zls;
  const double* __restrict__ dxk = wn.DX;
  double* __restrict__ dx = wn.DX + 6 * wn.pb.Nkf + (DIM == 3 ? 0 : 3 * (size_t)wn.pb.Npt);   // [n_lm][DIM]
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_lm * DIM) return;
  const int j = e / DIM, i = e % DIM;
  double v = Z[e];
  for (int a = obs_off[j]; a < obs_off[j + 1]; a++) {
    const int k = obs_kf[a];
    if (k < 0) continue;
#pragma unroll
    for (int c = 0; c < 6; c++) v -= Y[(size_t)a * DIM * 6 + i * 6 + c] * dxk[6 * k + c];
  }
  dx[e] = v;
}

// every window of the batch: lambda, Schur complement, dense solve, back substitution.  max_*: the largest window's sizes
int launch_lba_solve(gfpl_ctx* ctx, const LbaWin* wins, int W, int max_kf, int max_pt, int max_ls, double lambda, int scale_by_hmax) {
  const int nk = 6 * max_kf;
  lba_lambda_kernel<<<W, 256, 0, ctx->stream>>>(wins, lambda, scale_by_hmax);
  GFPL_LAUNCH_OK(ctx);
  if (nk > 0) {
    const int blocks = (int)(((size_t)nk * nk + 255) / 256);
    lba_schur_init_kernel<<<dim3(blocks < 1024 ? blocks : 1024, W), 256, 0, ctx->stream>>>(wins);
    GFPL_LAUNCH_OK(ctx);
  }
  if (max_pt > 0) {
    lba_eliminate_kernel<3><<<dim3((max_pt + 3) / 4, W), 128, 0, ctx->stream>>>(wins);
    GFPL_LAUNCH_OK(ctx);
  }
  if (max_ls > 0) {
    lba_eliminate_kernel<6><<<dim3((max_ls + 3) / 4, W), 128, 0, ctx->stream>>>(wins);
    GFPL_LAUNCH_OK(ctx);
  }
  if (nk > 0) {
    const size_t smem = nk <= LBA_SOLVE_SMEM_N ? ((size_t)nk * nk + nk) * sizeof(double) : 0;
    if (smem > 48 * 1024)
      GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(lba_dense_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lba_dense_solve_kernel<<<W, 256, smem, ctx->stream>>>(wins, nk <= LBA_SOLVE_SMEM_N ? 1 : 0);
    GFPL_LAUNCH_OK(ctx);
  }
  if (max_pt > 0) {
    lba_backsub_kernel<3><<<dim3((max_pt * 3 + 127) / 128, W), 128, 0, ctx->stream>>>(wins);
    GFPL_LAUNCH_OK(ctx);
  }
  if (max_ls > 0) {
    lba_backsub_kernel<6><<<dim3((max_ls * 6 + 127) / 128, W), 128, 0, ctx->stream>>>(wins);
    GFPL_LAUNCH_OK(ctx);
  }
  return GFPL_OK;
}

}  // namespace gfpl
