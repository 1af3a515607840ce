// gfpl_math.cuh — small fixed-size fp64 algebra for the sm_100a kernels.
//
// Everything here is register-resident device code (fully unrolled loops, no local-memory arrays with
// dynamic indexing where it can be avoided).  The library is compiled with -fmad=false so that every
// elementwise fp64 expression rounds exactly like IEEE scalar code does; only the order of parallel
// reductions can differ from a sequential CPU evaluation.  Sums inside one product run in ascending
// index order.  Reference formulas: auxiliar.cpp:154-238 (SE(3)), pinholeStereoCamera.cpp:133-170.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/gfpl_c.h"

namespace gfpl {

#define GFPL_DEV __device__ __forceinline__

template <int N, int K, int M>
GFPL_DEV void matmul(const double* A, const double* B, double* C) {
  double tmp[N * M];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < M; j++) {
      double s = A[i * K + 0] * B[0 * M + j];
#pragma unroll
      for (int k = 1; k < K; k++) s = s + A[i * K + k] * B[k * M + j];
      tmp[i * M + j] = s;
    }
#pragma unroll
  for (int i = 0; i < N * M; i++) C[i] = tmp[i];
}

GFPL_DEV void eye4(double* T) {
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// y = R*x + t of a row-major 4x4
GFPL_DEV void se3_apply(const double* T, const double* x, double* y) {
  double r[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double s = T[i * 4 + 0] * x[0];
    s = s + T[i * 4 + 1] * x[1];
    s = s + T[i * 4 + 2] * x[2];
    r[i] = s + T[i * 4 + 3];
  }
  y[0] = r[0]; y[1] = r[1]; y[2] = r[2];
}

GFPL_DEV void skew(const double* v, double* S) {
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

// auxiliar.cpp:154-163
GFPL_DEV void inverse_se3(const double* T, double* Tinv) {
  double out[16];
  eye4(out);
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) out[i * 4 + j] = T[j * 4 + i];
    double s = (-T[0 * 4 + i]) * T[0 * 4 + 3];
    s = s + (-T[1 * 4 + i]) * T[1 * 4 + 3];
    s = s + (-T[2 * 4 + i]) * T[2 * 4 + 3];
    out[i * 4 + 3] = s;
  }
#pragma unroll
  for (int i = 0; i < 16; i++) Tinv[i] = out[i];
}

// auxiliar.cpp:165-182, x = (t; w)
GFPL_DEV void expmap_se3(const double* x, double* T) {
  double w[3] = {x[3], x[4], x[5]}, t[3] = {x[0], x[1], x[2]};
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (!(theta < 0.000001)) {
    double s[9], ss[9], V[9], sk[9];
    skew(w, sk);
#pragma unroll
    for (int i = 0; i < 9; i++) s[i] = sk[i] / theta;
    matmul<3, 3, 3>(s, s, ss);
    double sn = sin(theta), cs = cos(theta);
    double omc = 1.0 - cs, tms = theta - sn;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = (I + s[i] * sn) + ss[i] * omc;
      V[i] = (I + (s[i] * omc) / theta) + (ss[i] * tms) / theta;
    }
    double tn[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double a = V[i * 3 + 0] * t[0];
      a = a + V[i * 3 + 1] * t[1];
      a = a + V[i * 3 + 2] * t[2];
      tn[i] = a;
    }
    t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
  }
  eye4(T);
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
    T[i * 4 + 3] = t[i];
  }
}

// auxiliar.cpp:475-477
GFPL_DEV bool is_finite16(const double* x) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 16; i++) ok = ok && ((x[i] - x[i]) == (x[i] - x[i]));
  return ok;
}

// general 4x4 inverse by cofactors (Matrix4d::inverse())
GFPL_DEV void inverse4(const double* m, double* invOut) {
  double inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  double id = 1.0 / det;
#pragma unroll
  for (int i = 0; i < 16; i++) invOut[i] = inv[i] * id;
}

// ---- camera ------------------------------------------------------------------------------------
GFPL_DEV void back_projection(const gfpl_camera& c, double u, double v, double disp, double* P) {
  double bd = c.b / disp;
  P[0] = bd * (u - c.cx);
  P[1] = bd * (v - c.cy);
  P[2] = bd * c.fx;
}
GFPL_DEV void projection(const gfpl_camera& c, const double* P, double* uv) {
  uv[0] = c.cx + c.fx * P[0] / P[2];
  uv[1] = c.cy + c.fy * P[1] / P[2];
}
GFPL_DEV double get_disparity(const gfpl_camera& c, double pZ) { return c.fx * c.b / pZ; }

// the 6-vector shared by optimizeFunctions, getPoseInfoOnLine and getPoseInfoPoint
// (stereoFrameHandler.cpp:2150-2155, 1383-1388, 1438-1443)
GFPL_DEV void jac6(double fgz2, double gx, double gy, double gz, double lx, double ly, double* J) {
  J[0] = +fgz2 * lx * gz;
  J[1] = +fgz2 * ly * gz;
  J[2] = -fgz2 * (gx * lx + gy * ly);
  J[3] = -fgz2 * (gx * gy * lx + gy * gy * ly + gz * gz * ly);
  J[4] = +fgz2 * (gx * gx * lx + gz * gz * lx + gx * gy * ly);
  J[5] = +fgz2 * (gx * gz * ly - gy * gz * lx);
}

// ---- symmetric eigenvalues: cyclic Jacobi, fixed sweep order, ascending output -------------------
// stands in for SelfAdjointEigenSolver (stereoFrame.cpp:994-997, stereoFrameHandler.cpp:1921) and the
// EigenSolver real parts of a symmetric matrix (linespec.cpp:53-66).
template <int N>
__device__ __noinline__ void sym_eigenvalues(const double* Ain, double* ev) {
  double A[N * N];
  for (int i = 0; i < N * N; i++) A[i] = Ain[i];
  for (int sweep = 0; sweep < 16; sweep++) {
    double off = 0.0;
    for (int p = 0; p < N; p++)
      for (int q = p + 1; q < N; q++) off = off + A[p * N + q] * A[p * N + q];
    if (off == 0.0) break;
    for (int p = 0; p < N; p++)
      for (int q = p + 1; q < N; q++) {
        double apq = A[p * N + q];
        if (apq == 0.0) continue;
        double app = A[p * N + p], aqq = A[q * N + q];
        double tau = (aqq - app) / (2.0 * apq);
        double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < N; k++) {
          double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; k++) {
          double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        A[p * N + q] = 0.0;
        A[q * N + p] = 0.0;
      }
  }
  for (int i = 0; i < N; i++) ev[i] = A[i * N + i];
  for (int i = 1; i < N; i++) {  // insertion sort ascending
    double v = ev[i];
    int j = i - 1;
    while (j >= 0 && ev[j] > v) { ev[j + 1] = ev[j]; j--; }
    ev[j + 1] = v;
  }
}

// same arithmetic, p/q/k loops unrolled so that A stays in registers (3x3 endpoint covariances in the accept loop)
template <int N>
__device__ __forceinline__ void sym_eigenvalues_unrolled(const double* Ain, double* ev) {
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
        if (apq != 0.0) {
          const double app = A[p * N + p], aqq = A[q * N + q];
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
#pragma unroll
          for (int k = 0; k < N; k++) {
            const double akp = A[k * N + p], akq = A[k * N + q];
            A[k * N + p] = c * akp - s * akq;
            A[k * N + q] = s * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < N; k++) {
            const double apk = A[p * N + k], aqk = A[q * N + k];
            A[p * N + k] = c * apk - s * aqk;
            A[q * N + k] = s * apk + c * aqk;
          }
          A[p * N + q] = 0.0;
          A[q * N + p] = 0.0;
        }
      }
  }
#pragma unroll
  for (int i = 0; i < N; i++) ev[i] = A[i * N + i];
}

// ---- 6x6 pivoted LDLT solve (LDLT<Matrix6d>::solve, stereoFrameHandler.cpp:2045-2046) --------------
static __device__ __noinline__ void ldlt6_solve(const double* Hin, const double* g, double* x) {
  const int n = 6;
  double A[36];
  int perm[6];
  for (int i = 0; i < 36; i++) A[i] = Hin[i];
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++) {
      double v = fabs(A[i * n + i]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {
      for (int j = 0; j < n; j++) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
      for (int i = 0; i < n; i++) { double t = A[i * n + k]; A[i * n + k] = A[i * n + p]; A[i * n + p] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    double temp[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; j++) temp[j] = A[j * n + j] * A[k * n + j];
    if (k > 0) {
      double dot = A[k * n + 0] * temp[0];
      for (int j = 1; j < k; j++) dot = dot + A[k * n + j] * temp[j];
      A[k * n + k] = A[k * n + k] - dot;
    }
    double d = A[k * n + k];
    for (int i = k + 1; i < n; i++) {
      double s = A[i * n + k];
      if (k > 0) {
        double dot = A[i * n + 0] * temp[0];
        for (int j = 1; j < k; j++) dot = dot + A[i * n + j] * temp[j];
        s = s - dot;
      }
      A[i * n + k] = (fabs(d) > 0.0) ? s / d : s;
    }
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = g[perm[i]];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] = y[i] - A[i * n + j] * y[j];
  const double tol = 1.0 / 1.7976931348623157e308;
  for (int i = 0; i < n; i++) y[i] = (fabs(A[i * n + i]) > tol) ? y[i] / A[i * n + i] : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] = y[i] - A[j * n + i] * y[j];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

// ---- 6x6 inverse by partial-pivot LU (Matrix6d::inverse(), stereoFrameHandler.cpp:2054) -----------
static __device__ __noinline__ void inverse6(const double* Hin, double* inv) {
  const int n = 6;
  double A[36], B[36];
  for (int i = 0; i < 36; i++) { A[i] = Hin[i]; B[i] = (i % 7 == 0) ? 1.0 : 0.0; }
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++) {
      double v = fabs(A[i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k)
      for (int j = 0; j < n; j++) {
        double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t;
        t = B[k * n + j]; B[k * n + j] = B[p * n + j]; B[p * n + j] = t;
      }
    double piv = A[k * n + k];
    for (int i = k + 1; i < n; i++) {
      double f = A[i * n + k] / piv;
      A[i * n + k] = f;
      for (int j = k + 1; j < n; j++) A[i * n + j] = A[i * n + j] - f * A[k * n + j];
      for (int j = 0; j < n; j++) B[i * n + j] = B[i * n + j] - f * B[k * n + j];
    }
  }
  for (int c = 0; c < n; c++)
    for (int i = n - 1; i >= 0; i--) {
      double s = B[i * n + c];
      for (int j = i + 1; j < n; j++) s = s - A[i * n + j] * B[j * n + c];
      B[i * n + c] = s / A[i * n + i];
    }
  for (int i = 0; i < 36; i++) inv[i] = B[i];
}

// lower-triangle packed index, i >= j
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// logdet (linespec.h:43-56) on a lower-triangle-packed symmetric 6x6 (21 values): Eigen's LLT reads only
// the lower triangle.  Unblocked left-looking kernel incl. its early return on a non-positive pivot
// (SURVEY.md A.4 D7); fully unrolled so L stays in registers.
GFPL_DEV double logdet6_packed(const double* Min) {
  double L[21];
#pragma unroll
  for (int i = 0; i < 21; i++) L[i] = Min[i];
  bool alive = true;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double x = L[tri(k, k)];
    if (k > 0) {
      double sn = L[tri(k, 0)] * L[tri(k, 0)];
#pragma unroll
      for (int j = 1; j < k; j++) sn = sn + L[tri(k, j)] * L[tri(k, j)];
      x = x - sn;
    }
    if (alive && x <= 0.0) alive = false;  // `return k`: the rest of the matrix is left untouched
    if (alive) {
      x = sqrt(x);
      L[tri(k, k)] = x;
#pragma unroll
      for (int i = k + 1; i < 6; i++) {
        double s = L[tri(i, k)];
        if (k > 0) {
          double d = L[tri(i, 0)] * L[tri(k, 0)];
#pragma unroll
          for (int j = 1; j < k; j++) d = d + L[tri(i, j)] * L[tri(k, j)];
          s = s - d;
        }
        L[tri(i, k)] = s / x;
      }
    }
  }
  double sum = log(L[tri(0, 0)]);
#pragma unroll
  for (int i = 1; i < 6; i++) sum = sum + log(L[tri(i, i)]);
  return 2.0 * sum;
}

// ---- warp helpers --------------------------------------------------------------------------------
GFPL_DEV double shfl_xor_f64(unsigned mask, double v, int lane_mask, int width = 32) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(mask, lo, lane_mask, width);
  hi = __shfl_xor_sync(mask, hi, lane_mask, width);
  return __hiloint2double(hi, lo);
}
GFPL_DEV double shfl_f64(unsigned mask, double v, int src, int width = 32) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(mask, lo, src, width);
  hi = __shfl_sync(mask, hi, src, width);
  return __hiloint2double(hi, lo);
}


// ---- C1: estimateStereoUncertainty (stereoFrame.cpp:1448-1484) with getCovMat2D_3D / getJacob2D_3D (:1375-1392, 1434-1446) ----
__device__ __forceinline__ void cov_mat_2d_3d(const gfpl_camera& cam, double u, double v, double u_std, double d,
                                              double d_std, double* cov) {
  double cov2D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // D5: off-diagonals are zero
  cov2D[0] = u_std * u_std; cov2D[4] = u_std * u_std; cov2D[8] = d_std * d_std;
  const double b = cam.b, d_2 = d * d;
  double J[9], Jt[9], tmp[9];
  J[0] = b / d; J[3] = 0.0; J[6] = 0.0;
  J[1] = 0.0; J[4] = b / d; J[7] = 0.0;
  J[2] = -(u - cam.cx) * b / d_2;
  J[5] = -(v - cam.cy) * b / d_2;
  J[8] = -cam.fx * b / d_2;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Jt[j * 3 + i] = J[i * 3 + j];
  matmul<3, 3, 3>(J, cov2D, tmp);
  matmul<3, 3, 3>(tmp, Jt, cov);
}


// covSpt3D / covEpt3D of one stereo line (stereoFrame.cpp:1458-1482)
__device__ __forceinline__ void line_cov_from_feature(const gfpl_camera& cam, const gfpl_config& cfg, const gfpl_line_feature& l,
                                                      gfpl_line_cov& c) {
  const double ratio = fabs(l.le[0]) > 0.15 ? cfg.ratio_disp_std : cfg.ratio_disp_std_hor;
  const double sdisp_std = ratio * l.sdisp, edisp_std = ratio * l.edisp;
  cov_mat_2d_3d(cam, l.spl[0], l.spl[1], 1.0, l.sdisp, sdisp_std, c.covS);
  cov_mat_2d_3d(cam, l.epl[0], l.epl[1], 1.0, l.edisp, edisp_std, c.covE);
}

}  // namespace gfpl
