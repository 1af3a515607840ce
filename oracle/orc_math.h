// TEST INFRASTRUCTURE — CPU oracle of the GF-PL-SLAM tracking hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
// The product (gf-pl-slam_b200/) never includes, links or calls anything in oracle/.
//
// orc_math.h — dependency-free small dense algebra standing in for the Eigen3 pieces the reference
// uses on the hot path (Eigen is NOT vendored in /root/reference and not installed here; "tested
// with 3.2.92", reference README.md:87).  Every routine states which Eigen facility it restates.
// All matrices are row-major.  Sums run in ascending index order, the order of Eigen's scalar
// (non-vectorised) product/redux paths.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>

namespace orc {

// C(n x m) = A(n x k) * B(k x m)
template <int N, int K, int M>
inline void matmul(const double* A, const double* B, double* C) {
  double tmp[N * M];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < M; j++) {
      double s = A[i * K + 0] * B[0 * M + j];
      for (int k = 1; k < K; k++) s = s + A[i * K + k] * B[k * M + j];
      tmp[i * M + j] = s;
    }
  std::memcpy(C, tmp, sizeof(tmp));
}

template <int N, int M>
inline void transpose(const double* A, double* At) {
  double tmp[N * M];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < M; j++) tmp[j * N + i] = A[i * M + j];
  std::memcpy(At, tmp, sizeof(tmp));
}

inline void eye4(double* T) {
  for (int i = 0; i < 16; i++) T[i] = 0.0;
  T[0] = T[5] = T[10] = T[15] = 1.0;
}

// y = R*x + t for a row-major 4x4 (DT.block(0,0,3,3) * P + DT.col(3).head(3))
inline void se3_apply(const double* T, const double* x, double* y) {
  double r[3];
  for (int i = 0; i < 3; i++) {
    double s = T[i * 4 + 0] * x[0];
    s = s + T[i * 4 + 1] * x[1];
    s = s + T[i * 4 + 2] * x[2];
    r[i] = s + T[i * 4 + 3];
  }
  y[0] = r[0]; y[1] = r[1]; y[2] = r[2];
}

// auxiliar.cpp:248-262
inline void skew(const double* v, double* S) {
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

// auxiliar.cpp:154-163
inline void inverse_se3(const double* T, double* Tinv) {
  double R[9], t[3], out[16];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R[i * 3 + j] = T[i * 4 + j];
    t[i] = T[i * 4 + 3];
  }
  eye4(out);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) out[i * 4 + j] = R[j * 3 + i];
    double s = (-R[0 * 3 + i]) * t[0];
    s = s + (-R[1 * 3 + i]) * t[1];
    s = s + (-R[2 * 3 + i]) * t[2];
    out[i * 4 + 3] = s;
  }
  std::memcpy(Tinv, out, sizeof(out));
}

// auxiliar.cpp:165-182.  x = (t; w)
inline void expmap_se3(const double* x, double* T) {
  double w[3] = {x[3], x[4], x[5]}, t[3] = {x[0], x[1], x[2]};
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (!(theta < 0.000001)) {
    double s[9], ss[9], V[9], sk[9];
    skew(w, sk);
    for (int i = 0; i < 9; i++) s[i] = sk[i] / theta;
    matmul<3, 3, 3>(s, s, ss);
    double sn = std::sin(theta), cs = std::cos(theta);
    double omc = 1.0 - cs, tms = theta - sn;
    for (int i = 0; i < 9; i++) {
      double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = (I + s[i] * sn) + ss[i] * omc;
      V[i] = (I + (s[i] * omc) / theta) + (ss[i] * tms) / theta;
    }
    double tn[3];
    for (int i = 0; i < 3; i++) {
      double a = V[i * 3 + 0] * t[0];
      a = a + V[i * 3 + 1] * t[1];
      a = a + V[i * 3 + 2] * t[2];
      tn[i] = a;
    }
    t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
  }
  eye4(T);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
    T[i * 4 + 3] = t[i];
  }
}

// Matrix3d::inverse(): Eigen's fixed-size 3x3 path is cofactors / determinant.
inline void inverse3(const double* m, double* inv) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c10 = m[5] * m[6] - m[3] * m[8];
  double c20 = m[3] * m[7] - m[4] * m[6];
  double det = (m[0] * c00 + m[1] * c10) + m[2] * c20;
  double id = 1.0 / det;
  double out[9];
  out[0] = c00 * id;
  out[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  out[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  out[3] = c10 * id;
  out[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  out[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  out[6] = c20 * id;
  out[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  out[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  std::memcpy(inv, out, sizeof(out));
}

// auxiliar.cpp:184-214
inline void logmap_se3(const double* T, double* x) {
  double R[9], Vt[3] = {T[3], T[7], T[11]}, w[3] = {0, 0, 0};
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = T[i * 4 + j];
  double cosine = (((R[0] + R[4]) + R[8]) - 1.0) / 2.0;
  if (cosine > 1.0) cosine = 1.0;
  else if (cosine < -1.0) cosine = -1.0;
  double sine = std::sqrt(1.0 - cosine * cosine);
  if (sine > 1.0) sine = 1.0;
  else if (sine < -1.0) sine = -1.0;
  double theta = std::acos(cosine);
  if (theta > 0.000001) {
    double w_hat[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) w_hat[i * 3 + j] = (theta * (R[i * 3 + j] - R[j * 3 + i])) / (2.0 * sine);
    // skewcoords (auxiliar.cpp:264-270): (w_hat(2,1), w_hat(0,2), w_hat(1,0))
    w[0] = w_hat[7]; w[1] = w_hat[2]; w[2] = w_hat[3];
    double sk[9], s[9], ss[9];
    skew(w, sk);
    for (int i = 0; i < 9; i++) s[i] = sk[i] / theta;
    matmul<3, 3, 3>(s, s, ss);
    double omc = 1.0 - cosine, tms = theta - sine;
    for (int i = 0; i < 9; i++) {
      double I = (i % 4 == 0) ? 1.0 : 0.0;
      V[i] = (I + (s[i] * omc) / theta) + (ss[i] * tms) / theta;
    }
  }
  double Vi[9];
  inverse3(V, Vi);
  for (int i = 0; i < 3; i++) {
    double a = Vi[i * 3 + 0] * Vt[0];
    a = a + Vi[i * 3 + 1] * Vt[1];
    a = a + Vi[i * 3 + 2] * Vt[2];
    x[i] = a;
  }
  x[3] = w[0]; x[4] = w[1]; x[5] = w[2];
}

// auxiliar.cpp:216-223 (6x6 row-major)
inline void adjoint_se3(const double* T, double* Ad) {
  double R[9], t[3], sk[9], skR[9];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R[i * 3 + j] = T[i * 4 + j];
    t[i] = T[i * 4 + 3];
  }
  skew(t, sk);
  matmul<3, 3, 3>(sk, R, skR);
  for (int i = 0; i < 36; i++) Ad[i] = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Ad[i * 6 + j] = R[i * 3 + j];
      Ad[i * 6 + 3 + j] = skR[i * 3 + j];
      Ad[(i + 3) * 6 + 3 + j] = R[i * 3 + j];
    }
}

// auxiliar.cpp:233-238
inline void unccomp_se3(const double* T1, const double* covT1, const double* covTinc, double* covT2) {
  double Ad[36], AdT[36], tmp[36], tmp2[36];
  adjoint_se3(T1, Ad);
  transpose<6, 6>(Ad, AdT);
  matmul<6, 6, 6>(Ad, covTinc, tmp);
  matmul<6, 6, 6>(tmp, AdT, tmp2);
  for (int i = 0; i < 36; i++) covT2[i] = covT1[i] + tmp2[i];
}

// auxiliar.cpp:475-477
inline bool is_finite(const double* x, int n) {
  for (int i = 0; i < n; i++)
    if (!((x[i] - x[i]) == (x[i] - x[i]))) return false;
  return true;
}

// Matrix4d::inverse(): general 4x4 inverse by cofactor expansion (Eigen's fixed-size 4x4 path is a
// cofactor formula too; its exact association is not reproducible without the source, so 1e-12-level
// differences against a real Eigen build are expected and covered by the fp64 tolerance).
inline void inverse4(const double* m, double* invOut) {
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
  for (int i = 0; i < 16; i++) invOut[i] = inv[i] * id;
}

// LDLT<Matrix6d>(H).solve(g) (stereoFrameHandler.cpp:2045-2046): Eigen's robust Cholesky with
// symmetric diagonal pivoting (largest |diagonal| first).  Restated with the same pivot rule;
// zero pivots are skipped in the D^-1 step like Eigen's pseudo-inverse of D.
inline void ldlt6_solve(const double* Hin, const double* g, double* x) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Hin, sizeof(A));
  int perm[6];
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++) {
      double v = std::fabs(A[i * n + i]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {  // symmetric swap of rows/cols k and p
      for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[p * n + j]);
      for (int i = 0; i < n; i++) std::swap(A[i * n + k], A[i * n + p]);
      std::swap(perm[k], perm[p]);
    }
    // left-looking step as in Eigen's ldlt_inplace<Lower>::unblocked:
    //   temp_j = D_j * L_kj ; d_k = A_kk - (L_k,0:k . temp) ; L_ik = (A_ik - (L_i,0:k . temp)) / d_k
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
      A[i * n + k] = (std::fabs(d) > 0.0) ? s / d : s;
    }
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = g[perm[i]];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] = y[i] - A[i * n + j] * y[j];
  const double tol = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < n; i++) y[i] = (std::fabs(A[i * n + i]) > tol) ? y[i] / A[i * n + i] : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] = y[i] - A[j * n + i] * y[j];
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

// Matrix6d::inverse() (stereoFrameHandler.cpp:2054): Eigen uses PartialPivLU for fixed sizes > 4.
inline void inverse6(const double* Hin, double* inv) {
  const int n = 6;
  double A[36], B[36];
  std::memcpy(A, Hin, sizeof(A));
  for (int i = 0; i < 36; i++) B[i] = (i % 7 == 0) ? 1.0 : 0.0;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++) {
      double v = std::fabs(A[i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k)
      for (int j = 0; j < n; j++) { std::swap(A[k * n + j], A[p * n + j]); std::swap(B[k * n + j], B[p * n + j]); }
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
  std::memcpy(inv, B, sizeof(B));
}

// SelfAdjointEigenSolver<MatrixNd>::eigenvalues() (stereoFrame.cpp:994-997; stereoFrameHandler.cpp:1921)
// and the real parts EigenSolver returns for a symmetric matrix (linespec.cpp:53-66).
// Restated as cyclic Jacobi (fixed sweep order, fixed sweep count) + ascending sort: eigenvalues agree
// with any backward-stable symmetric solver to O(eps*||A||).
template <int N>
inline void sym_eigenvalues(const double* Ain, double* ev) {
  double A[N * N];
  std::memcpy(A, Ain, sizeof(A));
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
        double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < N; k++) {  // columns p,q
          double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; k++) {  // rows p,q
          double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        A[p * N + q] = 0.0;
        A[q * N + p] = 0.0;
      }
  }
  for (int i = 0; i < N; i++) ev[i] = A[i * N + i];
  std::sort(ev, ev + N);
}

// logdet (linespec.h:43-56): 2 * sum(log(diag(L))) with Eigen's LLT left in whatever state its
// unblocked kernel reached: on a non-positive pivot at step k the kernel returns and the
// remaining diagonal still holds the input values.
inline double logdet6(const double* Min) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Min, sizeof(A));
  for (int k = 0; k < n; k++) {
    double x = A[k * n + k];
    if (k > 0) {
      double sn = A[k * n + 0] * A[k * n + 0];
      for (int j = 1; j < k; j++) sn = sn + A[k * n + j] * A[k * n + j];
      x = x - sn;
    }
    if (x <= 0.0) break;  // Eigen: `if (x <= 0) return k;` — NaN falls through and propagates
    x = std::sqrt(x);
    A[k * n + k] = x;
    for (int i = k + 1; i < n; i++) {
      double s = A[i * n + k];
      if (k > 0) {
        double d = A[i * n + 0] * A[k * n + 0];
        for (int j = 1; j < k; j++) d = d + A[i * n + j] * A[k * n + j];
        s = s - d;
      }
      A[i * n + k] = s / x;
    }
  }
  double sum = std::log(A[0]);
  for (int i = 1; i < n; i++) sum = sum + std::log(A[i * n + i]);
  return 2.0 * sum;
}

}  // namespace orc
