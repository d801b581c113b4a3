// TEST INFRASTRUCTURE — CPU oracle for the FastAPDGICP hot path. Not part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Small fixed-size fp64 linear algebra standing in for the Eigen calls on the reference path
// (Eigen is a third-party dependency that is absent from /root/reference and from this image):
//   JacobiSVD<Matrix3d>            fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:337
//   Matrix4d::inverse (3x3 block)  fast_apdgicp_impl.hpp:191
//   Matrix3d::inverse              fast_apdgicp_impl.hpp:332-334
//   LDLT<Matrix<double,6,6>>       fast_apdgicp/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:112,137
// The algorithms are the published ones (cyclic Jacobi for a symmetric 3x3; adjugate inverse;
// LDL^T with diagonal pivoting and Eigen's zero-pivot rule in solve).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

namespace apd_oracle {

struct Mat3 {
  double m[3][3];
  double& operator()(int r, int c) { return m[r][c]; }
  double operator()(int r, int c) const { return m[r][c]; }
  static Mat3 zero() {
    Mat3 z;
    for (auto& r : z.m)
      for (auto& v : r) v = 0.0;
    return z;
  }
  static Mat3 identity() {
    Mat3 z = zero();
    z.m[0][0] = z.m[1][1] = z.m[2][2] = 1.0;
    return z;
  }
};

inline Mat3 operator*(const Mat3& a, const Mat3& b) {
  Mat3 c = Mat3::zero();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) c.m[i][j] += a.m[i][k] * b.m[k][j];
  return c;
}
inline Mat3 operator+(const Mat3& a, const Mat3& b) {
  Mat3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][j] + b.m[i][j];
  return c;
}
inline Mat3 transpose(const Mat3& a) {
  Mat3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = a.m[j][i];
  return c;
}

// General 3x3 inverse by adjugate / determinant.
inline Mat3 inverse(const Mat3& a) {
  Mat3 c;
  c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
  c.m[0][1] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
  c.m[0][2] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
  c.m[1][0] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
  c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
  c.m[1][2] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
  c.m[2][0] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  c.m[2][1] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
  c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
  const double det = a.m[0][0] * c.m[0][0] + a.m[0][1] * c.m[1][0] + a.m[0][2] * c.m[2][0];
  const double inv = 1.0 / det;
  for (auto& r : c.m)
    for (auto& v : r) v *= inv;
  return c;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations; eigenvalues sorted DESCENDING
// (the order JacobiSVD returns singular values), columns of V the matching unit eigenvectors.
// For a symmetric PSD matrix this is its SVD with U == V.
inline void sym_eig3(const Mat3& a_in, double w[3], Mat3& V) {
  Mat3 a = a_in;
  V = Mat3::identity();
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = a.m[0][1] * a.m[0][1] + a.m[0][2] * a.m[0][2] + a.m[1][2] * a.m[1][2];
    const double diag = a.m[0][0] * a.m[0][0] + a.m[1][1] * a.m[1][1] + a.m[2][2] * a.m[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++) {
      for (int q = p + 1; q < 3; q++) {
        const double apq = a.m[p][q];
        if (apq == 0.0) continue;
        const double theta = (a.m[q][q] - a.m[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0);
        const double s = t * c;
        const int r = 3 - p - q;  // the untouched index
        const double app = a.m[p][p], aqq = a.m[q][q];
        const double arp = a.m[r][p], arq = a.m[r][q];
        a.m[p][p] = app - t * apq;
        a.m[q][q] = aqq + t * apq;
        a.m[p][q] = a.m[q][p] = 0.0;
        a.m[r][p] = a.m[p][r] = c * arp - s * arq;
        a.m[r][q] = a.m[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; k++) {
          const double vkp = V.m[k][p], vkq = V.m[k][q];
          V.m[k][p] = c * vkp - s * vkq;
          V.m[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  w[0] = a.m[0][0];
  w[1] = a.m[1][1];
  w[2] = a.m[2][2];
  // sort descending (3-element network), swapping columns of V alongside
  auto swap_cols = [&](int i, int j) {
    std::swap(w[i], w[j]);
    for (int k = 0; k < 3; k++) std::swap(V.m[k][i], V.m[k][j]);
  };
  if (w[0] < w[1]) swap_cols(0, 1);
  if (w[1] < w[2]) swap_cols(1, 2);
  if (w[0] < w[1]) swap_cols(0, 1);
}

// V diag(d) V^T
inline Mat3 recompose(const Mat3& V, const double d[3]) {
  Mat3 c = Mat3::zero();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) c.m[i][j] += V.m[i][k] * d[k] * V.m[j][k];
  return c;
}

// LDL^T of a symmetric 6x6 with diagonal pivoting, then solve A x = rhs.
// Pivot = largest remaining |diagonal|; in the solve a pivot with |D_i| <= tolerance contributes 0
// (the rule Eigen's LDLT::solve applies), tolerance = max(max|D| * eps, 1/highest).
inline void ldlt6_solve(const double A_in[6][6], const double rhs[6], double x[6]) {
  double A[6][6];
  int perm[6];
  for (int i = 0; i < 6; i++) {
    perm[i] = i;
    for (int j = 0; j < 6; j++) A[i][j] = A_in[i][j];
  }
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < 6; i++)
      if (std::fabs(A[i][i]) > best) {
        best = std::fabs(A[i][i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < 6; j++) std::swap(A[k][j], A[piv][j]);
      for (int i = 0; i < 6; i++) std::swap(A[i][k], A[i][piv]);
      std::swap(perm[k], perm[piv]);
    }
    const double d = A[k][k];
    if (d == 0.0) continue;  // remaining Schur complement column left as is (zero pivot)
    for (int i = k + 1; i < 6; i++) A[i][k] /= d;  // L(i,k)
    for (int i = k + 1; i < 6; i++)
      for (int j = k + 1; j <= i; j++) {
        A[i][j] -= A[i][k] * d * A[j][k];
        A[j][i] = A[i][j];
      }
  }
  double maxd = 0.0;
  for (int i = 0; i < 6; i++) maxd = std::max(maxd, std::fabs(A[i][i]));
  const double tol = std::max(maxd * std::numeric_limits<double>::epsilon(), 1.0 / std::numeric_limits<double>::max());
  double y[6];
  for (int i = 0; i < 6; i++) y[i] = rhs[perm[i]];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i][j] * y[j];  // L y = P b
  for (int i = 0; i < 6; i++) y[i] = (std::fabs(A[i][i]) > tol) ? y[i] / A[i][i] : 0.0;
  for (int i = 5; i >= 0; i--)
    for (int j = i + 1; j < 6; j++) y[i] -= A[j][i] * y[j];  // L^T z = y
  for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
}

}  // namespace apd_oracle
