// Device-side fixed-size fp64 linear algebra and the bit-exact fp32 helpers of the FastAPDGICP path.
// Replaces the Eigen calls of the reference (JacobiSVD / Matrix4d::inverse / LDLT / AngleAxis,
// fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:174-191,337 and
// fast_apdgicp/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:112,137) and so3_exp
// (fast_apdgicp/include/fast_gicp/so3/so3.hpp:59-78).
//
// The header also compiles as plain C++ (tests/host_harness.cpp builds it with g++
// -ffp-contract=off) so the search and the small solvers can be unit-tested without a GPU; that
// build is test infrastructure only, the product links the nvcc build and nothing else.
#pragma once
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define APD_HD __host__ __device__ __forceinline__
#define APD_HD_NOINLINE __host__ __device__ inline
#else
#define APD_HD inline
#define APD_HD_NOINLINE inline
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace apd {

// ---------- one rounding per operation (no FMA contraction) ----------
// On the device these are the _rn intrinsics, which the compiler never fuses; on the host the
// translation unit is built with -ffp-contract=off.
#if defined(__CUDA_ARCH__)
APD_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
APD_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
APD_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
APD_HD float fsqrt(float a) { return __fsqrt_rn(a); }
APD_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
APD_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
APD_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
APD_HD unsigned f2u(float f) { return __float_as_uint(f); }
APD_HD float u2f(unsigned u) { return __uint_as_float(u); }
#else
APD_HD float fmul(float a, float b) { return a * b; }
APD_HD float fadd(float a, float b) { return a + b; }
APD_HD float fsub(float a, float b) { return a - b; }
APD_HD float fsqrt(float a) { return std::sqrt(a); }
APD_HD double dmul(double a, double b) { return a * b; }
APD_HD double dadd(double a, double b) { return a + b; }
APD_HD double dsub(double a, double b) { return a - b; }
APD_HD unsigned f2u(float f) { union { float f; unsigned u; } c; c.f = f; return c.u; }
APD_HD float u2f(unsigned u) { union { float f; unsigned u; } c; c.u = u; return c.f; }
#endif

// FLANN L2_Simple<float>: ((dx*dx + dy*dy) + dz*dz); the reference is built without FMA
// (fast_apdgicp/CMakeLists.txt:11-13), so the device must not contract either.
APD_HD float sqdist_rn(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = fsub(qx, px);
  const float dy = fsub(qy, py);
  const float dz = fsub(qz, pz);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// float isometry * float point: ((R0*x + R1*y) + R2*z) + t  (SURVEY.md §8c convention for
// fast_apdgicp_impl.hpp:149 and for pcl::transformPointCloud at lsq_registration_impl.hpp:79)
APD_HD float xform_row_rn(float r0, float r1, float r2, float t, float x, float y, float z) {
  return fadd(fadd(fadd(fmul(r0, x), fmul(r1, y)), fmul(r2, z)), t);
}

// correctly rounded float arctangent: double evaluation rounded once (the definition the oracle uses
// for the atan2f calls at fast_apdgicp_impl.hpp:168,172-173)
APD_HD double atan2_f32(float y, float x) { return (double)(float)atan2((double)y, (double)x); }

// ---------- symmetric 3x3 in 6 doubles: xx xy xz yy yz zz ----------
struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};

APD_HD Sym3 operator+(const Sym3& a, const Sym3& b) { return Sym3{a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz}; }

// R * S * R^T for a general 3x3 R (row-major) and symmetric S
APD_HD Sym3 rsrt(const double* R, const Sym3& S) {
  double A[9];  // A = R * S
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double r0 = R[i * 3 + 0], r1 = R[i * 3 + 1], r2 = R[i * 3 + 2];
    A[i * 3 + 0] = r0 * S.xx + r1 * S.xy + r2 * S.xz;
    A[i * 3 + 1] = r0 * S.xy + r1 * S.yy + r2 * S.yz;
    A[i * 3 + 2] = r0 * S.xz + r1 * S.yz + r2 * S.zz;
  }
  Sym3 o;
  o.xx = A[0] * R[0] + A[1] * R[1] + A[2] * R[2];
  o.xy = A[0] * R[3] + A[1] * R[4] + A[2] * R[5];
  o.xz = A[0] * R[6] + A[1] * R[7] + A[2] * R[8];
  o.yy = A[3] * R[3] + A[4] * R[4] + A[5] * R[5];
  o.yz = A[3] * R[6] + A[4] * R[7] + A[5] * R[8];
  o.zz = A[6] * R[6] + A[7] * R[7] + A[8] * R[8];
  return o;
}

// inverse of a symmetric 3x3 by adjugate / determinant (Matrix4d::inverse of the block-diagonal
// 4x4 at fast_apdgicp_impl.hpp:189-192 reduces to this)
APD_HD Sym3 inverse(const Sym3& a) {
  Sym3 c;
  c.xx = a.yy * a.zz - a.yz * a.yz;
  c.xy = a.xz * a.yz - a.xy * a.zz;
  c.xz = a.xy * a.yz - a.xz * a.yy;
  c.yy = a.xx * a.zz - a.xz * a.xz;
  c.yz = a.xy * a.xz - a.xx * a.yz;
  c.zz = a.xx * a.yy - a.xy * a.xy;
  const double det = a.xx * c.xx + a.xy * c.xy + a.xz * c.xz;
  const double inv = 1.0 / det;
  c.xx *= inv; c.xy *= inv; c.xz *= inv; c.yy *= inv; c.yz *= inv; c.zz *= inv;
  return c;
}

// One Jacobi rotation annihilating a(p,q); r is the third index. Scalars are passed by reference so
// the whole decomposition stays in registers. Every operation rounds once (dmul/dadd/dsub), in the
// same order as the CPU oracle (oracle/linalg.hpp sym_eig3), so both follow the same rotation
// sequence bit for bit even on ill-conditioned neighbourhoods.
// The rotation angle of one Jacobi step: the divisions and square roots expand to a few hundred SASS
// instructions, so the three rotations of a sweep share ONE out-of-line copy (instruction-cache footprint of
// the kNN + covariance kernel); the cheap application of the rotation stays inline.
struct JacobiCS {
  double c, s, tapq;
};
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
static inline
#endif
JacobiCS jacobi_cs(double app, double aqq, double apq) {
  const double theta = dsub(aqq, app) / dmul(2.0, apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / dadd(fabs(theta), sqrt(dadd(dmul(theta, theta), 1.0)));
  JacobiCS r;
  r.c = 1.0 / sqrt(dadd(dmul(t, t), 1.0));
  r.s = dmul(t, r.c);
  r.tapq = dmul(t, apq);
  return r;
}

APD_HD void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq,
                       double& v0p, double& v0q, double& v1p, double& v1q, double& v2p, double& v2q) {
  if (apq == 0.0) return;
  const JacobiCS r = jacobi_cs(app, aqq, apq);
  const double c = r.c, s = r.s, tapq = r.tapq;
  app = dsub(app, tapq);
  aqq = dadd(aqq, tapq);
  apq = 0.0;
  const double nrp = dsub(dmul(c, arp), dmul(s, arq));
  const double nrq = dadd(dmul(s, arp), dmul(c, arq));
  arp = nrp;
  arq = nrq;
  double a, b;
  a = v0p; b = v0q; v0p = dsub(dmul(c, a), dmul(s, b)); v0q = dadd(dmul(s, a), dmul(c, b));
  a = v1p; b = v1q; v1p = dsub(dmul(c, a), dmul(s, b)); v1q = dadd(dmul(s, a), dmul(c, b));
  a = v2p; b = v2q; v2p = dsub(dmul(c, a), dmul(s, b)); v2q = dadd(dmul(s, a), dmul(c, b));
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues DESCENDING in w, matching unit
// eigenvectors in the columns of V (row-major V[r*3+c]). For symmetric PSD input this is the SVD
// JacobiSVD returns at fast_apdgicp_impl.hpp:337 (U == V).
APD_HD void sym_eig3(const Sym3& A, double w[3], double V[9]) {
  double a00 = A.xx, a01 = A.xy, a02 = A.xz, a11 = A.yy, a12 = A.yz, a22 = A.zz;
  double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = dadd(dadd(dmul(a01, a01), dmul(a02, a02)), dmul(a12, a12));
    const double diag = dadd(dadd(dmul(a00, a00), dmul(a11, a11)), dmul(a22, a22));
    if (off <= dmul(1e-34, diag) || off == 0.0) break;
    jacobi_rot(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);  // (p,q)=(0,1), r=2
    jacobi_rot(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);  // (0,2), r=1
    jacobi_rot(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);  // (1,2), r=0
  }
  double w0 = a00, w1 = a11, w2 = a22;
  // sort descending (3-element network), swapping the columns of V alongside
#define APD_SWAPCOL(wa, wb, va0, vb0, va1, vb1, va2, vb2) \
  {                                                        \
    double t_ = wa; wa = wb; wb = t_;                      \
    t_ = va0; va0 = vb0; vb0 = t_;                         \
    t_ = va1; va1 = vb1; vb1 = t_;                         \
    t_ = va2; va2 = vb2; vb2 = t_;                         \
  }
  if (w0 < w1) APD_SWAPCOL(w0, w1, v00, v01, v10, v11, v20, v21)
  if (w1 < w2) APD_SWAPCOL(w1, w2, v01, v02, v11, v12, v21, v22)
  if (w0 < w1) APD_SWAPCOL(w0, w1, v00, v01, v10, v11, v20, v21)
#undef APD_SWAPCOL
  w[0] = w0; w[1] = w1; w[2] = w2;
  V[0] = v00; V[1] = v01; V[2] = v02; V[3] = v10; V[4] = v11; V[5] = v12; V[6] = v20; V[7] = v21; V[8] = v22;
}

// V diag(d) V^T, summed over k = 0,1,2 in that order with ((V_ik * d_k) * V_jk) like the oracle
APD_HD Sym3 recompose(const double V[9], const double d[3]) {
  Sym3 c;
#define APD_RC(i, j) dadd(dadd(dadd(0.0, dmul(dmul(V[i * 3 + 0], d[0]), V[j * 3 + 0])), dmul(dmul(V[i * 3 + 1], d[1]), V[j * 3 + 1])), dmul(dmul(V[i * 3 + 2], d[2]), V[j * 3 + 2]))
  c.xx = APD_RC(0, 0);
  c.xy = APD_RC(0, 1);
  c.xz = APD_RC(0, 2);
  c.yy = APD_RC(1, 1);
  c.yz = APD_RC(1, 2);
  c.zz = APD_RC(2, 2);
#undef APD_RC
  return c;
}

// LDL^T with diagonal pivoting of a symmetric 6x6 (row-major A[36], destroyed) and solve A x = rhs.
// Zero-pivot rule of Eigen's LDLT::solve: |D_i| <= max|D| * eps contributes 0.
// One thread of the align kernel runs this between two block barriers while the rest of the team waits, so it is written
// for LATENCY: only the lower triangle is kept (21 entries: the algorithm never reads the upper one after it goes stale),
// every loop is unrolled and every index is a compile-time constant (the pivot, the only dynamic quantity, selects one of at
// most five statically indexed swap blocks), so the matrix, the permutation and the right-hand side stay in registers. The
// rolled version indexed a local-memory 6x6 with the pivot: a few hundred dependent local loads per solve. The arithmetic,
// its order and the pivoting rule are those of the rolled version (and of oracle/linalg.hpp).
#ifdef __CUDACC__
#define APD_UNROLL _Pragma("unroll")
#else
#define APD_UNROLL
#endif
#define APD_L(i, j) L[(i) * ((i) + 1) / 2 + (j)]  // i >= j
#define APD_SWAPD(a, b) { const double t_ = a; a = b; b = t_; }
// Solves (Ain + diag_add * I) x = rhs_sign * rhs: the damped system of step_lm (lsq_registration_impl.hpp:137) is read straight
// from the reduced H and b without a copy.
APD_HD_NOINLINE void ldlt6_solve(const double* Ain, const double* rhs, double* x, double diag_add = 0.0, double rhs_sign = 1.0) {
  double L[21], y[6];
  int perm[6];
  APD_UNROLL
  for (int i = 0; i < 6; i++) {
    APD_UNROLL
    for (int j = 0; j <= i; j++) APD_L(i, j) = Ain[i * 6 + j];
    if (diag_add != 0.0) APD_L(i, i) += diag_add;
    perm[i] = i;
    y[i] = rhs_sign < 0.0 ? -rhs[i] : rhs[i];
  }
  APD_UNROLL
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double best = fabs(APD_L(k, k));
    APD_UNROLL
    for (int i = k + 1; i < 6; i++) {
      const double v = fabs(APD_L(i, i));
      if (v > best) { best = v; piv = i; }
    }
    APD_UNROLL
    for (int p = k + 1; p < 6; p++) {
      if (piv == p) {  // symmetric transposition k <-> p on the lower triangle; the right-hand side moves along (y[i] = rhs[perm[i]])
        APD_SWAPD(APD_L(k, k), APD_L(p, p))
        APD_UNROLL
        for (int j = 0; j < k; j++) APD_SWAPD(APD_L(k, j), APD_L(p, j))
        APD_UNROLL
        for (int i = k + 1; i < p; i++) APD_SWAPD(APD_L(i, k), APD_L(p, i))
        APD_UNROLL
        for (int i = p + 1; i < 6; i++) APD_SWAPD(APD_L(i, k), APD_L(i, p))
        const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        APD_SWAPD(y[k], y[p])
      }
    }
    const double d = APD_L(k, k);
    if (d != 0.0) {
      APD_UNROLL
      for (int i = k + 1; i < 6; i++) APD_L(i, k) /= d;
      APD_UNROLL
      for (int i = k + 1; i < 6; i++) {
        APD_UNROLL
        for (int j = k + 1; j <= i; j++) APD_L(i, j) -= APD_L(i, k) * d * APD_L(j, k);
      }
    }
  }
  double maxd = 0.0;
  APD_UNROLL
  for (int i = 0; i < 6; i++) maxd = fmax(maxd, fabs(APD_L(i, i)));
  const double tol = fmax(maxd * DBL_EPSILON, 1.0 / DBL_MAX);
  APD_UNROLL
  for (int i = 0; i < 6; i++) {
    APD_UNROLL
    for (int j = 0; j < i; j++) y[i] -= APD_L(i, j) * y[j];
  }
  APD_UNROLL
  for (int i = 0; i < 6; i++) y[i] = (fabs(APD_L(i, i)) > tol) ? y[i] / APD_L(i, i) : 0.0;
  APD_UNROLL
  for (int i = 5; i >= 0; i--) {
    APD_UNROLL
    for (int j = i + 1; j < 6; j++) y[i] -= APD_L(j, i) * y[j];
  }
  // x[perm[i]] = y[i] without a dynamically indexed store
  APD_UNROLL
  for (int o = 0; o < 6; o++) {
    double v = 0.0;
    APD_UNROLL
    for (int i = 0; i < 6; i++) v = (perm[i] == o) ? y[i] : v;
    x[o] = v;
  }
}
#undef APD_L
#undef APD_SWAPD

// so3_exp (so3.hpp:59-78) followed by Quaterniond::toRotationMatrix; R row-major
APD_HD_NOINLINE void so3_exp_matrix(const double* w, double* R) {
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

}  // namespace apd
