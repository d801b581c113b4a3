// TEST INFRASTRUCTURE — CPU oracle for the FastAPDGICP hot path. Not part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY: pinned to the reference's own sources. The reference holds no golden vector, known-answer test or fixture for
// FastAPDGICP (its only test, fast_apdgicp/src/test/gicp_test.cpp:99-128, never builds it and its data/ directory is
// absent), and PCL, Eigen, FLANN and Boost are not installed. What compiles from the reference tree (`make ref`, outputs in
// oracle/_ref): (1) the exact kd-tree it vendors (radar_graph_slam/include/scan_context/nanoflann.hpp, nanoflann 1.3.2:
// FLANN's L2_Simple float metric; oracle/ref_nanoflann.cpp), which pins the kNN / 1-NN results of this file
// (tests/test_reference_knn.py, tests/golden/knn_nanoflann_v1.npz); (2) FastAPDGICP itself - fast_apdgicp.hpp,
// lsq_registration.hpp, their impl/ files, so3.hpp - UNMODIFIED, over stand-in Eigen / PCL / Boost headers
// (oracle/ref_apdgicp.cpp, oracle/ref_standins/), which pins covariances, APD model, Mahalanobis, H / b, compute_error and
// the whole LM loop of this file to the reference's text at 1e-9 (tests/test_reference_apdgicp.py) and generates
// tests/golden/apd_ref_golden_v1.npz for the GPU box. The third-party arithmetic under that text (Eigen's product, inverse,
// JacobiSVD and LDLT kernels) remains a restatement of the published algorithms on both sides; oracle/pyref.py
// (numpy / LAPACK) is an independent third opinion.
//
// Restates, line by line (paths relative to /root/reference/fast_apdgicp/include/fast_gicp):
//   FastAPDGICP            gicp/impl/fast_apdgicp_impl.hpp:14-363   (APD_I)
//   LsqRegistration        gicp/impl/lsq_registration_impl.hpp:11-173 (LSQ_I)
//   skewd / so3_exp        so3/so3.hpp:21-31,59-78 (SO3)
// and the PCL behaviours the callers observe (pcl::Registration::align / getFitnessScore,
// pcl::transformPointCloud, pcl::search::KdTree) as documented in SURVEY.md Appendix B.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "kdtree.hpp"
#include "linalg.hpp"

namespace apd_oracle {

// gicp/gicp_settings.hpp:6
enum Regularization { REG_NONE = 0, REG_MIN_EIG = 1, REG_NORMALIZED_MIN_EIG = 2, REG_PLANE = 3, REG_FROBENIUS = 4 };
// gicp/lsq_registration.hpp:13
enum Optimizer { OPT_GAUSS_NEWTON = 0, OPT_LEVENBERG_MARQUARDT = 1 };

struct Params {
  int num_threads = 0;                  // APD_I:16,34-42 (0 -> all)
  int k_correspondences = 20;           // APD_I:21
  int regularization = REG_PLANE;       // APD_I:25
  double max_corr_dist = FLT_MAX;       // APD_I:23 (corr_dist_threshold_, a double in PCL)
  int max_iterations = 64;              // LSQ_I:13
  double rotation_epsilon = 2e-3;       // LSQ_I:14
  double transformation_epsilon = 5e-4; // LSQ_I:15
  int optimizer = OPT_LEVENBERG_MARQUARDT;  // LSQ_I:17
  int lm_max_iterations = 10;           // LSQ_I:19
  double lm_init_lambda_factor = 1e-9;  // LSQ_I:20
  double dist_var = 0.86;               // fast_apdgicp.hpp:109
  double azimuth_var = 0.5;             // fast_apdgicp.hpp:107
  double elevation_var = 1.0;           // fast_apdgicp.hpp:108
};

struct Iso {  // Eigen::Isometry3d: rotation + translation, last row (0,0,0,1)
  double R[3][3];
  double t[3];
  static Iso identity() {
    Iso x;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) x.R[i][j] = (i == j);
      x.t[i] = 0;
    }
    return x;
  }
};

inline Iso mul(const Iso& a, const Iso& b) {
  Iso c;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) c.R[i][j] = a.R[i][0] * b.R[0][j] + a.R[i][1] * b.R[1][j] + a.R[i][2] * b.R[2][j];
    c.t[i] = a.R[i][0] * b.t[0] + a.R[i][1] * b.t[1] + a.R[i][2] * b.t[2] + a.t[i];
  }
  return c;
}

// SO3:59-78 followed by Eigen's Quaterniond::toRotationMatrix()
inline void so3_exp_matrix(const double w[3], double R[3][3]) {
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0][0] = 1 - (tyy + tzz);
  R[0][1] = txy - twz;
  R[0][2] = txz + twy;
  R[1][0] = txy + twz;
  R[1][1] = 1 - (txx + tzz);
  R[1][2] = tyz - twx;
  R[2][0] = txz - twy;
  R[2][1] = tyz + twx;
  R[2][2] = 1 - (txx + tyy);
}

struct TraceRow {  // one LM trial (the table LSQ_I:148-154 prints)
  double outer, inner, y0, yi, rho, lambda, dnorm, accepted;
};

class FastAPDGICP {
public:
  Params prm;

  // ---- cloud setters: APD_I:90-108 (kd-tree build = setInputCloud) ----
  void setInputSource(const float* xyz, int stride_floats, int n) {
    load(src_, xyz, stride_floats, n);
    src_tree_.build(&src_);
    src_covs_.clear();
  }
  void setInputTarget(const float* xyz, int stride_floats, int n) {
    load(tgt_, xyz, stride_floats, n);
    tgt_tree_.build(&tgt_);
    tgt_covs_.clear();
  }
  void swapSourceAndTarget() {  // APD_I:68-75
    src_.swap(tgt_);
    std::swap(src_tree_, tgt_tree_);  // source_kdtree_.swap(target_kdtree_), APD_I:71
    src_tree_.rebind(&src_);
    tgt_tree_.rebind(&tgt_);
    src_covs_.swap(tgt_covs_);
    src_knn_.swap(tgt_knn_);
    corr_.clear();
    sq_dist_.clear();
  }
  void clearSource() { src_.clear(); src_covs_.clear(); }  // APD_I:78-81
  void clearTarget() { tgt_.clear(); tgt_covs_.clear(); }  // APD_I:84-87
  void setSourceCovariances(const std::vector<Mat3>& c) { src_covs_ = c; }  // APD_I:111-113
  void setTargetCovariances(const std::vector<Mat3>& c) { tgt_covs_ = c; }  // APD_I:116-118

  int threads() const {
#ifdef _OPENMP
    return prm.num_threads > 0 ? prm.num_threads : omp_get_max_threads();
#else
    return 1;
#endif
  }

  // ---- APD_I:303-363 ----
  bool calculate_covariances(const std::vector<P3>& cloud, const KdTree& tree, std::vector<Mat3>& covs, std::vector<int>& knn_out) {
    const int n = (int)cloud.size();
    const int k = prm.k_correspondences;
    if (n < k) return false;  // reference reads uninitialised columns here (APD_I:318-321); rejected by convention
    covs.resize(n);
    knn_out.resize((size_t)n * k);
    const int nt = threads();
#pragma omp parallel for num_threads(nt) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      std::vector<int> k_indices;
      std::vector<float> k_sq;
      tree.knn(cloud[i], k, k_indices, k_sq);  // APD_I:316 (includes i itself)
      // Fewer than k results: a non-finite query point (no neighbour at all) or a cloud with fewer than k finite
      // points. The reference then reads whatever PCL's resize left in k_indices; the convention here (and in
      // the CUDA kernel) pads with the query itself and reports -1 in the index set.
      const int found = (int)k_indices.size();
      k_indices.resize(k, i);
      double mean[3] = {0, 0, 0};
      for (int j = 0; j < k; j++) {
        const P3& p = cloud[k_indices[j]];
        mean[0] += (double)p.x;
        mean[1] += (double)p.y;
        mean[2] += (double)p.z;
        knn_out[(size_t)i * k + j] = j < found ? k_indices[j] : -1;
      }
      for (double& m : mean) m /= k;  // APD_I:323 rowwise().mean()
      Mat3 cov = Mat3::zero();
      for (int j = 0; j < k; j++) {
        const P3& p = cloud[k_indices[j]];
        const double d[3] = {(double)p.x - mean[0], (double)p.y - mean[1], (double)p.z - mean[2]};
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) cov.m[a][b] += d[a] * d[b];
      }
      for (auto& r : cov.m)
        for (auto& v : r) v /= k;  // APD_I:324 divides by k_correspondences_
      covs[i] = regularize(cov);
    }
    return true;
  }

  Mat3 regularize(const Mat3& cov) const {  // APD_I:326-359
    if (prm.regularization == REG_NONE) return cov;
    if (prm.regularization == REG_FROBENIUS) {
      Mat3 C = cov;
      for (int i = 0; i < 3; i++) C.m[i][i] += 1e-3;
      Mat3 Ci = inverse(C);
      double fro = 0;
      for (auto& r : Ci.m)
        for (auto& v : r) fro += v * v;
      fro = std::sqrt(fro);
      for (auto& r : Ci.m)
        for (auto& v : r) v /= fro;
      return inverse(Ci);
    }
    double w[3];
    Mat3 V;
    sym_eig3(cov, w, V);
    double values[3];
    switch (prm.regularization) {
      case REG_PLANE:
        values[0] = 1; values[1] = 1; values[2] = 1e-3;
        break;
      case REG_MIN_EIG:
        for (int i = 0; i < 3; i++) values[i] = std::max(w[i], 1e-3);
        break;
      case REG_NORMALIZED_MIN_EIG: {
        const double mx = std::max(w[0], std::max(w[1], w[2]));
        for (int i = 0; i < 3; i++) values[i] = std::max(w[i] / mx, 1e-3);
        break;
      }
      default:
        std::fprintf(stderr, "here must not be reached\n");
        std::abort();
    }
    return recompose(V, values);
  }

  // Float isometry applied to a float point: the documented convention for Eigen's
  // Isometry3f * Vector4f at APD_I:149: ((R0*x + R1*y) + R2*z) + t, no FMA.
  static P3 transform_f(const float Rf[3][3], const float tf[3], const P3& a) {
    P3 q;
    float* o[3] = {&q.x, &q.y, &q.z};
    for (int j = 0; j < 3; j++) {
      float s = Rf[j][0] * a.x;
      s = s + Rf[j][1] * a.y;
      s = s + Rf[j][2] * a.z;
      s = s + tf[j];
      *o[j] = s;
    }
    return q;
  }

  // ---- APD_I:133-194 ----
  void update_correspondences(const Iso& T) {
    const int N = (int)src_.size();
    float Rf[3][3], tf[3];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Rf[i][j] = (float)T.R[i][j];
      tf[i] = (float)T.t[i];
    }
    corr_.assign(N, -1);
    sq_dist_.assign(N, 0.f);
    mahal_.resize(N);
    const double thr = prm.max_corr_dist * prm.max_corr_dist;  // APD_I:156, product in double
    const double sin_az = std::sin(prm.azimuth_var / 180 * M_PI);
    const double sin_el = std::sin(prm.elevation_var / 180 * M_PI);
    Mat3 Rm;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rm.m[i][j] = T.R[i][j];
    const Mat3 RmT = transpose(Rm);
    const int nt = threads();
#pragma omp parallel for num_threads(nt) schedule(guided, 8)
    for (int i = 0; i < N; i++) {
      const P3 pt = transform_f(Rf, tf, src_[i]);
      std::vector<int> k_indices;
      std::vector<float> k_sq;
      tgt_tree_.knn(pt, 1, k_indices, k_sq);
      if (k_indices.empty()) {  // non-finite query (kdtree.hpp): no neighbour, no correspondence
        sq_dist_[i] = INFINITY;
        continue;
      }
      sq_dist_[i] = k_sq[0];
      corr_[i] = ((double)k_sq[0] < thr) ? k_indices[0] : -1;
      if (corr_[i] < 0) continue;
      const Mat3& cov_A = src_covs_[i];
      const Mat3& cov_B = tgt_covs_[corr_[i]];
      // APD_I:167-173. atan2/sqrt on float arguments resolve to the float overloads
      // (`using namespace std;` at APD_I:7).
      const double dist = std::sqrt((double)pt.x * (double)pt.x + (double)pt.y * (double)pt.y + (double)pt.z * (double)pt.z);
      const double aoa = (double)atan2_f32(pt.x, std::sqrt(pt.y * pt.y + pt.z * pt.z));
      const double s_x = dist * prm.dist_var / 400;
      const double s_y = dist * sin_az / std::cos(aoa);
      const double s_z = dist * sin_el / std::cos(aoa);
      const double elev = (double)atan2_f32(std::sqrt(pt.x * pt.x + pt.y * pt.y), pt.z);
      const double azim = (double)atan2_f32(pt.y, pt.x);
      // R = AngleAxis(azim, Z) * AngleAxis(elev, Y)  (APD_I:174-177)
      const double ca = std::cos(azim), sa = std::sin(azim), ce = std::cos(elev), se = std::sin(elev);
      const double Rr[3][3] = {{ca * ce, -sa, ca * se}, {sa * ce, ca, sa * se}, {-se, 0.0, ce}};
      const double s[3] = {s_x, s_y, s_z};
      Mat3 Cd = Mat3::zero();  // (R S)(R S)^T, APD_I:181-182
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
          for (int c = 0; c < 3; c++) Cd.m[a][b] += (Rr[a][c] * s[c]) * (Rr[b][c] * s[c]);
      // APD_I:187-192: 3x3 block of (C_B + C_d) + T (C_A + C_d) T^T, inverted
      const Mat3 RCR = (cov_B + Cd) + Rm * (cov_A + Cd) * RmT;
      mahal_[i] = inverse(RCR);
    }
  }

  // atan2f: correctly-rounded float arctangent (double evaluation rounded once). glibc's atan2f is
  // within 1 ulp of this; fixing the rounding makes oracle and device agree bit for bit.
  static float atan2_f32(float y, float x) { return (float)std::atan2((double)y, (double)x); }

  // ---- APD_I:198-272 ----
  double linearize(const Iso& T, double H[6][6], double b[6]) {
    update_correspondences(T);
    const int N = (int)src_.size();
    const int nt = threads();
    std::vector<double> Hs((size_t)nt * 36, 0.0), bs((size_t)nt * 6, 0.0);
    double sum_errors = 0.0;
#pragma omp parallel for num_threads(nt) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < N; i++) {
      const int j = corr_[i];
      if (j < 0) continue;
      const double a[3] = {(double)src_[i].x, (double)src_[i].y, (double)src_[i].z};
      const double bb[3] = {(double)tgt_[j].x, (double)tgt_[j].y, (double)tgt_[j].z};
      double p[3], e[3];
      for (int r = 0; r < 3; r++) {
        p[r] = T.R[r][0] * a[0] + T.R[r][1] * a[1] + T.R[r][2] * a[2] + T.t[r];
        e[r] = bb[r] - p[r];
      }
      const Mat3& M = mahal_[i];
      double Me[3];
      for (int r = 0; r < 3; r++) Me[r] = M.m[r][0] * e[0] + M.m[r][1] * e[1] + M.m[r][2] * e[2];
      sum_errors += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
      if (!H || !b) continue;
      // J = [ skew(p) | -I ]  (APD_I:248-251, SO3:21-31)
      double J[3][6] = {{0, -p[2], p[1], -1, 0, 0}, {p[2], 0, -p[0], 0, -1, 0}, {-p[1], p[0], 0, 0, 0, -1}};
      double MJ[3][6];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 6; c++) MJ[r][c] = M.m[r][0] * J[0][c] + M.m[r][1] * J[1][c] + M.m[r][2] * J[2][c];
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      double* Ht = &Hs[(size_t)tid * 36];
      double* bt = &bs[(size_t)tid * 6];
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) Ht[r * 6 + c] += J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c] + J[2][r] * MJ[2][c];
        bt[r] += J[0][r] * Me[0] + J[1][r] * Me[1] + J[2][r] * Me[2];
      }
    }
    if (H && b) {
      for (int r = 0; r < 6; r++) {
        b[r] = 0;
        for (int c = 0; c < 6; c++) H[r][c] = 0;
      }
      for (int t = 0; t < nt; t++)
        for (int r = 0; r < 6; r++) {
          b[r] += bs[(size_t)t * 6 + r];
          for (int c = 0; c < 6; c++) H[r][c] += Hs[(size_t)t * 36 + r * 6 + c];
        }
    }
    return sum_errors;
  }

  // ---- APD_I:275-298: stale correspondences and Mahalanobis ----
  double compute_error(const Iso& T) const {
    const int N = (int)src_.size();
    const int nt = threads();
    double sum_errors = 0.0;
#pragma omp parallel for num_threads(nt) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < N; i++) {
      const int j = corr_[i];
      if (j < 0) continue;
      const double a[3] = {(double)src_[i].x, (double)src_[i].y, (double)src_[i].z};
      const double bb[3] = {(double)tgt_[j].x, (double)tgt_[j].y, (double)tgt_[j].z};
      double e[3];
      for (int r = 0; r < 3; r++) e[r] = bb[r] - (T.R[r][0] * a[0] + T.R[r][1] * a[1] + T.R[r][2] * a[2] + T.t[r]);
      const Mat3& M = mahal_[i];
      double Me[3];
      for (int r = 0; r < 3; r++) Me[r] = M.m[r][0] * e[0] + M.m[r][1] * e[1] + M.m[r][2] * e[2];
      sum_errors += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
    }
    return sum_errors;
  }

  // ---- LSQ_I:83-92 ----
  bool is_converged(const Iso& delta) const {
    double rmax = 0, tmax = 0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) rmax = std::max(rmax, 1.0 / prm.rotation_epsilon * std::fabs(delta.R[i][j] - (i == j ? 1.0 : 0.0)));
      tmax = std::max(tmax, 1.0 / prm.transformation_epsilon * std::fabs(delta.t[i]));
    }
    return std::max(rmax, tmax) < 1;
  }

  static Iso delta_from(const double d[6]) {
    Iso delta = Iso::identity();
    so3_exp_matrix(d, delta.R);
    delta.t[0] = d[3];
    delta.t[1] = d[4];
    delta.t[2] = d[5];
    return delta;
  }

  // ---- LSQ_I:107-123 ----
  bool step_gn(Iso& x0, Iso& delta) {
    double H[6][6], b[6], nb[6], d[6];
    linearize(x0, H, b);
    for (int i = 0; i < 6; i++) nb[i] = -b[i];
    ldlt6_solve(H, nb, d);
    delta = delta_from(d);
    x0 = mul(delta, x0);
    std::memcpy(final_hessian_, H, sizeof(H));
    return true;
  }

  // ---- LSQ_I:127-173 ----
  bool step_lm(Iso& x0, Iso& delta, int outer) {
    double H[6][6], b[6];
    const double y0 = linearize(x0, H, b);
    if (lm_lambda_ < 0.0) {
      double mx = 0;
      for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i][i]));
      lm_lambda_ = prm.lm_init_lambda_factor * mx;
    }
    double nu = 2.0;
    for (int i = 0; i < prm.lm_max_iterations; i++) {
      double A[6][6], nb[6], d[6];
      for (int r = 0; r < 6; r++) {
        nb[r] = -b[r];
        for (int c = 0; c < 6; c++) A[r][c] = H[r][c] + (r == c ? lm_lambda_ : 0.0);
      }
      ldlt6_solve(A, nb, d);
      delta = delta_from(d);
      const Iso xi = mul(delta, x0);
      const double yi = compute_error(xi);
      double denom = 0, dn = 0;
      for (int r = 0; r < 6; r++) {
        denom += d[r] * (lm_lambda_ * d[r] - b[r]);
        dn += d[r] * d[r];
      }
      const double rho = (y0 - yi) / denom;
      const bool reject = rho < 0;
      trace_.push_back(TraceRow{(double)outer, (double)i, y0, yi, rho, lm_lambda_, std::sqrt(dn), reject ? 0.0 : 1.0});
      if (reject) {
        if (is_converged(delta)) return true;  // x0 NOT updated (LSQ_I:156-159)
        lm_lambda_ = nu * lm_lambda_;
        nu = 2 * nu;
        continue;
      }
      x0 = xi;
      lm_lambda_ = lm_lambda_ * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(final_hessian_, H, sizeof(H));
      return true;
    }
    return false;
  }

  // ---- pcl::Registration::align + APD_I:121-130 + LSQ_I:55-80 ----
  // returns 0 ok, -1 no target/source, -2 cloud smaller than k
  int align(const float guess[16]) {
    if (tgt_.empty() || src_.empty()) return -1;
    converged_ = false;
    for (int i = 0; i < 16; i++) final_T_[i] = (i % 5 == 0) ? 1.f : 0.f;
    if (src_covs_.size() != src_.size())
      if (!calculate_covariances(src_, src_tree_, src_covs_, src_knn_)) return -2;
    if (tgt_covs_.size() != tgt_.size())
      if (!calculate_covariances(tgt_, tgt_tree_, tgt_covs_, tgt_knn_)) return -2;
    Iso x0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) x0.R[i][j] = (double)guess[i * 4 + j];
      x0.t[i] = (double)guess[i * 4 + 3];
    }
    lm_lambda_ = -1.0;
    trace_.clear();
    nr_iterations_ = 0;
    lm_failed_ = false;
    for (int i = 0; i < prm.max_iterations && !converged_; i++) {
      nr_iterations_ = i;
      Iso delta = Iso::identity();
      const bool ok = (prm.optimizer == OPT_GAUSS_NEWTON) ? step_gn(x0, delta) : step_lm(x0, delta, i);
      if (!ok) {
        if (verbose_) std::fprintf(stderr, "lm not converged!!\n");
        lm_failed_ = true;  // LSQ_I:71-74: converged_ stays false
        break;
      }
      converged_ = is_converged(delta);
    }
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) final_T_[i * 4 + j] = (float)x0.R[i][j];
      final_T_[i * 4 + 3] = (float)x0.t[i];
    }
    final_T_[12] = final_T_[13] = final_T_[14] = 0.f;
    final_T_[15] = 1.f;
    return 0;
  }

  // pcl::transformPointCloud(*input_, output, final_transformation_) at LSQ_I:79
  void transformed_source(const float T[16], std::vector<P3>& out) const {
    float Rf[3][3], tf[3];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Rf[i][j] = T[i * 4 + j];
      tf[i] = T[i * 4 + 3];
    }
    out.resize(src_.size());
    for (size_t i = 0; i < src_.size(); i++) out[i] = transform_f(Rf, tf, src_[i]);
  }

  // pcl::Registration::getFitnessScore(max_range)
  double getFitnessScore(double max_range = DBL_MAX) const {
    std::vector<P3> tr;
    transformed_source(final_T_, tr);
    double sum = 0;
    long nr = 0;
    std::vector<int> ki;
    std::vector<float> kd;
    for (size_t i = 0; i < tr.size(); i++) {
      tgt_tree_.knn(tr[i], 1, ki, kd);
      if (ki.empty()) continue;
      if ((double)kd[0] <= max_range) {
        sum += (double)kd[0];
        nr++;
      }
    }
    return nr > 0 ? sum / (double)nr : DBL_MAX;
  }

  // LSQ_I:50-52
  double evaluateCost(const float pose[16], double H[6][6], double b[6]) {
    Iso x;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) x.R[i][j] = (double)pose[i * 4 + j];
      x.t[i] = (double)pose[i * 4 + 3];
    }
    return linearize(x, H, b);
  }

  bool ensure_covariances() {
    if (src_covs_.size() != src_.size() && !src_.empty())
      if (!calculate_covariances(src_, src_tree_, src_covs_, src_knn_)) return false;
    if (tgt_covs_.size() != tgt_.size() && !tgt_.empty())
      if (!calculate_covariances(tgt_, tgt_tree_, tgt_covs_, tgt_knn_)) return false;
    return true;
  }

  // observable state
  std::vector<P3> src_, tgt_;
  KdTree src_tree_, tgt_tree_;
  std::vector<Mat3> src_covs_, tgt_covs_, mahal_;
  std::vector<int> src_knn_, tgt_knn_;
  std::vector<int> corr_;
  std::vector<float> sq_dist_;
  std::vector<TraceRow> trace_;
  double final_hessian_[6][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {0, 0, 0, 1, 0, 0}, {0, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 1}};
  float final_T_[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  bool converged_ = false;
  int nr_iterations_ = 0;
  double lm_lambda_ = -1.0;
  bool verbose_ = false;
  bool lm_failed_ = false;  // the last align printed "lm not converged!!" (LSQ_I:72)

private:
  static void load(std::vector<P3>& dst, const float* xyz, int stride_floats, int n) {
    dst.resize(n);
    for (int i = 0; i < n; i++) dst[i] = P3{xyz[(size_t)i * stride_floats], xyz[(size_t)i * stride_floats + 1], xyz[(size_t)i * stride_floats + 2]};
  }
};

}  // namespace apd_oracle
