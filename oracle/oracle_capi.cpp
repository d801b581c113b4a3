// TEST INFRASTRUCTURE — C entry points of the CPU oracle, for ctypes (tests/, smoke(), bench.py's
// cpu_baseline and --impl reference legs). Not part of the product; the product never links this.
#include <chrono>
#include <cstring>

#include "apd_oracle.hpp"
#include "preprocess_oracle.hpp"

using apd_oracle::FastAPDGICP;
using apd_oracle::Mat3;

extern "C" {

struct oracle_params {
  int num_threads;
  int k_correspondences;
  int regularization;
  int max_iterations;
  int optimizer;
  int lm_max_iterations;
  double max_corr_dist;
  double rotation_epsilon;
  double transformation_epsilon;
  double lm_init_lambda_factor;
  double dist_var;
  double azimuth_var;
  double elevation_var;
};

void* oracle_create() { return new FastAPDGICP(); }
void oracle_destroy(void* h) { delete static_cast<FastAPDGICP*>(h); }

void oracle_default_params(oracle_params* p) {
  apd_oracle::Params d;
  p->num_threads = d.num_threads;
  p->k_correspondences = d.k_correspondences;
  p->regularization = d.regularization;
  p->max_iterations = d.max_iterations;
  p->optimizer = d.optimizer;
  p->lm_max_iterations = d.lm_max_iterations;
  p->max_corr_dist = d.max_corr_dist;
  p->rotation_epsilon = d.rotation_epsilon;
  p->transformation_epsilon = d.transformation_epsilon;
  p->lm_init_lambda_factor = d.lm_init_lambda_factor;
  p->dist_var = d.dist_var;
  p->azimuth_var = d.azimuth_var;
  p->elevation_var = d.elevation_var;
}

void oracle_set_params(void* h, const oracle_params* p) {
  auto& d = static_cast<FastAPDGICP*>(h)->prm;
  d.num_threads = p->num_threads;
  d.k_correspondences = p->k_correspondences;
  d.regularization = p->regularization;
  d.max_iterations = p->max_iterations;
  d.optimizer = p->optimizer;
  d.lm_max_iterations = p->lm_max_iterations;
  d.max_corr_dist = p->max_corr_dist;
  d.rotation_epsilon = p->rotation_epsilon;
  d.transformation_epsilon = p->transformation_epsilon;
  d.lm_init_lambda_factor = p->lm_init_lambda_factor;
  d.dist_var = p->dist_var;
  d.azimuth_var = p->azimuth_var;
  d.elevation_var = p->elevation_var;
}

void oracle_set_source(void* h, const float* xyz, int stride_floats, int n) { static_cast<FastAPDGICP*>(h)->setInputSource(xyz, stride_floats, n); }
void oracle_set_target(void* h, const float* xyz, int stride_floats, int n) { static_cast<FastAPDGICP*>(h)->setInputTarget(xyz, stride_floats, n); }
void oracle_swap(void* h) { static_cast<FastAPDGICP*>(h)->swapSourceAndTarget(); }
void oracle_clear_source(void* h) { static_cast<FastAPDGICP*>(h)->clearSource(); }
void oracle_clear_target(void* h) { static_cast<FastAPDGICP*>(h)->clearTarget(); }

int oracle_align(void* h, const float* guess16, float* T16, int* converged, int* iterations) {
  auto* o = static_cast<FastAPDGICP*>(h);
  static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const int rc = o->align(guess16 ? guess16 : ident);
  if (T16) std::memcpy(T16, o->final_T_, sizeof(float) * 16);
  if (converged) *converged = o->converged_ ? 1 : 0;
  if (iterations) *iterations = o->nr_iterations_;
  return rc;
}

int oracle_lm_failed(void* h) { return static_cast<FastAPDGICP*>(h)->lm_failed_ ? 1 : 0; }

// InformationMatrixCalculator::calc_fitness_score (information_matrix_calculator.cpp:55-86): same pass with an explicit pose
double oracle_fitness_score(void* h, const float* T16, double max_range) {
  auto* o = static_cast<FastAPDGICP*>(h);
  float saved[16];
  std::memcpy(saved, o->final_T_, sizeof(saved));
  std::memcpy(o->final_T_, T16, sizeof(saved));
  const double f = o->getFitnessScore(max_range);
  std::memcpy(o->final_T_, saved, sizeof(saved));
  return f;
}

double oracle_fitness(void* h, double max_range) { return static_cast<FastAPDGICP*>(h)->getFitnessScore(max_range); }

int oracle_compute_covariances(void* h) { return static_cast<FastAPDGICP*>(h)->ensure_covariances() ? 0 : -2; }

double oracle_linearize(void* h, const float* pose16, double* H36, double* b6) {
  auto* o = static_cast<FastAPDGICP*>(h);
  if (!o->ensure_covariances()) return -1.0;
  double H[6][6], b[6];
  const double e = o->evaluateCost(pose16, H36 ? H : nullptr, b6 ? b : nullptr);
  if (H36 && b6) {
    std::memcpy(H36, H, sizeof(H));
    std::memcpy(b6, b, sizeof(b));
  }
  return e;
}

// the protected hooks at a DOUBLE pose (they take an Eigen::Isometry3d): linearize (APD_I:198-272) and compute_error (APD_I:275-298)
static apd_oracle::Iso iso_from_d16(const double* p) {
  apd_oracle::Iso x;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) x.R[i][j] = p[i * 4 + j];
    x.t[i] = p[i * 4 + 3];
  }
  return x;
}
double oracle_linearize_d(void* h, const double* pose16, double* H36, double* b6) {
  auto* o = static_cast<FastAPDGICP*>(h);
  if (!o->ensure_covariances()) return -1.0;
  double H[6][6], b[6];
  const double e = o->linearize(iso_from_d16(pose16), H, b);
  if (H36) std::memcpy(H36, H, sizeof(H));
  if (b6) std::memcpy(b6, b, sizeof(b));
  return e;
}
double oracle_compute_error_d(void* h, const double* pose16) { return static_cast<FastAPDGICP*>(h)->compute_error(iso_from_d16(pose16)); }

// which: 0 = source, 1 = target
int oracle_get_knn(void* h, int which, int* out) {
  auto* o = static_cast<FastAPDGICP*>(h);
  const auto& v = which ? o->tgt_knn_ : o->src_knn_;
  std::memcpy(out, v.data(), v.size() * sizeof(int));
  return (int)v.size();
}

int oracle_get_covariances(void* h, int which, double* out9) {
  auto* o = static_cast<FastAPDGICP*>(h);
  const auto& v = which ? o->tgt_covs_ : o->src_covs_;
  for (size_t i = 0; i < v.size(); i++) std::memcpy(out9 + i * 9, v[i].m, sizeof(double) * 9);
  return (int)v.size();
}

void oracle_set_covariances(void* h, int which, const double* in9, int n) {
  auto* o = static_cast<FastAPDGICP*>(h);
  std::vector<Mat3> v(n);
  for (int i = 0; i < n; i++) std::memcpy(v[i].m, in9 + (size_t)i * 9, sizeof(double) * 9);
  if (which) o->setTargetCovariances(v); else o->setSourceCovariances(v);
}

int oracle_get_correspondences(void* h, int* corr, float* sq_dist) {
  auto* o = static_cast<FastAPDGICP*>(h);
  if (corr) std::memcpy(corr, o->corr_.data(), o->corr_.size() * sizeof(int));
  if (sq_dist) std::memcpy(sq_dist, o->sq_dist_.data(), o->sq_dist_.size() * sizeof(float));
  return (int)o->corr_.size();
}

int oracle_get_mahalanobis(void* h, double* out9) {
  auto* o = static_cast<FastAPDGICP*>(h);
  for (size_t i = 0; i < o->mahal_.size(); i++) std::memcpy(out9 + i * 9, o->mahal_[i].m, sizeof(double) * 9);
  return (int)o->mahal_.size();
}

void oracle_get_final_hessian(void* h, double* H36) { std::memcpy(H36, static_cast<FastAPDGICP*>(h)->final_hessian_, sizeof(double) * 36); }

int oracle_get_trace(void* h, double* out8, int max_rows) {
  auto* o = static_cast<FastAPDGICP*>(h);
  const int n = std::min<int>(max_rows, (int)o->trace_.size());
  if (out8) std::memcpy(out8, o->trace_.data(), sizeof(apd_oracle::TraceRow) * n);
  return (int)o->trace_.size();
}

void oracle_transform_source(void* h, const float* T16, float* out_xyz) {
  auto* o = static_cast<FastAPDGICP*>(h);
  std::vector<apd_oracle::P3> tr;
  o->transformed_source(T16, tr);
  std::memcpy(out_xyz, tr.data(), tr.size() * sizeof(apd_oracle::P3));
}

// exact kNN of arbitrary queries in the target/source cloud via brute force (cross-check of the kd-tree)
void oracle_knn_bruteforce(const float* cloud_xyz, int n, const float* query_xyz, int nq, int k, int* out_idx, float* out_d2) {
  std::vector<apd_oracle::P3> pts(n);
  std::memcpy(pts.data(), cloud_xyz, sizeof(apd_oracle::P3) * n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; i++) {
    std::vector<int> ki;
    std::vector<float> kd;
    apd_oracle::knn_bruteforce(pts, apd_oracle::P3{query_xyz[3 * i], query_xyz[3 * i + 1], query_xyz[3 * i + 2]}, k, ki, kd);
    for (size_t j = 0; j < ki.size(); j++) {
      out_idx[(size_t)i * k + j] = ki[j];
      out_d2[(size_t)i * k + j] = kd[j];
    }
  }
}

void oracle_knn_kdtree(const float* cloud_xyz, int n, const float* query_xyz, int nq, int k, int* out_idx, float* out_d2) {
  std::vector<apd_oracle::P3> pts(n);
  std::memcpy(pts.data(), cloud_xyz, sizeof(apd_oracle::P3) * n);
  apd_oracle::KdTree tree;
  tree.build(&pts);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; i++) {
    std::vector<int> ki;
    std::vector<float> kd;
    tree.knn(apd_oracle::P3{query_xyz[3 * i], query_xyz[3 * i + 1], query_xyz[3 * i + 2]}, k, ki, kd);
    for (size_t j = 0; j < ki.size(); j++) {
      out_idx[(size_t)i * k + j] = ki[j];
      out_d2[(size_t)i * k + j] = kd[j];
    }
  }
}

// One full registration as BASELINE.md defines it (set target + set source + align + fitness),
// timed with steady_clock; returns seconds. Used by the CPU baseline legs of bench.py.
double oracle_timed_registration(void* h, const float* src_xyz, int ns, const float* tgt_xyz, int nt, int stride_floats,
                                 const float* guess16, int reuse_target, float* T16, int* converged, int* iterations, double* fitness) {
  auto* o = static_cast<FastAPDGICP*>(h);
  const auto t0 = std::chrono::steady_clock::now();
  if (!reuse_target) o->setInputTarget(tgt_xyz, stride_floats, nt);
  o->setInputSource(src_xyz, stride_floats, ns);
  oracle_align(h, guess16, T16, converged, iterations);
  const double f = o->getFitnessScore();
  if (fitness) *fitness = f;
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- "next" rows 8(f)-2 / 8(f)-4: preprocessing filters and submap accumulation (preprocess_oracle.hpp) ----
// clouds are (n, 4) float arrays x, y, z, intensity; outputs are written to `out` (capacity n_in) and the count returned
static std::vector<apd_oracle::PointI> to_pts(const float* xyzi, int n) {
  std::vector<apd_oracle::PointI> v(n);
  std::memcpy(v.data(), xyzi, sizeof(apd_oracle::PointI) * n);
  return v;
}
int oracle_distance_filter(const float* xyzi, int n, double near_thresh, double far_thresh, double z_low, double z_high, float* out) {
  const auto r = apd_oracle::distance_filter(to_pts(xyzi, n), near_thresh, far_thresh, z_low, z_high);
  std::memcpy(out, r.data(), sizeof(apd_oracle::PointI) * r.size());
  return (int)r.size();
}
int oracle_voxel_grid(const float* xyzi, int n, float leaf, float* out) {
  std::vector<apd_oracle::PointI> r;
  apd_oracle::voxel_grid(to_pts(xyzi, n), leaf, r);
  std::memcpy(out, r.data(), sizeof(apd_oracle::PointI) * r.size());
  return (int)r.size();
}
int oracle_approx_voxel_grid(const float* xyzi, int n, float leaf, float* out) {
  const auto r = apd_oracle::approx_voxel_grid(to_pts(xyzi, n), leaf);
  std::memcpy(out, r.data(), sizeof(apd_oracle::PointI) * r.size());
  return (int)r.size();
}
int oracle_radius_outlier_removal(const float* xyzi, int n, double radius, int min_pts, float* out) {
  const auto r = apd_oracle::radius_outlier_removal(to_pts(xyzi, n), radius, min_pts);
  std::memcpy(out, r.data(), sizeof(apd_oracle::PointI) * r.size());
  return (int)r.size();
}
int oracle_statistical_outlier_removal(const float* xyzi, int n, int mean_k, double stddev_mult, float* out) {
  const auto r = apd_oracle::statistical_outlier_removal(to_pts(xyzi, n), mean_k, stddev_mult);
  std::memcpy(out, r.data(), sizeof(apd_oracle::PointI) * r.size());
  return (int)r.size();
}
int oracle_accumulate_submap_m(const float* xyzi, const int* offsets, int n_clouds, const double* rel_poses16, float leaf, int approx, float* out);
int oracle_accumulate_submap(const float* xyzi, const int* offsets, int n_clouds, const double* rel_poses16, float leaf, float* out) {
  return oracle_accumulate_submap_m(xyzi, offsets, n_clouds, rel_poses16, leaf, 0, out);
}
// approx != 0: downsample_method APPROX_VOXELGRID (scan_matching_odometry_nodelet.cpp:156-160)
int oracle_accumulate_submap_m(const float* xyzi, const int* offsets, int n_clouds, const double* rel_poses16, float leaf, int approx, float* out) {
  std::vector<std::vector<apd_oracle::PointI>> clouds(n_clouds);
  std::vector<const double*> poses(n_clouds);
  for (int k = 0; k < n_clouds; k++) {
    clouds[k] = to_pts(xyzi + (size_t)offsets[k] * 4, offsets[k + 1] - offsets[k]);
    poses[k] = rel_poses16 + (size_t)k * 16;
  }
  auto acc = apd_oracle::accumulate_submap(clouds, poses);
  if (leaf > 0.f) {
    std::vector<apd_oracle::PointI> r;
    if (approx) r = apd_oracle::approx_voxel_grid(acc, leaf); else apd_oracle::voxel_grid(acc, leaf, r);
    acc.swap(r);
  }
  std::memcpy(out, acc.data(), sizeof(apd_oracle::PointI) * acc.size());
  return (int)acc.size();
}

int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
