// GPU-backed base of the drop-in fast_gicp::FastAPDGICP: the part of fast_gicp::LsqRegistration
// (reference fast_apdgicp/include/fast_gicp/gicp/lsq_registration.hpp:16-84, impl/lsq_registration_impl.hpp) the callers and
// subclasses see, on top of the apdgicp_b200 C ABI. The optimisation loop itself (lsq_registration_impl.hpp:55-173) runs
// inside the GPU align kernel.
//
// Deliberately NOT named LsqRegistration and NOT at <fast_gicp/gicp/lsq_registration.hpp>: the reference's FastGICP and
// FastVGICP (registrations.cpp:13-14) keep deriving from the reference's own CPU base class, which this header must not
// shadow. Both can be included in one translation unit (tests/cpp/coexist_test.cpp).
//
// Compiles against real PCL + Eigen (ROS machine) and against include/pcl_shim (in-container stand-ins, tests/cpp). Eigen
// members used: Matrix4f / Matrix4d / Matrix<double,6,6> ::data(), operator()(r,c), ::Identity(), Isometry3d::matrix().
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>

#include <apdgicp_b200.h>

#define APDGICP_B200_DROPIN 1

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class LsqRegistrationB200 : public pcl::Registration<PointSource, PointTarget, float> {
public:
  using Scalar = float;
  using Base = pcl::Registration<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename Base::Matrix4;
  using PointCloudSource = typename Base::PointCloudSource;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename Base::PointCloudTarget;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

protected:
  using Base::converged_;
  using Base::final_transformation_;
  using Base::max_iterations_;
  using Base::nr_iterations_;
  using Base::transformation_epsilon_;

public:
  explicit LsqRegistrationB200(int device = 0) {
    this->reg_name_ = "LsqRegistration";
    max_iterations_ = 64;            // lsq_registration_impl.hpp:13
    transformation_epsilon_ = 5e-4;  // :15
    final_hessian_.setIdentity();    // :23
    if (apd_create(device, &handle_) != APD_OK) {
      // no CPU fallback: the object stays unusable and every align() reports not-converged
      std::fprintf(stderr, "[apdgicp_b200] %s\n", apd_last_error(nullptr));
      handle_ = nullptr;
    }
    apd_default_params(&params_);
  }
  virtual ~LsqRegistrationB200() { apd_destroy(handle_); }
  LsqRegistrationB200(const LsqRegistrationB200&) = delete;
  LsqRegistrationB200& operator=(const LsqRegistrationB200&) = delete;

  void setRotationEpsilon(double eps) { params_.rotation_epsilon = eps; }                    // :30
  void setInitialLambdaFactor(double f) { params_.lm_init_lambda_factor = f; }              // :35
  void setDebugPrint(bool on) { lm_debug_print_ = on; }                                     // :40
  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const { return final_hessian_; }     // :45

  // evaluateCost(relative_pose, H, b) = linearize(Isometry3d(relative_pose.cast<double>()), H, b) (:50-52)
  double evaluateCost(const Eigen::Matrix4f& relative_pose, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) {
    Eigen::Isometry3d x;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) x.matrix()(r, c) = (double)relative_pose(r, c);
    return this->linearize(x, H, b);
  }

  virtual void swapSourceAndTarget() {}
  virtual void clearSource() {}
  virtual void clearTarget() {}

  apd_handle nativeHandle() const { return handle_; }
  // Status of the last align (apd_status, or APD_STATUS_LM_FAILED for "lm not converged!!")
  int lastStatus() const { return last_.status; }
  double lastFitnessScore() const { return last_.fitness; }  // getFitnessScore(DBL_MAX), computed inside the align kernel
  // publish_scan_matching_status's inlier pass (scan_matching_odometry_nodelet.cpp:698-712) without N CPU kd-tree queries
  long long lastInlierCount(double max_correspondence_dist = 0.5) {
    int64_t n = 0;
    if (!handle_ || apd_inlier_count(handle_, nullptr, max_correspondence_dist, &n) != APD_OK) return 0;
    return (long long)n;
  }

protected:
  // pcl::Registration::align -> computeTransformation (reference lsq_registration_impl.hpp:55-80). step_optimize / step_lm /
  // step_gn (:95-173) are not separate host calls here: the whole loop stays on the device (BASELINE.json north_star).
  void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    converged_ = false;
    nr_iterations_ = 0;
    if (!handle_ || !sync_inputs()) return;
    float g[16];
    to_row_major(guess, g);
    const int rc = apd_align(handle_, g, &last_);
    if (rc != APD_OK) {
      std::fprintf(stderr, "[apdgicp_b200] align failed: %s\n", apd_last_error(handle_));
      return;
    }
    if (last_.status == APD_STATUS_LM_FAILED) std::fprintf(stderr, "lm not converged!!\n");  // :72
    nr_iterations_ = last_.iterations;
    converged_ = last_.converged != 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) final_transformation_(r, c) = last_.T[r * 4 + c];
    double Hh[36];
    if (apd_get_final_hessian(handle_, Hh) == APD_OK) for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) final_hessian_(r, c) = Hh[r * 6 + c];
    if (lm_debug_print_) print_trace();
    // pcl::transformPointCloud(*input_, output, final_transformation_) (:79): xyz from the GPU, the other fields were copied by align()
    if (output.points.size() == this->input_->points.size() && !output.points.empty())
      apd_transform_source(handle_, nullptr, reinterpret_cast<float*>(output.points.data()), (int)sizeof(PointSource), APD_MEM_HOST);
  }

  // ---- the protected virtual surface of the reference (lsq_registration.hpp:64-76, fast_apdgicp.hpp:77-83) ----
  // is_converged (lsq_registration_impl.hpp:83-92), on the host: pure arithmetic on a 4x4
  bool is_converged(const Eigen::Isometry3d& delta) const {
    double rmax = 0.0, tmax = 0.0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) rmax = std::fmax(rmax, 1.0 / params_.rotation_epsilon * std::fabs(delta.matrix()(i, j) - (i == j ? 1.0 : 0.0)));
      tmax = std::fmax(tmax, 1.0 / transformation_epsilon_ * std::fabs(delta.matrix()(i, 3)));
    }
    return std::fmax(rmax, tmax) < 1.0;
  }
  // linearize at a double pose: update_correspondences + H, b, error (fast_apdgicp_impl.hpp:198-272)
  virtual double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) {
    if (!handle_ || !sync_inputs()) return 0.0;
    double pose[16], Hh[36], bh[6], err = 0.0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) pose[r * 4 + c] = trans.matrix()(r, c);
    if (apd_linearize_d(handle_, pose, Hh, bh, &err) != APD_OK) return 0.0;
    if (H) for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) (*H)(r, c) = Hh[r * 6 + c];
    if (b) for (int r = 0; r < 6; r++) (*b)(r) = bh[r];
    return err;
  }
  // compute_error at a double pose with the correspondences of the last linearize (fast_apdgicp_impl.hpp:275-298)
  virtual double compute_error(const Eigen::Isometry3d& trans) {
    if (!handle_) return 0.0;
    double pose[16], err = 0.0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) pose[r * 4 + c] = trans.matrix()(r, c);
    return apd_compute_error(handle_, pose, &err) == APD_OK ? err : 0.0;
  }

  // push parameters to the device; derived classes add their clouds
  virtual bool sync_inputs() {
    params_.max_iterations = max_iterations_;
    params_.transformation_epsilon = transformation_epsilon_;
    params_.max_corr_dist = this->corr_dist_threshold_;
    return apd_set_params(handle_, &params_) == APD_OK;
  }

  template <typename M>
  static void to_row_major(const M& m, float out[16]) {
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[r * 4 + c] = (float)m(r, c);
  }

  void print_trace() {
    int n = 0;
    apd_get_lm_trace(handle_, nullptr, 0, &n);
    if (n <= 0) return;
    std::vector<double> rows((size_t)n * 8);
    apd_get_lm_trace(handle_, rows.data(), n, &n);
    for (int i = 0; i < n; i++) {
      const double* r = &rows[(size_t)i * 8];
      if (r[1] == 0) std::printf("--- LM optimization ---\n%5s %15s %15s %15s %15s %15s %5s\n", "i", "y0", "yi", "rho", "lambda", "|delta|", "dec");
      std::printf("%5d %15g %15g %15g %15g %15g %5s\n", (int)r[1], r[2], r[3], r[4], r[5], r[6], r[7] != 0 ? "true" : "false");  // :148-154
    }
  }

  apd_handle handle_ = nullptr;
  apd_params params_;
  apd_result last_ = {};
  bool lm_debug_print_ = false;
  Eigen::Matrix<double, 6, 6> final_hessian_;
};

}  // namespace fast_gicp
