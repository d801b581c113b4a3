// Drop-in for fast_gicp::FastAPDGICP<pcl::PointXYZI, pcl::PointXYZI>
// (reference fast_apdgicp/include/fast_gicp/gicp/fast_apdgicp.hpp:33-110, impl/fast_apdgicp_impl.hpp).
// Same class name, namespace, include path and public surface, so
//   #include <fast_gicp/gicp/fast_apdgicp.hpp>
// in radar_graph_slam/src/radar_graph_slam/registrations.cpp:15,38-50 picks this header up when this
// repository's include/ directory precedes the reference's, and the nodelets keep holding a
// pcl::Registration<PointXYZI, PointXYZI>::Ptr. It is the ONLY header this repository puts on a reference
// include path: <fast_gicp/gicp/gicp_settings.hpp> below is the reference's own file, and the GPU base class lives
// under a name of its own (apdgicp_b200/lsq_registration_b200.hpp), so registrations.cpp:13-14 (the reference's
// fast_gicp.hpp / fast_vgicp.hpp on the reference's LsqRegistration) keeps compiling and linking libfast_gicp.so
// unchanged (tests/cpp/coexist_test.cpp). Header-only: everything forwards to the C ABI of libapdgicp_b200.so (kNN,
// covariances, APD Mahalanobis, H/b reduction and the LM loop all run on the GPU). There is no CPU fallback.
#pragma once
#include <limits>
#include <memory>
#include <vector>

#include <fast_gicp/gicp/gicp_settings.hpp>
#include <apdgicp_b200/lsq_registration_b200.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class FastAPDGICP : public LsqRegistrationB200<PointSource, PointTarget> {
public:
  using Scalar = float;
  using Lsq = LsqRegistrationB200<PointSource, PointTarget>;
  using Matrix4 = typename Lsq::Matrix4;
  using PointCloudSource = typename Lsq::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename Lsq::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using CovarianceList = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;
#if PCL_VERSION >= PCL_VERSION_CALC(1, 10, 0)  // fast_apdgicp.hpp:33-39
  using Ptr = pcl::shared_ptr<FastAPDGICP<PointSource, PointTarget>>;
  using ConstPtr = pcl::shared_ptr<const FastAPDGICP<PointSource, PointTarget>>;
#else
  using Ptr = boost::shared_ptr<FastAPDGICP<PointSource, PointTarget>>;
  using ConstPtr = boost::shared_ptr<const FastAPDGICP<PointSource, PointTarget>>;
#endif

protected:
  using Lsq::handle_;
  using Lsq::params_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::input_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::target_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::corr_dist_threshold_;

public:
  explicit FastAPDGICP(int device = 0) : Lsq(device) {
    this->reg_name_ = "FastAPDGICP";
    corr_dist_threshold_ = std::numeric_limits<float>::max();  // fast_apdgicp_impl.hpp:23
  }
  ~FastAPDGICP() override {}

  // fast_apdgicp_impl.hpp:34-65
  void setNumThreads(int n) { params_.num_threads = n; }  // kept for source compatibility; the GPU path ignores it
  // Deliberate deviation: the reference keeps covariances it has already computed until the CLOUD changes
  // (fast_apdgicp_impl.hpp:122-127 tests only the sizes), so a new k / regularisation silently applies to later clouds only.
  // Here the device recomputes stale covariances at the next align (the result a caller of these setters expects).
  void setCorrespondenceRandomness(int k) { params_.k_correspondences = k; invalidate_covariances(); }
  void setRegularizationMethod(RegularizationMethod method) { params_.regularization = static_cast<int>(method); invalidate_covariances(); }
  void setAzimuthVar(double var) { params_.azimuth_var = var; }
  void setElevationVar(double var) { params_.elevation_var = var; }
  void setDistVar(double var) { params_.dist_var = var; }

  // fast_apdgicp_impl.hpp:68-87
  void swapSourceAndTarget() override {
    input_.swap(target_);
    source_covs_.swap(target_covs_);
    std::swap(src_dirty_, tgt_dirty_);
    std::swap(src_covs_injected_, tgt_covs_injected_);
    std::swap(src_cov_dirty_, tgt_cov_dirty_);  // injected covariances that were not pushed yet travel with their cloud
    if (handle_) apd_swap_source_and_target(handle_);
    this->target_cloud_updated_ = true;  // PCL rebuilds its own tree_ on the next align
  }
  void clearSource() override {
    input_.reset();
    source_covs_.clear();
    src_covs_injected_ = src_cov_dirty_ = src_dirty_ = false;
    if (handle_) apd_clear_source(handle_);
  }
  void clearTarget() override {
    target_.reset();
    target_covs_.clear();
    tgt_covs_injected_ = tgt_cov_dirty_ = tgt_dirty_ = false;
    if (handle_) apd_clear_target(handle_);
  }

  // fast_apdgicp_impl.hpp:90-108: identical pointer -> nothing to do; otherwise the cloud goes to the device AT ONCE
  // (asynchronous upload, where the reference builds its kd-tree) with the pointer as the device cache key, so the scan
  // that was the source of the previous registration keeps its grid and covariances when it becomes the target
  // (scan_matching_odometry_nodelet.cpp:591-592). Pushing eagerly also means the device never holds the key of a cloud
  // this object no longer keeps alive: an address the allocator hands out again can not hit a stale cache entry.
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (input_ == cloud) return;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputSource(cloud);
    source_covs_.clear();
    src_covs_injected_ = src_cov_dirty_ = false;
    src_dirty_ = true;
    push_clouds();
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (target_ == cloud) return;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputTarget(cloud);
    target_covs_.clear();
    tgt_covs_injected_ = tgt_cov_dirty_ = false;
    tgt_dirty_ = true;
    push_clouds();
  }

  // Scan-to-map target built on the device: what ScanMatchingOdometryNodelet does at
  // scan_matching_odometry_nodelet.cpp:606-616 - transform the selected keyframe clouds by rel_pose (double),
  // concatenate, pcl::VoxelGrid with leaf `downsample_resolution` (<= 0: no down-sampling), setInputTarget - in one
  // call on a cloud set that already lives in HBM (apd_cloudset_create with sizeof(PointT) stride). The submap is
  // returned as a PCL cloud (keyframe_cloud_s2m) and registered as the target WITHOUT being uploaded again.
  typename pcl::PointCloud<PointTarget>::Ptr setInputTargetFromKeyframes(apd_cloudset keyframes, const std::vector<int>& which,
                                                                         const std::vector<Eigen::Matrix4d>& rel_poses, double downsample_resolution) {
    typename pcl::PointCloud<PointTarget>::Ptr cloud(new pcl::PointCloud<PointTarget>());
    if (!handle_ || which.size() != rel_poses.size()) return cloud;
    int64_t total = 0;
    if (apd_cloudset_info(keyframes, nullptr, &total) != APD_OK) return cloud;
    std::vector<double> poses(16 * which.size());
    for (size_t k = 0; k < which.size(); k++)
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) poses[16 * k + 4 * r + c] = rel_poses[k](r, c);
    std::vector<float> xyzi(4 * (size_t)std::max<int64_t>(total, 1));
    int n = 0;
    if (apd_build_submap(handle_, keyframes, which.data(), (int)which.size(), poses.data(), downsample_resolution, reinterpret_cast<uint64_t>(cloud.get()),
                         xyzi.data(), (int)total, &n) != APD_OK) {
      std::fprintf(stderr, "[apdgicp_b200] build_submap failed: %s\n", apd_last_error(handle_));
      return cloud;
    }
    cloud->resize(n);
    for (int i = 0; i < n; i++) {
      PointTarget& p = cloud->points[i];
      p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
    }
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputTarget(cloud);  // PCL's own bookkeeping (target_, tree_)
    target_covs_.clear();
    tgt_covs_injected_ = false;
    tgt_dirty_ = false;  // the device already holds this cloud under the key cloud.get()
    return cloud;
  }

  // fast_apdgicp_impl.hpp:111-118
  virtual void setSourceCovariances(const CovarianceList& covs) { source_covs_ = covs; src_covs_injected_ = true; src_cov_dirty_ = true; }
  virtual void setTargetCovariances(const CovarianceList& covs) { target_covs_ = covs; tgt_covs_injected_ = true; tgt_cov_dirty_ = true; }
  // fast_apdgicp.hpp:68-74 (the reference returns the host copies; here they are fetched from the device on demand)
  const CovarianceList& getSourceCovariances() { fetch_covariances(0, input_ ? input_->points.size() : 0, source_covs_); return source_covs_; }
  const CovarianceList& getTargetCovariances() { fetch_covariances(1, target_ ? target_->points.size() : 0, target_covs_); return target_covs_; }

protected:
  // fast_apdgicp.hpp:77-83: the protected virtuals of the reference class, same signatures (linearize and compute_error are
  // inherited from the GPU base with the reference's signatures; computeTransformation likewise)
  virtual void update_correspondences(const Eigen::Isometry3d& trans) { this->linearize(trans, nullptr, nullptr); }

  // upload whichever cloud changed; a cloud that is null / empty clears its device slot
  bool push_clouds() {
    if (!handle_) return false;
    // target first: when it is the previous source the device moves its data across instead of re-uploading
    if (tgt_dirty_) {
      const int rc = (target_ && !target_->points.empty())
                         ? apd_set_target(handle_, reinterpret_cast<const float*>(target_->points.data()), (int)sizeof(PointTarget), (int)target_->points.size(),
                                          reinterpret_cast<uint64_t>(target_.get()), APD_MEM_HOST)
                         : apd_clear_target(handle_);
      if (rc != APD_OK) return false;
      tgt_dirty_ = false;
    }
    if (src_dirty_) {
      const int rc = (input_ && !input_->points.empty())
                         ? apd_set_source(handle_, reinterpret_cast<const float*>(input_->points.data()), (int)sizeof(PointSource), (int)input_->points.size(),
                                          reinterpret_cast<uint64_t>(input_.get()), APD_MEM_HOST)
                         : apd_clear_source(handle_);
      if (rc != APD_OK) return false;
      src_dirty_ = false;
    }
    return true;
  }

  bool sync_inputs() override {
    if (!Lsq::sync_inputs()) return false;
    if (!input_ || !target_ || input_->points.empty() || target_->points.empty()) return false;
    if (!push_clouds()) return false;
    if (src_cov_dirty_ && src_covs_injected_ && source_covs_.size() == input_->points.size()) {
      if (apd_set_covariances(handle_, 0, reinterpret_cast<const double*>(source_covs_.data()), (int)source_covs_.size()) != APD_OK) return false;
      src_cov_dirty_ = false;
    }
    if (tgt_cov_dirty_ && tgt_covs_injected_ && target_covs_.size() == target_->points.size()) {
      if (apd_set_covariances(handle_, 1, reinterpret_cast<const double*>(target_covs_.data()), (int)target_covs_.size()) != APD_OK) return false;
      tgt_cov_dirty_ = false;
    }
    return true;
  }

  void invalidate_covariances() {
    // k / regularisation changed: device covariances are recomputed lazily (apd_set_params marks them stale)
    if (!src_covs_injected_) source_covs_.clear();
    if (!tgt_covs_injected_) target_covs_.clear();
  }

  void fetch_covariances(int which, size_t n, CovarianceList& out) {
    if ((which == 0 ? src_covs_injected_ : tgt_covs_injected_) || !handle_ || n == 0) return;
    if (!sync_inputs()) return;
    out.resize(n);
    // Eigen::Matrix4d is 16 contiguous doubles; the matrices are symmetric, so row/column order is immaterial
    if (apd_get_covariances(handle_, which, reinterpret_cast<double*>(out.data())) != APD_OK) out.clear();
  }

  CovarianceList source_covs_, target_covs_;
  bool src_dirty_ = false, tgt_dirty_ = false;
  bool src_covs_injected_ = false, tgt_covs_injected_ = false;
  bool src_cov_dirty_ = false, tgt_cov_dirty_ = false;
};

}  // namespace fast_gicp
