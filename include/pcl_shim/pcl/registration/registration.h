// IN-CONTAINER STAND-IN for pcl::Registration<PointSource, PointTarget, Scalar>: the members,
// virtuals and the align()/getFitnessScore() behaviour the RIV-SLAM callers observe, restated from
// PCL 1.10 registration.h / registration.hpp as documented in SURVEY.md Appendix B (PCL itself is not
// installed in the build image). Used by tests/cpp to compile and run the drop-in class through a
// base-class pointer exactly like scan_matching_odometry_nodelet.cpp:827-828 holds it.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/search/kdtree.h>

namespace pcl {

template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using Ptr = std::shared_ptr<Registration<PointSource, PointTarget, Scalar>>;
  using ConstPtr = std::shared_ptr<const Registration<PointSource, PointTarget, Scalar>>;
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

  Registration() : tree_(new KdTree) {
    final_transformation_.setIdentity();
    transformation_.setIdentity();
    previous_transformation_.setIdentity();
  }
  virtual ~Registration() {}

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; source_cloud_updated_ = true; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; target_cloud_updated_ = true; }
  PointCloudSourceConstPtr const getInputSource() { return input_; }
  PointCloudTargetConstPtr const getInputTarget() { return target_; }
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) {
    tree_ = tree;
    force_no_recompute_ = force_no_recompute;
    target_cloud_updated_ = true;
  }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  int getMaximumIterations() { return max_iterations_; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  double getTransformationEpsilon() { return transformation_epsilon_; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  double getMaxCorrespondenceDistance() { return corr_dist_threshold_; }
  bool hasConverged() const { return converged_; }

  // mean squared 1-NN distance of the transformed input (squared distance compared with max_range)
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    if (!input_ || !target_) return std::numeric_limits<double>::max();
    double sum = 0.0;
    int nr = 0;
    std::vector<int> idx(1);
    std::vector<float> sq(1);
    for (std::size_t i = 0; i < input_->size(); i++) {
      const PointSource& a = input_->points[i];
      PointTarget q;
      const Matrix4& T = final_transformation_;
      q.x = ((T(0, 0) * a.x + T(0, 1) * a.y) + T(0, 2) * a.z) + T(0, 3);
      q.y = ((T(1, 0) * a.x + T(1, 1) * a.y) + T(1, 2) * a.z) + T(1, 3);
      q.z = ((T(2, 0) * a.x + T(2, 1) * a.y) + T(2, 2) * a.z) + T(2, 3);
      tree_->nearestKSearch(q, 1, idx, sq);
      if (idx.empty()) continue;
      if (sq[0] <= max_range) { sum += sq[0]; nr++; }
    }
    return nr > 0 ? sum / nr : std::numeric_limits<double>::max();
  }

  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!initCompute()) return;
    output.points.resize(input_->points.size());
    output.width = input_->width;
    output.height = input_->height;
    output.is_dense = input_->is_dense;
    for (std::size_t i = 0; i < input_->points.size(); i++) {
      output.points[i] = input_->points[i];  // every field, intensity included
      output.points[i].pad_ = 1.0f;          // data[3] = 1
    }
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }

protected:
  bool initCompute() {
    if (!target_) {
      std::fprintf(stderr, "[pcl::registration::%s::compute] No input target dataset was given!\n", reg_name_.c_str());
      return false;
    }
    if (!input_) return false;
    if (target_cloud_updated_ && !force_no_recompute_) {
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
    return true;
  }
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;

  std::string reg_name_ = "Registration";
  KdTreePtr tree_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0;
  int max_iterations_ = 10;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = std::sqrt(std::numeric_limits<double>::max());
  bool converged_ = false;
  bool target_cloud_updated_ = true, source_cloud_updated_ = true, force_no_recompute_ = false;
};

}  // namespace pcl
