// TEST INFRASTRUCTURE - stand-in for pcl::transformPointCloud (pcl/common/impl/transforms.hpp): every point is copied
// with all its fields and its xyz replaced by the float 4x4 transform of (x, y, z, 1), summed in the order
// ((m0 x + m1 y) + m2 z) + m3 - PCL 1.10's scalar Transformer::se3. See oracle/ref_standins/Eigen/Core.
#pragma once
#include <Eigen/Core>
#include <pcl/point_cloud.h>

namespace pcl {

template <typename PointT, typename Scalar>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<Scalar, 4, 4>& tf) {
  if (&in != &out) {
    out.points = in.points;
    out.width = in.width;
    out.height = in.height;
    out.is_dense = in.is_dense;
  }
  for (std::size_t i = 0; i < out.points.size(); i++) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = (float)(((tf(0, 0) * x + tf(0, 1) * y) + tf(0, 2) * z) + tf(0, 3));
    out.points[i].y = (float)(((tf(1, 0) * x + tf(1, 1) * y) + tf(1, 2) * z) + tf(1, 3));
    out.points[i].z = (float)(((tf(2, 0) * x + tf(2, 1) * y) + tf(2, 2) * z) + tf(2, 3));
    out.points[i].data[3] = 1.f;
  }
}

}  // namespace pcl
