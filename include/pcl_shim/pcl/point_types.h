// IN-CONTAINER STAND-IN for <pcl/point_types.h>: the memory layout of pcl::PointXYZI (32 bytes:
// x y z pad(=1) | intensity + 3 pad). PCL is not installed in the build image.
#pragma once

namespace pcl {

struct alignas(16) PointXYZ {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
};

struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, pad_ = 1.f;
  float intensity = 0, pad2_[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");

}  // namespace pcl
