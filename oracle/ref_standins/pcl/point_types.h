// TEST INFRASTRUCTURE - stand-in for <pcl/point_types.h>: pcl::PointXYZI with the memory layout and the Eigen map
// accessors of PCL 1.8 / 1.10 (32 bytes: x y z 1 | intensity + padding), as far as the reference's FastAPDGICP sources
// use them. See oracle/ref_standins/Eigen/Core.
#pragma once
#include <Eigen/Core>

namespace pcl {

struct alignas(16) PointXYZI {
  union {
    float data[4];
    struct { float x, y, z, pad_; };   // pad_: the name include/pcl_shim's Registration stand-in uses for data[3]
  };
  union {
    struct { float intensity; };
    float data_c[4];
  };
  PointXYZI() { x = y = z = 0.f; data[3] = 1.f; data_c[0] = data_c[1] = data_c[2] = data_c[3] = 0.f; }
  Eigen::Map<Eigen::Vector4f> getVector4fMap() { return Eigen::Map<Eigen::Vector4f>(data); }
  Eigen::Map<const Eigen::Vector4f> getVector4fMap() const { return Eigen::Map<const Eigen::Vector4f>(data); }
  Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(data); }
  Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(data); }
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");

}  // namespace pcl
