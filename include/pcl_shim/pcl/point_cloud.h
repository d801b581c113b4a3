// IN-CONTAINER STAND-IN for <pcl/point_cloud.h>.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

// pcl/pcl_config.h: the reference headers switch their Ptr aliases on these (fast_apdgicp.hpp:33-39)
#ifndef PCL_VERSION_CALC
#define PCL_VERSION_CALC(MAJ, MIN, PATCH) (MAJ * 100000 + MIN * 100 + PATCH)
#define PCL_VERSION PCL_VERSION_CALC(1, 10, 0)
#endif

namespace pcl {

template <typename T> using shared_ptr = std::shared_ptr<T>;

template <typename PointT>
class PointCloud {
public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};

}  // namespace pcl
