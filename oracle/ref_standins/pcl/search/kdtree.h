// TEST INFRASTRUCTURE - stand-in for pcl::search::KdTree<PointT> (a FLANN KDTreeSingleIndex behind PCL's wrapper) on top
// of the exact kd-tree the reference itself vendors: radar_graph_slam/include/scan_context/nanoflann.hpp (nanoflann 1.3.2,
// the header-only descendant of that FLANN index), included from where it lies under /root/reference. Same float metric
// as FLANN's L2_Simple (result += diff * diff per dimension), exact search, results sorted by distance; equal distances
// are ordered by index (NANOFLANN_FIRST_MATCH), the convention of SURVEY 8c. See oracle/ref_standins/Eigen/Core.
#pragma once
#ifndef NANOFLANN_FIRST_MATCH
#define NANOFLANN_FIRST_MATCH
#endif
#include <scan_context/nanoflann.hpp>

#include <memory>
#include <vector>

#include <pcl/point_cloud.h>

namespace pcl {
namespace search {

template <typename PointT>
class KdTree {
  struct Adaptor {
    const PointCloud<PointT>* c = nullptr;
    inline size_t kdtree_get_point_count() const { return c->points.size(); }
    inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return c->points[idx].data[dim]; }
    template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
  };
  typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Adaptor>, Adaptor, 3, int> Index;

public:
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  void setInputCloud(const PointCloudConstPtr& cloud) {
    cloud_ = cloud;
    index_.reset();
    if (!cloud_ || cloud_->points.empty()) return;
    adaptor_.c = cloud_.get();
    index_.reset(new Index(3, adaptor_, nanoflann::KDTreeSingleIndexAdaptorParams(15)));   // pcl::KdTreeFLANN: 15 points per leaf
    index_->buildIndex();
  }
  PointCloudConstPtr getInputCloud() const { return cloud_; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& sq) const {
    const int n = cloud_ ? (int)cloud_->points.size() : 0;
    if (k > n) k = n;   // pcl::KdTreeFLANN::nearestKSearch: "if (k > total_nr_points_) k = total_nr_points_"
    idx.resize(k);
    sq.resize(k);
    if (k == 0) return 0;
    nanoflann::KNNResultSet<float, int> rs(k);
    rs.init(idx.data(), sq.data());
    index_->findNeighbors(rs, q.data, nanoflann::SearchParams(32, 0.f, true));
    return k;
  }

private:
  PointCloudConstPtr cloud_;
  Adaptor adaptor_;
  std::unique_ptr<Index> index_;
};

}  // namespace search
}  // namespace pcl
