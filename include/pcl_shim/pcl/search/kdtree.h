// IN-CONTAINER STAND-IN for pcl::search::KdTree: exact nearest neighbours by brute force, enough for
// the base-class getFitnessScore of the registration stand-in. Test infrastructure only.
#pragma once
#include <algorithm>
#include <vector>

#include <pcl/point_cloud.h>

namespace pcl {
namespace search {

template <typename PointT>
class KdTree {
public:
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  void setInputCloud(const PointCloudConstPtr& cloud) { cloud_ = cloud; }
  PointCloudConstPtr getInputCloud() const { return cloud_; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& sq) const {
    idx.clear();
    sq.clear();
    if (!cloud_) return 0;
    std::vector<std::pair<float, int>> all;
    all.reserve(cloud_->size());
    for (std::size_t i = 0; i < cloud_->size(); i++) {
      const PointT& p = cloud_->points[i];
      const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
      float d = dx * dx;
      d = d + dy * dy;
      d = d + dz * dz;
      all.emplace_back(d, (int)i);
    }
    k = std::min<int>(k, (int)all.size());
    std::partial_sort(all.begin(), all.begin() + k, all.end());
    for (int j = 0; j < k; j++) { idx.push_back(all[j].second); sq.push_back(all[j].first); }
    return k;
  }

private:
  PointCloudConstPtr cloud_;
};

}  // namespace search
}  // namespace pcl
