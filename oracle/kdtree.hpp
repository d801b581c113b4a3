// TEST INFRASTRUCTURE — CPU oracle for the FastAPDGICP hot path. Not part of the product.
//
// Exact k-nearest-neighbour search standing in for pcl::search::KdTree<PointXYZI> (FLANN
// KDTreeSingleIndex, epsilon = 0, sorted results, L2_Simple<float>), which the reference calls at
// fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:96,106,151,316. PCL and FLANN are
// third-party dependencies absent from /root/reference; their published behaviour is restated:
//   d2 = ((dx*dx + dy*dy) + dz*dz) accumulated in float (L2_Simple), no FMA contraction
//   results ascending; k clamped to the cloud size
// FLANN's order among exactly equal distances depends on tree traversal; the convention fixed for
// this project (BASELINE.json north_star) is ascending (d2, index).
//
// Non-finite points: pcl::KdTreeFLANN::convertCloudToArray (PCL 1.8-1.10 kdtree_flann.hpp, not under
// /root/reference) leaves every point whose coordinates are not all finite out of the index, so such a
// point is never anybody's neighbour; a query with a non-finite coordinate only produces NaN distances,
// which never beat a result-set entry in FLANN, i.e. it finds nothing. Both are restated here: build()
// skips non-finite points and knn() / knn_bruteforce() return an EMPTY result for a non-finite query.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace apd_oracle {

struct P3 {
  float x, y, z;
};

inline float sqdist(const P3& a, const P3& b) {
  const float dx = a.x - b.x;
  const float dy = a.y - b.y;
  const float dz = a.z - b.z;
  float r = dx * dx;  // build with -ffp-contract=off: three roundings, like L2_Simple<float>
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

// Bounded, sorted candidate list ordered by (d2, index).
struct KBest {
  int k;
  int count = 0;
  std::vector<float> d2;
  std::vector<int> idx;
  explicit KBest(int k_) : k(k_), d2(k_), idx(k_) {}
  bool full() const { return count == k; }
  float worst_d2() const { return d2[count - 1]; }
  int worst_idx() const { return idx[count - 1]; }
  static bool less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }
  void offer(float d, int i) {
    if (full() && !less(d, i, d2[k - 1], idx[k - 1])) return;
    int pos = full() ? k - 1 : count++;
    while (pos > 0 && less(d, i, d2[pos - 1], idx[pos - 1])) {
      d2[pos] = d2[pos - 1];
      idx[pos] = idx[pos - 1];
      pos--;
    }
    d2[pos] = d;
    idx[pos] = i;
  }
};

inline bool finite_pt(const P3& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

inline void knn_bruteforce(const std::vector<P3>& pts, const P3& q, int k, std::vector<int>& out_idx, std::vector<float>& out_d2) {
  int n_finite = 0;
  for (const P3& p : pts) n_finite += finite_pt(p) ? 1 : 0;
  k = std::min<int>(k, n_finite);
  if (!finite_pt(q)) k = 0;
  KBest best(k);
  for (int i = 0; i < (int)pts.size() && k > 0; i++)
    if (finite_pt(pts[i])) best.offer(sqdist(q, pts[i]), i);
  out_idx.assign(best.idx.begin(), best.idx.begin() + best.count);
  out_d2.assign(best.d2.begin(), best.d2.begin() + best.count);
}

class KdTree {
public:
  void build(const std::vector<P3>* pts) {
    pts_ = pts;
    order_.clear();
    order_.reserve(pts->size());
    for (int i = 0; i < (int)pts->size(); i++)
      if (finite_pt((*pts)[i])) order_.push_back(i);  // convertCloudToArray: invalid points are not indexed
    const int n = (int)order_.size();
    nodes_.clear();
    nodes_.reserve(n / 4 + 16);
    if (n > 0) build_rec(0, n);
  }
  void rebind(const std::vector<P3>* pts) { pts_ = pts; }  // same points, moved to another vector
  bool empty() const { return pts_ == nullptr || pts_->empty(); }
  const std::vector<P3>* cloud() const { return pts_; }

  void knn(const P3& q, int k, std::vector<int>& out_idx, std::vector<float>& out_d2) const {
    k = std::min<int>(k, (int)order_.size());
    if (!finite_pt(q)) k = 0;  // NaN distances never enter FLANN's result set
    KBest best(k);
    if (k > 0) search(0, q, best);
    out_idx.assign(best.idx.begin(), best.idx.begin() + best.count);
    out_d2.assign(best.d2.begin(), best.d2.begin() + best.count);
  }

private:
  struct Node {
    int lo, hi;        // leaf: range in order_
    int left, right;   // children (or -1)
    int dim;
    float split;
  };
  static constexpr int kLeaf = 12;

  int build_rec(int lo, int hi) {
    const int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, 0, 0.f});
    if (hi - lo <= kLeaf) return id;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; i++) {
      const P3& p = (*pts_)[order_[i]];
      const float c[3] = {p.x, p.y, p.z};
      for (int d = 0; d < 3; d++) {
        mn[d] = std::min(mn[d], c[d]);
        mx[d] = std::max(mx[d], c[d]);
      }
    }
    int dim = 0;
    for (int d = 1; d < 3; d++)
      if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    if (!(mx[dim] > mn[dim])) return id;  // all points identical: keep as one leaf
    const int mid = (lo + hi) / 2;
    auto coord = [&](int i) {
      const P3& p = (*pts_)[i];
      return dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
    };
    std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi, [&](int a, int b) { return coord(a) < coord(b); });
    const float split = coord(order_[mid]);
    const int l = build_rec(lo, mid);
    const int r = build_rec(mid, hi);
    nodes_[id].left = l;
    nodes_[id].right = r;
    nodes_[id].dim = dim;
    nodes_[id].split = split;
    return id;
  }

  void search(int id, const P3& q, KBest& best) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; i++) {
        const int pi = order_[i];
        best.offer(sqdist(q, (*pts_)[pi]), pi);
      }
      return;
    }
    const float qc = nd.dim == 0 ? q.x : (nd.dim == 1 ? q.y : q.z);
    const double diff = (double)qc - (double)nd.split;
    const int near = diff < 0 ? nd.left : nd.right;
    const int far = diff < 0 ? nd.right : nd.left;
    search(near, q, best);
    // Every point of the far side is at least |diff| away along this axis. Prune only when that
    // bound clears the current k-th distance by a margin larger than float rounding of d2, so that
    // ties and last-ulp effects can never change the exact (d2, index) result.
    if (!best.full() || diff * diff <= (double)best.worst_d2() * (1.0 + 1e-5) + 1e-30) search(far, q, best);
  }

  const std::vector<P3>* pts_ = nullptr;
  std::vector<int> order_;
  std::vector<Node> nodes_;
};

}  // namespace apd_oracle
