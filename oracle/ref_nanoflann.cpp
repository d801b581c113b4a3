// TEST INFRASTRUCTURE — thin C wrapper around the exact kd-tree the REFERENCE ITSELF vendors:
// radar_graph_slam/include/scan_context/nanoflann.hpp (nanoflann 1.3.2, header-only, STL-only), compiled
// from where it lies under /root/reference (include path only; nothing is copied into this repository).
// It is the one piece of reference code on the nearest-neighbour path that builds in this image (PCL / FLANN do
// not exist here), so it pins the search of the oracle: same metric arithmetic as FLANN's L2_Simple
// (nanoflann.hpp:432-440: result += diff * diff per dimension, float), exact search (eps = 0), results sorted by
// distance; NANOFLANN_FIRST_MATCH (nanoflann.hpp:178-182) orders equal distances by index inside the result set.
// Built by `make -C oracle ref` into oracle/_ref/libref_nanoflann.so (git-ignored).
#define NANOFLANN_FIRST_MATCH
#include <scan_context/nanoflann.hpp>

#include <cstddef>
#include <vector>

namespace {
struct Cloud {
  const float* p;
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return p[idx * 3 + dim]; }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};
using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Cloud>, Cloud, 3, int>;
}  // namespace

extern "C" {

// k nearest neighbours of m queries (xyz packed) in a cloud of n points (xyz packed); idx / d2 are m x k, ascending.
// Returns the number of neighbours found per query (min(k, n)).
int ref_nanoflann_knn(const float* pts, int n, const float* queries, int m, int k, int leaf_size, int* idx, float* d2) {
  Cloud cloud{pts, (size_t)n};
  Tree tree(3, cloud, nanoflann::KDTreeSingleIndexAdaptorParams(leaf_size > 0 ? leaf_size : 10));
  tree.buildIndex();
  const int kk = k < n ? k : n;
  std::vector<int> i(kk);
  std::vector<float> d(kk);
  for (int q = 0; q < m; q++) {
    nanoflann::KNNResultSet<float, int> rs(kk);
    rs.init(i.data(), d.data());
    tree.findNeighbors(rs, queries + (size_t)q * 3, nanoflann::SearchParams(32, 0.f, true));
    for (int j = 0; j < k; j++) {
      idx[(size_t)q * k + j] = j < kk ? i[j] : -1;
      d2[(size_t)q * k + j] = j < kk ? d[j] : 0.f;
    }
  }
  return kk;
}

const char* ref_nanoflann_version() { return "nanoflann 1.3.2 (radar_graph_slam/include/scan_context/nanoflann.hpp), NANOFLANN_FIRST_MATCH"; }
}
