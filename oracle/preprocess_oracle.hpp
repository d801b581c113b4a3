// TEST INFRASTRUCTURE — CPU oracle for the "next" rows SURVEY.md §8(f)-2 and §8(f)-4. Not part of the product.
//
// Restates, from the cited reference lines and from the published behaviour of the PCL 1.10 filters
// they call (PCL is a third-party dependency that is absent from /root/reference and from this image;
// PARITY UNPINNED: nothing in the reference pins these outputs):
//   distance_filter             radar_graph_slam/apps/preprocessing_nodelet.cpp:880-896
//   pcl::VoxelGrid<PointXYZI>   (preprocessing_nodelet.cpp:137-144,850-866; scan_matching_odometry_nodelet.cpp:148-155)
//                               pcl/filters/impl/voxel_grid.hpp applyFilter, downsample_all_data_ = true, no filter field
//   pcl::RadiusOutlierRemoval   (preprocessing_nodelet.cpp:176-184,868-878) pcl/filters/impl/radius_outlier_removal.hpp,
//                               dense-input branch (the voxel filter marks its output dense): nearest-k with k = min_pts + 1
//   submap accumulation         radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:606-616
//                               (pcl::transformPointCloud with a double transform, clouds concatenated, then downsample())
// Conventions where PCL leaves the result to the implementation: points of one voxel are accumulated in
// ascending input order (std::sort on the voxel index alone is not stable); float sums round after every
// operation (no FMA), like the rest of the oracle.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace apd_oracle {

struct PointI {
  float x, y, z, intensity;
};

inline bool finite3(const PointI& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

// preprocessing_nodelet.cpp:884-889: d = p.getVector3fMap().norm() (float), compared as double
inline std::vector<PointI> distance_filter(const std::vector<PointI>& in, double near_thresh, double far_thresh, double z_low, double z_high) {
  std::vector<PointI> out;
  out.reserve(in.size());
  for (const PointI& p : in) {
    float s = p.x * p.x;
    s = s + p.y * p.y;
    s = s + p.z * p.z;
    const double d = (double)std::sqrt(s);
    const double z = (double)p.z;
    if (d > near_thresh && d < far_thresh && z < z_high && z > z_low) out.push_back(p);
  }
  return out;
}

// pcl::VoxelGrid::applyFilter. Returns false (and copies the input) when the leaf is too small for 32-bit voxel indices.
inline bool voxel_grid(const std::vector<PointI>& in, float leaf, std::vector<PointI>& out) {
  out.clear();
  const float inv = 1.0f / leaf;
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  bool any = false;
  for (const PointI& p : in) {
    if (!finite3(p)) continue;
    const float c[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], c[a]);
      mx[a] = std::max(mx[a], c[a]);
    }
    any = true;
  }
  if (!any) return true;
  const std::int64_t dx = (std::int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (std::int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (std::int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (std::int64_t)std::numeric_limits<std::int32_t>::max()) {
    out = in;
    return false;
  }
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; a++) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    div_b[a] = (int)std::floor(mx[a] * inv) - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<unsigned, unsigned>> iv;  // (voxel index, point index)
  iv.reserve(in.size());
  for (unsigned i = 0; i < in.size(); i++) {
    const PointI& p = in[i];
    if (!finite3(p)) continue;
    const int i0 = (int)(std::floor(p.x * inv) - (float)min_b[0]);
    const int i1 = (int)(std::floor(p.y * inv) - (float)min_b[1]);
    const int i2 = (int)(std::floor(p.z * inv) - (float)min_b[2]);
    iv.emplace_back((unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), i);
  }
  std::sort(iv.begin(), iv.end());  // (voxel, point index): ascending input order inside a voxel (convention)
  size_t a = 0;
  while (a < iv.size()) {
    size_t b = a;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    while (b < iv.size() && iv[b].first == iv[a].first) {
      const PointI& p = in[iv[b].second];
      sx = sx + p.x; sy = sy + p.y; sz = sz + p.z; si = si + p.intensity;  // AccumulatorXYZ / AccumulatorIntensity
      b++;
    }
    const float n = (float)(b - a);
    out.push_back(PointI{sx / n, sy / n, sz / n, si / n});
    a = b;
  }
  return true;
}

// pcl::ApproximateVoxelGrid::applyFilter (pcl/filters/impl/approximate_voxel_grid.hpp, PCL 1.10; the APPROX_VOXELGRID branch of
// preprocessing_nodelet.cpp:145-149 and scan_matching_odometry_nodelet.cpp:156-160), restated as PCL writes it: one pass over the
// points with a history table of histsize_ = 512 entries indexed by (ix * 7171 + iy * 3079 + iz * 4231) & 511; a point whose voxel
// differs from the one stored in its entry flushes that entry (centroid / count becomes the next output point) and restarts it; the
// non-empty entries are flushed in table order at the end. All four fields are averaged (downsample_all_data_). PARITY UNPINNED:
// PCL is absent; the table size and hash constants are quoted from PCL 1.10's header. ix = static_cast<int>(floor(x * inv)) of a
// non-finite / out-of-range value is taken as INT_MIN (what cvttss2si yields on x86; undefined in C++).
inline int approx_cell(float v, float inv) {
  const float f = std::floor(v * inv);
  return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : std::numeric_limits<int>::min();
}
inline std::vector<PointI> approx_voxel_grid(const std::vector<PointI>& in, float leaf) {
  const unsigned histsize = 512;
  struct He { int ix, iy, iz, count; float c[4]; };
  std::vector<He> history(histsize, He{0, 0, 0, 0, {0.f, 0.f, 0.f, 0.f}});
  const float inv = 1.0f / leaf;
  std::vector<PointI> out;
  auto flush = [&](He& h) {
    const float n = (float)h.count;
    out.push_back(PointI{h.c[0] / n, h.c[1] / n, h.c[2] / n, h.c[3] / n});
  };
  for (const PointI& p : in) {
    const int ix = approx_cell(p.x, inv), iy = approx_cell(p.y, inv), iz = approx_cell(p.z, inv);
    const unsigned hash = ((unsigned)ix * 7171u + (unsigned)iy * 3079u + (unsigned)iz * 4231u) & (histsize - 1);
    He& h = history[hash];
    if (h.count && (ix != h.ix || iy != h.iy || iz != h.iz)) {
      flush(h);
      h.count = 0;
      h.c[0] = h.c[1] = h.c[2] = h.c[3] = 0.f;
    }
    h.ix = ix; h.iy = iy; h.iz = iz;
    h.count++;
    h.c[0] = h.c[0] + p.x; h.c[1] = h.c[1] + p.y; h.c[2] = h.c[2] + p.z; h.c[3] = h.c[3] + p.intensity;
  }
  for (He& h : history)
    if (h.count) flush(h);
  return out;
}

// pcl::RadiusOutlierRemoval on dense input: keep a point iff its (min_pts + 1)-th nearest point (itself included)
// exists and lies within the radius: !(r*r < d2), d2 = L2_Simple<float>
inline std::vector<PointI> radius_outlier_removal(const std::vector<PointI>& in, double radius, int min_pts) {
  std::vector<PointI> out;
  const int mean_k = min_pts + 1;
  const double r2 = radius * radius;
  std::vector<float> d2(in.size());
  for (size_t i = 0; i < in.size(); i++) {
    for (size_t j = 0; j < in.size(); j++) {
      const float dx = in[i].x - in[j].x, dy = in[i].y - in[j].y, dz = in[i].z - in[j].z;
      float s = dx * dx;
      s = s + dy * dy;
      s = s + dz * dz;
      d2[j] = s;
    }
    if ((int)in.size() < mean_k) continue;  // k != mean_k: removed
    std::nth_element(d2.begin(), d2.begin() + (mean_k - 1), d2.end());
    if (!(r2 < (double)d2[mean_k - 1])) out.push_back(in[i]);
  }
  return out;
}

// pcl::StatisticalOutlierRemoval (preprocessing_nodelet.cpp:165-174, the nodelet's code default when the rosparam is absent;
// pcl/filters/impl/statistical_outlier_removal.hpp applyFilterIndices, PCL 1.10): mean distance of every point to its
// mean_k nearest neighbours (the search asks for mean_k + 1 and skips the point itself), then
// threshold = mean + stddev_mult * stddev over all points; a point is removed when its mean distance exceeds it.
// Arithmetic as written there: dist_sum (double) += sqrt(nn_dists[k]) with nn_dists float (unqualified sqrt on a float
// promotes to double with <cmath> alone, the convention taken here), distances[i] = float(dist_sum / mean_k),
// sum (double) += distance, sq_sum (double) += distance * distance (a float product), both in index order.
// Clouds with fewer than mean_k + 1 points are returned unchanged (the search cannot deliver mean_k neighbours).
inline std::vector<PointI> statistical_outlier_removal(const std::vector<PointI>& in, int mean_k, double stddev_mult) {
  const size_t n = in.size();
  if (mean_k < 1 || n < (size_t)mean_k + 1) return in;
  std::vector<float> distances(n, 0.f);
  std::vector<float> d2(n);
  int valid = 0;
  for (size_t i = 0; i < n; i++) {
    if (!finite3(in[i])) continue;
    for (size_t j = 0; j < n; j++) {
      const float dx = in[i].x - in[j].x, dy = in[i].y - in[j].y, dz = in[i].z - in[j].z;
      float s = dx * dx;
      s = s + dy * dy;
      s = s + dz * dz;
      d2[j] = s;
    }
    std::partial_sort(d2.begin(), d2.begin() + mean_k + 1, d2.end());
    double dist_sum = 0.0;
    for (int k = 1; k < mean_k + 1; k++) dist_sum += std::sqrt((double)d2[k]);  // k = 0 is the query point
    distances[i] = (float)(dist_sum / mean_k);
    valid++;
  }
  double sum = 0.0, sq_sum = 0.0;
  for (const float d : distances) {
    sum += d;
    sq_sum += d * d;
  }
  const double mean = sum / (double)valid;
  const double variance = (sq_sum - sum * sum / (double)valid) / ((double)valid - 1);
  const double threshold = mean + stddev_mult * std::sqrt(variance);
  std::vector<PointI> out;
  for (size_t i = 0; i < n; i++)
    if (!((double)distances[i] > threshold)) out.push_back(in[i]);
  return out;
}

// scan_matching_odometry_nodelet.cpp:606-616: keyframe clouds moved by rel_pose (double 4x4, row-major here) and concatenated
inline std::vector<PointI> accumulate_submap(const std::vector<std::vector<PointI>>& clouds, const std::vector<const double*>& rel_poses) {
  std::vector<PointI> out;
  for (size_t k = 0; k < clouds.size(); k++) {
    const double* T = rel_poses[k];
    for (const PointI& p : clouds[k]) {
      // pcl::transformPointCloud<PointT, double>: double arithmetic, result stored as float; intensity copied
      const double x = p.x, y = p.y, z = p.z;
      PointI q;
      q.x = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
      q.y = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
      q.z = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
      q.intensity = p.intensity;
      out.push_back(q);
    }
  }
  return out;
}

}  // namespace apd_oracle
