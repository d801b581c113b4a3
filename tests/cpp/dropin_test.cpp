// Runs the drop-in fast_gicp::FastAPDGICP through a pcl::Registration base pointer, configured the
// way the reference factory does it (radar_graph_slam/src/radar_graph_slam/registrations.cpp:38-50
// with the launch-file values) and driven the way scan_matching_odometry_nodelet.cpp:437-482 and
// loop_detector.cpp:222-236 drive it. Input: a binary file of float32 PointXYZI-layout scans written by
// tests/test_cpp_dropin.py. Output: one text line per registration.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <fast_gicp/gicp/fast_apdgicp.hpp>

using PointT = pcl::PointXYZI;
using Cloud = pcl::PointCloud<PointT>;

static pcl::Registration<PointT, PointT>::Ptr select_registration_method() {
  auto gicp = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
  gicp->setNumThreads(0);
  gicp->setTransformationEpsilon(0.1);
  gicp->setMaximumIterations(64);
  gicp->setMaxCorrespondenceDistance(2.0);
  gicp->setCorrespondenceRandomness(20);
  gicp->setDistVar(0.86);
  gicp->setAzimuthVar(1.0);
  gicp->setElevationVar(1.0);
  return gicp;
}

// the protected virtual surface of the reference class (fast_apdgicp.hpp:77-83, lsq_registration.hpp:64-76), reached the way a
// subclass reaches it
struct Probe : fast_gicp::FastAPDGICP<PointT, PointT> {
  using fast_gicp::FastAPDGICP<PointT, PointT>::compute_error;
  using fast_gicp::FastAPDGICP<PointT, PointT>::is_converged;
  using fast_gicp::FastAPDGICP<PointT, PointT>::linearize;
  using fast_gicp::FastAPDGICP<PointT, PointT>::update_correspondences;
};

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int n_scans = 0;
  if (std::fread(&n_scans, sizeof(int), 1, f) != 1) return 2;
  std::vector<Cloud::Ptr> scans;
  for (int s = 0; s < n_scans; s++) {
    int n = 0;
    if (std::fread(&n, sizeof(int), 1, f) != 1) return 2;
    auto c = std::make_shared<Cloud>();
    c->resize(n);
    if (std::fread(c->points.data(), sizeof(PointT), n, f) != (size_t)n) return 2;
    scans.push_back(c);
  }
  std::fclose(f);

  pcl::Registration<PointT, PointT>::Ptr registration = select_registration_method();
  Cloud::Ptr aligned(new Cloud());
  // scan-to-scan odometry: the previous source becomes the target (SMO:591-592)
  registration->setInputTarget(scans[0]);
  for (int t = 1; t < n_scans; t++) {
    registration->setInputSource(scans[t]);
    registration->align(*aligned);
    const auto T = registration->getFinalTransformation();
    const double fit = registration->getFitnessScore();  // the non-virtual PCL implementation on tree_
    auto* apd = dynamic_cast<fast_gicp::FastAPDGICP<PointT, PointT>*>(registration.get());
    std::printf("pair %d converged %d T", t - 1, registration->hasConverged() ? 1 : 0);
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) std::printf(" %.9g", T(r, c));
    std::printf(" fitness_pcl %.17g fitness_gpu %.17g out0 %.9g %.9g %.9g %.9g n_out %zu\n", fit, apd->lastFitnessScore(), aligned->points[0].x,
                aligned->points[0].y, aligned->points[0].z, aligned->points[0].intensity, aligned->size());
    registration->setInputTarget(scans[t]);
  }
  // the FastAPDGICP-specific surface through the derived type (fast_apdgicp.hpp:51-74, lsq_registration.hpp:51-57)
  {
    auto gicp = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
    gicp->setMaxCorrespondenceDistance(2.0);
    gicp->setTransformationEpsilon(1e-6);
    gicp->setRotationEpsilon(1e-6);
    gicp->setAzimuthVar(1.0);
    gicp->setRegularizationMethod(fast_gicp::RegularizationMethod::PLANE);
    gicp->setInputTarget(scans[0]);
    gicp->setInputSource(scans[1]);
    Eigen::Matrix<double, 6, 6> H;
    Eigen::Matrix<double, 6, 1> b;
    const double e = gicp->evaluateCost(Eigen::Matrix4f::Identity(), &H, &b);
    std::printf("cost %.17g H00 %.17g H35 %.17g H53 %.17g b0 %.17g b5 %.17g\n", e, H(0, 0), H(3, 5), H(5, 3), b(0), b(5));
    gicp->align(*aligned);
    const auto Tf = gicp->getFinalTransformation();
    const auto& Hf = gicp->getFinalHessian();
    std::printf("fwd converged %d T03 %.9g T13 %.9g Hf00 %.17g\n", gicp->hasConverged() ? 1 : 0, Tf(0, 3), Tf(1, 3), Hf(0, 0));
    // covariances out, swap, covariances in: the swapped object must reproduce a fresh backward registration
    const auto cs = gicp->getSourceCovariances();
    const auto ct = gicp->getTargetCovariances();
    std::printf("covs %zu %zu c0 %.17g %.17g %.17g\n", cs.size(), ct.size(), cs[0](0, 0), cs[0](1, 0), cs[0](3, 3));
    gicp->swapSourceAndTarget();
    gicp->align(*aligned);
    const auto Tb = gicp->getFinalTransformation();
    auto fresh2 = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
    fresh2->setMaxCorrespondenceDistance(2.0);
    fresh2->setTransformationEpsilon(1e-6);
    fresh2->setRotationEpsilon(1e-6);
    fresh2->setAzimuthVar(1.0);
    fresh2->setInputSource(scans[0]);
    fresh2->setInputTarget(scans[1]);
    fresh2->setSourceCovariances(ct);   // injected: what the swapped object holds for the same clouds
    fresh2->setTargetCovariances(cs);
    fresh2->align(*aligned);
    const auto Tb2 = fresh2->getFinalTransformation();
    double dmax = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) dmax = std::max(dmax, (double)std::fabs(Tb(r, c) - Tb2(r, c)));
    std::printf("bwd converged %d %d maxdiff %.3g T03 %.9g\n", gicp->hasConverged() ? 1 : 0, fresh2->hasConverged() ? 1 : 0, dmax, Tb(0, 3));
    gicp->clearSource();
    gicp->align(*aligned);  // no source any more: PCL returns from initCompute (converged_ keeps its previous value)
    std::printf("cleared converged %d\n", gicp->hasConverged() ? 1 : 0);
  }
  // scan-to-map: the submap of the first three scans, built on the device, becomes the target (SMO:606-616)
  if (n_scans >= 4) {
    auto gicp = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
    gicp->setMaxCorrespondenceDistance(2.0);
    gicp->setTransformationEpsilon(0.1);
    gicp->setAzimuthVar(1.0);
    std::vector<PointT> all;
    std::vector<int32_t> off(1, 0);
    for (int s = 0; s < 3; s++) {
      all.insert(all.end(), scans[s]->points.begin(), scans[s]->points.end());
      off.push_back((int32_t)all.size());
    }
    apd_cloudset ks = nullptr;
    if (apd_cloudset_create(gicp->nativeHandle(), reinterpret_cast<const float*>(all.data()), (int)sizeof(PointT), off.data(), 3, APD_MEM_HOST, &ks) != APD_OK) return 3;
    std::vector<Eigen::Matrix4d> rel(3, Eigen::Matrix4d::Identity());
    rel[0](0, 3) = 0.30; rel[0](1, 3) = -0.02;  // made-up keyframe offsets; the Python side uses the same numbers
    rel[1](0, 3) = 0.15; rel[1](1, 3) = 0.01;
    auto submap = gicp->setInputTargetFromKeyframes(ks, std::vector<int>{0, 1, 2}, rel, 0.1);
    gicp->setInputSource(scans[3]);
    gicp->align(*aligned);
    const auto Tm = gicp->getFinalTransformation();
    std::printf("submap n %zu p0 %.9g %.9g %.9g %.9g converged %d T03 %.9g T13 %.9g T23 %.9g fitness %.17g\n", submap->size(), submap->points[0].x, submap->points[0].y,
                submap->points[0].z, submap->points[0].intensity, gicp->hasConverged() ? 1 : 0, Tm(0, 3), Tm(1, 3), Tm(2, 3), gicp->lastFitnessScore());
    apd_cloudset_destroy(gicp->nativeHandle(), ks);
  }
  // protected virtuals at a DOUBLE pose; the status message's inlier count
  {
    Probe p;
    p.setMaxCorrespondenceDistance(2.0);
    p.setTransformationEpsilon(0.1);
    p.setAzimuthVar(1.0);
    p.setInputTarget(scans[0]);
    p.setInputSource(scans[1]);
    Eigen::Isometry3d x;  // a pose that is NOT representable in float: the double path must see all of it
    x.matrix()(0, 3) = 0.1 + 1e-9;
    x.matrix()(1, 3) = -0.05 - 3e-10;
    Eigen::Matrix<double, 6, 6> H;
    Eigen::Matrix<double, 6, 1> b;
    const double e = p.linearize(x, &H, &b);
    Eigen::Isometry3d y = x;
    y.matrix()(0, 3) += 0.02;
    const double e_same = p.compute_error(x), e_moved = p.compute_error(y);
    p.update_correspondences(y);
    const double e_after = p.compute_error(y);
    Eigen::Isometry3d small, big;
    small.matrix()(0, 3) = 0.05;
    big.matrix()(0, 3) = 0.2;
    std::printf("probe lin %.17g H00 %.17g b3 %.17g err_same %.17g err_moved %.17g err_after %.17g conv_small %d conv_big %d\n", e, H(0, 0), b(3), e_same, e_moved, e_after,
                p.is_converged(small) ? 1 : 0, p.is_converged(big) ? 1 : 0);
    p.align(*aligned);
    std::printf("inliers %lld of %zu at 0.5 ; %lld at 2.0\n", p.lastInlierCount(0.5), aligned->size(), p.lastInlierCount(2.0));
  }
  // stale cache keys: clouds die and new ones of the same size appear (possibly at the same address); every align must see the
  // cloud it was given. Also setSourceCovariances + swapSourceAndTarget BEFORE any align (the injected set must travel).
  {
    auto gicp = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
    gicp->setMaxCorrespondenceDistance(2.0);
    gicp->setTransformationEpsilon(0.1);
    gicp->setAzimuthVar(1.0);
    gicp->setInputTarget(scans[0]);
    for (int rep = 0; rep < 6; rep++) {
      auto tmp = std::make_shared<Cloud>(*scans[1 + rep % 2]);  // a fresh copy: freed at the end of the iteration
      if (rep % 2) for (auto& q : tmp->points) q.x += 0.125f;
      gicp->setInputSource(tmp);
      gicp->align(*aligned);
      const auto T = gicp->getFinalTransformation();
      std::printf("stale rep %d T03 %.9g T13 %.9g\n", rep, T(0, 3), T(1, 3));
    }
    auto a = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
    a->setMaxCorrespondenceDistance(2.0);
    a->setTransformationEpsilon(1e-6);
    a->setRotationEpsilon(1e-6);
    a->setAzimuthVar(1.0);
    a->setInputSource(scans[0]);
    a->setInputTarget(scans[1]);
    fast_gicp::FastAPDGICP<PointT, PointT>::CovarianceList wide(scans[0]->size());
    for (auto& m : wide) { m.setZero(); m(0, 0) = 4.0; m(1, 1) = 4.0; m(2, 2) = 4.0; }
    a->setSourceCovariances(wide);   // belongs to scans[0]
    a->swapSourceAndTarget();        // scans[0] is now the TARGET and must keep the injected set
    a->align(*aligned);
    const auto Ta = a->getFinalTransformation();
    std::printf("swapinj converged %d T03 %.9g T13 %.9g T23 %.9g\n", a->hasConverged() ? 1 : 0, Ta(0, 3), Ta(1, 3), Ta(2, 3));
  }
  // per-call latency through the C++ class: setInputTarget (cached: the previous source) + setInputSource + align + the fitness
  // score computed on the GPU; PCL's own kd-tree build and getFitnessScore are excluded (force_no_recompute, lastFitnessScore)
  if (argc >= 3 && std::string(argv[2]) == "--latency") {
    auto gicp = std::make_shared<fast_gicp::FastAPDGICP<PointT, PointT>>();
    gicp->setMaxCorrespondenceDistance(2.0);
    gicp->setTransformationEpsilon(0.1);
    gicp->setAzimuthVar(1.0);
    gicp->setSearchMethodTarget(gicp->getSearchMethodTarget(), true);
    std::vector<double> ms;
    gicp->setInputTarget(scans[0]);
    for (int rep = 0; rep < 60; rep++) {
      const int t = 1 + rep % (n_scans - 1), prev = rep == 0 ? 0 : 1 + (rep - 1) % (n_scans - 1);
      const auto t0 = std::chrono::steady_clock::now();
      gicp->setInputTarget(scans[prev]);
      gicp->setInputSource(scans[t]);
      gicp->align(*aligned);
      volatile double f = gicp->lastFitnessScore();
      (void)f;
      ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    std::sort(ms.begin() + 5, ms.end());
    std::printf("latency_cpp_p50_ms %.4f n_points %zu\n", ms[5 + (ms.size() - 5) / 2], scans[1]->size());
  }
  // loop-closure style call without a target: align must print and leave hasConverged() false
  pcl::Registration<PointT, PointT>::Ptr fresh = select_registration_method();
  fresh->setInputSource(scans[0]);
  fresh->align(*aligned);
  std::printf("notarget converged %d\n", fresh->hasConverged() ? 1 : 0);
  return 0;
}
