// Runs the drop-in fast_gicp::FastAPDGICP through a pcl::Registration base pointer, configured the
// way the reference factory does it (radar_graph_slam/src/radar_graph_slam/registrations.cpp:38-50
// with the launch-file values) and driven the way scan_matching_odometry_nodelet.cpp:437-482 and
// loop_detector.cpp:222-236 drive it. Input: a binary file of float32 PointXYZI-layout scans written by
// tests/test_cpp_dropin.py. Output: one text line per registration.
#include <cstdio>
#include <cstdlib>
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
  // loop-closure style call without a target: align must print and leave hasConverged() false
  pcl::Registration<PointT, PointT>::Ptr fresh = select_registration_method();
  fresh->setInputSource(scans[0]);
  fresh->align(*aligned);
  std::printf("notarget converged %d\n", fresh->hasConverged() ? 1 : 0);
  return 0;
}
