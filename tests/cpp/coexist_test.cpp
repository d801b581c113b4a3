// Compile-only: the drop-in header next to the REFERENCE's own headers in one translation unit, with this repository's
// include/ directory FIRST on the include path — the situation of radar_graph_slam/src/radar_graph_slam/registrations.cpp:13-15,
// which includes the reference's fast_gicp.hpp and fast_vgicp.hpp (CPU classes on the reference's LsqRegistration and
// gicp_settings.hpp) and then fast_apdgicp.hpp. Round 1 shipped its own lsq_registration.hpp / gicp_settings.hpp on the same
// paths and broke exactly this. Built by tests/test_cpp_dropin.py with
//   g++ -fsyntax-only -I include -I include/pcl_shim -I /root/reference/fast_apdgicp/include
// (fast_vgicp.hpp itself needs Boost and the full Eigen, which the image lacks: its use of the shared headers is restated below).
#include <fast_gicp/gicp/gicp_settings.hpp>     // must be the reference's: all three enums
#include <fast_gicp/gicp/lsq_registration.hpp>  // must be the reference's: pure virtual linearize / compute_error, LSQ_OPTIMIZER_TYPE
#include <fast_gicp/gicp/fast_gicp.hpp>         // the reference's FastGICP declaration: `linearize(...) override` on that base
#include <fast_gicp/gicp/fast_apdgicp.hpp>      // this repository's drop-in

#include <type_traits>

#ifndef APDGICP_B200_DROPIN
#error "fast_apdgicp.hpp did not resolve to the drop-in"
#endif

using PointT = pcl::PointXYZI;

// what fast_vgicp.hpp:60-75 and fast_vgicp_voxel.hpp:10-20 need from gicp_settings.hpp
static fast_gicp::NeighborSearchMethod search_method = fast_gicp::NeighborSearchMethod::DIRECT7;
static fast_gicp::VoxelAccumulationMode voxel_mode = fast_gicp::VoxelAccumulationMode::ADDITIVE;
static fast_gicp::LSQ_OPTIMIZER_TYPE optimizer = fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt;

// a CPU registration on the REFERENCE's base class, the way FastGICP / FastVGICP derive from it (fast_gicp.hpp:77-79)
class CpuRegistration : public fast_gicp::LsqRegistration<PointT, PointT> {
protected:
  double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) override { return 0.0; }
  double compute_error(const Eigen::Isometry3d& trans) override { return 0.0; }
};

// the reference's base is still the reference's (abstract, with its LM state), and the drop-in does not derive from it
static_assert(std::is_abstract<fast_gicp::LsqRegistration<PointT, PointT>>::value, "reference LsqRegistration must stay the reference's");
static_assert(!std::is_base_of<fast_gicp::LsqRegistration<PointT, PointT>, fast_gicp::FastAPDGICP<PointT, PointT>>::value, "drop-in must not hijack the CPU base");
static_assert(std::is_base_of<pcl::Registration<PointT, PointT, float>, fast_gicp::FastAPDGICP<PointT, PointT>>::value, "callers hold a pcl::Registration::Ptr");
static_assert(std::is_base_of<fast_gicp::LsqRegistration<PointT, PointT>, fast_gicp::FastGICP<PointT, PointT>>::value, "reference FastGICP untouched");
static_assert(static_cast<int>(fast_gicp::RegularizationMethod::PLANE) == APD_REG_PLANE && static_cast<int>(fast_gicp::RegularizationMethod::FROBENIUS) == APD_REG_FROBENIUS,
              "C ABI enum values follow the reference's declaration order");
static_assert(static_cast<int>(fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt) == APD_OPT_LEVENBERG_MARQUARDT, "optimizer enum");

// the factory branch compiles against the drop-in type and returns the base pointer (registrations.cpp:38-50)
pcl::Registration<PointT, PointT>::Ptr make_apd() {
  fast_gicp::FastAPDGICP<PointT, PointT>::Ptr apdgicp(new fast_gicp::FastAPDGICP<PointT, PointT>());
  apdgicp->setNumThreads(0);
  apdgicp->setTransformationEpsilon(0.01);
  apdgicp->setMaximumIterations(64);
  apdgicp->setMaxCorrespondenceDistance(2.5);
  apdgicp->setCorrespondenceRandomness(20);
  apdgicp->setDistVar(0.86);
  apdgicp->setAzimuthVar(0.5);
  apdgicp->setElevationVar(1.0);
  (void)search_method; (void)voxel_mode; (void)optimizer;
  return apdgicp;
}
