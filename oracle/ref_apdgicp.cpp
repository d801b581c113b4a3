// TEST INFRASTRUCTURE - the REFERENCE's own FastAPDGICP, compiled UNMODIFIED from where its sources lie under
// /root/reference (include path only; nothing is copied into this repository):
//   fast_apdgicp/include/fast_gicp/gicp/fast_apdgicp.hpp, impl/fast_apdgicp_impl.hpp        (APD_H / APD_I)
//   fast_apdgicp/include/fast_gicp/gicp/lsq_registration.hpp, impl/lsq_registration_impl.hpp (LSQ_H / LSQ_I)
//   fast_apdgicp/include/fast_gicp/so3/so3.hpp, gicp/gicp_settings.hpp
//   radar_graph_slam/include/scan_context/nanoflann.hpp (the kd-tree behind the pcl::search::KdTree stand-in)
// Eigen, PCL and Boost are not installed in this image; the headers those sources include resolve to the stand-ins under
// oracle/ref_standins/ (Eigen: fixed-size dense algebra restated from Eigen 3.3's published algorithms; pcl / boost: the
// handful of types the sources touch) and include/pcl_shim (pcl::Registration, pcl::PointCloud). So what this library
// pins is the reference's own TEXT - every formula, loop, branch and default of APD_I:14-363, LSQ_I:11-173 and so3.hpp,
// executed as written - while the third-party arithmetic underneath (matrix products, 4x4 inverse, JacobiSVD, LDLT,
// kd-tree) is a restatement. The C entry points below only move data in and out and expose protected members through a
// derived class. Built by `make -C oracle ref` into oracle/_ref/libref_apdgicp.so (git-ignored); used by
// tests/test_reference_apdgicp.py and tests/golden/make_ref_golden.py only. Never used by the product.
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include <boost/format.hpp>
#include <pcl/registration/registration.h>   // the stand-ins (must precede the reference's headers: omp.h, iostream, SVD, LDLT)

#include <fast_gicp/gicp/fast_apdgicp.hpp>
#include <fast_gicp/gicp/impl/lsq_registration_impl.hpp>
#include <fast_gicp/gicp/impl/fast_apdgicp_impl.hpp>

// The one libm call on the path whose result depends on the C library: APD_I:168,172,173 call atan2 on float arguments
// (`using namespace std;`, APD_I:7), i.e. atan2f. glibc < 2.41 (Ubuntu 18.04 / 20.04 ship 2.27 / 2.31, this image 2.39) returns a
// result within 1 ulp that is not always the correctly rounded one; SURVEY 8c fixes the convention "correctly rounded" for
// the oracle and the device. This definition interposes the symbol inside this library only: mode 0 forwards to the C library's
// atan2f (the reference exactly as it would run on this machine), mode 1 rounds the double result once (the convention).
static int g_atan2f_mode = 0;
extern "C" float atan2f(float y, float x) noexcept {
  typedef float (*fn_t)(float, float);
  static fn_t libm = (fn_t)dlsym(RTLD_NEXT, "atan2f");
  return g_atan2f_mode == 0 && libm ? libm(y, x) : (float)std::atan2((double)y, (double)x);
}
extern "C" void ref_apd_set_atan2f_mode(int correctly_rounded) { g_atan2f_mode = correctly_rounded; }

namespace {

typedef pcl::PointXYZI P;
typedef std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> CovVec;

class RefAPD : public fast_gicp::FastAPDGICP<P, P> {
  typedef fast_gicp::FastAPDGICP<P, P> Base;

public:
  using Base::lm_max_iterations_;
  using Base::lsq_optimizer_type_;
  using Base::correspondences_;
  using Base::sq_distances_;
  using Base::mahalanobis_;
  using Base::source_covs_;
  using Base::target_covs_;
  using Base::nr_iterations_;
  using Base::input_;
  using Base::target_;
  using Base::source_kdtree_;
  using Base::target_kdtree_;
  double call_linearize(const Eigen::Isometry3d& x, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) { return this->linearize(x, H, b); }
  double call_compute_error(const Eigen::Isometry3d& x) { return this->compute_error(x); }
  // computeTransformation's first lines (APD_I:122-127)
  void ensure_covariances() {
    if (source_covs_.size() != input_->size()) this->template calculate_covariances<P>(input_, *source_kdtree_, source_covs_);
    if (target_covs_.size() != target_->size()) this->template calculate_covariances<P>(target_, *target_kdtree_, target_covs_);
  }
  std::string debug_text, err_text;
};

struct ref_params {   // == oracle_params (oracle/oracle_capi.cpp)
  int num_threads, k_correspondences, regularization, max_iterations, optimizer, lm_max_iterations;
  double max_corr_dist, rotation_epsilon, transformation_epsilon, lm_init_lambda_factor, dist_var, azimuth_var, elevation_var;
};

pcl::PointCloud<P>::Ptr make_cloud(const float* xyz, int stride, int n) {
  pcl::PointCloud<P>::Ptr c(new pcl::PointCloud<P>);
  c->resize(n);
  for (int i = 0; i < n; i++) {
    P& p = c->points[i];
    p.x = xyz[(size_t)i * stride];
    p.y = xyz[(size_t)i * stride + 1];
    p.z = xyz[(size_t)i * stride + 2];
  }
  return c;
}

Eigen::Isometry3d iso_from_d16(const double* p) {   // row-major 4x4
  Eigen::Matrix4d m;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m(i, j) = p[i * 4 + j];
  return Eigen::Isometry3d(m);
}

// numeric values of RegularizationMethod in the oracle / C ABI: NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS (gicp_settings.hpp:6)
fast_gicp::RegularizationMethod reg_of(int r) {
  switch (r) {
    case 0: return fast_gicp::RegularizationMethod::NONE;
    case 1: return fast_gicp::RegularizationMethod::MIN_EIG;
    case 2: return fast_gicp::RegularizationMethod::NORMALIZED_MIN_EIG;
    case 3: return fast_gicp::RegularizationMethod::PLANE;
    default: return fast_gicp::RegularizationMethod::FROBENIUS;
  }
}

}  // namespace

extern "C" {

void* ref_apd_create() { return new RefAPD(); }
void ref_apd_destroy(void* h) { delete static_cast<RefAPD*>(h); }

// the constructor's defaults as the reference sets them (APD_I:14-28, LSQ_I:11-24, APD_H:107-109)
void ref_apd_get_defaults(ref_params* p) {
  RefAPD r;
  struct Peek : RefAPD {
    using RefAPD::k_correspondences_; using RefAPD::regularization_method_; using RefAPD::max_iterations_; using RefAPD::corr_dist_threshold_;
    using RefAPD::rotation_epsilon_; using RefAPD::transformation_epsilon_; using RefAPD::lm_init_lambda_factor_;
    using RefAPD::distance_variance_; using RefAPD::azimuth_variance_; using RefAPD::elevation_variance_; using RefAPD::num_threads_;
  };
  Peek& q = static_cast<Peek&>(r);
  p->num_threads = q.num_threads_;
  p->k_correspondences = q.k_correspondences_;
  p->regularization = (int)q.regularization_method_;
  p->max_iterations = q.max_iterations_;
  p->optimizer = q.lsq_optimizer_type_ == fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt ? 1 : 0;
  p->lm_max_iterations = q.lm_max_iterations_;
  p->max_corr_dist = q.corr_dist_threshold_;
  p->rotation_epsilon = q.rotation_epsilon_;
  p->transformation_epsilon = q.transformation_epsilon_;
  p->lm_init_lambda_factor = q.lm_init_lambda_factor_;
  p->dist_var = q.distance_variance_;
  p->azimuth_var = q.azimuth_variance_;
  p->elevation_var = q.elevation_variance_;
}

void ref_apd_set_params(void* h, const ref_params* p) {
  RefAPD* r = static_cast<RefAPD*>(h);
  r->setNumThreads(p->num_threads);
  r->setCorrespondenceRandomness(p->k_correspondences);
  r->setRegularizationMethod(reg_of(p->regularization));
  r->setMaximumIterations(p->max_iterations);
  r->lsq_optimizer_type_ = p->optimizer ? fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt : fast_gicp::LSQ_OPTIMIZER_TYPE::GaussNewton;
  r->lm_max_iterations_ = p->lm_max_iterations;
  r->setMaxCorrespondenceDistance(p->max_corr_dist);
  r->setRotationEpsilon(p->rotation_epsilon);
  r->setTransformationEpsilon(p->transformation_epsilon);
  r->setInitialLambdaFactor(p->lm_init_lambda_factor);
  r->setDistVar(p->dist_var);
  r->setAzimuthVar(p->azimuth_var);
  r->setElevationVar(p->elevation_var);
}

void ref_apd_set_source(void* h, const float* xyz, int stride_floats, int n) { static_cast<RefAPD*>(h)->setInputSource(make_cloud(xyz, stride_floats, n)); }
void ref_apd_set_target(void* h, const float* xyz, int stride_floats, int n) { static_cast<RefAPD*>(h)->setInputTarget(make_cloud(xyz, stride_floats, n)); }
void ref_apd_swap(void* h) { static_cast<RefAPD*>(h)->swapSourceAndTarget(); }
void ref_apd_clear_source(void* h) { static_cast<RefAPD*>(h)->clearSource(); }
void ref_apd_clear_target(void* h) { static_cast<RefAPD*>(h)->clearTarget(); }

// pcl::Registration::align(output, guess) -> computeTransformation (APD_I:121-131 -> LSQ_I:55-81). With debug != 0 the
// reference's own LM table (LSQ_I:148-155) is captured from std::cout, floating-point columns at 17 significant digits.
int ref_apd_align(void* h, const float* guess16, int debug, float* T16, int* converged, int* iterations, float* out_xyz) {
  RefAPD* r = static_cast<RefAPD*>(h);
  if (!r->input_ || !r->target_) return -1;
  Eigen::Matrix4f g = Eigen::Matrix4f::Identity();
  if (guess16) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) g(i, j) = guess16[i * 4 + j];
  pcl::PointCloud<P> out;
  std::ostringstream cap, ecap;
  std::streambuf* saved = nullptr;
  std::streambuf* esaved = std::cerr.rdbuf(ecap.rdbuf());   // "lm not converged!!" (LSQ_I:72) goes to std::cerr
  r->setDebugPrint(debug != 0);
  if (debug) {
    saved = std::cout.rdbuf(cap.rdbuf());
    boost::format_stand_in_precision() = 17;
  }
  r->align(out, g);
  std::cerr.rdbuf(esaved);
  r->err_text = ecap.str();
  if (debug) {
    std::cout.rdbuf(saved);
    boost::format_stand_in_precision() = 0;
    r->debug_text = cap.str();
  }
  const Eigen::Matrix4f T = r->getFinalTransformation();
  if (T16) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T16[i * 4 + j] = T(i, j);
  if (converged) *converged = r->hasConverged() ? 1 : 0;
  if (iterations) *iterations = r->nr_iterations_;
  if (out_xyz) for (size_t i = 0; i < out.points.size(); i++) { out_xyz[3 * i] = out.points[i].x; out_xyz[3 * i + 1] = out.points[i].y; out_xyz[3 * i + 2] = out.points[i].z; }
  return 0;
}

// pcl::Registration::getFitnessScore (the base class's own 1-NN pass over tree_, which align()'s initCompute built)
double ref_apd_fitness(void* h, double max_range) { return static_cast<RefAPD*>(h)->getFitnessScore(max_range); }

int ref_apd_get_debug_text(void* h, char* buf, int cap) {
  const std::string& s = static_cast<RefAPD*>(h)->debug_text;
  if (buf && cap > 0) { const int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1; std::memcpy(buf, s.data(), n); buf[n] = 0; }
  return (int)s.size();
}

// 1 when the last align printed "lm not converged!!" (step_optimize returned false, LSQ_I:71-74)
int ref_apd_lm_failed(void* h) { return static_cast<RefAPD*>(h)->err_text.find("lm not converged!!") != std::string::npos ? 1 : 0; }

int ref_apd_compute_covariances(void* h) {
  RefAPD* r = static_cast<RefAPD*>(h);
  if (!r->input_ || !r->target_) return -1;
  r->ensure_covariances();
  return 0;
}

int ref_apd_get_covariances(void* h, int which, double* out9) {   // the 3x3 block, row-major; the rest of the 4x4 is returned by ref_apd_get_covariances4
  const CovVec& v = which ? static_cast<RefAPD*>(h)->target_covs_ : static_cast<RefAPD*>(h)->source_covs_;
  if (out9) for (size_t n = 0; n < v.size(); n++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out9[n * 9 + i * 3 + j] = v[n](i, j);
  return (int)v.size();
}
int ref_apd_get_covariances4(void* h, int which, double* out16) {
  const CovVec& v = which ? static_cast<RefAPD*>(h)->target_covs_ : static_cast<RefAPD*>(h)->source_covs_;
  if (out16) for (size_t n = 0; n < v.size(); n++) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out16[n * 16 + i * 4 + j] = v[n](i, j);
  return (int)v.size();
}
void ref_apd_set_covariances(void* h, int which, const double* in9, int n) {   // setSourceCovariances / setTargetCovariances (APD_I:110-118)
  CovVec v((size_t)n, Eigen::Matrix4d::Zero());
  for (int k = 0; k < n; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[k](i, j) = in9[(size_t)k * 9 + i * 3 + j];
  if (which) static_cast<RefAPD*>(h)->setTargetCovariances(v); else static_cast<RefAPD*>(h)->setSourceCovariances(v);
}

static double finish_lin(const Eigen::Matrix<double, 6, 6>& H, const Eigen::Matrix<double, 6, 1>& b, double e, double* H36, double* b6) {
  if (H36) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H36[i * 6 + j] = H(i, j);
  if (b6) for (int i = 0; i < 6; i++) b6[i] = b(i);
  return e;
}
// evaluateCost (LSQ_I:50-52): float pose -> linearize
double ref_apd_linearize(void* h, const float* pose16, double* H36, double* b6) {
  RefAPD* r = static_cast<RefAPD*>(h);
  r->ensure_covariances();
  Eigen::Matrix4f m;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m(i, j) = pose16[i * 4 + j];
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  const double e = r->evaluateCost(m, &H, &b);
  return finish_lin(H, b, e, H36, b6);
}
// the protected hooks at a double pose: linearize (APD_I:198-272), compute_error (APD_I:275-298)
double ref_apd_linearize_d(void* h, const double* pose16, double* H36, double* b6) {
  RefAPD* r = static_cast<RefAPD*>(h);
  r->ensure_covariances();
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  const double e = r->call_linearize(iso_from_d16(pose16), &H, &b);
  return finish_lin(H, b, e, H36, b6);
}
double ref_apd_compute_error_d(void* h, const double* pose16) { return static_cast<RefAPD*>(h)->call_compute_error(iso_from_d16(pose16)); }

int ref_apd_get_correspondences(void* h, int* corr, float* sq_dist) {
  RefAPD* r = static_cast<RefAPD*>(h);
  if (corr) std::memcpy(corr, r->correspondences_.data(), r->correspondences_.size() * sizeof(int));
  if (sq_dist) std::memcpy(sq_dist, r->sq_distances_.data(), r->sq_distances_.size() * sizeof(float));
  return (int)r->correspondences_.size();
}
// the 3x3 block of mahalanobis_[i]; rows without a correspondence are whatever resize() left there (zero)
int ref_apd_get_mahalanobis(void* h, double* out9) {
  RefAPD* r = static_cast<RefAPD*>(h);
  if (out9) for (size_t n = 0; n < r->mahalanobis_.size(); n++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out9[n * 9 + i * 3 + j] = r->mahalanobis_[n](i, j);
  return (int)r->mahalanobis_.size();
}
void ref_apd_get_final_hessian(void* h, double* H36) {
  const Eigen::Matrix<double, 6, 6>& H = static_cast<RefAPD*>(h)->getFinalHessian();
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H36[i * 6 + j] = H(i, j);
}

// so3.hpp:21-31, 59-78
void ref_skewd(const double* x3, double* out9) {
  const Eigen::Matrix3d m = fast_gicp::skewd(Eigen::Vector3d(x3[0], x3[1], x3[2]));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out9[i * 3 + j] = m(i, j);
}
void ref_so3_exp(const double* omega3, double* quat_wxyz, double* R9) {
  const Eigen::Quaterniond q = fast_gicp::so3_exp(Eigen::Vector3d(omega3[0], omega3[1], omega3[2]));
  quat_wxyz[0] = q.w(); quat_wxyz[1] = q.x(); quat_wxyz[2] = q.y(); quat_wxyz[3] = q.z();
  const Eigen::Matrix3d m = q.toRotationMatrix();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R9[i * 3 + j] = m(i, j);
}

// ---- the stand-in linear algebra itself (oracle/ref_standins/Eigen), exposed so that tests can hold it against numpy / LAPACK ----
void ref_eigen_svd3(const double* in9, double* U9, double* S3, double* V9) {
  Eigen::Matrix3d m;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = in9[i * 3 + j];
  Eigen::JacobiSVD<Eigen::Matrix3d> svd(m, Eigen::ComputeFullU | Eigen::ComputeFullV);
  for (int i = 0; i < 3; i++) {
    S3[i] = svd.singularValues()(i);
    for (int j = 0; j < 3; j++) { U9[i * 3 + j] = svd.matrixU()(i, j); V9[i * 3 + j] = svd.matrixV()(i, j); }
  }
}
void ref_eigen_ldlt6_solve(const double* A36, const double* b6, double* x6) {
  Eigen::Matrix<double, 6, 6> A;
  Eigen::Matrix<double, 6, 1> b;
  for (int i = 0; i < 6; i++) { b(i) = b6[i]; for (int j = 0; j < 6; j++) A(i, j) = A36[i * 6 + j]; }
  Eigen::LDLT<Eigen::Matrix<double, 6, 6>> solver(A);
  const Eigen::Matrix<double, 6, 1> x = solver.solve(b);
  for (int i = 0; i < 6; i++) x6[i] = x(i);
}
void ref_eigen_inverse(const double* in, int n, double* out) {
  if (n == 3) {
    Eigen::Matrix3d m;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = in[i * 3 + j];
    const Eigen::Matrix3d r = m.inverse();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[i * 3 + j] = r(i, j);
  } else {
    Eigen::Matrix4d m;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m(i, j) = in[i * 4 + j];
    const Eigen::Matrix4d r = m.inverse();
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i * 4 + j] = r(i, j);
  }
}
// AngleAxisd(yaw, Z) * AngleAxisd(pitch, Y) -> Matrix3d, as APD_I:174-177 writes it
void ref_eigen_yaw_pitch(double yaw, double pitch, double* R9) {
  Eigen::AngleAxisd pitchAngle(Eigen::AngleAxisd(pitch, Eigen::Vector3d::UnitY()));
  Eigen::AngleAxisd yawAngle(Eigen::AngleAxisd(yaw, Eigen::Vector3d::UnitZ()));
  Eigen::Matrix3d R;
  R = yawAngle * pitchAngle;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R9[i * 3 + j] = R(i, j);
}

const char* ref_apd_version() { return "fast_gicp::FastAPDGICP<pcl::PointXYZI, pcl::PointXYZI> compiled from /root/reference/fast_apdgicp/include over stand-in Eigen / PCL / Boost headers"; }
}
