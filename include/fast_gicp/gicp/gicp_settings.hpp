// Enumerations of the drop-in FastAPDGICP surface. Values match apd_regularization / apd_optimizer in
// apdgicp_b200.h, which in turn match the reference's declaration order
// (reference fast_apdgicp/include/fast_gicp/gicp/gicp_settings.hpp:6, lsq_registration.hpp:13).
#pragma once

namespace fast_gicp {

enum class RegularizationMethod { NONE = 0, MIN_EIG = 1, NORMALIZED_MIN_EIG = 2, PLANE = 3, FROBENIUS = 4 };

enum class LSQ_OPTIMIZER_TYPE { GaussNewton = 0, LevenbergMarquardt = 1 };

}  // namespace fast_gicp
