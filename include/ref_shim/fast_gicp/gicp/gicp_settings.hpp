// IN-CONTAINER STAND-IN for the REFERENCE's own header fast_apdgicp/include/fast_gicp/gicp/gicp_settings.hpp:6-10, used only
// where the reference tree is absent (the GPU test box). NOT for integration and NOT on the drop-in's include path: in a
// RIV-SLAM workspace <fast_gicp/gicp/gicp_settings.hpp> resolves to the reference's file, which the drop-in header
// includes unchanged (shipping a second header of that path would shadow it for the reference's FastGICP / FastVGICP).
// Same include guard and the same three enums in the same order, so seeing both files in one TU is harmless.
#ifndef FAST_GICP_GICP_SETTINGS_HPP
#define FAST_GICP_GICP_SETTINGS_HPP

namespace fast_gicp {

enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };

enum class NeighborSearchMethod { DIRECT27, DIRECT7, DIRECT1, DIRECT_RADIUS };

enum class VoxelAccumulationMode { ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE };

}  // namespace fast_gicp

#endif
