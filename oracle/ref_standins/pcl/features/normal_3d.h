// TEST INFRASTRUCTURE - fast_apdgicp_impl.hpp:5 includes <pcl/features/normal_3d.h> and uses nothing from it.
#pragma once
