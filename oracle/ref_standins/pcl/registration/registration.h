// TEST INFRASTRUCTURE - stand-in for <pcl/registration/registration.h> when the REFERENCE's FastAPDGICP sources are
// compiled (oracle/ref_apdgicp.cpp): brings in what the genuine header brings in transitively and the reference relies
// on (OpenMP, iostream, Eigen's SVD and Cholesky modules, pcl::transformPointCloud), then continues with the
// pcl::Registration stand-in of include/pcl_shim (the next directory on the include path).
#pragma once
#include <omp.h>
#include <cstdlib>
#include <iostream>
#include <Eigen/Dense>
#include <pcl/point_types.h>
#include <pcl/point_cloud.h>
#include <pcl/common/transforms.h>
#include <pcl/search/kdtree.h>
#include_next <pcl/registration/registration.h>
