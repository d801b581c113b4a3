/*
 * apdgicp_b200 — C ABI of the B200-native FastAPDGICP scan-matching hot path.
 *
 * The reference has no C ABI for this path: its "plugin" is the C++ class
 * fast_gicp::FastAPDGICP<pcl::PointXYZI, pcl::PointXYZI> (reference
 * fast_apdgicp/include/fast_gicp/gicp/fast_apdgicp.hpp:33-110, derived from LsqRegistration,
 * fast_apdgicp/include/fast_gicp/gicp/lsq_registration.hpp:16-84, derived from pcl::Registration).
 * Every entry point below cites the member of that class (or the PCL base-class behaviour the
 * callers observe) that it replaces. The drop-in C++ class in include/apdgicp_b200/fast_apdgicp.hpp
 * forwards to these functions and to nothing else; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; opaque handles; every function returns an apd_status (0 = ok) and never
 *     throws or aborts across the boundary. apd_last_error() gives the text of the last failure.
 *   - 4x4 transforms are ROW-MAJOR float[16] (Eigen::Matrix4f is column-major: transpose at the
 *     binding). 6x6 Hessians are row-major double[36] (symmetric).
 *   - point clouds are read as three consecutive floats x,y,z every `stride_bytes` bytes
 *     (32 for pcl::PointXYZI, 16 for packed float4, 12 for packed xyz). Intensity is never read
 *     (the reference touches only getVector4fMap()/getVector3fMap(), fast_apdgicp_impl.hpp:149,167,229,320).
 *   - there is NO CPU fallback: without a CUDA device apd_create fails with APD_ERR_NO_DEVICE.
 *   - a handle owns one CUDA stream context and is not re-entrant (like the reference object,
 *     SURVEY.md §8b "Threading"); different handles may be used from different host threads.
 */
#ifndef APDGICP_B200_H
#define APDGICP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APDGICP_B200_ABI_VERSION 1

typedef struct apd_context* apd_handle;      /* one registration object (FastAPDGICP instance)   */
typedef struct apd_cloudset_s* apd_cloudset; /* a batch of clouds resident in HBM, with grid+covs */

typedef enum apd_status {
  APD_OK = 0,
  APD_ERR_NO_DEVICE = 1,     /* no CUDA device / driver: the product has no CPU path            */
  APD_ERR_INVALID = 2,       /* bad argument                                                    */
  APD_ERR_NO_INPUT = 3,      /* align without source/target (pcl::Registration::align prints and returns) */
  APD_ERR_TOO_FEW_POINTS = 4,/* cloud smaller than k (reference reads uninitialised memory, fast_apdgicp_impl.hpp:318-321) */
  APD_ERR_CUDA = 5,          /* a CUDA call failed; see apd_last_error                          */
  APD_ERR_UNSUPPORTED = 6
} apd_status;

/* fast_gicp::RegularizationMethod, gicp/gicp_settings.hpp:6 (same numeric values) */
typedef enum apd_regularization {
  APD_REG_NONE = 0, APD_REG_MIN_EIG = 1, APD_REG_NORMALIZED_MIN_EIG = 2, APD_REG_PLANE = 3, APD_REG_FROBENIUS = 4
} apd_regularization;

/* fast_gicp::LSQ_OPTIMIZER_TYPE, gicp/lsq_registration.hpp:13 */
typedef enum apd_optimizer { APD_OPT_GAUSS_NEWTON = 0, APD_OPT_LEVENBERG_MARQUARDT = 1 } apd_optimizer;

typedef enum apd_mem { APD_MEM_HOST = 0, APD_MEM_DEVICE = 1 } apd_mem;

/* All tunables of the reference object. Defaults (apd_default_params) are the constructor values:
 * fast_apdgicp_impl.hpp:14-28, lsq_registration_impl.hpp:11-24, fast_apdgicp.hpp:107-109. */
typedef struct apd_params {
  int32_t k_correspondences;      /* setCorrespondenceRandomness, fast_apdgicp_impl.hpp:45 (20)              */
  int32_t regularization;         /* setRegularizationMethod, :50 (PLANE)                                    */
  int32_t max_iterations;         /* pcl setMaximumIterations; lsq_registration_impl.hpp:13 (64)             */
  int32_t optimizer;              /* lsq_optimizer_type_, lsq_registration_impl.hpp:17 (LM; no setter)       */
  int32_t lm_max_iterations;      /* lsq_registration_impl.hpp:19 (10)                                       */
  int32_t num_threads;            /* setNumThreads, fast_apdgicp_impl.hpp:34 — accepted and ignored on GPU   */
  double max_corr_dist;           /* pcl setMaxCorrespondenceDistance; fast_apdgicp_impl.hpp:23 (FLT_MAX)    */
  double rotation_epsilon;        /* setRotationEpsilon, lsq_registration_impl.hpp:14,30 (2e-3)              */
  double transformation_epsilon;  /* pcl setTransformationEpsilon; lsq_registration_impl.hpp:15 (5e-4)       */
  double lm_init_lambda_factor;   /* setInitialLambdaFactor, lsq_registration_impl.hpp:20,35 (1e-9)          */
  double dist_var;                /* setDistVar, fast_apdgicp.hpp:109 (0.86)                                 */
  double azimuth_var;             /* setAzimuthVar [deg], fast_apdgicp.hpp:107 (0.5)                         */
  double elevation_var;           /* setElevationVar [deg], fast_apdgicp.hpp:108 (1.0)                       */
} apd_params;

/* Result of one registration: 96 bytes, the record gathered across GPUs. */
typedef struct apd_result {
  float T[16];          /* final_transformation_ (row-major), lsq_registration_impl.hpp:78        */
  double fitness;       /* pcl getFitnessScore(): mean squared 1-NN distance, DBL_MAX if none     */
  double error;         /* last y0 = sum e^T M e returned by linearize, fast_apdgicp_impl.hpp:240 */
  int32_t converged;    /* converged_, lsq_registration_impl.hpp:75                               */
  int32_t iterations;   /* nr_iterations_, lsq_registration_impl.hpp:68 (0-based index of the last outer iteration) */
  int32_t status;       /* apd_status of this pair; 0 ok; 100 = "lm not converged!!" (lsq_registration_impl.hpp:71-74) */
  int32_t num_inliers;  /* correspondences inside the gate at the last linearize                  */
} apd_result;

#define APD_STATUS_LM_FAILED 100

/* ---- object lifetime: FastAPDGICP::FastAPDGICP / ~FastAPDGICP, fast_apdgicp.hpp:48-49 ---- */
int apd_create(int device_id, apd_handle* out);
int apd_destroy(apd_handle h);
const char* apd_last_error(apd_handle h);
int apd_abi_version(void);
/* Run all work of this handle on an existing CUDA stream (cudaStream_t); NULL = the handle's own. */
int apd_set_stream(apd_handle h, void* cuda_stream);

/* ---- parameters: the setters of fast_apdgicp.hpp:51-57, lsq_registration.hpp:51-53 and the PCL
 *      base-class setters the factory calls (radar_graph_slam/src/radar_graph_slam/registrations.cpp:38-50) ---- */
int apd_default_params(apd_params* p);
int apd_set_params(apd_handle h, const apd_params* p);
int apd_get_params(apd_handle h, apd_params* p);

/* ---- inputs: setInputSource / setInputTarget, fast_apdgicp_impl.hpp:90-108.
 * cache_key carries the reference's pointer-identity cache (fast_apdgicp_impl.hpp:91,102): a call
 * with the key the slot already holds returns at once; a key held by the OTHER slot moves its grid
 * and covariances across instead of recomputing them (results are identical). 0 = no caching.
 * A key names DATA: the caller must not reuse a key for other points while a slot still holds it (the reference compares
 * shared_ptrs it keeps alive; the drop-in class pushes every new cloud to the device at once, so the device only ever holds
 * keys of clouds the object itself keeps alive).
 * mem says whether xyz is a host or a device pointer. ---- */
int apd_set_source(apd_handle h, const float* xyz, int stride_bytes, int n, uint64_t cache_key, int mem);
int apd_set_target(apd_handle h, const float* xyz, int stride_bytes, int n, uint64_t cache_key, int mem);
int apd_swap_source_and_target(apd_handle h); /* fast_apdgicp_impl.hpp:68-75 */
int apd_clear_source(apd_handle h);           /* fast_apdgicp_impl.hpp:78-81 */
int apd_clear_target(apd_handle h);           /* fast_apdgicp_impl.hpp:84-87 */

/* ---- the hot path ---- */
/* pcl::Registration::align(output, guess) -> FastAPDGICP::computeTransformation
 * (fast_apdgicp_impl.hpp:121-130) -> LsqRegistration::computeTransformation (lsq_registration_impl.hpp:55-80).
 * guess == NULL means identity. Also fills out->fitness (getFitnessScore with max_range = DBL_MAX). */
int apd_align(apd_handle h, const float guess[16], apd_result* out);
/* pcl::Registration::getFitnessScore(max_range) on the last alignment. */
int apd_fitness(apd_handle h, double max_range, double* score);
/* "Next" row SURVEY.md §8(f)-1: InformationMatrixCalculator::calc_fitness_score(cloud1, cloud2, relpose, max_range)
 * (radar_graph_slam/src/radar_graph_slam/information_matrix_calculator.cpp:55-86; callers radar_graph_slam_nodelet.cpp:419,704,
 * loop_detector.cpp:315): mean squared 1-NN distance of T * cloud2 in cloud1, counting pairs with d2 <= max_range; DBL_MAX if none.
 * cloud1 is the handle's target, cloud2 its source; T == NULL is identity; n_used (may be NULL) receives the pair count. */
int apd_fitness_score(apd_handle h, const float T[16], double max_range, double* score, int64_t* n_used);
/* pcl::transformPointCloud(*input_, output, T) (lsq_registration_impl.hpp:79): writes x,y,z every
 * out_stride_bytes; T == NULL uses the last final transformation. */
int apd_transform_source(apd_handle h, const float T[16], float* out_xyz, int out_stride_bytes, int mem);
/* LsqRegistration::evaluateCost(pose, H, b) (lsq_registration_impl.hpp:50-52) = linearize at an
 * arbitrary pose (fast_apdgicp_impl.hpp:198-272). H, b may be NULL. */
int apd_linearize(apd_handle h, const float pose[16], double H[36], double b[6], double* error);
int apd_get_final_hessian(apd_handle h, double H[36]); /* getFinalHessian, lsq_registration_impl.hpp:45-47 */
/* The protected hooks of the reference class at a DOUBLE pose (they take an Eigen::Isometry3d; row-major double[16]):
 * FastAPDGICP::linearize (fast_apdgicp_impl.hpp:198-272: update_correspondences, then H, b, error; H and b may be NULL, which is
 * also all update_correspondences (:133-194) needs) and FastAPDGICP::compute_error (:275-298: sum e^T M e at `pose` with the
 * correspondences and Mahalanobis matrices the LAST linearize left behind). */
int apd_linearize_d(apd_handle h, const double pose[16], double H[36], double b[6], double* error);
int apd_compute_error(apd_handle h, const double pose[16], double* error);
/* ScanMatchingOdometryNodelet::publish_scan_matching_status (radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:698-712): number of
 * points of the aligned cloud (source moved by T; NULL = the last final transformation) whose nearest target point is STRICTLY closer
 * than max_dist (k_sq_dists[0] < max_dist * max_dist; the nodelet uses 0.5 m). After apd_align and with T == NULL no search runs:
 * the align kernel keeps every point's squared 1-NN distance at the final pose. */
int apd_inlier_count(apd_handle h, const float T[16], double max_dist, int64_t* n_inliers);

/* ---- state the reference exposes or that parity tests read. which: 0 = source, 1 = target ---- */
int apd_compute_covariances(apd_handle h);             /* calculate_covariances for stale clouds, fast_apdgicp_impl.hpp:122-127 */
int apd_get_knn(apd_handle h, int which, int32_t* idx_out /* n*k, (d2,index)-ascending */);
int apd_get_covariances(apd_handle h, int which, double* c16_out /* n*16, Matrix4d layout */); /* getSource/TargetCovariances, fast_apdgicp.hpp:68-74 */
int apd_set_covariances(apd_handle h, int which, const double* c16, int n);                    /* setSource/TargetCovariances, fast_apdgicp_impl.hpp:111-118 */
int apd_get_correspondences(apd_handle h, int32_t* corr_out /* n_src, -1 = none */, float* sq_dist_out /* may be NULL */);
int apd_get_mahalanobis(apd_handle h, double* m16_out /* n_src*16, zero for unmatched */);
/* LM trial table (the rows setDebugPrint prints, lsq_registration_impl.hpp:148-154):
 * 8 doubles per row = outer, inner, y0, yi, rho, lambda, |d|, accepted. */
int apd_get_lm_trace(apd_handle h, double* rows8, int max_rows, int* n_rows);

/* ---- batched path: many independent pairs in one launch sequence (configs C2 and C4). ----
 * A cloud set is a ragged batch of clouds kept in HBM; building it uploads (or copies) the points;
 * grids and covariances are computed on first use with the handle's parameters and then cached. */
int apd_cloudset_create(apd_handle h, const float* xyz, int stride_bytes, const int32_t* offsets /* host, n_clouds+1 */,
                        int n_clouds, int mem, apd_cloudset* out);
int apd_cloudset_destroy(apd_handle h, apd_cloudset cs);
int apd_cloudset_prepare(apd_handle h, apd_cloudset cs); /* grid + kNN + covariances now (asynchronous on the stream) */
/* Align n_pairs pairs (src_idx[i] of `src`) -> (tgt_idx[i] of `tgt`); src and tgt may be the same set
 * (scan-to-scan odometry: every scan is the source of one pair and the target of the next, covariances
 * computed once). NULL index arrays mean i -> i. guesses: n_pairs*16 floats (host) or NULL = identity.
 * out: n_pairs records in host (out_mem = APD_MEM_HOST, the call synchronises) or device memory
 * (APD_MEM_DEVICE, asynchronous on the handle's stream). */
int apd_align_pairs(apd_handle h, apd_cloudset src, apd_cloudset tgt, const int32_t* src_idx, const int32_t* tgt_idx,
                    const float* guesses, int n_pairs, apd_result* out, int out_mem);
/* Convenience: ragged host arrays in, host results out (builds two cloud sets, aligns i -> i). */
int apd_batch_align(apd_handle h, const float* pts_src, const int32_t* off_src, const float* pts_tgt, const int32_t* off_tgt,
                    int stride_bytes, const float* guesses, int n_pairs, apd_result* out);

/* Block until everything this handle has enqueued (device-output calls) is complete. */
int apd_synchronize(apd_handle h);

/* ---- tuning knobs with no counterpart in the reference (results never depend on them) ----
 *   "cells_per_point"  voxel-grid cell budget per point (default 4)
 *   "team_size"        CTAs cooperating on one pair: 0 = automatic, 1 = one CTA, 2..16 = cluster
 *   "force_unstaged"   1 = never stage the target grid in shared memory
 *   "max_teams"        cap on concurrently processed pairs (0 = as many as fit)
 *   "knn_packed"       1 = kNN collects candidates in the packed 32-bit list first (default), 0 = exact list only
 *   "fitness_max_range" max_range of the getFitnessScore the batched calls fill into apd_result.fitness (default DBL_MAX)
 *   "fused_build"      1 = small clouds are gridded by one launch for all levels (default), 0 = the multi-kernel pipeline
 *   "knn_fine_rings"   kNN: rings searched on one level of the grid pyramid before restarting on the next coarser one
 *   "knn_leaf_parts"   leaf-mode kNN: warps that share the 32 queries of one leaf (0 = automatic: 1 for batches, up to 4 for one scan)
 *   "timeline"         1 = the align kernel stamps its phases for apd_get_timeline (profiling aid)
 *   "kernel_timing"    1 = CUDA events around the hot launches (apd_get_kernel_times) */
int apd_set_option(apd_handle h, const char* name, double value);

/* Batched calc_fitness_score over cloud sets (one launch for a sliding window of keyframe pairs): scores[i] for cloud src_idx[i]
 * of `src` (cloud2) against cloud tgt_idx[i] of `tgt` (cloud1) at poses[i] (n_pairs*16 floats, NULL = identity). */
int apd_fitness_pairs(apd_handle h, apd_cloudset src, apd_cloudset tgt, const int32_t* src_idx, const int32_t* tgt_idx, const float* poses,
                      int n_pairs, double max_range, double* scores);

/* "Next" row SURVEY.md §8(f)-3: LoopDetector::matching over ALL loop candidates (radar_graph_slam/src/radar_graph_slam/loop_detector.cpp:379-441,
 * disabled in the reference because it is too slow on the CPU) as ONE launch: target = cloud keyframe_idx of `keyframes`
 * (registration->setInputTarget(new_keyframe->cloud)), sources = clouds cand_idx[0..n) of `candidates` (NULL = 0..n-1), guesses = n*16 floats
 * or NULL. Selection as at :415-423 (skip a candidate that did not converge or scores above the best so far: the earliest of equal scores
 * wins) and :431-434 (no loop if the best score exceeds fitness_score_thresh): *best = position in cand_idx or -1, relative_pose = its
 * final transformation, *best_score = its getFitnessScore(fitness_score_max_range) (DBL_MAX if none converged); records (may be NULL)
 * receives all n results. The two cloud-set arguments may be the same set. */
int apd_match_candidates(apd_handle h, apd_cloudset candidates, const int32_t* cand_idx, int n_candidates, apd_cloudset keyframes, int keyframe_idx,
                         const float* guesses, double fitness_score_max_range, double fitness_score_thresh, int32_t* best, float relative_pose[16],
                         double* best_score, apd_result* records);

/* ---- "next" rows SURVEY.md §8(f)-4 and §8(f)-2: the filters in front of the scan matcher and the submap behind it ----
 * Parameters of PreprocessingNodelet (radar_graph_slam/apps/preprocessing_nodelet.cpp:137-205; apd_default_preprocess_params gives
 * the code defaults, the launch file overrides several: radar_graph_slam/launch/radar_graph_slam.launch:50-63). */
typedef struct apd_preprocess_params {
  int32_t use_distance_filter;      /* preprocessing_nodelet.cpp:201 (true)                                   */
  int32_t outlier_removal;          /* 0 = NONE, 1 = RADIUS (pcl::RadiusOutlierRemoval), 2 = STATISTICAL (pcl::StatisticalOutlierRemoval, the nodelet's code default :166) */
  int32_t radius_min_neighbors;     /* :178 (2)                                                                */
  int32_t statistical_mean_k;       /* :168 (20); at most 31                                                    */
  double distance_near_thresh;      /* :202 (1.0)   */
  double distance_far_thresh;       /* :203 (100.0) */
  double z_low_thresh;              /* :204 (-5.0)  */
  double z_high_thresh;             /* :205 (20.0)  */
  double downsample_resolution;     /* :138 (0.1); <= 0 = downsample_method NONE; otherwise the leaf of pcl::VoxelGrid, or of pcl::ApproximateVoxelGrid
                                       after apd_set_option(h, "downsample_method", 1) (:137-149; 0 = VOXELGRID, the default) */
  double radius_radius;             /* :177 (0.8)   */
  double statistical_stddev;        /* :169 (1.0)   */
} apd_preprocess_params;
int apd_default_preprocess_params(apd_preprocess_params* p);
/* distance_filter -> downsample -> outlier_removal (preprocessing_nodelet.cpp:812-815) of one cloud on the GPU.
 * Points are read as x,y,z at the start of every stride_bytes record and the intensity at intensity_offset_bytes inside it
 * (pcl::PointXYZI: stride 32, intensity offset 16; packed xyzi: stride 16, offset 12); the output uses the same record layout.
 * out must hold n records; *n_out receives the number of points kept. */
int apd_preprocess(apd_handle h, const float* points, int stride_bytes, int intensity_offset_bytes, int n, const apd_preprocess_params* p, float* out, int* n_out);
/* Submap accumulation (radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:606-616): clouds which[0..n_sel) of a cloud set, each moved by
 * its rel_pose (row-major double[16], pcl::transformPointCloud in double), concatenated, then pcl::VoxelGrid if downsample_resolution > 0.
 * The result becomes the handle's TARGET without leaving the device (setInputTarget(keyframe_cloud_s2m), :615); out_xyzi (host, packed
 * x y z intensity, capacity out_capacity points, may be NULL) receives a copy. The 4th float of the set's points is taken as intensity. */
/* Number of clouds and total number of points of a cloud set (sizes the out buffer of apd_build_submap). */
int apd_cloudset_info(apd_cloudset cs, int32_t* n_clouds, int64_t* total_points);
int apd_build_submap(apd_handle h, apd_cloudset keyframes, const int32_t* which, int n_sel, const double* rel_poses, double downsample_resolution,
                     uint64_t cache_key, float* out_xyzi, int out_capacity, int* n_out);

/* Scan-to-scan odometry over one host array of n_scans scans (config C2; the call pattern of
 * radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:449-468,584-592 replayed over a recorded drive): pair i registers scan
 * i+1 onto scan i, every scan is uploaded, gridded and given covariances once, out receives n_scans-1 records. guesses: (n_scans-1)*16
 * floats or NULL = identity. Like apd_batch_align it pipelines chunks over two streams so uploads overlap compute. */
int apd_odometry_align(apd_handle h, const float* pts, const int32_t* offsets /* host, n_scans+1 */, int n_scans, int stride_bytes,
                       const float* guesses, apd_result* out);

/* ---- introspection for benchmarks ---- */
/* The streaming kernels of the path at n_points, each timed `reps` times with CUDA events on the handle's stream, L2 evicted (a 256 MB
 * read) before every repetition: [0] pack_points (pcl::PointXYZI records -> float4), [1] transform_points (the output cloud,
 * lsq_registration_impl.hpp:79), [2] cov_export (getSource/TargetCovariances), [3] cov_import (setSource/TargetCovariances).
 * gbps = algorithmic bytes (48 / 28 / 192 / 192 per point) / time; ms (may be NULL) = time per launch. */
int apd_bench_streaming(apd_handle h, int n_points, int reps, double gbps[4], double ms[4]);
/* Point-to-point distance evaluations executed by the leaf-mode searches since the last call (and reset): kNN + covariance kernel,
 * and the 1-NN searches of the align kernel (correspondences + fitness). Every lane of a warp counts: a broadcast leaf scan is
 * 32 queries x 32 candidates, a transposed turn 32 candidates of one query. */
int apd_get_search_counters(apd_handle h, int64_t* knn_evals, int64_t* nn1_evals);
/* With apd_set_option(h, "kernel_timing", 1) every hot launch is bracketed by CUDA events on the stream it is launched on; this call
 * synchronises, returns the summed durations since the last call (ms) and the launch counts per kind - 0 pack_points, 1 grid / leaf
 * build, 2 kNN + covariance, 3 align (+ fitness) - and clears them. Chunks of a pipelined call that ran on the helper stream are included. */
int apd_get_kernel_times(apd_handle h, double ms[4], int64_t launches[4]);
/* Profiling aid: with apd_set_option(h, "timeline", 1) the align kernel records (phase, %globaltimer ns) stamps for the first pair of a
 * launch: 0 kernel entered, 1 target staged, 2 iteration starts, 3 correspondences done, 4 H/b reduced, 5 LM trial done, 6 fitness done. */
int apd_get_timeline(apd_handle h, uint64_t* phase_ns /* 2 values per stamp */, int max_stamps, int* n_stamps);
/* Same option: counters of the leaf-mode correspondence searches of the last align launch, [0..7] first (unseeded) pass, [8..15] seeded
 * passes: warp groups, leaves taken from the schedule, broadcast scans, transposed turns, most turns in one group, most leaves in one
 * group, queries. */
int apd_get_debug_counters(apd_handle h, uint64_t out[16]);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches claim). */
int apd_get_launch_count(apd_handle h, int64_t* n);
/* Iteration counters of the last apd_align_pairs call, summed over pairs: outer iterations (linearize
 * passes) and LM trials (compute_error passes); used for the algorithmic-bytes figure. */
int apd_get_work_counters(apd_handle h, int64_t* linearize_passes, int64_t* error_passes, int64_t* pairs);

#ifdef __cplusplus
}
#endif
#endif /* APDGICP_B200_H */
