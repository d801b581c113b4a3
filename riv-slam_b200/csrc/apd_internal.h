// Internal interface between the C ABI (apd_capi.cu) and the kernel launchers
// (apd_build.cu, apd_knn_cov.cu, apd_align.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/apdgicp_b200.h"
#include "apd_grid.cuh"

namespace apd {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a driver call of a microsecond or two in front of every launch of the
// single-pair path; the attribute only ever has to grow, so remember per (kernel, device) what was granted.
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kern, size_t bytes) {
  static std::mutex m;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  cudaGetDevice(&dev);
  const auto key = std::make_pair(reinterpret_cast<const void*>(kern), dev);
  {
    std::lock_guard<std::mutex> lk(m);
    auto it = granted.find(key);
    if (it != granted.end() && it->second >= bytes) return cudaSuccess;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(m);
  size_t& g = granted[key];
  if (g < bytes) g = bytes;
  return cudaSuccess;
}


// Coarser levels of the grid pyramid (cell edge x4 and x16): the same points sorted by a coarser cell
// id. A search that is not finished after a few rings of the fine grid restarts here instead of
// walking (2r+1)^2 mostly empty rows per ring through a sparse neighbourhood.
constexpr int kCoarseLevels = 2;
constexpr int kFineRings = 3;     // Top1 searches: rings tried on a level before going coarser
constexpr int kFineRingsKnn = 5;  // kNN: measured optimum (profiles/): coarse cells hold many points, so only sparse neighbourhoods go coarser
struct CoarseLevel {
  float4* spts;
  unsigned* cells;
  GridParams* grid;
  const long long* cell_off;
};

// Device-side description of a ragged batch of clouds ("cloud set"). All arrays live in HBM.
struct CloudSetView {
  int n_clouds;
  int total_points;
  const int* pt_off;          // [n_clouds+1] point offsets
  const long long* cell_off;  // [n_clouds+1] offsets into `cells`; cloud c owns grid[c].ncells+1 entries from cell_off[c]
  float4* pts;                // original order, (x, y, z, 1)
  float4* spts;               // cell-sorted, w = original local index (bit pattern)
  unsigned* cells;            // per-cloud exclusive prefix sums of cell populations (local indices)
  GridParams* grid;           // [n_clouds]
  double2* cov0;              // sorted order: (xx, xy)
  double2* cov1;              //               (xz, yy)
  double2* cov2;              //               (yz, zz)
  int* inv0;                  // original local index -> position in spts
  CoarseLevel coarse[kCoarseLevels];
  // leaf mode (apd_leaf.cuh; every cloud of the set fits shared memory): spts is in Hilbert order, cut into leaves of 32 points;
  // there is no cell table and no pyramid. lbox == nullptr: grid mode.
  float4* lbox;               // 2 float4 per leaf (lo, hi)
  const int* leaf_off;        // [n_clouds+1] offsets into lbox / 2 (cloud c owns ceil(n_c / 32) leaves)
  // The shared-memory IMAGE of every cloud, written by the leaf build: kLeafImage float4 per leaf, cloud c at limg + kLeafImage * leaf_off[c],
  // [nleaf * 32 float4: points in the pair layout of apd_leaf.cuh][nleaf * 2 float4: boxes] - byte for byte what the search kernels keep in
  // shared memory, so staging a cloud is ONE bulk asynchronous copy (cp.async.bulk -> mbarrier, apd_leaf.cuh leaf_stage_bulk) instead of
  // a load / negate / interleave / store loop through registers. nullptr: option "bulk_stage" off (the register path, leaf_stage).
  float4* limg;
};
constexpr int kLeafImage = 34;  // float4 per leaf in limg: 32 points + 2 box corners

#ifdef __CUDACC__
__device__ __forceinline__ GridView<unsigned> coarse_view(const CloudSetView& cs, int level, int cloud) {
  GridView<unsigned> G;
  const int base = cs.pt_off[cloud];
  G.spts = cs.coarse[level].spts + base;
  G.cells = cs.coarse[level].cells + cs.coarse[level].cell_off[cloud];
  G.g = cs.coarse[level].grid[cloud];
  G.n = cs.pt_off[cloud + 1] - base;
  return G;
}

// Stage one cloud's fine grid in shared memory: sorted points (16 B each) and the cell table narrowed to
// 16-bit entries. Loads are issued in independent batches (4 x 16 B per thread for the points, 8 x 16 B
// for the table, whose per-cloud offset is a multiple of four entries) so the copy is bandwidth- rather
// than latency-bound: on a single scan it is on the critical path of every CTA.
__device__ __forceinline__ void stage_grid(float4* __restrict__ s_pts, uint16_t* __restrict__ s_cells, const float4* __restrict__ gp,
                                           const unsigned* __restrict__ gc, int n, int ncells) {
  const int T = blockDim.x, t = threadIdx.x;
  for (int i = t; i < n; i += 4 * T) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i + u * T < n) v[u] = gp[i + u * T];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i + u * T < n) s_pts[i + u * T] = v[u];
  }
  const int total = ncells + 1;
  const int quads = total >> 2;
  const uint4* gc4 = reinterpret_cast<const uint4*>(gc);
  uint2* sc2 = reinterpret_cast<uint2*>(s_cells);
  for (int i = t; i < quads; i += 8 * T) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (i + u * T < quads) v[u] = gc4[i + u * T];
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (i + u * T < quads) sc2[i + u * T] = make_uint2(v[u].x | (v[u].y << 16), v[u].z | (v[u].w << 16));
  }
  for (int i = (quads << 2) + t; i < total; i += T) s_cells[i] = (uint16_t)gc[i];
}

// Exact unbounded (or gate-bounded) search through the pyramid: fine grid first, then coarser levels.
// A TopK visitor must be reset between levels (the same points would be offered twice); a Top1
// visitor keeps its key, which only tightens the bound. Returns the level that completed the search.
template <typename CellT, typename Visitor, bool RESET>
__device__ __forceinline__ int pyramid_search(const GridView<CellT>& G0, const CloudSetView& cs, int cloud, float qx, float qy, float qz, float limit2, Visitor& vis,
                                              int RINGS = kFineRings) {
  if (grid_search(G0, qx, qy, qz, limit2, vis, RINGS)) return 0;
  if (RESET) vis.init();
  const GridView<unsigned> G1 = coarse_view(cs, 0, cloud);
  if (grid_search(G1, qx, qy, qz, limit2, vis, RINGS)) return 1;
  if (RESET) vis.init();
  const GridView<unsigned> G2 = coarse_view(cs, 1, cloud);
  grid_search(G2, qx, qy, qz, limit2, vis);
  return 2;
}
#endif

struct DeviceParams {
  int k;
  int regularization;
  int max_iterations;
  int optimizer;
  int lm_max_iterations;
  float corr_limit2;        // float upper bound of max_corr_dist^2 for search pruning
  float corr_wide2;         // (1.4 * max_corr_dist)^2: radius of the first, anchoring search (see apd_align.cu)
  double corr_thr2;         // max_corr_dist^2 (double product, fast_apdgicp_impl.hpp:156)
  double rotation_epsilon;
  double transformation_epsilon;
  double lm_init_lambda_factor;
  double dist_var;
  double sin_az;            // sin(azimuth_var / 180 * pi)
  double sin_el;            // sin(elevation_var / 180 * pi)
  int knn_packed;           // kNN: use the packed 32-bit candidate list (tuning; results identical)
  int knn_fine_rings;       // kNN: rings tried on a pyramid level before restarting on the next coarser one
};

// Per concurrent pair slot scratch of the align kernel (sorted-source order).
struct AlignScratch {
  int max_src;        // points per slot
  int* corr;          // [slots * max_src] sorted position in the target, -1 = none
  float* sqd;         // [slots * max_src]
  double2* m0;        // Mahalanobis (xx, xy)
  double2* m1;        //             (xz, yy)
  double2* m2;        //             (yz, zz)
  float* fit;         // [slots * max_src] squared 1-NN distance at the final pose (fitness pass)
  float4* anchor;     // unmatched points: (query position, lower bound of the distance to ANY target point) of their last full search
};

enum TeamKind { TEAM_CTA = 0, TEAM_CLUSTER = 1, TEAM_GRID = 2 };
#ifndef APD_ALIGN_THREADS
#define APD_ALIGN_THREADS 512
#endif
constexpr int kAlignThreads = APD_ALIGN_THREADS;
#ifndef APD_ALIGN_MIN_BLOCKS
#define APD_ALIGN_MIN_BLOCKS 1   // CTAs per SM the align kernel is compiled for (experiments: 256 threads x 2)
#endif
#ifndef APD_KNN_THREADS
#define APD_KNN_THREADS 640
#endif
constexpr int kKnnThreads = APD_KNN_THREADS;
constexpr int kNRed = 30;  // doubles per reduction record (see apd_align.cu)

struct AlignBatch {
  CloudSetView src, tgt;
  const int* src_idx;      // [n_pairs] or nullptr (identity map)
  const int* tgt_idx;
  int src_base, tgt_base;  // without an index array pair p registers cloud p + src_base onto cloud p + tgt_base (odometry: 1 and 0)
  const float* guesses;    // [n_pairs*16] row-major or nullptr (identity)
  const double* guesses64; // [n_pairs*16] row-major double poses (the protected linearize / compute_error hooks take an Isometry3d); wins over guesses
  apd_result* out;         // [n_pairs] device
  double* final_hessian;   // [n_pairs*36] or nullptr
  double* lin_b;           // [n_pairs*6] or nullptr (mode 1: the gradient of evaluateCost)
  double* trace;           // [n_pairs * trace_rows * 8] or nullptr
  int* trace_count;        // [n_pairs] or nullptr
  int trace_rows;
  int n_pairs;
  int* work_counter;       // dynamic pair fetch (TEAM_CTA)
  unsigned long long* counters;  // [0] linearize passes, [1] error passes
  double* grid_partials;   // TEAM_GRID: [2 * gridDim.x * kNRed]
  AlignScratch scratch;
  DeviceParams prm;
  int min_points;          // pairs with a cloud smaller than this report APD_ERR_TOO_FEW_POINTS (k; 0 = no check)
  int mode;                // 0 = align, 1 = linearize only (evaluateCost): out->error, final_hessian, lin_b,
                           // 2 = fitness score only at the given pose (calc_fitness_score): out->fitness, out->T = pose,
                           // 3 = compute_error only at the given pose, with the correspondences and Mahalanobis matrices the
                           //     last linearize left in the slot (fast_apdgicp_impl.hpp:275-298): out->error
  double max_range;        // fitness gate (getFitnessScore max_range)
  unsigned long long* nn1_evals;  // nullable: += distance evaluations of the leaf-mode 1-NN searches (bench.py)
  unsigned long long* timeline;  // nullable (option "timeline"): thread 0 of team 0's first CTA stamps %globaltimer at phase boundaries, [0] = count
};

struct LaunchStats {
  long long launches = 0;
};

// ---- launchers (all asynchronous on `stream`) ----
struct BuildWorkspace {
  unsigned* bbox;    // [6 * n_clouds]
  int* cellid;       // [total_points]
  unsigned* cursor;  // [total_cells]
};
// tiles: device int4 (cloud, first point, count, 0) covering every point of the set
cudaError_t launch_grid_build(const CloudSetView& cs, const BuildWorkspace& ws, const int4* tiles, int n_tiles, const int* cell_cap /*device [n_clouds]*/,
                              long long total_cells, bool finest_level, cudaStream_t stream, LaunchStats* st);
// all pyramid levels of every cloud in ONE launch (clouds small enough for one CTA per cloud and level)
cudaError_t launch_grid_build_fused(const CloudSetView& cs, const int* const cap[1 + kCoarseLevels], int* const cellid[1 + kCoarseLevels],
                                    unsigned* const cursor[1 + kCoarseLevels], size_t smem_bytes /*0: global-memory version*/, cudaStream_t stream, LaunchStats* st);
// Hilbert order + leaf boxes of every cloud of a leaf-mode set (one CTA per cloud)
cudaError_t launch_leaf_build(const CloudSetView& cs, int max_n, cudaStream_t stream, LaunchStats* st, bool stamps = false);
cudaError_t leaf_build_stamps(unsigned long long out[16]);  // profiling aid
cudaError_t launch_knn_cov_leaf(const CloudSetView& cs, const int4* tiles, int n_tiles, int max_n, const DeviceParams& prm, int* knn_out /*nullable*/,
                                unsigned long long* evals /*nullable: += distance evaluations*/, cudaStream_t stream, LaunchStats* st, bool stamps = false);
cudaError_t knn_leaf_stamps(unsigned long long out[16]);  // profiling aid
cudaError_t launch_knn_cov(const CloudSetView& cs, const int4* tiles, int n_tiles, bool staged, size_t smem_bytes, const DeviceParams& prm,
                           int* knn_out /*nullable: total*k, original order rows*/, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_align(const AlignBatch& b, int team_kind, int team_size, int n_teams, bool stage_target, size_t smem_bytes, cudaStream_t stream,
                         LaunchStats* st);
// strict: count d2 < max_range (the status message's inlier test) instead of d2 <= max_range (getFitnessScore)
cudaError_t launch_fitness(const CloudSetView& src, int s, const CloudSetView& tgt, int t, const float* T16 /*device*/, double max_range, bool strict,
                           double* partials /*device [2*blocks]*/, int blocks, double* out /*device [2]: score, count*/, cudaStream_t stream, LaunchStats* st);
// out[0] = number of entries of d2[0..n) with (double)d2 < thr
cudaError_t launch_count_below(const float* d2, int n, double thr, double* out /*device [1]*/, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_pack_points(const float* xyz, int stride_floats, long long n, float4* out, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_transform_points(const float4* pts, int n, const float* T16 /*device*/, float* out, int out_stride_floats, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_l2_flush(const void* buf, size_t bytes, float* sink, cudaStream_t stream);  // evict L2 by READING a large buffer
cudaError_t launch_iota_w(float4* pts, int n, cudaStream_t stream, LaunchStats* st);  // pts[i].w = bits(i)
// gather/scatter between original and sorted order for the covariance getters / setters
cudaError_t launch_cov_export(const CloudSetView& cs, int cloud, double* out16 /*device n*16*/, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_cov_import(const CloudSetView& cs, int cloud, const double* in16 /*device n*16*/, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_corr_export(const AlignBatch& b, int slot, int s /*source cloud*/, int t /*target cloud*/, int* corr_out, float* sqd_out, double* m16_out, cudaStream_t stream, LaunchStats* st);

// ---- preprocessing filters and submap accumulation (apd_preprocess.cu); clouds are float4 x, y, z, intensity ----
cudaError_t launch_distance_filter(const float4* in, int n, double near_t, double far_t, double z_low, double z_high, int mode, unsigned char* flag, float4* out,
                                   int* n_out, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_voxel_grid(const float4* in, const int* n_in_dev, int n_in_host, float leaf, unsigned* ws_u32 /*4*n*/, int* seg_start /*n+1*/, float4* out, int* n_out,
                              cudaStream_t stream, LaunchStats* st);
cudaError_t launch_approx_voxel_grid(const float4* in, const int* n_in_dev, int n_in_host, float leaf, unsigned* ws_u32 /*4*n*/, int* seg_start /*n+1*/, float4* out, int* n_out,
                              cudaStream_t stream, LaunchStats* st);
cudaError_t launch_radius_flags(const CloudSetView& cs, int n, double radius, int min_pts, unsigned char* flag, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_statistical_flags(const CloudSetView& cs, int n, const int* n_dev, int mean_k, double stddev_mult, float* dist /*n*/, double* thr /*1*/, unsigned char* flag,
                                     cudaStream_t stream, LaunchStats* st);
cudaError_t launch_compact(const float4* in, const unsigned char* flag, const int* n_dev, float4* out, int* n_out, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_submap_gather(const CloudSetView& cs, const float4* xyzi, const int* which, const int* out_off, int n_sel, int total, const double* poses, float4* out,
                                 cudaStream_t stream, LaunchStats* st);
cudaError_t launch_pack_xyzi(const float* raw, int stride_floats, int intensity_offset, int n, float4* out, cudaStream_t stream, LaunchStats* st);
cudaError_t launch_unpack_xyzi(const float4* in, const int* n_dev, int n_max, int stride_floats, int intensity_offset, float* raw, cudaStream_t stream, LaunchStats* st);

size_t align_static_smem();
size_t knn_leaf_smem_bytes(int max_n);  // dynamic shared memory of the leaf-mode kNN kernel for clouds up to max_n points
int align_max_teams(int team_kind, int team_size, bool stage_target, size_t smem_bytes);

}  // namespace apd
