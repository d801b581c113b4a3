// C ABI of the B200-native FastAPDGICP path (include/apdgicp_b200.h): handle and cloud-set
// management, parameter plumbing and launch orchestration. No computation happens on the host:
// every entry point either moves bytes or enqueues the kernels of apd_build.cu / apd_knn_cov.cu /
// apd_align.cu on the handle's stream. There is no CPU fallback; without a device apd_create fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "apd_internal.h"
#include "apd_leaf.cuh"

using namespace apd;

namespace {

// Grow-only cache of device allocations shared by a handle and its cloud sets: cudaMalloc/cudaFree
// cost milliseconds and serialise the device, so blocks released by one step are handed to the next.
// All work of a handle is ordered on one stream, which makes reuse of a released block safe.
struct Pool {
  std::multimap<size_t, void*> free_;
  std::mutex m;
  ~Pool() {
    for (auto& kv : free_) cudaFree(kv.second);
  }
  static size_t bucket(size_t bytes) {
    const size_t g = bytes <= (64u << 10) ? 512 : (bytes <= (16u << 20) ? (64u << 10) : (1u << 20));
    return (bytes + g - 1) / g * g;
  }
  cudaError_t acquire(size_t bytes, void** p, size_t* cap) {
    const size_t want = bucket(bytes);
    {
      std::lock_guard<std::mutex> lk(m);
      auto it = free_.lower_bound(want);
      if (it != free_.end() && it->first <= want + want / 2 + (1u << 20)) {
        *p = it->second;
        *cap = it->first;
        free_.erase(it);
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {  // give cached blocks back to the driver and retry once
      cudaGetLastError();
      {
        std::lock_guard<std::mutex> lk(m);
        for (auto& kv : free_) cudaFree(kv.second);
        free_.clear();
      }
      e = cudaMalloc(p, want);
      if (e != cudaSuccess) return e;
    }
    *cap = want;
    return cudaSuccess;
  }
  void release(void* p, size_t cap) {
    std::lock_guard<std::mutex> lk(m);
    free_.emplace(cap, p);
  }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  std::shared_ptr<Pool> pool;
  ~DevBuf() { drop(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void drop() {
    if (!p) return;
    if (pool) pool->release(p, cap);
    else cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    drop();
    if (pool) return pool->acquire(bytes, &p, &cap);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return e; }
    cap = bytes;
    return cudaSuccess;
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct apd_cloudset_s {
  int n_clouds = 0;
  long long total = 0;
  int max_n = 0, min_n = 0, max_cap = 0;
  long long total_cells = 0;
  std::vector<int> h_off;
  DevBuf pts, spts, cells, grid, cov0, cov1, cov2, inv0;
  // the small per-cloud tables live in ONE device block filled by one copy from pinned staging memory (cloudset_layout)
  DevBuf tables;
  int* d_pt_off = nullptr;
  long long* d_cell_off = nullptr;
  int* d_cell_cap = nullptr;
  int4 *d_tiles_build = nullptr, *d_tiles_knn = nullptr;
  long long* d_c_cell_off[kCoarseLevels] = {nullptr, nullptr};
  int* d_c_cell_cap[kCoarseLevels] = {nullptr, nullptr};
  // coarse pyramid levels (cell budget / 64 and / 4096): own sorted copy, cell table and grid parameters
  DevBuf c_spts[kCoarseLevels], c_cells[kCoarseLevels], c_grid[kCoarseLevels];
  long long c_total_cells[kCoarseLevels] = {0, 0};
  int n_tiles_build = 0, n_tiles_knn = 0;
  bool grid_built = false, cov_valid = false;
  int cov_k = -1, cov_reg = -1;
  bool staged = false;      // LEAF mode (apd_leaf.cuh): every cloud fits the shared-memory staging area; Hilbert-sorted leaves, no cell tables
  size_t staged_smem = 0;   // bytes the align kernel stages for the largest cloud
  DevBuf lbox;              // leaf mode: 2 float4 per leaf
  DevBuf limg;              // leaf mode: the shared-memory image of every cloud (CloudSetView::limg), source of the bulk staging copies
  bool bulk_stage = true;   // handle option "bulk_stage" at the time the set was made
  int* d_leaf_off = nullptr;
  long long total_leaves = 0;
  explicit apd_cloudset_s(const std::shared_ptr<Pool>& pool) {
    for (DevBuf* b : {&pts, &spts, &cells, &grid, &cov0, &cov1, &cov2, &inv0, &tables, &lbox, &limg}) b->pool = pool;
    for (int l = 0; l < kCoarseLevels; l++)
      for (DevBuf* b : {&c_spts[l], &c_cells[l], &c_grid[l]}) b->pool = pool;
  }
  // the view the build kernels use to construct coarse level l (same points, that level's tables)
  CloudSetView level_view(int l) const {
    CloudSetView v = view();
    v.spts = c_spts[l].as<float4>();
    v.cells = c_cells[l].as<unsigned>();
    v.grid = c_grid[l].as<GridParams>();
    v.cell_off = d_c_cell_off[l];
    return v;
  }
  CloudSetView view() const {
    CloudSetView v;
    v.n_clouds = n_clouds;
    v.total_points = (int)total;
    v.pt_off = d_pt_off;
    v.cell_off = d_cell_off;
    v.pts = pts.as<float4>();
    v.spts = spts.as<float4>();
    v.cells = cells.as<unsigned>();
    v.grid = grid.as<GridParams>();
    v.cov0 = cov0.as<double2>();
    v.cov1 = cov1.as<double2>();
    v.cov2 = cov2.as<double2>();
    v.inv0 = inv0.as<int>();
    v.lbox = staged ? lbox.as<float4>() : nullptr;
    v.limg = staged && bulk_stage ? limg.as<float4>() : nullptr;
    v.leaf_off = d_leaf_off;
    for (int l = 0; l < kCoarseLevels; l++) {
      v.coarse[l].spts = c_spts[l].as<float4>();
      v.coarse[l].cells = c_cells[l].as<unsigned>();
      v.coarse[l].grid = c_grid[l].as<GridParams>();
      v.coarse[l].cell_off = d_c_cell_off[l];
    }
    return v;
  }
};

struct apd_context {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  apd_params prm;
  std::string err;
  std::shared_ptr<Pool> pool = std::make_shared<Pool>();
  std::shared_ptr<apd_cloudset_s> src, tgt;
  uint64_t src_key = 0, tgt_key = 0;
  LaunchStats stats;
  int sm_count = 0;
  size_t smem_optin = 0;
  // tunables (apd_set_option)
  double cells_per_point = 4.0;
  int team_size = 0;      // 0 = automatic
  int force_unstaged = 0;
  int max_teams_opt = 0;  // 0 = as many as fit
  int knn_packed = 1;
  int no_fused_build = 0;
  int no_smem_build = 0;
  double fitness_max_range = DBL_MAX;  // getFitnessScore(max_range) used by the batched calls
  int knn_fine_rings = kFineRingsKnn;
  int knn_leaf_parts = 0;  // 0 = automatic; 1, 2, 4, 8: warps per leaf in the leaf kNN kernel (experiments)
  int timeline_opt = 0;  // profiling aid: the align kernel stamps its phases (apd_get_timeline)
  // kernel timing (option "kernel_timing"): CUDA events around the hot launches, on the stream they are launched on
  int kernel_timing = 0;
  int pipeline_mem = APD_MEM_HOST;  // option "pipeline_device_input": apd_odometry_align / apd_batch_align read their point arrays from device memory
  bool bulk_stage = true;     // option "bulk_stage": leaf-mode clouds are staged by one cp.async.bulk from the image the build wrote (0: register path)
  int downsample_method = 0;  // option "downsample_method": 0 = VOXELGRID (pcl::VoxelGrid), 1 = APPROX_VOXELGRID (pcl::ApproximateVoxelGrid)
  struct TimedLaunch { int kind; cudaEvent_t e0, e1; };
  std::vector<TimedLaunch> timed;
  std::vector<cudaEvent_t> event_pool;
  // scratch (grow-only)
  DevBuf raw_upload, ws_bbox, ws_cellid, ws_cursor, sc_corr, sc_sqd, sc_m0, sc_m1, sc_m2, sc_anchor, sc_fit, results, guesses, idx_src, idx_tgt, fh, lin_b, trace, trace_count,
      counters, search_counters, grid_partials, misc, timeline, oneblock, knn_tmp, cov_tmp, pp_a, pp_b, pp_flag, pp_n, pp_ws, pp_seg, pp_tab;
  int scratch_slots = 0, scratch_max_src = 0;
  // last single-pair alignment
  bool has_last = false;
  apd_result last;
  double final_hessian[36];
  std::vector<double> lm_trace;
  bool last_lin_valid = false;  // scratch slot 0 holds correspondences of the current src/tgt
  bool fit_valid = false;       // scratch slot 0 also holds the squared 1-NN distances of the last align's final pose
  long long work_lin = 0, work_err = 0, work_pairs = 0;
  // pinned staging buffer for small host->device table uploads (grow-only) and the event that guards its reuse
  unsigned char* stage_host[2] = {nullptr, nullptr};  // slot 0: cloud-set tables, slot 1: small point uploads
  size_t stage_cap[2] = {0, 0};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  std::map<std::tuple<int, int, bool, size_t>, int> team_fit;  // cached occupancy answers (max_teams_cached)
  unsigned char* down_host = nullptr;  // pinned landing area of apd_align's result block
  apd_handle helper = nullptr;   // second stream/pool for pipelined batches (pipelined_align)
  // upload-ahead of pipelined host batches: the whole input array crosses PCIe on a copy stream of its own, one event per chunk
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  DevBuf raw_all[2];
  bool upload_ahead = true;      // option "upload_ahead"
  int pipe_chunks = 4, pipe_first_waves = 1;   // options "pipeline_chunks" / "pipeline_first_waves" (pipelined_align)
  bool kernel_uploads = false;      // stage_upload fetches its block with a kernel instead of the copy engine (set while a call uploads ahead)
  cudaEvent_t fill_wait = nullptr;  // make_cloudset: the stream waits for this event between the table upload and the first read of the points
  long long helper_launches_seen = 0;
};

namespace {

thread_local std::string g_err;

int fail(apd_handle h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_err = msg;
  return code;
}

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      cudaGetLastError();                                                                            \
      return fail(h, APD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
    }                                                                                                \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Brackets one hot launch with CUDA events on the handle's stream when kernel timing is on (bench.py's per-kernel figures).
// kind: 0 pack_points, 1 grid / leaf build, 2 kNN + covariance, 3 align (+ fitness)
struct KernelTimer {
  apd_handle h;
  int slot = -1;
  KernelTimer(apd_handle h_, int kind) : h(h_) {
    if (!h->kernel_timing) return;
    cudaEvent_t ev[2];
    for (auto& e : ev) {
      if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
      else if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
    }
    cudaEventRecord(ev[0], h->stream);
    h->timed.push_back({kind, ev[0], ev[1]});
    slot = (int)h->timed.size() - 1;
  }
  ~KernelTimer() {
    if (slot >= 0) cudaEventRecord(h->timed[slot].e1, h->stream);
  }
};

DeviceParams device_params(const apd_params& p) {
  DeviceParams d;
  d.k = p.k_correspondences;
  d.regularization = p.regularization;
  d.max_iterations = p.max_iterations;
  d.optimizer = p.optimizer;
  d.lm_max_iterations = p.lm_max_iterations;
  d.corr_thr2 = p.max_corr_dist * p.max_corr_dist;  // product in double (fast_apdgicp_impl.hpp:156)
  // smallest float not below the double threshold: pruning with it can never drop a candidate that passes the gate
  float f = (float)d.corr_thr2;
  if (!(d.corr_thr2 < 3.0e38)) f = INFINITY;
  else if ((double)f < d.corr_thr2) f = std::nextafterf(f, INFINITY);
  d.corr_limit2 = f;
  d.corr_wide2 = f < 3.0e38f ? f * 1.96f : f;
  d.rotation_epsilon = p.rotation_epsilon;
  d.transformation_epsilon = p.transformation_epsilon;
  d.lm_init_lambda_factor = p.lm_init_lambda_factor;
  d.dist_var = p.dist_var;
  d.sin_az = std::sin(p.azimuth_var / 180 * M_PI);
  d.sin_el = std::sin(p.elevation_var / 180 * M_PI);
  d.knn_packed = 1;
  d.knn_fine_rings = kFineRingsKnn;
  return d;
}

// shared memory a staged grid may take: the opt-in maximum minus the align kernel's static state and
// minus room for the kNN kernel's per-thread neighbour lists (up to 32 two-byte entries per thread)
size_t staging_budget(apd_handle h) { return h->smem_optin - align_static_smem() - 512 - 32 * sizeof(uint16_t) * kKnnThreads; }

// Copy `bytes` of small tables to the device through the handle's pinned staging buffer: `fill` writes them into the
// buffer, one asynchronous copy moves them. The buffer is reused by the next call, so its previous copy must have
// completed (an event, normally long signalled); no stream synchronisation is needed and the caller's own host
// vectors may die at once.
// While pipelined_align's copy stream keeps the host-to-device copy engine busy with the call's points, a small copy submitted on
// another stream is not served before that stream runs dry (measured: a chunk's 2 KB table block waited for 110 MB of points). The
// staging buffer is page-locked, i.e. readable by the SMs through the unified address space: such blocks are then fetched by a
// kernel (option / handle flag kernel_uploads), which needs no copy engine at all.
__global__ void host_block_fetch_kernel(unsigned* __restrict__ dst, const unsigned* __restrict__ src_host, size_t n_words) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_words) dst[i] = src_host[i];
}

template <typename Fill>
int stage_upload(apd_handle h, int slot, size_t bytes, Fill fill, void* dst) {
  if (bytes == 0) return APD_OK;
  if (!h->stage_ev[slot]) CK(cudaEventCreateWithFlags(&h->stage_ev[slot], cudaEventDisableTiming));
  else CK(cudaEventSynchronize(h->stage_ev[slot]));
  if (bytes > h->stage_cap[slot]) {
    if (h->stage_host[slot]) cudaFreeHost(h->stage_host[slot]);
    h->stage_host[slot] = nullptr;
    h->stage_cap[slot] = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 2, 64u << 10);
    CK(cudaMallocHost(reinterpret_cast<void**>(&h->stage_host[slot]), want));
    h->stage_cap[slot] = want;
  }
  fill(h->stage_host[slot]);
  if (h->kernel_uploads && bytes % 4 == 0 && bytes <= (256u << 10)) {
    const size_t words = bytes / 4;
    host_block_fetch_kernel<<<(unsigned)((words + 255) / 256), 256, 0, h->stream>>>(static_cast<unsigned*>(dst), reinterpret_cast<const unsigned*>(h->stage_host[slot]), words);
    h->stats.launches++;
    CK(cudaGetLastError());
  } else {
    CK(cudaMemcpyAsync(dst, h->stage_host[slot], bytes, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaEventRecord(h->stage_ev[slot], h->stream));
  return APD_OK;
}

// Build the device-side description of a ragged batch: offsets, per-cloud cell budgets (grid mode) or leaf offsets (leaf mode), tile lists.
// force_grid: the caller needs the voxel grid whatever the size (the pre-processing filters count neighbours on it).
int cloudset_layout(apd_handle h, apd_cloudset_s* cs, bool force_grid = false) {
  const int nc = cs->n_clouds;
  std::vector<int> cap(nc);
  std::vector<long long> cell_off(nc + 1);
  cs->max_n = 0;
  cs->min_n = nc ? INT32_MAX : 0;
  for (int c = 0; c < nc; c++) {
    const int n = cs->h_off[c + 1] - cs->h_off[c];
    cs->max_n = std::max(cs->max_n, n);
    cs->min_n = std::min(cs->min_n, n);
  }
  // LEAF mode (apd_leaf.cuh): every cloud of the set, staged as 32-point leaves, fits the shared memory of both the align kernel
  // (staged target + its static state) and the kNN kernel (staged cloud + per-lane pending lists)
  const size_t stage_bytes = (size_t)((cs->max_n + kLeaf - 1) / kLeaf) * (kLeaf * 16 + 32);
  const bool leaf = nc > 0 && !force_grid && !h->force_unstaged && cs->max_n <= kLeafMaxPoints && stage_bytes + align_static_smem() + 1024 <= h->smem_optin &&
                    knn_leaf_smem_bytes(cs->max_n) + 14336 <= h->smem_optin;  // + the kNN kernel's static query slots (12.8 KB)
  cs->staged = leaf;
  cs->staged_smem = leaf ? ((stage_bytes + 15) & ~(size_t)15) : 0;
  for (int c = 0; c < nc; c++) {
    const int n = cs->h_off[c + 1] - cs->h_off[c];
    long long want = leaf ? 1 : std::max<long long>(64, (long long)std::llround(h->cells_per_point * (double)n));
    cap[c] = (int)std::min<long long>(want, 1ll << 28);
  }
  cs->max_cap = 0;
  for (int c = 0; c < nc; c++) cs->max_cap = std::max(cs->max_cap, cap[c]);
  long long off = 0;
  for (int c = 0; c < nc; c++) {
    cell_off[c] = off;
    off += leaf ? 0 : (((long long)cap[c] + 1 + 3) & ~3ll);  // 16-byte aligned tables
  }
  cell_off[nc] = off;
  cs->total_cells = off;
  // coarse pyramid levels (grid mode): 64x and 4096x fewer cells than the fine budget
  std::vector<int> ccap[kCoarseLevels];
  std::vector<long long> ccell_off[kCoarseLevels];
  for (int l = 0; l < kCoarseLevels; l++) {
    ccap[l].resize(nc);
    ccell_off[l].resize(nc + 1);
    long long o = 0;
    for (int c = 0; c < nc; c++) {
      const int n = cs->h_off[c + 1] - cs->h_off[c];
      const long long fine = std::min<long long>(std::max<long long>(64, (long long)std::llround(h->cells_per_point * (double)n)), 1ll << 28);
      ccap[l][c] = (int)std::max<long long>(8, fine >> (6 * (l + 1)));
      ccell_off[l][c] = o;
      o += leaf ? 0 : (long long)ccap[l][c] + 1;
    }
    ccell_off[l][nc] = o;
    cs->c_total_cells[l] = o;
  }
  // leaf offsets (leaf mode)
  std::vector<int> leaf_off(nc + 1, 0);
  for (int c = 0; c < nc; c++) leaf_off[c + 1] = leaf_off[c] + (leaf ? (cs->h_off[c + 1] - cs->h_off[c] + kLeaf - 1) / kLeaf : 0);
  cs->total_leaves = leaf_off[nc];

  // tiles for the per-point build kernels (grid mode): 1024 points per CTA
  std::vector<int4> tb;
  if (!leaf)
    for (int c = 0; c < nc; c++) {
      const int n = cs->h_off[c + 1] - cs->h_off[c];
      for (int s = 0; s < n; s += 1024) tb.push_back(int4{c, s, std::min(1024, n - s), 0});
    }
  // tiles for kNN + covariance: whole clouds per CTA when the batch alone fills the GPU, otherwise split so that one wave of CTAs
  // exists. Grid mode: (cloud, first query, queries); leaf mode: (cloud, first leaf, leaves).
  std::vector<int4> tk;
  const long long target_ctas = h->sm_count;  // one CTA per SM fits (shared memory): a single wave
  if (leaf) {
    const long long per = nc >= target_ctas ? (1ll << 30) : std::max<long long>((cs->total_leaves + target_ctas - 1) / std::max<long long>(target_ctas, 1), 1);
    // fewer leaves than warps (a single scan): every leaf's queries are cut into `sub` parts, one warp each (knn_cov_leaf_kernel)
    int sub = 1;
    while (sub < 4 && cs->total_leaves * sub * 2 <= (long long)h->sm_count * 20) sub *= 2;  // measured on one 5000-point scan: 79 / 73 / 68 / 67 us for 1 / 2 / 4 / 8 parts
    if (h->knn_leaf_parts > 0) sub = h->knn_leaf_parts;
    for (int c = 0; c < nc; c++) {
      const int nl = leaf_off[c + 1] - leaf_off[c];
      for (long long s = 0; s < nl; s += per) tk.push_back(int4{c, (int)s, (int)std::min<long long>(per, nl - s), sub});
    }
  } else {
    long long tile_q = nc >= target_ctas ? (long long)cs->max_n : std::max<long long>((cs->total + target_ctas - 1) / std::max<long long>(target_ctas, 1), 16);
    tile_q = std::max<long long>(tile_q, 1);
    for (int c = 0; c < nc; c++) {
      const int n = cs->h_off[c + 1] - cs->h_off[c];
      for (long long s = 0; s < n; s += tile_q) tk.push_back(int4{c, (int)s, (int)std::min<long long>(tile_q, n - s), 0});
    }
  }
  cs->n_tiles_build = (int)tb.size();
  cs->n_tiles_knn = (int)tk.size();

  // one device block for the small tables, one copy from the handle's pinned staging buffer: a dozen synchronous
  // copies from pageable vectors plus a stream synchronisation cost ~50 us per cloud set, most of setInputSource
  struct Part { const void* src; size_t bytes; size_t off; };
  Part parts[6 + 2 * kCoarseLevels];
  int np_ = 0;
  size_t off_b = 0;
  auto add = [&](const void* src, size_t bytes) {
    parts[np_] = Part{src, bytes, off_b};
    off_b = (off_b + std::max<size_t>(bytes, 16) + 15) & ~(size_t)15;
    return np_++;
  };
  const int i_pt = add(cs->h_off.data(), sizeof(int) * (nc + 1));
  const int i_co = add(cell_off.data(), sizeof(long long) * (nc + 1));
  const int i_cc = add(cap.data(), sizeof(int) * nc);
  const int i_tb = add(tb.data(), sizeof(int4) * tb.size());
  const int i_tk = add(tk.data(), sizeof(int4) * tk.size());
  const int i_lf = add(leaf_off.data(), sizeof(int) * (nc + 1));
  int i_lo[kCoarseLevels], i_lc[kCoarseLevels];
  for (int l = 0; l < kCoarseLevels; l++) {
    i_lo[l] = add(ccell_off[l].data(), sizeof(long long) * (nc + 1));
    i_lc[l] = add(ccap[l].data(), sizeof(int) * nc);
  }
  CK(cs->tables.reserve(off_b));
  unsigned char* dbase = cs->tables.as<unsigned char>();
  cs->d_pt_off = reinterpret_cast<int*>(dbase + parts[i_pt].off);
  cs->d_cell_off = reinterpret_cast<long long*>(dbase + parts[i_co].off);
  cs->d_cell_cap = reinterpret_cast<int*>(dbase + parts[i_cc].off);
  cs->d_tiles_build = reinterpret_cast<int4*>(dbase + parts[i_tb].off);
  cs->d_tiles_knn = reinterpret_cast<int4*>(dbase + parts[i_tk].off);
  cs->d_leaf_off = reinterpret_cast<int*>(dbase + parts[i_lf].off);
  for (int l = 0; l < kCoarseLevels; l++) {
    cs->d_c_cell_off[l] = reinterpret_cast<long long*>(dbase + parts[i_lo[l]].off);
    cs->d_c_cell_cap[l] = reinterpret_cast<int*>(dbase + parts[i_lc[l]].off);
  }
  int rc_stage = stage_upload(h, 0, off_b, [&](unsigned char* host) {
    for (int i = 0; i < np_; i++)
      if (parts[i].bytes) memcpy(host + parts[i].off, parts[i].src, parts[i].bytes);
  }, dbase);
  if (rc_stage) return rc_stage;
  const size_t np1 = (size_t)std::max<long long>(cs->total, 1);
  CK(cs->grid.reserve(sizeof(GridParams) * std::max(nc, 1)));
  CK(cs->pts.reserve(sizeof(float4) * np1));
  CK(cs->spts.reserve(sizeof(float4) * np1));
  CK(cs->cov0.reserve(sizeof(double2) * np1));
  CK(cs->cov1.reserve(sizeof(double2) * np1));
  CK(cs->cov2.reserve(sizeof(double2) * np1));
  CK(cs->inv0.reserve(sizeof(int) * np1));
  if (leaf) {
    CK(cs->lbox.reserve(sizeof(float4) * 2 * (size_t)std::max<long long>(cs->total_leaves, 1)));
    cs->bulk_stage = h->bulk_stage;
    if (cs->bulk_stage) CK(cs->limg.reserve(sizeof(float4) * kLeafImage * (size_t)std::max<long long>(cs->total_leaves, 1)));
    return APD_OK;  // no cell tables, no pyramid
  }
  CK(cs->cells.reserve(sizeof(unsigned) * (size_t)std::max<long long>(cs->total_cells, 1)));
  for (int l = 0; l < kCoarseLevels; l++) {
    CK(cs->c_spts[l].reserve(sizeof(float4) * np1));
    CK(cs->c_cells[l].reserve(sizeof(unsigned) * (size_t)std::max<long long>(cs->c_total_cells[l], 1)));
    CK(cs->c_grid[l].reserve(sizeof(GridParams) * std::max(nc, 1)));
  }
  return APD_OK;
}

int cloudset_fill(apd_handle h, apd_cloudset_s* cs, const float* xyz, int stride_bytes, int mem) {
  const long long n = cs->total;
  if (n == 0) return APD_OK;
  // a small host cloud (one scan) goes through the pinned staging buffer: a host memcpy plus a truly asynchronous copy
  // instead of the driver's blocking pageable path
  constexpr size_t kStageMax = 1u << 20;
  if (stride_bytes == 16) {
    const size_t bytes = sizeof(float4) * (size_t)n;
    if (mem == APD_MEM_HOST && bytes <= kStageMax) return stage_upload(h, 1, bytes, [&](unsigned char* host) { memcpy(host, xyz, bytes); }, cs->pts.p);
    CK(cudaMemcpyAsync(cs->pts.p, xyz, bytes, mem == APD_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
    return APD_OK;
  }
  const float* dev_xyz = xyz;
  if (mem == APD_MEM_HOST) {
    const size_t bytes = (size_t)(n - 1) * stride_bytes + 12;
    CK(h->raw_upload.reserve(bytes + 16));
    if (bytes <= kStageMax) {
      int rc = stage_upload(h, 1, bytes, [&](unsigned char* host) { memcpy(host, xyz, bytes); }, h->raw_upload.p);
      if (rc) return rc;
    } else {
      CK(cudaMemcpyAsync(h->raw_upload.p, xyz, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    dev_xyz = h->raw_upload.as<float>();
  }
  KernelTimer kt(h, 0);
  CK(launch_pack_points(dev_xyz, stride_bytes / 4, n, cs->pts.as<float4>(), h->stream, &h->stats));
  return APD_OK;
}

constexpr int kFusedBuildMaxPoints = 16384;  // clouds up to this size are built by one CTA per pyramid level

int cloudset_build_grid(apd_handle h, apd_cloudset_s* cs) {
  if (cs->grid_built || cs->n_clouds == 0) { cs->grid_built = true; return APD_OK; }
  KernelTimer kt(h, 1);
  if (cs->staged) {  // leaf mode: Hilbert order + leaf boxes, one CTA per cloud
    CK(launch_leaf_build(cs->view(), cs->max_n, h->stream, &h->stats, h->timeline_opt != 0));
    cs->grid_built = true;
    return APD_OK;
  }
  if (cs->max_n <= kFusedBuildMaxPoints && !h->no_fused_build) {
    // one launch for every level of every cloud
    const size_t np = (size_t)std::max<long long>(cs->total, 1);
    CK(h->ws_cellid.reserve(sizeof(int) * np * (1 + kCoarseLevels)));
    size_t cur_total = (size_t)cs->total_cells;
    for (int l = 0; l < kCoarseLevels; l++) cur_total += (size_t)cs->c_total_cells[l];
    CK(h->ws_cursor.reserve(sizeof(unsigned) * cur_total));
    const int* cap[1 + kCoarseLevels];
    int* cellid[1 + kCoarseLevels];
    unsigned* cursor[1 + kCoarseLevels];
    cap[0] = cs->d_cell_cap;
    cellid[0] = h->ws_cellid.as<int>();
    cursor[0] = h->ws_cursor.as<unsigned>();
    size_t off = (size_t)cs->total_cells;
    for (int l = 0; l < kCoarseLevels; l++) {
      cap[l + 1] = cs->d_c_cell_cap[l];
      cellid[l + 1] = h->ws_cellid.as<int>() + np * (l + 1);
      cursor[l + 1] = h->ws_cursor.as<unsigned>() + off;
      off += (size_t)cs->c_total_cells[l];
    }
    // shared-memory build when the largest cloud's sorted points and fine cell table fit one CTA
    size_t build_smem = sizeof(float4) * (size_t)cs->max_n + sizeof(unsigned) * ((size_t)cs->max_cap + 4) + 16;
    if (build_smem + 2048 > h->smem_optin || h->no_smem_build) build_smem = 0;
    CK(launch_grid_build_fused(cs->view(), cap, cellid, cursor, build_smem, h->stream, &h->stats));
    cs->grid_built = true;
    return APD_OK;
  }
  CK(h->ws_bbox.reserve(sizeof(unsigned) * 6 * cs->n_clouds));
  CK(h->ws_cellid.reserve(sizeof(int) * std::max<long long>(cs->total, 1)));
  CK(h->ws_cursor.reserve(sizeof(unsigned) * (size_t)cs->total_cells));
  BuildWorkspace ws{h->ws_bbox.as<unsigned>(), h->ws_cellid.as<int>(), h->ws_cursor.as<unsigned>()};
  CK(launch_grid_build(cs->view(), ws, cs->d_tiles_build, cs->n_tiles_build, cs->d_cell_cap, cs->total_cells, true, h->stream, &h->stats));
  for (int l = 0; l < kCoarseLevels; l++)
    CK(launch_grid_build(cs->level_view(l), ws, cs->d_tiles_build, cs->n_tiles_build, cs->d_c_cell_cap[l], cs->c_total_cells[l], false, h->stream,
                         &h->stats));
  cs->grid_built = true;
  return APD_OK;
}

int cloudset_prepare(apd_handle h, apd_cloudset_s* cs, int* knn_out = nullptr) {
  const int k = h->prm.k_correspondences;
  if (k < 1 || k > 32) return fail(h, APD_ERR_UNSUPPORTED, "k_correspondences must be in [1, 32]");
  if (h->prm.regularization < 0 || h->prm.regularization > 4) return fail(h, APD_ERR_INVALID, "unknown regularization method");
  int rc = cloudset_build_grid(h, cs);
  if (rc) return rc;
  if (cs->cov_valid && cs->cov_k == k && cs->cov_reg == h->prm.regularization && !knn_out) return APD_OK;
  if (cs->cov_valid && cs->cov_k < 0 && !knn_out) return APD_OK;  // covariances injected by the caller (setSource/TargetCovariances)
  DeviceParams dp = device_params(h->prm);
  dp.knn_packed = h->knn_packed;
  dp.knn_fine_rings = h->knn_fine_rings;
  KernelTimer kt(h, 2);
  if (!h->search_counters.p) {
    CK(h->search_counters.reserve(sizeof(unsigned long long) * 4));
    CK(cudaMemsetAsync(h->search_counters.p, 0, sizeof(unsigned long long) * 4, h->stream));
  }
  if (cs->staged) CK(launch_knn_cov_leaf(cs->view(), cs->d_tiles_knn, cs->n_tiles_knn, cs->max_n, dp, knn_out, h->search_counters.as<unsigned long long>(), h->stream, &h->stats, h->timeline_opt != 0));
  else CK(launch_knn_cov(cs->view(), cs->d_tiles_knn, cs->n_tiles_knn, false, 0, dp, knn_out, h->stream, &h->stats));
  cs->cov_valid = true;
  cs->cov_k = k;
  cs->cov_reg = h->prm.regularization;
  return APD_OK;
}

int make_cloudset(apd_handle h, const float* xyz, int stride_bytes, const int32_t* offsets, int n_clouds, int mem, std::shared_ptr<apd_cloudset_s>* out,
                  bool force_grid = false) {
  if (n_clouds < 0 || (n_clouds > 0 && !offsets)) return fail(h, APD_ERR_INVALID, "bad cloud offsets");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, APD_ERR_INVALID, "stride_bytes must be a multiple of 4 and at least 12");
  auto cs = std::make_shared<apd_cloudset_s>(h->pool);
  cs->n_clouds = n_clouds;
  cs->h_off.assign(n_clouds + 1, 0);
  for (int c = 0; c <= n_clouds && n_clouds > 0; c++) {
    cs->h_off[c] = offsets[c] - offsets[0];
    if (c && cs->h_off[c] < cs->h_off[c - 1]) return fail(h, APD_ERR_INVALID, "cloud offsets must be non-decreasing");
  }
  cs->total = cs->h_off[n_clouds];
  if (cs->total > 0 && !xyz) return fail(h, APD_ERR_INVALID, "null point pointer");
  int rc = cloudset_layout(h, cs.get(), force_grid);
  if (rc) return rc;
  const float* first = xyz ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(xyz) + (size_t)(n_clouds ? offsets[0] : 0) * stride_bytes) : nullptr;
  if (h->fill_wait) {   // upload-ahead (pipelined_align): the tables above are already on their way; only the points wait for the copy stream
    cudaEvent_t e = h->fill_wait;
    h->fill_wait = nullptr;
    CK(cudaStreamWaitEvent(h->stream, e, 0));
  }
  rc = cloudset_fill(h, cs.get(), first, stride_bytes, mem);
  if (rc) return rc;
  *out = cs;
  return APD_OK;
}

int ensure_align_scratch(apd_handle h, int slots, int max_src) {
  if (slots <= h->scratch_slots && max_src <= h->scratch_max_src) return APD_OK;
  slots = std::max(slots, h->scratch_slots);
  max_src = std::max(max_src, h->scratch_max_src);
  const size_t n = (size_t)slots * max_src;
  CK(h->sc_corr.reserve(sizeof(int) * n));
  CK(h->sc_sqd.reserve(sizeof(float) * n));
  CK(h->sc_m0.reserve(sizeof(double2) * n));
  CK(h->sc_m1.reserve(sizeof(double2) * n));
  CK(h->sc_m2.reserve(sizeof(double2) * n));
  CK(h->sc_anchor.reserve(sizeof(float4) * n));
  CK(h->sc_fit.reserve(sizeof(float) * n));
  h->scratch_slots = slots;
  h->scratch_max_src = max_src;
  return APD_OK;
}

struct TeamPlan {
  int kind = TEAM_CTA, size = 1, teams = 1;
  bool staged = false;
  size_t smem = 0;
};

// Pick the launch shape for n_pairs pairs: one CTA per pair when the batch fills the GPU, clusters
// (or the whole cooperative grid for very large sources) when few pairs must be made fast.
// plan_for_pairs: the batch size the team shape is chosen for (a pipelined call plans once for the whole
// batch so that every chunk sums in the same order and the records do not depend on the chunking)
// occupancy queries are driver calls (several microseconds each) and their answers never change for a handle: remember them
int max_teams_cached(apd_handle h, int kind, int size, bool staged, size_t smem) {
  const auto key = std::make_tuple(kind, size, staged, smem);
  auto it = h->team_fit.find(key);
  if (it != h->team_fit.end()) return it->second;
  const int fit = align_max_teams(kind, size, staged, smem);
  h->team_fit.emplace(key, fit);
  return fit;
}

int plan_teams(apd_handle h, const apd_cloudset_s* src, const apd_cloudset_s* tgt, int n_pairs, int plan_for_pairs, TeamPlan* plan) {
  TeamPlan p;
  p.staged = tgt->staged;
  p.smem = p.staged ? tgt->staged_smem : 0;
  int size = h->team_size;
  if (size <= 0) {
    size = 1;
    while (size < 16 && (long long)std::max(n_pairs, plan_for_pairs) * size * 2 <= h->sm_count) size *= 2;  // 16 = the non-portable cluster maximum
    // a CTA pass handles kAlignThreads points at a time: no point in more CTAs than that
    while (size > 1 && (long long)src->max_n < (long long)(size / 2) * kAlignThreads) size /= 2;
  }
  const bool force_grid_team = h->team_size >= 1000;  // option team_size = 1000 + n: a cooperative grid of up to n CTAs on one pair (experiments)
  if (force_grid_team) size = 1;
  // One pair: a cooperative grid. Measured on a 5000-point pair (scripts/latency_probe.py): 16-CTA cluster 147 us, grid of 74 CTAs 118 us,
  // of 148 CTAs 123 us per align; a warp then owns a handful of queries and the transposed leaf scan costs what those queries need.
  const bool single_staged = p.staged && h->team_size <= 0 && n_pairs == 1 && plan_for_pairs <= 1 && src->max_n >= 256;
  if ((force_grid_team && n_pairs == 1) || single_staged || (!p.staged && h->team_size <= 0 && n_pairs == 1 && src->max_n > 16 * kAlignThreads * 2)) {
    // large source against a large target: the whole GPU on one pair
    const int max_blocks = max_teams_cached(h, TEAM_GRID, 0, p.staged, p.smem);
    if (max_blocks >= 2) {
      p.kind = TEAM_GRID;
      p.size = std::min(max_blocks, (src->max_n + kAlignThreads - 1) / kAlignThreads);
      if (force_grid_team) p.size = std::max(2, std::min(max_blocks, h->team_size - 1000));
      else if (single_staged) p.size = std::max(2, std::min(max_blocks, (src->max_n + 63) / 64));  // ~64 points per CTA: 4 queries per warp
      p.teams = 1;
      *plan = p;
      return APD_OK;
    }
  }
  if (size > 1) {
    if (size > 16) size = 16;
    int fit = max_teams_cached(h, TEAM_CLUSTER, size, p.staged, p.smem);
    while (fit < 1 && size > 1) {
      size /= 2;
      fit = size > 1 ? max_teams_cached(h, TEAM_CLUSTER, size, p.staged, p.smem) : 0;
    }
    if (size > 1) {
      p.kind = TEAM_CLUSTER;
      p.size = size;
      p.teams = std::max(1, std::min(n_pairs, fit));
      if (h->max_teams_opt > 0) p.teams = std::min(p.teams, h->max_teams_opt);
      *plan = p;
      return APD_OK;
    }
  }
  const int fit = max_teams_cached(h, TEAM_CTA, 1, p.staged, p.smem);
  if (fit < 1) return fail(h, APD_ERR_CUDA, "align kernel does not fit on this device");
  p.kind = TEAM_CTA;
  p.size = 1;
  p.teams = std::max(1, std::min(n_pairs, fit));
  if (h->max_teams_opt > 0) p.teams = std::min(p.teams, h->max_teams_opt);
  *plan = p;
  return APD_OK;
}

struct AlignCall {
  apd_cloudset_s *src, *tgt;
  const int32_t *src_idx = nullptr, *tgt_idx = nullptr;  // host
  int src_base = 0, tgt_base = 0;                        // without index arrays: pair p = cloud p + src_base onto cloud p + tgt_base
  const float* guesses = nullptr;                        // host, n_pairs*16
  const double* guesses64 = nullptr;                     // host, n_pairs*16 (wins over guesses)
  int n_pairs = 0;
  int mode = 0;
  int min_points = 0;
  int plan_for_pairs = 0;
  double max_range = DBL_MAX;
  bool want_trace = false;
  bool want_hessian = false;
  // single-pair calls: result record, work counters, trace count, final Hessian, b and the trace rows live in ONE device block
  // (layout kOB_*), cleared by one memset and fetched by one copy into pinned memory
  unsigned char* block = nullptr;
  int block_trace_rows = 0;
};

constexpr size_t kOB_result = 0, kOB_counters = 128, kOB_tcount = 160, kOB_fh = 176, kOB_linb = 464, kOB_trace = 512;
constexpr int kOB_fast_rows = 40;  // trace rows fetched together with the header (a registration rarely has more LM trials)
static_assert(sizeof(apd_result) <= kOB_counters, "result record must fit its slot");

// the device block and its pinned landing area for a single-pair call (AlignCall::block)
int single_block(apd_handle h, AlignCall* c) {
  c->block_trace_rows = std::max(1, std::min(4096, h->prm.max_iterations * std::max(1, h->prm.lm_max_iterations)));
  CK(h->oneblock.reserve(kOB_trace + sizeof(double) * 8 * (size_t)c->block_trace_rows));
  c->block = h->oneblock.as<unsigned char>();
  if (!h->down_host) CK(cudaHostAlloc(reinterpret_cast<void**>(&h->down_host), kOB_trace + sizeof(double) * 8 * kOB_fast_rows, cudaHostAllocDefault));
  return APD_OK;
}

// Enqueue the align kernel for a batch; results land in h->results (device).
int run_align(apd_handle h, const AlignCall& c, AlignBatch* used = nullptr, TeamPlan* used_plan = nullptr) {
  // mode 2 (fitness score only) needs the grids but no covariances
  int rc = c.mode == 2 ? cloudset_build_grid(h, c.src) : cloudset_prepare(h, c.src);
  if (rc) return rc;
  if (c.tgt != c.src) {
    rc = c.mode == 2 ? cloudset_build_grid(h, c.tgt) : cloudset_prepare(h, c.tgt);
    if (rc) return rc;
  }
  const int np = c.n_pairs;
  TeamPlan plan;
  rc = plan_teams(h, c.src, c.tgt, np, c.plan_for_pairs, &plan);
  if (rc) return rc;
  rc = ensure_align_scratch(h, plan.kind == TEAM_GRID ? 1 : plan.teams, std::max(c.src->max_n, 1));
  if (rc) return rc;
  unsigned long long* d_counters = nullptr;
  if (c.block) {
    CK(cudaMemsetAsync(c.block, 0, kOB_trace, h->stream));
    d_counters = reinterpret_cast<unsigned long long*>(c.block + kOB_counters);
  } else {
    CK(h->results.reserve(sizeof(apd_result) * np));
    CK(h->counters.reserve(sizeof(unsigned long long) * 4));
    CK(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * 4, h->stream));
    d_counters = h->counters.as<unsigned long long>();
  }
  AlignBatch b;
  memset(&b, 0, sizeof(b));
  b.src = c.src->view();
  b.tgt = c.tgt->view();
  b.src_base = c.src_base;
  b.tgt_base = c.tgt_base;
  if (c.src_idx) {
    CK(h->idx_src.reserve(sizeof(int) * np));
    CK(cudaMemcpyAsync(h->idx_src.p, c.src_idx, sizeof(int) * np, cudaMemcpyHostToDevice, h->stream));
    b.src_idx = h->idx_src.as<int>();
  }
  if (c.tgt_idx) {
    CK(h->idx_tgt.reserve(sizeof(int) * np));
    CK(cudaMemcpyAsync(h->idx_tgt.p, c.tgt_idx, sizeof(int) * np, cudaMemcpyHostToDevice, h->stream));
    b.tgt_idx = h->idx_tgt.as<int>();
  }
  if (c.guesses64) {
    CK(h->guesses.reserve(sizeof(double) * 16 * np));
    CK(cudaMemcpyAsync(h->guesses.p, c.guesses64, sizeof(double) * 16 * np, cudaMemcpyHostToDevice, h->stream));
    b.guesses64 = h->guesses.as<double>();
  } else if (c.guesses) {
    CK(h->guesses.reserve(sizeof(float) * 16 * np));
    if (h->kernel_uploads) {  // the copy engine is busy with the call's points (pipelined_align): through the pinned staging block and a kernel
      const float* g = c.guesses;
      const size_t gb = sizeof(float) * 16 * np;
      int rcg = stage_upload(h, 1, gb, [&](unsigned char* host) { memcpy(host, g, gb); }, h->guesses.p);
      if (rcg) return rcg;
    } else {
      CK(cudaMemcpyAsync(h->guesses.p, c.guesses, sizeof(float) * 16 * np, cudaMemcpyHostToDevice, h->stream));
    }
    b.guesses = h->guesses.as<float>();
  }
  if (c.block) {
    b.out = reinterpret_cast<apd_result*>(c.block + kOB_result);
    b.final_hessian = reinterpret_cast<double*>(c.block + kOB_fh);
    b.lin_b = reinterpret_cast<double*>(c.block + kOB_linb);
    b.trace_rows = c.block_trace_rows;
    b.trace = reinterpret_cast<double*>(c.block + kOB_trace);
    b.trace_count = reinterpret_cast<int*>(c.block + kOB_tcount);
  } else {
    b.out = h->results.as<apd_result>();
    if (c.want_hessian || c.mode == 1) {
      CK(h->fh.reserve(sizeof(double) * 36 * np));
      CK(h->lin_b.reserve(sizeof(double) * 6 * np));
      CK(cudaMemsetAsync(h->fh.p, 0, sizeof(double) * 36 * np, h->stream));
      b.final_hessian = h->fh.as<double>();
      b.lin_b = h->lin_b.as<double>();
    }
    if (c.want_trace) {
      b.trace_rows = std::max(1, std::min(4096, h->prm.max_iterations * std::max(1, h->prm.lm_max_iterations)));
      CK(h->trace.reserve(sizeof(double) * 8 * b.trace_rows * np));
      CK(h->trace_count.reserve(sizeof(int) * np));
      CK(cudaMemsetAsync(h->trace_count.p, 0, sizeof(int) * np, h->stream));
      b.trace = h->trace.as<double>();
      b.trace_count = h->trace_count.as<int>();
    }
  }
  b.n_pairs = np;
  b.work_counter = reinterpret_cast<int*>(d_counters + 2);
  b.counters = d_counters;
  if (plan.kind == TEAM_GRID) {
    CK(h->grid_partials.reserve(sizeof(double) * 2 * plan.size * kNRed));
    b.grid_partials = h->grid_partials.as<double>();
  }
  b.scratch.max_src = h->scratch_max_src;
  b.scratch.corr = h->sc_corr.as<int>();
  b.scratch.sqd = h->sc_sqd.as<float>();
  b.scratch.m0 = h->sc_m0.as<double2>();
  b.scratch.m1 = h->sc_m1.as<double2>();
  b.scratch.m2 = h->sc_m2.as<double2>();
  b.scratch.anchor = h->sc_anchor.as<float4>();
  b.scratch.fit = h->sc_fit.as<float>();
  b.prm = device_params(h->prm);
  if (!h->search_counters.p) {
    CK(h->search_counters.reserve(sizeof(unsigned long long) * 4));
    CK(cudaMemsetAsync(h->search_counters.p, 0, sizeof(unsigned long long) * 4, h->stream));
  }
  b.nn1_evals = h->search_counters.as<unsigned long long>() + 1;
  if (h->timeline_opt) {
    CK(h->timeline.reserve(sizeof(unsigned long long) * 1024));
    CK(cudaMemsetAsync(h->timeline.p, 0, sizeof(unsigned long long) * 1024, h->stream));
    b.timeline = h->timeline.as<unsigned long long>();
  }
  b.mode = c.mode;
  b.min_points = c.min_points;
  b.max_range = c.max_range;
  {
    KernelTimer kt(h, 3);
    CK(launch_align(b, plan.kind, plan.size, plan.teams, plan.staged, plan.smem, h->stream, &h->stats));
  }
  if (used) *used = b;
  if (used_plan) *used_plan = plan;
  return APD_OK;
}

int fetch_counters(apd_handle h, int np) {
  unsigned long long c[4];
  CK(cudaMemcpyAsync(c, h->counters.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->work_lin = (long long)c[0];
  h->work_err = (long long)c[1];
  h->work_pairs = np;
  return APD_OK;
}

int set_cloud(apd_handle h, bool is_source, const float* xyz, int stride_bytes, int n, uint64_t key, int mem) {
  if (n < 0) return fail(h, APD_ERR_INVALID, "negative point count");
  auto& slot = is_source ? h->src : h->tgt;
  auto& slot_key = is_source ? h->src_key : h->tgt_key;
  auto& other = is_source ? h->tgt : h->src;
  auto& other_key = is_source ? h->tgt_key : h->src_key;
  h->last_lin_valid = h->fit_valid = false;
  if (key != 0 && slot && slot_key == key && slot->total == n) return APD_OK;  // same pointer: fast_apdgicp_impl.hpp:91,102
  if (key != 0 && other && other_key == key && other->total == n) {
    slot = other;  // the other slot already holds this cloud: share its grid and covariances
    slot_key = key;
    return APD_OK;
  }
  const int32_t off[2] = {0, n};
  std::shared_ptr<apd_cloudset_s> cs;
  int rc = make_cloudset(h, xyz, stride_bytes, off, 1, mem, &cs);
  if (rc) return rc;
  slot = cs;
  slot_key = key;
  h->has_last = false;
  return APD_OK;
}

}  // namespace

extern "C" {

int apd_abi_version(void) { return APDGICP_B200_ABI_VERSION; }

int apd_default_params(apd_params* p) {
  if (!p) return APD_ERR_INVALID;
  p->k_correspondences = 20;
  p->regularization = APD_REG_PLANE;
  p->max_iterations = 64;
  p->optimizer = APD_OPT_LEVENBERG_MARQUARDT;
  p->lm_max_iterations = 10;
  p->num_threads = 0;
  p->max_corr_dist = (double)FLT_MAX;
  p->rotation_epsilon = 2e-3;
  p->transformation_epsilon = 5e-4;
  p->lm_init_lambda_factor = 1e-9;
  p->dist_var = 0.86;
  p->azimuth_var = 0.5;
  p->elevation_var = 1.0;
  return APD_OK;
}

int apd_create(int device_id, apd_handle* out) {
  if (!out) return APD_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    g_err = "no CUDA device: apdgicp_b200 has no CPU path";
    return APD_ERR_NO_DEVICE;
  }
  if (device_id < 0 || device_id >= count) {
    g_err = "device id out of range";
    return APD_ERR_INVALID;
  }
  apd_handle h = new (std::nothrow) apd_context();
  if (!h) return APD_ERR_INVALID;
  h->device = device_id;
  DeviceGuard guard(device_id);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess || prop.major < 10) {
    g_err = "apdgicp_b200 is built for sm_100a (Blackwell B200) only";
    delete h;
    return APD_ERR_NO_DEVICE;
  }
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_err = "cudaStreamCreate failed";
    delete h;
    return APD_ERR_CUDA;
  }
  h->stream = h->own_stream;
  apd_default_params(&h->prm);
  memset(&h->last, 0, sizeof(h->last));
  for (int i = 0; i < 36; i++) h->final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;  // final_hessian_.setIdentity(), lsq_registration_impl.hpp:23
  *out = h;
  return APD_OK;
}

int apd_destroy(apd_handle h) {
  if (!h) return APD_OK;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  h->src.reset();
  h->tgt.reset();
  if (h->helper) apd_destroy(h->helper);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (auto& e : h->copy_events) cudaEventDestroy(e);
  for (int i = 0; i < 2; i++) {
    if (h->stage_host[i]) cudaFreeHost(h->stage_host[i]);
    if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
  }
  if (h->down_host) cudaFreeHost(h->down_host);
  for (auto& t : h->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  for (auto& e : h->event_pool) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return APD_OK;
}

const char* apd_last_error(apd_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int apd_set_stream(apd_handle h, void* cuda_stream) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return APD_OK;
}

int apd_set_params(apd_handle h, const apd_params* p) {
  if (!h || !p) return APD_ERR_INVALID;
  if (p->k_correspondences < 1) return fail(h, APD_ERR_INVALID, "k_correspondences < 1");
  if (p->k_correspondences > 32) return fail(h, APD_ERR_UNSUPPORTED, "k_correspondences > 32 is not supported");
  if (p->regularization < 0 || p->regularization > 4) return fail(h, APD_ERR_INVALID, "unknown regularization method");
  if (p->optimizer < 0 || p->optimizer > 1) return fail(h, APD_ERR_INVALID, "unknown optimizer");
  h->prm = *p;
  h->last_lin_valid = h->fit_valid = false;
  return APD_OK;
}

int apd_get_params(apd_handle h, apd_params* p) {
  if (!h || !p) return APD_ERR_INVALID;
  *p = h->prm;
  return APD_OK;
}

int apd_set_option(apd_handle h, const char* name, double value) {
  if (!h || !name) return APD_ERR_INVALID;
  const std::string n(name);
  if (n == "cells_per_point") { if (!(value > 0)) return fail(h, APD_ERR_INVALID, "cells_per_point must be positive"); h->cells_per_point = value; }
  else if (n == "team_size") h->team_size = (int)value;
  else if (n == "force_unstaged") h->force_unstaged = value != 0;
  else if (n == "max_teams") h->max_teams_opt = (int)value;
  else if (n == "knn_packed") h->knn_packed = value != 0;
  else if (n == "fused_build") h->no_fused_build = value == 0;
  else if (n == "smem_build") h->no_smem_build = value == 0;
  else if (n == "fitness_max_range") h->fitness_max_range = value;
  else if (n == "knn_fine_rings") h->knn_fine_rings = std::max(0, (int)value);
  else if (n == "timeline") h->timeline_opt = value != 0;
  else if (n == "knn_leaf_parts") {
    const int v = (int)value;
    if (v != 0 && v != 1 && v != 2 && v != 4 && v != 8) return fail(h, APD_ERR_INVALID, "knn_leaf_parts must be 0, 1, 2, 4 or 8");
    h->knn_leaf_parts = v;
  }
  else if (n == "kernel_timing") h->kernel_timing = value != 0;
  else if (n == "upload_ahead") h->upload_ahead = value != 0;
  else if (n == "pipeline_chunks") h->pipe_chunks = std::max(1, (int)value);
  else if (n == "pipeline_first_waves") h->pipe_first_waves = std::max(1, (int)value);
  else if (n == "pipeline_device_input") h->pipeline_mem = value != 0 ? APD_MEM_DEVICE : APD_MEM_HOST;
  else if (n == "bulk_stage") h->bulk_stage = value != 0;   // takes effect for cloud sets made afterwards
  else if (n == "downsample_method") {   // the rosparam of preprocessing_nodelet.cpp:137 / scan_matching_odometry_nodelet.cpp:149 (NONE = downsample_resolution <= 0)
    if (value != 0 && value != 1) return fail(h, APD_ERR_UNSUPPORTED, "downsample_method must be 0 (VOXELGRID) or 1 (APPROX_VOXELGRID)");
    h->downsample_method = (int)value;
  }
  else return fail(h, APD_ERR_INVALID, "unknown option " + n);
  return APD_OK;
}

int apd_set_source(apd_handle h, const float* xyz, int stride_bytes, int n, uint64_t cache_key, int mem) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return set_cloud(h, true, xyz, stride_bytes, n, cache_key, mem);
}

int apd_set_target(apd_handle h, const float* xyz, int stride_bytes, int n, uint64_t cache_key, int mem) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return set_cloud(h, false, xyz, stride_bytes, n, cache_key, mem);
}

int apd_swap_source_and_target(apd_handle h) {
  if (!h) return APD_ERR_INVALID;
  std::swap(h->src, h->tgt);
  std::swap(h->src_key, h->tgt_key);
  h->last_lin_valid = h->fit_valid = false;  // correspondences_.clear(), fast_apdgicp_impl.hpp:73
  return APD_OK;
}

int apd_clear_source(apd_handle h) {
  if (!h) return APD_ERR_INVALID;
  h->src.reset();
  h->src_key = 0;
  h->last_lin_valid = h->fit_valid = false;
  return APD_OK;
}

int apd_clear_target(apd_handle h) {
  if (!h) return APD_ERR_INVALID;
  h->tgt.reset();
  h->tgt_key = 0;
  h->last_lin_valid = h->fit_valid = false;
  return APD_OK;
}

static int single_pair_checks(apd_handle h) {
  if (!h->src || !h->tgt || h->src->total == 0 || h->tgt->total == 0) return fail(h, APD_ERR_NO_INPUT, "source or target cloud not set");
  const int k = h->prm.k_correspondences;
  const bool src_needs = !(h->src->cov_valid && h->src->cov_k < 0);
  const bool tgt_needs = !(h->tgt->cov_valid && h->tgt->cov_k < 0);
  if ((src_needs && h->src->total < k) || (tgt_needs && h->tgt->total < k)) return fail(h, APD_ERR_TOO_FEW_POINTS, "cloud has fewer points than k_correspondences");
  return APD_OK;
}

int apd_compute_covariances(apd_handle h) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  const int k = h->prm.k_correspondences;
  for (auto* cs : {h->src.get(), h->tgt.get()}) {
    if (!cs || cs->total == 0) continue;
    if (!(cs->cov_valid && cs->cov_k < 0) && cs->total < k) return fail(h, APD_ERR_TOO_FEW_POINTS, "cloud has fewer points than k_correspondences");
    int rc = cloudset_prepare(h, cs);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_align(apd_handle h, const float guess[16], apd_result* out) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  // pcl::Registration::align resets converged_ and final_transformation_ before anything else
  memset(&h->last, 0, sizeof(h->last));
  for (int i = 0; i < 16; i++) h->last.T[i] = (i % 5 == 0) ? 1.f : 0.f;
  h->last.fitness = DBL_MAX;
  h->has_last = false;
  h->lm_trace.clear();
  int rc = single_pair_checks(h);
  if (rc) {
    h->last.status = rc;
    if (out) *out = h->last;
    return rc;
  }
  AlignCall c;
  c.src = h->src.get();
  c.tgt = h->tgt.get();
  c.guesses = guess;
  c.n_pairs = 1;
  c.want_trace = true;
  c.want_hessian = true;
  // One device block, one memset, one copy into pinned memory, one synchronisation. (Five separate copies into pageable host
  // variables - record, trace count, Hessian, counters, trace - each blocked the host for a driver round trip: ~40 us of a
  // 165 us call around a 90 us kernel.)
  rc = single_block(h, &c);
  if (rc) return rc;
  const size_t fast_bytes = kOB_trace + sizeof(double) * 8 * (size_t)std::min(c.block_trace_rows, kOB_fast_rows);
  AlignBatch b;
  rc = run_align(h, c, &b);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->down_host, c.block, fast_bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(&h->last, h->down_host + kOB_result, sizeof(apd_result));
  {
    unsigned long long cnt[4];
    memcpy(cnt, h->down_host + kOB_counters, sizeof(cnt));
    h->work_lin = (long long)cnt[0];
    h->work_err = (long long)cnt[1];
    h->work_pairs = 1;
  }
  int n_rows = 0;
  memcpy(&n_rows, h->down_host + kOB_tcount, sizeof(int));
  double fh[36];
  memcpy(fh, h->down_host + kOB_fh, sizeof(fh));
  if (n_rows > 0) {
    h->lm_trace.resize((size_t)n_rows * 8);
    const int fast = std::min(n_rows, kOB_fast_rows);
    memcpy(h->lm_trace.data(), h->down_host + kOB_trace, sizeof(double) * 8 * fast);
    if (n_rows > fast)
      CK(cudaMemcpy(h->lm_trace.data() + (size_t)fast * 8, c.block + kOB_trace + sizeof(double) * 8 * fast, sizeof(double) * 8 * (n_rows - fast), cudaMemcpyDeviceToHost));
  }
  bool any = false;
  for (int i = 0; i < 36; i++) any |= fh[i] != 0.0;
  if (any) memcpy(h->final_hessian, fh, sizeof(fh));  // only an accepted step writes final_hessian_
  h->has_last = true;
  h->last_lin_valid = true;
  h->fit_valid = h->last.status == APD_OK || h->last.status == APD_STATUS_LM_FAILED;  // the fitness pass ran at the final pose
  if (out) *out = h->last;
  return APD_OK;
}

int apd_fitness(apd_handle h, double max_range, double* score) {
  if (!h || !score) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (!h->src || !h->tgt || h->src->total == 0 || h->tgt->total == 0) return fail(h, APD_ERR_NO_INPUT, "source or target cloud not set");
  int rc = cloudset_build_grid(h, h->src.get());
  if (rc) return rc;
  rc = cloudset_build_grid(h, h->tgt.get());
  if (rc) return rc;
  float T[16];
  if (h->has_last) memcpy(T, h->last.T, sizeof(T));
  else for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.f : 0.f;
  const int blocks = std::max(1, std::min(2 * h->sm_count, (int)((h->src->total + 255) / 256)));
  CK(h->misc.reserve(sizeof(double) * (2 * blocks + 2) + sizeof(float) * 16));
  double* partials = h->misc.as<double>();
  double* res = partials + 2 * blocks;
  float* dT = reinterpret_cast<float*>(res + 2);
  CK(cudaMemcpyAsync(dT, T, sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CK(launch_fitness(h->src->view(), 0, h->tgt->view(), 0, dT, max_range, false, partials, blocks, res, h->stream, &h->stats));
  double r[2];
  CK(cudaMemcpyAsync(r, res, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *score = r[0];
  return APD_OK;
}

// InformationMatrixCalculator::calc_fitness_score (radar_graph_slam/src/radar_graph_slam/information_matrix_calculator.cpp:55-86):
// target slot = cloud1 (the kd-tree side), source slot = cloud2, T = relpose.cast<float>()
int apd_fitness_score(apd_handle h, const float T[16], double max_range, double* score, int64_t* n_used) {
  if (!h || !score) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (!h->src || !h->tgt || h->src->total == 0 || h->tgt->total == 0) return fail(h, APD_ERR_NO_INPUT, "source or target cloud not set");
  int rc = cloudset_build_grid(h, h->src.get());
  if (rc) return rc;
  rc = cloudset_build_grid(h, h->tgt.get());
  if (rc) return rc;
  float Th[16];
  if (T) memcpy(Th, T, sizeof(Th));
  else for (int i = 0; i < 16; i++) Th[i] = (i % 5 == 0) ? 1.f : 0.f;
  const int blocks = std::max(1, std::min(2 * h->sm_count, (int)((h->src->total + 255) / 256)));
  CK(h->misc.reserve(sizeof(double) * (2 * blocks + 2) + sizeof(float) * 16));
  double* partials = h->misc.as<double>();
  double* res = partials + 2 * blocks;
  float* dT = reinterpret_cast<float*>(res + 2);
  CK(cudaMemcpyAsync(dT, Th, sizeof(Th), cudaMemcpyHostToDevice, h->stream));
  CK(launch_fitness(h->src->view(), 0, h->tgt->view(), 0, dT, max_range, false, partials, blocks, res, h->stream, &h->stats));
  double r[2];
  CK(cudaMemcpyAsync(r, res, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *score = r[0];
  if (n_used) *n_used = (int64_t)r[1];
  return APD_OK;
}

// publish_scan_matching_status (radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:698-712): for every point of the aligned cloud
// (the source moved by the final transformation) one 1-NN query in the target, counted when k_sq_dists[0] < max_dist * max_dist.
int apd_inlier_count(apd_handle h, const float T[16], double max_dist, int64_t* n_inliers) {
  if (!h || !n_inliers) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (!h->src || !h->tgt || h->src->total == 0 || h->tgt->total == 0) return fail(h, APD_ERR_NO_INPUT, "source or target cloud not set");
  const double thr = max_dist * max_dist;
  double r[2] = {0, 0};
  if (!T && h->fit_valid) {
    // the align kernel left the squared 1-NN distance of every source point at the final pose in its scratch: count, no search
    CK(h->misc.reserve(sizeof(double) * 2));
    CK(launch_count_below(h->sc_fit.as<float>(), (int)h->src->total, thr, h->misc.as<double>(), h->stream, &h->stats));
    CK(cudaMemcpyAsync(r, h->misc.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *n_inliers = (int64_t)r[0];
    return APD_OK;
  }
  int rc = cloudset_build_grid(h, h->src.get());
  if (rc) return rc;
  rc = cloudset_build_grid(h, h->tgt.get());
  if (rc) return rc;
  float Th[16];
  if (T) memcpy(Th, T, sizeof(Th));
  else if (h->has_last) memcpy(Th, h->last.T, sizeof(Th));
  else for (int i = 0; i < 16; i++) Th[i] = (i % 5 == 0) ? 1.f : 0.f;
  const int blocks = std::max(1, std::min(2 * h->sm_count, (int)((h->src->total + 255) / 256)));
  CK(h->misc.reserve(sizeof(double) * (2 * blocks + 2) + sizeof(float) * 16));
  double* partials = h->misc.as<double>();
  double* res = partials + 2 * blocks;
  float* dT = reinterpret_cast<float*>(res + 2);
  CK(cudaMemcpyAsync(dT, Th, sizeof(Th), cudaMemcpyHostToDevice, h->stream));
  CK(launch_fitness(h->src->view(), 0, h->tgt->view(), 0, dT, thr, true, partials, blocks, res, h->stream, &h->stats));
  CK(cudaMemcpyAsync(r, res, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *n_inliers = (int64_t)r[1];
  return APD_OK;
}

// Batched calc_fitness_score: pair i scores cloud src_idx[i] of `src` (cloud2) against cloud tgt_idx[i] of `tgt`
// (cloud1) at poses[i]; one launch for a whole sliding window of keyframe pairs.
int apd_fitness_pairs(apd_handle h, apd_cloudset src, apd_cloudset tgt, const int32_t* src_idx, const int32_t* tgt_idx, const float* poses, int n_pairs,
                      double max_range, double* scores) {
  if (!h || !src || !tgt || !scores || n_pairs < 0) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (n_pairs == 0) return APD_OK;
  apd_cloudset_s* s = reinterpret_cast<std::shared_ptr<apd_cloudset_s>*>(src)->get();
  apd_cloudset_s* t = reinterpret_cast<std::shared_ptr<apd_cloudset_s>*>(tgt)->get();
  for (int i = 0; i < n_pairs; i++) {
    const int si = src_idx ? src_idx[i] : i, ti = tgt_idx ? tgt_idx[i] : i;
    if (si < 0 || si >= s->n_clouds || ti < 0 || ti >= t->n_clouds) return fail(h, APD_ERR_INVALID, "pair index out of range");
  }
  AlignCall c;
  c.src = s;
  c.tgt = t;
  c.src_idx = src_idx;
  c.tgt_idx = tgt_idx;
  c.guesses = poses;
  c.n_pairs = n_pairs;
  c.mode = 2;
  c.max_range = max_range;
  int rc = run_align(h, c);
  if (rc) return rc;
  std::vector<apd_result> r(n_pairs);
  CK(cudaMemcpyAsync(r.data(), h->results.p, sizeof(apd_result) * n_pairs, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n_pairs; i++) scores[i] = r[i].fitness;
  h->last_lin_valid = h->fit_valid = false;
  return APD_OK;
}

int apd_transform_source(apd_handle h, const float T[16], float* out_xyz, int out_stride_bytes, int mem) {
  if (!h || !out_xyz) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (!h->src || h->src->total == 0) return fail(h, APD_ERR_NO_INPUT, "source cloud not set");
  if (out_stride_bytes < 12 || out_stride_bytes % 4) return fail(h, APD_ERR_INVALID, "bad output stride");
  const int n = (int)h->src->total;
  float Th[16];
  if (T) memcpy(Th, T, sizeof(Th));
  else if (h->has_last) memcpy(Th, h->last.T, sizeof(Th));
  else for (int i = 0; i < 16; i++) Th[i] = (i % 5 == 0) ? 1.f : 0.f;
  CK(h->misc.reserve(sizeof(float) * 16 + 64));
  float* dT = h->misc.as<float>();
  CK(cudaMemcpyAsync(dT, Th, sizeof(Th), cudaMemcpyHostToDevice, h->stream));
  if (mem == APD_MEM_DEVICE) {
    CK(launch_transform_points(h->src->pts.as<float4>(), n, dT, out_xyz, out_stride_bytes / 4, h->stream, &h->stats));
    CK(cudaStreamSynchronize(h->stream));
    return APD_OK;
  }
  CK(h->cov_tmp.reserve(sizeof(float) * 3 * (size_t)n));
  CK(launch_transform_points(h->src->pts.as<float4>(), n, dT, h->cov_tmp.as<float>(), 3, h->stream, &h->stats));
  if (out_stride_bytes == 12) {
    CK(cudaMemcpyAsync(out_xyz, h->cov_tmp.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  } else {
    CK(cudaMemcpy2DAsync(out_xyz, out_stride_bytes, h->cov_tmp.p, 12, 12, n, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

static int linearize_common(apd_handle h, const float* pose_f, const double* pose_d, double H[36], double b[6], double* error);
int apd_linearize(apd_handle h, const float pose[16], double H[36], double b[6], double* error) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return linearize_common(h, pose, nullptr, H, b, error);
}

// the protected hooks of the reference class, callable at a DOUBLE pose (they take an Eigen::Isometry3d):
// linearize (fast_apdgicp_impl.hpp:198-272) and compute_error (:275-298)
static int linearize_common(apd_handle h, const float* pose_f, const double* pose_d, double H[36], double b[6], double* error) {
  int rc = single_pair_checks(h);
  if (rc) return rc;
  AlignCall c;
  c.src = h->src.get();
  c.tgt = h->tgt.get();
  c.guesses = pose_f;
  c.guesses64 = pose_d;
  c.n_pairs = 1;
  c.mode = 1;
  // record, H and b in one device block, fetched by one copy into pinned memory (as apd_align)
  rc = single_block(h, &c);
  if (rc) return rc;
  rc = run_align(h, c);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->down_host, c.block, kOB_trace, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  apd_result r;
  memcpy(&r, h->down_host + kOB_result, sizeof(r));
  if (H) memcpy(H, h->down_host + kOB_fh, sizeof(double) * 36);
  if (b) memcpy(b, h->down_host + kOB_linb, sizeof(double) * 6);
  if (error) *error = r.error;
  h->last_lin_valid = true;
  h->fit_valid = false;
  return APD_OK;
}

int apd_linearize_d(apd_handle h, const double pose[16], double H[36], double b[6], double* error) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return linearize_common(h, nullptr, pose, H, b, error);
}

int apd_compute_error(apd_handle h, const double pose[16], double* error) {
  if (!h || !error) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (!h->last_lin_valid || !h->src || !h->tgt) return fail(h, APD_ERR_NO_INPUT, "compute_error needs the correspondences of a previous linearize on the current clouds");
  AlignCall c;
  c.src = h->src.get();
  c.tgt = h->tgt.get();
  c.guesses64 = pose;
  c.n_pairs = 1;
  c.mode = 3;
  int rc = run_align(h, c);
  if (rc) return rc;
  apd_result r;
  CK(cudaMemcpyAsync(&r, h->results.p, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *error = r.error;
  return APD_OK;
}

int apd_get_final_hessian(apd_handle h, double H[36]) {
  if (!h || !H) return APD_ERR_INVALID;
  memcpy(H, h->final_hessian, sizeof(h->final_hessian));
  return APD_OK;
}

int apd_get_knn(apd_handle h, int which, int32_t* idx_out) {
  if (!h || !idx_out) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  apd_cloudset_s* cs = which ? h->tgt.get() : h->src.get();
  if (!cs || cs->total == 0) return fail(h, APD_ERR_NO_INPUT, "cloud not set");
  const int k = h->prm.k_correspondences;
  if (cs->total < k) return fail(h, APD_ERR_TOO_FEW_POINTS, "cloud has fewer points than k_correspondences");
  CK(h->knn_tmp.reserve(sizeof(int) * (size_t)cs->total * k));
  const bool injected = cs->cov_valid && cs->cov_k < 0;
  if (injected) return fail(h, APD_ERR_UNSUPPORTED, "covariances were injected by the caller; kNN sets are not available");
  int rc = cloudset_prepare(h, cs, h->knn_tmp.as<int>());
  if (rc) return rc;
  CK(cudaMemcpyAsync(idx_out, h->knn_tmp.p, sizeof(int) * (size_t)cs->total * k, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_get_covariances(apd_handle h, int which, double* c16_out) {
  if (!h || !c16_out) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  apd_cloudset_s* cs = which ? h->tgt.get() : h->src.get();
  if (!cs || cs->total == 0) return fail(h, APD_ERR_NO_INPUT, "cloud not set");
  if (!(cs->cov_valid && cs->cov_k < 0) && cs->total < h->prm.k_correspondences)
    return fail(h, APD_ERR_TOO_FEW_POINTS, "cloud has fewer points than k_correspondences");
  int rc = cloudset_prepare(h, cs);
  if (rc) return rc;
  CK(h->cov_tmp.reserve(sizeof(double) * 16 * (size_t)cs->total));
  CK(launch_cov_export(cs->view(), 0, h->cov_tmp.as<double>(), h->stream, &h->stats));
  CK(cudaMemcpyAsync(c16_out, h->cov_tmp.p, sizeof(double) * 16 * (size_t)cs->total, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_set_covariances(apd_handle h, int which, const double* c16, int n) {
  if (!h || !c16) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  apd_cloudset_s* cs = which ? h->tgt.get() : h->src.get();
  if (!cs || cs->total == 0) return fail(h, APD_ERR_NO_INPUT, "cloud not set");
  if (n != cs->total) return fail(h, APD_ERR_INVALID, "covariance count does not match the cloud size");
  if (h->src && h->src == h->tgt) {
    // Both slots share one device cloud (same cache key: the previous source became the target). The reference keeps
    // source_covs_ and target_covs_ apart (fast_apdgicp_impl.hpp:111-118), so the slot being written gets its own copy
    // first: same points (device-to-device), its own grid, then the injected covariances.
    auto& slot = which ? h->tgt : h->src;
    const int32_t off[2] = {0, n};
    std::shared_ptr<apd_cloudset_s> own;
    int rc0 = make_cloudset(h, reinterpret_cast<const float*>(cs->pts.p), 16, off, 1, APD_MEM_DEVICE, &own);
    if (rc0) return rc0;
    slot = own;
    cs = own.get();
  }
  int rc = cloudset_build_grid(h, cs);
  if (rc) return rc;
  CK(h->cov_tmp.reserve(sizeof(double) * 16 * (size_t)n));
  CK(cudaMemcpyAsync(h->cov_tmp.p, c16, sizeof(double) * 16 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  CK(launch_cov_import(cs->view(), 0, h->cov_tmp.as<double>(), h->stream, &h->stats));
  CK(cudaStreamSynchronize(h->stream));
  cs->cov_valid = true;
  cs->cov_k = -1;  // marks "provided by the caller": never recomputed
  cs->cov_reg = -1;
  h->last_lin_valid = h->fit_valid = false;
  return APD_OK;
}

static int export_corr(apd_handle h, int32_t* corr_out, float* sq_dist_out, double* m16_out) {
  if (!h->last_lin_valid || !h->src || !h->tgt) return fail(h, APD_ERR_NO_INPUT, "no linearization has run on the current clouds");
  const int n = (int)h->src->total;
  CK(h->cov_tmp.reserve(sizeof(double) * 16 * (size_t)n));
  CK(h->knn_tmp.reserve(sizeof(int) * 2 * (size_t)n));
  AlignBatch b;
  memset(&b, 0, sizeof(b));
  b.src = h->src->view();
  b.tgt = h->tgt->view();
  b.scratch.max_src = h->scratch_max_src;
  b.scratch.corr = h->sc_corr.as<int>();
  b.scratch.sqd = h->sc_sqd.as<float>();
  b.scratch.m0 = h->sc_m0.as<double2>();
  b.scratch.m1 = h->sc_m1.as<double2>();
  b.scratch.m2 = h->sc_m2.as<double2>();
  int* dcorr = h->knn_tmp.as<int>();
  float* dsqd = reinterpret_cast<float*>(dcorr + n);
  CK(launch_corr_export(b, 0, 0, 0, dcorr, dsqd, m16_out ? h->cov_tmp.as<double>() : nullptr, h->stream, &h->stats));
  if (corr_out) CK(cudaMemcpyAsync(corr_out, dcorr, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  if (sq_dist_out) CK(cudaMemcpyAsync(sq_dist_out, dsqd, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
  if (m16_out) CK(cudaMemcpyAsync(m16_out, h->cov_tmp.p, sizeof(double) * 16 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_get_correspondences(apd_handle h, int32_t* corr_out, float* sq_dist_out) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return export_corr(h, corr_out, sq_dist_out, nullptr);
}

int apd_get_mahalanobis(apd_handle h, double* m16_out) {
  if (!h || !m16_out) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return export_corr(h, nullptr, nullptr, m16_out);
}

int apd_get_lm_trace(apd_handle h, double* rows8, int max_rows, int* n_rows) {
  if (!h) return APD_ERR_INVALID;
  const int n = (int)(h->lm_trace.size() / 8);
  if (n_rows) *n_rows = n;
  if (rows8 && max_rows > 0) memcpy(rows8, h->lm_trace.data(), sizeof(double) * 8 * std::min(n, max_rows));
  return APD_OK;
}

// ---- batched path ----

int apd_cloudset_create(apd_handle h, const float* xyz, int stride_bytes, const int32_t* offsets, int n_clouds, int mem, apd_cloudset* out) {
  if (!h || !out) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  std::shared_ptr<apd_cloudset_s> cs;
  int rc = make_cloudset(h, xyz, stride_bytes, offsets, n_clouds, mem, &cs);
  if (rc) return rc;
  // hand the caller a heap-allocated shared_ptr so the set's lifetime is explicit
  *out = reinterpret_cast<apd_cloudset>(new std::shared_ptr<apd_cloudset_s>(cs));
  return APD_OK;
}

static apd_cloudset_s* deref(apd_cloudset cs) { return cs ? reinterpret_cast<std::shared_ptr<apd_cloudset_s>*>(cs)->get() : nullptr; }

int apd_cloudset_destroy(apd_handle h, apd_cloudset cs) {
  if (!h) return APD_ERR_INVALID;
  if (!cs) return APD_OK;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  delete reinterpret_cast<std::shared_ptr<apd_cloudset_s>*>(cs);
  return APD_OK;
}

int apd_cloudset_prepare(apd_handle h, apd_cloudset cs) {
  if (!h || !cs) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  return cloudset_prepare(h, deref(cs));
}

int apd_align_pairs(apd_handle h, apd_cloudset src, apd_cloudset tgt, const int32_t* src_idx, const int32_t* tgt_idx, const float* guesses, int n_pairs,
                    apd_result* out, int out_mem) {
  if (!h || !src || !tgt || !out || n_pairs < 0) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (n_pairs == 0) return APD_OK;
  apd_cloudset_s* s = deref(src);
  apd_cloudset_s* t = deref(tgt);
  for (int i = 0; i < n_pairs; i++) {
    const int si = src_idx ? src_idx[i] : i, ti = tgt_idx ? tgt_idx[i] : i;
    if (si < 0 || si >= s->n_clouds || ti < 0 || ti >= t->n_clouds) return fail(h, APD_ERR_INVALID, "pair index out of range");
  }
  AlignCall c;
  c.src = s;
  c.tgt = t;
  c.src_idx = src_idx;
  c.tgt_idx = tgt_idx;
  c.guesses = guesses;
  c.n_pairs = n_pairs;
  c.min_points = h->prm.k_correspondences;
  c.max_range = h->fitness_max_range;
  int rc = run_align(h, c);
  if (rc) return rc;
  h->last_lin_valid = h->fit_valid = false;
  if (out_mem == APD_MEM_DEVICE) {
    CK(cudaMemcpyAsync(out, h->results.p, sizeof(apd_result) * n_pairs, cudaMemcpyDeviceToDevice, h->stream));
    h->work_pairs = n_pairs;
    return APD_OK;
  }
  CK(cudaMemcpyAsync(out, h->results.p, sizeof(apd_result) * n_pairs, cudaMemcpyDeviceToHost, h->stream));
  return fetch_counters(h, n_pairs);
}

// LoopDetector::matching over ALL candidates (radar_graph_slam/src/radar_graph_slam/loop_detector.cpp:379-441, #if 0 in the reference because
// it is too slow on the CPU): registration->setInputTarget(new_keyframe->cloud); for every candidate setInputSource, align(guess),
// getFitnessScore(fitness_score_max_range); keep the best converged score (:415-423); reject above fitness_score_thresh (:431-434).
int apd_match_candidates(apd_handle h, apd_cloudset candidates, const int32_t* cand_idx, int n_candidates, apd_cloudset keyframes, int keyframe_idx, const float* guesses,
                         double fitness_score_max_range, double fitness_score_thresh, int32_t* best, float relative_pose[16], double* best_score, apd_result* records) {
  if (!h || !candidates || !keyframes || !best || n_candidates < 0) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  *best = -1;
  if (best_score) *best_score = DBL_MAX;
  if (n_candidates == 0) return APD_OK;  // :391-393
  apd_cloudset_s* s = deref(candidates);
  apd_cloudset_s* t = deref(keyframes);
  if (keyframe_idx < 0 || keyframe_idx >= t->n_clouds) return fail(h, APD_ERR_INVALID, "keyframe index out of range");
  std::vector<int32_t> ti((size_t)n_candidates, keyframe_idx);
  for (int i = 0; i < n_candidates; i++) {
    const int si = cand_idx ? cand_idx[i] : i;
    if (si < 0 || si >= s->n_clouds) return fail(h, APD_ERR_INVALID, "candidate index out of range");
  }
  AlignCall c;
  c.src = s;
  c.tgt = t;
  c.src_idx = cand_idx;
  c.tgt_idx = ti.data();
  c.guesses = guesses;
  c.n_pairs = n_candidates;
  c.min_points = h->prm.k_correspondences;
  c.max_range = fitness_score_max_range;
  int rc = run_align(h, c);
  if (rc) return rc;
  h->last_lin_valid = h->fit_valid = false;
  std::vector<apd_result> r((size_t)n_candidates);
  CK(cudaMemcpyAsync(r.data(), h->results.p, sizeof(apd_result) * n_candidates, cudaMemcpyDeviceToHost, h->stream));
  rc = fetch_counters(h, n_candidates);
  if (rc) return rc;
  double bs = DBL_MAX;
  int bi = -1;
  for (int i = 0; i < n_candidates; i++) {
    const double score = r[i].fitness;
    if (!r[i].converged || r[i].status != APD_OK || score > bs) continue;  // :415-417 (a pair the reference could not even align never converges)
    bs = score;
    bi = i;
  }
  if (records) memcpy(records, r.data(), sizeof(apd_result) * n_candidates);
  if (best_score) *best_score = bs;
  if (bi < 0 || bs > fitness_score_thresh) return APD_OK;  // "loop not found..." (:431-434)
  *best = bi;
  if (relative_pose) memcpy(relative_pose, r[bi].T, sizeof(float) * 16);
  return APD_OK;
}

// ---- pipelined host-to-host batches ----
// A large batch is cut into chunks that alternate between the handle and a lazily created helper
// handle (same device, own stream and allocation pool): while one chunk's kernels run, the next
// chunk's points cross PCIe on the other stream. Results are bitwise identical to one big launch:
// pairs are independent and every pair is processed by the same team shape.
namespace {

constexpr int kMinChunkPairs = 256;  // smallest chunk worth a launch sequence of its own

struct ChunkSlot {
  apd_handle h = nullptr;
  std::shared_ptr<apd_cloudset_s> a, b;  // kept alive until the slot's stream has drained
  int pairs = 0;
};

int helper_of(apd_handle h, apd_handle* out) {
  if (!h->helper) {
    apd_handle x = nullptr;
    int rc = apd_create(h->device, &x);
    if (rc) return fail(h, rc, "could not create the pipeline helper handle");
    h->helper = x;
  }
  apd_handle x = h->helper;
  x->prm = h->prm;
  x->cells_per_point = h->cells_per_point;
  x->team_size = h->team_size;
  x->force_unstaged = h->force_unstaged;
  x->max_teams_opt = h->max_teams_opt;
  x->knn_packed = h->knn_packed;
  x->no_fused_build = h->no_fused_build;
  x->no_smem_build = h->no_smem_build;
  x->kernel_timing = h->kernel_timing;
  x->bulk_stage = h->bulk_stage;
  x->fitness_max_range = h->fitness_max_range;
  x->knn_fine_rings = h->knn_fine_rings;
  x->knn_leaf_parts = h->knn_leaf_parts;
  *out = x;
  return APD_OK;
}

// wait for a slot's previous chunk, collect its counters, release its clouds
int retire(apd_handle owner, ChunkSlot& s, long long* lin, long long* err) {
  if (!s.h || s.pairs == 0) return APD_OK;
  int rc = fetch_counters(s.h, s.pairs);
  if (rc) return s.h == owner ? rc : fail(owner, rc, s.h->err);
  *lin += s.h->work_lin;
  *err += s.h->work_err;
  s.a.reset();
  s.b.reset();
  s.pairs = 0;
  return APD_OK;
}

// odometry == true: one ragged array of n_pairs + 1 scans, pair i = scan i+1 -> scan i
int pipelined_align(apd_handle h, const float* pts_src, const int32_t* off_src, const float* pts_tgt, const int32_t* off_tgt, int stride_bytes, const float* guesses,
                    int n_pairs, apd_result* out, bool odometry) {
  // about four chunks per call, each a whole number of waves of one-pair-per-SM teams: enough to hide
  // all but the first upload, few enough that launch tails and per-chunk host work stay small
  int kChunkPairs = std::max(kMinChunkPairs, (n_pairs + h->pipe_chunks - 1) / h->pipe_chunks);
  kChunkPairs = (kChunkPairs + h->sm_count - 1) / h->sm_count * h->sm_count;
  ChunkSlot slots[2];
  slots[0].h = h;
  if (n_pairs > kChunkPairs) {
    int rc = helper_of(h, &slots[1].h);
    if (rc) return rc;
  }
  long long lin = 0, err = 0;
  std::vector<int32_t> idx_s, idx_t;
  int chunk = 0;
  // Every early return below leaves through this guard: the other slot may still have kernels and an asynchronous copy into
  // the caller's `out` in flight on its own stream, and its cloud sets go back to the pool when the slots die.
  struct Drain {
    ChunkSlot* s;
    ~Drain() {
      for (int i = 0; i < 2; i++)
        if (s[i].h) cudaStreamSynchronize(s[i].h->stream);
    }
  } drain{slots};
  // the first chunk's upload is the one nothing can hide: keep it to a single wave of teams when the batch is large
  const int first_pairs = (slots[1].h && n_pairs > 3 * h->sm_count) ? std::min(kChunkPairs, h->pipe_first_waves * h->sm_count) : kChunkPairs;  // measured: 74 / 148 / 256 pairs -> 12.85 / 12.31 / 12.54 ms per 1000 pairs
  // A chunk of np odometry pairs holds np + 1 scans, and the kNN kernel takes whole scans per CTA, one CTA per SM: 297 scans are
  // three waves where 296 are two (a 1000-pair call ran 2 + 3 + 3 + 2 = 10 waves of kNN instead of 7, +1.4 ms). Odometry chunks
  // therefore hold one pair less than a multiple of the SM count.
  const int odo = odometry && n_pairs > kChunkPairs ? 1 : 0;
  // the chunks of this call
  std::vector<std::pair<int, int>> chunks;  // (first pair, pairs)
  // Sizes ramp up geometrically from the first wave: chunk k + 1's points must have crossed the link by the time chunk k's kernels
  // finish, and a pair uploads in about half the time it computes (C4: 320 KB at 54 GB/s = 5.9 us against 11.4 us), so a chunk may be
  // twice its predecessor but not seven times (measured on C4: 148 then 1036 pairs left the GPU idle for 3.9 ms of a 52 ms call).
  for (int p0 = 0, np = 0, cap = first_pairs; p0 < n_pairs; p0 += np, cap = std::min(kChunkPairs, 2 * cap)) {
    np = std::min(cap - odo, n_pairs - p0);
    chunks.emplace_back(p0, np);
  }
  // UPLOAD-AHEAD. With per-chunk uploads on the chunk's own stream, chunk k + 2 cannot start crossing PCIe before chunk k has been
  // retired (its slot is reused), and the measured timeline (APD_PIPE_TRACE) showed SMs waiting for points twice per call. The input
  // is one contiguous host array, so when it is page-locked the whole of it is sent at once on a copy stream of its own, in chunk
  // order, with an event behind every chunk's last byte; a chunk's stream only waits for its event. The link then runs at full rate
  // from the first microsecond of the call, whatever the kernels do (device raw buffers: stride_bytes per point, the caller's layout).
  // The copy engine serves its queue in submission order ACROSS streams, and every chunk's cloud set sends a small table block
  // through the same engine: chunk k's big copy is therefore submitted just before chunk k's tables (k = 0, 1), and everything
  // that is left right after chunk 1 has been enqueued (measured: with all copies submitted up front the first kernel of the call
  // waited 2.9 ms for its 2 KB of tables behind 160 MB of points).
  struct Uploader {
    int n_arrays = 0, stride = 0, next = 0;
    bool odometry = false;
    const char* host_base[2] = {nullptr, nullptr};
    char* dev_base[2] = {nullptr, nullptr};
    const int32_t* offs[2] = {nullptr, nullptr};
    size_t sent[2] = {0, 0};
    cudaError_t submit_through(int k, const std::vector<std::pair<int, int>>& chunks, apd_handle h) {
      for (; next <= k && next < (int)chunks.size(); next++) {
        const int end_cloud = chunks[next].first + chunks[next].second + (odometry ? 1 : 0);   // clouds [.., end_cloud) must be on the device
        for (int a = 0; a < n_arrays; a++) {
          const size_t upto = (size_t)offs[a][end_cloud] * stride;
          if (upto > sent[a]) {
            cudaError_t e = cudaMemcpyAsync(dev_base[a] + sent[a], host_base[a] + sent[a], upto - sent[a], cudaMemcpyHostToDevice, h->copy_stream);
            if (e != cudaSuccess) return e;
            sent[a] = upto;
          }
        }
        cudaError_t e = cudaEventRecord(h->copy_events[next], h->copy_stream);
        if (e != cudaSuccess) return e;
      }
      return cudaSuccess;
    }
  } up;
  up.stride = stride_bytes;
  up.odometry = odometry;
  int mem_in = h->pipeline_mem;
  const float* dev_src = pts_src;
  const float* dev_tgt = pts_tgt;
  bool ahead = false;
  // (the device copy of the raw input is bounded: beyond 4 GB a call keeps the per-chunk path, whose staging buffer is one chunk long)
  const int last_cloud_in = odometry ? n_pairs : n_pairs - 1;
  const size_t raw_total = (size_t)(off_src[last_cloud_in + 1] - off_src[0]) * stride_bytes + (odometry ? 0 : (size_t)(off_tgt[last_cloud_in + 1] - off_tgt[0]) * stride_bytes);
  if (mem_in == APD_MEM_HOST && h->upload_ahead && slots[1].h && chunks.size() > 1 && raw_total <= ((size_t)4 << 30)) {
    cudaPointerAttributes at{};
    const bool pinned_s = cudaPointerGetAttributes(&at, pts_src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    const bool pinned_t = odometry || (cudaPointerGetAttributes(&at, pts_tgt) == cudaSuccess && at.type == cudaMemoryTypeHost);
    cudaGetLastError();
    ahead = pinned_s && pinned_t;
  }
  if (ahead) {
    if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    while (h->copy_events.size() < chunks.size()) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->copy_events.push_back(e);
    }
    const int n_arrays = odometry ? 1 : 2;
    const float* host_base[2] = {pts_src, pts_tgt};
    const int32_t* offs[2] = {off_src, off_tgt};
    const int last_cloud = odometry ? n_pairs : n_pairs - 1;   // index of the last cloud a pair touches
    const char* dev_base[2] = {nullptr, nullptr};
    for (int a = 0; a < n_arrays; a++) {
      const size_t first_byte = (size_t)offs[a][0] * stride_bytes, end_byte = (size_t)offs[a][last_cloud + 1] * stride_bytes;
      CK(h->raw_all[a].reserve(std::max<size_t>(end_byte - first_byte, 16) + 16));
      dev_base[a] = static_cast<const char*>(h->raw_all[a].p) - first_byte;   // so that dev_base + offset * stride addresses like the host array
    }
    up.n_arrays = n_arrays;
    for (int a = 0; a < 2; a++) { up.host_base[a] = reinterpret_cast<const char*>(host_base[a]); up.dev_base[a] = const_cast<char*>(dev_base[a]); up.offs[a] = offs[a]; }
    up.sent[0] = (size_t)offs[0][0] * stride_bytes;
    up.sent[1] = odometry ? 0 : (size_t)offs[1][0] * stride_bytes;
    dev_src = reinterpret_cast<const float*>(dev_base[0]);
    dev_tgt = odometry ? nullptr : reinterpret_cast<const float*>(dev_base[1]);
    mem_in = APD_MEM_DEVICE;
  }
  struct DrainCopy {
    apd_handle h; bool on;
    ~DrainCopy() {
      if (on && h->copy_stream) cudaStreamSynchronize(h->copy_stream);
      h->kernel_uploads = false;
      if (h->helper) h->helper->kernel_uploads = false;
    }
  } drain_copy{h, ahead};
  if (ahead) h->kernel_uploads = slots[1].h->kernel_uploads = true;
  // APD_PIPE_TRACE=1: host and device time stamps of every chunk on stderr (diagnostic)
  static const bool trace = getenv("APD_PIPE_TRACE") != nullptr;
  struct Tr { double h0, h1, h2, h3; cudaEvent_t e0, e1, e2; int np; };
  std::vector<Tr> tr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (chunk = 0; chunk < (int)chunks.size(); chunk++) {
    const int p0 = chunks[chunk].first, np = chunks[chunk].second;
    ChunkSlot& s = slots[slots[1].h ? (chunk & 1) : 0];
    Tr t{};
    if (trace) t.h0 = now();
    int rc = retire(h, s, &lin, &err);
    if (rc) return rc;
    apd_handle hc = s.h;
    if (trace) { t.h1 = now(); t.np = np; cudaEventCreate(&t.e0); cudaEventCreate(&t.e1); cudaEventCreate(&t.e2); cudaEventRecord(t.e0, hc->stream); }
    if (ahead) {
      CK(up.submit_through(chunk, chunks, h));
      hc->fill_wait = h->copy_events[chunk];   // consumed by the first make_cloudset below, after its table upload
    }
    AlignCall c;
    if (odometry) {
      rc = make_cloudset(hc, dev_src, stride_bytes, off_src + p0, np + 1, mem_in, &s.a);
      if (rc) return hc == h ? rc : fail(h, rc, hc->err);
      c.src = c.tgt = s.a.get();
      c.src_base = 1;   // pair i: scan i + 1 onto scan i; no index arrays to upload
      c.tgt_base = 0;
    } else {
      rc = make_cloudset(hc, dev_src, stride_bytes, off_src + p0, np, mem_in, &s.a);
      if (rc) return hc == h ? rc : fail(h, rc, hc->err);
      if (ahead) hc->fill_wait = h->copy_events[chunk];
      rc = make_cloudset(hc, dev_tgt, stride_bytes, off_tgt + p0, np, mem_in, &s.b);
      if (rc) return hc == h ? rc : fail(h, rc, hc->err);
      c.src = s.a.get();
      c.tgt = s.b.get();
    }
    if (trace) { t.h2 = now(); cudaEventRecord(t.e1, hc->stream); }
    c.guesses = guesses ? guesses + (size_t)p0 * 16 : nullptr;
    c.n_pairs = np;
    c.plan_for_pairs = n_pairs;
    c.max_range = h->fitness_max_range;
    c.min_points = h->prm.k_correspondences;
    rc = run_align(hc, c);
    if (rc) return hc == h ? rc : fail(h, rc, hc->err);
    if (cudaMemcpyAsync(out + p0, hc->results.p, sizeof(apd_result) * np, cudaMemcpyDeviceToHost, hc->stream) != cudaSuccess) {
      cudaGetLastError();
      return fail(h, APD_ERR_CUDA, "result copy failed");
    }
    s.pairs = np;
    if (ahead) CK(up.submit_through(chunk >= 1 ? (int)chunks.size() - 1 : chunk + 1, chunks, h));
    if (trace) { t.h3 = now(); cudaEventRecord(t.e2, hc->stream); tr.push_back(t); }
  }
  for (ChunkSlot& s : slots) {
    int rc = retire(h, s, &lin, &err);
    if (rc) return rc;
  }
  if (trace && !tr.empty()) {
    const double hend = now();
    for (size_t i = 0; i < tr.size(); i++) {
      float a = 0, b = 0, c2 = 0;
      cudaEventElapsedTime(&a, tr[0].e0, tr[i].e0); cudaEventElapsedTime(&b, tr[0].e0, tr[i].e1); cudaEventElapsedTime(&c2, tr[0].e0, tr[i].e2);
      fprintf(stderr, "chunk %zu pairs %d | host: enter %.3f retired %.3f prepared-enqueued %.3f align-enqueued %.3f | device: start %.3f upload+build+knn done %.3f align+d2h done %.3f\n", i,
              tr[i].np, tr[i].h0 - tr[0].h0, tr[i].h1 - tr[0].h0, tr[i].h2 - tr[0].h0, tr[i].h3 - tr[0].h0, a, b, c2);
    }
    fprintf(stderr, "call returned at host %.3f ms\n", hend - tr[0].h0);
    for (Tr& t : tr) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
  }
  h->work_lin = lin;
  h->work_err = err;
  h->work_pairs = n_pairs;
  h->last_lin_valid = h->fit_valid = false;
  if (slots[1].h) h->stats.launches += slots[1].h->stats.launches - h->helper_launches_seen, h->helper_launches_seen = slots[1].h->stats.launches;
  return APD_OK;
}

}  // namespace

int apd_batch_align(apd_handle h, const float* pts_src, const int32_t* off_src, const float* pts_tgt, const int32_t* off_tgt, int stride_bytes,
                    const float* guesses, int n_pairs, apd_result* out) {
  if (!h || !out || n_pairs < 0 || !off_src || !off_tgt) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (n_pairs == 0) return APD_OK;
  return pipelined_align(h, pts_src, off_src, pts_tgt, off_tgt, stride_bytes, guesses, n_pairs, out, false);
}

int apd_odometry_align(apd_handle h, const float* pts, const int32_t* offsets, int n_scans, int stride_bytes, const float* guesses, apd_result* out) {
  if (!h || !out || n_scans < 0 || !offsets) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (n_scans < 2) return APD_OK;
  return pipelined_align(h, pts, offsets, nullptr, nullptr, stride_bytes, guesses, n_scans - 1, out, true);
}

// ---- "next" rows 8(f)-4 / 8(f)-2: preprocessing filters and submap accumulation ----

int apd_default_preprocess_params(apd_preprocess_params* p) {
  if (!p) return APD_ERR_INVALID;
  memset(p, 0, sizeof(*p));
  p->use_distance_filter = 1;
  p->outlier_removal = 1;  // launch file: outlier_removal_method RADIUS (radar_graph_slam.launch:59)
  p->radius_min_neighbors = 2;
  p->distance_near_thresh = 1.0;
  p->distance_far_thresh = 100.0;
  p->z_low_thresh = -5.0;
  p->z_high_thresh = 20.0;
  p->downsample_resolution = 0.1;
  p->radius_radius = 0.8;
  p->statistical_mean_k = 20;
  p->statistical_stddev = 1.0;
  return APD_OK;
}

static int preprocess_scratch(apd_handle h, size_t n) {
  n = std::max<size_t>(n, 1);
  CK(h->pp_a.reserve(sizeof(float4) * n));
  CK(h->pp_b.reserve(sizeof(float4) * n));
  CK(h->pp_flag.reserve(n));
  CK(h->pp_n.reserve(sizeof(int) * 8));
  CK(h->pp_ws.reserve(sizeof(unsigned) * 4 * n));
  CK(h->pp_seg.reserve(sizeof(int) * (n + 1)));
  return APD_OK;
}

int apd_preprocess(apd_handle h, const float* points, int stride_bytes, int intensity_offset_bytes, int n, const apd_preprocess_params* p, float* out, int* n_out) {
  if (!h || !p || !n_out || n < 0) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  *n_out = 0;
  if (stride_bytes < 16 || stride_bytes % 4) return fail(h, APD_ERR_INVALID, "stride_bytes must be a multiple of 4 and at least 16 (x, y, z, intensity)");
  if (intensity_offset_bytes < 12 || intensity_offset_bytes % 4 || intensity_offset_bytes + 4 > stride_bytes)
    return fail(h, APD_ERR_INVALID, "intensity_offset_bytes must address a float inside the record, after x, y, z");
  if (p->outlier_removal < 0 || p->outlier_removal > 2) return fail(h, APD_ERR_UNSUPPORTED, "outlier_removal must be 0 (NONE), 1 (RADIUS) or 2 (STATISTICAL)");
  if (p->outlier_removal == 2 && (p->statistical_mean_k < 1 || p->statistical_mean_k > 31)) return fail(h, APD_ERR_UNSUPPORTED, "statistical_mean_k must be in [1, 31]");
  if (p->outlier_removal == 1 && (p->radius_min_neighbors < 0 || !(p->radius_radius > 0))) return fail(h, APD_ERR_INVALID, "bad radius outlier parameters");
  if (n == 0) return APD_OK;
  if (!points || !out) return fail(h, APD_ERR_INVALID, "null point pointer");
  const int sf = stride_bytes / 4, io = intensity_offset_bytes / 4;
  int rc = preprocess_scratch(h, (size_t)n);
  if (rc) return rc;
  const size_t raw_bytes = (size_t)n * stride_bytes;
  CK(h->raw_upload.reserve(raw_bytes));
  CK(cudaMemcpyAsync(h->raw_upload.p, points, raw_bytes, cudaMemcpyHostToDevice, h->stream));
  float4 *A = h->pp_a.as<float4>(), *B = h->pp_b.as<float4>();
  int* nd = h->pp_n.as<int>();
  unsigned char* flag = h->pp_flag.as<unsigned char>();
  CK(launch_pack_xyzi(h->raw_upload.as<float>(), sf, io, n, A, h->stream, &h->stats));
  // distance_filter (preprocessing_nodelet.cpp:812); without it only the NaN removal of downsample() (:852-856) can drop points
  const bool voxel = p->downsample_resolution > 0;
  const int mode = p->use_distance_filter ? 0 : (voxel ? 2 : 1);
  CK(launch_distance_filter(A, n, p->distance_near_thresh, p->distance_far_thresh, p->z_low_thresh, p->z_high_thresh, mode, flag, B, nd + 0, h->stream, &h->stats));
  float4 *cur = B, *other = A;
  int* cur_n = nd + 0;
  if (voxel) {
    CK((h->downsample_method == 1 ? launch_approx_voxel_grid : launch_voxel_grid)(cur, cur_n, n, (float)p->downsample_resolution, h->pp_ws.as<unsigned>(), h->pp_seg.as<int>(), other,
                                                                                    nd + 1, h->stream, &h->stats));
    std::swap(cur, other);
    cur_n = nd + 1;
  }
  if (p->outlier_removal != 0) {
    int m = 0;
    CK(cudaMemcpyAsync(&m, cur_n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (m > 0) {
      // neighbour counting on the same uniform grid the scan matcher searches
      const int32_t off[2] = {0, m};
      std::shared_ptr<apd_cloudset_s> cs;
      rc = make_cloudset(h, reinterpret_cast<const float*>(cur), 16, off, 1, APD_MEM_DEVICE, &cs, /*force_grid=*/true);
      if (rc) return rc;
      rc = cloudset_build_grid(h, cs.get());
      if (rc) return rc;
      if (p->outlier_removal == 1) {
        CK(launch_radius_flags(cs->view(), m, p->radius_radius, p->radius_min_neighbors, flag, h->stream, &h->stats));
      } else if (m >= p->statistical_mean_k + 1) {
        CK(h->pp_seg.reserve(sizeof(float) * (size_t)m + 16));  // reused as the mean-distance array
        CK(launch_statistical_flags(cs->view(), m, cur_n, p->statistical_mean_k, p->statistical_stddev, h->pp_seg.as<float>(),
                                    reinterpret_cast<double*>(h->pp_n.as<int>() + 4), flag, h->stream, &h->stats));
      } else {
        CK(cudaMemsetAsync(flag, 1, (size_t)m, h->stream));  // fewer than mean_k + 1 points: returned unchanged
      }
      CK(launch_compact(cur, flag, cur_n, other, nd + 2, h->stream, &h->stats));
      std::swap(cur, other);
      cur_n = nd + 2;
      CK(cudaStreamSynchronize(h->stream));  // cs (pooled buffers) dies here; its kernels are done
    }
  }
  CK(launch_unpack_xyzi(cur, cur_n, n, sf, io, h->raw_upload.as<float>(), h->stream, &h->stats));
  int m = 0;
  CK(cudaMemcpyAsync(&m, cur_n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (m > 0) {
    CK(cudaMemcpyAsync(out, h->raw_upload.p, (size_t)m * stride_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  *n_out = m;
  return APD_OK;
}

int apd_cloudset_info(apd_cloudset cs, int32_t* n_clouds, int64_t* total_points) {
  if (!cs) return APD_ERR_INVALID;
  const apd_cloudset_s* s = deref(cs);
  if (n_clouds) *n_clouds = s->n_clouds;
  if (total_points) *total_points = s->total;
  return APD_OK;
}

int apd_build_submap(apd_handle h, apd_cloudset keyframes, const int32_t* which, int n_sel, const double* rel_poses, double downsample_resolution, uint64_t cache_key,
                     float* out_xyzi, int out_capacity, int* n_out) {
  if (!h || !keyframes || !n_out || n_sel < 0 || (n_sel > 0 && (!which || !rel_poses))) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  *n_out = 0;
  apd_cloudset_s* ks = deref(keyframes);
  std::vector<int> off(n_sel + 1, 0);
  for (int i = 0; i < n_sel; i++) {
    if (which[i] < 0 || which[i] >= ks->n_clouds) return fail(h, APD_ERR_INVALID, "keyframe index out of range");
    off[i + 1] = off[i] + (ks->h_off[which[i] + 1] - ks->h_off[which[i]]);
  }
  const int total = off[n_sel];
  h->last_lin_valid = h->fit_valid = false;
  h->has_last = false;
  if (total == 0) {  // an empty submap: the target is cleared (PCL would refuse an empty target at align time)
    h->tgt.reset();
    h->tgt_key = 0;
    return APD_OK;
  }
  int rc = preprocess_scratch(h, (size_t)total);
  if (rc) return rc;
  // selection tables: [which n_sel | out_off n_sel+1] ints, then poses (16 doubles each), 16-byte aligned
  const size_t ints = (size_t)(2 * n_sel + 1), pose_at = (ints * sizeof(int) + 15) & ~(size_t)15;
  std::vector<unsigned char> tab(pose_at + sizeof(double) * 16 * n_sel);
  memcpy(tab.data(), which, sizeof(int) * n_sel);
  memcpy(tab.data() + sizeof(int) * n_sel, off.data(), sizeof(int) * (n_sel + 1));
  memcpy(tab.data() + pose_at, rel_poses, sizeof(double) * 16 * n_sel);
  CK(h->pp_tab.reserve(tab.size()));
  CK(cudaMemcpyAsync(h->pp_tab.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, h->stream));
  const int* d_which = h->pp_tab.as<int>();
  const int* d_off = d_which + n_sel;
  const double* d_pose = reinterpret_cast<const double*>(h->pp_tab.as<unsigned char>() + pose_at);
  float4 *A = h->pp_a.as<float4>(), *B = h->pp_b.as<float4>();
  CK(launch_submap_gather(ks->view(), ks->pts.as<float4>(), d_which, d_off, n_sel, total, d_pose, A, h->stream, &h->stats));
  float4* cur = A;
  int m = total;
  if (downsample_resolution > 0) {
    int* nd = h->pp_n.as<int>();
    CK((h->downsample_method == 1 ? launch_approx_voxel_grid : launch_voxel_grid)(A, nullptr, total, (float)downsample_resolution, h->pp_ws.as<unsigned>(), h->pp_seg.as<int>(), B, nd,
                                                                                    h->stream, &h->stats));
    CK(cudaMemcpyAsync(&m, nd, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    cur = B;
  }
  CK(cudaStreamSynchronize(h->stream));  // also: the host table `tab` has been consumed
  // registration_s2m->setInputTarget(keyframe_cloud_s2m), scan_matching_odometry_nodelet.cpp:615, without leaving the device
  const int32_t one[2] = {0, m};
  std::shared_ptr<apd_cloudset_s> cs;
  rc = make_cloudset(h, reinterpret_cast<const float*>(cur), 16, one, 1, APD_MEM_DEVICE, &cs);
  if (rc) return rc;
  h->tgt = cs;
  h->tgt_key = cache_key;
  if (out_xyzi && out_capacity > 0) {
    CK(cudaMemcpyAsync(out_xyzi, cur, sizeof(float4) * (size_t)std::min(m, out_capacity), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  *n_out = m;
  return APD_OK;
}

int apd_synchronize(apd_handle h) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_get_timeline(apd_handle h, uint64_t* phase_ns /* 2 per stamp */, int max_stamps, int* n_stamps) {
  if (!h || !n_stamps) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  *n_stamps = 0;
  if (!h->timeline.p) return APD_OK;
  unsigned long long buf[512];
  CK(cudaMemcpyAsync(buf, h->timeline.p, sizeof(buf), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  int n = (int)std::min<unsigned long long>(buf[0], 230);
  // the last leaf build's phases (cloud 0), as phases 100 + k
  unsigned long long bs[16];
  CK(leaf_build_stamps(bs));
  for (int k = 0; k < 11; k++)
    if (bs[k]) { buf[1 + 2 * n] = 100 + k; buf[2 + 2 * n] = bs[k]; n++; }
  CK(knn_leaf_stamps(bs));  // the last leaf kNN launch's first group
  for (int k = 0; k < 8; k++)
    if (bs[k]) { buf[1 + 2 * n] = 120 + k; buf[2 + 2 * n] = bs[k]; n++; }
  if (bs[8]) { buf[1 + 2 * n] = 128; buf[2 + 2 * n] = bs[8]; n++; }              // latest group finish of the launch
  if (bs[9]) { buf[1 + 2 * n] = 129; buf[2 + 2 * n] = bs[0] + bs[9]; n++; }      // enter + duration of the longest group
  *n_stamps = std::min(n, max_stamps);
  if (phase_ns) memcpy(phase_ns, buf + 1, sizeof(unsigned long long) * 2 * std::min(n, max_stamps));
  return APD_OK;
}

// Streaming kernels of the path at a given size, timed with CUDA events on the handle's stream (bench.py "streaming" section): the only
// kernels whose roofline is HBM bandwidth by construction. Algorithmic bytes per point: pack_points 32 (one pcl::PointXYZI record) + 16;
// transform_points 16 + 12; cov_export 16 + 48 + 128; cov_import 128 + 16 + 48.
int apd_bench_streaming(apd_handle h, int n_points, int reps, double gbps[4], double ms[4]) {
  if (!h || !gbps || n_points <= 0 || reps <= 0) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  const size_t n = (size_t)n_points;
  DevBuf raw, cov16, out3, flush;
  for (DevBuf* b : {&raw, &cov16, &out3, &flush}) b->pool = h->pool;
  CK(raw.reserve(32 * n));
  CK(cov16.reserve(128 * n));
  CK(out3.reserve(12 * n));
  const size_t flush_bytes = 256u << 20;  // larger than the 126 MB L2: every timed repetition starts cold
  CK(flush.reserve(flush_bytes));
  CK(cudaMemsetAsync(raw.p, 0x3c, 32 * n, h->stream));
  CK(cudaMemsetAsync(flush.p, 0, flush_bytes, h->stream));
  CK(cudaMemsetAsync(cov16.p, 0, 128 * n, h->stream));
  const int32_t off[2] = {0, n_points};
  std::shared_ptr<apd_cloudset_s> cs;
  const int saved_unstaged = h->force_unstaged;
  int rc = make_cloudset(h, raw.as<float>(), 32, off, 1, APD_MEM_DEVICE, &cs, /*force_grid=*/true);
  if (rc) return rc;
  h->force_unstaged = saved_unstaged;
  // spts.w must hold a permutation for the export / import kernels: the identity is as good as any (the grid is not built here)
  CK(cudaMemcpyAsync(cs->spts.p, cs->pts.p, sizeof(float4) * n, cudaMemcpyDeviceToDevice, h->stream));
  CK(launch_iota_w(cs->spts.as<float4>(), n_points, h->stream, &h->stats));
  CK(h->misc.reserve(sizeof(float) * 16 + 64));
  const float T[16] = {0.99f, -0.1f, 0.f, 0.3f, 0.1f, 0.99f, 0.f, -0.2f, 0.f, 0.f, 1.f, 0.05f, 0.f, 0.f, 0.f, 1.f};
  CK(cudaMemcpyAsync(h->misc.p, T, sizeof(T), cudaMemcpyHostToDevice, h->stream));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double bytes[4] = {48.0, 28.0, 192.0, 192.0};
  for (int k = 0; k < 4; k++) {
    double total = 0.0;
    for (int r = -1; r < reps; r++) {  // r = -1: warm-up
      CK(launch_l2_flush(flush.p, flush_bytes, out3.as<float>(), h->stream));  // clean eviction: every repetition starts cold
      CK(cudaEventRecord(e0, h->stream));
      if (k == 0) CK(launch_pack_points(raw.as<float>(), 8, (long long)n, cs->pts.as<float4>(), h->stream, &h->stats));
      else if (k == 1) CK(launch_transform_points(cs->pts.as<float4>(), n_points, h->misc.as<float>(), out3.as<float>(), 3, h->stream, &h->stats));
      else if (k == 2) CK(launch_cov_export(cs->view(), 0, cov16.as<double>(), h->stream, &h->stats));
      else CK(launch_cov_import(cs->view(), 0, cov16.as<double>(), h->stream, &h->stats));
      CK(cudaEventRecord(e1, h->stream));
      CK(cudaEventSynchronize(e1));
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, e0, e1));
      if (r >= 0) total += t;
    }
    const double per = total / reps;
    if (ms) ms[k] = per;
    gbps[k] = bytes[k] * (double)n / (per * 1e-3) / 1e9;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_get_search_counters(apd_handle h, int64_t* knn_evals, int64_t* nn1_evals) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  long long k = 0, n = 0;
  for (apd_handle x : {h, h->helper}) {
    if (!x || !x->search_counters.p) continue;
    unsigned long long c[2] = {0, 0};
    CK(cudaMemcpyAsync(c, x->search_counters.p, sizeof(c), cudaMemcpyDeviceToHost, x->stream));
    CK(cudaMemsetAsync(x->search_counters.p, 0, sizeof(unsigned long long) * 4, x->stream));
    CK(cudaStreamSynchronize(x->stream));
    k += (long long)c[0];
    n += (long long)c[1];
  }
  if (knn_evals) *knn_evals = k;
  if (nn1_evals) *nn1_evals = n;
  return APD_OK;
}

int apd_get_kernel_times(apd_handle h, double ms[4], int64_t launches[4]) {
  if (!h || !ms) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  for (int k = 0; k < 4; k++) { ms[k] = 0.0; if (launches) launches[k] = 0; }
  for (apd_handle x : {h, h->helper}) {
    if (!x) continue;
    CK(cudaStreamSynchronize(x->stream));
    for (auto& t : x->timed) {
      float e = 0.f;
      if (cudaEventElapsedTime(&e, t.e0, t.e1) == cudaSuccess && t.kind >= 0 && t.kind < 4) {
        ms[t.kind] += (double)e;
        if (launches) launches[t.kind]++;
      } else {
        cudaGetLastError();
      }
      x->event_pool.push_back(t.e0);
      x->event_pool.push_back(t.e1);
    }
    x->timed.clear();
  }
  return APD_OK;
}

int apd_get_debug_counters(apd_handle h, uint64_t out[16]) {
  if (!h || !out) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  memset(out, 0, sizeof(uint64_t) * 16);
  if (!h->timeline.p) return APD_OK;
  CK(cudaMemcpyAsync(out, h->timeline.as<unsigned long long>() + 600, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return APD_OK;
}

int apd_get_launch_count(apd_handle h, int64_t* n) {
  if (!h || !n) return APD_ERR_INVALID;
  *n = h->stats.launches;
  return APD_OK;
}

int apd_get_work_counters(apd_handle h, int64_t* linearize_passes, int64_t* error_passes, int64_t* pairs) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (h->work_pairs > 0 && h->work_lin == 0 && h->counters.p) {  // a device-output call: counters not fetched yet
    int rc = fetch_counters(h, (int)h->work_pairs);
    if (rc) return rc;
  }
  if (linearize_passes) *linearize_passes = h->work_lin;
  if (error_passes) *error_passes = h->work_err;
  if (pairs) *pairs = h->work_pairs;
  return APD_OK;
}

}  // extern "C"
