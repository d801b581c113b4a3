// "Next" rows of SURVEY.md §8(f): the point-cloud filters in front of the scan matcher and the submap
// accumulation behind it, so that a scan can stay on the device from the radar message to the pose.
//   distance filter            radar_graph_slam/apps/preprocessing_nodelet.cpp:880-896
//   pcl::VoxelGrid             preprocessing_nodelet.cpp:137-144,850-866 (centroid of every 0.1 m voxel, all fields)
//   pcl::RadiusOutlierRemoval  preprocessing_nodelet.cpp:176-184,868-878 (dense input: (min_pts+1)-th neighbour within r)
//   submap accumulation        radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:606-616
// Clouds here are float4 (x, y, z, intensity). The filters work on one cloud of at most a few ten thousand
// points at sensor rate, so each is ONE CTA (1024 threads) that keeps its bookkeeping in shared memory and
// orders its phases with block barriers; outputs keep PCL's order (input order for the two index filters,
// ascending voxel index for the voxel grid), sums inside a voxel run in ascending input order.
#include "apd_internal.h"

namespace apd {

namespace {

constexpr int kPT = 1024;  // threads of the single-CTA preprocessing kernels

// exclusive block scan of one value per thread (blockDim == kPT); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* s_warp, unsigned* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  __syncthreads();  // s_warp may still be read by the previous call
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const unsigned w = s_warp[threadIdx.x];
    unsigned wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (threadIdx.x >= d) wi += t;
    }
    s_warp[threadIdx.x] = wi - w;
    if (threadIdx.x == 31) s_warp[32] = wi;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp] + incl - v;
}

// order-preserving compaction of the points whose flag is set: every thread owns a contiguous segment
__device__ __forceinline__ int compact_by_flag(const float4* __restrict__ in, const unsigned char* __restrict__ flag, int n, float4* __restrict__ out, unsigned* s_warp) {
  const int seg = (n + kPT - 1) / kPT;
  const int j0 = min((int)threadIdx.x * seg, n), j1 = min(j0 + seg, n);
  unsigned cnt = 0;
  for (int j = j0; j < j1; j++) cnt += flag[j] ? 1u : 0u;
  unsigned total;
  unsigned pos = block_excl_scan(cnt, s_warp, &total);
  for (int j = j0; j < j1; j++)
    if (flag[j]) out[pos++] = in[j];
  return (int)total;
}

// preprocessing_nodelet.cpp:884-889
// mode 0: the distance filter; 1: keep finite points (pcl::removeNaNFromPointCloud in downsample(), :852-856); 2: keep everything
__global__ void __launch_bounds__(kPT) distance_filter_kernel(const float4* __restrict__ in, int n, double near_t, double far_t, double z_low, double z_high, int mode,
                                                              unsigned char* __restrict__ flag, float4* __restrict__ out, int* __restrict__ n_out) {
  __shared__ unsigned s_warp[33];
  for (int i = threadIdx.x; i < n; i += kPT) {
    const float4 p = in[i];
    const double d = (double)fsqrt(fadd(fadd(fmul(p.x, p.x), fmul(p.y, p.y)), fmul(p.z, p.z)));  // getVector3fMap().norm()
    const double z = (double)p.z;
    bool keep = d > near_t && d < far_t && z < z_high && z > z_low;
    if (mode == 1) keep = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
    if (mode == 2) keep = true;
    flag[i] = keep ? 1 : 0;
  }
  __syncthreads();
  const int m = compact_by_flag(in, flag, n, out, s_warp);
  if (threadIdx.x == 0) *n_out = m;
}

// pcl::VoxelGrid::applyFilter (pcl/filters/impl/voxel_grid.hpp, PCL 1.10; downsample_all_data, no filter field): bounding box of the finite
// points, voxel index from floor(coordinate * inverse leaf), points grouped by voxel index, one centroid (all fields) per voxel.
// Workspace: keys/vals ping-pong (4 x n u32), seg_start (n + 1 int).
struct VoxelWs {
  unsigned *key_a, *val_a, *key_b, *val_b;
  int* seg_start;
};

__device__ __forceinline__ unsigned enc_f(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// Stable LSD radix sort of (key, value) pairs by one CTA of kPT threads, 8 bits per pass, as many passes as maxkey needs. On return
// ka / va point at the sorted arrays (the caller's pointers are swapped once per pass).
__device__ __forceinline__ void block_radix_sort(unsigned*& ka, unsigned*& va, unsigned*& kb, unsigned*& vb, int n, unsigned maxkey, unsigned (*s_hist)[256], unsigned* s_warp) {
  // Stable LSD radix sort, 8 bits per pass. Every warp owns a contiguous band of 32-element rows (coalesced
  // loads and stores) and a private 256-bin histogram; bins are ranked digit-major / warp-minor, and inside a
  // row the lanes that share a digit are ordered by lane (__match_any_sync), so equal keys keep their input order.
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = (n + 31) >> 5;
  const int band = (rows + (kPT >> 5) - 1) / (kPT >> 5);
  const int r0 = min(warp * band, rows), r1 = min(r0 + band, rows);
  for (int shift = 0; shift < 32 && (maxkey >> shift) != 0u; shift += 8) {
    for (int d = lane; d < 256; d += 32) s_hist[warp][d] = 0u;
    __syncwarp();
    for (int r = r0; r < r1; r++) {
      const int j = (r << 5) + lane;
      if (j < n) atomicAdd(&s_hist[warp][(ka[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    {  // exclusive scan of the 256 x 32 bins in (digit, warp) order: 8 consecutive bins per thread
      unsigned loc[8], sum = 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int e = tid * 8 + k;
        loc[k] = s_hist[e & 31][e >> 5];
        sum += loc[k];
      }
      unsigned total;
      unsigned run = block_excl_scan(sum, s_warp, &total);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int e = tid * 8 + k;
        s_hist[e & 31][e >> 5] = run;
        run += loc[k];
      }
    }
    __syncthreads();
    for (int r = r0; r < r1; r++) {
      const int j = (r << 5) + lane;
      const bool valid = j < n;
      const unsigned k = valid ? ka[j] : 0u;
      const unsigned d = valid ? ((k >> shift) & 255u) : (0x10000u + (unsigned)lane);
      const unsigned same = __match_any_sync(0xFFFFFFFFu, d);
      const unsigned rank = __popc(same & ((1u << lane) - 1u));
      unsigned dst = 0u;
      if (valid) {
        dst = s_hist[warp][d] + rank;
        kb[dst] = k;
        vb[dst] = va[j];
      }
      __syncwarp();
      if (valid && rank == 0u) s_hist[warp][d] += __popc(same);
      __syncwarp();
    }
    __syncthreads();
    unsigned* t = ka; ka = kb; kb = t;
    t = va; va = vb; vb = t;
  }
}

__global__ void __launch_bounds__(kPT) voxel_grid_kernel(const float4* __restrict__ in, const int* __restrict__ n_in_dev, int n_in_host, float leaf, VoxelWs ws,
                                                         float4* __restrict__ out, int* __restrict__ n_out) {
  __shared__ unsigned s_hist[kPT / 32][256];  // per-warp digit histograms / scatter cursors
  __shared__ unsigned s_warp[33];
  __shared__ unsigned s_box[6];
  __shared__ int s_minb[3], s_mul[3], s_small;
  __shared__ unsigned s_maxkey;
  const int tid = threadIdx.x;
  const int n = n_in_dev ? *n_in_dev : n_in_host;
  const float inv = 1.0f / leaf;

  // bounding box of the finite points (getMinMax3D on a non-dense cloud)
  if (tid < 6) s_box[tid] = tid < 3 ? 0xFFFFFFFFu : 0u;
  if (tid == 0) s_maxkey = 0u;
  __syncthreads();
  unsigned mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  for (int i = tid; i < n; i += kPT) {
    const float4 p = in[i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const unsigned e[3] = {enc_f(p.x), enc_f(p.y), enc_f(p.z)};
#pragma unroll
      for (int a = 0; a < 3; a++) { mn[a] = min(mn[a], e[a]); mx[a] = max(mx[a], e[a]); }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = __reduce_min_sync(0xFFFFFFFFu, mn[a]);
    mx[a] = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { atomicMin(&s_box[a], mn[a]); atomicMax(&s_box[3 + a], mx[a]); }
  }
  __syncthreads();
  if (tid == 0) {
    s_small = 0;
    s_minb[0] = s_minb[1] = s_minb[2] = 0;
    s_mul[0] = s_mul[1] = s_mul[2] = 1;
    if (s_box[0] <= s_box[3]) {  // at least one finite point
      int div[3];
      long long d[3];
      for (int a = 0; a < 3; a++) {
        const float lo = dec_f(s_box[a]), hi = dec_f(s_box[3 + a]);
        d[a] = (long long)fmul(fsub(hi, lo), inv) + 1;
        s_minb[a] = (int)floorf(fmul(lo, inv));
        div[a] = (int)floorf(fmul(hi, inv)) - s_minb[a] + 1;
      }
      s_mul[1] = div[0];
      s_mul[2] = div[0] * div[1];
      if (d[0] * d[1] * d[2] > 2147483647ll) s_small = 1;  // "Leaf size is too small": PCL returns the input unchanged
    }
  }
  __syncthreads();
  if (s_small) {
    for (int i = tid; i < n; i += kPT) out[i] = in[i];
    if (tid == 0) *n_out = n;
    return;
  }

  // voxel index per finite point (non-finite points get key 0xFFFFFFFF and are dropped after the sort)
  unsigned local_max = 0u;
  for (int i = tid; i < n; i += kPT) {
    const float4 p = in[i];
    unsigned key = 0xFFFFFFFFu;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const int i0 = (int)fsub(floorf(fmul(p.x, inv)), (float)s_minb[0]);
      const int i1 = (int)fsub(floorf(fmul(p.y, inv)), (float)s_minb[1]);
      const int i2 = (int)fsub(floorf(fmul(p.z, inv)), (float)s_minb[2]);
      key = (unsigned)(i0 * s_mul[0] + i1 * s_mul[1] + i2 * s_mul[2]);
    }
    ws.key_a[i] = key;
    ws.val_a[i] = (unsigned)i;
    local_max = max(local_max, key);
  }
  local_max = __reduce_max_sync(0xFFFFFFFFu, local_max);
  if ((tid & 31) == 0) atomicMax(&s_maxkey, local_max);
  __syncthreads();
  const unsigned maxkey = s_maxkey;

  // stable sort by voxel index (block_radix_sort above); ka / va hold the result
  unsigned *ka = ws.key_a, *va = ws.val_a, *kb = ws.key_b, *vb = ws.val_b;
  block_radix_sort(ka, va, kb, vb, n, maxkey, s_hist, s_warp);
  const int lane = tid & 31, warp = tid >> 5;
  const int rows = (n + 31) >> 5;
  const int band = (rows + (kPT >> 5) - 1) / (kPT >> 5);
  const int r0 = min(warp * band, rows), r1 = min(r0 + band, rows);

  // voxel heads -> segment starts (valid keys only), then one thread per voxel accumulates its points in order
  unsigned carry = 0u;
  for (int r = r0; r < r1; r++) {
    const int j = (r << 5) + lane;
    const bool head = j < n && ka[j] != 0xFFFFFFFFu && (j == 0 || ka[j] != ka[j - 1]);
    carry += __popc(__ballot_sync(0xFFFFFFFFu, head));
  }
  unsigned total;
  unsigned pos = block_excl_scan(lane == 0 ? carry : 0u, s_warp, &total);
  pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
  for (int r = r0; r < r1; r++) {
    const int j = (r << 5) + lane;
    const bool head = j < n && ka[j] != 0xFFFFFFFFu && (j == 0 || ka[j] != ka[j - 1]);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, head);
    if (head) ws.seg_start[pos + __popc(m & ((1u << lane) - 1u))] = j;
    pos += __popc(m);
  }
  __syncthreads();
  const int n_vox = (int)total;
  // number of valid (finite) points = first index with key 0xFFFFFFFF; find it from the last segment
  for (int s = tid; s < n_vox; s += kPT) {
    const int a = ws.seg_start[s];
    const unsigned key = ka[a];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int b = a;
    while (b < n && ka[b] == key) {
      const float4 p = in[va[b]];
      sx = fadd(sx, p.x); sy = fadd(sy, p.y); sz = fadd(sz, p.z); si = fadd(si, p.w);
      b++;
    }
    const float cnt = (float)(b - a);
    out[s] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
  }
  if (tid == 0) *n_out = n_vox;
}

// pcl::ApproximateVoxelGrid::applyFilter (pcl/filters/impl/approximate_voxel_grid.hpp, PCL 1.10; downsample_all_data, histsize_ = 512),
// the APPROX_VOXELGRID branch of preprocessing_nodelet.cpp:145-149 / scan_matching_odometry_nodelet.cpp:156-160. PCL walks the points
// once with a 512-entry history table indexed by hash(ix, iy, iz) = (ix * 7171 + iy * 3079 + iz * 4231) & 511: a point whose voxel
// differs from the one its entry holds FLUSHES that entry (its centroid becomes the next output point) and starts a new one; at the end
// the non-empty entries are flushed in table order. Sequential as written, but every entry only ever sees the points that hash to it,
// in input order, so the result is a function of per-entry runs:
//   * stable sort of the points by hash -> per entry, its points in input order; a RUN = maximal stretch with the same (ix, iy, iz)
//   * every run is one output point: sum of its points in input order (float, all four fields), divided by the count
//   * a run that is not the last of its entry is flushed by the first point of the next run: output slot = number of flushing points
//     before that point in INPUT order (exclusive scan of a flag array in input order)
//   * the last run of an entry is flushed at the end: slot = (number of flushing points) + (number of non-empty entries before it)
// static_cast<int>(floor(x * inv)) of a non-finite or out-of-range float is what cvttss2si returns on the reference's x86: INT_MIN.
__device__ __forceinline__ int approx_cell(float v, float inv) {
  const float f = floorf(fmul(v, inv));
  return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000;
}
__device__ __forceinline__ void approx_cells(const float4& p, float inv, int c[3]) {
  c[0] = approx_cell(p.x, inv); c[1] = approx_cell(p.y, inv); c[2] = approx_cell(p.z, inv);
}

__global__ void __launch_bounds__(kPT) approx_voxel_grid_kernel(const float4* __restrict__ in, const int* __restrict__ n_in_dev, int n_in_host, float leaf, VoxelWs ws,
                                                                float4* __restrict__ out, int* __restrict__ n_out) {
  __shared__ unsigned s_hist[kPT / 32][256];
  __shared__ unsigned s_warp[33];
  __shared__ unsigned s_entry[512];   // entry non-empty flag, then rank among the non-empty entries
  const int tid = threadIdx.x;
  const int n = n_in_dev ? *n_in_dev : n_in_host;
  const float inv = 1.0f / leaf;      // inverse_leaf_size_ = Array3f::Ones() / leaf_size_
  for (int i = tid; i < n; i += kPT) {
    int c[3];
    approx_cells(in[i], inv, c);
    ws.key_a[i] = ((unsigned)c[0] * 7171u + (unsigned)c[1] * 3079u + (unsigned)c[2] * 4231u) & 511u;
    ws.val_a[i] = (unsigned)i;
  }
  if (tid < 512) s_entry[tid] = 0u;
  __syncthreads();
  unsigned *ka = ws.key_a, *va = ws.val_a, *kb = ws.key_b, *vb = ws.val_b;
  block_radix_sort(ka, va, kb, vb, n, 511u, s_hist, s_warp);
  const int lane = tid & 31, warp = tid >> 5;
  const int rows = (n + 31) >> 5;
  const int band = (rows + (kPT >> 5) - 1) / (kPT >> 5);
  const int r0 = min(warp * band, rows), r1 = min(r0 + band, rows);

  // run heads -> seg_start; flushing points flagged in INPUT order (kb); non-empty entries
  unsigned carry = 0u;
  for (int pass = 0; pass < 2; pass++) {
    unsigned pos = 0u, total = 0u;
    if (pass == 1) {
      pos = block_excl_scan(lane == 0 ? carry : 0u, s_warp, &total);
      pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
    }
    for (int r = r0; r < r1; r++) {
      const int j = (r << 5) + lane;
      bool head = false, first = false;
      if (j < n) {
        first = j == 0 || ka[j] != ka[j - 1];
        head = first;
        if (!first) {
          int c[3], q[3];
          approx_cells(in[va[j]], inv, c);
          approx_cells(in[va[j - 1]], inv, q);
          head = c[0] != q[0] || c[1] != q[1] || c[2] != q[2];
        }
      }
      const unsigned m = __ballot_sync(0xFFFFFFFFu, head);
      if (pass == 0) {
        carry += __popc(m);
        if (j < n) kb[va[j]] = (head && !first) ? 1u : 0u;
        if (first) s_entry[ka[j]] = 1u;
      } else {
        if (head) ws.seg_start[pos + __popc(m & ((1u << lane) - 1u))] = j;
        pos += __popc(m);
      }
    }
    if (pass == 1 && tid == 0) ws.seg_start[total] = n;
    if (pass == 1) carry = total;
    __syncthreads();
  }
  const int n_runs = (int)__shfl_sync(0xFFFFFFFFu, carry, 0);   // every thread holds `total` already; keeps the value warp-uniform
  // rank of every entry among the non-empty ones
  unsigned n_entries;
  {
    const unsigned v = tid < 512 ? s_entry[tid] : 0u;
    const unsigned ex = block_excl_scan(v, s_warp, &n_entries);
    __syncthreads();
    if (tid < 512) s_entry[tid] = ex;
  }
  // exclusive scan of the flush flags in input order: every thread owns a contiguous segment; vb[i] = flushes before point i
  unsigned n_flush;
  {
    const int seg = (n + kPT - 1) / kPT;
    const int j0 = min(tid * seg, n), j1 = min(j0 + seg, n);
    unsigned cnt = 0u;
    for (int j = j0; j < j1; j++) cnt += kb[j];
    unsigned run = block_excl_scan(cnt, s_warp, &n_flush);
    for (int j = j0; j < j1; j++) { vb[j] = run; run += kb[j]; }
  }
  __syncthreads();
  for (int s = tid; s < n_runs; s += kPT) {
    const int a = ws.seg_start[s], b = ws.seg_start[s + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int j = a; j < b; j++) {
      const float4 p = in[va[j]];
      sx = fadd(sx, p.x); sy = fadd(sy, p.y); sz = fadd(sz, p.z); si = fadd(si, p.w);   // hhe->centroid += scratch
    }
    const float cnt = (float)(b - a);
    const unsigned slot = (b < n && ka[b] == ka[a]) ? vb[va[b]] : n_flush + s_entry[ka[a]];
    out[slot] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);                      // hhe->centroid /= count
  }
  if (tid == 0) *n_out = n_runs;
}

// count of target points within the radius (d2 <= r2 as in PCL's dense branch), for RadiusOutlierRemoval
struct CountWithin {
  double r2;
  int count;
  __device__ __forceinline__ float bound2() const { return __int_as_float(0x7f800000); }
  __device__ __forceinline__ void offer(float d2, unsigned, int) { count += ((double)d2 <= r2) ? 1 : 0; }
  __device__ __forceinline__ void end_run() {}
};

__global__ void radius_flag_kernel(CloudSetView cs, double r2, float r2_up, int min_pts, unsigned char* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // position in the sorted order of cloud 0
  const int n = cs.pt_off[1] - cs.pt_off[0];
  if (i >= n) return;
  GridView<unsigned> G;
  G.g = cs.grid[0];
  G.n = n;
  G.spts = cs.spts;
  G.cells = cs.cells + cs.cell_off[0];
  const float4 p = cs.spts[i];
  CountWithin v{r2, 0};
  grid_ball_search(G, p.x, p.y, p.z, r2_up, v);
  // nn_dists[min_pts] (0-based, the query itself is entry 0) must exist and not exceed r^2
  flag[__float_as_uint(p.w)] = (v.count >= min_pts + 1) ? 1 : 0;
}

// pcl::StatisticalOutlierRemoval, first pass (statistical_outlier_removal.hpp applyFilterIndices): the mean distance of every
// point to its mean_k nearest neighbours. The exact search keeps the 32 smallest (d2, index) keys; entries 1 .. mean_k are the
// neighbours (entry 0 is the point itself), summed in ascending order as sqrt in double, stored as float in ORIGINAL order.
__global__ void __launch_bounds__(256) stat_mean_dist_kernel(CloudSetView cs, int mean_k, float* __restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // position in the sorted order of cloud 0
  const int n = cs.pt_off[1] - cs.pt_off[0];
  if (i >= n) return;
  GridView<unsigned> G;
  G.g = cs.grid[0];
  G.n = n;
  G.spts = cs.spts;
  G.cells = cs.cells + cs.cell_off[0];
  const float4 p = cs.spts[i];
  TopK<32> tk;
  tk.init();
  pyramid_search<unsigned, TopK<32>, true>(G, cs, 0, p.x, p.y, p.z, __int_as_float(0x7f800000), tk, kFineRingsKnn);
  double dist_sum = 0.0;
#pragma unroll
  for (int k = 1; k < 32; k++)
    if (k <= mean_k) dist_sum = dadd(dist_sum, sqrt((double)__uint_as_float((unsigned)(tk.key[k] >> 32))));
  dist[__float_as_uint(p.w)] = (float)(dist_sum / (double)mean_k);
}

// second pass: mean and standard deviation of the distances, accumulated by ONE thread in index order (the sums are
// order-dependent doubles; a few ten thousand additions), then the keep flag of every point
__global__ void stat_threshold_kernel(const float* __restrict__ dist, const int* __restrict__ n_dev, double stddev_mult, double* __restrict__ thr) {
  const int n = *n_dev;
  double sum = 0.0, sq_sum = 0.0;
  for (int i = 0; i < n; i++) {
    const float d = dist[i];
    sum = dadd(sum, (double)d);
    sq_sum = dadd(sq_sum, (double)fmul(d, d));
  }
  const double mean = sum / (double)n;
  const double variance = dsub(sq_sum, dmul(sum, sum) / (double)n) / dsub((double)n, 1.0);
  *thr = dadd(mean, dmul(stddev_mult, sqrt(variance)));
}

__global__ void stat_flag_kernel(const float* __restrict__ dist, const int* __restrict__ n_dev, const double* __restrict__ thr, unsigned char* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_dev) return;
  flag[i] = ((double)dist[i] > *thr) ? 0 : 1;
}

__global__ void __launch_bounds__(kPT) compact_kernel(const float4* __restrict__ in, const unsigned char* __restrict__ flag, const int* __restrict__ n_dev,
                                                      float4* __restrict__ out, int* __restrict__ n_out) {
  __shared__ unsigned s_warp[33];
  const int m = compact_by_flag(in, flag, *n_dev, out, s_warp);
  if (threadIdx.x == 0) *n_out = m;
}

// scan_matching_odometry_nodelet.cpp:609-613: pcl::transformPointCloud with a double transform, then concatenation
__global__ void submap_gather_kernel(CloudSetView cs, const int* __restrict__ which, const int* __restrict__ out_off, int n_sel, const double* __restrict__ poses,
                                     float4* __restrict__ out, const float4* __restrict__ intensity_src) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = out_off[n_sel];
  if (j >= total) return;
  int lo = 0, hi = n_sel - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (out_off[mid] <= j) lo = mid; else hi = mid - 1;
  }
  const int c = which[lo];
  const int src = cs.pt_off[c] + (j - out_off[lo]);
  const float4 p = intensity_src[src];
  const double* T = poses + (size_t)lo * 16;
  const double x = p.x, y = p.y, z = p.z;
  float4 q;
  q.x = (float)dadd(dadd(dadd(dmul(T[0], x), dmul(T[1], y)), dmul(T[2], z)), T[3]);
  q.y = (float)dadd(dadd(dadd(dmul(T[4], x), dmul(T[5], y)), dmul(T[6], z)), T[7]);
  q.z = (float)dadd(dadd(dadd(dmul(T[8], x), dmul(T[9], y)), dmul(T[10], z)), T[11]);
  q.w = p.w;
  out[j] = q;
}

__global__ void pack_xyzi_kernel(const float* __restrict__ raw, int stride_floats, int intensity_offset, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = raw + (size_t)i * stride_floats;
  out[i] = make_float4(p[0], p[1], p[2], p[intensity_offset]);
}

__global__ void unpack_xyzi_kernel(const float4* __restrict__ in, const int* __restrict__ n_dev, int stride_floats, int intensity_offset, float* __restrict__ raw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_dev) return;
  const float4 p = in[i];
  float* o = raw + (size_t)i * stride_floats;
  for (int k = 0; k < stride_floats; k++) o[k] = 0.f;
  o[0] = p.x; o[1] = p.y; o[2] = p.z;
  if (stride_floats >= 8) o[3] = 1.0f;  // pcl::PointXYZI: data[3] = 1
  o[intensity_offset] = p.w;
}

}  // namespace

#define APD_LAUNCH_CHECK()                      \
  do {                                          \
    if (st) st->launches++;                     \
    cudaError_t e_ = cudaGetLastError();        \
    if (e_ != cudaSuccess) return e_;           \
  } while (0)

cudaError_t launch_distance_filter(const float4* in, int n, double near_t, double far_t, double z_low, double z_high, int mode, unsigned char* flag, float4* out,
                                   int* n_out, cudaStream_t stream, LaunchStats* st) {
  distance_filter_kernel<<<1, kPT, 0, stream>>>(in, n, near_t, far_t, z_low, z_high, mode, flag, out, n_out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_voxel_grid(const float4* in, const int* n_in_dev, int n_in_host, float leaf, unsigned* ws_u32 /*4*n*/, int* seg_start /*n+1*/, float4* out, int* n_out,
                              cudaStream_t stream, LaunchStats* st) {
  VoxelWs ws{ws_u32, ws_u32 + n_in_host, ws_u32 + 2 * (size_t)n_in_host, ws_u32 + 3 * (size_t)n_in_host, seg_start};
  voxel_grid_kernel<<<1, kPT, 0, stream>>>(in, n_in_dev, n_in_host, leaf, ws, out, n_out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_approx_voxel_grid(const float4* in, const int* n_in_dev, int n_in_host, float leaf, unsigned* ws_u32 /*4*n*/, int* seg_start /*n+1*/, float4* out,
                                     int* n_out, cudaStream_t stream, LaunchStats* st) {
  VoxelWs ws{ws_u32, ws_u32 + n_in_host, ws_u32 + 2 * (size_t)n_in_host, ws_u32 + 3 * (size_t)n_in_host, seg_start};
  approx_voxel_grid_kernel<<<1, kPT, 0, stream>>>(in, n_in_dev, n_in_host, leaf, ws, out, n_out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_radius_flags(const CloudSetView& cs, int n, double radius, int min_pts, unsigned char* flag, cudaStream_t stream, LaunchStats* st) {
  if (n == 0) return cudaSuccess;
  const double r2 = radius * radius;
  float up = (float)r2;
  if ((double)up < r2) up = nextafterf(up, INFINITY);
  radius_flag_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cs, r2, up, min_pts, flag);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_statistical_flags(const CloudSetView& cs, int n, const int* n_dev, int mean_k, double stddev_mult, float* dist, double* thr, unsigned char* flag,
                                     cudaStream_t stream, LaunchStats* st) {
  if (n == 0) return cudaSuccess;
  stat_mean_dist_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cs, mean_k, dist);
  APD_LAUNCH_CHECK();
  stat_threshold_kernel<<<1, 1, 0, stream>>>(dist, n_dev, stddev_mult, thr);
  APD_LAUNCH_CHECK();
  stat_flag_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dist, n_dev, thr, flag);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_compact(const float4* in, const unsigned char* flag, const int* n_dev, float4* out, int* n_out, cudaStream_t stream, LaunchStats* st) {
  compact_kernel<<<1, kPT, 0, stream>>>(in, flag, n_dev, out, n_out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_submap_gather(const CloudSetView& cs, const float4* xyzi, const int* which, const int* out_off, int n_sel, int total, const double* poses, float4* out,
                                 cudaStream_t stream, LaunchStats* st) {
  if (total == 0) return cudaSuccess;
  submap_gather_kernel<<<(total + 255) / 256, 256, 0, stream>>>(cs, which, out_off, n_sel, poses, out, xyzi);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_pack_xyzi(const float* raw, int stride_floats, int intensity_offset, int n, float4* out, cudaStream_t stream, LaunchStats* st) {
  if (n == 0) return cudaSuccess;
  pack_xyzi_kernel<<<(n + 255) / 256, 256, 0, stream>>>(raw, stride_floats, intensity_offset, n, out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_unpack_xyzi(const float4* in, const int* n_dev, int n_max, int stride_floats, int intensity_offset, float* raw, cudaStream_t stream, LaunchStats* st) {
  if (n_max == 0) return cudaSuccess;
  unpack_xyzi_kernel<<<(n_max + 255) / 256, 256, 0, stream>>>(in, n_dev, stride_floats, intensity_offset, raw);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

}  // namespace apd
