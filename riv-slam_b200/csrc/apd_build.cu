// Grid construction for a ragged batch of clouds, plus the small utility kernels of the C ABI.
// This is the GPU counterpart of pcl::search::KdTree::setInputCloud, which the reference runs in
// setInputSource / setInputTarget (fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:96,106)
// and again in pcl::Registration::initCompute: a uniform voxel grid per cloud, points counting-sorted
// by cell, built entirely on the device (no host round trip: the cell size is chosen on the GPU
// from the bounding box and a per-cloud cell budget).
//
// Pipeline (all kernels are tile-driven: one CTA per (cloud, chunk of points)):
//   bbox -> grid_params -> count -> scan -> scatter -> cell_sort
// cell_sort orders every cell's run by original index so the layout, and with it every later
// summation order, is deterministic regardless of atomic scheduling.
#include <algorithm>

#include "apd_internal.h"
#include "apd_leaf.cuh"

namespace apd {

namespace {

constexpr int kTileThreads = 256;

__device__ __forceinline__ unsigned enc_f(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void bbox_init_kernel(unsigned* bbox, int n_clouds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clouds * 6) bbox[i] = (i % 6 < 3) ? 0xFFFFFFFFu : 0u;
}

__global__ void __launch_bounds__(kTileThreads) bbox_kernel(CloudSetView cs, const int4* __restrict__ tiles, unsigned* __restrict__ bbox) {
  const int4 tile = tiles[blockIdx.x];
  const int base = cs.pt_off[tile.x];
  unsigned mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  for (int i = threadIdx.x; i < tile.z; i += blockDim.x) {
    const float4 p = cs.pts[base + tile.y + i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const unsigned e[3] = {enc_f(p.x), enc_f(p.y), enc_f(p.z)};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        mn[a] = min(mn[a], e[a]);
        mx[a] = max(mx[a], e[a]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = __reduce_min_sync(0xFFFFFFFFu, mn[a]);
    mx[a] = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&bbox[tile.x * 6 + a], mn[a]);
      atomicMax(&bbox[tile.x * 6 + 3 + a], mx[a]);
    }
  }
}

// Bounding box -> cell edge h and grid dimensions with nx*ny*nz <= cap.
__device__ __forceinline__ GridParams make_grid_params(const float lo[3], const float hi[3], long long cap) {
  const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  if (cap < 1) cap = 1;
  // start from the edge that spends the whole budget on the box volume, then grow until it fits
  const float emax = fmaxf(fmaxf(ex, ey), ez);
  const float floor_h = fmaxf(emax * 1e-4f, 1e-6f);
  const float dx = fmaxf(ex, floor_h), dy = fmaxf(ey, floor_h), dz = fmaxf(ez, floor_h);
  float h = fmaxf(cbrtf(dx * dy * dz / (float)cap), floor_h);
  int nx = 1, ny = 1, nz = 1;
  for (int it = 0; it < 4096; it++) {
    const float fx = floorf(ex / h) + 1.f, fy = floorf(ey / h) + 1.f, fz = floorf(ez / h) + 1.f;
    if (fx * fy * fz <= (float)cap && fx < 2e9f && fy < 2e9f && fz < 2e9f) {
      nx = (int)fx; ny = (int)fy; nz = (int)fz;
      if ((long long)nx * ny * nz <= cap) break;
    }
    h *= 1.02f;
  }
  if ((long long)nx * ny * nz > cap) { nx = ny = nz = 1; h = fmaxf(emax, floor_h) * 2.f; }
  GridParams g;
  g.lox = lo[0]; g.loy = lo[1]; g.loz = lo[2];
  g.h = h;
  g.inv_h = 1.0f / h;
  g.nx = nx; g.ny = ny; g.nz = nz;
  g.ncells = nx * ny * nz;
  float amax = 0.f;
  for (int a = 0; a < 3; a++) amax = fmaxf(amax, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
  g.slack = 4e-6f * (amax + emax + h) + 1e-30f;
  return g;
}

// One thread per cloud.
__global__ void grid_params_kernel(CloudSetView cs, const unsigned* __restrict__ bbox, const int* __restrict__ cell_cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cs.n_clouds) return;
  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    const unsigned l = bbox[c * 6 + a], h = bbox[c * 6 + 3 + a];
    if (l > h) { lo[a] = 0.f; hi[a] = 0.f; }  // no finite point
    else { lo[a] = dec_f(l); hi[a] = dec_f(h); }
  }
  cs.grid[c] = make_grid_params(lo, hi, cell_cap[c]);
}

__global__ void __launch_bounds__(kTileThreads) count_kernel(CloudSetView cs, const int4* __restrict__ tiles, int* __restrict__ cellid) {
  const int4 tile = tiles[blockIdx.x];
  const int base = cs.pt_off[tile.x];
  const GridParams g = cs.grid[tile.x];
  unsigned* cells = cs.cells + cs.cell_off[tile.x];
  for (int i = threadIdx.x; i < tile.z; i += blockDim.x) {
    const float4 p = cs.pts[base + tile.y + i];
    const int cell = cell_index(g, p.x, p.y, p.z);
    cellid[base + tile.y + i] = cell;
    atomicAdd(&cells[cell], 1u);
  }
}

// One CTA per cloud: in-place exclusive scan of the ncells+1 counters; a copy goes to `cursor`.
__global__ void __launch_bounds__(1024) scan_kernel(CloudSetView cs, unsigned* __restrict__ cursor) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry;
  const int c = blockIdx.x;
  const long long off = cs.cell_off[c];
  unsigned* cells = cs.cells + off;
  unsigned* cur = cursor + off;
  const int total = cs.grid[c].ncells + 1;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  // coalesced chunks of blockDim elements: block-wide scan per chunk with a running carry
  for (int start = 0; start < total; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const unsigned v = (i < total) ? cells[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      const unsigned w = warp_sums[threadIdx.x];
      unsigned wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
        if (threadIdx.x >= d) wi += t;
      }
      warp_sums[threadIdx.x] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const unsigned excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
    if (i < total) {
      cells[i] = excl;
      cur[i] = excl;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kTileThreads) scatter_kernel(CloudSetView cs, const int4* __restrict__ tiles, const int* __restrict__ cellid,
                                                               unsigned* __restrict__ cursor) {
  const int4 tile = tiles[blockIdx.x];
  const int base = cs.pt_off[tile.x];
  unsigned* cur = cursor + cs.cell_off[tile.x];
  for (int i = threadIdx.x; i < tile.z; i += blockDim.x) {
    const int li = tile.y + i;
    const float4 p = cs.pts[base + li];
    const unsigned pos = atomicAdd(&cur[cellid[base + li]], 1u);
    cs.spts[base + pos] = make_float4(p.x, p.y, p.z, __uint_as_float((unsigned)li));
  }
}

// Order every cell's run by original index (the scatter's arrival order depends on atomic timing).
// A warp inspects 32 cells at a time and sorts the ones holding two or more points cooperatively by
// rank counting: every element counts the smaller keys of its cell, all reads happen before any
// write, so there is no dependent chain of global-memory round trips (the first version's per-thread
// insertion sort cost 180 us on one 5000-point cloud because of it).
__device__ __forceinline__ void warp_sort_cells(const unsigned* __restrict__ cells, float4* __restrict__ sp, int ncells, int warp, int n_warps) {
  const int lane = threadIdx.x & 31;
  for (int cell0 = warp * 32; cell0 < ncells; cell0 += n_warps * 32) {
    const int cell = cell0 + lane;
    int s = 0, m = 0;
    if (cell < ncells) {
      s = (int)cells[cell];
      m = (int)cells[cell + 1] - s;
    }
    unsigned todo = __ballot_sync(0xFFFFFFFFu, m >= 2);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int cs0 = __shfl_sync(0xFFFFFFFFu, s, src), cm = __shfl_sync(0xFFFFFFFFu, m, src);
      if (cm <= 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned key = 0xFFFFFFFFu;
        if (lane < cm) {
          v = sp[cs0 + lane];
          key = __float_as_uint(v.w);
        }
        int rank = 0;
        for (int j = 0; j < cm; j++) rank += (__shfl_sync(0xFFFFFFFFu, key, j) < key) ? 1 : 0;
        __syncwarp();
        if (lane < cm) sp[cs0 + rank] = v;
      } else if (cm <= 256) {
        float4 v[8];
        int rank[8];
#pragma unroll
        for (int t = 0; t < 8; t++) {
          rank[t] = 0;
          v[t] = (lane + 32 * t < cm) ? sp[cs0 + lane + 32 * t] : make_float4(0.f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu));
        }
        for (int j = 0; j < cm; j++) {
          const unsigned kj = __float_as_uint(sp[cs0 + j].w);  // nothing has been written yet
#pragma unroll
          for (int t = 0; t < 8; t++) rank[t] += (kj < __float_as_uint(v[t].w)) ? 1 : 0;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; t++)
          if (lane + 32 * t < cm) sp[cs0 + rank[t]] = v[t];
      } else if (lane == 0) {  // a pathological cell (hundreds of points): plain insertion sort
        for (int i = cs0 + 1; i < cs0 + cm; i++) {
          const float4 v = sp[i];
          const unsigned key = __float_as_uint(v.w);
          int j = i - 1;
          while (j >= cs0 && __float_as_uint(sp[j].w) > key) {
            sp[j + 1] = sp[j];
            j--;
          }
          sp[j + 1] = v;
        }
      }
      __syncwarp();
    }
  }
}

__global__ void cell_sort_kernel(CloudSetView cs, int cloud_begin) {
  const int c = cloud_begin + blockIdx.y;
  if (c >= cs.n_clouds) return;
  warp_sort_cells(cs.cells + cs.cell_off[c], cs.spts + cs.pt_off[c], cs.grid[c].ncells, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5),
                  gridDim.x * (blockDim.x >> 5));
}

// ---- fused build for small clouds: ONE CTA builds one pyramid level of one cloud ----
// bbox -> grid parameters -> count -> scan -> scatter -> (finest level) cell sort + inverse order, with
// block barriers instead of kernel boundaries. Replaces ~23 launches per cloud set by one; on a single
// 5000-point scan that is the difference between ~130 us and ~20 us of the align latency.
struct FusedLevels {
  const int* cap[1 + kCoarseLevels];      // per-cloud cell budgets of each level
  int* cellid[1 + kCoarseLevels];         // per-point cell ids (workspace, one array per level)
  unsigned* cursor[1 + kCoarseLevels];    // scatter cursors (workspace, laid out like the level's cell table)
};

// Step 1 of both fused builds: bounding box of the cloud's finite points (block-wide), grid parameters from the box and
// the level's cell budget (thread 0), published to the level's GridParams slot and returned to every thread.
__device__ __forceinline__ GridParams block_grid_params(const float4* __restrict__ pts, int n, long long cap, GridParams* out_slot, unsigned (*s_box)[6], GridParams* s_g) {
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  unsigned mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  for (int i = tid; i < n; i += T) {
    const float4 p = pts[i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const unsigned e[3] = {enc_f(p.x), enc_f(p.y), enc_f(p.z)};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        mn[a] = min(mn[a], e[a]);
        mx[a] = max(mx[a], e[a]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = __reduce_min_sync(0xFFFFFFFFu, mn[a]);
    mx[a] = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
  }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { s_box[warp][a] = mn[a]; s_box[warp][3 + a] = mx[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      unsigned l = 0xFFFFFFFFu, h = 0u;
      for (int w = 0; w < (T >> 5); w++) { l = min(l, s_box[w][a]); h = max(h, s_box[w][3 + a]); }
      if (l > h) { lo[a] = 0.f; hi[a] = 0.f; }
      else { lo[a] = dec_f(l); hi[a] = dec_f(h); }
    }
    *s_g = make_grid_params(lo, hi, cap);
    *out_slot = *s_g;
  }
  __syncthreads();
  return *s_g;
}

__global__ void __launch_bounds__(1024) build_fused_kernel(CloudSetView cs, FusedLevels L) {
  __shared__ unsigned s_box[32][6];
  __shared__ GridParams s_g;
  __shared__ unsigned s_warp[32];
  const int c = blockIdx.x, level = blockIdx.y;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int base = cs.pt_off[c];
  const int n = cs.pt_off[c + 1] - base;
  const float4* pts = cs.pts + base;
  float4* spts = (level == 0 ? cs.spts : cs.coarse[level - 1].spts) + base;
  const long long coff = level == 0 ? cs.cell_off[c] : cs.coarse[level - 1].cell_off[c];
  unsigned* cells = (level == 0 ? cs.cells : cs.coarse[level - 1].cells) + coff;
  unsigned* cursor = L.cursor[level] + coff;
  int* cellid = L.cellid[level] + base;

  // 1. bounding box of the finite points -> grid parameters
  const GridParams g = block_grid_params(pts, n, L.cap[level][c], &(level == 0 ? cs.grid : cs.coarse[level - 1].grid)[c], s_box, &s_g);
  const int total = g.ncells + 1;

  // 2. zero the counters, 3. count
  for (int j = tid; j < total; j += T) cells[j] = 0u;
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const float4 p = pts[i];
    const int cell = cell_index(g, p.x, p.y, p.z);
    cellid[i] = cell;
    atomicAdd(&cells[cell], 1u);
  }
  __syncthreads();

  // 4. exclusive scan: every thread owns a contiguous segment, the segment totals are scanned by the block
  const int seg = (total + T - 1) / T;
  const int j0 = min(tid * seg, total), j1 = min(j0 + seg, total);
  unsigned sum = 0;
  for (int j = j0; j < j1; j++) sum += cells[j];
  unsigned incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (tid < 32) {
    const unsigned w = tid < (T >> 5) ? s_warp[tid] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (tid >= d) wi += t;
    }
    s_warp[tid] = wi - w;
  }
  __syncthreads();
  unsigned run = s_warp[warp] + incl - sum;
  for (int j = j0; j < j1; j++) {
    const unsigned v = cells[j];
    cells[j] = run;
    cursor[j] = run;
    run += v;
  }
  __syncthreads();

  // 5. scatter
  for (int i = tid; i < n; i += T) {
    const float4 p = pts[i];
    const unsigned pos = atomicAdd(&cursor[cellid[i]], 1u);
    spts[pos] = make_float4(p.x, p.y, p.z, __uint_as_float((unsigned)i));
  }
  if (level != 0) return;
  __syncthreads();

  // 6. deterministic order inside every cell, then the inverse permutation
  warp_sort_cells(cells, spts, g.ncells, warp, T >> 5);
  __syncthreads();
  for (int i = tid; i < n; i += T) cs.inv0[base + __float_as_uint(spts[i].w)] = i;
}

// Shared-memory variant of the fused build for clouds whose sorted points (16 B each) and cell table (4 B per
// cell) fit the CTA's shared memory together (a 5000-point scan at 4 cells per point: 160 KB): counting, the
// scan, the scatter and the in-cell sort all run on shared memory; HBM/L2 sees three coalesced reads of the
// points and one coalesced write each of the cell table, the sorted points and the inverse permutation.
// The global-memory version above spent its time in L2 atomics and in strided scan segments
// (1.08 ms per 1001 scans x 3 levels).
// s_arr[0] = 0 and s_arr[1 + c] is cell c's counter, then (after the scan) its first slot, then (after the
// scatter, which advances it to the cell's end) the first slot of cell c + 1: at every stage s_arr is laid out
// so that the finished array IS the cell table cells[0 .. ncells].
__global__ void __launch_bounds__(1024) build_fused_smem_kernel(CloudSetView cs, FusedLevels L) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ unsigned s_box[32][6];
  __shared__ GridParams s_g;
  __shared__ unsigned s_warp[32];
  const int c = blockIdx.x, level = blockIdx.y;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int base = cs.pt_off[c];
  const int n = cs.pt_off[c + 1] - base;
  const float4* pts = cs.pts + base;
  float4* spts = (level == 0 ? cs.spts : cs.coarse[level - 1].spts) + base;
  const long long coff = level == 0 ? cs.cell_off[c] : cs.coarse[level - 1].cell_off[c];
  unsigned* cells = (level == 0 ? cs.cells : cs.coarse[level - 1].cells) + coff;
  float4* s_pts = reinterpret_cast<float4*>(sm_raw);
  unsigned* s_arr = reinterpret_cast<unsigned*>(sm_raw + sizeof(float4) * (size_t)n);

  // 1. bounding box of the finite points -> grid parameters
  const GridParams g = block_grid_params(pts, n, L.cap[level][c], &(level == 0 ? cs.grid : cs.coarse[level - 1].grid)[c], s_box, &s_g);
  const int E = g.ncells + 1;  // entries of the cell table

  // 2. zero the counters, 3. count
  for (int j = tid; j < E + 1; j += T) s_arr[j] = 0u;
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const float4 p = pts[i];
    atomicAdd(&s_arr[1 + cell_index(g, p.x, p.y, p.z)], 1u);
  }
  __syncthreads();

  // 4. exclusive scan of s_arr[1 .. E]: every warp owns a contiguous band of 32-entry rows
  unsigned* a = s_arr + 1;
  const int rows = (E + 31) >> 5;
  const int band = (rows + (T >> 5) - 1) / (T >> 5);
  const int r0 = min(warp * band, rows), r1 = min(r0 + band, rows);
  unsigned carry = 0u;
  for (int r = r0; r < r1; r++) {
    const int j = (r << 5) + lane;
    const unsigned v = j < E ? a[j] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += t;
    }
    if (j < E) a[j] = carry + incl - v;
    carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  if (lane == 0) s_warp[warp] = carry;
  __syncthreads();
  if (tid < 32) {
    const unsigned w = tid < (T >> 5) ? s_warp[tid] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (tid >= d) wi += t;
    }
    s_warp[tid] = wi - w;
  }
  __syncthreads();
  const unsigned woff = s_warp[warp];
  for (int r = r0; r < r1; r++) {
    const int j = (r << 5) + lane;
    if (j < E) {
      const unsigned v = a[j] + woff;
      a[j] = v;
      cells[j] = v;  // the finished table goes out now; the scatter below turns a[] into cursors
    }
  }
  __syncthreads();

  // 5. scatter into shared memory
  for (int i = tid; i < n; i += T) {
    const float4 p = pts[i];
    const unsigned pos = atomicAdd(&a[cell_index(g, p.x, p.y, p.z)], 1u);
    s_pts[pos] = make_float4(p.x, p.y, p.z, __uint_as_float((unsigned)i));
  }
  __syncthreads();

  // 6. finest level: deterministic order inside every cell (s_arr is the cell table again, see above)
  if (level == 0) {
    warp_sort_cells(s_arr, s_pts, g.ncells, warp, T >> 5);
    __syncthreads();
  }
  for (int i = tid; i < n; i += T) {
    const float4 v = s_pts[i];
    spts[i] = v;
    if (level == 0) cs.inv0[base + __float_as_uint(v.w)] = i;
  }
}

// original index -> position in the cell-sorted order (needed when a coarse pyramid level finds a
// neighbour and the caller wants its level-0 position)
__global__ void inverse_order_kernel(CloudSetView cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cs.total_points) return;
  // binary search of the cloud that owns sorted slot i
  int lo = 0, hi = cs.n_clouds - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (cs.pt_off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  const int base = cs.pt_off[lo];
  cs.inv0[base + __float_as_uint(cs.spts[i].w)] = i - base;
}

__global__ void pack_points_kernel(const float* __restrict__ xyz, int stride_floats, long long n, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + i * stride_floats;
  // .w is never read by the matcher; for pcl::PointXYZI records (stride >= 20 bytes) it carries the intensity for apd_build_submap
  out[i] = make_float4(p[0], p[1], p[2], stride_floats >= 5 ? p[4] : 1.0f);
}

// pcl::PointXYZI records (32 bytes, 16-byte aligned): every record is two float4 (x y z pad | intensity pad pad pad), read with
// 128-bit loads - a warp reads 1 KB contiguous per instruction instead of three strided scalars per lane - four records per thread
__global__ void __launch_bounds__(256) pack_points_xyzi_kernel(const float4* __restrict__ rec, long long n, float4* __restrict__ out) {
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float4 a[4], b[4];
#pragma unroll
  for (int u = 0; u < 4; u++)
    if (i0 + u < n) { a[u] = __ldcs(rec + 2 * (i0 + u)); b[u] = __ldcs(rec + 2 * (i0 + u) + 1); }
#pragma unroll
  for (int u = 0; u < 4; u++)
    if (i0 + u < n) out[i0 + u] = make_float4(a[u].x, a[u].y, a[u].z, b[u].x);
}

// pcl::transformPointCloud(*input_, output, T) at lsq_registration_impl.hpp:79 (float arithmetic)
__global__ void transform_points_kernel(const float4* __restrict__ pts, int n, const float* __restrict__ T, float* __restrict__ out, int out_stride_floats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = pts[i];
  float* o = out + (size_t)i * out_stride_floats;
  o[0] = xform_row_rn(T[0], T[1], T[2], T[3], a.x, a.y, a.z);
  o[1] = xform_row_rn(T[4], T[5], T[6], T[7], a.x, a.y, a.z);
  o[2] = xform_row_rn(T[8], T[9], T[10], T[11], a.x, a.y, a.z);
}

// The same for a packed xyz output (12 bytes per point): a CTA stages its 256 results in shared memory and writes them as 768
// consecutive floats with 128-bit stores (three scalar stores per lane at stride 12 B used a third of every sector write)
__global__ void __launch_bounds__(256) transform_points_packed_kernel(const float4* __restrict__ pts, int n, const float* __restrict__ T, float* __restrict__ out) {
  __shared__ __align__(16) float s[256 * 3];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    const float4 a = __ldcs(pts + i);
    s[threadIdx.x * 3 + 0] = xform_row_rn(T[0], T[1], T[2], T[3], a.x, a.y, a.z);
    s[threadIdx.x * 3 + 1] = xform_row_rn(T[4], T[5], T[6], T[7], a.x, a.y, a.z);
    s[threadIdx.x * 3 + 2] = xform_row_rn(T[8], T[9], T[10], T[11], a.x, a.y, a.z);
  }
  __syncthreads();
  const int first = blockIdx.x * 256, cnt = min(256, n - first);   // cnt * 3 floats, starting at a multiple of 768 floats: 16-byte aligned
  float* o = out + (size_t)first * 3;
  const int nf = cnt * 3;
  if (threadIdx.x < 192) {
    const int f = threadIdx.x * 4;
    if (f + 3 < nf) *reinterpret_cast<float4*>(o + f) = *reinterpret_cast<const float4*>(s + f);
    else for (int k = f; k < nf; k++) o[k] = s[k];
  }
}

// sorted 6-double covariances <-> Eigen::Matrix4d layout (16 doubles, symmetric so row/column order is moot) in original order.
// Eight lanes per point: lane j of a point's group moves doubles 2j, 2j+1 of the 4x4, so a warp writes (reads) four complete
// 128-byte matrices with one 16-byte access per lane instead of sixteen scattered 8-byte accesses per lane.
__device__ __forceinline__ double cov16_entry(const double2& a, const double2& b, const double2& c, int e) {
  // row-major 4x4: [xx xy xz 0 | xy yy yz 0 | xz yz zz 0 | 0 0 0 0]; a = (xx, xy), b = (xz, yy), c = (yz, zz)
  switch (e) {
    case 0: return a.x; case 1: return a.y; case 2: return b.x;
    case 4: return a.y; case 5: return b.y; case 6: return c.x;
    case 8: return b.x; case 9: return c.x; case 10: return c.y;
    default: return 0.0;
  }
}
__global__ void __launch_bounds__(256) cov_export_kernel(CloudSetView cs, int cloud, double* __restrict__ out16) {
  const int n = cs.pt_off[cloud + 1] - cs.pt_off[cloud];
  const int base = cs.pt_off[cloud];
  // four (point, lane-of-8) items per thread, a whole grid apart: four independent chains of loads in flight
  const long long total = (long long)n * 8, stride = (long long)gridDim.x * blockDim.x;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t < total; t += 4 * stride) {
    unsigned orig[4];
    double2 a[4], b[4], c[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long tu = t + u * stride;
      if (tu < total) {
        const int gi = base + (int)(tu >> 3);
        orig[u] = __float_as_uint(cs.spts[gi].w);
        a[u] = cs.cov0[gi]; b[u] = cs.cov1[gi]; c[u] = cs.cov2[gi];  // the 8 lanes of a point read the same 48 bytes (broadcast)
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long tu = t + u * stride;
      if (tu < total) {
        const int j = (int)(tu & 7);
        __stcs(reinterpret_cast<double2*>(out16 + (size_t)orig[u] * 16) + j, make_double2(cov16_entry(a[u], b[u], c[u], 2 * j), cov16_entry(a[u], b[u], c[u], 2 * j + 1)));
      }
    }
  }
}

// one thread per point: the six doubles kept (entries 0, 1, 2, 5, 6, 10 of the 4x4) sit in five 16-byte pieces of the record, loaded
// independently; three 16-byte stores
__global__ void __launch_bounds__(256) cov_import_kernel(CloudSetView cs, int cloud, const double* __restrict__ in16) {
  const int n = cs.pt_off[cloud + 1] - cs.pt_off[cloud];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int gi = cs.pt_off[cloud] + i;
  const unsigned orig = __float_as_uint(cs.spts[gi].w);
  const double2* r = reinterpret_cast<const double2*>(in16 + (size_t)orig * 16);
  const double2 p0 = __ldcs(r), p1 = __ldcs(r + 1), p2 = __ldcs(r + 2), p3 = __ldcs(r + 3), p5 = __ldcs(r + 5);
  cs.cov0[gi] = make_double2(p0.x, p0.y);   // xx, xy
  cs.cov1[gi] = make_double2(p1.x, p2.y);   // xz, yy
  cs.cov2[gi] = make_double2(p3.x, p5.x);   // yz, zz
}

// reads `n` float4 and keeps nothing: evicts the L2 with CLEAN lines (a memset would leave 126 MB of dirty lines whose write-back then
// competes with the kernel being timed)
__global__ void l2_flush_read_kernel(const float4* __restrict__ buf, size_t n, float* __restrict__ sink) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(buf + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) *sink = acc;
}

__global__ void iota_w_kernel(float4* __restrict__ pts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pts[i].w = __uint_as_float((unsigned)i);
}

// correspondences_ / sq_distances_ / mahalanobis_ of the last linearize in ORIGINAL source order with
// ORIGINAL target indices (fast_apdgicp.hpp:102-105)
__global__ void corr_export_kernel(AlignBatch b, int slot, int s, int t, int* __restrict__ corr_out, float* __restrict__ sqd_out, double* __restrict__ m16_out) {
  const int ns = b.src.pt_off[s + 1] - b.src.pt_off[s];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  const size_t si = (size_t)slot * b.scratch.max_src + i;
  const unsigned orig = __float_as_uint(b.src.spts[b.src.pt_off[s] + i].w);
  const int c = b.scratch.corr[si];
  if (corr_out) corr_out[orig] = (c >= 0) ? (int)__float_as_uint(b.tgt.spts[b.tgt.pt_off[t] + c].w) : -1;
  if (sqd_out) sqd_out[orig] = b.scratch.sqd[si];
  if (m16_out) {
    double* o = m16_out + (size_t)orig * 16;
    for (int j = 0; j < 16; j++) o[j] = 0.0;
    if (c >= 0) {
      const double2 m0 = b.scratch.m0[si], m1 = b.scratch.m1[si], m2 = b.scratch.m2[si];
      o[0] = m0.x; o[1] = m0.y; o[2] = m1.x;
      o[4] = m0.y; o[5] = m1.y; o[6] = m2.x;
      o[8] = m1.x; o[9] = m2.x; o[10] = m2.y;
    }
  }
}

// ---- leaf build (apd_leaf.cuh): Hilbert order + one bounding box per 32 points, ONE CTA per cloud ----
// Replaces the three-level grid pyramid for clouds that fit shared memory: bounding box -> 30-bit Hilbert key per point ->
// stable LSD radix sort (4 x 8 bits, keys and 16-bit indices in shared memory) -> sorted points, inverse permutation and
// leaf boxes. Non-finite points get the key 2^30 and end up behind every finite point; they are left out of the boxes
// (pcl::KdTreeFLANN::convertCloudToArray leaves them out of the index). Equal keys keep their input order (the sort is
// stable), so the layout - and with it every later summation order - is deterministic.

__device__ __forceinline__ unsigned spread10(unsigned v) {  // bit b -> bit 3b, for 10-bit v
  v = (v | v << 16) & 0x030000FFu;
  v = (v | v << 8) & 0x0300F00Fu;
  v = (v | v << 4) & 0x030C30C3u;
  v = (v | v << 2) & 0x09249249u;
  return v;
}

// Skilling's transform: 3 x 10 bit coordinates -> 30-bit Hilbert index (consecutive indices are adjacent cells)
__device__ __forceinline__ unsigned hilbert30(unsigned x, unsigned y, unsigned z) {
  unsigned X[3] = {x, y, z};
#pragma unroll
  for (unsigned Q = 512u; Q > 1u; Q >>= 1) {
    const unsigned P = Q - 1u;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (X[i] & Q) X[0] ^= P;
      else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0];
  X[2] ^= X[1];
  unsigned t = 0u;
#pragma unroll
  for (unsigned Q = 512u; Q > 1u; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1u;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  // interleave: bit b of X[0] -> 3b + 2, of X[1] -> 3b + 1, of X[2] -> 3b (spread by magic masks: 9 operations per coordinate)
  return spread10(X[0]) << 2 | spread10(X[1]) << 1 | spread10(X[2]);
}

// bounding box of leaf l (finite points only), by one warp; the sorted order comes either as 64-bit words (index in the low 13 bits)
// or as 16-bit indices
__device__ __forceinline__ void leaf_box_of(const float4* __restrict__ pts, const unsigned long long* __restrict__ w64, const uint16_t* __restrict__ v16, int n, int l,
                                            int lane, float4* __restrict__ box) {
  const int i = l * kLeaf + lane;
  unsigned lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
  if (i < n) {
    const unsigned idx = w64 ? (unsigned)(w64[i] & ((1u << kLeafPosBits) - 1u)) : (unsigned)v16[i];
    const float4 p = pts[idx];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      lo[0] = hi[0] = enc_f(p.x); lo[1] = hi[1] = enc_f(p.y); lo[2] = hi[2] = enc_f(p.z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
    hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
  }
  if (lane == 0) {
    const float inf = __int_as_float(0x7f800000);
    const bool empty = lo[0] > hi[0];
    box[2 * l] = empty ? make_float4(inf, inf, inf, 0.f) : make_float4(dec_f(lo[0]), dec_f(lo[1]), dec_f(lo[2]), 0.f);
    box[2 * l + 1] = empty ? make_float4(-inf, -inf, -inf, 0.f) : make_float4(dec_f(hi[0]), dec_f(hi[1]), dec_f(hi[2]), 0.f);
  }
}

// profiling aid (option "timeline"): %globaltimer stamps of the build phases of cloud 0, appended to apd_get_timeline as phases 100+
__device__ unsigned long long g_leaf_build_stamps[16];
__device__ __forceinline__ void bstamp(bool on, int k) {
  if (on && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_leaf_build_stamps[k] = t;
  }
}

__global__ void __launch_bounds__(1024) leaf_build_kernel(CloudSetView cs, bool bitonic, bool stamps) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ unsigned s_hist[32][256];  // per-warp digit histograms / scatter cursors
  __shared__ unsigned s_box[32][6];
  __shared__ unsigned s_warp[33];
  __shared__ float s_lo[3], s_scale;
  const int c = blockIdx.x;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int base = cs.pt_off[c];
  const int n = cs.pt_off[c + 1] - base;
  if (n == 0) return;
  bstamp(stamps, 0);
  const float4* gpts = cs.pts + base;
  // the cloud itself is kept in shared memory for the whole build: the key pass, the final gather and the leaf boxes read it again
  // in sorted (random) order, which from global memory cost one exposed L2 round trip per point and thread (7 + 3 us of a 42 us build)
  float4* pts = reinterpret_cast<float4*>(sm_raw);
  unsigned* ka = reinterpret_cast<unsigned*>(pts + n);
  unsigned* kb = ka + n;
  uint16_t* va = reinterpret_cast<uint16_t*>(kb + n);
  uint16_t* vb = va + n;

  // 1. stage the points; bounding box of the finite ones
  unsigned mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  for (int i = tid; i < n; i += T) {
    const float4 p = gpts[i];
    pts[i] = p;
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
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { s_box[warp][a] = mn[a]; s_box[warp][3 + a] = mx[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    float ext = 0.f;
    for (int a = 0; a < 3; a++) {
      unsigned l = 0xFFFFFFFFu, h = 0u;
      for (int w = 0; w < (T >> 5); w++) { l = min(l, s_box[w][a]); h = max(h, s_box[w][3 + a]); }
      const float lo = l > h ? 0.f : dec_f(l), hi = l > h ? 0.f : dec_f(h);
      s_lo[a] = lo;
      ext = fmaxf(ext, hi - lo);
    }
    s_scale = ext > 0.f ? 1023.0f / ext : 0.f;  // cubic cells: the curve's locality is isotropic
  }
  __syncthreads();
  bstamp(stamps, 1);  // bounding box

  // 2. Hilbert keys
  const float lox = s_lo[0], loy = s_lo[1], loz = s_lo[2], scale = s_scale;
  for (int i = tid; i < n; i += T) {
    const float4 p = pts[i];
    unsigned key = 0x40000000u;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const unsigned qx = min(1023u, (unsigned)fmaxf((p.x - lox) * scale, 0.f));
      const unsigned qy = min(1023u, (unsigned)fmaxf((p.y - loy) * scale, 0.f));
      const unsigned qz = min(1023u, (unsigned)fmaxf((p.z - loz) * scale, 0.f));
      key = hilbert30(qx, qy, qz);
    }
    ka[i] = key;
    va[i] = (uint16_t)i;
  }
  __syncthreads();
  bstamp(stamps, 2);  // keys

  // 3a. few clouds (a single scan: latency matters, the GPU is otherwise idle): bitonic sort of 64-bit (key << 13 | index) words in
  //     shared memory - 91 barrier-separated steps of four compare-exchanges per thread for 8192 slots (~5 us) against ~44 us for
  //     the four radix passes below, whose histogram / scan / ranked-scatter phases are long dependent chains. Many more
  //     instructions in total, so batches keep the radix sort. Same order either way: the index breaks ties, as stability does.
  if (bitonic) {
    unsigned long long* w = reinterpret_cast<unsigned long long*>(ka);  // behind the staged points (8-byte aligned: 16 n bytes in front)
    int n2 = 64;
    while (n2 < n) n2 <<= 1;
    __syncthreads();
    // (keys were written to ka / va above, which alias w: read them back into registers first)
    unsigned long long mine[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = tid + u * T;
      mine[u] = i < n ? ((unsigned long long)ka[i] << kLeafPosBits) | (unsigned long long)va[i] : ~0ull;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = tid + u * T;
      if (i < n2) w[i] = mine[u];
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int p = tid; p < (n2 >> 1); p += T) {
          const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));   // lower element of the pair (bit j clear)
          const int l = i | j;
          const unsigned long long a = w[i], b = w[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { w[i] = b; w[l] = a; }
        }
        __syncthreads();
      }
    }
    float4* spts = cs.spts + base;
    for (int i = tid; i < n; i += T) {
      const unsigned idx = (unsigned)(w[i] & ((1u << kLeafPosBits) - 1u));
      const float4 p = pts[idx];
      spts[i] = make_float4(p.x, p.y, p.z, __uint_as_float(idx));
      cs.inv0[base + idx] = i;
    }
    const int nleaf_b = (n + kLeaf - 1) / kLeaf;
    float4* box_b = cs.lbox + 2 * (size_t)cs.leaf_off[c];
    for (int l = warp; l < nleaf_b; l += (T >> 5)) leaf_box_of(pts, w, nullptr, n, l, lane, box_b);
    if (cs.limg) {  // the shared-memory image (CloudSetView::limg)
      float4* img = cs.limg + (size_t)kLeafImage * cs.leaf_off[c];
      for (int l = warp; l < nleaf_b; l += (T >> 5)) {
        const int i = l * kLeaf + lane;
        const float qnan = __int_as_float(0x7fc00000);
        float4 p = make_float4(qnan, qnan, qnan, 0.f);
        unsigned idx = 0x7FFFFu;
        if (i < n) { idx = (unsigned)(w[i] & ((1u << kLeafPosBits) - 1u)); p = pts[idx]; }
        leaf_image_store_points(img, i, p.x, p.y, p.z, idx, lane);
        __syncwarp();
        if (lane < 2) img[(size_t)nleaf_b * kLeaf + 2 * l + lane] = box_b[2 * l + lane];   // written by lane 0 of this warp above
      }
    }
    return;
  }

  // 3b. stable LSD radix sort, 8 bits per pass: every warp owns a contiguous band of 32-element rows and a private histogram;
  //     bins are ranked digit-major / warp-minor, lanes of a row that share a digit are ordered by lane (__match_any_sync)
  const int rows = (n + 31) >> 5;
  const int band = (rows + (T >> 5) - 1) / (T >> 5);
  const int r0 = min(warp * band, rows), r1 = min(r0 + band, rows);
  for (int shift = 0; shift < 32; shift += 8) {
    for (int d = lane; d < 256; d += 32) s_hist[warp][d] = 0u;
    __syncwarp();
    for (int r = r0; r < r1; r++) {
      const int j = (r << 5) + lane;
      if (j < n) atomicAdd(&s_hist[warp][(ka[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    bstamp(stamps && shift == 0, 9);  // pass 0: histograms
    {  // exclusive scan of the 256 x 32 bins in (digit, warp) order: 8 consecutive bins per thread
      unsigned loc[8], sum = 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int e = tid * 8 + k;
        loc[k] = s_hist[e & 31][e >> 5];
        sum += loc[k];
      }
      unsigned incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      if (tid < 32) {
        const unsigned w = s_warp[tid];
        unsigned wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
          if (tid >= d) wi += t;
        }
        s_warp[tid] = wi - w;
      }
      __syncthreads();
      unsigned run = s_warp[warp] + incl - sum;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int e = tid * 8 + k;
        s_hist[e & 31][e >> 5] = run;
        run += loc[k];
      }
    }
    __syncthreads();
    bstamp(stamps && shift == 0, 10);  // pass 0: bins ranked
    for (int r = r0; r < r1; r++) {
      const int j = (r << 5) + lane;
      const bool valid = j < n;
      const unsigned k = valid ? ka[j] : 0u;
      const unsigned d = (k >> shift) & 255u;
      // lanes of this row with the same digit: eight ballots (one per digit bit). __match_any_sync resolves one distinct value per
      // step: ~1 us per row when the 32 digits of a row are all different (the low bytes of a Hilbert key), 5 of the 7 us of a pass.
      unsigned same = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const unsigned bit = (d >> b) & 1u;
        const unsigned v = __ballot_sync(0xFFFFFFFFu, bit);
        same &= bit ? v : ~v;
      }
      const unsigned rank = __popc(same & ((1u << lane) - 1u));
      if (valid) {
        const unsigned dst = s_hist[warp][d] + rank;
        kb[dst] = k;
        vb[dst] = va[j];
      }
      __syncwarp();
      if (valid && rank == 0u) s_hist[warp][d] += __popc(same);
      __syncwarp();
    }
    __syncthreads();
    unsigned* tk = ka; ka = kb; kb = tk;
    uint16_t* tv = va; va = vb; vb = tv;
    bstamp(stamps, 3 + (shift >> 3));  // radix pass done
  }

  // 4. sorted points, inverse permutation, leaf boxes (finite points only): a warp's 32 consecutive sorted positions ARE one leaf
  float4* spts = cs.spts + base;
  const int nleaf = (n + kLeaf - 1) / kLeaf;
  float4* box = cs.lbox + 2 * (size_t)cs.leaf_off[c];
  float4* img = cs.limg ? cs.limg + (size_t)kLeafImage * cs.leaf_off[c] : nullptr;  // the shared-memory image (CloudSetView::limg)
  for (int l = warp; l < nleaf; l += (T >> 5)) {
    const int i = l * kLeaf + lane;
    unsigned lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    const float qnan = __int_as_float(0x7fc00000);
    float4 p = make_float4(qnan, qnan, qnan, 0.f);
    unsigned idx = 0x7FFFFu;
    if (i < n) {
      idx = va[i];
      p = pts[idx];
      spts[i] = make_float4(p.x, p.y, p.z, __uint_as_float(idx));
      cs.inv0[base + idx] = i;
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        lo[0] = hi[0] = enc_f(p.x); lo[1] = hi[1] = enc_f(p.y); lo[2] = hi[2] = enc_f(p.z);
      }
    }
    if (img) leaf_image_store_points(img, i, p.x, p.y, p.z, idx, lane);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
      hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
    }
    if (lane == 0) {
      const float inf = __int_as_float(0x7f800000);
      const bool empty = lo[0] > hi[0];
      const float4 blo = empty ? make_float4(inf, inf, inf, 0.f) : make_float4(dec_f(lo[0]), dec_f(lo[1]), dec_f(lo[2]), 0.f);
      const float4 bhi = empty ? make_float4(-inf, -inf, -inf, 0.f) : make_float4(dec_f(hi[0]), dec_f(hi[1]), dec_f(hi[2]), 0.f);
      box[2 * l] = blo;
      box[2 * l + 1] = bhi;
      if (img) {
        img[(size_t)nleaf * kLeaf + 2 * l] = blo;
        img[(size_t)nleaf * kLeaf + 2 * l + 1] = bhi;
      }
    }
  }
  bstamp(stamps, 8);  // sorted points, inverse order and boxes written
}

}  // namespace

#define APD_LAUNCH_CHECK()                      \
  do {                                          \
    if (st) st->launches++;                     \
    cudaError_t e_ = cudaGetLastError();        \
    if (e_ != cudaSuccess) return e_;           \
  } while (0)

cudaError_t launch_grid_build(const CloudSetView& cs, const BuildWorkspace& ws, const int4* tiles, int n_tiles, const int* cell_cap, long long total_cells,
                              bool finest_level, cudaStream_t stream, LaunchStats* st) {
  if (cs.n_clouds == 0 || n_tiles == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(cs.cells, 0, sizeof(unsigned) * (size_t)total_cells, stream);
  if (e != cudaSuccess) return e;
  bbox_init_kernel<<<(cs.n_clouds * 6 + 255) / 256, 256, 0, stream>>>(ws.bbox, cs.n_clouds);
  APD_LAUNCH_CHECK();
  bbox_kernel<<<n_tiles, kTileThreads, 0, stream>>>(cs, tiles, ws.bbox);
  APD_LAUNCH_CHECK();
  grid_params_kernel<<<(cs.n_clouds + 127) / 128, 128, 0, stream>>>(cs, ws.bbox, cell_cap);
  APD_LAUNCH_CHECK();
  count_kernel<<<n_tiles, kTileThreads, 0, stream>>>(cs, tiles, ws.cellid);
  APD_LAUNCH_CHECK();
  scan_kernel<<<cs.n_clouds, 1024, 0, stream>>>(cs, ws.cursor);
  APD_LAUNCH_CHECK();
  scatter_kernel<<<n_tiles, kTileThreads, 0, stream>>>(cs, tiles, ws.cellid, ws.cursor);
  APD_LAUNCH_CHECK();
  // Only the finest level fixes the order in which the align kernel sums its partials; coarse levels
  // are searched through (d2, index) keys alone, so their in-cell order is irrelevant.
  if (!finest_level) return cudaSuccess;
  for (int c0 = 0; c0 < cs.n_clouds; c0 += 32768) {
    const int ny = min(cs.n_clouds - c0, 32768);
    // enough threads to cover the largest per-cloud table once for batches; grid-stride otherwise
    const long long avg_cells = total_cells / cs.n_clouds + 1;
    const long long want_bx = (avg_cells + 255) / 256;
    const int bx = (int)(want_bx < 4096 ? want_bx : 4096);
    cell_sort_kernel<<<dim3(bx, ny), 256, 0, stream>>>(cs, c0);
    APD_LAUNCH_CHECK();
  }
  inverse_order_kernel<<<(cs.total_points + 255) / 256, 256, 0, stream>>>(cs);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_grid_build_fused(const CloudSetView& cs, const int* const cap[1 + kCoarseLevels], int* const cellid[1 + kCoarseLevels],
                                    unsigned* const cursor[1 + kCoarseLevels], size_t smem_bytes, cudaStream_t stream, LaunchStats* st) {
  if (cs.n_clouds == 0) return cudaSuccess;
  FusedLevels L;
  for (int l = 0; l <= kCoarseLevels; l++) { L.cap[l] = cap[l]; L.cellid[l] = cellid[l]; L.cursor[l] = cursor[l]; }
  if (smem_bytes > 0) {  // every cloud's sorted points + cell table fit shared memory
    cudaError_t e = ensure_dynamic_smem(build_fused_smem_kernel, smem_bytes);
    if (e != cudaSuccess) return e;
    build_fused_smem_kernel<<<dim3(cs.n_clouds, 1 + kCoarseLevels), 1024, smem_bytes, stream>>>(cs, L);
    APD_LAUNCH_CHECK();
    return cudaSuccess;
  }
  build_fused_kernel<<<dim3(cs.n_clouds, 1 + kCoarseLevels), 1024, 0, stream>>>(cs, L);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t leaf_build_stamps(unsigned long long out[16]) { return cudaMemcpyFromSymbol(out, g_leaf_build_stamps, sizeof(unsigned long long) * 16); }

cudaError_t launch_leaf_build(const CloudSetView& cs, int max_n, cudaStream_t stream, LaunchStats* st, bool stamps) {
  if (cs.n_clouds == 0) return cudaSuccess;
  int n2 = 64;
  while (n2 < max_n) n2 <<= 1;
  // Measured on one 5000-point scan: the bitonic path takes 97 us against 44 us for the four radix passes (91 barrier-separated
  // steps of conflicting 64-bit shared-memory exchanges): kept for the record, not used.
  const bool bitonic = false;
  size_t smem = (size_t)max_n * (sizeof(float4) + 2 * sizeof(unsigned) + 2 * sizeof(uint16_t)) + 16;
  if (bitonic) smem = std::max(smem, (size_t)max_n * sizeof(float4) + (size_t)n2 * sizeof(unsigned long long) + 16);
  cudaError_t e = ensure_dynamic_smem(leaf_build_kernel, smem);
  if (e != cudaSuccess) return e;
  leaf_build_kernel<<<cs.n_clouds, 1024, smem, stream>>>(cs, bitonic, stamps);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_pack_points(const float* xyz, int stride_floats, long long n, float4* out, cudaStream_t stream, LaunchStats* st) {
  if (n == 0) return cudaSuccess;
  if (stride_floats == 8 && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0)
    pack_points_xyzi_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(reinterpret_cast<const float4*>(xyz), n, out);
  else
    pack_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(xyz, stride_floats, n, out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_transform_points(const float4* pts, int n, const float* T16, float* out, int out_stride_floats, cudaStream_t stream, LaunchStats* st) {
  if (n == 0) return cudaSuccess;
  if (out_stride_floats == 3 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    transform_points_packed_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, n, T16, out);
  else
    transform_points_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, n, T16, out, out_stride_floats);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_l2_flush(const void* buf, size_t bytes, float* sink, cudaStream_t stream) {
  l2_flush_read_kernel<<<148 * 8, 256, 0, stream>>>(static_cast<const float4*>(buf), bytes / 16, sink);
  return cudaGetLastError();
}

cudaError_t launch_iota_w(float4* pts, int n, cudaStream_t stream, LaunchStats* st) {
  if (n == 0) return cudaSuccess;
  iota_w_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, n);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_cov_export(const CloudSetView& cs, int cloud, double* out16, cudaStream_t stream, LaunchStats* st) {
  // n is read on the device; size the grid from the set total (clouds exported this way are single-cloud sets)
  const int n = cs.total_points;
  if (n == 0) return cudaSuccess;
  const long long want = ((long long)n * 8 + 1023) / 1024;   // 4 items per thread
  cov_export_kernel<<<(unsigned)(want < 148 * 32 ? want : 148 * 32), 256, 0, stream>>>(cs, cloud, out16);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_cov_import(const CloudSetView& cs, int cloud, const double* in16, cudaStream_t stream, LaunchStats* st) {
  const int n = cs.total_points;
  if (n == 0) return cudaSuccess;
  cov_import_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cs, cloud, in16);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_corr_export(const AlignBatch& b, int slot, int s, int t, int* corr_out, float* sqd_out, double* m16_out, cudaStream_t stream, LaunchStats* st) {
  const int n = b.scratch.max_src;  // >= the source size; the kernel bounds itself
  if (n == 0) return cudaSuccess;
  corr_export_kernel<<<(n + 255) / 256, 256, 0, stream>>>(b, slot, s, t, corr_out, sqd_out, m16_out);
  APD_LAUNCH_CHECK();
  return cudaSuccess;
}

}  // namespace apd
