// k-nearest-neighbour search fused with covariance estimation and regularisation:
// FastAPDGICP::calculate_covariances, fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:303-363.
//
// One thread per query point, queries taken in cell-sorted order so the lanes of a warp walk the
// same neighbourhood. For clouds that fit, the whole grid (sorted float4 points + a uint16 cell
// table) is staged once per CTA in shared memory and every candidate read is an LDS; larger clouds
// read the same structures from HBM/L2. The top-k list lives in registers as 64-bit
// (d2 bits << 32 | index) keys, which makes the (d2, index) order a single integer compare.
// The covariance follows the reference's arithmetic in the reference's order (fp64, one rounding
// per operation) so that the Jacobi sweep sees the same matrix as the CPU oracle.
#include <cstdint>

#include "apd_internal.h"

namespace apd {

namespace {

// fast_apdgicp_impl.hpp:326-359
__device__ __forceinline__ Sym3 regularize(const Sym3& cov, int method) {
  if (method == APD_REG_NONE) return cov;
  if (method == APD_REG_FROBENIUS) {
    Sym3 C = cov;
    C.xx = dadd(C.xx, 1e-3); C.yy = dadd(C.yy, 1e-3); C.zz = dadd(C.zz, 1e-3);
    Sym3 Ci = inverse(C);
    const double fro = sqrt(Ci.xx * Ci.xx + Ci.yy * Ci.yy + Ci.zz * Ci.zz + 2.0 * (Ci.xy * Ci.xy + Ci.xz * Ci.xz + Ci.yz * Ci.yz));
    Ci.xx /= fro; Ci.xy /= fro; Ci.xz /= fro; Ci.yy /= fro; Ci.yz /= fro; Ci.zz /= fro;
    return inverse(Ci);
  }
  double w[3], V[9], values[3];
  sym_eig3(cov, w, V);
  if (method == APD_REG_PLANE) {
    values[0] = 1.0; values[1] = 1.0; values[2] = 1e-3;
  } else if (method == APD_REG_MIN_EIG) {
    for (int i = 0; i < 3; i++) values[i] = fmax(w[i], 1e-3);
  } else {  // NORMALIZED_MIN_EIG
    const double mx = fmax(w[0], fmax(w[1], w[2]));
    for (int i = 0; i < 3; i++) values[i] = fmax(w[i] / mx, 1e-3);
  }
  return recompose(V, values);
}

// ---- cold paths, kept OUT OF LINE ----
// The kernel's hot loop (fine-grid candidate scan + packed insertion + exact sort + covariance + Jacobi) has
// to live in the 32 KB instruction cache; with everything inlined the kernel was 12 000 SASS instructions
// (190 KB) and ncu showed instruction fetch (no_instruction) as its top stall. The rare paths - coarse
// pyramid levels for isolated points, the exact 64-bit list for crowded distance buckets - are separate
// functions that hand their result back through local memory.
template <int K, int M>
__device__ __noinline__ void knn_coarse_packed(GridView<unsigned> G1, GridView<unsigned> G2, float qx, float qy, float qz, int kbits, int rings, unsigned* out) {
  TopKPacked<K, M> ap;
  ap.kbits = kbits;
  ap.sh = kbits - 1;
#pragma unroll 1
  for (int l = 0; l < 2; l++) {
    ap.init();
    if (grid_search(l ? G2 : G1, qx, qy, qz, __int_as_float(0x7f800000), ap, l ? 0x7fffffff : rings)) break;
  }
#pragma unroll
  for (int j = 0; j < M; j++) out[j] = ap.a[j];
}

template <int K, typename CellT>
__device__ __noinline__ void knn_exact(GridView<CellT> G0, GridView<unsigned> G1, GridView<unsigned> G2, float qx, float qy, float qz, int rings,
                                       unsigned long long* out) {
  TopK<K> tk;
  tk.init();
  if (!grid_search(G0, qx, qy, qz, __int_as_float(0x7f800000), tk, rings)) {
#pragma unroll 1
    for (int l = 0; l < 2; l++) {
      tk.init();
      if (grid_search(l ? G2 : G1, qx, qy, qz, __int_as_float(0x7f800000), tk, l ? 0x7fffffff : rings)) break;
    }
  }
#pragma unroll
  for (int j = 0; j < K; j++) out[j] = tk.key[j];
}

constexpr int kMaxRankedChunks = 1024;  // tiles with more chunks than this are not ranked (6 KB of shared memory)

template <int K, bool STAGED>
__global__ void __launch_bounds__(kKnnThreads, 1)
knn_cov_kernel(CloudSetView cs, const int4* __restrict__ tiles, int k, int method, int packed_path, int fine_rings, int idx_off, int* __restrict__ knn_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_next;  // next unclaimed chunk of the tile
  __shared__ unsigned s_span[kMaxRankedChunks];
  __shared__ uint16_t s_order[kMaxRankedChunks];
  const int4 tile = tiles[blockIdx.x];
  if (threadIdx.x == 0) s_next = 0;
  const int c = tile.x;
  const int base = cs.pt_off[c];
  const int n = cs.pt_off[c + 1] - base;
  const GridParams g = cs.grid[c];
  const float4* gspts = cs.spts + base;
  const unsigned* gcells = cs.cells + cs.cell_off[c];

  typedef typename std::conditional<STAGED, uint16_t, unsigned>::type CellT;  // also the type of a stored neighbour index
  GridView<CellT> G;
  G.g = g;
  G.n = n;
  // this thread's neighbour list (entry j at nbr[j * blockDim.x]: conflict-free), behind the staged grid
  CellT* nbr = reinterpret_cast<CellT*>(smem_raw + idx_off) + threadIdx.x;
  if (STAGED) {
    float4* s_pts = reinterpret_cast<float4*>(smem_raw);
    uint16_t* s_cells = reinterpret_cast<uint16_t*>(smem_raw + sizeof(float4) * (size_t)n);
    stage_grid(s_pts, s_cells, gspts, gcells, n, g.ncells);
    __syncthreads();
    G.spts = s_pts;
    G.cells = reinterpret_cast<const CellT*>(s_cells);
  } else {
    G.spts = gspts;
    G.cells = reinterpret_cast<const CellT*>(gcells);
    __syncthreads();
  }
  const int nstride = blockDim.x;

  const float4* opts = cs.pts + base;  // original order, for the neighbour gather
  const double inv_div = (double)k;
  // index bits of the packed key; beyond ~1M points the distance part becomes too coarse to filter
  int kbits = 1;
  while (kbits < 31 && (1u << kbits) < (unsigned)n) kbits++;
  const bool use_packed = packed_path && kbits <= 20;
  // Lanes of a warp take ADJACENT cell-sorted queries: they walk the same rings at the same time. Two
  // alternatives were measured in round 1 and rejected (profiles/): per-thread runs of consecutive
  // queries chained by the triangle inequality (2x slower: the warp loses its spatial coherence) and
  // per-ring queues drained by the whole warp (1.2x slower).
  // Warps pull chunks of `qpw` adjacent queries from a shared counter: the cost of a chunk varies a lot
  // (sparse neighbourhoods walk more rings), and with a static round-robin the CTA idled ~20 % of its
  // warp slots waiting for the slowest warp (ncu sm__warps_active 12.8 of 16). A tile smaller than the
  // CTA (single-pair latency mode) is spread thinly, few queries per warp, so that every scheduler of
  // the SM has warps and a warp serialises fewer divergent lanes.
  const int tile_end = tile.y + tile.z;
  const int lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  const int qpw = imax(1, imin(32, (tile.z + n_warps - 1) / n_warps));
  // Longest chunk first: a chunk of adjacent queries that spreads over many cells lies in a sparse region, and its
  // searches walk more rings (or fall back to the coarse levels). The chunks are ranked by that cell span and handed
  // out in descending order, so the long ones do not end up alone in the tail of the CTA (with the plain sorted order
  // the sparse high-z end came last: 12 % of the warp slots idle; from the end first: -7 %; ranked: see DESIGN.md).
  const int n_chunks = (tile.z + qpw - 1) / qpw;
  const bool ranked = n_chunks <= kMaxRankedChunks;
  if (ranked) {
    for (int ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
      const float4 a = G.spts[tile.y + ch * qpw], b = G.spts[imin(tile.y + ch * qpw + qpw, tile_end) - 1];
      s_span[ch] = (unsigned)(cell_index(g, b.x, b.y, b.z) - cell_index(g, a.x, a.y, a.z));
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < n_chunks; ch += blockDim.x) {
      const unsigned mine = s_span[ch];
      int rank = 0;
      for (int j = 0; j < n_chunks; j++) {
        const unsigned o = s_span[j];
        rank += (o > mine || (o == mine && j < ch)) ? 1 : 0;
      }
      s_order[rank] = (uint16_t)ch;
    }
    __syncthreads();
  }
  for (;;) {
    int k0 = 0;
    if (lane == 0) k0 = atomicAdd(&s_next, 1);
    k0 = __shfl_sync(0xFFFFFFFFu, k0, 0);
    if (k0 >= n_chunks) break;
    // unranked (very many chunks): from the end of the cell-sorted order, where the sparse high-z cells are
    const int chunk = ranked ? (int)s_order[k0] : n_chunks - 1 - k0;
    const int q = tile.y + chunk * qpw + lane;
    if (lane >= qpw || q >= tile_end) continue;
    const float4 p = G.spts[q];
    const unsigned self = __float_as_uint(p.w);
    // Fast path: collect candidates in a packed 32-bit list (min/max insertion, see TopKPacked), then
    // recompute and sort the exact keys of its K+4 entries. The list proves its own completeness;
    // when it cannot (a crowded distance bucket: lattice-like ties) the query is redone with the exact
    // 64-bit list. Either way the result is the exact (d2, index)-ordered top-k.
    // Both searches go through the grid pyramid: fine grid for a few rings, coarser levels for the
    // rare isolated point.
    bool done = false;
    if (use_packed) {
      TopKPacked<K, K + 4> ap;
      ap.kbits = kbits;
      ap.sh = kbits - 1;
      ap.init();
      if (!grid_search(G, p.x, p.y, p.z, __int_as_float(0x7f800000), ap, fine_rings)) {
        unsigned tmp[K + 4];
        knn_coarse_packed<K, K + 4>(coarse_view(cs, 0, c), coarse_view(cs, 1, c), p.x, p.y, p.z, kbits, fine_rings, tmp);
#pragma unroll
        for (int j = 0; j < K + 4; j++) ap.a[j] = tmp[j];
      }
      if (ap.complete()) {
        TopK<K> tk;
        exact_from_packed(ap, p.x, p.y, p.z, opts, tk);
#pragma unroll
        for (int j = 0; j < K; j++) nbr[j * nstride] = (CellT)(unsigned)(tk.key[j] & 0xFFFFFFFFull);
        done = true;
      }
    }
    if (!done) {
      unsigned long long fb[K];
      knn_exact<K, CellT>(G, coarse_view(cs, 0, c), coarse_view(cs, 1, c), p.x, p.y, p.z, fine_rings, fb);
#pragma unroll
      for (int j = 0; j < K; j++) nbr[j * nstride] = (CellT)(unsigned)(fb[j] & 0xFFFFFFFFull);
    }
    // (an entry that is not a valid index - all ones, truncated to CellT - marks "no neighbour": non-finite query)

    // neighbours -> mean -> covariance / k   (fast_apdgicp_impl.hpp:318-324); rolled loops over the stored list
    double mx = 0.0, my = 0.0, mz = 0.0;
#pragma unroll 4
    for (int j = 0; j < k; j++) {
      unsigned idx = (unsigned)nbr[j * nstride];
      if (idx >= (unsigned)n) idx = self;
      const float4 nb = opts[idx];
      mx = dadd(mx, (double)nb.x);
      my = dadd(my, (double)nb.y);
      mz = dadd(mz, (double)nb.z);
    }
    mx = mx / inv_div; my = my / inv_div; mz = mz / inv_div;
    Sym3 cov{0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int j = 0; j < k; j++) {
      unsigned idx = (unsigned)nbr[j * nstride];
      if (idx >= (unsigned)n) idx = self;
      const float4 nb = opts[idx];
      const double dx = dsub((double)nb.x, mx), dy = dsub((double)nb.y, my), dz = dsub((double)nb.z, mz);
      cov.xx = dadd(cov.xx, dmul(dx, dx));
      cov.xy = dadd(cov.xy, dmul(dx, dy));
      cov.xz = dadd(cov.xz, dmul(dx, dz));
      cov.yy = dadd(cov.yy, dmul(dy, dy));
      cov.yz = dadd(cov.yz, dmul(dy, dz));
      cov.zz = dadd(cov.zz, dmul(dz, dz));
    }
    cov.xx /= inv_div; cov.xy /= inv_div; cov.xz /= inv_div; cov.yy /= inv_div; cov.yz /= inv_div; cov.zz /= inv_div;
    const Sym3 r = regularize(cov, method);
    cs.cov0[base + q] = make_double2(r.xx, r.xy);
    cs.cov1[base + q] = make_double2(r.xz, r.yy);
    cs.cov2[base + q] = make_double2(r.yz, r.zz);
    if (knn_out) {
      int* row = knn_out + ((size_t)base + self) * k;
#pragma unroll 1
      for (int j = 0; j < k; j++) {
        const unsigned idx = (unsigned)nbr[j * nstride];
        row[j] = idx < (unsigned)n ? (int)idx : -1;
      }
    }
  }
}

template <int K>
cudaError_t launch_k(const CloudSetView& cs, const int4* tiles, int n_tiles, bool staged, size_t smem_bytes, const DeviceParams& prm, int* knn_out,
                     cudaStream_t stream) {
  // dynamic shared memory: [staged grid | per-thread neighbour lists: K entries of 2 (staged) or 4 bytes]
  if (staged) {
    const size_t idx_off = (smem_bytes + 15) & ~(size_t)15;
    const size_t total = idx_off + sizeof(uint16_t) * K * kKnnThreads;
    cudaError_t e = cudaFuncSetAttribute(knn_cov_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
    if (e != cudaSuccess) return e;
    knn_cov_kernel<K, true><<<n_tiles, kKnnThreads, total, stream>>>(cs, tiles, prm.k, prm.regularization, prm.knn_packed, prm.knn_fine_rings, (int)idx_off, knn_out);
  } else {
    const size_t total = sizeof(unsigned) * K * kKnnThreads;
    cudaError_t e = cudaFuncSetAttribute(knn_cov_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
    if (e != cudaSuccess) return e;
    knn_cov_kernel<K, false><<<n_tiles, kKnnThreads, total, stream>>>(cs, tiles, prm.k, prm.regularization, prm.knn_packed, prm.knn_fine_rings, 0, knn_out);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_knn_cov(const CloudSetView& cs, const int4* tiles, int n_tiles, bool staged, size_t smem_bytes, const DeviceParams& prm, int* knn_out,
                           cudaStream_t stream, LaunchStats* st) {
  if (n_tiles == 0) return cudaSuccess;
  if (st) st->launches++;
  const int k = prm.k;
  if (k <= 8) return launch_k<8>(cs, tiles, n_tiles, staged, smem_bytes, prm, knn_out, stream);
  if (k <= 10) return launch_k<10>(cs, tiles, n_tiles, staged, smem_bytes, prm, knn_out, stream);
  if (k <= 15) return launch_k<15>(cs, tiles, n_tiles, staged, smem_bytes, prm, knn_out, stream);
  if (k <= 20) return launch_k<20>(cs, tiles, n_tiles, staged, smem_bytes, prm, knn_out, stream);
  if (k <= 32) return launch_k<32>(cs, tiles, n_tiles, staged, smem_bytes, prm, knn_out, stream);
  return cudaErrorInvalidValue;
}

}  // namespace apd
