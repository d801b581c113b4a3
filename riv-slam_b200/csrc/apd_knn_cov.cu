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
#include "apd_leaf.cuh"
#include "apd_merge_net.cuh"

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

// ======================================================================================================================
// Leaf-mode kNN + covariance (clouds that fit shared memory; apd_leaf.cuh). A warp owns the 32 queries of one leaf.
//
// Candidate bookkeeping. Round 1 inserted every candidate that passed the gate into a sorted register list at once: a
// 48-instruction min/max network executed by the 7 lanes (of 32) that happened to pass. Here a lane that passes only
// APPENDS a packed 32-bit key ((top 19 bits of d2) << 13 | slot) to its own pending list in shared memory (one
// predicated store); the lists are merged into the sorted register list by ALL lanes together at warp-uniform points
// (before a leaf scan that could overflow a list, and at the end), so the network runs converged. The gate only tightens
// at a merge; a stale gate costs a few extra appends, never correctness.
// The packed order is a filter (quantised distance), exactly as in TopKPacked: the list keeps K + 4 entries, proves its own
// completeness (bucket of entry K+3 strictly beyond the bucket of entry K-1), then the exact 64-bit keys
// (d2 bits, original index) of its entries are recomputed from the staged points and sorted; a crowded bucket (lattice
// data, duplicates) falls back to a bounded exact scan. Either way the result is the exact (d2, index)-ordered top-k.
#ifndef APD_KNN_LEAF_THREADS
#define APD_KNN_LEAF_THREADS 640
#endif
constexpr int kKnnLeafThreads = APD_KNN_LEAF_THREADS;
constexpr int kKnnTransposeMax = 10;  // a leaf that at most this many of the warp's queries can reach is scanned transposed (one turn per query)
#ifndef APD_PEND_CAP
#define APD_PEND_CAP 40
#endif
constexpr int kPendCap = APD_PEND_CAP;  // pending keys per lane; a leaf scan appends at most 32

template <int K, int M>
struct LeafTopK {
  unsigned a[M];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < M; i++) a[i] = 0xFFFFFFFFu;
  }
  __device__ __forceinline__ void insert(unsigned key) {
#pragma unroll
    for (int j = M - 1; j > 0; j--) a[j] = umin32(a[j], umax32(a[j - 1], key));
    a[0] = umin32(a[0], key);
  }
  // upper edge of the K-th entry's distance bucket; an empty slot (all ones) decodes to a NaN = "no bound yet"
  __device__ __forceinline__ float bound2() const {
    constexpr int sh = kLeafPosBits - 1;
    return __uint_as_float(((a[K - 1] >> kLeafPosBits) << sh) | ((1u << sh) - 1u));
  }
  __device__ __forceinline__ bool complete() const { return a[K - 1] == 0xFFFFFFFFu || (a[M - 1] >> kLeafPosBits) > (a[K - 1] >> kLeafPosBits); }
};

__device__ __forceinline__ unsigned leaf_pack_key(float d2, int pos) { return ((__float_as_uint(d2) >> (kLeafPosBits - 1)) << kLeafPosBits) | (unsigned)pos; }

// exact 64-bit key of a staged slot: (d2 bits) << 32 | tag, tag = original index << 13 | slot (apd_leaf.cuh)
__device__ __forceinline__ unsigned long long leaf_exact_key(const LeafView& L, int pos, float qx, float qy, float qz) {
  const float4 t = leaf_point_gather(L, pos);
  const float d2 = sqdist_rn(qx, qy, qz, t.x, t.y, t.z);
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(t.w);
}

// Rare path (crowded distance bucket): exact top-K of one lane by a scan of every staged point inside `bound2`.
template <int K>
__device__ __noinline__ void knn_leaf_exact(LeafView L, float qx, float qy, float qz, float bound2, unsigned long long* out) {
  TopK<K> tk;
  tk.init();
  const int npad = L.nleaf * kLeaf;
  for (int p = 0; p < npad; p++) {
    const float4 t = leaf_point(L, p);
    const float d2 = sqdist_rn(qx, qy, qz, t.x, t.y, t.z);
    if (d2 <= bound2) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(t.w);
      if (key < tk.key[K - 1]) {
        tk.key[K - 1] = key;
#pragma unroll
        for (int j = K - 1; j > 0; j--) {
          const unsigned long long x = tk.key[j - 1], y = tk.key[j];
          const bool sw = y < x;
          tk.key[j - 1] = sw ? y : x;
          tk.key[j] = sw ? x : y;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < K; j++) out[j] = tk.key[j];
}

// profiling aid (option "timeline"): %globaltimer stamps of the phases of the first group of CTA 0, appended to apd_get_timeline as phases 120+
__device__ unsigned long long g_knn_stamps[16];
__device__ __forceinline__ void kstamp(bool on, int k) {
  if (on) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_knn_stamps[k] = t;
  }
}

template <int K>
__global__ void __launch_bounds__(kKnnLeafThreads, 1)
knn_cov_leaf_kernel(CloudSetView cs, const int4* __restrict__ tiles, int k, int method, int* __restrict__ knn_out, unsigned long long* __restrict__ evals, bool stamps) {
  constexpr int M = K + 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_next;
  __shared__ float4 s_q[kKnnLeafThreads];    // transposed scans: every lane's query (x, y, z, gate) ...
  __shared__ unsigned s_qc[kKnnLeafThreads];  // ... and the fill of its pending list
  const int4 tile = tiles[blockIdx.x];  // (cloud, first leaf, leaves, parts per leaf)
  if (threadIdx.x == 0) s_next = 0;
  const int c = tile.x;
  const int base = cs.pt_off[c];
  const int n = cs.pt_off[c + 1] - base;
  kstamp(stamps && blockIdx.x == 0 && threadIdx.x == 0, 0);
  LeafView L;
  L.n = n;
  L.nleaf = (n + kLeaf - 1) / kLeaf;
  float4* sP = reinterpret_cast<float4*>(smem_raw);
  float4* sbox = sP + (size_t)L.nleaf * kLeaf;
  unsigned* s_list = reinterpret_cast<unsigned*>(sbox + 2 * (size_t)L.nleaf);
  __shared__ unsigned long long s_stage_bar;
  if (cs.limg) {
    // one bulk asynchronous copy of the image the build wrote (points in the pair layout + boxes), completion on an mbarrier
    unsigned parity = 0u;
    if (threadIdx.x == 0) mbar_init(&s_stage_bar, 1u);
    __syncthreads();
    leaf_stage_bulk(sP, cs.limg + (size_t)kLeafImage * cs.leaf_off[c], n, &s_stage_bar, parity);
  } else {
    leaf_stage(sP, sbox, cs.spts + base, cs.lbox + 2 * (size_t)cs.leaf_off[c], n);
  }
  L.P = sP;
  L.box = sbox;
  __syncthreads();
  kstamp(stamps && blockIdx.x == 0 && threadIdx.x == 0, 1);  // staged

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NT = 32;  // every warp owns its own region: entry j of lane l at [j * 32 + l] (conflict-free)
  unsigned* lst = s_list + warp * 32 * kPendCap + lane;
  // after the search the warp's region holds the K neighbour slots of its lanes (16-bit entries, entry j at nbr[j * 32])
  uint16_t* nbr = reinterpret_cast<uint16_t*>(s_list + warp * 32 * kPendCap) + lane;
  const double inv_div = (double)k;

  for (;;) {
    int k0 = 0;
    if (lane == 0) k0 = atomicAdd(&s_next, 1);
    k0 = __shfl_sync(0xFFFFFFFFu, k0, 0);
    // A launch with fewer leaves than the GPU has warps (one scan: 157 leaves for 2960 warps) cuts every leaf's 32 queries into
    // `sub` parts of 32 / sub lanes, one warp each. The instruction count of a search does not depend on how many lanes take part,
    // but the union of leaves a warp must scan does: the longest group of a radar scan (32 clutter points spread over tens of
    // metres) took 76 us against 24 us for a typical one, and the kernel lasts as long as its longest warp.
    const int sub = tile.w > 1 ? tile.w : 1, qpw = kLeaf / sub;
    if (k0 >= tile.z * sub) break;
    const int g = tile.y + (tile.z - 1 - k0 / sub);  // from the end of the curve first (no particular reason beyond determinism)
    const bool st = stamps && blockIdx.x == 0 && k0 == 0 && lane == 0;
    unsigned long long t_group = 0;
    if (stamps && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_group));
    const bool mine = lane < qpw;           // lanes beyond the part's queries idle along (they repeat its first query, invalid)
    const int q = g * kLeaf + (k0 % sub) * qpw + (mine ? lane : 0);
    const float4 p = leaf_point(L, q);      // padding slots: NaN
    const bool valid = mine && q < n && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
    const f32x2_t qx2 = f2_pack(p.x, p.x), qy2 = f2_pack(p.y, p.y), qz2 = f2_pack(p.z, p.z);

    LeafTopK<K, M> tk;
    tk.init();
    float gate = valid ? FLT_MAX : -1.f;   // a distance above it cannot belong to the result
    int cnt = 0;

    // Merge the pending keys of every lane into its sorted list, all lanes together. Batches of 8 go through a sorting network
    // and a pruned Batcher merge (apd_merge_net.cuh: 158 min/max per 8 keys); a remainder of one or two keys is inserted one at
    // a time (48 min/max each). A lane with fewer pending keys than the busiest lane merges empty slots (all ones).
    auto merge = [&]() {
      const int maxc = __reduce_max_sync(0xFFFFFFFFu, cnt);
      int j = 0;
      for (; maxc - j >= 3; j += 8) {
        unsigned b[8];
#pragma unroll
        for (int t = 0; t < 8; t++) b[t] = (j + t < cnt) ? lst[(j + t) * NT] : 0xFFFFFFFFu;
        sort8_net(b);
        MergeNet8<M>::run(tk.a, b);
      }
      for (; j < maxc; j++) tk.insert(j < cnt ? lst[j * NT] : 0xFFFFFFFFu);
      cnt = 0;
      if (valid) gate = fminf(tk.bound2(), FLT_MAX);  // NaN (list not full) -> FLT_MAX
    };
    auto scan = [&](int leaf) {
      const ulonglong2* P = reinterpret_cast<const ulonglong2*>(L.P) + leaf * kLeaf;
#pragma unroll 2
      for (int j = 0; j < kLeaf / 2; j += 2) {   // four candidates per vote
        const ulonglong2 A0 = P[2 * j], B0 = P[2 * j + 1], A1 = P[2 * j + 2], B1 = P[2 * j + 3];
        float d0, d1, d2, d3;
        leaf_pair_d2(qx2, qy2, qz2, A0, B0.x, d0, d1);
        leaf_pair_d2(qx2, qy2, qz2, A1, B1.x, d2, d3);
        // NaN (padding, non-finite points) fails every test (fminf drops it)
        if (__any_sync(0xFFFFFFFFu, fminf(fminf(d0, d1), fminf(d2, d3)) <= gate)) {
          if (d0 <= gate) { lst[cnt * NT] = leaf_pack_key(d0, leaf * kLeaf + 2 * j); cnt++; }
          if (d1 <= gate) { lst[cnt * NT] = leaf_pack_key(d1, leaf * kLeaf + 2 * j + 1); cnt++; }
          if (d2 <= gate) { lst[cnt * NT] = leaf_pack_key(d2, leaf * kLeaf + 2 * j + 2); cnt++; }
          if (d3 <= gate) { lst[cnt * NT] = leaf_pack_key(d3, leaf * kLeaf + 2 * j + 3); cnt++; }
        }
      }
    };

    // The same leaf TRANSPOSED, for a leaf only a few of the warp's queries can reach (bit mask `need`): every lane holds one
    // CANDIDATE, the queries take turns (two per trip: independent chains). A turn costs ~20 instructions against ~250 for the
    // broadcast scan, and only the queries that need the leaf pay. The lanes that pass a query's gate append to THAT query's
    // pending list (any lane may write any list: shared memory), at consecutive slots by their rank in the vote.
    float4* qs = s_q + warp * 32;
    unsigned* qc = s_qc + warp * 32;
    unsigned* lst_w = s_list + warp * 32 * kPendCap;
    auto scan_transposed = [&](int leaf, unsigned need) {
      const int cpos = leaf * kLeaf + lane;
      const float4 c = leaf_point(L, cpos);   // NaN coordinates in padding slots and for non-finite points: never pass
      qs[lane] = make_float4(p.x, p.y, p.z, gate);
      qc[lane] = (unsigned)cnt;
      __syncwarp();
      const unsigned lt = (1u << lane) - 1u;
      while (need) {
        const int q0 = __ffs(need) - 1;
        need &= need - 1;
        const bool two = need != 0u;
        const int q1 = two ? __ffs(need) - 1 : q0;
        need &= need - 1;
        const float4 Q0 = qs[q0], Q1 = qs[q1];
        const float d0 = sqdist_rn(Q0.x, Q0.y, Q0.z, c.x, c.y, c.z);
        const float d1 = sqdist_rn(Q1.x, Q1.y, Q1.z, c.x, c.y, c.z);
        const bool p0 = d0 <= Q0.w, p1 = two && d1 <= Q1.w;
        const unsigned m0 = __ballot_sync(0xFFFFFFFFu, p0), m1 = __ballot_sync(0xFFFFFFFFu, p1);
        if (m0) {
          const unsigned c0 = qc[q0];
          __syncwarp();  // every lane holds the fill before the owner advances it (write after read)
          if (p0) lst_w[(c0 + __popc(m0 & lt)) * NT + q0] = leaf_pack_key(d0, cpos);
          if (lane == q0) qc[q0] = c0 + __popc(m0);
        }
        if (m1) {
          const unsigned c1 = qc[q1];
          __syncwarp();
          if (p1) lst_w[(c1 + __popc(m1 & lt)) * NT + q1] = leaf_pack_key(d1, cpos);
          if (lane == q1) qc[q1] = c1 + __popc(m1);
        }
      }
      __syncwarp();
      cnt = (int)qc[lane];
    };

    if (__any_sync(0xFFFFFFFFu, valid)) {
      // The query's own leaf first (it fills the list: 32 points >= K, and gives every lane a first bound), then the other
      // leaves nearest first. ONE broadcast scan site and ONE merge site: the two are the bulk of the loop's instruction footprint.
      LeafSchedule S;
      int l = g;
      bool first = true;
      unsigned n_evals = 0, need = 0xFFFFFFFFu;
      for (;;) {
        if (l >= 0) {
          if (__popc(need) > kKnnTransposeMax) { scan(l); n_evals += kLeaf * 32; }
          else { n_evals += __popc(need) * 32; scan_transposed(l, need); }
        }
        // merge when the next scan (up to 32 appends) could overflow a list, after the own leaf, and at the very end
        if (first || __any_sync(0xFFFFFFFFu, l < 0 ? cnt > 0 : cnt > kPendCap - kLeaf)) merge();
        if (l < 0) break;
        if (first) kstamp(st, 2);  // own leaf scanned and merged
        if (first) {
          float glo[3], ghi[3];
          leaf_group_box(p.x, p.y, p.z, valid, glo, ghi);
          S.init(L, glo, ghi, g);
          first = false;
        }
        for (;;) {  // the nearest unvisited leaf that some lane's bound reaches
          const float G = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, valid ? __float_as_uint(gate) : 0u));
          l = S.next(G);
          if (l < 0) break;
          const float dl = leaf_point_box2(p.x, p.y, p.z, L.box[2 * l], L.box[2 * l + 1]);
          need = __ballot_sync(0xFFFFFFFFu, valid && dl <= gate);
          if (need) break;
        }
      }
      // distance evaluations executed by this group (bench.py's figure): 32 lanes x 32 candidates per broadcast scan, 32 per turn
      if (evals && lane == 0) atomicAdd(evals, (unsigned long long)n_evals);
    }

    kstamp(st, 3);  // search done
    // exact (d2, original index) keys of the K + 4 survivors, sorted; or the bounded exact scan when the packed list cannot
    // prove that it holds the whole K-th bucket
    unsigned long long k64[M];
    if (tk.complete()) {
#pragma unroll
      for (int j = 0; j < M; j++) k64[j] = tk.a[j] != 0xFFFFFFFFu ? leaf_exact_key(L, (int)(tk.a[j] & ((1u << kLeafPosBits) - 1u)), p.x, p.y, p.z) : APD_KEY_INF;
      bool swapped = true;
      while (swapped) {
        swapped = false;
#pragma unroll
        for (int j = 0; j + 1 < M; j += 2) {
          const unsigned long long x = k64[j], y = k64[j + 1];
          const bool sw = y < x;
          k64[j] = sw ? y : x; k64[j + 1] = sw ? x : y;
          swapped |= sw;
        }
#pragma unroll
        for (int j = 1; j + 1 < M; j += 2) {
          const unsigned long long x = k64[j], y = k64[j + 1];
          const bool sw = y < x;
          k64[j] = sw ? y : x; k64[j + 1] = sw ? x : y;
          swapped |= sw;
        }
      }
    } else {
      unsigned long long fb[K];
      knn_leaf_exact<K>(L, p.x, p.y, p.z, fminf(tk.bound2(), FLT_MAX), fb);
#pragma unroll
      for (int j = 0; j < K; j++) k64[j] = fb[j];
    }
    __syncwarp();  // the pending lists are dead from here on: their memory now holds the neighbour slots
    kstamp(st, 4);  // exact keys sorted
#pragma unroll
    for (int j = 0; j < K; j++) nbr[j * NT] = (k64[j] >> 32) >= 0x7F800000ull ? (uint16_t)0xFFFFu : (uint16_t)(k64[j] & ((1u << kLeafPosBits) - 1u));
    // (0xFFFF marks "no neighbour": a non-finite query, or fewer than k finite points in the cloud)

    if (mine && q < n) {
      // neighbours -> mean -> covariance / k   (fast_apdgicp_impl.hpp:318-324), in (d2, index) order from shared memory
      double mx = 0.0, my = 0.0, mz = 0.0;
#pragma unroll 4
      for (int j = 0; j < k; j++) {
        int s = nbr[j * NT];
        if (s == 0xFFFF) s = q;
        const float4 nb = leaf_point_gather(L, s);
        mx = dadd(mx, (double)nb.x);
        my = dadd(my, (double)nb.y);
        mz = dadd(mz, (double)nb.z);
      }
      mx = mx / inv_div; my = my / inv_div; mz = mz / inv_div;
      Sym3 cov{0, 0, 0, 0, 0, 0};
#pragma unroll 4
      for (int j = 0; j < k; j++) {
        int s = nbr[j * NT];
        if (s == 0xFFFF) s = q;
        const float4 nb = leaf_point_gather(L, s);
        const double dx = dsub((double)nb.x, mx), dy = dsub((double)nb.y, my), dz = dsub((double)nb.z, mz);
        cov.xx = dadd(cov.xx, dmul(dx, dx));
        cov.xy = dadd(cov.xy, dmul(dx, dy));
        cov.xz = dadd(cov.xz, dmul(dx, dz));
        cov.yy = dadd(cov.yy, dmul(dy, dy));
        cov.yz = dadd(cov.yz, dmul(dy, dz));
        cov.zz = dadd(cov.zz, dmul(dz, dz));
      }
      cov.xx /= inv_div; cov.xy /= inv_div; cov.xz /= inv_div; cov.yy /= inv_div; cov.yz /= inv_div; cov.zz /= inv_div;
      kstamp(st, 5);  // covariance
      const Sym3 r = regularize(cov, method);
      kstamp(st, 6);  // regularised
      cs.cov0[base + q] = make_double2(r.xx, r.xy);
      cs.cov1[base + q] = make_double2(r.xz, r.yy);
      cs.cov2[base + q] = make_double2(r.yz, r.zz);
      if (knn_out) {
        int* row = knn_out + ((size_t)base + leaf_tag_index(p.w)) * k;
#pragma unroll 1
        for (int j = 0; j < k; j++) {
          const int s = nbr[j * NT];
          row[j] = s == 0xFFFF ? -1 : (int)leaf_tag_index(leaf_point_gather(L, s).w);
        }
      }
    }
    __syncwarp();  // the next group's pending lists reuse the neighbour slots
    kstamp(st, 7);  // group done
    if (stamps && lane == 0) {  // latest finish and longest group of the whole launch
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMax(&g_knn_stamps[8], t);
      atomicMax(&g_knn_stamps[9], t - t_group);
    }
  }
}

template <int K>
cudaError_t launch_k_leaf(const CloudSetView& cs, const int4* tiles, int n_tiles, int max_n, const DeviceParams& prm, int* knn_out, unsigned long long* evals,
                          cudaStream_t stream, bool stamps) {
  const int nleaf = (max_n + kLeaf - 1) / kLeaf;
  const size_t total = (size_t)nleaf * (kLeaf * 16 + 32) + sizeof(unsigned) * kPendCap * kKnnLeafThreads;
  cudaError_t e = ensure_dynamic_smem(knn_cov_leaf_kernel<K>, total);
  if (e != cudaSuccess) return e;
  knn_cov_leaf_kernel<K><<<n_tiles, kKnnLeafThreads, total, stream>>>(cs, tiles, prm.k, prm.regularization, knn_out, evals, stamps);
  return cudaGetLastError();
}

template <int K>
cudaError_t launch_k(const CloudSetView& cs, const int4* tiles, int n_tiles, bool staged, size_t smem_bytes, const DeviceParams& prm, int* knn_out,
                     cudaStream_t stream) {
  // dynamic shared memory: [staged grid | per-thread neighbour lists: K entries of 2 (staged) or 4 bytes]
  if (staged) {
    const size_t idx_off = (smem_bytes + 15) & ~(size_t)15;
    const size_t total = idx_off + sizeof(uint16_t) * K * kKnnThreads;
    cudaError_t e = ensure_dynamic_smem(knn_cov_kernel<K, true>, total);
    if (e != cudaSuccess) return e;
    knn_cov_kernel<K, true><<<n_tiles, kKnnThreads, total, stream>>>(cs, tiles, prm.k, prm.regularization, prm.knn_packed, prm.knn_fine_rings, (int)idx_off, knn_out);
  } else {
    const size_t total = sizeof(unsigned) * K * kKnnThreads;
    cudaError_t e = ensure_dynamic_smem(knn_cov_kernel<K, false>, total);
    if (e != cudaSuccess) return e;
    knn_cov_kernel<K, false><<<n_tiles, kKnnThreads, total, stream>>>(cs, tiles, prm.k, prm.regularization, prm.knn_packed, prm.knn_fine_rings, 0, knn_out);
  }
  return cudaGetLastError();
}

}  // namespace

size_t knn_leaf_smem_bytes(int max_n) { return (size_t)((max_n + kLeaf - 1) / kLeaf) * (kLeaf * 16 + 32) + sizeof(unsigned) * kPendCap * kKnnLeafThreads; }

cudaError_t knn_leaf_stamps(unsigned long long out[16]) {  // read and reset
  cudaError_t e = cudaMemcpyFromSymbol(out, g_knn_stamps, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return e;
  const unsigned long long zero[16] = {0};
  return cudaMemcpyToSymbol(g_knn_stamps, zero, sizeof(zero));
}

cudaError_t launch_knn_cov_leaf(const CloudSetView& cs, const int4* tiles, int n_tiles, int max_n, const DeviceParams& prm, int* knn_out,
                                unsigned long long* evals, cudaStream_t stream, LaunchStats* st, bool stamps) {
  if (n_tiles == 0) return cudaSuccess;
  if (st) st->launches++;
  const int k = prm.k;
  if (k <= 8) return launch_k_leaf<8>(cs, tiles, n_tiles, max_n, prm, knn_out, evals, stream, stamps);
  if (k <= 10) return launch_k_leaf<10>(cs, tiles, n_tiles, max_n, prm, knn_out, evals, stream, stamps);
  if (k <= 15) return launch_k_leaf<15>(cs, tiles, n_tiles, max_n, prm, knn_out, evals, stream, stamps);
  if (k <= 20) return launch_k_leaf<20>(cs, tiles, n_tiles, max_n, prm, knn_out, evals, stream, stamps);
  if (k <= 32) return launch_k_leaf<32>(cs, tiles, n_tiles, max_n, prm, knn_out, evals, stream, stamps);
  return cudaErrorInvalidValue;
}

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
