// Uniform voxel grid over one cloud and the exact ring-expansion neighbour search on it.
// Replaces pcl::search::KdTree::nearestKSearch (FLANN, exact, sorted) as the reference uses it at
// fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:151 (1-NN per iteration) and :316
// (k-NN for covariances). Result order is ascending (d2, original index); d2 is the fp32
// L2_Simple value with no FMA contraction, so index sets are bit-exact against the CPU oracle.
//
// Layout: points are sorted by linear cell id (x fastest), so the cells x0..x1 of one (y,z) row are
// ONE contiguous run of points: a ring of the search is visited as (2r+1)^2 row lookups, not
// (2r+1)^3 cell lookups. spts[i].w carries the point's original index (bit pattern).
#pragma once
#include "apd_math.cuh"

namespace apd {

struct GridParams {
  float lox, loy, loz;  // lower corner
  float h, inv_h;       // cell edge
  int nx, ny, nz;
  int ncells;
  float slack;          // absolute safety margin for cell-boundary bounds (covers fp32 rounding of cell assignment)
};

// A cloud as the kernels see it. CellT is uint32_t in global memory and uint16_t when the table is
// staged in shared memory for clouds below 65536 points.
template <typename CellT>
struct GridView {
  const float4* spts;   // sorted points of this cloud
  const CellT* cells;   // ncells+1 exclusive prefix sums (local indices)
  GridParams g;
  int n;
};

APD_HD int imax(int a, int b) { return a > b ? a : b; }
APD_HD int imin(int a, int b) { return a < b ? a : b; }

APD_HD int cell_coord(float p, float lo, float inv_h, int n) {
  const float f = floorf(fmul(fsub(p, lo), inv_h));
  // NaN and -inf -> 0, +inf -> n-1
  return (f >= 0.0f) ? ((f < (float)(n - 1)) ? (int)f : (n - 1)) : 0;
}

APD_HD int cell_index(const GridParams& g, float x, float y, float z) {
  const int cx = cell_coord(x, g.lox, g.inv_h, g.nx);
  const int cy = cell_coord(y, g.loy, g.inv_h, g.ny);
  const int cz = cell_coord(z, g.loz, g.inv_h, g.nz);
  return (cz * g.ny + cy) * g.nx + cx;
}

// ---- visitors ----
// Keys are (float bits of d2) << 32 | original index: d2 >= +0, so unsigned integer order on the key
// IS the (d2, index) lexicographic order.
APD_HD unsigned long long make_key(float d2, unsigned idx) { return ((unsigned long long)f2u(d2) << 32) | idx; }
#define APD_KEY_INF 0x7F800000FFFFFFFFull

template <int K>
struct TopK {
  unsigned long long key[K];
  APD_HD void init() {
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = APD_KEY_INF;
  }
  APD_HD float bound2() const { return u2f((unsigned)(key[K - 1] >> 32)); }
  APD_HD void end_run() {}
  APD_HD void offer(float d2, unsigned idx, int /*pos*/) {
    const unsigned long long k = make_key(d2, idx);
    if (k < key[K - 1]) {
      key[K - 1] = k;
#pragma unroll
      for (int j = K - 1; j > 0; j--) {
        const unsigned long long a = key[j - 1], b = key[j];
        const bool sw = b < a;
        key[j - 1] = sw ? b : a;
        key[j] = sw ? a : b;
      }
    }
  }
};

// ---- packed 32-bit candidate list (fast path of the kNN kernel) ----
// Inserting into the exact 64-bit list costs ~6 instructions per slot in a dependent chain, and the
// round-1 profile shows that this insertion is 60-70 % of the kNN kernel. A 32-bit key
//   (top bits of d2) << kbits | index        kbits = ceil(log2 n)
// can be inserted with one unsigned min and one max per slot (independent, no predicates, no payload).
// It orders candidates by a QUANTISED distance, so it is only a filter: the list keeps the M = K + 4
// smallest packed keys, and afterwards
//   - complete() tells whether every candidate of the K-th entry's distance bucket is in the list
//     (the bucket of entry M-1 is strictly larger): then the exact top-K is a subset of the list and
//     exact_from_packed() recomputes exact 64-bit keys for the M entries and sorts them;
//   - otherwise (many ties, e.g. lattice data) the caller redoes the query with the exact list.
// Every candidate that could still belong to the exact top-K passes the gate "not beyond the K-th
// entry's bucket", so nothing is lost; bound2() is the upper edge of that bucket, a valid (slightly
// loose) bound for row pruning and for the stopping rule.
APD_HD unsigned umin32(unsigned a, unsigned b) { return a < b ? a : b; }
APD_HD unsigned umax32(unsigned a, unsigned b) { return a > b ? a : b; }

template <int K, int M>
struct TopKPacked {
  unsigned a[M];
  float gate;       // a distance above it cannot belong to the result: upper edge of the K-th entry's bucket, FLT_MAX while the
                    // list is not full (NaN / inf distances compare false and never enter)
  int kbits, sh;
  APD_HD void setup(int n) {
    int b = 1;
    while (b < 31 && (1u << b) < (unsigned)n) b++;
    kbits = b;
    sh = b - 1;  // 31 significant bits of d2 (it is >= 0) minus the 32 - kbits we keep
  }
  APD_HD void init() {  // kbits / sh must be set
#pragma unroll
    for (int i = 0; i < M; i++) a[i] = 0xFFFFFFFFu;
    gate = FLT_MAX;
  }
  // (Feeding the list FMA-contracted distances, with one bucket of margin in the gate, the bound and the
  // completeness test, was measured: the looser bound costs what the two saved instructions gain.)
  APD_HD float bound2() const {
    // all-ones (empty slot) decodes to a NaN: comparisons against it are false, i.e. "no bound yet"
    return u2f(((a[K - 1] >> kbits) << sh) | ((1u << sh) - 1u));
  }
  // The gate is ONE float compare per candidate (for non-negative floats the order of the values is the order
  // of their bit patterns, so "d2 <= upper edge of the bucket" is a bucket compare); the packed key is only
  // built for the few candidates that pass.
  // (Parking the candidates that pass the gate and inserting them behind the candidate loop of a run was
  // measured: the lanes of a warp are not converged there either, 7.8 -> 8.5 ms per 1001 scans.)
  APD_HD void offer(float d2, unsigned idx, int /*pos*/) {
    if (d2 <= gate) {
      const unsigned key = ((f2u(d2) >> sh) << kbits) | idx;
#pragma unroll
      for (int j = M - 1; j > 0; j--) a[j] = umin32(a[j], umax32(a[j - 1], key));
      a[0] = umin32(a[0], key);
      gate = fminf(bound2(), FLT_MAX);  // NaN (list not full) or beyond the finite range -> FLT_MAX
    }
  }
  APD_HD void end_run() {}
  APD_HD bool complete() const { return a[K - 1] == 0xFFFFFFFFu || (a[M - 1] >> kbits) > (a[K - 1] >> kbits); }
};

// Exact (d2, index)-ordered top-K from a complete packed list: recompute the exact keys of its M
// entries from the points (original order, `opts`) and sort them. The packed order already agrees
// with the exact one except inside a distance bucket, so an odd-even transposition sort needs one
// or two passes.
template <int K, int M>
APD_HD void exact_from_packed(const TopKPacked<K, M>& ap, float qx, float qy, float qz, const float4* opts, TopK<K>& tk) {
  unsigned long long k64[M];
  const unsigned mask = (1u << ap.kbits) - 1u;
#pragma unroll
  for (int j = 0; j < M; j++) {
    k64[j] = APD_KEY_INF;
    if (ap.a[j] != 0xFFFFFFFFu) {
      const unsigned idx = ap.a[j] & mask;
      const float4 t = opts[idx];
      k64[j] = make_key(sqdist_rn(qx, qy, qz, t.x, t.y, t.z), idx);
    }
  }
  bool swapped = true;
  while (swapped) {
    swapped = false;
#pragma unroll
    for (int j = 0; j + 1 < M; j += 2) {
      const unsigned long long x = k64[j], y = k64[j + 1];
      const bool sw = y < x;
      k64[j] = sw ? y : x;
      k64[j + 1] = sw ? x : y;
      swapped |= sw;
    }
#pragma unroll
    for (int j = 1; j + 1 < M; j += 2) {
      const unsigned long long x = k64[j], y = k64[j + 1];
      const bool sw = y < x;
      k64[j] = sw ? y : x;
      k64[j + 1] = sw ? x : y;
      swapped |= sw;
    }
  }
#pragma unroll
  for (int j = 0; j < K; j++) tk.key[j] = k64[j];
}

struct Top1 {
  float d2;      // best squared distance so far (+inf: none)
  unsigned idx;  // its original index (tie-break)
  int pos;       // its position in the searched (sorted) order
  APD_HD void init() { d2 = u2f(0x7f800000u); idx = 0xFFFFFFFFu; pos = -1; }
  APD_HD float bound2() const { return d2; }
  APD_HD void end_run() {}
  // (d2, index) lexicographic minimum; the float compare alone settles almost every candidate
  APD_HD void offer(float d, unsigned i, int p) {
    if (d <= d2) {
      if (d < d2 || i < idx) { d2 = d; idx = i; pos = p; }
    }
  }
};

template <typename CellT, typename Visitor>
APD_HD void visit_run(const GridView<CellT>& G, int c0, int c1, float qx, float qy, float qz, Visitor& vis) {
  const int s = (int)G.cells[c0];
  const int e = (int)G.cells[c1];
  for (int p = s; p < e; p++) {
    const float4 t = G.spts[p];
    vis.offer(sqdist_rn(qx, qy, qz, t.x, t.y, t.z), f2u(t.w), p);
  }
  vis.end_run();
}

// Exact search. `limit2` bounds the radius of interest (squared): once every unvisited point is
// provably farther than both the visitor's current bound and limit2, the search stops. Pass
// +inf for an unbounded search. NaN query coordinates fall into cell 0 and produce NaN distances,
// which never beat a key (the visitor stays empty), like a kd-tree that finds nothing.
// `max_ring` caps the expansion: the function returns true when the search is COMPLETE (stopping rule
// met or the whole grid covered) and false when it gave up after ring max_ring, in which case the
// caller restarts on a coarser level of the grid pyramid (a sparse neighbourhood would otherwise walk
// (2r+1)^2 mostly empty rows per ring).
template <typename CellT, typename Visitor>
APD_HD bool grid_search(const GridView<CellT>& G, float qx, float qy, float qz, float limit2, Visitor& vis, int max_ring = 0x7fffffff,
                        float* unexplored = nullptr) {
  // *unexplored: on a complete search, a lower bound of the distance from the query to every point that
  // was NOT offered to the visitor: points beyond the explored cube and points of rows that were skipped
  // (+inf when every point was offered)
  if (unexplored) *unexplored = 0.f;
  float skipped2 = FLT_MAX;  // smallest lower bound (squared) among the skipped rows
  const GridParams& g = G.g;
  const int cx = cell_coord(qx, g.lox, g.inv_h, g.nx);
  const int cy = cell_coord(qy, g.loy, g.inv_h, g.ny);
  const int cz = cell_coord(qz, g.loz, g.inv_h, g.nz);
  const int rmax = imax(imax(imax(cx, g.nx - 1 - cx), imax(cy, g.ny - 1 - cy)), imax(cz, g.nz - 1 - cz));
  for (int r = 0; r <= rmax; r++) {
    const int z0 = imax(cz - r, 0), z1 = imin(cz + r, g.nz - 1);
    const int y0 = imax(cy - r, 0), y1 = imin(cy + r, g.ny - 1);
    const int xa = cx - r, xb = cx + r;
    const int x0 = imax(xa, 0), x1 = imin(xb, g.nx - 1);
    for (int z = z0; z <= z1; z++) {
      // lower bound of |dz| to any point of this z-slab (0 for the query's own slab)
      float dz = 0.f;
      if (z < cz) dz = qz - (g.loz + (float)(z + 1) * g.h);
      else if (z > cz) dz = (g.loz + (float)z * g.h) - qz;
      dz = fmaxf(dz - g.slack, 0.f);
      const bool zface = (z == cz - r) || (z == cz + r);
      for (int y = y0; y <= y1; y++) {
        float dy = 0.f;
        if (y < cy) dy = qy - (g.loy + (float)(y + 1) * g.h);
        else if (y > cy) dy = (g.loy + (float)y * g.h) - qy;
        dy = fmaxf(dy - g.slack, 0.f);
        // whole row out of reach (strictly farther than the current bound): skip. 0.99999 absorbs the
        // rounding of this bound itself; the bound is re-read per row so it tightens inside a ring.
        // (Trimming the row to the chord of the bounding ball was measured in round 1: the extra sqrt and
        // two cell lookups per row cost more than the candidates they save.)
        const float row_lb2 = (dz * dz + dy * dy) * 0.99999f;
        if (row_lb2 > fminf(vis.bound2(), limit2)) {
          skipped2 = fminf(skipped2, row_lb2);
          continue;
        }
        const bool face = zface || (y == cy - r) || (y == cy + r);
        const int rowbase = (z * g.ny + y) * g.nx;
        // face row: the full x-run of this ring belongs to the shell (r == 0: the query's own cell);
        // interior row: only the two end cells at x = cx-r and cx+r are new.
        // Both shapes go through ONE visit site (a rolled two-part loop): the visitor's insertion code is
        // the bulk of the kernels' instruction footprint, and three inlined copies per search level thrashed
        // the 32 KB instruction cache (ncu: no_instruction was the top stall of the kNN kernel).
        int ca0 = rowbase + x0, ca1 = rowbase + x1 + 1, cb0 = 0, cb1 = 0;
        if (!face) {
          ca0 = rowbase + xa; ca1 = ca0 + (xa >= 0 ? 1 : 0);
          cb0 = rowbase + xb; cb1 = cb0 + (xb <= g.nx - 1 ? 1 : 0);
        }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (int part = 0; part < 2; part++) {
          const int c0 = part ? cb0 : ca0, c1 = part ? cb1 : ca1;
          if (c1 > c0) visit_run(G, c0, c1, qx, qy, qz, vis);
        }
      }
    }
    // Every unvisited point lies outside the cube of cells [c-r, c+r] along at least one axis, hence
    // at least `reach` away, where reach = min over the cube faces that still have cells beyond them
    // of the distance from the query to that face. Stop when that clears the current bound.
    float reach = FLT_MAX;
    if (cx - r > 0) reach = fminf(reach, qx - (g.lox + (float)(cx - r) * g.h));
    if (cx + r < g.nx - 1) reach = fminf(reach, (g.lox + (float)(cx + r + 1) * g.h) - qx);
    if (cy - r > 0) reach = fminf(reach, qy - (g.loy + (float)(cy - r) * g.h));
    if (cy + r < g.ny - 1) reach = fminf(reach, (g.loy + (float)(cy + r + 1) * g.h) - qy);
    if (cz - r > 0) reach = fminf(reach, qz - (g.loz + (float)(cz - r) * g.h));
    if (cz + r < g.nz - 1) reach = fminf(reach, (g.loz + (float)(cz + r + 1) * g.h) - qz);
    if (reach == FLT_MAX) {  // the cube covers the whole grid
      if (unexplored) *unexplored = sqrtf(skipped2) * 0.999995f;
      return true;
    }
    reach = reach - g.slack;
    if (reach > 0.f) {
      const float reach2 = reach * reach * 0.99999f;
      if (reach2 > fminf(vis.bound2(), limit2)) {
        if (unexplored) *unexplored = fminf(reach, sqrtf(skipped2)) * 0.999995f;
        return true;
      }
    }
    if (r >= max_ring) return false;
  }
  if (unexplored) *unexplored = sqrtf(skipped2) * 0.999995f;
  return true;
}

// Exact search inside a ball whose squared radius `B` is KNOWN to contain the answer (the k-th
// neighbour for a TopK visitor, the nearest neighbour for Top1): one pass over the (y,z) rows that
// intersect the ball, each row's run trimmed in x to the ball's chord. No ring bookkeeping and no
// early termination are needed because B already bounds the result. B comes from
//   - the triangle inequality between consecutive queries, r_k(q') <= r_k(q) + |q - q'|, or
//   - the distance to a seed point (the previous iteration's correspondence).
// Correctness only needs cell_coord to be monotone (it is: fsub, fmul by a positive constant and
// floorf are monotone) and the radius below to be rounded up; points farther than B are dropped
// before they reach the visitor, so a visitor that is not yet full never fills up with them.
template <typename CellT, typename Visitor>
APD_HD void grid_ball_search(const GridView<CellT>& G, float qx, float qy, float qz, float B, Visitor& vis) {
  const GridParams& g = G.g;
  const float rad = sqrtf(B) * 1.00002f + g.slack;
  const float rad2 = rad * rad * 1.00002f;
  const int z0 = cell_coord(qz - rad, g.loz, g.inv_h, g.nz), z1 = cell_coord(qz + rad, g.loz, g.inv_h, g.nz);
  const int y0 = cell_coord(qy - rad, g.loy, g.inv_h, g.ny), y1 = cell_coord(qy + rad, g.loy, g.inv_h, g.ny);
  for (int z = z0; z <= z1; z++) {
    const float zl = g.loz + (float)z * g.h;
    const float dz = fmaxf(fmaxf(zl - qz, qz - (zl + g.h)) - g.slack, 0.f);  // lower bound of |dz| into this slab
    const float remz = rad2 - dz * dz;
    if (remz < 0.f) continue;
    for (int y = y0; y <= y1; y++) {
      const float yl = g.loy + (float)y * g.h;
      const float dy = fmaxf(fmaxf(yl - qy, qy - (yl + g.h)) - g.slack, 0.f);
      const float rem = remz - dy * dy;
      if (rem < 0.f) continue;
      const float w = sqrtf(rem) + g.slack;  // half chord of the ball in this row (rounded up via rad2)
      const int xa = cell_coord(qx - w, g.lox, g.inv_h, g.nx), xb = cell_coord(qx + w, g.lox, g.inv_h, g.nx);
      const int rowbase = (z * g.ny + y) * g.nx;
      const int s = (int)G.cells[rowbase + xa];
      const int e = (int)G.cells[rowbase + xb + 1];
      for (int p = s; p < e; p++) {
        const float4 t = G.spts[p];
        const float d2 = sqdist_rn(qx, qy, qz, t.x, t.y, t.z);
        if (d2 <= B) vis.offer(d2, f2u(t.w), p);
      }
      vis.end_run();
    }
  }
}

// Upper bound of the squared k-th neighbour distance of q' from the result at q (triangle
// inequality), inflated so that fp32 rounding of either distance can never make it too small.
APD_HD float chained_bound2(float rk2_prev, float step2) {
  const float r = sqrtf(rk2_prev) + sqrtf(step2);
  return r * r * 1.0002f + 1e-30f;
}

}  // namespace apd
