// Leaf structure for clouds that fit one SM's shared memory (a radar scan: a few thousand points), and the
// warp-cooperative exact searches on it. Replaces pcl::search::KdTree::nearestKSearch as the reference uses it at
// fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:151 (1-NN per iteration) and :316 (k-NN for covariances)
// for such clouds; larger clouds keep the voxel grid of apd_grid.cuh.
//
// Why not the grid: round 1 ran one query per thread through a ring search on a uniform grid. Every lane walked its own
// rows and runs, and ncu showed 11.7 (kNN) / 12.8 (align) of 32 lanes active per instruction, at 70 % issue utilisation.
// A radar scan occupies 4 % of its bounding box's cells (points lie on surfaces), so no cell size fits.
//
// Layout: the cloud is sorted along a Hilbert curve and cut into LEAVES of 32 consecutive points, each with its
// axis-aligned bounding box. A warp owns the 32 queries of one leaf (kNN) or 32 consecutive cell-sorted source points
// (1-NN): queries that are close in space. All control flow is WARP-UNIFORM:
//   - the warp picks the unvisited leaf nearest to the group's bounding box (one REDUX over cached box distances) until
//     that distance exceeds the largest per-lane bound: leaves come nearest first, so bounds tighten early;
//   - a leaf is scanned by all lanes together: the candidate is a shared-memory BROADCAST (one wavefront), each lane
//     computes the distance to its own query. Two candidates per instruction (add / mul .f32x2, sm_100 packed fp32), with
//     every operation rounded like the scalar one: d2 = ((dx*dx + dy*dy) + dz*dz), FLANN L2_Simple, no contraction;
//   - a lane skips nothing by itself; the warp skips a leaf when NO lane's bound reaches its box.
// A lane therefore evaluates candidates it does not need (the union of what the 32 queries need: ~400 instead of ~150),
// but every evaluation is a converged, branch-free affair of a few issue slots instead of a divergent walk.
// Results are ordered by (d2, original index) exactly as before: index sets stay bit-exact against the CPU oracle.
#pragma once
#include "apd_math.cuh"

namespace apd {

constexpr int kLeaf = 32;              // points per leaf = lanes per warp
constexpr int kLeafMaxPoints = 8192;   // 13-bit positions in packed keys; 256 leaves = 8 rounds of cached box distances
constexpr int kLeafPosBits = 13;
constexpr int kLeafMaxRounds = kLeafMaxPoints / kLeaf / 32;

#ifdef __CUDACC__

// order-preserving float <-> unsigned (for REDUX min / max on coordinates)
__device__ __forceinline__ unsigned leaf_enc(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float leaf_dec(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// Packed fp32 (sm_100 add / mul .f32x2): two candidates per instruction, each half rounded to nearest like the scalar
// operation. Two traps, both measured: the float2 intrinsics (__fmul2_rn + __fadd2_rn) are contracted by the compiler into
// FFMA2, and so is inline PTX with explicit .rn modifiers (ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2, unlike the scalar
// forms). A fused multiply-add changes the last bit of a distance (the reference is built without FMA,
// fast_apdgicp/CMakeLists.txt:11-13). So the differences and the squares are packed (FADD2, FMUL2) and the two SUMS of a
// distance are scalar __fadd_rn on the halves, which ptxas cannot merge with a packed multiply: 10 instructions per
// candidate pair instead of 16, and bit-exact L2_Simple.
typedef unsigned long long f32x2_t;  // two floats in one 64-bit register: low half = first
__device__ __forceinline__ f32x2_t f2_pack(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// squared distances of one query (coordinates duplicated in both halves) to the two candidates of a staged pair:
// ((dx*dx + dy*dy) + dz*dz) per half, FLANN L2_Simple order; A = (-x0,-x1 | -y0,-y1), Bz = (-z0,-z1)
__device__ __forceinline__ void leaf_pair_d2(f32x2_t qx2, f32x2_t qy2, f32x2_t qz2, const ulonglong2& A, f32x2_t Bz, float& d0, float& d1) {
  const f32x2_t dx = f2_add(qx2, A.x), dy = f2_add(qy2, A.y), dz = f2_add(qz2, Bz);
  float x0, x1, y0, y1, z0, z1;
  f2_unpack(f2_mul(dx, dx), x0, x1);
  f2_unpack(f2_mul(dy, dy), y0, y1);
  f2_unpack(f2_mul(dz, dz), z0, z1);
  d0 = __fadd_rn(__fadd_rn(x0, y0), z0);
  d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
}

// A staged cloud: candidate pairs in shared memory, NEGATED so that a difference is one packed add:
//   pair j = points 2j, 2j+1:  P[2j] = (-x0, -x1, -y0, -y1)   P[2j+1] = (-z0, -z1, orig0, orig1)   (orig = original index bits)
// Slots beyond n (the last leaf is padded to 32) hold NaN coordinates: every distance to them is NaN and fails every
// `<=` test. Non-finite input points behave the same way and are sorted to the end of the cloud by the build.
struct LeafView {
  const float4* P;    // shared memory: 2 float4 per pair, 16 pairs per leaf
  const float4* box;  // shared memory: 2 float4 per leaf: (lox, loy, loz, -), (hix, hiy, hiz, -); an empty leaf has lo = +inf, hi = -inf
  int n;              // points
  int nleaf;          // ceil(n / 32)
};

__device__ __forceinline__ size_t leaf_stage_bytes(int n) {
  const int nleaf = (n + kLeaf - 1) / kLeaf;
  return (size_t)nleaf * (kLeaf * 16 + 32);
}

// Copy one cloud's sorted points and leaf boxes into shared memory in the pair layout (whole CTA; caller syncs).
__device__ __forceinline__ void leaf_stage(float4* __restrict__ sP, float4* __restrict__ sbox, const float4* __restrict__ gspts, const float4* __restrict__ gbox, int n) {
  const int nleaf = (n + kLeaf - 1) / kLeaf;
  const int npairs = nleaf * (kLeaf / 2);
  const float qnan = __int_as_float(0x7fc00000);
  for (int j = threadIdx.x; j < npairs; j += blockDim.x) {
    const int i0 = 2 * j, i1 = 2 * j + 1;
    float4 a = make_float4(qnan, qnan, qnan, __uint_as_float(0xFFFFFFFFu)), b = a;
    if (i0 < n) a = gspts[i0];
    if (i1 < n) b = gspts[i1];
    sP[2 * j] = make_float4(-a.x, -b.x, -a.y, -b.y);
    sP[2 * j + 1] = make_float4(-a.z, -b.z, a.w, b.w);
  }
  for (int j = threadIdx.x; j < 2 * nleaf; j += blockDim.x) sbox[j] = gbox[j];
}

// the point in slot `pos` of a staged cloud: (x, y, z, original index bits)
__device__ __forceinline__ float4 leaf_point(const LeafView& L, int pos) {
  const float4 A = L.P[2 * (pos >> 1)], B = L.P[2 * (pos >> 1) + 1];
  return (pos & 1) ? make_float4(-A.y, -A.w, -B.y, B.w) : make_float4(-A.x, -A.z, -B.x, B.z);
}

// squared distance from a point to a box (0 inside), as a lower bound of the distance to every point of the box;
// scaled down so that float rounding of this bound itself can never exclude a candidate (same 0.99999 as the grid search)
__device__ __forceinline__ float leaf_point_box2(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return (dx * dx + dy * dy + dz * dz) * 0.99999f;  // an empty box (lo = +inf) gives +inf
}
__device__ __forceinline__ float leaf_box_box2(const float (&alo)[3], const float (&ahi)[3], const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(lo.x - ahi[0], alo[0] - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - ahi[1], alo[1] - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - ahi[2], alo[2] - hi.z), 0.f);
  return (dx * dx + dy * dy + dz * dz) * 0.99999f;
}

// The nearest-first leaf schedule of one warp. Every lane caches the (group box -> leaf box) distances of the leaves
// lane, lane + 32, ...; next() returns the unvisited leaf closest to the group, or -1 once that distance exceeds G.
struct LeafSchedule {
  float d[kLeafMaxRounds];
  int rounds;
  __device__ __forceinline__ void init(const LeafView& L, const float (&glo)[3], const float (&ghi)[3], int skip_leaf) {
    const int lane = threadIdx.x & 31;
    rounds = (L.nleaf + 31) >> 5;
#pragma unroll
    for (int t = 0; t < kLeafMaxRounds; t++) {
      d[t] = __int_as_float(0x7f800000);
      if (t < rounds) {
        const int l = t * 32 + lane;
        if (l < L.nleaf && l != skip_leaf) d[t] = leaf_box_box2(glo, ghi, L.box[2 * l], L.box[2 * l + 1]);
      }
    }
  }
  // G: the largest bound any lane still has (squared). All lanes get the same answer.
  __device__ __forceinline__ int next(float G) {
    const int lane = threadIdx.x & 31;
    unsigned best = 0xFFFFFFFFu;
#pragma unroll
    for (int t = 0; t < kLeafMaxRounds; t++)
      if (t < rounds) best = min(best, (__float_as_uint(d[t]) & 0xFFFFFF00u) | (unsigned)t << 5 | (unsigned)lane);  // d >= 0: bit order is value order; rounded DOWN
    best = __reduce_min_sync(0xFFFFFFFFu, best);
    if ((best & 0xFFFFFF00u) >= 0x7F800000u) return -1;  // every leaf visited (or empty: box distance +inf)
    if (!(__uint_as_float(best & 0xFFFFFF00u) <= G)) return -1;
    const int l = (int)(best & 0xFFu);
#pragma unroll
    for (int t = 0; t < kLeafMaxRounds; t++)
      if (t == (l >> 5) && lane == (l & 31)) d[t] = __int_as_float(0x7f800000);
    return l;
  }
};

// Bounding box of the warp's valid queries (REDUX on order-preserving encodings). Lanes with valid == false do not count.
__device__ __forceinline__ void leaf_group_box(float qx, float qy, float qz, bool valid, float (&glo)[3], float (&ghi)[3]) {
  const float q[3] = {qx, qy, qz};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const unsigned e = leaf_enc(q[a]);
    glo[a] = leaf_dec(__reduce_min_sync(0xFFFFFFFFu, valid ? e : 0xFFFFFFFFu));
    ghi[a] = leaf_dec(__reduce_max_sync(0xFFFFFFFFu, valid ? e : 0u));
  }
}

// ---- 1-NN (update_correspondences, getFitnessScore) ----
// Per-lane state of an exact nearest-neighbour search ordered by (d2, original index).
struct LeafTop1 {
  float d2;       // best squared distance so far; also the lane's search bound
  unsigned idx;   // original index of the best (tie-break)
  int pos;        // its slot in the staged (sorted) target, -1 = none
};

// Scan one leaf for all 32 lanes: candidates are broadcast, two per packed instruction.
__device__ __forceinline__ void leaf_scan_top1(const LeafView& L, int leaf, f32x2_t qx2, f32x2_t qy2, f32x2_t qz2, LeafTop1& v) {
  const ulonglong2* P = reinterpret_cast<const ulonglong2*>(L.P) + leaf * kLeaf;  // 2 x 16 bytes per pair, 16 pairs
#pragma unroll 4
  for (int j = 0; j < kLeaf / 2; j++) {
    const ulonglong2 A = P[2 * j], B = P[2 * j + 1];
    float d0, d1;
    leaf_pair_d2(qx2, qy2, qz2, A, B.x, d0, d1);
    // (d2, index) lexicographic minimum; NaN (padding, non-finite points) compares false
    if (d0 <= v.d2) {
      const unsigned i = (unsigned)(B.y & 0xFFFFFFFFull);
      if (d0 < v.d2 || i < v.idx) { v.d2 = d0; v.idx = i; v.pos = leaf * kLeaf + 2 * j; }
    }
    if (d1 <= v.d2) {
      const unsigned i = (unsigned)(B.y >> 32);
      if (d1 < v.d2 || i < v.idx) { v.d2 = d1; v.idx = i; v.pos = leaf * kLeaf + 2 * j + 1; }
    }
  }
}

// Exact nearest neighbour of every lane's query among the staged target, inside the lane's initial bound v.d2
// (+inf: unbounded; a seed candidate may already sit in v). valid == false lanes take no part. Warp-collective.
__device__ __forceinline__ void leaf_nn1(const LeafView& L, float qx, float qy, float qz, bool valid, LeafTop1& v) {
  if (!__any_sync(0xFFFFFFFFu, valid)) return;
  float glo[3], ghi[3];
  leaf_group_box(qx, qy, qz, valid, glo, ghi);
  LeafSchedule S;
  S.init(L, glo, ghi, -1);
  const f32x2_t qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
  if (!valid) v.d2 = -1.f;  // nothing passes `<= -1`
  for (;;) {
    // the largest bound of the group (bit patterns of non-negative floats order like the values; -1 counts as 0)
    const float G = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, valid ? __float_as_uint(v.d2) : 0u));
    const int l = S.next(G);
    if (l < 0) break;
    const float dl = leaf_point_box2(qx, qy, qz, L.box[2 * l], L.box[2 * l + 1]);
    if (!__any_sync(0xFFFFFFFFu, valid && dl <= v.d2)) continue;
    leaf_scan_top1(L, l, qx2, qy2, qz2, v);
  }
}

#endif  // __CUDACC__

}  // namespace apd
