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
constexpr int kLeafMaxPoints = 6144;   // 13-bit positions in packed keys; 192 leaves = 6 rounds of cached box distances
constexpr int kLeafPosBits = 13;
constexpr int kLeafMaxRounds = kLeafMaxPoints / kLeaf / 32;
constexpr int kLeafTransposeMax = 20;  // 1-NN: up to this many queries needing a leaf take turns (transposed scan); more: one broadcast scan

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
//   pair j = points 2j, 2j+1:  P[2j] = (-x0, -x1, -y0, -y1)   P[2j+1] = (-z0, -z1, tag0, tag1)
//   tag = original index << 13 | slot: the low word of a candidate's exact 64-bit key (d2 bits << 32 | tag), whose unsigned order
//   IS the (d2, original index) order of the results (original indices are below 8192 in leaf mode)
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
    float4 a = make_float4(qnan, qnan, qnan, __uint_as_float(0x7FFFFu)), b = a;
    if (i0 < n) a = gspts[i0];
    if (i1 < n) b = gspts[i1];
    sP[2 * j] = make_float4(-a.x, -b.x, -a.y, -b.y);
    sP[2 * j + 1] = make_float4(-a.z, -b.z, __uint_as_float(__float_as_uint(a.w) << kLeafPosBits | (unsigned)i0), __uint_as_float(__float_as_uint(b.w) << kLeafPosBits | (unsigned)i1));
  }
  for (int j = threadIdx.x; j < 2 * nleaf; j += blockDim.x) sbox[j] = gbox[j];
}

// ---- staging by the copy engine: cp.async.bulk (global -> shared, completion on an mbarrier) of the image the leaf build wrote ----
// One elected thread arms the barrier with the byte count and issues the copy; every thread then waits on the barrier's phase. The
// bytes never pass through registers or the LSU, the issuing warp is free at once, and the data lands in the async proxy's order:
// the mbarrier wait is what makes it visible to the generic-proxy loads that follow. SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LEAF_MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LEAF_MBAR_DONE_%=;\n"
      "bra LEAF_MBAR_WAIT_%=;\n"
      "LEAF_MBAR_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// Whole CTA; the caller has synchronised the CTA since the last read of the destination. `image` = limg of the cloud (16-byte aligned),
// the destination receives leaf_stage_bytes(n) bytes: points in the pair layout, then the boxes. Returns after the data is visible.
__device__ __forceinline__ void leaf_stage_bulk(float4* __restrict__ sP, const float4* __restrict__ image, int n, unsigned long long* bar, unsigned& parity) {
  const unsigned bytes = (unsigned)leaf_stage_bytes(n);
  if (bytes == 0) return;
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of the destination are ordered before the copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sP)), "l"(image), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
  }
  mbar_wait(bar, parity);
  parity ^= 1u;
}
// The build's side: one leaf's 32 sorted points (lane = slot in the leaf, idx = original index, 0x7FFFF and NaN coordinates beyond n)
// written into the image in the pair layout: even lanes store (-x0, -x1, -y0, -y1), odd lanes (-z0, -z1, tag0, tag1).
__device__ __forceinline__ void leaf_image_store_points(float4* __restrict__ image, int slot, float x, float y, float z, unsigned idx, int lane) {
  const unsigned tag = idx << kLeafPosBits | (unsigned)slot;
  const float px = __shfl_xor_sync(0xFFFFFFFFu, x, 1), py = __shfl_xor_sync(0xFFFFFFFFu, y, 1), pz = __shfl_xor_sync(0xFFFFFFFFu, z, 1);
  const unsigned ptag = __shfl_xor_sync(0xFFFFFFFFu, tag, 1);
  image[slot] = (lane & 1) ? make_float4(-pz, -z, __uint_as_float(ptag), __uint_as_float(tag)) : make_float4(-x, -px, -y, -py);
}

// the point in slot `pos` of a staged cloud: (x, y, z, tag bits); tag >> 13 = original index
__device__ __forceinline__ float4 leaf_point(const LeafView& L, int pos) {
  const float4 A = L.P[2 * (pos >> 1)], B = L.P[2 * (pos >> 1) + 1];
  return (pos & 1) ? make_float4(-A.y, -A.w, -B.y, B.w) : make_float4(-A.x, -A.z, -B.x, B.z);
}
__device__ __forceinline__ unsigned leaf_tag_index(float w) { return __float_as_uint(w) >> kLeafPosBits; }
// The same point for a GATHER (every lane a different, unrelated slot: neighbour lists, seeds, correspondences): four scalar
// loads of exactly the words wanted instead of two 16-byte loads of which half is dropped by selects - fewer instructions, and
// about half the shared-memory wavefronts (a random 16-byte access pattern serialises into 8-10 wavefronts per load; the
// compiler drops the tag load where only the coordinates are used). Slots in order (pos = base + lane) keep leaf_point.
__device__ __forceinline__ float4 leaf_point_gather(const LeafView& L, int pos) {
  const float* w = reinterpret_cast<const float*>(L.P) + ((pos >> 1) << 3) + (pos & 1);
  return make_float4(-w[0], -w[2], -w[4], w[6]);
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
      float v = __int_as_float(0x7f800000);
      if (t < rounds) {
        const int l = t * 32 + lane;
        if (l < L.nleaf && l != skip_leaf) v = leaf_box_box2(glo, ghi, L.box[2 * l], L.box[2 * l + 1]);
      }
      d[t] = v;
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
    // (a select per slot, not a conditional store: the compiler merges conditional stores into ONE dynamically indexed
    // store, which moves the whole array from registers to local memory)
    const int mine = lane == (l & 31) ? (l >> 5) : -1;
#pragma unroll
    for (int t = 0; t < kLeafMaxRounds; t++) d[t] = (t == mine) ? __int_as_float(0x7f800000) : d[t];
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
// Per-lane state of an exact nearest-neighbour search: the smallest 64-bit key (d2 bits << 32 | original index << 13 | slot)
// seen so far. Unsigned order of the key = (d2, original index) order, so ties break by index like the oracle.
// Before anything is found the low word is all ones and the high word is the search bound: a candidate AT the bound still wins.
struct LeafTop1 {
  unsigned long long key;
  __device__ __forceinline__ void init(float bound2) { key = ((unsigned long long)__float_as_uint(bound2) << 32) | 0xFFFFFFFFull; }
  __device__ __forceinline__ float d2() const { return __uint_as_float((unsigned)(key >> 32)); }  // best distance, or the bound while nothing is found
  __device__ __forceinline__ bool found() const { return (unsigned)key != 0xFFFFFFFFu; }
  __device__ __forceinline__ int pos() const { return found() ? (int)((unsigned)key & ((1u << kLeafPosBits) - 1u)) : -1; }
};

// Scan one leaf for all 32 lanes: candidates are broadcast, two per packed instruction. The update (a 64-bit unsigned minimum)
// sits behind a warp vote: once the first candidates are in, most pairs improve nobody's result.
__device__ __forceinline__ void leaf_scan_top1(const LeafView& L, int leaf, f32x2_t qx2, f32x2_t qy2, f32x2_t qz2, LeafTop1& v) {
  const ulonglong2* P = reinterpret_cast<const ulonglong2*>(L.P) + leaf * kLeaf;  // 2 x 16 bytes per pair, 16 pairs
#pragma unroll 2
  for (int j = 0; j < kLeaf / 2; j += 2) {   // four candidates per vote
    const ulonglong2 A0 = P[2 * j], B0 = P[2 * j + 1], A1 = P[2 * j + 2], B1 = P[2 * j + 3];
    float d0, d1, d2, d3;
    leaf_pair_d2(qx2, qy2, qz2, A0, B0.x, d0, d1);
    leaf_pair_d2(qx2, qy2, qz2, A1, B1.x, d2, d3);
    const float best = v.d2();
    // NaN distances (padding, non-finite points) compare false; an invalid lane's key is 0 (best = +0, and its distances are NaN)
    if (__any_sync(0xFFFFFFFFu, fminf(fminf(d0, d1), fminf(d2, d3)) <= best)) {
      const unsigned long long k0 = ((unsigned long long)__float_as_uint(d0) << 32) | (B0.y & 0xFFFFFFFFull);
      const unsigned long long k1 = ((unsigned long long)__float_as_uint(d1) << 32) | (B0.y >> 32);
      const unsigned long long k2 = ((unsigned long long)__float_as_uint(d2) << 32) | (B1.y & 0xFFFFFFFFull);
      const unsigned long long k3 = ((unsigned long long)__float_as_uint(d3) << 32) | (B1.y >> 32);
      const unsigned long long ka = k0 < k1 ? k0 : k1, kb = k2 < k3 ? k2 : k3;  // NaN bit patterns are above +inf: they never win
      const unsigned long long k = ka < kb ? ka : kb;
      if (k < v.key) v.key = k;
    }
  }
}

// The same leaf, TRANSPOSED: every lane holds one CANDIDATE of the leaf, and the queries that need the leaf (bit mask `need`)
// take turns. One turn = broadcast the query from the warp's shared-memory slots, 32 distances in parallel, one REDUX for the
// minimum; only when that minimum reaches the query's current best is anything updated. A turn costs ~15 instructions, so the
// transposed form wins whenever fewer than about 20 of the 32 queries need the leaf - the normal case once a search is seeded
// (a query's ball touches 2-3 leaf boxes, the group's 32 balls together touch 7-8).
// qslot: the warp's 32 x float4 (qx, qy, qz, bits of the query's best d2 / bound), kept current by the owning lane.
__device__ __forceinline__ void leaf_scan_top1_transposed(const LeafView& L, int leaf, unsigned need, float4* __restrict__ qslot, LeafTop1& v) {
  const int lane = threadIdx.x & 31;
  const float4 c = leaf_point(L, leaf * kLeaf + lane);  // this lane's candidate (NaN coordinates in padding slots)
  const unsigned ctag = __float_as_uint(c.w);
  // Two queries per turn: the loads, the distance chains and the two reductions of a pair are independent, which halves the
  // dependent latency a lone warp sees (single-pair latency mode: one or two groups per warp, nothing else to hide it).
  while (need) {
    const int q0 = __ffs(need) - 1;
    need &= need - 1;
    const int q1 = need ? __ffs(need) - 1 : q0;
    need &= need - 1;
    const float4 Q0 = qslot[q0], Q1 = qslot[q1];
    const unsigned d0 = __float_as_uint(sqdist_rn(Q0.x, Q0.y, Q0.z, c.x, c.y, c.z));
    const unsigned d1 = __float_as_uint(sqdist_rn(Q1.x, Q1.y, Q1.z, c.x, c.y, c.z));
    const unsigned m0 = __reduce_min_sync(0xFFFFFFFFu, d0);  // d2 >= +0: bit order is value order; NaN patterns are above +inf
    const unsigned m1 = __reduce_min_sync(0xFFFFFFFFu, d1);
    const bool h0 = m0 <= __float_as_uint(Q0.w), h1 = q1 != q0 && m1 <= __float_as_uint(Q1.w);  // warp-uniform
    if (h0 || h1) {
      __syncwarp();  // every lane has read the two slots before an owner updates its bound (write after read)
      // lowest tag (= lowest original index) among the candidates at that distance, then the 64-bit (d2, index) comparison
      if (h0) {
        const unsigned wtag = __reduce_min_sync(0xFFFFFFFFu, d0 == m0 ? ctag : 0xFFFFFFFFu);
        const unsigned long long nk = ((unsigned long long)m0 << 32) | wtag;
        if (lane == q0 && nk < v.key) { v.key = nk; qslot[q0].w = __uint_as_float(m0); }
      }
      if (h1) {
        const unsigned wtag = __reduce_min_sync(0xFFFFFFFFu, d1 == m1 ? ctag : 0xFFFFFFFFu);
        const unsigned long long nk = ((unsigned long long)m1 << 32) | wtag;
        if (lane == q1 && nk < v.key) { v.key = nk; qslot[q1].w = __uint_as_float(m1); }
      }
      __syncwarp();
    }
  }
  __syncwarp();  // the slots are rewritten by their owners after a broadcast scan: no lane may still be reading them here
}

// Exact nearest neighbour of every lane's query among the staged target, inside the bound v was initialised with
// (+inf: unbounded). valid == false lanes take no part. Warp-collective. qslot: 32 float4 of shared memory owned by this warp.
__device__ __forceinline__ void leaf_nn1(const LeafView& L, float qx, float qy, float qz, bool valid, LeafTop1& v, float4* __restrict__ qslot,
                                         unsigned long long* dbg = nullptr, unsigned long long* evals = nullptr) {
  if (!__any_sync(0xFFFFFFFFu, valid)) return;
  unsigned n_next = 0, n_bcast = 0, n_turns = 0;
  float glo[3], ghi[3];
  leaf_group_box(qx, qy, qz, valid, glo, ghi);
  LeafSchedule S;
  S.init(L, glo, ghi, -1);
  // an invalid lane must neither win a candidate nor hold the group's bound up: NaN coordinates, key 0
  const float nanq = __int_as_float(0x7fc00000);
  if (!valid) { qx = nanq; qy = nanq; qz = nanq; v.key = 0ull; }
  const f32x2_t qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
  const int lane = threadIdx.x & 31;
  __syncwarp();
  qslot[lane] = make_float4(qx, qy, qz, __uint_as_float((unsigned)(v.key >> 32)));
  __syncwarp();
  for (;;) {
    // the largest bound of the group (bit patterns of non-negative floats order like the values)
    const float G = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, (unsigned)(v.key >> 32)));
    const int l = S.next(G);
    if (l < 0) break;
    n_next++;
    const float dl = leaf_point_box2(qx, qy, qz, L.box[2 * l], L.box[2 * l + 1]);  // NaN for an invalid lane: compares false
    // (valid: fmaxf drops the NaN of an invalid lane's box distance, so that lane must be masked explicitly)
    const unsigned need = __ballot_sync(0xFFFFFFFFu, valid && dl <= v.d2());
    if (need == 0u) continue;
    if (__popc(need) > kLeafTransposeMax) {
      leaf_scan_top1(L, l, qx2, qy2, qz2, v);
      qslot[lane].w = __uint_as_float((unsigned)(v.key >> 32));
      __syncwarp();
      n_bcast++;
    } else {
      leaf_scan_top1_transposed(L, l, need, qslot, v);
      n_turns += __popc(need);
    }
  }
  if (evals && lane == 0) atomicAdd(evals, (unsigned long long)n_bcast * (kLeaf * 32) + (unsigned long long)n_turns * 32);
  const unsigned n_valid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
  if (dbg && lane == 0) {  // profiling aid: groups, leaves popped, broadcast scans, transposed turns, worst group
    atomicAdd(dbg + 0, 1ull);
    atomicAdd(dbg + 1, (unsigned long long)n_next);
    atomicAdd(dbg + 2, (unsigned long long)n_bcast);
    atomicAdd(dbg + 3, (unsigned long long)n_turns);
    atomicMax(dbg + 4, (unsigned long long)n_turns);
    atomicMax(dbg + 5, (unsigned long long)n_next);
    atomicAdd(dbg + 6, (unsigned long long)n_valid);
  }
}

#endif  // __CUDACC__

}  // namespace apd
