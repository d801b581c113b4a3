// The GPU-resident align loop: LsqRegistration::computeTransformation
// (fast_apdgicp/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:55-80), step_lm / step_gn
// (:107-173), is_converged (:83-92) and FastAPDGICP::update_correspondences / linearize /
// compute_error (fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:133-298), plus the
// pcl::Registration::getFitnessScore pass the callers run afterwards.
//
// One persistent kernel runs the whole optimisation of a scan pair without returning to the host.
// A pair is owned by a TEAM of CTAs:
//   TEAM_CTA      one CTA per pair; CTAs pull pairs from a global counter (batched throughput mode)
//   TEAM_CLUSTER  a thread-block cluster per pair; partial Hessians are exchanged through
//                 distributed shared memory (single-pair latency mode)
//   TEAM_GRID     the whole cooperative grid on one pair; partials go through HBM/L2 and a grid
//                 barrier (large clouds)
// Every CTA of a team evaluates the 6x6 solve and all accept/reject/convergence decisions
// redundantly from bit-identical reduced sums, so control flow stays uniform without extra traffic.
//
// Reduction record (kNRed = 30 doubles): [0..5] S^T M S (xx xy xz yy yz zz), [6..14] N = S^T M (row
// major), [15..20] sum M, [21..23] S^T M e, [24..26] M e, [27] e^T M e, [28] inlier count, [29] spare.
// Per thread the partials are accumulated in fp64 registers, combined inside a warp with shuffles,
// across warps by a fixed-order shared-memory tree and across CTAs in rank order: the sum is
// deterministic for a given launch shape.
#include <cooperative_groups.h>

#include <cstdint>
#include <type_traits>

#include "apd_internal.h"
#include "apd_leaf.cuh"

namespace cg = cooperative_groups;

namespace apd {

namespace {

struct AlignShared {
  double x0[12];   // current pose: R row-major (9) + t (3)
  double xi[12];   // trial pose
  double H[36];
  double b[6];
  double d[6];
  double delta[12];
  double red[kNRed];         // team-reduced record
  double part[2][kNRed];     // this CTA's partial (double-buffered for DSMEM readers)
  double warp_part[kAlignThreads / 32][kNRed];
  double lambda, nu, y0, yi;
  float Tf[12];
  int decision;    // LM trial: 0 rejected, 1 accepted, 2 rejected but converged
  int converged;
  int pair;
  int staged_target;  // cloud index currently staged, -1 none
  int next;           // next unclaimed source point of the running search pass (dynamic chunks)
  float4 qslot[kAlignThreads / 32][32];  // leaf mode: every warp's 32 queries (x, y, z, bits of the best d2) for the transposed scans
};

// profiling aid (apd_set_option "timeline"): nanosecond stamps of the phases of the first pair, read back by apd_get_timeline
__device__ __forceinline__ void stamp_tl(unsigned long long* tl, int phase) {
  if (tl && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long n = tl[0];
    if (n < 250) {
      tl[1 + 2 * n] = (unsigned long long)phase;
      tl[2 + 2 * n] = t;
      tl[0] = n + 1;
    }
  }
}
__device__ __forceinline__ void stamp(const AlignBatch& B, int phase) { stamp_tl(B.timeline, phase); }

template <int TEAM>
struct TeamCtx {
  int size, rank, id, count;
  int buf;
  double* grid_partials;
  unsigned long long* tl;  // timeline (profiling aid), usually null
};

template <int TEAM>
__device__ __forceinline__ void team_sync() {
  if (TEAM == TEAM_CLUSTER) cg::this_cluster().sync();
  else if (TEAM == TEAM_GRID) { __threadfence(); cg::this_grid().sync(); }
  else __syncthreads();
}

// Warp-wide sums of NV <= 32 per-lane values in 31 shuffles instead of 5 * NV: a reduce-scatter. In the step with offset o every
// lane keeps the half of its values whose index has bit o equal to its own lane bit and hands the other half to lane ^ o, so the
// value count halves with every step and lane l ends up with the warp total of value l. (Shuffles issue at one warp per clock per
// SM: the 29 butterflies of a linearization cost 16 warps x 290 SHFL = 2.4 us per reduction, a third of a single-pair iteration.)
// The order of the additions is fixed by the lane numbers, so the result is deterministic.
template <int NV>
__device__ __forceinline__ double warp_reduce_scatter(const double (&acc)[kNRed]) {
  const int lane = threadIdx.x & 31;
  double a[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const double lo = j < NV ? acc[j] : 0.0, hi = j + 16 < NV ? acc[j + 16] : 0.0;
      const double recv = __shfl_xor_sync(0xFFFFFFFFu, up ? lo : hi, 16);
      a[j] = (up ? hi : lo) + recv;
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const bool up = lane & o;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j < o) {
        const double recv = __shfl_xor_sync(0xFFFFFFFFu, up ? a[j] : a[j + o], o);
        a[j] = (up ? a[j + o] : a[j]) + recv;
      }
    }
  }
  return a[0];
}

// acc: per-thread partial record. On return S.red holds the team-wide sums (same bits in every CTA).
template <int TEAM, int NV>
__device__ __forceinline__ void team_reduce(double (&acc)[kNRed], AlignShared& S, TeamCtx<TEAM>& tc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (NV > 4) {
    const double v = warp_reduce_scatter<NV>(acc);
    if (lane < NV) S.warp_part[warp][lane] = v;
  } else {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      if (lane == 0) S.warp_part[warp][i] = v;
    }
  }
  __syncthreads();
  if (NV > 4) stamp_tl(tc.tl, 30);  // warp sums done
  if (threadIdx.x < NV) {
    double v = 0.0;
    for (int w = 0; w < kAlignThreads / 32; w++) v += S.warp_part[w][threadIdx.x];
    if (TEAM == TEAM_CTA) S.red[threadIdx.x] = v;
    else if (TEAM == TEAM_CLUSTER) S.part[tc.buf][threadIdx.x] = v;
    else tc.grid_partials[((size_t)tc.buf * tc.size + tc.rank) * kNRed + threadIdx.x] = v;
  }
  if (TEAM == TEAM_CTA) {
    __syncthreads();
    return;
  }
  if (NV > 4) stamp_tl(tc.tl, 31);  // CTA partial written
  team_sync<TEAM>();
  if (NV > 4) stamp_tl(tc.tl, 32);  // team barrier passed
  // Cross-CTA sum in a fixed order that is the same in every CTA, so all of them end up with identical bits.
  if (TEAM == TEAM_CLUSTER && tc.size <= 16) {
    // a cluster has at most 16 ranks: each HALF-warp sums one value (lane & 15 = rank), so the 16 warps cover all 29 values of a
    // linearization in one round of remote reads instead of two
    cg::cluster_group cl = cg::this_cluster();
    const int r = lane & 15;
    for (int i0 = 2 * warp; i0 < NV; i0 += 2 * (kAlignThreads / 32)) {  // warp-uniform trip count: the shuffles below need all 32 lanes
      const int i = i0 + (lane >> 4);
      double v = (i < NV && r < tc.size) ? *cl.map_shared_rank(&S.part[tc.buf][i], r) : 0.0;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      if (r == 0 && i < NV) S.red[i] = v;
    }
  } else if (TEAM == TEAM_GRID) {
    // A half-warp per value, so all values of a linearization are summed in ONE round: lane s of the half adds the partials of
    // ranks s, s + 16, ... (ascending; the loads of a batch of 8 are independent and in flight together), then a fixed butterfly.
    // (One thread per value walking all ranks cost 148 dependent L2 round trips per reduction: 30 us of a 65 us iteration; one
    // warp per value still paid two rounds of three dependent round trips on a 79-CTA team.)
    const int s16 = lane & 15;
    for (int i0 = 2 * warp; i0 < NV; i0 += 2 * (kAlignThreads / 32)) {  // warp-uniform trip count: the shuffles below need all 32 lanes
      const int i = i0 + (lane >> 4);
      double v = 0.0;
      if (i < NV) {
        for (int r0 = s16; r0 < tc.size; r0 += 16 * 8) {
          double x[8];
#pragma unroll
          for (int m = 0; m < 8; m++) {
            const int r = r0 + 16 * m;
            x[m] = r < tc.size ? __ldcg(&tc.grid_partials[((size_t)tc.buf * tc.size + r) * kNRed + i]) : 0.0;
          }
#pragma unroll
          for (int m = 0; m < 8; m++) v += x[m];
        }
      }
      __syncwarp();
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      if (s16 == 0 && i < NV) S.red[i] = v;
    }
  } else {
    // clusters of more than 16 CTAs do not exist today; kept general: one warp per value, lane l adds ranks l, l + 32, ...
    for (int i = warp; i < NV; i += kAlignThreads / 32) {
      double v = 0.0;
      cg::cluster_group cl = cg::this_cluster();
      for (int r = lane; r < tc.size; r += 32) v += *cl.map_shared_rank(&S.part[tc.buf][i], r);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      if (lane == 0) S.red[i] = v;
    }
  }
  tc.buf ^= 1;
  __syncthreads();
}

// The search passes hand out their points in chunks of adjacent (cell-sorted) points from a shared counter:
// the cost of a chunk varies a lot (sparse neighbourhoods, anchors that skip the search), and with a static
// round-robin 14 % of the kernel's samples sat at the next barrier waiting for the slowest warp (round-1 ncu).
// The passes only write per-point results, so the assignment does not influence any sum. A CTA with fewer
// points than threads (cluster and grid teams) spreads them thinly over all its warps.
// Which source points a CTA of a team owns: blocks of 32 consecutive sorted points dealt round-robin over the team's CTAs
// (a contiguous split gave one CTA the dense near-range points and another the sparse clutter, whose searches take
// several times longer: the whole cluster then waited for it at every reduction). A CTA addresses its points by a LOCAL
// index j in [0, n_local); map(j) is the point, or >= ns past the end of the cloud. Team of one CTA: the identity.
// A large team on a small cloud (the cooperative grid on one scan pair: ~64 points per CTA, 4 per warp) deals blocks of 8 instead:
// with two 32-point blocks per CTA the slowest CTA's search pass took twice the average, and every CTA waited for it at the grid
// barrier of the reduction (3.8 us of a 6 us reduction).
struct Own {
  int size, rank, ns, n_local, sh;
  __device__ __forceinline__ int map(int j) const { return (((j >> sh) * size + rank) << sh) | (j & ((1 << sh) - 1)); }
};
__device__ __forceinline__ Own make_own(int ns, int size, int rank) {
  Own o;
  o.size = size; o.rank = rank; o.ns = ns;
  o.sh = (size > 16 && ns <= size * 8 * (kAlignThreads / 32)) ? 3 : 5;
  const int blocks = (ns + (1 << o.sh) - 1) >> o.sh;
  o.n_local = blocks > rank ? ((blocks - rank + size - 1) / size) << o.sh : 0;
  return o;
}

struct ChunkPlan {
  int chunk, lane, n_local;
};
__device__ __forceinline__ ChunkPlan chunk_begin(AlignShared& S, const Own& own) {
  __syncthreads();
  if (threadIdx.x == 0) S.next = 0;
  __syncthreads();
  const int n_warps = blockDim.x >> 5;
  ChunkPlan c;
  // a power of two: a group must not straddle two of the CTA's 32-point blocks (with a team they are far apart in space)
  const int want = max(1, min(1 << own.sh, (own.n_local + n_warps - 1) / n_warps));
  c.chunk = 1 << (31 - __clz(want));
  c.lane = threadIdx.x & 31;
  c.n_local = own.n_local;
  return c;
}
// returns false when the pass is over; otherwise i is this lane's point or -1 (idle lane of the chunk)
__device__ __forceinline__ bool chunk_next(AlignShared& S, const ChunkPlan& c, const Own& own, int& i) {
  int i0 = 0;
  if (c.lane == 0) i0 = atomicAdd(&S.next, c.chunk);
  i0 = __shfl_sync(0xFFFFFFFFu, i0, 0);
  if (i0 >= c.n_local) return false;
  // handed out from the end of the sorted order first
  const int j = c.n_local - 1 - i0 - c.lane;
  i = -1;
  if (c.lane < c.chunk && j >= 0) {
    const int g = own.map(j);
    if (g < own.ns) i = g;
  }
  return true;
}

// APD measurement covariance of the transformed source point (fast_apdgicp_impl.hpp:167-184)
__device__ __forceinline__ Sym3 apd_cov(float qx, float qy, float qz, const DeviceParams& P) {
  const double dist = sqrt(dadd(dadd((double)qx * (double)qx, (double)qy * (double)qy), (double)qz * (double)qz));
  const double aoa = atan2_f32(qx, fsqrt(fadd(fmul(qy, qy), fmul(qz, qz))));
  const double cosa = cos(aoa);
  const double s0 = dist * P.dist_var / 400;
  const double s1 = dist * P.sin_az / cosa;
  const double s2 = dist * P.sin_el / cosa;
  const double elev = atan2_f32(fsqrt(fadd(fmul(qx, qx), fmul(qy, qy))), qz);
  const double azim = atan2_f32(qy, qx);
  double sa, ca, se, ce;
  sincos(azim, &sa, &ca);
  sincos(elev, &se, &ce);
  // R = Rz(azim) * Ry(elev); A = R * diag(s); Cd = A A^T
  const double a00 = ca * ce * s0, a01 = -sa * s1, a02 = ca * se * s2;
  const double a10 = sa * ce * s0, a11 = ca * s1, a12 = sa * se * s2;
  const double a20 = -se * s0, a22 = ce * s2;
  Sym3 c;
  c.xx = a00 * a00 + a01 * a01 + a02 * a02;
  c.xy = a00 * a10 + a01 * a11 + a02 * a12;
  c.xz = a00 * a20 + a02 * a22;
  c.yy = a10 * a10 + a11 * a11 + a12 * a12;
  c.yz = a10 * a20 + a12 * a22;
  c.zz = a20 * a20 + a22 * a22;
  return c;
}

__device__ __forceinline__ void pose_mul(const double* a, const double* b, double* c) {  // c = a * b (isometries)
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) c[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
    c[9 + i] = a[i * 3 + 0] * b[9] + a[i * 3 + 1] * b[10] + a[i * 3 + 2] * b[11] + a[9 + i];
  }
}

// lsq_registration_impl.hpp:83-92
__device__ __forceinline__ bool is_converged(const double* delta, const DeviceParams& P) {
  double rmax = 0.0, tmax = 0.0;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) rmax = fmax(rmax, 1.0 / P.rotation_epsilon * fabs(delta[i * 3 + j] - (i == j ? 1.0 : 0.0)));
    tmax = fmax(tmax, 1.0 / P.transformation_epsilon * fabs(delta[9 + i]));
  }
  return fmax(rmax, tmax) < 1.0;
}

__device__ __forceinline__ void set_float_pose(AlignShared& S, const double* x) {
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) S.Tf[i * 4 + j] = (float)x[i * 3 + j];
    S.Tf[i * 4 + 3] = (float)x[9 + i];
  }
}

// unpack the reduced record into the symmetric 6x6 H, b and y0
__device__ __forceinline__ void unpack_record(AlignShared& S) {
  const double* r = S.red;
  double* H = S.H;
  const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      H[i * 6 + j] = r[sym[i][j]];                 // S^T M S
      H[i * 6 + 3 + j] = -r[6 + i * 3 + j];        // -S^T M
      H[(3 + j) * 6 + i] = -r[6 + i * 3 + j];      // (-S^T M)^T
      H[(3 + i) * 6 + 3 + j] = r[15 + sym[i][j]];  // M
    }
  for (int i = 0; i < 3; i++) {
    S.b[i] = r[21 + i];
    S.b[3 + i] = -r[24 + i];
  }
  S.y0 = r[27];
}

template <typename CellT>
struct TargetView {
  GridView<CellT> G;                  // grid mode (target in HBM)
  LeafView L;                         // leaf mode (target staged in shared memory, apd_leaf.cuh)
  const double2 *cov0, *cov1, *cov2;  // global, sorted target order
  const int* inv0;                    // original index -> sorted position (this cloud)
  int cloud;                          // index in B.tgt (for the coarse pyramid levels)
};

// covariances of a matched pair -> Mahalanobis matrix of the point (fast_apdgicp_impl.hpp:159-192)
template <typename CellT>
__device__ __forceinline__ void store_mahalanobis(const AlignBatch& B, const AlignShared& S, const TargetView<CellT>& T, const double2* __restrict__ c0,
                                                  const double2* __restrict__ c1, const double2* __restrict__ c2, int i, int pos, size_t sbase, float qx, float qy, float qz) {
  const double2 a0 = c0[i], a1 = c1[i], a2 = c2[i];
  const double2 b0 = T.cov0[pos], b1 = T.cov1[pos], b2 = T.cov2[pos];
  const Sym3 Cd = apd_cov(qx, qy, qz, B.prm);
  const Sym3 CA = Sym3{a0.x, a0.y, a1.x, a1.y, a2.x, a2.y} + Cd;
  const Sym3 CB = Sym3{b0.x, b0.y, b1.x, b1.y, b2.x, b2.y} + Cd;
  const Sym3 RCR = CB + rsrt(S.x0, CA);
  const Sym3 M = inverse(RCR);
  B.scratch.m0[sbase + i] = make_double2(M.xx, M.xy);
  B.scratch.m1[sbase + i] = make_double2(M.xz, M.yy);
  B.scratch.m2[sbase + i] = make_double2(M.yz, M.zz);
}

// update_correspondences on a LEAF-mode target (staged in shared memory): a warp takes 32 adjacent sorted source points and
// searches for all of them together (leaf_nn1: warp-uniform control flow, broadcast candidates). Same rules as the grid
// version below: seeded bound from the previous correspondence, anchors for points without one, strict gate d2 < dmax^2.
template <typename CellT>
__device__ __forceinline__ void correspondence_pass_leaf(const AlignBatch& B, AlignShared& S, const TargetView<CellT>& T, const float4* __restrict__ sspts,
                                                         const double2* __restrict__ c0, const double2* __restrict__ c1, const double2* __restrict__ c2,
                                                         const Own& own, size_t sbase, bool seeded) {
  const float* Tf = S.Tf;
  const float r00 = Tf[0], r01 = Tf[1], r02 = Tf[2], t0 = Tf[3];
  const float r10 = Tf[4], r11 = Tf[5], r12 = Tf[6], t1 = Tf[7];
  const float r20 = Tf[8], r21 = Tf[9], r22 = Tf[10], t2 = Tf[11];
  const float inf = __int_as_float(0x7f800000);
  const bool gated = B.prm.corr_limit2 < 3.0e38f;
  const ChunkPlan cp = chunk_begin(S, own);
  for (int i; chunk_next(S, cp, own, i);) {   // warp-uniform: every lane of the warp gets a point or -1
    const bool act = i >= 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) a = sspts[i];
    const float qx = xform_row_rn(r00, r01, r02, t0, a.x, a.y, a.z);
    const float qy = xform_row_rn(r10, r11, r12, t1, a.x, a.y, a.z);
    const float qz = xform_row_rn(r20, r21, r22, t2, a.x, a.y, a.z);
    const bool finite = isfinite(qx) && isfinite(qy) && isfinite(qz);  // a non-finite query finds nothing (NaN distances)
    const int prev = (seeded && act) ? B.scratch.corr[sbase + i] : -2;
    float bound = inf;
    bool search = act && finite, full_ring = false, anchored = false;
    float unexplored = 0.f;
    if (prev >= 0) {
      // the previous correspondence bounds the nearest neighbour (capped by the gate: beyond it nothing can match)
      const float4 t = leaf_point_gather(T.L, prev);
      bound = fminf(sqdist_rn(qx, qy, qz, t.x, t.y, t.z), B.prm.corr_limit2);
    } else if (prev == -1 && gated) {
      // ANCHOR of a point that had no correspondence: where it was searched and how far every target point is from there at
      // least. If it has moved by less than the slack between that bound and the gate, nothing can have come inside the gate.
      const float4 an = B.scratch.anchor[sbase + i];
      const float moved = sqrtf(sqdist_rn(qx, qy, qz, an.x, an.y, an.z));
      if (an.w * 0.99999f - moved * 1.00001f - 1e-3f > sqrtf(B.prm.corr_limit2) * 1.00001f) { search = false; anchored = true; }
      bound = B.prm.corr_limit2;
      unexplored = sqrtf(B.prm.corr_limit2);  // everything inside the gate radius is examined
      full_ring = true;
    } else if (gated) {
      // first pass: a little beyond the gate, so that a point without correspondence learns how far the target really is
      bound = B.prm.corr_wide2;
      unexplored = sqrtf(B.prm.corr_wide2);
      full_ring = true;
    }
    LeafTop1 v;
    v.init(bound);
    stamp(B, 10);
    leaf_nn1(T.L, qx, qy, qz, search, v, S.qslot[threadIdx.x >> 5], B.timeline ? B.timeline + 600 + (seeded ? 8 : 0) : nullptr, B.nn1_evals);
    stamp(B, 11);
    if (!act) continue;
    if (anchored) {
      B.scratch.sqd[sbase + i] = inf;
      continue;  // corr stays -1, the anchor stays valid
    }
    const bool found = search && v.found();
    const float d2 = found ? v.d2() : inf;
    const bool ok = found && (double)d2 < B.prm.corr_thr2;
    const int vpos = v.pos();
    B.scratch.corr[sbase + i] = ok ? vpos : -1;
    B.scratch.sqd[sbase + i] = d2;
    if (!ok) {
      // anchor: every target point is at least min(nearest found, radius that was examined) away; after a seeded search that
      // lost its correspondence nothing is known (bound 0: always search)
      const float lb = full_ring ? fminf(found ? sqrtf(d2) : FLT_MAX, unexplored) : 0.f;
      B.scratch.anchor[sbase + i] = make_float4(qx, qy, qz, lb);
      continue;
    }
    store_mahalanobis(B, S, T, c0, c1, c2, i, vpos, sbase, qx, qy, qz);
    stamp(B, 12);
  }
  stamp(B, 13);
  __syncthreads();
  stamp(B, 14);
}

// pcl::Registration::getFitnessScore on a LEAF-mode target: per-point squared 1-NN distances (the fixed-order sum follows in fitness_pass)
template <typename CellT>
__device__ __forceinline__ void fitness_search_leaf(const AlignBatch& B, AlignShared& S, const TargetView<CellT>& T, const float4* __restrict__ sspts, const Own& own,
                                                    size_t sbase, bool seeded) {
  const float* Tf = S.Tf;
  const float inf = __int_as_float(0x7f800000);
  const ChunkPlan cp = chunk_begin(S, own);
  for (int i; chunk_next(S, cp, own, i);) {
    const bool act = i >= 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) a = sspts[i];
    const float qx = xform_row_rn(Tf[0], Tf[1], Tf[2], Tf[3], a.x, a.y, a.z);
    const float qy = xform_row_rn(Tf[4], Tf[5], Tf[6], Tf[7], a.x, a.y, a.z);
    const float qz = xform_row_rn(Tf[8], Tf[9], Tf[10], Tf[11], a.x, a.y, a.z);
    const bool finite = isfinite(qx) && isfinite(qy) && isfinite(qz);
    const int prev = (seeded && act) ? B.scratch.corr[sbase + i] : -1;
    float bound = inf;
    if (prev >= 0) {  // seeded by the correspondence of the last linearization
      const float4 t = leaf_point_gather(T.L, prev);
      bound = sqdist_rn(qx, qy, qz, t.x, t.y, t.z);
    }
    LeafTop1 v;
    v.init(bound);
    leaf_nn1(T.L, qx, qy, qz, act && finite, v, S.qslot[threadIdx.x >> 5], nullptr, B.nn1_evals);
    if (act) B.scratch.fit[sbase + i] = (finite && v.found()) ? v.d2() : __int_as_float(0x7fc00000);  // NaN: no neighbour at all
  }
  __syncthreads();
}

// update_correspondences for the calling thread's points (fast_apdgicp_impl.hpp:146-193)
template <typename CellT>
__device__ __forceinline__ void correspondence_pass(const AlignBatch& B, AlignShared& S, const TargetView<CellT>& T, const float4* __restrict__ sspts,
                                                    const double2* __restrict__ c0, const double2* __restrict__ c1, const double2* __restrict__ c2,
                                                    const Own& own, size_t sbase, bool seeded) {
  const float* Tf = S.Tf;
  const float r00 = Tf[0], r01 = Tf[1], r02 = Tf[2], t0 = Tf[3];
  const float r10 = Tf[4], r11 = Tf[5], r12 = Tf[6], t1 = Tf[7];
  const float r20 = Tf[8], r21 = Tf[9], r22 = Tf[10], t2 = Tf[11];
  const ChunkPlan cp = chunk_begin(S, own);
  for (int i; chunk_next(S, cp, own, i);) {
    if (i < 0) continue;
    const float4 a = sspts[i];
    const float qx = xform_row_rn(r00, r01, r02, t0, a.x, a.y, a.z);
    const float qy = xform_row_rn(r10, r11, r12, t1, a.x, a.y, a.z);
    const float qz = xform_row_rn(r20, r21, r22, t2, a.x, a.y, a.z);
    Top1 v;
    v.init();
    // From the second linearization on, the previous correspondence seeds the search: its distance
    // from the new query bounds the nearest neighbour, so one pass over the rows of that ball is
    // exact (grid_ball_search); a point that had none searches the ball of the gate radius.
    const int prev = seeded ? B.scratch.corr[sbase + i] : -2;
    bool full_ring = false;
    float unexplored = 0.f;
    if (prev >= 0) {
      const float4 t = T.G.spts[prev];
      grid_ball_search(T.G, qx, qy, qz, fminf(sqdist_rn(qx, qy, qz, t.x, t.y, t.z), B.prm.corr_limit2), v);
    } else if (prev == -1 && B.prm.corr_limit2 < 3.0e38f) {
      // A point that had no correspondence keeps an ANCHOR: where it was searched and how far every
      // target point is from there at least. If it has moved by less than the slack between that
      // bound and the gate, nothing can have come inside the gate: no search at all (clutter and
      // non-overlapping regions, a quarter of a radar scan, otherwise pay the widest search every iteration).
      const float4 an = B.scratch.anchor[sbase + i];
      const float moved = sqrtf(sqdist_rn(qx, qy, qz, an.x, an.y, an.z));
      if (an.w * 0.99999f - moved * 1.00001f - 1e-3f > sqrtf(B.prm.corr_limit2) * 1.00001f) {
        B.scratch.sqd[sbase + i] = __int_as_float(0x7f800000);
        continue;  // corr stays -1, the anchor stays valid
      }
      grid_ball_search(T.G, qx, qy, qz, B.prm.corr_limit2, v);
      unexplored = sqrtf(B.prm.corr_limit2);  // everything inside the gate radius was visited
      full_ring = true;
    } else if (B.prm.corr_limit2 < 3.0e38f) {
      // searched a little beyond the gate so that a point without correspondence learns how far the target
      // really is (its anchor); the gate itself is applied below
      grid_search(T.G, qx, qy, qz, B.prm.corr_wide2, v, 0x7fffffff, &unexplored);
      full_ring = true;
    } else {
      // no gate (constructor default FLT_MAX): unbounded search through the pyramid; a hit on a coarse
      // level is mapped back to its position in the fine order
      if (pyramid_search<CellT, Top1, false>(T.G, B.tgt, T.cloud, qx, qy, qz, B.prm.corr_limit2, v) > 0 && v.pos >= 0)
        v.pos = T.inv0[v.idx];
    }
    const float d2 = v.bound2();
    const bool ok = v.pos >= 0 && (double)d2 < B.prm.corr_thr2;
    B.scratch.corr[sbase + i] = ok ? v.pos : -1;
    B.scratch.sqd[sbase + i] = d2;
    if (!ok) {
      // anchor: every target point is at least min(nearest visited, distance to the unexplored region) away;
      // after a seeded ball search that lost its correspondence nothing is known (bound 0: always search)
      const float lb = full_ring ? fminf(v.pos >= 0 ? sqrtf(d2) : FLT_MAX, unexplored) : 0.f;
      B.scratch.anchor[sbase + i] = make_float4(qx, qy, qz, lb);
      continue;
    }
    store_mahalanobis(B, S, T, c0, c1, c2, i, v.pos, sbase, qx, qy, qz);
  }
  __syncthreads();  // the accumulation passes read these records with a different (static) point-to-thread map
}

// H/b/error accumulation (FULL, fast_apdgicp_impl.hpp:221-258) or error only (:278-296) at pose x
template <bool FULL, bool LEAF, typename CellT>
__device__ __forceinline__ void accumulate_pass(const AlignBatch& B, const double* x, const TargetView<CellT>& T, const float4* __restrict__ sspts, const Own& own,
                                                size_t sbase, double (&acc)[kNRed]) {
  const double R00 = x[0], R01 = x[1], R02 = x[2], R10 = x[3], R11 = x[4], R12 = x[5], R20 = x[6], R21 = x[7], R22 = x[8];
  const double tx = x[9], ty = x[10], tz = x[11];
  // fixed point-to-thread order (whatever warp searched a point): the sums are deterministic for a launch shape
  for (int j = threadIdx.x; j < own.n_local; j += blockDim.x) {
    const int i = own.map(j);
    if (i >= own.ns) continue;
    const int c = B.scratch.corr[sbase + i];
    if (c < 0) continue;
    const float4 a = sspts[i];
    const float4 bt = LEAF ? leaf_point_gather(T.L, c) : T.G.spts[c];
    const double2 m0 = B.scratch.m0[sbase + i], m1 = B.scratch.m1[sbase + i], m2 = B.scratch.m2[sbase + i];
    const double Mxx = m0.x, Mxy = m0.y, Mxz = m1.x, Myy = m1.y, Myz = m2.x, Mzz = m2.y;
    const double ax = (double)a.x, ay = (double)a.y, az = (double)a.z;
    const double px = R00 * ax + R01 * ay + R02 * az + tx;
    const double py = R10 * ax + R11 * ay + R12 * az + ty;
    const double pz = R20 * ax + R21 * ay + R22 * az + tz;
    const double ex = (double)bt.x - px, ey = (double)bt.y - py, ez = (double)bt.z - pz;
    const double Me0 = Mxx * ex + Mxy * ey + Mxz * ez;
    const double Me1 = Mxy * ex + Myy * ey + Myz * ez;
    const double Me2 = Mxz * ex + Myz * ey + Mzz * ez;
    acc[27] += ex * Me0 + ey * Me1 + ez * Me2;
    if (FULL) {
      acc[28] += 1.0;
      // N = S^T M with S = skew(p):  S^T = [[0, pz, -py], [-pz, 0, px], [py, -px, 0]]
      const double N00 = pz * Mxy - py * Mxz, N01 = pz * Myy - py * Myz, N02 = pz * Myz - py * Mzz;
      const double N10 = px * Mxz - pz * Mxx, N11 = px * Myz - pz * Mxy, N12 = px * Mzz - pz * Mxz;
      const double N20 = py * Mxx - px * Mxy, N21 = py * Mxy - px * Myy, N22 = py * Mxz - px * Myz;
      // (N S)_i0 = N_i1 pz - N_i2 py ; (N S)_i1 = N_i2 px - N_i0 pz ; (N S)_i2 = N_i0 py - N_i1 px
      acc[0] += N01 * pz - N02 * py;
      acc[1] += N02 * px - N00 * pz;
      acc[2] += N00 * py - N01 * px;
      acc[3] += N12 * px - N10 * pz;
      acc[4] += N10 * py - N11 * px;
      acc[5] += N20 * py - N21 * px;
      acc[6] += N00; acc[7] += N01; acc[8] += N02;
      acc[9] += N10; acc[10] += N11; acc[11] += N12;
      acc[12] += N20; acc[13] += N21; acc[14] += N22;
      acc[15] += Mxx; acc[16] += Mxy; acc[17] += Mxz; acc[18] += Myy; acc[19] += Myz; acc[20] += Mzz;
      acc[21] += pz * Me1 - py * Me2;
      acc[22] += px * Me2 - pz * Me0;
      acc[23] += py * Me0 - px * Me1;
      acc[24] += Me0; acc[25] += Me1; acc[26] += Me2;
    }
  }
}

// pcl::Registration::getFitnessScore(max_range): mean squared 1-NN distance of the transformed source
template <bool LEAF, typename CellT>
__device__ __forceinline__ void fitness_pass(const AlignBatch& B, AlignShared& S, const TargetView<CellT>& T, const float4* __restrict__ sspts, const Own& own,
                                             size_t sbase, bool seeded, double (&acc)[kNRed]) {
  const float* Tf = S.Tf;
  if (LEAF) {
    fitness_search_leaf(B, S, T, sspts, own, sbase, seeded);
  } else {
    const ChunkPlan cp = chunk_begin(S, own);
    for (int i; chunk_next(S, cp, own, i);) {
      if (i < 0) continue;
      const float4 a = sspts[i];
      const float qx = xform_row_rn(Tf[0], Tf[1], Tf[2], Tf[3], a.x, a.y, a.z);
      const float qy = xform_row_rn(Tf[4], Tf[5], Tf[6], Tf[7], a.x, a.y, a.z);
      const float qz = xform_row_rn(Tf[8], Tf[9], Tf[10], Tf[11], a.x, a.y, a.z);
      Top1 v;
      v.init();
      const int prev = seeded ? B.scratch.corr[sbase + i] : -1;
      if (prev >= 0) {  // seeded by the correspondence of the last linearization
        const float4 t = T.G.spts[prev];
        grid_ball_search(T.G, qx, qy, qz, sqdist_rn(qx, qy, qz, t.x, t.y, t.z), v);
      } else {
        pyramid_search<CellT, Top1, false>(T.G, B.tgt, T.cloud, qx, qy, qz, __int_as_float(0x7f800000), v);  // only the distance is used
      }
      B.scratch.fit[sbase + i] = v.pos >= 0 ? v.bound2() : __int_as_float(0x7fc00000);  // NaN: no neighbour at all
    }
    __syncthreads();
  }
  // the sum runs in a fixed point-to-thread order, whatever warp searched the point
  for (int j = threadIdx.x; j < own.n_local; j += blockDim.x) {
    const int i = own.map(j);
    if (i >= own.ns) continue;
    const float d2 = B.scratch.fit[sbase + i];
    if ((double)d2 <= B.max_range) {
      acc[0] += (double)d2;
      acc[1] += 1.0;
    }
  }
}

template <int TEAM, bool STAGED>
__global__ void __launch_bounds__(kAlignThreads, APD_ALIGN_MIN_BLOCKS) align_kernel(const __grid_constant__ AlignBatch B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ AlignShared S;
  typedef unsigned CellT;  // grid mode reads the 32-bit cell table from HBM; leaf mode (STAGED) has no cell table at all

  TeamCtx<TEAM> tc;
  tc.buf = 0;
  tc.grid_partials = B.grid_partials;
  tc.tl = B.timeline;
  if (TEAM == TEAM_CTA) {
    tc.size = 1; tc.rank = 0; tc.id = blockIdx.x; tc.count = gridDim.x;
  } else if (TEAM == TEAM_CLUSTER) {
    cg::cluster_group cl = cg::this_cluster();
    tc.size = (int)cl.num_blocks(); tc.rank = (int)cl.block_rank(); tc.id = blockIdx.x / tc.size; tc.count = gridDim.x / tc.size;
  } else {
    tc.size = gridDim.x; tc.rank = blockIdx.x; tc.id = 0; tc.count = 1;
  }
  const bool leader = (tc.rank == 0 && threadIdx.x == 0);
  const DeviceParams& P = B.prm;
  __shared__ unsigned long long s_stage_bar;   // completion barrier of the bulk staging copies (leaf_stage_bulk)
  unsigned stage_parity = 0u;
  if (threadIdx.x == 0) {
    S.staged_target = -1;
    if (STAGED) mbar_init(&s_stage_bar, 1u);
  }
  __syncthreads();

  int pair = tc.id;
  stamp(B, 0);  // kernel entered
  for (;;) {
    if (TEAM == TEAM_CTA) {
      __syncthreads();
      if (threadIdx.x == 0) S.pair = atomicAdd(B.work_counter, 1);
      __syncthreads();
      pair = S.pair;
    }
    if (pair >= B.n_pairs) break;
    const int s = B.src_idx ? B.src_idx[pair] : pair + B.src_base;
    const int t = B.tgt_idx ? B.tgt_idx[pair] : pair + B.tgt_base;
    const int sb = B.src.pt_off[s], ns = B.src.pt_off[s + 1] - sb;
    const int tb = B.tgt.pt_off[t], nt = B.tgt.pt_off[t + 1] - tb;
    const float4* sspts = B.src.spts + sb;
    const double2 *c0 = B.src.cov0 + sb, *c1 = B.src.cov1 + sb, *c2 = B.src.cov2 + sb;
    const size_t sbase = (size_t)tc.id * B.scratch.max_src;

    TargetView<CellT> T;
    if (!STAGED) T.G.g = B.tgt.grid[t];
    T.G.n = nt;
    T.cov0 = B.tgt.cov0 + tb; T.cov1 = B.tgt.cov1 + tb; T.cov2 = B.tgt.cov2 + tb;
    T.inv0 = B.tgt.inv0 + tb;
    T.cloud = t;
    if (STAGED) {
      // leaf mode: the target's Hilbert-sorted points (pair layout) and leaf boxes live in shared memory for the whole pair
      float4* sP = reinterpret_cast<float4*>(smem_raw);
      const int nleaf = (nt + kLeaf - 1) / kLeaf;
      float4* sbox = sP + (size_t)nleaf * kLeaf;
      if (S.staged_target != t) {  // consecutive pairs on the same target (scan-to-submap) reuse the staged cloud
        if (B.tgt.limg) {
          // one bulk asynchronous copy of the image the build wrote (points in the pair layout + boxes): no register round trip
          leaf_stage_bulk(sP, B.tgt.limg + (size_t)kLeafImage * B.tgt.leaf_off[t], nt, &s_stage_bar, stage_parity);
        } else {
          leaf_stage(sP, sbox, B.tgt.spts + tb, B.tgt.lbox + 2 * (size_t)B.tgt.leaf_off[t], nt);
        }
        __syncthreads();
        if (threadIdx.x == 0) S.staged_target = t;
      }
      T.L.P = sP;
      T.L.box = sbox;
      T.L.n = nt;
      T.L.nleaf = nleaf;
    } else {
      T.G.spts = B.tgt.spts + tb;
      T.G.cells = reinterpret_cast<const CellT*>(B.tgt.cells + B.tgt.cell_off[t]);
    }

    stamp(B, 1);  // target staged
    const Own own = make_own(ns, tc.size, tc.rank);

    if (threadIdx.x == 0) {
      // x0 = Isometry3d(guess.cast<double>())  (lsq_registration_impl.hpp:56)
      if (B.guesses64) {
        const double* g = B.guesses64 + (size_t)pair * 16;
        for (int i = 0; i < 3; i++) {
          for (int j = 0; j < 3; j++) S.x0[i * 3 + j] = g[i * 4 + j];
          S.x0[9 + i] = g[i * 4 + 3];
        }
      } else if (B.guesses) {
        const float* g = B.guesses + (size_t)pair * 16;
        for (int i = 0; i < 3; i++) {
          for (int j = 0; j < 3; j++) S.x0[i * 3 + j] = (double)g[i * 4 + j];
          S.x0[9 + i] = (double)g[i * 4 + 3];
        }
      } else {
        for (int i = 0; i < 12; i++) S.x0[i] = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
      }
      S.lambda = -1.0;
      S.converged = 0;
      set_float_pose(S, S.x0);
    }
    __syncthreads();

    int iterations = 0, status = APD_OK, n_trace = 0;
    double last_y0 = 0.0, last_inl = 0.0;
    double acc[kNRed];
    bool have_input = ns > 0 && nt > 0;
    if (!have_input) status = APD_ERR_NO_INPUT;
    else if (ns < B.min_points || nt < B.min_points) {  // the reference reads uninitialised columns here (fast_apdgicp_impl.hpp:318-321)
      have_input = false;
      status = APD_ERR_TOO_FEW_POINTS;
    }

    if (have_input && B.mode == 3) {
      // ---- compute_error(x0) alone: stale correspondences and Mahalanobis of the last linearize (fast_apdgicp_impl.hpp:275-298) ----
#pragma unroll
      for (int i = 0; i < kNRed; i++) acc[i] = 0.0;
      accumulate_pass<false, STAGED>(B, S.x0, T, sspts, own, sbase, acc);
      acc[0] = acc[27];
      team_reduce<TEAM, 1>(acc, S, tc);
      last_y0 = S.red[0];
      __syncthreads();
    }
    for (int it = 0; have_input && it < (B.mode == 1 ? 1 : (B.mode >= 2 ? 0 : P.max_iterations)); it++) {
      iterations = it;
      // ---- linearize(x0) ----
      stamp(B, 2);  // iteration starts
      if (STAGED) correspondence_pass_leaf(B, S, T, sspts, c0, c1, c2, own, sbase, it > 0);
      else correspondence_pass(B, S, T, sspts, c0, c1, c2, own, sbase, it > 0);
      stamp(B, 3);  // correspondences + Mahalanobis done
#pragma unroll
      for (int i = 0; i < kNRed; i++) acc[i] = 0.0;
      accumulate_pass<true, STAGED>(B, S.x0, T, sspts, own, sbase, acc);
      stamp(B, 20);  // H, b partials accumulated
      team_reduce<TEAM, 29>(acc, S, tc);
      stamp(B, 4);  // H, b reduced
      if (threadIdx.x == 0) unpack_record(S);
      if (leader) atomicAdd(&B.counters[0], 1ull);
      __syncthreads();
      stamp(B, 23);  // record unpacked
      last_y0 = S.y0;
      last_inl = S.red[28];
      if (B.mode == 1) break;

      bool step_ok = false;
      if (P.optimizer == APD_OPT_GAUSS_NEWTON) {
        // ---- step_gn (lsq_registration_impl.hpp:107-123) ----
        if (threadIdx.x == 0) {
          ldlt6_solve(S.H, S.b, S.d, 0.0, -1.0);
          so3_exp_matrix(S.d, S.delta);
          S.delta[9] = S.d[3]; S.delta[10] = S.d[4]; S.delta[11] = S.d[5];
          pose_mul(S.delta, S.x0, S.xi);
          for (int i = 0; i < 12; i++) S.x0[i] = S.xi[i];
          set_float_pose(S, S.x0);
          S.converged = is_converged(S.delta, P) ? 1 : 0;
          if (tc.rank == 0 && B.final_hessian)
            for (int i = 0; i < 36; i++) B.final_hessian[(size_t)pair * 36 + i] = S.H[i];
        }
        __syncthreads();
        step_ok = true;
      } else {
        // ---- step_lm (lsq_registration_impl.hpp:127-173) ----
        if (threadIdx.x == 0) {
          if (S.lambda < 0.0) {
            double mx = 0.0;
            for (int i = 0; i < 6; i++) mx = fmax(mx, fabs(S.H[i * 6 + i]));
            S.lambda = P.lm_init_lambda_factor * mx;
          }
          S.nu = 2.0;
        }
        for (int li = 0; li < P.lm_max_iterations; li++) {
          if (threadIdx.x == 0) {
            ldlt6_solve(S.H, S.b, S.d, S.lambda, -1.0);
            stamp(B, 24);  // LDLT done
            so3_exp_matrix(S.d, S.delta);
            stamp(B, 25);  // so3_exp done
            S.delta[9] = S.d[3]; S.delta[10] = S.d[4]; S.delta[11] = S.d[5];
            pose_mul(S.delta, S.x0, S.xi);
          }
          stamp(B, 21);  // damped system solved
          __syncthreads();
          // ---- compute_error(xi): stale correspondences and Mahalanobis ----
#pragma unroll
          for (int i = 0; i < kNRed; i++) acc[i] = 0.0;
          accumulate_pass<false, STAGED>(B, S.xi, T, sspts, own, sbase, acc);
          acc[0] = acc[27];
          stamp(B, 22);  // error partials accumulated
          team_reduce<TEAM, 1>(acc, S, tc);
          stamp(B, 5);  // LM solve + error pass reduced
          if (leader) atomicAdd(&B.counters[1], 1ull);
          if (threadIdx.x == 0) {
            const double yi = S.red[0];
            double denom = 0.0, dn = 0.0;
            for (int r = 0; r < 6; r++) {
              denom += S.d[r] * (S.lambda * S.d[r] - S.b[r]);
              dn += S.d[r] * S.d[r];
            }
            const double rho = (S.y0 - yi) / denom;
            const bool reject = rho < 0.0;
            if (tc.rank == 0 && B.trace && n_trace < B.trace_rows) {
              double* row = B.trace + ((size_t)pair * B.trace_rows + n_trace) * 8;
              row[0] = (double)it; row[1] = (double)li; row[2] = S.y0; row[3] = yi; row[4] = rho; row[5] = S.lambda; row[6] = sqrt(dn);
              row[7] = reject ? 0.0 : 1.0;
            }
            if (reject) {
              if (is_converged(S.delta, P)) {
                S.decision = 2;  // success, x0 NOT updated (:156-159)
              } else {
                S.lambda = S.nu * S.lambda;
                S.nu = 2.0 * S.nu;
                S.decision = 0;
              }
            } else {
              for (int i = 0; i < 12; i++) S.x0[i] = S.xi[i];
              set_float_pose(S, S.x0);
              const double f = 2.0 * rho - 1.0;
              S.lambda = S.lambda * fmax(1.0 / 3.0, 1.0 - f * f * f);
              if (tc.rank == 0 && B.final_hessian)
                for (int i = 0; i < 36; i++) B.final_hessian[(size_t)pair * 36 + i] = S.H[i];
              S.decision = 1;
            }
          }
          n_trace++;
          __syncthreads();
          if (S.decision != 0) { step_ok = true; break; }
        }
        if (step_ok && threadIdx.x == 0) S.converged = is_converged(S.delta, P) ? 1 : 0;
        __syncthreads();
      }
      if (!step_ok) {  // "lm not converged!!" (:71-74): converged_ stays false
        status = APD_STATUS_LM_FAILED;
        break;
      }
      if (S.converged) break;
    }

    // ---- final_transformation_ = x0.cast<float>() (:78) and getFitnessScore ----
#pragma unroll
    for (int i = 0; i < kNRed; i++) acc[i] = 0.0;
    if (have_input && B.mode != 1 && B.mode != 3) fitness_pass<STAGED>(B, S, T, sspts, own, sbase, B.mode == 0 && P.max_iterations > 0, acc);
    team_reduce<TEAM, 2>(acc, S, tc);
    stamp(B, 6);  // fitness done
    if (leader) {
      apd_result r;
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 4; j++) r.T[i * 4 + j] = S.Tf[i * 4 + j];
      }
      r.T[12] = 0.f; r.T[13] = 0.f; r.T[14] = 0.f; r.T[15] = 1.f;
      if (!have_input)
        for (int i = 0; i < 16; i++) r.T[i] = (i % 5 == 0) ? 1.f : 0.f;  // pcl align returns with final_transformation_ = I
      r.fitness = (S.red[1] > 0.0) ? S.red[0] / S.red[1] : DBL_MAX;
      r.error = last_y0;
      r.converged = have_input ? S.converged : 0;
      r.iterations = iterations;
      r.status = status;
      r.num_inliers = (int)last_inl;
      B.out[pair] = r;
      if (B.trace_count) B.trace_count[pair] = min(n_trace, B.trace_rows);
      if (B.mode == 1) {
        if (B.final_hessian)
          for (int i = 0; i < 36; i++) B.final_hessian[(size_t)pair * 36 + i] = S.H[i];
        if (B.lin_b)
          for (int i = 0; i < 6; i++) B.lin_b[(size_t)pair * 6 + i] = S.b[i];
      }
    }
    __syncthreads();
    if (TEAM != TEAM_CTA) pair += tc.count;
  }
  if (TEAM == TEAM_CLUSTER) cg::this_cluster().sync();  // peers may still be reading this CTA's partials
}

template <int TEAM, bool STAGED>
cudaError_t launch_t(const AlignBatch& b, int team_size, int n_teams, size_t smem_bytes, cudaStream_t stream) {
  auto kern = align_kernel<TEAM, STAGED>;
  cudaError_t e = ensure_dynamic_smem(kern, smem_bytes);
  if (e != cudaSuccess) return e;
  if (TEAM == TEAM_CTA) {
    kern<<<n_teams, kAlignThreads, smem_bytes, stream>>>(b);
    return cudaGetLastError();
  }
  if (TEAM == TEAM_CLUSTER) {
    if (team_size > 8) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_teams * team_size);
    cfg.blockDim = dim3(kAlignThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, b);
  }
  // TEAM_GRID: cooperative launch, team_size CTAs
  void* args[] = {(void*)&b};
  return cudaLaunchCooperativeKernel((void*)kern, dim3(team_size), dim3(kAlignThreads), args, smem_bytes, stream);
}

template <int TEAM, bool STAGED>
int max_teams_t(int team_size, size_t smem_bytes) {
  auto kern = align_kernel<TEAM, STAGED>;
  if (ensure_dynamic_smem(kern, smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (TEAM == TEAM_CLUSTER) {
    if (team_size > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(team_size);
    cfg.blockDim = dim3(kAlignThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return n;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAlignThreads, smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return per_sm * sms;  // TEAM_CTA: concurrent teams; TEAM_GRID: the largest cooperative grid
}

}  // namespace

size_t align_static_smem() { return sizeof(AlignShared); }

#define APD_DISPATCH(FN, ...)                                                                  \
  (team_kind == TEAM_CTA ? (stage_target ? FN<TEAM_CTA, true>(__VA_ARGS__) : FN<TEAM_CTA, false>(__VA_ARGS__))            \
   : team_kind == TEAM_CLUSTER ? (stage_target ? FN<TEAM_CLUSTER, true>(__VA_ARGS__) : FN<TEAM_CLUSTER, false>(__VA_ARGS__)) \
                               : (stage_target ? FN<TEAM_GRID, true>(__VA_ARGS__) : FN<TEAM_GRID, false>(__VA_ARGS__)))

cudaError_t launch_align(const AlignBatch& b, int team_kind, int team_size, int n_teams, bool stage_target, size_t smem_bytes, cudaStream_t stream,
                         LaunchStats* st) {
  if (b.n_pairs == 0) return cudaSuccess;
  if (st) st->launches++;
  return APD_DISPATCH(launch_t, b, team_size, n_teams, smem_bytes, stream);
}

int align_max_teams(int team_kind, int team_size, bool stage_target, size_t smem_bytes) { return APD_DISPATCH(max_teams_t, team_size, smem_bytes); }

// ---- stand-alone getFitnessScore(max_range) for an arbitrary transform ----
namespace {

__global__ void __launch_bounds__(256) fitness_kernel(CloudSetView src, int s, CloudSetView tgt, int t, const float* __restrict__ Tf, double max_range, bool strict,
                                                      double* __restrict__ partials) {
  __shared__ double ws[8][2];
  GridView<unsigned> G;
  const int tb = tgt.pt_off[t];
  G.g = tgt.grid[t];
  G.n = tgt.pt_off[t + 1] - tb;
  G.spts = tgt.spts + tb;
  G.cells = tgt.cells + tgt.cell_off[t];
  const int sb = src.pt_off[s], ns = src.pt_off[s + 1] - sb;
  double sum = 0.0, cnt = 0.0;
  if (G.n > 0) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
      const float4 a = src.spts[sb + i];
      const float qx = xform_row_rn(Tf[0], Tf[1], Tf[2], Tf[3], a.x, a.y, a.z);
      const float qy = xform_row_rn(Tf[4], Tf[5], Tf[6], Tf[7], a.x, a.y, a.z);
      const float qz = xform_row_rn(Tf[8], Tf[9], Tf[10], Tf[11], a.x, a.y, a.z);
      Top1 v;
      v.init();
      pyramid_search<unsigned, Top1, false>(G, tgt, t, qx, qy, qz, __int_as_float(0x7f800000), v);
      if (v.pos >= 0 && (strict ? (double)v.bound2() < max_range : (double)v.bound2() <= max_range)) {
        sum += (double)v.bound2();
        cnt += 1.0;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5][0] = sum; ws[threadIdx.x >> 5][1] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) { a += ws[w][0]; b += ws[w][1]; }
    partials[blockIdx.x * 2] = a;
    partials[blockIdx.x * 2 + 1] = b;
  }
}

// the same on a LEAF-mode target: every CTA stages the target (shared memory), its warps take 32 sorted source points at a time
__global__ void __launch_bounds__(256) fitness_leaf_kernel(CloudSetView src, int s, CloudSetView tgt, int t, const float* __restrict__ Tf, double max_range, bool strict,
                                                           double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double ws[8][2];
  __shared__ float4 qslot[8][32];
  const int tb = tgt.pt_off[t], nt = tgt.pt_off[t + 1] - tb;
  const int sb = src.pt_off[s], ns = src.pt_off[s + 1] - sb;
  LeafView L;
  L.n = nt;
  L.nleaf = (nt + kLeaf - 1) / kLeaf;
  float4* sP = reinterpret_cast<float4*>(smem_raw);
  float4* sbox = sP + (size_t)L.nleaf * kLeaf;
  __shared__ unsigned long long s_stage_bar;
  if (tgt.limg) {
    unsigned parity = 0u;
    if (threadIdx.x == 0) mbar_init(&s_stage_bar, 1u);
    __syncthreads();
    leaf_stage_bulk(sP, tgt.limg + (size_t)kLeafImage * tgt.leaf_off[t], nt, &s_stage_bar, parity);
  } else {
    leaf_stage(sP, sbox, tgt.spts + tb, tgt.lbox + 2 * (size_t)tgt.leaf_off[t], nt);
  }
  L.P = sP;
  L.box = sbox;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double sum = 0.0, cnt = 0.0;
  if (nt > 0) {
    for (int i0 = (blockIdx.x * nw + warp) * 32; i0 < ns; i0 += gridDim.x * nw * 32) {
      const int i = i0 + lane;
      const bool act = i < ns;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) a = src.spts[sb + i];
      const float qx = xform_row_rn(Tf[0], Tf[1], Tf[2], Tf[3], a.x, a.y, a.z);
      const float qy = xform_row_rn(Tf[4], Tf[5], Tf[6], Tf[7], a.x, a.y, a.z);
      const float qz = xform_row_rn(Tf[8], Tf[9], Tf[10], Tf[11], a.x, a.y, a.z);
      const bool valid = act && isfinite(qx) && isfinite(qy) && isfinite(qz);
      LeafTop1 v;
      v.init(__int_as_float(0x7f800000));
      leaf_nn1(L, qx, qy, qz, valid, v, qslot[warp]);
      if (valid && v.found() && (strict ? (double)v.d2() < max_range : (double)v.d2() <= max_range)) {
        sum += (double)v.d2();
        cnt += 1.0;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
  }
  if (lane == 0) { ws[warp][0] = sum; ws[warp][1] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < nw; w++) { a += ws[w][0]; b += ws[w][1]; }
    partials[blockIdx.x * 2] = a;
    partials[blockIdx.x * 2 + 1] = b;
  }
}

__global__ void fitness_final_kernel(const double* __restrict__ partials, int blocks, double* __restrict__ out) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < blocks; i++) { a += partials[i * 2]; b += partials[i * 2 + 1]; }
  out[0] = (b > 0.0) ? a / b : DBL_MAX;
  out[1] = b;
}

// one CTA: a scan is a few thousand values, and the count has to be exact (integer), so no atomics on doubles
__global__ void __launch_bounds__(1024) count_below_kernel(const float* __restrict__ d2, int n, double thr, double* __restrict__ out) {
  __shared__ int ws[32];
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += ((double)d2[i] < thr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += ws[w];
    out[0] = (double)t;
  }
}

}  // namespace

cudaError_t launch_count_below(const float* d2, int n, double thr, double* out, cudaStream_t stream, LaunchStats* st) {
  count_below_kernel<<<1, 1024, 0, stream>>>(d2, n, thr, out);
  if (st) st->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fitness(const CloudSetView& src, int s, const CloudSetView& tgt, int t, const float* T16, double max_range, bool strict, double* partials,
                           int blocks, double* out, cudaStream_t stream, LaunchStats* st) {
  if (tgt.lbox) {  // leaf-mode target: staged per CTA, so few CTAs (the staging copy is the fixed cost)
    const size_t smem = (size_t)((tgt.total_points + kLeaf - 1) / kLeaf + tgt.n_clouds) * (kLeaf * 16 + 32);  // >= the largest cloud of the set
    cudaError_t e = ensure_dynamic_smem(fitness_leaf_kernel, smem);
    if (e != cudaSuccess) return e;
    fitness_leaf_kernel<<<blocks, 256, smem, stream>>>(src, s, tgt, t, T16, max_range, strict, partials);
  } else {
    fitness_kernel<<<blocks, 256, 0, stream>>>(src, s, tgt, t, T16, max_range, strict, partials);
  }
  if (st) st->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fitness_final_kernel<<<1, 1, 0, stream>>>(partials, blocks, out);
  if (st) st->launches++;
  return cudaGetLastError();
}

}  // namespace apd
