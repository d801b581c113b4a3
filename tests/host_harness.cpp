// TEST INFRASTRUCTURE — compiles the product's search/math headers (riv-slam_b200/csrc/apd_grid.cuh,
// apd_math.cuh) as plain C++ so the exact ring-expansion search, the Jacobi eigen-solver, the LDL^T
// solve and so3_exp can be unit-tested on a machine without a GPU. Nothing here is shipped or
// called by the product; the GPU parity tests remain the real gate.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC tests/host_harness.cpp -o tests/_host_harness.so
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../riv-slam_b200/csrc/apd_grid.cuh"

using namespace apd;

namespace {

struct HostGrid {
  std::vector<float4> spts;
  std::vector<unsigned> cells;
  GridParams g;
};

// same rules as grid_params_kernel / count / scan / scatter / cell_sort in apd_build.cu
HostGrid build(const float* xyz, int n, int cell_cap) {
  HostGrid G;
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  bool any = false;
  for (int i = 0; i < n; i++) {
    const float* p = xyz + 3 * i;
    if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
    for (int a = 0; a < 3; a++) {
      if (!any) { lo[a] = hi[a] = p[a]; }
      lo[a] = std::min(lo[a], p[a]);
      hi[a] = std::max(hi[a], p[a]);
    }
    any = true;
  }
  const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  const long long cap = std::max(cell_cap, 1);
  const float emax = std::max(std::max(ex, ey), ez);
  const float floor_h = std::max(emax * 1e-4f, 1e-6f);
  const float dx = std::max(ex, floor_h), dy = std::max(ey, floor_h), dz = std::max(ez, floor_h);
  float h = std::max(cbrtf(dx * dy * dz / (float)cap), floor_h);
  int nx = 1, ny = 1, nz = 1;
  for (int it = 0; it < 4096; it++) {
    const float fx = floorf(ex / h) + 1.f, fy = floorf(ey / h) + 1.f, fz = floorf(ez / h) + 1.f;
    if (fx * fy * fz <= (float)cap && fx < 2e9f && fy < 2e9f && fz < 2e9f) {
      nx = (int)fx; ny = (int)fy; nz = (int)fz;
      if ((long long)nx * ny * nz <= cap) break;
    }
    h *= 1.02f;
  }
  GridParams& g = G.g;
  g.lox = lo[0]; g.loy = lo[1]; g.loz = lo[2];
  g.h = h; g.inv_h = 1.0f / h;
  g.nx = nx; g.ny = ny; g.nz = nz; g.ncells = nx * ny * nz;
  float amax = 0.f;
  for (int a = 0; a < 3; a++) amax = std::max(amax, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
  g.slack = 4e-6f * (amax + emax + h) + 1e-30f;
  std::vector<int> cid(n);
  G.cells.assign(g.ncells + 1, 0);
  for (int i = 0; i < n; i++) {
    cid[i] = cell_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    G.cells[cid[i] + 1]++;
  }
  for (int c = 0; c < g.ncells; c++) G.cells[c + 1] += G.cells[c];
  std::vector<unsigned> cur(G.cells.begin(), G.cells.end() - 1);
  G.spts.resize(n);
  for (int i = 0; i < n; i++) G.spts[cur[cid[i]]++] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], u2f((unsigned)i));  // ascending index per cell
  return G;
}

template <int K>
void knn_t(const HostGrid& HG, const float* q, int nq, int k, int* idx, float* d2) {
  GridView<unsigned> G{HG.spts.data(), HG.cells.data(), HG.g, (int)HG.spts.size()};
  for (int i = 0; i < nq; i++) {
    TopK<K> tk;
    tk.init();
    grid_search(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], INFINITY, tk);
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = (int)(unsigned)(tk.key[j] & 0xFFFFFFFFull);
      d2[(size_t)i * k + j] = u2f((unsigned)(tk.key[j] >> 32));
    }
  }
}

// the kNN kernel's per-thread schedule: the first query by ring expansion, the following ones inside
// the ball given by the triangle inequality from the previous result (chunks of `chunk` queries)
template <int K>
void knn_chained_t(const HostGrid& HG, int chunk, int k, int* idx, float* d2) {
  GridView<unsigned> G{HG.spts.data(), HG.cells.data(), HG.g, (int)HG.spts.size()};
  const int n = (int)HG.spts.size();
  for (int q0 = 0; q0 < n; q0 += chunk) {
    float px = 0, py = 0, pz = 0, prk2 = -1.f;
    for (int q = q0; q < std::min(n, q0 + chunk); q++) {
      const float4 p = HG.spts[q];
      TopK<K> tk;
      tk.init();
      if (prk2 >= 0.f) grid_ball_search(G, p.x, p.y, p.z, chained_bound2(prk2, sqdist_rn(p.x, p.y, p.z, px, py, pz)), tk);
      else grid_search(G, p.x, p.y, p.z, INFINITY, tk);
      px = p.x; py = p.y; pz = p.z; prk2 = tk.bound2();
      const unsigned self = f2u(p.w);
      for (int j = 0; j < k; j++) {
        idx[(size_t)self * k + j] = (int)(unsigned)(tk.key[j] & 0xFFFFFFFFull);
        d2[(size_t)self * k + j] = u2f((unsigned)(tk.key[j] >> 32));
      }
    }
  }
}

// the grid pyramid of apd_internal.h (pyramid_search): fine grid for `rings` rings, then cap/64, then cap/4096
template <int K>
int knn_pyramid_t(const HostGrid* L, int rings, const float* q, int nq, int k, int* idx, float* d2) {
  int coarse_used = 0;
  for (int i = 0; i < nq; i++) {
    TopK<K> tk;
    tk.init();
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    int level = 0;
    for (; level < 3; level++) {
      GridView<unsigned> G{L[level].spts.data(), L[level].cells.data(), L[level].g, (int)L[level].spts.size()};
      if (level) tk.init();
      if (grid_search(G, qx, qy, qz, INFINITY, tk, level < 2 ? rings : 0x7fffffff)) break;
    }
    coarse_used += level > 0;
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = (int)(unsigned)(tk.key[j] & 0xFFFFFFFFull);
      d2[(size_t)i * k + j] = u2f((unsigned)(tk.key[j] >> 32));
    }
  }
  return coarse_used;
}

// the kNN kernel's fast path: packed 32-bit candidate list, completeness check, exact fix-up, else exact redo
template <int K>
int knn_packed_t(const HostGrid& HG, const std::vector<float4>& opts, const float* q, int nq, int k, int* idx, float* d2) {
  GridView<unsigned> G{HG.spts.data(), HG.cells.data(), HG.g, (int)HG.spts.size()};
  int fallbacks = 0;
  for (int i = 0; i < nq; i++) {
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    TopKPacked<K, K + 4> ap;
    ap.setup(G.n);
    ap.init();
    grid_search(G, qx, qy, qz, INFINITY, ap);
    TopK<K> tk;
    if (ap.complete()) {
      exact_from_packed(ap, qx, qy, qz, opts.data(), tk);
    } else {
      fallbacks++;
      tk.init();
      grid_search(G, qx, qy, qz, INFINITY, tk);
    }
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = (int)(unsigned)(tk.key[j] & 0xFFFFFFFFull);
      d2[(size_t)i * k + j] = u2f((unsigned)(tk.key[j] >> 32));
    }
  }
  return fallbacks;
}

}  // namespace

extern "C" {

// kNN with the packed-key fast path; returns the number of queries that fell back to the exact list
int hh_knn_packed(const float* cloud_xyz, int n, int cell_cap, const float* q, int nq, int k, int* idx, float* d2) {
  const HostGrid G = build(cloud_xyz, n, cell_cap);
  std::vector<float4> opts(n);
  for (int i = 0; i < n; i++) opts[i] = make_float4(cloud_xyz[3 * i], cloud_xyz[3 * i + 1], cloud_xyz[3 * i + 2], 1.f);
  if (k == 10) return knn_packed_t<10>(G, opts, q, nq, k, idx, d2);
  if (k == 20) return knn_packed_t<20>(G, opts, q, nq, k, idx, d2);
  return -1;
}

// kNN through the three-level pyramid; returns how many queries needed a coarse level
int hh_knn_pyramid(const float* cloud_xyz, int n, int cell_cap, int rings, const float* q, int nq, int k, int* idx, float* d2) {
  const HostGrid L[3] = {build(cloud_xyz, n, cell_cap), build(cloud_xyz, n, std::max(8, cell_cap >> 6)), build(cloud_xyz, n, std::max(8, cell_cap >> 12))};
  if (k <= 10) return knn_pyramid_t<10>(L, rings, q, nq, k, idx, d2);
  if (k <= 20) return knn_pyramid_t<20>(L, rings, q, nq, k, idx, d2);
  return knn_pyramid_t<32>(L, rings, q, nq, k, idx, d2);
}

// self-kNN of a cloud with the chained-ball schedule; rows in ORIGINAL point order
void hh_knn_chained(const float* cloud_xyz, int n, int cell_cap, int chunk, int k, int* idx, float* d2) {
  const HostGrid G = build(cloud_xyz, n, cell_cap);
  if (k <= 10) knn_chained_t<10>(G, chunk, k, idx, d2);
  else if (k <= 20) knn_chained_t<20>(G, chunk, k, idx, d2);
  else knn_chained_t<32>(G, chunk, k, idx, d2);
}

// seeded 1-NN: ball of radius |q - cloud[seed]| (seed < 0: ball of radius sqrt(limit2))
void hh_nn1_seeded(const float* cloud_xyz, int n, int cell_cap, const float* q, int nq, const int* seed, float limit2, int* idx, float* d2) {
  const HostGrid HG = build(cloud_xyz, n, cell_cap);
  GridView<unsigned> G{HG.spts.data(), HG.cells.data(), HG.g, n};
  for (int i = 0; i < nq; i++) {
    Top1 v;
    v.init();
    float B = limit2;
    if (seed[i] >= 0) B = sqdist_rn(q[3 * i], q[3 * i + 1], q[3 * i + 2], cloud_xyz[3 * seed[i]], cloud_xyz[3 * seed[i] + 1], cloud_xyz[3 * seed[i] + 2]);
    grid_ball_search(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], B, v);
    idx[i] = v.pos >= 0 ? (int)f2u(HG.spts[v.pos].w) : -1;
    d2[i] = v.bound2();
  }
}

int hh_knn(const float* cloud_xyz, int n, int cell_cap, const float* q, int nq, int k, int* idx, float* d2) {
  const HostGrid G = build(cloud_xyz, n, cell_cap);
  if (k <= 10) knn_t<10>(G, q, nq, k, idx, d2);
  else if (k <= 20) knn_t<20>(G, q, nq, k, idx, d2);
  else knn_t<32>(G, q, nq, k, idx, d2);
  return G.g.ncells;
}

// bounded 1-NN: idx = -1 when nothing lies within limit2 (the search may stop early)
void hh_nn1(const float* cloud_xyz, int n, int cell_cap, const float* q, int nq, float limit2, int* idx, float* d2) {
  const HostGrid HG = build(cloud_xyz, n, cell_cap);
  GridView<unsigned> G{HG.spts.data(), HG.cells.data(), HG.g, n};
  for (int i = 0; i < nq; i++) {
    Top1 v;
    v.init();
    grid_search(G, q[3 * i], q[3 * i + 1], q[3 * i + 2], limit2, v);
    idx[i] = v.pos >= 0 ? (int)f2u(HG.spts[v.pos].w) : -1;
    d2[i] = v.bound2();
  }
}

void hh_sym_eig3(const double* c6, double* w3, double* V9) {
  Sym3 A{c6[0], c6[1], c6[2], c6[3], c6[4], c6[5]};
  sym_eig3(A, w3, V9);
}

void hh_ldlt6(const double* A36, const double* rhs, double* x) {
  double A[36];
  memcpy(A, A36, sizeof(A));
  ldlt6_solve(A, rhs, x);
}

void hh_so3_exp(const double* w, double* R9) { so3_exp_matrix(w, R9); }
}
