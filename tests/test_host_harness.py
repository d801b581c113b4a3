"""CPU unit tests of the product's search and small-solver headers compiled as plain C++
(tests/host_harness.cpp): the exact ring-expansion grid search against brute force, and the
fp64 helpers against numpy. The GPU runs the same source through nvcc."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle.oracle import knn_bruteforce

_fp, _ip, _dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)


def load_harness():
    """Compile (when stale) and load tests/_host_harness.so: the product's search header as plain C++."""
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    out = os.path.join(ROOT, "tests", "_host_harness.so")
    deps = [src] + [os.path.join(ROOT, "riv-slam_b200", "csrc", f) for f in ("apd_grid.cuh", "apd_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", out], check=True)
    L = C.CDLL(out)
    L.hh_knn.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _ip, _fp]
    L.hh_nn1.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_float, _ip, _fp]
    L.hh_knn_chained.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _fp]
    L.hh_nn1_seeded.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, _ip, C.c_float, _ip, _fp]
    L.hh_sym_eig3.argtypes = [_dp, _dp, _dp]
    L.hh_ldlt6.argtypes = [_dp, _dp, _dp]
    L.hh_so3_exp.argtypes = [_dp, _dp]
    return L


@pytest.fixture(scope="module")
def hh():
    return load_harness()


def _knn(L, cloud, q, k, cap):
    c = np.ascontiguousarray(cloud[:, :3], np.float32)
    qq = np.ascontiguousarray(q[:, :3], np.float32)
    idx = np.zeros((len(qq), k), np.int32)
    d2 = np.zeros((len(qq), k), np.float32)
    L.hh_knn(c.ctypes.data_as(_fp), len(c), cap, qq.ctypes.data_as(_fp), len(qq), k, idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
    return idx, d2


@pytest.mark.parametrize("cap", [1, 64, 3000, 24000, 100000])
def test_grid_knn_is_exact(hh, small_pair, cap):
    src, tgt, _ = small_pair
    for k in (10, 20):
        for q in (tgt, src):
            i0, d0 = knn_bruteforce(tgt, q, k)
            i1, d1 = _knn(hh, tgt, q, k, cap)
            assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_grid_knn_ties_and_outside_queries(hh, small_pair):
    lat = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lat = lat[np.random.default_rng(0).permutation(len(lat))]
    for cap in (1, 50, 320, 5000):
        i0, d0 = knn_bruteforce(lat, lat, 20)
        i1, d1 = _knn(hh, lat, lat, 20, cap)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    src, tgt, _ = small_pair
    far = src.copy()
    far[:, 0] += 1000.0
    far[::2, 2] -= 500.0
    i0, d0 = knn_bruteforce(tgt, far, 10)
    i1, d1 = _knn(hh, tgt, far, 10, 20000)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_bounded_nn1(hh, small_pair):
    src, tgt, _ = small_pair
    c = np.ascontiguousarray(tgt[:, :3])
    q = np.ascontiguousarray(src[:, :3])
    for limit2 in (4.0, 0.25, np.inf):
        idx = np.zeros(len(q), np.int32)
        d2 = np.zeros(len(q), np.float32)
        hh.hh_nn1(c.ctypes.data_as(_fp), len(c), 10000, q.ctypes.data_as(_fp), len(q), limit2, idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
        i0, d0 = knn_bruteforce(tgt, src, 1)
        ok = d0[:, 0] < limit2
        assert np.array_equal(idx[ok], i0[ok, 0]) and np.array_equal(d2[ok], d0[ok, 0])
        assert not (d2[~ok] < limit2).any()   # nothing inside the gate is ever invented


@pytest.mark.parametrize("cap,chunk", [(1, 7), (2000, 1), (5000, 10), (5000, 64), (40000, 5)])
def test_chained_ball_knn_is_exact(hh, small_pair, cap, chunk):
    """The kNN kernel's schedule: ring search for the first query of a chunk, triangle-inequality ball for the rest."""
    _, tgt, _ = small_pair
    lat = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lat = lat[np.random.default_rng(0).permutation(len(lat))]
    for cloud in (tgt, lat):
        c = np.ascontiguousarray(cloud[:, :3], np.float32)
        for k in (10, 20):
            idx = np.zeros((len(c), k), np.int32)
            d2 = np.zeros((len(c), k), np.float32)
            hh.hh_knn_chained(c.ctypes.data_as(_fp), len(c), cap, chunk, k, idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
            i0, d0 = knn_bruteforce(cloud, cloud, k)
            assert np.array_equal(i0, idx) and np.array_equal(d0, d2)


@pytest.mark.parametrize("cap,rings", [(5000, 3), (40000, 1), (40000, 0), (300000, 2)])
def test_pyramid_knn_is_exact(hh, small_pair, cap, rings):
    """Fine grid for a few rings, then the coarser pyramid levels (apd_internal.h pyramid_search)."""
    src, tgt, _ = small_pair
    hh.hh_knn_pyramid.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _ip, _fp]
    c = np.ascontiguousarray(tgt[:, :3], np.float32)
    far = src[:, :3].copy()
    far[::3, 0] += 300.0
    for q in (c, np.ascontiguousarray(far, np.float32)):
        for k in (10, 20):
            idx = np.zeros((len(q), k), np.int32)
            d2 = np.zeros((len(q), k), np.float32)
            used = hh.hh_knn_pyramid(c.ctypes.data_as(_fp), len(c), cap, rings, q.ctypes.data_as(_fp), len(q), k, idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
            i0, d0 = knn_bruteforce(tgt, q, k)
            assert np.array_equal(i0, idx) and np.array_equal(d0, d2)
            assert used > 0   # the coarse levels really are exercised


def test_packed_key_fast_path_is_exact(hh, small_pair):
    """32-bit packed candidate list + completeness check + exact fix-up (or exact redo) == brute force."""
    src, tgt, _ = small_pair
    hh.hh_knn_packed.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _ip, _fp]
    lat = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lat = lat[np.random.default_rng(0).permutation(len(lat))]
    dup = np.concatenate([tgt[:300], tgt[:300], tgt[:300]])          # exact duplicates: ties everywhere
    for cloud, queries, expect_fallbacks in ((tgt, tgt, False), (tgt, src, False), (lat, lat, True), (dup, dup, False)):
        c = np.ascontiguousarray(cloud[:, :3], np.float32)
        q = np.ascontiguousarray(queries[:, :3], np.float32)
        for k in (10, 20):
            idx = np.zeros((len(q), k), np.int32)
            d2 = np.zeros((len(q), k), np.float32)
            fb = hh.hh_knn_packed(c.ctypes.data_as(_fp), len(c), 4 * len(c), q.ctypes.data_as(_fp), len(q), k, idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
            i0, d0 = knn_bruteforce(cloud, queries, k)
            assert np.array_equal(i0, idx) and np.array_equal(d0, d2)
            if expect_fallbacks:
                assert fb > 0            # the exact redo path is exercised
            else:
                assert fb <= len(q) // 100   # and is rare on radar-like data


def test_seeded_ball_nn1_is_exact(hh, small_pair):
    src, tgt, _ = small_pair
    c = np.ascontiguousarray(tgt[:, :3])
    q = np.ascontiguousarray(src[:, :3])
    i0, d0 = knn_bruteforce(tgt, src, 1)
    rng = np.random.default_rng(5)
    for limit2 in (4.0, np.inf):
        # seeds: the true neighbour, a random point, or none
        seed = np.where(rng.uniform(size=len(q)) < 0.4, i0[:, 0], rng.integers(0, len(c), len(q))).astype(np.int32)
        seed[rng.uniform(size=len(q)) < 0.2] = -1
        idx = np.zeros(len(q), np.int32)
        d2 = np.zeros(len(q), np.float32)
        hh.hh_nn1_seeded(c.ctypes.data_as(_fp), len(c), 6000, q.ctypes.data_as(_fp), len(q), seed.ctypes.data_as(_ip), limit2,
                         idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
        ok = (seed >= 0) | (d0[:, 0] <= limit2)
        assert np.array_equal(idx[ok], i0[ok, 0]) and np.array_equal(d2[ok], d0[ok, 0])
        assert (idx[~ok] == -1).all()


def test_small_solvers(hh):
    rng = np.random.default_rng(3)
    for _ in range(200):
        A = rng.normal(size=(3, 3))
        S = A @ A.T
        c6 = np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])
        w = np.zeros(3)
        V = np.zeros(9)
        hh.hh_sym_eig3(c6.ctypes.data_as(_dp), w.ctypes.data_as(_dp), V.ctypes.data_as(_dp))
        V = V.reshape(3, 3)
        assert np.allclose(np.sort(w)[::-1], w) and np.allclose(w, np.linalg.eigvalsh(S)[::-1], rtol=1e-10, atol=1e-12)
        assert np.allclose(V @ np.diag(w) @ V.T, S, atol=1e-10)
        B = rng.normal(size=(6, 6))
        H = B @ B.T + 1e-3 * np.eye(6)
        rhs = rng.normal(size=6)
        x = np.zeros(6)
        hh.hh_ldlt6(np.ascontiguousarray(H).ctypes.data_as(_dp), rhs.ctypes.data_as(_dp), x.ctypes.data_as(_dp))
        assert np.allclose(H @ x, rhs, atol=1e-8)
        from scipy.spatial.transform import Rotation
        wv = rng.normal(size=3) * rng.choice([1e-7, 0.01, 1.0])
        R = np.zeros(9)
        hh.hh_so3_exp(wv.ctypes.data_as(_dp), R.ctypes.data_as(_dp))
        assert np.allclose(R.reshape(3, 3), Rotation.from_rotvec(wv).as_matrix(), atol=1e-13)
    # singular system: zero pivots contribute nothing (Eigen LDLT::solve rule) -> x = 0 for H = 0
    x = np.ones(6)
    hh.hh_ldlt6(np.zeros(36).ctypes.data_as(_dp), np.zeros(6).ctypes.data_as(_dp), x.ctypes.data_as(_dp))
    assert np.array_equal(x, np.zeros(6))
