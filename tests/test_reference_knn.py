"""The nearest-neighbour search pinned against REFERENCE code: the exact kd-tree the reference vendors
(radar_graph_slam/include/scan_context/nanoflann.hpp, nanoflann 1.3.2), compiled from /root/reference into
oracle/_ref (`make -C oracle ref`) and frozen into tests/golden/knn_nanoflann_v1.npz
(tests/golden/make_knn_ref_golden.py). PCL / FLANN, which FastAPDGICP calls for the same job, do not exist in
this image; nanoflann uses FLANN's L2_Simple float arithmetic and is exact, so the lists must be identical
wherever the k-th distance is not tied (nanoflann admits a candidate only if it is strictly closer than its
current k-th, so a tie exactly at rank k is resolved by visiting order there, by index here).

CPU: the live library (when oracle/_ref is built, i.e. in the build container) and the committed fixture against
the oracle's kd-tree and brute-force searches. GPU: the CUDA kNN and 1-NN against the fixture through the C ABI.
"""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden", "knn_nanoflann_v1.npz")


def _same_lists(idx, d2, idx_ref, d2_ref):
    """identical distances everywhere; identical indices wherever a distance is unique in its list"""
    assert np.array_equal(d2, d2_ref)
    uniq = np.ones_like(d2_ref, dtype=bool)
    uniq[:, 1:] &= d2_ref[:, 1:] != d2_ref[:, :-1]
    uniq[:, :-1] &= d2_ref[:, :-1] != d2_ref[:, 1:]
    uniq[:, -1] = False  # the k-th entry may be tied with the first one left out
    assert np.array_equal(idx[uniq], idx_ref[uniq])
    return float(uniq.mean())


def _clouds():
    from riv_slam_b200 import datagen
    rng = np.random.default_rng(5)
    radar = datagen.make_pair(1, 11, n_src=2500)[0][:, :3]
    uniform = rng.uniform(-30, 30, (2000, 3)).astype(np.float32)
    lattice = np.stack(np.meshgrid(np.arange(12), np.arange(11), np.arange(9), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.5
    return {"radar": radar, "uniform": uniform, "lattice": lattice}


@pytest.mark.parametrize("name", ["radar", "uniform", "lattice"])
@pytest.mark.parametrize("k", [1, 20])
def test_oracle_search_matches_reference_kdtree(name, k):
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref is built only where /root/reference is mounted (make -C oracle ref)")
    cloud = _clouds()[name]
    rng = np.random.default_rng(9)
    queries = np.concatenate([cloud[::3], cloud[::7] + rng.normal(0, 0.3, cloud[::7].shape).astype(np.float32)])
    ri, rd = O.ref_nanoflann_knn(cloud, queries, k)
    for fn in (O.knn_kdtree, O.knn_bruteforce):
        oi, od = fn(cloud, queries, k)
        frac = _same_lists(oi, od, ri, rd)
        if name != "lattice" and k > 1:
            assert frac > (k - 1) / k - 0.01  # generic data: practically no ties, so every index but the k-th is compared
    # the tie rule itself: equal distances are ordered by index in both
    if name == "lattice" and k > 1:
        oi, od = O.knn_kdtree(cloud, queries, k)
        tied = od[:, 1:] == od[:, :-1]
        assert tied.any() and (oi[:, 1:][tied] > oi[:, :-1][tied]).all()


def test_fixture_matches_oracle():
    from oracle import oracle as O
    g = np.load(GOLDEN)
    src, tgt = g["src"], g["tgt"]
    for k in (10, 20):
        oi, od = O.knn_kdtree(src, src, k)
        assert _same_lists(oi, od, g[f"knn{k}_src_idx"].astype(np.int32), g[f"knn{k}_src_d2"]) > (k - 1) / k - 0.01
    oi, od = O.knn_kdtree(tgt, tgt, 20)
    assert _same_lists(oi, od, g["knn20_tgt_idx"].astype(np.int32), g["knn20_tgt_d2"]) > 0.94
    oi, od = O.knn_bruteforce(tgt, g["queries_1nn"], 1)
    assert np.array_equal(od[:, 0], g["nn1_d2"]) and np.array_equal(oi[:, 0], g["nn1_idx"].astype(np.int32))
    # the fixture was generated from the seeded generator: the inputs are reproducible
    from riv_slam_b200 import datagen
    c, i, ns, nt = (int(v) for v in g["seed"])
    s2, t2, _ = datagen.make_pair(c, i, n_src=ns, n_tgt=nt)
    assert np.array_equal(s2, src) and np.array_equal(t2, tgt)


@pytest.mark.gpu
def test_gpu_search_matches_reference_fixture():
    from riv_slam_b200.fast_apdgicp import FastAPDGICP
    g = np.load(GOLDEN)
    src, tgt = g["src"], g["tgt"]
    for k in (10, 20):
        for unstaged in (0, 1):
            reg = FastAPDGICP(0)
            reg.handle().set_params(k_correspondences=k)
            reg.setOption("force_unstaged", unstaged)
            reg.setInputSource(src); reg.setInputTarget(tgt)
            ref = g[f"knn{k}_src_idx"].astype(np.int32)
            d2 = g[f"knn{k}_src_d2"]
            got = reg.getKnn(0)
            uniq = np.ones_like(d2, dtype=bool)
            uniq[:, 1:] &= d2[:, 1:] != d2[:, :-1]
            uniq[:, :-1] &= d2[:, :-1] != d2[:, 1:]
            uniq[:, -1] = False
            assert np.array_equal(got[uniq], ref[uniq]) and uniq.mean() > (k - 1) / k - 0.01
            if k == 20:
                gt_idx, gt_d2 = g["knn20_tgt_idx"].astype(np.int32), g["knn20_tgt_d2"]
                u2 = np.ones_like(gt_d2, dtype=bool)
                u2[:, 1:] &= gt_d2[:, 1:] != gt_d2[:, :-1]
                u2[:, :-1] &= gt_d2[:, :-1] != gt_d2[:, 1:]
                u2[:, -1] = False
                assert np.array_equal(reg.getKnn(1)[u2], gt_idx[u2])
    # 1-NN correspondences (update_correspondences without a gate: constructor default FLT_MAX)
    reg = FastAPDGICP(0)
    q = np.ascontiguousarray(np.concatenate([g["queries_1nn"], np.ones((len(g["queries_1nn"]), 1), np.float32)], axis=1))
    reg.setInputSource(q); reg.setInputTarget(tgt)
    reg.evaluateCost(np.eye(4))
    corr, sq = reg.getCorrespondences()
    assert np.array_equal(sq, g["nn1_d2"]) and np.array_equal(corr, g["nn1_idx"].astype(np.int32))


def test_product_search_header_matches_reference_kdtree():
    """The product's own search code (riv-slam_b200/csrc/apd_grid.cuh, compiled as plain C++ by tests/host_harness.cpp:
    the same source the GPU runs through nvcc) against the reference's vendored kd-tree, and against the frozen fixture
    where oracle/_ref is not built."""
    import test_host_harness as hh_mod
    L = hh_mod.load_harness()
    from oracle import oracle as O
    g = np.load(GOLDEN)
    src, tgt = g["src"], g["tgt"]
    for cap in (6000, 400):   # cells_per_point 4 and a much coarser grid
        idx, d2 = hh_mod._knn(L, src, src, 20, cap)
        assert _same_lists(idx, d2, g["knn20_src_idx"].astype(np.int32), g["knn20_src_d2"]) > 0.94
        idx, d2 = hh_mod._knn(L, tgt, g["queries_1nn"], 1, cap)
        assert np.array_equal(d2[:, 0], g["nn1_d2"]) and np.array_equal(idx[:, 0], g["nn1_idx"].astype(np.int32))
    if O.ref_available():
        for name, cloud in _clouds().items():
            ri, rd = O.ref_nanoflann_knn(cloud, cloud, 20)
            idx, d2 = hh_mod._knn(L, cloud, cloud, 20, 4 * len(cloud))
            _same_lists(idx, d2, ri, rd)
