"""CPU tests of the oracle itself: C++ restatement vs its independent numpy/LAPACK twin."""
import numpy as np
import pytest

from conftest import LAUNCH_PARAMS, TIGHT_PARAMS
from oracle import pyref
from oracle.oracle import Oracle, knn_bruteforce, knn_kdtree


def _pyref(params, reg=pyref.REG_PLANE):
    p = dict(params)
    k = p.pop("k_correspondences")
    return pyref.PyRef(k=k, reg=reg, **p)


def test_kdtree_matches_bruteforce_and_numpy(small_pair):
    src, tgt, _ = small_pair
    i_bf, d_bf = knn_bruteforce(tgt, src, 20)
    i_kd, d_kd = knn_kdtree(tgt, src, 20)
    i_np, d_np = pyref.knn(src[:, :3], tgt[:, :3], 20)
    assert np.array_equal(i_bf, i_kd) and np.array_equal(d_bf, d_kd)
    assert np.array_equal(i_bf, i_np) and np.array_equal(d_bf, d_np)


def test_knn_ties_break_by_index():
    # lattice cloud: many exactly equal distances
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(0)
    g = g[rng.permutation(g.shape[0])]
    i_bf, d_bf = knn_bruteforce(g, g, 10)
    i_kd, d_kd = knn_kdtree(g, g, 10)
    i_np, _ = pyref.knn(g, g, 10)
    assert np.array_equal(i_bf, i_kd) and np.array_equal(i_bf, i_np)
    # within equal distances indices ascend
    for r in range(g.shape[0]):
        for j in range(9):
            if d_bf[r, j] == d_bf[r, j + 1]:
                assert i_bf[r, j] < i_bf[r, j + 1]
    assert np.array_equal(i_bf[:, 0], np.arange(g.shape[0]))  # self is the nearest (d2 = 0)


@pytest.mark.parametrize("reg", [pyref.REG_PLANE, pyref.REG_NONE, pyref.REG_MIN_EIG, pyref.REG_NORMALIZED_MIN_EIG, pyref.REG_FROBENIUS])
def test_covariances_match_twin(small_pair, reg):
    src, _, _ = small_pair
    o = Oracle(regularization=reg, **LAUNCH_PARAMS)
    o.set_source(src)
    o.set_target(src)
    assert o.compute_covariances() == 0
    C_ref, knn_ref = pyref.covariances(src, 20, reg)
    assert np.array_equal(o.knn(0), knn_ref)
    C = o.covariances(0)
    if reg == pyref.REG_PLANE:
        # tolerance scales with the eigen-gap of the raw covariance (SURVEY.md §7 hard parts)
        C_raw, _ = pyref.covariances(src, 20, pyref.REG_NONE)
        w = np.linalg.eigvalsh(C_raw)
        gap = (w[:, 1] - w[:, 0]) / w[:, 2]
        ok = gap > 1e-6
        assert ok.mean() > 0.99
        err = np.abs(C - C_ref).max(axis=(1, 2))
        assert (err[ok] * gap[ok]).max() < 1e-12
        ev = np.linalg.eigvalsh(C)
        assert np.allclose(ev, [1e-3, 1, 1], atol=1e-12)
    else:
        scale = np.abs(C_ref).max(axis=(1, 2))
        assert (np.abs(C - C_ref).max(axis=(1, 2)) / scale).max() < 1e-9


def test_linearize_matches_twin(small_pair):
    src, tgt, T_gt = small_pair
    o = Oracle(**LAUNCH_PARAMS)
    o.set_source(src)
    o.set_target(tgt)
    r = _pyref(LAUNCH_PARAMS)
    r.set_source(src)
    r.set_target(tgt)
    for pose in (np.eye(4), T_gt):
        e, H, b = o.linearize(pose)
        e2, H2, b2 = r.linearize(np.asarray(pose, dtype=np.float32))
        corr, sq = o.correspondences()
        assert np.array_equal(corr, r.corr)
        assert np.array_equal(sq, r.sq)
        assert (corr >= 0).sum() > 100 and (corr < 0).sum() > 0  # the 2 m gate is active
        assert abs(e - e2) <= 1e-9 * abs(e2)
        assert np.abs(H - H2).max() <= 1e-9 * np.abs(H2).max()
        assert np.abs(b - b2).max() <= 1e-9 * np.abs(b2).max()
        M = o.mahalanobis()
        v = corr >= 0
        assert (np.abs(M[v] - r.M[v]).max(axis=(1, 2)) / np.abs(r.M[v]).max(axis=(1, 2))).max() < 1e-9
        assert np.allclose(H, H.T, rtol=0, atol=1e-9 * np.abs(H).max())


@pytest.mark.parametrize("params", [LAUNCH_PARAMS, TIGHT_PARAMS], ids=["launch", "tight"])
def test_align_matches_twin(small_pair, params):
    src, tgt, T_gt = small_pair
    o = Oracle(**params)
    o.set_source(src)
    o.set_target(tgt)
    rc, T, conv, it = o.align()
    r = _pyref(params)
    r.set_source(src)
    r.set_target(tgt)
    T2, conv2, it2 = r.align()
    assert rc == 0 and conv == conv2 and it == it2
    tr, tr2 = o.trace(), r.trace
    assert tr.shape == tr2.shape
    assert np.array_equal(tr[:, [0, 1, 7]], tr2[:, [0, 1, 7]])  # same accept/reject sequence
    assert np.allclose(tr[:, 2:4], tr2[:, 2:4], rtol=1e-7)
    assert np.abs(T - T2)[:3, :3].max() < 1e-6 and np.abs(T - T2)[:3, 3].max() < 1e-5
    assert abs(o.fitness() - r.fitness()) <= 1e-6 * r.fitness()
    # sanity envelope of the reference's own test (fast_apdgicp/src/test/gicp_test.cpp:148-149) is
    # 0.05 m / 1 deg on lidar data; radar-like clouds are noisier, so only a loose bound is asserted
    assert np.abs(T - T_gt)[:3, 3].max() < 0.3


def test_align_edge_cases(small_pair):
    src, tgt, _ = small_pair
    o = Oracle(**LAUNCH_PARAMS)
    rc, *_ = o.align()
    assert rc == -1  # no clouds
    o.set_source(src[:10])
    o.set_target(tgt)
    assert o.align()[0] == -2  # fewer points than k
    # far-away source: no correspondence inside the gate -> H = b = 0, delta = I, converged at once
    far = src.copy()
    far[:, 0] += 1000.0
    o.set_source(far)
    rc, T, conv, it = o.align()
    assert rc == 0 and conv and it == 0
    assert np.array_equal(T, np.eye(4, dtype=np.float32))
    assert (o.correspondences()[0] < 0).all()


def test_swap_and_backward(small_pair):
    src, tgt, _ = small_pair
    o = Oracle(**TIGHT_PARAMS)
    o.set_source(src)
    o.set_target(tgt)
    _, T_fwd, conv, _ = o.align()
    o.swap()
    _, T_bwd, conv2, _ = o.align()
    assert conv and conv2
    # forward and backward estimates are near-inverse (different correspondences, so only loosely)
    E = T_fwd.astype(np.float64) @ T_bwd.astype(np.float64)
    assert np.abs(E - np.eye(4)).max() < 0.05
    o2 = Oracle(**TIGHT_PARAMS)
    o2.set_source(tgt)
    o2.set_target(src)
    _, T_b2, _, _ = o2.align()
    assert np.array_equal(T_bwd, T_b2)  # swap == fresh object with the roles exchanged
