"""GPU parity tests, second file: the branches and inputs the first round left untested.

  * LM control flow (lsq_registration_impl.hpp:127-173): rejected trials, lambda * nu growth, rejected-but-converged
    (x0 not moved), "lm not converged!!" (APD_STATUS_LM_FAILED), non-default initial lambda factor — tests/lm_cases.py
  * non-finite rows (NaN / +-inf) and exact on-axis points (y = z = 0, cos(AoA) -> 0, fast_apdgicp_impl.hpp:168-171)
    through the whole align
  * covariances (not only index sets) on degenerate clouds, with LAPACK as a third opinion
  * full-size configs of BASELINE.json: C3 (5k scan vs 100k submap) and C5 (200k vs 1M, k = 10 / 15 / 20)
"""
import os

import numpy as np
import pytest

from conftest import LAUNCH_PARAMS, TIGHT_PARAMS, ROOT
from test_gpu_parity import _gpu, _oracle, _assert_same_transform, _rot_angle, REL_TOL, _degenerate_clouds

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- LM branches

def _lm_names():
    import lm_cases
    return list(lm_cases.CASES)


@pytest.mark.parametrize("team", [0, 1, 4])
@pytest.mark.parametrize("name", _lm_names())
def test_lm_branches_match_oracle_and_golden(name, team):
    """Full 8-column LM trace, final transform, converged flag, iteration count, status and final Hessian against the
    live oracle AND the committed vectors, for every team shape (cluster, single CTA)."""
    import lm_cases
    from riv_slam_b200 import fast_apdgicp as F
    gold = np.load(os.path.join(ROOT, "tests", "golden", "apd_golden_lm_v1.npz"))
    src, tgt, _ = lm_cases.make_pair()
    r = lm_cases.run_oracle(name, src, tgt)
    g = _gpu(lm_cases.case_params(name))
    g.setOption("team_size", team)
    g.setInputSource(src); g.setInputTarget(tgt)
    g.setSourceCovariances(r["cov_src"]); g.setTargetCovariances(r["cov_tgt"])
    g.align(want_output=False)
    tr = g.getLMTrace()
    for tr0 in (r["trace"], gold[f"{name}_trace"]):
        assert tr.shape == tr0.shape, (tr, tr0)
        assert np.array_equal(tr[:, [0, 1, 7]], tr0[:, [0, 1, 7]])            # outer, inner, accepted: same decisions
        assert np.allclose(tr[:, [2, 3]], tr0[:, [2, 3]], rtol=1e-7)           # y0, yi
        assert np.allclose(tr[:, [5, 6]], tr0[:, [5, 6]], rtol=1e-6)           # lambda (incl. nu growth), |d|
        assert np.allclose(tr[:, 4], tr0[:, 4], rtol=1e-4, atol=1e-4)          # rho
    assert g.hasConverged() == r["converged"] and g.nr_iterations() == r["iterations"]
    assert [int(g.hasConverged()), g.nr_iterations(), int(g.status() == F.APD_STATUS_LM_FAILED)] == list(gold[f"{name}_state"])
    assert (g.status() == F.APD_STATUS_LM_FAILED) == r["lm_failed"]
    _assert_same_transform(g.getFinalTransformation(), r["T"])
    _assert_same_transform(g.getFinalTransformation(), gold[f"{name}_T"])
    assert abs(g.getFitnessScore() - r["fitness"]) <= REL_TOL * r["fitness"]
    Hf0 = r["final_hessian"]
    assert np.abs(g.getFinalHessian() - Hf0).max() <= REL_TOL * np.abs(Hf0).max()
    if name.startswith("lm_failed"):
        # nothing was accepted: final_hessian_ keeps the constructor's identity, x0 keeps the guess (:23, :71-74)
        assert np.array_equal(g.getFinalHessian(), np.eye(6))
        assert np.array_equal(g.getFinalTransformation(), np.eye(4, dtype=np.float32))


def test_pair_without_inliers_between_ordinary_pairs_in_a_batch():
    """TEAM_CTA batch: a pair with nothing inside the gate (H = b = 0, converged at iteration 0) sits between two
    ordinary pairs; the per-pair LM state (lambda, converged flag, staged target) must not leak across pairs."""
    import lm_cases
    from riv_slam_b200 import fast_apdgicp as F
    src, tgt, _ = lm_cases.make_pair()
    p = dict(LAUNCH_PARAMS)
    H = F.Handle(0)
    H.set_params(**p)
    far = src.copy(); far[:, 0] += 500.0
    res = F.batch_align(H, [src, far, src], [tgt, tgt, tgt])
    o = _oracle(p)
    o.set_source(src); o.set_target(tgt)
    rc, T0, conv0, it0 = o.align()
    for i in (0, 2):
        assert bool(res[i]["converged"]) == conv0 and int(res[i]["iterations"]) == it0
        _assert_same_transform(res[i]["T"].reshape(4, 4), T0)
    assert np.array_equal(res[0]["T"], res[2]["T"])
    assert bool(res[1]["converged"]) and int(res[1]["iterations"]) == 0 and int(res[1]["num_inliers"]) == 0


# ---------------------------------------------------------------- non-finite rows and on-axis points

def _poisoned_pair(small_pair):
    src, tgt, T_gt = small_pair
    src = np.array(src, copy=True); tgt = np.array(tgt, copy=True)
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    src[5, 0] = nan
    src[77, :3] = nan
    src[300, 2] = inf
    src[301, 1] = -inf
    tgt[0, 1] = nan
    tgt[640, :3] = nan
    tgt[900, 0] = inf
    tgt[1299, 2] = -inf
    bad_s = np.array([5, 77, 300, 301]); bad_t = np.array([0, 640, 900, 1299])
    return src, tgt, bad_s, bad_t


@pytest.mark.parametrize("unstaged", [0, 1])
def test_non_finite_rows_through_align(small_pair, unstaged):
    """NaN / inf rows: never anybody's neighbour (PCL leaves them out of the kd-tree), never matched as a query (NaN
    distances never enter FLANN's result set) — oracle/kdtree.hpp. Everything else must be untouched by them."""
    src, tgt, bad_s, bad_t = _poisoned_pair(small_pair)
    ok_s = np.setdiff1d(np.arange(len(src)), bad_s); ok_t = np.setdiff1d(np.arange(len(tgt)), bad_t)
    for params in (LAUNCH_PARAMS, TIGHT_PARAMS):
        g = _gpu(params); g.setOption("force_unstaged", unstaged)
        o = _oracle(params)
        g.setInputSource(src); g.setInputTarget(tgt)
        o.set_source(src); o.set_target(tgt)
        assert o.compute_covariances() == 0
        ks, kt = g.getKnn(0), g.getKnn(1)
        assert np.array_equal(ks[ok_s], o.knn(0)[ok_s]) and np.array_equal(kt[ok_t], o.knn(1)[ok_t])
        assert not np.isin(ks[ok_s], bad_s).any() and not np.isin(kt[ok_t], bad_t).any()
        assert (ks[bad_s] == -1).all() and (kt[bad_t] == -1).all() and (o.knn(0)[bad_s] == -1).all()
        Cs, Cs0 = g.getSourceCovariances()[:, :3, :3], o.covariances(0)
        assert np.abs(Cs[ok_s] - Cs0[ok_s]).max() <= REL_TOL
        e, H, b = g.evaluateCost(np.eye(4))
        e0, H0, b0 = o.linearize(np.eye(4))
        corr, sq = g.getCorrespondences()
        corr0, sq0 = o.correspondences()
        assert np.array_equal(corr, corr0) and (corr[bad_s] == -1).all() and not np.isin(corr, bad_t).any()
        assert np.isfinite(H).all() and np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max() and abs(e - e0) <= REL_TOL * abs(e0)
        g.align(want_output=False)
        rc, T0, conv0, it0 = o.align()
        assert g.hasConverged() == conv0 and g.nr_iterations() == it0
        assert np.isfinite(g.getFinalTransformation()).all()
        _assert_same_transform(g.getFinalTransformation(), T0)
        f0 = o.fitness()
        assert np.isfinite(f0) and abs(g.getFitnessScore() - f0) <= REL_TOL * f0
        assert abs(g.getFitnessScore(1.5) - o.fitness(1.5)) <= REL_TOL * o.fitness(1.5)
        assert g.result().num_inliers == int((o.correspondences()[0] >= 0).sum())


def test_non_finite_rows_in_a_batch(small_pair):
    from riv_slam_b200 import fast_apdgicp as F
    src, tgt, bad_s, bad_t = _poisoned_pair(small_pair)
    H = F.Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    res = F.batch_align(H, [src, small_pair[0]], [tgt, small_pair[1]])
    for i, (s, t) in enumerate([(src, tgt), (small_pair[0], small_pair[1])]):
        o = _oracle(LAUNCH_PARAMS)
        o.set_source(s); o.set_target(t)
        rc, T0, conv0, it0 = o.align()
        assert bool(res[i]["converged"]) == conv0 and int(res[i]["iterations"]) == it0
        _assert_same_transform(res[i]["T"].reshape(4, 4), T0)
        assert abs(float(res[i]["fitness"]) - o.fitness()) <= REL_TOL * o.fitness()


def test_on_axis_points_through_align(small_pair):
    """Points exactly on the sensor's x axis (y = z = 0): AoA = atan2f(x, 0) = float(pi/2), cos(AoA) = -4.4e-8, so
    s_y and s_z are ~2e7 * r (fast_apdgicp_impl.hpp:168-171): huge but finite covariances, weights ~ 0. The kernels
    must propagate exactly that (no clamp); the origin (r = 0, every angle atan2f(0, 0) = 0) rides along."""
    src, tgt, T_gt = small_pair
    src = np.array(src, copy=True); tgt = np.array(tgt, copy=True)
    xs = np.array([3.0, 7.5, 12.25, 20.0, 41.0, 63.5, -4.0], dtype=np.float32)
    src[:7, 0] = xs; src[:7, 1] = 0.0; src[:7, 2] = 0.0
    src[7, :3] = 0.0
    tgt[:7, 0] = xs + np.float32(0.05); tgt[:7, 1] = 0.0; tgt[:7, 2] = 0.0
    g = _gpu(LAUNCH_PARAMS); o = _oracle(LAUNCH_PARAMS)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    e, H, b = g.evaluateCost(np.eye(4))       # identity: the float transform leaves the points exactly on the axis
    e0, H0, b0 = o.linearize(np.eye(4))
    corr, sq = g.getCorrespondences()
    corr0, sq0 = o.correspondences()
    assert np.array_equal(corr, corr0) and (corr0[:7] >= 0).all()
    M, M0 = g.getMahalanobis()[:, :3, :3], o.mahalanobis()
    # the on-axis weights really collapse: s_y (azimuthal tangent = y) and s_z (radial = x here) blow up
    assert np.isfinite(M0[:8]).all() and np.abs(M0[:7, 0, 0]).max() < 1e-9 and np.abs(M0[:7, 1, 1]).max() < 1e-9 and (M0[:7, 2, 2] > 0.1).all()
    d, d0 = np.diagonal(M[:7], axis1=1, axis2=2), np.diagonal(M0[:7], axis1=1, axis2=2)
    assert (np.abs(d - d0) <= REL_TOL * np.abs(d0)).all()   # also the collapsed entries themselves, to 1e-5 relative
    m = corr0 >= 0
    scale = np.abs(M0[m]).max(axis=(1, 2))
    assert (np.abs(M[m] - M0[m]).max(axis=(1, 2)) <= REL_TOL * scale).all()
    assert np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max() and np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()
    for params in (LAUNCH_PARAMS, TIGHT_PARAMS):
        g = _gpu(params); o = _oracle(params)
        g.setInputSource(src); g.setInputTarget(tgt)
        o.set_source(src); o.set_target(tgt)
        g.align(want_output=False)
        rc, T0, conv0, it0 = o.align()
        assert g.hasConverged() == conv0 and g.nr_iterations() == it0
        _assert_same_transform(g.getFinalTransformation(), T0)
        tr, tr0 = g.getLMTrace(), o.trace()
        assert tr.shape == tr0.shape and np.array_equal(tr[:, [0, 1, 7]], tr0[:, [0, 1, 7]])


# ---------------------------------------------------------------- covariances on degenerate clouds

def _lapack_third_opinion(cloud, knn, reg):
    """numpy/LAPACK covariances from the SAME index sets; returns (U diag V^T, V diag V^T, eigenvalues descending)."""
    P = np.asarray(cloud, np.float32)[:, :3]
    X = P[knn].astype(np.float64)
    Xc = X - X.mean(axis=1, keepdims=True)
    C = np.einsum("nka,nkb->nab", Xc, Xc) / knn.shape[1]
    U, S, Vt = np.linalg.svd(C)
    if reg == 3:
        vals = np.broadcast_to(np.array([1.0, 1.0, 1e-3]), S.shape)
    elif reg == 1:
        vals = np.maximum(S, 1e-3)
    else:
        vals = np.maximum(S / np.maximum(S.max(axis=1, keepdims=True), 1e-300), 1e-3)
    usv = np.einsum("nab,nb,nbc->nac", U, vals, Vt)
    vsv = np.einsum("nba,nb,nbc->nac", Vt, vals, Vt)
    return usv, vsv, S, vals


@pytest.mark.parametrize("reg", [3, 1, 0, 4], ids=["PLANE", "MIN_EIG", "NONE", "FROBENIUS"])
def test_covariances_on_degenerate_clouds(reg):
    """Rank-deficient neighbourhoods (flat, collinear, duplicated, identical points): the kernel against the oracle
    everywhere, and both against LAPACK's SVD wherever the answer is well defined. U diag V^T of a symmetric PSD
    matrix does not depend on the decomposition exactly when the substituted values are constant on every cluster
    of (numerically) equal singular values; points where that fails are excluded and their fraction is reported
    ('compare with a tolerance scaled by the eigen-gap')."""
    report = []
    for name, cloud in _degenerate_clouds():
        if name == "huge_coordinates":
            continue  # float32 coordinates at 4e5: the covariance itself only carries ~3 digits; index sets are tested elsewhere
        g = _gpu(k_correspondences=20, regularization=reg)
        o = _oracle(k_correspondences=20, regularization=reg)
        g.setInputSource(cloud); o.set_source(cloud); o.set_target(cloud)
        assert o.compute_covariances() == 0
        knn = g.getKnn(0)
        assert np.array_equal(knn, o.knn(0)), name
        C, C0 = g.getSourceCovariances()[:, :3, :3], o.covariances(0)
        assert np.isfinite(C0).all(), name
        scale = np.maximum(np.abs(C0).max(axis=(1, 2)), 1e-300)
        tol0 = np.full(len(C0), REL_TOL)
        if reg == 4:
            # FROBENIUS inverts (C + 1e-3 I) and the normalised inverse again (fast_apdgicp_impl.hpp:330-333): two inversions of a
            # matrix whose condition number reaches 1e7 on a neighbourhood stretched over kilometres (the 'one_outlier' point), by
            # cofactors (Eigen's 3x3 inverse does the same). The cofactors of a nearly rank-1 matrix cancel, so the first inverse
            # carries entrywise errors of cond * eps, which the second inversion amplifies by cond again: oracle (no FMA
            # contraction) and kernel (FMA) agree to cond^2 * eps there, not to 1e-5. Scale the tolerance accordingly.
            P = np.asarray(cloud, np.float32)[:, :3]
            X = P[knn].astype(np.float64)
            Xc = X - X.mean(axis=1, keepdims=True)
            cond = np.linalg.cond(np.einsum("nka,nkb->nab", Xc, Xc) / 20 + 1e-3 * np.eye(3))
            tol0 = np.maximum(REL_TOL, 1e-15 * cond * cond)
            report.append((name, "points with a cond-scaled tolerance", float((tol0 > REL_TOL).mean())))
        assert (np.abs(C - C0).max(axis=(1, 2)) <= tol0 * scale).all(), (name, reg)   # kernel == oracle, all points
        if reg in (0, 4):
            continue
        usv, vsv, S, vals = _lapack_third_opinion(cloud, knn, reg)
        smax = np.maximum(S[:, :1], 1e-300)
        rel_gap = np.abs(S[:, :-1] - S[:, 1:]) / smax            # gaps (s0-s1), (s1-s2) relative to the largest
        same_val = np.abs(vals[:, :-1] - vals[:, 1:]) <= 1e-12   # substituted values equal across the pair
        well_defined = ((rel_gap > 1e-6) | same_val).all(axis=1)
        err_v = np.abs(C - vsv).max(axis=(1, 2)) / np.maximum(np.abs(vsv).max(axis=(1, 2)), 1e-300)
        # tolerance scaled by the smallest relevant gap: eigenvectors move by eps / gap
        gap = np.where(same_val, 1.0, rel_gap).min(axis=1)
        tol = np.maximum(REL_TOL, 1e-12 / np.maximum(gap, 1e-300))
        bad = well_defined & (err_v > tol)
        assert not bad.any(), (name, reg, int(bad.sum()), float(err_v[bad].max()))
        # where LAPACK's U and V disagree (sign of a null-space vector), U diag V^T is not even symmetric: count those
        usv_asym = (np.abs(usv - vsv).max(axis=(1, 2)) > 1e-9).mean()
        report.append((name, float(1.0 - well_defined.mean()), float(usv_asym)))
    print("degenerate covariances: (cloud, fraction excluded as ill-defined, fraction where LAPACK U != V)", report)


# ---------------------------------------------------------------- full-size configs C3 and C5

def test_c3_scan_vs_100k_submap_full_size():
    """BASELINE.json config 3: 5k-point scans against a ~100k-point accumulated submap. kNN sets of the submap
    bit-exact, covariances, H/b at the odometry guess, final transforms, iteration counts, fitness."""
    from riv_slam_b200 import datagen
    from riv_slam_b200 import fast_apdgicp as F
    n_q, per = 6, 5000
    scans, poses = datagen.make_drive(3, 0, 20 + n_q, per, workers=min(16, os.cpu_count() or 1))
    sub = []
    for t in range(20):
        Trel = np.linalg.inv(poses[0]) @ poses[t]
        p = scans[t].copy()
        p[:, :3] = (scans[t][:, :3].astype(np.float64) @ Trel[:3, :3].T + Trel[:3, 3]).astype(np.float32)
        sub.append(p)
    submap = np.concatenate(sub)
    assert submap.shape[0] == 100000
    queries = scans[20:20 + n_q]
    guess = (np.linalg.inv(poses[0]) @ poses[19]).astype(np.float32)
    o = _oracle(LAUNCH_PARAMS)
    o.set_target(submap); o.set_source(queries[0])
    assert o.compute_covariances() == 0
    g = _gpu(LAUNCH_PARAMS)
    g.setInputTarget(submap, cache_key=1); g.setInputSource(queries[0], cache_key=2)
    assert np.array_equal(g.getKnn(1), o.knn(1))
    C, C0 = g.getTargetCovariances()[:, :3, :3], o.covariances(1)
    assert (np.abs(C - C0).max(axis=(1, 2)) <= REL_TOL * np.abs(C0).max(axis=(1, 2))).all()
    e, H, b = g.evaluateCost(guess)
    e0, H0, b0 = o.linearize(guess)
    assert np.array_equal(g.getCorrespondences()[0], o.correspondences()[0])
    assert np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max() and np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()
    for i, q in enumerate(queries):
        g.setInputTarget(submap, cache_key=1); g.setInputSource(q, cache_key=10 + i)
        o.set_source(q)
        g.align(guess, want_output=False)
        rc, T0, conv0, it0 = o.align(guess)
        assert g.hasConverged() == conv0 and g.nr_iterations() == it0
        _assert_same_transform(g.getFinalTransformation(), T0)
        assert abs(g.getFitnessScore() - o.fitness()) <= REL_TOL * o.fitness()
    # the batched form of the same thing: every query against the one target
    Hn = F.Handle(0); Hn.set_params(**LAUNCH_PARAMS)
    S, T = F.CloudSet(Hn, queries), F.CloudSet(Hn, [submap])
    res = F.align_pairs(Hn, S, T, tgt_idx=np.zeros(n_q, np.int32), guesses=np.stack([guess] * n_q))
    o.set_source(queries[-1])
    rc, T0, conv0, it0 = o.align(guess)
    _assert_same_transform(res[-1]["T"].reshape(4, 4), T0)
    assert int(res[-1]["iterations"]) == it0


def test_c5_200k_vs_1m_full_size():
    """BASELINE.json config 5: 200k-point source vs 1M-point target (accumulated keyframe maps), k = 10 / 15 / 20:
    kNN index sets of all 1.2M points bit-exact, covariances, H/b/error at identity, the final transform."""
    from riv_slam_b200 import datagen
    w = min(16, os.cpu_count() or 1)
    tgt, pose_t = datagen.make_map(5, 0, 330, 5000, frame=150, workers=w, total_scans=340, speed=1.0)
    src, pose_s = datagen.make_map(5, 0, 80, 5000, frame=151, first=112, workers=w, total_scans=340, speed=1.0, resample=1)
    rng = np.random.default_rng(5)
    tgt = tgt[np.sort(rng.choice(tgt.shape[0], min(1000000, tgt.shape[0]), replace=False))]
    src = src[np.sort(rng.choice(src.shape[0], min(200000, src.shape[0]), replace=False))]
    assert tgt.shape[0] == 1000000 and src.shape[0] == 200000
    for k in (10, 15, 20):
        p = dict(LAUNCH_PARAMS, k_correspondences=k)
        g = _gpu(p); o = _oracle(p)
        g.setInputTarget(tgt, cache_key=1); g.setInputSource(src, cache_key=2)
        o.set_target(tgt); o.set_source(src)
        assert o.compute_covariances() == 0
        for which in (0, 1):
            assert np.array_equal(g.getKnn(which), o.knn(which)), (k, which)
        for which, C in ((0, g.getSourceCovariances()), (1, g.getTargetCovariances())):
            C0 = o.covariances(which)
            assert (np.abs(C[:, :3, :3] - C0).max(axis=(1, 2)) <= REL_TOL * np.abs(C0).max(axis=(1, 2))).all(), (k, which)
        e, H, b = g.evaluateCost(np.eye(4))
        e0, H0, b0 = o.linearize(np.eye(4))
        assert np.array_equal(g.getCorrespondences()[0], o.correspondences()[0])
        assert abs(e - e0) <= REL_TOL * abs(e0)
        assert np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max() and np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()
        if k == 20:
            g.align(want_output=False)
            rc, T0, conv0, it0 = o.align()
            assert g.hasConverged() == conv0 and g.nr_iterations() == it0
            _assert_same_transform(g.getFinalTransformation(), T0)
            assert abs(g.getFitnessScore() - o.fitness()) <= REL_TOL * o.fitness()


# ---------------------------------------------------------------- small surface checks

def test_python_select_registration_method_matches_factory_defaults(small_pair):
    """registrations.cpp:38-50: the FAST_APDGICP branch with its rosparam defaults, then one registration."""
    from riv_slam_b200 import fast_apdgicp as F
    reg = F.select_registration_method({})
    p = reg.handle().get_params()
    assert (p.k_correspondences, p.max_iterations) == (20, 64)
    assert (p.max_corr_dist, p.transformation_epsilon, p.dist_var, p.azimuth_var, p.elevation_var) == (2.5, 0.01, 0.86, 0.5, 1.0)
    reg = F.select_registration_method({"reg_max_correspondence_distance": 2.0, "reg_transformation_epsilon": 0.1, "azimuth_var": 1.0, "reg_num_threads": 4})
    src, tgt, _ = small_pair
    reg.setInputTarget(tgt); reg.setInputSource(src)
    reg.align(want_output=False)
    o = _oracle(LAUNCH_PARAMS)
    o.set_source(src); o.set_target(tgt)
    rc, T0, conv0, it0 = o.align()
    assert reg.hasConverged() == conv0 and reg.nr_iterations() == it0
    _assert_same_transform(reg.getFinalTransformation(), T0)


def test_protected_hooks_and_inlier_count(small_pair):
    """linearize / update_correspondences / compute_error at a double pose (fast_apdgicp.hpp:77-83) and the status message's
    inlier pass (scan_matching_odometry_nodelet.cpp:698-712) through the C ABI."""
    src, tgt, T_gt = small_pair
    g = _gpu(LAUNCH_PARAMS); o = _oracle(LAUNCH_PARAMS)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    x = np.array(T_gt, dtype=np.float64); x[0, 3] += 1e-9
    y = x.copy(); y[1, 3] += 0.03
    e, H, b = g.linearize(x)
    e0, H0, b0 = o.linearize_d(x)
    assert abs(e - e0) <= 1e-10 * abs(e0) and np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max() and np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()
    assert abs(g.compute_error(x) - e0) <= 1e-10 * abs(e0)
    assert abs(g.compute_error(y) - o.compute_error_d(y)) <= 1e-9 * abs(e0)
    g.update_correspondences(y); o.linearize_d(y)
    assert np.array_equal(g.getCorrespondences()[0], o.correspondences()[0])
    assert abs(g.compute_error(y) - o.compute_error_d(y)) <= 1e-9 * abs(e0)
    g.align(want_output=False)
    rc, T0, conv0, it0 = o.align()
    T = g.getFinalTransformation()
    for d in (0.5, 0.1, 2.0):
        n_fast = g.inlierCount(d)           # from the align kernel's per-point distances
        n_slow = g.inlierCount(d, T=T)      # a fresh search at the same pose
        assert n_fast == n_slow == o.inlier_count(T, d), d
    g2 = _gpu(LAUNCH_PARAMS)
    g2.setInputSource(src); g2.setInputTarget(tgt)
    with pytest.raises(Exception):
        g2.compute_error(x)                 # no linearize yet: refused, not garbage


@pytest.mark.gpu
def test_leaf_knn_warp_parts_give_identical_results():
    """knn_cov_leaf_kernel cuts the 32 queries of a leaf into 1 / 2 / 4 / 8 warp parts for a single scan (fewer leaves than warps).
    The split only changes which warp serves a query, and with it whether a leaf is scanned broadcast or transposed: neighbour
    lists and covariances must come out bit-identical, and equal to the oracle's."""
    from oracle.oracle import Oracle
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import FastAPDGICP
    scans, _ = datagen.make_sequence(2, 3, n_scans=1, n_points=3000)
    cloud = np.ascontiguousarray(scans[0][:, :3])
    o = Oracle(**LAUNCH_PARAMS)
    o.set_source(cloud); o.set_target(cloud)
    o.compute_covariances()
    ref_knn = o.knn(0)
    ref_cov = o.covariances(0)
    got = {}
    for parts in (1, 2, 4, 8):
        reg = FastAPDGICP(0)
        reg.handle().set_params(**LAUNCH_PARAMS)
        reg.setOption("knn_leaf_parts", parts)
        reg.setInputSource(cloud, cache_key=10 + parts)
        reg.setInputTarget(cloud, cache_key=10 + parts)
        reg.computeCovariances()
        got[parts] = (reg.getKnn(0), reg.getSourceCovariances())
        assert np.array_equal(got[parts][0], ref_knn), parts
        assert np.abs(got[parts][1][:, :3, :3] - ref_cov[:, :3, :3]).max() <= 1e-9
    for parts in (2, 4, 8):
        assert np.array_equal(got[parts][0], got[1][0])
        assert np.array_equal(got[parts][1], got[1][1])   # bit-identical covariances


# ---------------------------------------------------------------- staging by the copy engine

def test_bulk_staging_is_bit_identical_to_the_register_path(small_pair):
    """Leaf-mode clouds reach shared memory by ONE cp.async.bulk of the image the build wrote (option "bulk_stage", default on)
    instead of a load / negate / interleave / store loop. Same bytes in shared memory, so everything downstream is bit-identical:
    kNN, covariances, H / b, whole registrations for every team shape, a batch in which every CTA stages many targets in turn
    (the mbarrier's phase flips per staging), and the stand-alone fitness kernel."""
    from riv_slam_b200 import datagen
    from riv_slam_b200 import fast_apdgicp as F
    src, tgt, _ = small_pair
    outs = []
    for bulk in (1, 0):
        r = {}
        for team in (0, 1, 4):
            g = _gpu(LAUNCH_PARAMS)
            g.setOption("bulk_stage", bulk)
            g.setOption("team_size", team)
            g.setInputSource(src); g.setInputTarget(tgt)
            r[f"knn{team}"] = g.getKnn(0).copy()
            r[f"cov{team}"] = g.getTargetCovariances().copy()
            e, H, b = g.evaluateCost(np.eye(4))
            r[f"lin{team}"] = np.r_[e, H.ravel(), b]
            g.align(want_output=False)
            r[f"T{team}"] = g.getFinalTransformation().copy()
            r[f"fit{team}"] = np.array([g.getFitnessScore(), g.getFitnessScore(1.5), g.nr_iterations()])
            r[f"trace{team}"] = g.getLMTrace().copy()
        H = F.Handle(0)
        H.set_params(**LAUNCH_PARAMS)
        H.set_option("bulk_stage", bulk)
        pairs = [datagen.make_pair(2, i % 7, n_src=900 + 37 * (i % 5)) for i in range(14)]
        srcs = [pairs[i % 14][0] for i in range(330)]     # 330 pairs on 148 SMs: every CTA stages two or three different targets
        tgts = [pairs[(i * 5) % 14][1] for i in range(330)]
        res = F.batch_align(H, srcs, tgts)
        r["batch_T"] = np.array(res["T"]); r["batch_fit"] = np.array(res["fitness"]); r["batch_it"] = np.array(res["iterations"])
        cs_s, cs_t = F.CloudSet(H, srcs[:40]), F.CloudSet(H, tgts[:40])
        r["fitness_pairs"] = np.array(F.fitness_pairs(H, cs_s, cs_t, poses=np.tile(np.eye(4, dtype=np.float32), (40, 1, 1))))
        outs.append(r)
    a, b = outs
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k
    assert (a["batch_it"] > 0).all()


# ---------------------------------------------------------------- the host pipeline's upload-ahead path

def test_upload_ahead_pipeline_is_bit_identical():
    """apd_odometry_align / apd_batch_align on PAGE-LOCKED input send the whole array ahead on a copy stream, fetch their table blocks
    and guesses with a kernel and address odometry pairs by index bases; pageable input (and option upload_ahead = 0) keeps the
    per-chunk path. Same records, bit for bit, with and without initial guesses, and equal to one un-pipelined launch."""
    import ctypes as C
    import torch
    from riv_slam_b200 import datagen
    from riv_slam_b200 import fast_apdgicp as F
    n_scans = 701                                            # 700 pairs: chunks of 147, 295, 258
    base, _ = datagen.make_drive(2, 5, 24, 700, workers=4)
    scans = [base[i % 24] for i in range(n_scans)]
    pts, off = F._ragged([np.ascontiguousarray(s[:, :4]) for s in scans])
    pinned = torch.from_numpy(pts.copy()).pin_memory()
    rng = np.random.default_rng(3)
    guesses = np.tile(np.eye(4, dtype=np.float32), (n_scans - 1, 1, 1))
    guesses[:, :3, 3] += rng.normal(0, 0.05, (n_scans - 1, 3)).astype(np.float32)
    H = F.Handle(0)
    H.set_params(**LAUNCH_PARAMS)

    def run(points, ahead, g):
        H.set_option("upload_ahead", ahead)
        return F.odometry_align(H, (points, off), guesses=g).copy()

    for g in (None, guesses):
        a = run(pinned, 1, g)                                # page-locked: upload-ahead
        b = run(pinned, 0, g)                                # page-locked, per-chunk uploads
        c = run(pts, 1, g)                                   # pageable: per-chunk uploads whatever the option says
        assert a.tobytes() == b.tobytes() == c.tobytes()
        assert (a["status"] == 0).all() and (a["iterations"] >= 0).all()
        cs = F.CloudSet(H, [np.ascontiguousarray(s[:, :4]) for s in scans])
        idx = np.arange(n_scans - 1, dtype=np.int32)
        one = F.align_pairs(H, cs, cs, src_idx=idx + 1, tgt_idx=idx, guesses=g)
        assert np.array_equal(one["T"], a["T"]) and np.array_equal(one["iterations"], a["iterations"]) and np.array_equal(one["fitness"], a["fitness"])
    # the batch entry point: separate source and target arrays, both page-locked
    src_pts, src_off = F._ragged([np.ascontiguousarray(s[:, :4]) for s in scans[1:]])
    tgt_pts, tgt_off = F._ragged([np.ascontiguousarray(s[:, :4]) for s in scans[:-1]])
    ps, pt = torch.from_numpy(src_pts.copy()).pin_memory(), torch.from_numpy(tgt_pts.copy()).pin_memory()
    res = np.zeros(n_scans - 1, dtype=F.RESULT_DTYPE)
    ip = C.POINTER(C.c_int32)
    for ahead in (1, 0):
        H.set_option("upload_ahead", ahead)
        H.check(H.L.apd_batch_align(H.h, C.c_void_p(ps.data_ptr()), src_off.ctypes.data_as(ip), C.c_void_p(pt.data_ptr()), tgt_off.ctypes.data_as(ip), 16,
                                    guesses.ctypes.data_as(C.POINTER(C.c_float)), n_scans - 1, C.c_void_p(res.ctypes.data)))
        assert res.tobytes() == a.tobytes()


# ---------------------------------------------------------------- randomised sweep

def test_randomised_sweep_against_the_oracle():
    """The 80 seeded cases of tests/test_reference_apdgicp.py::test_randomised_sweep (there: oracle == the reference's compiled sources)
    through the CUDA path: cloud sizes 25..700, every regularisation, k = 5..20, gates from 0.3 m to unbounded, LM and Gauss-Newton,
    iteration caps down to 1, APD variances on and off, non-identity guesses, clouds metres apart."""
    from riv_slam_b200 import datagen
    from riv_slam_b200 import fast_apdgicp as F
    rng = np.random.default_rng(2026)
    done = 0
    for _ in range(80):
        n_s, n_t = int(rng.integers(25, 700)), int(rng.integers(25, 700))
        src, tgt, _ = datagen.make_pair(int(rng.choice([1, 2, 4])), int(rng.integers(0, 50)), n_src=n_s, n_tgt=n_t)
        if rng.random() < 0.2:
            src = src.copy(); src[:, 0] += rng.uniform(1.5, 30)
        p = dict(k_correspondences=int(rng.choice([5, 10, 15, 20])), regularization=int(rng.integers(0, 5)),
                 max_corr_dist=float(rng.choice([0.3, 1.0, 2.0, 5.0, 3.4e38])), max_iterations=int(rng.choice([1, 3, 16, 64])),
                 optimizer=int(rng.random() < 0.8), lm_max_iterations=int(rng.choice([1, 3, 10])),
                 transformation_epsilon=float(rng.choice([0.1, 5e-4, 1e-3])), rotation_epsilon=float(rng.choice([2e-3, 1e-3])),
                 lm_init_lambda_factor=float(rng.choice([1e-9, 1e-6, 1e-2])), dist_var=float(rng.choice([0.0, 0.86, 2.0])),
                 azimuth_var=float(rng.choice([0.0, 0.5, 1.0])), elevation_var=float(rng.choice([0.0, 1.0, 3.0])))
        if min(n_s, n_t) <= p["k_correspondences"]:
            continue
        G = np.eye(4, dtype=np.float32); G[:3, 3] = rng.normal(0, 0.2, 3)
        o = _oracle(p)
        o.set_source(src); o.set_target(tgt)
        rc0, T0, conv0, it0 = o.align(G)
        g = _gpu(p)
        g.setInputSource(src); g.setInputTarget(tgt)
        g.align(G, want_output=False)
        assert rc0 == 0 and (g.hasConverged(), g.nr_iterations()) == (conv0, it0), p
        assert (g.status() == F.APD_STATUS_LM_FAILED) == o.lm_failed(), p
        _assert_same_transform(g.getFinalTransformation(), T0)
        assert np.array_equal(g.getKnn(0), o.knn(0)) and np.array_equal(g.getKnn(1), o.knn(1)), p
        tr0, tr1 = o.trace(), g.getLMTrace()
        assert tr0.shape == tr1.shape, p
        if tr0.size:
            noise = np.abs(tr0[:, 2] - tr0[:, 3]) <= 1e-9 * np.abs(tr0[:, 2])
            assert np.array_equal(tr0[~noise, 7], tr1[~noise, 7]), p
            assert np.allclose(tr0[:, [2, 3]], tr1[:, [2, 3]], rtol=1e-7, atol=1e-300), p
        f0 = o.fitness()
        assert abs(g.getFitnessScore() - f0) <= REL_TOL * f0, p
        done += 1
    assert done >= 60
