"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Gates (BASELINE.json north_star): kNN index sets bit-exact; covariances and H/b within 1e-5
relative; final transforms within 1e-5 rad / 1e-4 m; same converged flag and iteration count.
"""
import numpy as np
import pytest

from conftest import LAUNCH_PARAMS, TIGHT_PARAMS

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5    # rad
TRANS_TOL = 1e-4  # m
REL_TOL = 1e-5    # covariances, H, b


def _gpu(params=None, **extra):
    from riv_slam_b200.fast_apdgicp import FastAPDGICP
    reg = FastAPDGICP(0)
    p = dict(params or {})
    p.update(extra)
    if p:
        reg.handle().set_params(**p)
    return reg


def _oracle(params=None, **extra):
    from oracle.oracle import Oracle
    p = dict(params or {})
    p.update(extra)
    return Oracle(**p)


def _rot_angle(Ra, Rb):
    # small-angle safe: sin(angle) from the skew part (arccos of the trace loses half the digits near 0)
    R = Ra.astype(np.float64) @ Rb.astype(np.float64).T
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(v))))


def _assert_same_transform(T, T_ref):
    assert _rot_angle(T[:3, :3], T_ref[:3, :3]) <= ROT_TOL
    assert np.abs(T[:3, 3].astype(np.float64) - T_ref[:3, 3]).max() <= TRANS_TOL


@pytest.fixture(scope="module")
def pair5k():
    from riv_slam_b200 import datagen
    return datagen.make_pair(4, 0, n_src=5000)


# ---------------------------------------------------------------- kNN + covariances

@pytest.mark.parametrize("k", [10, 15, 20])
def test_knn_bit_exact(small_pair, k):
    src, tgt, _ = small_pair
    g = _gpu(LAUNCH_PARAMS, k_correspondences=k)
    o = _oracle(LAUNCH_PARAMS, k_correspondences=k)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    assert o.compute_covariances() == 0
    for which in (0, 1):
        assert np.array_equal(g.getKnn(which), o.knn(which))


def test_knn_bit_exact_5k_and_unstaged(pair5k):
    src, tgt, _ = pair5k
    o = _oracle(LAUNCH_PARAMS)
    o.set_source(src); o.set_target(tgt)
    assert o.compute_covariances() == 0
    for unstaged in (0, 1):
        g = _gpu(LAUNCH_PARAMS)
        g.setOption("force_unstaged", unstaged)
        g.setInputSource(src); g.setInputTarget(tgt)
        assert np.array_equal(g.getKnn(0), o.knn(0))
        assert np.array_equal(g.getKnn(1), o.knn(1))


def test_knn_ties_break_by_index():
    from oracle.oracle import knn_bruteforce
    lat = np.stack(np.meshgrid(np.arange(7), np.arange(6), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lat = lat[np.random.default_rng(1).permutation(lat.shape[0])]
    for cpp in (0.5, 8.0, 64.0):
        g = _gpu(k_correspondences=20)
        g.setOption("cells_per_point", cpp)
        g.setInputSource(lat)
        ref, _ = knn_bruteforce(lat, lat, 20)
        assert np.array_equal(g.getKnn(0), ref)


@pytest.mark.parametrize("reg", [3, 0, 1, 2, 4], ids=["PLANE", "NONE", "MIN_EIG", "NORMALIZED_MIN_EIG", "FROBENIUS"])
def test_covariances(small_pair, reg):
    src, tgt, _ = small_pair
    g = _gpu(LAUNCH_PARAMS, regularization=reg)
    o = _oracle(LAUNCH_PARAMS, regularization=reg)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    assert o.compute_covariances() == 0
    for which, get in ((0, g.getSourceCovariances), (1, g.getTargetCovariances)):
        C = get()
        Cref = o.covariances(which)
        assert np.all(C[:, 3, :] == 0) and np.all(C[:, :, 3] == 0)  # Matrix4d layout: zero last row/column
        scale = np.abs(Cref).max(axis=(1, 2))
        err = np.abs(C[:, :3, :3] - Cref).max(axis=(1, 2)) / scale
        assert err.max() <= REL_TOL
        # same arithmetic in the same order: agreement is in fact at rounding level
        assert np.median(err) < 1e-12


# ---------------------------------------------------------------- linearize (evaluateCost)

def test_linearize_matches_oracle(small_pair):
    src, tgt, T_gt = small_pair
    g = _gpu(LAUNCH_PARAMS)
    o = _oracle(LAUNCH_PARAMS)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    for pose in (np.eye(4), T_gt):
        e, H, b = g.evaluateCost(pose)
        e0, H0, b0 = o.linearize(pose)
        corr, sq = g.getCorrespondences()
        corr0, sq0 = o.correspondences()
        assert np.array_equal(corr, corr0)
        m = corr0 >= 0
        assert m.sum() > 100 and (~m).sum() > 0
        assert np.array_equal(sq[m], sq0[m])
        assert abs(e - e0) <= REL_TOL * abs(e0)
        assert np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max()
        assert np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()
        assert np.array_equal(H, H.T)
        M = g.getMahalanobis()[:, :3, :3]
        M0 = o.mahalanobis()
        rel = np.abs(M[m] - M0[m]).max(axis=(1, 2)) / np.abs(M0[m]).max(axis=(1, 2))
        assert rel.max() <= REL_TOL
        assert np.all(M[~m] == 0)


def test_linearize_default_params_unbounded_gate(small_pair):
    # constructor defaults: max_corr_dist = FLT_MAX -> every point has a correspondence
    src, tgt, _ = small_pair
    g = _gpu()
    o = _oracle()
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    e, H, b = g.evaluateCost(np.eye(4))
    e0, H0, b0 = o.linearize(np.eye(4))
    corr, sq = g.getCorrespondences()
    corr0, sq0 = o.correspondences()
    assert (corr0 >= 0).all() and np.array_equal(corr, corr0) and np.array_equal(sq, sq0)
    assert np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max()
    assert np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()


# ---------------------------------------------------------------- align

@pytest.mark.parametrize("params", [LAUNCH_PARAMS, TIGHT_PARAMS], ids=["launch", "tight"])
@pytest.mark.parametrize("team", [0, 1, 2, 8])
def test_align_matches_oracle(small_pair, params, team):
    src, tgt, T_gt = small_pair
    g = _gpu(params)
    g.setOption("team_size", team)
    o = _oracle(params)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    out = g.align()
    rc, T0, conv0, it0 = o.align()
    assert rc == 0
    assert g.hasConverged() == conv0 and g.nr_iterations() == it0
    tr, tr0 = g.getLMTrace(), o.trace()
    assert tr.shape == tr0.shape
    assert np.array_equal(tr[:, [0, 1, 7]], tr0[:, [0, 1, 7]])  # same accept / reject sequence
    assert np.allclose(tr[:, 2:4], tr0[:, 2:4], rtol=1e-7)
    T = g.getFinalTransformation()
    _assert_same_transform(T, T0)
    f0 = o.fitness()
    assert abs(g.getFitnessScore() - f0) <= REL_TOL * f0
    assert abs(g.getFitnessScore(1.5) - o.fitness(1.5)) <= REL_TOL * o.fitness(1.5)
    assert np.array_equal(out, o.transform_source(T))  # pcl::transformPointCloud in float
    Hf, Hf0 = g.getFinalHessian(), o.final_hessian()
    assert np.abs(Hf - Hf0).max() <= REL_TOL * np.abs(Hf0).max()
    assert g.result().num_inliers == int((o.correspondences()[0] >= 0).sum())


def test_align_with_guess_and_5k(pair5k):
    src, tgt, T_gt = pair5k
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = T_gt[:3, 3] * 0.5
    for params in (LAUNCH_PARAMS, TIGHT_PARAMS):
        g = _gpu(params)
        o = _oracle(params)
        g.setInputSource(src); g.setInputTarget(tgt)
        o.set_source(src); o.set_target(tgt)
        g.align(guess, want_output=False)
        rc, T0, conv0, it0 = o.align(guess)
        assert g.hasConverged() == conv0 and g.nr_iterations() == it0
        _assert_same_transform(g.getFinalTransformation(), T0)
        assert np.abs(g.getFinalTransformation()[:3, 3] - T_gt[:3, 3]).max() < 0.3


def test_align_gauss_newton(small_pair):
    src, tgt, _ = small_pair
    g = _gpu(TIGHT_PARAMS, optimizer=0, max_iterations=12)
    o = _oracle(TIGHT_PARAMS, optimizer=0, max_iterations=12)
    g.setInputSource(src); g.setInputTarget(tgt)
    o.set_source(src); o.set_target(tgt)
    g.align(want_output=False)
    rc, T0, conv0, it0 = o.align()
    assert g.hasConverged() == conv0 and g.nr_iterations() == it0
    _assert_same_transform(g.getFinalTransformation(), T0)


def test_align_unstaged_target_and_large_target():
    from riv_slam_b200 import datagen
    src, tgt, _ = datagen.make_pair(3, 1, n_src=2500, n_tgt=20000)
    o = _oracle(LAUNCH_PARAMS)
    o.set_source(src); o.set_target(tgt)
    rc, T0, conv0, it0 = o.align()
    for team in (0, 1, 4):
        g = _gpu(LAUNCH_PARAMS)
        g.setOption("team_size", team)
        g.setInputSource(src); g.setInputTarget(tgt)
        g.align(want_output=False)
        assert g.hasConverged() == conv0 and g.nr_iterations() == it0
        _assert_same_transform(g.getFinalTransformation(), T0)
        assert abs(g.getFitnessScore() - o.fitness()) <= REL_TOL * o.fitness()


def test_edge_cases(small_pair):
    from riv_slam_b200 import fast_apdgicp as F
    src, tgt, _ = small_pair
    g = _gpu(LAUNCH_PARAMS)
    assert g.align() is None and not g.hasConverged()          # no clouds: pcl prints and returns
    assert g.status() == F.APD_ERR_NO_INPUT
    assert np.array_equal(g.getFinalTransformation(), np.eye(4, dtype=np.float32))
    g.setInputSource(src[:10]); g.setInputTarget(tgt)
    assert g.align() is None and g.status() == F.APD_ERR_TOO_FEW_POINTS
    # far-away source: nothing inside the 2 m gate -> H = b = 0, delta = I, converged at iteration 0
    far = src.copy()
    far[:, 0] += 1000.0
    g.setInputSource(far)
    g.align(want_output=False)
    assert g.hasConverged() and g.nr_iterations() == 0
    assert np.array_equal(g.getFinalTransformation(), np.eye(4, dtype=np.float32))
    assert (g.getCorrespondences()[0] < 0).all()
    o = _oracle(LAUNCH_PARAMS)
    o.set_source(far); o.set_target(tgt)
    o.align()
    assert abs(g.getFitnessScore() - o.fitness()) <= REL_TOL * o.fitness()
    # k above the supported range is refused loudly, never silently clamped
    with pytest.raises(F.ApdError):
        g.setCorrespondenceRandomness(64)


def test_swap_cache_and_injected_covariances(small_pair):
    src, tgt, _ = small_pair
    g = _gpu(TIGHT_PARAMS)
    g.setInputSource(src); g.setInputTarget(tgt)
    g.align(want_output=False)
    T_fwd = g.getFinalTransformation()
    g.swapSourceAndTarget()
    g.align(want_output=False)
    T_bwd = g.getFinalTransformation()
    o = _oracle(TIGHT_PARAMS)
    o.set_source(tgt); o.set_target(src)
    _, T0, conv0, it0 = o.align()
    assert g.hasConverged() == conv0 and g.nr_iterations() == it0
    _assert_same_transform(T_bwd, T0)
    assert np.abs(T_fwd.astype(np.float64) @ T_bwd.astype(np.float64) - np.eye(4)).max() < 0.05
    # odometry pattern: the previous source becomes the target (same cache key -> shared device data)
    g2 = _gpu(TIGHT_PARAMS)
    g2.setInputSource(tgt, cache_key=11); g2.setInputTarget(src, cache_key=12)
    launches0 = g2.handle().launch_count()
    g2.align(want_output=False)
    per_align = g2.handle().launch_count() - launches0
    g2.setInputTarget(tgt, cache_key=11)   # held by the source slot: no rebuild
    g2.setInputSource(src, cache_key=12)   # held by the (old) target slot... now replaced: rebuild allowed
    g2.swapSourceAndTarget()
    # injected covariances are used as given (setSourceCovariances / setTargetCovariances)
    g3 = _gpu(TIGHT_PARAMS)
    g3.setInputSource(src); g3.setInputTarget(tgt)
    Cs, Ct = g.getTargetCovariances(), g.getSourceCovariances()  # g is swapped: its target is `src`
    g3.setSourceCovariances(Cs); g3.setTargetCovariances(Ct)
    assert np.array_equal(g3.getSourceCovariances(), Cs)
    g3.align(want_output=False)
    _assert_same_transform(g3.getFinalTransformation(), T_fwd)
    assert per_align > 0


# ---------------------------------------------------------------- batched path

def _oracle_results(pairs, params, guesses=None):
    out = []
    for i, (s, t) in enumerate(pairs):
        o = _oracle(params)
        o.set_source(s); o.set_target(t)
        rc, T, conv, it = o.align(None if guesses is None else guesses[i])
        out.append((rc, T, conv, it, o.fitness() if rc == 0 else None))
    return out


def test_batch_align_matches_oracle():
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs, batch_align
    sizes = [(900, 1000), (1500, 1400), (1200, 1200), (2000, 1800), (1000, 2200), (1300, 900), (1100, 1100)]
    pairs = [datagen.make_pair(4, 100 + i, n_src=a, n_tgt=b)[:2] for i, (a, b) in enumerate(sizes)]
    ref = _oracle_results(pairs, LAUNCH_PARAMS)
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    res = batch_align(H, [p[0] for p in pairs], [p[1] for p in pairs])
    lin, err, n = H.work_counters()
    assert n == len(pairs) and lin >= len(pairs) and err >= lin
    for r, (rc, T0, conv0, it0, f0) in zip(res, ref):
        assert r["status"] == 0 and bool(r["converged"]) == conv0 and r["iterations"] == it0
        _assert_same_transform(r["T"], T0)
        assert abs(r["fitness"] - f0) <= REL_TOL * f0
    # cloud sets + explicit index maps, every team shape, staged or not: same answers
    S = CloudSet(H, [p[0] for p in pairs])
    T = CloudSet(H, [p[1] for p in pairs])
    idx = np.array([3, 0, 6, 2], dtype=np.int32)
    for team, unstaged in ((0, 0), (1, 0), (4, 0), (1, 1), (2, 1)):
        H.set_option("team_size", team)
        H.set_option("force_unstaged", unstaged)
        S2 = CloudSet(H, [p[0] for p in pairs]) if unstaged else S
        T2 = CloudSet(H, [p[1] for p in pairs]) if unstaged else T
        r2 = align_pairs(H, S2, T2, src_idx=idx, tgt_idx=idx)
        for j, i in enumerate(idx):
            assert bool(r2[j]["converged"]) == ref[i][2] and r2[j]["iterations"] == ref[i][3]
            _assert_same_transform(r2[j]["T"], ref[i][1])
    H.set_option("team_size", 0)
    H.set_option("force_unstaged", 0)
    # ragged batch with an empty and an undersized cloud: per-pair status, the others unaffected
    from riv_slam_b200 import fast_apdgicp as F
    srcs = [pairs[0][0], pairs[1][0][:0], pairs[2][0][:7], pairs[3][0]]
    tgts = [pairs[0][1], pairs[1][1], pairs[2][1], pairs[3][1]]
    r3 = batch_align(H, srcs, tgts)
    assert r3[1]["status"] == F.APD_ERR_NO_INPUT and r3[2]["status"] == F.APD_ERR_TOO_FEW_POINTS
    assert not r3[1]["converged"] and not r3[2]["converged"]
    for j, i in ((0, 0), (3, 3)):
        _assert_same_transform(r3[j]["T"], ref[i][1])


def test_odometry_chain_shared_set():
    """Scan-to-scan odometry over one cloud set: scan t+1 -> scan t, guesses chained on the host."""
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs
    scans, poses = datagen.make_sequence(2, 0, n_scans=5, n_points=1500)
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    S = CloudSet(H, scans)
    n = len(scans) - 1
    res = align_pairs(H, S, S, src_idx=np.arange(1, n + 1), tgt_idx=np.arange(0, n))
    for t in range(n):
        o = _oracle(LAUNCH_PARAMS)
        o.set_source(scans[t + 1]); o.set_target(scans[t])
        rc, T0, conv0, it0 = o.align()
        assert bool(res[t]["converged"]) == conv0 and res[t]["iterations"] == it0
        _assert_same_transform(res[t]["T"], T0)
        gt = datagen.relative_gt(poses, t)
        assert np.abs(res[t]["T"][:3, 3] - gt[:3, 3]).max() < 0.3


def test_full_size_properties(pair5k):
    """Size-independent properties at the benchmark size (5k points)."""
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs
    src, tgt, T_gt = pair5k
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    S = CloudSet(H, [src] * 6)
    T = CloudSet(H, [tgt] * 6)
    r1 = align_pairs(H, S, T)
    r2 = align_pairs(H, S, T)
    assert r1.tobytes() == r2.tobytes()                     # deterministic
    assert all(r1[i].tobytes() == r1[0].tobytes() for i in range(6))   # identical pairs -> identical records
    assert r1[0]["converged"] and np.abs(r1[0]["T"][:3, 3] - T_gt[:3, 3]).max() < 0.3
    # idempotence: restarting from the converged transform stops at once and stays put
    r3 = align_pairs(H, S, T, guesses=np.stack([r1[i]["T"] for i in range(6)]))
    assert all(r3["converged"]) and r3["iterations"].max() <= 1
    assert np.abs(r3[0]["T"] - r1[0]["T"]).max() < 0.05
    # a rigidly moved copy of the target registers back onto it (encode -> decode round trip)
    from riv_slam_b200 import datagen
    Tm = datagen.pose_matrix([0.3, -0.1, 0.05], [0.2, -0.3, 1.0])
    moved = tgt.copy()
    moved[:, :3] = (tgt[:, :3].astype(np.float64) - Tm[:3, 3]) @ Tm[:3, :3]   # p_src = Tm^-1 p_tgt
    H.set_params(**TIGHT_PARAMS)
    r4 = align_pairs(H, CloudSet(H, [moved]), CloudSet(H, [tgt]))
    assert r4[0]["converged"]
    assert np.abs(r4[0]["T"].astype(np.float64) - Tm).max() < 2e-3
    assert r4[0]["fitness"] < 1e-5


def test_drive_parity_statistics():
    """SURVEY 8(d) parity gates over a drive at the benchmark size: 96 consecutive 5000-point pairs (config C2, identity
    guess) and the same pairs with chained guesses. Every pair: same converged flag and iteration count as the oracle,
    transform within 1e-5 rad / 1e-4 m, fitness within 1e-5 relative."""
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import Handle, odometry_align
    n_pairs = 96
    scans, poses = datagen.make_drive(2, 5, n_pairs + 1, 5000)  # generated in-process: no fork next to a live CUDA context
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    res = odometry_align(H, scans)
    o = _oracle(LAUNCH_PARAMS)
    ref, guesses = [], [np.eye(4, dtype=np.float32)]
    for t in range(n_pairs):
        if t == 0:
            o.set_target(scans[0])
        else:
            o.swap()  # the previous source becomes the target with its covariances
        o.set_source(scans[t + 1])
        rc, T0, conv0, it0 = o.align()
        ref.append((T0, conv0, it0, o.fitness()))
        guesses.append(T0.astype(np.float32))  # constant-velocity guess for the next pair (SMO:461-465 without the ego-velocity term)
    bad = []
    for t, (T0, conv0, it0, f0) in enumerate(ref):
        if bool(res[t]["converged"]) != conv0 or int(res[t]["iterations"]) != it0:
            bad.append((t, int(res[t]["iterations"]), it0))
        _assert_same_transform(res[t]["T"], T0)
        assert abs(res[t]["fitness"] - f0) <= 1e-5 * f0
    assert not bad, bad
    # chained guesses: pair t starts from the oracle's result of pair t-1
    res2 = odometry_align(H, scans, guesses=np.stack(guesses[:n_pairs]))
    o2 = _oracle(LAUNCH_PARAMS)
    for t in range(0, n_pairs, 4):
        o2.set_source(scans[t + 1]); o2.set_target(scans[t])
        rc, T0, conv0, it0 = o2.align(guesses[t])
        assert bool(res2[t]["converged"]) == conv0 and int(res2[t]["iterations"]) == it0
        _assert_same_transform(res2[t]["T"], T0)


# ---------------------------------------------------------------- large clouds (configs C3 / C5)

def test_large_clouds_grid_team():
    """Source and target far beyond the shared-memory staging limit: global-memory grid, the
    cooperative whole-GPU team (TEAM_GRID), and every other team shape must agree with the oracle."""
    from riv_slam_b200 import datagen
    src, tgt, _ = datagen.make_pair(5, 0, n_src=40000, n_tgt=70000, voxel=None)
    o = _oracle(LAUNCH_PARAMS)
    o.set_source(src); o.set_target(tgt)
    rc, T0, conv0, it0 = o.align()
    assert rc == 0
    knn_ref = o.knn(0)
    for team in (0, 1, 8):
        g = _gpu(LAUNCH_PARAMS)
        g.setOption("team_size", team)
        g.setInputSource(src); g.setInputTarget(tgt)
        assert np.array_equal(g.getKnn(0), knn_ref)
        g.align(want_output=False)
        assert g.hasConverged() == conv0 and g.nr_iterations() == it0
        _assert_same_transform(g.getFinalTransformation(), T0)
        assert abs(g.getFitnessScore() - o.fitness()) <= REL_TOL * o.fitness()
        e, H, b = g.evaluateCost(T0)
        e0, H0, b0 = o.linearize(T0)
        assert np.abs(H - H0).max() <= REL_TOL * np.abs(H0).max() and np.abs(b - b0).max() <= REL_TOL * np.abs(b0).max()


def test_scan_to_submap_reuses_target():
    """Config C3: several scans against one large accumulated target; the target is prepared once."""
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs
    scans, poses = datagen.make_sequence(3, 0, n_scans=6, n_points=3000)
    # submap = first four scans moved into the frame of scan 0
    sub = []
    for t in range(4):
        Trel = np.linalg.inv(poses[0]) @ poses[t]
        p = scans[t].copy()
        p[:, :3] = (scans[t][:, :3].astype(np.float64) @ Trel[:3, :3].T + Trel[:3, 3]).astype(np.float32)
        sub.append(p)
    submap = np.concatenate(sub)
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    S = CloudSet(H, scans[4:6])
    T = CloudSet(H, [submap])
    guesses = np.stack([(np.linalg.inv(poses[0]) @ poses[3]).astype(np.float32)] * 2)   # last known pose as the guess (SMO:461-465)
    res = align_pairs(H, S, T, tgt_idx=np.zeros(2, dtype=np.int32), guesses=guesses)
    for j in range(2):
        o = _oracle(LAUNCH_PARAMS)
        o.set_source(scans[4 + j]); o.set_target(submap)
        rc, T0, conv0, it0 = o.align(guesses[j])
        assert bool(res[j]["converged"]) == conv0 and res[j]["iterations"] == it0
        _assert_same_transform(res[j]["T"], T0)
        # sanity only: with the launch-file epsilon (0.1 m) the loop stops as soon as a step is below 10 cm
        gt = np.linalg.inv(poses[0]) @ poses[4 + j]
        assert np.abs(res[j]["T"][:3, 3] - gt[:3, 3]).max() < 1.5


def test_pipelined_host_batches_equal_one_launch():
    """apd_odometry_align / apd_batch_align cut big host batches into chunks over two streams; the records
    must equal those of a single apd_align_pairs launch bit for bit."""
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs, batch_align, odometry_align
    base, _ = datagen.make_drive(2, 3, 24, 700, workers=0)
    order = []
    t, d = 0, 1
    for _ in range(601):                       # 600 pairs > 2 chunks of 256
        order.append(t)
        if t + d < 0 or t + d >= len(base):
            d = -d
        t += d
    scans = [base[i] for i in order]
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    S = CloudSet(H, scans)
    n = len(scans) - 1
    ref = align_pairs(H, S, S, src_idx=np.arange(1, n + 1), tgt_idx=np.arange(0, n))
    got = odometry_align(H, scans)
    assert got.tobytes() == ref.tobytes()
    lin, err, pairs = H.work_counters()
    assert pairs == n and lin >= n
    got2 = batch_align(H, scans[1:], scans[:-1])
    assert got2.tobytes() == ref.tobytes()
    # guesses are routed per chunk
    g = np.stack([ref[i]["T"] for i in range(n)])
    got3 = odometry_align(H, scans, guesses=g)
    ref3 = align_pairs(H, S, S, src_idx=np.arange(1, n + 1), tgt_idx=np.arange(0, n), guesses=g)
    assert got3.tobytes() == ref3.tobytes()


def test_information_matrix_fitness_score(small_pair):
    """InformationMatrixCalculator::calc_fitness_score / calc_information_matrix."""
    from riv_slam_b200.information_matrix import InformationMatrixCalculator
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, fitness_pairs
    src, tgt, T_gt = small_pair
    calc = InformationMatrixCalculator()
    o = _oracle()
    o.set_source(src); o.set_target(tgt)
    for pose, rng in ((np.eye(4), 1e300), (T_gt, 1e300), (T_gt, 1.0), (T_gt, 1e-9)):
        f = calc.calc_fitness_score(tgt, src, pose, max_range=rng)
        f0 = o.fitness_score(pose, rng)
        assert abs(f - f0) <= REL_TOL * abs(f0)
    assert calc.calc_fitness_score(tgt, src, T_gt, max_range=-1.0) == np.finfo(np.float64).max   # nothing in range
    inf = calc.calc_information_matrix(tgt, src, T_gt)
    f0 = o.fitness_score(T_gt)
    w = lambda lo, hi: np.float32(1e-8 * (lo ** 2 + (hi ** 2 - lo ** 2) * (1 - np.exp(-20.0 * f0)) / (1 - np.exp(-20.0 * 0.5))))
    assert np.allclose(np.diag(inf)[:3], 1.0 / float(w(0.1, 5.0)), rtol=1e-5) and np.allclose(np.diag(inf)[3:], 1.0 / float(w(0.05, 0.2)), rtol=1e-5)
    # batched over cloud sets
    H = Handle(0)
    S, T = CloudSet(H, [src, src[:700]]), CloudSet(H, [tgt])
    poses = np.stack([T_gt, np.eye(4)]).astype(np.float32)
    sc = fitness_pairs(H, S, T, tgt_idx=[0, 0], poses=poses, max_range=4.0)
    o2 = _oracle()
    o2.set_source(src[:700]); o2.set_target(tgt)
    assert abs(sc[0] - o.fitness_score(T_gt, 4.0)) <= REL_TOL * sc[0] and abs(sc[1] - o2.fitness_score(np.eye(4), 4.0)) <= REL_TOL * sc[1]


def test_fused_and_multikernel_grid_builds_agree(pair5k):
    """The one-launch build for small clouds and the multi-kernel pipeline must give identical registrations."""
    from riv_slam_b200.fast_apdgicp import Handle, CloudSet, align_pairs
    src, tgt, _ = pair5k
    out = []
    for fused in (1, 0):
        H = Handle(0)
        H.set_params(**TIGHT_PARAMS)
        H.set_option("fused_build", fused)
        out.append(align_pairs(H, CloudSet(H, [src, src[:3000]]), CloudSet(H, [tgt, tgt[:2500]])).tobytes())
    assert out[0] == out[1]


def test_batched_loop_candidate_matching():
    """LoopDetector::matching (loop_detector.cpp:379-441) over all candidates in one launch."""
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import Handle
    from riv_slam_b200.loop_matching import matching, relative_guess
    scans, poses = datagen.make_sequence(4, 5, n_scans=7, n_points=1500)
    new_kf, new_pose = scans[3], poses[3]
    cands = [scans[i] for i in (0, 1, 2, 4, 5, 6)]
    cand_poses = [poses[i] for i in (0, 1, 2, 4, 5, 6)]
    guesses = np.stack([relative_guess(new_pose, p) for p in cand_poses])
    H = Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    best, rel, score, res = matching(H, cands, new_kf, guesses, fitness_score_max_range=4.0, fitness_score_thresh=6.0)
    # the reference's sequential loop on the CPU oracle
    best0, score0, rel0 = None, np.finfo(np.float64).max, None
    for i, c in enumerate(cands):
        o = _oracle(LAUNCH_PARAMS)
        o.set_target(new_kf); o.set_source(c)
        rc, T0, conv0, it0 = o.align(guesses[i])
        s0 = o.fitness(4.0)
        assert bool(res[i]["converged"]) == conv0 and res[i]["iterations"] == it0
        assert abs(res[i]["fitness"] - s0) <= REL_TOL * s0
        if not conv0 or s0 > score0:
            continue
        best0, score0, rel0 = i, s0, T0
    assert best == best0 and abs(score - score0) <= REL_TOL * score0
    _assert_same_transform(rel, rel0)
    # a threshold below the best score rejects the loop
    assert matching(H, cands, new_kf, guesses, 4.0, fitness_score_thresh=score0 * 0.5)[0] is None
    assert matching(H, [], new_kf)[0] is None


def _degenerate_clouds():
    rng = np.random.default_rng(11)
    n = 400
    yield "planar_z0", np.c_[rng.uniform(-20, 20, (n, 2)), np.zeros(n)].astype(np.float32)
    yield "collinear", np.c_[np.linspace(0, 50, n), np.full(n, 3.0), np.full(n, -1.0)].astype(np.float32)
    yield "all_identical", np.tile(np.array([[1.5, -2.0, 0.25]], np.float32), (n, 1))
    yield "duplicates", np.repeat(rng.uniform(-5, 5, (n // 4, 3)).astype(np.float32), 4, axis=0)
    yield "huge_coordinates", (rng.uniform(-1, 1, (n, 3)) + np.array([4.0e5, -3.0e5, 1.0e5])).astype(np.float32)
    yield "tiny_extent", (rng.uniform(0, 1e-4, (n, 3)) + 7.0).astype(np.float32)
    yield "two_far_clusters", np.r_[rng.normal(0, 0.2, (n // 2, 3)), rng.normal(0, 0.2, (n // 2, 3)) + 900.0].astype(np.float32)
    yield "exactly_k_points", rng.uniform(-3, 3, (20, 3)).astype(np.float32)
    yield "k_plus_one", rng.uniform(-3, 3, (21, 3)).astype(np.float32)
    yield "one_outlier", np.r_[rng.uniform(-2, 2, (n - 1, 3)), [[5000.0, 5000.0, 5000.0]]].astype(np.float32)


@pytest.mark.parametrize("unstaged", [0, 1])
def test_knn_exact_on_degenerate_clouds(unstaged):
    """Index sets stay bit-exact where a grid is at its worst: flat, collinear, duplicated, far-apart, tiny."""
    from oracle.oracle import knn_bruteforce
    for name, cloud in _degenerate_clouds():
        for cpp in (4.0, 0.05, 200.0):
            g = _gpu(k_correspondences=20)
            g.setOption("cells_per_point", cpp)
            g.setOption("force_unstaged", unstaged)
            g.setInputSource(cloud)
            ref, _ = knn_bruteforce(cloud, cloud, 20)
            got = g.getKnn(0)
            assert np.array_equal(got, ref), (name, cpp, unstaged, int((got != ref).any(axis=1).sum()))
    # and a 1-NN / gate check against the same clouds as targets
    rng = np.random.default_rng(3)
    for name, cloud in _degenerate_clouds():
        q = (cloud[rng.integers(0, len(cloud), 300)] + rng.normal(0, 0.3, (300, 3))).astype(np.float32)
        g = _gpu(LAUNCH_PARAMS)
        g.setOption("force_unstaged", unstaged)
        g.setInputSource(q); g.setInputTarget(cloud)
        o = _oracle(LAUNCH_PARAMS)
        o.set_source(q); o.set_target(cloud)
        if len(cloud) < 20 or len(q) < 20:
            continue
        g.evaluateCost(np.eye(4))
        o.linearize(np.eye(4))
        corr, sq = g.getCorrespondences()
        corr0, sq0 = o.correspondences()
        assert np.array_equal(corr, corr0), name
        assert np.array_equal(sq[corr0 >= 0], sq0[corr0 >= 0]), name
