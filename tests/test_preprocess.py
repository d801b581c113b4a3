""""Next" rows SURVEY.md §8(f)-4 (preprocessing filters) and §8(f)-2 (submap accumulation).

CPU part (not gpu): the C++ oracle (oracle/preprocess_oracle.hpp) against an independent numpy /
pure-Python twin written from the same reference lines (preprocessing_nodelet.cpp:812-815,850-896,
scan_matching_odometry_nodelet.cpp:606-616) and PCL's published filter behaviour.
GPU part: apd_preprocess / apd_build_submap through the C ABI against the oracle, bit-exact
(same points, same order, same float bits).
"""
import numpy as np
import pytest

from conftest import LAUNCH_PARAMS


# ---------------------------------------------------------------- inputs

def raw_scan(seed, n=3000, with_bad=True):
    """A raw radar-like cloud (x y z intensity): several returns per 0.1 m voxel, near/far/z outliers, NaN/inf rows."""
    from riv_slam_b200 import datagen
    src, _, _ = datagen.make_pair(9, seed, n_src=n)
    rng = np.random.default_rng(1234 + seed)
    base = np.asarray(src[:, :3], dtype=np.float32)
    reps = [base]
    for _ in range(2):  # extra returns of the same scatterers -> voxels with 2-3 points
        keep = rng.random(len(base)) < 0.5
        reps.append(base[keep] + rng.normal(0, 0.02, (int(keep.sum()), 3)).astype(np.float32))
    xyz = np.concatenate(reps)
    lone = rng.uniform([-20, -40, -8], [110, 40, 25], (200, 3)).astype(np.float32)  # isolated clutter, some outside the z band
    near = rng.normal(0, 0.4, (50, 3)).astype(np.float32)                           # inside distance_near_thresh
    xyz = np.concatenate([xyz, lone, near])
    pts = np.concatenate([xyz, rng.uniform(0, 40, (len(xyz), 1)).astype(np.float32)], axis=1)
    pts = pts[rng.permutation(len(pts))]
    if with_bad:
        bad = rng.choice(len(pts), 12, replace=False)
        pts[bad[:4], 0] = np.nan
        pts[bad[4:8], 1] = np.inf
        pts[bad[8:], 2] = -np.inf
    return np.ascontiguousarray(pts, dtype=np.float32)


def oracle_pipeline(cloud, use_distance_filter=1, distance_near_thresh=1.0, distance_far_thresh=100.0, z_low_thresh=-5.0, z_high_thresh=20.0,
                    downsample_resolution=0.1, outlier_removal=1, radius_radius=0.8, radius_min_neighbors=2, statistical_mean_k=20, statistical_stddev=1.0,
                    downsample_method="VOXELGRID"):
    from oracle import oracle as O
    c = cloud
    if use_distance_filter:
        c = O.distance_filter(c, distance_near_thresh, distance_far_thresh, z_low_thresh, z_high_thresh)
    if downsample_resolution > 0:
        c = O.voxel_grid(c, downsample_resolution) if downsample_method == "VOXELGRID" else O.approx_voxel_grid(c, downsample_resolution)
    else:
        c = c[np.isfinite(c[:, :3]).all(axis=1)]  # pcl::removeNaNFromPointCloud, preprocessing_nodelet.cpp:852-856
    if outlier_removal == 1:
        c = O.radius_outlier_removal(c, radius_radius, radius_min_neighbors)
    elif outlier_removal == 2:
        c = O.statistical_outlier_removal(c, statistical_mean_k, statistical_stddev)
    return c


# ---------------------------------------------------------------- twins (independent of the C++ oracle)

def twin_distance_filter(c, near, far, zlo, zhi):
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    with np.errstate(invalid="ignore", over="ignore"):
        d = np.sqrt(((x * x).astype(np.float32) + (y * y).astype(np.float32)).astype(np.float32) + (z * z).astype(np.float32), dtype=np.float32).astype(np.float64)
        keep = (d > near) & (d < far) & (z.astype(np.float64) < zhi) & (z.astype(np.float64) > zlo)
    return c[keep]


def twin_voxel_grid(c, leaf):
    f32 = np.float32
    inv = f32(1.0) / f32(leaf)
    fin = np.isfinite(c[:, :3]).all(axis=1)
    p = c[fin]
    mn, mx = p[:, :3].min(axis=0), p[:, :3].max(axis=0)
    minb = np.floor(mn * inv).astype(np.int64)
    div = np.floor(mx * inv).astype(np.int64) - minb + 1
    ijk = (np.floor(p[:, :3] * inv) - minb.astype(f32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    groups = {}
    for i, k in enumerate(key.tolist()):  # ascending input order inside a voxel
        groups.setdefault(k, []).append(i)
    out = []
    for k in sorted(groups):
        s = np.zeros(4, dtype=f32)
        for i in groups[k]:
            s = (s + p[i]).astype(f32)
        out.append(s / f32(len(groups[k])))
    return np.array(out, dtype=f32).reshape(-1, 4)


def _approx_cells(c, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    with np.errstate(invalid="ignore", over="ignore"):
        f = np.floor(c[:, :3].astype(np.float32) * inv)
    ok = np.isfinite(f) & (f >= -2147483648.0) & (f < 2147483648.0)
    cells = np.where(ok, np.where(ok, f, 0).astype(np.int64), -2**31)              # cvttss2si: INT_MIN when not representable
    h = ((cells[:, 0] * 7171 + cells[:, 1] * 3079 + cells[:, 2] * 4231) & 0xFFFFFFFF) & 511
    return cells, h


def twin_approx_voxel_grid(c, leaf):
    """pcl::ApproximateVoxelGrid as PCL writes it: a pure-Python walk over the points with the 512-entry history table."""
    cells, hs = _approx_cells(c, leaf)
    table = {}
    out = []

    def flush(e):
        n = np.float32(e[1])
        out.append(e[2] / n)
    for i in range(len(c)):
        key = tuple(cells[i]); h = int(hs[i])
        e = table.get(h)
        if e is not None and e[0] != key:
            flush(e)
            e = None
        if e is None:
            e = [key, 0, np.zeros(4, np.float32)]
            table[h] = e
        e[1] += 1
        e[2] = e[2] + c[i].astype(np.float32)
    for h in sorted(table):
        flush(table[h])
    return np.array(out, dtype=np.float32).reshape(-1, 4)


def parallel_approx_voxel_grid(c, leaf):
    """The data-parallel form the CUDA kernel uses (apd_preprocess.cu, approx_voxel_grid_kernel), in numpy: stable sort by table entry,
    runs of equal voxels, a flush slot per run from an exclusive scan of the flushing points in input order."""
    n = len(c)
    if n == 0:
        return np.zeros((0, 4), np.float32)
    cells, hs = _approx_cells(c, leaf)
    order = np.argsort(hs, kind="stable")
    sh, sc = hs[order], cells[order]
    first = np.r_[True, sh[1:] != sh[:-1]]
    head = first | np.r_[False, (sc[1:] != sc[:-1]).any(axis=1)]
    flushing = np.zeros(n, np.int64)
    flushing[order[head & ~first]] = 1
    before = np.cumsum(flushing) - flushing                     # exclusive scan in INPUT order
    n_flush = int(flushing.sum())
    entry_rank = np.cumsum(first) - 1                           # rank of the entry among the non-empty ones, per sorted position
    starts = np.flatnonzero(head)
    ends = np.r_[starts[1:], n]
    out = np.zeros((len(starts), 4), np.float32)
    for a, b in zip(starts, ends):
        acc = np.zeros(4, np.float32)
        for j in range(a, b):
            acc = acc + c[order[j]].astype(np.float32)
        slot = before[order[b]] if b < n and sh[b] == sh[a] else n_flush + entry_rank[a]
        out[slot] = acc / np.float32(b - a)
    return out


def twin_radius_outlier(c, radius, min_pts):
    x = c[:, :3]
    d = (x[:, None, :] - x[None, :, :]).astype(np.float32)
    d2 = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32) + (d[..., 2] * d[..., 2]).astype(np.float32)
    if len(c) < min_pts + 1:
        return c[:0]
    kth = np.partition(d2, min_pts, axis=1)[:, min_pts].astype(np.float64)
    return c[~(radius * radius < kth)]


def twin_statistical_outlier(c, mean_k, stddev_mult):
    if len(c) < mean_k + 1:
        return c
    x = c[:, :3]
    d = (x[:, None, :] - x[None, :, :]).astype(np.float32)
    d2 = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32) + (d[..., 2] * d[..., 2]).astype(np.float32)
    nn = np.sort(d2, axis=1)[:, 1:mean_k + 1].astype(np.float64)
    acc = np.zeros(len(c))
    for k in range(mean_k):            # ascending neighbours, one addition at a time
        acc = acc + np.sqrt(nn[:, k])
    dist = (acc / mean_k).astype(np.float32)
    s = q = 0.0
    for v in dist:                     # index order, like the reference loop
        s += float(v)
        q += float(np.float32(v * v))
    n = len(c)
    thr = s / n + stddev_mult * np.sqrt((q - s * s / n) / (n - 1.0))
    return c[~(dist.astype(np.float64) > thr)]


# ---------------------------------------------------------------- CPU: oracle vs twins

@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_distance_filter_matches_twin(seed):
    from oracle import oracle as O
    c = raw_scan(seed, 1500)
    for args in [(1.0, 100.0, -5.0, 20.0), (2.0, 50.0, -2.0, 6.0)]:
        assert np.array_equal(O.distance_filter(c, *args), twin_distance_filter(c, *args))


@pytest.mark.parametrize("leaf", [0.1, 0.5, 2.0])
def test_oracle_voxel_grid_matches_twin(leaf):
    from oracle import oracle as O
    c = raw_scan(2, 1500)
    got, want = O.voxel_grid(c, leaf), twin_voxel_grid(c, leaf)
    assert got.shape == want.shape and got.tobytes() == want.tobytes()
    # every output point is the centroid of a distinct voxel: idempotent up to centroid rounding in count
    assert len(O.voxel_grid(got, leaf)) == len(got)


def test_oracle_voxel_grid_edge_cases():
    from oracle import oracle as O
    assert O.voxel_grid(np.zeros((0, 4), np.float32), 0.1).shape == (0, 4)
    nan = np.full((5, 4), np.nan, np.float32)
    assert O.voxel_grid(nan, 0.1).shape == (0, 4)
    one = np.array([[1, 2, 3, 4]], np.float32)
    assert np.array_equal(O.voxel_grid(one, 0.1), one)
    # leaf too small for 32-bit voxel indices: PCL warns and returns the input cloud unchanged
    far = np.array([[0, 0, 0, 1], [9000, 9000, 9000, 2]], np.float32)
    assert np.array_equal(O.voxel_grid(far, 0.001), far)


@pytest.mark.parametrize("leaf", [0.1, 0.5, 2.0, 25.0])
def test_oracle_approx_voxel_grid_matches_twins(leaf):
    """pcl::ApproximateVoxelGrid: the C++ oracle == a pure-Python walk with the history table == the data-parallel form of the kernel.
    Finite input (what the nodelet's distance filter leaves) and raw input with NaN / inf rows (use_distance_filter = false)."""
    from oracle import oracle as O
    for seed, with_bad in ((0, False), (1, False), (2, True)):
        c = raw_scan(seed, 1200, with_bad=with_bad)
        got = O.approx_voxel_grid(c, leaf)
        want = twin_approx_voxel_grid(c, leaf)
        assert got.shape == want.shape and got.tobytes() == want.tobytes()
        par = parallel_approx_voxel_grid(c, leaf)
        assert par.shape == got.shape and par.tobytes() == got.tobytes()
        if not with_bad:
            assert len(got) >= len(O.voxel_grid(c, leaf))       # evictions split voxels, they never merge them
    # collisions are what makes it approximate: a voxel that is evicted and comes back yields two output points
    a = np.array([[0.05, 0.05, 0.05, 1.0]], np.float32)
    cells, h = _approx_cells(a, 0.1)
    k = next(k for k in range(1, 4000) if ((k * 7171) & 511) == 0)                 # same table entry, different voxel
    b = a + np.array([[0.1 * k, 0, 0, 1.0]], np.float32)
    assert _approx_cells(b, 0.1)[1][0] == h[0]
    seq = np.concatenate([a, b, a, a]).astype(np.float32)
    got = O.approx_voxel_grid(seq, 0.1)
    assert len(got) == 3 and np.array_equal(got[0], a[0]) and np.array_equal(got[1], b[0]) and np.array_equal(got[2], a[0])
    assert parallel_approx_voxel_grid(seq, 0.1).tobytes() == got.tobytes()
    assert O.approx_voxel_grid(np.zeros((0, 4), np.float32), 0.1).shape == (0, 4)


@pytest.mark.parametrize("radius,min_pts", [(0.8, 2), (0.5, 5), (0.3, 1)])
def test_oracle_radius_outlier_matches_twin(radius, min_pts):
    from oracle import oracle as O
    c = O.voxel_grid(raw_scan(3, 800, with_bad=False), 0.1)
    got, want = O.radius_outlier_removal(c, radius, min_pts), twin_radius_outlier(c, radius, min_pts)
    assert np.array_equal(got, want)
    assert 0 < len(got) < len(c)


@pytest.mark.parametrize("mean_k,mult", [(20, 1.0), (8, 0.5), (30, 1.2)])
def test_oracle_statistical_outlier_matches_twin(mean_k, mult):
    from oracle import oracle as O
    c = O.voxel_grid(raw_scan(4, 700, with_bad=False), 0.1)
    got, want = O.statistical_outlier_removal(c, mean_k, mult), twin_statistical_outlier(c, mean_k, mult)
    assert np.array_equal(got, want)
    assert 0 < len(got) < len(c)
    few = c[:mean_k]                    # fewer than mean_k + 1 points: unchanged
    assert np.array_equal(O.statistical_outlier_removal(few, mean_k, mult), few)


def test_oracle_submap_matches_twin():
    from oracle import oracle as O
    from riv_slam_b200 import datagen
    clouds = [raw_scan(10 + i, 600, with_bad=False) for i in range(3)]
    poses = [datagen.pose_matrix([0.4 * i, 0.05 * i, 0.0], [0.0, 0.2 * i, 1.0 * i]) for i in range(3)]
    got = O.accumulate_submap(clouds, poses, 0.0)
    want = []
    for c, T in zip(clouds, poses):
        x = c[:, :3].astype(np.float64)
        q = ((T[:3, 0] * x[:, :1] + T[:3, 1] * x[:, 1:2]) + T[:3, 2] * x[:, 2:3]) + T[:3, 3]
        want.append(np.concatenate([q.astype(np.float32), c[:, 3:4]], axis=1))
    want = np.concatenate(want)
    assert got.tobytes() == want.tobytes()
    assert O.accumulate_submap(clouds, poses, 0.1).tobytes() == twin_voxel_grid(want, 0.1).tobytes()


# ---------------------------------------------------------------- GPU: the CUDA path vs the oracle, bit-exact

def _handle():
    from riv_slam_b200.fast_apdgicp import Handle
    return Handle(0)


CASES = [
    dict(),                                                         # code defaults + RADIUS (launch file)
    dict(outlier_removal=0),
    dict(downsample_resolution=0.0),                                # downsample_method NONE -> NaN removal only
    dict(use_distance_filter=0, outlier_removal=0),
    dict(use_distance_filter=0, downsample_resolution=0.0, outlier_removal=0),
    dict(distance_near_thresh=2.0, distance_far_thresh=60.0, z_low_thresh=-2.0, z_high_thresh=6.0, downsample_resolution=0.25, radius_radius=0.5,
         radius_min_neighbors=5),
    dict(outlier_removal=2),                                        # the nodelet's code defaults: STATISTICAL 20 / 1.0
    dict(outlier_removal=2, statistical_mean_k=30, statistical_stddev=1.2, downsample_resolution=0.2),   # launch-file values of the statistical parameters
    dict(downsample_method="APPROX_VOXELGRID"),                     # pcl::ApproximateVoxelGrid (preprocessing_nodelet.cpp:145-149) + RADIUS
    dict(downsample_method="APPROX_VOXELGRID", outlier_removal=0, downsample_resolution=0.5),
    dict(downsample_method="APPROX_VOXELGRID", use_distance_filter=0, outlier_removal=0),   # NaN / inf rows reach the filter
    dict(downsample_method="APPROX_VOXELGRID", outlier_removal=2, downsample_resolution=2.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("width", [4, 8])
def test_preprocess_bit_exact(case, width):
    from riv_slam_b200.fast_apdgicp import preprocess
    kw = dict(CASES[case])
    c = raw_scan(20 + case, 3000)
    want = oracle_pipeline(c, **kw)
    kw.setdefault("downsample_method", "VOXELGRID")
    cloud = c
    if width == 8:  # pcl::PointXYZI memory: x y z 1 | intensity pad pad pad
        cloud = np.zeros((len(c), 8), np.float32)
        cloud[:, :3], cloud[:, 3], cloud[:, 4] = c[:, :3], 1.0, c[:, 3]
    got = preprocess(_handle(), cloud, **kw)
    if width == 8:
        assert np.all(got[:, 3] == 1.0) and np.all(got[:, 5:] == 0.0)
        got = np.ascontiguousarray(got[:, [0, 1, 2, 4]])
    assert got.shape == want.shape
    if np.isnan(want).any():   # raw NaN / inf rows averaged by ApproximateVoxelGrid: NaN where the oracle has NaN (the payload bits of a NaN are not data)
        assert kw["downsample_method"] == "APPROX_VOXELGRID" and not kw.get("use_distance_filter", 1)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.where(np.isnan(got), np.float32(0), got).tobytes() == np.where(np.isnan(want), np.float32(0), want).tobytes()
    else:
        assert got.tobytes() == want.tobytes()
    assert 0 < len(got) < len(c)
    if kw["downsample_method"] == "APPROX_VOXELGRID" and kw.get("outlier_removal", 1) == 0:   # and it is not the exact filter in disguise
        assert got.tobytes() != oracle_pipeline(c, **dict(kw, downsample_method="VOXELGRID")).tobytes()


@pytest.mark.gpu
def test_preprocess_edge_cases():
    from riv_slam_b200.fast_apdgicp import preprocess, ApdError
    H = _handle()
    assert preprocess(H, np.zeros((0, 4), np.float32)).shape == (0, 4)
    assert preprocess(H, np.full((7, 4), np.nan, np.float32)).shape == (0, 4)          # everything filtered
    lone = np.array([[5, 0, 0, 1], [50, 0, 0, 2]], np.float32)
    assert preprocess(H, lone).shape == (0, 4)                                          # fewer points than min_neighbors + 1
    assert np.array_equal(preprocess(H, lone, outlier_removal=0), lone)
    with pytest.raises(ApdError):
        preprocess(H, lone, outlier_removal=3)                                          # unknown method: loud, no fallback
    with pytest.raises(ApdError):
        preprocess(H, lone, outlier_removal=2, statistical_mean_k=40)                   # more neighbours than the kernel keeps
    assert np.array_equal(preprocess(H, lone, outlier_removal=2), lone)                 # fewer than mean_k + 1 points: unchanged
    # a 20k-point raw scan (several CTA-sized segments per thread)
    big = np.concatenate([raw_scan(40 + i, 5000) for i in range(3)])
    assert preprocess(H, big).tobytes() == oracle_pipeline(big).tobytes()
    # downsample_method: APPROX_VOXELGRID on the same scan, empty / single-point input, and an unknown method refused loudly
    assert preprocess(H, big, downsample_method="APPROX_VOXELGRID").tobytes() == oracle_pipeline(big, downsample_method="APPROX_VOXELGRID").tobytes()
    assert preprocess(H, np.zeros((0, 4), np.float32)).shape == (0, 4)
    assert np.array_equal(preprocess(H, lone, outlier_removal=0), lone)
    with pytest.raises(ApdError):
        preprocess(H, lone, downsample_method="OCTREE")
    assert preprocess(H, big, downsample_method="VOXELGRID").tobytes() == oracle_pipeline(big).tobytes()   # sticky option switched back


@pytest.mark.gpu
def test_build_submap_approx_voxelgrid_bit_exact():
    """scan_matching_odometry_nodelet.cpp:156-160: the submap downsampled by pcl::ApproximateVoxelGrid (80 k accumulated points)."""
    from oracle import oracle as O
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import CloudSet, Handle, build_submap
    scans, poses = datagen.make_drive(9, 4, 17, n_points=5000)
    clouds = [np.ascontiguousarray(np.concatenate([s[:, :3], np.full((len(s), 1), 1.0 + i, np.float32)], axis=1)) for i, s in enumerate(scans)]
    which = list(range(16))
    rel = [np.linalg.inv(poses[i]) @ poses[16] for i in which]
    H = Handle(0)
    ks = CloudSet(H, clouds)
    for leaf in (0.1, 1.0):
        want = O.accumulate_submap([clouds[i] for i in which], rel, leaf, approx=True)
        got = build_submap(H, ks, which, rel, leaf, cache_key=5, downsample_method="APPROX_VOXELGRID")
        assert got.shape == want.shape and got.tobytes() == want.tobytes()
        assert len(want) > len(O.accumulate_submap([clouds[i] for i in which], rel, leaf))


@pytest.mark.gpu
@pytest.mark.parametrize("leaf", [0.1, 0.0])
def test_build_submap_bit_exact_and_becomes_target(leaf):
    from oracle import oracle as O
    from oracle.oracle import Oracle
    from riv_slam_b200 import datagen
    from riv_slam_b200.fast_apdgicp import CloudSet, FastAPDGICP, build_submap
    scans, poses = datagen.make_drive(9, 3, 6, n_points=3000)
    clouds = [np.ascontiguousarray(np.concatenate([s[:, :3], np.full((len(s), 1), 7.0 + i, np.float32)], axis=1)) for i, s in enumerate(scans)]
    which = [1, 2, 3, 4]
    last = poses[5]
    rel = [np.linalg.inv(poses[i]) @ last for i in which]  # scan_matching_odometry_nodelet.cpp:610 (as written in the reference)
    want = O.accumulate_submap([clouds[i] for i in which], rel, leaf)
    reg = FastAPDGICP(0)
    reg.handle().set_params(**LAUNCH_PARAMS)
    ks = CloudSet(reg.handle(), clouds)
    got = build_submap(reg.handle(), ks, which, rel, leaf, cache_key=77)
    assert got.shape == want.shape and got.tobytes() == want.tobytes()
    # the submap is now the target: aligning against it equals aligning against the same cloud set from the host
    reg.setInputSource(clouds[5])
    reg.align(want_output=False)
    assert reg.hasConverged()
    Ta = reg.getFinalTransformation().copy()
    fit_a, it_a = reg.getFitnessScore(), reg.nr_iterations()
    ref = FastAPDGICP(0)
    ref.handle().set_params(**LAUNCH_PARAMS)
    ref.setInputTarget(want)
    ref.setInputSource(clouds[5])
    ref.align(want_output=False)
    assert np.array_equal(Ta, ref.getFinalTransformation()) and fit_a == ref.getFitnessScore() and it_a == ref.nr_iterations()
    o = Oracle(**LAUNCH_PARAMS)
    o.set_source(clouds[5]); o.set_target(want)
    rc, To, conv, it = o.align()
    assert rc == 0 and conv and it == it_a
    assert np.abs(Ta.astype(np.float64) - To).max() < 1e-4
