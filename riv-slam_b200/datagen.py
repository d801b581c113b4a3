"""Deterministic synthetic 4D-radar clouds for the FastAPDGICP hot path.

The reference ships no data for this path (its only test never builds FastAPDGICP and its
``data/`` directory is absent), so every input in this repository comes from here.

Scene model (static world, metres, sensor x forward / y left / z up):
  * ground plane z = -sensor_height (reference: radar_graph_slam/launch/radar_graph_slam.launch:190)
  * two side walls y = +8 and y = -6, optional far wall
  * random axis-aligned boxes and vertical poles, plus volumetric clutter
Sensor model (Oculii-Eagle-like): azimuth +-56.5 deg
(radar_graph_slam/include/scan_context/Scancontext.h:110), elevation +-20 deg, range 2..100 m
(launch:51-52). First-hit ray casting of uniformly drawn (azimuth, elevation), polar noise
(sigma_r = 0.86/400 * r, the s_x model of fast_apdgicp_impl.hpp:169; 0.1 deg angular), 0.03 m
Cartesian jitter, keep-one 0.1 m voxel filter (launch:56-57), on-axis reject |y|+|z| < 1e-3.

Everything is numpy float64 internally and float32 (n, 4) = x, y, z, intensity on output.
"""
from __future__ import annotations

import dataclasses
import numpy as np

SEED_BASE = 20260000

AZ_HALF_FOV = np.deg2rad(56.5)
EL_HALF_FOV = np.deg2rad(20.0)
RANGE_MIN = 2.0
RANGE_MAX = 100.0


def pair_seed(config: int, pair_index: int) -> int:
    """seed = 20260000 + 1000*config + pair_index."""
    return SEED_BASE + 1000 * int(config) + int(pair_index)


@dataclasses.dataclass
class Scene:
    ground_z: float
    wall_y_pos: float
    wall_y_neg: float
    far_wall_x: float  # np.inf -> none
    box_lo: np.ndarray  # (B, 3)
    box_hi: np.ndarray  # (B, 3)
    pole_xy: np.ndarray  # (P, 2)
    pole_r: np.ndarray  # (P,)
    pole_h: np.ndarray  # (P,) top z
    scale: float = 1.0


def make_scene(rng: np.random.Generator, x_min=-20.0, x_max=120.0, far_wall=True, scale=1.0,
               boxes_per_100m=40, poles_per_100m=60, clear_radius=3.0, clear_lane=None) -> Scene:
    """Random static world. Obstacles that would swallow the sensor are dropped: everything within
    ``clear_radius`` (scaled) of the origin, and, for drives, everything intersecting the lane
    ``clear_lane = (y_lo, y_hi)`` the vehicle moves in."""
    length = (x_max - x_min)
    nb = max(1, int(round(boxes_per_100m * length / 100.0 / scale)))
    npole = max(1, int(round(poles_per_100m * length / 100.0 / scale)))
    wy_p, wy_n = 8.0 * scale, -6.0 * scale
    c = np.stack([rng.uniform(x_min, x_max, nb), rng.uniform(wy_n + 0.5, wy_p - 0.5, nb),
                  np.full(nb, -2.0)], axis=1)
    size = np.stack([rng.uniform(0.5, 3.0, nb), rng.uniform(0.5, 2.0, nb), rng.uniform(0.5, 3.5, nb)], axis=1) * scale
    lo = c - np.array([0.5, 0.5, 0.0]) * size
    hi = c + np.array([0.5, 0.5, 1.0]) * size
    pxy = np.stack([rng.uniform(x_min, x_max, npole), rng.uniform(wy_n + 0.3, wy_p - 0.3, npole)], axis=1)
    pr = rng.uniform(0.08, 0.25, npole) * scale
    ph = -2.0 + rng.uniform(2.0, 6.0, npole) * scale
    cr = clear_radius * scale
    keep_b = ~((lo[:, 0] < cr) & (hi[:, 0] > -cr) & (lo[:, 1] < cr) & (hi[:, 1] > -cr))
    keep_p = np.hypot(pxy[:, 0], pxy[:, 1]) > cr + pr
    if clear_lane is not None:
        keep_b &= ~((lo[:, 1] < clear_lane[1]) & (hi[:, 1] > clear_lane[0]))
        keep_p &= ~((pxy[:, 1] - pr < clear_lane[1]) & (pxy[:, 1] + pr > clear_lane[0]))
    lo, hi, pxy, pr, ph = lo[keep_b], hi[keep_b], pxy[keep_p], pr[keep_p], ph[keep_p]
    return Scene(ground_z=-2.0, wall_y_pos=wy_p, wall_y_neg=wy_n,
                 far_wall_x=(x_min + 80.0 * scale if far_wall else np.inf),
                 box_lo=lo, box_hi=hi, pole_xy=pxy, pole_r=pr, pole_h=ph, scale=scale)


def pose_matrix(t, rpy_deg) -> np.ndarray:
    """4x4 world_from_sensor with R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = np.deg2rad(np.asarray(rpy_deg, dtype=np.float64))
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T


def _cast(scene: Scene, origin: np.ndarray, dirs: np.ndarray, rmax: float) -> np.ndarray:
    """First-hit range along unit rays (n,3) from origin (3,); inf where nothing is hit."""
    if dirs.shape[0] > 65536:  # bound the (rays x obstacles) temporaries of the million-point configs
        return np.concatenate([_cast(scene, origin, dirs[i:i + 65536], rmax) for i in range(0, dirs.shape[0], 65536)])
    n = dirs.shape[0]
    best = np.full(n, np.inf)
    eps = 1e-12

    def plane(axis, value):
        nonlocal best
        d = dirs[:, axis]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (value - origin[axis]) / np.where(np.abs(d) < eps, np.nan, d)
        ok = np.isfinite(t) & (t > 1e-6)
        best = np.where(ok & (t < best), t, best)

    plane(2, scene.ground_z)
    plane(1, scene.wall_y_pos)
    plane(1, scene.wall_y_neg)
    if np.isfinite(scene.far_wall_x):
        plane(0, scene.far_wall_x)

    # axis-aligned boxes, slab method (n, B)
    if scene.box_lo.shape[0]:
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / np.where(np.abs(dirs) < eps, eps, dirs)  # (n,3)
        t0 = (scene.box_lo[None, :, :] - origin[None, None, :]) * inv[:, None, :]
        t1 = (scene.box_hi[None, :, :] - origin[None, None, :]) * inv[:, None, :]
        tn = np.minimum(t0, t1).max(axis=2)
        tf = np.maximum(t0, t1).min(axis=2)
        hit = (tn <= tf) & (tn > 1e-6)
        tb = np.where(hit, tn, np.inf).min(axis=1)
        best = np.minimum(best, tb)

    # vertical finite cylinders (n, P)
    if scene.pole_xy.shape[0]:
        ox = origin[0] - scene.pole_xy[:, 0][None, :]
        oy = origin[1] - scene.pole_xy[:, 1][None, :]
        dx = dirs[:, 0][:, None]
        dy = dirs[:, 1][:, None]
        a = dx * dx + dy * dy
        b = 2.0 * (ox * dx + oy * dy)
        c = ox * ox + oy * oy - (scene.pole_r ** 2)[None, :]
        disc = b * b - 4 * a * c
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * np.where(a < eps, np.nan, a))
        z = origin[2] + t * dirs[:, 2][:, None]
        ok = np.isfinite(t) & (t > 1e-6) & (z >= scene.ground_z) & (z <= scene.pole_h[None, :])
        tp = np.where(ok, t, np.inf).min(axis=1)
        best = np.minimum(best, tp)

    return np.where(best <= rmax, best, np.inf)


def scan(scene: Scene, pose: np.ndarray, n_points: int, rng: np.random.Generator,
         voxel: float | None = 0.1, clutter_frac: float = 0.15) -> np.ndarray:
    """One radar scan of exactly ``n_points`` points in the SENSOR frame, float32 (n,4)."""
    s = scene.scale
    rmin, rmax = RANGE_MIN * s, RANGE_MAX * s
    R = pose[:3, :3]
    origin = pose[:3, 3]
    out = np.empty((0, 4), dtype=np.float64)
    seen = np.empty(0, dtype=np.int64)
    attempts = 0
    while out.shape[0] < n_points:
        attempts += 1
        if attempts > 64:
            raise RuntimeError("scene too empty to draw the requested number of points")
        m = int((n_points - out.shape[0]) * 1.8) + 64
        az = rng.uniform(-AZ_HALF_FOV, AZ_HALF_FOV, m)
        el = rng.uniform(-EL_HALF_FOV, EL_HALF_FOV, m)
        d_s = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        rng_hit = _cast(scene, origin, d_s @ R.T, rmax)
        # volumetric clutter replaces a fraction of the returns (ghost targets / multipath)
        is_clutter = rng.uniform(size=m) < clutter_frac
        r_clutter = rng.uniform(rmin, 0.6 * rmax, m)
        r = np.where(is_clutter, np.minimum(r_clutter, np.where(np.isfinite(rng_hit), rng_hit, np.inf)), rng_hit)
        ok = np.isfinite(r) & (r >= rmin) & (r <= rmax)
        # polar measurement noise
        r_n = r + rng.normal(0.0, 1.0, m) * (0.86 / 400.0) * np.where(ok, r, 0.0)
        az_n = az + rng.normal(0.0, np.deg2rad(0.1), m)
        el_n = el + rng.normal(0.0, np.deg2rad(0.1), m)
        p = np.stack([r_n * np.cos(el_n) * np.cos(az_n), r_n * np.cos(el_n) * np.sin(az_n), r_n * np.sin(el_n)], axis=1)
        p = p + rng.normal(0.0, 0.03 * s, (m, 3))
        inten = rng.uniform(0.0, 40.0, m)
        ok &= (np.abs(p[:, 1]) + np.abs(p[:, 2])) >= 1e-3
        p = p[ok]
        inten = inten[ok]
        if voxel is not None:
            kv = np.floor(p / (voxel * s)).astype(np.int64) + (1 << 20)
            keys = (kv[:, 0] << 42) | (kv[:, 1] << 21) | kv[:, 2]
            _, first = np.unique(keys, return_index=True)  # first occurrence per voxel
            keep = np.zeros(p.shape[0], dtype=bool)
            keep[first] = True
            keep &= ~np.isin(keys, seen)
            seen = np.concatenate([seen, keys[keep]])
            p = p[keep]
            inten = inten[keep]
        out = np.concatenate([out, np.concatenate([p, inten[:, None]], axis=1)], axis=0)
    return np.ascontiguousarray(out[:n_points].astype(np.float32))


def random_relative_pose(rng: np.random.Generator, max_trans=1.0, max_rot_deg=3.0) -> np.ndarray:
    """T_gt: translation U(0, max_trans) mostly along x, rotation <= max_rot_deg (launch:83-84)."""
    tx = rng.uniform(0.0, max_trans)
    ty = rng.uniform(-0.2, 0.2) * max_trans
    tz = rng.uniform(-0.05, 0.05) * max_trans
    rpy = rng.uniform(-1.0, 1.0, 3) * np.array([0.3, 0.3, 1.0]) * max_rot_deg
    return pose_matrix([tx, ty, tz], rpy)


def make_pair(config: int, pair_index: int, n_src: int = 5000, n_tgt: int | None = None,
              scale: float = 1.0, voxel: float | None = 0.1):
    """Independent (source, target, T_gt) with p_target = T_gt * p_source.

    Target is the scene seen from the identity pose, source the same scene seen from T_gt with
    independent ray sampling and noise (SURVEY.md §8d "Pair").
    """
    n_tgt = n_src if n_tgt is None else n_tgt
    rng = np.random.Generator(np.random.PCG64(pair_seed(config, pair_index)))
    scene = make_scene(rng, x_min=-20.0 * scale, x_max=120.0 * scale, far_wall=True, scale=scale)
    T_gt = random_relative_pose(rng, max_trans=1.0 * scale)
    tgt = scan(scene, np.eye(4), n_tgt, rng, voxel=voxel)
    src = scan(scene, T_gt, n_src, rng, voxel=voxel)
    return src, tgt, T_gt


def make_sequence(config: int, seq_index: int, n_scans: int, n_points: int = 5000,
                  speed: float = 0.5, voxel: float | None = 0.1):
    """A smooth drive through a long corridor: ``n_scans`` scans and their world poses.

    Pair t of the odometry workload is (source = scan t+1, target = scan t); its ground truth is
    inv(pose_t) @ pose_{t+1}. Step length ~ ``speed`` m with slowly varying yaw (<= 3 deg/frame).
    """
    rng = np.random.Generator(np.random.PCG64(pair_seed(config, seq_index)))
    length = speed * n_scans * 1.15 + 140.0
    scene = make_scene(rng, x_min=-20.0, x_max=length, far_wall=False, clear_lane=(-1.0, 3.0))
    poses = []
    x, y, z, yaw = 0.0, 1.0, 0.0, 0.0
    yaw_rate = 0.0
    for t in range(n_scans):
        poses.append(pose_matrix([x, y, z], [0.2 * np.sin(0.05 * t), 0.2 * np.cos(0.07 * t), yaw]))
        step = speed * (1.0 + 0.2 * np.sin(0.03 * t)) + rng.normal(0.0, 0.02)
        yaw_rate = 0.9 * yaw_rate + rng.normal(0.0, 0.15)
        yaw_rate = float(np.clip(yaw_rate, -2.0, 2.0))
        # steer back to the corridor centre line so the walls stay in view
        yaw_rate -= 0.05 * yaw + 0.2 * (y - 1.0)
        yaw += float(np.clip(yaw_rate, -3.0, 3.0))
        x += step * np.cos(np.deg2rad(yaw))
        y += step * np.sin(np.deg2rad(yaw))
    scans = [scan(scene, P, n_points, rng, voxel=voxel) for P in poses]
    return scans, poses


def relative_gt(poses, t: int) -> np.ndarray:
    """Ground-truth transform mapping scan t+1 coordinates into scan t coordinates."""
    return np.linalg.inv(poses[t]) @ poses[t + 1]


# ---------------------------------------------------------------------------------------------
# Benchmark-scale sequences: every scan has its own seed, so scans can be generated in parallel
# ---------------------------------------------------------------------------------------------

def _drive_poses(n_scans: int, speed: float, rng: np.random.Generator):
    poses = []
    x, y, yaw, yaw_rate = 0.0, 1.0, 0.0, 0.0
    for t in range(n_scans):
        poses.append(pose_matrix([x, y, 0.0], [0.2 * np.sin(0.05 * t), 0.2 * np.cos(0.07 * t), yaw]))
        step = speed * (1.0 + 0.2 * np.sin(0.03 * t)) + rng.normal(0.0, 0.02)
        yaw_rate = float(np.clip(0.9 * yaw_rate + rng.normal(0.0, 0.15), -2.0, 2.0))
        yaw_rate -= 0.05 * yaw + 0.2 * (y - 1.0)
        yaw += float(np.clip(yaw_rate, -3.0, 3.0))
        x += step * np.cos(np.deg2rad(yaw))
        y += step * np.sin(np.deg2rad(yaw))
    return poses


def _drive_scan(args):
    config, seq_index, t, n_scans, n_points, speed, voxel, resample = args
    rng0 = np.random.Generator(np.random.PCG64(pair_seed(config, seq_index)))
    length = speed * n_scans * 1.15 + 140.0
    scene = make_scene(rng0, x_min=-20.0, x_max=length, far_wall=False, clear_lane=(-1.0, 3.0))
    poses = _drive_poses(n_scans, speed, rng0)
    rng = np.random.Generator(np.random.PCG64([pair_seed(config, seq_index), t + 1 + 1000003 * resample]))
    return scan(scene, poses[t], n_points, rng, voxel=voxel)


def make_drive(config: int, seq_index: int, n_scans: int, n_points: int = 5000, speed: float = 0.5,
               voxel: float | None = 0.1, workers: int = 0, resample: int = 0):
    """Like :func:`make_sequence` but scan t draws from its own stream PCG64([seed, t+1]), so the scans
    can be generated by a process pool. Returns (scans, poses)."""
    rng0 = np.random.Generator(np.random.PCG64(pair_seed(config, seq_index)))
    length = speed * n_scans * 1.15 + 140.0
    make_scene(rng0, x_min=-20.0, x_max=length, far_wall=False, clear_lane=(-1.0, 3.0))  # advance the stream exactly as the workers do
    poses = _drive_poses(n_scans, speed, rng0)
    # resample > 0 re-draws every scan (same scene, same poses, independent rays and noise)
    jobs = [(config, seq_index, t, n_scans, n_points, speed, voxel, resample) for t in range(n_scans)]
    if workers and workers > 1 and n_scans > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, n_scans)) as pool:
            scans = pool.map(_drive_scan, jobs, chunksize=max(1, n_scans // (4 * workers)))
    else:
        scans = [_drive_scan(j) for j in jobs]
    return scans, poses


def voxel_keep_one(points: np.ndarray, voxel: float) -> np.ndarray:
    """Keep the first point of every voxel (the generator's stand-in for the 0.1 m down-sampling of launch:56-57)."""
    kv = np.floor(points[:, :3].astype(np.float64) / voxel).astype(np.int64) + (1 << 20)
    keys = (kv[:, 0] << 42) | (kv[:, 1] << 21) | kv[:, 2]
    _, first = np.unique(keys, return_index=True)
    return points[np.sort(first)]


def make_map(config: int, seq_index: int, n_scans: int, n_points: int = 5000, frame: int = 0, first: int = 0, stride: int = 1,
             voxel: float = 0.1, workers: int = 0, speed: float = 0.5, total_scans: int | None = None, resample: int = 0):
    """An accumulated, voxel-filtered map: scans first, first+stride, ... of a drive moved into the frame of
    pose `frame` and de-duplicated per voxel — the kind of cloud a keyframe map holds (bounded density,
    unlike a raw million-point scan). Returns (cloud float32 (m,4), pose of `frame`)."""
    total = total_scans or (first + stride * n_scans + 1)
    scans, poses = make_drive(config, seq_index, total, n_points, speed=speed, workers=workers, resample=resample)
    inv = np.linalg.inv(poses[frame])
    parts = []
    for j in range(n_scans):
        t = first + stride * j
        Trel = inv @ poses[t]
        p = scans[t].copy()
        p[:, :3] = (scans[t][:, :3].astype(np.float64) @ Trel[:3, :3].T + Trel[:3, 3]).astype(np.float32)
        parts.append(p)
    return voxel_keep_one(np.concatenate(parts), voxel), poses[frame]
