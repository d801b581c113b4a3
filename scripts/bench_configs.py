#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that bench.py's headline line does not cover
(C3 scan-to-submap, C4 batched loop-closure verification, C5 large-cloud sweep) and for the
"next" row 8(f)-1 (fitness scoring), each with the CPU oracle timed beside it on a bounded sample.
One JSON line per measurement on stdout; run on the GPU box:

    python scripts/bench_configs.py [c1] [c3] [c4] [c5] [fit] [pre] [submap]  > gpurun_out/configs.jsonl
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import LAUNCH_PARAMS  # noqa: E402
from riv_slam_b200 import datagen  # noqa: E402
from riv_slam_b200 import fast_apdgicp as F  # noqa: E402

PEAK = 6545.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def emit(**kw):
    print(json.dumps(kw), flush=True)


def gpu_time(H, fn, reps=3):
    """best-of wall clock around a synchronised call (the calls themselves end with a stream sync)"""
    best = 1e30
    for _ in range(reps):
        H.synchronize()
        t0 = time.perf_counter()
        fn()
        H.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def oracle(**kw):
    from oracle.oracle import Oracle
    p = dict(LAUNCH_PARAMS)
    p.update(kw)
    return Oracle(**p)


def c1():
    """Single align of two ~3k-point clouds: the reference-path config (CPU/OpenMP), with the GPU beside it."""
    src, tgt, T_gt = datagen.make_pair(1, 0, n_src=3000)
    from oracle.oracle import max_threads
    rows = {}
    for name, prm in (("launch", LAUNCH_PARAMS), ("tight", dict(LAUNCH_PARAMS, transformation_epsilon=1e-6, rotation_epsilon=1e-6))):
        o = oracle(**{k: v for k, v in prm.items() if k not in LAUNCH_PARAMS or prm[k] != LAUNCH_PARAMS[k]})
        cpu = []
        for _ in range(7):
            sec, Tc, conv, it, fit = o.timed_registration(src, tgt)
            cpu.append(sec)
        o1 = oracle(num_threads=1, **{k: v for k, v in prm.items() if k not in LAUNCH_PARAMS or prm[k] != LAUNCH_PARAMS[k]})
        sec1 = min(o1.timed_registration(src, tgt)[0] for _ in range(3))
        reg = F.FastAPDGICP(0)
        reg.handle().set_params(**prm)
        gpu = []
        for i in range(12):
            t0 = time.perf_counter()
            reg.setInputTarget(tgt, cache_key=2 * i + 1)
            reg.setInputSource(src, cache_key=2 * i + 2)
            reg.align(None, want_output=False)
            reg.getFitnessScore()
            gpu.append(time.perf_counter() - t0)
        Tg = reg.getFinalTransformation()
        rows[name] = dict(cpu_ms_all_cores=float(np.median(cpu[2:]) * 1e3), cpu_ms_1_thread=sec1 * 1e3, cpu_cores=max_threads(), gpu_ms=float(np.median(gpu[3:]) * 1e3),
                          iterations=it, converged=bool(conv), gpu_iterations=reg.nr_iterations(), T_abs_err_vs_oracle=float(np.abs(Tc - Tg).max()),
                          fitness=fit, gpu_fitness=reg.getFitnessScore())
    emit(config="C1 single align, N = M = 3000, cold (both clouds new: upload + grid + kNN + covariances + align + fitness)", **{f"{k}_{kk}": vv for k, r in rows.items() for kk, vv in r.items()})


def c3():
    """5k-point scans against a ~100k-point accumulated submap, target prepared once and reused."""
    n_scans, per = 24, 5000
    scans, poses = datagen.make_drive(3, 0, n_scans + 20, per, workers=min(16, os.cpu_count() or 1))
    sub = []
    for t in range(20):  # 20 scans accumulated in the frame of scan 0 (SMO:606-617 keeps the last keyframes)
        Trel = np.linalg.inv(poses[0]) @ poses[t]
        p = scans[t].copy()
        p[:, :3] = (scans[t][:, :3].astype(np.float64) @ Trel[:3, :3].T + Trel[:3, 3]).astype(np.float32)
        sub.append(p)
    submap = np.concatenate(sub)
    queries = scans[20:20 + n_scans]
    guesses = np.stack([(np.linalg.inv(poses[0]) @ poses[19]).astype(np.float32)] * n_scans)
    H = F.Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    W = F.CloudSet(H, [submap])   # warm-up: first use of these kernels at this size (module load, pool growth)
    W.prepare(); H.synchronize(); W.destroy()
    T = F.CloudSet(H, [submap])
    t_prep = gpu_time(H, lambda: (setattr(T, "_x", 0), T.prepare())[1], reps=1)
    S = F.CloudSet(H, queries)
    S.prepare()
    H.synchronize()
    res = None

    def run():
        nonlocal res
        res = F.align_pairs(H, S, T, tgt_idx=np.zeros(n_scans, np.int32), guesses=guesses)
    t_batch = gpu_time(H, run)
    # one scan at a time through the reference-shaped object (target cached by key)
    reg = F.FastAPDGICP(0)
    reg.handle().set_params(**LAUNCH_PARAMS)
    reg.setInputTarget(submap, cache_key=1)
    lat = []
    for i, q in enumerate(queries):
        t0 = time.perf_counter()
        reg.setInputTarget(submap, cache_key=1)
        reg.setInputSource(q, cache_key=100 + i)
        reg.align(guesses[i], want_output=False)
        reg.getFitnessScore()
        lat.append(time.perf_counter() - t0)
    o = oracle()
    t0 = time.perf_counter()
    o.set_target(submap)
    o.set_source(queries[0])
    o.compute_covariances()
    t_cpu_prep = time.perf_counter() - t0
    cpu = []
    for i in range(min(8, n_scans)):
        t0 = time.perf_counter()
        o.set_source(queries[i])
        rc, Tc, conv, it = o.align(guesses[i])
        o.fitness()
        cpu.append(time.perf_counter() - t0)
        assert np.abs(Tc - res[i]["T"]).max() < 1e-3
    emit(config="C3 scan-to-submap", n_src=per, n_tgt=int(submap.shape[0]), target_prepare_ms=t_prep * 1e3, batch_reg_per_s=n_scans / t_batch,
         single_p50_latency_ms=float(np.median(lat[2:]) * 1e3), cpu_target_prepare_ms=t_cpu_prep * 1e3, cpu_p50_latency_ms=float(np.median(cpu) * 1e3),
         cpu_reg_per_s=len(cpu) / sum(cpu), converged=int(res["converged"].sum()), mean_iterations=float(res["iterations"].mean()))


def c4():
    """4096 independent 5k-point pairs, identity guess (loop-closure candidate verification)."""
    n_pairs, uniq = 4096, 64
    base = [datagen.make_pair(4, i, n_src=5000) for i in range(uniq)]
    srcs = [base[i % uniq][0] for i in range(n_pairs)]
    tgts = [base[i % uniq][1] for i in range(n_pairs)]
    H = F.Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    ps, os_ = F._ragged(srcs)
    pt, ot = F._ragged(tgts)
    res = None

    def e2e():
        nonlocal res
        res = F.batch_align(H, (ps, os_), (pt, ot))
    e2e()
    t_e2e = gpu_time(H, e2e, reps=2)
    lin, err, _ = H.work_counters()
    # device-resident path: sets uploaded once, then grid + kNN + covariances ("prepare") and one align launch.
    # prepare is idempotent, so it is timed on fresh sets after one warm-up round (first-use allocations).
    S, T = F.CloudSet(H, (ps, os_)), F.CloudSet(H, (pt, ot))
    S.prepare(); T.prepare(); H.synchronize()
    S.destroy(); T.destroy()
    S, T = F.CloudSet(H, (ps, os_)), F.CloudSet(H, (pt, ot))
    t_prep = gpu_time(H, lambda: (S.prepare(), T.prepare()), reps=1)
    t_align = gpu_time(H, lambda: F.align_pairs(H, S, T), reps=2)
    o = oracle()
    cpu = []
    for i in range(40):
        sec, Tc, conv, it, fit = o.timed_registration(srcs[i], tgts[i])
        cpu.append(sec)
        assert np.abs(Tc - res[i]["T"]).max() < 1e-3 and bool(res[i]["converged"]) == conv
    bytes_align = 5000 * (148 * lin + 84 * err + 32 * n_pairs)
    emit(config="C4 batched loop-closure verification", n_pairs=n_pairs, unique_pairs=uniq, e2e_reg_per_s=n_pairs / t_e2e, e2e_ms=t_e2e * 1e3,
         prepare_ms=t_prep * 1e3, align_ms=t_align * 1e3, device_reg_per_s=n_pairs / (t_prep + t_align), align_algorithmic_GBps=bytes_align / t_align / 1e9,
         align_frac_of_measured_hbm=bytes_align / t_align / 1e9 / PEAK, cpu_reg_per_s=len(cpu) / sum(cpu), cpu_sample="first 40 pairs, all host cores",
         converged_frac=float(res["converged"].mean()), mean_iterations=float(res["iterations"].mean()))


def c5():
    """200k-point source vs 1M-point target, k in {10, 15, 20}. Both are accumulated keyframe maps of one
    drive (330 resp. 80 scans of 5000 points moved into one frame, 0.1 m voxel de-duplication as in the
    reference pipeline, launch:56-57), independently sampled, 0.8 m / 1 deg apart."""
    w = min(16, os.cpu_count() or 1)
    tgt, pose_t = datagen.make_map(5, 0, 330, 5000, frame=150, workers=w, total_scans=340, speed=1.0)
    src, pose_s = datagen.make_map(5, 0, 80, 5000, frame=151, first=112, workers=w, total_scans=340, speed=1.0, resample=1)
    rng = np.random.default_rng(5)
    tgt = tgt[np.sort(rng.choice(tgt.shape[0], min(1000000, tgt.shape[0]), replace=False))]
    src = src[np.sort(rng.choice(src.shape[0], min(200000, src.shape[0]), replace=False))]
    T_gt = np.linalg.inv(pose_t) @ pose_s
    for k in (10, 15, 20):
        reg = F.FastAPDGICP(0)
        reg.handle().set_params(**dict(LAUNCH_PARAMS, k_correspondences=k))
        H = reg.handle()
        # warm-up on the same handle: first use of this k's kernels and the pool's growth are not part of the figure
        reg.setInputTarget(tgt, cache_key=1); reg.setInputSource(src, cache_key=2)
        reg.computeCovariances(); H.synchronize()
        reg.setInputTarget(tgt, cache_key=3)   # new keys: the clouds are uploaded, gridded and searched again
        reg.setInputSource(src, cache_key=4)
        H.synchronize()
        t_cov = gpu_time(H, reg.computeCovariances, reps=1)       # grid + kNN + covariances of both clouds (1.2M points)
        t_lin = gpu_time(H, lambda: reg.evaluateCost(np.eye(4)), reps=3)
        # the kernel alone (CUDA events on the launch stream), without the host side of the call
        H.set_option("kernel_timing", 1); H.kernel_times()
        for _ in range(3):
            reg.evaluateCost(np.eye(4))
        k_ms, k_n = H.kernel_times()
        H.set_option("kernel_timing", 0)
        t_lin_kernel = k_ms["align"] / max(k_n["align"], 1) * 1e-3
        t_align = gpu_time(H, lambda: reg.align(None, want_output=False), reps=2)
        lin, err, _ = H.work_counters()
        n_all = src.shape[0] + tgt.shape[0]
        row = dict(config="C5 large-cloud sweep", k=k, n_src=int(src.shape[0]), n_tgt=int(tgt.shape[0]), grid_knn_cov_ms=t_cov * 1e3,
                   knn_cov_points_per_s=n_all / t_cov, knn_cov_algorithmic_GBps=n_all * 84 / t_cov / 1e9, linearize_ms=t_lin * 1e3, linearize_kernel_ms=t_lin_kernel * 1e3,
                   linearize_algorithmic_GBps=src.shape[0] * 148 / t_lin / 1e9, linearize_frac_of_measured_hbm=src.shape[0] * 148 / t_lin / 1e9 / PEAK,
                   align_ms=t_align * 1e3, linearize_passes=int(lin), error_passes=int(err), converged=bool(reg.hasConverged()), iterations=reg.nr_iterations())
        if k == 20:
            o = oracle(k_correspondences=k)
            o.set_target(tgt)
            o.set_source(src)
            t0 = time.perf_counter()
            o.compute_covariances()
            row["cpu_knn_cov_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            e0, H0, b0 = o.linearize(np.eye(4))
            row["cpu_linearize_ms"] = (time.perf_counter() - t0) * 1e3
            e, Hg, bg = reg.evaluateCost(np.eye(4))
            row["H_rel_err_vs_oracle"] = float(np.abs(Hg - H0).max() / np.abs(H0).max())
            t0 = time.perf_counter()
            rc, Tc, conv, it = o.align()
            row["cpu_align_ms"] = (time.perf_counter() - t0) * 1e3
            reg.align(None, want_output=False)
            row["T_abs_err_vs_oracle"] = float(np.abs(Tc - reg.getFinalTransformation()).max())
        emit(**row)


def fit():
    """8(f)-1: fitness scoring of a sliding window of keyframe pairs in one launch vs one by one on the CPU."""
    n, per = 256, 5000
    scans, poses = datagen.make_drive(2, 7, 33, per, workers=min(16, os.cpu_count() or 1))
    idx_a = np.arange(n) % 32
    idx_b = idx_a + 1
    rel = np.stack([(np.linalg.inv(poses[a]) @ poses[b]).astype(np.float32) for a, b in zip(idx_a, idx_b)])
    H = F.Handle(0)
    S = F.CloudSet(H, scans)
    sc = None

    def run():
        nonlocal sc
        sc = F.fitness_pairs(H, S, S, src_idx=idx_b, tgt_idx=idx_a, poses=rel, max_range=1e300)
    run()
    t = gpu_time(H, run)
    o = oracle()
    cpu = []
    for i in range(16):
        t0 = time.perf_counter()
        o.set_target(scans[idx_a[i]])
        o.set_source(scans[idx_b[i]])
        f0 = o.fitness_score(rel[i])
        cpu.append(time.perf_counter() - t0)
        assert abs(f0 - sc[i]) <= 1e-5 * f0
    emit(config="8(f)-1 fitness scoring (calc_fitness_score)", n_pairs=n, points=per, gpu_scores_per_s=n / t, gpu_ms=t * 1e3,
         algorithmic_GBps=n * per * 32 / t / 1e9, cpu_scores_per_s=len(cpu) / sum(cpu), cpu_note="kd-tree build + single-threaded queries, as the reference")


def pre():
    """8(f)-4: distance filter -> 0.1 m voxel grid -> radius outlier removal of raw scans (preprocessing_nodelet.cpp:812-815)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_preprocess import raw_scan, oracle_pipeline
    H = F.Handle(0)
    for n in (5000, 20000):
        clouds = [np.concatenate([raw_scan(100 + 7 * i + j, 5000) for j in range(n // 5000)]) for i in range(6)]
        outs = [F.preprocess(H, c) for c in clouds]  # warm-up
        ts = []
        for _ in range(5):
            for c in clouds:
                t0 = time.perf_counter()
                F.preprocess(H, c)
                ts.append(time.perf_counter() - t0)
        cpu = []
        for c, g in zip(clouds[:3], outs[:3]):
            t0 = time.perf_counter()
            want = oracle_pipeline(c)
            cpu.append(time.perf_counter() - t0)
            assert want.tobytes() == g.tobytes()
        emit(config="8(f)-4 preprocessing filters (distance, VoxelGrid 0.1 m, RadiusOutlierRemoval 0.8 m / 2)", points_in=int(len(clouds[0])), points_out=int(len(outs[0])),
             gpu_ms_p50=float(np.median(ts) * 1e3), cpu_oracle_ms=float(np.median(cpu) * 1e3), bit_exact=True,
             note="host PointXYZI-like array in, host array out, per call; the CPU oracle's outlier filter is a brute-force O(n^2) restatement, not PCL's kd-tree")


def submap():
    """8(f)-2: accumulate the last keyframes into the scan-to-map target (scan_matching_odometry_nodelet.cpp:606-616) on the device."""
    from oracle import oracle as O
    per = 5000
    for n_key in (5, 20):
        scans, poses = datagen.make_drive(3, 1, n_key + 1, per, workers=min(16, os.cpu_count() or 1))
        clouds = [np.ascontiguousarray(np.concatenate([s[:, :3], np.full((len(s), 1), 1.0, np.float32)], axis=1)) for s in scans]
        which = list(range(n_key))
        rel = [np.linalg.inv(poses[i]) @ poses[n_key] for i in which]
        reg = F.FastAPDGICP(0)
        reg.handle().set_params(**LAUNCH_PARAMS)
        H = reg.handle()
        ks = F.CloudSet(H, clouds)
        got = F.build_submap(H, ks, which, rel, 0.1, cache_key=1)
        t_build = gpu_time(H, lambda: F.build_submap(H, ks, which, rel, 0.1, cache_key=0, want_cloud=False), reps=5)
        t0 = time.perf_counter()
        want = O.accumulate_submap([clouds[i] for i in which], rel, 0.1)
        t_cpu = time.perf_counter() - t0
        assert want.tobytes() == got.tobytes()

        def host_path():  # what the reference does: build the submap on the host, hand it to setInputTarget
            reg.setInputTarget(want, cache_key=0)
            reg.computeCovariances()
        def dev_path():
            F.build_submap(H, ks, which, rel, 0.1, cache_key=0, want_cloud=False)
            reg.computeCovariances()
        reg.setInputSource(clouds[n_key])
        host_path(); dev_path()
        t_host = gpu_time(H, host_path, reps=5)
        t_dev = gpu_time(H, dev_path, reps=5)
        emit(config="8(f)-2 submap accumulation (transform + concatenate + VoxelGrid 0.1 m -> target)", keyframes=n_key, points_in=n_key * per, points_out=int(len(got)),
             gpu_build_ms=t_build * 1e3, cpu_oracle_build_ms=t_cpu * 1e3, bit_exact=True,
             target_ready_ms_device_resident=t_dev * 1e3, target_ready_ms_host_cloud_upload=t_host * 1e3,
             reference_flow_ms=(t_cpu + t_host) * 1e3,
             note="target_ready = submap + grid + kNN + covariances, i.e. until align can start; host_cloud_upload takes a submap that already exists on "
                  "the host, reference_flow adds the CPU oracle's time to build it there (what the nodelet does before setInputTarget)")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c4", "fit", "pre", "submap", "c5"]
    for w in which:
        {"c1": c1, "c3": c3, "c4": c4, "c5": c5, "fit": fit, "pre": pre, "submap": submap}[w]()
