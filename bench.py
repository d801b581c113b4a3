#!/usr/bin/env python
"""Benchmark of the FastAPDGICP scan-matching hot path (BASELINE.json metric:
"scan-pair registrations/sec @5k-pt 4D radar scans; p50 align latency (ms)").

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the CPU path (oracle port) on the host cores

A *step* is one pass of the hot path over one batch of synthetic input: config C2 of BASELINE.json —
scan-to-scan odometry over `--pairs` (default 1000) consecutive 5000-point radar scan pairs
(scan t+1 -> scan t). Every scan is gridded, kNN-searched and given covariances once (it is the source
of one pair and the target of the next, the reference's own usage pattern,
radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:449-468,584-592), then all pairs are
aligned and scored (getFitnessScore). A registration = set target + set source + align + fitness.

Reported on one JSON line:
  value   registrations/s with the raw scans already resident in HBM (CUDA events on the launch stream)
  e2e     the same through the C-ABI with PINNED HOST buffers in pcl::PointXYZI layout (32 B/point):
          host->device copy of every scan and device->host read of every result inside the timed region
  roofline  the align kernel: algorithmic bytes (SURVEY.md §8d / DESIGN.md) / measured launch time
  cpu_baseline  the CPU oracle (a port of the reference; the reference itself cannot be built here)
                on a bounded sample of the same workload, all host cores
With N > 1 every rank runs its own segment of the drive (weak scaling, no data-path collective)
and the 96-byte result records are all-gathered over NCCL inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# radar_graph_slam/launch/radar_graph_slam.launch:34-36,95-101
LAUNCH_PARAMS = dict(k_correspondences=20, max_corr_dist=2.0, max_iterations=64, transformation_epsilon=0.1,
                     rotation_epsilon=2e-3, dist_var=0.86, azimuth_var=1.0, elevation_var=1.0)
METRIC = "scan-pair registrations/sec @5k-pt 4D radar scans"
UNIT = "registrations/s"
N_POINTS = 5000
# algorithmic bytes per source point
B_LINEARIZE, B_ERROR, B_FITNESS, B_PREPARE = 148, 84, 32, 84


def make_workload(n_pairs: int, unique: int, seq_index: int, workers: int):
    """scans[0..n_pairs] of a drive; `unique` distinct scans generated, then replayed forth and back
    (every consecutive pair stays a physically adjacent scan pair)."""
    from riv_slam_b200 import datagen
    u = max(2, min(unique, n_pairs + 1))
    scans, _ = datagen.make_drive(2, seq_index, u, N_POINTS, workers=workers)
    order = []
    t, d = 0, 1
    for _ in range(n_pairs + 1):
        order.append(t)
        if t + d < 0 or t + d >= u:
            d = -d
        t += d
    return scans, np.asarray(order)


def to_pointxyzi(scans, order):
    """(sum n, 8) float32 in pcl::PointXYZI memory layout: x y z 1 | intensity 0 0 0, plus offsets."""
    n = len(order)
    out = np.zeros((n * N_POINTS, 8), dtype=np.float32)
    for i, s in enumerate(order):
        blk = out[i * N_POINTS:(i + 1) * N_POINTS]
        blk[:, :3] = scans[s][:, :3]
        blk[:, 3] = 1.0
        blk[:, 4] = scans[s][:, 3]
    off = (np.arange(n + 1) * N_POINTS).astype(np.int32)
    return out, off


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    """Threads the CPU arm uses: every core this process may run on. torchrun exports OMP_NUM_THREADS=1, which must
    not shrink the reference arm to one thread, so the count is taken from the affinity mask, not from OpenMP."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_registrations(scans, order, n_sample, threads, time_budget_s):
    """The CPU path on pairs 0..n_sample-1 of the same drive: kd-tree build + kNN/covariances once per
    scan, align + fitness per pair. Returns (registrations/s, pairs done, cores, per-pair seconds)."""
    from oracle.oracle import Oracle
    cores = threads or host_cores()
    o = Oracle(num_threads=cores, **LAUNCH_PARAMS)
    per = []
    t_start = time.perf_counter()
    prev = None
    done = 0
    for i in range(n_sample):
        tgt = scans[order[i]]
        src = scans[order[i + 1]]
        t0 = time.perf_counter()
        if prev is None:
            o.set_target(tgt)      # first frame: setInputTarget (SMO:437)
        else:
            o.swap()               # the previous source becomes the target with its covariances (cache-by-pointer equivalent)
        o.set_source(src)
        rc, T, conv, it = o.align()
        o.fitness()
        per.append(time.perf_counter() - t0)
        prev = src
        done += 1
        if time.perf_counter() - t_start > time_budget_s:
            break
    total = sum(per)
    return done / total, done, cores, per


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, all host threads) on this arm's workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = min(args.pairs, args.cpu_sample)
    scans, order = make_workload(args.pairs, args.unique, 0, args.workers)
    cores = host_cores()
    rates = []
    for s in range(args.warmup + args.steps):
        rate, done, cores, per = cpu_registrations(scans, order, n_sample, cores, args.cpu_budget)
        if s >= args.warmup:
            rates.append((rate, done, per))
    total_pairs = sum(d for _, d, _ in rates)
    total_time = sum(sum(p) for _, _, p in rates)
    value = total_pairs / total_time
    allper = np.concatenate([np.asarray(p) for _, _, p in rates])
    sample = f"first {rates[0][1]} of {args.pairs} pairs of the C2 drive per step, covariances reused scan to scan"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_time / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "p50_align_latency_ms": float(np.median(allper) * 1e3),
        "config": {"workload": f"C2 sequential scan-to-scan odometry, {args.pairs} pairs x {N_POINTS} pts (bounded sample)", "params": "launch file"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1000, help="scan pairs per step per GPU")
    ap.add_argument("--unique", type=int, default=96, help="distinct scans generated (replayed forth and back)")
    ap.add_argument("--workers", type=int, default=0, help="processes for data generation (0 = host cores)")
    ap.add_argument("--cpu-sample", type=int, default=200)
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--latency-pairs", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="name=value tuning option for apd_set_option")
    args = ap.parse_args()
    if args.workers <= 0:
        args.workers = min(32, os.cpu_count() or 1)
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    P = args.pairs
    # synthetic scans first: the generator forks worker processes, which must happen before this process owns a CUDA context
    scans, order = make_workload(P, args.unique, rank, max(1, args.workers // max(1, min(world, 8))))   # each rank drives its own segment
    host_np, off = to_pointxyzi(scans, order)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from riv_slam_b200 import fast_apdgicp as F
    from riv_slam_b200 import build
    build.build_library()  # no-op when the prebuilt .so is current; raises if it cannot be built
    H = F.Handle(local_rank)
    H.set_params(**LAUNCH_PARAMS)
    for kv in args.option:
        k, v = kv.split("=")
        H.set_option(k, float(v))
    stream = torch.cuda.Stream(device=dev)
    H.set_stream(stream.cuda_stream)
    L = H.L
    import ctypes as C

    host = torch.from_numpy(host_np).pin_memory()
    dev_pts = host.to(dev, non_blocking=False)
    res_dev = torch.zeros(P * 96, dtype=torch.uint8, device=dev)
    res_host = torch.zeros(P * 96, dtype=torch.uint8).pin_memory()
    gathered = torch.zeros(world * P * 96, dtype=torch.uint8, device=dev) if world > 1 else None
    src_idx = np.arange(1, P + 1, dtype=np.int32)
    tgt_idx = np.arange(0, P, dtype=np.int32)
    ip = C.POINTER(C.c_int32)

    def step(pts_tensor, mem, out_host):
        """One pass of the hot path. Returns (cloudset handle to destroy, event pairs)."""
        cs = C.c_void_p()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            H.check(L.apd_cloudset_create(H.h, C.c_void_p(pts_tensor.data_ptr()), 32, off.ctypes.data_as(ip), P + 1, mem, C.byref(cs)))
            H.check(L.apd_cloudset_prepare(H.h, cs))
            ev[1].record(stream)
            H.check(L.apd_align_pairs(H.h, cs, cs, src_idx.ctypes.data_as(ip), tgt_idx.ctypes.data_as(ip), None, P,
                                      C.c_void_p(res_dev.data_ptr()), F.MEM_DEVICE))
            ev[2].record(stream)
            if world > 1:
                dist.all_gather_into_tensor(gathered, res_dev)
            if out_host:
                res_host.copy_(res_dev, non_blocking=True)
            ev[3].record(stream)
        return cs, ev

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(pts_tensor, mem, out_host, steps):
        barrier()
        t_wall0 = time.perf_counter()
        tot = prep = align = 0.0
        # steps run back to back; each step's cloud set is freed after its results are complete
        for _ in range(steps):
            cs, ev = step(pts_tensor, mem, out_host)
            ev[3].synchronize()
            tot += ev[0].elapsed_time(ev[3])
            prep += ev[0].elapsed_time(ev[1])
            align += ev[1].elapsed_time(ev[2])
            L.apd_cloudset_destroy(H.h, cs)
        barrier()
        wall = time.perf_counter() - t_wall0
        return tot * 1e-3, prep * 1e-3, align * 1e-3, wall

    # ---- warm-up (>= 3 steps requested by the contract), then the device-resident timed region ----
    timed(dev_pts, F.MEM_DEVICE, False, max(args.warmup, 1))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = H.launch_count()
    t0 = time.perf_counter()
    tot, prep, align, wall = timed(dev_pts, F.MEM_DEVICE, False, args.steps)
    t1 = time.perf_counter()
    launches = H.launch_count() - launches0
    lin, err, _ = H.work_counters()   # of the last step (all steps do identical work)

    # ---- end to end: pinned host PointXYZI buffers in, host results out, through the public call a user makes ----
    # apd_odometry_align uploads every scan from (pinned) host memory, builds grids + covariances, aligns
    # all pairs and writes the 96-byte records to host memory; it pipelines chunks over two streams.
    res_np = np.zeros(P, dtype=F.RESULT_DTYPE)
    res_pin = torch.from_numpy(res_np.view(np.uint8).reshape(-1)).pin_memory()
    res_view = np.frombuffer(memoryview(res_pin.numpy()), dtype=F.RESULT_DTYPE)

    def e2e_step():
        H.check(L.apd_odometry_align(H.h, C.c_void_p(host.data_ptr()), off.ctypes.data_as(ip), P + 1, 32, None, C.c_void_p(res_pin.data_ptr())))

    def e2e_timed(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                e2e_step()          # returns when the results are in host memory
                if world > 1:
                    dist.all_gather_into_tensor(gathered, res_dev)
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0

    e2e_timed(max(args.warmup, 1))
    e_tot, e_wall = e2e_timed(args.steps)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None   # samples cover both timed regions (device-resident and end to end)
    e_tot = max(e_tot, e_wall)    # the helper stream's work is not on `stream`: the host clock bounds the region
    results = res_view.copy()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tot_m, e_m, align_m, prep_m = allmax(tot), allmax(e_tot), allmax(align), allmax(prep)
    # e2e uses the larger of the device-event time and the host wall clock around the same region
    e_m = max(e_m, 0.0)
    e_wall_m = allmax(e_wall)

    # ---- single-pair latency through the reference-shaped object (setInputTarget/Source + align + fitness) ----
    lat = []
    reg = F.FastAPDGICP(local_rank)
    reg.handle().set_params(**LAUNCH_PARAMS)
    n_lat = min(args.latency_pairs, P)
    clouds = [np.ascontiguousarray(host_np[i * N_POINTS:(i + 1) * N_POINTS]) for i in range(n_lat + 1)]
    guess = np.eye(4, dtype=np.float32)
    for i in range(n_lat):
        ta = time.perf_counter()
        reg.setInputTarget(clouds[i], cache_key=i + 1)        # the previous source: device data reused
        reg.setInputSource(clouds[i + 1], cache_key=i + 2)
        reg.align(guess, want_output=False)
        reg.getFitnessScore()
        lat.append(time.perf_counter() - ta)
    p50 = float(np.median(lat[3:]) * 1e3) if len(lat) > 3 else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = world * P * args.steps / tot_m
    e2e_value = world * P * args.steps / max(e_m, 1e-12)
    # roofline of the dominant kernel (align): algorithmic bytes of one launch / its measured duration
    align_bytes = N_POINTS * (B_LINEARIZE * lin + B_ERROR * err + B_FITNESS * P)
    align_s = align_m / args.steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    else:
        peak, peak_src = 6650.0, "fallback"
    achieved = align_bytes / align_s / 1e9
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of one align launch at this shape, from the committed ncu capture
    try:
        if P == 1000:
            traffic = int(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))["align_kernel"]["dram_bytes_per_launch"])
    except Exception:
        traffic = None
    prep_bytes = (P + 1) * N_POINTS * B_PREPARE
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_m / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": f"synthetic ({len(scans)} generated 4D-radar scans of one drive, replayed forth and back to {P + 1} scans per GPU)",
        "config": {"workload": f"C2 sequential scan-to-scan odometry: {P} pairs x {N_POINTS} pts per GPU per step (scan t+1 -> scan t, identity guess)",
                   "params": "launch file (k=20, dmax=2.0, eps 0.1/2e-3, 64 iters, vars 0.86/1.0/1.0, PLANE)",
                   "l2": "inputs larger than L2 (per-step working set ~0.6 GB per GPU)", "parallelism": f"pair-sharded x{world}"},
        "p50_align_latency_ms": p50,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_np.nbytes), "d2h_bytes_per_step": int(P * 96),
                "ms_per_step": 1e3 * e_m / args.steps, "wall_ms_per_step": 1e3 * e_wall_m / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "align_kernel<TEAM_CTA,staged>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(align_bytes), "launch_ms": 1e3 * align_s,
                     "linearize_passes": int(lin), "error_passes": int(err)},
        "phases_ms_per_step": {"upload_or_copy+grid+knn_cov": 1e3 * prep_m / args.steps, "align+fitness": 1e3 * align_m / args.steps,
                               "prepare_algorithmic_GBps": prep_bytes / (prep_m / args.steps) / 1e9},
        "results": {"converged_frac": float(np.mean(results["converged"] != 0)), "mean_iterations": float(np.mean(results["iterations"])),
                    "mean_fitness": float(np.mean(results["fitness"])), "status_ok_frac": float(np.mean(results["status"] == 0))},
    }
    if not args.no_cpu:
        rate, done, cores, per = cpu_registrations(scans, order, min(P, args.cpu_sample), 0, args.cpu_budget)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {done} of {P} pairs of the same drive, covariances reused scan to scan",
                                "p50_latency_ms": float(np.median(per) * 1e3)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
