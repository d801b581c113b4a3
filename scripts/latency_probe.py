#!/usr/bin/env python
"""Wall-clock breakdown of the single-pair (odometry) call sequence through the C ABI."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from riv_slam_b200 import datagen
from riv_slam_b200.fast_apdgicp import FastAPDGICP
from bench import LAUNCH_PARAMS, to_pointxyzi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
opts = [a.split("=") for a in sys.argv[2:]]
scans, _ = datagen.make_drive(2, 0, n + 1, 5000, workers=8)
pts, off = to_pointxyzi(scans, np.arange(n + 1))
clouds = [np.ascontiguousarray(pts[i * 5000:(i + 1) * 5000]) for i in range(n + 1)]
reg = FastAPDGICP(0)
reg.handle().set_params(**LAUNCH_PARAMS)
for k, v in opts:
    reg.setOption(k, float(v))
H = reg.handle()
rows = []
for i in range(n):
    t = [time.perf_counter()]
    reg.setInputTarget(clouds[i], cache_key=i + 1); H.synchronize(); t.append(time.perf_counter())
    reg.setInputSource(clouds[i + 1], cache_key=i + 2); H.synchronize(); t.append(time.perf_counter())
    reg.computeCovariances(); t.append(time.perf_counter())
    reg.align(None, want_output=False); t.append(time.perf_counter())
    reg.getFitnessScore(); t.append(time.perf_counter())
    rows.append(np.diff(t))
r = np.array(rows[3:]) * 1e3
names = ["setInputTarget(cached)", "setInputSource(upload)", "grid+knn+cov", "align(+fitness)", "getFitnessScore"]
for nm, m, p in zip(names, np.median(r, 0), np.percentile(r, 95, 0)):
    print(f"{nm:26s} p50 {m:7.3f} ms   p95 {p:7.3f} ms")
print(f"{'total':26s} p50 {np.median(r.sum(1)):7.3f} ms   iterations {reg.nr_iterations()} launches/pair {H.launch_count() / n:.1f}")
# the caller's flow (no synchronisation between the calls) with the library's CUDA-event timers on: GPU time per kernel and pair
reg.setOption("kernel_timing", 1)
H.kernel_times()
wall = []
for i in range(n):
    t0 = time.perf_counter()
    reg.setInputTarget(clouds[i], cache_key=i + 1)
    reg.setInputSource(clouds[i + 1], cache_key=i + 2)
    reg.align(None, want_output=False)
    reg.getFitnessScore()
    wall.append(time.perf_counter() - t0)
k_ms, k_n = H.kernel_times()
reg.setOption("kernel_timing", 0)
print("caller's flow: wall p50 %.3f ms; GPU time per pair (CUDA events): %s" % (np.median(wall) * 1e3, "  ".join(f"{k} {1e3 * v / max(k_n[k], 1):.1f} us x{k_n[k] / n:.1f}" for k, v in k_ms.items())))
# in-kernel timeline of one more align (phase stamps from %globaltimer)
reg.setOption("timeline", 1)
reg.setInputTarget(clouds[0], cache_key=1001); reg.setInputSource(clouds[1], cache_key=1002)
reg.computeCovariances()
reg.align(None, want_output=False)
tl = H.timeline()
names = {0: "enter", 1: "staged", 2: "iter", 3: "corr", 4: "H/b", 5: "LM trial", 6: "fitness", 10: "nn1>", 11: "<nn1", 12: "mahal", 13: "passend", 14: "synced", 20: "acc", 21: "solved", 22: "erracc", 23: "unpacked", 24: "ldlt", 25: "so3exp", 30: "wsum", 31: "part", 32: "barrier"}
if tl:
    t0 = tl[0][1]
    print("timeline (us):", " ".join(f"{names.get(p, p)}@{(t - t0) / 1e3:.1f}" for p, t in tl if p < 100))
    kl = [(p, t) for p, t in tl if p >= 120]
    if kl:
        kn = {120: "enter", 121: "staged", 122: "own leaf", 123: "searched", 124: "exact", 125: "cov", 126: "regularised", 127: "done", 128: "last group done", 129: "longest group (as if started at enter)"}
        print("leaf kNN, first group (us):", " ".join(f"{kn.get(p, p)}@{(t - kl[0][1]) / 1e3:.1f}" for p, t in kl))
    bl = [(p, t) for p, t in tl if 100 <= p < 120]
    if bl:
        bn = {100: "enter", 101: "bbox", 102: "keys", 103: "pass0", 104: "pass1", 105: "pass2", 106: "pass3", 107: "gather", 108: "boxes", 109: "p0 hist", 110: "p0 ranked"}
        print("leaf build (us):", " ".join(f"{bn.get(p, p)}@{(t - bl[0][1]) / 1e3:.1f}" for p, t in sorted(bl, key=lambda x: x[1])))
dc = H.debug_counters()
for nm, c in (("first pass", dc[:8]), ("seeded passes", dc[8:])):
    if c[0]:
        print(f"{nm}: groups {c[0]} queries/group {c[6] / c[0]:.1f} leaves popped/group {c[1] / c[0]:.1f} broadcast scans/group {c[2] / c[0]:.2f} turns/group {c[3] / c[0]:.1f} max turns {c[4]} max leaves {c[5]}")
