#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): single pair in every
team shape, a ragged batch, the pipelined host call and the fitness entry points."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from riv_slam_b200 import datagen
from riv_slam_b200 import fast_apdgicp as F
from bench import LAUNCH_PARAMS

src, tgt, _ = datagen.make_pair(1, 3, n_src=700, n_tgt=800)
for team in (0, 1, 2, 4):
    for unstaged in (0, 1):
        r = F.FastAPDGICP(0)
        r.handle().set_params(**LAUNCH_PARAMS)
        r.setOption("team_size", team)
        r.setOption("force_unstaged", unstaged)
        r.setInputTarget(tgt); r.setInputSource(src)
        r.align(want_output=True)
        r.getFitnessScore(1.0); r.getKnn(0); r.getCorrespondences(); r.evaluateCost(np.eye(4))
        print("team", team, "unstaged", unstaged, r.hasConverged(), r.nr_iterations())
# leaf kNN with its queries cut into 1 / 4 / 8 warp parts (transposed scans write into other lanes' pending lists), and the
# profiling stamps of the three kernels
for parts in (1, 4, 8):
    r = F.FastAPDGICP(0)
    r.handle().set_params(**LAUNCH_PARAMS)
    r.setOption("knn_leaf_parts", parts)
    r.setOption("timeline", 1 if parts == 4 else 0)
    r.setInputTarget(tgt, cache_key=100 + parts); r.setInputSource(src, cache_key=200 + parts)
    r.computeCovariances()
    r.align(want_output=False)
    print("parts", parts, r.hasConverged(), len(r.handle().timeline()))
H = F.Handle(0)
H.set_params(**LAUNCH_PARAMS)
clouds = [src[:300], src, tgt[:500], src[:0], tgt]
res = F.batch_align(H, clouds, [tgt, tgt[:600], src, tgt, src[:5]])
print(res["status"], res["converged"])
print(F.odometry_align(H, [src, tgt, src[:400], tgt[:350]])["iterations"])
S = F.CloudSet(H, clouds[:3])
print(F.fitness_pairs(H, S, S, src_idx=[0, 1], tgt_idx=[2, 2], max_range=4.0))
# pre-processing filters and submap accumulation (single-CTA kernels with shared-memory histograms / warp votes)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_preprocess import raw_scan
raw = raw_scan(5, 1500)
print(len(F.preprocess(H, raw)), len(F.preprocess(H, raw, outlier_removal=0, downsample_resolution=0.25)), len(F.preprocess(H, raw, use_distance_filter=0, downsample_resolution=0.0)))
xyzi = [np.ascontiguousarray(np.concatenate([c[:, :3], np.ones((len(c), 1), np.float32)], axis=1)) for c in (src, tgt, src[:400])]
K = F.CloudSet(H, xyzi)
print(len(F.build_submap(H, K, [0, 1, 2], [np.eye(4)] * 3, 0.1)), len(F.build_submap(H, K, [2, 0], [np.eye(4)] * 2, 0.0)))
# fallback global-memory build (the shared-memory one is the default for these sizes)
H2 = F.Handle(0)
H2.set_params(**LAUNCH_PARAMS)
H2.set_option("smem_build", 0)
print(F.batch_align(H2, [src, tgt[:500]], [tgt, src])["iterations"])
print("done")

# pcl::ApproximateVoxelGrid path (single-CTA radix sort + run scan) and the register staging path (bulk_stage off)
raw = np.concatenate([np.asarray(src[:, :3], np.float32), np.random.default_rng(0).uniform(0, 30, (len(src), 1)).astype(np.float32)], axis=1)
Hh = F.Handle(0)
print("approx voxel grid", len(F.preprocess(Hh, raw, downsample_method="APPROX_VOXELGRID", downsample_resolution=0.5)))
F.set_downsample_method(Hh, "VOXELGRID")
r = F.FastAPDGICP(0)
r.handle().set_params(**LAUNCH_PARAMS)
r.setOption("bulk_stage", 0)
r.setInputTarget(tgt, cache_key=901); r.setInputSource(src, cache_key=902)
r.align(want_output=False); r.getFitnessScore(1.0)
print("register staging", r.hasConverged(), r.nr_iterations())

# the upload-ahead host pipeline: page-locked input (copy stream + events, kernel-fetched table blocks and guesses, index bases)
import torch
drive, _ = datagen.make_drive(2, 1, 12, 400, workers=2)
many = [drive[i % 12] for i in range(331)]            # 330 pairs: three chunks
pts, off = F._ragged([np.ascontiguousarray(s[:, :4]) for s in many])
pin = torch.from_numpy(pts.copy()).pin_memory()
Hp = F.Handle(0); Hp.set_params(**LAUNCH_PARAMS)
gs = np.tile(np.eye(4, dtype=np.float32), (330, 1, 1))
ra = F.odometry_align(Hp, (pin, off), guesses=gs).copy()
Hp.set_option("upload_ahead", 0)
rb = F.odometry_align(Hp, (pin, off), guesses=gs).copy()
print("upload ahead", ra.tobytes() == rb.tobytes(), int((ra["converged"] != 0).sum()))
