#!/usr/bin/env python
"""Generates tests/golden/apd_golden_v1.npz from the CPU oracle (oracle/liboracle.so).

These vectors pin the ORACLE - they guard it against drift and give the GPU tests a committed target that does
not depend on rebuilding it. The reference holds no golden vector for FastAPDGICP; the vectors made
by the reference's own compiled sources are tests/golden/apd_ref_golden_v1.npz (make_ref_golden.py), and
tests/test_reference_apdgicp.py holds the oracle to those sources directly.
Inputs come from the deterministic generator (riv_slam_b200.datagen, seed 20260000+1000*config+index).

    python tests/golden/make_golden.py        # rewrites the .npz next to this script
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

LAUNCH = dict(k_correspondences=20, max_corr_dist=2.0, max_iterations=64, transformation_epsilon=0.1,
              rotation_epsilon=2e-3, dist_var=0.86, azimuth_var=1.0, elevation_var=1.0)
TIGHT = dict(LAUNCH, transformation_epsilon=1e-6, rotation_epsilon=1e-6)
CONFIG, INDEX, N_SRC, N_TGT = 1, 42, 600, 640


def main():
    from oracle.oracle import Oracle
    from riv_slam_b200 import datagen
    src, tgt, T_gt = datagen.make_pair(CONFIG, INDEX, n_src=N_SRC, n_tgt=N_TGT)
    out = {"src": src, "tgt": tgt, "T_gt": T_gt, "seed": np.array([CONFIG, INDEX, N_SRC, N_TGT])}
    o = Oracle(**LAUNCH)
    o.set_source(src)
    o.set_target(tgt)
    assert o.compute_covariances() == 0
    out["knn_src"] = o.knn(0).astype(np.int16)
    out["knn_tgt"] = o.knn(1).astype(np.int16)
    out["cov_src"] = o.covariances(0)
    out["cov_tgt"] = o.covariances(1)
    for name, pose in (("I", np.eye(4)), ("gt", T_gt)):
        e, H, b = o.linearize(pose)
        corr, sq = o.correspondences()
        out[f"lin_{name}_pose"] = np.asarray(pose, dtype=np.float32)
        out[f"lin_{name}_err"] = np.array(e)
        out[f"lin_{name}_H"] = H
        out[f"lin_{name}_b"] = b
        out[f"lin_{name}_corr"] = corr.astype(np.int16)
        out[f"lin_{name}_sq"] = sq
    for name, prm in (("launch", LAUNCH), ("tight", TIGHT)):
        o = Oracle(**prm)
        o.set_source(src)
        o.set_target(tgt)
        rc, T, conv, it = o.align()
        assert rc == 0
        out[f"align_{name}_T"] = T
        out[f"align_{name}_conv_it"] = np.array([int(conv), it])
        out[f"align_{name}_trace"] = o.trace()
        out[f"align_{name}_fitness"] = np.array([o.fitness(), o.fitness(1.5)])
        out[f"align_{name}_final_hessian"] = o.final_hessian()
    np.savez_compressed(os.path.join(HERE, "apd_golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "apd_golden_v1.npz"))


if __name__ == "__main__":
    main()
