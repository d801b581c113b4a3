#!/usr/bin/env python
"""Generates tests/golden/knn_nanoflann_v1.npz with the REFERENCE's own vendored exact kd-tree
(radar_graph_slam/include/scan_context/nanoflann.hpp, nanoflann 1.3.2, compiled from /root/reference into
oracle/_ref by `make -C oracle ref`; L2_Simple float metric as FLANN's, NANOFLANN_FIRST_MATCH).

This is the one golden vector of the path that comes from reference code run in this container: it pins the
nearest-neighbour search (k = 20 lists for the covariances, k = 1 correspondences) of the oracle and of the
CUDA path. Everything after the search (covariances, Mahalanobis, LM) stays pinned by the oracle only.

    make -C oracle ref && python tests/golden/make_knn_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from oracle import oracle as O
    from riv_slam_b200 import datagen
    assert O.ref_available(), "build oracle/_ref first: make -C oracle ref (needs /root/reference)"
    src, tgt, T_gt = datagen.make_pair(1, 77, n_src=1500, n_tgt=1600)
    moved = (src[:, :3].astype(np.float64) @ T_gt[:3, :3].T + T_gt[:3, 3]).astype(np.float32)  # roughly aligned queries for the 1-NN vector
    out = {"src": src, "tgt": tgt, "queries_1nn": moved, "seed": np.array([1, 77, 1500, 1600])}
    for k in (10, 20):
        idx, d2 = O.ref_nanoflann_knn(src, src, k)
        out[f"knn{k}_src_idx"], out[f"knn{k}_src_d2"] = idx.astype(np.int16), d2
    idx, d2 = O.ref_nanoflann_knn(tgt, tgt, 20)
    out["knn20_tgt_idx"], out["knn20_tgt_d2"] = idx.astype(np.int16), d2
    idx, d2 = O.ref_nanoflann_knn(tgt, moved, 1)
    out["nn1_idx"], out["nn1_d2"] = idx[:, 0].astype(np.int16), d2[:, 0]
    path = os.path.join(HERE, "knn_nanoflann_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
