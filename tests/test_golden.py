"""Committed golden vectors (tests/golden/apd_golden_v1.npz, written by tests/golden/make_golden.py
from the CPU oracle): the oracle must keep reproducing them (CPU), and the CUDA path must match them
(GPU) — kNN sets bit-exact, covariances/H/b within 1e-5 relative, transforms within 1e-5 rad / 1e-4 m."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "apd_golden_v1.npz")
LAUNCH = dict(k_correspondences=20, max_corr_dist=2.0, max_iterations=64, transformation_epsilon=0.1,
              rotation_epsilon=2e-3, dist_var=0.86, azimuth_var=1.0, elevation_var=1.0)
TIGHT = dict(LAUNCH, transformation_epsilon=1e-6, rotation_epsilon=1e-6)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_generator_is_deterministic(gold):
    from riv_slam_b200 import datagen
    c, i, ns, nt = (int(v) for v in gold["seed"])
    src, tgt, T_gt = datagen.make_pair(c, i, n_src=ns, n_tgt=nt)
    assert np.array_equal(src, gold["src"]) and np.array_equal(tgt, gold["tgt"]) and np.array_equal(T_gt, gold["T_gt"])


def test_oracle_reproduces_golden(gold):
    from oracle.oracle import Oracle
    o = Oracle(**LAUNCH)
    o.set_source(gold["src"]); o.set_target(gold["tgt"])
    assert o.compute_covariances() == 0
    assert np.array_equal(o.knn(0), gold["knn_src"]) and np.array_equal(o.knn(1), gold["knn_tgt"])
    assert np.allclose(o.covariances(0), gold["cov_src"], rtol=0, atol=1e-13)
    e, H, b = o.linearize(gold["lin_gt_pose"])
    assert np.array_equal(o.correspondences()[0], gold["lin_gt_corr"])
    assert np.allclose(H, gold["lin_gt_H"], rtol=1e-9) and np.allclose(b, gold["lin_gt_b"], rtol=1e-9, atol=1e-9 * np.abs(gold["lin_gt_b"]).max())
    for name, prm in (("launch", LAUNCH), ("tight", TIGHT)):
        o = Oracle(**prm)
        o.set_source(gold["src"]); o.set_target(gold["tgt"])
        rc, T, conv, it = o.align()
        assert [int(conv), it] == list(gold[f"align_{name}_conv_it"])
        assert np.abs(T - gold[f"align_{name}_T"]).max() < 1e-6
        assert np.allclose([o.fitness(), o.fitness(1.5)], gold[f"align_{name}_fitness"], rtol=1e-9)


@pytest.mark.gpu
def test_cuda_path_matches_golden(gold):
    from riv_slam_b200.fast_apdgicp import FastAPDGICP
    g = FastAPDGICP(0)
    g.handle().set_params(**LAUNCH)
    g.setInputSource(gold["src"]); g.setInputTarget(gold["tgt"])
    assert np.array_equal(g.getKnn(0), gold["knn_src"]) and np.array_equal(g.getKnn(1), gold["knn_tgt"])
    for C, ref in ((g.getSourceCovariances(), gold["cov_src"]), (g.getTargetCovariances(), gold["cov_tgt"])):
        assert (np.abs(C[:, :3, :3] - ref).max(axis=(1, 2)) / np.abs(ref).max(axis=(1, 2))).max() <= 1e-5
    for name in ("I", "gt"):
        e, H, b = g.evaluateCost(gold[f"lin_{name}_pose"])
        corr, sq = g.getCorrespondences()
        assert np.array_equal(corr, gold[f"lin_{name}_corr"])
        m = corr >= 0
        assert np.array_equal(sq[m], gold[f"lin_{name}_sq"][m])
        assert abs(e - gold[f"lin_{name}_err"]) <= 1e-5 * abs(gold[f"lin_{name}_err"])
        assert np.abs(H - gold[f"lin_{name}_H"]).max() <= 1e-5 * np.abs(gold[f"lin_{name}_H"]).max()
        assert np.abs(b - gold[f"lin_{name}_b"]).max() <= 1e-5 * np.abs(gold[f"lin_{name}_b"]).max()
    for name, prm in (("launch", LAUNCH), ("tight", TIGHT)):
        g = FastAPDGICP(0)
        g.handle().set_params(**prm)
        g.setInputSource(gold["src"]); g.setInputTarget(gold["tgt"])
        g.align(want_output=False)
        T, T0 = g.getFinalTransformation(), gold[f"align_{name}_T"]
        assert [int(g.hasConverged()), g.nr_iterations()] == list(gold[f"align_{name}_conv_it"])
        assert np.abs(T[:3, :3] - T0[:3, :3]).max() <= 1e-5 and np.abs(T[:3, 3] - T0[:3, 3]).max() <= 1e-4
        tr, tr0 = g.getLMTrace(), gold[f"align_{name}_trace"]
        assert tr.shape == tr0.shape and np.array_equal(tr[:, [0, 1, 7]], tr0[:, [0, 1, 7]])
        f = gold[f"align_{name}_fitness"]
        assert abs(g.getFitnessScore() - f[0]) <= 1e-5 * f[0] and abs(g.getFitnessScore(1.5) - f[1]) <= 1e-5 * f[1]
        Hf = gold[f"align_{name}_final_hessian"]
        assert np.abs(g.getFinalHessian() - Hf).max() <= 1e-5 * np.abs(Hf).max()
