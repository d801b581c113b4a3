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


# ---------------------------------------------------------------- LM branches (tests/lm_cases.py)

GOLD_LM = os.path.join(ROOT, "tests", "golden", "apd_golden_lm_v1.npz")


@pytest.fixture(scope="module")
def gold_lm():
    return np.load(GOLD_LM)


def _lm_case_names():
    import lm_cases
    return list(lm_cases.CASES)


@pytest.mark.parametrize("name", _lm_case_names())
def test_oracle_reproduces_lm_branch_golden(gold_lm, name):
    """The oracle keeps walking lsq_registration_impl.hpp:127-173 the same way: rejected trials with lambda * nu
    growth, rejected-but-converged, "lm not converged!!" and a non-default initial lambda factor."""
    import lm_cases
    src, tgt, _ = lm_cases.make_pair()
    r = lm_cases.run_oracle(name, src, tgt)
    tr, tr0 = r["trace"], gold_lm[f"{name}_trace"]
    assert tr.shape == tr0.shape and np.array_equal(tr[:, [0, 1, 7]], tr0[:, [0, 1, 7]])
    assert np.allclose(tr[:, [2, 3, 5, 6]], tr0[:, [2, 3, 5, 6]], rtol=1e-8)
    assert np.allclose(tr[:, 4], tr0[:, 4], rtol=1e-4, atol=1e-4)  # rho = (y0 - yi) / ...: cancellation near convergence, summation order varies (OpenMP)
    assert [int(r["converged"]), r["iterations"], int(r["lm_failed"])] == list(gold_lm[f"{name}_state"])
    assert np.abs(r["T"] - gold_lm[f"{name}_T"]).max() < 1e-6
    # the case really is what its name says
    rej = tr[:, 7] == 0
    if name.startswith("rejected_then_accepted"):
        assert rej.sum() >= 6 and tr[:, 1].max() >= 6 and tr[-1, 7] == 1 and r["converged"]
        # lambda doubles, then x4, x8 ... inside the rejected run (nu *= 2 per rejection, :160-163)
        run = tr[(tr[:, 0] == 1)]
        assert np.allclose(run[1:, 5] / run[:-1, 5], 2.0 ** np.arange(1, len(run)), rtol=1e-12)
    elif name == "rejected_but_converged":
        assert tr[-1, 7] == 0 and r["converged"] and not r["lm_failed"]
    elif name.startswith("lm_failed"):
        assert rej.all() and r["lm_failed"] and not r["converged"] and len(tr) == lm_cases.CASES[name][3]["lm_max_iterations"]
    else:
        assert not rej.any() and np.isclose(tr[0, 5], 1e-3 * np.abs(np.diag(r["final_hessian"])).max(), rtol=0.5)


@pytest.mark.parametrize("name", ["rejected_then_accepted", "rejected_but_converged", "lm_failed_2"])
def test_numpy_twin_walks_the_same_lm_branches(gold_lm, name):
    """oracle/pyref.py (LAPACK solve instead of LDL^T, rotation-vector exponential instead of the quaternion formula)
    takes the same accept / reject / give-up decisions on the injected-covariance cases."""
    import lm_cases
    from oracle.pyref import PyRef
    src, tgt, _ = lm_cases.make_pair()
    p = lm_cases.case_params(name)
    r = lm_cases.run_oracle(name, src, tgt)
    tw = PyRef(k=p["k_correspondences"], max_corr_dist=p["max_corr_dist"], max_iterations=p["max_iterations"], rotation_epsilon=p["rotation_epsilon"],
               transformation_epsilon=p["transformation_epsilon"], lm_max_iterations=p["lm_max_iterations"], dist_var=p["dist_var"],
               azimuth_var=p["azimuth_var"], elevation_var=p["elevation_var"])
    tw.src = np.asarray(src, dtype=np.float32)[:, :3]
    tw.tgt = np.asarray(tgt, dtype=np.float32)[:, :3]
    tw.cov_src, tw.cov_tgt = r["cov_src"], r["cov_tgt"]
    T, conv, it = tw.align()
    tr0 = gold_lm[f"{name}_trace"]
    assert tw.trace.shape == tr0.shape and np.array_equal(tw.trace[:, [0, 1, 7]], tr0[:, [0, 1, 7]])
    assert np.allclose(tw.trace[:, [2, 3, 5, 6]], tr0[:, [2, 3, 5, 6]], rtol=1e-6)
    assert np.allclose(tw.trace[:, 4], tr0[:, 4], rtol=1e-3, atol=1e-3)
    assert [int(conv), it] == list(gold_lm[f"{name}_state"][:2])
    assert np.abs(T - gold_lm[f"{name}_T"]).max() < 1e-5
