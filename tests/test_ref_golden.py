"""The oracle against tests/golden/apd_ref_golden_v1.npz - vectors produced by the REFERENCE'S OWN FastAPDGICP sources, compiled
unmodified over stand-in Eigen / PCL / Boost headers (oracle/ref_apdgicp.cpp, tests/golden/make_ref_golden.py). Runs anywhere:
the vectors stand in for /root/reference. The same vectors are the GPU path's target in tests/test_gpu_ref_golden.py."""
import os

import numpy as np
import pytest

from conftest import ROOT

import ref_cases as R

TIGHT = 1e-9


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "apd_ref_golden_v1.npz"))


@pytest.fixture(scope="module")
def pair():
    return R.make_pair()


def _oracle(params, pair):
    from oracle.oracle import Oracle
    o = Oracle(**params)
    o.set_source(pair[0]); o.set_target(pair[1])
    return o


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


def test_vectors_come_from_the_reference_sources(gold):
    assert "compiled from /root/reference/fast_apdgicp/include" in str(gold["version"])
    assert len(gold.files) >= 180


@pytest.mark.parametrize("name", list(R.COV_CASES))
def test_covariances(gold, pair, name):
    o = _oracle(R.COV_CASES[name], pair)
    assert o.compute_covariances() == 0
    for which, side in ((0, "src"), (1, "tgt")):
        key = f"cov_{name}_{side}"
        if key not in gold.files:
            continue
        C0, C1 = o.covariances(which), gold[key]
        assert (np.abs(C0 - C1).max(axis=(1, 2)) <= TIGHT * np.abs(C1).max(axis=(1, 2))).all()


@pytest.mark.parametrize("name", list(R.LIN_CASES))
def test_linearize(gold, pair, name):
    for i, P in enumerate(R.poses()):
        o = _oracle(R.LIN_CASES[name], pair)
        e, H, b = o.linearize_d(P)
        k = f"lin_{name}_{i}"
        corr, sq = o.correspondences()
        assert np.array_equal(corr, gold[k + "_corr"]) and np.array_equal(sq, gold[k + "_sq"])
        Q = np.array(P); Q[:3, 3] += [0.01, 0.02, -0.01]
        assert np.allclose([e, o.compute_error_d(Q)], gold[k + "_e"], rtol=TIGHT, atol=0)
        assert _rel(H, gold[k + "_H"]) <= TIGHT and _rel(b, gold[k + "_b"]) <= TIGHT
        if k + "_mahal" in gold.files:
            m = corr >= 0
            M0, M1 = o.mahalanobis()[m], gold[k + "_mahal"][m]
            assert (np.abs(M0 - M1).max(axis=(1, 2)) <= TIGHT * np.abs(M1).max(axis=(1, 2))).all()


def test_libc_atan2f_vector_is_a_ulp_away(gold):
    """The stored linearization made with the C library's atan2f differs from the convention's by what 1 ulp of a float angle explains."""
    rel = _rel(gold["lin_launch_1_libc_H"], gold["lin_launch_1_H"])
    assert 0 < rel <= 1e-5


def check_align(got, gold, key, t_tol=1e-7, h_tol=1e-8):
    """got: dict(T, converged, iterations, lm_failed, trace (n, 8) or None, final_hessian)."""
    assert [int(got["converged"]), got["iterations"], int(got["lm_failed"])] == list(gold[key + "_state"]), key
    assert np.abs(np.asarray(got["T"], np.float64) - gold[key + "_T"]).max() <= t_tol, key
    tr1 = gold[key + "_trace"]
    tr0 = got["trace"]
    if tr0 is not None and tr1.size:
        assert tr0.shape == tr1.shape, key
        assert np.array_equal(tr0[:, [0, 1, 7]], tr1[:, [0, 1, 7]]), key                    # outer, inner, accepted
        assert np.allclose(tr0[:, [2, 3]], tr1[:, [2, 3]], rtol=got.get("y_rtol", 1e-8), atol=0), key      # y0, yi
        assert np.allclose(tr0[:, 5], tr1[:, 5], rtol=got.get("lambda_rtol", 1e-8), atol=0), key            # lambda
        assert np.allclose(tr0[:, 6], tr1[:, 6], rtol=1e-5, atol=1e-11), key                                # |delta|
        big = np.abs(tr1[:, 2] - tr1[:, 3]) > 1e-6 * np.abs(tr1[:, 2])
        assert np.allclose(tr0[big, 4], tr1[big, 4], rtol=1e-4, atol=1e-4), key                             # rho
    assert _rel(got["final_hessian"], gold[key + "_final_hessian"]) <= h_tol, key


def _run(o, g=None):
    rc, T, conv, it = o.align(g)
    assert rc == 0
    return dict(T=T, converged=conv, iterations=it, lm_failed=o.lm_failed(), trace=o.trace(), final_hessian=o.final_hessian())


@pytest.mark.parametrize("name", list(R.ALIGN_CASES))
def test_align(gold, pair, name):
    o = _oracle(R.ALIGN_CASES[name], pair)
    check_align(_run(o), gold, f"align_{name}")
    T = gold[f"align_{name}_T"]
    assert np.array_equal(o.transform_source(T)[:64], gold[f"align_{name}_aligned_head"])
    o = _oracle(R.ALIGN_CASES[name], pair)
    check_align(_run(o, R.guess()), gold, f"align_{name}_guess")


def test_lm_branches(gold, pair):
    import lm_cases
    base = _oracle(lm_cases.LAUNCH, pair)
    assert base.compute_covariances() == 0
    cov_src = base.covariances(0)
    for name in lm_cases.CASES:
        cov_tgt = lm_cases.injected_target_covariances(name, base.covariances(1))
        o = _oracle(lm_cases.case_params(name), pair)
        o.set_covariances(0, cov_src); o.set_covariances(1, cov_tgt)
        check_align(_run(o), gold, f"lm_{name}")
    rej = sum(int((gold[f"lm_{n}_trace"][:, 7] == 0).sum()) for n in lm_cases.CASES)
    assert rej >= 9 and list(gold["lm_lm_failed_1_state"]) == [0, 0, 1] and gold["lm_rejected_but_converged_trace"][-1, 7] == 0


def test_baseline_size_pair_and_chain(gold):
    """BASELINE.json's size: a 5000-point pair (covariances, H / b at two poses, two registrations) and an odometry chain of
    5000-point scans with swapSourceAndTarget between the pairs, as the reference's own sources computed them."""
    from oracle.oracle import Oracle
    s5, t5, _ = R.make_pair5k()
    o = Oracle(**R.LIN_CASES["launch"])
    o.set_source(s5); o.set_target(t5)
    assert o.compute_covariances() == 0
    for which, side in ((0, "src"), (1, "tgt")):
        C0, C1 = o.covariances(which), gold[f"p5k_cov_{side}"]
        assert (np.abs(C0 - C1).max(axis=(1, 2)) <= TIGHT * np.abs(C1).max(axis=(1, 2))).all()
    for i, P in enumerate(R.poses()[:2]):
        e, H, b = o.linearize_d(P)
        corr, sq = o.correspondences()
        assert np.array_equal(corr, gold[f"p5k_lin_{i}_corr"]) and np.array_equal(sq, gold[f"p5k_lin_{i}_sq"])
        assert abs(e - float(gold[f"p5k_lin_{i}_e"])) <= TIGHT * abs(e) and _rel(H, gold[f"p5k_lin_{i}_H"]) <= TIGHT and _rel(b, gold[f"p5k_lin_{i}_b"]) <= TIGHT
    for key, name in (("p5k_align", "launch"), ("p5k_align_eps_1e-4", "eps_1e-4")):
        o = Oracle(**R.ALIGN_CASES[name])
        o.set_source(s5); o.set_target(t5)
        check_align(_run(o), gold, key)
    scans = R.make_chain5k()
    o = Oracle(**R.LIN_CASES["launch"])
    o.set_target(scans[0])
    for i in range(1, len(scans)):
        if i > 1:
            o.swap()      # the oracle (and the product) keep the previous source's covariances; same numbers as a fresh setInputTarget
        o.set_source(scans[i])
        rc, T, conv, it = o.align()
        assert rc == 0 and [int(conv), it] == list(gold["chain5k_state"][i - 1, :2].astype(int))
        assert np.abs(T.astype(np.float64) - gold["chain5k_T"][i - 1]).max() <= 1e-7
        assert abs(o.fitness() - gold["chain5k_state"][i - 1, 2]) <= 1e-6 * gold["chain5k_state"][i - 1, 2]
