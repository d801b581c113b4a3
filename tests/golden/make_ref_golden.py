#!/usr/bin/env python
"""Generates tests/golden/apd_ref_golden_v1.npz by running the REFERENCE'S OWN FastAPDGICP sources, compiled unmodified from
/root/reference over stand-in Eigen / PCL / Boost headers (oracle/ref_apdgicp.cpp -> oracle/_ref/libref_apdgicp.so), on the
cases of tests/ref_cases.py and tests/lm_cases.py. One thread (per-thread partial sums added in index order), atan2f correctly
rounded (the SURVEY 8c convention; the same linearization with the C library's atan2f is stored beside it as *_libc).

Needs /root/reference; the vectors travel to the GPU box in its place.

    python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import lm_cases
    import ref_cases as R
    from oracle import refapd
    refapd.build(force=True)
    refapd.set_atan2f_mode(True)
    src, tgt, _ = R.make_pair()
    out = {"version": np.array(refapd.lib().ref_apd_version().decode())}

    clouds = {"src": src, "tgt": tgt}

    def fresh(params):
        r = refapd.RefAPD(**params)
        r.set_source(clouds["src"]); r.set_target(clouds["tgt"])
        return r

    for name, p in R.COV_CASES.items():
        r = fresh(p)
        assert r.compute_covariances() == 0
        out[f"cov_{name}_src"] = r.covariances(0)
        if name == "plane_k20":
            out[f"cov_{name}_tgt"] = r.covariances(1)

    for name, p in R.LIN_CASES.items():
        for i, P in enumerate(R.poses()):
            r = fresh(p)
            e, H, b = r.linearize_d(P)
            corr, sq = r.correspondences()
            Q = np.array(P); Q[:3, 3] += [0.01, 0.02, -0.01]
            out[f"lin_{name}_{i}_e"] = np.array([e, r.compute_error_d(Q)])
            out[f"lin_{name}_{i}_H"] = H
            out[f"lin_{name}_{i}_b"] = b
            out[f"lin_{name}_{i}_corr"] = corr
            out[f"lin_{name}_{i}_sq"] = sq
            if name == "launch":
                out[f"lin_{name}_{i}_mahal"] = r.mahalanobis()
    # the reference as it runs on this machine: the C library's atan2f (not correctly rounded before glibc 2.41)
    refapd.set_atan2f_mode(False)
    r = fresh(R.LIN_CASES["launch"])
    e, H, b = r.linearize_d(R.poses()[1])
    out["lin_launch_1_libc_e"] = np.array(e); out["lin_launch_1_libc_H"] = H; out["lin_launch_1_libc_b"] = b
    refapd.set_atan2f_mode(True)

    def store_align(key, r, g):
        rc, T, conv, it = r.align(g)
        assert rc == 0
        tr = r.trace()
        if tr.size:   # no decision inside rounding noise: the vectors must be reproducible by any faithful implementation
            assert (np.abs(tr[:, 2] - tr[:, 3]) > 1e-11 * np.abs(tr[:, 2])).all(), (key, tr)
            assert (np.abs(tr[:, 4]) > 1e-3).all(), (key, tr)
        out[f"{key}_T"] = T
        out[f"{key}_state"] = np.array([int(conv), it, int(r.lm_failed())])
        out[f"{key}_trace"] = tr
        out[f"{key}_final_hessian"] = r.final_hessian()
        out[f"{key}_aligned_head"] = r.aligned[:64].copy()
        print(f"{key}: converged {conv} iterations {it} rows {len(tr)} rejected {int((tr[:, 7] == 0).sum()) if tr.size else 0} lm_failed {r.lm_failed()}")

    for name, p in R.ALIGN_CASES.items():
        store_align(f"align_{name}", fresh(p), None)
        store_align(f"align_{name}_guess", fresh(p), R.guess())

    base = fresh(lm_cases.LAUNCH)
    assert base.compute_covariances() == 0
    cov_src = base.covariances(0)
    for name in lm_cases.CASES:
        cov_tgt = lm_cases.injected_target_covariances(name, base.covariances(1))
        r = fresh(lm_cases.case_params(name))
        r.set_covariances(0, cov_src); r.set_covariances(1, cov_tgt)
        store_align(f"lm_{name}", r, None)

    # BASELINE size: a 5000-point pair and an odometry chain with swapSourceAndTarget between the pairs
    s5, t5, _ = R.make_pair5k()
    r = refapd.RefAPD(**R.LIN_CASES["launch"])
    r.set_source(s5); r.set_target(t5)
    assert r.compute_covariances() == 0
    out["p5k_cov_src"] = r.covariances(0); out["p5k_cov_tgt"] = r.covariances(1)
    for i, P in enumerate(R.poses()[:2]):
        e, H, b = r.linearize_d(P)
        corr, sq = r.correspondences()
        out[f"p5k_lin_{i}_e"] = np.array(e); out[f"p5k_lin_{i}_H"] = H; out[f"p5k_lin_{i}_b"] = b
        out[f"p5k_lin_{i}_corr"] = corr; out[f"p5k_lin_{i}_sq"] = sq
    clouds["src"], clouds["tgt"] = s5, t5
    store_align("p5k_align", fresh(R.ALIGN_CASES["launch"]), None)
    store_align("p5k_align_eps_1e-4", fresh(R.ALIGN_CASES["eps_1e-4"]), None)
    scans = R.make_chain5k()
    r = refapd.RefAPD(**R.LIN_CASES["launch"])
    Ts, st = [], []
    for i in range(1, len(scans)):
        # the nodelet's own sequence (scan_matching_odometry_nodelet.cpp:449-468, 584-592): setInputTarget(previous) + setInputSource(new).
        # (After swapSourceAndTarget the reference's getFitnessScore would use PCL's tree_ of the OLD target - swap does not raise
        # target_cloud_updated_ - see tests/test_reference_apdgicp.py::test_fitness_after_swap_uses_pcls_stale_tree; no RIV-SLAM caller swaps.)
        r.set_target(scans[i - 1])
        r.set_source(scans[i])
        rc, T, conv, it = r.align()
        assert rc == 0
        tr = r.trace()
        assert (np.abs(tr[:, 2] - tr[:, 3]) > 1e-11 * np.abs(tr[:, 2])).all() and (np.abs(tr[:, 4]) > 1e-3).all()
        Ts.append(T); st.append([int(conv), it, r.fitness()])
    out["chain5k_T"] = np.array(Ts); out["chain5k_state"] = np.array(st)
    print("chain:", st)

    path = os.path.join(HERE, "apd_ref_golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
