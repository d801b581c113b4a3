"""The C++ drop-in class (include/fast_gicp/gicp/fast_apdgicp.hpp) compiled against the in-container
PCL/Eigen stand-ins and driven through a pcl::Registration base pointer like the reference nodelets."""
import os
import subprocess

import numpy as np
import pytest

from conftest import LAUNCH_PARAMS, ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "_dropin_test")


def _build():
    from riv_slam_b200 import build
    build.build_library()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp")
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include", "pcl_shim"), "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", os.path.join(ROOT, "riv-slam_b200"), "-lapdgicp_b200", "-Wl,-rpath," + os.path.join(ROOT, "riv-slam_b200")]
    subprocess.run(cmd, check=True, capture_output=True)
    return EXE


def test_dropin_header_compiles_and_links():
    exe = _build()
    assert os.path.exists(exe)
    # the executable depends on the product library only through the C ABI
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    used = sorted({l.split()[-1] for l in out.splitlines() if " apd_" in l})
    assert "apd_align" in used and "apd_set_source" in used and "apd_set_target" in used
    assert all(u.startswith("apd_") for u in used)


@pytest.mark.gpu
def test_dropin_matches_python_path_and_oracle(tmp_path):
    from oracle.oracle import Oracle
    from riv_slam_b200 import datagen
    exe = _build()
    scans, _ = datagen.make_sequence(2, 1, n_scans=4, n_points=1500)
    path = tmp_path / "scans.bin"
    with open(path, "wb") as f:
        f.write(np.int32(len(scans)).tobytes())
        for s in scans:
            blk = np.zeros((s.shape[0], 8), dtype=np.float32)
            blk[:, :3] = s[:, :3]
            blk[:, 3] = 1.0
            blk[:, 4] = s[:, 3]
            f.write(np.int32(s.shape[0]).tobytes())
            f.write(blk.tobytes())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.splitlines()]
    pairs = [l for l in lines if l[0] == "pair"]
    assert len(pairs) == len(scans) - 1
    for t, l in enumerate(pairs):
        conv = int(l[3])
        T = np.array(l[5:21], dtype=np.float64).reshape(4, 4)
        fit_pcl, fit_gpu = float(l[22]), float(l[24])
        out0 = np.array(l[26:30], dtype=np.float32)
        o = Oracle(**LAUNCH_PARAMS)
        o.set_source(scans[t + 1]); o.set_target(scans[t])
        rc, T0, conv0, it0 = o.align()
        assert conv == int(conv0)
        assert np.abs(T[:3, :3] - T0[:3, :3]).max() < 1e-5 and np.abs(T[:3, 3] - T0[:3, 3]).max() < 1e-4
        f0 = o.fitness()
        assert abs(fit_gpu - f0) <= 1e-5 * f0 and abs(fit_pcl - f0) <= 1e-5 * f0
        ref0 = o.transform_source(T0)[0]
        assert np.abs(out0[:3] - ref0).max() < 1e-4
        assert out0[3] == scans[t + 1][0, 3]          # intensity rides along in the output cloud
        assert int(l[31]) == scans[t + 1].shape[0]
    # derived-type surface: evaluateCost / getFinalHessian / covariances / swap / clear against the oracle
    T = dict(LAUNCH_PARAMS, transformation_epsilon=1e-6, rotation_epsilon=1e-6)
    o = Oracle(**T)
    o.set_source(scans[1]); o.set_target(scans[0])
    e0, H0, b0 = o.linearize(np.eye(4))
    cost = next(l for l in lines if l[0] == "cost")
    vals = dict(zip(cost[0::2], cost[1::2]))
    assert abs(float(vals["cost"]) - e0) <= 1e-5 * abs(e0)
    for key, ref in (("H00", H0[0, 0]), ("H35", H0[3, 5]), ("H53", H0[5, 3]), ("b0", b0[0]), ("b5", b0[5])):
        assert abs(float(vals[key]) - ref) <= 1e-5 * max(np.abs(H0).max() if key[0] == "H" else np.abs(b0).max(), 1e-30)
    rc, Tf0, conv0, it0 = o.align()
    fwd = next(l for l in lines if l[0] == "fwd")
    assert int(fwd[2]) == int(conv0) and abs(float(fwd[4]) - Tf0[0, 3]) < 1e-4 and abs(float(fwd[6]) - Tf0[1, 3]) < 1e-4
    assert abs(float(fwd[8]) - o.final_hessian()[0, 0]) <= 1e-5 * abs(o.final_hessian()[0, 0])
    covs = next(l for l in lines if l[0] == "covs")
    C0 = o.covariances(0)
    assert int(covs[1]) == scans[1].shape[0] and int(covs[2]) == scans[0].shape[0]
    assert abs(float(covs[4]) - C0[0][0, 0]) < 1e-9 and abs(float(covs[5]) - C0[0][1, 0]) < 1e-9 and float(covs[6]) == 0.0
    o2 = Oracle(**T)
    o2.set_source(scans[0]); o2.set_target(scans[1])
    rc, Tb0, convb, itb = o2.align()
    bwd = next(l for l in lines if l[0] == "bwd")
    assert int(bwd[2]) == int(convb) and int(bwd[3]) == int(convb)
    assert float(bwd[5]) < 1e-6                      # swapped object == fresh object with injected covariances
    assert abs(float(bwd[7]) - Tb0[0, 3]) < 1e-4
    # scan-to-map target built on the device from three keyframes (setInputTargetFromKeyframes)
    from oracle import oracle as O
    rel = [np.eye(4) for _ in range(3)]
    rel[0][0, 3], rel[0][1, 3] = 0.30, -0.02
    rel[1][0, 3], rel[1][1, 3] = 0.15, 0.01
    want = O.accumulate_submap([np.ascontiguousarray(s[:, :4]) for s in scans[:3]], rel, 0.1)
    sm = next(l for l in lines if l[0] == "submap")
    assert int(sm[2]) == len(want)
    assert np.array_equal(np.array(sm[4:8], dtype=np.float32), want[0])
    om = Oracle(**LAUNCH_PARAMS)
    om.set_source(scans[3]); om.set_target(want)
    rc, Tm0, convm, itm = om.align()
    assert int(sm[9]) == int(convm)
    assert np.abs(np.array(sm[11:16:2], dtype=np.float64) - Tm0[:3, 3]).max() < 1e-4
    fm = om.fitness()
    assert abs(float(sm[17]) - fm) <= 1e-5 * fm
    # after clearSource() PCL's align returns from initCompute before touching converged_ (stale value, as in PCL)
    assert any(l[:2] == ["cleared", "converged"] for l in lines)
    assert ["notarget", "converged", "0"] in lines
    assert "No input target dataset" in r.stderr
