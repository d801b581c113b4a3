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
    assert ["notarget", "converged", "0"] in lines
    assert "No input target dataset" in r.stderr
