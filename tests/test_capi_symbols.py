"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
that include/apdgicp_b200.h declares. No compute entry point is called without a GPU — except to
confirm that the product refuses to run without one (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from riv_slam_b200 import build
    return build.build_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "apdgicp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    syms = _declared_symbols()
    assert len(syms) >= 30
    L = C.CDLL(lib_path)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_python_binding_covers_the_header(lib_path):
    from riv_slam_b200 import fast_apdgicp as F
    assert sorted(F._PROTOTYPES) == _declared_symbols()
    L = F.load_library()
    assert L.apd_abi_version() == 1
    p = F.ApdParams()
    assert L.apd_default_params(C.byref(p)) == 0
    # constructor defaults of the reference: fast_apdgicp_impl.hpp:14-28, lsq_registration_impl.hpp:11-24
    assert (p.k_correspondences, p.regularization, p.max_iterations, p.optimizer, p.lm_max_iterations) == (20, F.PLANE, 64, F.LevenbergMarquardt, 10)
    assert (p.rotation_epsilon, p.transformation_epsilon, p.lm_init_lambda_factor) == (2e-3, 5e-4, 1e-9)
    assert (p.dist_var, p.azimuth_var, p.elevation_var) == (0.86, 0.5, 1.0)
    assert p.max_corr_dist == float(C.c_float(3.4028234663852886e38).value)
    assert C.sizeof(F.ApdResult) == 96


def test_no_cpu_fallback(lib_path):
    import torch
    from riv_slam_b200 import fast_apdgicp as F
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(F.ApdError) as e:
        F.FastAPDGICP(0)
    assert e.value.code == F.APD_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "riv-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("CPU oracle", "").replace("the oracle", "").replace("oracle/linalg.hpp", ""), f


def test_build_staleness_is_decided_by_content_hash(lib_path):
    """A snapshot copied to the GPU box does not keep a usable mtime order: the library is stale only when the hash
    of its sources and flags differs from the one stored beside it (an 8-rank bench once raced on a needless rebuild)."""
    from riv_slam_b200 import build
    assert os.path.exists(build.HASH_PATH)
    assert not build.needs_build()
    src = os.path.join(build.CSRC, build.SOURCES[0])
    st = os.stat(src)
    try:
        os.utime(src, (st.st_atime, st.st_mtime + 3600))   # "newer" source, same content
        assert not build.needs_build()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    assert open(build.HASH_PATH).read().strip() == build._source_hash()
