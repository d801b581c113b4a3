"""TEST INFRASTRUCTURE - ctypes face of oracle/_ref/libref_apdgicp.so: the REFERENCE's own FastAPDGICP sources compiled
unmodified from /root/reference over stand-in Eigen / PCL / Boost headers (oracle/ref_apdgicp.cpp, `make -C oracle ref`).

Exists only where /root/reference is mounted (the build container); it validates the oracle restatement
(tests/test_reference_apdgicp.py) and generates tests/golden/apd_ref_golden_v1.npz (tests/golden/make_ref_golden.py),
which travels to the GPU box. The class mirrors oracle.oracle.Oracle method for method. Never used by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import OracleParams, _f32, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref_apdgicp.so")
REFERENCE = os.environ.get("REFERENCE", "/root/reference")


def available() -> bool:
    """True when the library exists or can be built (the reference tree is mounted)."""
    return os.path.exists(_LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "fast_apdgicp", "include"))


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "_ref/libref_apdgicp.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ref_apd_create.restype = C.c_void_p
        L.ref_apd_destroy.argtypes = [C.c_void_p]
        L.ref_apd_get_defaults.argtypes = [C.POINTER(OracleParams)]
        L.ref_apd_set_params.argtypes = [C.c_void_p, C.POINTER(OracleParams)]
        L.ref_apd_set_source.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
        L.ref_apd_set_target.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
        for name in ("ref_apd_swap", "ref_apd_clear_source", "ref_apd_clear_target", "ref_apd_lm_failed", "ref_apd_compute_covariances"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_apd_align.argtypes = [C.c_void_p, fp, C.c_int, fp, ip, ip, fp]
        L.ref_apd_fitness.argtypes = [C.c_void_p, C.c_double]
        L.ref_apd_fitness.restype = C.c_double
        L.ref_apd_get_debug_text.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_apd_get_covariances.argtypes = [C.c_void_p, C.c_int, dp]
        L.ref_apd_get_covariances4.argtypes = [C.c_void_p, C.c_int, dp]
        L.ref_apd_set_covariances.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.ref_apd_linearize.argtypes = [C.c_void_p, fp, dp, dp]
        L.ref_apd_linearize.restype = C.c_double
        L.ref_apd_linearize_d.argtypes = [C.c_void_p, dp, dp, dp]
        L.ref_apd_linearize_d.restype = C.c_double
        L.ref_apd_compute_error_d.argtypes = [C.c_void_p, dp]
        L.ref_apd_compute_error_d.restype = C.c_double
        L.ref_apd_get_correspondences.argtypes = [C.c_void_p, ip, fp]
        L.ref_apd_get_mahalanobis.argtypes = [C.c_void_p, dp]
        L.ref_apd_get_final_hessian.argtypes = [C.c_void_p, dp]
        L.ref_skewd.argtypes = [dp, dp]
        L.ref_so3_exp.argtypes = [dp, dp, dp]
        L.ref_apd_version.restype = C.c_char_p
        L.ref_eigen_svd3.argtypes = [dp, dp, dp, dp]
        L.ref_eigen_ldlt6_solve.argtypes = [dp, dp, dp]
        L.ref_eigen_inverse.argtypes = [dp, C.c_int, dp]
        L.ref_eigen_yaw_pitch.argtypes = [C.c_double, C.c_double, dp]
        L.ref_apd_set_atan2f_mode.argtypes = [C.c_int]
        _lib = L
    return _lib


def set_atan2f_mode(correctly_rounded: bool):
    """APD_I:168,172,173 call atan2f. False: the C library's (the reference as it runs on this machine; within 1 ulp).
    True: the double result rounded once - the convention SURVEY 8c fixes for the oracle and the device."""
    lib().ref_apd_set_atan2f_mode(int(bool(correctly_rounded)))


def defaults() -> OracleParams:
    """The values the reference's constructors leave behind (APD_I:14-28, LSQ_I:11-24, APD_H:107-109)."""
    p = OracleParams()
    lib().ref_apd_get_defaults(C.byref(p))
    return p


def skewd(x):
    a = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(9)
    lib().ref_skewd(_ptr(a, C.c_double), _ptr(out, C.c_double))
    return out.reshape(3, 3)


def so3_exp(omega):
    a = np.ascontiguousarray(omega, dtype=np.float64)
    q = np.zeros(4)
    R = np.zeros(9)
    lib().ref_so3_exp(_ptr(a, C.c_double), _ptr(q, C.c_double), _ptr(R, C.c_double))
    return q, R.reshape(3, 3)


# ---- the stand-in Eigen pieces on their own (tests/test_reference_apdgicp.py holds them against numpy / LAPACK) ----

def eigen_svd3(m):
    a = np.ascontiguousarray(m, dtype=np.float64).reshape(9)
    U, S, V = np.zeros(9), np.zeros(3), np.zeros(9)
    lib().ref_eigen_svd3(_ptr(a, C.c_double), _ptr(U, C.c_double), _ptr(S, C.c_double), _ptr(V, C.c_double))
    return U.reshape(3, 3), S, V.reshape(3, 3)


def eigen_ldlt6_solve(A, b):
    a = np.ascontiguousarray(A, dtype=np.float64).reshape(36)
    bb = np.ascontiguousarray(b, dtype=np.float64).reshape(6)
    x = np.zeros(6)
    lib().ref_eigen_ldlt6_solve(_ptr(a, C.c_double), _ptr(bb, C.c_double), _ptr(x, C.c_double))
    return x


def eigen_inverse(m):
    a = np.ascontiguousarray(m, dtype=np.float64)
    n = a.shape[0]
    out = np.zeros(n * n)
    lib().ref_eigen_inverse(_ptr(a.reshape(-1), C.c_double), n, _ptr(out, C.c_double))
    return out.reshape(n, n)


def eigen_yaw_pitch(yaw, pitch):
    R = np.zeros(9)
    lib().ref_eigen_yaw_pitch(float(yaw), float(pitch), _ptr(R, C.c_double))
    return R.reshape(3, 3)


def parse_lm_table(text: str) -> np.ndarray:
    """The reference's own LM debug table (LSQ_I:148-155) -> rows (outer, inner, y0, yi, rho, lambda, |delta|, accepted) like the
    oracle's trace. `accepted` is the branch LSQ_I:156 takes (NOT rho < 0), not the printed 'x' (rho > 0)."""
    rows = []
    outer = -1
    for line in text.splitlines():
        t = line.split()
        if line.startswith("--- LM optimization ---"):
            outer += 1
            continue
        if len(t) < 6 or not t[0].lstrip("-").isdigit():
            continue
        inner = int(t[0])
        y0, yi, rho, lam, dn = (float(v) for v in t[1:6])
        rows.append((outer, inner, y0, yi, rho, lam, dn, 0.0 if rho < 0 else 1.0))
    return np.array(rows, dtype=np.float64).reshape(-1, 8)


class RefAPD:
    """fast_gicp::FastAPDGICP<pcl::PointXYZI, pcl::PointXYZI> itself, behind the same surface as oracle.oracle.Oracle."""

    def __init__(self, **params):
        self.L = lib()
        self.h = C.c_void_p(self.L.ref_apd_create())
        self.p = defaults()
        self.p.num_threads = 1   # one thread: the per-thread H / b partial sums (APD_I:201-206, 262-270) are then added in index order
        self.set_params(**params)
        self.n_src = self.n_tgt = 0
        self._trace = np.zeros((0, 8))

    def __del__(self):
        try:
            self.L.ref_apd_destroy(self.h)
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.p, k):
                raise AttributeError(k)
            setattr(self.p, k, v)
        self.L.ref_apd_set_params(self.h, C.byref(self.p))

    @property
    def k(self):
        return self.p.k_correspondences

    def set_source(self, pts):
        a = _f32(pts)
        self.n_src = a.shape[0]
        self.L.ref_apd_set_source(self.h, _ptr(a, C.c_float), a.shape[1], a.shape[0])

    def set_target(self, pts):
        a = _f32(pts)
        self.n_tgt = a.shape[0]
        self.L.ref_apd_set_target(self.h, _ptr(a, C.c_float), a.shape[1], a.shape[0])

    def swap(self):
        self.L.ref_apd_swap(self.h)
        self.n_src, self.n_tgt = self.n_tgt, self.n_src

    def clear_source(self):
        self.L.ref_apd_clear_source(self.h)

    def clear_target(self):
        self.L.ref_apd_clear_target(self.h)

    def align(self, guess=None, debug=True):
        g = _f32(np.eye(4) if guess is None else guess).reshape(16)
        T = np.zeros(16, dtype=np.float32)
        conv, it = C.c_int(0), C.c_int(0)
        self.aligned = np.zeros((self.n_src, 3), dtype=np.float32)
        rc = self.L.ref_apd_align(self.h, _ptr(g, C.c_float), int(debug), _ptr(T, C.c_float), C.byref(conv), C.byref(it), _ptr(self.aligned, C.c_float))
        if debug and rc == 0:
            n = self.L.ref_apd_get_debug_text(self.h, None, 0)
            buf = C.create_string_buffer(n + 1)
            self.L.ref_apd_get_debug_text(self.h, buf, n + 1)
            self.debug_text = buf.value.decode()
            self._trace = parse_lm_table(self.debug_text)
        return rc, T.reshape(4, 4), bool(conv.value), it.value

    def lm_failed(self) -> bool:
        return bool(self.L.ref_apd_lm_failed(self.h))

    def fitness(self, max_range=float(np.finfo(np.float64).max)):
        return self.L.ref_apd_fitness(self.h, max_range)

    def trace(self):
        return self._trace

    def compute_covariances(self):
        return self.L.ref_apd_compute_covariances(self.h)

    def covariances(self, which):
        n = self.L.ref_apd_get_covariances(self.h, which, None)
        out = np.zeros((n, 3, 3))
        self.L.ref_apd_get_covariances(self.h, which, _ptr(out, C.c_double))
        return out

    def covariances4(self, which):
        n = self.L.ref_apd_get_covariances4(self.h, which, None)
        out = np.zeros((n, 4, 4))
        self.L.ref_apd_get_covariances4(self.h, which, _ptr(out, C.c_double))
        return out

    def set_covariances(self, which, covs):
        a = np.ascontiguousarray(covs, dtype=np.float64)
        self.L.ref_apd_set_covariances(self.h, which, _ptr(a, C.c_double), a.shape[0])

    def linearize(self, pose):
        g = _f32(pose).reshape(16)
        H, b = np.zeros(36), np.zeros(6)
        e = self.L.ref_apd_linearize(self.h, _ptr(g, C.c_float), _ptr(H, C.c_double), _ptr(b, C.c_double))
        return e, H.reshape(6, 6), b

    def linearize_d(self, pose):
        g = np.ascontiguousarray(pose, dtype=np.float64).reshape(16)
        H, b = np.zeros(36), np.zeros(6)
        e = self.L.ref_apd_linearize_d(self.h, _ptr(g, C.c_double), _ptr(H, C.c_double), _ptr(b, C.c_double))
        return e, H.reshape(6, 6), b

    def compute_error_d(self, pose):
        g = np.ascontiguousarray(pose, dtype=np.float64).reshape(16)
        return self.L.ref_apd_compute_error_d(self.h, _ptr(g, C.c_double))

    def correspondences(self):
        corr = np.zeros(self.n_src, dtype=np.int32)
        sq = np.zeros(self.n_src, dtype=np.float32)
        self.L.ref_apd_get_correspondences(self.h, _ptr(corr, C.c_int), _ptr(sq, C.c_float))
        return corr, sq

    def mahalanobis(self):
        out = np.zeros((self.n_src, 3, 3))
        self.L.ref_apd_get_mahalanobis(self.h, _ptr(out, C.c_double))
        return out

    def final_hessian(self):
        H = np.zeros(36)
        self.L.ref_apd_get_final_hessian(self.h, _ptr(H, C.c_double))
        return H.reshape(6, 6)
