"""TEST INFRASTRUCTURE — ctypes face of the CPU oracle (oracle/liboracle.so).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module. The product (``riv_slam_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
OPT_GAUSS_NEWTON, OPT_LEVENBERG_MARQUARDT = 0, 1


class OracleParams(C.Structure):
    _fields_ = [
        ("num_threads", C.c_int),
        ("k_correspondences", C.c_int),
        ("regularization", C.c_int),
        ("max_iterations", C.c_int),
        ("optimizer", C.c_int),
        ("lm_max_iterations", C.c_int),
        ("max_corr_dist", C.c_double),
        ("rotation_epsilon", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("lm_init_lambda_factor", C.c_double),
        ("dist_var", C.c_double),
        ("azimuth_var", C.c_double),
        ("elevation_var", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (no-op when up to date)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_default_params.argtypes = [C.POINTER(OracleParams)]
        L.oracle_set_params.argtypes = [C.c_void_p, C.POINTER(OracleParams)]
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.oracle_set_source.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
        L.oracle_set_target.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
        for name in ("oracle_swap", "oracle_clear_source", "oracle_clear_target"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_align.argtypes = [C.c_void_p, fp, fp, ip, ip]
        L.oracle_align.restype = C.c_int
        L.oracle_fitness.argtypes = [C.c_void_p, C.c_double]
        L.oracle_fitness.restype = C.c_double
        L.oracle_fitness_score.argtypes = [C.c_void_p, fp, C.c_double]
        L.oracle_fitness_score.restype = C.c_double
        L.oracle_compute_covariances.argtypes = [C.c_void_p]
        L.oracle_linearize.argtypes = [C.c_void_p, fp, dp, dp]
        L.oracle_linearize.restype = C.c_double
        L.oracle_linearize_d.argtypes = [C.c_void_p, dp, dp, dp]
        L.oracle_linearize_d.restype = C.c_double
        L.oracle_compute_error_d.argtypes = [C.c_void_p, dp]
        L.oracle_compute_error_d.restype = C.c_double
        L.oracle_lm_failed.argtypes = [C.c_void_p]
        L.oracle_get_knn.argtypes = [C.c_void_p, C.c_int, ip]
        L.oracle_get_covariances.argtypes = [C.c_void_p, C.c_int, dp]
        L.oracle_set_covariances.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.oracle_get_correspondences.argtypes = [C.c_void_p, ip, fp]
        L.oracle_get_mahalanobis.argtypes = [C.c_void_p, dp]
        L.oracle_get_final_hessian.argtypes = [C.c_void_p, dp]
        L.oracle_get_trace.argtypes = [C.c_void_p, dp, C.c_int]
        L.oracle_transform_source.argtypes = [C.c_void_p, fp, fp]
        L.oracle_knn_bruteforce.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, ip, fp]
        L.oracle_knn_kdtree.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, ip, fp]
        L.oracle_timed_registration.argtypes = [C.c_void_p, fp, C.c_int, fp, C.c_int, C.c_int, fp, C.c_int, fp, ip, ip, dp]
        L.oracle_timed_registration.restype = C.c_double
        L.oracle_max_threads.restype = C.c_int
        L.oracle_distance_filter.argtypes = [fp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, fp]
        L.oracle_voxel_grid.argtypes = [fp, C.c_int, C.c_float, fp]
        L.oracle_approx_voxel_grid.argtypes = [fp, C.c_int, C.c_float, fp]
        L.oracle_accumulate_submap_m.argtypes = [fp, ip, C.c_int, dp, C.c_float, C.c_int, fp]
        L.oracle_radius_outlier_removal.argtypes = [fp, C.c_int, C.c_double, C.c_int, fp]
        L.oracle_statistical_outlier_removal.argtypes = [fp, C.c_int, C.c_int, C.c_double, fp]
        L.oracle_accumulate_submap.argtypes = [fp, ip, C.c_int, dp, C.c_float, fp]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Mirror of the FastAPDGICP surface on the CPU oracle."""

    def __init__(self, **params):
        self.L = lib()
        self.h = C.c_void_p(self.L.oracle_create())
        self.p = OracleParams()
        self.L.oracle_default_params(C.byref(self.p))
        self.set_params(**params)
        self.n_src = self.n_tgt = 0

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.p, k):
                raise AttributeError(k)
            setattr(self.p, k, v)
        self.L.oracle_set_params(self.h, C.byref(self.p))

    @property
    def k(self):
        return self.p.k_correspondences

    def set_source(self, pts):
        a = _f32(pts)
        self.n_src = a.shape[0]
        self.L.oracle_set_source(self.h, _ptr(a, C.c_float), a.shape[1], a.shape[0])

    def set_target(self, pts):
        a = _f32(pts)
        self.n_tgt = a.shape[0]
        self._tgt_np = np.ascontiguousarray(a[:, :3])
        self.L.oracle_set_target(self.h, _ptr(a, C.c_float), a.shape[1], a.shape[0])

    def swap(self):
        self.L.oracle_swap(self.h)
        self.n_src, self.n_tgt = self.n_tgt, self.n_src

    def align(self, guess=None):
        g = _f32(np.eye(4) if guess is None else guess).reshape(16)
        T = np.zeros(16, dtype=np.float32)
        conv = C.c_int(0)
        it = C.c_int(0)
        rc = self.L.oracle_align(self.h, _ptr(g, C.c_float), _ptr(T, C.c_float), C.byref(conv), C.byref(it))
        return rc, T.reshape(4, 4), bool(conv.value), it.value

    def lm_failed(self) -> bool:
        """True when the last align ended with "lm not converged!!" (LSQ_I:71-74)."""
        return bool(self.L.oracle_lm_failed(self.h))

    def fitness(self, max_range=float(np.finfo(np.float64).max)):
        return self.L.oracle_fitness(self.h, max_range)

    def fitness_score(self, T, max_range=float(np.finfo(np.float64).max)):
        """calc_fitness_score(target=cloud1, source=cloud2, relpose=T, max_range)"""
        g = _f32(T).reshape(16)
        return self.L.oracle_fitness_score(self.h, _ptr(g, C.c_float), max_range)

    def compute_covariances(self):
        return self.L.oracle_compute_covariances(self.h)

    def linearize(self, pose):
        g = _f32(pose).reshape(16)
        H = np.zeros(36)
        b = np.zeros(6)
        e = self.L.oracle_linearize(self.h, _ptr(g, C.c_float), _ptr(H, C.c_double), _ptr(b, C.c_double))
        return e, H.reshape(6, 6), b

    def linearize_d(self, pose):
        """linearize at a double pose (the protected hook takes an Isometry3d, APD_I:198-272)."""
        g = np.ascontiguousarray(pose, dtype=np.float64).reshape(16)
        H = np.zeros(36)
        b = np.zeros(6)
        e = self.L.oracle_linearize_d(self.h, _ptr(g, C.c_double), _ptr(H, C.c_double), _ptr(b, C.c_double))
        return e, H.reshape(6, 6), b

    def compute_error_d(self, pose):
        """compute_error at a double pose with the correspondences of the last linearize (APD_I:275-298)."""
        g = np.ascontiguousarray(pose, dtype=np.float64).reshape(16)
        return self.L.oracle_compute_error_d(self.h, _ptr(g, C.c_double))

    def inlier_count(self, T, max_dist):
        """publish_scan_matching_status (scan_matching_odometry_nodelet.cpp:698-712): aligned points whose nearest target point
        is strictly closer than max_dist (float squared distance < double max_dist^2)."""
        q = self.transform_source(T)
        idx, d2 = knn_kdtree(self._tgt_np, q, 1)
        return int((d2[:, 0].astype(np.float64) < float(max_dist) * float(max_dist)).sum())

    def knn(self, which):
        n = self.n_tgt if which else self.n_src
        out = np.zeros((n, self.k), dtype=np.int32)
        self.L.oracle_get_knn(self.h, which, _ptr(out, C.c_int))
        return out

    def covariances(self, which):
        n = self.n_tgt if which else self.n_src
        out = np.zeros((n, 3, 3))
        got = self.L.oracle_get_covariances(self.h, which, _ptr(out, C.c_double))
        return out[:got]

    def set_covariances(self, which, covs):
        a = np.ascontiguousarray(covs, dtype=np.float64)
        self.L.oracle_set_covariances(self.h, which, _ptr(a, C.c_double), a.shape[0])

    def correspondences(self):
        corr = np.zeros(self.n_src, dtype=np.int32)
        sq = np.zeros(self.n_src, dtype=np.float32)
        self.L.oracle_get_correspondences(self.h, _ptr(corr, C.c_int), _ptr(sq, C.c_float))
        return corr, sq

    def mahalanobis(self):
        out = np.zeros((self.n_src, 3, 3))
        self.L.oracle_get_mahalanobis(self.h, _ptr(out, C.c_double))
        return out

    def final_hessian(self):
        H = np.zeros(36)
        self.L.oracle_get_final_hessian(self.h, _ptr(H, C.c_double))
        return H.reshape(6, 6)

    def trace(self):
        n = self.L.oracle_get_trace(self.h, None, 0)
        out = np.zeros((n, 8))
        if n:
            self.L.oracle_get_trace(self.h, _ptr(out, C.c_double), n)
        return out

    def transform_source(self, T):
        g = _f32(T).reshape(16)
        out = np.zeros((self.n_src, 3), dtype=np.float32)
        self.L.oracle_transform_source(self.h, _ptr(g, C.c_float), _ptr(out, C.c_float))
        return out

    def timed_registration(self, src, tgt, guess=None, reuse_target=False):
        s = _f32(src)
        t = _f32(tgt)
        g = _f32(np.eye(4) if guess is None else guess).reshape(16)
        T = np.zeros(16, dtype=np.float32)
        conv = C.c_int(0)
        it = C.c_int(0)
        fit = C.c_double(0)
        self.n_src, self.n_tgt = s.shape[0], t.shape[0]
        sec = self.L.oracle_timed_registration(self.h, _ptr(s, C.c_float), s.shape[0], _ptr(t, C.c_float), t.shape[0], s.shape[1],
                                               _ptr(g, C.c_float), int(reuse_target), _ptr(T, C.c_float), C.byref(conv), C.byref(it), C.byref(fit))
        return sec, T.reshape(4, 4), bool(conv.value), it.value, fit.value


def knn_bruteforce(cloud, queries, k):
    c = _f32(np.asarray(cloud)[:, :3])
    q = _f32(np.asarray(queries)[:, :3])
    idx = np.zeros((q.shape[0], k), dtype=np.int32)
    d2 = np.zeros((q.shape[0], k), dtype=np.float32)
    lib().oracle_knn_bruteforce(_ptr(c, C.c_float), c.shape[0], _ptr(q, C.c_float), q.shape[0], k, _ptr(idx, C.c_int), _ptr(d2, C.c_float))
    return idx, d2


def knn_kdtree(cloud, queries, k):
    c = _f32(np.asarray(cloud)[:, :3])
    q = _f32(np.asarray(queries)[:, :3])
    idx = np.zeros((q.shape[0], k), dtype=np.int32)
    d2 = np.zeros((q.shape[0], k), dtype=np.float32)
    lib().oracle_knn_kdtree(_ptr(c, C.c_float), c.shape[0], _ptr(q, C.c_float), q.shape[0], k, _ptr(idx, C.c_int), _ptr(d2, C.c_float))
    return idx, d2


def max_threads() -> int:
    return lib().oracle_max_threads()


# ---- preprocessing filters and submap accumulation (oracle/preprocess_oracle.hpp); clouds are (n, 4) x y z intensity ----

def _xyzi(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def distance_filter(cloud, near_thresh, far_thresh, z_low, z_high):
    a = _xyzi(cloud)
    out = np.zeros_like(a)
    n = lib().oracle_distance_filter(_ptr(a, C.c_float), a.shape[0], near_thresh, far_thresh, z_low, z_high, _ptr(out, C.c_float))
    return out[:n]


def voxel_grid(cloud, leaf):
    a = _xyzi(cloud)
    out = np.zeros_like(a)
    n = lib().oracle_voxel_grid(_ptr(a, C.c_float), a.shape[0], leaf, _ptr(out, C.c_float))
    return out[:n]


def approx_voxel_grid(cloud, leaf):
    """pcl::ApproximateVoxelGrid (downsample_method APPROX_VOXELGRID)."""
    a = _xyzi(cloud)
    out = np.zeros_like(a)
    n = lib().oracle_approx_voxel_grid(_ptr(a, C.c_float), a.shape[0], leaf, _ptr(out, C.c_float))
    return out[:n]


def radius_outlier_removal(cloud, radius, min_pts):
    a = _xyzi(cloud)
    out = np.zeros_like(a)
    n = lib().oracle_radius_outlier_removal(_ptr(a, C.c_float), a.shape[0], radius, min_pts, _ptr(out, C.c_float))
    return out[:n]


def statistical_outlier_removal(cloud, mean_k, stddev_mult):
    a = _xyzi(cloud)
    out = np.zeros_like(a)
    n = lib().oracle_statistical_outlier_removal(_ptr(a, C.c_float), a.shape[0], mean_k, stddev_mult, _ptr(out, C.c_float))
    return out[:n]


def accumulate_submap(clouds, rel_poses, leaf=0.0, approx=False):
    off = np.zeros(len(clouds) + 1, dtype=np.int32)
    off[1:] = np.cumsum([c.shape[0] for c in clouds])
    a = _xyzi(np.concatenate(clouds))
    P = np.ascontiguousarray(rel_poses, dtype=np.float64).reshape(len(clouds), 16)
    out = np.zeros_like(a)
    n = lib().oracle_accumulate_submap_m(_ptr(a, C.c_float), _ptr(off, C.c_int), len(clouds), _ptr(P, C.c_double), leaf, int(approx), _ptr(out, C.c_float))
    return out[:n]


# ---- oracle/_ref: the reference's own vendored exact kd-tree (nanoflann 1.3.2), see oracle/ref_nanoflann.cpp ----

_REF_PATH = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref_nanoflann_knn(cloud, queries, k, leaf_size=10):
    """kNN by the kd-tree the reference vendors (compiled from /root/reference by `make -C oracle ref`): (idx int32 (m, k), d2 float32 (m, k))."""
    global _ref
    if _ref is None:
        L = C.CDLL(_REF_PATH)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.ref_nanoflann_knn.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, ip, fp]
        _ref = L
    c = _f32(np.asarray(cloud)[:, :3])
    q = _f32(np.asarray(queries)[:, :3])
    idx = np.zeros((q.shape[0], k), dtype=np.int32)
    d2 = np.zeros((q.shape[0], k), dtype=np.float32)
    _ref.ref_nanoflann_knn(_ptr(c, C.c_float), c.shape[0], _ptr(q, C.c_float), q.shape[0], k, leaf_size, _ptr(idx, C.c_int), _ptr(d2, C.c_float))
    return idx, d2
