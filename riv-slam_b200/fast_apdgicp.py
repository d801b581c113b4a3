"""Host-side mirror of ``fast_gicp::FastAPDGICP<pcl::PointXYZI, pcl::PointXYZI>`` over the C ABI.

Method names, argument meaning and error behaviour follow the reference class
(fast_apdgicp/include/fast_gicp/gicp/fast_apdgicp.hpp:33-110, lsq_registration.hpp:16-84) and the
``pcl::Registration`` members its callers use (radar_graph_slam/apps/scan_matching_odometry_nodelet.cpp:449-482,
radar_graph_slam/src/radar_graph_slam/loop_detector.cpp:222-236), so the parity tests read like
code written against the reference. All computation happens in ``libapdgicp_b200.so`` on the GPU;
this module only marshals pointers. There is no CPU fallback: importing works anywhere, creating
an object without the library or without a CUDA device raises.

Matrices cross this boundary as numpy row-major 4x4 float32 / 6x6 float64, clouds as ``(n, c)``
float32 arrays with x, y, z in the first three columns (``c`` = 3, 4 or 8 for packed xyz, xyzi or
pcl::PointXYZI memory) or as CUDA ``torch`` tensors of the same shape.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("APDGICP_B200_LIB") or os.path.join(_PKG, "libapdgicp_b200.so")  # the override serves kernel-variant experiments

# fast_gicp::RegularizationMethod (gicp/gicp_settings.hpp:6)
NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS = range(5)
# fast_gicp::LSQ_OPTIMIZER_TYPE (gicp/lsq_registration.hpp:13)
GaussNewton, LevenbergMarquardt = 0, 1

APD_OK, APD_ERR_NO_DEVICE, APD_ERR_INVALID, APD_ERR_NO_INPUT, APD_ERR_TOO_FEW_POINTS, APD_ERR_CUDA, APD_ERR_UNSUPPORTED = range(7)
APD_STATUS_LM_FAILED = 100
MEM_HOST, MEM_DEVICE = 0, 1
DBL_MAX = float(np.finfo(np.float64).max)


class ApdParams(C.Structure):
    _fields_ = [
        ("k_correspondences", C.c_int32), ("regularization", C.c_int32), ("max_iterations", C.c_int32),
        ("optimizer", C.c_int32), ("lm_max_iterations", C.c_int32), ("num_threads", C.c_int32),
        ("max_corr_dist", C.c_double), ("rotation_epsilon", C.c_double), ("transformation_epsilon", C.c_double),
        ("lm_init_lambda_factor", C.c_double), ("dist_var", C.c_double), ("azimuth_var", C.c_double),
        ("elevation_var", C.c_double),
    ]


class ApdResult(C.Structure):
    _fields_ = [
        ("T", C.c_float * 16), ("fitness", C.c_double), ("error", C.c_double), ("converged", C.c_int32),
        ("iterations", C.c_int32), ("status", C.c_int32), ("num_inliers", C.c_int32),
    ]


class ApdPreprocessParams(C.Structure):
    """apd_preprocess_params: the PreprocessingNodelet parameters (preprocessing_nodelet.cpp:137-205)."""
    _fields_ = [
        ("use_distance_filter", C.c_int32), ("outlier_removal", C.c_int32), ("radius_min_neighbors", C.c_int32), ("statistical_mean_k", C.c_int32),
        ("distance_near_thresh", C.c_double), ("distance_far_thresh", C.c_double), ("z_low_thresh", C.c_double), ("z_high_thresh", C.c_double),
        ("downsample_resolution", C.c_double), ("radius_radius", C.c_double), ("statistical_stddev", C.c_double),
    ]


RESULT_DTYPE = np.dtype([("T", np.float32, (4, 4)), ("fitness", np.float64), ("error", np.float64), ("converged", np.int32),
                         ("iterations", np.int32), ("status", np.int32), ("num_inliers", np.int32)])
assert RESULT_DTYPE.itemsize == C.sizeof(ApdResult) == 96


class ApdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"apdgicp_b200 error {code}: {msg}")
        self.code = code


_lib = None
_fp, _dp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)

# every symbol include/apdgicp_b200.h declares: (restype, argtypes)
_PROTOTYPES = {
    "apd_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "apd_destroy": (C.c_int, [C.c_void_p]),
    "apd_last_error": (C.c_char_p, [C.c_void_p]),
    "apd_abi_version": (C.c_int, []),
    "apd_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_default_params": (C.c_int, [C.POINTER(ApdParams)]),
    "apd_set_params": (C.c_int, [C.c_void_p, C.POINTER(ApdParams)]),
    "apd_get_params": (C.c_int, [C.c_void_p, C.POINTER(ApdParams)]),
    "apd_set_source": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "apd_set_target": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "apd_swap_source_and_target": (C.c_int, [C.c_void_p]),
    "apd_clear_source": (C.c_int, [C.c_void_p]),
    "apd_clear_target": (C.c_int, [C.c_void_p]),
    "apd_align": (C.c_int, [C.c_void_p, _fp, C.POINTER(ApdResult)]),
    "apd_fitness": (C.c_int, [C.c_void_p, C.c_double, _dp]),
    "apd_transform_source": (C.c_int, [C.c_void_p, _fp, C.c_void_p, C.c_int, C.c_int]),
    "apd_linearize": (C.c_int, [C.c_void_p, _fp, _dp, _dp, _dp]),
    "apd_get_final_hessian": (C.c_int, [C.c_void_p, _dp]),
    "apd_linearize_d": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "apd_compute_error": (C.c_int, [C.c_void_p, _dp, _dp]),
    "apd_inlier_count": (C.c_int, [C.c_void_p, _fp, C.c_double, C.POINTER(C.c_int64)]),
    "apd_match_candidates": (C.c_int, [C.c_void_p, C.c_void_p, _ip, C.c_int, C.c_void_p, C.c_int, _fp, C.c_double, C.c_double, _ip, _fp, _dp, C.c_void_p]),
    "apd_compute_covariances": (C.c_int, [C.c_void_p]),
    "apd_get_knn": (C.c_int, [C.c_void_p, C.c_int, _ip]),
    "apd_get_covariances": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "apd_set_covariances": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int]),
    "apd_get_correspondences": (C.c_int, [C.c_void_p, _ip, _fp]),
    "apd_get_mahalanobis": (C.c_int, [C.c_void_p, _dp]),
    "apd_get_lm_trace": (C.c_int, [C.c_void_p, _dp, C.c_int, _ip]),
    "apd_cloudset_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _ip, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "apd_cloudset_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_cloudset_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_align_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _ip, _ip, _fp, C.c_int, C.c_void_p, C.c_int]),
    "apd_batch_align": (C.c_int, [C.c_void_p, C.c_void_p, _ip, C.c_void_p, _ip, C.c_int, _fp, C.c_int, C.c_void_p]),
    "apd_fitness_score": (C.c_int, [C.c_void_p, _fp, C.c_double, _dp, C.POINTER(C.c_int64)]),
    "apd_fitness_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _ip, _ip, _fp, C.c_int, C.c_double, _dp]),
    "apd_default_preprocess_params": (C.c_int, [C.POINTER(ApdPreprocessParams)]),
    "apd_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(ApdPreprocessParams), C.c_void_p, _ip]),
    "apd_cloudset_info": (C.c_int, [C.c_void_p, _ip, C.POINTER(C.c_int64)]),
    "apd_build_submap": (C.c_int, [C.c_void_p, C.c_void_p, _ip, C.c_int, _dp, C.c_double, C.c_uint64, C.c_void_p, C.c_int, _ip]),
    "apd_odometry_align": (C.c_int, [C.c_void_p, C.c_void_p, _ip, C.c_int, C.c_int, _fp, C.c_void_p]),
    "apd_synchronize": (C.c_int, [C.c_void_p]),
    "apd_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "apd_get_timeline": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, _ip]),
    "apd_get_debug_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "apd_get_kernel_times": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_int64)]),
    "apd_bench_streaming": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp]),
    "apd_get_search_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "apd_get_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "apd_get_work_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


def load_library(path: str | None = None):
    """dlopen the product library and bind every C-ABI symbol. Raises if the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -m riv_slam_b200.build` "
                           "(the product is the CUDA library; there is no CPU fallback)")
    L = C.CDLL(p)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    if L.apd_abi_version() != 1:
        raise RuntimeError("ABI version mismatch")
    if path is None:
        _lib = L
    return L


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _cloud_arg(pts):
    """-> (keepalive, void pointer, stride_bytes, n, mem)"""
    if _is_torch(pts):
        import torch
        t = pts
        if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] < 3:
            raise ValueError("cloud tensors must be float32 of shape (n, >=3)")
        t = t.contiguous()
        mem = MEM_DEVICE if t.is_cuda else MEM_HOST
        return t, C.c_void_p(t.data_ptr()), t.shape[1] * 4, t.shape[0], mem
    a = np.ascontiguousarray(pts, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("clouds must have shape (n, >=3)")
    return a, C.c_void_p(a.ctypes.data), a.shape[1] * 4, a.shape[0], MEM_HOST


def _mat16(T):
    if T is None:
        return None, None
    a = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    return a, a.ctypes.data_as(_fp)


class Handle:
    """One apd_handle (one registration object bound to one GPU and stream)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.apd_create(int(device), C.byref(self.h))
        if rc != APD_OK:
            raise ApdError(rc, self.L.apd_last_error(None).decode())
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.apd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, ok=(APD_OK,)):
        if rc not in ok:
            raise ApdError(rc, self.L.apd_last_error(self.h).decode())
        return rc

    def set_option(self, name: str, value: float):
        self.check(self.L.apd_set_option(self.h, name.encode(), float(value)))

    def set_stream(self, cuda_stream: int | None):
        self.check(self.L.apd_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self.check(self.L.apd_synchronize(self.h))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        self.check(self.L.apd_get_launch_count(self.h, C.byref(n)))
        return n.value

    def timeline(self):
        """(phase, ns) stamps of the last align launch's first pair (option "timeline")."""
        n = C.c_int32(0)
        buf = (C.c_uint64 * 500)()
        self.check(self.L.apd_get_timeline(self.h, buf, 250, C.byref(n)))
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)]

    def kernel_times(self):
        """({kind: ms}, {kind: launches}) since the last call (option "kernel_timing"): pack, build, knn_cov, align."""
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        self.check(self.L.apd_get_kernel_times(self.h, ms, n))
        names = ("pack_points", "build", "knn_cov", "align")
        return {k: float(v) for k, v in zip(names, ms)}, {k: int(v) for k, v in zip(names, n)}

    def bench_streaming(self, n_points: int, reps: int = 5):
        """{kernel: (GB/s algorithmic, ms)} of the streaming kernels at n_points (apd_bench_streaming)."""
        g = (C.c_double * 4)()
        m = (C.c_double * 4)()
        self.check(self.L.apd_bench_streaming(self.h, int(n_points), int(reps), g, m))
        names = ("pack_points", "transform_points", "cov_export", "cov_import")
        return {k: (float(a), float(b)) for k, a, b in zip(names, g, m)}

    def search_counters(self):
        """(kNN distance evaluations, 1-NN distance evaluations) since the last call."""
        a, b = C.c_int64(0), C.c_int64(0)
        self.check(self.L.apd_get_search_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_counters(self):
        buf = (C.c_uint64 * 16)()
        self.check(self.L.apd_get_debug_counters(self.h, buf))
        return [int(x) for x in buf]

    def work_counters(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.check(self.L.apd_get_work_counters(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def get_params(self) -> ApdParams:
        p = ApdParams()
        self.check(self.L.apd_get_params(self.h, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = self.get_params()
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        self.check(self.L.apd_set_params(self.h, C.byref(p)))


class FastAPDGICP:
    """Drop-in for the reference registration object (see module docstring).

    >>> reg = FastAPDGICP()
    >>> reg.setNumThreads(0); reg.setMaxCorrespondenceDistance(2.0); reg.setCorrespondenceRandomness(20)
    >>> reg.setInputTarget(target); reg.setInputSource(source)
    >>> aligned = reg.align()                       # pcl::Registration::align(output[, guess])
    >>> reg.hasConverged(), reg.getFinalTransformation(), reg.getFitnessScore()
    """

    def __init__(self, device: int = 0):
        self._H = Handle(device)
        self.L = self._H.L
        self.h = self._H.h
        self._result = None
        self._src = self._tgt = None
        self._n_src = self._n_tgt = 0

    # ---- fast_apdgicp.hpp:51-57 ----
    def setNumThreads(self, n: int):
        self._H.set_params(num_threads=int(n))  # accepted for source compatibility; the GPU path ignores it

    def setCorrespondenceRandomness(self, k: int):
        self._H.set_params(k_correspondences=int(k))

    def setRegularizationMethod(self, method: int):
        self._H.set_params(regularization=int(method))

    def setAzimuthVar(self, v: float):
        self._H.set_params(azimuth_var=float(v))

    def setElevationVar(self, v: float):
        self._H.set_params(elevation_var=float(v))

    def setDistVar(self, v: float):
        self._H.set_params(dist_var=float(v))

    # ---- pcl::Registration setters the factory uses (registrations.cpp:38-50) ----
    def setMaxCorrespondenceDistance(self, d: float):
        self._H.set_params(max_corr_dist=float(d))

    def setTransformationEpsilon(self, e: float):
        self._H.set_params(transformation_epsilon=float(e))

    def setMaximumIterations(self, n: int):
        self._H.set_params(max_iterations=int(n))

    # ---- lsq_registration.hpp:51-53 ----
    def setRotationEpsilon(self, e: float):
        self._H.set_params(rotation_epsilon=float(e))

    def setInitialLambdaFactor(self, f: float):
        self._H.set_params(lm_init_lambda_factor=float(f))

    def setDebugPrint(self, _on: bool):
        pass  # the LM table is always recorded; read it with getLMTrace()

    def setOptimizer(self, opt: int):
        """lsq_optimizer_type_ (lsq_registration_impl.hpp:17); the reference has no public setter."""
        self._H.set_params(optimizer=int(opt))

    def setLMMaxIterations(self, n: int):
        self._H.set_params(lm_max_iterations=int(n))

    # ---- clouds: fast_apdgicp_impl.hpp:68-108 ----
    def setInputSource(self, cloud, cache_key: int | None = None):
        keep, ptr, stride, n, mem = _cloud_arg(cloud)
        key = (id(cloud) if cache_key is None else cache_key) & 0xFFFFFFFFFFFFFFFF
        self._H.check(self.L.apd_set_source(self.h, ptr, stride, n, C.c_uint64(key), mem))
        self._src, self._n_src = cloud, n  # the reference keeps the shared pointer alive; so do we

    def setInputTarget(self, cloud, cache_key: int | None = None):
        keep, ptr, stride, n, mem = _cloud_arg(cloud)
        key = (id(cloud) if cache_key is None else cache_key) & 0xFFFFFFFFFFFFFFFF
        self._H.check(self.L.apd_set_target(self.h, ptr, stride, n, C.c_uint64(key), mem))
        self._tgt, self._n_tgt = cloud, n

    def swapSourceAndTarget(self):
        self._H.check(self.L.apd_swap_source_and_target(self.h))
        self._src, self._tgt = self._tgt, self._src
        self._n_src, self._n_tgt = self._n_tgt, self._n_src

    def clearSource(self):
        self._H.check(self.L.apd_clear_source(self.h))
        self._src, self._n_src = None, 0

    def clearTarget(self):
        self._H.check(self.L.apd_clear_target(self.h))
        self._tgt, self._n_tgt = None, 0

    # ---- pcl::Registration::align and its observers ----
    def align(self, guess=None, want_output: bool = True):
        """align(output, guess): returns the transformed source (n, 3) float32, or None.

        Like pcl::Registration::align this never raises for missing inputs or a failed LM step:
        it leaves hasConverged() false (SURVEY.md §8b error conventions).
        """
        keep, g = _mat16(guess)
        r = ApdResult()
        rc = self.L.apd_align(self.h, g, C.byref(r))
        self._result = r
        self._status = rc
        if rc in (APD_ERR_NO_INPUT, APD_ERR_TOO_FEW_POINTS):
            return None
        self._H.check(rc)
        if not want_output:
            return None
        return self.transformSource(None)

    def hasConverged(self) -> bool:
        return bool(self._result.converged) if self._result is not None else False

    def getFinalTransformation(self) -> np.ndarray:
        if self._result is None:
            return np.eye(4, dtype=np.float32)
        return np.array(self._result.T, dtype=np.float32).reshape(4, 4)

    def getFitnessScore(self, max_range: float = DBL_MAX) -> float:
        if self._result is not None and max_range == DBL_MAX and self._status == APD_OK:
            return float(self._result.fitness)  # computed inside the align kernel
        s = C.c_double(0)
        self._H.check(self.L.apd_fitness(self.h, float(max_range), C.byref(s)))
        return s.value

    def nr_iterations(self) -> int:
        return int(self._result.iterations) if self._result is not None else 0

    def status(self) -> int:
        return int(self._result.status) if self._result is not None else APD_OK

    def result(self) -> ApdResult:
        return self._result

    def transformSource(self, T=None) -> np.ndarray:
        keep, t = _mat16(T)
        out = np.empty((self._n_src, 3), dtype=np.float32)
        self._H.check(self.L.apd_transform_source(self.h, t, C.c_void_p(out.ctypes.data), 12, MEM_HOST))
        return out

    # ---- lsq_registration.hpp:54-57 ----
    def evaluateCost(self, pose, want_H: bool = True):
        keep, g = _mat16(pose)
        H = np.zeros(36)
        b = np.zeros(6)
        e = C.c_double(0)
        self._H.check(self.L.apd_linearize(self.h, g, H.ctypes.data_as(_dp), b.ctypes.data_as(_dp), C.byref(e)))
        return (e.value, H.reshape(6, 6), b) if want_H else e.value

    # ---- the protected hooks of the reference class (fast_apdgicp.hpp:77-83), at a double pose ----
    def linearize(self, trans, want_H: bool = True):
        """FastAPDGICP::linearize(Isometry3d, H, b) (fast_apdgicp_impl.hpp:198-272)."""
        a = np.ascontiguousarray(trans, dtype=np.float64).reshape(16)
        H = np.zeros(36)
        b = np.zeros(6)
        e = C.c_double(0)
        self._H.check(self.L.apd_linearize_d(self.h, a.ctypes.data_as(_dp), H.ctypes.data_as(_dp) if want_H else None, b.ctypes.data_as(_dp) if want_H else None, C.byref(e)))
        return (e.value, H.reshape(6, 6), b) if want_H else e.value

    def update_correspondences(self, trans):
        """FastAPDGICP::update_correspondences(Isometry3d) (fast_apdgicp_impl.hpp:133-194): the first half of linearize."""
        self.linearize(trans, want_H=False)

    def compute_error(self, trans) -> float:
        """FastAPDGICP::compute_error(Isometry3d) (fast_apdgicp_impl.hpp:275-298): stale correspondences of the last linearize."""
        a = np.ascontiguousarray(trans, dtype=np.float64).reshape(16)
        e = C.c_double(0)
        self._H.check(self.L.apd_compute_error(self.h, a.ctypes.data_as(_dp), C.byref(e)))
        return e.value

    def inlierCount(self, max_dist: float = 0.5, T=None) -> int:
        """publish_scan_matching_status (scan_matching_odometry_nodelet.cpp:698-712): aligned points with a target point closer than max_dist."""
        keep, t = _mat16(T)
        n = C.c_int64(0)
        self._H.check(self.L.apd_inlier_count(self.h, t, float(max_dist), C.byref(n)))
        return n.value

    def getFinalHessian(self) -> np.ndarray:
        H = np.zeros(36)
        self._H.check(self.L.apd_get_final_hessian(self.h, H.ctypes.data_as(_dp)))
        return H.reshape(6, 6)

    # ---- covariances: fast_apdgicp.hpp:64-74 ----
    def computeCovariances(self):
        return self.L.apd_compute_covariances(self.h)

    def _cov(self, which):
        n = self._n_tgt if which else self._n_src
        out = np.zeros((n, 4, 4))
        self._H.check(self.L.apd_get_covariances(self.h, which, out.ctypes.data_as(_dp)))
        return out

    def getSourceCovariances(self):
        return self._cov(0)

    def getTargetCovariances(self):
        return self._cov(1)

    def _set_cov(self, which, covs):
        a = np.ascontiguousarray(covs, dtype=np.float64)
        if a.shape[1:] == (3, 3):
            b = np.zeros((a.shape[0], 4, 4))
            b[:, :3, :3] = a
            a = b
        self._H.check(self.L.apd_set_covariances(self.h, which, a.ctypes.data_as(_dp), a.shape[0]))

    def setSourceCovariances(self, covs):
        self._set_cov(0, covs)

    def setTargetCovariances(self, covs):
        self._set_cov(1, covs)

    # ---- parity / debug getters ----
    def getKnn(self, which: int) -> np.ndarray:
        n = self._n_tgt if which else self._n_src
        k = self._H.get_params().k_correspondences
        out = np.zeros((n, k), dtype=np.int32)
        self._H.check(self.L.apd_get_knn(self.h, which, out.ctypes.data_as(_ip)))
        return out

    def getCorrespondences(self):
        corr = np.zeros(self._n_src, dtype=np.int32)
        sq = np.zeros(self._n_src, dtype=np.float32)
        self._H.check(self.L.apd_get_correspondences(self.h, corr.ctypes.data_as(_ip), sq.ctypes.data_as(_fp)))
        return corr, sq

    def getMahalanobis(self):
        out = np.zeros((self._n_src, 4, 4))
        self._H.check(self.L.apd_get_mahalanobis(self.h, out.ctypes.data_as(_dp)))
        return out

    def getLMTrace(self) -> np.ndarray:
        n = C.c_int32(0)
        self._H.check(self.L.apd_get_lm_trace(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 8))
        if n.value:
            self._H.check(self.L.apd_get_lm_trace(self.h, out.ctypes.data_as(_dp), n.value, C.byref(n)))
        return out

    def setOption(self, name, value):
        self._H.set_option(name, value)

    def handle(self) -> Handle:
        return self._H


def select_registration_method(params: dict | None = None, device: int = 0) -> FastAPDGICP:
    """The FAST_APDGICP branch of select_registration_method
    (radar_graph_slam/src/radar_graph_slam/registrations.cpp:38-50): rosparam names, factory defaults."""
    p = dict(params or {})
    reg = FastAPDGICP(device)
    reg.setNumThreads(int(p.get("reg_num_threads", 0)))
    reg.setTransformationEpsilon(float(p.get("reg_transformation_epsilon", 0.01)))
    reg.setMaximumIterations(int(p.get("reg_maximum_iterations", 64)))
    reg.setMaxCorrespondenceDistance(float(p.get("reg_max_correspondence_distance", 2.5)))
    reg.setCorrespondenceRandomness(int(p.get("reg_correspondence_randomness", 20)))
    reg.setDistVar(float(p.get("dist_var", 0.86)))
    reg.setAzimuthVar(float(p.get("azimuth_var", 0.5)))
    reg.setElevationVar(float(p.get("elevation_var", 1.0)))
    return reg


# ---------------------------------------------------------------------------------------------
# batched path (apd_cloudset_* / apd_align_pairs / apd_batch_align)
# ---------------------------------------------------------------------------------------------

def _ragged(clouds):
    """list of (n_i, c) arrays -> (concatenated (N, c) float32, offsets int32[n+1])"""
    if isinstance(clouds, tuple) and len(clouds) == 2:
        pts, off = clouds
        return pts, np.ascontiguousarray(off, dtype=np.int32)
    off = np.zeros(len(clouds) + 1, dtype=np.int32)
    off[1:] = np.cumsum([c.shape[0] for c in clouds])
    width = clouds[0].shape[1] if len(clouds) else 4
    pts = np.ascontiguousarray(np.concatenate(clouds, axis=0) if len(clouds) else np.zeros((0, width)), dtype=np.float32)
    return pts, off


class CloudSet:
    """A ragged batch of clouds resident in HBM (apd_cloudset)."""

    def __init__(self, handle: Handle, clouds):
        self.H = handle
        pts, off = _ragged(clouds)
        keep, ptr, stride, n, mem = _cloud_arg(pts)
        self.n_clouds = len(off) - 1
        self.offsets = off
        self.cs = C.c_void_p()
        handle.check(handle.L.apd_cloudset_create(handle.h, ptr, stride, off.ctypes.data_as(_ip), self.n_clouds, mem, C.byref(self.cs)))

    def prepare(self):
        self.H.check(self.H.L.apd_cloudset_prepare(self.H.h, self.cs))

    def destroy(self):
        if self.cs:
            self.H.L.apd_cloudset_destroy(self.H.h, self.cs)
            self.cs = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def align_pairs(handle: Handle, src: CloudSet, tgt: CloudSet, src_idx=None, tgt_idx=None, guesses=None, n_pairs=None, out=None):
    """apd_align_pairs: returns a structured array (RESULT_DTYPE); ``out`` may be a CUDA uint8 tensor of 96*n bytes."""
    si = None if src_idx is None else np.ascontiguousarray(src_idx, dtype=np.int32)
    ti = None if tgt_idx is None else np.ascontiguousarray(tgt_idx, dtype=np.int32)
    if n_pairs is None:
        n_pairs = len(si) if si is not None else (len(ti) if ti is not None else min(src.n_clouds, tgt.n_clouds))
    g = None if guesses is None else np.ascontiguousarray(guesses, dtype=np.float32).reshape(n_pairs, 16)
    if out is not None and _is_torch(out):
        handle.check(handle.L.apd_align_pairs(handle.h, src.cs, tgt.cs, None if si is None else si.ctypes.data_as(_ip),
                                              None if ti is None else ti.ctypes.data_as(_ip), None if g is None else g.ctypes.data_as(_fp),
                                              n_pairs, C.c_void_p(out.data_ptr()), MEM_DEVICE if out.is_cuda else MEM_HOST))
        if not out.is_cuda:
            handle.synchronize()
        return out
    res = np.zeros(n_pairs, dtype=RESULT_DTYPE) if out is None else out
    handle.check(handle.L.apd_align_pairs(handle.h, src.cs, tgt.cs, None if si is None else si.ctypes.data_as(_ip),
                                          None if ti is None else ti.ctypes.data_as(_ip), None if g is None else g.ctypes.data_as(_fp),
                                          n_pairs, C.c_void_p(res.ctypes.data), MEM_HOST))
    return res


def batch_align(handle: Handle, sources, targets, guesses=None):
    """apd_batch_align: host clouds in, host results out; pair i = (sources[i] -> targets[i])."""
    ps, os_ = _ragged(sources)
    pt, ot = _ragged(targets)
    n = len(os_) - 1
    if len(ot) - 1 != n:
        raise ValueError("sources and targets must hold the same number of clouds")
    if _is_torch(ps) or _is_torch(pt):
        raise ValueError("batch_align takes host arrays; use CloudSet + align_pairs for device tensors")
    if ps.shape[1] != pt.shape[1]:
        raise ValueError("sources and targets must share one point stride")
    g = None if guesses is None else np.ascontiguousarray(guesses, dtype=np.float32).reshape(n, 16)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    handle.check(handle.L.apd_batch_align(handle.h, C.c_void_p(ps.ctypes.data), os_.ctypes.data_as(_ip), C.c_void_p(pt.ctypes.data),
                                          ot.ctypes.data_as(_ip), ps.shape[1] * 4, None if g is None else g.ctypes.data_as(_fp), n,
                                          C.c_void_p(res.ctypes.data)))
    return res


def odometry_align(handle: Handle, scans, guesses=None, out=None):
    """apd_odometry_align: pair i registers scan i+1 onto scan i. ``scans`` is a list of (n_i, c) host arrays
    or a (points, offsets) tuple (points may be a pinned torch tensor); returns RESULT_DTYPE[n_scans - 1]."""
    pts, off = _ragged(scans)
    n = len(off) - 1
    keep, ptr, stride, _, mem = _cloud_arg(pts)
    if mem != MEM_HOST:
        raise ValueError("odometry_align takes host memory; use CloudSet + align_pairs for device tensors")
    g = None if guesses is None else np.ascontiguousarray(guesses, dtype=np.float32).reshape(max(n - 1, 0), 16)
    res = np.zeros(max(n - 1, 0), dtype=RESULT_DTYPE) if out is None else out
    optr = C.c_void_p(res.data_ptr()) if _is_torch(res) else C.c_void_p(res.ctypes.data)
    handle.check(handle.L.apd_odometry_align(handle.h, ptr, off.ctypes.data_as(_ip), n, stride, None if g is None else g.ctypes.data_as(_fp), optr))
    return res


def fitness_pairs(handle: Handle, src: CloudSet, tgt: CloudSet, src_idx=None, tgt_idx=None, poses=None, max_range: float = DBL_MAX, n_pairs=None):
    """apd_fitness_pairs: batched calc_fitness_score; returns float64[n_pairs]."""
    si = None if src_idx is None else np.ascontiguousarray(src_idx, dtype=np.int32)
    ti = None if tgt_idx is None else np.ascontiguousarray(tgt_idx, dtype=np.int32)
    if n_pairs is None:
        n_pairs = len(si) if si is not None else (len(ti) if ti is not None else min(src.n_clouds, tgt.n_clouds))
    g = None if poses is None else np.ascontiguousarray(poses, dtype=np.float32).reshape(n_pairs, 16)
    out = np.zeros(n_pairs)
    handle.check(handle.L.apd_fitness_pairs(handle.h, src.cs, tgt.cs, None if si is None else si.ctypes.data_as(_ip), None if ti is None else ti.ctypes.data_as(_ip),
                                            None if g is None else g.ctypes.data_as(_fp), n_pairs, float(max_range), out.ctypes.data_as(_dp)))
    return out


# ---------------------------------------------------------------------------------------------
# "next" rows SURVEY.md 8(f)-4 / 8(f)-2: the filters in front of the matcher and the submap behind it
# ---------------------------------------------------------------------------------------------

def preprocess_params(**kw) -> ApdPreprocessParams:
    """Code defaults of PreprocessingNodelet::initializeParams (preprocessing_nodelet.cpp:137-205) with keyword overrides."""
    p = ApdPreprocessParams()
    load_library().apd_default_preprocess_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(f"unknown preprocessing parameter {k}")
        setattr(p, k, v)
    return p


DOWNSAMPLE_METHODS = {"VOXELGRID": 0, "APPROX_VOXELGRID": 1}


def set_downsample_method(handle: Handle, method: str):
    """The ``downsample_method`` rosparam (preprocessing_nodelet.cpp:137-156, scan_matching_odometry_nodelet.cpp:149-167): VOXELGRID
    (pcl::VoxelGrid, the default and the launch file's choice) or APPROX_VOXELGRID (pcl::ApproximateVoxelGrid); NONE is
    ``downsample_resolution <= 0``. Anything else is refused (the nodelet falls back to a pass-through with a warning)."""
    if method not in DOWNSAMPLE_METHODS:
        raise ApdError(APD_ERR_UNSUPPORTED, f"unknown downsample_method {method!r}: VOXELGRID or APPROX_VOXELGRID (NONE = downsample_resolution <= 0)")
    handle.set_option("downsample_method", DOWNSAMPLE_METHODS[method])


def preprocess(handle: Handle, cloud, params: ApdPreprocessParams | None = None, downsample_method: str | None = None, **kw) -> np.ndarray:
    """distance_filter -> downsample -> outlier_removal (preprocessing_nodelet.cpp:812-815) of one host cloud on the GPU.

    ``cloud`` is (n, 4) packed x y z intensity or (n, 8) pcl::PointXYZI memory; the result has the same row layout.
    ``downsample_method`` (sticky on the handle): see set_downsample_method."""
    if downsample_method is not None:
        set_downsample_method(handle, downsample_method)
    a = np.ascontiguousarray(cloud, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("preprocess takes (n, 4) xyzi or (n, 8) pcl::PointXYZI arrays")
    p = params if params is not None else preprocess_params(**kw)
    out = np.zeros_like(a)
    n_out = C.c_int32(0)
    handle.check(handle.L.apd_preprocess(handle.h, C.c_void_p(a.ctypes.data), a.shape[1] * 4, 12 if a.shape[1] == 4 else 16, a.shape[0], C.byref(p),
                                         C.c_void_p(out.ctypes.data), C.byref(n_out)))
    return out[:n_out.value]


def build_submap(handle: Handle, keyframes: CloudSet, which, rel_poses, downsample_resolution: float = 0.1, cache_key: int = 0, want_cloud: bool = True,
                 downsample_method: str | None = None):
    """apd_build_submap: the keyframe clouds ``which`` moved by ``rel_poses`` (4x4 double each), concatenated and voxel-filtered,
    become the handle's target (scan_matching_odometry_nodelet.cpp:606-616). Returns the submap as (m, 4) xyzi (or its size).
    ``downsample_method`` (sticky on the handle): see set_downsample_method."""
    if downsample_method is not None:
        set_downsample_method(handle, downsample_method)
    w = np.ascontiguousarray(which, dtype=np.int32)
    P = np.ascontiguousarray(rel_poses, dtype=np.float64).reshape(len(w), 16)
    cap = int(sum(keyframes.offsets[i + 1] - keyframes.offsets[i] for i in w))
    out = np.zeros((max(cap, 1), 4), dtype=np.float32) if want_cloud else None
    n_out = C.c_int32(0)
    handle.check(handle.L.apd_build_submap(handle.h, keyframes.cs, w.ctypes.data_as(_ip), len(w), P.ctypes.data_as(_dp), float(downsample_resolution),
                                           C.c_uint64(cache_key), None if out is None else C.c_void_p(out.ctypes.data), cap if want_cloud else 0, C.byref(n_out)))
    return out[:n_out.value] if want_cloud else n_out.value
