"""TEST INFRASTRUCTURE — numpy/LAPACK twin of the C++ oracle, used only to test the oracle itself.

An independent second restatement of the same reference lines through different primitives
(LAPACK SVD instead of Jacobi, ``numpy.linalg.solve`` instead of LDL^T, scipy's rotation-vector
exponential instead of the quaternion formula, dense float32 distance matrices instead of a
kd-tree). Agreement between the two is the evidence that the oracle restates

  fast_apdgicp/include/fast_gicp/gicp/impl/fast_apdgicp_impl.hpp:121-363   (APD_I)
  fast_apdgicp/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:55-173 (LSQ_I)

correctly; it is O(n^2) in memory, so only for clouds of a few thousand points.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)


def sqdist_f32(q, p):
    """FLANN L2_Simple<float>: ((dx*dx + dy*dy) + dz*dz), one float rounding per operation."""
    q = np.asarray(q, dtype=np.float32)
    p = np.asarray(p, dtype=np.float32)
    d = q[:, None, :3] - p[None, :, :3]
    r = d[..., 0] * d[..., 0]
    r = r + d[..., 1] * d[..., 1]
    r = r + d[..., 2] * d[..., 2]
    return r


def knn(queries, cloud, k):
    """Exact kNN ordered by (d2, index). Returns (idx[nq,k], d2[nq,k])."""
    d2 = sqdist_f32(queries, cloud)
    n = cloud.shape[0]
    # stable argsort on d2 keeps the lower index first among equal distances
    order = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return order.astype(np.int32), np.take_along_axis(d2, order, axis=1)


def covariances(cloud, k=20, reg=REG_PLANE):
    """APD_I:303-363."""
    P = np.asarray(cloud, dtype=np.float32)[:, :3]
    idx, _ = knn(P, P, k)
    X = P[idx].astype(np.float64)  # (n,k,3)
    Xc = X - X.mean(axis=1, keepdims=True)
    C = np.einsum("nka,nkb->nab", Xc, Xc) / k
    if reg == REG_NONE:
        return C, idx
    if reg == REG_FROBENIUS:
        Ci = np.linalg.inv(C + 1e-3 * np.eye(3))
        Ci = Ci / np.linalg.norm(Ci, axis=(1, 2), keepdims=True)
        return np.linalg.inv(Ci), idx
    U, S, Vt = np.linalg.svd(C)
    if reg == REG_PLANE:
        vals = np.broadcast_to(np.array([1.0, 1.0, 1e-3]), S.shape)
    elif reg == REG_MIN_EIG:
        vals = np.maximum(S, 1e-3)
    elif reg == REG_NORMALIZED_MIN_EIG:
        vals = np.maximum(S / S.max(axis=1, keepdims=True), 1e-3)
    else:
        raise ValueError(reg)
    return np.einsum("nab,nb,nbc->nac", U, vals, Vt), idx


def transform_f32(T, P):
    """Float isometry times float point, ((R0*x + R1*y) + R2*z) + t (SURVEY.md §8c convention)."""
    Tf = np.asarray(T, dtype=np.float32)
    P = np.asarray(P, dtype=np.float32)
    out = np.empty((P.shape[0], 3), dtype=np.float32)
    for j in range(3):
        s = Tf[j, 0] * P[:, 0]
        s = s + Tf[j, 1] * P[:, 1]
        s = s + Tf[j, 2] * P[:, 2]
        out[:, j] = s + Tf[j, 3]
    return out


def _atan2_f32(y, x):
    return np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32).astype(np.float64)


def apd_cov(q, dist_var, az_var, el_var):
    """C_d of APD_I:167-184 for float32 points q (n,3)."""
    x, y, z = q[:, 0], q[:, 1], q[:, 2]
    qd = q.astype(np.float64)
    dist = np.sqrt(qd[:, 0] ** 2 + qd[:, 1] ** 2 + qd[:, 2] ** 2)
    aoa = _atan2_f32(x, np.sqrt(y * y + z * z))  # float32 arithmetic inside, like sqrtf
    s_x = dist * dist_var / 400
    s_y = dist * np.sin(az_var / 180 * np.pi) / np.cos(aoa)
    s_z = dist * np.sin(el_var / 180 * np.pi) / np.cos(aoa)
    el = _atan2_f32(np.sqrt(x * x + y * y), z)
    az = _atan2_f32(y, x)
    Rz = Rotation.from_rotvec(az[:, None] * np.array([0.0, 0.0, 1.0])).as_matrix()  # AngleAxis(az, Z)
    Ry = Rotation.from_rotvec(el[:, None] * np.array([0.0, 1.0, 0.0])).as_matrix()  # AngleAxis(el, Y)
    R = Rz @ Ry
    S2 = np.stack([s_x, s_y, s_z], axis=1) ** 2
    return np.einsum("nab,nb,ncb->nac", R, S2, R)


class PyRef:
    def __init__(self, k=20, reg=REG_PLANE, max_corr_dist=np.finfo(np.float32).max, max_iterations=64,
                 rotation_epsilon=2e-3, transformation_epsilon=5e-4, lm_max_iterations=10,
                 lm_init_lambda_factor=1e-9, dist_var=0.86, azimuth_var=0.5, elevation_var=1.0):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def set_source(self, P):
        self.src = np.asarray(P, dtype=np.float32)[:, :3]
        self.cov_src, self.knn_src = covariances(self.src, self.k, self.reg)

    def set_target(self, P):
        self.tgt = np.asarray(P, dtype=np.float32)[:, :3]
        self.cov_tgt, self.knn_tgt = covariances(self.tgt, self.k, self.reg)

    def update_correspondences(self, T):
        q = transform_f32(T, self.src)
        idx, d2 = knn(q, self.tgt, 1)
        idx, d2 = idx[:, 0], d2[:, 0]
        thr = float(self.max_corr_dist) * float(self.max_corr_dist)
        self.corr = np.where(d2.astype(np.float64) < thr, idx, -1).astype(np.int32)
        self.sq = d2
        Cd = apd_cov(q, self.dist_var, self.azimuth_var, self.elevation_var)
        R = np.asarray(T, dtype=np.float64)[:3, :3]
        j = np.where(self.corr >= 0, self.corr, 0)
        RCR = (self.cov_tgt[j] + Cd) + R @ (self.cov_src + Cd) @ R.T
        valid = self.corr >= 0
        self.M = np.zeros_like(RCR)
        self.M[valid] = np.linalg.inv(RCR[valid])

    def _residuals(self, T):
        T = np.asarray(T, dtype=np.float64)
        valid = self.corr >= 0
        a = self.src.astype(np.float64)[valid]
        b = self.tgt.astype(np.float64)[self.corr[valid]]
        p = a @ T[:3, :3].T + T[:3, 3]
        return p, b - p, self.M[valid]

    def linearize(self, T):
        self.update_correspondences(T)
        p, e, M = self._residuals(T)
        n = p.shape[0]
        J = np.zeros((n, 3, 6))
        J[:, 0, 1], J[:, 0, 2] = -p[:, 2], p[:, 1]
        J[:, 1, 0], J[:, 1, 2] = p[:, 2], -p[:, 0]
        J[:, 2, 0], J[:, 2, 1] = -p[:, 1], p[:, 0]
        J[:, :, 3:] = -np.eye(3)
        H = np.einsum("nar,nab,nbc->rc", J, M, J)
        g = np.einsum("nar,nab,nb->r", J, M, e)
        err = np.einsum("na,nab,nb->", e, M, e)
        return err, H, g

    def compute_error(self, T):
        _, e, M = self._residuals(T)
        return np.einsum("na,nab,nb->", e, M, e)

    def is_converged(self, delta):
        r = np.abs(delta[:3, :3] - np.eye(3)).max() / self.rotation_epsilon
        t = np.abs(delta[:3, 3]).max() / self.transformation_epsilon
        return max(r, t) < 1

    def align(self, guess=None):
        x = np.asarray(np.eye(4) if guess is None else guess, dtype=np.float32).astype(np.float64)
        lam = -1.0
        converged = False
        trace = []
        it = 0
        i = 0
        while i < self.max_iterations and not converged:
            it = i
            i += 1
            y0, H, g = self.linearize(x)
            if lam < 0:
                lam = self.lm_init_lambda_factor * np.abs(np.diag(H)).max()
            nu = 2.0
            ok = False
            for li in range(self.lm_max_iterations):
                d = np.linalg.solve(H + lam * np.eye(6), -g)
                delta = np.eye(4)
                delta[:3, :3] = Rotation.from_rotvec(d[:3]).as_matrix()
                delta[:3, 3] = d[3:]
                xi = delta @ x
                yi = self.compute_error(xi)
                rho = (y0 - yi) / d.dot(lam * d - g)
                trace.append((it, li, y0, yi, rho, lam, np.linalg.norm(d), float(not rho < 0)))
                if rho < 0:
                    if self.is_converged(delta):
                        ok = True
                        break
                    lam *= nu
                    nu *= 2
                    continue
                x = xi
                lam *= max(1.0 / 3.0, 1 - (2 * rho - 1) ** 3)
                ok = True
                break
            if not ok:
                break
            converged = self.is_converged(delta)
        self.final = x.astype(np.float32)
        self.converged = converged
        self.iterations = it
        self.trace = np.array(trace)
        return self.final, converged, it

    def fitness(self, max_range=np.finfo(np.float64).max):
        q = transform_f32(self.final, self.src)
        _, d2 = knn(q, self.tgt, 1)
        d2 = d2[:, 0].astype(np.float64)
        m = d2 <= max_range
        return d2[m].sum() / m.sum() if m.any() else np.finfo(np.float64).max
