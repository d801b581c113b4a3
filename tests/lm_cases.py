"""Inputs that drive LsqRegistration::step_lm (reference
fast_apdgicp/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:127-173) through the branches an
ordinary registration never takes: rejected trials (rho < 0: lambda *= nu, nu *= 2, :160-163), a rejected
trial whose step is already below the convergence thresholds (returns true WITHOUT moving x0, :156-159) and
ten (or lm_max_iterations) rejections in a row ("lm not converged!!", converged_ stays false, :71-74, :172).

Well-conditioned clouds never reject: with the correspondences and Mahalanobis matrices frozen during the
trials (compute_error, fast_apdgicp_impl.hpp:275-298) the cost is a convex quadratic in the translation.
The cases below make it NON-convex through setTargetCovariances (fast_apdgicp_impl.hpp:116-118): a seeded
random subset of the target covariances is replaced by -c * I, so (C_B + R C_A R^T) is negative definite
for those points (eigenvalues -c + {1, 1, 1e-3}: far from singular, the arithmetic stays well conditioned)
and the APD measurement term is switched off (dist/azimuth/elevation var = 0) so that nothing re-adds a
positive part. Found by a seed search against the CPU oracle; every row of every trace has |rho| > 0.5, so
no accept/reject decision sits near its threshold.

Shared by tests/golden/make_golden_lm.py (writes the vectors), tests/test_golden.py (the oracle must keep
reproducing them) and tests/test_gpu_lm_branches.py (the CUDA path must match them and the live oracle).
"""
import numpy as np

LAUNCH = dict(k_correspondences=20, max_corr_dist=2.0, max_iterations=64, transformation_epsilon=0.1,
              rotation_epsilon=2e-3, dist_var=0.86, azimuth_var=1.0, elevation_var=1.0)
NO_APD = dict(dist_var=0.0, azimuth_var=0.0, elevation_var=0.0)
PAIR = dict(config=1, index=7, n_src=1200, n_tgt=1300)   # = conftest.small_pair

# name -> (mask seed, fraction of target points made negative definite, c, parameter overrides, what must happen)
CASES = {
    "rejected_then_accepted": (30, 0.60, 1.5, dict(lm_max_iterations=10), "rejected rows, inner index up to 6, converges"),
    "rejected_then_accepted_loose": (19, 0.65, 1.5, dict(lm_max_iterations=10, transformation_epsilon=0.6, rotation_epsilon=0.05), "6 rejections inside one outer iteration"),
    "rejected_but_converged": (8, 0.65, 1.5, dict(lm_max_iterations=10, transformation_epsilon=0.6, rotation_epsilon=0.05), "last trial rejected, step below eps: success, x0 not moved"),
    "lm_failed_2": (2, 0.70, 1.5, dict(lm_max_iterations=2), "two rejections = lm_max_iterations: lm not converged!!"),
    "lm_failed_1": (2, 0.70, 1.5, dict(lm_max_iterations=1), "one rejection = lm_max_iterations: lm not converged!!"),
    # no injection (fraction 0, APD model on): an ordinary registration, but with setInitialLambdaFactor far from its default
    "big_initial_lambda": (0, 0.0, 0.0, dict(lm_init_lambda_factor=1e-3, transformation_epsilon=1e-4, rotation_epsilon=1e-5), "setInitialLambdaFactor at a non-default value"),
}


def make_pair():
    from riv_slam_b200 import datagen
    return datagen.make_pair(PAIR["config"], PAIR["index"], n_src=PAIR["n_src"], n_tgt=PAIR["n_tgt"])


def case_params(name):
    p = dict(LAUNCH)
    if CASES[name][1] > 0:
        p.update(NO_APD)
    p.update(CASES[name][3])
    return p


def injected_target_covariances(name, cov_tgt):
    """cov_tgt: (n, 3, 3) regularised covariances of the target; returns the copy with the seeded subset replaced by -c*I."""
    seed, frac, c = CASES[name][:3]
    mask = np.random.default_rng(seed).random(cov_tgt.shape[0]) < frac
    out = np.array(cov_tgt, dtype=np.float64, copy=True)
    out[mask] = -c * np.eye(3)
    return out


def run_oracle(name, src, tgt):
    """The oracle's answer for a case: dict(T, converged, iterations, trace, lm_failed, final_hessian, fitness, cov_src, cov_tgt)."""
    from oracle.oracle import Oracle
    base = Oracle(**LAUNCH)
    base.set_source(src); base.set_target(tgt)
    assert base.compute_covariances() == 0
    cov_src = base.covariances(0).copy()
    cov_tgt = injected_target_covariances(name, base.covariances(1))
    o = Oracle(**case_params(name))
    o.set_source(src); o.set_target(tgt)
    o.set_covariances(0, cov_src); o.set_covariances(1, cov_tgt)
    rc, T, conv, it = o.align()
    assert rc == 0
    return dict(T=T, converged=conv, iterations=it, trace=o.trace(), lm_failed=o.lm_failed(), final_hessian=o.final_hessian(),
                fitness=o.fitness(), cov_src=cov_src, cov_tgt=cov_tgt)
