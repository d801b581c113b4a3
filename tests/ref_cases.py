"""Cases of tests/golden/apd_ref_golden_v1.npz: vectors produced by the REFERENCE'S OWN FastAPDGICP sources (compiled unmodified
over stand-in Eigen / PCL / Boost headers, oracle/ref_apdgicp.cpp; one thread, correctly rounded atan2f - SURVEY 8c). Shared by
tests/golden/make_ref_golden.py (writes them, needs /root/reference), tests/test_ref_golden.py (the oracle against them, CPU)
and tests/test_gpu_ref_golden.py (the CUDA path against them through the C ABI, on the GPU box where /root/reference is absent)."""
import numpy as np

from conftest import LAUNCH_PARAMS, TIGHT_PARAMS

PAIR = dict(config=1, index=7, n_src=1200, n_tgt=1300)   # = conftest.small_pair

# calculate_covariances (fast_apdgicp_impl.hpp:300-363): regularisation x k, source cloud
COV_CASES = {"plane_k20": dict(regularization=3, k_correspondences=20), "plane_k10": dict(regularization=3, k_correspondences=10),
             "plane_k15": dict(regularization=3, k_correspondences=15), "none_k20": dict(regularization=0, k_correspondences=20),
             "min_eig_k20": dict(regularization=1, k_correspondences=20), "norm_min_eig_k20": dict(regularization=2, k_correspondences=20),
             "frobenius_k20": dict(regularization=4, k_correspondences=20)}


def poses():
    """Fixed double poses for linearize / compute_error (fast_apdgicp_impl.hpp:134-298)."""
    from scipy.spatial.transform import Rotation
    out = [np.eye(4)]
    P = np.eye(4); P[:3, 3] = [0.1, -0.05, 0.02]
    out.append(P)
    P = np.eye(4); P[:3, :3] = Rotation.from_rotvec([0.01, -0.02, 0.05]).as_matrix(); P[:3, 3] = [-0.4, 0.3, 0.05]
    out.append(P)
    return out


LIN_CASES = {"launch": LAUNCH_PARAMS, "defaults": {}, "gate_0.5": dict(LAUNCH_PARAMS, max_corr_dist=0.5)}

# whole registrations (computeTransformation, lsq_registration_impl.hpp:55-173); thresholds loose enough that no accept / reject
# decision sits in rounding noise (checked when the vectors are written)
ALIGN_CASES = {"launch": LAUNCH_PARAMS, "defaults": {}, "gn": dict(LAUNCH_PARAMS, optimizer=0),
               "eps_1e-4": dict(LAUNCH_PARAMS, transformation_epsilon=1e-4, rotation_epsilon=1e-5),
               "k10_min_eig": dict(LAUNCH_PARAMS, regularization=1, k_correspondences=10),
               "frobenius": dict(LAUNCH_PARAMS, regularization=4), "none": dict(LAUNCH_PARAMS, regularization=0),
               "lambda_1e-3": dict(LAUNCH_PARAMS, lm_init_lambda_factor=1e-3, transformation_epsilon=1e-4, rotation_epsilon=1e-5),
               "max_iter_2": dict(LAUNCH_PARAMS, max_iterations=2), "gate_0.5": dict(LAUNCH_PARAMS, max_corr_dist=0.5)}


def guess():
    from scipy.spatial.transform import Rotation
    G = np.eye(4); G[:3, :3] = Rotation.from_rotvec([0.0, 0.0, 0.03]).as_matrix(); G[:3, 3] = [0.3, -0.2, 0.0]
    return G.astype(np.float32)


def make_pair():
    from riv_slam_b200 import datagen
    return datagen.make_pair(PAIR["config"], PAIR["index"], n_src=PAIR["n_src"], n_tgt=PAIR["n_tgt"])


# BASELINE.json size: one C2 / C4 pair of 5000-point scans (= test_gpu_parity.pair5k) and a short odometry chain of 5000-point scans in
# which every scan is target once and source once (scan_matching_odometry_nodelet.cpp:449-468, 584-592)
PAIR5K = dict(config=4, index=0, n_src=5000)
CHAIN5K = dict(config=2, index=0, n_scans=5, n_points=5000)


def make_pair5k():
    from riv_slam_b200 import datagen
    return datagen.make_pair(PAIR5K["config"], PAIR5K["index"], n_src=PAIR5K["n_src"])


def make_chain5k():
    from riv_slam_b200 import datagen
    scans, _ = datagen.make_drive(CHAIN5K["config"], CHAIN5K["index"], CHAIN5K["n_scans"], CHAIN5K["n_points"], workers=4)
    return scans
