import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


LAUNCH_PARAMS = dict(  # radar_graph_slam/launch/radar_graph_slam.launch:34-36,95-101
    k_correspondences=20, max_corr_dist=2.0, max_iterations=64, transformation_epsilon=0.1,
    rotation_epsilon=2e-3, dist_var=0.86, azimuth_var=1.0, elevation_var=1.0)
TIGHT_PARAMS = dict(LAUNCH_PARAMS, transformation_epsilon=1e-6, rotation_epsilon=1e-6)


@pytest.fixture(scope="session")
def small_pair():
    from riv_slam_b200 import datagen
    return datagen.make_pair(1, 7, n_src=1200, n_tgt=1300)


@pytest.fixture(scope="session")
def c1_pair():
    from riv_slam_b200 import datagen
    return datagen.make_pair(1, 0, n_src=3000)
