"""Host-side mirror of ``radar_graph_slam::InformationMatrixCalculator`` (SURVEY.md §8(f) "next" row 1):
the fitness score — a 1-NN mean-squared-distance pass, the same primitive as getFitnessScore — runs on
the GPU through ``apd_fitness_score``; the edge information matrix is the reference's closed form on it.

Reference: radar_graph_slam/src/radar_graph_slam/information_matrix_calculator.cpp:14-86 and
radar_graph_slam/include/radar_graph_slam/information_matrix_calculator.hpp:27-56 (defaults of ``load``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .fast_apdgicp import DBL_MAX, FastAPDGICP, _dp, _mat16


class InformationMatrixCalculator:
    def __init__(self, params: dict | None = None, device: int = 0):
        self._reg = FastAPDGICP(device)
        self.load(params or {})

    def load(self, params: dict):
        """information_matrix_calculator.hpp:27-38 (same names and defaults)."""
        g = params.get
        self.use_const_inf_matrix = bool(g("use_const_inf_matrix", False))
        self.const_stddev_x = float(g("const_stddev_x", 0.5))
        self.const_stddev_q = float(g("const_stddev_q", 0.1))
        self.var_gain_a = float(g("var_gain_a", 20.0))
        self.min_stddev_x = float(g("min_stddev_x", 0.1))
        self.max_stddev_x = float(g("max_stddev_x", 5.0))
        self.min_stddev_q = float(g("min_stddev_q", 0.05))
        self.max_stddev_q = float(g("max_stddev_q", 0.2))
        # information_matrix_calculator.cpp:24 — the constructor both reference callers use (radar_graph_slam_nodelet.cpp, loop_detector.cpp)
        # reads 0.5; the 2.5 of the header's unused load() template (information_matrix_calculator.hpp:37) never takes effect
        self.fitness_score_thresh = float(g("fitness_score_thresh", 0.5))

    def calc_fitness_score(self, cloud1, cloud2, relpose, max_range: float = DBL_MAX) -> float:
        """information_matrix_calculator.cpp:55-86: kd-tree on cloud1, cloud2 transformed by relpose.cast<float>()."""
        r = self._reg
        r.setInputTarget(cloud1)
        r.setInputSource(cloud2)
        keep, t = _mat16(relpose)
        s = C.c_double(0)
        r._H.check(r.L.apd_fitness_score(r.h, t, float(max_range), C.byref(s), None))
        return s.value

    @staticmethod
    def _weight(a, max_x, min_y, max_y, x):  # information_matrix_calculator.hpp:43-46
        y = (1.0 - np.exp(-a * x)) / (1.0 - np.exp(-a * max_x))
        return min_y + (max_y - min_y) * y

    def information_from_fitness(self, fitness_score: float) -> np.ndarray:
        """information_matrix_calculator.cpp:39-52 (w_x, w_q are floats in the reference)."""
        w_x = np.float32(1.0e-8 * self._weight(self.var_gain_a, self.fitness_score_thresh, self.min_stddev_x ** 2, self.max_stddev_x ** 2, fitness_score))
        w_q = np.float32(1.0e-8 * self._weight(self.var_gain_a, self.fitness_score_thresh, self.min_stddev_q ** 2, self.max_stddev_q ** 2, fitness_score))
        inf = np.eye(6)
        inf[:3, :3] /= float(w_x)
        inf[3:, 3:] /= float(w_q)
        return inf

    def calc_information_matrix(self, cloud1, cloud2, relpose) -> np.ndarray:
        """information_matrix_calculator.cpp:29-53."""
        if self.use_const_inf_matrix:
            inf = np.eye(6)
            inf[:3, :3] /= self.const_stddev_x
            inf[3:, 3:] /= self.const_stddev_q
            return inf
        return self.information_from_fitness(self.calc_fitness_score(cloud1, cloud2, relpose))
