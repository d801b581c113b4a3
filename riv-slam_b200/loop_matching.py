"""Batched loop-closure candidate verification (SURVEY.md §8(f) "next" row 3).

The reference verifies ONE Scan-Context candidate at a time (LoopDetector::performScanContextLoopClosure,
radar_graph_slam/src/radar_graph_slam/loop_detector.cpp:192-236); its all-candidates strategy —
align the new keyframe against every candidate, keep the best fitness score — is ``#if 0``-ed
(LoopDetector::matching, loop_detector.cpp:379-441) because it is too slow on the CPU. On the GPU all
candidates are one ``apd_align_pairs`` launch, so the strategy is restored here with the reference's
selection rule, thresholds and guess construction.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .fast_apdgicp import RESULT_DTYPE, CloudSet, Handle, _fp, _ip


def relative_guess(new_keyframe_estimate: np.ndarray, candidate_estimate: np.ndarray) -> np.ndarray:
    """loop_detector.cpp:406-411: guess = (new^-1 * candidate).cast<float>() with guess(2,3) = 0
    (rotations re-normalised through a quaternion in the reference; inputs are expected orthonormal)."""
    g = (np.linalg.inv(np.asarray(new_keyframe_estimate, dtype=np.float64)) @ np.asarray(candidate_estimate, dtype=np.float64)).astype(np.float32)
    g[2, 3] = 0.0
    return g


def matching(handle: Handle, candidate_clouds, new_keyframe_cloud, guesses=None, fitness_score_max_range: float = float(np.finfo(np.float64).max),
             fitness_score_thresh: float = 0.5):
    """LoopDetector::matching over all candidates in one launch.

    Returns (best_index or None, relative_pose (4x4 float32) or None, best_score, records). Selection
    follows loop_detector.cpp:415-423: a candidate is skipped if it did not converge or its score is
    greater than the best so far (so the earliest of equal scores wins); the loop is rejected if the
    best score exceeds ``fitness_score_thresh`` (:431-434).
    """
    n = len(candidate_clouds)
    if n == 0:
        return None, None, float(np.finfo(np.float64).max), np.zeros(0)
    src = CloudSet(handle, list(candidate_clouds))      # registration->setInputSource(candidate->cloud)
    tgt = CloudSet(handle, [new_keyframe_cloud])        # registration->setInputTarget(new_keyframe->cloud)
    return match_candidates(handle, src, None, tgt, 0, guesses, fitness_score_max_range, fitness_score_thresh)


def match_candidates(handle: Handle, candidates: CloudSet, cand_idx, keyframes: CloudSet, keyframe_idx: int, guesses=None,
                     fitness_score_max_range: float = float(np.finfo(np.float64).max), fitness_score_thresh: float = 0.5):
    """apd_match_candidates on cloud sets that already live on the device (the C entry point a C++ LoopDetector links):
    selection and thresholds happen behind the C ABI. Same return value as ``matching``."""
    ci = None if cand_idx is None else np.ascontiguousarray(cand_idx, dtype=np.int32)
    n = candidates.n_clouds if ci is None else len(ci)
    g = None if guesses is None else np.ascontiguousarray(guesses, dtype=np.float32).reshape(n, 16)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    best = C.c_int32(-1)
    pose = np.zeros(16, dtype=np.float32)
    score = C.c_double(0)
    handle.check(handle.L.apd_match_candidates(handle.h, candidates.cs, None if ci is None else ci.ctypes.data_as(_ip), n, keyframes.cs, int(keyframe_idx),
                                               None if g is None else g.ctypes.data_as(_fp), float(fitness_score_max_range), float(fitness_score_thresh),
                                               C.byref(best), pose.ctypes.data_as(_fp), C.byref(score), C.c_void_p(res.ctypes.data)))
    if best.value < 0:
        return None, None, score.value, res
    return best.value, pose.reshape(4, 4), score.value, res
