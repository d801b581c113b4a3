"""Batched loop-closure candidate verification (SURVEY.md §8(f) "next" row 3).

The reference verifies ONE Scan-Context candidate at a time (LoopDetector::performScanContextLoopClosure,
radar_graph_slam/src/radar_graph_slam/loop_detector.cpp:192-236); its all-candidates strategy —
align the new keyframe against every candidate, keep the best fitness score — is ``#if 0``-ed
(LoopDetector::matching, loop_detector.cpp:379-441) because it is too slow on the CPU. On the GPU all
candidates are one ``apd_align_pairs`` launch, so the strategy is restored here with the reference's
selection rule, thresholds and guess construction.
"""
from __future__ import annotations

import numpy as np

from .fast_apdgicp import CloudSet, Handle, align_pairs


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
    handle.set_option("fitness_max_range", fitness_score_max_range)
    try:
        src = CloudSet(handle, list(candidate_clouds))      # registration->setInputSource(candidate->cloud)
        tgt = CloudSet(handle, [new_keyframe_cloud])        # registration->setInputTarget(new_keyframe->cloud)
        res = align_pairs(handle, src, tgt, tgt_idx=np.zeros(n, dtype=np.int32), guesses=guesses)
    finally:
        handle.set_option("fitness_max_range", float(np.finfo(np.float64).max))
    best_score = float(np.finfo(np.float64).max)
    best = None
    for i in range(n):
        score = float(res[i]["fitness"])
        if not res[i]["converged"] or score > best_score:
            continue
        best_score = score
        best = i
    if best is None or best_score > fitness_score_thresh:
        return None, None, best_score, res
    return best, np.array(res[best]["T"]), best_score, res
