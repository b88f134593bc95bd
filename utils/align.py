"""Drop-in for the reference's ``utils/align.py``: weighted Umeyama,
confidence x Huber IRLS on pixel correspondences, norm-ratio Umeyama, overlap extraction.
Same names, positional order, defaults and return types; the arithmetic runs on the B200."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from da3slam_b200 import _lib as _L
from da3slam_b200 import host as _host
from .geometry import depth_to_point_cloud_vectorized, apply_sim3_transform  # noqa: F401  (re-exported like the reference)


def weighted_umeyama_alignment(src, dst, w):
    """dst ~= s R src + t with weights w (utils/align.py:14-40; eps 1e-8 in both places)."""
    return _host.umeyama(src, dst, w, _L.UMEYAMA_WEIGHTED)


def weighted_umeyama_alignment0(points1: np.ndarray, points2: np.ndarray, weights: np.ndarray):
    """Legacy twin (utils/align.py:42-92), reproduced as the reference computes it: same centroids,
    covariance and rotation as `weighted_umeyama_alignment`, but scale = trace(Sigma) / (var + 1e-8) — wrong for
    rotated data, and what callers of this name get upstream."""
    return _host.umeyama(points1, points2, weights, _L.UMEYAMA_LEGACY_TRACE)


def huber_weight(residual: float, delta: float = 1.0) -> float:
    """rho'(r)/r of the Huber loss (utils/align.py:94-109).  Scalar helper; the per-point
    weights of the IRLS loop are computed inside the fused kernel."""
    a = abs(residual)
    return 1.0 if a <= delta else delta / a


def align_two_point_clouds_irls(point_map1: np.ndarray, point_map2: np.ndarray, conf1: np.ndarray, conf2: np.ndarray,
                                min_points: int = 100, max_iterations: int = 20,
                                convergence_threshold: float = 1e-6) -> Tuple[float, np.ndarray, np.ndarray]:
    """IRLS Sim(3) mapping point_map2 onto point_map1 (utils/align.py:111-218), including the
    reference's independent masks and its <= 5000-pair draw from the global numpy RNG."""
    return _host.irls_pixel(point_map1, point_map2, conf1, conf2, min_points, max_iterations, convergence_threshold,
                            delta=1.0, compat="reference")


def align_two_point_clouds_umeyama(point_map2: np.ndarray, point_map1: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """Norm-ratio scale + Kabsch on frame 0 (utils/align.py:224-276).  Parameter names are
    swapped in the reference (:224) and kept: the FIRST argument is the target."""
    target = np.asarray(point_map2)[0].reshape(-1, 3)
    source = np.asarray(point_map1)[0].reshape(-1, 3)
    return _host.umeyama(source, target, None, _L.UMEYAMA_NORMRATIO)


def align_two_point_clouds(point_map1: np.ndarray, point_map2: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """The reference dispatches to the norm-ratio variant (utils/align.py:301)."""
    return align_two_point_clouds_umeyama(point_map1, point_map2)


def extract_overlap_chunk_prediction(prev_chunk_prediction: dict, cur_chunk_prediction: dict, overlap_size: int):
    """World-frame point maps of prev's last and cur's first `overlap_size` frames; the
    reference returns None for both confidences (utils/align.py:307-343)."""
    pm1 = depth_to_point_cloud_vectorized(prev_chunk_prediction["depth"][-overlap_size:],
                                          prev_chunk_prediction["intrinsics"][-overlap_size:],
                                          prev_chunk_prediction["extrinsics"][-overlap_size:])
    pm2 = depth_to_point_cloud_vectorized(cur_chunk_prediction["depth"][:overlap_size],
                                          cur_chunk_prediction["intrinsics"][:overlap_size],
                                          cur_chunk_prediction["extrinsics"][:overlap_size])
    return pm1, pm2, None, None
