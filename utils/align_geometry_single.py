"""Drop-in for the reference's ``utils/align_geometry_single.py`` (what ``solver.py``
imports): dict-or-object access, guarded depth scale, single-overlap-frame alignment and
the chunk extrinsics chain.  Per-pixel work on the B200."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from da3slam_b200 import _lib as _L
from da3slam_b200 import host as _host


def to4x4(E3x4: np.ndarray) -> np.ndarray:
    E = np.eye(4, dtype=np.float64)
    E[:3, :4] = E3x4
    return E


def to3x4(E4x4: np.ndarray) -> np.ndarray:
    return E4x4[:3, :4]


def _get(pred, key: str):
    """dict or attribute access (utils/align_geometry_single.py:18-22)."""
    return pred[key] if isinstance(pred, dict) else getattr(pred, key)


def image_to_chw01(pred, idx: int) -> np.ndarray:
    """processed_images[idx] (H,W,3) uint8 -> (3,H,W) in [0,1] (:25-28)."""
    return _get(pred, "processed_images")[idx].transpose(2, 0, 1) / 255.0


def estimate_depth_scale(prev_chunk, cur_chunk, conf_th=0.2, eps=1e-6) -> float:
    """Guarded median depth ratio of the overlap frame (:31-49): < 50 valid pixels or a
    non-finite / non-positive median give 1.0."""
    return _host.depth_scale(prev_chunk, cur_chunk, conf_th, eps, guarded=True)


def depth_to_point_cloud_vectorized(depth, intrinsics, extrinsics, device=None, in_coords="camera"):
    """Same contract as align_geometry.depth_to_point_cloud_vectorized (:52-102)."""
    import align_geometry as _ag
    return _ag.depth_to_point_cloud_vectorized(depth, intrinsics, extrinsics, device, in_coords)


def extract_single_overlap_point_cloud(prev_chunk_prediction, cur_chunk_prediction) -> Tuple[np.ndarray, np.ndarray]:
    """Camera-frame clouds of prev[-1] and cur[0] (:105-122)."""
    pc_prev = depth_to_point_cloud_vectorized(_get(prev_chunk_prediction, "depth")[-1:],
                                              _get(prev_chunk_prediction, "intrinsics")[-1:],
                                              _get(prev_chunk_prediction, "extrinsics")[-1:], in_coords="camera")
    pc_cur = depth_to_point_cloud_vectorized(_get(cur_chunk_prediction, "depth")[:1],
                                             _get(cur_chunk_prediction, "intrinsics")[:1],
                                             _get(cur_chunk_prediction, "extrinsics")[:1], in_coords="camera")
    return pc_prev, pc_cur


def align_two_point_clouds_icp(source: np.ndarray, target: np.ndarray, threshold: float, max_iterations: int,
                               verbose: bool = True) -> Tuple[float, np.ndarray, np.ndarray]:
    """target ~= R source + t, s == 1 (:126-180): point-to-point ICP; correspondences as selected by
    align_geometry.CORRESPONDENCES (nearest neighbour by default, as Open3D does)."""
    import align_geometry as _ag
    n_src, n_tgt = source.shape[0], target.shape[0]
    s, R, t = _ag.align_two_point_clouds_icp(source, target, threshold, max_iterations)
    if verbose:
        print(f"[ICP] threshold={threshold}, max_iter={max_iterations}")
        print(f"[ICP] source points: {n_src}, target points: {n_tgt}")
        print(f"[ICP] det(R)={np.linalg.det(R):.6f} (should be close to +1)")
        print(f"[ICP] t={t}")
    return s, R, t


def align_two_point_clouds(source: np.ndarray, target: np.ndarray, threshold: float,
                           max_iterations: int) -> Tuple[float, np.ndarray, np.ndarray]:
    return align_two_point_clouds_icp(source, target, threshold, max_iterations)


def get_aligned_chunk_extrinsics_single_overlap(prev_overlap_aligned_3x4: np.ndarray, prev_chunk_prediction,
                                                cur_chunk_prediction, icp_threshold: float = 0.1,
                                                icp_max_iter: int = 50):
    """Global w2c of every frame of the current chunk from the previous chunk's last global
    w2c and the overlap-frame registration (:192-255).  Returns (E_global [N,3,4] float64,
    E_last [3,4], (s, R, t))."""
    if prev_overlap_aligned_3x4 is None:
        raise ValueError("prev_overlap_aligned_3x4 is None. You must initialize it from the first chunk.")
    pc_prev, pc_cur = extract_single_overlap_point_cloud(prev_chunk_prediction, cur_chunk_prediction)
    s, R, t = align_two_point_clouds(pc_cur.reshape(-1, 3), pc_prev.reshape(-1, 3), threshold=icp_threshold,
                                     max_iterations=icp_max_iter)
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R
    T[:3, 3] = t
    Eg = np.linalg.inv(T) @ to4x4(prev_overlap_aligned_3x4)
    E_local = _get(cur_chunk_prediction, "extrinsics")
    n = E_local.shape[0]
    out = np.zeros((n, 3, 4), dtype=np.float64)
    out[0] = to3x4(Eg)
    for i in range(1, n):
        Eg = to4x4(E_local[i]) @ np.linalg.inv(to4x4(E_local[i - 1])) @ Eg
        out[i] = to3x4(Eg)
    return out, out[-1], (float(s), np.asarray(R, np.float64), np.asarray(t, np.float64))
