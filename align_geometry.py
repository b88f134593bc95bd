"""Drop-in for the reference's root module ``align_geometry.py`` (same public names,
positional order and defaults — SURVEY.md section 8b), with every per-pixel computation on
the B200 through libda3s.so.  ``main_align.py`` imports exactly these names.

Correspondences of ``align_two_point_clouds`` / ``_icp`` / ``_umeyama`` (module attribute
``CORRESPONDENCES``, environment variable ``DA3S_CORRESPONDENCES``):
  * ``"nearest"`` (default): the reference's own semantics — nearest-neighbour search
    (Open3D ICP, align_geometry.py:8-56; KD-tree Umeyama loop :84-140) — on the GPU
    (``da3s_icp_points``: uniform-grid exact nearest neighbour + float64 moments + closed form,
    one launch per iteration, no host round trip).  Open3D is not vendored: parity is against
    the restatement in oracle/ref_port.py (scipy cKDTree), see oracle/SPEC.md.
  * ``"pixel"``: every call site passes the two clouds of the SAME overlap frame in pixel order
    (main_align.py:37-44), so the pixel correspondences can be used directly — closed-form
    Umeyama (:59-82) over all finite pairs, or SE(3) for "icp" — in one fused pass instead of
    30 rounds of M nearest-neighbour queries.  Not what the reference computes.
Also different from the reference, on purpose:
  * the stubs ``align_two_point_clouds_irls`` / ``_turboreg`` (reference :143-159 return
    None) are implemented: Huber-IRLS with unit confidences.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

import os

from da3slam_b200 import _lib as _L
from da3slam_b200 import host as _host

CORRESPONDENCES = os.environ.get("DA3S_CORRESPONDENCES", "nearest")


def _nearest() -> bool:
    if CORRESPONDENCES not in ("nearest", "pixel"):
        raise ValueError(f"align_geometry.CORRESPONDENCES must be 'nearest' or 'pixel', got {CORRESPONDENCES!r}")
    return CORRESPONDENCES == "nearest"


def _finite_pairs(source, target):
    source = np.asarray(source)
    target = np.asarray(target)
    if source.shape != target.shape:
        raise ValueError("pixel-correspondence alignment needs clouds of equal shape "
                         f"(got {source.shape} and {target.shape})")
    keep = np.isfinite(source).all(axis=1) & np.isfinite(target).all(axis=1)   # align_geometry.py:94-95
    return source[keep], target[keep]


def _umeyama_sim3(X: np.ndarray, Y: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """Solve Y ~= s R X + t (align_geometry.py:59-82) on the GPU."""
    return _host.umeyama(X, Y, None, _L.UMEYAMA_MEAN)


def align_two_point_clouds_icp(source: np.ndarray, target: np.ndarray, threshold: float = 0.0001,
                               max_iterations: int = 50) -> Tuple[float, np.ndarray, np.ndarray]:
    """Rigid registration, s == 1.0 (align_geometry.py:8-56): point-to-point ICP from the identity with
    nearest-neighbour correspondences within `threshold` (or the pixel correspondences, module docstring)."""
    if _nearest():
        return _host.icp(source, target, threshold, max_iterations, rigid=True)
    src, tgt = _finite_pairs(source, target)
    s, R, t = _host.umeyama(src, tgt, None, _L.UMEYAMA_MEAN)
    # SE(3): keep the rotation, recompute t for unit scale (Kabsch); the rotation of the
    # similarity solution and of the rigid solution coincide (both are U D V^T of the same covariance)
    t = tgt.mean(axis=0) - R @ src.mean(axis=0)
    return 1.0, R, t


def align_two_point_clouds_umeyama(source: np.ndarray, target: np.ndarray, threshold: float = 0.001,
                                   max_iterations: int = 30) -> Tuple[float, np.ndarray, np.ndarray]:
    """target ~= s R source + t (align_geometry.py:84-140): up to `max_iterations` rounds of nearest neighbour
    (d^2 < threshold^2, fewer than 20 inliers stops) + closed-form Sim(3), composed."""
    if _nearest():
        return _host.icp(source, target, threshold, max_iterations, rigid=False)
    src, tgt = _finite_pairs(source, target)
    return _host.umeyama(src, tgt, None, _L.UMEYAMA_MEAN)


def align_two_point_clouds_irls(source: np.ndarray, target: np.ndarray, threshold: float = 0.001,
                                max_iterations: int = 30) -> Tuple[float, np.ndarray, np.ndarray]:
    """Huber-IRLS Umeyama with unit confidences (the reference stub, :152-159, returns None)."""
    src, tgt = _finite_pairs(source, target)
    ones = np.ones(len(src), np.float32)
    return _host.irls_pixel(tgt, src, ones, ones, min_points=3, max_iterations=max_iterations, compat="joint")


def align_two_point_clouds_turboreg(source: np.ndarray, target: np.ndarray, threshold: float = 0.001,
                                    max_iterations: int = 30) -> Tuple[float, np.ndarray, np.ndarray]:
    return align_two_point_clouds_irls(source, target, threshold, max_iterations)


def align_two_point_clouds(source: np.ndarray, target: np.ndarray, threshold: float = 0.001,
                           max_iterations: int = 30, method: str = "icp") -> Tuple[float, np.ndarray, np.ndarray]:
    """Register source onto target (align_geometry.py:162-187).  The reference ignores
    `method` and returns the Umeyama result (:182-183); so does this, and it prints it (:184-186)."""
    s, R, t = align_two_point_clouds_umeyama(source, target, threshold, max_iterations)
    print(f"s: {s}")
    print(f"R: {R}")
    print(f"t: {t}")
    return s, R, t


def depth_to_point_cloud_vectorized(depth, intrinsics, extrinsics, device=None, in_coords="camera"):
    """[N,H,W] depth -> [N,H,W,3] float32 points, camera or world frame
    (align_geometry.py:192-256).  numpy in -> numpy out, torch in -> torch out."""
    assert in_coords in ("camera", "world")
    import torch
    is_np = isinstance(depth, np.ndarray)
    pts = _host.unproject(depth.detach().cpu().numpy() if not is_np else depth,
                          intrinsics.detach().cpu().numpy() if not is_np else intrinsics,
                          extrinsics.detach().cpu().numpy() if not is_np else extrinsics,
                          world=(in_coords == "world"), out_f64=False)
    if is_np:
        return pts
    out = torch.from_numpy(pts)
    return out.to(device if device is not None else depth.device)


def extract_overlap_point_cloud(prev_chunk_prediction, cur_chunk_prediction) -> Tuple[np.ndarray, np.ndarray]:
    """Camera-frame clouds of prev's last frame and cur's first frame, each [1,H,W,3]
    (align_geometry.py:259-290)."""
    pm1 = depth_to_point_cloud_vectorized(prev_chunk_prediction.depth[-1:], prev_chunk_prediction.intrinsics[-1:],
                                          prev_chunk_prediction.extrinsics[-1:], in_coords="camera")
    pm2 = depth_to_point_cloud_vectorized(cur_chunk_prediction.depth[:1], cur_chunk_prediction.intrinsics[:1],
                                          cur_chunk_prediction.extrinsics[:1], in_coords="camera")
    print(f"point_map1: {pm1.shape}")
    print(f"point_map2: {pm1.shape}")
    return pm1, pm2


def images_to_chw01(images) -> np.ndarray:
    """(N,H,W,3) uint8 -> (N,3,H,W) in [0,1] (align_geometry.py:293-304)."""
    return images.transpose(0, 3, 1, 2) / 255.0


def estimate_depth_scale(prev_chunk, cur_chunk, conf_th=0.2, eps=1e-6) -> float:
    """median(d_prev / d_cur) over the valid, confident pixels of the overlap frame
    (align_geometry.py:307-330): exact selection on the GPU, bit-identical to np.median."""
    return _host.depth_scale(prev_chunk, cur_chunk, conf_th, eps, guarded=False)


def compute_aligned_chunk_extrinsics_from_prev_overlap(overlap_frame_global_extrinsics_for_next_align: np.ndarray,
                                                       cur_chunk_local_extrinsics: np.ndarray,
                                                       point_cloud_transform: np.ndarray) -> np.ndarray:
    """E0g = inv(T) E_prev;  Ei_g = Ei_l inv(E0_l) E0g  (align_geometry.py:333-368).
    N 4x4 float64 products: host arithmetic, as SURVEY.md 8a row E classifies it."""
    def to4(E):
        M = np.eye(4, dtype=np.float64)
        M[:3, :4] = E
        return M
    E0g = np.linalg.inv(point_cloud_transform) @ to4(overlap_frame_global_extrinsics_for_next_align)
    E0l_inv = np.linalg.inv(to4(cur_chunk_local_extrinsics[0]))
    return np.stack([(to4(E) @ E0l_inv @ E0g)[:3, :4] for E in cur_chunk_local_extrinsics], axis=0)


def make_image_chunks(image_paths: List[str], chunk_size: int, overlap: int = 1) -> List[List[str]]:
    """Chunks of chunk_size sharing `overlap` items, last start forced to n - chunk_size
    (align_geometry.py:371-392)."""
    assert chunk_size >= 2, "chunk_size must be at least 2"
    assert 0 <= overlap < chunk_size, "need 0 <= overlap < chunk_size"
    n = len(image_paths)
    if n < chunk_size:
        return []
    starts = list(range(0, n - chunk_size + 1, chunk_size - overlap))
    if starts[-1] != n - chunk_size:
        starts.append(n - chunk_size)
    return [image_paths[s:s + chunk_size] for s in starts]
