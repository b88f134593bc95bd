"""Drop-in for the reference's ``utils/geometry.py``: float64
world unprojection, Sim(3) application, Sim(3) chain accumulation, extrinsic re-basing.
The per-point work runs on the B200; the O(#chunks) 3x3 algebra stays on the host."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from da3slam_b200 import host as _host
from da3slam_b200.pipeline import accumulate_sim3 as _accumulate


def depth_to_point_cloud_vectorized(depth: np.ndarray, intrinsics: np.ndarray, extrinsics: np.ndarray) -> np.ndarray:
    """[N,H,W] depth, [N,3,3] K, [N,3,4] w2c -> [N,H,W,3] float64 world points
    (utils/geometry.py:4-40: K^-1 and the full 4x4 inverse, float64 throughout)."""
    return _host.unproject(depth, intrinsics, extrinsics, world=True, out_f64=True, mode="kinv", general_inverse=True)


def apply_sim3_transform(points: np.ndarray, s: float, R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """s * (p R^T) + t for [N,H,W,3] or [M,3]; float64 out (utils/geometry.py:43-70)."""
    return _host.apply_sim3(points, s, R, t)


def accumulate_sim3_transforms(sim3_transforms: List[Tuple[float, np.ndarray, np.ndarray]]):
    """Identity first, then left-to-right composition; len+1 entries (utils/geometry.py:73-119)."""
    return _accumulate(sim3_transforms)


def transform_camara_extrinsics(extrinsic: np.ndarray, s: float, R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """w2c of the current chunk re-expressed in the reference chunk:
    w2c_ref = w2c_cur @ [R^T / s | -R^T t / s]  (utils/geometry.py:121-160)."""
    w2c = np.eye(4)
    w2c[:3, :4] = extrinsic
    back = np.eye(4)
    back[:3, :3] = (1.0 / s) * R.T
    back[:3, 3] = -R.T @ t / s
    return (w2c @ back)[:3, :4]
