"""``utils`` as a real package.  The reference ships BOTH a module ``utils.py`` (image
listing, key-frame stride, per-chunk debug colours) and a directory ``utils/`` without
``__init__.py``; the module shadows the directory, so ``solver.py``'s
``from utils.align_geometry_single import ...`` cannot be imported as published
(SURVEY.md section 0 item 6).  Here the directory is the package and re-exports the four
helpers of ``utils.py`` that ``main_align.py`` / ``solver.py`` import (utils.py:7,31,62,81).
These helpers are IO glue outside the alignment hot path and run on the host.
"""
from __future__ import annotations

import glob
import os
import re
from typing import List

import numpy as np

_IMAGE_PATTERNS = ("*.png", "*.jpg", "*.jpeg", "*.bmp", "*.tiff", "*.tif")
_PALETTE = ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0),
            (1.0, 0.0, 1.0), (0.0, 1.0, 1.0), (1.0, 0.5, 0.0), (0.5, 0.0, 1.0))
chunk_colors: list = []


def load_image(folder_path: str) -> List[str]:
    """Image paths of a folder ordered by the digits in their file names (utils.py:7-28)."""
    found: List[str] = []
    for pat in _IMAGE_PATTERNS:
        found += glob.glob(os.path.join(folder_path, pat))

    def number(path):
        digits = re.sub(r"\D", "", os.path.splitext(os.path.basename(path))[0])
        return int(digits) if digits else 0

    found.sort(key=number)
    if not found:
        print(f"Warning: No images found in {folder_path}")
        return []
    print(f"Found {len(found)} images in {folder_path}")
    return found


def extract_keyframe(image_paths: List[str], num_keyframe: int) -> List[str]:
    """Every num_keyframe-th path; a non-positive stride returns everything (utils.py:31-56)."""
    if not image_paths:
        return []
    if num_keyframe <= 0:
        print(f"Warning: num_keyframe must be positive, got {num_keyframe}")
        return image_paths
    picked = image_paths[::num_keyframe]
    print(f"Extracted {len(picked)} keyframes from {len(image_paths)} total frames (interval={num_keyframe})")
    return picked


def get_distinct_color(chunk_idx):
    """One of eight saturated colours, cycling (utils.py:62-78)."""
    return _PALETTE[chunk_idx % len(_PALETTE)]


def apply_chunk_color_to_images_batch(img_chw, chunk_idx):
    """Paint every image of a chunk in the chunk's debug colour; (B,3,H,W) in [0,1] out
    (utils.py:81-114; the reference's tail is cut off in the published file — the return
    value used by main_align.py:92-99 is the CHW float batch)."""
    while chunk_idx >= len(chunk_colors):
        chunk_colors.append(get_distinct_color(len(chunk_colors)))
    color = np.asarray(chunk_colors[chunk_idx], dtype=np.float64)
    try:
        import torch
        if isinstance(img_chw, torch.Tensor):
            img_chw = img_chw.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    arr = np.asarray(img_chw)
    if arr.ndim == 4 and arr.shape[1] == 3:
        b, _, h, w = arr.shape
    else:
        b, h, w = arr.shape[0], arr.shape[1], arr.shape[2]
    print(f"Chunk {chunk_idx} apply color: RGB{tuple(int(c * 255) for c in color)}")
    return np.broadcast_to(color.reshape(1, 3, 1, 1), (b, 3, h, w)).copy()
