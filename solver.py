"""Drop-in for the reference's ``solver.py::SLAMSolver`` (same constructor and method names,
solver.py:22-246): a deque of frame paths, one network call per chunk, depth-scale + single
overlap-frame registration against the previous chunk, extrinsics chaining, viewer update.
Orchestration is host Python as in the reference; every per-pixel stage it calls
(``estimate_depth_scale``, ``get_aligned_chunk_extrinsics_single_overlap``, the viewer) runs
on the B200.  The network stays the reference's ``depth_anything_3`` model (out of scope)."""
from __future__ import annotations

import time
from collections import deque
from typing import Dict, List, Tuple

import numpy as np
import torch

from utils import extract_keyframe, load_image
from utils.align_geometry_single import (estimate_depth_scale, get_aligned_chunk_extrinsics_single_overlap,
                                         image_to_chw01)


def _say(msg: str) -> None:
    print(f"[solver] {msg}")


class SLAMSolver:
    def __init__(self, image_dir, config):
        self.config = config
        self.chunk_size = self.config["Model"]["chunk_size"]
        self.overlap_size = self.config["Model"]["overlap_size"]      # kept; alignment uses ONE overlap frame
        self.image_dir = image_dir
        self.chunk_count = 0
        self.frame_buffer: deque = deque(maxlen=self.chunk_size * 2)
        self.chunk_prediction_list: List[Dict] = []
        self.prev_overlap_aligned_3x4 = None      # global w2c of the previous chunk's last frame
        self.model = None
        self.load_model()
        self.viewer = None
        self.init_viewer()

    def load_model(self):
        """solver.py:49-67: cuda | mps | cpu, DepthAnything3.from_pretrained(Weights.DA3)."""
        if torch.cuda.is_available():
            self.device = "cuda"
        else:
            self.device = "mps" if torch.backends.mps.is_available() else "cpu"
        _say(f"network device: {self.device}")
        from depth_anything_3.api import DepthAnything3        # the reference's network, not part of this path
        weights = self.config["Weights"]["DA3"]
        self.model = DepthAnything3.from_pretrained(weights).to(self.device).eval()
        _say(f"DA3 weights loaded from {weights}")

    def init_viewer(self):
        """solver.py:69-78."""
        try:
            from viewer import SLAMViewer
        except ImportError as err:                              # viser missing: run headless
            _say(f"no viewer ({err})")
            self.viewer = None
            return
        self.viewer = SLAMViewer(port=self.config["Model"]["port"])

    def update_buffer_after_chunk_processed(self):
        """Drop chunk - overlap frames so the next chunk starts on the overlap frame (solver.py:80-85)."""
        if len(self.frame_buffer) <= self.overlap_size:
            return
        drop = min(self.chunk_size - self.overlap_size, len(self.frame_buffer))
        for _ in range(drop):
            self.frame_buffer.popleft()

    def update_viewer(self, chunk_prediction: Dict):
        """Every frame of the chunk (overlap frame included, as the reference does) with its global
        extrinsic (solver.py:87-114)."""
        if self.viewer is None:
            return
        poses = chunk_prediction.get("extrinsics_global")
        if poses is None:                                       # only legitimate for the first chunk
            _say("chunk has no global extrinsics yet: showing it in its local frame")
            poses = chunk_prediction["extrinsics"]
        for i, _path in enumerate(chunk_prediction["image_paths"]):
            self.viewer.add_frame(image_to_chw01(chunk_prediction, i), chunk_prediction["depth"][i], chunk_prediction["conf"][i],
                                  poses[i], chunk_prediction["intrinsics"][i])

    def process_chunk_alignment(self, prev_chunk_prediction: Dict, cur_chunk_prediction: Dict) -> Tuple[float, np.ndarray, np.ndarray]:
        """solver.py:116-153: depth scale (mutates cur depth), overlap registration, extrinsics chain."""
        s_depth = estimate_depth_scale(prev_chunk_prediction, cur_chunk_prediction, conf_th=0.2)
        cur_chunk_prediction["depth"] = cur_chunk_prediction["depth"] * s_depth
        extrinsics_global, prev_overlap_for_next, (s, R, t) = get_aligned_chunk_extrinsics_single_overlap(
            prev_overlap_aligned_3x4=self.prev_overlap_aligned_3x4,
            prev_chunk_prediction=prev_chunk_prediction,
            cur_chunk_prediction=cur_chunk_prediction)
        cur_chunk_prediction["extrinsics_global"] = extrinsics_global
        self.prev_overlap_aligned_3x4 = prev_overlap_for_next
        return s, R, t

    def run_single_chunk_prediction(self, chunk_image_paths: List[str]) -> Dict:
        """One network call; the Prediction fields the hot path consumes (solver.py:155-177)."""
        _say(f"chunk {self.chunk_count}: network inference on {len(chunk_image_paths)} frames")
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        with torch.no_grad():
            pred = self.model.inference(image=chunk_image_paths, process_res_method="upper_bound_resize")
        out = {name: getattr(pred, name) for name in ("processed_images", "depth", "conf", "extrinsics", "intrinsics")}
        out.update(chunk_idx=self.chunk_count, image_paths=chunk_image_paths)
        return out

    def load_chunk_image_paths(self) -> List[str]:
        return list(self.frame_buffer)[:self.chunk_size]

    def should_run_chunk_prediction(self) -> bool:
        return len(self.frame_buffer) >= self.chunk_size

    def process_frame(self, image_path: str):
        """solver.py:192-228."""
        self.frame_buffer.append(image_path)
        if not self.should_run_chunk_prediction():
            return
        cur = self.run_single_chunk_prediction(self.load_chunk_image_paths())
        self.chunk_prediction_list.append(cur)
        if self.chunk_count == 0:
            cur["extrinsics_global"] = cur["extrinsics"]          # first chunk defines the global frame
            self.prev_overlap_aligned_3x4 = cur["extrinsics_global"][-1]
        else:
            self.process_chunk_alignment(self.chunk_prediction_list[self.chunk_count - 1], cur)
        self.update_viewer(cur)
        self.update_buffer_after_chunk_processed()
        self.chunk_count += 1
        time.sleep(self.config["Model"]["sleep_between_chunk"])   # the reference pauses so the viewer can be watched

    def run(self):
        """solver.py:230-246."""
        frames = load_image(self.image_dir)
        if not frames:
            _say(f"nothing to do: {self.image_dir} holds no images")
            return
        for path in extract_keyframe(frames, self.config["Model"]["keyframe_interval"]):
            self.process_frame(path)
        _say(f"done: {self.chunk_count} chunks")
