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
        self.device = "cuda" if torch.cuda.is_available() else "mps" if torch.backends.mps.is_available() else "cpu"
        print(f"Using device: {self.device}")
        try:
            from depth_anything_3.api import DepthAnything3
            model_path = self.config["Weights"]["DA3"]
            print(f"Loading DA3 model from {model_path}...")
            self.model = DepthAnything3.from_pretrained(model_path).to(self.device)
            self.model.eval()
            print("Model loaded successfully")
        except ImportError as e:
            print(f"Failed to load DA3 model: {e}")
            raise
        except Exception as e:
            print(f"Error loading model: {e}")
            raise

    def init_viewer(self):
        """solver.py:69-78."""
        port = self.config["Model"]["port"]
        try:
            from viewer import SLAMViewer
            self.viewer = SLAMViewer(port=port)
            print(f"Viewer initialized on port {port}")
        except ImportError as e:
            print(f"Failed to initialize viewer: {e}")
            self.viewer = None

    def update_buffer_after_chunk_processed(self):
        """Drop chunk - overlap frames so the next chunk starts on the overlap frame (solver.py:80-85)."""
        if len(self.frame_buffer) > self.overlap_size:
            for _ in range(self.chunk_size - self.overlap_size):
                if self.frame_buffer:
                    self.frame_buffer.popleft()

    def update_viewer(self, chunk_prediction: Dict):
        """Every frame of the chunk (overlap frame included, as the reference does) with its global
        extrinsic (solver.py:87-114)."""
        if self.viewer is None:
            return
        extrinsics_global = chunk_prediction.get("extrinsics_global", None)
        if extrinsics_global is None:
            print("warn: no extrinsics_global; if is not the first chunk then error")
            extrinsics_global = chunk_prediction["extrinsics"]
        for i in range(len(chunk_prediction["image_paths"])):
            self.viewer.add_frame(image=image_to_chw01(chunk_prediction, i), depth=chunk_prediction["depth"][i],
                                  conf=chunk_prediction["conf"][i], extrinsic=extrinsics_global[i],
                                  intrinsic=chunk_prediction["intrinsics"][i])

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
        print(f"  Predict single chunk with {len(chunk_image_paths)} images through da3...")
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        with torch.no_grad():
            prediction = self.model.inference(image=chunk_image_paths, process_res_method="upper_bound_resize")
        return {"chunk_idx": self.chunk_count, "image_paths": chunk_image_paths,
                "processed_images": prediction.processed_images, "depth": prediction.depth, "conf": prediction.conf,
                "extrinsics": prediction.extrinsics, "intrinsics": prediction.intrinsics}

    def load_chunk_image_paths(self) -> List[str]:
        return list(self.frame_buffer)[:self.chunk_size]

    def should_run_chunk_prediction(self) -> bool:
        return len(self.frame_buffer) >= self.chunk_size

    def process_frame(self, image_path: str):
        """solver.py:192-228."""
        self.frame_buffer.append(image_path)
        if not self.should_run_chunk_prediction():
            return
        print("=" * 50)
        print(f"\n  Processing chunk {self.chunk_count}...")
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
        time.sleep(self.config["Model"]["sleep_between_chunk"])
        print("  Sleep for observation")
        print("=" * 50)

    def run(self):
        """solver.py:230-246."""
        print("=" * 50)
        print("Starting DA3-SLAM ...")
        print("=" * 50)
        image_paths = load_image(self.image_dir)
        if not image_paths:
            print(f"Warning: No images found in {self.image_dir}")
            return
        for img_path in extract_keyframe(image_paths, self.config["Model"]["keyframe_interval"]):
            self.process_frame(img_path)
        print("=" * 50)
        print("SLAM process completed!")
