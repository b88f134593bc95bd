"""Drop-in for the reference's ``viewer.py::SLAMViewer`` (constructor, ``add_frame``,
``clear``, ``run``) with the per-point arithmetic on the B200 and the map kept RESIDENT on
the device.

Reference behaviour kept (viewer.py:156-247, :317-356):
  * unproject with the VGGT closed form to world coordinates, stride mask, validity
    ``0.1 < z_world < 50`` and finite;
  * threshold = percentile(conf[conf > 0], min(slider, 99.9)) over the WHOLE map, keep
    ``conf >= threshold``; optional single-frame filter.
Changed on purpose: the reference re-stacks every stored array on every ``add_frame``
(O(frames^2), viewer.py:323-330); here a frame is unprojected straight into the tail of append-only device
buffers (capacity doubles: O(frame) per call), the threshold is one exact selection over the resident
confidences, and the push is one ordered filter + compaction kernel (``da3s_filter_points``).  Setting ``vis_voxel`` (scene units,
an attribute — the constructor keeps the reference's signature) makes ``visible_points()`` hand over the
voxel-downsampled subset of the filtered map instead of every point (oracle/SPEC.md section 5).
viser is optional — without it the viewer runs headless and ``visible_points()`` returns what would
have been pushed.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from da3slam_b200 import _lib as _L
from da3slam_b200 import ops as _ops

try:  # presentation layer only
    import viser  # type: ignore
except Exception:  # pragma: no cover
    viser = None


class SLAMViewer:
    def __init__(self, port: int = 8080, vis_stride: int = 1, vis_point_size: float = 0.003):
        if not torch.cuda.is_available():
            raise RuntimeError("SLAMViewer keeps the map on the GPU: CUDA device required (no CPU fallback)")
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.vis_stride = int(vis_stride)
        self.vis_point_size = vis_point_size
        self.conf_percent = 65.0                 # viewer.py:86-88 slider default
        self.frame_selector = "All"              # viewer.py:90-92 dropdown default
        self.vis_voxel = None                    # > 0: push one averaged point per occupied voxel of this size
        self._lock = threading.Lock()            # the reference mutates its lists from GUI threads unguarded
        self.server = None
        self.point_cloud = None
        if viser is not None:  # pragma: no cover
            self.server = viser.ViserServer(host="0.0.0.0", port=port)
            self.server.gui.configure_theme(titlebar_content=None, control_layout="collapsible")
        self.clear()

    # -- the map: append-only device buffers (capacity doubles), one contiguous block of H*W entries per frame
    def clear(self):
        with self._lock:
            self._cap = 0
            self._used = 0
            self._xyz = self._rgb = self._conf = self._valid = None
            self.frames = []         # (start, count, frame_idx) of every stored frame's block
            self.camera_poses = []
            self.next_frame_id = 0
            self.total_points = 0

    def _reserve(self, n_new):
        """Room for n_new more entries: grows geometrically, so appending F frames copies O(total) bytes overall
        (the reference re-stacks the whole map on every frame, viewer.py:323-330: O(F^2))."""
        need = self._used + n_new
        if need <= self._cap:
            return
        cap = max(need, 2 * self._cap, 1 << 18)
        dev = self.device
        new = (torch.empty((cap, 3), dtype=torch.float32, device=dev), torch.empty((cap, 3), dtype=torch.uint8, device=dev),
               torch.empty((cap,), dtype=torch.float32, device=dev), torch.empty((cap,), dtype=torch.uint8, device=dev))
        if self._used:
            for dst, src in zip(new, (self._xyz, self._rgb, self._conf, self._valid)):
                dst[:self._used].copy_(src[:self._used])
        self._xyz, self._rgb, self._conf, self._valid = new
        self._cap = cap

    def add_frame(self, image: np.ndarray, depth: np.ndarray, conf: np.ndarray, extrinsic: np.ndarray, intrinsic: np.ndarray):
        """image (3,H,W) or (H,W,3) in [0,1]; depth (H,W) or (H,W,1); conf (H,W); extrinsic (3,4) w2c;
        intrinsic (3,3)  (viewer.py:156-175).  Work and bytes moved are O(this frame): the new block is unprojected
        straight into the tail of the resident buffers."""
        frame_idx = self.next_frame_id
        self.next_frame_id += 1
        depth = np.asarray(depth, np.float32)
        if depth.ndim == 3:
            depth = depth[:, :, 0]
        H, W = depth.shape
        colors = np.asarray(image)
        if colors.shape[0] == 3:
            colors = colors.transpose(1, 2, 0)
        if colors.shape[:2] != (H, W):           # viewer.py:192-193 (cv2.INTER_LINEAR)
            import cv2
            colors = cv2.resize(np.ascontiguousarray(colors, np.float32), (W, H), interpolation=cv2.INTER_LINEAR)
        col_u8 = np.ascontiguousarray((colors * 255).astype(np.uint8))
        dev = self.device
        n = H * W
        conf = np.asarray(conf)
        c_host = conf.astype(np.float32) if conf.shape == (H, W) else np.ones((H, W), np.float32)   # viewer.py:211
        with self._lock:
            self._reserve(n)
            lo, hi = self._used, self._used + n
            d = torch.from_numpy(depth).to(dev)[None]
            self._conf[lo:hi].copy_(torch.from_numpy(c_host).view(-1), non_blocking=True)
            self._rgb[lo:hi].copy_(torch.from_numpy(col_u8).view(-1, 3), non_blocking=True)
            cams = _ops.build_cams(torch.from_numpy(np.asarray(intrinsic, np.float32))[None].to(dev),
                                   torch.from_numpy(np.asarray(extrinsic, np.float32))[None].to(dev))
            # closed-form unprojection to world + validity 0.1 < z < 50 & finite (viewer.py:198-218), written in place
            _ops.unproject_filter(d, self._conf[lo:hi].view(1, H, W), cams, mode="closed", world=True, world_z=True, want_count=False,
                                  xyz_out=self._xyz[lo:hi].view(1, H, W, 3), mask_out=self._valid[lo:hi].view(1, H, W))
            if self.vis_stride > 1:              # viewer.py:205-206
                v2 = self._valid[lo:hi].view(H, W)
                keep = torch.zeros((H, W), dtype=torch.uint8, device=dev)
                keep[::self.vis_stride, ::self.vis_stride] = 1
                v2.mul_(keep)
            self._conf[lo:hi].mul_(self._valid[lo:hi])               # 0 = "not in the map": the percentile below skips those entries
            n_new = int(self._valid[lo:hi].sum().item())
            if n_new > 0:                        # viewer.py:220-234: a frame without a valid point is not stored
                self._used = hi
                self.frames.append((lo, n, frame_idx))
                self.camera_poses.append(np.asarray(extrinsic))
                self.total_points += n_new
        self._update_point_cloud()

    # -- map-wide confidence filter (viewer.py:317-356)
    def _threshold(self):
        """percentile(conf[stored & conf > 0], min(slider, 99.9)) over the whole map: one exact selection over the
        resident confidences (no copy; entries that are not in the map hold 0 and SEL_POSITIVE skips them)."""
        if not self._used:
            return None
        sel = _ops.select([dict(a=self._conf[:self._used], kind=_L.SEL_POSITIVE, stat=_L.SEL_PERCENTILE,
                                percent=float(min(self.conf_percent, 99.9)))], self.device)[0]
        if sel["n_valid"] == 0:
            return None                          # no positive confidence: everything is shown (viewer.py:337-338)
        return np.float32(sel["value"])

    def visible_points(self):
        """(points [n,3] float32, colors [n,3] uint8) that the reference would hand to viser: one ordered
        filter + compaction kernel over the resident map (da3s_filter_points), or over one frame's block."""
        with self._lock:
            if not self._used:
                return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8)
            thr = self._threshold()
            lo, hi = 0, self._used
            if self.frame_selector != "All":
                try:
                    want = int(self.frame_selector)
                    hit = [(s, c) for s, c, f in self.frames if f == want]
                    if not hit:
                        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8)
                    lo, hi = hit[0][0], hit[0][0] + hit[0][1]
                except ValueError:
                    pass
            all_pts, all_cols = _ops.filter_points(self._xyz[lo:hi], self._rgb[lo:hi], self._conf[lo:hi], self._valid[lo:hi],
                                                   thr=None if thr is None else float(thr))
            if self.vis_voxel and all_pts.shape[0] > 0:
                vx, vc, _, _ = _ops.voxel_downsample([(all_pts.contiguous(), all_cols.contiguous(), None)], float(self.vis_voxel))
                return vx.cpu().numpy(), vc.cpu().numpy()
            return all_pts.cpu().numpy(), all_cols.cpu().numpy()

    def _update_point_cloud(self):
        if self.server is None:
            return
        pts, cols = self.visible_points()  # pragma: no cover
        if len(pts):  # pragma: no cover
            self.point_cloud = self.server.scene.add_point_cloud(name="slam_pcd", points=pts, colors=cols,
                                                                 point_size=self.vis_point_size, point_shape="circle")

    def run(self, background: bool = False):
        """The reference blocks in a sleep loop unless background (viewer.py:417-433)."""
        if background or self.server is None:
            return
        import time  # pragma: no cover
        while True:  # pragma: no cover
            time.sleep(0.01)
