"""Drop-in for the reference's ``viewer.py::SLAMViewer`` (constructor, ``add_frame``,
``clear``, ``run``) with the per-point arithmetic on the B200 and the map kept RESIDENT on
the device.

Reference behaviour kept (viewer.py:156-247, :317-356):
  * unproject with the VGGT closed form to world coordinates, stride mask, validity
    ``0.1 < z_world < 50`` and finite;
  * threshold = percentile(conf[conf > 0], min(slider, 99.9)) over the WHOLE map, keep
    ``conf >= threshold``; optional single-frame filter.
Changed on purpose: the reference re-stacks every stored array on every ``add_frame``
(O(frames^2), viewer.py:323-330); here a frame is appended once to device buffers and the
threshold is one exact selection over the resident confidences.  Setting ``vis_voxel`` (scene units,
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

    # -- reference attribute names (viewer.py:32-47) kept for callers that inspect them
    def clear(self):
        with self._lock:
            self.points = []         # per frame [H,W,3] float32 world points (device)
            self.colors = []         # per frame [H,W,3] uint8 (device)
            self.conf = []           # per frame [H,W] float32 with 0 where the pixel is not stored (device)
            self.camera_poses = []
            self.next_frame_id = 0
            self.total_points = 0

    def add_frame(self, image: np.ndarray, depth: np.ndarray, conf: np.ndarray, extrinsic: np.ndarray, intrinsic: np.ndarray):
        """image (3,H,W) or (H,W,3) in [0,1]; depth (H,W) or (H,W,1); conf (H,W); extrinsic (3,4) w2c;
        intrinsic (3,3)  (viewer.py:156-175)."""
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
        col_u8 = (colors * 255).astype(np.uint8)
        dev = self.device
        d = torch.from_numpy(depth).to(dev)[None]
        conf = np.asarray(conf)
        c_host = conf.astype(np.float32) if conf.shape == (H, W) else np.ones((H, W), np.float32)   # viewer.py:211
        c = torch.from_numpy(c_host).to(dev)[None]
        cams = _ops.build_cams(torch.from_numpy(np.asarray(intrinsic, np.float32))[None].to(dev),
                               torch.from_numpy(np.asarray(extrinsic, np.float32))[None].to(dev))
        # closed-form unprojection to world + validity 0.1 < z < 50 & finite (viewer.py:198-218)
        xyz, valid, _ = _ops.unproject_filter(d, c, cams, mode="closed", world=True, world_z=True, want_count=False)
        if self.vis_stride > 1:                  # viewer.py:205-206
            stride = torch.zeros((H, W), dtype=torch.bool, device=dev)
            stride[::self.vis_stride, ::self.vis_stride] = True
            valid = valid & stride[None]
        stored_conf = torch.where(valid[0], c[0], torch.zeros_like(c[0]))       # 0 = "not in the map"
        n_new = int(valid.sum().item())
        with self._lock:
            if n_new > 0:                        # viewer.py:220-234
                self.points.append(xyz[0])
                self.colors.append(torch.from_numpy(col_u8).to(dev))
                self.conf.append((stored_conf, valid[0], frame_idx))
                self.camera_poses.append(np.asarray(extrinsic))
                self.total_points += n_new
        self._update_point_cloud()

    # -- map-wide confidence filter (viewer.py:317-356)
    def _threshold(self):
        """percentile(conf[stored & conf > 0], min(slider, 99.9)) over the whole map: one exact selection."""
        if not self.conf:
            return None
        allc = torch.cat([c.reshape(-1) for c, _, _ in self.conf])
        sel = _ops.select([dict(a=allc, kind=_L.SEL_POSITIVE, stat=_L.SEL_PERCENTILE,
                                percent=float(min(self.conf_percent, 99.9)))], self.device)[0]
        if sel["n_valid"] == 0:
            return None                          # no positive confidence: everything is shown (viewer.py:337-338)
        return np.float32(sel["value"])

    def visible_points(self):
        """(points [n,3] float32, colors [n,3] uint8) that the reference would hand to viser."""
        with self._lock:
            if not self.points:
                return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8)
            thr = self._threshold()
            pts, cols = [], []
            for xyz, col, (c, valid, fidx) in zip(self.points, self.colors, self.conf):
                m = valid if thr is None else (valid & (c >= float(thr)))
                if self.frame_selector != "All":
                    try:
                        if int(self.frame_selector) != fidx:
                            continue
                    except ValueError:
                        pass
                pts.append(xyz[m])
                cols.append(col[m])
            if not pts:
                return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8)
            all_pts, all_cols = torch.cat(pts), torch.cat(cols)
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
