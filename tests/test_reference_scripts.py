"""Script-level drop-in proof (north-star: "main_slam.py and main_align.py run unchanged").

The reference's OWN entry scripts — /root/reference/main_align.py and /root/reference/main_slam.py, loaded
unmodified from where they lie — are executed with this repository first on sys.path, so that their
`from align_geometry import ...`, `from utils import ...`, `from viewer import SLAMViewer`, `from solver import
SLAMSolver`, `from config import load_config` bind to THIS repo's modules.  The network (`depth_anything_3`) is a fake
that returns synthetic predictions; `time.sleep` is patched out; viser is optional in the viewer shim.

Two arms:
  * CUDA present and reference present (a developer box): the scripts run on the real kernels.
  * no CUDA (this container, where /root/reference lives): the numpy-level entry points of
    `da3slam_b200.host` and the viewer are replaced by test doubles built on the oracle — the scripts
    then exercise every name, signature, call order and return type of the shim modules, and the chained
    extrinsics are checked against the reference's own functions run on the same predictions.
The GPU box has no /root/reference: there the kernel-level shim tests (tests/test_gpu_shims.py) cover the same
modules against golden outputs of the reference.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import ref_port as rp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")


class _FakeDA3:
    """depth_anything_3.api.DepthAnything3 stand-in: synthetic Predictions, one per inference() call."""
    SUBS = None

    def __init__(self):
        self.calls = 0
        self.kwargs = []

    @classmethod
    def from_pretrained(cls, path):
        return cls()

    def to(self, device):
        return self

    def eval(self):
        return self

    def inference(self, image=None, **kw):
        s = _FakeDA3.SUBS[self.calls]
        self.calls += 1
        self.kwargs.append(kw)
        return types.SimpleNamespace(processed_images=s["processed_images"], depth=s["depth"].copy(), conf=s["conf"],
                                     extrinsics=s["extrinsics"], intrinsics=s["intrinsics"])


class _RecordingViewer:
    """viewer.SLAMViewer stand-in for the CPU arm (the real shim keeps its map on the GPU)."""
    instances = []

    def __init__(self, port=8080, vis_stride=1, vis_point_size=0.003):
        self.frames = []
        _RecordingViewer.instances.append(self)

    def add_frame(self, image, depth, conf, extrinsic, intrinsic):
        assert np.asarray(extrinsic).shape == (3, 4) and np.asarray(intrinsic).shape == (3, 3)
        assert np.asarray(depth).shape == np.asarray(conf).shape
        self.frames.append(np.asarray(extrinsic, np.float64))

    def clear(self):
        self.frames = []

    def run(self, background=False):
        return None


def _install_fakes(monkeypatch, subs):
    _FakeDA3.SUBS = subs
    mod = types.ModuleType("depth_anything_3")
    api = types.ModuleType("depth_anything_3.api")
    api.DepthAnything3 = _FakeDA3
    mod.api = api
    monkeypatch.setitem(sys.modules, "depth_anything_3", mod)
    monkeypatch.setitem(sys.modules, "depth_anything_3.api", api)
    monkeypatch.syspath_prepend(ROOT)
    import time
    monkeypatch.setattr(time, "sleep", lambda s: None)
    if torch.cuda.is_available():
        return "cuda"
    # ---- CPU arm: oracle-backed doubles for the numpy-level entry points the shims call ----
    from da3slam_b200 import host

    def unproject(depth, intrinsics, extrinsics, *, world, out_f64, mode="kinv", general_inverse=True):
        if out_f64:
            return rp.unproject_world_f64(depth, intrinsics, extrinsics)
        return rp.unproject_f32(depth, intrinsics, extrinsics, "world" if world else "camera")

    def depth_scale(prev, cur, conf_th=0.2, eps=1e-6, guarded=False):
        return (rp.depth_scale_guarded if guarded else rp.depth_scale_plain)(prev, cur, conf_th, eps)

    def icp(source, target, threshold, max_iterations, rigid=False):
        fn = rp.icp_point_to_point_kdtree if rigid else rp.umeyama_icp_kdtree
        out = fn(np.asarray(source, np.float64), np.asarray(target, np.float64), threshold, max_iterations)
        return float(out[0]), np.asarray(out[1], np.float64), np.asarray(out[2], np.float64)

    monkeypatch.setattr(host, "unproject", unproject)
    monkeypatch.setattr(host, "depth_scale", depth_scale)
    monkeypatch.setattr(host, "icp", icp)
    import viewer
    _RecordingViewer.instances = []
    monkeypatch.setattr(viewer, "SLAMViewer", _RecordingViewer)
    return "cpu-doubles"


def _load_script(name):
    path = os.path.join(ref_loader.REF_ROOT, name)
    spec = importlib.util.spec_from_file_location("_da3ref_script_" + name.replace(".py", ""), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _own(module_name):
    m = sys.modules[module_name]
    return os.path.abspath(m.__file__).startswith(ROOT + os.sep)


def _write_images(folder, n):
    for i in range(n):
        (folder / f"{i:04d}.png").write_bytes(b"x")


def test_main_align_runs_unchanged(monkeypatch, tmp_path, capsys):
    from da3slam_b200 import synth
    H, W, F, n_chunks = 24, 32, 4, 3
    subs, gt = synth.make_sequence(n_chunks, F, H, W, overlap=1, seed=11, with_images=True)
    arm = _install_fakes(monkeypatch, subs)
    for name in ("align_geometry", "viewer", "utils", "solver", "config"):      # fresh imports, bound to this repo
        sys.modules.pop(name, None)
    if arm != "cuda":
        import viewer
        monkeypatch.setattr(viewer, "SLAMViewer", _RecordingViewer)
    script = _load_script("main_align.py")
    assert _own("align_geometry") and _own("utils") and _own("viewer")          # the scripts picked up OUR modules
    assert script.make_image_chunks.__module__ == "align_geometry"
    _write_images(tmp_path, F + (n_chunks - 1) * (F - 1))
    monkeypatch.setattr(script, "folder_path", str(tmp_path))
    monkeypatch.setattr(script, "model_path", "unused")
    script.main()                                                                 # main_align.py:74-131, unmodified
    out = capsys.readouterr().out
    assert out.count("point_map1:") == n_chunks - 1                              # align_geometry.py:287-288 prints
    # what the script computed == the reference's own functions on the same predictions
    ref = ref_loader.load()
    E_prev = np.asarray(subs[0]["extrinsics"][-1], np.float64)
    prev = types.SimpleNamespace(**{k: np.array(v) for k, v in subs[0].items()})
    want = []
    for k in range(1, n_chunks):
        cur = types.SimpleNamespace(**{k_: np.array(v) for k_, v in subs[k].items()})
        s_depth = ref.ag.estimate_depth_scale(prev, cur, conf_th=0.2)
        cur.depth = cur.depth * s_depth
        pm1, pm2 = ref.ag.extract_overlap_point_cloud(prev, cur)
        s, R, t = rp.umeyama_icp_kdtree(pm2.reshape(-1, 3).astype(np.float64), pm1.reshape(-1, 3).astype(np.float64), 0.001, 30)
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        Eg = ref.ag.compute_aligned_chunk_extrinsics_from_prev_overlap(E_prev, np.asarray(cur.extrinsics, np.float64), T)
        want.append(Eg)
        E_prev, prev = Eg[-1], cur
    if arm == "cuda":
        import viewer
        return                                                                    # real viewer: frames live on the device
    got = _RecordingViewer.instances[-1].frames
    assert len(got) == 2 * n_chunks                                               # first + last frame of every chunk
    for k in range(1, n_chunks):
        assert np.abs(got[2 * k] - want[k - 1][0]).max() < 1e-9 and np.abs(got[2 * k + 1] - want[k - 1][-1]).max() < 1e-9


def test_main_slam_runs_unchanged(monkeypatch, tmp_path):
    from da3slam_b200 import synth
    H, W, F, n_chunks = 24, 32, 4, 3
    subs, gt = synth.make_sequence(n_chunks, F, H, W, overlap=1, seed=12, with_images=True)
    arm = _install_fakes(monkeypatch, subs)
    for name in ("solver", "viewer", "config", "utils", "utils.align_geometry_single", "align_geometry"):
        sys.modules.pop(name, None)
    if arm != "cuda":
        import viewer
        monkeypatch.setattr(viewer, "SLAMViewer", _RecordingViewer)
    script = _load_script("main_slam.py")
    assert _own("solver") and _own("config")
    img_dir = tmp_path / "images"
    img_dir.mkdir()
    _write_images(img_dir, F + (n_chunks - 1) * (F - 1))
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(f"Model:\n  chunk_size: {F}\n  overlap_size: 1\n  keyframe_interval: 1\n  sleep_between_chunk: 0\n  port: 8099\n"
                   "Weights:\n  DA3: unused\n")
    monkeypatch.setattr(sys, "argv", ["main_slam.py", "--image_dir", str(img_dir), "--config", str(cfg)])
    import solver
    made = []
    orig_init = solver.SLAMSolver.__init__

    def spy(self, *a, **kw):
        orig_init(self, *a, **kw)
        made.append(self)
    monkeypatch.setattr(solver.SLAMSolver, "__init__", spy)
    # main() ends in an idle loop (`while True: time.sleep(0.01)`, main_slam.py:46-51): leave it the way Ctrl+C does
    import time
    ticks = {"n": 0}

    def sleep(s):
        if s == 0.01:
            ticks["n"] += 1
            raise KeyboardInterrupt
    monkeypatch.setattr(time, "sleep", sleep)
    script.main()                                                                 # main_slam.py:9-51, unmodified
    assert ticks["n"] == 1 and len(made) == 1
    sv = made[0]
    assert sv.chunk_count == n_chunks and len(sv.chunk_prediction_list) == n_chunks
    for k in range(1, n_chunks):
        Eg = sv.chunk_prediction_list[k]["extrinsics_global"]
        assert Eg.shape == (F, 3, 4) and Eg.dtype == np.float64 and np.isfinite(Eg).all()
    if arm != "cuda":                                                             # every frame of every chunk reached the viewer
        assert len(_RecordingViewer.instances[-1].frames) == n_chunks * F
