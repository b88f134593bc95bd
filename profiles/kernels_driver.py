"""Runs the kernels that bench.py's step does not reach, at BASELINE shapes, so that ncu can capture them:

    ncu --set full -k regex:'^(apply_sim3|icp_|points_moments|unproject_filter)' python profiles/kernels_driver.py

apply_sim3_kernel (K5, utils/geometry.py:43-70) on one 16-frame 518 x 518 submap, points_moments_kernel (array-level
Umeyama / IRLS, utils/align.py:14-40, :169-211) and icp_build / icp_iter (align_geometry.py:84-140) on one 518 x 518
overlap frame, unproject_filter_kernel (K1) on one submap.  Prints CUDA-event timings (not under ncu: plain run)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from da3slam_b200 import _lib as L  # noqa: E402
from da3slam_b200 import ops, synth  # noqa: E402
from da3slam_b200.pipeline import DeviceSubmap  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    H = W = 518
    subs, gt = synth.make_sequence_device(2, 16, H, W, 1, seed=7, with_images=False, device=dev)
    a, b = (DeviceSubmap.from_prediction(s, dev) for s in subs)
    out = {}
    xyz, mask, _ = ops.unproject_filter(a.depth, a.conf, a.cams, mode="fast", world=True, conf_cmp=">", conf_thr=0.1, depth_eps=1e-6)
    n_pts = xyz.numel() // 3
    out["unproject_filter_16x518x518_ms"] = timed(lambda: ops.unproject_filter(a.depth, a.conf, a.cams, mode="fast", world=True, conf_cmp=">",
                                                                                conf_thr=0.1, depth_eps=1e-6, xyz_out=xyz, mask_out=mask.view(torch.uint8),
                                                                                want_count=False))
    row = ops.sim3_row(1.1, torch.eye(3).numpy(), [0.1, 0.2, 0.3], dev)
    ms = timed(lambda: ops.apply_sim3(xyz, row, out_f64=False))
    out["apply_sim3_f32_ms"], out["apply_sim3_f32_GBps"] = ms, 24.0 * n_pts / ms / 1e6
    ms = timed(lambda: ops.apply_sim3(xyz, row, out_f64=True))
    out["apply_sim3_f64out_ms"], out["apply_sim3_f64out_GBps"] = ms, 36.0 * n_pts / ms / 1e6
    # one overlap frame as materialised clouds (the reference's array-level API)
    pa, _, _ = ops.unproject_filter(a.depth[-1:], None, a.cams[-1:], mode="fast", world=False, want_mask=False, want_count=False)
    pb, _, _ = ops.unproject_filter(b.depth[:1], None, b.cams[:1], mode="fast", world=False, want_mask=False, want_count=False)
    src, dst = pb.view(-1, 3), pa.view(-1, 3)
    n = src.shape[0]
    w = torch.ones(n, dtype=torch.float32, device=dev)
    ms = timed(lambda: ops.umeyama_points(src, dst, w, L.UMEYAMA_WEIGHTED))
    out["points_moments_f32_268k_ms"], out["points_moments_f32_GBps"] = ms, 28.0 * n / ms / 1e6
    big_s, big_d = src.repeat(64, 1).contiguous(), dst.repeat(64, 1).contiguous()
    big_w = w.repeat(64).contiguous()
    ms = timed(lambda: ops.umeyama_points(big_s, big_d, big_w, L.UMEYAMA_WEIGHTED))
    out["points_moments_f32_17M_ms"], out["points_moments_f32_17M_GBps"] = ms, 28.0 * big_s.shape[0] / ms / 1e6
    conf = torch.ones(n, dtype=torch.float32, device=dev)
    out["irls_points_268k_ms"] = timed(lambda: ops.irls_points(src, dst, conf, conf), reps=3)
    out["icp_sim3_30it_268k_ms"] = timed(lambda: ops.icp_points(src, dst, 0.05, 30, L.ICP_SIM3), reps=2)
    out["icp_rigid_50it_268k_ms"] = timed(lambda: ops.icp_points(src, dst, 0.1, 50, L.ICP_RIGID), reps=2)
    out["fp32_peak_tflops"] = ops.fp32_peak_tflops(dev)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
