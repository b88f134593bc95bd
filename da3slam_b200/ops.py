"""torch-facing wrappers over the C ABI (include/da3s.h).

torch is plumbing only: it owns device memory and streams; every computation on the
path is a kernel of libda3s.so launched on torch's current stream.  All functions take
CUDA tensors and raise on CPU tensors (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

_contexts = {}
_retired = []        # outgrown contexts stay alive: objects created earlier (VoxelGrid, plans) keep using theirs


class Context:
    """One da3s_ctx per device (+ its workspace)."""

    def __init__(self, device: torch.device, workspace_bytes: int):
        self.lib = L.load()
        self.device = device
        h = C.c_void_p()
        L.check(self.lib.da3s_create(device.index or 0, workspace_bytes, C.byref(h)), "da3s_create")
        self.h = h
        self.workspace_bytes = workspace_bytes

    def close(self):
        if self.h:
            self.lib.da3s_destroy(self.h)
            self.h = None

    @property
    def launches(self) -> int:
        return int(self.lib.da3s_launch_count(self.h))

    def kernel_timers(self, on: bool = True) -> None:
        """Event pairs around the library's four big kernels (include/da3s.h: da3s_kernel_timers)."""
        L.check(self.lib.da3s_kernel_timers(self.h, int(bool(on))), "da3s_kernel_timers")

    def kernel_time(self, which: int):
        """(sum in ms of the `timed` most recent launch durations, timed, launches since the last read, work units all
        those launches executed — RANSAC scoring only); synchronises with the last launch."""
        ms, timed, n, work = C.c_double(0.0), C.c_int(0), C.c_int(0), C.c_double(0.0)
        L.check(self.lib.da3s_kernel_time(self.h, int(which), C.byref(ms), C.byref(timed), C.byref(n), C.byref(work)), "da3s_kernel_time")
        return float(ms.value), int(timed.value), int(n.value), float(work.value)


def context(device=None, workspace_bytes: int | None = None) -> Context:
    if not torch.cuda.is_available():
        raise RuntimeError("da3slam_b200 needs a CUDA device: the alignment path has no CPU fallback")
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if device.type != "cuda":
        raise RuntimeError("da3slam_b200 runs on CUDA devices only")
    if device.index is None:
        device = torch.device(f"cuda:{torch.cuda.current_device()}")
    ctx = _contexts.get(device.index)
    want = workspace_bytes or (1 << 30)
    if ctx is None or ctx.workspace_bytes < want:
        if ctx is not None:
            _retired.append(ctx)
        with torch.cuda.device(device):
            ctx = Context(device, want)
        _contexts[device.index] = ctx
    return ctx


class _NoRange:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NVTX = None


def nvtx_range(name: str):
    """NVTX range around a stage (tracing row): visible in nsys / ncu --nvtx timelines.  Enabled with
    DA3S_NVTX=1 (a push/pop pair costs ~1 us of host time per stage, so it is off in the timed default)."""
    global _NVTX
    if _NVTX is None:
        import os
        _NVTX = os.environ.get("DA3S_NVTX", "0") not in ("", "0")
    if not _NVTX:
        return _NoRange()
    return torch.cuda.nvtx.range(name)


def fp32_peak_tflops(device=None, iters: int = 64) -> float:
    """float32 FMA throughput of the device measured with the library's microbenchmark kernel (synchronises)."""
    ctx = context(device)
    out = C.c_double(0.0)
    st = C.c_void_p(torch.cuda.current_stream(ctx.device).cuda_stream)
    L.check(ctx.lib.da3s_measure_fp32_peak(ctx.h, int(iters), C.byref(out), st), "da3s_measure_fp32_peak")
    return float(out.value)


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t, name="tensor", dtype=None):
    if t is None:
        return C.c_void_p(0)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------
# camera table
# ------------------------------------------------------------------------------------------
def build_cams(intrinsics: torch.Tensor, extrinsics: torch.Tensor, general_inverse: bool = False, out=None) -> torch.Tensor:
    """[n,3,3] f32, [n,3,4] f32 -> opaque uint8 tensor [n, 200] holding da3s_cam records (`out`: refresh in place)."""
    K = intrinsics.to(torch.float32).contiguous()
    E = extrinsics.to(torch.float32).contiguous()
    n = K.shape[0]
    assert K.shape == (n, 3, 3) and E.shape == (n, 3, 4)
    ctx = context(K.device)
    cams = out if out is not None else torch.empty((n, L.CAM_BYTES), dtype=torch.uint8, device=K.device)
    assert cams.shape == (n, L.CAM_BYTES) and cams.dtype == torch.uint8 and cams.is_contiguous()
    rc = ctx.lib.da3s_build_cams(ctx.h, _ptr(K), _ptr(E), n, L.CAM_GENERAL_INV if general_inverse else L.CAM_CLOSED_FORM,
                                 _ptr(cams), _stream(K))
    L.check(rc, "da3s_build_cams")
    return cams


# ------------------------------------------------------------------------------------------
# K1 / K5
# ------------------------------------------------------------------------------------------
def unproject_filter(depth, conf, cams, *, mode="closed", world=False, out_f64=False, conf_cmp=None,
                     conf_thr=0.0, conf_thr_dev=None, conf_floor=None, depth_eps=None, world_z=False,
                     sim3=None, want_mask=True, want_count=True, xyz_out=None, mask_out=None):
    """depth [N,H,W] f32 (+ conf) -> (xyz [N,H,W,3], mask [N,H,W] bool or None, n_kept tensor or None).
    xyz_out / mask_out (uint8) may be preallocated buffers of the right shape."""
    depth = depth.contiguous()
    N, H, W = depth.shape
    ctx = context(depth.device)
    flags = {"closed": L.UNPROJ_CLOSED, "kinv": L.UNPROJ_KINV, "fast": L.UNPROJ_FAST}[mode]
    if world:
        flags |= L.UNPROJ_WORLD
    if out_f64:
        flags |= L.UNPROJ_OUT_F64
    if conf_cmp == ">":
        flags |= L.MASK_CONF_GT
    elif conf_cmp == ">=":
        flags |= L.MASK_CONF_GE
    elif conf_cmp is not None:
        raise ValueError("conf_cmp must be '>' or '>='")
    if conf_floor is not None:
        flags |= L.MASK_CONF_FLOOR
    if depth_eps is not None:
        flags |= L.MASK_DEPTH
    if world_z:
        flags |= L.MASK_WORLD_Z
    s3 = None
    if sim3 is not None:
        s3 = sim3.to(torch.float64).contiguous()
        if s3.dim() == 2:
            assert s3.shape == (N, 13)
            flags |= L.SIM3_PER_FRAME
        else:
            assert s3.shape == (13,)
    xyz = xyz_out if xyz_out is not None else torch.empty((N, H, W, 3), dtype=torch.float64 if out_f64 else torch.float32,
                                                          device=depth.device)
    assert xyz.numel() == N * H * W * 3 and xyz.dtype == (torch.float64 if out_f64 else torch.float32)
    mask = mask_out if mask_out is not None else (torch.empty((N, H, W), dtype=torch.uint8, device=depth.device) if want_mask else None)
    cnt = torch.zeros((1,), dtype=torch.int64, device=depth.device) if want_count else None
    rc = ctx.lib.da3s_unproject_filter(
        ctx.h, _ptr(depth, "depth", torch.float32), _ptr(conf, "conf", torch.float32), _ptr(cams, "cams"), N, H, W, flags,
        float(conf_thr), _ptr(conf_thr_dev, "conf_thr_dev", torch.float32), float(conf_floor or 0.0), float(depth_eps or 0.0),
        _ptr(s3), _ptr(xyz), _ptr(mask), _ptr(cnt), _stream(depth))
    L.check(rc, "da3s_unproject_filter")
    return xyz, (mask.view(torch.bool) if mask is not None else None), cnt


def make_frame_jobs(jobs, device) -> torch.Tensor:
    """jobs: list of dicts(depth, conf, cam, sim3, conf_thr, xyz, mask) of CUDA tensors / None (views into
    persistent buffers).  Returns the device job table for unproject_filter_jobs."""
    arr = (L.FrameJob * len(jobs))()
    for i, j in enumerate(jobs):
        for name in ("depth", "conf", "cam", "sim3", "conf_thr", "xyz", "mask"):
            t = j.get(name)
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError("frame job tensors must live on the GPU")
            if name in ("depth", "conf", "xyz") and t.data_ptr() % 16:
                raise L.Da3sError(L.EALIGN, "make_frame_jobs", f"{name} is not 16-byte aligned")
            setattr(arr[i], name, t.data_ptr())
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


def unproject_filter_jobs(job_table: torch.Tensor, n_frames: int, H: int, W: int, *, mode="fast", world=True, out_f64=False,
                          conf_cmp=None, conf_thr=0.0, conf_floor=None, depth_eps=None, world_z=False, n_kept=None):
    """One launch over all frames of the job table (see da3s_unproject_filter_jobs)."""
    ctx = context(job_table.device)
    flags = {"closed": L.UNPROJ_CLOSED, "kinv": L.UNPROJ_KINV, "fast": L.UNPROJ_FAST}[mode]
    flags |= (L.UNPROJ_WORLD if world else 0) | (L.UNPROJ_OUT_F64 if out_f64 else 0)
    flags |= {None: 0, ">": L.MASK_CONF_GT, ">=": L.MASK_CONF_GE}[conf_cmp]
    flags |= (L.MASK_CONF_FLOOR if conf_floor is not None else 0) | (L.MASK_DEPTH if depth_eps is not None else 0)
    flags |= L.MASK_WORLD_Z if world_z else 0
    rc = ctx.lib.da3s_unproject_filter_jobs(ctx.h, _ptr(job_table), n_frames, H, W, flags, float(conf_thr),
                                            float(conf_floor or 0.0), float(depth_eps or 0.0), _ptr(n_kept), _stream(job_table))
    L.check(rc, "da3s_unproject_filter_jobs")


def sim3_row(s, R, t, device) -> torch.Tensor:
    row = np.empty(13, np.float64)
    row[0] = s
    row[1:10] = np.asarray(R, np.float64).reshape(-1)
    row[10:13] = np.asarray(t, np.float64).reshape(-1)
    return torch.from_numpy(row).to(device)


def apply_sim3(points: torch.Tensor, sim3: torch.Tensor, out_f64: bool | None = None) -> torch.Tensor:
    """points [...,3] f32/f64, sim3 [13] f64 (s, R row-major, t) -> s * (p R^T) + t."""
    points = points.contiguous()
    assert points.shape[-1] == 3 and points.dtype in (torch.float32, torch.float64)
    in_f64 = points.dtype == torch.float64
    if out_f64 is None:
        out_f64 = True                     # utils/geometry.py:43-70 returns float64 for any input
    ctx = context(points.device)
    out = torch.empty(points.shape, dtype=torch.float64 if out_f64 else torch.float32, device=points.device)
    s3 = sim3.to(torch.float64).contiguous()
    rc = ctx.lib.da3s_apply_sim3(ctx.h, _ptr(points), int(in_f64), points.numel() // 3, _ptr(s3), _ptr(out), int(out_f64),
                                 _stream(points))
    L.check(rc, "da3s_apply_sim3")
    return out


def filter_points(xyz, rgb, conf, valid, thr=None, xyz_out=None, rgb_out=None):
    """Ordered compaction of a resident cloud (da3s_filter_points): keep point i iff valid[i] (None = all) and, when thr is
    given, conf[i] >= thr.  xyz [n,3] f32, rgb [n,3] u8 or None, conf [n] f32, valid [n] u8/bool or None.  Returns
    (xyz_kept, rgb_kept) trimmed to the kept count (this helper synchronises to learn it)."""
    n = xyz.shape[0]
    dev = xyz.device
    ctx = context(dev)
    if n == 0:
        return xyz[:0], (rgb[:0] if rgb is not None else None)
    if valid is not None and valid.dtype == torch.bool:
        valid = valid.view(torch.uint8)
    out_xyz = xyz_out if xyz_out is not None else torch.empty((n, 3), dtype=torch.float32, device=dev)
    out_rgb = None
    if rgb is not None:
        out_rgb = rgb_out if rgb_out is not None else torch.empty((n, 3), dtype=torch.uint8, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
    rc = ctx.lib.da3s_filter_points(ctx.h, _ptr(xyz, "xyz", torch.float32), _ptr(rgb, "rgb", torch.uint8), _ptr(conf, "conf", torch.float32),
                                    _ptr(valid, "valid", torch.uint8), n, int(thr is not None), float(thr if thr is not None else 0.0),
                                    out_xyz.shape[0], _ptr(out_xyz), _ptr(out_rgb), _ptr(cnt), _stream(xyz))
    L.check(rc, "da3s_filter_points")
    k = int(cnt.item())
    return out_xyz[:k], (out_rgb[:k] if out_rgb is not None else None)


# ------------------------------------------------------------------------------------------
# exact selection
# ------------------------------------------------------------------------------------------
SELECT_OUT_BYTES = C.sizeof(L.SelectOut)
SELECT_VALUE_OFFSET = L.SelectOut.value.offset


def select(segments, device):
    """segments: list of dicts(a=tensor, [b, ca, cb], kind, stat, percent, conf_th, eps).
    Returns a structured numpy array (n_valid, lo, hi, value, gamma) — this helper
    synchronises; select_async and the pair pipeline do not."""
    host = select_async(segments, device).cpu().numpy()
    dt = np.dtype([("n_valid", np.int64), ("lo", np.float32), ("hi", np.float32), ("value", np.float32), ("gamma", np.float32)])
    return host.view(dt).reshape(len(segments))


def select_value_view(d_out: torch.Tensor) -> torch.Tensor:
    """[n] float32 strided view of the `value` field of a select_async result (stays on device)."""
    return d_out.view(torch.float32).view(d_out.shape[0], SELECT_OUT_BYTES // 4)[:, SELECT_VALUE_OFFSET // 4]


class SelectPlan:
    """A prebuilt device segment table: run() only enqueues the three selection passes."""

    def __init__(self, segments, device, private_ctx: bool = False):
        self.device = torch.device(device)
        self.n = len(segments)
        self.d_segs, self.max_n, self._keep = _build_segs(segments, device)
        self.out = torch.empty((self.n, SELECT_OUT_BYTES), dtype=torch.uint8, device=device)
        self.ctx = None
        if private_ctx:                  # its own da3s_ctx (= its own scratch): may run on a side stream next to other calls
            need = (4 << 20) + self.n * ((self.max_n // 8 + (1 << 16)) * 4 + (1 << 16))
            with torch.cuda.device(self.device):
                self.ctx = Context(self.device, need)

    def run(self) -> torch.Tensor:
        ctx = self.ctx or context(self.device)
        rc = ctx.lib.da3s_select(ctx.h, _ptr(self.d_segs), self.n, self.max_n, _ptr(self.out), _stream(self.d_segs))
        L.check(rc, "da3s_select")
        return self.out

    def value_ptr_tensor(self, i) -> torch.Tensor:
        """1-element float32 view of record i's `value` — pass as conf_thr_dev."""
        return select_value_view(self.out)[i:i + 1]


def select_async(segments, device) -> torch.Tensor:
    """Same as select() but returns the raw device records [n, 24] uint8 without synchronising
    the stream (building the segment table does copy a few bytes to the device)."""
    ctx = context(device)
    d_segs, max_n, keep = _build_segs(segments, device)
    n = len(segments)
    d_out = torch.empty((n, C.sizeof(L.SelectOut)), dtype=torch.uint8, device=device)
    rc = ctx.lib.da3s_select(ctx.h, _ptr(d_segs), n, max_n, _ptr(d_out), _stream(d_segs))
    L.check(rc, "da3s_select")
    return d_out


def _build_segs(segments, device):
    n = len(segments)
    segs = (L.SelectSeg * n)()
    keep = []
    max_n = 0
    for i, s in enumerate(segments):
        a = s["a"].contiguous().view(-1)
        keep.append(a)
        segs[i].a = a.data_ptr()
        for nm in ("b", "ca", "cb"):
            t = s.get(nm)
            if t is not None:
                t = t.contiguous().view(-1)
                keep.append(t)
                setattr(segs[i], nm, t.data_ptr())
        segs[i].n = a.numel()
        segs[i].kind = s.get("kind", L.SEL_VALUES)
        segs[i].stat = s.get("stat", L.SEL_MEDIAN)
        segs[i].percent = s.get("percent", 50.0)
        segs[i].conf_th = s.get("conf_th", 0.0)
        segs[i].eps = s.get("eps", 0.0)
        max_n = max(max_n, a.numel())
    raw = np.frombuffer(bytes(segs), dtype=np.uint8).copy()
    d_segs = torch.from_numpy(raw).to(device)
    return d_segs, max_n, keep


# ------------------------------------------------------------------------------------------
# pairs
# ------------------------------------------------------------------------------------------
def make_pairs(entries, device) -> torch.Tensor:
    """entries: list of (depth_a, conf_a, depth_b, conf_b, cams_a, cams_b) CUDA tensors (views are
    fine as long as they are contiguous and 16-byte aligned).  Returns the device pair table."""
    n = len(entries)
    arr = (L.Pair * n)()
    for i, (da, ca, db, cb, cama, camb) in enumerate(entries):
        for t in (da, ca, db, cb):
            if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32:
                raise RuntimeError("pair maps must be contiguous float32 CUDA tensors")
            if t.data_ptr() % 16:
                raise L.Da3sError(L.EALIGN, "make_pairs", "depth/conf view is not 16-byte aligned")
        arr[i].depth_a, arr[i].conf_a = da.data_ptr(), ca.data_ptr()
        arr[i].depth_b, arr[i].conf_b = db.data_ptr(), cb.data_ptr()
        arr[i].cam_a, arr[i].cam_b = cama.data_ptr(), camb.data_ptr()
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


def align_pairs(pairs: torch.Tensor, n_pairs: int, overlap: int, H: int, W: int, opts: L.AlignOpts,
                sample_idx: torch.Tensor | None = None, want_aux=False, want_counts=False, rows_out=None):
    """Runs the whole pair pipeline asynchronously; returns (rows [n,16] f64, aux [n,8] f64 or None,
    counts [n,n_hyp] i32 or None) as CUDA tensors.  rows_out: preallocated [n,16] float64 result buffer."""
    dev = pairs.device
    ctx = context(dev)
    rows = rows_out if rows_out is not None else torch.empty((n_pairs, L.ROW_LEN), dtype=torch.float64, device=dev)
    assert rows.shape == (n_pairs, L.ROW_LEN) and rows.dtype == torch.float64 and rows.is_contiguous()
    aux = torch.zeros((n_pairs, L.AUX_DOUBLES), dtype=torch.float64, device=dev) if want_aux else None
    counts = None
    if opts.n_hyp > 0:
        if sample_idx is None:
            raise ValueError("n_hyp > 0 needs sample_idx")
        sample_idx = sample_idx.to(torch.int32).contiguous()
        assert sample_idx.shape == (n_pairs, opts.n_hyp, 3)
        if want_counts:
            counts = torch.empty((n_pairs, opts.n_hyp), dtype=torch.int32, device=dev)
    rc = ctx.lib.da3s_align_pairs(ctx.h, _ptr(pairs), n_pairs, overlap, H, W, C.byref(opts), _ptr(sample_idx), _ptr(rows),
                                  _ptr(aux), _ptr(counts), _stream(pairs))
    L.check(rc, "da3s_align_pairs")
    return rows, aux, counts


def accumulate_sim3(rows: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[n,16] pair rows -> [n+1,13] cumulative Sim(3) (identity first), on the device."""
    rows = rows.contiguous()
    n = rows.shape[0]
    ctx = context(rows.device)
    cum = out if out is not None else torch.empty((n + 1, 13), dtype=torch.float64, device=rows.device)
    assert cum.shape == (n + 1, 13) and cum.dtype == torch.float64 and cum.is_contiguous()
    L.check(ctx.lib.da3s_accumulate_sim3(ctx.h, _ptr(rows, "rows", torch.float64), n, _ptr(cum), _stream(rows)),
            "da3s_accumulate_sim3")
    return cum


def pair_thresholds(pairs, n_pairs, overlap, H, W, opts):
    dev = pairs.device
    ctx = context(dev)
    thr = torch.empty((n_pairs,), dtype=torch.float32, device=dev)
    ds = torch.empty((n_pairs,), dtype=torch.float32, device=dev)
    aux = torch.zeros((n_pairs, L.AUX_DOUBLES), dtype=torch.float64, device=dev)
    rc = ctx.lib.da3s_pair_thresholds(ctx.h, _ptr(pairs), n_pairs, overlap, H, W, C.byref(opts), _ptr(thr), _ptr(ds),
                                      _ptr(aux), _stream(pairs))
    L.check(rc, "da3s_pair_thresholds")
    return thr, ds, aux


def ransac_hypotheses(pairs, n_pairs, overlap, H, W, thr, ds, sample_idx, *, world=True, valid_depth=True, depth_eps=1e-6):
    dev = pairs.device
    ctx = context(dev)
    sample_idx = sample_idx.to(torch.int32).contiguous()
    n_hyp = sample_idx.shape[1]
    A = torch.empty((n_pairs, n_hyp, 9), dtype=torch.float32, device=dev)
    t = torch.empty((n_pairs, n_hyp, 3), dtype=torch.float32, device=dev)
    ok = torch.empty((n_pairs, n_hyp), dtype=torch.uint8, device=dev)
    s3 = torch.empty((n_pairs, n_hyp, 13), dtype=torch.float64, device=dev)
    rc = ctx.lib.da3s_ransac_hypotheses(ctx.h, _ptr(pairs), n_pairs, overlap, H, W, int(world), int(valid_depth), depth_eps,
                                        _ptr(thr), _ptr(ds), _ptr(sample_idx), n_hyp, _ptr(A), _ptr(t), _ptr(ok), _ptr(s3),
                                        _stream(pairs))
    L.check(rc, "da3s_ransac_hypotheses")
    return A, t, ok, s3


def ransac_score(pairs, n_pairs, overlap, H, W, thr, ds, A, t, ok, ransac_thr, *, world=True, valid_depth=True, depth_eps=1e-6):
    dev = pairs.device
    ctx = context(dev)
    n_hyp = A.shape[1]
    counts = torch.empty((n_pairs, n_hyp), dtype=torch.int32, device=dev)
    A, t, ok = A.contiguous(), t.contiguous(), ok.contiguous()      # named: must outlive the launch
    rc = ctx.lib.da3s_ransac_score(ctx.h, _ptr(pairs), n_pairs, overlap, H, W, int(world), int(valid_depth), depth_eps,
                                   _ptr(thr), _ptr(ds), _ptr(A), _ptr(t), _ptr(ok),
                                   n_hyp, float(ransac_thr), _ptr(counts), _stream(pairs))
    L.check(rc, "da3s_ransac_score")
    return counts


def ransac_inlier_mask(pairs, n_pairs, overlap, H, W, thr, ds, best_A, best_t, best_ok, ransac_thr, *, world=True,
                       valid_depth=True, depth_eps=1e-6):
    dev = pairs.device
    ctx = context(dev)
    mask = torch.empty((n_pairs, overlap * H * W), dtype=torch.uint8, device=dev)
    best_A, best_t, best_ok = best_A.contiguous(), best_t.contiguous(), best_ok.contiguous()   # must outlive the launch
    rc = ctx.lib.da3s_ransac_inlier_mask(ctx.h, _ptr(pairs), n_pairs, overlap, H, W, int(world), int(valid_depth), depth_eps,
                                         _ptr(thr), _ptr(ds), _ptr(best_A), _ptr(best_t),
                                         _ptr(best_ok), float(ransac_thr), _ptr(mask), _stream(pairs))
    L.check(rc, "da3s_ransac_inlier_mask")
    return mask.view(torch.bool)


# ------------------------------------------------------------------------------------------
# materialised correspondences
# ------------------------------------------------------------------------------------------
def umeyama_points(src, dst, weights=None, variant=L.UMEYAMA_WEIGHTED, idx_src=None, idx_dst=None) -> torch.Tensor:
    src = src.contiguous()
    dst = dst.contiguous()
    assert src.dtype == dst.dtype and src.dtype in (torch.float32, torch.float64)
    n = src.numel() // 3
    ctx = context(src.device)
    row = torch.zeros((L.ROW_LEN,), dtype=torch.float64, device=src.device)
    w64 = 0
    if weights is not None:
        weights = weights.contiguous()
        assert weights.dtype in (torch.float32, torch.float64)
        w64 = int(weights.dtype == torch.float64)
    n_idx = 0
    if idx_src is not None:
        idx_src = idx_src.to(torch.int64).contiguous()
        idx_dst = idx_dst.to(torch.int64).contiguous()
        n_idx = idx_src.numel()
    rc = ctx.lib.da3s_umeyama_points(ctx.h, _ptr(src), _ptr(dst), int(src.dtype == torch.float64), _ptr(weights), w64, n,
                                     _ptr(idx_src), _ptr(idx_dst), n_idx, variant, _ptr(row), _stream(src))
    L.check(rc, "da3s_umeyama_points")
    return row


def icp_points(src, dst, threshold: float, max_iterations: int, mode=L.ICP_SIM3, cell_size=None) -> torch.Tensor:
    """Nearest-neighbour registration of two UNORDERED clouds src [n,3] -> dst [m,3] (da3s_icp_points):
    mode ICP_SIM3 = align_geometry.py:84-140, ICP_RIGID = Open3D point-to-point ICP.  Returns the [16] row."""
    src = src.contiguous()
    dst = dst.contiguous()
    assert src.dtype == dst.dtype and src.dtype in (torch.float32, torch.float64)
    n, m = src.numel() // 3, dst.numel() // 3
    need = 4 * m * 12 + m * 4 + (64 << 20)
    ctx = context(src.device, need if need > (1 << 30) else None)
    row = torch.zeros((L.ROW_LEN,), dtype=torch.float64, device=src.device)
    if cell_size is None:
        # grid cells of about twice the target's point spacing (a surface sampled by m points over its bounding box)
        d = dst.view(-1, 3)
        fin = torch.isfinite(d).all(dim=1)
        ext = (d[fin].amax(dim=0) - d[fin].amin(dim=0)).max().item() if bool(fin.any()) else 0.0
        cell_size = 2.0 * ext / max(1.0, math.sqrt(m)) if ext > 0 else 0.0
    rc = ctx.lib.da3s_icp_points(ctx.h, _ptr(src), n, _ptr(dst), m, int(src.dtype == torch.float64), int(mode), float(threshold),
                                 float(cell_size), int(max_iterations), _ptr(row), _stream(src))
    L.check(rc, "da3s_icp_points")
    return row


def irls_points(src, dst, conf_src, conf_dst, idx_src=None, idx_dst=None, delta=1.0, max_iterations=20, tol=1e-6):
    src = src.contiguous()
    dst = dst.contiguous()
    assert src.dtype == dst.dtype and src.dtype in (torch.float32, torch.float64)
    n = src.numel() // 3
    ctx = context(src.device)
    row = torch.zeros((L.ROW_LEN,), dtype=torch.float64, device=src.device)
    n_idx = 0
    if idx_src is not None:
        idx_src = idx_src.to(torch.int64).contiguous()
        idx_dst = idx_dst.to(torch.int64).contiguous()
        n_idx = idx_src.numel()
    conf_src, conf_dst = conf_src.contiguous(), conf_dst.contiguous()
    rc = ctx.lib.da3s_irls_points(ctx.h, _ptr(src), _ptr(dst), int(src.dtype == torch.float64),
                                  _ptr(conf_src, "conf_src", torch.float32),
                                  _ptr(conf_dst, "conf_dst", torch.float32), n, _ptr(idx_src), _ptr(idx_dst),
                                  n_idx, float(delta), int(max_iterations), float(tol), _ptr(row), _stream(src))
    L.check(rc, "da3s_irls_points")
    return row


# ------------------------------------------------------------------------------------------
# voxel grid
# ------------------------------------------------------------------------------------------
class VoxelGrid:
    """Stage-level handle on the context's hash grid: begin() / insert() / finish() only enqueue
    kernels; read() synchronises and trims the outputs."""

    def __init__(self, device, table_slots: int, max_voxels: int, with_rgb: bool, private_ctx: bool = False):
        assert table_slots & (table_slots - 1) == 0
        device = torch.device(device)
        self.device = device
        self.table_slots = table_slots
        self.max_voxels = max_voxels
        need = table_slots * 66 + (256 << 20)
        if private_ctx:                  # its own da3s_ctx (= its own table): several grids on one device
            with torch.cuda.device(device):
                self.ctx = Context(device, need)
        else:
            self.ctx = context(device, need)
        self.xyz = torch.empty((max_voxels, 3), dtype=torch.float32, device=device)
        self.rgb = torch.empty((max_voxels, 3), dtype=torch.uint8, device=device) if with_rgb else None
        self.count = torch.empty((max_voxels,), dtype=torch.int32, device=device)
        self.key = torch.empty((max_voxels,), dtype=torch.int64, device=device)
        self.nv = torch.zeros((2,), dtype=torch.int64, device=device)      # [n_voxels, n_dropped]

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def begin(self):
        L.check(self.ctx.lib.da3s_voxel_begin(self.ctx.h, self.table_slots, self._st()), "da3s_voxel_begin")

    def insert(self, xyz, rgb, mask, voxel):
        n = xyz.numel() // 3
        rc = self.ctx.lib.da3s_voxel_insert(self.ctx.h, _ptr(xyz, "xyz", torch.float32), _ptr(rgb, "rgb", torch.uint8),
                                            _ptr(mask, "mask"), n, float(voxel), self._st())
        L.check(rc, "da3s_voxel_insert")

    def make_jobs(self, clouds):
        """clouds: list of (xyz, rgb or None, mask or None) CUDA tensors that stay alive.  Returns the device job
        table (and the largest n) for insert_jobs: every cloud goes into the grid in ONE launch."""
        arr = (L.VoxelJob * len(clouds))()
        max_n = 0
        for i, (xyz, rgb, mask) in enumerate(clouds):
            arr[i].xyz = _ptr(xyz, "xyz", torch.float32).value
            arr[i].rgb = _ptr(rgb, "rgb", torch.uint8).value if rgb is not None else None
            arr[i].mask = _ptr(mask, "mask").value if mask is not None else None
            arr[i].n = xyz.numel() // 3
            max_n = max(max_n, arr[i].n)
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        return torch.from_numpy(raw).to(self.device), len(clouds), max_n

    def insert_jobs(self, jobs, voxel, width=0):
        """width > 0: the clouds are image sequences with rows of `width` points (patch-wise traversal)."""
        table, n_jobs, max_n = jobs
        rc = self.ctx.lib.da3s_voxel_insert_jobs(self.ctx.h, _ptr(table), n_jobs, max_n, int(width), float(voxel), self._st())
        L.check(rc, "da3s_voxel_insert_jobs")

    def make_export_jobs(self, frames):
        """frames: list of dicts(depth, conf, cam, sim3, conf_thr, rgb) of CUDA tensors / None that stay alive
        (one per image frame).  Returns the device job table for insert_frames."""
        arr = (L.ExportJob * len(frames))()
        for i, j in enumerate(frames):
            for name in ("depth", "conf", "cam", "sim3", "conf_thr", "rgb"):
                t = j.get(name)
                if t is None:
                    continue
                if not t.is_cuda or not t.is_contiguous():
                    raise RuntimeError("export job tensors must be contiguous CUDA tensors")
                if name in ("depth", "conf") and t.data_ptr() % 16:
                    raise L.Da3sError(L.EALIGN, "make_export_jobs", f"{name} is not 16-byte aligned")
                setattr(arr[i], name, t.data_ptr())
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        return torch.from_numpy(raw).to(self.device), len(frames)

    def insert_frames(self, jobs, H, W, voxel, *, world=True, conf_cmp=None, conf_thr=0.0, conf_floor=None, depth_eps=None):
        """Fused unproject (fast float32) + Sim(3) + filter + insert of every frame of the job table: the grid
        receives exactly what unproject_filter_jobs + insert_jobs would give it, the points are never written."""
        table, n_frames = jobs
        flags = L.UNPROJ_FAST | (L.UNPROJ_WORLD if world else 0)
        flags |= {None: 0, ">": L.MASK_CONF_GT, ">=": L.MASK_CONF_GE}[conf_cmp]
        flags |= (L.MASK_CONF_FLOOR if conf_floor is not None else 0) | (L.MASK_DEPTH if depth_eps is not None else 0)
        rc = self.ctx.lib.da3s_unproject_voxel_jobs(self.ctx.h, _ptr(table), n_frames, H, W, flags, float(conf_thr),
                                                    float(conf_floor or 0.0), float(depth_eps or 0.0), float(voxel), self._st())
        L.check(rc, "da3s_unproject_voxel_jobs")

    def send(self, world, rank, inboxes, counts, flags, step, cap):
        """Multi-GPU merge, step 1: compact this rank's table and write every record into the inbox of the rank
        that owns its key.  inboxes[d] / counts[d] / flags[d]: rank d's [world, cap, 6] int64 inbox, [world] int64 record
        counts and [world] int64 arrival flags (this step's half) as tensors mapped in THIS process (own tensors for
        d == rank, CUDA-IPC mappings for the peers).  After its stores the kernel writes `step` into flags[d][rank]."""
        ip = (C.c_void_p * world)(*[t.data_ptr() for t in inboxes])
        cp = (C.c_void_p * world)(*[t.data_ptr() for t in counts])
        fp = (C.c_void_p * world)(*[t.data_ptr() for t in flags])
        rc = self.ctx.lib.da3s_voxel_send(self.ctx.h, world, rank, ip, cp, fp, int(step), cap, self._st())
        L.check(rc, "da3s_voxel_send")

    def merge_inbox(self, inbox, counts, world, cap, flags=None, step=0):
        """Multi-GPU merge, step 2: wait on the device until every rank's records of `step` have arrived (flags given),
        then fold the own inbox into the table."""
        rc = self.ctx.lib.da3s_voxel_merge_inbox(self.ctx.h, _ptr(inbox), _ptr(counts), _ptr(flags), int(step), world, cap, self._st())
        L.check(rc, "da3s_voxel_merge_inbox")

    def finish(self, voxel):
        rc = self.ctx.lib.da3s_voxel_finish(self.ctx.h, float(voxel), self.max_voxels, _ptr(self.xyz), _ptr(self.rgb),
                                            _ptr(self.count), _ptr(self.key), C.c_void_p(self.nv.data_ptr()),
                                            C.c_void_p(self.nv.data_ptr() + 8), self._st())
        L.check(rc, "da3s_voxel_finish")

    def read(self, sort=False):
        n, dropped = (int(v) for v in self.nv.cpu())
        if dropped:
            raise L.Da3sError(L.ENOMEM, "VoxelGrid", f"hash table full: {dropped} points dropped")
        if n > self.max_voxels:
            raise L.Da3sError(L.ENOMEM, "VoxelGrid", f"{n} voxels > max_voxels {self.max_voxels}")
        xyz, cnt, key = self.xyz[:n], self.count[:n], self.key[:n]
        rgb = self.rgb[:n] if self.rgb is not None else None
        if sort:
            order = torch.argsort(key)
            xyz, cnt, key = xyz[order], cnt[order], key[order]
            rgb = rgb[order] if rgb is not None else None
        return xyz, rgb, cnt, key


def voxel_downsample(clouds, voxel: float, table_slots: int | None = None, max_voxels: int | None = None, sort=True):
    """clouds: list of (xyz [n,3] f32, rgb [n,3] u8 or None, mask [n] bool/u8 or None) CUDA tensors, all
    accumulated into one grid.  Returns (xyz [m,3] f32, rgb [m,3] u8 or None, count [m] i32, key [m] i64)."""
    dev = clouds[0][0].device
    total = sum(c[0].numel() // 3 for c in clouds)
    if table_slots is None:
        table_slots = 1 << max(10, int(math.ceil(math.log2(max(2 * total, 1024)))))
        table_slots = min(table_slots, 1 << 28)
    need = table_slots * 66 + (64 << 20)
    ctx = context(dev, need)
    if max_voxels is None:
        max_voxels = min(total, table_slots)
    has_rgb = clouds[0][1] is not None
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(ctx.lib.da3s_voxel_begin(ctx.h, table_slots, st), "da3s_voxel_begin")
    for xyz, rgb, mask in clouds:
        xyz = xyz.contiguous().view(-1, 3)
        if mask is not None:
            mask = mask.contiguous().view(-1).view(torch.uint8)
        if rgb is not None:
            rgb = rgb.contiguous()
        rc = ctx.lib.da3s_voxel_insert(ctx.h, _ptr(xyz, "xyz", torch.float32), _ptr(rgb),
                                       _ptr(mask), xyz.shape[0], float(voxel), st)
        L.check(rc, "da3s_voxel_insert")
    out_xyz = torch.empty((max_voxels, 3), dtype=torch.float32, device=dev)
    out_rgb = torch.empty((max_voxels, 3), dtype=torch.uint8, device=dev) if has_rgb else None
    out_cnt = torch.empty((max_voxels,), dtype=torch.int32, device=dev)
    out_key = torch.empty((max_voxels,), dtype=torch.int64, device=dev)
    nv = torch.zeros((2,), dtype=torch.int64, device=dev)
    rc = ctx.lib.da3s_voxel_finish(ctx.h, float(voxel), max_voxels, _ptr(out_xyz), _ptr(out_rgb), _ptr(out_cnt), _ptr(out_key),
                                   C.c_void_p(nv.data_ptr()), C.c_void_p(nv.data_ptr() + 8), st)
    L.check(rc, "da3s_voxel_finish")
    n, dropped = (int(v) for v in nv.cpu())
    if dropped:
        raise L.Da3sError(L.ENOMEM, "voxel_downsample", f"hash table full: {dropped} points dropped")
    if n > max_voxels:
        raise L.Da3sError(L.ENOMEM, "voxel_downsample", f"{n} voxels > max_voxels {max_voxels}")
    out_xyz, out_cnt, out_key = out_xyz[:n], out_cnt[:n], out_key[:n]
    if out_rgb is not None:
        out_rgb = out_rgb[:n]
    if sort:
        order = torch.argsort(out_key)     # canonical order = ascending packed key (glue, not the hot path)
        out_xyz, out_cnt, out_key = out_xyz[order], out_cnt[order], out_key[order]
        if out_rgb is not None:
            out_rgb = out_rgb[order]
    return out_xyz, out_rgb, out_cnt, out_key
