"""numpy-in / numpy-out entry points: what the reference-named shim modules at the repo
root (align_geometry.py, utils/align.py, utils/geometry.py, ...) call.

Each function stages its numpy inputs into (pinned ->) device memory, runs kernels of
libda3s.so and reads the result back; argument meaning, return types and the error
conventions of the reference function it stands in for are kept (cited per function).
There is no CPU fallback: without CUDA these raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops
from .pipeline import DeviceSubmap, pair_entry, rows_to_sim3


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("da3slam_b200: CUDA device required (no CPU fallback for the alignment path)")
    return torch.device(f"cuda:{torch.cuda.current_device()}")


def _up(x, dtype=None):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(_device(), non_blocking=True).contiguous()


def _row_to_tuple(row):
    r = row.cpu().numpy()
    return float(r[0]), r[1:10].reshape(3, 3).copy(), r[10:13].copy()


# ------------------------------------------------------------------------------------------
# unprojection
# ------------------------------------------------------------------------------------------
def unproject(depth, intrinsics, extrinsics, *, world: bool, out_f64: bool, mode: str = "kinv",
              general_inverse: bool = True):
    """depth [N,H,W], K [N,3,3], w2c [N,3,4] -> [N,H,W,3] numpy.
    mode 'kinv' + general_inverse follows utils/geometry.py:4-40 / align_geometry.py:192-256
    (K^-1 and the full matrix inverse); mode 'closed' follows src/vggt/utils/geometry.py:14-116."""
    d = _up(depth, torch.float32)
    cams = ops.build_cams(_up(intrinsics, torch.float32), _up(extrinsics, torch.float32), general_inverse)
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode=mode, world=world, out_f64=out_f64, want_mask=False, want_count=False)
    return xyz.cpu().numpy()


def apply_sim3(points, s, R, t):
    """utils/geometry.py:43-70: float64 out for any float input, same shape."""
    pts = np.asarray(points)
    p = _up(pts if pts.dtype in (np.float32, np.float64) else pts.astype(np.float64))
    row = ops.sim3_row(s, R, t, p.device)
    return ops.apply_sim3(p, row, out_f64=True).cpu().numpy()


# ------------------------------------------------------------------------------------------
# depth scale (align_geometry.py:307-330, utils/align_geometry_single.py:31-49)
# ------------------------------------------------------------------------------------------
def _get(pred, key):
    return pred[key] if isinstance(pred, dict) else getattr(pred, key)


def _has(pred, key):
    return (key in pred) if isinstance(pred, dict) else hasattr(pred, key)


def depth_scale(prev, cur, conf_th=0.2, eps=1e-6, guarded=False) -> float:
    d_prev = _up(np.asarray(_get(prev, "depth"))[-1], torch.float32)
    d_cur = _up(np.asarray(_get(cur, "depth"))[0], torch.float32)
    seg = dict(a=d_prev, b=d_cur, kind=L.SEL_RATIO, stat=L.SEL_MEDIAN, conf_th=float(conf_th), eps=float(eps))
    if _has(prev, "conf") and _has(cur, "conf"):
        seg["ca"] = _up(np.asarray(_get(prev, "conf"))[-1], torch.float32)
        seg["cb"] = _up(np.asarray(_get(cur, "conf"))[0], torch.float32)
    out = ops.select([seg], d_prev.device)[0]
    if guarded:
        if out["n_valid"] < 50:
            return 1.0
        s = out["value"]
        if not np.isfinite(s) or s <= 0:
            return 1.0
        return float(s)
    return float(out["value"])          # nan for an empty mask, as np.median gives


# ------------------------------------------------------------------------------------------
# Umeyama family on arrays
# ------------------------------------------------------------------------------------------
def _pts(x):
    x = np.asarray(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    return x.reshape(-1, 3)


def umeyama(src, dst, weights=None, variant=L.UMEYAMA_WEIGHTED):
    """variant WEIGHTED: utils/align.py:14-40; MEAN: align_geometry.py:59-82;
    NORMRATIO: utils/align.py:224-276.  target ~= s R source + t."""
    a, b = _pts(src), _pts(dst)
    if a.dtype != b.dtype:
        a, b = a.astype(np.float64), b.astype(np.float64)
    w = None
    if weights is not None:
        w = np.asarray(weights)
        w = _up(w if w.dtype in (np.float32, np.float64) else w.astype(np.float64))
    return _row_to_tuple(ops.umeyama_points(_up(a), _up(b), w, variant))


def icp(source, target, threshold, max_iterations, rigid=False):
    """Nearest-neighbour registration target ~= s R source + t of two unordered clouds [n,3], [m,3]:
    rigid=False -> align_geometry.py:84-140 (KD-tree Umeyama loop); rigid=True -> Open3D point-to-point ICP
    (align_geometry.py:29-45, utils/align_geometry_single.py:146-160), s == 1."""
    a, b = _pts(source), _pts(target)
    if a.dtype != b.dtype:
        a, b = a.astype(np.float64), b.astype(np.float64)
    row = ops.icp_points(_up(a), _up(b), threshold, max_iterations, L.ICP_RIGID if rigid else L.ICP_SIM3)
    return _row_to_tuple(row)


def irls_pixel(point_map1, point_map2, conf1, conf2, min_points=100, max_iterations=20,
               convergence_threshold=1e-6, delta=1.0, compat="reference", indices=None):
    """utils/align.py:111-218.  compat='reference' keeps the reference's behaviour bit for
    bit in its discrete decisions: float32 threshold from the two medians, masks applied
    INDEPENDENTLY to the two clouds, <= 5000 pairs drawn with np.random.choice from the
    global numpy RNG (pass `indices` to fix the draw).  compat='joint' uses the joint mask
    and every pixel (the upstream VGGT-Long behaviour).  Returns (s, R, t): map2 -> map1."""
    p1, p2 = _pts(point_map1), _pts(point_map2)
    if p1.dtype != p2.dtype:
        p1, p2 = p1.astype(np.float64), p2.astype(np.float64)
    c1 = np.ascontiguousarray(np.asarray(conf1, np.float32).reshape(-1))
    c2 = np.ascontiguousarray(np.asarray(conf2, np.float32).reshape(-1))
    d1, d2, dc1, dc2 = _up(p1), _up(p2), _up(c1), _up(c2)
    med = ops.select([dict(a=dc1, stat=L.SEL_MEDIAN), dict(a=dc2, stat=L.SEL_MEDIAN)], d1.device)["value"]
    thr = min(med[0], med[1]) * np.float32(0.1)
    m1 = dc1 > float(thr)               # float32 compare: float(thr) is exactly the float32 value
    m2 = dc2 > float(thr)
    if compat == "joint":
        m1 = m2 = m1 & m2
    nz1 = torch.nonzero(m1).view(-1)    # index bookkeeping for the reference's boolean indexing
    nz2 = torch.nonzero(m2).view(-1)
    n1, n2 = int(nz1.numel()), int(nz2.numel())
    if n1 < min_points or n2 < min_points:
        print(f"  Warning: Not enough points for alignment: {n1} vs {n2}")
        return 1.0, np.eye(3), np.zeros(3)
    if compat == "reference":
        k = min(5000, n1, n2)
        if indices is None:
            indices = np.random.choice(min(n1, n2), k, replace=False)
        idx = _up(np.asarray(indices, np.int64))
        i1, i2 = nz1[idx], nz2[idx]
    else:
        i1, i2 = nz1, nz2
    row = ops.irls_points(d2, d1, dc2, dc1, idx_src=i2, idx_dst=i1, delta=delta, max_iterations=max_iterations,
                          tol=convergence_threshold)
    return _row_to_tuple(row)


# ------------------------------------------------------------------------------------------
# submap pairs from host predictions (the end-to-end path bench.py times)
# ------------------------------------------------------------------------------------------
def align_pairs_host_arrays(dA, cA, KA, EA, dB, cB, KB, EB, sample_idx=None, **opt_kw):
    """da3s_align_pairs_host on prepared HOST arrays (pageable numpy, or numpy views of pinned torch tensors):
    depth/conf [n,overlap,H,W] float32 per side, K [n,overlap,3,3], E [n,overlap,3,4] float32.  The library copies
    them in, runs the batched device pipeline, copies the [n,16] float64 rows out and synchronises."""
    import ctypes as C
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (dA, cA, KA, EA, dB, cB, KB, EB)]
    dA, cA, KA, EA, dB, cB, KB, EB = arrs
    n, o, H, W = dA.shape
    for a, shp in zip(arrs, [(n, o, H, W), (n, o, H, W), (n, o, 3, 3), (n, o, 3, 4)] * 2):
        if a.shape != shp:
            raise ValueError(f"align_pairs_host_arrays: expected shape {shp}, got {a.shape}")
    opts = L.default_opts(**opt_kw)
    need = 4 * dA.nbytes + (256 << 20)
    dev = _device()
    ctx = ops.context(dev, need)
    rows = np.empty((n, L.ROW_LEN), np.float64)
    si = None
    if opts.n_hyp > 0:
        if sample_idx is None:
            raise ValueError("n_hyp > 0 needs sample_idx [n_pairs, n_hyp, 3]")
        si = np.ascontiguousarray(sample_idx, np.int32)
        if si.shape != (n, opts.n_hyp, 3):
            raise ValueError(f"sample_idx must be [{n}, {opts.n_hyp}, 3], got {si.shape}")
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = C.c_void_p
    rc = ctx.lib.da3s_align_pairs_host(ctx.h, n, o, H, W, P(dA.ctypes.data), P(cA.ctypes.data), P(KA.ctypes.data),
                                       P(EA.ctypes.data), P(dB.ctypes.data), P(cB.ctypes.data), P(KB.ctypes.data),
                                       P(EB.ctypes.data), C.byref(opts), P(si.ctypes.data if si is not None else 0),
                                       P(rows.ctypes.data), st)
    L.check(rc, "da3s_align_pairs_host")
    return rows


def align_prediction_pairs(pairs, overlap=1, sample_idx=None, **opt_kw):
    """pairs: list of (prev_prediction, cur_prediction) with HOST numpy fields.  Goes through
    da3s_align_pairs_host: host->device copies, the batched device pipeline, rows back.
    Returns [n,16] float64 rows (include/da3s.h DA3S_ROW_*)."""
    o = overlap

    def stack(which, key, sl):
        return np.ascontiguousarray(np.stack([np.asarray(_get(p[which], key))[sl] for p in pairs]), dtype=np.float32)
    tail, head = slice(-o, None), slice(0, o)
    dA, cA, KA, EA = (stack(0, k, tail) for k in ("depth", "conf", "intrinsics", "extrinsics"))
    dB, cB, KB, EB = (stack(1, k, head) for k in ("depth", "conf", "intrinsics", "extrinsics"))
    return align_pairs_host_arrays(dA, cA, KA, EA, dB, cB, KB, EB, sample_idx=sample_idx, **opt_kw)


def align_two_predictions(prev, cur, overlap=1, **opt_kw):
    rows = align_prediction_pairs([(prev, cur)], overlap, **opt_kw)
    return rows_to_sim3(rows)[0], rows[0]
