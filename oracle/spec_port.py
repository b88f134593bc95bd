"""CPU statement of the parts of the hot path the reference does NOT contain.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference
has no fused f32 unprojection, no dense joint-mask IRLS, no RANSAC and no voxel
grid, so nothing in /root/reference pins these.
The normative text is oracle/SPEC.md; this file is its executable form and the
CUDA kernels are checked against it (bit-exact for masks / inliers / counts /
voxel keys, 1e-6 relative for Sim(3)).  Wherever a reference function exists it
is called from ref_port (weighted_umeyama, umeyama_sim3, apply_sim3, ...), so the
unpinned part is only the glue between pinned pieces.
"""
from __future__ import annotations

import numpy as np

from . import ref_port as rp

F32 = np.float32


# --------------------------------------------------------------------------
# SPEC 1 — fast f32 camera-frame unprojection (the per-pixel form the fused
# kernels use; numerically a twin of align_geometry.py:232-239 in float32)
# --------------------------------------------------------------------------
def cam_fast_f32(depth, intrinsics):
    """depth [N,H,W] f32, K [N,3,3] -> cam [N,H,W,3] f32 with every operation a
    single float32 rounding, in this order:
        x = ((f32(u) - cu) * d) * inv_fu,  inv_fu = f32(1) / fu
        y = ((f32(v) - cv) * d) * inv_fv,  inv_fv = f32(1) / fv
        z = d
    """
    depth = np.asarray(depth, dtype=F32)
    K = np.asarray(intrinsics, dtype=F32)
    N, H, W = depth.shape
    u = np.arange(W, dtype=F32)[None, None, :]
    v = np.arange(H, dtype=F32)[None, :, None]
    fu, fv = K[:, 0, 0][:, None, None], K[:, 1, 1][:, None, None]
    cu, cv = K[:, 0, 2][:, None, None], K[:, 1, 2][:, None, None]
    inv_fu = F32(1) / fu
    inv_fv = F32(1) / fv
    x = ((u - cu) * depth) * inv_fu
    y = ((v - cv) * depth) * inv_fv
    out = np.empty((N, H, W, 3), dtype=F32)
    out[..., 0] = x
    out[..., 1] = y
    out[..., 2] = depth
    return out


def c2w_closed_form(extrinsics):
    """[N,3,4] w2c (any float dtype) -> (R [N,3,3], t [N,3]) float64 of the
    camera-to-world transform, closed form [R^T | -R^T t] evaluated in float64
    (src/vggt/utils/geometry.py:119-168 semantics with float64 inputs)."""
    E = np.asarray(extrinsics, dtype=np.float64)
    Rt = np.transpose(E[:, :3, :3], (0, 2, 1))
    t = -np.einsum("nij,nj->ni", Rt, E[:, :3, 3])
    return Rt, t


def world_from_cam_f64(cam_f32, extrinsics):
    """world = R_c2w @ cam + t_c2w in float64 from float32 camera points."""
    R, t = c2w_closed_form(extrinsics)
    return np.einsum("nij,nhwj->nhwi", R, cam_f32.astype(np.float64)) + t[:, None, None, :]


def fmaf(a, b, c):
    """float32 fused multiply-add emulated through float64 (the product of two
    float32 is exact in float64; the single float64 addition can double-round in
    ~2^-29 of cases — the C oracle uses the real fmaf and is the arbiter)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def world_from_cam_f32(cam_f32, extrinsics):
    """SPEC 4 world points for RANSAC scoring: c2w rounded to float32, then per
    axis  w_i = fma(M_i0, x, fma(M_i1, y, fma(M_i2, z, m_i)))."""
    R, t = c2w_closed_form(extrinsics)
    M = R.astype(F32)
    m = t.astype(F32)
    x, y, z = cam_f32[..., 0], cam_f32[..., 1], cam_f32[..., 2]
    out = np.empty_like(cam_f32)
    for i in range(3):
        Mi0, Mi1, Mi2 = (M[:, i, j][:, None, None] for j in range(3))
        out[..., i] = fmaf(Mi0, x, fmaf(Mi1, y, fmaf(Mi2, z, m[:, i][:, None, None])))
    return out


# --------------------------------------------------------------------------
# SPEC 2 — pair correspondences: overlap frames, joint mask, weights
# --------------------------------------------------------------------------
def pair_correspondences(prev, cur, overlap=1, world=True, depth_scale=None,
                         valid_depth=True, eps=1e-6, thr=None):
    """prev/cur: dicts with depth [F,H,W] f32, conf [F,H,W] f32, intrinsics
    [F,3,3], extrinsics [F,3,4].  Returns a dict with
      y  [M,3] f64  target points  (prev submap, last `overlap` frames)
      x  [M,3] f64  source points  (cur submap, first `overlap` frames)
      xf, yf        the float32 camera-frame points (before any world transform)
      c  [M]  f32   sqrt(conf_prev * conf_cur) in float32
      mask [M] bool conf_prev > thr & conf_cur > thr (& depth validity)
      thr  f32      min(median(conf_prev), median(conf_cur)) * 0.1   (utils/align.py:140-142)
    Pixel i of prev overlap frame k corresponds to pixel i of cur overlap frame k.
    `depth_scale` (float32) multiplies the cur depth first (solver.py:125-126).
    """
    o = overlap
    dA = np.asarray(rp.field(prev, "depth"))[-o:].astype(F32)
    cA = np.asarray(rp.field(prev, "conf"))[-o:].astype(F32)
    KA = np.asarray(rp.field(prev, "intrinsics"))[-o:]
    EA = np.asarray(rp.field(prev, "extrinsics"))[-o:]
    dB = np.asarray(rp.field(cur, "depth"))[:o].astype(F32)
    cB = np.asarray(rp.field(cur, "conf"))[:o].astype(F32)
    KB = np.asarray(rp.field(cur, "intrinsics"))[:o]
    EB = np.asarray(rp.field(cur, "extrinsics"))[:o]
    if depth_scale is not None:
        dB = dB * F32(depth_scale)
    if thr is None:
        thr = rp.irls_conf_threshold(cA.reshape(-1), cB.reshape(-1))
    thr = F32(thr)
    mask = (cA > thr) & (cB > thr)
    if valid_depth:
        mask &= (dA > F32(eps)) & (dB > F32(eps)) & np.isfinite(dA) & np.isfinite(dB)
    yf = cam_fast_f32(dA, KA)
    xf = cam_fast_f32(dB, KB)
    if world:
        y = world_from_cam_f64(yf, EA)
        x = world_from_cam_f64(xf, EB)
    else:
        y = yf.astype(np.float64)
        x = xf.astype(np.float64)
    c = np.sqrt(cA * cB)
    return {"x": x.reshape(-1, 3), "y": y.reshape(-1, 3), "xf": xf.reshape(-1, 3),
            "yf": yf.reshape(-1, 3), "c": c.reshape(-1), "mask": mask.reshape(-1),
            "thr": thr, "EA": EA, "EB": EB, "shape": dA.shape}


# --------------------------------------------------------------------------
# SPEC 3 — dense IRLS (reference loop utils/align.py:169-211 without the 5000
# subsample, joint mask, configurable delta, optional extra gate mask)
# --------------------------------------------------------------------------
def irls_dense(x, y, c, mask, delta=1.0, max_iterations=20, tol=1e-6, min_points=100,
               huber=True):
    """x (source) -> y (target).  Returns (s, R, t, info).  With huber=False one
    confidence-weighted Umeyama solve is done (iteration count 1)."""
    n = int(mask.sum())
    info = {"n_valid": n, "iters": 0, "status": 0}
    if n < min_points:
        info["status"] = 1
        return 1.0, np.eye(3), np.zeros(3), info
    xs, ys, cs = x[mask], y[mask], c[mask]
    if not huber:
        s, R, t = rp.weighted_umeyama(xs, ys, cs)
        info["iters"] = 1
        return s, R, t, info
    s, R, t = 1.0, np.eye(3), np.zeros(3)
    for it in range(max_iterations):
        res = np.linalg.norm(ys - rp.apply_sim3(xs, s, R, t), axis=1)
        hub = np.ones_like(res)
        big = res > delta
        hub[big] = delta / res[big]
        w = cs * hub
        w = w / (np.max(w) + 1e-8)
        s_n, R_n, t_n = rp.weighted_umeyama(xs, ys, w)
        change = np.abs(s_n - s) + np.linalg.norm(R_n - R) + np.linalg.norm(t_n - t)
        s, R, t = s_n, R_n, t_n
        info["iters"] = it + 1
        info["mean_residual"] = float(np.mean(res))
        if change < tol:
            break
    return s, R, t, info


# --------------------------------------------------------------------------
# SPEC 4 — RANSAC: 3-point hypotheses + float32-FMA inlier scoring
# --------------------------------------------------------------------------
def ransac_points(corr, world):
    """float32 source/target points used for hypothesis generation and scoring."""
    if world:
        shp = corr["shape"] + (3,)
        xs = world_from_cam_f32(corr["xf"].reshape(shp), corr["EB"]).reshape(-1, 3)
        ys = world_from_cam_f32(corr["yf"].reshape(shp), corr["EA"]).reshape(-1, 3)
        return xs, ys
    return corr["xf"], corr["yf"]


def ransac_hypotheses(xs, ys, mask, sample_idx):
    """sample_idx [n_hyp,3] pixel indices.  A hypothesis is INVALID when any of its
    three pixels is masked out, when two of them coincide, or when the 3-point
    Umeyama (align_geometry.py:59-82 on the float32 points promoted to float64)
    is not finite.  Returns A [n_hyp,3,3] f32 (= f32(s*R)), t [n_hyp,3] f32,
    valid [n_hyp] bool, sim3 [n_hyp,13] f64 (s, R row-major, t)."""
    n_hyp = sample_idx.shape[0]
    A = np.zeros((n_hyp, 3, 3), F32)
    T = np.zeros((n_hyp, 3), F32)
    ok = np.zeros(n_hyp, bool)
    sim3 = np.zeros((n_hyp, 13), np.float64)
    for h in range(n_hyp):
        i = sample_idx[h]
        if not mask[i].all() or len(set(int(v) for v in i)) < 3:
            continue
        X = xs[i].astype(np.float64)
        Y = ys[i].astype(np.float64)
        s, R, t = rp.umeyama_sim3(X, Y)
        if not (np.isfinite(s) and np.isfinite(R).all() and np.isfinite(t).all()):
            continue
        sim3[h, 0] = s
        sim3[h, 1:10] = R.reshape(-1)
        sim3[h, 10:13] = t
        A[h] = (s * R).astype(F32)
        T[h] = t.astype(F32)
        ok[h] = True
    return A, T, ok, sim3


def residual2_f32(A, t, xs, ys):
    """Squared residual of one hypothesis at every correspondence, float32 FMA:
        p_i = fma(A_i0, x0, fma(A_i1, x1, fma(A_i2, x2, t_i)))
        d_i = p_i - y_i
        r2  = fma(d_0, d_0, fma(d_1, d_1, d_2 * d_2))"""
    x0, x1, x2 = xs[:, 0], xs[:, 1], xs[:, 2]
    d = []
    for i in range(3):
        p = fmaf(A[i, 0], x0, fmaf(A[i, 1], x1, fmaf(A[i, 2], x2, t[i])))
        d.append((p - ys[:, i]).astype(F32))
    return fmaf(d[0], d[0], fmaf(d[1], d[1], (d[2] * d[2]).astype(F32)))


def ransac_score(A, T, ok, xs, ys, mask, thr):
    """Inlier iff mask & r2 < f32(thr*thr) (strict, as align_geometry.py:123).
    Invalid hypotheses score 0.  Returns counts [n_hyp] int32."""
    thr2 = F32(float(thr) * float(thr))
    counts = np.zeros(A.shape[0], np.int32)
    for h in range(A.shape[0]):
        if not ok[h]:
            continue
        r2 = residual2_f32(A[h], T[h], xs, ys)
        counts[h] = int(np.count_nonzero(mask & (r2 < thr2)))
    return counts


def ransac_best(counts, ok, min_inliers=20):
    """argmax over valid hypotheses, ties to the lowest index; fewer than
    `min_inliers` (align_geometry.py:124) -> no model (-1)."""
    c = np.where(ok, counts, -1)
    best = int(np.argmax(c))
    if c[best] < min_inliers:
        return -1, int(max(c[best], 0))
    return best, int(c[best])


def ransac_inlier_mask(A, T, best, xs, ys, mask, thr):
    thr2 = F32(float(thr) * float(thr))
    if best < 0:
        return np.zeros_like(mask)
    return mask & (residual2_f32(A[best], T[best], xs, ys) < thr2)


# --------------------------------------------------------------------------
# SPEC 5 — voxel-grid downsample with exact integer accumulation
# --------------------------------------------------------------------------
VOX_BIAS = 1 << 20
VOX_FRAC = 4294967296.0  # 2^32


def voxel_keys(points_f32, voxel):
    """Per axis k = floor(f64(p) / f64(f32(voxel))) as int64; a point is usable
    when it is finite and |k| < 2^20 on every axis.  Returns (key [n] int64 packed
    21 bits/axis with bias 2^20, usable [n] bool, k [n,3] int64, frac [n,3] f64)."""
    p = np.asarray(points_f32, dtype=F32).astype(np.float64)
    v = np.float64(F32(voxel))
    finite = np.isfinite(p).all(axis=1)
    q = np.where(finite[:, None], p, 0.0) / v
    kf = np.floor(q)
    usable = finite & (np.abs(kf) < VOX_BIAS).all(axis=1)
    k = np.where(usable[:, None], kf, 0).astype(np.int64)
    frac = np.where(usable[:, None], q - kf, 0.0)
    key = ((k[:, 0] + VOX_BIAS) << 42) | ((k[:, 1] + VOX_BIAS) << 21) | (k[:, 2] + VOX_BIAS)
    return key, usable, k, frac


def voxel_downsample(points_f32, voxel, rgb=None, mask=None):
    """One output point per occupied voxel, sorted by packed key.
      position = f32((k + (sum_q / count) / 2^32) * f64(f32(voxel))),
                 sum_q = sum over points of llrint(frac * 2^32)   (exact int64)
      colour   = (2*sum_c + count) // (2*count)                    (round half up)
    Returns (xyz [m,3] f32, rgb [m,3] u8 or None, count [m] int32, key [m] int64)."""
    key, usable, k, frac = voxel_keys(points_f32, voxel)
    if mask is not None:
        usable = usable & np.asarray(mask, bool).reshape(-1)
    idx = np.flatnonzero(usable)
    if idx.size == 0:
        return (np.zeros((0, 3), F32), None if rgb is None else np.zeros((0, 3), np.uint8),
                np.zeros(0, np.int32), np.zeros(0, np.int64))
    order = idx[np.argsort(key[idx], kind="stable")]
    skey = key[order]
    starts = np.flatnonzero(np.concatenate(([True], skey[1:] != skey[:-1])))
    count = np.diff(np.concatenate((starts, [skey.size]))).astype(np.int64)
    q = np.rint(frac[order] * VOX_FRAC).astype(np.int64)
    sum_q = np.add.reduceat(q, starts, axis=0)
    kk = k[order][starts]
    v = np.float64(F32(voxel))
    mean_frac = sum_q.astype(np.float64) / count[:, None].astype(np.float64) / VOX_FRAC
    xyz = ((kk.astype(np.float64) + mean_frac) * v).astype(F32)
    out_rgb = None
    if rgb is not None:
        c = np.asarray(rgb, np.uint8).reshape(-1, 3)[order].astype(np.int64)
        sum_c = np.add.reduceat(c, starts, axis=0)
        out_rgb = ((2 * sum_c + count[:, None]) // (2 * count[:, None])).astype(np.uint8)
    return xyz, out_rgb, count.astype(np.int32), skey[starts]


# --------------------------------------------------------------------------
# SPEC 6 — whole pair alignment (what one "submap pair aligned" means)
# --------------------------------------------------------------------------
def align_pair(prev, cur, overlap=1, world=True, use_depth_scale=False, delta=1.0,
               max_iterations=20, tol=1e-6, min_points=100, huber=True,
               ransac=None, valid_depth=True):
    """ransac: None or dict(sample_idx=[n_hyp,3], thr=float, min_inliers=20).
    Returns dict(s, R, t, iters, n_valid, status, depth_scale, thr, ...)."""
    ds = None
    if use_depth_scale:
        ds = F32(rp.depth_scale_guarded(prev, cur))
    corr = pair_correspondences(prev, cur, overlap, world, ds, valid_depth)
    mask = corr["mask"]
    out = {"thr": corr["thr"], "depth_scale": 1.0 if ds is None else float(ds)}
    if ransac is not None:
        xs, ys = ransac_points(corr, world)
        A, T, ok, hyp = ransac_hypotheses(xs, ys, mask, ransac["sample_idx"])
        counts = ransac_score(A, T, ok, xs, ys, mask, ransac["thr"])
        best, nbest = ransac_best(counts, ok, ransac.get("min_inliers", 20))
        out.update(counts=counts, best=best, best_count=nbest, hyp_ok=ok, hyp_A=A, hyp_t=T, hyp_sim3=hyp)
        if best < 0:
            out.update(s=1.0, R=np.eye(3), t=np.zeros(3), iters=0, n_valid=0, status=2)
            return out
        mask = ransac_inlier_mask(A, T, best, xs, ys, mask, ransac["thr"])
        out["inlier_mask"] = mask
    s, R, t, info = irls_dense(corr["x"], corr["y"], corr["c"], mask, delta, max_iterations,
                               tol, min_points, huber)
    out.update(s=s, R=R, t=t, iters=info["iters"], n_valid=info["n_valid"], status=info["status"])
    return out


# --------------------------------------------------------------------------
# C-backed twins of the SPEC 4 loops (oracle/c/oracle.c, real fmaf): exact and fast
# enough for full-size inputs; the numpy forms above are the readable statement.
# --------------------------------------------------------------------------
def _c_arrays(A, T, ok, xs, ys, mask):
    A = np.ascontiguousarray(A, F32).reshape(-1, 9)
    T = np.ascontiguousarray(T, F32).reshape(-1, 3)
    ok = np.ascontiguousarray(ok, np.uint8)
    xs = np.ascontiguousarray(xs, F32)
    ys = np.ascontiguousarray(ys, F32)
    mask = np.ascontiguousarray(mask, np.uint8)
    return A, T, ok, xs, ys, mask


def ransac_score_c(A, T, ok, xs, ys, mask, thr):
    from . import build as ob
    lib = ob.load()
    A, T, ok, xs, ys, mask = _c_arrays(A, T, ok, xs, ys, mask)
    counts = np.zeros(A.shape[0], np.int32)
    lib.oracle_ransac_score(A.ctypes.data, T.ctypes.data, ok.ctypes.data, A.shape[0], xs.ctypes.data, ys.ctypes.data,
                            mask.ctypes.data, xs.shape[0], F32(float(thr) * float(thr)), counts.ctypes.data)
    return counts


def ransac_inlier_mask_c(A, T, best, xs, ys, mask, thr):
    from . import build as ob
    lib = ob.load()
    if best < 0:
        return np.zeros(len(mask), bool)
    A, T, _, xs, ys, mask = _c_arrays(A, T, np.ones(len(np.asarray(A).reshape(-1, 9)), np.uint8), xs, ys, mask)
    out = np.zeros(xs.shape[0], np.uint8)
    lib.oracle_ransac_inlier_mask(A[best].ctypes.data, T[best].ctypes.data, xs.ctypes.data, ys.ctypes.data,
                                  mask.ctypes.data, xs.shape[0], F32(float(thr) * float(thr)), out.ctypes.data)
    return out.astype(bool)


def cam_fast_f32_c(depth, intrinsics):
    from . import build as ob
    lib = ob.load()
    depth = np.ascontiguousarray(depth, F32)
    K = np.ascontiguousarray(intrinsics, F32)
    n, H, W = depth.shape
    out = np.empty((n, H, W, 3), F32)
    lib.oracle_cam_fast_f32(depth.ctypes.data, n, H, W, K.ctypes.data, out.ctypes.data)
    return out
