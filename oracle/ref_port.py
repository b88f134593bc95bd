"""numpy / torch-CPU restatement of the reference functions on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference file:line it restates (paths relative to /root/reference).  The
arithmetic — dtype promotions, operation order, epsilons, strictness of every
comparison — follows the reference exactly; only the code organisation is ours.
Pinned by tests/golden/*.npz (written from the real reference by
tests/golden/make_golden.py) and replayed by tests/test_oracle_golden.py.

Naming: ``prev`` = previous submap (target), ``cur`` = current submap (source);
every solver returns (s, R, t) with  target ~= s * R @ source + t.
"""
from __future__ import annotations

import numpy as np

try:  # torch is only needed for the float32 unprojection twin (row U1)
    import torch
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------
# helpers shared by the prediction containers (dict or attribute object)
# --------------------------------------------------------------------------
def field(pred, key):
    """utils/align_geometry_single.py:18-22 — dict or attribute access."""
    return pred[key] if isinstance(pred, dict) else getattr(pred, key)


def has_field(pred, key) -> bool:
    """utils/align_geometry_single.py:35-36 — `"conf" in d` / hasattr."""
    return (key in pred) if isinstance(pred, dict) else hasattr(pred, key)


def as_4x4(E3x4):
    """align_geometry.py:347-350, utils/align_geometry_single.py:8-11."""
    out = np.eye(4, dtype=np.float64)
    out[:3, :4] = E3x4
    return out


# --------------------------------------------------------------------------
# U2 — float64 numpy unprojection to WORLD   (utils/geometry.py:4-40)
# --------------------------------------------------------------------------
def unproject_world_f64(depth, intrinsics, extrinsics):
    """depth [N,H,W], K [N,3,3], w2c [N,3,4] -> world points [N,H,W,3] float64.

    utils/geometry.py:17-23  pixel grid (int64 u, v and float64 ones, concatenated
                             -> float64 homogeneous pixels)
    utils/geometry.py:26-28  cam = inv(K) . pix  (einsum), then * depth
    utils/geometry.py:32-38  4x4 float64 extrinsic, full matrix inverse, einsum
    """
    depth = np.asarray(depth)
    N, H, W = depth.shape
    pix = np.empty((N, H, W, 3), dtype=np.float64)
    pix[..., 0] = np.arange(W)[None, None, :]
    pix[..., 1] = np.arange(H)[None, :, None]
    pix[..., 2] = 1.0
    k_inv = np.linalg.inv(intrinsics)                      # keeps the input dtype
    cam = np.einsum("nij,nhwj->nhwi", k_inv, pix)
    cam = cam * depth[..., None]
    cam_h = np.concatenate([cam, np.ones((N, H, W, 1))], axis=-1)
    w2c = np.zeros((N, 4, 4))
    w2c[:, :3, :4] = extrinsics
    w2c[:, 3, 3] = 1.0
    c2w = np.linalg.inv(w2c)
    world_h = np.einsum("nij,nhwj->nhwi", c2w, cam_h)
    return world_h[..., :3]


# --------------------------------------------------------------------------
# U1 — float32 torch unprojection, camera or world
#      (align_geometry.py:192-256; twin utils/align_geometry_single.py:52-102)
# --------------------------------------------------------------------------
def unproject_f32(depth, intrinsics, extrinsics, in_coords="camera"):
    """numpy in -> numpy float32 out.  align_geometry.py:208-210 casts everything
    to float32; :232-239 pixel grid, torch.inverse(K), einsum, * depth;
    :242-251 4x4 w2c, torch.inverse, einsum (world only)."""
    assert in_coords in ("camera", "world")
    d = torch.tensor(np.asarray(depth), dtype=torch.float32)
    K = torch.tensor(np.asarray(intrinsics), dtype=torch.float32)
    E = torch.tensor(np.asarray(extrinsics), dtype=torch.float32)
    N, H, W = d.shape
    u = torch.arange(W).float().view(1, 1, W, 1).expand(N, H, W, 1)
    v = torch.arange(H).float().view(1, H, 1, 1).expand(N, H, W, 1)
    one = torch.ones((N, H, W, 1))
    pix = torch.cat([u, v, one], dim=-1)
    cam = torch.einsum("nij,nhwj->nhwi", torch.inverse(K), pix) * d.unsqueeze(-1)
    if in_coords == "camera":
        return cam.numpy()
    w2c = torch.zeros(N, 4, 4)
    w2c[:, :3, :4] = E
    w2c[:, 3, 3] = 1.0
    world_h = torch.einsum("nij,nhwj->nhwi", torch.inverse(w2c), torch.cat([cam, one], dim=-1))
    return world_h[..., :3].numpy()


# --------------------------------------------------------------------------
# U3 — VGGT closed-form unprojection (src/vggt/utils/geometry.py:14-168)
# --------------------------------------------------------------------------
def se3_inverse_closed_form(se3):
    """src/vggt/utils/geometry.py:119-168 (numpy branch): [R^T | -R^T t], the
    4x4 filled into a float64 identity (np.tile(np.eye(4)) :159)."""
    se3 = np.asarray(se3)
    R = se3[:, :3, :3]
    T = se3[:, :3, 3:]
    Rt = np.transpose(R, (0, 2, 1))
    out = np.tile(np.eye(4), (len(R), 1, 1))
    out[:, :3, :3] = Rt
    out[:, :3, 3:] = -np.matmul(Rt, T)
    return out


def cam_points_closed_form(depth_hw, K):
    """src/vggt/utils/geometry.py:86-116: x=(u-cu)*d/fu, y=(v-cv)*d/fv, z=d with
    int64 grids and numpy scalar intrinsics (=> float64 intermediates under NEP-50),
    stacked and cast to float32 (:114).  Zero skew is asserted (:99)."""
    H, W = depth_hw.shape
    assert K.shape == (3, 3)
    assert K[0, 1] == 0 and K[1, 0] == 0
    fu, fv, cu, cv = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    x = (u - cu) * depth_hw / fu
    y = (v - cv) * depth_hw / fv
    return np.stack((x, y, depth_hw), axis=-1).astype(np.float32)


def unproject_world_vggt(depth, extrinsics, intrinsics, eps=1e-8):
    """src/vggt/utils/geometry.py:14-43 + :46-83.  Returns (world [S,H,W,3] float64,
    cam [S,H,W,3] float32, valid [S,H,W] bool).  The reference wrapper returns only
    the world array; the other two are what depth_to_world_coords_points gives."""
    depth = np.asarray(depth)
    if depth.ndim == 4:
        depth = depth[..., 0]
    world, cams, valid = [], [], []
    for f in range(depth.shape[0]):
        d = depth[f]
        cam = cam_points_closed_form(d, intrinsics[f])
        c2w = se3_inverse_closed_form(extrinsics[f][None])[0]
        world.append(np.dot(cam, c2w[:3, :3].T) + c2w[:3, 3])          # :80
        cams.append(cam)
        valid.append(d > eps)                                           # :67
    return np.stack(world, 0), np.stack(cams, 0), np.stack(valid, 0)


# --------------------------------------------------------------------------
# D — depth-scale median between the overlap frames
# --------------------------------------------------------------------------
def _depth_scale_mask(prev, cur, conf_th, eps):
    d_prev = field(prev, "depth")[-1]
    d_cur = field(cur, "depth")[0]
    m = (d_prev > eps) & (d_cur > eps) & np.isfinite(d_prev) & np.isfinite(d_cur)
    if has_field(prev, "conf") and has_field(cur, "conf"):
        m &= (field(prev, "conf")[-1] > conf_th) & (field(cur, "conf")[0] > conf_th)
    return d_prev, d_cur, m


def depth_scale_plain(prev, cur, conf_th=0.2, eps=1e-6) -> float:
    """align_geometry.py:307-330: median of d_prev/d_cur over the joint mask,
    no guards (an empty mask gives nan, as numpy does)."""
    d_prev, d_cur, m = _depth_scale_mask(prev, cur, conf_th, eps)
    return float(np.median(d_prev[m] / d_cur[m]))


def depth_scale_guarded(prev, cur, conf_th=0.2, eps=1e-6) -> float:
    """utils/align_geometry_single.py:31-49: fewer than 50 valid pixels -> 1.0;
    non-finite or non-positive median -> 1.0."""
    d_prev, d_cur, m = _depth_scale_mask(prev, cur, conf_th, eps)
    if m.sum() < 50:
        return 1.0
    s = np.median(d_prev[m] / d_cur[m])
    if not np.isfinite(s) or s <= 0:
        return 1.0
    return float(s)


# --------------------------------------------------------------------------
# W — weighted Umeyama (utils/align.py:14-40) and its legacy twin (:42-92)
# --------------------------------------------------------------------------
def weighted_umeyama(src, dst, w):
    """utils/align.py:14-40.  eps=1e-8 in the weight normaliser (:17) and in the
    variance denominator (:37); reflection decided on det(U @ Vt) (:30)."""
    eps = 1e-8
    w = w.astype(np.float64)
    w = w / (np.sum(w) + eps)
    wc = w[:, None]
    mu_s = np.sum(src * wc, axis=0)
    mu_d = np.sum(dst * wc, axis=0)
    X = src - mu_s
    Y = dst - mu_d
    cov = (Y * wc).T @ X
    U, S, Vt = np.linalg.svd(cov)
    D = np.eye(3)
    if np.linalg.det(U @ Vt) < 0:
        D[2, 2] = -1.0
    R = U @ D @ Vt
    var_s = np.sum(w * np.sum(X * X, axis=1))
    s = float((S @ np.diag(D)) / (var_s + eps))
    t = mu_d - s * (R @ mu_s)
    return s, R, t


def weighted_umeyama_legacy(p1, p2, weights):
    """utils/align.py:42-92 (scale from trace(S), Vt-row flip): kept because it is
    public API; wrong for rotated data, as SURVEY.md section 8a notes."""
    weights = weights / (np.sum(weights) + 1e-8)
    c1 = np.sum(p1 * weights[:, None], axis=0)
    c2 = np.sum(p2 * weights[:, None], axis=0)
    a = p1 - c1
    b = p2 - c2
    S = np.dot(b.T, np.dot(np.diag(weights), a))
    U, _, Vt = np.linalg.svd(S)
    R = np.dot(U, Vt)
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = np.dot(U, Vt)
    s = np.trace(S) / (np.sum(weights * np.sum(a ** 2, axis=1)) + 1e-8)
    t = c2 - s * np.dot(R, c1)
    return s, R, t


def huber_weight(residual: float, delta: float = 1.0) -> float:
    """utils/align.py:94-109."""
    a = abs(residual)
    return 1.0 if a <= delta else delta / a


# --------------------------------------------------------------------------
# unweighted Umeyama (align_geometry.py:59-82)
# --------------------------------------------------------------------------
def umeyama_sim3(X, Y):
    """align_geometry.py:59-82: Y ~= s R X + t; covariance and variance divided by
    N; reflection decided on det(U)*det(Vt) (:73); eps 1e-12 (:79)."""
    n = X.shape[0]
    mx = X.mean(axis=0)
    my = Y.mean(axis=0)
    Xc = X - mx
    Yc = Y - my
    U, Dg, Vt = np.linalg.svd((Yc.T @ Xc) / n)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1.0
    R = U @ S @ Vt
    var_x = (Xc ** 2).sum() / n
    s = float((Dg * np.diag(S)).sum() / (var_x + 1e-12))
    t = my - s * (R @ mx)
    return s, R, t


# --------------------------------------------------------------------------
# S / A / E — apply, accumulate, re-base
# --------------------------------------------------------------------------
def apply_sim3(points, s, R, t):
    """utils/geometry.py:43-70: s * (p @ R.T) + t; float64 out for float32 in
    because R is float64.  Accepts [N,H,W,3] or [M,3]."""
    shp = points.shape
    out = np.dot(points.reshape(-1, 3), R.T)
    out = s * out
    out = out + t
    return out.reshape(shp)


def accumulate_sim3(chain):
    """utils/geometry.py:73-119: identity first, then left-to-right composition;
    output length = len(chain) + 1 (empty in -> empty out)."""
    if not chain:
        return []
    acc = [(1.0, np.eye(3), np.zeros(3))]
    acc.append(tuple(chain[0]))
    for i in range(1, len(chain)):
        s_n, R_n, t_n = chain[i]
        s_p, R_p, t_p = acc[i]
        acc.append((s_p * s_n, R_p @ R_n, s_p * (R_p @ t_n) + t_p))
    return acc


def rebase_extrinsic_sim3(extrinsic, s, R, t):
    """utils/geometry.py:121-160 (transform_camara_extrinsics): w2c_ref =
    w2c_cur @ [R^T/s | -R^T t / s]."""
    w2c = np.eye(4)
    w2c[:3, :4] = extrinsic
    T = np.eye(4)
    T[:3, :3] = (1.0 / s) * R.T
    T[:3, 3] = -R.T @ t / s
    return (w2c @ T)[:3, :4]


def chain_extrinsics_from_overlap(E_prev_global_3x4, E_local_Nx3x4, T_4x4):
    """align_geometry.py:333-368: E0g = inv(T) @ E_prev; Ei_g = Ei_l @ inv(E0_l) @ E0g."""
    E0g = np.linalg.inv(T_4x4) @ as_4x4(E_prev_global_3x4)
    E0l_inv = np.linalg.inv(as_4x4(E_local_Nx3x4[0]))
    out = [(as_4x4(E_local_Nx3x4[i]) @ E0l_inv @ E0g)[:3, :4] for i in range(E_local_Nx3x4.shape[0])]
    return np.stack(out, axis=0)


def chain_extrinsics_single_overlap(E_prev_global_3x4, E_local_Nx3x4, R, t):
    """utils/align_geometry_single.py:224-252: same chain, frame-to-frame form
    Ei_g = (Ei_l @ inv(E(i-1)_l)) @ E(i-1)_g."""
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R
    T[:3, 3] = t
    Eg = np.linalg.inv(T) @ as_4x4(E_prev_global_3x4)
    n = E_local_Nx3x4.shape[0]
    out = np.zeros((n, 3, 4), dtype=np.float64)
    out[0] = Eg[:3, :4]
    for i in range(1, n):
        Eg = as_4x4(E_local_Nx3x4[i]) @ np.linalg.inv(as_4x4(E_local_Nx3x4[i - 1])) @ Eg
        out[i] = Eg[:3, :4]
    return out


def image_chunks(items, chunk_size, overlap=1):
    """align_geometry.py:371-392: stride chunk-overlap, forced last start n-chunk."""
    assert chunk_size >= 2
    assert 0 <= overlap < chunk_size
    n = len(items)
    if n < chunk_size:
        return []
    starts = list(range(0, n - chunk_size + 1, chunk_size - overlap))
    if starts[-1] != n - chunk_size:
        starts.append(n - chunk_size)
    return [items[s:s + chunk_size] for s in starts]


def solver_chunk_starts(n_frames, chunk_size, overlap):
    """solver.py:34,80-85,186-196: the deque advances by chunk-overlap and trailing
    frames that never fill a chunk are dropped (300,16,1 -> 19 chunks)."""
    starts, buf_start, buf_len = [], 0, 0
    for _ in range(n_frames):
        buf_len += 1
        if buf_len >= chunk_size:
            starts.append(buf_start)
            if buf_len > overlap:
                adv = min(chunk_size - overlap, buf_len)
                buf_start += adv
                buf_len -= adv
    return starts


# --------------------------------------------------------------------------
# G + I — IRLS on pixel correspondences (utils/align.py:111-218)
# --------------------------------------------------------------------------
def irls_conf_threshold(conf1_flat, conf2_flat):
    """utils/align.py:140-142: min(median, median) * 0.1 — np.float32 for float32
    confidences (NumPy >= 2 weak-scalar promotion)."""
    return min(np.median(conf1_flat), np.median(conf2_flat)) * 0.1


def irls_reference(point_map1, point_map2, conf1, conf2, min_points=100,
                   max_iterations=20, convergence_threshold=1e-6,
                   delta=1.0, indices=None, return_trace=False):
    """utils/align.py:111-218, including the reference behaviours SURVEY.md
    section 0 items 7-8 pin: masks applied independently to the two clouds
    (:145-151) and then indexed with the SAME random indices (:159-165) drawn from
    the global numpy RNG.  `indices` overrides the RNG draw (that is what the GPU
    path receives); `delta` is hard-wired to 1.0 in the reference (:94,189)."""
    p1 = point_map1.reshape(-1, 3)
    p2 = point_map2.reshape(-1, 3)
    c1 = conf1.reshape(-1)
    c2 = conf2.reshape(-1)
    thr = irls_conf_threshold(c1, c2)
    m1 = c1 > thr
    m2 = c2 > thr
    p1f, p2f, c1f, c2f = p1[m1], p2[m2], c1[m1], c2[m2]
    if len(p1f) < min_points or len(p2f) < min_points:
        out = (1.0, np.eye(3), np.zeros(3))
        return (out, {"iters": 0, "thr": thr}) if return_trace else out
    k = min(5000, len(p1f), len(p2f))
    if indices is None:
        indices = np.random.choice(min(len(p1f), len(p2f)), k, replace=False)
    a = p1f[indices]
    b = p2f[indices]
    c = np.sqrt(c1f[indices] * c2f[indices])
    s, R, t = 1.0, np.eye(3), np.zeros(3)
    iters = 0
    for it in range(max_iterations):
        res = np.linalg.norm(a - apply_sim3(b, s, R, t), axis=1)
        # utils/align.py:180-191, vectorised: rho'(r)/r = 1 for r<=delta (incl. 0)
        hub = np.ones_like(res)
        big = res > delta
        hub[big] = delta / res[big]
        w = np.zeros_like(res)
        w[:] = c * hub
        w = w / (np.max(w) + 1e-8)
        s_n, R_n, t_n = weighted_umeyama(b, a, w)
        change = np.abs(s_n - s) + np.linalg.norm(R_n - R) + np.linalg.norm(t_n - t)
        s, R, t = s_n, R_n, t_n
        iters = it + 1
        if change < convergence_threshold:
            break
    out = (s, R, t)
    if return_trace:
        return out, {"iters": iters, "thr": thr, "indices": np.asarray(indices),
                     "n1": int(m1.sum()), "n2": int(m2.sum())}
    return out


# --------------------------------------------------------------------------
# N — norm-ratio Umeyama on pixel correspondences (utils/align.py:224-276)
# --------------------------------------------------------------------------
def umeyama_norm_ratio(point_map2, point_map1):
    """utils/align.py:224-276.  NOTE the swapped parameter names at :224: the first
    positional argument is called point_map2.  Frame 0 only (:238-239); scale =
    sum||y-mu_y|| / sum||x-mu_x|| (:250-252); Kabsch on s*Xc^T Yc (:256-271)."""
    a = point_map1[0].reshape(-1, 3)
    b = point_map2[0].reshape(-1, 3)
    ca = np.mean(a, axis=0)
    cb = np.mean(b, axis=0)
    a0 = a - ca
    b0 = b - cb
    da = np.linalg.norm(a0, axis=1).sum()
    db = np.linalg.norm(b0, axis=1).sum()
    s = db / da if da > 0 else 1.0
    Hm = (a0 * s).T @ b0
    U, _, Vt = np.linalg.svd(Hm)
    Sg = np.eye(3)
    if np.linalg.det(Vt.T @ U.T) < 0:
        Sg[2, 2] = -1
    R = Vt.T @ Sg @ U.T
    t = cb - s * (R @ ca.T).T
    return float(s), R, t


# --------------------------------------------------------------------------
# K — KD-tree Umeyama-ICP (align_geometry.py:84-140).  PARITY UNPINNED: the
# reference uses Open3D's KDTreeFlann (un-vendored, absent); scipy's cKDTree gives
# the same exact nearest neighbour except under exact distance ties.
# --------------------------------------------------------------------------
def umeyama_icp_kdtree(source, target, threshold=0.001, max_iterations=30):
    from scipy.spatial import cKDTree
    source = source[np.isfinite(source).all(axis=1)]
    target = target[np.isfinite(target).all(axis=1)]
    tree = cKDTree(target.astype(np.float64))
    s_c, R_c, t_c = 1.0, np.eye(3), np.zeros(3)
    src = source.astype(np.float64)
    for _ in range(max_iterations):
        warped = (s_c * (R_c @ src.T)).T + t_c
        dist, idx = tree.query(warped, k=1)
        inl = (dist * dist) < (threshold * threshold)
        if inl.sum() < 20:
            break
        s_u, R_u, t_u = umeyama_sim3(warped[inl], target[idx[inl]])
        s_c, R_c, t_c = s_u * s_c, R_u @ R_c, s_u * (R_u @ t_c) + t_u
    return float(s_c), R_c, t_c


def icp_point_to_point_kdtree(source, target, threshold=0.0001, max_iterations=50):
    """C — Open3D registration_icp(TransformationEstimationPointToPoint(), ICPConvergenceCriteria(max_iteration))
    as called at align_geometry.py:29-45 / utils/align_geometry_single.py:146-160.  Open3D is not vendored and not
    installed: PARITY UNPINNED.  Restated from its published algorithm (pipelines/registration/Registration.cpp):
    evaluate correspondences (nearest neighbour within `threshold`), then up to max_iteration times
    {Kabsch update without scale on the correspondences, apply, re-evaluate, stop when |d fitness| < 1e-6 and
    |d inlier_rmse| < 1e-6}.  Returns (1.0, R, t)."""
    from scipy.spatial import cKDTree
    src = np.asarray(source, np.float64)
    tgt = np.asarray(target, np.float64)
    src = src[np.isfinite(src).all(axis=1)]
    tgt = tgt[np.isfinite(tgt).all(axis=1)]
    tree = cKDTree(tgt)
    R, t = np.eye(3), np.zeros(3)

    def evaluate(R, t):
        w = src @ R.T + t
        dist, idx = tree.query(w, k=1)
        inl = (dist * dist) < threshold * threshold
        n = int(inl.sum())
        fit = n / max(len(src), 1)
        rmse = float(np.sqrt((dist[inl] ** 2).sum() / n)) if n else 0.0
        return w, idx, inl, fit, rmse

    w, idx, inl, fit, rmse = evaluate(R, t)
    for _ in range(max_iterations):
        if inl.sum() < 1:
            break
        X, Y = w[inl], tgt[idx[inl]]
        mx, my = X.mean(0), Y.mean(0)
        cov = (Y - my).T @ (X - mx) / len(X)
        U, D, Vt = np.linalg.svd(cov)
        S = np.eye(3)
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            S[2, 2] = -1.0
        Ru = U @ S @ Vt
        tu = my - Ru @ mx
        R, t = Ru @ R, Ru @ t + tu
        w, idx, inl, fit2, rmse2 = evaluate(R, t)
        stop = abs(fit - fit2) < 1e-6 and abs(rmse - rmse2) < 1e-6
        fit, rmse = fit2, rmse2
        if stop:
            break
    return 1.0, R, t


# --------------------------------------------------------------------------
# P / V — viewer-side filtering (viewer.py:198-234, :317-356;
#          utils/viser_server.py:107-108, :182-186)
# --------------------------------------------------------------------------
def viewer_frame_points(depth_hw, conf_hw, extrinsic_3x4, intrinsic_3x3, vis_stride=1):
    """viewer.py:198-218: U3 world points, stride mask, then keep
    0.1 < z_world < 50 and all-finite.  Returns (points [n,3] f64, conf [n],
    flat pixel index [n])."""
    d = depth_hw[..., 0] if depth_hw.ndim == 3 else depth_hw
    world = unproject_world_vggt(d[None], extrinsic_3x4[None], intrinsic_3x3[None])[0][0]
    H, W = d.shape
    stride = np.zeros((H, W), dtype=bool)
    stride[::vis_stride, ::vis_stride] = True
    pts = world[stride]
    cf = conf_hw[stride] if conf_hw.shape == (H, W) else np.ones(len(pts))
    ok = (pts[:, 2] > 0.1) & (pts[:, 2] < 50.0) & np.all(np.isfinite(pts), axis=1)
    pix = np.flatnonzero(stride.reshape(-1))[ok]
    return pts[ok], cf[ok], pix


def viewer_conf_mask(all_conf, percent):
    """viewer.py:333-338: threshold = percentile(conf[conf>0], min(percent, 99.9)),
    keep conf >= threshold; all-true when no confidence is positive."""
    if len(all_conf) > 0 and np.any(all_conf > 0):
        thr = np.percentile(all_conf[all_conf > 0], min(percent, 99.9))
        return all_conf >= thr, thr
    return np.ones(len(all_conf), dtype=bool), None


def viser_conf_mask(conf_flat, percent, floor):
    """utils/viser_server.py:107-108 (floor 0.1) and :182-186 (floor 1e-5):
    percentile over ALL confidences, keep >= threshold and > floor."""
    thr = np.percentile(conf_flat, percent)
    return (conf_flat >= thr) & (conf_flat > floor), thr


# --------------------------------------------------------------------------
# exact restatement of np.median / np.percentile for float32 data from two order
# statistics — what the GPU path reproduces after exact selection (SURVEY.md
# appendix A rows "D", "IRLS threshold", "Percentile").
# --------------------------------------------------------------------------
def median_from_order_stats(sorted_lo, sorted_hi):
    """np.median of float32 data: odd n -> the middle element (lo==hi); even n ->
    np.mean of the two middles = float32 add then divide by 2."""
    lo = np.float32(sorted_lo)
    hi = np.float32(sorted_hi)
    return np.float32(np.float32(lo + hi) / np.float32(2.0))


def percentile_indices_f32(n, percent):
    """numpy/lib/_function_base_impl.py (2.3): q = percent / float32(100) is float32;
    virtual index (n-1)*q is float32; previous=floor, next=previous+1, both clamped
    to the last element when virtual >= n-1; gamma = virtual - previous (float32)."""
    q = np.true_divide(percent, np.float32(100))
    virt = (n - 1) * q
    prev = np.floor(virt)
    nxt = prev + 1
    if virt >= n - 1:
        prev_i = nxt_i = n - 1
    elif virt < 0:
        prev_i = nxt_i = 0
    else:
        prev_i, nxt_i = int(prev), int(nxt)
    gamma = np.float32(virt - prev)
    return prev_i, nxt_i, gamma


def percentile_from_order_stats(a_prev, a_next, gamma):
    """numpy _lerp in float32: a + (b-a)*g, replaced by b - (b-a)*(1-g) where g>=0.5."""
    a = np.float32(a_prev)
    b = np.float32(a_next)
    g = np.float32(gamma)
    diff = np.float32(b - a)
    out = np.float32(a + np.float32(diff * g))
    if g >= 0.5:
        out = np.float32(b - np.float32(diff * np.float32(np.float32(1) - g)))
    return out
