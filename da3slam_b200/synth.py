"""Synthetic submaps with the shapes BASELINE.json names.

A *submap* is what Depth Anything 3 returns for one chunk of frames
(solver.py:168-176): ``depth [F,H,W] f32``, ``conf [F,H,W] f32`` (>= 1),
``extrinsics [F,3,4] f32`` world-to-camera in the submap's own world frame,
``intrinsics [F,3,3] f32`` and ``processed_images [F,H,W,3] u8``.

Consecutive submaps share ``overlap`` frames.  Submap k+1's copy of a shared frame
is submap k's copy re-expressed under a ground-truth Sim(3)  p_k = s R p_{k+1} + t:
    c2w_k = (R R_{k+1}, s R t_{k+1} + t),   depth_k = s * depth_{k+1}
so the pixel correspondences of the overlap frames recover (s, R, t) exactly up to
the injected depth noise / outliers.

numpy only (host, seeded, reproducible); ``to_device`` moves a submap to CUDA.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def rotvec_to_matrix(rv):
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv)
    if th < 1e-15:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_intrinsics(n, H, W):
    K = np.zeros((n, 3, 3), F32)
    K[:, 0, 0] = K[:, 1, 1] = 0.9 * W
    K[:, 0, 2] = (W - 1) / 2
    K[:, 1, 2] = (H - 1) / 2
    K[:, 2, 2] = 1
    return K


def smooth_depth(rng, n, H, W):
    """d0 + a sin(2 pi u f1 / W) cos(2 pi v f2 / H) in [0.5, 3.0] + N(0, 0.002)."""
    u = np.arange(W, dtype=np.float64)[None, None, :] / W
    v = np.arange(H, dtype=np.float64)[None, :, None] / H
    d0 = rng.uniform(1.2, 2.2, size=(n, 1, 1))
    a = rng.uniform(0.2, 0.7, size=(n, 1, 1))
    f1 = rng.uniform(0.5, 2.5, size=(n, 1, 1))
    f2 = rng.uniform(0.5, 2.5, size=(n, 1, 1))
    ph = rng.uniform(0, 2 * np.pi, size=(n, 1, 1))
    d = d0 + a * np.sin(2 * np.pi * u * f1 + ph) * np.cos(2 * np.pi * v * f2)
    d = np.clip(d, 0.5, 3.0) + rng.normal(0, 0.002, size=(n, H, W))
    return d.astype(F32)


def da3_like_conf(rng, n, H, W, border=6, offset=0.0):
    """offset + exp(N(0, 0.75)), 5 % of pixels forced to `offset`, low-confidence
    border band.  DA3 confidences are >= 1 (offset=1); the alignment stage upstream
    sees conf - 1 (utils/da3_streaming.py:276), which is the default here so that
    the 0.1 x median threshold actually masks pixels."""
    c = offset + np.exp(rng.normal(0, 0.75, size=(n, H, W)))
    c[rng.random((n, H, W)) < 0.05] = offset
    b = min(border, H // 4, W // 4)
    if b > 0:
        band = np.full((H, W), False)
        band[:b] = band[-b:] = True
        band[:, :b] = band[:, -b:] = True
        c[:, band] = offset + 0.01 * rng.random((n, int(band.sum())))
    return c.astype(F32)


def random_sim3(rng, s_range=(0.7, 1.4), max_rot=0.3, max_t=1.0):
    s = float(rng.uniform(*s_range))
    rv = rng.normal(size=3)
    rv = rv / np.linalg.norm(rv) * rng.uniform(0, max_rot)
    t = rng.uniform(-max_t, max_t, size=3)
    return s, rotvec_to_matrix(rv), t


def trajectory_w2c(rng, n, rot_sigma=0.02, step=0.05):
    """Smooth camera walk; frame 0 is close to identity.  Returns [n,3,4] float64 w2c."""
    E = np.zeros((n, 3, 4))
    R = rotvec_to_matrix(rng.normal(0, 0.01, size=3))
    c = rng.normal(0, 0.01, size=3)           # camera centre in world
    for i in range(n):
        E[i, :, :3] = R.T                     # w2c rotation = (c2w rotation)^T
        E[i, :, 3] = -R.T @ c
        R = R @ rotvec_to_matrix(rng.normal(0, rot_sigma, size=3))
        c = c + R @ np.array([0, 0, step]) + rng.normal(0, 0.2 * step, size=3)
    return E


def c2w_of(E):
    R = np.transpose(E[..., :3, :3], (*range(E.ndim - 2), E.ndim - 1, E.ndim - 2))
    t = -np.einsum("...ij,...j->...i", R, E[..., :3, 3])
    return R, t


def w2c_from_c2w(R, t):
    E = np.zeros(R.shape[:-2] + (3, 4))
    Rt = np.swapaxes(R, -1, -2)
    E[..., :3, :3] = Rt
    E[..., :3, 3] = -np.einsum("...ij,...j->...i", Rt, t)
    return E


def make_sequence(n_submaps, frames, H, W, overlap=1, seed=1234, outlier_ratio=0.0,
                  depth_noise=0.002, with_images=False):
    """Returns (submaps, gt) where submaps is a list of dicts (the Prediction
    fields) and gt[k] = (s, R, t) maps submap k+1 coordinates into submap k's."""
    rng = np.random.default_rng(seed)
    subs, gt = [], []
    K = make_intrinsics(frames, H, W)
    prev = None
    for k in range(n_submaps):
        E = trajectory_w2c(rng, frames)
        depth = smooth_depth(rng, frames, H, W)
        conf = da3_like_conf(rng, frames, H, W)
        if prev is not None:
            s, R, t = random_sim3(rng)
            gt.append((s, R, t))
            # overlap frames: cur[:o] observes the same scene as prev[-o:]
            Rp, tp = c2w_of(prev["extrinsics"][-overlap:].astype(np.float64))   # c2w_prev
            # c2w_prev = (R R_cur, s R t_cur + t)  =>  R_cur = R^T Rp, t_cur = R^T (tp - t) / s
            Rc = np.einsum("ij,njk->nik", R.T, Rp)
            tc = np.einsum("ij,nj->ni", R.T, tp - t) / s
            E[:overlap] = w2c_from_c2w(Rc, tc)
            d = prev["depth"][-overlap:].astype(np.float64) / s
            d = d + rng.normal(0, depth_noise, size=d.shape)
            if outlier_ratio > 0:
                bad = rng.random(d.shape) < outlier_ratio
                d[bad] = rng.uniform(0.3, 4.0, size=int(bad.sum()))
            depth[:overlap] = d.astype(F32)
        sub = {"depth": depth, "conf": conf, "extrinsics": E.astype(F32), "intrinsics": K.copy()}
        if with_images:
            sub["processed_images"] = rng.integers(0, 256, size=(frames, H, W, 3), dtype=np.uint8)
        subs.append(sub)
        prev = sub
    return subs, gt


def make_pair(H, W, frames=2, overlap=1, seed=1234, **kw):
    subs, gt = make_sequence(2, frames, H, W, overlap, seed, **kw)
    return subs[0], subs[1], gt[0]


def to_device(sub, device="cuda"):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in sub.items()}


# ------------------------------------------------------------------------------------------
# device-side generator (same distributions, torch RNG on the GPU): used by bench.py where
# 10^8 pixels would take minutes with numpy.  Extrinsics / ground truth are drawn on the host.
# ------------------------------------------------------------------------------------------
def make_sequence_device(n_submaps, frames, H, W, overlap=1, seed=1234, outlier_ratio=0.0, depth_noise=0.002,
                         with_images=True, device="cuda", conf_offset=0.0, abs_scale=None):
    """Returns (submaps, gt): submaps are dicts of CUDA tensors with the Prediction fields.

    abs_scale=(lo, hi): every submap draws its OWN metric scale sigma_k in [lo, hi] (sigma_0 = 1) and the pair scale is
    sigma_k / sigma_{k-1} — what a network that predicts each chunk at an arbitrary scale produces.  Without it the pair
    scales are independent draws, so the accumulated scale is a random walk over the sequence (after 64 pairs anything
    from 0.2 to 5: map extent and voxel count then vary 20x between seeds)."""
    import torch
    rng = np.random.default_rng(seed)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    dev = torch.device(device)
    u = (torch.arange(W, device=dev, dtype=torch.float32) / W).view(1, 1, W)
    v = (torch.arange(H, device=dev, dtype=torch.float32) / H).view(1, H, 1)
    K = torch.from_numpy(make_intrinsics(frames, H, W)).to(dev)
    b = min(6, H // 4, W // 4)
    band = torch.zeros((H, W), dtype=torch.bool, device=dev)
    if b > 0:
        band[:b] = True
        band[-b:] = True
        band[:, :b] = True
        band[:, -b:] = True
    subs, gt, prev = [], [], None
    sigma_prev = 1.0
    for k in range(n_submaps):
        E = trajectory_w2c(rng, frames)
        p = {name: torch.from_numpy(rng.uniform(lo, hi, size=(frames, 1, 1)).astype(np.float32)).to(dev)
             for name, (lo, hi) in dict(d0=(1.2, 2.2), a=(0.2, 0.7), f1=(0.5, 2.5), f2=(0.5, 2.5), ph=(0, 2 * np.pi)).items()}
        depth = p["d0"] + p["a"] * torch.sin(2 * np.pi * u * p["f1"] + p["ph"]) * torch.cos(2 * np.pi * v * p["f2"])
        depth = depth.clamp_(0.5, 3.0) + 0.002 * torch.randn((frames, H, W), device=dev, generator=g)
        conf = conf_offset + torch.exp(0.75 * torch.randn((frames, H, W), device=dev, generator=g))
        conf[torch.rand((frames, H, W), device=dev, generator=g) < 0.05] = conf_offset
        conf[:, band] = conf_offset + 0.01 * torch.rand((frames, int(band.sum())), device=dev, generator=g)
        if prev is not None:
            s, R, t = random_sim3(rng)
            if abs_scale is not None:
                sigma = float(rng.uniform(*abs_scale))
                s, sigma_prev = sigma / sigma_prev, sigma
            gt.append((s, R, t))
            Ep = prev["extrinsics"][-overlap:].double().cpu().numpy()
            Rp, tp = c2w_of(Ep)
            Rc = np.einsum("ij,njk->nik", R.T, Rp)
            tc = np.einsum("ij,nj->ni", R.T, tp - t) / s
            E[:overlap] = w2c_from_c2w(Rc, tc)
            d = prev["depth"][-overlap:] / float(s) + depth_noise * torch.randn((overlap, H, W), device=dev, generator=g)
            if outlier_ratio > 0:
                bad = torch.rand((overlap, H, W), device=dev, generator=g) < outlier_ratio
                d = torch.where(bad, 0.3 + 3.7 * torch.rand((overlap, H, W), device=dev, generator=g), d)
            depth[:overlap] = d
        sub = {"depth": depth.contiguous(), "conf": conf.contiguous(),
               "extrinsics": torch.from_numpy(E.astype(F32)).to(dev), "intrinsics": K.clone()}
        if with_images:
            sub["processed_images"] = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
        subs.append(sub)
        prev = sub
    return subs, gt


def submap_to_host(sub):
    return {k: v.cpu().numpy() for k, v in sub.items()}
