"""Write tests/golden/*.npz from the REAL reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The GPU box has no /root/reference; the tests replay these files.  Every array in
a fixture is either an input we generated (seeded) or an output of an unmodified
reference function called through oracle/ref_loader.py.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def rot(rv):
    th = np.linalg.norm(rv)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def scene(rng, n, H, W):
    depth = (1.0 + rng.random((n, H, W)) * 2.0).astype(np.float32)
    conf = (1.0 + np.exp(rng.normal(0, 0.75, (n, H, W)))).astype(np.float32)
    K = np.zeros((n, 3, 3), np.float32)
    K[:, 0, 0] = 0.9 * W
    K[:, 1, 1] = 0.8 * W
    K[:, 0, 2] = (W - 1) / 2 + 0.3
    K[:, 1, 2] = (H - 1) / 2 - 0.2
    K[:, 2, 2] = 1
    E = np.zeros((n, 3, 4), np.float32)
    for i in range(n):
        E[i, :, :3] = rot(rng.normal(0, 0.2, 3))
        E[i, :, 3] = rng.normal(0, 0.5, 3)
    return depth, conf, K, E


def main():
    ref = ref_loader.load()
    ag, ags, geo, al, vg = ref.ag, ref.ags, ref.geo, ref.al, ref.vg
    out = {}
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- unprojection (U1, U2, U3) ------------------------------------------
    rng = np.random.default_rng(101)
    depth, conf, K, E = scene(rng, 3, 20, 28)
    g = dict(depth=depth, conf=conf, K=K, E=E)
    g["u1_camera"] = ag.depth_to_point_cloud_vectorized(depth, K, E, in_coords="camera")
    g["u1_world"] = ag.depth_to_point_cloud_vectorized(depth, K, E, in_coords="world")
    g["u1s_camera"] = ags.depth_to_point_cloud_vectorized(depth, K, E, in_coords="camera")
    g["u1s_world"] = ags.depth_to_point_cloud_vectorized(depth, K, E, in_coords="world")
    g["u2_world"] = geo.depth_to_point_cloud_vectorized(depth, K, E)
    g["u3_world"] = vg.unproject_depth_map_to_point_map(depth[..., None], E, K)
    w, c, m = vg.depth_to_world_coords_points(depth[1], E[1], K[1])
    g["u3_world_1"], g["u3_cam_1"], g["u3_mask_1"] = w, c, m
    g["se3_inv"] = vg.closed_form_inverse_se3(E)
    np.savez_compressed(os.path.join(HERE, "unproject.npz"), **g)

    # ---- depth scale (D) ----------------------------------------------------
    rng = np.random.default_rng(202)
    g = {}
    for case in range(4):
        dA, cA, _, _ = scene(rng, 2, 24, 30)
        dB, cB, _, _ = scene(rng, 2, 24, 30)
        cA = cA - 1.0
        cB = cB - 1.0
        if case == 1:                      # holes / non-finite values
            dA[-1, :5] = 0.0
            dB[0, 5:8] = np.nan
            dA[-1, 8:10] = np.inf
        if case == 2:                      # fewer than 50 valid pixels -> guarded path gives 1.0
            cA[-1] = 0.0
            cA[-1, 0, :30] = 5.0
        if case == 3:                      # odd count
            cA[-1, 0, 0] = 0.0
        prev_d = dict(depth=dA, conf=cA)
        cur_d = dict(depth=dB, conf=cB)
        prev_o = types.SimpleNamespace(depth=dA, conf=cA)
        cur_o = types.SimpleNamespace(depth=dB, conf=cB)
        g[f"dA{case}"], g[f"cA{case}"], g[f"dB{case}"], g[f"cB{case}"] = dA, cA, dB, cB
        with np.errstate(all="ignore"):
            g[f"plain{case}"] = np.float64(ag.estimate_depth_scale(prev_o, cur_o, conf_th=0.2))
            g[f"guard{case}"] = np.float64(ags.estimate_depth_scale(prev_d, cur_d, conf_th=0.2))
            g[f"guard_obj{case}"] = np.float64(ags.estimate_depth_scale(prev_o, cur_o))
    # no-conf container
    g["noconf"] = np.float64(ags.estimate_depth_scale(dict(depth=dA), dict(depth=dB)))
    np.savez_compressed(os.path.join(HERE, "depth_scale.npz"), **g)

    # ---- Umeyama family (W, W0, _umeyama_sim3, huber, N) ---------------------
    rng = np.random.default_rng(303)
    g = {}
    src = rng.normal(0, 1, (400, 3))
    s0, R0, t0 = 1.3, rot(np.array([0.2, -0.1, 0.25])), np.array([0.3, -0.2, 0.9])
    dst = s0 * src @ R0.T + t0 + rng.normal(0, 0.01, (400, 3))
    wts = rng.random(400).astype(np.float32)
    g.update(src=src, dst=dst, w=wts)
    s, R, t = al.weighted_umeyama_alignment(src, dst, wts)
    g["W_s"], g["W_R"], g["W_t"] = np.float64(s), R, t
    s, R, t = al.weighted_umeyama_alignment(src.astype(np.float32), dst.astype(np.float32), wts)
    g["W32_s"], g["W32_R"], g["W32_t"] = np.float64(s), R, t
    s, R, t = al.weighted_umeyama_alignment0(src, dst, wts.astype(np.float64))
    g["W0_s"], g["W0_R"], g["W0_t"] = np.float64(s), R, t
    s, R, t = ag._umeyama_sim3(src, dst)
    g["U_s"], g["U_R"], g["U_t"] = np.float64(s), R, t
    # reflection case: mirror one axis so det(U Vt) < 0 without the fix
    dst_m = dst.copy()
    dst_m[:, 2] *= -1
    s, R, t = al.weighted_umeyama_alignment(src, dst_m, wts)
    g["Wm_s"], g["Wm_R"], g["Wm_t"] = np.float64(s), R, t
    s, R, t = ag._umeyama_sim3(src, dst_m)
    g["Um_s"], g["Um_R"], g["Um_t"] = np.float64(s), R, t
    g["dst_m"] = dst_m
    # 3-point (RANSAC minimal sample) case
    s, R, t = ag._umeyama_sim3(src[:3], dst[:3])
    g["U3_s"], g["U3_R"], g["U3_t"] = np.float64(s), R, t
    rs = np.array([0.0, 0.3, 1.0, 1.0000001, 2.5, -4.0])
    g["huber_r"] = rs
    g["huber_w"] = np.array([al.huber_weight(float(r)) for r in rs])
    g["huber_w_d05"] = np.array([al.huber_weight(float(r), 0.5) for r in rs])
    pm1 = dst.reshape(1, 20, 20, 3).astype(np.float32)
    pm2 = src.reshape(1, 20, 20, 3).astype(np.float32)
    s, R, t = al.align_two_point_clouds_umeyama(pm1, pm2)          # (point_map2=pm1, point_map1=pm2)
    g["N_s"], g["N_R"], g["N_t"] = np.float64(s), R, t
    s, R, t = al.align_two_point_clouds(pm1, pm2)
    g["Napi_s"], g["Napi_R"], g["Napi_t"] = np.float64(s), R, t
    g["pm1"], g["pm2"] = pm1, pm2
    np.savez_compressed(os.path.join(HERE, "umeyama.npz"), **g)

    # ---- IRLS (G + I) --------------------------------------------------------
    rng = np.random.default_rng(404)
    g = {}
    H, W = 60, 100                                   # 6000 px > 5000 so the subsample is real
    pts2 = rng.normal(0, 1, (1, H, W, 3)).astype(np.float32)
    s0, R0, t0 = 0.8, rot(np.array([-0.1, 0.3, 0.05])), np.array([-0.4, 0.1, 0.2])
    pts1 = (s0 * pts2.reshape(-1, 3) @ R0.T + t0).reshape(1, H, W, 3)
    pts1 = (pts1 + rng.normal(0, 0.02, pts1.shape)).astype(np.float32)
    far = rng.random((1, H, W)) < 0.1                 # gross outliers -> Huber branch is exercised
    pts1[far] += rng.normal(0, 3.0, (int(far.sum()), 3)).astype(np.float32)
    c1 = np.exp(rng.normal(0, 0.75, (1, H, W))).astype(np.float32)
    c2 = np.exp(rng.normal(0, 0.75, (1, H, W))).astype(np.float32)
    g.update(pm1=pts1, pm2=pts2, c1=c1, c2=c2)
    for tag, (a1, a2) in {"same": (c1, c1), "indep": (c1, c2)}.items():
        for seed in (0, 1):
            np.random.seed(seed)
            with quiet:
                s, R, t = al.align_two_point_clouds_irls(pts1, pts2, a1, a2)
            # replay the reference's RNG draw (utils/align.py:159-160)
            thr = min(np.median(a1.reshape(-1)), np.median(a2.reshape(-1))) * 0.1
            n1 = int((a1.reshape(-1) > thr).sum())
            n2 = int((a2.reshape(-1) > thr).sum())
            np.random.seed(seed)
            idx = np.random.choice(min(n1, n2), min(5000, n1, n2), replace=False)
            g[f"{tag}{seed}_s"], g[f"{tag}{seed}_R"], g[f"{tag}{seed}_t"] = np.float64(s), R, t
            g[f"{tag}{seed}_idx"] = idx
            g[f"{tag}{seed}_thr"] = np.asarray(thr)
    # too-few-points early return (utils/align.py:154-156)
    with quiet:
        s, R, t = al.align_two_point_clouds_irls(pts1[:, :5, :10], pts2[:, :5, :10], c1[:, :5, :10], c1[:, :5, :10])
    g["few_s"], g["few_R"], g["few_t"] = np.float64(s), R, t
    np.savez_compressed(os.path.join(HERE, "irls.npz"), **g)

    # ---- apply / accumulate / extrinsics chaining (S, A, E) ----------------------
    rng = np.random.default_rng(505)
    g = {}
    P4 = rng.normal(0, 1, (2, 6, 7, 3)).astype(np.float32)
    P2 = rng.normal(0, 1, (50, 3))
    g.update(P4=P4, P2=P2, s=np.float64(1.7), R=R0, t=t0)
    g["S4"] = geo.apply_sim3_transform(P4, 1.7, R0, t0)
    g["S2"] = geo.apply_sim3_transform(P2, 1.7, R0, t0)
    chain = []
    for i in range(5):
        chain.append((float(rng.uniform(0.7, 1.4)), rot(rng.normal(0, 0.3, 3)), rng.normal(0, 1, 3)))
    acc = geo.accumulate_sim3_transforms(chain)
    g["chain_s"] = np.array([c[0] for c in chain])
    g["chain_R"] = np.stack([c[1] for c in chain])
    g["chain_t"] = np.stack([c[2] for c in chain])
    g["acc_s"] = np.array([c[0] for c in acc])
    g["acc_R"] = np.stack([c[1] for c in acc])
    g["acc_t"] = np.stack([c[2] for c in acc])
    g["acc_empty_len"] = np.int64(len(geo.accumulate_sim3_transforms([])))
    _, _, K, E = scene(rng, 5, 4, 4)
    g["E_local"] = E
    g["rebase"] = geo.transform_camara_extrinsics(E[2], 1.7, R0, t0)
    T = np.eye(4)
    T[:3, :3] = R0
    T[:3, 3] = t0
    E_prev = E[4].astype(np.float64)
    g["chain_overlap"] = ag.compute_aligned_chunk_extrinsics_from_prev_overlap(E_prev, E.astype(np.float64), T)
    g["T"] = T
    np.savez_compressed(os.path.join(HERE, "sim3_chain.npz"), **g)

    # ---- chunking ---------------------------------------------------------------
    g = {}
    for n, c, o in [(300, 16, 1), (2000, 32, 1), (2000, 64, 1), (10, 4, 1), (3, 4, 1), (9, 4, 2), (7, 7, 0)]:
        ch = ag.make_image_chunks(list(range(n)), c, o)
        g[f"chunks_{n}_{c}_{o}"] = np.array([x[0] for x in ch], dtype=np.int64)
        g[f"chunklen_{n}_{c}_{o}"] = np.array([len(x) for x in ch], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "chunks.npz"), **g)

    # ---- overlap extraction + single-overlap chunk alignment plumbing (X) --------
    rng = np.random.default_rng(606)
    g = {}
    dA, cA, KA, EA = scene(rng, 4, 12, 16)
    dB, cB, KB, EB = scene(rng, 4, 12, 16)
    prev = dict(depth=dA, conf=cA, intrinsics=KA, extrinsics=EA)
    cur = dict(depth=dB, conf=cB, intrinsics=KB, extrinsics=EB)
    g.update({f"prev_{k}": v for k, v in prev.items()})
    g.update({f"cur_{k}": v for k, v in cur.items()})
    with quiet:
        a, b = ag.extract_overlap_point_cloud(types.SimpleNamespace(**prev), types.SimpleNamespace(**cur))
    g["X_root_prev"], g["X_root_cur"] = a, b
    a, b = ags.extract_single_overlap_point_cloud(prev, cur)
    g["X_single_prev"], g["X_single_cur"] = a, b
    a, b, ca, cb = al.extract_overlap_chunk_prediction(prev, cur, 2)
    g["X_align_pm1"], g["X_align_pm2"] = a, b
    assert ca is None and cb is None
    np.savez_compressed(os.path.join(HERE, "overlap.npz"), **g)

    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"golden fixtures written to {HERE} ({total / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
