"""Pin the oracle (oracle/ref_port.py) to the reference.

tests/golden/*.npz hold outputs of the UNMODIFIED reference functions (written by
tests/golden/make_golden.py through oracle/ref_loader.py).  Bit-exact where the
port performs the same numpy/torch calls; 1e-12 where LAPACK/BLAS call order may
differ.  CPU only.
"""
import types

import numpy as np
import pytest

from oracle import ref_loader
from oracle import ref_port as rp


def same(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape
    assert np.array_equal(a, b, equal_nan=True)


def close(a, b, tol=1e-12):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


def test_unproject_u1_u2_u3(golden):
    g = golden("unproject")
    d, K, E = g["depth"], g["K"], g["E"]
    same(rp.unproject_f32(d, K, E, "camera"), g["u1_camera"])
    same(rp.unproject_f32(d, K, E, "world"), g["u1_world"])
    same(g["u1s_camera"], g["u1_camera"])            # the two reference twins agree
    same(g["u1s_world"], g["u1_world"])
    same(rp.unproject_world_f64(d, K, E), g["u2_world"])
    w, c, m = rp.unproject_world_vggt(d, E, K)
    same(w, g["u3_world"])
    same(w[1], g["u3_world_1"])
    same(c[1], g["u3_cam_1"])
    same(m[1], g["u3_mask_1"])
    same(rp.se3_inverse_closed_form(E), g["se3_inv"])
    # the three reference unprojections agree with each other to ~5e-7
    assert np.abs(g["u1_world"] - g["u2_world"]).max() < 5e-6
    assert np.abs(g["u3_world"] - g["u2_world"]).max() < 5e-6


def test_depth_scale(golden):
    g = golden("depth_scale")
    for case in range(4):
        prev = dict(depth=g[f"dA{case}"], conf=g[f"cA{case}"])
        cur = dict(depth=g[f"dB{case}"], conf=g[f"cB{case}"])
        po, co = types.SimpleNamespace(**prev), types.SimpleNamespace(**cur)
        with np.errstate(all="ignore"):
            same(np.float64(rp.depth_scale_plain(po, co, conf_th=0.2)), g[f"plain{case}"])
            same(np.float64(rp.depth_scale_guarded(prev, cur, conf_th=0.2)), g[f"guard{case}"])
            same(np.float64(rp.depth_scale_guarded(po, co)), g[f"guard_obj{case}"])
    assert g["guard2"] == 1.0                         # < 50 valid pixels
    same(np.float64(rp.depth_scale_guarded(dict(depth=g["dA3"]), dict(depth=g["dB3"]))), g["noconf"])


def test_umeyama_family(golden):
    g = golden("umeyama")
    src, dst, w = g["src"], g["dst"], g["w"]
    for tag, (a, b) in {"W": (src, dst), "W32": (src.astype(np.float32), dst.astype(np.float32)),
                        "Wm": (src, g["dst_m"])}.items():
        s, R, t = rp.weighted_umeyama(a, b, w)
        same(np.float64(s), g[tag + "_s"]); same(R, g[tag + "_R"]); same(t, g[tag + "_t"])
        assert abs(np.linalg.det(R) - 1) < 1e-12
    s, R, t = rp.weighted_umeyama_legacy(src, dst, w.astype(np.float64))
    same(np.float64(s), g["W0_s"]); same(R, g["W0_R"]); same(t, g["W0_t"])
    for tag, (a, b) in {"U": (src, dst), "Um": (src, g["dst_m"]), "U3": (src[:3], dst[:3])}.items():
        s, R, t = rp.umeyama_sim3(a, b)
        same(np.float64(s), g[tag + "_s"]); same(R, g[tag + "_R"]); same(t, g[tag + "_t"])
    same(np.array([rp.huber_weight(float(r)) for r in g["huber_r"]]), g["huber_w"])
    same(np.array([rp.huber_weight(float(r), 0.5) for r in g["huber_r"]]), g["huber_w_d05"])
    s, R, t = rp.umeyama_norm_ratio(g["pm1"], g["pm2"])
    same(np.float64(s), g["N_s"]); same(R, g["N_R"]); same(t, g["N_t"])
    same(g["Napi_R"], g["N_R"])                       # utils/align.py:301 returns variant N


def test_irls_reference_quirks(golden):
    g = golden("irls")
    pm1, pm2 = g["pm1"], g["pm2"]
    for tag, (a1, a2) in {"same": (g["c1"], g["c1"]), "indep": (g["c1"], g["c2"])}.items():
        for seed in (0, 1):
            # explicit indices (what the GPU path receives) ...
            (s, R, t), tr = rp.irls_reference(pm1, pm2, a1, a2, indices=g[f"{tag}{seed}_idx"], return_trace=True)
            same(np.float64(s), g[f"{tag}{seed}_s"]); same(R, g[f"{tag}{seed}_R"]); same(t, g[f"{tag}{seed}_t"])
            same(np.asarray(tr["thr"]), g[f"{tag}{seed}_thr"])
            assert np.asarray(tr["thr"]).dtype == np.float32      # NumPy >= 2 promotion (SURVEY appx A)
            # ... and the global-RNG draw itself (utils/align.py:159-160)
            np.random.seed(seed)
            s2, R2, t2 = rp.irls_reference(pm1, pm2, a1, a2)
            same(R2, g[f"{tag}{seed}_R"])
    s, R, t = rp.irls_reference(pm1[:, :5, :10], pm2[:, :5, :10], g["c1"][:, :5, :10], g["c1"][:, :5, :10])
    same(np.float64(s), g["few_s"]); same(R, g["few_R"]); same(t, g["few_t"])
    # the independent-mask quirk really changes the answer
    assert np.abs(g["indep0_R"] - g["same0_R"]).max() > 1e-3


def test_apply_accumulate_chain(golden):
    g = golden("sim3_chain")
    s, R, t = float(g["s"]), g["R"], g["t"]
    same(rp.apply_sim3(g["P4"], s, R, t), g["S4"])
    same(rp.apply_sim3(g["P2"], s, R, t), g["S2"])
    assert g["S4"].dtype == np.float64                # f32 in -> f64 out
    chain = [(float(a), b, c) for a, b, c in zip(g["chain_s"], g["chain_R"], g["chain_t"])]
    acc = rp.accumulate_sim3(chain)
    assert len(acc) == len(chain) + 1
    same(np.array([a[0] for a in acc]), g["acc_s"])
    same(np.stack([a[1] for a in acc]), g["acc_R"])
    same(np.stack([a[2] for a in acc]), g["acc_t"])
    assert len(rp.accumulate_sim3([])) == int(g["acc_empty_len"]) == 0
    same(rp.rebase_extrinsic_sim3(g["E_local"][2], s, R, t), g["rebase"])
    E = g["E_local"].astype(np.float64)
    same(rp.chain_extrinsics_from_overlap(E[4], E, g["T"]), g["chain_overlap"])
    # the frame-to-frame formulation of utils/align_geometry_single.py agrees
    close(rp.chain_extrinsics_single_overlap(E[4], E, g["T"][:3, :3], g["T"][:3, 3]), g["chain_overlap"], 1e-12)


def test_chunking(golden):
    g = golden("chunks")
    for key in g.files:
        if not key.startswith("chunks_"):
            continue
        n, c, o = (int(x) for x in key.split("_")[1:])
        ch = rp.image_chunks(list(range(n)), c, o)
        same(np.array([x[0] for x in ch], dtype=np.int64), g[key])
        same(np.array([len(x) for x in ch], dtype=np.int64), g[f"chunklen_{n}_{c}_{o}"])
    assert len(rp.image_chunks(list(range(300)), 16, 1)) == 20
    assert len(rp.image_chunks(list(range(2000)), 32, 1)) == 65
    assert len(rp.image_chunks(list(range(2000)), 64, 1)) == 32
    assert len(rp.solver_chunk_starts(300, 16, 1)) == 19      # solver.py deque semantics
    assert rp.solver_chunk_starts(300, 16, 1)[-1] == 270


def test_overlap_extraction(golden):
    g = golden("overlap")
    prev = {k[5:]: g[k] for k in g.files if k.startswith("prev_")}
    cur = {k[4:]: g[k] for k in g.files if k.startswith("cur_")}
    same(rp.unproject_f32(prev["depth"][-1:], prev["intrinsics"][-1:], prev["extrinsics"][-1:]), g["X_root_prev"])
    same(rp.unproject_f32(cur["depth"][:1], cur["intrinsics"][:1], cur["extrinsics"][:1]), g["X_root_cur"])
    same(g["X_single_prev"], g["X_root_prev"])
    same(rp.unproject_world_f64(prev["depth"][-2:], prev["intrinsics"][-2:], prev["extrinsics"][-2:]), g["X_align_pm1"])
    same(rp.unproject_world_f64(cur["depth"][:2], cur["intrinsics"][:2], cur["extrinsics"][:2]), g["X_align_pm2"])


def test_percentile_and_median_from_order_stats():
    rng = np.random.default_rng(5)
    for n in (1, 2, 5, 100, 1001, 50000):
        a = (rng.random(n) * 5).astype(np.float32)
        srt = np.sort(a)
        for p in (0, 10, 33.3, 50, 65, 99.9, 100):
            i, j, gam = rp.percentile_indices_f32(n, p)
            same(rp.percentile_from_order_stats(srt[i], srt[j], gam), np.percentile(a, p))
        same(rp.median_from_order_stats(srt[(n - 1) // 2], srt[n // 2]), np.median(a))


def test_viewer_filter_restatement():
    # V has no importable reference (viser absent): restated from viewer.py:198-218,333-338
    rng = np.random.default_rng(9)
    d = (rng.random((10, 14)) * 3).astype(np.float32)
    d[0, :3] = 0.0
    c = rng.random((10, 14)).astype(np.float32)
    K = np.array([[12, 0, 6.5], [0, 12, 4.5], [0, 0, 1]], np.float32)
    E = np.eye(4, dtype=np.float32)[:3]
    pts, cf, pix = rp.viewer_frame_points(d, c, E, K, vis_stride=2)
    assert pts.dtype == np.float64 and len(pts) == len(cf) == len(pix)
    assert (pts[:, 2] > 0.1).all() and (pts[:, 2] < 50).all()
    m, thr = rp.viewer_conf_mask(cf, 65)
    assert thr.dtype == np.float32 and m.sum() == (cf >= thr).sum()
    m2, _ = rp.viewer_conf_mask(np.zeros(4, np.float32), 65)
    assert m2.all()


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_live_reference_random_inputs():
    """Beyond the frozen fixtures: fresh random inputs straight against the reference."""
    ref = ref_loader.load()
    rng = np.random.default_rng(77)
    for _ in range(5):
        src = rng.normal(0, 1, (257, 3))
        dst = rng.normal(0, 1, (257, 3))
        w = rng.random(257).astype(np.float32)
        a = ref.al.weighted_umeyama_alignment(src, dst, w)
        b = rp.weighted_umeyama(src, dst, w)
        same(np.float64(a[0]), np.float64(b[0])); same(a[1], b[1]); same(a[2], b[2])
        a = ref.ag._umeyama_sim3(src, dst)
        b = rp.umeyama_sim3(src, dst)
        same(np.float64(a[0]), np.float64(b[0])); same(a[1], b[1]); same(a[2], b[2])
