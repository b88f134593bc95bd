"""CPU checks of the oracle's own (unpinned) specs: numpy statement == C restatement,
known-answer recovery, voxel invariants.  No GPU."""
import numpy as np

from da3slam_b200 import synth
from oracle import ref_port as rp
from oracle import spec_port as sp


def test_numpy_and_c_statements_agree():
    rng = np.random.default_rng(0)
    H, W = 40, 48
    A_, B_, gt = synth.make_pair(H, W, frames=2, seed=3, outlier_ratio=0.3)
    for world in (True, False):
        corr = sp.pair_correspondences(A_, B_, 1, world)
        xs, ys = sp.ransac_points(corr, world)
        si = rng.integers(0, H * W, size=(64, 3))
        A, T, ok, _ = sp.ransac_hypotheses(xs, ys, corr["mask"], si)
        c_np = sp.ransac_score(A, T, ok, xs, ys, corr["mask"], 0.02)
        c_c = sp.ransac_score_c(A, T, ok, xs, ys, corr["mask"], 0.02)
        assert np.array_equal(c_np, c_c)
        best, _ = sp.ransac_best(c_np, ok)
        assert np.array_equal(sp.ransac_inlier_mask(A, T, best, xs, ys, corr["mask"], 0.02),
                              sp.ransac_inlier_mask_c(A, T, best, xs, ys, corr["mask"], 0.02))
    assert np.array_equal(sp.cam_fast_f32(A_["depth"], A_["intrinsics"]), sp.cam_fast_f32_c(A_["depth"], A_["intrinsics"]))


def test_fast_unprojection_tracks_the_reference_paths():
    rng = np.random.default_rng(1)
    d = synth.smooth_depth(rng, 2, 30, 36)
    K = synth.make_intrinsics(2, 30, 36)
    E = synth.trajectory_w2c(rng, 2).astype(np.float32)
    fast = sp.cam_fast_f32(d, K)
    assert np.abs(fast - rp.unproject_f32(d, K, E, "camera")).max() < 2e-6          # vs align_geometry.py:192-256
    _, cam, _ = rp.unproject_world_vggt(d, E, K)
    assert np.abs(fast - cam).max() < 1e-6                                            # vs src/vggt (closed form)
    w64 = sp.world_from_cam_f64(fast, E)
    assert np.abs(w64 - rp.unproject_world_f64(d, K, E)).max() < 5e-6                 # vs utils/geometry.py:4-40


def test_dense_irls_recovers_ground_truth_and_matches_reference_when_masks_agree():
    A_, B_, (s, R, t) = synth.make_pair(48, 64, frames=2, seed=5)
    o = sp.align_pair(A_, B_, 1, world=True)
    assert o["status"] == 0 and abs(o["s"] - s) < 1e-3 and np.abs(o["R"] - R).max() < 1e-3
    # on inputs where the reference's two independent masks coincide and nothing is subsampled,
    # the dense joint-mask IRLS *is* the reference IRLS (utils/align.py:111-218)
    corr = sp.pair_correspondences(A_, B_, 1, True)
    conf = np.sqrt(A_["conf"][-1:] * B_["conf"][:1])                                # same conf on both sides
    n = int((conf.reshape(-1) > rp.irls_conf_threshold(conf.reshape(-1), conf.reshape(-1))).sum())
    if n <= 5000:
        pm1 = corr["y"].reshape(1, 48, 64, 3)
        pm2 = corr["x"].reshape(1, 48, 64, 3)
        ref = rp.irls_reference(pm1, pm2, conf, conf, indices=np.arange(n))
        thr = rp.irls_conf_threshold(conf.reshape(-1), conf.reshape(-1))
        mine = sp.irls_dense(corr["x"], corr["y"], np.sqrt(conf * conf).reshape(-1), conf.reshape(-1) > thr)
        assert abs(ref[0] - mine[0]) < 1e-12 and np.abs(ref[1] - mine[1]).max() < 1e-12


def test_voxel_invariants():
    rng = np.random.default_rng(2)
    p = rng.normal(0, 1, (20000, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (20000, 3), dtype=np.uint8)
    xyz, col, cnt, key = sp.voxel_downsample(p, 0.1, rgb)
    assert cnt.sum() == 20000 and np.all(np.diff(key) > 0)
    # every mean lies inside its voxel; permutation invariance (integer accumulation)
    k = np.floor(xyz.astype(np.float64) / np.float64(np.float32(0.1)))
    _, _, kk, _ = sp.voxel_keys(xyz, 0.1)
    perm = rng.permutation(20000)
    xyz2, col2, cnt2, key2 = sp.voxel_downsample(p[perm], 0.1, rgb[perm])
    assert np.array_equal(xyz, xyz2) and np.array_equal(col, col2) and np.array_equal(cnt, cnt2) and np.array_equal(key, key2)
    # idempotence of the voxel SET: downsampling the means again keeps one point per voxel
    xyz3, _, cnt3, key3 = sp.voxel_downsample(xyz, 0.1)
    assert len(key3) <= len(key) and cnt3.sum() == len(key)
    # the C key function agrees with the numpy one
    from oracle import build as ob
    import ctypes as C
    lib = ob.load()
    keys_np, usable, _, frac = sp.voxel_keys(p[:500], 0.1)
    for i in range(500):
        kk_ = C.c_int64()
        q = (C.c_int64 * 3)()
        pi = np.ascontiguousarray(p[i])
        ok = lib.oracle_voxel_key(pi.ctypes.data, np.float32(0.1), C.byref(kk_), q)
        assert bool(ok) == bool(usable[i])
        if ok:
            assert kk_.value == keys_np[i]
            assert list(q) == [int(v) for v in np.rint(frac[i] * sp.VOX_FRAC)]
