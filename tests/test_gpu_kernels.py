"""GPU parity tests: every kernel of libda3s.so against the CPU oracle, through the C ABI
(da3slam_b200.ops is a thin ctypes layer).  Bit-exact for masks / counts / order
statistics / inliers / voxel keys; 1e-6 relative (stated per test) for floating results.
"""
import numpy as np
import pytest
import torch

from da3slam_b200 import _lib as L
from da3slam_b200 import ops, synth
from da3slam_b200.pipeline import DeviceSubmap
from oracle import ref_port as rp
from oracle import spec_port as sp

pytestmark = pytest.mark.gpu

REL = 1e-6          # north-star tolerance for scale / rotation / translation


def dev_t(x, cuda, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(cuda)
    return t if dtype is None else t.to(dtype)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


# ----------------------------------------------------------------------------------------
# K1 unprojection + filtering
# ----------------------------------------------------------------------------------------
def test_unproject_against_reference_golden(golden, cuda):
    g = golden("unproject")
    d, c, K, E = (dev_t(g[k], cuda) for k in ("depth", "conf", "K", "E"))
    cams = ops.build_cams(K, E)
    # closed form, camera frame: BIT-EXACT vs src/vggt/utils/geometry.py:86-116
    xyz, _, _ = ops.unproject_filter(d[1:2], None, cams[1:2], mode="closed", world=False, want_mask=False, want_count=False)
    assert np.array_equal(xyz[0].cpu().numpy(), g["u3_cam_1"])
    # closed form, world, float64 out vs VGGT world.  VGGT evaluates -R^T t in float32
    # (src/vggt/utils/geometry.py:150-156 on float32 extrinsics) where this library uses float64,
    # so the two differ at float32 rounding of the translation (~1e-8), inside the 1e-6 contract
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode="closed", world=True, out_f64=True, want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), g["u3_world"]) < REL
    cam64, _, _ = ops.unproject_filter(d, None, cams, mode="closed", world=False, out_f64=True, want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), sp.world_from_cam_f64(cam64.cpu().numpy().astype(np.float32), g["E"])) < 1e-14
    # general K^-1 path vs the float64 numpy reference (utils/geometry.py:4-40)
    cams_g = ops.build_cams(K, E, general_inverse=True)
    xyz, _, _ = ops.unproject_filter(d, None, cams_g, mode="kinv", world=True, out_f64=True, want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), g["u2_world"]) < REL
    # float32 outputs vs the float32 torch reference (align_geometry.py:192-256)
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode="closed", world=False, want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), g["u1_camera"]) < REL
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode="fast", world=True, want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), g["u1_world"]) < 2e-6


@pytest.mark.parametrize("H,W", [(20, 28), (37, 41), (518, 518)])
def test_unproject_fast_bit_exact_and_masks(cuda, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    n = 3
    depth = synth.smooth_depth(rng, n, H, W)
    depth[0, 0, :5] = 0.0
    depth[1, 1, 2] = np.nan
    depth[2, 2, 3] = np.inf
    conf = synth.da3_like_conf(rng, n, H, W)
    K = synth.make_intrinsics(n, H, W)
    E = synth.trajectory_w2c(rng, n).astype(np.float32)
    d, c = dev_t(depth, cuda), dev_t(conf, cuda)
    cams = ops.build_cams(dev_t(K, cuda), dev_t(E, cuda))
    thr = np.float32(0.37)
    # SPEC 1: float32 camera points, bit-exact vs the oracle statement
    xyz, mask, cnt = ops.unproject_filter(d, c, cams, mode="fast", world=False, conf_cmp=">", conf_thr=float(thr), depth_eps=1e-6)
    ref = sp.cam_fast_f32(depth, K)
    got = xyz.cpu().numpy()
    assert np.array_equal(got, ref, equal_nan=True)
    ref_mask = (conf > thr) & (depth > np.float32(1e-6)) & np.isfinite(depth)
    assert np.array_equal(mask.cpu().numpy(), ref_mask)
    assert int(cnt.item()) == int(ref_mask.sum())
    # '>=' with a positive floor (viewer.py:334-336) and the threshold passed from device memory
    thr_dev = torch.tensor([float(thr)], dtype=torch.float32, device=cuda)
    _, mask, cnt = ops.unproject_filter(d, c, cams, mode="fast", conf_cmp=">=", conf_thr_dev=thr_dev, conf_floor=0.0)
    ref_mask = (conf >= thr) & (conf > 0)
    assert np.array_equal(mask.cpu().numpy(), ref_mask) and int(cnt.item()) == int(ref_mask.sum())
    # threshold exactly equal to a data value: '>' and '>=' must differ exactly there
    v = conf[0, H // 2, W // 2]
    _, m_gt, _ = ops.unproject_filter(d, c, cams, mode="fast", conf_cmp=">", conf_thr=float(v))
    _, m_ge, _ = ops.unproject_filter(d, c, cams, mode="fast", conf_cmp=">=", conf_thr=float(v))
    assert np.array_equal(m_gt.cpu().numpy(), conf > v) and np.array_equal(m_ge.cpu().numpy(), conf >= v)
    # world mode in float32 (SPEC 4 world points)
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode="fast", world=True, want_mask=False, want_count=False)
    finite = np.isfinite(depth)
    ref_w = sp.world_from_cam_f32(np.where(finite[..., None], ref, 0), E)
    assert np.array_equal(xyz.cpu().numpy()[finite], ref_w[finite])


def test_unproject_fused_sim3_and_viewer_validity(cuda):
    rng = np.random.default_rng(3)
    n, H, W = 4, 24, 32
    depth = synth.smooth_depth(rng, n, H, W)
    depth[0, :2] = 0.01                                   # z < 0.1 after unprojection with identity-ish pose
    conf = synth.da3_like_conf(rng, n, H, W)
    K = synth.make_intrinsics(n, H, W)
    E = synth.trajectory_w2c(rng, n).astype(np.float32)
    d, c = dev_t(depth, cuda), dev_t(conf, cuda)
    cams = ops.build_cams(dev_t(K, cuda), dev_t(E, cuda))
    _, cam, _ = rp.unproject_world_vggt(depth, E, K)
    world = sp.world_from_cam_f64(cam, E)                  # VGGT camera points, float64 closed-form c2w (see above)
    assert rel_err(world, rp.unproject_world_vggt(depth, E, K)[0]) < REL
    # viewer.py:214-218 validity on world z
    xyz, mask, _ = ops.unproject_filter(d, c, cams, mode="closed", world=True, out_f64=True, world_z=True)
    ref_mask = (world[..., 2] > 0.1) & (world[..., 2] < 50.0) & np.all(np.isfinite(world), axis=-1)
    assert rel_err(xyz.cpu().numpy(), world) < 1e-12
    assert np.array_equal(mask.cpu().numpy(), ref_mask)
    # one Sim(3) for all frames, and one per frame, fused behind the unprojection
    s, R, t = synth.random_sim3(rng)
    row = ops.sim3_row(s, R, t, cuda)
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode="closed", world=True, out_f64=True, sim3=row, want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), rp.apply_sim3(world, s, R, t)) < 1e-12
    rows, ref = [], []
    for f in range(n):
        s, R, t = synth.random_sim3(rng)
        rows.append(ops.sim3_row(s, R, t, cuda))
        ref.append(rp.apply_sim3(world[f], s, R, t))
    xyz, _, _ = ops.unproject_filter(d, None, cams, mode="closed", world=True, out_f64=True, sim3=torch.stack(rows), want_mask=False, want_count=False)
    assert rel_err(xyz.cpu().numpy(), np.stack(ref)) < 1e-12


def test_apply_sim3_golden_and_sizes(golden, cuda):
    g = golden("sim3_chain")
    row = ops.sim3_row(float(g["s"]), g["R"], g["t"], cuda)
    out = ops.apply_sim3(dev_t(g["P4"], cuda), row)
    assert out.dtype == torch.float64 and rel_err(out.cpu().numpy(), g["S4"]) < 1e-14
    out = ops.apply_sim3(dev_t(g["P2"], cuda), row)
    assert rel_err(out.cpu().numpy(), g["S2"]) < 1e-14
    rng = np.random.default_rng(0)
    for n in (1, 3, 4, 5, 1023, 1024, 1025, 100003):
        p = rng.normal(0, 2, (n, 3)).astype(np.float32)
        out = ops.apply_sim3(dev_t(p, cuda), row, out_f64=False)
        ref = rp.apply_sim3(p, float(g["s"]), g["R"], g["t"])
        assert out.dtype == torch.float32 and np.abs(out.cpu().numpy() - ref).max() < 1e-6 * 8


# ----------------------------------------------------------------------------------------
# exact selection
# ----------------------------------------------------------------------------------------
def test_select_median_percentile_bit_exact(cuda):
    rng = np.random.default_rng(11)
    segs, refs = [], []
    for n in (1, 2, 3, 4, 7, 100, 4095, 4096, 4097, 268324):
        a = (np.exp(rng.normal(0, 0.75, n))).astype(np.float32)
        if n > 50:
            a[rng.integers(0, n, n // 20)] = 0.0          # ties and zeros
            a[rng.integers(0, n, 5)] = -1.5               # negatives
        t = dev_t(a, cuda)
        segs.append(dict(a=t, stat=L.SEL_MEDIAN)); refs.append(np.median(a))
        for p in (0.0, 10.0, 50.0, 65.0, 99.9, 100.0):
            segs.append(dict(a=t, stat=L.SEL_PERCENTILE, percent=p)); refs.append(np.percentile(a, p))
        pos = a[a > 0]
        if len(pos):
            segs.append(dict(a=t, kind=L.SEL_POSITIVE, stat=L.SEL_PERCENTILE, percent=65.0))
            refs.append(np.percentile(pos, 65.0))          # viewer.py:334-335
    out = ops.select(segs, cuda)
    for i, r in enumerate(refs):
        assert out["value"][i] == np.float32(r), (i, out[i], r)
    # all-equal data, and an empty positive set
    z = dev_t(np.zeros(1000, np.float32), cuda)
    out = ops.select([dict(a=z, stat=L.SEL_MEDIAN), dict(a=z, kind=L.SEL_POSITIVE, stat=L.SEL_PERCENTILE, percent=65.0)], cuda)
    assert out["value"][0] == 0 and out["n_valid"][1] == 0 and np.isnan(out["value"][1])


def test_select_sorted_and_smooth_segments(cuda):
    """Values concentrated around the wanted rank in one part of the segment (sorted input, a smooth ramp, a plateau at
    the median): the block that holds them cannot stage all its candidates in shared memory and writes them to the
    candidate array directly — the answers stay exact."""
    rng = np.random.default_rng(17)
    n = 700_001
    ramp = np.linspace(0.5, 3.0, n).astype(np.float32)
    cases = [np.sort(rng.normal(1, 0.3, n).astype(np.float32)), ramp, ramp[::-1].copy(),
             np.where(np.arange(n) % 3 == 0, np.float32(1.25), rng.random(n).astype(np.float32) * 2.5).astype(np.float32),
             np.repeat(rng.random(n // 5000 + 1).astype(np.float32), 5000)[:n]]
    segs, want = [], []
    for a in cases:
        for stat, pct in ((L.SEL_MEDIAN, 0.0), (L.SEL_PERCENTILE, 65.0), (L.SEL_PERCENTILE, 3.0)):
            segs.append(dict(a=dev_t(a, cuda), kind=L.SEL_VALUES, stat=stat, percent=pct))
            want.append(np.median(a) if stat == L.SEL_MEDIAN else np.percentile(a, pct))
    out = ops.select(segs, cuda)
    for i, w_ in enumerate(want):
        assert out["n_valid"][i] == n and np.float32(out["value"][i]) == np.float32(w_), i


def test_select_depth_ratio_median(golden, cuda):
    g = golden("depth_scale")
    segs, refs, counts = [], [], []
    for case in range(4):
        dA, cA, dB, cB = (g[f"{k}{case}"] for k in ("dA", "cA", "dB", "cB"))
        segs.append(dict(a=dev_t(dA[-1], cuda), b=dev_t(dB[0], cuda), ca=dev_t(cA[-1], cuda), cb=dev_t(cB[0], cuda),
                         kind=L.SEL_RATIO, stat=L.SEL_MEDIAN, conf_th=0.2, eps=1e-6))
        refs.append(g[f"plain{case}"])
        m = (dA[-1] > 1e-6) & (dB[0] > 1e-6) & np.isfinite(dA[-1]) & np.isfinite(dB[0]) & (cA[-1] > 0.2) & (cB[0] > 0.2)
        counts.append(int(m.sum()))
    out = ops.select(segs, cuda)
    for i in range(4):
        assert out["n_valid"][i] == counts[i]
        assert np.float64(out["value"][i]) == refs[i]      # align_geometry.py:329-330, bit-exact


# ----------------------------------------------------------------------------------------
# pair alignment
# ----------------------------------------------------------------------------------------
def make_dev_pairs(subs, cuda, overlap):
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    from da3slam_b200.pipeline import pair_entry
    entries = [pair_entry(dsubs[k], dsubs[k + 1], overlap) for k in range(len(dsubs) - 1)]
    return dsubs, ops.make_pairs(entries, cuda), len(entries)


def check_rows_against_oracle(rows, subs, overlap, rtol=REL, **kw):
    rows = rows.cpu().numpy()
    for k in range(len(subs) - 1):
        o = sp.align_pair(subs[k], subs[k + 1], overlap=overlap, **kw)
        r = rows[k]
        assert int(r[15]) == o["status"], (k, r[15], o["status"])
        assert int(r[13]) == o["n_valid"], (k, r[13], o["n_valid"])            # bit-exact mask count
        if o["status"] == 0:
            assert int(r[14]) == o["iters"], (k, r[14], o["iters"])
            assert abs(r[0] - o["s"]) <= rtol * abs(o["s"])
            assert rel_err(r[1:10].reshape(3, 3), o["R"]) <= rtol
            assert np.abs(r[10:13] - o["t"]).max() <= rtol * max(1.0, np.abs(o["t"]).max())
        else:
            assert r[0] == 1.0 and np.array_equal(r[1:10].reshape(3, 3), np.eye(3)) and not r[10:13].any()


@pytest.mark.parametrize("world", [True, False])
@pytest.mark.parametrize("overlap", [1, 2])
def test_align_pairs_irls_vs_oracle(cuda, world, overlap):
    subs, gt = synth.make_sequence(4, 3, 40, 52, overlap=overlap, seed=21 + overlap)
    keep_alive, table, n = make_dev_pairs(subs, cuda, overlap)
    opts = L.default_opts(world=int(world))
    rows, aux, _ = ops.align_pairs(table, n, overlap, 40, 52, opts, want_aux=True)
    check_rows_against_oracle(rows, subs, overlap, world=world)
    aux = aux.cpu().numpy()
    for k in range(n):
        c = sp.pair_correspondences(subs[k], subs[k + 1], overlap, world)
        assert np.float32(aux[k, 0]) == c["thr"]                               # utils/align.py:142, bit-exact
    if world:                                                                   # the estimate recovers ground truth
        r = rows.cpu().numpy()
        for k in range(n):
            assert abs(r[k, 0] - gt[k][0]) < 2e-3 and np.abs(r[k, 1:10].reshape(3, 3) - gt[k][1]).max() < 2e-3


def test_align_pairs_huber_tail_and_single_solve(cuda):
    # gross outliers put many residuals above delta so the Huber branch matters
    subs, _ = synth.make_sequence(3, 2, 48, 64, overlap=1, seed=5, outlier_ratio=0.25)
    keep_alive, table, n = make_dev_pairs(subs, cuda, 1)
    for delta in (1.0, 0.1):
        opts = L.default_opts(world=1, huber_delta=delta)
        rows, _, _ = ops.align_pairs(table, n, 1, 48, 64, opts)
        check_rows_against_oracle(rows, subs, 1, world=True, delta=delta)
    opts = L.default_opts(world=1, huber=0)
    rows, _, _ = ops.align_pairs(table, n, 1, 48, 64, opts)
    check_rows_against_oracle(rows, subs, 1, world=True, huber=False)
    # max_iterations cap is honoured
    opts = L.default_opts(world=1, max_iterations=2, tol=0.0)
    rows, _, _ = ops.align_pairs(table, n, 1, 48, 64, opts)
    check_rows_against_oracle(rows, subs, 1, world=True, max_iterations=2, tol=0.0)


def test_align_pairs_depth_scale_and_too_few(cuda):
    subs, _ = synth.make_sequence(3, 2, 36, 44, overlap=1, seed=9)
    subs[2]["conf"][0] = 0.0                                # pair 1: nothing passes the threshold
    keep_alive, table, n = make_dev_pairs(subs, cuda, 1)
    opts = L.default_opts(world=0, depth_scale_mode=1)
    rows, aux, _ = ops.align_pairs(table, n, 1, 36, 44, opts, want_aux=True)
    check_rows_against_oracle(rows, subs, 1, world=False, use_depth_scale=True)
    aux = aux.cpu().numpy()
    assert np.float32(aux[0, 1]) == np.float32(rp.depth_scale_guarded(subs[0], subs[1]))   # bit-exact median ratio
    assert aux[1, 1] == 1.0                                                                   # < 50 valid -> 1.0
    assert int(rows[1, 15].item()) == 1


def test_align_pairs_batch_order_invariance(cuda):
    subs, _ = synth.make_sequence(5, 2, 32, 40, overlap=1, seed=2)
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    from da3slam_b200.pipeline import pair_entry
    entries = [pair_entry(dsubs[k], dsubs[k + 1], 1) for k in range(4)]
    opts = L.default_opts(world=1)
    rows_a, _, _ = ops.align_pairs(ops.make_pairs(entries, cuda), 4, 1, 32, 40, opts)
    perm = [2, 0, 3, 1]
    rows_b, _, _ = ops.align_pairs(ops.make_pairs([entries[i] for i in perm], cuda), 4, 1, 32, 40, opts)
    rows_c, _, _ = ops.align_pairs(ops.make_pairs(entries[1:3], cuda), 2, 1, 32, 40, opts)
    a, b, c = rows_a.cpu().numpy(), rows_b.cpu().numpy(), rows_c.cpu().numpy()
    assert np.array_equal(a[perm], b)          # bit-identical regardless of batch order ...
    assert np.array_equal(a[1:3], c)           # ... and of how the batch is sharded (multi-GPU contract)


def test_ransac_stages_bit_exact(cuda):
    H, W, n_hyp = 40, 48, 96
    subs, gt = synth.make_sequence(3, 2, H, W, overlap=1, seed=31, outlier_ratio=0.3)
    keep_alive, table, n = make_dev_pairs(subs, cuda, 1)
    rng = np.random.default_rng(7)
    si = rng.integers(0, H * W, size=(n, n_hyp, 3)).astype(np.int32)
    si[0, 5] = [3, 3, 9]                                    # repeated pixel -> invalid hypothesis
    opts = L.default_opts(world=1)
    thr, ds, _ = ops.pair_thresholds(table, n, 1, H, W, opts)
    for world in (True, False):
        A, t, ok, s3 = ops.ransac_hypotheses(table, n, 1, H, W, thr, ds, dev_t(si, cuda), world=world)
        counts = ops.ransac_score(table, n, 1, H, W, thr, ds, A, t, ok, 0.02, world=world)
        A_h, t_h, ok_h, s3_h, cnt_h = (x.cpu().numpy() for x in (A, t, ok, s3, counts))
        for k in range(n):
            corr = sp.pair_correspondences(subs[k], subs[k + 1], 1, world)
            xs, ys = sp.ransac_points(corr, world)
            oA, oT, ook, osim = sp.ransac_hypotheses(xs, ys, corr["mask"], si[k])
            assert np.array_equal(ok_h[k].astype(bool), ook)                    # validity flags: exact
            v = ook
            assert rel_err(s3_h[k][v], osim[v]) < 1e-9                           # 3-point Umeyama vs LAPACK route
            # scoring: feed the GPU's own float32 hypotheses to the oracle -> counts must be bit-exact
            ref_counts = sp.ransac_score(A_h[k].reshape(-1, 3, 3), t_h[k], ok_h[k].astype(bool), xs, ys, corr["mask"], 0.02)
            assert np.array_equal(cnt_h[k], ref_counts)
            best, nbest = sp.ransac_best(ref_counts, ook, 20)
            assert best >= 0
            m = ops.ransac_inlier_mask(table[k * L.PAIR_BYTES:(k + 1) * L.PAIR_BYTES], 1, 1, H, W, thr[k:k + 1], ds[k:k + 1],
                                       A[k, best:best + 1], t[k, best:best + 1], ok[k, best:best + 1], 0.02, world=world)
            ref_m = sp.ransac_inlier_mask(A_h[k].reshape(-1, 3, 3), t_h[k], best, xs, ys, corr["mask"], 0.02)
            assert np.array_equal(m.cpu().numpy()[0], ref_m) and int(ref_m.sum()) == nbest


def test_align_pairs_with_ransac_end_to_end(cuda):
    H, W, n_hyp = 40, 48, 128
    subs, gt = synth.make_sequence(3, 2, H, W, overlap=1, seed=33, outlier_ratio=0.3)
    keep_alive, table, n = make_dev_pairs(subs, cuda, 1)
    rng = np.random.default_rng(8)
    si = rng.integers(0, H * W, size=(n, n_hyp, 3)).astype(np.int32)
    opts = L.default_opts(world=1, n_hyp=n_hyp, ransac_thr=0.02)
    rows, aux, counts = ops.align_pairs(table, n, 1, H, W, opts, dev_t(si, cuda), want_aux=True, want_counts=True)
    rows, aux, counts = rows.cpu().numpy(), aux.cpu().numpy(), counts.cpu().numpy()
    for k in range(n):
        o = sp.align_pair(subs[k], subs[k + 1], overlap=1, world=True, ransac=dict(sample_idx=si[k], thr=0.02))
        # hypothesis tables can differ in the last float32 bit between the device Jacobi SVD and LAPACK, so the
        # end-to-end check is on the decisions and the refined transform, not on every count
        assert int(aux[k, 4]) == o["best"] and int(aux[k, 5]) == o["best_count"]
        assert int(rows[k, 13]) == o["n_valid"]
        assert abs(rows[k, 0] - o["s"]) <= REL * o["s"] and rel_err(rows[k, 1:10].reshape(3, 3), o["R"]) <= REL
        assert abs(rows[k, 0] - gt[k][0]) < 5e-3              # RANSAC + IRLS recovers ground truth despite 30 % outliers
    # no model: threshold so tight that fewer than min_inliers agree
    opts = L.default_opts(world=1, n_hyp=n_hyp, ransac_thr=1e-7)
    rows, _, _ = ops.align_pairs(table, n, 1, H, W, opts, dev_t(si, cuda))
    assert (rows[:, 15].cpu().numpy() == 2).all() and (rows[:, 0].cpu().numpy() == 1.0).all()


# ----------------------------------------------------------------------------------------
# array-level Umeyama / IRLS against the reference's golden outputs
# ----------------------------------------------------------------------------------------
def row_close(row, s, R, t, tol=REL):
    r = row.cpu().numpy()
    assert abs(r[0] - s) <= tol * abs(s), (r[0], s)
    assert rel_err(r[1:10].reshape(3, 3), R) <= tol
    assert np.abs(r[10:13] - t).max() <= tol * max(1.0, np.abs(t).max())


def test_umeyama_points_against_reference_golden(golden, cuda):
    g = golden("umeyama")
    src, dst, w = dev_t(g["src"], cuda), dev_t(g["dst"], cuda), dev_t(g["w"], cuda)
    row_close(ops.umeyama_points(src, dst, w), float(g["W_s"]), g["W_R"], g["W_t"])
    row_close(ops.umeyama_points(src.float(), dst.float(), w), float(g["W32_s"]), g["W32_R"], g["W32_t"])
    row_close(ops.umeyama_points(src, dev_t(g["dst_m"], cuda), w), float(g["Wm_s"]), g["Wm_R"], g["Wm_t"])    # reflection fix
    row_close(ops.umeyama_points(src, dst, None, L.UMEYAMA_MEAN), float(g["U_s"]), g["U_R"], g["U_t"])
    row_close(ops.umeyama_points(src, dev_t(g["dst_m"], cuda), None, L.UMEYAMA_MEAN), float(g["Um_s"]), g["Um_R"], g["Um_t"])
    row_close(ops.umeyama_points(src[:3].contiguous(), dst[:3].contiguous(), None, L.UMEYAMA_MEAN),
              float(g["U3_s"]), g["U3_R"], g["U3_t"], tol=1e-9)                                              # rank-2 covariance
    pm1, pm2 = dev_t(g["pm1"], cuda), dev_t(g["pm2"], cuda)
    # utils/align.py:224: first argument ("point_map2") is the TARGET, second the source
    row_close(ops.umeyama_points(pm2.view(-1, 3), pm1.view(-1, 3), None, L.UMEYAMA_NORMRATIO), float(g["N_s"]), g["N_R"], g["N_t"])


def test_irls_points_against_reference_golden(golden, cuda):
    g = golden("irls")
    pm1, pm2 = g["pm1"].reshape(-1, 3), g["pm2"].reshape(-1, 3)
    for tag, (a1, a2) in {"same": (g["c1"], g["c1"]), "indep": (g["c1"], g["c2"])}.items():
        c1, c2 = a1.reshape(-1), a2.reshape(-1)
        thr = g[f"{tag}0_thr"]
        nz1, nz2 = np.flatnonzero(c1 > thr), np.flatnonzero(c2 > thr)       # utils/align.py:145-151
        for seed in (0, 1):
            idx = g[f"{tag}{seed}_idx"]
            row = ops.irls_points(dev_t(pm2, cuda), dev_t(pm1, cuda), dev_t(c2, cuda), dev_t(c1, cuda),
                                  idx_src=dev_t(nz2[idx], cuda), idx_dst=dev_t(nz1[idx], cuda))
            row_close(row, float(g[f"{tag}{seed}_s"]), g[f"{tag}{seed}_R"], g[f"{tag}{seed}_t"])


# ----------------------------------------------------------------------------------------
# voxel grid
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 1000, 200003])
def test_voxel_downsample_bit_exact(cuda, n):
    rng = np.random.default_rng(n + 1)
    pts = rng.normal(0, 1.0, (n, 3)).astype(np.float32)
    if n >= 1000:
        pts[:200] = pts[0] + rng.normal(0, 1e-4, (200, 3)).astype(np.float32)   # heavy collisions in one voxel
        pts[300] = [np.nan, 0, 0]
        pts[301] = [1e30, 0, 0]                                                  # key out of range
        pts[302] = [-0.0, 0.02, -0.02]                                           # exact voxel boundaries
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    mask = rng.random(n) < 0.9
    if n == 0:
        return
    for use_rgb, use_mask in ((True, True), (False, False)):
        out = ops.voxel_downsample([(dev_t(pts, cuda), dev_t(rgb, cuda) if use_rgb else None,
                                     dev_t(mask, cuda) if use_mask else None)], 0.02)
        xyz, col, cnt, key = sp.voxel_downsample(pts, 0.02, rgb if use_rgb else None, mask if use_mask else None)
        assert np.array_equal(out[3].cpu().numpy(), key)            # voxel set: bit-exact keys
        assert np.array_equal(out[2].cpu().numpy(), cnt)            # counts: bit-exact
        assert np.array_equal(out[0].cpu().numpy(), xyz)            # integer accumulation -> positions bit-exact too
        if use_rgb:
            assert np.array_equal(out[1].cpu().numpy(), col)


@pytest.mark.parametrize("centre", [0.0, -37.5, 812.25])
def test_voxel_dense_cloud_warp_reduction(cuda, centre):
    """Many points per voxel (the hires case): batches of 32 points with <= 4 distinct voxels merge with masked warp
    reductions on the 16-bit limbs of the 32-bit fixed-point offsets; per-voxel counts go past 2^16."""
    rng = np.random.default_rng(17)
    n = 300007
    cells = rng.integers(0, 3, (n, 3)).astype(np.float32)                      # 27 voxels, ~11 k points each
    cells[: n // 2] = np.repeat(cells[: n // 2: 64], 64, axis=0)[: n // 2]       # runs of one voxel / mixed batches
    pts = (np.float32(centre) + (cells + rng.random((n, 3), dtype=np.float32) * 0.98 + 0.01) * np.float32(0.05)).astype(np.float32)
    pts[-70000:] = np.float32(centre) + np.float32(0.051)                       # one voxel with 70 000 identical points
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    for use_rgb in (True, False):
        out = ops.voxel_downsample([(dev_t(pts, cuda), dev_t(rgb, cuda) if use_rgb else None, None)], 0.05)
        xyz, col, cnt, key = sp.voxel_downsample(pts, 0.05, rgb if use_rgb else None, None)
        assert cnt.max() >= 70000 and len(key) <= 64
        assert np.array_equal(out[3].cpu().numpy(), key) and np.array_equal(out[2].cpu().numpy(), cnt)
        assert np.array_equal(out[0].cpu().numpy(), xyz)
        if use_rgb:
            assert np.array_equal(out[1].cpu().numpy(), col)


@pytest.mark.parametrize("voxel", [0.02, 0.05, 1.0 / 3.0, 0.1, 1.0, 0.0078125, 0.3, 1e-3, 7.25])
def test_voxel_quantisation_is_the_correctly_rounded_division(cuda, voxel):
    """SPEC 5 keys and fractions come from q = f64(p) / f64(f32(voxel)) CORRECTLY ROUNDED; the kernel evaluates it with two
    Markstein corrections instead of the division routine.  Adversarial coordinates: exact multiples of the voxel size and
    their float32 neighbours (floor flips with the last bit of q), every binade from 1e-30 to the key range, both signs."""
    rng = np.random.default_rng(int(voxel * 1e6) % 9973)
    v32 = np.float32(voxel)
    k = rng.integers(-(1 << 19), 1 << 19, size=60000).astype(np.float64)
    edge = (k * np.float64(v32)).astype(np.float32)
    near = np.concatenate([edge, np.nextafter(edge, np.float32(np.inf)), np.nextafter(edge, np.float32(-np.inf))])
    mags = (10.0 ** rng.uniform(-30, np.log10(float(v32) * (1 << 19)), size=60000) * rng.choice([-1.0, 1.0], size=60000)).astype(np.float32)
    col = np.concatenate([near, mags, rng.normal(0, 5 * float(v32), 60000).astype(np.float32)])
    pts = np.stack([col, rng.permutation(col), rng.permutation(col)], axis=1)
    xyz, _, cnt, key = ops.voxel_downsample([(dev_t(pts, cuda), None, None)], float(v32), table_slots=1 << 21)
    e_xyz, _, e_cnt, e_key = sp.voxel_downsample(pts, float(v32), None, None)
    assert np.array_equal(key.cpu().numpy(), e_key) and np.array_equal(cnt.cpu().numpy(), e_cnt)
    assert np.array_equal(xyz.cpu().numpy(), e_xyz)


def test_voxel_accumulates_across_clouds(cuda):
    rng = np.random.default_rng(4)
    a = rng.normal(0, 0.5, (5000, 3)).astype(np.float32)
    b = rng.normal(0, 0.5, (7000, 3)).astype(np.float32)
    out = ops.voxel_downsample([(dev_t(a, cuda), None, None), (dev_t(b, cuda), None, None)], 0.05)
    xyz, _, cnt, key = sp.voxel_downsample(np.concatenate([a, b]), 0.05)
    assert np.array_equal(out[3].cpu().numpy(), key) and np.array_equal(out[2].cpu().numpy(), cnt)
    assert np.array_equal(out[0].cpu().numpy(), xyz)


def test_voxel_insert_jobs_one_launch(cuda):
    """da3s_voxel_insert_jobs: ragged clouds (empty, unaligned mask, sparse and dense masks) in one launch
    == the oracle on their concatenation."""
    rng = np.random.default_rng(9)
    sizes = [0, 1, 31, 129, 5000, 20011]
    keep = [1.0, 1.0, 0.5, 0.05, 0.35, 0.9]
    pts = [rng.normal(0, 0.4, (n, 3)).astype(np.float32) for n in sizes]
    rgb = [rng.integers(0, 256, (n, 3), dtype=np.uint8) for n in sizes]
    msk = [(rng.random(n) < k) for n, k in zip(sizes, keep)]
    pts[4][:300] = pts[4][0] + rng.normal(0, 1e-4, (300, 3)).astype(np.float32)
    pts[5][7] = [np.inf, 0, 0]
    grid = ops.VoxelGrid(cuda, 1 << 16, 1 << 16, True)
    backing = torch.zeros(sum(sizes) + 64, dtype=torch.uint8, device=cuda)
    clouds, off = [], 1                                        # masks at odd offsets: the unaligned path
    for p_, c_, m_ in zip(pts, rgb, msk):
        mv = backing[off:off + len(m_)]
        mv.copy_(torch.from_numpy(m_.astype(np.uint8)))
        off += len(m_) + 1
        clouds.append((dev_t(p_, cuda), dev_t(c_, cuda), mv))
    jobs = grid.make_jobs(clouds)
    for width in (0, 37):                                      # runs of 128 points / 8x16 patches of rows of 37
        grid.begin()
        grid.insert_jobs(jobs, 0.03, width=width)
        grid.finish(0.03)
        xyz, col, cnt, key = grid.read(sort=True)
        e_xyz, e_col, e_cnt, e_key = sp.voxel_downsample(np.concatenate(pts), 0.03, np.concatenate(rgb), np.concatenate(msk))
        assert np.array_equal(key.cpu().numpy(), e_key) and np.array_equal(cnt.cpu().numpy(), e_cnt)
        assert np.array_equal(xyz.cpu().numpy(), e_xyz) and np.array_equal(col.cpu().numpy(), e_col)


def test_voxel_multi_rank_merge_is_bit_exact(cuda):
    """SURVEY 8e map export: three ranks (three private tables on this one GPU) fill their grids, send every
    record to the rank that owns its key (da3s_voxel_send writes straight into the owner's inbox) and merge;
    the union of their shares == the grid of all points on one rank, bit for bit."""
    from da3slam_b200.sharding import VoxelExchange
    rng = np.random.default_rng(21)
    world, voxel = 3, 0.04
    pts = [rng.normal(0, 0.5, (n, 3)).astype(np.float32) for n in (9000, 14000, 1)]      # the same region: shared voxels
    rgb = [rng.integers(0, 256, (len(p_), 3), dtype=np.uint8) for p_ in pts]
    grids = [ops.VoxelGrid(cuda, 1 << 15, 1 << 15, True, private_ctx=True) for _ in range(world)]
    cap = 1 << 14
    ex = [VoxelExchange(cuda, world, r, cap, local=True) for r in range(world)]
    for e in ex:
        e.attach_local(ex)
    for r in range(world):
        grids[r].begin()
        grids[r].insert(dev_t(pts[r], cuda), dev_t(rgb[r], cuda), None, voxel)
    for r in range(world):
        ex[r].send(grids[r])                                                  # every send is enqueued before the first fold:
    shares = []                                                               # the device-side waits below find their flags set
    for r in range(world):
        ex[r].fold(grids[r])
        grids[r].finish(voxel)
        shares.append([t.cpu().numpy() for t in grids[r].read(sort=True)])
    assert int(sum(int(e.counts[1].sum()) for e in ex)) >= sum(len(s_[3]) for s_ in shares)       # step 1 -> parity 1
    assert all(int(v) == 1 for e in ex for v in e.flags[1].cpu())             # every rank saw every rank's arrival flag
    key = np.concatenate([s_[3] for s_ in shares])
    assert len(np.unique(key)) == len(key)                                   # every voxel has exactly one owner
    order = np.argsort(key)
    e_xyz, e_col, e_cnt, e_key = sp.voxel_downsample(np.concatenate(pts), voxel, np.concatenate(rgb), None)
    assert np.array_equal(key[order], e_key)
    assert np.array_equal(np.concatenate([s_[2] for s_ in shares])[order], e_cnt)
    assert np.array_equal(np.concatenate([s_[0] for s_ in shares])[order], e_xyz)
    assert np.array_equal(np.concatenate([s_[1] for s_ in shares])[order], e_col)
    assert min(len(s_[3]) for s_ in shares) > 0.2 * len(key) / world          # the key hash spreads the voxels
    for g in grids:
        g.ctx.close()


def test_voxel_merge_across_gpus():
    """The real thing: one process per GPU, inboxes mapped through CUDA IPC, records stored over NVLink
    (tests/multi_gpu_voxel_check.py): five merges back to back without host synchronisation, one rank held back per step.
    Needs two GPUs; profiles/r2_multi_gpu_voxel_merge_*.json are recorded runs on 2, 4 and 8 GPUs."""
    import json, os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, POINTS="500000", STEPS="5")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(here, "multi_gpu_voxel_check.py")], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    assert json.loads(line)["bit_identical_to_single_gpu"] is True


def test_errors_are_loud(cuda):
    with pytest.raises(RuntimeError):
        ops.unproject_filter(torch.zeros(1, 4, 4), None, torch.zeros(1, 200, dtype=torch.uint8))      # CPU tensor
    d = torch.zeros(1, 4, 4, device=cuda)
    cams = torch.zeros(1, L.CAM_BYTES, dtype=torch.uint8, device=cuda)
    with pytest.raises(L.Da3sError):
        ctx = ops.context(cuda)
        rc = ctx.lib.da3s_unproject_filter(ctx.h, 0, 0, 0, 1, 4, 4, 0, 0.0, 0, 0.0, 0.0, 0, 0, 0, 0, 0)
        L.check(rc, "null args")
    with pytest.raises(ValueError):
        ops.unproject_filter(d, d, cams, conf_cmp="<")


# ----------------------------------------------------------------------------------------
# whole step: SequencePlan (what bench.py times) against the oracle pipeline
# ----------------------------------------------------------------------------------------
def test_sequence_plan_end_to_end(cuda):
    from da3slam_b200.pipeline import SequencePlan
    H, W, F, n = 48, 64, 3, 4
    subs, gt = synth.make_sequence(n, F, H, W, overlap=1, seed=77, with_images=True)
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    # two-kernel export (keeps the per-point arrays, checked below) and the fused export (depth -> grid)
    # ... and: percentile thresholds selected in line vs on the side stream next to the alignment
    plan = SequencePlan(dsubs, overlap=1, voxel=0.05, conf_percentile=65.0, table_slots=1 << 16, world=1, fuse_export=False,
                        overlap_percentile=False)
    fused = SequencePlan(dsubs, overlap=1, voxel=0.05, conf_percentile=65.0, table_slots=1 << 16, world=1, fuse_export=True)
    assert fused.fuse_export and not plan.fuse_export and fused.side is not None and plan.side is None
    for _ in range(2):                                        # twice: the grid must come back clean
        plan.run()
        out = plan.read(sort=True)
        fused.run()
        out_f = fused.read(sort=True)
    rows = out["rows"]
    chain = []
    for k in range(n - 1):
        o = sp.align_pair(subs[k], subs[k + 1], overlap=1, world=True)
        assert int(rows[k, 13]) == o["n_valid"] and abs(rows[k, 0] - o["s"]) <= REL * o["s"]
        chain.append((o["s"], o["R"], o["t"]))
    acc = rp.accumulate_sim3(chain)
    cum = out["cum"]
    for k in range(n):
        assert abs(cum[k, 0] - acc[k][0]) <= REL * acc[k][0] and rel_err(cum[k, 1:10].reshape(3, 3), acc[k][1]) <= REL
        assert np.abs(cum[k, 10:13] - acc[k][2]).max() <= REL * max(1.0, np.abs(acc[k][2]).max())
    all_xyz, all_rgb, all_mask = [], [], []
    for k in range(n):
        f0 = plan.first[k]
        conf, depth = subs[k]["conf"][f0:], subs[k]["depth"][f0:]
        ref_mask, thr = rp.viewer_conf_mask(conf.reshape(-1), 65.0)         # viewer.py:333-336
        ref_mask = ref_mask & (depth.reshape(-1) > np.float32(1e-6))
        got_mask = plan.mask[k].cpu().numpy().astype(bool).reshape(-1)
        assert np.array_equal(got_mask, ref_mask)                            # exact percentile + '>=' mask: bit-exact
        cam = sp.cam_fast_f32(depth, subs[k]["intrinsics"][f0:])
        world = rp.apply_sim3(sp.world_from_cam_f64(cam, subs[k]["extrinsics"][f0:]), *acc[k])
        got = plan.xyz[k].cpu().numpy()
        assert np.abs(got - world).max() <= 2e-6 * max(1.0, np.abs(world).max())
        all_xyz.append(got.reshape(-1, 3)); all_mask.append(got_mask)
        all_rgb.append(subs[k]["processed_images"][f0:].reshape(-1, 3))
    xyz, col, cnt, key = sp.voxel_downsample(np.concatenate(all_xyz), 0.05, np.concatenate(all_rgb), np.concatenate(all_mask))
    assert np.array_equal(out["voxel_key"].cpu().numpy(), key) and np.array_equal(out["voxel_count"].cpu().numpy(), cnt)
    assert np.array_equal(out["voxel_xyz"].cpu().numpy(), xyz) and np.array_equal(out["voxel_rgb"].cpu().numpy(), col)
    # the fused kernel never writes the points, yet fills the grid with exactly the same integers
    assert np.array_equal(out_f["rows"], out["rows"]) and np.array_equal(out_f["cum"], out["cum"])
    assert np.array_equal(out_f["voxel_key"].cpu().numpy(), key) and np.array_equal(out_f["voxel_count"].cpu().numpy(), cnt)
    assert np.array_equal(out_f["voxel_xyz"].cpu().numpy(), xyz) and np.array_equal(out_f["voxel_rgb"].cpu().numpy(), col)


def test_sequence_stream_matches_resident_plan(cuda):
    """SequenceStream (upload k+1 | compute k | download k-1 on three streams, two slots) returns, for every
    sequence and in order, exactly what a SequencePlan on resident data returns."""
    from da3slam_b200.pipeline import SequencePlan, SequenceStream
    H, W, F, n = 40, 48, 3, 3
    seqs = [synth.make_sequence(n, F, H, W, overlap=1, seed=500 + i, with_images=True)[0] for i in range(5)]
    kw = dict(overlap=1, voxel=0.05, conf_percentile=65.0, table_slots=1 << 15, world=1)
    stream = SequenceStream(seqs[0], cuda, slots=2, with_keys=True, **kw)
    got = []
    for res in stream.process(seqs):
        got.append({k: (np.array(v) if v is not None and not isinstance(v, int) else v) for k, v in res.items()})   # copy: slots are reused
    assert len(got) == len(seqs)
    for subs, res in zip(seqs, got):
        plan = SequencePlan([DeviceSubmap.from_prediction(s_, cuda) for s_ in subs], **kw)
        plan.run()
        ref = plan.read(sort=True)                               # canonical order: ascending key (slot order depends on probing races)
        assert np.array_equal(res["rows"], ref["rows"]) and np.array_equal(res["cum"], ref["cum"])
        assert res["n_voxels"] == ref["voxel_key"].shape[0]
        order = np.argsort(res["voxel_key"])
        assert np.array_equal(res["voxel_key"][order], ref["voxel_key"].cpu().numpy())
        assert np.array_equal(res["voxel_xyz"][order], ref["voxel_xyz"].cpu().numpy())
        assert np.array_equal(res["voxel_rgb"][order], ref["voxel_rgb"].cpu().numpy())
        assert np.array_equal(res["voxel_count"][order], ref["voxel_count"].cpu().numpy())


def test_full_size_properties(cuda):
    """BASELINE-size frames (518 x 518, 268 324 correspondences per pair), where the oracle would take minutes:
    size-independent properties instead — ground-truth recovery, bit-identical rows under re-sharding, Sim(3)
    equivariance under an exact power-of-two depth scaling, exact order statistics against a full sort, and
    conservation of points through the voxel grid (fused and two-kernel export)."""
    from da3slam_b200.pipeline import SequencePlan, pair_entry
    H = W = 518
    subs, gt = synth.make_sequence_device(4, 3, H, W, overlap=1, seed=4321, with_images=True, device=cuda)
    dsubs = [DeviceSubmap.from_prediction(s_, cuda) for s_ in subs]
    entries = [pair_entry(dsubs[k], dsubs[k + 1], 1) for k in range(3)]
    opts = L.default_opts(world=1)
    rows, aux, _ = ops.align_pairs(ops.make_pairs(entries, cuda), 3, 1, H, W, opts, want_aux=True)
    r = rows.cpu().numpy()
    for k in range(3):                                           # (1) the estimate recovers the generating Sim(3)
        assert r[k, 15] == 0 and abs(r[k, 0] - gt[k][0]) < 1e-3 * gt[k][0]
        assert np.abs(r[k, 1:10].reshape(3, 3) - gt[k][1]).max() < 1e-3
    for k in range(3):                                           # (2) one pair alone == the same pair inside the batch
        alone, _, _ = ops.align_pairs(ops.make_pairs(entries[k:k + 1], cuda), 1, 1, H, W, opts)
        assert np.array_equal(alone.cpu().numpy()[0], r[k])
    # (3) camera-frame alignment, source depths x 2 (exact in float32): every source point doubles exactly, so s
    #     halves, R and t stay (up to the rounding of the float32 micro-batches, far inside the 1e-6 contract)
    cam = L.default_opts(world=0)
    scaled = DeviceSubmap(dsubs[1].depth * 2.0, dsubs[1].conf, dsubs[1].cams, dsubs[1].intrinsics, dsubs[1].extrinsics, dsubs[1].images)
    r1, _, _ = ops.align_pairs(ops.make_pairs(entries[:1], cuda), 1, 1, H, W, cam)
    r2, _, _ = ops.align_pairs(ops.make_pairs([pair_entry(dsubs[0], scaled, 1)], cuda), 1, 1, H, W, cam)
    r1, r2 = r1.cpu().numpy()[0], r2.cpu().numpy()[0]
    assert abs(2.0 * r2[0] - r1[0]) <= REL * r1[0] and np.abs(r2[1:13] - r1[1:13]).max() <= REL
    assert r2[13] == r1[13] == r[0, 13]                         # the same correspondences were kept
    # (4) exact median threshold of a full frame against a full sort (numpy >= 2 float32 arithmetic, utils/align.py:140-142)
    cA, cB = dsubs[0].conf[-1].flatten(), dsubs[1].conf[0].flatten()
    med = []
    for c in (cA, cB):
        srt = torch.sort(c).values
        n = srt.numel()
        med.append(((srt[n // 2 - 1] + srt[n // 2]) / 2.0) if n % 2 == 0 else srt[n // 2])
    want_thr = (torch.minimum(med[0], med[1]) * torch.tensor(0.1, dtype=torch.float32, device=cuda)).item()
    assert np.float32(aux.cpu().numpy()[0, 0]) == np.float32(want_thr)
    assert r[0, 13] == int(((cA > want_thr) & (cB > want_thr) & (dsubs[0].depth[-1].flatten() > 1e-6)
                            & (dsubs[1].depth[0].flatten() > 1e-6)).sum().item())
    # (5) every kept point lands in exactly one voxel: counts sum to the kept points; fused == two-kernel export
    kw = dict(overlap=1, voxel=0.02, conf_percentile=65.0, table_slots=1 << 22, world=1)
    two = SequencePlan(dsubs, fuse_export=False, **kw)
    two.run()
    out2 = two.read(sort=True)
    kept = sum(int(m_.sum().item()) for m_ in two.mask)
    assert int(out2["voxel_count"].sum().item()) == kept and kept > 0
    fused = SequencePlan(dsubs, fuse_export=True, **kw)
    fused.run()
    outf = fused.read(sort=True)
    for name in ("voxel_key", "voxel_count", "voxel_xyz", "voxel_rgb"):
        assert torch.equal(outf[name], out2[name]), name
    assert np.array_equal(outf["rows"], out2["rows"]) and np.array_equal(out2["rows"], r)


def _reexpress(sub, frames, sim):
    """The same frames predicted in another frame L with p_old = s R p_L + t: depth / s, w2c [R_e R | (t_e + R_e t) / s]."""
    s_, R_, t_ = sim
    E = np.asarray(sub["extrinsics"], np.float64)[frames]
    out = {k: np.ascontiguousarray(np.asarray(sub[k])[frames]) for k in ("depth", "conf", "intrinsics")}
    out["depth"] = (out["depth"] / np.float32(s_)).astype(np.float32)
    Rn = E[:, :, :3] @ R_
    tn = (E[:, :, 3] + E[:, :, :3] @ t_) / s_
    out["extrinsics"] = np.concatenate([Rn, tn[:, :, None]], axis=2).astype(np.float32)
    return out


def test_loop_constraints_and_pose_graph(cuda):
    """Loop closure end to end: a jointly predicted loop chunk holding the last frame of chunk 0 and the
    first frame of chunk 3 gives, through two batched GPU alignments, the constraint 'chunk 3 in chunk 0'; the pose graph
    then pulls a drifting chain back."""
    from da3slam_b200 import pipeline, posegraph
    rng = np.random.default_rng(8)
    n, F, H, W = 4, 2, 60, 80
    subs, gt = synth.make_sequence(n, F, H, W, overlap=1, seed=91)
    A = posegraph.sequential_to_absolute(gt)                                   # chunk k in chunk 0's frame
    S_L = synth.random_sim3(rng)                                               # loop frame: p_0 = S_L p_L
    own_a = _reexpress(subs[0], [F - 1], (1.0, np.eye(3), np.zeros(3)))
    own_b = _reexpress(subs[3], [0], (1.0, np.eye(3), np.zeros(3)))
    loop_a = _reexpress(subs[0], [F - 1], S_L)                                 # p_0 = S_L p_L
    loop_b = _reexpress(subs[3], [0], posegraph._compose(posegraph._inverse(A[3]), posegraph._as_sim3(S_L)))   # p_3 = A_3^-1 S_L p_L
    dev = [DeviceSubmap.from_prediction(x, cuda) for x in (own_a, loop_a, own_b, loop_b)]
    loops = pipeline.loop_constraints([(0, 3, dev[0], dev[1], dev[2], dev[3])], world=1)
    (a, b, T), = loops
    assert (a, b) == (0, 3)
    assert abs(T[0] - A[3][0]) < 2e-3 * A[3][0] and np.abs(T[1] - A[3][1]).max() < 2e-3 and np.abs(T[2] - A[3][2]).max() < 2e-2
    # drifting odometry + the measured loop: the end of the chain comes back
    noisy = [posegraph._compose(posegraph._as_sim3(g), posegraph._unchart(np.concatenate([rng.normal(0, 0.02, 3), [rng.normal(0, 0.02)],
                                                                                               rng.normal(0, 0.05, 3)]))) for g in gt]
    fixed = pipeline.close_loops(noisy, loops)

    def end_err(seq):
        E_ = posegraph.sequential_to_absolute(seq)[-1]
        return np.linalg.norm(posegraph._chart(posegraph._compose(posegraph._inverse(A[3]), E_)))
    assert end_err(fixed) < 0.5 * end_err(noisy)


def test_unproject_jobs_equals_flat_launch(cuda):
    rng = np.random.default_rng(12)
    n, H, W = 5, 30, 44
    depth = synth.smooth_depth(rng, n, H, W)
    conf = synth.da3_like_conf(rng, n, H, W)
    d, c = dev_t(depth, cuda), dev_t(conf, cuda)
    cams = ops.build_cams(dev_t(synth.make_intrinsics(n, H, W), cuda), dev_t(synth.trajectory_w2c(rng, n).astype(np.float32), cuda))
    rows = torch.stack([ops.sim3_row(*synth.random_sim3(rng), cuda) for _ in range(n)])
    thr = torch.tensor([0.3, 0.5], dtype=torch.float32, device=cuda)
    ref_xyz, ref_mask, ref_cnt = [], [], 0
    for f in range(n):
        x, m, k = ops.unproject_filter(d[f:f + 1], c[f:f + 1], cams[f:f + 1], mode="fast", world=True, sim3=rows[f], conf_cmp=">=",
                                       conf_thr_dev=thr[f % 2:f % 2 + 1], conf_floor=0.0, depth_eps=1e-6)
        ref_xyz.append(x); ref_mask.append(m); ref_cnt += int(k.item())
    xyz = torch.empty((n, H, W, 3), dtype=torch.float32, device=cuda)
    mask = torch.empty((n, H, W), dtype=torch.uint8, device=cuda)
    jobs = [dict(depth=d[f], conf=c[f], cam=cams[f], sim3=rows[f], conf_thr=thr[f % 2:f % 2 + 1], xyz=xyz[f], mask=mask[f]) for f in range(n)]
    kept = torch.zeros((1,), dtype=torch.int64, device=cuda)
    ops.unproject_filter_jobs(ops.make_frame_jobs(jobs, cuda), n, H, W, mode="fast", world=True, conf_cmp=">=", conf_floor=0.0,
                              depth_eps=1e-6, n_kept=kept)
    assert torch.equal(xyz, torch.cat(ref_xyz)) and torch.equal(mask.bool(), torch.cat(ref_mask)) and int(kept.item()) == ref_cnt


@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 300001])
def test_filter_points_is_an_ordered_compaction(cuda, n):
    """da3s_filter_points == points[valid & (conf >= thr)] in input order (viewer.py:333-355), bit for bit."""
    rng = np.random.default_rng(n + 3)
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    conf = rng.random(n).astype(np.float32)
    valid = rng.random(n) < 0.7
    d = [dev_t(a, cuda) for a in (xyz, rgb, conf, valid)]
    for thr in (None, 0.35, 2.0):
        want = valid if thr is None else valid & (conf >= np.float32(thr))
        p, c = ops.filter_points(d[0], d[1], d[2], d[3], thr=thr)
        assert np.array_equal(p.cpu().numpy(), xyz[want]) and np.array_equal(c.cpu().numpy(), rgb[want])
    p, c = ops.filter_points(d[0], None, d[2], None, thr=0.5)                   # no validity bytes, no colours
    assert c is None and np.array_equal(p.cpu().numpy(), xyz[conf >= np.float32(0.5)])
