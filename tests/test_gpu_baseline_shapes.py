"""GPU parity AT THE BASELINE SHAPES (518 x 518, 1036 x 1036, 1024 hypotheses), default kernel.

tests/test_gpu_kernels.py compares the kernels with the oracle on small frames; these tests do the same
comparison on the frame sizes BASELINE.json names, through the C ABI (da3s_align_pairs, opts.precise = 0 — the
mixed-precision kernel bench.py times), against oracle/spec_port.align_pair which follows
utils/align.py:14-40 (weighted Umeyama) and :169-211 (IRLS) in float64:

    n_valid, iterations, status, winning hypothesis and its inlier count   exact
    scale, rotation, translation                                           <= 1e-6 relative (north-star tolerance)

One 518 x 518 pair costs the numpy oracle 0.3 - 2 s, so every row of the batches below is checked.
"""
import numpy as np
import pytest
import torch

from da3slam_b200 import _lib as L
from da3slam_b200 import ops, synth
from da3slam_b200.pipeline import DeviceSubmap, pair_entry
from oracle import spec_port as sp

pytestmark = pytest.mark.gpu

REL = 1e-6          # north-star tolerance for scale / rotation / translation


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def dev_pairs(subs, cuda, overlap):
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    entries = [pair_entry(dsubs[k], dsubs[k + 1], overlap) for k in range(len(dsubs) - 1)]
    return dsubs, ops.make_pairs(entries, cuda), len(entries)


def check_row(r, o, tag=""):
    assert int(r[15]) == o["status"], (tag, r[15], o["status"])
    assert int(r[13]) == o["n_valid"], (tag, r[13], o["n_valid"])                 # the joint confidence mask, bit-exact
    if o["status"] != 0:
        assert r[0] == 1.0 and np.array_equal(r[1:10].reshape(3, 3), np.eye(3)) and not r[10:13].any()
        return
    assert int(r[14]) == o["iters"], (tag, r[14], o["iters"])
    assert abs(r[0] - o["s"]) <= REL * abs(o["s"]), (tag, r[0], o["s"])
    assert rel_err(r[1:10].reshape(3, 3), o["R"]) <= REL, tag
    assert np.abs(r[10:13] - o["t"]).max() <= REL * max(1.0, np.abs(o["t"]).max()), tag


def test_align_8_pairs_518_vs_oracle(cuda):
    """configs[0]/[1] shape: 518 x 518, one overlap frame, defaults of utils/align.py (delta 1.0, <= 20 it, tol 1e-6)."""
    H = W = 518
    subs, gt = synth.make_sequence(9, 2, H, W, overlap=1, seed=20518)
    keep, table, n = dev_pairs(subs, cuda, 1)
    for world in (1, 0):
        rows, _, _ = ops.align_pairs(table, n, 1, H, W, L.default_opts(world=world))
        rows = rows.cpu().numpy()
        for k in range(n):
            check_row(rows[k], sp.align_pair(subs[k], subs[k + 1], overlap=1, world=bool(world)), f"pair {k} world={world}")
        if world:
            for k in range(n):                                                     # and the generating Sim(3) is recovered
                assert abs(rows[k, 0] - gt[k][0]) < 1e-3 * gt[k][0]


@pytest.mark.parametrize("delta", [1.0, 0.1])
def test_align_518_overlap2_outliers_huber(cuda, delta):
    """Two overlap frames, 25 % gross outliers: many residuals above delta, so the Huber branch decides the weights."""
    H = W = 518
    subs, _ = synth.make_sequence(4, 3, H, W, overlap=2, seed=31518, outlier_ratio=0.25)
    keep, table, n = dev_pairs(subs, cuda, 2)
    rows, _, _ = ops.align_pairs(table, n, 2, H, W, L.default_opts(world=1, huber_delta=delta))
    rows = rows.cpu().numpy()
    for k in range(n):
        check_row(rows[k], sp.align_pair(subs[k], subs[k + 1], overlap=2, world=True, delta=delta), f"pair {k} delta={delta}")


def test_align_2_pairs_1036_vs_oracle(cuda):
    """configs[4] shape: 1036 x 1036 (1 073 296 correspondences per pair)."""
    H = W = 1036
    subs, _ = synth.make_sequence(3, 2, H, W, overlap=1, seed=41036)
    keep, table, n = dev_pairs(subs, cuda, 1)
    rows, _, _ = ops.align_pairs(table, n, 1, H, W, L.default_opts(world=1))
    rows = rows.cpu().numpy()
    for k in range(n):
        check_row(rows[k], sp.align_pair(subs[k], subs[k + 1], overlap=1, world=True), f"pair {k}")


def test_huber_knee_within_one_ulp(cuda):
    """Residuals of the FIRST pass engineered to sit at delta * (1 +- a few ulp): the float32 branch `rr > delta^2`
    of the default kernel may take the other side of the knee than the float64 oracle does (utils/align.py:186-191),
    which must not matter — the Huber weight is continuous there."""
    H, W, delta = 518, 518, 0.25
    rng = np.random.default_rng(7)
    A, B, _ = synth.make_pair(H, W, frames=2, overlap=1, seed=518)
    # camera mode, identity start: r = |dA - dB| * sqrt(1 + a^2 + b^2), a = (u - cu)/fu, b = (v - cv)/fv
    K = A["intrinsics"][-1].astype(np.float64)
    u = (np.arange(W)[None, :] - K[0, 2]) / K[0, 0]
    v = (np.arange(H)[:, None] - K[1, 2]) / K[1, 1]
    nrm = np.sqrt(1.0 + u * u + v * v)
    dA = A["depth"][-1].astype(np.float64)
    dB = (dA - delta / nrm).astype(np.float32)
    ulps = rng.integers(-2, 3, size=dB.shape)                                       # -2 .. +2 ulp around the knee
    dB = (dB.view(np.int32) + ulps.astype(np.int32)).view(np.float32)
    B["depth"][0] = dB
    B["intrinsics"][0] = A["intrinsics"][-1]
    keep, table, n = dev_pairs([A, B], cuda, 1)
    rows, _, _ = ops.align_pairs(table, n, 1, H, W, L.default_opts(world=0, huber_delta=delta))
    check_row(rows.cpu().numpy()[0], sp.align_pair(A, B, overlap=1, world=False, delta=delta), "knee")


def test_ransac_1024_hypotheses_64_pairs(cuda):
    """configs[2] shape: 64 pairs x 1024 hypotheses at 518 x 518, 30 % outliers, through da3s_align_pairs.
    Every pair: ground truth recovered and the winner's count equals its own score table.  Pairs 0, 21, 42, 63:
    winner, inlier count, kept correspondences and the refined Sim(3) against the oracle run on the same indices."""
    H = W = 518
    n_sub, n_hyp, thr = 65, 1024, 0.02
    subs, gt = synth.make_sequence_device(n_sub, 2, H, W, overlap=1, seed=2000, outlier_ratio=0.3, with_images=False, device=cuda)
    dsubs = [DeviceSubmap.from_prediction(s_, cuda) for s_ in subs]
    entries = [pair_entry(dsubs[k], dsubs[k + 1], 1) for k in range(n_sub - 1)]
    n = len(entries)
    rng = np.random.default_rng(99)
    si = rng.integers(0, H * W, size=(n, n_hyp, 3)).astype(np.int32)
    opts = L.default_opts(world=1, n_hyp=n_hyp, ransac_thr=thr)
    rows, aux, counts = ops.align_pairs(ops.make_pairs(entries, cuda), n, 1, H, W, opts, torch.from_numpy(si).to(cuda),
                                        want_aux=True, want_counts=True)
    # the same call without the count table scores in rounds and drops hypotheses that can no longer win (pair_align.cu,
    # rs_round_of): winner, its count and every row must come out bit for bit the same
    rows_r, aux_r, _ = ops.align_pairs(ops.make_pairs(entries, cuda), n, 1, H, W, opts, torch.from_numpy(si).to(cuda), want_aux=True)
    assert torch.equal(rows_r, rows) and torch.equal(aux_r, aux)
    rows, aux, counts = rows.cpu().numpy(), aux.cpu().numpy(), counts.cpu().numpy()
    for k in range(n):
        assert rows[k, 15] == 0 and abs(rows[k, 0] - gt[k][0]) < 5e-3 * gt[k][0], k
        best = int(aux[k, 4])
        assert best == int(np.argmax(counts[k])) and int(aux[k, 5]) == int(counts[k].max())    # ties -> lowest index
    for k in (0, 21, 42, 63):
        prev, cur = synth.submap_to_host(subs[k]), synth.submap_to_host(subs[k + 1])
        corr = sp.pair_correspondences(prev, cur, 1, True)
        xs, ys = sp.ransac_points(corr, True)
        A, T, ok, _ = sp.ransac_hypotheses(xs, ys, corr["mask"], si[k])
        ref_counts = sp.ransac_score_c(A, T, ok, xs, ys, corr["mask"], thr)
        best, nbest = sp.ransac_best(ref_counts, ok)
        assert int(aux[k, 4]) == best and int(aux[k, 5]) == nbest, k
        # hypothesis tables agree to 1e-9 before their float32 rounding (device Jacobi SVD vs LAPACK): a count may move by
        # the few correspondences that sit within that distance of the threshold, never more
        assert np.abs(counts[k].astype(np.int64) - ref_counts).max() <= 2, k
        mask = sp.ransac_inlier_mask_c(A, T, best, xs, ys, corr["mask"], thr)
        s, R, t, info = sp.irls_dense(corr["x"], corr["y"], corr["c"], mask)
        check_row(rows[k], dict(s=s, R=R, t=t, iters=info["iters"], n_valid=info["n_valid"], status=info["status"]), f"pair {k}")


@pytest.mark.parametrize("outlier", [0.1, 0.45])
def test_ransac_rounds_two_overlap_frames(cuda, outlier):
    """The round scoring (pair_align.cu: rs_round_of, leader pass, pruning) on a second geometry: 384 x 384 (9 tiles per
    frame), TWO overlap frames, 10 % outliers (almost every valid hypothesis survives) and 45 % (the leader's ratio is
    below what round 0 can prune on: round 1 is the first to drop anything).  With and without the count table the rows,
    the winner and its count must be identical; the winner of every pair is checked against the oracle's table."""
    H = W = 384
    n_sub, n_hyp, thr, ov = 4, 192, 0.02, 2
    subs, gt = synth.make_sequence_device(n_sub, 3, H, W, overlap=ov, seed=31, outlier_ratio=outlier, with_images=False, device=cuda)
    dsubs, table, n = dev_pairs(subs, cuda, ov)
    rng = np.random.default_rng(17)
    si = rng.integers(0, ov * H * W, size=(n, n_hyp, 3)).astype(np.int32)
    opts = L.default_opts(world=1, n_hyp=n_hyp, ransac_thr=thr)
    rows_f, aux_f, counts = ops.align_pairs(table, n, ov, H, W, opts, torch.from_numpy(si).to(cuda), want_aux=True, want_counts=True)
    rows_r, aux_r, _ = ops.align_pairs(table, n, ov, H, W, opts, torch.from_numpy(si).to(cuda), want_aux=True)
    assert torch.equal(rows_r, rows_f) and torch.equal(aux_r, aux_f)
    aux, counts = aux_r.cpu().numpy(), counts.cpu().numpy()
    for k in range(n):
        prev, cur = synth.submap_to_host(subs[k]), synth.submap_to_host(subs[k + 1])
        corr = sp.pair_correspondences(prev, cur, ov, True)
        xs, ys = sp.ransac_points(corr, True)
        A, T, ok, _ = sp.ransac_hypotheses(xs, ys, corr["mask"], si[k])
        ref_counts = sp.ransac_score_c(A, T, ok, xs, ys, corr["mask"], thr)
        best, nbest = sp.ransac_best(ref_counts, ok)
        assert int(aux[k, 4]) == best and abs(int(aux[k, 5]) - nbest) <= 2, (k, aux[k, 4:6], best, nbest)
        assert int(aux[k, 5]) == int(counts[k].max())
