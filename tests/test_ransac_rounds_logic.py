"""CPU tests of the round scoring da3s_align_pairs uses for RANSAC (csrc/pair_align.cu: rs_round_of, ransac_lead_kernel,
ransac_leader_kernel, ransac_prune_kernel; include/da3s.h: da3s_ransac_round_of).

The dealing rule is the library's own host function; the pruning is restated in numpy on per-tile inlier counts and
checked against plain full scoring: same winner (most inliers, ties to the lowest index), same winning count, and no
dropped hypothesis could have reached the leader's count.  The CUDA path itself is compared bit for bit with full scoring
in tests/test_gpu_baseline_shapes.py."""
import numpy as np
import pytest

from da3slam_b200 import _lib as L
from da3slam_b200 import synth
from oracle import spec_port as sp

ROUNDS = 3          # DA3S_RANSAC_ROUNDS
TILE = 16384        # DA3S_RANSAC_TILE


def round_of(tile, tpf):
    return L.load().da3s_ransac_round_of(tile, tpf)


def test_dealing_rule_partitions_every_frame():
    lib = L.load()
    assert lib.da3s_ransac_round_of(-1, 4) < 0 and lib.da3s_ransac_round_of(4, 4) < 0 and lib.da3s_ransac_round_of(0, 0) < 0
    for tpf in list(range(1, 80)) + [128, 263, 600]:
        rounds = np.array([round_of(t, tpf) for t in range(tpf)])
        assert rounds.min() >= 0 and rounds.max() < ROUNDS
        assert rounds[tpf // 2] == 0                                  # round 0 is never empty: a leader always exists
        if tpf >= 16:
            share = np.bincount(rounds, minlength=ROUNDS) / tpf
            assert 0.28 <= share[0] <= 0.45 and 0.15 <= share[1] <= 0.30 and share[2] >= 0.35, (tpf, share)
            # interleaved over the image: no run of 4 consecutive tiles without a round-0 tile
            gaps = np.diff(np.flatnonzero(rounds == 0))
            assert gaps.max() <= 4, (tpf, gaps.max())
    assert [round_of(t, 17) for t in range(17)].count(0) == 6         # 518 x 518: 6 + 4 + 7 tiles


def simulate_rounds(C, kept, rounds):
    """C [tiles, n_hyp] inliers per tile, kept [tiles] correspondences per tile, rounds [tiles].  Returns (winner, its count,
    evaluations executed, survivors per round) following ransac_score_rounds."""
    n_hyp = C.shape[1]
    counts = np.zeros(n_hyp, np.int64)
    alive = np.ones(n_hyp, bool)
    r0 = rounds == 0
    counts += C[r0].sum(0)
    work = int(kept[r0].sum()) * n_hyp
    lead = int(np.argmax(counts))                                     # ties -> lowest index (np.argmax)
    counts[lead] += C[~r0, lead].sum()                                # the leader on every other tile
    alive[lead] = False
    survivors = []
    for r in range(1, ROUNDS):
        remaining = int(kept[rounds >= r].sum())
        alive &= (counts + remaining >= counts[lead])
        survivors.append(int(alive.sum()))
        sel = rounds == r
        counts[alive] += C[sel][:, alive].sum(0)
        work += int(kept[sel].sum()) * int(alive.sum())
    complete = alive.copy()
    complete[lead] = True
    best = int(np.argmax(np.where(complete, counts, -1)))
    # what ransac_best_kernel does: it looks at EVERY valid hypothesis' count, complete or partial
    assert best == int(np.argmax(counts))
    return best, int(counts[best]), work, survivors, complete, counts


@pytest.mark.parametrize("seed", range(40))
def test_pruning_keeps_the_winner_on_random_tables(seed):
    rng = np.random.default_rng(seed)
    tpf = int(rng.integers(8, 40))
    n_hyp = int(rng.integers(2, 96))
    rounds = np.array([round_of(t, tpf) for t in range(tpf)])
    kept = rng.integers(0, 400, size=tpf)
    # a mixture: a few good hypotheses (ratio near `top`), many bad ones, deliberate duplicates so that ties happen
    top = rng.uniform(0.2, 0.95)
    ratio = np.where(rng.random(n_hyp) < 0.3, top * rng.uniform(0.9, 1.0, n_hyp), rng.uniform(0.0, 0.3, n_hyp))
    C = rng.binomial(kept[:, None], ratio[None, :])
    if n_hyp > 3:
        C[:, n_hyp - 1] = C[:, int(np.argmax(ratio))]                 # an exact tie with the best one, at a higher index
    full = C.sum(0)
    best, nbest, work, survivors, complete, counts = simulate_rounds(C, kept, rounds)
    assert best == int(np.argmax(full)) and nbest == int(full.max())
    assert np.array_equal(counts[complete], full[complete])           # survivors are completely counted
    assert (full[~complete] < nbest).all()                            # whatever was dropped could not even tie
    assert work <= int(kept.sum()) * n_hyp


def test_pruning_on_oracle_counts_384():
    """Per-tile counts of the CPU oracle on a 384 x 384 pair with 30 % outliers (9 tiles, 96 hypotheses): the rounds find the
    oracle's winner and skip more than a third of the evaluations."""
    H = W = 384
    subs, _ = synth.make_sequence(2, 2, H, W, 1, seed=77, outlier_ratio=0.3)
    corr = sp.pair_correspondences(subs[0], subs[1], 1, True)
    xs, ys = sp.ransac_points(corr, True)
    rng = np.random.default_rng(3)
    si = rng.integers(0, H * W, size=(96, 3))
    A, T, ok, _ = sp.ransac_hypotheses(xs, ys, corr["mask"], si)
    tpf = (H * W + TILE - 1) // TILE
    C = np.zeros((tpf, 96), np.int64)
    kept = np.zeros(tpf, np.int64)
    for t in range(tpf):
        m = np.zeros(H * W, bool)
        m[t * TILE:(t + 1) * TILE] = True
        m &= corr["mask"].reshape(-1)
        kept[t] = int(m.sum())
        C[t] = sp.ransac_score_c(A, T, ok, xs, ys, m, 0.02)
    full = sp.ransac_score_c(A, T, ok, xs, ys, corr["mask"], 0.02)
    assert np.array_equal(C.sum(0), full)
    valid = ok.astype(bool)
    rounds = np.array([round_of(t, tpf) for t in range(tpf)])
    best, nbest, work, survivors, complete, _ = simulate_rounds(C[:, valid], kept, rounds)
    ref_best, ref_n = sp.ransac_best(full, ok)
    assert int(np.flatnonzero(valid)[best]) == ref_best and nbest == ref_n
    assert work < 0.67 * int(kept.sum()) * int(valid.sum()), (work, survivors)
