"""Sim(3) pose graph with loop constraints (parity unpinned — upstream's optimiser is not
vendored): exact data is a fixed point, loops pull a drifting chain back, and the minimum agrees with
scipy.optimize.least_squares run on the same residuals."""
import numpy as np
from scipy.optimize import least_squares

from da3slam_b200 import posegraph as pg
from da3slam_b200 import synth


def _chain(rng, n):
    return [synth.random_sim3(rng) for _ in range(n - 1)]


def _perturb(rng, T, rot=0.01, scale=0.01, trans=0.02):
    d = np.concatenate([rng.normal(0, rot, 3), [rng.normal(0, scale)], rng.normal(0, trans, 3)])
    return pg._compose(pg._as_sim3(T), pg._unchart(d))


def test_exact_constraints_are_a_fixed_point():
    rng = np.random.default_rng(0)
    seq = _chain(rng, 8)
    A = pg.sequential_to_absolute(seq)
    loops = [(0, 7, pg._compose(pg._inverse(A[0]), A[7])), (2, 6, pg._compose(pg._inverse(A[2]), A[6]))]
    out, info = pg.optimize(seq, loops, return_info=True)
    assert info["cost_before"] < 1e-20 and info["cost_after"] <= info["cost_before"] + 1e-20
    for a, b in zip(out, seq):
        assert abs(a[0] - b[0]) < 1e-9 and np.abs(a[1] - b[1]).max() < 1e-9 and np.abs(a[2] - b[2]).max() < 1e-9


def test_loops_remove_drift_and_match_scipy():
    rng = np.random.default_rng(1)
    n = 12
    truth = _chain(rng, n)
    A_true = pg.sequential_to_absolute(truth)
    noisy = [_perturb(rng, T) for T in truth]                                   # odometry drifts
    loops = [(a, b, pg._compose(pg._inverse(A_true[a]), A_true[b])) for a, b in ((0, 11), (1, 9), (3, 10))]   # loops are exact
    out, info = pg.optimize(noisy, loops, max_iterations=30, lambda_init=1e-6, return_info=True)
    assert info["cost_after"] < 0.5 * info["cost_before"]

    def end_error(seq):
        A = pg.sequential_to_absolute(seq)
        return np.linalg.norm(pg._chart(pg._compose(pg._inverse(A_true[-1]), A[-1])))
    assert end_error(out) < 0.2 * end_error(noisy)

    # independent minimiser on the same residual function
    edges = pg._edges([pg._as_sim3(x) for x in noisy], loops)
    A0 = pg.sequential_to_absolute(noisy)

    def fun(x):
        A = [A0[0]] + [pg._compose(A0[k], pg._unchart(x[7 * (k - 1):7 * k])) for k in range(1, n)]
        return pg.residuals(A, edges)
    ref = least_squares(fun, np.zeros(7 * (n - 1)), method="lm", xtol=1e-14, ftol=1e-14)
    assert abs(info["cost_after"] - 2 * ref.cost) <= 1e-6 * max(1e-12, 2 * ref.cost) + 1e-12


def test_bad_loop_index_is_loud():
    rng = np.random.default_rng(2)
    seq = _chain(rng, 4)
    try:
        pg.optimize(seq, [(0, 9, seq[0])])
    except ValueError:
        return
    raise AssertionError("out-of-range loop accepted")
