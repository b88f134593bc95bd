"""Loop-closure Sim(3) pose graph over chunk transforms — the consumer of the loop-candidate
alignments (`bench.py --workload loop512` produces them on the GPU).

Upstream calls the un-vendored `loop_utils.sim3loop.Sim3LoopOptimizer.optimize(sim3_list, loop_sim3_list)`
(utils/da3_streaming.py:35,183,617 — commented out in the reference's own run, so nothing pins it: PARITY
UNPINNED, the statement below is this project's).  Configuration keys kept: `max_iterations` 30,
`lambda_init` 1e-6 (configs/config1.yaml:23-26).

Conventions (the reference's): `sequential[k]` = (s, R, t) maps chunk k+1 into chunk k (cur -> prev,
align_geometry.py:91); a loop `(a, b, (s, R, t))` maps chunk b into chunk a.  With absolute transforms
A_0 = I, A_{k+1} = A_k o sequential[k] (utils/geometry.py:73-119), the optimiser minimises

    sum_k |chart(sequential[k]^-1 o A_k^-1 o A_{k+1})|^2 + sum_loops |chart(T_ab^-1 o A_a^-1 o A_b)|^2

over A_1..A_{n-1} by Levenberg-Marquardt (damping lambda * diag(J^T J), lambda /10 on success, x10 on
failure), with the local chart  (s, R, t) -> (rotation vector, log s, t)  and the update A <- A o chart^-1(d).
Returns the re-derived sequential list, so it drops into `accumulate_sim3_transforms` unchanged.
Small sparse problem (7 unknowns per chunk): float64 on the host, scipy.sparse normal equations.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import spsolve
from scipy.spatial.transform import Rotation


def _compose(a, b):
    """a o b: apply b first, then a."""
    sa, Ra, ta = a
    sb, Rb, tb = b
    return sa * sb, Ra @ Rb, sa * (Ra @ tb) + ta


def _inverse(a):
    s, R, t = a
    return 1.0 / s, R.T, -(R.T @ t) / s


def _chart(a):
    s, R, t = a
    return np.concatenate([Rotation.from_matrix(R).as_rotvec(), [np.log(s)], t])


def _unchart(d):
    return float(np.exp(d[3])), Rotation.from_rotvec(d[:3]).as_matrix(), np.asarray(d[4:7], np.float64)


def _as_sim3(x):
    s, R, t = x
    return float(s), np.asarray(R, np.float64).reshape(3, 3), np.asarray(t, np.float64).reshape(3)


def sequential_to_absolute(sequential):
    """[(s,R,t)] * (n-1) -> n absolute transforms, identity first (utils/geometry.py:73-119)."""
    out = [(1.0, np.eye(3), np.zeros(3))]
    for rel in sequential:
        out.append(_compose(out[-1], _as_sim3(rel)))
    return out


def absolute_to_sequential(absolute):
    return [_compose(_inverse(absolute[k]), absolute[k + 1]) for k in range(len(absolute) - 1)]


def _edges(sequential, loops):
    e = [(k, k + 1, _as_sim3(rel)) for k, rel in enumerate(sequential)]
    e += [(int(a), int(b), _as_sim3(T)) for a, b, T in loops]
    return e


def residuals(absolute, edges):
    r = np.empty(7 * len(edges))
    for i, (a, b, T) in enumerate(edges):
        r[7 * i:7 * i + 7] = _chart(_compose(_inverse(T), _compose(_inverse(absolute[a]), absolute[b])))
    return r


def optimize(sequential, loops, max_iterations=30, lambda_init=1e-6, tol=1e-12, return_info=False):
    """Sim(3) pose-graph optimisation of a chunk chain with loop constraints.  `sequential`: n-1 relative
    transforms; `loops`: [(a, b, (s, R, t))].  Returns the optimised sequential list (and an info dict)."""
    sequential = [_as_sim3(x) for x in sequential]
    edges = _edges(sequential, loops)
    A = sequential_to_absolute(sequential)
    n = len(A)
    for a, b, _ in edges:
        if not (0 <= a < n and 0 <= b < n and a != b):
            raise ValueError(f"loop ({a}, {b}) does not index two different chunks of {n}")
    lam = float(lambda_init)
    r = residuals(A, edges)
    cost = float(r @ r)
    cost0, it = cost, 0
    eps = 1e-6
    for it in range(1, max_iterations + 1):
        rows, cols, vals = [], [], []
        for i, (a, b, T) in enumerate(edges):                      # central differences in the local chart: 14 columns per edge
            Ti = _inverse(T)
            for node in (a, b):
                if node == 0:
                    continue                                        # gauge: chunk 0 stays the identity
                for k in range(7):
                    d = np.zeros(7)
                    d[k] = eps
                    plus = _compose(A[node], _unchart(d))
                    minus = _compose(A[node], _unchart(-d))
                    Aa_p, Ab_p = (plus, A[b]) if node == a else (A[a], plus)
                    Aa_m, Ab_m = (minus, A[b]) if node == a else (A[a], minus)
                    col = (_chart(_compose(Ti, _compose(_inverse(Aa_p), Ab_p))) -
                           _chart(_compose(Ti, _compose(_inverse(Aa_m), Ab_m)))) / (2 * eps)
                    rows += list(range(7 * i, 7 * i + 7))
                    cols += [7 * (node - 1) + k] * 7
                    vals += list(col)
        J = sparse.csr_matrix((vals, (rows, cols)), shape=(7 * len(edges), 7 * (n - 1)))
        H = (J.T @ J).tocsc()
        g = J.T @ r
        improved = False
        for _ in range(10):
            step = spsolve(H + lam * sparse.diags(H.diagonal() + 1e-12), -g)
            trial = [A[0]] + [_compose(A[k], _unchart(step[7 * (k - 1):7 * k])) for k in range(1, n)]
            r_t = residuals(trial, edges)
            c_t = float(r_t @ r_t)
            if c_t < cost:
                A, r, improved = trial, r_t, True
                done = cost - c_t < tol * max(1.0, cost)
                cost = c_t
                lam = max(lam / 10.0, 1e-15)
                break
            lam *= 10.0
        if not improved or done:
            break
    out = absolute_to_sequential(A)
    if return_info:
        return out, {"iterations": it, "cost_before": cost0, "cost_after": cost, "lambda": lam}
    return out
