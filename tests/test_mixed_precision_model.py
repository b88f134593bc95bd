"""Error model of the default (mixed-precision) IRLS kernel, on the CPU.

`pair_moments_mixed_kernel` (da3slam_b200/csrc/pair_align.cu) accumulates the 22 weighted moments of a pair in float32
FMA arithmetic, in micro-batches of 16 correspondences per thread that are flushed into float64 accumulators; tiles are
combined, un-pivoted and moved to world coordinates in float64.  The reference is float64 throughout
(utils/align.py:14-40, :169-211).  This test restates that arithmetic in numpy in both forms the kernel has —
  scalar: points by SPEC 1, centred at the frame-centre correspondence (all three coordinates);
  packed (default for even W): points from the pixel's ray, x = d (a, b, 1) with a = (u - cu) * (1/fu) rounded once,
          only the depth coordinate pivoted, the even and the odd pixel of a pair in separate float32 accumulators
          (the two halves of a packed register) that are added just before the float64 flush —
with the same micro-batch membership (thread t of 128 takes the float4 groups t, t + 128, ... of a 16384-pixel tile,
four groups per flush), the same operation order inside a micro-batch and the float32 Huber branch on rr > delta^2; it runs
the whole IRLS loop and bounds its distance to the float64 oracle: <= 1e-7 relative on (s, R, t) where the contract is
1e-6.  The GPU tests (tests/test_gpu_baseline_shapes.py) check the real kernel against the same oracle at 518^2 and 1036^2."""
import numpy as np
import pytest

from da3slam_b200 import synth
from oracle import ref_port as rp
from oracle import spec_port as sp

F32 = np.float32
THREADS, GROUPS_PER_TILE, GROUPS_PER_FLUSH = 128, 4096, 4


def fma32(acc, a, b):
    """acc + a * b with one float32 rounding (exact product in float64, one rounding; double rounding ~2^-29 of cases)."""
    return (acc.astype(np.float64) + a.astype(np.float64) * b.astype(np.float64)).astype(F32)


def microbatch_order(n_pix):
    """[n_mb, 16] pixel indices (or -1): the correspondences each (tile, thread, flush) accumulates, in kernel order."""
    n_groups = n_pix // 4
    rows = []
    for t0 in range(0, n_groups, GROUPS_PER_TILE):
        g = np.arange(t0, min(n_groups, t0 + GROUPS_PER_TILE))
        pad = (-len(g)) % (THREADS * GROUPS_PER_FLUSH)
        g = np.concatenate([g, np.full(pad, -1)])
        it = g.reshape(-1, THREADS)                                   # iteration j, thread t -> group t0 + j * 128 + t
        it = it.reshape(-1, GROUPS_PER_FLUSH, THREADS).transpose(0, 2, 1)       # [flush, thread, 4 groups]
        rows.append(it.reshape(-1, GROUPS_PER_FLUSH))
    grp = np.concatenate(rows)                                        # [n_mb, 4] group ids
    pix = np.where(grp[:, :, None] >= 0, grp[:, :, None] * 4 + np.arange(4)[None, None, :], -1)
    return pix.reshape(len(grp), 16)


def mixed_moments(xf, yf, w32, pivot_x, pivot_y, order, packed=False):
    """Raw float64 camera-frame moments (25) the kernel would produce for one frame."""
    pad_x = np.vstack([xf, np.zeros((1, 3), F32)])
    pad_y = np.vstack([yf, np.zeros((1, 3), F32)])
    pad_w = np.concatenate([w32, np.zeros(1, F32)])
    if packed:                                                        # even / odd pixel of a pair: separate accumulators
        halves = [mixed_moments_f32(pad_x, pad_y, pad_w, pivot_x, pivot_y, order[:, par::2]) for par in (0, 1)]
        acc = (halves[0] + halves[1]).astype(F32)
    else:
        acc = mixed_moments_f32(pad_x, pad_y, pad_w, pivot_x, pivot_y, order)
    return unpivot(acc.astype(np.float64).sum(0), pivot_x, pivot_y)


def mixed_moments_f32(pad_x, pad_y, pad_w, pivot_x, pivot_y, order):
    acc = np.zeros((order.shape[0], 22), F32)
    for i in range(order.shape[1]):
        idx = order[:, i]                                             # -1 -> the zero-weight pad entry
        wi = pad_w[idx]
        xc = (pad_x[idx] - pivot_x).astype(F32)
        yc = (pad_y[idx] - pivot_y).astype(F32)
        wx = (wi[:, None] * xc).astype(F32)
        acc[:, 0] = (acc[:, 0] + wi).astype(F32)
        for k in range(3):
            acc[:, 1 + k] = (acc[:, 1 + k] + wx[:, k]).astype(F32)
            acc[:, 4 + k] = fma32(acc[:, 4 + k], wi, yc[:, k])
        c = 7
        for a in range(3):
            for b in range(3):
                acc[:, c] = fma32(acc[:, c], yc[:, a], wx[:, b]); c += 1
        for a in range(3):
            for b in range(a, 3):
                acc[:, c] = fma32(acc[:, c], wx[:, a], xc[:, b]); c += 1
    return acc


def unpivot(m, pivot_x, pivot_y):
    """float64 flush + tile combination happened in `m`; undo the pivot (exact polynomial identities in float64)."""
    S0, Sx, Sy, Syx, q = m[0], m[1:4], m[4:7], m[7:16].reshape(3, 3), m[16:22]
    Sxx = np.array([[q[0], q[1], q[2]], [q[1], q[3], q[4]], [q[2], q[4], q[5]]])
    px, py = pivot_x.astype(np.float64), pivot_y.astype(np.float64)   # un-pivot: exact polynomial identities in float64
    Sx_o, Sy_o = Sx + S0 * px, Sy + S0 * py
    Syx_o = Syx + np.outer(Sy, px) + np.outer(py, Sx) + S0 * np.outer(py, px)
    Sxx_o = Sxx + np.outer(Sx, px) + np.outer(px, Sx) + S0 * np.outer(px, px)
    out = np.zeros(25)
    out[0], out[1:4], out[4:7], out[7:16] = S0, Sx_o, Sy_o, Syx_o.reshape(-1)
    out[16:22] = [Sxx_o[0, 0], Sxx_o[0, 1], Sxx_o[0, 2], Sxx_o[1, 1], Sxx_o[1, 2], Sxx_o[2, 2]]
    return out


def to_world(m, EB, EA):
    Rx, tx = sp.c2w_closed_form(EB); Ry, ty = sp.c2w_closed_form(EA)
    Mx, mx, My, my = Rx[0], tx[0], Ry[0], ty[0]
    S0, Sx, Sy, Syx, q = m[0], m[1:4], m[4:7], m[7:16].reshape(3, 3), m[16:22]
    Sxx = np.array([[q[0], q[1], q[2]], [q[1], q[3], q[4]], [q[2], q[4], q[5]]])
    o = np.zeros(25)
    MSx, MSy = Mx @ Sx, My @ Sy
    o[0], o[1:4], o[4:7] = S0, MSx + mx * S0, MSy + my * S0
    o[7:16] = (My @ Syx @ Mx.T + np.outer(MSy, mx) + np.outer(my, MSx) + S0 * np.outer(my, mx)).reshape(-1)
    X = Mx @ Sxx @ Mx.T + np.outer(MSx, mx) + np.outer(mx, MSx) + S0 * np.outer(mx, mx)
    o[16:22] = [X[0, 0], X[0, 1], X[0, 2], X[1, 1], X[1, 2], X[2, 2]]
    return o


def solve(m, wscale):
    S0 = m[0] / wscale; den = S0 + 1e-8
    Sx, Sy = m[1:4] / wscale, m[4:7] / wscale
    mx, my = Sx / den, Sy / den
    cov = (m[7:16].reshape(3, 3) / wscale - np.outer(my, Sx) - np.outer(Sy, mx) + S0 * np.outer(my, mx)) / den
    var = ((m[16] + m[19] + m[21]) / wscale - 2 * mx @ Sx + S0 * mx @ mx) / den
    U, S, Vt = np.linalg.svd(cov)
    D = np.eye(3)
    if np.linalg.det(U @ Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    s = (S @ np.diag(D)) / (var + 1e-8)
    return s, R, my - s * R @ mx


def ray_points(sub, frame, H, W):
    """packed form: x = d * (a, b, 1), a = (u - cu) * (1/fu), b = (v - cv) * (1/fv), every operation rounded once."""
    K = np.asarray(sub["intrinsics"][frame], F32)
    d = np.asarray(sub["depth"][frame], F32)
    a = ((np.arange(W, dtype=F32) - K[0, 2]) * (F32(1) / K[0, 0])).astype(F32)[None, :]
    b = ((np.arange(H, dtype=F32) - K[1, 2]) * (F32(1) / K[1, 1])).astype(F32)[:, None]
    return np.stack([(d * a).astype(F32), (d * b).astype(F32), d], axis=-1).reshape(-1, 3)


def mixed_irls(corr, H, W, delta=1.0, max_it=20, tol=1e-6, packed_points=None):
    xf, yf, c, mask = corr["xf"], corr["yf"], corr["c"], corr["mask"]
    packed = packed_points is not None
    if packed:
        xf, yf = packed_points
    order = microbatch_order(H * W)
    centre = (H // 2) * W + W // 2                                      # the pivot: the correspondence at the frame's centre pixel
    px = np.where(np.isfinite(xf[centre]), xf[centre], 0).astype(F32)
    py = np.where(np.isfinite(yf[centre]), yf[centre], 0).astype(F32)
    if packed:
        px[:2] = 0; py[:2] = 0                                          # depth coordinate only
    Rx, tx = sp.c2w_closed_form(corr["EB"]); Ry, ty = sp.c2w_closed_form(corr["EA"])
    xs = np.where(mask[:, None], xf, F32(0)); ys = np.where(mask[:, None], yf, F32(0))
    s, R, t = 1.0, np.eye(3), np.zeros(3)
    for it in range(max_it):
        By, Bx, cc = Ry[0], -(s * R) @ Rx[0], ty[0] - (s * R) @ tx[0] - t
        Bp, cp = (np.linalg.inv(By) @ Bx).astype(F32), (np.linalg.inv(By) @ cc).astype(F32)      # |r| = |y + Bp x + cp|
        r = np.zeros((len(xs), 3), F32)
        for i in range(3):
            acc = (ys[:, i] + cp[i]).astype(F32)
            for k in (2, 1, 0):
                acc = fma32(acc, np.full(len(xs), Bp[i, k], F32), xs[:, k])
            r[:, i] = acc
        rr = fma32(fma32((r[:, 2] * r[:, 2]).astype(F32), r[:, 1], r[:, 1]), r[:, 0], r[:, 0])
        w = np.where(mask, c, F32(0)).astype(F32)
        big = rr > F32(delta * delta)                                  # the float32 branch of the kernel
        hub = (F32(delta) / np.sqrt(np.maximum(rr, F32(1e-30)))).astype(F32)
        w = np.where(big, (w * hub).astype(F32), w)
        m = to_world(mixed_moments(xs, ys, w, px, py, order, packed), corr["EB"], corr["EA"])
        s_n, R_n, t_n = solve(m, float(w.max()) + 1e-8)
        change = abs(s_n - s) + np.linalg.norm(R_n - R) + np.linalg.norm(t_n - t)
        s, R, t = s_n, R_n, t_n
        if change < tol:
            break
    return s, R, t, it + 1


@pytest.mark.parametrize("H,W,seed,outliers", [(200, 260, 102, 0.0), (518, 518, 100, 0.0), (518, 518, 101, 0.1)])
def test_mixed_precision_irls_stays_far_inside_the_contract(H, W, seed, outliers):
    A, B, _ = synth.make_pair(H, W, frames=2, seed=seed, outlier_ratio=outliers)
    corr = sp.pair_correspondences(A, B, 1, True)
    s0, R0, t0, info = sp.irls_dense(corr["x"], corr["y"], corr["c"], corr["mask"])
    for points in (None, (ray_points(B, 0, H, W), ray_points(A, -1, H, W))):     # scalar form, packed form
        s, R, t, iters = mixed_irls(corr, H, W, packed_points=points)
        assert iters == info["iters"]
        assert abs(s - s0) <= 1e-7 * s0 and np.abs(R - R0).max() <= 1e-7 and np.abs(t - t0).max() <= 1e-7 * max(1.0, np.abs(t0).max())


def test_microbatch_membership():
    order = microbatch_order(518 * 518)
    used = order[order >= 0]
    assert len(used) == 518 * 518 and len(np.unique(used)) == 518 * 518          # every pixel in exactly one micro-batch
    assert np.array_equal(order[0], np.concatenate([np.arange(4) + 4 * 128 * j for j in range(4)]))   # thread 0: groups 0, 128, 256, 384
