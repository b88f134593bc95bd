// sim3_math.cuh — float64 3x3 algebra, one-sided Jacobi SVD and the closed-form
// Umeyama solve from raw weighted moments.  Runs on ONE device thread per pair (the
// data-parallel work is the moment reduction; this is ~1k flops).
//
// Moment vector layout (MOM_LEN doubles), raw sums over the kept correspondences
// (x = source point, y = target point, w = weight):
//   [0]      S0  = sum w
//   [1..3]   Sx  = sum w x
//   [4..6]   Sy  = sum w y
//   [7..15]  Syx = sum w y x^T   (row-major, Syx[i][j] = sum w y_i x_j)
//   [16..21] Sxx = sum w x x^T   (symmetric: xx, xy, xz, yy, yz, zz).  The full matrix, not
//            just its trace, so that the move to world coordinates is exact for ANY
//            camera-to-world matrix (float32 rotations are orthonormal only to ~6e-8)
//   [22]     Sr  = sum residual (diagnostic, IRLS only)
//   [23]     wmax (max weight; combined with max, not +)
//   [24]     n   (count, exact in float64)
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

#define MOM_S0 0
#define MOM_SX 1
#define MOM_SY 4
#define MOM_SYX 7
#define MOM_SXX 16
#define MOM_SR 22
#define MOM_WMAX 23
#define MOM_N 24
#define MOM_LEN 25
#define EFF_LEN 33          // per (pair, frame): By[9] | Bx[9] | c[3] | Bp[9] | cp[3]
                            //   exact residual  r = By y + Bx x + c            (float64 kernel)
                            //   |r| ~= |y + Bp x + cp|, Bp = By^-1 Bx, cp = By^-1 c   (float32 kernel; equal up to the
                            //   ~6e-8 non-orthonormality of a float32 rotation, i.e. at float32 rounding level)

HD double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
           M[2] * (M[3] * M[7] - M[4] * M[6]);
}
HD void mat3_mul(const double* A, const double* B, double* C) {           // C = A B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
HD void mat3_mul_bt(const double* A, const double* B, double* C) {        // C = A B^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
HD void mat3_tmul(const double* A, const double* B, double* C) {          // C = A^T B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
HD void mat3_vec(const double* A, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
HD void mat3_tvec(const double* A, const double* v, double* o) {          // o = A^T v
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
HD bool mat3_inv(const double* M, double* I) {
    double d = det3(M);
    if (d == 0.0 || !(fabs(d) < INFINITY)) return false;
    double r = 1.0 / d;
    I[0] = (M[4] * M[8] - M[5] * M[7]) * r;  I[1] = (M[2] * M[7] - M[1] * M[8]) * r;  I[2] = (M[1] * M[5] - M[2] * M[4]) * r;
    I[3] = (M[5] * M[6] - M[3] * M[8]) * r;  I[4] = (M[0] * M[8] - M[2] * M[6]) * r;  I[5] = (M[2] * M[3] - M[0] * M[5]) * r;
    I[6] = (M[3] * M[7] - M[4] * M[6]) * r;  I[7] = (M[1] * M[6] - M[0] * M[7]) * r;  I[8] = (M[0] * M[4] - M[1] * M[3]) * r;
    return true;
}

// A = U diag(S) V^T, S descending, U and V orthonormal (columns).  Row-major 3x3.
HD void svd3(const double* A, double* U, double* S, double* V) {
    double u[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) u[i] = A[i];
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < 3; ++i) {
                    a += u[3 * i + p] * u[3 * i + p];
                    b += u[3 * i + q] * u[3 * i + q];
                    g += u[3 * i + p] * u[3 * i + q];
                }
                if (g == 0.0 || fabs(g) <= 2.5e-16 * sqrt(a * b)) continue;
                rotated = true;
                double zeta = (b - a) / (2.0 * g);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    double up = u[3 * i + p], uq = u[3 * i + q];
                    u[3 * i + p] = c * up - s * uq;
                    u[3 * i + q] = s * up + c * uq;
                    double vp = v[3 * i + p], vq = v[3 * i + q];
                    v[3 * i + p] = c * vp - s * vq;
                    v[3 * i + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sg[3];
    for (int j = 0; j < 3; ++j)
        sg[j] = sqrt(u[j] * u[j] + u[3 + j] * u[3 + j] + u[6 + j] * u[6 + j]);
    int o0 = 0, o1 = 1, o2 = 2, tmp;                         // sort descending
    if (sg[o0] < sg[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
    if (sg[o0] < sg[o2]) { tmp = o0; o0 = o2; o2 = tmp; }
    if (sg[o1] < sg[o2]) { tmp = o1; o1 = o2; o2 = tmp; }
    int ord[3] = {o0, o1, o2};
    for (int j = 0; j < 3; ++j) {
        int k = ord[j];
        S[j] = sg[k];
        double inv = sg[k] > 0 ? 1.0 / sg[k] : 0.0;
        for (int i = 0; i < 3; ++i) {
            U[3 * i + j] = u[3 * i + k] * inv;
            V[3 * i + j] = v[3 * i + k];
        }
    }
    // rank-deficient input: rebuild the null-space columns of U so that U stays orthonormal
    double tiny = 1e-13 * S[0];
    if (S[0] <= 0) {
        for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    if (S[1] <= tiny) {                                      // rank 1: any unit vector orthogonal to u0
        double a0 = fabs(U[0]), a1 = fabs(U[3]), a2 = fabs(U[6]);
        double e[3] = {0, 0, 0};
        e[(a0 <= a1 && a0 <= a2) ? 0 : (a1 <= a2 ? 1 : 2)] = 1.0;
        double d = e[0] * U[0] + e[1] * U[3] + e[2] * U[6];
        double w0 = e[0] - d * U[0], w1 = e[1] - d * U[3], w2 = e[2] - d * U[6];
        double n = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
        U[1] = w0 / n; U[4] = w1 / n; U[7] = w2 / n;
    }
    if (S[2] <= tiny) {                                      // rank <= 2: u2 = u0 x u1
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

// Variants of the closed form (which epsilons, which determinant test):
//   0  utils/align.py:14-40   weights normalised by (sum w + 1e-8); var + 1e-8; det(U Vt)
//   1  align_geometry.py:59-82  unweighted means; cov, var divided by N; var + 1e-12; det(U) det(Vt)
//   2  utils/align.py:42-92   as 0, but the scale is trace(Sigma) / (var + 1e-8) — the reference's legacy solver,
//      wrong for rotated data yet part of its API (weighted_umeyama_alignment0); reproduced, not corrected
#define SOLVE_WEIGHTED 0
#define SOLVE_MEAN 1
#define SOLVE_LEGACY_TRACE 2

// mom: raw moments (world or camera frame).  wscale: every weight is divided by this first
// (utils/align.py:194: max(w) + 1e-8; 1.0 when not in IRLS).  Writes s, R[9], t[3].
// Returns false (identity written) when the moments are unusable.
HD bool umeyama_from_moments(const double* mom, double wscale, int variant, double* s_out, double* R, double* t) {
    double S0 = mom[MOM_S0] / wscale;
    double den, eps_var;
    if (variant == SOLVE_MEAN) { den = mom[MOM_N]; eps_var = 1e-12; }
    else                       { den = S0 + 1e-8; eps_var = 1e-8; }
    bool ok = den > 0 && S0 > 0;
    double mx[3], my[3], Sx[3], Sy[3];
    for (int i = 0; i < 3; ++i) {
        Sx[i] = mom[MOM_SX + i] / wscale;
        Sy[i] = mom[MOM_SY + i] / wscale;
        mx[i] = ok ? Sx[i] / den : 0.0;
        my[i] = ok ? Sy[i] / den : 0.0;
    }
    double cov[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            cov[3 * i + j] = (mom[MOM_SYX + 3 * i + j] / wscale - my[i] * Sx[j] - Sy[i] * mx[j] + S0 * my[i] * mx[j]) / den;
    double var = ((mom[MOM_SXX] + mom[MOM_SXX + 3] + mom[MOM_SXX + 5]) / wscale - 2.0 * (mx[0] * Sx[0] + mx[1] * Sx[1] + mx[2] * Sx[2]) +
                  S0 * (mx[0] * mx[0] + mx[1] * mx[1] + mx[2] * mx[2])) / den;
    double U[9], Sg[3], V[9];
    svd3(cov, U, Sg, V);
    double dsign = 1.0;
    if (variant != SOLVE_MEAN) {
        double UVt[9];
        mat3_mul_bt(U, V, UVt);
        if (det3(UVt) < 0) dsign = -1.0;
    } else {
        if (det3(U) * det3(V) < 0) dsign = -1.0;
    }
    double Ud[9];
    for (int i = 0; i < 3; ++i) { Ud[3 * i] = U[3 * i]; Ud[3 * i + 1] = U[3 * i + 1]; Ud[3 * i + 2] = U[3 * i + 2] * dsign; }
    mat3_mul_bt(Ud, V, R);
    double s = (Sg[0] + Sg[1] + dsign * Sg[2]) / (var + eps_var);
    if (variant == SOLVE_LEGACY_TRACE) s = (cov[0] + cov[4] + cov[8]) / (var + eps_var);      // utils/align.py:83-87
    double Rm[3];
    mat3_vec(R, mx, Rm);
    for (int i = 0; i < 3; ++i) t[i] = my[i] - s * Rm[i];
    *s_out = s;
    bool finite = (fabs(s) < INFINITY);
    for (int i = 0; i < 9; ++i) finite = finite && (fabs(R[i]) < INFINITY);
    for (int i = 0; i < 3; ++i) finite = finite && (fabs(t[i]) < INFINITY);
    if (!ok || !finite) {
        *s_out = 1.0;
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        t[0] = t[1] = t[2] = 0.0;
        return false;
    }
    return true;
}

// Raw camera-frame moments of one overlap frame -> world-frame moments under the two frames'
// camera-to-world maps x_w = Mx x + mx, y_w = My y + my, ADDED into `acc`.  Every moment is a
// polynomial in the points, so this is exact (up to float64 rounding) for any affine map.
HD void moments_to_world_add(const double* m, const double* c2w_x, const double* c2w_y, double* acc) {
    double Mx[9], My[9], tx[3], ty[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { Mx[3 * i + j] = c2w_x[4 * i + j]; My[3 * i + j] = c2w_y[4 * i + j]; }
        tx[i] = c2w_x[4 * i + 3];
        ty[i] = c2w_y[4 * i + 3];
    }
    double S0 = m[MOM_S0];
    double MSx[3], MSy[3];
    mat3_vec(Mx, &m[MOM_SX], MSx);
    mat3_vec(My, &m[MOM_SY], MSy);
    acc[MOM_S0] += S0;
    for (int i = 0; i < 3; ++i) {
        acc[MOM_SX + i] += MSx[i] + tx[i] * S0;
        acc[MOM_SY + i] += MSy[i] + ty[i] * S0;
    }
    double T1[9], T2[9];
    mat3_mul(My, &m[MOM_SYX], T1);           // My Syx
    mat3_mul_bt(T1, Mx, T2);                 // My Syx Mx^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            acc[MOM_SYX + 3 * i + j] += T2[3 * i + j] + MSy[i] * tx[j] + ty[i] * MSx[j] + S0 * ty[i] * tx[j];
    // sum w x_w x_w^T = Mx Sxx Mx^T + (Mx Sx) mx^T + mx (Mx Sx)^T + S0 mx mx^T
    const double* q = &m[MOM_SXX];
    double Sm[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
    mat3_mul(Mx, Sm, T1);
    mat3_mul_bt(T1, Mx, T2);
    int idx = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            acc[MOM_SXX + idx++] += T2[3 * i + j] + MSx[i] * tx[j] + tx[i] * MSx[j] + S0 * tx[i] * tx[j];
    acc[MOM_SR] += m[MOM_SR];
    acc[MOM_WMAX] = fmax(acc[MOM_WMAX], m[MOM_WMAX]);
    acc[MOM_N] += m[MOM_N];
}

// Residual of one overlap frame under the world Sim(3) (s,R,t), evaluated from CAMERA-frame
// points:  r_w = y_w - (s R x_w + t) = By y + Bx x + c  with  By = My, Bx = -s R Mx,
// c = my - s R mx - t.  Exact for any camera-to-world matrices.
HD void effective_residual_transform(double s, const double* R, const double* t,
                                     const double* c2w_x, const double* c2w_y, double* eff) {
    double Mx[9], tx[3], sR[9];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { Mx[3 * i + j] = c2w_x[4 * i + j]; eff[3 * i + j] = c2w_y[4 * i + j]; sR[3 * i + j] = s * R[3 * i + j]; }
        tx[i] = c2w_x[4 * i + 3];
    }
    double T1[9], v[3];
    mat3_mul(sR, Mx, T1);
    for (int k = 0; k < 9; ++k) eff[9 + k] = -T1[k];
    mat3_vec(sR, tx, v);
    for (int i = 0; i < 3; ++i) eff[18 + i] = c2w_y[4 * i + 3] - v[i] - t[i];
    double Byi[9];
    if (!mat3_inv(eff, Byi)) { for (int k = 0; k < 9; ++k) Byi[k] = (k % 4 == 0) ? 1.0 : 0.0; }
    mat3_mul(Byi, eff + 9, eff + 21);
    mat3_vec(Byi, eff + 18, eff + 30);
}
