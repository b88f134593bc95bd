/* oracle.c — plain-C restatement of the two loops numpy cannot state exactly or quickly.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): used by tests/ and by bench.py's CPU
 * baseline legs only.  PARITY UNPINNED: the reference has no RANSAC and no voxel grid;
 * the normative text is oracle/SPEC.md sections 4 and 5.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/build.py).
 * -ffp-contract=off keeps every * and + a separate rounding; the ONLY fused operations
 * are the explicit fmaf() calls, which is exactly what the CUDA kernel's fmaf() does.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* SPEC 4: squared residual of one hypothesis at one correspondence, float32 FMA
 *   p_i = fma(A_i0, x0, fma(A_i1, x1, fma(A_i2, x2, t_i)));  d_i = p_i - y_i
 *   r2  = fma(d0, d0, fma(d1, d1, d2 * d2))                                        */
static inline float residual2(const float* A, const float* t, const float* x, const float* y) {
    float d0 = fmaf(A[0], x[0], fmaf(A[1], x[1], fmaf(A[2], x[2], t[0]))) - y[0];
    float d1 = fmaf(A[3], x[0], fmaf(A[4], x[1], fmaf(A[5], x[2], t[1]))) - y[1];
    float d2 = fmaf(A[6], x[0], fmaf(A[7], x[1], fmaf(A[8], x[2], t[2]))) - y[2];
    return fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
}

/* counts[h] = #{ i : mask[i] && r2(h, i) < thr2 }  (strict, align_geometry.py:123);
 * invalid hypotheses score 0. */
void oracle_ransac_score(const float* A, const float* t, const uint8_t* ok, int n_hyp,
                         const float* xs, const float* ys, const uint8_t* mask, long long m,
                         float thr2, int32_t* counts) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int h = 0; h < n_hyp; ++h) {
        int32_t c = 0;
        if (ok[h]) {
            const float* Ah = A + 9 * (size_t)h;
            const float* th = t + 3 * (size_t)h;
            for (long long i = 0; i < m; ++i)
                if (mask[i] && residual2(Ah, th, xs + 3 * i, ys + 3 * i) < thr2) ++c;
        }
        counts[h] = c;
    }
}

void oracle_ransac_inlier_mask(const float* A, const float* t, const float* xs, const float* ys,
                               const uint8_t* mask, long long m, float thr2, uint8_t* out) {
#pragma omp parallel for
    for (long long i = 0; i < m; ++i)
        out[i] = (mask[i] && residual2(A, t, xs + 3 * i, ys + 3 * i) < thr2) ? 1 : 0;
}

/* SPEC 1: float32 camera-frame unprojection, one rounding per operation. */
void oracle_cam_fast_f32(const float* depth, int n, int H, int W, const float* K /* [n,3,3] */, float* out) {
#pragma omp parallel for
    for (int f = 0; f < n; ++f) {
        const float* k = K + 9 * (size_t)f;
        float fu = k[0], fv = k[4], cu = k[2], cv = k[5];
        float ifu = 1.0f / fu, ifv = 1.0f / fv;
        for (int v = 0; v < H; ++v)
            for (int u = 0; u < W; ++u) {
                size_t i = ((size_t)f * H + v) * W + u;
                float d = depth[i];
                out[3 * i] = (((float)u - cu) * d) * ifu;
                out[3 * i + 1] = (((float)v - cv) * d) * ifv;
                out[3 * i + 2] = d;
            }
    }
}

/* SPEC 5 keys and fixed-point fractions for one point; returns 0 when the point is unusable. */
int oracle_voxel_key(const float* p, float voxel, int64_t* key, int64_t* q /* [3] */) {
    double v = (double)voxel;
    int64_t k[3];
    for (int a = 0; a < 3; ++a) {
        if (!isfinite(p[a])) return 0;
        double quo = (double)p[a] / v;
        double fl = floor(quo);
        if (!(fabs(fl) < 1048576.0)) return 0;
        k[a] = (int64_t)fl;
        q[a] = llrint((quo - fl) * 4294967296.0);
    }
    *key = ((k[0] + 1048576) << 42) | ((k[1] + 1048576) << 21) | (k[2] + 1048576);
    return 1;
}
