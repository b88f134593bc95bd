// unproject_frame.cuh — per-frame constants and the per-pixel unprojection shared by K1
// (unproject.cu) and the fused export kernel (voxel.cu): one definition, so that both produce
// the same float32 points bit for bit.
#pragma once
#include "common.cuh"
#include "sim3_math.cuh"

struct UnprojFrame {            // per-frame constants, staged in shared memory
    double fu, fv, cu, cv;      // closed form
    double kinv[9];             // general
    double M[9], m[3];          // output transform: identity, c2w, or sim3 o c2w
    float cuf, cvf, ifu, ifv;   // fast float32 path
    float Mf[9], mf[3];
};

template <int MODE> struct K1Val { typedef double type; };
template <> struct K1Val<DA3S_UNPROJ_FAST> { typedef float type; };      // the fast path never leaves float32

template <int MODE>             // DA3S_UNPROJ_CLOSED / KINV / FAST
__device__ __forceinline__ void unproject_pixel(const UnprojFrame& f, int u, int v, float d, bool xform,
                                                typename K1Val<MODE>::type& X, typename K1Val<MODE>::type& Y,
                                                typename K1Val<MODE>::type& Z) {
    if constexpr (MODE == DA3S_UNPROJ_FAST) {
        float x, y;
        cam_fast((float)u, (float)v, d, f.cuf, f.cvf, f.ifu, f.ifv, x, y);
        if (xform) {
            X = fmaf(f.Mf[0], x, fmaf(f.Mf[1], y, fmaf(f.Mf[2], d, f.mf[0])));
            Y = fmaf(f.Mf[3], x, fmaf(f.Mf[4], y, fmaf(f.Mf[5], d, f.mf[1])));
            Z = fmaf(f.Mf[6], x, fmaf(f.Mf[7], y, fmaf(f.Mf[8], d, f.mf[2])));
        } else {
            X = x; Y = y; Z = d;
        }
    } else {
        double x, y, z;
        if (MODE == DA3S_UNPROJ_CLOSED) {
            // src/vggt/utils/geometry.py:109-114: float64 sub, mul, div (each rounded once), then float32
            double dd = (double)d;
            x = (double)__double2float_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)u, f.cu), dd), f.fu));
            y = (double)__double2float_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)v, f.cv), dd), f.fv));
            z = dd;
        } else {
            // utils/geometry.py:26-28: K^-1 [u, v, 1] then * depth, float64 throughout
            double uu = (double)u, vv = (double)v, dd = (double)d;
            x = (f.kinv[0] * uu + f.kinv[1] * vv + f.kinv[2]) * dd;
            y = (f.kinv[3] * uu + f.kinv[4] * vv + f.kinv[5]) * dd;
            z = (f.kinv[6] * uu + f.kinv[7] * vv + f.kinv[8]) * dd;
        }
        if (xform) {
            X = f.M[0] * x + f.M[1] * y + f.M[2] * z + f.m[0];
            Y = f.M[3] * x + f.M[4] * y + f.M[5] * z + f.m[1];
            Z = f.M[6] * x + f.M[7] * y + f.M[8] * z + f.m[2];
        } else {
            X = x; Y = y; Z = z;
        }
    }
}


// output transform of a frame: identity, camera-to-world, or Sim(3) o camera-to-world, composed in
// float64 once per frame (utils/da3_streaming.py:639-644 applies the two back to back)
__device__ __forceinline__ void compose_unproj_frame(const da3s_cam& c, const double* s3, bool world, UnprojFrame& fr) {
    fr.fu = c.fu; fr.fv = c.fv; fr.cu = c.cu; fr.cv = c.cv;
    for (int k = 0; k < 9; ++k) fr.kinv[k] = c.kinv[k];
    fr.cuf = c.cu; fr.cvf = c.cv; fr.ifu = c.inv_fu; fr.ifv = c.inv_fv;
    double M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, m[3] = {0, 0, 0};
    if (world) {
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) M[3 * r + k] = c.c2w[4 * r + k];
            m[r] = c.c2w[4 * r + 3];
        }
    }
    if (s3) {
        double s = s3[0];
        double sR[9], M2[9], m2[3];
        for (int k = 0; k < 9; ++k) sR[k] = s * s3[1 + k];
        mat3_mul(sR, M, M2);
        mat3_vec(sR, m, m2);
        for (int k = 0; k < 9; ++k) M[k] = M2[k];
        for (int k = 0; k < 3; ++k) m[k] = m2[k] + s3[10 + k];
    }
    for (int k = 0; k < 9; ++k) { fr.M[k] = M[k]; fr.Mf[k] = (float)M[k]; }
    for (int k = 0; k < 3; ++k) { fr.m[k] = m[k]; fr.mf[k] = (float)m[k]; }
}
