// pair_align.cu — the submap-pair alignment hot path.
//
//   thresholds   exact medians (select.cu) -> conf threshold, depth scale        (device only)
//   K4 RANSAC    3-point hypotheses, float32-FMA inlier counts, winner           (ALU bound)
//   K2/K3 IRLS   fused unproject + joint mask + conf/Huber weight + 20 float64 moments,
//                warp-shuffle -> shared -> per-block partials -> fixed-order final sum and
//                closed-form Umeyama by the LAST block of each pair (ticket)      (HBM bound)
//
// One launch covers every pair of the batch: grid = (overlap * tiles_per_frame, n_pairs).
// Reductions use no floating-point atomics and a tiling that depends only on H*W, so the
// rows are bit-identical from run to run and for any sharding of the pairs over GPUs.
//
// Algorithmic bytes per correspondence and pass: 16 B read (depth+conf of both sides),
// 0 B written.  Moments are accumulated in each overlap frame's CAMERA coordinates and
// moved to world coordinates once per (pair, frame) — the transform is linear in the
// moments — so world mode costs nothing per pixel.
#include "common.cuh"
#include "sim3_math.cuh"
#include <cstdlib>

int da3s_select_impl(da3s_ctx* ctx, const da3s_select_seg* segs, int n_segs, long long max_n,
                     da3s_select_out* out, cudaStream_t st);

#define PA_THREADS 256
#define PA_GROUPS_PER_BLOCK 4096            // 16384 pixels per block; depends on nothing but this constant

struct PairState {
    double s, R[9], t[3];
    double change, mean_res, n_valid;
    int iters, done, status, gate_on;
};

struct PairArgs {
    const da3s_pair* pairs;
    int n_pairs, overlap, H, W;
    long long P;
    int tiles_per_frame;
    int world, valid_depth, huber, variant;
    float depth_eps;
    const float* thr;                       // [n_pairs]
    const float* dscale;                    // [n_pairs] (nullable = 1)
    PairState* state;                       // [n_pairs]
    double* eff;                            // [n_pairs][overlap][EFF_LEN]  residual map of the current estimate
    const float* gate;                      // [n_pairs][12] winning hypothesis (nullable)
    float gate_thr2;
    double* partials;                       // [n_pairs][overlap*tiles][MOM_LEN]
    unsigned int* tickets;                  // [n_pairs]
    int* n_active;                          // pairs active when the IRLS kernel starts (constant while it runs)
    int* done_count;                        // [max_iterations] pairs finished during pass p (persistent kernel's exit test)
    unsigned long long* work_counter;       // [max_iterations] next (pair, tile) item of pass p ([0]: the queue's item ticket)
    int* q_pair;                            // [q_cap] IRLS work queue: entry e = a pair whose next pass may run (-1: not yet published)
    unsigned int* q_reserve;                // next free queue entry
    int q_cap;
    double huber_delta, tol;
    float delta_f, delta2_f, W_f;           // float32 copies read straight from the constant bank in the inner loop
    int max_iterations, min_points, precise;
    double* rows;                           // [n_pairs][16]
    da3s_pair_aux* aux;
};

struct FrameConst {
    float cuA, cvA, ifuA, ifvA, cuB, cvB, ifuB, ifvB;
    float MfA[12], MfB[12];                 // float32 c2w (RANSAC world points, SPEC 4)
    float gate[12];
    double By[9], Bx[9], c[3];              // residual = By y + Bx x + c  (By = I in camera mode)
    float Bpf[9], cpf[3];                   // float32 form |r| = |y + Bp x + cp| (mixed-precision kernel)
    float thr, ds;
    int gate_on;
};

__device__ __forceinline__ void load_frame_const(FrameConst& fc, const da3s_pair& pr, int frame, const PairArgs& a, int pair) {
    const da3s_cam& ca = pr.cam_a[frame];
    const da3s_cam& cb = pr.cam_b[frame];
    fc.cuA = ca.cu; fc.cvA = ca.cv; fc.ifuA = ca.inv_fu; fc.ifvA = ca.inv_fv;
    fc.cuB = cb.cu; fc.cvB = cb.cv; fc.ifuB = cb.inv_fu; fc.ifvB = cb.inv_fv;
    for (int k = 0; k < 12; ++k) { fc.MfA[k] = (float)ca.c2w[k]; fc.MfB[k] = (float)cb.c2w[k]; }
    fc.thr = a.thr[pair];
    fc.ds = a.dscale ? a.dscale[pair] : 1.0f;
    fc.gate_on = 0;
    if (a.gate && a.state && a.state[pair].gate_on) {
        fc.gate_on = 1;
        for (int k = 0; k < 12; ++k) fc.gate[k] = a.gate[12 * (size_t)pair + k];
    }
    if (a.eff) {
        const double* e = a.eff + ((size_t)pair * a.overlap + frame) * EFF_LEN;
        // written by another block in the previous pass of the persistent kernel: read through L2
        for (int k = 0; k < 9; ++k) { fc.By[k] = __ldcg(e + k); fc.Bx[k] = __ldcg(e + 9 + k); fc.Bpf[k] = (float)__ldcg(e + 21 + k); }
        for (int k = 0; k < 3; ++k) { fc.c[k] = __ldcg(e + 18 + k); fc.cpf[k] = (float)__ldcg(e + 30 + k); }
    }
}

// One correspondence: joint mask, float32 camera-frame points of both sides (SPEC 1, 2).
__device__ __forceinline__ bool corr_points(const FrameConst& fc, bool valid_depth, float eps, int u, int v,
                                            float dA, float cA, float dB, float cB,
                                            float* x /* source = B */, float* y /* target = A */, float& dBs) {
    dBs = __fmul_rn(dB, fc.ds);             // solver.py:126: depth * s_depth in float32
    bool keep = (cA > fc.thr) && (cB > fc.thr);
    if (valid_depth) keep = keep && (dA > eps) && (dBs > eps) && is_finite_f(dA) && is_finite_f(dBs);
    float uf = (float)u, vf = (float)v;
    cam_fast(uf, vf, dA, fc.cuA, fc.cvA, fc.ifuA, fc.ifvA, y[0], y[1]);
    y[2] = dA;
    cam_fast(uf, vf, dBs, fc.cuB, fc.cvB, fc.ifuB, fc.ifvB, x[0], x[1]);
    x[2] = dBs;
    return keep;
}

__device__ __forceinline__ void to_world_f32(const float* M, const float* p, float* w) {
    w[0] = fmaf(M[0], p[0], fmaf(M[1], p[1], fmaf(M[2], p[2], M[3])));
    w[1] = fmaf(M[4], p[0], fmaf(M[5], p[1], fmaf(M[6], p[2], M[7])));
    w[2] = fmaf(M[8], p[0], fmaf(M[9], p[1], fmaf(M[10], p[2], M[11])));
}

// SPEC 4 residual: p_i = fma(A_i0,x0, fma(A_i1,x1, fma(A_i2,x2, t_i))); d = p - y; r2 = fma(d0,d0, fma(d1,d1, d2*d2))
__device__ __forceinline__ float residual2_f32(const float* A /* 9 + t 3 */, const float* x, const float* y) {
    float d0 = __fsub_rn(fmaf(A[0], x[0], fmaf(A[1], x[1], fmaf(A[2], x[2], A[9]))), y[0]);
    float d1 = __fsub_rn(fmaf(A[3], x[0], fmaf(A[4], x[1], fmaf(A[5], x[2], A[10]))), y[1]);
    float d2 = __fsub_rn(fmaf(A[6], x[0], fmaf(A[7], x[1], fmaf(A[8], x[2], A[11]))), y[2]);
    return fmaf(d0, d0, fmaf(d1, d1, __fmul_rn(d2, d2)));
}

// RANSAC points of a correspondence (camera or float32-world), written to xs/ys
__device__ __forceinline__ void ransac_points(const FrameConst& fc, int world, const float* x, const float* y, float* xs, float* ys) {
    if (world) { to_world_f32(fc.MfB, x, xs); to_world_f32(fc.MfA, y, ys); }
    else { for (int k = 0; k < 3; ++k) { xs[k] = x[k]; ys[k] = y[k]; } }
}

// ---------------------------------------------------------------------------------
// per-pair solve, run by one thread of the last block
// ---------------------------------------------------------------------------------
__device__ __noinline__ void write_row(const PairArgs& a, int pair, const PairState& st) {
    double* r = a.rows + (size_t)pair * DA3S_ROW_LEN;
    r[DA3S_ROW_S] = st.s;
    for (int k = 0; k < 9; ++k) r[DA3S_ROW_R + k] = st.R[k];
    for (int k = 0; k < 3; ++k) r[DA3S_ROW_T + k] = st.t[k];
    r[DA3S_ROW_NVALID] = st.n_valid;
    r[DA3S_ROW_ITERS] = (double)st.iters;
    r[DA3S_ROW_STATUS] = (double)st.status;
    if (a.aux) { a.aux[pair].mean_residual = st.mean_res; a.aux[pair].last_change = st.change; }
}

__device__ __noinline__ void set_effective(const PairArgs& a, int pair, const PairState& st) {
    for (int f = 0; f < a.overlap; ++f) {
        double* e = a.eff + ((size_t)pair * a.overlap + f) * EFF_LEN;
        if (a.world) {
            const da3s_pair pr = a.pairs[pair];
            effective_residual_transform(st.s, st.R, st.t, pr.cam_b[f].c2w, pr.cam_a[f].c2w, e);
        } else {
            for (int k = 0; k < 9; ++k) { e[k] = (k % 4 == 0) ? 1.0 : 0.0; e[9 + k] = -st.s * st.R[k]; e[21 + k] = e[9 + k]; }
            for (int k = 0; k < 3; ++k) { e[18 + k] = -st.t[k]; e[30 + k] = e[18 + k]; }
        }
    }
}

__device__ __noinline__ void solve_pair(const PairArgs& a, int pair, const double* mom /* world or camera moments */, int pass = 0) {
    PairState st = a.state[pair];
    const double n = mom[MOM_N];
    st.n_valid = n;
    if (n < (double)a.min_points) {                         // utils/align.py:154-156
        st.s = 1.0;
        for (int k = 0; k < 9; ++k) st.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        st.t[0] = st.t[1] = st.t[2] = 0.0;
        st.status = 1; st.done = 1; st.change = 0.0;
        a.state[pair] = st;
        write_row(a, pair, st);
        if (a.done_count) atomicAdd(&a.done_count[pass], 1);
        return;
    }
    double s, R[9], t[3];
    double wscale = a.huber ? (mom[MOM_WMAX] + 1e-8) : 1.0; // utils/align.py:194
    umeyama_from_moments(mom, wscale, a.variant, &s, R, t);
    double dR = 0, dt = 0;
    for (int k = 0; k < 9; ++k) dR += (R[k] - st.R[k]) * (R[k] - st.R[k]);
    for (int k = 0; k < 3; ++k) dt += (t[k] - st.t[k]) * (t[k] - st.t[k]);
    st.change = fabs(s - st.s) + sqrt(dR) + sqrt(dt);       // utils/align.py:200
    st.s = s;
    for (int k = 0; k < 9; ++k) st.R[k] = R[k];
    for (int k = 0; k < 3; ++k) st.t[k] = t[k];
    st.iters += 1;
    st.mean_res = a.precise ? mom[MOM_SR] / n : sqrt(mom[MOM_SR] / n);   // mean |r| (float64 kernel) or rms (mixed kernel)
    if (!a.huber || st.change < a.tol || st.iters >= a.max_iterations) st.done = 1;
    if (st.done) write_row(a, pair, st);
    else set_effective(a, pair, st);
    __threadfence();                        // rows / residual map before the flag that releases them
    a.state[pair] = st;
    if (st.done && a.done_count) atomicAdd(&a.done_count[pass], 1);
}

// ---------------------------------------------------------------------------------
// K2 / K3: fused moments (one IRLS iteration, or the single weighted solve)
// ---------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(PA_THREADS, 2)
pair_moments_kernel(PairArgs a) {
    __shared__ FrameConst fc;
    __shared__ double red[PA_THREADS / 32][MOM_LEN];
    __shared__ double fmom[MOM_LEN];
    __shared__ bool is_last;
    const int pair = blockIdx.y;
    if (a.state[pair].done) return;                          // block-uniform: converged pairs cost nothing
    const da3s_pair pr = a.pairs[pair];
    const int frame = blockIdx.x / a.tiles_per_frame;
    const int tile = blockIdx.x - frame * a.tiles_per_frame;
    if (threadIdx.x == 0) load_frame_const(fc, pr, frame, a, pair);
    __syncthreads();

    double acc[MOM_LEN];
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) acc[k] = 0.0;

    const size_t foff = (size_t)frame * (size_t)a.P;
    const float* dA = pr.depth_a + foff; const float* cA = pr.conf_a + foff;
    const float* dB = pr.depth_b + foff; const float* cB = pr.conf_b + foff;
    const bool huber = a.huber;
    const double delta = a.huber_delta;

    auto accumulate = [&](int u, int v, float da, float ca, float db, float cb) {
        float x[3], y[3], dbs;
        bool keep = corr_points(fc, a.valid_depth, a.depth_eps, u, v, da, ca, db, cb, x, y, dbs);
        if (keep && fc.gate_on) {
            float xs[3], ys[3];
            ransac_points(fc, a.world, x, y, xs, ys);
            keep = residual2_f32(fc.gate, xs, ys) < a.gate_thr2;
        }
        if (!keep) return;
        double w = (double)__fsqrt_rn(__fmul_rn(ca, cb));    // utils/align.py:166 in float32
        const double X0 = x[0], X1 = x[1], X2 = x[2], Y0 = y[0], Y1 = y[1], Y2 = y[2];
        if (huber) {
            double r0 = fc.Bx[0] * X0 + fc.Bx[1] * X1 + fc.Bx[2] * X2 + fc.c[0];
            double r1 = fc.Bx[3] * X0 + fc.Bx[4] * X1 + fc.Bx[5] * X2 + fc.c[1];
            double r2 = fc.Bx[6] * X0 + fc.Bx[7] * X1 + fc.Bx[8] * X2 + fc.c[2];
            if (a.world) {                                  // block-uniform
                r0 += fc.By[0] * Y0 + fc.By[1] * Y1 + fc.By[2] * Y2;
                r1 += fc.By[3] * Y0 + fc.By[4] * Y1 + fc.By[5] * Y2;
                r2 += fc.By[6] * Y0 + fc.By[7] * Y1 + fc.By[8] * Y2;
            } else { r0 += Y0; r1 += Y1; r2 += Y2; }
            double r = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
            if (r > delta) w *= delta / r;                   // utils/align.py:94-109, :186-191
            acc[MOM_SR] += r;
        }
        const double wx0 = w * X0, wx1 = w * X1, wx2 = w * X2;
        const double wy0 = w * Y0, wy1 = w * Y1, wy2 = w * Y2;
        acc[MOM_S0] += w;
        acc[MOM_SX] += wx0; acc[MOM_SX + 1] += wx1; acc[MOM_SX + 2] += wx2;
        acc[MOM_SY] += wy0; acc[MOM_SY + 1] += wy1; acc[MOM_SY + 2] += wy2;
        acc[MOM_SYX + 0] += wy0 * X0; acc[MOM_SYX + 1] += wy0 * X1; acc[MOM_SYX + 2] += wy0 * X2;
        acc[MOM_SYX + 3] += wy1 * X0; acc[MOM_SYX + 4] += wy1 * X1; acc[MOM_SYX + 5] += wy1 * X2;
        acc[MOM_SYX + 6] += wy2 * X0; acc[MOM_SYX + 7] += wy2 * X1; acc[MOM_SYX + 8] += wy2 * X2;
        acc[MOM_SXX + 0] += wx0 * X0; acc[MOM_SXX + 1] += wx0 * X1; acc[MOM_SXX + 2] += wx0 * X2;
        acc[MOM_SXX + 3] += wx1 * X1; acc[MOM_SXX + 4] += wx1 * X2; acc[MOM_SXX + 5] += wx2 * X2;
        acc[MOM_WMAX] = fmax(acc[MOM_WMAX], w);
        acc[MOM_N] += 1.0;
    };

    if (VEC) {
        const long long n_groups = a.P >> 2;
        const long long g_begin = (long long)tile * PA_GROUPS_PER_BLOCK;
        long long g_end = g_begin + PA_GROUPS_PER_BLOCK;
        if (g_end > n_groups) g_end = n_groups;
        const float4* dA4 = reinterpret_cast<const float4*>(dA); const float4* cA4 = reinterpret_cast<const float4*>(cA);
        const float4* dB4 = reinterpret_cast<const float4*>(dB); const float4* cB4 = reinterpret_cast<const float4*>(cB);
#pragma unroll 2
        for (long long g = g_begin + threadIdx.x; g < g_end; g += PA_THREADS) {
            const float4 da = ldg_stream(dA4 + g), ca = ldg_stream(cA4 + g);
            const float4 db = ldg_stream(dB4 + g), cb = ldg_stream(cB4 + g);
            long long pix = g << 2;
            int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
            accumulate(u, v, da.x, ca.x, db.x, cb.x); if (++u == a.W) { u = 0; ++v; }
            accumulate(u, v, da.y, ca.y, db.y, cb.y); if (++u == a.W) { u = 0; ++v; }
            accumulate(u, v, da.z, ca.z, db.z, cb.z); if (++u == a.W) { u = 0; ++v; }
            accumulate(u, v, da.w, ca.w, db.w, cb.w);
        }
    } else {
        const long long p_begin = (long long)tile * PA_GROUPS_PER_BLOCK * 4;
        long long p_end = p_begin + (long long)PA_GROUPS_PER_BLOCK * 4;
        if (p_end > a.P) p_end = a.P;
        for (long long pix = p_begin + threadIdx.x; pix < p_end; pix += PA_THREADS) {
            int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
            accumulate(u, v, dA[pix], cA[pix], dB[pix], cB[pix]);
        }
    }

    // ---- block reduction: shuffle -> shared -> one partial row, fixed order ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) {
        double v = (k == MOM_WMAX) ? warp_max(acc[k]) : warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    const int n_tiles = a.overlap * a.tiles_per_frame;
    double* prow = a.partials + ((size_t)pair * n_tiles + blockIdx.x) * MOM_LEN;
    if (threadIdx.x < MOM_LEN) {
        double v = red[0][threadIdx.x];
        for (int w = 1; w < PA_THREADS / 32; ++w)
            v = (threadIdx.x == MOM_WMAX) ? fmax(v, red[w][threadIdx.x]) : v + red[w][threadIdx.x];
        prow[threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&a.tickets[pair], 1u);
        is_last = (t == (unsigned int)n_tiles - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- last block of the pair: sum partials frame by frame (tile order), to world, solve ----
    __shared__ double wmom[MOM_LEN];
    if (threadIdx.x < MOM_LEN) wmom[threadIdx.x] = 0.0;
    __syncthreads();
    for (int f = 0; f < a.overlap; ++f) {
        if (threadIdx.x < MOM_LEN) {
            const double* src = a.partials + ((size_t)pair * n_tiles + (size_t)f * a.tiles_per_frame) * MOM_LEN + threadIdx.x;
            double v = 0.0;
            for (int t = 0; t < a.tiles_per_frame; ++t) {
                double p = __ldcg(src + (size_t)t * MOM_LEN);
                v = (threadIdx.x == MOM_WMAX) ? fmax(v, p) : v + p;
            }
            fmom[threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (a.world) {
                double m[MOM_LEN], acc2[MOM_LEN];
                for (int k = 0; k < MOM_LEN; ++k) { m[k] = fmom[k]; acc2[k] = wmom[k]; }
                moments_to_world_add(m, pr.cam_b[f].c2w, pr.cam_a[f].c2w, acc2);
                for (int k = 0; k < MOM_LEN; ++k) wmom[k] = acc2[k];
            } else {
                for (int k = 0; k < MOM_LEN; ++k)
                    wmom[k] = (k == MOM_WMAX) ? fmax(wmom[k], fmom[k]) : wmom[k] + fmom[k];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double m[MOM_LEN];
        for (int k = 0; k < MOM_LEN; ++k) m[k] = wmom[k];
        solve_pair(a, pair, m);
        a.tickets[pair] = 0;
    }
}

// ---------------------------------------------------------------------------------
// K2 / K3, mixed precision (default): the per-correspondence arithmetic runs in float32
// FMA on points centred at a per-tile pivot, in micro-batches of 8 correspondences whose 22
// partial sums are flushed into float64 accumulators; tiles are un-pivoted and combined in
// float64.  Rationale (profiles/r1_baseline_*): with float64 per-point arithmetic the kernel is
// bound by the FP64 pipe (64 lanes/clk/SM) and the conversion unit, not by HBM.  Deviation from
// the all-float64 oracle: <= 1e-8 relative on (s, R, t) at 5e4..3e5 correspondences
// (scratch emulation + tests), two orders inside the 1e-6 contract.  `precise = 1` selects
// the all-float64 kernel above.
// ---------------------------------------------------------------------------------
#define PM_THREADS 128
#define PM_NMOM 22                          // S0, Sx[3], Sy[3], Syx[9], Sxx[6]
#ifndef PM_DEPTH
#define PM_DEPTH 3                          // groups (64 B each) a thread keeps in flight through cp.async; 0 = two register buffers
#endif
#ifndef PM_PK_BLOCKS
#define PM_PK_BLOCKS 3                      // resident blocks per SM the packed kernel is compiled for (registers)
#endif
#define PM_RING_BYTES ((PM_DEPTH > 0 ? (PM_DEPTH + 1) : 0) * 4 * PM_THREADS * 16)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }


__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// IRLS work queue.  An ENTRY is a pair whose next pass may run (its residual map is published); entry e stands for the
// items e * n_tiles .. (e + 1) * n_tiles - 1 of the global item ticket.  The pairs active at the start fill the first
// entries (this kernel, pair order); the block that solves a pair appends it again unless the pair has finished.
__global__ void __launch_bounds__(1024)
pm_queue_init_kernel(PairArgs a) {
    __shared__ int base_sh;
    if (threadIdx.x == 0) base_sh = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        int base = 0;
        for (int p0 = 0; p0 < a.n_pairs; p0 += 32) {
            const int p = p0 + (int)threadIdx.x;
            const bool act = p < a.n_pairs && a.state[p].done == 0;
            const unsigned int m = __ballot_sync(0xffffffffu, act);
            if (act) a.q_pair[base + __popc(m & ((1u << threadIdx.x) - 1u))] = p;
            base += __popc(m);
        }
        if (threadIdx.x == 0) { base_sh = base; *a.q_reserve = (unsigned int)base; *a.n_active = base; a.done_count[0] = 0; a.work_counter[0] = 0ull; }
    }
    __syncthreads();
    for (int e = base_sh + (int)threadIdx.x; e < a.q_cap; e += 1024) a.q_pair[e] = -1;
}

// thread 0 of a block: the pair of queue entry e, waiting until the entry is published; -1 once every pair has finished
__device__ __forceinline__ int pm_wait_entry(const PairArgs& a, long long e) {
    for (;;) {
        if (e < (long long)a.q_cap) {
            const int p = ld_acquire_s32(a.q_pair + e);
            if (p >= 0) return p;
        }
        if (*((volatile int*)&a.done_count[0]) >= *((volatile int*)a.n_active)) return -1;      // nothing will be appended any more
        __nanosleep(200);
    }
}

// PERSISTENT: one launch runs every IRLS iteration of every pair.  Blocks draw (queue entry, tile) items from one global
// ticket; the block that finishes the last tile of a pair's pass (per-pair ticket) solves the pair and, unless it has
// finished, appends it to the queue again — pairs iterate independently, there is no barrier between passes — and the
// kernel ends when every active pair has finished: no empty launches, no host involvement between iterations.
//
// PK (default whenever W is even): the two horizontally adjacent pixels of a float4 half share every float32
// instruction — packed FFMA2 / FMUL2 / FADD2 on register pairs (.x = pixel 2i, .y = pixel 2i + 1; W even means a pair
// never straddles a row).  The kernel is bound by instruction issue, not by the float32 pipe (profiles/r2_fp32_pipes.txt:
// a scalar FFMA and a packed FFMA2 both issue once per clock per sub-partition, the packed one doing twice the work), so
// halving the floating-point instruction count is what moves it.  In this form the points are built from the pixel's ray,
// x = d (a, b, 1) with a = (u - cu) / fu computed once per pair, and only the depth coordinate is pivoted (the lateral
// coordinates are centred on the principal point already).
template <bool VEC, bool GATE, bool HUBER, bool PK>
__global__ void __launch_bounds__(PM_THREADS, PK ? PM_PK_BLOCKS : 4)
pair_moments_mixed_kernel(PairArgs a) {
    __shared__ double red[MOM_LEN][PM_THREADS];     // block reduction scratch (25.6 KB)
    __shared__ FrameConst fc;
    __shared__ float piv[6];                        // pivot: x (source) then y (target), float32 camera-frame point
    __shared__ double fmom[MOM_LEN];
    __shared__ double wmom[MOM_LEN];
    __shared__ double part[MOM_LEN][PM_THREADS / 32];
    __shared__ bool is_last;
    __shared__ int cur_pair;                       // the current item's pair, -1 once every pair has finished
    __shared__ long long cur_item;
    const int n_tiles_all = a.overlap * a.tiles_per_frame;
    const long long n_items = (long long)a.n_pairs * n_tiles_all;
  __shared__ float fcm[32];                         // per-item constants, filled by warp 0 (one value per lane)
  // with few items per block an early request would take work away from idle blocks
  const bool prefetch = n_items >= 4ll * gridDim.x;
  const int pass = 0;
  {
   // dynamic work distribution over the queue's items: the block that happens to solve a pair (serial epilogue) simply
   // takes fewer tiles.  The NEXT ticket is requested while the current item is processed (the atomic's round trip is
   // never waited for); a ticket whose entry is not published yet is waited for by thread 0 (pm_wait_entry).
   __syncthreads();
   if (threadIdx.x == 0) {
       const long long it = (long long)atomicAdd(&a.work_counter[0], 1ull);
       cur_item = it;
       cur_pair = pm_wait_entry(a, it / n_tiles_all);
   }
   __syncthreads();
   for (;;) {
    const long long item = cur_item;
    const int pair = cur_pair;
    if (pair < 0) break;
    long long next_item = 0;
    if (prefetch && threadIdx.x == 0) next_item = (long long)atomicAdd(&a.work_counter[0], 1ull);
   do {
    const int item_tile = (int)(item % n_tiles_all);
    const da3s_pair pr = a.pairs[pair];
    const int frame = item_tile / a.tiles_per_frame;
    const int tile = item_tile - frame * a.tiles_per_frame;
    const size_t foff = (size_t)frame * (size_t)a.P;
    const float* dA = pr.depth_a + foff; const float* cA = pr.conf_a + foff;
    const float* dB = pr.depth_b + foff; const float* cB = pr.conf_b + foff;
    const long long p_begin = (long long)tile * PA_GROUPS_PER_BLOCK * 4;
    long long p_end = p_begin + (long long)PA_GROUPS_PER_BLOCK * 4;
    if (p_end > a.P) p_end = a.P;
    if (threadIdx.x < 32) {
        // item constants, one independent load per lane (a single round trip instead of a serial prologue)
        const int t = threadIdx.x;
        const int vc = a.H / 2, uc = a.W / 2;
        const long long pc = (long long)vc * a.W + uc;
        const double* e = a.eff + ((size_t)pair * a.overlap + frame) * EFF_LEN;     // written in the previous pass: read through L2
        float val = 0.0f;
        if (t < 8) val = reinterpret_cast<const float*>(t < 4 ? &pr.cam_a[frame] : &pr.cam_b[frame])[2 + (t & 3)];   // cu cv 1/fu 1/fv
        else if (t < 17) val = (float)__ldcg(e + 21 + (t - 8));
        else if (t < 20) val = (float)__ldcg(e + 30 + (t - 17));
        else if (t == 20) val = a.thr[pair];
        else if (t == 21) val = a.dscale ? a.dscale[pair] : 1.0f;
        else if (t == 22) val = dA[pc];
        else if (t == 23) val = dB[pc];
        else if (t == 24) val = (GATE && a.gate && *((volatile int*)&a.state[pair].gate_on)) ? 1.0f : 0.0f;
        fcm[t] = val;
        __syncwarp();
        // pivot = the correspondence at the frame's centre pixel (corr_points arithmetic): the same for every tile of
        // the frame, so per-tile partials add directly and are un-pivoted once per frame by the last block
        if (t < 6) {
            const bool is_x = t < 3;
            const float d = is_x ? __fmul_rn(fcm[23], fcm[21]) : fcm[22];
            const float* in = is_x ? &fcm[4] : &fcm[0];
            float p0, p1;
            cam_fast((float)uc, (float)vc, d, in[0], in[1], in[2], in[3], p0, p1);
            const float v = (t % 3 == 0) ? p0 : ((t % 3 == 1) ? p1 : d);
            piv[t] = (is_finite_f(v) && (!PK || t % 3 == 2)) ? v : 0.0f;       // PK: depth coordinate only
        }
    } else if (GATE && threadIdx.x == 32) {
        load_frame_const(fc, pr, frame, a, pair);    // float32 c2w + winning hypothesis for the RANSAC gate
    }
    __syncthreads();
    // block-uniform constants into registers
    const float px0 = piv[0], px1 = piv[1], px2 = piv[2], py0 = piv[3], py1 = piv[4], py2 = piv[5];
    const float cuA = fcm[0], cvA = fcm[1], ifuA = fcm[2], ifvA = fcm[3], cuB = fcm[4], cvB = fcm[5], ifuB = fcm[6], ifvB = fcm[7];
    const float thr = fcm[20], ds = fcm[21], eps = a.depth_eps;
    const bool any_depth = !a.valid_depth, gate_on = GATE && fcm[24] != 0.0f;
    float B[9], c3[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) B[k] = fcm[8 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) c3[k] = fcm[17 + k];
    // per-thread float64 accumulators live in the reduction scratch (column = thread): keeps the
    // kernel at <= 128 registers (4 blocks/SM) and the final block reduction reads them in place
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) red[k][threadIdx.x] = 0.0;
    float m[PM_NMOM];
#pragma unroll
    for (int k = 0; k < PM_NMOM; ++k) m[k] = 0.0f;
    float wmax = 0.0f, r2sum = 0.0f;
    int cnt = 0;

    // one correspondence, branch-free: a rejected pixel contributes weight 0 on sanitised depths
    auto accumulate = [&](float uf, float vf, float da, float ca, float db, float cb) {
        float dbs = __fmul_rn(db, ds);
        // predicate logic only (no short-circuit): a data-dependent branch here diverges in every warp
        const bool depth_ok = (da > eps) & (dbs > eps) & is_finite_f(da) & is_finite_f(dbs);
        bool keep = (ca > thr) & (cb > thr) & (depth_ok | any_depth);
        // sanitise the depths of a rejected pixel (its points become finite zeros): nothing
        // non-finite can reach the sums through 0 * x
        da = keep ? da : 0.0f; dbs = keep ? dbs : 0.0f;
        float x0, x1, y0, y1;
        cam_fast(uf, vf, da, cuA, cvA, ifuA, ifvA, y0, y1);
        cam_fast(uf, vf, dbs, cuB, cvB, ifuB, ifvB, x0, x1);
        if (GATE && gate_on) {                                  // block-uniform
            float x[3] = {x0, x1, dbs}, y[3] = {y0, y1, da}, xs[3], ys[3];
            ransac_points(fc, a.world, x, y, xs, ys);
            keep = keep & (residual2_f32(fc.gate, xs, ys) < a.gate_thr2);
        }
        const float x2 = dbs, y2 = da;
        float w = sqrt_approx(keep ? ca * cb : 0.0f);           // utils/align.py:166 (<= 1 ulp from sqrt_f32); sqrt(0) = 0
        if (HUBER) {
            const float r0 = fmaf(B[0], x0, fmaf(B[1], x1, fmaf(B[2], x2, y0 + c3[0])));
            const float r1 = fmaf(B[3], x0, fmaf(B[4], x1, fmaf(B[5], x2, y1 + c3[1])));
            const float r2 = fmaf(B[6], x0, fmaf(B[7], x1, fmaf(B[8], x2, y2 + c3[2])));
            const float rr = fmaf(r0, r0, fmaf(r1, r1, r2 * r2));
            const float hub = a.delta_f * rsqrt_approx(rr);      // Huber: delta / r  (utils/align.py:94-109)
            w = (rr > a.delta2_f) ? w * hub : w;
            r2sum += keep ? rr : 0.0f;
        }
        const float xc0 = x0 - px0, xc1 = x1 - px1, xc2 = x2 - px2;
        const float yc0 = y0 - py0, yc1 = y1 - py1, yc2 = y2 - py2;
        const float wx0 = w * xc0, wx1 = w * xc1, wx2 = w * xc2;
        m[0] += w;
        m[1] += wx0; m[2] += wx1; m[3] += wx2;
        m[4] = fmaf(w, yc0, m[4]); m[5] = fmaf(w, yc1, m[5]); m[6] = fmaf(w, yc2, m[6]);
        m[7] = fmaf(yc0, wx0, m[7]);   m[8] = fmaf(yc0, wx1, m[8]);   m[9] = fmaf(yc0, wx2, m[9]);
        m[10] = fmaf(yc1, wx0, m[10]); m[11] = fmaf(yc1, wx1, m[11]); m[12] = fmaf(yc1, wx2, m[12]);
        m[13] = fmaf(yc2, wx0, m[13]); m[14] = fmaf(yc2, wx1, m[14]); m[15] = fmaf(yc2, wx2, m[15]);
        m[16] = fmaf(wx0, xc0, m[16]); m[17] = fmaf(wx0, xc1, m[17]); m[18] = fmaf(wx0, xc2, m[18]);
        m[19] = fmaf(wx1, xc1, m[19]); m[20] = fmaf(wx1, xc2, m[20]); m[21] = fmaf(wx2, xc2, m[21]);
        wmax = fmaxf(wmax, w);
        cnt += keep ? 1 : 0;
    };
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < PM_NMOM; ++k) { red[k][threadIdx.x] += (double)m[k]; m[k] = 0.0f; }
        red[MOM_SR][threadIdx.x] += (double)r2sum; r2sum = 0.0f;
    };
    auto group = [&](float uf, float vf, const float4& da, const float4& ca, const float4& db, const float4& cb) {
        accumulate(uf, vf, da.x, ca.x, db.x, cb.x); uf += 1.0f; if (uf == a.W_f) { uf = 0.0f; vf += 1.0f; }
        accumulate(uf, vf, da.y, ca.y, db.y, cb.y); uf += 1.0f; if (uf == a.W_f) { uf = 0.0f; vf += 1.0f; }
        accumulate(uf, vf, da.z, ca.z, db.z, cb.z); uf += 1.0f; if (uf == a.W_f) { uf = 0.0f; vf += 1.0f; }
        accumulate(uf, vf, da.w, ca.w, db.w, cb.w);
    };

    // ---- PK: two adjacent pixels per instruction ----
    float2 m2[PK ? PM_NMOM : 1];
    float2 r2sum2 = make_float2(0.0f, 0.0f);
    if (PK) {
#pragma unroll
        for (int k = 0; k < (PK ? PM_NMOM : 1); ++k) m2[k] = make_float2(0.0f, 0.0f);
    }
    const float2 ds2 = make_float2(ds, ds), pz2x = make_float2(px2, px2), pz2y = make_float2(py2, py2);
    auto dup = [](float v) { return make_float2(v, v); };
    auto accumulate2 = [&](float uf, float vf, float2 da, const float2 ca, const float2 db, const float2 cb) {
        float2 dbs = __fmul2_rn(db, ds2);
        const bool ok0 = (da.x > eps) & (dbs.x > eps) & is_finite_f(da.x) & is_finite_f(dbs.x);
        const bool ok1 = (da.y > eps) & (dbs.y > eps) & is_finite_f(da.y) & is_finite_f(dbs.y);
        bool k0 = (ca.x > thr) & (cb.x > thr) & (ok0 | any_depth);
        bool k1 = (ca.y > thr) & (cb.y > thr) & (ok1 | any_depth);
        da = make_float2(k0 ? da.x : 0.0f, k1 ? da.y : 0.0f);
        dbs = make_float2(k0 ? dbs.x : 0.0f, k1 ? dbs.y : 0.0f);
        // rays of the two pixels: a = (u - cu) / fu (pixel 2i, 2i + 1), b = (v - cv) / fv (shared)
        const float aA = (uf - cuA) * ifuA, aB = (uf - cuB) * ifuB;
        const float2 aA2 = make_float2(aA, aA + ifuA), aB2 = make_float2(aB, aB + ifuB);
        const float2 bA2 = dup((vf - cvA) * ifvA), bB2 = dup((vf - cvB) * ifvB);
        const float2 y0 = __fmul2_rn(da, aA2), y1 = __fmul2_rn(da, bA2), x0 = __fmul2_rn(dbs, aB2), x1 = __fmul2_rn(dbs, bB2);
        if (GATE && gate_on) {                                  // block-uniform; SPEC 4 arithmetic on SPEC 1 points, per pixel
            float xa[3], ya[3], xs[3], ys[3];
            cam_fast(uf, vf, da.x, cuA, cvA, ifuA, ifvA, ya[0], ya[1]); ya[2] = da.x;
            cam_fast(uf, vf, dbs.x, cuB, cvB, ifuB, ifvB, xa[0], xa[1]); xa[2] = dbs.x;
            ransac_points(fc, a.world, xa, ya, xs, ys);
            k0 = k0 & (residual2_f32(fc.gate, xs, ys) < a.gate_thr2);
            cam_fast(uf + 1.0f, vf, da.y, cuA, cvA, ifuA, ifvA, ya[0], ya[1]); ya[2] = da.y;
            cam_fast(uf + 1.0f, vf, dbs.y, cuB, cvB, ifuB, ifvB, xa[0], xa[1]); xa[2] = dbs.y;
            ransac_points(fc, a.world, xa, ya, xs, ys);
            k1 = k1 & (residual2_f32(fc.gate, xs, ys) < a.gate_thr2);
        }
        const float2 cc = __fmul2_rn(ca, cb);
        float2 w = make_float2(sqrt_approx(k0 ? cc.x : 0.0f), sqrt_approx(k1 ? cc.y : 0.0f));   // utils/align.py:166
        if (HUBER) {
            const float2 r0 = __ffma2_rn(dup(B[0]), x0, __ffma2_rn(dup(B[1]), x1, __ffma2_rn(dup(B[2]), dbs, __fadd2_rn(y0, dup(c3[0])))));
            const float2 r1 = __ffma2_rn(dup(B[3]), x0, __ffma2_rn(dup(B[4]), x1, __ffma2_rn(dup(B[5]), dbs, __fadd2_rn(y1, dup(c3[1])))));
            const float2 r2 = __ffma2_rn(dup(B[6]), x0, __ffma2_rn(dup(B[7]), x1, __ffma2_rn(dup(B[8]), dbs, __fadd2_rn(da, dup(c3[2])))));
            const float2 rr = __ffma2_rn(r0, r0, __ffma2_rn(r1, r1, __fmul2_rn(r2, r2)));
            const float2 wh = __fmul2_rn(w, make_float2(a.delta_f * rsqrt_approx(rr.x), a.delta_f * rsqrt_approx(rr.y)));
            w = make_float2((rr.x > a.delta2_f) ? wh.x : w.x, (rr.y > a.delta2_f) ? wh.y : w.y);         // Huber: delta / r
            r2sum2 = __fadd2_rn(r2sum2, make_float2(k0 ? rr.x : 0.0f, k1 ? rr.y : 0.0f));
        }
        const float2 xc2 = __fadd2_rn(dbs, make_float2(-pz2x.x, -pz2x.y)), yc2 = __fadd2_rn(da, make_float2(-pz2y.x, -pz2y.y));
        const float2 wx0 = __fmul2_rn(w, x0), wx1 = __fmul2_rn(w, x1), wx2 = __fmul2_rn(w, xc2);
        m2[0] = __fadd2_rn(m2[0], w);
        m2[1] = __fadd2_rn(m2[1], wx0); m2[2] = __fadd2_rn(m2[2], wx1); m2[3] = __fadd2_rn(m2[3], wx2);
        m2[4] = __ffma2_rn(w, y0, m2[4]); m2[5] = __ffma2_rn(w, y1, m2[5]); m2[6] = __ffma2_rn(w, yc2, m2[6]);
        m2[7] = __ffma2_rn(y0, wx0, m2[7]);    m2[8] = __ffma2_rn(y0, wx1, m2[8]);    m2[9] = __ffma2_rn(y0, wx2, m2[9]);
        m2[10] = __ffma2_rn(y1, wx0, m2[10]);  m2[11] = __ffma2_rn(y1, wx1, m2[11]);  m2[12] = __ffma2_rn(y1, wx2, m2[12]);
        m2[13] = __ffma2_rn(yc2, wx0, m2[13]); m2[14] = __ffma2_rn(yc2, wx1, m2[14]); m2[15] = __ffma2_rn(yc2, wx2, m2[15]);
        m2[16] = __ffma2_rn(wx0, x0, m2[16]);  m2[17] = __ffma2_rn(wx0, x1, m2[17]);  m2[18] = __ffma2_rn(wx0, xc2, m2[18]);
        m2[19] = __ffma2_rn(wx1, x1, m2[19]);  m2[20] = __ffma2_rn(wx1, xc2, m2[20]); m2[21] = __ffma2_rn(wx2, xc2, m2[21]);
        wmax = fmaxf(wmax, fmaxf(w.x, w.y));
        cnt += (k0 ? 1 : 0) + (k1 ? 1 : 0);
    };
    auto flush2 = [&]() {
#pragma unroll
        for (int k = 0; k < (PK ? PM_NMOM : 1); ++k) { red[k][threadIdx.x] += (double)(m2[k].x + m2[k].y); m2[k] = make_float2(0.0f, 0.0f); }
        red[MOM_SR][threadIdx.x] += (double)(r2sum2.x + r2sum2.y); r2sum2 = make_float2(0.0f, 0.0f);
    };
    auto group2 = [&](float uf, float vf, const float4& da, const float4& ca, const float4& db, const float4& cb) {
        accumulate2(uf, vf, make_float2(da.x, da.y), make_float2(ca.x, ca.y), make_float2(db.x, db.y), make_float2(cb.x, cb.y));
        uf += 2.0f; if (uf >= a.W_f) { uf -= a.W_f; vf += 1.0f; }        // W even: a pair never straddles a row
        accumulate2(uf, vf, make_float2(da.z, da.w), make_float2(ca.z, ca.w), make_float2(db.z, db.w), make_float2(cb.z, cb.w));
    };

    if (VEC) {
        const long long n_groups = a.P >> 2;
        const long long g_begin = (long long)tile * PA_GROUPS_PER_BLOCK;
        long long g_end = g_begin + PA_GROUPS_PER_BLOCK;
        if (g_end > n_groups) g_end = n_groups;
        const float4* dA4 = reinterpret_cast<const float4*>(dA); const float4* cA4 = reinterpret_cast<const float4*>(cA);
        const float4* dB4 = reinterpret_cast<const float4*>(dB); const float4* cB4 = reinterpret_cast<const float4*>(cB);
        // pixel coordinates are carried incrementally (float): one division per thread per tile
        const int step_px = PM_THREADS * 4, step_vi = step_px / a.W;
        const float step_v = (float)step_vi, step_u = (float)(step_px - step_vi * a.W);
        float u0f, v0f;
        { const long long p0 = (g_begin + threadIdx.x) << 2; const int v0 = (int)(p0 / a.W); v0f = (float)v0; u0f = (float)(int)(p0 - (long long)v0 * a.W); }
        const long long g_first = g_begin + threadIdx.x;
        int left = g_first < g_end ? (int)((g_end - g_first + PM_THREADS - 1) / PM_THREADS) : 0;
        const float4* pdA = dA4 + g_first; const float4* pcA = cA4 + g_first;
        const float4* pdB = dB4 + g_first; const float4* pcB = cB4 + g_first;
        auto advance = [&]() {
            u0f += step_u; v0f += step_v;
            if (u0f >= a.W_f) { u0f -= a.W_f; v0f += 1.0f; }
        };
#if PM_DEPTH > 0
        // asynchronous copy ring (cp.async, 16 bytes per array and group, straight into shared memory): PM_DEPTH groups
        // = PM_DEPTH x 64 bytes per thread are in flight while the current group is accumulated, without holding them in
        // registers.  Measured on loop512 (512 pairs, 2.69 passes): with the scalar inner loop the kernel is issue bound and
        // the ring changes nothing (IRLS 1.64 ms at depth 0, 2, 3); with the packed inner loop the issue rate drops to 44 %,
        // global-load latency becomes the first stall, and depth 3 gives 1.50 ms (depth 2: 1.57, 4: 1.53, 6: 1.74).
        // Every thread reads back only what it copied itself (no block barrier); the slot refilled in an iteration is the one
        // consumed in the PREVIOUS iteration.  One (possibly empty) commit per iteration keeps wait_group's count uniform.
        extern __shared__ float4 pm_ring[];                     // [PM_DEPTH + 1][4][PM_THREADS]
        auto slot_ptr = [&](int sl, int arr) { return pm_ring + ((size_t)(sl * 4 + arr) * PM_THREADS + threadIdx.x); };
        auto issue = [&](int j) {                                // group j of this thread (j < left) into slot j % (PM_DEPTH + 1)
            if (j < left) {
                const int sl = j % (PM_DEPTH + 1);
                const size_t off = (size_t)j * PM_THREADS;
                cp_async16(slot_ptr(sl, 0), pdA + off); cp_async16(slot_ptr(sl, 1), pcA + off);
                cp_async16(slot_ptr(sl, 2), pdB + off); cp_async16(slot_ptr(sl, 3), pcB + off);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int j = 0; j < PM_DEPTH; ++j) issue(j);
        for (int j = 0; j < left; ++j) {
            asm volatile("cp.async.wait_group %0;" ::"n"(PM_DEPTH - 1) : "memory");
            const int sl = j % (PM_DEPTH + 1);
            const float4 a0 = *slot_ptr(sl, 0), a1 = *slot_ptr(sl, 1), a2 = *slot_ptr(sl, 2), a3 = *slot_ptr(sl, 3);
            issue(j + PM_DEPTH);
            if (PK) group2(u0f, v0f, a0, a1, a2, a3); else group(u0f, v0f, a0, a1, a2, a3); advance();
            if ((j & 3) == 3) { if (PK) flush2(); else flush(); }
        }
        if (left & 3) { if (PK) flush2(); else flush(); }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
        // software pipeline: the 4 loads (64 B) of the thread's NEXT group are in flight while the current
        // group (4 correspondences) is accumulated; two register buffers alternate (no copies), running
        // pointers (no per-group address arithmetic); float64 flush every 4 groups (16 correspondences)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a0 = z4, a1 = z4, a2 = z4, a3 = z4, b0 = z4, b1 = z4, b2 = z4, b3 = z4;
        if (left > 0) { a0 = ldg_stream(pdA); a1 = ldg_stream(pcA); a2 = ldg_stream(pdB); a3 = ldg_stream(pcB); }
        while (left > 0) {                                           // 4 groups per trip
            if (left > 1) { b0 = ldg_stream(pdA + PM_THREADS); b1 = ldg_stream(pcA + PM_THREADS); b2 = ldg_stream(pdB + PM_THREADS); b3 = ldg_stream(pcB + PM_THREADS); }
            if (PK) group2(u0f, v0f, a0, a1, a2, a3); else group(u0f, v0f, a0, a1, a2, a3); advance();
            if (left > 1) {
                if (left > 2) { a0 = ldg_stream(pdA + 2 * PM_THREADS); a1 = ldg_stream(pcA + 2 * PM_THREADS); a2 = ldg_stream(pdB + 2 * PM_THREADS); a3 = ldg_stream(pcB + 2 * PM_THREADS); }
                if (PK) group2(u0f, v0f, b0, b1, b2, b3); else group(u0f, v0f, b0, b1, b2, b3); advance();
            }
            if (left > 2) {
                if (left > 3) { b0 = ldg_stream(pdA + 3 * PM_THREADS); b1 = ldg_stream(pcA + 3 * PM_THREADS); b2 = ldg_stream(pdB + 3 * PM_THREADS); b3 = ldg_stream(pcB + 3 * PM_THREADS); }
                if (PK) group2(u0f, v0f, a0, a1, a2, a3); else group(u0f, v0f, a0, a1, a2, a3); advance();
            }
            if (left > 3) {
                if (left > 4) { a0 = ldg_stream(pdA + 4 * PM_THREADS); a1 = ldg_stream(pcA + 4 * PM_THREADS); a2 = ldg_stream(pdB + 4 * PM_THREADS); a3 = ldg_stream(pcB + 4 * PM_THREADS); }
                if (PK) group2(u0f, v0f, b0, b1, b2, b3); else group(u0f, v0f, b0, b1, b2, b3); advance();
            }
            if (PK) flush2(); else flush();
            pdA += 4 * PM_THREADS; pcA += 4 * PM_THREADS; pdB += 4 * PM_THREADS; pcB += 4 * PM_THREADS;
            left -= 4;
        }
#endif
    } else {
        int since = 0;
        for (long long pix = p_begin + threadIdx.x; pix < p_end; pix += PM_THREADS) {
            const int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
            accumulate((float)u, (float)v, dA[pix], cA[pix], dB[pix], cB[pix]);
            if (++since == 8) { flush(); since = 0; }
        }
        flush();
    }

    // ---- block reduction through shared memory (float64), fixed order ----
    red[MOM_WMAX][threadIdx.x] = (double)wmax;
    red[MOM_N][threadIdx.x] = (double)cnt;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < MOM_LEN * (PM_THREADS / 32)) {               // 100 threads: (moment k, 32-thread slice)
        const int k = threadIdx.x / (PM_THREADS / 32), sl = threadIdx.x % (PM_THREADS / 32);
        const double* src = &red[k][sl * 32];
        // rotated start index per lane avoids bank conflicts; four independent chains, fixed order
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        const bool is_max = (k == MOM_WMAX);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const double p0 = src[(j + lane) & 31], p1 = src[(j + 1 + lane) & 31], p2 = src[(j + 2 + lane) & 31], p3 = src[(j + 3 + lane) & 31];
            v0 = is_max ? fmax(v0, p0) : v0 + p0; v1 = is_max ? fmax(v1, p1) : v1 + p1;
            v2 = is_max ? fmax(v2, p2) : v2 + p2; v3 = is_max ? fmax(v3, p3) : v3 + p3;
        }
        const double v = is_max ? fmax(fmax(v0, v1), fmax(v2, v3)) : (v0 + v1) + (v2 + v3);
        part[k][sl] = v;
    }
    __syncthreads();
    const int n_tiles = a.overlap * a.tiles_per_frame;
    if (threadIdx.x < MOM_LEN) {
        double v = part[threadIdx.x][0];
        for (int w = 1; w < PM_THREADS / 32; ++w) v = (threadIdx.x == MOM_WMAX) ? fmax(v, part[threadIdx.x][w]) : v + part[threadIdx.x][w];
        a.partials[((size_t)pair * n_tiles + item_tile) * MOM_LEN + threadIdx.x] = v;      // still about the frame pivot
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&a.tickets[pair], 1u);
        is_last = (t == (unsigned int)n_tiles - 1);
    }
    __syncthreads();
    if (!is_last) continue;
    __threadfence();

    // ---- last block of the pair: sum partials frame by frame (tile order), un-pivot, to world, solve ----
    if (threadIdx.x < MOM_LEN) wmom[threadIdx.x] = 0.0;
    __syncthreads();
    for (int f = 0; f < a.overlap; ++f) {
        if (threadIdx.x < MOM_LEN) {
            const double* src = a.partials + ((size_t)pair * n_tiles + (size_t)f * a.tiles_per_frame) * MOM_LEN + threadIdx.x;
            double v = 0.0;
            for (int t = 0; t < a.tiles_per_frame; ++t) {
                double p = __ldcg(src + (size_t)t * MOM_LEN);
                v = (threadIdx.x == MOM_WMAX) ? fmax(v, p) : v + p;
            }
            fmom[threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // frame f's pivot (same rule as above), un-pivot (exact polynomial identities, float64), then to world
            FrameConst& g = fc;
            load_frame_const(g, pr, f, a, pair);
            const int vc = a.H / 2, uc = a.W / 2;
            const size_t pc = (size_t)f * (size_t)a.P + (size_t)vc * a.W + uc;
            float x[3], y[3], dbs;
            corr_points(g, false, 0.0f, uc, vc, pr.depth_a[pc], 1.0f, pr.depth_b[pc], 1.0f, x, y, dbs);
            double sx[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, sy[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
            for (int k = 0; k < 3; ++k) {
                const bool use = !PK || k == 2;                  // the packed kernel pivots the depth coordinate only
                sx[4 * k + 3] = (use && is_finite_f(x[k])) ? (double)x[k] : 0.0;
                sy[4 * k + 3] = (use && is_finite_f(y[k])) ? (double)y[k] : 0.0;
            }
            double mm[MOM_LEN], cam[MOM_LEN];
            for (int k = 0; k < MOM_LEN; ++k) { mm[k] = fmom[k]; cam[k] = 0.0; }
            moments_to_world_add(mm, sx, sy, cam);
            if (a.world) {
                double acc2[MOM_LEN];
                for (int k = 0; k < MOM_LEN; ++k) acc2[k] = wmom[k];
                moments_to_world_add(cam, pr.cam_b[f].c2w, pr.cam_a[f].c2w, acc2);
                for (int k = 0; k < MOM_LEN; ++k) wmom[k] = acc2[k];
            } else {
                for (int k = 0; k < MOM_LEN; ++k)
                    wmom[k] = (k == MOM_WMAX) ? fmax(wmom[k], cam[k]) : wmom[k] + cam[k];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double mm[MOM_LEN];
        for (int k = 0; k < MOM_LEN; ++k) mm[k] = wmom[k];
        a.tickets[pair] = 0;
        solve_pair(a, pair, mm, pass);              // ends with: fence, state, done_count
        if (!*((volatile int*)&a.state[pair].done)) {
            // the pair's next pass may start: its residual map, state and ticket are written; publish the entry
            __threadfence();
            const unsigned int e = atomicAdd(a.q_reserve, 1u);
            if (e < (unsigned int)a.q_cap) st_release_s32(a.q_pair + e, pair);
        }
    }
   } while (0);
    __syncthreads();                                // everyone is done with cur_item and the shared scratch
    if (threadIdx.x == 0) {
        if (!prefetch) next_item = (long long)atomicAdd(&a.work_counter[0], 1ull);
        cur_item = next_item;
        cur_pair = pm_wait_entry(a, next_item / n_tiles_all);
    }
    __syncthreads();
   }   // items
  }
}

// ---------------------------------------------------------------------------------
// state initialisation / thresholds
// ---------------------------------------------------------------------------------
__global__ void pair_state_init_kernel(PairArgs a) {
    int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= a.n_pairs) return;
    PairState st;
    st.s = 1.0;
    for (int k = 0; k < 9; ++k) st.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    st.t[0] = st.t[1] = st.t[2] = 0.0;
    st.change = 0.0; st.mean_res = 0.0; st.n_valid = 0.0;
    st.iters = 0; st.done = 0; st.status = 0; st.gate_on = 0;
    a.state[pair] = st;
    a.tickets[pair] = 0;
    set_effective(a, pair, st);
    if (pair == 0 && a.n_active) {
        *a.n_active = a.n_pairs;
        for (int q = 0; q < a.max_iterations; ++q) { a.done_count[q] = 0; a.work_counter[q] = 0ull; }
    }
}

__global__ void make_pair_segs_kernel(const da3s_pair* pairs, int n_pairs, long long M, int depth_mode,
                                      float conf_th, float eps, da3s_select_seg* segs) {
    int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    const da3s_pair pr = pairs[pair];
    da3s_select_seg s;
    s.b = nullptr; s.ca = nullptr; s.cb = nullptr; s.n = M; s.kind = DA3S_SEL_VALUES; s.stat = DA3S_SEL_MEDIAN;
    s.percent = 50.0f; s.conf_th = 0.0f; s.eps = 0.0f; s.reserved = 0.0f;
    s.a = pr.conf_a; segs[3 * pair + 0] = s;
    s.a = pr.conf_b; segs[3 * pair + 1] = s;
    // depth-scale median over the single overlap frame prev[-1] / cur[0] (align_geometry.py:319-329):
    // A's LAST overlap frame and B's FIRST overlap frame
    s.kind = DA3S_SEL_RATIO;
    s.n = depth_mode ? M : 0;
    s.a = pr.depth_a; s.b = pr.depth_b; s.ca = pr.conf_a; s.cb = pr.conf_b; s.conf_th = conf_th; s.eps = eps;
    segs[3 * pair + 2] = s;
}

__global__ void pair_prepare_kernel(const da3s_select_out* sel, int n_pairs, int depth_mode, float thr_override,
                                    float* thr, float* dscale, da3s_pair_aux* aux) {
    int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    float ma = sel[3 * pair].value, mb = sel[3 * pair + 1].value;
    // utils/align.py:142: min(median1, median2) * 0.1 in float32 (NumPy >= 2 weak scalar)
    float t = __fmul_rn(fminf(ma, mb), 0.1f);
    if (!isnan(thr_override)) t = thr_override;
    thr[pair] = t;
    float ds = 1.0f;
    if (depth_mode) {
        const da3s_select_out r = sel[3 * pair + 2];
        ds = r.value;
        if (depth_mode == 1) {                               // utils/align_geometry_single.py:42-48
            if (r.n_valid < 50) ds = 1.0f;
            else if (!is_finite_f(ds) || ds <= 0.0f) ds = 1.0f;
        }
    }
    dscale[pair] = ds;
    if (aux) {
        aux[pair].conf_thr = t; aux[pair].depth_scale = ds; aux[pair].median_a = ma; aux[pair].median_b = mb;
        aux[pair].best_hyp = -1.0; aux[pair].best_count = 0.0; aux[pair].mean_residual = 0.0; aux[pair].last_change = 0.0;
    }
}

// ---------------------------------------------------------------------------------
// K4: RANSAC
// ---------------------------------------------------------------------------------
struct RansacArgs {
    const da3s_pair* pairs;
    int n_pairs, overlap, H, W;
    long long P;
    int tiles_per_frame, world, valid_depth, n_hyp, hyp_base;
    float depth_eps, thr2;
    const float* thr; const float* dscale;
    const int32_t* sample_idx;
    float* hyp_A; float* hyp_t; uint8_t* hyp_ok; double* hyp_sim3;
    int32_t* counts;
    unsigned long long* work;       // optional (kernel timers on): += (hypothesis, correspondence) evaluations executed
    int round_mask, n_sel;          // tiles scored by this launch: those whose rs_round_of() bit is set; n_sel of them per frame
    int list_stride;                // hypotheses per pair in the (chyp, cidx) list handed to the scoring kernel
    int split;                      // blocks per tile (a block scores 1 / split of a tile's correspondences); divides PA_GROUPS_PER_BLOCK
    int32_t* kept_tile;             // optional [n_pairs][overlap * tiles_per_frame]: correspondences that pass the mask, per tile
};

__device__ __forceinline__ void frame_const_basic(FrameConst& fc, const da3s_pair& pr, int frame, const float* thr,
                                                  const float* dscale, int pair) {
    const da3s_cam& ca = pr.cam_a[frame];
    const da3s_cam& cb = pr.cam_b[frame];
    fc.cuA = ca.cu; fc.cvA = ca.cv; fc.ifuA = ca.inv_fu; fc.ifvA = ca.inv_fv;
    fc.cuB = cb.cu; fc.cvB = cb.cv; fc.ifuB = cb.inv_fu; fc.ifvB = cb.inv_fv;
    for (int k = 0; k < 12; ++k) { fc.MfA[k] = (float)ca.c2w[k]; fc.MfB[k] = (float)cb.c2w[k]; }
    fc.thr = thr[pair];
    fc.ds = dscale ? dscale[pair] : 1.0f;
    fc.gate_on = 0;
}

__global__ void ransac_hyp_kernel(RansacArgs a) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    int pair = blockIdx.y;
    if (h >= a.n_hyp) return;
    const da3s_pair pr = a.pairs[pair];
    const int32_t* si = a.sample_idx + ((size_t)pair * a.n_hyp + h) * 3;
    const long long M = a.P * a.overlap;
    bool ok = true;
    double X[3][3], Y[3][3];
    long long idx[3];
    for (int j = 0; j < 3; ++j) {
        idx[j] = si[j];
        if (idx[j] < 0 || idx[j] >= M) { ok = false; idx[j] = 0; }
        int frame = (int)(idx[j] / a.P);
        long long pix = idx[j] - (long long)frame * a.P;
        int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
        FrameConst fc;
        frame_const_basic(fc, pr, frame, a.thr, a.dscale, pair);
        float x[3], y[3], xs[3], ys[3], dbs;
        bool keep = corr_points(fc, a.valid_depth, a.depth_eps, u, v, pr.depth_a[idx[j]], pr.conf_a[idx[j]],
                                pr.depth_b[idx[j]], pr.conf_b[idx[j]], x, y, dbs);
        ok = ok && keep;
        ransac_points(fc, a.world, x, y, xs, ys);
        for (int k = 0; k < 3; ++k) { X[j][k] = xs[k]; Y[j][k] = ys[k]; }
    }
    if (idx[0] == idx[1] || idx[0] == idx[2] || idx[1] == idx[2]) ok = false;
    double s = 1.0, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    if (ok) {
        // align_geometry.py:59-82 on three pairs, centred exactly as the reference does
        double mx[3], my[3];
        for (int k = 0; k < 3; ++k) { mx[k] = (X[0][k] + X[1][k] + X[2][k]) / 3.0; my[k] = (Y[0][k] + Y[1][k] + Y[2][k]) / 3.0; }
        double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, var = 0.0;
        for (int j = 0; j < 3; ++j) {
            double xc[3], yc[3];
            for (int k = 0; k < 3; ++k) { xc[k] = X[j][k] - mx[k]; yc[k] = Y[j][k] - my[k]; }
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) cov[3 * i + k] += yc[i] * xc[k];
            var += xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2];
        }
        for (int k = 0; k < 9; ++k) cov[k] /= 3.0;
        var /= 3.0;
        double U[9], Sg[3], V[9];
        svd3(cov, U, Sg, V);
        double dsign = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;
        double Ud[9];
        for (int i = 0; i < 3; ++i) { Ud[3 * i] = U[3 * i]; Ud[3 * i + 1] = U[3 * i + 1]; Ud[3 * i + 2] = U[3 * i + 2] * dsign; }
        mat3_mul_bt(Ud, V, R);
        s = (Sg[0] + Sg[1] + dsign * Sg[2]) / (var + 1e-12);
        double Rm[3];
        mat3_vec(R, mx, Rm);
        for (int k = 0; k < 3; ++k) t[k] = my[k] - s * Rm[k];
        bool fin = fabs(s) < INFINITY;
        for (int k = 0; k < 9; ++k) fin = fin && (fabs(R[k]) < INFINITY);
        for (int k = 0; k < 3; ++k) fin = fin && (fabs(t[k]) < INFINITY);
        ok = fin;
    }
    size_t o = (size_t)pair * a.n_hyp + h;
    for (int k = 0; k < 9; ++k) a.hyp_A[o * 9 + k] = ok ? (float)(s * R[k]) : 0.0f;
    for (int k = 0; k < 3; ++k) a.hyp_t[o * 3 + k] = ok ? (float)t[k] : 0.0f;
    a.hyp_ok[o] = ok ? 1 : 0;
    if (a.hyp_sim3) {
        double* r = a.hyp_sim3 + o * 13;
        r[0] = ok ? s : 0.0;
        for (int k = 0; k < 9; ++k) r[1 + k] = ok ? R[k] : 0.0;
        for (int k = 0; k < 3; ++k) r[10 + k] = ok ? t[k] : 0.0;
    }
}

#ifndef RS_THREADS
#define RS_THREADS 64                       // small blocks: the hypothesis range is cut to the VALID hypotheses with 128-wide granularity
#endif
#define RS_SUB 512                          // correspondences staged per shared-memory sub-tile
#ifndef RS_HPT
#define RS_HPT 4                            // hypotheses per thread (even: two share every packed instruction)
#endif
#define RS_HYP_PER_BLOCK (RS_THREADS * RS_HPT)
#ifndef RS_SPLIT
#define RS_SPLIT 8                          // blocks per tile in the round launches (PA_GROUPS_PER_BLOCK * 4 / RS_SPLIT correspondences each)
#endif

// Scoring in ROUNDS with exact pruning (da3s_align_pairs needs the winner, not every count).  The tiles of a frame are
// dealt into RS_ROUNDS = 3 rounds, interleaved over the image: round 0 takes the middle tile and 5 of every 16 others
// (35 % of a 518 x 518 frame), round 1 another 4 of 16 (59 % seen), round 2 the rest.  After round 0 the hypothesis with
// the most inliers so far (the LEADER) is scored on all other tiles at once, which also counts the correspondences every
// tile keeps.  Its complete count L bounds the winner from below, so after each round a hypothesis whose count so far
// plus ALL correspondences of the tiles it has not seen yet stays below L can neither win nor tie and is dropped from the
// list (ransac_prune_kernel); the survivors' counts are complete at the end and the winner (most inliers, ties to the
// lowest index) is the one full scoring finds.  With a leader at 70 % inliers, round 0 drops everything below 14 % of
// the leader's count; round 1 catches up to 32 % when the leader itself is weaker.  Few, large launches: every launch
// pays a ramp (coefficients, first staging) and a tail, which is why the rounds are not finer.
#define RS_ROUNDS 3
__host__ __device__ __forceinline__ int rs_round_of(int tile, int tiles_per_frame) {
    if (tile == tiles_per_frame / 2) return 0;
    return (int)((0x2102201201202102ull >> (4 * (tile & 15))) & 15ull);
}
static_assert(RS_ROUNDS == DA3S_RANSAC_ROUNDS && PA_GROUPS_PER_BLOCK * 4 == DA3S_RANSAC_TILE, "include/da3s.h documents the rounds");
extern "C" int da3s_ransac_round_of(int tile, int tiles_per_frame) {
    if (tile < 0 || tiles_per_frame <= 0 || tile >= tiles_per_frame) return DA3S_EINVAL;
    return rs_round_of(tile, tiles_per_frame);
}
static int rs_round_tiles(int tiles_per_frame, int mask) {
    int n = 0;
    for (int t = 0; t < tiles_per_frame; ++t) n += (mask >> rs_round_of(t, tiles_per_frame)) & 1;
    return n;
}

// ordered compaction of the entries i of a list for which keep(i) holds (block per pair, 256 threads)
template <typename Keep>
__device__ __forceinline__ void rs_filter_list(const float* __restrict__ chyp_in, const int32_t* __restrict__ cidx_in, int n_in, int stride,
                                               int pair, float* __restrict__ chyp_out, int32_t* __restrict__ cidx_out,
                                               int32_t* __restrict__ n_out, int* wsum, int* base_sh, Keep keep) {
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *base_sh = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_in; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const int h = i < n_in ? cidx_in[(size_t)pair * stride + i] : -1;
        const bool ok = h >= 0 && keep(h);
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int before = *base_sh;
        for (unsigned int w = 0; w < warp; ++w) before += wsum[w];
        if (ok) {
            const int pos = before + __popc(m & ((1u << lane) - 1u));
            const float* src = chyp_in + ((size_t)pair * stride + i) * 12;
            float* dst = chyp_out + ((size_t)pair * stride + pos) * 12;
#pragma unroll
            for (int k = 0; k < 12; ++k) dst[k] = src[k];
            cidx_out[(size_t)pair * stride + pos] = h;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wsum[w]; *base_sh += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) n_out[pair] = *base_sh;
}

// leader after round 0 (most inliers so far, ties to the lowest index) as a one-entry list, and the list without it
__global__ void __launch_bounds__(256)
ransac_lead_kernel(const int32_t* __restrict__ counts, int n_hyp, const float* __restrict__ chyp_in, const int32_t* __restrict__ cidx_in,
                   const int32_t* __restrict__ n_in, float* __restrict__ chyp_lead, int32_t* __restrict__ cidx_lead, int32_t* __restrict__ n_lead,
                   float* __restrict__ chyp_out, int32_t* __restrict__ cidx_out, int32_t* __restrict__ n_out) {
    __shared__ unsigned long long best[256];
    __shared__ int wsum[8];
    __shared__ int base_sh;
    const int pair = blockIdx.x, n = n_in[pair];
    unsigned long long b = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int h = cidx_in[(size_t)pair * n_hyp + i];
        const unsigned long long key = ((unsigned long long)(unsigned int)(counts[(size_t)pair * n_hyp + h] + 1) << 32) | (unsigned int)(0x7fffffff - h);
        if (key > b) b = key;
    }
    best[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s && best[threadIdx.x + s] > best[threadIdx.x]) best[threadIdx.x] = best[threadIdx.x + s];
        __syncthreads();
    }
    const int lead = best[0] ? (int)(0x7fffffff - (unsigned int)(best[0] & 0xffffffffu)) : -1;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256)
        if (cidx_in[(size_t)pair * n_hyp + i] == lead) {
            for (int k = 0; k < 12; ++k) chyp_lead[(size_t)pair * 12 + k] = chyp_in[((size_t)pair * n_hyp + i) * 12 + k];
        }
    if (threadIdx.x == 0) { cidx_lead[pair] = lead; n_lead[pair] = lead >= 0 ? 1 : 0; }
    rs_filter_list(chyp_in, cidx_in, n, n_hyp, pair, chyp_out, cidx_out, n_out, wsum, &base_sh, [&](int h) { return h != lead; });
}

// after round `after_round`: keep the hypotheses that can still reach the leader's complete count
__global__ void __launch_bounds__(256)
ransac_prune_kernel(const int32_t* __restrict__ counts, int n_hyp, const int32_t* __restrict__ cidx_lead, const int32_t* __restrict__ kept_tile,
                    int overlap, int tiles_per_frame, int after_round, const float* __restrict__ chyp_in, const int32_t* __restrict__ cidx_in,
                    const int32_t* __restrict__ n_in, float* __restrict__ chyp_out, int32_t* __restrict__ cidx_out, int32_t* __restrict__ n_out) {
    __shared__ int wsum[8];
    __shared__ int base_sh;
    __shared__ long long rem_sh;
    const int pair = blockIdx.x;
    if (threadIdx.x == 0) rem_sh = 0;
    __syncthreads();
    long long rem = 0;                      // correspondences of the tiles the list has not been scored on yet
    for (int i = threadIdx.x; i < overlap * tiles_per_frame; i += 256)
        if (rs_round_of(i % tiles_per_frame, tiles_per_frame) > after_round) rem += kept_tile[(size_t)pair * overlap * tiles_per_frame + i];
    for (int o = 16; o > 0; o >>= 1) rem += __shfl_down_sync(0xffffffffu, rem, o);
    if ((threadIdx.x & 31) == 0 && rem) atomicAdd((unsigned long long*)&rem_sh, (unsigned long long)rem);
    __syncthreads();
    const long long remaining = rem_sh;
    const int lead = cidx_lead[pair];
    const long long target = lead >= 0 ? (long long)counts[(size_t)pair * n_hyp + lead] : 0;
    rs_filter_list(chyp_in, cidx_in, n_in[pair], n_hyp, pair, chyp_out, cidx_out, n_out, wsum, &base_sh,
                   [&](int h) { return (long long)counts[(size_t)pair * n_hyp + h] + remaining >= target; });
}

// Valid hypotheses of every pair, compacted (stable): coefficient rows [n_pairs][n_hyp][12] (A row-major, then t), their
// original indices, and the count.  About a third of uniformly drawn 3-pixel samples touch a masked pixel (SPEC 4: such a
// hypothesis scores 0), and scoring them costs exactly what a valid one costs — so the scoring kernel never sees them.
__global__ void __launch_bounds__(256)
ransac_compact_kernel(const float* __restrict__ hyp_A, const float* __restrict__ hyp_t, const uint8_t* __restrict__ hyp_ok, int n_hyp,
                      float* __restrict__ chyp, int32_t* __restrict__ cidx, int32_t* __restrict__ n_valid) {
    __shared__ int wsum[8];
    __shared__ int base_sh;
    const int pair = blockIdx.x;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_sh = 0;
    __syncthreads();
    for (int h0 = 0; h0 < n_hyp; h0 += 256) {
        const int h = h0 + threadIdx.x;
        const bool ok = h < n_hyp && hyp_ok[(size_t)pair * n_hyp + h];
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int before = base_sh;
        for (unsigned int w = 0; w < warp; ++w) before += wsum[w];
        if (ok) {
            const int pos = before + __popc(m & ((1u << lane) - 1u));
            const size_t o = (size_t)pair * n_hyp + h, d = ((size_t)pair * n_hyp + pos) * 12;
#pragma unroll
            for (int k = 0; k < 9; ++k) chyp[d + k] = hyp_A[o * 9 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) chyp[d + 9 + k] = hyp_t[o * 3 + k];
            cidx[(size_t)pair * n_hyp + pos] = h;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wsum[w]; base_sh += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) n_valid[pair] = base_sh;
}

// 1.0f / 0.0f for a < b in ONE instruction (FSET.BF): inlier counts are accumulated as packed float32 (exact: a block
// counts at most 16384 correspondences), which replaces compare + predicated integer add per hypothesis by 1.5 instructions
__device__ __forceinline__ float set_lt(float a, float b) { float r; asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }


// Threads own hypotheses (coefficients in registers), points are broadcast from shared
// memory.  Two hypotheses share every arithmetic instruction: sm_100's packed float32 pipe
// (FFMA2 / FADD2 / FMUL2 on register pairs, each half an ordinary IEEE operation, so SPEC 4's
// fma chain is reproduced bit for bit) evaluates the residual of a point under a PAIR of
// hypotheses in 15 instructions; the point is stored duplicated (v, v) so that one LDS.128
// delivers two broadcast operands.  Only correspondences that passed the joint mask are staged (warp-aggregated
// compaction into the sub-tile), only valid hypotheses are scored (ransac_compact_kernel), blockIdx.z selects a slice of
// RS_HYP_PER_BLOCK of them and warps beyond the last valid hypothesis leave at once.
__global__ void __launch_bounds__(RS_THREADS)
ransac_score_kernel(RansacArgs a, const float* __restrict__ chyp, const int32_t* __restrict__ cidx, const int32_t* __restrict__ n_valid) {
    __shared__ FrameConst fc;
    __shared__ float4 pts[RS_SUB][3];       // (x0,x0,x1,x1) (x2,x2,-y0,-y0) (-y1,-y1,-y2,-y2)
    __shared__ int n_sh;
    const int pair = blockIdx.y;
    const int nv = n_valid[pair];
    const int hyp0 = blockIdx.z * RS_HYP_PER_BLOCK;
    if (hyp0 >= nv) return;                 // block-uniform
    const da3s_pair pr = a.pairs[pair];
    __shared__ int tile_sh;
    const int bx = blockIdx.x / a.split, part = blockIdx.x - bx * a.split;
    const int frame = bx / a.n_sel;
    if (threadIdx.x == 0) {
        frame_const_basic(fc, pr, frame, a.thr, a.dscale, pair); n_sh = 0;
        const int j = bx - frame * a.n_sel;             // the j-th tile of this launch's rounds
        int t = j;
        if (a.n_sel != a.tiles_per_frame) {
            int c = 0;
            for (t = 0; t < a.tiles_per_frame; ++t)
                if ((a.round_mask >> rs_round_of(t, a.tiles_per_frame)) & 1) { if (c == j) break; ++c; }
        }
        tile_sh = t;
    }
    __syncthreads();
    const int tile = tile_sh;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this warp's hypotheses: lane-interleaved inside the warp so that a warp covers 32 * RS_HPT consecutive valid ones
    const int wbase = hyp0 + (int)warp * 32 * RS_HPT;
    const bool warp_live = wbase < nv;      // warp-uniform; dead warps still help staging

    float2 A[RS_HPT / 2][12];               // .x: hypothesis 2q, .y: hypothesis 2q + 1
    float2 cnt[RS_HPT / 2];
    int hid[RS_HPT];
#pragma unroll
    for (int m = 0; m < RS_HPT; ++m) {
        const int j = wbase + m * 32 + (int)lane;
        hid[m] = j < nv ? j : -1;
        const float* hc = chyp + ((size_t)pair * a.list_stride + (j < nv ? j : 0)) * 12;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float c = warp_live ? hc[k] : 0.0f;
            if (m & 1) A[m >> 1][k].y = c; else A[m >> 1][k].x = c;
        }
    }
#pragma unroll
    for (int q = 0; q < RS_HPT / 2; ++q) cnt[q] = make_float2(0.0f, 0.0f);

    const size_t foff = (size_t)frame * (size_t)a.P;
    const long long part_len = (long long)(PA_GROUPS_PER_BLOCK * 4) / a.split;
    const long long p_begin = (long long)tile * PA_GROUPS_PER_BLOCK * 4 + part * part_len;
    long long p_end = p_begin + part_len;
    if (p_end > a.P) p_end = a.P;
    const float2 thr2 = make_float2(a.thr2, a.thr2);
    int n_staged = 0;
    for (long long sb = p_begin; sb < p_end; sb += RS_SUB) {
        // stage the KEPT correspondences of this sub-tile (RS_SUB / RS_THREADS per thread), compacted
#pragma unroll 2
        for (int j = 0; j < RS_SUB / RS_THREADS; ++j) {
            const long long pix = sb + j * RS_THREADS + threadIdx.x;
            float xs[3] = {0, 0, 0}, ys[3] = {0, 0, 0};
            bool keep = false;
            if (pix < p_end) {
                int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
                float x[3], y[3], dbs;
                keep = corr_points(fc, a.valid_depth, a.depth_eps, u, v,
                                   ldg_stream1(pr.depth_a + foff + pix), ldg_stream1(pr.conf_a + foff + pix),
                                   ldg_stream1(pr.depth_b + foff + pix), ldg_stream1(pr.conf_b + foff + pix), x, y, dbs);
                if (keep) ransac_points(fc, a.world, x, y, xs, ys);
            }
            const unsigned int m = __ballot_sync(0xffffffffu, keep);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&n_sh, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                pts[slot][0] = make_float4(xs[0], xs[0], xs[1], xs[1]);
                pts[slot][1] = make_float4(xs[2], xs[2], -ys[0], -ys[0]);      // d = p - y  ==  p + (-y), exactly
                pts[slot][2] = make_float4(-ys[1], -ys[1], -ys[2], -ys[2]);
            }
        }
        __syncthreads();
        const int n_here = n_sh;
        n_staged += n_here;
        if (warp_live) {
#pragma unroll 8
            for (int i = 0; i < n_here; ++i) {
                const float4 p0 = pts[i][0], p1 = pts[i][1], p2 = pts[i][2];
                const float2 X0 = make_float2(p0.x, p0.y), X1 = make_float2(p0.z, p0.w), X2 = make_float2(p1.x, p1.y);
                const float2 N0 = make_float2(p1.z, p1.w), N1 = make_float2(p2.x, p2.y), N2 = make_float2(p2.z, p2.w);
#pragma unroll
                for (int q = 0; q < RS_HPT / 2; ++q) {
                    // residual2_f32 (SPEC 4) for two hypotheses at once
                    const float2 d0 = __fadd2_rn(__ffma2_rn(A[q][0], X0, __ffma2_rn(A[q][1], X1, __ffma2_rn(A[q][2], X2, A[q][9]))), N0);
                    const float2 d1 = __fadd2_rn(__ffma2_rn(A[q][3], X0, __ffma2_rn(A[q][4], X1, __ffma2_rn(A[q][5], X2, A[q][10]))), N1);
                    const float2 d2 = __fadd2_rn(__ffma2_rn(A[q][6], X0, __ffma2_rn(A[q][7], X1, __ffma2_rn(A[q][8], X2, A[q][11]))), N2);
                    const float2 r2 = __ffma2_rn(d0, d0, __ffma2_rn(d1, d1, __fmul2_rn(d2, d2)));
                    cnt[q] = __fadd2_rn(cnt[q], make_float2(set_lt(r2.x, thr2.x), set_lt(r2.y, thr2.y)));
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) n_sh = 0;
        __syncthreads();
    }
    if (!warp_live) return;
    if (threadIdx.x == 0) {
        if (a.work) atomicAdd(a.work, (unsigned long long)n_staged * (unsigned long long)min(RS_HYP_PER_BLOCK, nv - hyp0));
        if (a.kept_tile && blockIdx.z == 0 && n_staged)
            atomicAdd(&a.kept_tile[((size_t)pair * a.overlap + frame) * a.tiles_per_frame + tile], n_staged);
    }
#pragma unroll
    for (int m = 0; m < RS_HPT; ++m) {
        const int c = (int)((m & 1) ? cnt[m >> 1].y : cnt[m >> 1].x);
        if (hid[m] >= 0 && c) atomicAdd(&a.counts[(size_t)pair * a.n_hyp + cidx[(size_t)pair * a.list_stride + hid[m]]], c);
    }
}

// winner per pair: max count, ties to the lowest index; fewer than min_inliers -> no model
__global__ void __launch_bounds__(256)
ransac_best_kernel(const int32_t* counts, const uint8_t* hyp_ok, const float* hyp_A, const float* hyp_t, int n_hyp,
                   int min_inliers, PairState* state, float* gate, double* rows, da3s_pair_aux* aux, int* n_active) {
    __shared__ unsigned long long best[256];
    const int pair = blockIdx.x;
    unsigned long long b = 0;
    for (int h = threadIdx.x; h < n_hyp; h += 256) {
        size_t o = (size_t)pair * n_hyp + h;
        if (!hyp_ok[o]) continue;
        unsigned long long key = ((unsigned long long)(unsigned int)(counts[o] + 1) << 32) | (unsigned int)(0x7fffffff - h);
        if (key > b) b = key;
    }
    best[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s && best[threadIdx.x + s] > best[threadIdx.x]) best[threadIdx.x] = best[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long k = best[0];
        int cnt = (int)(k >> 32) - 1;
        int h = k ? (int)(0x7fffffff - (unsigned int)(k & 0xffffffffu)) : -1;
        PairState st = state[pair];
        if (h < 0 || cnt < min_inliers) {                    // align_geometry.py:124
            st.done = 1; st.status = 2; st.gate_on = 0; st.n_valid = 0.0;
            state[pair] = st;
            if (n_active) atomicSub(n_active, 1);
            double* r = rows + (size_t)pair * DA3S_ROW_LEN;
            r[DA3S_ROW_S] = 1.0;
            for (int q = 0; q < 9; ++q) r[DA3S_ROW_R + q] = (q % 4 == 0) ? 1.0 : 0.0;
            for (int q = 0; q < 3; ++q) r[DA3S_ROW_T + q] = 0.0;
            r[DA3S_ROW_NVALID] = 0.0; r[DA3S_ROW_ITERS] = 0.0; r[DA3S_ROW_STATUS] = 2.0;
            if (aux) { aux[pair].best_hyp = -1.0; aux[pair].best_count = cnt < 0 ? 0.0 : (double)cnt; }
        } else {
            st.gate_on = 1;
            state[pair] = st;
            size_t o = (size_t)pair * n_hyp + h;
            for (int q = 0; q < 9; ++q) gate[12 * (size_t)pair + q] = hyp_A[o * 9 + q];
            for (int q = 0; q < 3; ++q) gate[12 * (size_t)pair + 9 + q] = hyp_t[o * 3 + q];
            if (aux) { aux[pair].best_hyp = (double)h; aux[pair].best_count = (double)cnt; }
        }
    }
}

__global__ void __launch_bounds__(256)
ransac_mask_kernel(RansacArgs a, const float* best_A, const float* best_t, const uint8_t* best_ok, uint8_t* mask_out) {
    __shared__ FrameConst fc;
    __shared__ float g[12];
    const int pair = blockIdx.y;
    const da3s_pair pr = a.pairs[pair];
    const int frame = blockIdx.x / a.tiles_per_frame;
    const int tile = blockIdx.x - frame * a.tiles_per_frame;
    if (threadIdx.x == 0) {
        frame_const_basic(fc, pr, frame, a.thr, a.dscale, pair);
        for (int k = 0; k < 9; ++k) g[k] = best_A[9 * (size_t)pair + k];
        for (int k = 0; k < 3; ++k) g[9 + k] = best_t[3 * (size_t)pair + k];
    }
    __syncthreads();
    const bool ok = best_ok[pair];
    const size_t foff = (size_t)frame * (size_t)a.P;
    const long long p_begin = (long long)tile * PA_GROUPS_PER_BLOCK * 4;
    long long p_end = p_begin + (long long)PA_GROUPS_PER_BLOCK * 4;
    if (p_end > a.P) p_end = a.P;
    for (long long pix = p_begin + threadIdx.x; pix < p_end; pix += 256) {
        int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
        float x[3], y[3], xs[3], ys[3], dbs;
        bool keep = corr_points(fc, a.valid_depth, a.depth_eps, u, v, pr.depth_a[foff + pix], pr.conf_a[foff + pix],
                                pr.depth_b[foff + pix], pr.conf_b[foff + pix], x, y, dbs);
        if (keep && ok) {
            ransac_points(fc, a.world, x, y, xs, ys);
            keep = residual2_f32(g, xs, ys) < a.thr2;
        }
        mask_out[(size_t)pair * a.P * a.overlap + foff + pix] = (keep && ok) ? 1 : 0;
    }
}

// The leader on the tiles of the rounds in a.round_mask, point-parallel (threads own correspondences, the one hypothesis is
// block-uniform): its inlier count (same float32 chain as the scoring kernel: residual2_f32) and, per tile, the number of
// correspondences that pass the mask.  A scoring-kernel warp evaluates 128 hypothesis slots whatever the list length, so a
// one-entry list would cost a fifth of full scoring; this pass is a plain 16 B / correspondence stream.
__global__ void __launch_bounds__(256)
ransac_leader_kernel(RansacArgs a, const float* __restrict__ chyp_lead, const int32_t* __restrict__ cidx_lead) {
    __shared__ FrameConst fc;
    __shared__ float g[12];
    __shared__ int tile_sh;
    __shared__ int red[2][8];
    const int pair = blockIdx.y;
    const int lead = cidx_lead[pair];
    if (lead < 0) return;                   // block-uniform
    const da3s_pair pr = a.pairs[pair];
    const int frame = blockIdx.x / a.n_sel;
    if (threadIdx.x == 0) {
        frame_const_basic(fc, pr, frame, a.thr, a.dscale, pair);
        for (int k = 0; k < 12; ++k) g[k] = chyp_lead[(size_t)pair * 12 + k];
        const int j = blockIdx.x - frame * a.n_sel;
        int c = 0, t;
        for (t = 0; t < a.tiles_per_frame; ++t)
            if ((a.round_mask >> rs_round_of(t, a.tiles_per_frame)) & 1) { if (c == j) break; ++c; }
        tile_sh = t;
    }
    __syncthreads();
    const int tile = tile_sh;
    const size_t foff = (size_t)frame * (size_t)a.P;
    const long long p_begin = (long long)tile * PA_GROUPS_PER_BLOCK * 4;
    long long p_end = p_begin + (long long)PA_GROUPS_PER_BLOCK * 4;
    if (p_end > a.P) p_end = a.P;
    int kept = 0, inl = 0;
    for (long long pix = p_begin + threadIdx.x; pix < p_end; pix += 256) {
        const int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
        float x[3], y[3], xs[3], ys[3], dbs;
        const bool keep = corr_points(fc, a.valid_depth, a.depth_eps, u, v, ldg_stream1(pr.depth_a + foff + pix), ldg_stream1(pr.conf_a + foff + pix),
                                      ldg_stream1(pr.depth_b + foff + pix), ldg_stream1(pr.conf_b + foff + pix), x, y, dbs);
        if (keep) {
            ransac_points(fc, a.world, x, y, xs, ys);
            ++kept;
            inl += residual2_f32(g, xs, ys) < a.thr2 ? 1 : 0;
        }
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    inl = __reduce_add_sync(0xffffffffu, inl);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = kept; red[1][threadIdx.x >> 5] = inl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0, n = 0;
        for (int w = 0; w < 8; ++w) { k += red[0][w]; n += red[1][w]; }
        if (a.kept_tile) a.kept_tile[((size_t)pair * a.overlap + frame) * a.tiles_per_frame + tile] = k;
        if (n) atomicAdd(&a.counts[(size_t)pair * a.n_hyp + lead], n);
    }
}

// ---------------------------------------------------------------------------------
// host-side orchestration
// ---------------------------------------------------------------------------------
static bool pair_geometry(int overlap, int H, int W, long long& P, int& tiles_per_frame) {
    if (overlap <= 0 || H <= 0 || W <= 0) return false;
    P = (long long)H * W;
    long long groups = (P + 3) / 4;
    tiles_per_frame = (int)((groups + PA_GROUPS_PER_BLOCK - 1) / PA_GROUPS_PER_BLOCK);
    return (long long)tiles_per_frame * overlap <= 2147483647LL;
}

extern "C" void da3s_align_opts_default(da3s_align_opts* o) {
    if (!o) return;
    o->world = 1; o->depth_scale_mode = 0; o->depth_conf_th = 0.2f; o->depth_eps = 1e-6f; o->valid_depth = 1;
    o->conf_thr_override = nanf(""); o->huber = 1; o->huber_delta = 1.0; o->max_iterations = 20; o->tol = 1e-6;
    o->min_points = 100; o->n_hyp = 0; o->ransac_thr = 0.05f; o->ransac_min_inliers = 20; o->precise = 0;
}

static int thresholds_impl(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                           const da3s_align_opts* opts, float* thr, float* dscale, da3s_pair_aux* aux, cudaStream_t st) {
    long long P; int tpf;
    if (!pair_geometry(overlap, H, W, P, tpf)) return DA3S_EINVAL;
    size_t save_top = ctx->ws_top;
    WS_ALLOC(ctx, da3s_select_seg, segs, 3 * (size_t)n_pairs);
    WS_ALLOC(ctx, da3s_select_out, sel, 3 * (size_t)n_pairs);
    int threads = 128, blocks = (n_pairs + threads - 1) / threads;
    // the confidence medians run over ALL overlap frames (utils/align.py:136-141); the depth
    // ratio over ONE frame: A's last overlap frame vs B's first (align_geometry.py:319-320)
    make_pair_segs_kernel<<<blocks, threads, 0, st>>>(pairs, n_pairs, P * overlap, opts->depth_scale_mode,
                                                      opts->depth_conf_th, opts->depth_eps, segs);
    DA3S_LAUNCH_CHECK(ctx);
    int rc = da3s_select_impl(ctx, segs, 3 * n_pairs, P * overlap, sel, st);
    if (rc != DA3S_OK) return rc;
    pair_prepare_kernel<<<blocks, threads, 0, st>>>(sel, n_pairs, opts->depth_scale_mode, opts->conf_thr_override, thr, dscale, aux);
    DA3S_LAUNCH_CHECK(ctx);
    ctx->ws_top = save_top;
    return DA3S_OK;
}

extern "C" int da3s_pair_thresholds(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                                    const da3s_align_opts* opts, float* conf_thr, float* depth_scale,
                                    da3s_pair_aux* aux, void* stream) {
    if (!ctx || !pairs || !opts || !conf_thr || !depth_scale || n_pairs <= 0) return DA3S_EINVAL;
    if (opts->depth_scale_mode && overlap != 1) return DA3S_EINVAL;
    ws_reset(ctx);
    return thresholds_impl(ctx, pairs, n_pairs, overlap, H, W, opts, conf_thr, depth_scale, aux, (cudaStream_t)stream);
}

static void fill_ransac_args(RansacArgs& r, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W, long long P,
                             int tpf, int world, int valid_depth, float depth_eps, const float* thr, const float* dscale,
                             int n_hyp, float ransac_thr) {
    r.pairs = pairs; r.n_pairs = n_pairs; r.overlap = overlap; r.H = H; r.W = W; r.P = P; r.tiles_per_frame = tpf;
    r.world = world; r.valid_depth = valid_depth; r.n_hyp = n_hyp; r.hyp_base = 0; r.depth_eps = depth_eps;
    r.thr2 = (float)((double)ransac_thr * (double)ransac_thr);
    r.thr = thr; r.dscale = dscale; r.sample_idx = nullptr; r.hyp_A = nullptr; r.hyp_t = nullptr; r.hyp_ok = nullptr;
    r.hyp_sim3 = nullptr; r.counts = nullptr; r.work = nullptr;
    r.round_mask = (1 << RS_ROUNDS) - 1; r.n_sel = tpf; r.list_stride = n_hyp; r.split = 1; r.kept_tile = nullptr;
}

extern "C" int da3s_ransac_hypotheses(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                                      int world, int valid_depth, float depth_eps, const float* conf_thr, const float* depth_scale,
                                      const int32_t* sample_idx, int n_hyp, float* hyp_A, float* hyp_t, uint8_t* hyp_ok,
                                      double* hyp_sim3, void* stream) {
    if (!ctx || !pairs || !conf_thr || !sample_idx || !hyp_A || !hyp_t || !hyp_ok || n_pairs <= 0 || n_hyp <= 0) return DA3S_EINVAL;
    if (n_pairs > 65535) return DA3S_EINVAL;
    long long P; int tpf;
    if (!pair_geometry(overlap, H, W, P, tpf)) return DA3S_EINVAL;
    RansacArgs r;
    fill_ransac_args(r, pairs, n_pairs, overlap, H, W, P, tpf, world, valid_depth, depth_eps, conf_thr, depth_scale, n_hyp, 0.0f);
    r.sample_idx = sample_idx; r.hyp_A = hyp_A; r.hyp_t = hyp_t; r.hyp_ok = hyp_ok; r.hyp_sim3 = hyp_sim3;
    dim3 grid((n_hyp + 127) / 128, n_pairs);
    ransac_hyp_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(r);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_ransac_score(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                                 int world, int valid_depth, float depth_eps, const float* conf_thr, const float* depth_scale,
                                 const float* hyp_A, const float* hyp_t, const uint8_t* hyp_ok, int n_hyp,
                                 float ransac_thr, int32_t* counts, void* stream) {
    if (!ctx || !pairs || !conf_thr || !hyp_A || !hyp_t || !hyp_ok || !counts || n_pairs <= 0 || n_hyp <= 0) return DA3S_EINVAL;
    if (n_pairs > 65535) return DA3S_EINVAL;
    long long P; int tpf;
    if (!pair_geometry(overlap, H, W, P, tpf)) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    RansacArgs r;
    fill_ransac_args(r, pairs, n_pairs, overlap, H, W, P, tpf, world, valid_depth, depth_eps, conf_thr, depth_scale, n_hyp, ransac_thr);
    r.hyp_A = const_cast<float*>(hyp_A); r.hyp_t = const_cast<float*>(hyp_t); r.hyp_ok = const_cast<uint8_t*>(hyp_ok); r.counts = counts;
    r.work = ctx->prof_on ? ctx->prof_work + DA3S_TIMED_RANSAC_SCORE : nullptr;
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_pairs * n_hyp, st));
    size_t save_top = ctx->ws_top;
    WS_ALLOC(ctx, float, chyp, (size_t)n_pairs * n_hyp * 12);
    WS_ALLOC(ctx, int32_t, cidx, (size_t)n_pairs * n_hyp);
    WS_ALLOC(ctx, int32_t, n_valid, n_pairs);
    ransac_compact_kernel<<<n_pairs, 256, 0, st>>>(hyp_A, hyp_t, hyp_ok, n_hyp, chyp, cidx, n_valid);
    DA3S_LAUNCH_CHECK(ctx);
    const int zs = (n_hyp + RS_HYP_PER_BLOCK - 1) / RS_HYP_PER_BLOCK;
    if (zs > 65535) return DA3S_EINVAL;
    dim3 grid(tpf * overlap, n_pairs, zs);
    prof_begin(ctx, DA3S_TIMED_RANSAC_SCORE, st);
    ransac_score_kernel<<<grid, RS_THREADS, 0, st>>>(r, chyp, cidx, n_valid);
    prof_end(ctx, DA3S_TIMED_RANSAC_SCORE, st);
    DA3S_LAUNCH_CHECK(ctx);
    ctx->ws_top = save_top;     // consumed in stream order
    return DA3S_OK;
}

// Scoring for da3s_align_pairs: rounds with exact pruning (comment above rs_round_of).  Same winner as da3s_ransac_score;
// the counts of dropped hypotheses stay partial (below the winner's), so this path is not used when the caller asks for
// the count table.
static int ransac_score_rounds(da3s_ctx* ctx, RansacArgs r, const float* hyp_A, const float* hyp_t, const uint8_t* hyp_ok,
                               int32_t* counts, cudaStream_t st) {
    const int n_pairs = r.n_pairs, n_hyp = r.n_hyp, tpf = r.tiles_per_frame, overlap = r.overlap;
    r.counts = counts;
    r.work = ctx->prof_on ? ctx->prof_work + DA3S_TIMED_RANSAC_SCORE : nullptr;
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_pairs * n_hyp, st));
    size_t save_top = ctx->ws_top;
    float* chyp[2]; int32_t* cidx[2]; int32_t* n_list[2];
    for (int k = 0; k < 2; ++k) {
        WS_ALLOC(ctx, float, c, (size_t)n_pairs * n_hyp * 12);
        WS_ALLOC(ctx, int32_t, i, (size_t)n_pairs * n_hyp);
        WS_ALLOC(ctx, int32_t, n, n_pairs);
        chyp[k] = c; cidx[k] = i; n_list[k] = n;
    }
    WS_ALLOC(ctx, float, chyp_lead, (size_t)n_pairs * 12);
    WS_ALLOC(ctx, int32_t, cidx_lead, n_pairs);
    WS_ALLOC(ctx, int32_t, n_lead, n_pairs);
    WS_ALLOC(ctx, int32_t, kept_tile, (size_t)n_pairs * overlap * tpf);
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(kept_tile, 0, sizeof(int32_t) * (size_t)n_pairs * overlap * tpf, st));
    const int zs = (n_hyp + RS_HYP_PER_BLOCK - 1) / RS_HYP_PER_BLOCK;
    if (zs > 65535) return DA3S_EINVAL;
    // a round has few tiles: every tile is cut into RS_SPLIT blocks so that one launch still fills the device
    auto score = [&](int mask, const float* ch, const int32_t* ci, const int32_t* nl, int stride, int z, bool count_kept) -> int {
        RansacArgs q = r;
        q.round_mask = mask; q.n_sel = rs_round_tiles(tpf, mask); q.list_stride = stride; q.split = RS_SPLIT;
        q.kept_tile = count_kept ? kept_tile : nullptr;
        if (q.n_sel == 0) return DA3S_OK;
        prof_begin(ctx, DA3S_TIMED_RANSAC_SCORE, st);
        ransac_score_kernel<<<dim3(q.n_sel * overlap * RS_SPLIT, n_pairs, z), RS_THREADS, 0, st>>>(q, ch, ci, nl);
        prof_end(ctx, DA3S_TIMED_RANSAC_SCORE, st);
        DA3S_LAUNCH_CHECK(ctx);
        return DA3S_OK;
    };
    ransac_compact_kernel<<<n_pairs, 256, 0, st>>>(hyp_A, hyp_t, hyp_ok, n_hyp, chyp[0], cidx[0], n_list[0]);
    DA3S_LAUNCH_CHECK(ctx);
    const int all_rounds = (1 << RS_ROUNDS) - 1;
    int rc = score(1 << 0, chyp[0], cidx[0], n_list[0], n_hyp, zs, true);                   // round 0, every valid hypothesis
    if (rc != DA3S_OK) return rc;
    ransac_lead_kernel<<<n_pairs, 256, 0, st>>>(counts, n_hyp, chyp[0], cidx[0], n_list[0], chyp_lead, cidx_lead, n_lead,
                                                chyp[1], cidx[1], n_list[1]);
    DA3S_LAUNCH_CHECK(ctx);
    {                                                                                       // the leader on every other tile
        RansacArgs q = r;
        q.round_mask = all_rounds & ~1; q.n_sel = rs_round_tiles(tpf, q.round_mask); q.kept_tile = kept_tile;
        if (q.n_sel > 0) {
            ransac_leader_kernel<<<dim3(q.n_sel * overlap, n_pairs), 256, 0, st>>>(q, chyp_lead, cidx_lead);
            DA3S_LAUNCH_CHECK(ctx);
        }
    }
    int cur = 1;
    for (int round = 1; round < RS_ROUNDS; ++round) {
        if (rs_round_tiles(tpf, 1 << round) == 0) continue;
        ransac_prune_kernel<<<n_pairs, 256, 0, st>>>(counts, n_hyp, cidx_lead, kept_tile, overlap, tpf, round - 1, chyp[cur], cidx[cur],
                                                     n_list[cur], chyp[cur ^ 1], cidx[cur ^ 1], n_list[cur ^ 1]);
        DA3S_LAUNCH_CHECK(ctx);
        cur ^= 1;
        rc = score(1 << round, chyp[cur], cidx[cur], n_list[cur], n_hyp, zs, false);
        if (rc != DA3S_OK) return rc;
    }
    ctx->ws_top = save_top;     // consumed in stream order
    return DA3S_OK;
}

extern "C" int da3s_ransac_inlier_mask(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                                       int world, int valid_depth, float depth_eps, const float* conf_thr, const float* depth_scale,
                                       const float* best_A, const float* best_t, const uint8_t* best_ok, float ransac_thr,
                                       uint8_t* mask_out, void* stream) {
    if (!ctx || !pairs || !conf_thr || !best_A || !best_t || !best_ok || !mask_out || n_pairs <= 0) return DA3S_EINVAL;
    if (n_pairs > 65535) return DA3S_EINVAL;
    long long P; int tpf;
    if (!pair_geometry(overlap, H, W, P, tpf)) return DA3S_EINVAL;
    RansacArgs r;
    fill_ransac_args(r, pairs, n_pairs, overlap, H, W, P, tpf, world, valid_depth, depth_eps, conf_thr, depth_scale, 1, ransac_thr);
    dim3 grid(tpf * overlap, n_pairs);
    ransac_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(r, best_A, best_t, best_ok, mask_out);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_align_pairs(da3s_ctx* ctx, const da3s_pair* pairs, int n_pairs, int overlap, int H, int W,
                                const da3s_align_opts* opts, const int32_t* sample_idx, double* sim3_rows,
                                da3s_pair_aux* aux, int32_t* hyp_counts_out, void* stream) {
    if (!ctx || !pairs || !opts || !sim3_rows || n_pairs <= 0) return DA3S_EINVAL;
    if (n_pairs > 65535) return DA3S_EINVAL;
    if (opts->n_hyp < 0 || (opts->n_hyp > 0 && !sample_idx)) return DA3S_EINVAL;
    if (opts->max_iterations <= 0 || opts->min_points < 0) return DA3S_EINVAL;
    if (opts->depth_scale_mode < 0 || opts->depth_scale_mode > 2) return DA3S_EINVAL;
    if (opts->depth_scale_mode && overlap != 1) return DA3S_EINVAL;
    long long P; int tpf;
    if (!pair_geometry(overlap, H, W, P, tpf)) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ws_reset(ctx);
    const int n_tiles = tpf * overlap;
    WS_ALLOC(ctx, float, thr, n_pairs);
    WS_ALLOC(ctx, float, dscale, n_pairs);
    WS_ALLOC(ctx, PairState, state, n_pairs);
    WS_ALLOC(ctx, double, eff, (size_t)n_pairs * overlap * EFF_LEN);
    WS_ALLOC(ctx, float, gate, (size_t)n_pairs * 12);
    WS_ALLOC(ctx, double, partials, (size_t)n_pairs * n_tiles * MOM_LEN);
    WS_ALLOC(ctx, unsigned int, tickets, n_pairs);
    WS_ALLOC(ctx, int, n_active, 1);
    WS_ALLOC(ctx, int, done_count, opts->max_iterations);
    WS_ALLOC(ctx, unsigned long long, work_counter, opts->max_iterations);
    const int q_cap = n_pairs * (opts->max_iterations + 1);
    WS_ALLOC(ctx, int, q_pair, (size_t)q_cap);
    WS_ALLOC(ctx, unsigned int, q_reserve, 1);

    int rc = thresholds_impl(ctx, pairs, n_pairs, overlap, H, W, opts, thr, dscale, aux, st);
    if (rc != DA3S_OK) return rc;

    PairArgs a;
    a.pairs = pairs; a.n_pairs = n_pairs; a.overlap = overlap; a.H = H; a.W = W; a.P = P; a.tiles_per_frame = tpf;
    a.world = opts->world; a.valid_depth = opts->valid_depth; a.huber = opts->huber; a.variant = SOLVE_WEIGHTED;
    a.depth_eps = opts->depth_eps; a.thr = thr; a.dscale = dscale; a.state = state; a.eff = eff;
    a.gate = opts->n_hyp > 0 ? gate : nullptr;
    a.gate_thr2 = (float)((double)opts->ransac_thr * (double)opts->ransac_thr);
    a.partials = partials; a.tickets = tickets; a.n_active = n_active; a.done_count = done_count; a.work_counter = work_counter; a.huber_delta = opts->huber_delta; a.tol = opts->tol;
    a.delta_f = (float)opts->huber_delta; a.delta2_f = a.delta_f * a.delta_f; a.W_f = (float)W;
    a.max_iterations = opts->max_iterations; a.min_points = opts->min_points; a.precise = opts->precise;
    a.rows = sim3_rows; a.aux = aux;
    a.q_pair = q_pair; a.q_reserve = q_reserve; a.q_cap = q_cap;

    int threads = 128, blocks = (n_pairs + threads - 1) / threads;
    pair_state_init_kernel<<<blocks, threads, 0, st>>>(a);
    DA3S_LAUNCH_CHECK(ctx);

    if (opts->n_hyp > 0) {
        const int nh = opts->n_hyp;
        WS_ALLOC(ctx, float, hyp_A, (size_t)n_pairs * nh * 9);
        WS_ALLOC(ctx, float, hyp_t, (size_t)n_pairs * nh * 3);
        WS_ALLOC(ctx, uint8_t, hyp_ok, (size_t)n_pairs * nh);
        int32_t* counts = hyp_counts_out;
        if (!counts) { WS_ALLOC(ctx, int32_t, c2, (size_t)n_pairs * nh); counts = c2; }
        rc = da3s_ransac_hypotheses(ctx, pairs, n_pairs, overlap, H, W, opts->world, opts->valid_depth, opts->depth_eps,
                                    thr, dscale, sample_idx, nh, hyp_A, hyp_t, hyp_ok, nullptr, stream);
        if (rc != DA3S_OK) return rc;
        // the winner is all that is needed unless the caller asked for the count table: rounds with exact pruning
        static const bool no_rounds = getenv("DA3S_RANSAC_FULL") && getenv("DA3S_RANSAC_FULL")[0] == '1';
        if (!hyp_counts_out && tpf >= 8 && !no_rounds) {
            RansacArgs r;
            fill_ransac_args(r, pairs, n_pairs, overlap, H, W, P, tpf, opts->world, opts->valid_depth, opts->depth_eps, thr, dscale, nh,
                             opts->ransac_thr);
            rc = ransac_score_rounds(ctx, r, hyp_A, hyp_t, hyp_ok, counts, st);
        } else {
            rc = da3s_ransac_score(ctx, pairs, n_pairs, overlap, H, W, opts->world, opts->valid_depth, opts->depth_eps,
                                   thr, dscale, hyp_A, hyp_t, hyp_ok, nh, opts->ransac_thr, counts, stream);
        }
        if (rc != DA3S_OK) return rc;
        ransac_best_kernel<<<n_pairs, 256, 0, st>>>(counts, hyp_ok, hyp_A, hyp_t, nh, opts->ransac_min_inliers, state, gate,
                                                    sim3_rows, aux, n_active);
        DA3S_LAUNCH_CHECK(ctx);
    }

    const bool vec = (P % 4 == 0);          // per-pair pointer alignment is the caller's contract (EALIGN documented)
    dim3 grid(n_tiles, n_pairs);
    const int iters = opts->huber ? opts->max_iterations : 1;
    if (opts->precise) {
        for (int it = 0; it < iters; ++it) {
            if (vec) pair_moments_kernel<true><<<grid, PA_THREADS, 0, st>>>(a);
            else     pair_moments_kernel<false><<<grid, PA_THREADS, 0, st>>>(a);
            DA3S_LAUNCH_CHECK(ctx);
        }
        return DA3S_OK;
    }
    // persistent launch: as many blocks as can be resident, never more than there are items
    const bool use_gate = a.gate != nullptr;
    // PK (packed pixel pairs) needs the float4 path and an even row length; DA3S_IRLS_SCALAR=1 forces the scalar form (A/B runs)
    static const bool force_scalar = getenv("DA3S_IRLS_SCALAR") && getenv("DA3S_IRLS_SCALAR")[0] == '1';
    const bool pk = vec && (W % 2 == 0) && !force_scalar;
    static const void* const table[12] = {
        (const void*)pair_moments_mixed_kernel<false, false, false, false>, (const void*)pair_moments_mixed_kernel<false, false, true, false>,
        (const void*)pair_moments_mixed_kernel<false, true, false, false>,  (const void*)pair_moments_mixed_kernel<false, true, true, false>,
        (const void*)pair_moments_mixed_kernel<true, false, false, false>,  (const void*)pair_moments_mixed_kernel<true, false, true, false>,
        (const void*)pair_moments_mixed_kernel<true, true, false, false>,   (const void*)pair_moments_mixed_kernel<true, true, true, false>,
        (const void*)pair_moments_mixed_kernel<true, false, false, true>,   (const void*)pair_moments_mixed_kernel<true, false, true, true>,
        (const void*)pair_moments_mixed_kernel<true, true, false, true>,    (const void*)pair_moments_mixed_kernel<true, true, true, true>};
    const void* fn = table[(pk ? 8 : (vec ? 4 : 0)) | (use_gate ? 2 : 0) | (opts->huber ? 1 : 0)];
    int per_sm = 0;
    const size_t ring_bytes = vec ? (size_t)PM_RING_BYTES : 0;
    if (ring_bytes)
        DA3S_CHECK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
    DA3S_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PM_THREADS, ring_bytes));
    if (per_sm < 1) return DA3S_ECUDA;
    long long n_blocks = (long long)per_sm * ctx->sm_count;
    const long long items = (long long)n_pairs * n_tiles;
    if (n_blocks > items) n_blocks = items;
    void* kargs[] = {(void*)&a};
    pm_queue_init_kernel<<<1, 1024, 0, st>>>(a);
    DA3S_LAUNCH_CHECK(ctx);
    // no grid-wide barrier inside: an ordinary launch (a block only ever waits for entries that running blocks produce)
    prof_begin(ctx, DA3S_TIMED_IRLS, st);
    DA3S_CHECK_CUDA(ctx, cudaLaunchKernel(fn, dim3((unsigned int)n_blocks), dim3(PM_THREADS), kargs, ring_bytes, st));
    prof_end(ctx, DA3S_TIMED_IRLS, st);
    ctx->launches++;
    return DA3S_OK;
}

// ---------------------------------------------------------------------------------
// Umeyama / IRLS on MATERIALISED correspondences (the reference's array-level API:
// utils/align.py:14-40, :111-218, :224-276; align_geometry.py:59-82).  Same reduction and
// solve as above with one "pair", camera mode, and optional gather lists.
// Algorithmic bytes: 12+12+4 = 28 B per point pair (float32), 56 B (float64).
// ---------------------------------------------------------------------------------
#define PT_THREADS 256
#define PT_PER_BLOCK 4096

struct PointsArgs {
    const void* src; const void* dst;
    const void* weights; int weights_f64;
    const float* conf_src; const float* conf_dst;
    long long n; const long long* idx_src; const long long* idx_dst; long long count;
    int irls;
    double* norm_state;                     // [8]: mx[3], my[3], n, unused  (NORMRATIO)
    PairArgs pa;
};

template <typename T>
__device__ __forceinline__ void load_pt(const void* base, long long i, double* p) {
    const T* q = (const T*)base + 3 * i;
    p[0] = (double)q[0]; p[1] = (double)q[1]; p[2] = (double)q[2];
}

template <typename T, int STAGE>            // STAGE 0: moments (+solve); 1: centred norm sums (NORMRATIO pass 2)
__global__ void __launch_bounds__(PT_THREADS)
points_moments_kernel(PointsArgs a) {
    __shared__ double red[PT_THREADS / 32][MOM_LEN];
    __shared__ bool is_last;
    __shared__ double eff[EFF_LEN];
    __shared__ double cen[6];
    PairArgs& pa = a.pa;
    if (STAGE == 0 && pa.state[0].done) return;
    if (threadIdx.x < EFF_LEN && pa.eff) eff[threadIdx.x] = pa.eff[threadIdx.x];
    if (STAGE == 1 && threadIdx.x < 6) cen[threadIdx.x] = a.norm_state[threadIdx.x];
    __syncthreads();
    double acc[MOM_LEN];
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) acc[k] = 0.0;
    const long long begin = (long long)blockIdx.x * PT_PER_BLOCK;
    long long end = begin + PT_PER_BLOCK;
    if (end > a.count) end = a.count;
    for (long long i = begin + threadIdx.x; i < end; i += PT_THREADS) {
        const long long is = a.idx_src ? a.idx_src[i] : i;
        const long long id = a.idx_dst ? a.idx_dst[i] : i;
        double X[3], Y[3];
        load_pt<T>(a.src, is, X);
        load_pt<T>(a.dst, id, Y);
        if (STAGE == 1) {
            double dx0 = X[0] - cen[0], dx1 = X[1] - cen[1], dx2 = X[2] - cen[2];
            double dy0 = Y[0] - cen[3], dy1 = Y[1] - cen[4], dy2 = Y[2] - cen[5];
            acc[0] += sqrt(dx0 * dx0 + dx1 * dx1 + dx2 * dx2);
            acc[1] += sqrt(dy0 * dy0 + dy1 * dy1 + dy2 * dy2);
            continue;
        }
        double w = 1.0;
        if (a.weights) w = a.weights_f64 ? ((const double*)a.weights)[i] : (double)((const float*)a.weights)[i];
        if (a.irls) {
            w = (double)__fsqrt_rn(__fmul_rn(a.conf_dst[id], a.conf_src[is]));       // utils/align.py:166
            double r0 = Y[0] + (eff[9] * X[0] + eff[10] * X[1] + eff[11] * X[2] + eff[18]);
            double r1 = Y[1] + (eff[12] * X[0] + eff[13] * X[1] + eff[14] * X[2] + eff[19]);
            double r2 = Y[2] + (eff[15] * X[0] + eff[16] * X[1] + eff[17] * X[2] + eff[20]);
            double r = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
            if (r > pa.huber_delta) w *= pa.huber_delta / r;
            acc[MOM_SR] += r;
        }
        const double wx0 = w * X[0], wx1 = w * X[1], wx2 = w * X[2];
        const double wy0 = w * Y[0], wy1 = w * Y[1], wy2 = w * Y[2];
        acc[MOM_S0] += w;
        acc[MOM_SX] += wx0; acc[MOM_SX + 1] += wx1; acc[MOM_SX + 2] += wx2;
        acc[MOM_SY] += wy0; acc[MOM_SY + 1] += wy1; acc[MOM_SY + 2] += wy2;
        acc[MOM_SYX + 0] += wy0 * X[0]; acc[MOM_SYX + 1] += wy0 * X[1]; acc[MOM_SYX + 2] += wy0 * X[2];
        acc[MOM_SYX + 3] += wy1 * X[0]; acc[MOM_SYX + 4] += wy1 * X[1]; acc[MOM_SYX + 5] += wy1 * X[2];
        acc[MOM_SYX + 6] += wy2 * X[0]; acc[MOM_SYX + 7] += wy2 * X[1]; acc[MOM_SYX + 8] += wy2 * X[2];
        acc[MOM_SXX + 0] += wx0 * X[0]; acc[MOM_SXX + 1] += wx0 * X[1]; acc[MOM_SXX + 2] += wx0 * X[2];
        acc[MOM_SXX + 3] += wx1 * X[1]; acc[MOM_SXX + 4] += wx1 * X[2]; acc[MOM_SXX + 5] += wx2 * X[2];
        acc[MOM_WMAX] = fmax(acc[MOM_WMAX], w);
        acc[MOM_N] += 1.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) {
        double v = (k == MOM_WMAX) ? warp_max(acc[k]) : warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    double* prow = pa.partials + (size_t)blockIdx.x * MOM_LEN;
    if (threadIdx.x < MOM_LEN) {
        double v = red[0][threadIdx.x];
        for (int w = 1; w < PT_THREADS / 32; ++w)
            v = (threadIdx.x == MOM_WMAX) ? fmax(v, red[w][threadIdx.x]) : v + red[w][threadIdx.x];
        prow[threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&pa.tickets[0], 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double tot[MOM_LEN];
    if (threadIdx.x < MOM_LEN) {
        double v = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) {
            double p = __ldcg(pa.partials + (size_t)b * MOM_LEN + threadIdx.x);
            v = (threadIdx.x == MOM_WMAX) ? fmax(v, p) : v + p;
        }
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    pa.tickets[0] = 0;
    double m[MOM_LEN];
    for (int k = 0; k < MOM_LEN; ++k) m[k] = tot[k];
    if (STAGE == 1) {
        // utils/align.py:250-276: s = sum|y-cy| / sum|x-cx|; Kabsch on H = s Xc^T Yc; R = V S U^T
        PairState st = pa.state[0];
        const double* ns = a.norm_state;
        double mx[3] = {ns[0], ns[1], ns[2]}, my[3] = {ns[3], ns[4], ns[5]};
        double s = m[0] > 0 ? m[1] / m[0] : 1.0;
        double Hm[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Hm[3 * i + j] = s * ns[8 + 3 * j + i];          // (C^T)[i][j], C = centred sum y x^T
        double U[9], Sg[3], V[9], VUt[9];
        svd3(Hm, U, Sg, V);
        mat3_mul_bt(V, U, VUt);
        double dsign = det3(VUt) < 0 ? -1.0 : 1.0;
        double Vd[9];
        for (int i = 0; i < 3; ++i) { Vd[3 * i] = V[3 * i]; Vd[3 * i + 1] = V[3 * i + 1]; Vd[3 * i + 2] = V[3 * i + 2] * dsign; }
        mat3_mul_bt(Vd, U, st.R);
        double Rm[3];
        mat3_vec(st.R, mx, Rm);
        for (int k = 0; k < 3; ++k) st.t[k] = my[k] - s * Rm[k];
        st.s = s; st.iters = 1; st.done = 1; st.n_valid = ns[6];
        pa.state[0] = st;
        write_row(pa, 0, st);
        return;
    }
    if (a.norm_state) {
        // NORMRATIO pass 1: centroids (utils/align.py:242-243) and the centred cross sums
        double n = m[MOM_N];
        double* ns = a.norm_state;
        for (int k = 0; k < 3; ++k) { ns[k] = m[MOM_SX + k] / n; ns[3 + k] = m[MOM_SY + k] / n; }
        ns[6] = n; ns[7] = 0.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) ns[8 + 3 * i + j] = m[MOM_SYX + 3 * i + j] - n * ns[3 + i] * ns[j];
        return;
    }
    solve_pair(pa, 0, m);
}

static int points_common(da3s_ctx* ctx, PointsArgs& a, long long count, int huber, int variant, double delta,
                         int max_it, double tol, int min_points, double* row) {
    WS_ALLOC(ctx, PairState, state, 1);
    WS_ALLOC(ctx, double, eff, EFF_LEN);
    WS_ALLOC(ctx, unsigned int, tickets, 1);
    long long nb = (count + PT_PER_BLOCK - 1) / PT_PER_BLOCK;
    if (nb < 1) nb = 1;
    if (nb > 2147483647LL) return DA3S_EINVAL;
    WS_ALLOC(ctx, double, partials, (size_t)nb * MOM_LEN);
    PairArgs& pa = a.pa;
    pa.pairs = nullptr; pa.n_pairs = 1; pa.overlap = 1; pa.H = 1; pa.W = 1; pa.P = 1; pa.tiles_per_frame = (int)nb;
    pa.world = 0; pa.valid_depth = 0; pa.huber = huber; pa.variant = variant; pa.depth_eps = 0; pa.thr = nullptr;
    pa.dscale = nullptr; pa.state = state; pa.eff = eff; pa.gate = nullptr; pa.gate_thr2 = 0; pa.partials = partials;
    pa.tickets = tickets; pa.n_active = nullptr; pa.done_count = nullptr; pa.work_counter = nullptr; pa.q_pair = nullptr; pa.q_reserve = nullptr; pa.q_cap = 0; pa.huber_delta = delta; pa.tol = tol; pa.max_iterations = max_it; pa.min_points = min_points;
    pa.precise = 1;
    pa.rows = row; pa.aux = nullptr;
    a.count = count;
    return (int)nb;
}

extern "C" int da3s_umeyama_points(da3s_ctx* ctx, const void* src, const void* dst, int points_f64,
                                   const void* weights, int weights_f64, long long n,
                                   const long long* idx_src, const long long* idx_dst, long long n_idx,
                                   int variant, double* sim3_row, void* stream) {
    if (!ctx || !src || !dst || !sim3_row || n <= 0) return DA3S_EINVAL;
    if ((idx_src == nullptr) != (idx_dst == nullptr)) return DA3S_EINVAL;
    if (variant < 0 || variant > 3) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ws_reset(ctx);
    PointsArgs a;
    a.src = src; a.dst = dst; a.weights = (variant == DA3S_UMEYAMA_WEIGHTED || variant == DA3S_UMEYAMA_LEGACY_TRACE) ? weights : nullptr; a.weights_f64 = weights_f64;
    a.conf_src = nullptr; a.conf_dst = nullptr; a.n = n; a.idx_src = idx_src; a.idx_dst = idx_dst; a.irls = 0;
    a.norm_state = nullptr;
    long long count = idx_src ? n_idx : n;
    if (count <= 0) return DA3S_EINVAL;
    const int solve = variant == DA3S_UMEYAMA_WEIGHTED ? SOLVE_WEIGHTED : (variant == DA3S_UMEYAMA_LEGACY_TRACE ? SOLVE_LEGACY_TRACE : SOLVE_MEAN);
    int nb = points_common(ctx, a, count, 0, solve, 1.0, 1, 0.0, 0, sim3_row);
    if (nb < 0) return nb;
    pair_state_init_kernel<<<1, 32, 0, st>>>(a.pa);
    DA3S_LAUNCH_CHECK(ctx);
    if (variant == DA3S_UMEYAMA_NORMRATIO) {
        WS_ALLOC(ctx, double, ns, 17);
        a.norm_state = ns;
        if (points_f64) points_moments_kernel<double, 0><<<nb, PT_THREADS, 0, st>>>(a);
        else            points_moments_kernel<float, 0><<<nb, PT_THREADS, 0, st>>>(a);
        DA3S_LAUNCH_CHECK(ctx);
        if (points_f64) points_moments_kernel<double, 1><<<nb, PT_THREADS, 0, st>>>(a);
        else            points_moments_kernel<float, 1><<<nb, PT_THREADS, 0, st>>>(a);
        DA3S_LAUNCH_CHECK(ctx);
        return DA3S_OK;
    }
    if (points_f64) points_moments_kernel<double, 0><<<nb, PT_THREADS, 0, st>>>(a);
    else            points_moments_kernel<float, 0><<<nb, PT_THREADS, 0, st>>>(a);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_irls_points(da3s_ctx* ctx, const void* src, const void* dst, int points_f64,
                                const float* conf_src, const float* conf_dst, long long n,
                                const long long* idx_src, const long long* idx_dst, long long n_idx,
                                double huber_delta, int max_iterations, double tol, double* sim3_row, void* stream) {
    if (!ctx || !src || !dst || !conf_src || !conf_dst || !sim3_row || n <= 0 || max_iterations <= 0) return DA3S_EINVAL;
    if ((idx_src == nullptr) != (idx_dst == nullptr)) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ws_reset(ctx);
    PointsArgs a;
    a.src = src; a.dst = dst; a.weights = nullptr; a.weights_f64 = 0; a.conf_src = conf_src; a.conf_dst = conf_dst;
    a.n = n; a.idx_src = idx_src; a.idx_dst = idx_dst; a.irls = 1; a.norm_state = nullptr;
    long long count = idx_src ? n_idx : n;
    if (count <= 0) return DA3S_EINVAL;
    int nb = points_common(ctx, a, count, 1, SOLVE_WEIGHTED, huber_delta, max_iterations, tol, 0, sim3_row);
    if (nb < 0) return nb;
    pair_state_init_kernel<<<1, 32, 0, st>>>(a.pa);
    DA3S_LAUNCH_CHECK(ctx);
    for (int it = 0; it < max_iterations; ++it) {
        if (points_f64) points_moments_kernel<double, 0><<<nb, PT_THREADS, 0, st>>>(a);
        else            points_moments_kernel<float, 0><<<nb, PT_THREADS, 0, st>>>(a);
        DA3S_LAUNCH_CHECK(ctx);
    }
    return DA3S_OK;
}

// ---------------------------------------------------------------------------------
// Sim(3) chain (utils/geometry.py:73-119): n tiny compositions, one thread.
// ---------------------------------------------------------------------------------
// rows are staged through shared memory in chunks (coalesced loads/stores by the whole block);
// the n sequential 3x3 compositions themselves run on one thread out of shared memory.
#define ACC_CHUNK 128
__global__ void __launch_bounds__(256)
accumulate_sim3_kernel(const double* __restrict__ rows, int n, double* __restrict__ cum) {
    __shared__ double in[ACC_CHUNK * DA3S_ROW_LEN];
    __shared__ double out[ACC_CHUNK * 13];
    __shared__ double cur[13];
    if (threadIdx.x < 13) {
        double v = (threadIdx.x == 0 || threadIdx.x == 1 || threadIdx.x == 5 || threadIdx.x == 9) ? 1.0 : 0.0;
        cur[threadIdx.x] = v;
        cum[threadIdx.x] = v;
    }
    __syncthreads();
    for (int base = 0; base < n; base += ACC_CHUNK) {
        const int m = (n - base) < ACC_CHUNK ? (n - base) : ACC_CHUNK;
        for (int i = threadIdx.x; i < m * DA3S_ROW_LEN; i += blockDim.x) in[i] = rows[(size_t)base * DA3S_ROW_LEN + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = cur[0], R[9], t[3];
            for (int k = 0; k < 9; ++k) R[k] = cur[1 + k];
            for (int k = 0; k < 3; ++k) t[k] = cur[10 + k];
            for (int i = 0; i < m; ++i) {
                const double* r = in + i * DA3S_ROW_LEN;
                const double sn = r[DA3S_ROW_S];
                const double* Rn = r + DA3S_ROW_R;
                const double* tn = r + DA3S_ROW_T;
                double Rt[3], R2[9];
                mat3_vec(R, tn, Rt);                              // t' = s_prev (R_prev t_next) + t_prev
                for (int k = 0; k < 3; ++k) t[k] = s * Rt[k] + t[k];
                mat3_mul(R, Rn, R2);                              // R' = R_prev R_next
                for (int k = 0; k < 9; ++k) R[k] = R2[k];
                s = s * sn;                                       // s' = s_prev s_next
                double* o = out + i * 13;
                o[0] = s;
                for (int k = 0; k < 9; ++k) o[1 + k] = R[k];
                for (int k = 0; k < 3; ++k) o[10 + k] = t[k];
            }
            cur[0] = s;
            for (int k = 0; k < 9; ++k) cur[1 + k] = R[k];
            for (int k = 0; k < 3; ++k) cur[10 + k] = t[k];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < m * 13; i += blockDim.x) cum[(size_t)(base + 1) * 13 + i] = out[i];
        __syncthreads();
    }
}

extern "C" int da3s_accumulate_sim3(da3s_ctx* ctx, const double* rows, int n_rows, double* cum, void* stream) {
    if (!ctx || !cum || n_rows < 0 || (n_rows > 0 && !rows)) return DA3S_EINVAL;
    accumulate_sim3_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(rows, n_rows, cum);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}
