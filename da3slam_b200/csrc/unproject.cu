// unproject.cu — camera table, K1 (depth -> xyz fused with filtering and an optional
// Sim(3)), K5 (apply Sim(3)).  HBM-bound streaming kernels: 128-bit coalesced loads,
// xyz staged through shared memory so that every store instruction of a warp writes
// 512 contiguous bytes, grid = (tiles per frame, frames).
//
// Algorithmic bytes: K1 = 4 (depth) + 4 (conf) read + 12 (xyz f32) + 1 (mask) write
// = 21 B / pixel; K5 = 12 + 12 = 24 B / point.
#include "common.cuh"
#include "sim3_math.cuh"
#include "unproject_frame.cuh"

// ---------------------------------------------------------------------------------
// camera table
// ---------------------------------------------------------------------------------
__global__ void build_cams_kernel(const float* __restrict__ K9, const float* __restrict__ E12, int n,
                                  int inverse_mode, da3s_cam* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* K = K9 + 9 * (size_t)i;
    const float* E = E12 + 12 * (size_t)i;
    da3s_cam c;
    c.fu = K[0]; c.fv = K[4]; c.cu = K[2]; c.cv = K[5];
    c.inv_fu = __fdiv_rn(1.0f, c.fu);
    c.inv_fv = __fdiv_rn(1.0f, c.fv);
    c.skew_flag = (K[1] != 0.0f || K[3] != 0.0f) ? 1.0f : 0.0f;
    c.reserved = 0.0f;
    double Kd[9], Ki[9];
    for (int k = 0; k < 9; ++k) Kd[k] = (double)K[k];
    if (!mat3_inv(Kd, Ki))
        for (int k = 0; k < 9; ++k) Ki[k] = nan("");
    for (int k = 0; k < 9; ++k) c.kinv[k] = Ki[k];
    double R[9], t[3], Rinv[9];
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) R[3 * r + k] = (double)E[4 * r + k];
        t[r] = (double)E[4 * r + 3];
    }
    if (inverse_mode == DA3S_CAM_GENERAL_INV) {
        if (!mat3_inv(R, Rinv))
            for (int k = 0; k < 9; ++k) Rinv[k] = nan("");
    } else {
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) Rinv[3 * r + k] = R[3 * k + r];
    }
    double ti[3];
    mat3_vec(Rinv, t, ti);
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) c.c2w[4 * r + k] = Rinv[3 * r + k];
        c.c2w[4 * r + 3] = -ti[r];
    }
    out[i] = c;
}

extern "C" int da3s_build_cams(da3s_ctx* ctx, const float* K9, const float* E12, int n_frames, int inverse_mode,
                               da3s_cam* cams_out, void* stream) {
    if (!ctx || !K9 || !E12 || !cams_out || n_frames <= 0) return DA3S_EINVAL;
    if (inverse_mode != DA3S_CAM_CLOSED_FORM && inverse_mode != DA3S_CAM_GENERAL_INV) return DA3S_EINVAL;
    int threads = 128;
    build_cams_kernel<<<(n_frames + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
        K9, E12, n_frames, inverse_mode, cams_out);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

// ---------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------
#define K1_THREADS 256
#define K1_PIX_PER_THREAD 4

struct K1Args {
    const da3s_frame_job* jobs;  // nullable: per-frame pointers (batched launch over frames of many submaps)
    const float* depth; const float* conf; const da3s_cam* cams;
    int H, W; long long P;      // P = H*W
    int flags;
    float conf_thr; const float* conf_thr_dev; float conf_floor; float depth_eps;
    const double* sim3;
    void* xyz; uint8_t* mask; unsigned long long* n_kept;
    int groups_per_block;       // float4 groups handled by one block (multiple of K1_THREADS)
};

template <int MODE, typename OutT, bool VEC>
__global__ void __launch_bounds__(K1_THREADS, (MODE == DA3S_UNPROJ_FAST && sizeof(OutT) == 4) ? 4 : 2)
unproject_filter_kernel(K1Args a) {
    __shared__ UnprojFrame fr;
    __shared__ __align__(16) float stage[VEC ? (K1_THREADS / 32) * 32 * 12 : 4];     // per-warp staging for f32 xyz
    __shared__ unsigned int blk_kept;
    __shared__ float thr_sh;
    const int frame = blockIdx.y;
    const bool world = a.flags & DA3S_UNPROJ_WORLD;
    // per-frame pointers: from the job table (batched launch) or from the flat arrays
    const float* depth; const float* conf; const da3s_cam* cam; const double* s3; const float* thr_dev;
    OutT* xyz; uint8_t* mask;
    if (a.jobs) {
        const da3s_frame_job j = a.jobs[frame];
        depth = j.depth; conf = j.conf; cam = j.cam; s3 = j.sim3; thr_dev = j.conf_thr;
        xyz = (OutT*)j.xyz; mask = j.mask;
    } else {
        const size_t frame_off = (size_t)frame * (size_t)a.P;
        depth = a.depth + frame_off; conf = a.conf ? a.conf + frame_off : nullptr; cam = a.cams + frame;
        s3 = a.sim3 ? a.sim3 + ((a.flags & DA3S_SIM3_PER_FRAME) ? 13 * (size_t)frame : 0) : nullptr;
        thr_dev = a.conf_thr_dev;
        xyz = (OutT*)a.xyz + frame_off * 3; mask = a.mask ? a.mask + frame_off : nullptr;
    }
    const bool xform = world || s3 != nullptr;
    if (threadIdx.x == 0) {
        compose_unproj_frame(*cam, s3, world, fr);
        blk_kept = 0;
        thr_sh = thr_dev ? *thr_dev : a.conf_thr;
    }
    __syncthreads();

    const float thr = thr_sh;
    unsigned int kept = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // block-uniform switches, evaluated once; the per-pixel test below is branch-free
    // a NaN threshold is what da3s_select returns for a segment without usable confidences: the reference then keeps
    // every point (viewer.py:333-338: `else: conf_mask = np.ones(...)`), so the confidence tests are switched off
    const bool use_conf = conf && !(thr != thr);
    const bool f_gt = use_conf && (a.flags & DA3S_MASK_CONF_GT), f_ge = use_conf && (a.flags & DA3S_MASK_CONF_GE);
    const bool f_floor = use_conf && (a.flags & DA3S_MASK_CONF_FLOOR), f_depth = a.flags & DA3S_MASK_DEPTH;
    const bool f_wz = a.flags & DA3S_MASK_WORLD_Z;
    const float cfloor = a.conf_floor, deps = a.depth_eps;
    typedef typename K1Val<MODE>::type CT;
    auto keep_of = [&](float d, float c, CT X, CT Y, CT Z) -> bool {
        bool k = (!f_gt || c > thr) & (!f_ge || c >= thr) & (!f_floor || c > cfloor);
        k = k & (!f_depth || ((d > deps) & is_finite_f(d)));
        if (f_wz)
            k = k && (Z > (CT)0.1) && (Z < (CT)50.0) && (fabs((double)X) < INFINITY) && (fabs((double)Y) < INFINITY) &&
                (fabs((double)Z) < INFINITY);
        return k;
    };

    if (VEC) {
        const long long n_groups = a.P >> 2;
        const long long g_begin = (long long)blockIdx.x * a.groups_per_block;
        long long g_end = g_begin + a.groups_per_block;
        if (g_end > n_groups) g_end = n_groups;
        const float4* d4 = reinterpret_cast<const float4*>(depth);
        const float4* c4 = reinterpret_cast<const float4*>(conf);
        // one float4 group = 4 pixels -> 12 output values; staged per warp so every store is coalesced
        auto emit = [&](long long gb, long long g, int u, int v, bool active, const float* dv, const float* cv) {
            const long long pix = g << 2;
            unsigned int mbits = 0;
            OutT o[12];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                CT X, Y, Z;
                unproject_pixel<MODE>(fr, u, v, dv[j], xform, X, Y, Z);
                o[3 * j] = (OutT)X; o[3 * j + 1] = (OutT)Y; o[3 * j + 2] = (OutT)Z;
                if (active && keep_of(dv[j], cv[j], X, Y, Z)) mbits |= (1u << (8 * j));
                if (++u == a.W) { u = 0; ++v; }
            }
            kept += __popc(mbits);
            if (mask && active) reinterpret_cast<unsigned int*>(mask)[g] = mbits;
            if (sizeof(OutT) == 4) {
                float* ws = stage + warp * (32 * 12);
                float4* ws4 = reinterpret_cast<float4*>(ws);
                ws4[lane * 3 + 0] = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
                ws4[lane * 3 + 1] = make_float4((float)o[4], (float)o[5], (float)o[6], (float)o[7]);
                ws4[lane * 3 + 2] = make_float4((float)o[8], (float)o[9], (float)o[10], (float)o[11]);
                __syncwarp();
                const long long warp_g0 = gb + warp * 32;                  // first group of this warp
                float4* out4 = reinterpret_cast<float4*>(xyz) + warp_g0 * 3;
                long long warp_groups = g_end - warp_g0;                   // groups this warp really owns
                if (warp_groups > 32) warp_groups = 32;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int idx = j * 32 + lane;
                    if (idx < warp_groups * 3) stg_stream(out4 + idx, ws4[idx]);
                }
                __syncwarp();
            } else if (active) {
                double* out = (double*)xyz + (size_t)pix * 3;
#pragma unroll
                for (int j = 0; j < 12; ++j) out[j] = (double)o[j];
            }
        };
        // pixel coordinates are carried incrementally: one division per thread, then adds with carry
        const int W = a.W;
        const int step_px = K1_THREADS * 4, step_v = step_px / W, step_u = step_px - step_v * W;
        int v0, u0;
        { const long long p0 = (g_begin + threadIdx.x) << 2; v0 = (int)(p0 / W); u0 = (int)(p0 - (long long)v0 * W); }
        for (long long gb = g_begin; gb < g_end; gb += 2 * K1_THREADS) {  // block-uniform; two groups in flight per thread
            const long long g0 = gb + threadIdx.x, g1 = g0 + K1_THREADS;
            int u1 = u0 + step_u, v1 = v0 + step_v;
            if (u1 >= W) { u1 -= W; ++v1; }
            const bool a0 = g0 < g_end, a1 = g1 < g_end;
            float d0[4] = {0, 0, 0, 0}, c0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
            float4 t0, t1, q0, q1;
            if (a0) t0 = ldg_stream(d4 + g0);
            if (a1) t1 = ldg_stream(d4 + g1);
            if (conf && a0) q0 = ldg_stream(c4 + g0);
            if (conf && a1) q1 = ldg_stream(c4 + g1);
            if (a0) { d0[0] = t0.x; d0[1] = t0.y; d0[2] = t0.z; d0[3] = t0.w; }
            if (a1) { d1[0] = t1.x; d1[1] = t1.y; d1[2] = t1.z; d1[3] = t1.w; }
            if (conf && a0) { c0[0] = q0.x; c0[1] = q0.y; c0[2] = q0.z; c0[3] = q0.w; }
            if (conf && a1) { c1[0] = q1.x; c1[1] = q1.y; c1[2] = q1.z; c1[3] = q1.w; }
            emit(gb, g0, u0, v0, a0, d0, c0);
            if (gb + K1_THREADS < g_end) emit(gb + K1_THREADS, g1, u1, v1, a1, d1, c1);
            u0 = u1 + step_u; v0 = v1 + step_v;
            if (u0 >= W) { u0 -= W; ++v0; }
        }
    } else {
        const long long p_begin = (long long)blockIdx.x * a.groups_per_block * 4;
        long long p_end = p_begin + (long long)a.groups_per_block * 4;
        if (p_end > a.P) p_end = a.P;
        for (long long pix = p_begin + threadIdx.x; pix < p_end; pix += K1_THREADS) {
            float d = depth[pix], c = conf ? conf[pix] : 0.0f;
            int v = (int)(pix / a.W), u = (int)(pix - (long long)v * a.W);
            CT X, Y, Z;
            unproject_pixel<MODE>(fr, u, v, d, xform, X, Y, Z);
            bool k = keep_of(d, c, X, Y, Z);
            kept += k;
            if (mask) mask[pix] = k ? 1 : 0;
            OutT* out = xyz + (size_t)pix * 3;
            out[0] = (OutT)X; out[1] = (OutT)Y; out[2] = (OutT)Z;
        }
    }
    if (a.n_kept) {
        unsigned int wk = __reduce_add_sync(0xffffffffu, kept);
        if (lane == 0 && wk) atomicAdd(&blk_kept, wk);
        __syncthreads();
        if (threadIdx.x == 0 && blk_kept) atomicAdd(a.n_kept, (unsigned long long)blk_kept);
    }
}

template <int MODE, typename OutT>
static int launch_k1(da3s_ctx* ctx, const K1Args& a0, int n_frames, bool vec, cudaStream_t st) {
    K1Args a = a0;
    // tile: 8192 pixels per block (2048 float4 groups = 8 iterations of 256 threads)
    a.groups_per_block = 2048;
    long long groups = (a.P + 3) / 4;
    int tiles = (int)((groups + a.groups_per_block - 1) / a.groups_per_block);
    dim3 grid(tiles, n_frames);
    if (vec) unproject_filter_kernel<MODE, OutT, true><<<grid, K1_THREADS, 0, st>>>(a);
    else     unproject_filter_kernel<MODE, OutT, false><<<grid, K1_THREADS, 0, st>>>(a);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_unproject_filter(da3s_ctx* ctx, const float* depth, const float* conf, const da3s_cam* cams,
                                     int n_frames, int H, int W, int flags,
                                     float conf_thr, const float* conf_thr_dev, float conf_floor, float depth_eps,
                                     const double* sim3, void* xyz_out, uint8_t* mask_out,
                                     unsigned long long* n_kept, void* stream) {
    if (!ctx || !depth || !cams || !xyz_out || n_frames <= 0 || H <= 0 || W <= 0) return DA3S_EINVAL;
    if ((flags & DA3S_MASK_CONF_GT) && (flags & DA3S_MASK_CONF_GE)) return DA3S_EINVAL;
    int mode = flags & DA3S_UNPROJ_MODEMASK;
    if (mode == 3) return DA3S_EINVAL;
    if (n_frames > 65535) return DA3S_EINVAL;
    K1Args a;
    a.jobs = nullptr;
    a.depth = depth; a.conf = conf; a.cams = cams; a.H = H; a.W = W; a.P = (long long)H * W;
    a.flags = flags; a.conf_thr = conf_thr; a.conf_thr_dev = conf_thr_dev; a.conf_floor = conf_floor;
    a.depth_eps = depth_eps; a.sim3 = sim3; a.xyz = xyz_out; a.mask = mask_out; a.n_kept = n_kept;
    a.groups_per_block = 0;
    const bool f64 = flags & DA3S_UNPROJ_OUT_F64;
    // the vector path needs 16-byte aligned frames: P % 4 == 0 and aligned bases
    bool vec = (a.P % 4 == 0) && aligned16(depth) && (!conf || aligned16(conf)) && aligned16(xyz_out) &&
               (!mask_out || ((uintptr_t)mask_out & 3) == 0);
    cudaStream_t st = (cudaStream_t)stream;
#define K1_DISPATCH(M)                                                        \
    return f64 ? launch_k1<M, double>(ctx, a, n_frames, vec, st)              \
               : launch_k1<M, float>(ctx, a, n_frames, vec, st)
    if (mode == DA3S_UNPROJ_CLOSED) { K1_DISPATCH(DA3S_UNPROJ_CLOSED); }
    if (mode == DA3S_UNPROJ_KINV)   { K1_DISPATCH(DA3S_UNPROJ_KINV); }
    K1_DISPATCH(DA3S_UNPROJ_FAST);
#undef K1_DISPATCH
}

extern "C" int da3s_unproject_filter_jobs(da3s_ctx* ctx, const da3s_frame_job* jobs, int n_frames, int H, int W, int flags,
                                          float conf_thr, float conf_floor, float depth_eps,
                                          unsigned long long* n_kept, void* stream) {
    if (!ctx || !jobs || n_frames <= 0 || H <= 0 || W <= 0) return DA3S_EINVAL;
    if ((flags & DA3S_MASK_CONF_GT) && (flags & DA3S_MASK_CONF_GE)) return DA3S_EINVAL;
    int mode = flags & DA3S_UNPROJ_MODEMASK;
    if (mode == 3 || n_frames > 65535) return DA3S_EINVAL;
    K1Args a;
    a.jobs = jobs; a.depth = nullptr; a.conf = nullptr; a.cams = nullptr; a.H = H; a.W = W; a.P = (long long)H * W;
    a.flags = flags; a.conf_thr = conf_thr; a.conf_thr_dev = nullptr; a.conf_floor = conf_floor; a.depth_eps = depth_eps;
    a.sim3 = nullptr; a.xyz = nullptr; a.mask = nullptr; a.n_kept = n_kept; a.groups_per_block = 0;
    const bool f64 = flags & DA3S_UNPROJ_OUT_F64;
    const bool vec = (a.P % 4 == 0);        // per-job pointer alignment (16 B) is the caller's contract
    cudaStream_t st = (cudaStream_t)stream;
#define K1_DISPATCH(M)                                                        \
    return f64 ? launch_k1<M, double>(ctx, a, n_frames, vec, st)              \
               : launch_k1<M, float>(ctx, a, n_frames, vec, st)
    if (mode == DA3S_UNPROJ_CLOSED) { K1_DISPATCH(DA3S_UNPROJ_CLOSED); }
    if (mode == DA3S_UNPROJ_KINV)   { K1_DISPATCH(DA3S_UNPROJ_KINV); }
    K1_DISPATCH(DA3S_UNPROJ_FAST);
#undef K1_DISPATCH
}

// ---------------------------------------------------------------------------------
// K5: apply Sim(3):  out = s * (p R^T) + t   (utils/geometry.py:60-62, same operation order:
// rotate, scale, translate — not pre-multiplied — so float64 results track numpy's)
// ---------------------------------------------------------------------------------
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
apply_sim3_kernel(const InT* __restrict__ in, long long n, const double* __restrict__ sim3, OutT* __restrict__ out) {
    __shared__ double sm[13];
    if (threadIdx.x < 13) sm[threadIdx.x] = sim3[threadIdx.x];
    __syncthreads();
    const double s = sm[0];
    const double* R = sm + 1;
    const double* t = sm + 10;
    const bool f32io = sizeof(InT) == 4 && sizeof(OutT) == 4;
    if (f32io) {
        // 4 points = 3 float4 per thread; loads and stores both coalesced through shared memory
        __shared__ float4 stg[256 * 3];
        const long long n_groups = n >> 2;                         // full groups of 4 points
        const float4* in4 = reinterpret_cast<const float4*>(in);
        float4* out4 = reinterpret_cast<float4*>(out);
        const long long per_block = 256;
        for (long long gb = (long long)blockIdx.x * per_block; gb < n_groups; gb += (long long)gridDim.x * per_block) {
            long long gcount = n_groups - gb; if (gcount > per_block) gcount = per_block;
            const long long f4_count = gcount * 3;
            for (int j = 0; j < 3; ++j) {
                int idx = j * 256 + threadIdx.x;
                if (idx < f4_count) stg[idx] = ldg_stream(in4 + gb * 3 + idx);
            }
            __syncthreads();
            if (threadIdx.x < gcount) {
                float* p = reinterpret_cast<float*>(&stg[threadIdx.x * 3]);
                float q[12];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
                    q[3 * k]     = (float)(s * (x * R[0] + y * R[1] + z * R[2]) + t[0]);
                    q[3 * k + 1] = (float)(s * (x * R[3] + y * R[4] + z * R[5]) + t[1]);
                    q[3 * k + 2] = (float)(s * (x * R[6] + y * R[7] + z * R[8]) + t[2]);
                }
#pragma unroll
                for (int k = 0; k < 12; ++k) p[k] = q[k];
            }
            __syncthreads();
            for (int j = 0; j < 3; ++j) {
                int idx = j * 256 + threadIdx.x;
                if (idx < f4_count) stg_stream(out4 + gb * 3 + idx, stg[idx]);
            }
            __syncthreads();
        }
        // tail points (n % 4)
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            long long i = (n_groups << 2) + threadIdx.x;
            double x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
            out[3 * i]     = (OutT)(s * (x * R[0] + y * R[1] + z * R[2]) + t[0]);
            out[3 * i + 1] = (OutT)(s * (x * R[3] + y * R[4] + z * R[5]) + t[1]);
            out[3 * i + 2] = (OutT)(s * (x * R[6] + y * R[7] + z * R[8]) + t[2]);
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            double x = (double)in[3 * i], y = (double)in[3 * i + 1], z = (double)in[3 * i + 2];
            out[3 * i]     = (OutT)(s * (x * R[0] + y * R[1] + z * R[2]) + t[0]);
            out[3 * i + 1] = (OutT)(s * (x * R[3] + y * R[4] + z * R[5]) + t[1]);
            out[3 * i + 2] = (OutT)(s * (x * R[6] + y * R[7] + z * R[8]) + t[2]);
        }
    }
}

extern "C" int da3s_apply_sim3(da3s_ctx* ctx, const void* xyz_in, int in_f64, long long n_points,
                               const double* sim3, void* xyz_out, int out_f64, void* stream) {
    if (!ctx || !xyz_in || !xyz_out || !sim3 || n_points < 0) return DA3S_EINVAL;
    if (n_points == 0) return DA3S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long work = in_f64 || out_f64 ? (n_points + 255) / 256 : ((n_points >> 2) + 255) / 256;
    long long cap = (long long)ctx->sm_count * 16;
    int blocks = (int)(work < 1 ? 1 : (work > cap ? cap : work));
    if (!in_f64 && !out_f64) {
        if (!aligned16(xyz_in) || !aligned16(xyz_out)) return DA3S_EALIGN;
        apply_sim3_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)xyz_in, n_points, sim3, (float*)xyz_out);
    } else if (!in_f64 && out_f64) {
        apply_sim3_kernel<float, double><<<blocks, 256, 0, st>>>((const float*)xyz_in, n_points, sim3, (double*)xyz_out);
    } else if (in_f64 && out_f64) {
        apply_sim3_kernel<double, double><<<blocks, 256, 0, st>>>((const double*)xyz_in, n_points, sim3, (double*)xyz_out);
    } else {
        apply_sim3_kernel<double, float><<<blocks, 256, 0, st>>>((const double*)xyz_in, n_points, sim3, (float*)xyz_out);
    }
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

// ---------------------------------------------------------------------------------
// Ordered filter + compaction of a resident cloud (the viewer's map push, viewer.py:333-355): keep point i iff
// valid[i] (nullable) and conf[i] >= thr (when use_thr), write the kept points and colours densely IN INPUT ORDER.
//   count  one block per 2048 points: kept count
//   scan   single block over the block counts (exclusive offsets + total)
//   write  every block recomputes its predicate, ranks its points with ballot / popc and stores
// Algorithmic bytes: 5 B/point read twice (conf + valid) + 15 B per kept point read and written.
// ---------------------------------------------------------------------------------
#define FP_THREADS 256
#define FP_PER_BLOCK 2048

__device__ __forceinline__ bool fp_keep(const float* conf, const uint8_t* valid, long long i, bool use_thr, float thr) {
    bool k = valid ? valid[i] != 0 : true;
    if (use_thr) k = k && (conf[i] >= thr);
    return k;
}

__global__ void __launch_bounds__(FP_THREADS)
filter_count_kernel(const float* __restrict__ conf, const uint8_t* __restrict__ valid, long long n, int use_thr, float thr,
                    unsigned int* __restrict__ block_counts) {
    __shared__ unsigned int wsum[FP_THREADS / 32];
    const long long base = (long long)blockIdx.x * FP_PER_BLOCK;
    unsigned int c = 0;
    for (int j = 0; j < FP_PER_BLOCK / FP_THREADS; ++j) {
        const long long i = base + j * FP_THREADS + threadIdx.x;
        c += (i < n && fp_keep(conf, valid, i, use_thr, thr)) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < FP_THREADS / 32; ++w) t += wsum[w];
        block_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
filter_scan_kernel(const unsigned int* __restrict__ counts, int n, unsigned long long* __restrict__ offsets, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long before = carry;
        for (unsigned int w = 0; w < warp; ++w) before += wsum[w];
        if (i < n) offsets[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(FP_THREADS)
filter_write_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ rgb, const float* __restrict__ conf,
                    const uint8_t* __restrict__ valid, long long n, int use_thr, float thr,
                    const unsigned long long* __restrict__ offsets, long long max_out,
                    float* __restrict__ xyz_out, uint8_t* __restrict__ rgb_out) {
    __shared__ unsigned int wsum[FP_THREADS / 32];
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * FP_PER_BLOCK;
    unsigned long long out = offsets[blockIdx.x];
    for (int j = 0; j < FP_PER_BLOCK / FP_THREADS; ++j) {          // rounds in input order
        const long long i = base + j * FP_THREADS + threadIdx.x;
        const bool k = i < n && fp_keep(conf, valid, i, use_thr, thr);
        const unsigned int m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        unsigned int before = 0, round_total = 0;
        for (unsigned int w = 0; w < FP_THREADS / 32; ++w) { if (w < warp) before += wsum[w]; round_total += wsum[w]; }
        if (k) {
            const unsigned long long o = out + before + __popc(m & ((1u << lane) - 1u));
            if ((long long)o < max_out) {
                xyz_out[3 * o] = xyz[3 * i]; xyz_out[3 * o + 1] = xyz[3 * i + 1]; xyz_out[3 * o + 2] = xyz[3 * i + 2];
                if (rgb_out) { rgb_out[3 * o] = rgb[3 * i]; rgb_out[3 * o + 1] = rgb[3 * i + 1]; rgb_out[3 * o + 2] = rgb[3 * i + 2]; }
            }
        }
        out += round_total;
        __syncthreads();
    }
}

extern "C" int da3s_filter_points(da3s_ctx* ctx, const float* xyz, const uint8_t* rgb, const float* conf, const uint8_t* valid,
                                  long long n, int use_thr, float thr, long long max_out, float* xyz_out, uint8_t* rgb_out,
                                  unsigned long long* n_out, void* stream) {
    if (!ctx || !xyz || !xyz_out || !n_out || n < 0 || max_out < 0 || (use_thr && !conf) || ((rgb == nullptr) != (rgb_out == nullptr)))
        return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(n_out, 0, 8, st)); return DA3S_OK; }
    const long long nb = (n + FP_PER_BLOCK - 1) / FP_PER_BLOCK;
    if (nb > 2147483647LL) return DA3S_EINVAL;
    size_t save_top = ctx->ws_top;
    WS_ALLOC(ctx, unsigned int, counts, (size_t)nb);
    WS_ALLOC(ctx, unsigned long long, offsets, (size_t)nb);
    filter_count_kernel<<<(unsigned int)nb, FP_THREADS, 0, st>>>(conf, valid, n, use_thr, thr, counts);
    DA3S_LAUNCH_CHECK(ctx);
    filter_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)nb, offsets, n_out);
    DA3S_LAUNCH_CHECK(ctx);
    filter_write_kernel<<<(unsigned int)nb, FP_THREADS, 0, st>>>(xyz, rgb, conf, valid, n, use_thr, thr, offsets, max_out, xyz_out, rgb_out);
    DA3S_LAUNCH_CHECK(ctx);
    ctx->ws_top = save_top;     // consumed in stream order
    return DA3S_OK;
}
