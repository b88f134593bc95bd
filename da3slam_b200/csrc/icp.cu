// icp.cu — nearest-neighbour registration of two unordered clouds:
//   mode DA3S_ICP_SIM3   the KD-tree Umeyama loop of align_geometry.py:84-140
//                        (warp source, nearest target point, keep d^2 < thr^2, < 20 inliers -> stop,
//                         _umeyama_sim3 on (warped source, matched target), compose)
//   mode DA3S_ICP_RIGID  Open3D point-to-point ICP as called at align_geometry.py:29-45 and
//                        utils/align_geometry_single.py:146-160 (un-vendored; restated from its published
//                        algorithm: correspondences within max distance, Kabsch update without scale,
//                        stop when |d fitness| and |d inlier_rmse| < 1e-6 or after max_iteration)
//
// A point can only be an inlier if a target point lies within `thr`, so the exact nearest neighbour is
// needed only inside that radius: the target goes into a uniform grid (hash table cell -> linked list of
// points, cell size c <= thr chosen by the caller from the point density) and a query walks the shells of
// cells around its own cell outwards; it stops as soon as the best distance found cannot be beaten by the
// next shell ((r-1) c >= best) or the shell lies beyond thr.  Equal distances resolve to the lowest target
// index (a KD-tree's choice among exact ties is implementation defined).
// One launch per iteration, no host round trip: the last block of a launch (ticket) reduces the
// float64 moments of the inlier pairs, solves, composes and decides whether later launches still work.
#include "common.cuh"
#include "sim3_math.cuh"

#define ICP_EMPTY 0xFFFFFFFFFFFFFFFFull
#define ICP_BIAS (1 << 20)
#define ICP_THREADS 256
#define ICP_MAX_PROBE 1024

struct IcpState {
    double s, R[9], t[3];
    double fitness, rmse;           // of the last evaluation (RIGID)
    double n_inliers;
    int iters, done, status, evaluated;
};

struct IcpArgs {
    const void* src; const void* dst; int points_f64;
    long long n, m;
    double thr, thr2, inv_h, cell;
    int rings;                      // shells to visit at most: ceil(thr / cell)
    unsigned long long* keys; int* head; int* next; long long slots;
    IcpState* state; double* partials; unsigned int* ticket; unsigned long long* n_src_valid;
    int mode, max_iterations, min_inliers;
    double* row;
};

__device__ __forceinline__ unsigned long long icp_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__device__ __forceinline__ void icp_load(const void* base, int f64, long long i, double* p) {
    if (f64) { const double* q = (const double*)base + 3 * i; p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; }
    else { const float* q = (const float*)base + 3 * i; p[0] = (double)q[0]; p[1] = (double)q[1]; p[2] = (double)q[2]; }
}

__device__ __forceinline__ bool icp_cell(const double* p, double inv_h, long long* c) {
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double q = floor(p[a] * inv_h);
        ok = ok && (fabs(q) < (double)(ICP_BIAS - 2));
        c[a] = ok ? (long long)q : 0;
    }
    return ok;
}

__device__ __forceinline__ unsigned long long icp_key(long long cx, long long cy, long long cz) {
    return ((unsigned long long)(cx + ICP_BIAS) << 42) | ((unsigned long long)(cy + ICP_BIAS) << 21) | (unsigned long long)(cz + ICP_BIAS);
}

__global__ void icp_clear_kernel(unsigned long long* keys, int* head, long long slots, IcpState* st, unsigned int* ticket,
                                 unsigned long long* n_src_valid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < slots) { keys[i] = ICP_EMPTY; head[i] = -1; }
    if (i == 0) {
        st->s = 1.0;
        for (int k = 0; k < 9; ++k) st->R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        st->t[0] = st->t[1] = st->t[2] = 0.0;
        st->fitness = 0.0; st->rmse = 0.0; st->n_inliers = 0.0;
        st->iters = 0; st->done = 0; st->status = 0; st->evaluated = 0;
        *ticket = 0u; *n_src_valid = 0ull;
    }
}

// target points -> grid; points with a non-finite coordinate (align_geometry.py:95) or outside +-2^20 cells never match
__global__ void __launch_bounds__(ICP_THREADS)
icp_build_kernel(IcpArgs a) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.m) return;
    double p[3];
    icp_load(a.dst, a.points_f64, j, p);
    a.next[j] = -1;
    if (!(fabs(p[0]) < INFINITY && fabs(p[1]) < INFINITY && fabs(p[2]) < INFINITY)) return;
    long long c[3];
    if (!icp_cell(p, a.inv_h, c)) return;
    const unsigned long long key = icp_key(c[0], c[1], c[2]);
    unsigned long long slot = icp_hash(key) & (unsigned long long)(a.slots - 1);
    for (int probe = 0; probe < ICP_MAX_PROBE; ++probe) {
        unsigned long long cur = *((volatile unsigned long long*)(a.keys + slot));
        if (cur == ICP_EMPTY) {
            cur = atomicCAS(a.keys + slot, ICP_EMPTY, key);
            if (cur == ICP_EMPTY) cur = key;
        }
        if (cur == key) {
            a.next[j] = atomicExch(a.head + slot, (int)j);
            return;
        }
        slot = (slot + 1) & (unsigned long long)(a.slots - 1);
    }
}

__device__ void icp_write_row(const IcpArgs& a, const IcpState& st) {
    double* r = a.row;
    r[DA3S_ROW_S] = st.s;
    for (int k = 0; k < 9; ++k) r[DA3S_ROW_R + k] = st.R[k];
    for (int k = 0; k < 3; ++k) r[DA3S_ROW_T + k] = st.t[k];
    r[DA3S_ROW_NVALID] = st.n_inliers;
    r[DA3S_ROW_ITERS] = (double)st.iters;
    r[DA3S_ROW_STATUS] = (double)st.status;
}

// one evaluation under the current transform (+ the update that follows it)
__global__ void __launch_bounds__(ICP_THREADS)
icp_iter_kernel(IcpArgs a, int launch_index) {
    __shared__ double red[ICP_THREADS / 32][MOM_LEN];
    __shared__ IcpState cur;
    __shared__ bool is_last;
    if (threadIdx.x == 0) cur = *a.state;                           // written by the previous launch
    __syncthreads();
    if (cur.done) return;                                            // block-uniform (written before this launch started)
    double acc[MOM_LEN];
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) acc[k] = 0.0;
    unsigned long long valid = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        double x[3];
        icp_load(a.src, a.points_f64, i, x);
        if (!(fabs(x[0]) < INFINITY && fabs(x[1]) < INFINITY && fabs(x[2]) < INFINITY)) continue;   // align_geometry.py:94
        ++valid;
        double r[3], w[3];
        mat3_vec(cur.R, x, r);
        for (int k = 0; k < 3; ++k) w[k] = cur.s * r[k] + cur.t[k];  // (s (R x)) + t, align_geometry.py:109
        long long c[3];
        if (!icp_cell(w, a.inv_h, c)) continue;
        double best = INFINITY;
        int best_j = -1;
        auto visit = [&](long long cx, long long cy, long long cz) {
            const unsigned long long key = icp_key(cx, cy, cz);
            unsigned long long slot = icp_hash(key) & (unsigned long long)(a.slots - 1);
            int h = -1;
            for (int probe = 0; probe < ICP_MAX_PROBE; ++probe) {
                const unsigned long long k2 = a.keys[slot];
                if (k2 == key) { h = a.head[slot]; break; }
                if (k2 == ICP_EMPTY) break;
                slot = (slot + 1) & (unsigned long long)(a.slots - 1);
            }
            for (int j = h; j >= 0; j = a.next[j]) {
                double q[3];
                icp_load(a.dst, a.points_f64, j, q);
                const double d0 = w[0] - q[0], d1 = w[1] - q[1], d2 = w[2] - q[2];
                const double dd = d0 * d0 + d1 * d1 + d2 * d2;
                if (dd < best || (dd == best && j < best_j)) { best = dd; best_j = j; }
            }
        };
        for (int r2 = 0; r2 <= a.rings; ++r2) {
            if (r2 >= 2) {                                           // every point of shell r2 is at least (r2 - 1) c away
                const double lower = (double)(r2 - 1) * a.cell;
                if (lower * lower > best || lower >= a.thr) break;
            }
            for (int dz = -r2; dz <= r2; ++dz)
                for (int dy = -r2; dy <= r2; ++dy) {
                    const bool face = (dz == -r2 || dz == r2 || dy == -r2 || dy == r2);
                    if (face) { for (int dx = -r2; dx <= r2; ++dx) visit(c[0] + dx, c[1] + dy, c[2] + dz); }
                    else { visit(c[0] - r2, c[1] + dy, c[2] + dz); if (r2) visit(c[0] + r2, c[1] + dy, c[2] + dz); }
                }
        }
        if (best_j < 0 || !(best < a.thr2)) continue;                // strict, align_geometry.py:123
        double y[3];
        icp_load(a.dst, a.points_f64, best_j, y);
        acc[MOM_S0] += 1.0;
        acc[MOM_N] += 1.0;
        acc[MOM_SR] += best;                                         // sum of squared distances (inlier_rmse)
#pragma unroll
        for (int k = 0; k < 3; ++k) { acc[MOM_SX + k] += w[k]; acc[MOM_SY + k] += y[k]; }
#pragma unroll
        for (int i2 = 0; i2 < 3; ++i2)
#pragma unroll
            for (int j2 = 0; j2 < 3; ++j2) acc[MOM_SYX + 3 * i2 + j2] += y[i2] * w[j2];
        acc[MOM_SXX + 0] += w[0] * w[0]; acc[MOM_SXX + 1] += w[0] * w[1]; acc[MOM_SXX + 2] += w[0] * w[2];
        acc[MOM_SXX + 3] += w[1] * w[1]; acc[MOM_SXX + 4] += w[1] * w[2]; acc[MOM_SXX + 5] += w[2] * w[2];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < MOM_LEN; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    unsigned long long wv = valid;
    for (int o = 16; o > 0; o >>= 1) wv += __shfl_down_sync(0xffffffffu, wv, o);
    if (lane == 0 && wv && launch_index == 0) atomicAdd(a.n_src_valid, wv);
    __syncthreads();
    double* prow = a.partials + (size_t)blockIdx.x * MOM_LEN;
    if (threadIdx.x < MOM_LEN) {
        double v = 0.0;
        for (int w2 = 0; w2 < ICP_THREADS / 32; ++w2) v += red[w2][threadIdx.x];
        prow[threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double tot[MOM_LEN];
    if (threadIdx.x < MOM_LEN) {
        double v = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(a.partials + (size_t)b * MOM_LEN + threadIdx.x);   // block order: deterministic
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    *a.ticket = 0u;
    IcpState st = cur;
    double m[MOM_LEN];
    for (int k = 0; k < MOM_LEN; ++k) m[k] = tot[k];
    m[MOM_WMAX] = 1.0;
    const double n_inl = m[MOM_N];
    const double n_src = (double)*((volatile unsigned long long*)a.n_src_valid);
    st.n_inliers = n_inl;
    bool update = true;
    if (a.mode == DA3S_ICP_SIM3) {
        if (n_inl < (double)a.min_inliers) { st.done = 1; update = false; if (st.iters == 0) st.status = 1; }     // :124 break
    } else {
        const double fitness = n_src > 0 ? n_inl / n_src : 0.0;
        const double rmse = n_inl > 0 ? sqrt(m[MOM_SR] / n_inl) : 0.0;
        if (st.evaluated && fabs(st.fitness - fitness) < 1e-6 && fabs(st.rmse - rmse) < 1e-6) { st.done = 1; update = false; }
        st.fitness = fitness; st.rmse = rmse; st.evaluated = 1;
        if (n_inl < 1.0) { st.done = 1; update = false; if (st.iters == 0) st.status = 1; }                        // no correspondence
        if (st.iters >= a.max_iterations) { st.done = 1; update = false; }
    }
    if (update) {
        double su, Ru[9], tu[3];
        umeyama_from_moments(m, 1.0, SOLVE_MEAN, &su, Ru, tu);       // align_geometry.py:59-82 on (warped source, matched target)
        if (a.mode == DA3S_ICP_RIGID) {                              // Kabsch without scale: same rotation, t = mu_y - R mu_x
            su = 1.0;
            double mx[3], my[3], Rm[3];
            for (int k = 0; k < 3; ++k) { mx[k] = m[MOM_SX + k] / n_inl; my[k] = m[MOM_SY + k] / n_inl; }
            mat3_vec(Ru, mx, Rm);
            for (int k = 0; k < 3; ++k) tu[k] = my[k] - Rm[k];
        }
        // compose (:136-138): s' = s_u s;  R' = R_u R;  t' = s_u (R_u t) + t_u
        double R2[9], Rt[3];
        mat3_mul(Ru, st.R, R2);
        mat3_vec(Ru, st.t, Rt);
        for (int k = 0; k < 9; ++k) st.R[k] = R2[k];
        for (int k = 0; k < 3; ++k) st.t[k] = su * Rt[k] + tu[k];
        st.s = su * st.s;
        st.iters += 1;
        if (a.mode == DA3S_ICP_SIM3 && st.iters >= a.max_iterations) st.done = 1;
    }
    *a.state = st;
    icp_write_row(a, st);
}

extern "C" int da3s_icp_points(da3s_ctx* ctx, const void* src, long long n_src, const void* dst, long long n_dst, int points_f64,
                               int mode, double threshold, double cell_size, int max_iterations, double* sim3_row, void* stream) {
    if (!ctx || !src || !dst || !sim3_row || n_src <= 0 || n_dst <= 0 || !(threshold > 0.0) || max_iterations <= 0) return DA3S_EINVAL;
    if (mode != DA3S_ICP_SIM3 && mode != DA3S_ICP_RIGID) return DA3S_EINVAL;
    if (n_dst > 0x7FFFFFFFll) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ws_reset(ctx);
    long long slots = 1024;
    while (slots < 2 * n_dst) slots <<= 1;
    long long blocks = (n_src + ICP_THREADS - 1) / ICP_THREADS, cap = (long long)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    WS_ALLOC(ctx, unsigned long long, keys, slots);
    WS_ALLOC(ctx, int, head, slots);
    WS_ALLOC(ctx, int, next, n_dst);
    WS_ALLOC(ctx, IcpState, state, 1);
    WS_ALLOC(ctx, double, partials, (size_t)blocks * MOM_LEN);
    WS_ALLOC(ctx, unsigned int, ticket, 2);
    WS_ALLOC(ctx, unsigned long long, n_valid, 1);
    IcpArgs a;
    a.src = src; a.dst = dst; a.points_f64 = points_f64; a.n = n_src; a.m = n_dst;
    double cell = (cell_size > 0.0 && cell_size < threshold) ? cell_size : threshold;
    if (threshold / cell > 64.0) cell = threshold / 64.0;            // bounds the shell walk of a query without neighbours
    a.thr = threshold; a.thr2 = threshold * threshold; a.cell = cell; a.inv_h = 1.0 / cell;
    a.rings = (int)ceil(threshold / cell);
    a.keys = keys; a.head = head; a.next = next; a.slots = slots;
    a.state = state; a.partials = partials; a.ticket = ticket; a.n_src_valid = n_valid;
    a.mode = mode; a.max_iterations = max_iterations; a.min_inliers = 20;
    a.row = sim3_row;
    icp_clear_kernel<<<(unsigned int)((slots + 255) / 256), 256, 0, st>>>(keys, head, slots, state, ticket, n_valid);
    DA3S_LAUNCH_CHECK(ctx);
    icp_build_kernel<<<(unsigned int)((n_dst + ICP_THREADS - 1) / ICP_THREADS), ICP_THREADS, 0, st>>>(a);
    DA3S_LAUNCH_CHECK(ctx);
    // SIM3: max_iterations evaluations, each followed by its update; RIGID: one more evaluation closes the last update
    const int launches = mode == DA3S_ICP_SIM3 ? max_iterations : max_iterations + 1;
    for (int it = 0; it < launches; ++it) {
        icp_iter_kernel<<<(unsigned int)blocks, ICP_THREADS, 0, st>>>(a, it);
        DA3S_LAUNCH_CHECK(ctx);
    }
    return DA3S_OK;
}
