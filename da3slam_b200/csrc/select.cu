// select.cu — exact, batched order statistics (median / percentile) of float32 data.
//
// Segmented MSD radix select, 3 passes over the keys (11 + 11 + 10 bits).  Each pass is
// ONE kernel: every block histograms its slice of its segment in shared memory
// (warp-aggregated atomics), merges into the segment's global histogram, and the last
// block to finish (ticket) scans the histogram, narrows the (prefix, rank) state of the
// segment's two queries and clears the histogram for the next pass.  No host round trip.
//
// The finishing arithmetic reproduces numpy >= 2 on float32 input bit for bit
// (SURVEY.md appendix A): median of an even count = (a + b) / 2 in float32; percentile =
// float32 virtual index (n-1) * (p / 100f), float32 lerp with the gamma >= 0.5 branch.
//
// Algorithmic bytes: 4 B x 3 passes per element (16 B x 3 for the depth-ratio kind).
#include "common.cuh"

#define SEL_THREADS 256
#define SEL_ITEMS 16                       // elements per thread per block (4 float4)
#define SEL_BINS 2048

struct SelState {
    unsigned int prefix[2];
    long long rank[2];                     // rank of the wanted element inside the current prefix class
    long long n_valid;
    float gamma;
    int empty;
};

__device__ __forceinline__ int sel_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int sel_bits(int pass) { return pass == 2 ? 10 : 11; }

// key of element i of a segment; returns false when the element does not take part
__device__ __forceinline__ bool sel_key(const da3s_select_seg& s, long long i, unsigned int& key) {
    float a = s.a[i];
    if (s.kind == DA3S_SEL_VALUES) { key = f32_to_key(a); return true; }
    if (s.kind == DA3S_SEL_POSITIVE) { key = f32_to_key(a); return a > 0.0f; }
    float b = s.b[i];
    bool ok = (a > s.eps) && (b > s.eps) && is_finite_f(a) && is_finite_f(b);
    if (s.ca && s.cb) ok = ok && (s.ca[i] > s.conf_th) && (s.cb[i] > s.conf_th);
    key = f32_to_key(__fdiv_rn(a, b));     // align_geometry.py:329: float32 / float32
    return ok;
}

__device__ __forceinline__ void hist_add(unsigned int* hist, unsigned int digit, bool pred) {
    // warp-aggregated shared-memory atomic: one atomic per distinct digit in the warp
    unsigned int act = __ballot_sync(0xffffffffu, pred);
    if (!pred) return;
    unsigned int peers = __match_any_sync(act, digit);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
}

template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS)
select_pass_kernel(const da3s_select_seg* __restrict__ segs, SelState* state,
                   unsigned int* ghist /* [n_segs][2][SEL_BINS] */,
                   unsigned int* tickets, da3s_select_out* out) {
    __shared__ unsigned int hist[2][SEL_BINS];
    __shared__ bool is_last;
    __shared__ long long scan_tot[SEL_THREADS];
    __shared__ unsigned int new_prefix[2];
    __shared__ long long new_rank[2];
    __shared__ long long sh_rank[2];
    __shared__ int sh_empty;
    const int seg_id = blockIdx.y;
    const da3s_select_seg seg = segs[seg_id];
    SelState st;
    if (PASS > 0) st = state[seg_id];
    const bool two = PASS > 0 && (st.prefix[0] != st.prefix[1]);       // queries diverged
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += SEL_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();

    const int shift = sel_shift(PASS), bits = sel_bits(PASS);
    const unsigned int dmask = (1u << bits) - 1u;
    const long long base = (long long)blockIdx.x * (SEL_THREADS * SEL_ITEMS);
    const bool skip = (PASS > 0 && st.empty);
    if (!skip && base < seg.n) {
        const bool vec = (seg.kind != DA3S_SEL_RATIO) && aligned16(seg.a);
        unsigned int keys[SEL_ITEMS];
        unsigned int okmask = 0;
        // all loads first (4 x 128-bit in flight per thread), then the histogram updates
        if (vec && base + (long long)SEL_THREADS * SEL_ITEMS <= seg.n) {
            float4 v[SEL_ITEMS / 4];
#pragma unroll
            for (int it = 0; it < SEL_ITEMS / 4; ++it)
                v[it] = ldg_stream(reinterpret_cast<const float4*>(seg.a + base + ((long long)it * SEL_THREADS + threadIdx.x) * 4));
#pragma unroll
            for (int it = 0; it < SEL_ITEMS / 4; ++it) {
                const float f[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    keys[it * 4 + j] = f32_to_key(f[j]);
                    if (seg.kind == DA3S_SEL_VALUES || f[j] > 0.0f) okmask |= 1u << (it * 4 + j);
                }
            }
        } else {
#pragma unroll
            for (int it = 0; it < SEL_ITEMS / 4; ++it)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long i = base + ((long long)it * SEL_THREADS + threadIdx.x) * 4 + j;
                    keys[it * 4 + j] = 0;
                    if (i < seg.n && sel_key(seg, i, keys[it * 4 + j])) okmask |= 1u << (it * 4 + j);
                }
        }
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {
            const unsigned int digit = (keys[e] >> shift) & dmask;
            const bool ok = (okmask >> e) & 1u;
            if (PASS == 0) {
                hist_add(hist[0], digit, ok);                    // every element counts: aggregate per warp
            } else {
                // only the elements inside a query's prefix class count (a small fraction): plain atomics
                const unsigned int hi = keys[e] >> (shift + bits);
                if (ok && hi == st.prefix[0]) atomicAdd(&hist[0][digit], 1u);
                if (two && ok && hi == st.prefix[1]) atomicAdd(&hist[1][digit], 1u);
            }
        }
    }
    __syncthreads();
    unsigned int* gh = ghist + (size_t)seg_id * 2 * SEL_BINS;
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += SEL_THREADS) {
        unsigned int c = (&hist[0][0])[i];
        if (c) atomicAdd(&gh[i], c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&tickets[seg_id], 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- last block of this segment: narrow the two queries ----
    const int nb = 1 << bits;
    const int per = SEL_BINS / SEL_THREADS;                    // 8 bins per thread
    for (int q = 0; q < 2; ++q) {
        const unsigned int* src = gh + ((PASS > 0 && two) ? q : 0) * SEL_BINS;
        long long loc[per], tot = 0;
#pragma unroll
        for (int k = 0; k < per; ++k) {
            int b = threadIdx.x * per + k;
            loc[k] = b < nb ? (long long)__ldcg(&src[b]) : 0;
            tot += loc[k];
        }
        scan_tot[threadIdx.x] = tot;
        __syncthreads();
        if (threadIdx.x == 0) {                                // 256-entry serial exclusive scan
            long long run = 0;
            for (int i = 0; i < SEL_THREADS; ++i) { long long v = scan_tot[i]; scan_tot[i] = run; run += v; }
            if (PASS == 0 && q == 0) {
                // ranks from the count (total = number of participating elements)
                st.n_valid = run; st.empty = (run == 0); st.gamma = 0.0f;
                st.prefix[0] = st.prefix[1] = 0;
                long long n = run;
                if (n > 0) {
                    if (seg.stat == DA3S_SEL_MEDIAN) { st.rank[0] = (n - 1) / 2; st.rank[1] = n / 2; }
                    else {
                        // numpy: q = p / float32(100); virt = (n-1) * q in float32; floor; clamp
                        float qf = __fdiv_rn(seg.percent, 100.0f);
                        float virt = __fmul_rn((float)(n - 1), qf);
                        float prev = floorf(virt);
                        if (virt >= (float)(n - 1)) { st.rank[0] = st.rank[1] = n - 1; }
                        else if (virt < 0.0f) { st.rank[0] = st.rank[1] = 0; }
                        else { st.rank[0] = (long long)prev; st.rank[1] = (long long)prev + 1; }
                        st.gamma = __fsub_rn(virt, prev);
                    }
                } else { st.rank[0] = st.rank[1] = 0; }
                state[seg_id].n_valid = st.n_valid; state[seg_id].empty = st.empty; state[seg_id].gamma = st.gamma;
                sh_rank[0] = st.rank[0]; sh_rank[1] = st.rank[1]; sh_empty = st.empty;
            }
        }
        __syncthreads();
        if (PASS == 0) { st.rank[0] = sh_rank[0]; st.rank[1] = sh_rank[1]; st.empty = sh_empty; }
        __syncthreads();
        const long long want = st.rank[q];
        long long run = scan_tot[threadIdx.x];
        if (!st.empty) {
#pragma unroll
            for (int k = 0; k < per; ++k) {
                if (want >= run && want < run + loc[k]) {      // exactly one (thread, k) matches
                    unsigned int b = threadIdx.x * per + k;
                    unsigned int old = PASS == 0 ? 0u : st.prefix[q];
                    new_prefix[q] = (old << bits) | b;
                    new_rank[q] = want - run;
                }
                run += loc[k];
            }
        }
        __syncthreads();
    }
    // clear the global histogram and the ticket for the next pass / next call
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += SEL_THREADS) gh[i] = 0;
    if (threadIdx.x == 0) {
        tickets[seg_id] = 0;
        if (!st.empty) {                                       // thread 0 alone publishes the narrowed state
            state[seg_id].prefix[0] = new_prefix[0]; state[seg_id].prefix[1] = new_prefix[1];
            state[seg_id].rank[0] = new_rank[0];     state[seg_id].rank[1] = new_rank[1];
        }
    }
    if (PASS == 2 && threadIdx.x == 0) {
        SelState f;
        f.prefix[0] = new_prefix[0]; f.prefix[1] = new_prefix[1];
        f.n_valid = st.n_valid; f.gamma = st.gamma; f.empty = st.empty;
        da3s_select_out o;
        o.n_valid = f.n_valid;
        o.gamma = f.gamma;
        if (f.empty) {
            o.lo = o.hi = o.value = __int_as_float(0x7fc00000);
        } else {
            float a = key_to_f32(f.prefix[0]), b = key_to_f32(f.prefix[1]);
            o.lo = a; o.hi = b;
            if (seg.stat == DA3S_SEL_MEDIAN) {
                // np.median: odd -> the element; even -> np.mean of the two = (a+b)/2 in float32
                o.value = (f.n_valid & 1) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
            } else {
                // numpy _lerp in float32
                float diff = __fsub_rn(b, a);
                float v = __fadd_rn(a, __fmul_rn(diff, f.gamma));
                if (f.gamma >= 0.5f) v = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, f.gamma)));
                o.value = v;
            }
        }
        out[seg_id] = o;
    }
}

// internal entry used by pair_align.cu as well
int da3s_select_impl(da3s_ctx* ctx, const da3s_select_seg* segs, int n_segs, long long max_n,
                     da3s_select_out* out, cudaStream_t st) {
    if (n_segs <= 0) return DA3S_OK;
    if (n_segs > 65535) return DA3S_EINVAL;
    size_t save_top = ctx->ws_top;
    WS_ALLOC(ctx, SelState, state, n_segs);
    WS_ALLOC(ctx, unsigned int, ghist, (size_t)n_segs * 2 * SEL_BINS);
    WS_ALLOC(ctx, unsigned int, tickets, n_segs);
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(ghist, 0, sizeof(unsigned int) * (size_t)n_segs * 2 * SEL_BINS, st));
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(tickets, 0, sizeof(unsigned int) * n_segs, st));
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(SelState) * n_segs, st));
    long long per_block = (long long)SEL_THREADS * SEL_ITEMS;
    long long bx = (max_n + per_block - 1) / per_block;
    if (bx < 1) bx = 1;
    if (bx > 2147483647LL) return DA3S_EINVAL;
    dim3 grid((unsigned int)bx, n_segs);
    select_pass_kernel<0><<<grid, SEL_THREADS, 0, st>>>(segs, state, ghist, tickets, out);
    DA3S_LAUNCH_CHECK(ctx);
    select_pass_kernel<1><<<grid, SEL_THREADS, 0, st>>>(segs, state, ghist, tickets, out);
    DA3S_LAUNCH_CHECK(ctx);
    select_pass_kernel<2><<<grid, SEL_THREADS, 0, st>>>(segs, state, ghist, tickets, out);
    DA3S_LAUNCH_CHECK(ctx);
    ctx->ws_top = save_top;     // scratch is free again once the three passes are queued (stream order)
    return DA3S_OK;
}

extern "C" int da3s_select(da3s_ctx* ctx, const da3s_select_seg* segs, int n_segs, long long max_n,
                           da3s_select_out* out, void* stream) {
    if (!ctx || !segs || !out || n_segs < 0 || max_n < 0) return DA3S_EINVAL;
    ws_reset(ctx);
    return da3s_select_impl(ctx, segs, n_segs, max_n, out, (cudaStream_t)stream);
}
