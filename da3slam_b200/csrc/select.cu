// select.cu — exact, batched order statistics (median / percentile) of float32 data.
//
// One full pass over the data instead of three:
//   A  sample    one block per segment takes a 4096-element strided sample into shared memory and
//                radix-selects (in shared memory) the two sample keys [lo, hi] that bracket the
//                wanted rank (+-5 sigma of the sampling error).  Segments of <= 4096 elements are
//                answered here.
//   B  count     ONE streaming pass: per element 2 compares; counts elements below / equal to
//                the pivots and appends the few (~8 %) strictly inside the bracket to a
//                candidate buffer (block-aggregated append).  The last block of a segment
//                (ticket) computes the exact ranks from the exact count and maps them into the
//                candidates — or, if the bracket missed or the buffer overflowed (heavy ties,
//                adversarial data), redirects the next stage to the full segment.  Exactness
//                never depends on the sample; only the amount of work does.
//   C  resolve   one block per segment: 4-pass MSD radix select (8 bits each, shared-memory
//                histograms) over the candidates (or, on fallback, over the whole segment).
//                The end reproduces numpy >= 2's float32 finishing arithmetic bit for bit
//                (SURVEY.md appendix A): median of an even count = (a + b) / 2; percentile =
//                float32 virtual index (n-1) * (p / 100f), lerp with the gamma >= 0.5 branch.
// Everything is device-side: no host round trip, fixed launch sequence (3 kernels).
//
// Algorithmic bytes: 4 B per element (16 B for the depth-ratio kind) read once.
#include "common.cuh"

#define SEL_THREADS 256
#define SEL_ITEMS 16                       // elements per thread per block iteration (4 float4)
#define SEL_BINS 2048
#define SEL_SAMPLE 4096
#define SEL_KIND_KEYS 100                  // internal: the segment already holds ordered uint32 keys

struct SelWork {                           // per segment, device resident
    const void* data;                      // stage C input: candidate keys (or the original `a` on fallback)
    long long n;                           // elements in `data`
    int kind;                              // SEL_KIND_KEYS or the original kind (fallback)
    int done;                              // 1: answered by the sample stage; 2: no valid element
    unsigned int lo_key, hi_key;           // bracket from the sample
    long long rank[2];                     // ranks to find inside `data`
    int resolved[2];                       // the order statistic is already known (a pivot)
    unsigned int resolved_key[2];
    long long n_valid;
    float gamma;
    int overflow;
    // stage B counters
    unsigned long long c_valid, c_less, c_eqlo, c_eqhi, c_cand;
    // multi-block stage C state
    unsigned int prefix[2];
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

// ranks of the two order statistics numpy uses, and the interpolation weight
__device__ __forceinline__ void sel_ranks(long long n, int stat, float percent, long long& k0, long long& k1, float& gamma) {
    gamma = 0.0f;
    if (n <= 0) { k0 = k1 = 0; return; }
    if (stat == DA3S_SEL_MEDIAN) { k0 = (n - 1) / 2; k1 = n / 2; return; }
    // numpy: q = p / float32(100); virt = (n-1) * q in float32; floor; clamp
    float qf = __fdiv_rn(percent, 100.0f);
    float virt = __fmul_rn((float)(n - 1), qf);
    float prev = floorf(virt);
    if (virt >= (float)(n - 1)) { k0 = k1 = n - 1; }
    else if (virt < 0.0f) { k0 = k1 = 0; }
    else { k0 = (long long)prev; k1 = (long long)prev + 1; }
    gamma = __fsub_rn(virt, prev);
}

__device__ __forceinline__ void sel_finish(const da3s_select_seg& seg, long long n_valid, float gamma, bool empty,
                                           unsigned int key0, unsigned int key1, da3s_select_out* out) {
    da3s_select_out o;
    o.n_valid = n_valid;
    o.gamma = gamma;
    if (empty) {
        o.lo = o.hi = o.value = __int_as_float(0x7fc00000);
    } else {
        float a = key_to_f32(key0), b = key_to_f32(key1);
        o.lo = a; o.hi = b;
        if (seg.stat == DA3S_SEL_MEDIAN) {
            // np.median: odd -> the element; even -> np.mean of the two = (a+b)/2 in float32
            o.value = (n_valid & 1) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
        } else {
            // numpy _lerp in float32
            float diff = __fsub_rn(b, a);
            float v = __fadd_rn(a, __fmul_rn(diff, gamma));
            if (gamma >= 0.5f) v = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
            o.value = v;
        }
    }
    *out = o;
}

// ---------------------------------------------------------------------------------
// block-level exact selection of TWO ranks in one sweep per digit: the keys of rank r[0] and r[1]
// among the keys produced by `get(i, k)`, i in [0, n) (get returns false for elements that do not
// take part).  MSD radix, 8 bits per pass, histograms in shared memory; keys are known to be
// < 2^(32-lz), so only ceil((32-lz)/8) passes run and the leading pass already spreads over all
// 256 bins.  While both ranks share a prefix one histogram serves both.  Every thread of the
// block must call it.
// ---------------------------------------------------------------------------------
struct SelScratch { unsigned int hist[2][256]; unsigned int digit[2]; unsigned int rest[2]; };

template <typename Get>
__device__ void block_select2(long long n, long long r0, long long r1, Get get, SelScratch& sc, int lz,
                              unsigned int& key0, unsigned int& key1) {
    unsigned int prefix0 = 0, prefix1 = 0;
    const int passes = (32 - lz + 7) / 8;
    const unsigned int lane = threadIdx.x & 31;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 24 - 8 * pass;
        const bool same = prefix0 == prefix1;
        for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sc.hist[0][0])[i] = 0;
        __syncthreads();
        auto visit = [&](unsigned int k, bool ok) {                  // whole warps call it
            const unsigned int kk = k << lz;
            const unsigned int top = pass == 0 ? 0u : (kk >> (shift + 8));
            const unsigned int digit = (kk >> shift) & 255u;
            const bool in0 = ok && top == prefix0;
            int uniform;
            __match_all_sync(0xffffffffu, in0 ? digit : 0xFFFFu, &uniform);
            if (uniform) { if (in0 && lane == 0) atomicAdd(&sc.hist[0][digit], 32u); }   // typical for un-bracketed keys in the leading pass
            else if (in0) atomicAdd(&sc.hist[0][digit], 1u);
            if (!same && ok && top == prefix1) atomicAdd(&sc.hist[1][digit], 1u);
        };
        const long long step = (long long)blockDim.x * 4;
        const long long n_round = (n + step - 1) / step * step;
        for (long long i0 = threadIdx.x; i0 < n_round; i0 += step) {  // 4 independent loads in flight per thread
            unsigned int k[4]; bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const long long i = i0 + (long long)j * blockDim.x; k[j] = 0; ok[j] = i < n && get(i, k[j]); }
#pragma unroll
            for (int j = 0; j < 4; ++j) visit(k[j], ok[j]);
        }
        __syncthreads();
        if (threadIdx.x < 64) {                                      // warp q locates rank q: 8 bins per lane, exclusive scan
            const int q = threadIdx.x >> 5;
            const unsigned int* h = sc.hist[(q == 1 && !same) ? 1 : 0];
            long long rank = q ? r1 : r0;
            unsigned int loc[8], tot = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { loc[b] = h[lane * 8 + b]; tot += loc[b]; }
            unsigned int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (unsigned int)o) incl += t;
            }
            long long run = (long long)(incl - tot);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (rank >= run && rank < run + (long long)loc[b]) { sc.digit[q] = lane * 8 + b; sc.rest[q] = (unsigned int)(rank - run); }
                run += loc[b];
            }
        }
        __syncthreads();
        prefix0 = (prefix0 << 8) | sc.digit[0]; r0 = (long long)sc.rest[0];
        prefix1 = (prefix1 << 8) | sc.digit[1]; r1 = (long long)sc.rest[1];
        __syncthreads();
    }
    const int drop = 32 - 8 * passes;                                // digits not visited are zero
    key0 = (drop >= 0 ? prefix0 << drop : prefix0) >> lz;
    key1 = (drop >= 0 ? prefix1 << drop : prefix1) >> lz;
}

// ---------------------------------------------------------------------------------
// A: sample, bracket
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS)
select_sample_kernel(const da3s_select_seg* __restrict__ segs, SelWork* work, da3s_select_out* out) {
    __shared__ unsigned int keys[SEL_SAMPLE];
    __shared__ SelScratch sc;
    __shared__ unsigned int n_ok;
    const int seg_id = blockIdx.x;
    const da3s_select_seg seg = segs[seg_id];
    if (threadIdx.x == 0) n_ok = 0;
    __syncthreads();
    const long long n = seg.n;
    const long long m = n < SEL_SAMPLE ? n : SEL_SAMPLE;
    unsigned int mine = 0;
    for (int j = threadIdx.x; j < SEL_SAMPLE; j += SEL_THREADS) {
        unsigned int k = 0xFFFFFFFFu;
        if (j < m) {
            const long long i = (n <= SEL_SAMPLE) ? j : (long long)(((unsigned long long)j * (unsigned long long)n) / SEL_SAMPLE);
            unsigned int kk;
            if (sel_key(seg, i, kk)) { k = kk; ++mine; }
        }
        keys[j] = k;
    }
    atomicAdd(&n_ok, mine);
    __syncthreads();
    const long long mv = (long long)n_ok;
    // valid samples are the ones != 0xFFFFFFFF; a real key 0xFFFFFFFF (a NaN pattern) among them is counted as rejected,
    // which only widens the bracket (exactness comes from stage B/C, never from the sample)
    auto get = [&](long long i, unsigned int& k) -> bool { k = keys[i]; return k != 0xFFFFFFFFu; };
    SelWork w;
    w.data = nullptr; w.n = 0; w.kind = SEL_KIND_KEYS; w.done = 0; w.lo_key = 0u; w.hi_key = 0xFFFFFFFFu;
    w.rank[0] = w.rank[1] = 0; w.resolved[0] = w.resolved[1] = 0; w.resolved_key[0] = w.resolved_key[1] = 0;
    w.prefix[0] = w.prefix[1] = 0;
    w.n_valid = 0; w.gamma = 0.0f; w.overflow = 0; w.c_valid = w.c_less = w.c_eqlo = w.c_eqhi = w.c_cand = 0ull;
    if (n <= SEL_SAMPLE) {
        // the sample is the whole segment: answer now (count the valid ones exactly, incl. NaN-pattern keys)
        long long k0, k1; float gamma;
        sel_ranks(mv, seg.stat, seg.percent, k0, k1, gamma);
        auto get_all = [&](long long i, unsigned int& k) -> bool { unsigned int kk; bool ok = i < m && sel_key(seg, i, kk); k = kk; return ok; };
        unsigned int a = 0, b = 0;
        if (mv > 0) block_select2(m, k0, k1, get_all, sc, 0, a, b);  // block-uniform
        if (threadIdx.x == 0) {
            sel_finish(seg, mv, gamma, mv == 0, a, b, out + seg_id);
            w.done = 1; w.n_valid = mv; w.gamma = gamma;
            work[seg_id] = w;
        }
        return;
    }
    if (mv > 0) {                                                // block-uniform
        // bracket the wanted quantile of the VALID elements: +-5 sigma of the binomial sampling error
        const double q = (seg.stat == DA3S_SEL_MEDIAN) ? 0.5 : fmin(fmax((double)seg.percent / 100.0, 0.0), 1.0);
        const double c = q * (double)(mv - 1);
        const double margin = 5.0 * sqrt((double)mv * q * (1.0 - q)) + 8.0;
        const long long lo_i = (long long)floor(c - margin), hi_i = (long long)ceil(c + margin);
        unsigned int klo, khi;
        block_select2(SEL_SAMPLE, lo_i > 0 ? lo_i : 0, hi_i < mv - 1 ? hi_i : mv - 1, get, sc, 0, klo, khi);
        if (lo_i > 0) w.lo_key = klo;
        if (hi_i < mv - 1) w.hi_key = khi;
    }
    if (threadIdx.x == 0) work[seg_id] = w;
}

// ---------------------------------------------------------------------------------
// B: one streaming pass — count below / at the pivots, collect the candidates.
// ~15 instructions per element: key, three predicated counters, one range test; the rare
// candidates are staged in shared memory (warp-aggregated) and flushed with one global
// atomic per block.
// ---------------------------------------------------------------------------------
#define SEL_SINGLE_MAX (1ll << 21)                   // up to here one block per segment resolves the candidates
#define SEL_CHUNKS 8                                 // chunks of 4096 elements per block: amortises the block epilogue
#define SEL_STAGE (2 * SEL_THREADS * SEL_ITEMS)      // candidate staging; flushed whenever another chunk might not fit

__global__ void __launch_bounds__(SEL_THREADS)
select_count_kernel(const da3s_select_seg* __restrict__ segs, SelWork* work, unsigned int* cand, long long cand_cap,
                    unsigned int* tickets, int chunks /* per block, <= SEL_CHUNKS */) {
    __shared__ unsigned int stage[SEL_STAGE];
    __shared__ unsigned int n_stage, stage_limit;
    __shared__ unsigned long long blk[4];
    __shared__ unsigned long long cand_base;
    __shared__ bool is_last;
    const int seg_id = blockIdx.y;
    const da3s_select_seg seg = segs[seg_id];
    SelWork* w = work + seg_id;
    if (w->done) return;                                             // block-uniform
    const unsigned int lo = w->lo_key, hi = w->hi_key;
    if (threadIdx.x < 4) blk[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) { n_stage = 0; stage_limit = SEL_STAGE; }
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31;
    unsigned int n_valid = 0, n_less = 0, n_eqlo = 0, n_eqhi = 0;
    // 16 keys of one thread.  Pass 1 touches every key but only classifies it as below / inside the closed
    // bracket [lo, hi] (bit masks, counted with popc once per 16); the few keys inside (~8 %) are then split
    // into == lo, == hi and strict candidates, and the candidates are staged with ONE warp scan + one shared
    // atomic per warp (a warp almost always holds a candidate, so per-element voting would cost more than
    // the counting).
    const unsigned int width = hi - lo;                              // lo <= hi
    auto visit16 = [&](const unsigned int (&keys)[SEL_ITEMS], unsigned int okmask) {
        unsigned int lessmask = 0, inmask = 0, lomask = 0, himask = 0;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; ++e) {                       // branch-free: bit e of each mask
            const unsigned int k = keys[e];
            lessmask |= (k < lo ? 1u : 0u) << e;
            inmask |= ((k - lo) <= width ? 1u : 0u) << e;           // unsigned: k < lo wraps above width
            lomask |= (k == lo ? 1u : 0u) << e;
            himask |= (k == hi ? 1u : 0u) << e;
        }
        lomask &= okmask;
        himask &= okmask & ~lomask;                                  // lo == hi: counted once, as == lo
        n_valid += __popc(okmask);
        n_less += __popc(lessmask & okmask);
        n_eqlo += __popc(lomask);
        n_eqhi += __popc(himask);
        const unsigned int cmask = inmask & okmask & ~lomask & ~himask;
        const unsigned int cnt = __popc(cmask);
        unsigned int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned int)o) incl += t;
        }
        const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {                                                 // warp-uniform
            unsigned int pos = 0;
            if (lane == 31) pos = atomicAdd(&n_stage, total);
            const unsigned int warp_pos = __shfl_sync(0xffffffffu, pos, 31);
            if (warp_pos + total > SEL_STAGE) {
                // the block's staging area is full (values concentrated around the wanted rank in this part of the
                // segment, e.g. sorted or smooth data): this warp's candidates go straight to the segment's
                // candidate array; everything staged below `stage_limit` is still valid
                unsigned long long gpos = 0;
                if (lane == 31) { atomicMin(&stage_limit, warp_pos); gpos = atomicAdd(&w->c_cand, (unsigned long long)total); }
                gpos = __shfl_sync(0xffffffffu, gpos, 31) + incl - cnt;
                unsigned int* dst = cand + (size_t)seg_id * cand_cap;
                bool over = false;
#pragma unroll
                for (int e = 0; e < SEL_ITEMS; ++e)
                    if ((cmask >> e) & 1u) {
                        if ((long long)gpos < cand_cap) dst[gpos] = keys[e]; else over = true;
                        ++gpos;
                    }
                if (over) w->overflow = 1;
                return;
            }
            pos = warp_pos + incl - cnt;
            // one shared-window address, then a predicated store + add per key (the generic form re-derives the
            // window base for every store)
            unsigned int saddr = (unsigned int)__cvta_generic_to_shared(stage) + 4u * pos;
#pragma unroll
            for (int e = 0; e < SEL_ITEMS; ++e)
                if ((cmask >> e) & 1u) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(keys[e]) : "memory");
                    saddr += 4u;
                }
        }
    };
    auto flush = [&]() {                                             // every thread of the block calls it
        __syncthreads();
        const unsigned int total = n_stage < stage_limit ? n_stage : stage_limit;  // reservations from stage_limit on went to global directly
        if (threadIdx.x == 0) cand_base = total ? atomicAdd(&w->c_cand, (unsigned long long)total) : 0ull;
        __syncthreads();
        if (total) {
            unsigned int* dst = cand + (size_t)seg_id * cand_cap;
            bool over = false;
            for (unsigned int i = threadIdx.x; i < total; i += SEL_THREADS) {
                const unsigned long long pos = cand_base + i;
                if ((long long)pos < cand_cap) dst[pos] = stage[i]; else over = true;
            }
            if (over) w->overflow = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) n_stage = 0;
        __syncthreads();
    };
    const long long chunk = (long long)SEL_THREADS * SEL_ITEMS;
    for (int ch = 0; ch < chunks; ++ch) {                            // block-uniform
        const long long base = ((long long)blockIdx.x * chunks + ch) * chunk;
        if (base >= seg.n) break;
        const bool vec = (seg.kind != DA3S_SEL_RATIO) && aligned16(seg.a) && base + chunk <= seg.n;
        unsigned int keys[SEL_ITEMS];
        unsigned int okmask = 0;
        if (vec) {
            float4 v[SEL_ITEMS / 4];
#pragma unroll
            for (int it = 0; it < SEL_ITEMS / 4; ++it)
                v[it] = ldg_stream(reinterpret_cast<const float4*>(seg.a + base + ((long long)it * SEL_THREADS + threadIdx.x) * 4));
            if (seg.kind == DA3S_SEL_VALUES) {                       // block-uniform: every element takes part
#pragma unroll
                for (int it = 0; it < SEL_ITEMS / 4; ++it) {
                    keys[it * 4 + 0] = f32_to_key(v[it].x); keys[it * 4 + 1] = f32_to_key(v[it].y);
                    keys[it * 4 + 2] = f32_to_key(v[it].z); keys[it * 4 + 3] = f32_to_key(v[it].w);
                }
                okmask = 0xFFFFu;
            } else {                                                 // DA3S_SEL_POSITIVE
#pragma unroll
                for (int it = 0; it < SEL_ITEMS / 4; ++it) {
                    const float f[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        keys[it * 4 + j] = f32_to_key(f[j]);
                        okmask |= (f[j] > 0.0f ? 1u : 0u) << (it * 4 + j);
                    }
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < SEL_ITEMS; ++e) {
                const long long i = base + ((long long)(e >> 2) * SEL_THREADS + threadIdx.x) * 4 + (e & 3);
                keys[e] = 0;
                if (i < seg.n && sel_key(seg, i, keys[e])) okmask |= 1u << e;
            }
        }
        visit16(keys, okmask);                                       // no block barrier between chunks: warps run ahead freely
    }
    flush();
    unsigned int r0 = __reduce_add_sync(0xffffffffu, n_valid), r1 = __reduce_add_sync(0xffffffffu, n_less);
    unsigned int r2 = __reduce_add_sync(0xffffffffu, n_eqlo), r3 = __reduce_add_sync(0xffffffffu, n_eqhi);
    if (lane == 0) {
        if (r0) atomicAdd(&blk[0], (unsigned long long)r0);
        if (r1) atomicAdd(&blk[1], (unsigned long long)r1);
        if (r2) atomicAdd(&blk[2], (unsigned long long)r2);
        if (r3) atomicAdd(&blk[3], (unsigned long long)r3);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blk[0]) atomicAdd(&w->c_valid, blk[0]);
        if (blk[1]) atomicAdd(&w->c_less, blk[1]);
        if (blk[2]) atomicAdd(&w->c_eqlo, blk[2]);
        if (blk[3]) atomicAdd(&w->c_eqhi, blk[3]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&tickets[seg_id], 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence();
    // ---- last block of the segment: exact ranks, map into the candidates (or fall back) ----
    tickets[seg_id] = 0;
    volatile SelWork* vw = w;
    const long long nv = (long long)vw->c_valid, nl = (long long)vw->c_less, ne0 = (long long)vw->c_eqlo;
    const long long ne1 = (long long)vw->c_eqhi, nc = (long long)vw->c_cand;
    long long k[2]; float gamma;
    sel_ranks(nv, seg.stat, seg.percent, k[0], k[1], gamma);
    w->n_valid = nv; w->gamma = gamma;
    if (nv == 0) { w->done = 2; return; }                            // empty: stage C's last pass writes NaN
    bool fallback = vw->overflow != 0;
    for (int q = 0; q < 2; ++q) fallback = fallback || (k[q] < nl) || (k[q] >= nl + ne0 + nc + ne1);
    if (fallback) {
        w->data = seg.a; w->n = seg.n; w->kind = seg.kind;
        for (int q = 0; q < 2; ++q) { w->rank[q] = k[q]; w->resolved[q] = 0; }
        return;
    }
    w->data = cand + (size_t)seg_id * cand_cap; w->n = nc; w->kind = SEL_KIND_KEYS;
    for (int q = 0; q < 2; ++q) {
        if (k[q] < nl + ne0) { w->resolved[q] = 1; w->resolved_key[q] = lo; }
        else if (k[q] < nl + ne0 + nc) { w->resolved[q] = 0; w->rank[q] = k[q] - nl - ne0; }
        else { w->resolved[q] = 1; w->resolved_key[q] = hi; }
    }
}

// ---------------------------------------------------------------------------------
// C (large segments): multi-block 3-pass MSD radix select (11+11+10 bits) over SelWork with a
// global histogram per segment; the last block of a segment (ticket) narrows the live ranks.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void hist_add(unsigned int* hist, unsigned int digit, bool pred) {
    // warp-aggregated shared-memory atomic: one atomic per distinct digit in the warp
    unsigned int act = __ballot_sync(0xffffffffu, pred);
    if (!pred) return;
    unsigned int peers = __match_any_sync(act, digit);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
}

template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS)
select_pass_kernel(const da3s_select_seg* __restrict__ segs, SelWork* work, unsigned int* ghist /* [n_segs][2][SEL_BINS] */,
                   unsigned int* tickets, da3s_select_out* out) {
    __shared__ unsigned int hist[2][SEL_BINS];
    __shared__ bool is_last;
    __shared__ long long scan_tot[SEL_THREADS];
    __shared__ unsigned int new_prefix[2];
    __shared__ long long new_rank[2];
    const int seg_id = blockIdx.y;
    SelWork* wp = work + seg_id;
    const int done = wp->done;
    if (done == 1) return;                                           // answered by the sample stage
    const da3s_select_seg seg = segs[seg_id];
    const bool empty = (done == 2);
    const long long n = empty ? 0 : wp->n;
    const int kind = wp->kind;
    const long long rank0 = wp->rank[0], rank1 = wp->rank[1];
    const bool res0 = wp->resolved[0] != 0, res1 = wp->resolved[1] != 0;
    const unsigned int pre0 = wp->prefix[0], pre1 = wp->prefix[1];
    // query 1 needs its own histogram when the live queries diverged, or when query 0 is already resolved
    const bool q1_own = PASS > 0 && !res1 && (res0 || pre0 != pre1);
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += SEL_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();

    const int shift = sel_shift(PASS), bits = sel_bits(PASS);
    const unsigned int dmask = (1u << bits) - 1u;
    const bool any_live = !(res0 && res1) && n > 0;
    da3s_select_seg src = seg;                                       // element source for this stage
    src.n = n; src.kind = kind;
    if (kind != SEL_KIND_KEYS) src.a = (const float*)wp->data;
    const unsigned int* kdata = (const unsigned int*)wp->data;
    const long long chunk = (long long)SEL_THREADS * SEL_ITEMS;
    if (any_live)
        for (long long base = (long long)blockIdx.x * chunk; base < n; base += (long long)gridDim.x * chunk) {   // block-uniform
            unsigned int keys[SEL_ITEMS];
            unsigned int okmask = 0;
            if (kind == SEL_KIND_KEYS) {
#pragma unroll
                for (int it = 0; it < SEL_ITEMS; ++it) {
                    const long long i = base + (long long)it * SEL_THREADS + threadIdx.x;
                    keys[it] = 0;
                    if (i < n) { keys[it] = kdata[i]; okmask |= 1u << it; }
                }
            } else {
#pragma unroll
                for (int it = 0; it < SEL_ITEMS; ++it) {
                    const long long i = base + (long long)it * SEL_THREADS + threadIdx.x;
                    keys[it] = 0;
                    if (i < n && sel_key(src, i, keys[it])) okmask |= 1u << it;
                }
            }
#pragma unroll
            for (int e = 0; e < SEL_ITEMS; ++e) {
                const unsigned int digit = (keys[e] >> shift) & dmask;
                const bool ok = (okmask >> e) & 1u;
                if (PASS == 0) {
                    hist_add(hist[0], digit, ok);                    // one shared histogram serves both queries
                } else {
                    const unsigned int hi = keys[e] >> (shift + bits);
                    if (ok && !res0 && hi == pre0) atomicAdd(&hist[0][digit], 1u);
                    if (ok && q1_own && hi == pre1) atomicAdd(&hist[1][digit], 1u);
                }
            }
        }
    __syncthreads();
    unsigned int* gh = ghist + (size_t)seg_id * 2 * SEL_BINS;
    if (any_live)
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

    // ---- last block of this segment: narrow the live queries ----
    const int nb = 1 << bits;
    const int per = SEL_BINS / SEL_THREADS;                          // 8 bins per thread
    if (threadIdx.x < 2) { new_prefix[threadIdx.x] = threadIdx.x ? pre1 : pre0; new_rank[threadIdx.x] = threadIdx.x ? rank1 : rank0; }
    __syncthreads();
    for (int q = 0; q < 2; ++q) {
        const bool live = any_live && !(q == 0 ? res0 : res1);
        const int hsel = (q == 1 && q1_own) ? 1 : 0;
        const unsigned int* srch = gh + hsel * SEL_BINS;
        long long loc[per], tot = 0;
#pragma unroll
        for (int k = 0; k < per; ++k) {
            int b = threadIdx.x * per + k;
            loc[k] = (live && b < nb) ? (long long)__ldcg(&srch[b]) : 0;
            tot += loc[k];
        }
        scan_tot[threadIdx.x] = tot;
        __syncthreads();
        if (threadIdx.x == 0) {                                      // 256-entry serial exclusive scan
            long long run = 0;
            for (int i = 0; i < SEL_THREADS; ++i) { long long v = scan_tot[i]; scan_tot[i] = run; run += v; }
        }
        __syncthreads();
        const long long want = (q == 0) ? rank0 : rank1;
        long long run = scan_tot[threadIdx.x];
        if (live) {
#pragma unroll
            for (int k = 0; k < per; ++k) {
                if (want >= run && want < run + loc[k]) {            // exactly one (thread, k) matches
                    unsigned int b = threadIdx.x * per + k;
                    unsigned int old = PASS == 0 ? 0u : (q == 0 ? pre0 : pre1);
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
        wp->prefix[0] = new_prefix[0]; wp->prefix[1] = new_prefix[1];
        wp->rank[0] = new_rank[0];     wp->rank[1] = new_rank[1];
        if (PASS == 2) {
            const unsigned int key0 = res0 ? wp->resolved_key[0] : new_prefix[0];
            const unsigned int key1 = res1 ? wp->resolved_key[1] : new_prefix[1];
            sel_finish(seg, wp->n_valid, wp->gamma, empty, key0, key1, out + seg_id);
        }
    }
}

// ---------------------------------------------------------------------------------
// C: resolve — one block per segment selects the live ranks among the candidates (a few
// per cent of the segment, L2 resident) or, on fallback, in the whole segment.
// ---------------------------------------------------------------------------------
#define SEL_RESOLVE_THREADS 1024             // one block per segment: all the parallelism a segment gets
__global__ void __launch_bounds__(SEL_RESOLVE_THREADS)
select_resolve_kernel(const da3s_select_seg* __restrict__ segs, SelWork* work, da3s_select_out* out) {
    __shared__ SelScratch sc;
    const int seg_id = blockIdx.x;
    const SelWork w = work[seg_id];
    if (w.done == 1) return;                                         // answered by the sample stage
    const da3s_select_seg seg = segs[seg_id];
    const bool empty = (w.done == 2);
    unsigned int key[2] = {w.resolved_key[0], w.resolved_key[1]};
    if (!empty) {
        da3s_select_seg src = seg;
        src.n = w.n; src.kind = w.kind;
        if (w.kind != SEL_KIND_KEYS) src.a = (const float*)w.data;
        const unsigned int* kdata = (const unsigned int*)w.data;
        // candidates lie strictly inside (lo_key, hi_key): select on k - lo_key, whose leading zero bits are known
        const bool rel = (w.kind == SEL_KIND_KEYS);
        const unsigned int base = rel ? w.lo_key : 0u;
        const int lz = rel ? __clz((w.hi_key - w.lo_key) | 1u) : 0;
        auto get = [&](long long i, unsigned int& k) -> bool {
            if (rel) { k = kdata[i] - base; return true; }
            return sel_key(src, i, k);
        };
        if (!w.resolved[0] || !w.resolved[1]) {                       // block-uniform
            const long long r0 = w.resolved[0] ? w.rank[1] : w.rank[0], r1 = w.resolved[1] ? w.rank[0] : w.rank[1];
            unsigned int k0, k1;
            block_select2(w.n, r0, r1, get, sc, lz, k0, k1);
            if (!w.resolved[0]) key[0] = base + k0;
            if (!w.resolved[1]) key[1] = base + k1;
        }
    }
    if (threadIdx.x == 0) sel_finish(seg, w.n_valid, w.gamma, empty, key[0], key[1], out + seg_id);
}

// internal entry used by pair_align.cu as well
int da3s_select_impl(da3s_ctx* ctx, const da3s_select_seg* segs, int n_segs, long long max_n,
                     da3s_select_out* out, cudaStream_t st) {
    if (n_segs <= 0) return DA3S_OK;
    if (n_segs > 65535) return DA3S_EINVAL;
    size_t save_top = ctx->ws_top;
    const long long per_block = (long long)SEL_THREADS * SEL_ITEMS;
    // candidate capacity per segment: 1/8 of the largest segment (the bracket holds ~8 % on average)
    long long cand_cap = (max_n / 8 + per_block - 1) / per_block * per_block;
    if (cand_cap < per_block) cand_cap = per_block;
    WS_ALLOC(ctx, SelWork, work, n_segs);
    const bool big = max_n > SEL_SINGLE_MAX;                     // candidates too many for one block per segment
    WS_ALLOC(ctx, unsigned int, tickets, n_segs);
    WS_ALLOC(ctx, unsigned int, ghist, big ? (size_t)n_segs * 2 * SEL_BINS : 1);
    WS_ALLOC(ctx, unsigned int, cand, (size_t)n_segs * (size_t)cand_cap);
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(tickets, 0, sizeof(unsigned int) * n_segs, st));
    if (big) DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(ghist, 0, sizeof(unsigned int) * (size_t)n_segs * 2 * SEL_BINS, st));
    select_sample_kernel<<<n_segs, SEL_THREADS, 0, st>>>(segs, work, out);
    DA3S_LAUNCH_CHECK(ctx);
    if (max_n > SEL_SAMPLE) {
        // chunks per block: SEL_CHUNKS for big batches (amortises the block epilogue), fewer when that would leave
        // the GPU with less than ~8 blocks per SM (a handful of overlap frames: latency, not bandwidth, is the cost)
        int chunks = SEL_CHUNKS;
        const long long n_chunks = (max_n + per_block - 1) / per_block;
        while (chunks > 1 && (n_chunks + chunks - 1) / chunks * n_segs < (long long)ctx->sm_count * 8) chunks >>= 1;
        long long bx = (n_chunks + chunks - 1) / chunks;
        if (bx > 2147483647LL) return DA3S_EINVAL;
        select_count_kernel<<<dim3((unsigned int)bx, n_segs), SEL_THREADS, 0, st>>>(segs, work, cand, cand_cap, tickets, chunks);
        DA3S_LAUNCH_CHECK(ctx);
        if (!big) {
            select_resolve_kernel<<<n_segs, SEL_RESOLVE_THREADS, 0, st>>>(segs, work, out);
            DA3S_LAUNCH_CHECK(ctx);
        } else {
            dim3 grid((unsigned int)(cand_cap / per_block), n_segs);
            select_pass_kernel<0><<<grid, SEL_THREADS, 0, st>>>(segs, work, ghist, tickets, out);
            DA3S_LAUNCH_CHECK(ctx);
            select_pass_kernel<1><<<grid, SEL_THREADS, 0, st>>>(segs, work, ghist, tickets, out);
            DA3S_LAUNCH_CHECK(ctx);
            select_pass_kernel<2><<<grid, SEL_THREADS, 0, st>>>(segs, work, ghist, tickets, out);
            DA3S_LAUNCH_CHECK(ctx);
        }
    }
    ctx->ws_top = save_top;     // scratch is free again once the kernels are queued (stream order)
    return DA3S_OK;
}

extern "C" int da3s_select(da3s_ctx* ctx, const da3s_select_seg* segs, int n_segs, long long max_n,
                           da3s_select_out* out, void* stream) {
    if (!ctx || !segs || !out || n_segs < 0 || max_n < 0) return DA3S_EINVAL;
    ws_reset(ctx);
    return da3s_select_impl(ctx, segs, n_segs, max_n, out, (cudaStream_t)stream);
}
