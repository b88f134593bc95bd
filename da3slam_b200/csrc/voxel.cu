// voxel.cu — K6: voxel-grid downsample on a GPU hash grid (oracle/SPEC.md section 5).
//
// key    = floor(f64(p) / f64(voxel)) per axis, 21 bits per axis, bias 2^20 (63 bits; all ones = EMPTY)
// table  = rec[slots], one 64-byte record (two 32-byte sectors) per slot:
//            sector 0:  key | sum_qx | sum_qy | sum_qz          sum_q = sum of llrint(frac * 2^32), exact int64
//            sector 1:  (count, sum_r) | (sum_g, sum_b) | pad | pad
//          followed by ONE OCCUPANCY BIT per slot (slots / 8 bytes: 4 MB for 2^25 slots, L2 resident).
// Integer accumulation makes the result independent of insertion order (bit-identical run to run and
// across GPUs), which floating-point atomics would not be.
//
// Protocol (vox_commit).  The occupancy bit arbitrates: `atomicOr` on the bitmap word is the ONLY read a claim
// needs, so claiming a cold slot never fetches its record from DRAM.  The winner writes both sectors of the record
// as whole 256-bit stores (sector 0 with key | BUSY) and then publishes the key with st.release.gpu; a thread
// that finds the bit already set reads the key with ld.acquire.gpu (spinning while bit 63 is set: EMPTY or BUSY,
// i.e. while the claimer has not published), and if it is its own key adds five 64-bit integers with RED — ordered
// after the claimer's stores by the release / acquire pair.  Records are never cleared: the compaction resets
// only the key word of the records it emits and the bitmap words it has walked.
//
// Contention is cut before it reaches L2: points of one warp that fall in the same voxel (neighbouring pixels
// of a frame usually do) are combined (match.any + warp reductions) and issue one commit per distinct voxel.
//
// Algorithmic bytes: 12 (+3 rgb, +1 mask) per input point read (fused export: 8 per pixel + 3 per kept point);
// 12 (+3) + 4 (+8 key) per occupied voxel written.  Table traffic is not algorithmic; per voxel it is one 64-byte
// record written by the claim and read by the compaction.
#include "common.cuh"
#include "unproject_frame.cuh"

#define VOX_EMPTY 0xFFFFFFFFFFFFFFFFull
#define VOX_BUSY 0x8000000000000000ull      // keys use 63 bits; EMPTY has the bit set too: "bit 63 set" == not (yet) a published key
#ifndef VOX_EXP_NOFENCE
#define VOX_EXP_NOFENCE 0                   // measurement only (NOT correct): publish without release semantics
#endif
#ifndef VOX_EXP_NOCOMMIT
#define VOX_EXP_NOCOMMIT 0                  // measurement only: 1 = no table access at all (what the kernel costs before the table)
#endif
#ifndef VOX_MERGE_PULL
#define VOX_MERGE_PULL 1                    // 1: leaders pull their peers' 32-bit point values (4 shuffles per round); 0: 64-bit partial sums (10)
#endif
#define VOX_BIAS (1 << 20)
#define VOX_REC 8                           // u64 words per record
#define VOX_MAX_PROBE 4096
#ifndef VOX_REDUX_GROUPS
#define VOX_REDUX_GROUPS 4                  // batches with at most this many distinct voxels merge by masked warp reductions
#endif
#ifndef VOX_DIV_FAST
#define VOX_DIV_FAST 1                      // 1: correctly rounded p / voxel by two Markstein corrections (5 DP ops); 0: __ddiv_rn
#endif
#ifndef VOX_HASH32
#define VOX_HASH32 1                        // 1: 32-bit multiplicative hash of the coarse cell; 0: 64-bit murmur finaliser
#endif
#define VOX_REC_PTR(acc, s) ((acc) + (size_t)(s) * VOX_REC)
#define VOX_BITMAP(acc, slots) (reinterpret_cast<unsigned int*>(const_cast<unsigned long long*>(acc) + (size_t)(slots) * VOX_REC))

// Locality-preserving slot: the 4 x 4 x 4 block of voxels a key belongs to is hashed to a REGION of 64 consecutive
// slots (4 KB of records, two bitmap words), the position inside the region is the voxel's position inside its block;
// on a collision the probe moves to the next region (same position).  Neighbouring voxels — what a warp's batch of
// neighbouring pixels hits — then share DRAM rows and bitmap sectors instead of being scattered over the whole table.
#define VOX_LOCAL_BITS 2                    // 4 x 4 x 4 voxels per region
#define VOX_STEP (1ull << (3 * VOX_LOCAL_BITS))

__device__ __forceinline__ unsigned long long vox_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ unsigned long long vox_slot0(unsigned long long key, long long slots) {
    const unsigned long long m = (1ull << VOX_LOCAL_BITS) - 1ull;
    const unsigned long long local = (((key >> 42) & m) << (2 * VOX_LOCAL_BITS)) | (((key >> 21) & m) << VOX_LOCAL_BITS) | (key & m);
#if VOX_HASH32
    // 19-bit coarse cell coordinates -> one 32-bit word (three multiplies, one finishing mix)
    const unsigned int cx = (unsigned int)(key >> (42 + VOX_LOCAL_BITS)), cy = (unsigned int)(key >> (21 + VOX_LOCAL_BITS)) & 0x7FFFFu,
                       cz = (unsigned int)(key >> VOX_LOCAL_BITS) & 0x7FFFFu;
    unsigned int h = (cx * 0x9E3779B1u) ^ (cy * 0x85EBCA77u) ^ (cz * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return (((unsigned long long)h << (3 * VOX_LOCAL_BITS)) | local) & (unsigned long long)(slots - 1);
#else
    const unsigned long long coarse = key & ~((m << 42) | (m << 21) | m);
    return ((vox_hash(coarse) << (3 * VOX_LOCAL_BITS)) | local) & (unsigned long long)(slots - 1);
#endif
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void voxel_clear_kernel(unsigned long long* acc, long long slots) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one slot per thread: key word + its share of the bitmap
    unsigned int* bm = VOX_BITMAP(acc, slots);
    for (; i < slots; i += (long long)gridDim.x * blockDim.x) {
        *VOX_REC_PTR(acc, i) = VOX_EMPTY;
        if ((i & 31) == 0) bm[i >> 5] = 0u;
    }
}

// Add one (already merged) contribution to the voxel `key` (protocol: file header), probing from `slot`.
__device__ __forceinline__ bool vox_commit_from(unsigned long long* __restrict__ acc, long long slots, unsigned long long key,
                                                unsigned long long sx, unsigned long long sy, unsigned long long sz,
                                                unsigned long long cr, unsigned long long gb, unsigned long long slot) {
    unsigned int* __restrict__ bm = VOX_BITMAP(acc, slots);
    for (int probe = 0; probe < VOX_MAX_PROBE; ++probe) {
        unsigned long long* rec = VOX_REC_PTR(acc, slot);
        const unsigned int bit = 1u << (slot & 31);
        const unsigned int old = atomicOr(bm + (slot >> 5), bit);
        if (!(old & bit)) {
            // ours: both sectors as whole 256-bit stores (a full-sector write allocates in L2 without fetching the line
            // from DRAM; a partial one does not) — sector 0 carries the key with VOX_BUSY set — then the key alone with
            // release semantics: everything above is visible device-wide before the key loses its BUSY bit
            asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec + 4), "l"(cr), "l"(gb), "l"(0ull), "l"(0ull) : "memory");
            asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec), "l"(key | VOX_BUSY), "l"(sx), "l"(sy), "l"(sz) : "memory");
#if VOX_EXP_NOFENCE
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(rec), "l"(key) : "memory");
#else
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(rec), "l"(key) : "memory");
#endif
            return true;
        }
        unsigned long long cur = ld_acquire_u64(rec);
        while (cur & VOX_BUSY) { __nanosleep(32); cur = ld_acquire_u64(rec); }       // EMPTY or key | BUSY: claimed, not yet published
        if (cur == key) {
            red_add_u64(rec + 1, sx); red_add_u64(rec + 2, sy); red_add_u64(rec + 3, sz);
            red_add_u64(rec + 4, cr);
            if (gb) red_add_u64(rec + 5, gb);
            return true;
        }
        slot = (slot + VOX_STEP) & (unsigned long long)(slots - 1);
    }
    return false;
}
__device__ __forceinline__ bool vox_commit(unsigned long long* __restrict__ acc, long long slots, unsigned long long key,
                                           unsigned long long sx, unsigned long long sy, unsigned long long sz,
                                           unsigned long long cr, unsigned long long gb) {
    return vox_commit_from(acc, slots, key, sx, sy, sz, cr, gb, vox_slot0(key, slots));
}

// SPEC 5 quantisation of one coordinate: q = f64(p) / f64(voxel) correctly rounded, k = floor(q), f = llrint((q - k) 2^32).
// VOX_DIV_FAST: q by Markstein's iteration instead of the division routine: q0 = p * y with y = RN(1 / voxel), then twice
// r = fma(-q, v, p) (the exact residual), q = fma(r, y, q).  After the first correction q is within one ulp of p / v, and
// for such a q the second correction returns the correctly rounded quotient (Markstein 1990, Theorem: y = RN(1 / v), v's
// significand not all ones — checked on the host, the division routine is used otherwise).  5 DP operations instead of ~40.
struct VoxQuant { double v, y; bool fast; };
__device__ __forceinline__ double vox_div(double p, const VoxQuant& qz) {
#if VOX_DIV_FAST
    if (qz.fast) {
        double q = p * qz.y;
        double r = fma(-q, qz.v, p);
        q = fma(r, qz.y, q);
        r = fma(-q, qz.v, p);
        return fma(r, qz.y, q);
    }
#endif
    return __ddiv_rn(p, qz.v);
}
static VoxQuant make_quant(float voxel) {
    VoxQuant q;
    q.v = (double)voxel;
    q.y = 1.0 / q.v;                                         // IEEE division on the host: correctly rounded reciprocal
    unsigned long long bits;
    memcpy(&bits, &q.v, sizeof(bits));
    q.fast = (bits & 0xFFFFFFFFFFFFFull) != 0xFFFFFFFFFFFFFull;
    return q;
}

// one batch of up to 32 points held in registers: float64 quantisation, in-warp merge of the lanes that fall
// in the same voxel, then probe / claim / five additions by the merged lanes.  Every lane of the warp calls it.
#define VOX_SCRATCH 48                      // uint4 per warp: one published point per lane, 24 bytes each (merge)
// quantise + merge; returns true on the lanes that carry a merged contribution (key, sums) afterwards
__device__ __forceinline__ bool vox_merge_batch(bool active, float px, float py, float pz, unsigned int rgb, bool has_rgb, const VoxQuant& qz,
                                                uint4* __restrict__ gather, unsigned long long& key, unsigned long long& sx,
                                                unsigned long long& sy, unsigned long long& sz, unsigned long long& cr, unsigned long long& gb) {
    const unsigned int lane = threadIdx.x & 31;
    key = 0; sx = 0; sy = 0; sz = 0; cr = 1ull << 32; gb = 0ull;
    if (active) {                                                   // NaN / infinite coordinates fail the range test below
        const double qx = vox_div((double)px, qz), qy = vox_div((double)py, qz), qzz = vox_div((double)pz, qz);
        const double kx = floor(qx), ky = floor(qy), kz = floor(qzz);
        active = (fabs(kx) < (double)VOX_BIAS) && (fabs(ky) < (double)VOX_BIAS) && (fabs(kz) < (double)VOX_BIAS);
        if (active) {
            key = ((unsigned long long)((long long)kx + VOX_BIAS) << 42) | ((unsigned long long)((long long)ky + VOX_BIAS) << 21) |
                  (unsigned long long)((long long)kz + VOX_BIAS);
            sx = (unsigned long long)__double2ll_rn((qx - kx) * 4294967296.0);
            sy = (unsigned long long)__double2ll_rn((qy - ky) * 4294967296.0);
            sz = (unsigned long long)__double2ll_rn((qzz - kz) * 4294967296.0);
            cr |= (unsigned long long)(rgb & 0xFFu);
            gb = ((unsigned long long)((rgb >> 8) & 0xFFu) << 32) | (unsigned long long)((rgb >> 16) & 0xFFu);
        }
    }
    const unsigned int act = __ballot_sync(0xffffffffu, active);
    if (!active) return false;
    // combine the lanes of this warp that hit the same voxel; a lane alone in its voxel skips the exchange
    const unsigned int peers = __match_any_sync(act, key);
    const unsigned int leader = __ffs(peers) - 1;
    const unsigned int groups = __popc(__ballot_sync(act, lane == leader));     // distinct voxels in this batch
    if (groups <= VOX_REDUX_GROUPS) {
        // few, large groups (dense clouds: many pixels per voxel): one masked warp reduction per 16-bit limb instead
        // of (group size - 1) rounds of shuffles.  A point's sums are < 2^32, a group has <= 32 lanes: no limb overflows.
        if (peers != (1u << lane)) {
            const unsigned int n = __popc(peers);
            const unsigned long long tx = ((unsigned long long)__reduce_add_sync(peers, (unsigned int)(sx >> 16)) << 16) +
                                          __reduce_add_sync(peers, (unsigned int)(sx & 0xFFFFull));
            const unsigned long long ty = ((unsigned long long)__reduce_add_sync(peers, (unsigned int)(sy >> 16)) << 16) +
                                          __reduce_add_sync(peers, (unsigned int)(sy & 0xFFFFull));
            const unsigned long long tz = ((unsigned long long)__reduce_add_sync(peers, (unsigned int)(sz >> 16)) << 16) +
                                          __reduce_add_sync(peers, (unsigned int)(sz & 0xFFFFull));
            unsigned long long tr = 0, tg = 0, tb = 0;
            if (has_rgb) {                                       // warp-uniform
                tr = __reduce_add_sync(peers, (unsigned int)(cr & 0xFFull));
                tg = __reduce_add_sync(peers, (unsigned int)(gb >> 32));
                tb = __reduce_add_sync(peers, (unsigned int)(gb & 0xFFull));
            }
            if (lane != leader) return false;
            sx = tx; sy = ty; sz = tz;
            cr = ((unsigned long long)n << 32) | tr;
            gb = (tg << 32) | tb;
        }
    } else if (peers != (1u << lane)) {
        unsigned int rest = peers & ~(1u << leader);                // identical for every lane of the group
#if VOX_MERGE_PULL
        // every lane still holds ONE point: its three fractions are <= 2^32 (llrint of a fraction just below 1 gives 2^32)
        // and its colour is 24 bits.  Every lane of a multi-point group publishes its point as three 64-bit words in the
        // warp's scratch, fraction in bits 0..39 and one colour channel in bits 40..: a group has at most 32 lanes, so the
        // fraction sums stay below 2^38 and the channel sums below 2^13 — plain 64-bit additions accumulate both without
        // a carry between the fields.  The leader reads one LDS.128 + one LDS.64 per peer and adds three words (the peers
        // never accumulate, and leave); the count is the group's population.
        {
            const unsigned long long wa = sx | ((unsigned long long)(rgb & 0xFFu) << 40);
            const unsigned long long wb = sy | ((unsigned long long)((rgb >> 8) & 0xFFu) << 40);
            const unsigned long long wc = sz | ((unsigned long long)((rgb >> 16) & 0xFFu) << 40);
            uint2* gather2 = reinterpret_cast<uint2*>(gather + 32);
            gather[lane] = make_uint4((unsigned int)wa, (unsigned int)(wa >> 32), (unsigned int)wb, (unsigned int)(wb >> 32));
            gather2[lane] = make_uint2((unsigned int)wc, (unsigned int)(wc >> 32));
            __syncwarp(peers);
            if (lane != leader) return false;
            unsigned long long ta = wa, tb = wb, tc = wc;
            while (rest) {
                const int src = __ffs(rest) - 1;
                rest &= rest - 1;
                const uint4 g = gather[src];
                const uint2 h = gather2[src];
                ta += ((unsigned long long)g.y << 32) | g.x;
                tb += ((unsigned long long)g.w << 32) | g.z;
                tc += ((unsigned long long)h.y << 32) | h.x;
            }
            const unsigned long long fmask = (1ull << 40) - 1ull;
            sx = ta & fmask; sy = tb & fmask; sz = tc & fmask;
            cr = ((unsigned long long)__popc(peers) << 32) | (ta >> 40);
            gb = ((tb >> 40) << 32) | (tc >> 40);
        }
#else
        while (rest) {
            const int src = __ffs(rest) - 1;
            rest &= rest - 1;
            const unsigned long long ax = __shfl_sync(peers, sx, src), ay = __shfl_sync(peers, sy, src);
            const unsigned long long az = __shfl_sync(peers, sz, src), ac = __shfl_sync(peers, cr, src);
            const unsigned long long ag = __shfl_sync(peers, gb, src);
            if (lane == leader) { sx += ax; sy += ay; sz += az; cr += ac; gb += ag; }
        }
#endif
        if (lane != leader) return false;
    }
    return true;
}

// One batch of up to 32 points: quantise + merge (above), then every carrying lane commits its contribution (vox_commit).
// Every lane of the warp calls it.  (A warp-cooperative variant — claimers stage their records in shared memory and lane
// pairs store both halves of a record in one instruction, one 64-byte L2 request instead of two — cut the write requests
// by 20 % but cost more in extra instructions and warp synchronisation than it saved: profiles/r2_voxel_experiments.md.)
__device__ __forceinline__ void voxel_insert_xyz(bool active, float px, float py, float pz, unsigned int rgb, bool has_rgb, const VoxQuant& qz,
                                                 unsigned long long* __restrict__ acc, long long slots,
                                                 unsigned long long* __restrict__ counters, uint4* __restrict__ gather /* [VOX_SCRATCH], this warp's */) {
    unsigned long long key, sx, sy, sz, cr, gb;
    if (!vox_merge_batch(active, px, py, pz, rgb, has_rgb, qz, gather, key, sx, sy, sz, cr, gb)) return;
#if VOX_EXP_NOCOMMIT == 1
    return;
#endif
    if (vox_commit(acc, slots, key, sx, sy, sz, cr, gb)) return;
    atomicAdd(&counters[1], cr >> 32);                              // table full: reported by finish
}

__device__ __forceinline__ void voxel_insert_point(bool active, long long i, const da3s_voxel_job& job, const VoxQuant& qz,
                                                   unsigned long long* __restrict__ acc,
                                                   long long slots, unsigned long long* __restrict__ counters, uint4* __restrict__ gather) {
    float px = 0.0f, py = 0.0f, pz = 0.0f;
    unsigned int rgb = 0u;
    if (active) {
        px = job.xyz[3 * i]; py = job.xyz[3 * i + 1]; pz = job.xyz[3 * i + 2];
        if (job.rgb) rgb = (unsigned int)job.rgb[3 * i] | ((unsigned int)job.rgb[3 * i + 1] << 8) | ((unsigned int)job.rgb[3 * i + 2] << 16);
    }
    voxel_insert_xyz(active, px, py, pz, rgb, job.rgb != nullptr, qz, acc, slots, counters, gather);
}

// One launch inserts any number of clouds (a job table) — e.g. every submap of a sequence.  Work unit of a warp = a
// TILE of 128 points: 128 consecutive points of an unstructured cloud, or — when the cloud is an image sequence of row
// length `width` — 8 rows x 16 pixels.  The 2-D tile matters: a voxel's footprint is a patch of pixels, and the in-warp
// combination (match.any) only sees the batch of 32 queued points; with row-major tiles it merged ~2 points per voxel,
// with patches several times more, and every merged point saves one probe and five L2 atomics.  Each warp streams the
// mask bytes of a tile (one 4-byte load per lane), queues the indices of the points that pass in shared memory, and runs
// the expensive part — coordinates, float64 quantisation, hash probe, atomics — only on full batches of 32 queued points.
// Blocks sweep the tiles of all jobs in order (block b takes block-chunks b, b + G, ...).
#define VI_THREADS 256
#ifndef VI_MIN_BLOCKS
#define VI_MIN_BLOCKS 6
#endif
#define VI_QUEUE 256                        // per-warp ring of point indices (>= 31 left over + 128 new)
#define VI_TILES_PER_WARP 2
#define VI_TILES_PER_BLOCK (VI_TILES_PER_WARP * VI_THREADS / 32)

__global__ void __launch_bounds__(VI_THREADS, VI_MIN_BLOCKS)
voxel_insert_kernel(const da3s_voxel_job* __restrict__ jobs, da3s_voxel_job single, int n_jobs, long long chunks_per_job,
                    int width, VoxQuant qz, unsigned long long* __restrict__ acc,
                    long long slots, unsigned long long* __restrict__ counters /* [0]=voxels (set by finish) [1]=dropped */) {
    __shared__ long long queue[VI_THREADS / 32][VI_QUEUE];
    __shared__ uint4 gather_sh[VI_THREADS / 32][VOX_SCRATCH];
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long* q = queue[warp];
    uint4* gather = gather_sh[warp];
    const long long total = chunks_per_job * n_jobs;
    const int tiles_per_band = width > 0 ? (width + 15) / 16 : 1;
    for (long long c = blockIdx.x; c < total; c += gridDim.x) {
        const long long j = c / chunks_per_job;
        const da3s_voxel_job job = jobs ? jobs[j] : single;
        const long long n = job.n;
        const long long t_begin = (c - j * chunks_per_job) * VI_TILES_PER_BLOCK + (long long)warp * VI_TILES_PER_WARP;
        unsigned int head = 0, count = 0;                           // warp-uniform ring state
#pragma unroll 1
        for (int tt = 0; tt < VI_TILES_PER_WARP; ++tt) {
            const long long t = t_begin + tt;
            // the 4 consecutive points of this lane: [i0, i0 + lim)
            long long i0; int lim = 4;
            if (width > 0) {
                const long long band = t / tiles_per_band;
                const int tx = (int)(t - band * tiles_per_band);
                const int col = tx * 16 + (int)(lane & 3) * 4;
                i0 = (band * 8 + (lane >> 2)) * (long long)width + col;
                lim = width - col;                                  // the last tile of a band is narrower
            } else {
                i0 = t * 128 + (long long)lane * 4;
            }
            if (lim > 4) lim = 4;
            if (n - i0 < lim) lim = (int)(n - i0 < 0 ? 0 : n - i0);
            if (__ballot_sync(0xffffffffu, lim > 0) == 0) break;    // past the end of this job (warp-uniform)
            unsigned int flags = 0;
            if (lim > 0) {
                if (!job.mask) flags = (1u << lim) - 1u;
                else if (lim == 4 && ((reinterpret_cast<uintptr_t>(job.mask + i0) & 3) == 0)) {
                    const uchar4 m = *reinterpret_cast<const uchar4*>(job.mask + i0);
                    flags = (m.x ? 1u : 0u) | (m.y ? 2u : 0u) | (m.z ? 4u : 0u) | (m.w ? 8u : 0u);
                } else {
                    for (int b = 0; b < lim; ++b) flags |= job.mask[i0 + b] ? (1u << b) : 0u;
                }
            }
            const unsigned int cnt = __popc(flags);
            unsigned int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (unsigned int)o) incl += v;
            }
            const unsigned int total_new = __shfl_sync(0xffffffffu, incl, 31);
            unsigned int pos = head + count + incl - cnt;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if ((flags >> b) & 1u) { q[pos & (VI_QUEUE - 1)] = i0 + b; ++pos; }
            count += total_new;
            __syncwarp();
            while (count >= 32) {
                const long long i = q[(head + lane) & (VI_QUEUE - 1)];
                voxel_insert_point(true, i, job, qz, acc, slots, counters, gather);
                __syncwarp();
                head += 32; count -= 32;
            }
        }
        if (count) {                                                // the queue never crosses a job boundary
            const bool active = lane < count;
            const long long i = active ? q[(head + lane) & (VI_QUEUE - 1)] : 0;
            voxel_insert_point(active, i, job, qz, acc, slots, counters, gather);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------
// Fused export: depth -> world point (+ Sim(3)) -> confidence / validity filter -> voxel grid, without
// ever writing the points.  Same arithmetic as K1's fast float32 path followed by the insert above
// (shared device functions), so the grid is bit-identical to the two-kernel route; what disappears
// is 13 B/pixel of K1 stores and ~16 B/kept point of insert loads.
// ---------------------------------------------------------------------------------
#define EX_CONST 24                         // floats per frame: cu cv 1/fu 1/fv | Mf[9] | mf[3] | thr_ge thr_gt d_lo d_hi | pad

// Per-frame constants of the fused export.  The filter flags become four thresholds, so the per-pixel test is four
// compares and no flag logic:  keep = conf >= thr_ge & conf > thr_gt & depth > d_lo & depth <= d_hi  (a disabled test
// gets -inf / +inf; a frame without confidences is given conf = 1).
__global__ void export_frame_const_kernel(const da3s_export_job* __restrict__ jobs, int n_frames, int world, int flags, float conf_thr,
                                          float conf_floor, float depth_eps, float* __restrict__ fcs) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const da3s_export_job j = jobs[f];
    UnprojFrame fr;
    compose_unproj_frame(*j.cam, j.sim3, world != 0, fr);
    float* o = fcs + (size_t)f * EX_CONST;
    o[0] = fr.cuf; o[1] = fr.cvf; o[2] = fr.ifu; o[3] = fr.ifv;
    for (int k = 0; k < 9; ++k) o[4 + k] = fr.Mf[k];
    for (int k = 0; k < 3; ++k) o[13 + k] = fr.mf[k];
    const float thr = j.conf_thr ? *j.conf_thr : conf_thr;
    // a NaN threshold = a selection over no usable confidence (da3s_select): the reference keeps every point then
    // (viewer.py:333-338), so the confidence tests are switched off for this frame
    const bool use_conf = j.conf != nullptr && !(thr != thr);
    const float ninf = __int_as_float(0xff800000), pinf = __int_as_float(0x7f800000);
    float thr_ge = ninf, thr_gt = ninf;
    if (use_conf) {
        if (flags & DA3S_MASK_CONF_GE) thr_ge = thr;
        if (flags & DA3S_MASK_CONF_GT) thr_gt = thr;
        if (flags & DA3S_MASK_CONF_FLOOR) thr_gt = fmaxf(thr_gt, conf_floor);
    }
    o[16] = thr_ge; o[17] = thr_gt;
    o[18] = (flags & DA3S_MASK_DEPTH) ? depth_eps : ninf;          // depth > eps and finite  ==  eps < depth <= FLT_MAX
    o[19] = (flags & DA3S_MASK_DEPTH) ? 3.402823466e+38f : pinf;   // without the test non-finite depths go on and are dropped as points
    o[20] = use_conf ? 1.0f : 0.0f;
    o[21] = o[22] = o[23] = 0.0f;
}

struct ExportArgs {
    const da3s_export_job* jobs; const float* fcs;
    int n_frames, H, W;
    VoxQuant qz;
    int chunks_per_frame, tiles_per_band;
    unsigned long long* acc; long long slots; unsigned long long* counters;
};

__global__ void __launch_bounds__(VI_THREADS, VI_MIN_BLOCKS)
export_voxel_kernel(ExportArgs a) {
    __shared__ unsigned int qd[VI_THREADS / 32][VI_QUEUE];        // depth bits of the queued pixels
    __shared__ unsigned int qp[VI_THREADS / 32][VI_QUEUE];        // (v << 16) | u
    __shared__ float fc_sh[VI_THREADS / 32][EX_CONST];
    __shared__ uint4 gather_sh[VI_THREADS / 32][VOX_SCRATCH];
    const VoxQuant qz = a.qz;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int* qdw = qd[warp];
    unsigned int* qpw = qp[warp];
    float* fc = fc_sh[warp];
    uint4* gather = gather_sh[warp];
    const int W = a.W, H = a.H;
    const int total = a.chunks_per_frame * a.n_frames;
    const int lrow = (int)(lane >> 2), lcol = (int)(lane & 3) * 4;
    for (int c = blockIdx.x; c < total; c += gridDim.x) {
        const int f = c / a.chunks_per_frame;
        const da3s_export_job job = a.jobs[f];
        __syncwarp();
        if (lane < EX_CONST) fc[lane] = a.fcs[(size_t)f * EX_CONST + lane];
        __syncwarp();
        const float thr_ge = fc[16], thr_gt = fc[17], d_lo = fc[18], d_hi = fc[19];
        const bool use_conf = fc[20] != 0.0f;
        const int t_begin = (c - f * a.chunks_per_frame) * VI_TILES_PER_BLOCK + (int)warp * VI_TILES_PER_WARP;
        unsigned int head = 0, count = 0;                           // warp-uniform ring state
        auto batch = [&](bool active, unsigned int dbits, unsigned int uv) {
            const int u = (int)(uv & 0xFFFFu), v = (int)(uv >> 16);
            const float d = __uint_as_float(dbits);
            unsigned int rgb = 0u;
            if (active && job.rgb) {
                const unsigned int off = 3u * (unsigned int)(v * W + u);
                const uint8_t* p = job.rgb + off;
                rgb = (unsigned int)p[0] | ((unsigned int)p[1] << 8) | ((unsigned int)p[2] << 16);
            }
            // K1 fast path (unproject_pixel<DA3S_UNPROJ_FAST> with the composed float32 transform)
            float x, y;
            cam_fast((float)u, (float)v, d, fc[0], fc[1], fc[2], fc[3], x, y);
            const float X = fmaf(fc[4], x, fmaf(fc[5], y, fmaf(fc[6], d, fc[13])));
            const float Y = fmaf(fc[7], x, fmaf(fc[8], y, fmaf(fc[9], d, fc[14])));
            const float Z = fmaf(fc[10], x, fmaf(fc[11], y, fmaf(fc[12], d, fc[15])));
            voxel_insert_xyz(active, X, Y, Z, rgb, job.rgb != nullptr, qz, a.acc, a.slots, a.counters, gather);
        };
#pragma unroll 1
        for (int tt = 0; tt < VI_TILES_PER_WARP; ++tt) {
            const int t = t_begin + tt;
            const int band = t / a.tiles_per_band;
            const int col = (t - band * a.tiles_per_band) * 16 + lcol;
            const int row = band * 8 + lrow;
            if (band * 8 >= H) break;                               // past the end of this frame (warp-uniform)
            int lim = W - col;
            if (lim > 4) lim = 4;
            if (row >= H) lim = 0;
            float d4[4] = {0.f, 0.f, 0.f, 0.f}, c4[4] = {1.f, 1.f, 1.f, 1.f};
            if (lim > 0) {
                const unsigned int i0 = (unsigned int)(row * W + col);
                if (lim == 4 && ((i0 & 1u) == 0u)) {                // 8-byte aligned pairs (frames are 16-byte aligned)
                    const float2 da = ldg_stream(reinterpret_cast<const float2*>(job.depth + i0));
                    const float2 db = ldg_stream(reinterpret_cast<const float2*>(job.depth + i0 + 2));
                    d4[0] = da.x; d4[1] = da.y; d4[2] = db.x; d4[3] = db.y;
                    if (use_conf) {
                        const float2 ca = ldg_stream(reinterpret_cast<const float2*>(job.conf + i0));
                        const float2 cb = ldg_stream(reinterpret_cast<const float2*>(job.conf + i0 + 2));
                        c4[0] = ca.x; c4[1] = ca.y; c4[2] = cb.x; c4[3] = cb.y;
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (b < lim) { d4[b] = job.depth[i0 + b]; if (use_conf) c4[b] = job.conf[i0 + b]; }
                }
            }
            unsigned int flags = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const bool k = (b < lim) & (c4[b] >= thr_ge) & (c4[b] > thr_gt) & (d4[b] > d_lo) & (d4[b] <= d_hi);
                flags |= k ? (1u << b) : 0u;
            }
            const unsigned int cnt = __popc(flags);
            unsigned int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int w = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (unsigned int)o) incl += w;
            }
            const unsigned int total_new = __shfl_sync(0xffffffffu, incl, 31);
            unsigned int pos = head + count + incl - cnt;
            const unsigned int uv0 = ((unsigned int)row << 16) | (unsigned int)col;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if ((flags >> b) & 1u) {
                    qdw[pos & (VI_QUEUE - 1)] = __float_as_uint(d4[b]);
                    qpw[pos & (VI_QUEUE - 1)] = uv0 + b;
                    ++pos;
                }
            count += total_new;
            __syncwarp();
            while (count >= 32) {
                const unsigned int qi = (head + lane) & (VI_QUEUE - 1);
                batch(true, qdw[qi], qpw[qi]);
                __syncwarp();
                head += 32; count -= 32;
            }
        }
        if (count) {                                                // the queue never crosses a frame boundary
            const bool active = lane < count;
            const unsigned int qi = (head + lane) & (VI_QUEUE - 1);
            batch(active, active ? qdw[qi] : 0u, active ? qpw[qi] : 0u);
            __syncwarp();
        }
    }
}

extern "C" int da3s_unproject_voxel_jobs(da3s_ctx* ctx, const da3s_export_job* jobs_dev, int n_frames, int H, int W, int flags,
                                         float conf_thr, float conf_floor, float depth_eps, float voxel, void* stream) {
    if (!ctx || !jobs_dev || n_frames < 0 || H <= 0 || W <= 0 || H > 65535 || W > 65535 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if ((flags & DA3S_UNPROJ_MODEMASK) != DA3S_UNPROJ_FAST || (flags & (DA3S_UNPROJ_OUT_F64 | DA3S_MASK_WORLD_Z))) return DA3S_EINVAL;
    if ((flags & DA3S_MASK_CONF_GT) && (flags & DA3S_MASK_CONF_GE)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    if (n_frames == 0) return DA3S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    size_t save_top = ctx->ws_top;
    WS_ALLOC(ctx, float, fcs, (size_t)n_frames * EX_CONST);
    export_frame_const_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(jobs_dev, n_frames, (flags & DA3S_UNPROJ_WORLD) ? 1 : 0, flags, conf_thr,
                                                                      conf_floor, depth_eps, fcs);
    DA3S_LAUNCH_CHECK(ctx);
    ExportArgs a;
    a.jobs = jobs_dev; a.fcs = fcs; a.n_frames = n_frames; a.H = H; a.W = W;
    a.qz = make_quant(voxel);
    a.tiles_per_band = (W + 15) / 16;
    const long long tiles = (long long)((H + 7) / 8) * a.tiles_per_band;
    const long long cpf = (tiles + VI_TILES_PER_BLOCK - 1) / VI_TILES_PER_BLOCK;
    if (cpf * n_frames > 2147483647LL || (long long)H * W > 1400000000LL) return DA3S_EINVAL;     // 32-bit tile / pixel indices in the kernel
    a.chunks_per_frame = (int)cpf;
    a.acc = ctx->vox_acc; a.slots = ctx->vox_slots; a.counters = ctx->vox_counters;
    int per_sm = 0;
    DA3S_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, export_voxel_kernel, VI_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const long long total = (long long)a.chunks_per_frame * n_frames, cap = (long long)ctx->sm_count * per_sm;
    prof_begin(ctx, DA3S_TIMED_EXPORT_VOXEL, st);
    export_voxel_kernel<<<(unsigned int)(total > cap ? cap : total), VI_THREADS, 0, st>>>(a);
    prof_end(ctx, DA3S_TIMED_EXPORT_VOXEL, st);
    DA3S_LAUNCH_CHECK(ctx);
    ctx->ws_top = save_top;     // the constants are consumed in stream order
    return DA3S_OK;
}

// Compaction without atomics or barriers, driven by the occupancy bitmap (never by the records):
//   count  every WARP counts the set bits of its fixed range of 512 slots (16 bitmap words)
//   scan   two levels: per-chunk scans of the group counts, then the chunk totals (offsets and the total)
//   emit   every warp walks the occupied slots of its range densely, reads key + record, emits, resets the
//          KEY and finally its 16 bitmap words: records are rewritten by the next claimer, so the table is
//          ready for the next begin() without a clearing pass.
// Output order = slot order (which of two colliding keys gets the earlier slot depends on the insertion race;
// the canonical order is ascending key — ops.VoxelGrid.read(sort=True)).
#define VC_THREADS 256
#define VC_ROUNDS 16
#define VC_PER_WARP (32 * VC_ROUNDS)        // 512 slots per warp
#define VC_PER_BLOCK (VC_PER_WARP * VC_THREADS / 32)

__global__ void __launch_bounds__(VC_THREADS)
voxel_count_kernel(const unsigned int* __restrict__ bitmap, long long n_warps, unsigned int* __restrict__ warp_counts) {
    // one THREAD per 512-slot group: four 128-bit loads of its 16 bitmap words
    const long long wid = (long long)blockIdx.x * VC_THREADS + threadIdx.x;
    if (wid >= n_warps) return;
    const uint4* p = reinterpret_cast<const uint4*>(bitmap + wid * VC_ROUNDS);
    unsigned int c = 0;
#pragma unroll
    for (int j = 0; j < VC_ROUNDS / 4; ++j) {
        const uint4 w = p[j];
        c += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
    warp_counts[wid] = c;
}

#define VS_ITEMS 8                          // counts per thread: one block scans a chunk of 8192 group counts
#define VS_CHUNK (1024 * VS_ITEMS)
// Two-level exclusive scan of the per-group counts: every block scans its chunk (offsets local to the chunk) and writes the
// chunk's total; voxel_scan_top_kernel then turns the <= 512 chunk totals into chunk offsets and the grand total.  (One
// block walking all 524 288 groups of a 2^28-slot table took 0.40 ms of the 4.1 ms compaction.)
__global__ void __launch_bounds__(1024)
voxel_scan_kernel(const unsigned int* __restrict__ counts, int n, unsigned long long* __restrict__ offsets,
                  unsigned long long* __restrict__ chunk_tot) {
    __shared__ unsigned long long wsum[32];
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * VS_CHUNK + threadIdx.x * VS_ITEMS;
    unsigned int c[VS_ITEMS];
    unsigned long long v = 0ull;
#pragma unroll
    for (int k = 0; k < VS_ITEMS; ++k) { c[k] = (i0 + k < n) ? counts[i0 + k] : 0u; v += c[k]; }
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long before = 0ull;
    for (unsigned int w = 0; w < warp; ++w) before += wsum[w];
    unsigned long long run = before + incl - v;
#pragma unroll
    for (int k = 0; k < VS_ITEMS; ++k) { if (i0 + k < n) offsets[i0 + k] = run; run += c[k]; }
    if (threadIdx.x == 1023) chunk_tot[blockIdx.x] = before + incl;
}

__global__ void __launch_bounds__(1024)
voxel_scan_top_kernel(unsigned long long* __restrict__ chunk_tot /* in: totals, out: exclusive offsets */, int n_chunks,
                      unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long wsum[32];
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long v = (int)threadIdx.x < n_chunks ? chunk_tot[threadIdx.x] : 0ull;     // n_chunks <= 1024
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long before = 0ull;
    for (unsigned int w = 0; w < warp; ++w) before += wsum[w];
    if ((int)threadIdx.x < n_chunks) chunk_tot[threadIdx.x] = before + incl - v;
    if (threadIdx.x == 1023) counters[0] = before + incl;
}

// rank-th occupied slot of a warp's 512-slot range from its 16 occupancy words (shared memory: words + exclusive counts)
__device__ __forceinline__ int vox_nth_slot(const unsigned int* s_word, const unsigned int* s_excl, unsigned int rank) {
    int w = 0;                                                   // word holding the rank-th occupied slot: last excl <= rank
#pragma unroll
    for (int step = VC_ROUNDS / 2; step >= 1; step >>= 1)
        if (s_excl[w + step] <= rank) w += step;
    const unsigned int word = s_word[w];
    unsigned int n = rank - s_excl[w], pos = 0;
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) {                       // position of the n-th set bit of `word`
        const unsigned int c = __popc((word >> pos) & ((1u << sh) - 1u));
        if (n >= c) { n -= c; pos += sh; }
    }
    return w * 32 + (int)pos;
}

#ifndef VE_Q
#define VE_Q 4                              // dense rounds (32 records each) a warp has in flight
#endif
#ifndef VE_MIN_BLOCKS
#define VE_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(VC_THREADS, VE_MIN_BLOCKS)
voxel_emit_kernel(unsigned long long* __restrict__ acc, long long slots,
                  const unsigned long long* __restrict__ warp_offsets, const unsigned long long* __restrict__ chunk_offsets,
                  const unsigned int* __restrict__ warp_counts, float voxel, long long max_voxels,
                  float* __restrict__ xyz_out, uint8_t* __restrict__ rgb_out, int32_t* __restrict__ count_out,
                  long long* __restrict__ key_out) {
    __shared__ unsigned int s_word[VC_THREADS / 32][VC_ROUNDS], s_excl[VC_THREADS / 32][VC_ROUNDS];
    const double vd = (double)voxel;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * (VC_THREADS / 32) + warp;
    const long long base = wid * VC_PER_WARP;
    if (base >= slots) return;
    unsigned int* bm = VOX_BITMAP(acc, slots) + wid * VC_ROUNDS;
    // The warp's 512 slots as 16 occupancy words.  The occupied slots are handled DENSELY: lane l of dense round r
    // takes the (32 r + l)-th occupied slot, so the number of rounds — and of record loads in flight per warp —
    // follows the voxels, not the table size.  The three per-group loads are independent: issued together.
    const unsigned int my_word = lane < VC_ROUNDS ? __ldcs(bm + lane) : 0u;
    const unsigned long long out0 = __ldcs(warp_offsets + wid) + chunk_offsets[wid / VS_CHUNK];
    if (__ldcs(warp_counts + wid) == 0u) return;                  // nothing in these 512 slots (warp-uniform)
    const unsigned int my_cnt = __popc(my_word);
    unsigned int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < VC_ROUNDS; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, VC_ROUNDS - 1);
    if (lane < VC_ROUNDS) { s_word[warp][lane] = my_word; s_excl[warp][lane] = incl - my_cnt; bm[lane] = 0u; }
    __syncwarp();
    for (unsigned int r0 = 0; r0 < total; r0 += 32u * VE_Q) {     // VE_Q dense rounds at a time: their record loads overlap
        ulonglong2 r01[VE_Q], r23[VE_Q], r45[VE_Q];                        // (key, sum_x), (sum_y, sum_z), ((n, sum_r), (sum_g, sum_b))
        unsigned long long* recs[VE_Q];
        bool occ[VE_Q];
#pragma unroll
        for (int q = 0; q < VE_Q; ++q) {
            const unsigned int rank = r0 + 32u * q + lane;
            occ[q] = rank < total;
            if (occ[q]) {
                unsigned long long* rec = VOX_REC_PTR(acc, base + vox_nth_slot(s_word[warp], s_excl[warp], rank));
                r01[q] = *reinterpret_cast<const ulonglong2*>(rec);
                r23[q] = *reinterpret_cast<const ulonglong2*>(rec + 2);
                r45[q] = *reinterpret_cast<const ulonglong2*>(rec + 4);
                recs[q] = rec;
            }
        }
#pragma unroll
        for (int q = 0; q < VE_Q; ++q) {
            if (!occ[q]) continue;
            const unsigned long long o = out0 + r0 + 32u * q + lane;
            if ((long long)o >= max_voxels) continue;
            const unsigned long long k64 = r01[q].x, cr = r45[q].x, gb = r45[q].y, cnt = cr >> 32;
            const double inv = 1.0 / 4294967296.0;
            const unsigned long long sq[3] = {r01[q].y, r23[q].x, r23[q].y};
            const double k[3] = {(double)((long long)((k64 >> 42) & 0x1FFFFF) - VOX_BIAS),
                                 (double)((long long)((k64 >> 21) & 0x1FFFFF) - VOX_BIAS),
                                 (double)((long long)(k64 & 0x1FFFFF) - VOX_BIAS)};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // position = f32((k + (sum_q / count) / 2^32) * voxel), each step rounded once
                double mf = __dmul_rn(__ddiv_rn((double)sq[c], (double)cnt), inv);
                xyz_out[3 * o + c] = (float)__dmul_rn(__dadd_rn(k[c], mf), vd);
            }
            if (rgb_out) {
                // (2*sum + n) / (2*n) = floor(sum/n) + (2*rem >= n), in 32-bit arithmetic
                const unsigned int n32 = (unsigned int)cnt;
                const unsigned int ch[3] = {(unsigned int)(cr & 0xFFFFFFFFull), (unsigned int)(gb >> 32), (unsigned int)(gb & 0xFFFFFFFFull)};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned int d = ch[c] / n32, rem = ch[c] - d * n32;
                    rgb_out[3 * o + c] = (uint8_t)(d + ((rem >= n32 - rem) ? 1u : 0u));
                }
            }
            count_out[o] = (int32_t)cnt;
            if (key_out) key_out[o] = (long long)k64;
        }
        // key resets LAST, when the loads above have long returned (a store to a sector whose fill is still in flight
        // stalls behind it); the record itself is rewritten by the next claimer (vox_commit)
#pragma unroll
        for (int q = 0; q < VE_Q; ++q)
            if (occ[q]) *recs[q] = VOX_EMPTY;
    }
}

// ---------------------------------------------------------------------------------
// Multi-GPU merge of voxel grids (map export): every rank fills its own grid, then ONE
// kernel compacts it and writes every record straight into the inbox of the rank that owns its
// key — peer memory over NVLink, no staging buffer and no collective in the data path — and the
// owner folds its inbox into its (clean) table.  The sums are integers, so the merged grid is
// bit-identical to inserting all points on one GPU.
//   count  per warp and destination: occupied slots whose owner is d
//   scan   one block per destination: exclusive offsets in slot order (deterministic inbox content); its last
//          thread publishes this rank's record count for destination d in the owner's `counts`
//   send   read records, reset keys and bitmap, store 48-byte records at inbox[owner][rank*cap + offset], then
//          signal arrival: every block, after a device-wide fence over its stores, bumps a local ticket; the LAST
//          block writes the step number into flags[rank] of every destination (system-scope release).  The owner's
//          merge kernel spins on its `world` flags (system-scope acquire) before it reads the inbox: no host
//          synchronisation and no process-group barrier in the data path.
// Inboxes are double-buffered by step parity: a sender's stores of step k + 1 can never land in the buffer the
// owner is still merging for step k (a rank cannot start step k + 2 before every rank has merged step k, because
// its own merge of step k + 1 waits for every peer's send of step k + 1, which follows that peer's merge of step k).
// ---------------------------------------------------------------------------------
#define VOX_MAX_WORLD 16
struct VoxPeers {
    unsigned long long* inbox[VOX_MAX_WORLD];       // this step's inbox half of every rank
    unsigned long long* counts[VOX_MAX_WORLD];      // [world] record counts, this step's half
    unsigned long long* flags[VOX_MAX_WORLD];       // [world] arrival flags (step numbers), this step's half
};

__device__ __forceinline__ int vox_owner(unsigned long long key, int world) {
    return (int)((vox_hash(key) >> 40) % (unsigned long long)world);      // high bits: independent of the slot index
}

// this warp's 512-slot group as dense occupancy ranks (shared: words + exclusive popcounts); returns the occupied count
__device__ __forceinline__ unsigned int vox_group_ranks(const unsigned int* bm, unsigned int lane, unsigned int* s_word, unsigned int* s_excl) {
    const unsigned int my_word = lane < VC_ROUNDS ? bm[lane] : 0u;
    const unsigned int my_cnt = __popc(my_word);
    unsigned int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < VC_ROUNDS; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    if (lane < VC_ROUNDS) { s_word[lane] = my_word; s_excl[lane] = incl - my_cnt; }
    __syncwarp();
    return __shfl_sync(0xffffffffu, incl, VC_ROUNDS - 1);
}

__global__ void __launch_bounds__(VC_THREADS)
voxel_count_dest_kernel(const unsigned long long* __restrict__ acc, long long slots, int world,
                        unsigned int* __restrict__ warp_counts /* [n_warps][world] */) {
    __shared__ unsigned int s_word[VC_THREADS / 32][VC_ROUNDS], s_excl[VC_THREADS / 32][VC_ROUNDS];
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * (VC_THREADS / 32) + warp;
    const long long base = wid * VC_PER_WARP;
    if (base >= slots) return;
    const unsigned int total = vox_group_ranks(VOX_BITMAP(acc, slots) + wid * VC_ROUNDS, lane, s_word[warp], s_excl[warp]);
    unsigned int mine = 0;                                       // lane d counts destination d
    for (unsigned int r0 = 0; r0 < total; r0 += 128u) {          // dense walk, four rounds of key loads in flight
        unsigned long long k[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned int rank = r0 + 32u * q + lane;
            k[q] = rank < total ? __ldcs(VOX_REC_PTR(acc, base + vox_nth_slot(s_word[warp], s_excl[warp], rank))) : VOX_EMPTY;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int owner = k[q] != VOX_EMPTY ? vox_owner(k[q], world) : -1;
            for (int d = 0; d < world; ++d) {
                const unsigned int m = __ballot_sync(0xffffffffu, owner == d);
                if ((int)lane == d) mine += __popc(m);
            }
        }
    }
    if ((int)lane < world) warp_counts[wid * world + lane] = mine;
}

__global__ void __launch_bounds__(1024)
voxel_scan_dest_kernel(const unsigned int* __restrict__ counts, int n, int world, int rank, long long cap,
                       unsigned long long* __restrict__ offsets /* [n][world] */, VoxPeers peers,
                       unsigned long long* __restrict__ counters /* [1] += records that did not fit */) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    const int d = blockIdx.x;                                    // one block per destination column
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n ? (unsigned long long)counts[(size_t)i * world + d] : 0ull;
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
        if (i < n) offsets[(size_t)i * world + d] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long total = carry;
        if ((long long)total > cap) { atomicAdd(&counters[1], total - (unsigned long long)cap); total = (unsigned long long)cap; }
        peers.counts[d][rank] = total;                           // peer store; made visible by the send kernel's flag (stream order + fence)
    }
}

__global__ void __launch_bounds__(VC_THREADS)
voxel_send_kernel(unsigned long long* __restrict__ acc, long long slots, const unsigned long long* __restrict__ offsets,
                  int world, int rank, long long cap, VoxPeers peers, unsigned int* __restrict__ ticket, unsigned long long step) {
    __shared__ unsigned int s_word[VC_THREADS / 32][VC_ROUNDS], s_excl[VC_THREADS / 32][VC_ROUNDS];
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * (VC_THREADS / 32) + warp;
    const long long base = wid * VC_PER_WARP;
    if (base < slots) {
        unsigned int* bm = VOX_BITMAP(acc, slots) + wid * VC_ROUNDS;
        const unsigned int total = vox_group_ranks(bm, lane, s_word[warp], s_excl[warp]);
        if (total) {
            unsigned long long out = (int)lane < world ? offsets[wid * world + lane] : 0ull;    // lane d: next index for destination d
            for (unsigned int r0 = 0; r0 < total; r0 += 128u) {  // dense walk in slot order (the order the count pass used)
                ulonglong2 r01[4], r23[4], r45[4];
                unsigned long long* recs[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned int rk = r0 + 32u * q + lane;
                    recs[q] = nullptr;
                    r01[q] = make_ulonglong2(VOX_EMPTY, 0ull);
                    if (rk < total) {
                        unsigned long long* rec = VOX_REC_PTR(acc, base + vox_nth_slot(s_word[warp], s_excl[warp], rk));
                        r01[q] = *reinterpret_cast<const ulonglong2*>(rec);
                        r23[q] = *reinterpret_cast<const ulonglong2*>(rec + 2);
                        r45[q] = *reinterpret_cast<const ulonglong2*>(rec + 4);
                        recs[q] = rec;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool occ = recs[q] != nullptr;
                    const int owner = occ ? vox_owner(r01[q].x, world) : -1;
                    unsigned long long my_idx = 0;
                    for (int d = 0; d < world; ++d) {
                        const unsigned int m = __ballot_sync(0xffffffffu, owner == d);
                        const unsigned long long start = __shfl_sync(0xffffffffu, out, d);
                        if (owner == d) my_idx = start + __popc(m & ((1u << lane) - 1u));
                        if ((int)lane == d) out += __popc(m);
                    }
                    if (occ && (long long)my_idx < cap) {
                        unsigned long long* dst = peers.inbox[owner] + ((size_t)rank * (size_t)cap + my_idx) * 6;   // NVLink stores
                        *reinterpret_cast<ulonglong2*>(dst) = r01[q];
                        *reinterpret_cast<ulonglong2*>(dst + 2) = r23[q];
                        *reinterpret_cast<ulonglong2*>(dst + 4) = r45[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)                        // key resets last (see voxel_emit_kernel)
                    if (recs[q]) *recs[q] = VOX_EMPTY;
            }
            if (lane < VC_ROUNDS) bm[lane] = 0u;
        }
    }
    // arrival: all of this block's peer stores are performed (system scope) before its ticket; the last block signals
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && (int)threadIdx.x < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flags[threadIdx.x] + rank), "l"(step) : "memory");
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// owner side: wait until every rank's records of this step have arrived (device-side, no host involvement)
__global__ void voxel_wait_kernel(const unsigned long long* __restrict__ flags, int world, unsigned long long step) {
    if ((int)threadIdx.x < world) {
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v < step) __nanosleep(200);
        } while (v < step);
    }
}

__global__ void __launch_bounds__(256)
voxel_merge_kernel(const unsigned long long* __restrict__ inbox, const unsigned long long* __restrict__ counts, int world, long long cap,
                   unsigned long long* __restrict__ acc, long long slots, unsigned long long* __restrict__ counters) {
    const int src = blockIdx.y;
    const long long n = (long long)counts[src];
    const unsigned long long* seg = inbox + (size_t)src * (size_t)cap * 6;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 r0 = *reinterpret_cast<const ulonglong2*>(seg + i * 6);
        const ulonglong2 r1 = *reinterpret_cast<const ulonglong2*>(seg + i * 6 + 2);
        const ulonglong2 r2 = *reinterpret_cast<const ulonglong2*>(seg + i * 6 + 4);
        const unsigned long long key = r0.x;
        const bool placed = vox_commit(acc, slots, key, r0.y, r1.x, r1.y, r2.x, r2.y);
        if (!placed) atomicAdd(&counters[1], r2.x >> 32);
    }
}

extern "C" int da3s_voxel_send(da3s_ctx* ctx, int world, int rank, void* const* inbox_ptrs, void* const* count_ptrs,
                               void* const* flag_ptrs, unsigned long long step, long long cap, void* stream) {
    if (!ctx || world < 1 || world > VOX_MAX_WORLD || rank < 0 || rank >= world || !inbox_ptrs || !count_ptrs || !flag_ptrs || cap < 1)
        return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    VoxPeers peers;
    for (int d = 0; d < VOX_MAX_WORLD; ++d) {
        peers.inbox[d] = d < world ? (unsigned long long*)inbox_ptrs[d] : nullptr;
        peers.counts[d] = d < world ? (unsigned long long*)count_ptrs[d] : nullptr;
        peers.flags[d] = d < world ? (unsigned long long*)flag_ptrs[d] : nullptr;
        if (d < world && (!peers.inbox[d] || !peers.counts[d] || !peers.flags[d])) return DA3S_EINVAL;
    }
    const int n_warps = (int)((ctx->vox_slots + VC_PER_WARP - 1) / VC_PER_WARP);
    const int n_cblocks = (n_warps + VC_THREADS / 32 - 1) / (VC_THREADS / 32);
    size_t save_top = ctx->ws_top;
    WS_ALLOC(ctx, unsigned int, warp_counts, (size_t)n_warps * world);
    WS_ALLOC(ctx, unsigned long long, warp_offsets, (size_t)n_warps * world);
    voxel_count_dest_kernel<<<n_cblocks, VC_THREADS, 0, st>>>(ctx->vox_acc, ctx->vox_slots, world, warp_counts);
    DA3S_LAUNCH_CHECK(ctx);
    voxel_scan_dest_kernel<<<world, 1024, 0, st>>>(warp_counts, n_warps, world, rank, cap, warp_offsets, peers, ctx->vox_counters);
    DA3S_LAUNCH_CHECK(ctx);
    voxel_send_kernel<<<n_cblocks, VC_THREADS, 0, st>>>(ctx->vox_acc, ctx->vox_slots, warp_offsets, world, rank, cap, peers,
                                                       (unsigned int*)(ctx->vox_counters + 2), step);
    DA3S_LAUNCH_CHECK(ctx);
    ctx->ws_top = save_top;
    ctx->vox_clean = false;         // clean again, but the table stays active: the merge inserts into it
    return DA3S_OK;
}

extern "C" int da3s_voxel_merge_inbox(da3s_ctx* ctx, const void* inbox, const void* counts, const void* flags,
                                      unsigned long long step, int world, long long cap, void* stream) {
    if (!ctx || !inbox || !counts || world < 1 || world > VOX_MAX_WORLD || cap < 1) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (flags) {                                                  // null: the caller has synchronised by other means
        voxel_wait_kernel<<<1, 32, 0, st>>>((const unsigned long long*)flags, world, step);
        DA3S_LAUNCH_CHECK(ctx);
    }
    long long want = (cap + 255) / 256, lim = (long long)ctx->sm_count * 8;
    dim3 grid((unsigned int)(want > lim ? lim : want), (unsigned int)world);
    voxel_merge_kernel<<<grid, 256, 0, st>>>((const unsigned long long*)inbox, (const unsigned long long*)counts, world, cap,
                                             ctx->vox_acc, ctx->vox_slots, ctx->vox_counters);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_voxel_begin(da3s_ctx* ctx, long long table_slots, void* stream) {
    if (!ctx || table_slots < 1024 || (table_slots & (table_slots - 1)) || table_slots > (1ll << 31)) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    // layout of the reserved tail: records [slots][8] u64 | occupancy bitmap [slots/32] u32 | counters [32] u64 (256 B:
    // voxels, dropped, send ticket) | group counts [slots/512] u32 | group offsets [slots/512] u64
    const long long n_groups = (table_slots + VC_PER_WARP - 1) / VC_PER_WARP;
    size_t bytes = (size_t)table_slots * VOX_REC * 8 + (size_t)table_slots / 8 + 256 + (size_t)n_groups * 12 + 8192 + 512;   // + <= 1024 chunk offsets
    const bool reuse = (ctx->vox_slots == table_slots) && ctx->vox_clean && ctx->vox_bytes > 0;
    if (!reuse) {
        if (bytes > ctx->ws_bytes) return DA3S_ENOMEM;
        ctx->vox_bytes = 0;
        ws_reset(ctx);
        size_t start = (ctx->ws_bytes - bytes) & ~(size_t)255;
        ctx->vox_bytes = ctx->ws_bytes - start;               // stays reserved until a different size is requested
        ctx->vox_acc = (unsigned long long*)(ctx->ws + start);
        ctx->vox_counters = (unsigned long long*)(ctx->ws + start + (size_t)table_slots * VOX_REC * 8 + (size_t)table_slots / 8);
        ctx->vox_groups = (unsigned int*)(ctx->vox_counters + 32);             // group counts, then group offsets
        ctx->vox_slots = table_slots;
        long long want = (table_slots + 255) / 256, cap = (long long)ctx->sm_count * 32;
        voxel_clear_kernel<<<(int)(want > cap ? cap : want), 256, 0, st>>>(ctx->vox_acc, table_slots);
        DA3S_LAUNCH_CHECK(ctx);
        DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->vox_counters, 0, 256, st));
    }
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->vox_counters, 0, 16, st));    // voxels, dropped (the send ticket resets itself)
    ctx->vox_clean = false;
    ctx->vox_active = true;
    return DA3S_OK;
}

static int voxel_insert_launch(da3s_ctx* ctx, const da3s_voxel_job* jobs_dev, int n_jobs, long long max_n, int width,
                               const da3s_voxel_job& single, float voxel, void* stream) {
    long long tiles;
    if (width > 0) {
        const long long rows = (max_n + width - 1) / width;
        tiles = ((rows + 7) / 8) * ((width + 15) / 16);
    } else {
        tiles = (max_n + 127) / 128;
    }
    const long long chunks_per_job = (tiles + VI_TILES_PER_BLOCK - 1) / VI_TILES_PER_BLOCK;
    // exactly the blocks that are resident together, so that "block b takes chunks b, b + G, ..." is a sweep in order
    int per_sm = 0;
    DA3S_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxel_insert_kernel, VI_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const long long total = chunks_per_job * n_jobs, cap = (long long)ctx->sm_count * per_sm;
    voxel_insert_kernel<<<(unsigned int)(total > cap ? cap : total), VI_THREADS, 0, (cudaStream_t)stream>>>(
        jobs_dev, single, n_jobs, chunks_per_job, width, make_quant(voxel), ctx->vox_acc, ctx->vox_slots, ctx->vox_counters);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_voxel_insert(da3s_ctx* ctx, const float* xyz, const uint8_t* rgb, const uint8_t* mask,
                                 long long n, float voxel, void* stream) {
    if (!ctx || !xyz || n < 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    if (n == 0) return DA3S_OK;
    da3s_voxel_job single;
    single.xyz = xyz; single.rgb = rgb; single.mask = mask; single.n = n;
    return voxel_insert_launch(ctx, nullptr, 1, n, 0, single, voxel, stream);
}

extern "C" int da3s_voxel_insert_jobs(da3s_ctx* ctx, const da3s_voxel_job* jobs_dev, int n_jobs, long long max_n,
                                      int width, float voxel, void* stream) {
    if (!ctx || !jobs_dev || n_jobs < 0 || n_jobs > 65535 || max_n < 0 || width < 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    if (n_jobs == 0 || max_n == 0) return DA3S_OK;
    da3s_voxel_job none;
    none.xyz = nullptr; none.rgb = nullptr; none.mask = nullptr; none.n = 0;
    return voxel_insert_launch(ctx, jobs_dev, n_jobs, max_n, width, none, voxel, stream);
}

extern "C" int da3s_voxel_finish(da3s_ctx* ctx, float voxel, long long max_voxels, float* xyz_out, uint8_t* rgb_out,
                                 int32_t* count_out, long long* key_out, unsigned long long* n_voxels,
                                 unsigned long long* n_dropped, void* stream) {
    if (!ctx || !xyz_out || !count_out || !n_voxels || max_voxels <= 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_warps = (ctx->vox_slots + VC_PER_WARP - 1) / VC_PER_WARP;
    const int n_cblocks = (int)((n_warps + VC_THREADS / 32 - 1) / (VC_THREADS / 32));
    unsigned int* warp_counts = ctx->vox_groups;
    unsigned long long* warp_offsets = (unsigned long long*)(warp_counts + ((n_warps + 1) & ~1ll));
    unsigned long long* chunk_offsets = warp_offsets + n_warps;
    const int n_chunks = (int)((n_warps + VS_CHUNK - 1) / VS_CHUNK);
    voxel_count_kernel<<<(int)((n_warps + VC_THREADS - 1) / VC_THREADS), VC_THREADS, 0, st>>>(VOX_BITMAP(ctx->vox_acc, ctx->vox_slots), n_warps,
                                                                                             warp_counts);
    DA3S_LAUNCH_CHECK(ctx);
    voxel_scan_kernel<<<n_chunks, 1024, 0, st>>>(warp_counts, (int)n_warps, warp_offsets, chunk_offsets);
    DA3S_LAUNCH_CHECK(ctx);
    voxel_scan_top_kernel<<<1, 1024, 0, st>>>(chunk_offsets, n_chunks, ctx->vox_counters);
    DA3S_LAUNCH_CHECK(ctx);
    prof_begin(ctx, DA3S_TIMED_VOXEL_EMIT, st);
    voxel_emit_kernel<<<n_cblocks, VC_THREADS, 0, st>>>(ctx->vox_acc, ctx->vox_slots, warp_offsets, chunk_offsets, warp_counts, voxel, max_voxels,
                                                       xyz_out, rgb_out, count_out, key_out);
    prof_end(ctx, DA3S_TIMED_VOXEL_EMIT, st);
    DA3S_LAUNCH_CHECK(ctx);
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_voxels, ctx->vox_counters, 8, cudaMemcpyDeviceToDevice, st));
    if (n_dropped)
        DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_dropped, ctx->vox_counters + 1, 8, cudaMemcpyDeviceToDevice, st));
    ctx->vox_active = false;
    ctx->vox_clean = true;          // every occupied key and every bitmap word was reset in place by the compaction pass
    return DA3S_OK;
}
