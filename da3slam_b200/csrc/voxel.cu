// voxel.cu — K6: voxel-grid downsample on a GPU hash grid (oracle/SPEC.md section 5).
//
// key    = floor(f64(p) / f64(voxel)) per axis, 21 bits per axis, bias 2^20
// record = 64 B, one L2 line sector pair: key | sum_qx | sum_qy | sum_qz | (count,sum_r) | (sum_g,sum_b)
//          sum_q = sum of llrint(frac * 2^32), frac = p/voxel - floor(p/voxel)  (exact int64)
// Integer accumulation makes the result independent of insertion order (bit-identical
// run to run and across GPUs), which floating-point atomics would not be.
//
// Contention is cut before it reaches L2: points of one warp that fall in the same voxel
// (neighbouring pixels of a frame usually do) are combined (match.any + shuffles) and issue
// one set of atomics per distinct voxel.  Compaction streams the table once and resets the
// records it emits, so no separate clearing pass is needed between uses.
//
// Algorithmic bytes: 12 (+3 rgb, +1 mask) per input point read; 12 (+3) + 4 (+8 key) per
// occupied voxel written.  Hash-table traffic (random 64 B records in L2/HBM) is what
// actually bounds this kernel and is not counted as algorithmic.
#include "common.cuh"

#define VOX_EMPTY 0xFFFFFFFFFFFFFFFFull
#define VOX_BIAS (1 << 20)
#define VOX_REC 8                           // u64 words per record
#define VOX_MAX_PROBE 4096

__device__ __forceinline__ unsigned long long vox_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__global__ void voxel_clear_kernel(unsigned long long* table, long long slots) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one u64 word per thread
    long long words = slots * VOX_REC;
    for (; i < words; i += (long long)gridDim.x * blockDim.x) table[i] = ((i & (VOX_REC - 1)) == 0) ? VOX_EMPTY : 0ull;
}

__global__ void __launch_bounds__(256)
voxel_insert_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ mask,
                    long long n, float voxel, unsigned long long* __restrict__ table, long long slots,
                    unsigned long long* __restrict__ counters /* [0]=voxels (set by finish) [1]=dropped */) {
    const double vd = (double)voxel;
    const unsigned int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mask && mask[i] == 0) continue;                    // masked points never touch their coordinates
        const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        if (!(is_finite_f(px) && is_finite_f(py) && is_finite_f(pz))) continue;
        double qx = __ddiv_rn((double)px, vd), qy = __ddiv_rn((double)py, vd), qz = __ddiv_rn((double)pz, vd);
        double kx = floor(qx), ky = floor(qy), kz = floor(qz);
        if (!((fabs(kx) < (double)VOX_BIAS) && (fabs(ky) < (double)VOX_BIAS) && (fabs(kz) < (double)VOX_BIAS))) continue;
        unsigned long long key = ((unsigned long long)((long long)kx + VOX_BIAS) << 42) |
                                 ((unsigned long long)((long long)ky + VOX_BIAS) << 21) |
                                 (unsigned long long)((long long)kz + VOX_BIAS);
        unsigned long long sx = (unsigned long long)__double2ll_rn((qx - kx) * 4294967296.0);
        unsigned long long sy = (unsigned long long)__double2ll_rn((qy - ky) * 4294967296.0);
        unsigned long long sz = (unsigned long long)__double2ll_rn((qz - kz) * 4294967296.0);
        unsigned long long cr = 1ull << 32, gb = 0ull;
        if (rgb) {
            cr |= (unsigned long long)rgb[3 * i];
            gb = ((unsigned long long)rgb[3 * i + 1] << 32) | (unsigned long long)rgb[3 * i + 2];
        }
        // combine the lanes of this warp that hit the same voxel; a lane alone in its voxel (the
        // common case) skips the exchange entirely
        const unsigned int act = __activemask();
        const unsigned int peers = __match_any_sync(act, key);
        const unsigned int leader = __ffs(peers) - 1;
        if (peers != (1u << lane)) {
            unsigned int rest = peers & ~(1u << leader);            // identical for every lane of the group
            while (rest) {
                const int src = __ffs(rest) - 1;
                rest &= rest - 1;
                const unsigned long long ax = __shfl_sync(peers, sx, src), ay = __shfl_sync(peers, sy, src);
                const unsigned long long az = __shfl_sync(peers, sz, src), ac = __shfl_sync(peers, cr, src);
                const unsigned long long ag = __shfl_sync(peers, gb, src);
                if (lane == leader) { sx += ax; sy += ay; sz += az; cr += ac; gb += ag; }
            }
            if (lane != leader) continue;
        }
        unsigned long long slot = vox_hash(key) & (unsigned long long)(slots - 1);
        bool placed = false;
        for (int probe = 0; probe < VOX_MAX_PROBE; ++probe) {
            unsigned long long* rec = table + slot * VOX_REC;
            unsigned long long cur = *((volatile unsigned long long*)rec);
            if (cur == VOX_EMPTY) {
                cur = atomicCAS(rec, VOX_EMPTY, key);
                if (cur == VOX_EMPTY) cur = key;
            }
            if (cur == key) {
                atomicAdd(rec + 1, sx); atomicAdd(rec + 2, sy); atomicAdd(rec + 3, sz);
                atomicAdd(rec + 4, cr);
                if (rgb) atomicAdd(rec + 5, gb);
                placed = true;
                break;
            }
            slot = (slot + 1) & (unsigned long long)(slots - 1);
        }
        if (!placed) atomicAdd(&counters[1], cr >> 32);
    }
}

// Compaction: one streaming pass over the table.  Empty slots cost one 32-byte sector; an
// occupied record is emitted and RESET in place, so the table is clean for the next begin()
// without a separate clearing pass.  Output positions come from a block-wide prefix sum and ONE
// global atomic per block iteration (1024 slots) — a per-warp atomic on a single counter
// serialises at L2 and was the whole cost of this kernel (profiles/r1_*).
#define VC_THREADS 256
#define VC_PER_THREAD 4
__global__ void __launch_bounds__(VC_THREADS)
voxel_compact_kernel(unsigned long long* __restrict__ table, long long slots, unsigned long long* __restrict__ counters,
                     float voxel, long long max_voxels, float* __restrict__ xyz_out, uint8_t* __restrict__ rgb_out,
                     int32_t* __restrict__ count_out, long long* __restrict__ key_out) {
    __shared__ unsigned int warp_tot[VC_THREADS / 32];
    __shared__ unsigned long long blk_base;
    const double vd = (double)voxel;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long chunk = (long long)VC_THREADS * VC_PER_THREAD;
    for (long long c0 = (long long)blockIdx.x * chunk; c0 < slots; c0 += (long long)gridDim.x * chunk) {     // block-uniform
        unsigned long long keys[VC_PER_THREAD];
        unsigned int mine = 0;
#pragma unroll
        for (int j = 0; j < VC_PER_THREAD; ++j) {
            const long long s = c0 + (long long)j * VC_THREADS + threadIdx.x;
            keys[j] = (s < slots) ? table[(size_t)s * VOX_REC] : VOX_EMPTY;
            mine += (keys[j] != VOX_EMPTY);
        }
        // exclusive prefix of `mine` over the block
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < VC_THREADS / 32; ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
        if (threadIdx.x == 0 && total) blk_base = atomicAdd(&counters[0], (unsigned long long)total);
        __syncthreads();
        unsigned long long o = blk_base + before + (incl - mine);
        if (total) {
#pragma unroll
            for (int j = 0; j < VC_PER_THREAD; ++j) {
                const unsigned long long key = keys[j];
                if (key == VOX_EMPTY) continue;
                const long long s = c0 + (long long)j * VC_THREADS + threadIdx.x;
                unsigned long long* rec = table + (size_t)s * VOX_REC;
                const unsigned long long s_x = rec[1];
                const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(rec + 2);     // sum_y, sum_z
                const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(rec + 4);     // (n, sum_r), (sum_g, sum_b)
                *reinterpret_cast<ulonglong2*>(rec) = make_ulonglong2(VOX_EMPTY, 0ull);
                *reinterpret_cast<ulonglong2*>(rec + 2) = make_ulonglong2(0ull, 0ull);
                *reinterpret_cast<ulonglong2*>(rec + 4) = make_ulonglong2(0ull, 0ull);
                const unsigned long long oo = o++;
                if ((long long)oo >= max_voxels) continue;
                const unsigned long long cr = b.x, gb = b.y, cnt = cr >> 32;
                const double inv = 1.0 / 4294967296.0;
                const unsigned long long sq[3] = {s_x, a.x, a.y};
                double k[3] = {(double)((long long)((key >> 42) & 0x1FFFFF) - VOX_BIAS),
                               (double)((long long)((key >> 21) & 0x1FFFFF) - VOX_BIAS),
                               (double)((long long)(key & 0x1FFFFF) - VOX_BIAS)};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    // position = f32((k + (sum_q / count) / 2^32) * voxel), each step rounded once
                    double mf = __dmul_rn(__ddiv_rn((double)sq[q], (double)cnt), inv);
                    xyz_out[3 * oo + q] = (float)__dmul_rn(__dadd_rn(k[q], mf), vd);
                }
                if (rgb_out) {
                    // (2*sum + n) / (2*n) = floor(sum/n) + (2*rem >= n), in 32-bit arithmetic (64-bit integer
                    // division is emulated and was a large part of this kernel)
                    const unsigned int n32 = (unsigned int)cnt;
                    const unsigned int ch[3] = {(unsigned int)(cr & 0xFFFFFFFFull), (unsigned int)(gb >> 32), (unsigned int)(gb & 0xFFFFFFFFull)};
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const unsigned int d = ch[q] / n32, rem = ch[q] - d * n32;
                        rgb_out[3 * oo + q] = (uint8_t)(d + ((rem >= n32 - rem) ? 1u : 0u));
                    }
                }
                count_out[oo] = (int32_t)cnt;
                if (key_out) key_out[oo] = (long long)key;
            }
        }
        __syncthreads();                                       // blk_base / warp_tot are reused next iteration
    }
}

extern "C" int da3s_voxel_begin(da3s_ctx* ctx, long long table_slots, void* stream) {
    if (!ctx || table_slots < 1024 || (table_slots & (table_slots - 1)) || table_slots > (1ll << 31)) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    // layout of the reserved tail: records [slots][8] u64 | counters [4] u64
    size_t bytes = (size_t)table_slots * VOX_REC * 8 + 256;
    const bool reuse = (ctx->vox_slots == table_slots) && ctx->vox_clean && ctx->vox_bytes > 0;
    if (!reuse) {
        if (bytes > ctx->ws_bytes) return DA3S_ENOMEM;
        ctx->vox_bytes = 0;
        ws_reset(ctx);
        size_t start = (ctx->ws_bytes - bytes) & ~(size_t)255;
        ctx->vox_bytes = ctx->ws_bytes - start;               // stays reserved until a different size is requested
        ctx->vox_keys = (unsigned long long*)(ctx->ws + start);
        ctx->vox_dropped = ctx->vox_keys + (size_t)table_slots * VOX_REC;      // counters[4]
        ctx->vox_occ = nullptr;
        ctx->vox_slots = table_slots;
        long long words = table_slots * VOX_REC;
        long long want = (words + 255) / 256, cap = (long long)ctx->sm_count * 32;
        voxel_clear_kernel<<<(int)(want > cap ? cap : want), 256, 0, st>>>(ctx->vox_keys, table_slots);
        DA3S_LAUNCH_CHECK(ctx);
    }
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->vox_dropped, 0, 32, st));
    ctx->vox_clean = false;
    ctx->vox_active = true;
    return DA3S_OK;
}

extern "C" int da3s_voxel_insert(da3s_ctx* ctx, const float* xyz, const uint8_t* rgb, const uint8_t* mask,
                                 long long n, float voxel, void* stream) {
    if (!ctx || !xyz || n < 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    if (n == 0) return DA3S_OK;
    long long want = (n + 255) / 256, cap = (long long)ctx->sm_count * 32;
    int blocks = (int)(want > cap ? cap : want);
    voxel_insert_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(xyz, rgb, mask, n, voxel, ctx->vox_keys, ctx->vox_slots,
                                                                  ctx->vox_dropped);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_voxel_finish(da3s_ctx* ctx, float voxel, long long max_voxels, float* xyz_out, uint8_t* rgb_out,
                                 int32_t* count_out, long long* key_out, unsigned long long* n_voxels,
                                 unsigned long long* n_dropped, void* stream) {
    if (!ctx || !xyz_out || !count_out || !n_voxels || max_voxels <= 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    voxel_compact_kernel<<<ctx->sm_count * 8, VC_THREADS, 0, st>>>(ctx->vox_keys, ctx->vox_slots, ctx->vox_dropped, voxel, max_voxels,
                                                            xyz_out, rgb_out, count_out, key_out);
    DA3S_LAUNCH_CHECK(ctx);
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_voxels, ctx->vox_dropped, 8, cudaMemcpyDeviceToDevice, st));
    if (n_dropped)
        DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_dropped, ctx->vox_dropped + 1, 8, cudaMemcpyDeviceToDevice, st));
    ctx->vox_active = false;
    ctx->vox_clean = true;          // every occupied record was reset in place by the compaction pass
    return DA3S_OK;
}
