// voxel.cu — K6: voxel-grid downsample on a GPU hash grid (oracle/SPEC.md section 5).
//
// key    = floor(f64(p) / f64(voxel)) per axis, 21 bits per axis, bias 2^20
// table  = keys[slots] u64, DENSE (compaction scans 8 B per slot instead of a 64-B DRAM atom)
//        + acc[slots] 64-byte records: sum_qx | sum_qy | sum_qz | (count,sum_r) | (sum_g,sum_b) | pad
//          sum_q = sum of llrint(frac * 2^32), frac = p/voxel - floor(p/voxel)  (exact int64)
// Integer accumulation makes the result independent of insertion order (bit-identical
// run to run and across GPUs), which floating-point atomics would not be.
//
// Contention is cut before it reaches L2: points of one warp that fall in the same voxel
// (neighbouring pixels of a frame usually do) are combined (match.any + shuffles) and issue
// one set of atomics per distinct voxel.  Compaction (count / scan / emit over the dense key
// array) resets what it emits, so no separate clearing pass is needed between uses.
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

__global__ void voxel_clear_kernel(unsigned long long* keys, unsigned long long* acc, long long slots) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one u64 word per thread
    const long long words = slots * (VOX_REC + 1);
    for (; i < words; i += (long long)gridDim.x * blockDim.x) {
        if (i < slots) keys[i] = VOX_EMPTY; else acc[i - slots] = 0ull;
    }
}

__global__ void __launch_bounds__(256)
voxel_insert_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ mask,
                    long long n, float voxel, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ acc,
                    long long slots, unsigned long long* __restrict__ counters /* [0]=voxels (set by finish) [1]=dropped */) {
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
            unsigned long long cur = *((volatile unsigned long long*)(keys + slot));
            if (cur == VOX_EMPTY) {
                cur = atomicCAS(keys + slot, VOX_EMPTY, key);
                if (cur == VOX_EMPTY) cur = key;
            }
            if (cur == key) {
                unsigned long long* rec = acc + slot * VOX_REC;
                atomicAdd(rec + 0, sx); atomicAdd(rec + 1, sy); atomicAdd(rec + 2, sz);
                atomicAdd(rec + 3, cr);
                if (rgb) atomicAdd(rec + 4, gb);
                placed = true;
                break;
            }
            slot = (slot + 1) & (unsigned long long)(slots - 1);
        }
        if (!placed) atomicAdd(&counters[1], cr >> 32);
    }
}

// Compaction without atomics or barriers, two streaming passes over the DENSE key array:
//   count  every WARP counts the occupied slots of its fixed range of 512 slots
//   scan   one block turns the per-warp counts into output offsets (and the total)
//   emit   every warp re-reads its 16 x 32 keys (all loads in flight together), derives each
//          occupied slot's output index from warp ballots + its offset, then reads, emits and
//          RESETS the records, so the table is clean for the next begin() without a clearing
//          pass.  No __syncthreads: the previous block-synchronous version was bound by exposed
//          DRAM latency (profiles/r1_*: long_scoreboard 70, issue 8 %).
// Output order = slot order: deterministic for a given table size.
#define VC_THREADS 256
#define VC_ROUNDS 16
#define VC_PER_WARP (32 * VC_ROUNDS)        // 512 slots per warp
#define VC_PER_BLOCK (VC_PER_WARP * VC_THREADS / 32)

__global__ void __launch_bounds__(VC_THREADS)
voxel_count_kernel(const unsigned long long* __restrict__ keys, long long slots, unsigned int* __restrict__ warp_counts) {
    const unsigned int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (VC_THREADS / 32) + (threadIdx.x >> 5);
    const long long base = wid * VC_PER_WARP;
    if (base >= slots) return;
    unsigned int c = 0;
#pragma unroll
    for (int j = 0; j < VC_ROUNDS / 2; ++j) {
        const long long s = base + ((long long)j * 32 + lane) * 2;       // two keys = one 16-byte load
        if (s + 1 < slots) {
            const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(keys + s);
            c += (k.x != VOX_EMPTY) + (k.y != VOX_EMPTY);
        } else if (s < slots) c += keys[s] != VOX_EMPTY;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) warp_counts[wid] = c;
}

__global__ void __launch_bounds__(1024)
voxel_scan_kernel(const unsigned int* __restrict__ counts, int n, unsigned long long* __restrict__ offsets,
                  unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n ? (unsigned long long)counts[i] : 0ull;
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
    if (threadIdx.x == 0) counters[0] = carry;
}

__global__ void __launch_bounds__(VC_THREADS)
voxel_emit_kernel(unsigned long long* __restrict__ keys, unsigned long long* __restrict__ acc, long long slots,
                  const unsigned long long* __restrict__ warp_offsets, float voxel, long long max_voxels,
                  float* __restrict__ xyz_out, uint8_t* __restrict__ rgb_out, int32_t* __restrict__ count_out,
                  long long* __restrict__ key_out) {
    const double vd = (double)voxel;
    const unsigned int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (VC_THREADS / 32) + (threadIdx.x >> 5);
    const long long base = wid * VC_PER_WARP;
    if (base >= slots) return;
    unsigned long long key[VC_ROUNDS];
#pragma unroll
    for (int j = 0; j < VC_ROUNDS; ++j) {                       // 16 independent coalesced loads per lane
        const long long s = base + (long long)j * 32 + lane;
        key[j] = s < slots ? keys[s] : VOX_EMPTY;
    }
    unsigned long long out = warp_offsets[wid];
#pragma unroll
    for (int j0 = 0; j0 < VC_ROUNDS; j0 += 4) {                 // 4 rounds at a time: their record loads overlap
        ulonglong2 ra[4], rb[4];
        unsigned long long rg[4], oo[4];
        bool occ[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + q;
            occ[q] = key[j] != VOX_EMPTY;
            const unsigned int m = __ballot_sync(0xffffffffu, occ[q]);
            oo[q] = out + __popc(m & ((1u << lane) - 1u));
            out += __popc(m);
            if (occ[q]) {
                const long long s = base + (long long)j * 32 + lane;
                unsigned long long* rec = acc + (size_t)s * VOX_REC;
                ra[q] = *reinterpret_cast<const ulonglong2*>(rec);          // sum_x, sum_y
                rb[q] = *reinterpret_cast<const ulonglong2*>(rec + 2);      // sum_z, (n, sum_r)
                rg[q] = rec[4];                                              // (sum_g, sum_b)
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!occ[q]) continue;
            const int j = j0 + q;
            const long long s = base + (long long)j * 32 + lane;
            unsigned long long* rec = acc + (size_t)s * VOX_REC;
            *reinterpret_cast<ulonglong2*>(rec) = make_ulonglong2(0ull, 0ull);
            *reinterpret_cast<ulonglong2*>(rec + 2) = make_ulonglong2(0ull, 0ull);
            rec[4] = 0ull;
            keys[s] = VOX_EMPTY;
            if ((long long)oo[q] >= max_voxels) continue;
            const unsigned long long k64 = key[j], cr = rb[q].y, gb = rg[q], cnt = cr >> 32, o = oo[q];
            const double inv = 1.0 / 4294967296.0;
            const unsigned long long sq[3] = {ra[q].x, ra[q].y, rb[q].x};
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
    }
}

extern "C" int da3s_voxel_begin(da3s_ctx* ctx, long long table_slots, void* stream) {
    if (!ctx || table_slots < 1024 || (table_slots & (table_slots - 1)) || table_slots > (1ll << 31)) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    // layout of the reserved tail: keys [slots] u64 | acc [slots][8] u64 | counters [4] u64 (256 B) |
    // warp counts [slots/512] u32 | warp offsets [slots/512] u64
    const long long n_cblocks = (table_slots + VC_PER_WARP - 1) / VC_PER_WARP;
    size_t bytes = (size_t)table_slots * (VOX_REC + 1) * 8 + 256 + (size_t)n_cblocks * 16 + 512;
    const bool reuse = (ctx->vox_slots == table_slots) && ctx->vox_clean && ctx->vox_bytes > 0;
    if (!reuse) {
        if (bytes > ctx->ws_bytes) return DA3S_ENOMEM;
        ctx->vox_bytes = 0;
        ws_reset(ctx);
        size_t start = (ctx->ws_bytes - bytes) & ~(size_t)255;
        ctx->vox_bytes = ctx->ws_bytes - start;               // stays reserved until a different size is requested
        ctx->vox_keys = (unsigned long long*)(ctx->ws + start);
        ctx->vox_acc = ctx->vox_keys + (size_t)table_slots;
        ctx->vox_dropped = ctx->vox_acc + (size_t)table_slots * VOX_REC;       // counters[4]
        ctx->vox_occ = (unsigned int*)(ctx->vox_dropped + 32);                 // block counts, then block offsets
        ctx->vox_slots = table_slots;
        long long words = table_slots * (VOX_REC + 1);
        long long want = (words + 255) / 256, cap = (long long)ctx->sm_count * 32;
        voxel_clear_kernel<<<(int)(want > cap ? cap : want), 256, 0, st>>>(ctx->vox_keys, ctx->vox_acc, table_slots);
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
    voxel_insert_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(xyz, rgb, mask, n, voxel, ctx->vox_keys, ctx->vox_acc,
                                                                  ctx->vox_slots, ctx->vox_dropped);
    DA3S_LAUNCH_CHECK(ctx);
    return DA3S_OK;
}

extern "C" int da3s_voxel_finish(da3s_ctx* ctx, float voxel, long long max_voxels, float* xyz_out, uint8_t* rgb_out,
                                 int32_t* count_out, long long* key_out, unsigned long long* n_voxels,
                                 unsigned long long* n_dropped, void* stream) {
    if (!ctx || !xyz_out || !count_out || !n_voxels || max_voxels <= 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_active) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_warps = (int)((ctx->vox_slots + VC_PER_WARP - 1) / VC_PER_WARP);
    const int n_cblocks = (n_warps + VC_THREADS / 32 - 1) / (VC_THREADS / 32);
    unsigned int* warp_counts = ctx->vox_occ;
    unsigned long long* warp_offsets = (unsigned long long*)(warp_counts + ((n_warps + 1) & ~1));
    voxel_count_kernel<<<n_cblocks, VC_THREADS, 0, st>>>(ctx->vox_keys, ctx->vox_slots, warp_counts);
    DA3S_LAUNCH_CHECK(ctx);
    voxel_scan_kernel<<<1, 1024, 0, st>>>(warp_counts, n_warps, warp_offsets, ctx->vox_dropped);
    DA3S_LAUNCH_CHECK(ctx);
    voxel_emit_kernel<<<n_cblocks, VC_THREADS, 0, st>>>(ctx->vox_keys, ctx->vox_acc, ctx->vox_slots, warp_offsets, voxel, max_voxels,
                                                       xyz_out, rgb_out, count_out, key_out);
    DA3S_LAUNCH_CHECK(ctx);
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_voxels, ctx->vox_dropped, 8, cudaMemcpyDeviceToDevice, st));
    if (n_dropped)
        DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_dropped, ctx->vox_dropped + 1, 8, cudaMemcpyDeviceToDevice, st));
    ctx->vox_active = false;
    ctx->vox_clean = true;          // every occupied record was reset in place by the compaction pass
    return DA3S_OK;
}
