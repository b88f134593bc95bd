// voxel.cu — K6: voxel-grid downsample on a GPU hash grid (oracle/SPEC.md section 5).
//
// key    = floor(f64(p) / f64(voxel)) per axis, 21 bits per axis, bias 2^20
// record = 64 B, one L2 line sector pair: key | sum_qx | sum_qy | sum_qz | (count,sum_r) | (sum_g,sum_b)
//          sum_q = sum of llrint(frac * 2^32), frac = p/voxel - floor(p/voxel)  (exact int64)
// Integer accumulation makes the result independent of insertion order (bit-identical
// run to run and across GPUs), which floating-point atomics would not be.
//
// Contention is cut before it reaches L2: points of one warp that fall in the same voxel
// (neighbouring pixels of a frame usually do) are combined with a labelled warp partition
// and issue one set of atomics per distinct voxel.
//
// Algorithmic bytes: 12 (+3 rgb, +1 mask) per input point read; 12 (+3) + 4 (+8 key) per
// occupied voxel written.  Hash-table traffic (random 64 B records in L2/HBM) is what
// actually bounds this kernel and is not counted as algorithmic.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
namespace cg = cooperative_groups;

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
                    unsigned long long* __restrict__ dropped) {
    const double vd = (double)voxel;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        bool ok = is_finite_f(px) && is_finite_f(py) && is_finite_f(pz);
        if (mask) ok = ok && (mask[i] != 0);
        double qx = __ddiv_rn((double)px, vd), qy = __ddiv_rn((double)py, vd), qz = __ddiv_rn((double)pz, vd);
        double kx = floor(qx), ky = floor(qy), kz = floor(qz);
        ok = ok && (fabs(kx) < (double)VOX_BIAS) && (fabs(ky) < (double)VOX_BIAS) && (fabs(kz) < (double)VOX_BIAS);
        if (!ok) continue;
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
        // combine the lanes of this warp that hit the same voxel
        cg::coalesced_group active = cg::coalesced_threads();
        cg::coalesced_group same = cg::labeled_partition(active, key);
        sx = cg::reduce(same, sx, cg::plus<unsigned long long>());
        sy = cg::reduce(same, sy, cg::plus<unsigned long long>());
        sz = cg::reduce(same, sz, cg::plus<unsigned long long>());
        cr = cg::reduce(same, cr, cg::plus<unsigned long long>());
        gb = cg::reduce(same, gb, cg::plus<unsigned long long>());
        if (same.thread_rank() != 0) continue;
        unsigned long long slot = vox_hash(key) & (unsigned long long)(slots - 1);
        bool placed = false;
        for (int probe = 0; probe < VOX_MAX_PROBE; ++probe) {
            unsigned long long* rec = table + slot * VOX_REC;
            unsigned long long cur = *((volatile unsigned long long*)rec);
            if (cur == VOX_EMPTY) cur = atomicCAS(rec, VOX_EMPTY, key);
            if (cur == VOX_EMPTY || cur == key) {
                atomicAdd(rec + 1, sx); atomicAdd(rec + 2, sy); atomicAdd(rec + 3, sz);
                atomicAdd(rec + 4, cr);
                if (rgb) atomicAdd(rec + 5, gb);
                placed = true;
                break;
            }
            slot = (slot + 1) & (unsigned long long)(slots - 1);
        }
        if (!placed) atomicAdd(dropped, cr >> 32);
    }
}

__global__ void __launch_bounds__(256)
voxel_compact_kernel(const unsigned long long* __restrict__ table, long long slots, float voxel, long long max_voxels,
                     float* __restrict__ xyz_out, uint8_t* __restrict__ rgb_out, int32_t* __restrict__ count_out,
                     long long* __restrict__ key_out, unsigned long long* __restrict__ n_voxels) {
    const double vd = (double)voxel;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (long long)gridDim.x * blockDim.x) {
        const unsigned long long* rec = table + s * VOX_REC;
        unsigned long long key = rec[0];
        if (key == VOX_EMPTY) continue;
        unsigned long long o = atomicAdd(n_voxels, 1ull);
        if ((long long)o >= max_voxels) continue;
        const unsigned long long cr = rec[4], gb = rec[5];
        const unsigned long long cnt = cr >> 32;
        const double inv = 1.0 / 4294967296.0;
        double k[3] = {(double)((long long)((key >> 42) & 0x1FFFFF) - VOX_BIAS),
                       (double)((long long)((key >> 21) & 0x1FFFFF) - VOX_BIAS),
                       (double)((long long)(key & 0x1FFFFF) - VOX_BIAS)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // position = f32((k + (sum_q / count) / 2^32) * voxel), each step rounded once
            double mf = __dmul_rn(__ddiv_rn((double)rec[1 + a], (double)cnt), inv);
            xyz_out[3 * o + a] = (float)__dmul_rn(__dadd_rn(k[a], mf), vd);
        }
        if (rgb_out) {
            unsigned long long r = cr & 0xFFFFFFFFull, g = gb >> 32, b = gb & 0xFFFFFFFFull;
            rgb_out[3 * o]     = (uint8_t)((2 * r + cnt) / (2 * cnt));
            rgb_out[3 * o + 1] = (uint8_t)((2 * g + cnt) / (2 * cnt));
            rgb_out[3 * o + 2] = (uint8_t)((2 * b + cnt) / (2 * cnt));
        }
        count_out[o] = (int32_t)cnt;
        if (key_out) key_out[o] = (long long)key;
    }
}

extern "C" int da3s_voxel_begin(da3s_ctx* ctx, long long table_slots, void* stream) {
    if (!ctx || table_slots < 1024 || (table_slots & (table_slots - 1))) return DA3S_EINVAL;
    size_t bytes = (size_t)table_slots * VOX_REC * 8 + 256;
    if (bytes > ctx->ws_bytes) return DA3S_ENOMEM;
    ctx->vox_bytes = 0;
    ws_reset(ctx);
    // the table owns the tail of the workspace until da3s_voxel_finish
    size_t start = (ctx->ws_bytes - bytes) & ~(size_t)255;
    ctx->vox_bytes = ctx->ws_bytes - start;
    ctx->vox_keys = (unsigned long long*)(ctx->ws + start);
    ctx->vox_dropped = ctx->vox_keys + (size_t)table_slots * VOX_REC;
    ctx->vox_slots = table_slots;
    cudaStream_t st = (cudaStream_t)stream;
    long long words = table_slots * VOX_REC;
    int blocks = (int)((words + 255) / 256 > (long long)ctx->sm_count * 32 ? (long long)ctx->sm_count * 32 : (words + 255) / 256);
    voxel_clear_kernel<<<blocks, 256, 0, st>>>(ctx->vox_keys, table_slots);
    DA3S_LAUNCH_CHECK(ctx);
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->vox_dropped, 0, 8, st));
    return DA3S_OK;
}

extern "C" int da3s_voxel_insert(da3s_ctx* ctx, const float* xyz, const uint8_t* rgb, const uint8_t* mask,
                                 long long n, float voxel, void* stream) {
    if (!ctx || !xyz || n < 0 || !(voxel > 0.0f)) return DA3S_EINVAL;
    if (!ctx->vox_slots) return DA3S_EINVAL;
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
    if (!ctx->vox_slots) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    DA3S_CHECK_CUDA(ctx, cudaMemsetAsync(n_voxels, 0, 8, st));
    long long want = (ctx->vox_slots + 255) / 256, cap = (long long)ctx->sm_count * 32;
    int blocks = (int)(want > cap ? cap : want);
    voxel_compact_kernel<<<blocks, 256, 0, st>>>(ctx->vox_keys, ctx->vox_slots, voxel, max_voxels, xyz_out, rgb_out,
                                                 count_out, key_out, n_voxels);
    DA3S_LAUNCH_CHECK(ctx);
    if (n_dropped)
        DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(n_dropped, ctx->vox_dropped, 8, cudaMemcpyDeviceToDevice, st));
    ctx->vox_slots = 0;
    ctx->vox_bytes = 0;                      // the tail of the workspace is free again (stream order)
    return DA3S_OK;
}
