// common.cuh — context, error plumbing and device helpers shared by every kernel file.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#include "../../include/da3s.h"

#define DA3S_SM_COUNT_FALLBACK 148

struct da3s_ctx {
    int device;
    int sm_count;
    size_t l2_bytes;
    unsigned char* ws;          // workspace base (cudaMalloc, 256-byte aligned)
    size_t ws_bytes;
    size_t ws_top;              // bump pointer for the current call
    size_t ws_floor;            // ws_reset() returns here: a caller's staging area below it survives nested entry points
    int last_cuda_error;
    unsigned long long launches;
    // voxel hash-table state (lives at the END of the workspace between begin/finish; layout: voxel.cu)
    unsigned long long* vox_acc;        // records [slots][8] u64, then the occupancy bitmap [slots/32] u32
    long long vox_slots;
    unsigned long long* vox_counters;   // [32] u64: [0] voxels emitted, [1] points dropped (table full), [2] send ticket
    unsigned int* vox_groups;           // per 512-slot group: occupied count [n] u32, then output offset [n] u64
    size_t vox_bytes;
    bool vox_clean, vox_active;
    // per-kernel timers (da3s_kernel_timers / da3s_kernel_time): event pairs around the named launches, a ring per kernel
    bool prof_on;
    cudaEvent_t prof_ev[DA3S_TIMED_KERNELS][DA3S_TIMER_RING][2];
    unsigned int prof_n[DA3S_TIMED_KERNELS];
    unsigned long long* prof_work;      // device [DA3S_TIMED_KERNELS]: work units the timed kernels report (RANSAC: evaluations)
};

// Event pair around one launch of a timed kernel (no-ops unless da3s_kernel_timers(ctx, 1) was called)
static inline void prof_begin(da3s_ctx* c, int which, cudaStream_t st) {
    if (c->prof_on) cudaEventRecord(c->prof_ev[which][c->prof_n[which] % DA3S_TIMER_RING][0], st);
}
static inline void prof_end(da3s_ctx* c, int which, cudaStream_t st) {
    if (c->prof_on) { cudaEventRecord(c->prof_ev[which][c->prof_n[which] % DA3S_TIMER_RING][1], st); c->prof_n[which]++; }
}

#define DA3S_CHECK_CUDA(ctx, expr)                                   \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) {                                     \
            (ctx)->last_cuda_error = (int)_e;                        \
            return DA3S_ECUDA;                                       \
        }                                                            \
    } while (0)

#define DA3S_LAUNCH_CHECK(ctx)                                       \
    do {                                                             \
        (ctx)->launches++;                                           \
        cudaError_t _e = cudaGetLastError();                         \
        if (_e != cudaSuccess) {                                     \
            (ctx)->last_cuda_error = (int)_e;                        \
            return DA3S_ECUDA;                                       \
        }                                                            \
    } while (0)

// ---- workspace bump allocator (per call; nothing is allocated after create) -------
static inline void ws_reset(da3s_ctx* c) { c->ws_top = c->ws_floor; }
static inline void* ws_alloc(da3s_ctx* c, size_t bytes) {
    size_t start = (c->ws_top + 255) & ~(size_t)255;
    size_t limit = c->ws_bytes - c->vox_bytes;      // the voxel table owns the tail while active
    if (start + bytes > limit) return nullptr;
    c->ws_top = start + bytes;
    return c->ws + start;
}
#define WS_ALLOC(ctx, T, var, count)                                             \
    T* var = (T*)ws_alloc((ctx), sizeof(T) * (size_t)(count));                   \
    if (!(var)) return DA3S_ENOMEM

__host__ __device__ static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// ---- device helpers ----------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit load that does not allocate in L1 (data is touched once per pass)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// order-preserving float32 <-> uint32 key (total order: -inf < ... < -0 < +0 < ... < +inf < NaN+)
__device__ __forceinline__ unsigned int f32_to_key(float f) {
    const unsigned int u = __float_as_uint(f);
    return u ^ ((unsigned int)((int)u >> 31) | 0x80000000u);        // negative: ~u, else u | sign bit (two instructions)
}
__device__ __forceinline__ float key_to_f32(unsigned int k) {
    unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// SPEC 1 (oracle/SPEC.md): float32 camera-frame unprojection, every op rounded once.
__device__ __forceinline__ void cam_fast(float u, float v, float d, float cu, float cv,
                                         float inv_fu, float inv_fv, float& x, float& y) {
    x = __fmul_rn(__fmul_rn(__fsub_rn(u, cu), d), inv_fu);
    y = __fmul_rn(__fmul_rn(__fsub_rn(v, cv), d), inv_fv);
}

__device__ __forceinline__ bool is_finite_f(float x) { return fabsf(x) <= 3.402823466e+38f; }
