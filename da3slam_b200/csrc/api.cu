// api.cu — context lifetime, error strings, and the host-buffer entry point.
#include "common.cuh"
#include <new>

extern "C" int da3s_version(void) { return DA3S_VERSION; }

extern "C" const char* da3s_strerror(int code) {
    switch (code) {
        case DA3S_OK: return "ok";
        case DA3S_EINVAL: return "invalid argument";
        case DA3S_EALIGN: return "pointer not 16-byte aligned";
        case DA3S_ENOMEM: return "context workspace too small";
        case DA3S_ECUDA: return "CUDA runtime error";
        case DA3S_ETOOFEW: return "too few usable points";
        default: return "unknown da3s error";
    }
}

extern "C" int da3s_create(int device, size_t workspace_bytes, da3s_ctx** out) {
    if (!out) return DA3S_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return DA3S_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return DA3S_ECUDA;
    da3s_ctx* c = new (std::nothrow) da3s_ctx();
    if (!c) return DA3S_ENOMEM;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return DA3S_ECUDA; }
    c->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : DA3S_SM_COUNT_FALLBACK;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    if (workspace_bytes < (1u << 20)) workspace_bytes = (1u << 20);
    c->ws_bytes = workspace_bytes;
    c->ws = nullptr;
    cudaError_t e = cudaMalloc((void**)&c->ws, workspace_bytes);
    if (e != cudaSuccess) { int rc = (e == cudaErrorMemoryAllocation) ? DA3S_ENOMEM : DA3S_ECUDA; cudaGetLastError(); delete c; return rc; }
    c->ws_top = 0; c->ws_floor = 0; c->last_cuda_error = 0; c->launches = 0;
    c->vox_acc = nullptr; c->vox_slots = 0; c->vox_counters = nullptr; c->vox_groups = nullptr; c->vox_bytes = 0; c->vox_clean = false; c->vox_active = false;
    *out = c;
    return DA3S_OK;
}

extern "C" int da3s_destroy(da3s_ctx* ctx) {
    if (!ctx) return DA3S_EINVAL;
    cudaSetDevice(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    for (int k = 0; k < DA3S_TIMED_KERNELS; ++k)
        for (int i = 0; i < DA3S_TIMER_RING; ++i)
            for (int j = 0; j < 2; ++j) if (ctx->prof_ev[k][i][j]) cudaEventDestroy(ctx->prof_ev[k][i][j]);
    if (ctx->prof_work) cudaFree(ctx->prof_work);
    delete ctx;
    return DA3S_OK;
}

extern "C" int da3s_last_cuda_error(const da3s_ctx* ctx) { return ctx ? ctx->last_cuda_error : 0; }
extern "C" unsigned long long da3s_launch_count(const da3s_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" int da3s_kernel_timers(da3s_ctx* ctx, int on) {
    if (!ctx) return DA3S_EINVAL;
    cudaSetDevice(ctx->device);
    if (on && !ctx->prof_ev[0][0][0]) {
        for (int k = 0; k < DA3S_TIMED_KERNELS; ++k)
            for (int i = 0; i < DA3S_TIMER_RING; ++i)
                for (int j = 0; j < 2; ++j) DA3S_CHECK_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[k][i][j]));
        DA3S_CHECK_CUDA(ctx, cudaMalloc((void**)&ctx->prof_work, sizeof(unsigned long long) * DA3S_TIMED_KERNELS));
    }
    if (ctx->prof_work) DA3S_CHECK_CUDA(ctx, cudaMemset(ctx->prof_work, 0, sizeof(unsigned long long) * DA3S_TIMED_KERNELS));
    for (int k = 0; k < DA3S_TIMED_KERNELS; ++k) ctx->prof_n[k] = 0;
    ctx->prof_on = on != 0;
    return DA3S_OK;
}

extern "C" int da3s_kernel_time(da3s_ctx* ctx, int which, double* sum_ms_out, int* timed_out, int* launches_out, double* work_out) {
    if (!ctx || which < 0 || which >= DA3S_TIMED_KERNELS || !sum_ms_out || !timed_out || !launches_out) return DA3S_EINVAL;
    *sum_ms_out = 0.0; *timed_out = 0; *launches_out = 0;
    if (work_out) *work_out = 0.0;
    if (!ctx->prof_ev[0][0][0]) return DA3S_OK;
    const unsigned int n = ctx->prof_n[which], kept = n < DA3S_TIMER_RING ? n : DA3S_TIMER_RING;
    for (unsigned int i = 0; i < kept; ++i) {
        const unsigned int e = (n - 1 - i) % DA3S_TIMER_RING;
        if (i == 0) DA3S_CHECK_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[which][e][1]));
        float ms = 0.f;
        DA3S_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[which][e][0], ctx->prof_ev[which][e][1]));
        *sum_ms_out += ms;
    }
    *timed_out = (int)kept;
    *launches_out = (int)n;
    if (work_out && ctx->prof_work && n) {      // the last launch has completed (event synchronised above)
        unsigned long long w = 0;
        DA3S_CHECK_CUDA(ctx, cudaMemcpy(&w, ctx->prof_work + which, sizeof(w), cudaMemcpyDeviceToHost));
        DA3S_CHECK_CUDA(ctx, cudaMemset(ctx->prof_work + which, 0, sizeof(w)));
        *work_out = (double)w;
    }
    ctx->prof_n[which] = 0;
    return DA3S_OK;
}

extern "C" int da3s_enable_peer_access(da3s_ctx* ctx, int peer_device) {
    if (!ctx || peer_device < 0) return DA3S_EINVAL;
    if (peer_device == ctx->device) return DA3S_OK;
    int prev = 0, can = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaDeviceCanAccessPeer(&can, ctx->device, peer_device);
    if (e == cudaSuccess && can) {
        e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    } else if (e == cudaSuccess) {
        e = cudaErrorPeerAccessUnsupported;
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) { ctx->last_cuda_error = (int)e; cudaGetLastError(); return DA3S_ECUDA; }
    return DA3S_OK;
}

// ---------------------------------------------------------------------------------
// float32 FMA peak of this device, measured (the roofline denominator of the ALU-bound RANSAC scoring kernel)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fp32_peak_kernel(int iters, float seed, float* sink) {
    // 8 independent packed chains per thread hide the 4-cycle FFMA latency; 8 resident blocks of 256 threads per SM
    // packed float32 (FFMA2 on register pairs: two IEEE FMAs per lane and instruction) — the form the RANSAC kernel uses
    float2 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = make_float2(seed + (float)(threadIdx.x + k), seed - (float)k);
    const float2 m = make_float2(1.0000001f + seed * 1e-9f, 0.9999999f), c = make_float2(1e-7f * seed, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 256; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = __ffma2_rn(a[k], m, c);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
    if (s == 123.456f) sink[0] = s;             // never true: keeps the chains alive
}

extern "C" int da3s_measure_fp32_peak(da3s_ctx* ctx, int iters, double* tflops_out, void* stream) {
    if (!ctx || !tflops_out || iters <= 0) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ws_reset(ctx);
    WS_ALLOC(ctx, float, sink, 1);
    cudaEvent_t e0, e1;
    DA3S_CHECK_CUDA(ctx, cudaEventCreate(&e0));
    DA3S_CHECK_CUDA(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->sm_count * 8;
    fp32_peak_kernel<<<blocks, 256, 0, st>>>(2, 1.0f, sink);            // warm-up
    DA3S_LAUNCH_CHECK(ctx);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        fp32_peak_kernel<<<blocks, 256, 0, st>>>(iters, 1.0f, sink);
        DA3S_LAUNCH_CHECK(ctx);
        cudaEventRecord(e1, st);
        DA3S_CHECK_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 4096.0 * (double)iters * 256.0 * (double)blocks;
        if (ms > 0.0f && flop / (ms * 1e-3) / 1e12 > best) best = flop / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops_out = best;
    return DA3S_OK;
}

// ---------------------------------------------------------------------------------
// host-buffer path: H2D copies, the device pipeline, D2H of the rows, one synchronise.
// ---------------------------------------------------------------------------------
__global__ void fill_pairs_kernel(da3s_pair* pairs, int n_pairs, long long M, int overlap,
                                  const float* dA, const float* cA, const float* dB, const float* cB,
                                  const da3s_cam* camA, const da3s_cam* camB) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    da3s_pair p;
    p.depth_a = dA + (size_t)i * M; p.conf_a = cA + (size_t)i * M;
    p.depth_b = dB + (size_t)i * M; p.conf_b = cB + (size_t)i * M;
    p.cam_a = camA + (size_t)i * overlap; p.cam_b = camB + (size_t)i * overlap;
    pairs[i] = p;
}

extern "C" int da3s_align_pairs_host(da3s_ctx* ctx, int n_pairs, int overlap, int H, int W,
                                     const float* depth_a, const float* conf_a, const float* K_a, const float* E_a,
                                     const float* depth_b, const float* conf_b, const float* K_b, const float* E_b,
                                     const da3s_align_opts* opts, const int32_t* sample_idx,
                                     double* sim3_rows, void* stream) {
    if (!ctx || !depth_a || !conf_a || !K_a || !E_a || !depth_b || !conf_b || !K_b || !E_b || !opts || !sim3_rows)
        return DA3S_EINVAL;
    if (n_pairs <= 0 || overlap <= 0 || H <= 0 || W <= 0) return DA3S_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long M = (long long)overlap * H * W;
    if ((M * 4) % 16 != 0) return DA3S_EALIGN;             // packed pairs must keep 16-byte alignment
    const size_t map_bytes = sizeof(float) * (size_t)n_pairs * M;
    const size_t nf = (size_t)n_pairs * overlap;
    // input staging lives at the start of the workspace; da3s_align_pairs resets the bump pointer to
    // ctx->ws_floor, which is raised above the staging area for the duration of the nested call
    ctx->ws_floor = 0;
    ws_reset(ctx);
    WS_ALLOC(ctx, float, d_dA, (size_t)n_pairs * M);
    WS_ALLOC(ctx, float, d_cA, (size_t)n_pairs * M);
    WS_ALLOC(ctx, float, d_dB, (size_t)n_pairs * M);
    WS_ALLOC(ctx, float, d_cB, (size_t)n_pairs * M);
    WS_ALLOC(ctx, float, d_K, nf * 9 * 2);
    WS_ALLOC(ctx, float, d_E, nf * 12 * 2);
    WS_ALLOC(ctx, da3s_cam, d_cam, nf * 2);
    WS_ALLOC(ctx, da3s_pair, d_pairs, n_pairs);
    WS_ALLOC(ctx, double, d_rows, (size_t)n_pairs * DA3S_ROW_LEN);
    int32_t* d_idx = nullptr;
    if (opts->n_hyp > 0) {
        if (!sample_idx) return DA3S_EINVAL;
        WS_ALLOC(ctx, int32_t, tmp, (size_t)n_pairs * opts->n_hyp * 3);
        d_idx = tmp;
    }
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dA, depth_a, map_bytes, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_cA, conf_a, map_bytes, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dB, depth_b, map_bytes, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_cB, conf_b, map_bytes, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_K, K_a, sizeof(float) * nf * 9, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_K + nf * 9, K_b, sizeof(float) * nf * 9, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_E, E_a, sizeof(float) * nf * 12, cudaMemcpyHostToDevice, st));
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_E + nf * 12, E_b, sizeof(float) * nf * 12, cudaMemcpyHostToDevice, st));
    if (d_idx)
        DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(d_idx, sample_idx, sizeof(int32_t) * (size_t)n_pairs * opts->n_hyp * 3,
                                             cudaMemcpyHostToDevice, st));
    int rc = da3s_build_cams(ctx, d_K, d_E, (int)(2 * nf), DA3S_CAM_CLOSED_FORM, d_cam, stream);
    if (rc != DA3S_OK) return rc;
    fill_pairs_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(d_pairs, n_pairs, M, overlap, d_dA, d_cA, d_dB, d_cB,
                                                             d_cam, d_cam + nf);
    DA3S_LAUNCH_CHECK(ctx);
    // run the device pipeline on the remainder of the workspace (an active voxel table keeps its tail)
    ctx->ws_floor = ctx->ws_top;
    rc = da3s_align_pairs(ctx, d_pairs, n_pairs, overlap, H, W, opts, d_idx, d_rows, nullptr, nullptr, stream);
    ctx->ws_floor = 0;
    if (rc != DA3S_OK) return rc;
    DA3S_CHECK_CUDA(ctx, cudaMemcpyAsync(sim3_rows, d_rows, sizeof(double) * (size_t)n_pairs * DA3S_ROW_LEN,
                                         cudaMemcpyDeviceToHost, st));
    DA3S_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    return DA3S_OK;
}
