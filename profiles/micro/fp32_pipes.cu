// fp32_pipes.cu — what the float32 pipe of this GPU sustains per instruction form (a diagnostic, not product code):
//   scalar FFMA with three register operands, scalar FFMA with an immediate, packed FFMA2, and the RANSAC-like mix
//   (FFMA2 + FSET + FADD2).  Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
// Prints TFLOP/s (2 flop per FMA lane) and warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float s, float* sink) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(s + threadIdx.x + i, s - i);
    float2 m = make_float2(1.0000001f + s * 1e-9f, 0.9999999f + s * 1e-9f), c = make_float2(1e-7f * s, -1e-7f * s);
    float2 cnt = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 128; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }          // 2 scalar FFMA, 3 regs
                if (MODE == 1) { a[i].x = fmaf(a[i].x, 1.0000001f, 1e-7f); a[i].y = fmaf(a[i].y, 0.9999999f, -1e-7f); }   // immediates
                if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); }                                              // 1 packed FFMA2
                if (MODE == 3) {                                                                               // RANSAC-like: 15 FP2 : 2 FSET : 1 FADD2
                    a[i] = __ffma2_rn(a[i], m, c);
                    if ((i & 7) == 7) {
                        float x, y;
                        asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(x) : "f"(a[i].x), "f"(s));
                        asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(y) : "f"(a[i].y), "f"(s));
                        cnt = __fadd2_rn(cnt, make_float2(x, y));
                    }
                }
            }
        }
    }
    float t = cnt.x + cnt.y;
    for (int i = 0; i < 8; ++i) t += a[i].x + a[i].y;
    if (t == 123.456f) sink[0] = t;
}

template <int MODE>
void run(const char* name, int sms, double clock_ghz) {
    float* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, iters = 64;
    k<MODE><<<blocks, 256>>>(2, 1.f, sink);
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(iters, 1.f, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double fma_lanes = 2.0 * 8 * 128 * (double)iters * 256.0 * blocks;          // FMA lane-operations
    const double fp_instr = (MODE == 2 || MODE == 3 ? 1.0 : 2.0) * 8 * 128 * (double)iters * 8.0 * blocks;   // warp instructions (FP only)
    printf("%-34s %7.3f ms  %6.1f TFLOP/s  %5.3f FP warp-instr/clk/SMSP (at %.3f GHz)\n", name, best, 2.0 * fma_lanes / (best * 1e-3) / 1e12,
           fp_instr / (best * 1e-3) / (clock_ghz * 1e9) / (sms * 4.0), clock_ghz);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<0>("scalar FFMA, 3 register operands", p.multiProcessorCount, ghz);
    run<1>("scalar FFMA, immediate operands", p.multiProcessorCount, ghz);
    run<2>("packed FFMA2", p.multiProcessorCount, ghz);
    run<3>("FFMA2 + (2 FSET + FADD2) per 8", p.multiProcessorCount, ghz);
    return 0;
}
