// write_fill.cu — does a write that misses in L2 fetch its line from DRAM?  (diagnostic, not product code)
// N random 128-byte-aligned lines of a 4 GB buffer are written in different ways; ncu reports dram__bytes_read per kernel:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_fill write_fill.cu
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ./write_fill
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long mix(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k;
}
__device__ __forceinline__ void st32(unsigned long long* p, unsigned long long v) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v), "l"(v + 1), "l"(v + 2), "l"(v + 3) : "memory");
}
// MODE 0: one thread, one 32-byte sector of the line          (32 of 128 bytes)
// MODE 1: one thread, two 32-byte stores = one 64-byte record  (two instructions)
// MODE 2: two lanes, 32 bytes each, ONE instruction = 64 bytes contiguous
// MODE 3: four lanes, 32 bytes each, ONE instruction = the whole 128-byte line
// MODE 4: one thread, four 32-byte stores = the whole line      (four instructions)
// MODE 5: one thread, one 8-byte store                           (partial sector)
template <int MODE>
__global__ void k(unsigned long long* buf, unsigned long long lines, unsigned long long n, unsigned long long salt) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 2) i >>= 1;
    if (MODE == 3) i >>= 2;
    if (i >= n) return;
    unsigned long long line = mix(i * 0x9E3779B97F4A7C15ull + salt) % lines;
    unsigned long long* p = buf + line * 16;
    if (MODE == 0) st32(p, i);
    if (MODE == 1) { st32(p, i); st32(p + 4, i); }
    if (MODE == 2) st32(p + 4 * (threadIdx.x & 1), i);
    if (MODE == 3) st32(p + 4 * (threadIdx.x & 3), i);
    if (MODE == 4) { st32(p, i); st32(p + 4, i); st32(p + 8, i); st32(p + 12, i); }
    if (MODE == 5) p[0] = i;
}

int main() {
    const unsigned long long bytes = 4ull << 30, lines = bytes / 128, n = 8ull << 20;
    unsigned long long* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes); cudaDeviceSynchronize();
    const int T = 256;
    k<0><<<(unsigned)((n + T - 1) / T), T>>>(buf, lines, n, 1); cudaDeviceSynchronize();
    k<1><<<(unsigned)((n + T - 1) / T), T>>>(buf, lines, n, 2); cudaDeviceSynchronize();
    k<2><<<(unsigned)((2 * n + T - 1) / T), T>>>(buf, lines, n, 3); cudaDeviceSynchronize();
    k<3><<<(unsigned)((4 * n + T - 1) / T), T>>>(buf, lines, n, 4); cudaDeviceSynchronize();
    k<4><<<(unsigned)((n + T - 1) / T), T>>>(buf, lines, n, 5); cudaDeviceSynchronize();
    k<5><<<(unsigned)((n + T - 1) / T), T>>>(buf, lines, n, 6); cudaDeviceSynchronize();
    printf("%llu random lines written per kernel (modes 0..5); useful bytes: 32, 64, 64, 128, 128, 8 per line\n", n);
    return 0;
}
