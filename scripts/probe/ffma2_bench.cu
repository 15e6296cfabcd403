// Issue-rate probe: fma.rn.f32 vs fma.rn.f32x2 (and add / mul) on sm_100a.  nvcc -arch=sm_100a -o ffma2 ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float x = threadIdx.x * 1e-3f;
    float a[16];
    unsigned long long p[8];
    for (int i = 0; i < 16; ++i) a[i] = x + i;
    for (int i = 0; i < 8; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const unsigned long long m = pack(1.0001f, 0.9999f), c = pack(1e-3f, -1e-3f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 1.0001f, 1e-3f);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], m, c);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] + 1e-3f;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], c);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    for (int i = 0; i < 8; ++i) s += (float)(p[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const char* names[4] = {"FFMA  x16 scalar", "FFMA2 x8 packed ", "FADD  x16 scalar", "FADD2 x8 packed "};
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 4; ++mode) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
                if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
                if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
                if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            // 16 fp32 results per thread and iteration
            printf("%s  warps/SM %2d: %8lld cycles, %.2f fp32 results per clock per SM\n", names[mode], warps, h,
                   16.0 * iters * warps * 32 / (double)h);
        }
    }
    return 0;
}
