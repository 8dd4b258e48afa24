// Throughput of packed fp32 (FFMA2 / FADD2 / FMUL2) against scalar FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu
// Prints warp-instructions per clock per SM for a given number of resident warps.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int kIters = 2048, kChains = 8;

template <int MODE>
__global__ void bench(float* out, long long* cycles, float seed) {
    float a[kChains], b[kChains];
    u64 p[kChains];
    for (int i = 0; i < kChains; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * i; p[i] = pk(a[i], b[i]); }
    const float w = seed * 0.999f;
    const u64 w2 = pk(w, -w);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) { a[i] = fma1(a[i], w, b[i]); }                               // FFMA
            if (MODE == 1) { p[i] = fma2(p[i], w2, p[i]); }                              // FFMA2
            if (MODE == 2) { a[i] = add1(a[i], b[i]); }                                  // FADD
            if (MODE == 3) { p[i] = add2(p[i], w2); }                                    // FADD2
            if (MODE == 4) { a[i] = fma1(a[i], w, b[i]); b[i] = add1(b[i], w); }         // FFMA + FADD
            if (MODE == 5) { p[i] = fma2(p[i], w2, p[i]); a[i] = fma1(a[i], w, b[i]); }  // FFMA2 + FFMA
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < kChains; ++i) { float x, y; up(p[i], x, y); s += a[i] + b[i] + x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int warps_per_sm, int sms) {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    bench<MODE><<<sms, warps_per_sm * 32>>>(out, cyc, 1.0f);
    bench<MODE><<<sms, warps_per_sm * 32>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < sms; ++i) mean += h[i]; mean /= sms;
    double instr = (double)kIters * kChains * per_iter * warps_per_sm;
    printf("%-14s warps/SM %2d : %.3f warp-instr/clk/SM (%.0f clk)\n", name, warps_per_sm, instr / mean, mean);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA", 1, w, sms);
        run<1>("FFMA2", 1, w, sms);
        run<2>("FADD", 1, w, sms);
        run<3>("FADD2", 1, w, sms);
        run<4>("FFMA+FADD", 2, w, sms);
        run<5>("FFMA2+FFMA", 2, w, sms);
    }
    return 0;
}
