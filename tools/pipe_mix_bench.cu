// Which pipes of an sm_100a SM overlap?  Packed fp32 (FFMA2) against shared-memory
// loads, warp shuffles and integer ALU work, each alone and interleaved 1:1.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix_bench pipe_mix_bench.cu
// Prints warp-instructions per clock per SM (all classes summed) per mode.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float2 lds64(unsigned addr) { float2 v; asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ float4 lds128(unsigned addr) { float4 v; asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ float shfl(float v, int lane) { float r; asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(v), "r"(lane)); return r; }
__device__ __forceinline__ int iadd(int a, int b) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

constexpr int kIters = 1024, kChains = 8;
enum { M_FFMA2, M_LDS64, M_LDS128, M_SHFL, M_IADD, M_FFMA2_LDS64, M_FFMA2_LDS128, M_FFMA2_SHFL, M_FFMA2_IADD, M_LDS64_SHFL, M_LDS128_SHFL, M_FFMA2_LDS128_IADD, M_COUNT };

template <int MODE>
__global__ void bench(float* out, long long* cycles, float seed) {
    extern __shared__ float smem[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem[i] = seed * i;
    __syncthreads();
    u64 p[kChains];
    float f[kChains];
    int n[kChains];
    for (int i = 0; i < kChains; ++i) { p[i] = pk(seed + i + threadIdx.x, seed * i); f[i] = seed * i; n[i] = i + threadIdx.x; }
    const u64 w2 = pk(seed * 0.999f, -seed * 0.999f);
    const unsigned base = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 1024;
    const unsigned base64 = (unsigned)__cvta_generic_to_shared(smem) + (threadIdx.x & 31) * 8 + (threadIdx.x >> 5) * 1024;
    const int src = (threadIdx.x + 5) & 31;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            constexpr bool F = MODE == M_FFMA2 || MODE == M_FFMA2_LDS64 || MODE == M_FFMA2_LDS128 || MODE == M_FFMA2_SHFL || MODE == M_FFMA2_IADD || MODE == M_FFMA2_LDS128_IADD;
            constexpr bool L64 = MODE == M_LDS64 || MODE == M_FFMA2_LDS64 || MODE == M_LDS64_SHFL;
            constexpr bool L128 = MODE == M_LDS128 || MODE == M_FFMA2_LDS128 || MODE == M_LDS128_SHFL || MODE == M_FFMA2_LDS128_IADD;
            constexpr bool S = MODE == M_SHFL || MODE == M_FFMA2_SHFL || MODE == M_LDS64_SHFL || MODE == M_LDS128_SHFL;
            constexpr bool I = MODE == M_IADD || MODE == M_FFMA2_IADD || MODE == M_FFMA2_LDS128_IADD;
            if (F) p[i] = fma2(p[i], w2, p[i]);
            if (L64) { float2 v = lds64(base64 + ((i + it) & 3) * 256); acc += v.x; }
            if (L128) { float4 v = lds128(base + ((i + it) & 1) * 512); acc += v.x; }
            if (S) f[i] = shfl(f[i], src);
            if (I) n[i] = iadd(n[i], it);
        }
    }
    long long t1 = clock64();
    float s = acc;
    for (int i = 0; i < kChains; ++i) { float x, y; up(p[i], x, y); s += x + y + f[i] + n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int warps_per_sm, int sms) {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 2; ++rep) bench<MODE><<<sms, warps_per_sm * 32, 65536>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < sms; ++i) mean += h[i]; mean /= sms;
    double instr = (double)kIters * kChains * per_iter * warps_per_sm;
    printf("%-20s warps/SM %2d : %.3f warp-instr/clk/SM  (%.3f clk per group of %d)\n", name, warps_per_sm, instr / mean, mean / ((double)kIters * kChains * warps_per_sm), per_iter);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {8, 16}) {
        run<M_FFMA2>("FFMA2", 1, w, sms);
        run<M_LDS64>("LDS.64", 1, w, sms);
        run<M_LDS128>("LDS.128", 1, w, sms);
        run<M_SHFL>("SHFL", 1, w, sms);
        run<M_IADD>("IADD", 1, w, sms);
        run<M_FFMA2_LDS64>("FFMA2+LDS.64", 2, w, sms);
        run<M_FFMA2_LDS128>("FFMA2+LDS.128", 2, w, sms);
        run<M_FFMA2_SHFL>("FFMA2+SHFL", 2, w, sms);
        run<M_FFMA2_IADD>("FFMA2+IADD", 2, w, sms);
        run<M_LDS64_SHFL>("LDS.64+SHFL", 2, w, sms);
        run<M_LDS128_SHFL>("LDS.128+SHFL", 2, w, sms);
        run<M_FFMA2_LDS128_IADD>("FFMA2+LDS.128+IADD", 3, w, sms);
    }
    return 0;
}
