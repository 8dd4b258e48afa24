// Probe for the "transposed" conv formulation (groundwork for raising the conv
// stack above the 109 clk / N=80 MMA limit, profiles/r01w_ncu.md):
//   D^T[out channel (TMEM lane)][row (column)] = sum_{tap, ci} W[tap][ci][co] * X[row + tap - 1][ci]
// with the WEIGHTS as the M-side operand held in TMEM (tcgen05.cp from shared
// memory, then tcgen05.mma with A in TMEM) and the ACTIVATIONS as the N-side
// operand in the same K-major [k-group][row][8 ch] layout the current kernel
// uses (tap shift = 16-byte start offset).  Checks the numerics against the
// host and times the 15-MMA tile-layer for several N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ts_conv_probe tools/ts_conv_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int C = 80, KG = C / 8, KS = 3, MROWS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// smem: W operand  [tap][kg][128 rows (out ch)][8 in]  bf16   (3*10*128*16 = 61,440 B)
//       X operand  [kg][N + 2 rows][8 ch]              bf16
template <int N>
__global__ void __launch_bounds__(128, 1) probe(
    const __nv_bfloat16* __restrict__ w_packed, const __nv_bfloat16* __restrict__ x_packed,
    float* __restrict__ out, long long* cycles, int reps, int mode) {
    constexpr int RB = N + 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sw = smem;
    uint8_t* sx = smem + KS * KG * MROWS * 16;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < KS * KG * MROWS * 16 / 16; i += 128)
        reinterpret_cast<uint4*>(sw)[i] = reinterpret_cast<const uint4*>(w_packed)[i];
    for (int i = tid; i < KG * RB * 16 / 16; i += 128)
        reinterpret_cast<uint4*>(sx)[i] = reinterpret_cast<const uint4*>(x_packed)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmem_w = tmem + 256;          // 120 columns of weights
    // instruction descriptor: D fp32, A/B bf16, K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    uint32_t parity = 0;
    long long best = 1ll << 60;
    if (warp == 0) {
        for (int rep = 0; rep < reps; ++rep) {
            long long t0 = clock64();
            uint32_t elected;
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(elected));
            if (elected) {
                // weights -> TMEM: 15 chunks of 128 rows x 32 bytes (K = 16)
                // (mode 1: only on the first repetition, mode 2: copies only)
                if (mode == 3 && rep == 0) {      // first buffer filled once
#pragma unroll
                    for (int c = 0; c < 15; ++c) {
                        const uint64_t src = make_desc(
                            smem_u32(sw) + ((c / 5) * KG + 2 * (c % 5)) * MROWS * 16, MROWS * 16, 128);
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;"
                                     ::"r"(tmem_w + c * 8), "l"(src) : "memory");
                    }
                }
                if (mode == 0 || mode == 2 || (mode == 1 && rep == 0))
#pragma unroll
                for (int tap = 0; tap < KS; ++tap)
#pragma unroll
                    for (int kk = 0; kk < C / 16; ++kk) {
                        const uint64_t src = make_desc(
                            smem_u32(sw) + (tap * KG + 2 * kk) * MROWS * 16, MROWS * 16, 128);
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;"
                                     ::"r"(tmem_w + (tap * 5 + kk) * 8), "l"(src) : "memory");
                    }
                if (mode != 2)
#pragma unroll
                for (int tap = 0; tap < KS; ++tap)
#pragma unroll
                    for (int kk = 0; kk < C / 16; ++kk) {
                        const uint64_t db = make_desc(
                            smem_u32(sx) + (2 * kk) * RB * 16 + tap * 16, RB * 16, 128);
                        const uint32_t acc = (tap | kk) != 0;
                        asm volatile(
                            "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n\t}"
                            ::"r"(tmem), "r"(tmem_w + (tap * 5 + kk) * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    }
                if (mode == 3) {                  // next layer's weights into the OTHER buffer
#pragma unroll
                    for (int c = 0; c < 15; ++c) {
                        const uint64_t src = make_desc(
                            smem_u32(sw) + ((c / 5) * KG + 2 * (c % 5)) * MROWS * 16, MROWS * 16, 128);
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;"
                                     ::"r"(tmem_w + 120 + c * 8), "l"(src) : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
            } while (!done);
            parity ^= 1;
            long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (lane == 0) cycles[blockIdx.x] = best;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D^T: lane = out channel, column = row
    if (blockIdx.x == 0) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                  "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                  "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <int N>
static void run(int grid, int mode) {
    constexpr int RB = N + 2;
    std::vector<float> w(KS * C * C), x((size_t)RB * C);
    srand(1);
    for (auto& v : w) v = bf((rand() / (float)RAND_MAX - 0.5f) * 0.5f);
    for (auto& v : x) v = bf((rand() / (float)RAND_MAX - 0.5f) * 2.f);
    std::vector<__nv_bfloat16> wp((size_t)KS * KG * MROWS * 8, __float2bfloat16_rn(0.f)), xp((size_t)KG * RB * 8);
    for (int tap = 0; tap < KS; ++tap)
        for (int ci = 0; ci < C; ++ci)
            for (int co = 0; co < C; ++co)       // [tap][kg][row = co][8 in]
                wp[((size_t)(tap * KG + ci / 8) * MROWS + co) * 8 + ci % 8] =
                    __float2bfloat16_rn(w[(tap * C + ci) * C + co]);
    for (int r = 0; r < RB; ++r)
        for (int c = 0; c < C; ++c)              // [kg][row][8 ch]
            xp[((size_t)(c / 8) * RB + r) * 8 + c % 8] = __float2bfloat16_rn(x[(size_t)r * C + c]);
    __nv_bfloat16 *dw, *dx; float* dout; long long* dcyc;
    cudaMalloc(&dw, wp.size() * 2); cudaMalloc(&dx, xp.size() * 2);
    cudaMalloc(&dout, (size_t)MROWS * N * 4); cudaMalloc(&dcyc, grid * 8);
    cudaMemcpy(dw, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, xp.data(), xp.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)KS * KG * MROWS * 16 + (size_t)KG * RB * 16 + 1024;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<N><<<grid, 128, smem>>>(dw, dx, dout, dcyc, 20, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d ERROR %s\n", N, cudaGetErrorString(e)); exit(1); }
    std::vector<float> out((size_t)MROWS * N);
    std::vector<long long> cyc(grid);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int co = 0; co < C; ++co)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int tap = 0; tap < KS; ++tap)
                for (int ci = 0; ci < C; ++ci)
                    ref += (double)w[(tap * C + ci) * C + co] * x[(size_t)(n + tap) * C + ci];
            worst = fmax(worst, fabs(ref - out[(size_t)co * N + n]));
            scale = fmax(scale, fabs(ref));
        }
    long long sum = 0, mx = 0;
    for (auto c : cyc) { sum += c; if (c > mx) mx = c; }
    printf("N=%3d grid=%3d %-22s: max |err| %.3e (scale %.2f)  issue..commit seen: avg %.0f clk, max %lld clk\n",
           N, grid, mode == 0 ? "15 cp + 15 MMA" : mode == 1 ? "15 MMA (weights kept)" : mode == 2 ? "15 cp only" : "15 MMA, then 15 cp",
           (mode == 2) ? 0.0 : worst, scale, (double)sum / grid, mx);
    cudaFree(dw); cudaFree(dx); cudaFree(dout); cudaFree(dcyc);
}

int main() {
    for (int mode = 0; mode < 4; ++mode) {
        run<128>(148, mode);
        if (mode != 3) {                 // two weight buffers + N > 128 exceed the 512 columns
            run<192>(148, mode);
            run<256>(148, mode);
        }
    }
    return 0;
}
