// Second probe for the transposed conv formulation: a whole fused stack of L
// layers on one 192-row tile per CTA, weights in TMEM (tcgen05.cp), activations
// as the N-side operand, and the TRANSPOSED epilogue (TMEM lane = output
// channel): tcgen05.ld of 16 rows of one channel -> + bias -> ReLU -> bf16 ->
// 2-byte stores into the next layer's [k-group][row][8 ch] operand buffer.
// Checks the result against the host and times the MMA and epilogue phases.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ts_conv_stack_probe tools/ts_conv_stack_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int C = 80, KG = C / 8, KS = 3, MROWS = 128, N = 192, RB = N + 2, L = 7;
constexpr int W_BYTES = KS * KG * MROWS * 16;      // 61,440 per layer
constexpr int X_BYTES = KG * RB * 16;              // 31,040

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) probe(
    const __nv_bfloat16* __restrict__ w_packed,   // [L][tap][kg][128][8]
    const float* __restrict__ bias,               // [L][C]
    const __nv_bfloat16* __restrict__ x_packed,   // [kg][RB][8]
    float* __restrict__ out,                      // [N][C] of CTA 0
    long long* cycles,                            // [grid][3]: MMA phase, epilogue phase, weight phase (sums over layers)
    int mode) {                                   // 0: tcgen05.cp from shared memory, 1: tcgen05.st from registers
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sw = smem;
    uint8_t* sx = smem + W_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int THREADS = 32 * WARPS;
    constexpr int SPLIT = WARPS / 4;               // warps per TMEM lane quadrant
    for (int i = tid; i < X_BYTES / 16; i += THREADS)
        reinterpret_cast<uint4*>(sx)[i] = reinterpret_cast<const uint4*>(x_packed)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s, tmem_w = tmem + 256;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    uint32_t parity = 0;
    long long mma_clk = 0, epi_clk = 0, wgt_clk = 0;

    for (int layer = 0; layer < L; ++layer) {
        long long tw = clock64();
        if (mode == 0) {
            // this layer's weights: global -> shared (a plain copy; the product kernel streams them with TMA)
            for (int i = tid; i < W_BYTES / 16; i += THREADS)
                reinterpret_cast<uint4*>(sw)[i] =
                    reinterpret_cast<const uint4*>(w_packed + (size_t)layer * W_BYTES / 2)[i];
        } else {
            // weights straight from global (L2) into TMEM through registers: thread = TMEM lane = weight
            // row m; chunk c = (tap, kk) is the 32-byte K16 slice [kg 2kk | kg 2kk+1] of that row -> 8 columns
            const uint8_t* wl = reinterpret_cast<const uint8_t*>(w_packed) + (size_t)layer * W_BYTES;
            const uint32_t lane_base = tmem_w + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
            for (int c = 0; c < 15; ++c) {
                if (c % SPLIT != (warp >> 2)) continue;
                const uint8_t* src = wl + ((size_t)((c / 5) * KG + 2 * (c % 5)) * MROWS + (tid & 127)) * 16;
                const uint4 lo = *reinterpret_cast<const uint4*>(src);
                const uint4 hi = *reinterpret_cast<const uint4*>(src + MROWS * 16);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                    ::"r"(lane_base + c * 8), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
                      "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        long long t0 = clock64();
        wgt_clk += t0 - tw;
        if (warp == 0) {
            uint32_t elected;
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(elected));
            if (elected) {
                if (mode == 0)
#pragma unroll
                for (int c = 0; c < 15; ++c) {
                    const uint64_t src = make_desc(
                        smem_u32(sw) + ((c / 5) * KG + 2 * (c % 5)) * MROWS * 16, MROWS * 16, 128);
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_w + c * 8), "l"(src) : "memory");
                }
#pragma unroll
                for (int c = 0; c < 15; ++c) {
                    const int tap = c / 5, kk = c % 5;
                    const uint64_t db = make_desc(smem_u32(sx) + (2 * kk) * RB * 16 + tap * 16, RB * 16, 128);
                    asm volatile(
                        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n\t}"
                        ::"r"(tmem), "r"(tmem_w + c * 8), "l"(db), "r"(idesc), "r"((uint32_t)(c != 0)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
        }
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        } while (!done);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        long long t1 = clock64();

        // transposed epilogue: thread = TMEM lane = output channel
        const int c = tid & 127;
        const bool live = c < C;
        const float b = live ? bias[layer * C + c] : 0.f;
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint8_t* column = sx + ((c >> 3) * RB + 1) * 16 + (c & 7) * 2;     // row n -> + n * 16
        for (int n0 = (warp >> 2) * (N / SPLIT); n0 < ((warp >> 2) + 1) * (N / SPLIT); n0 += 32) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr + n0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (live) {
                if (layer + 1 < L) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = fmaxf(__uint_as_float(r[j]) + b, 0.f);
                        *reinterpret_cast<__nv_bfloat16*>(column + (n0 + j) * 16) = __float2bfloat16_rn(v);
                    }
                } else if (blockIdx.x == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        out[(size_t)(n0 + j) * C + c] = fmaxf(__uint_as_float(r[j]) + b, 0.f);
                }
            }
        }
        long long t2 = clock64();
        mma_clk += t1 - t0;
        epi_clk += t2 - t1;
    }
    if (tid == 0) {
        cycles[3 * blockIdx.x] = mma_clk;
        cycles[3 * blockIdx.x + 1] = epi_clk;
        cycles[3 * blockIdx.x + 2] = wgt_clk;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main() {
    std::vector<float> w((size_t)L * KS * C * C), bias((size_t)L * C), x((size_t)RB * C, 0.f);
    srand(3);
    for (auto& v : w) v = bf((rand() / (float)RAND_MAX - 0.5f) * 0.25f);
    for (auto& v : bias) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    for (int r = 1; r <= N; ++r)
        for (int c = 0; c < C; ++c) x[(size_t)r * C + c] = bf((rand() / (float)RAND_MAX - 0.5f) * 2.f);
    std::vector<__nv_bfloat16> wp((size_t)L * W_BYTES / 2, __float2bfloat16_rn(0.f)), xp((size_t)X_BYTES / 2);
    for (int l = 0; l < L; ++l)
        for (int tap = 0; tap < KS; ++tap)
            for (int ci = 0; ci < C; ++ci)
                for (int co = 0; co < C; ++co)
                    wp[(size_t)l * W_BYTES / 2 + ((size_t)(tap * KG + ci / 8) * MROWS + co) * 8 + ci % 8] =
                        __float2bfloat16_rn(w[((size_t)(l * KS + tap) * C + ci) * C + co]);
    for (int r = 0; r < RB; ++r)
        for (int c = 0; c < C; ++c)
            xp[((size_t)(c / 8) * RB + r) * 8 + c % 8] = __float2bfloat16_rn(x[(size_t)r * C + c]);
    // host reference with the same bf16 rounding of the activations between layers
    std::vector<float> cur = x, nxt((size_t)RB * C, 0.f);
    for (int l = 0; l < L; ++l) {
        for (int n = 1; n <= N; ++n)
            for (int co = 0; co < C; ++co) {
                double acc = bias[(size_t)l * C + co];
                for (int tap = 0; tap < KS; ++tap)
                    for (int ci = 0; ci < C; ++ci)
                        acc += (double)w[((size_t)(l * KS + tap) * C + ci) * C + co] *
                               cur[(size_t)(n + tap - 1) * C + ci];
                float v = acc > 0 ? (float)acc : 0.f;
                nxt[(size_t)n * C + co] = l + 1 < L ? bf(v) : v;
            }
        cur = nxt;
    }
    const int grid = 148;
    __nv_bfloat16 *dw, *dx; float *dbias, *dout; long long* dcyc;
    cudaMalloc(&dw, wp.size() * 2); cudaMalloc(&dx, xp.size() * 2);
    cudaMalloc(&dbias, bias.size() * 4); cudaMalloc(&dout, (size_t)N * C * 4); cudaMalloc(&dcyc, grid * 24);
    cudaMemcpy(dw, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, xp.data(), xp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dbias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = W_BYTES + X_BYTES + 1024;
    cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int run = 0; run < 3; ++run) {
        const int mode = run == 0 ? 0 : 1, warps = run == 2 ? 8 : 4;
        cudaMemset(dout, 0, (size_t)N * C * 4);
        if (warps == 4) probe<4><<<grid, 128, smem>>>(dw, dbias, dx, dout, dcyc, mode);
        else probe<8><<<grid, 256, smem>>>(dw, dbias, dx, dout, dcyc, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> out((size_t)N * C);
        std::vector<long long> cyc(3 * grid);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cyc.data(), dcyc, grid * 24, cudaMemcpyDeviceToHost);
        double worst = 0, scale = 0;
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < C; ++c) {
                worst = fmax(worst, fabs(out[(size_t)n * C + c] - cur[(size_t)(n + 1) * C + c]));
                scale = fmax(scale, fabs(cur[(size_t)(n + 1) * C + c]));
            }
        double mma = 0, epi = 0, wgt = 0;
        for (int i = 0; i < grid; ++i) { mma += cyc[3 * i]; epi += cyc[3 * i + 1]; wgt += cyc[3 * i + 2]; }
        printf("%s, %d warps: %d fused layers on a %d-row tile: max |err| %.3e (scale %.2f)\n",
               mode == 0 ? "weights via shared memory + tcgen05.cp" : "weights via registers + tcgen05.st   ",
               warps, L, N, worst, scale);
        printf("   per layer: weight phase (all threads) %.0f clk, tensor phase (%s15 MMA, commit, wait) %.0f clk, "
               "transposed epilogue %.0f clk\n",
               wgt / grid / L, mode == 0 ? "15 cp, " : "", mma / grid / L, epi / grid / L);
    }
    return 0;
}
