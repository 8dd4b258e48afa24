// Transposed tcgen05 conv stack for sm_100a (bf16 operands, fp32 accumulate),
// opt-in with EMPHASES_B200_TC=transposed.  Same op as conv_tc.cu -- n_layers x
// [Conv1d(80 -> 80, k = 3, 'same') + activation] over packed rows, emphases/
// model/core.py:92-94 and model/layers/convolution.py:13-37 -- in the
// formulation profiles/r01z_ts_conv_probe.md validated:
//
//   D^T[out channel (TMEM lane)][row (column)] = sum_{tap, ci} W[tap][ci][co] X[row + tap - 1][ci]
//
// * The WEIGHTS are the M-side operand and live in TMEM (120 columns per
//   layer, double buffered).  A weight warpgroup writes them there through
//   registers (global / L2 -> tcgen05.st), so the tensor pipe runs nothing but
//   MMAs and there is no weight ring in shared memory.
// * The ACTIVATIONS are the N-side operand in the same K-major
//   [k-group][row][8 ch] shared-memory layout as conv_tc.cu (tap shift = 16-byte
//   start offset); with A in TMEM an MMA of N = 128 rows costs N / 2 = 64 clk
//   (the math floor) instead of the 109 clk of the N = 80 SS-mode MMA.
// * The epilogue is transposed: thread = TMEM lane = output channel, so the bias
//   is one register; two warps per lane quadrant drain half of the 128 rows
//   each and write the next layer's operand with 2-byte stores.
//
// Warp roles (672 threads, one persistent CTA per SM, two 128-row tiles in flight):
//   warps 0-7   epilogue of tile slot 0 (quadrant = warp % 4, column half = warp / 4)
//   warps 8-15  epilogue of tile slot 1
//   warps 16-19 weight group (quadrant = warp % 4): layer l+1's weights while layer l computes
//   warp 20     MMA issuer
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace emph {

namespace tct {

#ifdef EXP_TRACE
__device__ long long g_trace_t[12][256];
#define TRACE_T(ev, idx) do { if (blockIdx.x == 0 && (idx) < 256) g_trace_t[ev][idx] = clock64(); } while (0)
#else
#define TRACE_T(ev, idx) do {} while (0)
#endif

constexpr int C = 80;
constexpr int KS = 3;
constexpr int KG = C / 8;
constexpr int NT = 128;               // rows per tile = UMMA N
constexpr int RB = NT + 2;            // operand buffer rows (one pad row each side)
constexpr int kSlots = 2;
constexpr int kChunks = KS * C / 16;  // 15 K16 slices per layer
constexpr int kMaxLayers = 16;
constexpr int ACT_BYTES = KG * RB * 16;          // 20,800
constexpr int kEpiWarps = 8;                     // per slot
constexpr int kWeightWarp0 = kSlots * kEpiWarps; // 16
constexpr int kMmaWarp = kWeightWarp0 + 4;       // 20
constexpr int kThreads = 32 * (kMmaWarp + 1);    // 672
constexpr int kAccCols = 128;                    // accumulator columns per slot
constexpr int kWeightCol0 = kSlots * kAccCols;   // 256
constexpr int kWeightCols = kChunks * 8;         // 120 per buffer
constexpr int W_LAYER_BYTES = kChunks * 128 * 32;   // [chunk][128 rows][16 bf16]: 61,440

struct __align__(128) Smem {
    uint8_t act[kSlots][ACT_BYTES + 96];     // +96 keeps 128-byte alignment
    float bias[kMaxLayers][C];
    uint32_t row_mask[kSlots][4];            // bit n % 32 of word n / 32: row n is a real frame
    uint64_t act_ready[kSlots];
    uint64_t mma_done[kSlots];
    uint64_t w_ready[2];
    uint64_t w_free[2];
    uint32_t tmem_base;
};

struct Acts {
    int act[kMaxLayers];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b64 state;\n\t"
        "mbarrier.arrive.shared::cta.b64 state, [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t elected;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(elected));
    return elected != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

#ifdef TCT_STSM
// 16 TMEM lanes x 32 columns in the mma-fragment layout: thread t gets lane t / 4, columns
// 8 i + 2 (t % 4) and + 1 in r[4 i], r[4 i + 1], and lane t / 4 + 8, same columns, in r[4 i + 2], r[4 i + 3]
__device__ __forceinline__ void tmem_ld16x256(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Four 8 x 8 b16 fragments stored transposed: lane 8 k + j supplies the address of stored row j of matrix k
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
                 ::"r"(addr), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
}
#endif

// D fp32, A / B bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kInstrDesc =
    (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

#ifdef TCT_STSM
__device__ __noinline__ void activate_pairs(uint32_t (&r)[16], float b_lo, float b_hi, int a) {
#pragma unroll 1
    for (int j = 0; j < 16; ++j)
        r[j] = __float_as_uint(apply_activation(__uint_as_float(r[j]) + ((j & 2) ? b_hi : b_lo), a));
}
#endif

// Any activation other than ReLU (rolled and out of line on purpose)
__device__ __noinline__ void activate_rows(uint32_t (&r)[32], float b, int a) {
#pragma unroll 1
    for (int j = 0; j < 32; ++j)
        r[j] = __float_as_uint(apply_activation(__uint_as_float(r[j]) + b, a));
}

__global__ void __launch_bounds__(kThreads, 1)
conv_stack_tct_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ row_seq, int total_rows,
    const uint8_t* __restrict__ weights,   // [layer][chunk][128 rows][16 bf16]
    const float* __restrict__ bias,        // [layer][C] fp32
    Acts acts, int n_layers, int tile_rows, int n_tiles, float* __restrict__ y) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int halo = n_layers;                       // (KS - 1) / 2 per layer
    const int rounds = (n_tiles + gridDim.x * kSlots - 1) / (gridDim.x * kSlots);

    // ---- one-time setup ----
    for (int i = tid; i < kSlots * KG * 2 * 4; i += kThreads) {    // pad rows stay zero
        const int slot = i / (KG * 8), rem = i % (KG * 8);
        const int kg = rem / 8, edge = (rem / 4) & 1, word = rem & 3;
        reinterpret_cast<uint32_t*>(sm.act[slot] + (kg * RB + (edge ? RB - 1 : 0)) * 16)[word] = 0u;
    }
    for (int i = tid; i < n_layers * C; i += kThreads) sm.bias[i / C][i % C] = bias[i];
    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(&sm.act_ready[s], kEpiWarps);
            mbar_init(&sm.mma_done[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.w_ready[b], 4);
            mbar_init(&sm.w_free[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     ::"r"(smem_u32(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < kWeightWarp0) {
        // =========================== epilogue groups ===========================
        const int slot = warp / kEpiWarps;
        const int quad = warp & 3, half = (warp >> 2) & 1;
        const int gtid = tid - slot * (32 * kEpiWarps);        // 0..255 within the group
        const int c = quad * 32 + lane;                         // output channel = TMEM lane
        const bool live = c < C;
        uint8_t* act = sm.act[slot];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * kAccCols;
        uint8_t* column = act + ((c >> 3) * RB + 1) * 16 + (c & 7) * 2;    // row n: + 16 n
        uint32_t done_parity = 0;

        for (int round = 0; round < rounds; ++round) {
            const int tile = (round * gridDim.x + blockIdx.x) * kSlots + slot;
            if (tile >= n_tiles) break;
            const int row0 = tile * tile_rows - halo;           // global row of local row 0

            // fp32 rows -> bf16 operand buffer: 128 x 20 float4, 10 per thread
#pragma unroll
            for (int it = 0; it < 10; ++it) {
                const int i = gtid + 256 * it;
                const int r = i / (C / 4), c4 = i % (C / 4);
                const int gr = row0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr >= 0 && gr < total_rows) v = ld_stream4(x + (size_t)gr * C + 4 * c4);
                uint8_t* dst = act + ((c4 >> 1) * RB + r + 1) * 16 + (c4 & 1) * 8;
                *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            }
            if (gtid < NT) {                                     // validity bits of the 128 rows
                const int gr = row0 + gtid;
                const bool valid = gr >= 0 && gr < total_rows && __ldg(row_seq + gr) >= 0;
                const uint32_t bits = __ballot_sync(0xffffffffu, valid);
                if (lane == 0) sm.row_mask[slot][gtid >> 5] = bits;
            }
            // the 8 warps of the group publish together: named barrier 1 + slot
            asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(32 * kEpiWarps) : "memory");
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.act_ready[slot]);

            for (int layer = 0; layer < n_layers; ++layer) {
                const int a = acts.act[layer];
                const bool last = layer + 1 == n_layers;
                const float b = live ? sm.bias[layer][c] : 0.f;
                if (tid == 0) TRACE_T(5, round * n_layers + layer);
                mbar_wait(&sm.mma_done[slot], done_parity);
                if (tid == 0) TRACE_T(6, round * n_layers + layer);
                done_parity ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int n0 = half * (NT / 2); n0 < (half + 1) * (NT / 2); n0 += 32) {
#ifdef TCT_STSM
                    if (!last) {
                        // Experimental epilogue (profiles/r01z_ts_conv_probe.md, last section):
                        // accumulators in the mma-fragment layout, operand rows written by
                        // transposed 8 x 8 stores - 16-byte rows instead of 2-byte elements
                        const uint32_t mask = sm.row_mask[slot][n0 >> 5];
                        const int q = lane & 3, k = lane >> 3, jrow = lane & 7;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int ch0 = quad * 32 + h * 16;           // warp-uniform
                            if (ch0 >= C) break;
                            uint32_t f[16];
                            tmem_ld16x256(tmem_base + ((uint32_t)ch0 << 16) + slot * kAccCols + n0, f);
                            const float b_lo = sm.bias[layer][ch0 + (lane >> 2)];
                            const float b_hi = sm.bias[layer][ch0 + 8 + (lane >> 2)];
                            if (a == EMPH_ACT_RELU) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    f[j] = __float_as_uint(fmaxf(
                                        __uint_as_float(f[j]) + ((j & 2) ? b_hi : b_lo), 0.f));
                            } else if (a == EMPH_ACT_NONE) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    f[j] = __float_as_uint(__uint_as_float(f[j]) + ((j & 2) ? b_hi : b_lo));
                            } else {
                                uint32_t t[16];   // only this branch goes through local memory
#pragma unroll
                                for (int j = 0; j < 16; ++j) t[j] = f[j];
                                activate_pairs(t, b_lo, b_hi, a);
#pragma unroll
                                for (int j = 0; j < 16; ++j) f[j] = t[j];
                            }
                            uint32_t m[8];                                 // [block i][lo / hi channel group]
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                m[2 * i] = pack_bf16(__uint_as_float(f[4 * i]), __uint_as_float(f[4 * i + 1]));
                                m[2 * i + 1] = pack_bf16(__uint_as_float(f[4 * i + 2]), __uint_as_float(f[4 * i + 3]));
                            }
                            if (mask != 0xffffffffu) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t two = (mask >> (8 * i + 2 * q)) & 3u;
                                    const uint32_t keep = ((two & 1u) ? 0x0000ffffu : 0u) | ((two & 2u) ? 0xffff0000u : 0u);
                                    m[2 * i] &= keep;
                                    m[2 * i + 1] &= keep;
                                }
                            }
                            // matrix k of a store: channel group (k & 1), row block (k >> 1) of the pair
                            const int kg = (ch0 >> 3) + (k & 1);
#pragma unroll
                            for (int pair = 0; pair < 2; ++pair) {
                                const int n = n0 + 8 * (2 * pair + (k >> 1)) + jrow;
                                stmatrix_x4_trans(smem_u32(act + (kg * RB + 1 + n) * 16),
                                                  m[4 * pair], m[4 * pair + 1], m[4 * pair + 2], m[4 * pair + 3]);
                            }
                        }
                        continue;
                    }
#endif
                    uint32_t r[32];
                    tmem_ld32(taddr + n0, r);
                    if (tid == 0 && n0 == 0) TRACE_T(10, round * n_layers + layer);
                    if (tid == 0 && n0 == 32) TRACE_T(11, round * n_layers + layer);
                    const uint32_t mask = sm.row_mask[slot][n0 >> 5];
                    if (!live) continue;
                    // the activation is chosen OUTSIDE the unrolled loops: the
                    // generic one (erf / exp) stays rolled so the hot code is small;
                    // rows are zeroed only in the (rare) chunks that hold a gap row
                    if (a == EMPH_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]) + b, 0.f));
                    } else if (a == EMPH_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            r[j] = __float_as_uint(__uint_as_float(r[j]) + b);
                    } else {
                        uint32_t t[32];       // only this branch goes through local memory
#pragma unroll
                        for (int j = 0; j < 32; ++j) t[j] = r[j];
                        activate_rows(t, b, a);
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = t[j];
                    }
                    if (mask != 0xffffffffu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (!((mask >> j) & 1u)) r[j] = 0u;
                    }
                    if (!last) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            *reinterpret_cast<__nv_bfloat16*>(column + (n0 + j) * 16) =
                                __float2bfloat16_rn(__uint_as_float(r[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + j, g = row0 + n;
                            if (n < halo || n >= NT - halo || g < 0 || g >= total_rows) continue;
                            y[(size_t)g * C + c] = __uint_as_float(r[j]);
                        }
                    }
                }
                if (tid == 0) TRACE_T(7, round * n_layers + layer);
                if (!last) {
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.act_ready[slot]);
                } else {
                    // the next tile overwrites the operand buffer: every warp of the
                    // group must be past its last TMEM read and smem write
                    tc_fence_before();
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(32 * kEpiWarps) : "memory");
                }
            }
        }
    } else if (warp < kMmaWarp) {
        // ============================ weight group ============================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;                          // weight row = TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + kWeightCol0;
        uint32_t free_parity = 3u;                               // bit b; first use of a buffer never waits
        int g = 0;                                               // global layer counter
        for (int round = 0; round < rounds; ++round) {
            const int tile0 = (round * gridDim.x + blockIdx.x) * kSlots;
            if (tile0 >= n_tiles) break;
            for (int layer = 0; layer < n_layers; ++layer, ++g) {
                const int buffer = g & 1;
                mbar_wait(&sm.w_free[buffer], (free_parity >> buffer) & 1);
                if (tid == kWeightWarp0 * 32) TRACE_T(8, g);
                free_parity ^= 1u << buffer;
                tc_fence_after();
                if (quad * 32 < C) {       // warp-uniform: the store is .sync.aligned
                    // (rows 80..95 of the packed blob are zero)
                    const uint8_t* src = weights + (size_t)layer * W_LAYER_BYTES + (size_t)m * 32;
#pragma unroll
                    for (int chunk = 0; chunk < kChunks; ++chunk) {
                        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(src + chunk * 128 * 32));
                        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(src + chunk * 128 * 32 + 16));
                        asm volatile(
                            "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                            ::"r"(lane_base + buffer * kWeightCols + chunk * 8), "r"(lo.x), "r"(lo.y),
                              "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.w_ready[buffer]);
                if (tid == kWeightWarp0 * 32) TRACE_T(9, g);
            }
        }
    } else {
        // ============================= MMA issuer =============================
        uint32_t ready_parity = 0;                   // bit s = parity of slot s
        uint32_t weight_parity = 0;                  // bit b = parity of weight buffer b
        uint64_t d_act[kSlots];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) d_act[s] = make_desc(smem_u32(sm.act[s]), RB * 16, 128);
        int g = 0;
        for (int round = 0; round < rounds; ++round) {
            const int tile0 = (round * gridDim.x + blockIdx.x) * kSlots;
            if (tile0 >= n_tiles) break;
            const int active = min(kSlots, n_tiles - tile0);
            for (int layer = 0; layer < n_layers; ++layer, ++g) {
                const int buffer = g & 1;
                if (lane == 0) TRACE_T(0, g);
                mbar_wait(&sm.w_ready[buffer], (weight_parity >> buffer) & 1);
                if (lane == 0) TRACE_T(1, g);
                weight_parity ^= 1u << buffer;
                tc_fence_after();
                const uint32_t a_base = tmem_base + kWeightCol0 + buffer * kWeightCols;
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    if (s < active) {
                        mbar_wait(&sm.act_ready[s], (ready_parity >> s) & 1);
                        if (lane == 0) TRACE_T(2 + s, g);
                        ready_parity ^= 1u << s;
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem_base + s * kAccCols;
#pragma unroll
                            for (int chunk = 0; chunk < kChunks; ++chunk) {
                                const int tap = chunk / (C / 16), kk = chunk % (C / 16);
                                const uint64_t db =
                                    d_act[s] + (uint64_t)(((2 * kk) * RB * 16 + tap * 16) >> 4);
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\t"
                                    "setp.ne.b32 p, %4, 0;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                    ::"r"(d), "r"(a_base + chunk * 8), "l"(db), "r"(kInstrDesc),
                                      "r"((uint32_t)(chunk != 0)) : "memory");
                            }
                            umma_commit(&sm.mma_done[s]);
                        }
                        __syncwarp();
                    }
                }
                if (elect_one()) umma_commit(&sm.w_free[buffer]);   // weights consumed by both slots
                __syncwarp();
                if (lane == 0) TRACE_T(4, g);
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// fp32 weights [L][tap][in][out] -> bf16 [L][chunk = tap * 5 + kk][128 rows = out][16 = in % 16]
// (rows 80..127 zero: they are never copied to TMEM)
__global__ void pack_weights_tct_kernel(
    const float* __restrict__ w, int n_layers, __nv_bfloat16* __restrict__ out) {
    const int per_layer = W_LAYER_BYTES / 2;
    const int total = n_layers * per_layer;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int layer = i / per_layer, local = i % per_layer;
        const int e = local & 15, m = (local >> 4) & 127, chunk = local >> 11;
        const int tap = chunk / (C / 16), ci = (chunk % (C / 16)) * 16 + e;
        float v = 0.f;
        if (m < C) v = w[((size_t)(layer * KS + tap) * C + ci) * C + m];
        out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace tct

#ifdef EXP_TRACE
extern "C" int emph_conv_tct_trace_read(long long* host) {
    return (int)cudaMemcpyFromSymbol(host, tct::g_trace_t, sizeof(long long) * 12 * 256);
}
#endif

int conv_weights_tct_bytes(int n_layers) { return n_layers * tct::W_LAYER_BYTES; }

int pack_conv_weights_tct(
    const float* weights, int n_layers, void* packed, cudaStream_t stream) {
    tct::pack_weights_tct_kernel<<<64, 256, 0, stream>>>(
        weights, n_layers, reinterpret_cast<__nv_bfloat16*>(packed));
    EMPH_CHECK_LAUNCH("emph_pack_conv_weights_tc(transposed)");
    return EMPH_OK;
}

int conv_stack_bf16_tct(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const float* bias, const int32_t* acts_host, int32_t n_layers,
    float* y, cudaStream_t stream) {
    EMPH_REQUIRE(n_layers <= tct::kMaxLayers, "emph_conv_stack(tc transposed): too many layers");
    EMPH_REQUIRE(bias != nullptr, "emph_conv_stack(tc transposed): the fp32 bias is required");
    const int halo = n_layers;
    const int tile_rows = tct::NT - 2 * halo;
    EMPH_REQUIRE(tile_rows >= 32, "emph_conv_stack(tc transposed): %d layers leave no tile", n_layers);
    tct::Acts acts;
    for (int i = 0; i < tct::kMaxLayers; ++i) acts.act[i] = i < n_layers ? acts_host[i] : 0;
    const size_t smem = sizeof(tct::Smem) + 128;
    int s = check_cuda(
        cudaFuncSetAttribute(tct::conv_stack_tct_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
        "conv_tct smem attribute");
    if (s != EMPH_OK) return s;
    const int n_tiles = (total_rows + tile_rows - 1) / tile_rows;
    const int want = (n_tiles + tct::kSlots - 1) / tct::kSlots;
    const int grid = want < sm_count() ? want : sm_count();
    tct::conv_stack_tct_kernel<<<grid, tct::kThreads, smem, stream>>>(
        x, row_seq, total_rows, reinterpret_cast<const uint8_t*>(weights), bias, acts,
        n_layers, tile_rows, n_tiles, y);
    EMPH_CHECK_LAUNCH("emph_conv_stack(bf16 tc transposed)");
    return EMPH_OK;
}

}  // namespace emph
