// bf16 tensor-core fused Conv1d stack, "wide-N" formulation (sm_100a).
//
// Same job as conv_tc.cu (input_layer + frame_encoder, emphases/model/core.py:
// 92-94, emphases/model/layers/convolution.py:13-37) but organised around what
// tools/umma_bench.cu measured: every tcgen05.mma costs >= ~75 clk and SS-mode
// operand fetch sustains ~70 B/clk, so 16 MMAs of N = 80 per layer tile cannot
// pass ~40 % of the tensor peak.  Here the three taps become OUTPUT columns:
//     D[128 rows][240] = A[128][80] * [W_0 | W_1 | W_2]         (5 + 1 MMAs)
// and the +-1 row shift moves to the accumulator side, applied by the epilogue
// with warp shuffles:
//     out[m] = D_0[m-1] + D_1[m] + D_2[m+1]     (bias rides in the D_1 columns).
// A is read once per layer instead of three times.
//
// Warp roles (one persistent CTA per SM, 2 tile slots, 576 threads):
//   warps 8s .. 8s+7 : epilogue of slot s; warp e handles TMEM lane quadrant
//                      e % 4 and channels [40 (e / 4), +40) in 5 chunks of 8:
//                      3 x tcgen05.ld.x8, 16 shuffles, 16 adds, relu+bf16 pack,
//                      one 16-byte store into the next layer's A operand.  Row
//                      m-1 / m+1 across warp borders goes through a tiny smem
//                      mailbox + a 128-thread named barrier per chunk.
//   warp 16          : MMA issuer (warp-uniform loop, elect.sync)
//   warp 17          : weight producer (cp.async.bulk into a 3-stage ring)
#include <cuda_bf16.h>

#include "common.cuh"

namespace emph {

namespace tc240 {

constexpr int C = 80;
constexpr int KS = 3;
constexpr int KG = C / 8;
constexpr int N = KS * C;             // 240 output columns per MMA
constexpr int M = 128;
constexpr int kSlots = 2;
constexpr int kStages = 3;
constexpr int kMaxLayers = 16;
constexpr int ACT_BYTES = KG * M * 16;            // 20,480 per slot
constexpr int W_CONV_BYTES = KG * N * 16;         // 38,400
constexpr int W_BIAS_BYTES = 2 * N * 16;          // 7,680
constexpr int W_LAYER_BYTES = W_CONV_BYTES + W_BIAS_BYTES;   // 46,080
constexpr int kWarpsPerSlot = 8;
constexpr int kEpilogueThreads = kSlots * kWarpsPerSlot * 32;   // 512
constexpr int kThreads = kEpilogueThreads + 64;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;

struct __align__(128) Smem {
    uint8_t act[kSlots][ACT_BYTES];
    uint8_t w[kStages][W_LAYER_BYTES];
    uint8_t ones[2 * M * 16];
    // mailbox[slot][half][parity][quad][0: D_0 of lane 31 | 1: D_2 of lane 0][8]
    float mailbox[kSlots][2][2][4][2][8];
    uint64_t w_full[kStages];
    uint64_t w_empty[kStages];
    uint64_t act_ready[kSlots];
    uint64_t mma_done[kSlots];
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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .b64 state;\n\t"
        "mbarrier.arrive.expect_tx.shared::cta.b64 state, [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(
    uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// 32 lanes x 8 consecutive fp32 columns
// (results are only valid after tmem_ld_wait(): they land directly in the float
// registers so that no instruction touches them before the wait)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
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
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr uint32_t kInstrDesc =
    (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(kThreads, 1)
conv_stack_tc240_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ row_seq, int total_rows,
    const uint8_t* __restrict__ weights, Acts acts, int n_layers, int tile_rows, int n_tiles,
    float* __restrict__ y) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int halo = n_layers * ((KS - 1) / 2);
    const int rounds = (n_tiles + gridDim.x * kSlots - 1) / (gridDim.x * kSlots);

    // bias chunk A operand: k-group 0 = {1, 1, 0, ...} for every row, k-group 1 = 0
    for (int i = tid; i < 2 * M * 4; i += kThreads) {
        const int kg = i / (M * 4), word = i & 3;
        reinterpret_cast<uint32_t*>(sm.ones)[i] = (kg == 0 && word == 0) ? 0x3F803F80u : 0u;
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&sm.w_full[i], 1);
            mbar_init(&sm.w_empty[i], 1);
        }
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(&sm.act_ready[s], kWarpsPerSlot);
            mbar_init(&sm.mma_done[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == kSlots * kWarpsPerSlot) tmem_alloc(&sm.tmem_base, kTmemCols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < kSlots * kWarpsPerSlot) {
        // =========================== epilogue group ===========================
        const int slot = warp / kWarpsPerSlot;
        const int e = warp % kWarpsPerSlot;
        const int quad = e & 3;                     // TMEM lane quadrant
        const int half = e >> 2;                    // channel half: [40 half, +40)
        const int row = quad * 32 + lane;           // tile-local row == TMEM lane
        const int gtid = tid % (kWarpsPerSlot * 32);
        uint8_t* act = sm.act[slot];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * kAccStride;
        const int barrier_id = 1 + slot * 2 + half;     // 128 threads: 4 quads of one half
        uint32_t done_parity = 0;

        for (int round = 0; round < rounds; ++round) {
            const int tile = (round * gridDim.x + blockIdx.x) * kSlots + slot;
            if (tile >= n_tiles) break;
            const int row0 = tile * tile_rows - halo;
            const int g = row0 + row;
            const bool in_range = g >= 0 && g < total_rows;
            const bool valid = in_range && __ldg(row_seq + g) >= 0;

            // fp32 rows -> bf16 A operand (256 threads, 10 float4 each)
            {
                float4 v[10];
#pragma unroll
                for (int it = 0; it < 10; ++it) {
                    const int i = gtid + 256 * it;
                    const int gr = row0 + i / (C / 4);
                    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gr >= 0 && gr < total_rows)
                        v[it] = ld_stream4(x + (size_t)gr * C + 4 * (i % (C / 4)));
                }
#pragma unroll
                for (int it = 0; it < 10; ++it) {
                    const int i = gtid + 256 * it;
                    const int r = i / (C / 4), c4 = i % (C / 4);
                    *reinterpret_cast<uint2*>(act + ((c4 >> 1) * M + r) * 16 + (c4 & 1) * 8) =
                        make_uint2(pack_bf16(v[it].x, v[it].y), pack_bf16(v[it].z, v[it].w));
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.act_ready[slot]);

            {   // pull the next round's tile of this slot into L2
                const int next = tile + gridDim.x * kSlots;
                if (next < n_tiles) {
                    const long lo = (long)max(next * tile_rows - halo, 0) * C * 4;
                    const long hi = (long)min(next * tile_rows - halo + M, total_rows) * C * 4;
                    const char* base = reinterpret_cast<const char*>(x);
                    for (long off = lo + 128 * gtid; off < hi; off += 128 * 256)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
                }
            }
            const bool zero = !valid;

            for (int layer = 0; layer < n_layers; ++layer) {
                const int a = acts.act[layer];
                const bool last = layer + 1 == n_layers;
                const bool relu = a == EMPH_ACT_RELU;
                const bool simple = relu || a == EMPH_ACT_NONE;
                const bool store = last && in_range && row >= halo && row < M - halo;
                mbar_wait(&sm.mma_done[slot], done_parity);
                done_parity ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int chunk = 0; chunk < 5; ++chunk) {
                    const int c0 = half * 40 + chunk * 8;
                    float d0[8], d1[8], d2[8];
                    tmem_ld8(taddr + c0, d0);                 // tap 0: needs row m-1
                    tmem_ld8(taddr + 2 * C + c0, d2);         // tap 2: needs row m+1
                    tmem_ld8(taddr + C + c0, d1);             // tap 1 (+ bias)
                    tmem_ld_wait();
                    // rows across warp borders go through the mailbox
                    float* box = sm.mailbox[slot][half][chunk & 1][quad][0];
                    if (lane == 31) {
                        *reinterpret_cast<float4*>(box) = make_float4(d0[0], d0[1], d0[2], d0[3]);
                        *reinterpret_cast<float4*>(box + 4) = make_float4(d0[4], d0[5], d0[6], d0[7]);
                    }
                    if (lane == 0) {
                        *reinterpret_cast<float4*>(box + 8) = make_float4(d2[0], d2[1], d2[2], d2[3]);
                        *reinterpret_cast<float4*>(box + 12) = make_float4(d2[4], d2[5], d2[6], d2[7]);
                    }
                    named_barrier(barrier_id, 128);
                    float up[8], dn[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        up[j] = __shfl_up_sync(0xffffffffu, d0[j], 1);
                        dn[j] = __shfl_down_sync(0xffffffffu, d2[j], 1);
                    }
                    if (lane == 0) {
                        if (quad > 0) {
                            const float* src = sm.mailbox[slot][half][chunk & 1][quad - 1][0];
#pragma unroll
                            for (int j = 0; j < 8; ++j) up[j] = src[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) up[j] = 0.f;
                        }
                    }
                    if (lane == 31) {
                        if (quad < 3) {
                            const float* src = sm.mailbox[slot][half][chunk & 1][quad + 1][1];
#pragma unroll
                            for (int j = 0; j < 8; ++j) dn[j] = src[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) dn[j] = 0.f;
                        }
                    }
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = (up[j] + d1[j]) + dn[j];
                        if (!simple) v[j] = apply_activation(v[j], a);
                    }
                    if (!last) {
                        uint32_t p[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            p[j] = relu ? pack_bf16_relu(v[2 * j], v[2 * j + 1])
                                        : pack_bf16(v[2 * j], v[2 * j + 1]);
                        if (zero) p[0] = p[1] = p[2] = p[3] = 0u;
                        *reinterpret_cast<uint4*>(act + ((c0 >> 3) * M + row) * 16) =
                            make_uint4(p[0], p[1], p[2], p[3]);
                    } else if (store) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (relu) v[j] = fmaxf(v[j], 0.f);
                            if (zero) v[j] = 0.f;
                        }
                        float4* dst = reinterpret_cast<float4*>(y + (size_t)g * C + c0);
                        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                }
                if (!last) {
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.act_ready[slot]);
                }
            }
        }
    } else if (warp == kSlots * kWarpsPerSlot) {
        // ============================ MMA issuer ============================
        uint32_t ready_parity = 0;
        int stage = 0;
        uint32_t full_parity = 0;
        const uint64_t d_ones = make_desc(smem_u32(sm.ones), M * 16, 128);
        uint64_t d_act[kSlots];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) d_act[s] = make_desc(smem_u32(sm.act[s]), M * 16, 128);
        for (int round = 0; round < rounds; ++round) {
            const int tile0 = (round * gridDim.x + blockIdx.x) * kSlots;
            if (tile0 >= n_tiles) break;
            const int active = min(kSlots, n_tiles - tile0);
            for (int layer = 0; layer < n_layers; ++layer) {
                mbar_wait(&sm.w_full[stage], full_parity);
                const uint64_t d_w = make_desc(smem_u32(sm.w[stage]), N * 16, 128);
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    if (s < active) {
                        mbar_wait(&sm.act_ready[s], (ready_parity >> s) & 1);
                        ready_parity ^= 1u << s;
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem_base + s * kAccStride;
                            umma_bf16(d, d_ones, d_w + (W_CONV_BYTES >> 4), kInstrDesc, 0);
#pragma unroll
                            for (int kk = 0; kk < C / 16; ++kk)
                                umma_bf16(d, d_act[s] + (uint64_t)(((2 * kk) * M * 16) >> 4),
                                          d_w + (uint64_t)(((2 * kk) * N * 16) >> 4), kInstrDesc, 1);
                            umma_commit(&sm.mma_done[s]);
                        }
                        __syncwarp();
                    }
                }
                if (elect_one()) umma_commit(&sm.w_empty[stage]);
                __syncwarp();
                if (++stage == kStages) { stage = 0; full_parity ^= 1; }
            }
        }
    } else {
        // ========================== weight producer ==========================
        if (lane == 0) {
            int stage = 0;
            uint32_t empty_parity = 1;
            for (int round = 0; round < rounds; ++round) {
                const int tile0 = (round * gridDim.x + blockIdx.x) * kSlots;
                if (tile0 >= n_tiles) break;
                for (int layer = 0; layer < n_layers; ++layer) {
                    mbar_wait(&sm.w_empty[stage], empty_parity);
                    mbar_arrive_expect_tx(&sm.w_full[stage], W_LAYER_BYTES);
                    bulk_load(sm.w[stage], weights + (size_t)layer * W_LAYER_BYTES,
                              W_LAYER_BYTES, &sm.w_full[stage]);
                    if (++stage == kStages) { stage = 0; empty_parity ^= 1; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kSlots * kWarpsPerSlot) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// fp32 weights [L][tap][in][out] + bias [L][out] -> per layer
//   bf16 [kg][n = tap * 80 + out][8 in], then the bias chunk bf16 [2][n][8]:
//   k-group 0 of the tap-1 columns holds (bias_hi, bias_lo, 0, ...)
__global__ void pack_weights_kernel(
    const float* __restrict__ w, const float* __restrict__ bias, int n_layers,
    __nv_bfloat16* __restrict__ out) {
    const int per_layer = W_LAYER_BYTES / 2;
    const int conv = W_CONV_BYTES / 2;
    const int total = n_layers * per_layer;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int layer = i / per_layer, local = i % per_layer;
        float v = 0.f;
        if (local < conv) {
            const int e = local & 7, n = (local >> 3) % N, kg = (local >> 3) / N;
            const int tap = n / C, co = n % C, ci = kg * 8 + e;
            v = w[((size_t)(layer * KS + tap) * C + ci) * C + co];
        } else {
            const int rem = local - conv;
            const int e = rem & 7, n = (rem >> 3) % N, kg = (rem >> 3) / N;
            if (kg == 0 && e < 2 && n / C == 1) {
                const float b = bias[layer * C + n % C];
                const float hi = __bfloat162float(__float2bfloat16_rn(b));
                v = e == 0 ? hi : b - hi;
            }
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace tc240

int conv_stack_bf16_tc240(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const int32_t* acts_host, int32_t n_layers, float* y,
    cudaStream_t stream) {
    EMPH_REQUIRE(n_layers <= tc240::kMaxLayers, "emph_conv_stack(bf16 tc): too many layers");
    const int halo = n_layers * ((tc240::KS - 1) / 2);
    const int tile_rows = tc240::M - 2 * halo;
    EMPH_REQUIRE(tile_rows >= 32, "emph_conv_stack(bf16 tc): %d layers leave no tile", n_layers);
    tc240::Acts acts;
    for (int i = 0; i < tc240::kMaxLayers; ++i) acts.act[i] = i < n_layers ? acts_host[i] : 0;
    const size_t smem = sizeof(tc240::Smem) + 128;
    int s = check_cuda(
        cudaFuncSetAttribute(tc240::conv_stack_tc240_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
        "conv_tc240 smem attribute");
    if (s != EMPH_OK) return s;
    const int n_tiles = (total_rows + tile_rows - 1) / tile_rows;
    const int want = (n_tiles + tc240::kSlots - 1) / tc240::kSlots;
    const int grid = want < sm_count() ? want : sm_count();
    tc240::conv_stack_tc240_kernel<<<grid, tc240::kThreads, smem, stream>>>(
        x, row_seq, total_rows, reinterpret_cast<const uint8_t*>(weights), acts,
        n_layers, tile_rows, n_tiles, y);
    EMPH_CHECK_LAUNCH("emph_conv_stack(bf16 tc, N=240)");
    return EMPH_OK;
}

int conv_weights_tc240_bytes(int n_layers) { return n_layers * tc240::W_LAYER_BYTES; }

int pack_conv_weights_tc240(
    const float* weights, const float* bias, int n_layers, void* packed, cudaStream_t stream) {
    tc240::pack_weights_kernel<<<64, 256, 0, stream>>>(
        weights, bias, n_layers, reinterpret_cast<__nv_bfloat16*>(packed));
    EMPH_CHECK_LAUNCH("emph_pack_conv_weights_tc(N=240)");
    return EMPH_OK;
}

}  // namespace emph
