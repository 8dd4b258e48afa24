// Pieces shared by the tensor-core attention (attention_tc.cu) and the fused
// Transformer-layer kernels that write its K / V records (transformer_tc.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"

namespace emph {
namespace attn_tc {

constexpr int kThreads = 256;
constexpr int kQueries = 128;      // per CTA = the host's query block
constexpr int kKeys = 64;          // per shared-memory tile
constexpr int kStages = 3;         // key tiles in flight
enum { kPlainFp16 = 0, kSplitBf16 = 1 };

template <int D, int MODE>
struct Layout {
    static constexpr int DP = (D + 15) / 16 * 16;          // K dims padded to whole k-steps
    static constexpr int NP = MODE == kSplitBf16 ? 2 : 1;  // operand parts
    static constexpr int kOffV = NP * DP;                   // in 16-bit elements
    static constexpr int kContent = NP * (DP + D) * 2;      // bytes
    static constexpr int kRecord = kContent % 32 == 16 ? kContent : kContent + 16;
    static constexpr int kTileBytes = kKeys * kRecord;
    static constexpr int kSmemBytes = kStages * kTileBytes + 16 * kStages;  // + mbarriers
    static_assert(D % 8 == 0, "head dim must be a multiple of 8");
    static_assert(kRecord % 32 == 16, "ldmatrix rows must land on distinct banks");
};

__device__ __forceinline__ float exp2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// {lo -> bits 0..15, hi -> bits 16..31}
template <int MODE>
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi) {
    uint32_t r;
    if (MODE == kPlainFp16)
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// what the rounding to bf16 dropped, as a second bf16 pair
__device__ __forceinline__ uint32_t pack_residual(float lo, float hi, uint32_t rounded) {
    return pack_pair<kSplitBf16>(
        lo - __uint_as_float(rounded << 16), hi - __uint_as_float(rounded & 0xffff0000u));
}

template <int MODE>
__device__ __forceinline__ void mma_16816(
    float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (MODE == kPlainFp16)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 "
            "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
            "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t address) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(address));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t address) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(address));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t address) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(address));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t address) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(address));
}

}  // namespace attn_tc
}  // namespace emph
