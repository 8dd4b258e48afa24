// Frame -> word segment pooling, sm_100a.  Replaces emphases.downsample
// (emphases/core.py:426-469), whose per-word Python loop costs >= 2 host syncs
// and >= 3 launches per word on a GPU, by one launch driven by int32 offset
// arrays: a warp owns a word row, lanes own float4 channel groups, the frame
// loop is unrolled for memory-level parallelism.  HBM-bound: 4*C bytes per
// frame in, 4*C bytes per word out.
#include <math_constants.h>

#include "common.cuh"

namespace emph {

constexpr int kPoolWarps = 8;

template <int METHOD>
__device__ __forceinline__ float4 pool_combine(float4 a, float4 b) {
    if (METHOD == EMPH_POOL_MAX)
        return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__device__ __forceinline__ float4 ldg_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int METHOD>
__global__ void __launch_bounds__(kPoolWarps * 32)
pool_words_kernel(
    const float* __restrict__ x, int channels,
    const int32_t* __restrict__ row_start, const int32_t* __restrict__ n_rows,
    const int32_t* __restrict__ word_seq, const int32_t* __restrict__ word_lo,
    const int32_t* __restrict__ word_hi, int total_word_rows, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int groups = channels >> 2;                // float4 groups per row
    const int warps_total = gridDim.x * kPoolWarps;
    for (int w = blockIdx.x * kPoolWarps + (threadIdx.x >> 5); w < total_word_rows;
         w += warps_total) {
        const int u = __ldg(word_seq + w);
        float* dst = y + (size_t)w * channels;
        const int lo = __ldg(word_lo + w), hi = __ldg(word_hi + w);
        const bool padded_slot = (lo == -1 && hi == -1);
        const bool masked_slot = (lo == -2 && hi == -2);
        if (u < 0 || masked_slot || (padded_slot && METHOD != EMPH_POOL_CENTER)) {
            for (int g = lane; g < groups; g += 32)
                *reinterpret_cast<float4*>(dst + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const int n = __ldg(n_rows + u);
        const float* base = x + (size_t)__ldg(row_start + u) * channels;
        if (METHOD == EMPH_POOL_CENTER) {
            // padded slots carry bounds (0, 0) in the reference -> frame 0
            const int idx = padded_slot ? 0 : (lo + hi) >> 1;   // floor for idx >= 0
            for (int g = lane; g < groups; g += 32) {
                float4 v = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
                if (idx >= 0 && idx < n) v = ldg_stream(base + (size_t)idx * channels + 4 * g);
                *reinterpret_cast<float4*>(dst + 4 * g) = v;
            }
            continue;
        }
        // torch slice semantics: both ends clipped to [0, n]
        const int s = min(max(lo, 0), n), e = min(max(hi, 0), n);
        const int count = max(e - s, 0);
        for (int g = lane; g < groups; g += 32) {
            const float init = METHOD == EMPH_POOL_MAX ? -CUDART_INF_F : 0.f;
            float4 acc0 = make_float4(init, init, init, init), acc1 = acc0, acc2 = acc0, acc3 = acc0;
            const float* p = base + (size_t)s * channels + 4 * g;
            int f = 0;
            for (; f + 4 <= count; f += 4) {
                float4 v0 = ldg_stream(p + (size_t)(f + 0) * channels);
                float4 v1 = ldg_stream(p + (size_t)(f + 1) * channels);
                float4 v2 = ldg_stream(p + (size_t)(f + 2) * channels);
                float4 v3 = ldg_stream(p + (size_t)(f + 3) * channels);
                acc0 = pool_combine<METHOD>(acc0, v0);
                acc1 = pool_combine<METHOD>(acc1, v1);
                acc2 = pool_combine<METHOD>(acc2, v2);
                acc3 = pool_combine<METHOD>(acc3, v3);
            }
            for (; f < count; ++f)
                acc0 = pool_combine<METHOD>(acc0, ldg_stream(p + (size_t)f * channels));
            float4 acc = pool_combine<METHOD>(pool_combine<METHOD>(acc0, acc1),
                                              pool_combine<METHOD>(acc2, acc3));
            if (METHOD == EMPH_POOL_AVERAGE) {
                const float d = (float)count;         // 0 -> NaN like torch.mean([])
                acc = make_float4(acc.x / d, acc.y / d, acc.z / d, acc.w / d);
            }
            *reinterpret_cast<float4*>(dst + 4 * g) = acc;
        }
    }
}

}  // namespace emph

extern "C" int emph_pool_words(
    const float* x, int32_t channels,
    const int32_t* row_start, const int32_t* n_rows,
    const int32_t* word_seq, const int32_t* word_lo, const int32_t* word_hi,
    int32_t total_word_rows, int32_t method, float* y, void* stream) {
    EMPH_REQUIRE(channels > 0 && channels % 4 == 0, "emph_pool_words: channels %d not a multiple of 4", channels);
    EMPH_REQUIRE(total_word_rows >= 0, "emph_pool_words: negative size");
    if (total_word_rows == 0) return EMPH_OK;
    long want = ((long)total_word_rows + emph::kPoolWarps - 1) / emph::kPoolWarps;
    long cap = (long)emph::sm_count() * 8;
    int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = (cudaStream_t)stream;
#define EMPH_POOL_LAUNCH(M)                                                       \
    emph::pool_words_kernel<M><<<grid, emph::kPoolWarps * 32, 0, st>>>(           \
        x, channels, row_start, n_rows, word_seq, word_lo, word_hi, total_word_rows, y)
    switch (method) {
        case EMPH_POOL_AVERAGE: EMPH_POOL_LAUNCH(EMPH_POOL_AVERAGE); break;
        case EMPH_POOL_MAX: EMPH_POOL_LAUNCH(EMPH_POOL_MAX); break;
        case EMPH_POOL_SUM: EMPH_POOL_LAUNCH(EMPH_POOL_SUM); break;
        case EMPH_POOL_CENTER: EMPH_POOL_LAUNCH(EMPH_POOL_CENTER); break;
        default:
            emph::set_error("emph_pool_words: unknown method %d", method);
            return EMPH_EINVAL;
    }
#undef EMPH_POOL_LAUNCH
    EMPH_CHECK_LAUNCH("emph_pool_words");
    return EMPH_OK;
}
