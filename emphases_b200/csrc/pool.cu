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

// ---- pooling fused into the tensor-core conv stack (conv_tc.cu) ----
int conv_stack_tc_pool(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const int32_t* acts_host, int32_t n_layers, int32_t channels,
    int32_t kernel_size, int32_t precision, const int32_t* row_word, int32_t pool_mode,
    long long* fixed, float* out, float* y, cudaStream_t stream);

// row_word[g] = the word row that frame row g belongs to (pre-set to -1): a warp
// per word row marks [lo, hi) clipped to the sequence (or only the centre row).
template <bool CENTER>
__global__ void __launch_bounds__(kPoolWarps * 32)
row_words_kernel(
    const int32_t* __restrict__ row_start, const int32_t* __restrict__ n_rows,
    const int32_t* __restrict__ word_seq, const int32_t* __restrict__ word_lo,
    const int32_t* __restrict__ word_hi, int total_word_rows,
    int32_t* __restrict__ row_word, int32_t* __restrict__ word_count) {
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * kPoolWarps;
    for (int w = blockIdx.x * kPoolWarps + (threadIdx.x >> 5); w < total_word_rows;
         w += warps_total) {
        const int u = __ldg(word_seq + w);
        int s = 0, e = 0;
        if (u >= 0) {
            const int n = __ldg(n_rows + u);
            const int lo = __ldg(word_lo + w), hi = __ldg(word_hi + w);
            if (CENTER) {
                const int idx = (lo + hi) >> 1;
                if (idx >= 0 && idx < n) { s = idx; e = idx + 1; }
            } else {
                s = min(max(lo, 0), n);
                e = min(max(hi, 0), n);
            }
            const int base = __ldg(row_start + u);
            for (int f = s + lane; f < e; f += 32) row_word[base + f] = w;
        }
        if (lane == 0) word_count[w] = u < 0 ? -1 : max(e - s, 0);
    }
}

// fixed-point sums -> fp32 (sum) or fp32 mean (sum / count; 0 rows: NaN like
// torch.mean of an empty slice)
__global__ void pool_fixed_finish_kernel(
    const long long* __restrict__ fixed, const int32_t* __restrict__ word_count,
    long long total, int channels, int average, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const double sum = (double)fixed[i] * (1.0 / 268435456.0);     // 2^-28 units
        const int count = word_count[i / channels];               // -1: separator word row
        out[i] = count < 0 ? 0.f : average ? (float)(sum / (double)count) : (float)sum;
    }
}

}  // namespace emph

extern "C" int emph_conv_stack_pool(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const int32_t* acts_host, int32_t n_layers, int32_t channels,
    int32_t kernel_size, int32_t precision,
    const int32_t* row_start, const int32_t* n_rows,
    const int32_t* word_seq, const int32_t* word_lo, const int32_t* word_hi,
    int32_t total_word_rows, int32_t method,
    int32_t* row_word, int32_t* word_count, long long* fixed,
    float* pooled, float* y, void* stream) {
    EMPH_REQUIRE(total_rows >= 0 && total_word_rows >= 0, "emph_conv_stack_pool: negative size");
    EMPH_REQUIRE(n_layers > 0, "emph_conv_stack_pool: no layers");
    EMPH_REQUIRE(method == EMPH_POOL_SUM || method == EMPH_POOL_AVERAGE ||
                     method == EMPH_POOL_MAX || method == EMPH_POOL_CENTER,
                 "emph_conv_stack_pool: unknown method %d", method);
    EMPH_REQUIRE(x != y, "emph_conv_stack_pool: x and y must not alias (halo rows)");
    if (total_rows == 0 || total_word_rows == 0) return EMPH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool sums = method == EMPH_POOL_SUM || method == EMPH_POOL_AVERAGE;
    const size_t elements = (size_t)total_word_rows * channels;
    int s = emph::check_cuda(
        cudaMemsetAsync(row_word, 0xFF, sizeof(int32_t) * (size_t)total_rows, st),
        "emph_conv_stack_pool: memset");
    if (s != EMPH_OK) return s;
    s = emph::check_cuda(
        sums ? cudaMemsetAsync(fixed, 0, sizeof(long long) * elements, st)
             : cudaMemsetAsync(pooled, 0, sizeof(float) * elements, st),
        "emph_conv_stack_pool: memset");
    if (s != EMPH_OK) return s;
    long want = ((long)total_word_rows + emph::kPoolWarps - 1) / emph::kPoolWarps;
    long cap = (long)emph::sm_count() * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (method == EMPH_POOL_CENTER)
        emph::row_words_kernel<true><<<grid, emph::kPoolWarps * 32, 0, st>>>(
            row_start, n_rows, word_seq, word_lo, word_hi, total_word_rows, row_word, word_count);
    else
        emph::row_words_kernel<false><<<grid, emph::kPoolWarps * 32, 0, st>>>(
            row_start, n_rows, word_seq, word_lo, word_hi, total_word_rows, row_word, word_count);
    EMPH_CHECK_LAUNCH("emph_conv_stack_pool(row words)");
    // tc::kPoolSum = 1, kPoolMax = 2, kPoolCenter = 3
    const int mode = sums ? 1 : method == EMPH_POOL_MAX ? 2 : 3;
    s = emph::conv_stack_tc_pool(
        x, row_seq, total_rows, weights, acts_host, n_layers, channels, kernel_size, precision,
        row_word, mode, fixed, pooled, y, st);
    if (s != EMPH_OK) return s;
    if (sums) {
        const long long total = (long long)elements;
        const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        emph::pool_fixed_finish_kernel<<<blocks, 256, 0, st>>>(
            fixed, word_count, total, channels, method == EMPH_POOL_AVERAGE, pooled);
        EMPH_CHECK_LAUNCH("emph_conv_stack_pool(finish)");
    }
    return EMPH_OK;
}

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
