// Per-file sums for the evaluation caller (emphases/evaluate/core.py:27-110,
// emphases/evaluate/metrics.py:13-111): one warp per file over its packed
// word rows, accumulated in fp64 in a fixed order (deterministic).
//
//   pass 0:  sums[u] = { sum p, sum t, 0, 0, 0 }
//   pass 1:  sums[u] = { sum (p - mean_p)^2, sum (t - mean_t)^2,
//                        sum (p - mean_p)(t - mean_t), sum bce, sum (p - t)^2 }
// with p = postprocess(logit): sigmoid (LOSS 'bce') or clamp to [0, 1] ('mse'),
// and bce = BCE-with-logits ('bce', metrics.py:62-68) or the clamped-probability
// form -(t log(x + 1e-6) + (1 - t) log(1 - x + 1e-6)) ('mse', metrics.py:70-75).
#include "common.cuh"

namespace emph {

constexpr int kEvalWarps = 8;

__global__ void __launch_bounds__(kEvalWarps * 32)
word_metric_sums_kernel(
    const float* __restrict__ logits, const float* __restrict__ targets,
    const int32_t* __restrict__ word_row_start, const int32_t* __restrict__ n_words,
    int n_seq, int loss_mode, int pass, double mean_p, double mean_t,
    double* __restrict__ sums) {
    const int lane = threadIdx.x & 31;
    const int u = blockIdx.x * kEvalWarps + (threadIdx.x >> 5);
    if (u >= n_seq) return;
    const int first = word_row_start[u], count = n_words[u];
    double acc[5] = {0., 0., 0., 0., 0.};
    for (int i = lane; i < count; i += 32) {
        const float x = logits[first + i];
        const float t = targets[first + i];
        float p;
        if (loss_mode == 0) p = 1.f / (1.f + expf(-x));
        else p = fminf(fmaxf(x, 0.f), 1.f);
        if (pass == 0) {
            acc[0] += (double)p;
            acc[1] += (double)t;
        } else {
            const double dp = (double)p - mean_p, dt = (double)t - mean_t;
            double bce;
            if (loss_mode == 0) {
                bce = fmax((double)x, 0.) - (double)x * (double)t +
                      log1p(exp(-fabs((double)x)));
            } else {
                bce = -((double)t * log((double)p + 1e-6) +
                        (1. - (double)t) * log(1. - (double)p + 1e-6));
            }
            const double d = (double)p - (double)t;
            acc[0] += dp * dp;
            acc[1] += dt * dt;
            acc[2] += dp * dt;
            acc[3] += bce;
            acc[4] += d * d;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int offset = 16; offset > 0; offset >>= 1)
            acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], offset);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) sums[(size_t)u * 5 + k] = acc[k];
    }
}

}  // namespace emph

extern "C" int emph_word_metric_sums(
    const float* logits, const float* targets,
    const int32_t* word_row_start, const int32_t* n_words, int32_t n_seq,
    int32_t loss_mode, int32_t pass, double mean_p, double mean_t,
    double* sums, void* stream) {
    using namespace emph;
    EMPH_REQUIRE(n_seq >= 0, "emph_word_metric_sums: negative size");
    EMPH_REQUIRE(loss_mode == 0 || loss_mode == 1,
                 "emph_word_metric_sums: loss_mode %d is not 0 (bce) or 1 (mse)", loss_mode);
    EMPH_REQUIRE(pass == 0 || pass == 1, "emph_word_metric_sums: pass %d is not 0 or 1", pass);
    if (n_seq == 0) return EMPH_OK;
    const int grid = (n_seq + kEvalWarps - 1) / kEvalWarps;
    word_metric_sums_kernel<<<grid, kEvalWarps * 32, 0, (cudaStream_t)stream>>>(
        logits, targets, word_row_start, n_words, n_seq, loss_mode, pass,
        mean_p, mean_t, sums);
    EMPH_CHECK_LAUNCH("emph_word_metric_sums");
    return EMPH_OK;
}
