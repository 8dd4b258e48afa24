// Output projection Conv1d(C -> 1, k, 'same') over packed rows + postprocess,
// sm_100a.  Replaces output_layer (emphases/model/core.py:33-37,138) and
// emphases.postprocess (emphases/core.py:335-342).  One warp per row; the
// k neighbouring rows are read straight from L2 (tiny: 2.5 words per second
// of audio).
#include "common.cuh"

namespace emph {

constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32)
output_head_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ row_seq, int total_rows,
    int channels, int kernel_size, const float* __restrict__ weight, float bias,
    int mode, float* __restrict__ logits, float* __restrict__ scores) {
    const int lane = threadIdx.x & 31;
    const int half = (kernel_size - 1) / 2;
    for (int r = blockIdx.x * kHeadWarps + (threadIdx.x >> 5); r < total_rows;
         r += gridDim.x * kHeadWarps) {
        if (__ldg(row_seq + r) < 0) {
            if (lane == 0) {
                if (logits) logits[r] = 0.f;
                if (scores) scores[r] = 0.f;
            }
            continue;
        }
        float acc = 0.f;
        for (int tap = 0; tap < kernel_size; ++tap) {
            const int g = r + tap - half;
            if (g < 0 || g >= total_rows) continue;
            const float* src = x + (size_t)g * channels;
            const float* w = weight + tap * channels;
            for (int c = lane; c < channels; c += 32) acc = fmaf(__ldg(w + c), src[c], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float logit = acc + bias;
            if (logits) logits[r] = logit;
            if (scores) {
                float s = logit;
                if (mode == EMPH_HEAD_SIGMOID) s = 1.f / (1.f + expf(-logit));
                else if (mode == EMPH_HEAD_CLAMP) s = fminf(fmaxf(logit, 0.f), 1.f);
                scores[r] = s;
            }
        }
    }
}

}  // namespace emph

extern "C" int emph_output_head(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    int32_t channels, int32_t kernel_size, const float* weight, float bias_host,
    int32_t mode, float* logits, float* scores, void* stream) {
    EMPH_REQUIRE(total_rows >= 0 && channels > 0, "emph_output_head: bad size");
    EMPH_REQUIRE(kernel_size >= 1 && (kernel_size & 1), "emph_output_head: kernel_size must be odd");
    if (total_rows == 0) return EMPH_OK;
    long want = ((long)total_rows + emph::kHeadWarps - 1) / emph::kHeadWarps;
    long cap = (long)emph::sm_count() * 8;
    int grid = (int)(want < cap ? want : cap);
    emph::output_head_kernel<<<grid, emph::kHeadWarps * 32, 0, (cudaStream_t)stream>>>(
        x, row_seq, total_rows, channels, kernel_size, weight, bias_host, mode, logits, scores);
    EMPH_CHECK_LAUNCH("emph_output_head");
    return EMPH_OK;
}
