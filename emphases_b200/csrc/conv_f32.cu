// fp32 (CUDA-core FFMA) fused Conv1d stack over the packed row axis, sm_100a.
//
// Replaces input_layer + frame_encoder (emphases/model/core.py:92-94,
// emphases/model/layers/convolution.py:13-37) and word_decoder
// (model/core.py:105-107): n_layers x [Conv1d(C -> C, k, 'same') + activation]
// with all layers fused in one kernel.  A CTA owns a tile of 128 consecutive
// rows including a halo of n_layers * (k-1)/2 rows each side; activations stay
// in shared memory across layers (updated in place through registers), weights
// of layer l+1 are prefetched with cp.async while layer l computes.  Separator
// rows (row_seq < 0) are re-zeroed after every layer, which is exactly the
// per-utterance zero 'same' padding of the reference.
//
// This is the exact-parity mode (max-abs 1e-5 on scores); the tensor-core
// mode lives in conv_tc.cu.
#include "common.cuh"

namespace emph {

constexpr int kConvMaxLayers = 16;

struct ConvActs {
    int act[kConvMaxLayers];
};

// DB: double-buffer the layer weights (prefetch layer l+1 during layer l);
// kernel sizes 5 and 7 only fit shared memory single-buffered
template <int C, int KS, bool DB = true>
struct ConvF32 {
    static constexpr int R = 128;                 // rows per CTA per layer
    static constexpr int HALF = (KS - 1) / 2;
    static constexpr int LD = C + 4;              // smem row stride (floats)
    static constexpr int THREADS = 256;
    static constexpr int CT = C / 8;              // output channels per thread
    static constexpr int WFLOATS = KS * C * C;    // one layer's weights
    static constexpr int ACT_FLOATS = (R + 2 * HALF) * LD;
    static constexpr int WBUFS = DB ? 2 : 1;
    static constexpr size_t SMEM =
        (size_t)(ACT_FLOATS + WBUFS * WFLOATS) * sizeof(float) + R * sizeof(int);
    static_assert(C % 8 == 0 && CT % 2 == 0, "channels must be a multiple of 16");
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;\n" ::);
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int C, int KS, bool DB>
__global__ void __launch_bounds__(ConvF32<C, KS, DB>::THREADS, 1)
conv_stack_f32_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ row_seq, int total_rows,
    const float* __restrict__ weights, const float* __restrict__ bias,
    ConvActs acts, int n_layers, int tile_rows, float* __restrict__ y) {
    using Cfg = ConvF32<C, KS, DB>;
    constexpr int R = Cfg::R, HALF = Cfg::HALF, LD = Cfg::LD, CT = Cfg::CT;
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                              // [(R + 2 HALF)][LD]
    float* wbuf = smem + Cfg::ACT_FLOATS;           // [WBUFS][KS*C][C]
    int* valid = reinterpret_cast<int*>(wbuf + Cfg::WBUFS * Cfg::WFLOATS);   // [R]

    const int tid = threadIdx.x;
    const int halo = n_layers * HALF;
    const int row0 = blockIdx.x * tile_rows - halo;   // global row of local row 0

    // weights of layer 0
    for (int i = tid; i < Cfg::WFLOATS / 4; i += Cfg::THREADS)
        cp_async16(wbuf + 4 * i, weights + 4 * i);
    cp_async_commit();

    // zero pad rows, load the input tile
    for (int i = tid; i < HALF * LD; i += Cfg::THREADS) {
        act[i] = 0.f;
        act[(R + HALF) * LD + i] = 0.f;
    }
    for (int i = tid; i < R * (C / 4); i += Cfg::THREADS) {
        int r = i / (C / 4), c4 = i % (C / 4);
        int g = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= 0 && g < total_rows)
            v = *reinterpret_cast<const float4*>(x + (size_t)g * C + 4 * c4);
        *reinterpret_cast<float4*>(act + (r + HALF) * LD + 4 * c4) = v;
    }
    for (int r = tid; r < R; r += Cfg::THREADS) {
        int g = row0 + r;
        valid[r] = (g >= 0 && g < total_rows) ? (__ldg(row_seq + g) >= 0) : 0;
    }

    const int tx = tid & 7;        // channel group: co = tx * CT .. + CT
    const int ty = tid >> 3;       // rows ty + 32 i, i = 0..3

    for (int layer = 0; layer < n_layers; ++layer) {
        float* w = wbuf + (DB ? (layer & 1) : 0) * Cfg::WFLOATS;
        if (DB && layer + 1 < n_layers) {
            float* wn = wbuf + ((layer + 1) & 1) * Cfg::WFLOATS;
            const float* src = weights + (size_t)(layer + 1) * Cfg::WFLOATS;
            for (int i = tid; i < Cfg::WFLOATS / 4; i += Cfg::THREADS)
                cp_async16(wn + 4 * i, src + 4 * i);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        float acc[4][CT];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;

#pragma unroll
        for (int tap = 0; tap < KS; ++tap) {
            const float* a0 = act + (ty + tap) * LD;     // (+HALF pad, -HALF shift)
            const float* wt = w + tap * C * C + tx * CT;
#pragma unroll 2
            for (int c4 = 0; c4 < C / 4; ++c4) {
                float4 a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[i] = *reinterpret_cast<const float4*>(a0 + 32 * i * LD + 4 * c4);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float2 wv[CT / 2];
#pragma unroll
                    for (int j = 0; j < CT / 2; ++j)
                        wv[j] = *reinterpret_cast<const float2*>(wt + (4 * c4 + cc) * C + 2 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float av = cc == 0 ? a[i].x : cc == 1 ? a[i].y : cc == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int j = 0; j < CT / 2; ++j) {
                            acc[i][2 * j] = fmaf(av, wv[j].x, acc[i][2 * j]);
                            acc[i][2 * j + 1] = fmaf(av, wv[j].y, acc[i][2 * j + 1]);
                        }
                    }
                }
            }
        }
        __syncthreads();   // every read of this layer's input (and weights) is done
        if (!DB && layer + 1 < n_layers) {
            // single weight buffer: the next layer's weights load while the
            // epilogue below runs
            const float* src = weights + (size_t)(layer + 1) * Cfg::WFLOATS;
            for (int i = tid; i < Cfg::WFLOATS / 4; i += Cfg::THREADS)
                cp_async16(wbuf + 4 * i, src + 4 * i);
            cp_async_commit();
        }

        const int a = acts.act[layer];
        const float* b = bias + layer * C + tx * CT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r = ty + 32 * i;
            float keep = valid[r] ? 1.f : 0.f;
            float* dst = act + (r + HALF) * LD + tx * CT;
#pragma unroll
            for (int j = 0; j < CT; j += 2) {
                float v0 = apply_activation(acc[i][j] + __ldg(b + j), a);
                float v1 = apply_activation(acc[i][j + 1] + __ldg(b + j + 1), a);
                // separator rows are forced to exact zero (not NaN * 0)
                v0 = keep != 0.f ? v0 : 0.f;
                v1 = keep != 0.f ? v1 : 0.f;
                *reinterpret_cast<float2*>(dst + j) = make_float2(v0, v1);
            }
        }
        __syncthreads();
    }

    // coalesced copy of the exact inner rows
    for (int i = tid; i < tile_rows * (C / 4); i += Cfg::THREADS) {
        int r = halo + i / (C / 4), c4 = i % (C / 4);
        int g = row0 + r;
        if (g < total_rows)
            *reinterpret_cast<float4*>(y + (size_t)g * C + 4 * c4) =
                *reinterpret_cast<const float4*>(act + (r + HALF) * LD + 4 * c4);
    }
}

template <int C, int KS, bool DB = true>
int launch_conv_f32(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const float* bias, const int32_t* acts_host,
    int32_t n_layers, float* y, cudaStream_t stream) {
    using Cfg = ConvF32<C, KS, DB>;
    const int halo = n_layers * Cfg::HALF;
    const int tile_rows = Cfg::R - 2 * halo;
    EMPH_REQUIRE(tile_rows >= 32, "emph_conv_stack: %d layers of kernel %d leave no tile", n_layers, KS);
    ConvActs acts;
    for (int i = 0; i < kConvMaxLayers; ++i) acts.act[i] = i < n_layers ? acts_host[i] : 0;
    int s = check_cuda(
        cudaFuncSetAttribute(
            conv_stack_f32_kernel<C, KS, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)Cfg::SMEM),
        "conv_f32 smem attribute");
    if (s != EMPH_OK) return s;
    int grid = (total_rows + tile_rows - 1) / tile_rows;
    conv_stack_f32_kernel<C, KS, DB><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(
        x, row_seq, total_rows, weights, bias, acts, n_layers, tile_rows, y);
    EMPH_CHECK_LAUNCH("emph_conv_stack(fp32)");
    return EMPH_OK;
}

// ---------------------------------------------------------------------------
// Wide models (CHANNELS = 128 of the reference's hyper-parameter sweep): a
// layer's weights (k x 128 x 128 fp32 = 196 KB at k = 3) no longer fit shared
// memory, so they stream through it one TAP at a time (64 KB), double
// buffered with cp.async: tap q + 1 loads while tap q is consumed.  Same
// tiling, register blocking and separator handling as the kernel above.
template <int C, int KS>
struct ConvF32Tap {
    static constexpr int R = 128;
    static constexpr int HALF = (KS - 1) / 2;
    static constexpr int LD = C + 4;
    static constexpr int THREADS = 256;
    static constexpr int CT = C / 8;
    static constexpr int TAP_FLOATS = C * C;
    static constexpr int ACT_FLOATS = (R + 2 * HALF) * LD;
    static constexpr size_t SMEM =
        (size_t)(ACT_FLOATS + 2 * TAP_FLOATS) * sizeof(float) + R * sizeof(int);
    static_assert(C % 8 == 0 && CT % 2 == 0, "channels must be a multiple of 16");
};

template <int C, int KS>
__global__ void __launch_bounds__(ConvF32Tap<C, KS>::THREADS, 1)
conv_stack_f32_tap_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ row_seq, int total_rows,
    const float* __restrict__ weights, const float* __restrict__ bias,
    ConvActs acts, int n_layers, int tile_rows, float* __restrict__ y) {
    using Cfg = ConvF32Tap<C, KS>;
    constexpr int R = Cfg::R, HALF = Cfg::HALF, LD = Cfg::LD, CT = Cfg::CT;
    extern __shared__ __align__(16) float smem[];
    float* act = smem;                              // [(R + 2 HALF)][LD]
    float* wbuf = smem + Cfg::ACT_FLOATS;           // [2][C][C]
    int* valid = reinterpret_cast<int*>(wbuf + 2 * Cfg::TAP_FLOATS);   // [R]

    const int tid = threadIdx.x;
    const int halo = n_layers * HALF;
    const int row0 = blockIdx.x * tile_rows - halo;
    const int n_taps = n_layers * KS;               // taps in stream order [layer][tap]

    auto fetch_tap = [&](int q) {
        float* dst = wbuf + (q & 1) * Cfg::TAP_FLOATS;
        const float* src = weights + (size_t)q * Cfg::TAP_FLOATS;
        for (int i = tid; i < Cfg::TAP_FLOATS / 4; i += Cfg::THREADS)
            cp_async16(dst + 4 * i, src + 4 * i);
        cp_async_commit();
    };
    fetch_tap(0);

    for (int i = tid; i < HALF * LD; i += Cfg::THREADS) {
        act[i] = 0.f;
        act[(R + HALF) * LD + i] = 0.f;
    }
    for (int i = tid; i < R * (C / 4); i += Cfg::THREADS) {
        int r = i / (C / 4), c4 = i % (C / 4);
        int g = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= 0 && g < total_rows)
            v = *reinterpret_cast<const float4*>(x + (size_t)g * C + 4 * c4);
        *reinterpret_cast<float4*>(act + (r + HALF) * LD + 4 * c4) = v;
    }
    for (int r = tid; r < R; r += Cfg::THREADS) {
        int g = row0 + r;
        valid[r] = (g >= 0 && g < total_rows) ? (__ldg(row_seq + g) >= 0) : 0;
    }

    const int tx = tid & 7;        // channel group: co = tx * CT .. + CT
    const int ty = tid >> 3;       // rows ty + 32 i, i = 0..3

    for (int layer = 0; layer < n_layers; ++layer) {
        float acc[4][CT];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;

#pragma unroll 1
        for (int tap = 0; tap < KS; ++tap) {
            const int q = layer * KS + tap;
            // the other buffer was last read by tap q - 1, whose trailing barrier
            // every thread has passed
            if (q + 1 < n_taps) {
                fetch_tap(q + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float* a0 = act + (ty + tap) * LD;
            const float* wt = wbuf + (q & 1) * Cfg::TAP_FLOATS + tx * CT;
#pragma unroll 2
            for (int c4 = 0; c4 < C / 4; ++c4) {
                float4 a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[i] = *reinterpret_cast<const float4*>(a0 + 32 * i * LD + 4 * c4);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float2 wv[CT / 2];
#pragma unroll
                    for (int j = 0; j < CT / 2; ++j)
                        wv[j] = *reinterpret_cast<const float2*>(wt + (4 * c4 + cc) * C + 2 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float av = cc == 0 ? a[i].x : cc == 1 ? a[i].y : cc == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int j = 0; j < CT / 2; ++j) {
                            acc[i][2 * j] = fmaf(av, wv[j].x, acc[i][2 * j]);
                            acc[i][2 * j + 1] = fmaf(av, wv[j].y, acc[i][2 * j + 1]);
                        }
                    }
                }
            }
            __syncthreads();   // this tap's weights (and, after the last tap, the input) are consumed
        }

        const int a = acts.act[layer];
        const float* b = bias + layer * C + tx * CT;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r = ty + 32 * i;
            float keep = valid[r] ? 1.f : 0.f;
            float* dst = act + (r + HALF) * LD + tx * CT;
#pragma unroll
            for (int j = 0; j < CT; j += 2) {
                float v0 = apply_activation(acc[i][j] + __ldg(b + j), a);
                float v1 = apply_activation(acc[i][j + 1] + __ldg(b + j + 1), a);
                v0 = keep != 0.f ? v0 : 0.f;
                v1 = keep != 0.f ? v1 : 0.f;
                *reinterpret_cast<float2*>(dst + j) = make_float2(v0, v1);
            }
        }
        __syncthreads();
    }

    for (int i = tid; i < tile_rows * (C / 4); i += Cfg::THREADS) {
        int r = halo + i / (C / 4), c4 = i % (C / 4);
        int g = row0 + r;
        if (g < total_rows)
            *reinterpret_cast<float4*>(y + (size_t)g * C + 4 * c4) =
                *reinterpret_cast<const float4*>(act + (r + HALF) * LD + 4 * c4);
    }
}

template <int C, int KS>
int launch_conv_f32_tap(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const float* bias, const int32_t* acts_host,
    int32_t n_layers, float* y, cudaStream_t stream) {
    using Cfg = ConvF32Tap<C, KS>;
    const int halo = n_layers * Cfg::HALF;
    const int tile_rows = Cfg::R - 2 * halo;
    EMPH_REQUIRE(tile_rows >= 32, "emph_conv_stack: %d layers of kernel %d leave no tile", n_layers, KS);
    ConvActs acts;
    for (int i = 0; i < kConvMaxLayers; ++i) acts.act[i] = i < n_layers ? acts_host[i] : 0;
    int s = check_cuda(
        cudaFuncSetAttribute(
            conv_stack_f32_tap_kernel<C, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)Cfg::SMEM),
        "conv_f32 (tap-streamed) smem attribute");
    if (s != EMPH_OK) return s;
    int grid = (total_rows + tile_rows - 1) / tile_rows;
    conv_stack_f32_tap_kernel<C, KS><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(
        x, row_seq, total_rows, weights, bias, acts, n_layers, tile_rows, y);
    EMPH_CHECK_LAUNCH("emph_conv_stack(fp32, tap-streamed)");
    return EMPH_OK;
}

int conv_stack_f32(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const float* bias, const int32_t* acts_host,
    int32_t n_layers, int32_t channels, int32_t kernel_size, float* y,
    cudaStream_t stream) {
    if (channels == 80 && kernel_size == 3)
        return launch_conv_f32<80, 3>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 80 && kernel_size == 1)
        return launch_conv_f32<80, 1>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 80 && kernel_size == 5)
        return launch_conv_f32<80, 5, false>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 80 && kernel_size == 7)
        return launch_conv_f32<80, 7, false>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 128 && kernel_size == 3)
        return launch_conv_f32_tap<128, 3>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 128 && kernel_size == 1)
        return launch_conv_f32_tap<128, 1>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 128 && kernel_size == 5)
        return launch_conv_f32_tap<128, 5>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    if (channels == 128 && kernel_size == 7)
        return launch_conv_f32_tap<128, 7>(x, row_seq, total_rows, weights, bias, acts_host, n_layers, y, stream);
    set_error("emph_conv_stack(fp32): channels=%d kernel_size=%d not compiled in", channels, kernel_size);
    return EMPH_ENOSYS;
}

int conv_stack_bf16_tc(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const float* bias, const int32_t* acts_host,
    int32_t n_layers, int32_t channels, int32_t kernel_size, float* y,
    cudaStream_t stream);
int conv_stack_bf16x3_tc(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const int32_t* acts_host,
    int32_t n_layers, int32_t channels, int32_t kernel_size, float* y,
    cudaStream_t stream);
int conv_stack_bf16x6_tc(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const int32_t* acts_host,
    int32_t n_layers, int32_t channels, int32_t kernel_size, float* y,
    cudaStream_t stream);

}  // namespace emph

extern "C" int emph_conv_stack(
    const float* x, const int32_t* row_seq, int32_t total_rows,
    const float* weights, const float* bias, const int32_t* acts_host,
    int32_t n_layers, int32_t channels, int32_t kernel_size,
    int32_t precision, float* y, void* stream) {
    EMPH_REQUIRE(total_rows >= 0, "emph_conv_stack: negative rows");
    EMPH_REQUIRE(n_layers > 0 && n_layers <= emph::kConvMaxLayers,
                 "emph_conv_stack: n_layers %d out of range", n_layers);
    EMPH_REQUIRE(x != y, "emph_conv_stack: x and y must not alias (halo rows)");
    if (total_rows == 0) return EMPH_OK;
    if (precision == EMPH_PREC_FP32)
        return emph::conv_stack_f32(x, row_seq, total_rows, weights, bias, acts_host,
                                    n_layers, channels, kernel_size, y, (cudaStream_t)stream);
    if (precision == EMPH_PREC_BF16_TC)
        return emph::conv_stack_bf16_tc(x, row_seq, total_rows, weights, bias, acts_host,
                                        n_layers, channels, kernel_size, y, (cudaStream_t)stream);
    if (precision == EMPH_PREC_BF16X3_TC)
        return emph::conv_stack_bf16x3_tc(x, row_seq, total_rows, weights, acts_host,
                                          n_layers, channels, kernel_size, y, (cudaStream_t)stream);
    if (precision == EMPH_PREC_BF16X6_TC)
        return emph::conv_stack_bf16x6_tc(x, row_seq, total_rows, weights, acts_host,
                                          n_layers, channels, kernel_size, y, (cudaStream_t)stream);
    emph::set_error("emph_conv_stack: unknown precision %d", precision);
    return EMPH_EINVAL;
}
