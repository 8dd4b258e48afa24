// Transformer-variant kernels over packed rows, fp32, sm_100a.
//
// Replace the pieces of emphases/model/layers/transformer.py:13-52
// (nn.TransformerEncoder of post-norm layers, d_model 80, 2 heads,
// dim_feedforward 80, key-padding mask from the sequence lengths) that are
// not plain per-row linear maps (those run through emph_conv_stack with
// kernel_size 1):
//   emph_add_positional   x + PE[:T]                      transformer.py:50-52
//   emph_attention_rows   softmax(Q K^T / sqrt(d) + key mask) V, per head
//   emph_add_layernorm    LayerNorm(x + residual), eps 1e-5
// In the packed layout the key-padding mask is "keys of the same sequence
// with index < n_keys[u]" -- block-diagonal attention with no padded FLOPs.
#include <math_constants.h>

#include "common.cuh"

namespace emph {

constexpr int kAttnThreads = 128;

// 2^x, x <= 0 here (ex2.approx: relative error 2^-22; exp2(-inf) = 0)
__device__ __forceinline__ float exp2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 128 queries per CTA (the host cuts query blocks of that size): a lane PAIR owns two queries
constexpr int kAttnK = 64;     // keys per shared-memory tile

// Block-diagonal attention with online softmax, fp32.  Two lanes share two
// queries: lane h of the pair holds dims [h D/2, (h+1) D/2) of both queries'
// q and o vectors, so every 16-byte K / V shared-memory load (a broadcast: the
// 16 pairs of a warp read two addresses) feeds 8 FMAs instead of 4 and the
// partial dot products meet through one shuffle per query and key.  (One
// query per thread, the first version, was bound by the shared-memory pipe at
// 75 % with the FP32 pipe at 45 %, profiles/r01z_attention.md.)
template <int D>
__global__ void __launch_bounds__(kAttnThreads)
attention_rows_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    int channels, const int32_t* __restrict__ row_start, const int32_t* __restrict__ n_queries,
    const int32_t* __restrict__ n_keys,
    const int32_t* __restrict__ block_seq, const int32_t* __restrict__ block_q0,
    float scale, float* __restrict__ out) {
    constexpr int H = D / 2;                   // dims per lane
    static_assert(H % 4 == 0, "head dim must be a multiple of 8");
    __shared__ __align__(16) float ks[kAttnK][D];
    __shared__ __align__(16) float vs[kAttnK][D];
    const int u = block_seq[blockIdx.x];
    const int q0 = block_q0[blockIdx.x];
    const int head = blockIdx.y;
    const int base = row_start[u];
    const int nk = n_keys[u];
    const int nq = n_queries[u];
    const int half = threadIdx.x & 1;
    const int qa = q0 + 2 * (threadIdx.x >> 1), qb = qa + 1;
    const bool active_a = qa < nq, active_b = qb < nq;
    const size_t column = (size_t)head * D + half * H;

    // scores are kept in the log2 domain (log2 e folded into the query scale):
    // every exponential is then one MUFU.EX2
    scale *= 1.4426950408889634f;
    float qra[H], qrb[H], oa[H], ob[H];
#pragma unroll
    for (int d = 0; d < H; ++d) {
        qra[d] = active_a ? q[(size_t)(base + qa) * channels + column + d] * scale : 0.f;
        qrb[d] = active_b ? q[(size_t)(base + qb) * channels + column + d] * scale : 0.f;
        oa[d] = 0.f;
        ob[d] = 0.f;
    }
    float ma = -CUDART_INF_F, la = 0.f, mb = -CUDART_INF_F, lb = 0.f;

    for (int kt = 0; kt < nk; kt += kAttnK) {
        const int count = min(kAttnK, nk - kt);
        __syncthreads();
        for (int i = threadIdx.x; i < count * (D / 4); i += kAttnThreads) {
            const int j = i / (D / 4), d4 = i % (D / 4);
            const size_t src = (size_t)(base + kt + j) * channels + (size_t)head * D + 4 * d4;
            *reinterpret_cast<float4*>(&ks[j][4 * d4]) = *reinterpret_cast<const float4*>(k + src);
            *reinterpret_cast<float4*>(&vs[j][4 * d4]) = *reinterpret_cast<const float4*>(v + src);
        }
        __syncthreads();
        // (inactive pairs run along: the shuffles below need every lane)
        for (int j = 0; j < count; ++j) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < H / 4; ++d4) {
                const float4 kk = *reinterpret_cast<const float4*>(&ks[j][half * H + 4 * d4]);
                sa = fmaf(qra[4 * d4], kk.x, sa);
                sa = fmaf(qra[4 * d4 + 1], kk.y, sa);
                sa = fmaf(qra[4 * d4 + 2], kk.z, sa);
                sa = fmaf(qra[4 * d4 + 3], kk.w, sa);
                sb = fmaf(qrb[4 * d4], kk.x, sb);
                sb = fmaf(qrb[4 * d4 + 1], kk.y, sb);
                sb = fmaf(qrb[4 * d4 + 2], kk.z, sb);
                sb = fmaf(qrb[4 * d4 + 3], kk.w, sb);
            }
            sa += __shfl_xor_sync(0xffffffffu, sa, 1);
            sb += __shfl_xor_sync(0xffffffffu, sb, 1);
            if (sa > ma) {                     // new running maximum: rescale
                const float correction = exp2_fast(ma - sa);
                la *= correction;
#pragma unroll
                for (int d = 0; d < H; ++d) oa[d] *= correction;
                ma = sa;
            }
            if (sb > mb) {
                const float correction = exp2_fast(mb - sb);
                lb *= correction;
#pragma unroll
                for (int d = 0; d < H; ++d) ob[d] *= correction;
                mb = sb;
            }
            const float pa = exp2_fast(sa - ma), pb = exp2_fast(sb - mb);
            la += pa;
            lb += pb;
#pragma unroll
            for (int d4 = 0; d4 < H / 4; ++d4) {
                const float4 vv = *reinterpret_cast<const float4*>(&vs[j][half * H + 4 * d4]);
                oa[4 * d4] = fmaf(pa, vv.x, oa[4 * d4]);
                oa[4 * d4 + 1] = fmaf(pa, vv.y, oa[4 * d4 + 1]);
                oa[4 * d4 + 2] = fmaf(pa, vv.z, oa[4 * d4 + 2]);
                oa[4 * d4 + 3] = fmaf(pa, vv.w, oa[4 * d4 + 3]);
                ob[4 * d4] = fmaf(pb, vv.x, ob[4 * d4]);
                ob[4 * d4 + 1] = fmaf(pb, vv.y, ob[4 * d4 + 1]);
                ob[4 * d4 + 2] = fmaf(pb, vv.z, ob[4 * d4 + 2]);
                ob[4 * d4 + 3] = fmaf(pb, vv.w, ob[4 * d4 + 3]);
            }
        }
    }
    if (active_a) {
        const float inv = 1.f / la;
        float* dst = out + (size_t)(base + qa) * channels + column;
#pragma unroll
        for (int d = 0; d < H; ++d) dst[d] = oa[d] * inv;
    }
    if (active_b) {
        const float inv = 1.f / lb;
        float* dst = out + (size_t)(base + qb) * channels + column;
#pragma unroll
        for (int d = 0; d < H; ++d) dst[d] = ob[d] * inv;
    }
}

// separator rows are zero
__global__ void attention_clear_kernel(
    const int32_t* __restrict__ row_seq, int total_rows, int channels, float* __restrict__ out) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= total_rows) return;
    if (row_seq[r] >= 0) return;
    for (int c = threadIdx.x & 31; c < channels; c += 32) out[(size_t)r * channels + c] = 0.f;
}

__global__ void add_positional_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ row_start,
    const int32_t* __restrict__ row_seq, int total_rows, int channels,
    const float* __restrict__ table, int table_rows, float* __restrict__ y) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= total_rows) return;
    const int u = row_seq[r];
    for (int c = threadIdx.x & 31; c < channels; c += 32) {
        float value = 0.f;
        if (u >= 0) {
            const int t = r - row_start[u];
            value = x[(size_t)r * channels + c] +
                    (t < table_rows ? table[(size_t)t * channels + c] : 0.f);
        }
        y[(size_t)r * channels + c] = value;
    }
}

// y = LayerNorm(x + residual) * gamma + beta, one warp per row
__global__ void add_layernorm_kernel(
    const float* __restrict__ x, const float* __restrict__ residual,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    const int32_t* __restrict__ row_seq, int total_rows, int channels, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= total_rows) return;
    if (row_seq[r] < 0) {
        for (int c = lane; c < channels; c += 32) y[(size_t)r * channels + c] = 0.f;
        return;
    }
    float values[4];                                // channels <= 128
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        values[i] = c < channels
            ? x[(size_t)r * channels + c] + residual[(size_t)r * channels + c] : 0.f;
        sum += values[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)channels;
    float square = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        const float d = c < channels ? values[i] - mean : 0.f;
        square = fmaf(d, d, square);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) square += __shfl_xor_sync(0xffffffffu, square, o);
    const float inv = rsqrtf(square / (float)channels + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = lane + 32 * i;
        if (c < channels)
            y[(size_t)r * channels + c] = (values[i] - mean) * inv * gamma[c] + beta[c];
    }
}

}  // namespace emph

extern "C" {

int emph_add_positional(
    const float* x, const int32_t* row_start, const int32_t* row_seq, int32_t total_rows,
    int32_t channels, const float* table, int32_t table_rows, float* y, void* stream) {
    EMPH_REQUIRE(total_rows >= 0 && channels > 0, "emph_add_positional: bad size");
    if (total_rows == 0) return EMPH_OK;
    emph::add_positional_kernel<<<(total_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        x, row_start, row_seq, total_rows, channels, table, table_rows, y);
    EMPH_CHECK_LAUNCH("emph_add_positional");
    return EMPH_OK;
}

int emph_attention_rows(
    const float* q, const float* k, const float* v, int32_t channels, int32_t heads,
    const int32_t* row_start, const int32_t* n_queries, const int32_t* n_keys,
    const int32_t* row_seq, int32_t total_rows,
    const int32_t* block_seq, const int32_t* block_q0, int32_t n_blocks,
    float scale, float* out, void* stream) {
    EMPH_REQUIRE(heads > 0 && channels % heads == 0, "emph_attention_rows: bad head count");
    const int head_dim = channels / heads;
    if (total_rows == 0) return EMPH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    emph::attention_clear_kernel<<<(total_rows + 7) / 8, 256, 0, st>>>(
        row_seq, total_rows, channels, out);
    EMPH_CHECK_LAUNCH("emph_attention_rows(clear)");
    if (n_blocks == 0) return EMPH_OK;
    dim3 grid(n_blocks, heads);
    if (head_dim == 40) {
        emph::attention_rows_kernel<40><<<grid, emph::kAttnThreads, 0, st>>>(
            q, k, v, channels, row_start, n_queries, n_keys, block_seq, block_q0, scale, out);
    } else if (head_dim == 32) {
        emph::attention_rows_kernel<32><<<grid, emph::kAttnThreads, 0, st>>>(
            q, k, v, channels, row_start, n_queries, n_keys, block_seq, block_q0, scale, out);
    } else if (head_dim == 64) {
        emph::attention_rows_kernel<64><<<grid, emph::kAttnThreads, 0, st>>>(
            q, k, v, channels, row_start, n_queries, n_keys, block_seq, block_q0, scale, out);
    } else {
        emph::set_error("emph_attention_rows: head_dim %d not compiled in", head_dim);
        return EMPH_ENOSYS;
    }
    EMPH_CHECK_LAUNCH("emph_attention_rows");
    return EMPH_OK;
}

int emph_add_layernorm(
    const float* x, const float* residual, const float* gamma, const float* beta, float eps,
    const int32_t* row_seq, int32_t total_rows, int32_t channels, float* y, void* stream) {
    EMPH_REQUIRE(channels > 0 && channels <= 128, "emph_add_layernorm: channels %d > 128", channels);
    if (total_rows == 0) return EMPH_OK;
    emph::add_layernorm_kernel<<<(total_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        x, residual, gamma, beta, eps, row_seq, total_rows, channels, y);
    EMPH_CHECK_LAUNCH("emph_add_layernorm");
    return EMPH_OK;
}

}  // extern "C"
