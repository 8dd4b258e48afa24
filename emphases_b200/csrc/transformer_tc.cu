// Fused per-row passes of one Transformer encoder layer
// (emphases/model/layers/transformer.py:18-23: nn.TransformerEncoderLayer,
// d_model 80, 2 heads, dim_feedforward 80, post-norm, ReLU), sm_100a.
//
// Around the attention kernel (attention_tc.cu) a layer is three passes over the
// packed rows instead of five linear-map launches, two LayerNorm passes and a
// staging pass:
//   emph_transformer_qkv        q = x Wq^T + bq as fp32 rows; k and v straight
//                               into the 16-bit records the attention kernel reads
//   emph_transformer_proj_norm  y = LayerNorm(residual + x Wo^T + bo)
//   emph_transformer_ffn_norm   y = LayerNorm(x + relu(x W1^T + b1) W2^T + b2)
//
// All of them are 80 x 80 matrix products per 16-row warp tile on mma.sync
// m16n8k16 with fp32 accumulators, at fp32 grade: activations and weights are
// split into NP bf16 parts (NP = 2: hi*hi + hi*lo + lo*hi, 2^-17 per product;
// NP = 3: six products, 2^-24) exactly like the tcgen05 conv stack's bf16x3 /
// bf16x6 modes.  NP = 1 is one fp16 value per operand (2^-12, like the
// attention kernel's fp16 form): the 2e-3 'bf16' mode, bound by HBM alone.  The activation fragments are built in registers from fp32 row
// loads (an A fragment's elements are the thread's own (row, column) pairs, the
// same positions the accumulators use, so the feed-forward block chains its two
// products without leaving registers and the residual lines up with the
// output); the weight parts sit in shared memory as [out][in + 8] rows (176
// bytes = 16 mod 32: conflict-free ldmatrix).  LayerNorm statistics of a row
// live in one lane quad: two shuffles per reduction.
#include "attention_tc.cuh"

namespace emph {
namespace xf_tc {

using attn_tc::kPlainFp16;
using attn_tc::kSplitBf16;
using attn_tc::smem_u32;

constexpr int C = 80;                       // channels = d_model
constexpr int KS = C / 16;                  // k-steps
constexpr int NT = C / 8;                   // n-tiles
constexpr int kWStride = C + 8;             // weight row, 16-bit elements
constexpr int kWPartBytes = C * kWStride * 2;          // 14,080
constexpr int kThreads = 256;
constexpr int kRowsPerCta = 16 * (kThreads / 32);      // 128

// operand type of the NP-part form: one fp16, or NP bf16 parts
template <int NP>
struct Operand {
    static constexpr int kMode = NP == 1 ? kPlainFp16 : kSplitBf16;
};

// (lo, hi) fp32 pair -> NP bf16 pairs whose sum is the pair to 2^-(8 NP + 1)
// (NP = 1: one fp16 pair)
template <int NP>
__device__ __forceinline__ void split_pair(float lo, float hi, uint32_t (&parts)[NP]) {
    if (NP == 1) {
        parts[0] = attn_tc::pack_pair<kPlainFp16>(lo, hi);
        return;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        parts[p] = attn_tc::pack_pair<kSplitBf16>(lo, hi);
        lo -= __uint_as_float(parts[p] << 16);
        hi -= __uint_as_float(parts[p] & 0xffff0000u);
    }
}

// A fragments (all k-steps, all parts) of 16 rows held as accumulator-layout
// values v[n-tile][4]: step s takes tiles 2s and 2s + 1
template <int NP>
__device__ __forceinline__ void fragments_from_tiles(
    const float (&v)[NT][4], uint32_t (&a)[NP][KS][4]) {
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t parts[NP];
            split_pair<NP>(v[2 * s + (i >> 1)][2 * (i & 1)], v[2 * s + (i >> 1)][2 * (i & 1) + 1], parts);
#pragma unroll
            for (int p = 0; p < NP; ++p) a[p][s][i] = parts[p];
        }
}

// fp32 rows -> accumulator-layout values: (row g: [0], [1]; row g + 8: [2], [3]) x columns 8j + 2t
__device__ __forceinline__ void load_tiles(
    const float* __restrict__ x, int row0, int total_rows, int lane, float (&v)[NT][4]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = row0 + g + 8 * r;
        const float* src = x + (size_t)row * C + 2 * t;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            float2 pair = make_float2(0.f, 0.f);
            if (row < total_rows) pair = *reinterpret_cast<const float2*>(src + 8 * j);
            v[j][2 * r] = pair.x;
            v[j][2 * r + 1] = pair.y;
        }
    }
}

// acc[j] += A W^T for one 80 x 80 matrix whose NP parts start at `w` in shared
// memory; products with part indices pa + pw < NP, the SMALLEST terms first: the
// tensor core truncates every addition at the accumulator's magnitude, so the
// low-order products (2^-9, 2^-18 of the result) are summed while the
// accumulator is still that small and only the KS hi * hi steps run at full
// magnitude (measured on the LayerNorm output: 3.6e-6 -> see the kernel test)
template <int NP>
__device__ __forceinline__ void product(
    float (&acc)[NT][4], const uint32_t (&a)[NP][KS][4], uint32_t w, int lane) {
    const uint32_t lane_rows = (uint32_t)(lane & 7) * (kWStride * 2) + 16u * (lane >> 3);
    const uint32_t lane_tail = (uint32_t)(lane & 7) * (kWStride * 2) + 16u * ((lane >> 3) & 1);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int level = NP - 1; level >= 0; --level) {
#pragma unroll
            for (int pw = 0; pw <= level; ++pw) {
                const int pa = level - pw;
                const uint32_t rows = w + pw * kWPartBytes + 8 * j * (kWStride * 2);
                uint32_t b[KS][2];
#pragma unroll
                for (int s2 = 0; s2 < KS / 2; ++s2) {
                    uint32_t r4[4];
                    attn_tc::ldmatrix_x4(r4, rows + lane_rows + 64 * s2);
                    b[2 * s2][0] = r4[0]; b[2 * s2][1] = r4[1];
                    b[2 * s2 + 1][0] = r4[2]; b[2 * s2 + 1][1] = r4[3];
                }
                if (KS & 1) attn_tc::ldmatrix_x2(b[KS - 1], rows + lane_tail + 32 * (KS - 1));
#pragma unroll
                for (int s = 0; s < KS; ++s)
                    attn_tc::mma_16816<Operand<NP>::kMode>(acc[j], a[pa][s], b[s][0], b[s][1]);
            }
        }
    }
}

__device__ __forceinline__ void clear(float (&acc)[NT][4]) {
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
}

__device__ __forceinline__ void add_bias(
    float (&acc)[NT][4], const float* __restrict__ bias, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(bias + 8 * j + 2 * t));
        acc[j][0] += b.x; acc[j][2] += b.x;
        acc[j][1] += b.y; acc[j][3] += b.y;
    }
}

// weight parts of `matrices` matrices: global [matrix][part][out][in + 8] -> shared
__device__ __forceinline__ void load_weights(
    unsigned char* smem, const unsigned char* __restrict__ weights, int bytes) {
    for (int i = threadIdx.x * 16; i < bytes; i += kThreads * 16)
        *reinterpret_cast<uint4*>(smem + i) = __ldg(reinterpret_cast<const uint4*>(weights + i));
    __syncthreads();
}

// v <- LayerNorm(v) * gamma + beta for the warp's 16 rows (a row's 80 values sit
// in one lane quad)
__device__ __forceinline__ void layernorm_values(
    float (&v)[NT][4], const float* __restrict__ gamma, const float* __restrict__ beta,
    float eps, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) sum += v[j][2 * r] + v[j][2 * r + 1];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum * (1.f / C);
        float square = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float d0 = v[j][2 * r] - mean, d1 = v[j][2 * r + 1] - mean;
            square = fmaf(d0, d0, square);
            square = fmaf(d1, d1, square);
        }
        square += __shfl_xor_sync(0xffffffffu, square, 1);
        square += __shfl_xor_sync(0xffffffffu, square, 2);
        const float inv = rsqrtf(square * (1.f / C) + eps);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma + 8 * j + 2 * t));
            const float2 bt = __ldg(reinterpret_cast<const float2*>(beta + 8 * j + 2 * t));
            v[j][2 * r] = (v[j][2 * r] - mean) * inv * gm.x + bt.x;
            v[j][2 * r + 1] = (v[j][2 * r + 1] - mean) * inv * gm.y + bt.y;
        }
    }
}

// the warp's 16 rows to y; zeros on separator rows
__device__ __forceinline__ void store_rows(
    const float (&v)[NT][4], const int32_t* __restrict__ row_seq, int row0, int total_rows,
    int lane, float* __restrict__ y) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = row0 + g + 8 * r;
        if (row >= total_rows) continue;
        const bool separator = row_seq[row] < 0;
        float* dst = y + (size_t)row * C + 2 * t;
#pragma unroll
        for (int j = 0; j < NT; ++j)
            *reinterpret_cast<float2*>(dst + 8 * j) = make_float2(
                separator ? 0.f : v[j][2 * r], separator ? 0.f : v[j][2 * r + 1]);
    }
}

__device__ __forceinline__ void layernorm_store(
    float (&v)[NT][4], const float* __restrict__ gamma, const float* __restrict__ beta,
    float eps, const int32_t* __restrict__ row_seq, int row0, int total_rows, int lane,
    float* __restrict__ y) {
    layernorm_values(v, gamma, beta, eps, lane);
    store_rows(v, row_seq, row0, total_rows, lane, y);
}

// ---------------------------------------------------------------------------
// q, k, v projections of one layer.  weights: [3][NP][80][88] bf16 (q, k, v),
// bias [240].  q goes out as fp32 rows; k and v as the attention kernel's
// records (attention_tc.cuh Layout<40, ATT>), so no staging pass follows.  Every
// byte of a row's records that the attention kernel reads is written here; the
// caller keeps the 64 records of slack behind each head's rows zero.
// ---------------------------------------------------------------------------
template <int NP, int ATT>
__global__ void __launch_bounds__(kThreads, NP == 1 ? 2 : 1)
qkv_kernel(
    const float* __restrict__ x, int total_rows, const unsigned char* __restrict__ weights,
    const float* __restrict__ bias, float* __restrict__ q, unsigned char* __restrict__ staged,
    int padded_rows) {
    constexpr int D = C / 2;                                 // two heads
    using L = attn_tc::Layout<D, ATT>;
    extern __shared__ __align__(128) unsigned char smem[];
    load_weights(smem, weights, 3 * NP * kWPartBytes);
    const uint32_t w = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int tile = blockIdx.x; tile * kRowsPerCta < total_rows; tile += gridDim.x) {
        const int row0 = tile * kRowsPerCta + 16 * warp;
        if (row0 >= total_rows) continue;
        uint32_t a[NP][KS][4];
        {
            float v[NT][4];
            load_tiles(x, row0, total_rows, lane, v);
            fragments_from_tiles<NP>(v, a);
        }
#pragma unroll 1
        for (int matrix = 0; matrix < 3; ++matrix) {
            float acc[NT][4];
            clear(acc);
            product<NP>(acc, a, w + matrix * NP * kWPartBytes, lane);
            add_bias(acc, bias + matrix * C, lane);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + g + 8 * r;
                if (row >= total_rows) continue;
                if (matrix == 0) {
                    float* dst = q + (size_t)row * C + 2 * t;
#pragma unroll
                    for (int j = 0; j < NT; ++j)
                        *reinterpret_cast<float2*>(dst + 8 * j) =
                            make_float2(acc[j][2 * r], acc[j][2 * r + 1]);
                    continue;
                }
                // element offset of this matrix's first part inside a record
                const int first = matrix == 1 ? 0 : L::kOffV;
                const int next_part = matrix == 1 ? L::DP : D;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int head = (8 * j) / D, dim = 8 * j - head * D + 2 * t;
                    unsigned char* record =
                        staged + ((size_t)head * padded_rows + row) * L::kRecord;
                    const uint32_t hi = attn_tc::pack_pair<ATT>(acc[j][2 * r], acc[j][2 * r + 1]);
                    *reinterpret_cast<uint32_t*>(record + 2 * (first + dim)) = hi;
                    if (ATT == kSplitBf16)
                        *reinterpret_cast<uint32_t*>(record + 2 * (first + next_part + dim)) =
                            attn_tc::pack_residual(acc[j][2 * r], acc[j][2 * r + 1], hi);
                }
                // K's padded dims [D, DP) are zeros in every part (the lane quad
                // of a row covers the 8 of them)
                if (matrix == 1 && L::DP - D == 8) {
#pragma unroll
                    for (int head = 0; head < 2; ++head) {
                        unsigned char* record =
                            staged + ((size_t)head * padded_rows + row) * L::kRecord;
#pragma unroll
                        for (int part = 0; part < L::NP; ++part)
                            *reinterpret_cast<uint32_t*>(
                                record + 2 * (part * L::DP + D + 2 * t)) = 0u;
                    }
                }
            }
        }
    }
}

// y = LayerNorm(residual + x W^T + b).  weights: [1][NP][80][88]
template <int NP>
__global__ void __launch_bounds__(kThreads, 2)
proj_norm_kernel(
    const float* __restrict__ x, const float* __restrict__ residual, int total_rows,
    const unsigned char* __restrict__ weights, const float* __restrict__ bias,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    const int32_t* __restrict__ row_seq, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    load_weights(smem, weights, NP * kWPartBytes);
    const uint32_t w = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int tile = blockIdx.x; tile * kRowsPerCta < total_rows; tile += gridDim.x) {
        const int row0 = tile * kRowsPerCta + 16 * warp;
        if (row0 >= total_rows) continue;
        uint32_t a[NP][KS][4];
        float v[NT][4];
        load_tiles(x, row0, total_rows, lane, v);
        fragments_from_tiles<NP>(v, a);
        float acc[NT][4];
        clear(acc);
        product<NP>(acc, a, w, lane);
        add_bias(acc, bias, lane);
        load_tiles(residual, row0, total_rows, lane, v);
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] += v[j][i];
        layernorm_store(acc, gamma, beta, eps, row_seq, row0, total_rows, lane, y);
    }
}

// y = LayerNorm(x + relu(x W1^T + b1) W2^T + b2).  weights: [2][NP][80][88], bias [160]
template <int NP>
__global__ void __launch_bounds__(kThreads, 2)
ffn_norm_kernel(
    const float* __restrict__ x, int total_rows, const unsigned char* __restrict__ weights,
    const float* __restrict__ bias, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, const int32_t* __restrict__ row_seq,
    float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    load_weights(smem, weights, 2 * NP * kWPartBytes);
    const uint32_t w = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int tile = blockIdx.x; tile * kRowsPerCta < total_rows; tile += gridDim.x) {
        const int row0 = tile * kRowsPerCta + 16 * warp;
        if (row0 >= total_rows) continue;
        uint32_t a[NP][KS][4];
        float v[NT][4];
        load_tiles(x, row0, total_rows, lane, v);
        fragments_from_tiles<NP>(v, a);
        float acc[NT][4];
        clear(acc);
        product<NP>(acc, a, w, lane);
        add_bias(acc, bias, lane);
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaxf(acc[j][i], 0.f);
        fragments_from_tiles<NP>(acc, a);          // the hidden layer never leaves registers
        clear(acc);
        product<NP>(acc, a, w + NP * kWPartBytes, lane);
        add_bias(acc, bias + C, lane);
        load_tiles(x, row0, total_rows, lane, v);  // the residual (an L2 hit)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] += v[j][i];
        layernorm_store(acc, gamma, beta, eps, row_seq, row0, total_rows, lane, y);
    }
}

// Everything of a layer behind the attention in one pass:
//   n = LayerNorm1(residual + x Wo^T + bo);  y = LayerNorm2(n + relu(n W1^T + b1) W2^T + b2)
// weights: [3][NP][80][88] (out-projection, linear1, linear2), bias [240]; the
// intermediate n never leaves registers
template <int NP>
__global__ void __launch_bounds__(kThreads, 1)
layer_tail_kernel(
    const float* __restrict__ x, const float* __restrict__ residual, int total_rows,
    const unsigned char* __restrict__ weights, const float* __restrict__ bias,
    const float* __restrict__ gamma1, const float* __restrict__ beta1,
    const float* __restrict__ gamma2, const float* __restrict__ beta2, float eps,
    const int32_t* __restrict__ row_seq, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem[];
    load_weights(smem, weights, 3 * NP * kWPartBytes);
    const uint32_t w = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int tile = blockIdx.x; tile * kRowsPerCta < total_rows; tile += gridDim.x) {
        const int row0 = tile * kRowsPerCta + 16 * warp;
        if (row0 >= total_rows) continue;
        uint32_t a[NP][KS][4];
        float normed[NT][4];
        {
            float v[NT][4];
            load_tiles(x, row0, total_rows, lane, v);
            fragments_from_tiles<NP>(v, a);
            clear(normed);
            product<NP>(normed, a, w, lane);
            add_bias(normed, bias, lane);
            load_tiles(residual, row0, total_rows, lane, v);
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) normed[j][i] += v[j][i];
        }
        layernorm_values(normed, gamma1, beta1, eps, lane);
        fragments_from_tiles<NP>(normed, a);
        float acc[NT][4];
        clear(acc);
        product<NP>(acc, a, w + NP * kWPartBytes, lane);
        add_bias(acc, bias + C, lane);
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaxf(acc[j][i], 0.f);
        fragments_from_tiles<NP>(acc, a);
        clear(acc);
        product<NP>(acc, a, w + 2 * NP * kWPartBytes, lane);
        add_bias(acc, bias + 2 * C, lane);
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] += normed[j][i];
        layernorm_store(acc, gamma2, beta2, eps, row_seq, row0, total_rows, lane, y);
    }
}

template <typename Kernel>
int configure(Kernel kernel, int smem_bytes, const char* what) {
    return check_cuda(
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes),
        what);
}

inline int grid_for(int total_rows, int ctas_per_sm) {
    const int tiles = (total_rows + kRowsPerCta - 1) / kRowsPerCta;
    return tiles < sm_count() * ctas_per_sm ? tiles : sm_count() * ctas_per_sm;
}

}  // namespace xf_tc
}  // namespace emph

extern "C" {

#define EMPH_XF_REQUIRE_SHAPE(name)                                                      \
    EMPH_REQUIRE(channels == emph::xf_tc::C, name ": compiled for 80 channels, got %d",  \
                 channels);                                                              \
    EMPH_REQUIRE(parts >= 1 && parts <= 3, name ": parts must be 1, 2 or 3, got %d", parts)

int emph_transformer_qkv(
    const float* x, int32_t total_rows, int32_t channels, const void* weights, const float* bias,
    int32_t parts, int32_t attention_mode, float* q, void* staged, int64_t staged_bytes,
    void* stream) {
    using namespace emph::xf_tc;
    EMPH_XF_REQUIRE_SHAPE("emph_transformer_qkv");
    EMPH_REQUIRE(attention_mode == kPlainFp16 || attention_mode == kSplitBf16,
                 "emph_transformer_qkv: unknown attention operand mode %d", attention_mode);
    if (total_rows <= 0) return EMPH_OK;
    const int padded_rows = total_rows + emph::attn_tc::kKeys;
    const int64_t record = attention_mode == kSplitBf16
        ? emph::attn_tc::Layout<C / 2, kSplitBf16>::kRecord
        : emph::attn_tc::Layout<C / 2, kPlainFp16>::kRecord;
    EMPH_REQUIRE(staged_bytes >= 2 * (int64_t)padded_rows * record,
                 "emph_transformer_qkv: record buffer of %lld bytes is too small",
                 (long long)staged_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 3 * parts * kWPartBytes;
    const int grid = grid_for(total_rows, parts == 1 ? 2 : 1);
#define EMPH_XF_QKV(NP, ATT)                                                              \
    do {                                                                                  \
        const int status = configure(qkv_kernel<NP, ATT>, smem, "emph_transformer_qkv"); \
        if (status != EMPH_OK) return status;                                             \
        qkv_kernel<NP, ATT><<<grid, kThreads, smem, st>>>(                                \
            x, total_rows, (const unsigned char*)weights, bias, q, (unsigned char*)staged, \
            padded_rows);                                                                 \
    } while (0)
    if (parts == 1 && attention_mode == kPlainFp16) EMPH_XF_QKV(1, kPlainFp16);
    else if (parts == 1) EMPH_XF_QKV(1, kSplitBf16);
    else if (parts == 2 && attention_mode == kPlainFp16) EMPH_XF_QKV(2, kPlainFp16);
    else if (parts == 2) EMPH_XF_QKV(2, kSplitBf16);
    else if (attention_mode == kPlainFp16) EMPH_XF_QKV(3, kPlainFp16);
    else EMPH_XF_QKV(3, kSplitBf16);
#undef EMPH_XF_QKV
    EMPH_CHECK_LAUNCH("emph_transformer_qkv");
    return EMPH_OK;
}

int emph_transformer_proj_norm(
    const float* x, const float* residual, int32_t total_rows, int32_t channels,
    const void* weights, const float* bias, int32_t parts, const float* gamma,
    const float* beta, float eps, const int32_t* row_seq, float* y, void* stream) {
    using namespace emph::xf_tc;
    EMPH_XF_REQUIRE_SHAPE("emph_transformer_proj_norm");
    if (total_rows <= 0) return EMPH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = parts * kWPartBytes;
    const int grid = grid_for(total_rows, 2);
    if (parts == 1) {
        const int status = configure(proj_norm_kernel<1>, smem, "emph_transformer_proj_norm");
        if (status != EMPH_OK) return status;
        proj_norm_kernel<1><<<grid, kThreads, smem, st>>>(
            x, residual, total_rows, (const unsigned char*)weights, bias, gamma, beta, eps,
            row_seq, y);
    } else if (parts == 2) {
        const int status = configure(proj_norm_kernel<2>, smem, "emph_transformer_proj_norm");
        if (status != EMPH_OK) return status;
        proj_norm_kernel<2><<<grid, kThreads, smem, st>>>(
            x, residual, total_rows, (const unsigned char*)weights, bias, gamma, beta, eps,
            row_seq, y);
    } else {
        const int status = configure(proj_norm_kernel<3>, smem, "emph_transformer_proj_norm");
        if (status != EMPH_OK) return status;
        proj_norm_kernel<3><<<grid, kThreads, smem, st>>>(
            x, residual, total_rows, (const unsigned char*)weights, bias, gamma, beta, eps,
            row_seq, y);
    }
    EMPH_CHECK_LAUNCH("emph_transformer_proj_norm");
    return EMPH_OK;
}

int emph_transformer_ffn_norm(
    const float* x, int32_t total_rows, int32_t channels, const void* weights, const float* bias,
    int32_t parts, const float* gamma, const float* beta, float eps, const int32_t* row_seq,
    float* y, void* stream) {
    using namespace emph::xf_tc;
    EMPH_XF_REQUIRE_SHAPE("emph_transformer_ffn_norm");
    if (total_rows <= 0) return EMPH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 2 * parts * kWPartBytes;
    const int grid = grid_for(total_rows, 2);
    if (parts == 1) {
        const int status = configure(ffn_norm_kernel<1>, smem, "emph_transformer_ffn_norm");
        if (status != EMPH_OK) return status;
        ffn_norm_kernel<1><<<grid, kThreads, smem, st>>>(
            x, total_rows, (const unsigned char*)weights, bias, gamma, beta, eps, row_seq, y);
    } else if (parts == 2) {
        const int status = configure(ffn_norm_kernel<2>, smem, "emph_transformer_ffn_norm");
        if (status != EMPH_OK) return status;
        ffn_norm_kernel<2><<<grid, kThreads, smem, st>>>(
            x, total_rows, (const unsigned char*)weights, bias, gamma, beta, eps, row_seq, y);
    } else {
        const int status = configure(ffn_norm_kernel<3>, smem, "emph_transformer_ffn_norm");
        if (status != EMPH_OK) return status;
        ffn_norm_kernel<3><<<grid, kThreads, smem, st>>>(
            x, total_rows, (const unsigned char*)weights, bias, gamma, beta, eps, row_seq, y);
    }
    EMPH_CHECK_LAUNCH("emph_transformer_ffn_norm");
    return EMPH_OK;
}

int emph_transformer_layer_tail(
    const float* x, const float* residual, int32_t total_rows, int32_t channels,
    const void* weights, const float* bias, int32_t parts, const float* gamma1,
    const float* beta1, const float* gamma2, const float* beta2, float eps,
    const int32_t* row_seq, float* y, void* stream) {
    using namespace emph::xf_tc;
    EMPH_XF_REQUIRE_SHAPE("emph_transformer_layer_tail");
    if (total_rows <= 0) return EMPH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 3 * parts * kWPartBytes;
    const int grid = grid_for(total_rows, 1);
#define EMPH_XF_TAIL(NP)                                                                  \
    do {                                                                                  \
        const int status =                                                                \
            configure(layer_tail_kernel<NP>, smem, "emph_transformer_layer_tail");        \
        if (status != EMPH_OK) return status;                                             \
        layer_tail_kernel<NP><<<grid, kThreads, smem, st>>>(                              \
            x, residual, total_rows, (const unsigned char*)weights, bias, gamma1, beta1,  \
            gamma2, beta2, eps, row_seq, y);                                              \
    } while (0)
    if (parts == 1) EMPH_XF_TAIL(1);
    else if (parts == 2) EMPH_XF_TAIL(2);
    else EMPH_XF_TAIL(3);
#undef EMPH_XF_TAIL
    EMPH_CHECK_LAUNCH("emph_transformer_layer_tail");
    return EMPH_OK;
}

#undef EMPH_XF_REQUIRE_SHAPE

}  // extern "C"
