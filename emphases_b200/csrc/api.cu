// C-ABI plumbing: error string, version, small utility kernels.
#include <stdarg.h>

#include <mutex>
#include <string.h>

#include "common.cuh"

namespace emph {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list args;
    va_start(args, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, args);
    va_end(args);
}

// row -> sequence map by binary search over the sorted row_start array
__global__ void row_index_kernel(
    const int32_t* __restrict__ row_start, const int32_t* __restrict__ n_rows,
    int32_t n_seq, int32_t* __restrict__ row_seq, int32_t total_rows) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total_rows) return;
    int lo = 0, hi = n_seq;  // last u with row_start[u] <= r
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(row_start + mid) <= r) lo = mid + 1; else hi = mid;
    }
    int u = lo - 1;
    int seq = -1;
    if (u >= 0 && r < __ldg(row_start + u) + __ldg(n_rows + u)) seq = u;
    row_seq[r] = seq;
}

// (out, in, k) -> [k][in][out]
__global__ void pack_conv_weights_kernel(
    const float* __restrict__ w, int co, int ci, int ks, float* __restrict__ packed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = co * ci * ks;
    if (i >= total) return;
    int o = i % co;
    int c = (i / co) % ci;
    int k = i / (co * ci);
    packed[i] = w[(o * ci + c) * ks + k];
}

// (out, in, k) -> the ADJOINT conv in the same packed layout: [k][out][in] with
// the taps flipped (dX = conv(dY, W^T flipped), emphases_b200/training.py)
__global__ void pack_conv_weights_adjoint_kernel(
    const float* __restrict__ w, int co, int ci, int ks, float* __restrict__ packed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = co * ci * ks;
    if (i >= total) return;
    int c = i % ci;                 // output channel of the adjoint = input channel of the conv
    int o = (i / ci) % co;          // input channel of the adjoint = output channel of the conv
    int k = i / (co * ci);
    packed[i] = w[(o * ci + c) * ks + (ks - 1 - k)];
}

// (B, C, T) -> packed rows; 32x32 smem transpose tiles per sequence
__global__ void pack_rows_kernel(
    const float* __restrict__ bct, int channels, int frames,
    const int32_t* __restrict__ row_start, const int32_t* __restrict__ n_rows,
    float* __restrict__ rows) {
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int n = n_rows[b];
    if (t0 >= n) return;
    const float* src = bct + (size_t)b * channels * frames;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < channels && t < n) ? src[(size_t)c * frames + t] : 0.f;
    }
    __syncthreads();
    float* dst = rows + (size_t)row_start[b] * channels;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int t = t0 + i, c = c0 + threadIdx.x;
        if (t < n && c < channels) dst[(size_t)t * channels + c] = tile[threadIdx.x][i];
    }
}

__global__ void zero_separator_rows_kernel(
    const int32_t* __restrict__ row_seq, int total_rows, int channels,
    float* __restrict__ rows) {
    int r = blockIdx.x;
    if (r >= total_rows || row_seq[r] >= 0) return;
    for (int c = threadIdx.x; c < channels; c += blockDim.x)
        rows[(size_t)r * channels + c] = 0.f;
}

__global__ void unpack_rows_kernel(
    const float* __restrict__ rows, const int32_t* __restrict__ row_start,
    const int32_t* __restrict__ n_rows, int channels, int frames,
    float* __restrict__ bct) {
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int n = n_rows[b];
    const float* src = rows + (size_t)row_start[b] * channels;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (t < n && c < channels) ? src[(size_t)t * channels + c] : 0.f;
    }
    __syncthreads();
    float* dst = bct + (size_t)b * channels * frames;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, t = t0 + threadIdx.x;
        if (c < channels && t < frames) dst[(size_t)c * frames + t] = tile[threadIdx.x][i];
    }
}

// one warp per destination row
__global__ void segment_rows_kernel(
    const float* __restrict__ x, int channels,
    const int32_t* __restrict__ src_row, const int32_t* __restrict__ count,
    const int32_t* __restrict__ dst_row_start,
    const int32_t* __restrict__ dst_row_seq, int total_dst_rows,
    float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int groups = channels >> 2;
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         r < total_dst_rows; r += gridDim.x * (blockDim.x >> 5)) {
        const int q = __ldg(dst_row_seq + r);
        const float* src = nullptr;
        if (q >= 0) {
            int i = r - __ldg(dst_row_start + q);
            if (i < __ldg(count + q))
                src = x + (size_t)(__ldg(src_row + q) + i) * channels;
        }
        for (int g = lane; g < groups; g += 32) {
            float4 v = src ? *reinterpret_cast<const float4*>(src + 4 * g)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(y + (size_t)r * channels + 4 * g) = v;
        }
    }
}

// [rows][c_in] -> [rows][c_out], new channels zero (models wider than the 80
// log-mel features: CHANNELS=128 of the reference's hyper-parameter sweep)
__global__ void widen_rows_kernel(
    const float* __restrict__ x, int rows, int quads_in, int quads_out, float* __restrict__ y) {
    const long total = (long)rows * quads_out;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
         i += (long)gridDim.x * blockDim.x) {
        const long r = i / quads_out;
        const int q = (int)(i % quads_out);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < quads_in) v = reinterpret_cast<const float4*>(x)[r * quads_in + q];
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

}  // namespace emph

namespace emph {

namespace {
struct StagingSlot {
    void* buffer = nullptr;
    size_t capacity = 0;
    cudaEvent_t consumed = nullptr;
};
constexpr int kStagingSlots = 8;
constexpr int kStagingDevices = 16;        // an event belongs to the device it was made on
StagingSlot g_staging[kStagingDevices][kStagingSlots];
int g_staging_next[kStagingDevices] = {};
std::mutex g_staging_mutex;
}  // namespace

int staged_upload(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return EMPH_OK;
    int device = 0;
    cudaGetDevice(&device);
    device = device < 0 ? 0 : device % kStagingDevices;
    std::lock_guard<std::mutex> lock(g_staging_mutex);
    StagingSlot& slot = g_staging[device][g_staging_next[device]];
    g_staging_next[device] = (g_staging_next[device] + 1) % kStagingSlots;
    if (slot.consumed == nullptr) {
        int s = check_cuda(
            cudaEventCreateWithFlags(&slot.consumed, cudaEventDisableTiming), "staging event");
        if (s != EMPH_OK) return s;
    } else {
        // the copy that last used this slot must have been consumed by its stream
        int s = check_cuda(cudaEventSynchronize(slot.consumed), "staging wait");
        if (s != EMPH_OK) return s;
    }
    if (slot.capacity < bytes) {
        if (slot.buffer) cudaFreeHost(slot.buffer);
        slot.buffer = nullptr;
        slot.capacity = 0;
        const size_t capacity = bytes < (1u << 16) ? (1u << 16) : bytes * 2;
        int s = check_cuda(cudaHostAlloc(&slot.buffer, capacity, cudaHostAllocDefault), "staging alloc");
        if (s != EMPH_OK) return s;
        slot.capacity = capacity;
    }
    memcpy(slot.buffer, src, bytes);
    int s = check_cuda(
        cudaMemcpyAsync(dst, slot.buffer, bytes, cudaMemcpyHostToDevice, stream), "staged upload");
    if (s != EMPH_OK) return s;
    return check_cuda(cudaEventRecord(slot.consumed, stream), "staging record");
}

}  // namespace emph

extern "C" {

int emph_segment_rows(
    const float* x, int32_t channels,
    const int32_t* src_row, const int32_t* count, const int32_t* dst_row_start,
    int32_t n_seg, const int32_t* dst_row_seq, int32_t total_dst_rows,
    float* y, void* stream) {
    EMPH_REQUIRE(channels > 0 && channels % 4 == 0, "emph_segment_rows: channels %d not a multiple of 4", channels);
    EMPH_REQUIRE(n_seg >= 0 && total_dst_rows >= 0, "emph_segment_rows: negative size");
    if (total_dst_rows == 0) return EMPH_OK;
    long want = ((long)total_dst_rows + 7) / 8;
    long cap = (long)emph::sm_count() * 8;
    int grid = (int)(want < cap ? want : cap);
    emph::segment_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        x, channels, src_row, count, dst_row_start, dst_row_seq, total_dst_rows, y);
    EMPH_CHECK_LAUNCH("emph_segment_rows");
    return EMPH_OK;
}

int emph_version(void) { return 100; }

const char* emph_last_error(void) { return emph::g_error; }

int emph_device_sm_count(void) { return emph::sm_count(); }

int emph_row_index(
    const int32_t* row_start, const int32_t* n_rows, int32_t n_seq,
    int32_t* row_seq, int32_t total_rows, void* stream) {
    EMPH_REQUIRE(n_seq >= 0 && total_rows >= 0, "emph_row_index: negative size");
    if (total_rows == 0) return EMPH_OK;
    int threads = 256;
    int blocks = (total_rows + threads - 1) / threads;
    emph::row_index_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        row_start, n_rows, n_seq, row_seq, total_rows);
    EMPH_CHECK_LAUNCH("emph_row_index");
    return EMPH_OK;
}

int emph_pack_conv_weights(
    const float* conv_weight, int32_t out_channels, int32_t in_channels,
    int32_t kernel_size, float* packed, void* stream) {
    int total = out_channels * in_channels * kernel_size;
    EMPH_REQUIRE(total > 0, "emph_pack_conv_weights: empty weight");
    emph::pack_conv_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        conv_weight, out_channels, in_channels, kernel_size, packed);
    EMPH_CHECK_LAUNCH("emph_pack_conv_weights");
    return EMPH_OK;
}

int emph_pack_conv_weights_adjoint(
    const float* conv_weight, int32_t out_channels, int32_t in_channels,
    int32_t kernel_size, float* packed, void* stream) {
    int total = out_channels * in_channels * kernel_size;
    EMPH_REQUIRE(total > 0, "emph_pack_conv_weights_adjoint: empty weight");
    emph::pack_conv_weights_adjoint_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        conv_weight, out_channels, in_channels, kernel_size, packed);
    EMPH_CHECK_LAUNCH("emph_pack_conv_weights_adjoint");
    return EMPH_OK;
}

int emph_pack_rows(
    const float* bct, int32_t batch, int32_t channels, int32_t frames,
    const int32_t* row_start, const int32_t* n_rows,
    const int32_t* row_seq, int32_t total_rows, float* rows, void* stream) {
    EMPH_REQUIRE(batch > 0 && channels > 0 && frames > 0, "emph_pack_rows: empty input");
    dim3 grid((frames + 31) / 32, (channels + 31) / 32, batch), block(32, 8);
    emph::pack_rows_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        bct, channels, frames, row_start, n_rows, rows);
    EMPH_CHECK_LAUNCH("emph_pack_rows");
    emph::zero_separator_rows_kernel<<<total_rows, 32, 0, (cudaStream_t)stream>>>(
        row_seq, total_rows, channels, rows);
    EMPH_CHECK_LAUNCH("emph_pack_rows(separators)");
    return EMPH_OK;
}

int emph_widen_rows(
    const float* x, int32_t rows, int32_t channels_in, int32_t channels_out, float* y,
    void* stream) {
    EMPH_REQUIRE(rows >= 0 && channels_in > 0 && channels_out >= channels_in &&
                 channels_in % 4 == 0 && channels_out % 4 == 0,
                 "emph_widen_rows: bad shape (%d rows, %d -> %d channels)", rows,
                 channels_in, channels_out);
    if (rows == 0) return EMPH_OK;
    const long quads = (long)rows * (channels_out / 4);
    const long want = (quads + 255) / 256;
    const long cap = (long)emph::sm_count() * 16;
    emph::widen_rows_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        x, rows, channels_in / 4, channels_out / 4, y);
    EMPH_CHECK_LAUNCH("emph_widen_rows");
    return EMPH_OK;
}

int emph_unpack_rows(
    const float* rows, const int32_t* row_start, const int32_t* n_rows,
    int32_t batch, int32_t channels, int32_t frames, float* bct, void* stream) {
    EMPH_REQUIRE(batch > 0 && channels > 0 && frames > 0, "emph_unpack_rows: empty input");
    dim3 grid((frames + 31) / 32, (channels + 31) / 32, batch), block(32, 8);
    emph::unpack_rows_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        rows, row_start, n_rows, channels, frames, bct);
    EMPH_CHECK_LAUNCH("emph_unpack_rows");
    return EMPH_OK;
}

}  // extern "C"
