// Polyphase windowed-sinc resampler, sm_100a.  Replaces emphases.resample
// (emphases/core.py:613-619 -> torchaudio.transforms.Resample, i.e. a strided
// conv1d with `new_freq` filters of 2 * width + orig_freq taps):
//   y[q * new + p] = sum_k kernel[p][k] * xpad[q * orig + k],
//   xpad = zeros(width) ++ x ++ zeros(width + orig)
// The filter bank is built on the host exactly as torchaudio builds it
// (emphases_b200/resampling.py) and passed in.  One thread per output sample;
// a block's outputs share their input window through L1.
#include "common.cuh"

namespace emph {

__global__ void __launch_bounds__(256)
resample_kernel(
    const float* __restrict__ x, long long length, const float* __restrict__ kernel,
    int orig, int fresh, int width, int taps, float* __restrict__ y, long long target) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= target) return;
    const long long q = n / fresh;
    const int p = (int)(n - q * fresh);
    const float* w = kernel + (size_t)p * taps;
    const long long first = q * orig - width;          // x index of tap 0
    float acc = 0.f;
    int k0 = first < 0 ? (int)(-first) : 0;
    long long k1 = length - first;                      // taps with x index < length
    if (k1 > taps) k1 = taps;
    for (int k = k0; k < (int)k1; ++k) acc = fmaf(__ldg(w + k), __ldg(x + first + k), acc);
    y[n] = acc;
}

// The same filter over a whole packed int16 corpus in one launch: output
// sample n of the packed 16 kHz buffer belongs to the utterance found by a
// binary search over out_off (ascending); samples are x / 32768 exactly as
// torchaudio.load returns them.  Gaps between utterances are zeroed.
template <typename T>
__device__ __forceinline__ float packed_sample(const T* p);
template <>
__device__ __forceinline__ float packed_sample<int16_t>(const int16_t* p) {
    return (float)__ldg(p) * (1.f / 32768.f);
}
template <>
__device__ __forceinline__ float packed_sample<float>(const float* p) { return __ldg(p); }

template <typename T>
__global__ void __launch_bounds__(256)
resample_packed_kernel(
    const T* __restrict__ x, const int64_t* __restrict__ in_off,
    const int64_t* __restrict__ in_len, const int64_t* __restrict__ out_off,
    const int64_t* __restrict__ out_len, int n_utterances,
    const float* __restrict__ kernel, int orig, int fresh, int width, int taps,
    float* __restrict__ y, long long total) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= total) return;
    int lo = 0, hi = n_utterances;               // last u with out_off[u] <= n
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(out_off + mid) <= n) lo = mid; else hi = mid;
    }
    const long long local = n - __ldg(out_off + lo);
    if (local < 0 || local >= __ldg(out_len + lo)) {
        y[n] = 0.f;
        return;
    }
    const T* src = x + __ldg(in_off + lo);
    const long long length = __ldg(in_len + lo);
    const long long q = local / fresh;
    const int p = (int)(local - q * fresh);
    const float* w = kernel + (size_t)p * taps;
    const long long first = q * orig - width;
    float acc = 0.f;
    int k0 = first < 0 ? (int)(-first) : 0;
    long long k1 = length - first;
    if (k1 > taps) k1 = taps;
    for (int k = k0; k < (int)k1; ++k)
        acc = fmaf(__ldg(w + k), packed_sample<T>(src + first + k), acc);
    y[n] = acc;
}

}  // namespace emph

namespace emph {
template <typename T>
int launch_resample_packed(
    const T* x, const int64_t* in_off, const int64_t* in_len,
    const int64_t* out_off, const int64_t* out_len, int32_t n_utterances,
    const float* kernel, int32_t orig_freq, int32_t new_freq, int32_t width,
    float* y, int64_t total_out, void* stream) {
    EMPH_REQUIRE(orig_freq > 0 && new_freq > 0 && width >= 0, "emph_resample_packed: bad filter");
    EMPH_REQUIRE(n_utterances >= 0 && total_out >= 0, "emph_resample_packed: negative size");
    if (total_out == 0 || n_utterances == 0) return EMPH_OK;
    const int taps = 2 * width + orig_freq;
    const long long blocks = (total_out + 255) / 256;
    resample_packed_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        x, in_off, in_len, out_off, out_len, n_utterances, kernel, orig_freq, new_freq,
        width, taps, y, total_out);
    EMPH_CHECK_LAUNCH("emph_resample_packed");
    return EMPH_OK;
}
}  // namespace emph

extern "C" int emph_resample_packed_i16(
    const int16_t* x, const int64_t* in_off, const int64_t* in_len,
    const int64_t* out_off, const int64_t* out_len, int32_t n_utterances,
    const float* kernel, int32_t orig_freq, int32_t new_freq, int32_t width,
    float* y, int64_t total_out, void* stream) {
    return emph::launch_resample_packed<int16_t>(
        x, in_off, in_len, out_off, out_len, n_utterances, kernel, orig_freq, new_freq, width,
        y, total_out, stream);
}

extern "C" int emph_resample_packed_f32(
    const float* x, const int64_t* in_off, const int64_t* in_len,
    const int64_t* out_off, const int64_t* out_len, int32_t n_utterances,
    const float* kernel, int32_t orig_freq, int32_t new_freq, int32_t width,
    float* y, int64_t total_out, void* stream) {
    return emph::launch_resample_packed<float>(
        x, in_off, in_len, out_off, out_len, n_utterances, kernel, orig_freq, new_freq, width,
        y, total_out, stream);
}

extern "C" int emph_resample_f32(
    const float* x, int64_t length, const float* kernel, int32_t orig_freq,
    int32_t new_freq, int32_t width, float* y, int64_t target_length, void* stream) {
    EMPH_REQUIRE(orig_freq > 0 && new_freq > 0 && width >= 0, "emph_resample_f32: bad filter");
    EMPH_REQUIRE(length >= 0 && target_length >= 0, "emph_resample_f32: negative length");
    if (target_length == 0) return EMPH_OK;
    const int taps = 2 * width + orig_freq;
    const long long blocks = (target_length + 255) / 256;
    emph::resample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        x, length, kernel, orig_freq, new_freq, width, taps, y, target_length);
    EMPH_CHECK_LAUNCH("emph_resample_f32");
    return EMPH_OK;
}
