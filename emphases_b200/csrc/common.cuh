// Shared helpers for the emphases_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/emphases_b200.h"

namespace emph {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t status, const char* what) {
    if (status != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(status));
        return EMPH_ECUDA;
    }
    return EMPH_OK;
}

#define EMPH_CHECK_LAUNCH(what)                                  \
    do {                                                         \
        int _s = ::emph::check_cuda(cudaGetLastError(), what);   \
        if (_s != EMPH_OK) return _s;                            \
    } while (0)

#define EMPH_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            ::emph::set_error(__VA_ARGS__);  \
            return EMPH_EINVAL;              \
        }                                    \
    } while (0)

// Host -> device copy of a small, freshly built host array through a ring of
// page-locked staging buffers (api.cu): truly asynchronous (a cudaMemcpyAsync
// from pageable memory synchronises the stream first), and `src` may be reused
// as soon as the call returns.
int staged_upload(void* dst, const void* src, size_t bytes, cudaStream_t stream);

inline int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int device = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, device);
        if (cached <= 0) cached = 148;
    }
    return cached;
}

__device__ __forceinline__ float apply_activation(float v, int act) {
    switch (act) {
        case EMPH_ACT_RELU: return fmaxf(v, 0.f);
        case EMPH_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
        case EMPH_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
        case EMPH_ACT_SILU: return v / (1.f + expf(-v));
        default: return v;
    }
}

}  // namespace emph
