// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace emph {
int conv_stack_bf16_tc(
    const float*, const int32_t*, int32_t, const float*, const float*, const int32_t*,
    int32_t, int32_t, int32_t, float*, cudaStream_t) {
    set_error("emph_conv_stack: bf16 tensor-core mode not built yet");
    return EMPH_ENOSYS;
}
}  // namespace emph
