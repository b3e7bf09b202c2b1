// Library-level entry points: version and the thread-local error string.
#include <stdarg.h>

#include "common.cuh"

namespace b2s {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b2s

namespace b2s {
// streaming zero fill (128-bit stores that bypass L1): what a gradient buffer needs before sparse or atomic accumulation
__global__ void __launch_bounds__(256) zero_fill_kernel(float4* __restrict__ p, int64_t n4) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) stg_stream4(p + i, z);
}
}  // namespace b2s

extern "C" int b2s_zero_fill(void* p, int64_t bytes, b2s_stream_t stream) {
    B2S_REQUIRE(bytes >= 0 && (p != nullptr || bytes == 0), "bad args");
    if (bytes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    char* c = reinterpret_cast<char*>(p);
    const int64_t head = (16 - (reinterpret_cast<uintptr_t>(c) & 15u)) & 15u;      // unaligned edges: plain memsets
    if (head >= bytes) { B2S_CUDA(cudaMemsetAsync(c, 0, (size_t)bytes, st)); return 0; }
    if (head) B2S_CUDA(cudaMemsetAsync(c, 0, (size_t)head, st));
    const int64_t n4 = (bytes - head) / 16, tail = bytes - head - n4 * 16;
    if (n4) {
        b2s::zero_fill_kernel<<<(unsigned)b2s::cdiv(n4 < 148 * 32 * 256 ? n4 : 148 * 32 * 256, 256), 256, 0, st>>>(
            reinterpret_cast<float4*>(c + head), n4);
        B2S_LAUNCH_CHECK();
    }
    if (tail) B2S_CUDA(cudaMemsetAsync(c + head + n4 * 16, 0, (size_t)tail, st));
    return 0;
}

extern "C" {
int b2s_version(void) { return B2S_VERSION; }
const char* b2s_last_error(void) { return b2s::g_err; }
}
