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
// streaming zero fill (128-bit stores that bypass L1): what a gradient buffer needs before sparse or atomic accumulation.
// Eight independent stores per thread and iteration; 32 CTAs of 256 threads per SM's worth of grid, the shape that writes
// 6.4 TB/s in sh_zero_kernel.  (A deliberately thin variant -- two 128-thread CTAs per SM -- was built for running the
// fill on a second stream beside compositing; every placement of that overlap measured as a net loss,
// profiles/r02_variants_c9.txt, so the fill runs on the caller's stream.)
__global__ void __launch_bounds__(256) zero_fill_kernel(float4* __restrict__ p, int64_t n4) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
#pragma unroll
        for (int k = 0; k < 8; ++k) stg_stream4(p + i + k * stride, z);
    }
    for (; i < n4; i += stride) stg_stream4(p + i, z);
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
        const int64_t grid = b2s::cdiv(n4, 256 * 8);
        b2s::zero_fill_kernel<<<(unsigned)(grid < 148 * 32 ? grid : 148 * 32), 256, 0, st>>>(reinterpret_cast<float4*>(c + head), n4);
        B2S_LAUNCH_CHECK();
    }
    if (tail) B2S_CUDA(cudaMemsetAsync(c + head + n4 * 16, 0, (size_t)tail, st));
    return 0;
}

extern "C" {
int b2s_version(void) { return B2S_VERSION; }
const char* b2s_last_error(void) { return b2s::g_err; }
}
