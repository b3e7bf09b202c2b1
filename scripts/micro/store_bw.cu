// Micro-benchmark: write-only bandwidth of different global store flavours (1.11 GB, the size of the v_sh gradient).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a store_bw.cu -o store_bw
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ void st4(float4* p, float4 v) {
    if (MODE == 0) *p = v;
    else if (MODE == 1) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (MODE == 2) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else if (MODE == 3) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int MODE>
__global__ void fill_stride(float4* o, long n) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) st4<MODE>(o + i, z);
}
template <int MODE>
__global__ void fill_flat(float4* o, long n) {   // one CTA per 256*4 float4, 4 stores per thread
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    long base = (long)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k * 256 < n) st4<MODE>(o + base + k * 256, z);
}
// the pattern of sh_rows_store: a warp writes a contiguous 6 KB chunk, lane + 32 i
template <int MODE>
__global__ void fill_warpchunk(float4* o, long n_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long row0 = ((long)blockIdx.x * 8 + warp) * 32;
    if (row0 >= n_rows) return;
    float4* g = o + row0 * 12;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 12; ++i) st4<MODE>(g + lane + 32 * i, z);
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main() {
    const long rows = 5800000, n4 = rows * 12;       // 48 floats per row
    const double gb = n4 * 16 / 1e9;
    float4* buf; cudaMalloc(&buf, n4 * 16);
    float ms;
    ms = timeit([&] { cudaMemsetAsync(buf, 0, n4 * 16); }); printf("cudaMemsetAsync            %.3f ms %.0f GB/s\n", ms, gb / ms * 1e3);
#define RUN(NAME, KERNEL, GRID, BLOCK, ARG) ms = timeit([&] { KERNEL<<<GRID, BLOCK>>>(buf, ARG); }); printf("%-26s %.3f ms %.0f GB/s\n", NAME, ms, gb / ms * 1e3);
    RUN("stride plain 1184x256", fill_stride<0>, 148 * 8, 256, n4)
    RUN("stride no_allocate", fill_stride<1>, 148 * 8, 256, n4)
    RUN("stride .cs", fill_stride<2>, 148 * 8, 256, n4)
    RUN("stride .wt", fill_stride<3>, 148 * 8, 256, n4)
    RUN("stride plain 148x1024", fill_stride<0>, 148, 1024, n4)
    RUN("stride plain 4736x256", fill_stride<0>, 148 * 32, 256, n4)
    RUN("flat plain", fill_flat<0>, (unsigned)((n4 + 1023) / 1024), 256, n4)
    RUN("flat no_allocate", fill_flat<1>, (unsigned)((n4 + 1023) / 1024), 256, n4)
    RUN("flat .cs", fill_flat<2>, (unsigned)((n4 + 1023) / 1024), 256, n4)
    RUN("warpchunk plain", fill_warpchunk<0>, (unsigned)((rows + 255) / 256), 256, rows)
    RUN("warpchunk no_allocate", fill_warpchunk<1>, (unsigned)((rows + 255) / 256), 256, rows)
    RUN("warpchunk .cs", fill_warpchunk<2>, (unsigned)((rows + 255) / 256), 256, rows)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
