// Shared helpers for the b2s kernels (error plumbing, vector load/store, warp utilities).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2s.h"

namespace b2s {

void set_error(const char* fmt, ...);  // api.cu: thread-local message for b2s_last_error()

#define B2S_REQUIRE(cond, msg)                                            \
    do {                                                                  \
        if (!(cond)) {                                                    \
            ::b2s::set_error("%s: %s", __func__, msg);                    \
            return B2S_E_BADARG;                                          \
        }                                                                 \
    } while (0)

#define B2S_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::b2s::set_error("%s: %s -> %s", __func__, #expr, cudaGetErrorString(_e));         \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

#define B2S_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            ::b2s::set_error("%s: kernel launch -> %s", __func__, cudaGetErrorString(_e));     \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

// streaming 128-bit load that does not pollute L1 (data read exactly once)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(FULL, v, 16);
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    return v;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b2s
