// Pieces shared by the stage-1 kernels (project.cu: projection + SH colour; route.cu: the same fused with the multi-GPU
// exchange): the shared-memory coefficient tile, camera setup, and the tensor-memory-accelerator plumbing.
#pragma once
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#include <stdlib.h>

#include "b2s_math.cuh"
#include "common.cuh"

namespace b2s {

constexpr int FWD_THREADS = 256;

// Row stride of the smem SH tile (floats).  Degree 3 (48 floats per row): 52, i.e. 16-byte aligned rows whose per-thread
// 128-bit reads are conflict free (8 consecutive rows start in banks 0,20,8,28,16,4,24,12: disjoint 4-bank groups), so the
// tile is filled, read and drained with LDS.128 / STS.128.  Other degrees: odd stride, conflict-free scalar rows.
__host__ __device__ constexpr int sh_row_stride(int K) { return K == 16 ? 52 : ((3 * K) | 1); }

// ---- warp-cooperative movement of 32 SH rows between global memory and the padded smem tile -------
// gbase points at row 0 of the warp's chunk, `rows` valid rows, row_mask: which rows to move.
template <int KVEC>  // 16: 3*K == 48 floats and stride_k == K -> 128-bit path; 0: generic
__device__ __forceinline__ void sh_rows_load(float* tile, const float* gbase, int rows, unsigned row_mask, int K,
                                             int stride_k, int lane) {
    const int rs = sh_row_stride(K);
    if (KVEC == 16) {
        const float4* g4 = reinterpret_cast<const float4*>(gbase);
        float4 v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            int q = lane + 32 * i;
            int r = q / 12;
            bool on = (r < rows) && ((row_mask >> r) & 1u);
            v[i] = on ? ldg_stream4(g4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            int q = lane + 32 * i;
            int r = q / 12, e = (q - r * 12) * 4;
            *reinterpret_cast<float4*>(tile + r * rs + e) = v[i];
        }
    } else {
        const int per = 3 * K;
        for (int f = lane; f < 32 * per; f += 32) {
            int r = f / per, e = f - r * per;
            bool on = (r < rows) && ((row_mask >> r) & 1u);
            tile[r * rs + e] = on ? ldg_stream1(gbase + (size_t)r * stride_k * 3 + e) : 0.f;
        }
    }
}

template <int KVEC>
__device__ __forceinline__ void sh_rows_store(const float* tile, float* gbase, int rows, int K, int stride_k, int lane,
                                              bool accumulate) {
    const int rs = sh_row_stride(K);
    if (KVEC == 16 && !accumulate) {
        float4* g4 = reinterpret_cast<float4*>(gbase);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            int q = lane + 32 * i;
            int r = q / 12, e = (q - r * 12) * 4;
            if (r < rows) stg_stream4(g4 + q, *reinterpret_cast<const float4*>(tile + r * rs + e));
        }
    } else {
        const int per = 3 * K;
        for (int f = lane; f < 32 * per; f += 32) {
            int r = f / per, e = f - r * per;
            if (r < rows) {
                float* g = gbase + (size_t)r * stride_k * 3 + e;
                *g = accumulate ? (*g + tile[r * rs + e]) : tile[r * rs + e];
            }
        }
    }
}

__device__ __forceinline__ void load_cameras(Camera* cams, int C, const float* viewmats, const float* Ks,
                                             const float* campos, float W, float H) {
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        camera_setup(cams[c], viewmats + 16 * c, Ks + 9 * c, campos ? campos + 3 * c : nullptr, W, H);
    __syncthreads();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();      // a copy that never completes is a bug: fail, do not hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}

constexpr int TMA_A_BYTES = FWD_THREADS * 128, TMA_B_BYTES = FWD_THREADS * 64;   // 32 KB + 16 KB per CTA

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda: the library must load on a box without a driver)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn tensor_map_encoder() {
    const char* env = getenv("B2S_PROJECT_TMA");
    if (env && env[0] == '0') return nullptr;                // A/B switch: the LDG/STS kernel
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// the two boxes of the [N,48] fp32 coefficient tensor: floats [0,32) (128-byte swizzle) and [32,48) (64-byte swizzle) of 256 rows
static inline bool sh_tensor_maps(const float* coeffs, int64_t N, CUtensorMap* map_a, CUtensorMap* map_b) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || N <= 0 || N >= (1LL << 31)) return false;
    const cuuint64_t gdim[2] = {48, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {192};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box_a[2] = {32, (cuuint32_t)FWD_THREADS}, box_b[2] = {16, (cuuint32_t)FWD_THREADS};
    void* g = const_cast<float*>(coeffs);
    if (enc(map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstride, box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    return enc(map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstride, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace b2s
