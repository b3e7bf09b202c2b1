// Sparse gradient synchronisation for camera data-parallel training (BASELINE config 4: replicated
// Gaussians, every rank renders its own camera, gradients summed over ranks).  gsplat / the reference's trainer would
// all-reduce the dense gradients -- 236 B per Gaussian, 1.37 GB at 5.8 M.  In a scene that saturates, compositing reaches
// a few percent of the Gaussians of a view, so almost all of that volume is zeros.  Here every rank compacts the rows
// that carry a gradient into records
//     [row (int bits), v_means 3, v_quats 4, v_scales 3, v_opacity 1, v_sh 3 K]
// which are all-gathered (a few tens of MB) and added into the dense gradients of every rank:
//   b2s_grad_compact     : one pass over the small gradients (a row is live when any of v_means / v_quats / v_scales /
//                          v_opacity / its first SH gradient row is non-zero: the degree-0 basis is a non-zero constant,
//                          so a live colour cotangent always shows there), ONE atomic per CTA for the base, records
//                          written up to `cap`; total[0] = number of live rows (may exceed cap: the caller then falls
//                          back to the dense all-reduce)
//   b2s_grad_scatter_add : records -> atomic adds into the dense gradients (two ranks may report the same row)
#include "common.cuh"

namespace b2s {

constexpr int GS_THREADS = 256;

__global__ void __launch_bounds__(GS_THREADS) grad_compact_kernel(int64_t N, int sh_floats, const float* __restrict__ v_means,
                                                                  const float* __restrict__ v_quats, const float* __restrict__ v_scales,
                                                                  const float* __restrict__ v_opac, const float* __restrict__ v_sh,
                                                                  unsigned long long* __restrict__ total, float* __restrict__ records,
                                                                  int64_t cap) {
    __shared__ unsigned s_w[GS_THREADS / 32];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * GS_THREADS + threadIdx.x;
    bool live = false;
    if (n < N) {
        live |= (v_means[3 * n] != 0.f) | (v_means[3 * n + 1] != 0.f) | (v_means[3 * n + 2] != 0.f);
        const float4 q = reinterpret_cast<const float4*>(v_quats)[n];
        live |= (q.x != 0.f) | (q.y != 0.f) | (q.z != 0.f) | (q.w != 0.f);
        live |= (v_scales[3 * n] != 0.f) | (v_scales[3 * n + 1] != 0.f) | (v_scales[3 * n + 2] != 0.f);
        if (v_opac) live |= v_opac[n] != 0.f;
        if (v_sh) { const float* s = v_sh + n * sh_floats; live |= (s[0] != 0.f) | (s[1] != 0.f) | (s[2] != 0.f); }
    }
    const unsigned b = __ballot_sync(FULL, live);
    if (lane == 0) s_w[warp] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < GS_THREADS / 32; ++w) { const unsigned x = s_w[w]; s_w[w] = t; t += x; }
        s_base = t ? atomicAdd(total, (unsigned long long)t) : 0ull;
    }
    __syncthreads();
    if (!live) return;
    const int64_t slot = (int64_t)s_base + s_w[warp] + __popc(b & ((1u << lane) - 1u));
    if (slot >= cap) return;
    const int R = 12 + sh_floats;
    float* o = records + slot * R;
    o[0] = __int_as_float((int)n);
    o[1] = v_means[3 * n]; o[2] = v_means[3 * n + 1]; o[3] = v_means[3 * n + 2];
    const float4 q = reinterpret_cast<const float4*>(v_quats)[n];
    o[4] = q.x; o[5] = q.y; o[6] = q.z; o[7] = q.w;
    o[8] = v_scales[3 * n]; o[9] = v_scales[3 * n + 1]; o[10] = v_scales[3 * n + 2];
    o[11] = v_opac ? v_opac[n] : 0.f;
    for (int k = 0; k < sh_floats; ++k) o[12 + k] = v_sh[n * sh_floats + k];
}

// one warp per record: the 12 + 3K values of a row are added by consecutive lanes (coalesced record reads, and the
// atomics of a row land in a handful of cache lines)
__global__ void __launch_bounds__(GS_THREADS) grad_scatter_add_kernel(int64_t n_records, int sh_floats, const float* __restrict__ records,
                                                                      float* __restrict__ v_means, float* __restrict__ v_quats,
                                                                      float* __restrict__ v_scales, float* __restrict__ v_opac,
                                                                      float* __restrict__ v_sh) {
    const int lane = threadIdx.x & 31;
    const int R = 12 + sh_floats;
    for (int64_t r = (int64_t)blockIdx.x * (GS_THREADS / 32) + (threadIdx.x >> 5); r < n_records; r += (int64_t)gridDim.x * (GS_THREADS / 32)) {
        const float* p = records + r * R;
        const int64_t n = (int64_t)__float_as_int(p[0]);
        for (int k = 1 + lane; k < R; k += 32) {
            const float v = p[k];
            if (v == 0.f) continue;
            float* dst;
            if (k < 4) dst = v_means + 3 * n + (k - 1);
            else if (k < 8) dst = v_quats + 4 * n + (k - 4);
            else if (k < 11) dst = v_scales + 3 * n + (k - 8);
            else if (k == 11) dst = v_opac ? v_opac + n : nullptr;
            else dst = v_sh + n * sh_floats + (k - 12);
            if (dst) atomicAdd(dst, v);
        }
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_grad_compact(int64_t N, int32_t sh_floats, const float* v_means, const float* v_quats, const float* v_scales,
                                const float* v_opacities, const float* v_sh, int64_t* total, float* records, int64_t cap,
                                b2s_stream_t stream) {
    B2S_REQUIRE(N >= 0 && sh_floats >= 0 && cap >= 0 && total != nullptr, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    B2S_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
    if (N == 0) return 0;
    B2S_REQUIRE(v_means && v_quats && v_scales && (sh_floats == 0 || v_sh) && (cap == 0 || records), "null pointer");
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(v_quats) & 15u) == 0, "v_quats misaligned");
    grad_compact_kernel<<<(unsigned)cdiv(N, GS_THREADS), GS_THREADS, 0, st>>>(N, sh_floats, v_means, v_quats, v_scales, v_opacities, v_sh,
                                                                             reinterpret_cast<unsigned long long*>(total), records, cap);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_grad_scatter_add(int64_t n_records, int32_t sh_floats, const float* records, float* v_means, float* v_quats,
                                    float* v_scales, float* v_opacities, float* v_sh, b2s_stream_t stream) {
    B2S_REQUIRE(n_records >= 0 && sh_floats >= 0, "bad args");
    if (n_records == 0) return 0;
    B2S_REQUIRE(records && v_means && v_quats && v_scales && (sh_floats == 0 || v_sh), "null pointer");
    const unsigned grid = (unsigned)min((int64_t)148 * 16, cdiv(n_records, GS_THREADS / 32));
    grad_scatter_add_kernel<<<grid, GS_THREADS, 0, (cudaStream_t)stream>>>(n_records, sh_floats, records, v_means, v_quats, v_scales,
                                                                          v_opacities, v_sh);
    B2S_LAUNCH_CHECK();
    return 0;
}
