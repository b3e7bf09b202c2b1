// SPZ packing on the GPU: the quantisation half of the reference's Spz-Export component
// (`splat_converter splat.ply`, source/container/src/main.py:1532-1545), i.e. packGaussians() of the vendored Niantic
// codec (source/container/src/pipeline/post_processing/spz/src/cc/load-spz.cc:213-289): 24-bit fixed-point positions
// (12 fractional bits), 8-bit log-scales, 3 x 8-bit quaternion xyz (w made non-negative), 8-bit sigmoid(alpha), 8-bit
// wide-range DC colour, SH quantised to 5 bits (degree 1) / 4 bits (rest) stored in 8.  236 B -> 64 B per Gaussian,
// pure HBM streaming.  The gzip container around the packed arrays is written on the host (export.py).
//
// Byte-exact against the reference's own converter (oracle/_ref/splat_converter): every step is a single-rounding fp32
// operation in the reference's order (no FMA contraction, IEEE sqrt / division), std::round = roundf, and exp() is
// evaluated in fp64 and rounded once, which is what a correctly rounded expf returns.
#include "common.cuh"

namespace b2s {

__device__ __forceinline__ uint8_t spz_u8(float x) {            // load-spz.cc:74 toUint8
    return (uint8_t)fminf(fmaxf(roundf(x), 0.f), 255.f);
}
__device__ __forceinline__ uint8_t spz_sh(float x, int bucket) {  // load-spz.cc:77-81 quantizeSH
    int q = (int)__fadd_rn(roundf(__fmul_rn(x, 128.f)), 128.f);
    q = (q + bucket / 2) / bucket * bucket;                       // C division: truncates toward zero, as the reference
    return (uint8_t)min(max(q, 0), 255);
}

// one thread per Gaussian for the small fields, then the SH block cooperatively (coalesced byte stores)
__global__ void __launch_bounds__(256) spz_pack_kernel(int64_t N, int sh_dim, const float* __restrict__ means,
                                                       const float* __restrict__ scales, const float* __restrict__ quats,
                                                       const float* __restrict__ opac, const float* __restrict__ sh0,
                                                       const float* __restrict__ shN, int shN_stride, uint8_t* __restrict__ o_pos,
                                                       uint8_t* __restrict__ o_alpha, uint8_t* __restrict__ o_col,
                                                       uint8_t* __restrict__ o_scale, uint8_t* __restrict__ o_rot,
                                                       uint8_t* __restrict__ o_sh) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int32_t fx = (int32_t)roundf(__fmul_rn(means[3 * i + k], 4096.f));
            o_pos[9 * i + 3 * k] = fx & 0xff; o_pos[9 * i + 3 * k + 1] = (fx >> 8) & 0xff; o_pos[9 * i + 3 * k + 2] = (fx >> 16) & 0xff;
            o_scale[3 * i + k] = spz_u8(__fmul_rn(__fadd_rn(scales[3 * i + k], 10.f), 16.f));
            o_col[3 * i + k] = spz_u8(__fadd_rn(__fmul_rn(sh0[3 * i + k], 0.15f * 255.0f), 0.5f * 255.0f));
        }
        // checkpoint order is (w, x, y, z); the codec works on (x, y, z, w)   (load-spz.cc:675: rot_1, rot_2, rot_3, rot_0)
        const float w = quats[4 * i], x = quats[4 * i + 1], y = quats[4 * i + 2], z = quats[4 * i + 3];
        const float n2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)), __fmul_rn(w, w));
        const float n = __fsqrt_rn(n2);
        const float qx = __fdiv_rn(x, n), qy = __fdiv_rn(y, n), qz = __fdiv_rn(z, n), qw = __fdiv_rn(w, n);
        const float s = qw < 0.f ? -127.5f : 127.5f;
        o_rot[3 * i] = spz_u8(__fadd_rn(__fmul_rn(qx, s), 127.5f));
        o_rot[3 * i + 1] = spz_u8(__fadd_rn(__fmul_rn(qy, s), 127.5f));
        o_rot[3 * i + 2] = spz_u8(__fadd_rn(__fmul_rn(qz, s), 127.5f));
        const float e = (float)exp(-(double)opac[i]);
        o_alpha[i] = spz_u8(__fmul_rn(__fdiv_rn(1.f, __fadd_rn(1.f, e)), 255.f));
    }
    if (sh_dim > 0) {
        const int per = sh_dim * 3;
        const int64_t base = (int64_t)blockIdx.x * blockDim.x;
        const int64_t rows = min((int64_t)blockDim.x, N - base);
        for (int64_t t = threadIdx.x; t < rows * per; t += blockDim.x) {
            const int64_t r = t / per;
            const int e = (int)(t - r * per);
            o_sh[base * per + t] = spz_sh(shN[(base + r) * shN_stride + e], e < 9 ? 8 : 16);
        }
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_spz_pack(int64_t N, int32_t sh_dim, const float* means, const float* log_scales, const float* quats_wxyz,
                            const float* opacity_logits, const float* sh0, const float* shN, int32_t shN_rows,
                            uint8_t* positions, uint8_t* alphas, uint8_t* colors, uint8_t* scales, uint8_t* rotations,
                            uint8_t* sh, b2s_stream_t stream) {
    B2S_REQUIRE(N >= 0 && (sh_dim == 0 || sh_dim == 3 || sh_dim == 8 || sh_dim == 15), "sh_dim must be 0, 3, 8 or 15");
    if (N == 0) return 0;
    B2S_REQUIRE(means && log_scales && quats_wxyz && opacity_logits && sh0, "null input");
    B2S_REQUIRE(positions && alphas && colors && scales && rotations, "null output");
    B2S_REQUIRE(sh_dim == 0 || (shN && sh && shN_rows >= sh_dim), "SH arrays missing");
    spz_pack_kernel<<<(unsigned)cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(N, sh_dim, means, log_scales, quats_wxyz,
                                                                             opacity_logits, sh0, shN, shN_rows * 3, positions,
                                                                             alphas, colors, scales, rotations, sh);
    B2S_LAUNCH_CHECK();
    return 0;
}
