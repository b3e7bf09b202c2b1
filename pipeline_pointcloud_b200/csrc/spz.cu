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
#include <algorithm>

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

// ----------------------------------------------------------------------------------------------------------
// The decode direction (`splat_converter splat.spz`, spz/src/main.cpp:29-35; ConvertSpzToPly, spz_converter_api.cpp:7-30):
// unpackGaussians (load-spz.cc:373-440) fused with saveSplatToPly's row interleave (load-spz.cc:752-790).  64 B in ->
// (17 + 3*shDim)*4 B out per Gaussian; one thread per OUTPUT float, so every store is coalesced and the byte loads of a
// warp fall into a handful of lines.  Each value is the reference's sequence of single fp32 operations; std::log on the
// 256 possible alpha bytes = the fp64 logarithm rounded once (checked against glibc's logf for all 256 in
// tests/test_spz_cpu.py).
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float spz_rot_xyz(uint8_t b) { return __fadd_rn(__fmul_rn((float)b, 1.0f / 127.5f), -1.0f); }

__global__ void __launch_bounds__(256) spz_unpack_kernel(int64_t N, int sh_dim, float pos_scale, const uint8_t* __restrict__ pos,
                                                         const uint8_t* __restrict__ alpha, const uint8_t* __restrict__ col,
                                                         const uint8_t* __restrict__ scale, const uint8_t* __restrict__ rot,
                                                         const uint8_t* __restrict__ sh, float* __restrict__ rows) {
    const int D = 17 + 3 * sh_dim;
    const int64_t total = N * D;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / D;
        const int c = (int)(e - i * D);
        float v = 0.f;                                                           // nx, ny, nz
        if (c < 3) {                                                             // 24-bit fixed point, sign-extended
            const uint8_t* p = pos + 9 * i + 3 * c;
            int32_t f = (int32_t)p[0] | ((int32_t)p[1] << 8) | ((int32_t)p[2] << 16);
            f |= (f & 0x800000) ? (int32_t)0xff000000 : 0;
            v = __fmul_rn((float)f, pos_scale);
        } else if (c >= 6 && c < 9) {                                            // DC colour
            v = __fdiv_rn(__fadd_rn(__fdiv_rn((float)col[3 * i + (c - 6)], 255.0f), -0.5f), 0.15f);
        } else if (c >= 9 && c < 9 + 3 * sh_dim) {                               // f_rest: channel-major from [N, shDim, 3]
            const int t = c - 9, ch = t / sh_dim, j = t - ch * sh_dim;
            v = __fdiv_rn(__fadd_rn((float)sh[(i * sh_dim + j) * 3 + ch], -128.0f), 128.0f);
        } else if (c == 9 + 3 * sh_dim) {                                        // invSigmoid(alpha / 255)
            const float x = __fdiv_rn((float)alpha[i], 255.0f);
            v = (float)log((double)__fdiv_rn(x, __fadd_rn(1.0f, -x)));
        } else if (c >= 10 + 3 * sh_dim && c < 13 + 3 * sh_dim) {                // log-scales
            v = __fadd_rn(__fdiv_rn((float)scale[3 * i + (c - 10 - 3 * sh_dim)], 16.0f), -10.0f);
        } else if (c >= 13 + 3 * sh_dim) {                                       // rot_0 = w, rot_1..3 = x, y, z
            const int k = c - 13 - 3 * sh_dim;
            if (k > 0) {
                v = spz_rot_xyz(rot[3 * i + k - 1]);
            } else {
                const float x = spz_rot_xyz(rot[3 * i]), y = spz_rot_xyz(rot[3 * i + 1]), z = spz_rot_xyz(rot[3 * i + 2]);
                const float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
                v = __fsqrt_rn(fmaxf(0.0f, __fadd_rn(1.0f, -sq)));
            }
        }
        rows[e] = v;
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_spz_unpack(int64_t N, int32_t sh_dim, int32_t fractional_bits, const uint8_t* positions, const uint8_t* alphas,
                              const uint8_t* colors, const uint8_t* scales, const uint8_t* rotations, const uint8_t* sh,
                              float* ply_rows, b2s_stream_t stream) {
    B2S_REQUIRE(N >= 0 && (sh_dim == 0 || sh_dim == 3 || sh_dim == 8 || sh_dim == 15), "sh_dim must be 0, 3, 8 or 15");
    B2S_REQUIRE(fractional_bits >= 0 && fractional_bits < 31, "fractional_bits out of range");
    if (N == 0) return 0;
    B2S_REQUIRE(positions && alphas && colors && scales && rotations && ply_rows, "null array");
    B2S_REQUIRE(sh_dim == 0 || sh, "SH array missing");
    const int64_t total = N * (17 + 3 * sh_dim);
    const float pos_scale = (float)(1.0 / (double)(1 << fractional_bits));       // load-spz.cc:402
    spz_unpack_kernel<<<(unsigned)std::min<int64_t>(cdiv(total, 256), 148 * 64), 256, 0, (cudaStream_t)stream>>>(
        N, sh_dim, pos_scale, positions, alphas, colors, scales, rotations, sh, ply_rows);
    B2S_LAUNCH_CHECK();
    return 0;
}


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
