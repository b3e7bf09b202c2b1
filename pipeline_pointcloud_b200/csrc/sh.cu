// Stand-alone stage-level ops: SH colour fwd/bwd over arbitrary directions (gsplat
// `spherical_harmonics`) and `quat_scale_to_covar_preci`.
// The training path uses the fused kernels in project.cu; these exist for API parity.
#include "b2s_math.cuh"
#include "common.cuh"

namespace b2s {

__global__ void sh_fwd_kernel(int64_t M, int deg, int stride_k, const float* __restrict__ dirs,
                              const float* __restrict__ coeffs, const uint8_t* __restrict__ masks,
                              float* __restrict__ colors) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float r = 0.f, g = 0.f, b = 0.f;
    if (!masks || masks[i]) {
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        float Y[16];
        sh_basis(deg, x * inv, y * inv, z * inv, Y);
        const int K = (deg + 1) * (deg + 1);
        const float* c = coeffs + (size_t)i * stride_k * 3;
        for (int k = 0; k < K; ++k) { r += Y[k] * c[3 * k]; g += Y[k] * c[3 * k + 1]; b += Y[k] * c[3 * k + 2]; }
    }
    colors[3 * i] = r; colors[3 * i + 1] = g; colors[3 * i + 2] = b;
}

__global__ void sh_bwd_kernel(int64_t M, int deg, int stride_k, const float* __restrict__ dirs,
                              const float* __restrict__ coeffs, const uint8_t* __restrict__ masks,
                              const float* __restrict__ v_colors, float* __restrict__ v_coeffs,
                              float* __restrict__ v_dirs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int K = (deg + 1) * (deg + 1);
    float* vc = v_coeffs + (size_t)i * stride_k * 3;
    if (masks && !masks[i]) {
        for (int k = 0; k < 3 * K; ++k) vc[k] = 0.f;
        if (v_dirs) { v_dirs[3 * i] = 0.f; v_dirs[3 * i + 1] = 0.f; v_dirs[3 * i + 2] = 0.f; }
        return;
    }
    float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    float ux = x * inv, uy = y * inv, uz = z * inv;
    float Y[16], w[16];
    sh_basis(deg, ux, uy, uz, Y);
    const float vr = v_colors[3 * i], vg = v_colors[3 * i + 1], vb = v_colors[3 * i + 2];
    const float* c = coeffs + (size_t)i * stride_k * 3;
    for (int k = 0; k < K; ++k) {
        vc[3 * k] = Y[k] * vr; vc[3 * k + 1] = Y[k] * vg; vc[3 * k + 2] = Y[k] * vb;
        w[k] = c[3 * k] * vr + c[3 * k + 1] * vg + c[3 * k + 2] * vb;
    }
    if (v_dirs) {
        float vx, vy, vz;
        sh_basis_vjp(deg, ux, uy, uz, w, vx, vy, vz);
        float dp = ux * vx + uy * vy + uz * vz;
        v_dirs[3 * i] = (vx - ux * dp) * inv; v_dirs[3 * i + 1] = (vy - uy * dp) * inv; v_dirs[3 * i + 2] = (vz - uz * dp) * inv;
    }
}

__global__ void covar_preci_kernel(int64_t N, const float* __restrict__ quats, const float* __restrict__ scales,
                                   float* __restrict__ covars, float* __restrict__ precis) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float q[4] = {quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]};
    float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    Rot R; float qn[4], nrm, M[9]; Sym3 S;
    quat_to_rotmat(q, R, qn, nrm);
    if (covars) {
        covar_from_rot_scale(R, s, M, S);
        float* o = covars + 9 * i;
        o[0] = S.c00; o[1] = S.c01; o[2] = S.c02; o[3] = S.c01; o[4] = S.c11; o[5] = S.c12; o[6] = S.c02; o[7] = S.c12; o[8] = S.c22;
    }
    if (precis) {
        float is[3] = {1.f / s[0], 1.f / s[1], 1.f / s[2]};
        covar_from_rot_scale(R, is, M, S);
        float* o = precis + 9 * i;
        o[0] = S.c00; o[1] = S.c01; o[2] = S.c02; o[3] = S.c01; o[4] = S.c11; o[5] = S.c12; o[6] = S.c02; o[7] = S.c12; o[8] = S.c22;
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_sh_fwd(int64_t M, int32_t degree, int32_t stride_k, const float* dirs, const float* coeffs,
                          const uint8_t* masks, float* colors, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && degree >= 0 && degree <= 3, "bad sizes (degree 0..3)");
    B2S_REQUIRE(stride_k >= (degree + 1) * (degree + 1), "stride_k < (deg+1)^2");
    if (M == 0) return 0;
    B2S_REQUIRE(dirs && coeffs && colors, "null pointer");
    sh_fwd_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(M, degree, stride_k, dirs, coeffs, masks, colors);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_sh_bwd(int64_t M, int32_t degree, int32_t stride_k, const float* dirs, const float* coeffs,
                          const uint8_t* masks, const float* v_colors, float* v_coeffs, float* v_dirs,
                          b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && degree >= 0 && degree <= 3, "bad sizes (degree 0..3)");
    B2S_REQUIRE(stride_k >= (degree + 1) * (degree + 1), "stride_k < (deg+1)^2");
    if (M == 0) return 0;
    B2S_REQUIRE(dirs && coeffs && v_colors && v_coeffs, "null pointer");
    sh_bwd_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(M, degree, stride_k, dirs, coeffs, masks,
                                                                           v_colors, v_coeffs, v_dirs);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_quat_scale_to_covar_preci(int64_t N, const float* quats, const float* scales, float* covars,
                                             float* precis, b2s_stream_t stream) {
    B2S_REQUIRE(N >= 0, "bad size");
    if (N == 0) return 0;
    B2S_REQUIRE(quats && scales, "null pointer");
    covar_preci_kernel<<<(unsigned)cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(N, quats, scales, covars, precis);
    B2S_LAUNCH_CHECK();
    return 0;
}
