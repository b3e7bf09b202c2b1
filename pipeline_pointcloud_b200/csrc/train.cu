// The step either side of rasterization() in the training loop: the photometric loss that
// produces the render cotangent, and the optimizer update that consumes the gradients.
//
// Reference anchors: the Train-Stage1 component runs gsplat's `examples/simple_trainer.py`
// (source/container/src/main.py:1328-1347), whose loss is  (1 - lambda) * L1 + lambda * (1 - SSIM)  with
// lambda = 0.2 (SSIM: 11x11 Gaussian window, sigma 1.5, "valid" padding, C1 = 0.01^2, C2 = 0.03^2, as
// fused_ssim / pytorch-msssim) and whose optimizers are torch.optim.Adam per parameter group (SelectiveAdam with
// `--packed` visibility masks).  Neither lives in /root/reference; the oracle (oracle/train_oracle.py) restates
// them with torch ops.
//
//   l1_ssim_fwd : one CTA per 16x16 tile of SSIM outputs; the 26x26 input windows of both images are staged in
//                 shared memory per channel, every thread accumulates the five window moments of its pixel, the
//                 CTA reduces its partial L1 / SSIM sums (written per CTA: the host-side sum is deterministic)
//                 and stores the three partial derivatives of SSIM w.r.t. the moments of the render;
//   l1_ssim_bwd : one CTA per 16x16 tile of INPUT pixels: full correlation of the three derivative maps with the
//                 window, plus the L1 sign term; writes d loss / d render.
//   adam_step   : fused Adam over [rows, width] fp32 parameters, optional per-row visibility mask.
#include "common.cuh"

namespace b2s {

constexpr int SS_WIN = 11, SS_HALO = SS_WIN - 1, SS_TILE = 16, SS_IN = SS_TILE + SS_HALO;   // 26
constexpr float SS_C1 = 0.01f * 0.01f, SS_C2 = 0.03f * 0.03f;

struct SsimWindow {
    float g[SS_WIN];
};

static SsimWindow make_window() {
    SsimWindow w;
    double s = 0.0, v[SS_WIN];
    for (int i = 0; i < SS_WIN; ++i) {
        const double d = i - SS_WIN / 2;
        v[i] = exp(-(d * d) / (2.0 * 1.5 * 1.5));
        s += v[i];
    }
    for (int i = 0; i < SS_WIN; ++i) w.g[i] = (float)(v[i] / s);
    return w;
}

// x, y: [C, H, W, ch] (the layout rasterization() renders in).  Outputs of the SSIM map: (H-10) x (W-10).
__global__ void __launch_bounds__(SS_TILE * SS_TILE) l1_ssim_fwd_kernel(int C, int H, int W, int ch, const float* __restrict__ x,
                                                                        const float* __restrict__ y, const SsimWindow win,
                                                                        float* __restrict__ dmaps /* [C,H',W',ch,3] or NULL */,
                                                                        float* __restrict__ partial /* [ctas][2] */) {
    __shared__ float sx[SS_IN][SS_IN + 1], sy[SS_IN][SS_IN + 1];
    __shared__ float s_red[2][SS_TILE * SS_TILE / 32];
    const int Ho = H - SS_HALO, Wo = W - SS_HALO;
    const int cam = blockIdx.z;
    const int ox0 = blockIdx.x * SS_TILE, oy0 = blockIdx.y * SS_TILE;
    const int tx = threadIdx.x % SS_TILE, ty = threadIdx.x / SS_TILE;
    const int ox = ox0 + tx, oy = oy0 + ty;
    const bool valid = ox < Wo && oy < Ho;
    // L1 is summed over the INPUT pixels: tile (bx, by) owns input pixels [16bx, 16bx+16) x [16by, 16by+16), and the
    // last tile row / column also owns the 10-pixel rim that no SSIM output tile starts in
    const bool last_x = blockIdx.x == gridDim.x - 1, last_y = blockIdx.y == gridDim.y - 1;
    float l1 = 0.f, ss = 0.f;
    for (int c = 0; c < ch; ++c) {
        __syncthreads();
        for (int t = threadIdx.x; t < SS_IN * SS_IN; t += SS_TILE * SS_TILE) {
            const int ly = t / SS_IN, lx = t - ly * SS_IN;
            const int gx = ox0 + lx, gy = oy0 + ly;
            float vx = 0.f, vy = 0.f;
            if (gx < W && gy < H) {
                const int64_t idx = (((int64_t)cam * H + gy) * W + gx) * ch + c;
                vx = x[idx]; vy = y[idx];
                const bool own = (lx < SS_TILE || last_x) && (ly < SS_TILE || last_y);
                if (own) l1 += fabsf(vx - vy);
            }
            sx[ly][lx] = vx; sy[ly][lx] = vy;
        }
        __syncthreads();
        if (valid) {
            float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int a = 0; a < SS_WIN; ++a) {
                float r1 = 0.f, r2 = 0.f, r11 = 0.f, r22 = 0.f, r12 = 0.f;
#pragma unroll
                for (int b = 0; b < SS_WIN; ++b) {
                    const float vx = sx[ty + a][tx + b], vy = sy[ty + a][tx + b], g = win.g[b];
                    r1 = fmaf(g, vx, r1); r2 = fmaf(g, vy, r2);
                    r11 = fmaf(g, vx * vx, r11); r22 = fmaf(g, vy * vy, r22); r12 = fmaf(g, vx * vy, r12);
                }
                const float ga = win.g[a];
                mu1 = fmaf(ga, r1, mu1); mu2 = fmaf(ga, r2, mu2);
                e11 = fmaf(ga, r11, e11); e22 = fmaf(ga, r22, e22); e12 = fmaf(ga, r12, e12);
            }
            const float s1 = e11 - mu1 * mu1, s2 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
            const float A1 = 2.f * mu1 * mu2 + SS_C1, A2 = 2.f * s12 + SS_C2;
            const float B1 = mu1 * mu1 + mu2 * mu2 + SS_C1, B2 = s1 + s2 + SS_C2;
            const float inv = 1.f / (B1 * B2);
            const float S = A1 * A2 * inv;
            ss += S;
            if (dmaps) {
                // partial derivatives of S w.r.t. the three window moments of the render: mu1, E[x^2], E[xy]
                const float d_mu1 = 2.f * mu2 * (A2 - A1) * inv - S * 2.f * mu1 * (B2 - B1) * inv;
                const float d_e11 = -S / B2;
                const float d_e12 = 2.f * A1 * inv;
                float* o = dmaps + ((((int64_t)cam * Ho + oy) * Wo + ox) * ch + c) * 3;
                o[0] = d_mu1; o[1] = d_e11; o[2] = d_e12;
            }
        }
    }
    l1 = warp_sum(l1); ss = warp_sum(ss);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = l1; s_red[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < SS_TILE * SS_TILE / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
        const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[2 * cta] = a; partial[2 * cta + 1] = b;
    }
}

// d loss / d x = w_l1 * sign(x - y) + w_ssim * sum_q win(p - q) * [d_mu1(q) + 2 x(p) d_e11(q) + y(p) d_e12(q)]
// (w_l1 = v * (1-lambda) / (C*H*W*ch), w_ssim = -v * lambda / (C*Ho*Wo*ch), v = upstream cotangent of the loss)
__global__ void __launch_bounds__(SS_TILE * SS_TILE) l1_ssim_bwd_kernel(int C, int H, int W, int ch, const float* __restrict__ x,
                                                                        const float* __restrict__ y, const SsimWindow win,
                                                                        const float* __restrict__ dmaps, float w_l1, float w_ssim,
                                                                        float* __restrict__ v_x) {
    __shared__ float sd[3][SS_IN][SS_IN + 1];
    const int Ho = H - SS_HALO, Wo = W - SS_HALO;
    const int cam = blockIdx.z;
    const int px0 = blockIdx.x * SS_TILE, py0 = blockIdx.y * SS_TILE;
    const int tx = threadIdx.x % SS_TILE, ty = threadIdx.x / SS_TILE;
    const int px = px0 + tx, py = py0 + ty;
    const bool inside = px < W && py < H;
    for (int c = 0; c < ch; ++c) {
        __syncthreads();
        // derivative-map window: outputs q = p - (a, b), a, b in [0, 10]  ->  q in [p0 - 10, p0 + 15]
        for (int t = threadIdx.x; t < SS_IN * SS_IN; t += SS_TILE * SS_TILE) {
            const int ly = t / SS_IN, lx = t - ly * SS_IN;
            const int qx = px0 - SS_HALO + lx, qy = py0 - SS_HALO + ly;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (qx >= 0 && qy >= 0 && qx < Wo && qy < Ho) {
                const float* o = dmaps + ((((int64_t)cam * Ho + qy) * Wo + qx) * ch + c) * 3;
                d0 = o[0]; d1 = o[1]; d2 = o[2];
            }
            sd[0][ly][lx] = d0; sd[1][ly][lx] = d1; sd[2][ly][lx] = d2;
        }
        __syncthreads();
        if (inside) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int a = 0; a < SS_WIN; ++a) {
                float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                for (int b = 0; b < SS_WIN; ++b) {
                    // output q = p - (a, b) sits at local (ty + 10 - a, tx + 10 - b); its weight for p is g[a] g[b]
                    const float g = win.g[b];
                    r0 = fmaf(g, sd[0][ty + SS_HALO - a][tx + SS_HALO - b], r0);
                    r1 = fmaf(g, sd[1][ty + SS_HALO - a][tx + SS_HALO - b], r1);
                    r2 = fmaf(g, sd[2][ty + SS_HALO - a][tx + SS_HALO - b], r2);
                }
                const float ga = win.g[a];
                a0 = fmaf(ga, r0, a0); a1 = fmaf(ga, r1, a1); a2 = fmaf(ga, r2, a2);
            }
            const int64_t idx = (((int64_t)cam * H + py) * W + px) * ch + c;
            const float vx = x[idx], vy = y[idx];
            const float d = vx - vy;
            const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            v_x[idx] = w_l1 * sgn + w_ssim * (a0 + 2.f * vx * a1 + vy * a2);
        }
    }
}

// fused Adam (torch.optim.Adam semantics: no weight decay, no amsgrad); mask: one byte per row, 0 = row untouched
__global__ void __launch_bounds__(256) adam_kernel(int64_t n, int width, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, const uint8_t* __restrict__ mask,
                                                   float lr, float b1, float b2, float eps, float inv_bc1, float inv_sqrt_bc2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i / width]) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= (lr * inv_bc1) * (mi / denom);
}

}  // namespace b2s

using namespace b2s;

extern "C" int64_t b2s_l1_ssim_ctas(int32_t C, int32_t H, int32_t W) {
    if (H <= SS_HALO || W <= SS_HALO || C <= 0) return 0;
    return (int64_t)C * cdiv(H - SS_HALO, SS_TILE) * cdiv(W - SS_HALO, SS_TILE);
}

extern "C" int b2s_l1_ssim_fwd(int32_t C, int32_t H, int32_t W, int32_t channels, const float* render, const float* target,
                               float* dmaps, float* partial, b2s_stream_t stream) {
    B2S_REQUIRE(C > 0 && channels > 0 && H > SS_HALO && W > SS_HALO, "image must be larger than the 11x11 window");
    B2S_REQUIRE(render && target && partial, "null pointer");
    static const SsimWindow win = make_window();
    dim3 grid((unsigned)cdiv(W - SS_HALO, SS_TILE), (unsigned)cdiv(H - SS_HALO, SS_TILE), (unsigned)C);
    l1_ssim_fwd_kernel<<<grid, SS_TILE * SS_TILE, 0, (cudaStream_t)stream>>>(C, H, W, channels, render, target, win, dmaps, partial);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_l1_ssim_bwd(int32_t C, int32_t H, int32_t W, int32_t channels, const float* render, const float* target,
                               const float* dmaps, float w_l1, float w_ssim, float* v_render, b2s_stream_t stream) {
    B2S_REQUIRE(C > 0 && channels > 0 && H > SS_HALO && W > SS_HALO, "image must be larger than the 11x11 window");
    B2S_REQUIRE(render && target && dmaps && v_render, "null pointer");
    static const SsimWindow win = make_window();
    dim3 grid((unsigned)cdiv(W, SS_TILE), (unsigned)cdiv(H, SS_TILE), (unsigned)C);
    l1_ssim_bwd_kernel<<<grid, SS_TILE * SS_TILE, 0, (cudaStream_t)stream>>>(C, H, W, channels, render, target, win, dmaps, w_l1,
                                                                             w_ssim, v_render);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_adam_step(int64_t rows, int32_t width, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             const uint8_t* row_mask, float lr, float beta1, float beta2, float eps, int64_t step,
                             b2s_stream_t stream) {
    B2S_REQUIRE(rows >= 0 && width > 0 && step >= 1, "bad sizes");
    if (rows == 0) return 0;
    B2S_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const int64_t n = rows * width;
    adam_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, width, param, grad, exp_avg, exp_avg_sq, row_mask, lr,
                                                                         beta1, beta2, eps, (float)(1.0 / bc1),
                                                                         (float)(1.0 / sqrt(bc2)));
    B2S_LAUNCH_CHECK();
    return 0;
}
