// Stage 3: per-16x16-tile front-to-back alpha compositing, forward and backward.
//
// Replaces gsplat `rasterize_to_pixels` fwd/bwd; reached in the reference
// from source/container/src/main.py:1308-1316 / :1339-1347 on every training iteration.
//
// One CTA per (camera, tile), 256 threads, one pixel per thread.  A warp owns a compact 8x4 pixel
// block (not two 16-pixel rows) so that the Gaussians it can skip entirely are maximised.
// Gaussians of the tile are staged through shared memory in batches of 256 (each thread gathers one
// splat record by flatten_id), all lanes then read the same record (broadcast, conflict free).
// A splat is rejected *before* the exponential with sigma > ln(255*opacity); the exact alpha < 1/255
// test still runs afterwards, so semantics are unchanged.  Early termination: a warp whose 32 pixels
// are saturated skips the batch by vote, the CTA stops when all 256 are.
#include "common.cuh"

namespace b2s {

constexpr int TILE = 16;
constexpr int CTA = TILE * TILE;
constexpr float ALPHA_MAX = 0.999f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float T_STOP = 1e-4f;
constexpr float TAU_MARGIN = 1e-3f;

struct PixelMap {
    int px, py;
    bool inside;
};
__device__ __forceinline__ PixelMap pixel_of_thread(int tile_x, int tile_y, int W, int H) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PixelMap p;
    p.px = tile_x * TILE + (warp & 1) * 8 + (lane & 7);
    p.py = tile_y * TILE + (warp >> 1) * 4 + (lane >> 3);
    p.inside = p.px < W && p.py < H;
    return p;
}

// =================================================================================================
// forward
// =================================================================================================
template <int CH>
__global__ void __launch_bounds__(CTA) composite_fwd_kernel(const b2s_composite_fwd_args a, float* __restrict__ final_T) {
    __shared__ float4 s_xyo[CTA];  // x, y, opacity, tau = ln(255*opacity) + margin
    __shared__ float4 s_con[CTA];  // conic a, b, c, -
    __shared__ float s_col[CTA * CH];

    const int tile_x = blockIdx.x, tile_y = blockIdx.y, cam = blockIdx.z;
    const PixelMap pm = pixel_of_thread(tile_x, tile_y, a.width, a.height);
    const float px = (float)pm.px + 0.5f, py = (float)pm.py + 0.5f;
    const int64_t tile_id = ((int64_t)cam * a.tile_h + tile_y) * a.tile_w + tile_x;
    const int64_t n_tiles_total = (int64_t)a.C * a.tile_h * a.tile_w;
    const int lo = a.isect_offsets[tile_id];
    const int hi = (tile_id == n_tiles_total - 1) ? (int)a.n_isects : a.isect_offsets[tile_id + 1];
    const int nb = (hi - lo + CTA - 1) / CTA;

    float T = 1.f;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    int last = 0;
    bool done = !pm.inside;

    for (int b = 0; b < nb; ++b) {
        if (__syncthreads_count(done) == CTA) break;  // also fences the previous batch's smem reads
        const int start = lo + b * CTA;
        const int idx = start + threadIdx.x;
        if (idx < hi) {
            const int g = a.flatten_ids[idx];
            const float2 xy = reinterpret_cast<const float2*>(a.means2d)[g];
            const float op = a.opacities[g];
            s_xyo[threadIdx.x] = make_float4(xy.x, xy.y, op, __logf(255.f * op) + TAU_MARGIN);
            s_con[threadIdx.x] = make_float4(a.conics[3 * g], a.conics[3 * g + 1], a.conics[3 * g + 2], 0.f);
#pragma unroll
            for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)g * CH + k];
        }
        __syncthreads();
        if (__all_sync(FULL, done)) continue;  // early termination by warp vote
        const int bs = min(CTA, hi - start);
        for (int j = 0; j < bs && !done; ++j) {
            const float4 xyo = s_xyo[j];
            const float4 con = s_con[j];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
            if (sigma < 0.f || sigma > xyo.w) continue;
            const float alpha = fminf(ALPHA_MAX, xyo.z * __expf(-sigma));
            if (alpha < ALPHA_MIN) continue;
            const float nT = T * (1.f - alpha);
            if (nT <= T_STOP) { done = true; break; }
            const float w = alpha * T;
#pragma unroll
            for (int k = 0; k < CH; ++k) acc[k] += w * s_col[j * CH + k];
            T = nT;
            last = start + j;
        }
    }
    if (pm.inside) {
        const int64_t pix = ((int64_t)cam * a.height + pm.py) * a.width + pm.px;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            a.render_colors[pix * CH + k] = a.backgrounds ? acc[k] + T * a.backgrounds[cam * CH + k] : acc[k];
        a.render_alphas[pix] = 1.f - T;
        a.last_ids[pix] = last;
        if (final_T) final_T[pix] = T;
    }
}

// =================================================================================================
// backward: replay back-to-front from the last contributor, warp-reduce, one atomic per warp
// =================================================================================================
template <int CH>
__global__ void __launch_bounds__(CTA) composite_bwd_kernel(const b2s_composite_bwd_args a, const float* __restrict__ final_T) {
    __shared__ float4 s_xyo[CTA];
    __shared__ float4 s_con[CTA];
    __shared__ float s_col[CTA * CH];
    __shared__ int s_id[CTA];
    __shared__ int s_max[CTA / 32];

    const int tile_x = blockIdx.x, tile_y = blockIdx.y, cam = blockIdx.z;
    const PixelMap pm = pixel_of_thread(tile_x, tile_y, a.width, a.height);
    const float px = (float)pm.px + 0.5f, py = (float)pm.py + 0.5f;
    const int64_t tile_id = ((int64_t)cam * a.tile_h + tile_y) * a.tile_w + tile_x;
    const int64_t n_tiles_total = (int64_t)a.C * a.tile_h * a.tile_w;
    const int lo = a.isect_offsets[tile_id];
    const int hi = (tile_id == n_tiles_total - 1) ? (int)a.n_isects : a.isect_offsets[tile_id + 1];
    if (hi <= lo) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pix = pm.inside ? ((int64_t)cam * a.height + pm.py) * a.width + pm.px : 0;

    float T_final = 1.f, v_alpha_out = 0.f;
    float v_out[CH], bg[CH], S[CH];
    int bin_final = -1;
#pragma unroll
    for (int k = 0; k < CH; ++k) { v_out[k] = 0.f; S[k] = 0.f; bg[k] = a.backgrounds ? a.backgrounds[cam * CH + k] : 0.f; }
    if (pm.inside) {
        T_final = final_T ? final_T[pix] : 1.f - a.render_alphas[pix];
        bin_final = a.last_ids[pix];
        v_alpha_out = a.v_render_alphas ? a.v_render_alphas[pix] : 0.f;
#pragma unroll
        for (int k = 0; k < CH; ++k) v_out[k] = a.v_render_colors ? a.v_render_colors[pix * CH + k] : 0.f;
        // a pixel that composited nothing has last == 0 by convention; T_final == 1 identifies it
        if (T_final == 1.f) bin_final = -1;
    }
    float T = T_final;
    float v_bg_dot = 0.f;  // sum_k bg_k * v_out_k
#pragma unroll
    for (int k = 0; k < CH; ++k) v_bg_dot += bg[k] * v_out[k];

    int warp_final = bin_final;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) warp_final = max(warp_final, __shfl_xor_sync(FULL, warp_final, d));
    if (lane == 0) s_max[warp] = warp_final;
    __syncthreads();
    int cta_final = -1;
#pragma unroll
    for (int w = 0; w < CTA / 32; ++w) cta_final = max(cta_final, s_max[w]);
    if (cta_final < lo) return;

    // only [lo, cta_final] can contribute; walk it back to front in batches of 256
    const int range_hi = min(hi, cta_final + 1);
    const int nb = (range_hi - lo + CTA - 1) / CTA;
    for (int b = 0; b < nb; ++b) {
        __syncthreads();
        const int batch_end = range_hi - 1 - b * CTA;               // highest index of this batch
        const int bs = min(CTA, batch_end + 1 - lo);
        const int idx = batch_end - threadIdx.x;                    // smem slot t holds index batch_end - t
        if (idx >= lo) {
            const int g = a.flatten_ids[idx];
            s_id[threadIdx.x] = g;
            const float2 xy = reinterpret_cast<const float2*>(a.means2d)[g];
            const float op = a.opacities[g];
            s_xyo[threadIdx.x] = make_float4(xy.x, xy.y, op, __logf(255.f * op) + TAU_MARGIN);
            s_con[threadIdx.x] = make_float4(a.conics[3 * g], a.conics[3 * g + 1], a.conics[3 * g + 2], 0.f);
#pragma unroll
            for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)g * CH + k];
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_final); t < bs; ++t) {
            const float4 xyo = s_xyo[t];
            const float4 con = s_con[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
            bool valid = pm.inside && (batch_end - t <= bin_final) && !(sigma < 0.f || sigma > xyo.w);
            float vis = 0.f, alpha = 0.f;
            if (valid) {
                vis = __expf(-sigma);
                alpha = fminf(ALPHA_MAX, xyo.z * vis);
                valid = alpha >= ALPHA_MIN;
            }
            if (!__any_sync(FULL, valid)) continue;

            float v_col[CH];
            float v_ca = 0.f, v_cb = 0.f, v_cc = 0.f, v_x = 0.f, v_y = 0.f, v_op = 0.f;
#pragma unroll
            for (int k = 0; k < CH; ++k) v_col[k] = 0.f;
            if (valid) {
                const float ra = 1.f / (1.f - alpha);
                T *= ra;
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const float c = s_col[t * CH + k];
                    v_col[k] = fac * v_out[k];
                    v_alpha += (c * T - S[k] * ra) * v_out[k];
                    S[k] += c * fac;
                }
                v_alpha += T_final * ra * v_alpha_out;
                v_alpha -= T_final * ra * v_bg_dot;
                if (xyo.z * vis <= ALPHA_MAX) {
                    const float v_sigma = -xyo.z * vis * v_alpha;
                    v_ca = 0.5f * v_sigma * dx * dx;
                    v_cb = v_sigma * dx * dy;
                    v_cc = 0.5f * v_sigma * dy * dy;
                    v_x = v_sigma * (con.x * dx + con.y * dy);
                    v_y = v_sigma * (con.y * dx + con.z * dy);
                    v_op = vis * v_alpha;
                }
            }
            float v_ax = 0.f, v_ay = 0.f;
            if (a.v_means2d_abs) { v_ax = warp_sum(fabsf(v_x)); v_ay = warp_sum(fabsf(v_y)); }
#pragma unroll
            for (int k = 0; k < CH; ++k) v_col[k] = warp_sum(v_col[k]);
            v_ca = warp_sum(v_ca); v_cb = warp_sum(v_cb); v_cc = warp_sum(v_cc);
            v_x = warp_sum(v_x); v_y = warp_sum(v_y); v_op = warp_sum(v_op);
            if (lane == 0) {
                const int64_t g = s_id[t];
#pragma unroll
                for (int k = 0; k < CH; ++k) atomicAdd(a.v_colors + g * CH + k, v_col[k]);
                atomicAdd(a.v_conics + 3 * g, v_ca);
                atomicAdd(a.v_conics + 3 * g + 1, v_cb);
                atomicAdd(a.v_conics + 3 * g + 2, v_cc);
                atomicAdd(a.v_means2d + 2 * g, v_x);
                atomicAdd(a.v_means2d + 2 * g + 1, v_y);
                atomicAdd(a.v_opacities + g, v_op);
                if (a.v_means2d_abs) {
                    atomicAdd(a.v_means2d_abs + 2 * g, v_ax);
                    atomicAdd(a.v_means2d_abs + 2 * g + 1, v_ay);
                }
            }
        }
    }
}

template <int CH>
static int launch_fwd(const b2s_composite_fwd_args* a, float* final_T, cudaStream_t st) {
    dim3 grid(a->tile_w, a->tile_h, a->C);
    composite_fwd_kernel<CH><<<grid, CTA, 0, st>>>(*a, final_T);
    B2S_LAUNCH_CHECK();
    return 0;
}
template <int CH>
static int launch_bwd(const b2s_composite_bwd_args* a, const float* final_T, cudaStream_t st) {
    dim3 grid(a->tile_w, a->tile_h, a->C);
    composite_bwd_kernel<CH><<<grid, CTA, 0, st>>>(*a, final_T);
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s

using namespace b2s;

// The exact final transmittance is carried in an optional side buffer (args.final_T):
// 1 - alpha loses up to 6e-4 relative on nearly opaque pixels, which the backward would inherit.
extern "C" int b2s_composite_fwd(const b2s_composite_fwd_args* a, b2s_stream_t stream) {
    B2S_REQUIRE(a != nullptr, "null args");
    B2S_REQUIRE(a->tile_size == TILE, "tile_size must be 16");
    B2S_REQUIRE(a->C >= 0 && a->width > 0 && a->height > 0, "bad sizes");
    B2S_REQUIRE(a->tile_w == (a->width + TILE - 1) / TILE && a->tile_h == (a->height + TILE - 1) / TILE, "tile grid mismatch");
    if (a->C == 0) return 0;
    B2S_REQUIRE(a->isect_offsets && a->render_colors && a->render_alphas && a->last_ids, "null pointer");
    B2S_REQUIRE(a->n_isects == 0 || (a->means2d && a->conics && a->colors && a->opacities && a->flatten_ids), "null input");
    B2S_REQUIRE(a->n_isects < 2147483647LL, "n_isects must fit int32");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_FWD(CHN) (launch_fwd<CHN>(a, a->final_T, st))
    switch (a->channels) {
        case 1: return CALL_FWD(1);
        case 2: return CALL_FWD(2);
        case 3: return CALL_FWD(3);
        case 4: return CALL_FWD(4);
        case 8: return CALL_FWD(8);
        case 16: return CALL_FWD(16);
        case 32: return CALL_FWD(32);
        default: set_error("composite: channels must be 1,2,3,4,8,16,32"); return B2S_E_UNSUPPORTED;
    }
#undef CALL_FWD
}

extern "C" int b2s_composite_bwd(const b2s_composite_bwd_args* a, b2s_stream_t stream) {
    B2S_REQUIRE(a != nullptr, "null args");
    B2S_REQUIRE(a->tile_size == TILE, "tile_size must be 16");
    B2S_REQUIRE(a->C >= 0 && a->width > 0 && a->height > 0, "bad sizes");
    if (a->C == 0 || a->n_isects == 0) return 0;
    B2S_REQUIRE(a->isect_offsets && a->flatten_ids && a->means2d && a->conics && a->colors && a->opacities, "null input");
    B2S_REQUIRE(a->last_ids && (a->final_T || a->render_alphas), "forward state missing");
    B2S_REQUIRE(a->v_means2d && a->v_conics && a->v_colors && a->v_opacities, "null gradient output");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL_BWD(CHN) (launch_bwd<CHN>(a, a->final_T, st))
    switch (a->channels) {
        case 1: return CALL_BWD(1);
        case 2: return CALL_BWD(2);
        case 3: return CALL_BWD(3);
        case 4: return CALL_BWD(4);
        case 8: return CALL_BWD(8);
        case 16: return CALL_BWD(16);
        case 32: return CALL_BWD(32);
        default: set_error("composite: channels must be 1,2,3,4,8,16,32"); return B2S_E_UNSUPPORTED;
    }
#undef CALL_BWD
}
