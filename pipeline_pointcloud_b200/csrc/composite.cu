// Stage 3: per-16x16-tile front-to-back alpha compositing, forward and backward.
//
// Replaces gsplat `rasterize_to_pixels` fwd/bwd; reached in the reference
// from source/container/src/main.py:1308-1316 / :1339-1347 on every training iteration.
//
// One CTA per (camera, tile), 256 threads, one pixel per thread.  A warp owns a compact 8x4 pixel
// block (not two 16-pixel rows) so that the Gaussians it can skip entirely are maximised.
// Gaussians of the tile are staged through shared memory in batches of 256 (each thread gathers one
// splat record by flatten_id), all lanes then read the same record (broadcast, conflict free).
//
// The staged record is pre-digested so the per-(pixel, Gaussian) reject path is ~15 instructions:
//   rec0 = (x, y, opacity, tau2)      tau2 = log2(255 * opacity) + margin
//   rec1 = (A, B, C, id)              (A, B, C) = log2(e) * (a/2, b, c/2)
//   sigma2 = dx*(A*dx + B*dy) + C*dy*dy = sigma * log2(e);   alpha = opacity * 2^-sigma2
// A splat is rejected before the exponential when sigma2 > tau2 (i.e. alpha < 1/255 by a margin); the
// exact `alpha < 1/255` test still runs afterwards, so the reference semantics are unchanged.
// Early termination: a warp whose 32 pixels are saturated skips the batch by vote, the CTA stops
// when all 256 are.
#include "common.cuh"

namespace b2s {

constexpr int TILE = 16;
constexpr int CTA = TILE * TILE;
constexpr float ALPHA_MAX = 0.999f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float T_STOP = 1e-4f;
constexpr float TAU_MARGIN = 2e-3f;          // in log2 units; only widens the set that reaches the exact test
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

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

// keeps a value in a register: without it ptxas rematerialises the pixel centre from %ctaid/%tid
// inside the hot loop (14 extra instructions per Gaussian)
__device__ __forceinline__ void pin(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// gathers one splat into the staged record format
__device__ __forceinline__ void stage_splat(int g, const float* __restrict__ means2d, const float* __restrict__ conics,
                                            const float* __restrict__ opacities, float4& rec0, float4& rec1) {
    const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
    const float op = opacities[g];
    rec0 = make_float4(xy.x, xy.y, op, __log2f(255.f * op) + TAU_MARGIN);
    rec1 = make_float4(0.5f * LOG2E * conics[3 * g], LOG2E * conics[3 * g + 1], 0.5f * LOG2E * conics[3 * g + 2],
                       __int_as_float(g));
}

// =================================================================================================
// forward
// =================================================================================================
template <int CH>
__global__ void __launch_bounds__(CTA) composite_fwd_kernel(const b2s_composite_fwd_args a, float* __restrict__ final_T) {
    __shared__ float4 s_rec0[CTA];
    __shared__ float4 s_rec1[CTA];
    __shared__ float s_col[CTA * CH];

    const int tile_x = blockIdx.x, tile_y = blockIdx.y, cam = blockIdx.z;
    const PixelMap pm = pixel_of_thread(tile_x, tile_y, a.width, a.height);
    float px = (float)pm.px + 0.5f, py = (float)pm.py + 0.5f;
    pin(px); pin(py);
    const int64_t tile_id = ((int64_t)cam * a.tile_h + tile_y) * a.tile_w + tile_x;
    const int64_t n_tiles_total = (int64_t)a.C * a.tile_h * a.tile_w;
    const int lo = a.isect_offsets[tile_id];
    const int hi = (tile_id == n_tiles_total - 1) ? (int)a.n_isects : a.isect_offsets[tile_id + 1];
    const int nb = (hi - lo + CTA - 1) / CTA;
    const bool chunked = a.chunk_index >= 0;
    if (chunked && a.chunk_index > 0 && (hi <= lo || a.tile_done[tile_id])) return;   // state of this tile is unchanged

    float T = 1.f;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    int last = 0;
    bool done = !pm.inside;
    const int64_t pix0 = ((int64_t)cam * a.height + pm.py) * a.width + pm.px;
    if (chunked && a.chunk_index > 0 && pm.inside) {       // resume from the state the previous chunks left
        const float Ts = final_T[pix0];
        done = Ts < 0.f;
        T = fabsf(Ts);
#pragma unroll
        for (int k = 0; k < CH; ++k) acc[k] = a.render_colors[pix0 * CH + k];
        last = a.last_ids[pix0];
    }
    const int last_tag = chunked ? (a.chunk_index << 28) : 0;

    for (int b = 0; b < nb; ++b) {
        if (__syncthreads_count(done) == CTA) break;  // also fences the previous batch's smem reads
        const int start = lo + b * CTA;
        const int idx = start + threadIdx.x;
        if (idx < hi) {
            const int g = a.flatten_ids[idx];
            float4 r0, r1;
            stage_splat(g, a.means2d, a.conics, a.opacities, r0, r1);
            s_rec0[threadIdx.x] = r0;
            s_rec1[threadIdx.x] = r1;
#pragma unroll
            for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)g * CH + k];
        }
        __syncthreads();
        if (__all_sync(FULL, done)) continue;  // early termination by warp vote
        const int bs = min(CTA, hi - start);
#pragma unroll 2
        for (int j = 0; j < bs && !done; ++j) {
            const float4 r0 = s_rec0[j];
            const float4 r1 = s_rec1[j];
            const float dx = r0.x - px, dy = r0.y - py;
            const float s2 = fmaf(dx, fmaf(r1.x, dx, r1.y * dy), r1.z * dy * dy);
            if (s2 < 0.f || s2 > r0.w) continue;
            const float alpha = fminf(ALPHA_MAX, r0.z * ex2_approx(-s2));
            if (alpha < ALPHA_MIN) continue;
            const float nT = T * (1.f - alpha);
            if (nT <= T_STOP) { done = true; break; }
            const float w = alpha * T;
#pragma unroll
            for (int k = 0; k < CH; ++k) acc[k] = fmaf(w, s_col[j * CH + k], acc[k]);
            T = nT;
            last = last_tag | (start + j);
        }
    }
    if (chunked) {
        const int n_done = __syncthreads_count(done);
        if (threadIdx.x == 0 && n_done == CTA) a.tile_done[tile_id] = 1;
    }
    if (pm.inside) {
        const int64_t pix = ((int64_t)cam * a.height + pm.py) * a.width + pm.px;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            a.render_colors[pix * CH + k] = (a.backgrounds && !chunked) ? acc[k] + T * a.backgrounds[cam * CH + k] : acc[k];
        a.render_alphas[pix] = 1.f - T;
        a.last_ids[pix] = last;
        if (final_T) final_T[pix] = done ? -T : T;    // sign bit = the pixel saturated (stopped at T_next <= 1e-4)
    }
}

// -------------------------------------------------------------------------------------------------
// Multi-value warp reduction.  Reducing V per-lane values with V independent butterflies costs 5*V
// shuffles.  Instead the lanes trade halves: at mask 16 every lane keeps one half of the slots and
// sends the other half to its partner, at mask 8 a quarter, ... so P slots cost P-1 shuffles plus one
// butterfly step per unused mask.  Afterwards slot s lives (fully summed) in the lanes whose high
// bits spell s, and each owner issues ONE red.global for its slot: the 9 atomics of a Gaussian become
// a single instruction with 9 active lanes.
// -------------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ float halving_reduce(float (&v)[P], int lane) {
    static_assert(P == 8 || P == 16 || P == 32, "P must be 8, 16 or 32");
    int m = 16;
#pragma unroll
    for (int n = P; n > 1; n >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[n / 2 + i];
            const float keep = up ? v[n / 2 + i] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, m);
        }
        m >>= 1;
    }
    float r = v[0];
    for (; m > 0; m >>= 1) r += __shfl_xor_sync(FULL, r, m);
    return r;
}
// slot held by `lane` after halving_reduce<P>, and whether this lane is the one that publishes it
template <int P>
__device__ __forceinline__ int halving_slot(int lane, bool& owner) {
    int slot = 0, m = 16;
#pragma unroll
    for (int n = P; n > 1; n >>= 1) {
        if (lane & m) slot += n / 2;
        m >>= 1;
    }
    owner = (lane & ((m << 1) - 1)) == 0;   // low (unused) bits zero
    return slot;
}

// =================================================================================================
// backward: replay back-to-front from the last contributor, multi-value warp reduce, one RED per warp
// =================================================================================================
template <int CH, bool ABS>
__global__ void __launch_bounds__(CTA) composite_bwd_kernel(const b2s_composite_bwd_args a, const float* __restrict__ final_T) {
    constexpr int V = CH + 6 + (ABS ? 2 : 0);   // values reduced per (warp, Gaussian)
    __shared__ float4 s_rec0[CTA];
    __shared__ float4 s_rec1[CTA];
    __shared__ float s_col[CTA * CH];
    __shared__ int s_max[CTA / 32];

    const int tile_x = blockIdx.x, tile_y = blockIdx.y, cam = blockIdx.z;
    const PixelMap pm = pixel_of_thread(tile_x, tile_y, a.width, a.height);
    float px = (float)pm.px + 0.5f, py = (float)pm.py + 0.5f;
    pin(px); pin(py);
    const int64_t tile_id = ((int64_t)cam * a.tile_h + tile_y) * a.tile_w + tile_x;
    const int64_t n_tiles_total = (int64_t)a.C * a.tile_h * a.tile_w;
    const int lo = a.isect_offsets[tile_id];
    const int hi = (tile_id == n_tiles_total - 1) ? (int)a.n_isects : a.isect_offsets[tile_id + 1];
    if (hi <= lo) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pix = pm.inside ? ((int64_t)cam * a.height + pm.py) * a.width + pm.px : 0;

    // which gradient slot this lane publishes after the multi-value reduction (lane-constant)
    // slots: 0-1 v_xy | 2-4 v_conic | 5..5+CH v_colour | 5+CH v_opacity | then |v_xy| (absgrad)
    float* slot_ptr = nullptr;
    unsigned slot_stride = 0;
    if (V <= 16) {
        bool owner;
        int slot = (V <= 9) ? halving_slot<8>(lane, owner) : halving_slot<16>(lane, owner);
        if (V == 9 && lane == 1) { slot = 8; owner = true; }
        if (owner) {
            if (slot < 2) { slot_ptr = a.v_means2d + slot; slot_stride = 2; }
            else if (slot < 5) { slot_ptr = a.v_conics + (slot - 2); slot_stride = 3; }
            else if (slot < 5 + CH) { slot_ptr = a.v_colors + (slot - 5); slot_stride = CH; }
            else if (slot == 5 + CH) { slot_ptr = a.v_opacities; slot_stride = 1; }
            else if (ABS && slot < 8 + CH) { slot_ptr = a.v_means2d_abs + (slot - 6 - CH); slot_stride = 2; }
        }
    }

    float T_final = 1.f, v_alpha_out = 0.f;
    float v_out[CH], S[CH];
    int bin_final = -1;
    float v_bg_dot = 0.f;  // sum_k bg_k * v_out_k
#pragma unroll
    for (int k = 0; k < CH; ++k) { v_out[k] = 0.f; S[k] = 0.f; }
    const bool chunked = a.chunk_index >= 0;
    if (pm.inside) {
        T_final = final_T ? fabsf(final_T[pix]) : 1.f - a.render_alphas[pix];
        bin_final = a.last_ids[pix];
        if (chunked) {      // last_ids = chunk << 28 | index into that chunk's list
            const int last_chunk = bin_final >> 28;
            bin_final = a.chunk_index > last_chunk ? -1 : (a.chunk_index == last_chunk ? (bin_final & ((1 << 28) - 1)) : hi - 1);
        }
        v_alpha_out = a.v_render_alphas ? a.v_render_alphas[pix] : 0.f;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            v_out[k] = a.v_render_colors ? a.v_render_colors[pix * CH + k] : 0.f;
            if (a.backgrounds) v_bg_dot = fmaf(a.backgrounds[cam * CH + k], v_out[k], v_bg_dot);
        }
        // a pixel that composited nothing has last == 0 by convention; T_final == 1 identifies it
        if (T_final == 1.f) bin_final = -1;
    }
    float T = T_final;
    if (chunked && pm.inside) {   // running state left by the later chunks
        T = a.state_T[pix];
#pragma unroll
        for (int k = 0; k < CH; ++k) S[k] = a.state_S[pix * CH + k];
    }
    const float v_T_final = T_final * (v_alpha_out - v_bg_dot);   // d(loss)/d(alpha_i) picks up  v_T_final / (1 - alpha_i)

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
        const int idx = batch_end - (int)threadIdx.x;               // smem slot t holds index batch_end - t
        if (idx >= lo) {
            const int g = a.flatten_ids[idx];
            float4 r0, r1;
            stage_splat(g, a.means2d, a.conics, a.opacities, r0, r1);
            s_rec0[threadIdx.x] = r0;
            s_rec1[threadIdx.x] = r1;
#pragma unroll
            for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)g * CH + k];
        }
        __syncthreads();
        const int t_min = pm.inside ? batch_end - bin_final : CTA;   // this pixel takes part from slot t_min on
        for (int t = max(0, batch_end - warp_final); t < bs; ++t) {
            const float4 r0 = s_rec0[t];
            const float4 r1 = s_rec1[t];
            const float dx = r0.x - px, dy = r0.y - py;
            const float s2 = fmaf(dx, fmaf(r1.x, dx, r1.y * dy), r1.z * dy * dy);
            // invalid lanes run the same straight-line code with alpha = 0, vis = 0: every term vanishes
            float vis = 0.f, alpha = 0.f;
            if (t >= t_min && s2 >= 0.f && s2 <= r0.w) {
                vis = ex2_approx(-s2);
                alpha = fminf(ALPHA_MAX, r0.z * vis);
                if (alpha < ALPHA_MIN) { alpha = 0.f; vis = 0.f; }
            }
            if (!__any_sync(FULL, alpha > 0.f)) continue;

            const float ra = rcp_approx(1.f - alpha);
            if (alpha > 0.f) T *= ra;                               // (never trust rcp(1) == 1 over thousands of steps)
            const float fac = alpha * T;
            float v_alpha = v_T_final * ra;
            float v_col[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const float c = s_col[t * CH + k];
                v_col[k] = fac * v_out[k];
                v_alpha = fmaf(fmaf(c, T, -S[k] * ra), v_out[k], v_alpha);
                S[k] = fmaf(c, fac, S[k]);
            }
            if (r0.z * vis > ALPHA_MAX) vis = 0.f;                   // clamped alpha: no gradient through sigma / opacity
            const float v_op = vis * v_alpha;
            const float v_s2 = -r0.z * v_op;                         // d/d(sigma)
            // sigma = 0.5*(a dx^2 + c dy^2) + b dx dy, with (a, b, c) = (2 A, B, 2 C) / log2(e)
            const float wx = v_s2 * dx, wy = v_s2 * dy;
            const float v_ca = 0.5f * wx * dx, v_cb = wx * dy, v_cc = 0.5f * wy * dy;
            const float k2 = 2.f * LN2, k1 = LN2;
            const float v_x = v_s2 * fmaf(k2 * r1.x, dx, k1 * r1.y * dy);
            const float v_y = v_s2 * fmaf(k1 * r1.y, dx, k2 * r1.z * dy);

            const unsigned g = (unsigned)__float_as_int(r1.w);
            if (V <= 9) {
                float r8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r8[k] = 0.f;
                r8[0] = v_x; r8[1] = v_y; r8[2] = v_ca; r8[3] = v_cb; r8[4] = v_cc;
#pragma unroll
                for (int k = 0; k < CH && k < 3; ++k) r8[5 + k] = v_col[k];
                if (V < 9) r8[5 + CH] = v_op;
                float mine = halving_reduce<8>(r8, lane);
                if (V == 9) {                      // ninth value: plain butterfly, published by lane 1
                    const float ninth = warp_sum(v_op);
                    if (lane == 1) mine = ninth;
                }
                if (slot_ptr) atomicAdd(slot_ptr + (uint64_t)g * slot_stride, mine);
            } else if (V <= 16) {
                float r16[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) r16[k] = 0.f;
                r16[0] = v_x; r16[1] = v_y; r16[2] = v_ca; r16[3] = v_cb; r16[4] = v_cc;
#pragma unroll
                for (int k = 0; k < CH && 5 + k < 16; ++k) r16[5 + k] = v_col[k];
                if (5 + CH < 16) r16[5 + CH] = v_op;
                if (ABS && 7 + CH < 16) { r16[6 + CH] = fabsf(v_x); r16[7 + CH] = fabsf(v_y); }
                const float mine = halving_reduce<16>(r16, lane);
                if (slot_ptr) atomicAdd(slot_ptr + (uint64_t)g * slot_stride, mine);
            } else {                               // wide feature vectors: one butterfly per value
                float v_ax = 0.f, v_ay = 0.f;
                if (ABS) { v_ax = warp_sum(fabsf(v_x)); v_ay = warp_sum(fabsf(v_y)); }
#pragma unroll
                for (int k = 0; k < CH; ++k) v_col[k] = warp_sum(v_col[k]);
                const float r_ca = warp_sum(v_ca), r_cb = warp_sum(v_cb), r_cc = warp_sum(v_cc);
                const float r_x = warp_sum(v_x), r_y = warp_sum(v_y), r_op = warp_sum(v_op);
                if (lane == 0) {
                    const uint64_t gg = g;
#pragma unroll
                    for (int k = 0; k < CH; ++k) atomicAdd(a.v_colors + gg * CH + k, v_col[k]);
                    atomicAdd(a.v_conics + 3 * gg, r_ca);
                    atomicAdd(a.v_conics + 3 * gg + 1, r_cb);
                    atomicAdd(a.v_conics + 3 * gg + 2, r_cc);
                    atomicAdd(a.v_means2d + 2 * gg, r_x);
                    atomicAdd(a.v_means2d + 2 * gg + 1, r_y);
                    atomicAdd(a.v_opacities + gg, r_op);
                    if (ABS) {
                        atomicAdd(a.v_means2d_abs + 2 * gg, v_ax);
                        atomicAdd(a.v_means2d_abs + 2 * gg + 1, v_ay);
                    }
                }
            }
        }
    }
    if (chunked && pm.inside) {
        a.state_T[pix] = T;
#pragma unroll
        for (int k = 0; k < CH; ++k) a.state_S[pix * CH + k] = S[k];
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
    if (a->v_means2d_abs) composite_bwd_kernel<CH, true><<<grid, CTA, 0, st>>>(*a, final_T);
    else composite_bwd_kernel<CH, false><<<grid, CTA, 0, st>>>(*a, final_T);
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
    if (a->chunk_index >= 0) {
        B2S_REQUIRE(a->chunk_index < 8 && a->n_isects < (1LL << 28), "chunked: at most 8 chunks of < 2^28 intersections");
        B2S_REQUIRE(a->tile_done && a->final_T, "chunked: tile_done and final_T are required");
    }
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
    if (a->chunk_index >= 0) B2S_REQUIRE(a->state_T && a->state_S && a->final_T, "chunked: state_T, state_S, final_T are required");
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
