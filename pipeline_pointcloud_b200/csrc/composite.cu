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
//
// Two list sources (template parameter BINS):
//   BINS = false  the reference's layout: the tile's Gaussians are flatten_ids[isect_offsets[t] .. isect_offsets[t+1])
//                 (the stage-level op `rasterize_to_pixels`);
//   BINS = true   two-level binning (binning.cu, `b2s_coarse_bin`): the CTA walks the depth-ordered list of its
//                 8x8-tile coarse bin, 1024 records per round (each thread tests 128 bits of records against the
//                 tile's mask bits), compacts the hits -- in order -- into a shared-memory ring and composites them
//                 256 at a time.  The hits are exactly the tile's slice of the reference's sorted list, but only
//                 the part in front of the saturation point is ever generated.
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

// ---- two-level binning: in-kernel expansion of a tile's list from its coarse bin ------------------------------
constexpr int CB = B2S_COARSE;               // tiles per coarse-bin side
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN = CTA * SCAN_PER_THREAD;  // records examined per round
constexpr int QRING = 2048;                  // ring of pending hits (>= CTA - 1 + SCAN, power of two)
constexpr int BINS_MAX_CH = 4;               // wider feature vectors use the classic lists (shared-memory budget)

// One round: examine the records [wbase, wbase + SCAN) (wbase a multiple of 4) that also lie in [lo, hi), keep those
// whose masks cover tile (lx, ly) of the coarse bin, and append (flat id, record position) to the ring at `qtail`
// in list order (REVERSE: descending positions, for the backward).  Returns the number of hits (CTA-uniform).
// s_wtot: 8 ints, alternate between two buffers on consecutive rounds (one barrier per round).
template <bool REVERSE>
__device__ __forceinline__ int scan_round(const uint32_t* __restrict__ recs, const int32_t* __restrict__ ids, int wbase,
                                          int lo, int hi, unsigned lx, unsigned ly, int2* s_q, int qtail, int* s_wtot) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // thread t owns 4 consecutive records (one aligned 128-bit group); in REVERSE, thread 0 owns the highest 4
    static_assert(SCAN_PER_THREAD == 4, "one 128-bit load per thread");
    const int g0 = REVERSE ? wbase + SCAN - SCAN_PER_THREAD * (t + 1) : wbase + SCAN_PER_THREAD * t;
    uint32_t r[SCAN_PER_THREAD];
    int id[SCAN_PER_THREAD];
    {
        uint4 rv = make_uint4(0u, 0u, 0u, 0u);
        int4 iv = make_int4(0, 0, 0, 0);
        if (g0 >= 0 && g0 < hi && g0 + 4 > lo) {     // the lists are padded to a multiple of 4 entries
            rv = __ldg(reinterpret_cast<const uint4*>(recs + g0));
            iv = __ldg(reinterpret_cast<const int4*>(ids + g0));
        }
        r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
        id[0] = iv.x; id[1] = iv.y; id[2] = iv.z; id[3] = iv.w;
    }
    unsigned hits = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        const int p = g0 + k;
        const unsigned h = (r[k] >> (16 + lx)) & (r[k] >> (24 + ly)) & 1u;
        hits |= ((p >= lo && p < hi) ? h : 0u) << k;
    }
    const int c = __popc(hits);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CTA / 32; ++w) {
        const int x = s_wtot[w];
        if (w < warp) before += x;
        total += x;
    }
    int slot = qtail + before + incl - c;
#pragma unroll
    for (int kk = 0; kk < SCAN_PER_THREAD; ++kk) {
        const int k = REVERSE ? SCAN_PER_THREAD - 1 - kk : kk;
        if ((hits >> k) & 1u) {
            s_q[slot & (QRING - 1)] = make_int2(id[k], g0 + k);
            ++slot;
        }
    }
    return total;
}

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

// pixel-centre bounds of the 8x4 block a warp owns
struct WarpBlock {
    float xmin, xmax, ymin, ymax;
};
__device__ __forceinline__ WarpBlock warp_block(int tile_x, int tile_y) {
    const int warp = threadIdx.x >> 5;
    WarpBlock b;
    b.xmin = (float)(tile_x * TILE + (warp & 1) * 8) + 0.5f;
    b.xmax = b.xmin + 7.f;
    b.ymin = (float)(tile_y * TILE + (warp >> 1) * 4) + 0.5f;
    b.ymax = b.ymin + 3.f;
    return b;
}
__device__ __forceinline__ bool box_hits(const float4& box, const WarpBlock& b) {
    return box.x <= b.xmax && box.y >= b.xmin && box.z <= b.ymax && box.w >= b.ymin;
}

// keeps a value in a register: without it ptxas rematerialises the pixel centre from %ctaid/%tid
// inside the hot loop (14 extra instructions per Gaussian)
__device__ __forceinline__ void pin(float& v) { asm volatile("" : "+f"(v)); }
// The staged splats live in ONE shared-memory block addressed from a pinned 32-bit shared-window base: ptxas otherwise
// rebuilds the window address of every array (S2R SR_CgaCtaId + MOV + LEA) inside the hot loop to stay under the register
// cap -- 12 of the forward's 56 instructions per (warp, splat) pair (profiles/r02_ncu_full_summary.md).
template <int CH>
struct __align__(16) Staged {
    float4 rec0[CTA];
    float4 rec1[CTA];
    float4 box[CTA];
    float col[CTA * CH];
};
constexpr int ST_REC1 = CTA * 16, ST_BOX = 2 * CTA * 16, ST_COL = 3 * CTA * 16;   // byte offsets inside Staged
__device__ __forceinline__ void pin_u32(uint32_t& v) { asm volatile("" : "+r"(v)); }
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
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

// gathers one splat into the staged record format; `box` = (xmin, xmax, ymin, ymax) of the region where the splat
// can pass the `sigma2 <= tau2` test (the axis-aligned bounding box of that ellipse, padded), used to skip the splat
// for a whole warp when it cannot reach the warp's 8x4 pixel block
__device__ __forceinline__ void stage_splat(int g, const float* __restrict__ means2d, const float* __restrict__ conics,
                                            const float* __restrict__ opacities, float4& rec0, float4& rec1, float4& box) {
    const float2 xy = reinterpret_cast<const float2*>(means2d)[g];
    const float op = opacities[g];
    const float tau = __log2f(255.f * op) + TAU_MARGIN;
    const float A = 0.5f * LOG2E * conics[3 * g], B = LOG2E * conics[3 * g + 1], C = 0.5f * LOG2E * conics[3 * g + 2];
    rec0 = make_float4(xy.x, xy.y, op, tau);
    rec1 = make_float4(A, B, C, __int_as_float(g));
    // { A dx^2 + B dx dy + C dy^2 <= tau }: |dx| <= sqrt(4 C tau / (4AC - B^2)), |dy| <= sqrt(4 A tau / (4AC - B^2))
    const float det4 = 4.f * A * C - B * B;
    float ex = 3.0e38f, ey = 3.0e38f;              // degenerate conic: never culled here (the exact test decides)
    if (!(tau >= 0.f)) { ex = -1.f; ey = -1.f; }   // alpha < 1/255 everywhere (also NaN opacity): empty region
    else if (det4 > 0.f && A > 0.f && C > 0.f) {
        const float k = 4.f * tau / det4;
        ex = sqrtf(k * C) * 1.0001f + 1e-3f;
        ey = sqrtf(k * A) * 1.0001f + 1e-3f;
    }
    box = make_float4(xy.x - ex, xy.x + ex, xy.y - ey, xy.y + ey);
}

// =================================================================================================
// forward
// =================================================================================================
// 5 resident CTAs per SM: ptxas then fits the kernels in 47 registers and schedules them measurably better than the
// unconstrained 48-register build (composite_fwd 0.372 -> 0.351 ms, composite_bwd 0.734 -> 0.663 ms on the same box;
// 6 CTAs / 40 registers: 0.358 / 0.706), profiles/r02_variants_c9.txt
#ifndef B2S_COMP_MINBLOCKS
#define B2S_COMP_MINBLOCKS 5
#endif
template <int CH, bool BINS>
__global__ void __launch_bounds__(CTA, B2S_COMP_MINBLOCKS) composite_fwd_kernel(const b2s_composite_fwd_args a, float* __restrict__ final_T) {
    __shared__ Staged<CH> s_st;
    __shared__ int2 s_q[BINS ? QRING : 1];
    __shared__ int s_wtot[2][CTA / 32];
    float4* const s_rec0 = s_st.rec0;
    float4* const s_rec1 = s_st.rec1;
    float4* const s_box = s_st.box;
    float* const s_col = s_st.col;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_st);
    pin_u32(sbase);

    // blockIdx.y counts from the first tile row of the launch's window (a tile stripe of a multi-GPU job, or the whole grid);
    // tile ids, bin keys and pixel centres are those of the whole image, the output tensors hold the window's rows only
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + a.tile_row_lo, cam = blockIdx.z;
    const PixelMap pm = pixel_of_thread(tile_x, tile_y, a.width, a.height);
    const int out_y0 = a.tile_row_lo * TILE, out_h = min(a.height, a.tile_row_hi * TILE) - out_y0;
    float px = (float)pm.px + 0.5f, py = (float)pm.py + 0.5f;
    pin(px); pin(py);
    const int lane = threadIdx.x & 31;
    const WarpBlock wb = warp_block(tile_x, tile_y);
    int lo, hi;
    if (BINS) {
        const int bins_w = (a.tile_w + CB - 1) / CB, bins_h = (a.tile_h + CB - 1) / CB;
        const int key = (cam * bins_h + tile_y / CB) * bins_w + tile_x / CB;
        lo = a.bin_offsets[key];
        hi = a.bin_offsets[key + 1];
    } else {
        const int64_t tile_id = ((int64_t)cam * a.tile_h + tile_y) * a.tile_w + tile_x;
        const int64_t n_tiles_total = (int64_t)a.C * a.tile_h * a.tile_w;
        lo = a.isect_offsets[tile_id];
        hi = (tile_id == n_tiles_total - 1) ? (int)a.n_isects : a.isect_offsets[tile_id + 1];
    }

    // A pixel that is done (saturated: stopped at T_next <= 1e-4; or outside the image) carries its transmittance NEGATED:
    // one sign test replaces a flag register in the hot loop, and -T is exactly what final_T stores for a saturated pixel.
    float T = pm.inside ? 1.f : -1.f;
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = 0.f;
    int last = 0;

    // One batch of `bs` staged splats.  32 splats at a time, each lane tests one splat's box against the warp's pixel
    // block; only the survivors are evaluated per pixel.  `last` records s_rec1[j].w (BINS: the record position) or
    // start + j.
    auto composite_batch = [&](int bs, int start) {
        for (int j0 = 0; j0 < bs; j0 += 32) {
            const int jl = j0 + lane;
            unsigned m = __ballot_sync(FULL, jl < bs && box_hits(s_box[jl], wb));
            while (m) {
                const int j = j0 + __ffs(m) - 1;
                m &= m - 1;
                if (T > 0.f) {
                    const uint32_t sa = sbase + (uint32_t)j * 16u;
                    const float4 r0 = lds128<0>(sa);
                    const float4 r1 = lds128<ST_REC1>(sa);
                    const float dx = r0.x - px, dy = r0.y - py;
                    const float s2 = fmaf(dx, fmaf(r1.x, dx, r1.y * dy), r1.z * dy * dy);
                    if (!(s2 < 0.f || s2 > r0.w)) {
                        const float alpha = fminf(ALPHA_MAX, r0.z * ex2_approx(-s2));
                        if (!(alpha < ALPHA_MIN)) {
                            const float nT = T * (1.f - alpha);
                            if (nT <= T_STOP) {
                                T = -T;
                            } else {
                                const float w = alpha * T;
                                const uint32_t sc = sbase + (uint32_t)j * (4u * CH);
                                if constexpr (CH % 4 == 0) {
#pragma unroll
                                    for (int k = 0; k < CH; k += 4) {
                                        const float4 c4 = k == 0 ? lds128<ST_COL>(sc) : lds128<ST_COL>(sc + 4u * k);
                                        acc[k] = fmaf(w, c4.x, acc[k]); acc[k + 1] = fmaf(w, c4.y, acc[k + 1]);
                                        acc[k + 2] = fmaf(w, c4.z, acc[k + 2]); acc[k + 3] = fmaf(w, c4.w, acc[k + 3]);
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < CH; ++k)
                                        acc[k] = fmaf(w, k == 0 ? lds32<ST_COL>(sc) : lds32<ST_COL>(sc + 4u * k), acc[k]);
                                }
                                T = nT;
                                last = BINS ? __float_as_int(r1.w) : start + j;
                            }
                        }
                    }
                }
            }
            if (__all_sync(FULL, !(T > 0.f))) break;
        }
    };

    if (BINS) {
        const unsigned lx = tile_x % CB, ly = tile_y % CB;
        int pos = lo & ~3;                       // next window of the coarse bin's list (4-aligned)
        int qhead = 0, qcount = 0, round = 0;
        while (true) {
            if (__syncthreads_count(!(T > 0.f)) == CTA) break;   // also fences the previous batch's smem reads
            bool scanned = false;
            while (qcount < CTA && pos < hi) {
                qcount += scan_round<false>(a.bin_recs, a.bin_ids, pos, lo, hi, lx, ly, s_q, qhead + qcount, s_wtot[round & 1]);
                pos += SCAN;
                ++round;
                scanned = true;
            }
            if (qcount == 0) break;
            if (scanned) __syncthreads();                  // ring writes of the last round -> staging reads
            const int bs = min(CTA, qcount);
            if ((int)threadIdx.x < bs) {
                const int2 e = s_q[(qhead + threadIdx.x) & (QRING - 1)];
                float4 r0, r1, bx;
                stage_splat(e.x, a.means2d, a.conics, a.opacities, r0, r1, bx);
                r1.w = __int_as_float(e.y);
                s_rec0[threadIdx.x] = r0;
                s_rec1[threadIdx.x] = r1;
                s_box[threadIdx.x] = bx;
#pragma unroll
                for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)e.x * CH + k];
            }
            __syncthreads();
            qhead += bs;
            qcount -= bs;
            if (__all_sync(FULL, !(T > 0.f))) continue;          // early termination by warp vote
            composite_batch(bs, 0);
        }
    } else {
        const int nb = (hi - lo + CTA - 1) / CTA;
        for (int b = 0; b < nb; ++b) {
            if (__syncthreads_count(!(T > 0.f)) == CTA) break;  // also fences the previous batch's smem reads
            const int start = lo + b * CTA;
            const int idx = start + threadIdx.x;
            if (idx < hi) {
                const int g = a.flatten_ids[idx];
                float4 r0, r1, bx;
                stage_splat(g, a.means2d, a.conics, a.opacities, r0, r1, bx);
                s_rec0[threadIdx.x] = r0;
                s_rec1[threadIdx.x] = r1;
                s_box[threadIdx.x] = bx;
#pragma unroll
                for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)g * CH + k];
            }
            __syncthreads();
            if (__all_sync(FULL, !(T > 0.f))) continue;  // early termination by warp vote
            composite_batch(min(CTA, hi - start), start);
        }
    }
    if (pm.inside) {
        const int64_t pix = ((int64_t)cam * out_h + (pm.py - out_y0)) * a.width + pm.px;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            a.render_colors[pix * CH + k] = a.backgrounds ? acc[k] + fabsf(T) * a.backgrounds[cam * CH + k] : acc[k];
        a.render_alphas[pix] = 1.f - fabsf(T);
        a.last_ids[pix] = last;
        if (final_T) final_T[pix] = T;               // sign bit = the pixel saturated (stopped at T_next <= 1e-4)
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
template <int CH, bool ABS, bool BINS>
__global__ void __launch_bounds__(CTA, B2S_COMP_MINBLOCKS) composite_bwd_kernel(const b2s_composite_bwd_args a, const float* __restrict__ final_T) {
    constexpr int V = CH + 6 + (ABS ? 2 : 0);   // values reduced per (warp, Gaussian)
    __shared__ Staged<CH> s_st;
    float4* const s_rec0 = s_st.rec0;
    float4* const s_rec1 = s_st.rec1;
    float4* const s_box = s_st.box;
    float* const s_col = s_st.col;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_st);
    pin_u32(sbase);
    __shared__ int s_pos[BINS ? CTA : 1];       // BINS: list position of each staged splat (descending)
    __shared__ int2 s_q[BINS ? QRING : 1];
    __shared__ int s_wtot[2][CTA / 32];
    __shared__ int s_max[CTA / 32];

    // blockIdx.y counts from the first tile row of the launch's window (a tile stripe of a multi-GPU job, or the whole grid);
    // tile ids, bin keys and pixel centres are those of the whole image, the output tensors hold the window's rows only
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + a.tile_row_lo, cam = blockIdx.z;
    const PixelMap pm = pixel_of_thread(tile_x, tile_y, a.width, a.height);
    const int out_y0 = a.tile_row_lo * TILE, out_h = min(a.height, a.tile_row_hi * TILE) - out_y0;
    float px = (float)pm.px + 0.5f, py = (float)pm.py + 0.5f;
    pin(px); pin(py);
    int lo, hi;
    if (BINS) {
        const int bins_w = (a.tile_w + CB - 1) / CB, bins_h = (a.tile_h + CB - 1) / CB;
        const int key = (cam * bins_h + tile_y / CB) * bins_w + tile_x / CB;
        lo = a.bin_offsets[key];
        hi = a.bin_offsets[key + 1];
    } else {
        const int64_t tile_id = ((int64_t)cam * a.tile_h + tile_y) * a.tile_w + tile_x;
        const int64_t n_tiles_total = (int64_t)a.C * a.tile_h * a.tile_w;
        lo = a.isect_offsets[tile_id];
        hi = (tile_id == n_tiles_total - 1) ? (int)a.n_isects : a.isect_offsets[tile_id + 1];
    }
    if (hi <= lo) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WarpBlock wb = warp_block(tile_x, tile_y);
    const int64_t pix = pm.inside ? ((int64_t)cam * out_h + (pm.py - out_y0)) * a.width + pm.px : 0;

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
    if (pm.inside) {
        T_final = final_T ? fabsf(final_T[pix]) : 1.f - a.render_alphas[pix];
        bin_final = a.last_ids[pix];
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

    // One staged splat (smem slot t) against this thread's pixel; `mine` = the pixel takes part in it.
    auto splat_bwd = [&](int t, bool mine) {
        const uint32_t sa = sbase + (uint32_t)t * 16u;
        const float4 r0 = lds128<0>(sa);
        const float4 r1 = lds128<ST_REC1>(sa);
        const float dx = r0.x - px, dy = r0.y - py;
        const float s2 = fmaf(dx, fmaf(r1.x, dx, r1.y * dy), r1.z * dy * dy);
        // invalid lanes run the same straight-line code with alpha = 0, vis = 0: every term vanishes
        float vis = 0.f, alpha = 0.f;
        if (mine && s2 >= 0.f && s2 <= r0.w) {
            vis = ex2_approx(-s2);
            alpha = fminf(ALPHA_MAX, r0.z * vis);
            if (alpha < ALPHA_MIN) { alpha = 0.f; vis = 0.f; }
        }
        if (!__any_sync(FULL, alpha > 0.f)) return;

        const float ra = rcp_approx(1.f - alpha);
        if (alpha > 0.f) T *= ra;                               // (never trust rcp(1) == 1 over thousands of steps)
        const float fac = alpha * T;
        float v_alpha = v_T_final * ra;
        float v_col[CH];
        const uint32_t sc = sbase + (uint32_t)t * (4u * CH);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const float c = k == 0 ? lds32<ST_COL>(sc) : lds32<ST_COL>(sc + 4u * k);
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
        // (tried, round 2: letting the lanes of a pair with <= 2 / 4 / 8 contributing pixels add straight into the gradients
        // instead of paying the 32-lane reduction -- 0.661 -> 0.678 / 0.672 / 0.666 ms: no gain, profiles/r02_variants_c9.txt)
        if (V <= 9) {
            float r8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r8[k] = 0.f;
            r8[0] = v_x; r8[1] = v_y; r8[2] = v_ca; r8[3] = v_cb; r8[4] = v_cc;
#pragma unroll
            for (int k = 0; k < CH && k < 3; ++k) r8[5 + k] = v_col[k];
            if (V < 9) r8[5 + CH] = v_op;
            float mine_v = halving_reduce<8>(r8, lane);
            if (V == 9) {                      // ninth value: plain butterfly, published by lane 1
                const float ninth = warp_sum(v_op);
                if (lane == 1) mine_v = ninth;
            }
            if (slot_ptr) atomicAdd(slot_ptr + g * slot_stride, mine_v);    // 32-bit element offset: rows x width < 2^32 (include/b2s.h)
        } else if (V <= 16) {
            float r16[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) r16[k] = 0.f;
            r16[0] = v_x; r16[1] = v_y; r16[2] = v_ca; r16[3] = v_cb; r16[4] = v_cc;
#pragma unroll
            for (int k = 0; k < CH && 5 + k < 16; ++k) r16[5 + k] = v_col[k];
            if (5 + CH < 16) r16[5 + CH] = v_op;
            if (ABS && 7 + CH < 16) { r16[6 + CH] = fabsf(v_x); r16[7 + CH] = fabsf(v_y); }
            const float mine_v = halving_reduce<16>(r16, lane);
            if (slot_ptr) atomicAdd(slot_ptr + g * slot_stride, mine_v);    // 32-bit element offset: rows x width < 2^32 (include/b2s.h)
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
    };

    if (BINS) {
        // walk the coarse bin's list back to front from the CTA's last contributor
        const unsigned lx = tile_x % CB, ly = tile_y % CB;
        const int range_hi = cta_final + 1;                    // records [lo, range_hi) can contribute
        int wbase = ((range_hi + 3) & ~3) - SCAN;              // window [wbase, wbase + SCAN), walked downwards
        int qhead = 0, qcount = 0, round = 0;
        while (true) {
            __syncthreads();                                   // fences the previous batch's smem reads
            bool scanned = false;
            while (qcount < CTA && wbase + SCAN > lo) {
                qcount += scan_round<true>(a.bin_recs, a.bin_ids, wbase, lo, range_hi, lx, ly, s_q, qhead + qcount,
                                           s_wtot[round & 1]);
                wbase -= SCAN;
                ++round;
                scanned = true;
            }
            if (qcount == 0) break;
            if (scanned) __syncthreads();
            const int bs = min(CTA, qcount);
            if ((int)threadIdx.x < bs) {
                const int2 e = s_q[(qhead + threadIdx.x) & (QRING - 1)];
                float4 r0, r1, bx;
                stage_splat(e.x, a.means2d, a.conics, a.opacities, r0, r1, bx);
                s_rec0[threadIdx.x] = r0;
                s_rec1[threadIdx.x] = r1;
                s_box[threadIdx.x] = bx;
                s_pos[threadIdx.x] = e.y;
#pragma unroll
                for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)e.x * CH + k];
            }
            __syncthreads();
            qhead += bs;
            qcount -= bs;
            for (int t0 = 0; t0 < bs; t0 += 32) {
                const int tl = t0 + lane;
                // skip what lies behind every pixel of this warp, and what cannot reach its pixel block
                unsigned m = __ballot_sync(FULL, tl < bs && s_pos[tl] <= warp_final && box_hits(s_box[tl], wb));
                while (m) {
                    const int t = t0 + __ffs(m) - 1;
                    m &= m - 1;
                    splat_bwd(t, s_pos[t] <= bin_final);
                }
            }
        }
    } else {
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
                float4 r0, r1, bx;
                stage_splat(g, a.means2d, a.conics, a.opacities, r0, r1, bx);
                s_rec0[threadIdx.x] = r0;
                s_rec1[threadIdx.x] = r1;
                s_box[threadIdx.x] = bx;
#pragma unroll
                for (int k = 0; k < CH; ++k) s_col[threadIdx.x * CH + k] = a.colors[(int64_t)g * CH + k];
            }
            __syncthreads();
            const int t_min = pm.inside ? batch_end - bin_final : CTA;   // this pixel takes part from slot t_min on
            const int t_first = batch_end - warp_final;                  // slots before it lie behind every pixel of this warp
            for (int t0 = max(0, t_first) & ~31; t0 < bs; t0 += 32) {
                const int tl = t0 + lane;
                unsigned m = __ballot_sync(FULL, tl < bs && tl >= t_first && box_hits(s_box[tl], wb));
                while (m) {
                    const int t = t0 + __ffs(m) - 1;
                    m &= m - 1;
                    splat_bwd(t, t >= t_min);
                }
            }
        }
    }
}

template <int CH>
static int launch_fwd(const b2s_composite_fwd_args* a, float* final_T, cudaStream_t st) {
    dim3 grid(a->tile_w, a->tile_row_hi - a->tile_row_lo, a->C);
    if (a->bin_offsets) {
        if constexpr (CH <= BINS_MAX_CH) composite_fwd_kernel<CH, true><<<grid, CTA, 0, st>>>(*a, final_T);
        else { set_error("composite: two-level binning supports up to %d channels", BINS_MAX_CH); return B2S_E_UNSUPPORTED; }
    } else composite_fwd_kernel<CH, false><<<grid, CTA, 0, st>>>(*a, final_T);
    B2S_LAUNCH_CHECK();
    return 0;
}
template <int CH>
static int launch_bwd(const b2s_composite_bwd_args* a, const float* final_T, cudaStream_t st) {
    dim3 grid(a->tile_w, a->tile_row_hi - a->tile_row_lo, a->C);
    if (a->bin_offsets) {
        if constexpr (CH <= BINS_MAX_CH) {
            if (a->v_means2d_abs) composite_bwd_kernel<CH, true, true><<<grid, CTA, 0, st>>>(*a, final_T);
            else composite_bwd_kernel<CH, false, true><<<grid, CTA, 0, st>>>(*a, final_T);
        } else { set_error("composite: two-level binning supports up to %d channels", BINS_MAX_CH); return B2S_E_UNSUPPORTED; }
    } else {
        if (a->v_means2d_abs) composite_bwd_kernel<CH, true, false><<<grid, CTA, 0, st>>>(*a, final_T);
        else composite_bwd_kernel<CH, false, false><<<grid, CTA, 0, st>>>(*a, final_T);
    }
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s

using namespace b2s;

// The exact final transmittance is carried in an optional side buffer (args.final_T):
// 1 - alpha loses up to 6e-4 relative on nearly opaque pixels, which the backward would inherit.
extern "C" int b2s_composite_fwd(const b2s_composite_fwd_args* a_in, b2s_stream_t stream) {
    B2S_REQUIRE(a_in != nullptr, "null args");
    b2s_composite_fwd_args args = *a_in;
    const b2s_composite_fwd_args* a = &args;
    if (args.tile_row_hi == 0 && args.tile_row_lo == 0) args.tile_row_hi = args.tile_h;   // no window: the whole grid
    B2S_REQUIRE(args.tile_row_lo >= 0 && args.tile_row_lo <= args.tile_row_hi && args.tile_row_hi <= args.tile_h, "bad tile-row window");
    B2S_REQUIRE(a->tile_size == TILE, "tile_size must be 16");
    B2S_REQUIRE(a->C >= 0 && a->width > 0 && a->height > 0, "bad sizes");
    B2S_REQUIRE(a->tile_w == (a->width + TILE - 1) / TILE && a->tile_h == (a->height + TILE - 1) / TILE, "tile grid mismatch");
    if (a->C == 0 || a->tile_row_hi == a->tile_row_lo) return 0;
    B2S_REQUIRE((a->isect_offsets || a->bin_offsets) && a->render_colors && a->render_alphas && a->last_ids, "null pointer");
    B2S_REQUIRE(a->n_isects == 0 || (a->means2d && a->conics && a->colors && a->opacities), "null input");
    B2S_REQUIRE(a->n_isects == 0 || (a->bin_offsets ? (a->bin_recs && a->bin_ids) : a->flatten_ids != nullptr), "null list");
    B2S_REQUIRE(a->n_isects < 2147483647LL, "n_isects must fit int32");
    if (a->bin_offsets)
        B2S_REQUIRE(((reinterpret_cast<uintptr_t>(a->bin_recs) | reinterpret_cast<uintptr_t>(a->bin_ids)) & 15u) == 0,
                    "bin_recs / bin_ids must be 16-byte aligned");
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

extern "C" int b2s_composite_bwd(const b2s_composite_bwd_args* a_in, b2s_stream_t stream) {
    B2S_REQUIRE(a_in != nullptr, "null args");
    b2s_composite_bwd_args args = *a_in;
    const b2s_composite_bwd_args* a = &args;
    if (args.tile_row_hi == 0 && args.tile_row_lo == 0) args.tile_row_hi = args.tile_h;
    B2S_REQUIRE(args.tile_row_lo >= 0 && args.tile_row_lo <= args.tile_row_hi && args.tile_row_hi <= args.tile_h, "bad tile-row window");
    B2S_REQUIRE(a->tile_size == TILE, "tile_size must be 16");
    B2S_REQUIRE(a->C >= 0 && a->width > 0 && a->height > 0, "bad sizes");
    if (a->C == 0 || a->n_isects == 0 || a->tile_row_hi == a->tile_row_lo) return 0;
    B2S_REQUIRE(a->means2d && a->conics && a->colors && a->opacities, "null input");
    B2S_REQUIRE(a->bin_offsets ? (a->bin_recs && a->bin_ids) : (a->isect_offsets && a->flatten_ids), "null list");
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
