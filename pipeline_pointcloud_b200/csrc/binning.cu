// Stage 2: tile binning -- count, scan, key fill, (tile, depth) radix sort, per-tile offsets.
//
// Replaces gsplat `isect_tiles` (two launches + cumsum), the CUB radix sort it calls, and
// `isect_offset_encode`.  Integer work: every output is bit-exact
// against oracle/torch_oracle.py::isect_tiles / isect_offset_encode.
#include "b2s_math.cuh"
#include "common.cuh"
#include "radix.cuh"

#ifdef B2S_SORT_CUB   // A/B switch: build with -DB2S_SORT_CUB to route the sorts through cub::DeviceRadixSort
#include <cub/device/device_radix_sort.cuh>
#endif

namespace b2s {

// -------------------------------------------------------------------------------------------------
// count (only used by the stage-level op; the training path gets tiles_per_gauss from stage 1)
// -------------------------------------------------------------------------------------------------
__global__ void isect_count_kernel(int64_t M, const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                   float tile_size, int tile_w, int row_lo, int row_hi, int32_t* __restrict__ tiles) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int r = radii[i];
    int cnt = 0;
    if (r > 0) {
        float2 m = means2d[i];
        int x0, y0, x1, y1;
        tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
        cnt = (x1 - x0) * (y1 - y0);
    }
    tiles[i] = cnt;
}

// -------------------------------------------------------------------------------------------------
// single-pass exclusive scan (decoupled look-back, dynamic tile ids so predecessors always run)
// -------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
// status word: bits 63..62 = flag (0 none, 1 tile aggregate, 2 inclusive prefix), bits 61..0 = value
constexpr unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VAL_MASK = (1ull << 62) - 1;

// Decoupled look-back, executed by one full warp: lane l inspects predecessor (tile - 1 - l), so a window of
// 32 CTAs costs one memory round trip instead of 32.  Publishes this CTA's aggregate first and its inclusive
// prefix at the end; returns the exclusive prefix.  CTA ids must be handed out dynamically (atomic counter)
// so that every predecessor is already running.
__device__ __forceinline__ unsigned long long lookback_warp(volatile unsigned long long* st, unsigned tile,
                                                            unsigned long long block_sum, int lane) {
    if (tile == 0) {
        if (lane == 0) st[0] = FLAG_INC | block_sum;
        return 0ull;
    }
    if (lane == 0) st[tile] = FLAG_AGG | block_sum;   // flag and value travel in one word: no fence needed
    unsigned long long excl = 0ull;
    long long base = (long long)tile - 1;
    while (true) {
        const long long k = base - lane;
        unsigned long long sv = 2ull << 62;                              // before tile 0: inclusive prefix 0
        if (k >= 0) sv = st[k];   // before tile 0: inclusive prefix 0
        const unsigned long long flag = sv & ~VAL_MASK;
        const unsigned ready = __ballot_sync(FULL, flag != 0ull);
        const unsigned inc = __ballot_sync(FULL, flag == FLAG_INC);
        const int p = inc ? (__ffs(inc) - 1) : 31;                    // nearest predecessor with an inclusive prefix
        const unsigned need = (p == 31) ? FULL : ((1u << (p + 1)) - 1u);
        if ((ready & need) != need) continue;                         // someone in the window has not published yet
        unsigned long long v = (lane <= p) ? (sv & VAL_MASK) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
        excl += v;
        if (inc) break;
        base -= 32;
    }
    if (lane == 0) st[tile] = FLAG_INC | (excl + block_sum);
    return excl;
}

template <bool GATHER>  // GATHER: scan in[idx[i]] (values taken in a permuted order)
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(int64_t M, const int32_t* __restrict__ in,
                                                            const int32_t* __restrict__ idx,
                                                            int32_t* __restrict__ out, int64_t* __restrict__ total,
                                                            unsigned long long* status, unsigned int* tile_counter) {
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    unsigned long long tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < M) ? (GATHER ? in[idx[base + k]] : in[base + k]) : 0;
        tsum += (unsigned long long)v[k];
    }
    // block exclusive scan of per-thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        unsigned long long x = s_warp[w];
        if (w < warp) warp_off += x;
        block_sum += x;
    }
    if (warp == 0) {
        const unsigned long long excl = lookback_warp(status, tile, block_sum, lane);
        if (lane == 0) {
            s_prefix = excl;
            if ((int64_t)(tile + 1) * SCAN_TILE >= M) *total = (int64_t)(excl + block_sum);
        }
    }
    __syncthreads();
    unsigned long long run = s_prefix + warp_off + (inc - tsum);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < M) out[base + k] = (int32_t)run;
        run += (unsigned long long)v[k];
    }
}

// -------------------------------------------------------------------------------------------------
// packed layout: indices of the visible rows (radius > 0), ascending -- camera-major, Gaussian-ascending, the row order
// of gsplat's packed mode -- by the same single-pass look-back scan (replaces torch.nonzero, a library kernel)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) select_visible_kernel(int64_t M, const int32_t* __restrict__ radii,
                                                                      int64_t* __restrict__ sel, int64_t cap, int64_t* __restrict__ total,
                                                                      unsigned long long* status, unsigned int* tile_counter) {
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    unsigned flags = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < M && radii[base + k] > 0) flags |= 1u << k;
    const unsigned long long tsum = __popc(flags);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        unsigned long long x = s_warp[w];
        if (w < warp) warp_off += x;
        block_sum += x;
    }
    if (warp == 0) {
        const unsigned long long excl = lookback_warp(status, tile, block_sum, lane);
        if (lane == 0) {
            s_prefix = excl;
            if ((int64_t)(tile + 1) * SCAN_TILE >= M) *total = (int64_t)(excl + block_sum);
        }
    }
    __syncthreads();
    int64_t run = (int64_t)(s_prefix + warp_off + (inc - tsum));
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if ((flags >> k) & 1u) {
            if (run < cap) sel[run] = base + k;
            ++run;
        }
}

// -------------------------------------------------------------------------------------------------
// fill: one key/value per covered tile, row-major inside the rectangle (y outer, x inner)
// key = cam << (32 + tile_bits) | tile << 32 | bits(depth)
// -------------------------------------------------------------------------------------------------
__global__ void isect_fill_kernel(int64_t M, int64_t N, const int64_t* __restrict__ camera_ids,
                                  const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                  const float* __restrict__ depths, const int32_t* __restrict__ cum, float tile_size,
                                  int tile_w, int row_lo, int row_hi, int tile_bits, int64_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    int r = radii[i];
    if (r <= 0) return;
    float2 m = means2d[i];
    int x0, y0, x1, y1;
    tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
    if (x1 <= x0 || y1 <= y0) return;
    int64_t cam = camera_ids ? camera_ids[i] : (i / N);
    uint64_t hi = (uint64_t)cam << tile_bits;
    uint64_t dbits = (uint64_t)__float_as_uint(depths[i]);
    int64_t o = cum[i];
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            uint64_t tile = (uint64_t)(y * tile_w + x);
            keys[o] = (int64_t)(((hi | tile) << 32) | dbits);
            vals[o] = (int32_t)i;
            ++o;
        }
}

// -------------------------------------------------------------------------------------------------
// offsets: first sorted index of every (camera, tile); empty tiles take the next tile's start
// -------------------------------------------------------------------------------------------------
__global__ void isect_offsets_kernel(int64_t n, const int64_t* __restrict__ keys, int n_tiles, int tile_bits,
                                     int64_t total_tiles, int32_t* __restrict__ offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t cur = (uint64_t)keys[i] >> 32;
    int64_t cur_lin = (int64_t)(cur >> tile_bits) * n_tiles + (int64_t)(cur & ((1ull << tile_bits) - 1));
    if (i == 0) {
        for (int64_t t = 0; t <= cur_lin; ++t) offsets[t] = 0;
    } else {
        uint64_t prev = (uint64_t)keys[i - 1] >> 32;
        if (prev != cur) {
            int64_t prev_lin = (int64_t)(prev >> tile_bits) * n_tiles + (int64_t)(prev & ((1ull << tile_bits) - 1));
            for (int64_t t = prev_lin + 1; t <= cur_lin; ++t) offsets[t] = (int32_t)i;
        }
    }
    if (i == n - 1)
        for (int64_t t = cur_lin + 1; t < total_tiles; ++t) offsets[t] = (int32_t)n;
}


// =================================================================================================
// Depth-first binning (the training path).
//
// The reference sorts n_isects 64-bit (camera | tile | depth) keys: 6 radix passes over 12-byte records.
// The same order is reached with far less traffic by sorting *Gaussians* by depth once (M 32-bit keys),
// emitting the tile intersections in that order, and stable-splitting them by (camera, tile) with two
// 8-bit radix passes over 8-byte records:
//     within a tile: ascending depth bits, ties by ascending flat index   == the reference's stable sort.
// isect_ids (the 64-bit keys) are never needed by compositing and are rebuilt on demand.
// =================================================================================================
__global__ void depth_keys_kernel(int64_t M, const float* __restrict__ depths, const int32_t* __restrict__ radii,
                                  uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    keys[i] = radii[i] > 0 ? __float_as_uint(depths[i]) : 0xFFFFFFFFu;  // culled rows sort last
    vals[i] = (int32_t)i;
}

// Expansion with a fused single-pass scan.  Every thread takes one depth-ordered Gaussian, recomputes its
// tile rectangle, the CTA scans the counts and obtains its global base by decoupled look-back (dynamic
// CTA ids, as scan_kernel); then each warp writes its Gaussians' (tile key, flat id) records with lanes
// striding over the warp's output range, so the stores are fully coalesced.
constexpr int EXP_THREADS = 256;
template <typename KeyT>
__global__ void __launch_bounds__(EXP_THREADS) expand_kernel(int64_t M, int64_t N, const int64_t* __restrict__ camera_ids,
                                                             const int32_t* __restrict__ sorted_ids,
                                                             const float2* __restrict__ means2d,
                                                             const int32_t* __restrict__ radii, float tile_size, int tile_w,
                                                             int row_lo, int row_hi, int tile_bits, unsigned long long* status,
                                                             unsigned int* cta_counter, KeyT* __restrict__ tile_keys,
                                                             int32_t* __restrict__ vals) {
    __shared__ int s_start[EXP_THREADS];   // output offset of each Gaussian relative to its warp's first
    __shared__ int s_x0[EXP_THREADS], s_w[EXP_THREADS], s_id[EXP_THREADS];
    __shared__ uint32_t s_base[EXP_THREADS];  // cam << tile_bits | (y0 * tile_w)
    __shared__ unsigned int s_cta;
    __shared__ int s_warp_tot[EXP_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_cta = atomicAdd(cta_counter, 1u);
    __syncthreads();
    const unsigned int cta = s_cta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int64_t j = (int64_t)cta * EXP_THREADS + threadIdx.x;
    int cnt = 0, id = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (j < M) {
        id = sorted_ids[j];
        int r = radii[id];
        if (r > 0) {
            float2 m = means2d[id];
            tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
            cnt = max(0, x1 - x0) * max(0, y1 - y0);
        }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (lane == 31) s_warp_tot[warp] = total;
    int64_t cam = 0;
    if (cnt > 0) cam = camera_ids ? camera_ids[id] : (id / N);
    s_start[threadIdx.x] = incl - cnt;
    s_x0[threadIdx.x] = x0;
    s_w[threadIdx.x] = max(1, x1 - x0);
    s_id[threadIdx.x] = id;
    s_base[threadIdx.x] = ((uint32_t)cam << tile_bits) + (uint32_t)(y0 * tile_w);
    __syncthreads();
    int warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < EXP_THREADS / 32; ++w) {
        int x = s_warp_tot[w];
        if (w < warp) warp_off += x;
        block_sum += x;
    }
    if (warp == 0) {
        const unsigned long long excl = lookback_warp(status, cta, (unsigned long long)block_sum, lane);
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    if (total == 0) return;
    const int64_t out0 = (int64_t)s_prefix + warp_off;
    for (int e = lane; e < total; e += 32) {
        // owner g = last Gaussian of the warp with start <= e (binary search over 32 entries)
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (s_start[wbase + lo + step] <= e) lo += step;
        const int g = wbase + lo;
        const int k = e - s_start[g];
        const int w = s_w[g];
        const int dy = k / w, dx = k - dy * w;
        tile_keys[out0 + e] = (KeyT)(s_base[g] + (uint32_t)(dy * tile_w + s_x0[g] + dx));
        vals[out0 + e] = s_id[g];
    }
}

// offsets[t] = lower bound of key(t) in the sorted (camera << tile_bits | tile) keys: one thread per tile,
// ~log2(n) dependent loads each -- tiny next to a pass over the keys.
template <typename KeyT>
__global__ void offsets_search_kernel(int64_t n, const KeyT* __restrict__ keys, int n_tiles, int tile_bits,
                                      int64_t total_tiles, int32_t* __restrict__ offsets) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_tiles) return;
    const uint32_t want = ((uint32_t)(t / n_tiles) << tile_bits) | (uint32_t)(t % n_tiles);
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((uint32_t)keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    offsets[t] = (int32_t)lo;
}

template <typename KeyT>
__global__ void isect_ids_kernel(int64_t n, const KeyT* __restrict__ tile_keys, const int32_t* __restrict__ flat,
                                 const float* __restrict__ depths, int64_t* __restrict__ isect_ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    isect_ids[i] = (int64_t)(((uint64_t)tile_keys[i] << 32) | (uint64_t)__float_as_uint(depths[flat[i]]));
}


// =================================================================================================
// Two-level binning (the path rasterization() runs): coarse bins of B2S_COARSE x B2S_COARSE tiles.
//
// One record per (depth-ordered row, coarse bin it overlaps):
//     rec = mask_y << 24 | mask_x << 16 | coarse key,   coarse key = (cam * bins_h + by) * bins_w + bx
// mask_x bit i <=> the row's tile rectangle covers tile column 8*bx + i.  The records are stable-sorted by
// coarse key (depth order kept inside a bin); the compositing kernels expand the per-tile lists on the fly.
// =================================================================================================
constexpr int CB = B2S_COARSE;
static_assert(CB == 8, "the record format packs two 8-bit tile masks");

__device__ __forceinline__ int coarse_span(int t0, int t1) { return (t1 + CB - 1) / CB - t0 / CB; }   // t1 > t0

__global__ void coarse_count_kernel(int64_t M, const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                    float tile_size, int tile_w, int row_lo, int row_hi, unsigned long long* __restrict__ total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (i < M) {
        const int r = radii[i];
        if (r > 0) {
            const float2 m = means2d[i];
            int x0, y0, x1, y1;
            tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
            if (x1 > x0 && y1 > y0) cnt = coarse_span(x0, x1) * coarse_span(y0, y1);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(total, (unsigned long long)cnt);
}

// Expansion in depth order with a fused single-pass scan (same structure as expand_kernel): a warp writes
// the records of its 32 rows with lanes striding over the warp's output range (coalesced).
__global__ void __launch_bounds__(EXP_THREADS) coarse_expand_kernel(int64_t M, int64_t N, const int64_t* __restrict__ camera_ids,
                                                                    const int32_t* __restrict__ sorted_ids,
                                                                    const float2* __restrict__ means2d,
                                                                    const int32_t* __restrict__ radii, float tile_size,
                                                                    int tile_w, int row_lo, int row_hi, int bins_w, int bins_h,
                                                                    unsigned long long* status, unsigned int* cta_counter,
                                                                    uint32_t* __restrict__ recs, int32_t* __restrict__ vals) {
    __shared__ int s_start[EXP_THREADS];
    __shared__ int s_id[EXP_THREADS];
    __shared__ uint32_t s_xs[EXP_THREADS], s_ys[EXP_THREADS];   // x0 | x1 << 16, y0 | y1 << 16 (tile units)
    __shared__ uint32_t s_cam[EXP_THREADS];                     // cam * bins_h * bins_w
    __shared__ unsigned int s_cta;
    __shared__ int s_warp_tot[EXP_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_cta = atomicAdd(cta_counter, 1u);
    __syncthreads();
    const unsigned int cta = s_cta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int64_t j = (int64_t)cta * EXP_THREADS + threadIdx.x;
    int cnt = 0, id = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (j < M) {
        id = sorted_ids[j];
        const int r = radii[id];
        if (r > 0) {
            const float2 m = means2d[id];
            tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
            if (x1 > x0 && y1 > y0) cnt = coarse_span(x0, x1) * coarse_span(y0, y1);
        }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (lane == 31) s_warp_tot[warp] = total;
    int64_t cam = 0;
    if (cnt > 0) cam = camera_ids ? camera_ids[id] : (id / N);
    s_start[threadIdx.x] = incl - cnt;
    s_id[threadIdx.x] = id;
    s_xs[threadIdx.x] = (uint32_t)x0 | ((uint32_t)x1 << 16);
    s_ys[threadIdx.x] = (uint32_t)y0 | ((uint32_t)y1 << 16);
    s_cam[threadIdx.x] = (uint32_t)cam * (uint32_t)(bins_w * bins_h);
    __syncthreads();
    int warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < EXP_THREADS / 32; ++w) {
        int x = s_warp_tot[w];
        if (w < warp) warp_off += x;
        block_sum += x;
    }
    if (warp == 0) {
        const unsigned long long excl = lookback_warp(status, cta, (unsigned long long)block_sum, lane);
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    if (total == 0) return;
    const int64_t out0 = (int64_t)s_prefix + warp_off;
    for (int e = lane; e < total; e += 32) {
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (s_start[wbase + lo + step] <= e) lo += step;
        const int g = wbase + lo;
        const int k = e - s_start[g];
        const int gx0 = (int)(s_xs[g] & 0xFFFFu), gx1 = (int)(s_xs[g] >> 16);
        const int gy0 = (int)(s_ys[g] & 0xFFFFu), gy1 = (int)(s_ys[g] >> 16);
        const int bx0 = gx0 / CB, by0 = gy0 / CB;
        const int w = coarse_span(gx0, gx1);
        const int dy = k / w, dx = k - dy * w;
        const int bx = bx0 + dx, by = by0 + dy;
        // tile columns / rows of this coarse bin covered by the rectangle, as bit masks
        const int lx0 = max(gx0 - bx * CB, 0), lx1 = min(gx1 - bx * CB, CB);
        const int ly0 = max(gy0 - by * CB, 0), ly1 = min(gy1 - by * CB, CB);
        const uint32_t mx = ((1u << lx1) - 1u) & ~((1u << lx0) - 1u);
        const uint32_t my = ((1u << ly1) - 1u) & ~((1u << ly0) - 1u);
        recs[out0 + e] = (my << 24) | (mx << 16) | (s_cam[g] + (uint32_t)(by * bins_w + bx));
        vals[out0 + e] = s_id[g];
    }
}

// ---- two-kernel expansion (no chained scan) -------------------------------------------------------------------------
// coarse_expand_kernel above spends half its time at the barrier behind the decoupled look-back (ncu: 53 % stall_barrier).
// Here the per-row work is split so that no CTA waits on another:
//   coarse_rects_kernel : gathers each depth-ordered row once, stores its tile rectangle (x0|x1<<16, y0|y1<<16) in depth
//                         order and the CTA's record count;
//   coarse_scan_kernel  : exclusive scan of the CTA counts (one CTA);
//   coarse_emit_kernel  : streams the rectangles (coalesced, no gather), scans inside the CTA and writes the records.
__global__ void __launch_bounds__(EXP_THREADS) coarse_rects_kernel(int64_t M, const int32_t* __restrict__ sorted_ids,
                                                                   const float2* __restrict__ means2d,
                                                                   const int32_t* __restrict__ radii, float tile_size, int tile_w,
                                                                   int row_lo, int row_hi, uint32_t* __restrict__ xs,
                                                                   uint32_t* __restrict__ ys, uint32_t* __restrict__ cta_tot) {
    __shared__ int s_w[EXP_THREADS / 32];
    const int64_t j = (int64_t)blockIdx.x * EXP_THREADS + threadIdx.x;
    int cnt = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (j < M) {
        const int id = sorted_ids[j];
        const int r = radii[id];
        if (r > 0) {
            const float2 m = means2d[id];
            tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
            if (x1 > x0 && y1 > y0) cnt = coarse_span(x0, x1) * coarse_span(y0, y1);
            else x0 = x1 = y0 = y1 = 0;
        }
        xs[j] = (uint32_t)x0 | ((uint32_t)x1 << 16);
        ys[j] = (uint32_t)y0 | ((uint32_t)y1 << 16);
    }
    cnt = __reduce_add_sync(FULL, cnt);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < EXP_THREADS / 32; ++w) t += s_w[w];
        cta_tot[blockIdx.x] = (uint32_t)t;
    }
}

// exclusive scan of the CTA counts, one CTA: coalesced 1024-wide sweeps with a running carry (the chunk-per-thread form
// read with a stride of n/1024 words and took 20 us for 22 K counts)
__global__ void __launch_bounds__(1024) coarse_scan_kernel(uint32_t* __restrict__ cta_tot, int64_t n) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t k = base + threadIdx.x;
        const unsigned v = k < n ? cta_tot[k] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_w[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(FULL, wi, d);
                if (lane >= d) wi += t;
            }
            s_w[lane] = wi - w;                      // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned carry = s_carry;
        if (k < n) cta_tot[k] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_w[31] + incl;   // total of this sweep (thread 1023 holds the last inclusive)
        __syncthreads();
    }
}

__global__ void __launch_bounds__(EXP_THREADS) coarse_emit_kernel(int64_t M, int64_t N, const int64_t* __restrict__ camera_ids,
                                                                  const int32_t* __restrict__ sorted_ids,
                                                                  const uint32_t* __restrict__ xs, const uint32_t* __restrict__ ys,
                                                                  const uint32_t* __restrict__ cta_base, int bins_w, int bins_h,
                                                                  uint32_t* __restrict__ recs, int32_t* __restrict__ vals) {
    __shared__ int s_start[EXP_THREADS];
    __shared__ int s_id[EXP_THREADS];
    __shared__ uint32_t s_xs[EXP_THREADS], s_ys[EXP_THREADS];
    __shared__ uint32_t s_cam[EXP_THREADS];
    __shared__ int s_warp_tot[EXP_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const int64_t j = (int64_t)blockIdx.x * EXP_THREADS + threadIdx.x;
    uint32_t px = 0, py = 0;
    int id = 0, cnt = 0;
    if (j < M) {
        px = xs[j]; py = ys[j];
        id = sorted_ids[j];
        const int x0 = (int)(px & 0xFFFFu), x1 = (int)(px >> 16), y0 = (int)(py & 0xFFFFu), y1 = (int)(py >> 16);
        if (x1 > x0 && y1 > y0) cnt = coarse_span(x0, x1) * coarse_span(y0, y1);
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (lane == 31) s_warp_tot[warp] = total;
    int64_t cam = 0;
    if (cnt > 0) cam = camera_ids ? camera_ids[id] : (id / N);
    s_start[threadIdx.x] = incl - cnt;
    s_id[threadIdx.x] = id;
    s_xs[threadIdx.x] = px;
    s_ys[threadIdx.x] = py;
    s_cam[threadIdx.x] = (uint32_t)cam * (uint32_t)(bins_w * bins_h);
    __syncthreads();
    if (total == 0) return;
    int warp_off = 0;
#pragma unroll
    for (int w = 0; w < EXP_THREADS / 32; ++w)
        if (w < warp) warp_off += s_warp_tot[w];
    const int64_t out0 = (int64_t)cta_base[blockIdx.x] + warp_off;
    for (int e = lane; e < total; e += 32) {
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (s_start[wbase + lo + step] <= e) lo += step;
        const int g = wbase + lo;
        const int k = e - s_start[g];
        const int gx0 = (int)(s_xs[g] & 0xFFFFu), gx1 = (int)(s_xs[g] >> 16);
        const int gy0 = (int)(s_ys[g] & 0xFFFFu), gy1 = (int)(s_ys[g] >> 16);
        const int bx0 = gx0 / CB, by0 = gy0 / CB;
        const int w = coarse_span(gx0, gx1);
        const int dy = k / w, dx = k - dy * w;
        const int bx = bx0 + dx, by = by0 + dy;
        const int lx0 = max(gx0 - bx * CB, 0), lx1 = min(gx1 - bx * CB, CB);
        const int ly0 = max(gy0 - by * CB, 0), ly1 = min(gy1 - by * CB, CB);
        const uint32_t mx = ((1u << lx1) - 1u) & ~((1u << lx0) - 1u);
        const uint32_t my = ((1u << ly1) - 1u) & ~((1u << ly0) - 1u);
        recs[out0 + e] = (my << 24) | (mx << 16) | (s_cam[g] + (uint32_t)(by * bins_w + bx));
        vals[out0 + e] = s_id[g];
    }
}

// ---- fused path (b2s_depth_order_rects + b2s_coarse_bin_rects) ------------------------------------------------------
// coarse_rects_kernel above is a pass of dependent gathers (sorted id -> radius, mean: two 32-byte sectors per row, DRAM
// bound at ~65 us for 5.8 M rows wherever it runs -- fused into the sort's last pass it cost the same).  Here the histogram
// kernel of the depth sort, which streams the rows in row order, derives every row's tile rectangle from coalesced loads and
// stores it, packed into 32 bits, by row (x0 | x1 << 8 | y0 << 16 | y1 << 24 in tile units: grids up to 255 x 255 tiles,
// i.e. 4080 x 4080 pixels); the expansion then gathers ONE 4-byte word per depth-ordered row.  The last pass does not
// write the sorted keys, which nobody reads.
struct RectSide {
    static constexpr bool kActive = true;
    const float2* means2d;
    const int32_t* radii;
    float tile_size, inv_tile;       // inv_tile > 0: tile_size is a power of two and x / tile_size == x * inv_tile exactly
    int tile_w, row_lo, row_hi;
    uint32_t* rects;                 // [M], by row
    __device__ __forceinline__ void operator()(int64_t /*i*/, int32_t id) const {
        const int r = __ldg(radii + id);
        const float2 m = __ldg(means2d + id);
        int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        if (r > 0) {
            if (inv_tile > 0.f) {    // the same values as tile_rect, without three IEEE divisions per row
                const float mx = __fmul_rn(m.x, inv_tile), my = __fmul_rn(m.y, inv_tile), rr = __fmul_rn((float)r, inv_tile);
                x0 = iclamp((int)floorf(__fsub_rn(mx, rr)), 0, tile_w); y0 = iclamp((int)floorf(__fsub_rn(my, rr)), row_lo, row_hi);
                x1 = iclamp((int)ceilf(__fadd_rn(mx, rr)), 0, tile_w); y1 = iclamp((int)ceilf(__fadd_rn(my, rr)), row_lo, row_hi);
            } else {
                tile_rect(m.x, m.y, r, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
            }
            if (!(x1 > x0 && y1 > y0)) x0 = x1 = y0 = y1 = 0;
        }
        rects[id] = (uint32_t)x0 | ((uint32_t)x1 << 8) | ((uint32_t)y0 << 16) | ((uint32_t)y1 << 24);
    }
};
struct IdStore {
    int32_t* sorted_ids;
    __device__ __forceinline__ void operator()(bool active, unsigned pos, uint32_t /*key*/, int32_t id) const {
        if (active) sorted_ids[pos] = id;
    }
};

// Expansion of the depth-ordered rows into coarse-bin records, one kernel: every CTA takes EMIT_ROWS rows, gathers their
// packed rectangles (one word per row: four times the rows per CTA and half the sectors per row of the fused look-back
// kernel that stalled behind its chained scan) and obtains its output base by decoupled look-back.  A row overlaps 1.4 bins on average, so every lane
// writes the (up to EMIT_DIRECT) records of its own row itself -- the slots of neighbouring lanes are neighbours, the
// stores coalesce -- and only the rare larger rows are expanded by the whole warp.  The digits of the bin sort are counted
// on the way (shared-memory histogram, flushed once per CTA), which replaces the sort's own histogram pass.
constexpr int EMIT_GROUPS = 4;                               // 32-row groups per warp
constexpr int EMIT_ROWS = EXP_THREADS * EMIT_GROUPS;         // rows per CTA
constexpr int EMIT_DIRECT = 4;

__device__ __forceinline__ uint32_t coarse_record(int gx0, int gx1, int gy0, int gy1, int bx, int by, uint32_t cam_base, int bins_w) {
    const int lx0 = max(gx0 - bx * CB, 0), lx1 = min(gx1 - bx * CB, CB);
    const int ly0 = max(gy0 - by * CB, 0), ly1 = min(gy1 - by * CB, CB);
    const uint32_t mx = ((1u << lx1) - 1u) & ~((1u << lx0) - 1u);
    const uint32_t my = ((1u << ly1) - 1u) & ~((1u << ly0) - 1u);
    return (my << 24) | (mx << 16) | (cam_base + (uint32_t)(by * bins_w + bx));
}

__global__ void __launch_bounds__(EXP_THREADS, 5) coarse_emit_lb_kernel(int64_t M, int64_t N, const int64_t* __restrict__ camera_ids,
                                                                     const int32_t* __restrict__ sorted_ids,
                                                                     const uint32_t* __restrict__ rects, int bins_w, int bins_h,
                                                                     unsigned long long* status, unsigned int* cta_counter,
                                                                     const radix::Plan pl, unsigned long long* __restrict__ hist,
                                                                     uint32_t* __restrict__ recs, int32_t* __restrict__ vals) {
    __shared__ unsigned int s_cta;
    __shared__ int s_warp_tot[EXP_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned s_hist[2 * radix::RADIX];
    if (threadIdx.x == 0) s_cta = atomicAdd(cta_counter, 1u);
    for (int i = threadIdx.x; i < 2 * radix::RADIX; i += EXP_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const unsigned int cta = s_cta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)cta * EMIT_ROWS + (int64_t)warp * (32 * EMIT_GROUPS) + lane;
    uint32_t rc[EMIT_GROUPS];
    int id[EMIT_GROUPS], cnt[EMIT_GROUPS], excl[EMIT_GROUPS];
#pragma unroll
    for (int g = 0; g < EMIT_GROUPS; ++g) {
        const int64_t j = row0 + g * 32;
        rc[g] = 0u; id[g] = 0;
        if (j < M) {
            id[g] = sorted_ids[j];
            rc[g] = __ldg(rects + id[g]);
        }
    }
    int run = 0;                                               // records of the warp's earlier groups
#pragma unroll
    for (int g = 0; g < EMIT_GROUPS; ++g) {
        const int x0 = (int)(rc[g] & 0xFFu), x1 = (int)((rc[g] >> 8) & 0xFFu), y0 = (int)((rc[g] >> 16) & 0xFFu), y1 = (int)(rc[g] >> 24);
        cnt[g] = (x1 > x0 && y1 > y0) ? coarse_span(x0, x1) * coarse_span(y0, y1) : 0;
        int incl = cnt[g];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        excl[g] = run + incl - cnt[g];
        run += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) s_warp_tot[warp] = run;
    __syncthreads();
    int warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < EXP_THREADS / 32; ++w) {
        const int x = s_warp_tot[w];
        if (w < warp) warp_off += x;
        block_sum += x;
    }
    // (a look-back by the whole CTA -- 256 predecessors per round trip -- measured slower: 119 vs 98 us; the walk is short
    // in practice and the wider poll only adds traffic, profiles/r02_variants_c15.txt)
    if (warp == 0) {
        const unsigned long long e = lookback_warp(status, cta, (unsigned long long)block_sum, lane);
        if (lane == 0) s_prefix = e;
    }
    __syncthreads();
    const int64_t out0 = (int64_t)s_prefix + warp_off;
    const unsigned m0 = (1u << pl.bits[0]) - 1u, m1 = pl.passes > 1 ? (1u << pl.bits[1]) - 1u : 0u;
    const int sh1 = pl.shift[1];
    const uint32_t per_cam = (uint32_t)(bins_w * bins_h);
    auto put = [&](int64_t slot, uint32_t rec, int v) {
        recs[slot] = rec;
        vals[slot] = v;
        atomicAdd(&s_hist[rec & m0], 1u);
        if (m1) atomicAdd(&s_hist[radix::RADIX + ((rec >> sh1) & m1)], 1u);
    };
#pragma unroll
    for (int g = 0; g < EMIT_GROUPS; ++g) {
        const int gx0 = (int)(rc[g] & 0xFFu), gx1 = (int)((rc[g] >> 8) & 0xFFu), gy0 = (int)((rc[g] >> 16) & 0xFFu), gy1 = (int)(rc[g] >> 24);
        uint32_t cam_base = 0u;                                 // (32-bit division, and none at all with one camera)
        if (cnt[g] > 0 && (camera_ids || M > N)) cam_base = (camera_ids ? (uint32_t)camera_ids[id[g]] : (uint32_t)id[g] / (uint32_t)N) * per_cam;
        const int bx0 = gx0 / CB, by0 = gy0 / CB, w = cnt[g] > 0 ? coarse_span(gx0, gx1) : 1;
        // small rows: each lane writes its own records
        if (cnt[g] <= EMIT_DIRECT) {
            int dx = 0, dy = 0;
            for (int k = 0; k < cnt[g]; ++k) {
                put(out0 + excl[g] + k, coarse_record(gx0, gx1, gy0, gy1, bx0 + dx, by0 + dy, cam_base, bins_w), id[g]);
                if (++dx == w) { dx = 0; ++dy; }
            }
        }
        // large rows: one at a time, the warp strides over the row's bins
        unsigned big = __ballot_sync(FULL, cnt[g] > EMIT_DIRECT);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t r = __shfl_sync(FULL, rc[g], src);
            const int n = __shfl_sync(FULL, cnt[g], src), v = __shfl_sync(FULL, id[g], src);
            const int ex = __shfl_sync(FULL, excl[g], src), ww = __shfl_sync(FULL, w, src);
            const uint32_t cb = __shfl_sync(FULL, cam_base, src);
            const int rx0 = (int)(r & 0xFFu), rx1 = (int)((r >> 8) & 0xFFu), ry0 = (int)((r >> 16) & 0xFFu), ry1 = (int)(r >> 24);
            for (int k = lane; k < n; k += 32) {
                const int dy = k / ww, dx = k - dy * ww;
                put(out0 + ex + k, coarse_record(rx0, rx1, ry0, ry1, rx0 / CB + dx, ry0 / CB + dy, cb, bins_w), v);
            }
        }
    }
    __syncthreads();
    hist += (size_t)(cta % radix::HIST_COPIES) * sizeof(uint32_t) * radix::RADIX;      // [copies][4 passes][256]
    for (int i = threadIdx.x; i < pl.passes * radix::RADIX; i += EXP_THREADS)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

// one digit pass sorts the records: the scanned digit histogram of that pass IS the offset table
__global__ void coarse_offsets_from_hist_kernel(const unsigned long long* __restrict__ scanned, int64_t n_bins, int64_t n_records,
                                                int32_t* __restrict__ offsets) {
    for (int64_t k = threadIdx.x; k <= n_bins; k += blockDim.x) offsets[k] = k < n_bins ? (int32_t)scanned[k] : (int32_t)n_records;
}

// bin_offsets[k] = lower bound of coarse key k in the sorted records, k = 0 .. n_bins (inclusive: the end)
__global__ void coarse_offsets_kernel(int64_t n, const uint32_t* __restrict__ recs, int64_t n_bins, int32_t* __restrict__ offsets) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_bins) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(recs[mid] & 0xFFFFu) < k) lo = mid + 1; else hi = mid;
    }
    offsets[k] = (int32_t)lo;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// scratch bytes of one (key, int32) pair sort, and the sort itself: the hand-written onesweep of radix.cuh
template <typename KeyT>
static size_t cub_pairs_bytes(int64_t n) {
#ifdef B2S_SORT_CUB
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const KeyT*)nullptr, (KeyT*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, n > 0 ? n : 1, 0, (int)sizeof(KeyT) * 8, (cudaStream_t)0);
    return bytes;
#else
    return radix::workspace_bytes<KeyT>(n > 0 ? n : 1);
#endif
}
static size_t cub_pairs32_bytes(int64_t n) { return cub_pairs_bytes<uint32_t>(n); }

template <typename KeyT>
static int sort_pairs(void* tmp, size_t tmp_bytes, const KeyT* keys_in, KeyT* keys_out, const int32_t* vals_in,
                      int32_t* vals_out, int64_t n, int end_bit, cudaStream_t st) {
#ifdef B2S_SORT_CUB
    B2S_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, st));
    return 0;
#else
    (void)tmp_bytes;
    return radix::sort_pairs<KeyT>(n, end_bit, keys_in, vals_in, keys_out, vals_out, tmp, st);
#endif
}

static int bit_length(int64_t v) {
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b;
}

}  // namespace b2s

using namespace b2s;

// tile-row window of a call: NULL = the whole grid
static inline bool row_window(const b2s_tile_rows* rows, int tile_h, int& lo, int& hi) {
    lo = 0; hi = tile_h;
    if (rows) { lo = rows->row_lo; hi = rows->row_hi; }
    return lo >= 0 && hi >= lo && hi <= tile_h;
}

extern "C" int b2s_isect_count(int64_t M, const float* means2d, const int32_t* radii, int32_t tile_size, int32_t tile_w,
                               int32_t tile_h, int32_t* tiles_per_gauss, const b2s_tile_rows* rows, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0, "bad sizes");
    int row_lo, row_hi;
    B2S_REQUIRE(row_window(rows, tile_h, row_lo, row_hi), "bad tile-row window");
    if (M == 0) return 0;
    B2S_REQUIRE(means2d && radii && tiles_per_gauss, "null pointer");
    isect_count_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(
        M, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_w, row_lo, row_hi, tiles_per_gauss);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t b2s_isect_scan_workspace_bytes(int64_t M) {
    int64_t tiles = cdiv(M > 0 ? M : 1, SCAN_TILE);
    return 16 + tiles * 8;  // [tile counter, pad] + one status word per tile
}

extern "C" int b2s_isect_scan(int64_t M, const int32_t* tiles_per_gauss, int32_t* cum_tiles, int64_t* total,
                              void* workspace, int64_t workspace_bytes, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && total != nullptr, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        B2S_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return 0;
    }
    B2S_REQUIRE(tiles_per_gauss && cum_tiles && workspace, "null pointer");
    if (workspace_bytes < b2s_isect_scan_workspace_bytes(M)) {
        set_error("b2s_isect_scan: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
    int64_t tiles = cdiv(M, SCAN_TILE);
    B2S_CUDA(cudaMemsetAsync(workspace, 0, (size_t)(16 + tiles * 8), st));
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 16);
    scan_kernel<false><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(M, tiles_per_gauss, nullptr, cum_tiles, total, status, counter);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_select_visible(int64_t M, const int32_t* radii, int64_t* sel, int64_t cap, int64_t* total, void* workspace,
                                  int64_t workspace_bytes, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && cap >= 0 && total != nullptr, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        B2S_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
        return 0;
    }
    B2S_REQUIRE(radii && (sel || cap == 0) && workspace, "null pointer");
    if (workspace_bytes < b2s_isect_scan_workspace_bytes(M)) {
        set_error("b2s_select_visible: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
    const int64_t tiles = cdiv(M, SCAN_TILE);
    B2S_CUDA(cudaMemsetAsync(workspace, 0, (size_t)(16 + tiles * 8), st));
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 16);
    select_visible_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(M, radii, sel, cap, total, status, counter);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_isect_fill(int64_t M, int64_t N, const int64_t* camera_ids, const float* means2d, const int32_t* radii,
                              const float* depths, const int32_t* cum_tiles, int32_t tile_size, int32_t tile_w,
                              int32_t tile_h, int64_t* isect_ids, int32_t* flatten_ids, const b2s_tile_rows* rows,
                              b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0, "bad sizes");
    int row_lo, row_hi;
    B2S_REQUIRE(row_window(rows, tile_h, row_lo, row_hi), "bad tile-row window");
    if (M == 0) return 0;
    B2S_REQUIRE(camera_ids != nullptr || N > 0, "dense rows need N");
    B2S_REQUIRE(means2d && radii && depths && cum_tiles, "null pointer");
    int tile_bits = bit_length((int64_t)tile_w * tile_h);
    isect_fill_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(
        M, N, camera_ids, reinterpret_cast<const float2*>(means2d), radii, depths, cum_tiles, (float)tile_size, tile_w,
        row_lo, row_hi, tile_bits, isect_ids, flatten_ids);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t b2s_isect_sort_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    return (int64_t)cub_pairs_bytes<uint64_t>(n) + 256;
}

extern "C" int b2s_isect_sort(int64_t n, int32_t key_bits, const int64_t* keys_in, const int32_t* vals_in, int64_t* keys_out,
                              int32_t* vals_out, void* workspace, int64_t workspace_bytes, b2s_stream_t stream) {
    B2S_REQUIRE(n >= 0 && key_bits > 0 && key_bits <= 64, "bad sizes");
    if (n == 0) return 0;
    B2S_REQUIRE(keys_in && vals_in && keys_out && vals_out && workspace, "null pointer");
    if (workspace_bytes < b2s_isect_sort_workspace_bytes(n)) {
        set_error("b2s_isect_sort: workspace too small");
        return B2S_E_WORKSPACE;
    }
    void* w = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    return sort_pairs<uint64_t>(w, (size_t)workspace_bytes - 256, reinterpret_cast<const uint64_t*>(keys_in),
                                reinterpret_cast<uint64_t*>(keys_out), vals_in, vals_out, n, key_bits, (cudaStream_t)stream);
}

extern "C" int b2s_isect_offsets(int64_t n, const int64_t* keys, int32_t C, int32_t tile_w, int32_t tile_h,
                                 int32_t* offsets, b2s_stream_t stream) {
    B2S_REQUIRE(n >= 0 && C >= 0 && tile_w > 0 && tile_h > 0 && offsets, "bad args");
    int n_tiles = tile_w * tile_h;
    int64_t total_tiles = (int64_t)C * n_tiles;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (total_tiles > 0) B2S_CUDA(cudaMemsetAsync(offsets, 0, (size_t)total_tiles * 4, st));
        return 0;
    }
    B2S_REQUIRE(keys != nullptr, "null keys");
    int tile_bits = bit_length(n_tiles);
    isect_offsets_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(n, keys, n_tiles, tile_bits, total_tiles, offsets);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// depth-first binning entry points
// ---------------------------------------------------------------------------------------------------
extern "C" int b2s_isect_depth_keys(int64_t M, const float* depths, const int32_t* radii, uint32_t* keys, int32_t* vals,
                                    b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0, "bad size");
    if (M == 0) return 0;
    B2S_REQUIRE(depths && radii && keys && vals, "null pointer");
    depth_keys_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(M, depths, radii, keys, vals);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t b2s_depth_order_workspace_bytes(int64_t M) {
    if (M <= 0) return 256;
    return (int64_t)(align256((size_t)M * 4) + align256(cub_pairs32_bytes(M)));  // sorted keys (discarded) + CUB temp
}

extern "C" int b2s_depth_order(int64_t M, const uint32_t* depth_keys, const int32_t* row_ids, int32_t* sorted_ids,
                               void* workspace, int64_t workspace_bytes, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) return 0;
    B2S_REQUIRE(depth_keys && row_ids && sorted_ids && workspace, "null pointer");
    if (workspace_bytes < b2s_depth_order_workspace_bytes(M)) {
        set_error("b2s_depth_order: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    char* w = reinterpret_cast<char*>(workspace);
    uint32_t* keys_sorted = reinterpret_cast<uint32_t*>(w);
    w += align256((size_t)M * 4);
    size_t cub_bytes = cub_pairs32_bytes(M);
    return sort_pairs<uint32_t>(w, cub_bytes, depth_keys, keys_sorted, row_ids, sorted_ids, M, 32, st);
}

extern "C" int32_t b2s_tile_key_bytes(int32_t C, int32_t tile_w, int32_t tile_h) {
    return (bit_length((int64_t)tile_w * tile_h) + bit_length(C) <= 16) ? 2 : 4;
}

extern "C" int64_t b2s_tile_bin_workspace_bytes(int64_t M, int64_t n_isects) {
    if (n_isects <= 0 || M <= 0) return 256;
    // unsorted keys + unsorted vals + CUB temp + look-back status (counter + one word per CTA)
    size_t cub = cub_pairs32_bytes(n_isects);
    size_t cub16 = cub_pairs_bytes<uint16_t>(n_isects);
    return (int64_t)(2 * align256((size_t)n_isects * 4) + align256(cub > cub16 ? cub : cub16) +
                     align256(16 + (size_t)cdiv(M, EXP_THREADS) * 8));
}

template <typename KeyT>
static int tile_bin_impl(int64_t M, int64_t N, const int64_t* camera_ids, int64_t n_isects, const int32_t* sorted_ids,
                         const float* means2d, const int32_t* radii, int tile_size, int tile_w, int tile_h, int row_lo, int row_hi, int tile_bits,
                         int key_bits, int64_t total_tiles, void* tile_keys_sorted, int32_t* flatten_ids, int32_t* offsets,
                         char* w, cudaStream_t st) {
    const int n_tiles = tile_w * tile_h;
    KeyT* keys_u = reinterpret_cast<KeyT*>(w);
    w += align256((size_t)n_isects * 4);
    int32_t* vals_u = reinterpret_cast<int32_t*>(w);
    w += align256((size_t)n_isects * 4);
    size_t cub_bytes = cub_pairs_bytes<KeyT>(n_isects);
    void* cub_tmp = w;
    size_t cub32 = cub_pairs32_bytes(n_isects), cub16 = cub_pairs_bytes<uint16_t>(n_isects);
    w += align256(cub32 > cub16 ? cub32 : cub16);
    const int64_t ctas = cdiv(M, EXP_THREADS);
    B2S_CUDA(cudaMemsetAsync(w, 0, (size_t)(16 + ctas * 8), st));
    unsigned int* counter = reinterpret_cast<unsigned int*>(w);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(w + 16);
    expand_kernel<KeyT><<<(unsigned)ctas, EXP_THREADS, 0, st>>>(
        M, N, camera_ids, sorted_ids, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_w, row_lo, row_hi,
        tile_bits, status, counter, keys_u, vals_u);
    B2S_LAUNCH_CHECK();
    {
        int rc = sort_pairs<KeyT>(cub_tmp, cub_bytes, keys_u, reinterpret_cast<KeyT*>(tile_keys_sorted), vals_u, flatten_ids,
                                  n_isects, key_bits, st);
        if (rc != 0) return rc;
    }
    offsets_search_kernel<KeyT><<<(unsigned)cdiv(total_tiles, 128), 128, 0, st>>>(
        n_isects, reinterpret_cast<const KeyT*>(tile_keys_sorted), n_tiles, tile_bits, total_tiles, offsets);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_tile_bin(int64_t M, int64_t N, const int64_t* camera_ids, int64_t n_isects, const int32_t* sorted_ids,
                            const float* means2d, const int32_t* radii, int32_t C, int32_t tile_size, int32_t tile_w,
                            int32_t tile_h, void* tile_keys_sorted, int32_t* flatten_ids, int32_t* offsets,
                            void* workspace, int64_t workspace_bytes, const b2s_tile_rows* rows, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && n_isects >= 0 && C >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0 && offsets, "bad args");
    int row_lo, row_hi;
    B2S_REQUIRE(row_window(rows, tile_h, row_lo, row_hi), "bad tile-row window");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = tile_w * tile_h;
    const int64_t total_tiles = (int64_t)C * n_tiles;
    if (n_isects == 0 || M == 0) {
        if (total_tiles > 0) B2S_CUDA(cudaMemsetAsync(offsets, 0, (size_t)total_tiles * 4, st));
        return 0;
    }
    B2S_REQUIRE(camera_ids != nullptr || N > 0, "dense rows need N");
    B2S_REQUIRE(sorted_ids && means2d && radii && tile_keys_sorted && flatten_ids && workspace, "null pointer");
    const int tile_bits = bit_length(n_tiles);
    const int key_bits = tile_bits + bit_length(C);
    B2S_REQUIRE(key_bits <= 32, "camera x tile ids exceed 32 bits");
    if (workspace_bytes < b2s_tile_bin_workspace_bytes(M, n_isects)) {
        set_error("b2s_tile_bin: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    char* w = reinterpret_cast<char*>(workspace);
    if (key_bits <= 16)
        return tile_bin_impl<uint16_t>(M, N, camera_ids, n_isects, sorted_ids, means2d, radii, tile_size, tile_w, tile_h,
                                       row_lo, row_hi, tile_bits, key_bits, total_tiles, tile_keys_sorted, flatten_ids, offsets, w, st);
    return tile_bin_impl<uint32_t>(M, N, camera_ids, n_isects, sorted_ids, means2d, radii, tile_size, tile_w, tile_h,
                                   row_lo, row_hi, tile_bits, key_bits, total_tiles, tile_keys_sorted, flatten_ids, offsets, w, st);
}

extern "C" int b2s_isect_ids_from_bins(int64_t n_isects, int32_t key_bytes, const void* tile_keys_sorted,
                                       const int32_t* flatten_ids, const float* depths, int64_t* isect_ids,
                                       b2s_stream_t stream) {
    B2S_REQUIRE(n_isects >= 0 && (key_bytes == 2 || key_bytes == 4), "bad args");
    if (n_isects == 0) return 0;
    B2S_REQUIRE(tile_keys_sorted && flatten_ids && depths && isect_ids, "null pointer");
    const unsigned grid = (unsigned)cdiv(n_isects, 256);
    if (key_bytes == 2)
        isect_ids_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
            n_isects, reinterpret_cast<const uint16_t*>(tile_keys_sorted), flatten_ids, depths, isect_ids);
    else
        isect_ids_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
            n_isects, reinterpret_cast<const uint32_t*>(tile_keys_sorted), flatten_ids, depths, isect_ids);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// two-level binning entry points
// ---------------------------------------------------------------------------------------------------
extern "C" int b2s_coarse_count(int64_t M, const float* means2d, const int32_t* radii, int32_t tile_size, int32_t tile_w,
                                int32_t tile_h, int64_t* total, const b2s_tile_rows* rows, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0 && total, "bad args");
    int row_lo, row_hi;
    B2S_REQUIRE(row_window(rows, tile_h, row_lo, row_hi), "bad tile-row window");
    if (M == 0) return 0;
    B2S_REQUIRE(means2d && radii, "null pointer");
    coarse_count_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(
        M, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_w, row_lo, row_hi,
        reinterpret_cast<unsigned long long*>(total));
    B2S_LAUNCH_CHECK();
    return 0;
}

static size_t pad4(int64_t n) { return (size_t)((n + 3) & ~(int64_t)3); }

extern "C" int64_t b2s_coarse_bin_workspace_bytes(int64_t M, int64_t n_records) {
    if (n_records <= 0 || M <= 0) return 256;
    // unsorted records + unsorted ids + sort scratch + look-back status (counter + one word per CTA)
    return (int64_t)(2 * align256(pad4(n_records) * 4) + align256(cub_pairs32_bytes(n_records)) +
                     align256(16 + (size_t)cdiv(M, EXP_THREADS) * 8) + 2 * align256((size_t)M * 4));
}

extern "C" int b2s_coarse_bin(int64_t M, int64_t N, const int64_t* camera_ids, int64_t n_records, const int32_t* sorted_ids,
                              const float* means2d, const int32_t* radii, int32_t C, int32_t tile_size, int32_t tile_w,
                              int32_t tile_h, uint32_t* bin_recs, int32_t* bin_ids, int32_t* bin_offsets, void* workspace,
                              int64_t workspace_bytes, const b2s_tile_rows* rows, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && n_records >= 0 && C >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0 && bin_offsets, "bad args");
    int row_lo, row_hi;
    B2S_REQUIRE(row_window(rows, tile_h, row_lo, row_hi), "bad tile-row window");
    cudaStream_t st = (cudaStream_t)stream;
    const int bins_w = (tile_w + CB - 1) / CB, bins_h = (tile_h + CB - 1) / CB;
    const int64_t n_bins = (int64_t)C * bins_w * bins_h;
    if (n_bins > 65536) {
        set_error("b2s_coarse_bin: C * bins_w * bins_h = %lld exceeds 65536", (long long)n_bins);
        return B2S_E_UNSUPPORTED;
    }
    if (n_records == 0 || M == 0) {
        B2S_CUDA(cudaMemsetAsync(bin_offsets, 0, (size_t)(n_bins + 1) * 4, st));
        return 0;
    }
    B2S_REQUIRE(camera_ids != nullptr || N > 0, "dense rows need N");
    B2S_REQUIRE(sorted_ids && means2d && radii && bin_recs && bin_ids && workspace, "null pointer");
    B2S_REQUIRE(n_records < 2147483647LL, "n_records must fit int32");
    if (workspace_bytes < b2s_coarse_bin_workspace_bytes(M, n_records)) {
        set_error("b2s_coarse_bin: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    B2S_REQUIRE(((reinterpret_cast<uintptr_t>(bin_recs) | reinterpret_cast<uintptr_t>(bin_ids)) & 15u) == 0,
                "bin_recs / bin_ids must be 16-byte aligned");
    char* w = reinterpret_cast<char*>(workspace);
    uint32_t* recs_u = reinterpret_cast<uint32_t*>(w);
    w += align256(pad4(n_records) * 4);
    int32_t* vals_u = reinterpret_cast<int32_t*>(w);
    w += align256(pad4(n_records) * 4);
    void* sort_tmp = w;
    const size_t sort_bytes = cub_pairs32_bytes(n_records);
    w += align256(sort_bytes);
    const int64_t ctas = cdiv(M, EXP_THREADS);
    uint32_t* cta_tot = reinterpret_cast<uint32_t*>(w);                 // (the look-back status area of the fused variant)
    w += align256(16 + (size_t)ctas * 8);
    uint32_t* xs = reinterpret_cast<uint32_t*>(w);
    w += align256((size_t)M * 4);
    uint32_t* ys = reinterpret_cast<uint32_t*>(w);
    if (tile_w < 65536 && tile_h < 65536 && n_records < 4294967295LL) {
        coarse_rects_kernel<<<(unsigned)ctas, EXP_THREADS, 0, st>>>(M, sorted_ids, reinterpret_cast<const float2*>(means2d), radii,
                                                                    (float)tile_size, tile_w, row_lo, row_hi, xs, ys, cta_tot);
        B2S_LAUNCH_CHECK();
        coarse_scan_kernel<<<1, 1024, 0, st>>>(cta_tot, ctas);
        B2S_LAUNCH_CHECK();
        coarse_emit_kernel<<<(unsigned)ctas, EXP_THREADS, 0, st>>>(M, N, camera_ids, sorted_ids, xs, ys, cta_tot, bins_w, bins_h,
                                                                   recs_u, vals_u);
        B2S_LAUNCH_CHECK();
    } else {
        B2S_CUDA(cudaMemsetAsync(cta_tot, 0, (size_t)(16 + ctas * 8), st));
        unsigned int* counter = reinterpret_cast<unsigned int*>(cta_tot);
        unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(cta_tot) + 16);
        coarse_expand_kernel<<<(unsigned)ctas, EXP_THREADS, 0, st>>>(
            M, N, camera_ids, sorted_ids, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_w, row_lo, row_hi,
            bins_w, bins_h, status, counter, recs_u, vals_u);
        B2S_LAUNCH_CHECK();
    }
    const int key_bits = bit_length(n_bins > 1 ? n_bins - 1 : 1);
    {
        int rc = sort_pairs<uint32_t>(sort_tmp, sort_bytes, recs_u, bin_recs, vals_u, bin_ids, n_records, key_bits, st);
        if (rc != 0) return rc;
    }
    // the kernels read the lists 4 entries at a time: define the (at most 3) pad entries
    if (pad4(n_records) > (size_t)n_records) {
        B2S_CUDA(cudaMemsetAsync(bin_recs + n_records, 0, (pad4(n_records) - (size_t)n_records) * 4, st));
        B2S_CUDA(cudaMemsetAsync(bin_ids + n_records, 0, (pad4(n_records) - (size_t)n_records) * 4, st));
    }
    coarse_offsets_kernel<<<(unsigned)cdiv(n_bins + 1, 128), 128, 0, st>>>(n_records, bin_recs, n_bins, bin_offsets);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// fused two-level binning: depth sort whose last pass emits the rectangles, one-kernel expansion
// ---------------------------------------------------------------------------------------------------
extern "C" int64_t b2s_depth_order_rects_workspace_bytes(int64_t M) {
    if (M <= 0) return 256;
    return (int64_t)align256(radix::workspace_bytes<uint32_t>(M));
}

extern "C" int b2s_depth_order_rects(int64_t M, const uint32_t* depth_keys, const int32_t* row_ids, const float* means2d,
                                     const int32_t* radii, int32_t tile_size, int32_t tile_w, int32_t tile_h,
                                     const b2s_tile_rows* rows, int32_t* sorted_ids, uint32_t* rects, void* workspace,
                                     int64_t workspace_bytes, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0, "bad args");
    if (tile_w > 255 || tile_h > 255) {
        set_error("b2s_depth_order_rects: the packed rectangles need a tile grid of at most 255 x 255 (use b2s_depth_order + b2s_coarse_bin)");
        return B2S_E_UNSUPPORTED;
    }
    int row_lo, row_hi;
    B2S_REQUIRE(row_window(rows, tile_h, row_lo, row_hi), "bad tile-row window");
    if (M == 0) return 0;
    B2S_REQUIRE(depth_keys && row_ids && means2d && radii && sorted_ids && rects && workspace, "null pointer");
    if (workspace_bytes < b2s_depth_order_rects_workspace_bytes(M)) {
        set_error("b2s_depth_order_rects: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    const bool pow2 = (tile_size & (tile_size - 1)) == 0;
    const RectSide first{reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, pow2 ? 1.0f / (float)tile_size : 0.f,
                         tile_w, row_lo, row_hi, rects};
    const IdStore last{sorted_ids};
    return radix::sort_ex<uint32_t, RectSide, IdStore>(M, 32, depth_keys, row_ids, first, last, false, workspace,
                                                       (cudaStream_t)stream);
}

extern "C" int64_t b2s_coarse_bin_rects_workspace_bytes(int64_t M, int64_t n_records) {
    if (n_records <= 0 || M <= 0) return 256;
    // unsorted records + unsorted ids + sort scratch + look-back status (counter + one word per CTA)
    return (int64_t)(2 * align256(pad4(n_records) * 4) + align256(radix::workspace_bytes<uint32_t>(n_records)) +
                     align256(16 + (size_t)cdiv(M, EMIT_ROWS) * 8));
}

extern "C" int b2s_coarse_bin_rects(int64_t M, int64_t N, const int64_t* camera_ids, int64_t n_records, const int32_t* sorted_ids,
                                    const uint32_t* rects, int32_t C, int32_t tile_w, int32_t tile_h,
                                    uint32_t* bin_recs, int32_t* bin_ids, int32_t* bin_offsets, void* workspace,
                                    int64_t workspace_bytes, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && n_records >= 0 && C >= 0 && tile_w > 0 && tile_h > 0 && bin_offsets, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    const int bins_w = (tile_w + CB - 1) / CB, bins_h = (tile_h + CB - 1) / CB;
    const int64_t n_bins = (int64_t)C * bins_w * bins_h;
    if (n_bins > 65536) {
        set_error("b2s_coarse_bin_rects: C * bins_w * bins_h = %lld exceeds 65536", (long long)n_bins);
        return B2S_E_UNSUPPORTED;
    }
    if (n_records == 0 || M == 0) {
        B2S_CUDA(cudaMemsetAsync(bin_offsets, 0, (size_t)(n_bins + 1) * 4, st));
        return 0;
    }
    B2S_REQUIRE(camera_ids != nullptr || N > 0, "dense rows need N");
    B2S_REQUIRE(sorted_ids && rects && bin_recs && bin_ids && workspace, "null pointer");
    B2S_REQUIRE(n_records < 2147483647LL, "n_records must fit int32");
    if (workspace_bytes < b2s_coarse_bin_rects_workspace_bytes(M, n_records)) {
        set_error("b2s_coarse_bin_rects: workspace too small");
        return B2S_E_WORKSPACE;
    }
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
    B2S_REQUIRE(((reinterpret_cast<uintptr_t>(bin_recs) | reinterpret_cast<uintptr_t>(bin_ids)) & 15u) == 0,
                "bin_recs / bin_ids must be 16-byte aligned");
    char* w = reinterpret_cast<char*>(workspace);
    uint32_t* recs_u = reinterpret_cast<uint32_t*>(w);
    w += align256(pad4(n_records) * 4);
    int32_t* vals_u = reinterpret_cast<int32_t*>(w);
    w += align256(pad4(n_records) * 4);
    void* sort_ws = w;
    w += align256(radix::workspace_bytes<uint32_t>(n_records));
    const int64_t ctas = cdiv(M, EMIT_ROWS);
    unsigned int* counter = reinterpret_cast<unsigned int*>(w);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(w + 16);
    const int key_bits = bit_length(n_bins > 1 ? n_bins - 1 : 1);
    const radix::Plan pl = radix::make_plan(key_bits);                       // at most 2 passes (16 key bits)
    unsigned long long* hist = radix::hist_ptr<uint32_t>(sort_ws, n_records);
    B2S_CUDA(cudaMemsetAsync(counter, 0, (size_t)(16 + ctas * 8), st));
    B2S_CUDA(cudaMemsetAsync(hist, 0, radix::hist_bytes<uint32_t>(), st));
    coarse_emit_lb_kernel<<<(unsigned)ctas, EXP_THREADS, 0, st>>>(M, N, camera_ids, sorted_ids, rects, bins_w, bins_h, status, counter, pl,
                                                                  hist, recs_u, vals_u);
    B2S_LAUNCH_CHECK();
    {
        const radix::StorePairs<uint32_t> out{bin_recs, bin_ids};
        int rc = radix::sort_pairs_ex<uint32_t, radix::StorePairs<uint32_t>>(n_records, key_bits, recs_u, vals_u, out, true, sort_ws, st);
        if (rc != 0) return rc;
    }
    if (pad4(n_records) > (size_t)n_records) {       // the kernels read the lists 4 entries at a time: define the pad entries
        B2S_CUDA(cudaMemsetAsync(bin_recs + n_records, 0, (pad4(n_records) - (size_t)n_records) * 4, st));
        B2S_CUDA(cudaMemsetAsync(bin_ids + n_records, 0, (pad4(n_records) - (size_t)n_records) * 4, st));
    }
    if (pl.passes == 1) coarse_offsets_from_hist_kernel<<<1, 256, 0, st>>>(hist, n_bins, n_records, bin_offsets);
    else coarse_offsets_kernel<<<(unsigned)cdiv(n_bins + 1, 128), 128, 0, st>>>(n_records, bin_recs, n_bins, bin_offsets);
    B2S_LAUNCH_CHECK();
    return 0;
}
