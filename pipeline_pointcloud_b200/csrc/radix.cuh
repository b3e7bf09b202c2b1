// Hand-written stable LSD radix sort of (key, int32 value) pairs for sm_100a -- "onesweep" style:
// one histogram kernel for all digits, then ONE kernel per 8-bit digit that ranks, looks back and
// scatters in a single pass over the data (read 1x, write 1x per digit).
//
// Per digit pass, a CTA owns a tile of 256 threads x ITEMS consecutive elements:
//   1. warp-striped coalesced load; each warp owns a contiguous slice, so (warp, item, lane) order is
//      memory order and ranking in that order keeps the sort stable;
//   2. per-warp ranking: lanes holding the same digit find each other through a shared-memory mask word they all
//      `atomicOr` their lane bit into (B2S_RADIX_RANK: `match.any` and per-bit ballots are the measured alternatives),
//      the group leader bumps the warp's digit counter, rank = counter + popc(lower peers);
//   3. per-digit exclusive scan across the 8 warps (thread d <-> digit d); the tile's aggregate is published at once;
//   4. local reorder through shared memory (it needs nothing from other tiles and frees the item registers), then the
//      decoupled look-back across tiles, one thread per digit (256 chains in parallel, coalesced status rows), tile ids
//      handed out dynamically so predecessors are always resident;
//   5. coalesced scatter: runs of equal digits leave contiguously.  The last pass may hand each element to an epilogue
//      instead (the depth sort stores only the ids); the histogram kernel may carry a hook that sees every input element
//      (the depth sort derives the rows' tile rectangles there).
// Replaces cub::DeviceRadixSort on the training path (depth sort of Gaussians, (camera,tile) split of the
// intersections) and in the reference-algorithm path (64-bit isect keys).
#pragma once
#include "common.cuh"

namespace b2s {
namespace radix {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
// The CTAs flush their shared-memory digit counts into HIST_COPIES copies of the global histogram (copy = CTA % copies), which
// the scan kernel adds up: same-address atomics serialise in L2 (~27 cycles each), and with a few hot digits (a depth
// key's top byte, 77 coarse bins) one copy made the flush the whole cost of the kernel -- 18 us for 1 184 CTAs, 80 us for the
// 5 665 CTAs of the coarse-bin expansion.
constexpr int HIST_COPIES = 16;
constexpr unsigned long long R_FLAG_AGG = 1ull << 62, R_FLAG_INC = 2ull << 62, R_VAL_MASK = (1ull << 62) - 1;

#ifndef B2S_RADIX_ITEMS32
#define B2S_RADIX_ITEMS32 16
#endif
template <typename KeyT> struct Cfg { static constexpr int ITEMS = B2S_RADIX_ITEMS32; };
template <> struct Cfg<uint64_t> { static constexpr int ITEMS = 8; };

template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift, unsigned mask = RADIX - 1) {
    return (unsigned)(k >> shift) & mask;
}

// Digit widths: the key bits are split evenly over the passes (14 bits -> 7 + 7 rather than 8 + 6), which
// doubles the length of the runs the 8-bit pass would scatter and keeps both passes equally cheap.
struct Plan {
    int passes;
    int shift[8];
    int bits[8];
};
static inline Plan make_plan(int key_bits) {
    Plan pl;
    pl.passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    int done = 0;
    for (int p = 0; p < pl.passes; ++p) {
        int b = (key_bits - done + (pl.passes - p) - 1) / (pl.passes - p);
        pl.shift[p] = done;
        pl.bits[p] = b;
        done += b;
    }
    return pl;
}

// ---- histogram of every digit position in one pass over the keys ------------------------------------------
// A hook that sees every (index, value) of the input once, in input order, from the histogram kernel -- a streaming pass at
// full occupancy whose time goes to shared-memory atomics, so independent coalesced loads and stores ride along for free.
// The depth sort uses it to derive each row's tile rectangle on the way (binning.cu).  (Tried, round 2: the same hook inside
// the first digit pass -- +71 us there, the pass is latency bound at 35 % occupancy; and carrying the rectangle through the
// passes as a second 32-bit payload -- +19 us per pass, register spills included: profiles/r02_variants_c15.txt.)
struct NoSide {
    static constexpr bool kActive = false;
    __device__ __forceinline__ void operator()(int64_t /*i*/, int32_t /*val*/) const {}
};

template <typename KeyT, int MAX_PASSES, typename Side>
__global__ void __launch_bounds__(THREADS) hist_kernel(int64_t n, const KeyT* __restrict__ keys, const int32_t* __restrict__ vals,
                                                       const Side side, const Plan pl,
                                                       unsigned long long* __restrict__ hist /* [copies][MAX_PASSES][256] */) {
    const int passes = pl.passes;
    hist += (size_t)(blockIdx.x % HIST_COPIES) * MAX_PASSES * RADIX;
    __shared__ unsigned s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += THREADS) s_hist[i] = 0;
    __syncthreads();
    // U consecutive elements per thread and iteration: their loads (and the dependent loads of the hook) are in flight
    // together and a warp still touches one contiguous run.  (U elements a grid stride apart measured SLOWER than one at a
    // time: 62 vs 45 us with the rectangle hook at 5.8 M rows, profiles/r02_variants_c15.txt.)
#ifndef B2S_HIST_U
#define B2S_HIST_U 4
#endif
    constexpr int U = B2S_HIST_U;
    const int64_t stride = (int64_t)gridDim.x * THREADS * U;
    for (int64_t i0 = ((int64_t)blockIdx.x * THREADS + threadIdx.x) * U; i0 < n; i0 += stride) {
        KeyT k[U];
        int32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u;
            k[u] = i < n ? keys[i] : (KeyT)0;
            v[u] = (Side::kActive && i < n) ? vals[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (Side::kActive && i0 + u < n) side(i0 + u, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u >= n) break;
#pragma unroll
            for (int p = 0; p < MAX_PASSES; ++p)
                if (p < passes) atomicAdd(&s_hist[p * RADIX + digit_of(k[u], pl.shift[p], (1u << pl.bits[p]) - 1u)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += THREADS)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

// sum of the copies, then exclusive scan of each 256-bin histogram (one CTA per pass); the result replaces copy 0
static __global__ void __launch_bounds__(RADIX) scan_hist_kernel(unsigned long long* __restrict__ hist, int max_passes) {
    __shared__ unsigned long long s[RADIX];
    unsigned long long* h = hist + (size_t)blockIdx.x * RADIX;
    unsigned long long v = 0ull;
#pragma unroll 4
    for (int c = 0; c < HIST_COPIES; ++c) v += h[(size_t)c * max_passes * RADIX + threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < RADIX; d <<= 1) {
        unsigned long long t = threadIdx.x >= d ? s[threadIdx.x - d] : 0ull;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

// ---- one digit pass ---------------------------------------------------------------------------------------
// The kernel is issue-bound, so everything is kept in 32 bits: indices (n < 2^31 enforced by the host driver) and
// the look-back status words (StatusT = uint32 with a 30-bit value when n < 2^30, else uint64).
template <typename StatusT> struct St;
template <> struct St<unsigned> {
    static constexpr int SHIFT = 30;
    static constexpr unsigned MASK = (1u << 30) - 1u;
};
template <> struct St<unsigned long long> {
    static constexpr int SHIFT = 62;
    static constexpr unsigned long long MASK = (1ull << 62) - 1ull;
};

#ifndef B2S_RADIX_MINBLOCKS
#define B2S_RADIX_MINBLOCKS 3
#endif
// What the scatter of a pass does with an element whose final position is known.  The default stores the pair; the last
// pass of the depth sort swaps in an epilogue that stores only what the next stage reads (binning.cu).
template <typename KeyT>
struct StorePairs {
    KeyT* keys_out;
    int32_t* vals_out;
    __device__ __forceinline__ void operator()(bool active, unsigned pos, KeyT key, int32_t val) const {
        if (active) {
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
    }
};
// Ranking inside a warp (A/B on config 3, profiles/r02_variants_c15.txt: depth sort 0.393 / 0.346 / 0.399 ms, bin pass alike):
//   0  `match.any` on the digit: its cost grows with the number of distinct digits in the warp (~28 of 32 on the uniform
//      low digits; the concentrated top digit ran twice as fast, 41 vs 72-76 us)
//   1  peer masks built with shared-memory atomicOr (two alternating mask arrays that alias the reorder buffers) -- the
//      product default
//   2  one ballot per digit bit
#ifndef B2S_RADIX_RANK
#define B2S_RADIX_RANK 1
#endif

template <typename KeyT, typename StatusT, typename Epi>
__global__ void __launch_bounds__(THREADS, B2S_RADIX_MINBLOCKS) onesweep_kernel(unsigned n, const KeyT* __restrict__ keys_in,
                                                              const int32_t* __restrict__ vals_in, const Epi epi,
                                                              int shift, unsigned mask,
                                                              const unsigned long long* __restrict__ digit_base /* [256] */,
                                                              StatusT* status /* [tiles][256] */, unsigned int* tile_counter) {
    constexpr int ITEMS = Cfg<KeyT>::ITEMS;
    constexpr unsigned TILE = THREADS * ITEMS;
    constexpr int HS = RADIX + 1;                        // +1: dump bin for the out-of-range lanes of the last tile
    __shared__ unsigned s_warp_hist[WARPS * HS];         // per-warp digit counters -> exclusive prefix across warps
    __shared__ unsigned s_cta_start[RADIX];              // first slot of each digit inside the locally sorted tile
    __shared__ unsigned s_off[RADIX];                    // (global position of the digit's first element) - cta_start
    __shared__ KeyT s_keys[TILE];
    __shared__ int32_t s_vals[TILE];
    __shared__ unsigned s_wsum[WARPS];
    __shared__ unsigned int s_tile;
#if B2S_RADIX_RANK == 1
    static_assert(4 * TILE >= (size_t)WARPS * RADIX * 4, "mask arrays alias the reorder buffers");
    unsigned* s_mask0 = reinterpret_cast<unsigned*>(s_keys);     // [WARPS][RADIX], dead before the reorder starts
    unsigned* s_mask1 = reinterpret_cast<unsigned*>(s_vals);
#endif

    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < WARPS * HS; i += THREADS) s_warp_hist[i] = 0;
#if B2S_RADIX_RANK == 1
    for (int i = threadIdx.x; i < WARPS * RADIX; i += THREADS) {
        s_mask0[i] = 0;
        s_mask1[i] = 0;
    }
#endif
    __syncthreads();
    const unsigned tile = s_tile;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tile_base = tile * TILE;
    const unsigned warp_base = tile_base + warp * 32 * ITEMS + lane;
    const bool full = tile_base + TILE <= n;
    const unsigned tile_count = full ? TILE : n - tile_base;

    KeyT key[ITEMS];
    int32_t val[ITEMS];
    unsigned dig[ITEMS];
    if (full) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = keys_in[warp_base + i * 32];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) val[i] = vals_in[warp_base + i * 32];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) dig[i] = digit_of(key[i], shift, mask);
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const unsigned g = warp_base + i * 32;
            const bool ok = g < n;
            key[i] = ok ? keys_in[g] : (KeyT)0;
            val[i] = ok ? vals_in[g] : 0;
            dig[i] = ok ? digit_of(key[i], shift, mask) : (unsigned)RADIX;   // out-of-range lanes share the dump bin
        }
    }
    // 2. rank inside the warp, in memory order.  All MATCH instructions are issued first (independent); then, item
    // by item, every lane reads its digit's counter (a broadcast within the group) and the group leader bumps it.
    unsigned* my_hist = s_warp_hist + warp * HS;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned rank[ITEMS];
#if B2S_RADIX_RANK == 1
    // peers through shared memory: every lane ORs its bit into its digit's mask word; the group leader clears the word
    // again.  Consecutive items use different mask arrays, so the clear of item i never races the ORs of item i + 1.
    {
        unsigned* mk0 = s_mask0 + warp * RADIX;
        unsigned* mk1 = s_mask1 + warp * RADIX;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            unsigned* mk = (i & 1) ? mk1 : mk0;
            const bool live = full || dig[i] < RADIX;            // the dump-bin lanes of the last tile are never stored
            if (live) atomicOr(&mk[dig[i]], 1u << lane);
            __syncwarp();
            const unsigned peers = live ? mk[dig[i]] : (1u << lane);
            const unsigned pre = my_hist[dig[i]];
            __syncwarp();
            if (live && lane == __ffs(peers) - 1) {
                mk[dig[i]] = 0u;
                my_hist[dig[i]] = pre + __popc(peers);
            }
            rank[i] = pre + __popc(peers & lt_mask);
        }
        __syncwarp();
    }
#else
    unsigned peers[ITEMS];
#if B2S_RADIX_RANK == 2
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        unsigned m = FULL;
#pragma unroll
        for (int b = 0; b <= RADIX_BITS; ++b) {          // bit 8: the dump bin of the last tile
            if (b == RADIX_BITS && full) break;
            const bool bit = (dig[i] >> b) & 1u;
            const unsigned bal = __ballot_sync(FULL, bit);
            m &= bit ? bal : ~bal;
        }
        peers[i] = m;
    }
#else
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) peers[i] = __match_any_sync(FULL, dig[i]);
#endif
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const unsigned pre = my_hist[dig[i]];
        __syncwarp();
        if (lane == __ffs(peers[i]) - 1) my_hist[dig[i]] = pre + __popc(peers[i]);
        rank[i] = pre + __popc(peers[i] & lt_mask);
        __syncwarp();
    }
#endif
    __syncthreads();
    // 3. thread d <-> digit d: exclusive prefix across warps, tile total
    const unsigned d = threadIdx.x;
    unsigned cnt = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const unsigned t = s_warp_hist[w * HS + d];
        s_warp_hist[w * HS + d] = cnt;
        cnt += t;
    }
    // 4a. publish the tile's aggregate for digit d at once (the status word carries flag and value together: no fence
    // needed) -- the sooner it is visible, the less the tiles behind this one wait
    constexpr int SH = St<StatusT>::SHIFT;
    volatile StatusT* st = status;
    st[tile * RADIX + d] = ((StatusT)(tile == 0 ? 2 : 1) << SH) | (StatusT)cnt;
    // exclusive scan of the tile's digit counts (256 values, one per thread)
    unsigned start;
    {
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        unsigned woff = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w)
            if (w < (int)warp) woff += s_wsum[w];
        start = woff + incl - cnt;
        s_cta_start[d] = start;
    }
    __syncthreads();
    // 5a. local reorder: needs nothing from other tiles, and it ends the life of the 4 x ITEMS item registers
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (full || dig[i] < RADIX) {
            const unsigned p = s_cta_start[dig[i]] + s_warp_hist[warp * HS + dig[i]] + rank[i];
            s_keys[p] = key[i];
            s_vals[p] = val[i];
        }
    }
    // 4b. look-back for digit d, LB tiles per round trip (independent loads).  (Windows of 16 and 32 tiles measured no
    // better / worse than 8: 0.272 / 0.295 vs 0.272 ms for the depth sort -- the walk is short in practice and a wider poll
    // only adds traffic; publishing early and reordering first is what gained, 0.285 -> 0.272 ms.)
    {
        unsigned long long excl = 0;
        if (tile != 0) {
#ifndef B2S_RADIX_LB
#define B2S_RADIX_LB 8
#endif
            constexpr int LB = B2S_RADIX_LB;
            int k = (int)tile - 1;
            while (true) {
                StatusT sv[LB];
#pragma unroll
                for (int u = 0; u < LB; ++u) sv[u] = (k - u >= 0) ? st[(unsigned)(k - u) * RADIX + d] : ((StatusT)2 << SH);
                unsigned ready = 0, inc = 0;
#pragma unroll
                for (int u = 0; u < LB; ++u) {
                    const unsigned f = (unsigned)(sv[u] >> SH);
                    ready |= (unsigned)(f != 0u) << u;
                    inc |= (unsigned)(f == 2u) << u;
                }
                const int first_wait = (LB == 32 && ready == FULL) ? 32 : __ffs(~ready) - 1;   // LB when all are published
                const int first_inc = inc ? __ffs(inc) - 1 : LB;
                const int usable = min(first_wait, first_inc + 1);              // tiles [k, k-usable) can be consumed
#pragma unroll
                for (int u = 0; u < LB; ++u)
                    if (u < usable) excl += (unsigned long long)(sv[u] & St<StatusT>::MASK);
                if (first_inc < usable) break;
                k -= usable;
            }
            st[tile * RADIX + d] = ((StatusT)2 << SH) | (StatusT)(excl + cnt);
        }
        s_off[d] = (unsigned)(digit_base[d] + excl) - start;
    }
    __syncthreads();
    // 5b. coalesced scatter
#pragma unroll 4
    for (unsigned p = threadIdx.x; p - threadIdx.x < tile_count; p += THREADS) {     // whole warps: the epilogue may vote
        const KeyT k = s_keys[p];
        const unsigned pos = s_off[digit_of(k, shift, mask)] + p;
        epi(p < tile_count, pos, k, s_vals[p]);
    }
}

// ---- host driver --------------------------------------------------------------------------------------------
template <typename KeyT>
static inline int64_t tiles_of(int64_t n) { return cdiv(n > 0 ? n : 1, (int64_t)THREADS * Cfg<KeyT>::ITEMS); }

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

template <typename KeyT>
static inline int passes_of(int key_bits) { return (key_bits + RADIX_BITS - 1) / RADIX_BITS; }

// workspace: 2 temp key buffers + 2 temp value buffers + histograms + status + tile counters
template <typename KeyT>
static inline size_t workspace_bytes(int64_t n) {
    const int max_passes = (int)sizeof(KeyT);
    return 2 * up256((size_t)n * sizeof(KeyT)) + 2 * up256((size_t)n * 4) + up256((size_t)HIST_COPIES * max_passes * RADIX * 8) +
           up256((size_t)tiles_of<KeyT>(n) * RADIX * 8) + up256((size_t)max_passes * 4);
}

// where the [HIST_COPIES][sizeof(KeyT)][256] digit histograms live inside the workspace (a caller that counts the digits
// itself -- the coarse-bin expansion -- zeroes all copies (hist_bytes), adds into any of them, then sorts with
// hist_ready = true)
template <typename KeyT>
static inline size_t hist_bytes() { return (size_t)HIST_COPIES * sizeof(KeyT) * RADIX * 8; }
template <typename KeyT>
static inline unsigned long long* hist_ptr(void* workspace, int64_t n) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 2 * up256((size_t)n * sizeof(KeyT)) +
                                                 2 * up256((size_t)n * 4));
}
template <typename KeyT, typename StatusT, typename Epi>
static inline cudaError_t launch_pass(int64_t tiles, unsigned n, const KeyT* k, const int32_t* v, const Epi& epi, int shift,
                                      unsigned mask, const unsigned long long* base, void* status, unsigned int* counter,
                                      cudaStream_t st) {
    onesweep_kernel<KeyT, StatusT, Epi><<<(unsigned)tiles, THREADS, 0, st>>>(n, k, v, epi, shift, mask, base,
                                                                            reinterpret_cast<StatusT*>(status), counter);
    return cudaGetLastError();
}

// Sorts bits [0, key_bits) ascending, stable.  keys_in/vals_in are not modified.  Every pass but the last stores its
// output into the workspace; the last pass hands each element with its final position to `last` (StorePairs = the plain
// sort); `side` sees every input element once, from the histogram kernel (NoSide = nothing).  hist_ready: the workspace
// histograms already hold the digit counts of `make_plan(key_bits)` (then `side` is not called).
template <typename KeyT, typename Side, typename Epi>
static int sort_ex(int64_t n, int key_bits, const KeyT* keys_in, const int32_t* vals_in, const Side& side, const Epi& last,
                   bool hist_ready, void* workspace, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n >= 2147483647LL) {
        set_error("radix::sort_pairs: n must be < 2^31");
        return B2S_E_BADARG;
    }
    constexpr int MAXP = (int)sizeof(KeyT);
    const Plan pl = make_plan(key_bits);
    const int passes = pl.passes;
    if (passes < 1 || passes > MAXP) {
        set_error("radix::sort_pairs: key_bits out of range");
        return B2S_E_BADARG;
    }
    char* w = reinterpret_cast<char*>(workspace);
    KeyT* tk[2];
    int32_t* tv[2];
    tk[0] = reinterpret_cast<KeyT*>(w); w += up256((size_t)n * sizeof(KeyT));
    tk[1] = reinterpret_cast<KeyT*>(w); w += up256((size_t)n * sizeof(KeyT));
    tv[0] = reinterpret_cast<int32_t*>(w); w += up256((size_t)n * 4);
    tv[1] = reinterpret_cast<int32_t*>(w); w += up256((size_t)n * 4);
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(w); w += up256(hist_bytes<KeyT>());
    unsigned long long* status = reinterpret_cast<unsigned long long*>(w);
    const int64_t tiles = tiles_of<KeyT>(n);
    w += up256((size_t)tiles * RADIX * 8);
    unsigned int* counters = reinterpret_cast<unsigned int*>(w);

    if (!hist_ready) {
        B2S_CUDA(cudaMemsetAsync(hist, 0, hist_bytes<KeyT>(), st));
        const int hist_grid = (int)min((int64_t)148 * 8, cdiv(n, THREADS));
        hist_kernel<KeyT, MAXP, Side><<<hist_grid, THREADS, 0, st>>>(n, keys_in, vals_in, side, pl, hist);
        B2S_LAUNCH_CHECK();
    }
    B2S_CUDA(cudaMemsetAsync(counters, 0, (size_t)MAXP * 4, st));
    scan_hist_kernel<<<passes, RADIX, 0, st>>>(hist, MAXP);
    B2S_LAUNCH_CHECK();
    using Mid = StorePairs<KeyT>;
    const KeyT* src_k = keys_in;
    const int32_t* src_v = vals_in;
    const bool small = n < (1LL << 30);
    for (int p = 0; p < passes; ++p) {
        B2S_CUDA(cudaMemsetAsync(status, 0, (size_t)tiles * RADIX * (small ? 4 : 8), st));
        const unsigned mask = (1u << pl.bits[p]) - 1u;
        const unsigned long long* base = hist + (size_t)p * RADIX;
        const Mid mid{tk[p & 1], tv[p & 1]};
        cudaError_t e;
#define B2S_PASS(EPI, epi_)                                                                                                      \
    (small ? launch_pass<KeyT, unsigned, EPI>(tiles, (unsigned)n, src_k, src_v, epi_, pl.shift[p], mask, base, status, counters + p, st) \
           : launch_pass<KeyT, unsigned long long, EPI>(tiles, (unsigned)n, src_k, src_v, epi_, pl.shift[p], mask, base, status,   \
                                                        counters + p, st))
        if (p == passes - 1) e = B2S_PASS(Epi, last);
        else e = B2S_PASS(Mid, mid);
#undef B2S_PASS
        if (e != cudaSuccess) {
            set_error("radix::sort: digit pass %d -> %s", p, cudaGetErrorString(e));
            return (int)e;
        }
        src_k = tk[p & 1];
        src_v = tv[p & 1];
    }
    return 0;
}

template <typename KeyT, typename Epi>
static int sort_pairs_ex(int64_t n, int key_bits, const KeyT* keys_in, const int32_t* vals_in, const Epi& last, bool hist_ready,
                         void* workspace, cudaStream_t st) {
    return sort_ex<KeyT, NoSide, Epi>(n, key_bits, keys_in, vals_in, NoSide{}, last, hist_ready, workspace, st);
}

template <typename KeyT>
static int sort_pairs(int64_t n, int key_bits, const KeyT* keys_in, const int32_t* vals_in, KeyT* keys_out, int32_t* vals_out,
                      void* workspace, cudaStream_t st) {
    return sort_pairs_ex<KeyT, StorePairs<KeyT>>(n, key_bits, keys_in, vals_in, StorePairs<KeyT>{keys_out, vals_out}, false,
                                                 workspace, st);
}

}  // namespace radix
}  // namespace b2s
