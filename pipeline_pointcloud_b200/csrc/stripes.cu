// Tile-stripe routing for the multi-GPU path (BASELINE config 5; no counterpart in the reference,
// whose multi-GPU branch -- gsplat `rasterization(distributed=True)` behind source/container/src/main.py:1317-1347 --
// routes to camera owners; pack.cu holds that scheme).
//
// Every rank owns a shard of the Gaussians and a stripe of tile rows of the shared camera(s).  After projecting
// its shard a rank must hand every visible splat to each stripe its tile rectangle touches, and get the gradients
// back.  The exchange is folded into the producing kernels over NVLink peer mappings (symmetric memory):
//
//   stripe_count   one pass over the dense rows: destination mask per row, per-CTA row counts per destination,
//                  and per-destination totals (rows, tile intersections, coarse-bin records) so that the receiver
//                  can size every buffer of its binning from the W x W table -- no second host round trip;
//   stripe_scan    exclusive scan of the per-CTA counts, one CTA per destination;
//   p2p_put        tiny peer stores (the totals row into every rank's table, counters into their owners' tables);
//   stripe_pack    one pass over the dense rows again: every row computes its slot in each destination's list
//                  (table base + CTA prefix + rank inside the CTA, ascending row order => deterministic, and per
//                  camera ascending global Gaussian id -- the reference's tie-break order) and stores its record,
//                  field by field (SoA), straight into the destination rank's receive buffer;
//   stripe_unpack  receiver: copies the fields out of the (reused) exchange buffer into the caller's tensors and
//                  derives what binning needs (tile count inside the stripe, depth key, row id) in the same pass;
//   stripe_return  backward: only rows compositing actually reached carry a gradient (a few percent in a
//                  saturated scene), so the owner of a stripe sends back sparse (source row, gradient) records,
//                  appended per source rank with a warp-aggregated counter, straight into the source's buffer;
//   stripe_scatter_add  source: adds the returned records into the dense (zero-filled) gradients of its rows; a
//                  splat that straddled a stripe boundary receives one record per stripe.
#include "b2s_math.cuh"
#include "common.cuh"

namespace b2s {

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int CB_ = B2S_COARSE;

struct StripeTable {
    int world;
    int cams_per_rank;                 // > 0: camera owners (rank d owns cameras [d*cpr, (d+1)*cpr), whole tile grid)
    int bounds[B2S_MAX_RANKS + 1];     // cams_per_rank == 0: tile rows of the stripes
};
struct PeerPtrs {
    void* p[B2S_MAX_RANKS];
};
struct RankI64 {
    int64_t v[B2S_MAX_RANKS + 1];
};

__device__ __forceinline__ int coarse_span_(int t0, int t1) { return (t1 + CB_ - 1) / CB_ - t0 / CB_; }

// -------------------------------------------------------------------------------------------------------------
// count
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS) stripe_count_kernel(int64_t M, const float2* __restrict__ means2d,
                                                                  const int32_t* __restrict__ radii, int64_t N, float tile_size,
                                                                  int tile_w, int tile_h, const StripeTable sb,
                                                                  uint16_t* __restrict__ dest_mask,
                                                                  uint32_t* __restrict__ cta_counts, int64_t n_ctas,
                                                                  unsigned long long* __restrict__ totals, long long tag) {
    __shared__ unsigned s_acc[B2S_MAX_RANKS * 3];
    if (threadIdx.x < B2S_MAX_RANKS * 3) s_acc[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    bool vis = false;
    if (i < M) {
        const int r = radii[i];
        if (r > 0) {
            const float2 m = means2d[i];
            tile_rect(m.x, m.y, r, tile_size, tile_w, tile_h, x0, y0, x1, y1);
            vis = x1 > x0 && y1 > y0;
        }
    }
    unsigned mask = 0;
    const int owner = sb.cams_per_rank > 0 ? (int)((i < M ? i : 0) / N) / sb.cams_per_rank : -1;
    for (int d = 0; d < sb.world; ++d) {
        const int lo = sb.cams_per_rank > 0 ? 0 : sb.bounds[d], hi = sb.cams_per_rank > 0 ? tile_h : sb.bounds[d + 1];
        const int cy0 = min(max(y0, lo), hi), cy1 = min(max(y1, lo), hi);
        const bool hit = vis && cy1 > cy0 && (owner < 0 || owner == d);
        const unsigned b = __ballot_sync(FULL, hit);
        if (b == 0u) continue;                                      // warp-uniform
        unsigned tiles = hit ? (unsigned)((x1 - x0) * (cy1 - cy0)) : 0u;
        unsigned coarse = hit ? (unsigned)(coarse_span_(x0, x1) * coarse_span_(cy0, cy1)) : 0u;
        tiles = __reduce_add_sync(FULL, tiles);
        coarse = __reduce_add_sync(FULL, coarse);
        if (hit) mask |= 1u << d;
        if (lane == 0) {
            atomicAdd(&s_acc[3 * d + 0], (unsigned)__popc(b));
            atomicAdd(&s_acc[3 * d + 1], tiles);
            atomicAdd(&s_acc[3 * d + 2], coarse);
        }
    }
    if (i < M) dest_mask[i] = (uint16_t)mask;
    __syncthreads();
    // per-CTA partial sums, reduced by stripe_scan_kernel (same-address global atomics from every CTA serialise in
    // L2: 0.26 ms for 350 K of them, independent of the problem size)
    if ((int)threadIdx.x < sb.world) {
        const int d = threadIdx.x;
        if (blockIdx.x == 0) totals[4 * d + 3] = (unsigned long long)tag;
        const int64_t plane = (int64_t)sb.world * n_ctas;
        cta_counts[(int64_t)d * n_ctas + blockIdx.x] = s_acc[3 * d];
        cta_counts[plane + (int64_t)d * n_ctas + blockIdx.x] = s_acc[3 * d + 1];
        cta_counts[2 * plane + (int64_t)d * n_ctas + blockIdx.x] = s_acc[3 * d + 2];
    }
}

// exclusive scan of one destination's per-CTA counts (one CTA per destination; every thread scans a contiguous
// chunk, the chunk totals are scanned in shared memory)
constexpr int SCAN_T = 1024;
__global__ void __launch_bounds__(SCAN_T) stripe_scan_kernel(uint32_t* __restrict__ cta_counts, int64_t n_ctas,
                                                             unsigned long long* __restrict__ totals) {
    __shared__ unsigned s_tot[SCAN_T];
    __shared__ unsigned long long s_red[2][SCAN_T / 32];
    const int64_t plane = (int64_t)gridDim.x * n_ctas;
    uint32_t* c = cta_counts + (int64_t)blockIdx.x * n_ctas;
    const int64_t per = (n_ctas + SCAN_T - 1) / SCAN_T;
    const int64_t b = min(n_ctas, (int64_t)threadIdx.x * per), e = min(n_ctas, b + per);
    unsigned sum = 0;
    for (int64_t k = b; k < e; ++k) sum += c[k];
    // totals of the two other planes (tile intersections, coarse records): coalesced strided reads
    unsigned long long t1 = 0, t2 = 0;
    for (int64_t k = threadIdx.x; k < n_ctas; k += SCAN_T) { t1 += c[plane + k]; t2 += c[2 * plane + k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t1 += __shfl_xor_sync(FULL, t1, o); t2 += __shfl_xor_sync(FULL, t2, o); }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = t1; s_red[1][threadIdx.x >> 5] = t2; }
    s_tot[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a1 = 0, a2 = 0;
        for (int w = 0; w < SCAN_T / 32; ++w) { a1 += s_red[0][w]; a2 += s_red[1][w]; }
        totals[4 * blockIdx.x + 1] = a1;
        totals[4 * blockIdx.x + 2] = a2;
    }
    for (int d = 1; d < SCAN_T; d <<= 1) {
        const unsigned t = (int)threadIdx.x >= d ? s_tot[threadIdx.x - d] : 0u;
        __syncthreads();
        s_tot[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned run = s_tot[threadIdx.x] - sum;
    if (threadIdx.x == SCAN_T - 1) totals[4 * blockIdx.x] = (unsigned long long)s_tot[SCAN_T - 1];   // rows for this destination
    for (int64_t k = b; k < e; ++k) {
        const unsigned v = c[k];
        c[k] = run;
        run += v;
    }
}

// words_per_peer 64-bit words from src + d * src_stride to peer d's pointer (src_stride 0: the same words to everybody)
__global__ void p2p_put_kernel(int world, const unsigned long long* __restrict__ src, int words_per_peer, int src_stride,
                               const PeerPtrs dst) {
    for (int d = blockIdx.x; d < world; d += gridDim.x) {
        unsigned long long* o = reinterpret_cast<unsigned long long*>(dst.p[d]);
        if (o == nullptr) continue;
        for (int k = threadIdx.x; k < words_per_peer; k += blockDim.x) o[k] = src[(int64_t)d * src_stride + k];
    }
}

// -------------------------------------------------------------------------------------------------------------
// pack + exchange.  Receive buffer of a rank: one record of F = 10 + ch 32-bit words per row,
//   [x, y, depth, conic a b c, opacity, colour[ch], radius, camera, global Gaussian id]
// (AoS: what crosses NVLink is one contiguous run per (CTA, destination), not one run per field).
// -------------------------------------------------------------------------------------------------------------
struct PackArgs {
    int64_t M, N, n_ctas, gid_offset;
    int ch, world, cams_per_rank;
    const float2* means2d; const float* depths; const float* conics; const float* opac; const float* colors;
    const int32_t* radii; const uint16_t* dest_mask; const uint32_t* cta_prefix;
    int64_t base[B2S_MAX_RANKS];     // first row this rank owns in each destination's buffer
    float* peer[B2S_MAX_RANKS];
};

// Rows of a CTA are first grouped by destination in shared memory (slot order kept), then every destination's group
// leaves as ONE contiguous run of records: a CTA's ~256/W rows for a peer become a ~1.5 KB NVLink burst written with
// fully coalesced stores, instead of the 4 B-per-lane scatter a direct store would produce (the Gaussians of a warp
// land in unrelated stripes).
constexpr int PACK_STAGE = ST_THREADS + ST_THREADS / 2;    // staged records per CTA (straddlers are staged once per stripe)
constexpr int PACK_MAXCH = 4;                              // wider colour vectors take the direct path
constexpr int PACK_MAXF = 10 + PACK_MAXCH;

__global__ void __launch_bounds__(ST_THREADS) stripe_pack_kernel(const PackArgs a) {
    __shared__ unsigned s_wtot[B2S_MAX_RANKS][ST_WARPS];
    __shared__ int s_seg[B2S_MAX_RANKS + 1];               // first staged slot of each destination's group
    __shared__ float s_rec[PACK_STAGE * PACK_MAXF];
    const int64_t i = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mask = i < a.M ? (unsigned)a.dest_mask[i] : 0u;
    const unsigned lt = (1u << lane) - 1u;
    const int F = 10 + a.ch;
    for (int d = 0; d < a.world; ++d) {
        const unsigned b = __ballot_sync(FULL, (mask >> d) & 1u);
        if (lane == 0) s_wtot[d][warp] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int d = 0; d < a.world; ++d) {
            s_seg[d] = run;
            for (int w = 0; w < ST_WARPS; ++w) run += (int)s_wtot[d][w];
        }
        s_seg[a.world] = run;
    }
    __syncthreads();
    const bool stage = s_seg[a.world] <= PACK_STAGE && a.ch <= PACK_MAXCH;   // CTA-uniform
    // the row's record, loaded once (a straddling splat is stored several times)
    float rec[PACK_MAXF];
#pragma unroll
    for (int k = 0; k < PACK_MAXF; ++k) rec[k] = 0.f;
    if (mask) {
        const float2 m = a.means2d[i];
        rec[0] = m.x; rec[1] = m.y; rec[2] = a.depths[i];
        rec[3] = a.conics[3 * i]; rec[4] = a.conics[3 * i + 1]; rec[5] = a.conics[3 * i + 2];
        rec[6] = a.opac[i];
        int cam = (int)(i / a.N);
        const int gid = (int)(i - (int64_t)cam * a.N + a.gid_offset);
        if (a.cams_per_rank > 0) cam %= a.cams_per_rank;      // camera owners: the receiver numbers its own cameras from 0
        if (a.ch <= PACK_MAXCH) {
#pragma unroll
            for (int k = 0; k < PACK_MAXCH; ++k)
                if (k < a.ch) rec[7 + k] = a.colors[i * a.ch + k];
#pragma unroll
            for (int k = 0; k <= PACK_MAXCH; ++k)             // the three integers follow the colours
                if (k == a.ch) { rec[7 + k] = __int_as_float(a.radii[i]); rec[8 + k] = __int_as_float(cam); rec[9 + k] = __int_as_float(gid); }
        } else {
            rec[7] = __int_as_float(a.radii[i]); rec[8] = __int_as_float(cam); rec[9] = __int_as_float(gid);   // (direct path order)
        }
    }
    for (int d = 0; d < a.world; ++d) {
        const bool mine = (mask >> d) & 1u;
        const unsigned b = __ballot_sync(FULL, mine);
        if (!mine) continue;
        unsigned before = 0;
        for (int w = 0; w < warp; ++w) before += s_wtot[d][w];
        const int local = (int)before + __popc(b & lt);              // rank among this CTA's rows for d
        if (stage) {
            float* r = s_rec + (s_seg[d] + local) * F;
#pragma unroll
            for (int k = 0; k < PACK_MAXF; ++k)
                if (k < F) r[k] = rec[k];
        } else {
            const int64_t s = a.base[d] + a.cta_prefix[(int64_t)d * a.n_ctas + blockIdx.x] + local;
            float* o = a.peer[d] + s * F;
            if (a.ch <= PACK_MAXCH) {
#pragma unroll
                for (int k = 0; k < PACK_MAXF; ++k)
                    if (k < F) o[k] = rec[k];
            } else {
#pragma unroll
                for (int k = 0; k < 7; ++k) o[k] = rec[k];
                for (int k = 0; k < a.ch; ++k) o[7 + k] = a.colors[i * a.ch + k];
                o[7 + a.ch] = rec[7]; o[8 + a.ch] = rec[8]; o[9 + a.ch] = rec[9];
            }
        }
    }
    if (!stage) return;
    __syncthreads();
    for (int d = 0; d < a.world; ++d) {
        const int g0 = s_seg[d], cnt = s_seg[d + 1] - g0;
        if (cnt == 0) continue;
        const int64_t s0 = a.base[d] + a.cta_prefix[(int64_t)d * a.n_ctas + blockIdx.x];
        float* o = a.peer[d] + s0 * F;
        const float* r = s_rec + g0 * F;
        for (int t = threadIdx.x; t < cnt * F; t += ST_THREADS) o[t] = r[t];
    }
}

// receiver: exchange buffer -> the caller's tensors, plus what binning needs
__global__ void __launch_bounds__(ST_THREADS) stripe_unpack_kernel(int64_t n, int ch, const float* __restrict__ buf,
                                                                   float tile_size, int tile_w, int row_lo, int row_hi,
                                                                   float2* __restrict__ means2d, float* __restrict__ depths,
                                                                   float* __restrict__ conics, float* __restrict__ opac,
                                                                   float* __restrict__ colors, int32_t* __restrict__ radii,
                                                                   int64_t* __restrict__ camera_ids, int64_t* __restrict__ gaussian_ids,
                                                                   int32_t* __restrict__ tiles, uint32_t* __restrict__ depth_keys,
                                                                   int32_t* __restrict__ row_ids) {
    const int64_t j = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x;
    if (j >= n) return;
    const float* p = buf + j * (10 + ch);
    const float2 m = make_float2(p[0], p[1]);
    const float depth = p[2];
    means2d[j] = m;
    depths[j] = depth;
    conics[3 * j] = p[3]; conics[3 * j + 1] = p[4]; conics[3 * j + 2] = p[5];
    opac[j] = p[6];
    for (int k = 0; k < ch; ++k) colors[j * ch + k] = p[7 + k];
    const int radius = __float_as_int(p[7 + ch]);
    radii[j] = radius;
    camera_ids[j] = (int64_t)__float_as_int(p[8 + ch]);
    gaussian_ids[j] = (int64_t)__float_as_int(p[9 + ch]);
    int x0, y0, x1, y1;
    tile_rect(m.x, m.y, radius, tile_size, tile_w, row_lo, row_hi, x0, y0, x1, y1);
    tiles[j] = max(0, x1 - x0) * max(0, y1 - y0);
    depth_keys[j] = radius > 0 ? __float_as_uint(depth) : 0xFFFFFFFFu;
    row_ids[j] = (int32_t)j;
}

// -------------------------------------------------------------------------------------------------------------
// backward: sparse return of the gradients.  record = [source row (int32 bits), v_xy(2), v_depth, v_conic(3),
// v_opacity, v_colour(ch)] = 8 + ch words, appended to the region this rank owns in the source rank's buffer.
// -------------------------------------------------------------------------------------------------------------
struct ReturnArgs {
    int64_t n, cam_offset;                    // cam_offset: first global camera of this rank (camera owners), else 0
    int ch, world;
    const float2* g_means2d; const float* g_depths; const float* g_conics; const float* g_opac; const float* g_colors;
    const int64_t* camera_ids; const int64_t* gaussian_ids;
    int64_t recv_prefix[B2S_MAX_RANKS + 1];   // rows [recv_prefix[s], recv_prefix[s+1]) came from rank s
    int64_t gid_offset[B2S_MAX_RANKS];        // first global Gaussian id of rank s
    int64_t n_src[B2S_MAX_RANKS];             // Gaussians of rank s (rows per camera of its dense layout)
    int64_t region_base[B2S_MAX_RANKS];       // first record slot of this rank's region inside rank s's buffer
    unsigned long long* counters;             // [world], zeroed: records appended per source rank
    float* peer[B2S_MAX_RANKS];
};

__global__ void __launch_bounds__(ST_THREADS) stripe_return_kernel(const ReturnArgs a) {
    const int64_t j = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float g[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool nz = false;
    if (j < a.n) {
        // every load is issued before the first test (one memory round trip per row, not two)
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (a.g_colors && a.ch == 3) { c0 = a.g_colors[3 * j]; c1 = a.g_colors[3 * j + 1]; c2 = a.g_colors[3 * j + 2]; }
        if (a.g_means2d) { const float2 v = a.g_means2d[j]; g[0] = v.x; g[1] = v.y; }
        if (a.g_depths) g[2] = a.g_depths[j];
        if (a.g_conics) { g[3] = a.g_conics[3 * j]; g[4] = a.g_conics[3 * j + 1]; g[5] = a.g_conics[3 * j + 2]; }
        if (a.g_opac) g[6] = a.g_opac[j];
#pragma unroll
        for (int k = 0; k < 7; ++k) nz |= g[k] != 0.f;
        nz |= (c0 != 0.f) | (c1 != 0.f) | (c2 != 0.f);
        if (a.g_colors && a.ch != 3 && !nz)
            for (int k = 0; k < a.ch; ++k) nz |= a.g_colors[j * a.ch + k] != 0.f;
    }
    int src = 0;
    if (nz)
        while (src + 1 < a.world && j >= a.recv_prefix[src + 1]) ++src;
    unsigned todo = __ballot_sync(FULL, nz);
    const unsigned lt = (1u << lane) - 1u;
    const int G = 8 + a.ch;
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int s = __shfl_sync(FULL, src, leader);
        const unsigned grp = __ballot_sync(FULL, nz && src == s);
        unsigned long long base = 0ull;
        if (lane == leader) base = atomicAdd(&a.counters[s], (unsigned long long)__popc(grp));
        base = __shfl_sync(FULL, base, leader);
        if (nz && src == s) {
            const int64_t slot = a.region_base[s] + (int64_t)base + __popc(grp & lt);
            float* o = a.peer[s] + slot * G;
            const int64_t row = (a.camera_ids[j] + a.cam_offset) * a.n_src[s] + (a.gaussian_ids[j] - a.gid_offset[s]);
            o[0] = __int_as_float((int)row);
#pragma unroll
            for (int k = 0; k < 7; ++k) o[1 + k] = g[k];
            for (int k = 0; k < a.ch; ++k) o[8 + k] = a.g_colors ? a.g_colors[j * a.ch + k] : 0.f;
        }
        todo &= ~grp;
    }
}

struct ScatterArgs {
    int ch, world;
    const float* buf;                          // this rank's return buffer
    const unsigned long long* counts;          // [world]: records rank d appended (written by rank d)
    int64_t region_base[B2S_MAX_RANKS];        // first record slot of rank d's region
    int64_t region_cap[B2S_MAX_RANKS];         // rows this rank sent to rank d (capacity of the region)
    float* v_means2d; float* v_depths; float* v_conics; float* v_opac; float* v_colors;
};

__global__ void __launch_bounds__(ST_THREADS) stripe_scatter_add_kernel(const ScatterArgs a) {
    const int d = blockIdx.y;
    const int64_t cnt = min((int64_t)a.counts[d], a.region_cap[d]);
    const int G = 8 + a.ch;
    for (int64_t k = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; k < cnt; k += (int64_t)gridDim.x * ST_THREADS) {
        const float* p = a.buf + (a.region_base[d] + k) * G;
        const int64_t r = (int64_t)__float_as_int(p[0]);
        if (p[1] != 0.f) atomicAdd(a.v_means2d + 2 * r, p[1]);
        if (p[2] != 0.f) atomicAdd(a.v_means2d + 2 * r + 1, p[2]);
        if (p[3] != 0.f) atomicAdd(a.v_depths + r, p[3]);
        if (p[4] != 0.f) atomicAdd(a.v_conics + 3 * r, p[4]);
        if (p[5] != 0.f) atomicAdd(a.v_conics + 3 * r + 1, p[5]);
        if (p[6] != 0.f) atomicAdd(a.v_conics + 3 * r + 2, p[6]);
        if (p[7] != 0.f) atomicAdd(a.v_opac + r, p[7]);
        for (int c = 0; c < a.ch; ++c)
            if (p[8 + c] != 0.f) atomicAdd(a.v_colors + r * a.ch + c, p[8 + c]);
    }
}

static bool fill_stripes(StripeTable& sb, const b2s_stripes* s, int tile_h) {
    if (!s || s->world < 1 || s->world > B2S_MAX_RANKS || s->cams_per_rank < 0) return false;
    sb.world = s->world;
    sb.cams_per_rank = s->cams_per_rank;
    if (s->cams_per_rank > 0) {
        for (int d = 0; d <= B2S_MAX_RANKS; ++d) sb.bounds[d] = 0;
        return true;
    }
    for (int d = 0; d <= s->world; ++d) {
        sb.bounds[d] = s->bounds[d];
        if (d > 0 && sb.bounds[d] < sb.bounds[d - 1]) return false;
    }
    return sb.bounds[0] == 0 && sb.bounds[s->world] == tile_h;
}

}  // namespace b2s

using namespace b2s;

extern "C" int64_t b2s_stripe_ctas(int64_t M) { return cdiv(M > 0 ? M : 1, ST_THREADS); }

extern "C" int b2s_stripe_count(int64_t M, int64_t N, const float* means2d, const int32_t* radii, int32_t tile_size,
                                int32_t tile_w, int32_t tile_h, const b2s_stripes* stripes, uint16_t* dest_mask, uint32_t* cta_counts,
                                int64_t* totals, int64_t tag, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && N >= 0 && (N > 0 || M == 0) && tile_size > 0 && tile_w > 0 && tile_h > 0 && totals && cta_counts, "bad args");
    if (N == 0) N = 1;      // an empty shard (a rank that pruned all of its Gaussians): zero counts, no division by zero
    StripeTable sb;
    B2S_REQUIRE(fill_stripes(sb, stripes, tile_h), "stripes: 1..16 ranks, ascending tile rows from 0 to tile_h");
    B2S_REQUIRE(sb.cams_per_rank == 0 || M <= N * sb.cams_per_rank * sb.world, "more cameras than cams_per_rank * world");
    cudaStream_t st = (cudaStream_t)stream;
    B2S_CUDA(cudaMemsetAsync(totals, 0, (size_t)sb.world * 4 * sizeof(int64_t), st));
    const int64_t n_ctas = b2s_stripe_ctas(M);
    B2S_REQUIRE(M == 0 || (means2d && radii && dest_mask), "null pointer");
    stripe_count_kernel<<<(unsigned)n_ctas, ST_THREADS, 0, st>>>(M, reinterpret_cast<const float2*>(means2d), radii, N,
                                                                (float)tile_size, tile_w, tile_h, sb, dest_mask, cta_counts, n_ctas,
                                                                reinterpret_cast<unsigned long long*>(totals), (long long)tag);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_stripe_scan(int64_t M, int32_t world, uint32_t* cta_counts, int64_t* totals, b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && world >= 1 && world <= B2S_MAX_RANKS && cta_counts && totals, "bad args");
    stripe_scan_kernel<<<world, SCAN_T, 0, (cudaStream_t)stream>>>(cta_counts, b2s_stripe_ctas(M),
                                                                  reinterpret_cast<unsigned long long*>(totals));
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_p2p_put(int32_t world, const int64_t* src, int32_t words_per_peer, int32_t src_stride,
                           void* const* host_peer_dst, b2s_stream_t stream) {
    B2S_REQUIRE(world >= 1 && world <= B2S_MAX_RANKS && src && words_per_peer > 0 && src_stride >= 0 && host_peer_dst, "bad args");
    PeerPtrs pp;
    for (int d = 0; d < B2S_MAX_RANKS; ++d) pp.p[d] = d < world ? host_peer_dst[d] : nullptr;
    p2p_put_kernel<<<world, 128, 0, (cudaStream_t)stream>>>(world, reinterpret_cast<const unsigned long long*>(src), words_per_peer,
                                                           src_stride, pp);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_stripe_pack(int64_t M, int64_t N, int32_t channels, int64_t gid_offset, const float* means2d,
                               const float* depths, const float* conics, const float* opacities, const float* colors,
                               const int32_t* radii, const uint16_t* dest_mask, const uint32_t* cta_prefix, int32_t world,
                               int32_t cams_per_rank, const int64_t* host_dest_base, void* const* host_peer_buffers,
                               b2s_stream_t stream) {
    B2S_REQUIRE(M >= 0 && N >= 0 && channels >= 0 && world >= 1 && world <= B2S_MAX_RANKS, "bad sizes");
    if (M == 0) return 0;
    B2S_REQUIRE(N > 0, "rows without Gaussians");
    B2S_REQUIRE(means2d && depths && conics && opacities && colors && radii && dest_mask && cta_prefix, "null pointer");
    B2S_REQUIRE(host_dest_base && host_peer_buffers, "null routing table");
    PackArgs a;
    a.M = M; a.N = N; a.n_ctas = b2s_stripe_ctas(M); a.gid_offset = gid_offset;
    a.ch = channels; a.world = world; a.cams_per_rank = cams_per_rank;
    a.means2d = reinterpret_cast<const float2*>(means2d); a.depths = depths; a.conics = conics; a.opac = opacities;
    a.colors = colors; a.radii = radii; a.dest_mask = dest_mask; a.cta_prefix = cta_prefix;
    for (int d = 0; d < B2S_MAX_RANKS; ++d) {
        a.base[d] = d < world ? host_dest_base[d] : 0;
        a.peer[d] = d < world ? reinterpret_cast<float*>(host_peer_buffers[d]) : nullptr;
    }
    stripe_pack_kernel<<<(unsigned)a.n_ctas, ST_THREADS, 0, (cudaStream_t)stream>>>(a);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_stripe_unpack(int64_t n, int32_t channels, const float* recv_buffer, int32_t tile_size,
                                 int32_t tile_w, int32_t tile_h, const b2s_tile_rows* rows, float* means2d, float* depths,
                                 float* conics, float* opacities, float* colors, int32_t* radii, int64_t* camera_ids,
                                 int64_t* gaussian_ids, int32_t* tiles_per_gauss, uint32_t* depth_keys, int32_t* row_ids,
                                 b2s_stream_t stream) {
    B2S_REQUIRE(n >= 0 && channels >= 0 && tile_size > 0 && tile_w > 0 && tile_h > 0, "bad sizes");
    if (n == 0) return 0;
    B2S_REQUIRE(recv_buffer && means2d && depths && conics && opacities && colors && radii && camera_ids && gaussian_ids &&
                tiles_per_gauss && depth_keys && row_ids, "null pointer");
    int lo = 0, hi = tile_h;
    if (rows) { lo = rows->row_lo; hi = rows->row_hi; }
    B2S_REQUIRE(lo >= 0 && hi >= lo && hi <= tile_h, "bad tile-row window");
    stripe_unpack_kernel<<<(unsigned)cdiv(n, ST_THREADS), ST_THREADS, 0, (cudaStream_t)stream>>>(
        n, channels, recv_buffer, (float)tile_size, tile_w, lo, hi, reinterpret_cast<float2*>(means2d), depths, conics,
        opacities, colors, radii, camera_ids, gaussian_ids, tiles_per_gauss, depth_keys, row_ids);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_stripe_return(int64_t n, int32_t channels, const float* g_means2d, const float* g_depths,
                                 const float* g_conics, const float* g_opacities, const float* g_colors,
                                 const int64_t* camera_ids, const int64_t* gaussian_ids, int64_t cam_offset, int32_t world,
                                 const int64_t* host_recv_prefix, const int64_t* host_gid_offset, const int64_t* host_n_src,
                                 const int64_t* host_region_base, int64_t* counters, void* const* host_peer_buffers,
                                 b2s_stream_t stream) {
    B2S_REQUIRE(n >= 0 && channels >= 0 && world >= 1 && world <= B2S_MAX_RANKS && counters, "bad args");
    cudaStream_t st = (cudaStream_t)stream;
    B2S_CUDA(cudaMemsetAsync(counters, 0, (size_t)world * 8, st));
    if (n == 0) return 0;
    B2S_REQUIRE(camera_ids && gaussian_ids && host_recv_prefix && host_gid_offset && host_n_src && host_region_base &&
                host_peer_buffers, "null pointer");
    ReturnArgs a;
    a.n = n; a.cam_offset = cam_offset; a.ch = channels; a.world = world;
    a.g_means2d = reinterpret_cast<const float2*>(g_means2d); a.g_depths = g_depths; a.g_conics = g_conics;
    a.g_opac = g_opacities; a.g_colors = g_colors; a.camera_ids = camera_ids; a.gaussian_ids = gaussian_ids;
    a.counters = reinterpret_cast<unsigned long long*>(counters);
    for (int d = 0; d < B2S_MAX_RANKS; ++d) {
        a.recv_prefix[d] = d <= world ? host_recv_prefix[d] : 0;
        a.gid_offset[d] = d < world ? host_gid_offset[d] : 0;
        a.n_src[d] = d < world ? host_n_src[d] : 0;
        a.region_base[d] = d < world ? host_region_base[d] : 0;
        a.peer[d] = d < world ? reinterpret_cast<float*>(host_peer_buffers[d]) : nullptr;
    }
    a.recv_prefix[B2S_MAX_RANKS] = world == B2S_MAX_RANKS ? host_recv_prefix[world] : 0;
    B2S_REQUIRE(host_recv_prefix[world] == n, "receive prefix does not end at n");
    stripe_return_kernel<<<(unsigned)cdiv(n, ST_THREADS), ST_THREADS, 0, st>>>(a);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_stripe_scatter_add(int32_t world, int32_t channels, const float* ret_buffer, const int64_t* host_region_base,
                                      const int64_t* host_region_cap, const int64_t* ret_counts, float* v_means2d,
                                      float* v_depths, float* v_conics, float* v_opacities, float* v_colors,
                                      b2s_stream_t stream) {
    B2S_REQUIRE(world >= 1 && world <= B2S_MAX_RANKS && channels >= 0, "bad args");
    B2S_REQUIRE(ret_buffer && host_region_base && host_region_cap && ret_counts, "null pointer");
    B2S_REQUIRE(v_means2d && v_depths && v_conics && v_opacities && v_colors, "null gradient output");
    ScatterArgs a;
    a.ch = channels; a.world = world; a.buf = ret_buffer;
    a.counts = reinterpret_cast<const unsigned long long*>(ret_counts);
    int64_t max_cap = 0;
    for (int d = 0; d < B2S_MAX_RANKS; ++d) {
        a.region_base[d] = d < world ? host_region_base[d] : 0;
        a.region_cap[d] = d < world ? host_region_cap[d] : 0;
        if (a.region_cap[d] > max_cap) max_cap = a.region_cap[d];
    }
    if (max_cap == 0) return 0;
    a.v_means2d = v_means2d; a.v_depths = v_depths; a.v_conics = v_conics; a.v_opac = v_opacities; a.v_colors = v_colors;
    dim3 grid((unsigned)min((int64_t)148 * 4, cdiv(max_cap, ST_THREADS)), (unsigned)world);
    stripe_scatter_add_kernel<<<grid, ST_THREADS, 0, (cudaStream_t)stream>>>(a);
    B2S_LAUNCH_CHECK();
    return 0;
}
