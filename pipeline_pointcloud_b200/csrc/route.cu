// Stage 1 fused with the multi-GPU exchange ("direct" route; SURVEY.md 8e, BASELINE configs 3-5).
//
// `rasterization(distributed=True)` -- the scheme the reference's multi-GPU branch runs through gsplat
// (source/container/src/main.py:1317-1347) -- projects a rank's shard of Gaussians into every camera and hands each
// visible splat to the rank that composites it (camera owner, or owner of a tile-row stripe).  Upstream does
// projection -> pack -> NCCL all-to-all (counts, then payload) -> split.  Here ONE kernel does the projection, the SH
// colour and the exchange: every CTA stores its visible rows, as finished 52-byte records, straight into the
// destination rank's receive buffer over the NVLink peer mapping (symmetric memory), while the next CTAs are still
// projecting -- the transfer overlaps the arithmetic tile by tile and nothing is staged in local HBM first.
//
// No counts are needed up front because every (source rank, camera) pair owns a fixed-capacity REGION of the
// receiver's buffer and every source CTA a fixed 256-row slot inside it:
//     slot of row (camera c, Gaussian n) for destination d = region(src, c) * cap + (n / 256) * 256 + rank inside the CTA
// so placement is deterministic (ascending Gaussian index inside a region: the reference's tie-break order) and needs
// no atomics, no look-back and no host round trip.  The CTA also stores its per-destination counts (rows, tile
// intersections, coarse-bin records) into the receiver's count table.  After ONE barrier the receiver scans the count
// table (route_scan) and gathers the occupied slots into compact arrays (route_gather), deriving the tile counts,
// depth keys and row ids of its binning in the same pass.  Regions that are mostly empty cost address space (180 GB
// of HBM per GPU), not bandwidth: only occupied slots are ever read.
//
// Both buffers are double-buffered by the caller (parity), which is what lets a single barrier per exchange suffice.
#include <string.h>

#include "project_common.cuh"

namespace b2s {

constexpr int RT_MAXCH = 4;                    // colour channels carried in a record
constexpr int RT_MAXF = 10 + RT_MAXCH;         // [x, y, depth, conic(3), opacity, colour(ch), radius, camera, local Gaussian index]
constexpr int CBR = B2S_COARSE;

struct RouteDev {
    int world, rank, cams_per_rank, region_cams, ch, color_mode, n_color_in, depth_channel, overflow;
    int bounds[B2S_MAX_RANKS + 1];
    long long cap_rows, n_ctas_cap;
    const float* colors_in;
    float* rec[B2S_MAX_RANKS];
    int32_t* cnt[B2S_MAX_RANKS];
};

__device__ __forceinline__ int coarse_span_r(int t0, int t1) { return (t1 + CBR - 1) / CBR - t0 / CBR; }

// SHMODE: 0 = no SH (colours from colors_in / depth), 1 = SH through the LDG shared-memory tile (any degree / stride),
//         3 = degree-3 rows through the tensor-memory accelerator (swizzled boxes, see project.cu)
template <int SHMODE>
__global__ void __launch_bounds__(FWD_THREADS) project_route_kernel(const b2s_project_fwd_args a, const RouteDev rp,
                                                                    const __grid_constant__ CUtensorMap map_a,
                                                                    const __grid_constant__ CUtensorMap map_b) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_w[3][FWD_THREADS / 32];
    const float W = (float)a.width, H = (float)a.height;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = (a.sh_degree + 1) * (a.sh_degree + 1);

    // ---- shared-memory carve-up: [coefficient tiles][staging of one destination's records][cameras]
    unsigned char* p = smem_raw;
    unsigned char *tile_a = nullptr, *tile_b = nullptr;
    float* sh_tile = nullptr;
    if (SHMODE == 3) {
        const uint32_t raw = smem_u32(smem_raw);
        p += ((raw + 1023u) & ~1023u) - raw;
        tile_a = p; tile_b = p + TMA_A_BYTES;
        p += TMA_A_BYTES + TMA_B_BYTES;
    } else if (SHMODE == 1) {
        sh_tile = reinterpret_cast<float*>(p) + (size_t)warp * 32 * sh_row_stride(K);
        p += (size_t)(FWD_THREADS / 32) * 32 * sh_row_stride(K) * sizeof(float);
        p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    }
    float* s_stage = reinterpret_cast<float*>(p);
    p += (size_t)FWD_THREADS * RT_MAXF * sizeof(float);
    Camera* cams = reinterpret_cast<Camera*>(p);

    const uint32_t mbar = smem_u32(&s_mbar);
    if (SHMODE == 3 && threadIdx.x == 0) mbar_init(mbar, 1);
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, W, H);
    const int64_t row0 = (int64_t)blockIdx.x * FWD_THREADS;
    if (SHMODE == 3 && threadIdx.x == 0) {
        mbar_expect_tx(mbar, TMA_A_BYTES + TMA_B_BYTES);
        tma_load_2d(smem_u32(tile_a), &map_a, 0, (int)row0, mbar);
        tma_load_2d(smem_u32(tile_b), &map_b, 32, (int)row0, mbar);
    }

    const int64_t n = row0 + threadIdx.x;
    const bool live = n < a.N;
    const int64_t nc = live ? n : (a.N > 0 ? a.N - 1 : 0);
    float mean[3] = {0.f, 0.f, 0.f}, q[4] = {1.f, 0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
    float opac = 1.f;
    if (a.N > 0) {
        mean[0] = __ldg(a.means + 3 * nc); mean[1] = __ldg(a.means + 3 * nc + 1); mean[2] = __ldg(a.means + 3 * nc + 2);
        const float4 qq = __ldg(reinterpret_cast<const float4*>(a.quats) + nc);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
        s[0] = __ldg(a.scales + 3 * nc); s[1] = __ldg(a.scales + 3 * nc + 1); s[2] = __ldg(a.scales + 3 * nc + 2);
        if (a.opacities) opac = __ldg(a.opacities + nc);
    }
    Rot R; float qn[4], qnorm, M[9]; Sym3 S;
    quat_to_rotmat(q, R, qn, qnorm);
    covar_from_rot_scale(R, s, M, S);

    // the coefficient rows of the CTA (all live rows: visibility is not known yet, and in a sharded job nearly every
    // Gaussian is visible in some camera)
    if (SHMODE == 1) {
        const int64_t wrow0 = row0 + warp * 32;
        const int rows = (int)max((int64_t)0, min((int64_t)32, a.N - wrow0));
        if (rows > 0)
            sh_rows_load<0>(sh_tile, a.sh_coeffs + (size_t)wrow0 * a.sh_stride_k * 3, rows, __ballot_sync(FULL, live), K, a.sh_stride_k, lane);
        __syncwarp();
    }
    if (SHMODE == 3) mbar_wait(mbar, 0);

    const int F = 10 + rp.ch;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned r = threadIdx.x;
    for (int c = 0; c < a.C; ++c) {
        ProjOut o;
        project_one(cams[c], mean, S, W, H, a.eps2d, a.near_plane, a.far_plane, a.radius_clip, o, nullptr);
        if (live) a.radii[(int64_t)c * a.N + n] = o.radius;
        int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        bool vis = live && o.radius > 0 && !rp.overflow;
        if (vis) {
            tile_rect(o.x2d, o.y2d, o.radius, (float)a.tile_size, a.tile_w, a.tile_h, x0, y0, x1, y1);
            vis = x1 > x0 && y1 > y0;
        }
        float rec[RT_MAXF];
#pragma unroll
        for (int k = 0; k < RT_MAXF; ++k) rec[k] = 0.f;
        if (vis) {
            rec[0] = o.x2d; rec[1] = o.y2d; rec[2] = o.depth;
            rec[3] = o.conic_a; rec[4] = o.conic_b; rec[5] = o.conic_c;
            rec[6] = a.antialiased ? opac * o.comp : opac;
            float col[RT_MAXCH] = {0.f, 0.f, 0.f, 0.f};
            int nc_ = 0;
            if (SHMODE != 0) {
                const float dx = mean[0] - cams[c].campos[0], dy = mean[1] - cams[c].campos[1], dz = mean[2] - cams[c].campos[2];
                const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
                float Y[16];
                sh_basis(SHMODE == 3 ? 3 : a.sh_degree, dx * inv, dy * inv, dz * inv, Y);
                float cr = 0.f, cg = 0.f, cb = 0.f;
                if (SHMODE == 3) {
                    const float4* row_a = reinterpret_cast<const float4*>(tile_a + r * 128);
                    const float4* row_b = reinterpret_cast<const float4*>(tile_b + r * 64);
                    const unsigned xa = r & 7u, xb = (r >> 1) & 3u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float4 chk[3];
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const int j = 3 * g + t;
                            chk[t] = j < 8 ? row_a[(unsigned)j ^ xa] : row_b[(unsigned)(j - 8) ^ xb];
                        }
                        const float4 c0 = chk[0], c1 = chk[1], c2 = chk[2];
                        cr += Y[4 * g] * c0.x; cg += Y[4 * g] * c0.y; cb += Y[4 * g] * c0.z;
                        cr += Y[4 * g + 1] * c0.w; cg += Y[4 * g + 1] * c1.x; cb += Y[4 * g + 1] * c1.y;
                        cr += Y[4 * g + 2] * c1.z; cg += Y[4 * g + 2] * c1.w; cb += Y[4 * g + 2] * c2.x;
                        cr += Y[4 * g + 3] * c2.y; cg += Y[4 * g + 3] * c2.z; cb += Y[4 * g + 3] * c2.w;
                    }
                } else {
                    const float* row = sh_tile + lane * sh_row_stride(K);
#pragma unroll 4
                    for (int k = 0; k < K; ++k) { cr += Y[k] * row[3 * k]; cg += Y[k] * row[3 * k + 1]; cb += Y[k] * row[3 * k + 2]; }
                }
                col[0] = fmaxf(cr + 0.5f, 0.f); col[1] = fmaxf(cg + 0.5f, 0.f); col[2] = fmaxf(cb + 0.5f, 0.f);
                nc_ = 3;
            } else if (rp.color_mode == 1) {
#pragma unroll
                for (int k = 0; k < RT_MAXCH; ++k)
                    if (k < rp.n_color_in) col[k] = __ldg(rp.colors_in + n * rp.n_color_in + k);
                nc_ = rp.n_color_in;
            }
            if (rp.depth_channel) {
#pragma unroll
                for (int k = 0; k < RT_MAXCH; ++k)
                    if (k == nc_) col[k] = o.depth;
            }
            const int lc = rp.cams_per_rank > 0 ? c % rp.cams_per_rank : c;
#pragma unroll
            for (int k = 0; k < RT_MAXCH; ++k)
                if (k < rp.ch) rec[7 + k] = col[k];
#pragma unroll
            for (int k = 0; k <= RT_MAXCH; ++k)           // the three integers follow the colours
                if (k == rp.ch) { rec[7 + k] = __int_as_float(o.radius); rec[8 + k] = __int_as_float(lc); rec[9 + k] = __int_as_float((int)n); }
        }

        // ---- destinations of camera c: its owner, or every stripe
        const int d_lo = rp.cams_per_rank > 0 ? c / rp.cams_per_rank : 0;
        const int d_hi = rp.cams_per_rank > 0 ? d_lo + 1 : rp.world;
        const int lc_u = rp.cams_per_rank > 0 ? c % rp.cams_per_rank : c;
        const long long reg = (long long)rp.rank * rp.region_cams + lc_u;
        for (int d = d_lo; d < d_hi; ++d) {
            const int lo = rp.cams_per_rank > 0 ? 0 : rp.bounds[d], hi = rp.cams_per_rank > 0 ? a.tile_h : rp.bounds[d + 1];
            const int cy0 = min(max(y0, lo), hi), cy1 = min(max(y1, lo), hi);
            const bool hit = vis && cy1 > cy0;
            const unsigned b = __ballot_sync(FULL, hit);
            int tiles = hit ? (x1 - x0) * (cy1 - cy0) : 0;
            int coarse = hit ? coarse_span_r(x0, x1) * coarse_span_r(cy0, cy1) : 0;
            tiles = __reduce_add_sync(FULL, tiles);
            coarse = __reduce_add_sync(FULL, coarse);
            if (lane == 0) { s_w[0][warp] = __popc(b); s_w[1][warp] = tiles; s_w[2][warp] = coarse; }
            __syncthreads();
            int before = 0, total = 0, t_tiles = 0, t_coarse = 0;
#pragma unroll
            for (int w = 0; w < FWD_THREADS / 32; ++w) {
                const int x = s_w[0][w];
                if (w < warp) before += x;
                total += x; t_tiles += s_w[1][w]; t_coarse += s_w[2][w];
            }
            if (hit) {
                float* sr = s_stage + (before + __popc(b & lt)) * F;
#pragma unroll
                for (int k = 0; k < RT_MAXF; ++k)
                    if (k < F) sr[k] = rec[k];
            }
            __syncthreads();
            // one contiguous, coalesced run of `total` records into the destination's slot of this CTA
            // (128-bit peer stores; the slot starts 16-byte aligned and the last store may run up to 3 floats past the
            // CTA's records, still inside its own 256-row slot)
            float4* dst = reinterpret_cast<float4*>(rp.rec[d] + ((reg * rp.cap_rows + row0) * (long long)F));
            const float4* src4 = reinterpret_cast<const float4*>(s_stage);
            for (int t = threadIdx.x; t < ((total * F + 3) >> 2); t += FWD_THREADS) dst[t] = src4[t];
            if (threadIdx.x == 0) {
                int4 e = make_int4(total, t_tiles, t_coarse, blockIdx.x == 0 ? (int)a.N : 0);
                *reinterpret_cast<int4*>(rp.cnt[d] + (reg * rp.n_ctas_cap + blockIdx.x) * 4) = e;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// receiver: scan of the count table, one CTA per source rank.  row_base[region][source CTA] = exclusive prefix of the row
// counts INSIDE the source (camera-major, then source CTA); info[4 + 4 s ..] = (rows, Gaussian count, tile intersections,
// coarse-bin records) of source s -- the host adds the sources up when it reads them (its one sync).
// Every thread takes 8 consecutive entries per sweep (independent 128-bit loads), so a source of 8 K CTAs is one sweep.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 1024, SC_PER = 8;
__global__ void __launch_bounds__(SC_THREADS) route_scan_kernel(int region_cams, long long n_ctas_cap, const int32_t* __restrict__ cnt,
                                                                int32_t* __restrict__ row_base, long long* __restrict__ info) {
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned long long s_tl[32], s_co[32];
    __shared__ unsigned long long s_carry, s_tiles, s_coarse;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int src = blockIdx.x;
    const long long n_src = cnt[((long long)src * region_cams * n_ctas_cap) * 4 + 3];     // tag of (src, camera 0, CTA 0)
    const long long n_ctas = min((max(n_src, 0LL) + FWD_THREADS - 1) / FWD_THREADS, n_ctas_cap);
    const long long items = n_ctas * region_cams;
    if (threadIdx.x == 0) { s_carry = 0ull; s_tiles = 0ull; s_coarse = 0ull; }
    __syncthreads();
    for (long long base = 0; base < items; base += (long long)SC_THREADS * SC_PER) {
        const long long k0 = base + (long long)threadIdx.x * SC_PER;
        unsigned rows[SC_PER];
        long long addr[SC_PER];
        unsigned tsum = 0;
        unsigned long long tl = 0, co = 0;
#pragma unroll
        for (int i = 0; i < SC_PER; ++i) {
            const long long k = k0 + i;
            int4 e = make_int4(0, 0, 0, 0);
            addr[i] = -1;
            if (k < items) {
                const long long lc = k / n_ctas, b = k - lc * n_ctas;
                addr[i] = ((long long)src * region_cams + lc) * n_ctas_cap + b;
                e = *reinterpret_cast<const int4*>(cnt + addr[i] * 4);
            }
            rows[i] = (unsigned)e.x;
            tsum += rows[i];
            tl += (unsigned long long)e.y;
            co += (unsigned long long)e.z;
        }
        unsigned incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { tl += __shfl_xor_sync(FULL, tl, d); co += __shfl_xor_sync(FULL, co, d); }
        if (lane == 31) s_wsum[warp] = incl;
        if (lane == 0) { s_tl[warp] = tl; s_co[warp] = co; }
        __syncthreads();
        if (warp == 0) {
            const unsigned w = s_wsum[lane];
            unsigned wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(FULL, wi, d);
                if (lane >= d) wi += t;
            }
            s_wsum[lane] = wi - w;                     // exclusive prefix of the warp totals
            unsigned long long a = s_tl[lane], c = s_co[lane];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(FULL, a, d); c += __shfl_xor_sync(FULL, c, d); }
            if (lane == 0) { s_tiles += a; s_coarse += c; }
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        unsigned run = (unsigned)carry + s_wsum[warp] + incl - tsum;
#pragma unroll
        for (int i = 0; i < SC_PER; ++i) {
            if (addr[i] >= 0) row_base[addr[i]] = (int32_t)run;
            run += rows[i];
        }
        __syncthreads();
        if (threadIdx.x == SC_THREADS - 1) s_carry = carry + s_wsum[31] + incl;      // thread 1023 holds the sweep's last inclusive sum
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        info[4 + 4 * src] = (long long)s_carry; info[4 + 4 * src + 1] = n_src;
        info[4 + 4 * src + 2] = (long long)s_tiles; info[4 + 4 * src + 3] = (long long)s_coarse;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// receiver: occupied slots -> compact tensors (one CTA per source CTA, one warp per 32 of its rows), plus what binning
// needs.  A source CTA's rows are one contiguous run of records: each warp moves its 32 records through shared memory
// with coalesced 128-bit loads, then every lane takes its row apart and the fields leave as coalesced stores.
// ---------------------------------------------------------------------------------------------------------------------
struct GatherArgs {
    int world, region_cams, ch;
    long long cap_rows, n_ctas_cap;
    long long n_src[B2S_MAX_RANKS], gid_offset[B2S_MAX_RANKS], row_offset[B2S_MAX_RANKS], cta_base[B2S_MAX_RANKS + 1];
    const float* rec; const int32_t* cnt; const int32_t* row_base;
    float tile_size; int tile_w, row_lo, row_hi;
    float2* means2d; float* depths; float* conics; float* opac; float* colors; int32_t* radii;
    int64_t* camera_ids; int64_t* gaussian_ids; int32_t* tiles; uint32_t* depth_keys; int32_t* row_ids;
};

template <int WPS>      // warps per slot: 8 when the slots are nearly full (camera owners), fewer when they are sparse (tile stripes)
__global__ void __launch_bounds__(256) route_gather_kernel(const GatherArgs a) {
    __shared__ __align__(16) float s_rows[8][32 * RT_MAXF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // item = (source, camera, source CTA), flattened; 8 / WPS items per CTA; warp w of an item's group: rows [32 w, ..) step 32 WPS
    const long long item = (long long)blockIdx.x * (8 / WPS) + warp / WPS;
    if (item >= a.cta_base[a.world]) return;
    int src = 0;
    while (src + 1 < a.world && item >= a.cta_base[src + 1]) ++src;
    const long long n_ctas = (a.n_src[src] + FWD_THREADS - 1) / FWD_THREADS;
    const long long local = item - a.cta_base[src];
    const int lc = (int)(local / n_ctas);
    const long long b = local - (long long)lc * n_ctas;
    const long long reg = (long long)src * a.region_cams + lc;
    const int count = a.cnt[(reg * a.n_ctas_cap + b) * 4];
    const long long base = a.row_offset[src] + a.row_base[reg * a.n_ctas_cap + b];
    const int F = 10 + a.ch;
    const float* slot = a.rec + (reg * a.cap_rows + b * FWD_THREADS) * (long long)F;     // 16-byte aligned: 256 F floats per slot
    float* sm = s_rows[warp];
    for (int t0 = (warp % WPS) * 32; t0 < count; t0 += 32 * WPS) {
        const int rows = min(32, count - t0);
        const int n4 = (rows * F + 3) >> 2;                             // (reads at most 3 floats past the chunk, inside the slot)
        const float4* g4 = reinterpret_cast<const float4*>(slot + (long long)t0 * F);
        float4 v[4];                                                    // 32 rows x at most 14 words = at most 112 float4: four per lane,
#pragma unroll                                                          // all in flight before the first shared-memory store
        for (int i = 0; i < 4; ++i) v[i] = (lane + 32 * i < n4) ? ldg_stream4(g4 + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (lane + 32 * i < n4) reinterpret_cast<float4*>(sm)[lane + 32 * i] = v[i];
        __syncwarp();
        if (lane < rows) {
            const float* p = sm + lane * F;
            const long long j = base + t0 + lane;
            const float2 m = make_float2(p[0], p[1]);
            const float depth = p[2];
            a.means2d[j] = m;
            a.depths[j] = depth;
            a.conics[3 * j] = p[3]; a.conics[3 * j + 1] = p[4]; a.conics[3 * j + 2] = p[5];
            a.opac[j] = p[6];
            for (int k = 0; k < a.ch; ++k) a.colors[j * a.ch + k] = p[7 + k];
            const int radius = __float_as_int(p[7 + a.ch]);
            a.radii[j] = radius;
            a.camera_ids[j] = (int64_t)__float_as_int(p[8 + a.ch]);
            a.gaussian_ids[j] = a.gid_offset[src] + (int64_t)__float_as_int(p[9 + a.ch]);
            int x0, y0, x1, y1;
            tile_rect(m.x, m.y, radius, a.tile_size, a.tile_w, a.row_lo, a.row_hi, x0, y0, x1, y1);
            a.tiles[j] = max(0, x1 - x0) * max(0, y1 - y0);
            a.depth_keys[j] = radius > 0 ? __float_as_uint(depth) : 0xFFFFFFFFu;
            a.row_ids[j] = (int32_t)j;
        }
        __syncwarp();
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int64_t b2s_route_rec_floats(int32_t world, int32_t region_cams, int64_t cap_rows) {
    return (int64_t)world * region_cams * cap_rows * RT_MAXF;
}
extern "C" int64_t b2s_route_cnt_ints(int32_t world, int32_t region_cams, int64_t cap_rows) {
    return (int64_t)world * region_cams * (cap_rows / FWD_THREADS) * 4;
}

extern "C" int b2s_route_project(const b2s_route_args* ra, b2s_stream_t stream) {
    B2S_REQUIRE(ra != nullptr, "null args");
    const b2s_project_fwd_args* a = &ra->proj;
    B2S_REQUIRE(a->N >= 0 && a->C > 0 && a->C <= B2S_MAX_CAMERAS_PER_CALL, "bad sizes");
    B2S_REQUIRE(ra->world >= 1 && ra->world <= B2S_MAX_RANKS && ra->rank >= 0 && ra->rank < ra->world, "bad world / rank");
    B2S_REQUIRE(ra->channels >= 1 && ra->channels <= RT_MAXCH, "route: 1..4 colour channels");
    B2S_REQUIRE(ra->cap_rows > 0 && ra->cap_rows % FWD_THREADS == 0, "cap_rows must be a positive multiple of 256");
    B2S_REQUIRE(ra->region_cams > 0 && (ra->cams_per_rank == 0 ? ra->region_cams == a->C
                                                                : (ra->region_cams == ra->cams_per_rank && a->C == ra->cams_per_rank * ra->world)),
                "region_cams / cams_per_rank / C mismatch");
    B2S_REQUIRE(a->tile_size > 0 && a->tile_w > 0 && a->tile_h > 0, "bad tile geometry");
    B2S_REQUIRE(a->N == 0 || (a->means && a->quats && a->scales && a->radii), "null input");
    B2S_REQUIRE(a->viewmats && a->Ks, "null cameras");
    B2S_REQUIRE(a->N == 0 || aligned16(a->quats), "quats misaligned");
    const bool overflow = ra->overflow != 0 || a->N > ra->cap_rows;
    const bool sh = ra->color_mode == 0;
    int K = 1;
    if (sh) {
        B2S_REQUIRE(a->sh_coeffs != nullptr || a->N == 0, "SH colour mode needs sh_coeffs");
        B2S_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 3, "sh_degree must be 0..3");
        K = (a->sh_degree + 1) * (a->sh_degree + 1);
        B2S_REQUIRE(a->sh_stride_k >= K, "sh_stride_k < (deg+1)^2");
        B2S_REQUIRE(ra->channels == 3 + (ra->depth_channel ? 1 : 0), "SH colour mode carries 3 (+1 depth) channels");
    } else {
        B2S_REQUIRE(ra->n_color_in >= 0 && ra->n_color_in + (ra->depth_channel ? 1 : 0) == ra->channels, "channel count mismatch");
        B2S_REQUIRE(ra->n_color_in == 0 || ra->colors_in != nullptr || a->N == 0, "null colours");
    }
    RouteDev rp;
    rp.world = ra->world; rp.rank = ra->rank; rp.cams_per_rank = ra->cams_per_rank; rp.region_cams = ra->region_cams;
    rp.ch = ra->channels; rp.color_mode = ra->color_mode; rp.n_color_in = ra->n_color_in; rp.depth_channel = ra->depth_channel;
    rp.overflow = overflow ? 1 : 0;
    rp.cap_rows = ra->cap_rows; rp.n_ctas_cap = ra->cap_rows / FWD_THREADS; rp.colors_in = ra->colors_in;
    for (int d = 0; d <= B2S_MAX_RANKS; ++d) rp.bounds[d] = d <= ra->world ? ra->bounds[d] : 0;
    if (ra->cams_per_rank == 0) {
        B2S_REQUIRE(rp.bounds[0] == 0 && rp.bounds[ra->world] == a->tile_h, "stripe bounds must run from 0 to tile_h");
        for (int d = 0; d < ra->world; ++d) B2S_REQUIRE(rp.bounds[d + 1] >= rp.bounds[d], "stripe bounds must ascend");
    }
    for (int d = 0; d < B2S_MAX_RANKS; ++d) {
        rp.rec[d] = d < ra->world ? reinterpret_cast<float*>(ra->peer_rec[d]) : nullptr;
        rp.cnt[d] = d < ra->world ? reinterpret_cast<int32_t*>(ra->peer_cnt[d]) : nullptr;
        if (d < ra->world) B2S_REQUIRE(rp.rec[d] && rp.cnt[d] && aligned16(rp.cnt[d]), "null / misaligned peer buffer");
    }
    // an empty shard still publishes its (zero) counts; an overflowing one only its tag (CTA 0)
    const unsigned grid = overflow ? 1u : (unsigned)max((int64_t)1, cdiv(a->N, FWD_THREADS));
    const size_t cam_bytes = ((a->C * sizeof(Camera) + 15) / 16) * 16;
    const size_t stage_bytes = (size_t)FWD_THREADS * RT_MAXF * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    alignas(64) CUtensorMap map_a, map_b;
    memset(&map_a, 0, sizeof(map_a)); memset(&map_b, 0, sizeof(map_b));
    if (sh && a->N > 0 && K == 16 && a->sh_stride_k == 16 && aligned16(a->sh_coeffs) && sh_tensor_maps(a->sh_coeffs, a->N, &map_a, &map_b)) {
        const size_t smem = 1024 + TMA_A_BYTES + TMA_B_BYTES + stage_bytes + cam_bytes;
        B2S_CUDA(cudaFuncSetAttribute(project_route_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        project_route_kernel<3><<<grid, FWD_THREADS, smem, st>>>(*a, rp, map_a, map_b);
    } else if (sh && a->N > 0) {
        const size_t smem = (size_t)(FWD_THREADS / 32) * 32 * sh_row_stride(K) * sizeof(float) + 16 + stage_bytes + cam_bytes;
        B2S_CUDA(cudaFuncSetAttribute(project_route_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        project_route_kernel<1><<<grid, FWD_THREADS, smem, st>>>(*a, rp, map_a, map_b);
    } else {
        const size_t smem = stage_bytes + cam_bytes;
        B2S_CUDA(cudaFuncSetAttribute(project_route_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        project_route_kernel<0><<<grid, FWD_THREADS, smem, st>>>(*a, rp, map_a, map_b);
    }
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_route_scan(int32_t world, int32_t region_cams, int64_t cap_rows, const int32_t* cnt, int32_t* row_base,
                              int64_t* info, b2s_stream_t stream) {
    B2S_REQUIRE(world >= 1 && world <= B2S_MAX_RANKS && region_cams > 0 && cap_rows > 0 && cap_rows % FWD_THREADS == 0, "bad sizes");
    B2S_REQUIRE(cnt && row_base && info && aligned16(cnt), "null / misaligned pointer");
    route_scan_kernel<<<world, SC_THREADS, 0, (cudaStream_t)stream>>>(region_cams, cap_rows / FWD_THREADS, cnt, row_base,
                                                                     reinterpret_cast<long long*>(info));
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_route_gather(int32_t world, int32_t region_cams, int64_t cap_rows, int32_t channels, const float* rec,
                                const int32_t* cnt, const int32_t* row_base, const int64_t* host_n_src,
                                const int64_t* host_rows_from, int64_t n_rows,
                                int32_t tile_size, int32_t tile_w, int32_t tile_h, const b2s_tile_rows* rows, float* means2d,
                                float* depths, float* conics, float* opacities, float* colors, int32_t* radii, int64_t* camera_ids,
                                int64_t* gaussian_ids, int32_t* tiles_per_gauss, uint32_t* depth_keys, int32_t* row_ids,
                                b2s_stream_t stream) {
    B2S_REQUIRE(world >= 1 && world <= B2S_MAX_RANKS && region_cams > 0 && cap_rows > 0 && cap_rows % FWD_THREADS == 0, "bad sizes");
    B2S_REQUIRE(channels >= 1 && channels <= RT_MAXCH && host_n_src && host_rows_from && tile_size > 0 && tile_w > 0 && tile_h > 0,
                "bad args");
    if (n_rows == 0) return 0;
    B2S_REQUIRE(rec && cnt && row_base && means2d && depths && conics && opacities && colors && radii && camera_ids && gaussian_ids &&
                tiles_per_gauss && depth_keys && row_ids, "null pointer");
    GatherArgs g;
    g.world = world; g.region_cams = region_cams; g.ch = channels; g.cap_rows = cap_rows; g.n_ctas_cap = cap_rows / FWD_THREADS;
    long long off = 0, items = 0, rows_before = 0;
    for (int s = 0; s < B2S_MAX_RANKS; ++s) {
        const long long ns = s < world ? host_n_src[s] : 0;
        B2S_REQUIRE(ns >= 0 && ns <= cap_rows, "a source's Gaussian count exceeds the region capacity");
        g.n_src[s] = ns; g.gid_offset[s] = off; g.cta_base[s] = items; g.row_offset[s] = rows_before;
        off += ns;
        items += cdiv(ns, FWD_THREADS) * region_cams;
        rows_before += s < world ? host_rows_from[s] : 0;
    }
    B2S_REQUIRE(rows_before == n_rows, "rows per source do not add up to n_rows");
    g.cta_base[B2S_MAX_RANKS] = items;
    for (int s = world; s <= B2S_MAX_RANKS; ++s) g.cta_base[s] = items;
    g.rec = rec; g.cnt = cnt; g.row_base = row_base;
    int lo = 0, hi = tile_h;
    if (rows) { lo = rows->row_lo; hi = rows->row_hi; }
    B2S_REQUIRE(lo >= 0 && hi >= lo && hi <= tile_h, "bad tile-row window");
    g.tile_size = (float)tile_size; g.tile_w = tile_w; g.row_lo = lo; g.row_hi = hi;
    g.means2d = reinterpret_cast<float2*>(means2d); g.depths = depths; g.conics = conics; g.opac = opacities; g.colors = colors;
    g.radii = radii; g.camera_ids = camera_ids; g.gaussian_ids = gaussian_ids; g.tiles = tiles_per_gauss; g.depth_keys = depth_keys;
    g.row_ids = row_ids;
    if (items == 0) return 0;
    const long long avg = n_rows / items;                       // rows per occupied slot, on average
    cudaStream_t st = (cudaStream_t)stream;
    if (avg > 128) route_gather_kernel<8><<<(unsigned)items, 256, 0, st>>>(g);
    else if (avg > 64) route_gather_kernel<4><<<(unsigned)cdiv(items, 2), 256, 0, st>>>(g);
    else if (avg > 32) route_gather_kernel<2><<<(unsigned)cdiv(items, 4), 256, 0, st>>>(g);
    else route_gather_kernel<1><<<(unsigned)cdiv(items, 8), 256, 0, st>>>(g);
    B2S_LAUNCH_CHECK();
    return 0;
}
