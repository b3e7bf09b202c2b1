// Stage 1: fused EWA projection + SH colour (+ tile count), forward and backward.
//
// Replaces gsplat `fully_fused_projection` fwd/bwd + `spherical_harmonics` fwd/bwd + the counting
// half of `isect_tiles`.  Reached in the reference from
// source/container/src/main.py:1308-1316 / :1339-1347 on every training iteration.
//
// Design (HBM-bound streaming kernel):
//  * one thread per Gaussian, cameras looped inside: the 44 B of parameters are read once,
//  * SH coefficients ([N,K,3], 192 B per Gaussian at degree 3 -- the largest tensor of the path) are
//    moved by the whole warp as one contiguous 32-row chunk with 128-bit streaming loads into a
//    padded shared-memory tile (row stride odd => the per-thread row reads are conflict free);
//    the backward writes v_coeffs through the same tile so the 1.1 GB gradient goes out coalesced,
//  * rows are only fetched when some lane of the warp sees its Gaussian in some camera,
//  * the forward also emits tiles_per_gauss so binning needs no extra pass over means2d/radii.
#include "project_common.cuh"

namespace b2s {

constexpr int BWD_THREADS = 128;

// =================================================================================================
// forward
// =================================================================================================
template <int KVEC>
__global__ void __launch_bounds__(FWD_THREADS) project_fwd_kernel(const b2s_project_fwd_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Camera* cams = reinterpret_cast<Camera*>(smem_raw);
    float* sh_tiles = reinterpret_cast<float*>(smem_raw + ((a.C * sizeof(Camera) + 15) / 16) * 16);
    const float W = (float)a.width, H = (float)a.height;
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, W, H);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * FWD_THREADS + threadIdx.x;
    const bool live = n < a.N;
    const int64_t nc = live ? n : a.N - 1;  // clamp so every lane can run the same code

    float mean[3], q[4], s[3];
    mean[0] = __ldg(a.means + 3 * nc); mean[1] = __ldg(a.means + 3 * nc + 1); mean[2] = __ldg(a.means + 3 * nc + 2);
    {
        float4 qq = __ldg(reinterpret_cast<const float4*>(a.quats) + nc);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
    }
    s[0] = __ldg(a.scales + 3 * nc); s[1] = __ldg(a.scales + 3 * nc + 1); s[2] = __ldg(a.scales + 3 * nc + 2);
    const float opac = a.opacities ? __ldg(a.opacities + nc) : 1.f;

    Rot R; float qn[4], qnorm, M[9]; Sym3 S;
    quat_to_rotmat(q, R, qn, qnorm);
    covar_from_rot_scale(R, s, M, S);

    unsigned long long vis = 0ull;
    int my_tiles = 0, my_coarse = 0;
    for (int c = 0; c < a.C; ++c) {
        ProjOut o;
        project_one(cams[c], mean, S, W, H, a.eps2d, a.near_plane, a.far_plane, a.radius_clip, o, nullptr);
        if (!live) continue;
        const int64_t i = (int64_t)c * a.N + n;
        a.radii[i] = o.radius;
        reinterpret_cast<float2*>(a.means2d)[i] = make_float2(o.x2d, o.y2d);
        a.depths[i] = o.depth;
        a.conics[3 * i] = o.conic_a; a.conics[3 * i + 1] = o.conic_b; a.conics[3 * i + 2] = o.conic_c;
        if (a.compensations) a.compensations[i] = o.comp;
        if (a.opacities_out) a.opacities_out[i] = a.antialiased ? opac * o.comp : opac;  // comp == 0 on culled rows
        if (a.tiles_per_gauss) {
            int cnt = 0;
            if (o.radius > 0) {
                int x0, y0, x1, y1;
                tile_rect(o.x2d, o.y2d, o.radius, (float)a.tile_size, a.tile_w, a.tile_h, x0, y0, x1, y1);
                cnt = (x1 - x0) * (y1 - y0);
                if (cnt > 0)
                    my_coarse += (((x1 + B2S_COARSE - 1) / B2S_COARSE) - x0 / B2S_COARSE) *
                                 (((y1 + B2S_COARSE - 1) / B2S_COARSE) - y0 / B2S_COARSE);
            }
            a.tiles_per_gauss[i] = cnt;
            my_tiles += cnt;
        }
        if (a.depth_keys) a.depth_keys[i] = o.radius > 0 ? __float_as_uint(o.depth) : 0xFFFFFFFFu;
        if (a.row_ids) a.row_ids[i] = (int32_t)i;
        if (o.radius > 0) vis |= 1ull << c;
    }

    if (a.total_isects) {
        // The data-dependent sizes the host reads: sum of tile counts, number of visible rows, number of coarse-bin
        // records.  One atomic per CTA and counter, spread over B2S_TOTAL_SLOTS counter rows (same-address atomics
        // serialise in L2: with one per warp they cost 0.3 ms at 5.8 M Gaussians); the host sums the slots.
        __shared__ int s_tot[3][FWD_THREADS / 32];
        int wsum = my_tiles, vcount = __popcll(vis), csum = my_coarse;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            wsum += __shfl_xor_sync(FULL, wsum, d);
            vcount += __shfl_xor_sync(FULL, vcount, d);
            csum += __shfl_xor_sync(FULL, csum, d);
        }
        if (lane == 0) { s_tot[0][warp] = wsum; s_tot[1][warp] = vcount; s_tot[2][warp] = csum; }
        __syncthreads();
        if (threadIdx.x < 3) {
            int tsum = 0;
#pragma unroll
            for (int w = 0; w < FWD_THREADS / 32; ++w) tsum += s_tot[threadIdx.x][w];
            if (tsum) atomicAdd(reinterpret_cast<unsigned long long*>(a.total_isects) + 4 * (blockIdx.x % B2S_TOTAL_SLOTS) + threadIdx.x,
                                (unsigned long long)tsum);
        }
    }

    if (a.sh_coeffs == nullptr || a.colors == nullptr) return;
    const int K = (a.sh_degree + 1) * (a.sh_degree + 1);
    const int rs = sh_row_stride(K);
    float* tile = sh_tiles + (size_t)warp * 32 * rs;
    const unsigned row_mask = __ballot_sync(FULL, live && vis != 0ull);
    const int64_t row0 = (int64_t)blockIdx.x * FWD_THREADS + warp * 32;
    if (row_mask) {
        int rows = (int)min((int64_t)32, a.N - row0);
        sh_rows_load<KVEC>(tile, a.sh_coeffs + (size_t)row0 * a.sh_stride_k * 3, rows, row_mask, K, a.sh_stride_k, lane);
    }
    __syncwarp();
    if (!live) return;
    const float* row = tile + lane * rs;
    for (int c = 0; c < a.C; ++c) {
        const int64_t i = (int64_t)c * a.N + n;
        float r = 0.f, g = 0.f, b = 0.f;
        if ((vis >> c) & 1ull) {
            float dx = mean[0] - cams[c].campos[0], dy = mean[1] - cams[c].campos[1], dz = mean[2] - cams[c].campos[2];
            float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
            float Y[16];
            sh_basis(a.sh_degree, dx * inv, dy * inv, dz * inv, Y);
            if (KVEC == 16) {                      // 4 bases = 12 floats = three 128-bit shared loads
                const float4* row4 = reinterpret_cast<const float4*>(row);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 c0 = row4[3 * q], c1 = row4[3 * q + 1], c2 = row4[3 * q + 2];
                    r += Y[4 * q] * c0.x; g += Y[4 * q] * c0.y; b += Y[4 * q] * c0.z;
                    r += Y[4 * q + 1] * c0.w; g += Y[4 * q + 1] * c1.x; b += Y[4 * q + 1] * c1.y;
                    r += Y[4 * q + 2] * c1.z; g += Y[4 * q + 2] * c1.w; b += Y[4 * q + 2] * c2.x;
                    r += Y[4 * q + 3] * c2.y; g += Y[4 * q + 3] * c2.z; b += Y[4 * q + 3] * c2.w;
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    r += Y[k] * row[3 * k]; g += Y[k] * row[3 * k + 1]; b += Y[k] * row[3 * k + 2];
                }
            }
            r = fmaxf(r + 0.5f, 0.f); g = fmaxf(g + 0.5f, 0.f); b = fmaxf(b + 0.5f, 0.f);
        }
        a.colors[3 * i] = r; a.colors[3 * i + 1] = g; a.colors[3 * i + 2] = b;
    }
}

// =================================================================================================
// forward, TMA variant (degree-3 coefficient tensors, [N,16,3] contiguous): the CTA's 256 coefficient rows (48 KB) are
// fetched by the tensor-memory accelerator as two 2-D boxes -- floats [0,32) of every row with the 128-byte swizzle,
// floats [32,48) with the 64-byte swizzle -- issued by ONE thread at kernel start, so the 1.1 GB coefficient stream is
// in flight while all 256 threads do the projection arithmetic, and lands in shared memory without passing through
// registers (the LDG.128 + STS.128 round trip of the kernel above).  The hardware XOR-swizzle (16-byte chunk index ^ row
// bits) makes the per-thread row reads conflict free on dense, unpadded rows.  Completion is an mbarrier transaction count.
// SASS: UTMALDG + SYNCS.  Rows beyond N are zero-filled by the accelerator (out-of-bounds fill), so the tail needs no mask.
// =================================================================================================
__global__ void __launch_bounds__(FWD_THREADS) project_fwd_tma_kernel(const b2s_project_fwd_args a,
                                                                      const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_b) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;
    // 1024-byte aligned tile base (the 128-byte swizzle pattern repeats every 1 KB of shared-memory ADDRESS)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
    unsigned char* tile_a = smem_raw + pad;
    unsigned char* tile_b = tile_a + TMA_A_BYTES;
    Camera* cams = reinterpret_cast<Camera*>(tile_b + TMA_B_BYTES);
    const float W = (float)a.width, H = (float)a.height;
    const uint32_t mbar = smem_u32(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, W, H);      // ends with __syncthreads: the barrier is initialised for all
    const int64_t row0 = (int64_t)blockIdx.x * FWD_THREADS;
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar, TMA_A_BYTES + TMA_B_BYTES);
        tma_load_2d(smem_u32(tile_a), &map_a, 0, (int)row0, mbar);
        tma_load_2d(smem_u32(tile_b), &map_b, 32, (int)row0, mbar);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = row0 + threadIdx.x;
    const bool live = n < a.N;
    const int64_t nc = live ? n : a.N - 1;  // clamp so every lane can run the same code

    float mean[3], q[4], s[3];
    mean[0] = __ldg(a.means + 3 * nc); mean[1] = __ldg(a.means + 3 * nc + 1); mean[2] = __ldg(a.means + 3 * nc + 2);
    {
        float4 qq = __ldg(reinterpret_cast<const float4*>(a.quats) + nc);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
    }
    s[0] = __ldg(a.scales + 3 * nc); s[1] = __ldg(a.scales + 3 * nc + 1); s[2] = __ldg(a.scales + 3 * nc + 2);
    const float opac = a.opacities ? __ldg(a.opacities + nc) : 1.f;

    Rot R; float qn[4], qnorm, M[9]; Sym3 S;
    quat_to_rotmat(q, R, qn, qnorm);
    covar_from_rot_scale(R, s, M, S);

    unsigned long long vis = 0ull;
    int my_tiles = 0, my_coarse = 0;
    for (int c = 0; c < a.C; ++c) {
        ProjOut o;
        project_one(cams[c], mean, S, W, H, a.eps2d, a.near_plane, a.far_plane, a.radius_clip, o, nullptr);
        if (!live) continue;
        const int64_t i = (int64_t)c * a.N + n;
        a.radii[i] = o.radius;
        reinterpret_cast<float2*>(a.means2d)[i] = make_float2(o.x2d, o.y2d);
        a.depths[i] = o.depth;
        a.conics[3 * i] = o.conic_a; a.conics[3 * i + 1] = o.conic_b; a.conics[3 * i + 2] = o.conic_c;
        if (a.compensations) a.compensations[i] = o.comp;
        if (a.opacities_out) a.opacities_out[i] = a.antialiased ? opac * o.comp : opac;  // comp == 0 on culled rows
        if (a.tiles_per_gauss) {
            int cnt = 0;
            if (o.radius > 0) {
                int x0, y0, x1, y1;
                tile_rect(o.x2d, o.y2d, o.radius, (float)a.tile_size, a.tile_w, a.tile_h, x0, y0, x1, y1);
                cnt = (x1 - x0) * (y1 - y0);
                if (cnt > 0)
                    my_coarse += (((x1 + B2S_COARSE - 1) / B2S_COARSE) - x0 / B2S_COARSE) *
                                 (((y1 + B2S_COARSE - 1) / B2S_COARSE) - y0 / B2S_COARSE);
            }
            a.tiles_per_gauss[i] = cnt;
            my_tiles += cnt;
        }
        if (a.depth_keys) a.depth_keys[i] = o.radius > 0 ? __float_as_uint(o.depth) : 0xFFFFFFFFu;
        if (a.row_ids) a.row_ids[i] = (int32_t)i;
        if (o.radius > 0) vis |= 1ull << c;
    }

    if (a.total_isects) {      // (see project_fwd_kernel)
        __shared__ int s_tot[3][FWD_THREADS / 32];
        int wsum = my_tiles, vcount = __popcll(vis), csum = my_coarse;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            wsum += __shfl_xor_sync(FULL, wsum, d);
            vcount += __shfl_xor_sync(FULL, vcount, d);
            csum += __shfl_xor_sync(FULL, csum, d);
        }
        if (lane == 0) { s_tot[0][warp] = wsum; s_tot[1][warp] = vcount; s_tot[2][warp] = csum; }
        __syncthreads();
        if (threadIdx.x < 3) {
            int tsum = 0;
#pragma unroll
            for (int w = 0; w < FWD_THREADS / 32; ++w) tsum += s_tot[threadIdx.x][w];
            if (tsum) atomicAdd(reinterpret_cast<unsigned long long*>(a.total_isects) + 4 * (blockIdx.x % B2S_TOTAL_SLOTS) + threadIdx.x,
                                (unsigned long long)tsum);
        }
    }

    mbar_wait(mbar, 0);            // the 48 KB of coefficient rows have landed (every thread observes the phase flip itself)
    if (!live) return;
    // row r = threadIdx.x; 16-byte chunk j of the row: j < 8 in box A at chunk (j ^ (r & 7)), j >= 8 in box B at ((j - 8) ^ ((r >> 1) & 3))
    const unsigned r = threadIdx.x;
    const float4* row_a = reinterpret_cast<const float4*>(tile_a + r * 128);
    const float4* row_b = reinterpret_cast<const float4*>(tile_b + r * 64);
    const unsigned xa = r & 7u, xb = (r >> 1) & 3u;
    for (int c = 0; c < a.C; ++c) {
        const int64_t i = (int64_t)c * a.N + n;
        float cr = 0.f, cg = 0.f, cb = 0.f;
        if ((vis >> c) & 1ull) {
            float dx = mean[0] - cams[c].campos[0], dy = mean[1] - cams[c].campos[1], dz = mean[2] - cams[c].campos[2];
            float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
            float Y[16];
            sh_basis(3, dx * inv, dy * inv, dz * inv, Y);
#pragma unroll
            for (int g = 0; g < 4; ++g) {          // 4 bases = 12 floats = three 16-byte chunks
                float4 ch[3];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int j = 3 * g + t;
                    ch[t] = j < 8 ? row_a[(unsigned)j ^ xa] : row_b[(unsigned)(j - 8) ^ xb];
                }
                const float4 c0 = ch[0], c1 = ch[1], c2 = ch[2];
                cr += Y[4 * g] * c0.x; cg += Y[4 * g] * c0.y; cb += Y[4 * g] * c0.z;
                cr += Y[4 * g + 1] * c0.w; cg += Y[4 * g + 1] * c1.x; cb += Y[4 * g + 1] * c1.y;
                cr += Y[4 * g + 2] * c1.z; cg += Y[4 * g + 2] * c1.w; cb += Y[4 * g + 2] * c2.x;
                cr += Y[4 * g + 3] * c2.y; cg += Y[4 * g + 3] * c2.z; cb += Y[4 * g + 3] * c2.w;
            }
            cr = fmaxf(cr + 0.5f, 0.f); cg = fmaxf(cg + 0.5f, 0.f); cb = fmaxf(cb + 0.5f, 0.f);
        }
        a.colors[3 * i] = cr; a.colors[3 * i + 1] = cg; a.colors[3 * i + 2] = cb;
    }
}

// =================================================================================================
// backward, part 1: SH colour VJP (streams 12K B/Gaussian in and out -- 85 % of the backward's bytes --
// with few registers, so it runs at high occupancy), writes v_sh_coeffs and the view-direction part of
// v_means.  MULTI: more than one camera => separate in/out tiles; one camera => in place.
// =================================================================================================
constexpr int SHB_THREADS = 256;

// sparse backward mode (see bwd_classify_kernel below): list lengths, written by the classify pass of the same call
struct BwdCounts {
    unsigned n_geom, n_sh;
    unsigned pad[62];
};
__device__ __forceinline__ bool bwd_sparse(unsigned n_live, int64_t N) { return (int64_t)n_live * 4 <= N; }

template <int KVEC, bool MULTI>
__global__ void __launch_bounds__(SHB_THREADS, 4) sh_bwd_fused_kernel(const b2s_project_bwd_args a, const BwdCounts* cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Camera* cams = reinterpret_cast<Camera*>(smem_raw);
    if (cnt && bwd_sparse(cnt->n_sh, a.N)) return;          // sh_zero_kernel + sh_sparse_kernel do this call's work
    float* sh_tiles = reinterpret_cast<float*>(smem_raw + ((a.C * sizeof(Camera) + 15) / 16) * 16);
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, (float)a.width, (float)a.height);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // persistent: a fixed grid walks the 256-row blocks (in sparse mode the whole grid returns above after a handful of
    // CTA launches instead of N / 256 of them)
    const int64_t n_blocks = (a.N + SHB_THREADS - 1) / SHB_THREADS;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const int64_t n = blk * SHB_THREADS + threadIdx.x;
    const bool live = n < a.N;
    const int64_t nc = live ? n : a.N - 1;
    float mean[3];
    mean[0] = __ldg(a.means + 3 * nc); mean[1] = __ldg(a.means + 3 * nc + 1); mean[2] = __ldg(a.means + 3 * nc + 2);
    unsigned long long vis = 0ull;
    if (live)
        for (int c = 0; c < a.C; ++c) {
            const int64_t i = (int64_t)c * a.N + n;
            if (__ldg(a.radii + i) <= 0) continue;
            if (a.v_colors) {   // a zero colour cotangent (occluded row) contributes exactly nothing: skip its 12K bytes
                const float v0 = __ldg(a.v_colors + 3 * i), v1 = __ldg(a.v_colors + 3 * i + 1), v2 = __ldg(a.v_colors + 3 * i + 2);
                if (v0 == 0.f && v1 == 0.f && v2 == 0.f) continue;
            }
            vis |= 1ull << c;
        }

    const int K = (a.sh_degree + 1) * (a.sh_degree + 1);
    const int rs = sh_row_stride(K);
    float* tile_in = sh_tiles + (size_t)warp * (MULTI ? 2 : 1) * 32 * rs;
    float* tile_out = MULTI ? tile_in + 32 * rs : tile_in;
    const int64_t row0 = blk * SHB_THREADS + warp * 32;
    const int rows = (int)max((int64_t)0, min((int64_t)32, a.N - row0));
    const unsigned row_mask = a.v_colors ? __ballot_sync(FULL, live && vis != 0ull) : 0u;
    if (row_mask == 0u && !a.accumulate) {
        // nothing in this warp's 32 rows can have a gradient (culled / occluded): stream zeros straight out
        if (live) { a.v_means[3 * n] = 0.f; a.v_means[3 * n + 1] = 0.f; a.v_means[3 * n + 2] = 0.f; }
        if (rows > 0) {
            float* g = a.v_sh_coeffs + (size_t)row0 * a.sh_stride_k * 3;
            if (KVEC == 16) {
                float4* g4 = reinterpret_cast<float4*>(g);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    const int q = lane + 32 * i;
                    if (q < rows * 12) stg_stream4(g4 + q, z);
                }
            } else {
                const int per = 3 * K;
                for (int f = lane; f < 32 * per; f += 32) {
                    const int r = f / per;
                    if (r < rows) g[(size_t)r * a.sh_stride_k * 3 + (f - r * per)] = 0.f;
                }
            }
        }
        continue;
    }
    // masked-off rows are zero-filled by the loader, so culled Gaussians get zero gradients for free
    sh_rows_load<KVEC>(tile_in, a.sh_coeffs + (size_t)row0 * a.sh_stride_k * 3, rows, row_mask, K, a.sh_stride_k, lane);
    if (MULTI)
        for (int f = lane; f < 32 * rs; f += 32) tile_out[f] = 0.f;
    __syncwarp();

    float v_mean[3] = {0.f, 0.f, 0.f};
    if (live && vis != 0ull && a.v_colors) {
        const float* rin = tile_in + lane * rs;
        float* rout = tile_out + lane * rs;
        for (int c = 0; c < a.C; ++c) {
            if (!((vis >> c) & 1ull)) continue;
            const int64_t i = (int64_t)c * a.N + n;
            float dx = mean[0] - cams[c].campos[0], dy = mean[1] - cams[c].campos[1], dz = mean[2] - cams[c].campos[2];
            float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
            float ux = dx * inv, uy = dy * inv, uz = dz * inv;
            float Y[16], w[16];
            sh_basis(a.sh_degree, ux, uy, uz, Y);
            float cr = 0.5f, cg = 0.5f, cb = 0.5f;
            if (KVEC == 16) {
                const float4* rin4 = reinterpret_cast<const float4*>(rin);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 c0 = rin4[3 * q], c1 = rin4[3 * q + 1], c2 = rin4[3 * q + 2];
                    cr += Y[4 * q] * c0.x; cg += Y[4 * q] * c0.y; cb += Y[4 * q] * c0.z;
                    cr += Y[4 * q + 1] * c0.w; cg += Y[4 * q + 1] * c1.x; cb += Y[4 * q + 1] * c1.y;
                    cr += Y[4 * q + 2] * c1.z; cg += Y[4 * q + 2] * c1.w; cb += Y[4 * q + 2] * c2.x;
                    cr += Y[4 * q + 3] * c2.y; cg += Y[4 * q + 3] * c2.z; cb += Y[4 * q + 3] * c2.w;
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < K; ++k) { cr += Y[k] * rin[3 * k]; cg += Y[k] * rin[3 * k + 1]; cb += Y[k] * rin[3 * k + 2]; }
            }
            const float vr = cr > 0.f ? __ldg(a.v_colors + 3 * i) : 0.f;        // clamp_min(., 0) gate
            const float vg = cg > 0.f ? __ldg(a.v_colors + 3 * i + 1) : 0.f;
            const float vb = cb > 0.f ? __ldg(a.v_colors + 3 * i + 2) : 0.f;
            if (KVEC == 16) {
                const float4* rin4 = reinterpret_cast<const float4*>(rin);
                float4* rout4 = reinterpret_cast<float4*>(rout);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 c0 = rin4[3 * q], c1 = rin4[3 * q + 1], c2 = rin4[3 * q + 2];
                    w[4 * q] = c0.x * vr + c0.y * vg + c0.z * vb;
                    w[4 * q + 1] = c0.w * vr + c1.x * vg + c1.y * vb;
                    w[4 * q + 2] = c1.z * vr + c1.w * vg + c2.x * vb;
                    w[4 * q + 3] = c2.y * vr + c2.z * vg + c2.w * vb;
                    float4 o0 = make_float4(Y[4 * q] * vr, Y[4 * q] * vg, Y[4 * q] * vb, Y[4 * q + 1] * vr);
                    float4 o1 = make_float4(Y[4 * q + 1] * vg, Y[4 * q + 1] * vb, Y[4 * q + 2] * vr, Y[4 * q + 2] * vg);
                    float4 o2 = make_float4(Y[4 * q + 2] * vb, Y[4 * q + 3] * vr, Y[4 * q + 3] * vg, Y[4 * q + 3] * vb);
                    if (MULTI) {
                        const float4 p0 = rout4[3 * q], p1 = rout4[3 * q + 1], p2 = rout4[3 * q + 2];
                        o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
                        o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
                        o2.x += p2.x; o2.y += p2.y; o2.z += p2.z; o2.w += p2.w;
                    }
                    rout4[3 * q] = o0; rout4[3 * q + 1] = o1; rout4[3 * q + 2] = o2;
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    w[k] = rin[3 * k] * vr + rin[3 * k + 1] * vg + rin[3 * k + 2] * vb;
                    if (MULTI) { rout[3 * k] += Y[k] * vr; rout[3 * k + 1] += Y[k] * vg; rout[3 * k + 2] += Y[k] * vb; }
                    else { rout[3 * k] = Y[k] * vr; rout[3 * k + 1] = Y[k] * vg; rout[3 * k + 2] = Y[k] * vb; }
                }
            }
            float vx, vy, vz;
            sh_basis_vjp(a.sh_degree, ux, uy, uz, w, vx, vy, vz);
            const float dp = ux * vx + uy * vy + uz * vz;
            v_mean[0] += (vx - ux * dp) * inv; v_mean[1] += (vy - uy * dp) * inv; v_mean[2] += (vz - uz * dp) * inv;
        }
    }
    if (live) {
        if (a.accumulate) { a.v_means[3 * n] += v_mean[0]; a.v_means[3 * n + 1] += v_mean[1]; a.v_means[3 * n + 2] += v_mean[2]; }
        else { a.v_means[3 * n] = v_mean[0]; a.v_means[3 * n + 1] = v_mean[1]; a.v_means[3 * n + 2] = v_mean[2]; }
    }
    __syncwarp();
    if (rows > 0)
        sh_rows_store<KVEC>(tile_out, a.v_sh_coeffs + (size_t)row0 * a.sh_stride_k * 3, rows, K, a.sh_stride_k, lane,
                            a.accumulate != 0);
    __syncwarp();      // the next block reuses this warp's tile
    }
}

// =================================================================================================
// backward, sparse mode.  In a dense scene compositing saturates after the first few layers, so most visible rows
// come back with an all-zero cotangent (2.7 % are reached in the 5.8 M benchmark scene).  `bwd_classify_kernel`
// zero-fills the small gradients, tests every row once and appends the live ones to two lists (geometry, SH colour);
// when a list holds at most a quarter of the rows the dense kernel of that stage returns immediately and a
// list-driven kernel does the work instead (for SH: a streaming zero fill of v_sh plus one thread per live row).
// The choice is made on the device from the list lengths: no host round trip.
// =================================================================================================

constexpr int CLS_ROWS = 4;      // rows per thread: 1024 rows per CTA, ONE global atomic per CTA and list (same-address
                                 // atomics serialise in L2: one per warp cost 0.18 ms at 5.8 M rows)
__global__ void __launch_bounds__(256) bwd_classify_kernel(const b2s_project_bwd_args a, bool with_sh, BwdCounts* cnt,
                                                           int32_t* __restrict__ list_geom, int32_t* __restrict__ list_sh) {
    __shared__ unsigned s_g[CLS_ROWS * 8 + 1], s_s[CLS_ROWS * 8 + 1];      // per (k, warp) counts -> exclusive prefix
    __shared__ unsigned s_base[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned bg[CLS_ROWS], bs[CLS_ROWS];
    bool lg[CLS_ROWS], ls[CLS_ROWS];
#pragma unroll
    for (int k = 0; k < CLS_ROWS; ++k) {
        const int64_t n = ((int64_t)blockIdx.x * CLS_ROWS + k) * 256 + threadIdx.x;
        bool live_g = false, live_s = false;
        if (n < a.N) {
            for (int c = 0; c < a.C; ++c) {
                const int64_t i = (int64_t)c * a.N + n;
                // every load is issued before the radius is looked at: behind a `radius <= 0 -> continue` the cotangent loads
                // wait one memory latency for the radius and the kernel sat 67 % of its stall samples on the scoreboard
                // (4.5 TB/s); 8 % of the rows are culled, so the loads that turn out unnecessary are few
                const int r = __ldg(a.radii + i);
#ifdef B2S_CLASSIFY_DEPENDENT      // A/B: the old order
                if (r <= 0) continue;
#endif
                bool g = false, sc = false;
                if (a.v_means2d) { const float2 v = __ldg(reinterpret_cast<const float2*>(a.v_means2d) + i); g |= (v.x != 0.f) | (v.y != 0.f); }
                if (a.v_conics) g |= (__ldg(a.v_conics + 3 * i) != 0.f) | (__ldg(a.v_conics + 3 * i + 1) != 0.f) | (__ldg(a.v_conics + 3 * i + 2) != 0.f);
                if (a.v_opacities_out) g |= __ldg(a.v_opacities_out + i) != 0.f;
                if (a.v_depths) g |= __ldg(a.v_depths + i) != 0.f;
                if (a.v_compensations) g |= __ldg(a.v_compensations + i) != 0.f;
                if (with_sh && a.v_colors)
                    sc = (__ldg(a.v_colors + 3 * i) != 0.f) | (__ldg(a.v_colors + 3 * i + 1) != 0.f) | (__ldg(a.v_colors + 3 * i + 2) != 0.f);
                live_g |= g & (r > 0);
                live_s |= sc & (r > 0);
            }
            a.v_means[3 * n] = 0.f; a.v_means[3 * n + 1] = 0.f; a.v_means[3 * n + 2] = 0.f;
            if (a.v_scales) { a.v_scales[3 * n] = 0.f; a.v_scales[3 * n + 1] = 0.f; a.v_scales[3 * n + 2] = 0.f; }
            if (a.v_quats) reinterpret_cast<float4*>(a.v_quats)[n] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.v_opacities) a.v_opacities[n] = 0.f;
        }
        lg[k] = live_g; ls[k] = live_s;
        bg[k] = __ballot_sync(FULL, live_g); bs[k] = __ballot_sync(FULL, live_s);
        if (lane == 0) { s_g[k * 8 + warp] = __popc(bg[k]); s_s[k * 8 + warp] = __popc(bs[k]); }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned* c = threadIdx.x == 0 ? s_g : s_s;
        unsigned run = 0;
        for (int j = 0; j < CLS_ROWS * 8; ++j) { const unsigned v = c[j]; c[j] = run; run += v; }
        s_base[threadIdx.x] = run ? atomicAdd(threadIdx.x == 0 ? &cnt->n_geom : &cnt->n_sh, run) : 0u;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < CLS_ROWS; ++k) {
        const int64_t n = ((int64_t)blockIdx.x * CLS_ROWS + k) * 256 + threadIdx.x;
        if (lg[k]) list_geom[s_base[0] + s_g[k * 8 + warp] + __popc(bg[k] & lt)] = (int32_t)n;
        if (ls[k]) list_sh[s_base[1] + s_s[k * 8 + warp] + __popc(bs[k] & lt)] = (int32_t)n;
    }
}

__global__ void __launch_bounds__(256) sh_zero_kernel(const b2s_project_bwd_args a, const BwdCounts* cnt) {
    if (!bwd_sparse(cnt->n_sh, a.N)) return;
    const int64_t total = a.N * a.sh_stride_k * 3;            // floats; the tensor is 16-byte aligned and a multiple of 4 here
    float4* o = reinterpret_cast<float4*>(a.v_sh_coeffs);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += (int64_t)gridDim.x * blockDim.x)
        stg_stream4(o + i, z);
    if (blockIdx.x == 0 && threadIdx.x < (total & 3)) a.v_sh_coeffs[(total & ~(int64_t)3) + threadIdx.x] = 0.f;
}

// one thread per live row: SH colour VJP straight from / to global memory (few rows: no staging needed).
// VEC: degree 3, 16-byte aligned rows of 48 floats -> twelve 128-bit loads and stores per row.
template <bool VEC>
__global__ void __launch_bounds__(128) sh_sparse_kernel(const b2s_project_bwd_args a, const BwdCounts* cnt,
                                                        const int32_t* __restrict__ list_sh) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Camera* cams = reinterpret_cast<Camera*>(smem_raw);
    if (!bwd_sparse(cnt->n_sh, a.N)) return;
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, (float)a.width, (float)a.height);
    const int K = (a.sh_degree + 1) * (a.sh_degree + 1);
    const int64_t total = cnt->n_sh;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = list_sh[idx];
        const float* cf = a.sh_coeffs + n * a.sh_stride_k * 3;
        float* out = a.v_sh_coeffs + n * a.sh_stride_k * 3;
        float co[48], acc[48];
        if (VEC) {
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const float4 v = ldg_stream4(reinterpret_cast<const float4*>(cf) + q);
                co[4 * q] = v.x; co[4 * q + 1] = v.y; co[4 * q + 2] = v.z; co[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int e = 0; e < 48; ++e) acc[e] = 0.f;
        }
        const float mx = __ldg(a.means + 3 * n), my = __ldg(a.means + 3 * n + 1), mz = __ldg(a.means + 3 * n + 2);
        float v_mean[3] = {0.f, 0.f, 0.f};
        bool first = true;
        for (int c = 0; c < a.C; ++c) {
            const int64_t i = (int64_t)c * a.N + n;
            if (__ldg(a.radii + i) <= 0) continue;
            const float g0 = __ldg(a.v_colors + 3 * i), g1 = __ldg(a.v_colors + 3 * i + 1), g2 = __ldg(a.v_colors + 3 * i + 2);
            if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
            const float dx = mx - cams[c].campos[0], dy = my - cams[c].campos[1], dz = mz - cams[c].campos[2];
            const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
            const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
            float Y[16], w[16];
            sh_basis(a.sh_degree, ux, uy, uz, Y);
            float cr = 0.5f, cg = 0.5f, cb = 0.5f;
            if (VEC) {
#pragma unroll
                for (int k = 0; k < 16; ++k) { cr += Y[k] * co[3 * k]; cg += Y[k] * co[3 * k + 1]; cb += Y[k] * co[3 * k + 2]; }
            } else {
                for (int k = 0; k < K; ++k) { cr += Y[k] * __ldg(cf + 3 * k); cg += Y[k] * __ldg(cf + 3 * k + 1); cb += Y[k] * __ldg(cf + 3 * k + 2); }
            }
            const float vr = cr > 0.f ? g0 : 0.f, vg = cg > 0.f ? g1 : 0.f, vb = cb > 0.f ? g2 : 0.f;   // clamp_min(., 0) gate
            if (VEC) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    w[k] = co[3 * k] * vr + co[3 * k + 1] * vg + co[3 * k + 2] * vb;
                    acc[3 * k] += Y[k] * vr; acc[3 * k + 1] += Y[k] * vg; acc[3 * k + 2] += Y[k] * vb;
                }
            } else {
                for (int k = 0; k < K; ++k) {
                    w[k] = __ldg(cf + 3 * k) * vr + __ldg(cf + 3 * k + 1) * vg + __ldg(cf + 3 * k + 2) * vb;
                    if (first) { out[3 * k] = Y[k] * vr; out[3 * k + 1] = Y[k] * vg; out[3 * k + 2] = Y[k] * vb; }
                    else { out[3 * k] += Y[k] * vr; out[3 * k + 1] += Y[k] * vg; out[3 * k + 2] += Y[k] * vb; }
                }
            }
            first = false;
            float vx, vy, vz;
            sh_basis_vjp(a.sh_degree, ux, uy, uz, w, vx, vy, vz);
            const float dp = ux * vx + uy * vy + uz * vz;
            v_mean[0] += (vx - ux * dp) * inv; v_mean[1] += (vy - uy * dp) * inv; v_mean[2] += (vz - uz * dp) * inv;
        }
        if (VEC) {
#pragma unroll
            for (int q = 0; q < 12; ++q)
                stg_stream4(reinterpret_cast<float4*>(out) + q, make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
        }
        a.v_means[3 * n] += v_mean[0]; a.v_means[3 * n + 1] += v_mean[1]; a.v_means[3 * n + 2] += v_mean[2];   // zero-filled by classify
    }
}

// projection VJP of one Gaussian over the call's cameras (shared by the dense and the list-driven kernel)
__device__ __forceinline__ bool project_bwd_row(const b2s_project_bwd_args& a, const Camera* cams, int64_t n, float (&v_mean)[3],
                                                float (&v_quat)[4], float (&v_scale)[3], float& v_opac) {
    float mean[3], q[4], s[3];
    mean[0] = __ldg(a.means + 3 * n); mean[1] = __ldg(a.means + 3 * n + 1); mean[2] = __ldg(a.means + 3 * n + 2);
    {
        float4 qq = __ldg(reinterpret_cast<const float4*>(a.quats) + n);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
    }
    s[0] = __ldg(a.scales + 3 * n); s[1] = __ldg(a.scales + 3 * n + 1); s[2] = __ldg(a.scales + 3 * n + 2);
    const float opac = a.opacities ? __ldg(a.opacities + n) : 1.f;

    Rot R; float qn[4], qnorm, M[9]; Sym3 S;
    quat_to_rotmat(q, R, qn, qnorm);
    covar_from_rot_scale(R, s, M, S);

    v_mean[0] = v_mean[1] = v_mean[2] = 0.f;
    Sym3 v_S = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    v_opac = 0.f;
    bool any = false;
    for (int c = 0; c < a.C; ++c) {
        const int64_t i = (int64_t)c * a.N + n;
        if (__ldg(a.radii + i) <= 0) continue;   // the forward's decision is authoritative
        ProjMid mid;
        project_mid_fast(cams[c], mean, S, a.eps2d, mid);
        if (!(mid.det > 0.f)) continue;
        any = true;
        ProjGradIn g;
        if (a.v_means2d) { float2 v = __ldg(reinterpret_cast<const float2*>(a.v_means2d) + i); g.v_x2d = v.x; g.v_y2d = v.y; }
        else { g.v_x2d = g.v_y2d = 0.f; }
        g.v_depth = a.v_depths ? __ldg(a.v_depths + i) : 0.f;
        if (a.v_conics) { g.v_conic_a = __ldg(a.v_conics + 3 * i); g.v_conic_b = __ldg(a.v_conics + 3 * i + 1); g.v_conic_c = __ldg(a.v_conics + 3 * i + 2); }
        else { g.v_conic_a = g.v_conic_b = g.v_conic_c = 0.f; }
        g.v_comp = a.v_compensations ? __ldg(a.v_compensations + i) : 0.f;
        if (a.v_opacities_out) {
            float vo = __ldg(a.v_opacities_out + i);
            if (a.antialiased) {
                g.v_comp += vo * opac;
                v_opac += vo * sqrtf(fmaxf(mid.det_orig / mid.det, 0.f));
            }
            else v_opac += vo;
        }
        project_one_bwd(cams[c], mean, S, a.eps2d, mid, g, v_mean, v_S);
    }
    v_quat[0] = v_quat[1] = v_quat[2] = v_quat[3] = 0.f;
    v_scale[0] = v_scale[1] = v_scale[2] = 0.f;
    if (any) covar_bwd(q, qn, qnorm, R, s, M, v_S, v_quat, v_scale);
    return any;
}

// =================================================================================================
// backward, part 2: projection VJP (little traffic, heavy arithmetic)
// =================================================================================================
template <bool MEANS_ACC>   // MEANS_ACC: v_means already holds the SH part (or an earlier camera chunk)
__global__ void __launch_bounds__(BWD_THREADS) project_bwd_kernel(const b2s_project_bwd_args a, const BwdCounts* cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Camera* cams = reinterpret_cast<Camera*>(smem_raw);
    if (cnt && bwd_sparse(cnt->n_geom, a.N)) return;        // the list-driven kernel does this call's work
    const float W = (float)a.width, H = (float)a.height;
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, W, H);

    const int64_t n = (int64_t)blockIdx.x * BWD_THREADS + threadIdx.x;
    if (n >= a.N) return;
    if (!a.accumulate) {
        // Rows that are culled, or whose cotangents are all exactly zero (occluded: compositing never reached them),
        // have zero gradients: write the zeros and skip the parameter loads and the arithmetic.
        bool live = false;
        for (int c = 0; c < a.C && !live; ++c) {
            const int64_t i = (int64_t)c * a.N + n;
            if (__ldg(a.radii + i) <= 0) continue;
            if (a.v_means2d) { const float2 v = __ldg(reinterpret_cast<const float2*>(a.v_means2d) + i); live |= (v.x != 0.f) | (v.y != 0.f); }
            if (a.v_conics) live |= (__ldg(a.v_conics + 3 * i) != 0.f) | (__ldg(a.v_conics + 3 * i + 1) != 0.f) | (__ldg(a.v_conics + 3 * i + 2) != 0.f);
            if (a.v_opacities_out) live |= __ldg(a.v_opacities_out + i) != 0.f;
            if (a.v_depths) live |= __ldg(a.v_depths + i) != 0.f;
            if (a.v_compensations) live |= __ldg(a.v_compensations + i) != 0.f;
        }
        if (!live) {
            for (int k = 0; k < 3; ++k) { if (!MEANS_ACC) a.v_means[3 * n + k] = 0.f; a.v_scales[3 * n + k] = 0.f; }
            reinterpret_cast<float4*>(a.v_quats)[n] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.v_opacities) a.v_opacities[n] = 0.f;
            return;
        }
    }
    float v_mean[3], v_quat[4], v_scale[3], v_opac;
    const bool any = project_bwd_row(a, cams, n, v_mean, v_quat, v_scale, v_opac);
    if (a.accumulate) {
        for (int k = 0; k < 3; ++k) { a.v_means[3 * n + k] += v_mean[k]; a.v_scales[3 * n + k] += v_scale[k]; }
        for (int k = 0; k < 4; ++k) a.v_quats[4 * n + k] += v_quat[k];
        if (a.v_opacities) a.v_opacities[n] += v_opac;
    } else {
        for (int k = 0; k < 3; ++k) {
            if (MEANS_ACC) { if (any) a.v_means[3 * n + k] += v_mean[k]; }
            else a.v_means[3 * n + k] = v_mean[k];
            a.v_scales[3 * n + k] = v_scale[k];
        }
        reinterpret_cast<float4*>(a.v_quats)[n] = make_float4(v_quat[0], v_quat[1], v_quat[2], v_quat[3]);
        if (a.v_opacities) a.v_opacities[n] = v_opac;
    }
}


// list-driven projection VJP: one thread per live row (the rest was zero-filled by bwd_classify_kernel).  Used whenever
// the workspace is present, sparse or not: the list keeps 1024-row blocks in ascending order, so loads stay coalesced.
__global__ void __launch_bounds__(BWD_THREADS) project_bwd_sparse_kernel(const b2s_project_bwd_args a, const BwdCounts* cnt,
                                                                        const int32_t* __restrict__ list_geom) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Camera* cams = reinterpret_cast<Camera*>(smem_raw);
    const int64_t total = cnt->n_geom;
    if ((int64_t)blockIdx.x * BWD_THREADS >= total) return;
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, (float)a.width, (float)a.height);
    for (int64_t idx = (int64_t)blockIdx.x * BWD_THREADS + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * BWD_THREADS) {
        const int64_t n = list_geom[idx];
        float v_mean[3], v_quat[4], v_scale[3], v_opac;
        const bool any = project_bwd_row(a, cams, n, v_mean, v_quat, v_scale, v_opac);
        if (any) { a.v_means[3 * n] += v_mean[0]; a.v_means[3 * n + 1] += v_mean[1]; a.v_means[3 * n + 2] += v_mean[2]; }
        a.v_scales[3 * n] = v_scale[0]; a.v_scales[3 * n + 1] = v_scale[1]; a.v_scales[3 * n + 2] = v_scale[2];
        reinterpret_cast<float4*>(a.v_quats)[n] = make_float4(v_quat[0], v_quat[1], v_quat[2], v_quat[3]);
        if (a.v_opacities) a.v_opacities[n] = v_opac;
    }
}


// =================================================================================================
// camera gradients (gsplat's `v_viewmats`; nerfstudio's camera optimizer behind source/container/src/main.py:1270-1316 reads
// them).  A separate pass, launched only when the caller asks for it, so the kernels above stay as they are: per
// (camera, Gaussian) with a live cotangent it re-derives the projection VJP's camera terms (12 values: dR, dt) and the SH
// colour's view-direction term (3 values: d campos), reduces them over the warp and the CTA (shuffles, then shared
// memory) and issues ONE atomic per value, camera and CTA.  v_view: [C,16] = dR (9, row-major), dt (3), dcampos (3), pad.
// =================================================================================================
constexpr int VW_THREADS = 256, VW_VALS = 15;
__global__ void __launch_bounds__(VW_THREADS) viewmat_bwd_kernel(const b2s_project_bwd_args a, float* __restrict__ v_view) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Camera* cams = reinterpret_cast<Camera*>(smem_raw);
    __shared__ float s_red[VW_THREADS / 32][VW_VALS];
    load_cameras(cams, a.C, a.viewmats, a.Ks, a.campos, (float)a.width, (float)a.height);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * VW_THREADS + threadIdx.x;
    const bool live = n < a.N;
    const int64_t nc = live ? n : a.N - 1;
    float mean[3], q[4], s[3];
    mean[0] = __ldg(a.means + 3 * nc); mean[1] = __ldg(a.means + 3 * nc + 1); mean[2] = __ldg(a.means + 3 * nc + 2);
    {
        const float4 qq = __ldg(reinterpret_cast<const float4*>(a.quats) + nc);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
    }
    s[0] = __ldg(a.scales + 3 * nc); s[1] = __ldg(a.scales + 3 * nc + 1); s[2] = __ldg(a.scales + 3 * nc + 2);
    const float opac = a.opacities ? __ldg(a.opacities + nc) : 1.f;
    Rot R; float qn[4], qnorm, M[9]; Sym3 S;
    quat_to_rotmat(q, R, qn, qnorm);
    covar_from_rot_scale(R, s, M, S);
    const bool with_sh = a.sh_coeffs != nullptr && a.v_colors != nullptr;
    const int K = (a.sh_degree + 1) * (a.sh_degree + 1);
    for (int c = 0; c < a.C; ++c) {
        float acc[VW_VALS];
#pragma unroll
        for (int k = 0; k < VW_VALS; ++k) acc[k] = 0.f;
        const int64_t i = (int64_t)c * a.N + nc;
        if (live && __ldg(a.radii + i) > 0) {
            ProjMid mid;
            project_mid_fast(cams[c], mean, S, a.eps2d, mid);
            if (mid.det > 0.f) {
                ProjGradIn g;
                if (a.v_means2d) { const float2 v = __ldg(reinterpret_cast<const float2*>(a.v_means2d) + i); g.v_x2d = v.x; g.v_y2d = v.y; }
                else { g.v_x2d = g.v_y2d = 0.f; }
                g.v_depth = a.v_depths ? __ldg(a.v_depths + i) : 0.f;
                if (a.v_conics) { g.v_conic_a = __ldg(a.v_conics + 3 * i); g.v_conic_b = __ldg(a.v_conics + 3 * i + 1); g.v_conic_c = __ldg(a.v_conics + 3 * i + 2); }
                else { g.v_conic_a = g.v_conic_b = g.v_conic_c = 0.f; }
                g.v_comp = a.v_compensations ? __ldg(a.v_compensations + i) : 0.f;
                if (a.v_opacities_out && a.antialiased) g.v_comp += __ldg(a.v_opacities_out + i) * opac;
                float vm[3] = {0.f, 0.f, 0.f};
                Sym3 vS = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                project_one_bwd(cams[c], mean, S, a.eps2d, mid, g, vm, vS, acc);
            }
            if (with_sh) {
                const float g0 = __ldg(a.v_colors + 3 * i), g1 = __ldg(a.v_colors + 3 * i + 1), g2 = __ldg(a.v_colors + 3 * i + 2);
                if (g0 != 0.f || g1 != 0.f || g2 != 0.f) {
                    const float* cf = a.sh_coeffs + nc * a.sh_stride_k * 3;
                    const float dx = mean[0] - cams[c].campos[0], dy = mean[1] - cams[c].campos[1], dz = mean[2] - cams[c].campos[2];
                    const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
                    const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
                    float Y[16], w[16];
                    sh_basis(a.sh_degree, ux, uy, uz, Y);
                    float cr = 0.5f, cg = 0.5f, cb = 0.5f;
                    for (int k = 0; k < K; ++k) { cr += Y[k] * __ldg(cf + 3 * k); cg += Y[k] * __ldg(cf + 3 * k + 1); cb += Y[k] * __ldg(cf + 3 * k + 2); }
                    const float vr = cr > 0.f ? g0 : 0.f, vg = cg > 0.f ? g1 : 0.f, vb = cb > 0.f ? g2 : 0.f;
                    for (int k = 0; k < K; ++k) w[k] = __ldg(cf + 3 * k) * vr + __ldg(cf + 3 * k + 1) * vg + __ldg(cf + 3 * k + 2) * vb;
                    float vx, vy, vz;
                    sh_basis_vjp(a.sh_degree, ux, uy, uz, w, vx, vy, vz);
                    const float dp = ux * vx + uy * vy + uz * vz;
                    // dirs = mean - campos: d/d campos = -d/d dirs
                    acc[12] -= (vx - ux * dp) * inv; acc[13] -= (vy - uy * dp) * inv; acc[14] -= (vz - uz * dp) * inv;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < VW_VALS; ++k) {
            const float r = warp_sum(acc[k]);
            if (lane == 0) s_red[warp][k] = r;
        }
        __syncthreads();
        if (threadIdx.x < VW_VALS) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < VW_THREADS / 32; ++w) t += s_red[w][threadIdx.x];
            if (t != 0.f) atomicAdd(v_view + 16 * c + threadIdx.x, t);
        }
        __syncthreads();
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_project_fwd(const b2s_project_fwd_args* a, b2s_stream_t stream) {
    B2S_REQUIRE(a != nullptr, "null args");
    B2S_REQUIRE(a->N >= 0 && a->C >= 0, "negative size");
    B2S_REQUIRE(a->C <= B2S_MAX_CAMERAS_PER_CALL, "too many cameras in one call (loop over chunks)");
    if (a->N == 0 || a->C == 0) return 0;
    B2S_REQUIRE(a->means && a->quats && a->scales && a->viewmats && a->Ks, "null input");
    B2S_REQUIRE(a->radii && a->means2d && a->depths && a->conics, "null output");
    B2S_REQUIRE(aligned16(a->quats) && (reinterpret_cast<uintptr_t>(a->means2d) & 7u) == 0, "quats/means2d misaligned");
    const bool sh = a->sh_coeffs != nullptr && a->colors != nullptr;
    int K = 1;
    if (sh) {
        B2S_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 3, "sh_degree must be 0..3");
        K = (a->sh_degree + 1) * (a->sh_degree + 1);
        B2S_REQUIRE(a->sh_stride_k >= K, "sh_stride_k < (deg+1)^2");
    }
    if (a->tiles_per_gauss) B2S_REQUIRE(a->tile_size > 0 && a->tile_w > 0 && a->tile_h > 0, "bad tile geometry");
    const bool vec = sh && K == 16 && a->sh_stride_k == 16 && aligned16(a->sh_coeffs);
    if (vec) {      // Blackwell data path: the coefficient rows arrive through the tensor-memory accelerator
        alignas(64) CUtensorMap map_a, map_b;
        if (sh_tensor_maps(a->sh_coeffs, a->N, &map_a, &map_b)) {
            const size_t smem_tma = 1024 + TMA_A_BYTES + TMA_B_BYTES + ((a->C * sizeof(Camera) + 15) / 16) * 16;
            B2S_CUDA(cudaFuncSetAttribute(project_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
            project_fwd_tma_kernel<<<(unsigned)cdiv(a->N, FWD_THREADS), FWD_THREADS, smem_tma, (cudaStream_t)stream>>>(*a, map_a, map_b);
            B2S_LAUNCH_CHECK();
            return 0;
        }
    }
    size_t smem = ((a->C * sizeof(Camera) + 15) / 16) * 16;
    if (sh) smem += (size_t)(FWD_THREADS / 32) * 32 * sh_row_stride(K) * sizeof(float);
    dim3 grid((unsigned)cdiv(a->N, FWD_THREADS));
    auto kern = vec ? project_fwd_kernel<16> : project_fwd_kernel<0>;
    B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, FWD_THREADS, smem, (cudaStream_t)stream>>>(*a);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t b2s_project_bwd_workspace_bytes(int64_t N) {
    return 256 + 2 * ((N + 63) / 64) * 64 * (int64_t)sizeof(int32_t);
}

extern "C" int b2s_project_bwd(const b2s_project_bwd_args* a, b2s_stream_t stream) {
    B2S_REQUIRE(a != nullptr, "null args");
    B2S_REQUIRE(a->N >= 0 && a->C >= 0, "negative size");
    B2S_REQUIRE(a->C <= B2S_MAX_CAMERAS_PER_CALL, "too many cameras in one call (loop over chunks)");
    if (a->N == 0) return 0;
    B2S_REQUIRE(a->means && a->quats && a->scales && a->viewmats && a->Ks && a->radii, "null input");
    B2S_REQUIRE(a->stages >= 0 && a->stages <= 4, "stages must be 0..4");
    const bool sh_only = a->stages == 1 || a->stages == 4;
    B2S_REQUIRE(a->v_means && (sh_only || (a->v_quats && a->v_scales)), "null gradient output");
    B2S_REQUIRE(aligned16(a->quats) && (sh_only || aligned16(a->v_quats)), "quats misaligned");
    // stages: 0 everything | 1 classify + SH | 2 projection VJP (after 1, or after 3 and 4) | 3 classify only | 4 SH only (after 3)
    const bool has_sh = a->sh_coeffs != nullptr && a->v_sh_coeffs != nullptr;
    const bool sh = has_sh && (a->stages == 0 || a->stages == 1 || a->stages == 4);
    const bool geom = a->stages == 0 || a->stages == 2;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cam_bytes = ((a->C * sizeof(Camera) + 15) / 16) * 16;
    // sparse mode: needs the caller's workspace, a single camera chunk, and (for the streaming zero fill) an aligned v_sh
    BwdCounts* cnt = nullptr;
    int32_t *list_geom = nullptr, *list_sh = nullptr;
    if (a->workspace && !a->accumulate && a->N < 2147483647LL && a->v_quats && a->v_scales) {
        if (a->workspace_bytes < b2s_project_bwd_workspace_bytes(a->N)) {
            set_error("b2s_project_bwd: workspace too small");
            return B2S_E_WORKSPACE;
        }
        B2S_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255u) == 0, "workspace must be 256-byte aligned");
        if (!has_sh || aligned16(a->v_sh_coeffs)) {
            cnt = reinterpret_cast<BwdCounts*>(a->workspace);
            list_geom = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(a->workspace) + 256);
            list_sh = list_geom + ((a->N + 63) / 64) * 64;
        }
    }
    if (a->stages == 3) B2S_REQUIRE(cnt != nullptr, "stages == 3 needs the sparse-backward workspace");
    if (cnt && (a->stages == 0 || a->stages == 1 || a->stages == 3 || (a->stages == 2 && !has_sh))) {
        // first (or only) call of this backward: zero the small gradients, build the two lists
        B2S_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BwdCounts), st));
        bwd_classify_kernel<<<(unsigned)cdiv(a->N, 256 * CLS_ROWS), 256, 0, st>>>(*a, has_sh, cnt, list_geom, list_sh);
        B2S_LAUNCH_CHECK();
    }
    if (sh) {
        B2S_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 3, "sh_degree must be 0..3");
        const int K = (a->sh_degree + 1) * (a->sh_degree + 1);
        B2S_REQUIRE(a->sh_stride_k >= K, "sh_stride_k < (deg+1)^2");
        const bool vec = K == 16 && a->sh_stride_k == 16 && aligned16(a->sh_coeffs) && aligned16(a->v_sh_coeffs);
        const bool multi = a->C > 1;
        const size_t smem = cam_bytes + (size_t)(SHB_THREADS / 32) * (multi ? 2 : 1) * 32 * sh_row_stride(K) * sizeof(float);
        auto kern = vec ? (multi ? sh_bwd_fused_kernel<16, true> : sh_bwd_fused_kernel<16, false>)
                        : (multi ? sh_bwd_fused_kernel<0, true> : sh_bwd_fused_kernel<0, false>);
        B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)min((int64_t)148 * 4, cdiv(a->N, SHB_THREADS)), SHB_THREADS, smem, st>>>(*a, cnt);
        B2S_LAUNCH_CHECK();
        if (cnt) {
            if (!a->v_sh_prezeroed) {
                sh_zero_kernel<<<148 * 32, 256, 0, st>>>(*a, cnt);
                B2S_LAUNCH_CHECK();
            }
            const unsigned grid = (unsigned)min((int64_t)148 * 16, cdiv(cdiv(a->N, 4), 128));
            if (vec) sh_sparse_kernel<true><<<grid, 128, cam_bytes, st>>>(*a, cnt, list_sh);
            else sh_sparse_kernel<false><<<grid, 128, cam_bytes, st>>>(*a, cnt, list_sh);
            B2S_LAUNCH_CHECK();
        }
    }
    if (geom) {
        // stages == 2 with the SH pointers set: the SH part is already in v_means (an earlier stages == 1 call)
        const bool means_acc = has_sh;
        if (cnt) {      // the small gradients are zero-filled and the live rows listed: one list-driven kernel, whatever the density
            project_bwd_sparse_kernel<<<(unsigned)min((int64_t)148 * 16, cdiv(a->N, BWD_THREADS)), BWD_THREADS, cam_bytes, st>>>(
                *a, cnt, list_geom);
        } else {
            auto kern2 = means_acc ? project_bwd_kernel<true> : project_bwd_kernel<false>;
            kern2<<<(unsigned)cdiv(a->N, BWD_THREADS), BWD_THREADS, cam_bytes, st>>>(*a, nullptr);
        }
        B2S_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int b2s_viewmat_bwd(const b2s_project_bwd_args* a, float* v_view, b2s_stream_t stream) {
    B2S_REQUIRE(a != nullptr && v_view != nullptr, "null args");
    B2S_REQUIRE(a->N >= 0 && a->C >= 0 && a->C <= B2S_MAX_CAMERAS_PER_CALL, "bad sizes");
    if (a->N == 0 || a->C == 0) return 0;
    B2S_REQUIRE(a->means && a->quats && a->scales && a->viewmats && a->Ks && a->radii, "null input");
    B2S_REQUIRE(aligned16(a->quats), "quats misaligned");
    if (a->sh_coeffs && a->v_colors) {
        B2S_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 3, "sh_degree must be 0..3");
        B2S_REQUIRE(a->sh_stride_k >= (a->sh_degree + 1) * (a->sh_degree + 1), "sh_stride_k < (deg+1)^2");
    }
    const size_t cam_bytes = ((a->C * sizeof(Camera) + 15) / 16) * 16;
    viewmat_bwd_kernel<<<(unsigned)cdiv(a->N, VW_THREADS), VW_THREADS, cam_bytes, (cudaStream_t)stream>>>(*a, v_view);
    B2S_LAUNCH_CHECK();
    return 0;
}
