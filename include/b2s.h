/* b2s -- B200-native Gaussian-splat rasterizer, C ABI.
 *
 * This is the drop-in boundary for the `rasterization()` hot path that the reference's
 * Train-Stage1 component reaches through gsplat (reference call sites:
 * source/container/src/main.py:1308-1316 `ns-train splatfacto`, and
 * source/container/src/main.py:1339-1347 `gsplat/examples/simple_trainer.py --packed`).
 * gsplat binds its kernels through a torch C++ extension (`gsplat/cuda/_wrapper.py` ->
 * `gsplat.csrc`); each entry point below names the gsplat op it replaces.  gsplat itself is
 * not in /root/reference (unpinned pip install, source/container/Dockerfile:220-229), so the
 * cited op names are the upstream ones, see SURVEY.md section 8a.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*.
 *  - the caller owns every buffer, including scratch: query `*_workspace_bytes` first.
 *    The library never allocates device memory, never creates streams, never calls
 *    cudaSetDevice, and holds no global mutable state (re-entrant; forward is called on the
 *    Python thread, backward on torch's autograd thread).  The one exception is the file-level
 *    converter at the end (b2s_convert_spz_to_ply / ConvertSpzToPly: host paths in, file out).
 *  - every call takes the CUDA stream explicitly (`cudaStream_t` passed as void*) and only
 *    enqueues work; nothing synchronises.
 *  - return value: 0 = ok, negative = B2S_E_* argument error, positive = cudaError_t.
 *    Never aborts, never throws.  `b2s_last_error()` gives a thread-local message.
 *  - all floating point is fp32, ids are int32, sort keys int64 (as gsplat).
 *  - layouts are C-contiguous; "pair" = one (camera, Gaussian) row, dense index c*N + n.
 */
#ifndef B2S_H
#define B2S_H
#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B2S_API
#else
#define B2S_API __attribute__((visibility("default")))
#endif

#define B2S_VERSION 100
#define B2S_E_BADARG (-1)
#define B2S_E_UNSUPPORTED (-2)
#define B2S_E_WORKSPACE (-3)
#define B2S_TOTAL_SLOTS 32
#define B2S_MAX_CAMERAS_PER_CALL 64 /* project_* process cameras in chunks of this size (the host loops) */

typedef void* b2s_stream_t; /* cudaStream_t */

B2S_API int b2s_version(void);
B2S_API const char* b2s_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 forward: fused EWA projection (+ SH colour, + tile count).
 * Replaces gsplat `fully_fused_projection` fwd (+ `spherical_harmonics` fwd + the count half of
 * `isect_tiles`), SURVEY.md 8a rows 1-4.  One thread per Gaussian, cameras looped inside so the
 * 44 B of parameters and the 12*K B of SH coefficients are read once.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t N;              /* Gaussians */
    int32_t C;              /* cameras in this call, <= B2S_MAX_CAMERAS_PER_CALL */
    const float* means;     /* [N,3] */
    const float* quats;     /* [N,4] (w,x,y,z), un-normalised */
    const float* scales;    /* [N,3] post-exp */
    const float* opacities; /* [N] post-sigmoid, or NULL */
    const float* viewmats;  /* [C,4,4] world->camera */
    const float* Ks;        /* [C,3,3] */
    int32_t width, height;
    float eps2d, near_plane, far_plane, radius_clip;
    int32_t antialiased;    /* 1: opacities_out = opacity * compensation */
    /* SH colour (sh_coeffs == NULL: skipped) */
    const float* sh_coeffs; /* [N, sh_stride_k, 3] */
    const float* campos;    /* [C,3] camera centres in world space, or NULL: derived from viewmats */
    int32_t sh_degree;      /* 0..3, uses the first (deg+1)^2 rows */
    int32_t sh_stride_k;    /* rows per Gaussian in sh_coeffs */
    /* tile count (tiles_per_gauss == NULL: skipped) */
    int32_t tile_size, tile_w, tile_h;
    /* outputs, all [C,N,...] */
    int32_t* radii;           /* [C,N]   0 = culled */
    float* means2d;           /* [C,N,2] */
    float* depths;            /* [C,N] */
    float* conics;            /* [C,N,3] */
    float* compensations;     /* [C,N] or NULL */
    float* opacities_out;     /* [C,N] or NULL */
    float* colors;            /* [C,N,3] = max(SH + 0.5, 0), or NULL */
    int32_t* tiles_per_gauss; /* [C,N] or NULL */
    uint32_t* depth_keys;     /* [C,N] or NULL: depth bits, 0xFFFFFFFF on culled rows (for b2s_depth_order) */
    int32_t* row_ids;         /* [C,N] or NULL: iota, c*N + n */
    int64_t* total_isects;    /* device int64[B2S_TOTAL_SLOTS][4] or NULL, zero it first: slot sums give
                                 [0] sum of tiles_per_gauss, [1] visible rows, [2] coarse-bin records (b2s_coarse_bin),
                                 [3] unused (spread over slots to avoid same-address atomics) */
} b2s_project_fwd_args;
B2S_API int b2s_project_fwd(const b2s_project_fwd_args* a, b2s_stream_t stream);

/* Stage 1 backward.  Replaces gsplat `fully_fused_projection` bwd + `spherical_harmonics` bwd
 * (SURVEY.md 8a row 9).  Cotangent pointers may be NULL (treated as zero).  Gradients are
 * *written* (not accumulated) unless `accumulate` is set (used by the host loop over camera chunks). */
typedef struct {
    int64_t N;
    int32_t C;
    const float* means; const float* quats; const float* scales; const float* opacities;
    const float* viewmats; const float* Ks;
    int32_t width, height;
    float eps2d, near_plane, far_plane, radius_clip;
    int32_t antialiased;
    const float* sh_coeffs; const float* campos; int32_t sh_degree; int32_t sh_stride_k;
    const int32_t* radii;        /* [C,N] from the forward */
    /* cotangents [C,N,...] */
    const float* v_means2d; const float* v_depths; const float* v_conics;
    const float* v_compensations; /* only when the caller consumed `compensations` directly */
    const float* v_opacities_out; const float* v_colors;
    int32_t accumulate;
    /* gradients */
    float* v_means;     /* [N,3] */
    float* v_quats;     /* [N,4] */
    float* v_scales;    /* [N,3] */
    float* v_opacities; /* [N] or NULL */
    float* v_sh_coeffs; /* [N, sh_stride_k, 3] or NULL; rows >= (deg+1)^2 are not touched */
    int32_t stages;         /* 0 = both; 1 = SH colour VJP only (v_sh_coeffs, v_means); 2 = projection VJP only (when
                               sh_coeffs and v_sh_coeffs are set it ADDS to the v_means a stages == 1 call left);
                               3 = the classify pass alone (needs the workspace), 4 = the SH part of a stages == 1 call after
                               a stages == 3 call (1 == 3 then 4: lets a caller time the three parts separately) */
    /* optional scratch of b2s_project_bwd_workspace_bytes(N) bytes, 256-byte aligned (NULL: dense kernels only).
     * With it, rows whose cotangents are all zero (occluded: compositing never reached them) are filtered out once:
     * when at most a quarter of the rows is live the gradients are produced by a zero fill plus list-driven kernels.
     * A stages == 2 call must be given the workspace its stages == 1 call used. */
    void* workspace;
    int64_t workspace_bytes;
    int32_t v_sh_prezeroed;  /* != 0: the caller already zero-filled v_sh_coeffs (e.g. on a second stream, under the forward's
                                compositing kernel, which leaves HBM idle): the sparse mode skips its own zero fill */
} b2s_project_bwd_args;
/* Streaming zero fill of a device buffer (any alignment), for gradient buffers that are accumulated into. */
B2S_API int b2s_zero_fill(void* p, int64_t bytes, b2s_stream_t stream);
B2S_API int64_t b2s_project_bwd_workspace_bytes(int64_t N);
B2S_API int b2s_project_bwd(const b2s_project_bwd_args* a, b2s_stream_t stream);
/* Camera gradients of the same backward (gsplat's `v_viewmats`; only nerfstudio's optional camera optimizer asks for them):
 * v_view [C,16], zero-filled by the caller, accumulated with one atomic per value, camera and CTA after a warp + CTA
 * reduction: [0..8] d/dR (row-major world->camera rotation), [9..11] d/dt, [12..14] d/d(camera centre) from the SH view
 * directions (the caller chains it through campos = -R^-1 t), [15] unused.  Takes the arguments of b2s_project_bwd
 * (gradient outputs, stages and workspace are ignored). */
B2S_API int b2s_viewmat_bwd(const b2s_project_bwd_args* a, float* v_view, b2s_stream_t stream);

/* Stand-alone SH evaluation over M directions (gsplat `spherical_harmonics` fwd/bwd): the
 * stage-level op; the training path uses the fused kernels above.  masks may be NULL. */
B2S_API int b2s_sh_fwd(int64_t M, int32_t degree, int32_t stride_k, const float* dirs, const float* coeffs,
                       const uint8_t* masks, float* colors, b2s_stream_t stream);
B2S_API int b2s_sh_bwd(int64_t M, int32_t degree, int32_t stride_k, const float* dirs, const float* coeffs,
                       const uint8_t* masks, const float* v_colors, float* v_coeffs, float* v_dirs /* or NULL */,
                       b2s_stream_t stream);

/* gsplat `quat_scale_to_covar_preci` fwd: covars/precis [N,3,3] (either may be NULL). */
B2S_API int b2s_quat_scale_to_covar_preci(int64_t N, const float* quats, const float* scales, float* covars,
                                          float* precis, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2: tile binning.  Replaces gsplat `isect_tiles` (count / fill), the CUB radix sort it
 * calls, and `isect_offset_encode`.  All outputs bit-exact.
 * Two-phase because n_isects is data dependent:
 *   b2s_isect_count  -> tiles_per_gauss [M]                          (skipped when stage 1 made it)
 *   b2s_isect_scan   -> cum_tiles [M] (exclusive), *total (device int64; the caller copies it to
 *                       the host to size the next buffers -- the one host sync of a forward)
 *   b2s_isect_fill   -> isect_ids [I] int64, flatten_ids [I] int32   (unsorted, pair-major)
 *   b2s_isect_sort   -> sorted isect_ids / flatten_ids                (stable LSD radix)
 *   b2s_isect_offsets-> offsets [C*tile_h*tile_w] int32
 * M = number of rows (C*N dense, nnz packed).  camera_ids == NULL means dense rows (cam = row / N).
 *
 * Tile-row window (`rows`, a HOST struct read during the call; NULL = the whole grid): restricts every
 * tile rectangle to tile rows [row_lo, row_hi) of the image's tile grid.  This is how a rank of a multi-GPU
 * job bins only the tile stripe it owns: tile ids, keys and offsets stay those of the whole
 * image, so a stripe's sorted list is exactly the single-GPU list filtered by tile row.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t row_lo, row_hi; } b2s_tile_rows;
B2S_API int b2s_isect_count(int64_t M, const float* means2d, const int32_t* radii, int32_t tile_size,
                            int32_t tile_w, int32_t tile_h, int32_t* tiles_per_gauss, const b2s_tile_rows* rows,
                            b2s_stream_t stream);
B2S_API int64_t b2s_isect_scan_workspace_bytes(int64_t M);
B2S_API int b2s_isect_scan(int64_t M, const int32_t* tiles_per_gauss, int32_t* cum_tiles, int64_t* total,
                           void* workspace, int64_t workspace_bytes, b2s_stream_t stream);
/* Packed layout (gsplat `packed=True`): sel = the indices of the rows with radius > 0, ascending (camera-major,
 * Gaussian-ascending), at most cap of them; total[0] (device) = their number.  Workspace as b2s_isect_scan. */
B2S_API int b2s_select_visible(int64_t M, const int32_t* radii, int64_t* sel, int64_t cap, int64_t* total, void* workspace,
                               int64_t workspace_bytes, b2s_stream_t stream);
B2S_API int b2s_isect_fill(int64_t M, int64_t N, const int64_t* camera_ids /* or NULL */, const float* means2d,
                           const int32_t* radii, const float* depths, const int32_t* cum_tiles, int32_t tile_size,
                           int32_t tile_w, int32_t tile_h, int64_t* isect_ids, int32_t* flatten_ids,
                           const b2s_tile_rows* rows, b2s_stream_t stream);
B2S_API int64_t b2s_isect_sort_workspace_bytes(int64_t n_isects);
B2S_API int b2s_isect_sort(int64_t n_isects, int32_t key_bits, const int64_t* isect_ids_in, const int32_t* flatten_ids_in,
                           int64_t* isect_ids_out, int32_t* flatten_ids_out, void* workspace, int64_t workspace_bytes,
                           b2s_stream_t stream);
B2S_API int b2s_isect_offsets(int64_t n_isects, const int64_t* isect_ids_sorted, int32_t C, int32_t tile_w,
                              int32_t tile_h, int32_t* offsets, b2s_stream_t stream);

/* Depth-first binning -- the training path's replacement for fill + 64-bit sort + offset_encode.
 * Same sorted order as the reference (ascending (camera, tile, depth bits), ties by ascending flat
 * index) with ~3x less HBM traffic: Gaussians are depth-sorted once (M 32-bit keys), expanded to tile
 * intersections in that order, then stable-split by (camera, tile) with an LSD radix over
 * tile_bits+cam_bits bits of 8-byte records.  The 64-bit isect_ids are rebuilt on demand.
 *   b2s_isect_depth_keys : depth bits (0xFFFFFFFF on culled rows) + row ids   (stage 1 can emit both,
 *                          together with *total_isects, the data-dependent size the host must read)
 *   b2s_depth_order      : sorted_ids [M]  (32-bit LSD radix sort of the depth keys; enqueue it BEFORE
 *                          reading total_isects so the sort overlaps the host's wait)
 *   b2s_tile_bin         : fused scan + expansion in depth order, (camera,tile) radix sort, offsets:
 *                          tile_keys_sorted [I] (cam << tile_bits | tile; 16-bit when they fit, see
 *                          b2s_tile_key_bytes), flatten_ids [I], offsets [C*tile_h*tile_w]
 *   b2s_isect_ids_from_bins : isect_ids [I] = tile_key << 32 | bits(depths[flatten_ids])
 * Workspaces must be 256-byte aligned. */
B2S_API int b2s_isect_depth_keys(int64_t M, const float* depths, const int32_t* radii, uint32_t* keys, int32_t* vals,
                                 b2s_stream_t stream);
B2S_API int64_t b2s_depth_order_workspace_bytes(int64_t M);
B2S_API int b2s_depth_order(int64_t M, const uint32_t* depth_keys, const int32_t* row_ids, int32_t* sorted_ids,
                            void* workspace, int64_t workspace_bytes, b2s_stream_t stream);
B2S_API int32_t b2s_tile_key_bytes(int32_t C, int32_t tile_w, int32_t tile_h);
B2S_API int64_t b2s_tile_bin_workspace_bytes(int64_t M, int64_t n_isects);
B2S_API int b2s_tile_bin(int64_t M, int64_t N, const int64_t* camera_ids /* or NULL */, int64_t n_isects,
                         const int32_t* sorted_ids, const float* means2d, const int32_t* radii, int32_t C,
                         int32_t tile_size, int32_t tile_w, int32_t tile_h, void* tile_keys_sorted /* [I] x 4 bytes */,
                         int32_t* flatten_ids, int32_t* offsets, void* workspace, int64_t workspace_bytes,
                         const b2s_tile_rows* rows, b2s_stream_t stream);
B2S_API int b2s_isect_ids_from_bins(int64_t n_isects, int32_t key_bytes, const void* tile_keys_sorted,
                                    const int32_t* flatten_ids, const float* depths, int64_t* isect_ids,
                                    b2s_stream_t stream);

/* Two-level binning -- what `rasterization()` runs.  Compositing stops at the first few percent of a tile's
 * depth-sorted list once its pixels saturate, so materialising every (tile, Gaussian) intersection is mostly
 * wasted work.  Instead the depth-ordered rows are binned into COARSE bins of B2S_COARSE x B2S_COARSE tiles
 * (one record per (row, coarse bin) it overlaps, ~4x fewer records than tile intersections), and the compositing
 * kernels expand a tile's list on the fly: they walk the coarse bin's list in order and keep the records whose
 * tile mask covers the tile.  The kept records are exactly the tile's slice of the reference's sorted list
 * (same order: depth bits, ties by flat index), so renders and gradients are unchanged.
 *   record = mask_y(8) << 24 | mask_x(8) << 16 | coarse key(16), coarse key = (cam*bins_h + by)*bins_w + bx;
 *   mask_x bit i set <=> the row's tile rectangle covers tile column 8*bx + i (same for y).
 *   b2s_coarse_count : *total += number of (row, coarse bin) records        (stage 1 can emit this itself)
 *   b2s_coarse_bin   : bin_recs [n] / bin_ids [n] sorted by coarse key (stable: depth order inside a bin),
 *                      bin_offsets [C*bins_h*bins_w + 1].  bin_recs / bin_ids must be allocated with n rounded
 *                      up to a multiple of 4 entries and 16-byte aligned (the kernels read them 128 bits at a time).
 * Needs C * ceil(tile_w/8) * ceil(tile_h/8) <= 65536 (B2S_E_UNSUPPORTED otherwise). */
#define B2S_COARSE 8
B2S_API int b2s_coarse_count(int64_t M, const float* means2d, const int32_t* radii, int32_t tile_size, int32_t tile_w,
                             int32_t tile_h, int64_t* total /* device, accumulated */, const b2s_tile_rows* rows,
                             b2s_stream_t stream);
B2S_API int64_t b2s_coarse_bin_workspace_bytes(int64_t M, int64_t n_records);
B2S_API int b2s_coarse_bin(int64_t M, int64_t N, const int64_t* camera_ids /* or NULL */, int64_t n_records,
                           const int32_t* sorted_ids, const float* means2d, const int32_t* radii, int32_t C,
                           int32_t tile_size, int32_t tile_w, int32_t tile_h, uint32_t* bin_recs, int32_t* bin_ids,
                           int32_t* bin_offsets, void* workspace, int64_t workspace_bytes, const b2s_tile_rows* rows,
                           b2s_stream_t stream);

/* The same binning with half the gather and four launches fewer (what `rasterization()` runs; the pair above is kept
 * as the unfused reference and for A/B, B2S_BINNING_FUSED=0, and serves tile grids beyond 255 x 255):
 *   b2s_depth_order_rects : b2s_depth_order whose histogram kernel -- a streaming pass over the rows in row order, so the
 *                           loads coalesce -- also derives every row's tile rectangle and stores it BY ROW, packed into 32 bits:
 *                           rects [M] = x0 | x1 << 8 | y0 << 16 | y1 << 24 (tile units, rows clipped to the window, 0 on
 *                           culled rows; every row id in row_ids must be < M); the last pass stores sorted_ids [M] and
 *                           never writes the sorted keys.  Needs nothing the host has to wait for, so it too is
 *                           enqueued before the host reads the record count.  tile_w, tile_h <= 255.
 *   b2s_coarse_bin_rects  : one kernel gathers the rectangles of the depth-ordered rows (one word per row), scans the
 *                           record counts (decoupled look-back), writes the records and counts the digits of the bin
 *                           sort on the way; one radix pass (two above 256 bins) sorts them; with one pass the scanned
 *                           digit histogram is the offset table.
 * Outputs are bit-identical to b2s_depth_order + b2s_coarse_bin. */
B2S_API int64_t b2s_depth_order_rects_workspace_bytes(int64_t M);
B2S_API int b2s_depth_order_rects(int64_t M, const uint32_t* depth_keys, const int32_t* row_ids, const float* means2d,
                                  const int32_t* radii, int32_t tile_size, int32_t tile_w, int32_t tile_h,
                                  const b2s_tile_rows* rows, int32_t* sorted_ids, uint32_t* rects, void* workspace,
                                  int64_t workspace_bytes, b2s_stream_t stream);
B2S_API int64_t b2s_coarse_bin_rects_workspace_bytes(int64_t M, int64_t n_records);
B2S_API int b2s_coarse_bin_rects(int64_t M, int64_t N, const int64_t* camera_ids /* or NULL */, int64_t n_records,
                                 const int32_t* sorted_ids, const uint32_t* rects, int32_t C, int32_t tile_w, int32_t tile_h,
                                 uint32_t* bin_recs, int32_t* bin_ids, int32_t* bin_offsets, void* workspace,
                                 int64_t workspace_bytes, b2s_stream_t stream);

/* Packed layout (gsplat `packed=True`): gather the visible rows `sel` (ascending dense indices c*N+n) of the
 * stage-1 outputs into nnz-row tensors (+ camera_ids / gaussian_ids), and scatter gradients back into
 * zero-filled dense tensors.  Any output pointer may be NULL (skipped). */
B2S_API int b2s_pack_rows(int64_t nnz, int64_t N, const int64_t* sel, int32_t channels, const float* means2d,
                          const float* depths, const float* conics, const float* opacities, const float* colors,
                          const int32_t* radii, const int32_t* tiles, float* means2d_p, float* depths_p, float* conics_p,
                          float* opacities_p, float* colors_p, int32_t* radii_p, int32_t* tiles_p, int64_t* camera_ids,
                          int64_t* gaussian_ids, b2s_stream_t stream);
B2S_API int b2s_unpack_rows(int64_t nnz, const int64_t* sel, int32_t channels, const float* g_means2d,
                            const float* g_depths, const float* g_conics, const float* g_opacities, const float* g_colors,
                            float* v_means2d, float* v_depths, float* v_conics, float* v_opacities, float* v_colors,
                            b2s_stream_t stream);

/* Multi-GPU routing, UNFUSED reference implementation (B2S_ROUTE=nccl; the default path is the fused stripe / camera-owner
 * exchange further down): the visible splats of the local shard travel to their owners as ONE fused record each, [x, y, depth, conic(3), opacity, colour(ch), radius, tiles, local
 * camera, global Gaussian id] = 11 + ch floats (integers bit-cast), so a single NCCL all-to-all moves them;
 * gradients come back as [7 + ch] records by the reverse all-to-all and are scattered into dense tensors. */
B2S_API int b2s_pack_payload(int64_t nnz, int64_t N, const int64_t* sel, int32_t channels, int32_t cams_per_rank,
                             int64_t gid_offset, const float* means2d, const float* depths, const float* conics,
                             const float* opacities, const float* colors, const int32_t* radii, const int32_t* tiles,
                             float* payload, b2s_stream_t stream);
B2S_API int b2s_split_payload(int64_t n, int32_t channels, const float* payload, float* means2d, float* depths,
                              float* conics, float* opacities, float* colors, int32_t* radii, int32_t* tiles,
                              int64_t* camera_ids, int64_t* gaussian_ids, b2s_stream_t stream);
B2S_API int b2s_join_grads(int64_t n, int32_t channels, const float* g_means2d, const float* g_depths,
                           const float* g_conics, const float* g_opacities, const float* g_colors, float* out,
                           b2s_stream_t stream);
B2S_API int b2s_scatter_grads(int64_t nnz, const int64_t* sel, int32_t channels, const float* in, float* v_means2d,
                              float* v_depths, float* v_conics, float* v_opacities, float* v_colors,
                              int32_t accumulate /* 1: rows may repeat in sel (tile stripes), records are summed */,
                              b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU tile stripes (rasterization(distributed=True, tile_stripes=True); SURVEY.md 8e-ii, BASELINE config 5):
 * every rank owns a shard of the Gaussians and the tile rows [bounds[r], bounds[r+1]) of the shared camera(s).
 * The exchange is fused into the producing kernels over NVLink peer mappings of symmetric memory; the caller
 * provides the barriers.  No counterpart in the reference (its multi-GPU branch routes to camera owners, above).
 *
 *   b2s_stripe_count : dest_mask [M] (bit d: the row's tile rectangle touches stripe d) and per-CTA partial sums per
 *                      destination, cta_counts [3][world][b2s_stripe_ctas(M)] = rows, tile intersections, coarse-bin
 *                      records; totals [world][4] int64: slot 3 = tag (the other slots are written by the scan)
 *   b2s_stripe_scan  : cta_counts[0] -> exclusive prefix along the CTAs, per destination (in place);
 *                      totals [world][4] = (rows, tile intersections, coarse-bin records, tag) per destination
 *   b2s_p2p_put      : words_per_peer 64-bit words from src + d*src_stride to host_peer_dst[d] (NULL: skipped)
 *   b2s_stripe_pack  : stores every row's record into each destination's receive buffer at
 *                      host_dest_base[d] + its rank among this rank's rows for d (ascending row order).
 *                      Receive buffer: one record of 10+ch 32-bit words per row = [x, y, depth, conic(3), opacity,
 *                      colour(ch), radius, camera, global Gaussian id]; a CTA's rows for one peer leave as one
 *                      contiguous, coalesced run
 *   b2s_stripe_unpack: receiver, rows [0,n) of its buffer -> tensors (+ int64 camera_ids / gaussian_ids), and the
 *                      inputs of binning: tiles_per_gauss inside the window `rows`, depth_keys, row_ids
 *   b2s_stripe_return: backward; rows with a non-zero gradient are appended as [source row, v_xy, v_depth, v_conic,
 *                      v_opacity, v_colour] = 8+ch words to this rank's region (first slot host_region_base[s]) of
 *                      source rank s's return buffer; counters [world] (zeroed by the call) = records per source
 *   b2s_stripe_scatter_add : source; adds the ret_counts[d] records of every region into the dense zero-filled gradients
 * ---------------------------------------------------------------------------------------------- */
#define B2S_MAX_RANKS 16
typedef struct {
    int32_t world;                      /* stripes = ranks, 1..B2S_MAX_RANKS */
    int32_t cams_per_rank;              /* 0: tile stripes.  > 0: CAMERA OWNERS instead (upstream's scheme): rank d owns
                                           cameras [d*cams_per_rank, (d+1)*cams_per_rank) and their whole tile grid;
                                           the same kernels then route every visible row to its camera's owner */
    int32_t bounds[B2S_MAX_RANKS + 1];  /* tile stripes: ascending tile rows, bounds[0] = 0, bounds[world] = tile_h */
} b2s_stripes;
B2S_API int64_t b2s_stripe_ctas(int64_t M);
B2S_API int b2s_stripe_count(int64_t M, int64_t N /* rows per camera */, const float* means2d, const int32_t* radii,
                             int32_t tile_size, int32_t tile_w, int32_t tile_h, const b2s_stripes* stripes /* host */, uint16_t* dest_mask, uint32_t* cta_counts,
                             int64_t* totals, int64_t tag /* stored in totals[d][3], e.g. this rank's Gaussian count */,
                             b2s_stream_t stream);
B2S_API int b2s_stripe_scan(int64_t M, int32_t world, uint32_t* cta_counts, int64_t* totals, b2s_stream_t stream);
B2S_API int b2s_p2p_put(int32_t world, const int64_t* src, int32_t words_per_peer, int32_t src_stride,
                        void* const* host_peer_dst, b2s_stream_t stream);
B2S_API int b2s_stripe_pack(int64_t M, int64_t N, int32_t channels, int64_t gid_offset, const float* means2d,
                            const float* depths, const float* conics, const float* opacities, const float* colors,
                            const int32_t* radii, const uint16_t* dest_mask, const uint32_t* cta_prefix, int32_t world,
                            int32_t cams_per_rank, const int64_t* host_dest_base, void* const* host_peer_buffers,
                            b2s_stream_t stream);
B2S_API int b2s_stripe_unpack(int64_t n, int32_t channels, const float* recv_buffer, int32_t tile_size,
                              int32_t tile_w, int32_t tile_h, const b2s_tile_rows* rows, float* means2d, float* depths,
                              float* conics, float* opacities, float* colors, int32_t* radii, int64_t* camera_ids,
                              int64_t* gaussian_ids, int32_t* tiles_per_gauss, uint32_t* depth_keys, int32_t* row_ids,
                              b2s_stream_t stream);
B2S_API int b2s_stripe_return(int64_t n, int32_t channels, const float* g_means2d, const float* g_depths,
                              const float* g_conics, const float* g_opacities, const float* g_colors,
                              const int64_t* camera_ids, const int64_t* gaussian_ids,
                              int64_t cam_offset /* camera owners: this rank's first global camera; stripes: 0 */,
                              int32_t world, const int64_t* host_recv_prefix /* [world+1] */, const int64_t* host_gid_offset,
                              const int64_t* host_n_src, const int64_t* host_region_base, int64_t* counters,
                              void* const* host_peer_buffers, b2s_stream_t stream);
B2S_API int b2s_stripe_scatter_add(int32_t world, int32_t channels, const float* ret_buffer, const int64_t* host_region_base,
                                   const int64_t* host_region_cap, const int64_t* ret_counts /* device [world] */,
                                   float* v_means2d, float* v_depths, float* v_conics, float* v_opacities, float* v_colors,
                                   b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse gradient synchronisation for camera data-parallel training (BASELINE config 4; csrc/gradsync.cu):
 * instead of all-reducing 236 B per Gaussian, every rank compacts the rows that carry a gradient into records
 * [row, v_means 3, v_quats 4, v_scales 3, v_opacity, v_sh sh_floats] (12 + sh_floats words), the records are
 * all-gathered and added into every rank's dense gradients.  total[0] (device, zeroed by the call) = live rows; at most
 * `cap` records are written -- when total exceeds cap the caller falls back to the dense all-reduce.
 * ---------------------------------------------------------------------------------------------- */
B2S_API int b2s_grad_compact(int64_t N, int32_t sh_floats, const float* v_means, const float* v_quats, const float* v_scales,
                             const float* v_opacities /* or NULL */, const float* v_sh /* [N, sh_floats] or NULL */,
                             int64_t* total, float* records, int64_t cap, b2s_stream_t stream);
B2S_API int b2s_grad_scatter_add(int64_t n_records, int32_t sh_floats, const float* records, float* v_means, float* v_quats,
                                 float* v_scales, float* v_opacities, float* v_sh, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 FUSED with the multi-GPU exchange (the "direct" route; csrc/route.cu).  Replaces, for
 * `rasterization(distributed=True)`, upstream's projection -> pack -> all_to_all(counts) -> all_to_all(payload) ->
 * split (gsplat/rendering.py + gsplat/distributed.py behind source/container/src/main.py:1317-1347) and this library's
 * own b2s_project_fwd + b2s_stripe_count/scan/pack/unpack: ONE kernel projects the rank's shard into every camera,
 * evaluates the colour and stores each visible row, as a finished record, straight into the receive buffer of the rank
 * that composites it (peer pointers of symmetric memory), plus per-CTA counts into that rank's count table.
 *
 * Receive buffer of a rank: world * region_cams REGIONS of cap_rows records of (10 + channels) words
 *   [x, y, depth, conic(3), opacity, colour(ch), radius, local camera, LOCAL Gaussian index];
 * region (source s, camera lc) = s * region_cams + lc; inside it source CTA b (Gaussians [256 b, 256 b + 256)) owns the
 * slots [256 b, 256 b + 256) and fills them from the front in ascending Gaussian order.  Count table: int32[4] per
 * (region, source CTA) = rows, tile intersections, coarse-bin records (inside the destination's window), tag
 * (CTA 0 of a region: the source's Gaussian count).  cap_rows: a multiple of 256, >= every rank's Gaussian count;
 * a source whose count exceeds it (or overflow != 0) publishes only its tag -- the receivers see it in info[3] and
 * the caller grows the buffers collectively and repeats the call.  The caller provides ONE barrier between
 * b2s_route_project and b2s_route_scan and alternates two buffer sets (parity) between consecutive exchanges.
 *   b2s_route_scan   : row_base[region][source CTA] = exclusive prefix of the row counts inside the source (camera, then
 *                      Gaussian order); info (int64 [4 + 4 world]): info[4 + 4 s ..] = (rows, Gaussian count, tile
 *                      intersections, coarse-bin records) of source s (one CTA per source; the host adds them up)
 *   b2s_route_gather : occupied slots -> compact tensors (camera_ids local, gaussian_ids GLOBAL = sum of the lower
 *                      ranks' counts + local index) and the inputs of binning, as b2s_stripe_unpack
 * The backward returns the gradients with b2s_stripe_return / b2s_stripe_scatter_add.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    b2s_project_fwd_args proj;   /* N, C (ALL cameras of the job), parameters, cameras, sh_coeffs / sh_degree / sh_stride_k,
                                    tile geometry; of the outputs only `radii` (dense [C,N], read by the backward) is written */
    int32_t world, rank;
    int32_t cams_per_rank;       /* > 0: camera owners (camera c is composited by rank c / cams_per_rank); 0: tile stripes */
    int32_t region_cams;         /* cameras per destination: cams_per_rank, or C for tile stripes */
    int32_t channels;            /* colour channels of a record, 1..4 */
    int32_t color_mode;          /* 0: SH colour (3 channels); 1: per-Gaussian colours_in [N, n_color_in] */
    int32_t n_color_in;
    int32_t depth_channel;       /* 1: the depth is appended as the last channel (render modes RGB+D, D) */
    int32_t overflow;            /* != 0: publish the tag only */
    int32_t bounds[B2S_MAX_RANKS + 1];   /* tile stripes: ascending tile rows from 0 to tile_h */
    int64_t cap_rows;
    const float* colors_in;
    void* peer_rec[B2S_MAX_RANKS];       /* every rank's receive buffer (this parity), peer-mapped */
    void* peer_cnt[B2S_MAX_RANKS];       /* every rank's count table (this parity), peer-mapped, 16-byte aligned */
} b2s_route_args;
B2S_API int64_t b2s_route_rec_floats(int32_t world, int32_t region_cams, int64_t cap_rows);   /* sized for 4 channels */
B2S_API int64_t b2s_route_cnt_ints(int32_t world, int32_t region_cams, int64_t cap_rows);
B2S_API int b2s_route_project(const b2s_route_args* a, b2s_stream_t stream);
B2S_API int b2s_route_scan(int32_t world, int32_t region_cams, int64_t cap_rows, const int32_t* cnt, int32_t* row_base,
                           int64_t* info, b2s_stream_t stream);
B2S_API int b2s_route_gather(int32_t world, int32_t region_cams, int64_t cap_rows, int32_t channels, const float* rec,
                             const int32_t* cnt, const int32_t* row_base, const int64_t* host_n_src,
                             const int64_t* host_rows_from /* [world]: rows per source, from info */, int64_t n_rows,
                             int32_t tile_size, int32_t tile_w, int32_t tile_h, const b2s_tile_rows* rows, float* means2d,
                             float* depths, float* conics, float* opacities, float* colors, int32_t* radii, int64_t* camera_ids,
                             int64_t* gaussian_ids, int32_t* tiles_per_gauss, uint32_t* depth_keys, int32_t* row_ids,
                             b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3: per-16x16-tile front-to-back compositing.  Replaces gsplat `rasterize_to_pixels`
 * fwd / bwd.  Rows of means2d/conics/colors/opacities are addressed by
 * flatten_ids, so dense [C*N] and packed [nnz] layouts both work.  channels in {1,2,3,4,8,16,32}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t C, width, height, tile_size, tile_w, tile_h, channels;
    int64_t n_isects;
    const float* means2d;       /* [M,2] */
    const float* conics;        /* [M,3] */
    const float* colors;        /* [M,channels] */
    const float* opacities;     /* [M] */
    const float* backgrounds;   /* [C,channels] or NULL */
    const int32_t* isect_offsets; /* [C,tile_h,tile_w] */
    const int32_t* flatten_ids; /* [n_isects] */
    float* render_colors;       /* [C,H,W,channels] */
    float* render_alphas;       /* [C,H,W,1] */
    int32_t* last_ids;          /* [C,H,W] index into flatten_ids of the last contributor */
    float* final_T;             /* [C,H,W] exact final transmittance for the backward (negated on pixels that
                                   saturated, i.e. stopped at T_next <= 1e-4), or NULL */
    /* two-level binning (b2s_coarse_bin): when bin_offsets != NULL the tile lists are expanded from the coarse bins
     * inside the kernel; isect_offsets / flatten_ids are not read and last_ids holds positions in bin_ids. */
    const int32_t* bin_offsets; /* [C*bins_h*bins_w + 1] or NULL */
    const uint32_t* bin_recs;   /* [n_isects] (n_isects = number of coarse records) */
    const int32_t* bin_ids;     /* [n_isects] */
    /* tile-row window (multi-GPU tile stripes): only tile rows [tile_row_lo, tile_row_hi) of the grid are composited
     * and the per-pixel tensors (render_colors, render_alphas, last_ids, final_T) hold just those rows:
     * [C, min(height, 16*tile_row_hi) - 16*tile_row_lo, W, ..].  0, 0 = the whole grid. */
    int32_t tile_row_lo, tile_row_hi;
} b2s_composite_fwd_args;
B2S_API int b2s_composite_fwd(const b2s_composite_fwd_args* a, b2s_stream_t stream);

/* Backward.  The gradient tensors are accumulated into with atomics: zero them first (b2s_zero_fill).  The element offset
 * of a gradient, row * width, is formed in 32 bits: rows * max(channels, 3) must be below 2^32 (the Python host checks it). */

typedef struct {
    int32_t C, width, height, tile_size, tile_w, tile_h, channels;
    int64_t n_isects;
    const float* means2d; const float* conics; const float* colors; const float* opacities;
    const float* backgrounds;
    const int32_t* isect_offsets; const int32_t* flatten_ids;
    const float* render_alphas; const int32_t* last_ids;
    const float* final_T;         /* from the forward, or NULL (then 1 - render_alphas is used) */
    const float* v_render_colors; /* [C,H,W,channels] */
    const float* v_render_alphas; /* [C,H,W,1] */
    /* gradients: ACCUMULATED with atomics, the caller zero-fills them first */
    float* v_means2d;     /* [M,2] */
    float* v_means2d_abs; /* [M,2] or NULL (absgrad) */
    float* v_conics;      /* [M,3] */
    float* v_colors;      /* [M,channels] */
    float* v_opacities;   /* [M] */
    const int32_t* bin_offsets; const uint32_t* bin_recs; const int32_t* bin_ids;   /* as in the forward */
    int32_t tile_row_lo, tile_row_hi; /* as in the forward: the per-pixel tensors hold the window's rows only */
} b2s_composite_bwd_args;
B2S_API int b2s_composite_bwd(const b2s_composite_bwd_args* a, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The step either side of rasterization() in the training loop.  The reference's
 * Train-Stage1 component runs gsplat's examples/simple_trainer.py (source/container/src/main.py:1328-1347):
 * loss = (1-lambda) * L1 + lambda * (1 - SSIM) (fused_ssim: 11x11 Gaussian window, sigma 1.5, "valid" padding,
 * C1 = 0.01^2, C2 = 0.03^2), optimizers = torch.optim.Adam per parameter group.
 *   b2s_l1_ssim_fwd : render / target [C,H,W,ch]; partial [b2s_l1_ssim_ctas(C,H,W)][2] = per-CTA sums of |x-y| and of
 *                     the SSIM map (the caller adds them up: deterministic); dmaps [C,H-10,W-10,ch,3] = partial
 *                     derivatives of the SSIM map w.r.t. the render's window moments (NULL: forward only)
 *   b2s_l1_ssim_bwd : v_render [C,H,W,ch] = w_l1 * sign(x-y) + w_ssim * correlate(dmaps, window)
 *   b2s_adam_step   : p, g, exp_avg, exp_avg_sq [rows, width]; row_mask (1 byte per row, NULL = all rows) skips rows
 *                     entirely (SelectiveAdam: rows invisible this step keep parameters AND moments); step counts from 1
 * ---------------------------------------------------------------------------------------------- */
B2S_API int64_t b2s_l1_ssim_ctas(int32_t C, int32_t H, int32_t W);
B2S_API int b2s_l1_ssim_fwd(int32_t C, int32_t H, int32_t W, int32_t channels, const float* render, const float* target,
                            float* dmaps, float* partial, b2s_stream_t stream);
B2S_API int b2s_l1_ssim_bwd(int32_t C, int32_t H, int32_t W, int32_t channels, const float* render, const float* target,
                            const float* dmaps, float w_l1, float w_ssim, float* v_render, b2s_stream_t stream);
B2S_API int b2s_adam_step(int64_t rows, int32_t width, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                          const uint8_t* row_mask, float lr, float beta1, float beta2, float eps, int64_t step,
                          b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SPZ packing: packGaussians() of the reference's vendored Niantic codec
 * (source/container/src/pipeline/post_processing/spz/src/cc/load-spz.cc:213-289), the quantisation behind the Spz-Export
 * component (`splat_converter splat.ply`, source/container/src/main.py:1532-1545).  Inputs are the checkpoint tensors
 * (log-scales, wxyz quaternions, opacity logits, sh0 [N,1,3], shN [N,shN_rows,3] of which the first sh_dim rows are
 * used); outputs are the codec's byte arrays: positions [9N], alphas [N], colors [3N], scales [3N], rotations [3N],
 * sh [3*sh_dim*N].  Byte-exact against the reference's converter (oracle/_ref/splat_converter).
 * ---------------------------------------------------------------------------------------------- */
B2S_API int b2s_spz_pack(int64_t N, int32_t sh_dim, const float* means, const float* log_scales, const float* quats_wxyz,
                         const float* opacity_logits, const float* sh0, const float* shN, int32_t shN_rows,
                         uint8_t* positions, uint8_t* alphas, uint8_t* colors, uint8_t* scales, uint8_t* rotations,
                         uint8_t* sh, b2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SPZ decoding: unpackGaussians() (load-spz.cc:373-440) fused with the row interleave of saveSplatToPly()
 * (load-spz.cc:752-790).  Inputs are the codec's byte arrays as b2s_spz_pack writes them (24-bit positions only: the
 * float16 layout of SPZ version 1 is rejected by the file-level call); output is the PLY vertex table
 * [N, 17 + 3*sh_dim] fp32 in the reference's column order x y z nx ny nz f_dc_0..2 f_rest_* (channel-major) opacity
 * scale_0..2 rot_0..3 (w first).  Bit-exact against the reference's converter.
 *
 * b2s_convert_spz_to_ply / ConvertSpzToPly: the reference's exported C entry point
 * (spz/src/spz_converter_api.cpp:7-30, same name and signature) and the `.spz` branch of `splat_converter`
 * (spz/src/main.cpp:29-35): HOST paths in, file out.  Reads + inflates on the host, decodes on the current device (the
 * only entry of this library that allocates device memory and synchronises), writes the PLY.  A missing or malformed
 * file returns an error (the reference returns false for a missing file and divides by zero on an empty cloud).
 * ---------------------------------------------------------------------------------------------- */
B2S_API int b2s_spz_unpack(int64_t N, int32_t sh_dim, int32_t fractional_bits, const uint8_t* positions, const uint8_t* alphas,
                           const uint8_t* colors, const uint8_t* scales, const uint8_t* rotations, const uint8_t* sh,
                           float* ply_rows, b2s_stream_t stream);
B2S_API int b2s_convert_spz_to_ply(const char* host_spz_path, const char* host_ply_path);
/* The other branch of `splat_converter` (spz/src/main.cpp:22-28): loadSplatFromPly (load-spz.cc:568-750: the same header
 * rules -- two fixed lines, `comment` lines skipped, float properties only, fields found by name, 1 .. 10 Mi vertices) +
 * saveSpz.  The interleaved vertex table is uploaded as it lies in the file; a gather kernel splits it into the codec's
 * arrays and b2s_spz_pack quantises them; gzip with the reference's parameters: the file is byte-identical to the
 * reference binary's.  `pipeline_pointcloud_b200/splat_converter` is the reference's command line on top of both calls. */
B2S_API int b2s_convert_ply_to_spz(const char* host_ply_path, const char* host_spz_path);
B2S_API bool ConvertSpzToPly(const char* spzFilePath, const char* plyFilePath);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H */
