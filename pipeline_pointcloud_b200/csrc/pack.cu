// Packed layout support (gsplat `packed=True`): gathers the visible
// (camera, Gaussian) rows of the dense stage-1 outputs into nnz-row tensors, camera-major and
// Gaussian-ascending, and scatters their gradients back.  Row indices are unique, so the backward is a
// plain scatter into zero-filled dense gradients (no atomics, no sort).
#include "common.cuh"

namespace b2s {

__global__ void pack_rows_kernel(int64_t nnz, int64_t N, const int64_t* __restrict__ sel, int ch,
                                 const float2* __restrict__ means2d, const float* __restrict__ depths,
                                 const float* __restrict__ conics, const float* __restrict__ opac,
                                 const float* __restrict__ colors, const int32_t* __restrict__ radii,
                                 const int32_t* __restrict__ tiles, float2* __restrict__ means2d_p,
                                 float* __restrict__ depths_p, float* __restrict__ conics_p, float* __restrict__ opac_p,
                                 float* __restrict__ colors_p, int32_t* __restrict__ radii_p, int32_t* __restrict__ tiles_p,
                                 int64_t* __restrict__ camera_ids, int64_t* __restrict__ gaussian_ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t r = sel[i];
    if (means2d_p) means2d_p[i] = means2d[r];
    if (depths_p) depths_p[i] = depths[r];
    if (conics_p) { conics_p[3 * i] = conics[3 * r]; conics_p[3 * i + 1] = conics[3 * r + 1]; conics_p[3 * i + 2] = conics[3 * r + 2]; }
    if (opac_p) opac_p[i] = opac[r];
    if (colors_p) for (int k = 0; k < ch; ++k) colors_p[i * ch + k] = colors[r * ch + k];
    if (radii_p) radii_p[i] = radii[r];
    if (tiles_p) tiles_p[i] = tiles[r];
    if (camera_ids) { camera_ids[i] = r / N; gaussian_ids[i] = r % N; }
}

__global__ void unpack_rows_kernel(int64_t nnz, const int64_t* __restrict__ sel, int ch, const float2* __restrict__ g_means2d,
                                   const float* __restrict__ g_depths, const float* __restrict__ g_conics,
                                   const float* __restrict__ g_opac, const float* __restrict__ g_colors,
                                   float2* __restrict__ v_means2d, float* __restrict__ v_depths, float* __restrict__ v_conics,
                                   float* __restrict__ v_opac, float* __restrict__ v_colors) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t r = sel[i];
    if (g_means2d) v_means2d[r] = g_means2d[i];
    if (g_depths) v_depths[r] = g_depths[i];
    if (g_conics) { v_conics[3 * r] = g_conics[3 * i]; v_conics[3 * r + 1] = g_conics[3 * i + 1]; v_conics[3 * r + 2] = g_conics[3 * i + 2]; }
    if (g_opac) v_opac[r] = g_opac[i];
    if (g_colors) for (int k = 0; k < ch; ++k) v_colors[r * ch + k] = g_colors[i * ch + k];
}

// ---- multi-GPU routing: one fused record per visible splat, so one all-to-all moves everything ----------
// record = [x, y, depth, conic a b c, opacity, colour[ch], radius, tiles, local camera, global Gaussian id]
// (the four integers travel bit-cast in float slots); F = 11 + ch floats.
__global__ void pack_payload_kernel(int64_t nnz, int64_t N, const int64_t* __restrict__ sel, int ch, int cams_per_rank,
                                    int64_t gid_offset, const float2* __restrict__ means2d, const float* __restrict__ depths,
                                    const float* __restrict__ conics, const float* __restrict__ opac,
                                    const float* __restrict__ colors, const int32_t* __restrict__ radii,
                                    const int32_t* __restrict__ tiles, float* __restrict__ payload) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int64_t r = sel[i];
    const int F = 11 + ch;
    float* o = payload + i * F;
    const float2 m = means2d[r];
    o[0] = m.x; o[1] = m.y; o[2] = depths[r];
    o[3] = conics[3 * r]; o[4] = conics[3 * r + 1]; o[5] = conics[3 * r + 2];
    o[6] = opac[r];
    for (int k = 0; k < ch; ++k) o[7 + k] = colors[r * ch + k];
    const int64_t cam = r / N, gid = r % N;
    o[7 + ch] = __int_as_float(radii[r]);
    o[8 + ch] = __int_as_float(tiles[r]);
    o[9 + ch] = __int_as_float((int)(cam % cams_per_rank));
    o[10 + ch] = __int_as_float((int)(gid + gid_offset));
}

__global__ void split_payload_kernel(int64_t n, int ch, const float* __restrict__ payload, float2* __restrict__ means2d,
                                     float* __restrict__ depths, float* __restrict__ conics, float* __restrict__ opac,
                                     float* __restrict__ colors, int32_t* __restrict__ radii, int32_t* __restrict__ tiles,
                                     int64_t* __restrict__ camera_ids, int64_t* __restrict__ gaussian_ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int F = 11 + ch;
    const float* p = payload + i * F;
    means2d[i] = make_float2(p[0], p[1]);
    depths[i] = p[2];
    conics[3 * i] = p[3]; conics[3 * i + 1] = p[4]; conics[3 * i + 2] = p[5];
    opac[i] = p[6];
    for (int k = 0; k < ch; ++k) colors[i * ch + k] = p[7 + k];
    radii[i] = __float_as_int(p[7 + ch]);
    tiles[i] = __float_as_int(p[8 + ch]);
    camera_ids[i] = (int64_t)__float_as_int(p[9 + ch]);
    gaussian_ids[i] = (int64_t)__float_as_int(p[10 + ch]);
}

// gradients of the received rows -> one [n, 7+ch] record (NULL pieces count as zero)
__global__ void join_grads_kernel(int64_t n, int ch, const float2* __restrict__ g_means2d, const float* __restrict__ g_depths,
                                  const float* __restrict__ g_conics, const float* __restrict__ g_opac,
                                  const float* __restrict__ g_colors, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int G = 7 + ch;
    float* o = out + i * G;
    const float2 m = g_means2d ? g_means2d[i] : make_float2(0.f, 0.f);
    o[0] = m.x; o[1] = m.y;
    o[2] = g_depths ? g_depths[i] : 0.f;
    o[3] = g_conics ? g_conics[3 * i] : 0.f; o[4] = g_conics ? g_conics[3 * i + 1] : 0.f; o[5] = g_conics ? g_conics[3 * i + 2] : 0.f;
    o[6] = g_opac ? g_opac[i] : 0.f;
    for (int k = 0; k < ch; ++k) o[7 + k] = g_colors ? g_colors[i * ch + k] : 0.f;
}

// returned gradient records -> dense (zero-filled) gradients of the stage-1 outputs.  ADD: a row may occur several
// times in `sel` (tile stripes: a splat that straddles a stripe boundary was sent to every stripe it touches), so the
// records are summed with red.global.add instead of stored.
template <bool ADD>
__global__ void scatter_grads_kernel(int64_t nnz, const int64_t* __restrict__ sel, int ch, const float* __restrict__ in,
                                     float* __restrict__ v_means2d, float* __restrict__ v_depths, float* __restrict__ v_conics,
                                     float* __restrict__ v_opac, float* __restrict__ v_colors) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int G = 7 + ch;
    const float* p = in + i * G;
    const int64_t r = sel[i];
    auto put = [](float* dst, float v) {
        if (ADD) atomicAdd(dst, v);
        else *dst = v;
    };
    put(v_means2d + 2 * r, p[0]); put(v_means2d + 2 * r + 1, p[1]);
    put(v_depths + r, p[2]);
    put(v_conics + 3 * r, p[3]); put(v_conics + 3 * r + 1, p[4]); put(v_conics + 3 * r + 2, p[5]);
    put(v_opac + r, p[6]);
    for (int k = 0; k < ch; ++k) put(v_colors + r * ch + k, p[7 + k]);
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_pack_payload(int64_t nnz, int64_t N, const int64_t* sel, int32_t channels, int32_t cams_per_rank,
                                int64_t gid_offset, const float* means2d, const float* depths, const float* conics,
                                const float* opacities, const float* colors, const int32_t* radii, const int32_t* tiles,
                                float* payload, b2s_stream_t stream) {
    B2S_REQUIRE(nnz >= 0 && N > 0 && channels >= 0 && cams_per_rank > 0, "bad sizes");
    if (nnz == 0) return 0;
    B2S_REQUIRE(sel && means2d && depths && conics && opacities && colors && radii && tiles && payload, "null pointer");
    pack_payload_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
        nnz, N, sel, channels, cams_per_rank, gid_offset, reinterpret_cast<const float2*>(means2d), depths, conics, opacities,
        colors, radii, tiles, payload);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_split_payload(int64_t n, int32_t channels, const float* payload, float* means2d, float* depths,
                                 float* conics, float* opacities, float* colors, int32_t* radii, int32_t* tiles,
                                 int64_t* camera_ids, int64_t* gaussian_ids, b2s_stream_t stream) {
    B2S_REQUIRE(n >= 0 && channels >= 0, "bad sizes");
    if (n == 0) return 0;
    B2S_REQUIRE(payload && means2d && depths && conics && opacities && colors && radii && tiles && camera_ids && gaussian_ids,
                "null pointer");
    split_payload_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, channels, payload, reinterpret_cast<float2*>(means2d), depths, conics, opacities, colors, radii, tiles, camera_ids,
        gaussian_ids);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_join_grads(int64_t n, int32_t channels, const float* g_means2d, const float* g_depths,
                              const float* g_conics, const float* g_opacities, const float* g_colors, float* out,
                              b2s_stream_t stream) {
    B2S_REQUIRE(n >= 0 && channels >= 0, "bad sizes");
    if (n == 0) return 0;
    B2S_REQUIRE(out != nullptr, "null pointer");
    join_grads_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, channels, reinterpret_cast<const float2*>(g_means2d), g_depths, g_conics, g_opacities, g_colors, out);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_scatter_grads(int64_t nnz, const int64_t* sel, int32_t channels, const float* in, float* v_means2d,
                                 float* v_depths, float* v_conics, float* v_opacities, float* v_colors, int32_t accumulate,
                                 b2s_stream_t stream) {
    B2S_REQUIRE(nnz >= 0 && channels >= 0, "bad sizes");
    if (nnz == 0) return 0;
    B2S_REQUIRE(sel && in && v_means2d && v_depths && v_conics && v_opacities && v_colors, "null pointer");
    if (accumulate)
        scatter_grads_kernel<true><<<(unsigned)cdiv(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
            nnz, sel, channels, in, v_means2d, v_depths, v_conics, v_opacities, v_colors);
    else
        scatter_grads_kernel<false><<<(unsigned)cdiv(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
            nnz, sel, channels, in, v_means2d, v_depths, v_conics, v_opacities, v_colors);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_pack_rows(int64_t nnz, int64_t N, const int64_t* sel, int32_t channels, const float* means2d,
                             const float* depths, const float* conics, const float* opacities, const float* colors,
                             const int32_t* radii, const int32_t* tiles, float* means2d_p, float* depths_p, float* conics_p,
                             float* opacities_p, float* colors_p, int32_t* radii_p, int32_t* tiles_p, int64_t* camera_ids,
                             int64_t* gaussian_ids, b2s_stream_t stream) {
    B2S_REQUIRE(nnz >= 0 && N > 0 && channels >= 0, "bad sizes");
    if (nnz == 0) return 0;
    B2S_REQUIRE(sel != nullptr, "null sel");
    B2S_REQUIRE((camera_ids == nullptr) == (gaussian_ids == nullptr), "camera_ids and gaussian_ids go together");
    pack_rows_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
        nnz, N, sel, channels, reinterpret_cast<const float2*>(means2d), depths, conics, opacities, colors, radii, tiles,
        reinterpret_cast<float2*>(means2d_p), depths_p, conics_p, opacities_p, colors_p, radii_p, tiles_p, camera_ids,
        gaussian_ids);
    B2S_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2s_unpack_rows(int64_t nnz, const int64_t* sel, int32_t channels, const float* g_means2d,
                               const float* g_depths, const float* g_conics, const float* g_opacities, const float* g_colors,
                               float* v_means2d, float* v_depths, float* v_conics, float* v_opacities, float* v_colors,
                               b2s_stream_t stream) {
    B2S_REQUIRE(nnz >= 0 && channels >= 0, "bad sizes");
    if (nnz == 0) return 0;
    B2S_REQUIRE(sel != nullptr, "null sel");
    unpack_rows_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
        nnz, sel, channels, reinterpret_cast<const float2*>(g_means2d), g_depths, g_conics, g_opacities, g_colors,
        reinterpret_cast<float2*>(v_means2d), v_depths, v_conics, v_opacities, v_colors);
    B2S_LAUNCH_CHECK();
    return 0;
}
