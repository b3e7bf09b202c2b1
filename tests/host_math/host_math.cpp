// TEST TOOL, not a product path.  Compiles the host+device math header of the CUDA kernels
// (pipeline_pointcloud_b200/csrc/b2s_math.cuh) with g++ so that the per-element projection / SH math
// and their hand-derived VJPs can be checked against the torch oracle on a box without a GPU.
// Built by tests/host_math/build.py into tests/host_math/_host_math.so; only tests/ load it.
#include "b2s_math.cuh"
#include <cstring>
using namespace b2s;

extern "C" {

// forward projection of all (camera, Gaussian) pairs; outputs [C,N,...]; rects [C,N,4] = x0,y0,x1,y1
void hm_project_fwd(long N, int C, const float* means, const float* quats, const float* scales, const float* viewmats,
                    const float* Ks, int W, int H, float eps2d, float near_plane, float far_plane, float radius_clip,
                    int tile_size, int tile_w, int tile_h, int* radii, float* means2d, float* depths, float* conics,
                    float* comps, int* rects) {
    for (int c = 0; c < C; ++c) {
        Camera cam;
        camera_setup(cam, viewmats + 16 * c, Ks + 9 * c, nullptr, (float)W, (float)H);
        for (long n = 0; n < N; ++n) {
            Rot R; float qn[4], nrm, M[9]; Sym3 S;
            quat_to_rotmat(quats + 4 * n, R, qn, nrm);
            covar_from_rot_scale(R, scales + 3 * n, M, S);
            ProjOut o;
            project_one(cam, means + 3 * n, S, (float)W, (float)H, eps2d, near_plane, far_plane, radius_clip, o, nullptr);
            long i = (long)c * N + n;
            radii[i] = o.radius; means2d[2 * i] = o.x2d; means2d[2 * i + 1] = o.y2d; depths[i] = o.depth;
            conics[3 * i] = o.conic_a; conics[3 * i + 1] = o.conic_b; conics[3 * i + 2] = o.conic_c; comps[i] = o.comp;
            int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
            if (o.radius > 0) tile_rect(o.x2d, o.y2d, o.radius, (float)tile_size, tile_w, tile_h, x0, y0, x1, y1);
            rects[4 * i] = x0; rects[4 * i + 1] = y0; rects[4 * i + 2] = x1; rects[4 * i + 3] = y1;
        }
    }
}

// backward: cotangents [C,N,...] -> v_means [N,3], v_quats [N,4], v_scales [N,3]
void hm_project_bwd(long N, int C, const float* means, const float* quats, const float* scales, const float* viewmats,
                    const float* Ks, int W, int H, float eps2d, float near_plane, float far_plane, float radius_clip,
                    const float* v_means2d, const float* v_depths, const float* v_conics, const float* v_comps,
                    float* v_means, float* v_quats, float* v_scales) {
    for (long n = 0; n < N; ++n) {
        Rot R; float qn[4], nrm, M[9]; Sym3 S;
        quat_to_rotmat(quats + 4 * n, R, qn, nrm);
        covar_from_rot_scale(R, scales + 3 * n, M, S);
        float vm[3] = {0, 0, 0}; Sym3 vS = {0, 0, 0, 0, 0, 0};
        for (int c = 0; c < C; ++c) {
            Camera cam;
            camera_setup(cam, viewmats + 16 * c, Ks + 9 * c, nullptr, (float)W, (float)H);
            ProjOut o; ProjMid mid;
            if (!project_one(cam, means + 3 * n, S, (float)W, (float)H, eps2d, near_plane, far_plane, radius_clip, o, nullptr)) continue;
            project_mid_fast(cam, means + 3 * n, S, eps2d, mid);   // what the backward kernel uses
            long i = (long)c * N + n;
            ProjGradIn g = {v_means2d[2 * i], v_means2d[2 * i + 1], v_depths[i], v_conics[3 * i], v_conics[3 * i + 1],
                            v_conics[3 * i + 2], v_comps ? v_comps[i] : 0.f};
            project_one_bwd(cam, means + 3 * n, S, eps2d, mid, g, vm, vS);
        }
        float vq[4], vs[3];
        covar_bwd(quats + 4 * n, qn, nrm, R, scales + 3 * n, M, vS, vq, vs);
        for (int k = 0; k < 3; ++k) { v_means[3 * n + k] = vm[k]; v_scales[3 * n + k] = vs[k]; }
        for (int k = 0; k < 4; ++k) v_quats[4 * n + k] = vq[k];
    }
}

// SH colour for dirs [M,3] (un-normalised) and coeffs [M,Kt,3]:  out [M,3] (no +0.5, no clamp)
void hm_sh_fwd(long M, int deg, int Kt, const float* dirs, const float* coeffs, float* out) {
    int K = (deg + 1) * (deg + 1);
    for (long i = 0; i < M; ++i) {
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        float Y[16]; sh_basis(deg, x * inv, y * inv, z * inv, Y);
        float r = 0, g = 0, b = 0;
        for (int k = 0; k < K; ++k) {
            const float* c = coeffs + ((long)i * Kt + k) * 3;
            r += Y[k] * c[0]; g += Y[k] * c[1]; b += Y[k] * c[2];
        }
        out[3 * i] = r; out[3 * i + 1] = g; out[3 * i + 2] = b;
    }
}

// SH VJP: v_out [M,3] -> v_coeffs [M,Kt,3], v_dirs [M,3]
void hm_sh_bwd(long M, int deg, int Kt, const float* dirs, const float* coeffs, const float* v_out, float* v_coeffs,
               float* v_dirs) {
    int K = (deg + 1) * (deg + 1);
    memset(v_coeffs, 0, sizeof(float) * M * Kt * 3);
    for (long i = 0; i < M; ++i) {
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float nrm = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f), inv = 1.f / nrm;
        float ux = x * inv, uy = y * inv, uz = z * inv;
        float Y[16]; sh_basis(deg, ux, uy, uz, Y);
        float w[16];
        const float* vo = v_out + 3 * i;
        for (int k = 0; k < K; ++k) {
            const float* c = coeffs + ((long)i * Kt + k) * 3;
            float* vc = v_coeffs + ((long)i * Kt + k) * 3;
            vc[0] = Y[k] * vo[0]; vc[1] = Y[k] * vo[1]; vc[2] = Y[k] * vo[2];
            w[k] = c[0] * vo[0] + c[1] * vo[1] + c[2] * vo[2];
        }
        float vx, vy, vz; sh_basis_vjp(deg, ux, uy, uz, w, vx, vy, vz);
        float dp = ux * vx + uy * vy + uz * vz;
        v_dirs[3 * i] = (vx - ux * dp) * inv; v_dirs[3 * i + 1] = (vy - uy * dp) * inv; v_dirs[3 * i + 2] = (vz - uz * dp) * inv;
    }
}
}
