// Per-element math of the rasterization() hot path: EWA projection (+VJP) and SH colour (+VJP).
//
// Everything here is `B2S_HD` (host + device) so that tests/ can compile the very same
// functions with g++ and check them against the torch oracle on a box without a GPU
// (tests/host_math/).  The product only ever runs them inside the sm_100a kernels.
//
// Reference semantics: gsplat 1.4 `fully_fused_projection`, `spherical_harmonics`
// (Appendix A), reached from the reference at
// source/container/src/main.py:1308-1316 and :1339-1347.
//
// Canonical order.  The forward projection feeds the *binning* (tile rectangles, sort keys),
// which must be bit-exact against the oracle, so every operation on that path is a single
// IEEE-754 round-to-nearest fp32 operation in a fixed order (RN_* below: never contracted to
// FMA).  oracle/torch_oracle.py::fully_fused_projection spells out the same sequence.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define RN_MUL(a, b) __fmul_rn((a), (b))
#define RN_ADD(a, b) __fadd_rn((a), (b))
#define RN_SUB(a, b) __fsub_rn((a), (b))
#define RN_DIV(a, b) __fdiv_rn((a), (b))
#define RN_SQRT(a) __fsqrt_rn((a))
#else  // host build: compile with -ffp-contract=off -fno-fast-math
#define RN_MUL(a, b) ((a) * (b))
#define RN_ADD(a, b) ((a) + (b))
#define RN_SUB(a, b) ((a) - (b))
#define RN_DIV(a, b) ((a) / (b))
#define RN_SQRT(a) sqrtf((a))
#endif

namespace b2s {

// ---------------------------------------------------------------------------------------------
// camera
// ---------------------------------------------------------------------------------------------
struct Camera {
    float R[9];  // world->camera rotation, row-major
    float t[3];  // world->camera translation
    float fx, fy, cx, cy;
    float lim_xp, lim_xn, lim_yp, lim_yn;  // clamp limits for x/z, y/z
    float campos[3];                       // camera centre in world coords (for SH view dirs)
};

// camera centre = translation column of inverse(viewmat) = -A^-1 t for the affine [A | t]
// (cofactor inverse, so a non-orthonormal A is handled like torch.linalg.inv does)
B2S_HD void campos_from_viewmat(const float* V, float out[3]) {
    const float R[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    const float t[3] = {V[3], V[7], V[11]};
    float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[2] * R[7] - R[1] * R[8], c02 = R[1] * R[5] - R[2] * R[4];
    float c10 = R[5] * R[6] - R[3] * R[8], c11 = R[0] * R[8] - R[2] * R[6], c12 = R[2] * R[3] - R[0] * R[5];
    float c20 = R[3] * R[7] - R[4] * R[6], c21 = R[1] * R[6] - R[0] * R[7], c22 = R[0] * R[4] - R[1] * R[3];
    float det = R[0] * c00 + R[1] * c10 + R[2] * c20;
    float id = 1.f / det;
    out[0] = -(c00 * t[0] + c01 * t[1] + c02 * t[2]) * id;
    out[1] = -(c10 * t[0] + c11 * t[1] + c12 * t[2]) * id;
    out[2] = -(c20 * t[0] + c21 * t[1] + c22 * t[2]) * id;
}

// viewmat: 16 floats row-major [4,4]; K: 9 floats row-major [3,3]
B2S_HD void camera_setup(Camera& c, const float* viewmat, const float* K, const float* campos, float W, float H) {
    c.R[0] = viewmat[0]; c.R[1] = viewmat[1]; c.R[2] = viewmat[2];  c.t[0] = viewmat[3];
    c.R[3] = viewmat[4]; c.R[4] = viewmat[5]; c.R[5] = viewmat[6];  c.t[1] = viewmat[7];
    c.R[6] = viewmat[8]; c.R[7] = viewmat[9]; c.R[8] = viewmat[10]; c.t[2] = viewmat[11];
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    float tan_fovx = RN_DIV(RN_MUL(0.5f, W), c.fx);
    float tan_fovy = RN_DIV(RN_MUL(0.5f, H), c.fy);
    c.lim_xp = RN_ADD(RN_DIV(RN_SUB(W, c.cx), c.fx), RN_MUL(0.3f, tan_fovx));
    c.lim_xn = RN_ADD(RN_DIV(c.cx, c.fx), RN_MUL(0.3f, tan_fovx));
    c.lim_yp = RN_ADD(RN_DIV(RN_SUB(H, c.cy), c.fy), RN_MUL(0.3f, tan_fovy));
    c.lim_yn = RN_ADD(RN_DIV(c.cy, c.fy), RN_MUL(0.3f, tan_fovy));
    if (campos) { c.campos[0] = campos[0]; c.campos[1] = campos[1]; c.campos[2] = campos[2]; }
    else campos_from_viewmat(viewmat, c.campos);
}

// ---------------------------------------------------------------------------------------------
// quaternion / scale -> rotation, covariance   
// ---------------------------------------------------------------------------------------------
struct Rot { float r[9]; };
struct Sym3 { float c00, c01, c02, c11, c12, c22; };

// q = (w,x,y,z) un-normalised.  Returns R and the normalised quaternion + 1/|q|.
B2S_HD void quat_to_rotmat(const float q[4], Rot& R, float qn[4], float& inv_norm_den) {
    float n2 = RN_ADD(RN_ADD(RN_ADD(RN_MUL(q[0], q[0]), RN_MUL(q[1], q[1])), RN_MUL(q[2], q[2])), RN_MUL(q[3], q[3]));
    float n = fmaxf(RN_SQRT(n2), 1e-12f);
    inv_norm_den = n;
    float w = RN_DIV(q[0], n), x = RN_DIV(q[1], n), y = RN_DIV(q[2], n), z = RN_DIV(q[3], n);
    qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
    float x2 = RN_MUL(x, x), y2 = RN_MUL(y, y), z2 = RN_MUL(z, z);
    float xy = RN_MUL(x, y), xz = RN_MUL(x, z), yz = RN_MUL(y, z);
    float wx = RN_MUL(w, x), wy = RN_MUL(w, y), wz = RN_MUL(w, z);
    R.r[0] = RN_SUB(1.f, RN_MUL(2.f, RN_ADD(y2, z2)));
    R.r[1] = RN_MUL(2.f, RN_SUB(xy, wz));
    R.r[2] = RN_MUL(2.f, RN_ADD(xz, wy));
    R.r[3] = RN_MUL(2.f, RN_ADD(xy, wz));
    R.r[4] = RN_SUB(1.f, RN_MUL(2.f, RN_ADD(x2, z2)));
    R.r[5] = RN_MUL(2.f, RN_SUB(yz, wx));
    R.r[6] = RN_MUL(2.f, RN_SUB(xz, wy));
    R.r[7] = RN_MUL(2.f, RN_ADD(yz, wx));
    R.r[8] = RN_SUB(1.f, RN_MUL(2.f, RN_ADD(x2, y2)));
}

#define B2S_DOT3(a0, b0, a1, b1, a2, b2) RN_ADD(RN_ADD(RN_MUL(a0, b0), RN_MUL(a1, b1)), RN_MUL(a2, b2))

// M = R diag(s);  Sigma = M M^T
B2S_HD void covar_from_rot_scale(const Rot& R, const float s[3], float M[9], Sym3& S) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = RN_MUL(R.r[3 * i + j], s[j]);
    S.c00 = B2S_DOT3(M[0], M[0], M[1], M[1], M[2], M[2]);
    S.c01 = B2S_DOT3(M[0], M[3], M[1], M[4], M[2], M[5]);
    S.c02 = B2S_DOT3(M[0], M[6], M[1], M[7], M[2], M[8]);
    S.c11 = B2S_DOT3(M[3], M[3], M[4], M[4], M[5], M[5]);
    S.c12 = B2S_DOT3(M[3], M[6], M[4], M[7], M[5], M[8]);
    S.c22 = B2S_DOT3(M[6], M[6], M[7], M[7], M[8], M[8]);
}

// ---------------------------------------------------------------------------------------------
// forward projection of one (camera, Gaussian) pair   
// ---------------------------------------------------------------------------------------------
struct ProjOut {
    int radius;     // 0 => culled
    float x2d, y2d, depth;
    float conic_a, conic_b, conic_c;
    float comp;     // sqrt(max(0, det_orig/det))
};

// intermediate values the backward re-derives
struct ProjMid {
    float tx, ty, tz;          // camera-space mean
    float txc, tyc;            // clamped
    bool clamp_x, clamp_y;     // clamped => txc = tz*lim (sign kept in lim_used)
    float limx_used, limy_used;
    float s00, s01, s02, s11, s12, s22;  // camera-space covariance
    float j00, j02, j11, j12;
    float q00, q01, q11;       // 2-D covariance before blur
    float a, d, det, det_orig;
};

B2S_HD bool project_one(const Camera& cam, const float mean[3], const Sym3& S, float W, float H, float eps2d,
                        float near_plane, float far_plane, float radius_clip, ProjOut& o, ProjMid* mid) {
    const float* R = cam.R;
    float tx = RN_ADD(B2S_DOT3(R[0], mean[0], R[1], mean[1], R[2], mean[2]), cam.t[0]);
    float ty = RN_ADD(B2S_DOT3(R[3], mean[0], R[4], mean[1], R[5], mean[2]), cam.t[1]);
    float tz = RN_ADD(B2S_DOT3(R[6], mean[0], R[7], mean[1], R[8], mean[2]), cam.t[2]);
    o.radius = 0; o.x2d = o.y2d = o.depth = 0.f; o.conic_a = o.conic_b = o.conic_c = 0.f; o.comp = 0.f;
    if (!(tz > near_plane && tz < far_plane)) return false;

    // A = R Sigma ; Sc = A R^T
    float Sg[9] = {S.c00, S.c01, S.c02, S.c01, S.c11, S.c12, S.c02, S.c12, S.c22};
    float A[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[3 * i + j] = B2S_DOT3(R[3 * i], Sg[j], R[3 * i + 1], Sg[3 + j], R[3 * i + 2], Sg[6 + j]);
    float s00 = B2S_DOT3(A[0], R[0], A[1], R[1], A[2], R[2]);
    float s01 = B2S_DOT3(A[0], R[3], A[1], R[4], A[2], R[5]);
    float s02 = B2S_DOT3(A[0], R[6], A[1], R[7], A[2], R[8]);
    float s11 = B2S_DOT3(A[3], R[3], A[4], R[4], A[5], R[5]);
    float s12 = B2S_DOT3(A[3], R[6], A[4], R[7], A[5], R[8]);
    float s22 = B2S_DOT3(A[6], R[6], A[7], R[7], A[8], R[8]);

    float txz = RN_DIV(tx, tz), tyz = RN_DIV(ty, tz);
    float cxr = fminf(cam.lim_xp, fmaxf(-cam.lim_xn, txz));
    float cyr = fminf(cam.lim_yp, fmaxf(-cam.lim_yn, tyz));
    float txc = RN_MUL(tz, cxr), tyc = RN_MUL(tz, cyr);
    float tz2 = RN_MUL(tz, tz);
    float j00 = RN_DIV(cam.fx, tz), j11 = RN_DIV(cam.fy, tz);
    float j02 = -RN_DIV(RN_MUL(cam.fx, txc), tz2);
    float j12 = -RN_DIV(RN_MUL(cam.fy, tyc), tz2);
    float b00 = RN_ADD(RN_MUL(j00, s00), RN_MUL(j02, s02));
    float b01 = RN_ADD(RN_MUL(j00, s01), RN_MUL(j02, s12));
    float b02 = RN_ADD(RN_MUL(j00, s02), RN_MUL(j02, s22));
    float b11 = RN_ADD(RN_MUL(j11, s11), RN_MUL(j12, s12));
    float b12 = RN_ADD(RN_MUL(j11, s12), RN_MUL(j12, s22));
    float q00 = RN_ADD(RN_MUL(b00, j00), RN_MUL(b02, j02));
    float q01 = RN_ADD(RN_MUL(b01, j11), RN_MUL(b02, j12));
    float q11 = RN_ADD(RN_MUL(b11, j11), RN_MUL(b12, j12));
    float x2d = RN_ADD(RN_MUL(cam.fx, txz), cam.cx);
    float y2d = RN_ADD(RN_MUL(cam.fy, tyz), cam.cy);

    float det_orig = RN_SUB(RN_MUL(q00, q11), RN_MUL(q01, q01));
    float a = RN_ADD(q00, eps2d), d = RN_ADD(q11, eps2d);
    float det = RN_SUB(RN_MUL(a, d), RN_MUL(q01, q01));
    if (mid) {
        mid->tx = tx; mid->ty = ty; mid->tz = tz; mid->txc = txc; mid->tyc = tyc;
        mid->clamp_x = (cxr != txz); mid->clamp_y = (cyr != tyz);
        mid->limx_used = cxr; mid->limy_used = cyr;
        mid->s00 = s00; mid->s01 = s01; mid->s02 = s02; mid->s11 = s11; mid->s12 = s12; mid->s22 = s22;
        mid->j00 = j00; mid->j02 = j02; mid->j11 = j11; mid->j12 = j12;
        mid->q00 = q00; mid->q01 = q01; mid->q11 = q11; mid->a = a; mid->d = d; mid->det = det; mid->det_orig = det_orig;
    }
    if (!(det > 0.f)) return false;
    float detc = fmaxf(det, 1e-10f);
    float midv = RN_MUL(0.5f, RN_ADD(a, d));
    float lam = RN_ADD(midv, RN_SQRT(fmaxf(RN_SUB(RN_MUL(midv, midv), det), 0.01f)));
    float radius = ceilf(RN_MUL(3.f, RN_SQRT(lam)));
    if (!(radius > radius_clip)) return false;
    if (!(RN_ADD(x2d, radius) > 0.f && RN_SUB(x2d, radius) < W && RN_ADD(y2d, radius) > 0.f && RN_SUB(y2d, radius) < H))
        return false;
    o.radius = (int)radius;
    o.x2d = x2d; o.y2d = y2d; o.depth = tz;
    o.conic_a = RN_DIV(d, detc);
    o.conic_b = -RN_DIV(q01, detc);
    o.conic_c = RN_DIV(a, detc);
    o.comp = RN_SQRT(fmaxf(RN_DIV(det_orig, detc), 0.f));
    return true;
}

// The backward only needs the intermediates (never the cull decision, which it takes from the forward's
// radii), so it recomputes them with ordinary contracted arithmetic: ~40 % fewer instructions than the
// canonical-order forward.  Differences from project_one are at the 1e-7 level.
B2S_HD void project_mid_fast(const Camera& cam, const float mean[3], const Sym3& S, float eps2d, ProjMid& m) {
    const float* R = cam.R;
    const float tx = R[0] * mean[0] + R[1] * mean[1] + R[2] * mean[2] + cam.t[0];
    const float ty = R[3] * mean[0] + R[4] * mean[1] + R[5] * mean[2] + cam.t[1];
    const float tz = R[6] * mean[0] + R[7] * mean[1] + R[8] * mean[2] + cam.t[2];
    // A = R Sigma ; Sc = A R^T
    const float a00 = R[0] * S.c00 + R[1] * S.c01 + R[2] * S.c02, a01 = R[0] * S.c01 + R[1] * S.c11 + R[2] * S.c12,
                a02 = R[0] * S.c02 + R[1] * S.c12 + R[2] * S.c22;
    const float a10 = R[3] * S.c00 + R[4] * S.c01 + R[5] * S.c02, a11 = R[3] * S.c01 + R[4] * S.c11 + R[5] * S.c12,
                a12 = R[3] * S.c02 + R[4] * S.c12 + R[5] * S.c22;
    const float a20 = R[6] * S.c00 + R[7] * S.c01 + R[8] * S.c02, a21 = R[6] * S.c01 + R[7] * S.c11 + R[8] * S.c12,
                a22 = R[6] * S.c02 + R[7] * S.c12 + R[8] * S.c22;
    m.s00 = a00 * R[0] + a01 * R[1] + a02 * R[2];
    m.s01 = a00 * R[3] + a01 * R[4] + a02 * R[5];
    m.s02 = a00 * R[6] + a01 * R[7] + a02 * R[8];
    m.s11 = a10 * R[3] + a11 * R[4] + a12 * R[5];
    m.s12 = a10 * R[6] + a11 * R[7] + a12 * R[8];
    m.s22 = a20 * R[6] + a21 * R[7] + a22 * R[8];
    const float rz = 1.f / tz;
    const float txz = tx * rz, tyz = ty * rz;
    const float cxr = fminf(cam.lim_xp, fmaxf(-cam.lim_xn, txz));
    const float cyr = fminf(cam.lim_yp, fmaxf(-cam.lim_yn, tyz));
    m.tx = tx; m.ty = ty; m.tz = tz;
    m.txc = tz * cxr; m.tyc = tz * cyr;
    m.clamp_x = (cxr != txz); m.clamp_y = (cyr != tyz);
    m.limx_used = cxr; m.limy_used = cyr;
    m.j00 = cam.fx * rz; m.j11 = cam.fy * rz;
    m.j02 = -cam.fx * m.txc * rz * rz;
    m.j12 = -cam.fy * m.tyc * rz * rz;
    const float b00 = m.j00 * m.s00 + m.j02 * m.s02, b01 = m.j00 * m.s01 + m.j02 * m.s12, b02 = m.j00 * m.s02 + m.j02 * m.s22;
    const float b11 = m.j11 * m.s11 + m.j12 * m.s12, b12 = m.j11 * m.s12 + m.j12 * m.s22;
    m.q00 = b00 * m.j00 + b02 * m.j02;
    m.q01 = b01 * m.j11 + b02 * m.j12;
    m.q11 = b11 * m.j11 + b12 * m.j12;
    m.det_orig = m.q00 * m.q11 - m.q01 * m.q01;
    m.a = m.q00 + eps2d; m.d = m.q11 + eps2d;
    m.det = m.a * m.d - m.q01 * m.q01;
}

B2S_HD int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Tile rectangle [min,max) of a projected Gaussian   (SURVEY 8a row 4), rows clipped to the window
// [row_lo, row_hi) of the tile grid (the whole grid: 0, tile_h; a tile stripe of a multi-GPU job otherwise)
B2S_HD void tile_rect(float x2d, float y2d, int radius, float tile_size, int tile_w, int row_lo, int row_hi, int& x0, int& y0,
                      int& x1, int& y1) {
    float mx = RN_DIV(x2d, tile_size), my = RN_DIV(y2d, tile_size), r = RN_DIV((float)radius, tile_size);
    int ax = (int)floorf(RN_SUB(mx, r)), ay = (int)floorf(RN_SUB(my, r));
    int bx = (int)ceilf(RN_ADD(mx, r)), by = (int)ceilf(RN_ADD(my, r));
    x0 = iclamp(ax, 0, tile_w); y0 = iclamp(ay, row_lo, row_hi);
    x1 = iclamp(bx, 0, tile_w); y1 = iclamp(by, row_lo, row_hi);
}
B2S_HD void tile_rect(float x2d, float y2d, int radius, float tile_size, int tile_w, int tile_h, int& x0, int& y0, int& x1,
                      int& y1) {
    tile_rect(x2d, y2d, radius, tile_size, tile_w, 0, tile_h, x0, y0, x1, y1);
}

// ---------------------------------------------------------------------------------------------
// backward of project_one: accumulates into v_mean[3] and v_Sigma (world covariance, Sym3 as the
// gradient w.r.t. the 6 unique entries, off-diagonals counting both symmetric positions).
// ---------------------------------------------------------------------------------------------
struct ProjGradIn {
    float v_x2d, v_y2d, v_depth;
    float v_conic_a, v_conic_b, v_conic_c;
    float v_comp;
};

// v_cam (optional, 12 floats, accumulated): gradient w.r.t. the camera's world->camera rotation (row-major 3x3) and
// translation -- the `v_viewmats` of gsplat's projection backward (t = R mean + t_cam; Sigma_cam = R Sigma R^T).
B2S_HD void project_one_bwd(const Camera& cam, const float mean[3], const Sym3& S, float eps2d, const ProjMid& m,
                            const ProjGradIn& g, float v_mean[3], Sym3& v_S, float* v_cam = nullptr) {
    const float* R = cam.R;
    // ---- conic = inverse([[a,b],[b,d]]) -> v_a, v_b (total for the off-diagonal), v_d
    float det = m.det;
    float inv_det = 1.f / det;
    float A = m.d * inv_det, B = -m.q01 * inv_det, Cc = m.a * inv_det;   // conic entries
    // G = [[vA, vB/2],[vB/2, vC]] ; v_Sigma2 = -Q G Q
    float g00 = g.v_conic_a, g01 = 0.5f * g.v_conic_b, g11 = g.v_conic_c;
    float h00 = A * g00 + B * g01, h01 = A * g01 + B * g11;   // H = Q G
    float h10 = B * g00 + Cc * g01, h11 = B * g01 + Cc * g11;
    float v_q00 = -(h00 * A + h01 * B);
    float v_q01 = -((h00 * B + h01 * Cc) + (h10 * A + h11 * B));
    float v_q11 = -(h10 * B + h11 * Cc);
    // ---- compensation = sqrt(max(0, det_orig/det))
    if (g.v_comp != 0.f) {
        float r = m.det_orig * inv_det;
        if (r > 0.f) {
            float comp = sqrtf(r);
            float v_r = g.v_comp * 0.5f / comp;
            float v_det_orig = v_r * inv_det;
            float v_det = -v_r * r * inv_det;
            v_q00 += v_det_orig * m.q11 + v_det * m.d;
            v_q11 += v_det_orig * m.q00 + v_det * m.a;
            v_q01 += -2.f * m.q01 * (v_det_orig + v_det);
        }
    }
    // ---- Sigma2 = J Sc J^T, J = [[j00,0,j02],[0,j11,j12]]
    float G00 = v_q00, G01 = 0.5f * v_q01, G11 = v_q11;
    // v_Sc = J^T G J  (symmetric 3x3, as gradient w.r.t. the full matrix)
    float j00 = m.j00, j02 = m.j02, j11 = m.j11, j12 = m.j12;
    float vs00 = j00 * G00 * j00;
    float vs01 = j00 * G01 * j11;
    float vs02 = j00 * (G00 * j02 + G01 * j12);
    float vs11 = j11 * G11 * j11;
    float vs12 = j11 * (G01 * j02 + G11 * j12);
    float vs22 = j02 * (G00 * j02 + G01 * j12) + j12 * (G01 * j02 + G11 * j12);
    // v_J = 2 G J Sc   (2x3)
    float js00 = j00 * m.s00 + j02 * m.s02, js01 = j00 * m.s01 + j02 * m.s12, js02 = j00 * m.s02 + j02 * m.s22;  // (J Sc) row 0
    float js10 = j11 * m.s01 + j12 * m.s02, js11 = j11 * m.s11 + j12 * m.s12, js12 = j11 * m.s12 + j12 * m.s22;  // row 1
    float vj00 = 2.f * (G00 * js00 + G01 * js10);
    float vj02 = 2.f * (G00 * js02 + G01 * js12);
    float vj11 = 2.f * (G01 * js01 + G11 * js11);
    float vj12 = 2.f * (G01 * js02 + G11 * js12);
    // ---- J(tz, txc, tyc)
    float rz = 1.f / m.tz, rz2 = rz * rz, rz3 = rz2 * rz;
    float v_tx = 0.f, v_ty = 0.f;
    float v_tz = -cam.fx * rz2 * vj00 - cam.fy * rz2 * vj11 + 2.f * cam.fx * m.txc * rz3 * vj02 + 2.f * cam.fy * m.tyc * rz3 * vj12;
    float v_txc = -cam.fx * rz2 * vj02, v_tyc = -cam.fy * rz2 * vj12;
    if (m.clamp_x) v_tz += v_txc * m.limx_used; else v_tx += v_txc;
    if (m.clamp_y) v_tz += v_tyc * m.limy_used; else v_ty += v_tyc;
    // ---- means2d = (fx tx/tz + cx, fy ty/tz + cy), depth = tz
    v_tx += cam.fx * rz * g.v_x2d;
    v_ty += cam.fy * rz * g.v_y2d;
    v_tz += -(cam.fx * m.tx * g.v_x2d + cam.fy * m.ty * g.v_y2d) * rz2 + g.v_depth;
    // ---- t = R mean + t
    v_mean[0] += R[0] * v_tx + R[3] * v_ty + R[6] * v_tz;
    v_mean[1] += R[1] * v_tx + R[4] * v_ty + R[7] * v_tz;
    v_mean[2] += R[2] * v_tx + R[5] * v_ty + R[8] * v_tz;
    // ---- Sc = R Sigma R^T  ->  v_Sigma = R^T v_Sc R   (full symmetric), then fold to 6 unique
    float VS[9] = {vs00, vs01, vs02, vs01, vs11, vs12, vs02, vs12, vs22};
    float T[9];  // T = R^T VS
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = R[i] * VS[j] + R[3 + i] * VS[3 + j] + R[6 + i] * VS[6 + j];
    float f00 = T[0] * R[0] + T[1] * R[3] + T[2] * R[6];
    float f01 = T[0] * R[1] + T[1] * R[4] + T[2] * R[7];
    float f02 = T[0] * R[2] + T[1] * R[5] + T[2] * R[8];
    float f11 = T[3] * R[1] + T[4] * R[4] + T[5] * R[7];
    float f12 = T[3] * R[2] + T[4] * R[5] + T[5] * R[8];
    float f22 = T[6] * R[2] + T[7] * R[5] + T[8] * R[8];
    v_S.c00 += f00; v_S.c11 += f11; v_S.c22 += f22;
    v_S.c01 += 2.f * f01; v_S.c02 += 2.f * f02; v_S.c12 += 2.f * f12;
    if (v_cam) {
        // t = R mean + t_cam
        v_cam[9] += v_tx; v_cam[10] += v_ty; v_cam[11] += v_tz;
        for (int j = 0; j < 3; ++j) { v_cam[j] += v_tx * mean[j]; v_cam[3 + j] += v_ty * mean[j]; v_cam[6 + j] += v_tz * mean[j]; }
        // Sigma_cam = R Sigma R^T  ->  v_R += (VS + VS^T) (R Sigma) = 2 VS (R Sigma)
        const float Sg[9] = {S.c00, S.c01, S.c02, S.c01, S.c11, S.c12, S.c02, S.c12, S.c22};
        float A[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) A[3 * i + j] = R[3 * i] * Sg[j] + R[3 * i + 1] * Sg[3 + j] + R[3 * i + 2] * Sg[6 + j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                v_cam[3 * i + j] += 2.f * (VS[3 * i] * A[j] + VS[3 * i + 1] * A[3 + j] + VS[3 * i + 2] * A[6 + j]);
    }
    (void)eps2d;
}

// Sigma = M M^T, M = R(q) diag(s):  v_Sigma (6 unique, off-diagonals already doubled) -> v_quat, v_scale
B2S_HD void covar_bwd(const float q_raw[4], const float qn[4], float norm, const Rot& R, const float s[3], const float M[9],
                      const Sym3& v_S, float v_quat[4], float v_scale[3]) {
    // full symmetric gradient matrix G (each off-diagonal gets half of the folded gradient)
    float G[9] = {v_S.c00, 0.5f * v_S.c01, 0.5f * v_S.c02, 0.5f * v_S.c01, v_S.c11, 0.5f * v_S.c12,
                  0.5f * v_S.c02, 0.5f * v_S.c12, v_S.c22};
    // v_M = 2 G M
    float vM[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) vM[3 * i + j] = 2.f * (G[3 * i] * M[j] + G[3 * i + 1] * M[3 + j] + G[3 * i + 2] * M[6 + j]);
    // M_ij = R_ij s_j
    float vR[9];
    for (int j = 0; j < 3; ++j) {
        v_scale[j] = R.r[j] * vM[j] + R.r[3 + j] * vM[3 + j] + R.r[6 + j] * vM[6 + j];
        for (int i = 0; i < 3; ++i) vR[3 * i + j] = vM[3 * i + j] * s[j];
    }
    float w = qn[0], x = qn[1], y = qn[2], z = qn[3];
    float vw = 2.f * (x * (vR[7] - vR[5]) + y * (vR[2] - vR[6]) + z * (vR[3] - vR[1]));
    float vx = 2.f * (-2.f * x * (vR[4] + vR[8]) + y * (vR[1] + vR[3]) + z * (vR[2] + vR[6]) + w * (vR[7] - vR[5]));
    float vy = 2.f * (x * (vR[1] + vR[3]) - 2.f * y * (vR[0] + vR[8]) + z * (vR[5] + vR[7]) + w * (vR[2] - vR[6]));
    float vz = 2.f * (x * (vR[2] + vR[6]) + y * (vR[5] + vR[7]) - 2.f * z * (vR[0] + vR[4]) + w * (vR[3] - vR[1]));
    // q_n = q / |q|
    float dotp = w * vw + x * vx + y * vy + z * vz;
    float inv = 1.f / norm;
    v_quat[0] = (vw - w * dotp) * inv;
    v_quat[1] = (vx - x * dotp) * inv;
    v_quat[2] = (vy - y * dotp) * inv;
    v_quat[3] = (vz - z * dotp) * inv;
    (void)q_raw;
}

// ---------------------------------------------------------------------------------------------
// spherical harmonics, degree 0..3   
// ---------------------------------------------------------------------------------------------
#define B2S_SH_C0 0.2820947917738781f
#define B2S_SH_C1 0.48860251190292f
#define B2S_SH_C20 1.092548430592079f
#define B2S_SH_C21 0.9461746957575601f
#define B2S_SH_C22 0.3153915652525201f
#define B2S_SH_C23 0.5462742152960395f
#define B2S_SH_C30 0.5900435899266435f
#define B2S_SH_C31 2.890611442640554f
#define B2S_SH_C32 2.285228997322329f
#define B2S_SH_C33 0.4570457994644658f
#define B2S_SH_C34 1.865881662950577f
#define B2S_SH_C35 1.119528997770346f
#define B2S_SH_C36 1.445305721320277f

// basis values Y[0..(deg+1)^2) for a unit direction
B2S_HD void sh_basis(int deg, float x, float y, float z, float* Y) {
    Y[0] = B2S_SH_C0;
    if (deg < 1) return;
    Y[1] = -B2S_SH_C1 * y; Y[2] = B2S_SH_C1 * z; Y[3] = -B2S_SH_C1 * x;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z;
    Y[4] = B2S_SH_C20 * x * y;
    Y[5] = -B2S_SH_C20 * y * z;
    Y[6] = B2S_SH_C21 * zz - B2S_SH_C22;
    Y[7] = -B2S_SH_C20 * x * z;
    Y[8] = B2S_SH_C23 * (xx - yy);
    if (deg < 3) return;
    Y[9] = -B2S_SH_C30 * y * (3.f * xx - yy);
    Y[10] = B2S_SH_C31 * x * y * z;
    Y[11] = y * (-B2S_SH_C32 * zz + B2S_SH_C33);
    Y[12] = z * (B2S_SH_C34 * zz - B2S_SH_C35);
    Y[13] = x * (-B2S_SH_C32 * zz + B2S_SH_C33);
    Y[14] = B2S_SH_C36 * z * (xx - yy);
    Y[15] = -B2S_SH_C30 * x * (xx - 3.f * yy);
}

// d(sum_k Y_k w_k)/d(x,y,z) for weights w_k (w_k = <coeff_k, v_colour>)
B2S_HD void sh_basis_vjp(int deg, float x, float y, float z, const float* w, float& vx, float& vy, float& vz) {
    vx = vy = vz = 0.f;
    if (deg < 1) return;
    vy += -B2S_SH_C1 * w[1]; vz += B2S_SH_C1 * w[2]; vx += -B2S_SH_C1 * w[3];
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z;
    vx += B2S_SH_C20 * y * w[4];               vy += B2S_SH_C20 * x * w[4];
    vy += -B2S_SH_C20 * z * w[5];              vz += -B2S_SH_C20 * y * w[5];
    vz += 2.f * B2S_SH_C21 * z * w[6];
    vx += -B2S_SH_C20 * z * w[7];              vz += -B2S_SH_C20 * x * w[7];
    vx += 2.f * B2S_SH_C23 * x * w[8];         vy += -2.f * B2S_SH_C23 * y * w[8];
    if (deg < 3) return;
    vx += -6.f * B2S_SH_C30 * x * y * w[9];    vy += -3.f * B2S_SH_C30 * (xx - yy) * w[9];
    vx += B2S_SH_C31 * y * z * w[10];          vy += B2S_SH_C31 * x * z * w[10];       vz += B2S_SH_C31 * x * y * w[10];
    vy += (-B2S_SH_C32 * zz + B2S_SH_C33) * w[11];                                      vz += -2.f * B2S_SH_C32 * y * z * w[11];
    vz += (3.f * B2S_SH_C34 * zz - B2S_SH_C35) * w[12];
    vx += (-B2S_SH_C32 * zz + B2S_SH_C33) * w[13];                                      vz += -2.f * B2S_SH_C32 * x * z * w[13];
    vx += 2.f * B2S_SH_C36 * x * z * w[14];    vy += -2.f * B2S_SH_C36 * y * z * w[14]; vz += B2S_SH_C36 * (xx - yy) * w[14];
    vx += -3.f * B2S_SH_C30 * (xx - yy) * w[15];                                        vy += 6.f * B2S_SH_C30 * x * y * w[15];
}

}  // namespace b2s
