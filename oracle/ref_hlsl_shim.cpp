// TEST INFRASTRUCTURE — executes, as C++, the REFERENCE's own shader functions that restate the per-Gaussian forward
// maths (the only statement of that maths inside /root/reference; the trainer's rasterizer itself is absent,
// SURVEY.md §0).  The function bodies are cut out of the reference's .hlsl files at build time (oracle/Makefile,
// `make ref` -> oracle/_ref/gen/*.inc); nothing of them is stored in this repo.  tests/test_oracle_vs_reference_hlsl.py
// checks oracle/dvs_oracle.c (and through it the CUDA kernels) against these.
#include "hlsl_prelude.hpp"

#define SH_DEGREE 3
namespace refhlsl {
#include "_ref/gen/intersect_cov.inc"  // gsplat_intersect.hlsl:61-134  computeCov3D / computeCov2D
#include "_ref/gen/sh.inc"             // gsplat_sh.hlsl:42-103         SH_C1.. constants, evalSH
#include "_ref/gen/ndc2pix.inc"        // gsplat_vs.hlsl:211-214        ndc2Pix
float aa_factor(const float3x3& cov2D) {
#include "_ref/gen/aa.inc"             // gsplat_vs.hlsl:297-300        detOrig / detBlur / corner_aaFactor
    return corner_aaFactor;
}
}  // namespace refhlsl

extern "C" {
#define EXPORT __attribute__((visibility("default")))

// scale[3] (activated), mod, rot[4] = (r,x,y,z) used as given -> cov3D[6] upper triangle
EXPORT void ref_hlsl_cov3d(const float* scale, float mod, const float* rot, float* cov6) {
    float3x3 c;
    refhlsl::computeCov3D(float3(scale[0], scale[1], scale[2]), mod, float4(rot[0], rot[1], rot[2], rot[3]), c);
    cov6[0] = c[0][0]; cov6[1] = c[0][1]; cov6[2] = c[0][2]; cov6[3] = c[1][1]; cov6[4] = c[1][2]; cov6[5] = c[2][2];
}

// view16 = the caller's flat matrix (element [4c+r] = row r, column c of the column-vector world->view matrix), which
// read as a row-major float4x4 is the row-vector form the shaders use (mul(world_pos, view), gsplat_vs.hlsl:241-245).
// Returns (cov00 + 0.3, cov11 + 0.3, cov01).
EXPORT void ref_hlsl_cov2d(const float* p_view, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                           const float* cov6, const float* view16, float* out3) {
    float3x3 c(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
    float4x4 v;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) v[i][j] = view16[4 * i + j];
    const float3 r = refhlsl::computeCov2D(float3(p_view[0], p_view[1], p_view[2]), focal_x, focal_y, tan_fovx, tan_fovy, c, v, 1.0f);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// sh_rest[15][3] (RGB interleaved), unit direction -> the degree 1..3 part of the colour
EXPORT void ref_hlsl_eval_sh(const float* sh_rest, const float* dir, float* rgb) {
    float3 sh[15];
    for (int k = 0; k < 15; k++) sh[k] = float3(sh_rest[3 * k], sh_rest[3 * k + 1], sh_rest[3 * k + 2]);
    const float3 r = refhlsl::evalSH(sh, float3(dir[0], dir[1], dir[2]));
    rgb[0] = r.x; rgb[1] = r.y; rgb[2] = r.z;
}

EXPORT float ref_hlsl_ndc2pix(float v, int S) { return refhlsl::ndc2Pix(v, S); }

// un-blurred 2x2 covariance (a, b, c) -> sqrt(max(det / det_blurred, 0))
EXPORT float ref_hlsl_aa_factor(float a, float b, float c) {
    float3x3 m(a, b, 0, b, c, 0, 0, 0, 0);
    return refhlsl::aa_factor(m);
}
}
