// aux_normal_ops.h — per-Gaussian arithmetic of the normal map, shared by aux_outputs.cu and a
// host-compiled test harness (tests/native/aux_normal_host.cpp).  The normal of a Gaussian is the shortest axis of its
// ellipsoid (column argmin(scale) of R(q), the convention of the surface-aligned 3DGS variants the closed trainer's
// `normalConsistencyLoss` option belongs to), expressed in view space and turned towards the camera.  The axis choice
// and the orientation are piecewise constant: they are decisions, not differentiated.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define DVS_AN_HD __host__ __device__ __forceinline__
#else
#define DVS_AN_HD inline
#endif

namespace dvs_aux {

// view: flat [4c + r] world -> camera (dvs_camera.view).  activated: quaternion used as given (DVS_FLAG_INPUT_ACTIVATED),
// otherwise normalised first.  Returns the view-space normal; *axis (0..2) and *flip (+1 / -1) are the decisions taken.
DVS_AN_HD void normal_forward(const float q_in[4], const float scale[3], const float mean[3], const float* view, bool activated,
                              float n_v[3], int* axis, float* flip) {
    float r = q_in[0], x = q_in[1], y = q_in[2], z = q_in[3];
    if (!activated) {
        const float inv = 1.0f / sqrtf(r * r + x * x + y * y + z * z);
        r *= inv; x *= inv; y *= inv; z *= inv;
    }
    int j = 0;
    if (scale[1] < scale[j]) j = 1;
    if (scale[2] < scale[j]) j = 2;
    float n[3];
    if (j == 0) { n[0] = 1.f - 2.f * (y * y + z * z); n[1] = 2.f * (x * y + r * z); n[2] = 2.f * (x * z - r * y); }
    else if (j == 1) { n[0] = 2.f * (x * y - r * z); n[1] = 1.f - 2.f * (x * x + z * z); n[2] = 2.f * (y * z + r * x); }
    else { n[0] = 2.f * (x * z + r * y); n[1] = 2.f * (y * z - r * x); n[2] = 1.f - 2.f * (x * x + y * y); }
    float t[3], nv[3];
    for (int a = 0; a < 3; a++) {  // row a of the view matrix
        nv[a] = view[a] * n[0] + view[4 + a] * n[1] + view[8 + a] * n[2];
        t[a] = view[a] * mean[0] + view[4 + a] * mean[1] + view[8 + a] * mean[2] + view[12 + a];
    }
    const float s = (nv[0] * t[0] + nv[1] * t[1] + nv[2] * t[2]) > 0.f ? -1.f : 1.f;  // towards the camera
    for (int a = 0; a < 3; a++) n_v[a] = s * nv[a];
    *axis = j; *flip = s;
}

// dL/dq (stored quaternion) from dL/dn_v, for the decisions (axis, flip) of the forward
DVS_AN_HD void normal_backward(const float q_in[4], int j, float flip, const float* view, bool activated, const float dn_v[3],
                               float dq[4]) {
    float r = q_in[0], x = q_in[1], y = q_in[2], z = q_in[3], len = 1.f;
    if (!activated) {
        len = sqrtf(r * r + x * x + y * y + z * z);
        const float inv = 1.0f / len;
        r *= inv; x *= inv; y *= inv; z *= inv;
    }
    float g[3];  // dL/dn_w = flip * V_rot^T dL/dn_v
    for (int c = 0; c < 3; c++) g[c] = flip * (view[4 * c] * dn_v[0] + view[4 * c + 1] * dn_v[1] + view[4 * c + 2] * dn_v[2]);
    float d[4];
    if (j == 0) {
        d[0] = 2.f * (z * g[1] - y * g[2]);
        d[1] = 2.f * (y * g[1] + z * g[2]);
        d[2] = -4.f * y * g[0] + 2.f * x * g[1] - 2.f * r * g[2];
        d[3] = -4.f * z * g[0] + 2.f * r * g[1] + 2.f * x * g[2];
    } else if (j == 1) {
        d[0] = -2.f * z * g[0] + 2.f * x * g[2];
        d[1] = 2.f * y * g[0] - 4.f * x * g[1] + 2.f * r * g[2];
        d[2] = 2.f * x * g[0] + 2.f * z * g[2];
        d[3] = -2.f * r * g[0] - 4.f * z * g[1] + 2.f * y * g[2];
    } else {
        d[0] = 2.f * y * g[0] - 2.f * x * g[1];
        d[1] = 2.f * z * g[0] - 2.f * r * g[1] - 4.f * x * g[2];
        d[2] = 2.f * r * g[0] + 2.f * z * g[1] - 4.f * y * g[2];
        d[3] = 2.f * x * g[0] + 2.f * y * g[1];
    }
    if (activated) {
        for (int k = 0; k < 4; k++) dq[k] = d[k];
    } else {  // through q / |q|
        const float qh[4] = {r, x, y, z};
        const float dot = qh[0] * d[0] + qh[1] * d[1] + qh[2] * d[2] + qh[3] * d[3];
        for (int k = 0; k < 4; k++) dq[k] = (d[k] - qh[k] * dot) / len;
    }
}

}  // namespace dvs_aux
