// aux_outputs.cu — depth / alpha maps and their gradients (outputs the closed trainer needs for
// `normalConsistencyLoss` and mesh extraction, docs/userGuide.md:52-58) WITHOUT a second compositing kernel.
//
// The compositor is linear in the per-splat colour: depth(p) = sum_k w_k(p) z_k and alpha(p) = sum_k w_k(p) are the image
// of the colour triple (z_k, 1, 0) over a zero background.  So the forward runs the same render_fwd kernel on a copy of the
// 48-byte screen records whose colour fields are replaced (aux_records_kernel), and the backward runs the same render_bwd
// kernel on that copy BEFORE the colour backward: its geometry sums land in the six geometry slots of the shared
// screen-gradient record, where the colour pass adds its own (the per-Gaussian backward then sees the gradient of the
// whole loss), and its "colour" sum is dL/dz_k, which aux_extract_kernel moves out of the record (slots 6-8 are zeroed
// for the colour pass) and aux_depth_grad_kernel adds to dL/dmean through row 2 of the view matrix.
// tests/aux_ref.py restates exactly this composition with the oracle's functions; tests/test_aux_outputs.py checks that
// restatement against float64 autograd on the CPU and this file against the restatement on the GPU.
// The normal map (sum_k w_k n_k, n_k = shortest ellipsoid axis in view space, aux_normal_ops.h) is a third pass of the
// same kind with the colour triple n_k; its colour sums are dL/dn_k, chained to the quaternion by aux_normal_grad_kernel.
#include "aux_normal_ops.h"
#include "common.cuh"
#include "kernels.h"

namespace dvs {

namespace {
constexpr int AUX_THREADS = 256;
inline int aux_grid(int64_t n) { return (int)std::min<int64_t>((n + AUX_THREADS - 1) / AUX_THREADS, 148 * 8); }

// record layout (preprocess_fwd.cu): q0 {mx, my, A2, B2}  q1 {C2, lo, r, g}  q2 {b, depth, radius, tiles | clamped}
__global__ void __launch_bounds__(AUX_THREADS) aux_records_kernel(int N, const float4* __restrict__ rec, float4* __restrict__ out) {
    for (int i = blockIdx.x * AUX_THREADS + threadIdx.x; i < N; i += gridDim.x * AUX_THREADS) {
        const float4 q0 = rec[3 * (size_t)i], q1 = rec[3 * (size_t)i + 1], q2 = rec[3 * (size_t)i + 2];
        out[3 * (size_t)i] = q0;
        out[3 * (size_t)i + 1] = make_float4(q1.x, q1.y, q2.y, 1.0f);  // colour (z, 1, .)
        out[3 * (size_t)i + 2] = make_float4(0.0f, q2.y, q2.z, q2.w);  // (., ., 0)
    }
}
// screen-gradient record (render_bwd.cu): {Sx, Sy, Sxx, Sxy} {Syy, S0, c0, c1} {c2, ax, ay, -}
__global__ void __launch_bounds__(AUX_THREADS) aux_extract_kernel(int N, float4* __restrict__ sgrad, float* __restrict__ dz) {
    for (int i = blockIdx.x * AUX_THREADS + threadIdx.x; i < N; i += gridDim.x * AUX_THREADS) {
        float4 s1 = sgrad[3 * (size_t)i + 1];
        dz[i] = s1.z;
        s1.z = 0.0f; s1.w = 0.0f;
        sgrad[3 * (size_t)i + 1] = s1;
        float4 s2 = sgrad[3 * (size_t)i + 2];
        s2.x = 0.0f;
        sgrad[3 * (size_t)i + 2] = s2;
    }
}
__global__ void __launch_bounds__(AUX_THREADS) aux_depth_grad_kernel(int N, const float* __restrict__ dz, float v0, float v1, float v2,
                                                                      float* __restrict__ dmeans) {
    for (int i = blockIdx.x * AUX_THREADS + threadIdx.x; i < N; i += gridDim.x * AUX_THREADS) {
        const float d = dz[i];
        if (d != 0.0f) {
            dmeans[3 * (size_t)i] += v0 * d;
            dmeans[3 * (size_t)i + 1] += v1 * d;
            dmeans[3 * (size_t)i + 2] += v2 * d;
        }
    }
}
struct View16 { float m[16]; };
__global__ void __launch_bounds__(AUX_THREADS) aux_normal_records_kernel(int N, View16 view, bool activated, const float* __restrict__ quats,
                                                                          const float* __restrict__ scales, const float* __restrict__ means,
                                                                          const float4* __restrict__ rec, float4* __restrict__ out) {
    for (int i = blockIdx.x * AUX_THREADS + threadIdx.x; i < N; i += gridDim.x * AUX_THREADS) {
        const float4 q4 = reinterpret_cast<const float4*>(quats)[i];
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float sc[3] = {scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]};
        const float mu[3] = {means[3 * (size_t)i], means[3 * (size_t)i + 1], means[3 * (size_t)i + 2]};
        float n[3], flip;
        int axis;
        dvs_aux::normal_forward(q, sc, mu, view.m, activated, n, &axis, &flip);
        const float4 q0 = rec[3 * (size_t)i], q1 = rec[3 * (size_t)i + 1], q2 = rec[3 * (size_t)i + 2];
        out[3 * (size_t)i] = q0;
        out[3 * (size_t)i + 1] = make_float4(q1.x, q1.y, n[0], n[1]);
        out[3 * (size_t)i + 2] = make_float4(n[2], q2.y, q2.z, q2.w);
    }
}
__global__ void __launch_bounds__(AUX_THREADS) aux_extract3_kernel(int N, float4* __restrict__ sgrad, float* __restrict__ dn) {
    for (int i = blockIdx.x * AUX_THREADS + threadIdx.x; i < N; i += gridDim.x * AUX_THREADS) {
        float4 s1 = sgrad[3 * (size_t)i + 1];
        float4 s2 = sgrad[3 * (size_t)i + 2];
        dn[3 * (size_t)i] = s1.z; dn[3 * (size_t)i + 1] = s1.w; dn[3 * (size_t)i + 2] = s2.x;
        s1.z = 0.0f; s1.w = 0.0f; s2.x = 0.0f;
        sgrad[3 * (size_t)i + 1] = s1;
        sgrad[3 * (size_t)i + 2] = s2;
    }
}
__global__ void __launch_bounds__(AUX_THREADS) aux_normal_grad_kernel(int N, View16 view, bool activated, const float* __restrict__ quats,
                                                                       const float* __restrict__ scales, const float* __restrict__ means,
                                                                       const float* __restrict__ dn, float* __restrict__ dquats) {
    for (int i = blockIdx.x * AUX_THREADS + threadIdx.x; i < N; i += gridDim.x * AUX_THREADS) {
        const float g[3] = {dn[3 * (size_t)i], dn[3 * (size_t)i + 1], dn[3 * (size_t)i + 2]};
        if (g[0] == 0.0f && g[1] == 0.0f && g[2] == 0.0f) continue;  // invisible, or no gradient reached it
        const float4 q4 = reinterpret_cast<const float4*>(quats)[i];
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float sc[3] = {scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]};
        const float mu[3] = {means[3 * (size_t)i], means[3 * (size_t)i + 1], means[3 * (size_t)i + 2]};
        float n[3], flip, dq[4];
        int axis;
        dvs_aux::normal_forward(q, sc, mu, view.m, activated, n, &axis, &flip);  // the decisions of the forward, recomputed
        dvs_aux::normal_backward(q, axis, flip, view.m, activated, g, dq);
        float4 o = reinterpret_cast<float4*>(dquats)[i];
        o.x += dq[0]; o.y += dq[1]; o.z += dq[2]; o.w += dq[3];
        reinterpret_cast<float4*>(dquats)[i] = o;
    }
}
}  // namespace

cudaError_t launch_aux_normal_records(const Cam& cam, int N, const Params& prm, const float4* rec, float4* rec_aux, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    View16 v;
    for (int k = 0; k < 16; k++) v.m[k] = cam.view[k];
    aux_normal_records_kernel<<<aux_grid(N), AUX_THREADS, 0, st>>>(N, v, (cam.flags & DVS_FLAG_INPUT_ACTIVATED) != 0, prm.quats, prm.scales,
                                                                  prm.means3D, rec, rec_aux);
    return cudaGetLastError();
}
cudaError_t launch_aux_extract3(int N, float4* sgrad, float* dn, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    aux_extract3_kernel<<<aux_grid(N), AUX_THREADS, 0, st>>>(N, sgrad, dn);
    return cudaGetLastError();
}
cudaError_t launch_aux_normal_grad(const Cam& cam, int N, const Params& prm, const float* dn, float* dquats, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    View16 v;
    for (int k = 0; k < 16; k++) v.m[k] = cam.view[k];
    aux_normal_grad_kernel<<<aux_grid(N), AUX_THREADS, 0, st>>>(N, v, (cam.flags & DVS_FLAG_INPUT_ACTIVATED) != 0, prm.quats, prm.scales,
                                                               prm.means3D, dn, dquats);
    return cudaGetLastError();
}

cudaError_t launch_aux_records(int N, const float4* rec, float4* rec_aux, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    aux_records_kernel<<<aux_grid(N), AUX_THREADS, 0, st>>>(N, rec, rec_aux);
    return cudaGetLastError();
}
cudaError_t launch_aux_extract(int N, float4* sgrad, float* dz, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    aux_extract_kernel<<<aux_grid(N), AUX_THREADS, 0, st>>>(N, sgrad, dz);
    return cudaGetLastError();
}
cudaError_t launch_aux_depth_grad(int N, const float* dz, const float view_row2[3], float* dmeans, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    aux_depth_grad_kernel<<<aux_grid(N), AUX_THREADS, 0, st>>>(N, dz, view_row2[0], view_row2[1], view_row2[2], dmeans);
    return cudaGetLastError();
}

__global__ void background_grad_kernel(int64_t P, const float* __restrict__ final_T, const float* __restrict__ dL_dpix,
                                       float* __restrict__ dL_dbg, const uint32_t* __restrict__ info) {
    const bool dead = info[2] != 0u;  // the forward overflowed its arena: no image, no transmittance
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const float T = dead ? 0.f : final_T[p];
        dL_dbg[p] = T * dL_dpix[p];
        dL_dbg[P + p] = T * dL_dpix[P + p];
        dL_dbg[2 * P + p] = T * dL_dpix[2 * P + p];
    }
}
cudaError_t launch_background_grad(int64_t P, const float* final_T, const float* dL_dpix, float* dL_dbg, const uint32_t* info,
                                   cudaStream_t st) {
    if (P > 0) background_grad_kernel<<<148 * 8, 256, 0, st>>>(P, final_T, dL_dpix, dL_dbg, info);
    return cudaGetLastError();
}

}  // namespace dvs
