// sh_grad_ops.h — the view-direction-factored SH gradient (SURVEY.md §8 e, multi-GPU exchange), shared by the kernel of
// sh_exchange.cu and a host-compiled test harness (tests/native/sh_grad_host.cpp).
//
// For one view the gradient of the higher SH bands is an outer product: dL/dshN[i][k][c] = B_k(dir_i) * dL/dcolour[i][c]
// (B_k = SH basis of the unit direction camera -> Gaussian, dL/dcolour after the colour clamp mask), and the band-0
// gradient the rasterizer already writes is the same dL/dcolour times a constant: dL/dsh0[i][c] = SH_C0 * dL/dcolour[i][c]
// (preprocess_bwd.cu:261).  So a rank that knows every view's camera centre and every view's dL/dsh0 (12 B per Gaussian
// and view) can form  sum_v B_k(dir_{v,i}) * dL/dsh0_v[i][c] / SH_C0  itself instead of receiving the 180 B per Gaussian
// of the summed dL/dshN.  Basis constants and signs as in preprocess_bwd.cu:266-288 (in-tree: gsplat_sh.hlsl:41-104).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define DVS_SG_HD __host__ __device__ __forceinline__
#else
#define DVS_SG_HD inline
#endif

namespace dvs_shx {

constexpr float kC0 = 0.28209479177387814f;

// b[k-1] = B_k(x, y, z) for k = 1 .. (deg+1)^2 - 1; (x, y, z) a unit vector
DVS_SG_HD void sh_rest_basis(int deg, float x, float y, float z, float b[15]) {
    if (deg < 1) return;
    const float C1 = 0.4886025119029199f;
    b[0] = -C1 * y; b[1] = C1 * z; b[2] = -C1 * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[3] = 1.0925484305920792f * xy;
    b[4] = -1.0925484305920792f * yz;
    b[5] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[6] = -1.0925484305920792f * xz;
    b[7] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[8] = -0.5900435899266435f * y * (3.0f * xx - yy);
    b[9] = 2.890611442640554f * xy * z;
    b[10] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    b[11] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[12] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    b[13] = 1.445305721320277f * z * (xx - yy);
    b[14] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// acc[3 * (k-1) + c] += B_k(dir) * dsh0[c] / SH_C0 for one view; mean / campos in world space.  A Gaussian the view did
// not see has dsh0 = 0 and is skipped (its direction may be degenerate).
DVS_SG_HD void accumulate_view(int deg, const float mean[3], const float campos[3], const float dsh0[3], float acc[45]) {
    if (dsh0[0] == 0.0f && dsh0[1] == 0.0f && dsh0[2] == 0.0f) return;
    const float ox = mean[0] - campos[0], oy = mean[1] - campos[1], oz = mean[2] - campos[2];
    const float li = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
    float b[15];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 15; k++) b[k] = 0.0f;
    sh_rest_basis(deg, ox * li, oy * li, oz * li, b);
    const float d0 = dsh0[0] / kC0, d1 = dsh0[1] / kC0, d2 = dsh0[2] / kC0;
    // bands above the active degree have b = 0: a fixed trip count keeps acc[] in registers on the device
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 15; k++) {
        acc[3 * k] += b[k] * d0;
        acc[3 * k + 1] += b[k] * d1;
        acc[3 * k + 2] += b[k] * d2;
    }
}

// ---- the kernel of sh_exchange.cu, written as the two phases a CTA runs between barriers, so that a host harness can run
// the very same indexing (tile loop, shared-memory rows, 128-bit copy with scalar tail) thread by thread.
struct ExchangeArgs {
    const float* means;     // [N,3]
    const float* campos;    // [V,3]
    const float* dsh0_all;  // [V,N,3]
    long long N;
    int V, deg, RW;         // RW = 3 * sh_rest_alloc words per output row (<= 45)
    float* out;             // [N,RW]
    int vec_ok;             // out is 16-byte aligned
};
// phase 1: thread `tid` of the CTA that owns Gaussians [base, base + cnt) accumulates its row into rows[tid * RW ..]
DVS_SG_HD void exchange_compute(const ExchangeArgs& a, float* rows, int tid, long long base, int cnt) {
    if (tid >= cnt) return;
    const long long i = base + tid;
    float acc[45];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 45; k++) acc[k] = 0.0f;
    const float mean[3] = {a.means[3 * i], a.means[3 * i + 1], a.means[3 * i + 2]};
    // views in groups of four: the twelve loads of a group are issued before any of its arithmetic, so a thread waits for
    // memory once per four views instead of once per view (the kernel runs at low occupancy — 45 accumulators per thread —
    // and was latency-bound with one dependent load per view)
    for (int v0 = 0; v0 < a.V; v0 += 4) {
        float dc[4][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 4; u++) {
            const int v = v0 + u < a.V ? v0 + u : a.V - 1;  // (clamped: a harmless repeat of the last view's load)
            const float* d = a.dsh0_all + ((size_t)v * (size_t)a.N + (size_t)i) * 3;
            dc[u][0] = d[0]; dc[u][1] = d[1]; dc[u][2] = d[2];
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 4; u++)
            if (v0 + u < a.V) accumulate_view(a.deg, mean, a.campos + 3 * (v0 + u), dc[u], acc);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 45; k++)
        if (k < a.RW) rows[tid * a.RW + k] = acc[k];
}
// phase 2: the CTA's rows are one contiguous span of the output: `nthreads` threads copy it, 128 bits at a time when the
// span starts on a 16-byte boundary (always for RW = 45: base is a multiple of 128), scalar tail
DVS_SG_HD void exchange_store(const ExchangeArgs& a, const float* rows, int tid, int nthreads, long long base, int cnt) {
    float* dst = a.out + base * a.RW;
    const int n_words = cnt * a.RW;
    if (a.vec_ok && ((base * a.RW) & 3) == 0) {
        struct alignas(16) V4 { float x, y, z, w; };
        const int n_vec = n_words >> 2;
        for (int k = tid; k < n_vec; k += nthreads) reinterpret_cast<V4*>(dst)[k] = reinterpret_cast<const V4*>(rows)[k];
        for (int k = (n_vec << 2) + tid; k < n_words; k += nthreads) dst[k] = rows[k];
    } else {
        for (int k = tid; k < n_words; k += nthreads) dst[k] = rows[k];
    }
}

}  // namespace dvs_shx
