// preprocess_fwd.cu — A1: per-Gaussian forward (3D->2D EWA covariance projection + SH evaluation)
// plus per-tile duplicate counting (first half of A2/A3).
//
// Replaces `preprocessCUDA` of the absent gsplatrast operator (algorithm: Appendix
// B.1; in-tree corroboration of the maths: diverse/assets/shaders/gaussian/gsplat_intersect.hlsl:61-134
// (cov3D, cov2D, 1.3*tanfov clamp, +0.3), gsplat_vs.hlsl:189-214 (R(q), ndc2Pix),
// gsplat_viewz_cs.hlsl:197-199 (1/(w+1e-7)), gsplat_sh.hlsl:42-103 (SH)).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every a*b+c below is two IEEE roundings and
// every fmaf() is one, so that radius / tile rect / depth key / mean2D / rgb are the literal
// operation sequence of SURVEY.md Appendix B.6 and come out bit-identical to the CPU oracle.
//
// Roofline: HBM.  Algorithmic bytes per Gaussian: (44+12K) read + 48 (record) + 16 (aux) written
// for visible ones.  One thread per Gaussian for the geometry; the 12*(K-1)-byte SH row is
// staged per warp through shared memory with 128-bit coalesced loads.
#include "common.cuh"
#include "kernels.h"

namespace dvs {

// ---- deterministic exp (identical op sequence on CPU and GPU) -------------------------------
__device__ __forceinline__ float det_expf(float x) {
    x = fminf(fmaxf(x, -87.0f), 88.0f);
    float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r) + 1.0f;
    int ni = (int)n;
    return e * __uint_as_float((uint32_t)(ni + 127) << 23);
}

__device__ __forceinline__ void sh_basis_dev(int deg, float x, float y, float z, float* b) {
    const float C1 = 0.4886025119029199f;
    b[0] = 0.28209479177387814f;
    if (deg < 1) return;
    b[1] = -C1 * y;
    b[2] = C1 * z;
    b[3] = -C1 * x;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy;
    b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (fmaf(2.0f, zz, -xx) - yy);
    b[7] = -1.0925484305920792f * xz;
    b[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[9] = -0.5900435899266435f * y * fmaf(3.0f, xx, -yy);
    b[10] = 2.890611442640554f * xy * z;
    b[11] = -0.4570457994644658f * y * (fmaf(4.0f, zz, -xx) - yy);
    b[12] = 0.3731763325901154f * z * (fmaf(2.0f, zz, -(3.0f * xx)) - 3.0f * yy);
    b[13] = -0.4570457994644658f * x * (fmaf(4.0f, zz, -xx) - yy);
    b[14] = 1.445305721320277f * z * (xx - yy);
    b[15] = -0.5900435899266435f * x * fmaf(-3.0f, yy, xx);
}

constexpr int PF_THREADS = 128;

template <int DEG>
__global__ void __launch_bounds__(PF_THREADS)
preprocess_fwd_kernel(Cam cam, int N, Params prm, float4* __restrict__ rec, uint4* __restrict__ aux,
                      uint32_t* __restrict__ tile_count, int32_t* __restrict__ out_radii,
                      unsigned long long* __restrict__ stats /* [0]=V, [1]=D */) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    extern __shared__ float sh_stage[];  // [warps][32 * 3 * KR] floats
    const int KR = cam.KR;
    const int row = 3 * KR;  // floats of shN per Gaussian
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * PF_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * PF_THREADS + warp * 32;

    // ---- stage this warp's shN rows: 32*row contiguous floats, 128-bit coalesced ----
    float* mysh = sh_stage + (size_t)warp * 32 * row;
    if (K > 1 && warp_first < N) {
        const int nrows = min(32, N - warp_first);
        const int nflt = nrows * row;
        const float* src = prm.shN + (size_t)warp_first * row;
        const int nvec = nflt >> 2;
        const float4* src4 = reinterpret_cast<const float4*>(src);
        float4* dst4 = reinterpret_cast<float4*>(mysh);
        for (int v = lane; v < nvec; v += 32) dst4[v] = ldg_nc_f4(src4 + v);
        for (int t = (nvec << 2) + lane; t < nflt; t += 32) mysh[t] = __ldg(src + t);
    }
    __syncwarp();

    bool visible = false;
    uint32_t tiles = 0;
    if (i < N) {
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0;
        uint4 ax = make_uint4(0u, 0u, 0u, 0u);
        int rad = 0;
        const float px = __ldg(prm.means3D + 3 * (size_t)i), py = __ldg(prm.means3D + 3 * (size_t)i + 1),
                    pz = __ldg(prm.means3D + 3 * (size_t)i + 2);
        const float* V = cam.view;
        const float* P = cam.proj;
        const float t0 = fmaf(V[0], px, fmaf(V[4], py, fmaf(V[8], pz, V[12])));
        const float t1 = fmaf(V[1], px, fmaf(V[5], py, fmaf(V[9], pz, V[13])));
        const float t2 = fmaf(V[2], px, fmaf(V[6], py, fmaf(V[10], pz, V[14])));
        do {
            if (t2 <= 0.2f) break;
            const float h0 = fmaf(P[0], px, fmaf(P[4], py, fmaf(P[8], pz, P[12])));
            const float h1 = fmaf(P[1], px, fmaf(P[5], py, fmaf(P[9], pz, P[13])));
            const float h3 = fmaf(P[3], px, fmaf(P[7], py, fmaf(P[11], pz, P[15])));
            const float w_inv = 1.0f / (h3 + 1e-7f);
            const float ndcx = h0 * w_inv, ndcy = h1 * w_inv;
            // activations
            float s0, s1, s2, qr, qx, qy, qz, o;
            {
                const float a0 = __ldg(prm.scales + 3 * (size_t)i), a1 = __ldg(prm.scales + 3 * (size_t)i + 1),
                            a2 = __ldg(prm.scales + 3 * (size_t)i + 2);
                const float4 qq = __ldg(reinterpret_cast<const float4*>(prm.quats) + i);
                const float oo = __ldg(prm.opacities + i);
                if (cam.flags & DVS_FLAG_INPUT_ACTIVATED) {
                    s0 = cam.scale_modifier * a0; s1 = cam.scale_modifier * a1; s2 = cam.scale_modifier * a2;
                    qr = qq.x; qx = qq.y; qy = qq.z; qz = qq.w;
                    o = oo;
                } else {
                    s0 = cam.scale_modifier * det_expf(a0);
                    s1 = cam.scale_modifier * det_expf(a1);
                    s2 = cam.scale_modifier * det_expf(a2);
                    const float n2 = fmaf(qq.x, qq.x, fmaf(qq.y, qq.y, fmaf(qq.z, qq.z, qq.w * qq.w)));
                    const float inv = 1.0f / sqrtf(n2);
                    qr = qq.x * inv; qx = qq.y * inv; qy = qq.z * inv; qz = qq.w * inv;
                    o = 1.0f / (1.0f + det_expf(-oo));
                }
            }
            // R(q), M = R S, Sigma = M M^T
            const float R00 = fmaf(-2.0f, fmaf(qz, qz, qy * qy), 1.0f);
            const float R01 = 2.0f * fmaf(qx, qy, -(qr * qz));
            const float R02 = 2.0f * fmaf(qx, qz, qr * qy);
            const float R10 = 2.0f * fmaf(qx, qy, qr * qz);
            const float R11 = fmaf(-2.0f, fmaf(qz, qz, qx * qx), 1.0f);
            const float R12 = 2.0f * fmaf(qy, qz, -(qr * qx));
            const float R20 = 2.0f * fmaf(qx, qz, -(qr * qy));
            const float R21 = 2.0f * fmaf(qy, qz, qr * qx);
            const float R22 = fmaf(-2.0f, fmaf(qy, qy, qx * qx), 1.0f);
            const float M00 = R00 * s0, M01 = R01 * s1, M02 = R02 * s2;
            const float M10 = R10 * s0, M11 = R11 * s1, M12 = R12 * s2;
            const float M20 = R20 * s0, M21 = R21 * s1, M22 = R22 * s2;
            const float S00 = fmaf(M00, M00, fmaf(M01, M01, M02 * M02));
            const float S01 = fmaf(M00, M10, fmaf(M01, M11, M02 * M12));
            const float S02 = fmaf(M00, M20, fmaf(M01, M21, M02 * M22));
            const float S11 = fmaf(M10, M10, fmaf(M11, M11, M12 * M12));
            const float S12 = fmaf(M10, M20, fmaf(M11, M21, M12 * M22));
            const float S22 = fmaf(M20, M20, fmaf(M21, M21, M22 * M22));
            // EWA
            const float fx = (float)cam.W / (2.0f * cam.tanfovx), fy = (float)cam.H / (2.0f * cam.tanfovy);
            const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
            const float txtz = t0 / t2, tytz = t1 / t2;
            const float tx = fminf(limx, fmaxf(-limx, txtz)) * t2;
            const float ty = fminf(limy, fmaxf(-limy, tytz)) * t2;
            const float tz2 = t2 * t2;
            const float J00 = fx / t2, J02 = -(fx * tx) / tz2, J11 = fy / t2, J12 = -(fy * ty) / tz2;
            const float T00 = fmaf(J00, V[0], J02 * V[2]), T01 = fmaf(J00, V[4], J02 * V[6]),
                        T02 = fmaf(J00, V[8], J02 * V[10]);
            const float T10 = fmaf(J11, V[1], J12 * V[2]), T11 = fmaf(J11, V[5], J12 * V[6]),
                        T12 = fmaf(J11, V[9], J12 * V[10]);
            const float U00 = fmaf(T00, S00, fmaf(T01, S01, T02 * S02));
            const float U01 = fmaf(T00, S01, fmaf(T01, S11, T02 * S12));
            const float U02 = fmaf(T00, S02, fmaf(T01, S12, T02 * S22));
            const float U10 = fmaf(T10, S00, fmaf(T11, S01, T12 * S02));
            const float U11 = fmaf(T10, S01, fmaf(T11, S11, T12 * S12));
            const float U12 = fmaf(T10, S02, fmaf(T11, S12, T12 * S22));
            const float ca = fmaf(U00, T00, fmaf(U01, T01, U02 * T02)) + 0.3f;
            const float cb = fmaf(U00, T10, fmaf(U01, T11, U02 * T12));
            const float cc = fmaf(U10, T10, fmaf(U11, T11, U12 * T12)) + 0.3f;
            const float det = fmaf(ca, cc, -(cb * cb));
            if (det == 0.0f) break;
            const float det_inv = 1.0f / det;
            const float cA = cc * det_inv, cB = -cb * det_inv, cC = ca * det_inv;
            const float mid = 0.5f * (ca + cc);
            const float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
            const float l1 = mid + sq, l2 = mid - sq;
            const float rad_f = ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
            const int radius = (int)rad_f;
            const float mx = fmaf(ndcx + 1.0f, (float)cam.W, -1.0f) * 0.5f;
            const float my = fmaf(ndcy + 1.0f, (float)cam.H, -1.0f) * 0.5f;
            const float radf = (float)radius;
            const int minx = min(cam.gx, max(0, (int)((mx - radf) * 0.0625f)));
            const int miny = min(cam.gy, max(0, (int)((my - radf) * 0.0625f)));
            const int maxx = min(cam.gx, max(0, (int)((mx + radf + 15.0f) * 0.0625f)));
            const int maxy = min(cam.gy, max(0, (int)((my + radf + 15.0f) * 0.0625f)));
            const long long area = (long long)(maxx - minx) * (long long)(maxy - miny);
            if (area <= 0) break;
            // colour
            float d0 = px - cam.campos[0], d1 = py - cam.campos[1], d2 = pz - cam.campos[2];
            const float len = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
            const float linv = 1.0f / len;
            d0 *= linv; d1 *= linv; d2 *= linv;
            float bas[16];
            sh_basis_dev(DEG, d0, d1, d2, bas);
            float col[3];
            uint32_t clamped = 0;
            const float* myrow = mysh + lane * row;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float acc = bas[0] * __ldg(prm.sh0 + 3 * (size_t)i + ch);
#pragma unroll
                for (int k = 1; k < K; k++) acc = fmaf(bas[k], myrow[3 * (k - 1) + ch], acc);
                acc += 0.5f;
                if (acc < 0.0f) clamped |= 1u << ch;
                col[ch] = fmaxf(acc, 0.0f);
            }
            // ---- everything below is outside the bit-exact contract (derived fields) ----
            visible = true;
            tiles = (uint32_t)area;
            rad = radius;
            const float lo = log2f(o);
            q0 = make_float4(mx, my, (-0.5f * LOG2E) * cA, (-LOG2E) * cB);
            q1 = make_float4((-0.5f * LOG2E) * cC, lo, col[0], col[1]);
            q2 = make_float4(col[2], t2, __int_as_float(radius), __uint_as_float(tiles | (clamped << 24)));
            // opacity-aware AABB half extents of {alpha >= 1/255}: d^T conic d <= 2 ln2 (lo - log2(1/255))
            const float m = lo - ALPHA_MIN_LOG2;
            float ex = -1.0f, ey = -1.0f;  // negative: never contributes
            if (m > 0.0f) {
                const float k2 = 2.0f * LN2 * m;
                ex = sqrtf(k2 * ca) * 1.0001f + 0.01f;
                ey = sqrtf(k2 * cc) * 1.0001f + 0.01f;
            }
            ax = make_uint4((uint32_t)minx | ((uint32_t)miny << 16), (uint32_t)maxx | ((uint32_t)maxy << 16),
                            __float_as_uint(ex), __float_as_uint(ey));
            // per-tile duplicate counts (RED, no return)
            for (int y = miny; y < maxy; y++)
                for (int x = minx; x < maxx; x++) atomicAdd(tile_count + y * cam.gx + x, 1u);
        } while (false);
        float4* r = rec + 3 * (size_t)i;
        r[0] = q0; r[1] = q1; r[2] = q2;
        aux[i] = ax;
        if (out_radii) out_radii[i] = rad;
    }
    // stats: V and D
    const unsigned vm = __ballot_sync(0xffffffffu, visible);
    uint32_t tsum = tiles;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, off);
    if (lane == 0 && vm) {
        atomicAdd(stats + 0, (unsigned long long)__popc(vm));
        atomicAdd(stats + 1, (unsigned long long)tsum);
    }
}

cudaError_t launch_preprocess_fwd(const Cam& cam, int N, const Params& prm, float4* rec, uint4* aux,
                                  uint32_t* tile_count, int32_t* out_radii, unsigned long long* stats,
                                  cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const int grid = (N + PF_THREADS - 1) / PF_THREADS;
    const size_t smem = (size_t)(PF_THREADS / 32) * 32 * 3 * cam.KR * sizeof(float);
#define DVS_LAUNCH_PF(D)                                                                                  \
    do {                                                                                                  \
        if (smem > 48 * 1024)                                                                             \
            cudaFuncSetAttribute(preprocess_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem);                                                              \
        preprocess_fwd_kernel<D><<<grid, PF_THREADS, smem, st>>>(cam, N, prm, rec, aux, tile_count,       \
                                                                  out_radii, stats);                      \
    } while (0)
    switch (cam.deg) {
        case 0: DVS_LAUNCH_PF(0); break;
        case 1: DVS_LAUNCH_PF(1); break;
        case 2: DVS_LAUNCH_PF(2); break;
        default: DVS_LAUNCH_PF(3); break;
    }
#undef DVS_LAUNCH_PF
    return cudaGetLastError();
}

}  // namespace dvs
