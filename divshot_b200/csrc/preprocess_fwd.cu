// preprocess_fwd.cu — A1: per-Gaussian forward (3D->2D EWA covariance projection + SH evaluation)
// plus per-tile duplicate counting (two-pass binning) or the duplicate emission itself (single-pass binning).
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
// Roofline: HBM.  Algorithmic bytes per Gaussian: (44+12K) read + 48 (record) + 16 (aux) written.
// One thread per Gaussian for the geometry; each warp's six parameter rows (incl. the 12*(K-1)-byte SH rows)
// are staged in shared memory by ONE lane issuing 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) that complete
// on the warp's mbarrier, and the records / aux words / radii leave by bulk shared->global stores.
#include "common.cuh"
#include "emit.cuh"
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
constexpr int PF_WARPS = PF_THREADS / 32;

// Per-warp staging layout (bytes).  Inputs arrive by ONE elected lane issuing six 1-D bulk copies
// (cp.async.bulk, TMA engine) that complete on the warp's mbarrier; the 48-byte screen records, the
// 16-byte aux words and the radii leave the same way (bulk shared->global stores).
struct PfLayout {
    int means, scales, quats, opac, sh0, shN, rec, aux, radii, total;
    __host__ __device__ explicit PfLayout(int row_floats) {
        int o = 0;
        means = o; o += 32 * 12;
        scales = o; o += 32 * 12;
        quats = o; o += 32 * 16;
        opac = o; o += 32 * 4;
        sh0 = o; o += 32 * 12;
        // the outgoing records / aux words / radii reuse the SH rows once those have been consumed
        shN = o;
        rec = o; aux = rec + 32 * 48; radii = aux + 32 * 16;
        const int out_bytes = 32 * (48 + 16 + 4), sh_bytes = 32 * row_floats * 4;
        o += sh_bytes > out_bytes ? sh_bytes : out_bytes;
        total = (o + 127) & ~127;
    }
};

template <int DEG>
__global__ void __launch_bounds__(PF_THREADS)
preprocess_fwd_kernel(Cam cam, int N, Params prm, float4* __restrict__ rec, uint4* __restrict__ aux,
                      uint32_t* __restrict__ tile_count, int32_t* __restrict__ out_radii,
                      unsigned long long* __restrict__ stats /* [0]=V, [1]=D */, FusedEmit fe) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    extern __shared__ __align__(128) unsigned char pf_smem[];
    const int KR = cam.KR;
    const int row = 3 * KR;  // floats of shN per Gaussian
    const PfLayout L(row);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * PF_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * PF_THREADS + warp * 32;
    uint64_t* bar = reinterpret_cast<uint64_t*>(pf_smem) + warp;
    unsigned char* base = pf_smem + 64 + (size_t)warp * L.total;
    float* s_means = reinterpret_cast<float*>(base + L.means);
    float* s_scales = reinterpret_cast<float*>(base + L.scales);
    float4* s_quats = reinterpret_cast<float4*>(base + L.quats);
    float* s_opac = reinterpret_cast<float*>(base + L.opac);
    float* s_sh0 = reinterpret_cast<float*>(base + L.sh0);
    float* mysh = reinterpret_cast<float*>(base + L.shN);
    float4* s_rec = reinterpret_cast<float4*>(base + L.rec);
    uint4* s_aux = reinterpret_cast<uint4*>(base + L.aux);
    int32_t* s_radii = reinterpret_cast<int32_t*>(base + L.radii);
    if (warp_first >= N) return;
    const int nrows = min(32, N - warp_first);
    const bool full = nrows == 32;  // full warps use the bulk-copy path, the single tail warp plain loads

    if (full) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            const uint32_t bytes = 32u * (12u + 12u + 16u + 4u + 12u) + (K > 1 ? 32u * (uint32_t)row * 4u : 0u);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(s_means, prm.means3D + 3 * (size_t)warp_first, 32 * 12, bar);
            bulk_g2s(s_scales, prm.scales + 3 * (size_t)warp_first, 32 * 12, bar);
            bulk_g2s(s_quats, prm.quats + 4 * (size_t)warp_first, 32 * 16, bar);
            bulk_g2s(s_opac, prm.opacities + warp_first, 32 * 4, bar);
            bulk_g2s(s_sh0, prm.sh0 + 3 * (size_t)warp_first, 32 * 12, bar);
            if (K > 1) bulk_g2s(mysh, prm.shN + (size_t)warp_first * row, 32u * (uint32_t)row * 4u, bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
    } else {
        for (int t = lane; t < nrows * 3; t += 32) {
            s_means[t] = __ldg(prm.means3D + 3 * (size_t)warp_first + t);
            s_scales[t] = __ldg(prm.scales + 3 * (size_t)warp_first + t);
            s_sh0[t] = __ldg(prm.sh0 + 3 * (size_t)warp_first + t);
        }
        if (lane < nrows) {
            s_quats[lane] = __ldg(reinterpret_cast<const float4*>(prm.quats) + warp_first + lane);
            s_opac[lane] = __ldg(prm.opacities + warp_first + lane);
        }
        if (K > 1)
            for (int t = lane; t < nrows * row; t += 32) mysh[t] = __ldg(prm.shN + (size_t)warp_first * row + t);
        __syncwarp();
    }

    bool visible = false;
    uint32_t tiles = 0;
    int minx = 0, miny = 0, maxx = 0, maxy = 0;
    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0;
    uint4 ax = make_uint4(0u, 0u, 0u, 0u);
    int rad = 0;
    if (i < N) {
        const float px = s_means[3 * lane], py = s_means[3 * lane + 1], pz = s_means[3 * lane + 2];
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
                const float a0 = s_scales[3 * lane], a1 = s_scales[3 * lane + 1], a2 = s_scales[3 * lane + 2];
                const float4 qq = s_quats[lane];
                const float oo = s_opac[lane];
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
            const float ca0 = fmaf(U00, T00, fmaf(U01, T01, U02 * T02));
            const float cb = fmaf(U00, T10, fmaf(U01, T11, U02 * T12));
            const float cc0 = fmaf(U10, T10, fmaf(U11, T11, U12 * T12));
            const float ca = ca0 + 0.3f, cc = cc0 + 0.3f;
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
            const int rminx = min(cam.gx, max(0, (int)((mx - radf) * 0.0625f)));
            const int rminy = min(cam.gy, max(0, (int)((my - radf) * 0.0625f)));
            const int rmaxx = min(cam.gx, max(0, (int)((mx + radf + 15.0f) * 0.0625f)));
            const int rmaxy = min(cam.gy, max(0, (int)((my + radf + 15.0f) * 0.0625f)));
            const long long area = (long long)(rmaxx - rminx) * (long long)(rmaxy - rminy);
            if (area <= 0) break;
            minx = rminx; miny = rminy; maxx = rmaxx; maxy = rmaxy;
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
                float acc = bas[0] * s_sh0[3 * lane + ch];
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
            if (cam.flags & DVS_FLAG_ANTIALIAS) {  // mip-splatting opacity compensation (gsplat_vs.hlsl:296-301,373)
                const float det0 = fmaf(ca0, cc0, -(cb * cb));
                o = o * sqrtf(fmaxf(0.0f, det0 / det));
            }
            const float lo = log2f(o);
            q0 = make_float4(mx, my, (-0.5f * LOG2E) * cA, (-0.5f * LOG2E) * cC);
            q1 = make_float4((-LOG2E) * cB, lo, col[0], col[1]);
            q2 = make_float4(col[2], t2, __int_as_float(radius), __uint_as_float(tiles | (clamped << 24)));
            // aux: tile rect, SH clamp mask (bits 29..31 of .y) and the depth key — all the emission kernel and
            // the preprocess backward need, so neither touches the 48-byte records
            ax = make_uint4((uint32_t)minx | ((uint32_t)miny << 16),
                            (uint32_t)maxx | ((uint32_t)maxy << 16) | (clamped << 29), __float_as_uint(t2), 0u);
        } while (false);
    }
    __syncwarp();  // every lane is done with its SH row: the region now becomes the output staging
    s_rec[3 * lane] = q0; s_rec[3 * lane + 1] = q1; s_rec[3 * lane + 2] = q2;
    s_aux[lane] = ax;
    s_radii[lane] = rad;
    // ---- records / aux / radii leave through the TMA engine (full warps) ----
    if (full) {
        fence_proxy_async();  // my generic-proxy smem writes -> visible to the async proxy
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(rec + 3 * (size_t)warp_first, s_rec, 32 * 48);
            bulk_s2g(aux + warp_first, s_aux, 32 * 16);
            if (out_radii) bulk_s2g(out_radii + warp_first, s_radii, 32 * 4);
            bulk_commit();
        }
    } else {
        __syncwarp();
        if (i < N) {
            float4* r = rec + 3 * (size_t)i;
            r[0] = s_rec[3 * lane]; r[1] = s_rec[3 * lane + 1]; r[2] = s_rec[3 * lane + 2];
            aux[i] = s_aux[lane];
            if (out_radii) out_radii[i] = s_radii[lane];
        }
    }
    if (fe.bin_stride) {
        // single-pass binning: emit the duplicates right here (the slot-claiming atomic is also the count);
        // the tile bins have a fixed stride sized from an earlier forward (see api.cu)
        const CullParams cp = cull_params(q0, q1);
        warp_emit(cam.gx, i, minx, miny, maxx - minx, visible ? (int)tiles : 0, __float_as_uint(q2.y), cp,
                  fe.tile_cursor, fe.bins, fe.bin_stride, 0u, fe.overflow_word, fe.tight != 0u);
    } else if (visible) {
        // two-pass binning: per-tile duplicate counts (RED, no return); overlaps the bulk stores
        for (int y = miny; y < maxy; y++)
            for (int x = minx; x < maxx; x++) atomicAdd(tile_count + (size_t)(y * cam.gx + x) * TILE_CTR_STRIDE, 1u);
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
    if (full && lane == 0) bulk_wait_read0();  // shared memory must outlive the bulk stores' reads
}

cudaError_t launch_preprocess_fwd(const Cam& cam, int N, const Params& prm, float4* rec, uint4* aux,
                                  uint32_t* tile_count, int32_t* out_radii, unsigned long long* stats,
                                  const FusedEmit& fe, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const int grid = (N + PF_THREADS - 1) / PF_THREADS;
    const size_t smem = 64 + (size_t)PF_WARPS * PfLayout(3 * cam.KR).total;
#define DVS_LAUNCH_PF(D)                                                                                  \
    do {                                                                                                  \
        if (smem > 48 * 1024)                                                                             \
            cudaFuncSetAttribute(preprocess_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem);                                                              \
        preprocess_fwd_kernel<D><<<grid, PF_THREADS, smem, st>>>(cam, N, prm, rec, aux, tile_count,       \
                                                                  out_radii, stats, fe);                  \
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

// =====================================================================================================================
// 2DGS ("surfel") per-Gaussian forward — GaussianTrainConfig::modelType = 1 (main.cpp:28, gs_train.cpp:68,
// docs/userGuide.md:38; DIVSHOT's own 2DGS rasterizer is in the closed plugin, so the algorithm is the published one:
// Huang et al., "2D Gaussian Splatting for Geometrically Accurate Radiance Fields", restated as S.1 in oracle/dvs_oracle.c).
// A Gaussian is a flat disk with tangents R[:,0], R[:,1] and scales (s_u, s_v); the homography
//     M = Npix * Proj * [ s_u t_u | s_v t_v | p ; 0 0 1 ]          (rows Tu, Tv, Tw: homogeneous PIXEL coordinates of (u, v, 1))
// is the surfel's screen record.  Bounds (3 sigma) come from M; tile rect, depth key, culling and colour are the 3DGS ones,
// so binning / sorting run unchanged on the `aux` words and on a 3DGS-shaped stand-in record (mean2D, colour, depth, radius,
// tiles; zero conic = "no sub-tile culling").  Same TU as A1 because it is the same contract: literal operation sequence,
// -fmad=false, bit-identical radii / rects / depth keys / centres / colours to the oracle.  One thread per Gaussian, plain
// loads: this variant is built for correctness first (DESIGN.md section 9 lists what a tuned version would change).
// rec2 (64 B): {Tu.x, Tu.y, Tu.z, Tv.x} {Tv.y, Tv.z, Tw.x, Tw.y} {Tw.z, cx, cy, opacity} {r, g, b, depth}
// =====================================================================================================================
template <int DEG>
__global__ void __launch_bounds__(128)
surfel_preprocess_fwd_kernel(Cam cam, int N, Params prm, float4* __restrict__ rec, float4* __restrict__ rec2,
                             float4* __restrict__ cull2 /* [N][2]: the sub-tile cull ellipse, in the layout cull_params() reads */,
                             uint4* __restrict__ aux, uint32_t* __restrict__ tile_count, int32_t* __restrict__ out_radii,
                             unsigned long long* __restrict__ stats) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool visible = false;
    uint32_t tiles = 0;
    if (i < N) {
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0;
        float4 s0r = q0, s1r = q0, s2r = q0, s3r = q0;
        float4 k0 = q0, k1 = q0;  // zero conic: "no culling" (cull_params: every box passes)
        uint4 ax = make_uint4(0u, 0u, 0u, 0u);
        int rad = 0;
        const float px = prm.means3D[3 * (size_t)i], py = prm.means3D[3 * (size_t)i + 1], pz = prm.means3D[3 * (size_t)i + 2];
        const float* V = cam.view;
        const float* P = cam.proj;
        const float t2 = fmaf(V[2], px, fmaf(V[6], py, fmaf(V[10], pz, V[14])));
        do {
            if (t2 <= 0.2f) break;
            float s[3], qr, qx, qy, qz, o;
            {
                const float a0 = prm.scales[3 * (size_t)i], a1 = prm.scales[3 * (size_t)i + 1], a2 = prm.scales[3 * (size_t)i + 2];
                const float4 qq = reinterpret_cast<const float4*>(prm.quats)[i];
                const float oo = prm.opacities[i];
                if (cam.flags & DVS_FLAG_INPUT_ACTIVATED) {
                    s[0] = cam.scale_modifier * a0; s[1] = cam.scale_modifier * a1; s[2] = cam.scale_modifier * a2;
                    qr = qq.x; qx = qq.y; qy = qq.z; qz = qq.w;
                    o = oo;
                } else {
                    s[0] = cam.scale_modifier * det_expf(a0);
                    s[1] = cam.scale_modifier * det_expf(a1);
                    s[2] = cam.scale_modifier * det_expf(a2);
                    const float n2 = fmaf(qq.x, qq.x, fmaf(qq.y, qq.y, fmaf(qq.z, qq.z, qq.w * qq.w)));
                    const float inv = 1.0f / sqrtf(n2);
                    qr = qq.x * inv; qx = qq.y * inv; qy = qq.z * inv; qz = qq.w * inv;
                    o = 1.0f / (1.0f + det_expf(-oo));
                }
            }
            float R[3][3];
            R[0][0] = fmaf(-2.0f, fmaf(qz, qz, qy * qy), 1.0f);
            R[0][1] = 2.0f * fmaf(qx, qy, -(qr * qz));
            R[0][2] = 2.0f * fmaf(qx, qz, qr * qy);
            R[1][0] = 2.0f * fmaf(qx, qy, qr * qz);
            R[1][1] = fmaf(-2.0f, fmaf(qz, qz, qx * qx), 1.0f);
            R[1][2] = 2.0f * fmaf(qy, qz, -(qr * qx));
            R[2][0] = 2.0f * fmaf(qx, qz, -(qr * qy));
            R[2][1] = 2.0f * fmaf(qy, qz, qr * qx);
            R[2][2] = fmaf(-2.0f, fmaf(qy, qy, qx * qx), 1.0f);
            // homography rows: T[j] = Tu_j, T[3 + j] = Tv_j, T[6 + j] = Tw_j for the columns j = u, v, 1
            const float hw = 0.5f * (float)cam.W, hh = 0.5f * (float)cam.H;
            const float ow = 0.5f * (float)(cam.W - 1), oh = 0.5f * (float)(cam.H - 1);
            float T[9];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                float v0, v1, v2, w1;
                if (j < 2) { v0 = R[0][j] * s[j]; v1 = R[1][j] * s[j]; v2 = R[2][j] * s[j]; w1 = 0.0f; }
                else { v0 = px; v1 = py; v2 = pz; w1 = 1.0f; }
                const float cx = fmaf(P[0], v0, fmaf(P[4], v1, fmaf(P[8], v2, P[12] * w1)));
                const float cy = fmaf(P[1], v0, fmaf(P[5], v1, fmaf(P[9], v2, P[13] * w1)));
                const float cw = fmaf(P[3], v0, fmaf(P[7], v1, fmaf(P[11], v2, P[15] * w1)));
                T[j] = fmaf(hw, cx, ow * cw);
                T[3 + j] = fmaf(hh, cy, oh * cw);
                T[6 + j] = cw;
            }
            const float tp0 = 9.0f, tp1 = 9.0f, tp2 = -1.0f;
            const float dist = fmaf(tp0 * T[6], T[6], fmaf(tp1 * T[7], T[7], tp2 * T[8] * T[8]));
            if (dist == 0.0f) break;
            const float f0 = tp0 / dist, f1 = tp1 / dist, f2 = tp2 / dist;
            const float cx = fmaf(f0 * T[0], T[6], fmaf(f1 * T[1], T[7], f2 * T[2] * T[8]));
            const float cy = fmaf(f0 * T[3], T[6], fmaf(f1 * T[4], T[7], f2 * T[5] * T[8]));
            const float qxx = fmaf(f0 * T[0], T[0], fmaf(f1 * T[1], T[1], f2 * T[2] * T[2]));
            const float qyy = fmaf(f0 * T[3], T[3], fmaf(f1 * T[4], T[4], f2 * T[5] * T[5]));
            const float ex = sqrtf(fmaxf(1e-4f, fmaf(cx, cx, -qxx))), ey = sqrtf(fmaxf(1e-4f, fmaf(cy, cy, -qyy)));
            const float rad_f = ceilf(fmaxf(fmaxf(ex, ey), 3.0f * 0.707106f));
            if (!(rad_f < 1.0e9f)) break;  // degenerate homography (NaN / inf extent)
            const int radius = (int)rad_f;
            const float radf = (float)radius;
            const int rminx = min(cam.gx, max(0, (int)((cx - radf) * 0.0625f)));
            const int rminy = min(cam.gy, max(0, (int)((cy - radf) * 0.0625f)));
            const int rmaxx = min(cam.gx, max(0, (int)((cx + radf + 15.0f) * 0.0625f)));
            const int rmaxy = min(cam.gy, max(0, (int)((cy + radf + 15.0f) * 0.0625f)));
            const long long area = (long long)(rmaxx - rminx) * (long long)(rmaxy - rminy);
            if (area <= 0) break;
            float d0 = px - cam.campos[0], d1 = py - cam.campos[1], d2 = pz - cam.campos[2];
            const float len = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
            const float linv = 1.0f / len;
            d0 *= linv; d1 *= linv; d2 *= linv;
            float bas[16];
            sh_basis_dev(DEG, d0, d1, d2, bas);
            float col[3];
            uint32_t clamped = 0;
            const float* myrow = prm.shN + (size_t)i * 3 * cam.KR;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float acc = bas[0] * prm.sh0[3 * (size_t)i + ch];
#pragma unroll
                for (int k = 1; k < K; k++) acc = fmaf(bas[k], myrow[3 * (k - 1) + ch], acc);
                acc += 0.5f;
                if (acc < 0.0f) clamped |= 1u << ch;
                col[ch] = fmaxf(acc, 0.0f);
            }
            visible = true;
            tiles = (uint32_t)area;
            rad = radius;
            // 3DGS-shaped stand-in (zero conic: every sub-tile box passes) for the emission kernel and the debug unpackers
            q0 = make_float4(cx, cy, 0.f, 0.f);
            q1 = make_float4(0.f, log2f(o), col[0], col[1]);
            q2 = make_float4(col[2], t2, __int_as_float(radius), __uint_as_float(tiles | (clamped << 24)));
            s0r = make_float4(T[0], T[1], T[2], T[3]);
            s1r = make_float4(T[4], T[5], T[6], T[7]);
            s2r = make_float4(T[8], cx, cy, o);
            s3r = make_float4(col[0], col[1], col[2], t2);
            // ---- sub-tile cull ellipse (ours; outside the bit-exact contract, conservative by construction + margins) ----
            // alpha >= 1/255 needs rho = min(rho3d, rho2d) <= m2 = 2 ln(255 o).  {rho3d <= m2} is the projected disk u^2 + v^2 <=
            // m2: an ellipse on screen as long as the disk lies in front of the camera plane (dist_k < 0), with centre k and
            // shape matrix S (the pixel set is (x - k)^T S^-1 (x - k) <= 1) given by the DUAL conic M diag(m2, m2, -1) M^T —
            // the same well-conditioned expressions as the 3-sigma bounds above, plus the xy term; they are evaluated in pixel
            // coordinates relative to the 3-sigma centre so that the products stay small.  {rho2d <= m2} is the disk of radius
            // sqrt(m2 / 2) around the 3-sigma centre.  The ellipse with shape S + r^2 I, r = that radius + |k - centre|, contains
            // both; its conic is what emission tests the tile's eight 8x4-pixel boxes against (emit.cuh: sub_tile_mask).
            {
                const float m2 = 2.0f * logf(255.0f * o) * 1.0001f + 1e-3f;
                if (!(m2 > 0.0f)) {
                    k0 = make_float4(cx, cy, -1.0f, -1.0f);
                    k1 = make_float4(0.f, ALPHA_MIN_LOG2 - 1.0f, 0.f, 0.f);  // m < 0: the mask is empty, no pixel can blend it
                } else {
                    const float dist_k = m2 * (T[6] * T[6] + T[7] * T[7]) - T[8] * T[8];
                    if (dist_k < -1e-6f * T[8] * T[8]) {
                        const float u0 = fmaf(-cx, T[6], T[0]), u1 = fmaf(-cx, T[7], T[1]), u2 = fmaf(-cx, T[8], T[2]);
                        const float v0 = fmaf(-cy, T[6], T[3]), v1 = fmaf(-cy, T[7], T[4]), v2 = fmaf(-cy, T[8], T[5]);
                        const float g0 = m2 / dist_k, g2 = -1.0f / dist_k;
                        const float kx = g0 * (u0 * T[6] + u1 * T[7]) + g2 * u2 * T[8];
                        const float ky = g0 * (v0 * T[6] + v1 * T[7]) + g2 * v2 * T[8];
                        float hx = kx * kx - (g0 * (u0 * u0 + u1 * u1) + g2 * u2 * u2);
                        float hy = ky * ky - (g0 * (v0 * v0 + v1 * v1) + g2 * v2 * v2);
                        float hxy = kx * ky - (g0 * (u0 * v0 + u1 * v1) + g2 * u2 * v2);
                        hx = fmaxf(hx, 0.0f); hy = fmaxf(hy, 0.0f);
                        const float lim = sqrtf(hx * hy);
                        hxy = fminf(fmaxf(hxy, -lim), lim);
                        const float r = sqrtf(0.5f * m2) + sqrtf(kx * kx + ky * ky) + 0.25f;
                        const float sxx = fmaf(hx, 1.002f, r * r), syy = fmaf(hy, 1.002f, r * r), sxy = hxy * 1.002f;
                        const float det = sxx * syy - sxy * sxy;
                        // the axis-aligned alternative (semi-axes sqrt(2) x the half sizes of the rectangle around both sets):
                        // smaller when the two centres are far apart relative to the ellipse (strong perspective)
                        const float exk = sqrtf(fmaxf(1e-4f, hx)), eyk = sqrtf(fmaxf(1e-4f, hy)), rf = sqrtf(0.5f * m2);
                        const float xlo = fminf(kx - exk, -rf), xhi = fmaxf(kx + exk, rf);
                        const float ylo = fminf(ky - eyk, -rf), yhi = fmaxf(ky + eyk, rf);
                        const float sx = 1.4143f * (0.5f * (xhi - xlo) * 1.001f + 0.05f), sy = 1.4143f * (0.5f * (yhi - ylo) * 1.001f + 0.05f);
                        if (sxx < 1.0e12f && syy < 1.0e12f && det > 0.0f && det < sx * sx * sy * sy) {
                            const float idet = 1.0f / det;
                            k0 = make_float4(cx + kx, cy + ky, -syy * idet, -sxx * idet);
                            k1 = make_float4(2.0f * sxy * idet, ALPHA_MIN_LOG2 + 1.0f, 0.f, 0.f);  // test: d^T (S + r^2 I)^-1 d <= ~1
                        } else if (sx < 1.0e6f && sy < 1.0e6f) {  // (NaN / inf bounds: keep "no culling")
                            k0 = make_float4(cx + 0.5f * (xlo + xhi), cy + 0.5f * (ylo + yhi), -1.0f / (sx * sx), -1.0f / (sy * sy));
                            k1 = make_float4(0.f, ALPHA_MIN_LOG2 + 1.0f, 0.f, 0.f);  // test: dx^2 / sx^2 + dy^2 / sy^2 <= ~1
                        }
                    }
                }
            }
            ax = make_uint4((uint32_t)rminx | ((uint32_t)rminy << 16), (uint32_t)rmaxx | ((uint32_t)rmaxy << 16) | (clamped << 29),
                            __float_as_uint(t2), 0u);
            for (int y = rminy; y < rmaxy; y++)
                for (int x = rminx; x < rmaxx; x++) atomicAdd(tile_count + (size_t)(y * cam.gx + x) * TILE_CTR_STRIDE, 1u);
        } while (false);
        float4* r = rec + 3 * (size_t)i;
        r[0] = q0; r[1] = q1; r[2] = q2;
        float4* r2 = rec2 + 4 * (size_t)i;
        r2[0] = s0r; r2[1] = s1r; r2[2] = s2r; r2[3] = s3r;
        cull2[2 * (size_t)i] = k0; cull2[2 * (size_t)i + 1] = k1;
        aux[i] = ax;
        if (out_radii) out_radii[i] = rad;
    }
    const unsigned vm = __ballot_sync(0xffffffffu, visible);
    uint32_t tsum = tiles;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, off);
    if (lane == 0 && vm) {
        atomicAdd(stats + 0, (unsigned long long)__popc(vm));
        atomicAdd(stats + 1, (unsigned long long)tsum);
    }
}

cudaError_t launch_surfel_preprocess_fwd(const Cam& cam, int N, const Params& prm, float4* rec, float4* rec2, float4* cull2, uint4* aux,
                                         uint32_t* tile_count, int32_t* out_radii, unsigned long long* stats, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const int grid = (N + 127) / 128;
    switch (cam.deg) {
        case 0: surfel_preprocess_fwd_kernel<0><<<grid, 128, 0, st>>>(cam, N, prm, rec, rec2, cull2, aux, tile_count, out_radii, stats); break;
        case 1: surfel_preprocess_fwd_kernel<1><<<grid, 128, 0, st>>>(cam, N, prm, rec, rec2, cull2, aux, tile_count, out_radii, stats); break;
        case 2: surfel_preprocess_fwd_kernel<2><<<grid, 128, 0, st>>>(cam, N, prm, rec, rec2, cull2, aux, tile_count, out_radii, stats); break;
        default: surfel_preprocess_fwd_kernel<3><<<grid, 128, 0, st>>>(cam, N, prm, rec, rec2, cull2, aux, tile_count, out_radii, stats); break;
    }
    return cudaGetLastError();
}

}  // namespace dvs
