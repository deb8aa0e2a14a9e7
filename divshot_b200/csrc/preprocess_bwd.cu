// preprocess_bwd.cu — A8: per-Gaussian backward (screen-space gradients -> stored parameters).
//
// Replaces `computeCov2DCUDA` (backward) + `preprocessCUDA` (backward) of the absent gsplatrast operator
// (forward counterparts corroborated in-tree at
// diverse/assets/shaders/gaussian/gsplat_intersect.hlsl:61-134 and gsplat_sh.hlsl:64-103), fused with
// the activation chain rule (exp / sigmoid / quaternion normalisation — gaussian_model.cpp:145-157) so the
// gradients are w.r.t. the stored raw parameters, and with the per-splat constant factors the
// compositing backward leaves out (see render_bwd.cu).
//
// Roofline: HBM.  Per Gaussian: reads 48 B screen-gradient record + 16 B of the screen record +
// (44+12K) B parameters (visible ones), writes (44+12K) B dense gradient; re-zeroes the
// screen-gradient record it consumed.  SH rows are staged per warp through shared memory so both
// the 12(K-1)-byte loads and stores are 128-bit coalesced.
#include "common.cuh"
#include "kernels.h"

namespace dvs {

constexpr int PB_THREADS = 128;
constexpr int PB_WARPS = PB_THREADS / 32;

// Per-warp staging layout (bytes).  Inputs (parameters, aux words, screen-gradient records) arrive through
// 1-D bulk copies on the warp's mbarrier; the dense gradients are written in place over the consumed
// inputs and leave through bulk shared->global stores, as does the re-zeroed screen-gradient block.
struct PbLayout {
    int means, scales, quats, opac, sh0, shN, sgrad, aux, total;
    __host__ __device__ explicit PbLayout(int row_floats) {
        int o = 0;
        means = o; o += 32 * 12;
        scales = o; o += 32 * 12;
        quats = o; o += 32 * 16;
        opac = o; o += 32 * 4;
        sh0 = o; o += 32 * 12;
        shN = o; o += 32 * row_floats * 4;
        sgrad = o; o += 32 * 48;
        aux = o; o += 32 * 16;
        total = (o + 127) & ~127;
    }
};

template <int DEG>
__global__ void __launch_bounds__(PB_THREADS)
preprocess_bwd_kernel(Cam cam, int N, Params prm, const uint4* __restrict__ aux, float4* __restrict__ sgrad,
                      Grads g, uint32_t flags) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    extern __shared__ __align__(128) unsigned char pb_smem[];
    const int KR = cam.KR;
    const int row = 3 * KR;
    const PbLayout L(row);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * PB_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * PB_THREADS + warp * 32;
    const bool accumulate = flags & DVS_FLAG_ACCUMULATE;
    // DVS_FLAG_SKIP_SHN_GRAD: dL/dshN is not written here (180 of the 236 B per Gaussian of this kernel's output at degree 3) —
    // the fused multi-GPU exchange forms the SUMMED dL/dshN of all views from their dL/dsh0 (collective.cu)
    const bool write_shn = !(flags & DVS_FLAG_SKIP_SHN_GRAD) || accumulate;
    uint64_t* bar = reinterpret_cast<uint64_t*>(pb_smem) + warp;
    unsigned char* base = pb_smem + 64 + (size_t)warp * L.total;
    float* s_means = reinterpret_cast<float*>(base + L.means);
    float* s_scales = reinterpret_cast<float*>(base + L.scales);
    float4* s_quats = reinterpret_cast<float4*>(base + L.quats);
    float* s_opac = reinterpret_cast<float*>(base + L.opac);
    float* s_sh0 = reinterpret_cast<float*>(base + L.sh0);
    float* mysh = reinterpret_cast<float*>(base + L.shN);
    float4* s_sg = reinterpret_cast<float4*>(base + L.sgrad);
    uint4* s_aux = reinterpret_cast<uint4*>(base + L.aux);
    if (warp_first >= N) return;
    const int nrows = min(32, N - warp_first);
    const int nflt = nrows * row;
    const bool full = nrows == 32;

    if (full) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            const uint32_t bytes = 32u * (12u + 12u + 16u + 4u + 48u + 16u) + (K > 1 ? 32u * (uint32_t)row * 4u : 0u);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(s_aux, aux + warp_first, 32 * 16, bar);
            bulk_g2s(s_sg, sgrad + 3 * (size_t)warp_first, 32 * 48, bar);
            bulk_g2s(s_means, prm.means3D + 3 * (size_t)warp_first, 32 * 12, bar);
            bulk_g2s(s_scales, prm.scales + 3 * (size_t)warp_first, 32 * 12, bar);
            bulk_g2s(s_quats, prm.quats + 4 * (size_t)warp_first, 32 * 16, bar);
            bulk_g2s(s_opac, prm.opacities + warp_first, 32 * 4, bar);
            if (K > 1) bulk_g2s(mysh, prm.shN + (size_t)warp_first * row, 32u * (uint32_t)row * 4u, bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
    } else {
        for (int t = lane; t < nrows * 3; t += 32) {
            s_means[t] = __ldg(prm.means3D + 3 * (size_t)warp_first + t);
            s_scales[t] = __ldg(prm.scales + 3 * (size_t)warp_first + t);
        }
        for (int t = lane; t < nrows * 3; t += 32) s_sg[t] = sgrad[3 * (size_t)warp_first + t];
        if (lane < nrows) {
            s_quats[lane] = __ldg(reinterpret_cast<const float4*>(prm.quats) + warp_first + lane);
            s_opac[lane] = __ldg(prm.opacities + warp_first + lane);
            s_aux[lane] = __ldg(aux + warp_first + lane);
        }
        if (K > 1)
            for (int t = lane; t < nflt; t += 32) mysh[t] = __ldg(prm.shN + (size_t)warp_first * row + t);
        __syncwarp();
    }

    // visibility (tile rect area > 0) and SH clamp mask come from the aux word
    uint4 ax = make_uint4(0u, 0u, 0u, 0u);
    if (i < N) ax = s_aux[lane];
    const bool vis = ((ax.y & 0xffffu) > (ax.x & 0xffffu)) && (((ax.y >> 16) & 0x1fffu) > (ax.x >> 16));
    const bool any_vis = __any_sync(0xffffffffu, vis);

    float dmean0 = 0.f, dmean1 = 0.f, dmean2 = 0.f;
    float dsc0 = 0.f, dsc1 = 0.f, dsc2 = 0.f;
    float dq0 = 0.f, dq1 = 0.f, dq2 = 0.f, dq3 = 0.f;
    float dop = 0.f;
    float dsh0[3] = {0.f, 0.f, 0.f};
    float gm2x = 0.f, gm2y = 0.f, gabx = 0.f, gaby = 0.f;
    float* myrow = mysh + lane * row;

    // consume my screen-gradient record, leave zeros behind (the whole block is stored back below)
    const float4 sg0 = s_sg[3 * lane], sg1 = s_sg[3 * lane + 1], sg2 = s_sg[3 * lane + 2];
    {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        s_sg[3 * lane] = z4; s_sg[3 * lane + 1] = z4; s_sg[3 * lane + 2] = z4;
    }
    if (vis) {
        const uint32_t clamped = ax.y >> 29;
        const float px = s_means[3 * lane], py = s_means[3 * lane + 1], pz = s_means[3 * lane + 2];
        const float a0 = s_scales[3 * lane], a1 = s_scales[3 * lane + 1], a2 = s_scales[3 * lane + 2];
        const float4 qq = s_quats[lane];
        const float oo = s_opac[lane];
        const bool activated = cam.flags & DVS_FLAG_INPUT_ACTIVATED;
        float s0, s1, s2, qr, qx, qy, qz, o, qlen = 1.0f;
        if (activated) {
            s0 = cam.scale_modifier * a0; s1 = cam.scale_modifier * a1; s2 = cam.scale_modifier * a2;
            qr = qq.x; qx = qq.y; qy = qq.z; qz = qq.w; o = oo;
        } else {
            s0 = cam.scale_modifier * expf(a0); s1 = cam.scale_modifier * expf(a1); s2 = cam.scale_modifier * expf(a2);
            qlen = sqrtf(qq.x * qq.x + qq.y * qq.y + qq.z * qq.z + qq.w * qq.w);
            const float inv = 1.0f / qlen;
            qr = qq.x * inv; qx = qq.y * inv; qy = qq.z * inv; qz = qq.w * inv;
            o = 1.0f / (1.0f + expf(-oo));
        }
        // screen-space sums left by render_bwd.cu (moments of s = dL/dpower over the splat's pixels):
        //   sg0 = {Sx, Sy, Sxx, Sxy}, sg1 = {Syy, S0, sum w*dL/dpix r, g}, sg2 = {b, sum|gx|, sum|gy|, -}
        const float dA = -0.5f * sg0.z, dBh = -0.5f * sg0.w, dC = -0.5f * sg1.x;  // dBh = half the off-diagonal
        const float dL_dopacity = sg1.y / o;
        const float dcol_in[3] = {sg1.z, sg1.w, sg2.x};
        gabx = 0.5f * (float)cam.W * sg2.y; gaby = 0.5f * (float)cam.H * sg2.z;

        const float* V = cam.view;
        const float* Pm = cam.proj;
        const float t0 = V[0] * px + V[4] * py + V[8] * pz + V[12];
        const float t1 = V[1] * px + V[5] * py + V[9] * pz + V[13];
        const float t2 = V[2] * px + V[6] * py + V[10] * pz + V[14];
        // forward recompute: R, M, Sigma, T, cov2D
        const float R[3][3] = {{1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qr * qz), 2.f * (qx * qz + qr * qy)},
                               {2.f * (qx * qy + qr * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qr * qx)},
                               {2.f * (qx * qz - qr * qy), 2.f * (qy * qz + qr * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
        const float sc[3] = {s0, s1, s2};
        float M[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) M[r][k] = R[r][k] * sc[k];
        float S[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) S[r][c] = M[r][0] * M[c][0] + M[r][1] * M[c][1] + M[r][2] * M[c][2];
        const float fx = (float)cam.W / (2.0f * cam.tanfovx), fy = (float)cam.H / (2.0f * cam.tanfovy);
        const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
        const float txtz = t0 / t2, tytz = t1 / t2;
        const float tx = fminf(limx, fmaxf(-limx, txtz)) * t2;
        const float ty = fminf(limy, fmaxf(-limy, tytz)) * t2;
        const float tzi = 1.0f / t2, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float J00 = fx * tzi, J02 = -(fx * tx) * tz2, J11 = fy * tzi, J12 = -(fy * ty) * tz2;
        float Tm[2][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Tm[0][j] = J00 * V[4 * j + 0] + J02 * V[4 * j + 2];
            Tm[1][j] = J11 * V[4 * j + 1] + J12 * V[4 * j + 2];
        }
        float TS[2][3];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int j = 0; j < 3; j++) TS[r][j] = Tm[r][0] * S[0][j] + Tm[r][1] * S[1][j] + Tm[r][2] * S[2][j];
        const float ca0 = TS[0][0] * Tm[0][0] + TS[0][1] * Tm[0][1] + TS[0][2] * Tm[0][2];
        const float cb = TS[0][0] * Tm[1][0] + TS[0][1] * Tm[1][1] + TS[0][2] * Tm[1][2];
        const float cc0 = TS[1][0] * Tm[1][0] + TS[1][1] * Tm[1][1] + TS[1][2] * Tm[1][2];
        const float ca = ca0 + 0.3f, cc = cc0 + 0.3f;
        // 1. conic -> cov2D
        const float det = ca * cc - cb * cb;
        // dL/dmean2D (ndc-scaled): dpower/ddx = -(A dx + B dy) with the conic (A,B,C) = (cc,-cb,ca)/det
        const float det_inv = 1.0f / det;
        const float gmx = -0.5f * (float)cam.W * det_inv * (cc * sg0.x - cb * sg0.y);
        const float gmy = -0.5f * (float)cam.H * det_inv * (ca * sg0.y - cb * sg0.x);
        gm2x = gmx; gm2y = gmy;
        const float kappa = 1.0f / (det * det + 1e-7f);
        float da = kappa * (-cc * cc * dA + 2.0f * cb * cc * dBh + (det - ca * cc) * dC);
        float dc = kappa * (-ca * ca * dC + 2.0f * ca * cb * dBh + (det - ca * cc) * dA);
        float db = kappa * 2.0f * (cb * cc * dA - (det + 2.0f * cb * cb) * dBh + ca * cb * dC);
        if (cam.flags & DVS_FLAG_ANTIALIAS) {
            // opacity' = o * rho, rho = sqrt(max(0, r)), r = det0/det.  S0 = sum s = o' dL/do', so dL/drho = S0 / rho and
            // dL/dr = S0 / (2 r); dL/do (used below) stays S0 / o.
            const float det0 = ca0 * cc0 - cb * cb, r = det0 * det_inv;
            if (r > 0.0f) {
                const float gr = 0.5f * sg1.y / r * det_inv * det_inv;
                da += gr * (cc0 * det - det0 * cc);
                dc += gr * (ca0 * det - det0 * ca);
                db += gr * (-2.0f * cb * (det - det0));
            }
        }
        // 2. cov2D -> Sigma (symmetric 3x3 gradient dS) and -> T -> J -> t
        float dS[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                dS[r][c] = Tm[0][r] * Tm[0][c] * da + 0.5f * (Tm[0][r] * Tm[1][c] + Tm[1][r] * Tm[0][c]) * db +
                           Tm[1][r] * Tm[1][c] * dc;
        float dT[2][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            dT[0][j] = 2.0f * da * TS[0][j] + db * TS[1][j];
            dT[1][j] = 2.0f * dc * TS[1][j] + db * TS[0][j];
        }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            dJ00 += V[4 * j + 0] * dT[0][j];
            dJ02 += V[4 * j + 2] * dT[0][j];
            dJ11 += V[4 * j + 1] * dT[1][j];
            dJ12 += V[4 * j + 2] * dT[1][j];
        }
        const float mxk = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
        const float myk = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
        const float dt0 = mxk * (-fx * tz2) * dJ02;
        const float dt1 = myk * (-fy * tz2) * dJ12;
        const float dt2 = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0f * fx * tx) * tz3 * dJ02 + (2.0f * fy * ty) * tz3 * dJ12;
        dmean0 = V[0] * dt0 + V[1] * dt1 + V[2] * dt2;
        dmean1 = V[4] * dt0 + V[5] * dt1 + V[6] * dt2;
        dmean2 = V[8] * dt0 + V[9] * dt1 + V[10] * dt2;
        // 3. projection
        {
            const float h0 = Pm[0] * px + Pm[4] * py + Pm[8] * pz + Pm[12];
            const float h1 = Pm[1] * px + Pm[5] * py + Pm[9] * pz + Pm[13];
            const float h3 = Pm[3] * px + Pm[7] * py + Pm[11] * pz + Pm[15];
            const float m_w = 1.0f / (h3 + 1e-7f);
            const float mul1 = h0 * m_w * m_w, mul2 = h1 * m_w * m_w;
            dmean0 += (Pm[0] * m_w - Pm[3] * mul1) * gmx + (Pm[1] * m_w - Pm[3] * mul2) * gmy;
            dmean1 += (Pm[4] * m_w - Pm[7] * mul1) * gmx + (Pm[5] * m_w - Pm[7] * mul2) * gmy;
            dmean2 += (Pm[8] * m_w - Pm[11] * mul1) * gmx + (Pm[9] * m_w - Pm[11] * mul2) * gmy;
        }
        // 4. SH
        {
            const float ox = px - cam.campos[0], oy = py - cam.campos[1], oz = pz - cam.campos[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float li = 1.0f / len;
            const float x = ox * li, y = oy * li, z = oz * li;
            const float dcol[3] = {(clamped & 1u) ? 0.f : dcol_in[0], (clamped & 2u) ? 0.f : dcol_in[1],
                                   (clamped & 4u) ? 0.f : dcol_in[2]};
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) dsh0[ch] = 0.28209479177387814f * dcol[ch];
            if (DEG >= 1) {
                // per coefficient k: s_k = sum_ch sh[k][ch]*dcol[ch]; basis b_k and its gradient
                float bk[16], gx_[16], gy_[16], gz_[16];
#pragma unroll
                for (int k = 0; k < 16; k++) { bk[k] = 0.f; gx_[k] = 0.f; gy_[k] = 0.f; gz_[k] = 0.f; }
                const float C1 = 0.4886025119029199f;
                bk[1] = -C1 * y; gy_[1] = -C1;
                bk[2] = C1 * z; gz_[2] = C1;
                bk[3] = -C1 * x; gx_[3] = -C1;
                if (DEG >= 2) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    const float c20 = 1.0925484305920792f, c21 = -1.0925484305920792f, c22 = 0.31539156525252005f,
                                c23 = -1.0925484305920792f, c24 = 0.5462742152960396f;
                    bk[4] = c20 * xy; gx_[4] = c20 * y; gy_[4] = c20 * x;
                    bk[5] = c21 * yz; gy_[5] = c21 * z; gz_[5] = c21 * y;
                    bk[6] = c22 * (2.0f * zz - xx - yy); gx_[6] = c22 * -2.0f * x; gy_[6] = c22 * -2.0f * y; gz_[6] = c22 * 4.0f * z;
                    bk[7] = c23 * xz; gx_[7] = c23 * z; gz_[7] = c23 * x;
                    bk[8] = c24 * (xx - yy); gx_[8] = c24 * 2.0f * x; gy_[8] = c24 * -2.0f * y;
                    if (DEG >= 3) {
                        const float c30 = -0.5900435899266435f, c31 = 2.890611442640554f, c32 = -0.4570457994644658f,
                                    c33 = 0.3731763325901154f, c34 = -0.4570457994644658f, c35 = 1.445305721320277f,
                                    c36 = -0.5900435899266435f;
                        bk[9] = c30 * y * (3.0f * xx - yy); gx_[9] = c30 * 6.0f * xy; gy_[9] = c30 * (3.0f * xx - 3.0f * yy);
                        bk[10] = c31 * xy * z; gx_[10] = c31 * yz; gy_[10] = c31 * xz; gz_[10] = c31 * xy;
                        bk[11] = c32 * y * (4.0f * zz - xx - yy); gx_[11] = c32 * -2.0f * xy;
                        gy_[11] = c32 * (4.0f * zz - xx - 3.0f * yy); gz_[11] = c32 * 8.0f * yz;
                        bk[12] = c33 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy); gx_[12] = c33 * -6.0f * xz;
                        gy_[12] = c33 * -6.0f * yz; gz_[12] = c33 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
                        bk[13] = c34 * x * (4.0f * zz - xx - yy); gx_[13] = c34 * (4.0f * zz - 3.0f * xx - yy);
                        gy_[13] = c34 * -2.0f * xy; gz_[13] = c34 * 8.0f * xz;
                        bk[14] = c35 * z * (xx - yy); gx_[14] = c35 * 2.0f * xz; gy_[14] = c35 * -2.0f * yz;
                        gz_[14] = c35 * (xx - yy);
                        bk[15] = c36 * x * (xx - 3.0f * yy); gx_[15] = c36 * (3.0f * xx - 3.0f * yy); gy_[15] = c36 * -6.0f * xy;
                    }
                }
#pragma unroll
                for (int k = 1; k < K; k++) {
                    const float c0 = myrow[3 * (k - 1)], c1 = myrow[3 * (k - 1) + 1], c2 = myrow[3 * (k - 1) + 2];
                    const float sk = c0 * dcol[0] + c1 * dcol[1] + c2 * dcol[2];
                    ddx = fmaf(gx_[k], sk, ddx); ddy = fmaf(gy_[k], sk, ddy); ddz = fmaf(gz_[k], sk, ddz);
                    myrow[3 * (k - 1)] = bk[k] * dcol[0];      // overwrite the staged coefficients with their gradient
                    myrow[3 * (k - 1) + 1] = bk[k] * dcol[1];
                    myrow[3 * (k - 1) + 2] = bk[k] * dcol[2];
                }
                for (int t = 3 * (K - 1); t < row; t++) myrow[t] = 0.f;  // allocated-but-inactive coefficients
            }
            const float dd = x * ddx + y * ddy + z * ddz;
            dmean0 += (ddx - x * dd) * li;
            dmean1 += (ddy - y * dd) * li;
            dmean2 += (ddz - z * dd) * li;
        }
        // 5. Sigma -> scale, rotation
        {
            float dM[3][3];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int k = 0; k < 3; k++) dM[r][k] = 2.0f * (dS[r][0] * M[0][k] + dS[r][1] * M[1][k] + dS[r][2] * M[2][k]);
            float ds[3], dR[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                ds[k] = R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k];
#pragma unroll
                for (int r = 0; r < 3; r++) dR[r][k] = dM[r][k] * sc[k];
            }
            float dq[4];
            dq[0] = 2.0f * (qz * (dR[1][0] - dR[0][1]) + qy * (dR[0][2] - dR[2][0]) + qx * (dR[2][1] - dR[1][2]));
            dq[1] = 2.0f * (qy * (dR[0][1] + dR[1][0]) + qz * (dR[0][2] + dR[2][0]) + qr * (dR[2][1] - dR[1][2])) -
                    4.0f * qx * (dR[1][1] + dR[2][2]);
            dq[2] = 2.0f * (qx * (dR[0][1] + dR[1][0]) + qr * (dR[0][2] - dR[2][0]) + qz * (dR[1][2] + dR[2][1])) -
                    4.0f * qy * (dR[0][0] + dR[2][2]);
            dq[3] = 2.0f * (qr * (dR[1][0] - dR[0][1]) + qx * (dR[0][2] + dR[2][0]) + qy * (dR[1][2] + dR[2][1])) -
                    4.0f * qz * (dR[0][0] + dR[1][1]);
            if (activated) {
                dsc0 = ds[0] * cam.scale_modifier; dsc1 = ds[1] * cam.scale_modifier; dsc2 = ds[2] * cam.scale_modifier;
                dq0 = dq[0]; dq1 = dq[1]; dq2 = dq[2]; dq3 = dq[3];
                dop = dL_dopacity;
            } else {
                dsc0 = ds[0] * s0; dsc1 = ds[1] * s1; dsc2 = ds[2] * s2;
                const float qd = qr * dq[0] + qx * dq[1] + qy * dq[2] + qz * dq[3];
                const float il = 1.0f / qlen;
                dq0 = (dq[0] - qr * qd) * il; dq1 = (dq[1] - qx * qd) * il;
                dq2 = (dq[2] - qy * qd) * il; dq3 = (dq[3] - qz * qd) * il;
                dop = sg1.y * (1.0f - o);  // (s/o) * o (1-o)
            }
        }
    }
    if (row > 0 && any_vis && (!vis || K == 1))
        for (int t = 0; t < row; t++) myrow[t] = 0.f;

    // ---- write dense gradients ----
    if (full && !accumulate) {
        // in place over the consumed inputs, then out through the TMA engine
        __syncwarp();
        s_means[3 * lane] = dmean0; s_means[3 * lane + 1] = dmean1; s_means[3 * lane + 2] = dmean2;
        s_scales[3 * lane] = dsc0; s_scales[3 * lane + 1] = dsc1; s_scales[3 * lane + 2] = dsc2;
        s_quats[lane] = make_float4(dq0, dq1, dq2, dq3);
        s_opac[lane] = dop;
        s_sh0[3 * lane] = dsh0[0]; s_sh0[3 * lane + 1] = dsh0[1]; s_sh0[3 * lane + 2] = dsh0[2];
        if (row > 0 && !any_vis)
            for (int t = 0; t < row; t++) myrow[t] = 0.f;
        if (g.mean2D) reinterpret_cast<float2*>(g.mean2D)[i] = make_float2(gm2x, gm2y);
        if (g.mean2D_abs) reinterpret_cast<float2*>(g.mean2D_abs)[i] = make_float2(gabx, gaby);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(g.means3D + 3 * (size_t)warp_first, s_means, 32 * 12);
            bulk_s2g(g.scales + 3 * (size_t)warp_first, s_scales, 32 * 12);
            bulk_s2g(g.quats + 4 * (size_t)warp_first, s_quats, 32 * 16);
            bulk_s2g(g.opacities + warp_first, s_opac, 32 * 4);
            bulk_s2g(g.sh0 + 3 * (size_t)warp_first, s_sh0, 32 * 12);
            if (row > 0 && write_shn) bulk_s2g(g.shN + (size_t)warp_first * row, mysh, 32u * (uint32_t)row * 4u);
            if (any_vis) bulk_s2g(sgrad + 3 * (size_t)warp_first, s_sg, 32 * 48);
            bulk_commit();
            bulk_wait_read0();
        }
        return;
    }
    // generic path: tail warp, or DVS_FLAG_ACCUMULATE (read-modify-write)
    if (vis) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        sgrad[3 * (size_t)i] = z4; sgrad[3 * (size_t)i + 1] = z4; sgrad[3 * (size_t)i + 2] = z4;
    }
    if (i < N) {
        float* gm = g.means3D + 3 * (size_t)i;
        float* gs = g.scales + 3 * (size_t)i;
        float4* gq = reinterpret_cast<float4*>(g.quats) + i;
        float* g0 = g.sh0 + 3 * (size_t)i;
        if (accumulate) {
            if (vis) {
                gm[0] += dmean0; gm[1] += dmean1; gm[2] += dmean2;
                gs[0] += dsc0; gs[1] += dsc1; gs[2] += dsc2;
                float4 q = *gq; q.x += dq0; q.y += dq1; q.z += dq2; q.w += dq3; *gq = q;
                g.opacities[i] += dop;
                g0[0] += dsh0[0]; g0[1] += dsh0[1]; g0[2] += dsh0[2];
                if (g.mean2D) { g.mean2D[2 * (size_t)i] += gm2x; g.mean2D[2 * (size_t)i + 1] += gm2y; }
                if (g.mean2D_abs) { g.mean2D_abs[2 * (size_t)i] += gabx; g.mean2D_abs[2 * (size_t)i + 1] += gaby; }
            }
        } else {
            gm[0] = dmean0; gm[1] = dmean1; gm[2] = dmean2;
            gs[0] = dsc0; gs[1] = dsc1; gs[2] = dsc2;
            *gq = make_float4(dq0, dq1, dq2, dq3);
            g.opacities[i] = dop;
            g0[0] = dsh0[0]; g0[1] = dsh0[1]; g0[2] = dsh0[2];
            if (g.mean2D) { g.mean2D[2 * (size_t)i] = gm2x; g.mean2D[2 * (size_t)i + 1] = gm2y; }
            if (g.mean2D_abs) { g.mean2D_abs[2 * (size_t)i] = gabx; g.mean2D_abs[2 * (size_t)i + 1] = gaby; }
        }
    }
    // SH rest gradients: coalesced stores of the staged rows
    if (row > 0 && write_shn) {
        __syncwarp();
        float* dst = g.shN + (size_t)warp_first * row;
        const int nvec = nflt >> 2;
        float4* dst4 = reinterpret_cast<float4*>(dst);
        const float4* s4 = reinterpret_cast<const float4*>(mysh);
        if (!any_vis) {
            if (!accumulate) {
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int v = lane; v < nvec; v += 32) dst4[v] = z4;
                for (int t = (nvec << 2) + lane; t < nflt; t += 32) dst[t] = 0.f;
            }
        } else if (accumulate) {
            for (int v = lane; v < nvec; v += 32) {
                float4 a = dst4[v]; const float4 b = s4[v];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; dst4[v] = a;
            }
            for (int t = (nvec << 2) + lane; t < nflt; t += 32) dst[t] += mysh[t];
        } else {
            for (int v = lane; v < nvec; v += 32) dst4[v] = s4[v];
            for (int t = (nvec << 2) + lane; t < nflt; t += 32) dst[t] = mysh[t];
        }
    }
}

cudaError_t launch_preprocess_bwd(const Cam& cam, int N, const Params& prm, const uint4* aux, float4* sgrad,
                                  const Grads& g, uint32_t flags, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const int grid = (N + PB_THREADS - 1) / PB_THREADS;
    const size_t smem = 64 + (size_t)PB_WARPS * PbLayout(3 * cam.KR).total;
#define DVS_LAUNCH_PB(D)                                                                                 \
    do {                                                                                                 \
        if (smem > 48 * 1024)                                                                            \
            cudaFuncSetAttribute(preprocess_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                 (int)smem);                                                             \
        preprocess_bwd_kernel<D><<<grid, PB_THREADS, smem, st>>>(cam, N, prm, aux, sgrad, g, flags);     \
    } while (0)
    switch (cam.deg) {
        case 0: DVS_LAUNCH_PB(0); break;
        case 1: DVS_LAUNCH_PB(1); break;
        case 2: DVS_LAUNCH_PB(2); break;
        default: DVS_LAUNCH_PB(3); break;
    }
#undef DVS_LAUNCH_PB
    return cudaGetLastError();
}

}  // namespace dvs
