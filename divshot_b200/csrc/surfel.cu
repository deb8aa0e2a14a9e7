// surfel.cu — the 2DGS ("surfel") compositing forward, reverse-walk backward and per-Gaussian backward:
// GaussianTrainConfig::modelType = 1 (application/diverseshot-cli/source/main.cpp:28, gs_train.cpp:68, docs/userGuide.md:38).
//
// DIVSHOT's 2DGS rasterizer is in the same closed plugin as its 3DGS one; the algorithm is the
// published one (Huang et al., "2D Gaussian Splatting for Geometrically Accurate Radiance Fields", SIGGRAPH 2024), restated
// as S.1-S.4 in oracle/dvs_oracle.c and pinned there by closed-form cases and a float64 autograd re-expression.  The
// per-Gaussian forward (homography M, screen bounds) is in preprocess_fwd.cu; binning and sorting are the 3DGS kernels.
//
//   per pixel (x, y) and surfel with rows Tu, Tv, Tw of M, projected centre c, opacity o:
//     k = x Tw - Tu,  l = y Tw - Tv,  pv = k x l  (pv.z == 0: skip),  (u, v) = pv.xy / pv.z,  rho3d = u^2 + v^2
//     rho2d = 2 |c - (x, y)|^2  (object-space low-pass filter),  rho = min(rho3d, rho2d)
//     depth = rho3d <= rho2d ? u Tw.x + v Tw.y + Tw.z : Tw.z  (< 0.2: skip),  alpha = min(0.99, o exp(-rho / 2))
//   then the 3DGS compositing rules (alpha < 1/255 skip, T (1 - alpha) < 1e-4 stop, n_contrib, final_T, background).
//
// Design.  One CTA per 16x16 tile, warp w = the 8x4-pixel sub-rectangle (w & 1, w >> 1), 64-byte records staged 256 per round
// in shared memory.  A warp visits only the surfels whose sub-tile mask bit is set: one ballot per 32 staged list entries,
// walked by find-first-set (the mask is computed at emission from the surfel's projected alpha >= 1/255 ellipse,
// preprocess_fwd.cu).  The pair evaluation uses reciprocals (rcp.approx) instead of IEEE divisions.  The backward also stops at
// the warp's deepest last contributor, sums its 15 per-pair values over the warp by RECURSIVE HALVING (20 shuffles instead of
// 15 x 5; skipped when no lane blended the pair) and sends them with one 128-bit vector reduction per four floats.  The
// per-Gaussian backward stages its shN / dL/dshN rows through shared memory.  Not applied here (DESIGN.md sections 8, 9): the
// 3DGS path's packed fp32x2 arithmetic, predicated straight-line loops and two-phase backward.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dvs {

namespace {
constexpr int SF_THREADS = 256;
constexpr float SF_FILTER_INV_SQ = 2.0f;

struct SurfelPair {
    float u, v, pz, ipz, G, alpha, dx, dy;
    float k0, k1, k2, l0, l1, l2;
    bool use3d;
};
// S.2 for one (pixel, surfel); false = skipped before the alpha test
__device__ __forceinline__ bool surfel_pair(const float4 a, const float4 b, const float4 c, float pxf, float pyf, SurfelPair& r) {
    // a = {Tu.x, Tu.y, Tu.z, Tv.x}  b = {Tv.y, Tv.z, Tw.x, Tw.y}  c = {Tw.z, cx, cy, opacity}
    r.k0 = fmaf(pxf, b.z, -a.x); r.k1 = fmaf(pxf, b.w, -a.y); r.k2 = fmaf(pxf, c.x, -a.z);
    r.l0 = fmaf(pyf, b.z, -a.w); r.l1 = fmaf(pyf, b.w, -b.x); r.l2 = fmaf(pyf, c.x, -b.y);
    const float p0 = fmaf(r.k1, r.l2, -(r.k2 * r.l1));
    const float p1 = fmaf(r.k2, r.l0, -(r.k0 * r.l2));
    const float p2 = fmaf(r.k0, r.l1, -(r.k1 * r.l0));
    if (p2 == 0.0f) return false;
    r.pz = p2;
    r.ipz = rcp_approx(p2);  // (a denormal p2 gives inf / NaN here: rho3d <= rho2d is then false and the low-pass branch is taken)
    r.u = p0 * r.ipz; r.v = p1 * r.ipz;
    const float rho3d = fmaf(r.u, r.u, r.v * r.v);
    r.dx = c.y - pxf; r.dy = c.z - pyf;
    const float rho2d = SF_FILTER_INV_SQ * fmaf(r.dx, r.dx, r.dy * r.dy);
    r.use3d = rho3d <= rho2d;
    const float rho = r.use3d ? rho3d : rho2d;
    const float dep = r.use3d ? fmaf(r.u, b.z, fmaf(r.v, b.w, c.x)) : c.x;
    if (dep < 0.2f) return false;
    r.G = __expf(-0.5f * rho);
    r.alpha = fminf(0.99f, c.w * r.G);
    return true;
}

__global__ void __launch_bounds__(SF_THREADS)
surfel_render_fwd_kernel(Cam cam, const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ plist,
                         const float4* __restrict__ rec2, float* __restrict__ out_color, float* __restrict__ final_T,
                         uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ info) {
    __shared__ float4 s_rec[SF_THREADS * 4];
    __shared__ uint32_t s_mask[SF_THREADS];
    if (info[2]) return;
    const int tile = blockIdx.x;
    const int tx = tile % cam.gx, ty = tile / cam.gx;
    const uint32_t r0 = tile_base[tile], n = tile_base[tile + 1] - r0;
    // warp w owns the 8x4-pixel sub-rectangle (w & 1, w >> 1) of the tile, like the 3DGS compositor: a whole warp skips a
    // surfel whose cull ellipse misses its 32 pixels (sub-tile mask bit w of the list entry, computed at emission)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbit = 1u << warp;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7), py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const float pxf = (float)px, pyf = (float)py;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    for (uint32_t base = 0; base < n; base += SF_THREADS) {
        if (__syncthreads_count(done) == SF_THREADS) break;
        const uint32_t idx = base + threadIdx.x;
        if (idx < n) {
            const uint32_t e = __ldg(plist + r0 + idx);
            s_mask[threadIdx.x] = e & 0xffu;
            if (e & 0xffu) {
                const float4* r = rec2 + 4 * (size_t)(e >> 8);
                s_rec[4 * threadIdx.x] = __ldg(r); s_rec[4 * threadIdx.x + 1] = __ldg(r + 1);
                s_rec[4 * threadIdx.x + 2] = __ldg(r + 2); s_rec[4 * threadIdx.x + 3] = __ldg(r + 3);
            }
        }
        __syncthreads();
        const uint32_t cnt = min((uint32_t)SF_THREADS, n - base);
        if (__all_sync(0xffffffffu, done)) continue;
        // warp-uniform walk: every lane evaluates the pair, finished lanes just stop updating (no per-lane branches around
        // the arithmetic); the warp leaves the round once all 32 pixels are done (checked every 8 visited surfels).
        // the warp's entries of this round as 32-bit masks (one ballot per 32 list entries), walked by find-first-set: the
        // loop costs nothing for the entries whose cull ellipse misses this warp's pixels
        uint32_t visited = 0;
        bool stop = false;
        for (uint32_t g0 = 0; g0 < cnt && !stop; g0 += 32) {
            const uint32_t jj = g0 + lane;
            uint32_t m = __ballot_sync(0xffffffffu, jj < cnt && (s_mask[jj] & wbit));
            while (m) {
                const uint32_t j = g0 + (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                SurfelPair q;
                bool ok = surfel_pair(s_rec[4 * j], s_rec[4 * j + 1], s_rec[4 * j + 2], pxf, pyf, q);
                ok = ok && !done && q.alpha >= 1.0f / 255.0f;
                const float test_T = T * (1.0f - q.alpha);
                if (ok && test_T < 1e-4f) { done = true; ok = false; }
                if (ok) {
                    const float w = q.alpha * T;
                    const float4 col = s_rec[4 * j + 3];
                    C0 = fmaf(col.x, w, C0); C1 = fmaf(col.y, w, C1); C2 = fmaf(col.z, w, C2);
                    T = test_T;
                    last = base + j + 1u;
                }
                if ((++visited & 7u) == 0u && __all_sync(0xffffffffu, done)) { stop = true; break; }
            }
        }
    }
    if (inside) {
        const size_t P = (size_t)cam.W * cam.H, pix = (size_t)py * cam.W + px;
        float b0 = cam.bg[0], b1 = cam.bg[1], b2 = cam.bg[2];
        if (cam.bg_image) { b0 = __ldg(cam.bg_image + pix); b1 = __ldg(cam.bg_image + P + pix); b2 = __ldg(cam.bg_image + 2 * P + pix); }
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = fmaf(T, b0, C0);
        out_color[P + pix] = fmaf(T, b1, C1);
        out_color[2 * P + pix] = fmaf(T, b2, C2);
    }
}

// S.3: reverse walk.  sgrad2 record per surfel (64 B, zero on entry, re-zeroed by the per-Gaussian backward):
//   {dTu.x, dTu.y, dTu.z, dTv.x} {dTv.y, dTv.z, dTw.x, dTw.y} {dTw.z, dcx, dcy, dopacity} {dr, dg, db, -}
__global__ void __launch_bounds__(SF_THREADS)
surfel_render_bwd_kernel(Cam cam, const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ plist,
                         const float4* __restrict__ rec2, const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                         const float* __restrict__ dL_dpix, float* __restrict__ sgrad2, const uint32_t* __restrict__ info) {
    __shared__ float4 s_rec[SF_THREADS * 4];
    __shared__ uint32_t s_id[SF_THREADS];
    if (info[2]) return;
    const int tile = blockIdx.x;
    const int tx = tile % cam.gx, ty = tile / cam.gx;
    const uint32_t r0 = tile_base[tile], n = tile_base[tile + 1] - r0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbit = 1u << warp;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7), py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t P = (size_t)cam.W * cam.H, pix = (size_t)py * cam.W + px;
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t last = inside ? n_contrib[pix] : 0u;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f, b0 = cam.bg[0], b1 = cam.bg[1], b2 = cam.bg[2];
    if (inside) {
        dp0 = __ldg(dL_dpix + pix); dp1 = __ldg(dL_dpix + P + pix); dp2 = __ldg(dL_dpix + 2 * P + pix);
        if (cam.bg_image) { b0 = __ldg(cam.bg_image + pix); b1 = __ldg(cam.bg_image + P + pix); b2 = __ldg(cam.bg_image + 2 * P + pix); }
    }
    const float bg_dot = b0 * dp0 + b1 * dp1 + b2 * dp2;
    float T = T_final, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    if (!__syncthreads_or(last != 0u)) return;  // no pixel of the tile blended anything
    const uint32_t wlast = __reduce_max_sync(0xffffffffu, last);  // deepest last contributor of the warp
    const int rounds = (int)((n + SF_THREADS - 1) / SF_THREADS);
    for (int rd = rounds - 1; rd >= 0; rd--) {
        const uint32_t base = (uint32_t)rd * SF_THREADS;
        __syncthreads();
        const uint32_t idx = base + threadIdx.x;
        if (idx < n) {
            const uint32_t e = __ldg(plist + r0 + idx);
            s_id[threadIdx.x] = e;  // id << 8 | sub-tile mask
            if (e & 0xffu) {
                const float4* r = rec2 + 4 * (size_t)(e >> 8);
                s_rec[4 * threadIdx.x] = __ldg(r); s_rec[4 * threadIdx.x + 1] = __ldg(r + 1);
                s_rec[4 * threadIdx.x + 2] = __ldg(r + 2); s_rec[4 * threadIdx.x + 3] = __ldg(r + 3);
            }
        }
        __syncthreads();
        const int cnt = (int)min((uint32_t)SF_THREADS, n - base);
        if (base >= wlast) continue;  // this warp's pixels all stopped earlier in the list
        // the warp's entries of this round as 32-bit masks (one ballot per 32 list entries: cull bit set, not past the warp's
        // deepest last contributor), walked from the top bit down
        for (int g0 = ((cnt - 1) >> 5) << 5; g0 >= 0; g0 -= 32) {
          const int jj = g0 + lane;
          uint32_t m = __ballot_sync(0xffffffffu, jj < cnt && (s_id[jj] & wbit) && base + (uint32_t)jj < wlast);
          while (m) {
            const int bit = 31 - __clz((int)m);
            m ^= 1u << bit;
            const int j = g0 + bit;
            const uint32_t contributor = base + (uint32_t)j;  // 0-based index of the entry in the tile's list
            float g[16];
#pragma unroll
            for (int k = 0; k < 16; k++) g[k] = 0.f;
            bool active = false;
            SurfelPair q;
            const float4 ra = s_rec[4 * j], rb = s_rec[4 * j + 1], rc = s_rec[4 * j + 2];
            if (contributor < last && surfel_pair(ra, rb, rc, pxf, pyf, q) && q.alpha >= 1.0f / 255.0f) {
                active = true;
                const float4 col = s_rec[4 * j + 3];
                const float inv_1ma = rcp_approx(1.0f - q.alpha);
                T = T * inv_1ma;
                acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0;
                acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1;
                acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2;
                lc0 = col.x; lc1 = col.y; lc2 = col.z;
                float dL_dalpha = (col.x - acc0) * dp0 + (col.y - acc1) * dp1 + (col.z - acc2) * dp2;
                const float wgt = q.alpha * T;
                g[12] = wgt * dp0; g[13] = wgt * dp1; g[14] = wgt * dp2;
                dL_dalpha *= T;
                last_alpha = q.alpha;
                dL_dalpha += (-T_final * inv_1ma) * bg_dot;
                const float dL_dG = rc.w * dL_dalpha;
                g[11] = q.G * dL_dalpha;
                if (q.use3d) {
                    const float dLu = dL_dG * -q.G * q.u, dLv = dL_dG * -q.G * q.v;
                    const float dsx = dLu * q.ipz, dsy = dLv * q.ipz;
                    const float d0 = dsx, d1 = dsy, d2 = -(dsx * q.u + dsy * q.v);
                    const float dk0 = q.l1 * d2 - q.l2 * d1, dk1 = q.l2 * d0 - q.l0 * d2, dk2 = q.l0 * d1 - q.l1 * d0;
                    const float dl0 = d1 * q.k2 - d2 * q.k1, dl1 = d2 * q.k0 - d0 * q.k2, dl2 = d0 * q.k1 - d1 * q.k0;
                    g[0] = -dk0; g[1] = -dk1; g[2] = -dk2;
                    g[3] = -dl0; g[4] = -dl1; g[5] = -dl2;
                    g[6] = pxf * dk0 + pyf * dl0; g[7] = pxf * dk1 + pyf * dl1; g[8] = pxf * dk2 + pyf * dl2;
                } else {
                    g[9] = dL_dG * -q.G * SF_FILTER_INV_SQ * q.dx;
                    g[10] = dL_dG * -q.G * SF_FILTER_INV_SQ * q.dy;
                }
            }
            if (!__any_sync(0xffffffffu, active)) continue;
            // warp sum of the 15 values by recursive halving: at each level a lane keeps one half of its values and trades the
            // other half with its partner, so 8 + 4 + 2 + 1 + 1 shuffles leave lane L with the warp total of value L >> 1
            // (instead of 15 x 5 shuffles for 15 full butterflies); four more shuffles gather them for the vector reductions
            {
                const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
                float a[8], b[4], c[2];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float send = h4 ? g[k] : g[k + 8], keep = h4 ? g[k + 8] : g[k];
                    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float send = h3 ? a[k] : a[k + 4], keep = h3 ? a[k + 4] : a[k];
                    b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const float send = h2 ? b[k] : b[k + 2], keep = h2 ? b[k + 2] : b[k];
                    c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                float t = (h1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c[0] : c[1], 2);
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                // lane L holds the total of value (L >> 1) & 15 with bit order: bit1 -> 1, bit2 -> 2, bit3 -> 4, bit4 -> 8
                const int src = 8 * (lane & 3);
                float4 v;
                v.x = __shfl_sync(0xffffffffu, t, src);
                v.y = __shfl_sync(0xffffffffu, t, src + 2);
                v.z = __shfl_sync(0xffffffffu, t, src + 4);
                v.w = __shfl_sync(0xffffffffu, t, src + 6);
                if (lane < 4) red_add_f4(sgrad2 + 16 * (size_t)(s_id[j] >> 8) + 4 * lane, v);
            }
          }
        }
    }
}

// SH basis and its gradient w.r.t. the (unit) direction, degree <= 3 (constants: gsplat_sh.hlsl:42-62)
__device__ __forceinline__ void sh_basis_grad(int deg, float X, float Y, float Z, float* bas, float (*gb)[3]) {
    const float C1 = 0.4886025119029199f;
    const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
    const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                         -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
    for (int k = 0; k < 16; k++) { bas[k] = 0.f; gb[k][0] = gb[k][1] = gb[k][2] = 0.f; }
    bas[0] = 0.28209479177387814f;
    if (deg < 1) return;
    bas[1] = -C1 * Y; bas[2] = C1 * Z; bas[3] = -C1 * X;
    gb[1][1] = -C1; gb[2][2] = C1; gb[3][0] = -C1;
    if (deg < 2) return;
    const float xx = X * X, yy = Y * Y, zz = Z * Z;
    bas[4] = C2[0] * X * Y; bas[5] = C2[1] * Y * Z; bas[6] = C2[2] * (2.f * zz - xx - yy); bas[7] = C2[3] * X * Z; bas[8] = C2[4] * (xx - yy);
    gb[4][0] = C2[0] * Y; gb[4][1] = C2[0] * X;
    gb[5][1] = C2[1] * Z; gb[5][2] = C2[1] * Y;
    gb[6][0] = C2[2] * -2.f * X; gb[6][1] = C2[2] * -2.f * Y; gb[6][2] = C2[2] * 4.f * Z;
    gb[7][0] = C2[3] * Z; gb[7][2] = C2[3] * X;
    gb[8][0] = C2[4] * 2.f * X; gb[8][1] = C2[4] * -2.f * Y;
    if (deg < 3) return;
    bas[9] = C3[0] * Y * (3.f * xx - yy); bas[10] = C3[1] * X * Y * Z; bas[11] = C3[2] * Y * (4.f * zz - xx - yy);
    bas[12] = C3[3] * Z * (2.f * zz - 3.f * xx - 3.f * yy); bas[13] = C3[4] * X * (4.f * zz - xx - yy);
    bas[14] = C3[5] * Z * (xx - yy); bas[15] = C3[6] * X * (xx - 3.f * yy);
    gb[9][0] = C3[0] * 6.f * X * Y; gb[9][1] = C3[0] * (3.f * xx - 3.f * yy);
    gb[10][0] = C3[1] * Y * Z; gb[10][1] = C3[1] * X * Z; gb[10][2] = C3[1] * X * Y;
    gb[11][0] = C3[2] * -2.f * X * Y; gb[11][1] = C3[2] * (4.f * zz - xx - 3.f * yy); gb[11][2] = C3[2] * 8.f * Y * Z;
    gb[12][0] = C3[3] * -6.f * X * Z; gb[12][1] = C3[3] * -6.f * Y * Z; gb[12][2] = C3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
    gb[13][0] = C3[4] * (4.f * zz - 3.f * xx - yy); gb[13][1] = C3[4] * -2.f * X * Y; gb[13][2] = C3[4] * 8.f * X * Z;
    gb[14][0] = C3[5] * 2.f * X * Z; gb[14][1] = C3[5] * -2.f * Y * Z; gb[14][2] = C3[5] * (xx - yy);
    gb[15][0] = C3[6] * (3.f * xx - 3.f * yy); gb[15][1] = C3[6] * -6.f * X * Y;
}

// S.4: per-surfel backward to the stored parameters; consumes and re-zeroes the sgrad2 record.
// STAGED: the only wide rows — shN in, dL/dshN out, 12 KR bytes per Gaussian each — are contiguous for the CTA's 128
// Gaussians, so they pass through shared memory: asynchronous 4-byte copies in (every sector fully used), each thread works on
// its own row there IN PLACE (reads coefficient k, then overwrites it with its gradient), and the span leaves with coalesced
// stores.  The direct form (thread i touching row i in global memory, 32 different 180-byte-strided sectors per warp
// instruction) took 0.51 ms at c3 against 0.12 ms for the 3DGS kernel, which stages the same way.
template <bool STAGED>
__global__ void __launch_bounds__(128)
surfel_preprocess_bwd_kernel(Cam cam, int N, Params prm, const uint4* __restrict__ aux, float4* __restrict__ sgrad2, Grads g,
                             uint32_t flags) {
    extern __shared__ float s_rows[];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool accumulate = flags & DVS_FLAG_ACCUMULATE;
    const bool acc_rows = accumulate && !STAGED;  // (staged rows are always written, the flush below adds or stores)
    const int deg = cam.deg, K = (deg + 1) * (deg + 1), KR = cam.KR;
    const int row_w = 3 * KR;
    const size_t blk_off = (size_t)blockIdx.x * blockDim.x * row_w;
    const int blk_words = min((int)blockDim.x, N - (int)(blockIdx.x * blockDim.x)) * row_w;
    if (STAGED && KR > 0) {
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(s_rows);
        for (int w = threadIdx.x; w < blk_words; w += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4u * (unsigned)w), "l"(prm.shN + blk_off + w) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const uint4 ax = i < N ? aux[i] : make_uint4(0u, 0u, 0u, 0u);
    if (STAGED && KR > 0) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    if (i < N) {
    const bool vis = ((ax.y & 0xffffu) > (ax.x & 0xffffu)) && (((ax.y >> 16) & 0x1fffu) > (ax.x >> 16));
    float dmean[3] = {0.f, 0.f, 0.f}, dsc[3] = {0.f, 0.f, 0.f}, dq4[4] = {0.f, 0.f, 0.f, 0.f}, dop = 0.f, dsh0[3] = {0.f, 0.f, 0.f};
    float gm2x = 0.f, gm2y = 0.f;
    float* shn_out = KR > 0 ? (STAGED ? s_rows + threadIdx.x * row_w : g.shN + (size_t)i * 3 * KR) : nullptr;
    if (vis) {
        float4* sg = sgrad2 + 4 * (size_t)i;
        const float4 g0 = sg[0], g1 = sg[1], g2 = sg[2], g3 = sg[3];
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        sg[0] = z4; sg[1] = z4; sg[2] = z4; sg[3] = z4;
        float dT[9] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x};
        const float dcx = g2.y, dcy = g2.z, dop_act = g2.w;
        const uint32_t clamped = ax.y >> 29;
        float dcol[3] = {(clamped & 1u) ? 0.f : g3.x, (clamped & 2u) ? 0.f : g3.y, (clamped & 4u) ? 0.f : g3.z};
        gm2x = dcx * 0.5f * (float)cam.W; gm2y = dcy * 0.5f * (float)cam.H;
        const float px = prm.means3D[3 * (size_t)i], py = prm.means3D[3 * (size_t)i + 1], pz = prm.means3D[3 * (size_t)i + 2];
        // activations
        float s[3], q[4], qlen = 1.f, o;
        const bool activated = cam.flags & DVS_FLAG_INPUT_ACTIVATED;
        {
            const float a0 = prm.scales[3 * (size_t)i], a1 = prm.scales[3 * (size_t)i + 1], a2 = prm.scales[3 * (size_t)i + 2];
            const float4 qq = reinterpret_cast<const float4*>(prm.quats)[i];
            const float oo = prm.opacities[i];
            if (activated) {
                s[0] = cam.scale_modifier * a0; s[1] = cam.scale_modifier * a1; s[2] = cam.scale_modifier * a2;
                q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w; o = oo;
            } else {
                s[0] = cam.scale_modifier * __expf(a0); s[1] = cam.scale_modifier * __expf(a1); s[2] = cam.scale_modifier * __expf(a2);
                qlen = sqrtf(qq.x * qq.x + qq.y * qq.y + qq.z * qq.z + qq.w * qq.w);
                const float inv = 1.0f / qlen;
                q[0] = qq.x * inv; q[1] = qq.y * inv; q[2] = qq.z * inv; q[3] = qq.w * inv;
                o = 1.0f / (1.0f + __expf(-oo));
            }
        }
        const float r = q[0], x = q[1], y = q[2], z = q[3];
        float R[3][3];
        R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
        R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
        R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
        const float* Pm = cam.proj;
        const float hw = 0.5f * (float)cam.W, hh = 0.5f * (float)cam.H, ow = 0.5f * (float)(cam.W - 1), oh = 0.5f * (float)(cam.H - 1);
        float T[9];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float v0, v1, v2, w1;
            if (j < 2) { v0 = R[0][j] * s[j]; v1 = R[1][j] * s[j]; v2 = R[2][j] * s[j]; w1 = 0.f; }
            else { v0 = px; v1 = py; v2 = pz; w1 = 1.f; }
            const float cx = Pm[0] * v0 + Pm[4] * v1 + Pm[8] * v2 + Pm[12] * w1;
            const float cy = Pm[1] * v0 + Pm[5] * v1 + Pm[9] * v2 + Pm[13] * w1;
            const float cw = Pm[3] * v0 + Pm[7] * v1 + Pm[11] * v2 + Pm[15] * w1;
            T[j] = hw * cx + ow * cw; T[3 + j] = hh * cy + oh * cw; T[6 + j] = cw;
        }
        {   // 1. projected centre -> T
            const float tp[3] = {9.f, 9.f, -1.f};
            const float dist = tp[0] * T[6] * T[6] + tp[1] * T[7] * T[7] + tp[2] * T[8] * T[8];
            float f[3], cx = 0.f, cy = 0.f;
#pragma unroll
            for (int j = 0; j < 3; j++) { f[j] = tp[j] / dist; cx += f[j] * T[j] * T[6 + j]; cy += f[j] * T[3 + j] * T[6 + j]; }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                dT[j] += dcx * f[j] * T[6 + j];
                dT[3 + j] += dcy * f[j] * T[6 + j];
                dT[6 + j] += dcx * (f[j] * T[j] - 2.f * cx * f[j] * T[6 + j]) + dcy * (f[j] * T[3 + j] - 2.f * cy * f[j] * T[6 + j]);
            }
        }
        // 2./3. T -> clip columns -> (L0, L1, p) through Proj^T
        float dvec[3][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float dcx_ = hw * dT[j], dcy_ = hh * dT[3 + j], dcw_ = ow * dT[j] + oh * dT[3 + j] + dT[6 + j];
#pragma unroll
            for (int rr = 0; rr < 3; rr++) dvec[j][rr] = Pm[4 * rr] * dcx_ + Pm[4 * rr + 1] * dcy_ + Pm[4 * rr + 3] * dcw_;
        }
        dmean[0] = dvec[2][0]; dmean[1] = dvec[2][1]; dmean[2] = dvec[2][2];
        {   // 4. SH
            float d0 = px - cam.campos[0], d1 = py - cam.campos[1], d2 = pz - cam.campos[2];
            const float len = sqrtf(d0 * d0 + d1 * d1 + d2 * d2), li = 1.0f / len;
            d0 *= li; d1 *= li; d2 *= li;
            float bas[16], gb[16][3];
            sh_basis_grad(deg, d0, d1, d2, bas, gb);
            dsh0[0] = bas[0] * dcol[0]; dsh0[1] = bas[0] * dcol[1]; dsh0[2] = bas[0] * dcol[2];
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
            const float* myrow = STAGED ? s_rows + threadIdx.x * row_w : prm.shN + (size_t)i * 3 * KR;
            for (int k = 1; k < K; k++) {
                const float sk = myrow[3 * (k - 1)] * dcol[0] + myrow[3 * (k - 1) + 1] * dcol[1] + myrow[3 * (k - 1) + 2] * dcol[2];
                ddx += gb[k][0] * sk; ddy += gb[k][1] * sk; ddz += gb[k][2] * sk;
                if (acc_rows) {
                    shn_out[3 * (k - 1)] += bas[k] * dcol[0]; shn_out[3 * (k - 1) + 1] += bas[k] * dcol[1]; shn_out[3 * (k - 1) + 2] += bas[k] * dcol[2];
                } else {
                    shn_out[3 * (k - 1)] = bas[k] * dcol[0]; shn_out[3 * (k - 1) + 1] = bas[k] * dcol[1]; shn_out[3 * (k - 1) + 2] = bas[k] * dcol[2];
                }
            }
            if (!acc_rows)
                for (int t = 3 * (K - 1); t < 3 * KR; t++) shn_out[t] = 0.f;
            const float dd = d0 * ddx + d1 * ddy + d2 * ddz;
            dmean[0] += (ddx - d0 * dd) * li; dmean[1] += (ddy - d1 * dd) * li; dmean[2] += (ddz - d2 * dd) * li;
        }
        {   // 5. (L0, L1) -> scales, rotation; 6. activations
            float ds[3] = {0.f, 0.f, 0.f}, dR[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int c = 0; c < 3; c++) dR[c][k] = 0.f;
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int c = 0; c < 3; c++) { ds[k] += R[c][k] * dvec[k][c]; dR[c][k] = s[k] * dvec[k][c]; }
            float dq[4];
            dq[0] = 2.f * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
            dq[1] = 2.f * (y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) - 4.f * x * (dR[1][1] + dR[2][2]);
            dq[2] = 2.f * (x * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + z * (dR[1][2] + dR[2][1])) - 4.f * y * (dR[0][0] + dR[2][2]);
            dq[3] = 2.f * (r * (dR[1][0] - dR[0][1]) + x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1])) - 4.f * z * (dR[0][0] + dR[1][1]);
            if (activated) {
                for (int k = 0; k < 3; k++) dsc[k] = ds[k] * cam.scale_modifier;
                for (int k = 0; k < 4; k++) dq4[k] = dq[k];
                dop = dop_act;
            } else {
                for (int k = 0; k < 3; k++) dsc[k] = ds[k] * s[k];
                const float qd = q[0] * dq[0] + q[1] * dq[1] + q[2] * dq[2] + q[3] * dq[3];
                for (int k = 0; k < 4; k++) dq4[k] = (dq[k] - q[k] * qd) / qlen;
                dop = dop_act * o * (1.0f - o);
            }
        }
    } else if (!acc_rows && KR > 0) {
        for (int t = 0; t < 3 * KR; t++) shn_out[t] = 0.f;
    }
    float* gm = g.means3D + 3 * (size_t)i;
    float* gs = g.scales + 3 * (size_t)i;
    float* gq = g.quats + 4 * (size_t)i;
    float* g0p = g.sh0 + 3 * (size_t)i;
    if (accumulate) {
        if (vis) {
            for (int k = 0; k < 3; k++) { gm[k] += dmean[k]; gs[k] += dsc[k]; g0p[k] += dsh0[k]; }
            for (int k = 0; k < 4; k++) gq[k] += dq4[k];
            g.opacities[i] += dop;
            if (g.mean2D) { g.mean2D[2 * (size_t)i] += gm2x; g.mean2D[2 * (size_t)i + 1] += gm2y; }
        }
    } else {
        for (int k = 0; k < 3; k++) { gm[k] = dmean[k]; gs[k] = dsc[k]; g0p[k] = dsh0[k]; }
        for (int k = 0; k < 4; k++) gq[k] = dq4[k];
        g.opacities[i] = dop;
        if (g.mean2D) { g.mean2D[2 * (size_t)i] = gm2x; g.mean2D[2 * (size_t)i + 1] = gm2y; }
        if (g.mean2D_abs) { g.mean2D_abs[2 * (size_t)i] = fabsf(gm2x); g.mean2D_abs[2 * (size_t)i + 1] = fabsf(gm2y); }
    }
    }  // i < N
    if (STAGED && KR > 0) {
        __syncthreads();
        float* dst = g.shN + blk_off;
        if (accumulate) {
            for (int w = threadIdx.x; w < blk_words; w += blockDim.x) dst[w] += s_rows[w];
        } else {
            for (int w = threadIdx.x; w < blk_words; w += blockDim.x) dst[w] = s_rows[w];
        }
    }
}
}  // namespace

cudaError_t launch_surfel_render_fwd(const Cam& cam, const uint32_t* tile_base, const uint32_t* plist, const float4* rec2,
                                     float* out_color, float* final_T, uint32_t* n_contrib, const uint32_t* info, cudaStream_t st) {
    const int T = cam.gx * cam.gy;
    if (T <= 0) return cudaSuccess;
    surfel_render_fwd_kernel<<<T, SF_THREADS, 0, st>>>(cam, tile_base, plist, rec2, out_color, final_T, n_contrib, info);
    return cudaGetLastError();
}
cudaError_t launch_surfel_render_bwd(const Cam& cam, const uint32_t* tile_base, const uint32_t* plist, const float4* rec2,
                                     const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float* sgrad2,
                                     const uint32_t* info, cudaStream_t st) {
    const int T = cam.gx * cam.gy;
    if (T <= 0) return cudaSuccess;
    surfel_render_bwd_kernel<<<T, SF_THREADS, 0, st>>>(cam, tile_base, plist, rec2, final_T, n_contrib, dL_dpix, sgrad2, info);
    return cudaGetLastError();
}
cudaError_t launch_surfel_preprocess_bwd(const Cam& cam, int N, const Params& prm, const uint4* aux, float4* sgrad2, const Grads& g,
                                         uint32_t flags, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    static const bool direct = [] { const char* e = getenv("DVS_SURFEL_PB_DIRECT"); return e && atoi(e) != 0; }();  // A/B switch
    const size_t smem = (size_t)128 * 3 * cam.KR * sizeof(float);
    if (direct || smem > 48 * 1024)
        surfel_preprocess_bwd_kernel<false><<<(N + 127) / 128, 128, 0, st>>>(cam, N, prm, aux, sgrad2, g, flags);
    else
        surfel_preprocess_bwd_kernel<true><<<(N + 127) / 128, 128, smem, st>>>(cam, N, prm, aux, sgrad2, g, flags);
    return cudaGetLastError();
}

}  // namespace dvs
