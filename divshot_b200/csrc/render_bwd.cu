// render_bwd.cu — A7: per-pixel reverse-walk gradient pass of the tile compositor.
//
// Replaces `renderCUDA` (backward) of the absent gsplatrast operator (no
// in-tree corroboration exists — the viewer has no backward pass).
//
// B200 design (not the upstream kernel, which issues 9 global atomics per (pixel, splat) pair):
//   * same tiling / sub-tile masks as the forward: warp w owns an 8x4-pixel sub-rectangle and only touches
//     splats whose footprint reaches it and that lie in front of the warp's deepest last-contributor;
//   * PHASE 1 (lane = pixel): the reverse walk proper.  Per (pixel, splat) pair only two numbers are
//     produced: s = dL/dpower and w = alpha*T (colour weight).  They go to two per-warp shared-memory
//     buffers S[slot][pixel], Wt[slot][pixel] (16 splats deep) — no cross-lane reduction here;
//   * PHASE 2 (lane = splat x pixel-half): every 16 buffered splats the lanes switch roles.  Lane (k, h)
//     walks 16 of the 32 pixels for splat k as 8 horizontally adjacent PAIRS held in packed fp32x2
//     registers (FFMA2 / FMUL2 / FADD2, new on sm_100) and accumulates the nine sums
//         S0 = sum s, Sx = sum s dx, Sy = sum s dy, Sxx, Sxy, Syy, and sum w*dL/dpix[0..2]
//     (10 packed operations per pixel pair, no shuffles); the two halves meet with one shuffle per sum and
//     each leaves with ONE 128-bit vector reduction (REDG.E.ADD.F32x4) into the 48-byte screen-gradient record;
//   * all per-splat constant factors (conic, W/2, -1/2, 1/opacity) are applied once per splat in the
//     preprocess backward, not per pair.
// Versus a shuffle-tree reduction per (warp, splat) this is ~2.5x fewer issued instructions and ~20x fewer
// SHFL.  Bound: instruction issue and shared-memory wavefronts (profiles/r1_v4_step_ncu_full.md), not HBM.
#include "common.cuh"
#include "kernels.h"

namespace dvs {

constexpr int RB_THREADS = 256;
#ifndef DVS_RB_ROUND
#define DVS_RB_ROUND 256
#endif
#ifndef DVS_RB_MINCTA
#define DVS_RB_MINCTA 4
#endif
constexpr int RB_ROUND = DVS_RB_ROUND;  // entries staged per round (128: +1%; 64 with 5 CTAs/SM: no faster)
constexpr int RB_NB = 16;      // splats buffered per warp between phase 1 and phase 2
constexpr int RB_SROW = 34;    // floats per buffer row (32 pixels + 2: the 64-bit pair loads of phase 2 are bank-conflict-free)

struct RbSmem {
    // byte offsets inside dynamic shared memory
    static constexpr int stage = 0;                              // RB_ROUND * 48
    static constexpr int ent = stage + RB_ROUND * 48;            // RB_ROUND * 4
    static constexpr int wmax = ent + RB_ROUND * 4;              // 8 * 4
    static constexpr int warp0 = wmax + 64;                      // per-warp region start
    static constexpr int S = 0;                                  // RB_NB rows of RB_SROW floats: s = dL/dpower per (slot, pixel)
    static constexpr int Wt = S + RB_NB * RB_SROW * 4;           // same shape: w = alpha*T (colour weight)
    static constexpr int meta = Wt + RB_NB * RB_SROW * 4;        // RB_NB * 8 words {entry word (id << 8 | mask), -, mx, my, A, B, C, -}
    static constexpr int dpA = meta + RB_NB * 8 * 4;             // 16 pixel pairs x {r0, r1, g0, g1} of dL/dpix
    static constexpr int dpB = dpA + 16 * 16;                    // 16 pixel pairs x {b0, b1}
    static constexpr int per_warp = dpB + 16 * 8;
    static constexpr int total = warp0 + (RB_THREADS / 32) * per_warp;
};

template <bool ABSGRAD>
__device__ __forceinline__ void rb_phase2(uint32_t wbase, int nbuf, int lane, float px0f, float py0f,
                                          float* __restrict__ sgrad) {
    __syncwarp();
    const int k = lane & 15, h = lane >> 4;
    const uint32_t mrow = wbase + RbSmem::meta + k * 32;
    const float2 mxy = lds_f2(mrow + 8);
    const float X = mxy.x - px0f;                     // dx of pixel column 0 of the sub-rectangle
    const float Y = mxy.y - (py0f + (float)(2 * h));  // dy of the first of my two pixel rows
    const uint32_t srow = wbase + RbSmem::S + (k * RB_SROW + 16 * h) * 4;
    const f32x2 Xp = pk2(X, X - 1.0f), m2 = pk2(-2.0f, -2.0f), m1 = pk2(-1.0f, -1.0f);
    // My 16 pixels are 8 horizontally adjacent PAIRS (2 rows x 4 pairs); every quantity is carried as a packed
    // fp32x2 {even pixel, odd pixel}: 10 FFMA2/FMUL2/FADD2 per pair instead of 24 scalar operations.
    // Pass A: the six geometric moments of s.
    f32x2 aS0 = 0ull, aSx = 0ull, aSy = 0ull, aSxx = 0ull, aSxy = 0ull, aSyy = 0ull;
    {
        f32x2 dyp = pk2(Y, Y);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            f32x2 dxp = Xp;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const f32x2 s = lds_p2(srow + (8 * r + 2 * q) * 4);
                const f32x2 sdx = mul2(s, dxp), sdy = mul2(s, dyp);
                aS0 = add2(aS0, s);
                aSx = add2(aSx, sdx);
                aSy = add2(aSy, sdy);
                aSxx = fma2(sdx, dxp, aSxx);
                aSxy = fma2(sdx, dyp, aSxy);
                aSyy = fma2(sdy, dyp, aSyy);
                if (q < 3) dxp = add2(dxp, m2);
            }
            if (r == 0) dyp = add2(dyp, m1);
        }
    }
    float S0 = sum2(aS0), Sx = sum2(aSx), Sy = sum2(aSy), Sxx = sum2(aSxx), Sxy = sum2(aSxy), Syy = sum2(aSyy);
    S0 += __shfl_xor_sync(0xffffffffu, S0, 16); Sx += __shfl_xor_sync(0xffffffffu, Sx, 16);
    Sy += __shfl_xor_sync(0xffffffffu, Sy, 16); Sxx += __shfl_xor_sync(0xffffffffu, Sxx, 16);
    Sxy += __shfl_xor_sync(0xffffffffu, Sxy, 16); Syy += __shfl_xor_sync(0xffffffffu, Syy, 16);
    // Pass B: colour sums  sum_pixels w * dL/dpix[ch].
    float c0, c1, c2;
    {
        f32x2 a0 = 0ull, a1 = 0ull, a2 = 0ull;
        const uint32_t wrow = srow + (RbSmem::Wt - RbSmem::S);
        const uint32_t pa = wbase + RbSmem::dpA + 8 * h * 16, pb = wbase + RbSmem::dpB + 8 * h * 8;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const f32x2 w = lds_p2(wrow + 8 * q);
            f32x2 rr, gg;
            lds_p4(pa + 16 * q, rr, gg);
            const f32x2 bb = lds_p2(pb + 8 * q);
            a0 = fma2(w, rr, a0);
            a1 = fma2(w, gg, a1);
            a2 = fma2(w, bb, a2);
        }
        c0 = sum2(a0); c1 = sum2(a1); c2 = sum2(a2);
    }
    c0 += __shfl_xor_sync(0xffffffffu, c0, 16); c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
    c2 += __shfl_xor_sync(0xffffffffu, c2, 16);
    float ax = 0.f, ay = 0.f;
    if (ABSGRAD) {  // Pass C: sum_pixels |dL/dmean2D contribution| (densification statistic), natural-units conic
        const f32x2 cA = pk2(lds_f1(mrow + 16), lds_f1(mrow + 16)), cB = pk2(lds_f1(mrow + 20), lds_f1(mrow + 20)),
                    cC = pk2(lds_f1(mrow + 24), lds_f1(mrow + 24));
        f32x2 bx = 0ull, by = 0ull;
        f32x2 dyp = pk2(Y, Y);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            f32x2 dxp = Xp;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const f32x2 s = lds_p2(srow + (8 * r + 2 * q) * 4);
                const f32x2 sdx = mul2(s, dxp), sdy = mul2(s, dyp);
                bx = add2(bx, abs2(fma2(cA, sdx, mul2(cB, sdy))));
                by = add2(by, abs2(fma2(cB, sdx, mul2(cC, sdy))));
                if (q < 3) dxp = add2(dxp, m2);
            }
            if (r == 0) dyp = add2(dyp, m1);
        }
        ax = sum2(bx); ay = sum2(by);
        ax += __shfl_xor_sync(0xffffffffu, ax, 16); ay += __shfl_xor_sync(0xffffffffu, ay, 16);
    }
    // both halves hold the totals: half 0 sends floats 0-3 of the 48-byte record, half 1 floats 4-7, as ONE
    // 128-bit vector reduction (REDG.E.ADD.F32x4) — 2 reduction instructions per 16 splats instead of 9, and
    // ~4x fewer L1TEX tag wavefronts (every lane's record is a different cache line)
    if (k < nbuf) {
        float* g = sgrad + 12 * (size_t)(lds_u1(mrow) >> 8);
        const float4 v = h ? make_float4(Syy, S0, c0, c1) : make_float4(Sx, Sy, Sxx, Sxy);
        red_add_f4(g + 4 * h, v);
        if (h == 0) {
            if (ABSGRAD) red_add_f4(g + 8, make_float4(c2, ax, ay, 0.f));
            else atomicAdd(g + 8, c2);
        }
    }
    __syncwarp();
}

template <bool ABSGRAD>
__global__ void __launch_bounds__(RB_THREADS, DVS_RB_MINCTA)
render_bwd_kernel(Cam cam, const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ plist,
                  const float4* __restrict__ rec, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  float* __restrict__ sgrad, const uint32_t* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char rb_smem[];
    if (info[2]) return;
    uint32_t sb0 = smem_u32(rb_smem);
    asm volatile("" : "+r"(sb0));  // keep the shared base address in a register (no re-derivation per pair)
    const uint32_t sb = sb0 + RbSmem::stage, se = sb0 + RbSmem::ent, swm = sb0 + RbSmem::wmax;
    const int tile = blockIdx.x;
    const int tx = tile % cam.gx, ty = tile / cam.gx;
    const uint32_t r0 = tile_base[tile], n = tile_base[tile + 1] - r0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = sb0 + RbSmem::warp0 + warp * RbSmem::per_warp;
    const int px0 = tx * TILE + (warp & 1) * 8, py0 = ty * TILE + (warp >> 1) * 4;
    const int px = px0 + (lane & 7), py = py0 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const float px0f = (float)px0, py0f = (float)py0;
    const f32x2 npxy = pk2(-(float)px, -(float)py);
    const size_t P = (size_t)cam.W * cam.H;
    const size_t pix = (size_t)py * cam.W + px;

    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t last = inside ? n_contrib[pix] : 0u;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) {
        dp0 = __ldg(dL_dpix + pix);
        dp1 = __ldg(dL_dpix + P + pix);
        dp2 = __ldg(dL_dpix + 2 * P + pix);
    }
    {   // dL/dpix of the warp's 32 pixels, pair-interleaved for the packed loads of phase 2
        const uint32_t pr = (uint32_t)(lane >> 1), odd = (uint32_t)(lane & 1);
        sts_f1(wbase + RbSmem::dpA + pr * 16 + odd * 4, dp0);
        sts_f1(wbase + RbSmem::dpA + pr * 16 + 8 + odd * 4, dp1);
        sts_f1(wbase + RbSmem::dpB + pr * 8 + odd * 4, dp2);
    }
    const float nTf_bg = -T_final * (cam.bg[0] * dp0 + cam.bg[1] * dp1 + cam.bg[2] * dp2);
    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;

    uint32_t wmax = last;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
    if (lane == 0) sts_u1(swm + warp * 4, wmax);
    __syncthreads();
    uint32_t cmax = 0;
#pragma unroll
    for (int w = 0; w < RB_THREADS / 32; w++) cmax = max(cmax, lds_u1(swm + w * 4));
    cmax = min(cmax, n);
    if (cmax == 0) return;
    const uint32_t wbit = 1u << warp;
    int nbuf = 0;
    // per-lane row addresses of the phase-1 buffer, kept in registers (no re-derivation from tid per pair)
    uint32_t mySW = wbase + RbSmem::S + lane * 4;
    uint32_t mmeta = wbase + RbSmem::meta;
    asm volatile("" : "+r"(mySW), "+r"(mmeta));

    // staging is software-pipelined: while a round is being walked, the next round's entry words and records
    // are already in flight into registers of the first RB_ROUND threads
    uint32_t e_n = 0;
    float4 q0_n = make_float4(0.f, 0.f, 0.f, 0.f), q1_n = q0_n;
    float b_n = 0.f;
    auto fetch = [&](int round) {
        const uint32_t idx = (uint32_t)round * RB_ROUND + threadIdx.x;
        e_n = 0;
        if (idx < cmax) e_n = __ldg(plist + r0 + idx);
        if (e_n & 0xffu) {
            const float4* r = rec + 3 * (size_t)(e_n >> 8);
            q0_n = __ldg(r);
            q1_n = __ldg(r + 1);
            b_n = __ldg(reinterpret_cast<const float*>(r + 2));
        }
    };
    const int rd0 = (int)((cmax - 1) / RB_ROUND);
    if (threadIdx.x < RB_ROUND) fetch(rd0);
    for (int rd = rd0; rd >= 0; rd--) {
        const uint32_t base_idx = (uint32_t)rd * RB_ROUND;
        __syncthreads();  // previous round fully consumed
        if (threadIdx.x < RB_ROUND) {
            sts_u1(se + threadIdx.x * 4, e_n);
            if (e_n & 0xffu) {
                // staged as {mx, my, A2, C2} {B2, lo, r, g} (register pairs for FADD2 / FMUL2, as in the forward)
                sts_f4(sb + threadIdx.x * 48, make_float4(q0_n.x, q0_n.y, q0_n.z, q1_n.x));
                sts_f4(sb + threadIdx.x * 48 + 16, make_float4(q0_n.w, q1_n.y, q1_n.z, q1_n.w));
                sts_f1(sb + threadIdx.x * 48 + 32, b_n);
            }
        }
        __syncthreads();
        if (rd > 0 && threadIdx.x < RB_ROUND) fetch(rd - 1);
        if (base_idx >= wmax) continue;  // this warp's pixels all stopped earlier in the list
        const int cnt = (int)min((uint32_t)RB_ROUND, cmax - base_idx);
        const int lastr = (int)last - (int)base_idx;  // entry k of this round is in front of my last contributor iff k < lastr
        for (int c = ((cnt - 1) >> 5) << 5; c >= 0; c -= 32) {
            if (base_idx + (uint32_t)c >= wmax) continue;
            const uint32_t myw = lds_u1(se + (c + lane) * 4);
            uint32_t bits = __ballot_sync(0xffffffffu, (myw & wbit) != 0u);
            while (bits) {
                const int j = 31 - __clz(bits);
                bits &= ~(1u << j);
                const int k = c + j;
                const uint32_t ea = sb + k * 48;
                f32x2 mxy, AC;
                lds_p4(ea, mxy, AC);
                const float4 q1 = lds_f4(ea + 16);  // {B2, lo, r, g}
                float dx, dy, adx, cdy;
                const f32x2 d = add2(mxy, npxy);    // same arithmetic, bit for bit, as the forward kernel
                upk2(d, dx, dy);
                upk2(mul2(AC, d), adx, cdy);
                const float t = fmaf(q1.x, dy, adx);
                const float pw = fmaf(cdy, dy, t * dx);
                const float ee = pw + q1.y;
                const bool act = k < lastr && pw <= 0.0f && ee >= ALPHA_MIN_LOG2;
                if (!__any_sync(0xffffffffu, act)) continue;
                float s = 0.f, wgt = 0.f;
                if (act) {
                    const float a_raw = ex2_approx(ee);
                    const float alpha = fminf(0.99f, a_raw);
                    const float rinv = rcp_approx(1.0f - alpha);
                    T = T * rinv;
                    wgt = alpha * T;
                    const float cb = lds_f1(ea + 32);
                    // R = colour accumulated behind this splat (what upstream calls accum_rec at the time of use);
                    // dL/dalpha = sum_ch (c - R) dL/dpix, then R <- alpha c + (1 - alpha) R = R + alpha (c - R)
                    const float d0 = q1.z - acc0, d1 = q1.w - acc1, d2 = cb - acc2;
                    float dL_dalpha = d0 * dp0;
                    dL_dalpha = fmaf(d1, dp1, dL_dalpha);
                    dL_dalpha = fmaf(d2, dp2, dL_dalpha);
                    acc0 = fmaf(alpha, d0, acc0);
                    acc1 = fmaf(alpha, d1, acc1);
                    acc2 = fmaf(alpha, d2, acc2);
                    dL_dalpha = fmaf(dL_dalpha, T, nTf_bg * rinv);
                    s = a_raw * dL_dalpha;  // dL/dpower (the 0.99 clamp is straight-through)
                }
                sts_f1(mySW + nbuf * (RB_SROW * 4), s);
                sts_f1(mySW + nbuf * (RB_SROW * 4) + (RbSmem::Wt - RbSmem::S), wgt);
                {   // slot metadata, stored by all lanes to one address with one value (a single wavefront each):
                    // the entry word comes by shuffle from the lane that owns the ballot bit
                    const uint32_t mrow = mmeta + nbuf * 32;
                    sts_u1(mrow, __shfl_sync(0xffffffffu, myw, j));
                    sts_p2(mrow + 8, mxy);
                    if (ABSGRAD) {  // natural-units conic for the |dL/dmean2D| statistic
                        float A2, C2;
                        upk2(AC, A2, C2);
                        sts_f2(mrow + 16, A2 * (-2.0f * LN2), q1.x * (-LN2));
                        sts_f1(mrow + 24, C2 * (-2.0f * LN2));
                    }
                }
                if (++nbuf == RB_NB) {
                    rb_phase2<ABSGRAD>(wbase, RB_NB, lane, px0f, py0f, sgrad);
                    nbuf = 0;
                }
            }
        }
    }
    if (nbuf) rb_phase2<ABSGRAD>(wbase, nbuf, lane, px0f, py0f, sgrad);  // rows k >= nbuf are masked at the RED
}

cudaError_t launch_render_bwd(const Cam& cam, const uint32_t* tile_base, const uint32_t* plist, const float4* rec,
                              const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float* sgrad,
                              bool absgrad, const uint32_t* info, cudaStream_t st) {
    const int T = cam.gx * cam.gy;
    if (T <= 0) return cudaSuccess;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(render_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RbSmem::total);
        cudaFuncSetAttribute(render_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RbSmem::total);
        attr_done = true;
    }
    if (absgrad)
        render_bwd_kernel<true><<<T, RB_THREADS, RbSmem::total, st>>>(cam, tile_base, plist, rec, final_T, n_contrib,
                                                                      dL_dpix, sgrad, info);
    else
        render_bwd_kernel<false><<<T, RB_THREADS, RbSmem::total, st>>>(cam, tile_base, plist, rec, final_T, n_contrib,
                                                                       dL_dpix, sgrad, info);
    return cudaGetLastError();
}

}  // namespace dvs
