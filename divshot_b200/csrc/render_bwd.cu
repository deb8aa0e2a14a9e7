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
//     buffers S[slot][pixel], Wt[slot][pixel] (16 splats deep) — no cross-lane reduction here.  The walk is one
//     straight-line predicated sequence (no divergent branch); the colour recursion is carried as ONE scalar per
//     pixel, (colour accumulated behind) . dL/dpix, instead of three channels; the slot's identity (its position in
//     the sorted list) and mean are captured in the registers of lane `slot` (no shuffle, no shared-memory
//     metadata).  ~45 issue slots and 11 shared-memory wavefronts per (warp, splat) visit (round 1: 56 and 14);
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
#define DVS_RB_ROUND 128
#endif
#ifndef DVS_RB_MINCTA
#define DVS_RB_MINCTA 4
#endif
constexpr int RB_ROUND = DVS_RB_ROUND;  // entries staged per round and buffer (two buffers)
constexpr int RB_NB = 16;      // splats buffered per warp between phase 1 and phase 2
constexpr int RB_SROW = 34;    // floats per buffer row (32 pixels + 2: the 64-bit pair loads of phase 2 are bank-conflict-free)

struct RbSmem {
    // byte offsets inside dynamic shared memory
    static constexpr int stage = 0;                              // 2 buffers of RB_ROUND * 48
    static constexpr int ent = stage + 2 * RB_ROUND * 48;        // 2 buffers of RB_ROUND * 4
    static constexpr int wmax = ent + 2 * RB_ROUND * 4;          // 8 * 4
    static constexpr int warp0 = wmax + 64;                      // per-warp region start
    static constexpr int S = 0;                                  // RB_NB rows of RB_SROW floats: s = dL/dpower per (slot, pixel)
    static constexpr int Wt = S + RB_NB * RB_SROW * 4;           // same shape: w = alpha*T (colour weight)
    static constexpr int dpA = Wt + RB_NB * RB_SROW * 4;         // 16 pixel pairs x {r0, r1, g0, g1} of dL/dpix
    static constexpr int dpB = dpA + 16 * 16;                    // 16 pixel pairs x {b0, b1}
    static constexpr int per_warp = dpB + 16 * 8;
    static constexpr int total = warp0 + (RB_THREADS / 32) * per_warp;
};

// Slot metadata captured by lane `slot` during phase 1 (lanes 16..31 hold nothing).
struct RbSlot {
    uint32_t gk;     // position of the entry in the sorted list (plist index); the mean is re-read from the record in phase 2
};

template <bool ABSGRAD>
__device__ __forceinline__ void rb_phase2(uint32_t wbase, int nbuf, int lane, float px0f, float py0f,
                                          const RbSlot& mine, const uint32_t* __restrict__ plist,
                                          const float4* __restrict__ rec, float* __restrict__ sgrad) {
    __syncwarp();
    const int k = lane & 15, h = lane >> 4;
    // both pixel-halves of slot k read the slot's identity from lane k; the entry word (Gaussian id) comes from the
    // sorted list itself (one L1/L2 hit per slot and flush, issued first so that the moment passes hide it)
    const uint32_t gk = __shfl_sync(0xffffffffu, mine.gk, k);
    uint32_t eword = 0;
    if (k < nbuf) eword = __ldg(plist + gk);
    // the slot's mean: the first 8 bytes of its record (an L2 hit: the record was staged for this tile moments ago); two
    // dependent loads per slot and FLUSH instead of two register moves per VISIT, issued before the colour pass hides them
    const float2 mxy = __ldg(reinterpret_cast<const float2*>(rec + 3 * (size_t)(eword >> 8)));
    const uint32_t srow = wbase + RbSmem::S + (k * RB_SROW + 16 * h) * 4;
    // My 16 pixels are 8 horizontally adjacent PAIRS (2 rows x 4 pairs); every quantity is carried as a packed
    // fp32x2 {even pixel, odd pixel}: 10 FFMA2/FMUL2/FADD2 per pair instead of 24 scalar operations.
    // Pass B first (it needs nothing of the slot's record): colour sums  sum_pixels w * dL/dpix[ch].
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
    // Pass A: the six geometric moments of s about the slot's mean.
    const float X = mxy.x - px0f;                     // dx of pixel column 0 of the sub-rectangle
    const float Y = mxy.y - (py0f + (float)(2 * h));  // dy of the first of my two pixel rows
    const f32x2 Xp = pk2(X, X - 1.0f), m2 = pk2(-2.0f, -2.0f), m1 = pk2(-1.0f, -1.0f);
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
    float ax = 0.f, ay = 0.f;
    if (ABSGRAD) {  // Pass C: sum_pixels |dL/dmean2D contribution| (densification statistic), natural-units conic
        // natural-units conic from the folded one in the record: A = -2 ln2 A2, B = -ln2 B2, C = -2 ln2 C2
        const float4* r = rec + 3 * (size_t)(eword >> 8);  // (eword = 0 for unused slots: record 0, result discarded)
        const float nA = __ldg(&r[0].z) * (-2.0f * LN2), nC = __ldg(&r[0].w) * (-2.0f * LN2), nB = __ldg(&r[1].x) * (-LN2);
        const f32x2 cA = pk2(nA, nA), cB = pk2(nB, nB), cC = pk2(nC, nC);
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
        float* g = sgrad + 12 * (size_t)(eword >> 8);
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
render_bwd_kernel(Cam cam, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ tile_base,
                  const uint32_t* __restrict__ plist, const float4* __restrict__ rec, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  float* __restrict__ sgrad, const uint32_t* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char rb_smem[];
    if (info[2]) return;
    uint32_t sb0 = smem_u32(rb_smem);
    asm volatile("" : "+r"(sb0));  // keep the shared base address in a register (no re-derivation per pair)
    const uint32_t sb = sb0 + RbSmem::stage, se = sb0 + RbSmem::ent, swm = sb0 + RbSmem::wmax;
    const int tile = tile_order ? (int)tile_order[blockIdx.x] : (int)blockIdx.x;
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
    float b0 = cam.bg[0], b1 = cam.bg[1], b2 = cam.bg[2];
    if (cam.bg_image && inside) {  // per-pixel background (sky model)
        b0 = __ldg(cam.bg_image + pix); b1 = __ldg(cam.bg_image + P + pix); b2 = __ldg(cam.bg_image + 2 * P + pix);
    }
    const float nTf_bg = -T_final * (b0 * dp0 + b1 * dp1 + b2 * dp2);
    float T = T_final;
    // accdp = (colour accumulated behind the current splat) . dL/dpix — the three-channel recursion
    // R <- R + alpha (c - R) collapses to one scalar because only R . dL/dpix is ever used
    float accdp = 0.f;

    // deepest last contributor of the warp / of the tile.  REDUX results are warp-uniform by construction, which lets
    // ptxas keep the whole loop nest below (bounds, buffer addresses, list positions) in uniform registers
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) sts_u1(swm + warp * 4, wmax);
    __syncthreads();
    uint32_t cmax = __reduce_max_sync(0xffffffffu, lds_u1(swm + (lane & (RB_THREADS / 32 - 1)) * 4));
    cmax = min(cmax, n);
    if (cmax == 0) return;
    const uint32_t wbit = 1u << warp;
    const uint32_t row0 = wbase + RbSmem::S + lane * 4;
    // byte offset of the buffer row being filled (slot * row pitch): a pure counter with a select-wrap, so that ptxas can
    // prove it warp-uniform and keep it (and the whole loop nest) on the uniform datapath
    constexpr uint32_t ROWB = RB_SROW * 4;
    uint32_t rowoff = 0;
    const uint32_t lane_row = (uint32_t)lane * ROWB;  // rowoff == lane_row <=> the slot being filled is `lane`
    RbSlot mine = {0u};
    const int glast = (int)(r0 + last);  // entry at list position g is in front of my last contributor iff g < glast

    // staging: two buffers of RB_ROUND records + entry words; the NEXT round (the walk goes back to front) is copied in
    // by 16-byte asynchronous copies (cp.async / LDGSTS) while the current one is walked — one CTA barrier per round and no
    // staging registers (round 1 held the prefetched record in 10 registers across the walk)
    auto stage = [&](uint32_t e, uint32_t buf) {
        sts_u1(se + (buf * RB_ROUND + threadIdx.x) * 4, e);
        if (e & 0xffu) {
            const float4* r = rec + 3 * (size_t)(e >> 8);
            const uint32_t dst = sb + buf * (RB_ROUND * 48) + threadIdx.x * 48;
            cp_async16(dst, r);
            cp_async16(dst + 16, r + 1);
            cp_async16(dst + 32, r + 2);
        }
        cp_async_commit();
    };
    auto entry_of = [&](int round) -> uint32_t {
        const uint32_t idx = (uint32_t)round * RB_ROUND + threadIdx.x;
        return (round >= 0 && idx < cmax) ? __ldg(plist + r0 + idx) : 0u;
    };
    const int rd0 = (int)((cmax - 1) / RB_ROUND);
    uint32_t e_n = 0;
    if (threadIdx.x < RB_ROUND) {
        stage(entry_of(rd0), (uint32_t)rd0 & 1u);
        e_n = entry_of(rd0 - 1);
    }
    for (int rd = rd0; rd >= 0; rd--) {
        const uint32_t base_idx = (uint32_t)rd * RB_ROUND;
        cp_async_wait0();  // my copies of round rd have landed
        __syncthreads();   // ... and everyone's; everyone has finished round rd+1 (whose buffer is refilled below)
        if (rd > 0 && threadIdx.x < RB_ROUND) {
            stage(e_n, (uint32_t)(rd - 1) & 1u);
            e_n = entry_of(rd - 2);
        }
        const uint32_t sbr = sb + ((uint32_t)rd & 1u) * (RB_ROUND * 48), ser = se + ((uint32_t)rd & 1u) * (RB_ROUND * 4);
        if (base_idx >= wmax) continue;  // this warp's pixels all stopped earlier in the list
        const int cnt = (int)min((uint32_t)RB_ROUND, cmax - base_idx);
        for (int c = ((cnt - 1) >> 5) << 5; c >= 0; c -= 32) {
            if (base_idx + (uint32_t)c >= wmax) continue;
            uint32_t bits = __ballot_sync(0xffffffffu, (lds_u1(ser + (c + lane) * 4) & wbit) != 0u);
            uint32_t ea0 = sbr + (uint32_t)c * 48u;
            int g0 = (int)(r0 + base_idx) + c;  // list position of the group's first entry
            asm volatile("" : "+r"(ea0), "+r"(g0));  // per-group values: keep them in registers, do not re-derive per visit
            while (bits) {
                uint32_t p, below;
                asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(bits));                     // highest set bit: FLO, no clz round trip
                asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(below) : "r"(0), "r"(p));  // bits [0, p)
                bits &= below;
                const uint32_t ea = ea0 + p * 48u;
                const int gk = g0 + (int)p;
                f32x2 mxy, AC;
                lds_p4(ea, mxy, AC);
                const float4 q1 = lds_f4(ea + 16);  // {B2, lo, r, g}
                const float cb = lds_f1(ea + 32);
                float dx, dy, adx, cdy;
                const f32x2 d = add2(mxy, npxy);    // same arithmetic, bit for bit, as the forward kernel
                upk2(d, dx, dy);
                upk2(mul2(AC, d), adx, cdy);
                const float t = fmaf(q1.x, dy, adx);
                const float pw = fmaf(cdy, dy, t * dx);
                const float ee = pw + q1.y;
                // pa = the pair was blended by the forward: in front of my last contributor, power <= 0, alpha >= 1/255.
                // The pair's raw alpha is SELECTED to 0 otherwise, which turns every update below into an exact no-op for it
                // (accdp + 0 * dd = accdp, s = w = 0) and T is stepped by a predicated multiply — one select and one predicated
                // instruction instead of four predicated commits:
                //   T <- T / (1 - alpha);  s = dL/dpower = 2^ee dL/dalpha (the 0.99 clamp is straight-through);  w = alpha T
                //   dL/dalpha = (c . dp - accdp) T - T_final (bg . dp) / (1 - alpha);  accdp <- accdp + alpha (c . dp - accdp)
                float a_raw = ex2_approx(ee), alpha, rinv;
                asm("{\n\t"
                    ".reg .pred pa;\n\t"
                    ".reg .f32 om;\n\t"
                    "setp.le.f32 pa, %4, 0f00000000;\n\t"
                    "setp.ge.and.f32 pa, %5, 0fC0FFD1BE, pa;\n\t"
                    "setp.lt.and.s32 pa, %6, %7, pa;\n\t"
                    "selp.f32 %0, %0, 0f00000000, pa;\n\t"
                    "min.f32 %1, %0, 0f3F7D70A4;\n\t"
                    "sub.rn.f32 om, 0f3F800000, %1;\n\t"
                    "rcp.approx.ftz.f32 %2, om;\n\t"
                    "@pa mul.rn.f32 %3, %3, %2;\n\t"
                    "}"
                    : "+f"(a_raw), "=f"(alpha), "=f"(rinv), "+f"(T)
                    : "f"(pw), "f"(ee), "r"(gk), "r"(glast));
                const float cdp = fmaf(cb, dp2, fmaf(q1.w, dp1, q1.z * dp0));
                const float dd = cdp - accdp;
                const float dLa = fmaf(dd, T, nTf_bg * rinv);
                accdp = fmaf(alpha, dd, accdp);
                const float s = a_raw * dLa, wgt = alpha * T;
                const uint32_t rowp = row0 + rowoff;
                sts_f1(rowp, s);
                sts_f1(rowp + (RbSmem::Wt - RbSmem::S), wgt);
                // lane `slot` keeps the slot's identity (its list position): one compare + one predicated move
                asm("{\n\t"
                    ".reg .pred ps;\n\t"
                    "setp.eq.u32 ps, %1, %2;\n\t"
                    "@ps mov.u32 %0, %3;\n\t"
                    "}"
                    : "+r"(mine.gk)
                    : "r"(rowoff), "r"(lane_row), "r"(gk));
                rowoff += ROWB;
                if (rowoff == RB_NB * ROWB) {  // 16 slots buffered: switch roles, reduce, send
                    rb_phase2<ABSGRAD>(wbase, RB_NB, lane, px0f, py0f, mine, plist, rec, sgrad);
                    rowoff = 0u;
                }
            }
        }
    }
    const int nbuf = (int)(rowoff / ROWB);
    if (nbuf) rb_phase2<ABSGRAD>(wbase, nbuf, lane, px0f, py0f, mine, plist, rec, sgrad);  // rows k >= nbuf are masked at the RED
}

cudaError_t launch_render_bwd(const Cam& cam, const uint32_t* tile_order, const uint32_t* tile_base,
                              const uint32_t* plist, const float4* rec, const float* final_T, const uint32_t* n_contrib,
                              const float* dL_dpix, float* sgrad, bool absgrad, const uint32_t* info, cudaStream_t st) {
    const int T = cam.gx * cam.gy;
    if (T <= 0) return cudaSuccess;
    // the dynamic shared-memory limit is a per-device, per-function attribute: set it on every launch (cheap, and right
    // for a process that drives several devices or several host threads)
    cudaError_t e;
    if (absgrad) {
        e = cudaFuncSetAttribute(render_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RbSmem::total);
        if (e != cudaSuccess) return e;
        render_bwd_kernel<true><<<T, RB_THREADS, RbSmem::total, st>>>(cam, tile_order, tile_base, plist, rec, final_T,
                                                                      n_contrib, dL_dpix, sgrad, info);
    } else {
        e = cudaFuncSetAttribute(render_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RbSmem::total);
        if (e != cudaSuccess) return e;
        render_bwd_kernel<false><<<T, RB_THREADS, RbSmem::total, st>>>(cam, tile_order, tile_base, plist, rec, final_T,
                                                                       n_contrib, dL_dpix, sgrad, info);
    }
    return cudaGetLastError();
}

}  // namespace dvs
