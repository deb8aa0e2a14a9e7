// render_fwd.cu — A6: tile-based front-to-back alpha compositing.
//
// Replaces `renderCUDA` (forward) of the absent gsplatrast operator (SURVEY.md §8 A6, Appendix B.3;
// alpha falloff / 0.99 clamp / 1-255 cut-off mirrored in-tree at
// diverse/assets/shaders/gaussian/gsplat_ps.hlsl:60-66,85; 16x16 tile constants gaussian_common.hlsl:162-163).
//
// B200 design (not the upstream kernel):
//   * one CTA per 16x16 tile, 8 warps; warp w owns the 8x4-pixel sub-rectangle (w&1, w>>1), so the
//     per-entry 8-bit sub-tile mask computed at emission lets a whole warp skip a splat whose
//     {alpha >= 1/255} footprint misses its 32 pixels — one ballot per 32 entries, no per-pixel work;
//   * CTAs take their tile from `tile_order` (longest lists first, written by the tile scan): the last wave of
//     CTAs is made of short tiles instead of whatever the raster order leaves;
//   * records are staged in shared memory 256 per round by 16-byte ASYNCHRONOUS copies (cp.async / LDGSTS, only for
//     entries whose mask is non-zero), double-buffered: round r+1 streams in while round r is composited, one CTA
//     barrier per round, no staging registers (the record layout in HBM is the layout the loop reads);
//   * alpha = ex2(A2 dx^2 + B2 dx dy + C2 dy^2 + lo): log2(e), -1/2 and the opacity are folded into the
//     record, so a pair costs 1 FADD2 + 1 FMUL2 + 1 FMUL + 2 FFMA + 1 FADD before the blend;
//   * the blend is one straight-line predicated sequence (no divergent branch, no per-lane `done` flag): the
//     transmittance is updated by EVERY pair that passes the alpha test, so a pixel that has stopped
//     (T' < 1e-4) keeps a T below the threshold and can never blend again, while `Tb` (T after the last
//     blended pair) is what final_T reports.  27 issue slots per (warp, splat) visit (round 1: 37).
// Bound: shared-memory bandwidth (the 36-byte record is broadcast to 32 lanes = 9 LSU wavefronts per visit)
// and issue, not HBM (256*D pair evaluations vs 4 B*D + 36 B*V + 20 B*P of traffic).
#include "common.cuh"
#include "kernels.h"

namespace dvs {

constexpr int RF_THREADS = 256;
constexpr int RF_STAGE = RF_THREADS * 48;  // bytes of one staging buffer (one 48-byte record per thread and round)

#ifdef DVS_RF_MINCTA
__global__ void __launch_bounds__(RF_THREADS, DVS_RF_MINCTA)
#else
__global__ void __launch_bounds__(RF_THREADS)
#endif
render_fwd_kernel(Cam cam, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ tile_base,
                  const uint32_t* __restrict__ plist, const float4* __restrict__ rec, float* __restrict__ out_color,
                  float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ info) {
    // two staging buffers of 256 records {mx, my, A2, C2} {B2, lo, r, g} {b, depth, radius, tiles} (the record as it lies
    // in HBM) + the entries' mask bytes; round r+1 is copied in asynchronously (LDGSTS) while round r is composited
    __shared__ __align__(16) unsigned char s_stage[2 * RF_STAGE];
    __shared__ uint32_t s_mask[2 * RF_THREADS];
    if (info[2]) return;
    uint32_t sb = smem_u32(s_stage), sm = smem_u32(s_mask);
    asm volatile("" : "+r"(sb), "+r"(sm));  // keep the shared base addresses in registers (no re-derivation per pair)
    const int tile = tile_order ? (int)tile_order[blockIdx.x] : (int)blockIdx.x;
    const int tx = tile % cam.gx, ty = tile / cam.gx;
    const uint32_t r0 = tile_base[tile], n = tile_base[tile + 1] - r0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const f32x2 npxy = pk2(-(float)px, -(float)py);
    // T: running transmittance, also the stop state (T < 1e-4 is absorbing: every later T' = T (1 - alpha) stays below);
    // Tb: transmittance after the last pair that was actually blended (= final_T).  Pixels outside the image start stopped.
    float T = inside ? 1.0f : 0.0f, Tb = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;
    bool warp_done = __all_sync(0xffffffffu, !inside);
    const uint32_t wbit = 1u << warp;

    const uint32_t rounds = (n + RF_THREADS - 1) / RF_THREADS;
    // stage(e, buf): this thread's entry of a round -> buffer `buf` (mask by a plain store, record by three 16-byte LDGSTS)
    auto stage = [&](uint32_t e, uint32_t buf) {
        sts_u1(sm + (buf * RF_THREADS + threadIdx.x) * 4, e & 0xffu);
        if (e & 0xffu) {
            const float4* r = rec + 3 * (size_t)(e >> 8);
            const uint32_t dst = sb + buf * RF_STAGE + threadIdx.x * 48;
            cp_async16(dst, r);
            cp_async16(dst + 16, r + 1);
            cp_async16(dst + 32, r + 2);
        }
        cp_async_commit();
    };
    uint32_t e_next = 0;
    if (rounds) {
        stage((threadIdx.x < n) ? __ldg(plist + r0 + threadIdx.x) : 0u, 0u);
        e_next = (RF_THREADS + threadIdx.x < n) ? __ldg(plist + r0 + RF_THREADS + threadIdx.x) : 0u;
    }
    for (uint32_t rd = 0; rd < rounds; rd++) {
        cp_async_wait0();  // my copies of round rd have landed
        // barrier: everyone's copies have landed AND everyone has finished round rd-1 (whose buffer is refilled below);
        // all 8 warps finished -> tile finished
        if (__syncthreads_count(warp_done) == RF_THREADS) break;
        if (rd + 1 < rounds) {
            stage(e_next, (rd + 1) & 1u);
            const uint32_t nxt = (rd + 2) * RF_THREADS + threadIdx.x;
            e_next = (nxt < n) ? __ldg(plist + r0 + nxt) : 0u;
        }
        if (!warp_done) {
            const uint32_t base_idx = rd * RF_THREADS;
            const int cnt = (int)min((uint32_t)RF_THREADS, n - base_idx);
            const uint32_t sbr = sb + (rd & 1u) * RF_STAGE, smr = sm + (rd & 1u) * (RF_THREADS * 4);
            uint32_t last_k = 0xffffffffu;  // index in this round of the last blended entry (none yet)
            for (int c = 0; c < cnt; c += 32) {
                // bit-reversed ballot: the highest set bit is the FIRST entry of the group (one FLO per visit finds it)
                uint32_t br = __brev(__ballot_sync(0xffffffffu, (lds_u1(smr + (c + lane) * 4) & wbit) != 0u));
                while (br) {
                    uint32_t p;
                    asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(br));  // FLO, no clz round trip
                    {
                        uint32_t below;
                        asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(below) : "r"(0), "r"(p));  // bits [0, p)
                        br &= below;
                    }
                    const uint32_t kk = (uint32_t)(c + 31) - p;    // entry index in the round
                    const uint32_t ea = sbr + kk * 48u;
                    f32x2 mxy, AC;
                    lds_p4(ea, mxy, AC);
                    const float4 q1 = lds_f4(ea + 16);  // {B2, lo, r, g}
                    float dx, dy, adx, cdy;
                    const f32x2 d = add2(mxy, npxy);    // {dx, dy}: one FADD2
                    upk2(d, dx, dy);
                    upk2(mul2(AC, d), adx, cdy);        // {A2 dx, C2 dy}: one FMUL2
                    const float t = fmaf(q1.x, dy, adx);
                    const float pw = fmaf(cdy, dy, t * dx);
                    const float ee = pw + q1.y;
                    // alpha = min(0.99, 2^ee), w = alpha T (computed for every lane, used under the predicates);
                    // pa = power <= 0 && alpha >= 1/255 (tested on the exponent): T -= w;
                    // pb = pa && T >= 1e-4: blend (C += rgb w), Tb = T, last = this entry
                    const float w = fminf(0.99f, ex2_approx(ee)) * T;
                    asm volatile(
                        "{\n\t"
                        ".reg .pred pa, pb;\n\t"
                        ".reg .f32 bl;\n\t"
                        "setp.le.f32 pa, %7, 0f00000000;\n\t"
                        "setp.ge.and.f32 pa, %6, 0fC0FFD1BE, pa;\n\t"
                        "@pa sub.rn.f32 %0, %0, %11;\n\t"
                        "setp.ge.and.f32 pb, %0, 0f38D1B717, pa;\n\t"
                        "@pb ld.shared.f32 bl, [%8+32];\n\t"
                        "@pb fma.rn.f32 %2, %9, %11, %2;\n\t"
                        "@pb fma.rn.f32 %3, %10, %11, %3;\n\t"
                        "@pb fma.rn.f32 %4, bl, %11, %4;\n\t"
                        "@pb mov.f32 %1, %0;\n\t"
                        "@pb mov.u32 %5, %12;\n\t"
                        "}"
                        : "+f"(T), "+f"(Tb), "+f"(C0), "+f"(C1), "+f"(C2), "+r"(last_k)
                        : "f"(ee), "f"(pw), "r"(ea), "f"(q1.z), "f"(q1.w), "f"(w), "r"(kk)
                        : "memory");
                }
                if (__all_sync(0xffffffffu, T < 1e-4f)) {
                    warp_done = true;
                    break;
                }
            }
            if (last_k != 0xffffffffu) last = base_idx + last_k + 1u;
        }
    }
    if (inside) {
        const size_t P = (size_t)cam.W * cam.H;
        const size_t pix = (size_t)py * cam.W + px;
        final_T[pix] = Tb;
        n_contrib[pix] = last;
        float b0 = cam.bg[0], b1 = cam.bg[1], b2 = cam.bg[2];
        if (cam.bg_image) {  // per-pixel background (sky model): out = C + T_final * bg(pixel)
            b0 = __ldg(cam.bg_image + pix); b1 = __ldg(cam.bg_image + P + pix); b2 = __ldg(cam.bg_image + 2 * P + pix);
        }
        out_color[pix] = fmaf(Tb, b0, C0);
        out_color[P + pix] = fmaf(Tb, b1, C1);
        out_color[2 * P + pix] = fmaf(Tb, b2, C2);
    }
}

cudaError_t launch_render_fwd(const Cam& cam, const uint32_t* tile_order, const uint32_t* tile_base,
                              const uint32_t* plist, const float4* rec, float* out_color, float* final_T,
                              uint32_t* n_contrib, const uint32_t* info, cudaStream_t st) {
    const int T = cam.gx * cam.gy;
    if (T <= 0) return cudaSuccess;
    render_fwd_kernel<<<T, RF_THREADS, 0, st>>>(cam, tile_order, tile_base, plist, rec, out_color, final_T, n_contrib,
                                                info);
    return cudaGetLastError();
}

}  // namespace dvs
