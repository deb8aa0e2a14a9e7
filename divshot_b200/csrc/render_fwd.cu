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
//   * records are gathered with 128-bit loads only for entries whose mask is non-zero and staged in
//     shared memory, 256 per round, the next round's entry words prefetched into registers;
//   * alpha = ex2(A2 dx^2 + B2 dx dy + C2 dy^2 + lo): log2(e), -1/2 and the opacity are folded into the
//     record, so a pair costs 2 FADD + 2 FMUL + 3 FFMA + 1 MUFU before the blend;
//   * the 1/255 cut-off is tested on the exponent (no MUFU for rejected pairs).
// Bound: issue / MUFU, not HBM (256*D pair evaluations vs 4 B*D + 36 B*V + 20 B*P of traffic).
#include "common.cuh"
#include "kernels.h"

namespace dvs {

constexpr int RF_THREADS = 256;

#ifdef DVS_RF_MINCTA
__global__ void __launch_bounds__(RF_THREADS, DVS_RF_MINCTA)
#else
__global__ void __launch_bounds__(RF_THREADS)
#endif
render_fwd_kernel(Cam cam, const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ plist,
                  const float4* __restrict__ rec, float* __restrict__ out_color, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ info) {
    // staged entries: 48 B each {q0, q1, b, -, -, -}; masks in their own bank-conflict-free array
    __shared__ __align__(16) unsigned char s_stage[RF_THREADS * 48];
    __shared__ uint32_t s_mask[RF_THREADS];
    if (info[2]) return;
    uint32_t sb = smem_u32(s_stage), sm = smem_u32(s_mask);
    asm volatile("" : "+r"(sb), "+r"(sm));  // keep the shared base addresses in registers (no re-derivation per pair)
    const int tile = blockIdx.x;
    const int tx = tile % cam.gx, ty = tile / cam.gx;
    const uint32_t r0 = tile_base[tile], n = tile_base[tile + 1] - r0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = tx * TILE + (warp & 1) * 8 + (lane & 7);
    const int py = ty * TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < cam.W && py < cam.H;
    const f32x2 npxy = pk2(-(float)px, -(float)py);
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);
    const uint32_t wbit = 1u << warp;

    const uint32_t rounds = (n + RF_THREADS - 1) / RF_THREADS;
    uint32_t e_next = (threadIdx.x < n) ? __ldg(plist + r0 + threadIdx.x) : 0u;
    for (uint32_t rd = 0; rd < rounds; rd++) {
        // all 8 warps finished -> tile finished (also the barrier that protects the staging buffers)
        if (__syncthreads_count(warp_done) == RF_THREADS) break;
        const uint32_t e = e_next;
        const uint32_t nxt = (rd + 1) * RF_THREADS + threadIdx.x;
        e_next = (nxt < n) ? __ldg(plist + r0 + nxt) : 0u;
        const uint32_t m = e & 0xffu;
        sts_u1(sm + threadIdx.x * 4, m);
        if (m) {
            const float4* r = rec + 3 * (size_t)(e >> 8);
            const float4 q0 = __ldg(r), q1 = __ldg(r + 1);
            const float b = __ldg(reinterpret_cast<const float*>(r + 2));
            // staged as {mx, my, A2, C2} {B2, lo, r, g}: {mx, my} and {A2, C2} are then register pairs for FADD2 / FMUL2
            sts_f4(sb + threadIdx.x * 48, make_float4(q0.x, q0.y, q0.z, q1.x));
            sts_f4(sb + threadIdx.x * 48 + 16, make_float4(q0.w, q1.y, q1.z, q1.w));
            sts_f1(sb + threadIdx.x * 48 + 32, b);
        }
        __syncthreads();
        if (!warp_done) {
            const uint32_t base_idx = rd * RF_THREADS;
            const int cnt = (int)min((uint32_t)RF_THREADS, n - base_idx);
            for (int c = 0; c < cnt; c += 32) {
                uint32_t bits = __ballot_sync(0xffffffffu, (lds_u1(sm + (c + lane) * 4) & wbit) != 0u);
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int k = c + j;
                    const uint32_t ea = sb + k * 48;
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
                    if (!done && pw <= 0.0f && ee >= ALPHA_MIN_LOG2) {
                        const float alpha = fminf(0.99f, ex2_approx(ee));
                        const float test_T = fmaf(-alpha, T, T);
                        if (test_T < 1e-4f) {
                            done = true;
                        } else {
                            const float w = alpha * T;
                            C0 = fmaf(q1.z, w, C0);
                            C1 = fmaf(q1.w, w, C1);
                            C2 = fmaf(lds_f1(ea + 32), w, C2);
                            T = test_T;
                            last = base_idx + (uint32_t)k + 1u;
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    break;
                }
            }
        }
    }
    if (inside) {
        const size_t P = (size_t)cam.W * cam.H;
        const size_t pix = (size_t)py * cam.W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = fmaf(T, cam.bg[0], C0);
        out_color[P + pix] = fmaf(T, cam.bg[1], C1);
        out_color[2 * P + pix] = fmaf(T, cam.bg[2], C2);
    }
}

cudaError_t launch_render_fwd(const Cam& cam, const uint32_t* tile_base, const uint32_t* plist, const float4* rec,
                              float* out_color, float* final_T, uint32_t* n_contrib, const uint32_t* info,
                              cudaStream_t st) {
    const int T = cam.gx * cam.gy;
    if (T <= 0) return cudaSuccess;
    render_fwd_kernel<<<T, RF_THREADS, 0, st>>>(cam, tile_base, plist, rec, out_color, final_T, n_contrib, info);
    return cudaGetLastError();
}

}  // namespace dvs
