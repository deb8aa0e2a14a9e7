// emit.cuh — sub-tile cull masks and the warp-cooperative duplicate emission (A3), shared by the two-pass
// emit_kernel (binning.cu) and the fused single-pass mode of preprocess_fwd.cu.
#pragma once
#include "common.cuh"

namespace dvs {

// ---------------------------------------------------------------------------------------------
// Sub-tile cull mask (ours; no upstream analogue).  For the entry (tile, Gaussian): which of the tile's
// eight 8x4-pixel sub-rectangles (bit w: x in [16tx+8(w&1), +7], y in [16ty+4(w>>1), +3]) can hold a
// pixel with alpha >= 1/255, i.e. Q(d) = a dx^2 + b dx dy + c dy^2 <= m for some d = mean - pixel in the
// box, with (a,b,c) = -(A2,B2,C2) and m = lo - log2(1/255) (+ a small conservative margin).  Q is convex
// with its minimum at d = 0, so the box minimum is on the edges facing the origin:
//   q = min( Q(ex, clamp(-b ex / 2c)),  Q(clamp(-b ey / 2a), ey) ),  (ex, ey) = box point nearest to 0
// evaluated branch-free for all 8 boxes (2 column ranges x 4 row ranges).  It is computed at emission —
// that kernel is bound by the L2 atomic rate and has ~80 % of its issue slots free — and rides in the low
// byte of the sort key, below the Gaussian id.
// ---------------------------------------------------------------------------------------------
struct CullParams {  // per Gaussian
    float mx, my, a, b, c, m, hbc, hba;
};
__device__ __forceinline__ CullParams cull_params(const float4 q0, const float4 q1) {
    CullParams p;
    p.mx = q0.x; p.my = q0.y;
    p.a = -q0.z; p.b = -q1.x; p.c = -q0.w;  // record layout {mx, my, A2, C2} {B2, lo, r, g}
    p.m = (q1.y - ALPHA_MIN_LOG2) * 1.0001f + 1e-3f;
    const bool ok = p.a > 0.0f && p.c > 0.0f;
    p.hbc = ok ? __fdividef(-0.5f * p.b, p.c) : 0.0f;
    p.hba = ok ? __fdividef(-0.5f * p.b, p.a) : 0.0f;
    if (!ok) p.m = 3.0e38f;  // degenerate conic: no culling (every box passes)
    return p;
}
__device__ __forceinline__ uint32_t sub_tile_mask(const CullParams& g, float X0, float Y0) {
    if (!(g.m > 0.0f)) return 0u;  // opacity < 1/255: never contributes
    const float mx = g.mx - X0, my = g.my - Y0;  // mean relative to the tile origin
    uint32_t mask = 0;
    float dxlo[2], dxhi[2], exn[2];
#pragma unroll
    for (int cx = 0; cx < 2; cx++) {  // d = mean - pixel, pixel x in [8cx, 8cx+7]
        dxhi[cx] = mx - (float)(8 * cx);
        dxlo[cx] = dxhi[cx] - 7.0f;
        exn[cx] = fminf(fmaxf(0.0f, dxlo[cx]), dxhi[cx]);
    }
#pragma unroll
    for (int ry = 0; ry < 4; ry++) {
        const float dyhi = my - (float)(4 * ry), dylo = dyhi - 3.0f;
        const float eyn = fminf(fmaxf(0.0f, dylo), dyhi);
        const float dxs = g.hba * eyn;  // unclamped minimiser along the horizontal edge
#pragma unroll
        for (int cx = 0; cx < 2; cx++) {
            const float ex = exn[cx];
            const float dy = fminf(fmaxf(g.hbc * ex, dylo), dyhi);
            const float q1v = fmaf(g.a * ex, ex, fmaf(g.b, ex, g.c * dy) * dy);
            const float dx = fminf(fmaxf(dxs, dxlo[cx]), dxhi[cx]);
            const float q2v = fmaf(g.c * eyn, eyn, fmaf(g.b, eyn, g.a * dx) * dx);
            if (fminf(q1v, q2v) <= g.m) mask |= 1u << (2 * ry + cx);
        }
    }
    return mask;
}

// ---------------------------------------------------------------------------------------------
// A3: emission.  Warp-cooperative: the 32 Gaussians of a warp flatten their tile rects into one work
// list (warp-shuffle prefix scan of the duplication counts) and every lane takes every 32nd item, so a
// big splat does not serialise its warp (upstream's per-thread loop does).  Each item claims a slot in its
// tile's bin with one atomic and writes  depth_bits<<32 | id<<8 | sub-tile mask.
//   bin_stride == 0 : two-pass mode, tile_cursor[t] was initialised to the tile's base offset (scan of the
//                     counts), slot is a global index, `cap` = arena capacity;
//   bin_stride  > 0 : single-pass mode, tile_cursor[t] starts at 0, tile t owns bins[t*bin_stride ..), a slot
//                     >= bin_stride sets the sticky overflow word (the host redoes the step in two-pass mode).
// All 32 lanes of the warp must call (lanes without work pass area = 0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_emit(int gx, int id, int minx, int miny, int w, int area, uint32_t depth_bits,
                                          const CullParams& cp, uint32_t* __restrict__ tile_cursor,
                                          unsigned long long* __restrict__ bins, uint32_t bin_stride, uint32_t cap,
                                          uint32_t* __restrict__ overflow_word, bool tight = false) {
    const int lane = threadIdx.x & 31;
    int incl = area;  // inclusive warp scan of the duplication counts
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    constexpr int EMIT_UNROLL = 2;
    for (int j0 = 0; j0 < total; j0 += 32 * EMIT_UNROLL) {
        uint32_t slot[EMIT_UNROLL], tile[EMIT_UNROLL];
        unsigned long long key[EMIT_UNROLL];
        bool ok[EMIT_UNROLL];
#pragma unroll
        for (int u = 0; u < EMIT_UNROLL; u++) {
            const int j = j0 + 32 * u + lane;
            // source lane = number of lanes whose inclusive offset is <= j (binary search over the sorted offsets)
            int pos = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, pos + step - 1);
                if (v <= j) pos += step;
            }
            const int src = min(pos, 31);
            const int s_incl = __shfl_sync(0xffffffffu, incl, src);
            const int s_area = __shfl_sync(0xffffffffu, area, src);
            const int s_minx = __shfl_sync(0xffffffffu, minx, src), s_miny = __shfl_sync(0xffffffffu, miny, src);
            const int s_w = __shfl_sync(0xffffffffu, w, src);
            const uint32_t s_depth = __shfl_sync(0xffffffffu, depth_bits, src);
            const int s_id = __shfl_sync(0xffffffffu, id, src);
            CullParams g;
            g.mx = __shfl_sync(0xffffffffu, cp.mx, src); g.my = __shfl_sync(0xffffffffu, cp.my, src);
            g.a = __shfl_sync(0xffffffffu, cp.a, src); g.b = __shfl_sync(0xffffffffu, cp.b, src);
            g.c = __shfl_sync(0xffffffffu, cp.c, src); g.m = __shfl_sync(0xffffffffu, cp.m, src);
            g.hbc = __shfl_sync(0xffffffffu, cp.hbc, src); g.hba = __shfl_sync(0xffffffffu, cp.hba, src);
            ok[u] = j < total;
            slot[u] = 0xffffffffu; tile[u] = 0u;
            key[u] = 0ull;
            if (ok[u]) {
                const int k = j - (s_incl - s_area);
                // row / column of item k inside the rect (k / w, k % w) without the ~20-instruction integer division:
                // float reciprocal estimate (k + 0.5 keeps the quotient away from integers by >= 0.5 / w, far above the
                // rounding error for k < 2^24), then an exact integer fix-up so the result never depends on rounding
                int q = __float2int_rz(((float)k + 0.5f) * __frcp_rn((float)s_w));
                int r = k - q * s_w;
                if (r < 0) { q -= 1; r += s_w; }
                if (r >= s_w) { q += 1; r -= s_w; }
                const int ty = s_miny + q, tx = s_minx + r;
                tile[u] = (uint32_t)(ty * gx + tx);
                // the mask computation overlaps the atomic's round trip — unless the lists are TIGHT (DVS_FLAG_TIGHT_LISTS,
                // single-pass mode only): an entry whose sub-tile mask is empty (no pixel of the tile reaches
                // alpha >= 1/255; ~39 % of D at c3, profiles/r1_pair_counts.md) is then not emitted at all — fewer
                // atomics, shorter sorts and list scans, the same image and gradients; the lists are the reference's
                // whole-rectangle lists with exactly those entries removed (tested), and n_contrib counts in them
                uint32_t m8;
                if (tight) {
                    m8 = sub_tile_mask(g, (float)(tx * TILE), (float)(ty * TILE));
                    if (m8 == 0u) { ok[u] = false; continue; }
                    slot[u] = atomicAdd(tile_cursor + (size_t)tile[u] * TILE_CTR_STRIDE, 1u);
                } else {
                    slot[u] = atomicAdd(tile_cursor + (size_t)tile[u] * TILE_CTR_STRIDE, 1u);
                    m8 = sub_tile_mask(g, (float)(tx * TILE), (float)(ty * TILE));
                }
                key[u] = ((unsigned long long)s_depth << 32) | (((uint32_t)s_id << 8) | m8);
            }
        }
#pragma unroll
        for (int u = 0; u < EMIT_UNROLL; u++) {
            if (!ok[u]) continue;
            if (bin_stride) {
                if (slot[u] < bin_stride) bins[(size_t)tile[u] * bin_stride + slot[u]] = key[u];
                else *overflow_word = 1u;
            } else if (slot[u] < cap) {
                bins[slot[u]] = key[u];
            }
        }
    }
}

}  // namespace dvs
