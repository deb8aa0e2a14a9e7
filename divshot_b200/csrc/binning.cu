// binning.cu — A2..A5: duplicate counting -> scan -> emission into per-tile bins -> tile-local sort.
//
// Replaces `InclusiveSum` + `duplicateWithKeys` + `DeviceRadixSort::SortPairs` + `identifyTileRanges` of
// the absent gsplatrast operator with a CUB-free design:
//   * per-tile counts come from the preprocess kernel (RED per (Gaussian,tile));
//   * one CTA scans the T counts with warp-shuffle prefix sums -> tile_base (= `ranges`);
//   * emission claims a slot in its tile's bin with one atomic and writes an 8-byte entry
//       depth_bits<<32 | id<<8 | submask
//     (submask: which of the tile's eight 8x4-pixel sub-rectangles the opacity-aware AABB of the
//     splat {alpha >= 1/255} overlaps — ours, used by A6/A7 to skip work; it never changes the lists);
//   * one CTA per tile sorts its bin in shared memory (normalised bitonic network, all-ascending
//     comparators, virtual +inf padding) and writes the low 32 bits (id<<8|submask) to plist.
// Sorting the 64-bit entry ascending == the credited total order (tile, depth bits, Gaussian index),
// because within a tile ids are unique and the mask sits below the id.
// The order of arrival in a bin is non-deterministic; the sort makes the output deterministic.
//
// Roofline: HBM-light (8 B written + 8 B read + 4 B written per duplicate); sort is shared-memory /
// issue bound.  Algorithmic bytes per SURVEY.md §8(d): 28 B per duplicate (we move 20).
#include "common.cuh"
#include "kernels.h"

namespace dvs {

// ---------------------------------------------------------------------------------------------
// A2/A5: single-CTA exclusive scan over tile counts (T <= ~65k), warp-shuffle based.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(int T, const uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_base,
                 uint32_t* __restrict__ tile_cursor, uint32_t* __restrict__ info, uint32_t dup_capacity) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t warp_max[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    uint32_t mx = 0;
    unsigned long long total64 = 0;
    for (int base = 0; base < T; base += SCAN_THREADS) {
        const int t = base + threadIdx.x;
        const uint32_t c = t < T ? tile_count[t] : 0u;
        mx = max(mx, c);
        uint32_t v = c;  // inclusive warp scan
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += n;
        }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += n;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = carry + (warp ? warp_sums[warp - 1] : 0u) + v - c;
        if (t < T) {
            tile_base[t] = excl;
            tile_cursor[t] = excl;
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) {
            total64 += (unsigned long long)warp_sums[31];
            carry_s = carry + warp_sums[31];
        }
        __syncthreads();
    }
    // max tile length
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 0) warp_max[warp] = mx;
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) {
        uint32_t m = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++) m = max(m, warp_max[w]);
        tile_base[T] = carry_s;
        info[0] = carry_s;
        info[1] = m;
        info[2] = (total64 > (unsigned long long)dup_capacity) ? 1u : 0u;
    }
}

cudaError_t launch_tile_scan(int T, const uint32_t* tile_count, uint32_t* tile_base, uint32_t* tile_cursor,
                             uint32_t* info, uint32_t dup_capacity, cudaStream_t st) {
    tile_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(T, tile_count, tile_base, tile_cursor, info, dup_capacity);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// A3: emission.  One lane per Gaussian for small rects; rects with more than 32 tiles are
// spread over the whole warp (upstream's per-thread loop serialises on them).
// ---------------------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;

__device__ __forceinline__ uint32_t sub_mask(float mx, float my, float ex, float ey, int tx, int ty) {
    // 8 sub-rectangles of 8x4 pixels: bit w -> x in [X0, X0+7], y in [Y0, Y0+3],
    // X0 = 16*tx + 8*(w&1), Y0 = 16*ty + 4*(w>>1).  AABB overlap factorises into columns x rows.
    const float X0 = (float)(tx * TILE), Y0 = (float)(ty * TILE);
    const float xl = mx - ex, xh = mx + ex, yl = my - ey, yh = my + ey;
    uint32_t col = 0, rowm = 0;
    if (xl <= X0 + 7.0f && xh >= X0) col |= 1u;
    if (xl <= X0 + 15.0f && xh >= X0 + 8.0f) col |= 2u;
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (yl <= Y0 + (float)(4 * r + 3) && yh >= Y0 + (float)(4 * r)) rowm |= 1u << r;
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (rowm & (1u << r)) m |= col << (2 * r);
    return (ex < 0.0f) ? 0u : m;
}

__device__ __forceinline__ void emit_one(int id, uint32_t depth_bits, float mx, float my, float ex, float ey,
                                         int tx, int ty, int gx, uint32_t* tile_cursor, unsigned long long* bins,
                                         uint32_t cap) {
    const uint32_t m = sub_mask(mx, my, ex, ey, tx, ty);
    const uint32_t slot = atomicAdd(tile_cursor + ty * gx + tx, 1u);
    if (slot < cap)
        bins[slot] = ((unsigned long long)depth_bits << 32) | ((unsigned long long)((uint32_t)id << 8) | m);
}

__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(Cam cam, int N, const float4* __restrict__ rec, const uint4* __restrict__ aux,
            uint32_t* __restrict__ tile_cursor, unsigned long long* __restrict__ bins, uint32_t cap) {
    const int i = blockIdx.x * EMIT_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint4 ax = make_uint4(0, 0, 0, 0);
    float mx = 0.f, my = 0.f;
    uint32_t depth_bits = 0;
    int minx = 0, miny = 0, maxx = 0, maxy = 0;
    if (i < N) {
        ax = __ldg(aux + i);
        minx = ax.x & 0xffff; miny = ax.x >> 16; maxx = ax.y & 0xffff; maxy = ax.y >> 16;
    }
    const int w = maxx - minx, h = maxy - miny;
    const int area = w * h;
    if (area > 0) {
        const float4 q0 = __ldg(rec + 3 * (size_t)i);
        const float4 q2 = __ldg(rec + 3 * (size_t)i + 2);
        mx = q0.x; my = q0.y;
        depth_bits = __float_as_uint(q2.y);
    }
    const float ex = __uint_as_float(ax.z), ey = __uint_as_float(ax.w);
    if (area > 0 && area <= 32) {
        for (int y = miny; y < maxy; y++)
            for (int x = minx; x < maxx; x++)
                emit_one(i, depth_bits, mx, my, ex, ey, x, y, cam.gx, tile_cursor, bins, cap);
    }
    // warp-cooperative path for big rects
    unsigned big = __ballot_sync(0xffffffffu, area > 32);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int b_id = __shfl_sync(0xffffffffu, i, src);
        const uint32_t b_depth = __shfl_sync(0xffffffffu, depth_bits, src);
        const float b_mx = __shfl_sync(0xffffffffu, mx, src), b_my = __shfl_sync(0xffffffffu, my, src);
        const float b_ex = __shfl_sync(0xffffffffu, ex, src), b_ey = __shfl_sync(0xffffffffu, ey, src);
        const int b_minx = __shfl_sync(0xffffffffu, minx, src), b_miny = __shfl_sync(0xffffffffu, miny, src);
        const int b_w = __shfl_sync(0xffffffffu, w, src), b_area = __shfl_sync(0xffffffffu, area, src);
        for (int k = lane; k < b_area; k += 32)
            emit_one(b_id, b_depth, b_mx, b_my, b_ex, b_ey, b_minx + k % b_w, b_miny + k / b_w, cam.gx,
                     tile_cursor, bins, cap);
    }
}

cudaError_t launch_emit(const Cam& cam, int N, const float4* rec, const uint4* aux, uint32_t* tile_cursor,
                        unsigned long long* bins, uint32_t dup_capacity, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    emit_kernel<<<(N + EMIT_THREADS - 1) / EMIT_THREADS, EMIT_THREADS, 0, st>>>(cam, N, rec, aux, tile_cursor,
                                                                                 bins, dup_capacity);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// A4: tile-local sort.  Normalised bitonic network: every comparator is ascending, so elements
// at virtual indices >= n (= +inf) never move and no physical padding is needed.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cmpx(T* a, int i, int j) {
    const T x = a[i], y = a[j];
    if (x > y) {
        a[i] = y;
        a[j] = x;
    }
}

template <int THREADS>
__device__ void bitonic_sort_u64(unsigned long long* a, int n) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    const int half = np2 >> 1;
    for (int k = 2; k <= np2; k <<= 1) {
        const int hk = k >> 1;
        for (int t = threadIdx.x; t < half; t += THREADS) {  // flip stage
            const int blk = t / hk, off = t - blk * hk;
            const int i = blk * k + off, j = blk * k + (k - 1 - off);
            if (j < n) cmpx(a, i, j);
        }
        __syncthreads();
        for (int s = hk >> 1; s > 0; s >>= 1) {  // half-cleaner stages
            for (int t = threadIdx.x; t < half; t += THREADS) {
                const int i = 2 * t - (t & (s - 1));
                const int j = i + s;
                if (j < n) cmpx(a, i, j);
            }
            __syncthreads();
        }
    }
}

// CTA handles its tile iff lo < n <= hi.  IN_SMEM: stage through dynamic shared memory.
template <int THREADS, bool IN_SMEM>
__global__ void __launch_bounds__(THREADS)
tile_sort_kernel(const uint32_t* __restrict__ tile_base, unsigned long long* __restrict__ bins,
                 uint32_t* __restrict__ plist, const uint32_t* __restrict__ info, uint32_t lo, uint32_t hi) {
    extern __shared__ unsigned long long s_keys[];
    if (info[2]) return;                  // arena overflow: forward is re-run by the host
    if (info[1] <= lo) return;            // no tile is this long
    const uint32_t b0 = tile_base[blockIdx.x], b1 = tile_base[blockIdx.x + 1];
    const uint32_t n = b1 - b0;
    if (n <= lo || n > hi) return;
    unsigned long long* g = bins + b0;
    if (IN_SMEM) {
        for (uint32_t t = threadIdx.x; t < n; t += THREADS) s_keys[t] = g[t];
        __syncthreads();
        bitonic_sort_u64<THREADS>(s_keys, (int)n);
        for (uint32_t t = threadIdx.x; t < n; t += THREADS) plist[b0 + t] = (uint32_t)s_keys[t];
    } else {
        __syncthreads();
        bitonic_sort_u64<THREADS>(g, (int)n);  // in place in global/L2 (rare: > 16384 entries in one tile)
        for (uint32_t t = threadIdx.x; t < n; t += THREADS) plist[b0 + t] = (uint32_t)g[t];
    }
}

cudaError_t launch_tile_sort(int T, const uint32_t* tile_base, unsigned long long* bins, uint32_t* plist,
                             const uint32_t* info, cudaStream_t st) {
    if (T <= 0) return cudaSuccess;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(tile_sort_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        cudaFuncSetAttribute(tile_sort_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        attr_done = true;
    }
    tile_sort_kernel<256, true><<<T, 256, 1024 * 8, st>>>(tile_base, bins, plist, info, 0u, 1024u);
    tile_sort_kernel<512, true><<<T, 512, 4096 * 8, st>>>(tile_base, bins, plist, info, 1024u, 4096u);
    tile_sort_kernel<1024, true><<<T, 1024, 16384 * 8, st>>>(tile_base, bins, plist, info, 4096u, 16384u);
    tile_sort_kernel<1024, false><<<T, 1024, 0, st>>>(tile_base, bins, plist, info, 16384u, 0xffffffffu);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// debug unpackers (parity tests)
// ---------------------------------------------------------------------------------------------
__global__ void unpack_kernel(int N, const float4* __restrict__ rec, int32_t* radii, uint32_t* tiles, float* depth,
                              float* mean2D, float* conic_opacity, float* rgb, uint8_t* clamped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 q0 = rec[3 * (size_t)i], q1 = rec[3 * (size_t)i + 1], q2 = rec[3 * (size_t)i + 2];
    const uint32_t tw = __float_as_uint(q2.w);
    const int rad = __float_as_int(q2.z);
    radii[i] = rad;
    tiles[i] = tw & 0xffffffu;
    depth[i] = q2.y;
    mean2D[2 * i] = q0.x; mean2D[2 * i + 1] = q0.y;
    conic_opacity[4 * i] = rad > 0 ? q0.z / (-0.5f * LOG2E) : 0.f;
    conic_opacity[4 * i + 1] = rad > 0 ? q0.w / (-LOG2E) : 0.f;
    conic_opacity[4 * i + 2] = rad > 0 ? q1.x / (-0.5f * LOG2E) : 0.f;
    conic_opacity[4 * i + 3] = rad > 0 ? exp2f(q1.y) : 0.f;
    rgb[3 * i] = q1.z; rgb[3 * i + 1] = q1.w; rgb[3 * i + 2] = q2.x;
    clamped[3 * i] = (tw >> 24) & 1; clamped[3 * i + 1] = (tw >> 25) & 1; clamped[3 * i + 2] = (tw >> 26) & 1;
}
cudaError_t launch_unpack(int N, const float4* rec, int32_t* radii, uint32_t* tiles, float* depth, float* mean2D,
                          float* conic_opacity, float* rgb, uint8_t* clamped, cudaStream_t st) {
    if (N > 0) unpack_kernel<<<(N + 255) / 256, 256, 0, st>>>(N, rec, radii, tiles, depth, mean2D, conic_opacity, rgb, clamped);
    return cudaGetLastError();
}
__global__ void unpack_plist_kernel(uint32_t D, const uint32_t* __restrict__ plist, uint32_t* ids, uint8_t* masks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    const uint32_t e = plist[i];
    if (ids) ids[i] = e >> 8;
    if (masks) masks[i] = (uint8_t)(e & 0xff);
}
cudaError_t launch_unpack_plist(uint32_t D, const uint32_t* plist, uint32_t* ids, uint8_t* masks, cudaStream_t st) {
    if (D > 0) unpack_plist_kernel<<<(D + 255) / 256, 256, 0, st>>>(D, plist, ids, masks);
    return cudaGetLastError();
}

}  // namespace dvs
