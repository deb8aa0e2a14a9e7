// binning.cu — A2..A5: duplicate counting -> scan -> emission into per-tile bins -> tile-local sort.
//
// Replaces `InclusiveSum` + `duplicateWithKeys` + `DeviceRadixSort::SortPairs` + `identifyTileRanges` of
// the absent gsplatrast operator with a CUB-free design:
//   * per-tile counts come from the preprocess kernel (RED per (Gaussian,tile));
//   * one CTA scans the T counts with warp-shuffle prefix sums -> tile_base (= `ranges`);
//   * emission claims a slot in its tile's bin with one atomic and writes an 8-byte entry
//       depth_bits<<32 | id<<8 | submask
//     where submask says which of the tile's eight 8x4-pixel sub-rectangles the splat's {alpha >= 1/255}
//     ellipse can reach (ours, used by A6/A7 to skip work; it never changes the lists);
//   * one CTA per tile sorts its bin in shared memory (adaptive two-level bucket sort; bitonic network as
//     the always-exact fallback) and writes the low word (id<<8 | submask) to plist.
// Sorting the 64-bit entry ascending == the credited total order (tile, depth bits, Gaussian index), because
// within a tile ids are unique and the mask sits below the id.
// The order of arrival in a bin is non-deterministic; the sort makes the output deterministic.
//
// Roofline: HBM-light (8 B written + 8 B read + 4 B written per duplicate); sort is shared-memory /
// issue bound.  Algorithmic bytes per SURVEY.md §8(d): 28 B per duplicate (we move 20).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"
#include "emit.cuh"
#include "kernels.h"

namespace dvs {
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------
// A2/A5: single-CTA exclusive scan over tile counts (T <= ~65k), warp-shuffle based.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;

// sort classes by tile-list length (see tile sort below)
constexpr int NUM_SORT_CLASSES = 5;
__host__ __device__ __forceinline__ int sort_class_of(uint32_t n) {
    return n <= 1024u ? 0 : n <= 2048u ? 1 : n <= 4096u ? 2 : n <= 16384u ? 3 : 4;
}

constexpr int SCAN_MAX_PER_THREAD = 8;  // register-blocked fast path for T <= 8192 tiles
// launch order of the compositing CTAs: a counting sort of the tiles by list length (descending, ORDER_BUCKETS buckets
// of ORDER_GRAIN entries, the last one open-ended), so that the last wave of CTAs is made of short lists
constexpr int ORDER_BUCKETS = 128;
constexpr uint32_t ORDER_GRAIN = 32;
__device__ __forceinline__ uint32_t order_bucket(uint32_t len) {  // bucket 0 = longest
    return (uint32_t)(ORDER_BUCKETS - 1) - min(len / ORDER_GRAIN, (uint32_t)(ORDER_BUCKETS - 1));
}

__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(int T, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_base,
                 uint32_t* __restrict__ tile_cursor /* nullptr: single-pass mode, counts ARE the cursors */,
                 uint32_t* __restrict__ info, uint32_t dup_capacity,
                 uint32_t* __restrict__ class_tiles /* [NUM_SORT_CLASSES][T] */,
                 uint32_t* __restrict__ tile_order /* [T] or nullptr: all tiles, longest lists first */,
                 unsigned long long* __restrict__ stats /* [2]: V, D accumulated by the preprocess kernel */) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t warp_max[32];
    __shared__ uint32_t s_hist[ORDER_BUCKETS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // housekeeping that used to be four memset launches per forward: the per-class counters are zeroed here, the tile
    // counters are zeroed as they are read (they are dead afterwards: the sort works from tile_base), and thread 0
    // publishes and clears the forward's V / D accumulators and the bin-overflow word at the end
    if (threadIdx.x < NUM_SORT_CLASSES) info[4 + threadIdx.x] = 0u;
    __syncthreads();
    // chunked over [0,T) in slabs of SCAN_THREADS * SCAN_MAX_PER_THREAD; within a slab thread t owns the
    // SCAN_MAX_PER_THREAD consecutive tiles starting at slab + t * SCAN_MAX_PER_THREAD (all loads in flight at once)
    unsigned long long carry = 0;
    uint32_t mx = 0;
    for (int slab = 0; slab < T; slab += SCAN_THREADS * SCAN_MAX_PER_THREAD) {
        const int first = slab + threadIdx.x * SCAN_MAX_PER_THREAD;
        uint32_t c[SCAN_MAX_PER_THREAD];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER_THREAD; k++) {
            const int t = first + k;
            c[k] = t < T ? tile_count[(size_t)t * TILE_CTR_STRIDE] : 0u;
        }
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER_THREAD; k++)
            if (first + k < T && c[k]) tile_count[(size_t)(first + k) * TILE_CTR_STRIDE] = 0u;
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER_THREAD; k++) { sum += c[k]; mx = max(mx, c[k]); }
        uint32_t v = sum;  // inclusive warp scan of the per-thread sums
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
        uint32_t run = (uint32_t)carry + (warp ? warp_sums[warp - 1] : 0u) + v - sum;
        const uint32_t slab_total = warp_sums[31];
        // per-class work lists for the sort kernels (info[4+q] = count of class q): count my tiles per class,
        // warp-scan the counts, ONE atomic per (warp, class) — the five atomics are independent and overlap
        uint32_t mycnt[NUM_SORT_CLASSES];
#pragma unroll
        for (int q = 0; q < NUM_SORT_CLASSES; q++) mycnt[q] = 0;
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER_THREAD; k++)
            if (first + k < T && c[k]) {
                const int cls = sort_class_of(c[k]);
#pragma unroll
                for (int q = 0; q < NUM_SORT_CLASSES; q++) mycnt[q] += (cls == q) ? 1u : 0u;
            }
        uint32_t slot[NUM_SORT_CLASSES];
#pragma unroll
        for (int q = 0; q < NUM_SORT_CLASSES; q++) {
            uint32_t inc = mycnt[q];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += n;
            }
            const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
            uint32_t b = 0;
            if (lane == 0 && wtot) b = atomicAdd(info + 4 + q, wtot);
            slot[q] = __shfl_sync(0xffffffffu, b, 0) + inc - mycnt[q];
        }
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER_THREAD; k++) {
            const int t = first + k;
            if (t < T) {
                tile_base[t] = run;
                if (tile_cursor) tile_cursor[(size_t)t * TILE_CTR_STRIDE] = run;
                if (c[k]) {
                    const int cls = sort_class_of(c[k]);
#pragma unroll
                    for (int q = 0; q < NUM_SORT_CLASSES; q++)
                        if (cls == q) class_tiles[(size_t)q * T + slot[q]++] = (uint32_t)t;
                }
            }
            run += c[k];
        }
        carry += slab_total;
        __syncthreads();
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 0) warp_max[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++) m = max(m, warp_max[w]);
        tile_base[T] = (uint32_t)carry;
        info[0] = (uint32_t)carry;
        info[1] = m;
        const bool ovf = carry > (unsigned long long)dup_capacity || info[10] != 0u;  // info[10]: a fixed-stride bin overflowed
        info[2] = ovf ? 1u : 0u;
        if (ovf) { info[3] += 1u; info[9] = (uint32_t)min(carry, 0xffffffffull); }  // sticky (deferred-check mode)
        info[10] = 0u;
        // V and D of this forward -> info[12..15] (what the host reads back), accumulators cleared for the next one
        const unsigned long long v = stats[0], d = stats[1];
        info[12] = (uint32_t)v; info[13] = (uint32_t)(v >> 32); info[14] = (uint32_t)d; info[15] = (uint32_t)(d >> 32);
        stats[0] = 0ull; stats[1] = 0ull;
    }
    if (tile_order) {
        // tile_base[] was written by this CTA above: visible to all its threads after the barrier
        for (int b = threadIdx.x; b <= ORDER_BUCKETS; b += SCAN_THREADS) s_hist[b] = 0u;
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += SCAN_THREADS) atomicAdd(&s_hist[order_bucket(tile_base[t + 1] - tile_base[t])], 1u);
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the ORDER_BUCKETS counts (4 per lane)
            uint32_t v[ORDER_BUCKETS / 32], sum = 0;
#pragma unroll
            for (int q = 0; q < ORDER_BUCKETS / 32; q++) { v[q] = s_hist[lane * (ORDER_BUCKETS / 32) + q]; sum += v[q]; }
            uint32_t inc = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t nn = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += nn;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int q = 0; q < ORDER_BUCKETS / 32; q++) { s_hist[lane * (ORDER_BUCKETS / 32) + q] = run; run += v[q]; }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += SCAN_THREADS)
            tile_order[atomicAdd(&s_hist[order_bucket(tile_base[t + 1] - tile_base[t])], 1u)] = (uint32_t)t;
    }
}

// The same scan spread over a thread-block CLUSTER of 8 CTAs (8 SMs).  The single-CTA form above is bound by what one SM's
// load/store path can do with sector-granular traffic — the counters are padded to one 32-byte sector each (so that the
// preprocess kernel's atomics do not collide), which makes T sector reads, T sector writes (zeroing) and, with eight
// consecutive tiles per thread, 32-sector tile_base / class-list stores per warp: ~20 us at c3 (6300 tiles), 2.3 % of the
// step and a serial link between the preprocess and sort kernels.  Here CTA r of the cluster owns tiles [r * 1024 * per,
// (r + 1) * 1024 * per), `per` consecutive tiles per thread (1 up to 8192 tiles: lanes <-> consecutive tiles, so the
// tile_base stores coalesce), publishes its total and maximum in its own shared memory, and after ONE cluster barrier
// reads the other seven through distributed shared memory to form its prefix: no global flags, no second launch.
constexpr int SCAN_CLUSTER = 8;
constexpr int SCAN_CLUSTER_MAX_PER = 4;  // up to 32768 tiles; beyond that the single-CTA kernel runs

__global__ void __cluster_dims__(SCAN_CLUSTER, 1, 1) __launch_bounds__(SCAN_THREADS)
tile_scan_cluster_kernel(int T, int per, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_base,
                         uint32_t* __restrict__ tile_cursor, uint32_t* __restrict__ info, uint32_t dup_capacity,
                         uint32_t* __restrict__ class_tiles, unsigned long long* __restrict__ stats) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t warp_max[32];
    __shared__ uint32_t s_pub[2];  // this CTA's total and maximum, read by the whole cluster
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (rank == 0 && threadIdx.x < NUM_SORT_CLASSES) info[4 + threadIdx.x] = 0u;  // (ordered before the atomics by the cluster barrier)
    const int first = ((int)rank * SCAN_THREADS + (int)threadIdx.x) * per;
    uint32_t c[SCAN_CLUSTER_MAX_PER];
    uint32_t sum = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < SCAN_CLUSTER_MAX_PER; k++) {
        const int t = first + k;
        c[k] = (k < per && t < T) ? tile_count[(size_t)t * TILE_CTR_STRIDE] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SCAN_CLUSTER_MAX_PER; k++) {
        if (c[k]) tile_count[(size_t)(first + k) * TILE_CTR_STRIDE] = 0u;  // dead after this read: zeroed for the next forward
        sum += c[k];
        mx = max(mx, c[k]);
    }
    uint32_t v = sum;  // inclusive warp scan of the per-thread sums
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v += n;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 31) warp_sums[warp] = v;
    if (lane == 0) warp_max[warp] = mx;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], m = warp_max[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += n;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
        warp_sums[lane] = w;  // inclusive over warps
        if (lane == 31) { s_pub[0] = w; s_pub[1] = m; }
    }
    cluster.sync();  // (also the CTA barrier for warp_sums)
    unsigned long long before = 0, total = 0;
    uint32_t gmax = 0;
#pragma unroll
    for (unsigned r = 0; r < SCAN_CLUSTER; r++) {
        const uint32_t* p = cluster.map_shared_rank(s_pub, r);
        const uint32_t t = p[0];
        if (r < rank) before += t;
        total += t;
        gmax = max(gmax, p[1]);
    }
    uint32_t run = (uint32_t)before + (warp ? warp_sums[warp - 1] : 0u) + v - sum;
    // per-class work lists for the sort kernels (info[4+q] = count of class q), as in the single-CTA kernel
    uint32_t mycnt[NUM_SORT_CLASSES];
#pragma unroll
    for (int q = 0; q < NUM_SORT_CLASSES; q++) mycnt[q] = 0;
#pragma unroll
    for (int k = 0; k < SCAN_CLUSTER_MAX_PER; k++)
        if (c[k]) {
            const int cls = sort_class_of(c[k]);
#pragma unroll
            for (int q = 0; q < NUM_SORT_CLASSES; q++) mycnt[q] += (cls == q) ? 1u : 0u;
        }
    uint32_t slot[NUM_SORT_CLASSES];
#pragma unroll
    for (int q = 0; q < NUM_SORT_CLASSES; q++) {
        uint32_t inc = mycnt[q];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += n;
        }
        const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t b = 0;
        if (lane == 0 && wtot) b = atomicAdd(info + 4 + q, wtot);
        slot[q] = __shfl_sync(0xffffffffu, b, 0) + inc - mycnt[q];
    }
#pragma unroll
    for (int k = 0; k < SCAN_CLUSTER_MAX_PER; k++) {
        const int t = first + k;
        if (k < per && t < T) {
            tile_base[t] = run;
            if (tile_cursor) tile_cursor[(size_t)t * TILE_CTR_STRIDE] = run;
            if (c[k]) {
                const int cls = sort_class_of(c[k]);
#pragma unroll
                for (int q = 0; q < NUM_SORT_CLASSES; q++)
                    if (cls == q) class_tiles[(size_t)q * T + slot[q]++] = (uint32_t)t;
            }
        }
        run += c[k];
    }
    if (rank == 0 && threadIdx.x == 0) {
        tile_base[T] = (uint32_t)total;
        info[0] = (uint32_t)total;
        info[1] = gmax;
        const bool ovf = total > (unsigned long long)dup_capacity || info[10] != 0u;  // info[10]: a fixed-stride bin overflowed
        info[2] = ovf ? 1u : 0u;
        if (ovf) { info[3] += 1u; info[9] = (uint32_t)min(total, 0xffffffffull); }  // sticky (deferred-check mode)
        info[10] = 0u;
        const unsigned long long vv = stats[0], d = stats[1];
        info[12] = (uint32_t)vv; info[13] = (uint32_t)(vv >> 32); info[14] = (uint32_t)d; info[15] = (uint32_t)(d >> 32);
        stats[0] = 0ull; stats[1] = 0ull;
    }
    cluster.sync();  // no CTA leaves while another may still read its s_pub
}

cudaError_t launch_tile_scan(int T, uint32_t* tile_count, uint32_t* tile_base, uint32_t* tile_cursor,
                             uint32_t* info, uint32_t dup_capacity, uint32_t* class_tiles, uint32_t* tile_order,
                             unsigned long long* stats, cudaStream_t st) {
    static const bool single = [] { const char* e = getenv("DVS_SCAN_SINGLE_CTA"); return e && atoi(e) != 0; }();  // A/B switch
    const int per = (T + SCAN_CLUSTER * SCAN_THREADS - 1) / (SCAN_CLUSTER * SCAN_THREADS);
    if (!single && !tile_order && per <= SCAN_CLUSTER_MAX_PER)
        tile_scan_cluster_kernel<<<SCAN_CLUSTER, SCAN_THREADS, 0, st>>>(T, per < 1 ? 1 : per, tile_count, tile_base, tile_cursor, info,
                                                                       dup_capacity, class_tiles, stats);
    else
        tile_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(T, tile_count, tile_base, tile_cursor, info, dup_capacity,
                                                     class_tiles, tile_order, stats);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// A3: emission (two-pass mode).  The warp-cooperative emission itself lives in emit.cuh and is shared with the
// fused single-pass mode of preprocess_fwd.cu.
// ---------------------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;

__global__ void __launch_bounds__(EMIT_THREADS)
emit_kernel(int gx, int N, const uint4* __restrict__ aux, const float4* __restrict__ rec, int stride,
            uint32_t* __restrict__ tile_cursor, unsigned long long* __restrict__ bins, uint32_t cap) {
    const int i = blockIdx.x * EMIT_THREADS + threadIdx.x;
    uint4 ax = make_uint4(0, 0, 0, 0);
    if (i < N) ax = __ldg(aux + i);
    const int minx = ax.x & 0xffff, miny = ax.x >> 16;
    const int w = (int)(ax.y & 0xffff) - minx, h = (int)((ax.y >> 16) & 0x1fff) - miny;
    const int area = (w > 0 && h > 0) ? w * h : 0;
    CullParams cp = {0.f, 0.f, 1.f, 0.f, 1.f, -1.f, 0.f, 0.f};
    if (area > 0) cp = cull_params(__ldg(rec + (size_t)stride * i), __ldg(rec + (size_t)stride * i + 1));
    warp_emit(gx, i, minx, miny, w, area, ax.z, cp, tile_cursor, bins, /*bin_stride=*/0u, cap, nullptr);
}

cudaError_t launch_emit(const Cam& cam, int N, const uint4* aux, const float4* cull, int cull_stride, uint32_t* tile_cursor,
                        unsigned long long* bins, uint32_t dup_capacity, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    emit_kernel<<<(N + EMIT_THREADS - 1) / EMIT_THREADS, EMIT_THREADS, 0, st>>>(cam.gx, N, aux, cull, cull_stride, tile_cursor,
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

// ---- adaptive two-level bucket sort (the common path) ---------------------------------------
// Depth keys are positive floats, so their bit patterns are monotone in depth and bucket =
// (bits - min) >> shift is a monotone, purely integer map.  Level 1: 256 coarse buckets over the tile's own
// [min,max]; level 2: each non-empty coarse bucket is subdivided over ITS OWN [min,max] into ~count
// sub-buckets (adapts to depth clusters: surfaces).  Entries are scattered to their fine bucket in
// arrival order and then ranked inside the bucket by the full 64-bit key (depth, id) — exact total
// order, O(1) expected work per entry.  If any fine bucket holds more than RANK_MAX entries (many
// equal / nearly equal depths) the tile falls back to the bitonic network above, which is always exact.
constexpr uint32_t RANK_MAX = 96;

__device__ __forceinline__ uint32_t next_pow2_u32(uint32_t c) { return c <= 1u ? 1u : 1u << (32 - __clz(c - 1u)); }

// In-place exclusive scan of a[0..len) by the whole CTA; returns the total (to every thread).
template <int THREADS>
__device__ uint32_t block_exclusive_scan(uint32_t* a, int len, uint32_t* s_part /* [THREADS/32 + 1] */) {
    constexpr int NW = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = ((len + NW - 1) / NW + 31) & ~31;  // contiguous region per warp, multiple of 32
    const int beg = warp * per, end = min(len, beg + per);
    uint32_t run = 0;
    for (int i = beg; i < end; i += 32) {
        const uint32_t c = (i + lane < end) ? a[i + lane] : 0u;
        uint32_t v = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += n;
        }
        if (i + lane < end) a[i + lane] = run + v - c;
        run += __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) s_part[warp] = run;
    __syncthreads();
    if (warp == 0) {
        const uint32_t c = lane < NW ? s_part[lane] : 0u;
        uint32_t v = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += n;
        }
        if (lane < NW) s_part[lane] = v - c;
        if (lane == 31) s_part[NW] = v;
    }
    __syncthreads();
    const uint32_t add = s_part[warp];
    if (add)
        for (int i = beg + lane; i < end; i += 32) a[i] += add;
    const uint32_t total = s_part[NW];
    __syncthreads();
    return total;
}

// KEYS_IN_SMEM = false (long lists): the unsorted keys stay in the global bin (L2-resident, re-read three
// times) so that two CTAs fit per SM.
template <int THREADS, int NMAX, bool KEYS_IN_SMEM>
__global__ void __launch_bounds__(THREADS)
tile_bucket_sort_kernel(int T, uint32_t bin_stride, int cls, const uint32_t* __restrict__ tile_base,
                        unsigned long long* __restrict__ bins_all, uint32_t* __restrict__ plist,
                        const uint32_t* __restrict__ info, const uint32_t* __restrict__ class_tiles) {
    extern __shared__ unsigned long long s_keys[];
    unsigned long long* tmp = s_keys + (KEYS_IN_SMEM ? NMAX : 0);  // [NMAX] keys grouped by fine bucket
    uint32_t* eb = reinterpret_cast<uint32_t*>(tmp + NMAX);  // [NMAX] fine bucket | arrival rank << 16
    uint32_t* cnt2 = eb + NMAX;                            // [2*NMAX + 32] fine-bucket counts -> bases
    uint32_t* cnt1 = cnt2 + 2 * NMAX + 32;                 // [256]
    uint32_t* min1 = cnt1 + 256;
    uint32_t* max1 = min1 + 256;
    uint32_t* sh2 = max1 + 256;
    uint32_t* off2 = sh2 + 256;                            // [257]
    uint16_t* eb2 = reinterpret_cast<uint16_t*>(off2 + 260);  // [NMAX] fine bucket of tmp[p]
    __shared__ uint32_t s_part[THREADS / 32 + 1];
    __shared__ uint32_t s_min[THREADS / 32], s_max[THREADS / 32];
    if (info[2]) return;
    const uint32_t ntiles = info[4 + cls];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const uint32_t tile = class_tiles[(size_t)cls * T + ti];
        const uint32_t b0 = tile_base[tile];
        const int n = (int)(tile_base[tile + 1] - b0);
        unsigned long long* bins = bins_all + (bin_stride ? (size_t)tile * bin_stride : (size_t)b0);  // this tile's bin
        unsigned long long* A = KEYS_IN_SMEM ? s_keys : bins;  // keys as emitted
        // phase 0: load, depth range
        uint32_t lmin = 0xffffffffu, lmax = 0u;
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const unsigned long long k = bins[i];
            if (KEYS_IN_SMEM) A[i] = k;
            const uint32_t d = (uint32_t)(k >> 32);
            lmin = min(lmin, d); lmax = max(lmax, d);
        }
        if (threadIdx.x < 256) { cnt1[threadIdx.x] = 0u; min1[threadIdx.x] = 0xffffffffu; max1[threadIdx.x] = 0u; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, off));
            lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
        }
        if (lane == 0) { s_min[warp] = lmin; s_max[warp] = lmax; }
        __syncthreads();
        uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) { kmin = min(kmin, s_min[w]); kmax = max(kmax, s_max[w]); }
        const uint32_t range = kmax - kmin;
        const uint32_t sh1 = range < 256u ? 0u : (uint32_t)(32 - __clz(range)) - 8u;
        // phase 1: coarse histogram + per-coarse-bucket range
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint32_t d = (uint32_t)(A[i] >> 32);
            const uint32_t b1 = (d - kmin) >> sh1;
            atomicAdd(cnt1 + b1, 1u);
            atomicMin(min1 + b1, d);
            atomicMax(max1 + b1, d);
        }
        __syncthreads();
        // phase 2: sub-bucket layout (threads 0..255 = coarse buckets), scan of sub-bucket counts
        if (threadIdx.x < 256) {
            const uint32_t c = cnt1[threadIdx.x];
            uint32_t nsub = 0u, sh = 0u;
            if (c) {
                const uint32_t r = max1[threadIdx.x] - min1[threadIdx.x];
                const uint32_t lg = 31u - (uint32_t)__clz(next_pow2_u32(c));
                const uint32_t bits = r ? (uint32_t)(32 - __clz(r)) : 0u;
                sh = bits > lg ? bits - lg : 0u;
                nsub = (r >> sh) + 1u;
            }
            sh2[threadIdx.x] = sh;
            off2[threadIdx.x] = nsub;
        }
        __syncthreads();
        const int NS = (int)block_exclusive_scan<THREADS>(off2, 256, s_part);  // off2[b1] = first fine bucket
        for (int i = threadIdx.x; i <= NS; i += THREADS) cnt2[i] = 0u;
        __syncthreads();
        // phase 3: fine histogram; remember (fine bucket, arrival rank)
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint32_t d = (uint32_t)(A[i] >> 32);
            const uint32_t b1 = (d - kmin) >> sh1;
            const uint32_t b2 = off2[b1] + ((d - min1[b1]) >> sh2[b1]);
            const uint32_t r = atomicAdd(cnt2 + b2, 1u);
            eb[i] = b2 | (r << 16);
        }
        __syncthreads();
        // phase 4: largest fine bucket (fallback decision) + exclusive scan -> fine bucket bases
        uint32_t mx = 0u;
        for (int i = threadIdx.x; i < NS; i += THREADS) mx = max(mx, cnt2[i]);
        const bool crowded = __syncthreads_or(mx > RANK_MAX);
        if (crowded) {
            bitonic_sort_u64<THREADS>(A, n);
            for (int i = threadIdx.x; i < n; i += THREADS) plist[b0 + i] = (uint32_t)A[i];
            __syncthreads();
            continue;
        }
        block_exclusive_scan<THREADS>(cnt2, NS + 1, s_part);  // cnt2[NS] = n afterwards
        // phase 5: group by fine bucket
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint32_t e = eb[i];
            const uint32_t b2 = e & 0xffffu;
            const uint32_t pos = cnt2[b2] + (e >> 16);
            tmp[pos] = A[i];
            eb2[pos] = (uint16_t)b2;
        }
        __syncthreads();
        // phase 6: rank inside the fine bucket by the full key, write the sorted list (low word = id<<8 | mask)
        for (int p = threadIdx.x; p < n; p += THREADS) {
            const unsigned long long k = tmp[p];
            const uint32_t b2 = eb2[p];
            const uint32_t s0 = cnt2[b2], s1 = cnt2[b2 + 1];
            uint32_t rank = 0;
            for (uint32_t q = s0; q < s1; q++) rank += tmp[q] < k ? 1u : 0u;
            plist[b0 + s0 + rank] = (uint32_t)k;
        }
        __syncthreads();
    }
}

// bitonic classes (long lists): shared memory up to 16384 entries, else in place in global/L2
template <int THREADS, bool IN_SMEM>
__global__ void __launch_bounds__(THREADS)
tile_bitonic_sort_kernel(int T, uint32_t bin_stride, int cls, const uint32_t* __restrict__ tile_base,
                         unsigned long long* __restrict__ bins, uint32_t* __restrict__ plist,
                         const uint32_t* __restrict__ info, const uint32_t* __restrict__ class_tiles) {
    extern __shared__ unsigned long long s_keys[];
    if (info[2]) return;
    const uint32_t ntiles = info[4 + cls];
    for (uint32_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const uint32_t tile = class_tiles[(size_t)cls * T + ti];
        const uint32_t b0 = tile_base[tile];
        const uint32_t n = tile_base[tile + 1] - b0;
        unsigned long long* g = bins + (bin_stride ? (size_t)tile * bin_stride : (size_t)b0);
        if (IN_SMEM) {
            for (uint32_t t = threadIdx.x; t < n; t += THREADS) s_keys[t] = g[t];
            __syncthreads();
            bitonic_sort_u64<THREADS>(s_keys, (int)n);
            for (uint32_t t = threadIdx.x; t < n; t += THREADS) plist[b0 + t] = (uint32_t)s_keys[t];
        } else {
            __syncthreads();
            bitonic_sort_u64<THREADS>(g, (int)n);  // rare: > 16384 entries in one tile
            for (uint32_t t = threadIdx.x; t < n; t += THREADS) plist[b0 + t] = (uint32_t)g[t];
        }
        __syncthreads();
    }
}

template <int NMAX, bool KEYS_IN_SMEM>
constexpr size_t bucket_smem_bytes() {
    return (size_t)NMAX * 8 * (KEYS_IN_SMEM ? 2 : 1) + (size_t)NMAX * 4 + (size_t)(2 * NMAX + 32) * 4 + 256 * 4 * 4 +
           260 * 4 + (size_t)NMAX * 2;
}

int tile_sort_launch_count(uint32_t bin_stride) {  // how many kernels launch_tile_sort enqueues (same conditions as below)
    const uint32_t max_len = bin_stride ? bin_stride : 0xffffffffu;
    return 2 + (max_len > 2048u ? 1 : 0) + (max_len > 4096u ? 1 : 0) + (max_len > 16384u ? 1 : 0);
}

cudaError_t launch_tile_sort(int T, uint32_t bin_stride, const uint32_t* tile_base, unsigned long long* bins,
                             uint32_t* plist, const uint32_t* info, const uint32_t* class_tiles, cudaStream_t st) {
    if (T <= 0) return cudaSuccess;
    // In single-pass mode no tile can hold more than bin_stride entries (a longer one overflows its bin and the
    // step is redone), so the kernels of the length classes above that bound are not launched at all.
    const uint32_t max_len = bin_stride ? bin_stride : 0xffffffffu;
    // (the dynamic shared-memory limit is a per-device, per-function attribute: once per device, not once per process)
    static bool attr_done_dev[64] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& attr_done = attr_done_dev[cur_dev & 63];
    if (!attr_done) {
        cudaFuncSetAttribute(tile_bucket_sort_kernel<256, 1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bucket_smem_bytes<1024, true>());
        cudaFuncSetAttribute(tile_bucket_sort_kernel<512, 2048, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bucket_smem_bytes<2048, true>());
        cudaFuncSetAttribute(tile_bucket_sort_kernel<1024, 4096, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bucket_smem_bytes<4096, false>());
        cudaFuncSetAttribute(tile_bitonic_sort_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             16384 * 8);
        attr_done = true;
    }
#ifndef DVS_SORT_GMULT
#define DVS_SORT_GMULT 1
#endif
    const int g_small = min(T, 148 * 6 * DVS_SORT_GMULT), g_mid = min(T, 148 * 3 * DVS_SORT_GMULT), g_long = min(T, 148 * 2), g_big = min(T, 148);
    tile_bucket_sort_kernel<256, 1024, true><<<g_small, 256, bucket_smem_bytes<1024, true>(), st>>>(T, bin_stride, 0, tile_base, bins, plist, info, class_tiles);
    tile_bucket_sort_kernel<512, 2048, true><<<g_mid, 512, bucket_smem_bytes<2048, true>(), st>>>(T, bin_stride, 1, tile_base, bins, plist, info, class_tiles);
    if (max_len > 2048u)
        tile_bucket_sort_kernel<1024, 4096, false><<<g_long, 1024, bucket_smem_bytes<4096, false>(), st>>>(T, bin_stride, 2, tile_base, bins, plist, info, class_tiles);
    if (max_len > 4096u)
        tile_bitonic_sort_kernel<1024, true><<<g_big, 1024, 16384 * 8, st>>>(T, bin_stride, 3, tile_base, bins, plist, info, class_tiles);
    if (max_len > 16384u)
        tile_bitonic_sort_kernel<1024, false><<<g_big, 1024, 0, st>>>(T, bin_stride, 4, tile_base, bins, plist, info, class_tiles);
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
    conic_opacity[4 * i + 1] = rad > 0 ? q1.x / (-LOG2E) : 0.f;
    conic_opacity[4 * i + 2] = rad > 0 ? q0.w / (-0.5f * LOG2E) : 0.f;
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
