// collective.cu — the one exchange step of the path: sum of the dense per-Gaussian gradient arena across the
// view-sharded ranks (SURVEY.md §8 e).  The reference has no multi-GPU code at all; NCCL's
// ncclAllReduce is the baseline.  This is the B200 / NVSwitch-native alternative: a two-shot all-reduce done by
// the switch itself on a multicast (NVLS) mapping of a symmetric buffer —
//     rank r owns shard r:  v = multimem.ld_reduce.add.v4.f32 [mc + i]   (switch sums the 8 replicas)
//                           multimem.st.v4.f32 [mc + i], v               (switch writes all 8 replicas)
// so every GPU moves ~(1 + 1/W) x the arena per direction instead of the ring's 2 (W-1)/W x.  The caller
// provides the multicast pointer of a symmetric allocation (torch.distributed._symmetric_memory) and
// brackets the call with cross-rank barriers (all ranks' gradients written before; all shards stored after).
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdlib.h>
#include <stdint.h>

#include "dvs_rast.h"
#include "sh_grad_ops.h"

namespace {

__device__ __forceinline__ float4 mm_ld_reduce(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mm_st(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

constexpr int AR_THREADS = 512;
constexpr int AR_UNROLL = 4;

__global__ void __launch_bounds__(AR_THREADS)
nvls_allreduce_kernel(float4* __restrict__ mc, size_t n_vec, int rank, int world) {
    // shard boundaries in units of float4, rank r owns [lo, hi)
    const size_t per = (n_vec + world - 1) / world;
    const size_t lo = per * rank, hi = lo + per < n_vec ? lo + per : n_vec;
    const size_t stride = (size_t)gridDim.x * AR_THREADS;
    size_t i = lo + (size_t)blockIdx.x * AR_THREADS + threadIdx.x;
    for (; i + (AR_UNROLL - 1) * stride < hi; i += AR_UNROLL * stride) {
        float4 v[AR_UNROLL];
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) v[u] = mm_ld_reduce(mc + i + u * stride);
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) mm_st(mc + i + u * stride, v[u]);
    }
    for (; i < hi; i += stride) mm_st(mc + i, mm_ld_reduce(mc + i));
}

}  // namespace

extern "C" DVS_API int dvs_coll_allreduce_nvls(void* multicast_ptr, size_t numel_f32, int rank, int world, int ctas,
                                              void* stream) {
    if (!multicast_ptr || (numel_f32 & 3) || world <= 0 || rank < 0 || rank >= world) return DVS_E_INVALID;
    if ((reinterpret_cast<uintptr_t>(multicast_ptr) & 15u) != 0) return DVS_E_INVALID;
    if (ctas <= 0) ctas = 148 * 2;
    nvls_allreduce_kernel<<<ctas, AR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float4*>(multicast_ptr), numel_f32 / 4, rank, world);
    return cudaGetLastError() == cudaSuccess ? DVS_OK : DVS_E_CUDA;
}


// ------------------------------------------------------------------------------------------------------------------
// The fused exchange (dvs_coll_exchange_fused, include/dvs_rast.h): in-switch reduction + peer reads + local SH accumulation
// in one persistent, co-resident grid with two device-side cross-rank barriers.
// ------------------------------------------------------------------------------------------------------------------
namespace {

static_assert(sizeof(dvs_coll_fused) == 488 && offsetof(dvs_coll_fused, sh0_tmp) == 144 && offsetof(dvs_coll_fused, N) == 384 &&
                  offsetof(dvs_coll_fused, rank) == 464,
              "dvs_coll_fused layout is part of the C-ABI (divshot_b200/_cabi.py: DvsCollFused)");
// CTA shapes of the fused kernel: 512 threads x 1 CTA per SM, or 256 threads x 2 CTAs per SM (two independent tile pipelines
// per SM: one CTA's NVLink round trip hides behind the other's arithmetic and stores).  Both are compiled; the 256 x 2 shape
// is the default, DVS_FX_THREADS=512 selects the other (A/B at 8 ranks in profiles/r2_scaling.md).
constexpr int FX_ROW_WORDS = 45;
constexpr unsigned long long FX_TIMEOUT_NS = 2000000000ull;

__device__ __forceinline__ void mm_red_release_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer memory: system-scope relaxed load — never served from this SM's L1, which is not coherent with another GPU's writes
__device__ __forceinline__ float ld_peer_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// wait until *p (monotonic counter) has reached `target`; wrap-safe comparison; false on timeout
template <bool SYS>
__device__ __forceinline__ bool wait_counter(const uint32_t* p, uint32_t target) {
    const unsigned long long t0 = now_ns();
    for (uint32_t spins = 0;; spins++) {
        const uint32_t v = SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p);
        if ((int32_t)(v - target) >= 0) return true;
        if ((spins & 1023u) == 1023u && now_ns() - t0 > FX_TIMEOUT_NS) return false;
    }
}

// Cross-rank barrier number `seq` (0, 1, 2, ... over the life of the buffers) of a co-resident grid of G CTAs on each of
// W ranks: every CTA arrives on the local counter; CTA 0 waits for its grid, then adds 1 to the symmetric counter in ALL
// replicas with one multimem.red.release; every CTA polls the local replica until all W ranks have done so.
__device__ __forceinline__ void cross_rank_barrier(const dvs_coll_fused& a, uint32_t seq, uint32_t* s_fail) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // this CTA's multimem / global stores and peer loads are ordered before its arrival
        atomicAdd(a.grid_counter, 1u);
        bool ok = true;
        if (blockIdx.x == 0) {
            ok = wait_counter<false>(a.grid_counter, (seq + 1u) * gridDim.x);
            __threadfence_system();
            mm_red_release_add_u32(a.signal_mc, 1u);
        }
        ok = wait_counter<true>(a.signal_local, (seq + 1u) * (uint32_t)a.world) && ok;
        if (!ok) {
            *s_fail = 1u;
            if (a.status) atomicExch(a.status, seq + 1u);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// step 3a for the FX_THREADS Gaussians [base, base + cnt): the CTA copies the dL/dsh0 rows of its tile from EVERY rank's arena
// into shared memory — the tile is one contiguous, 16-byte aligned span of 12 cnt bytes per rank, read as 128-bit peer loads
// (a row-wise 3 x 32-bit read would fetch every 32-byte sector three times over NVLink), up to eight ranks' loads in flight per
// thread before the first is consumed.  (A/B, 8 ranks, c3: the same copies as 16-byte cp.async / LDGSTS peer reads with two
// staging buffers — next tile in flight while the current one is processed — took 0.47 ms instead of 0.33 ms: asynchronous
// copies from PEER memory are slow on this platform; profiles/r2_d_bench_n8_cpasync_rejected.json.)
template <int FX_THREADS>
__device__ __forceinline__ void fused_tile_stage(const dvs_coll_fused& a, float* stage, int tid, long long base, int cnt) {
    const size_t row0 = (size_t)a.off_sh0 + 3 * (size_t)base;
    const int n_words = 3 * cnt, n_vec = n_words >> 2;
    for (int v0 = 0; v0 < a.world; v0 += 8) {
        float4 r[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (v0 + u < a.world && tid < n_vec) r[u] = ld_peer_f4(a.arena_peers[v0 + u] + row0 + 4 * (size_t)tid);
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (v0 + u < a.world && tid < n_vec) reinterpret_cast<float4*>(stage + (size_t)(v0 + u) * (3 * FX_THREADS))[tid] = r[u];
    }
    for (int w = (n_vec << 2) + tid; w < n_words; w += FX_THREADS)  // the last, partial tile only
        for (int v = 0; v < a.world; v++) stage[(size_t)v * (3 * FX_THREADS) + w] = ld_peer_f32(a.arena_peers[v] + row0 + w);
}
// step 3b: thread `tid` accumulates its Gaussian's dL/dshN row into the shared rows and the summed dL/dsh0 into sh0_tmp
template <int FX_THREADS>
__device__ __forceinline__ void fused_tile_compute(const dvs_coll_fused& a, const float* stage, float* rows, int tid, long long base,
                                                   int cnt) {
    if (tid >= cnt) return;
    const long long i = base + tid;
    float acc[45];
#pragma unroll
    for (int k = 0; k < 45; k++) acc[k] = 0.0f;
    const float mean[3] = {a.means[3 * i], a.means[3 * i + 1], a.means[3 * i + 2]};
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    for (int v = 0; v < a.world; v++) {
        const float* d = stage + (size_t)v * (3 * FX_THREADS) + 3 * tid;  // stride 3 words: conflict-free
        const float dc[3] = {d[0], d[1], d[2]};
        dvs_shx::accumulate_view(a.sh_degree, mean, a.campos + 3 * v, dc, acc);
        s0 += dc[0]; s1 += dc[1]; s2 += dc[2];
    }
    const int RW = 3 * a.sh_rest_alloc;
#pragma unroll
    for (int k = 0; k < 45; k++)
        if (k < RW) rows[tid * RW + k] = acc[k];
    a.sh0_tmp[3 * i] = s0; a.sh0_tmp[3 * i + 1] = s1; a.sh0_tmp[3 * i + 2] = s2;
}

template <int FX_THREADS>
__global__ void __launch_bounds__(FX_THREADS, 512 / FX_THREADS)
fused_exchange_kernel(const dvs_coll_fused a) {
    extern __shared__ __align__(16) float s_rows[];  // FX_THREADS rows of dL/dshN, then `world` x FX_THREADS staged dL/dsh0 rows
    float* s_stage = s_rows + FX_THREADS * FX_ROW_WORDS;
    __shared__ uint32_t s_fail;
    __shared__ long long s_tile;
    if (threadIdx.x == 0) s_fail = 0u;
    // the tile counter of step 3 (second local word): reset by CTA 0 before it arrives at the first barrier, which every
    // other CTA passes only after CTA 0 has arrived
    if (blockIdx.x == 0 && threadIdx.x == 0) a.grid_counter[1] = 0u;
    const uint32_t seq0 = (uint32_t)(a.launch_index * 2ull);

    // ---- 1. every rank's gradients are final
    cross_rank_barrier(a, seq0, &s_fail);
    if (s_fail) return;

    if ((int)blockIdx.x < a.reduce_ctas) {
        // ---- 2. in-switch sum of the reduced ranges: rank r owns shard r of each
        const size_t rtid = (size_t)blockIdx.x * FX_THREADS + threadIdx.x, rsize = (size_t)a.reduce_ctas * FX_THREADS;
        float4* mc = reinterpret_cast<float4*>(a.arena_mc);
#pragma unroll 1
        for (int r = 0; r < 3; r++) {
            const size_t v0 = (size_t)a.ranges[r][0] >> 2, v1 = (size_t)a.ranges[r][1] >> 2;
            if (v1 <= v0) continue;
            const size_t nv = v1 - v0, per = (nv + a.world - 1) / a.world;
            const size_t lo = v0 + per * a.rank, hi = (lo + per < v1) ? lo + per : v1;
            size_t i = lo + rtid;
            for (; i + 3 * rsize < hi; i += 4 * rsize) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = mm_ld_reduce(mc + i + u * rsize);
#pragma unroll
                for (int u = 0; u < 4; u++) mm_st(mc + i + u * rsize, v[u]);
            }
            for (; i < hi; i += rsize) mm_st(mc + i, mm_ld_reduce(mc + i));
        }
    }
    {
        // ---- 3. dL/dshN and the summed dL/dsh0 of all N Gaussians from every rank's dL/dsh0 (peer loads)
        float* out = a.arena_local + a.off_shN;
        const dvs_shx::ExchangeArgs x{a.means, a.campos, nullptr, (long long)a.N, a.world, a.sh_degree,
                                      3 * a.sh_rest_alloc, out, (reinterpret_cast<uintptr_t>(out) & 15u) == 0 ? 1 : 0};
        const long long n_tiles = (a.N + FX_THREADS - 1) / FX_THREADS;
        for (;;) {
            if (threadIdx.x == 0) s_tile = (long long)atomicAdd(a.grid_counter + 1, 1u);
            __syncthreads();
            const long long tile = s_tile;
            if (tile >= n_tiles) break;
            const long long base = tile * FX_THREADS;
            const int cnt = (int)(a.N - base < FX_THREADS ? a.N - base : FX_THREADS);
            fused_tile_stage<FX_THREADS>(a, s_stage, threadIdx.x, base, cnt);
            __syncthreads();
            fused_tile_compute<FX_THREADS>(a, s_stage, s_rows, threadIdx.x, base, cnt);
            __syncthreads();
            if (a.sh_rest_alloc > 0) dvs_shx::exchange_store(x, s_rows, threadIdx.x, FX_THREADS, base, cnt);
            __syncthreads();
        }
    }
    // ---- 4. every shard has been re-broadcast and nobody reads this rank's per-view dL/dsh0 any more
    cross_rank_barrier(a, seq0 + 1u, &s_fail);
    if (s_fail) return;
    {
        const size_t n = 3 * (size_t)a.N, gtid = (size_t)blockIdx.x * FX_THREADS + threadIdx.x, gsize = (size_t)gridDim.x * FX_THREADS;
        float* dst = a.arena_local + a.off_sh0;
        const size_t nv = n >> 2;  // off_sh0 is a multiple of 4 floats and sh0_tmp is 16-byte aligned
        for (size_t i = gtid; i < nv; i += gsize) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(a.sh0_tmp)[i];
        for (size_t i = (nv << 2) + gtid; i < n; i += gsize) dst[i] = a.sh0_tmp[i];
    }
}

int fused_threads() {
    static const int t = [] {
        const char* e = getenv("DVS_FX_THREADS");
        return (e && atoi(e) == 512) ? 512 : 256;
    }();
    return t;
}
size_t fused_smem(int world, int threads) { return (size_t)threads * (FX_ROW_WORDS + 3 * (size_t)(world > 0 ? world : 1)) * sizeof(float); }
const void* fused_kernel(int threads) {
    return threads == 512 ? reinterpret_cast<const void*>(fused_exchange_kernel<512>) : reinterpret_cast<const void*>(fused_exchange_kernel<256>);
}

// co-resident grid: `ctas` <= 0 -> every CTA slot of the device (SMs x CTAs per SM of the chosen shape)
int fused_grid(int ctas, int world) {
    int dev = 0, sms = 148, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = fused_threads();
    const size_t smem = fused_smem(world, threads);
    if (cudaFuncSetAttribute(fused_kernel(threads), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_kernel(threads), threads, smem) != cudaSuccess || per_sm < 1) return 0;
    const int cap = sms * per_sm;
    if (ctas <= 0) ctas = cap;
    return ctas < cap ? ctas : cap;
}

}  // namespace

extern "C" DVS_API int dvs_coll_exchange_fused_grid(int ctas, int world) { return fused_grid(ctas, world); }

extern "C" DVS_API int dvs_coll_exchange_fused(const dvs_coll_fused* args, void* stream) {
    if (!args) return DVS_E_INVALID;
    dvs_coll_fused a = *args;
    if (!a.arena_mc || !a.arena_local || !a.sh0_tmp || !a.signal_mc || !a.signal_local || !a.grid_counter || !a.means)
        return DVS_E_INVALID;
    if (a.world < 1 || a.world > 16 || a.rank < 0 || a.rank >= a.world || a.N < 0 || a.sh_degree < 0 || a.sh_degree > 3 ||
        a.sh_rest_alloc < 0 || a.sh_rest_alloc > 15 || (a.sh_degree + 1) * (a.sh_degree + 1) - 1 > a.sh_rest_alloc)
        return DVS_E_INVALID;
    for (int r = 0; r < a.world; r++)
        if (!a.arena_peers[r]) return DVS_E_INVALID;
    if (a.off_sh0 < 0 || (a.off_sh0 & 3) || a.off_shN < 0 || (a.off_shN & 3)) return DVS_E_INVALID;
    for (int r = 0; r < 3; r++)
        if (a.ranges[r][0] < 0 || (a.ranges[r][0] & 3) || (a.ranges[r][1] & 3) || a.ranges[r][1] < a.ranges[r][0]) return DVS_E_INVALID;
    const uintptr_t al = reinterpret_cast<uintptr_t>(a.arena_mc) | reinterpret_cast<uintptr_t>(a.arena_local) |
                         reinterpret_cast<uintptr_t>(a.sh0_tmp);
    if (al & 15u) return DVS_E_INVALID;
    if (a.N == 0) return DVS_OK;
    const int grid = fused_grid(a.ctas, a.world);
    if (grid < 1) return DVS_E_CUDA;
    a.ctas = grid;
    if (a.reduce_ctas <= 0) a.reduce_ctas = grid / 6;      // enough requests in flight for the switch (measured at 8 ranks: 24 of
                                                           // 148 CTAs 0.326 ms, 49: 0.335 ms); they join step 3 afterwards
    if (a.reduce_ctas > grid) a.reduce_ctas = grid;
    if (a.reduce_ctas < 1) a.reduce_ctas = 1;
    const int threads = fused_threads();
    const size_t smem = fused_smem(a.world, threads);
    void* kargs[] = {&a};
    // cooperative launch: the device-side barriers need every CTA resident (fails instead of deadlocking otherwise)
    const cudaError_t e = cudaLaunchCooperativeKernel(fused_kernel(threads), dim3(grid), dim3(threads), kargs, smem,
                                                      static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? DVS_OK : DVS_E_CUDA;
}
