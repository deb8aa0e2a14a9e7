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
#include <stdint.h>

#include "dvs_rast.h"

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
