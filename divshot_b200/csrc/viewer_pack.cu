// viewer_pack.cu — trainer -> viewer hand-off on the device (C-ABI in include/dvs_viewer_pack.h,
// per-Gaussian arithmetic in viewer_pack_ops.h).  One streaming kernel, HBM-bound: it reads the 236 B of raw parameters
// of every Gaussian once and writes 104 B of viewer records, 340 B/Gaussian of compulsory traffic.
//
// Layout of the work: a CTA of 128 threads owns 128 consecutive Gaussians per trip of a grid-stride loop.  The only wide
// row, shN (180 B per Gaussian), is a contiguous 23 KB span for the CTA: it is staged into shared memory with 128-bit
// loads (every sector fully used) and each thread then reads its own row at stride 45 words (odd: conflict-free).  The
// five narrow rows (12-16 B) are read directly; a warp's loads of one array cover one contiguous 384-512 B span.
// Records leave as 128-bit stores.  The bounding box is kept in registers across trips, reduced by shuffles, and leaves
// as six integer atomics per warp at the end (order-preserving float -> uint map).  Grid = a multiple of the SM count.
// Compiled with -fmad=false (viewer_pack_ops.h is a literal operation sequence).
#include <cuda_runtime.h>

#include <cfloat>

#include "dvs_viewer_pack.h"
#include "viewer_pack_ops.h"

namespace {
using namespace dvs_vp;
constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
viewer_pack_kernel(const float* __restrict__ means, const float* __restrict__ scales, const float* __restrict__ quats,
                   const float* __restrict__ opac, const float* __restrict__ sh0, const float* __restrict__ shN, int64_t N,
                   uint4* __restrict__ out_g, uint2* __restrict__ out_c, uint4* __restrict__ out_sh,
                   uint32_t* __restrict__ bbox, int shn_vec_ok) {
    __shared__ __align__(16) float s_shn[kThreads * kShRest];
    const PackArgs a{means, scales, quats, opac, sh0, shN, (long long)N, reinterpret_cast<uint32_t*>(out_g),
                     reinterpret_cast<uint32_t*>(out_c), reinterpret_cast<uint32_t*>(out_sh), shn_vec_ok};
    const int tid = threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const int64_t n_tiles = (N + kThreads - 1) / kThreads;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kThreads;
        const int cnt = (int)(N - base < kThreads ? N - base : kThreads);
        pack_stage(a, s_shn, tid, kThreads, base, cnt);
        __syncthreads();
        pack_compute(a, s_shn, tid, base, cnt, lo, hi);
        __syncthreads();  // the next trip overwrites s_shn
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t l = f32_to_ordered(lo[k]), h = f32_to_ordered(hi[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l = min(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(bbox + k, l);
            atomicMax(bbox + 3 + k, h);
        }
    }
}

__global__ void bbox_init_kernel(uint32_t* bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = f32_to_ordered(FLT_MAX);
    else if (threadIdx.x < 6) bbox[threadIdx.x] = f32_to_ordered(-FLT_MAX);
}
}  // namespace

#define DVS_VP_EXPORT extern "C" __attribute__((visibility("default")))

DVS_VP_EXPORT int dvs_viewer_pack(const float* means, const float* scales, const float* quats, const float* opacities,
                                  const float* sh0, const float* shN, int64_t N, void* out_gaussians, void* out_colors,
                                  void* out_sh, uint32_t* bbox_ordered, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || !bbox_ordered) return (int)cudaErrorInvalidValue;
    bbox_init_kernel<<<1, 32, 0, st>>>(bbox_ordered);
    if (N == 0) return (int)cudaGetLastError();
    if (!means || !scales || !quats || !opacities || !sh0 || !shN || !out_gaussians || !out_colors || !out_sh)
        return (int)cudaErrorInvalidValue;
    auto misaligned = [](const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; };
    if (misaligned(out_gaussians, 16) || misaligned(out_sh, 16) || misaligned(out_colors, 8) || misaligned(quats, 16))
        return (int)cudaErrorMisalignedAddress;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (N + kThreads - 1) / kThreads;
    const int64_t grid = n_tiles < (int64_t)sms * 8 ? n_tiles : (int64_t)sms * 8;  // 8 CTAs of 128 threads per SM
    viewer_pack_kernel<<<(unsigned)grid, kThreads, 0, st>>>(means, scales, quats, opacities, sh0, shN, N,
                                                           static_cast<uint4*>(out_gaussians), static_cast<uint2*>(out_colors),
                                                           static_cast<uint4*>(out_sh), bbox_ordered, misaligned(shN, 16) ? 0 : 1);
    return (int)cudaGetLastError();
}

DVS_VP_EXPORT void dvs_viewer_pack_decode_bbox(const uint32_t* b, float* min_xyz, float* max_xyz) {
    for (int a = 0; a < 3; a++) {
        min_xyz[a] = ordered_to_f32(b[a]);
        max_xyz[a] = ordered_to_f32(b[3 + a]);
    }
}
