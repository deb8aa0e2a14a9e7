// viewer_pack.cu — trainer -> viewer hand-off on the device (C-ABI in include/dvs_viewer_pack.h,
// per-Gaussian arithmetic in viewer_pack_ops.h).  One streaming kernel, HBM-bound: it reads the 236 B of raw parameters
// of every Gaussian once and writes 104 B of viewer records, 340 B/Gaussian of compulsory traffic.
//
// Layout of the work: a CTA of 128 threads owns 128 consecutive Gaussians per trip of a grid-stride loop.  The only wide
// row, shN (180 B per Gaussian), is a contiguous 23 KB span for the CTA: it is staged into shared memory with asynchronous
// 128-bit copies (every sector fully used, 12 in flight per thread) and each thread then reads its own row at stride 45
// words (odd: conflict-free).  The five narrow rows (12-16 B) are read directly and packed while those copies are in
// flight; a warp's loads of one array cover one contiguous 384-512 B span.  Records leave as 128-bit stores.  The bounding box is kept in registers across trips, reduced
// by shuffles, and leaves as six integer atomics per warp at the end (order-preserving float -> uint map).
// Grid = SM count x the CTAs that are resident at once (one wave).
// Compiled with -fmad=false (viewer_pack_ops.h is a literal operation sequence).
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdlib>

#include "dvs_viewer_pack.h"
#include "viewer_pack_ops.h"

namespace {
using namespace dvs_vp;
constexpr int kThreads = 128;
constexpr int kCtasPerSm = 6;  // default register budget: 6 CTAs per SM (85 registers; measured 0.0715 ms against 0.0755 ms with 8 x 64)
constexpr int kPrefetch = 1;   // DVS_VP_CTAS=5|6|8 and DVS_VP_PREFETCH=0|1 select the other instantiations (A/B)

// Stage the CTA's contiguous span of shN rows into shared memory with asynchronous 16-byte copies (LDGSTS): all of a thread's
// (up to) 12 copies are in flight at once and none passes through registers; the caller waits for them (cp.async.wait_group 0,
// then the CTA barrier) only after it has packed the narrow rows.  The plain loop of viewer_pack_ops.h (pack_stage,
// kept for the host harness: same indexing) left ONE 16-byte load per thread in flight — a tile's staging then cost 12 memory
// latencies back to back, two thirds of the kernel's time (round-2 ncu: long-scoreboard stalls 3.4 per issue, 2.8 TB/s).
__device__ __forceinline__ void stage_rows_async(const PackArgs& a, float* s_shn, int tid, long long base, int cnt) {
    const float* src = a.shN + base * kShRest;
    const int n_words = cnt * kShRest;
    if (a.shn_vec_ok) {
        const int n_vec = n_words >> 2;
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(s_shn);
#pragma unroll
        for (int it = 0; it < (kThreads * kShRest / 4 + kThreads - 1) / kThreads; it++) {
            const int i = tid + it * kThreads;
            if (i < n_vec) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16u * (unsigned)i), "l"(src + 4 * i) : "memory");
        }
        for (int i = (n_vec << 2) + tid; i < n_words; i += kThreads) s_shn[i] = src[i];
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        pack_stage(a, s_shn, tid, kThreads, base, cnt);
    }
}

// PREFETCH: the narrow rows of the CTA's NEXT tile are loaded into registers before the wide half of the current one runs,
// so their latency (the top stall of the plain form: 14 % of the samples) hides behind ~700 instructions of arithmetic.
template <int MIN_CTAS, bool PREFETCH>
__global__ void __launch_bounds__(kThreads, MIN_CTAS)
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
    NarrowRows rows{};
    if (PREFETCH && (int64_t)blockIdx.x * kThreads + tid < N) rows = load_narrow(a, (int64_t)blockIdx.x * kThreads + tid);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kThreads;
        const int cnt = (int)(N - base < kThreads ? N - base : kThreads);
        stage_rows_async(a, s_shn, tid, base, cnt);   // copies in flight ...
        if (PREFETCH) {                               // ... while the narrow rows are packed and stored
            if (tid < cnt) pack_narrow_rows(a, rows, base + tid, lo, hi);
            const int64_t nxt = (tile + gridDim.x) * kThreads + tid;
            if (nxt < N) rows = load_narrow(a, nxt);
        } else {
            pack_narrow(a, tid, base, cnt, lo, hi);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        pack_wide(a, s_shn, tid, base, cnt);
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
    // exactly one resident wave (ncu, round 2: 8 CTAs per SM were launched where 6 fit, and the 1/3 wave left over ran at a
    // third of the occupancy for as long as the first)
    using Kernel = void (*)(const float*, const float*, const float*, const float*, const float*, const float*, int64_t, uint4*,
                            uint2*, uint4*, uint32_t*, int);
    static Kernel kernel = nullptr;
    static int resident_dev[64] = {};  // (attributes and occupancy are per device)
    int& resident = resident_dev[dev & 63];
    if (!kernel || resident == 0) {
        int want = kCtasPerSm, prefetch = kPrefetch;
        if (const char* e = getenv("DVS_VP_CTAS")) want = atoi(e);
        if (const char* e = getenv("DVS_VP_PREFETCH")) prefetch = atoi(e);
        kernel = prefetch ? (want <= 5 ? viewer_pack_kernel<5, true> : want == 6 ? viewer_pack_kernel<6, true> : viewer_pack_kernel<8, true>)
                          : (want <= 5 ? viewer_pack_kernel<5, false> : want == 6 ? viewer_pack_kernel<6, false> : viewer_pack_kernel<8, false>);
        int occ = 0;
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0) != cudaSuccess || occ < 1) occ = 4;
        resident = occ;
    }
    const int64_t grid = n_tiles < (int64_t)sms * resident ? n_tiles : (int64_t)sms * resident;
    kernel<<<(unsigned)grid, kThreads, 0, st>>>(means, scales, quats, opacities, sh0, shN, N, static_cast<uint4*>(out_gaussians),
                                                static_cast<uint2*>(out_colors), static_cast<uint4*>(out_sh), bbox_ordered,
                                                misaligned(shN, 16) ? 0 : 1);
    return (int)cudaGetLastError();
}

DVS_VP_EXPORT void dvs_viewer_pack_decode_bbox(const uint32_t* b, float* min_xyz, float* max_xyz) {
    for (int a = 0; a < 3; a++) {
        min_xyz[a] = ordered_to_f32(b[a]);
        max_xyz[a] = ordered_to_f32(b[3 + a]);
    }
}
