// sh_exchange.cu — local half of the factored multi-GPU gradient exchange (SURVEY.md §8 e; arithmetic in sh_grad_ops.h).
// After every rank has all-gathered each view's dL/dsh0 (12 B per Gaussian and view) and camera centre, this kernel forms
// dL/dshN[i] = sum_v B(dir_{v,i}) (x) dL/dsh0_v[i] / SH_C0 for all N Gaussians, replacing the all-reduce of the 180 B per
// Gaussian dL/dshN tensor.  HBM-bound stream: reads 12 B (mean) + 12 V B, writes 3 KR * 4 B per Gaussian.
// One thread per Gaussian accumulates its row in registers; the CTA's 128 rows are contiguous in the output, so they are
// staged in shared memory (stride 3 KR words, odd for KR = 15: conflict-free) and leave as 128-bit stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dvs_rast.h"
#include "sh_grad_ops.h"

namespace {
constexpr int SX_THREADS = 128;
constexpr int SX_MAX_VIEWS = 64;
struct Campos { float p[SX_MAX_VIEWS * 3]; };

__global__ void __launch_bounds__(SX_THREADS)
sh_grad_from_dsh0_kernel(const float* __restrict__ means, Campos cams, const float* __restrict__ dsh0_all, int64_t N, int V, int deg,
                         int RW, float* __restrict__ out, int vec_ok) {
    __shared__ __align__(16) float s_rows[SX_THREADS * 45];
    const dvs_shx::ExchangeArgs a{means, cams.p, dsh0_all, (long long)N, V, deg, RW, out, vec_ok};
    const int tid = threadIdx.x;
    const int64_t n_tiles = (N + SX_THREADS - 1) / SX_THREADS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * SX_THREADS;
        const int cnt = (int)(N - base < SX_THREADS ? N - base : SX_THREADS);
        dvs_shx::exchange_compute(a, s_rows, tid, base, cnt);
        __syncthreads();
        dvs_shx::exchange_store(a, s_rows, tid, SX_THREADS, base, cnt);
        __syncthreads();  // the next trip overwrites s_rows
    }
}
}  // namespace

extern "C" DVS_API int dvs_coll_sh_grad_from_dsh0(const float* means, const float* campos_all_host, const float* dsh0_all, int64_t N,
                                                 int num_views, int sh_degree, int sh_rest_alloc, float* out_dshN, void* stream) {
    if (N < 0 || num_views < 1 || num_views > SX_MAX_VIEWS || sh_degree < 0 || sh_degree > 3 || sh_rest_alloc < 0 || sh_rest_alloc > 15)
        return DVS_E_INVALID;
    if ((sh_degree + 1) * (sh_degree + 1) - 1 > sh_rest_alloc) return DVS_E_INVALID;
    if (N == 0 || sh_rest_alloc == 0) return DVS_OK;
    if (!means || !campos_all_host || !dsh0_all || !out_dshN) return DVS_E_INVALID;
    Campos cams;
    for (int k = 0; k < 3 * num_views; k++) cams.p[k] = campos_all_host[k];
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (N + SX_THREADS - 1) / SX_THREADS;
    const int64_t grid = n_tiles < (int64_t)sms * 8 ? n_tiles : (int64_t)sms * 8;
    const int vec_ok = (reinterpret_cast<uintptr_t>(out_dshN) & 15u) == 0;
    sh_grad_from_dsh0_kernel<<<(unsigned)grid, SX_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        means, cams, dsh0_all, N, num_views, sh_degree, 3 * sh_rest_alloc, out_dshN, vec_ok);
    return cudaGetLastError() == cudaSuccess ? DVS_OK : DVS_E_CUDA;
}
