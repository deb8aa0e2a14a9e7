// densify.cu — the trainer's refinement step on the device (interface in densify.h, per-element
// maths in densify_ops.h).  Streaming kernels over N Gaussians (one thread per Gaussian, grid-stride), CUB only for the
// two data-parallel primitives that are not on the rasterizer path (an inclusive scan for the sampling CDF, a stream
// compaction for index lists).  Nothing here runs before iteration `warmupLength`, and only every `refineEvery`
// iterations, except mcmc_noise / mcmc_regularise / adc_accumulate which are O(N) streaming passes per step.
//
// The same source also compiles for the HOST (tests/native/densify_emul.cpp defines DVS_DENSIFY_HOST_EMULATION): kernels
// become serial loops, the two CUB primitives plain loops, the CUDA runtime a malloc shim.  That build is test
// infrastructure — it lets the CPU suite run the exact kernel bodies and the exact host orchestration (hole filling,
// capacity rules, sampling) against the checkers of tests/densify_ref.py.  The product (libgstrain.so) never has it.
#include "densify.h"

#ifndef DVS_DENSIFY_HOST_EMULATION
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#endif

#include <algorithm>
#include <vector>

#include "densify_ops.h"

namespace dvs_densify {

namespace {
constexpr int kThreads = 256;
inline int blocks_for(int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>((n + kThreads - 1) / kThreads, 1), 148 * 16); }
#ifdef DVS_DENSIFY_HOST_EMULATION
#define GRID_STRIDE(i, n) for (int64_t i = 0; i < (n); i++)
#define DVS_LAUNCH(kernel, n, st, ...) ((void)(st), (void)blocks_for(n), kernel(__VA_ARGS__))
#else
#define GRID_STRIDE(i, n) for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)
#define DVS_LAUNCH(kernel, n, st, ...) kernel<<<blocks_for(n), kThreads, 0, st>>>(__VA_ARGS__)
#endif
#define CKC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return e__; } while (0)

__device__ __forceinline__ void copy_row(const Tensors& t, int64_t d, int64_t s) {
    for (int k = 0; k < 3; k++) { t.means[3 * d + k] = t.means[3 * s + k]; t.scales[3 * d + k] = t.scales[3 * s + k]; t.sh0[3 * d + k] = t.sh0[3 * s + k]; }
    for (int k = 0; k < 4; k++) t.quats[4 * d + k] = t.quats[4 * s + k];
    t.opac[d] = t.opac[s];
    for (int k = 0; k < 45; k++) t.shN[45 * d + k] = t.shN[45 * s + k];
}
__device__ __forceinline__ void zero_row(const Tensors& t, int64_t d) {
    for (int k = 0; k < 3; k++) { t.means[3 * d + k] = 0.f; t.scales[3 * d + k] = 0.f; t.sh0[3 * d + k] = 0.f; }
    for (int k = 0; k < 4; k++) t.quats[4 * d + k] = 0.f;
    t.opac[d] = 0.f;
    for (int k = 0; k < 45; k++) t.shN[45 * d + k] = 0.f;
}

// ---------------------------------------------------------------------------------------------- MCMC kernels
__global__ void k_weights(const float* __restrict__ opac, int64_t N, float min_opacity, bool dead_get_zero,
                          double* __restrict__ weight, uint8_t* __restrict__ dead) {
    GRID_STRIDE(i, N) {
        const float o = sigmoidf_(opac[i]);
        const bool d = o <= min_opacity;
        if (dead) dead[i] = d;
        weight[i] = (d && dead_get_zero) ? 0.0 : (double)o;
    }
}
// The sample count is `M`, or — when `M_dev` is given — the count a preceding compaction left on the DEVICE (no host
// read-back between the kernels): *M_dev if 0 < *M_dev < N, else 0 (nothing dead, or everything dead: nothing to relocate).
__device__ __forceinline__ int64_t effective_count(int64_t M, const int64_t* M_dev, int64_t N) {
    if (!M_dev) return M;
    const int64_t d = *M_dev;
    return (d > 0 && d < N) ? d : 0;
}
__global__ void k_sample(const double* __restrict__ cdf, int64_t N, int64_t M, const int64_t* __restrict__ M_dev, uint64_t seed,
                         int64_t* __restrict__ src, int32_t* __restrict__ count) {
    const double total = cdf[N - 1];
    M = effective_count(M, M_dev, N);
    GRID_STRIDE(j, M) {
        int64_t idx = sample_cdf(cdf, N, uniform01(seed, (uint64_t)j) * total);
        while (idx > 0 && cdf[idx] == cdf[idx - 1]) idx--;  // never land on a zero-weight entry (u rounded up to the total)
        src[j] = idx;
        atomicAdd(count + idx, 1);
    }
}
__global__ void k_reloc_values(Tensors p, const int32_t* __restrict__ count, int64_t N, float min_opacity,
                               const float* __restrict__ binom, float* __restrict__ tmp_opac, float* __restrict__ tmp_scale) {
    GRID_STRIDE(i, N) {
        const int c = count[i];
        if (c <= 0) continue;
        const float s[3] = {expf(p.scales[3 * i]), expf(p.scales[3 * i + 1]), expf(p.scales[3 * i + 2])};
        float nl, ns[3];
        relocation(binom, sigmoidf_(p.opac[i]), s, c + 1, min_opacity, nl, ns);
        tmp_opac[i] = nl;
        for (int a = 0; a < 3; a++) tmp_scale[3 * i + a] = ns[a];
    }
}
// copy j: Gaussian src[j] -> slot (dst_list ? dst_list[j] : dst_base + j), with the relocation rule's opacity / scale
__global__ void k_copy_relocated(Tensors p, Tensors m1, Tensors m2, const int64_t* __restrict__ src,
                                 const int64_t* __restrict__ dst_list, int64_t dst_base, int64_t M, const int64_t* __restrict__ M_dev,
                                 int64_t N, const float* __restrict__ tmp_opac, const float* __restrict__ tmp_scale) {
    M = effective_count(M, M_dev, N);
    GRID_STRIDE(j, M) {
        const int64_t s = src[j], d = dst_list ? dst_list[j] : dst_base + j;
        copy_row(p, d, s);
        p.opac[d] = tmp_opac[s];
        for (int a = 0; a < 3; a++) p.scales[3 * d + a] = tmp_scale[3 * s + a];
        zero_row(m1, d);
        zero_row(m2, d);
    }
}
__global__ void k_commit_sources(Tensors p, Tensors m1, Tensors m2, const int32_t* __restrict__ count, int64_t N,
                                 const float* __restrict__ tmp_opac, const float* __restrict__ tmp_scale) {
    GRID_STRIDE(i, N) {
        if (count[i] <= 0) continue;
        p.opac[i] = tmp_opac[i];
        for (int a = 0; a < 3; a++) p.scales[3 * i + a] = tmp_scale[3 * i + a];
        zero_row(m1, i);
        zero_row(m2, i);
    }
}
__global__ void k_noise(Tensors p, int64_t N, float step, uint64_t seed) {
    GRID_STRIDE(i, N) {
        float e[4];
        normal2(seed, 2 * (uint64_t)i, e[0], e[1]);
        normal2(seed, 2 * (uint64_t)i + 1, e[2], e[3]);
        float d[3];
        mcmc_noise(p.scales + 3 * i, p.quats + 4 * i, p.opac[i], e, step, d);
        for (int a = 0; a < 3; a++) p.means[3 * i + a] += d[a];
    }
}
__global__ void k_regularise(Tensors p, Tensors g, int64_t N, float w_o, float w_s) {
    const float inv_n = 1.0f / (float)N, inv_3n = inv_n / 3.0f;
    GRID_STRIDE(i, N) {
        g.opac[i] += reg_grad_opacity(p.opac[i], w_o, inv_n);
        for (int a = 0; a < 3; a++) g.scales[3 * i + a] += reg_grad_scale(p.scales[3 * i + a], w_s, inv_3n);
    }
}

// ---------------------------------------------------------------------------------------------- ADC kernels
__global__ void k_adc_accumulate(const float* __restrict__ g2, const float* __restrict__ gabs, const int32_t* __restrict__ radii,
                                 float* __restrict__ accum, float* __restrict__ denom, int64_t N, const uint32_t* __restrict__ skip) {
    if (skip && *skip) return;  // the step's forward overflowed its arena: no gradients were produced, nothing to count
    GRID_STRIDE(i, N) {
        if (radii[i] <= 0) continue;
        const float* g = gabs ? gabs : g2;
        accum[i] += sqrtf(g[2 * i] * g[2 * i] + g[2 * i + 1] * g[2 * i + 1]);
        denom[i] += 1.0f;
    }
}
__global__ void k_adc_decide(Tensors p, const float* __restrict__ accum, const float* __restrict__ denom, int64_t N,
                             AdcConfig c, uint8_t* __restrict__ act, uint8_t* __restrict__ grow, uint8_t* __restrict__ pruned) {
    GRID_STRIDE(i, N) {
        const uint8_t a = adc_decide(accum[i], denom[i], p.scales + 3 * i, p.opac[i], c.grad_threshold, c.percent_dense,
                                     c.extent, c.prune_opacity, c.prune_scale3d);
        act[i] = a;
        pruned[i] = (a & ADC_PRUNE) ? 1 : 0;
        grow[i] = (!(a & ADC_PRUNE) && (a & (ADC_CLONE | ADC_SPLIT))) ? 1 : 0;
    }
}
// appended element j: a copy of grow_list[j] (clone) or its second split sample; Adam moments of the new slot zeroed
__global__ void k_adc_append(Tensors p, Tensors m1, Tensors m2, const int64_t* __restrict__ grow_list,
                             const uint8_t* __restrict__ act, int64_t base, int64_t M, uint64_t seed, bool revised) {
    GRID_STRIDE(j, M) {
        const int64_t s = grow_list[j], d = base + j;
        copy_row(p, d, s);
        if (revised) p.opac[d] = revised_opacity_logit(p.opac[s]);
        if (act[s] & ADC_SPLIT) {
            float e[4];
            normal2(seed, 4 * (uint64_t)s + 2, e[0], e[1]);
            normal2(seed, 4 * (uint64_t)s + 3, e[2], e[3]);
            float m[3], ls[3];
            adc_split_sample(p.means + 3 * s, p.scales + 3 * s, p.quats + 4 * s, e, m, ls);
            for (int a = 0; a < 3; a++) { p.means[3 * d + a] = m[a]; p.scales[3 * d + a] = ls[a]; }
        }
        zero_row(m1, d);
        zero_row(m2, d);
    }
}
// split originals become their own first sample (runs after k_adc_append, which reads the original mean / scale / opacity);
// with `revised` every grown original (clone or split) takes the revised opacity its copy already has
__global__ void k_adc_split_in_place(Tensors p, Tensors m1, Tensors m2, const uint8_t* __restrict__ act,
                                     const uint8_t* __restrict__ grow, int64_t N, uint64_t seed, bool revised) {
    GRID_STRIDE(i, N) {
        if (!grow[i]) continue;
        if (revised) p.opac[i] = revised_opacity_logit(p.opac[i]);
        if (!(act[i] & ADC_SPLIT)) continue;
        float e[4];
        normal2(seed, 4 * (uint64_t)i, e[0], e[1]);
        normal2(seed, 4 * (uint64_t)i + 1, e[2], e[3]);
        float m[3], ls[3];
        adc_split_sample(p.means + 3 * i, p.scales + 3 * i, p.quats + 4 * i, e, m, ls);
        for (int a = 0; a < 3; a++) { p.means[3 * i + a] = m[a]; p.scales[3 * i + a] = ls[a]; }
        zero_row(m1, i);
        zero_row(m2, i);
    }
}
// hole filling: flags over [0, total): hole = i < K && pruned, mover = i >= K && !pruned (pruned[i] = 0 for i >= N)
__global__ void k_hole_flags(const uint8_t* __restrict__ pruned, int64_t N, int64_t total, int64_t K,
                             uint8_t* __restrict__ hole, uint8_t* __restrict__ mover) {
    GRID_STRIDE(i, total) {
        const bool pr = i < N && pruned[i];
        hole[i] = (i < K && pr) ? 1 : 0;
        mover[i] = (i >= K && !pr) ? 1 : 0;
    }
}
__global__ void k_move_rows(Tensors p, Tensors m1, Tensors m2, float* __restrict__ accum, float* __restrict__ denom,
                            const int64_t* __restrict__ holes, const int64_t* __restrict__ movers, int64_t M) {
    GRID_STRIDE(j, M) {
        const int64_t d = holes[j], s = movers[j];
        copy_row(p, d, s);
        copy_row(m1, d, s);
        copy_row(m2, d, s);
        accum[d] = accum[s];
        denom[d] = denom[s];
    }
}
__global__ void k_reset_opacity(Tensors p, Tensors m1, Tensors m2, int64_t N, float cap_logit) {
    GRID_STRIDE(i, N) {
        p.opac[i] = fminf(p.opac[i], cap_logit);
        m1.opac[i] = 0.f;
        m2.opac[i] = 0.f;
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------- workspace
struct Workspace {
    int64_t cap = 0;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    double *weight = nullptr, *cdf = nullptr;
    uint8_t *flag_a = nullptr, *flag_b = nullptr, *act = nullptr;
    int64_t *list_a = nullptr, *list_b = nullptr;
    int32_t* count = nullptr;
    float *tmp_opac = nullptr, *tmp_scale = nullptr;
    int64_t* d_num = nullptr;  // [2]
    int64_t* h_num = nullptr;  // pinned [2]
    float* binom = nullptr;    // [kMaxRatio^2]
};

Workspace* workspace_create() { return new Workspace(); }

static void ws_free(Workspace* ws) {
    cudaFree(ws->cub_tmp); cudaFree(ws->weight); cudaFree(ws->cdf); cudaFree(ws->flag_a); cudaFree(ws->flag_b);
    cudaFree(ws->act); cudaFree(ws->list_a); cudaFree(ws->list_b); cudaFree(ws->count); cudaFree(ws->tmp_opac);
    cudaFree(ws->tmp_scale);
    ws->cub_tmp = nullptr; ws->cub_bytes = 0;
    ws->weight = ws->cdf = nullptr; ws->flag_a = ws->flag_b = ws->act = nullptr; ws->list_a = ws->list_b = nullptr;
    ws->count = nullptr; ws->tmp_opac = ws->tmp_scale = nullptr;
    ws->cap = 0;
}
void workspace_destroy(Workspace* ws) {
    if (!ws) return;
    ws_free(ws);
    cudaFree(ws->d_num); cudaFree(ws->binom);
    if (ws->h_num) cudaFreeHost(ws->h_num);
    delete ws;
}
static cudaError_t ws_ensure(Workspace* ws, int64_t cap) {
    if (!ws->d_num) {
        CKC(cudaMalloc(&ws->d_num, 2 * sizeof(int64_t)));
        CKC(cudaMallocHost(&ws->h_num, 2 * sizeof(int64_t)));
        std::vector<float> b((size_t)kMaxRatio * kMaxRatio, 0.f);  // Pascal's triangle, C(n,k) at [n*kMaxRatio + k]
        for (int n = 0; n < kMaxRatio; n++) {
            b[(size_t)n * kMaxRatio] = 1.f;
            for (int k = 1; k <= n; k++)
                b[(size_t)n * kMaxRatio + k] = b[(size_t)(n - 1) * kMaxRatio + k - 1] + (k <= n - 1 ? b[(size_t)(n - 1) * kMaxRatio + k] : 0.f);
        }
        CKC(cudaMalloc(&ws->binom, b.size() * sizeof(float)));
        CKC(cudaMemcpy(ws->binom, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (cap <= ws->cap) return cudaSuccess;
    ws_free(ws);
    const size_t n = (size_t)cap;
    CKC(cudaMalloc(&ws->weight, n * sizeof(double)));
    CKC(cudaMalloc(&ws->cdf, n * sizeof(double)));
    CKC(cudaMalloc(&ws->flag_a, n));
    CKC(cudaMalloc(&ws->flag_b, n));
    CKC(cudaMalloc(&ws->act, n));
    CKC(cudaMalloc(&ws->list_a, n * sizeof(int64_t)));
    CKC(cudaMalloc(&ws->list_b, n * sizeof(int64_t)));
    CKC(cudaMalloc(&ws->count, n * sizeof(int32_t)));
    CKC(cudaMalloc(&ws->tmp_opac, n * sizeof(float)));
    CKC(cudaMalloc(&ws->tmp_scale, 3 * n * sizeof(float)));
    ws->cap = cap;
    return cudaSuccess;
}
static cudaError_t ws_cub(Workspace* ws, size_t bytes) {
    if (bytes <= ws->cub_bytes) return cudaSuccess;
    cudaFree(ws->cub_tmp);
    ws->cub_tmp = nullptr;
    ws->cub_bytes = 0;
    CKC(cudaMalloc(&ws->cub_tmp, bytes));
    ws->cub_bytes = bytes;
    return cudaSuccess;
}
// indices i in [0, n) with flag[i] != 0 -> out (ascending); count -> d_num[slot]
#ifdef DVS_DENSIFY_HOST_EMULATION
static cudaError_t select_indices(Workspace* ws, const uint8_t* flag, int64_t n, int64_t* out, int slot, cudaStream_t) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++)
        if (flag[i]) out[m++] = i;
    ws->d_num[slot] = m;
    return cudaSuccess;
}
static cudaError_t inclusive_sum(Workspace*, const double* in, double* out, int64_t n, cudaStream_t) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) out[i] = (acc += in[i]);
    return cudaSuccess;
}
#else
static cudaError_t select_indices(Workspace* ws, const uint8_t* flag, int64_t n, int64_t* out, int slot, cudaStream_t st) {
    size_t bytes = 0;
    thrust::counting_iterator<int64_t> iota(0);
    CKC(cub::DeviceSelect::Flagged(nullptr, bytes, iota, flag, out, ws->d_num + slot, n, st));
    CKC(ws_cub(ws, bytes));
    return cub::DeviceSelect::Flagged(ws->cub_tmp, bytes, iota, flag, out, ws->d_num + slot, n, st);
}
static cudaError_t inclusive_sum(Workspace* ws, const double* in, double* out, int64_t n, cudaStream_t st) {
    size_t bytes = 0;
    CKC(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, st));
    CKC(ws_cub(ws, bytes));
    return cub::DeviceScan::InclusiveSum(ws->cub_tmp, bytes, in, out, n, st);
}
#endif
static cudaError_t read_counts(Workspace* ws, int n, cudaStream_t st) {
    CKC(cudaMemcpyAsync(ws->h_num, ws->d_num, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    return cudaStreamSynchronize(st);
}

// sample M sources from [0, N) with probability ~ weight, apply the relocation rule to sources and copies
// (M_dev != nullptr: the number of samples is on the device, M is only its upper bound — launch geometry)
static cudaError_t sample_and_copy(Workspace* ws, Tensors p, Tensors m1, Tensors m2, int64_t N, int64_t M, const int64_t* M_dev,
                                   const int64_t* dst_list, int64_t dst_base, float min_opacity, uint64_t seed,
                                   cudaStream_t st) {
    CKC(inclusive_sum(ws, ws->weight, ws->cdf, N, st));
    CKC(cudaMemsetAsync(ws->count, 0, (size_t)N * sizeof(int32_t), st));
    DVS_LAUNCH(k_sample, M, st, ws->cdf, N, M, M_dev, seed, ws->list_b, ws->count);
    DVS_LAUNCH(k_reloc_values, N, st, p, ws->count, N, min_opacity, ws->binom, ws->tmp_opac, ws->tmp_scale);
    DVS_LAUNCH(k_copy_relocated, M, st, p, m1, m2, ws->list_b, dst_list, dst_base, M, M_dev, N, ws->tmp_opac, ws->tmp_scale);
    DVS_LAUNCH(k_commit_sources, N, st, p, m1, m2, ws->count, N, ws->tmp_opac, ws->tmp_scale);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- MCMC
cudaError_t mcmc_refine(Workspace* ws, Tensors p, Tensors m1, Tensors m2, int64_t* N_io, int64_t capacity, int64_t cap_max,
                        float min_opacity, uint64_t seed, cudaStream_t st, RefineReport* rep) {
    const int64_t N = *N_io;
    if (N <= 0) return cudaSuccess;
    CKC(ws_ensure(ws, std::max(capacity, N)));
    // 1. relocate the dead onto the living
    DVS_LAUNCH(k_weights, N, st, p.opac, N, min_opacity, true, ws->weight, ws->flag_a);
    CKC(select_indices(ws, ws->flag_a, N, ws->list_a, 0, st));
    // the number of dead Gaussians stays on the device: the sampling / copy kernels read it there (launched for the upper
    // bound N), so the refinement queues without a host synchronisation.  Only a caller that wants the report pays for one.
    CKC(sample_and_copy(ws, p, m1, m2, N, N, ws->d_num, ws->list_a, 0, min_opacity, seed, st));
    if (rep) {
        CKC(read_counts(ws, 1, st));
        rep->dead = ws->h_num[0];
        rep->relocated = (rep->dead > 0 && rep->dead < N) ? rep->dead : 0;
    }
    // 2. grow by 5 %, bounded by capMax and by what the arenas hold
    const int64_t target = std::min<int64_t>(std::min<int64_t>(cap_max, capacity), (int64_t)(1.05 * (double)N));
    const int64_t n_new = std::max<int64_t>(0, target - N);
    if (n_new > 0) {
        DVS_LAUNCH(k_weights, N, st, p.opac, N, min_opacity, false, ws->weight, nullptr);
        CKC(sample_and_copy(ws, p, m1, m2, N, n_new, nullptr, nullptr, N, min_opacity, seed ^ 0xA5A5A5A5A5A5A5A5ull, st));
        *N_io = N + n_new;
        if (rep) rep->added = n_new;
    }
    return cudaGetLastError();
}

cudaError_t mcmc_noise(Tensors p, int64_t N, float step, uint64_t seed, cudaStream_t st) {
    if (N <= 0 || step == 0.f) return cudaSuccess;
    DVS_LAUNCH(k_noise, N, st, p, N, step, seed);
    return cudaGetLastError();
}

cudaError_t mcmc_regularise(Tensors p, Tensors g, int64_t N, float w_o, float w_s, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    DVS_LAUNCH(k_regularise, N, st, p, g, N, w_o, w_s);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- ADC
cudaError_t adc_accumulate(const float* mean2D_grad, const float* mean2D_abs, const int32_t* radii, float* accum,
                           float* denom, int64_t N, cudaStream_t st, const uint32_t* skip) {
    if (N <= 0) return cudaSuccess;
    DVS_LAUNCH(k_adc_accumulate, N, st, mean2D_grad, mean2D_abs, radii, accum, denom, N, skip);
    return cudaGetLastError();
}

cudaError_t adc_refine(Workspace* ws, Tensors p, Tensors m1, Tensors m2, float* accum, float* denom, int64_t* N_io,
                       int64_t capacity, int64_t cap_max, const AdcConfig& cfg, uint64_t seed, cudaStream_t st,
                       RefineReport* rep) {
    const int64_t N = *N_io;
    if (N <= 0) return cudaSuccess;
    CKC(ws_ensure(ws, std::max(capacity, N)));
    DVS_LAUNCH(k_adc_decide, N, st, p, accum, denom, N, cfg, ws->act, ws->flag_a, ws->flag_b);
    CKC(select_indices(ws, ws->flag_a, N, ws->list_a, 0, st));  // Gaussians that clone or split
    CKC(select_indices(ws, ws->flag_b, N, ws->list_b, 1, st));  // pruned (only the count is used here)
    CKC(read_counts(ws, 2, st));
    int64_t n_grow = ws->h_num[0];
    const int64_t n_pruned = ws->h_num[1];
    const int64_t limit = std::min(cap_max, capacity);
    if (N - n_pruned + n_grow > limit || N + n_grow > capacity) n_grow = 0;  // no room: prune only this round
    if (n_grow > 0) {
        DVS_LAUNCH(k_adc_append, n_grow, st, p, m1, m2, ws->list_a, ws->act, N, n_grow, seed, cfg.revised_opacity);
        DVS_LAUNCH(k_adc_split_in_place, N, st, p, m1, m2, ws->act, ws->flag_a, N, seed, cfg.revised_opacity);
    }
    const int64_t total = N + n_grow, K = total - n_pruned;
    // statistics restart after every refinement (also for the appended slots)
    if (n_pruned > 0 && K > 0) {
        DVS_LAUNCH(k_hole_flags, total, st, ws->flag_b, N, total, K, ws->flag_a, ws->act);
        CKC(select_indices(ws, ws->flag_a, total, ws->list_a, 0, st));
        CKC(select_indices(ws, ws->act, total, ws->list_b, 1, st));
        CKC(read_counts(ws, 2, st));
        if (ws->h_num[0] != ws->h_num[1]) return cudaErrorAssert;  // holes and movers always pair up
        if (ws->h_num[0] > 0)
            DVS_LAUNCH(k_move_rows, ws->h_num[0], st, p, m1, m2, accum, denom, ws->list_a, ws->list_b, ws->h_num[0]);
    }
    CKC(cudaMemsetAsync(accum, 0, (size_t)std::max(total, K) * sizeof(float), st));
    CKC(cudaMemsetAsync(denom, 0, (size_t)std::max(total, K) * sizeof(float), st));
    *N_io = K;
    if (rep) {
        rep->pruned = n_pruned;
        // clones and splits are not told apart on the host (one list); report the sum under `cloned`
        rep->cloned = n_grow;
    }
    return cudaGetLastError();
}

cudaError_t adc_reset_opacity(Tensors p, Tensors m1, Tensors m2, int64_t N, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    DVS_LAUNCH(k_reset_opacity, N, st, p, m1, m2, N, logitf_(0.01f));
    return cudaGetLastError();
}

}  // namespace dvs_densify

#define DVS_DENSIFY_EXPORT extern "C" __attribute__((visibility("default")))
using namespace dvs_densify;
// ---------------------------------------------------------------------------------------------- test hooks
// The refinement step without a trainer (device pointers; host pointers in the host-emulation build).  p6 / m1_6 / m2_6 / g6 point to
// {means, scales, quats, opac, sh0, shN}; report6 receives {dead, relocated, added, grown, split, pruned}.
static Tensors tensors6(float* const* t) { return Tensors{t[0], t[1], t[2], t[3], t[4], t[5]}; }
static void report6_out(const RefineReport& r, long long* o) {
    if (!o) return;
    o[0] = r.dead; o[1] = r.relocated; o[2] = r.added; o[3] = r.cloned; o[4] = r.split; o[5] = r.pruned;
}
DVS_DENSIFY_EXPORT int dvs_densify_test_mcmc_refine(float* const* p6, float* const* m1_6, float* const* m2_6, long long* N,
                                       long long capacity, long long cap_max, float min_opacity,
                                       unsigned long long seed, long long* report6, void* stream) {
    // (the workspace persists across calls like the trainer's; report6 == nullptr: no report, no host synchronisation)
    static Workspace* ws = workspace_create();
    RefineReport rep;
    int64_t n = *N;
    const cudaError_t e = mcmc_refine(ws, tensors6(p6), tensors6(m1_6), tensors6(m2_6), &n, capacity, cap_max,
                                                   min_opacity, seed, static_cast<cudaStream_t>(stream), report6 ? &rep : nullptr);
    if (report6) cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    *N = n;
    report6_out(rep, report6);
    return (int)e;
}
DVS_DENSIFY_EXPORT int dvs_densify_test_mcmc_noise(float* const* p6, long long N, float step, unsigned long long seed, void* stream) {
    return (int)mcmc_noise(tensors6(p6), N, step, seed, static_cast<cudaStream_t>(stream));
}
DVS_DENSIFY_EXPORT int dvs_densify_test_mcmc_regularise(float* const* p6, float* const* g6, long long N, float w_o, float w_s, void* stream) {
    return (int)mcmc_regularise(tensors6(p6), tensors6(g6), N, w_o, w_s, static_cast<cudaStream_t>(stream));
}
DVS_DENSIFY_EXPORT int dvs_densify_test_adc_accumulate(const float* mean2D_grad, const float* mean2D_abs, const int* radii, float* accum,
                                          float* denom, long long N, void* stream) {
    return (int)adc_accumulate(mean2D_grad, mean2D_abs, radii, accum, denom, N, static_cast<cudaStream_t>(stream));
}
DVS_DENSIFY_EXPORT int dvs_densify_test_adc_refine(float* const* p6, float* const* m1_6, float* const* m2_6, float* accum, float* denom,
                                      long long* N, long long capacity, long long cap_max, const float* cfg6,
                                      unsigned long long seed, long long* report6, void* stream) {
    Workspace* ws = workspace_create();
    RefineReport rep;
    int64_t n = *N;
    const AdcConfig c{cfg6[0], cfg6[1], cfg6[2], cfg6[3], cfg6[4], cfg6[5] != 0.0f};  // 6th: revisedOpacity
    const cudaError_t e = adc_refine(ws, tensors6(p6), tensors6(m1_6), tensors6(m2_6), accum, denom, &n, capacity,
                                                  cap_max, c, seed, static_cast<cudaStream_t>(stream), &rep);
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    workspace_destroy(ws);
    *N = n;
    report6_out(rep, report6);
    return (int)e;
}
DVS_DENSIFY_EXPORT int dvs_densify_test_adc_reset_opacity(float* const* p6, float* const* m1_6, float* const* m2_6, long long N, void* stream) {
    return (int)adc_reset_opacity(tensors6(p6), tensors6(m1_6), tensors6(m2_6), N, static_cast<cudaStream_t>(stream));
}
