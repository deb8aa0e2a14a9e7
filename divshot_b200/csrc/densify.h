// densify.h — host interface of the trainer's refinement step, implemented in densify.cu and
// used by gstrain.cu.  All pointers are device pointers into the trainer's capacity-sized arenas; the parameter arena
// and the two Adam-moment arenas share one layout, so one `Tensors` describes each.
#pragma once
#ifdef DVS_DENSIFY_HOST_EMULATION  // test build of densify.cu for the host (tests/native/densify_emul.cpp)
#include "cuda_host_shim.h"
#else
#include <cuda_runtime.h>
#endif

#include <cstdint>

namespace dvs_densify {

struct Tensors {
    float *means, *scales, *quats, *opac, *sh0, *shN;  // [cap,3] [cap,3] [cap,4] [cap] [cap,3] [cap,15,3]
};

struct Workspace;  // scratch buffers (CDF, sample lists, counters), grown on demand, owned by the trainer
Workspace* workspace_create();
void workspace_destroy(Workspace* ws);

struct RefineReport {
    int64_t dead = 0, relocated = 0, added = 0, cloned = 0, split = 0, pruned = 0;
};

// ---- densifyStrategy 1 (MCMC, the CLI default)
// Relocates every Gaussian with sigmoid(opacity) <= min_opacity onto a live one sampled with probability ~ opacity,
// then grows the set by 5 % (bounded by cap_max and capacity) the same way.  Sources and copies get the relocation
// rule's opacity/scale; Adam moments of every touched Gaussian are zeroed.  *N is updated.  One host synchronisation
// (the dead count).  `seed` makes the sampling reproducible.
cudaError_t mcmc_refine(Workspace* ws, Tensors p, Tensors m1, Tensors m2, int64_t* N, int64_t capacity, int64_t cap_max,
                        float min_opacity, uint64_t seed, cudaStream_t st, RefineReport* rep);
// x += Sigma eps gate(opacity) step   (step = noiselr * lr_xyz), every iteration after the optimizer step
cudaError_t mcmc_noise(Tensors p, int64_t N, float step, uint64_t seed, cudaStream_t st);
// g.opac += w_o/N sigmoid'(logit), g.scales += w_s/(3N) exp(log_scale)   (L1 regularisers of the MCMC strategy)
cudaError_t mcmc_regularise(Tensors p, Tensors g, int64_t N, float w_o, float w_s, cudaStream_t st);

// ---- densifyStrategy 0 / 2 (ADC): statistics after every backward, refinement every `refineEvery` iterations
// accum[i] += ||mean2D_grad[i]|| (or the abs-grad sum when `abs_grad` != nullptr), denom[i] += 1 for visible i
// (`skip`: optional device word, non-zero = this step produced no gradients — its forward overflowed — and is not counted)
cudaError_t adc_accumulate(const float* mean2D_grad, const float* mean2D_abs, const int32_t* radii, float* accum,
                           float* denom, int64_t N, cudaStream_t st, const uint32_t* skip = nullptr);
struct AdcConfig {
    float grad_threshold, percent_dense, extent, prune_opacity, prune_scale3d;
    bool revised_opacity = false;  // `revisedOpacity`: clones / split samples get opacity 1 - sqrt(1 - o)
};
// clone / split / prune in place (holes left by pruning are filled from the tail; clones and second split samples are
// appended).  Statistics are reset, Adam moments of new and split Gaussians zeroed.  *N is updated.
cudaError_t adc_refine(Workspace* ws, Tensors p, Tensors m1, Tensors m2, float* accum, float* denom, int64_t* N,
                       int64_t capacity, int64_t cap_max, const AdcConfig& cfg, uint64_t seed, cudaStream_t st,
                       RefineReport* rep);
// opacity <- min(opacity, logit(0.01)); Adam moments of the opacities zeroed (resetAlphaEvery)
cudaError_t adc_reset_opacity(Tensors p, Tensors m1, Tensors m2, int64_t N, cudaStream_t st);

}  // namespace dvs_densify
