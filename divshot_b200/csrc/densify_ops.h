// densify_ops.h — per-element arithmetic of the trainer's refinement step ("densify/prune stats,
// MCMC noise `noiselr`, `capMax`"), shared by the CUDA kernels of densify.cu and by a host-compiled test harness
// (tests/native/densify_host.cpp) so the maths is checked on the CPU against float64 numpy before it runs on a GPU.
//
// Reference surface: the closed trainer's options `densifyStrategy` (0 ADC, 1 MCMC — the CLI default, 2 ADC+),
// `refineEvery`, `warmupLength`, `refineStopIter`, `capMax`, `noiselr`, `min_opacity`, `growGrad2d`, `pruneOpacity`,
// `pruneScale3d`, `resetAlphaEvery` (application/diverseshot-cli/source/main.cpp:19-70, gs_train.cpp:50-99,
// docs/userGuide.md:39-42).  The implementation is absent from the reference; docs/userGuide.md:41
// names the algorithm: "3D Gaussian Splatting as Markov Chain Monte Carlo" (arXiv 2404.09591), restated here:
//   * relocation (eq. 9 of the paper): a Gaussian of opacity o and scale s that is split into n copies gets
//       o' = 1 - (1 - o)^(1/n),   s' = s * o / sum_{i=1..n} sum_{k=0..i-1} C(i-1,k) (-1)^k o'^(k+1) / sqrt(k+1)
//     so that the rendered contribution of the n copies matches the original;
//   * exploration noise: x += Sigma * eps * sigmoid(-100 (opacity - 0.005)) * noiselr * lr_xyz,  eps ~ N(0, I),
//     Sigma = R diag(s)^2 R^T;
//   * L1 regularisers 0.01 * mean(opacity) + 0.01 * mean(scale).
// ADC (clone / split / prune / opacity reset) is the classic 3DGS rule set (credited at README.md:99).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define DVS_HD __host__ __device__ __forceinline__
#else
#define DVS_HD inline
#endif

namespace dvs_densify {

constexpr int kMaxRatio = 51;  // copies of one Gaussian considered by the relocation rule (binomial table size)

// ---- counter-based RNG: one 64-bit mix per draw, no state (the same (seed, counter) gives the same number anywhere)
DVS_HD uint64_t mix64(uint64_t seed, uint64_t counter) {
    uint64_t z = seed + (counter + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
DVS_HD double uniform01(uint64_t seed, uint64_t counter) {  // [0, 1), 53 bits
    return (double)(mix64(seed, counter) >> 11) * (1.0 / 9007199254740992.0);
}
// two independent standard normals from one counter (Box-Muller on two 32-bit halves)
DVS_HD void normal2(uint64_t seed, uint64_t counter, float& a, float& b) {
    const uint64_t z = mix64(seed, counter);
    const float u1 = ((float)(uint32_t)(z >> 40) + 1.0f) * (1.0f / 16777216.0f);      // (0, 1]
    const float u2 = (float)(uint32_t)((z >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);   // [0, 1)
    const float r = sqrtf(-2.0f * logf(u1));
    const float t = 6.28318530717958647692f * u2;
    a = r * cosf(t);
    b = r * sinf(t);
}

// ---- sampling: index of the first cdf entry > u  (cdf = inclusive prefix sums of non-negative weights, n >= 1)
DVS_HD int64_t sample_cdf(const double* cdf, int64_t n, double u) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

DVS_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
DVS_HD float logitf_(float p) { return logf(p / (1.0f - p)); }

// ---- MCMC relocation rule.  binom = row-major Pascal triangle [kMaxRatio][kMaxRatio] (C(n,k) at [n*kMaxRatio+k]).
// In: activated opacity o in (0,1), activated scale s[3], n = number of copies (>= 1).  Out: raw (logit / log)
// parameters of each copy, opacity clamped to [min_opacity, 1 - 2^-24] like the credited implementation.
DVS_HD void relocation(const float* binom, float o, const float s[3], int n, float min_opacity, float& new_logit,
                       float new_log_scale[3]) {
    n = n < 1 ? 1 : (n > kMaxRatio ? kMaxRatio : n);
    const float o_new = 1.0f - powf(1.0f - o, 1.0f / (float)n);
    float denom = 0.0f;
    for (int i = 1; i <= n; i++) {
        float pw = o_new;  // o_new^(k+1)
        for (int k = 0; k <= i - 1; k++) {
            const float term = binom[(i - 1) * kMaxRatio + k] * pw / sqrtf((float)(k + 1));
            denom += (k & 1) ? -term : term;
            pw *= o_new;
        }
    }
    const float coeff = o / denom;
    const float oc = fminf(fmaxf(o_new, min_opacity), 1.0f - 5.9604645e-8f);
    new_logit = logitf_(oc);
    for (int a = 0; a < 3; a++) new_log_scale[a] = logf(coeff * s[a]);
}

// ---- rotation matrix of a (not necessarily unit) quaternion (r,x,y,z), normalised first — the trainer's convention
DVS_HD void quat_to_R(const float q[4], float R[9]) {
    const float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float r = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - r * z);       R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);       R[7] = 2.f * (y * z + r * x);       R[8] = 1.f - 2.f * (x * x + y * y);
}

// ---- MCMC exploration noise for one Gaussian: d = Sigma * eps * gate * step,  Sigma = R diag(exp(ls))^2 R^T,
// gate = sigmoid(-100 (opacity - 0.005)) = 1 / (1 + exp(-100 ((1 - opacity) - 0.995)))
DVS_HD void mcmc_noise(const float log_scale[3], const float quat[4], float logit_opacity, const float eps[3], float step,
                       float d[3]) {
    float R[9];
    quat_to_R(quat, R);
    const float s2[3] = {expf(2.f * log_scale[0]), expf(2.f * log_scale[1]), expf(2.f * log_scale[2])};
    const float gate = 1.0f / (1.0f + expf(-100.0f * ((1.0f - sigmoidf_(logit_opacity)) - 0.995f)));
    // t = R^T eps ; t *= s^2 ; d = R t
    float t[3];
    for (int k = 0; k < 3; k++) t[k] = (R[k] * eps[0] + R[3 + k] * eps[1] + R[6 + k] * eps[2]) * s2[k];
    const float g = gate * step;
    for (int r = 0; r < 3; r++) d[r] = (R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2]) * g;
}

// ---- gradients of the MCMC regularisers  w_o * mean(sigmoid(logit)) + w_s * mean(exp(log_scale))
DVS_HD float reg_grad_opacity(float logit, float w_o, float inv_n) {
    const float o = sigmoidf_(logit);
    return w_o * inv_n * o * (1.0f - o);
}
DVS_HD float reg_grad_scale(float log_scale, float w_s, float inv_3n) { return w_s * inv_3n * expf(log_scale); }

// ---- ADC decisions for one Gaussian (classic 3DGS densify_and_prune).
//   grad  = accumulated ||dL/dmean2D|| / visibility count (0 if never visible)
//   clone : grad >= thr and max scale <= percent_dense * extent       (under-reconstruction: duplicate in place)
//   split : grad >= thr and max scale >  percent_dense * extent       (over-reconstruction: two samples, scale / 1.6)
//   prune : opacity < prune_opacity, or max scale > prune_scale3d * extent
enum AdcAction : uint8_t { ADC_KEEP = 0, ADC_CLONE = 1, ADC_SPLIT = 2, ADC_PRUNE = 4 };
DVS_HD uint8_t adc_decide(float grad_accum, float denom, const float log_scale[3], float logit_opacity, float thr,
                          float percent_dense, float extent, float prune_opacity, float prune_scale3d) {
    const float g = denom > 0.f ? grad_accum / denom : 0.f;
    const float smax = expf(fmaxf(log_scale[0], fmaxf(log_scale[1], log_scale[2])));
    uint8_t act = ADC_KEEP;
    if (g >= thr) act = smax <= percent_dense * extent ? ADC_CLONE : ADC_SPLIT;
    if (sigmoidf_(logit_opacity) < prune_opacity || smax > prune_scale3d * extent) act |= ADC_PRUNE;
    return act;
}
// `revisedOpacity` ("Revising Densification in Gaussian Splatting", arXiv 2404.06109, eq. 9): a cloned or split Gaussian
// and its copy each get opacity 1 - sqrt(1 - o), so that the pair composites like the original instead of more opaquely
DVS_HD float revised_opacity_logit(float logit) {
    const float o = sigmoidf_(logit);
    const float o2 = fminf(fmaxf(1.0f - sqrtf(1.0f - o), 1e-7f), 1.0f - 5.9604645e-8f);
    return logitf_(o2);
}
// position of a split sample: mean + R diag(s) eps; its log-scale: log(s / 1.6)
DVS_HD void adc_split_sample(const float mean[3], const float log_scale[3], const float quat[4], const float eps[3],
                             float out_mean[3], float out_log_scale[3]) {
    float R[9];
    quat_to_R(quat, R);
    float t[3];
    for (int k = 0; k < 3; k++) t[k] = expf(log_scale[k]) * eps[k];
    for (int r = 0; r < 3; r++) out_mean[r] = mean[r] + R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2];
    for (int k = 0; k < 3; k++) out_log_scale[k] = log_scale[k] - 0.47000362924573563f;  // log(1.6)
}

}  // namespace dvs_densify
