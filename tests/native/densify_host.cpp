// TEST HARNESS (not product code): compiles divshot_b200/csrc/densify_ops.h — the per-element arithmetic the CUDA
// kernels of densify.cu call — for the host, so tests/test_densify_ops.py can check it against float64 numpy without
// a GPU.  The product never runs these on the CPU: densify.cu is the only caller in libgstrain.so.
#include <cstdint>
#include <vector>

#include "densify_ops.h"

using namespace dvs_densify;

static std::vector<float> pascal() {
    std::vector<float> b((size_t)kMaxRatio * kMaxRatio, 0.f);
    for (int n = 0; n < kMaxRatio; n++) {
        b[(size_t)n * kMaxRatio] = 1.f;
        for (int k = 1; k <= n; k++) b[(size_t)n * kMaxRatio + k] = b[(size_t)(n - 1) * kMaxRatio + k - 1] + b[(size_t)(n - 1) * kMaxRatio + k];
    }
    return b;
}

extern "C" {
double t_uniform01(uint64_t seed, uint64_t counter) { return uniform01(seed, counter); }
void t_normal2(uint64_t seed, uint64_t counter, float* out2) { normal2(seed, counter, out2[0], out2[1]); }
long long t_sample_cdf(const double* cdf, long long n, double u) { return sample_cdf(cdf, n, u); }
void t_relocation(float o, const float* s, int n, float min_opacity, float* out4) {
    static const std::vector<float> b = pascal();
    relocation(b.data(), o, s, n, min_opacity, out4[0], out4 + 1);
}
void t_mcmc_noise(const float* ls, const float* q, float logit, const float* eps, float step, float* d) {
    mcmc_noise(ls, q, logit, eps, step, d);
}
float t_reg_grad_opacity(float logit, float w, float inv_n) { return reg_grad_opacity(logit, w, inv_n); }
float t_reg_grad_scale(float ls, float w, float inv_3n) { return reg_grad_scale(ls, w, inv_3n); }
int t_adc_decide(float accum, float denom, const float* ls, float logit, float thr, float pd, float extent, float po, float ps) {
    return adc_decide(accum, denom, ls, logit, thr, pd, extent, po, ps);
}
void t_adc_split_sample(const float* mean, const float* ls, const float* q, const float* eps, float* out6) {
    adc_split_sample(mean, ls, q, eps, out6, out6 + 3);
}
}
