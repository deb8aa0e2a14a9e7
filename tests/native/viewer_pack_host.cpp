// TEST HARNESS (not product code): compiles divshot_b200/csrc/viewer_pack_ops.h — the per-Gaussian arithmetic the CUDA
// kernel of viewer_pack.cu calls — for the host, so tests/test_viewer_pack.py can compare it byte for byte with the
// reference's own quantisation (oracle/_ref/libviewerpack_ref.so) without a GPU.  Built with -ffp-contract=off.
#include <cfloat>
#include <cstdint>

#include "viewer_pack_ops.h"

using namespace dvs_vp;

extern "C" {
void t_viewer_pack(const float* means, const float* scales, const float* quats, const float* opac, const float* sh0,
                   const float* shN, long long N, uint32_t* out_g, uint32_t* out_c, uint32_t* out_sh, uint32_t* bbox_ordered) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (long long i = 0; i < N; i++) {
        pack_geometry(means + 3 * i, quats + 4 * i, scales + 3 * i, opac[i], out_g + 8 * i);
        pack_color(sh0 + 3 * i, out_c + 2 * i);
        float c[kShRest];
        for (int j = 0; j < kShRest; j++) c[j] = shN[kShRest * i + j];
        pack_sh_rest(c, out_sh + 16 * i);
        for (int a = 0; a < 3; a++) {
            const float p = means[3 * i + a];
            lo[a] = p < lo[a] ? p : lo[a];
            hi[a] = hi[a] < p ? p : hi[a];
        }
    }
    for (int a = 0; a < 3; a++) { bbox_ordered[a] = f32_to_ordered(lo[a]); bbox_ordered[3 + a] = f32_to_ordered(hi[a]); }
}
uint32_t t_f32_to_f16_glm(float f) { return f32_to_f16_glm(f); }
uint32_t t_f32_to_ordered(float f) { return f32_to_ordered(f); }
float t_ordered_to_f32(uint32_t o) { return ordered_to_f32(o); }
}
