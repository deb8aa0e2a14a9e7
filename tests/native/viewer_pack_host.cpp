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
// The kernel of viewer_pack.cu thread by thread: the same two phase functions and tile loop, 128 "threads" per CTA, `grid`
// CTAs, the shared rows a plain array poisoned before every trip; per-thread bounding boxes are reduced at the end (the
// kernel does that with shuffles and integer atomics).
void t_viewer_pack_kernel_emulation(const float* means, const float* scales, const float* quats, const float* opac, const float* sh0,
                                    const float* shN, long long N, uint32_t* out_g, uint32_t* out_c, uint32_t* out_sh,
                                    uint32_t* bbox_ordered, int grid) {
    const int T = 128;
    alignas(16) static float rows[128 * kShRest];
    const PackArgs a{means, scales, quats, opac, sh0, shN, N, out_g, out_c, out_sh, (int)((reinterpret_cast<unsigned long long>(shN) & 15ull) == 0)};
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    const long long n_tiles = (N + T - 1) / T;
    for (int block = 0; block < grid; block++) {
        float lo[128][3], hi[128][3];
        for (int t = 0; t < T; t++)
            for (int k = 0; k < 3; k++) { lo[t][k] = FLT_MAX; hi[t][k] = -FLT_MAX; }
        for (long long tile = block; tile < n_tiles; tile += grid) {
            const long long base = tile * T;
            const int cnt = (int)(N - base < T ? N - base : T);
            for (int k = 0; k < 128 * kShRest; k++) rows[k] = -12345.0f;
            for (int tid = 0; tid < T; tid++) pack_stage(a, rows, tid, T, base, cnt);
            for (int tid = 0; tid < T; tid++) pack_compute(a, rows, tid, base, cnt, lo[tid], hi[tid]);
        }
        for (int t = 0; t < T; t++)
            for (int k = 0; k < 3; k++) { blo[k] = lo[t][k] < blo[k] ? lo[t][k] : blo[k]; bhi[k] = bhi[k] < hi[t][k] ? hi[t][k] : bhi[k]; }
    }
    for (int k = 0; k < 3; k++) { bbox_ordered[k] = f32_to_ordered(blo[k]); bbox_ordered[3 + k] = f32_to_ordered(bhi[k]); }
}
uint32_t t_f32_to_f16_glm(float f) { return f32_to_f16_glm(f); }
uint32_t t_f32_to_ordered(float f) { return f32_to_ordered(f); }
float t_ordered_to_f32(uint32_t o) { return ordered_to_f32(o); }
}
