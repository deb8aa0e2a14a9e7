// TEST INFRASTRUCTURE (never linked into the product).  oracle/_ref/libviewerpack_ref.so: the reference's own CPU
// quantisation of a Gaussian model into the viewer's buffers, executed from the reference sources where they lie:
//   * diverse/source/assets/gaussian_model.cpp:14-22    the sigmoid lambda                    -> _ref/gen/vp_sigmoid.inc
//   * diverse/source/assets/gaussian_model.cpp:126-128  constants of create_gpu_buffer        -> _ref/gen/vp_consts.inc
//   * diverse/source/assets/gaussian_model.cpp:130-211  the body of its per-Gaussian lambda   -> _ref/gen/vp_body.inc
//   * diverse/source/assets/gaussian_model.cpp:292-298  the bounding-box loop of update_data  -> _ref/gen/vp_bbox.inc
//   * diverse/source/assets/gaussian_model.h:46-64      Gaussian / PackedVertexSH / PackedVertexColor -> _ref/gen/vp_structs.inc
// are cut out by line range at build time (oracle/Makefile) and compiled unmodified, against the reference's own glm
// (external/glm), base types (diverse_base/source/core/base_type.h) and packing helpers (utility/pack_utils.h), with
// the glm switches of diverse/CMakeLists.txt:89-91.  This file only supplies the surrounding declarations the cut
// lines refer to (the member vectors of GaussianModel, gaussian_model.h:134-139) and a C entry point.
#define GLM_FORCE_INTRINSICS
#define GLM_FORCE_DEPTH_ZERO_TO_ONE
#define GLM_FORCE_SWIZZLE
#include <math.h>

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include <glm/glm.hpp>
#include <glm/gtc/packing.hpp>

#include "core/base_type.h"
#include "utility/pack_utils.h"

namespace diverse {
#include "_ref/gen/vp_sigmoid.inc"
#include "_ref/gen/vp_structs.inc"
}  // namespace diverse

using namespace diverse;

#ifdef _OPENMP
#include <omp.h>
#endif
extern "C" __attribute__((visibility("default"))) int ref_viewer_pack_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

extern "C" __attribute__((visibility("default"))) int ref_viewer_pack(
    const float* pos_d, const float* scales_d, const float* rots_d, const float* opacities_d, const float* shs0_d,
    const float* shsn_d, long long num_gaussians, void* out_gaussians, void* out_colors, void* out_sh, float* bbox6) {
    static_assert(sizeof(Gaussian) == 32 && sizeof(PackedVertexColor) == 8 && sizeof(PackedVertexSH) == 64, "viewer layouts");
    // GaussianModel::update_from_cpu (gaussian_model.cpp:43-68): the members and how they are filled
    std::vector<glm::vec3> pos(num_gaussians);
    std::vector<std::array<float, 3>> shs_0(num_gaussians);
    std::vector<std::array<float, 45>> shs_n(num_gaussians);
    std::vector<float> opacities(num_gaussians);
    std::vector<glm::vec3> scales(num_gaussians);
    std::vector<glm::vec4> rot(num_gaussians);
    memcpy(pos.data(), pos_d, num_gaussians * sizeof(glm::vec3));
    memcpy(rot.data(), rots_d, num_gaussians * sizeof(glm::vec4));
    memcpy(scales.data(), scales_d, num_gaussians * sizeof(glm::vec3));
    memcpy(opacities.data(), opacities_d, num_gaussians * sizeof(f32));
    memcpy(shs_0.data(), shs0_d, num_gaussians * sizeof(f32) * 3);
    memcpy(shs_n.data(), shsn_d, num_gaussians * sizeof(f32) * 45);
    {
#include "_ref/gen/vp_bbox.inc"
        for (int a = 0; a < 3; a++) { bbox6[a] = minn[a]; bbox6[3 + a] = maxx[a]; }
    }
    // GaussianModel::create_gpu_buffer (gaussian_model.cpp:119-121): the three staging vectors
    std::vector<Gaussian> gaussians(pos.size());
    std::vector<PackedVertexColor> gaussians_sh_0(pos.size());
    std::vector<PackedVertexSH> gaussians_sh_n(pos.size());
#include "_ref/gen/vp_consts.inc"
    (void)t11; (void)t10;
    // parallel_for<size_t>(0, gaussians.size(), [&](size_t k) {  -- the reference's thread pool; OpenMP over the same body here
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < gaussians.size(); k++) {
#include "_ref/gen/vp_body.inc"
    }
    memcpy(out_gaussians, gaussians.data(), gaussians.size() * sizeof(Gaussian));
    memcpy(out_colors, gaussians_sh_0.data(), gaussians.size() * sizeof(PackedVertexColor));
    memcpy(out_sh, gaussians_sh_n.data(), gaussians.size() * sizeof(PackedVertexSH));
    return 0;
}
