/*
 * dvs_viewer_pack.h — C-ABI of the trainer -> viewer hand-off.
 *
 * Replaces, for a training run, the reference's path
 *     6 x GaussianTrainerScene::getGaussian*Cpu()                      (application/editor/source/editor.cpp:1559-1566)
 *  -> GaussianModel::update_from_cpu -> update_data -> create_gpu_buffer
 *                                            (diverse/source/assets/gaussian_model.cpp:43-68, 290-302, 115-212)
 * i.e. a 236 B/Gaussian device->host copy followed by a CPU quantisation pass, by ONE kernel that reads the raw
 * parameters where the trainer keeps them and writes the three buffers the splat viewer binds, byte-identical to what
 * create_gpu_buffer produces (struct layouts: diverse/source/assets/gaussian_model.h:46-64):
 *     gaussians  [N] x 32 B   Gaussian          { vec4 position (w = 0) ; uvec4 rotation_scale }
 *     colors     [N] x  8 B   PackedVertexColor  uvec2
 *     sh         [N] x 64 B   PackedVertexSH    { uvec4 sh1to3, sh4to7, sh8to11, sh12to15 }
 * and the bounding box of the positions (gaussian_model.cpp:292-299).
 *
 * All pointers are DEVICE pointers on the current device; out_* must be 16-byte aligned.  There is no CPU path.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVS_VP_GAUSSIAN_BYTES 32
#define DVS_VP_COLOR_BYTES 8
#define DVS_VP_SH_BYTES 64

/* Packs Gaussians [0, N).  means[N,3] scales[N,3] (log) quats[N,4] (r,x,y,z, any norm) opacities[N] (logit)
 * sh0[N,3] shN[N,15,3] — the layouts of gaussian_model.cpp:60-65.  bbox_ordered[6] (device, uint32) receives
 * min.xyz, max.xyz in the order-preserving integer encoding decoded by dvs_viewer_pack_decode_bbox; it is initialised
 * by this call.  Asynchronous on `stream`.  Returns 0 or a cudaError_t value. */
int dvs_viewer_pack(const float* means, const float* scales, const float* quats, const float* opacities, const float* sh0,
                    const float* shN, int64_t N, void* out_gaussians, void* out_colors, void* out_sh,
                    uint32_t* bbox_ordered, void* stream);

/* Host helper: bbox_ordered[6] (copied to the host by the caller) -> min.xyz, max.xyz.  For N == 0 the box is the
 * reference's empty box (+FLT_MAX.. , -FLT_MAX..). */
void dvs_viewer_pack_decode_bbox(const uint32_t* bbox_ordered_host, float* min_xyz, float* max_xyz);

#ifdef __cplusplus
}
#endif
