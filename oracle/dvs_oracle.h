/*
 * dvs_oracle.h — CPU ORACLE for the differentiable 3DGS rasterize path (TEST INFRASTRUCTURE).
 *
 * This is test infrastructure, not product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, link or call it.
 *
 * PARITY UNPINNED: the reference's implementation of this path (diverse_utils/gsplatrast,
 * named at /root/reference/diverse_utils/CMakeLists.txt:1-3, CMakeLists.txt:101-104) is NOT in
 * the reference tree (README.md:32,46), there is no binary, no test and no golden vector.
 * The arithmetic lives in a third-party dependency that is absent from /root/reference and
 * un-pinned: graphdeco-inria/diff-gaussian-rasterization, credited at README.md:95.  This
 * file restates that published algorithm and is anchored on the
 * in-tree corroboration each function cites (HLSL transliterations of the same forward maths
 * and the tensor conventions of the only in-tree consumer).  Its analytic backward is pinned
 * independently by a float64 torch-autograd re-expression (tests/autograd_ref.py) and by
 * closed-form known-answer cases (tests/test_oracle_kat.py).
 * PARTLY PINNED since: the reference's shader functions for cov3D, EWA cov2D (clamp, blur),
 * ndc2Pix, SH evaluation and the anti-aliasing factor are compiled as C++ from /root/reference
 * (oracle/_ref/libhlsl_ref.so) and orc_preprocess_fwd is checked against them
 * (tests/test_oracle_vs_reference_hlsl.py).  Culling, radius, tile rects, binning order,
 * compositing and both backward passes have no executable counterpart in the reference.
 *
 * Conventions: matrices are flat float[16], element m[4*c + r] = row r,
 * column c of the column-vector 4x4 (same indexing as gsplat_vs.hlsl:54-72).  Quaternion order
 * (r,x,y,z) (gsplat_vs.hlsl:189-200).  Raw parameters: log-scale, logit-opacity, un-normalised
 * quaternion, SH [N,K,3] RGB-interleaved (gaussian_model.cpp:43-68,145-157,579-583).
 */
#ifndef DVS_ORACLE_H
#define DVS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FLAG_INPUT_ACTIVATED 1 /* scales/opacity/quats are already activated (upstream-style inputs) */
#define ORC_FLAG_ANTIALIAS 2       /* mip-splatting opacity compensation sqrt(max(0, det(S')/det(S'+0.3I))), gsplat_vs.hlsl:296-301 */

typedef struct {
    float view[16];   /* world->view, flat m[4c+r] */
    float proj[16];   /* full view-projection, flat m[4c+r] */
    float campos[3];
    float tanfovx, tanfovy;
    int32_t width, height;
    float bg[3];
    float scale_modifier;
    int32_t sh_degree;     /* active degree 0..3 */
    int32_t sh_rest_alloc; /* rest coefficients allocated per Gaussian in shN (>= (deg+1)^2-1) */
    int32_t flags;
} orc_camera;

/* A1: per-Gaussian forward.  All outputs caller-allocated, length N (x components). */
void orc_preprocess_fwd(const orc_camera* cam, int32_t N,
                        const float* means3D, const float* scales, const float* quats,
                        const float* opacities, const float* sh0, const float* shN,
                        float* depth, int32_t* radii, float* mean2D /*[N,2]*/,
                        float* cov3D /*[N,6]*/, float* conic_opacity /*[N,4]*/,
                        float* rgb /*[N,3]*/, uint8_t* clamped /*[N,3]*/,
                        uint32_t* tiles_touched, int32_t* rect /*[N,4] minx,miny,maxx,maxy*/);

/* A2: inclusive scan; returns D. */
int64_t orc_scan_tiles(int32_t N, const uint32_t* tiles_touched, uint32_t* point_offsets);

/* A3-A5: duplicate, stable sort by (tile, depth bits), ranges. */
void orc_bin_sort(const orc_camera* cam, int32_t N, const float* depth, const int32_t* radii,
                  const int32_t* rect, const uint32_t* point_offsets, int64_t D,
                  uint64_t* keys_sorted /*[D]*/, uint32_t* point_list /*[D]*/,
                  uint32_t* ranges /*[T,2]*/);

/* literal A4/A5 (single stable radix sort of already-emitted keys); test cross-check of orc_bin_sort's parallel sort */
void orc_radix_check(int32_t T, int64_t D, uint64_t* keys, uint32_t* point_list, uint32_t* ranges);

/* A6: tile compositing forward.  fragile[p]=1 when a threshold decision of pixel p is within
 * a few ulp of flipping (alpha vs 1/255, T vs 1e-4, power vs 0) — integer outputs of such
 * pixels may legitimately differ on hardware with a different exp().  threads<=0: all cores. */
void orc_render_fwd(const orc_camera* cam, const uint32_t* ranges, const uint32_t* point_list,
                    const float* mean2D, const float* conic_opacity, const float* rgb,
                    float* out_color /*[3,H,W]*/, float* final_T /*[H*W]*/,
                    uint32_t* n_contrib /*[H*W]*/, uint8_t* fragile /*[H*W] or NULL*/,
                    int32_t threads);

/* A7: reverse-walk backward.  Outputs are zeroed then accumulated. */
void orc_render_bwd(const orc_camera* cam, int32_t N, const uint32_t* ranges,
                    const uint32_t* point_list, const float* mean2D, const float* conic_opacity,
                    const float* rgb, const float* final_T, const uint32_t* n_contrib,
                    const float* dL_dpix /*[3,H,W]*/,
                    float* dL_dmean2D /*[N,2] (ndc-scaled: x W/2, H/2)*/,
                    float* dL_dmean2D_abs /*[N,2] sum |.| or NULL*/,
                    float* dL_dconic /*[N,3]: dA, dB(total), dC*/, float* dL_dopacity /*[N]*/,
                    float* dL_dcolor /*[N,3]*/, int32_t threads);

/* A8: per-Gaussian backward to the stored (raw unless INPUT_ACTIVATED) parameters. */
void orc_preprocess_bwd(const orc_camera* cam, int32_t N,
                        const float* means3D, const float* scales, const float* quats,
                        const float* opacities, const float* sh0, const float* shN,
                        const int32_t* radii, const uint8_t* clamped,
                        const float* dL_dmean2D, const float* dL_dconic,
                        const float* dL_dopacity_act, const float* dL_dcolor,
                        float* dL_dmeans3D, float* dL_dscales, float* dL_dquats,
                        float* dL_dopacities, float* dL_dsh0, float* dL_dshN);

/* ---- 2DGS ("surfel") variant, GaussianTrainConfig::modelType = 1 (spec: the S.1-S.4 comment in dvs_oracle.c).  Binning and
 * sorting are shared with the 3DGS path (orc_scan_tiles, orc_bin_sort on depth / radii / rect). */
void orc2_preprocess_fwd(const orc_camera* cam, int32_t N, const float* means3D, const float* scales,
                         const float* quats, const float* opacities, const float* sh0, const float* shN,
                         float* depth, int32_t* radii, float* mean2D /*[N,2]*/, float* transmat /*[N,9]: Tu, Tv, Tw*/,
                         float* opacity_act /*[N]*/, float* rgb, uint8_t* clamped, uint32_t* tiles_touched, int32_t* rect);
void orc2_render_fwd(const orc_camera* cam, const uint32_t* ranges, const uint32_t* point_list, const float* mean2D,
                     const float* transmat, const float* opacity_act, const float* rgb, float* out_color,
                     float* final_T, uint32_t* n_contrib, uint8_t* fragile, int32_t threads);
void orc2_render_bwd(const orc_camera* cam, int32_t N, const uint32_t* ranges, const uint32_t* point_list,
                     const float* mean2D, const float* transmat, const float* opacity_act, const float* rgb,
                     const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float* dL_dT /*[N,9]*/,
                     float* dL_dmean2D /*[N,2], pixel units*/, float* dL_dopacity /*[N]*/, float* dL_dcolor /*[N,3]*/,
                     int32_t threads);
void orc2_preprocess_bwd(const orc_camera* cam, int32_t N, const float* means3D, const float* scales,
                         const float* quats, const float* opacities, const float* sh0, const float* shN,
                         const int32_t* radii, const uint8_t* clamped, const float* dL_dT, const float* dL_dmean2D,
                         const float* dL_dopacity_act, const float* dL_dcolor, float* dL_dmeans3D,
                         float* dL_dscales, float* dL_dquats, float* dL_dopacities, float* dL_dsh0, float* dL_dshN);

/* OpenMP thread count for all oracle stages (n <= 0: leave as is) */
void orc_set_threads(int32_t n);

/* deterministic fp32 exp used by the activations (bit-identical on CPU and GPU by construction) */
float orc_expf(float x);

#ifdef __cplusplus
}
#endif
#endif
