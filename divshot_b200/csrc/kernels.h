// kernels.h — host-side launcher declarations (one per stage of SURVEY.md §8 A1-A8).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace dvs {

// A1 (+ per-tile duplicate counts, or with fe.bin_stride > 0 the duplicate emission itself: single-pass binning)
cudaError_t launch_preprocess_fwd(const Cam& cam, int N, const Params& prm, float4* rec, uint4* aux,
                                  uint32_t* tile_count, int32_t* out_radii, unsigned long long* stats,
                                  const FusedEmit& fe, cudaStream_t st);

// A2/A5: exclusive scan of tile counts -> tile_base[T+1], cursor[T]; info[0]=D, info[1]=max len, info[2]=overflow,
// info[4..8] = tiles per sort class, class_tiles[5][T] = their ids; tile_order[T] (optional) = all tiles, longest lists
// first (a launch order for the compositing CTAs; measured useless at c3, profiles/r2_ab1_loops_tight_lpt.json).
// Also the forward's housekeeping: zeroes the tile counters it has read, publishes V / D (stats -> info[12..15]) and
// clears the accumulators and the bin-overflow word for the next forward (no memset launches in the step).
cudaError_t launch_tile_scan(int T, uint32_t* tile_count, uint32_t* tile_base, uint32_t* tile_cursor,
                             uint32_t* info, uint32_t dup_capacity, uint32_t* class_tiles, uint32_t* tile_order,
                             unsigned long long* stats, cudaStream_t st);

// A3 (two-pass mode): emit (depth | id | sub-tile mask) entries into per-tile bins
// (`cull`: where the sub-tile cull ellipse of Gaussian i lies — cull[cull_stride * i], cull[cull_stride * i + 1] in the layout
//  cull_params() reads: the 48-byte records themselves (stride 3) or the 2DGS path's own two words (stride 2))
cudaError_t launch_emit(const Cam& cam, int N, const uint4* aux, const float4* cull, int cull_stride, uint32_t* tile_cursor,
                        unsigned long long* bins, uint32_t dup_capacity, cudaStream_t st);

// A4: tile-local sort (CUB-free) -> plist (id<<8 | mask), tile-major
int tile_sort_launch_count(uint32_t bin_stride);
cudaError_t launch_tile_sort(int T, uint32_t bin_stride, const uint32_t* tile_base, unsigned long long* bins,
                             uint32_t* plist, const uint32_t* info, const uint32_t* class_tiles, cudaStream_t st);

// A6
cudaError_t launch_render_fwd(const Cam& cam, const uint32_t* tile_order, const uint32_t* tile_base,
                              const uint32_t* plist, const float4* rec, float* out_color, float* final_T,
                              uint32_t* n_contrib, const uint32_t* info, cudaStream_t st);

// A7
cudaError_t launch_render_bwd(const Cam& cam, const uint32_t* tile_order, const uint32_t* tile_base,
                              const uint32_t* plist, const float4* rec, const float* final_T, const uint32_t* n_contrib,
                              const float* dL_dpix, float* sgrad, bool absgrad, const uint32_t* info, cudaStream_t st);

// A8
cudaError_t launch_preprocess_bwd(const Cam& cam, int N, const Params& prm, const uint4* aux, float4* sgrad,
                                  const Grads& g, uint32_t flags, cudaStream_t st);

// F4 auxiliary outputs by linearity (aux_outputs.cu): records with the colour replaced by (depth, 1, 0); dL/ddepth per
// Gaussian moved out of the screen-gradient records; its contribution to dL/dmean
cudaError_t launch_aux_records(int N, const float4* rec, float4* rec_aux, cudaStream_t st);
cudaError_t launch_aux_extract(int N, float4* sgrad, float* dz, cudaStream_t st);
cudaError_t launch_aux_depth_grad(int N, const float* dz, const float view_row2[3], float* dmeans, cudaStream_t st);
// ... and the normal map: records with the colour replaced by the view-space normal; dL/dn per Gaussian; its chain to dL/dquat
cudaError_t launch_aux_normal_records(const Cam& cam, int N, const Params& prm, const float4* rec, float4* rec_aux, cudaStream_t st);
cudaError_t launch_aux_extract3(int N, float4* sgrad, float* dn, cudaStream_t st);
cudaError_t launch_aux_normal_grad(const Cam& cam, int N, const Params& prm, const float* dn, float* dquats, cudaStream_t st);

// 2DGS ("surfel") variant, GaussianTrainConfig::modelType = 1 (preprocess_fwd.cu / surfel.cu): per-Gaussian forward into the
// 64-byte homography records rec2 (+ a 3DGS-shaped stand-in in rec, aux, tile counts, so binning / sorting run unchanged),
// compositing forward, reverse-walk backward into the 64-byte records sgrad2, per-Gaussian backward
cudaError_t launch_surfel_preprocess_fwd(const Cam& cam, int N, const Params& prm, float4* rec, float4* rec2, float4* cull2, uint4* aux,
                                         uint32_t* tile_count, int32_t* out_radii, unsigned long long* stats, cudaStream_t st);
cudaError_t launch_surfel_render_fwd(const Cam& cam, const uint32_t* tile_base, const uint32_t* plist, const float4* rec2,
                                     float* out_color, float* final_T, uint32_t* n_contrib, const uint32_t* info, cudaStream_t st);
cudaError_t launch_surfel_render_bwd(const Cam& cam, const uint32_t* tile_base, const uint32_t* plist, const float4* rec2,
                                     const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float* sgrad2,
                                     const uint32_t* info, cudaStream_t st);
cudaError_t launch_surfel_preprocess_bwd(const Cam& cam, int N, const Params& prm, const uint4* aux, float4* sgrad2, const Grads& g,
                                         uint32_t flags, cudaStream_t st);

// background model: dL/dbg[ch][p] = final_T[p] * dL/dpix[ch][p]  (out = C + final_T * bg)
cudaError_t launch_background_grad(int64_t P, const float* final_T, const float* dL_dpix, float* dL_dbg, const uint32_t* info,
                                   cudaStream_t st);

// debug helpers (parity tests): unpack records into the upstream-style arrays
cudaError_t launch_unpack(int N, const float4* rec, int32_t* radii, uint32_t* tiles, float* depth, float* mean2D,
                          float* conic_opacity, float* rgb, uint8_t* clamped, cudaStream_t st);
cudaError_t launch_unpack_plist(uint32_t D, const uint32_t* plist, uint32_t* ids, uint8_t* masks, cudaStream_t st);

}  // namespace dvs
