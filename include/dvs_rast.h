/*
 * dvs_rast.h — C-ABI of the B200-native differentiable 3D-Gaussian-splatting rasterizer.
 *
 * This is the drop-in boundary for DIVSHOT's closed `diverse_utils/gsplatrast` operator
 * (named at /root/reference/diverse_utils/CMakeLists.txt:1, CMakeLists.txt:103,
 * premake-dependencies.lua:40; source absent from the reference tree — README.md:32,46).
 * The trainer plugin (`libgstrain.so`, include/gaussian_trainer_scene.hpp, symbols resolved at
 * application/diverseshot-cli/source/gs_train.cpp:24-179) and the libtorch
 * `torch::CustomClassHolder` operator both sit on top of exactly these entry points.
 *
 * Plain C: pointers and sizes only, no torch / CUDA types (streams travel as void*).
 * All data pointers are DEVICE pointers unless the name ends in `_host`; they must be
 * 16-byte aligned and contiguous fp32.  One context per (device, stream-at-a-time); a context
 * is not thread-safe.  Every entry point returns 0 on success or a negative DVS_E_* code and
 * records a message retrievable with dvs_rast_last_error().
 *
 * Parameter conventions (the tensors the trainer exports, proven by the only in-tree consumer
 * diverse/source/assets/gaussian_model.cpp:43-68,145-157,579-583):
 *   means3D [N,3]; log_scales [N,3] (exp activation); quats [N,4] (r,x,y,z, normalised inside);
 *   logit_opacities [N] (sigmoid activation); sh0 [N,3]; shN [N,sh_rest_alloc,3] RGB-interleaved.
 * With DVS_FLAG_INPUT_ACTIVATED the scale / quaternion / opacity inputs are taken as already
 * activated (the credited upstream operator's convention) and gradients are w.r.t. those.
 */
#ifndef DVS_RAST_H
#define DVS_RAST_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVS_API __attribute__((visibility("default")))

#define DVS_OK 0
#define DVS_E_INVALID (-1)     /* bad argument (null / misaligned pointer, bad size, bad degree) */
#define DVS_E_CUDA (-2)        /* a CUDA runtime call failed; see dvs_rast_last_error */
#define DVS_E_NOMEM (-3)       /* arena allocation failed */
#define DVS_E_STATE (-4)       /* backward without a matching forward, etc. */
#define DVS_E_UNSUPPORTED (-5) /* N >= 2^24, more than 2^24 tiles, ... */
#define DVS_E_OVERFLOW (-6)    /* a forward run with DVS_FLAG_DEFER_CHECK overflowed the binning arena: its outputs
                                  (and those of its backward) are invalid; the arena has been grown, redo the step */

#define DVS_FLAG_INPUT_ACTIVATED 1u /* scales/quats/opacities already activated */
#define DVS_FLAG_ACCUMULATE 2u      /* backward: add into the gradient buffers instead of overwriting */
#define DVS_FLAG_ABSGRAD 4u         /* backward: also write sum|dL/dmean2D| (densify statistic, main.cpp:44-45) */
#define DVS_FLAG_ANTIALIAS 16u      /* mip-splatting opacity compensation: opacity *= sqrt(max(0, det(S')/det(S'+0.3I)))
                                       (GaussianTrainConfig::mipAntiliased, docs/userGuide.md:58; gsplat_vs.hlsl:296-301) */
#define DVS_FLAG_DEFER_CHECK 8u     /* forward: do not synchronise the stream to validate the binning arena; the check is
                                       made by a later call (non-blocking) or by dvs_rast_get_stats (blocking).  Only
                                       honoured once a synchronous forward has sized the arena. */

#define DVS_FLAG_SKIP_SHN_GRAD 64u  /* backward: do not write grads->shN (180 of the 236 B per Gaussian at SH degree 3): the caller forms
                                     * the summed dL/dshN itself from every view's dL/dsh0 (dvs_coll_exchange_fused); not with ACCUMULATE */
#define DVS_FLAG_MODEL_2DGS 128u    /* forward (the backward follows the forward's model): 2D Gaussian splatting, GaussianTrainConfig::modelType
                                     * = 1 — a Gaussian is a flat disk (tangents R[:,0], R[:,1], scales[0..1]; scales[2] ignored, gradient 0),
                                     * rendered by ray-splat intersection with the object-space low-pass filter; same tensors, same outputs.
                                     * Two-pass binning, whole-rectangle lists (DEFER_CHECK is honoured, TIGHT_LISTS / ANTIALIAS ignored) */
#define DVS_FLAG_TIGHT_LISTS 32u    /* forward, only together with DVS_FLAG_DEFER_CHECK (the training-loop mode): entries whose
                                       {alpha >= 1/255} footprint misses their tile are not put into the tile lists.  Image,
                                       final_T and gradients are unchanged; point_list / ranges are the whole-rectangle lists
                                       of the credited rasterizer with exactly those entries removed, and n_contrib counts
                                       positions in the shortened lists.  Without the flag the lists are the reference's. */

typedef struct dvs_rast_ctx dvs_rast_ctx;

typedef struct dvs_camera {
    float view[16];  /* world->view, flat: element [4*c + r] = row r, column c */
    float proj[16];  /* full view-projection, same layout */
    float campos[3];
    float tanfovx, tanfovy;
    int32_t width, height;
    float bg[3];
    float scale_modifier;
    int32_t sh_degree;     /* active degree 0..3 */
    int32_t sh_rest_alloc; /* rest coefficients per Gaussian allocated in shN (>= (deg+1)^2 - 1) */
    uint32_t flags;        /* DVS_FLAG_* */
} dvs_camera;

typedef struct dvs_params { /* device pointers, caller-owned, read-only */
    const float* means3D;
    const float* scales;
    const float* quats;
    const float* opacities;
    const float* sh0;
    const float* shN; /* may be NULL when sh_rest_alloc == 0 */
} dvs_params;

typedef struct dvs_grads { /* device pointers, caller-owned, same shapes as dvs_params */
    float* means3D;
    float* scales;
    float* quats;
    float* opacities;
    float* sh0;
    float* shN;
    float* mean2D_abs; /* [N,2] optional (DVS_FLAG_ABSGRAD), may be NULL */
    float* mean2D;     /* [N,2] optional: screen-space dL/dmean2D (ndc-scaled), may be NULL */
} dvs_grads;

typedef struct dvs_stats {
    int64_t num_gaussians; /* N of the last forward */
    int64_t num_visible;   /* V: radius > 0 */
    int64_t num_dups;      /* D: sum of tiles_touched */
    int64_t dup_capacity;  /* entries the binning arena can hold */
    int64_t max_tile_len;  /* longest tile list */
    int32_t tiles_x, tiles_y;
    int32_t overflow;      /* 1 if the last forward needed a bigger arena (it was re-run) */
    int32_t reserved_;
    int64_t num_list_entries; /* entries actually in the tile lists (= num_dups unless DVS_FLAG_TIGHT_LISTS) */
} dvs_stats;

/* ids of internal buffers readable through dvs_rast_debug_read (parity tests only) */
enum {
    DVS_BUF_RADII = 0,         /* int32 [N] */
    DVS_BUF_TILES_TOUCHED = 1, /* uint32 [N] */
    DVS_BUF_DEPTH = 2,         /* float [N] */
    DVS_BUF_MEAN2D = 3,        /* float [N,2] */
    DVS_BUF_CONIC_OPACITY = 4, /* float [N,4]: A,B,C,opacity (un-scaled back from the record) */
    DVS_BUF_RGB = 5,           /* float [N,3] */
    DVS_BUF_CLAMPED = 6,       /* uint8 [N,3] */
    DVS_BUF_POINT_LIST = 7,    /* uint32 [D] sorted Gaussian ids, tile-major */
    DVS_BUF_RANGES = 8,        /* uint32 [T,2] */
    DVS_BUF_FINAL_T = 9,       /* float [H*W] */
    DVS_BUF_N_CONTRIB = 10,    /* uint32 [H*W] */
    DVS_BUF_CULL_MASK = 11,    /* uint8 [D] per-entry 8-bit sub-tile mask (ours; no upstream analogue) */
    DVS_BUF_SCREEN_GRADS = 12  /* float [N,12]: moments of s=dL/dpower {Sx,Sy,Sxx,Sxy,Syy,S0}, colour sums (3), |gx|,|gy|, pad */
};

/* Create a context on CUDA device `device`.  Scratch arenas grow on demand and persist. */
DVS_API int dvs_rast_create(int device, dvs_rast_ctx** out);
DVS_API void dvs_rast_destroy(dvs_rast_ctx* ctx);
DVS_API const char* dvs_rast_last_error(const dvs_rast_ctx* ctx);
DVS_API const char* dvs_rast_version(void);

/* Pre-size the arenas (optional; avoids the grow-and-rerun path and any allocation in the step). */
DVS_API int dvs_rast_reserve(dvs_rast_ctx* ctx, int64_t max_gaussians, int32_t max_width, int32_t max_height,
                             int64_t dup_capacity);

/*
 * Forward: preprocess (A1) -> tile binning (A2,A3) -> tile-local sort (A4,A5) -> compositing (A6).
 *   out_color : float [3,H,W] planar
 *   out_radii : int32 [N] or NULL
 * Replaces the forward half of the absent gsplatrast operator.
 * Synchronises `stream` once at the end to validate the binning arena size (the credited
 * upstream synchronises mid-pipeline to read D back); on overflow the arena is grown and the
 * forward re-run transparently.  With DVS_FLAG_DEFER_CHECK in cam->flags the call returns without any
 * host synchronisation (a training loop keeps the GPU queue full); an overflow is then reported as
 * DVS_E_OVERFLOW by a later call.
 */
DVS_API int dvs_rast_forward(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                             float* out_color, int32_t* out_radii, void* stream);

/*
 * Backward of the last forward on this context: reverse-walk compositing gradients (A7) ->
 * per-Gaussian backward (A8) to the stored parameters.
 *   dL_dpix : float [3,H,W]
 * `flags`: DVS_FLAG_ACCUMULATE, DVS_FLAG_ABSGRAD.
 */
DVS_API int dvs_rast_backward(dvs_rast_ctx* ctx, const dvs_params* params, const float* dL_dpix,
                              const dvs_grads* grads, uint32_t flags, void* stream);

/*
 * Auxiliary maps of the last forward (what depth- / normal-consistency losses and mesh extraction
 * read).  Either output may be NULL.
 *   out_aux    : float [2,H,W] = { depth = sum_k w_k z_k (view-space z, not normalised), alpha = sum_k w_k = 1 - T }
 *   out_normal : float [3,H,W] = sum_k w_k n_k, n_k = the shortest axis of Gaussian k's ellipsoid in view space, turned
 *                towards the camera (needs `params`: the tensors of the forward)
 * Computed by linearity with the same compositing kernel on records whose colour is (z, 1, 0) resp. n_k, over a zero
 * background; no host synchronisation.
 */
DVS_API int dvs_rast_forward_aux(dvs_rast_ctx* ctx, const dvs_params* params, float* out_aux, float* out_normal, void* stream);

/*
 * Backward of  <image, dL_dpix> + <depth, dL_daux[0]> + <alpha, dL_daux[1]> + <normal map, dL_dnormal>  (use instead of
 * dvs_rast_backward when the loss also reads the auxiliary maps).  dL_dpix : float [3,H,W]; dL_daux : float [2,H,W] or
 * NULL; dL_dnormal : float [3,H,W] or NULL; flags as dvs_rast_backward.
 */
DVS_API int dvs_rast_backward_aux(dvs_rast_ctx* ctx, const dvs_params* params, const float* dL_dpix, const float* dL_daux,
                                  const float* dL_dnormal, const dvs_grads* grads, uint32_t flags, void* stream);

/*
 * Background model (GaussianTrainConfig::enableBg, "Create Sky Model" docs/userGuide.md:53): a caller-owned per-pixel
 * background image [3,H,W] (device) replaces the constant camera background in every later forward / backward of this
 * context — out = C + final_T * bg(pixel) — until it is reset with NULL.  The image must match the camera's size and stay
 * valid from the forward to its backward.  dvs_rast_background_grad writes dL/dbg = final_T * dL/dpix [3,H,W] of the last
 * forward (what the caller's sky model differentiates through).
 */
DVS_API int dvs_rast_set_background(dvs_rast_ctx* ctx, const float* bg_image);
DVS_API int dvs_rast_background_grad(dvs_rast_ctx* ctx, const float* dL_dpix, float* dL_dbg, void* stream);

/*
 * Host-buffer step (the end-to-end path a trainer without device-resident images uses):
 * copies dL_dpix_host (pinned or pageable) to the device, runs forward + backward with the
 * device-resident parameters/gradients, copies the rendered image back to out_color_host.
 */
DVS_API int dvs_rast_step_host(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                               const dvs_grads* grads, const float* dL_dpix_host, float* out_color_host,
                               uint32_t bwd_flags, void* stream);

/*
 * The same step, PIPELINED: dvs_rast_step_host_async only queues the work of one step on `slot` (0 or 1: its own device
 * staging buffers and events) and returns; dvs_rast_step_host_wait(slot) blocks until that step's image has arrived in its
 * out_color_host.  A trainer alternates the slots and waits for step k only after queueing step k+1, so step k+1's H2D
 * overlaps step k's backward, step k's D2H overlaps step k+1's forward, and the launch queue never drains:
 *     async(slot 0); async(slot 1); wait(0); async(slot 0); wait(1); ...
 * The host buffers of a slot must stay untouched between its async and its wait; whatever else the caller queues on
 * `stream` after the call (the multi-GPU gradient exchange, the optimiser) runs after the step's backward with no host
 * synchronisation in between.  Use DVS_FLAG_DEFER_CHECK in cam->flags (a synchronous forward would stall the pipeline);
 * wait() reports an overflowed deferred-check forward (DVS_E_OVERFLOW) as soon as it is known.
 */
DVS_API int dvs_rast_step_host_async(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                                     const dvs_grads* grads, const float* dL_dpix_host, float* out_color_host,
                                     uint32_t bwd_flags, int slot, void* stream);
DVS_API int dvs_rast_step_host_wait(dvs_rast_ctx* ctx, int slot);

DVS_API int dvs_rast_get_stats(dvs_rast_ctx* ctx, dvs_stats* out);

/* DEVICE address of the word the last forward set non-zero if its binning arena overflowed (the compositing and backward
 * kernels of that step then exit early and produce nothing).  With DVS_FLAG_DEFER_CHECK the host learns of it one call
 * later (DVS_E_OVERFLOW, redo the step); kernels the caller queues behind the step (optimiser, statistics) read this word
 * and skip their update, so a step that produced no gradients does not move the model. */
DVS_API const uint32_t* dvs_rast_device_overflow_word(const dvs_rast_ctx* ctx);

/* Number of CUDA kernels of this library enqueued through the context so far (forward, backward, auxiliary passes; copies and
 * memsets are not kernels).  bench.py reports the difference over its timed region as `gpu_launches`. */
DVS_API uint64_t dvs_rast_kernel_launches(const dvs_rast_ctx* ctx);

/* Per-stage CUDA events (dvs_rast_stage_ms) are recorded only while profiling is on (default: on).  A training loop
 * turns it off: nine event records per step are launch-queue work the step does not need. */
DVS_API int dvs_rast_set_profiling(dvs_rast_ctx* ctx, int on);

/* Copy an internal buffer of the last forward/backward to HOST memory (parity tests). */
DVS_API int dvs_rast_debug_read(dvs_rast_ctx* ctx, int which, void* dst_host, size_t dst_bytes);

/* Per-stage device time of the last forward/backward in milliseconds (CUDA events; syncs). */
#define DVS_NUM_STAGES 8
DVS_API int dvs_rast_stage_ms(dvs_rast_ctx* ctx, float out_ms[DVS_NUM_STAGES]);
DVS_API const char* dvs_rast_stage_name(int i);

/*
 * NVSwitch in-switch all-reduce (sum, fp32, in place) of a symmetric buffer through its multicast mapping
 * (multimem.ld_reduce + multimem.st, two-shot: rank r reduces and re-broadcasts shard r).  `multicast_ptr` is the
 * NVLS multicast address of the buffer (e.g. torch.distributed._symmetric_memory handle.multicast_ptr), numel_f32 a
 * multiple of 4.  The caller must bracket the call with cross-rank barriers on `stream`.  New functionality: the
 * reference has no multi-GPU path; this is the exchange step of SURVEY.md section 8(e).
 */
DVS_API int dvs_coll_allreduce_nvls(void* multicast_ptr, size_t numel_f32, int rank, int world, int ctas,
                                    void* stream);

/*
 * Local half of the FACTORED gradient exchange: per view the gradient of the higher SH bands is the outer product
 * B(dir) (x) dL/dcolour, and the band-0 gradient the rasterizer writes is SH_C0 * dL/dcolour, so a rank that holds every
 * view's dL/dsh0 (all-gathered, 12 B per Gaussian and view) and camera centre forms the summed dL/dshN itself instead of
 * all-reducing that 180 B per Gaussian tensor:
 *   out_dshN[i][k][c] = sum_v B_{k+1}(normalize(means[i] - campos[v])) * dsh0_all[v][i][c] / SH_C0
 * means [N,3], dsh0_all [V,N,3], out_dshN [N,sh_rest_alloc,3]: device; campos_all_host [V,3]: HOST; V <= 64.
 * Valid when every view's dL/dsh0 is kept apart (one view per rank and step, or one slice per view).
 */
DVS_API int dvs_coll_sh_grad_from_dsh0(const float* means, const float* campos_all_host, const float* dsh0_all, int64_t N,
                                       int num_views, int sh_degree, int sh_rest_alloc, float* out_dshN, void* stream);

/*
 * The whole multi-GPU gradient exchange of SURVEY.md section 8(e) as ONE kernel over NVSwitch peer and multicast (NVLS)
 * mappings — in-switch reduction, peer reads, device-side cross-rank barriers and the local SH accumulation fused (one view
 * per rank and step).  Per view the gradient of the higher SH bands is the outer product B(dir) (x) dL/dcolour and the band-0
 * gradient the rasterizer writes is SH_C0 * dL/dcolour, so instead of summing the 180 B per Gaussian of dL/dshN over ranks
 * every rank forms that sum itself from the 12 B per Gaussian of each view's dL/dsh0:
 *   1. barrier: all ranks' backward passes are complete (multimem.red on a symmetric counter word + acquire polling of
 *               the local replica; no host involvement);
 *   2. the first `reduce_ctas` CTAs sum quats | means, scales | opacities (three ranges of the arena, 44 B per Gaussian) in
 *               the switch: rank r reduces shard r with multimem.ld_reduce and re-broadcasts it with multimem.st;
 *   3. all CTAs (those first ones as soon as their requests are issued) read every view's dL/dsh0 row straight from that
 *               rank's arena over NVLink (peer loads, L1 bypassed) and form
 *                   dL/dshN[i] = sum_v B(dir_{v,i}) (x) dL/dsh0_v[i] / SH_C0      and      dL/dsh0[i] = sum_v dL/dsh0_v[i]
 *               (same order on every rank: bit-identical results) — dL/dshN goes into the local arena, the summed dL/dsh0
 *               into `sh0_tmp` because the peers are still reading this rank's per-view dL/dsh0 (HBM-bound, tiles handed out
 *               by an atomic counter; the NVLink-bound and the HBM-bound work overlap);
 *   4. barrier: every shard has been re-broadcast and every peer read is done; `sh0_tmp` is copied into place.
 * Bytes received per GPU and Gaussian: 12 (W - 1) + 44 (1 + 1/W) (8 ranks: 134) instead of 236 (1 + 1/W) = 266 for the plain
 * in-switch all-reduce.
 * All pointers are device addresses of this rank.  `arena_mc` is the multicast address, `arena_local` this rank's unicast
 * address and `arena_peers[r]` rank r's arena as mapped into this process (torch.distributed._symmetric_memory:
 * multicast_ptr / buffer_ptrs, or cuMulticast* / cuMemMap); `sh0_tmp`: 3 N floats of plain device memory.  The words behind
 * `signal_*` and `grid_counter` (TWO words: grid arrivals, tile counter) must be zero before the FIRST call and are owned by
 * the kernel afterwards; `launch_index` is 0, 1, 2, ... and every rank must make the same sequence of calls.  `status`
 * (device, may be NULL) is set non-zero if a barrier timed out (~2 s): the kernel then ends without hanging and the results
 * are invalid.  Offsets are in floats from the start of the arena, multiples of 4.
 */
typedef struct dvs_coll_fused {
    void* arena_mc;
    float* arena_local;
    const float* arena_peers[16];
    float* sh0_tmp;
    uint32_t* signal_mc;
    uint32_t* signal_local;
    uint32_t* grid_counter;
    uint32_t* status;
    const float* means;         /* [N,3] parameters (view directions) */
    float campos[16 * 3];       /* camera centre of every rank's view */
    int64_t N;
    int64_t off_sh0, off_shN;   /* dL/dsh0 [N,3] (read from every rank, summed), dL/dshN [N,sh_rest_alloc,3] (written locally) */
    int64_t ranges[3][2];       /* [begin, end) of the ranges reduced in the switch (empty ranges allowed) */
    uint64_t launch_index;
    int32_t rank, world, sh_degree, sh_rest_alloc;
    int32_t ctas, reduce_ctas;  /* <= 0: defaults (one CTA per SM; a sixth of them issue the in-switch reduction first) */
} dvs_coll_fused;
DVS_API int dvs_coll_exchange_fused(const dvs_coll_fused* args, void* stream);
/* grid size dvs_coll_exchange_fused will use on the current device for `ctas` and `world` ranks (the co-residency bound applied) */
DVS_API int dvs_coll_exchange_fused_grid(int ctas, int world);

#ifdef __cplusplus
}
#endif
#endif
