/*
 * dvs_model_io.h — C-ABI of the Gaussian-model writers/readers (the formats the rest of
 * DIVSHOT consumes what the trainer produces in).  Implemented in divshot_b200/csrc/model_io.cpp, exported by
 * libgstrain.so.  Host pointers only; no CUDA, no torch.
 *
 * Reference interfaces replaced (all in /root/reference):
 *   external/tinygsplat/tiny_gsplat.hpp:600-687   save_ply / save_splat / save_compress_ply / save_dvs_splat /
 *                                                 save_spz_splats and the matching load_* functions
 *   external/tinygsplat/tiny_gsplat.cpp:168-241   PLY            :243-291  .splat      :293-395  compressed PLY
 *                                      :994-1117  .dvsplat       :1243-1272 .spz (-> external/spz/src/load-spz.cc)
 *                                      :398-630   reduced PLY (float rows) and its reader :817-992
 *   diverse/source/assets/gaussian_model.cpp:439-463   dispatch by file extension (mirrored by DVS_FMT_AUTO)
 *
 * Parity: byte-identical files to the reference writers and value-identical rows to the reference readers, checked
 * against the REAL reference code compiled into oracle/_ref/libtinygsplat_ref.so (tests/test_model_io.py).
 * Documented quirks of the reference that are reproduced for byte parity are listed in DESIGN.md §7.
 */
#ifndef DVS_MODEL_IO_H
#define DVS_MODEL_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DVS_API
#define DVS_API __attribute__((visibility("default")))
#endif

typedef enum dvs_model_format {
    DVS_FMT_AUTO = 0,           /* by path: ".compressed" + .ply -> 3, ".reduced" + .ply -> 6, .ply -> 1, .splat -> 2,
                                   .dvsplat -> 4, .spz -> 5 */
    DVS_FMT_PLY = 1,            /* 59 raw floats per vertex: x y z f_dc_0..2 f_rest_0..44 (channel-major) opacity scale_0..2 rot_0..3 */
    DVS_FMT_SPLAT = 2,          /* 32-byte records: pos f32x3, exp(scale) f32x3, RGBA u8x4, quaternion u8x4 */
    DVS_FMT_COMPRESSED_PLY = 3, /* 256-splat chunks in Morton order: 12 bound floats per chunk + 4 packed u32 per splat */
    DVS_FMT_DVSPLAT = 4,        /* 28-byte header, chunked 11-10-11 positions, u8-quantised attributes per SH degree block */
    DVS_FMT_SPZ = 5,            /* Niantic .spz v3 (gzip): 24-bit fixed-point positions, smallest-three quaternions */
    DVS_FMT_REDUCED_PLY = 6     /* ".reduced" + .ply: four vertex elements, one per SH degree, each row holding only the
                                   coefficients its degree uses (float rows: what the reference's dispatch writes) */
} dvs_model_format;

#define DVS_IO_ANTIALIASED 1u   /* model was trained with mip anti-aliasing (header comment / spz flag bit) */
#define DVS_IO_SPZ_SH_FIXED 2u  /* write: index the .spz SH block as [p][15][3] instead of the reference's overlapping
                                   [p*15+j+c] (tiny_gsplat.cpp:1262-1267).  Default (flag clear) = reference bytes. */

#define DVS_IO_REDUCED_SH_FIXED 4u /* write: coefficient j of a reduced-PLY row = shN[j][0..2]; default (flag clear) = the
                                   reference's bytes, which copy the overlapping window shN_flat[j..j+2]
                                   (tiny_gsplat.cpp:533-534) */

#define DVS_IO_ROW_FLOATS 59    /* reader row = the reference's RichPoint: pos[3] shs[48] opacity scale[3] rot[4] */

/* Resolves DVS_FMT_AUTO for `path`; returns 0 if the extension is not a model format. */
DVS_API int dvs_model_format_from_path(const char* path);

/* Writes N Gaussians given in the trainer's tensor layouts (raw / un-activated parameters):
 *   means3D[N,3], sh0[N,3], shN[N,15,3] (coefficient-major, RGB interleaved), logit_opac[N], log_scales[N,3],
 *   quats[N,4] (w,x,y,z), degrees[N] (active SH degree per Gaussian; NULL = all 3; used by .dvsplat only).
 * Returns 0, or a negative value (message via dvs_model_io_last_error). */
DVS_API int dvs_model_write(const char* path, int format, int64_t N, const float* means3D, const float* sh0,
                            const float* shN, const float* logit_opac, const float* log_scales, const float* quats,
                            const uint8_t* degrees, uint32_t flags);

/* Reads a model file.  Returns the number of Gaussians in the file (negative on error) and fills
 * rows[min(count, cap)][DVS_IO_ROW_FLOATS]; rows may be NULL to query the count.  *flags_out gets
 * DVS_IO_ANTIALIASED when the file says so. */
DVS_API int64_t dvs_model_read(const char* path, int format, float* rows, int64_t cap, uint32_t* flags_out);

DVS_API const char* dvs_model_io_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
