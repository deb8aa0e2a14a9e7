// api.cu — the C-ABI of include/dvs_rast.h: context, persistent arenas, stage sequencing.
//
// One context owns every scratch arena the rasterizer needs (screen records, tile bins, sorted lists,
// per-pixel compositing state, screen-gradient records) so a training step allocates nothing.  The
// credited upstream resizes byte tensors through allocator callbacks and reads the duplicate count D
// back to the host in the middle of the pipeline; here D stays on the device —
// the arena is sized ahead and validated once, after the forward has been enqueued.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "dvs_rast.h"
#include "kernels.h"

using namespace dvs;

// Launch order of the compositing CTAs: raster order.  -DDVS_TILE_ORDER_LPT makes the tile scan write a longest-list-first
// order and the compositors follow it; measured at c3 it changes nothing (profiles/r2_ab1_loops_tight_lpt.json: 0.8855 ms
// without vs 0.8889 ms with), so it is not in the default build.
#ifdef DVS_TILE_ORDER_LPT
#define DVS_TILE_ORDER(ctx) (ctx)->tile_order
#else
#define DVS_TILE_ORDER(ctx) nullptr
#endif

struct dvs_rast_ctx {
    int device = 0;
    char err[512] = {0};
    unsigned long long kernel_launches = 0;  // kernels of ours enqueued through this context so far
    // per-Gaussian arenas
    int64_t cap_gauss = 0;
    float4* rec = nullptr;
    uint4* aux = nullptr;
    float4* sgrad = nullptr;
    // 2DGS (DVS_FLAG_MODEL_2DGS): 64-byte homography records and screen-gradient records, allocated on first use
    int64_t cap_surfel = 0;
    float4* rec2 = nullptr;
    float4* sgrad2 = nullptr;
    float4* cull2 = nullptr;  // [cap][2] sub-tile cull ellipses of the surfels
    bool surfel_fwd = false;  // the last forward was a 2DGS one (the backward follows it)
    // per-tile
    int64_t cap_tiles = 0;
    uint32_t* tile_count = nullptr;
    uint32_t* tile_base = nullptr;  // T+1
    uint32_t* tile_cursor = nullptr;
    uint32_t* class_tiles = nullptr;  // [5][T] per-sort-class tile lists
    uint32_t* tile_order = nullptr;   // [T] launch order of the compositing CTAs (longest lists first)
    // per-duplicate
    int64_t cap_dups = 0;   // entries plist (and, in two-pass mode, bins) can hold
    int64_t cap_bins = 0;   // entries of the bins arena (>= cap_dups; >= T * bin_stride in single-pass mode)
    unsigned long long* bins = nullptr;
    uint32_t* plist = nullptr;
    // single-pass binning (only with DVS_FLAG_DEFER_CHECK): fixed per-tile bin stride sized by the last synchronous forward
    uint32_t bin_stride = 0;
    int64_t bin_stride_tiles = 0;
    // per-pixel
    int64_t cap_pix = 0;
    float* final_T = nullptr;
    uint32_t* n_contrib = nullptr;
    const float* bg_image = nullptr;  // caller-owned per-pixel background [3,H,W] (dvs_rast_set_background), or null
    float* h2d_grad[2] = {nullptr, nullptr};   // [3P] device staging of dL/dpix for dvs_rast_step_host*, one per pipeline slot
    float* d_image[2] = {nullptr, nullptr};    // [3P] rendered image before its D2H, one per pipeline slot
    // F4 auxiliary outputs (dvs_rast_forward_aux / dvs_rast_backward_aux), allocated on first use
    int64_t cap_aux_gauss = 0, cap_aux_pix = 0;
    float4* rec_aux = nullptr;   // [3 cap] records with the colour replaced by (depth, 1, 0)
    float* aux_dz = nullptr;     // [cap] dL/d(view-space depth) per Gaussian
    float* aux_dn = nullptr;     // [3 cap] dL/d(view-space normal) per Gaussian
    float* aux_img = nullptr;    // [3P] staging: the compositing kernels work on three planes
    float* aux_T = nullptr;      // [P]
    uint32_t* aux_nc = nullptr;  // [P]
    // small device words + pinned mirror
    uint32_t* info = nullptr;               // [16]: D, max len, overflow, -, tiles per sort class [5]
    unsigned long long* stats = nullptr;    // [2]: V, D
    uint32_t* h_info = nullptr;             // pinned [16]
    unsigned long long* h_stats = nullptr;  // pinned [2]
    // last forward
    bool have_fwd = false;
    Cam cam{};
    int64_t N = 0;
    dvs_stats st{};
    cudaEvent_t ev[DVS_NUM_STAGES + 2] = {};
    // dvs_rast_step_host: copies run on their own stream so the H2D overlaps the forward and the D2H the backward
    cudaStream_t copy_stream = nullptr;      // host -> device (dL/dpix)
    cudaStream_t copy_stream_out = nullptr;  // device -> host (image): its own stream, so a step's H2D is not queued behind the
                                             // previous step's D2H (PCIe is full duplex)
    cudaEvent_t ev_img = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr}, ev_bwd_done[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    bool ev_fwd = false, ev_bwd = false;
    // deferred arena validation (DVS_FLAG_DEFER_CHECK)
    cudaEvent_t ev_check = nullptr;
    bool pending_check = false;
    bool arena_sized = false;      // a synchronous forward has validated the capacity
    // the tile scan leaves the tile counters, the V / D accumulators and the bin-overflow word zeroed for the next forward;
    // anything that breaks that hand-over (first use, two-pass cursors, an error between launches) sets these and the
    // next forward clears them with memsets
    bool count_dirty = true, cursor_dirty = true, words_dirty = true;
    bool profiling = true;         // record the per-stage CUDA events (dvs_rast_set_profiling)
    uint32_t seen_overflows = 0;   // value of the sticky device counter info[3] already handled
};

static int fail(dvs_rast_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof c->err, fmt, ap);
        va_end(ap);
    }
    return code;
}
// (every CK(launch_*(...)) is one kernel launch of ours — launch_tile_sort adds its extra kernels itself — counted for
// dvs_rast_kernel_launches)
#define CK(call)                                                                                           \
    do {                                                                                                   \
        if (__builtin_strncmp(#call, "launch_", 7) == 0) ctx->kernel_launches++;                           \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail(ctx, DVS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// per-stage timing events (only while profiling is on: dvs_rast_set_profiling)
#define EV(i)                                                    \
    do {                                                         \
        if (ctx->profiling) CK(cudaEventRecord(ctx->ev[i], st)); \
    } while (0)

template <typename T>
static cudaError_t regrow(T*& p, size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
}

static int ensure_gauss(dvs_rast_ctx* ctx, int64_t N) {
    if (N <= ctx->cap_gauss) return DVS_OK;
    const int64_t cap = N + N / 8 + 1024;
    CK(regrow(ctx->rec, 3 * (size_t)cap));
    CK(regrow(ctx->aux, (size_t)cap));
    CK(regrow(ctx->sgrad, 3 * (size_t)cap));
    CK(cudaMemset(ctx->sgrad, 0, 3 * (size_t)cap * sizeof(float4)));
    ctx->cap_gauss = cap;
    return DVS_OK;
}
static int ensure_tiles(dvs_rast_ctx* ctx, int64_t T) {
    if (T <= ctx->cap_tiles) return DVS_OK;
    CK(regrow(ctx->tile_count, (size_t)T * TILE_CTR_STRIDE));
    CK(regrow(ctx->tile_base, (size_t)T + 1));
    CK(regrow(ctx->tile_cursor, (size_t)T * TILE_CTR_STRIDE));
    CK(regrow(ctx->class_tiles, 5 * (size_t)T));
    CK(regrow(ctx->tile_order, (size_t)T));
    ctx->cap_tiles = T;
    ctx->count_dirty = ctx->cursor_dirty = true;
    return DVS_OK;
}
static int ensure_dups(dvs_rast_ctx* ctx, int64_t D) {
    if (D <= ctx->cap_dups) return DVS_OK;
    if (D >= (int64_t)0xffffffffll) return fail(ctx, DVS_E_UNSUPPORTED, "duplicate count %lld exceeds 2^32", (long long)D);
    if (D > ctx->cap_bins) {
        CK(regrow(ctx->bins, (size_t)D));
        ctx->cap_bins = D;
    }
    CK(regrow(ctx->plist, (size_t)D));
    ctx->cap_dups = D;
    return DVS_OK;
}
static int ensure_bins(dvs_rast_ctx* ctx, int64_t entries) {
    if (entries <= ctx->cap_bins) return DVS_OK;
    CK(regrow(ctx->bins, (size_t)entries));
    ctx->cap_bins = entries;
    return DVS_OK;
}
static int ensure_pix(dvs_rast_ctx* ctx, int64_t P) {
    if (P <= ctx->cap_pix) return DVS_OK;
    CK(regrow(ctx->final_T, (size_t)P));
    CK(regrow(ctx->n_contrib, (size_t)P));
    for (int k = 0; k < 2; k++) {
        if (ctx->h2d_grad[k]) { cudaFree(ctx->h2d_grad[k]); ctx->h2d_grad[k] = nullptr; }
        if (ctx->d_image[k]) { cudaFree(ctx->d_image[k]); ctx->d_image[k] = nullptr; }
    }
    ctx->cap_pix = P;
    return DVS_OK;
}

static inline uint64_t info_v(const dvs_rast_ctx* ctx) { return (uint64_t)ctx->h_info[12] | ((uint64_t)ctx->h_info[13] << 32); }
static inline uint64_t info_d(const dvs_rast_ctx* ctx) { return (uint64_t)ctx->h_info[14] | ((uint64_t)ctx->h_info[15] << 32); }

static void publish_stats(dvs_rast_ctx* ctx) {
    ctx->st.num_visible = (int64_t)info_v(ctx);
    ctx->st.num_list_entries = (int64_t)ctx->h_info[0];
    // D of SURVEY.md section 8(d) = sum of tiles_touched, whatever the lists hold (with DVS_FLAG_TIGHT_LISTS they hold fewer)
    ctx->st.num_dups = (int64_t)info_d(ctx);
    ctx->st.dup_capacity = ctx->cap_dups;
    ctx->st.max_tile_len = ctx->h_info[1];
}

// Deferred validation: if the last DEFER_CHECK forward has finished (or `block`), look at the sticky overflow
// counter; on overflow grow the arena and report DVS_E_OVERFLOW.
static int resolve_pending(dvs_rast_ctx* ctx, bool block) {
    if (!ctx->pending_check) return DVS_OK;
    if (block) {
        CK(cudaEventSynchronize(ctx->ev_check));
    } else {
        cudaError_t q = cudaEventQuery(ctx->ev_check);
        if (q == cudaErrorNotReady) return DVS_OK;
        CK(q);
    }
    ctx->pending_check = false;
    if (ctx->h_info[3] != ctx->seen_overflows) {
        ctx->seen_overflows = ctx->h_info[3];
        const int64_t need = (int64_t)ctx->h_info[9];
        ctx->st.overflow = 1;
        ctx->have_fwd = false;
        ctx->arena_sized = false;
        int rc = ensure_dups(ctx, need + need / 4 + 4096);
        if (rc) return rc;
        return fail(ctx, DVS_E_OVERFLOW, "a deferred-check forward needed %lld binning entries (> capacity); arena grown, redo the step",
                    (long long)need);
    }
    publish_stats(ctx);
    return DVS_OK;
}

extern "C" {

const char* dvs_rast_version(void) { return "divshot_b200 rasterizer 0.1 (sm_100a)"; }

int dvs_rast_create(int device, dvs_rast_ctx** out) {
    if (!out) return DVS_E_INVALID;
    *out = nullptr;
    dvs_rast_ctx* ctx = new (std::nothrow) dvs_rast_ctx();
    if (!ctx) return DVS_E_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->info), 16 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->stats), 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_info), 16 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_stats), 2 * sizeof(unsigned long long));
    for (int i = 0; e == cudaSuccess && i < DVS_NUM_STAGES + 2; i++) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaMemset(ctx->info, 0, 16 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_check, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_img, cudaEventDisableTiming);
    for (int k = 0; e == cudaSuccess && k < 2; k++) {
        e = cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_d2h[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_bwd_done[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        // the product path must fail loudly without a usable CUDA device: there is no CPU fallback.
        fprintf(stderr, "dvs_rast_create: CUDA unavailable on device %d: %s\n", device, cudaGetErrorString(e));
        delete ctx;
        return DVS_E_CUDA;
    }
    *out = ctx;
    return DVS_OK;
}

void dvs_rast_destroy(dvs_rast_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->rec); cudaFree(ctx->aux); cudaFree(ctx->sgrad); cudaFree(ctx->rec2); cudaFree(ctx->sgrad2); cudaFree(ctx->cull2);
    cudaFree(ctx->tile_count); cudaFree(ctx->tile_base); cudaFree(ctx->tile_cursor); cudaFree(ctx->class_tiles); cudaFree(ctx->tile_order);
    cudaFree(ctx->bins); cudaFree(ctx->plist);
    cudaFree(ctx->final_T); cudaFree(ctx->n_contrib);
    for (int k = 0; k < 2; k++) { cudaFree(ctx->h2d_grad[k]); cudaFree(ctx->d_image[k]); }
    cudaFree(ctx->rec_aux); cudaFree(ctx->aux_dz); cudaFree(ctx->aux_dn); cudaFree(ctx->aux_img); cudaFree(ctx->aux_T); cudaFree(ctx->aux_nc);
    cudaFree(ctx->info); cudaFree(ctx->stats);
    cudaFreeHost(ctx->h_info); cudaFreeHost(ctx->h_stats);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_check) cudaEventDestroy(ctx->ev_check);
    if (ctx->ev_img) cudaEventDestroy(ctx->ev_img);
    for (int k = 0; k < 2; k++) {
        if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
        if (ctx->ev_d2h[k]) cudaEventDestroy(ctx->ev_d2h[k]);
        if (ctx->ev_bwd_done[k]) cudaEventDestroy(ctx->ev_bwd_done[k]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_stream_out) cudaStreamDestroy(ctx->copy_stream_out);
    delete ctx;
}

const char* dvs_rast_last_error(const dvs_rast_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int dvs_rast_reserve(dvs_rast_ctx* ctx, int64_t max_gaussians, int32_t max_width, int32_t max_height,
                     int64_t dup_capacity) {
    if (!ctx) return DVS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if (max_gaussians > 0 && (rc = ensure_gauss(ctx, max_gaussians))) return rc;
    if (max_width > 0 && max_height > 0) {
        const int64_t gx = (max_width + TILE - 1) / TILE, gy = (max_height + TILE - 1) / TILE;
        if ((rc = ensure_tiles(ctx, gx * gy))) return rc;
        if ((rc = ensure_pix(ctx, (int64_t)max_width * max_height))) return rc;
    }
    if (dup_capacity > 0 && (rc = ensure_dups(ctx, dup_capacity))) return rc;
    return DVS_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_params(dvs_rast_ctx* ctx, const dvs_params* p, int KR) {
    if (!p || !p->means3D || !p->scales || !p->quats || !p->opacities || !p->sh0 || (KR > 0 && !p->shN))
        return fail(ctx, DVS_E_INVALID, "null parameter pointer");
    if (!aligned16(p->means3D) || !aligned16(p->scales) || !aligned16(p->quats) || !aligned16(p->opacities) ||
        !aligned16(p->sh0) || (KR > 0 && !aligned16(p->shN)))
        return fail(ctx, DVS_E_INVALID, "parameter pointers must be 16-byte aligned");
    return DVS_OK;
}

int dvs_rast_forward(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                     float* out_color, int32_t* out_radii, void* stream) {
    if (!ctx) return DVS_E_INVALID;
    if (!cam || !out_color) return fail(ctx, DVS_E_INVALID, "null camera / output");
    if (N < 0 || N >= (int64_t)MAX_GAUSSIANS)
        return fail(ctx, DVS_E_UNSUPPORTED, "N=%lld outside [0, 2^24)", (long long)N);
    if (cam->width <= 0 || cam->height <= 0) return fail(ctx, DVS_E_INVALID, "bad image size");
    if (cam->sh_degree < 0 || cam->sh_degree > 3) return fail(ctx, DVS_E_INVALID, "sh_degree must be 0..3");
    const int K = (cam->sh_degree + 1) * (cam->sh_degree + 1);
    if (cam->sh_rest_alloc < K - 1) return fail(ctx, DVS_E_INVALID, "sh_rest_alloc < (deg+1)^2-1");
    int rc;
    if (N > 0 && (rc = check_params(ctx, params, cam->sh_rest_alloc))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    if ((rc = resolve_pending(ctx, false))) return rc;
    const bool defer = (cam->flags & DVS_FLAG_DEFER_CHECK) && ctx->arena_sized;

    Cam c{};
    memcpy(c.view, cam->view, sizeof c.view);
    memcpy(c.proj, cam->proj, sizeof c.proj);
    memcpy(c.campos, cam->campos, sizeof c.campos);
    c.tanfovx = cam->tanfovx; c.tanfovy = cam->tanfovy;
    c.W = cam->width; c.H = cam->height;
    memcpy(c.bg, cam->bg, sizeof c.bg);
    c.scale_modifier = cam->scale_modifier;
    c.deg = cam->sh_degree; c.KR = cam->sh_rest_alloc;
    c.flags = cam->flags;
    c.gx = (c.W + TILE - 1) / TILE; c.gy = (c.H + TILE - 1) / TILE;
    c.bg_image = ctx->bg_image;
    const int64_t T = (int64_t)c.gx * c.gy, P = (int64_t)c.W * c.H;
    if (T >= (1 << 24) || c.gx > 65535 || c.gy > 65535) return fail(ctx, DVS_E_UNSUPPORTED, "image too large");

    if ((rc = ensure_gauss(ctx, N > 0 ? N : 1))) return rc;
    if ((rc = ensure_tiles(ctx, T))) return rc;
    if ((rc = ensure_pix(ctx, P))) return rc;
    if (ctx->cap_dups == 0 && (rc = ensure_dups(ctx, (N > 0 ? 16 * N : 1) + 4096))) return rc;
    if (!ctx->arena_sized && ctx->st.num_dups > 0 && !ctx->pending_check &&
        (rc = ensure_dups(ctx, ctx->st.num_dups + ctx->st.num_dups / 2 + 4096)))
        return rc;
    if (!ctx->pending_check && ctx->bin_stride > 0 && ctx->bin_stride_tiles == T &&
        (rc = ensure_bins(ctx, (int64_t)ctx->bin_stride * T)))
        return rc;

    Params prm{};
    if (N > 0) prm = Params{params->means3D, params->scales, params->quats, params->opacities, params->sh0, params->shN};
    ctx->have_fwd = false;
    ctx->st.overflow = 0;
    const bool surfel = (cam->flags & DVS_FLAG_MODEL_2DGS) != 0;
    if (surfel && ctx->cap_gauss > ctx->cap_surfel) {
        CK(regrow(ctx->rec2, 4 * (size_t)ctx->cap_gauss));
        CK(regrow(ctx->sgrad2, 4 * (size_t)ctx->cap_gauss));
        CK(regrow(ctx->cull2, 2 * (size_t)ctx->cap_gauss));
        CK(cudaMemset(ctx->sgrad2, 0, 4 * (size_t)ctx->cap_gauss * sizeof(float4)));
        ctx->cap_surfel = ctx->cap_gauss;
    }
    // single-pass binning: only in deferred-check mode, with a bin stride sized by an earlier synchronous forward
    // of the same tile grid (an overflowing bin is reported like an arena overflow: DVS_E_OVERFLOW, redo the step)
    // (DVS_TWO_PASS=1 in the environment forces two-pass binning: A/B measurements only)
    static const bool force_two_pass = getenv("DVS_TWO_PASS") != nullptr;
    const bool fused = defer && ctx->bin_stride > 0 && ctx->bin_stride_tiles == T &&
                       (int64_t)ctx->bin_stride * T <= ctx->cap_bins && !force_two_pass && !surfel;
    for (int attempt = 0; attempt < 3; attempt++) {
        uint32_t* counters = fused ? ctx->tile_cursor : ctx->tile_count;
        if (fused ? ctx->cursor_dirty : ctx->count_dirty)
            CK(cudaMemsetAsync(counters, 0, (size_t)ctx->cap_tiles * TILE_CTR_STRIDE * sizeof(uint32_t), st));
        if (ctx->words_dirty) {
            CK(cudaMemsetAsync(ctx->stats, 0, 2 * sizeof(unsigned long long), st));
            CK(cudaMemsetAsync(ctx->info + 10, 0, sizeof(uint32_t), st));
        }
        ctx->count_dirty = ctx->cursor_dirty = ctx->words_dirty = true;  // until the scan below has been enqueued
        EV(0);
        const FusedEmit fe{ctx->tile_cursor, ctx->bins, fused ? ctx->bin_stride : 0u, ctx->info + 10,
                           (fused && (cam->flags & DVS_FLAG_TIGHT_LISTS)) ? 1u : 0u};
        if (surfel)
            CK(launch_surfel_preprocess_fwd(c, (int)N, prm, ctx->rec, ctx->rec2, ctx->cull2, ctx->aux, ctx->tile_count, out_radii, ctx->stats, st));
        else
            CK(launch_preprocess_fwd(c, (int)N, prm, ctx->rec, ctx->aux, ctx->tile_count, out_radii, ctx->stats, fe, st));
        EV(1);
        if (fused) {
            CK(launch_tile_scan((int)T, ctx->tile_cursor, ctx->tile_base, nullptr, ctx->info, (uint32_t)ctx->cap_dups,
                                ctx->class_tiles, DVS_TILE_ORDER(ctx), ctx->stats, st));
            ctx->cursor_dirty = false;  // zeroed as read
            ctx->count_dirty = false;   // untouched
            EV(2);
        } else {
            CK(launch_tile_scan((int)T, ctx->tile_count, ctx->tile_base, ctx->tile_cursor, ctx->info,
                                (uint32_t)ctx->cap_dups, ctx->class_tiles, DVS_TILE_ORDER(ctx), ctx->stats, st));
            ctx->count_dirty = false;   // zeroed as read; tile_cursor now holds the emission cursors (dirty)
            EV(2);
            CK(launch_emit(c, (int)N, ctx->aux, surfel ? ctx->cull2 : ctx->rec, surfel ? 2 : 3, ctx->tile_cursor, ctx->bins,
                           (uint32_t)ctx->cap_dups, st));
        }
        ctx->words_dirty = false;
        EV(3);
        CK(launch_tile_sort((int)T, fused ? ctx->bin_stride : 0u, ctx->tile_base, ctx->bins, ctx->plist, ctx->info,
                            ctx->class_tiles, st));
        ctx->kernel_launches += (unsigned long long)(tile_sort_launch_count(fused ? ctx->bin_stride : 0u) - 1);
        EV(4);
        if (surfel)
            CK(launch_surfel_render_fwd(c, ctx->tile_base, ctx->plist, ctx->rec2, out_color, ctx->final_T, ctx->n_contrib, ctx->info, st));
        else
            CK(launch_render_fwd(c, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec, out_color, ctx->final_T, ctx->n_contrib,
                                 ctx->info, st));
        EV(5);
        ctx->surfel_fwd = surfel;
        CK(cudaMemcpyAsync(ctx->h_info, ctx->info, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (defer) {  // no host synchronisation: validated later by resolve_pending()
            CK(cudaEventRecord(ctx->ev_check, st));
            ctx->pending_check = true;
            ctx->cam = c; ctx->N = N;
            ctx->st.num_gaussians = N; ctx->st.tiles_x = c.gx; ctx->st.tiles_y = c.gy;
            ctx->have_fwd = true; ctx->ev_fwd = ctx->profiling;
            return DVS_OK;
        }
        CK(cudaStreamSynchronize(st));
        ctx->seen_overflows = ctx->h_info[3];
        if (!ctx->h_info[2]) break;
        // arena too small: grow to the exact need (+25%) and run the forward again
        ctx->st.overflow = 1;
        const int64_t need = (int64_t)info_d(ctx);
        if ((rc = ensure_dups(ctx, need + need / 4 + 4096))) return rc;
        if (attempt == 2) return fail(ctx, DVS_E_NOMEM, "binning arena overflow persisted");
    }
    ctx->cam = c;
    ctx->N = N;
    ctx->st.num_gaussians = N;
    publish_stats(ctx);
    ctx->st.tiles_x = c.gx; ctx->st.tiles_y = c.gy;
    // size the fixed per-tile bin stride for single-pass binning of later deferred-check forwards
    {
        int64_t stride = ((int64_t)ctx->h_info[1] * 3 / 2 + 64 + 63) / 64 * 64;
        if (ctx->bin_stride_tiles == T) stride = std::max<int64_t>(stride, ctx->bin_stride);  // never shrink: views alternate
        const int64_t want = stride * T;
        if (want <= 4 * (int64_t)ctx->h_info[0] + (16 << 20)) {  // skewed scenes (one huge tile) stay two-pass
            ctx->bin_stride = (uint32_t)stride;
            ctx->bin_stride_tiles = T;
        } else {
            ctx->bin_stride = 0;
        }
    }
    // leave head-room so that slowly drifting parameters / other views do not overflow a deferred-check step
    if (ctx->cap_dups < (int64_t)ctx->h_info[0] + (int64_t)ctx->h_info[0] / 2) {
        // (arena contents are dead after the forward only if no backward follows; grow lazily at the next forward)
        ctx->arena_sized = false;
    } else {
        ctx->arena_sized = true;
    }
    ctx->have_fwd = true;
    ctx->ev_fwd = ctx->profiling;
    return DVS_OK;
}

int dvs_rast_backward(dvs_rast_ctx* ctx, const dvs_params* params, const float* dL_dpix, const dvs_grads* grads,
                      uint32_t flags, void* stream) {
    if (!ctx) return DVS_E_INVALID;
    {
        int rcp = resolve_pending(ctx, false);
        if (rcp) return rcp;
    }
    if (!ctx->have_fwd) return fail(ctx, DVS_E_STATE, "backward without a forward on this context");
    if (!dL_dpix || !grads) return fail(ctx, DVS_E_INVALID, "null dL_dpix / grads");
    const Cam& c = ctx->cam;
    const int64_t N = ctx->N;
    int rc;
    if (N > 0) {
        if ((rc = check_params(ctx, params, c.KR))) return rc;
        if (!grads->means3D || !grads->scales || !grads->quats || !grads->opacities || !grads->sh0 ||
            (c.KR > 0 && !grads->shN))
            return fail(ctx, DVS_E_INVALID, "null gradient pointer");
        if (!aligned16(grads->means3D) || !aligned16(grads->scales) || !aligned16(grads->quats) ||
            !aligned16(grads->opacities) || !aligned16(grads->sh0) || (c.KR > 0 && !aligned16(grads->shN)))
            return fail(ctx, DVS_E_INVALID, "gradient pointers must be 16-byte aligned");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    Params prm{};
    Grads g{};
    if (N > 0) {
        prm = Params{params->means3D, params->scales, params->quats, params->opacities, params->sh0, params->shN};
        g = Grads{grads->means3D, grads->scales, grads->quats, grads->opacities, grads->sh0, grads->shN,
                  (flags & DVS_FLAG_ABSGRAD) ? grads->mean2D_abs : nullptr, grads->mean2D};
    }
    EV(6);
    if (ctx->surfel_fwd) {  // the 2DGS forward's backward
        CK(launch_surfel_render_bwd(c, ctx->tile_base, ctx->plist, ctx->rec2, ctx->final_T, ctx->n_contrib, dL_dpix,
                                    reinterpret_cast<float*>(ctx->sgrad2), ctx->info, st));
        EV(7);
        CK(launch_surfel_preprocess_bwd(c, (int)N, prm, ctx->aux, ctx->sgrad2, g, flags, st));
        EV(8);
        ctx->ev_bwd = ctx->profiling;
        return DVS_OK;
    }
    CK(launch_render_bwd(c, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec, ctx->final_T, ctx->n_contrib, dL_dpix,
                         reinterpret_cast<float*>(ctx->sgrad), (flags & DVS_FLAG_ABSGRAD) && g.mean2D_abs,
                         ctx->info, st));
    EV(7);
    CK(launch_preprocess_bwd(c, (int)N, prm, ctx->aux, ctx->sgrad, g, flags, st));
    EV(8);
    ctx->ev_bwd = ctx->profiling;
    return DVS_OK;
}

// ---- F4: depth / alpha maps and their gradients by linearity (aux_outputs.cu explains the construction)
static int ensure_aux(dvs_rast_ctx* ctx, int64_t N, int64_t P) {
    if (N > ctx->cap_aux_gauss) {
        const int64_t cap = std::max<int64_t>(N, ctx->cap_gauss);
        CK(regrow(ctx->rec_aux, 3 * (size_t)cap));
        CK(regrow(ctx->aux_dz, (size_t)cap));
        CK(regrow(ctx->aux_dn, 3 * (size_t)cap));
        ctx->cap_aux_gauss = cap;
    }
    if (P > ctx->cap_aux_pix) {
        CK(regrow(ctx->aux_img, 3 * (size_t)P));
        CK(regrow(ctx->aux_T, (size_t)P));
        CK(regrow(ctx->aux_nc, (size_t)P));
        ctx->cap_aux_pix = P;
    }
    return DVS_OK;
}

int dvs_rast_forward_aux(dvs_rast_ctx* ctx, const dvs_params* params, float* out_aux, float* out_normal, void* stream) {
    if (!ctx) return DVS_E_INVALID;
    if (!out_aux && !out_normal) return fail(ctx, DVS_E_INVALID, "null outputs");
    if (!ctx->have_fwd) return fail(ctx, DVS_E_STATE, "forward_aux without a forward on this context");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    Cam c = ctx->cam;
    c.bg[0] = c.bg[1] = c.bg[2] = 0.0f;
    c.bg_image = nullptr;
    const int64_t N = ctx->N, P = (int64_t)c.W * c.H;
    int rc;
    if (out_normal && N > 0 && (rc = check_params(ctx, params, c.KR))) return rc;
    if ((rc = ensure_aux(ctx, N > 0 ? N : 1, P))) return rc;
    if (out_aux) {
        CK(launch_aux_records((int)N, ctx->rec, ctx->rec_aux, st));
        CK(launch_render_fwd(c, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec_aux, ctx->aux_img, ctx->aux_T, ctx->aux_nc, ctx->info, st));
        CK(cudaMemcpyAsync(out_aux, ctx->aux_img, 2 * (size_t)P * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (out_normal) {
        Params prm{};
        if (N > 0) prm = Params{params->means3D, params->scales, params->quats, params->opacities, params->sh0, params->shN};
        CK(launch_aux_normal_records(c, (int)N, prm, ctx->rec, ctx->rec_aux, st));
        CK(launch_render_fwd(c, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec_aux, out_normal, ctx->aux_T, ctx->aux_nc, ctx->info, st));
    }
    return DVS_OK;
}

int dvs_rast_backward_aux(dvs_rast_ctx* ctx, const dvs_params* params, const float* dL_dpix, const float* dL_daux,
                          const float* dL_dnormal, const dvs_grads* grads, uint32_t flags, void* stream) {
    if (!ctx) return DVS_E_INVALID;
    {
        int rcp = resolve_pending(ctx, false);
        if (rcp) return rcp;
    }
    if (!ctx->have_fwd) return fail(ctx, DVS_E_STATE, "backward without a forward on this context");
    if (!dL_dpix || !grads) return fail(ctx, DVS_E_INVALID, "null dL_dpix / grads");
    const Cam& c = ctx->cam;
    const int64_t N = ctx->N, P = (int64_t)c.W * c.H;
    int rc;
    if (N > 0) {
        if ((rc = check_params(ctx, params, c.KR))) return rc;
        if (!grads->means3D || !grads->scales || !grads->quats || !grads->opacities || !grads->sh0 ||
            (c.KR > 0 && !grads->shN))
            return fail(ctx, DVS_E_INVALID, "null gradient pointer");
        if (!aligned16(grads->means3D) || !aligned16(grads->scales) || !aligned16(grads->quats) ||
            !aligned16(grads->opacities) || !aligned16(grads->sh0) || (c.KR > 0 && !aligned16(grads->shN)))
            return fail(ctx, DVS_E_INVALID, "gradient pointers must be 16-byte aligned");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    if ((rc = ensure_aux(ctx, N > 0 ? N : 1, P))) return rc;
    Params prm{};
    Grads g{};
    if (N > 0) {
        prm = Params{params->means3D, params->scales, params->quats, params->opacities, params->sh0, params->shN};
        g = Grads{grads->means3D, grads->scales, grads->quats, grads->opacities, grads->sh0, grads->shN,
                  (flags & DVS_FLAG_ABSGRAD) ? grads->mean2D_abs : nullptr, grads->mean2D};
    }
    EV(6);
    Cam c0 = c;  // the auxiliary passes composite over a zero background
    c0.bg[0] = c0.bg[1] = c0.bg[2] = 0.0f;
    c0.bg_image = nullptr;
    if (dL_daux) {
        // 1. depth / alpha loss: same reverse walk, colour triple (depth, 1, 0), third plane of dL zero
        CK(launch_aux_records((int)N, ctx->rec, ctx->rec_aux, st));
        CK(cudaMemcpyAsync(ctx->aux_img, dL_daux, 2 * (size_t)P * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemsetAsync(ctx->aux_img + 2 * (size_t)P, 0, (size_t)P * sizeof(float), st));
        CK(launch_render_bwd(c0, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec_aux, ctx->final_T, ctx->n_contrib, ctx->aux_img,
                             reinterpret_cast<float*>(ctx->sgrad), false, ctx->info, st));
        // its colour sums are dL/dz: out of the records, so that the next pass finds slots 6-8 empty
        CK(launch_aux_extract((int)N, ctx->sgrad, ctx->aux_dz, st));
    }
    if (dL_dnormal) {
        // 2. normal-map loss: colour triple = view-space normal; its colour sums are dL/dn
        CK(launch_aux_normal_records(c0, (int)N, prm, ctx->rec, ctx->rec_aux, st));
        CK(launch_render_bwd(c0, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec_aux, ctx->final_T, ctx->n_contrib, dL_dnormal,
                             reinterpret_cast<float*>(ctx->sgrad), false, ctx->info, st));
        CK(launch_aux_extract3((int)N, ctx->sgrad, ctx->aux_dn, st));
    }
    // 3. the colour loss adds its geometry sums on top, then the per-Gaussian backward consumes the total
    CK(launch_render_bwd(c, DVS_TILE_ORDER(ctx), ctx->tile_base, ctx->plist, ctx->rec, ctx->final_T, ctx->n_contrib, dL_dpix,
                         reinterpret_cast<float*>(ctx->sgrad), (flags & DVS_FLAG_ABSGRAD) && g.mean2D_abs, ctx->info, st));
    EV(7);
    CK(launch_preprocess_bwd(c, (int)N, prm, ctx->aux, ctx->sgrad, g, flags, st));
    // 4. dL/dmean += (row 2 of the view matrix) * dL/dz ;  dL/dquat += d n / d quat ^T dL/dn
    if (N > 0 && dL_daux) {
        const float row2[3] = {c.view[2], c.view[6], c.view[10]};
        CK(launch_aux_depth_grad((int)N, ctx->aux_dz, row2, g.means3D, st));
    }
    if (N > 0 && dL_dnormal) CK(launch_aux_normal_grad(c, (int)N, prm, ctx->aux_dn, g.quats, st));
    EV(8);
    ctx->ev_bwd = ctx->profiling;
    return DVS_OK;
}

// One host-buffer step, QUEUED only (no host synchronisation): H2D of dL/dpix into the slot's device staging on the copy
// stream (overlaps the forward, and the previous step's backward: it only waits for the backward that last READ this slot's
// staging buffer), forward, D2H of the image on the copy stream (overlaps the backward), backward.
static int step_host_enqueue(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                             const dvs_grads* grads, const float* dL_dpix_host, float* out_color_host, uint32_t bwd_flags,
                             void* stream, int slot) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    const int64_t P = (int64_t)cam->width * cam->height;
    int rc;
    if (P > ctx->cap_pix) {  // the per-pixel arenas are about to be re-allocated: nothing of an earlier step may be in flight
        CK(cudaStreamSynchronize(st));
        CK(cudaStreamSynchronize(ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->copy_stream_out));
        ctx->slot_used[0] = ctx->slot_used[1] = false;
    }
    if ((rc = ensure_pix(ctx, P))) return rc;
    if (!ctx->h2d_grad[slot]) CK(cudaMalloc(reinterpret_cast<void**>(&ctx->h2d_grad[slot]), 3 * (size_t)ctx->cap_pix * sizeof(float)));
    if (!ctx->d_image[slot]) CK(cudaMalloc(reinterpret_cast<void**>(&ctx->d_image[slot]), 3 * (size_t)ctx->cap_pix * sizeof(float)));
    if (ctx->slot_used[slot]) {
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_bwd_done[slot], 0));  // the backward that read this staging buffer
        CK(cudaStreamWaitEvent(st, ctx->ev_d2h[slot], 0));                     // the D2H that read this image buffer
    } else {
        CK(cudaEventRecord(ctx->ev_img, st));  // first use: order the copy after whatever the caller queued before this step
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_img, 0));
    }
    CK(cudaMemcpyAsync(ctx->h2d_grad[slot], dL_dpix_host, 3 * (size_t)P * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_h2d[slot], ctx->copy_stream));
    if ((rc = dvs_rast_forward(ctx, cam, N, params, ctx->d_image[slot], nullptr, stream))) return rc;
    CK(cudaEventRecord(ctx->ev_img, st));
    CK(cudaStreamWaitEvent(ctx->copy_stream_out, ctx->ev_img, 0));
    CK(cudaMemcpyAsync(out_color_host, ctx->d_image[slot], 3 * (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream_out));
    CK(cudaEventRecord(ctx->ev_d2h[slot], ctx->copy_stream_out));
    CK(cudaStreamWaitEvent(st, ctx->ev_h2d[slot], 0));
    if ((rc = dvs_rast_backward(ctx, params, ctx->h2d_grad[slot], grads, bwd_flags, stream))) return rc;
    CK(cudaEventRecord(ctx->ev_bwd_done[slot], st));
    ctx->slot_used[slot] = true;
    return DVS_OK;
}

int dvs_rast_step_host(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                       const dvs_grads* grads, const float* dL_dpix_host, float* out_color_host, uint32_t bwd_flags,
                       void* stream) {
    if (!ctx) return DVS_E_INVALID;
    if (!cam || !dL_dpix_host || !out_color_host) return fail(ctx, DVS_E_INVALID, "null host buffer");
    int rc = step_host_enqueue(ctx, cam, N, params, grads, dL_dpix_host, out_color_host, bwd_flags, stream, 0);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaStreamWaitEvent(st, ctx->ev_d2h[0], 0));
    CK(cudaStreamSynchronize(st));
    return resolve_pending(ctx, true);  // DVS_E_OVERFLOW if a deferred-check forward overflowed (redo the step)
}

int dvs_rast_step_host_async(dvs_rast_ctx* ctx, const dvs_camera* cam, int64_t N, const dvs_params* params,
                             const dvs_grads* grads, const float* dL_dpix_host, float* out_color_host, uint32_t bwd_flags,
                             int slot, void* stream) {
    if (!ctx) return DVS_E_INVALID;
    if (!cam || !dL_dpix_host || !out_color_host) return fail(ctx, DVS_E_INVALID, "null host buffer");
    if (slot < 0 || slot > 1) return fail(ctx, DVS_E_INVALID, "slot must be 0 or 1");
    return step_host_enqueue(ctx, cam, N, params, grads, dL_dpix_host, out_color_host, bwd_flags, stream, slot);
}

int dvs_rast_step_host_wait(dvs_rast_ctx* ctx, int slot) {
    if (!ctx) return DVS_E_INVALID;
    if (slot < 0 || slot > 1) return fail(ctx, DVS_E_INVALID, "slot must be 0 or 1");
    if (!ctx->slot_used[slot]) return fail(ctx, DVS_E_STATE, "no step queued on slot %d", slot);
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev_d2h[slot]));  // the slot's image is in out_color_host
    return resolve_pending(ctx, false);           // report an overflowed deferred-check forward as soon as it is known
}

int dvs_rast_set_background(dvs_rast_ctx* ctx, const float* bg_image) {
    if (!ctx) return DVS_E_INVALID;
    ctx->bg_image = bg_image;
    return DVS_OK;
}

int dvs_rast_background_grad(dvs_rast_ctx* ctx, const float* dL_dpix, float* dL_dbg, void* stream) {
    if (!ctx) return DVS_E_INVALID;
    if (!dL_dpix || !dL_dbg) return fail(ctx, DVS_E_INVALID, "null dL_dpix / dL_dbg");
    if (!ctx->have_fwd) return fail(ctx, DVS_E_STATE, "background_grad without a forward on this context");
    CK(cudaSetDevice(ctx->device));
    CK(launch_background_grad((int64_t)ctx->cam.W * ctx->cam.H, ctx->final_T, dL_dpix, dL_dbg, ctx->info, static_cast<cudaStream_t>(stream)));
    return DVS_OK;
}

const uint32_t* dvs_rast_device_overflow_word(const dvs_rast_ctx* ctx) { return ctx ? ctx->info + 2 : nullptr; }
uint64_t dvs_rast_kernel_launches(const dvs_rast_ctx* ctx) { return ctx ? (uint64_t)ctx->kernel_launches : 0u; }

int dvs_rast_set_profiling(dvs_rast_ctx* ctx, int on) {
    if (!ctx) return DVS_E_INVALID;
    ctx->profiling = on != 0;
    if (!ctx->profiling) ctx->ev_fwd = ctx->ev_bwd = false;
    return DVS_OK;
}

int dvs_rast_get_stats(dvs_rast_ctx* ctx, dvs_stats* out) {
    if (!ctx || !out) return DVS_E_INVALID;
    int rc = resolve_pending(ctx, true);  // blocking: this is where a deferred check is finally settled
    *out = ctx->st;
    return rc;
}

int dvs_rast_debug_read(dvs_rast_ctx* ctx, int which, void* dst, size_t dst_bytes) {
    if (!ctx || !dst) return DVS_E_INVALID;
    {
        int rcp = resolve_pending(ctx, true);
        if (rcp) return rcp;
    }
    if (!ctx->have_fwd) return fail(ctx, DVS_E_STATE, "no forward to read from");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    const size_t N = (size_t)ctx->N, D = (size_t)ctx->st.num_list_entries;
    const size_t T = (size_t)ctx->cam.gx * ctx->cam.gy, P = (size_t)ctx->cam.W * ctx->cam.H;
    auto need = [&](size_t b) -> int {
        return dst_bytes >= b ? DVS_OK : fail(ctx, DVS_E_INVALID, "debug_read: need %zu bytes, got %zu", b, dst_bytes);
    };
    int rc;
    if (which <= DVS_BUF_CLAMPED) {
        // unpack the records into upstream-style arrays on the device, copy the requested one
        int32_t* radii; uint32_t* tiles; float *depth, *m2, *co, *rgb; uint8_t* cl;
        const size_t n1 = N ? N : 1;
        CK(cudaMalloc((void**)&radii, n1 * 4)); CK(cudaMalloc((void**)&tiles, n1 * 4));
        CK(cudaMalloc((void**)&depth, n1 * 4)); CK(cudaMalloc((void**)&m2, n1 * 8));
        CK(cudaMalloc((void**)&co, n1 * 16)); CK(cudaMalloc((void**)&rgb, n1 * 12)); CK(cudaMalloc((void**)&cl, n1 * 3));
        cudaError_t e = launch_unpack((int)N, ctx->rec, radii, tiles, depth, m2, co, rgb, cl, 0);
        const void* src = nullptr; size_t bytes = 0;
        switch (which) {
            case DVS_BUF_RADII: src = radii; bytes = N * 4; break;
            case DVS_BUF_TILES_TOUCHED: src = tiles; bytes = N * 4; break;
            case DVS_BUF_DEPTH: src = depth; bytes = N * 4; break;
            case DVS_BUF_MEAN2D: src = m2; bytes = N * 8; break;
            case DVS_BUF_CONIC_OPACITY: src = co; bytes = N * 16; break;
            case DVS_BUF_RGB: src = rgb; bytes = N * 12; break;
            default: src = cl; bytes = N * 3; break;
        }
        rc = need(bytes);
        if (e == cudaSuccess && rc == DVS_OK && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
        cudaFree(radii); cudaFree(tiles); cudaFree(depth); cudaFree(m2); cudaFree(co); cudaFree(rgb); cudaFree(cl);
        if (rc) return rc;
        CK(e);
        return DVS_OK;
    }
    switch (which) {
        case DVS_BUF_POINT_LIST:
        case DVS_BUF_CULL_MASK: {
            const bool ids = which == DVS_BUF_POINT_LIST;
            if ((rc = need(D * (ids ? 4 : 1)))) return rc;
            if (!D) return DVS_OK;
            void* tmp;
            CK(cudaMalloc(&tmp, D * (ids ? 4 : 1)));
            cudaError_t e = launch_unpack_plist((uint32_t)D, ctx->plist, ids ? (uint32_t*)tmp : nullptr,
                                                ids ? nullptr : (uint8_t*)tmp, 0);
            if (e == cudaSuccess) e = cudaMemcpy(dst, tmp, D * (ids ? 4 : 1), cudaMemcpyDeviceToHost);
            cudaFree(tmp);
            CK(e);
            return DVS_OK;
        }
        case DVS_BUF_RANGES: {
            if ((rc = need(T * 8))) return rc;
            uint32_t* h = new uint32_t[T + 1];
            cudaError_t e = cudaMemcpy(h, ctx->tile_base, (T + 1) * 4, cudaMemcpyDeviceToHost);
            uint32_t* o = static_cast<uint32_t*>(dst);
            for (size_t t = 0; t < T; t++) {
                // upstream leaves untouched tiles at (0,0)
                const bool empty = h[t + 1] == h[t];
                o[2 * t] = empty ? 0u : h[t];
                o[2 * t + 1] = empty ? 0u : h[t + 1];
            }
            delete[] h;
            CK(e);
            return DVS_OK;
        }
        case DVS_BUF_FINAL_T:
            if ((rc = need(P * 4))) return rc;
            CK(cudaMemcpy(dst, ctx->final_T, P * 4, cudaMemcpyDeviceToHost));
            return DVS_OK;
        case DVS_BUF_N_CONTRIB:
            if ((rc = need(P * 4))) return rc;
            CK(cudaMemcpy(dst, ctx->n_contrib, P * 4, cudaMemcpyDeviceToHost));
            return DVS_OK;
        case DVS_BUF_SCREEN_GRADS:
            if ((rc = need(N * 48))) return rc;
            CK(cudaMemcpy(dst, ctx->sgrad, N * 48, cudaMemcpyDeviceToHost));
            return DVS_OK;
        default:
            return fail(ctx, DVS_E_INVALID, "unknown buffer id %d", which);
    }
}

static const char* kStageNames[DVS_NUM_STAGES] = {"preprocess_fwd", "tile_scan", "emit", "tile_sort",
                                                  "render_fwd",     "render_bwd", "preprocess_bwd", "reserved"};
const char* dvs_rast_stage_name(int i) { return (i >= 0 && i < DVS_NUM_STAGES) ? kStageNames[i] : ""; }

int dvs_rast_stage_ms(dvs_rast_ctx* ctx, float out_ms[DVS_NUM_STAGES]) {
    if (!ctx || !out_ms) return DVS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < DVS_NUM_STAGES; i++) out_ms[i] = 0.f;
    if (ctx->ev_fwd)
        for (int i = 0; i < 5; i++) CK(cudaEventElapsedTime(&out_ms[i], ctx->ev[i], ctx->ev[i + 1]));
    if (ctx->ev_bwd) {
        CK(cudaEventElapsedTime(&out_ms[5], ctx->ev[6], ctx->ev[7]));
        CK(cudaEventElapsedTime(&out_ms[6], ctx->ev[7], ctx->ev[8]));
    }
    return DVS_OK;
}

}  // extern "C"
