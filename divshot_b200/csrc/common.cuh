// common.cuh — shared device-side definitions of the B200-native rasterizer (sm_100a only).
//
// HBM layout (all arenas owned by dvs_rast_ctx, SoA inputs owned by the caller):
//   rec   [N] x 48 B  screen record, three float4:
//           q0 = { mean2D.x, mean2D.y, A2, C2 }          A2 = -0.5*log2(e)*conicA, C2 = -0.5*log2(e)*conicC
//           q1 = { B2, lo, r, g }                        B2 = -log2(e)*conicB, lo = log2(opacity)
//         ({mx, my} and {A2, C2} are the operand pairs of the compositors' packed FADD2 / FMUL2, so the record is staged
//          into shared memory by verbatim 16-byte asynchronous copies and read back as aligned register pairs)
//           q2 = { b, depth, radius (int bits), tiles_touched | clamped<<24 (uint bits) }
//         so that alpha = ex2(A2*dx^2 + B2*dx*dy + C2*dy^2 + lo)  (one MUFU, no multiply by opacity)
//   aux   [N] x 16 B  { minx|miny<<16, maxx|maxy<<16|clamped<<29, depth bits, 0 }
//   bins  [max(Dcap, T*stride)] x 8 B  unsorted per-tile entries  depth_bits<<32 | id<<8 | submask
//   plist [Dcap] x 4 B  sorted entries id<<8 | submask (tile-major; the parity point_list is id)
//   tile_count / tile_base / tile_cursor [T]
//   final_T, n_contrib [P];  sgrad [N] x 48 B screen-space gradient record (see dvs_rast.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dvs_rast.h"

namespace dvs {

constexpr int TILE = 16;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float ALPHA_MIN_LOG2 = -7.994353436858858f;  // log2(1/255)
constexpr int ID_BITS = 24;
constexpr uint32_t MAX_GAUSSIANS = 1u << ID_BITS;
// per-tile atomic counters live 32 B apart: 7.6 M atomics onto 25 KB of packed counters serialise on a few
// L2 lines; one counter per sector spreads them over the L2 slices
constexpr int TILE_CTR_STRIDE = 8;

struct Cam {
    float view[16];
    float proj[16];
    float campos[3];
    float tanfovx, tanfovy;
    int W, H;
    float bg[3];
    float scale_modifier;
    int deg, KR;
    uint32_t flags;
    int gx, gy;
    const float* bg_image;  // optional per-pixel background [3,H,W] (dvs_rast_set_background: the trainer's sky model); null: bg[]
};

struct Params {
    const float* means3D;
    const float* scales;
    const float* quats;
    const float* opacities;
    const float* sh0;
    const float* shN;
};

struct Grads {
    float* means3D;
    float* scales;
    float* quats;
    float* opacities;
    float* sh0;
    float* shN;
    float* mean2D_abs;
    float* mean2D;
};

// single-pass binning (fused into preprocess_fwd): tile t owns bins[t * bin_stride ..); bin_stride == 0 -> two-pass
struct FusedEmit {
    uint32_t* tile_cursor;
    unsigned long long* bins;
    uint32_t bin_stride;
    uint32_t* overflow_word;
    uint32_t tight;  // DVS_FLAG_TIGHT_LISTS: entries whose sub-tile mask is empty are not emitted
};

// ---- small PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// explicit shared-window accessors (32-bit shared addresses; keeps address arithmetic to one IMAD)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
// ---- packed fp32x2 arithmetic (new on sm_100: FFMA2 / FMUL2 / FADD2 — two IEEE fp32 operations per issue slot).
// A pair lives in one 64-bit register, .x in the low half.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float sum2(f32x2 v) {  // lo + hi
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void sts_p2(uint32_t a, f32x2 v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ f32x2 abs2(f32x2 a) { return a & 0x7fffffff7fffffffull; }
__device__ __forceinline__ f32x2 lds_p2(uint32_t a) {  // 64-bit shared load of one pair
    f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_p4(uint32_t a, f32x2& v0, f32x2& v1) {  // 128-bit shared load of two pairs
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(a));
}
// 128-bit vector reduction to global memory (sm_90+: REDG.E.ADD.F32x4); `g` must be 16-byte aligned.
__device__ __forceinline__ void red_add_f4(float* g, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float lds_f1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u1(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u1(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// 16-byte asynchronous global -> shared copy (LDGSTS; L2 only, the gathered records have no reuse inside an SM)
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk store (TMA engine), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace dvs
