// gstrain.cu — the `gstrain` trainer plugin boundary: GaussianTrainerScene + the nine C symbols the
// unmodified diverseshot-cli resolves with dlsym (application/diverseshot-cli/source/gs_train.cpp:24-179;
// loader: diverse/diverse_base/source/core/plugin.cpp:36-166).
//
// Scope (SURVEY.md §8 b, F1/F2 "next"): this file exists so that the B200 rasterizer can be driven through the
// reference's own plugin boundary.  The step around the rasterizer is deliberately small — pick a view,
// rasterize forward (dvs_rast_forward), photometric loss (1-w)*L1 + w*(1-SSIM) with w = ssimWeight (main.cpp:24),
// rasterize backward (dvs_rast_backward), fused Adam with the per-group learning rates of GaussianTrainConfig —
// and keeps every tensor device-resident.
// Refinement (csrc/densify.cu): `densifyStrategy` 1 = MCMC (relocation of dead Gaussians, 5 % growth up
// to `capMax`, exploration noise `noiselr`, opacity/scale regularisers), 0 / 2 = ADC (clone / split / prune on the
// accumulated screen-space gradient, opacity reset every `resetAlphaEvery`).  It acts every `refineEvery` iterations
// for warmupLength < iteration < refineStopIter; a run that ends before `warmupLength` never enters it.
// COLMAP SfM / mesh export of the closed trainer are NOT rebuilt here.
//
// Data sources accepted by load_train_data:
//   "synthetic:N=100000,W=800,H=600,views=8,deg=1,seed=7"  — a random ground-truth splat scene is rendered with
//        this rasterizer into target views; training starts from a perturbed copy (no files needed);
//   a directory holding `cameras.txt` (one line per view:
//        image.ppm W H fx fy  r00 r01 r02 tx  r10 r11 r12 ty  r20 r21 r22 tz   — world->camera, +z forward)
//        binary PPM (P6) images, and optionally `points.txt` (x y z r g b per line) for initialisation.
// save_splat_model writes, by extension of modelPath, the PLY row layout of external/tinygsplat/tiny_gsplat.cpp:168-241
// (x y z, f_dc_0..2, f_rest_0..44 channel-major, opacity, scale_0..2, rot_0..3; raw parameters) or the 32-byte
// `.splat` records of tiny_gsplat.cpp:243-291.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "densify.h"
#include "dvs_model_io.h"
#include "dvs_rast.h"
#include "dvs_viewer_pack.h"
#include "gaussian_trainer_scene.hpp"
#include "sh_grad_ops.h"

#define GS_EXPORT __attribute__((visibility("default")))

namespace {

constexpr int KR = 15;  // rest coefficients stored per Gaussian (degree 3), as the trainer exports them

void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("gstrain: ") + what + ": " + cudaGetErrorString(e));
}
void ckr(int rc, dvs_rast_ctx* ctx, const char* what) {
    if (rc != DVS_OK) throw std::runtime_error(std::string("gstrain: ") + what + ": " + dvs_rast_last_error(ctx));
}

// ---------------------------------------------------------------------------------------------------------
// Photometric loss of the trainer step: L = (1 - w) * L1 + w * (1 - SSIM), w = ssimWeight
// (application/diverseshot-cli/source/main.cpp:24 "ssim", default 0.2), and its gradient dL/dpixel — the only
// thing the rasterizer backward needs.  SSIM uses the standard 11x11 Gaussian window (sigma 1.5), zero padding,
// C1 = 0.01^2, C2 = 0.03^2, mean over the 3*H*W map.  Two tiled kernels with separable convolutions in shared
// memory: (A) window statistics -> SSIM map -> per-pixel partials d/d(mu_x), d/d(E[x^2]), d/d(E[xy]);
// (B) convolve the partials back (the window is symmetric) and add the L1 term.
// ---------------------------------------------------------------------------------------------------------
constexpr int SS_T = 16, SS_H = 5, SS_W = SS_T + 2 * SS_H;  // tile, halo, padded tile
__constant__ float c_gauss[11] = {0.00102838f, 0.00759876f, 0.03600077f, 0.10936069f, 0.21300553f, 0.26601172f,
                                  0.21300553f, 0.10936069f, 0.03600077f, 0.00759876f, 0.00102838f};

__global__ void __launch_bounds__(SS_T* SS_T)
ssim_partials_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ dm,
                     float* __restrict__ d2, float* __restrict__ dxy, float* __restrict__ loss, int W, int H,
                     float w_ssim, float inv_n) {
    __shared__ float sx[SS_W][SS_W + 1], sy[SS_W][SS_W + 1];
    __shared__ float h[5][SS_W][SS_T + 1];
    const int ch = blockIdx.z, x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)ch * W * H;
    const int tid = threadIdx.y * SS_T + threadIdx.x;
    for (int t = tid; t < SS_W * SS_W; t += SS_T * SS_T) {
        const int r = t / SS_W, c = t % SS_W, gx = x0 + c - SS_H, gy = y0 + r - SS_H;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        sx[r][c] = in ? X[plane + (size_t)gy * W + gx] : 0.f;
        sy[r][c] = in ? Y[plane + (size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int t = tid; t < SS_W * SS_T; t += SS_T * SS_T) {  // horizontal pass
        const int r = t / SS_T, c = t % SS_T;
        float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k], xv = sx[r][c + k], yv = sy[r][c + k];
            a0 = fmaf(g, xv, a0); a1 = fmaf(g, yv, a1); a2 = fmaf(g, xv * xv, a2); a3 = fmaf(g, yv * yv, a3);
            a4 = fmaf(g, xv * yv, a4);
        }
        h[0][r][c] = a0; h[1][r][c] = a1; h[2][r][c] = a2; h[3][r][c] = a3; h[4][r][c] = a4;
    }
    __syncthreads();
    const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
    float local = 0.f;
    if (px < W && py < H) {
        float mx = 0, my = 0, ex2 = 0, ey2 = 0, exy = 0;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k];
            mx = fmaf(g, h[0][threadIdx.y + k][threadIdx.x], mx); my = fmaf(g, h[1][threadIdx.y + k][threadIdx.x], my);
            ex2 = fmaf(g, h[2][threadIdx.y + k][threadIdx.x], ex2); ey2 = fmaf(g, h[3][threadIdx.y + k][threadIdx.x], ey2);
            exy = fmaf(g, h[4][threadIdx.y + k][threadIdx.x], exy);
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float sxx = ex2 - mx * mx, syy = ey2 - my * my, sxy = exy - mx * my;
        const float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2, B1 = mx * mx + my * my + C1, B2 = sxx + syy + C2;
        const float iB = 1.f / (B1 * B2);
        const float ssim = A1 * A2 * iB;
        // dL/dssim = -w/n ; chain to the three window statistics that depend on X
        const float gs = -w_ssim * inv_n;
        const float dmu = 2.f * my * (A2 - A1) * iB - 2.f * mx * A1 * A2 * (B2 - B1) * iB * iB;
        const float de2 = -A1 * A2 * iB / B2;
        const float dex = 2.f * A1 * iB;
        const size_t o = plane + (size_t)py * W + px;
        dm[o] = gs * dmu; d2[o] = gs * de2; dxy[o] = gs * dex;
        const float dlt = sx[threadIdx.y + SS_H][threadIdx.x + SS_H] - sy[threadIdx.y + SS_H][threadIdx.x + SS_H];
        local = (w_ssim * (1.f - ssim) + (1.f - w_ssim) * fabsf(dlt)) * inv_n;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((tid & 31) == 0) atomicAdd(loss, local);
}

__global__ void __launch_bounds__(SS_T* SS_T)
ssim_backward_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ dm,
                     const float* __restrict__ d2, const float* __restrict__ dxy, float* __restrict__ dL_dpix, int W,
                     int H, float w_ssim, float inv_n) {
    __shared__ float s[3][SS_W][SS_W + 1];
    __shared__ float h[3][SS_W][SS_T + 1];
    const int ch = blockIdx.z, x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const size_t plane = (size_t)ch * W * H;
    const int tid = threadIdx.y * SS_T + threadIdx.x;
    for (int t = tid; t < SS_W * SS_W; t += SS_T * SS_T) {
        const int r = t / SS_W, c = t % SS_W, gx = x0 + c - SS_H, gy = y0 + r - SS_H;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t o = plane + (size_t)gy * W + gx;
        s[0][r][c] = in ? dm[o] : 0.f; s[1][r][c] = in ? d2[o] : 0.f; s[2][r][c] = in ? dxy[o] : 0.f;
    }
    __syncthreads();
    for (int t = tid; t < SS_W * SS_T; t += SS_T * SS_T) {
        const int r = t / SS_T, c = t % SS_T;
        float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k];
            a0 = fmaf(g, s[0][r][c + k], a0); a1 = fmaf(g, s[1][r][c + k], a1); a2 = fmaf(g, s[2][r][c + k], a2);
        }
        h[0][r][c] = a0; h[1][r][c] = a1; h[2][r][c] = a2;
    }
    __syncthreads();
    const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
    if (px < W && py < H) {
        float c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = c_gauss[k];
            c0 = fmaf(g, h[0][threadIdx.y + k][threadIdx.x], c0); c1 = fmaf(g, h[1][threadIdx.y + k][threadIdx.x], c1);
            c2 = fmaf(g, h[2][threadIdx.y + k][threadIdx.x], c2);
        }
        const size_t o = plane + (size_t)py * W + px;
        const float xv = X[o], yv = Y[o], d = xv - yv;
        const float l1 = (1.f - w_ssim) * inv_n * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        dL_dpix[o] = c0 + 2.f * xv * c1 + yv * c2 + l1;
    }
}

// loss + dL/dpix on `stream`; scratch: 9*W*H floats.  *loss (device) must be zeroed by the caller.
static void launch_photometric_loss(const float* render, const float* target, float* dL_dpix, float* loss,
                                    float* scratch, int W, int H, float w_ssim, cudaStream_t st) {
    const size_t P3 = (size_t)3 * W * H;
    const float inv_n = 1.f / (float)P3;
    const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, 3), block(SS_T, SS_T);
    ssim_partials_kernel<<<grid, block, 0, st>>>(render, target, scratch, scratch + P3, scratch + 2 * P3, loss, W, H,
                                                 w_ssim, inv_n);
    ssim_backward_kernel<<<grid, block, 0, st>>>(render, target, scratch, scratch + P3, scratch + 2 * P3, dL_dpix, W,
                                                 H, w_ssim, inv_n);
}

// useMask (main.cpp:69-70): a per-view mask M in [0,1] restricts the photometric loss to the masked region.  The loss sees
// render' = M render + (1 - M) target (identical to the target where M = 0, so neither L1 nor SSIM reports a difference
// there) and the chain rule gives dL/drender = M dL/drender'.  Two element-wise passes around the unchanged loss kernels.
__global__ void mask_blend_kernel(float* __restrict__ render, const float* __restrict__ target, const float* __restrict__ mask,
                                  size_t P) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 3 * P; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mask[i % P];
        render[i] = fmaf(m, render[i] - target[i], target[i]);
    }
}
__global__ void unpack_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i] / 255.f;  // the same value the fp32 path uploads (rgb / 255.f)
}
__global__ void mask_grad_kernel(float* __restrict__ dL_dpix, const float* __restrict__ mask, size_t P) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 3 * P; i += (size_t)gridDim.x * blockDim.x)
        dL_dpix[i] *= mask[i % P];
}

// Fused multi-tensor Adam: ONE launch steps all six parameter groups of the flat arenas (params / grads / two moments share
// one layout), each with its own learning rate — 16 B of the four streams per element and thread, 128-bit accesses (group
// offsets are multiples of 4 floats; a group's last partial vector is done element-wise).  `radii` != nullptr is `visibleAdam`
// (the "sparse Adam" of the 3DGS accelerations the closed trainer's flag is named after): only Gaussians the current view saw
// (radius > 0) are stepped, the moments of the others are left untouched instead of decaying.
// `skip` (device word, may be null): non-zero = the step's forward overflowed its binning arena and produced no image and no
// gradients (info[2] of the rasterizer context) — the update must not run (a zero gradient still moves every parameter
// through the decaying first moment); the step is redone by the host.
struct AdamGroup { size_t off, n; float lr; int width; };
struct AdamGroups { AdamGroup g[6]; };

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps, float c1, float c2) {
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    p -= lr * (m * c1) / (sqrtf(v * c2) + eps);
}

__global__ void __launch_bounds__(256)
adam_fused_kernel(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ M, float* __restrict__ V,
                  const AdamGroups groups, const int32_t* __restrict__ radii, const uint32_t* __restrict__ skip, float b1,
                  float b2, float eps, float c1, float c2) {
    if (skip && *skip) return;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
        const AdamGroup gr = groups.g[k];
        if (gr.n == 0) continue;
        float4* p4 = reinterpret_cast<float4*>(P + gr.off);
        const float4* g4 = reinterpret_cast<const float4*>(G + gr.off);
        float4* m4 = reinterpret_cast<float4*>(M + gr.off);
        float4* v4 = reinterpret_cast<float4*>(V + gr.off);
        const size_t nv = gr.n >> 2;
        for (size_t i = tid; i < nv; i += nthr) {
            float4 p = p4[i], m = m4[i], v = v4[i];
            const float4 g = g4[i];
            if (radii) {
                const size_t e = 4 * i;
                const bool s0 = radii[e / gr.width] > 0, s1 = radii[(e + 1) / gr.width] > 0, s2 = radii[(e + 2) / gr.width] > 0,
                           s3 = radii[(e + 3) / gr.width] > 0;
                if (!(s0 || s1 || s2 || s3)) continue;
                if (s0) adam_one(p.x, g.x, m.x, v.x, gr.lr, b1, b2, eps, c1, c2);
                if (s1) adam_one(p.y, g.y, m.y, v.y, gr.lr, b1, b2, eps, c1, c2);
                if (s2) adam_one(p.z, g.z, m.z, v.z, gr.lr, b1, b2, eps, c1, c2);
                if (s3) adam_one(p.w, g.w, m.w, v.w, gr.lr, b1, b2, eps, c1, c2);
            } else {
                adam_one(p.x, g.x, m.x, v.x, gr.lr, b1, b2, eps, c1, c2);
                adam_one(p.y, g.y, m.y, v.y, gr.lr, b1, b2, eps, c1, c2);
                adam_one(p.z, g.z, m.z, v.z, gr.lr, b1, b2, eps, c1, c2);
                adam_one(p.w, g.w, m.w, v.w, gr.lr, b1, b2, eps, c1, c2);
            }
            p4[i] = p; m4[i] = m; v4[i] = v;
        }
        for (size_t e = (nv << 2) + tid; e < gr.n; e += nthr) {  // the group's last partial vector
            if (radii && radii[e / gr.width] <= 0) continue;
            const size_t j = gr.off + e;
            float p = P[j], m = M[j], v = V[j];
            adam_one(p, G[j], m, v, gr.lr, b1, b2, eps, c1, c2);
            P[j] = p; M[j] = m; V[j] = v;
        }
    }
}

// ---- background ("sky") model: GaussianTrainConfig::enableBg, "Create Sky Model" (docs/userGuide.md:53) -------------------
// What lies behind the splats — sky, far scenery the point cloud does not cover — is a smooth function of the viewing
// direction: 9 real SH coefficients (degree 2) per colour channel, 27 learnable floats.  Every step the model is evaluated per
// pixel into a background image the rasterizer composites over (out = C + final_T * bg(pixel), dvs_rast_set_background), and
// its gradient is the SH-weighted sum of dL/dbg = final_T * dL/dpix over the pixels (dvs_rast_background_grad).  The closed
// trainer's sky model is absent from the reference: parity unpinned; the arithmetic is checked against
// numpy (tests/test_plugin.py).  The model file formats have no field for it: it lives with the trainer object.
constexpr int SKY_K = 9;
struct SkyCam { float Rt[9]; float tanx, tany; int W, H; };  // rotation rows of the view matrix (world -> camera)

__device__ __forceinline__ void sky_basis(const SkyCam& c, int px, int py, float Y[SKY_K]) {
    // pixel centre -> NDC (the inverse of ndc2Pix, gsplat_vs.hlsl:211-214) -> camera ray -> world direction
    const float nx = (2.0f * (float)px + 1.0f) / (float)c.W - 1.0f, ny = (2.0f * (float)py + 1.0f) / (float)c.H - 1.0f;
    const float rx = nx * c.tanx, ry = ny * c.tany, rz = 1.0f;
    float dx = c.Rt[0] * rx + c.Rt[3] * ry + c.Rt[6] * rz, dy = c.Rt[1] * rx + c.Rt[4] * ry + c.Rt[7] * rz,
          dz = c.Rt[2] * rx + c.Rt[5] * ry + c.Rt[8] * rz;  // R^T ray
    const float li = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= li; dy *= li; dz *= li;
    float b[15];
    dvs_shx::sh_rest_basis(2, dx, dy, dz, b);
    Y[0] = dvs_shx::kC0;
#pragma unroll
    for (int k = 1; k < SKY_K; k++) Y[k] = b[k - 1];
}
__global__ void sky_eval_kernel(SkyCam c, const float* __restrict__ coef /* [9][3] */, float* __restrict__ bg /* [3,H,W] */) {
    const size_t P = (size_t)c.W * c.H;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        float Y[SKY_K];
        sky_basis(c, (int)(p % c.W), (int)(p / c.W), Y);
        float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < SKY_K; k++) { r += Y[k] * coef[3 * k]; g += Y[k] * coef[3 * k + 1]; b += Y[k] * coef[3 * k + 2]; }
        bg[p] = r; bg[P + p] = g; bg[2 * P + p] = b;
    }
}
// dL/dcoef[k][ch] += sum_p Y_k(dir_p) * dL/dbg[ch][p]: per-thread partial sums over a grid-stride loop, warp shuffles, one
// atomic per (warp, coefficient, channel)
__global__ void sky_grad_kernel(SkyCam c, const float* __restrict__ dbg /* [3,H,W] */, float* __restrict__ dcoef /* [9][3] */) {
    const size_t P = (size_t)c.W * c.H;
    float acc[3 * SKY_K];
#pragma unroll
    for (int k = 0; k < 3 * SKY_K; k++) acc[k] = 0.f;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        float Y[SKY_K];
        sky_basis(c, (int)(p % c.W), (int)(p / c.W), Y);
        const float r = dbg[p], g = dbg[P + p], b = dbg[2 * P + p];
#pragma unroll
        for (int k = 0; k < SKY_K; k++) { acc[3 * k] += Y[k] * r; acc[3 * k + 1] += Y[k] * g; acc[3 * k + 2] += Y[k] * b; }
    }
#pragma unroll
    for (int k = 0; k < 3 * SKY_K; k++) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(dcoef + k, v);
    }
}
__global__ void sky_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float lr,
                                float b1, float b2, float eps, float c1, float c2) {
    const int i = threadIdx.x;
    if (i < 3 * SKY_K) {
        adam_one(p[i], g[i], m[i], v[i], lr, b1, b2, eps, c1, c2);
        g[i] = 0.f;  // the gradient accumulator starts the next step from zero
    }
}
static SkyCam sky_cam(const dvs_camera& cam) {
    SkyCam c;
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) c.Rt[3 * r + k] = cam.view[4 * k + r];  // view is flat [4 c + r]: rotation rows
    c.tanx = cam.tanfovx; c.tany = cam.tanfovy; c.W = cam.width; c.H = cam.height;
    return c;
}

// ---- normal-consistency loss: GaussianTrainConfig::normalConsistencyLoss (gs_train.cpp:79-84, docs/userGuide.md:52-58) -----
// The loss of the 2DGS paper (its eq. 14) on the maps dvs_rast_forward_aux renders: the rendered normal map N = sum_k w_k n_k
// must agree with the normal of the rendered SURFACE, taken from the depth map by finite differences:
//     D = depth / alpha (alpha > 1e-6, else 0);  P(x, y) = D(x, y) * ray(x, y),  ray = (ndc_x tan_x, ndc_y tan_y, 1);
//     n_d = normalize( (P(x+1, y) - P(x-1, y)) x (P(x, y+1) - P(x, y-1)) )   (interior pixels);
//     L = lambda / (W H) * sum_p ( 1 - alpha_p (N_p . n_d,p) )               (alpha_p as a constant weight).
// Gradients go to the normal map directly and to depth / alpha of the four neighbours through n_d; dvs_rast_backward_aux
// carries them on to the Gaussians.  The closed trainer's loss is absent from the reference: parity
// unpinned, the kernel is checked against torch autograd of the same formula (tests/test_plugin.py).
__global__ void normal_consistency_kernel(int W, int H, float tanx, float tany, const float* __restrict__ aux /* [2,H,W] */,
                                          const float* __restrict__ nrm /* [3,H,W] */, float lam_over_p, float* __restrict__ loss,
                                          float* __restrict__ d_aux /* [2,H,W], zeroed */, float* __restrict__ d_nrm /* [3,H,W] */) {
    const size_t P = (size_t)W * H;
    float local = 0.f;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % W), y = (int)(p / W);
        float gn0 = 0.f, gn1 = 0.f, gn2 = 0.f;
        if (x > 0 && y > 0 && x < W - 1 && y < H - 1) {
            const size_t q[4] = {p + 1, p - 1, p + (size_t)W, p - (size_t)W};  // x+1, x-1, y+1, y-1
            const int qx[4] = {x + 1, x - 1, x, x}, qy[4] = {y, y, y + 1, y - 1};
            float Dq[4], Aq[4], ray[4][2], Pq[4][3];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                Aq[k] = aux[P + q[k]];
                Dq[k] = Aq[k] > 1e-6f ? aux[q[k]] / Aq[k] : 0.f;
                ray[k][0] = ((2.f * (float)qx[k] + 1.f) / (float)W - 1.f) * tanx;
                ray[k][1] = ((2.f * (float)qy[k] + 1.f) / (float)H - 1.f) * tany;
                Pq[k][0] = Dq[k] * ray[k][0]; Pq[k][1] = Dq[k] * ray[k][1]; Pq[k][2] = Dq[k];
            }
            const float dx[3] = {Pq[0][0] - Pq[1][0], Pq[0][1] - Pq[1][1], Pq[0][2] - Pq[1][2]};
            const float dy[3] = {Pq[2][0] - Pq[3][0], Pq[2][1] - Pq[3][1], Pq[2][2] - Pq[3][2]};
            const float c[3] = {dx[1] * dy[2] - dx[2] * dy[1], dx[2] * dy[0] - dx[0] * dy[2], dx[0] * dy[1] - dx[1] * dy[0]};
            const float len = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            const float a = aux[P + p];
            if (len > 1e-20f) {
                const float n[3] = {c[0] / len, c[1] / len, c[2] / len};
                const float N0 = nrm[p], N1 = nrm[P + p], N2 = nrm[2 * P + p];
                local += lam_over_p * (1.f - a * (N0 * n[0] + N1 * n[1] + N2 * n[2]));
                gn0 = -lam_over_p * a * n[0]; gn1 = -lam_over_p * a * n[1]; gn2 = -lam_over_p * a * n[2];
                // dL/dn_d = -lam a N  ->  dL/dc = (I - n n^T) / |c| dL/dn_d  ->  dL/ddx = dy x dL/dc,  dL/ddy = dL/dc x dx
                const float g[3] = {-lam_over_p * a * N0, -lam_over_p * a * N1, -lam_over_p * a * N2};
                const float ng = n[0] * g[0] + n[1] * g[1] + n[2] * g[2];
                const float dc[3] = {(g[0] - n[0] * ng) / len, (g[1] - n[1] * ng) / len, (g[2] - n[2] * ng) / len};
                const float ddx[3] = {dy[1] * dc[2] - dy[2] * dc[1], dy[2] * dc[0] - dy[0] * dc[2], dy[0] * dc[1] - dy[1] * dc[0]};
                const float ddy[3] = {dc[1] * dx[2] - dc[2] * dx[1], dc[2] * dx[0] - dc[0] * dx[2], dc[0] * dx[1] - dc[1] * dx[0]};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float* dd = k < 2 ? ddx : ddy;
                    const float sgn = (k & 1) ? -1.f : 1.f;
                    const float dD = sgn * (ray[k][0] * dd[0] + ray[k][1] * dd[1] + dd[2]);  // P = D (ray_x, ray_y, 1)
                    if (Aq[k] > 1e-6f) {
                        atomicAdd(d_aux + q[k], dD / Aq[k]);                    // D = depth / alpha
                        atomicAdd(d_aux + P + q[k], -dD * Dq[k] / Aq[k]);
                    }
                }
            } else {
                local += lam_over_p;
            }
        }
        d_nrm[p] = gn0; d_nrm[P + p] = gn1; d_nrm[2 * P + p] = gn2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0.f) atomicAdd(loss, local);
}
static void launch_normal_consistency(int W, int H, float tanx, float tany, const float* aux, const float* nrm, float lambda, float* loss,
                                      float* d_aux, float* d_nrm, cudaStream_t st) {
    const size_t P = (size_t)W * H;
    cudaMemsetAsync(d_aux, 0, 2 * P * sizeof(float), st);
    normal_consistency_kernel<<<148 * 4, 256, 0, st>>>(W, H, tanx, tany, aux, nrm, lambda / (float)P, loss, d_aux, d_nrm);
}

struct View {
    dvs_camera cam;
    float* d_target = nullptr;  // [3,H,W] device, fp32 — or, with GSPackLevel::PackF32ToU8 and 8-bit source images:
    uint8_t* d_target_u8 = nullptr;  // [3,H,W] device, the image's own bytes (a quarter of the memory; unpacked per step)
    float* d_mask = nullptr;    // [H,W] device, optional (useMask)
    float Rt[12] = {0};         // world -> camera rows [R | t] as loaded
    float P[16] = {0};          // the perspective matrix alone, flat [4c+r]
    float fx = 0.f, fy = 0.f;
    std::string name;
    std::vector<uint8_t> rgba;  // lazily built RGBA8 copy of the target (getSplatImageView)
};

struct Arena {  // one flat buffer, six 16-byte-aligned views (same order as GradBuffers in rasterizer.py)
    float* flat = nullptr;
    size_t total = 0;
    size_t off_quats, off_shN, off_means, off_scales, off_sh0, off_opac;
    void layout(int64_t N) {
        size_t o = 0;
        auto take = [&](size_t n) { o = (o + 3) / 4 * 4; size_t r = o; o += n; return r; };
        off_quats = take(4 * N); off_shN = take(3 * KR * N); off_means = take(3 * N);
        off_scales = take(3 * N); off_sh0 = take(3 * N); off_opac = take(N);
        total = (o + 3) / 4 * 4;
    }
    void alloc(int64_t N) {  // N = capacity in Gaussians; the live count is GaussianTrainerImpl::N
        layout(N);
        ck(cudaMalloc(&flat, total * sizeof(float)), "cudaMalloc arena");
        ck(cudaMemset(flat, 0, total * sizeof(float)), "memset arena");
    }
    void release() { if (flat) cudaFree(flat); flat = nullptr; }
    float* quats() const { return flat + off_quats; }
    float* shN() const { return flat + off_shN; }
    float* means() const { return flat + off_means; }
    float* scales() const { return flat + off_scales; }
    float* sh0() const { return flat + off_sh0; }
    float* opac() const { return flat + off_opac; }
};

// groups in arena order: quats | shN | means | scales | sh0 | opac  (lrs[] in the same order)
void launch_adam_fused(const Arena& layout, const float* grads, float* m1, float* m2, int64_t N, const float lrs[6],
                       const int32_t* radii, const uint32_t* skip, float b1, float b2, float eps, float c1, float c2, cudaStream_t st) {
    AdamGroups g;
    const size_t offs[6] = {layout.off_quats, layout.off_shN, layout.off_means, layout.off_scales, layout.off_sh0, layout.off_opac};
    const int widths[6] = {4, 3 * KR, 3, 3, 3, 1};
    for (int k = 0; k < 6; k++) g.g[k] = AdamGroup{offs[k], (size_t)widths[k] * (size_t)N, lrs[k], widths[k]};
    adam_fused_kernel<<<148 * 8, 256, 0, st>>>(layout.flat, grads, m1, m2, g, radii, skip, b1, b2, eps, c1, c2);
}

void make_projection(dvs_camera& c, const float Rt[12], int W, int H, float fx, float fy, float Pflat[16] = nullptr) {
    // view (world->camera), flat [4c+r]
    float V[4][4] = {{Rt[0], Rt[1], Rt[2], Rt[3]}, {Rt[4], Rt[5], Rt[6], Rt[7]}, {Rt[8], Rt[9], Rt[10], Rt[11]},
                     {0, 0, 0, 1}};
    const float tanx = W / (2.f * fx), tany = H / (2.f * fy), zn = 0.01f, zf = 100.f;
    float P[4][4] = {{1.f / tanx, 0, 0, 0}, {0, 1.f / tany, 0, 0}, {0, 0, zf / (zf - zn), -(zf * zn) / (zf - zn)},
                     {0, 0, 1, 0}};
    float PV[4][4];
    for (int r = 0; r < 4; r++)
        for (int cc = 0; cc < 4; cc++) {
            float s = 0;
            for (int k = 0; k < 4; k++) s += P[r][k] * V[k][cc];
            PV[r][cc] = s;
        }
    for (int r = 0; r < 4; r++)
        for (int cc = 0; cc < 4; cc++) {
            c.view[4 * cc + r] = V[r][cc]; c.proj[4 * cc + r] = PV[r][cc];
            if (Pflat) Pflat[4 * cc + r] = P[r][cc];
        }
    // camera centre = -R^T t
    for (int k = 0; k < 3; k++) c.campos[k] = -(Rt[0 + k] * Rt[3] + Rt[4 + k] * Rt[7] + Rt[8 + k] * Rt[11]);
    c.tanfovx = tanx; c.tanfovy = tany; c.width = W; c.height = H;
    c.bg[0] = c.bg[1] = c.bg[2] = 0.f;
    c.scale_modifier = 1.f; c.sh_degree = 0; c.sh_rest_alloc = KR; c.flags = 0;
}

}  // namespace

// ---- view-sharded data parallelism behind the plugin boundary (SURVEY.md §8 e) ---------------------------------------------
// One process per GPU, each running the UNMODIFIED caller (diverseshot-cli / gstrain_driver) against this plugin; rank, world
// size and device come from the environment a launcher sets (DVS_RANK / DVS_WORLD_SIZE / DVS_LOCAL_RANK, else the RANK /
// WORLD_SIZE / LOCAL_RANK of torch.distributed.run).  Every rank holds the whole model, renders ITS view of the step's batch
// (dp.views_for_rank: view (step * world + rank) mod views) and the per-Gaussian gradients are summed over ranks before the
// optimiser — so all ranks step identically and the run equals a single-GPU run that accumulates the same batch of views.
// NCCL is loaded at run time (dlopen libnccl.so.2): a single-GPU run needs no NCCL at all.  The NCCL unique id travels through
// a file (DVS_NCCL_ID_FILE, default /tmp/dvs_nccl_id.<MASTER_PORT>): rank 0 writes it, the others wait for it.
struct DpComm {
    typedef struct ncclComm* comm_t;
    struct unique_id { char internal[128]; };
    int rank = 0, world = 1, local = 0;
    void* lib = nullptr;
    comm_t comm = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    static constexpr int kFloat32 = 7, kSum = 0;  // ncclFloat32, ncclSum (stable across NCCL 2.x)
    // DVS_DP_TIMING=1: device time of every gradient exchange (CUDA events, resolved when the scene is destroyed)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;

    static int env_int(const char* a, const char* b, int dflt) {
        const char* v = std::getenv(a);
        if (!v || !*v) v = std::getenv(b);
        return (v && *v) ? std::atoi(v) : dflt;
    }
    void read_env() {
        world = std::max(1, env_int("DVS_WORLD_SIZE", "WORLD_SIZE", 1));
        rank = env_int("DVS_RANK", "RANK", 0);
        local = env_int("DVS_LOCAL_RANK", "LOCAL_RANK", rank);
        if (rank < 0 || rank >= world) throw std::runtime_error("gstrain: rank outside [0, world size)");
        timing = std::getenv("DVS_DP_TIMING") != nullptr;
    }
    void nccl(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string("gstrain: ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "NCCL error"));
    }
    static std::string id_file() {
        if (const char* f = std::getenv("DVS_NCCL_ID_FILE")) return f;
        const char* port = std::getenv("MASTER_PORT");
        return std::string("/tmp/dvs_nccl_id.") + (port ? port : "default");
    }
    void init() {  // the CUDA device of this rank is already current
        if (world <= 1) return;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) throw std::runtime_error(std::string("gstrain: world size > 1 needs NCCL (dlopen libnccl.so.2 failed: ") + dlerror() + ")");
        auto sym = [&](const char* n) {
            void* p = dlsym(lib, n);
            if (!p) throw std::runtime_error(std::string("gstrain: NCCL symbol missing: ") + n);
            return p;
        };
        GetUniqueId = (int (*)(unique_id*))sym("ncclGetUniqueId");
        CommInitRank = (int (*)(comm_t*, int, unique_id, int))sym("ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, comm_t, cudaStream_t))sym("ncclAllReduce");
        GroupStart = (int (*)())sym("ncclGroupStart");
        GroupEnd = (int (*)())sym("ncclGroupEnd");
        CommDestroy = (int (*)(comm_t))sym("ncclCommDestroy");
        GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        unique_id id{};
        const std::string path = id_file();
        if (rank == 0) {
            nccl(GetUniqueId(&id), "ncclGetUniqueId");
            const std::string tmp = path + ".tmp";
            std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
            f.write(id.internal, sizeof id.internal);
            f.close();
            if (!f.good() || std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("gstrain: cannot write " + path);
        } else {
            bool got = false;
            for (int tries = 0; tries < 1200 && !got; tries++) {  // up to 120 s
                std::ifstream f(path, std::ios::binary);
                if (f.good()) {
                    f.read(id.internal, sizeof id.internal);
                    got = f.gcount() == (std::streamsize)sizeof id.internal;
                }
                if (!got) usleep(100000);
            }
            if (!got) throw std::runtime_error("gstrain: rank 0 never published the NCCL id at " + path);
        }
        nccl(CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
        if (rank == 0) {  // every rank has joined once ncclCommInitRank returns: the id file is spent
            std::remove(path.c_str());
        }
    }
    // in-place sum over ranks of several device ranges, one NCCL group (one fused launch)
    void all_reduce_sum(std::initializer_list<std::pair<float*, size_t>> ranges, cudaStream_t st, bool time_it = false) {
        if (world <= 1) return;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timing && time_it && timed.size() < 4096) {
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, st);
        }
        nccl(GroupStart(), "ncclGroupStart");
        for (auto& r : ranges)
            if (r.second) nccl(AllReduce(r.first, r.first, r.second, kFloat32, kSum, comm, st), "ncclAllReduce");
        nccl(GroupEnd(), "ncclGroupEnd");
        if (e0) {
            cudaEventRecord(e1, st);
            timed.emplace_back(e0, e1);
        }
    }
    void report_timing() {
        if (timed.empty()) return;
        cudaDeviceSynchronize();
        double sum = 0.0; size_t n = 0;
        for (size_t k = timed.size() / 4; k < timed.size(); k++) {  // the first quarter is warm-up
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, timed[k].first, timed[k].second) == cudaSuccess) { sum += ms; n++; }
        }
        for (auto& p : timed) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        timed.clear();
        if (n) std::fprintf(stderr, "gstrain: rank %d: gradient exchange %.3f ms per step (device time incl. waiting for the slowest rank, %zu steps)\n",
                            rank, sum / (double)n, n);
    }
    void destroy() {
        report_timing();
        if (comm && CommDestroy) CommDestroy(comm);
        comm = nullptr;
        if (lib) dlclose(lib);
        lib = nullptr;
    }
};

struct GaussianTrainerImpl {
    DpComm dp;                     // view-sharded data parallelism (world size 1: inert)
    int batch_views = 1;           // views rendered per step on this GPU before the optimiser (DVS_BATCH_VIEWS; gradients accumulate)
    dvs_rast_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    int64_t N = 0;
    int max_degree = 3;
    Arena params, grads, m1, m2;
    std::vector<View> views;
    float* d_render = nullptr;
    float* d_dLdpix = nullptr;
    float* d_scratch = nullptr;  // 9*W*H floats for the SSIM partial maps
    float* d_loss = nullptr;
    float* h_loss = nullptr;  // pinned
    size_t img_cap = 0;
    std::mt19937 rng{1234};
    float scene_extent = 1.f;
    // refinement (densify.cu): arenas hold `capacity` Gaussians, N of them live
    int64_t capacity = 0;
    dvs_densify::Workspace* dws = nullptr;
    float* d_accum = nullptr;      // [capacity] ADC: sum of ||dL/dmean2D||
    float* d_denom = nullptr;      // [capacity] ADC: visibility count
    float* d_mean2D = nullptr;     // [capacity,2] screen-space gradient of the last backward
    float* d_mean2D_abs = nullptr; // [capacity,2] (useAbsGrad)
    int32_t* d_radii = nullptr;    // [capacity] radii of the last forward
    bool refine_enabled = false;   // the schedule reaches the refinement window (set before upload)
    bool resync_next = false;      // N changed: the next forward re-sizes the binning arena synchronously
    float* d_target_f32 = nullptr;  // PackF32ToU8: the current view's image unpacked to fp32
    size_t target_f32_cap = 0;
    // normalConsistencyLoss: the auxiliary maps of the step and their gradients [2P | 3P | 2P | 3P]
    float* d_ncl = nullptr;
    size_t ncl_cap = 0;
    // background model (enableBg): 27 SH coefficients, their gradient and Adam moments [4][27], the per-pixel image and dL/dbg
    float* d_sky = nullptr;
    float* d_bg_img = nullptr;
    float* d_dbg = nullptr;
    size_t sky_img_cap = 0;
    int load_itr = -1;             // create_splat(config, loadItr): resume from the model file at config.modelPath at this iteration
    dvs_densify::RefineReport last_report;
    // The editor drives trainStep from a worker thread and reads the model from its UI thread (editor.cpp:1559-1574 vs
    // :1603-1620): every public entry that touches the device state takes this lock, so a reader never sees the model
    // half-way through a refinement (N and the rows would disagree).  Recursive: saveGaussianModel calls the getters.
    std::recursive_mutex mu;
    // editor surface: the initial model (resetGaussian), the initialisation points, time spent training
    std::vector<float> init[6];  // means, scales, quats, opac, sh0, shN as first uploaded
    std::vector<GsPoint3D> points3d;
    double train_seconds = 0.0;
    std::chrono::steady_clock::time_point last_step{};
    bool have_last_step = false;
    // viewer hand-off (viewer_pack.cu): two snapshots in flight at most, device staging + pinned host copy each
    struct ViewerSlot {
        uint8_t* d = nullptr;     // [vp_cap * 104 + 32] device: gaussians | colors | sh | bbox (ordered uint32 x 6)
        uint8_t* h = nullptr;     // pinned host mirror
        cudaEvent_t ready = nullptr;
        int64_t count = 0;
        int iteration = -1;
        bool requested = false;
    } vp[2];
    int64_t vp_cap = 0;
    int vp_next = 0, vp_last = -1;
    cudaStream_t vp_stream = nullptr;
    cudaEvent_t vp_packed = nullptr;

    dvs_densify::Tensors T(const Arena& a) const { return dvs_densify::Tensors{a.means(), a.scales(), a.quats(), a.opac(), a.sh0(), a.shN()}; }

    dvs_params P() const { return dvs_params{params.means(), params.scales(), params.quats(), params.opac(), params.sh0(), params.shN()}; }
    dvs_grads G() const { return dvs_grads{grads.means(), grads.scales(), grads.quats(), grads.opac(), grads.sh0(), grads.shN(), nullptr, nullptr}; }

    void ensure_images(size_t floats) {
        if (floats <= img_cap) return;
        if (d_render) cudaFree(d_render);
        if (d_dLdpix) cudaFree(d_dLdpix);
        if (d_scratch) cudaFree(d_scratch);
        ck(cudaMalloc(&d_render, floats * sizeof(float)), "cudaMalloc render");
        ck(cudaMalloc(&d_dLdpix, floats * sizeof(float)), "cudaMalloc dLdpix");
        ck(cudaMalloc(&d_scratch, 3 * floats * sizeof(float)), "cudaMalloc loss scratch");
        img_cap = floats;
    }
    void release_model() {
        params.release(); grads.release(); m1.release(); m2.release();
        cudaFree(d_accum); cudaFree(d_denom); cudaFree(d_mean2D); cudaFree(d_mean2D_abs); cudaFree(d_radii);
        d_accum = d_denom = d_mean2D = d_mean2D_abs = nullptr; d_radii = nullptr;
    }
    bool want_radii = false;  // visibleAdam: the forward's radii are kept even outside the refinement window
    void allocate(int64_t cap) {  // arenas for `cap` Gaussians, zero-filled
        capacity = cap;
        params.alloc(capacity); grads.alloc(capacity); m1.alloc(capacity); m2.alloc(capacity);
        if (want_radii && !refine_enabled) ck(cudaMalloc(&d_radii, capacity * sizeof(int32_t)), "cudaMalloc radii");
        if (refine_enabled) {  // refinement statistics and the buffers they are fed from
            ck(cudaMalloc(&d_accum, capacity * sizeof(float)), "cudaMalloc accum");
            ck(cudaMalloc(&d_denom, capacity * sizeof(float)), "cudaMalloc denom");
            ck(cudaMalloc(&d_mean2D, 2 * capacity * sizeof(float)), "cudaMalloc mean2D");
            ck(cudaMalloc(&d_mean2D_abs, 2 * capacity * sizeof(float)), "cudaMalloc mean2D_abs");
            ck(cudaMalloc(&d_radii, capacity * sizeof(int32_t)), "cudaMalloc radii");
            ck(cudaMemset(d_accum, 0, capacity * sizeof(float)), "memset accum");
            ck(cudaMemset(d_denom, 0, capacity * sizeof(float)), "memset denom");
            if (!dws) dws = dvs_densify::workspace_create();
        }
    }
    // n rows of raw parameters from the host into the (already allocated) arenas; n <= capacity
    void set_model(const float* means, const float* lscales, const float* quats, const float* logit, const float* sh0,
                   const float* shN, int64_t n) {
        N = n;
        auto up = [&](float* d, const float* h, size_t cnt) {
            if (cnt) ck(cudaMemcpy(d, h, cnt * sizeof(float), cudaMemcpyHostToDevice), "upload");
        };
        up(params.means(), means, 3 * (size_t)n); up(params.scales(), lscales, 3 * (size_t)n); up(params.quats(), quats, 4 * (size_t)n);
        up(params.opac(), logit, (size_t)n); up(params.sh0(), sh0, 3 * (size_t)n); up(params.shN(), shN, (size_t)3 * KR * n);
    }
    void upload(const std::vector<float>& means, const std::vector<float>& lscales, const std::vector<float>& quats,
                const std::vector<float>& logit, const std::vector<float>& sh0, const std::vector<float>& shN) {
        const int64_t n = (int64_t)logit.size();
        allocate(std::max(n, capacity));
        set_model(means.data(), lscales.data(), quats.data(), logit.data(), sh0.data(), shN.data(), n);
    }
    void zero_optimizer_state() {
        ck(cudaMemsetAsync(m1.flat, 0, m1.total * sizeof(float), stream), "memset m1");
        ck(cudaMemsetAsync(m2.flat, 0, m2.total * sizeof(float), stream), "memset m2");
        if (d_accum) {
            ck(cudaMemsetAsync(d_accum, 0, capacity * sizeof(float), stream), "memset accum");
            ck(cudaMemsetAsync(d_denom, 0, capacity * sizeof(float), stream), "memset denom");
        }
    }
    std::vector<float> download(const float* d, size_t n) const {
        std::vector<float> h(n);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost);
        return h;
    }
};

// -------------------------------------------------------------------------------------------------
GaussianTrainerScene::GaussianTrainerScene(const GaussianTrainConfig& config, int loadItr) : config_(config) {
    impl_ = new GaussianTrainerImpl();
    impl_->load_itr = loadItr;
    int dev = 0;
    impl_->dp.read_env();
    if (const char* b = std::getenv("DVS_BATCH_VIEWS")) impl_->batch_views = std::max(1, std::atoi(b));
    // no CPU fallback: without a CUDA device the constructor throws (the CLI then aborts, gs_train.cpp does not catch)
    if (impl_->dp.world > 1) {  // one process per GPU: this rank's device
        int ndev = 0;
        ck(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
        if (ndev <= 0) throw std::runtime_error("gstrain: no CUDA device");
        ck(cudaSetDevice(impl_->dp.local % ndev), "cudaSetDevice");
    }
    ck(cudaGetDevice(&dev), "cudaGetDevice");
    ck(cudaStreamCreateWithFlags(&impl_->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    int rc = dvs_rast_create(dev, &impl_->ctx);
    if (rc != DVS_OK) throw std::runtime_error("gstrain: dvs_rast_create failed (no usable CUDA device)");
    ck(cudaMalloc(&impl_->d_loss, sizeof(float)), "cudaMalloc loss");
    ck(cudaMallocHost(&impl_->h_loss, sizeof(float)), "cudaMallocHost loss");
    impl_->dp.init();  // world size > 1: joins the NCCL communicator of the run (collective: every rank constructs its scene)
}

GaussianTrainerScene::~GaussianTrainerScene() {
    if (!impl_) return;
    cudaDeviceSynchronize();
    for (auto& v : impl_->views) { cudaFree(v.d_target); cudaFree(v.d_target_u8); cudaFree(v.d_mask); }
    cudaFree(impl_->d_target_f32);
    impl_->release_model();
    cudaFree(impl_->d_render); cudaFree(impl_->d_dLdpix); cudaFree(impl_->d_scratch); cudaFree(impl_->d_loss);
    cudaFree(impl_->d_sky); cudaFree(impl_->d_bg_img); cudaFree(impl_->d_dbg); cudaFree(impl_->d_ncl);
    for (auto& v : impl_->vp) { cudaFree(v.d); if (v.h) cudaFreeHost(v.h); if (v.ready) cudaEventDestroy(v.ready); }
    if (impl_->vp_packed) cudaEventDestroy(impl_->vp_packed);
    if (impl_->vp_stream) cudaStreamDestroy(impl_->vp_stream);
    dvs_densify::workspace_destroy(impl_->dws);
    cudaFreeHost(impl_->h_loss);
    impl_->dp.destroy();
    if (impl_->ctx) dvs_rast_destroy(impl_->ctx);
    if (impl_->stream) cudaStreamDestroy(impl_->stream);
    delete impl_;
}

GaussianTrainerScene::GaussianTrainerScene(GaussianTrainerScene&& o) noexcept
    : ShowTrainView(o.ShowTrainView), curIteration(o.curIteration), pruenIteraions(std::move(o.pruenIteraions)),
      focus_region_position(o.focus_region_position), focus_region_rotation(o.focus_region_rotation),
      focus_region_scale(o.focus_region_scale), config_(std::move(o.config_)), status_(o.status_), train_(o.train_),
      terminate_(o.terminate_), loss_(o.loss_), impl_(o.impl_) {
    o.impl_ = nullptr;
}
GaussianTrainerScene& GaussianTrainerScene::operator=(GaussianTrainerScene&& o) noexcept {
    if (this != &o) {
        std::swap(impl_, o.impl_);  // the moved-from object releases our old device state in its destructor
        ShowTrainView = o.ShowTrainView; curIteration = o.curIteration; pruenIteraions = std::move(o.pruenIteraions);
        focus_region_position = o.focus_region_position; focus_region_rotation = o.focus_region_rotation;
        focus_region_scale = o.focus_region_scale;
        config_ = std::move(o.config_); status_ = o.status_; train_ = o.train_; terminate_ = o.terminate_; loss_ = o.loss_;
    }
    return *this;
}

// Refinement can only happen for warmupLength < iteration < min(refineStopIter, numIters): a run that never gets there
// keeps arenas of exactly N Gaussians (and never touches densify.cu).
static bool refinementPossible(const GaussianTrainConfig& c) {
    return c.refineEvery > 0 && c.warmupLength + 1 < std::min(c.refineStopIter, c.numIters);
}
// the rasterizer addresses Gaussians with 24 bits (dvs_rast_forward: N < 2^24), so growth stops there whatever capMax says
static int64_t effectiveCapMax(const GaussianTrainConfig& c) { return std::min<int64_t>(std::max(c.capMax, 1), (1 << 24) - 1); }
int64_t GaussianTrainerScene::plannedCapacity(int64_t N) const {
    return refinementPossible(config_) ? std::max<int64_t>(N, effectiveCapMax(config_)) : N;
}

static void parse_kv(const std::string& s, const char* key, long& out) {
    const std::string k = std::string(key) + "=";
    size_t p = s.find(k);
    if (p != std::string::npos) out = std::strtol(s.c_str() + p + k.size(), nullptr, 10);
}

bool GaussianTrainerScene::loadTrainData(const std::string& path) {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    status_ = TrainingStatus::Loading_Data;
    try {
        if (path.rfind("synthetic:", 0) == 0) {
            long N = 100000, W = 800, H = 600, nviews = 8, deg = 1, seed = 7;
            parse_kv(path, "N", N); parse_kv(path, "W", W); parse_kv(path, "H", H); parse_kv(path, "views", nviews);
            parse_kv(path, "deg", deg); parse_kv(path, "seed", seed);
            I.max_degree = (int)std::min(3l, std::max(0l, deg));
            I.rng.seed((unsigned)seed);
            std::uniform_real_distribution<float> U(0.f, 1.f);
            std::normal_distribution<float> G(0.f, 1.f);
            const float tanx = std::tan(0.5f * 60.f * 3.14159265f / 180.f), tany = tanx * H / W;
            const float mu_s = std::log(0.012f * std::cbrt(1.0e6f / (float)N));
            std::vector<float> means(3 * N), ls(3 * N), q(4 * N), lo(N), sh0(3 * N), shN((size_t)3 * KR * N, 0.f);
            for (long i = 0; i < N; i++) {
                const float z = 2.f + 8.f * U(I.rng);
                means[3 * i] = (2.3f * U(I.rng) - 1.15f) * tanx * z;
                means[3 * i + 1] = (2.3f * U(I.rng) - 1.15f) * tany * z;
                means[3 * i + 2] = z;
                for (int k = 0; k < 3; k++) ls[3 * i + k] = mu_s + 0.5f * G(I.rng);
                float n2 = 0;
                for (int k = 0; k < 4; k++) { q[4 * i + k] = G(I.rng); n2 += q[4 * i + k] * q[4 * i + k]; }
                for (int k = 0; k < 4; k++) q[4 * i + k] /= std::sqrt(n2);
                lo[i] = 1.5f * G(I.rng);
                for (int k = 0; k < 3; k++) sh0[3 * i + k] = G(I.rng);
                const int Kact = (I.max_degree + 1) * (I.max_degree + 1) - 1;
                for (int c = 0; c < Kact * 3; c++) shN[(size_t)3 * KR * i + c] = 0.2f * G(I.rng);
            }
            I.capacity = plannedCapacity((int64_t)lo.size());
            I.refine_enabled = refinementPossible(config_);
            I.want_radii = config_.visibleAdam;
            I.upload(means, ls, q, lo, sh0, shN);
            // target views on a ring, rendered from the ground truth with this rasterizer
            I.ensure_images((size_t)3 * W * H);
            for (long v = 0; v < nviews; v++) {
                const float ang = 2.f * 3.14159265f * v / std::max(1l, nviews), rad = nviews > 1 ? 0.5f : 0.f;
                const float eye[3] = {rad * std::cos(ang), rad * std::sin(ang), 0.f}, tgt[3] = {0, 0, 6};
                float f[3] = {tgt[0] - eye[0], tgt[1] - eye[1], tgt[2] - eye[2]};
                const float fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
                for (auto& x : f) x /= fl;
                float r[3] = {f[2], 0.f, -f[0]};  // up x f, up = (0,1,0)
                const float rl = std::sqrt(r[0] * r[0] + r[2] * r[2]);
                for (auto& x : r) x /= rl;
                const float u[3] = {f[1] * r[2] - f[2] * r[1], f[2] * r[0] - f[0] * r[2], f[0] * r[1] - f[1] * r[0]};
                float Rt[12] = {r[0], r[1], r[2], 0, u[0], u[1], u[2], 0, f[0], f[1], f[2], 0};
                for (int a = 0; a < 3; a++) Rt[4 * a + 3] = -(Rt[4 * a] * eye[0] + Rt[4 * a + 1] * eye[1] + Rt[4 * a + 2] * eye[2]);
                View vw;
                vw.fx = W / (2.f * tanx); vw.fy = H / (2.f * tany);
                std::memcpy(vw.Rt, Rt, sizeof Rt);
                vw.name = "synthetic_view_" + std::to_string(v);
                make_projection(vw.cam, Rt, (int)W, (int)H, vw.fx, vw.fy, vw.P);
                vw.cam.sh_degree = I.max_degree;
                if (config_.modelType == 1) vw.cam.flags |= DVS_FLAG_MODEL_2DGS;  // the ground truth of a 2DGS run is rendered as surfels
                ck(cudaMalloc(&vw.d_target, (size_t)3 * W * H * sizeof(float)), "cudaMalloc target");
                dvs_params P = I.P();
                ckr(dvs_rast_forward(I.ctx, &vw.cam, I.N, &P, vw.d_target, nullptr, I.stream), I.ctx, "render target");
                I.views.push_back(vw);
            }
            // training starts from a perturbed copy: jittered means, grey colours, thinner opacities
            for (long i = 0; i < N; i++) {
                for (int k = 0; k < 3; k++) means[3 * i + k] += 0.01f * G(I.rng);
                for (int k = 0; k < 3; k++) sh0[3 * i + k] = 0.f;
                lo[i] -= 0.5f;
            }
            std::fill(shN.begin(), shN.end(), 0.f);
            ck(cudaStreamSynchronize(I.stream), "sync");
            auto up = [&](float* d, const std::vector<float>& h) { ck(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice), "upload"); };
            up(I.params.means(), means); up(I.params.sh0(), sh0); up(I.params.opac(), lo); up(I.params.shN(), shN);
            I.scene_extent = 5.f;
        } else {
            std::ifstream cams(path + "/cameras.txt");
            if (!cams.good()) return false;
            std::string line;
            std::vector<float> host;
            while (std::getline(cams, line)) {
                if (line.empty() || line[0] == '#') continue;
                std::istringstream ss(line);
                std::string img; int W, H; float fx, fy, Rt[12];
                ss >> img >> W >> H >> fx >> fy;
                for (auto& x : Rt) ss >> x;
                if (!ss) return false;
                std::ifstream f(path + "/" + img, std::ios::binary);
                std::string magic; int w, h, maxv;
                f >> magic >> w >> h >> maxv;
                f.get();
                if (!f.good() || magic != "P6" || w != W || h != H || maxv != 255) return false;
                std::vector<unsigned char> rgb((size_t)3 * W * H);
                f.read(reinterpret_cast<char*>(rgb.data()), rgb.size());
                host.resize((size_t)3 * W * H);
                for (size_t p = 0; p < (size_t)W * H; p++)
                    for (int c = 0; c < 3; c++) host[c * (size_t)W * H + p] = rgb[3 * p + c] / 255.f;
                View vw;
                vw.fx = fx; vw.fy = fy; vw.name = img;
                std::memcpy(vw.Rt, Rt, sizeof Rt);
                make_projection(vw.cam, Rt, W, H, fx, fy, vw.P);
                if (config_.packLevel & PackF32ToU8) {
                    // `packLevel & PackF32ToU8` (gs_train.cpp:90-96, the CLI default): training images stay 8-bit on the device —
                    // 3 instead of 12 bytes per pixel, lossless for 8-bit sources — and are unpacked into one fp32 scratch
                    // image per step (a 5 us kernel at 1600x1000)
                    std::vector<unsigned char> planar((size_t)3 * W * H);
                    for (size_t p = 0; p < (size_t)W * H; p++)
                        for (int c = 0; c < 3; c++) planar[c * (size_t)W * H + p] = rgb[3 * p + c];
                    ck(cudaMalloc(&vw.d_target_u8, planar.size()), "cudaMalloc target (u8)");
                    ck(cudaMemcpy(vw.d_target_u8, planar.data(), planar.size(), cudaMemcpyHostToDevice), "upload image (u8)");
                } else {
                    ck(cudaMalloc(&vw.d_target, host.size() * sizeof(float)), "cudaMalloc target");
                    ck(cudaMemcpy(vw.d_target, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice), "upload image");
                }
                if (config_.useMask) {  // optional <image>.mask.pgm (binary P5, same size): 255 = train on this pixel
                    std::ifstream mf(path + "/" + img + ".mask.pgm", std::ios::binary);
                    std::string mm; int mw = 0, mh = 0, mv = 0;
                    mf >> mm >> mw >> mh >> mv;
                    mf.get();
                    if (mf.good() && mm == "P5" && mw == W && mh == H && mv == 255) {
                        std::vector<unsigned char> g((size_t)W * H);
                        mf.read(reinterpret_cast<char*>(g.data()), g.size());
                        std::vector<float> mk(g.size());
                        for (size_t p = 0; p < g.size(); p++) mk[p] = g[p] / 255.f;
                        ck(cudaMalloc(&vw.d_mask, mk.size() * sizeof(float)), "cudaMalloc mask");
                        ck(cudaMemcpy(vw.d_mask, mk.data(), mk.size() * sizeof(float), cudaMemcpyHostToDevice), "upload mask");
                    }
                }
                I.views.push_back(vw);
                I.ensure_images(host.size());
            }
            if (I.views.empty()) return false;
            // initial points: points.txt (x y z r g b) or nothing -> fail (SfM is out of scope)
            std::ifstream pts(path + "/points.txt");
            if (!pts.good()) return false;
            std::vector<float> means, ls, q, lo, sh0;
            float x, y, z, r, g, b;
            while (pts >> x >> y >> z >> r >> g >> b) {
                means.insert(means.end(), {x, y, z});
                for (int k = 0; k < 3; k++) ls.push_back(std::log(0.01f));
                q.insert(q.end(), {1.f, 0.f, 0.f, 0.f});
                lo.push_back(-2.1972246f);  // sigmoid^-1(0.1)
                sh0.insert(sh0.end(), {(r / 255.f - 0.5f) / 0.28209479f, (g / 255.f - 0.5f) / 0.28209479f, (b / 255.f - 0.5f) / 0.28209479f});
            }
            if (lo.empty()) return false;
            std::vector<float> shN((size_t)3 * KR * lo.size(), 0.f);
            I.capacity = plannedCapacity((int64_t)lo.size());
            I.refine_enabled = refinementPossible(config_);
            I.want_radii = config_.visibleAdam;
            I.upload(means, ls, q, lo, sh0, shN);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        status_ = TrainingStatus::Loading_Failed;
        return false;
    }
    // the editor surface: remember the initial model (resetGaussian) and expose it as the initialisation point cloud
    const float* src[6] = {I.params.means(), I.params.scales(), I.params.quats(), I.params.opac(), I.params.sh0(), I.params.shN()};
    const size_t width[6] = {3, 3, 4, 1, 3, (size_t)3 * KR};
    for (int k = 0; k < 6; k++) I.init[k] = I.download(src[k], width[k] * (size_t)I.N);
    I.points3d.resize((size_t)I.N);
    for (int64_t i = 0; i < I.N; i++) {
        GsPoint3D& pt = I.points3d[(size_t)i];
        pt.x = I.init[0][3 * i]; pt.y = I.init[0][3 * i + 1]; pt.z = I.init[0][3 * i + 2];
        uint8_t* rgb[3] = {&pt.r, &pt.g, &pt.b};
        for (int c = 0; c < 3; c++) {
            const float v = I.init[4][3 * i + c] * 0.28209479177387814f + 0.5f;
            *rgb[c] = (uint8_t)std::lround(255.f * std::min(1.f, std::max(0.f, v)));
        }
        pt.a = 255;
    }
    trainSetup();
    // `--load_itr K` (main.cpp:40-41 -> create_splat(config, K), gs_train.cpp:107): resume from the model this trainer saved
    // at config.modelPath (save_splat_model; any format of the F2 readers) and continue the schedule — learning-rate decay,
    // SH degree, refinement windows — at iteration K.  The checkpoint is the parameters only: the Adam
    // moments restart from zero.  A missing / unreadable file is an error: silently training from scratch at iteration K
    // would be worse.
    if (I.load_itr >= 0) {
        if (!resumeFromModelFile(config_.modelPath)) {
            std::fprintf(stderr, "gstrain: --load_itr %d: cannot resume from '%s': %s\n", I.load_itr, config_.modelPath.c_str(),
                         dvs_model_io_last_error());
            status_ = TrainingStatus::Loading_Failed;
            return false;
        }
        curIteration = I.load_itr;
    }
    return true;
}

bool GaussianTrainerScene::resumeFromModelFile(const std::string& path) {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (path.empty()) return false;
    const int fmt = dvs_model_format_from_path(path.c_str());
    const int64_t n = dvs_model_read(path.c_str(), fmt ? fmt : DVS_FMT_PLY, nullptr, 0, nullptr);
    if (n <= 0) return false;
    std::vector<float> rows((size_t)n * DVS_IO_ROW_FLOATS);
    if (dvs_model_read(path.c_str(), fmt ? fmt : DVS_FMT_PLY, rows.data(), n, nullptr) != n) return false;
    // reader row = RichPoint: pos[3] | f_dc[3] | f_rest[45] channel-major (f_rest[c*15 + j] = shN[j][c]) | opacity | scale[3] | rot[4]
    std::vector<float> pos(3 * (size_t)n), sh0(3 * (size_t)n), shn((size_t)3 * KR * n), op((size_t)n), sc(3 * (size_t)n), rot(4 * (size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const float* r = rows.data() + (size_t)i * DVS_IO_ROW_FLOATS;
        for (int k = 0; k < 3; k++) { pos[3 * i + k] = r[k]; sh0[3 * i + k] = r[3 + k]; sc[3 * i + k] = r[52 + k]; }
        for (int j = 0; j < KR; j++)
            for (int c = 0; c < 3; c++) shn[(size_t)3 * KR * i + 3 * j + c] = r[6 + c * KR + j];
        op[i] = r[51];
        for (int k = 0; k < 4; k++) rot[4 * i + k] = r[55 + k];
    }
    updateTensorFromHost(pos.data(), rot.data(), sc.data(), op.data(), sh0.data(), shn.data(), n);
    return true;
}

void GaussianTrainerScene::trainSetup() {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    int W = 0, H = 0;
    for (auto& v : I.views) { W = std::max(W, v.cam.width); H = std::max(H, v.cam.height); }
    ckr(dvs_rast_reserve(I.ctx, std::max(I.N, I.capacity), W, H, 0), I.ctx, "reserve");  // per-Gaussian records for every row refinement may add
    status_ = TrainingStatus::Preprocess_Done;
}

void GaussianTrainerScene::trainStep() {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (I.views.empty() || I.N == 0) throw std::runtime_error("gstrain: train_step without data");
    {   // time spent training: gaps between consecutive steps, pauses (> 2 s) dropped
        const auto now = std::chrono::steady_clock::now();
        if (I.have_last_step) {
            const double dt = std::chrono::duration<double>(now - I.last_step).count();
            if (dt < 2.0) I.train_seconds += dt;
        }
        I.last_step = now; I.have_last_step = true;
    }
    const int step = curIteration;
    // The step's batch: `batch_views` views on this GPU (gradients accumulate), times the ranks of a data-parallel run; rank r
    // takes views (step * world + r) * batch + j  (mod the view count) — dp.views_for_rank of the Python harness.
    const int B = I.batch_views, world = I.dp.world;
    const bool batched = B > 1 || world > 1;
    dvs_params P = I.P();
    dvs_grads G = I.G();
    // Steps run without any host synchronisation (DVS_FLAG_DEFER_CHECK; honoured once a synchronous forward has
    // sized the binning arena).  A late DVS_E_OVERFLOW means a deferred step overflowed the arena: its compositing and
    // backward kernels exited early, and the kernels queued behind them that would move the model on such a step — the
    // fused Adam update, the ADC statistics — read the rasterizer's device overflow word and do nothing (a zero gradient
    // would still move every parameter through the decaying first moment).  The MCMC regulariser only adds to the (unused)
    // gradients and the exploration noise is a zero-mean perturbation applied every step anyway.  The step is then redone.
    // (A batched / data-parallel step always applies its update: the overflowed view contributes zero gradients to the sum
    // and every rank must step identically.)
    // refinement window (densify.cu): warmupLength < step < refineStopIter; MCMC is strategy 1, ADC 0 and 2
    const bool refining = I.dws && step > config_.warmupLength && step < config_.refineStopIter;
    const bool mcmc = config_.densifyStrategy == 1;
    uint32_t cam_flags = 0u;
    if (!I.resync_next) cam_flags |= DVS_FLAG_DEFER_CHECK;
    I.resync_next = false;
    if (config_.mipAntiliased) cam_flags |= DVS_FLAG_ANTIALIAS;  // --mipAntiliased (main.cpp, docs/userGuide.md:58)
    if (config_.modelType == 1) cam_flags |= DVS_FLAG_MODEL_2DGS;  // --modelType 1: 2D Gaussian splatting (main.cpp:28, gs_train.cpp:68)
    uint32_t bwd_flags = 0u;
    if (refining && !mcmc) {  // ADC feeds on the screen-space gradient of every view
        G.mean2D = I.d_mean2D;
        if (config_.useAbsGrad) { G.mean2D_abs = I.d_mean2D_abs; bwd_flags |= DVS_FLAG_ABSGRAD; }
    }
    // visibleAdam steps only the rows the step's view saw; with several views per step there is no single view: dense Adam
    const bool sparse_adam = config_.visibleAdam && I.d_radii && !batched;
    const uint32_t* skip_word = batched ? nullptr : dvs_rast_device_overflow_word(I.ctx);
    ck(cudaMemsetAsync(I.d_loss, 0, sizeof(float), I.stream), "memset loss");
    for (int j = 0; j < B; j++) {
        const size_t vi = (((size_t)step * (size_t)world + (size_t)I.dp.rank) * (size_t)B + (size_t)j) % I.views.size();
        View& vw = I.views[vi];
        dvs_camera cam = vw.cam;
        cam.flags |= cam_flags;
        cam.sh_degree = std::min(I.max_degree, step / 1000);  // progressive SH degree (every 1000 iterations)
        const uint32_t flags_j = bwd_flags | (j > 0 ? DVS_FLAG_ACCUMULATE : 0u);
        if (j > 0 && refining && !mcmc) {  // the screen-space statistics are per view: start each view's from zero
            ck(cudaMemsetAsync(I.d_mean2D, 0, 2 * (size_t)I.N * sizeof(float), I.stream), "memset mean2D");
            if (config_.useAbsGrad) ck(cudaMemsetAsync(I.d_mean2D_abs, 0, 2 * (size_t)I.N * sizeof(float), I.stream), "memset mean2D_abs");
        }
        if (config_.enableBg) {  // evaluate the sky model for this camera; the rasterizer composites over it
            const size_t need = (size_t)3 * cam.width * cam.height;
            if (!I.d_sky) {
                ck(cudaMalloc(&I.d_sky, 4 * 3 * SKY_K * sizeof(float)), "cudaMalloc sky");
                ck(cudaMemsetAsync(I.d_sky, 0, 4 * 3 * SKY_K * sizeof(float), I.stream), "memset sky");
            }
            if (need > I.sky_img_cap) {
                ck(cudaStreamSynchronize(I.stream), "sync");
                cudaFree(I.d_bg_img); cudaFree(I.d_dbg);
                ck(cudaMalloc(&I.d_bg_img, need * sizeof(float)), "cudaMalloc sky image");
                ck(cudaMalloc(&I.d_dbg, need * sizeof(float)), "cudaMalloc sky gradient image");
                I.sky_img_cap = need;
            }
            sky_eval_kernel<<<148 * 4, 256, 0, I.stream>>>(sky_cam(cam), I.d_sky, I.d_bg_img);
            ckr(dvs_rast_set_background(I.ctx, I.d_bg_img), I.ctx, "set_background");
        }
        for (int attempt = 0;; attempt++) {
            int rc = dvs_rast_forward(I.ctx, &cam, I.N, &P, I.d_render, ((refining && !mcmc) || sparse_adam) ? I.d_radii : nullptr, I.stream);
            if (rc == DVS_E_OVERFLOW && attempt < 2) { cam.flags &= ~DVS_FLAG_DEFER_CHECK; continue; }
            ckr(rc, I.ctx, "forward");
            const size_t npix = (size_t)cam.width * cam.height;
            const float* target = vw.d_target;
            if (!target) {  // PackF32ToU8: the view's image is held as bytes
                if (3 * npix > I.target_f32_cap) {
                    ck(cudaStreamSynchronize(I.stream), "sync");
                    cudaFree(I.d_target_f32);
                    ck(cudaMalloc(&I.d_target_f32, 3 * npix * sizeof(float)), "cudaMalloc unpacked target");
                    I.target_f32_cap = 3 * npix;
                }
                unpack_u8_kernel<<<148 * 4, 256, 0, I.stream>>>(vw.d_target_u8, I.d_target_f32, 3 * npix);
                target = I.d_target_f32;
            }
            if (vw.d_mask) mask_blend_kernel<<<1184, 256, 0, I.stream>>>(I.d_render, target, vw.d_mask, npix);
            launch_photometric_loss(I.d_render, target, I.d_dLdpix, I.d_loss, I.d_scratch, cam.width, cam.height,
                                    std::min(1.f, std::max(0.f, config_.ssimWeight)), I.stream);
            if (vw.d_mask) mask_grad_kernel<<<1184, 256, 0, I.stream>>>(I.d_dLdpix, vw.d_mask, npix);
            // normalConsistencyLoss (3DGS model; from a quarter of the schedule on, at most iteration 7000 as in the 2DGS paper):
            // render depth / alpha / normal maps, add lambda * L_n to the loss, and backpropagate through the maps as well
            const bool ncl = config_.normalConsistencyLoss && config_.modelType == 0 && step >= std::min(7000, config_.numIters / 4);
            if (ncl) {
                if (5 * npix > I.ncl_cap) {
                    ck(cudaStreamSynchronize(I.stream), "sync");
                    cudaFree(I.d_ncl);
                    ck(cudaMalloc(&I.d_ncl, 10 * npix * sizeof(float)), "cudaMalloc normal-consistency maps");
                    I.ncl_cap = 5 * npix;
                }
                float *m_aux = I.d_ncl, *m_nrm = I.d_ncl + 2 * npix, *g_aux = I.d_ncl + 5 * npix, *g_nrm = I.d_ncl + 7 * npix;
                rc = dvs_rast_forward_aux(I.ctx, &P, m_aux, m_nrm, I.stream);
                ckr(rc, I.ctx, "forward_aux");
                launch_normal_consistency(cam.width, cam.height, cam.tanfovx, cam.tanfovy, m_aux, m_nrm, 0.05f, I.d_loss, g_aux, g_nrm, I.stream);
                rc = dvs_rast_backward_aux(I.ctx, &P, I.d_dLdpix, g_aux, g_nrm, &G, flags_j, I.stream);
            } else {
                rc = dvs_rast_backward(I.ctx, &P, I.d_dLdpix, &G, flags_j, I.stream);
            }
            if (rc == DVS_E_OVERFLOW && attempt < 2) { cam.flags &= ~DVS_FLAG_DEFER_CHECK; continue; }
            ckr(rc, I.ctx, "backward");
            break;
        }
        if (config_.enableBg) {  // dL/dsky += sum_p Y(dir_p) final_T(p) dL/dpix(p)
            ckr(dvs_rast_background_grad(I.ctx, I.d_dLdpix, I.d_dbg, I.stream), I.ctx, "background_grad");
            sky_grad_kernel<<<148 * 2, 256, 0, I.stream>>>(sky_cam(cam), I.d_dbg, I.d_sky + 3 * SKY_K);
        }
        if (refining && !mcmc)
            ck(dvs_densify::adc_accumulate(I.d_mean2D, config_.useAbsGrad ? I.d_mean2D_abs : nullptr, I.d_radii, I.d_accum,
                                           I.d_denom, I.N, I.stream, skip_word), "adc_accumulate");
    }
    // data parallel: ONE exchange step — the sum over ranks of the six live gradient ranges (one NCCL group = one fused launch)
    if (world > 1) {
        const size_t n = (size_t)I.N;
        I.dp.all_reduce_sum({{I.grads.quats(), 4 * n}, {I.grads.shN(), (size_t)3 * KR * n}, {I.grads.means(), 3 * n},
                             {I.grads.scales(), 3 * n}, {I.grads.sh0(), 3 * n}, {I.grads.opac(), n}}, I.stream, true);
    }
    if (world > 1 && config_.enableBg) I.dp.all_reduce_sum({{I.d_sky + 3 * SKY_K, (size_t)3 * SKY_K}}, I.stream);
    if (refining && mcmc)  // L1 regularisers of the MCMC strategy: 0.01 mean(opacity) + 0.01 mean(scale), once per step
        ck(dvs_densify::mcmc_regularise(I.T(I.params), I.T(I.grads), I.N, 0.01f, 0.01f, I.stream), "mcmc_regularise");
    // Adam, per-group learning rates (GaussianTrainConfig); position lr decays exponentially init -> final
    const float t = std::min(1.f, (float)step / (float)std::max(1, config_.numIters));
    const float lr_pos = std::exp((1.f - t) * std::log(config_.poslrInit) + t * std::log(config_.poslrFinal)) * I.scene_extent;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-15f;
    const float c1 = 1.f / (1.f - std::pow(b1, (float)(step + 1))), c2 = 1.f / (1.f - std::pow(b2, (float)(step + 1)));
    const bool visible_only = sparse_adam;
    {   // one launch for all six groups (per-group learning rates)
        const float lrs[6] = {config_.rotationlr, config_.featurelr / 20.f, lr_pos, config_.scalinglr, config_.featurelr, config_.opacitylr};
        launch_adam_fused(I.params, I.grads.flat, I.m1.flat, I.m2.flat, I.N, lrs, visible_only ? I.d_radii : nullptr, skip_word, b1, b2,
                          eps, c1, c2, I.stream);
    }
    if (config_.enableBg && I.d_sky)
        sky_adam_kernel<<<1, 32, 0, I.stream>>>(I.d_sky, I.d_sky + 3 * SKY_K, I.d_sky + 6 * SKY_K, I.d_sky + 9 * SKY_K, config_.featurelr, b1,
                                               b2, eps, c1, c2);
    if (refining) {
        const uint64_t seed = 0x5DEECE66Dull * (uint64_t)(step + 1);
        if (mcmc)  // exploration noise after the optimizer step: Sigma eps gate(opacity) noiselr lr_xyz
            ck(dvs_densify::mcmc_noise(I.T(I.params), I.N, config_.noiselr * lr_pos, seed, I.stream), "mcmc_noise");
        if (step % config_.refineEvery == 0) {
            const int64_t before = I.N;
            I.last_report = dvs_densify::RefineReport{};
            if (mcmc) {
                ck(dvs_densify::mcmc_refine(I.dws, I.T(I.params), I.T(I.m1), I.T(I.m2), &I.N, I.capacity, effectiveCapMax(config_),
                                            config_.min_opacity, seed, I.stream, config_.verbose ? &I.last_report : nullptr), "mcmc_refine");  // (the report costs a host sync)
            } else {
                if (world > 1)  // every rank accumulated the statistics of ITS views: refine on their sum, identically everywhere
                    I.dp.all_reduce_sum({{I.d_accum, (size_t)I.N}, {I.d_denom, (size_t)I.N}}, I.stream);
                const dvs_densify::AdcConfig ac{config_.growGrad2d, 0.01f, I.scene_extent, config_.pruneOpacity,
                                                config_.pruneScale3d, config_.revisedOpacity};
                ck(dvs_densify::adc_refine(I.dws, I.T(I.params), I.T(I.m1), I.T(I.m2), I.d_accum, I.d_denom, &I.N,
                                           I.capacity, effectiveCapMax(config_), ac, seed, I.stream, &I.last_report), "adc_refine");
                if (config_.resetAlphaEvery > 0 && step % config_.resetAlphaEvery == 0)
                    ck(dvs_densify::adc_reset_opacity(I.T(I.params), I.T(I.m1), I.T(I.m2), I.N, I.stream), "reset_opacity");
            }
            if (I.N != before) I.resync_next = true;
            if (config_.verbose)
                std::fprintf(stderr, "gstrain: refine @%d: N %lld -> %lld (dead %lld relocated %lld added %lld grown %lld pruned %lld)\n",
                             step, (long long)before, (long long)I.N, (long long)I.last_report.dead,
                             (long long)I.last_report.relocated, (long long)I.last_report.added,
                             (long long)I.last_report.cloned, (long long)I.last_report.pruned);
        }
    }
    if (step % 100 == 0 || config_.verbose) {  // the reported loss: mean over the step's views (all ranks')
        if (world > 1) I.dp.all_reduce_sum({{I.d_loss, 1}}, I.stream);
        ck(cudaMemcpyAsync(I.h_loss, I.d_loss, sizeof(float), cudaMemcpyDeviceToHost, I.stream), "loss D2H");
        ck(cudaStreamSynchronize(I.stream), "sync");
        loss_ = *I.h_loss / (float)(B * world);
    }
    ck(cudaGetLastError(), "train_step kernels");
    status_ = TrainingStatus::Training;
    curIteration++;
}

// ---- model writers: csrc/model_io.cpp behind include/dvs_model_io.h — PLY, .splat, compressed
// PLY, .dvsplat and .spz, byte-identical to external/tinygsplat's writers.  Dispatch by extension like
// diverse/source/assets/gaussian_model.cpp:439-463; anything else is written as PLY.
static bool write_model(const std::string& path, int64_t N, const float* pos, const float* sh0, const float* shn,
                        const float* op, const float* sc, const float* rot, bool antialiased) {
    int fmt = dvs_model_format_from_path(path.c_str());
    if (fmt == DVS_FMT_AUTO) fmt = DVS_FMT_PLY;
    const int rc = dvs_model_write(path.c_str(), fmt, N, pos, sh0, shn, op, sc, rot, nullptr,
                                   antialiased ? DVS_IO_ANTIALIASED : 0u);
    if (rc != 0) std::fprintf(stderr, "gstrain: %s\n", dvs_model_io_last_error());
    return rc == 0;
}

void GaussianTrainerScene::saveGaussianModel() {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (config_.modelPath.empty() || I.N == 0) return;
    const auto pos = getGaussianPositionCpu(), sh0 = getGaussianSH0Cpu(), shn = getGaussianSHNCpu();
    const auto op = getGaussianOpcaitiesCpu(), sc = getGaussianScalingsCpu(), rot = getGaussianRotationsCpu();
    if (!write_model(config_.modelPath, I.N, pos.data(), sh0.data(), shn.data(), op.data(), sc.data(), rot.data(),
                     config_.mipAntiliased))
        std::fprintf(stderr, "gstrain: cannot write %s\n", config_.modelPath.c_str());
}

void GaussianTrainerScene::exportMesh(const std::string&) {
    std::fprintf(stderr, "gstrain: mesh export is outside the rasterizer hot path (SURVEY.md section 8) - skipped\n");
}

// ---- trainer -> viewer hand-off: one fused pack kernel + one 104 B/Gaussian copy
static size_t vp_offset_colors(int64_t cap) { return (size_t)cap * DVS_VP_GAUSSIAN_BYTES; }
static size_t vp_offset_sh(int64_t cap) { return vp_offset_colors(cap) + (size_t)cap * DVS_VP_COLOR_BYTES; }
static size_t vp_offset_bbox(int64_t cap) { return vp_offset_sh(cap) + (size_t)cap * DVS_VP_SH_BYTES; }
static size_t vp_bytes(int64_t cap) { return vp_offset_bbox(cap) + 32; }

void GaussianTrainerScene::requestViewerPack() {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (I.N <= 0) return;
    if (!I.vp_stream) {
        ck(cudaStreamCreateWithFlags(&I.vp_stream, cudaStreamNonBlocking), "viewer stream");
        ck(cudaEventCreateWithFlags(&I.vp_packed, cudaEventDisableTiming), "viewer event");
        for (auto& v : I.vp) ck(cudaEventCreateWithFlags(&v.ready, cudaEventDisableTiming), "viewer event");
    }
    const int64_t need = std::max(I.N, I.capacity);
    if (need > I.vp_cap) {  // (re)size both slots; rows are 8 / 16-byte records, offsets stay 16-byte aligned for cap % 2 == 0
        ck(cudaStreamSynchronize(I.vp_stream), "viewer sync");
        const int64_t cap = (need + 1) / 2 * 2;
        for (auto& v : I.vp) {
            cudaFree(v.d); if (v.h) cudaFreeHost(v.h);
            v.d = nullptr; v.h = nullptr; v.requested = false; v.count = 0;
            ck(cudaMalloc(&v.d, vp_bytes(cap)), "cudaMalloc viewer pack");
            ck(cudaMallocHost(&v.h, vp_bytes(cap)), "cudaMallocHost viewer pack");
        }
        I.vp_cap = cap;
        I.vp_last = -1;
    }
    auto& v = I.vp[I.vp_next];
    // the slot's previous copy (two requests ago) must have left the device staging before it is overwritten
    if (v.requested) ck(cudaStreamWaitEvent(I.stream, v.ready, 0), "viewer wait");
    const dvs_params P = I.P();
    const int rc = dvs_viewer_pack(P.means3D, P.scales, P.quats, P.opacities, P.sh0, P.shN, I.N, v.d, v.d + vp_offset_colors(I.vp_cap),
                                   v.d + vp_offset_sh(I.vp_cap), reinterpret_cast<uint32_t*>(v.d + vp_offset_bbox(I.vp_cap)), I.stream);
    ck((cudaError_t)rc, "dvs_viewer_pack");
    ck(cudaEventRecord(I.vp_packed, I.stream), "viewer record");
    ck(cudaStreamWaitEvent(I.vp_stream, I.vp_packed, 0), "viewer wait");
    // four contiguous pieces: only the live rows of each buffer travel
    const size_t offs[4] = {0, vp_offset_colors(I.vp_cap), vp_offset_sh(I.vp_cap), vp_offset_bbox(I.vp_cap)};
    const size_t lens[4] = {(size_t)I.N * DVS_VP_GAUSSIAN_BYTES, (size_t)I.N * DVS_VP_COLOR_BYTES, (size_t)I.N * DVS_VP_SH_BYTES, 24};
    for (int k = 0; k < 4; k++)
        ck(cudaMemcpyAsync(v.h + offs[k], v.d + offs[k], lens[k], cudaMemcpyDeviceToHost, I.vp_stream), "viewer D2H");
    ck(cudaEventRecord(v.ready, I.vp_stream), "viewer record");
    v.count = I.N;
    v.iteration = curIteration;
    v.requested = true;
    I.vp_last = I.vp_next;
    I.vp_next ^= 1;
}

bool GaussianTrainerScene::acquireViewerPack(GaussianViewerPack& out, bool wait) {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (I.vp_last < 0) return false;
    int slot = I.vp_last;
    if (wait) {
        ck(cudaEventSynchronize(I.vp[slot].ready), "viewer sync");
    } else if (cudaEventQuery(I.vp[slot].ready) != cudaSuccess) {
        slot ^= 1;  // the newest is still in flight: fall back to the one before, if it exists and has landed
        if (!I.vp[slot].requested || cudaEventQuery(I.vp[slot].ready) != cudaSuccess) return false;
    }
    const auto& v = I.vp[slot];
    out.gaussians = v.h;
    out.colors = v.h + vp_offset_colors(I.vp_cap);
    out.sh = v.h + vp_offset_sh(I.vp_cap);
    out.count = v.count;
    out.iteration = v.iteration;
    dvs_viewer_pack_decode_bbox(reinterpret_cast<const uint32_t*>(v.h + vp_offset_bbox(I.vp_cap)), out.bboxMin, out.bboxMax);
    return true;
}

int64_t GaussianTrainerScene::getNumGaussians() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->N;
}
std::vector<float> GaussianTrainerScene::getGaussianPositionCpu() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->download(impl_->params.means(), 3 * impl_->N);
}
std::vector<float> GaussianTrainerScene::getGaussianSH0Cpu() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->download(impl_->params.sh0(), 3 * impl_->N);
}
std::vector<float> GaussianTrainerScene::getGaussianSHNCpu() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->download(impl_->params.shN(), (size_t)45 * impl_->N);
}
std::vector<float> GaussianTrainerScene::getGaussianOpcaitiesCpu() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->download(impl_->params.opac(), impl_->N);
}
std::vector<float> GaussianTrainerScene::getGaussianScalingsCpu() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->download(impl_->params.scales(), 3 * impl_->N);
}
std::vector<float> GaussianTrainerScene::getGaussianRotationsCpu() const {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    return impl_->download(impl_->params.quats(), 4 * impl_->N);
}
int GaussianTrainerScene::getNumCameras() const { return (int)impl_->views.size(); }
std::array<float, 16> GaussianTrainerScene::getCameraProjectionFlat(int i) const {
    std::array<float, 16> a; std::memcpy(a.data(), impl_->views.at(i).P, sizeof(float) * 16); return a;
}
std::array<float, 16> GaussianTrainerScene::getCameraView(int i) const {
    std::array<float, 16> a; std::memcpy(a.data(), impl_->views.at(i).cam.view, sizeof(float) * 16); return a;
}
void GaussianTrainerScene::getCameraPosXYZ(int i, float p[3]) const {
    for (int k = 0; k < 3; k++) p[k] = impl_->views.at(i).cam.campos[k];
}
// camera -> world rotation = R^T of the loaded world -> camera rows, as a unit quaternion (w, x, y, z)
void GaussianTrainerScene::getCameraRotationWXYZ(int i, float q[4]) const {
    const float* Rt = impl_->views.at(i).Rt;
    const float m[3][3] = {{Rt[0], Rt[4], Rt[8]}, {Rt[1], Rt[5], Rt[9]}, {Rt[2], Rt[6], Rt[10]}};  // transpose
    const float tr = m[0][0] + m[1][1] + m[2][2];
    if (tr > 0.f) {
        const float s = std::sqrt(tr + 1.f) * 2.f;
        q[0] = 0.25f * s; q[1] = (m[2][1] - m[1][2]) / s; q[2] = (m[0][2] - m[2][0]) / s; q[3] = (m[1][0] - m[0][1]) / s;
    } else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) {
        const float s = std::sqrt(1.f + m[0][0] - m[1][1] - m[2][2]) * 2.f;
        q[0] = (m[2][1] - m[1][2]) / s; q[1] = 0.25f * s; q[2] = (m[0][1] + m[1][0]) / s; q[3] = (m[0][2] + m[2][0]) / s;
    } else if (m[1][1] > m[2][2]) {
        const float s = std::sqrt(1.f + m[1][1] - m[0][0] - m[2][2]) * 2.f;
        q[0] = (m[0][2] - m[2][0]) / s; q[1] = (m[0][1] + m[1][0]) / s; q[2] = 0.25f * s; q[3] = (m[1][2] + m[2][1]) / s;
    } else {
        const float s = std::sqrt(1.f + m[2][2] - m[0][0] - m[1][1]) * 2.f;
        q[0] = (m[1][0] - m[0][1]) / s; q[1] = (m[0][2] + m[2][0]) / s; q[2] = (m[1][2] + m[2][1]) / s; q[3] = 0.25f * s;
    }
}

// ---- the rest of the editor surface
void GaussianTrainerScene::resetGaussian() {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (I.init[3].empty()) return;
    ck(cudaStreamSynchronize(I.stream), "sync");
    I.set_model(I.init[0].data(), I.init[1].data(), I.init[2].data(), I.init[3].data(), I.init[4].data(), I.init[5].data(),
                (int64_t)I.init[3].size());
    I.zero_optimizer_state();
    I.resync_next = true;
    I.train_seconds = 0.0; I.have_last_step = false;
    curIteration = 0;
    loss_ = 0.f;
    status_ = TrainingStatus::Preprocess_Done;
}
void GaussianTrainerScene::setDensifyStrategy(int strategy) {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    config_.densifyStrategy = std::min(2, std::max(0, strategy));
    if (I.d_accum) {  // the ADC statistics restart with the strategy
        ck(cudaMemsetAsync(I.d_accum, 0, I.capacity * sizeof(float), I.stream), "memset accum");
        ck(cudaMemsetAsync(I.d_denom, 0, I.capacity * sizeof(float), I.stream), "memset denom");
    }
}
float GaussianTrainerScene::getProgressOnCurrentPhase() const {
    switch (status_) {
        case TrainingStatus::Training: return std::min(1.f, (float)curIteration / (float)std::max(1, config_.numIters));
        case TrainingStatus::Training_Done: case TrainingStatus::Preprocess_Done: return 1.f;
        default: return 0.f;
    }
}
std::string GaussianTrainerScene::getCurrentTrainingPhaseName() const {
    switch (status_) {
        case TrainingStatus::Loading_Prepare: return "Preparing";
        case TrainingStatus::Loading_Data: return "Loading data";
        case TrainingStatus::Colmap_Sfm: return "Structure from motion";
        case TrainingStatus::Preprocess_Done: return "Ready";
        case TrainingStatus::Training: return "Training";
        case TrainingStatus::Training_Done: return "Done";
        case TrainingStatus::GS2Mesh: return "Extracting mesh";
        case TrainingStatus::Loading_Failed: return "Failed";
    }
    return "";
}
float GaussianTrainerScene::getTrainingElpasedTime() const { return (float)impl_->train_seconds; }
float GaussianTrainerScene::getEstimateTrainingTime() const {
    if (curIteration <= 0) return 0.f;
    const int left = std::max(0, config_.numIters - curIteration);
    return (float)(impl_->train_seconds / (double)curIteration * (double)left);
}
void GaussianTrainerScene::updateTensorFromHost(const float* pos, const float* rot, const float* scale, const float* opacity,
                                                const float* sh0, const float* shn, int64_t n) {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    if (n <= 0 || !pos || !rot || !scale || !opacity || !sh0 || !shn) throw std::runtime_error("gstrain: updateTensorFromHost with an empty model");
    ck(cudaStreamSynchronize(I.stream), "sync");
    if (n > I.capacity) {  // the edited model outgrew the arenas: re-allocate (pointers change, so nothing may be in flight)
        ck(cudaDeviceSynchronize(), "sync");
        I.release_model();
        I.allocate(plannedCapacity(n));
        ckr(dvs_rast_reserve(I.ctx, I.capacity, 0, 0, 0), I.ctx, "reserve");
        I.vp_last = -1;  // snapshots of the old model are stale
    }
    I.set_model(pos, scale, rot, opacity, sh0, shn, n);
    I.zero_optimizer_state();
    I.resync_next = true;
}
const std::vector<GsPoint3D>& GaussianTrainerScene::getPoints3D(int) const { return impl_->points3d; }
GsImageView GaussianTrainerScene::getSplatImageView(int id) {
    std::lock_guard<std::recursive_mutex> lock(impl_->mu);
    auto& I = *impl_;
    View& v = I.views.at((size_t)id);
    const size_t P = (size_t)v.cam.width * v.cam.height;
    if (v.rgba.empty()) {
        std::vector<float> chw;
        if (v.d_target) {
            chw = I.download(v.d_target, 3 * P);
        } else {  // PackF32ToU8: the bytes themselves
            std::vector<uint8_t> b(3 * P);
            cudaDeviceSynchronize();
            cudaMemcpy(b.data(), v.d_target_u8, 3 * P, cudaMemcpyDeviceToHost);
            chw.resize(3 * P);
            for (size_t k = 0; k < 3 * P; k++) chw[k] = b[k] / 255.f;
        }
        v.rgba.resize(4 * P);
        for (size_t p = 0; p < P; p++) {
            for (int c = 0; c < 3; c++) v.rgba[4 * p + c] = (uint8_t)std::lround(255.f * std::min(1.f, std::max(0.f, chw[c * P + p])));
            v.rgba[4 * p + 3] = 255;
        }
    }
    GsImageView out;
    out.width = v.cam.width; out.height = v.cam.height; out.data = v.rgba.data(); out.name = v.name;
    return out;
}
// cameras.json in the layout the public 3DGS tooling writes: one object per view with the camera centre and the
// camera -> world rotation rows
bool GaussianTrainerScene::saveCameraDatas(const std::string& jsonPath) const {
    std::ofstream f(jsonPath);
    if (!f.good()) return false;
    f.precision(9);
    f << "[";
    for (size_t i = 0; i < impl_->views.size(); i++) {
        const View& v = impl_->views[i];
        f << (i ? ",\n " : "\n ") << "{\"id\": " << i << ", \"img_name\": \"" << v.name << "\", \"width\": " << v.cam.width
          << ", \"height\": " << v.cam.height << ", \"position\": [" << v.cam.campos[0] << ", " << v.cam.campos[1] << ", "
          << v.cam.campos[2] << "], \"rotation\": [";
        for (int r = 0; r < 3; r++)  // row r of R^T
            f << (r ? ", " : "") << "[" << v.Rt[r] << ", " << v.Rt[4 + r] << ", " << v.Rt[8 + r] << "]";
        f << "], \"fx\": " << v.fx << ", \"fy\": " << v.fy << "}";
    }
    f << "\n]\n";
    return f.good();
}
bool GaussianTrainerScene::exportSparsePointCloud(const std::string& plyPath) const {
    std::ofstream f(plyPath, std::ios::binary);
    if (!f.good()) return false;
    const auto& pts = impl_->points3d;
    f << "ply\nformat binary_little_endian 1.0\nelement vertex " << pts.size()
      << "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n";
    for (const auto& p : pts) {
        f.write(reinterpret_cast<const char*>(&p.x), 12);
        f.write(reinterpret_cast<const char*>(&p.r), 3);
    }
    return f.good();
}
void GaussianTrainerScene::updateFocusRegion(const GsVec3& position, const GsVec3& rotationDegrees, const GsVec3& scale) {
    focus_region_position = position; focus_region_rotation = rotationDegrees; focus_region_scale = scale;
}
void GaussianTrainerScene::getFocusRegionMinMax(float mn[3], float mx[3]) const {
    const auto& pts = impl_->points3d;
    for (int k = 0; k < 3; k++) { mn[k] = pts.empty() ? -1.f : 3.4e38f; mx[k] = pts.empty() ? 1.f : -3.4e38f; }
    for (const auto& p : pts) {
        const float v[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], v[k]); mx[k] = std::max(mx[k], v[k]); }
    }
}
void GaussianTrainerScene::getFocusRegionTransformFlat(float m[16]) const {
    const float d2r = 3.14159265358979f / 180.f;
    const float cx = std::cos(focus_region_rotation.x * d2r), sx = std::sin(focus_region_rotation.x * d2r);
    const float cy = std::cos(focus_region_rotation.y * d2r), sy = std::sin(focus_region_rotation.y * d2r);
    const float cz = std::cos(focus_region_rotation.z * d2r), sz = std::sin(focus_region_rotation.z * d2r);
    const float R[3][3] = {{cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx},
                           {sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx},
                           {-sy, cy * sx, cy * cx}};  // Rz * Ry * Rx
    const float sc[3] = {focus_region_scale.x, focus_region_scale.y, focus_region_scale.z};
    const float t[3] = {focus_region_position.x, focus_region_position.y, focus_region_position.z};
    for (int c = 0; c < 3; c++) {
        for (int r = 0; r < 3; r++) m[4 * c + r] = R[r][c] * sc[c];
        m[4 * c + 3] = 0.f;
    }
    m[12] = t[0]; m[13] = t[1]; m[14] = t[2]; m[15] = 1.f;
}

// -------------------------------------------------------------------------------------------------
// The nine plugin symbols (gs_train.cpp:24,105-110,144-150,178-179) + the two the loader probes
// (plugin.cpp:97,110: get_description / create_instance — logged, not fatal, if missing).
extern "C" {
// test hook: `steps` fused-Adam updates (the trainer's own kernel and arena layout) of device arenas laid out for `capacity`
// rows with `N` live ones; lrs[] in arena order quats | shN | means | scales | sh0 | opac; returns the arena size in floats
GS_EXPORT int64_t gstrain_test_adam(float* params, const float* grads, float* m1, float* m2, int64_t N, int64_t capacity,
                                    const float* lrs, float b1, float b2, float eps, int first_step, int steps,
                                    const int32_t* radii, const uint32_t* skip, void* stream) {
    Arena lay;
    lay.layout(capacity);
    if (!params) return (int64_t)lay.total;
    lay.flat = params;
    for (int t = first_step; t < first_step + steps; t++) {
        const float c1 = 1.f / (1.f - std::pow(b1, (float)(t + 1))), c2 = 1.f / (1.f - std::pow(b2, (float)(t + 1)));
        launch_adam_fused(lay, grads, m1, m2, N, lrs, radii, skip, b1, b2, eps, c1, c2, static_cast<cudaStream_t>(stream));
    }
    return cudaGetLastError() == cudaSuccess ? (int64_t)lay.total : -1;
}
// test hook of the normal-consistency loss: loss (accumulated into *loss), dL/d(depth, alpha) [2,H,W], dL/dnormal [3,H,W]
GS_EXPORT int gstrain_test_normal_consistency(int W, int H, float tanx, float tany, const float* aux, const float* nrm, float lambda,
                                              float* loss, float* d_aux, float* d_nrm, void* stream) {
    launch_normal_consistency(W, H, tanx, tany, aux, nrm, lambda, loss, d_aux, d_nrm, static_cast<cudaStream_t>(stream));
    return (int)cudaGetLastError();
}
// test hooks of the background model: bg[3,H,W] = sky(coef[9][3]) for a camera; dcoef[9][3] += sum_p Y(dir_p) dbg[:, p]
GS_EXPORT int gstrain_test_sky_eval(const dvs_camera* cam, const float* coef, float* bg, void* stream) {
    sky_eval_kernel<<<148 * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(sky_cam(*cam), coef, bg);
    return (int)cudaGetLastError();
}
GS_EXPORT int gstrain_test_sky_grad(const dvs_camera* cam, const float* dbg, float* dcoef, void* stream) {
    sky_grad_kernel<<<148 * 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(sky_cam(*cam), dbg, dcoef);
    return (int)cudaGetLastError();
}
// float offsets of the six tensors inside an arena of `capacity` rows (arena order), for tests and tools
GS_EXPORT void gstrain_arena_offsets(int64_t capacity, int64_t out[6]) {
    Arena lay;
    lay.layout(capacity);
    const size_t o[6] = {lay.off_quats, lay.off_shN, lay.off_means, lay.off_scales, lay.off_sh0, lay.off_opac};
    for (int k = 0; k < 6; k++) out[k] = (int64_t)o[k];
}
GS_EXPORT void gstrain_init() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        std::fprintf(stderr, "gstrain_init: no CUDA device - this plugin has no CPU path\n");
}
GS_EXPORT void* create_splat(const GaussianTrainConfig& config, int loadItr) { return new GaussianTrainerScene(config, loadItr); }
GS_EXPORT bool load_train_data(GaussianTrainerScene* scene, const std::string& path) { return scene && scene->loadTrainData(path); }
GS_EXPORT void train_step(GaussianTrainerScene* scene) { scene->trainStep(); }
GS_EXPORT void save_splat_model(GaussianTrainerScene* scene) { scene->saveGaussianModel(); }
GS_EXPORT void export_mesh(GaussianTrainerScene* scene) { scene->exportMesh(""); }
GS_EXPORT void delete_splat(GaussianTrainerScene* scene) { delete scene; }
GS_EXPORT int get_cur_step(GaussianTrainerScene* scene) { return scene->getCurrentIterations(); }
GS_EXPORT void gstrain_destroy() { cudaDeviceSynchronize(); }
// test hook (device pointers): the trainer's photometric loss and its gradient; scratch = 9*W*H floats, *loss zeroed by the caller
GS_EXPORT void gstrain_photometric_loss(const float* render, const float* target, float* dL_dpix, float* loss,
                                        float* scratch, int W, int H, float ssim_weight, void* stream) {
    launch_photometric_loss(render, target, dL_dpix, loss, scratch, W, H, ssim_weight, static_cast<cudaStream_t>(stream));
}
// test hook (device pointers): the same with a per-pixel mask [H,W] (useMask); `render` is overwritten by the blended image
GS_EXPORT void gstrain_masked_photometric_loss(float* render, const float* target, const float* mask, float* dL_dpix,
                                               float* loss, float* scratch, int W, int H, float ssim_weight, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t npix = (size_t)W * H;
    mask_blend_kernel<<<1184, 256, 0, st>>>(render, target, mask, npix);
    launch_photometric_loss(render, target, dL_dpix, loss, scratch, W, H, ssim_weight, st);
    mask_grad_kernel<<<1184, 256, 0, st>>>(dL_dpix, mask, npix);
}
// test hook (host pointers): the model writers without a trainer / GPU
GS_EXPORT int gstrain_write_model(const char* path, long long N, const float* pos, const float* sh0, const float* shn,
                                  const float* op, const float* sc, const float* rot) {
    return write_model(path, N, pos, sh0, shn, op, sc, rot, false) ? 0 : -1;
}
GS_EXPORT const char* get_description() { return "gstrain: B200-native 3DGS trainer plugin (divshot_b200)"; }
GS_EXPORT void* create_instance() { return nullptr; }
}
