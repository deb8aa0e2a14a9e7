/*
 * dvs_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, "parity unpinned" — see dvs_oracle.h).
 *
 * Restates the credited 3DGS tile rasterizer (EXTERNAL SPEC of
 * graphdeco-inria/diff-gaussian-rasterization, credited at /root/reference/README.md:95)
 * in plain C.  Every function cites the in-tree corroboration it follows.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -mavx2 -fopenmp -fPIC -shared  (see oracle/Makefile).
 * -ffp-contract=off + explicit fmaf() makes the per-Gaussian arithmetic (A1) a literal
 * operation sequence; the CUDA kernels mirror that sequence (compiled -fmad=false), which is
 * what makes radii / tiles_touched / depth keys bit-exact.
 */
#include "dvs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* float -> int32 with the semantics of PTX cvt.rzi.s32.f32 (truncate, saturate, NaN -> 0) */
static inline int32_t f2i_rz(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (int32_t)0x80000000;
    return (int32_t)x;
}
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

/* Deterministic fp32 exp: Cody-Waite reduction + degree-5 minimax on r^2 (Cephes expf
 * coefficients), every step a single IEEE op or fmaf.  <= 2 ulp.  Used only by the
 * activations exp(log_scale) and sigmoid(logit) (gaussian_model.cpp:145-157). */
/* number of OpenMP threads used by every parallel region of the oracle (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

float orc_expf(float x) {
    x = fminf(fmaxf(x, -87.0f), 88.0f);
    float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r) + 1.0f;
    int32_t ni = (int32_t)n;
    return e * u2f((uint32_t)(ni + 127) << 23);
}

/* SH constants: gsplat_sh.hlsl:42-62, SH_C0 gaussian_model.cpp:128 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    float s[3];  /* activated scale * scale_modifier */
    float q[4];  /* unit quaternion r,x,y,z */
    float qlen;  /* |q_raw| */
    float o;     /* activated opacity */
} act_t;

static inline void activate(const orc_camera* cam, const float* scales, const float* quats,
                            const float* opac, int i, act_t* a) {
    if (cam->flags & ORC_FLAG_INPUT_ACTIVATED) {
        for (int k = 0; k < 3; k++) a->s[k] = cam->scale_modifier * scales[3 * i + k];
        for (int k = 0; k < 4; k++) a->q[k] = quats[4 * i + k];
        a->qlen = 1.0f;
        a->o = opac[i];
    } else {
        for (int k = 0; k < 3; k++) a->s[k] = cam->scale_modifier * orc_expf(scales[3 * i + k]);
        float q0 = quats[4 * i], q1 = quats[4 * i + 1], q2 = quats[4 * i + 2], q3 = quats[4 * i + 3];
        float n2 = fmaf(q0, q0, fmaf(q1, q1, fmaf(q2, q2, q3 * q3)));
        float len = sqrtf(n2);
        float inv = 1.0f / len;
        a->q[0] = q0 * inv; a->q[1] = q1 * inv; a->q[2] = q2 * inv; a->q[3] = q3 * inv;
        a->qlen = len;
        a->o = 1.0f / (1.0f + orc_expf(-opac[i]));
    }
}

/* R(q): gsplat_intersect.hlsl:71-82 / gsplat_vs.hlsl:189-200 (row-major rows listed there) */
static inline void quat_to_R(const float* q, float R[3][3]) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = fmaf(-2.0f, fmaf(z, z, y * y), 1.0f);
    R[0][1] = 2.0f * fmaf(x, y, -(r * z));
    R[0][2] = 2.0f * fmaf(x, z, r * y);
    R[1][0] = 2.0f * fmaf(x, y, r * z);
    R[1][1] = fmaf(-2.0f, fmaf(z, z, x * x), 1.0f);
    R[1][2] = 2.0f * fmaf(y, z, -(r * x));
    R[2][0] = 2.0f * fmaf(x, z, -(r * y));
    R[2][1] = 2.0f * fmaf(y, z, r * x);
    R[2][2] = fmaf(-2.0f, fmaf(y, y, x * x), 1.0f);
}

/* Sigma = (R S)(R S)^T, 6 upper-triangular values: gsplat_intersect.hlsl:61-95 */
static inline void cov3d_from(const float s[3], float R[3][3], float M[3][3], float c6[6]) {
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) M[i][k] = R[i][k] * s[k];
    c6[0] = fmaf(M[0][0], M[0][0], fmaf(M[0][1], M[0][1], M[0][2] * M[0][2]));
    c6[1] = fmaf(M[0][0], M[1][0], fmaf(M[0][1], M[1][1], M[0][2] * M[1][2]));
    c6[2] = fmaf(M[0][0], M[2][0], fmaf(M[0][1], M[2][1], M[0][2] * M[2][2]));
    c6[3] = fmaf(M[1][0], M[1][0], fmaf(M[1][1], M[1][1], M[1][2] * M[1][2]));
    c6[4] = fmaf(M[1][0], M[2][0], fmaf(M[1][1], M[2][1], M[1][2] * M[2][2]));
    c6[5] = fmaf(M[2][0], M[2][0], fmaf(M[2][1], M[2][1], M[2][2] * M[2][2]));
}

/* t = View * p (rows 0..2), m[4c+r]: gsplat_viewz_cs.hlsl:195-203 */
static inline void xform43(const float* m, const float* p, float t[3]) {
    for (int r = 0; r < 3; r++)
        t[r] = fmaf(m[r], p[0], fmaf(m[4 + r], p[1], fmaf(m[8 + r], p[2], m[12 + r])));
}
static inline void xform44(const float* m, const float* p, float h[4]) {
    for (int r = 0; r < 4; r++)
        h[r] = fmaf(m[r], p[0], fmaf(m[4 + r], p[1], fmaf(m[8 + r], p[2], m[12 + r])));
}

typedef struct {
    float fx, fy, limx, limy;
    float tx, ty, tz;       /* clamped t.x, t.y; t.z */
    float txtz, tytz;       /* unclamped ratios */
    float J00, J02, J11, J12;
    float T[2][3];
    float a, b, c;          /* cov2D with +0.3 */
    float a0, c0;           /* cov2D diagonal before the +0.3 low-pass */
} ewa_t;

/* EWA cov2D: gsplat_intersect.hlsl:96-134 (1.3*tanfov clamp, J, W, +0.3 low-pass) */
static inline void ewa_cov2d(const orc_camera* cam, const float t[3], const float c6[6], ewa_t* e) {
    const float* V = cam->view;
    e->fx = (float)cam->width / (2.0f * cam->tanfovx);
    e->fy = (float)cam->height / (2.0f * cam->tanfovy);
    e->limx = 1.3f * cam->tanfovx;
    e->limy = 1.3f * cam->tanfovy;
    e->tz = t[2];
    e->txtz = t[0] / t[2];
    e->tytz = t[1] / t[2];
    e->tx = fminf(e->limx, fmaxf(-e->limx, e->txtz)) * t[2];
    e->ty = fminf(e->limy, fmaxf(-e->limy, e->tytz)) * t[2];
    float tz2 = t[2] * t[2];
    e->J00 = e->fx / t[2];
    e->J02 = -(e->fx * e->tx) / tz2;
    e->J11 = e->fy / t[2];
    e->J12 = -(e->fy * e->ty) / tz2;
    /* W_rc = V[4c+r]; T = J W (2x3) */
    for (int j = 0; j < 3; j++) {
        float W0j = V[4 * j + 0], W1j = V[4 * j + 1], W2j = V[4 * j + 2];
        e->T[0][j] = fmaf(e->J00, W0j, e->J02 * W2j);
        e->T[1][j] = fmaf(e->J11, W1j, e->J12 * W2j);
    }
    /* Sigma symmetric 3x3 */
    float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float U[2][3];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++)
            U[i][j] = fmaf(e->T[i][0], S[0][j], fmaf(e->T[i][1], S[1][j], e->T[i][2] * S[2][j]));
    e->a0 = fmaf(U[0][0], e->T[0][0], fmaf(U[0][1], e->T[0][1], U[0][2] * e->T[0][2]));
    e->b = fmaf(U[0][0], e->T[1][0], fmaf(U[0][1], e->T[1][1], U[0][2] * e->T[1][2]));
    e->c0 = fmaf(U[1][0], e->T[1][0], fmaf(U[1][1], e->T[1][1], U[1][2] * e->T[1][2]));
    e->a = e->a0 + 0.3f;
    e->c = e->c0 + 0.3f;
}

/* SH basis values b[0..K) for unit direction d: gsplat_sh.hlsl:64-103 (signs and constants) */
static inline void sh_basis(int deg, const float d[3], float b[16]) {
    float x = d[0], y = d[1], z = d[2];
    b[0] = SH_C0;
    if (deg < 1) return;
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH_C2[0] * xy;
    b[5] = SH_C2[1] * yz;
    b[6] = SH_C2[2] * (fmaf(2.0f, zz, -xx) - yy);
    b[7] = SH_C2[3] * xz;
    b[8] = SH_C2[4] * (xx - yy);
    if (deg < 3) return;
    b[9] = SH_C3[0] * y * fmaf(3.0f, xx, -yy);
    b[10] = SH_C3[1] * xy * z;
    b[11] = SH_C3[2] * y * (fmaf(4.0f, zz, -xx) - yy);
    b[12] = SH_C3[3] * z * (fmaf(2.0f, zz, -(3.0f * xx)) - 3.0f * yy);
    b[13] = SH_C3[4] * x * (fmaf(4.0f, zz, -xx) - yy);
    b[14] = SH_C3[5] * z * (xx - yy);
    b[15] = SH_C3[6] * x * fmaf(-3.0f, yy, xx);
}

void orc_preprocess_fwd(const orc_camera* cam, int32_t N, const float* means3D, const float* scales,
                        const float* quats, const float* opacities, const float* sh0, const float* shN,
                        float* depth, int32_t* radii, float* mean2D, float* cov3D, float* conic_opacity,
                        float* rgb, uint8_t* clamped, uint32_t* tiles_touched, int32_t* rect) {
    const int gx = (cam->width + TILE - 1) / TILE, gy = (cam->height + TILE - 1) / TILE;
    const int deg = cam->sh_degree, K = (deg + 1) * (deg + 1), KR = cam->sh_rest_alloc;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        radii[i] = 0; tiles_touched[i] = 0; depth[i] = 0.0f;
        mean2D[2 * i] = mean2D[2 * i + 1] = 0.0f;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0.0f;
        for (int k = 0; k < 4; k++) { conic_opacity[4 * i + k] = 0.0f; rect[4 * i + k] = 0; }
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.0f; clamped[3 * i + k] = 0; }
        const float* p = means3D + 3 * (size_t)i;
        float t[3], h[4];
        xform43(cam->view, p, t);
        if (t[2] <= 0.2f) continue; /* B.1 step 1: only cull */
        xform44(cam->proj, p, h);
        float w_inv = 1.0f / (h[3] + 1e-7f); /* gsplat_viewz_cs.hlsl:197-199 */
        float ndcx = h[0] * w_inv, ndcy = h[1] * w_inv;
        act_t a; activate(cam, scales, quats, opacities, i, &a);
        float R[3][3], M[3][3], c6[6];
        quat_to_R(a.q, R);
        cov3d_from(a.s, R, M, c6);
        ewa_t e; ewa_cov2d(cam, t, c6, &e);
        float det = fmaf(e.a, e.c, -(e.b * e.b));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float cA = e.c * det_inv, cB = -e.b * det_inv, cC = e.a * det_inv;
        float mid = 0.5f * (e.a + e.c);
        float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float l1 = mid + sq, l2 = mid - sq;
        float rad_f = ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
        int32_t rad = f2i_rz(rad_f);
        /* ndc2Pix: gsplat_vs.hlsl:211-214 */
        float mx = fmaf(ndcx + 1.0f, (float)cam->width, -1.0f) * 0.5f;
        float my = fmaf(ndcy + 1.0f, (float)cam->height, -1.0f) * 0.5f;
        float radf = (float)rad;
        int32_t minx = imin(gx, imax(0, f2i_rz((mx - radf) * 0.0625f)));
        int32_t miny = imin(gy, imax(0, f2i_rz((my - radf) * 0.0625f)));
        int32_t maxx = imin(gx, imax(0, f2i_rz((mx + radf + 15.0f) * 0.0625f)));
        int32_t maxy = imin(gy, imax(0, f2i_rz((my + radf + 15.0f) * 0.0625f)));
        int64_t area = (int64_t)(maxx - minx) * (int64_t)(maxy - miny);
        if (area <= 0) continue;
        /* colour: direction from camera, SH eval, +0.5, clamp at 0 (gsplat_sh.hlsl:64-126) */
        float dir[3] = {p[0] - cam->campos[0], p[1] - cam->campos[1], p[2] - cam->campos[2]};
        float len = sqrtf(fmaf(dir[0], dir[0], fmaf(dir[1], dir[1], dir[2] * dir[2])));
        float linv = 1.0f / len;
        dir[0] *= linv; dir[1] *= linv; dir[2] *= linv;
        float bas[16]; sh_basis(deg, dir, bas);
        for (int ch = 0; ch < 3; ch++) {
            float acc = bas[0] * sh0[3 * (size_t)i + ch];
            for (int k = 1; k < K; k++) acc = fmaf(bas[k], shN[3 * ((size_t)i * KR + (k - 1)) + ch], acc);
            acc += 0.5f;
            clamped[3 * i + ch] = acc < 0.0f;
            rgb[3 * i + ch] = fmaxf(acc, 0.0f);
        }
        depth[i] = t[2];
        radii[i] = rad;
        mean2D[2 * i] = mx; mean2D[2 * i + 1] = my;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = c6[k];
        conic_opacity[4 * i] = cA; conic_opacity[4 * i + 1] = cB; conic_opacity[4 * i + 2] = cC;
        float opac = a.o;
        if (cam->flags & ORC_FLAG_ANTIALIAS) { /* mip-splatting compensation: gsplat_vs.hlsl:296-301,373 */
            const float det0 = fmaf(e.a0, e.c0, -(e.b * e.b));
            opac = a.o * sqrtf(fmaxf(0.0f, det0 / det));
        }
        conic_opacity[4 * i + 3] = opac;
        tiles_touched[i] = (uint32_t)area;
        rect[4 * i] = minx; rect[4 * i + 1] = miny; rect[4 * i + 2] = maxx; rect[4 * i + 3] = maxy;
    }
}

int64_t orc_scan_tiles(int32_t N, const uint32_t* tiles_touched, uint32_t* point_offsets) {
    uint64_t run = 0;
    for (int i = 0; i < N; i++) { run += tiles_touched[i]; point_offsets[i] = (uint32_t)run; }
    return (int64_t)run;
}

typedef struct { uint64_t key; uint32_t id; } orc_pair;
static int orc_pair_cmp(const void* a, const void* b) {
    const orc_pair* x = (const orc_pair*)a; const orc_pair* y = (const orc_pair*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

/* stable LSD radix sort of (key,value) pairs on the low `bits` bits (B.2: CUB SortPairs semantics) */
static void radix_sort_pairs(uint64_t* k, uint32_t* v, int64_t n, int bits) {
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    uint32_t* v2 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    for (int sh = 0; sh < bits; sh += 8) {
        size_t cnt[257]; memset(cnt, 0, sizeof cnt);
        for (int64_t i = 0; i < n; i++) cnt[((k[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; i++) { size_t p = cnt[(k[i] >> sh) & 255]++; k2[p] = k[i]; v2[p] = v[i]; }
        memcpy(k, k2, sizeof(uint64_t) * (size_t)n); memcpy(v, v2, sizeof(uint32_t) * (size_t)n);
    }
    free(k2); free(v2);
}

void orc_bin_sort(const orc_camera* cam, int32_t N, const float* depth, const int32_t* radii,
                  const int32_t* rect, const uint32_t* point_offsets, int64_t D, uint64_t* keys,
                  uint32_t* point_list, uint32_t* ranges) {
    const int gx = (cam->width + TILE - 1) / TILE, gy = (cam->height + TILE - 1) / TILE;
    const int T = gx * gy;
    /* A3 duplicateWithKeys: y outer, x inner; key = tile<<32 | depth bits; value = index */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        if (radii[i] <= 0) continue;
        int64_t off = i == 0 ? 0 : point_offsets[i - 1];
        for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; y++)
            for (int x = rect[4 * i]; x < rect[4 * i + 2]; x++) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key = (key << 32) | f2u(depth[i]);
                keys[off] = key; point_list[off] = (uint32_t)i; off++;
            }
    }
    /* A4: sort.  The credited design is one stable radix sort of all D (tile|depth) keys; the total order it
     * produces is (tile, depth bits, Gaussian index).  For the CPU baseline to use all host cores the same order is
     * produced here by a counting sort on the tile id (one serial O(D) pass) followed by independent per-tile sorts
     * on (depth bits, index), run in parallel over tiles.  orc_radix_check() below keeps the literal radix variant
     * for cross-checking in the tests. */
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)T);
    {
        uint32_t* cnt = (uint32_t*)calloc((size_t)T + 1, sizeof(uint32_t));
        for (int64_t j = 0; j < D; j++) cnt[(uint32_t)(keys[j] >> 32) + 1]++;
        for (int t = 0; t < T; t++) cnt[t + 1] += cnt[t];
        orc_pair* tmp = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(D > 0 ? D : 1));
        uint32_t* cur = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)T);
        memcpy(cur, cnt, sizeof(uint32_t) * (size_t)T);
        for (int64_t j = 0; j < D; j++) {
            const uint32_t t = (uint32_t)(keys[j] >> 32);
            tmp[cur[t]].key = keys[j]; tmp[cur[t]].id = point_list[j]; cur[t]++;
        }
#pragma omp parallel for schedule(dynamic, 8)
        for (int t = 0; t < T; t++) {
            const uint32_t b = cnt[t], e = cnt[t + 1];
            if (e > b) {
                qsort(tmp + b, e - b, sizeof(orc_pair), orc_pair_cmp);
                ranges[2 * t] = b; ranges[2 * t + 1] = e; /* A5 identifyTileRanges */
                for (uint32_t j = b; j < e; j++) { keys[j] = tmp[j].key; point_list[j] = tmp[j].id; }
            }
        }
        free(cnt); free(tmp); free(cur);
    }
}

/* The literal variant of A4/A5 (one stable LSD radix sort over all keys + range detection), kept so the tests can
 * check that the parallel per-tile sort above yields the identical lists. */
void orc_radix_check(int32_t T, int64_t D, uint64_t* keys, uint32_t* point_list, uint32_t* ranges) {
    int tb = 0; while ((1 << tb) < T) tb++; /* bits needed for tile ids < T */
    radix_sort_pairs(keys, point_list, D, 32 + tb + ((32 + tb) % 8 ? 8 - (32 + tb) % 8 : 0));
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)T);
    for (int64_t j = 0; j < D; j++) {
        uint32_t tile = (uint32_t)(keys[j] >> 32);
        if (j == 0 || tile != (uint32_t)(keys[j - 1] >> 32)) {
            ranges[2 * tile] = (uint32_t)j;
            if (j > 0) ranges[2 * (uint32_t)(keys[j - 1] >> 32) + 1] = (uint32_t)j;
        }
        if (j == D - 1) ranges[2 * tile + 1] = (uint32_t)D;
    }
}

/* A6: B.3; alpha falloff/clamp mirrored (different kernel scale) at gsplat_ps.hlsl:60-66,85 */
void orc_render_fwd(const orc_camera* cam, const uint32_t* ranges, const uint32_t* point_list,
                    const float* mean2D, const float* conic_opacity, const float* rgb, float* out_color,
                    float* final_T, uint32_t* n_contrib, uint8_t* fragile, int32_t threads) {
    const int W = cam->width, H = cam->height;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t P = (size_t)W * H;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int tx0 = (tile % gx) * TILE, ty0 = (tile / gx) * TILE;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx0 + lx, py = ty0 + ly;
                if (px >= W || py >= H) continue;
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, C[3] = {0, 0, 0};
                uint32_t contributor = 0, last = 0;
                uint8_t frag = 0;
                for (uint32_t j = r0; j < r1; j++) {
                    contributor++;
                    const uint32_t g = point_list[j];
                    const float dx = mean2D[2 * g] - pxf, dy = mean2D[2 * g + 1] - pyf;
                    const float cA = conic_opacity[4 * g], cB = conic_opacity[4 * g + 1];
                    const float cC = conic_opacity[4 * g + 2], o = conic_opacity[4 * g + 3];
                    const float power = -0.5f * (cA * dx * dx + cC * dy * dy) - cB * dx * dy;
                    const float mag = 0.5f * (fabsf(cA) * dx * dx + fabsf(cC) * dy * dy) + fabsf(cB * dx * dy);
                    if (fabsf(power) <= 4e-6f * mag && mag > 0.0f) frag = 1;
                    if (power > 0.0f) continue;
                    const float araw = o * expf(power);
                    const float alpha = fminf(0.99f, araw);
                    if (fabsf(araw - (1.0f / 255.0f)) <= (1.0f / 255.0f) * 3e-5f) frag = 1;
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (fabsf(test_T - 1e-4f) <= 1e-4f * 2e-3f) frag = 1;
                    if (test_T < 1e-4f) break;
                    const float w = alpha * T;
                    C[0] += rgb[3 * g] * w; C[1] += rgb[3 * g + 1] * w; C[2] += rgb[3 * g + 2] * w;
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                final_T[pix] = T;
                n_contrib[pix] = last;
                if (fragile) fragile[pix] = frag;
                for (int ch = 0; ch < 3; ch++) out_color[ch * P + pix] = C[ch] + T * cam->bg[ch];
            }
    }
}

/* Per-Gaussian sums are accumulated in double and rounded once: the credited upstream sums with fp32
 * atomics in a non-deterministic order, so every fp32 summation order is "the reference"; the double sum
 * is the centre of that family and keeps the oracle's own rounding noise out of the 1e-4 parity budget. */
static inline void atomic_addd(double* p, double v) {
#pragma omp atomic
    *p += v;
}

/* A7: B.4 reverse walk.  No in-tree corroboration (the viewer has no backward). */
void orc_render_bwd(const orc_camera* cam, int32_t N, const uint32_t* ranges, const uint32_t* point_list,
                    const float* mean2D, const float* conic_opacity, const float* rgb, const float* final_T,
                    const uint32_t* n_contrib, const float* dL_dpix, float* dL_dmean2D,
                    float* dL_dmean2D_abs, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                    int32_t threads) {
    const int W = cam->width, H = cam->height;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t P = (size_t)W * H;
    double* acc = (double*)calloc((size_t)(N > 0 ? N : 1) * 11, sizeof(double));
    double* a_m2 = acc;                      /* [N,2] */
    double* a_abs = acc + 2 * (size_t)N;     /* [N,2] */
    double* a_con = acc + 4 * (size_t)N;     /* [N,3] */
    double* a_op = acc + 7 * (size_t)N;      /* [N]   */
    double* a_col = acc + 8 * (size_t)N;     /* [N,3] */
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int tx0 = (tile % gx) * TILE, ty0 = (tile / gx) * TILE;
        if (r1 <= r0) continue;
        /* per-tile partial sums (11 per list entry), flushed with one atomic per (tile, splat, component): the
         * per-pixel atomics of the literal formulation do not scale beyond a few cores */
        double* loc = (double*)calloc((size_t)(r1 - r0) * 11, sizeof(double));
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx0 + lx, py = ty0 + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = final_T[pix];
                float T = T_final;
                const uint32_t last_contributor = n_contrib[pix];
                float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.0f;
                const float dp[3] = {dL_dpix[pix], dL_dpix[P + pix], dL_dpix[2 * P + pix]};
                uint32_t contributor = r1 - r0;
                for (uint32_t j = r1; j-- > r0;) {
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t g = point_list[j];
                    const float dx = mean2D[2 * g] - pxf, dy = mean2D[2 * g + 1] - pyf;
                    const float cA = conic_opacity[4 * g], cB = conic_opacity[4 * g + 1];
                    const float cC = conic_opacity[4 * g + 2], o = conic_opacity[4 * g + 3];
                    const float power = -0.5f * (cA * dx * dx + cC * dy * dy) - cB * dx * dy;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, o * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.0f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = rgb[3 * g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dp[ch];
                        loc[(size_t)(j - r0) * 11 + 8 + ch] += dchannel_dcolor * dp[ch];
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot = 0.0f;
                    for (int ch = 0; ch < 3; ch++) bg_dot += cam->bg[ch] * dp[ch];
                    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                    const float dL_dG = o * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * cA - gdy * cB;
                    const float dG_ddely = -gdy * cC - gdx * cB;
                    const float gmx = dL_dG * dG_ddelx * ddelx_dx, gmy = dL_dG * dG_ddely * ddely_dy;
                    double* lg = loc + (size_t)(j - r0) * 11;
                    lg[0] += gmx; lg[1] += gmy; lg[2] += fabsf(gmx); lg[3] += fabsf(gmy);
                    lg[4] += -0.5f * gdx * dx * dL_dG;
                    lg[5] += -gdx * dy * dL_dG; /* TOTAL off-diagonal gradient */
                    lg[6] += -0.5f * gdy * dy * dL_dG;
                    lg[7] += G * dL_dalpha;
                }
            }
        for (uint32_t j = r0; j < r1; j++) {
            const uint32_t g = point_list[j];
            const double* lg = loc + (size_t)(j - r0) * 11;
            if (lg[0] != 0.0) atomic_addd(&a_m2[2 * g], lg[0]);
            if (lg[1] != 0.0) atomic_addd(&a_m2[2 * g + 1], lg[1]);
            if (lg[2] != 0.0) atomic_addd(&a_abs[2 * g], lg[2]);
            if (lg[3] != 0.0) atomic_addd(&a_abs[2 * g + 1], lg[3]);
            for (int k = 0; k < 3; k++) if (lg[4 + k] != 0.0) atomic_addd(&a_con[3 * g + k], lg[4 + k]);
            if (lg[7] != 0.0) atomic_addd(&a_op[g], lg[7]);
            for (int k = 0; k < 3; k++) if (lg[8 + k] != 0.0) atomic_addd(&a_col[3 * g + k], lg[8 + k]);
        }
        free(loc);
    }
    for (size_t i = 0; i < (size_t)N; i++) {
        dL_dmean2D[2 * i] = (float)a_m2[2 * i]; dL_dmean2D[2 * i + 1] = (float)a_m2[2 * i + 1];
        if (dL_dmean2D_abs) { dL_dmean2D_abs[2 * i] = (float)a_abs[2 * i]; dL_dmean2D_abs[2 * i + 1] = (float)a_abs[2 * i + 1]; }
        for (int k = 0; k < 3; k++) { dL_dconic[3 * i + k] = (float)a_con[3 * i + k]; dL_dcolor[3 * i + k] = (float)a_col[3 * i + k]; }
        dL_dopacity[i] = (float)a_op[i];
    }
    free(acc);
}

/* A8: B.5.  Forward counterparts: gsplat_intersect.hlsl:61-134, gsplat_sh.hlsl:64-103.
 * Evaluated in double from the fp32 inputs and rounded once at the end: the chain conic -> cov2D -> T -> J / Sigma
 * -> (scale, rotation) cancels heavily, so two fp32 evaluation orders (with / without FMA contraction) of the
 * credited formulas differ by ~1e-4 relative on some elements; the double evaluation is the centre of that family
 * and keeps the oracle's own rounding out of the parity budget.  Nothing here is part of the bit-exact contract. */
void orc_preprocess_bwd(const orc_camera* cam, int32_t N, const float* means3D, const float* scales,
                        const float* quats, const float* opacities, const float* sh0, const float* shN,
                        const int32_t* radii, const uint8_t* clamped, const float* dL_dmean2D,
                        const float* dL_dconic, const float* dL_dopacity_act, const float* dL_dcolor,
                        float* dL_dmeans3D, float* dL_dscales, float* dL_dquats, float* dL_dopacities,
                        float* dL_dsh0, float* dL_dshN) {
    const int deg = cam->sh_degree, K = (deg + 1) * (deg + 1), KR = cam->sh_rest_alloc;
    double V[16], Pm[16];
    for (int k = 0; k < 16; k++) { V[k] = cam->view[k]; Pm[k] = cam->proj[k]; }
    (void)sh0;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < 3; k++) { dL_dmeans3D[3 * i + k] = 0; dL_dscales[3 * i + k] = 0; dL_dsh0[3 * i + k] = 0; }
        for (int k = 0; k < 4; k++) dL_dquats[4 * i + k] = 0;
        dL_dopacities[i] = 0;
        for (int k = 0; k < 3 * KR; k++) dL_dshN[3 * (size_t)KR * i + k] = 0;
        if (radii[i] <= 0) continue;
        const double p[3] = {means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1], means3D[3 * (size_t)i + 2]};
        /* activations */
        double s[3], q[4], qlen = 1.0, o;
        const int activated = cam->flags & ORC_FLAG_INPUT_ACTIVATED;
        if (activated) {
            for (int k = 0; k < 3; k++) s[k] = (double)cam->scale_modifier * scales[3 * i + k];
            for (int k = 0; k < 4; k++) q[k] = quats[4 * i + k];
            o = opacities[i];
        } else {
            for (int k = 0; k < 3; k++) s[k] = (double)cam->scale_modifier * exp((double)scales[3 * i + k]);
            double n2 = 0; for (int k = 0; k < 4; k++) n2 += (double)quats[4 * i + k] * quats[4 * i + k];
            qlen = sqrt(n2);
            for (int k = 0; k < 4; k++) q[k] = quats[4 * i + k] / qlen;
            o = 1.0 / (1.0 + exp(-(double)opacities[i]));
        }
        const double r = q[0], x = q[1], y = q[2], z = q[3];
        const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)},
                                {2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)},
                                {2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)}};
        double M[3][3], S[3][3];
        for (int a = 0; a < 3; a++) for (int k = 0; k < 3; k++) M[a][k] = R[a][k] * s[k];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) S[a][b] = M[a][0] * M[b][0] + M[a][1] * M[b][1] + M[a][2] * M[b][2];
        double t[3];
        for (int a = 0; a < 3; a++) t[a] = V[a] * p[0] + V[4 + a] * p[1] + V[8 + a] * p[2] + V[12 + a];
        const double fx = cam->width / (2.0 * cam->tanfovx), fy = cam->height / (2.0 * cam->tanfovy);
        const double limx = 1.3 * cam->tanfovx, limy = 1.3 * cam->tanfovy;
        const double txtz = t[0] / t[2], tytz = t[1] / t[2];
        const double tx = fmin(limx, fmax(-limx, txtz)) * t[2], ty = fmin(limy, fmax(-limy, tytz)) * t[2];
        const double tzi = 1.0 / t[2], tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const double J00 = fx * tzi, J02 = -fx * tx * tz2, J11 = fy * tzi, J12 = -fy * ty * tz2;
        double T[2][3], TS[2][3];
        for (int j = 0; j < 3; j++) {
            T[0][j] = J00 * V[4 * j + 0] + J02 * V[4 * j + 2];
            T[1][j] = J11 * V[4 * j + 1] + J12 * V[4 * j + 2];
        }
        for (int a = 0; a < 2; a++) for (int j = 0; j < 3; j++) TS[a][j] = T[a][0] * S[0][j] + T[a][1] * S[1][j] + T[a][2] * S[2][j];
        const double ca0 = TS[0][0] * T[0][0] + TS[0][1] * T[0][1] + TS[0][2] * T[0][2];
        const double cb = TS[0][0] * T[1][0] + TS[0][1] * T[1][1] + TS[0][2] * T[1][2];
        const double cc0 = TS[1][0] * T[1][0] + TS[1][1] * T[1][1] + TS[1][2] * T[1][2];
        const double ca = ca0 + 0.3, cc = cc0 + 0.3;
        double dmean[3] = {0, 0, 0};

        /* 1. conic -> cov2D: dSigma' = -Q G Q, Q = conic matrix, G = sym(dA, dB/2, dC) */
        const double det = ca * cc - cb * cb;
        const double kappa = 1.0 / (det * det + 1e-7);
        const double dA = dL_dconic[3 * i], dBh = 0.5 * dL_dconic[3 * i + 1], dC = dL_dconic[3 * i + 2];
        double da = kappa * (-cc * cc * dA + 2.0 * cb * cc * dBh + (det - ca * cc) * dC);
        double dc = kappa * (-ca * ca * dC + 2.0 * ca * cb * dBh + (det - ca * cc) * dA);
        double db = kappa * 2.0 * (cb * cc * dA - (det + 2.0 * cb * cb) * dBh + ca * cb * dC);
        double rho = 1.0; /* anti-aliasing factor: o' = o * rho, rho = sqrt(max(0, det0/det)) */
        if (cam->flags & ORC_FLAG_ANTIALIAS) {
            const double det0 = ca0 * cc0 - cb * cb, r = det0 / det;
            rho = r > 0.0 ? sqrt(r) : 0.0;
            if (r > 0.0) {
                /* dL/do' arrives as dL_dopacity_act; dL/drho = o * dL/do'; drho/dr = 1/(2 rho) */
                const double gr = o * (double)dL_dopacity_act[i] * 0.5 / rho;
                da += gr * (cc0 * det - det0 * cc) / (det * det);
                dc += gr * (ca0 * det - det0 * ca) / (det * det);
                db += gr * (-2.0 * cb * (det - det0)) / (det * det);
            }
        }
        /* 2. cov2D -> Sigma (symmetric gradient) and -> T -> J -> t */
        double dS[3][3];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++)
            dS[a][b] = T[0][a] * T[0][b] * da + 0.5 * (T[0][a] * T[1][b] + T[1][a] * T[0][b]) * db + T[1][a] * T[1][b] * dc;
        double dT[2][3];
        for (int j = 0; j < 3; j++) {
            dT[0][j] = 2.0 * da * TS[0][j] + db * TS[1][j];
            dT[1][j] = 2.0 * dc * TS[1][j] + db * TS[0][j];
        }
        double dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int j = 0; j < 3; j++) {
            dJ00 += V[4 * j + 0] * dT[0][j]; dJ02 += V[4 * j + 2] * dT[0][j];
            dJ11 += V[4 * j + 1] * dT[1][j]; dJ12 += V[4 * j + 2] * dT[1][j];
        }
        const double mx_ = (txtz < -limx || txtz > limx) ? 0.0 : 1.0;
        const double my_ = (tytz < -limy || tytz > limy) ? 0.0 : 1.0;
        double dt[3];
        dt[0] = mx_ * (-fx * tz2) * dJ02;
        dt[1] = my_ * (-fy * tz2) * dJ12;
        dt[2] = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0 * fx * tx) * tz3 * dJ02 + (2.0 * fy * ty) * tz3 * dJ12;
        for (int c = 0; c < 3; c++) dmean[c] += V[4 * c + 0] * dt[0] + V[4 * c + 1] * dt[1] + V[4 * c + 2] * dt[2];

        /* 3. projection: mean2D (ndc-scaled gradient) -> mean3D */
        {
            double h[4];
            for (int a = 0; a < 4; a++) h[a] = Pm[a] * p[0] + Pm[4 + a] * p[1] + Pm[8 + a] * p[2] + Pm[12 + a];
            const double m_w = 1.0 / (h[3] + 1e-7);
            const double mul1 = h[0] * m_w * m_w, mul2 = h[1] * m_w * m_w;
            const double g0 = dL_dmean2D[2 * i], g1 = dL_dmean2D[2 * i + 1];
            for (int c = 0; c < 3; c++)
                dmean[c] += (Pm[4 * c + 0] * m_w - Pm[4 * c + 3] * mul1) * g0 + (Pm[4 * c + 1] * m_w - Pm[4 * c + 3] * mul2) * g1;
        }

        /* 4. SH: colour -> coefficients and -> direction -> mean */
        {
            const double dir_o[3] = {p[0] - cam->campos[0], p[1] - cam->campos[1], p[2] - cam->campos[2]};
            const double len = sqrt(dir_o[0] * dir_o[0] + dir_o[1] * dir_o[1] + dir_o[2] * dir_o[2]);
            const double d[3] = {dir_o[0] / len, dir_o[1] / len, dir_o[2] / len};
            double dcol[3];
            for (int ch = 0; ch < 3; ch++) dcol[ch] = clamped[3 * i + ch] ? 0.0 : (double)dL_dcolor[3 * i + ch];
            const double X = d[0], Y = d[1], Z = d[2];
            double bas[16], gb[16][3]; memset(gb, 0, sizeof gb); memset(bas, 0, sizeof bas);
            bas[0] = SH_C0;
            if (deg >= 1) {
                bas[1] = -(double)SH_C1 * Y; bas[2] = (double)SH_C1 * Z; bas[3] = -(double)SH_C1 * X;
                gb[1][1] = -SH_C1; gb[2][2] = SH_C1; gb[3][0] = -SH_C1;
            }
            if (deg >= 2) {
                const double xx = X * X, yy = Y * Y, zz = Z * Z;
                bas[4] = SH_C2[0] * X * Y; bas[5] = SH_C2[1] * Y * Z; bas[6] = SH_C2[2] * (2 * zz - xx - yy);
                bas[7] = SH_C2[3] * X * Z; bas[8] = SH_C2[4] * (xx - yy);
                gb[4][0] = SH_C2[0] * Y; gb[4][1] = SH_C2[0] * X;
                gb[5][1] = SH_C2[1] * Z; gb[5][2] = SH_C2[1] * Y;
                gb[6][0] = SH_C2[2] * -2.0 * X; gb[6][1] = SH_C2[2] * -2.0 * Y; gb[6][2] = SH_C2[2] * 4.0 * Z;
                gb[7][0] = SH_C2[3] * Z; gb[7][2] = SH_C2[3] * X;
                gb[8][0] = SH_C2[4] * 2.0 * X; gb[8][1] = SH_C2[4] * -2.0 * Y;
                if (deg >= 3) {
                    bas[9] = SH_C3[0] * Y * (3 * xx - yy); bas[10] = SH_C3[1] * X * Y * Z;
                    bas[11] = SH_C3[2] * Y * (4 * zz - xx - yy); bas[12] = SH_C3[3] * Z * (2 * zz - 3 * xx - 3 * yy);
                    bas[13] = SH_C3[4] * X * (4 * zz - xx - yy); bas[14] = SH_C3[5] * Z * (xx - yy);
                    bas[15] = SH_C3[6] * X * (xx - 3 * yy);
                    gb[9][0] = SH_C3[0] * 6.0 * X * Y; gb[9][1] = SH_C3[0] * (3.0 * xx - 3.0 * yy);
                    gb[10][0] = SH_C3[1] * Y * Z; gb[10][1] = SH_C3[1] * X * Z; gb[10][2] = SH_C3[1] * X * Y;
                    gb[11][0] = SH_C3[2] * -2.0 * X * Y; gb[11][1] = SH_C3[2] * (4.0 * zz - xx - 3.0 * yy);
                    gb[11][2] = SH_C3[2] * 8.0 * Y * Z;
                    gb[12][0] = SH_C3[3] * -6.0 * X * Z; gb[12][1] = SH_C3[3] * -6.0 * Y * Z;
                    gb[12][2] = SH_C3[3] * (6.0 * zz - 3.0 * xx - 3.0 * yy);
                    gb[13][0] = SH_C3[4] * (4.0 * zz - 3.0 * xx - yy); gb[13][1] = SH_C3[4] * -2.0 * X * Y;
                    gb[13][2] = SH_C3[4] * 8.0 * X * Z;
                    gb[14][0] = SH_C3[5] * 2.0 * X * Z; gb[14][1] = SH_C3[5] * -2.0 * Y * Z; gb[14][2] = SH_C3[5] * (xx - yy);
                    gb[15][0] = SH_C3[6] * (3.0 * xx - 3.0 * yy); gb[15][1] = SH_C3[6] * -6.0 * X * Y;
                }
            }
            for (int ch = 0; ch < 3; ch++) dL_dsh0[3 * i + ch] = (float)(bas[0] * dcol[ch]);
            for (int k = 1; k < K; k++)
                for (int ch = 0; ch < 3; ch++) dL_dshN[3 * ((size_t)i * KR + (k - 1)) + ch] = (float)(bas[k] * dcol[ch]);
            double ddir[3] = {0, 0, 0};
            for (int k = 1; k < K; k++) {
                double sk = 0.0;
                for (int ch = 0; ch < 3; ch++) sk += (double)shN[3 * ((size_t)i * KR + (k - 1)) + ch] * dcol[ch];
                ddir[0] += gb[k][0] * sk; ddir[1] += gb[k][1] * sk; ddir[2] += gb[k][2] * sk;
            }
            const double dd = d[0] * ddir[0] + d[1] * ddir[1] + d[2] * ddir[2];
            for (int c = 0; c < 3; c++) dmean[c] += (ddir[c] - d[c] * dd) / len;
        }
        for (int c = 0; c < 3; c++) dL_dmeans3D[3 * i + c] = (float)dmean[c];

        /* 5. Sigma -> scale, rotation.  Sigma = Mr Mr^T, Mr = R diag(s) */
        {
            double dM[3][3];
            for (int a = 0; a < 3; a++) for (int k = 0; k < 3; k++)
                dM[a][k] = 2.0 * (dS[a][0] * M[0][k] + dS[a][1] * M[1][k] + dS[a][2] * M[2][k]);
            double ds[3], dR[3][3];
            for (int k = 0; k < 3; k++) {
                ds[k] = R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k];
                for (int a = 0; a < 3; a++) dR[a][k] = dM[a][k] * s[k];
            }
            double dq[4];
            dq[0] = 2.0 * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
            dq[1] = 2.0 * (y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) -
                    4.0 * x * (dR[1][1] + dR[2][2]);
            dq[2] = 2.0 * (x * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + z * (dR[1][2] + dR[2][1])) -
                    4.0 * y * (dR[0][0] + dR[2][2]);
            dq[3] = 2.0 * (r * (dR[1][0] - dR[0][1]) + x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1])) -
                    4.0 * z * (dR[0][0] + dR[1][1]);
            /* 6. activations (chain rule to the stored raw parameters) */
            if (activated) {
                for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = (float)(ds[k] * cam->scale_modifier);
                for (int k = 0; k < 4; k++) dL_dquats[4 * i + k] = (float)dq[k];
                dL_dopacities[i] = (float)(rho * (double)dL_dopacity_act[i]);
            } else {
                for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = (float)(ds[k] * s[k]);
                const double qd = q[0] * dq[0] + q[1] * dq[1] + q[2] * dq[2] + q[3] * dq[3];
                for (int k = 0; k < 4; k++) dL_dquats[4 * i + k] = (float)((dq[k] - q[k] * qd) / qlen);
                dL_dopacities[i] = (float)(rho * (double)dL_dopacity_act[i] * o * (1.0 - o));
            }
        }
    }
}

/* =====================================================================================================================
 * 2DGS ("surfel") variant — GaussianTrainConfig::modelType = 1 (application/diverseshot-cli/source/main.cpp:28,
 * gs_train.cpp:68, docs/userGuide.md:38).  TEST INFRASTRUCTURE like the rest of this file.
 *
 * PARITY UNPINNED: DIVSHOT's 2DGS rasterizer is in the same closed plugin as the 3DGS one.  This
 * restates the published algorithm the option is named after — "2D Gaussian Splatting for Geometrically Accurate Radiance
 * Fields" (Huang et al., SIGGRAPH 2024) and its public rasterizer's semantics:
 *   S.1 a Gaussian is a flat disk: tangents t_u = R[:,0], t_v = R[:,1] with scales (s_u, s_v) (the third scale is
 *       ignored), normal R[:,2].  The homography of the splat's local (u, v, 1) to homogeneous PIXEL coordinates is
 *           M = Npix * Proj * [ s_u t_u | s_v t_v | p ; 0 0 1 ],   rows Tu, Tv, Tw  (x_h, y_h, w_h),
 *       Npix = [[W/2, 0, 0, (W-1)/2], [0, H/2, 0, (H-1)/2], [0, 0, 0, 1]] (the same pixel-centre map as ndc2Pix).
 *       Screen bounds from M (3-sigma): tp = (9, 9, -1), dist = sum tp_j Tw_j^2 (0 -> invisible), f = tp / dist,
 *       centre c = (sum f Tu Tw, sum f Tv Tw), half extent e = sqrt(max(1e-4, c^2 - (sum f Tu^2, sum f Tv^2))),
 *       radius = ceil(max(e_x, e_y, 3 * 0.707106)); tile rect, depth key (view z), cull (z <= 0.2) and colour as 3DGS.
 *   S.2 per pixel (x, y) and splat: k = x Tw - Tu, l = y Tw - Tv, pv = k x l (pv.z == 0 -> skip), (u, v) = pv.xy / pv.z,
 *       rho3d = u^2 + v^2; object-space low-pass: rho2d = 2 |c - (x, y)|^2; rho = min(rho3d, rho2d);
 *       depth = rho3d <= rho2d ? u Tw.x + v Tw.y + Tw.z : Tw.z (< 0.2 -> skip); alpha = min(0.99, o exp(-rho / 2));
 *       then exactly the 3DGS compositing rules (alpha < 1/255 skip, T < 1e-4 stop, n_contrib, final_T).
 *   S.3 / S.4 analytic backward to M (9), c (2), opacity, colour and on to the stored parameters; decisions carry no
 *       gradient, the 0.99 clamp is straight-through.
 * Pinned by tests/autograd_ref_2dgs.py (float64 autograd of S.1-S.2) and closed-form cases (tests/test_oracle_2dgs.py).
 * ===================================================================================================================== */
#define SRF_FILTER_INV_SQ 2.0f
#define SRF_FILTER_3SIGMA (3.0f * 0.707106f)

/* homography rows from the activated surfel: T[0..2] = Tu, T[3..5] = Tv, T[6..8] = Tw (index j: u, v, 1) */
static inline void srf_transmat(const orc_camera* cam, const float* p, const float s[3], float R[3][3], float T[9]) {
    const float hw = 0.5f * (float)cam->width, hh = 0.5f * (float)cam->height;
    const float ow = 0.5f * (float)(cam->width - 1), oh = 0.5f * (float)(cam->height - 1);
    for (int j = 0; j < 3; j++) {
        float v[3], w1;
        if (j < 2) { for (int a = 0; a < 3; a++) v[a] = R[a][j] * s[j]; w1 = 0.0f; }
        else { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; w1 = 1.0f; }
        const float* m = cam->proj;
        const float cx = fmaf(m[0], v[0], fmaf(m[4], v[1], fmaf(m[8], v[2], m[12] * w1)));
        const float cy = fmaf(m[1], v[0], fmaf(m[5], v[1], fmaf(m[9], v[2], m[13] * w1)));
        const float cw = fmaf(m[3], v[0], fmaf(m[7], v[1], fmaf(m[11], v[2], m[15] * w1)));
        T[j] = fmaf(hw, cx, ow * cw);
        T[3 + j] = fmaf(hh, cy, oh * cw);
        T[6 + j] = cw;
    }
}

void orc2_preprocess_fwd(const orc_camera* cam, int32_t N, const float* means3D, const float* scales,
                         const float* quats, const float* opacities, const float* sh0, const float* shN,
                         float* depth, int32_t* radii, float* mean2D, float* transmat /*[N,9]*/, float* opacity_act,
                         float* rgb, uint8_t* clamped, uint32_t* tiles_touched, int32_t* rect) {
    const int gx = (cam->width + TILE - 1) / TILE, gy = (cam->height + TILE - 1) / TILE;
    const int deg = cam->sh_degree, K = (deg + 1) * (deg + 1), KR = cam->sh_rest_alloc;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        radii[i] = 0; tiles_touched[i] = 0; depth[i] = 0.0f; opacity_act[i] = 0.0f;
        mean2D[2 * i] = mean2D[2 * i + 1] = 0.0f;
        for (int k = 0; k < 9; k++) transmat[9 * i + k] = 0.0f;
        for (int k = 0; k < 4; k++) rect[4 * i + k] = 0;
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.0f; clamped[3 * i + k] = 0; }
        const float* p = means3D + 3 * (size_t)i;
        float t[3];
        xform43(cam->view, p, t);
        if (t[2] <= 0.2f) continue;
        act_t a; activate(cam, scales, quats, opacities, i, &a);
        float R[3][3], T[9];
        quat_to_R(a.q, R);
        srf_transmat(cam, p, a.s, R, T);
        const float tp[3] = {9.0f, 9.0f, -1.0f};
        const float dist = fmaf(tp[0] * T[6], T[6], fmaf(tp[1] * T[7], T[7], tp[2] * T[8] * T[8]));
        if (dist == 0.0f) continue;
        const float f[3] = {tp[0] / dist, tp[1] / dist, tp[2] / dist};
        const float cx = fmaf(f[0] * T[0], T[6], fmaf(f[1] * T[1], T[7], f[2] * T[2] * T[8]));
        const float cy = fmaf(f[0] * T[3], T[6], fmaf(f[1] * T[4], T[7], f[2] * T[5] * T[8]));
        const float qx = fmaf(f[0] * T[0], T[0], fmaf(f[1] * T[1], T[1], f[2] * T[2] * T[2]));
        const float qy = fmaf(f[0] * T[3], T[3], fmaf(f[1] * T[4], T[4], f[2] * T[5] * T[5]));
        const float ex = sqrtf(fmaxf(1e-4f, fmaf(cx, cx, -qx))), ey = sqrtf(fmaxf(1e-4f, fmaf(cy, cy, -qy)));
        const float rad_f = ceilf(fmaxf(fmaxf(ex, ey), SRF_FILTER_3SIGMA));
        if (!(rad_f < 1.0e9f)) continue; /* degenerate homography (NaN / inf extent) */
        const int32_t rad = f2i_rz(rad_f);
        const float radf = (float)rad;
        const int32_t minx = imin(gx, imax(0, f2i_rz((cx - radf) * 0.0625f)));
        const int32_t miny = imin(gy, imax(0, f2i_rz((cy - radf) * 0.0625f)));
        const int32_t maxx = imin(gx, imax(0, f2i_rz((cx + radf + 15.0f) * 0.0625f)));
        const int32_t maxy = imin(gy, imax(0, f2i_rz((cy + radf + 15.0f) * 0.0625f)));
        const int64_t area = (int64_t)(maxx - minx) * (int64_t)(maxy - miny);
        if (area <= 0) continue;
        float dir[3] = {p[0] - cam->campos[0], p[1] - cam->campos[1], p[2] - cam->campos[2]};
        const float len = sqrtf(fmaf(dir[0], dir[0], fmaf(dir[1], dir[1], dir[2] * dir[2])));
        const float linv = 1.0f / len;
        dir[0] *= linv; dir[1] *= linv; dir[2] *= linv;
        float bas[16]; sh_basis(deg, dir, bas);
        for (int ch = 0; ch < 3; ch++) {
            float acc = bas[0] * sh0[3 * (size_t)i + ch];
            for (int k = 1; k < K; k++) acc = fmaf(bas[k], shN[3 * ((size_t)i * KR + (k - 1)) + ch], acc);
            acc += 0.5f;
            clamped[3 * i + ch] = acc < 0.0f;
            rgb[3 * i + ch] = fmaxf(acc, 0.0f);
        }
        depth[i] = t[2];
        radii[i] = rad;
        mean2D[2 * i] = cx; mean2D[2 * i + 1] = cy;
        for (int k = 0; k < 9; k++) transmat[9 * i + k] = T[k];
        opacity_act[i] = a.o;
        tiles_touched[i] = (uint32_t)area;
        rect[4 * i] = minx; rect[4 * i + 1] = miny; rect[4 * i + 2] = maxx; rect[4 * i + 3] = maxy;
    }
}

/* one (pixel, surfel) evaluation of S.2; returns 0 if the pair is skipped before the alpha test */
typedef struct { float u, v, pz, rho3d, rho2d, dx, dy, G, alpha; float k[3], l[3]; int use3d; } srf_pair_t;
static inline int srf_pair(const float* T, float cx, float cy, float o, float pxf, float pyf, srf_pair_t* r) {
    for (int j = 0; j < 3; j++) { r->k[j] = fmaf(pxf, T[6 + j], -T[j]); r->l[j] = fmaf(pyf, T[6 + j], -T[3 + j]); }
    const float p0 = fmaf(r->k[1], r->l[2], -(r->k[2] * r->l[1]));
    const float p1 = fmaf(r->k[2], r->l[0], -(r->k[0] * r->l[2]));
    const float p2 = fmaf(r->k[0], r->l[1], -(r->k[1] * r->l[0]));
    if (p2 == 0.0f) return 0;
    r->pz = p2; r->u = p0 / p2; r->v = p1 / p2;
    r->rho3d = fmaf(r->u, r->u, r->v * r->v);
    r->dx = cx - pxf; r->dy = cy - pyf;
    r->rho2d = SRF_FILTER_INV_SQ * fmaf(r->dx, r->dx, r->dy * r->dy);
    r->use3d = r->rho3d <= r->rho2d;
    const float rho = r->use3d ? r->rho3d : r->rho2d;
    const float dep = r->use3d ? fmaf(r->u, T[6], fmaf(r->v, T[7], T[8])) : T[8];
    if (dep < 0.2f) return 0;
    r->G = expf(-0.5f * rho);
    r->alpha = fminf(0.99f, o * r->G);
    return 1;
}

void orc2_render_fwd(const orc_camera* cam, const uint32_t* ranges, const uint32_t* point_list, const float* mean2D,
                     const float* transmat, const float* opacity_act, const float* rgb, float* out_color,
                     float* final_T, uint32_t* n_contrib, uint8_t* fragile, int32_t threads) {
    const int W = cam->width, H = cam->height;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t P = (size_t)W * H;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int tx0 = (tile % gx) * TILE, ty0 = (tile / gx) * TILE;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx0 + lx, py = ty0 + ly;
                if (px >= W || py >= H) continue;
                float T = 1.0f, C[3] = {0, 0, 0};
                uint32_t contributor = 0, last = 0;
                uint8_t frag = 0;
                for (uint32_t j = r0; j < r1; j++) {
                    contributor++;
                    const uint32_t g = point_list[j];
                    srf_pair_t q;
                    if (!srf_pair(transmat + 9 * (size_t)g, mean2D[2 * g], mean2D[2 * g + 1], opacity_act[g], (float)px, (float)py, &q))
                        continue;
                    const float araw = opacity_act[g] * q.G;
                    if (fabsf(araw - (1.0f / 255.0f)) <= (1.0f / 255.0f) * 1e-4f) frag = 1;
                    if (fabsf(q.rho3d - q.rho2d) <= 1e-5f * fmaxf(q.rho3d, q.rho2d)) frag = 1;
                    if (q.alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1.0f - q.alpha);
                    if (fabsf(test_T - 1e-4f) <= 1e-4f * 2e-3f) frag = 1;
                    if (test_T < 1e-4f) break;
                    const float w = q.alpha * T;
                    C[0] += rgb[3 * g] * w; C[1] += rgb[3 * g + 1] * w; C[2] += rgb[3 * g + 2] * w;
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                final_T[pix] = T;
                n_contrib[pix] = last;
                if (fragile) fragile[pix] = frag;
                for (int ch = 0; ch < 3; ch++) out_color[ch * P + pix] = C[ch] + T * cam->bg[ch];
            }
    }
}

/* S.3: reverse walk; per-surfel sums in double (see atomic_addd).  Outputs zeroed then accumulated:
 * dL_dT [N,9], dL_dmean2D [N,2] (pixel units, NOT ndc-scaled), dL_dopacity [N] (w.r.t. the activated opacity), dL_dcolor [N,3] */
void orc2_render_bwd(const orc_camera* cam, int32_t N, const uint32_t* ranges, const uint32_t* point_list,
                     const float* mean2D, const float* transmat, const float* opacity_act, const float* rgb,
                     const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float* dL_dT,
                     float* dL_dmean2D, float* dL_dopacity, float* dL_dcolor, int32_t threads) {
    const int W = cam->width, H = cam->height;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t P = (size_t)W * H;
    double* acc = (double*)calloc((size_t)(N > 0 ? N : 1) * 15, sizeof(double));
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int tx0 = (tile % gx) * TILE, ty0 = (tile / gx) * TILE;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx0 + lx, py = ty0 + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = final_T[pix];
                float T = T_final;
                const uint32_t last = n_contrib[pix];
                const float dp[3] = {dL_dpix[pix], dL_dpix[P + pix], dL_dpix[2 * P + pix]};
                float accum[3] = {0, 0, 0}, last_alpha = 0.0f, last_color[3] = {0, 0, 0};
                for (uint32_t j = r1; j-- > r0;) {
                    const uint32_t contributor = j - r0; /* 0-based index of the entry */
                    if (contributor >= last) continue;
                    const uint32_t g = point_list[j];
                    const float* Tm = transmat + 9 * (size_t)g;
                    srf_pair_t q;
                    if (!srf_pair(Tm, mean2D[2 * g], mean2D[2 * g + 1], opacity_act[g], pxf, pyf, &q)) continue;
                    if (q.alpha < 1.0f / 255.0f) continue;
                    T = T / (1.0f - q.alpha);
                    float dL_dalpha = 0.0f;
                    double* A = acc + 15 * (size_t)g;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = rgb[3 * g + ch];
                        accum[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum[ch]) * dp[ch];
                        atomic_addd(A + 12 + ch, (double)(q.alpha * T * dp[ch]));
                    }
                    dL_dalpha *= T;
                    last_alpha = q.alpha;
                    float bg_dot = 0.0f;
                    for (int ch = 0; ch < 3; ch++) bg_dot += cam->bg[ch] * dp[ch];
                    dL_dalpha += (-T_final / (1.0f - q.alpha)) * bg_dot;
                    const float dL_dG = opacity_act[g] * dL_dalpha;
                    atomic_addd(A + 11, (double)(q.G * dL_dalpha));
                    if (q.use3d) {
                        /* G = exp(-(u^2 + v^2) / 2): dL/du = -G u dL/dG; (u, v) = pv.xy / pv.z, pv = k x l */
                        const float dLu = dL_dG * -q.G * q.u, dLv = dL_dG * -q.G * q.v;
                        const float dsx = dLu / q.pz, dsy = dLv / q.pz;
                        const float dpv[3] = {dsx, dsy, -(dsx * q.u + dsy * q.v)};
                        /* pv = k x l: dL/dk = l x dpv, dL/dl = dpv x k */
                        const float dk[3] = {q.l[1] * dpv[2] - q.l[2] * dpv[1], q.l[2] * dpv[0] - q.l[0] * dpv[2],
                                             q.l[0] * dpv[1] - q.l[1] * dpv[0]};
                        const float dl[3] = {dpv[1] * q.k[2] - dpv[2] * q.k[1], dpv[2] * q.k[0] - dpv[0] * q.k[2],
                                             dpv[0] * q.k[1] - dpv[1] * q.k[0]};
                        for (int c = 0; c < 3; c++) {
                            atomic_addd(A + c, (double)(-dk[c]));                          /* k = x Tw - Tu */
                            atomic_addd(A + 3 + c, (double)(-dl[c]));                      /* l = y Tw - Tv */
                            atomic_addd(A + 6 + c, (double)(pxf * dk[c] + pyf * dl[c]));
                        }
                    } else {
                        /* G = exp(-|c - pix|^2): low-pass branch, gradient to the projected centre */
                        atomic_addd(A + 9, (double)(dL_dG * -q.G * SRF_FILTER_INV_SQ * q.dx));
                        atomic_addd(A + 10, (double)(dL_dG * -q.G * SRF_FILTER_INV_SQ * q.dy));
                    }
                }
            }
    }
    for (size_t i = 0; i < (size_t)N; i++) {
        for (int k = 0; k < 9; k++) dL_dT[9 * i + k] = (float)acc[15 * i + k];
        dL_dmean2D[2 * i] = (float)acc[15 * i + 9]; dL_dmean2D[2 * i + 1] = (float)acc[15 * i + 10];
        dL_dopacity[i] = (float)acc[15 * i + 11];
        for (int k = 0; k < 3; k++) dL_dcolor[3 * i + k] = (float)acc[15 * i + 12 + k];
    }
    free(acc);
}

/* S.4: per-surfel backward to the stored parameters (double arithmetic like orc_preprocess_bwd).  Inputs: dL/dT [N,9],
 * dL/dmean2D [N,2] (pixel units), dL/dopacity (activated), dL/dcolor.  The third scale receives no gradient. */
void orc2_preprocess_bwd(const orc_camera* cam, int32_t N, const float* means3D, const float* scales,
                         const float* quats, const float* opacities, const float* sh0, const float* shN,
                         const int32_t* radii, const uint8_t* clamped, const float* dL_dT, const float* dL_dmean2D,
                         const float* dL_dopacity_act, const float* dL_dcolor, float* dL_dmeans3D,
                         float* dL_dscales, float* dL_dquats, float* dL_dopacities, float* dL_dsh0, float* dL_dshN) {
    const int deg = cam->sh_degree, K = (deg + 1) * (deg + 1), KR = cam->sh_rest_alloc;
    const double hw = 0.5 * cam->width, hh = 0.5 * cam->height, ow = 0.5 * (cam->width - 1), oh = 0.5 * (cam->height - 1);
    (void)sh0;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < 3; k++) { dL_dmeans3D[3 * i + k] = 0; dL_dscales[3 * i + k] = 0; dL_dsh0[3 * i + k] = 0; }
        for (int k = 0; k < 4; k++) dL_dquats[4 * i + k] = 0;
        dL_dopacities[i] = 0;
        for (int k = 0; k < 3 * KR; k++) dL_dshN[3 * (size_t)KR * i + k] = 0;
        if (radii[i] <= 0) continue;
        const float* pf = means3D + 3 * (size_t)i;
        const double p[3] = {pf[0], pf[1], pf[2]};
        act_t a; activate(cam, scales, quats, opacities, i, &a);
        float Rf[3][3], Tf[9];
        quat_to_R(a.q, Rf);
        srf_transmat(cam, pf, a.s, Rf, Tf);
        double T[9], dT[9];
        for (int k = 0; k < 9; k++) { T[k] = Tf[k]; dT[k] = dL_dT[9 * (size_t)i + k]; }
        /* 1. projected centre -> T */
        {
            const double tp[3] = {9.0, 9.0, -1.0};
            const double dist = tp[0] * T[6] * T[6] + tp[1] * T[7] * T[7] + tp[2] * T[8] * T[8];
            double cx = 0, cy = 0, f[3];
            for (int j = 0; j < 3; j++) { f[j] = tp[j] / dist; cx += f[j] * T[j] * T[6 + j]; cy += f[j] * T[3 + j] * T[6 + j]; }
            const double gx_ = dL_dmean2D[2 * i], gy_ = dL_dmean2D[2 * i + 1];
            for (int j = 0; j < 3; j++) {
                dT[j] += gx_ * f[j] * T[6 + j];
                dT[3 + j] += gy_ * f[j] * T[6 + j];
                dT[6 + j] += gx_ * (f[j] * T[j] - 2.0 * cx * f[j] * T[6 + j]) + gy_ * (f[j] * T[3 + j] - 2.0 * cy * f[j] * T[6 + j]);
            }
        }
        /* 2./3. T -> clip columns -> (L0, L1, p) through Proj^T */
        double dvec[3][3];
        for (int j = 0; j < 3; j++) {
            const double dcx = hw * dT[j], dcy = hh * dT[3 + j], dcw = ow * dT[j] + oh * dT[3 + j] + dT[6 + j];
            for (int r = 0; r < 3; r++)
                dvec[j][r] = (double)cam->proj[4 * r + 0] * dcx + (double)cam->proj[4 * r + 1] * dcy + (double)cam->proj[4 * r + 3] * dcw;
        }
        double dmean[3] = {dvec[2][0], dvec[2][1], dvec[2][2]};
        /* 4. SH: colour -> coefficients and -> direction -> mean (as orc_preprocess_bwd) */
        {
            const double dir_o[3] = {p[0] - cam->campos[0], p[1] - cam->campos[1], p[2] - cam->campos[2]};
            const double len = sqrt(dir_o[0] * dir_o[0] + dir_o[1] * dir_o[1] + dir_o[2] * dir_o[2]);
            const float df[3] = {(float)(dir_o[0] / len), (float)(dir_o[1] / len), (float)(dir_o[2] / len)};
            const double d[3] = {dir_o[0] / len, dir_o[1] / len, dir_o[2] / len};
            double dcol[3];
            for (int ch = 0; ch < 3; ch++) dcol[ch] = clamped[3 * i + ch] ? 0.0 : (double)dL_dcolor[3 * i + ch];
            float basf[16]; memset(basf, 0, sizeof basf);
            sh_basis(deg, df, basf);
            for (int ch = 0; ch < 3; ch++) dL_dsh0[3 * i + ch] = (float)((double)basf[0] * dcol[ch]);
            for (int k = 1; k < K; k++)
                for (int ch = 0; ch < 3; ch++) dL_dshN[3 * ((size_t)i * KR + (k - 1)) + ch] = (float)((double)basf[k] * dcol[ch]);
            /* d basis / d direction by central differences of the float64 basis would be circular: analytic, as in A8 */
            const double X = d[0], Y = d[1], Z = d[2];
            double gb[16][3]; memset(gb, 0, sizeof gb);
            if (deg >= 1) { gb[1][1] = -SH_C1; gb[2][2] = SH_C1; gb[3][0] = -SH_C1; }
            if (deg >= 2) {
                const double xx = X * X, yy = Y * Y, zz = Z * Z;
                gb[4][0] = SH_C2[0] * Y; gb[4][1] = SH_C2[0] * X;
                gb[5][1] = SH_C2[1] * Z; gb[5][2] = SH_C2[1] * Y;
                gb[6][0] = SH_C2[2] * -2.0 * X; gb[6][1] = SH_C2[2] * -2.0 * Y; gb[6][2] = SH_C2[2] * 4.0 * Z;
                gb[7][0] = SH_C2[3] * Z; gb[7][2] = SH_C2[3] * X;
                gb[8][0] = SH_C2[4] * 2.0 * X; gb[8][1] = SH_C2[4] * -2.0 * Y;
                if (deg >= 3) {
                    gb[9][0] = SH_C3[0] * 6.0 * X * Y; gb[9][1] = SH_C3[0] * (3.0 * xx - 3.0 * yy);
                    gb[10][0] = SH_C3[1] * Y * Z; gb[10][1] = SH_C3[1] * X * Z; gb[10][2] = SH_C3[1] * X * Y;
                    gb[11][0] = SH_C3[2] * -2.0 * X * Y; gb[11][1] = SH_C3[2] * (4.0 * zz - xx - 3.0 * yy);
                    gb[11][2] = SH_C3[2] * 8.0 * Y * Z;
                    gb[12][0] = SH_C3[3] * -6.0 * X * Z; gb[12][1] = SH_C3[3] * -6.0 * Y * Z;
                    gb[12][2] = SH_C3[3] * (6.0 * zz - 3.0 * xx - 3.0 * yy);
                    gb[13][0] = SH_C3[4] * (4.0 * zz - 3.0 * xx - yy); gb[13][1] = SH_C3[4] * -2.0 * X * Y;
                    gb[13][2] = SH_C3[4] * 8.0 * X * Z;
                    gb[14][0] = SH_C3[5] * 2.0 * X * Z; gb[14][1] = SH_C3[5] * -2.0 * Y * Z; gb[14][2] = SH_C3[5] * (xx - yy);
                    gb[15][0] = SH_C3[6] * (3.0 * xx - 3.0 * yy); gb[15][1] = SH_C3[6] * -6.0 * X * Y;
                }
            }
            double ddir[3] = {0, 0, 0};
            for (int k = 1; k < K; k++) {
                double sk = 0.0;
                for (int ch = 0; ch < 3; ch++) sk += (double)shN[3 * ((size_t)i * KR + (k - 1)) + ch] * dcol[ch];
                ddir[0] += gb[k][0] * sk; ddir[1] += gb[k][1] * sk; ddir[2] += gb[k][2] * sk;
            }
            const double dd = d[0] * ddir[0] + d[1] * ddir[1] + d[2] * ddir[2];
            for (int c = 0; c < 3; c++) dmean[c] += (ddir[c] - d[c] * dd) / len;
        }
        for (int c = 0; c < 3; c++) dL_dmeans3D[3 * i + c] = (float)dmean[c];
        /* 5. (L0, L1) = (s_u R[:,0], s_v R[:,1]) -> scales, rotation; 6. activations */
        {
            const double s[3] = {a.s[0], a.s[1], a.s[2]};
            const double q[4] = {a.q[0], a.q[1], a.q[2], a.q[3]};
            const double r = q[0], x = q[1], y = q[2], z = q[3];
            double ds[3] = {0, 0, 0}, dR[3][3];
            for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) dR[c][k] = 0.0;
            for (int k = 0; k < 2; k++)
                for (int c = 0; c < 3; c++) { ds[k] += (double)Rf[c][k] * dvec[k][c]; dR[c][k] = s[k] * dvec[k][c]; }
            double dq[4];
            dq[0] = 2.0 * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
            dq[1] = 2.0 * (y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) -
                    4.0 * x * (dR[1][1] + dR[2][2]);
            dq[2] = 2.0 * (x * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + z * (dR[1][2] + dR[2][1])) -
                    4.0 * y * (dR[0][0] + dR[2][2]);
            dq[3] = 2.0 * (r * (dR[1][0] - dR[0][1]) + x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1])) -
                    4.0 * z * (dR[0][0] + dR[1][1]);
            const double o = a.o;
            if (cam->flags & ORC_FLAG_INPUT_ACTIVATED) {
                for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = (float)(ds[k] * cam->scale_modifier);
                for (int k = 0; k < 4; k++) dL_dquats[4 * i + k] = (float)dq[k];
                dL_dopacities[i] = dL_dopacity_act[i];
            } else {
                for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = (float)(ds[k] * s[k]);
                const double qd = q[0] * dq[0] + q[1] * dq[1] + q[2] * dq[2] + q[3] * dq[3];
                for (int k = 0; k < 4; k++) dL_dquats[4 * i + k] = (float)((dq[k] - q[k] * qd) / (double)a.qlen);
                dL_dopacities[i] = (float)((double)dL_dopacity_act[i] * o * (1.0 - o));
            }
        }
    }
}
