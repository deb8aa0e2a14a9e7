// TEST INFRASTRUCTURE — the few words of HLSL vocabulary needed to compile, as C++, the functions of the reference's
// shaders that restate the rasterizer's per-Gaussian maths (SURVEY.md §8c "files a CPU restatement must follow"):
//   diverse/assets/shaders/gaussian/gsplat_intersect.hlsl:61-134   computeCov3D, computeCov2D (1.3 clamp, +0.3 blur)
//   diverse/assets/shaders/gaussian/gsplat_sh.hlsl:41-104           SH constants and evalSH
//   diverse/assets/shaders/gaussian/gsplat_vs.hlsl:211-214          ndc2Pix
//   diverse/assets/shaders/gaussian/gsplat_vs.hlsl:297-300          mip-splatting anti-aliasing factor
// The shader text itself is NOT in this repo: oracle/Makefile (`make ref`) cuts those line ranges out of
// /root/reference into oracle/_ref/gen/*.inc at build time and ref_hlsl_shim.cpp includes them.
// Semantics provided: value types with HLSL constructors (scalars of any arithmetic type), row-major
// float3x3/float4x4 indexed m[row][col], mul(A,B) = matrix product, component-wise operators.
#pragma once
#include <cmath>
#include <type_traits>

#define in
typedef unsigned int uint;

struct float2 {
    float x = 0, y = 0;
    float2() = default;
    template <class A, class B> float2(A a, B b) : x(float(a)), y(float(b)) {}
};
struct float3 {
    float x = 0, y = 0, z = 0;
    float3() = default;
    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>> float3(S s) : x(float(s)), y(float(s)), z(float(s)) {}
    template <class A, class B, class C> float3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
    float3& operator+=(const float3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
struct float4 {
    float x = 0, y = 0, z = 0, w = 0;
    float4() = default;
    template <class A, class B, class C, class D> float4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
};
inline float3 operator-(const float3& a) { return float3(-a.x, -a.y, -a.z); }
inline float3 operator+(const float3& a, const float3& b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3& a, const float3& b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
inline float3 operator*(const float3& a, S s) { const float f = float(s); return float3(a.x * f, a.y * f, a.z * f); }
template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
inline float3 operator*(S s, const float3& a) { return a * s; }

struct float3x3 {
    float m[3][3] = {};
    float3x3() = default;
    template <class A, class B, class C, class D, class E, class F, class G, class H, class I>
    float3x3(A a, B b, C c, D d, E e, F f, G g, H h, I i)
        : m{{float(a), float(b), float(c)}, {float(d), float(e), float(f)}, {float(g), float(h), float(i)}} {}
    float* operator[](int r) { return m[r]; }
    const float* operator[](int r) const { return m[r]; }
};
struct float4x4 {
    float m[4][4] = {};
    float* operator[](int r) { return m[r]; }
    const float* operator[](int r) const { return m[r]; }
};
inline float3x3 mul(const float3x3& a, const float3x3& b) {
    float3x3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0;
            for (int k = 0; k < 3; k++) s += a[i][k] * b[k][j];
            r[i][j] = s;
        }
    return r;
}
inline float3x3 transpose(const float3x3& a) {
    float3x3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i][j] = a[j][i];
    return r;
}
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline float min(A a, B b) { return float(a) < float(b) ? float(a) : float(b); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline float max(A a, B b) { return float(a) > float(b) ? float(a) : float(b); }
