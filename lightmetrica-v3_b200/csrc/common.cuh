// common.cuh — shared definitions for the sm_100a kernels of liblmcu.
//
// Everything marked LMCU_HD is plain arithmetic on plain data so that the per-item bodies of the
// kernels can also be compiled by g++ in tests/hostsim (a debugging aid for the GPU-less build
// container; it is never linked into the product and the product has no CPU path).
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define LMCU_HD __host__ __device__ __forceinline__
#define LMCU_D __device__ __forceinline__
#else
#define LMCU_HD inline
#define LMCU_D inline
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct double3 { double x, y, z; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
inline float3 make_float3(float x, float y, float z) { return float3{ x, y, z }; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{ x, y }; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }
inline double3 make_double3(double x, double y, double z) { return double3{ x, y, z }; }
#endif

namespace lmcu {

constexpr double kInf = 1e10;     // include/lm/math.h:54
constexpr double kEps = 1e-3;     // include/lm/math.h:55
constexpr float kPiF = 3.14159265358979323846f;

// -- bit casts --------------------------------------------------------------------------------
LMCU_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LMCU_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// -- FP64 arithmetic that must never be contracted into FMAs -----------------------------------
// The reference is compiled for baseline x86-64 (no FMA), so a*b+c is two roundings.  On the
// device the _rn intrinsics are never fused by ptxas; on the host hostsim is built with
// -ffp-contract=off.
LMCU_HD double mulrn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
LMCU_HD double addrn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
LMCU_HD double subrn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
LMCU_HD double divrn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
LMCU_HD double sqrtrn(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}

// x * copysign(1.0, d) (accel_sahbvh.cpp:66-68: u *= sign(det)): multiplying by +-1.0 is exact, so flipping the sign bit
// with one integer op gives the same bits for every finite, zero or infinite x.
LMCU_HD double mul_sign_of(double x, double d) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x) ^ (__double2hiint(d) & (int)0x80000000), __double2loint(x));
#else
    return x * copysign(1.0, d);
#endif
}

// glm::dot(vec3) = (x*x' + y*y') + z*z'; glm::cross as in func_geometric.inl
struct D3 { double x, y, z; };
LMCU_HD D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
LMCU_HD double ddot(D3 a, D3 b) { return addrn(addrn(mulrn(a.x, b.x), mulrn(a.y, b.y)), mulrn(a.z, b.z)); }
LMCU_HD D3 dcross(D3 x, D3 y) {
    return d3(subrn(mulrn(x.y, y.z), mulrn(y.y, x.z)),
              subrn(mulrn(x.z, y.x), mulrn(y.z, x.x)),
              subrn(mulrn(x.x, y.y), mulrn(y.x, x.y)));
}
LMCU_HD D3 dsub3(D3 a, D3 b) { return d3(subrn(a.x, b.x), subrn(a.y, b.y), subrn(a.z, b.z)); }

// -- float rounding toward -inf / +inf from double ----------------------------------------------
LMCU_HD float f32_down(double v) {
#if defined(__CUDA_ARCH__)
    return __double2float_rd(v);
#else
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return f;
#endif
}
LMCU_HD float f32_up(double v) {
#if defined(__CUDA_ARCH__)
    return __double2float_ru(v);
#else
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
#endif
}

// -- small float3 helpers (FP32 shading) ----------------------------------------------------------
struct F3 { float x, y, z; };
LMCU_HD F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
LMCU_HD F3 f3s(float s) { return f3(s, s, s); }
LMCU_HD F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
LMCU_HD F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
LMCU_HD F3 operator-(F3 a) { return f3(-a.x, -a.y, -a.z); }
LMCU_HD F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
LMCU_HD F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
LMCU_HD F3 operator*(float s, F3 a) { return f3(a.x * s, a.y * s, a.z * s); }
LMCU_HD float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LMCU_HD F3 cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
LMCU_HD F3 normalize(F3 a) { float s = 1.0f / sqrtf(dot(a, a)); return a * s; }
LMCU_HD float max3(F3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
LMCU_HD bool is_zero(F3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

}  // namespace lmcu
