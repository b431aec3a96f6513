// oracle/ref_shim.h — TEST INFRASTRUCTURE ONLY.
//
// A GLSL execution environment in C++ so that the reference's OWN shader files (read from /root/reference/data/shaders at
// build time by oracle/make_ref.py, never copied into this repo) compile with g++ and run on the CPU: oracle/_ref/libvhr_ref.so.
// The vector/matrix types, swizzles and built-in functions are the reference's own vendored glm
// (/root/reference/dependencies/glm, GLM_FORCE_SWIZZLE), the struct definitions are the reference's own dual-language
// header (src/rendering_backend/glsl_common.h). What this file adds is only what a Vulkan driver supplies at run time:
//   * image2D / sampler2D objects over plain arrays: imageLoad / imageStore with the texel formats of the hot path (fp16
//     stores round to nearest even, two-channel images read back (r, g, 0, 1)), texture() with the Vulkan spec's LOD-0
//     NEAREST / LINEAR formulas and the four address modes, textureSize / imageSize;
//   * traceRayEXT: the ray query itself is third-party driver arithmetic; it is bound to the double-precision
//     BVH of oracle/oracle_rt.cpp through two function pointers, and the miss / closest-hit SHADERS it invokes are again the
//     reference's files;
//   * a handful of operator overloads glm does not have but GLSL does (ivec2 * vec2, int-literal * vec3, swizzle /= scalar).
// Built with -ffp-contract=off: every GLSL operation is one IEEE fp32 operation, none fused.
#pragma once
#define GLM_FORCE_SWIZZLE
// GLM_FORCE_INTRINSICS only switches on glm's anonymous-struct swizzle members under gcc (setup.hpp:75-81,459); the default
// packed_highp types used here keep glm's scalar code paths.
#define GLM_FORCE_INTRINSICS
#include <cmath>
#include <cstdint>
#include <cstring>

#include "glm/glm.hpp"

namespace glm {
// ---- GLSL implicit conversions glm's templates do not deduce (int -> float, ivec -> vec, swizzle proxies) ---------------------
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> operator-(detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &s) { return -s(); }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline void operator/=(detail::_swizzle<N, T, Q, E0, E1, E2, E3> &s, T v) { s = s() / v; }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline void operator*=(detail::_swizzle<N, T, Q, E0, E1, E2, E3> &s, T v) { s = s() * v; }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> operator/(detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &s, T v) { return s() / v; }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> &operator+=(vec<N, T, Q> &a, detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &s) { return a += s(); }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> &operator-=(vec<N, T, Q> &a, detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &s) { return a -= s(); }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> abs(detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &s) { return abs(s()); }
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline vec<N, T, Q> normalize(detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &s) { return normalize(s()); }

// binary built-ins with a swizzle proxy on one side (cross(n, t.xyz), dot(t.xyz, n))
#define VHR_REF_SWZ_BINARY(fn, ret)                                                                                                          \
    template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>                                                                \
    inline ret fn(vec<N, T, Q> const &a, detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &b) { return fn(a, b()); }                        \
    template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>                                                                \
    inline ret fn(detail::_swizzle<N, T, Q, E0, E1, E2, E3> const &a, vec<N, T, Q> const &b) { return fn(a(), b); }
#define VHR_REF_COMMA ,
VHR_REF_SWZ_BINARY(cross, vec<N VHR_REF_COMMA T VHR_REF_COMMA Q>)
VHR_REF_SWZ_BINARY(dot, T)
#undef VHR_REF_SWZ_BINARY
#undef VHR_REF_COMMA
template <length_t L, qualifier Q> inline vec<L, float, Q> to_float(vec<L, int, Q> const &v) { return vec<L, float, Q>(v); }
// ivec (op) vec, vec (op) ivec
#define VHR_REF_MIXED(op)                                                                                                                   \
    template <length_t L, qualifier Q> inline vec<L, float, Q> operator op(vec<L, int, Q> const &a, vec<L, float, Q> const &b) { return to_float(a) op b; } \
    template <length_t L, qualifier Q> inline vec<L, float, Q> operator op(vec<L, float, Q> const &a, vec<L, int, Q> const &b) { return a op to_float(b); } \
    template <length_t L, qualifier Q> inline vec<L, float, Q> operator op(int a, vec<L, float, Q> const &b) { return (float)a op b; }        \
    template <length_t L, qualifier Q> inline vec<L, float, Q> operator op(vec<L, float, Q> const &a, int b) { return a op (float)b; }
VHR_REF_MIXED(+)
VHR_REF_MIXED(-)
VHR_REF_MIXED(*)
VHR_REF_MIXED(/)
#undef VHR_REF_MIXED
}  // namespace glm

namespace glsl {
using namespace glm;

// ---- built-ins whose corner cases GLSL leaves to the implementation: pinned to what the GPUs the reference runs on do -----------
// (the reference needs VK_KHR_ray_tracing_pipeline hardware; README: developed on an RTX GPU). These non-template overloads win
// over glm's templates for scalar floats; make_ref.py re-declares them inside every shader namespace (gpu_*), where they hide the
// C library's and glm's scalar versions while glm's vector versions stay reachable through argument-dependent lookup.
//   max / min: FMNMX returns the non-NaN operand (IEEE maxNum / minNum); glm's `(x < y) ? y : x` would pass a NaN first operand on.
//   pow(x, y): computed as exp2(y * log2(x)) -> NaN for x < 0 (GLSL: "undefined if x < 0"); C's powf(-1, 128) would be +1.
inline float gpu_max(float a, float b) { return std::fmax(a, b); }
inline float gpu_min(float a, float b) { return std::fmin(a, b); }
inline float gpu_pow(float x, float y) { return x < 0.0f ? std::nanf("") : std::pow(x, y); }

// ---- fp16 <-> fp32 (image stores round to nearest even, VK_FORMAT_R16G16B16A16_SFLOAT / R16G16_SFLOAT) ---------------------------
inline uint16_t float_to_half(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7e00u | ((ax >> 13) & 0x3ffu));       // NaN
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                              // rounds to infinity
    if (ax < 0x33000001u) return (uint16_t)sign;                                           // <= 2^-25: zero
    if (ax < 0x38800000u) {                                                                // subnormal half
        const uint32_t shift = 126u - (ax >> 23), m = (ax & 0x7fffffu) | 0x800000u;
        uint32_t r = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (r & 1u))) ++r;
        return (uint16_t)(sign | r);
    }
    uint32_t r = (ax - 0x38000000u) >> 13;
    const uint32_t rem = ax & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    return (uint16_t)(sign | r);
}
inline float half_to_float(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        float f = (float)m * 5.9604644775390625e-8f;          // m * 2^-24, exact
        std::memcpy(&x, &f, 4);
        x |= sign;
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 112u) << 23) | (m << 13);
    }
    float f; std::memcpy(&f, &x, 4);
    return f;
}

// ---- images ------------------------------------------------------------------------------------------------------------------------
enum Format {                       // VkFormat values
    FMT_R8G8B8A8_UNORM = 37, FMT_R8G8B8A8_SRGB = 43, FMT_B8G8R8A8_UNORM = 44, FMT_B8G8R8A8_SRGB = 50,
    FMT_R16G16_SFLOAT = 83, FMT_R16G16B16A16_SFLOAT = 97, FMT_D32_SFLOAT = 126,
};
struct ImageDesc {
    const void *rd;     // texels imageLoad / texture read
    void *wr;           // texels imageStore writes (== rd unless the harness gives the dispatch a snapshot to read)
    int w, h;
    int format;
};
struct image2D { const ImageDesc *d = nullptr; };
struct sampler2D {
    const ImageDesc *d = nullptr;
    int mag = 1, min = 1;           // VkFilter: 0 NEAREST, 1 LINEAR (default sampler: resource_manager.cpp:58-69)
    int wrap_u = 0, wrap_v = 0;     // VkSamplerAddressMode: 0 REPEAT, 1 MIRRORED_REPEAT, 2 CLAMP_TO_EDGE, 3 CLAMP_TO_BORDER (opaque black)
};

inline float srgb8_to_linear(uint8_t c) {     // Khronos Data Format 13.3.1, evaluated in double
    const double e = (double)c / 255.0;
    return (float)(e <= 0.04045 ? e / 12.92 : std::pow((e + 0.055) / 1.055, 2.4));
}
inline vec4 fetch(const ImageDesc &d, int x, int y) {
    const size_t i = (size_t)y * d.w + x;
    switch (d.format) {
        case FMT_R16G16B16A16_SFLOAT: { const uint16_t *p = (const uint16_t *)d.rd + 4 * i; return vec4(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]), half_to_float(p[3])); }
        case FMT_R16G16_SFLOAT: { const uint16_t *p = (const uint16_t *)d.rd + 2 * i; return vec4(half_to_float(p[0]), half_to_float(p[1]), 0.0f, 1.0f); }
        case FMT_D32_SFLOAT: return vec4(((const float *)d.rd)[i], 0.0f, 0.0f, 1.0f);
        case FMT_B8G8R8A8_UNORM: { const uint8_t *p = (const uint8_t *)d.rd + 4 * i; return vec4((float)p[2] / 255.0f, (float)p[1] / 255.0f, (float)p[0] / 255.0f, (float)p[3] / 255.0f); }
        case FMT_R8G8B8A8_UNORM: { const uint8_t *p = (const uint8_t *)d.rd + 4 * i; return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f); }
        case FMT_R8G8B8A8_SRGB: { const uint8_t *p = (const uint8_t *)d.rd + 4 * i; return vec4(srgb8_to_linear(p[0]), srgb8_to_linear(p[1]), srgb8_to_linear(p[2]), (float)p[3] / 255.0f); }
    }
    return vec4(0.0f);
}
inline vec4 imageLoad(image2D im, ivec2 c) { return fetch(*im.d, c.x, c.y); }
inline uint8_t unorm8(float v) {                 // VK UNORM conversion: clamp, scale, round to nearest even
    v = v != v ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v));
    return (uint8_t)std::nearbyint(v * 255.0f);
}
inline float linear_to_srgb(float l) {           // VK_FORMAT_*_SRGB attachment store (fixed function, not shader code)
    l = l != l ? 0.0f : (l < 0.0f ? 0.0f : (l > 1.0f ? 1.0f : l));
    return l <= 0.0031308f ? 12.92f * l : 1.055f * std::pow(l, 1.0f / 2.4f) - 0.055f;
}
inline void store(const ImageDesc &d, int x, int y, vec4 v) {
    const size_t i = (size_t)y * d.w + x;
    switch (d.format) {
        case FMT_R16G16B16A16_SFLOAT: { uint16_t *p = (uint16_t *)d.wr + 4 * i; p[0] = float_to_half(v.x); p[1] = float_to_half(v.y); p[2] = float_to_half(v.z); p[3] = float_to_half(v.w); break; }
        case FMT_R16G16_SFLOAT: { uint16_t *p = (uint16_t *)d.wr + 2 * i; p[0] = float_to_half(v.x); p[1] = float_to_half(v.y); break; }
        case FMT_D32_SFLOAT: ((float *)d.wr)[i] = v.x; break;
        case FMT_B8G8R8A8_UNORM: { uint8_t *p = (uint8_t *)d.wr + 4 * i; p[2] = unorm8(v.x); p[1] = unorm8(v.y); p[0] = unorm8(v.z); p[3] = unorm8(v.w); break; }
    }
}
inline void imageStore(image2D im, ivec2 c, vec4 v) { store(*im.d, c.x, c.y, v); }
inline ivec2 imageSize(image2D im) { return ivec2(im.d->w, im.d->h); }
inline ivec2 textureSize(sampler2D s, int) { return ivec2(s.d->w, s.d->h); }

// ---- texture(): Vulkan spec "Texel Coordinate Systems", LOD 0 (every image here has one mip level; compute / ray-tracing stages have
// no implicit derivatives, and the full-screen fragment pass magnifies 1:1) -----------------------------------------------------------
inline int wrap_coord(int i, int n, int mode) {
    auto mod = [](int a, int b) { int m = a % b; return m < 0 ? m + b : m; };
    switch (mode) {
        case 0: return mod(i, n);                                                           // REPEAT
        case 1: { int m = mod(i, 2 * n) - n; m = m >= 0 ? m : -(1 + m); return (n - 1) - m; }  // MIRRORED_REPEAT
        case 2: return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);                                 // CLAMP_TO_EDGE
    }
    return (i < 0 || i >= n) ? -1 : i;                                                      // CLAMP_TO_BORDER
}
inline int finite_int(float f) { return (f == f && std::fabs(f) < 1e9f) ? (int)f : 0; }     // non-finite coordinate (undefined in Vulkan): texel 0
inline vec4 fetch_or_border(const ImageDesc &d, int x, int y) { return (x < 0 || y < 0) ? vec4(0.0f, 0.0f, 0.0f, 1.0f) : fetch(d, x, y); }
inline vec4 texture(sampler2D s, vec2 uv) {
    const ImageDesc &d = *s.d;
    if (s.mag == 0) {
        const int i = finite_int(std::floor(uv.x * (float)d.w)), j = finite_int(std::floor(uv.y * (float)d.h));
        return fetch_or_border(d, wrap_coord(i, d.w, s.wrap_u), wrap_coord(j, d.h, s.wrap_v));
    }
    // the texture unit holds the filter coordinate in fixed point with 8 fractional bits, rounded to nearest (VkPhysicalDeviceLimits::
    // subTexelPrecisionBits = 8 on the RTX hardware the reference needs; CUDA C Programming Guide, "Linear Filtering")
    const float u = std::floor((uv.x * (float)d.w - 0.5f) * 256.0f + 0.5f) * 0.00390625f, v = std::floor((uv.y * (float)d.h - 0.5f) * 256.0f + 0.5f) * 0.00390625f;
    const float fu = std::floor(u), fv = std::floor(v);
    const float a = u - fu, b = v - fv;
    const int i = finite_int(fu), j = finite_int(fv);
    const int x0 = wrap_coord(i, d.w, s.wrap_u), x1 = wrap_coord(i + 1, d.w, s.wrap_u);
    const int y0 = wrap_coord(j, d.h, s.wrap_v), y1 = wrap_coord(j + 1, d.h, s.wrap_v);
    const vec4 t00 = fetch_or_border(d, x0, y0), t10 = fetch_or_border(d, x1, y0), t01 = fetch_or_border(d, x0, y1), t11 = fetch_or_border(d, x1, y1);
    // tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a) b t01 + a b t11
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}

// ---- built-in variables -------------------------------------------------------------------------------------------------------------
struct uvec3_xy {                    // gl_GlobalInvocationID / gl_LaunchIDEXT / gl_LaunchSizeEXT: .x .y .z and the .xy the shaders take
    uint x = 0, y = 0, z = 0;
    uvec2 xy = uvec2(0u, 0u);
    void set(uint x_, uint y_, uint z_ = 0) { x = x_; y = y_; z = z_; xy = uvec2(x_, y_); }
};

// ---- ray tracing --------------------------------------------------------------------------------------------------------------------
const uint gl_RayFlagsNoneEXT = 0u, gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsNoOpaqueEXT = 2u, gl_RayFlagsTerminateOnFirstHitEXT = 4u,
           gl_RayFlagsSkipClosestHitShaderEXT = 8u;
// The driver's acceleration structure, bound to oracle/oracle_rt.cpp (vo_trace_any / vo_trace_closest): the ONLY part of the ray
// passes that is not reference source.
struct accelerationStructureEXT {
    const void *scene = nullptr;
    int (*trace_any)(const void *scene, const float *o, const float *d, float tmin, float tmax) = nullptr;
    int (*trace_closest)(const void *scene, const float *o, const float *d, float tmin, float tmax, double *t_u_v, uint32_t *geom_prim) = nullptr;
    // closest hit with an any-hit stage: accept(user, geometry index, primitive id, u, v) == 0 <=> ignoreIntersectionEXT
    int (*trace_closest_filtered)(const void *scene, const float *o, const float *d, float tmin, float tmax,
                                  int (*accept)(void *user, uint32_t geometry_index, uint32_t primitive_id, double u, double v), void *user, double *t_u_v,
                                  uint32_t *geom_prim) = nullptr;
};
struct RayHit { bool hit = false; float t = 0.0f; vec2 attribs = vec2(0.0f); int geometry_index = 0, primitive_id = 0; };
// `anyhit` != nullptr: the hit group's any-hit shader runs on every candidate (gl_RayFlagsNoOpaqueEXT geometry)
inline RayHit trace_query(const accelerationStructureEXT &as, uint flags, vec3 o, float tmin, vec3 d, float tmax,
                          int (*anyhit)(void *, uint32_t, uint32_t, double, double) = nullptr, void *user = nullptr) {
    RayHit r;
    const float of[3] = {o.x, o.y, o.z}, df[3] = {d.x, d.y, d.z};
    if (anyhit) {
        double tuv[3]; uint32_t gp[2];
        if (as.trace_closest_filtered(as.scene, of, df, tmin, tmax, anyhit, user, tuv, gp)) {
            r.hit = true; r.t = (float)tuv[0]; r.attribs = vec2((float)tuv[1], (float)tuv[2]);
            r.geometry_index = (int)gp[0]; r.primitive_id = (int)gp[1];
        }
        return r;
    }
    if (flags & gl_RayFlagsTerminateOnFirstHitEXT) {
        r.hit = as.trace_any(as.scene, of, df, tmin, tmax) != 0;
    } else {
        double tuv[3]; uint32_t gp[2];
        if (as.trace_closest(as.scene, of, df, tmin, tmax, tuv, gp)) {
            r.hit = true; r.t = (float)tuv[0]; r.attribs = vec2((float)tuv[1], (float)tuv[2]);
            r.geometry_index = (int)gp[0]; r.primitive_id = (int)gp[1];
        }
    }
    return r;
}

}  // namespace glsl
