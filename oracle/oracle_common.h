// oracle/oracle_common.h — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference shaders).
//
// PINNED BY oracle/_ref: the reference (RMichelsen/VulkanHybridRenderer) ships no tests, golden images or known-answer
// vectors, but its shader files themselves compile for the CPU through oracle/ref_shim.h + the reference's vendored glm
// (oracle/make_ref.py -> oracle/_ref/libvhr_ref.so). tests/test_ref_pinning_cpu.py holds this restatement BIT-IDENTICAL to
// those compiled shaders for every pass both cover (common.glsl helpers, svgf.comp, svgf_atrous_filter.comp, ssao.comp,
// ssao_blur.comp, ssr.comp, composition.frag, raygen.rgen + miss shaders + reflection_hit.rchit), and tests/golden/ holds
// outputs of the compiled shaders. Still third-party and NOT pinned: the driver's BVH traversal / ray-triangle arithmetic
// behind traceRayEXT — both sides query the double-precision BVH of oracle_rt.cpp — and the rasteriser
// (the G-buffer producer restates gbuf.vert / gbuf.frag encodings on primary rays).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library. The product path (vulkanhybridrenderer_b200/) never links or calls it.
//
// All citations are relative to /root/reference.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace vo {

// ---------------------------------------------------------------------------------------------
// Host/shader shared structs — src/rendering_backend/glsl_common.h:31-99 (std140/scalar layouts
// verified: sizeof(PerFrameData)==584, Vertex 56, Material 44, Primitive 120, SVGFPushConstants 24).
// Matrices are glm column-major: m[c*4+r].
// ---------------------------------------------------------------------------------------------
struct DirectionalLight {   // glsl_common.h:52-57
    float projview[16];
    float direction[4];
    float color[4];
    float intensity[4];
};
struct PerFrameData {       // glsl_common.h:59-72
    float camera_view[16];
    float camera_proj[16];
    float camera_view_inverse[16];
    float camera_proj_inverse[16];
    float camera_viewproj_inverse[16];
    float camera_view_prev_frame[16];
    float camera_proj_prev_frame[16];
    DirectionalLight directional_light;
    float display_size[2];
    float display_size_inverse[2];
    uint32_t frame_index;
    int32_t blue_noise_texture_index;
};
static_assert(sizeof(PerFrameData) == 584, "PerFrameData layout");

struct Vertex {             // glsl_common.h:74-80
    float pos[3];
    float normal[3];
    float tangent[4];
    float uv0[2];
    float uv1[2];
};
static_assert(sizeof(Vertex) == 56, "Vertex layout");

struct Material {           // glsl_common.h:82-91
    float base_color[4];
    int32_t base_color_texture;
    int32_t metallic_roughness_texture;
    int32_t normal_map;
    float metallic_factor;
    float roughness_factor;
    int32_t alpha_mask;
    float alpha_cutoff;
};
static_assert(sizeof(Material) == 44, "Material layout");

struct Primitive {          // glsl_common.h:93-99
    float transform[16];
    Material material;
    uint32_t vertex_offset;
    uint32_t index_offset;
    uint32_t index_count;
};
static_assert(sizeof(Primitive) == 120, "Primitive layout");

// ---------------------------------------------------------------------------------------------
// fp16 storage emulation. Every inter-pass image of the hot path is half precision
// (hybrid_render_path.cpp:16-18,109-110,247-261); stores round to nearest even.
// ---------------------------------------------------------------------------------------------
static inline uint16_t f2h(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) {                       // inf / nan
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x200u | ((ax >> 13) & 0x3ffu)) : 0u));
    }
    if (ax >= 0x477ff000u) {                       // >= 65520 rounds to inf
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) {                        // subnormal half or zero (|f| < 2^-14)
        if (ax < 0x33000000u) return (uint16_t)sign;    // < 2^-25 -> 0  (2^-25 itself ties to even = 0)
        uint32_t e = ax >> 23;                     // biased exponent, 102..112
        uint32_t m = (ax & 0x7fffffu) | 0x800000u; // 24-bit significand
        uint32_t shift = 126u - e;                 // 14..24 : result = m >> shift in units of 2^-24
        uint32_t r = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (r & 1u))) r++;
        return (uint16_t)(sign | r);
    }
    uint32_t r = ((ax - 0x38000000u) >> 13);       // rebias 127->15, drop 13 bits
    uint32_t rem = ax & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
    return (uint16_t)(sign | r);
}

static inline float h2f(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) {
            x = sign;
        } else {                                    // subnormal: value = m * 2^-24
            float f = (float)m * 5.9604644775390625e-8f;
            std::memcpy(&x, &f, 4);
            x |= sign;
        }
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 112u) << 23) | (m << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

struct half4 { uint16_t x, y, z, w; };
struct half2 { uint16_t x, y; };
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

static inline vec4 load_rgba16f(const uint16_t *img, int W, int x, int y) {
    const uint16_t *p = img + ((size_t)y * W + x) * 4;
    return vec4{h2f(p[0]), h2f(p[1]), h2f(p[2]), h2f(p[3])};
}
static inline void store_rgba16f(uint16_t *img, int W, int x, int y, vec4 v) {
    uint16_t *p = img + ((size_t)y * W + x) * 4;
    p[0] = f2h(v.x); p[1] = f2h(v.y); p[2] = f2h(v.z); p[3] = f2h(v.w);
}
// A two-channel image read through imageLoad returns (r, g, 0, 1) — SURVEY Q2/Q13.
static inline vec4 load_rg16f(const uint16_t *img, int W, int x, int y) {
    const uint16_t *p = img + ((size_t)y * W + x) * 2;
    return vec4{h2f(p[0]), h2f(p[1]), 0.0f, 1.0f};
}
static inline void store_rg16f(uint16_t *img, int W, int x, int y, vec4 v) {
    uint16_t *p = img + ((size_t)y * W + x) * 2;
    p[0] = f2h(v.x); p[1] = f2h(v.y);
}

// ---------------------------------------------------------------------------------------------
// GLSL scalar/vector semantics used by the shaders (compiled with -ffp-contract=off).
// ---------------------------------------------------------------------------------------------
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }   // GLSL mix()
static inline float fractf(float x) { return x - std::floor(x); }                          // GLSL fract()
// GLSL max() on NVIDIA hardware (FMNMX) returns the non-NaN operand; C fmaxf has the same rule.
static inline float gl_max(float a, float b) { return std::fmax(a, b); }
// float -> int conversion: GLSL leaves out-of-range undefined; we pin it to the CUDA cvt.rzi.s32.f32
// rule (truncate, saturate, NaN -> 0) so oracle and kernels agree on degenerate motion vectors.
static inline int f2i_rz(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}

static inline vec3 v3(float x, float y, float z) { return vec3{x, y, z}; }
static inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
static inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
// GLSL normalize() = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x): the formula of the reference's own math library
// (dependencies/glm/detail/func_geometric.inl:88, func_exponential.inl:138), which is what oracle/_ref compiles the shaders with.
static inline vec3 normalize(vec3 a) { float r = 1.0f / std::sqrt(dot(a, a)); return {a.x * r, a.y * r, a.z * r}; }
static inline vec3 cross(vec3 a, vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// mat4 * vec4, column-major, summed pairwise: (col0*x + col1*y) + (col2*z + col3*w) — the order of the reference's own math library
// (dependencies/glm/detail/type_mat4x4.inl:561-571), i.e. of oracle/_ref. (GLSL itself leaves the order to the compiler.)
static inline vec4 mul44(const float *m, vec4 v) {
    vec4 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8]  * v.z + m[12] * v.w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9]  * v.z + m[13] * v.w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
    r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
    return r;
}
// mat3(m4) * vec3
static inline vec3 mul33_of44(const float *m, vec3 v) {
    vec3 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8]  * v.z;
    r.y = m[1] * v.x + m[5] * v.y + m[9]  * v.z;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z;
    return r;
}

// glsl_common.h:111-116 get_view_space_position
static inline vec3 get_view_space_position(const PerFrameData &pfd, float depth, vec2 uv) {
    vec4 p = mul44(pfd.camera_proj_inverse, vec4{uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, depth, 1.0f});
    return {p.x / p.w, p.y / p.w, p.z / p.w};
}
// glsl_common.h:118-122 get_world_space_position
static inline vec3 get_world_space_position(const PerFrameData &pfd, float depth, vec2 uv) {
    vec4 p = mul44(pfd.camera_viewproj_inverse, vec4{uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, depth, 1.0f});
    return {p.x / p.w, p.y / p.w, p.z / p.w};
}

// ---------------------------------------------------------------------------------------------
// data/shaders/common.glsl
// ---------------------------------------------------------------------------------------------
static const float COS_PI_4 = 0.70710678118654752440084f;
static const float PI_F = 3.14159265358979323846264f;
static const float TWO_PI_F = 6.28318530717958647692528f;
static const float PI_INVERSE_F = 0.31830988618379067153776f;

// common.glsl:47-56
static inline uint32_t seed_thread(uint32_t seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
// common.glsl:58-64
static inline uint32_t random_u32(uint32_t &state) {
    state ^= (state << 13);
    state ^= (state >> 17);
    state ^= (state << 5);
    return state;
}
// common.glsl:66-68
static inline float random01(uint32_t &state) {
    uint32_t b = 0x3f800000u | (random_u32(state) >> 9);
    float f;
    std::memcpy(&f, &b, 4);
    return f - 1.0f;
}
// common.glsl:29-34
static inline vec3 uniform_sample_cone(vec2 u, float cos_theta_max) {
    float cos_theta = (1.0f - u.x) + u.x * cos_theta_max;
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    float phi = u.y * TWO_PI_F;
    return {std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta};
}
// common.glsl:37-42
static inline vec3 uniform_sample_cosine_weighted_hemisphere(vec2 u) {
    float x = std::sqrt(u.x) * std::cos(TWO_PI_F * u.y);
    float y = std::sqrt(u.x) * std::sin(TWO_PI_F * u.y);
    float z = std::sqrt(1.0f - u.x);
    return {x, y, z};
}
// common.glsl:80-93 (Frisvad); returns columns M[0], M[1], M[2]
struct mat3 { vec3 c0, c1, c2; };
static inline mat3 onb_from_unit_vector(vec3 n) {
    mat3 M;
    M.c2 = n;
    if (n.z < -0.9999999f) {
        M.c0 = {0.0f, -1.0f, 0.0f};
        M.c1 = {-1.0f, 0.0f, 0.0f};
        return M;
    }
    float a = 1.0f / (1.0f + n.z);
    float b = -n.x * n.y * a;
    M.c0 = {1.0f - n.x * n.x * a, b, -n.x};
    M.c1 = {b, 1.0f - n.y * n.y * a, -n.y};
    return M;
}
static inline vec3 mul(const mat3 &M, vec3 v) {   // M * v = c0*v.x + c1*v.y + c2*v.z
    return {M.c0.x * v.x + M.c1.x * v.y + M.c2.x * v.z,
            M.c0.y * v.x + M.c1.y * v.y + M.c2.y * v.z,
            M.c0.z * v.x + M.c1.z * v.y + M.c2.z * v.z};
}

// common.glsl:116-150
// GLSL evaluates `(1 - f0) * (1 - HdotV) * ...` left to right; restated in that order.
static inline vec3 fresnel_schlick(vec3 f0, vec3 H, vec3 V) {
    float hv = gl_max(dot(H, V), 0.0f);
    float o = 1.0f - hv;
    vec3 r;
    r.x = f0.x + (1.0f - f0.x) * o * o * o * o * o;
    r.y = f0.y + (1.0f - f0.y) * o * o * o * o * o;
    r.z = f0.z + (1.0f - f0.z) * o * o * o * o * o;
    return r;
}
static inline float D_GGX(float roughness, vec3 N, vec3 H) {
    float a2 = roughness * roughness;
    float nh = gl_max(dot(N, H), 0.0f);
    float f = nh * nh * (a2 - 1.0f) + 1.0f;
    return a2 / (PI_F * f * f);
}
static inline float G_GGX(float roughness, vec3 N, vec3 V, vec3 L) {
    float k = ((roughness + 1.0f) * (roughness + 1.0f)) * 0.125f;
    float nv = gl_max(dot(N, V), 0.0f);
    float nl = gl_max(dot(N, L), 0.0f);
    float g_nvk = nv / (nv * (1.0f - k) + k);
    float g_nlk = nl / (nl * (1.0f - k) + k);
    return g_nvk * g_nlk;
}
static inline vec3 specular_brdf(float roughness, vec3 F, vec3 V, vec3 L, vec3 N, vec3 H) {
    float dg = D_GGX(roughness, N, H) * G_GGX(roughness, N, V, L);
    vec3 DFG = {dg * F.x, dg * F.y, dg * F.z};
    float denom = 4.0f * gl_max(dot(N, V), 0.0f) * gl_max(dot(N, L), 0.0f);
    float d = gl_max(denom, 1e-6f);
    return {DFG.x / d, DFG.y / d, DFG.z / d};
}
static inline vec3 diffuse_brdf(float metallic, vec3 albedo, vec3 F) {
    vec3 dp = {1.0f - F.x, 1.0f - F.y, 1.0f - F.z};
    float s = 1.0f - metallic;
    dp = dp * s;
    return {(dp.x * albedo.x) / PI_F, (dp.y * albedo.y) / PI_F, (dp.z * albedo.z) / PI_F};
}

}  // namespace vo
