// oracle/oracle_composition.cpp — TEST INFRASTRUCTURE ONLY. CPU restatement of the composition pass.
// Parity unpinned (see oracle_common.h). Citations relative to /root/reference.
//
//   vo_composition <- data/shaders/hybrid_render_path/composition.frag:60-161 drawn with composition.vert:5-8 as one
//                     full-screen triangle (src/render_paths/hybrid_render_path.cpp:333-379); shading helpers of
//                     data/shaders/common.glsl:116-150; attachment store = the swapchain's B8G8R8A8_SRGB
//                     (src/rendering_backend/vulkan_context.cpp:331) or, for parity measurements, linear fp16.
//
// texture() at a pixel centre through the default LINEAR sampler returns the texel itself; the shadow map
// is sampled at arbitrary coordinates, bilinear + REPEAT like oracle_svgf.cpp's sample_depth.
#include "oracle_common.h"

#include <cmath>

using namespace vo;

namespace {
inline int wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
inline void bilinear_setup(float u, int n, int &i0, int &i1, float &a) {
    // NVIDIA texture units — the hardware the reference needs (RTX) — hold the filter coordinate in fixed point with 8 fractional bits
    // (VkPhysicalDeviceLimits::subTexelPrecisionBits = 8; CUDA C Programming Guide, "Linear Filtering": 9-bit fixed point with 8 bits of
    // fractional value, 1.0 exactly representable). Consequence: a lookup at a texel centre returns that texel even when
    // fl(fl((x + .5) / W) * W) is an ulp off x + .5 (51 of the 1920 columns at 1080p) — SURVEY Q17.
    float uu = std::floor((u * (float)n - 0.5f) * 256.0f + 0.5f) * 0.00390625f;
    float fl = std::floor(uu);
    a = uu - fl;
    int i = (fl == fl && std::fabs(fl) < 1e9f) ? (int)fl : 0;
    i0 = wrap(i, n);
    i1 = wrap(i + 1, n);
}
inline float sample_r32f(const float *img, int W, int H, float u, float v) {
    int x0, x1, y0, y1; float a, b;
    bilinear_setup(u, W, x0, x1, a);
    bilinear_setup(v, H, y0, y1, b);
    float t00 = img[(size_t)y0 * W + x0], t10 = img[(size_t)y0 * W + x1];
    float t01 = img[(size_t)y1 * W + x0], t11 = img[(size_t)y1 * W + x1];
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
// Vulkan float -> UNORM8: NaN -> 0, clamp, round to nearest (ties to even)
inline uint8_t unorm8(float c) {
    c = (c == c) ? clampf(c, 0.0f, 1.0f) : 0.0f;
    return (uint8_t)std::nearbyint(c * 255.0f);
}
inline float srgb_encode(float c) {
    c = (c == c) ? clampf(c, 0.0f, 1.0f) : 0.0f;
    return c <= 0.0031308f ? 12.92f * c : 1.055f * std::pow(c, 1.0f / 2.4f) - 0.055f;
}
}  // namespace

extern "C" {

// out_format: 50 = B8G8R8A8_SRGB, 44 = B8G8R8A8_UNORM (4 bytes / pixel), 97 = R16G16B16A16_SFLOAT (linear, NaN -> 0).
// rt_channels: 2 = raw RG16F "Raytraced Shadows and Ambient Occlusion", 4 = the denoised RGBA16F image.
void vo_composition(const PerFrameData *pfd_, int W, int H, int shadow_mode, int ao_mode, int reflection_mode,
                    const uint8_t *albedo_bgra8, const uint16_t *normals, const uint16_t *motion, const float *depth,
                    const float *shadow_map, int shadow_w, int shadow_h, const uint16_t *ssao, const uint16_t *ssr,
                    const uint16_t *rt, int rt_channels, const uint16_t *refl, int out_format, void *out) {
    const PerFrameData &pfd = *pfd_;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const size_t pix = (size_t)y * W + x;
            const vec2 uv = {((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H};
            const uint8_t *a8 = albedo_bgra8 + pix * 4;
            const vec3 albedo = {(float)a8[2] / 255.0f, (float)a8[1] / 255.0f, (float)a8[0] / 255.0f};
            const float d = depth[pix];
            const vec3 P = get_world_space_position(pfd, d, uv);
            const vec4 n4 = load_rgba16f(normals, W, x, y);
            const vec3 N = {n4.x, n4.y, n4.z};
            const vec4 m4 = load_rgba16f(motion, W, x, y);
            const vec2 metallic_roughness = {m4.z, m4.w};

            vec2 rsa = {1.0f, 1.0f};
            if (shadow_mode == 0 || ao_mode == 0) {
                vec4 t = rt_channels == 2 ? load_rg16f(rt, W, x, y) : load_rgba16f(rt, W, x, y);
                rsa = {t.x, t.y};
            }
            const vec3 cam = {pfd.camera_view_inverse[12], pfd.camera_view_inverse[13], pfd.camera_view_inverse[14]};
            const vec3 V = normalize(cam - P);
            const vec3 L = -v3(pfd.directional_light.direction[0], pfd.directional_light.direction[1], pfd.directional_light.direction[2]);
            const vec3 Hh = normalize(L + V);

            float shadow = 1.0f;
            if (shadow_mode == 0) {
                shadow = rsa.x;
            } else if (shadow_mode == 1) {
                // SHADOW_BIAS_MATRIX * projview (mat4 * mat4 first, common.glsl:6-11), then * vec4(P, 1)
                const float *pv = pfd.directional_light.projview;
                float M[16];
                for (int c = 0; c < 4; ++c) {
                    M[c * 4 + 0] = 0.5f * pv[c * 4 + 0] + 0.5f * pv[c * 4 + 3];
                    M[c * 4 + 1] = 0.5f * pv[c * 4 + 1] + 0.5f * pv[c * 4 + 3];
                    M[c * 4 + 2] = pv[c * 4 + 2];
                    M[c * 4 + 3] = pv[c * 4 + 3];
                }
                const vec4 pl = mul44(M, vec4{P.x, P.y, P.z, 1.0f});
                const vec4 sc = {pl.x / pl.w, pl.y / pl.w, pl.z / pl.w, 1.0f};
                const float scale = 1.0f / 4096.0f;
                float acc = 0.0f;
                for (int i = 0; i < 16; ++i) {
                    const float ox = -1.5f + (float)(i >> 2), oy = -1.5f + (float)(i & 3);     // composition.frag:89-94
                    const float ds = sample_r32f(shadow_map, shadow_w, shadow_h, sc.x + ox * scale, sc.y + oy * scale);
                    acc += (sc.z < ds - 1e-4f) ? 0.0f : 1.0f;
                }
                shadow = acc / 16.0f;
            }
            float ao = 1.0f;
            if (ao_mode == 0) ao = rsa.y;
            else if (ao_mode == 1) ao = load_rgba16f(ssao, W, x, y).x;

            const float metallic = clampf(metallic_roughness.x, 0.0f, 1.0f);
            const float roughness = clampf(metallic_roughness.y, 0.04f, 1.0f);
            const float ambient_factor = PI_INVERSE_F;
            const vec3 li = {pfd.directional_light.intensity[0], pfd.directional_light.intensity[1], pfd.directional_light.intensity[2]};
            const vec3 lc = {pfd.directional_light.color[0], pfd.directional_light.color[1], pfd.directional_light.color[2]};
            const vec3 f0 = {mixf(0.04f, albedo.x, metallic), mixf(0.04f, albedo.y, metallic), mixf(0.04f, albedo.z, metallic)};
            const vec3 F = fresnel_schlick(f0, Hh, V);
            const float ndl = gl_max(dot(N, L), 0.0f);

            const vec3 ambient = (albedo * ao) * ambient_factor;
            const vec3 diffuse = (((diffuse_brdf(metallic, albedo, F) * ndl) * li) * lc) * shadow;
            vec3 specular = (((specular_brdf(roughness, F, V, L, N, Hh) * ndl) * li) * lc) * shadow;
            if (reflection_mode == 0 || reflection_mode == 1) {
                const vec4 r4 = load_rgba16f(reflection_mode == 0 ? refl : ssr, W, x, y);
                const vec3 reflections = v3(r4.x, r4.y, r4.z) * shadow;
                if (metallic == 1.0f) specular = reflections;
                else specular = {mixf(specular.x, reflections.x, roughness), mixf(specular.y, reflections.y, roughness),
                                 mixf(specular.z, reflections.z, roughness)};
            }
            const vec3 lighting = (ambient + diffuse) + specular;
            if (out_format == 97) {
                auto nz = [](float c) { return c == c ? c : 0.0f; };
                store_rgba16f(reinterpret_cast<uint16_t *>(out), W, x, y, vec4{nz(lighting.x), nz(lighting.y), nz(lighting.z), 1.0f});
            } else {
                float r = lighting.x, g = lighting.y, b = lighting.z;
                if (out_format == 50) { r = srgb_encode(r); g = srgb_encode(g); b = srgb_encode(b); }
                uint8_t *o = reinterpret_cast<uint8_t *>(out) + pix * 4;
                o[0] = unorm8(b); o[1] = unorm8(g); o[2] = unorm8(r); o[3] = 255;
            }
        }
    }
}

}  // extern "C"
