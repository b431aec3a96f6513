// oracle/oracle_ssr.cpp — TEST INFRASTRUCTURE ONLY. CPU restatement of the screen-space reflection pass.
// Parity unpinned (see oracle_common.h). Citations relative to /root/reference.
//
//   vo_ssr <- data/shaders/hybrid_render_path/ssr.comp:15-137, dispatched by the "SSR Pass" node
//             (src/render_paths/hybrid_render_path.cpp:202-243) with SSRPushConstants
//             (src/rendering_backend/glsl_common.h:41-46; defaults 25 / 0.1 / 0.5 / 10, hybrid_render_path.cpp:203-208).
//
// texture() goes through the default sampler (resource_manager.cpp:58-69: LINEAR, REPEAT, one mip level; compute
// shaders sample LOD 0) with the Vulkan spec's float weights, like oracle_svgf.cpp's SSAO.
#include "oracle_common.h"

#include <cmath>

using namespace vo;

namespace {

inline int wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
struct Taps { int x0, x1, y0, y1; float a, b; };
inline void axis(float u, int n, int &i0, int &i1, float &a) {
    // NVIDIA texture units — the hardware the reference needs (RTX) — hold the filter coordinate in fixed point with 8 fractional bits
    // (VkPhysicalDeviceLimits::subTexelPrecisionBits = 8; CUDA C Programming Guide, "Linear Filtering": 9-bit fixed point with 8 bits of
    // fractional value, 1.0 exactly representable). Consequence: a lookup at a texel centre returns that texel even when
    // fl(fl((x + .5) / W) * W) is an ulp off x + .5 (51 of the 1920 columns at 1080p) — SURVEY Q17.
    float uu = std::floor((u * (float)n - 0.5f) * 256.0f + 0.5f) * 0.00390625f;
    float fl = std::floor(uu);
    a = uu - fl;
    int i = (fl == fl && std::fabs(fl) < 1e9f) ? (int)fl : 0;      // non-finite coordinate: texel 0 (undefined in the reference)
    i0 = wrap(i, n);
    i1 = wrap(i + 1, n);
}
inline Taps taps(int W, int H, vec2 uv) {
    Taps t;
    axis(uv.x, W, t.x0, t.x1, t.a);
    axis(uv.y, H, t.y0, t.y1, t.b);
    return t;
}
inline float lerp4(const Taps &t, float t00, float t10, float t01, float t11) {
    return (1 - t.a) * (1 - t.b) * t00 + t.a * (1 - t.b) * t10 + (1 - t.a) * t.b * t01 + t.a * t.b * t11;
}

struct Frame {
    const PerFrameData &pfd;
    int W, H;
    const uint8_t *albedo;      // B8G8R8A8_UNORM
    const uint16_t *normals, *motion;
    const float *depth;
    float pv[16];               // camera_proj * camera_view (ssr.comp:23: `pfd.camera_proj * pfd.camera_view * vec4(v, 1.0)`, left to right)

    float tex_depth(vec2 uv) const {
        Taps t = taps(W, H, uv);
        return lerp4(t, depth[(size_t)t.y0 * W + t.x0], depth[(size_t)t.y0 * W + t.x1], depth[(size_t)t.y1 * W + t.x0], depth[(size_t)t.y1 * W + t.x1]);
    }
    vec4 tex_half4(const uint16_t *img, vec2 uv) const {
        Taps t = taps(W, H, uv);
        vec4 t00 = load_rgba16f(img, W, t.x0, t.y0), t10 = load_rgba16f(img, W, t.x1, t.y0);
        vec4 t01 = load_rgba16f(img, W, t.x0, t.y1), t11 = load_rgba16f(img, W, t.x1, t.y1);
        return {lerp4(t, t00.x, t10.x, t01.x, t11.x), lerp4(t, t00.y, t10.y, t01.y, t11.y), lerp4(t, t00.z, t10.z, t01.z, t11.z),
                lerp4(t, t00.w, t10.w, t01.w, t11.w)};
    }
    vec3 tex_albedo(vec2 uv) const {
        Taps t = taps(W, H, uv);
        auto rgb = [&](int x, int y) {
            const uint8_t *c = albedo + ((size_t)y * W + x) * 4;
            return vec3{(float)c[2] / 255.0f, (float)c[1] / 255.0f, (float)c[0] / 255.0f};
        };
        vec3 t00 = rgb(t.x0, t.y0), t10 = rgb(t.x1, t.y0), t01 = rgb(t.x0, t.y1), t11 = rgb(t.x1, t.y1);
        return {lerp4(t, t00.x, t10.x, t01.x, t11.x), lerp4(t, t00.y, t10.y, t01.y, t11.y), lerp4(t, t00.z, t10.z, t01.z, t11.z)};
    }
    // ssr.comp:22-26
    vec2 world_space_to_uv(vec3 v) const {
        vec4 clip = mul44(pv, vec4{v.x, v.y, v.z, 1.0f});
        return {(clip.x / clip.w) * 0.5f + 0.5f, (clip.y / clip.w) * 0.5f + 0.5f};
    }
    // ssr.comp:29-59
    vec3 compute_lighting(vec2 uv) const {
        vec3 albedo_ = tex_albedo(uv);
        vec3 position = get_world_space_position(pfd, tex_depth(uv), uv);
        vec4 mv = tex_half4(motion, uv);
        vec2 metallic_roughness = {mv.z, mv.w};
        vec3 camera_position = v3(pfd.camera_view_inverse[12], pfd.camera_view_inverse[13], pfd.camera_view_inverse[14]);
        vec3 V = normalize(camera_position - position);
        vec3 L = -v3(pfd.directional_light.direction[0], pfd.directional_light.direction[1], pfd.directional_light.direction[2]);
        vec4 n4 = tex_half4(normals, uv);
        vec3 N = v3(n4.x, n4.y, n4.z);
        vec3 H = normalize(L + V);
        float min_roughness = 0.04f;
        float metallic = std::fmin(std::fmax(metallic_roughness.x, 0.0f), 1.0f);
        float roughness = std::fmin(std::fmax(metallic_roughness.y, min_roughness), 1.0f);
        float ambient_factor = PI_INVERSE_F * 0.2f;
        vec3 li = v3(pfd.directional_light.intensity[0], pfd.directional_light.intensity[1], pfd.directional_light.intensity[2]);
        vec3 lc = v3(pfd.directional_light.color[0], pfd.directional_light.color[1], pfd.directional_light.color[2]);
        vec3 f0 = v3(mixf(0.04f, albedo_.x, metallic), mixf(0.04f, albedo_.y, metallic), mixf(0.04f, albedo_.z, metallic));
        vec3 F = fresnel_schlick(f0, H, V);
        vec3 ambient_lighting = albedo_ * ambient_factor;
        vec3 diffuse_lighting = diffuse_brdf(metallic, albedo_, F);
        vec3 specular_lighting = specular_brdf(roughness, F, V, L, N, H);
        return ambient_lighting + (diffuse_lighting + specular_lighting) * gl_max(dot(N, L), 0.0f) * li * lc;
    }
};

inline float distance(vec3 a, vec3 b) { return length(a - b); }

}  // namespace

extern "C" {

// ssr.comp:61-137 over rows [y0, y1). out = RGBA16F "Screen Space Reflections" (the shader declares r16f, the image is
// R16G16B16A16_SFLOAT, hybrid_render_path.cpp:219: all four channels are stored).
void vo_ssr(const PerFrameData *pfd_, int W, int H, int y0, int y1, float ray_distance, float step_size, float thickness, int bsearch_steps,
            const uint8_t *albedo_bgra8, const uint16_t *normals, const uint16_t *motion, const float *depth, uint16_t *out) {
    const PerFrameData &pfd = *pfd_;
    Frame f{pfd, W, H, albedo_bgra8, normals, motion, depth, {}};
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float acc = pfd.camera_proj[0 * 4 + r] * pfd.camera_view[c * 4 + 0];
            for (int k = 1; k < 4; ++k) acc = acc + pfd.camera_proj[k * 4 + r] * pfd.camera_view[c * 4 + k];
            f.pv[c * 4 + r] = acc;
        }
    const float q = ray_distance / step_size;
    const int n_steps = (q == q && q > 0.0f) ? (int)std::fmin(q, 1048576.0f) : 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int gy = y0; gy < y1; ++gy) {
        for (int gx = 0; gx < W; ++gx) {
            store_rgba16f(out, W, gx, gy, vec4{0, 0, 0, 0});                                       // :62-66
            vec2 coords = {(float)gx * pfd.display_size_inverse[0], (float)gy * pfd.display_size_inverse[1]};
            float fragment_depth = f.tex_depth(coords);
            vec3 camera_position = v3(pfd.camera_view_inverse[12], pfd.camera_view_inverse[13], pfd.camera_view_inverse[14]);
            vec3 P = get_world_space_position(pfd, fragment_depth, coords);
            vec4 n4 = f.tex_half4(normals, coords);
            vec3 N = v3(n4.x, n4.y, n4.z);
            vec3 I = normalize(P - camera_position);
            vec3 reflected_dir = normalize(I - N * (2.0f * dot(N, I)));                            // reflect(I, N)

            bool found = false;
            float prev_step = 0.0f, final_step = 0.0f;
            for (int i = 0; i < n_steps; ++i) {                                                    // :89-108
                float offset = step_size * (float)i;
                vec3 ray_position = P + reflected_dir * offset;
                float distance_to_ray = distance(camera_position, ray_position);
                vec2 sample_uv = f.world_space_to_uv(ray_position);
                vec3 screen_position = get_world_space_position(pfd, f.tex_depth(sample_uv), sample_uv);
                float distance_to_screen = distance(camera_position, screen_position);
                float delta_distance = distance_to_ray - distance_to_screen;
                if (delta_distance > 0.3f && delta_distance < thickness) {
                    final_step = offset;
                    found = true;
                    break;
                } else {
                    prev_step = offset;
                }
            }
            if (!found) continue;                                                                  // :110-112

            float mid_step = (prev_step + final_step) * 0.5f;                                      // :115-135
            vec2 final_uv = {0.0f, 0.0f};
            for (int i = 0; i < bsearch_steps; ++i) {
                float offset = mid_step;
                vec3 ray_position = P + reflected_dir * offset;
                float distance_to_ray = distance(camera_position, ray_position);
                final_uv = f.world_space_to_uv(ray_position);
                vec3 screen_position = get_world_space_position(pfd, f.tex_depth(final_uv), final_uv);
                float distance_to_screen = distance(camera_position, screen_position);
                float delta_distance = distance_to_ray - distance_to_screen;
                if (delta_distance > 0.3f && delta_distance < thickness) {
                    mid_step = (prev_step + mid_step) * 0.5f;
                } else {
                    float tmp = mid_step;
                    mid_step = mid_step + (mid_step - prev_step);
                    prev_step = tmp;
                }
            }
            vec3 c = f.compute_lighting(final_uv);
            store_rgba16f(out, W, gx, gy, vec4{c.x, c.y, c.z, 1.0f});
        }
    }
}

}  // extern "C"
