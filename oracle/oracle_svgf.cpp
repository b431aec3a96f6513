// oracle/oracle_svgf.cpp — TEST INFRASTRUCTURE ONLY. CPU restatement of the SVGF and SSAO shaders.
// Parity unpinned (see oracle_common.h). Citations relative to /root/reference.
//
//   vo_svgf_temporal     <- data/shaders/hybrid_render_path/svgf.comp:16-145
//   vo_svgf_atrous       <- data/shaders/hybrid_render_path/svgf_atrous_filter.comp:17-103
//   vo_svgf_pass         <- src/render_paths/hybrid_render_path.cpp:288-330 (dispatch/blit/ping-pong order)
//   vo_ssao / vo_ssao_blur <- data/shaders/hybrid_render_path/ssao.comp:14-53, ssao_blur.comp:11-26
//
// Images are dense row-major arrays of IEEE half bit patterns (uint16_t), row 0 = NDC y -1.
#include "oracle_common.h"

#include <algorithm>
#include <vector>

using namespace vo;

namespace {

// svgf.comp:16-39
inline bool is_valid_reprojection(const PerFrameData &pfd, const uint16_t *prev_normals, int W, int px, int py,
                                  int current_object_id, vec3 current_normal) {
    if (px < 0 || py < 0 || (float)px >= pfd.display_size[0] || (float)py >= pfd.display_size[1]) return false;
    vec4 pn = load_rgba16f(prev_normals, W, px, py);
    int prev_object_id = f2i_rz(pn.w);
    if (current_object_id != prev_object_id) return false;
    if (dot(current_normal, v3(pn.x, pn.y, pn.z)) < COS_PI_4) return false;
    return true;
}

// svgf_atrous_filter.comp:40-51
inline float edge_stopping_normal(vec3 np, vec3 nq) {
    // pow() with a non-positive base is undefined in GLSL; NVIDIA yields NaN -> max(0,NaN)=0.
    float d = dot(np, nq);
    if (!(d > 0.0f)) return 0.0f;
    return gl_max(0.0f, std::pow(d, 128.0f));
}
inline float edge_stopping_luminance(float variance_p, float lp, float lq) {
    float e = std::fabs(lp - lq) / (4.0f * std::sqrt(variance_p) + 1e-6f);
    return std::exp(-e);
}

}  // namespace

extern "C" {

// svgf.comp:41-145. Snapshot semantics for the moments image: reads `moments_in`
// (previous frame), writes `moments_out`.
void vo_svgf_temporal(const PerFrameData *pfd_, int W, int H,
                      const uint16_t *normals,       // RGBA16F "World Space Normals and Object IDs"
                      const uint16_t *motion,        // RGBA16F "Motion Vectors and Metallic Roughness"
                      const uint16_t *rt,            // RG16F   "Raytraced Shadows and Ambient Occlusion"
                      const uint16_t *prev_normals,  // RGBA16F prev_frame_normals_and_object_ids
                      const uint16_t *history,       // RGBA16F shadow_and_ao_history
                      const uint16_t *moments_in,    // RG16F   shadow_and_ao_moments_history (previous frame)
                      uint16_t *integrated_out,      // RGBA16F integrated_shadow_and_ao[0]
                      uint16_t *moments_out) {       // RG16F   shadow_and_ao_moments_history (this frame)
    const PerFrameData &pfd = *pfd_;
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < H; ++cy) {
        for (int cx = 0; cx < W; ++cx) {
            vec4 cn = load_rgba16f(normals, W, cx, cy);
            vec3 current_normal = v3(cn.x, cn.y, cn.z);
            int current_object_id = f2i_rz(cn.w);
            vec4 mv = load_rgba16f(motion, W, cx, cy);
            vec4 cur = load_rg16f(rt, W, cx, cy);
            float current_shadow = cur.x, current_ao = cur.y;

            // svgf.comp:52-55
            float pcx = (float)cx - mv.x * pfd.display_size[0] + 0.5f;
            float pcy = (float)cy - mv.y * pfd.display_size[1] + 0.5f;
            float x = fractf(pcx), y = fractf(pcy);
            int ax = f2i_rz(pcx), ay = f2i_rz(pcy);

            const float bw[4] = {(1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y};
            const int off[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};

            float prev_shadow = 0.0f, prev_ao = 0.0f, sum = 0.0f;
            float psm[2] = {0.0f, 0.0f}, pam[2] = {0.0f, 0.0f};
            for (int i = 0; i < 4; ++i) {
                int sx = ax + off[i][0], sy = ay + off[i][1];
                if (is_valid_reprojection(pfd, prev_normals, W, sx, sy, current_object_id, current_normal)) {
                    vec4 h = load_rgba16f(history, W, sx, sy);
                    prev_shadow += bw[i] * h.x;
                    prev_ao += bw[i] * h.y;
                    vec4 m = load_rg16f(moments_in, W, sx, sy);   // .zw = (0,1) — SURVEY Q2
                    psm[0] += bw[i] * m.x; psm[1] += bw[i] * m.y;
                    pam[0] += bw[i] * m.z; pam[1] += bw[i] * m.w;
                    sum += bw[i];
                }
            }
            bool valid = sum > 1e-6f;
            // svgf.comp:81-97 — accumulators are NOT reset before the 3x3 retry
            if (!valid) {
                for (int yy = -1; yy <= 1; ++yy) {
                    for (int xx = -1; xx <= 1; ++xx) {
                        int sx = ax + xx, sy = ay + yy;
                        if (is_valid_reprojection(pfd, prev_normals, W, sx, sy, current_object_id, current_normal)) {
                            vec4 h = load_rgba16f(history, W, sx, sy);
                            vec4 m = load_rg16f(moments_in, W, sx, sy);
                            prev_shadow += h.x;
                            prev_ao += h.y;
                            psm[0] += m.x; psm[1] += m.y;
                            pam[0] += m.z; pam[1] += m.w;
                            sum += 1.0f;
                        }
                    }
                }
                valid = sum > 1e-6f;
            }

            float sm[2] = {current_shadow, current_shadow * current_shadow};
            float am[2] = {current_ao, current_ao * current_ao};
            vec4 out;
            if (valid) {
                const float alpha = 0.2f, moments_alpha = 0.2f;
                prev_shadow /= sum;
                psm[0] /= sum; psm[1] /= sum;
                prev_ao /= sum;
                pam[0] /= sum; pam[1] /= sum;
                sm[0] = mixf(psm[0], sm[0], moments_alpha); sm[1] = mixf(psm[1], sm[1], moments_alpha);
                am[0] = mixf(pam[0], am[0], moments_alpha); am[1] = mixf(pam[1], am[1], moments_alpha);
                float sv = gl_max(0.0f, sm[1] - sm[0] * sm[0]);
                float av = gl_max(0.0f, am[1] - am[0] * am[0]);
                out = vec4{mixf(prev_shadow, current_shadow, alpha), mixf(prev_ao, current_ao, alpha), sv, av};
            } else {
                float sv = gl_max(0.0f, sm[1] - sm[0] * sm[0]);
                float av = gl_max(0.0f, am[1] - am[0] * am[0]);
                out = vec4{current_shadow, current_ao, sv, av};
            }
            store_rgba16f(integrated_out, W, cx, cy, out);
            store_rg16f(moments_out, W, cx, cy, vec4{sm[0], sm[1], am[0], am[1]});   // RG16F keeps .xy only
        }
    }
}

// svgf_atrous_filter.comp:53-103
void vo_svgf_atrous(const PerFrameData *pfd_, int W, int H, int step,
                    const uint16_t *normals, const uint16_t *integ_in, uint16_t *integ_out) {
    const PerFrameData &pfd = *pfd_;
    static const float gauss[9] = {1.0f / 16, 1.0f / 8, 1.0f / 16, 1.0f / 8, 1.0f / 4, 1.0f / 8, 1.0f / 16, 1.0f / 8, 1.0f / 16};
    static const float atrous[25] = {
        1.0f / 256, 1.0f / 64, 3.0f / 128, 1.0f / 64, 1.0f / 256,
        1.0f / 64,  1.0f / 16, 3.0f / 32,  1.0f / 16, 1.0f / 64,
        3.0f / 128, 3.0f / 32, 9.0f / 64,  3.0f / 32, 3.0f / 128,
        1.0f / 64,  1.0f / 16, 3.0f / 32,  1.0f / 16, 1.0f / 64,
        1.0f / 256, 1.0f / 64, 3.0f / 128, 1.0f / 64, 1.0f / 256};
    const float dsx = pfd.display_size[0], dsy = pfd.display_size[1];
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < H; ++cy) {
        for (int cx = 0; cx < W; ++cx) {
            vec4 np4 = load_rgba16f(normals, W, cx, cy);
            vec3 normal_p = v3(np4.x, np4.y, np4.z);
            int object_id_p = f2i_rz(np4.w);
            vec4 ip = load_rgba16f(integ_in, W, cx, cy);

            // gauss_3x3_filter (:17-38): OOB taps skipped, no renormalisation
            float var_s = 0.0f, var_a = 0.0f;
            for (int y = -1; y <= 1; ++y)
                for (int x = -1; x <= 1; ++x) {
                    int sx = cx + x, sy = cy + y;
                    if (sx < 0 || (float)sx >= dsx || sy < 0 || (float)sy >= dsy) continue;
                    float w = gauss[3 * (y + 1) + (x + 1)];
                    vec4 q = load_rgba16f(integ_in, W, sx, sy);
                    var_s += w * q.z;
                    var_a += w * q.w;
                }

            float sum_wx = 1.0f, sum_wy = 1.0f;
            vec4 sum = ip;
            for (int y = -2; y <= 2; ++y)
                for (int x = -2; x <= 2; ++x) {
                    int sx = cx + x * step, sy = cy + y * step;
                    if (sx < 0 || (float)sx >= dsx || sy < 0 || (float)sy >= dsy || (x == 0 && y == 0)) continue;
                    vec4 iq = load_rgba16f(integ_in, W, sx, sy);
                    float kernel = atrous[5 * (y + 2) + (x + 2)];
                    vec4 nq4 = load_rgba16f(normals, W, sx, sy);
                    float wn = edge_stopping_normal(normal_p, v3(nq4.x, nq4.y, nq4.z));
                    float wid = (object_id_p == f2i_rz(nq4.w)) ? 1.0f : 0.0f;
                    float wk = kernel * wn * wid;
                    float wx = wk * edge_stopping_luminance(var_s, ip.x, iq.x);
                    float wy = wk * edge_stopping_luminance(var_a, ip.y, iq.y);
                    sum_wx += wx;
                    sum_wy += wy;
                    sum.x += wx * iq.x;
                    sum.y += wy * iq.y;
                    sum.z += (wx * wx) * iq.z;
                    sum.w += (wy * wy) * iq.w;
                }
            vec4 out = {sum.x / sum_wx, sum.y / sum_wy, sum.z / (sum_wx * sum_wx), sum.w / (sum_wy * sum_wy)};
            store_rgba16f(integ_out, W, cx, cy, out);
        }
    }
}

// Persistent SVGF state: the five storage images of hybrid_render_path.cpp:247-261, zero-initialised
// (documented deviation) plus the second moments buffer of the snapshot semantics (Q11).
struct vo_svgf_state {
    int W, H;
    std::vector<uint16_t> integrated[2];   // integrated_shadow_and_ao.x / .y
    std::vector<uint16_t> prev_normals;
    std::vector<uint16_t> history;
    std::vector<uint16_t> moments[2];
    int ping;        // which of integrated[] is currently ".x"
    int moments_cur; // which of moments[] holds the previous frame
};

vo_svgf_state *vo_svgf_state_create(int W, int H) {
    vo_svgf_state *s = new vo_svgf_state();
    s->W = W; s->H = H;
    size_t n = (size_t)W * H;
    s->integrated[0].assign(n * 4, 0);
    s->integrated[1].assign(n * 4, 0);
    s->prev_normals.assign(n * 4, 0);
    s->history.assign(n * 4, 0);
    s->moments[0].assign(n * 2, 0);
    s->moments[1].assign(n * 2, 0);
    s->ping = 0;
    s->moments_cur = 0;
    return s;
}
void vo_svgf_state_destroy(vo_svgf_state *s) { delete s; }

// Copy state images in/out so tests can seed identical temporal state on both sides.
// which: 0 = integrated.x, 1 = integrated.y, 2 = prev_normals, 3 = history, 4 = moments (previous frame)
uint16_t *vo_svgf_state_image(vo_svgf_state *s, int which) {
    switch (which) {
        case 0: return s->integrated[s->ping].data();
        case 1: return s->integrated[s->ping ^ 1].data();
        case 2: return s->prev_normals.data();
        case 3: return s->history.data();
        case 4: return s->moments[s->moments_cur].data();
    }
    return nullptr;
}

// hybrid_render_path.cpp:288-330. `iter_outputs` (optional, 5 * W*H*4 halfs) receives every à-trous
// iteration's output; `denoised` receives what the reference blits to "Denoised Raytraced Shadows and
// Ambient Occlusion" — the output of iteration index 3.
void vo_svgf_pass(vo_svgf_state *s, const PerFrameData *pfd, const uint16_t *normals, const uint16_t *motion,
                  const uint16_t *rt, uint16_t *denoised, uint16_t *iter_outputs, uint16_t *temporal_out) {
    const int W = s->W, H = s->H;
    const size_t n4 = (size_t)W * H * 4;
    int x = s->ping, y = s->ping ^ 1;   // integrated_shadow_and_ao.x / .y
    // :291-297 svgf.comp writes integrated[0] (= .x) and the moments image
    vo_svgf_temporal(pfd, W, H, normals, motion, rt, s->prev_normals.data(), s->history.data(),
                     s->moments[s->moments_cur].data(), s->integrated[x].data(), s->moments[s->moments_cur ^ 1].data());
    s->moments_cur ^= 1;
    if (temporal_out) std::memcpy(temporal_out, s->integrated[x].data(), n4 * 2);
    // :299-319
    for (int i = 0; i < 5; ++i) {
        int step = 1 << i;
        vo_svgf_atrous(pfd, W, H, step, normals, s->integrated[x].data(), s->integrated[y].data());
        if (iter_outputs) std::memcpy(iter_outputs + (size_t)i * n4, s->integrated[y].data(), n4 * 2);
        if (i == 0) s->history = s->integrated[y];          // BlitImageStorageToStorage(.y -> history)
        std::swap(x, y);                                    // :318
    }
    s->prev_normals.assign(normals, normals + n4);          // :321
    if (denoised) std::memcpy(denoised, s->integrated[y].data(), n4 * 2);   // :322-325 (.y after the swaps = it3 output)
    std::swap(x, y);                                        // :328
    s->ping = x;
}

// ---------------------------------------------------------------------------------------------
// SSAO
// ---------------------------------------------------------------------------------------------
namespace {
// texture() through the default sampler (resource_manager.cpp:58-69): LINEAR min/mag, REPEAT addressing,
// single mip; compute shaders sample LOD 0. Weights follow the Vulkan spec's float formula.
inline int wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
inline void bilinear_setup(float u, int n, int &i0, int &i1, float &a) {
    // NVIDIA texture units — the hardware the reference needs (RTX) — hold the filter coordinate in fixed point with 8 fractional bits
    // (VkPhysicalDeviceLimits::subTexelPrecisionBits = 8; CUDA C Programming Guide, "Linear Filtering": 9-bit fixed point with 8 bits of
    // fractional value, 1.0 exactly representable). Consequence: a lookup at a texel centre returns that texel even when
    // fl(fl((x + .5) / W) * W) is an ulp off x + .5 (51 of the 1920 columns at 1080p) — SURVEY Q17.
    float uu = std::floor((u * (float)n - 0.5f) * 256.0f + 0.5f) * 0.00390625f;
    float fl = std::floor(uu);
    a = uu - fl;
    // guard against non-finite coordinates (reference behaviour undefined): treat as texel 0
    int i = (fl == fl && std::fabs(fl) < 1e9f) ? (int)fl : 0;
    i0 = wrap(i, n);
    i1 = wrap(i + 1, n);
}
inline float sample_depth(const float *depth, int W, int H, float u, float v) {
    int x0, x1, y0, y1; float a, b;
    bilinear_setup(u, W, x0, x1, a);
    bilinear_setup(v, H, y0, y1, b);
    float t00 = depth[(size_t)y0 * W + x0], t10 = depth[(size_t)y0 * W + x1];
    float t01 = depth[(size_t)y1 * W + x0], t11 = depth[(size_t)y1 * W + x1];
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}
inline vec3 sample_normal(const uint16_t *normals, int W, int H, float u, float v) {
    int x0, x1, y0, y1; float a, b;
    bilinear_setup(u, W, x0, x1, a);
    bilinear_setup(v, H, y0, y1, b);
    vec4 t00 = load_rgba16f(normals, W, x0, y0), t10 = load_rgba16f(normals, W, x1, y0);
    vec4 t01 = load_rgba16f(normals, W, x0, y1), t11 = load_rgba16f(normals, W, x1, y1);
    float w00 = (1 - a) * (1 - b), w10 = a * (1 - b), w01 = (1 - a) * b, w11 = a * b;
    return {w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x,
            w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y,
            w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z};
}
}  // namespace

// ssao.comp:14-53. `radius` is a parameter (intent 0.75).
void vo_ssao(const PerFrameData *pfd_, int W, int H, float radius, const float *depth, const uint16_t *normals,
             uint16_t *ssao_raw) {
    const PerFrameData &pfd = *pfd_;
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < H; ++gy) {
        for (int gx = 0; gx < W; ++gx) {
            vec2 coords = {(float)gx * pfd.display_size_inverse[0], (float)gy * pfd.display_size_inverse[1]};
            float current_depth = sample_depth(depth, W, H, coords.x, coords.y);
            if (current_depth == 0.0f) {
                store_rgba16f(ssao_raw, W, gx, gy, vec4{0, 0, 0, 0});
                continue;
            }
            vec3 P = get_view_space_position(pfd, current_depth, coords);
            vec3 N = mul33_of44(pfd.camera_view, sample_normal(normals, W, H, coords.x, coords.y));
            float perspective_radius = radius / P.z;
            const int sigma = 1;
            const float beta = 1e-4f;
            uint32_t rng = seed_thread(((uint32_t)gy * (uint32_t)pfd.display_size[1] + (uint32_t)gx) * pfd.frame_index);
            const int num_samples = 16;
            float sum = 0.0f;
            for (int i = 0; i < num_samples; ++i) {
                float ang = random01(rng) * 2.0f * PI_F;
                float dist = random01(rng) * perspective_radius;
                vec2 offset = {std::cos(ang) * dist, std::sin(ang) * dist};
                vec2 sc = {coords.x + offset.x, coords.y + offset.y};
                vec3 V = get_view_space_position(pfd, sample_depth(depth, W, H, sc.x, sc.y), sc) - P;
                sum += gl_max(dot(V, N) - beta, 0.0f) / (dot(V, V) + 1e-4f);
            }
            float ao = gl_max(1.0f - ((2.0f * sigma) / (float)num_samples) * sum, 0.0f);
            store_rgba16f(ssao_raw, W, gx, gy, vec4{ao, ao, ao, ao});
        }
    }
}

// ssao_blur.comp:11-26
void vo_ssao_blur(const PerFrameData *pfd_, int W, int H, const uint16_t *ssao_raw, uint16_t *ssao) {
    const PerFrameData &pfd = *pfd_;
    const float dsx = pfd.display_size[0], dsy = pfd.display_size[1];
#pragma omp parallel for schedule(static)
    for (int cy = 0; cy < H; ++cy) {
        for (int cx = 0; cx < W; ++cx) {
            float ao = 0.0f;
            for (int y = -6; y <= 6; ++y)
                for (int x = -6; x <= 6; ++x) {
                    int sx = cx + x, sy = cy + y;
                    if (sx < 0 || (float)sx >= dsx || sy < 0 || (float)sy >= dsy) continue;
                    ao += h2f(ssao_raw[((size_t)sy * W + sx) * 4]);
                }
            float o = ao / (13.0f * 13.0f);
            store_rgba16f(ssao, W, cx, cy, vec4{o, o, o, o});
        }
    }
}

// ---------------------------------------------------------------------------------------------
// KAT helpers (tests/golden): RNG and sampling functions of common.glsl
// ---------------------------------------------------------------------------------------------
uint32_t vo_seed_thread(uint32_t seed) { return seed_thread(seed); }
float vo_random01(uint32_t *state) { return random01(*state); }
void vo_uniform_sample_cone(float u0, float u1, float cos_theta_max, float *out3) {
    vec3 r = uniform_sample_cone(vec2{u0, u1}, cos_theta_max);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void vo_cosine_hemisphere(float u0, float u1, float *out3) {
    vec3 r = uniform_sample_cosine_weighted_hemisphere(vec2{u0, u1});
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void vo_onb(const float *n, float *out9) {
    mat3 M = onb_from_unit_vector(v3(n[0], n[1], n[2]));
    out9[0] = M.c0.x; out9[1] = M.c0.y; out9[2] = M.c0.z;
    out9[3] = M.c1.x; out9[4] = M.c1.y; out9[5] = M.c1.z;
    out9[6] = M.c2.x; out9[7] = M.c2.y; out9[8] = M.c2.z;
}
uint16_t vo_f2h(float f) { return f2h(f); }
float vo_h2f(uint16_t h) { return h2f(h); }
int vo_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(PerFrameData);
        case 1: return (int)sizeof(Vertex);
        case 2: return (int)sizeof(Material);
        case 3: return (int)sizeof(Primitive);
        case 4: return (int)sizeof(DirectionalLight);
    }
    return -1;
}

}  // extern "C"
