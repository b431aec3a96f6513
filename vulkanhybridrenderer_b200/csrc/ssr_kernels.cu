// ssr_kernels.cu — screen-space reflections of the hybrid path (sm_100a).
//
//   ssr_kernel <- /root/reference/data/shaders/hybrid_render_path/ssr.comp:61-137 (march + binary search) and
//                 :29-59 (compute_lighting), dispatched by the "SSR Pass" node
//                 (/root/reference/src/render_paths/hybrid_render_path.cpp:202-243) with SSRPushConstants
//                 (/root/reference/src/rendering_backend/glsl_common.h:41-46)
//
// One thread per pixel in 8x4-pixel warp tiles (neighbouring pixels march almost the same ray, so the warp leaves the
// loop together and the depth taps of a step share cache lines). Every step re-projects the ray point, takes one
// bilinear depth tap through the reference's default sampler (LINEAR / REPEAT, evaluated in software with the Vulkan
// float weights, like ssao_kernels.cu) and compares two camera distances against 0.3 / thickness — threshold tests on
// nearly equal numbers, so all of it is written with explicitly rounded operations in the oracle's order: one ulp of
// difference would pick another march step and change the whole pixel. The matrix product camera_proj * camera_view
// of world_space_to_uv (ssr.comp:22-26, GLSL evaluates it left to right) is formed once per dispatch.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "vhr_internal.h"


#ifndef VHR_SSR_SKIP_COOLDOWN
#define VHR_SSR_SKIP_COOLDOWN 7
#endif

namespace vhr {

struct SsrParams {
    int W, H;
    int x_end, y_begin, y_end;
    float step_size, thickness;
    float Wf, Hf;              // (float)W, (float)H
    int n_steps;               // int(ray_distance / step_size), ssr.comp:89
    int bsearch_steps;
    const uint32_t *albedo;    // binding 0 (BGRA8, sampled)
    const uint2 *normals;      // binding 1
    const uint2 *motion;       // binding 2 (.zw = metallic, roughness)
    const float *depth;        // binding 3
    const float4 *quads;       // the 2 x 2 bilinear footprint of every depth texel as one 16-byte word (build_depth_quads, ssao_kernels.cu)
    const float2 *tiles;       // (min, max) depth over texels [8 i - 2, 8 i + 10) x [8 j - 2, 8 j + 10) (REPEAT-wrapped) of tile (i, j), or nullptr: no skipping
    int tiles_w, tiles_h;
    float margin0;             // 1e-5 + 4e-6 |camera position|_1
    float pi14sq;
    float pi0, pi5, pi14, pi11, pi15;   // camera_proj_inverse entries m00, m11, m23, m32, m33 (column-major indices 0, 5, 14, 11, 15)
    uint2 *out;                // binding 4 (RGBA16F)
    float pv[16];              // camera_proj * camera_view, column-major
};

namespace {

struct Taps {
    size_t i00, i10, i01, i11;
    float a, b;
};
__device__ __forceinline__ Taps taps_for(const SsrParams &p, float u, float v) {
    int x0, x1, y0, y1;
    Taps t;
    bilinear_setup(u, p.W, p.Wf, x0, x1, t.a);
    bilinear_setup(v, p.H, p.Hf, y0, y1, t.b);
    t.i00 = (size_t)y0 * p.W + x0; t.i10 = (size_t)y0 * p.W + x1;
    t.i01 = (size_t)y1 * p.W + x0; t.i11 = (size_t)y1 * p.W + x1;
    return t;
}
// The depth tap of a probe: the four texels come as ONE 16-byte load from the quad image (the REPEAT wrap of the right / bottom neighbour
// is built into it, so only the first texel's index is formed) — same values as the four separate taps, a quarter of the L1 work.
__device__ __forceinline__ float sample_depth(const SsrParams &p, float u, float v) {
    int x0, x1, y0, y1;
    float a, b;
    bilinear_setup(u, p.W, p.Wf, x0, x1, a);
    bilinear_setup(v, p.H, p.Hf, y0, y1, b);
    const float4 q = __ldg(p.quads + (y0 * p.W + x0));
    return bilerp_rn(a, b, q.x, q.y, q.z, q.w);
}
__device__ __forceinline__ float3 sample_xyz16(const uint2 *img, const Taps &t) {
    const float4 t00 = unpack_rgba16f(__ldg(&img[t.i00])), t10 = unpack_rgba16f(__ldg(&img[t.i10]));
    const float4 t01 = unpack_rgba16f(__ldg(&img[t.i01])), t11 = unpack_rgba16f(__ldg(&img[t.i11]));
    return make_float3(bilerp_rn(t.a, t.b, t00.x, t10.x, t01.x, t11.x), bilerp_rn(t.a, t.b, t00.y, t10.y, t01.y, t11.y),
                       bilerp_rn(t.a, t.b, t00.z, t10.z, t01.z, t11.z));
}
__device__ __forceinline__ float2 sample_zw16(const uint2 *img, const Taps &t) {
    const float2 t00 = unpack_rg16f(__ldg(&img[t.i00].y)), t10 = unpack_rg16f(__ldg(&img[t.i10].y));
    const float2 t01 = unpack_rg16f(__ldg(&img[t.i01].y)), t11 = unpack_rg16f(__ldg(&img[t.i11].y));
    return make_float2(bilerp_rn(t.a, t.b, t00.x, t10.x, t01.x, t11.x), bilerp_rn(t.a, t.b, t00.y, t10.y, t01.y, t11.y));
}
// B8G8R8A8_UNORM: byte 0 = B; UNORM -> float is c / 255
__device__ __forceinline__ float3 bgra8_rgb(uint32_t c) {
    return make_float3(__fdiv_rn((float)((c >> 16) & 0xffu), 255.0f), __fdiv_rn((float)((c >> 8) & 0xffu), 255.0f),
                       __fdiv_rn((float)(c & 0xffu), 255.0f));
}
__device__ __forceinline__ float3 sample_albedo(const uint32_t *img, const Taps &t) {
    const float3 t00 = bgra8_rgb(__ldg(&img[t.i00])), t10 = bgra8_rgb(__ldg(&img[t.i10]));
    const float3 t01 = bgra8_rgb(__ldg(&img[t.i01])), t11 = bgra8_rgb(__ldg(&img[t.i11]));
    return make_float3(bilerp_rn(t.a, t.b, t00.x, t10.x, t01.x, t11.x), bilerp_rn(t.a, t.b, t00.y, t10.y, t01.y, t11.y),
                       bilerp_rn(t.a, t.b, t00.z, t10.z, t01.z, t11.z));
}
__device__ __forceinline__ float distance_rn(float3 a, float3 b) {
    const float3 d = make_float3(sub_rn(a.x, b.x), sub_rn(a.y, b.y), sub_rn(a.z, b.z));
    return sqrt_exact(dot3_rn(d, d));
}

// (min, max) of the depth image over every 8 x 8-texel tile plus a 2-texel apron, REPEAT-wrapped like the sampler.
__global__ void __launch_bounds__(128) depth_tiles_kernel(const float *__restrict__ depth, float2 *__restrict__ tiles, int W, int H, int tw, int th) {
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= tw * th) return;
    const int ti = t % tw, tj = t / tw;
    float mn = 3.0e38f, mx = -3.0e38f;
    for (int k = lane; k < 144; k += 32) {
        int x = ti * 8 - 2 + k % 12, y = tj * 8 - 2 + k / 12;
        x = x < 0 ? x + W : (x >= W ? x - W : x);
        y = y < 0 ? y + H : (y >= H ? y - H : y);
        if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
            const float d = __ldg(depth + (size_t)y * W + x);
            mn = fminf(mn, d); mx = fmaxf(mx, d);
            if (!(d == d)) { mn = -3.0e38f; mx = 3.0e38f; }      // a NaN texel: the tile decides nothing
        } else { mn = -3.0e38f; mx = 3.0e38f; }                  // image smaller than the apron
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) tiles[t] = make_float2(mn, mx);
}

// Can the march step at `offset` be decided WITHOUT evaluating it? The window test compares delta = |cam - rp| - |cam - sp|, where sp is
// the surface point under the ray point's screen position. For a perspective camera |cam - sp| = L(uv) / w(depth) with L = |(m00 x, m11 y,
// m23)| and w = m32 depth + m33 (the view-space form of the oracle's unprojection: the host has checked that camera_viewproj_inverse is
// camera_view_inverse * camera_proj_inverse with a rigid view and an inverse-perspective projection, m32 > 0, m33 >= 0), decreasing in
// depth. The depth the oracle samples is a convex combination of four texels of the tile the (approximate) uv falls in — the tile's
// 2-texel apron absorbs the difference between this approximate uv and the oracle's — so it lies in the tile's [min, max], and delta in
// [d1 - L / w(min), d1 - L / w(max)]. If that interval lies below 0.3 or above `thickness` by a margin of 1e-4 (d1 + d2) + 1e-5 + 4e-6 |cam|
// (a hundred times the rounding of both evaluations) the step is outside the window whatever the exact numbers are. ~45 instructions against
// ~200 for the probe; most steps of most rays are in open space in front of the geometry or well behind it.
__device__ __forceinline__ bool step_is_outside_window(const SsrParams &p, float3 P, float3 dir, float3 cam, float offset) {
    const float rx = fmaf(dir.x, offset, P.x), ry = fmaf(dir.y, offset, P.y), rz = fmaf(dir.z, offset, P.z);
    const float *m = p.pv;
    const float cx = fmaf(m[0], rx, fmaf(m[4], ry, fmaf(m[8], rz, m[12])));
    const float cy = fmaf(m[1], rx, fmaf(m[5], ry, fmaf(m[9], rz, m[13])));
    const float cw = fmaf(m[3], rx, fmaf(m[7], ry, fmaf(m[11], rz, m[15])));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(cw));
    const float x = cx * r, y = cy * r;                              // NDC = 2 uv - 1
    // texel coordinate u W = (x + 1) W / 2; one period of REPEAT, anything further out is left to the probe
    int tx = __float2int_rd(fmaf(x, 0.5f * p.Wf, 0.5f * p.Wf)), ty = __float2int_rd(fmaf(y, 0.5f * p.Hf, 0.5f * p.Hf));
    tx += tx < 0 ? p.W : 0; tx -= tx >= p.W ? p.W : 0;
    ty += ty < 0 ? p.H : 0; ty -= ty >= p.H ? p.H : 0;
    if (!((unsigned)tx < (unsigned)p.W && (unsigned)ty < (unsigned)p.H)) return false;      // also NaN / inf coordinates
    const float2 mm = __ldg(p.tiles + (ty >> 3) * p.tiles_w + (tx >> 3));
    if (!(mm.x >= 0.0f)) return false;
    const float lx = p.pi0 * x, ly = p.pi5 * y;
    const float Lq = fmaf(lx, lx, fmaf(ly, ly, p.pi14sq));
    float Li;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(Li) : "f"(Lq));
    const float L = Lq * Li;                                         // Lq >= m23^2 > 0
    const float w_far = fmaf(p.pi11, mm.x, p.pi15), w_near = fmaf(p.pi11, mm.y, p.pi15);
    float rf, rn;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(w_far));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"(w_near));
    const float d2_far = L * rf, d2_near = L * rn;                   // w_far = 0 (sky in the tile): +inf
    const float fx = cam.x - rx, fy = cam.y - ry, fz = cam.z - rz;
    const float d1q = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
    float d1i;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d1i) : "f"(d1q));
    const float d1 = d1q * d1i;                                      // d1q = 0 (the ray point is the camera): NaN, nothing is decided
    // margins: 1e-4 of the distances involved + 4e-6 of the camera's distance from the origin (the oracle subtracts world-space points)
    const float m_near = fmaf(1e-4f, d1 + fminf(d2_near, 1e30f), p.margin0), m_far = fmaf(1e-4f, d1 + fminf(d2_far, 1e30f), p.margin0);
    // in front of (or not far enough behind) the nearest surface of the tile: delta <= d1 - d2_near < 0.3 (an all-sky tile: d2_near = inf)
    if (d1 + m_near - 0.3f < d2_near) return true;
    // behind the farthest surface of the tile by more than the thickness (never with sky in the tile: d2_far = inf)
    if (d1 - p.thickness - m_far > d2_far) return true;
    return false;
}

// The tail of a probe in the oracle's operations: three IEEE quotients, two correctly rounded square roots (ssr.comp:92-99). Out of line:
// in_window() below only comes here when its cheap estimate of delta is within the error bound of a threshold.
__device__ __noinline__ float delta_exact_tail(float3 cam, float3 rp, float4 q) {
    const ExactDivisor dq = exact_divisor(q.w);
    const float3 sp = make_float3(div_exact(q.x, dq), div_exact(q.y, dq), div_exact(q.z, dq));
    return sub_rn(distance_rn(cam, rp), distance_rn(cam, sp));
}

// One probe of the march / the binary search (ssr.comp:90-99, 115-123): is delta_distance at `offset` along the ray inside the window
// (0.3, thickness)? — and the uv the ray point projects to. EXACT decisions at a fraction of the exact arithmetic:
//   * everything that selects the depth taps is computed in the oracle's operations and order: the ray point, clip.x / .y / .w (clip.z is
//     never formed), the two IEEE quotients (div_exact), su / sv, the fixed-point filter coordinate, the rounded bilinear interpolation
//     and the 4 x 4 unprojection product q — these are the oracle's bits;
//   * the rest of the oracle's probe (three quotients q.xyz / q.w, two distances with correctly rounded square roots, their difference)
//     only feeds two comparisons, so it is ESTIMATED: one MUFU reciprocal, fused dot products, rsqrt — each within a few ulps of the
//     rounded operation it replaces: |estimate - oracle's delta| stays below 8 * 2^-24 * (d1 + d2 + |sp.x| + |sp.y| + |sp.z|) (sum of the
//     per-operation bounds; 0.74 of it is the largest deviation on 1.6e7 random probes with worst-case MUFU errors,
//     tests/test_ssr_cpu.py), and the band used is four times that. When the estimate is further than the band from both thresholds the
//     comparisons are decided; otherwise (about one probe in 10^4, and NaN estimates) the oracle's tail is evaluated;
//   * q.w = 0 exactly — a tap on the sky, where most rays end — makes the oracle's quotients +-inf or NaN, its delta -inf or NaN and both
//     comparisons false: decided without arithmetic (q.w is the oracle's own value).
__device__ __forceinline__ bool in_window(const SsrParams &p, const PerFrameData &pfd, float3 P, float3 dir, float3 cam, float offset,
                                          float &su, float &sv) {
    const float3 rp = make_float3(add_rn(P.x, mul_rn(dir.x, offset)), add_rn(P.y, mul_rn(dir.y, offset)), add_rn(P.z, mul_rn(dir.z, offset)));
    const float *m = p.pv;
    const float cx = dot4_rn(m[0], rp.x, m[4], rp.y, m[8], rp.z, m[12], 1.0f);
    const float cy = dot4_rn(m[1], rp.x, m[5], rp.y, m[9], rp.z, m[13], 1.0f);
    const float cw = dot4_rn(m[3], rp.x, m[7], rp.y, m[11], rp.z, m[15], 1.0f);
    const ExactDivisor dw = exact_divisor(cw);
    su = add_rn(mul_rn(div_exact(cx, dw), 0.5f), 0.5f);
    sv = add_rn(mul_rn(div_exact(cy, dw), 0.5f), 0.5f);
    const float4 q = mul44_rn(pfd.camera_viewproj_inverse, make_float4(sub_rn(mul_rn(su, 2.0f), 1.0f), sub_rn(mul_rn(sv, 2.0f), 1.0f), sample_depth(p, su, sv), 1.0f));
    if (q.w == 0.0f) return false;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q.w));
    const float sx = q.x * r, sy = q.y * r, sz = q.z * r;
    const float ex = cam.x - sx, ey = cam.y - sy, ez = cam.z - sz;
    const float fx = sub_rn(cam.x, rp.x), fy = sub_rn(cam.y, rp.y), fz = sub_rn(cam.z, rp.z);
    const float d2q = fmaf(ex, ex, fmaf(ey, ey, ez * ez)), d1q = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
    float i2, i1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i2) : "f"(d2q));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(d1q));
    const float d2 = d2q * i2, d1 = d1q * i1;
    const float delta = d1 - d2;
    const float bound = (d1 + d2 + fabsf(sx) + fabsf(sy) + fabsf(sz)) * 1.9073486e-6f;      // 32 * 2^-24
    const float lo = delta - 0.3f, hi = p.thickness - delta;
    if (lo > bound && hi > bound) return true;
    if (lo < -bound || hi < -bound) return false;
    const float exact = delta_exact_tail(cam, rp, q);           // undecided (or NaN: 0 * inf, a zero distance)
    return exact > 0.3f && exact < p.thickness;
}

}  // namespace

// ssr.comp:62-86 for one pixel: the fragment's world position and the reflected direction. False (nothing to march): a sky pixel (depth
// 0 -> w = 0) has a non-finite P — every distance along its ray is inf or NaN, the window test can never pass — the output is zero.
__device__ __forceinline__ bool pixel_setup(const SsrParams &p, const PerFrameData &pfd, int gx, int gy, float3 cam, float3 &P, float3 &dir) {
    const float cu = mul_rn((float)gx, pfd.display_size_inverse[0]);                 // texel corner (ssr.comp:68)
    const float cv = mul_rn((float)gy, pfd.display_size_inverse[1]);
    const Taps t0 = taps_for(p, cu, cv);
    const float fragment_depth = bilerp_rn(t0.a, t0.b, __ldg(&p.depth[t0.i00]), __ldg(&p.depth[t0.i10]), __ldg(&p.depth[t0.i01]), __ldg(&p.depth[t0.i11]));
    P = unproject_rn(pfd.camera_viewproj_inverse, fragment_depth, cu, cv);
    const float3 N = sample_xyz16(p.normals, t0);
    const float3 I = normalize_rn(make_float3(sub_rn(P.x, cam.x), sub_rn(P.y, cam.y), sub_rn(P.z, cam.z)));
    const float k2 = mul_rn(2.0f, dot3_rn(N, I));
    dir = normalize_rn(make_float3(sub_rn(I.x, mul_rn(N.x, k2)), sub_rn(I.y, mul_rn(N.y, k2)), sub_rn(I.z, mul_rn(N.z, k2))));
    return fabsf(P.x) <= 3.0e38f && fabsf(P.y) <= 3.0e38f && fabsf(P.z) <= 3.0e38f;
}

// compute_lighting(final_uv), ssr.comp:29-59
__device__ __forceinline__ uint2 lighting(const SsrParams &p, const PerFrameData &pfd, float3 cam, float fu, float fv) {
    const Taps t = taps_for(p, fu, fv);
    const float3 albedo = sample_albedo(p.albedo, t);
    const float d = bilerp_rn(t.a, t.b, __ldg(&p.depth[t.i00]), __ldg(&p.depth[t.i10]), __ldg(&p.depth[t.i01]), __ldg(&p.depth[t.i11]));
    const float3 position = unproject_rn(pfd.camera_viewproj_inverse, d, fu, fv);
    const float2 mr = sample_zw16(p.motion, t);
    const float3 V = normalize_rn(make_float3(sub_rn(cam.x, position.x), sub_rn(cam.y, position.y), sub_rn(cam.z, position.z)));
    const float3 L = make_float3(-pfd.directional_light.direction[0], -pfd.directional_light.direction[1], -pfd.directional_light.direction[2]);
    const float3 Nh = sample_xyz16(p.normals, t);
    const float3 H = normalize_rn(make_float3(add_rn(L.x, V.x), add_rn(L.y, V.y), add_rn(L.z, V.z)));
    const float3 c = shade_direct_rn(albedo, mr.x, mr.y, Nh, V, L, H, pfd.directional_light.intensity, pfd.directional_light.color);
    return pack_rgba16f(make_float4(c.x, c.y, c.z, 1.0f));
}

// Variant 0 (default): one thread per pixel in 8x4-pixel warp tiles, the shader's control flow as written. A pixel that finds its hit
// after 20 steps idles while a neighbour walks all 250: 20 of 32 lanes are active on average (ncu) — and still the faster kernel, because
// the 32 lanes probe the SAME step of nearly parallel rays, so the four depth taps of a probe fall into a few cache lines.
__global__ void __launch_bounds__(128) ssr_kernel(const __grid_constant__ SsrParams p, const __grid_constant__ PerFrameData pfd) {
    // 16x8 block of four 8x4 warp tiles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = p.y_begin + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (gx >= p.x_end || gy >= p.y_end) return;
    const size_t pix = (size_t)gy * p.W + gx;
    const uint2 zero = make_uint2(0u, 0u);                                           // ssr.comp:62-66
    const float3 cam = make_float3(pfd.camera_view_inverse[12], pfd.camera_view_inverse[13], pfd.camera_view_inverse[14]);
    float3 P, dir;
    const bool finite_p = pixel_setup(p, pfd, gx, gy, cam, P, dir);
    bool found = false;
    float prev_step = 0.0f, final_step = 0.0f;
    // (two and four probes per round, examined in order afterwards, were measured: 4.89 / 5.66 ms against 4.84 — the kernel is bound by
    // instruction issue, not by latency)
    // The skip test only pays when the whole warp can skip the step (a probe issued for one lane costs as much as for 32): after a round
    // in which some lane needed its probe the warp goes `VHR_SSR_SKIP_COOLDOWN` rounds without testing (a warp-uniform counter; which
    // steps are tested changes nothing in the result, only how many probes are evaluated).
    int cooldown = 0;
    for (int i = 0; finite_p && i < p.n_steps; ++i) {                                            // ssr.comp:89-108
        const float offset = mul_rn(p.step_size, (float)i);
        float pu, pv;
        bool skip = false;
        if (p.tiles) {
            if (cooldown == 0) {
                skip = step_is_outside_window(p, P, dir, cam, offset);
                if (!__all_sync(__activemask(), skip)) cooldown = VHR_SSR_SKIP_COOLDOWN;
            } else --cooldown;
        }
        if (!skip && in_window(p, pfd, P, dir, cam, offset, pu, pv)) {
            final_step = offset;
            found = true;
            break;
        }
        prev_step = offset;
    }
    if (!found) {                                                                    // ssr.comp:110-112
        p.out[pix] = zero;
        return;
    }
    float mid_step = mul_rn(add_rn(prev_step, final_step), 0.5f);                    // ssr.comp:115-135
    float fu = 0.0f, fv = 0.0f;
    for (int i = 0; i < p.bsearch_steps; ++i) {
        if (in_window(p, pfd, P, dir, cam, mid_step, fu, fv)) {
            mid_step = mul_rn(add_rn(prev_step, mid_step), 0.5f);
        } else {
            const float tmp = mid_step;
            mid_step = add_rn(mid_step, sub_rn(mid_step, prev_step));
            prev_step = tmp;
        }
    }
    p.out[pix] = lighting(p, pfd, cam, fu, fv);
}

// Variant 1 (VHR_SSR_VARIANT=1, study): the same per-pixel computation with LANE REFILL. A warp owns a 32 x 8-pixel region (eight 8x4 tiles, walked in
// order) and keeps a pixel in every lane: a lane whose pixel is finished takes the next one of the region instead of idling until the
// slowest pixel of its tile has walked its 250 steps. Every lane runs the same loop body — one probe per round, of the march or of the
// binary search, which differ in how the offset is formed and what the answer updates (a per-lane state, ssr.comp:89-135 unrolled into
// MARCH / BSEARCH) — so unlike ray traversal there are no divergent phases to serialise; the two heavy one-off pieces (pixel set-up,
// ~300 instructions; lighting, ~250) run for the lanes that wait for them once GATHER have gathered or nothing else is left to do.
// Bit-identical images, and 6.7-7.5 ms against 4.1 ms for every region size and gather threshold tried (2 / 4 / 8 tiles, 4 / 8 lanes,
// profiles/r02/ssr_variants.log): lanes at different march steps probe different places, the tap loads of a warp spread over up to 32
// cache lines each, and the L1 pays more than the idle lanes cost.
template <int TX, int TY, int GATHER>
__global__ void __launch_bounds__(128) ssr_refill_kernel(const __grid_constant__ SsrParams p, const __grid_constant__ PerFrameData pfd) {
    enum : int { NEW = 0, MARCH = 1, BSEARCH = 2, LIGHT = 3, DONE = 4 };
    constexpr int REGION = 32 * TX * TY;          // TX x TY tiles of 8 x 4 pixels per warp
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rx0 = (blockIdx.x * 2 + (warp & 1)) * (8 * TX), ry0 = p.y_begin + (blockIdx.y * 2 + (warp >> 1)) * (4 * TY);
    if (rx0 >= p.x_end || ry0 >= p.y_end) return;                                    // warp-uniform
    const float3 cam = make_float3(pfd.camera_view_inverse[12], pfd.camera_view_inverse[13], pfd.camera_view_inverse[14]);
    int k = lane, next = 32, phase = NEW;
    int gx = 0, gy = 0, i = 0;
    float3 P = make_float3(0.0f, 0.0f, 0.0f), dir = P;
    float prev_step = 0.0f, mid_step = 0.0f, fu = 0.0f, fv = 0.0f;
    while (true) {
        unsigned m_run = __ballot_sync(FULL, phase == MARCH || phase == BSEARCH);
        const unsigned m_new = __ballot_sync(FULL, phase == NEW);
        if (m_new && (__popc(m_new) >= GATHER || !m_run)) {
            if (phase == NEW) {
                const int tile = k >> 5, l = k & 31;
                gx = rx0 + (tile % TX) * 8 + (l & 7);
                gy = ry0 + (tile / TX) * 4 + (l >> 3);
                phase = DONE;
                if (gx < p.x_end && gy < p.y_end) {
                    if (pixel_setup(p, pfd, gx, gy, cam, P, dir) && p.n_steps > 0) {
                        phase = MARCH; i = 0; prev_step = 0.0f;
                    } else {
                        p.out[(size_t)gy * p.W + gx] = make_uint2(0u, 0u);               // ssr.comp:62-66, 110-112
                    }
                }
            }
            m_run = __ballot_sync(FULL, phase == MARCH || phase == BSEARCH);
        }
        const unsigned m_light = __ballot_sync(FULL, phase == LIGHT);
        if (m_light && (__popc(m_light) >= GATHER || !m_run)) {
            if (phase == LIGHT) {
                p.out[(size_t)gy * p.W + gx] = lighting(p, pfd, cam, fu, fv);
                phase = DONE;
            }
        }
        const unsigned m_done = __ballot_sync(FULL, phase == DONE);
        if (m_done && next < REGION) {
            const int kk = next + __popc(m_done & ((1u << lane) - 1u));
            if (phase == DONE && kk < REGION) { k = kk; phase = NEW; }
            next += __popc(m_done);
        }
        if (__all_sync(FULL, phase == DONE)) break;
        if (phase == MARCH || phase == BSEARCH) {
            const float offset = phase == MARCH ? mul_rn(p.step_size, (float)i) : mid_step;
            float pu, pv;
            const bool hit = in_window(p, pfd, P, dir, cam, offset, pu, pv);
            if (phase == MARCH) {                                                     // ssr.comp:89-112
                if (hit) {
                    mid_step = mul_rn(add_rn(prev_step, offset), 0.5f);              // ssr.comp:115
                    fu = 0.0f; fv = 0.0f; i = 0;
                    phase = p.bsearch_steps > 0 ? BSEARCH : LIGHT;
                } else {
                    prev_step = offset;
                    if (++i >= p.n_steps) {
                        p.out[(size_t)gy * p.W + gx] = make_uint2(0u, 0u);
                        phase = DONE;
                    }
                }
            } else {                                                                 // ssr.comp:116-135
                fu = pu; fv = pv;
                if (hit) {
                    mid_step = mul_rn(add_rn(prev_step, mid_step), 0.5f);
                } else {
                    const float tmp = mid_step;
                    mid_step = add_rn(mid_step, sub_rn(mid_step, prev_step));
                    prev_step = tmp;
                }
                if (++i >= p.bsearch_steps) phase = LIGHT;
            }
        }
    }
}

// Does step_is_outside_window's bound hold for these matrices? camera_proj_inverse with the sparsity of an inverse perspective projection
// (m32 > 0, m33 >= 0: w = m32 depth + m33 grows with depth and is not negative), camera_view_inverse rigid, camera_viewproj_inverse their
// product (to 1e-4 of the largest entry) — then the distance from the camera to the oracle's unprojected point equals the length of the
// view-space point. Anything else (a hand-made test matrix, an orthographic camera): no skipping, every step is probed.
static bool skip_bound_applies(const PerFrameData &pfd) {
    const float *pi = pfd.camera_proj_inverse, *vi = pfd.camera_view_inverse, *vpi = pfd.camera_viewproj_inverse;
    for (int i = 0; i < 16; ++i)
        if (i != 0 && i != 5 && i != 11 && i != 14 && i != 15 && pi[i] != 0.0f) return false;
    if (!(pi[11] > 0.0f) || !(pi[15] >= 0.0f) || !(pi[0] != 0.0f) || !(pi[5] != 0.0f) || !(pi[14] != 0.0f)) return false;
    if (vi[3] != 0.0f || vi[7] != 0.0f || vi[11] != 0.0f || vi[15] != 1.0f) return false;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double d = 0.0;
            for (int k = 0; k < 3; ++k) d += (double)vi[a * 4 + k] * vi[b * 4 + k];      // columns of the rotation are orthonormal
            if (std::fabs(d - (a == b ? 1.0 : 0.0)) > 1e-4) return false;
        }
    double big = 0.0, prod[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += (double)vi[k * 4 + r] * pi[c * 4 + k];
            prod[c * 4 + r] = acc;
            big = std::max(big, std::fabs(acc));
        }
    for (int i = 0; i < 16; ++i)
        if (!(std::fabs(prod[i] - (double)vpi[i]) <= 1e-4 * big)) return false;
    return true;
}

int launch_ssr(vhr_context *ctx, uint32_t xg, uint32_t yg, const SSRPushConstants &pc) {
    // descriptor set 3 of the "SSR Pass" (hybrid_render_path.cpp:211-219): 0 albedo, 1 normals, 2 motion, 3 depth, 4 output
    if (ctx->n_bound < 5) return fail(VHR_ERR_STATE, "ssr.comp: pass images not bound (need bindings 0..4)");
    Image **b = ctx->bound;
    for (int i = 0; i < 5; ++i)
        if (!b[i]) return fail(VHR_ERR_STATE, "ssr.comp: unbound image");
    const int want[5] = {VHR_FORMAT_B8G8R8A8_UNORM, VHR_FORMAT_R16G16B16A16_SFLOAT, VHR_FORMAT_R16G16B16A16_SFLOAT, VHR_FORMAT_D32_SFLOAT,
                         VHR_FORMAT_R16G16B16A16_SFLOAT};
    for (int i = 0; i < 5; ++i) {
        if (b[i]->format != want[i]) return fail(VHR_ERR_INVALID, "ssr.comp: binding %d has format %d, expected %d", i, b[i]->format, want[i]);
        if (b[i]->width != b[4]->width || b[i]->height != b[4]->height) return fail(VHR_ERR_INVALID, "ssr.comp: image sizes differ");
    }
    if (ctx->pfd.display_size[0] != (float)b[4]->width || ctx->pfd.display_size[1] != (float)b[4]->height)
        return fail(VHR_ERR_INVALID, "ssr.comp: PerFrameData.display_size does not match the images");
    // int(pc.ray_distance / pc.step_size), ssr.comp:89; a non-finite or huge quotient (step_size = 0) has no defined value in GLSL
    const float q = pc.ray_distance / pc.step_size;
    if (!(q == q) || q > 1048576.0f) return fail(VHR_ERR_INVALID, "ssr.comp: ray_distance / step_size = %g", (double)q);
    if (pc.bsearch_steps < 0 || pc.bsearch_steps > 4096) return fail(VHR_ERR_INVALID, "ssr.comp: bsearch_steps = %d", pc.bsearch_steps);
    if (int rc = make_writable(ctx, b[4], covers_image(ctx, b[4], (uint64_t)xg * 8, (uint64_t)yg * 8))) return rc;
    SsrParams p;
    p.W = (int)b[4]->width; p.H = (int)b[4]->height;
    p.x_end = (int)std::min<uint64_t>(b[4]->width, (uint64_t)xg * 8);
    const int y_cov = (int)std::min<uint64_t>(b[4]->height, (uint64_t)yg * 8);
    p.y_begin = std::max(0, ctx->opt.row_begin);
    p.y_end = ctx->opt.row_end < 0 ? y_cov : std::min(y_cov, ctx->opt.row_end);
    if (p.x_end <= 0 || p.y_end <= p.y_begin) return VHR_OK;
    p.step_size = pc.step_size; p.thickness = pc.thickness;
    p.Wf = (float)p.W; p.Hf = (float)p.H;
    p.n_steps = q < 0.0f ? 0 : (int)q;
    p.bsearch_steps = pc.bsearch_steps;
    p.albedo = (const uint32_t *)b[0]->ptr; p.normals = (const uint2 *)b[1]->ptr; p.motion = (const uint2 *)b[2]->ptr;
    p.depth = (const float *)b[3]->ptr; p.out = (uint2 *)b[4]->ptr;
    if (int rc = build_depth_quads(ctx, p.depth, p.W, p.H, &p.quads)) return rc;
    // conservative step skipping (step_is_outside_window): only for a camera whose matrices have the structure the bound relies on
    p.tiles = nullptr; p.tiles_w = (p.W + 7) / 8; p.tiles_h = (p.H + 7) / 8;
    // Off by default (VHR_SSR_SKIP=1 turns it on): 79 % of the lane-steps of the bench frame can be skipped but only 41 % of the warp-steps
    // (a probe issued for one lane costs as much as for 32), and the test itself is a ~100-instruction dependent chain with a tile load
    // in the middle: 3.37 ms without, 3.83 / 3.49 / 3.31 / 3.22 ms with a cooldown of 0 / 1 / 3 / 7 rounds (profiles/r02/ssr_variants.log).
    // Bit-identical images either way.
    static const bool skip_enabled = [] { const char *e = getenv("VHR_SSR_SKIP"); return e && atoi(e) != 0; }();
    if (skip_enabled && p.W >= 16 && p.H >= 16 && skip_bound_applies(ctx->pfd)) {
        const int q = ctx->stream == ctx->queue[1] && ctx->queue[1] != nullptr ? 1 : 0;
        const size_t count = (size_t)p.tiles_w * p.tiles_h;
        if (ctx->depth_tiles_count[q] < count) {
            if (ctx->d_depth_tiles[q]) { VHR_CUDA_CHECK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_depth_tiles[q]); ctx->d_depth_tiles[q] = nullptr; }
            VHR_CUDA_CHECK(cudaMalloc(&ctx->d_depth_tiles[q], count * sizeof(float2)));
            ctx->depth_tiles_count[q] = count;
        }
        depth_tiles_kernel<<<(unsigned)((count + 3) / 4), 128, 0, ctx->stream>>>(p.depth, ctx->d_depth_tiles[q], p.W, p.H, p.tiles_w, p.tiles_h);
        VHR_CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
        p.tiles = ctx->d_depth_tiles[q];
        const float *pi = ctx->pfd.camera_proj_inverse;
        p.pi0 = pi[0]; p.pi5 = pi[5]; p.pi14 = pi[14]; p.pi11 = pi[11]; p.pi15 = pi[15]; p.pi14sq = pi[14] * pi[14];
        const float *vi = ctx->pfd.camera_view_inverse;
        p.margin0 = 1e-5f + 4e-6f * (std::fabs(vi[12]) + std::fabs(vi[13]) + std::fabs(vi[14]));
    }
    // camera_proj * camera_view, each element a left-to-right sum of rounded products (volatile keeps the host compiler from
    // contracting or reassociating; the oracle forms the same product)
    const float *A = ctx->pfd.camera_proj, *B = ctx->pfd.camera_view;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            volatile float acc = A[0 * 4 + r] * B[c * 4 + 0];
            for (int k = 1; k < 4; ++k) {
                volatile float prod = A[k * 4 + r] * B[c * 4 + k];
                acc = acc + prod;
            }
            p.pv[c * 4 + r] = acc;
        }
    static const int variant = [] { const char *e = getenv("VHR_SSR_VARIANT"); return e ? atoi(e) : 0; }();
    if (variant != 1) {
        dim3 block(128), grid((p.x_end + 15) / 16, (p.y_end - p.y_begin + 7) / 8);
        ssr_kernel<<<grid, block, 0, ctx->stream>>>(p, ctx->pfd);
    } else {
        // four warps per block, a (8 TX) x (4 TY)-pixel region each
        static const int region = [] { const char *e = getenv("VHR_SSR_REGION"); return e ? atoi(e) : 42; }();
#define VHR_SSR_CASE(ID, TX, TY, G) case ID: ssr_refill_kernel<TX, TY, G><<<dim3((p.x_end + 16 * TX - 1) / (16 * TX), (p.y_end - p.y_begin + 8 * TY - 1) / (8 * TY)), 128, 0, ctx->stream>>>(p, ctx->pfd); break;
        switch (region) {       // tiles across, tiles down, lanes gathered before a set-up / lighting round
            VHR_SSR_CASE(21, 2, 1, 8) VHR_SSR_CASE(22, 2, 1, 4) VHR_SSR_CASE(41, 4, 1, 8) VHR_SSR_CASE(42, 4, 1, 4) VHR_SSR_CASE(81, 4, 2, 8) VHR_SSR_CASE(82, 4, 2, 4) VHR_SSR_CASE(44, 2, 2, 4)
            default: return fail(VHR_ERR_INVALID, "VHR_SSR_REGION = %d", region);
        }
#undef VHR_SSR_CASE
    }
    VHR_CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    return VHR_OK;
}

}  // namespace vhr
