// svgf_kernels.cu — SVGF temporal accumulation + variance, and the à-trous wavelet filter (sm_100a).
//
//   svgf_temporal_kernel      <- /root/reference/data/shaders/hybrid_render_path/svgf.comp:16-145
//   atrous_direct_kernel      <- /root/reference/data/shaders/hybrid_render_path/svgf_atrous_filter.comp:17-103
//                                (one thread per pixel, 49 direct image loads: the reference's own dataflow)
//   atrous_tiled_kernel<S>    <- same arithmetic, B200 dataflow: a 64x8 output tile whose rows are S apart, staged
//                                once in shared memory as fp32 planes, so each texel is converted once instead of 25x
//                                and every tap is two 128-bit LDS; (shadow, ao) channel pairs run on packed
//                                fp32x2 instructions (FADD2/FMUL2/FFMA2).
//
// Images: dense row-major; RGBA16F texel = uint2, RG16F texel = uint32. fp32 math, fp16 RTE stores.
#include <algorithm>

#include "vhr_internal.h"

namespace vhr {

// ---------------------------------------------------------------------------------------------------------------
// Temporal pass
// ---------------------------------------------------------------------------------------------------------------
struct TemporalParams {
    int W, H;
    int x_end, y_begin, y_end;     // pixel range covered by this dispatch
    float dsx, dsy;                // pfd.display_size
    const uint2 *normals;          // binding 0
    const uint2 *motion;           // binding 1
    const uint32_t *rt;            // binding 3 (RG16F)
    const uint2 *prev_normals;     // storage_images[pc.prev_frame_normals_and_object_ids]
    const uint2 *history;          // storage_images[pc.shadow_and_ao_history]
    const uint32_t *moments_in;    // storage_images[pc.shadow_and_ao_moments_history] (previous frame, Q11 snapshot)
    uint2 *integrated_out;         // storage_images[pc.integrated_shadow_and_ao[0]]
    uint32_t *moments_out;
};

// svgf.comp:16-39
__device__ __forceinline__ bool is_valid_reprojection(const TemporalParams &p, int px, int py, int cur_id, float3 cur_n) {
    if (px < 0 || py < 0 || (float)px >= p.dsx || (float)py >= p.dsy) return false;
    float4 pn = unpack_rgba16f(__ldg(&p.prev_normals[(size_t)py * p.W + px]));
    if (cur_id != f2i_rz(pn.w)) return false;
    if (dot3_rn(cur_n, make_float3(pn.x, pn.y, pn.z)) < VHR_COS_PI_4) return false;
    return true;
}

__device__ __forceinline__ float mix_rn(float a, float b, float t) { return add_rn(mul_rn(a, sub_rn(1.0f, t)), mul_rn(b, t)); }

__global__ void __launch_bounds__(256) svgf_temporal_kernel(const __grid_constant__ TemporalParams p) {
    const int cx = blockIdx.x * 32 + threadIdx.x;
    const int cy = p.y_begin + blockIdx.y * 8 + threadIdx.y;
    if (cx >= p.x_end || cy >= p.y_end) return;
    const size_t pix = (size_t)cy * p.W + cx;

    float4 cn = unpack_rgba16f(__ldg(&p.normals[pix]));
    float3 cur_n = make_float3(cn.x, cn.y, cn.z);
    int cur_id = f2i_rz(cn.w);
    float2 mv = unpack_rg16f(__ldg(reinterpret_cast<const uint32_t *>(&p.motion[pix])));   // .xy only
    float2 cur = unpack_rg16f(__ldg(&p.rt[pix]));

    // svgf.comp:52-55
    float pcx = add_rn(sub_rn((float)cx, mul_rn(mv.x, p.dsx)), 0.5f);
    float pcy = add_rn(sub_rn((float)cy, mul_rn(mv.y, p.dsy)), 0.5f);
    float x = sub_rn(pcx, floorf(pcx)), y = sub_rn(pcy, floorf(pcy));
    int ax = f2i_rz(pcx), ay = f2i_rz(pcy);
    float omx = sub_rn(1.0f, x), omy = sub_rn(1.0f, y);
    const float bw[4] = {mul_rn(omx, omy), mul_rn(x, omy), mul_rn(omx, y), mul_rn(x, y)};

    float prev_s = 0.0f, prev_a = 0.0f, sum = 0.0f;
    float psm0 = 0.0f, psm1 = 0.0f, pam0 = 0.0f, pam1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int sx = ax + (i & 1), sy = ay + (i >> 1);
        if (is_valid_reprojection(p, sx, sy, cur_id, cur_n)) {
            size_t sp = (size_t)sy * p.W + sx;
            float2 h = unpack_rg16f(__ldg(reinterpret_cast<const uint32_t *>(&p.history[sp])));
            float2 m = unpack_rg16f(__ldg(&p.moments_in[sp]));
            prev_s = add_rn(prev_s, mul_rn(bw[i], h.x));
            prev_a = add_rn(prev_a, mul_rn(bw[i], h.y));
            psm0 = add_rn(psm0, mul_rn(bw[i], m.x));
            psm1 = add_rn(psm1, mul_rn(bw[i], m.y));
            // a two-channel image reads back (r, g, 0, 1): prev AO moments accumulate (0, w) — SURVEY Q2
            pam1 = add_rn(pam1, bw[i]);
            sum = add_rn(sum, bw[i]);
        }
    }
    bool valid = sum > 1e-6f;
    if (!valid) {   // svgf.comp:81-97, accumulators not reset
        for (int yy = -1; yy <= 1; ++yy)
            for (int xx = -1; xx <= 1; ++xx) {
                int sx = ax + xx, sy = ay + yy;
                if (is_valid_reprojection(p, sx, sy, cur_id, cur_n)) {
                    size_t sp = (size_t)sy * p.W + sx;
                    float2 h = unpack_rg16f(__ldg(reinterpret_cast<const uint32_t *>(&p.history[sp])));
                    float2 m = unpack_rg16f(__ldg(&p.moments_in[sp]));
                    prev_s = add_rn(prev_s, h.x);
                    prev_a = add_rn(prev_a, h.y);
                    psm0 = add_rn(psm0, m.x);
                    psm1 = add_rn(psm1, m.y);
                    pam1 = add_rn(pam1, 1.0f);
                    sum = add_rn(sum, 1.0f);
                }
            }
        valid = sum > 1e-6f;
    }

    float sm0 = cur.x, sm1 = mul_rn(cur.x, cur.x);
    float am0 = cur.y, am1 = mul_rn(cur.y, cur.y);
    float4 out;
    if (valid) {
        prev_s = __fdiv_rn(prev_s, sum);
        psm0 = __fdiv_rn(psm0, sum); psm1 = __fdiv_rn(psm1, sum);
        prev_a = __fdiv_rn(prev_a, sum);
        pam0 = __fdiv_rn(pam0, sum); pam1 = __fdiv_rn(pam1, sum);
        sm0 = mix_rn(psm0, sm0, 0.2f); sm1 = mix_rn(psm1, sm1, 0.2f);
        am0 = mix_rn(pam0, am0, 0.2f); am1 = mix_rn(pam1, am1, 0.2f);
        float sv = fmaxf(0.0f, sub_rn(sm1, mul_rn(sm0, sm0)));
        float av = fmaxf(0.0f, sub_rn(am1, mul_rn(am0, am0)));
        out = make_float4(mix_rn(prev_s, cur.x, 0.2f), mix_rn(prev_a, cur.y, 0.2f), sv, av);
    } else {
        float sv = fmaxf(0.0f, sub_rn(sm1, mul_rn(sm0, sm0)));
        float av = fmaxf(0.0f, sub_rn(am1, mul_rn(am0, am0)));
        out = make_float4(cur.x, cur.y, sv, av);
    }
    p.integrated_out[pix] = pack_rgba16f(out);
    p.moments_out[pix] = pack_rg16f(sm0, sm1);   // RG16F keeps the shadow moments only (Q2)
}

// ---------------------------------------------------------------------------------------------------------------
// À-trous, variant 0: the reference's dataflow (one thread per pixel, direct loads)
// ---------------------------------------------------------------------------------------------------------------
struct AtrousParams {
    int W, H;
    int x_end, y_begin, y_end;
    int step;
    float dsx, dsy;
    const uint2 *normals;      // binding 0
    const uint2 *integ_in;     // storage_images[pc.integrated_shadow_and_ao[0]]
    uint2 *integ_out;          // storage_images[pc.integrated_shadow_and_ao[1]]
};

__constant__ float c_atrous_h[5] = {1.0f / 16, 1.0f / 4, 3.0f / 8, 1.0f / 4, 1.0f / 16};

__device__ __forceinline__ bool oob(const AtrousParams &p, int sx, int sy) {
    return sx < 0 || (float)sx >= p.dsx || sy < 0 || (float)sy >= p.dsy;
}

__global__ void __launch_bounds__(256) atrous_direct_kernel(const __grid_constant__ AtrousParams p) {
    const int cx = blockIdx.x * 32 + threadIdx.x;
    const int cy = p.y_begin + blockIdx.y * 8 + threadIdx.y;
    if (cx >= p.x_end || cy >= p.y_end) return;
    const size_t pix = (size_t)cy * p.W + cx;
    float4 np4 = unpack_rgba16f(__ldg(&p.normals[pix]));
    int id_p = f2i_rz(np4.w);
    float4 ip = unpack_rgba16f(__ldg(&p.integ_in[pix]));

    // gauss_3x3_filter: weights 1/16 1/8 1/16 ..., OOB skipped, row-major accumulation order
    float var_s = 0.0f, var_a = 0.0f;
#pragma unroll
    for (int y = -1; y <= 1; ++y)
#pragma unroll
        for (int x = -1; x <= 1; ++x) {
            int sx = cx + x, sy = cy + y;
            if (oob(p, sx, sy)) continue;
            float w = (x == 0 ? 0.5f : 0.25f) * (y == 0 ? 0.5f : 0.25f);
            float4 q = unpack_rgba16f(__ldg(&p.integ_in[(size_t)sy * p.W + sx]));
            var_s = add_rn(var_s, mul_rn(w, q.z));
            var_a = add_rn(var_a, mul_rn(w, q.w));
        }
    float den_s = add_rn(mul_rn(4.0f, sqrtf(var_s)), 1e-6f);
    float den_a = add_rn(mul_rn(4.0f, sqrtf(var_a)), 1e-6f);

    float swx = 1.0f, swy = 1.0f;
    float4 sum = ip;
    for (int y = -2; y <= 2; ++y)
        for (int x = -2; x <= 2; ++x) {
            int sx = cx + x * p.step, sy = cy + y * p.step;
            if (oob(p, sx, sy) || (x == 0 && y == 0)) continue;
            size_t sp = (size_t)sy * p.W + sx;
            float4 iq = unpack_rgba16f(__ldg(&p.integ_in[sp]));
            float4 nq = unpack_rgba16f(__ldg(&p.normals[sp]));
            float kernel = c_atrous_h[y + 2] * c_atrous_h[x + 2];
            float d = add_rn(add_rn(mul_rn(np4.x, nq.x), mul_rn(np4.y, nq.y)), mul_rn(np4.z, nq.z));
            float wn = (d > 0.0f) ? fmaxf(0.0f, powf(d, 128.0f)) : 0.0f;
            float wid = (id_p == f2i_rz(nq.w)) ? 1.0f : 0.0f;
            float wk = mul_rn(mul_rn(kernel, wn), wid);
            float wx = mul_rn(wk, expf(-__fdiv_rn(fabsf(sub_rn(ip.x, iq.x)), den_s)));
            float wy = mul_rn(wk, expf(-__fdiv_rn(fabsf(sub_rn(ip.y, iq.y)), den_a)));
            swx = add_rn(swx, wx);
            swy = add_rn(swy, wy);
            sum.x = add_rn(sum.x, mul_rn(wx, iq.x));
            sum.y = add_rn(sum.y, mul_rn(wy, iq.y));
            sum.z = add_rn(sum.z, mul_rn(mul_rn(wx, wx), iq.z));
            sum.w = add_rn(sum.w, mul_rn(mul_rn(wy, wy), iq.w));
        }
    float4 out = make_float4(__fdiv_rn(sum.x, swx), __fdiv_rn(sum.y, swy), __fdiv_rn(sum.z, mul_rn(swx, swx)),
                             __fdiv_rn(sum.w, mul_rn(swy, swy)));
    p.integ_out[pix] = pack_rgba16f(out);
}

// ---------------------------------------------------------------------------------------------------------------
// À-trous, variant 1: shared-memory tile with rows S apart + packed fp32x2 math
// ---------------------------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE fp32 operations per issue slot.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int AT_TX = 64;    // output tile width (pixels, contiguous)
constexpr int AT_TY = 8;     // output tile height (rows, S apart)

template <int S>
__global__ void __launch_bounds__(AT_TX * AT_TY, 2) atrous_tiled_kernel(const __grid_constant__ AtrousParams p) {
    constexpr int COLS = AT_TX + 4 * S;
    constexpr int ROWS = AT_TY + 4;
    extern __shared__ float4 smem[];
    float4 *s_n = smem;                    // [ROWS][COLS] normal.xyz, object id (as float)
    float4 *s_i = smem + ROWS * COLS;      // [ROWS][COLS] shadow, ao, var_shadow, var_ao

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * AT_TX + tx;
    const int x0 = blockIdx.x * AT_TX;
    // blockIdx.y enumerates (super-tile k, residue r): rows y = y_begin + k*S*TY + r + S*j
    const int k = blockIdx.y / S, r = blockIdx.y % S;
    const int yb = p.y_begin + k * (S * AT_TY) + r;

    // ---- stage the tile: each texel is loaded and converted to fp32 once -----------------------------------
    for (int idx = tid; idx < ROWS * COLS; idx += AT_TX * AT_TY) {
        int row = idx / COLS, col = idx - row * COLS;
        int gx = x0 - 2 * S + col;
        int gy = yb + (row - 2) * S;
        float4 n = make_float4(0.0f, 0.0f, 0.0f, -1.0f);   // out of bounds: zero normal => weight 0 ("skipped")
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (gx >= 0 && gx < p.W && gy >= 0 && gy < p.H) {
            size_t sp = (size_t)gy * p.W + gx;
            n = unpack_rgba16f(__ldg(&p.normals[sp]));
            v = unpack_rgba16f(__ldg(&p.integ_in[sp]));
            n.w = (float)f2i_rz(n.w);
        }
        s_n[idx] = n;
        s_i[idx] = v;
    }
    __syncthreads();

    const int cx = x0 + tx;
    const int cy = yb + ty * S;
    if (cx >= p.x_end || cy >= p.y_end) return;

    const int crow = ty + 2, ccol = tx + 2 * S;
    const float4 np4 = s_n[crow * COLS + ccol];
    const float4 ip = s_i[crow * COLS + ccol];

    // ---- 3x3 gaussian of the variance (same accumulation order as the reference) ----------------------------
    float2 var = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int y = -1; y <= 1; ++y)
#pragma unroll
        for (int x = -1; x <= 1; ++x) {
            const float w = (x == 0 ? 0.5f : 0.25f) * (y == 0 ? 0.5f : 0.25f);
            float2 q;
            if (S == 1) {
                float4 t = s_i[(crow + y) * COLS + ccol + x];   // OOB texels staged as 0 => contribute 0
                q = make_float2(t.z, t.w);
            } else if (y == 0) {
                float4 t = s_i[crow * COLS + ccol + x];
                q = make_float2(t.z, t.w);
            } else {
                int sx = cx + x, sy = cy + y;
                q = make_float2(0.0f, 0.0f);
                if (sx >= 0 && sx < p.W && sy >= 0 && sy < p.H)
                    q = unpack_rg16f(__ldg(reinterpret_cast<const uint32_t *>(&p.integ_in[(size_t)sy * p.W + sx]) + 1));
            }
            var.x = add_rn(var.x, mul_rn(w, q.x));
            var.y = add_rn(var.y, mul_rn(w, q.y));
        }
    // exp(-|dl| / (4 sqrt(var) + 1e-6)) = exp2(|dl| * nk), nk = -log2(e) / (4 sqrt(var) + 1e-6)
    const float LOG2E = 1.4426950408889634f;
    float2 nk = make_float2(__fdiv_rn(-LOG2E, add_rn(mul_rn(4.0f, sqrtf(var.x)), 1e-6f)),
                            __fdiv_rn(-LOG2E, add_rn(mul_rn(4.0f, sqrtf(var.y)), 1e-6f)));
    const float2 lp = make_float2(ip.x, ip.y);
    const float2 neg_lp = make_float2(-ip.x, -ip.y);

    float2 sw = make_float2(1.0f, 1.0f);
    float2 sc = make_float2(ip.x, ip.y);     // colour sums
    float2 sv = make_float2(ip.z, ip.w);     // variance sums
    (void)lp;
#pragma unroll
    for (int y = -2; y <= 2; ++y) {
#pragma unroll
        for (int x = -2; x <= 2; ++x) {
            if (x == 0 && y == 0) continue;
            const int o = (crow + y) * COLS + ccol + x * S;
            const float4 nq = s_n[o];
            const float4 iq = s_i[o];
            const float kernel = ((y == 0) ? 0.375f : ((y == 1 || y == -1) ? 0.25f : 0.0625f)) *
                                 ((x == 0) ? 0.375f : ((x == 1 || x == -1) ? 0.25f : 0.0625f));
            float d = np4.x * nq.x + np4.y * nq.y + np4.z * nq.z;
            float d2 = d * d, d4 = d2 * d2, d8 = d4 * d4, d16 = d8 * d8, d32 = d16 * d16, d64 = d32 * d32;
            float wn = d64 * d64;                                                   // d^128
            float wk = (d > 0.0f && nq.w == np4.w) ? kernel * wn : 0.0f;            // Q10 + object-id edge stop
            float2 dl = add2(make_float2(iq.x, iq.y), neg_lp);                      // l_q - l_p
            float2 e = mul2(make_float2(fabsf(dl.x), fabsf(dl.y)), nk);
            float2 w = mul2(make_float2(wk, wk), make_float2(ex2_approx(e.x), ex2_approx(e.y)));
            sw = add2(sw, w);
            sc = fma2(w, make_float2(iq.x, iq.y), sc);
            sv = fma2(mul2(w, w), make_float2(iq.z, iq.w), sv);
        }
    }
    float4 out = make_float4(__fdiv_rn(sc.x, sw.x), __fdiv_rn(sc.y, sw.y), __fdiv_rn(sv.x, mul_rn(sw.x, sw.x)),
                             __fdiv_rn(sv.y, mul_rn(sw.y, sw.y)));
    p.integ_out[(size_t)cy * p.W + cx] = pack_rgba16f(out);
}

template <int S>
static int launch_tiled(vhr_context *ctx, const AtrousParams &p, int x_pixels, int y_pixels) {
    constexpr int COLS = AT_TX + 4 * S, ROWS = AT_TY + 4;
    constexpr size_t smem = (size_t)2 * ROWS * COLS * sizeof(float4);
    static_assert(smem <= 48 * 1024, "tile must fit the default dynamic shared memory limit");
    dim3 block(AT_TX, AT_TY);
    dim3 grid((x_pixels + AT_TX - 1) / AT_TX, ((y_pixels + S * AT_TY - 1) / (S * AT_TY)) * S);
    atrous_tiled_kernel<S><<<grid, block, smem, ctx->stream>>>(p);
    VHR_CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    return VHR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------------------------
static bool dispatch_range(vhr_context *ctx, const Image *ref, uint32_t xg, uint32_t yg, int &x_end, int &y0, int &y1) {
    x_end = (int)std::min<uint64_t>(ref->width, (uint64_t)xg * 8);
    int y_cov = (int)std::min<uint64_t>(ref->height, (uint64_t)yg * 8);
    y0 = std::max(0, ctx->opt.row_begin);
    y1 = ctx->opt.row_end < 0 ? y_cov : std::min(y_cov, ctx->opt.row_end);
    return x_end > 0 && y1 > y0;
}

static int check_image(const Image *im, int fmt, const Image *ref, const char *what) {
    if (!im || !im->ptr) return fail(VHR_ERR_INVALID, "%s: image not bound / unknown storage slot", what);
    if (im->format != fmt) return fail(VHR_ERR_INVALID, "%s: format %d, expected %d", what, im->format, fmt);
    if (ref && (im->width != ref->width || im->height != ref->height))
        return fail(VHR_ERR_INVALID, "%s: size %ux%u differs from %ux%u", what, im->width, im->height, ref->width, ref->height);
    return VHR_OK;
}

int launch_svgf_temporal(vhr_context *ctx, uint32_t xg, uint32_t yg, const SVGFPushConstants &pc) {
    // descriptor set 3 of the "SVGF Denoise Pass" (hybrid_render_path.cpp:264-272): 0 normals, 1 motion, 2 depth, 3 rt, 4 denoised
    if (ctx->n_bound < 4) return fail(VHR_ERR_STATE, "svgf.comp: pass images not bound (need bindings 0..3)");
    Image *normals = ctx->bound[0], *motion = ctx->bound[1], *rt = ctx->bound[3];
    Image *integ0 = storage_slot(ctx, pc.integrated_shadow_and_ao[0]);
    Image *prevn = storage_slot(ctx, pc.prev_frame_normals_and_object_ids);
    Image *hist = storage_slot(ctx, pc.shadow_and_ao_history);
    Image *mom = storage_slot(ctx, pc.shadow_and_ao_moments_history);
    int rc;
    if ((rc = check_image(normals, VHR_FORMAT_R16G16B16A16_SFLOAT, nullptr, "svgf.comp normals"))) return rc;
    if ((rc = check_image(motion, VHR_FORMAT_R16G16B16A16_SFLOAT, normals, "svgf.comp motion"))) return rc;
    if ((rc = check_image(rt, VHR_FORMAT_R16G16_SFLOAT, normals, "svgf.comp raytraced shadow/ao"))) return rc;
    if ((rc = check_image(integ0, VHR_FORMAT_R16G16B16A16_SFLOAT, normals, "svgf.comp integrated[0]"))) return rc;
    if ((rc = check_image(prevn, VHR_FORMAT_R16G16B16A16_SFLOAT, normals, "svgf.comp prev normals"))) return rc;
    if ((rc = check_image(hist, VHR_FORMAT_R16G16B16A16_SFLOAT, normals, "svgf.comp history"))) return rc;
    if ((rc = check_image(mom, VHR_FORMAT_R16G16_SFLOAT, normals, "svgf.comp moments"))) return rc;
    if (!mom->twin) {   // second moments buffer for the snapshot semantics
        VHR_CUDA_CHECK(cudaMalloc(&mom->twin, mom->bytes));
        VHR_CUDA_CHECK(cudaMemsetAsync(mom->twin, 0, mom->bytes, ctx->stream));
    }
    TemporalParams p;
    p.W = (int)normals->width; p.H = (int)normals->height;
    if (!dispatch_range(ctx, normals, xg, yg, p.x_end, p.y_begin, p.y_end)) return VHR_OK;
    p.dsx = ctx->pfd.display_size[0]; p.dsy = ctx->pfd.display_size[1];
    p.normals = (const uint2 *)normals->ptr; p.motion = (const uint2 *)motion->ptr; p.rt = (const uint32_t *)rt->ptr;
    p.prev_normals = (const uint2 *)prevn->ptr; p.history = (const uint2 *)hist->ptr;
    p.moments_in = (const uint32_t *)mom->ptr; p.integrated_out = (uint2 *)integ0->ptr; p.moments_out = (uint32_t *)mom->twin;
    dim3 block(32, 8), grid((p.x_end + 31) / 32, (p.y_end - p.y_begin + 7) / 8);
    svgf_temporal_kernel<<<grid, block, 0, ctx->stream>>>(p);
    VHR_CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    std::swap(mom->ptr, mom->twin);   // this frame's moments become "the" moments image
    return VHR_OK;
}

int launch_svgf_atrous(vhr_context *ctx, uint32_t xg, uint32_t yg, const SVGFPushConstants &pc) {
    if (ctx->n_bound < 1) return fail(VHR_ERR_STATE, "svgf_atrous_filter.comp: pass images not bound");
    Image *normals = ctx->bound[0];
    Image *in = storage_slot(ctx, pc.integrated_shadow_and_ao[0]);
    Image *out = storage_slot(ctx, pc.integrated_shadow_and_ao[1]);
    int rc;
    if ((rc = check_image(normals, VHR_FORMAT_R16G16B16A16_SFLOAT, nullptr, "atrous normals"))) return rc;
    if ((rc = check_image(in, VHR_FORMAT_R16G16B16A16_SFLOAT, normals, "atrous integrated[0]"))) return rc;
    if ((rc = check_image(out, VHR_FORMAT_R16G16B16A16_SFLOAT, normals, "atrous integrated[1]"))) return rc;
    if (in == out) return fail(VHR_ERR_INVALID, "atrous: integrated[0] and [1] are the same image");
    if (pc.atrous_step < 1) return fail(VHR_ERR_INVALID, "atrous: step %d < 1", pc.atrous_step);
    AtrousParams p;
    p.W = (int)normals->width; p.H = (int)normals->height;
    if (!dispatch_range(ctx, normals, xg, yg, p.x_end, p.y_begin, p.y_end)) return VHR_OK;
    p.step = pc.atrous_step;
    p.dsx = ctx->pfd.display_size[0]; p.dsy = ctx->pfd.display_size[1];
    p.normals = (const uint2 *)normals->ptr; p.integ_in = (const uint2 *)in->ptr; p.integ_out = (uint2 *)out->ptr;
    // The tiled kernel bounds-checks against the image size; the reference checks against pfd.display_size. They are
    // the same thing whenever the UBO matches the images, which the tiled path requires.
    bool tiled_ok = ctx->opt.atrous_variant == 1 && p.dsx == (float)p.W && p.dsy == (float)p.H;
    if (tiled_ok) {
        int xp = p.x_end, yp = p.y_end - p.y_begin;
        switch (p.step) {
            case 1: return launch_tiled<1>(ctx, p, xp, yp);
            case 2: return launch_tiled<2>(ctx, p, xp, yp);
            case 4: return launch_tiled<4>(ctx, p, xp, yp);
            case 8: return launch_tiled<8>(ctx, p, xp, yp);
            case 16: return launch_tiled<16>(ctx, p, xp, yp);
            default: break;   // other steps: direct kernel
        }
    }
    dim3 block(32, 8), grid((p.x_end + 31) / 32, (p.y_end - p.y_begin + 7) / 8);
    atrous_direct_kernel<<<grid, block, 0, ctx->stream>>>(p);
    VHR_CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    return VHR_OK;
}

}  // namespace vhr
