// ssao_kernels.cu — Alchemy screen-space ambient occlusion and its 13x13 box blur (sm_100a).
//
//   ssao_kernel       <- /root/reference/data/shaders/hybrid_render_path/ssao.comp:14-53
//   ssao_blur_kernel  <- /root/reference/data/shaders/hybrid_render_path/ssao_blur.comp:11-26
//
// `texture()` through the reference's default sampler (resource_manager.cpp:58-69: LINEAR, REPEAT) is evaluated in
// software with the Vulkan spec's float weights — CUDA texture units filter with 8-bit fixed-point weights, which
// would not match the CPU oracle.
#include <algorithm>

#include "vhr_internal.h"

namespace vhr {

struct SsaoParams {
    int W, H;
    int x_end, y_begin, y_end;
    float radius;
    const uint2 *normals;     // binding 0 (RGBA16F, sampled)
    const float *depth;       // binding 1 (D32F, sampled)
    uint2 *out;               // binding 2 (RGBA16F)
    HaloPush push;            // multi-GPU: the blur on the neighbours reads 6 rows of this output beyond their band
};

__device__ __forceinline__ void ssao_store(const SsaoParams &p, int gy, size_t pix, uint2 v) {
    p.out[pix] = v;
    if (p.push.rows) {
        if (p.push.up && gy < p.y_begin + p.push.rows) reinterpret_cast<uint2 *>(p.push.up)[pix] = v;
        if (p.push.down && gy >= p.y_end - p.push.rows) reinterpret_cast<uint2 *>(p.push.down)[pix] = v;
    }
}

__device__ __forceinline__ float sample_depth(const SsaoParams &p, float u, float v) {
    int x0, x1, y0, y1;
    float a, b;
    bilinear_setup(u, p.W, x0, x1, a);
    bilinear_setup(v, p.H, y0, y1, b);
    float t00 = __ldg(&p.depth[(size_t)y0 * p.W + x0]), t10 = __ldg(&p.depth[(size_t)y0 * p.W + x1]);
    float t01 = __ldg(&p.depth[(size_t)y1 * p.W + x0]), t11 = __ldg(&p.depth[(size_t)y1 * p.W + x1]);
    return bilerp_rn(a, b, t00, t10, t01, t11);
}

// get_view_space_position (glsl_common.h:111-116) for the 16 SAMPLES of a pixel. What must stay exact in this kernel is the sample POSITION
// (su, sv): the texture unit holds the filter coordinate with 8 fractional bits, so the depth tap is a step function of it — that is the
// centre unprojection (-> perspective radius), the RNG and sincosf, all kept as in the oracle. The unprojected sample itself only enters
// the occlusion sum continuously, so here the three IEEE divisions by w become one MUFU reciprocal + one Newton step (<= 1 ulp) and three
// products, and with PERSPECTIVE (the inverse projection has the sparsity of an inverse perspective matrix: only m00, m11, m23, m32, m33
// non-zero — checked on the host) the 16 products of the matrix-vector product that multiply exact zeros are not issued (same values: the
// pairwise sum of glm's mat4 * vec4 with zero terms is the remaining term).
template <bool PERSPECTIVE>
__device__ __forceinline__ float3 unproject_sample(const float *inv, float depth, float u, float v) {
    const float x = sub_rn(mul_rn(u, 2.0f), 1.0f), y = sub_rn(mul_rn(v, 2.0f), 1.0f);
    float4 q;
    if (PERSPECTIVE) {
        q = make_float4(mul_rn(inv[0], x), mul_rn(inv[5], y), inv[14], add_rn(mul_rn(inv[11], depth), inv[15]));
    } else {
        q = mul44_rn(inv, make_float4(x, y, depth, 1.0f));
    }
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q.w));
    const float rn = fmaf(r, fmaf(-q.w, r, 1.0f), r);        // one Newton step
    r = (q.w == 0.0f) ? r : rn;                              // w = 0 (a sky sample): x * (1 / 0) = x / 0 = +-inf or NaN, as the division gives
    return make_float3(q.x * r, q.y * r, q.z * r);
}

template <bool PERSPECTIVE>
__global__ void __launch_bounds__(256) ssao_kernel(const __grid_constant__ SsaoParams p, const __grid_constant__ PerFrameData pfd) {
    const int gx = blockIdx.x * 32 + threadIdx.x;
    const int gy = p.y_begin + blockIdx.y * 8 + threadIdx.y;
    if (gx >= p.x_end || gy >= p.y_end) return;
    const size_t pix = (size_t)gy * p.W + gx;
    const float cu = mul_rn((float)gx, pfd.display_size_inverse[0]);
    const float cv = mul_rn((float)gy, pfd.display_size_inverse[1]);
    float current_depth = sample_depth(p, cu, cv);
    if (current_depth == 0.0f) {
        ssao_store(p, gy, pix, make_uint2(0u, 0u));
        return;
    }
    float3 P = unproject_rn(pfd.camera_proj_inverse, current_depth, cu, cv);
    float3 nw;
    {
        int x0, x1, y0, y1;
        float a, b;
        bilinear_setup(cu, p.W, x0, x1, a);
        bilinear_setup(cv, p.H, y0, y1, b);
        float4 t00 = unpack_rgba16f(__ldg(&p.normals[(size_t)y0 * p.W + x0]));
        float4 t10 = unpack_rgba16f(__ldg(&p.normals[(size_t)y0 * p.W + x1]));
        float4 t01 = unpack_rgba16f(__ldg(&p.normals[(size_t)y1 * p.W + x0]));
        float4 t11 = unpack_rgba16f(__ldg(&p.normals[(size_t)y1 * p.W + x1]));
        nw = make_float3(bilerp_rn(a, b, t00.x, t10.x, t01.x, t11.x), bilerp_rn(a, b, t00.y, t10.y, t01.y, t11.y),
                         bilerp_rn(a, b, t00.z, t10.z, t01.z, t11.z));
    }
    float3 N = mul33_of44_rn(pfd.camera_view, nw);
    float perspective_radius = __fdiv_rn(p.radius, P.z);
    uint32_t rng = seed_thread(((uint32_t)gy * (uint32_t)pfd.display_size[1] + (uint32_t)gx) * pfd.frame_index);
    float sum = 0.0f;
    for (int i = 0; i < 16; ++i) {
        float ang = mul_rn(mul_rn(random01(rng), 2.0f), VHR_PI);
        float dist = mul_rn(random01(rng), perspective_radius);
        // full-precision sincosf: MUFU sin / cos (abs error ~5e-7) was measured to break the 1e-3 parity bar — across a depth edge the
        // bilinear depth tap is unprojected through znear / depth, which amplifies a 1e-6 shift of the tap (1.5e-3 on one pixel of
        // tests/test_ssao_gpu.py)
        float s, c;
        sincosf(ang, &s, &c);
        float su = add_rn(cu, mul_rn(c, dist)), sv = add_rn(cv, mul_rn(s, dist));
        float3 Q = unproject_sample<PERSPECTIVE>(pfd.camera_proj_inverse, sample_depth(p, su, sv), su, sv);
        float3 V = make_float3(sub_rn(Q.x, P.x), sub_rn(Q.y, P.y), sub_rn(Q.z, P.z));
        // fmaxf returns the non-NaN operand, like the GLSL max() on NVIDIA hardware the oracle restates
        float num = fmaxf(sub_rn(dot3_rn(V, N), 1e-4f), 0.0f);
        sum = add_rn(sum, __fdiv_rn(num, add_rn(dot3_rn(V, V), 1e-4f)));
    }
    float ao = fmaxf(sub_rn(1.0f, mul_rn(0.125f, sum)), 0.0f);
    ssao_store(p, gy, pix, pack_rgba16f(make_float4(ao, ao, ao, ao)));
}

// 13x13 box sum of .x, OOB skipped, always divided by 169. Tile 32x8 outputs; raw values staged in shared memory
// with a 6-texel apron; horizontal 13-sums, then vertical 13-sums.
struct BlurParams {
    int W, H;
    int x_end, y_begin, y_end;
    const uint2 *in;
    uint2 *out;
};

__global__ void __launch_bounds__(256) ssao_blur_kernel(const __grid_constant__ BlurParams p) {
    constexpr int TX = 32, TY = 8, R = 6;
    __shared__ float raw[TY + 2 * R][TX + 2 * R + 1];
    __shared__ float hsum[TY + 2 * R][TX];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = p.y_begin + blockIdx.y * TY;
    for (int idx = tid; idx < (TY + 2 * R) * (TX + 2 * R); idx += TX * TY) {
        int row = idx / (TX + 2 * R), col = idx - row * (TX + 2 * R);
        int gx = x0 - R + col, gy = y0 - R + row;
        float v = 0.0f;
        if (gx >= 0 && gx < p.W && gy >= 0 && gy < p.H)
            v = unpack_rg16f(__ldg(reinterpret_cast<const uint32_t *>(&p.in[(size_t)gy * p.W + gx]))).x;
        raw[row][col] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < (TY + 2 * R) * TX; idx += TX * TY) {
        int row = idx / TX, col = idx - row * TX;
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; ++k) s += raw[row][col + k];
        hsum[row][col] = s;
    }
    __syncthreads();
    const int cx = x0 + threadIdx.x, cy = y0 + threadIdx.y;
    if (cx >= p.x_end || cy >= p.y_end) return;
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) s += hsum[threadIdx.y + k][threadIdx.x];
    float o = __fdiv_rn(s, 169.0f);
    p.out[(size_t)cy * p.W + cx] = pack_rgba16f(make_float4(o, o, o, o));
}

static bool dispatch_range(vhr_context *ctx, const Image *ref, uint32_t xg, uint32_t yg, int &x_end, int &y0, int &y1) {
    x_end = (int)std::min<uint64_t>(ref->width, (uint64_t)xg * 8);
    int y_cov = (int)std::min<uint64_t>(ref->height, (uint64_t)yg * 8);
    y0 = std::max(0, ctx->opt.row_begin);
    y1 = ctx->opt.row_end < 0 ? y_cov : std::min(y_cov, ctx->opt.row_end);
    return x_end > 0 && y1 > y0;
}

int launch_ssao(vhr_context *ctx, uint32_t xg, uint32_t yg, float radius) {
    // descriptor set 3 of the "SSAO Pass" (hybrid_render_path.cpp:143-150): 0 normals, 1 depth, 2 raw output
    if (ctx->n_bound < 3) return fail(VHR_ERR_STATE, "ssao.comp: pass images not bound (need bindings 0..2)");
    Image *normals = ctx->bound[0], *depth = ctx->bound[1], *out = ctx->bound[2];
    if (!normals || !depth || !out) return fail(VHR_ERR_STATE, "ssao.comp: unbound image");
    if (normals->format != VHR_FORMAT_R16G16B16A16_SFLOAT || depth->format != VHR_FORMAT_D32_SFLOAT ||
        out->format != VHR_FORMAT_R16G16B16A16_SFLOAT)
        return fail(VHR_ERR_INVALID, "ssao.comp: unexpected image formats");
    if (depth->width != normals->width || depth->height != normals->height || out->width != normals->width ||
        out->height != normals->height)
        return fail(VHR_ERR_INVALID, "ssao.comp: image sizes differ");
    if (int rc = make_writable(ctx, out, covers_image(ctx, out, (uint64_t)xg * 8, (uint64_t)yg * 8))) return rc;
    SsaoParams p;
    p.W = (int)normals->width; p.H = (int)normals->height;
    if (!dispatch_range(ctx, normals, xg, yg, p.x_end, p.y_begin, p.y_end)) return VHR_OK;
    p.radius = radius;
    p.normals = (const uint2 *)normals->ptr; p.depth = (const float *)depth->ptr; p.out = (uint2 *)out->ptr;
    // multi-GPU (row bands; every rank holds the full depth / normal G-buffer the samples reach into): the 13x13 blur on the
    // neighbours reads 6 rows of this output beyond their band -> pushed by this kernel, then the flag-word round trip
    p.push = halo_push_for(ctx, out, false, 6);
    if (p.push.rows && ((p.push.up == nullptr && ctx->part.rank > 0) || (p.push.down == nullptr && ctx->part.rank + 1 < ctx->part.world)))
        return fail(VHR_ERR_STATE, "ssao.comp: the neighbours' '%s' image is not attached (vhr_image_attach_peer)", "Screen Space Ambient Occlusion Raw");
    dim3 block(32, 8), grid((p.x_end + 31) / 32, (p.y_end - p.y_begin + 7) / 8);
    // inverse perspective sparsity (column-major m[c * 4 + r]): everything but m00, m11, m23 (index 11), m32 (index 14), m33 (index 15) is zero
    bool perspective = true;
    for (int i = 0; i < 16; ++i)
        if (i != 0 && i != 5 && i != 11 && i != 14 && i != 15 && ctx->pfd.camera_proj_inverse[i] != 0.0f) perspective = false;
    if (perspective) ssao_kernel<true><<<grid, block, 0, ctx->stream>>>(p, ctx->pfd);
    else ssao_kernel<false><<<grid, block, 0, ctx->stream>>>(p, ctx->pfd);
    VHR_CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    return p.push.rows ? peer_sync_neighbours(ctx) : VHR_OK;
}

int launch_ssao_blur(vhr_context *ctx, uint32_t xg, uint32_t yg) {
    // "SSAO Blur Pass" (hybrid_render_path.cpp:170-177): 0 raw, 1 blurred
    if (ctx->n_bound < 2) return fail(VHR_ERR_STATE, "ssao_blur.comp: pass images not bound (need bindings 0..1)");
    Image *in = ctx->bound[0], *out = ctx->bound[1];
    if (!in || !out) return fail(VHR_ERR_STATE, "ssao_blur.comp: unbound image");
    if (in->format != VHR_FORMAT_R16G16B16A16_SFLOAT || out->format != VHR_FORMAT_R16G16B16A16_SFLOAT)
        return fail(VHR_ERR_INVALID, "ssao_blur.comp: unexpected image formats");
    if (in->width != out->width || in->height != out->height) return fail(VHR_ERR_INVALID, "ssao_blur.comp: image sizes differ");
    if (ctx->pfd.display_size[0] != (float)in->width || ctx->pfd.display_size[1] != (float)in->height)
        return fail(VHR_ERR_INVALID, "ssao_blur.comp: PerFrameData.display_size does not match the images");
    if (int rc = make_writable(ctx, out, covers_image(ctx, out, (uint64_t)xg * 8, (uint64_t)yg * 8))) return rc;
    BlurParams p;
    p.W = (int)in->width; p.H = (int)in->height;
    if (!dispatch_range(ctx, in, xg, yg, p.x_end, p.y_begin, p.y_end)) return VHR_OK;
    p.in = (const uint2 *)in->ptr; p.out = (uint2 *)out->ptr;
    dim3 block(32, 8), grid((p.x_end + 31) / 32, (p.y_end - p.y_begin + 7) / 8);
    ssao_blur_kernel<<<grid, block, 0, ctx->stream>>>(p);
    VHR_CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    return VHR_OK;
}

}  // namespace vhr
