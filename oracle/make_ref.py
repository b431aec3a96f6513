"""oracle/make_ref.py — TEST INFRASTRUCTURE ONLY. Builds oracle/_ref/libvhr_ref.so = the REFERENCE'S OWN SHADERS compiled for the CPU.

The GLSL files are read from /root/reference/data/shaders (and src/rendering_backend/glsl_common.h) where they lie, at build
time; nothing is copied into this repository: the generated C++ lives in oracle/_ref/gen/ (git-ignored) only while it is being compiled
(`--keep` on the command line leaves it there for debugging) and what stays is oracle/_ref/libvhr_ref.so.
Every shader becomes a C++ namespace whose body is the shader's text after a purely lexical translation:

  * `#include` resolved textually, `#version` / `#extension` dropped, glsl_common.h taken on its GLSL (`#ifndef __cplusplus`) side;
  * `layout(...) uniform image2D x;` -> `image2D x;` and the like (descriptor declarations become plain globals the harness
    binds), `layout(push_constant) uniform B { T pc; };` -> `T pc;`, ray payload / hit attribute / stage in-out declarations
    -> globals, specialisation constants -> settable globals;
  * `inout T x` -> `T &x`;
  * floating literals get an `f` suffix (a GLSL `0.5` is a 32-bit float, a C++ `0.5` is a double);
  * `main` -> `shader_main`; `discard` -> flag + return;
  * `vec3 albedo = texture(albedo, uv).rgb;` (a local shadowing the uniform it is initialised from: legal GLSL, ill-formed C++)
    gets the uniform qualified with its namespace.

Types, swizzles and built-ins come from the reference's vendored glm through oracle/ref_shim.h, which also supplies what a
Vulkan driver would: images, samplers, traceRayEXT (bound to the oracle's BVH). The harness below each shader (this file,
HARNESS_*) plays vkCmdDispatch / vkCmdTraceRaysKHR / vkCmdDraw: it binds the arrays and loops over the invocations.

Only tests/ (and __graft_entry__.build(), which merely compiles it) use the result. /root/reference does not exist on the GPU
box: the prebuilt library travels there with the repo snapshot.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VHR_REFERENCE_ROOT", "/root/reference")
SHADER_DIR = os.path.join(REF_ROOT, "data", "shaders")
OUT_DIR = os.path.join(HERE, "_ref")
GEN_DIR = os.path.join(OUT_DIR, "gen")
LIB = os.path.join(OUT_DIR, "libvhr_ref.so")
CXX = "g++"
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-w",
            "-I", os.path.join(REF_ROOT, "dependencies"), "-I", HERE]


def reference_available():
    return os.path.isfile(os.path.join(SHADER_DIR, "hybrid_render_path", "svgf.comp"))


# ---------------------------------------------------------------------------------------------------------------------------
# GLSL text -> C++ text
# ---------------------------------------------------------------------------------------------------------------------------
def strip_comments(text):
    def repl(m):
        s = m.group(0)
        return re.sub(r"[^\n]", " ", s) if s.startswith("/") else s
    return re.sub(r"//[^\n]*|/\*.*?\*/|\"(?:\\.|[^\"\\])*\"", repl, text, flags=re.S)


def load_glsl(path, seen=None):
    """Reads a GLSL file with its #includes inlined; `#ifdef/#ifndef __cplusplus` blocks are resolved as GLSL (undefined)."""
    seen = seen or set()
    out, stack = [], []          # stack of booleans: is the current conditional block live
    with open(path) as f:
        src = f.read().replace("\\\n", "\n")        # line continuations (svgf_atrous_filter.comp:40 ends in a stray backslash)
    for line in src.split("\n"):
        s = line.strip()
        if re.match(r"#\s*ifdef\s+__cplusplus", s):
            stack.append(False); continue
        if re.match(r"#\s*ifndef\s+__cplusplus", s):
            stack.append(True); continue
        if re.match(r"#\s*if(def|ndef)?\b", s):
            stack.append(None)                       # some other conditional: kept verbatim
        elif re.match(r"#\s*endif", s) and stack:
            top = stack.pop()
            if top is not None:
                continue
        elif re.match(r"#\s*else", s) and stack and stack[-1] is not None:
            stack[-1] = not stack[-1]; continue
        if any(v is False for v in stack):
            continue
        m = re.match(r'#\s*include\s+"([^"]+)"', s)
        if m:
            inc = os.path.normpath(os.path.join(os.path.dirname(path), m.group(1)))
            if inc not in seen:
                seen.add(inc)
                out.append(load_glsl(inc, seen))
            continue
        if re.match(r"#\s*(version|extension|pragma)\b", s):
            continue
        out.append(line)
    return "\n".join(out)


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)[fF]?(?![\w.])")


def translate(path, ns):
    """Returns (C++ text of the shader body, metadata about the declarations that became globals)."""
    text = strip_comments(load_glsl(path))
    meta = {"images": [], "samplers": [], "payload_out": {}, "payload_in": None, "stage_in": [], "stage_out": [], "spec": [], "uniform_names": []}
    TL = "thread_local "

    text = re.sub(r"layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;", "", text)

    def image_decl(m):
        binding = re.search(r"binding\s*=\s*(\d+)", m.group(1))
        sett = re.search(r"set\s*=\s*(\d+)", m.group(1))
        kind, name, arr = m.group(2), m.group(3), m.group(4)
        meta["uniform_names"].append(name)
        if arr:
            return f"{TL}{kind} *{name};"
        if int(sett.group(1)) == 3:
            meta["images" if kind == "image2D" else "samplers"].append((int(binding.group(1)), name))
        return f"{TL}{kind} {name};"
    text = re.sub(r"layout\s*\(([^)]*)\)\s*(?:readonly\s+|writeonly\s+)?uniform\s+(image2D|sampler2D|accelerationStructureEXT)\s+(\w+)\s*(\[\s*\])?\s*;",
                  image_decl, text)
    text = re.sub(r"layout\s*\([^)]*\)\s*buffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", lambda m: f"{TL}const {m.group(1)} *{m.group(2)};", text)
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*;\s*\}\s*;", lambda m: f"{TL}{m.group(1)} {m.group(2)};", text)

    def payload_decl(m):
        loc, kind, typ, name = int(m.group(1)), m.group(2), m.group(3), m.group(4)
        if kind == "rayPayloadEXT":
            meta["payload_out"][loc] = (typ, name)
        else:
            meta["payload_in"] = (loc, typ, name)
        return f"{TL}{typ} {name};"
    text = re.sub(r"layout\s*\(\s*location\s*=\s*(\d+)\s*\)\s*(rayPayloadEXT|rayPayloadInEXT)\s+(\w+)\s+(\w+)\s*;", payload_decl, text)
    text = re.sub(r"hitAttributeEXT\s+(\w+)\s+(\w+)\s*;", lambda m: f"{TL}{m.group(1)} {m.group(2)};", text)

    def stage_io(m):
        meta["stage_in" if m.group(2) == "in" else "stage_out"].append((int(m.group(1)), m.group(3), m.group(4)))
        return f"{TL}{m.group(3)} {m.group(4)};"
    text = re.sub(r"layout\s*\(\s*location\s*=\s*(\d+)\s*\)\s*(in|out)\s+(\w+)\s+(\w+)\s*;", stage_io, text)

    def spec_const(m):
        meta["spec"].append(m.group(2))
        return f"{TL}{m.group(1)} {m.group(2)} = {m.group(3)};"
    text = re.sub(r"layout\s*\(\s*constant_id\s*=\s*\w+\s*\)\s*const\s+(\w+)\s+(\w+)\s*=\s*([^;]+);", spec_const, text)
    if re.search(r"\blayout\s*\(", text):
        raise RuntimeError(f"{path}: untranslated layout declaration: " + re.search(r"\blayout\s*\([^;]*;", text).group(0))

    text = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1 &\2", text)
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", text)
    text = re.sub(r"\bdiscard\s*;", "{ gl_Discarded = true; return; }", text)
    text = re.sub(r"\bignoreIntersectionEXT\s*;", "{ gl_IgnoreIntersection = true; return; }", text)
    # a local initialised from the uniform it shadows
    for name in meta["uniform_names"]:
        text = re.sub(r"(\b\w+\s+%s\s*=\s*)([^;]*\b%s\b[^;]*;)" % (name, name),
                      lambda m: m.group(1) + re.sub(r"\b%s\b" % name, f"::{ns}::{name}", m.group(2)), text)
    return text, meta


PRELUDE = """// GENERATED by oracle/make_ref.py from {src} — do not edit, do not commit.
#include "ref_shim.h"
namespace {ns} {{
using namespace glsl;
thread_local uvec3_xy gl_GlobalInvocationID, gl_LaunchIDEXT, gl_LaunchSizeEXT;
thread_local int gl_GeometryIndexEXT, gl_PrimitiveID, gl_VertexIndex;
thread_local vec4 gl_FragCoord, gl_Position;
thread_local bool gl_Discarded, gl_IgnoreIntersection;
inline float max(float a, float b) {{ return gpu_max(a, b); }}      // GPU corner-case semantics, see ref_shim.h
inline float min(float a, float b) {{ return gpu_min(a, b); }}
inline float pow(float x, float y) {{ return gpu_pow(x, y); }}
{fwd}
// ---- shader text ------------------------------------------------------------------------------------------------------------
{body}
// ---- end of shader text -----------------------------------------------------------------------------------------------------
}}  // namespace {ns}
"""

TRACE_FWD = "void traceRayEXT(accelerationStructureEXT &as, uint flags, uint cull_mask, uint sbt_offset, uint sbt_stride, uint miss_index, vec3 origin, float tmin, vec3 dir, float tmax, int payload_loc);"


def shader_unit(rel, ns, fwd=""):
    path = os.path.join(SHADER_DIR, rel)
    body, meta = translate(path, ns)
    return PRELUDE.format(src=path, ns=ns, fwd=fwd, body=body), meta


# ---------------------------------------------------------------------------------------------------------------------------
# Harnesses: what the Vulkan driver + the reference's host code do around a shader (bind, dispatch)
# ---------------------------------------------------------------------------------------------------------------------------
COMMON_EXPORTS = """
extern "C" {
// data/shaders/common.glsl helpers and src/rendering_backend/glsl_common.h struct sizes, straight from the reference text
uint32_t vr_seed_thread(uint32_t s) { return ref_svgf::seed_thread(s); }
uint32_t vr_random(uint32_t *state) { return ref_svgf::random(*state); }
float vr_random01(uint32_t *state) { return ref_svgf::random01(*state); }
float vr_random01_inclusive(uint32_t *state) { return ref_svgf::random01_inclusive(*state); }
uint32_t vr_random_range(uint32_t *state, uint32_t lo, uint32_t hi) { return ref_svgf::random(*state, lo, hi); }
void vr_uniform_sample_cone(float u0, float u1, float c, float *o) { glm::vec3 r = ref_svgf::uniform_sample_cone(glm::vec2(u0, u1), c); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void vr_cosine_hemisphere(float u0, float u1, float *o) { glm::vec3 r = ref_svgf::uniform_sample_cosine_weighted_hemisphere(glm::vec2(u0, u1)); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void vr_onb(const float *n, float *o) { glm::mat3 M = ref_svgf::onb_from_unit_vector(glm::vec3(n[0], n[1], n[2])); for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) o[3 * c + r] = M[c][r]; }
void vr_oct_encode(const float *v, float *o) { glm::vec2 r = ref_svgf::vec3_encode_to_oct(glm::vec3(v[0], v[1], v[2])); o[0] = r.x; o[1] = r.y; }
void vr_oct_decode(const float *e, float *o) { glm::vec3 r = ref_svgf::oct_decode_to_vec3(glm::vec2(e[0], e[1])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void vr_fresnel_schlick(const float *f0, const float *H, const float *V, float *o) { glm::vec3 r = ref_svgf::fresnel_schlick(glm::vec3(f0[0], f0[1], f0[2]), glm::vec3(H[0], H[1], H[2]), glm::vec3(V[0], V[1], V[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
float vr_D_GGX(float rough, const float *N, const float *H) { return ref_svgf::D_GGX(rough, glm::vec3(N[0], N[1], N[2]), glm::vec3(H[0], H[1], H[2])); }
float vr_G_GGX(float rough, const float *N, const float *V, const float *L) { return ref_svgf::G_GGX(rough, glm::vec3(N[0], N[1], N[2]), glm::vec3(V[0], V[1], V[2]), glm::vec3(L[0], L[1], L[2])); }
void vr_get_world_space_position(const void *pfd, float depth, float u, float v, float *o) {
    std::memcpy(&ref_svgf::pfd, pfd, sizeof(ref_svgf::pfd));
    glm::vec3 r = ref_svgf::get_world_space_position(depth, glm::vec2(u, v)); o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void vr_get_view_space_position(const void *pfd, float depth, float u, float v, float *o) {
    std::memcpy(&ref_svgf::pfd, pfd, sizeof(ref_svgf::pfd));
    glm::vec3 r = ref_svgf::get_view_space_position(depth, glm::vec2(u, v)); o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
int vr_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(ref_svgf::PerFrameData);
        case 1: return (int)sizeof(ref_svgf::Vertex);
        case 2: return (int)sizeof(ref_svgf::Material);
        case 3: return (int)sizeof(ref_svgf::Primitive);
        case 4: return (int)sizeof(ref_svgf::SVGFPushConstants);
        case 5: return (int)sizeof(ref_svgf::SSRPushConstants);
        case 6: return (int)sizeof(ref_svgf::SSAOPushConstants);
        case 7: return (int)sizeof(ref_svgf::DirectionalLight);
        case 8: return (int)sizeof(ref_svgf::HybridPushConstants);
    }
    return -1;
}
uint16_t vr_f2h(float f) { return glsl::float_to_half(f); }
float vr_h2f(uint16_t h) { return glsl::half_to_float(h); }
// texture(sampler, uv) on an RGBA8 texture, for the sampler unit tests
void vr_sample_rgba8(const uint8_t *rgba, int w, int h, int vk_format, int mag, int min, int wrap_u, int wrap_v, float u, float v, float *o) {
    glsl::ImageDesc d{rgba, nullptr, w, h, vk_format};
    glsl::sampler2D s; s.d = &d; s.mag = mag; s.min = min; s.wrap_u = wrap_u; s.wrap_v = wrap_v;
    glm::vec4 r = glsl::texture(s, glm::vec2(u, v)); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
}
"""

# vkCmdDispatch of svgf.comp as HybridRenderPath records it (src/render_paths/hybrid_render_path.cpp:288-300): set 3 = normals,
# motion, depth (unused by the shader), raw ray-traced image, denoised (unused); storage_images[] by the push-constant indices.
# The shader reads the moments image at neighbouring pixels while other invocations overwrite it: the dispatch
# reads the previous frame's moments (`moments_in`) and writes `moments_out`.
HARNESS_SVGF = """
namespace ref_svgf {
static void bind(const void *pfd_, ImageDesc *st, ImageDesc *tr) {
    std::memcpy(&pfd, pfd_, sizeof(pfd));
    static thread_local image2D slots[5];
    for (int i = 0; i < 5; ++i) slots[i].d = &st[i];
    storage_images = slots;
    world_space_normals_and_object_ids.d = &tr[0];
    motion_vectors_and_metallic_roughness.d = &tr[1];
    raytraced_shadow_and_ao_texture.d = &tr[2];
    pc.integrated_shadow_and_ao = ivec2(0, 1);
    pc.prev_frame_normals_and_object_ids = 2;
    pc.shadow_and_ao_history = 3;
    pc.shadow_and_ao_moments_history = 4;
    pc.atrous_step = 1;
}
}
extern "C" void vr_svgf_temporal(const void *pfd, int W, int H, const uint16_t *normals, const uint16_t *motion, const uint16_t *rt, const uint16_t *prev_normals,
                                 const uint16_t *history, const uint16_t *moments_in, uint16_t *integrated_out, uint16_t *moments_out) {
    using namespace glsl;
    ImageDesc st[5] = {{integrated_out, integrated_out, W, H, FMT_R16G16B16A16_SFLOAT}, {nullptr, nullptr, W, H, FMT_R16G16B16A16_SFLOAT},
                       {prev_normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, {history, nullptr, W, H, FMT_R16G16B16A16_SFLOAT},
                       {moments_in, moments_out, W, H, FMT_R16G16_SFLOAT}};
    ImageDesc tr[3] = {{normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, {motion, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, {rt, nullptr, W, H, FMT_R16G16_SFLOAT}};
#pragma omp parallel
    {
        ref_svgf::bind(pfd, st, tr);
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                ref_svgf::gl_GlobalInvocationID.set(x, y);
                ref_svgf::shader_main();
            }
    }
}
"""

HARNESS_ATROUS = """
namespace ref_atrous {
static void bind(const void *pfd_, ImageDesc *st, ImageDesc *tr, int step) {
    std::memcpy(&pfd, pfd_, sizeof(pfd));
    static thread_local image2D slots[2];
    for (int i = 0; i < 2; ++i) slots[i].d = &st[i];
    storage_images = slots;
    world_space_normals_and_object_ids.d = &tr[0];
    pc.integrated_shadow_and_ao = ivec2(0, 1);
    pc.atrous_step = step;
}
}
extern "C" void vr_svgf_atrous(const void *pfd, int W, int H, int step, const uint16_t *normals, const uint16_t *integ_in, uint16_t *integ_out) {
    using namespace glsl;
    ImageDesc st[2] = {{integ_in, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, {integ_out, integ_out, W, H, FMT_R16G16B16A16_SFLOAT}};
    ImageDesc tr[1] = {{normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}};
#pragma omp parallel
    {
        ref_atrous::bind(pfd, st, tr, step);
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                ref_atrous::gl_GlobalInvocationID.set(x, y);
                ref_atrous::shader_main();
            }
    }
}
"""

# The "SVGF Denoise Pass" callback (hybrid_render_path.cpp:288-330) restated around the two compiled shaders: the dispatch
# order, the three blits and the ping-pong swaps. state = five persistent images (zero-initialised: documented deviation, the
# reference leaves them undefined).
HARNESS_SVGF_PASS = """
struct vr_svgf_state { int W, H; std::vector<uint16_t> img[5]; int ping[2]; };
extern "C" vr_svgf_state *vr_svgf_state_create(int W, int H) {
    vr_svgf_state *s = new vr_svgf_state();
    s->W = W; s->H = H;
    for (int i = 0; i < 5; ++i) s->img[i].assign((size_t)W * H * (i == 4 ? 2 : 4), 0);
    s->ping[0] = 0; s->ping[1] = 1;
    return s;
}
extern "C" void vr_svgf_state_destroy(vr_svgf_state *s) { delete s; }
extern "C" uint16_t *vr_svgf_state_image(vr_svgf_state *s, int which) { return s->img[which].data(); }
// out_iters (optional): the five a-trous outputs; out_temporal (optional): integrated[0] after svgf.comp
extern "C" void vr_svgf_pass(vr_svgf_state *s, const void *pfd, const uint16_t *normals, const uint16_t *motion, const uint16_t *rt, uint16_t *denoised,
                             uint16_t *out_iters, uint16_t *out_temporal) {
    const int W = s->W, H = s->H;
    const size_t n4 = (size_t)W * H * 4;
    int x = s->ping[0], y = s->ping[1];                   // pc.integrated_shadow_and_ao.{x,y}: slots 0 / 1
    std::vector<uint16_t> moments_out(s->img[4].size());
    vr_svgf_temporal(pfd, W, H, normals, motion, rt, s->img[2].data(), s->img[3].data(), s->img[4].data(), s->img[x].data(), moments_out.data());   // :299
    s->img[4].swap(moments_out);
    if (out_temporal) std::memcpy(out_temporal, s->img[x].data(), n4 * 2);
    for (int i = 0; i < 5; ++i) {                         // :300-319
        vr_svgf_atrous(pfd, W, H, 1 << i, normals, s->img[x].data(), s->img[y].data());
        if (out_iters) std::memcpy(out_iters + (size_t)i * n4, s->img[y].data(), n4 * 2);
        if (i == 0) s->img[3] = s->img[y];                // :309-315 BlitImageStorageToStorage(integrated.y -> history)
        std::swap(x, y);                                  // :318
    }
    s->img[2].assign(normals, normals + n4);              // :321 BlitImageTransientToStorage(normals -> prev normals)
    if (denoised) std::memcpy(denoised, s->img[y].data(), n4 * 2);   // :322-325 BlitImageStorageToTransient(integrated.y -> Denoised)
    std::swap(x, y);                                      // :328
    s->ping[0] = x; s->ping[1] = y;
}
"""

HARNESS_SSAO = """
extern "C" void vr_ssao(const void *pfd, int W, int H, float radius, const float *depth, const uint16_t *normals, uint16_t *out) {
    using namespace glsl;
    ImageDesc dn = {normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, dd = {depth, nullptr, W, H, FMT_D32_SFLOAT}, dout = {out, out, W, H, FMT_R16G16B16A16_SFLOAT};
#pragma omp parallel
    {
        std::memcpy(&ref_ssao::pfd, pfd, sizeof(ref_ssao::pfd));
        ref_ssao::world_space_normals_and_object_ids.d = &dn;       // default sampler: LINEAR, REPEAT (resource_manager.cpp:58-69)
        ref_ssao::depth.d = &dd;
        ref_ssao::screen_space_ambient_occlusion.d = &dout;
        ref_ssao::pc.radius = radius;
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                ref_ssao::gl_GlobalInvocationID.set(x, y);
                ref_ssao::shader_main();
            }
    }
}
extern "C" void vr_ssao_blur(const void *pfd, int W, int H, const uint16_t *raw, uint16_t *out) {
    using namespace glsl;
    ImageDesc din = {raw, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, dout = {out, out, W, H, FMT_R16G16B16A16_SFLOAT};
#pragma omp parallel
    {
        std::memcpy(&ref_ssao_blur::pfd, pfd, sizeof(ref_ssao_blur::pfd));
        ref_ssao_blur::screen_space_ambient_occlusion.d = &din;
        ref_ssao_blur::screen_space_ambient_occlusion_blurred.d = &dout;
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                ref_ssao_blur::gl_GlobalInvocationID.set(x, y);
                ref_ssao_blur::shader_main();
            }
    }
}
"""

HARNESS_SSR = """
extern "C" void vr_ssr(const void *pfd, int W, int H, int y0, int y1, float ray_distance, float step_size, float thickness, int bsearch_steps,
                       const uint8_t *albedo, const uint16_t *normals, const uint16_t *motion, const float *depth, uint16_t *out) {
    using namespace glsl;
    ImageDesc da = {albedo, nullptr, W, H, FMT_B8G8R8A8_UNORM}, dn = {normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, dm = {motion, nullptr, W, H, FMT_R16G16B16A16_SFLOAT},
              dd = {depth, nullptr, W, H, FMT_D32_SFLOAT}, dout = {out, out, W, H, FMT_R16G16B16A16_SFLOAT};
#pragma omp parallel
    {
        std::memcpy(&ref_ssr::pfd, pfd, sizeof(ref_ssr::pfd));
        ref_ssr::albedo.d = &da; ref_ssr::world_space_normals_and_object_ids.d = &dn; ref_ssr::motion_vectors_and_metallic_roughness.d = &dm;
        ref_ssr::depth.d = &dd; ref_ssr::screen_space_reflections.d = &dout;
        ref_ssr::pc.ray_distance = ray_distance; ref_ssr::pc.step_size = step_size; ref_ssr::pc.thickness = thickness; ref_ssr::pc.bsearch_steps = bsearch_steps;
#pragma omp for schedule(dynamic, 1)
        for (int y = y0; y < y1; ++y)
            for (int x = 0; x < W; ++x) {
                ref_ssr::gl_GlobalInvocationID.set(x, y);
                ref_ssr::shader_main();
            }
    }
}
"""

# vkCmdDraw(3, 1, 0, 0) of the composition pipeline (hybrid_render_path.cpp:333-379): composition.vert's full-screen triangle gives
# every fragment in_uv = pixel centre / size (the rasteriser's interpolation is fixed function); the attachment store is the
# format conversion of `out_format` (50 = B8G8R8A8_SRGB swapchain, 44 = UNORM, 97 = linear fp16 with NaN -> 0 for measurements).
HARNESS_COMPOSITION = """
extern "C" void vr_composition(const void *pfd, int W, int H, int shadow_mode, int ao_mode, int reflection_mode, const uint8_t *albedo, const uint16_t *normals,
                               const uint16_t *motion, const float *depth, const float *shadow_map, int shadow_w, int shadow_h, const uint16_t *ssao,
                               const uint16_t *ssr, const uint16_t *rt, int rt_channels, const uint16_t *refl, int out_format, void *out) {
    using namespace glsl;
    ImageDesc da = {albedo, nullptr, W, H, FMT_B8G8R8A8_UNORM}, dn = {normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, dm = {motion, nullptr, W, H, FMT_R16G16B16A16_SFLOAT},
              dd = {depth, nullptr, W, H, FMT_D32_SFLOAT}, dsm = {shadow_map, nullptr, shadow_w, shadow_h, FMT_D32_SFLOAT},
              dssao = {ssao, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, dssr = {ssr, nullptr, W, H, FMT_R16G16B16A16_SFLOAT},
              drt = {rt, nullptr, W, H, rt_channels == 2 ? FMT_R16G16_SFLOAT : FMT_R16G16B16A16_SFLOAT}, drefl = {refl, nullptr, W, H, FMT_R16G16B16A16_SFLOAT};
#pragma omp parallel
    {
        namespace S = ref_composition;
        std::memcpy(&S::pfd, pfd, sizeof(S::pfd));
        S::albedo.d = &da; S::world_space_normals_and_object_ids.d = &dn; S::motion_vectors_and_metallic_roughness.d = &dm; S::depth.d = &dd;
        S::shadow_map.d = &dsm; S::screen_space_ambient_occlusion.d = &dssao; S::screen_space_reflections.d = &dssr;
        S::raytraced_shadow_and_ao_texture.d = &drt; S::raytraced_reflections_texture.d = &drefl;
        S::shadow_mode = shadow_mode; S::ambient_occlusion_mode = ao_mode; S::reflection_mode = reflection_mode;
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                S::in_uv = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
                S::gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
                S::shader_main();
                const vec4 c = S::out_color;
                const size_t pix = (size_t)y * W + x;
                if (out_format == 97) {
                    auto nz = [](float v) { return v == v ? v : 0.0f; };
                    uint16_t *o = (uint16_t *)out + 4 * pix;
                    o[0] = float_to_half(nz(c.x)); o[1] = float_to_half(nz(c.y)); o[2] = float_to_half(nz(c.z)); o[3] = float_to_half(c.w);
                } else {
                    uint8_t *o = (uint8_t *)out + 4 * pix;
                    const bool srgb = out_format == 50;
                    o[0] = unorm8(srgb ? linear_to_srgb(c.z) : c.z); o[1] = unorm8(srgb ? linear_to_srgb(c.y) : c.y); o[2] = unorm8(srgb ? linear_to_srgb(c.x) : c.x);
                    o[3] = unorm8(c.w);
                }
            }
    }
}
"""

# vkCmdTraceRaysKHR of the "Raytrace Pipeline" (hybrid_render_path.cpp:101-136): raygen.rgen, miss.rmiss (index 0),
# reflection_miss.rmiss (index 1), one hit group with reflection_hit.rchit. traceRayEXT = the driver: query the acceleration
# structure (bound to the oracle's BVH), then run the reference's miss / closest-hit shader with the payload of `payload_loc`.
HARNESS_RAYGEN = """
namespace ref_raygen {
void traceRayEXT(accelerationStructureEXT &as, uint flags, uint cull_mask, uint sbt_offset, uint sbt_stride, uint miss_index, vec3 origin, float tmin, vec3 dir,
                 float tmax, int payload_loc) {
    vec4 &caller_payload = payload_loc == 0 ? payload : reflection_payload;
    const RayHit h = trace_query(as, flags, origin, tmin, dir, tmax);
    if (!h.hit) {
        if (miss_index == 0) { ref_miss::payload = caller_payload; ref_miss::shader_main(); caller_payload = ref_miss::payload; }
        else { ref_reflection_miss::reflection_payload = caller_payload; ref_reflection_miss::shader_main(); caller_payload = ref_reflection_miss::reflection_payload; }
        return;
    }
    if (flags & gl_RayFlagsSkipClosestHitShaderEXT) return;
    namespace Hs = ref_reflection_hit;
    Hs::gl_GeometryIndexEXT = h.geometry_index; Hs::gl_PrimitiveID = h.primitive_id;
    Hs::hit_attribs = vec3(h.attribs, 0.0f);
    Hs::reflection_payload = caller_payload;
    Hs::shader_main();
    caller_payload = Hs::reflection_payload;
}
}
struct vr_texture { const uint8_t *rgba; int w, h, vk_format, mag, min, wrap_u, wrap_v; };
extern "C" void vr_raygen(const void *scene, void *trace_any, void *trace_closest, const void *pfd, int W, int H, int y0, int y1, const float *depth,
                          const uint16_t *normals, uint16_t *shadow_ao_rg16f, uint16_t *reflections, const void *vertices, const uint32_t *indices,
                          const void *primitives, const vr_texture *textures, int n_textures) {
    using namespace glsl;
    ImageDesc dn = {normals, nullptr, W, H, FMT_R16G16B16A16_SFLOAT}, dd = {depth, nullptr, W, H, FMT_D32_SFLOAT},
              dsa = {shadow_ao_rg16f, shadow_ao_rg16f, W, H, FMT_R16G16_SFLOAT},      // hybrid_render_path.cpp:109: the image is R16G16_SFLOAT
              drefl = {reflections, reflections, W, H, FMT_R16G16B16A16_SFLOAT};
    std::vector<ImageDesc> tdesc(n_textures);
    std::vector<sampler2D> tsamp(n_textures);
    for (int i = 0; i < n_textures; ++i) {
        tdesc[i] = ImageDesc{textures[i].rgba, nullptr, textures[i].w, textures[i].h, textures[i].vk_format};
        tsamp[i].d = &tdesc[i]; tsamp[i].mag = textures[i].mag; tsamp[i].min = textures[i].min; tsamp[i].wrap_u = textures[i].wrap_u; tsamp[i].wrap_v = textures[i].wrap_v;
    }
#pragma omp parallel
    {
        namespace R = ref_raygen;
        namespace Hs = ref_reflection_hit;
        std::memcpy(&R::pfd, pfd, sizeof(R::pfd));
        std::memcpy(&Hs::pfd, pfd, sizeof(Hs::pfd));
        R::TLAS.scene = scene;
        R::TLAS.trace_any = (int (*)(const void *, const float *, const float *, float, float))trace_any;
        R::TLAS.trace_closest = (int (*)(const void *, const float *, const float *, float, float, double *, uint32_t *))trace_closest;
        R::world_space_normals_and_object_ids.d = &dn; R::depth.d = &dd;
        R::raytraced_shadow_and_ambient_occlusion.d = &dsa; R::raytraced_reflections.d = &drefl;
        Hs::vertices = (const Hs::Vertex *)vertices; Hs::indices = indices; Hs::primitives = (const Hs::Primitive *)primitives;
        Hs::textures = tsamp.data();
        R::gl_LaunchSizeEXT.set(W, H, 1);
#pragma omp for schedule(dynamic, 1)
        for (int y = y0; y < y1; ++y)
            for (int x = 0; x < W; ++x) {
                R::gl_LaunchIDEXT.set(x, y, 0);
                R::shader_main();
            }
    }
}
"""


# vkCmdTraceRaysKHR of the fully ray-traced path's "Raytracing Pipeline" (src/render_paths/raytraced_render_path.cpp:12-47): raygen.rgen /
# raygen_test_alpha.rgen, miss.rmiss (index 0), shadow_miss.rmiss (index 1), one hit group: closesthit.rchit, or closesthit_test_alpha.rchit +
# shadow_anyhit.rahit. The closest-hit shader itself calls traceRayEXT (the shadow ray, payload location 1 = a bool).
def harness_raytraced(sfx, alpha):
    anyhit = """
static int run_anyhit(void *, uint32_t geometry_index, uint32_t primitive_id, double u, double v) {
    namespace A = ref_rtp%(sfx)s_ahit;
    A::gl_GeometryIndexEXT = (int)geometry_index; A::gl_PrimitiveID = (int)primitive_id;
    A::hit_attribs = glm::vec3((float)u, (float)v, 0.0f);
    A::gl_IgnoreIntersection = false;
    A::shader_main();
    return A::gl_IgnoreIntersection ? 0 : 1;
}
""" % dict(sfx=sfx) if alpha else ""
    ah = "run_anyhit" if alpha else "nullptr"
    return anyhit + """
namespace ref_rtp%(sfx)s_chit {
void traceRayEXT(accelerationStructureEXT &as, uint flags, uint cull_mask, uint sbt_offset, uint sbt_stride, uint miss_index, vec3 origin, float tmin, vec3 dir,
                 float tmax, int payload_loc) {      // the shadow ray: payload location 1, miss index 1, closest-hit shader skipped
    const RayHit h = trace_query(as, flags, origin, tmin, dir, tmax, (flags & gl_RayFlagsNoOpaqueEXT) ? %(ah)s : nullptr);
    if (!h.hit) { ref_rtp%(sfx)s_smiss::payload = shadow_payload; ref_rtp%(sfx)s_smiss::shader_main(); shadow_payload = ref_rtp%(sfx)s_smiss::payload; }
}
}
namespace ref_rtp%(sfx)s_raygen {
void traceRayEXT(accelerationStructureEXT &as, uint flags, uint cull_mask, uint sbt_offset, uint sbt_stride, uint miss_index, vec3 origin, float tmin, vec3 dir,
                 float tmax, int payload_loc) {
    const RayHit h = trace_query(as, flags, origin, tmin, dir, tmax, (flags & gl_RayFlagsNoOpaqueEXT) ? %(ah)s : nullptr);
    if (!h.hit) { ref_rtp%(sfx)s_miss::payload = payload; ref_rtp%(sfx)s_miss::shader_main(); payload = ref_rtp%(sfx)s_miss::payload; return; }
    namespace Hs = ref_rtp%(sfx)s_chit;
    Hs::gl_GeometryIndexEXT = h.geometry_index; Hs::gl_PrimitiveID = h.primitive_id;
    Hs::hit_attribs = vec3(h.attribs, 0.0f);
    Hs::payload = payload;
    Hs::shader_main();
    payload = Hs::payload;
}
}
extern "C" void vr_raytraced%(sfx)s(const void *scene, void *trace_any, void *trace_closest, void *trace_closest_filtered, const void *pfd, int W, int H, uint8_t *out_bgra8,
                               const void *vertices, const uint32_t *indices, const void *primitives, const vr_texture *textures, int n_textures) {
    using namespace glsl;
    ImageDesc dout = {out_bgra8, out_bgra8, W, H, FMT_B8G8R8A8_UNORM};     // "RaytracedOutput" (raytraced_render_path.cpp:15; the shader says rgba8)
    // textures[]: slot -1 (a material without a base-colour texture; the alpha-tested shaders index it unconditionally, which is undefined in the
    // reference) holds one opaque white texel
    static const uint8_t white[4] = {255, 255, 255, 255};
    std::vector<ImageDesc> tdesc(n_textures + 1);
    std::vector<sampler2D> tsamp(n_textures + 1);
    tdesc[0] = ImageDesc{white, nullptr, 1, 1, FMT_R8G8B8A8_UNORM}; tsamp[0].d = &tdesc[0];
    for (int i = 0; i < n_textures; ++i) {
        tdesc[i + 1] = ImageDesc{textures[i].rgba, nullptr, textures[i].w, textures[i].h, textures[i].vk_format};
        tsamp[i + 1].d = &tdesc[i + 1]; tsamp[i + 1].mag = textures[i].mag; tsamp[i + 1].min = textures[i].min; tsamp[i + 1].wrap_u = textures[i].wrap_u; tsamp[i + 1].wrap_v = textures[i].wrap_v;
    }
#pragma omp parallel
    {
        namespace R = ref_rtp%(sfx)s_raygen;
        namespace Hs = ref_rtp%(sfx)s_chit;
        accelerationStructureEXT as;
        as.scene = scene;
        as.trace_any = (int (*)(const void *, const float *, const float *, float, float))trace_any;
        as.trace_closest = (int (*)(const void *, const float *, const float *, float, float, double *, uint32_t *))trace_closest;
        as.trace_closest_filtered = (int (*)(const void *, const float *, const float *, float, float, int (*)(void *, uint32_t, uint32_t, double, double), void *, double *, uint32_t *))trace_closest_filtered;
        std::memcpy(&R::pfd, pfd, sizeof(R::pfd)); std::memcpy(&Hs::pfd, pfd, sizeof(Hs::pfd));
        R::TLAS = as; Hs::TLAS = as;
        R::output_image.d = &dout;
        Hs::vertices = (const Hs::Vertex *)vertices; Hs::indices = indices; Hs::primitives = (const Hs::Primitive *)primitives; Hs::textures = tsamp.data() + 1;
%(ahbind)s
        R::gl_LaunchSizeEXT.set(W, H, 1);
#pragma omp for schedule(dynamic, 1)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                R::gl_LaunchIDEXT.set(x, y, 0);
                R::shader_main();
            }
    }
}
""" % dict(sfx=sfx, ah=ah, ahbind=("        { namespace A = ref_rtp%s_ahit; A::vertices = (const A::Vertex *)vertices; A::indices = indices; A::primitives = (const A::Primitive *)primitives; A::textures = tsamp.data() + 1; }" % sfx) if alpha else "")


def generate():
    """The translation units as (file name, C++ text); nothing is written here."""
    units = []
    H = "hybrid_render_path/"
    svgf, _ = shader_unit(H + "svgf.comp", "ref_svgf")
    atrous, _ = shader_unit(H + "svgf_atrous_filter.comp", "ref_atrous")
    units.append(("ref_svgf.cpp", "#include <vector>\n" + svgf + atrous + HARNESS_SVGF + HARNESS_ATROUS + HARNESS_SVGF_PASS + COMMON_EXPORTS))
    ssao, _ = shader_unit(H + "ssao.comp", "ref_ssao")
    blur, _ = shader_unit(H + "ssao_blur.comp", "ref_ssao_blur")
    units.append(("ref_ssao.cpp", ssao + blur + HARNESS_SSAO))
    ssr, _ = shader_unit(H + "ssr.comp", "ref_ssr")
    units.append(("ref_ssr.cpp", ssr + HARNESS_SSR))
    comp, _ = shader_unit(H + "composition.frag", "ref_composition")
    units.append(("ref_composition.cpp", comp + HARNESS_COMPOSITION))
    miss, _ = shader_unit(H + "miss.rmiss", "ref_miss")
    rmiss, _ = shader_unit(H + "reflection_miss.rmiss", "ref_reflection_miss")
    rhit, _ = shader_unit(H + "reflection_hit.rchit", "ref_reflection_hit", fwd=TRACE_FWD)
    rgen, _ = shader_unit(H + "raygen.rgen", "ref_raygen", fwd=TRACE_FWD)
    units.append(("ref_raygen.cpp", "#include <vector>\n" + miss + rmiss + rhit + rgen + HARNESS_RAYGEN))
    P = "raytraced_render_path/"
    tex_struct = "#include <cstdint>\nstruct vr_texture { const uint8_t *rgba; int w, h, vk_format, mag, min, wrap_u, wrap_v; };\n"
    for sfx, alpha in (("", False), ("_alpha", True)):
        parts = [shader_unit(P + "miss.rmiss", f"ref_rtp{sfx}_miss")[0], shader_unit(P + "shadow_miss.rmiss", f"ref_rtp{sfx}_smiss")[0]]
        if alpha:
            parts.append(shader_unit(P + "shadow_anyhit.rahit", f"ref_rtp{sfx}_ahit")[0])
        parts.append(shader_unit(P + ("closesthit_test_alpha.rchit" if alpha else "closesthit.rchit"), f"ref_rtp{sfx}_chit", fwd=TRACE_FWD)[0])
        parts.append(shader_unit(P + ("raygen_test_alpha.rgen" if alpha else "raygen.rgen"), f"ref_rtp{sfx}_raygen", fwd=TRACE_FWD)[0])
        units.append((f"ref_raytraced{sfx}.cpp", "#include <vector>\n" + tex_struct + "".join(parts) + harness_raytraced(sfx, alpha)))
    return units


def build(force=False, verbose=False, keep_generated=False):
    """Builds oracle/_ref/libvhr_ref.so when /root/reference is present; returns its path, or None when neither the reference nor a
    prebuilt library exists (the GPU box only ever sees the prebuilt one). The generated C++ (which carries the reference's shader text) only
    exists in oracle/_ref/gen/ for the duration of the compile: what stays is the library and a hash of its inputs (gen.stamp)."""
    import hashlib
    import shutil
    if not reference_available():
        return LIB if os.path.exists(LIB) else None
    units = generate()
    h = hashlib.sha256()
    for name, text in units:
        h.update(name.encode()); h.update(text.encode())
    for dep in (os.path.join(HERE, "ref_shim.h"), os.path.abspath(__file__)):
        with open(dep, "rb") as f:
            h.update(f.read())
    h.update(" ".join(CXXFLAGS).encode())
    stamp, digest = os.path.join(OUT_DIR, "gen.stamp"), h.hexdigest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    os.makedirs(GEN_DIR, exist_ok=True)
    objs, procs = [], []
    try:
        for name, text in units:
            s = os.path.join(GEN_DIR, name)
            with open(s, "w") as f:
                f.write(text)
            o = s[:-4] + ".o"
            objs.append(o)
            procs.append((s, subprocess.Popen([CXX] + CXXFLAGS + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        failed = False
        for s, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                failed = True
                sys.stderr.write(f"--- {s}\n{out}\n")
            elif verbose and out:
                print(out)
        if failed:
            raise RuntimeError("oracle/_ref: compiling the reference shaders failed")
        subprocess.run([CXX, "-shared", "-fopenmp", "-o", LIB] + objs, check=True)
        with open(stamp, "w") as f:
            f.write(digest + "\n")
    finally:
        if not keep_generated:
            shutil.rmtree(GEN_DIR, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, keep_generated="--keep" in sys.argv))
