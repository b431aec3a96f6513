"""ctypes binding of oracle/_ref/libvhr_ref.so — the REFERENCE'S OWN GLSL SHADERS compiled for the CPU by oracle/make_ref.py.
TEST INFRASTRUCTURE ONLY: tests/ use it to pin the hand-written oracle (oracle/*.cpp) and, through it, the CUDA kernels.

Same call shapes as oracle_lib so a test can run both on the same arrays. The library is built where /root/reference exists (this
container); on the GPU box the prebuilt .so that travelled with the snapshot is loaded. available() says whether there is one.
"""
import ctypes as C
import os

import numpy as np

import make_ref
import oracle_lib as O
from vulkanhybridrenderer_b200 import types as T

_lib = None


def available():
    return make_ref.reference_available() or os.path.exists(make_ref.LIB)


def lib():
    global _lib
    if _lib is None:
        path = make_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libvhr_ref.so is absent and /root/reference is not here to build it from")
        _lib = C.CDLL(path)
        _declare(_lib)
    return _lib


_p = O._p
_h = O._h


def _declare(L):
    vp = C.c_void_p
    L.vr_seed_thread.restype = C.c_uint32
    L.vr_seed_thread.argtypes = [C.c_uint32]
    L.vr_random.restype = C.c_uint32
    L.vr_random.argtypes = [C.POINTER(C.c_uint32)]
    L.vr_random01.restype = C.c_float
    L.vr_random01.argtypes = [C.POINTER(C.c_uint32)]
    L.vr_random01_inclusive.restype = C.c_float
    L.vr_random01_inclusive.argtypes = [C.POINTER(C.c_uint32)]
    L.vr_random_range.restype = C.c_uint32
    L.vr_random_range.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    L.vr_uniform_sample_cone.argtypes = [C.c_float, C.c_float, C.c_float, vp]
    L.vr_cosine_hemisphere.argtypes = [C.c_float, C.c_float, vp]
    L.vr_onb.argtypes = [vp, vp]
    L.vr_oct_encode.argtypes = [vp, vp]
    L.vr_oct_decode.argtypes = [vp, vp]
    L.vr_fresnel_schlick.argtypes = [vp, vp, vp, vp]
    L.vr_D_GGX.restype = C.c_float
    L.vr_D_GGX.argtypes = [C.c_float, vp, vp]
    L.vr_G_GGX.restype = C.c_float
    L.vr_G_GGX.argtypes = [C.c_float, vp, vp, vp]
    L.vr_get_world_space_position.argtypes = [vp, C.c_float, C.c_float, C.c_float, vp]
    L.vr_get_view_space_position.argtypes = [vp, C.c_float, C.c_float, C.c_float, vp]
    L.vr_sizeof.restype = C.c_int
    L.vr_sizeof.argtypes = [C.c_int]
    L.vr_f2h.restype = C.c_uint16
    L.vr_f2h.argtypes = [C.c_float]
    L.vr_h2f.restype = C.c_float
    L.vr_h2f.argtypes = [C.c_uint16]
    L.vr_sample_rgba8.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp]
    L.vr_svgf_temporal.argtypes = [vp, C.c_int, C.c_int] + [vp] * 8
    L.vr_svgf_atrous.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.vr_svgf_state_create.restype = vp
    L.vr_svgf_state_create.argtypes = [C.c_int, C.c_int]
    L.vr_svgf_state_destroy.argtypes = [vp]
    L.vr_svgf_state_image.restype = C.POINTER(C.c_uint16)
    L.vr_svgf_state_image.argtypes = [vp, C.c_int]
    L.vr_svgf_pass.argtypes = [vp] * 8
    L.vr_ssao.argtypes = [vp, C.c_int, C.c_int, C.c_float, vp, vp, vp]
    L.vr_ssao_blur.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.vr_ssr.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, vp, vp, vp, vp, vp]
    L.vr_composition.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, vp]
    L.vr_raygen.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    for name in ("vr_raytraced", "vr_raytraced_alpha"):
        getattr(L, name).argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int]


def svgf_temporal(pfd, normals, motion, rt, prev_normals, history, moments_in):
    H, W = normals.shape[:2]
    integ = np.zeros((H, W, 4), np.float16)
    mom = np.zeros((H, W, 2), np.float16)
    lib().vr_svgf_temporal(_p(pfd), W, H, _p(_h(normals)), _p(_h(motion)), _p(_h(rt)), _p(_h(prev_normals)), _p(_h(history)), _p(_h(moments_in)),
                           _p(integ), _p(mom))
    return integ, mom


def svgf_atrous(pfd, normals, integ_in, step):
    H, W = normals.shape[:2]
    out = np.zeros((H, W, 4), np.float16)
    lib().vr_svgf_atrous(_p(pfd), W, H, int(step), _p(_h(normals)), _p(_h(integ_in)), _p(out))
    return out


class SvgfState:
    """The "SVGF Denoise Pass" callback of hybrid_render_path.cpp:288-330 around the two compiled reference shaders."""

    def __init__(self, W, H):
        self.W, self.H = W, H
        self._s = lib().vr_svgf_state_create(W, H)

    def __del__(self):
        if getattr(self, "_s", None):
            lib().vr_svgf_state_destroy(self._s)
            self._s = None

    def image(self, which):
        ch = 2 if which == 4 else 4
        ptr = lib().vr_svgf_state_image(self._s, which)
        return np.ctypeslib.as_array(ptr, shape=(self.H, self.W, ch)).view(np.float16)

    def run(self, pfd, normals, motion, rt, want_iters=True):
        W, H = self.W, self.H
        den = np.zeros((H, W, 4), np.float16)
        iters = np.zeros((5, H, W, 4), np.float16) if want_iters else None
        temporal = np.zeros((H, W, 4), np.float16)
        lib().vr_svgf_pass(self._s, _p(pfd), _p(_h(normals)), _p(_h(motion)), _p(_h(rt)), _p(den), _p(iters), _p(temporal))
        return den, iters, temporal


def ssao(pfd, depth, normals, radius=0.75):
    H, W = depth.shape[:2]
    out = np.zeros((H, W, 4), np.float16)
    d = np.ascontiguousarray(depth, np.float32)
    lib().vr_ssao(_p(pfd), W, H, float(radius), _p(d), _p(_h(normals)), _p(out))
    return out


def ssao_blur(pfd, raw):
    H, W = raw.shape[:2]
    out = np.zeros((H, W, 4), np.float16)
    lib().vr_ssao_blur(_p(pfd), W, H, _p(_h(raw)), _p(out))
    return out


def ssr(pfd, albedo, normals, motion, depth, ray_distance=25.0, step_size=0.1, thickness=0.5, bsearch_steps=10, rows=None):
    H, W = depth.shape[:2]
    y0, y1 = (0, H) if rows is None else rows
    out = np.zeros((H, W, 4), np.float16)
    a8 = np.ascontiguousarray(albedo, np.uint8)
    d = np.ascontiguousarray(depth, np.float32)
    lib().vr_ssr(_p(pfd), W, H, y0, y1, float(ray_distance), float(step_size), float(thickness), int(bsearch_steps), _p(a8), _p(_h(normals)), _p(_h(motion)),
                 _p(d), _p(out))
    return out


def composition(pfd, albedo, normals, motion, depth, rt, shadow_mode=0, ao_mode=0, reflection_mode=2, ssao_img=None, ssr_img=None, refl=None,
                shadow_map=None, out_format=T.VK_FORMAT_R16G16B16A16_SFLOAT):
    H, W = depth.shape[:2]
    zero4 = np.zeros((H, W, 4), np.float16)
    ssao_img = zero4 if ssao_img is None else _h(ssao_img)
    ssr_img = zero4 if ssr_img is None else _h(ssr_img)
    refl = zero4 if refl is None else _h(refl)
    if shadow_map is None:
        shadow_map = np.zeros((4, 4), np.float32)
    sm = np.ascontiguousarray(shadow_map, np.float32)
    rt = _h(rt)
    out = np.zeros((H, W, 4), np.float16 if out_format == T.VK_FORMAT_R16G16B16A16_SFLOAT else np.uint8)
    a8 = np.ascontiguousarray(albedo, np.uint8)
    d = np.ascontiguousarray(depth, np.float32)
    lib().vr_composition(_p(pfd), W, H, int(shadow_mode), int(ao_mode), int(reflection_mode), _p(a8), _p(_h(normals)), _p(_h(motion)), _p(d), _p(sm),
                         sm.shape[1], sm.shape[0], _p(ssao_img), _p(ssr_img), _p(rt), rt.shape[-1], _p(refl), int(out_format), _p(out))
    return out


class _VrTexture(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("vk_format", C.c_int), ("mag", C.c_int), ("min", C.c_int), ("wrap_u", C.c_int),
                ("wrap_v", C.c_int)]


def _scene_arrays(scene):
    v = np.ascontiguousarray(scene.vertices)
    i = np.ascontiguousarray(scene.indices, np.uint32)
    p = np.ascontiguousarray(scene.primitives)
    texs = list(getattr(scene, "textures", []))
    keep = [np.ascontiguousarray(t.rgba, np.uint8) for t in texs]
    arr = (_VrTexture * max(len(texs), 1))()
    for k, t in enumerate(texs):
        mag, mn, wu, wv = (1, 1, 0, 0) if t.sampler is None else [int(x) for x in t.sampler]
        arr[k] = _VrTexture(keep[k].ctypes.data, keep[k].shape[1], keep[k].shape[0], int(t.format), mag, mn, wu, wv)
    return v, i, p, arr, len(texs), keep


def raytraced(scene, oracle_scene, pfd, W, H, alpha_test=False):
    """vkCmdTraceRaysKHR of the fully ray-traced path's "Raytracing Pipeline" compiled from the reference's raytraced_render_path/*.rgen / .rchit /
    .rahit / .rmiss: [H, W, 4] uint8 B8G8R8A8_UNORM "RaytracedOutput". traceRayEXT queries `oracle_scene`."""
    out = np.zeros((H, W, 4), np.uint8)
    v, i, p, arr, n, keep = _scene_arrays(scene)
    OL = O.lib()
    fn = lib().vr_raytraced_alpha if alpha_test else lib().vr_raytraced
    fn(oracle_scene._s, C.cast(OL.vo_trace_any, C.c_void_p), C.cast(OL.vo_trace_closest, C.c_void_p), C.cast(OL.vo_trace_closest_filtered, C.c_void_p),
       _p(pfd), W, H, _p(out), _p(v), _p(i), _p(p), C.cast(arr, C.c_void_p), n)
    return out


def raygen(scene, oracle_scene, pfd, depth, normals, rows=None):
    """vkCmdTraceRaysKHR of the reference's "Raytrace Pipeline" (raygen.rgen + both miss shaders + reflection_hit.rchit, compiled from the
    reference text; 2 AO samples and the 4x shadow loop exactly as written). traceRayEXT queries `oracle_scene` (oracle_lib.OracleScene)."""
    H, W = depth.shape[:2]
    y0, y1 = (0, H) if rows is None else rows
    sa = np.zeros((H, W, 2), np.float16)
    refl = np.zeros((H, W, 4), np.float16)
    d = np.ascontiguousarray(depth, np.float32)
    v = np.ascontiguousarray(scene.vertices)
    i = np.ascontiguousarray(scene.indices, np.uint32)
    p = np.ascontiguousarray(scene.primitives)
    texs = list(getattr(scene, "textures", []))
    keep = [np.ascontiguousarray(t.rgba, np.uint8) for t in texs]
    arr = (_VrTexture * max(len(texs), 1))()
    for k, t in enumerate(texs):
        mag, mn, wu, wv = (1, 1, 0, 0) if t.sampler is None else [int(x) for x in t.sampler]
        arr[k] = _VrTexture(keep[k].ctypes.data, keep[k].shape[1], keep[k].shape[0], int(t.format), mag, mn, wu, wv)
    OL = O.lib()
    lib().vr_raygen(oracle_scene._s, C.cast(OL.vo_trace_any, C.c_void_p), C.cast(OL.vo_trace_closest, C.c_void_p), _p(pfd), W, H, y0, y1, _p(d),
                    _p(_h(normals)), _p(sa), _p(refl), _p(v), _p(i), _p(p), C.cast(arr, C.c_void_p), len(texs))
    return dict(shadow_ao=sa, reflections=refl)
