"""CPU: pins the hand-written oracle (oracle/*.cpp) to oracle/_ref — the reference's OWN shader files compiled for the CPU
(oracle/make_ref.py reads /root/reference/data/shaders/*.{comp,rgen,rchit,rmiss,frag} + common.glsl + glsl_common.h at build time;
types and built-ins are the reference's vendored glm). Every comparison here is BIT FOR BIT: fp16 images as uint16, 8-bit images as
bytes, helper results as fp32 bit patterns.

Where the implementation is free (GLSL leaves normalize()/mat*vec evaluation order and NaN handling of max()/pow() open) both sides
follow one documented choice: glm's formulas for the former, the behaviour of the RT-capable GPUs the reference needs for the
latter (oracle/ref_shim.h).
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as Hh
import oracle_lib as O
import ref_lib as R
from vulkanhybridrenderer_b200 import camera, scenes
from vulkanhybridrenderer_b200 import types as T

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref cannot be built here (no /root/reference) and no prebuilt library travelled")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({2: np.uint16, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def assert_same(a, b, what):
    a, b = bits(a), bits(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = int(np.count_nonzero(a != b))
    assert bad == 0, f"{what}: {bad} of {a.size} values differ between oracle/ and oracle/_ref"


# ---- common.glsl + glsl_common.h ------------------------------------------------------------------------------------------------
def test_struct_sizes_from_the_reference_header():
    L = R.lib()
    want = {0: T.PerFrameData.itemsize, 1: T.Vertex.itemsize, 3: T.Primitive.itemsize, 4: T.SVGFPushConstants.itemsize, 5: 16, 6: 4, 7: 112}
    for which, size in want.items():
        assert L.vr_sizeof(which) == size, (which, L.vr_sizeof(which), size)
    assert L.vr_sizeof(2) == 44 and L.vr_sizeof(0) == 584 and L.vr_sizeof(1) == 56 and L.vr_sizeof(3) == 120


def test_rng_known_answers_come_from_the_reference_text():
    """The KATs of SURVEY Appendix A.3 / tests/golden/rng_kat.npz were derived by hand in round 1; here common.glsl:47-76 itself runs."""
    L, OL = R.lib(), O.lib()
    kat = np.load(os.path.join(GOLDEN, "rng_kat.npz"))
    for seed, state in zip(kat["seeds"], kat["states"]):
        assert L.vr_seed_thread(int(seed)) == int(state) == OL.vo_seed_thread(int(seed))
    for i, s in enumerate(kat["states"]):
        st = C.c_uint32(int(s))
        got = np.array([L.vr_random01(C.byref(st)) for _ in range(8)], np.float32)
        assert_same(got, kat["random01"][i], f"random01 stream of seed #{i}")
    rng = np.random.default_rng(11)
    for seed in rng.integers(0, 2**32, 2000, dtype=np.uint64):
        a, b = C.c_uint32(L.vr_seed_thread(int(seed))), C.c_uint32(OL.vo_seed_thread(int(seed)))
        assert a.value == b.value
        for _ in range(4):
            x, y = L.vr_random01(C.byref(a)), OL.vo_random01(C.byref(b))
            assert np.float32(x).view(np.uint32) == np.float32(y).view(np.uint32) and a.value == b.value
            assert 0.0 <= x < 1.0


def test_sampling_helpers_bit_exact():
    L, OL = R.lib(), O.lib()
    rng = np.random.default_rng(5)
    a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    for u0, u1 in rng.uniform(0, 1, (3000, 2)).astype(np.float32):
        L.vr_uniform_sample_cone(float(u0), float(u1), 0.999995, O._p(a)); OL.vo_uniform_sample_cone(float(u0), float(u1), 0.999995, O._p(b))
        assert_same(a, b, "uniform_sample_cone")
        L.vr_cosine_hemisphere(float(u0), float(u1), O._p(a)); OL.vo_cosine_hemisphere(float(u0), float(u1), O._p(b))
        assert_same(a, b, "uniform_sample_cosine_weighted_hemisphere")
    m, n = np.zeros(9, np.float32), np.zeros(9, np.float32)
    dirs = rng.standard_normal((3000, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs = np.concatenate([dirs, np.array([[0, 0, -1], [0, 0, 1], [1e-4, 0, -0.99999994]], np.float32)])
    for d in dirs:
        d = np.ascontiguousarray(d, np.float32)
        L.vr_onb(O._p(d), O._p(m)); OL.vo_onb(O._p(d), O._p(n))
        assert_same(m, n, f"onb_from_unit_vector({d})")


def test_fp16_conversion_tables_agree():
    L, OL = R.lib(), O.lib()
    for h in range(0, 65536):
        fa, fb = L.vr_h2f(h), OL.vo_h2f(h)
        assert np.float32(fa).view(np.uint32) == np.float32(fb).view(np.uint32) or (fa != fa and fb != fb)
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 20000).astype(np.float32),
                         np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, np.inf, -np.inf], np.float32)])
    for x in xs:
        assert L.vr_f2h(float(x)) == OL.vo_f2h(float(x)), x
    # numpy's float16 conversion is a third, independent implementation of round-to-nearest-even
    want = xs.astype(np.float16).view(np.uint16)
    got = np.array([L.vr_f2h(float(x)) for x in xs], np.uint16)
    assert np.array_equal(got, want)


def test_texture_unit_bit_exact_all_modes():
    """texture(textures[i], uv) of the shim (Vulkan LOD-0 formulas) vs the oracle's sampler: filters x address modes x UNORM / sRGB."""
    rng = np.random.default_rng(3)
    rgba = rng.integers(0, 256, (7, 5, 4), dtype=np.uint8)
    sc = scenes.sponza_like(500, seed=1, width=32, height=32, n_clutter=2)
    osc = O.OracleScene(sc)
    out_r, idx = np.zeros(4, np.float32), 0
    uvs = np.concatenate([rng.uniform(-2.5, 3.5, (400, 2)), np.array([[0, 0], [1, 1], [0.5, 0.5], [0.1, 0.9999999], [-1e-7, 1.0]])]).astype(np.float32)
    for fmt in (T.VK_FORMAT_R8G8B8A8_UNORM, 43):
        for mag in (0, 1):
            for wu in range(4):
                for wv in range(4):
                    idx = osc.add_texture(rgba, fmt, (mag, mag, wu, wv))
                    for u, v in uvs:
                        R.lib().vr_sample_rgba8(O._p(rgba), 5, 7, int(fmt), mag, mag, wu, wv, float(u), float(v), O._p(out_r))
                        assert_same(out_r, osc.sample_texture(idx, u, v), f"texture fmt {fmt} filter {mag} wrap ({wu},{wv}) uv ({u},{v})")


# ---- shader passes ----------------------------------------------------------------------------------------------------------------
def _frames(W, H, tris, seed, n_frames, textured=False):
    sc = scenes.sponza_like(tris, seed=seed, width=W, height=H, n_clutter=12)
    if textured:
        sc = scenes.add_procedural_textures(sc, size=32)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera
    for f in range(n_frames):
        if f:
            cam.set_pose(cam.position + np.array([0.06, 0.0, 0.02]), cam.yaw + 0.004, cam.pitch)
        pfd = seq.next(cam)
        yield sc, osc, pfd, osc.gbuffer(pfd, W, H)


@pytest.mark.parametrize("textured", [False, True])
@pytest.mark.parametrize("size", [(96, 64), (101, 59)])
def test_every_pass_of_the_hybrid_path_bit_exact_over_three_frames(size, textured):
    W, H = size
    so, sr = O.SvgfState(W, H), R.SvgfState(W, H)
    sm = np.random.default_rng(1).uniform(0, 1, (32, 32)).astype(np.float32)
    for f, (sc, osc, pfd, g) in enumerate(_frames(W, H, 6000, 21, 3, textured)):
        a = osc.raygen(pfd, g["depth"], g["normals"])                      # oracle: raygen.rgen restated
        b = R.raygen(sc, osc, pfd, g["depth"], g["normals"])               # the reference's raygen.rgen + miss + reflection_hit.rchit
        assert_same(a["shadow_ao"], b["shadow_ao"], f"frame {f} Raytraced Shadows and Ambient Occlusion")
        assert_same(a["reflections"], b["reflections"], f"frame {f} Raytraced Reflections")
        do, io, to = so.run(pfd, g["normals"], g["motion"], a["shadow_ao"])
        dr, ir, tr = sr.run(pfd, g["normals"], g["motion"], a["shadow_ao"])
        assert_same(to, tr, f"frame {f} svgf.comp integrated")
        assert_same(io, ir, f"frame {f} five a-trous iterations")
        assert_same(do, dr, f"frame {f} Denoised (= iteration 3, SURVEY Q1)")
        for k in range(5):
            assert_same(so.image(k), sr.image(k), f"frame {f} persistent SVGF image {k}")
        ra, rb = O.ssao(pfd, g["depth"], g["normals"]), R.ssao(pfd, g["depth"], g["normals"])
        assert_same(ra, rb, f"frame {f} ssao.comp")
        assert_same(O.ssao_blur(pfd, ra), R.ssao_blur(pfd, ra), f"frame {f} ssao_blur.comp")
        sa, sb = (F(pfd, g["albedo"], g["normals"], g["motion"], g["depth"]) for F in (O.ssr, R.ssr))
        assert_same(sa, sb, f"frame {f} ssr.comp")
        for modes in ((0, 0, 0), (1, 1, 1), (2, 2, 2), (0, 1, 2), (1, 0, 1)):
            for fmt in (T.VK_FORMAT_R16G16B16A16_SFLOAT, T.VK_FORMAT_B8G8R8A8_SRGB, T.VK_FORMAT_B8G8R8A8_UNORM):
                kw = dict(ssao_img=ra, ssr_img=sa, refl=a["reflections"], shadow_map=sm, out_format=fmt)
                x = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], do, *modes, **kw)
                y = R.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], do, *modes, **kw)
                assert_same(x, y, f"frame {f} composition.frag modes {modes} format {fmt}")
            x = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], a["shadow_ao"], *modes, refl=a["reflections"])
            y = R.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], a["shadow_ao"], *modes, refl=a["reflections"])
            assert_same(x, y, f"frame {f} composition.frag on the raw RG16F image, modes {modes}")


@pytest.mark.parametrize("step", [1, 2, 3, 4, 8, 16])
def test_atrous_on_worst_case_noise(step):
    W, H = 96, 64
    sc, osc, pfd, g = next(_frames(W, H, 6000, 21, 1))
    integ = Hh.noise_integrated(H, W, seed=2)
    assert_same(O.svgf_atrous(pfd, g["normals"], integ, step), R.svgf_atrous(pfd, g["normals"], integ, step), f"a-trous step {step} on noise")
    # opposing normals: pow(negative, 128) must give weight 0, not C's (+1)
    flipped = g["normals"].copy()
    flipped[::2, :, :3] *= np.float16(-1.0)
    assert_same(O.svgf_atrous(pfd, flipped, integ, step), R.svgf_atrous(pfd, flipped, integ, step), f"a-trous step {step}, opposing normals")


def test_temporal_with_random_history_and_large_motion():
    W, H = 120, 72
    it = _frames(W, H, 6000, 4, 2)
    sc, osc, pfd0, g0 = next(it)
    _, _, pfd1, g1 = next(it)
    rng = np.random.default_rng(7)
    rt = np.stack([rng.integers(0, 2, (H, W)), rng.integers(0, 3, (H, W)) * 0.5], -1).astype(np.float16)
    history = rng.uniform(0, 1, (H, W, 4)).astype(np.float16)
    moments = rng.uniform(0, 1, (H, W, 2)).astype(np.float16)
    for scale in (1.0, 8.0, -300.0):        # the last one throws most reprojections off screen
        motion = g1["motion"].copy()
        motion[..., :2] = (motion[..., :2].astype(np.float32) * scale).astype(np.float16)
        ai, am = O.svgf_temporal(pfd1, g1["normals"], motion, rt, g0["normals"], history, moments)
        bi, bm = R.svgf_temporal(pfd1, g1["normals"], motion, rt, g0["normals"], history, moments)
        assert_same(ai, bi, f"svgf.comp integrated, motion x{scale}")
        assert_same(am, bm, f"svgf.comp moments, motion x{scale}")


def test_ssao_radius_and_sky_samples():
    W, H = 96, 64
    sc, osc, pfd, g = next(_frames(W, H, 6000, 21, 1))
    assert np.count_nonzero(g["depth"] == 0) > 0, "the frame should contain sky (samples that unproject to infinity / NaN)"
    for radius in (0.75, 0.2, 3.0):
        assert_same(O.ssao(pfd, g["depth"], g["normals"], radius), R.ssao(pfd, g["depth"], g["normals"], radius), f"ssao.comp radius {radius}")


def test_golden_fixtures_are_outputs_of_the_reference_shaders():
    """tests/golden/hybrid_frames_96x64.npz and atrous_noise_96x64.npz: every shader-pass output stored there is reproduced by oracle/_ref
    from the stored inputs (tools/make_golden.py generates them from oracle/_ref)."""
    z = np.load(os.path.join(GOLDEN, "hybrid_frames_96x64.npz"))
    W, H = 96, 64

    class _Sc:
        vertices, indices, primitives, textures = z["vertices"], z["indices"], z["primitives"], []
    osc = O.OracleScene(_Sc)
    st = R.SvgfState(W, H)
    for f in range(3):
        pfd = z[f"f{f}_pfd"]
        b = R.raygen(_Sc, osc, pfd, z[f"f{f}_depth"], z[f"f{f}_normals"])
        assert_same(b["shadow_ao"], z[f"f{f}_shadow_ao"], f"golden frame {f} shadow/AO")
        assert_same(b["reflections"], z[f"f{f}_reflections"], f"golden frame {f} reflections")
        den, iters, temporal = st.run(pfd, z[f"f{f}_normals"], z[f"f{f}_motion"], z[f"f{f}_shadow_ao"])
        assert_same(temporal, z[f"f{f}_temporal"], f"golden frame {f} temporal")
        assert_same(iters, z[f"f{f}_atrous"], f"golden frame {f} a-trous")
        assert_same(den, z[f"f{f}_denoised"], f"golden frame {f} denoised")
        raw = R.ssao(pfd, z[f"f{f}_depth"], z[f"f{f}_normals"], 0.75)
        assert_same(raw, z[f"f{f}_ssao_raw"], f"golden frame {f} ssao raw")
        assert_same(R.ssao_blur(pfd, raw), z[f"f{f}_ssao"], f"golden frame {f} ssao")
    n = np.load(os.path.join(GOLDEN, "atrous_noise_96x64.npz"))
    for s in (1, 2, 3, 4, 8, 16):
        assert_same(R.svgf_atrous(n["pfd"], n["normals"], n["integ"], s), n[f"step{s}"], f"golden a-trous noise step {s}")
    t = np.load(os.path.join(GOLDEN, "textured_frame_96x64.npz"))
    assert_same(R.ssr(t["pfd"], t["albedo"], t["normals"], t["motion"], t["depth"]), t["ssr"], "golden textured frame ssr")


@pytest.mark.parametrize("textured", [False, True])
def test_fully_raytraced_path_bit_exact(textured):
    """raytraced_render_path/{raygen.rgen, closesthit.rchit, miss.rmiss, shadow_miss.rmiss} (and the alpha-tested pipeline: raygen_test_alpha.rgen,
    closesthit_test_alpha.rchit, shadow_anyhit.rahit run as the any-hit stage of the oracle's traversal) compiled from the reference: the
    8-bit "RaytracedOutput" of the port is identical. The alpha-tested shaders index textures[base_color_texture] unconditionally, which is undefined
    for a material without a texture (index -1): there the port falls back to the base colour, so the comparison is on the textured primitives."""
    W, H = 96, 64
    for f, (sc, osc, pfd, g) in enumerate(_frames(W, H, 6000, 21, 2, textured)):
        assert_same(osc.raytraced(pfd, W, H, False), R.raytraced(sc, osc, pfd, W, H, False), f"frame {f} Raytracing Pipeline (opaque)")
        a, b = osc.raytraced(pfd, W, H, True), R.raytraced(sc, osc, pfd, W, H, True)
        ids = osc.gbuffer(pfd, W, H, want_ids=True)["ids"][..., 0]
        tex = sc.primitives["material"]["base_color_texture"]
        on_textured = np.where(ids >= 0, tex[np.clip(ids, 0, len(tex) - 1)] >= 0, True)      # sky pixels (miss.rmiss) count too
        # `ids` comes from the G-buffer producer's primary ray, which is generated differently from this path's: on a silhouette pixel the two may
        # see different primitives, one of them untextured
        bad = np.any(a[on_textured] != b[on_textured], axis=-1)
        assert bad.sum() <= 3, f"frame {f} Raytracing Pipeline (alpha-tested): {int(bad.sum())} of {int(on_textured.sum())} pixels on textured primitives / sky differ"
        if textured:
            assert on_textured.sum() > 0.4 * W * H
