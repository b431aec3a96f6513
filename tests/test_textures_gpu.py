"""GPU parity with material textures: the G-buffer producer (gbuf.frag: base colour + alpha cut-outs, normal maps,
metallic-roughness) and the reflection hit shader (reflection_hit.rchit:26-39) through the C-ABI vs the CPU oracle."""
import numpy as np
import pytest

import helpers as Hh
import checker as O
from vulkanhybridrenderer_b200 import camera, capi, scenes
from vulkanhybridrenderer_b200 import types as T

pytestmark = pytest.mark.gpu
F4, F2 = T.VK_FORMAT_R16G16B16A16_SFLOAT, T.VK_FORMAT_R16G16_SFLOAT


@pytest.fixture(scope="module")
def textured():
    W, H = 256, 144
    sc = scenes.add_procedural_textures(scenes.sponza_like(40_000, seed=3, width=W, height=H, n_clutter=60))
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    seq.next(sc.camera)
    sc.camera.set_pose(sc.camera.position + np.array([0.05, 0.0, 0.01]), sc.camera.yaw + 0.002, sc.camera.pitch)
    pfd = seq.next(sc.camera)
    return W, H, sc, osc, pfd, osc.gbuffer(pfd, W, H)


def test_texture_upload_slots_and_validation(textured):
    W, H, sc, osc, pfd, g = textured
    with capi.Context(W, H) as ctx:
        # geometry naming a texture that was never uploaded is rejected
        with pytest.raises(capi.VhrError):
            ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        a = ctx.upload_texture_from_data(sc.textures[0].rgba, sc.textures[0].format, sc.textures[0].sampler)
        b = ctx.upload_texture_from_data(sc.textures[1].rgba, T.VK_FORMAT_R8G8B8A8_UNORM)          # default sampler
        assert (a, b) == (0, 1)                                                                     # first free slot (resource_manager.cpp:821-824)
        with pytest.raises(capi.VhrError):
            ctx.upload_texture_from_data(sc.textures[0].rgba, T.VK_FORMAT_B8G8R8A8_UNORM)
        with pytest.raises(capi.VhrError):
            ctx.upload_texture_from_data(sc.textures[0].rgba, T.VK_FORMAT_R8G8B8A8_SRGB, (1, 1, 7, 0))
        ctx.destroy_textures()
        ctx.load_scene(sc)                                                                          # slots 0..n-1 again


def test_gbuffer_pass_with_textures(textured):
    W, H, sc, osc, pfd, g = textured
    with capi.Context(W, H) as ctx:
        ctx.load_scene(sc)
        ctx.update_per_frame_ubo(pfd)
        for name, fmt in Hh.GBUF_IMAGES.items():
            ctx.actualize_image(name, fmt)
        ctx.bind_pass_images(list(Hh.GBUF_IMAGES))
        ctx.gbuffer_pass(W, H)
        got = {k: ctx.image_download(n) for k, n in (("albedo", "Albedo"), ("normals", Hh.N_NORMALS), ("motion", Hh.N_MOTION), ("depth", Hh.N_DEPTH))}
    same_obj = got["normals"][..., 3] == g["normals"][..., 3]
    sky_agree = np.mean((got["depth"] == 0) == (g["depth"] == 0))
    print(f"[gbuffer+textures] same object id on {same_obj.mean()*100:.3f}% of pixels; sky agreement {sky_agree*100:.3f}%")
    # the alpha test sits on texture edges: a hit point that differs in the last bits may land on the other side of a cut-out
    assert same_obj.mean() >= 0.998
    m = same_obj & (g["depth"] > 0)
    rel = np.abs(got["depth"][m] - g["depth"][m]) / g["depth"][m]
    assert np.quantile(rel, 0.999) < 1e-4
    # albedo: bilinear texels at a hit point that agrees to ~1e-6 in uv -> at most one code on 99.5 % of pixels
    da = np.abs(got["albedo"][m].astype(np.int32) - g["albedo"][m].astype(np.int32)).max(axis=-1)
    print(f"[gbuffer+textures] albedo: exact {np.mean(da == 0)*100:.2f}%, <= 1 code {np.mean(da <= 1)*100:.3f}%")
    assert np.mean(da <= 1) >= 0.995
    dn = np.abs(got["normals"][..., :3].astype(np.float32) - g["normals"][..., :3].astype(np.float32))[m]
    assert np.quantile(dn.max(axis=-1), 0.995) <= 4e-3          # normal maps amplify uv differences through the texel gradient
    dm = np.abs(got["motion"].astype(np.float32) - g["motion"].astype(np.float32))[m]
    assert np.quantile(dm.max(axis=-1), 0.995) <= 2e-3
    # the textured branches were really taken
    mat = sc.primitives["material"]
    ids = np.maximum(got["normals"][..., 3].astype(np.int32), 0)
    assert (m & (mat["base_color_texture"][ids] >= 0)).mean() > 0.2 and (m & (mat["normal_map"][ids] >= 0)).any()


def test_reflections_with_textures(textured):
    W, H, sc, osc, pfd, g = textured
    ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=7, want_t=True)
    with capi.Context(W, H) as ctx:
        ctx.load_scene(sc)
        ctx.update_per_frame_ubo(pfd)
        ctx.set_option(capi.OPT_DEBUG_REFLECTION_T, 1)
        ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
        ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
        ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
        ctx.trace_rays(W, H)
        sa, refl = ctx.image_download(Hh.N_RT), ctx.image_download(Hh.N_REFL)
        rt = ctx.download_reflection_t()
    # shadow / AO rays are opaque (gl_RayFlagsOpaqueEXT): textures must not change the masks
    agree = np.mean(np.all(sa == ref["shadow_ao"], axis=-1))
    assert agree >= 0.9999, agree
    ref_t = ref["refl_t"]
    both = (rt >= 0) & (ref_t >= 0)
    good = both & (np.abs(rt - ref_t) <= 1e-3 * np.maximum(1.0, ref_t))
    a = refl.astype(np.float32)[good]; b = ref["reflections"].astype(np.float32)[good]
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b) / 2.0)
    frac_ok = np.mean(err.max(axis=-1) <= 2e-3)
    print(f"[reflection+textures] radiance within tol on {frac_ok*100:.3f}% of {good.sum()} matched hits, psnr {Hh.psnr(a, b, peak=max(1.0, float(b.max()))):.1f} dB")
    # NEAREST-filtered and checker textures are discontinuous: a hit point that differs in the last bits can fetch the next texel
    assert frac_ok >= 0.99
    # against the same scene without textures the radiance differs where textured primitives are hit
    plain = scenes.sponza_like(40_000, seed=3, width=W, height=H, n_clutter=60)
    rp = O.OracleScene(plain).raygen(pfd, g["depth"], g["normals"], flags=4)
    assert np.abs(refl.astype(np.float32) - rp["reflections"].astype(np.float32)).max() > 0.05


def test_gltf_scene_path_renders_like_direct_upload(tmp_path):
    """SceneLoader::LoadScene (host/scene_loader.cpp) -> textures + UpdateGeometry -> one frame through the C++ render graph must
    equal the frame of the same scene uploaded array by array (texture slots are renumbered by first use; the images cannot tell)."""
    from vulkanhybridrenderer_b200 import gltf_export, host_api
    from vulkanhybridrenderer_b200 import hybrid_path as HP
    W, H = 160, 96
    sc = scenes.add_procedural_textures(scenes.sponza_like(12_000, seed=5, width=W, height=H, n_clutter=20), size=32)
    pfd = camera.FrameSequencer(W, H, sc.light).next(sc.camera)
    path = gltf_export.export(sc, tmp_path / "scene.glb")
    outs = []
    for via_gltf in (False, True):
        with host_api.Renderer(W, H) as r:
            if via_gltf:
                assert r.load_gltf(path) == len(sc.primitives)
            else:
                for i, t in enumerate(sc.textures):
                    assert r.ctx.upload_texture_from_data(t.rgba, t.format, t.sampler) == i
                r.load_scene(sc)
            r.set_modes(shadow=0, ao=0, reflection=0, denoise=True)
            r.set_gbuffer_producer(True)
            r.render(pfd)
            outs.append({n: r.ctx.image_download(n) for n in (HP.N_ALBEDO, HP.N_NORMALS, HP.N_MOTION, HP.N_DEPTH, HP.N_RT, HP.N_REFL, HP.N_RENDER_OUTPUT)})
    for n in outs[0]:
        a, b = outs[0][n], outs[1][n]
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), n
    assert outs[1][HP.N_REFL].astype(np.float32).max() > 0
    # a missing file leaves an empty scene (scene_loader.cpp:344-346), it does not throw
    with host_api.Renderer(W, H) as r:
        assert r.load_gltf(tmp_path / "nope.glb") == 0
