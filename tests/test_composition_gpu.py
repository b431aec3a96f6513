"""GPU parity: the composition kernel (vhr_draw, composition.frag) vs the CPU oracle on identical inputs."""
import numpy as np
import pytest

import helpers as Hh
import checker as O
from vulkanhybridrenderer_b200 import capi
from vulkanhybridrenderer_b200 import hybrid_path as HP
from vulkanhybridrenderer_b200 import types as T

pytestmark = pytest.mark.gpu

F4 = T.VK_FORMAT_R16G16B16A16_SFLOAT


def _inputs(W, H, seed=4):
    sc, osc, frames = Hh.scene_and_gbuffer(W, H)
    pfd, g = frames[1]
    rt = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=7)
    rng = np.random.default_rng(seed)
    ssao = rng.uniform(0, 1, (H, W, 4)).astype(np.float16)
    ssr = rng.uniform(0, 2, (H, W, 4)).astype(np.float16)
    den = np.concatenate([rng.uniform(0, 1, (H, W, 2)), rng.uniform(0, 0.2, (H, W, 2))], -1).astype(np.float16)
    shadow_map = rng.uniform(0.0, 1.0, (64, 48)).astype(np.float32)
    return pfd, g, rt, ssao, ssr, den, shadow_map


def _upload(ctx, path, g, rt, ssao, ssr, den, shadow_map):
    gs = path.gsets[0]
    ctx.image_upload(gs[HP.N_ALBEDO], g["albedo"]); ctx.image_upload(gs[HP.N_NORMALS], g["normals"])
    ctx.image_upload(gs[HP.N_MOTION], g["motion"]); ctx.image_upload(gs[HP.N_DEPTH], g["depth"])
    ctx.image_upload(HP.N_RT, rt["shadow_ao"]); ctx.image_upload(HP.N_REFL, rt["reflections"])
    ctx.image_upload(HP.N_SSAO, ssao); ctx.image_upload(HP.N_SSR, ssr); ctx.image_upload(HP.N_DENOISED, den)
    ctx.image_upload(HP.N_SHADOW_MAP, shadow_map)


@pytest.mark.parametrize("size", [(160, 96), (203, 77)])
@pytest.mark.parametrize("modes,denoised", [((0, 0, 2), True), ((0, 0, 0), False), ((2, 1, 1), True), ((1, 2, 0), True), ((2, 2, 2), False)])
def test_composition_hdr_parity(size, modes, denoised):
    W, H = size
    pfd, g, rt, ssao, ssr, den, sm = _inputs(W, H)
    with capi.Context(W, H) as ctx:
        ctx.update_per_frame_ubo(pfd)
        path = HP.HybridRenderPath(ctx, W, H, composition=F4, shadow_map_size=(sm.shape[1], sm.shape[0]))
        _upload(ctx, path, g, rt, ssao, ssr, den, sm)
        path.composition_pass(*modes, denoised=denoised)
        got = ctx.image_download(HP.N_RENDER_OUTPUT).astype(np.float32)
    ref = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], den if denoised else rt["shadow_ao"], *modes,
                        ssao_img=ssao, ssr_img=ssr, refl=rt["reflections"], shadow_map=sm).astype(np.float32)
    assert np.isfinite(ref).all() and np.isfinite(got).all()
    # linear HDR radiance: 1e-3 absolute up to 1.0, relative above (an fp16 ulp at 2.0 is already 2e-3)
    d = np.abs(got - ref)
    tol = 1e-3 * np.maximum(1.0, np.abs(ref))
    bad = d > tol
    print(f"[parity] composition {modes} {W}x{H}: max_abs={d.max():.3e} exact={np.mean(got == ref) * 100:.2f}% psnr={Hh.psnr(got, ref, peak=max(1.0, float(ref.max()))):.1f}dB")
    assert not bad.any(), (int(bad.sum()), float(d.max()))
    assert Hh.psnr(got, ref, peak=max(1.0, float(ref.max()))) >= Hh.PSNR_MIN_DB


@pytest.mark.parametrize("fmt", [T.VK_FORMAT_B8G8R8A8_SRGB, T.VK_FORMAT_B8G8R8A8_UNORM])
def test_composition_8bit_outputs(fmt):
    W, H = 160, 96
    pfd, g, rt, ssao, ssr, den, sm = _inputs(W, H)
    with capi.Context(W, H) as ctx:
        ctx.update_per_frame_ubo(pfd)
        path = HP.HybridRenderPath(ctx, W, H, composition=fmt, shadow_map_size=(sm.shape[1], sm.shape[0]))
        _upload(ctx, path, g, rt, ssao, ssr, den, sm)
        path.composition_pass(0, 0, 0, denoised=False)
        got = ctx.image_download(HP.N_RENDER_OUTPUT)
    ref = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], rt["shadow_ao"], 0, 0, 0, refl=rt["reflections"], out_format=fmt)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    print(f"[parity] composition 8-bit fmt {fmt}: exact={np.mean(diff == 0) * 100:.3f}% max code diff {diff.max()}")
    assert diff.max() <= 1                      # float -> UNORM8 ties / pow() ulps
    assert np.mean(diff == 0) >= 0.995


def test_composition_errors():
    W, H = 64, 32
    pfd, g, rt, ssao, ssr, den, sm = _inputs(W, H)
    with capi.Context(W, H) as ctx:
        ctx.update_per_frame_ubo(pfd)
        path = HP.HybridRenderPath(ctx, W, H, composition=F4)
        gs = path.gsets[0]
        names = [gs[HP.N_ALBEDO], gs[HP.N_NORMALS], gs[HP.N_MOTION], gs[HP.N_DEPTH], HP.N_SHADOW_MAP, HP.N_SSAO, HP.N_SSR, HP.N_DENOISED, HP.N_REFL,
                 HP.N_RENDER_OUTPUT]
        ctx.bind_pass_images(names)
        with pytest.raises(capi.VhrError):
            ctx.draw("hybrid_render_path/gbuf.frag", (0, 0, 0))          # no rasteriser
        with pytest.raises(capi.VhrError):
            ctx.draw(HP.SHADER_COMPOSITION, (0, 0, 0), vertex_count=6)    # not the full-screen triangle
        with pytest.raises(capi.VhrError):
            ctx.draw(HP.SHADER_COMPOSITION, (0, 0))                       # wrong number of specialization constants
        with pytest.raises(capi.VhrError):
            ctx.draw(HP.SHADER_COMPOSITION, (0, 3, 0))                    # mode out of range
        ctx.bind_pass_images(names[:9])
        with pytest.raises(capi.VhrError):
            ctx.draw(HP.SHADER_COMPOSITION, (0, 0, 0))                    # render output not bound
        ctx.bind_pass_images(names[:1] + [gs[HP.N_DEPTH]] + names[2:])
        with pytest.raises(capi.VhrError):
            ctx.draw(HP.SHADER_COMPOSITION, (0, 0, 0))                    # wrong format at binding 1
