"""GPU: the fused multi-GPU partition (interleaved ray blocks stored into the owners' images, halo rows pushed by the SVGF
kernels, flag-word ordering between streams) against the unpartitioned frames — bit for bit.

The ranks here are contexts of ONE process on ONE GPU (peers attached by device pointer), so the test runs on a single-GPU
box; the code path inside the library is the one `bench.py --partition rows` uses across GPUs (there the peers come from
CUDA IPC handles; `tools/fused_partition_parity.py` is the same check over real ranks)."""
import numpy as np
import pytest

import helpers as Hh
from vulkanhybridrenderer_b200 import camera, capi, scenes
from vulkanhybridrenderer_b200 import hybrid_path as HP
from vulkanhybridrenderer_b200 import multi_gpu as MG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_fused_partition_matches_single_context(world):
    W, H, n_frames = 320, 64 * world + 17, 4           # bands of >= 64 rows, ragged last band
    sc = scenes.sponza_like(20_000, seed=3, width=W, height=H, n_clutter=30)

    def make(rt_sets):
        ctx = capi.Context(W, H)
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.set_option(capi.OPT_AO_SPP, 2)
        ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 1)
        return ctx, HP.HybridRenderPath(ctx, W, H, rt_sets=rt_sets, ssao=True, composition=HP.T.VK_FORMAT_B8G8R8A8_SRGB, shadow_map_size=(8, 8))

    ref_ctx, ref_path = make(1)
    ranks = [make(2) for _ in range(world)]
    ctxs, paths = [c for c, _ in ranks], [p for _, p in ranks]
    try:
        halo = 24
        MG.setup_fused_partition_inprocess(ctxs, paths, motion_halo=halo)
        bands = [MG.band_rows(H, world, r) for r in range(world)]
        seq = camera.FrameSequencer(W, H, sc.light)
        cam = sc.camera
        for f in range(n_frames):
            if f:
                cam.set_pose(cam.position + np.array([0.05, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)
            pfd = seq.next(cam)
            # G-buffer: an input of the path; rendered once, handed to every rank in full
            ref_ctx.update_per_frame_ubo(pfd)
            g = ref_path.gsets[0]
            ref_ctx.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
            ref_ctx.gbuffer_pass(W, H)
            gb = {k: ref_ctx.image_download(g[k]) for k in (HP.N_ALBEDO, HP.N_NORMALS, HP.N_MOTION, HP.N_DEPTH)}
            assert MG.required_motion_halo(float(np.abs(gb[HP.N_MOTION][..., 1].astype(np.float32)).max()), H) <= halo
            for ctx, path in zip(ctxs, paths):
                for k, v in gb.items():
                    ctx.image_upload(path.gsets[0][k], v)
            ref_path.frame(pfd)
            ref_path.ssao_passes()
            ref_path.composition_pass(0, 1, 0, denoised=True)        # ray-traced shadows, SSAO, ray-traced reflections
            # every rank issues the plain single-GPU call sequence; nothing blocks on the host in between, the streams order
            # themselves through the flag words
            for path in paths:
                path.frame(pfd, rtset=f & 1)
            # SSAO (6-row halo of the raw image pushed by ssao.comp's kernel) and the composition pass on the same bands
            for path in paths:
                path.ssao_passes()
            for ctx, path in zip(ctxs, paths):
                g0 = path.gsets[0]
                ctx.bind_pass_images([g0[HP.N_ALBEDO], g0[HP.N_NORMALS], g0[HP.N_MOTION], g0[HP.N_DEPTH], HP.N_SHADOW_MAP, HP.N_SSAO, HP.N_SSR,
                                      HP.N_DENOISED, path.rt_sets[f & 1][1], HP.N_RENDER_OUTPUT])
                ctx.draw(HP.SHADER_COMPOSITION, (0, 1, 0))
            want = {k: ref_ctx.image_download(n) for k, n in (("rt", HP.N_RT), ("refl", HP.N_REFL), ("den", HP.N_DENOISED), ("ssao", HP.N_SSAO),
                                                              ("out", HP.N_RENDER_OUTPUT))}
            for r, (ctx, path) in enumerate(zip(ctxs, paths)):
                y0, y1 = bands[r]
                got = {"rt": ctx.image_download(path.rt_sets[f & 1][0]), "refl": ctx.image_download(path.rt_sets[f & 1][1]),
                       "den": ctx.image_download(HP.N_DENOISED), "ssao": ctx.image_download(HP.N_SSAO), "out": ctx.image_download(HP.N_RENDER_OUTPUT)}
                for k in want:
                    bad = int((got[k][y0:y1].view(np.uint8) != want[k][y0:y1].view(np.uint8)).sum())
                    assert bad == 0, f"frame {f} rank {r}/{world} image {k}: {bad} mismatching bytes in rows [{y0},{y1})"
    finally:
        for c in ctxs + [ref_ctx]:
            c.close()
