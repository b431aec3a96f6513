"""Two frames in flight (vhr_select_queue / vhr_queue_signal / vhr_queue_wait + HybridRenderPath.frame_overlapped): the Raytrace Pass
of frame k+1 runs on queue 1 under the SVGF Denoise Pass of frame k. Only the schedule changes, so every image — including the
temporal history that accumulates all earlier frames — must be bit-identical to the frames rendered one after the other."""
import numpy as np
import pytest

from vulkanhybridrenderer_b200 import camera, capi, scenes
from vulkanhybridrenderer_b200 import hybrid_path as HP

pytestmark = pytest.mark.gpu


def _render(schedule, n_frames, W=328, H=184, tris=20_000, read_every_frame=False):
    sc = scenes.sponza_like(tris, seed=3, width=W, height=H, n_clutter=40)
    per_frame = []
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.set_option(capi.OPT_TRACE_SHADOWS, 1); ctx.set_option(capi.OPT_TRACE_AO, 1)
        ctx.set_option(capi.OPT_AO_SPP, 2); ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 1)
        path = HP.HybridRenderPath(ctx, W, H, gbuffer_sets=2, rt_sets=2)
        seq = camera.FrameSequencer(W, H, sc.light)
        cam = sc.camera
        p0 = (cam.position.copy(), cam.yaw, cam.pitch)
        p1 = (cam.position + np.array([0.05, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)
        pfds = [None, None]
        for s in (1, 0, 1):                      # both poses rendered into the two resident G-buffer sets (as bench.py does)
            cam.set_pose(*(p1 if s else p0))
            pfd = seq.next(cam)
            ctx.update_per_frame_ubo(pfd)
            g = path.gsets[s]
            ctx.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
            ctx.gbuffer_pass(W, H)
            pfds[s] = pfd
        ctx.synchronize()
        for k in range(n_frames):
            s = k & 1
            pfd = pfds[s].copy()
            pfd["frame_index"] = 3 + k
            if schedule == "serial":
                path.frame(pfd, gset=s, rtset=s)
            else:
                path.frame_overlapped(pfd, k, gset=s)
            if read_every_frame:
                per_frame.append((ctx.image_download(HP.N_DENOISED), ctx.image_download(path.rt_sets[s][0])))
        ctx.select_queue(0)
        last = (n_frames - 1) & 1
        out = {"denoised": ctx.image_download(HP.N_DENOISED), "rt": ctx.image_download(path.rt_sets[last][0]),
               "refl": ctx.image_download(path.rt_sets[last][1])}
        pc = path.pc
        out["integrated0"] = ctx.storage_image_download(int(pc["integrated_shadow_and_ao"][0]))
        out["integrated1"] = ctx.storage_image_download(int(pc["integrated_shadow_and_ao"][1]))
        out["history"] = ctx.storage_image_download(int(pc["shadow_and_ao_history"]))
        out["moments"] = ctx.storage_image_download(int(pc["shadow_and_ao_moments_history"]))
        out["prev_normals"] = ctx.storage_image_download(int(pc["prev_frame_normals_and_object_ids"]))
    return out, per_frame


def _same(a, b, what):
    np.testing.assert_array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8), err_msg=what)


def test_overlapped_frames_equal_serial_frames():
    ref, _ = _render("serial", 9)
    for attempt in range(3):                     # a race would not show on every run
        got, _ = _render("overlapped", 9)
        for key in ref:
            _same(got[key], ref[key], f"{key} (attempt {attempt})")


def test_overlapped_frames_read_back_every_frame():
    """Reading frame k's images on queue 0 between the frames (the e2e pattern) sees frame k, not frame k+1's ray pass."""
    _, ref = _render("serial", 5, read_every_frame=True)
    _, got = _render("overlapped", 5, read_every_frame=True)
    for k, (r, g) in enumerate(zip(ref, got)):
        _same(g[0], r[0], f"denoised, frame {k}")
        _same(g[1], r[1], f"raw shadow/AO, frame {k}")


def test_queue_arguments():
    with capi.Context(64, 64) as ctx:
        with pytest.raises(capi.VhrError):
            ctx.select_queue(2)
        with pytest.raises(capi.VhrError):
            ctx.queue_signal(capi.MAX_SEMAPHORES)
        ctx.queue_wait(3)                        # never signalled: a no-op
        ctx.select_queue(1)
        with pytest.raises(capi.VhrError):       # the peer flag words are ordered on queue 0
            ctx.set_partition(1, 0, [0, 64])
        ctx.select_queue(0)
