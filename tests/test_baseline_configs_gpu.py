"""GPU parity at the FULL sizes of BASELINE.json's five configs, through the C-ABI.

The small-size tests (test_svgf_gpu / test_rt_gpu / ...) pin the arithmetic; these run the same comparisons at the sizes the
metric is quoted on — the CPU oracle still finishes each in seconds — plus size-independent properties where the whole frame
would be too slow on the CPU (determinism, sky rule, AO quantisation, partitioned == unpartitioned)."""
import numpy as np
import pytest

import helpers as Hh
import checker as O
from vulkanhybridrenderer_b200 import camera, capi, scenes
from vulkanhybridrenderer_b200 import hybrid_path as HP
from vulkanhybridrenderer_b200 import multi_gpu as MG
from vulkanhybridrenderer_b200 import types as T

pytestmark = pytest.mark.gpu

F4, F2 = T.VK_FORMAT_R16G16B16A16_SFLOAT, T.VK_FORMAT_R16G16_SFLOAT
MASK_MIN = 0.9999            # north_star: visibility masks agree on >= 99.99 % of pixels


def _gbuffer_on_gpu(ctx, path, pfd, W, H, gset=0):
    g = path.gsets[gset]
    ctx.update_per_frame_ubo(pfd)
    ctx.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
    ctx.gbuffer_pass(W, H)
    return {k: ctx.image_download(g[n]) for k, n in (("depth", HP.N_DEPTH), ("normals", HP.N_NORMALS), ("motion", HP.N_MOTION))}


# ---- config 1: SVGF a-trous, 5 iterations, 1280x720 -------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 2, 3])
@pytest.mark.parametrize("inputs", ["noise", "temporal"])
def test_config1_atrous_5_iterations_720p(variant, inputs):
    W, H = 1280, 720
    sc = scenes.sponza_like(260_000, seed=1, width=W, height=H)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    pfd = seq.next(sc.camera)
    g = osc.gbuffer(pfd, W, H)
    if inputs == "noise":                       # iid values and variances: worst case for the edge-stopping weights (seed 2)
        integ = Hh.noise_integrated(H, W, seed=2)
    else:                                       # realistic variance: one oracle temporal pass over a few warm-up frames
        state = O.SvgfState(W, H)
        cam = sc.camera
        for f in range(4):
            cam.set_pose(cam.position + np.array([0.02, 0.0, 0.01]), cam.yaw, cam.pitch)
            pfd = seq.next(cam)
            g = osc.gbuffer(pfd, W, H)
            rt = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=3)["shadow_ao"]
            _, _, integ = state.run(pfd, g["normals"], g["motion"], rt, want_iters=False)
    with capi.Context(W, H) as ctx:
        ctx.set_option(capi.OPT_ATROUS_VARIANT, variant)
        ctx.update_per_frame_ubo(pfd)
        ctx.actualize_image(Hh.N_NORMALS, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"])
        a, b = ctx.upload_new_storage_image(W, H, F4), ctx.upload_new_storage_image(W, H, F4)
        ctx.storage_image_upload(a, integ)
        ctx.bind_pass_images([Hh.N_NORMALS])
        cur, ref_chain = integ, integ
        for i in range(5):
            pc = np.zeros((), T.SVGFPushConstants)
            pc["integrated_shadow_and_ao"] = (a, b)
            pc["atrous_step"] = 1 << i
            ctx.dispatch("hybrid_render_path/svgf_atrous_filter.comp", HP.groups(W), HP.groups(H), 1, pc)
            got = ctx.storage_image_download(b)
            # per-iteration parity on the SAME input, and the chained oracle for the accumulated difference
            Hh.assert_parity(got, O.svgf_atrous(pfd, g["normals"], cur, 1 << i), f"config1 {inputs} v{variant} it{i} (same input)")
            ref_chain = O.svgf_atrous(pfd, g["normals"], ref_chain, 1 << i)
            Hh.assert_parity(got, ref_chain, f"config1 {inputs} v{variant} it{i} (chained)", max_abs=2e-3)
            cur = got
            a, b = b, a


# ---- config 2: RT hard shadows 1 spp, 1920x1080, ~260k triangles -------------------------------------------------------
def test_config2_shadows_1080p_260k():
    W, H = 1920, 1080
    sc = scenes.sponza_like(260_000, seed=3, width=W, height=H)
    sc.light = camera.directional_light((-0.3, -1.0, 0.2))
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.set_option(capi.OPT_TRACE_AO, 0); ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 0); ctx.set_option(capi.OPT_AO_SPP, 1)
        path = HP.HybridRenderPath(ctx, W, H)
        assert abs(ctx.bvh_stats().n_triangles - 260_000) <= 0.02 * 260_000
        for frame in range(1, 4):               # frame_index 0 would give every pixel the same seed (Q5)
            pfd = seq.next(sc.camera)
            pfd["frame_index"] = frame
            g = _gbuffer_on_gpu(ctx, path, pfd, W, H)
            path.raytrace_pass()
            sa = ctx.image_download(HP.N_RT)
            ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=1, flags=1)["shadow_ao"]
            agree = float(np.mean(sa[..., 0] == ref[..., 0]))
            lit = g["depth"] > 0
            print(f"[config2] frame_index {frame}: shadow mask agreement {agree * 100:.4f}% over {int(lit.sum())} rays, lit fraction {float(ref[..., 0].astype(np.float32)[lit].mean()):.3f}")
            assert agree >= MASK_MIN
            Hh.classify_mask_mismatches(osc, pfd, g["depth"], g["normals"], sa[..., :1], ref[..., :1], 0, f"config2 frame_index {frame}", flags=1)
            assert np.all(sa[..., 1].astype(np.float32) == 1.0)                       # AO off: written as 1
            assert np.all(sa[~lit].astype(np.float32) == 1.0)                         # sky rule (raygen.rgen:20-24)
            path.raytrace_pass()                                                       # determinism
            assert np.array_equal(ctx.image_download(HP.N_RT).view(np.uint16), sa.view(np.uint16))


# ---- config 3: AO 4 spp + SVGF temporal accumulation and variance, 1080p, ~1M triangles ---------------------------------
def test_config3_ao4_temporal_1080p_1M():
    W, H = 1920, 1080
    sc = scenes.sponza_like(1_000_000, seed=3, width=W, height=H)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.set_option(capi.OPT_TRACE_SHADOWS, 0); ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 0); ctx.set_option(capi.OPT_AO_SPP, 4)
        path = HP.HybridRenderPath(ctx, W, H)
        pc = path.pc
        state = O.SvgfState(W, H)
        for f in range(3):
            if f:
                cam.set_pose(cam.position + np.array([0.05, 0.0, 0.0]), cam.yaw, cam.pitch)      # translating camera, 0.05 units / frame
            pfd = seq.next(cam)
            g = _gbuffer_on_gpu(ctx, path, pfd, W, H)
            path.raytrace_pass()
            sa = ctx.image_download(HP.N_RT)
            ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=4, flags=2)["shadow_ao"]
            agree = float(np.mean(sa[..., 1] == ref[..., 1]))
            q = np.unique(sa[..., 1].astype(np.float32))
            print(f"[config3] frame {f}: AO agreement {agree * 100:.4f}% (8.29 M rays), levels {q.tolist()}")
            assert agree >= MASK_MIN
            Hh.classify_mask_mismatches(osc, pfd, g["depth"], g["normals"], sa[..., 1:], ref[..., 1:], 4, f"config3 frame {f}", flags=2)
            assert set(q.tolist()) <= {0.0, 0.25, 0.5, 0.75, 1.0}                      # mean of 4 binary samples
            # temporal accumulation + variance (svgf.comp) on the GPU's own ray output vs the oracle on the same inputs
            gx, gy = HP.groups(W), HP.groups(H)
            ctx.bind_pass_images([path.gsets[0][HP.N_NORMALS], path.gsets[0][HP.N_MOTION], path.gsets[0][HP.N_DEPTH], HP.N_RT, HP.N_DENOISED])
            path.svgf_denoise_pass()
            den = ctx.image_download(HP.N_DENOISED)
            ref_den, _, ref_temporal = state.run(pfd, g["normals"], g["motion"], sa, want_iters=False)
            Hh.assert_parity(den, ref_den, f"config3 frame {f} denoised")
            mom = ctx.storage_image_download(int(pc["shadow_and_ao_moments_history"]))
            Hh.assert_parity(mom, state.image(4), f"config3 frame {f} moments")


# ---- config 4: full hybrid frame at 3840x2160, ~3M triangles, row partition ------------------------------------------------
def test_config4_full_frame_4k_3M_partitioned_equals_single():
    W, H = 3840, 2160
    sc = scenes.sponza_like(3_000_000, seed=3, width=W, height=H)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera

    def make(rt_sets):
        ctx = capi.Context(W, H)
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)          # reference defaults: 1 shadow + 2 AO + 1 reflection ray
        return ctx, HP.HybridRenderPath(ctx, W, H, rt_sets=rt_sets)
    ref_ctx, ref_path = make(1)
    ranks = [make(2) for _ in range(2)]
    ctxs, paths = [c for c, _ in ranks], [p for _, p in ranks]
    try:
        halo = 40
        MG.setup_fused_partition_inprocess(ctxs, paths, motion_halo=halo)
        for f in range(2):
            if f:
                cam.set_pose(cam.position + np.array([0.05, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)
            pfd = seq.next(cam)
            g = _gbuffer_on_gpu(ref_ctx, ref_path, pfd, W, H)
            assert MG.required_motion_halo(float(np.abs(g["motion"][..., 1].astype(np.float32)).max()), H) <= halo
            for ctx, path in zip(ctxs, paths):
                for k, n in (("depth", HP.N_DEPTH), ("normals", HP.N_NORMALS), ("motion", HP.N_MOTION)):
                    ctx.image_upload(path.gsets[0][n], g[k])
            ref_path.frame(pfd)
            for path in paths:
                path.frame(pfd, rtset=f & 1)
            want = {k: ref_ctx.image_download(n) for k, n in (("rt", HP.N_RT), ("refl", HP.N_REFL), ("den", HP.N_DENOISED))}
            for r, (ctx, path) in enumerate(zip(ctxs, paths)):
                y0, y1 = MG.band_rows(H, 2, r)
                got = {"rt": ctx.image_download(path.rt_sets[f & 1][0]), "refl": ctx.image_download(path.rt_sets[f & 1][1]), "den": ctx.image_download(HP.N_DENOISED)}
                for k in want:
                    assert np.array_equal(got[k][y0:y1].view(np.uint16), want[k][y0:y1].view(np.uint16)), f"frame {f} rank {r} image {k}"
            # the oracle on a 64-row band of the same frame (the whole 4K frame is 33 M rays)
            rows = (1048, 1112)
            ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=7, rows=rows)
            agree = float(np.mean(np.all(want["rt"][rows[0]:rows[1]] == ref["shadow_ao"][rows[0]:rows[1]], axis=-1)))
            print(f"[config4] frame {f}: partitioned == single (bit-exact); oracle band rows {rows}: mask agreement {agree * 100:.4f}%")
            assert agree >= MASK_MIN
            band = slice(rows[0], rows[1])
            full_gpu, full_ref = np.ones_like(want["rt"]), np.ones_like(want["rt"])
            full_gpu[band], full_ref[band] = want["rt"][band], ref["shadow_ao"][band]
            Hh.classify_mask_mismatches(osc, pfd, g["depth"], g["normals"], full_gpu, full_ref, 2, f"config4 frame {f} rows {rows}")
            # 4K reflections on the band: same hit point => radiance within 2e-3 relative (HDR values, fp16 ulp above 1.0 exceeds 1e-3)
            a, b = want["refl"][band].astype(np.float32), ref["reflections"][band].astype(np.float32)
            err = np.abs(a - b) / np.maximum(1.0, np.abs(b) / 2.0)
            frac_ok = float(np.mean(err.max(axis=-1) <= 2e-3))
            print(f"[config4] frame {f}: 4K reflections within tolerance on {frac_ok * 100:.3f}% of the band's pixels")
            assert frac_ok >= 0.995
            # 4K denoised on the band: the checker's SVGF pass over rows [y0 - 72, y1 + 72) of the GPU's own ray output (the five a-trous
            # iterations reach 62 rows, the variance gaussian 1 more; frame 0 has no history to reproject)
            if f == 0:
                m = 72
                sub_ = slice(rows[0] - m, rows[1] + m)
                bpfd = pfd.copy()
                bpfd["display_size"] = (W, sub_.stop - sub_.start)
                bpfd["display_size_inverse"] = (np.float32(1) / np.float32(W), np.float32(1) / np.float32(sub_.stop - sub_.start))
                st = O.SvgfState(W, sub_.stop - sub_.start)
                den_ref, _, _ = st.run(bpfd, np.ascontiguousarray(g["normals"][sub_]), np.ascontiguousarray(g["motion"][sub_]),
                                       np.ascontiguousarray(want["rt"][sub_]), want_iters=False)
                Hh.assert_parity(want["den"][band], den_ref[m:-m], f"config4 frame {f} 4K denoised, rows {rows}")
            assert np.isfinite(want["den"].astype(np.float32)).all()
    finally:
        for c in ctxs + [ref_ctx]:
            c.close()


# ---- config 5: batch of 64 independent views, round-robin over the ranks -------------------------------------------------
def test_config5_view_batch_is_partitioned_exactly_and_views_render():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in MG.views_for_rank(64, world, r))
        assert seen == list(range(64))
    W, H = 1920, 1080
    sc = scenes.sponza_like(260_000, seed=3, width=W, height=H)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera
    base = cam.position.copy()
    sums = []
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.set_option(capi.OPT_AO_SPP, 1); ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 0)
        path = HP.HybridRenderPath(ctx, W, H)
        for view in MG.views_for_rank(64, 8, 3)[:3]:                 # rank 3 of 8 renders views 3, 11, 19, ...
            cam.set_pose(base + np.array([0.9 * view / 8.0, 0.0, 0.0]), cam.yaw + 0.01 * view, cam.pitch)
            pfd = seq.next(cam)
            g = _gbuffer_on_gpu(ctx, path, pfd, W, H)
            path.frame(pfd)
            # every view against the checker on a 96-row band: masks (mismatches classified) ...
            rows = (492, 588)
            sa = ctx.image_download(HP.N_RT)
            ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=1, flags=3, rows=rows)["shadow_ao"]
            band = slice(*rows)
            agree = float(np.mean(np.all(sa[band] == ref[band], axis=-1)))
            print(f"[config5] view {view}: mask agreement {agree * 100:.4f}% on rows {rows}")
            assert agree >= MASK_MIN
            full_gpu, full_ref = np.ones_like(sa), np.ones_like(sa)
            full_gpu[band], full_ref[band] = sa[band], ref[band]
            Hh.classify_mask_mismatches(osc, pfd, g["depth"], g["normals"], full_gpu, full_ref, 1, f"config5 view {view}")
            den = ctx.image_download(HP.N_DENOISED).astype(np.float32)
            assert np.isfinite(den).all() and 0.0 <= den[..., :2].min() and den[..., :2].max() <= 1.0 + 1e-3
            sums.append(float(den[..., :2].sum()))
    assert len(set(round(s, 1) for s in sums)) == len(sums), "different views must give different frames"


# ---- the rows around the path at the size the metric is quoted on -------------------------------------------------
def test_next_rows_1080p_260k():
    """ssao.comp + ssao_blur.comp (whole frame), ssr.comp (a 96-row band: 250 march steps per pixel on the CPU) and composition.frag (whole
    frame, screen-space modes) at 1920x1080 on the 260 k-triangle scene, from the G-buffer the CUDA producer wrote, against the
    reference's own shaders compiled for the CPU (checker). The small-size tests pin the arithmetic; this is the same comparison at
    the BASELINE size, where sample radii, march lengths and REPEAT wrap-arounds are those of a real frame."""
    W, H = 1920, 1080
    sc = scenes.sponza_like(260_000, seed=1, width=W, height=H)
    seq = camera.FrameSequencer(W, H, sc.light)
    seq.next(sc.camera)
    cam = sc.camera
    cam.set_pose(cam.position + np.array([0.03, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)
    pfd = seq.next(cam)
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        path = HP.HybridRenderPath(ctx, W, H, ssao=True, composition=F4, shadow_map_size=(64, 64))
        g = _gbuffer_on_gpu(ctx, path, pfd, W, H)
        g["albedo"] = ctx.image_download(path.gsets[0][HP.N_ALBEDO])
        path.ssao_passes()
        raw, blur = ctx.image_download(HP.N_SSAO_RAW), ctx.image_download(HP.N_SSAO)
        path.ssr_pass()
        ssr = ctx.image_download(HP.N_SSR)
        # composition in the screen-space modes (shadow mode 2 = off: no shadow map on this path): consumes the SSAO and SSR images above
        zero_rt = np.zeros((H, W, 2), np.float16)
        ctx.image_upload(HP.N_RT, zero_rt)
        path.composition_pass(2, 1, 1, denoised=False)
        comp = ctx.image_download(HP.N_RENDER_OUTPUT).astype(np.float32)
    ref_raw = O.ssao(pfd, g["depth"], g["normals"], 0.75)
    Hh.assert_parity(raw, ref_raw, "next rows 1080p: ssao raw", outlier_frac=1e-4)
    Hh.assert_parity(blur, O.ssao_blur(pfd, raw), "next rows 1080p: ssao blur (of the CUDA raw image)")
    y0, y1 = 600, 696
    ref_ssr = O.ssr(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], rows=(y0, y1))
    o, r = ssr[y0:y1].astype(np.float32), ref_ssr[y0:y1].astype(np.float32)
    fo, fr = o[..., 3] > 0, r[..., 3] > 0
    both = fo & fr
    err = np.abs(o[both][:, :3] - r[both][:, :3]) / np.maximum(1.0, np.abs(r[both][:, :3]))
    exact = float(np.mean(np.all(ssr[y0:y1].view(np.uint16) == ref_ssr[y0:y1].view(np.uint16), axis=-1)))
    print(f"[parity] next rows 1080p: ssr rows {y0}..{y1}: found {fr.mean()*100:.1f}%, mask agreement {np.mean(fo == fr)*100:.4f}%, bit-exact {exact*100:.4f}%")
    assert fr.mean() > 0.02 and np.mean(fo == fr) >= 0.999 and (not both.any() or np.mean(np.all(err <= 2e-3, axis=-1)) >= 0.999)
    ref_comp = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], zero_rt, 2, 1, 1, ssao_img=blur, ssr_img=ssr,
                             refl=np.zeros((H, W, 4), np.float16), shadow_map=np.zeros((64, 64), np.float32)).astype(np.float32)
    d = np.abs(comp - ref_comp)
    bad = d > 1e-3 * np.maximum(1.0, np.abs(ref_comp))
    print(f"[parity] next rows 1080p: composition (2, 1, 1): max_abs={d.max():.3e} exact={np.mean(comp == ref_comp)*100:.2f}% beyond tolerance: {int(bad.sum())}")
    assert np.isfinite(ref_comp).all() and np.isfinite(comp).all() and not bad.any()


# ---- the workload bench.py quotes its headline on ---------------------------------------------------------------------------------------
def test_bench_workload_frames_match_the_checker_on_a_band():
    """`hybrid_frame_1080p_3Mtri` exactly as bench.py drives it — its scene (2.997 M triangles), its two camera poses, its options (shadow
    1 + AO 1 spp, no reflections, copy-free blits), G-buffers from the CUDA producer, frame indices 3, 4, ...: the shadow / AO masks of the
    first two timed frames agree with the checker on a 64-row band (every mismatching pixel re-traced by brute force and classified), and
    the denoised image of the first frame (no history yet) is within the SVGF bar on that band. The number the bench reports is computed on
    these images."""
    import bench
    wl = "hybrid_frame_1080p_3Mtri"
    W, H, _, ao_spp, refl = bench.WORKLOADS[wl]
    sc, poses = bench.make_scene(wl)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera
    rows = (520, 584)
    band = slice(*rows)
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.set_option(capi.OPT_TRACE_SHADOWS, 1); ctx.set_option(capi.OPT_TRACE_AO, 1)
        ctx.set_option(capi.OPT_AO_SPP, ao_spp); ctx.set_option(capi.OPT_TRACE_REFLECTIONS, refl)
        path = HP.HybridRenderPath(ctx, W, H, gbuffer_sets=2, blit_alias=True)
        pfds, gs = [None, None], [None, None]
        for s in (1, 0, 1):                       # bench.GpuFrameLoop: pose 1 first, so each pose's previous camera is the other pose
            cam.set_pose(*poses[s])
            pfds[s] = seq.next(cam)
            gs[s] = _gbuffer_on_gpu(ctx, path, pfds[s], W, H, gset=s)
        for k in range(2):
            s = k & 1
            pfd = pfds[s]
            pfd["frame_index"] = 3 + k
            path.frame(pfd, gset=s)
            rt, den = ctx.image_download(HP.N_RT), ctx.image_download(HP.N_DENOISED)
            g = gs[s]
            ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=ao_spp, flags=3, rows=rows)
            agree = float(np.mean(np.all(rt[band] == ref["shadow_ao"][band], axis=-1)))
            print(f"[bench workload] frame {k}: mask agreement on rows {rows}: {agree * 100:.4f}%")
            assert agree >= MASK_MIN
            full_gpu, full_ref = np.ones_like(rt), np.ones_like(rt)
            full_gpu[band], full_ref[band] = rt[band], ref["shadow_ao"][band]
            Hh.classify_mask_mismatches(osc, pfd, g["depth"], g["normals"], full_gpu, full_ref, ao_spp, f"bench workload frame {k} rows {rows}")
            if k == 0:
                m = 72
                sub_ = slice(rows[0] - m, rows[1] + m)
                bpfd = pfd.copy()
                bpfd["display_size"] = (W, sub_.stop - sub_.start)
                bpfd["display_size_inverse"] = (np.float32(1) / np.float32(W), np.float32(1) / np.float32(sub_.stop - sub_.start))
                st = O.SvgfState(W, sub_.stop - sub_.start)
                den_ref, _, _ = st.run(bpfd, np.ascontiguousarray(g["normals"][sub_]), np.ascontiguousarray(g["motion"][sub_]),
                                       np.ascontiguousarray(rt[sub_]), want_iters=False)
                Hh.assert_parity(den[band], den_ref[m:-m], f"bench workload frame {k} denoised, rows {rows}")
            assert np.isfinite(den.astype(np.float32)).all()
