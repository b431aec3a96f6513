"""GPU parity: LBVH build + traversal + the ray-traced pass (through the C-ABI) vs the CPU oracle.

The reference's hit/no-hit decision is made inside the Vulkan driver ("parity unpinned"); the bar is
agreement with the oracle's exact (double precision, watertight) intersection on the same world-space triangle
soup: visibility masks agree on >= 99.99 % of pixels, mismatches only at grazing / epsilon cases.
"""
import numpy as np
import pytest

import helpers as Hh
import checker as O
from vulkanhybridrenderer_b200 import capi, scenes, camera
from vulkanhybridrenderer_b200 import types as T

pytestmark = pytest.mark.gpu
F4, F2 = T.VK_FORMAT_R16G16B16A16_SFLOAT, T.VK_FORMAT_R16G16_SFLOAT
MASK_AGREEMENT_MIN = 0.9999


def _random_rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(-19, 19, n), rng.uniform(0.1, 9, n), rng.uniform(-7.5, 7.5, n)], -1)
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, :3] = o; rays[:, 3] = 0.01; rays[:, 4:7] = d; rays[:, 7] = rng.choice([5.0, 10000.0], n)
    # a few axis-parallel rays (zero direction components exercise the slab test's inf handling)
    rays[:64, 4:7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 64)] * rng.choice([-1.0, 1.0], (64, 1))
    return rays


def test_bvh_build_and_explicit_rays():
    W, H = 64, 64
    sc = scenes.sponza_like(60_000, seed=11, width=W, height=H, n_clutter=60)
    osc = O.OracleScene(sc)
    rays = _random_rays(sc, 20000, 1)
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        st = ctx.bvh_stats()
        print(f"[bvh] tris={st.n_triangles} wide={st.n_wide_nodes} sah={st.sah_cost:.1f} build={st.build_ms:.2f}ms depth={st.wide_depth}")
        assert st.n_triangles == sc.num_triangles == osc.num_triangles
        assert 0 < st.n_wide_nodes < st.n_triangles
        any_t, _, _ = ctx.trace_explicit(rays, any_hit=True)
        cl_t, cl_ids, cl_uv = ctx.trace_explicit(rays, any_hit=False)
    ref_any = np.array([osc.trace_any(r[:3], r[4:7], r[3], r[7]) for r in rays])
    agree = np.mean((any_t > 0.5) == ref_any)
    print(f"[rays] any-hit agreement {agree*100:.4f}% hit rate {ref_any.mean():.3f}")
    bad = np.nonzero((any_t > 0.5) != ref_any)[0]
    if len(bad):
        tris = Hh.world_triangles(sc)
        real = 0
        for i in bad[:40]:
            r = rays[i]
            bf_any, bf_t, margin = Hh.brute_force_hits(tris, r[:3], r[4:7], r[3], r[7])
            grazing = margin < 1e-5
            real += not grazing
            print(f"  ray {i}: gpu={int(any_t[i] > 0.5)} oracle={int(ref_any[i])} brute={int(bf_any)} t={bf_t:.6f} margin={margin:.2e} "
                  f"{'grazing' if grazing else 'REAL'} o={r[:3]} d={r[4:7]} tmax={r[7]}")
        assert real == 0, f"{real} mismatches are not epsilon/grazing cases"
    assert agree >= MASK_AGREEMENT_MIN
    n_bad = 0
    for i, r in enumerate(rays[:4000]):
        ref = osc.trace_closest(r[:3], r[4:7], r[3], r[7])
        if ref is None:
            n_bad += cl_t[i] >= 0
            continue
        (t, u, v), (g, p) = ref
        if cl_t[i] < 0 or abs(cl_t[i] - t) > 1e-4 * max(1.0, t):
            n_bad += 1
            continue
        if (cl_ids[i, 0], cl_ids[i, 1]) == (g, p):
            assert abs(cl_uv[i, 0] - u) < 1e-3 and abs(cl_uv[i, 1] - v) < 1e-3
    print(f"[rays] closest-hit mismatches {n_bad} / 4000")
    assert n_bad <= 2
    # any-hit and closest-hit must be consistent with each other on the GPU
    assert np.array_equal(any_t > 0.5, cl_t >= 0)


def test_geometry_edge_cases():
    W, H = 32, 32
    with capi.Context(W, H) as ctx:
        # empty scene: traversal must miss
        ctx.update_geometry(np.zeros(0, T.Vertex), np.zeros(0, np.uint32), np.zeros(0, T.Primitive))
        assert ctx.bvh_stats().n_triangles == 0
        t, _, _ = ctx.trace_explicit(np.array([[0, 0, 0, 0.01, 0, 0, 1, 100]], np.float32), any_hit=True)
        assert t[0] == 0.0
        # one triangle, two triangles, three, and a pile of duplicates + a degenerate triangle
        v = np.zeros(4, T.Vertex)
        v["pos"] = [(-1, -1, 5), (1, -1, 5), (0, 1, 5), (0, 0, 5)]
        v["normal"] = (0, 0, -1)
        prim = np.zeros(1, T.Primitive)
        prim["transform"] = np.eye(4)
        prim["material"]["base_color"] = (1, 1, 1, 1)
        prim["material"]["base_color_texture"] = -1
        prim["material"]["metallic_roughness_texture"] = -1
        prim["material"]["normal_map"] = -1
        ray_hit = np.array([[0, 0, 0, 0.01, 0, 0, 1, 100]], np.float32)
        ray_miss = np.array([[3, 0, 0, 0.01, 0, 0, 1, 100]], np.float32)
        ray_short = np.array([[0, 0, 0, 0.01, 0, 0, 1, 4.5]], np.float32)
        for idx in ([0, 1, 2], [0, 1, 2, 0, 1, 2], [0, 1, 2] * 3, [0, 1, 2] * 40 + [3, 3, 3]):
            prim["index_count"] = len(idx)
            ctx.update_geometry(v, np.array(idx, np.uint32), prim)
            assert ctx.bvh_stats().n_triangles == len(idx) // 3
            t, ids, uv = ctx.trace_explicit(ray_hit, any_hit=False)
            assert abs(t[0] - 5.0) < 1e-5 and ids[0, 0] == 0
            assert ctx.trace_explicit(ray_hit, any_hit=True)[0][0] == 1.0
            assert ctx.trace_explicit(ray_miss, any_hit=True)[0][0] == 0.0
            assert ctx.trace_explicit(ray_short, any_hit=True)[0][0] == 0.0   # tMax is exclusive and in ray units
        # out-of-range index is rejected on the host (the reference would read out of bounds)
        prim["index_count"] = 3
        with pytest.raises(capi.VhrError):
            ctx.update_geometry(v, np.array([0, 1, 9], np.uint32), prim)
        # a shared edge must be watertight: rays through the common edge of two triangles hit one of them
        v2 = np.zeros(4, T.Vertex)
        v2["pos"] = [(-1, -1, 5), (1, -1, 5), (1, 1, 5), (-1, 1, 5)]
        prim["index_count"] = 6
        ctx.update_geometry(v2, np.array([0, 1, 2, 0, 2, 3], np.uint32), prim)
        n = 2001
        s = np.linspace(-0.999, 0.999, n, dtype=np.float32)
        rays = np.zeros((n, 8), np.float32)
        rays[:, 3] = 0.01; rays[:, 7] = 100
        rays[:, 4] = s * 0.2; rays[:, 5] = s * 0.2; rays[:, 6] = 1.0   # crosses the diagonal x == y exactly
        hits = ctx.trace_explicit(rays, any_hit=True)[0]
        assert hits.min() == 1.0, f"{int((hits == 0).sum())} rays leaked through a shared edge"


@pytest.mark.parametrize("size,tris,ao_spp", [((320, 184), 60_000, 2), ((203, 117), 20_000, 4)])
def test_raygen_masks_and_reflections(size, tris, ao_spp):
    W, H = size
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=tris, moving=True)
    pfd, g = frames[1]
    ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=ao_spp, flags=7, want_t=True)
    with capi.Context(W, H) as ctx:
        ctx.set_option(capi.OPT_AO_SPP, ao_spp)
        ctx.set_option(capi.OPT_DEBUG_REFLECTION_T, 1)
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.update_per_frame_ubo(pfd)
        ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
        ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
        ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
        ctx.trace_rays(W, H)
        sa = ctx.image_download(Hh.N_RT)
        refl = ctx.image_download(Hh.N_REFL)
        rt = ctx.download_reflection_t()
    shadow_agree = np.mean(sa[..., 0] == ref["shadow_ao"][..., 0])
    ao_agree = np.mean(sa[..., 1] == ref["shadow_ao"][..., 1])
    print(f"[raygen {W}x{H} {tris} tris ao_spp={ao_spp}] shadow agreement {shadow_agree*100:.4f}%  ao agreement {ao_agree*100:.4f}%"
          f"  (lit {ref['shadow_ao'][..., 0].astype(np.float32).mean():.3f}, ao {ref['shadow_ao'][..., 1].astype(np.float32).mean():.3f})")
    assert shadow_agree >= MASK_AGREEMENT_MIN
    assert ao_agree >= MASK_AGREEMENT_MIN
    Hh.classify_mask_mismatches(osc, pfd, g["depth"], g["normals"], sa, ref["shadow_ao"], ao_spp, f"raygen {W}x{H} ao_spp={ao_spp}")
    # reflection hit distance: same hit/miss classification and t within 1e-3 relative on >= 99.9 % of pixels
    ref_t = ref["refl_t"]
    same_class = (rt >= 0) == (ref_t >= 0)
    both = (rt >= 0) & (ref_t >= 0)
    t_ok = np.abs(rt - ref_t)[both] <= 1e-3 * np.maximum(1.0, ref_t[both])
    print(f"[reflection] class agreement {same_class.mean()*100:.4f}%  t within tol {t_ok.mean()*100:.4f}% of {both.sum()} hits")
    assert same_class.mean() >= 0.9995 and t_ok.mean() >= 0.999
    # radiance: HDR values (intensity 30) -> tolerance 1e-3 absolute or 2e-3 relative, on pixels that hit the same point
    good = both & (np.abs(rt - ref_t) <= 1e-3 * np.maximum(1.0, ref_t))
    a = refl.astype(np.float32)[good]; b = ref["reflections"].astype(np.float32)[good]
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b) / 2.0)
    frac_ok = np.mean(err.max(axis=-1) <= 2e-3)
    print(f"[reflection] radiance within tol on {frac_ok*100:.3f}% of matched hits, psnr {Hh.psnr(a, b, peak=max(1.0, float(b.max()))):.1f} dB")
    assert frac_ok >= 0.995
    # sky pixels: (1,1) and zero reflection (raygen.rgen:20-24)
    sky = g["depth"] == 0
    if sky.any():
        assert np.all(sa[sky].astype(np.float32) == 1.0) and np.all(refl[sky].astype(np.float32) == 0.0)


def test_raygen_option_flags_and_row_band():
    W, H = 128, 72
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=20_000)
    pfd, g = frames[0]
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.update_per_frame_ubo(pfd)
        ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
        ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
        ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
        ctx.trace_rays(W, H)
        full = ctx.image_download(Hh.N_RT).copy()
        # shadows only: AO channel becomes 1, shadow channel unchanged (the RNG stream is still consumed)
        ctx.set_option(capi.OPT_TRACE_AO, 0); ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 0)
        ctx.trace_rays(W, H)
        so = ctx.image_download(Hh.N_RT)
        assert np.array_equal(so[..., 0], full[..., 0]) and np.all(so[..., 1].astype(np.float32) == 1.0)
        assert np.all(ctx.image_download(Hh.N_REFL).astype(np.float32) == 0.0)
        # row band: only rows [20, 50) are written
        ctx.set_option(capi.OPT_TRACE_AO, 1)
        ctx.image_upload(Hh.N_RT, np.full((H, W, 2), 7.0, np.float16))
        ctx.set_option(capi.OPT_ROW_BEGIN, 20); ctx.set_option(capi.OPT_ROW_END, 50)
        ctx.trace_rays(W, H)
        band = ctx.image_download(Hh.N_RT)
        assert np.array_equal(band[20:50], full[20:50])
        assert np.all(band[:20].astype(np.float32) == 7.0) and np.all(band[50:].astype(np.float32) == 7.0)
        # wrong pipeline name / launch size are rejected
        with pytest.raises(capi.VhrError):
            ctx.trace_rays(W, H, pipeline="Some Other Pipeline")
        with pytest.raises(capi.VhrError):
            ctx.trace_rays(W + 8, H)


def test_gbuffer_pass_matches_oracle_encodings():
    W, H = 256, 144
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=60_000, moving=True)
    pfd, g = frames[1]
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.update_per_frame_ubo(pfd)
        for name, fmt in Hh.GBUF_IMAGES.items():
            ctx.actualize_image(name, fmt)
        ctx.bind_pass_images(list(Hh.GBUF_IMAGES))
        ctx.gbuffer_pass(W, H)
        got = {k: ctx.image_download(n) for k, n in (("albedo", "Albedo"), ("normals", Hh.N_NORMALS), ("motion", Hh.N_MOTION), ("depth", Hh.N_DEPTH))}
    same_obj = got["normals"][..., 3] == g["normals"][..., 3]
    print(f"[gbuffer] same object id on {same_obj.mean()*100:.3f}% of pixels; sky agreement {np.mean((got['depth'] == 0) == (g['depth'] == 0))*100:.3f}%")
    assert same_obj.mean() >= 0.999
    m = same_obj & (g["depth"] > 0)
    rel = np.abs(got["depth"][m] - g["depth"][m]) / g["depth"][m]
    assert np.quantile(rel, 0.999) < 1e-4
    dn = np.abs(got["normals"][..., :3].astype(np.float32) - g["normals"][..., :3].astype(np.float32))[m]
    assert np.quantile(dn.max(axis=-1), 0.999) <= 2e-3
    dm = np.abs(got["motion"].astype(np.float32) - g["motion"].astype(np.float32))[m]
    assert np.quantile(dm.max(axis=-1), 0.999) <= 1e-3
    assert np.mean(np.all(got["albedo"][m] == g["albedo"][m], axis=-1)) > 0.999


def test_persistent_raygen_variant_matches_per_pixel_kernel():
    """VHR_OPT_RAYGEN_VARIANT 1 (persistent warps, pixel queue) must produce the same images as variant 0."""
    W, H = 203, 117
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=20_000, moving=True)
    pfd, g = frames[1]
    outs = []
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.update_per_frame_ubo(pfd)
        ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
        ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
        ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
        for variant in (0, 1):
            ctx.set_option(capi.OPT_RAYGEN_VARIANT, variant)
            ctx.set_option(capi.OPT_ROW_BEGIN, 9); ctx.set_option(capi.OPT_ROW_END, 101)     # ragged band
            ctx.image_upload(Hh.N_RT, np.zeros((H, W, 2), np.float16)); ctx.image_upload(Hh.N_REFL, np.zeros((H, W, 4), np.float16))
            ctx.trace_rays(W, H)
            outs.append((ctx.image_download(Hh.N_RT), ctx.image_download(Hh.N_REFL)))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1].view(np.uint16), outs[1][1].view(np.uint16))
    assert outs[0][0][:9].astype(np.float32).sum() == 0 and outs[0][0][101:].astype(np.float32).sum() == 0   # rows outside the band untouched


@pytest.mark.parametrize("ao_spp,shadows,ao,refl,band", [(1, 1, 1, 0, None), (2, 1, 1, 1, None), (4, 1, 1, 0, (9, 101)), (2, 0, 1, 1, None), (1, 1, 0, 0, (16, 64)), (2, 1, 1, 1, (5, 117))])
def test_ray_queue_variant_matches_per_pixel_kernel(ao_spp, shadows, ao, refl, band):
    """VHR_OPT_RAYGEN_VARIANT 9 (two-phase kernel: rays generated into shared memory, traversal with ray-level lane refill; the reflection ray, if on,
    through the per-pixel kernel afterwards) traces exactly the rays of variant 0: every image is bit-identical, also on a ragged band of rows and
    with ray kinds switched off."""
    W, H = 203, 117
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=20_000, moving=True)
    pfd, g = frames[1]
    outs = []
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.update_per_frame_ubo(pfd)
        ctx.set_option(capi.OPT_DEBUG_REFLECTION_T, 1)
        ctx.set_option(capi.OPT_AO_SPP, ao_spp); ctx.set_option(capi.OPT_TRACE_SHADOWS, shadows); ctx.set_option(capi.OPT_TRACE_AO, ao)
        ctx.set_option(capi.OPT_TRACE_REFLECTIONS, refl)
        if band:
            ctx.set_option(capi.OPT_ROW_BEGIN, band[0]); ctx.set_option(capi.OPT_ROW_END, band[1])
        ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
        ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
        ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
        for v in (0, 9):
            ctx.set_option(capi.OPT_RAYGEN_VARIANT, v)
            ctx.image_upload(Hh.N_RT, np.full((H, W, 2), 0.25, np.float16)); ctx.image_upload(Hh.N_REFL, np.full((H, W, 4), 0.25, np.float16))
            ctx.trace_rays(W, H)
            outs.append((ctx.image_download(Hh.N_RT), ctx.image_download(Hh.N_REFL), ctx.download_reflection_t()))
    assert np.array_equal(outs[0][0].view(np.uint16), outs[1][0].view(np.uint16)), "shadow / AO image differs"
    assert np.array_equal(outs[0][1].view(np.uint16), outs[1][1].view(np.uint16)), "reflection image differs"
    if band is None:
        np.testing.assert_array_equal(outs[0][2], outs[1][2])
    if shadows and ao:
        rows = slice(*band) if band else slice(None)
        assert len(np.unique(outs[1][0][rows, :, 1].astype(np.float32))) >= 2 and len(np.unique(outs[1][0][rows, :, 0].astype(np.float32))) == 2


@pytest.mark.parametrize("variant", [2, 3, 4, 6, 7, 8, 10, 11, 12, 13, 14, 15])
def test_raygen_kernel_variants_match_default(variant):
    """VHR_OPT_RAYGEN_VARIANT 2 / 3 (other register budgets), 4 (postponed leaves: the warp runs the triangle block together) trace the
    same rays against the same tree: shadow / AO masks are identical; the closest-hit ray may report another triangle only on exact ties."""
    W, H = 203, 117
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=20_000, moving=True)
    pfd, g = frames[1]
    outs = []
    with capi.Context(W, H) as ctx:
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        ctx.update_per_frame_ubo(pfd)
        ctx.set_option(capi.OPT_DEBUG_REFLECTION_T, 1)
        ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
        ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
        ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
        ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
        for v in (0, variant):
            ctx.set_option(capi.OPT_RAYGEN_VARIANT, v)
            ctx.image_upload(Hh.N_RT, np.zeros((H, W, 2), np.float16)); ctx.image_upload(Hh.N_REFL, np.zeros((H, W, 4), np.float16))
            ctx.trace_rays(W, H)
            outs.append((ctx.image_download(Hh.N_RT), ctx.image_download(Hh.N_REFL), ctx.download_reflection_t()))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][2], outs[1][2])                                  # same closest-hit distance everywhere
    same = np.all(outs[0][1].view(np.uint16) == outs[1][1].view(np.uint16), axis=-1)
    assert same.mean() >= 0.9999, same.mean()


def test_ploc_hierarchy_gives_the_same_images(monkeypatch):
    """The binary hierarchy is built by parallel locally-ordered clustering (default, VHR_BVH_BUILDER=1) or as a radix tree
    (VHR_BVH_BUILDER=0). Visibility does not depend on the tree: masks and hit distances must be identical, the statistics must
    describe a complete tree, and two PLOC builds of the same scene must give the same tree (no atomics in the numbering)."""
    W, H = 203, 117
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=20_000, moving=True)
    pfd, g = frames[1]
    outs, stats = [], []
    for builder in ("0", "1", "1"):
        monkeypatch.setenv("VHR_BVH_BUILDER", builder)
        with capi.Context(W, H) as ctx:
            ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
            st = ctx.bvh_stats()
            stats.append((st.n_triangles, st.n_wide_nodes, st.sah_cost))
            ctx.update_per_frame_ubo(pfd)
            ctx.set_option(capi.OPT_DEBUG_REFLECTION_T, 1)
            ctx.actualize_image(Hh.N_NORMALS, F4); ctx.actualize_image(Hh.N_DEPTH, T.VK_FORMAT_D32_SFLOAT)
            ctx.actualize_image(Hh.N_RT, F2); ctx.actualize_image(Hh.N_REFL, F4)
            ctx.image_upload(Hh.N_NORMALS, g["normals"]); ctx.image_upload(Hh.N_DEPTH, g["depth"])
            ctx.bind_pass_images([Hh.N_NORMALS, Hh.N_DEPTH, Hh.N_RT, Hh.N_REFL])
            ctx.trace_rays(W, H)
            outs.append((ctx.image_download(Hh.N_RT), ctx.image_download(Hh.N_REFL), ctx.download_reflection_t()))
            # tiny inputs through the same builder
            v = np.zeros(4, T.Vertex)
            v["pos"] = [(-1, -1, 5), (1, -1, 5), (0, 1, 5), (0, 0, 5)]
            prim = np.zeros(1, T.Primitive)
            prim["transform"] = np.eye(4)
            for key in ("base_color_texture", "metallic_roughness_texture", "normal_map"):
                prim["material"][key] = -1
            for idx in ([0, 1, 2], [0, 1, 2, 0, 1, 3], [0, 1, 2] * 9):
                prim["index_count"] = len(idx)
                ctx.update_geometry(v, np.array(idx, np.uint32), prim)
                t, _, _ = ctx.trace_explicit(np.array([[0, -0.5, 0, 0.01, 0, 0, 1, 100], [3, 0, 0, 0.01, 0, 0, 1, 100]], np.float32), any_hit=False)
                assert abs(t[0] - 5.0) < 1e-5 and t[1] == -1.0
    print(f"[ploc] radix tree: {stats[0]}  ploc: {stats[1]}")
    assert stats[0][0] == stats[1][0]
    assert stats[1][:2] == stats[2][:2] and abs(stats[1][2] - stats[2][2]) < 1e-3      # the SAH figure is summed with float atomics
    for a, b in zip(outs[1], outs[2]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    assert np.mean(outs[0][2] == outs[1][2]) >= 0.9999
    same = np.all(outs[0][1].view(np.uint16) == outs[1][1].view(np.uint16), axis=-1)
    assert same.mean() >= 0.999, same.mean()


def test_strung_out_geometry_falls_back_to_the_radix_tree(monkeypatch):
    """Tiny triangles on a line with steadily growing gaps: every cluster's nearest neighbour is its left one, so PLOC merges one pair per
    round and would produce a tree as deep as the line is long. update_geometry must not fail — it rebuilds with the radix tree — and
    the rays must still find their triangles."""
    n = 3000
    x = 0.002 * np.arange(n, dtype=np.float64) ** 2
    v = np.zeros(3 * n, T.Vertex)
    pos = np.zeros((n, 3, 3), np.float32)
    pos[:, :, 2] = 5.0
    pos[:, 0, 0] = x; pos[:, 1, 0] = x + 0.01; pos[:, 2, 0] = x
    pos[:, 2, 1] = 0.01
    v["pos"] = pos.reshape(-1, 3)
    prim = np.zeros(1, T.Primitive)
    prim["transform"] = np.eye(4)
    prim["index_count"] = 3 * n
    for key in ("base_color_texture", "metallic_roughness_texture", "normal_map"):
        prim["material"][key] = -1
    idx = np.arange(3 * n, dtype=np.uint32)
    pick = np.array([0, 1, 7, 500, 1499, 2998, 2999])
    rays = np.zeros((len(pick) + 1, 8), np.float32)
    rays[:-1, 0] = x[pick] + 0.002; rays[:-1, 1] = 0.002
    rays[-1, 0] = x[10] + 0.02; rays[-1, 1] = 0.002           # between two triangles: a miss
    rays[:, 3] = 0.01; rays[:, 6] = 1.0; rays[:, 7] = 100.0
    stats = []
    for builder in ("1", "0"):
        monkeypatch.setenv("VHR_BVH_BUILDER", builder)
        with capi.Context(64, 64) as ctx:
            ctx.update_geometry(v, idx, prim)
            st = ctx.bvh_stats()
            stats.append((st.n_triangles, st.n_wide_nodes, st.wide_depth))
            t, ids, _ = ctx.trace_explicit(rays, any_hit=False)
            np.testing.assert_allclose(t[:-1], 5.0, atol=1e-5)
            assert t[-1] == -1.0
            t_any, _, _ = ctx.trace_explicit(rays, any_hit=True)
            assert (t_any[:-1] == 1.0).all() and t_any[-1] == 0.0          # any-hit: 1 = occluded
    assert stats[0] == stats[1] and stats[0][0] == n and stats[0][2] <= 40, stats
