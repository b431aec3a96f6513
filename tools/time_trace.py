"""Device timing of the ray pass (and the BVH build) through the C-ABI on the synthetic scenes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from vulkanhybridrenderer_b200 import capi, scenes, camera, types as T

F4, F2 = T.VK_FORMAT_R16G16B16A16_SFLOAT, T.VK_FORMAT_R16G16_SFLOAT
GB = {"Albedo": T.VK_FORMAT_B8G8R8A8_UNORM, "World Space Normals and Object IDs": F4,
      "Motion Vectors and Metallic Roughness": F4, "Depth": T.VK_FORMAT_D32_SFLOAT}


def elapsed(ctx, a, b):
    ms = C.c_double()
    capi._check(capi.lib().vhr_get_query_elapsed_ms(ctx._h, a, b, C.byref(ms)))
    return ms.value


def main(tris=260_000, W=1920, H=1080, reps=10):
    t0 = time.time()
    sc = scenes.sponza_like(tris, width=W, height=H)
    print(f"scene {sc.num_triangles} tris generated in {time.time()-t0:.2f}s")
    seq = camera.FrameSequencer(W, H, sc.light)
    pfd = seq.next(sc.camera)
    with capi.Context(W, H) as ctx:
        t0 = time.time()
        ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
        st = ctx.bvh_stats()
        print(f"update_geometry {time.time()-t0:.3f}s wall; build {st.build_ms:.2f} ms device; wide nodes {st.n_wide_nodes}; depth {st.wide_depth}; slots in use {st.n_used_slots / max(8 * st.n_wide_nodes, 1) * 100:.1f}%; sah {st.sah_cost:.1f}")
        ctx.update_per_frame_ubo(pfd)
        ctx.set_option(capi.OPT_RAYGEN_VARIANT, int(os.environ.get('VHR_RAYGEN_VARIANT', '0')))
        for n, f in GB.items():
            ctx.actualize_image(n, f)
        ctx.actualize_image("Raytraced Shadows and Ambient Occlusion", F2)
        ctx.actualize_image("Raytraced Reflections", F4)
        capi._check(capi.lib().vhr_create_query_pool(ctx._h, 2))
        ctx.bind_pass_images(list(GB))
        ctx.gbuffer_pass(W, H)
        capi.lib().vhr_write_timestamp(ctx._h, 0)
        for _ in range(reps):
            ctx.gbuffer_pass(W, H)
        capi.lib().vhr_write_timestamp(ctx._h, 1)
        ms = elapsed(ctx, 0, 1) / reps
        depth = ctx.image_download("Depth")
        nonsky = int((depth > 0).sum())
        print(f"gbuffer (primary closest-hit): {ms:.3f} ms  {W*H/ms/1e3:.1f} Mrays/s  non-sky {nonsky/(W*H)*100:.1f}%")
        ctx.bind_pass_images(["World Space Normals and Object IDs", "Depth", "Raytraced Shadows and Ambient Occlusion", "Raytraced Reflections"])
        for name, sh, ao, rf, spp in (("shadow only", 1, 0, 0, 1), ("ao 1spp", 0, 1, 0, 1), ("ao 2spp", 0, 1, 0, 2), ("shadow+ao1", 1, 1, 0, 1),
                                      ("reflection only", 0, 0, 1, 1), ("reference (s+2ao+refl)", 1, 1, 1, 2)):
            ctx.set_option(capi.OPT_TRACE_SHADOWS, sh); ctx.set_option(capi.OPT_TRACE_AO, ao)
            ctx.set_option(capi.OPT_TRACE_REFLECTIONS, rf); ctx.set_option(capi.OPT_AO_SPP, spp)
            ctx.trace_rays(W, H)
            capi.lib().vhr_write_timestamp(ctx._h, 0)
            for _ in range(reps):
                ctx.trace_rays(W, H)
            capi.lib().vhr_write_timestamp(ctx._h, 1)
            ms = elapsed(ctx, 0, 1) / reps
            rays = nonsky * (sh + ao * spp + rf)
            print(f"{name:24s}: {ms:8.3f} ms  {rays/ms/1e3:9.1f} Mrays/s")
        sa = ctx.image_download("Raytraced Shadows and Ambient Occlusion").astype(np.float32)
        print("lit fraction", sa[..., 0].mean(), "ao mean", sa[..., 1].mean())


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:]))
