import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from vulkanhybridrenderer_b200 import capi, scenes, camera, hybrid_path as HP
W, H = 64, 48
sc = scenes.sponza_like(3000, seed=5, width=W, height=H, n_clutter=5)
seq = camera.FrameSequencer(W, H, sc.light)
pfd = seq.next(sc.camera)
with capi.Context(W, H) as ctx:
    ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
    path = HP.HybridRenderPath(ctx, W, H)
    ctx.update_per_frame_ubo(pfd)
    g = path.gsets[0]
    ctx.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
    ctx.gbuffer_pass(W, H)
    ctx.synchronize(); print("gbuffer ok", flush=True)
    for v in (0, 1):
        ctx.set_option(capi.OPT_RAYGEN_VARIANT, v)
        path.raytrace_pass()
        ctx.synchronize()
        print("variant", v, "ok", ctx.image_download(HP.N_RT).astype(np.float32).mean(), flush=True)
