"""torchrun --nproc-per-node N tools/fused_partition_parity.py [W H TRIS FRAMES]

Fused multi-GPU partition (interleaved ray blocks stored straight into the owners' images, halo rows pushed by the SVGF
kernels, flag-word stream ordering — no collective in the frame) vs the same frames rendered on ONE GPU: every rank runs
the plain single-GPU call sequence on its partition; rank 0 also renders the full frames on a second context and compares
the gathered bands bit for bit (raw shadow/AO masks, reflections and the denoised image)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from vulkanhybridrenderer_b200 import camera, capi, scenes
from vulkanhybridrenderer_b200 import hybrid_path as HP
from vulkanhybridrenderer_b200 import multi_gpu as MG


def main():
    W, H, tris, n_frames = (int(x) for x in (sys.argv[1:5] + ["1920", "1080", "260000", "4"][len(sys.argv) - 1:]))
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.sponza_like(tris, seed=3, width=W, height=H)
    y0, y1 = MG.band_rows(H, world, rank)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        def make(rt_sets):
            ctx = capi.Context(W, H, device=local, stream=stream.cuda_stream)
            ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
            ctx.set_option(capi.OPT_AO_SPP, 2)
            ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 1)
            return ctx, HP.HybridRenderPath(ctx, W, H, rt_sets=rt_sets)
        ctx, path = make(2)
        # Dry run over the camera path: the halo of history / moments rows a frame's temporal pass needs was pushed during
        # the PREVIOUS frame, so the partition carries the maximum over the whole path (same G-buffer on every rank).
        seq0 = camera.FrameSequencer(W, H, sc.light)
        pos0, yaw0, pitch0 = sc.camera.position.copy(), sc.camera.yaw, sc.camera.pitch
        halo = 4
        for f in range(n_frames):
            if f:
                sc.camera.set_pose(sc.camera.position + np.array([0.05, 0.0, 0.01]), sc.camera.yaw + 0.002, sc.camera.pitch)
            ctx.update_per_frame_ubo(seq0.next(sc.camera))
            g = path.gsets[0]
            ctx.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
            ctx.gbuffer_pass(W, H)
            halo = max(halo, MG.required_motion_halo(float(np.abs(ctx.image_download(g[HP.N_MOTION])[..., 1].astype(np.float32)).max()), H))
        assert halo <= 64, f"camera motion needs {halo} halo rows (max 64)"
        sc.camera.set_pose(pos0, yaw0, pitch0)
        MG.setup_fused_partition(ctx, path, world, rank, motion_halo=halo)
        ref = make(1) if rank == 0 else None
        seq = camera.FrameSequencer(W, H, sc.light)
        cam = sc.camera
        ok = True
        rows = lambda name, c=None: torch.as_tensor(MG._DeviceRows((c or ctx).image_info(name)[0], H, (c or ctx).image_info(name)[1] *
                                                    HP.T.FORMAT_TEXEL_BYTES[(c or ctx).image_info(name)[3]] // 2), device="cuda")
        for f in range(n_frames):
            if f:
                cam.set_pose(cam.position + np.array([0.05, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)
            pfd = seq.next(cam)
            for c, p in ([(ctx, path)] + ([ref] if ref else [])):
                # the G-buffer is an input of the path: every rank is given the full one here (it needs its ray blocks and
                # its band + halo); the producer pass runs unpartitioned
                if c is ctx:
                    c.clear_partition(); c.set_option(capi.OPT_ROW_BEGIN, 0); c.set_option(capi.OPT_ROW_END, -1)
                c.update_per_frame_ubo(pfd)
                g = p.gsets[0]
                c.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
                c.gbuffer_pass(W, H)
                if c is ctx:
                    bands = [MG.band_rows(H, world, r)[0] for r in range(world)] + [H]
                    c.set_partition(world, rank, bands, ray_block_rows=8, motion_halo=halo)
            path.frame(pfd, gset=0, rtset=f & 1)
            if ref:
                ref[1].frame(pfd)
            outs = {}
            for key, name in (("den", HP.N_DENOISED), ("rt", path.rt_sets[f & 1][0]), ("refl", path.rt_sets[f & 1][1])):
                mine = rows(name)[y0:y1].contiguous()
                parts = [torch.empty((MG.band_rows(H, world, r)[1] - MG.band_rows(H, world, r)[0], mine.shape[1]), dtype=mine.dtype, device="cuda")
                         for r in range(world)]
                stream.synchronize()
                dist.all_gather(parts, mine)
                outs[key] = torch.cat(parts).cpu().numpy()
            if rank == 0:
                rc = ref[0]
                msg = []
                for key, name in (("rt", HP.N_RT), ("refl", HP.N_REFL), ("den", HP.N_DENOISED)):
                    want = rc.image_download(name).reshape(H, -1)
                    bad = int((outs[key].view(np.uint16) != want.view(np.uint16)).sum())
                    msg.append(f"{key} {bad}")
                    if bad:
                        rows_bad = np.nonzero((outs[key].view(np.uint16) != want.view(np.uint16)).any(axis=1))[0]
                        runs = np.split(rows_bad, np.nonzero(np.diff(rows_bad) > 1)[0] + 1)
                        msg.append("rows " + " ".join(f"{r[0]}-{r[-1]}" for r in runs[:12]))
                    ok &= bad == 0
                print(f"[fused x{world}] frame {f}: mismatching halfs: " + ", ".join(msg) + f" (of {outs['den'].size} denoised)")
        if rank == 0:
            print("[fused partition] PARITY", "OK (bit-exact)" if ok else "FAILED")
    dist.barrier()
    ctx.synchronize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
