"""torchrun --nproc-per-node N tools/rowband_parity.py [W H TRIS FRAMES]

Row-band split over N GPUs (NCCL halo exchange) vs the full frame rendered on ONE GPU: every rank renders its band of
`FRAMES` consecutive frames; rank 0 also renders the full frames on a second context and compares the gathered bands
bit for bit (raw shadow/AO masks and the denoised image)."""
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
    W, H, tris, n_frames = (int(x) for x in (sys.argv[1:5] + ["1920", "1080", "260000", "3"][len(sys.argv) - 1:]))
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = scenes.sponza_like(tris, seed=3, width=W, height=H)
    y0, y1 = MG.band_rows(H, world, rank)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        def make(full):
            ctx = capi.Context(W, H, device=local, stream=stream.cuda_stream)
            ctx.update_geometry(sc.vertices, sc.indices, sc.primitives)
            ctx.set_option(capi.OPT_AO_SPP, 1)
            ctx.set_option(capi.OPT_TRACE_REFLECTIONS, 0)
            return ctx, HP.HybridRenderPath(ctx, W, H)
        ctx, path = make(False)
        backend = MG.CabiBandBackend(ctx, path)
        drv = MG.RowBandSvgf(backend, H, world, rank, motion_halo=8)
        ref = make(True) if rank == 0 else None
        seq = camera.FrameSequencer(W, H, sc.light)
        cam = sc.camera
        ok = True
        for f in range(n_frames):
            if f:
                cam.set_pose(cam.position + np.array([0.05, 0.0, 0.01]), cam.yaw + 0.002, cam.pitch)
            pfd = seq.next(cam)
            for c, p in ([(ctx, path)] + ([ref] if ref else [])):
                full = c is not ctx
                c.update_per_frame_ubo(pfd)
                g0, g1 = (0, H) if full else (max(0, y0 - MG.GBUFFER_HALO), min(H, y1 + MG.GBUFFER_HALO))
                c.set_option(capi.OPT_ROW_BEGIN, g0); c.set_option(capi.OPT_ROW_END, g1)
                g = p.gsets[0]
                c.bind_pass_images([g[HP.N_ALBEDO], g[HP.N_NORMALS], g[HP.N_MOTION], g[HP.N_DEPTH]])
                c.gbuffer_pass(W, H)
                if full:
                    p.raytrace_pass(); p.svgf_denoise_pass()
            backend.trace((y0, y1))
            drv.run()
            den = backend.tensor("denoised")[y0:y1].contiguous()
            rt = backend.tensor("rt")[y0:y1].contiguous()
            stream.synchronize()
            dens = [torch.empty((MG.band_rows(H, world, r)[1] - MG.band_rows(H, world, r)[0], den.shape[1]), dtype=den.dtype, device="cuda") for r in range(world)]
            rts = [torch.empty((d.shape[0], rt.shape[1]), dtype=rt.dtype, device="cuda") for d in dens]
            dist.all_gather(dens, den); dist.all_gather(rts, rt)
            if rank == 0:
                rc = ref[0]
                want_den = rc.image_download(HP.N_DENOISED).reshape(H, -1)
                want_rt = rc.image_download(HP.N_RT).reshape(H, -1)
                got_den = torch.cat(dens).cpu().numpy(); got_rt = torch.cat(rts).cpu().numpy()
                e_rt = int((got_rt.view(np.uint16) != want_rt.view(np.uint16)).sum())
                e_den = int((got_den.view(np.uint16) != want_den.view(np.uint16)).sum())
                print(f"[rowband x{world}] frame {f}: mask texel mismatches {e_rt}, denoised half mismatches {e_den} of {got_den.size}; "
                      f"halo bytes sent by rank 0 so far {drv.x.bytes_sent}")
                ok &= e_rt == 0 and e_den == 0
        if rank == 0:
            print("[rowband] PARITY", "OK (bit-exact)" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
