"""Worker of tests/test_multi_gpu_gloo.py: one rank of the row-band SVGF driver on CPU tensors over gloo, with the
oracle as the per-band operator (test infrastructure; the product backend is multi_gpu.CabiBandBackend)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import oracle_lib as O  # noqa: E402
from vulkanhybridrenderer_b200 import multi_gpu as MG  # noqa: E402


class OracleBandBackend:
    def __init__(self, W, H):
        z4 = lambda: np.zeros((H, W, 4), np.float16)   # noqa: E731
        self.img = {"integ0": z4(), "integ1": z4(), "history": z4(), "prev_normals": z4(), "denoised": z4(),
                    "moments": np.zeros((H, W, 2), np.float16)}
        self.frame = None

    def set_frame(self, pfd, g, rt):
        self.frame = (pfd, g, rt)
        self.img["normals"] = g["normals"]

    def temporal(self, rows):
        pfd, g, rt = self.frame
        integ, mom = O.svgf_temporal(pfd, g["normals"], g["motion"], rt, self.img["prev_normals"], self.img["history"], self.img["moments"])
        self.img["integ0"][rows[0]:rows[1]] = integ[rows[0]:rows[1]]
        new_mom = self.img["moments"].copy()
        new_mom[rows[0]:rows[1]] = mom[rows[0]:rows[1]]
        self.img["moments"] = new_mom

    def atrous(self, i, rows):
        pfd, g, _ = self.frame
        out = O.svgf_atrous(pfd, g["normals"], self.img["integ0"], 1 << i)
        self.img["integ1"][rows[0]:rows[1]] = out[rows[0]:rows[1]]

    def swap_integrated(self):
        self.img["integ0"], self.img["integ1"] = self.img["integ1"], self.img["integ0"]

    def copy_rows(self, src, dst, rows):
        if not self.img[dst].flags.writeable or self.img[dst] is self.img.get("normals"):
            self.img[dst] = self.img[dst].copy()
        self.img[dst][rows[0]:rows[1]] = self.img[src][rows[0]:rows[1]]

    def tensor(self, name):
        a = self.img[name]
        return torch.from_numpy(a.view(np.int16).reshape(a.shape[0], -1))     # shares memory; gloo moves raw 16-bit words


def main():
    rank, world, port, out_dir, data = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(data, allow_pickle=True)
    frames = d["frames"]
    H, W = frames[0]["normals"].shape[:2]
    backend = OracleBandBackend(W, H)
    drv = MG.RowBandSvgf(backend, H, world, rank, motion_halo=int(d["motion_halo"]))
    y0, y1 = drv.y0, drv.y1
    outs = []
    for f in frames:
        g = {"normals": f["normals"].copy(), "motion": f["motion"]}
        backend.set_frame(f["pfd"], g, f["rt"])
        drv.run()
        outs.append(backend.img["denoised"][y0:y1].copy())
    np.save(os.path.join(out_dir, f"band_{rank}.npy"), np.stack(outs))
    np.save(os.path.join(out_dir, f"stats_{rank}.npy"), np.array([drv.x.exchanges, drv.x.bytes_sent, y0, y1]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
