"""CPU, world_size > 1 over gloo: the row-band split of the SVGF pass (multi_gpu.RowBandSvgf + HaloExchanger) stitched
from N ranks equals the single-process pass bit for bit. The per-band operator is the oracle; what is under test is the
product's band planning, halo widths and exchange order (the N > 1 host logic)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers as Hh
import oracle_lib as O
from vulkanhybridrenderer_b200 import multi_gpu as MG

HERE = os.path.dirname(os.path.abspath(__file__))


def test_band_plan_and_views():
    for H, N in ((2160, 8), (1080, 8), (97, 3), (5, 8)):
        rows = [MG.band_rows(H, N, r) for r in range(N)]
        assert rows[0][0] == 0 and rows[-1][1] == H
        assert all(rows[i][1] == rows[i + 1][0] for i in range(N - 1))
        sizes = [b - a for a, b in rows]
        assert max(sizes) - min(sizes) <= 1
    assert MG.views_for_rank(64, 8, 3) == list(range(3, 64, 8))
    assert sorted(sum((MG.views_for_rank(10, 4, r) for r in range(4)), [])) == list(range(10))
    assert MG.ATROUS_HALO == [2, 4, 8, 16, 32] and sum(MG.ATROUS_HALO) == 62     # 62 rows cumulative


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return str(p)


@pytest.mark.parametrize("world", [2, 3])      # 3 ranks: bands of 22/21/21 rows < the 32-row halo => multi-hop exchange
def test_row_band_svgf_equals_single_process(world, tmp_path):
    W, H = 96, 64
    sc, osc, frames = Hh.scene_and_gbuffer(W, H, tris=6000, moving=True)
    rng = np.random.default_rng(11)
    data = []
    for pfd, g in frames:
        ref = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=1, flags=3)
        rt = ref["shadow_ao"]
        # add noise so every texel matters
        rt = np.where(rng.uniform(size=rt.shape) < 0.15, rng.integers(0, 2, rt.shape), rt).astype(np.float16)
        data.append(dict(pfd=pfd, normals=g["normals"], motion=g["motion"], rt=rt))
    motion_px = max(float(np.abs(f["motion"][..., 1].astype(np.float32)).max() * H) for f in data[1:])
    motion_halo = int(np.ceil(motion_px)) + 2
    path = tmp_path / "frames.npz"
    np.savez(path, frames=np.array(data, dtype=object), motion_halo=motion_halo)
    # single-process reference
    st = O.SvgfState(W, H)
    want = [st.run(f["pfd"], f["normals"], f["motion"], f["rt"], want_iters=False)[0] for f in data]
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "mgpu_worker.py"), str(r), str(world), port, str(tmp_path), str(path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]
    got = np.concatenate([np.load(tmp_path / f"band_{r}.npy") for r in range(world)], axis=1)
    for f in range(len(data)):
        np.testing.assert_array_equal(got[f], want[f], err_msg=f"frame {f}")
    stats = [np.load(tmp_path / f"stats_{r}.npy") for r in range(world)]
    assert all(s[0] == 6 * len(data) for s in stats)      # 6 grouped exchanges per frame: temporal-out, moments, it0..it3 outputs
    print(f"[gloo x{world}] motion halo {motion_halo} rows, bytes sent per rank per frame: {[int(s[1]) // len(data) for s in stats]}")
