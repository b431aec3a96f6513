"""Generates tests/golden/*.npz — small frozen input/output vectors for every hot-path kernel.

The reference ships no fixtures. The shader-pass outputs stored here (shadow/AO, reflections, temporal, a-trous, denoised,
SSAO, SSR) are OUTPUTS OF THE REFERENCE'S OWN SHADERS: oracle/_ref compiles the GLSL files of /root/reference for the CPU
(oracle/make_ref.py) and this script runs them in this container, where /root/reference exists; the hand-written oracle is asserted
bit-identical on the way. The G-buffer inputs, hit distances and the fully ray-traced path come from the oracle (no reference shader
produces them from arrays). tests/test_ref_pinning_cpu.py re-derives every stored shader output from the stored inputs.
Regenerate only when the oracle or the shim is deliberately changed:  python tools/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np

import helpers as Hh
import oracle_lib as O
import ref_lib as R
from vulkanhybridrenderer_b200 import camera, scenes

OUT = os.path.join(ROOT, "tests", "golden")


def same(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"oracle and oracle/_ref differ on {what}"
    return b


def main():
    W, H = 96, 64
    sc = scenes.sponza_like(6000, seed=21, width=W, height=H, n_clutter=12)
    osc = O.OracleScene(sc)
    seq = camera.FrameSequencer(W, H, sc.light)
    cam = sc.camera
    st, st_ref = O.SvgfState(W, H), R.SvgfState(W, H)
    frames = []
    for f in range(3):
        if f:
            cam.set_pose(cam.position + np.array([0.06, 0.0, 0.02]), cam.yaw + 0.004, cam.pitch)
        pfd = seq.next(cam)
        g = osc.gbuffer(pfd, W, H)
        rg = osc.raygen(pfd, g["depth"], g["normals"], want_t=True)
        rr = R.raygen(sc, osc, pfd, g["depth"], g["normals"])                    # the reference's Raytrace Pipeline
        shadow_ao, reflections = same(rg["shadow_ao"], rr["shadow_ao"], "shadow/AO"), same(rg["reflections"], rr["reflections"], "reflections")
        den, iters, temporal = (same(a, b, "SVGF") for a, b in zip(st.run(pfd, g["normals"], g["motion"], shadow_ao),
                                                                   st_ref.run(pfd, g["normals"], g["motion"], shadow_ao)))
        ssao_raw = same(O.ssao(pfd, g["depth"], g["normals"], 0.75), R.ssao(pfd, g["depth"], g["normals"], 0.75), "ssao")
        frames.append(dict(pfd=pfd, depth=g["depth"], normals=g["normals"], motion=g["motion"], albedo=g["albedo"],
                           shadow_ao=shadow_ao, reflections=reflections, refl_t=rg["refl_t"], temporal=temporal,
                           atrous=iters, denoised=den, ssao_raw=ssao_raw, ssao=same(O.ssao_blur(pfd, ssao_raw), R.ssao_blur(pfd, ssao_raw), "ssao blur")))
    flat = {}
    for i, fr in enumerate(frames):
        for k, v in fr.items():
            flat[f"f{i}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "hybrid_frames_96x64.npz"), vertices=sc.vertices, indices=sc.indices, primitives=sc.primitives, **flat)
    # stand-alone a-trous vectors on worst-case noise, every step
    pfd, normals = frames[0]["pfd"], frames[0]["normals"]
    integ = Hh.noise_integrated(H, W, seed=2)
    np.savez_compressed(os.path.join(OUT, "atrous_noise_96x64.npz"), pfd=pfd, normals=normals, integ=integ,
                        **{f"step{s}": same(O.svgf_atrous(pfd, normals, integ, s), R.svgf_atrous(pfd, normals, integ, s), "a-trous") for s in (1, 2, 3, 4, 8, 16)})
    # textured scene: G-buffer with alpha cut-outs / normal maps, textured reflections, SSR, the fully ray-traced path
    tsc = scenes.add_procedural_textures(scenes.sponza_like(6000, seed=21, width=W, height=H, n_clutter=12), size=32)
    tosc = O.OracleScene(tsc)
    tseq = camera.FrameSequencer(W, H, tsc.light)
    tseq.next(tsc.camera)
    tsc.camera.set_pose(tsc.camera.position + np.array([0.06, 0.0, 0.02]), tsc.camera.yaw + 0.004, tsc.camera.pitch)
    tpfd = tseq.next(tsc.camera)
    tg = tosc.gbuffer(tpfd, W, H)
    trg = tosc.raygen(tpfd, tg["depth"], tg["normals"], want_t=True)
    trr = R.raygen(tsc, tosc, tpfd, tg["depth"], tg["normals"])
    same(trg["shadow_ao"], trr["shadow_ao"], "textured shadow/AO"); same(trg["reflections"], trr["reflections"], "textured reflections")
    tex = {}
    for i, t in enumerate(tsc.textures):
        tex[f"tex{i}_rgba"] = t.rgba
        tex[f"tex{i}_info"] = np.array([t.format, *t.sampler], np.int32)
    np.savez_compressed(os.path.join(OUT, "textured_frame_96x64.npz"), vertices=tsc.vertices, indices=tsc.indices, primitives=tsc.primitives,
                        n_textures=len(tsc.textures), pfd=tpfd, depth=tg["depth"], normals=tg["normals"], motion=tg["motion"], albedo=tg["albedo"],
                        shadow_ao=trg["shadow_ao"], reflections=trg["reflections"], refl_t=trg["refl_t"],
                        ssr=same(O.ssr(tpfd, tg["albedo"], tg["normals"], tg["motion"], tg["depth"]),
                                 R.ssr(tpfd, tg["albedo"], tg["normals"], tg["motion"], tg["depth"]), "ssr"),
                        raytraced=tosc.raytraced(tpfd, W, H, False), raytraced_alpha=tosc.raytraced(tpfd, W, H, True), **tex)
    # RNG / sampling KATs
    import ctypes as C
    seeds = np.array([0, 1, 16221, 0xdeadbeef, 12345678], np.uint32)
    states = np.array([O.lib().vo_seed_thread(int(s)) for s in seeds], np.uint32)
    r01 = np.zeros((len(seeds), 8), np.float32)
    for i, s in enumerate(states):
        st_ = C.c_uint32(int(s))
        for k in range(8):
            r01[i, k] = O.lib().vo_random01(C.byref(st_))
    np.savez_compressed(os.path.join(OUT, "rng_kat.npz"), seeds=seeds, states=states, random01=r01)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
