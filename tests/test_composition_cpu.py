"""Pins the composition oracle (oracle/oracle_composition.cpp) against an independent numpy restatement written from
composition.frag:60-161 and common.glsl:116-150, on an oracle G-buffer of the synthetic scene (CPU only)."""
import numpy as np
import pytest

import helpers as Hh
import oracle_lib as O
from vulkanhybridrenderer_b200 import types as T

f32 = np.float32


def _normalize(v):
    l = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]).astype(f32)
    with np.errstate(all="ignore"):
        return (v / l[..., None]).astype(f32)


def _dot(a, b):
    return ((a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]).astype(f32)


def numpy_composition(pfd, g, rt, shadow_mode, ao_mode, reflection_mode, ssao=None, refl=None):
    """Straight transcription of composition.frag for shadow_mode in {0, 2} (no shadow map), float32 throughout."""
    H, W = g["depth"].shape
    xs = (np.arange(W, dtype=f32) + f32(0.5)) / f32(W)
    ys = (np.arange(H, dtype=f32) + f32(0.5)) / f32(H)
    u, v = np.meshgrid(xs, ys)
    a8 = g["albedo"].astype(f32)
    albedo = np.stack([a8[..., 2] / f32(255), a8[..., 1] / f32(255), a8[..., 0] / f32(255)], -1).astype(f32)
    m = np.asarray(pfd["camera_viewproj_inverse"], f32).reshape(4, 4)       # m[c][r]
    ndc = [u * f32(2) - f32(1), v * f32(2) - f32(1), g["depth"].astype(f32), np.ones((H, W), f32)]
    with np.errstate(all="ignore"):
        p4 = [(((m[0, r] * ndc[0] + m[1, r] * ndc[1]) + m[2, r] * ndc[2]) + m[3, r] * ndc[3]).astype(f32) for r in range(4)]
        P = np.stack([p4[0] / p4[3], p4[1] / p4[3], p4[2] / p4[3]], -1).astype(f32)
    N = g["normals"][..., :3].astype(f32)
    mr = g["motion"][..., 2:].astype(f32)
    rsa = np.ones((H, W, 2), f32)
    if shadow_mode == 0 or ao_mode == 0:
        rsa = rt[..., :2].astype(f32)
    cam = np.asarray(pfd["camera_view_inverse"], f32).reshape(4, 4)[3, :3]
    with np.errstate(all="ignore"):
        V = _normalize(cam[None, None, :] - P)
        L = -np.asarray(pfd["directional_light"]["direction"], f32)[:3]
        Hv = _normalize(L[None, None, :] + V)
        shadow = rsa[..., 0] if shadow_mode == 0 else np.ones((H, W), f32)
        ao = rsa[..., 1] if ao_mode == 0 else (ssao[..., 0].astype(f32) if ao_mode == 1 else np.ones((H, W), f32))
        metallic = np.clip(mr[..., 0], 0, 1).astype(f32)
        rough = np.clip(mr[..., 1], f32(0.04), 1).astype(f32)
        li = np.asarray(pfd["directional_light"]["intensity"], f32)[:3]
        lc = np.asarray(pfd["directional_light"]["color"], f32)[:3]
        f0 = (f32(0.04) * (f32(1) - metallic[..., None]) + albedo * metallic[..., None]).astype(f32)
        hv = np.maximum(_dot(Hv, V), 0).astype(f32)
        o = (f32(1) - hv)[..., None]
        F = (f0 + (f32(1) - f0) * o * o * o * o * o).astype(f32)
        ndl = np.maximum(_dot(N, np.broadcast_to(L, N.shape)), 0).astype(f32)
        diffuse_brdf = ((f32(1) - F) * (f32(1) - metallic)[..., None] * albedo / f32(np.pi)).astype(f32)
        a2 = rough * rough
        nh = np.maximum(_dot(N, Hv), 0).astype(f32)
        ff = nh * nh * (a2 - f32(1)) + f32(1)
        D = a2 / (f32(np.pi) * ff * ff)
        k = ((rough + f32(1)) * (rough + f32(1))) * f32(0.125)
        nv = np.maximum(_dot(N, V), 0).astype(f32)
        G = (nv / (nv * (f32(1) - k) + k)) * (ndl / (ndl * (f32(1) - k) + k))
        denom = np.maximum(f32(4) * nv * ndl, f32(1e-6))
        spec_brdf = ((D * G)[..., None] * F / denom[..., None]).astype(f32)
        ambient = ao[..., None] * albedo * f32(0.31830988618379067)
        diffuse = diffuse_brdf * ndl[..., None] * li * lc * shadow[..., None]
        specular = spec_brdf * ndl[..., None] * li * lc * shadow[..., None]
        if reflection_mode == 0:
            r = refl[..., :3].astype(f32) * shadow[..., None]
            mixed = specular * (f32(1) - rough)[..., None] + r * rough[..., None]
            specular = np.where((metallic == 1)[..., None], r, mixed)
        lighting = (ambient + diffuse + specular).astype(f32)
    return np.nan_to_num(lighting, nan=0.0, posinf=np.inf, neginf=-np.inf)


@pytest.mark.parametrize("modes", [(0, 0, 2), (0, 0, 0), (2, 1, 0), (2, 2, 2)])
def test_oracle_composition_matches_numpy(modes):
    W, H = 96, 64
    sc, osc, frames = Hh.scene_and_gbuffer(W, H)
    pfd, g = frames[1]
    rng = np.random.default_rng(11)
    rt = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=7)
    ssao = rng.uniform(0, 1, (H, W, 4)).astype(np.float16)
    want = numpy_composition(pfd, g, rt["shadow_ao"], *modes, ssao=ssao, refl=rt["reflections"])
    got = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], rt["shadow_ao"], *modes, ssao_img=ssao,
                        refl=rt["reflections"]).astype(np.float32)
    lit = g["depth"] > 0
    assert lit.mean() > 0.5
    assert np.all(got[..., 3] == 1.0)
    assert np.all(got[~lit][:, :3] == 0.0)                      # sky: NaN radiance stores as 0
    d = np.abs(got[..., :3] - want)[lit]
    tol = 2e-3 * np.maximum(1.0, np.abs(want[lit]))             # fp16 store (rel 2^-11) + evaluation-order differences
    assert np.all(d <= tol), float(d.max())


def test_oracle_composition_srgb_store():
    W, H = 64, 40
    sc, osc, frames = Hh.scene_and_gbuffer(W, H)
    pfd, g = frames[0]
    rt = osc.raygen(pfd, g["depth"], g["normals"], ao_spp=2, flags=3)
    lin = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], rt["shadow_ao"]).astype(np.float32)
    srgb = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], rt["shadow_ao"], out_format=T.VK_FORMAT_B8G8R8A8_SRGB)
    unorm = O.composition(pfd, g["albedo"], g["normals"], g["motion"], g["depth"], rt["shadow_ao"], out_format=T.VK_FORMAT_B8G8R8A8_UNORM)
    assert np.all(srgb[..., 3] == 255) and np.all(unorm[..., 3] == 255)
    c = np.clip(lin[..., :3], 0, 1)
    enc = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055)
    # BGRA byte order; one code of slack for the fp16 rounding of the linear image used as the reference here
    assert np.max(np.abs(srgb[..., [2, 1, 0]].astype(np.int32) - np.rint(enc * 255).astype(np.int32))) <= 1
    assert np.max(np.abs(unorm[..., [2, 1, 0]].astype(np.int32) - np.rint(c * 255).astype(np.int32))) <= 1
