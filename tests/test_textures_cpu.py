"""Pins the oracle's texture(textures[i], uv) against an independent numpy restatement of the Vulkan texel-filtering rules
(unnormalised coordinates, LINEAR / NEAREST taps, the four address modes, sRGB decode before filtering), and checks the
texture-dependent branches of reflection_hit.rchit / gbuf.frag on a small scene. CPU only."""
import numpy as np
import pytest

import oracle_lib as O
from vulkanhybridrenderer_b200 import camera, scenes
from vulkanhybridrenderer_b200 import types as T

f32 = np.float32


def np_wrap(i, n, mode):
    if mode == 0:
        return i % n
    if mode == 1:
        m = (i % (2 * n)) - n
        m = m if m >= 0 else -(1 + m)
        return (n - 1) - m
    if mode == 2:
        return min(max(i, 0), n - 1)
    return i if 0 <= i < n else -1


def np_texel(tex, x, y):
    if x < 0 or y < 0:
        return np.array([0, 0, 0, 1], f32)
    c = tex.rgba[y, x].astype(np.float64) / 255.0
    out = (tex.rgba[y, x].astype(f32) / f32(255)).astype(f32)
    if tex.format == T.VK_FORMAT_R8G8B8A8_SRGB:
        lin = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
        out[:3] = lin[:3].astype(f32)
    return out


def np_sample(tex, u, v):
    H, W = tex.rgba.shape[:2]
    mag, _, wu, wv = tex.sampler if tex.sampler is not None else (1, 1, 0, 0)
    u, v = f32(u), f32(v)
    if mag == 0:
        i, j = int(np.floor(f32(u * f32(W)))), int(np.floor(f32(v * f32(H))))
        return np_texel(tex, np_wrap(i, W, wu), np_wrap(j, H, wv))
    uu, vv = f32(f32(u * f32(W)) - f32(0.5)), f32(f32(v * f32(H)) - f32(0.5))
    uu, vv = f32(np.floor(f32(f32(uu * f32(256.0)) + f32(0.5))) * f32(0.00390625)), f32(np.floor(f32(f32(vv * f32(256.0)) + f32(0.5))) * f32(0.00390625))      # 8 fractional bits of sub-texel precision
    fu, fv = np.floor(uu), np.floor(vv)
    a, b = f32(uu - fu), f32(vv - fv)
    i, j = int(fu), int(fv)
    x0, x1, y0, y1 = np_wrap(i, W, wu), np_wrap(i + 1, W, wu), np_wrap(j, H, wv), np_wrap(j + 1, H, wv)
    one = f32(1)
    t00, t10, t01, t11 = np_texel(tex, x0, y0), np_texel(tex, x1, y0), np_texel(tex, x0, y1), np_texel(tex, x1, y1)
    return (((f32(f32(one - a) * f32(one - b)) * t00).astype(f32) + (f32(a * f32(one - b)) * t10).astype(f32)).astype(f32)
            + (f32(f32(one - a) * b) * t01).astype(f32)).astype(f32) + (f32(a * b) * t11).astype(f32)


@pytest.fixture(scope="module")
def textured():
    W, H = 96, 64
    sc = scenes.add_procedural_textures(scenes.sponza_like(6000, seed=5, width=W, height=H, n_clutter=20))
    return W, H, sc, O.OracleScene(sc)


def test_texture_sampling_vs_numpy(textured):
    W, H, sc, osc = textured
    rng = np.random.default_rng(1)
    modes_seen = set()
    for idx, tex in enumerate(sc.textures):
        modes_seen.add(tuple(tex.sampler))
        uv = rng.uniform(-2.5, 3.5, (200, 2)).astype(f32)
        uv[:8] = [[0, 0], [1, 1], [0.5, 0.5], [-1, 2], [1e-7, 1 - 1e-7], [0.9999999, 0], [2.0, -2.0], [0.25, 0.75]]
        for u, v in uv:
            got, want = osc.sample_texture(idx, u, v), np_sample(tex, u, v)
            np.testing.assert_allclose(got, want, rtol=0, atol=1.5e-7, err_msg=f"texture {idx} uv=({u},{v})")
    assert len({m[2] for m in modes_seen} | {m[3] for m in modes_seen}) == 4, "all four address modes exercised"
    assert {m[0] for m in modes_seen} == {0, 1}


def test_srgb_decode_known_answers():
    sc = scenes.tiny_scene()
    osc = O.OracleScene(sc)
    px = np.zeros((1, 4, 4), np.uint8)
    px[0, :, 0] = [0, 10, 128, 255]
    px[..., 3] = [0, 64, 128, 255]
    i = osc.add_texture(px, T.VK_FORMAT_R8G8B8A8_SRGB, (0, 0, 2, 2))
    j = osc.add_texture(px, T.VK_FORMAT_R8G8B8A8_UNORM, (0, 0, 2, 2))
    want = [0.0, 10 / 255 / 12.92, ((128 / 255 + 0.055) / 1.055) ** 2.4, 1.0]
    for k in range(4):
        s = osc.sample_texture(i, (k + 0.5) / 4, 0.5)
        assert abs(s[0] - want[k]) < 1e-7 and abs(s[3] - px[0, k, 3] / 255) < 1e-7      # alpha stays linear
        assert abs(osc.sample_texture(j, (k + 0.5) / 4, 0.5)[0] - px[0, k, 0] / 255) < 1e-7


def test_gbuffer_and_reflections_use_textures(textured):
    W, H, sc, osc = textured
    plain = scenes.sponza_like(6000, seed=5, width=W, height=H, n_clutter=20)      # same geometry, constant materials
    oplain = O.OracleScene(plain)
    pfd = camera.FrameSequencer(W, H, sc.light).next(sc.camera)
    g, gp = osc.gbuffer(pfd, W, H, want_ids=True), oplain.gbuffer(pfd, W, H, want_ids=True)
    ids = g["normals"][..., 3].astype(np.int32)
    mat = sc.primitives["material"]
    lit = g["depth"] > 0
    tex_px = lit & (mat["base_color_texture"][np.maximum(ids, 0)] >= 0)
    assert tex_px.mean() > 0.2
    # untextured, unmasked primitives in front of nothing masked: identical texels to the plain scene
    same_hit = np.all(g["ids"] == gp["ids"], axis=-1)
    untex = lit & same_hit & (mat["base_color_texture"][np.maximum(ids, 0)] < 0) & (mat["normal_map"][np.maximum(ids, 0)] < 0)
    assert np.array_equal(g["albedo"][untex], gp["albedo"][untex]) and np.array_equal(g["normals"][untex].view(np.uint16), gp["normals"][untex].view(np.uint16))
    # textured pixels vary inside one primitive (a constant material cannot)
    big = np.bincount(ids[tex_px]).argmax()
    assert len(np.unique(g["albedo"][tex_px & (ids == big)].reshape(-1, 4), axis=0)) > 4
    # alpha cut-outs: some primary rays pass through masked primitives -> different hit than the opaque scene, and no
    # surviving texel of a masked primitive is below its cutoff
    assert (~same_hit & lit).sum() > 0
    masked = lit & (mat["alpha_mask"][np.maximum(ids, 0)] == 1)
    assert masked.any() and (g["albedo"][masked][:, 3].astype(np.float32) / 255.0 >= 0.5 - 1 / 255).all()
    # normal maps perturb the stored normal but keep it unit length
    nm_px = lit & same_hit & (mat["normal_map"][np.maximum(ids, 0)] >= 0)
    n, npl = g["normals"][nm_px][:, :3].astype(np.float32), gp["normals"][nm_px][:, :3].astype(np.float32)
    assert np.abs(n - npl).max() > 0.05 and np.abs(np.linalg.norm(n, axis=-1) - 1).max() < 2e-3
    # metallic-roughness textures scale the factors (factors are 1 where a texture is assigned)
    mr_px = lit & (mat["metallic_roughness_texture"][np.maximum(ids, 0)] >= 0)
    assert len(np.unique(g["motion"][mr_px][:, 2])) > 4
    # reflections: textured hits change the radiance
    r, rp = osc.raygen(pfd, g["depth"], g["normals"], flags=4), oplain.raygen(pfd, g["depth"], g["normals"], flags=4)
    assert np.abs(r["reflections"].astype(np.float32) - rp["reflections"].astype(np.float32)).max() > 0.05


# ---------------------------------------------------------------------------------------------------------------
# The alpha-tested ray-traced pipeline (raygen_test_alpha.rgen, closesthit_test_alpha.rchit, shadow_anyhit.rahit) transcribed in
# float64 with np_sample() above as texture(): candidate hits in distance order, the any-hit shader drops the ones whose base-colour
# alpha is below the cutoff, for the primary ray AND the shadow ray (gl_RayFlagsNoOpaqueEXT on both)
# ---------------------------------------------------------------------------------------------------------------
def _hits_in_order(tris64, o, d, tmin, tmax):
    e1 = tris64[:, 1] - tris64[:, 0]; e2 = tris64[:, 2] - tris64[:, 0]
    pv = np.cross(d[None, :], e2)
    det = np.einsum("ij,ij->i", e1, pv)
    ok = np.abs(det) > 0
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    tv = o[None, :] - tris64[:, 0]
    u = np.einsum("ij,ij->i", tv, pv) * inv
    qv = np.cross(tv, e1)
    v = np.einsum("j,ij->i", d, qv) * inv
    t = np.einsum("ij,ij->i", e2, qv) * inv
    inside = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > tmin) & (t < tmax)
    idx = np.nonzero(inside)[0]
    idx = idx[np.argsort(t[idx])]
    return [(int(k), float(u[k]), float(v[k]), float(t[k]), float(min(u[k], v[k], 1 - u[k] - v[k]))) for k in idx]


def test_alpha_tested_raytraced_pipeline_vs_transcription(textured):
    import helpers as Hh
    W, H, sc, osc = textured
    tris64 = Hh.world_triangles(sc).astype(np.float64)
    counts = np.array([int(p["index_count"]) // 3 for p in sc.primitives])
    first = np.concatenate([[0], np.cumsum(counts)])
    pfd = camera.FrameSequencer(W, H, sc.light).next(sc.camera)
    got = osc.raytraced(pfd, W, H, alpha_test=True)
    col = lambda name: np.asarray(pfd[name], np.float64).reshape(4, 4).T
    view_inv, proj_inv = col("camera_view_inverse"), col("camera_proj_inverse")
    light_dir = -np.asarray(pfd["directional_light"]["direction"], np.float64)[:3]
    lc = np.asarray(pfd["directional_light"]["color"], np.float64)[:3]
    to_bgra8 = lambda rgba: np.round(np.clip(np.asarray(rgba, np.float64), 0, 1)[[2, 1, 0, 3]] * 255.0)

    def surface(k, b1, b2):
        gi = int(np.searchsorted(first, k, side="right") - 1)
        p = sc.primitives[gi]
        base = int(p["index_offset"]) + 3 * (k - int(first[gi]))
        vi = [int(p["vertex_offset"]) + int(sc.indices[base + j]) for j in range(3)]
        bary = np.array([1.0 - b1 - b2, b1, b2])
        mix = lambda field: sum(sc.vertices[field][vi[j]].astype(np.float64) * bary[j] for j in range(3))
        return p, mix("uv0"), mix("normal"), (np.asarray(p["transform"], np.float64).reshape(4, 4).T @ np.append(mix("pos"), 1.0))[:3]

    def first_accepted(o, d):
        """Closest hit the any-hit shader keeps; ambiguous = a candidate grazes an edge or sits on the cutoff."""
        ambiguous = False
        cands = _hits_in_order(tris64, o, d, 0.1, 10000.0)
        for n, (k, b1, b2, t, edge) in enumerate(cands):
            p, uv, _, _ = surface(k, b1, b2)
            m = p["material"]
            ambiguous |= edge < 1e-3
            if n + 1 < len(cands) and cands[n + 1][3] - t < 1e-4 * max(1.0, t):
                ambiguous = True                           # coincident surfaces: either one may win
            if int(m["alpha_mask"]) == 1 and int(m["base_color_texture"]) >= 0:
                a = float(np_sample(sc.textures[int(m["base_color_texture"])], uv[0], uv[1])[3])
                ambiguous |= abs(a - float(m["alpha_cutoff"])) < 2e-2
                if a < float(m["alpha_cutoff"]):
                    continue                               # ignoreIntersectionEXT
            return (k, b1, b2), ambiguous
        return None, ambiguous

    n_checked = n_through = n_shadow_through = 0
    for y in range(0, H, 3):
        for x in range(0, W, 3):
            ndc = np.array([(x + 0.5) / W, (y + 0.5) / H]) * 2.0 - 1.0
            origin = (view_inv @ np.array([0, 0, 0, 1.0]))[:3]
            target = proj_inv @ np.array([ndc[0], ndc[1], 1.0, 1.0])
            direction = (view_inv @ np.append(target[:3] / np.linalg.norm(target[:3]), 0.0))[:3]
            cands = _hits_in_order(tris64, origin, direction, 0.1, 10000.0)
            hit, amb = first_accepted(origin, direction)
            if hit is None or amb:
                continue
            p, uv, normal, position = surface(*hit)
            m = p["material"]
            if int(m["base_color_texture"]) < 0 or int(m["normal_map"]) >= 0:
                continue                                   # closesthit_test_alpha.rchit samples textures[base_color_texture] unconditionally
            n_through += hit[0] != cands[0][0]
            albedo = np_sample(sc.textures[int(m["base_color_texture"])], uv[0], uv[1])[:3].astype(np.float64)
            occluder, amb_s = first_accepted(position, light_dir)
            if amb_s:
                continue
            n_shadow_through += occluder is None and len(_hits_in_order(tris64, position, light_dir, 0.1, 10000.0)) > 0
            rgb = 0.2 * albedo
            if occluder is None:
                rgb = rgb + max(normal @ light_dir, 0.0) * albedo * lc      # no light_intensity in the alpha-tested shader
            want = to_bgra8(np.append(rgb, 1.0))
            assert np.all(np.abs(got[y, x].astype(np.float64) - want) <= 1), (x, y, got[y, x], want)
            n_checked += 1
    assert n_checked > 60, n_checked
    assert n_through + n_shadow_through > 0, "no ray passed through a cut-out: the any-hit path was not exercised"
