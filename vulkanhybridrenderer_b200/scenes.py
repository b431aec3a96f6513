"""Deterministic procedural scenes for the hot path (no assets ship with the reference: data/models/ is
git-ignored, /root/reference/.gitignore:3; Sponza/Bistro are unavailable offline).

`sponza_like(target_triangles, seed)` builds a colonnaded hall — tiled floor, outer walls with relief, covered
side aisles, an open nave, two rows of fluted columns joined by arches, hanging drapes and a few hundred small
objects — as the flat arrays `ResourceManager::UpdateGeometry` consumes
(/root/reference/src/rendering_backend/resource_manager.cpp:291-360): `Vertex[]`, `uint32 indices[]` (relative to
each primitive's vertex_offset) and `Primitive[]` with a per-primitive world transform and constant material.
Vertices are shared (indexed grids, ~2 triangles per vertex) so a 3M-triangle scene fits the reference's 256 MB
vertex buffer (resource_manager.cpp:13). The flat primitive index is the G-buffer object id.
"""
import numpy as np

from . import types as T
from .camera import Camera, directional_light


class Texture:
    """One entry of textures[]: what the loader hands ResourceManager::UploadTextureFromData (scene_loader.cpp:277-309)."""

    def __init__(self, rgba, fmt, sampler=None):
        self.rgba = np.ascontiguousarray(rgba, np.uint8)       # [H, W, 4]
        self.format = fmt                                      # VK_FORMAT_R8G8B8A8_SRGB (base colour) / _UNORM
        self.sampler = sampler                                 # (mag, min, wrap_u, wrap_v) as VkFilter / VkSamplerAddressMode, None = default


class Scene:
    def __init__(self, vertices, indices, primitives, camera, light, name, textures=None):
        self.vertices = vertices
        self.indices = indices
        self.primitives = primitives
        self.camera = camera
        self.light = light
        self.name = name
        self.textures = textures or []          # index = the slot the materials name

    @property
    def num_triangles(self):
        return int(self.primitives["index_count"].sum() // 3)


def _grid(nu, nv, fn):
    """Tessellate fn(u, v) -> (pos[...,3]) over [0,1]^2 into an indexed grid; normals by central differences."""
    nu, nv = max(int(nu), 1), max(int(nv), 1)
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    p = fn(uu, vv)
    e = 1e-4
    du = fn(np.clip(uu + e, 0, 1), vv) - fn(np.clip(uu - e, 0, 1), vv)
    dv = fn(uu, np.clip(vv + e, 0, 1)) - fn(uu, np.clip(vv - e, 0, 1))
    n = np.cross(du, dv)
    ln = np.linalg.norm(n, axis=-1, keepdims=True)
    n = np.where(ln > 1e-20, n / np.maximum(ln, 1e-20), np.array([0.0, 1.0, 0.0]))
    vid = np.arange((nu + 1) * (nv + 1), dtype=np.uint32).reshape(nu + 1, nv + 1)
    a, b, c, d = vid[:-1, :-1], vid[1:, :-1], vid[1:, 1:], vid[:-1, 1:]
    tris = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    verts = np.zeros((nu + 1) * (nv + 1), T.Vertex)
    verts["pos"] = p.reshape(-1, 3)
    verts["normal"] = n.reshape(-1, 3)
    verts["tangent"] = (1.0, 0.0, 0.0, 1.0)
    verts["uv0"] = np.stack([uu, vv], -1).reshape(-1, 2)
    verts["uv1"] = verts["uv0"]
    return verts, tris.reshape(-1).astype(np.uint32)


def _orient(verts, inds, ref, toward=True):
    """Flip winding + normals so they face `ref` (toward=True) or away from it (closed objects: ref = centroid)."""
    d = np.asarray(ref, np.float64)[None, :] - verts["pos"].astype(np.float64)
    s = float(np.mean(np.sum(d * verts["normal"], axis=-1)))
    if (s < 0) == toward:
        verts["normal"] *= -1.0
        inds = inds.reshape(-1, 3)[:, ::-1].reshape(-1).copy()
    return verts, inds


def _trs(translation=(0, 0, 0), yaw=0.0, scale=(1, 1, 1), pitch=0.0):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    m = np.eye(4)
    m[:3, :3] = ry @ rx @ np.diag(scale)
    m[:3, 3] = translation
    return m


class _Builder:
    def __init__(self, rng):
        self.rng = rng
        self.verts, self.inds, self.prims = [], [], []
        self.v_off = 0
        self.i_off = 0

    def add(self, verts, inds, transform, color=None, metallic=None, roughness=None):
        rng = self.rng
        p = np.zeros((), T.Primitive)
        p["transform"] = np.asarray(transform, np.float32).T      # glm column-major
        m = p["material"]
        col = rng.uniform(0.25, 0.95, 3) if color is None else color
        m["base_color"] = (col[0], col[1], col[2], 1.0)
        m["base_color_texture"] = -1
        m["metallic_roughness_texture"] = -1
        m["normal_map"] = -1
        m["metallic_factor"] = rng.uniform(0.0, 1.0) if metallic is None else metallic
        m["roughness_factor"] = rng.uniform(0.1, 0.9) if roughness is None else roughness
        m["alpha_mask"] = 0
        m["alpha_cutoff"] = 0.5
        p["vertex_offset"] = self.v_off
        p["index_offset"] = self.i_off
        p["index_count"] = len(inds)
        self.verts.append(verts)
        self.inds.append(inds)
        self.prims.append(p)
        self.v_off += len(verts)
        self.i_off += len(inds)

    def finish(self):
        return (np.concatenate(self.verts), np.concatenate(self.inds), np.array(self.prims, T.Primitive))


def _dims(tris, aspect):
    """(nu, nv) with 2*nu*nv ~= tris and nu/nv ~= aspect."""
    cells = max(tris / 2.0, 1.0)
    nv = max(int(round(np.sqrt(cells / aspect))), 1)
    nu = max(int(round(cells / nv)), 1)
    return nu, nv


def sponza_like(target_triangles=260_000, seed=3, width=1920, height=1080, n_clutter=200):
    rng = np.random.default_rng(seed)
    b = _Builder(rng)
    L, HW, AW, HC, HT = 20.0, 4.0, 8.0, 5.0, 10.0   # half length, nave half width, aisle outer z, aisle ceiling y, wall top y
    n_cols = 13
    col_x = np.linspace(-L + 2.0, L - 2.0, n_cols)

    # triangle budget (fractions of the target)
    budget = {
        "floor": 0.16, "walls": 0.16, "ceil": 0.08, "gallery": 0.08,
        "columns": 0.18, "arches": 0.08, "drapes": 0.12, "clutter": 0.14,
    }
    tt = float(target_triangles)

    # floor: 8 x 2 tiles with shallow relief (stone slabs)
    tiles_x, tiles_z = 8, 2
    per = tt * budget["floor"] / (tiles_x * tiles_z)
    for ix in range(tiles_x):
        for iz in range(tiles_z):
            x0, x1 = -L + ix * (2 * L / tiles_x), -L + (ix + 1) * (2 * L / tiles_x)
            z0, z1 = -AW + iz * AW, -AW + (iz + 1) * AW
            nu, nv = _dims(per, (x1 - x0) / (z1 - z0))

            def f(u, v, x0=x0, x1=x1, z0=z0, z1=z1):
                x = x0 + u * (x1 - x0)
                z = z0 + v * (z1 - z0)
                y = 0.015 * np.sin(6.0 * x) * np.sin(6.0 * z) + 0.004 * np.sin(41.0 * x + 13.0 * z)
                return np.stack([x, y, z], -1)
            vs, ii = _orient(*_grid(nu, nv, f), ref=(0.0, 50.0, 0.0))
            b.add(vs, ii, _trs(), roughness=0.6, metallic=0.0)

    # outer walls (z = +-AW, 0..HT) and end walls (x = +-L), with brick-like relief; split in segments
    segs = 4
    per = tt * budget["walls"] / (2 * segs + 2)
    for side in (-1.0, 1.0):
        for s in range(segs):
            x0, x1 = -L + s * (2 * L / segs), -L + (s + 1) * (2 * L / segs)
            nu, nv = _dims(per, (x1 - x0) / HT)

            def f(u, v, x0=x0, x1=x1, side=side):
                x = x0 + u * (x1 - x0)
                y = v * HT
                z = side * (AW - 0.03 * np.sin(9.0 * x) * np.sin(14.0 * y))
                return np.stack([x, y, z], -1)
            vs, ii = _orient(*_grid(nu, nv, f), ref=(0.0, 3.0, 0.0))
            b.add(vs, ii, _trs(), metallic=0.0)
    for side in (-1.0, 1.0):
        nu, nv = _dims(per, (2 * AW) / HT)

        def f(u, v, side=side):
            z = -AW + u * 2 * AW
            y = v * HT
            x = side * (L - 0.03 * np.sin(9.0 * z) * np.sin(14.0 * y))
            return np.stack([x, y, z], -1)
        vs, ii = _orient(*_grid(nu, nv, f), ref=(0.0, 3.0, 0.0))
        b.add(vs, ii, _trs(), metallic=0.0)

    # aisle ceilings: shallow barrel vaults over z in [-AW,-HW] and [HW,AW] at y ~ HC
    per = tt * budget["ceil"] / (2 * segs)
    for side in (-1.0, 1.0):
        for s in range(segs):
            x0, x1 = -L + s * (2 * L / segs), -L + (s + 1) * (2 * L / segs)
            nu, nv = _dims(per, (x1 - x0) / (AW - HW))

            def f(u, v, x0=x0, x1=x1, side=side):
                x = x0 + u * (x1 - x0)
                zc = HW + v * (AW - HW)
                y = HC + 0.6 * np.sin(np.pi * v) + 0.05 * np.sin(3.0 * x)
                return np.stack([x, y, side * zc], -1)
            vs, ii = _orient(*_grid(nu, nv, f), ref=(0.0, -50.0, side * 6.0))
            b.add(vs, ii, _trs(), metallic=0.0)

    # gallery walls above the colonnades (z = +-HW, HC..HT): the nave stays open to the sky
    per = tt * budget["gallery"] / (2 * segs)
    for side in (-1.0, 1.0):
        for s in range(segs):
            x0, x1 = -L + s * (2 * L / segs), -L + (s + 1) * (2 * L / segs)
            nu, nv = _dims(per, (x1 - x0) / (HT - HC))

            def f(u, v, x0=x0, x1=x1, side=side):
                x = x0 + u * (x1 - x0)
                y = HC + 0.7 + v * (HT - HC - 0.7)
                z = side * (HW + 0.04 * np.sin(7.0 * x) * np.cos(9.0 * y))
                return np.stack([x, y, z], -1)
            vs, ii = _orient(*_grid(nu, nv, f), ref=(0.0, 7.5, 0.0))
            b.add(vs, ii, _trs(), metallic=0.0)

    # fluted columns (object space: unit-radius, unit-height cylinder; transform scales/positions it)
    per = tt * budget["columns"] / (2 * n_cols)
    nu, nv = _dims(per, 2.0)

    def col(u, v):
        ang = 2 * np.pi * u
        r = 1.0 + 0.06 * np.cos(16 * ang) + 0.25 * np.exp(-40.0 * v) + 0.25 * np.exp(-40.0 * (1 - v))
        return np.stack([r * np.cos(ang), v, -r * np.sin(ang)], -1)
    col_v, col_i = _orient(*_grid(nu, nv, col), ref=(0.0, 0.5, 0.0), toward=False)
    for side in (-1.0, 1.0):
        for x in col_x:
            b.add(col_v.copy(), col_i, _trs((x, 0.0, side * HW), yaw=rng.uniform(0, 6.28), scale=(0.35, HC, 0.35)),
                  roughness=0.5, metallic=0.0)

    # arches between neighbouring columns: half tori in the x-y plane
    per = tt * budget["arches"] / (2 * (n_cols - 1))
    nu, nv = _dims(per, 4.0)

    def arch(u, v):
        a = np.pi * u
        bb = 2 * np.pi * v
        R, r = 1.0, 0.18
        return np.stack([(R + r * np.cos(bb)) * np.cos(a), (R + r * np.cos(bb)) * np.sin(a), r * np.sin(bb)], -1)
    arch_v, arch_i = _orient(*_grid(nu, nv, arch), ref=(0.0, 0.6, 0.0), toward=False)
    gap = col_x[1] - col_x[0]
    for side in (-1.0, 1.0):
        for k in range(n_cols - 1):
            xc = 0.5 * (col_x[k] + col_x[k + 1])
            b.add(arch_v.copy(), arch_i, _trs((xc, HC - 0.05, side * HW), scale=(gap / 2, 0.8, 1.0)), metallic=0.0)

    # drapes hanging across the nave
    n_drapes = 8
    per = tt * budget["drapes"] / n_drapes
    for k in range(n_drapes):
        x = -L + 4.0 + k * (2 * L - 8.0) / (n_drapes - 1)
        ph = rng.uniform(0, 6.28)
        nu, nv = _dims(per, 2 * HW / 3.0)

        def f(u, v, ph=ph):
            z = (u - 0.5) * 2.0
            y = -v
            xx = 0.12 * np.sin(14.0 * z + ph) * (0.3 + v) + 0.05 * np.sin(5.0 * y + ph)
            return np.stack([xx, y, z], -1)
        vs, ii = _grid(nu, nv, f)
        b.add(vs, ii, _trs((x, HT - 1.0 - 0.3 * k % 2, 0.0), scale=(1.0, 3.0, HW * 0.45)),
              color=rng.uniform(0.2, 0.9, 3), roughness=0.9, metallic=0.0)

    # clutter: vases (surfaces of revolution) and blobs with random similarity/non-uniform transforms
    per = tt * budget["clutter"] / n_clutter
    nu, nv = _dims(per, 1.5)

    def vase(u, v):
        ang = 2 * np.pi * u
        r = 0.35 + 0.25 * np.sin(np.pi * v) ** 2 + 0.1 * np.sin(3 * np.pi * v)
        return np.stack([r * np.cos(ang), v, -r * np.sin(ang)], -1)

    def blob(u, v):
        th = np.pi * (v * 0.998 + 0.001)
        ph = 2 * np.pi * u
        r = 0.5 * (1.0 + 0.15 * np.sin(5 * ph) * np.sin(4 * th))
        return np.stack([r * np.sin(th) * np.cos(ph), 0.5 - r * np.cos(th) * 0.98, -r * np.sin(th) * np.sin(ph)], -1)
    vase_v, vase_i = _orient(*_grid(nu, nv, vase), ref=(0.0, 0.5, 0.0), toward=False)
    blob_v, blob_i = _orient(*_grid(nu, nv, blob), ref=(0.0, 0.5, 0.0), toward=False)
    for k in range(n_clutter):
        x = rng.uniform(-L + 1.0, L - 1.0)
        z = rng.uniform(-AW + 0.8, AW - 0.8)
        s = rng.uniform(0.3, 0.9)
        sc = (s * rng.uniform(0.7, 1.3), s * rng.uniform(0.8, 1.8), s * rng.uniform(0.7, 1.3))
        vs, ii = (vase_v, vase_i) if k % 2 == 0 else (blob_v, blob_i)
        b.add(vs.copy(), ii, _trs((x, 0.02, z), yaw=rng.uniform(0, 6.28), scale=sc))

    vertices, indices, primitives = b.finish()
    cam = Camera(position=(-15.0, 2.2, 0.6), yaw=np.deg2rad(-97.0), pitch=np.deg2rad(7.0),
                 yfov=np.deg2rad(60.0), aspect=width / height, znear=0.1)
    light = directional_light((-0.3, -1.0, 0.2), intensity=30.0)
    return Scene(vertices, indices, primitives, cam, light, f"sponza_like_{target_triangles}")


def tiny_scene(seed=0, width=64, height=48):
    """A few dozen triangles (floor quad grid + two blobs) for smoke tests and edge cases."""
    rng = np.random.default_rng(seed)
    b = _Builder(rng)

    def floor(u, v):
        return np.stack([(u - 0.5) * 8, 0 * u, (v - 0.5) * 8], -1)
    vs, ii = _orient(*_grid(4, 4, floor), ref=(0.0, 50.0, 0.0))
    b.add(vs, ii, _trs())

    def blob(u, v):
        th = np.pi * (v * 0.998 + 0.001)
        ph = 2 * np.pi * u
        return np.stack([0.5 * np.sin(th) * np.cos(ph), 0.5 - 0.5 * np.cos(th), -0.5 * np.sin(th) * np.sin(ph)], -1)
    vs, ii = _orient(*_grid(10, 8, blob), ref=(0.0, 0.5, 0.0), toward=False)
    b.add(vs.copy(), ii, _trs((0.5, 0.0, -1.0), yaw=0.3, scale=(1.0, 1.6, 0.8)))
    b.add(vs.copy(), ii, _trs((-1.2, 0.4, -2.0), yaw=1.1, scale=(0.7, 0.7, 1.4)))
    vertices, indices, primitives = b.finish()
    cam = Camera(position=(0.0, 1.2, 3.0), yaw=0.0, pitch=np.deg2rad(-12.0), yfov=np.deg2rad(60.0),
                 aspect=width / height, znear=0.1)
    light = directional_light((-0.3, -1.0, 0.2), intensity=30.0)
    return Scene(vertices, indices, primitives, cam, light, "tiny")


def add_procedural_textures(scene, seed=11, size=64, fraction=0.6, uv_scale=3.0):
    """Gives a deterministic subset of the primitives textured materials (no assets ship with the reference): a bank of
    procedural base-colour (sRGB, some with alpha cut-outs), metallic-roughness and normal-map textures with a mix of
    samplers, assigned round-robin; uv0 is scaled so REPEAT / MIRRORED_REPEAT / CLAMP addressing all get exercised.
    Returns the scene (modified in place)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    NEAREST, LINEAR = 0, 1
    REPEAT, MIRROR, CLAMP, BORDER = 0, 1, 2, 3
    tex = []

    def rgba(r, g, b, a=None):
        a = np.full_like(r, 255) if a is None else a
        return np.stack([r, g, b, a], -1).astype(np.uint8)
    # base colour (sRGB): checker, stripes with alpha holes, value noise
    chk = ((xx // 8 + yy // 8) % 2).astype(np.uint8)
    tex.append(Texture(rgba(60 + 150 * chk, 90 + 100 * chk, 200 - 120 * chk), T.VK_FORMAT_R8G8B8A8_SRGB, (LINEAR, LINEAR, REPEAT, REPEAT)))
    holes = (((xx % 16) - 8) ** 2 + ((yy % 16) - 8) ** 2 < 20)
    tex.append(Texture(rgba(200 - 3 * (xx % 32), 80 + 2 * yy, 40 + xx, np.where(holes, 30, 255)), T.VK_FORMAT_R8G8B8A8_SRGB, (LINEAR, LINEAR, MIRROR, REPEAT)))
    noise = rng.integers(0, 256, (size, size, 3))
    tex.append(Texture(rgba(noise[..., 0], noise[..., 1], noise[..., 2]), T.VK_FORMAT_R8G8B8A8_SRGB, (NEAREST, NEAREST, REPEAT, CLAMP)))
    tex.append(Texture(rgba(255 - 2 * xx, 128 + (yy % 64), 2 * yy + 40, 255 - 3 * ((xx + yy) % 64)), T.VK_FORMAT_R8G8B8A8_SRGB, (LINEAR, NEAREST, CLAMP, BORDER)))
    n_color = len(tex)
    # metallic-roughness (UNORM; .g = roughness factor... the shader reads metallic from .g and roughness from .b)
    tex.append(Texture(rgba(0 * xx, 40 + 3 * xx, 255 - 3 * yy), T.VK_FORMAT_R8G8B8A8_UNORM, (LINEAR, LINEAR, REPEAT, REPEAT)))
    tex.append(Texture(rgba(0 * xx, 255 * chk, 90 + 100 * chk), T.VK_FORMAT_R8G8B8A8_UNORM, (LINEAR, LINEAR, MIRROR, MIRROR)))
    n_mr = len(tex) - n_color
    # normal maps (UNORM): sinusoidal bumps, tangent-space
    ph = 2 * np.pi * xx / 16.0
    nx, ny = 0.35 * np.cos(ph), 0.35 * np.sin(2 * np.pi * yy / 16.0)
    nz = np.sqrt(np.maximum(1.0 - nx * nx - ny * ny, 0.0))
    enc = lambda c: np.clip(np.rint((c * 0.5 + 0.5) * 255.0), 0, 255)
    tex.append(Texture(rgba(enc(nx), enc(ny), enc(nz)), T.VK_FORMAT_R8G8B8A8_UNORM, (LINEAR, LINEAR, REPEAT, REPEAT)))
    n_nm = 1
    scene.textures = tex
    prims = scene.primitives
    k = 0
    for g in range(len(prims)):
        if rng.uniform() > fraction:
            continue
        m = prims[g]["material"]
        m["base_color_texture"] = k % n_color
        if k % 2 == 0:
            m["metallic_roughness_texture"] = n_color + (k // 2) % n_mr
            m["metallic_factor"] = 1.0
            m["roughness_factor"] = 1.0
        if k % 3 == 0:
            m["normal_map"] = n_color + n_mr + (k // 3) % n_nm
        if (k % n_color) in (1, 3) and k % 4 != 3:
            m["alpha_mask"] = 1
            m["alpha_cutoff"] = 0.5
        k += 1
        v0, n = int(prims[g]["vertex_offset"]), None
        v1 = int(prims[g + 1]["vertex_offset"]) if g + 1 < len(prims) else len(scene.vertices)
        scene.vertices["uv0"][v0:v1] = (scene.vertices["uv0"][v0:v1] * np.float32(uv_scale) - np.float32(0.6)).astype(np.float32)
    return scene
