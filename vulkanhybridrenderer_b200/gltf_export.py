"""Writes a `scenes.Scene` as glTF 2.0 (.gltf + .bin + .png files, or one .glb), so the scene-loader path
(host/scene_loader.cpp <- /root/reference/src/scene/scene_loader.cpp) can be exercised offline: no assets ship with the
reference (data/models/ is git-ignored). One glTF mesh + node per primitive (node.matrix = the primitive's transform),
a perspective camera node, a KHR_lights_punctual directional light node, PNG textures with samplers.
"""
import base64
import json
import os
import struct
import zlib

import numpy as np

from . import types as T

_FILTER = {0: 9728, 1: 9729}                            # VkFilter -> GL NEAREST / LINEAR
_WRAP = {0: 10497, 1: 33648, 2: 33071, 3: 33069}        # VkSamplerAddressMode -> GL REPEAT / MIRRORED_REPEAT / CLAMP_TO_EDGE / CLAMP_TO_BORDER


def encode_png(rgba, color_type=6, bit_depth=8, filter_type=None):
    """Minimal PNG writer (zlib). rgba: [H, W, 4] uint8. color_type 6 RGBA, 2 RGB, 0 grey, 4 grey+alpha; bit_depth 8 or 16;
    filter_type None = cycle through the five scanline filters (exercises the decoder), else 0..4."""
    a = np.ascontiguousarray(rgba, np.uint8)
    H, W = a.shape[:2]
    if color_type == 6:
        px = a
    elif color_type == 2:
        px = a[..., :3]
    elif color_type == 0:
        px = a[..., :1]
    elif color_type == 4:
        px = a[..., [0, 3]]
    else:
        raise ValueError(color_type)
    if bit_depth == 16:
        px = np.repeat(px[..., None], 2, axis=-1).reshape(H, W, -1)     # v -> v * 257 (high byte = low byte)
    elif bit_depth != 8:
        raise ValueError(bit_depth)
    rows = px.reshape(H, -1).astype(np.int32)
    bpp = rows.shape[1] // W
    raw = bytearray()
    prev = np.zeros(rows.shape[1], np.int32)
    for y in range(H):
        cur = rows[y]
        ft = (y % 5) if filter_type is None else filter_type
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = left
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (left + prev) >> 1
        else:
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
        raw.append(ft)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    comp = zlib.compress(bytes(raw), 6)
    half = len(comp) // 2            # two IDAT chunks: the decoder must concatenate them
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, bit_depth, color_type, 0, 0, 0)) +
            chunk(b"IDAT", comp[:half]) + chunk(b"IDAT", comp[half:]) + chunk(b"IEND", b""))


def _pad4(b, fill=b"\x00"):
    return b + fill * ((-len(b)) % 4)


def export(scene, path, camera_node=True, light_node=True, embed=False):
    """Writes `scene` to `path` (.gltf: side files <stem>.bin and <stem>_texN.png next to it; .glb: everything inside;
    embed=True with .gltf: base64 data URI for the buffer)."""
    path = str(path)
    glb = path.lower().endswith(".glb")
    stem = os.path.splitext(os.path.basename(path))[0]
    d = os.path.dirname(path)
    blob = bytearray()
    views, accessors = [], []

    def add_view(data, stride=None):
        off = len(blob)
        blob.extend(_pad4(data))
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        return len(views) - 1

    def add_accessor(view, comp, count, typ, offset=0, minmax=None):
        a = {"bufferView": view, "componentType": comp, "count": int(count), "type": typ}
        if offset:
            a["byteOffset"] = offset
        if minmax is not None:
            a["min"], a["max"] = minmax
        accessors.append(a)
        return len(accessors) - 1

    doc = {"asset": {"version": "2.0", "generator": "vulkanhybridrenderer_b200.gltf_export"}, "scene": 0, "scenes": [{"nodes": []}],
           "nodes": [], "meshes": [], "materials": [], "buffers": [], "bufferViews": views, "accessors": accessors}
    # textures
    if scene.textures:
        doc["images"], doc["textures"], doc["samplers"] = [], [], []
        for k, t in enumerate(scene.textures):
            png = encode_png(t.rgba)
            if glb:
                doc["images"].append({"bufferView": add_view(png), "mimeType": "image/png", "name": f"tex{k}"})
            else:
                fn = f"{stem}_tex{k}.png"
                with open(os.path.join(d, fn), "wb") as f:
                    f.write(png)
                doc["images"].append({"uri": fn, "name": f"tex{k}"})
            mag, mn, wu, wv = t.sampler if t.sampler is not None else (1, 1, 0, 0)
            doc["samplers"].append({"magFilter": _FILTER[mag], "minFilter": _FILTER[mn], "wrapS": _WRAP[wu], "wrapT": _WRAP[wv]})
            doc["textures"].append({"source": k, "sampler": k})
    # one mesh + node per primitive; the interleaved Vertex array is exported as is (byteStride 56)
    verts = np.ascontiguousarray(scene.vertices)
    for g, p in enumerate(scene.primitives):
        v0 = int(p["vertex_offset"])
        v1 = int(scene.primitives[g + 1]["vertex_offset"]) if g + 1 < len(scene.primitives) else len(verts)
        i0, n = int(p["index_offset"]), int(p["index_count"])
        vb = verts[v0:v1]
        view = add_view(vb.tobytes(), stride=T.Vertex.itemsize)
        pos = vb["pos"]
        attrs = {"POSITION": add_accessor(view, 5126, len(vb), "VEC3", 0, (pos.min(0).tolist(), pos.max(0).tolist())),
                 "NORMAL": add_accessor(view, 5126, len(vb), "VEC3", 12), "TANGENT": add_accessor(view, 5126, len(vb), "VEC4", 24),
                 "TEXCOORD_0": add_accessor(view, 5126, len(vb), "VEC2", 40), "TEXCOORD_1": add_accessor(view, 5126, len(vb), "VEC2", 48)}
        idx = np.ascontiguousarray(scene.indices[i0:i0 + n])
        small = len(vb) <= 65535 and g % 2 == 1                      # alternate 16- and 32-bit indices
        ia = add_accessor(add_view(idx.astype(np.uint16).tobytes() if small else idx.astype(np.uint32).tobytes()), 5123 if small else 5125, n, "SCALAR")
        m = p["material"]
        pbr = {"metallicFactor": float(m["metallic_factor"]), "roughnessFactor": float(m["roughness_factor"])}
        if m["base_color_texture"] >= 0:
            pbr["baseColorTexture"] = {"index": int(m["base_color_texture"])}
        else:
            pbr["baseColorFactor"] = [float(x) for x in m["base_color"]]
        if m["metallic_roughness_texture"] >= 0:
            pbr["metallicRoughnessTexture"] = {"index": int(m["metallic_roughness_texture"])}
        mat = {"pbrMetallicRoughness": pbr}
        if m["normal_map"] >= 0:
            mat["normalTexture"] = {"index": int(m["normal_map"])}
        if m["alpha_mask"] == 1:
            mat["alphaMode"], mat["alphaCutoff"] = "MASK", float(m["alpha_cutoff"])
        doc["materials"].append(mat)
        doc["meshes"].append({"primitives": [{"attributes": attrs, "indices": ia, "material": g, "mode": 4}]})
        doc["nodes"].append({"mesh": g, "matrix": [float(x) for x in np.asarray(p["transform"], np.float32).reshape(-1)]})
        doc["scenes"][0]["nodes"].append(len(doc["nodes"]) - 1)
    if camera_node and scene.camera is not None:
        cam = scene.camera
        doc["cameras"] = [{"type": "perspective", "perspective": {"yfov": float(cam.yfov), "aspectRatio": float(cam.aspect), "znear": float(cam.znear)}}]
        doc["nodes"].append({"camera": 0, "matrix": [float(x) for x in np.asarray(cam.transform, np.float32).T.reshape(-1)]})
        doc["scenes"][0]["nodes"].append(len(doc["nodes"]) - 1)
    if light_node and scene.light is not None:
        dvec = np.asarray(scene.light["direction"], np.float64)[:3]
        dvec = dvec / np.linalg.norm(dvec)
        # rotation taking (0,0,-1) to the light direction, as a quaternion on a child node under a scaled, translated parent
        z = np.array([0.0, 0.0, -1.0])
        axis = np.cross(z, dvec)
        s, c = np.linalg.norm(axis), float(np.dot(z, dvec))
        if s < 1e-12:
            q = [0.0, 0.0, 0.0, 1.0] if c > 0 else [1.0, 0.0, 0.0, 0.0]
        else:
            half = 0.5 * np.arctan2(s, c)
            ax = axis / s
            q = [float(ax[0] * np.sin(half)), float(ax[1] * np.sin(half)), float(ax[2] * np.sin(half)), float(np.cos(half))]
        doc["extensionsUsed"] = ["KHR_lights_punctual"]
        col = [float(x) for x in scene.light["color"][:3]]
        doc["extensions"] = {"KHR_lights_punctual": {"lights": [{"type": "directional", "color": col, "intensity": 3.0}]}}
        child = len(doc["nodes"])
        doc["nodes"].append({"rotation": q, "extensions": {"KHR_lights_punctual": {"light": 0}}})
        doc["nodes"].append({"translation": [1.0, 20.0, -3.0], "scale": [2.0, 2.0, 2.0], "children": [child]})
        doc["scenes"][0]["nodes"].append(len(doc["nodes"]) - 1)
    data = bytes(blob)
    if glb:
        doc["buffers"].append({"byteLength": len(data)})
        js = _pad4(json.dumps(doc, separators=(",", ":")).encode(), b" ")
        bn = _pad4(data)
        with open(path, "wb") as f:
            f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(bn)))
            f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
            f.write(struct.pack("<II", len(bn), 0x004E4942) + bn)
    else:
        if embed:
            doc["buffers"].append({"byteLength": len(data), "uri": "data:application/octet-stream;base64," + base64.b64encode(data).decode()})
        else:
            with open(os.path.join(d, stem + ".bin"), "wb") as f:
                f.write(data)
            doc["buffers"].append({"byteLength": len(data), "uri": stem + ".bin"})
        with open(path, "w") as f:
            json.dump(doc, f, indent=1)
    return path
