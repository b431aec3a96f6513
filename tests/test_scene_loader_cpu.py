"""The glTF scene path of the C++ host (host/scene_loader.cpp <- src/scene/scene_loader.cpp:40-349), CPU only: a procedural
scene is written as glTF 2.0 (.gltf + .bin + .png, embedded .gltf, .glb) and must come back as the same flat arrays, material
mapping, texture formats / samplers, camera and light the reference's loader would hand ResourceManager::UpdateGeometry."""
import io
import json
import os

import numpy as np
import pytest

from vulkanhybridrenderer_b200 import camera, capi, gltf_export, host_api, scenes
from vulkanhybridrenderer_b200 import types as T


@pytest.fixture(autouse=True, params=["cgltf", "own"])
def parser(request, monkeypatch):
    """Every test of this file runs twice: through the reference's vendored cgltf.h (the default when libvhr_host.so was built next to the
    reference tree) and through the loader's own JSON / .glb / accessor reader (VHR_GLTF_PARSER=own, the only one otherwise)."""
    if request.param == "cgltf" and not host_api.has_cgltf():
        pytest.skip("libvhr_host.so was built without the reference's vendored cgltf.h")
    monkeypatch.setenv("VHR_GLTF_PARSER", request.param)
    return request.param


@pytest.fixture(scope="module")
def scene():
    sc = scenes.add_procedural_textures(scenes.sponza_like(3000, seed=5, width=96, height=64, n_clutter=10), size=32)
    return sc


def _first_use_order(sm):
    order = []
    for g in range(len(sm)):
        for key in ("base_color_texture", "metallic_roughness_texture", "normal_map"):
            t = int(sm[key][g])
            if t >= 0 and t not in order:
                order.append(t)
    return order


@pytest.mark.parametrize("name,embed", [("scene.gltf", False), ("embedded.gltf", True), ("scene.glb", False)])
def test_round_trip(tmp_path, scene, name, embed):
    path = gltf_export.export(scene, tmp_path / name, embed=embed)
    got = host_api.parse_gltf(path)
    # geometry: one mesh per primitive, vertices / indices concatenated in node order, offsets rebuilt by the loader
    assert np.array_equal(got["vertices"].view(np.uint8), np.ascontiguousarray(scene.vertices).view(np.uint8))
    assert np.array_equal(got["indices"], scene.indices)
    assert len(got["primitives"]) == len(scene.primitives) and np.all(got["prims_per_mesh"] == 1)
    for key in ("vertex_offset", "index_offset", "index_count"):
        assert np.array_equal(got["primitives"][key], scene.primitives[key]), key
    assert np.array_equal(got["primitives"]["transform"], scene.primitives["transform"])
    # materials (scene_loader.cpp:182-218)
    gm, sm = got["primitives"]["material"], scene.primitives["material"]
    order = _first_use_order(sm)          # the loader uploads (and numbers) textures in order of first use
    for key in ("base_color_texture", "metallic_roughness_texture", "normal_map"):
        assert [int(x) for x in gm[key]] == [order.index(int(x)) if x >= 0 else -1 for x in sm[key]], key
    assert np.array_equal(gm["alpha_mask"], sm["alpha_mask"])
    assert np.allclose(gm["metallic_factor"], sm["metallic_factor"]) and np.allclose(gm["roughness_factor"], sm["roughness_factor"])
    untex = sm["base_color_texture"] < 0
    assert np.allclose(gm["base_color"][untex], sm["base_color"][untex])
    assert np.all(gm["base_color"][~untex] == 1.0)                      # with a texture the factor is ignored (:190-196)
    masked = sm["alpha_mask"] == 1
    assert np.allclose(gm["alpha_cutoff"][masked], sm["alpha_cutoff"][masked]) and np.all(gm["alpha_cutoff"][~masked] == 0.0)
    assert len(got["textures"]) == len(order)
    # formats: base colour sRGB, the others UNORM (scene_loader.cpp:249-272)
    for g in range(len(sm)):
        if sm["base_color_texture"][g] >= 0:
            assert got["textures"][gm["base_color_texture"][g]].format == T.VK_FORMAT_R8G8B8A8_SRGB
        if sm["metallic_roughness_texture"][g] >= 0:
            assert got["textures"][gm["metallic_roughness_texture"][g]].format == T.VK_FORMAT_R8G8B8A8_UNORM
        if sm["normal_map"][g] >= 0:
            assert got["textures"][gm["normal_map"][g]].format == T.VK_FORMAT_R8G8B8A8_UNORM
    # camera (scene_loader.cpp:43-72)
    cam = got["camera"]
    assert np.allclose(cam["perspective"], camera.to_glm(scene.camera.perspective), atol=1e-6)
    assert np.allclose(cam["transform"], camera.to_glm(scene.camera.transform), atol=2e-6)
    assert np.allclose(cam["view"], camera.to_glm(scene.camera.view), atol=2e-5)
    assert abs(float(cam["yaw"]) - (scene.camera.yaw % (2 * np.pi) - (2 * np.pi if scene.camera.yaw % (2 * np.pi) > np.pi else 0))) < 1e-5
    assert abs(float(cam["pitch"]) - scene.camera.pitch) < 1e-5 and abs(float(cam["roll"])) < 1e-5
    # light (scene_loader.cpp:74-100): direction from the node's world rotation (parent scale / translation ignored), intensity 30
    light = got["light"]
    d = np.asarray(scene.light["direction"][:3], np.float64)
    assert np.allclose(light["direction"][:3], d / np.linalg.norm(d), atol=1e-6) and light["direction"][3] == 0
    assert np.all(light["intensity"] == 30.0) and np.allclose(light["color"], [1, 1, 1, 1])
    pv = np.asarray(light["projview"], np.float64).T                      # row-major
    origin = pv @ np.array([0, 0, 0, 1.0])
    assert abs(origin[0]) < 1e-5 and abs(origin[1]) < 1e-5                # the light looks at the origin


def test_texture_slots_follow_first_use(tmp_path, scene):
    """Textures are uploaded in order of first use scanning meshes -> primitives: base colour, metallic-roughness, normal."""
    path = gltf_export.export(scene, tmp_path / "order.glb")
    got = host_api.parse_gltf(path)
    sm, gm = scene.primitives["material"], got["primitives"]["material"]
    order = _first_use_order(sm)
    assert len(got["textures"]) == len(order)
    for slot, src in enumerate(order):
        t, s = got["textures"][slot], scene.textures[src]
        assert np.array_equal(t.rgba, s.rgba), (slot, src)
        assert tuple(t.sampler) == tuple(s.sampler), (slot, src)
    for g in range(len(sm)):
        for key in ("base_color_texture", "metallic_roughness_texture", "normal_map"):
            assert int(gm[key][g]) == (order.index(int(sm[key][g])) if sm[key][g] >= 0 else -1)


def test_sampler_mapping_quirks(tmp_path, scene):
    """GetVkFilter / GetVkAddressMode (scene_loader.cpp:8-38): mip-mapped GL filters collapse, LINEAR_MIPMAP_NEAREST maps to NEAREST."""
    path = gltf_export.export(scene, tmp_path / "quirk.gltf")
    doc = json.load(open(path))
    doc["samplers"][0] = {"magFilter": 9729, "minFilter": 9985, "wrapS": 33071, "wrapT": 33648}      # LINEAR, LINEAR_MIPMAP_NEAREST
    doc["samplers"][1] = {"minFilter": 9987}                                                          # mag absent, LINEAR_MIPMAP_LINEAR, wraps default
    json.dump(doc, open(path, "w"))
    got = host_api.parse_gltf(path)
    by_content = {t.rgba.tobytes(): t for t in got["textures"]}
    t0, t1 = by_content[scene.textures[0].rgba.tobytes()], by_content[scene.textures[1].rgba.tobytes()]
    assert tuple(t0.sampler) == (capi.FILTER_LINEAR, capi.FILTER_NEAREST, capi.ADDRESS_MODE_CLAMP_TO_EDGE, capi.ADDRESS_MODE_MIRRORED_REPEAT)
    assert tuple(t1.sampler) == (capi.FILTER_LINEAR, capi.FILTER_LINEAR, capi.ADDRESS_MODE_REPEAT, capi.ADDRESS_MODE_REPEAT)


def test_default_light_and_errors(tmp_path, scene):
    path = gltf_export.export(scene, tmp_path / "nolight.glb", light_node=False)
    got = host_api.parse_gltf(path)
    assert np.allclose(got["light"]["direction"], [0, -1, 0.01, 0]) and np.allclose(got["light"]["color"], [1, 1, 1, 0])   # scene_loader.cpp:324-329
    with pytest.raises(capi.VhrError):
        host_api.parse_gltf(tmp_path / "missing.gltf")
    bad = tmp_path / "bad.gltf"
    bad.write_text('{"asset": {"version": "2.0"}, "meshes": [{"primitives": [{"attributes": {}}]}], "nodes": [{"mesh": 0}]}')
    with pytest.raises(capi.VhrError):
        host_api.parse_gltf(bad)                      # primitive without POSITION
    bad.write_text('{"asset": {"version": "1.0"}}')
    with pytest.raises(capi.VhrError):
        host_api.parse_gltf(bad)
    bad.write_text('{"asset": {"version": "2.0"}, "nodes": [')
    with pytest.raises(capi.VhrError):
        host_api.parse_gltf(bad)


@pytest.mark.parametrize("color_type,bit_depth", [(6, 8), (2, 8), (0, 8), (4, 8), (6, 16), (2, 16), (0, 16)])
def test_png_decoder(color_type, bit_depth):
    rng = np.random.default_rng(color_type * 31 + bit_depth)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    png = gltf_export.encode_png(img, color_type=color_type, bit_depth=bit_depth)
    got = host_api.decode_png(png)
    want = img.copy()
    if color_type in (0, 4):
        want[..., 1] = want[..., 2] = want[..., 0]
    if color_type in (0, 2):
        want[..., 3] = 255
    assert np.array_equal(got, want)
    # cross-check the test's own encoder against an independent decoder when one is installed
    try:
        from PIL import Image
    except ImportError:
        return
    if bit_depth == 8:
        pil = np.asarray(Image.open(io.BytesIO(png)).convert("RGBA"))
        assert np.array_equal(pil, want)


def test_png_from_independent_encoder():
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (40, 24, 4), dtype=np.uint8)
    for mode in ("RGBA", "RGB", "L", "LA", "P"):
        im = PIL.fromarray(img, "RGBA").convert(mode)
        buf = io.BytesIO()
        im.save(buf, format="PNG", optimize=(mode == "P"))
        got = host_api.decode_png(buf.getvalue())
        assert np.array_equal(got, np.asarray(im.convert("RGBA"))), mode
    with pytest.raises(capi.VhrError):
        host_api.decode_png(b"\xff\xd8\xff\xe0 not a png")      # JPEG magic: rejected loudly


def _hand_built_gltf(tmp_path, index_type):
    """Two triangles in one interleaved vertex buffer view (byteStride 28: POSITION f32x3, NORMAL f32x3, TEXCOORD_0 normalised u16x2),
    indices as u8 / u16 / u32, the mesh under a child node (matrix) of a parent node (TRS), plus a second primitive without normals."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    nrm = np.array([[0, 0, 1]] * 4, np.float32)
    uv16 = np.array([[0, 0], [65535, 0], [0, 65535], [65535, 32768]], np.uint16)
    inter = bytearray()
    for i in range(4):
        inter += pos[i].tobytes() + nrm[i].tobytes() + uv16[i].tobytes()
    assert len(inter) == 4 * 28
    idx = np.array([0, 1, 2, 2, 1, 3], {5121: np.uint8, 5123: np.uint16, 5125: np.uint32}[index_type])
    pos2 = (pos + np.float32(10)).astype(np.float32)
    blob = bytes(inter)
    off_idx = len(blob); blob += idx.tobytes(); blob += b"\0" * (-len(blob) % 4)
    off_pos2 = len(blob); blob += pos2.tobytes()
    (tmp_path / "hand.bin").write_bytes(blob)
    q = [0.0, float(np.sin(np.pi / 4)), 0.0, float(np.cos(np.pi / 4))]          # 90 degrees about +Y
    child = np.eye(4, dtype=np.float32); child[:3, 3] = [0.5, 0.0, 0.0]          # translation, written column-major below
    doc = {
        "asset": {"version": "2.0"},
        "buffers": [{"uri": "hand.bin", "byteLength": len(blob)}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 4 * 28, "byteStride": 28},
                        {"buffer": 0, "byteOffset": off_idx, "byteLength": idx.nbytes},
                        {"buffer": 0, "byteOffset": off_pos2, "byteLength": pos2.nbytes}],
        "accessors": [{"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 4, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 24, "componentType": 5123, "normalized": True, "count": 4, "type": "VEC2"},
                      {"bufferView": 1, "componentType": index_type, "count": 6, "type": "SCALAR"},
                      {"bufferView": 2, "componentType": 5126, "count": 4, "type": "VEC3"}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3},
                                   {"attributes": {"POSITION": 4}, "indices": 3}]}],
        "nodes": [{"children": [1], "translation": [1.0, 2.0, 3.0], "rotation": q, "scale": [2.0, 2.0, 2.0]},
                  {"mesh": 0, "matrix": [float(x) for x in child.T.reshape(-1)]}],
        "scenes": [{"nodes": [0]}], "scene": 0,
    }
    path = tmp_path / "hand.gltf"
    path.write_text(json.dumps(doc))
    return path, pos, nrm, uv16, idx, pos2, child


@pytest.mark.parametrize("index_type", [5121, 5123, 5125])
def test_interleaved_views_index_types_and_node_hierarchy(tmp_path, index_type):
    path, pos, nrm, uv16, idx, pos2, child = _hand_built_gltf(tmp_path, index_type)
    got = host_api.parse_gltf(path)
    v = got["vertices"]
    assert len(v) == 8 and len(got["indices"]) == 12 and len(got["primitives"]) == 2
    np.testing.assert_array_equal(v["pos"][:4], pos)
    np.testing.assert_array_equal(v["normal"][:4], nrm)
    np.testing.assert_allclose(v["uv0"][:4], uv16.astype(np.float32) / np.float32(65535), rtol=0, atol=1e-7)
    np.testing.assert_array_equal(v["pos"][4:], pos2)
    assert np.all(v["normal"][4:] == 0) and np.all(v["uv0"][4:] == 0) and np.all(v["tangent"] == 0)     # absent attributes stay zero (:150-173)
    np.testing.assert_array_equal(got["indices"], np.concatenate([idx, idx]).astype(np.uint32))         # relative to vertex_offset
    p = got["primitives"]
    assert list(p["vertex_offset"]) == [0, 4] and list(p["index_offset"]) == [0, 6] and list(p["index_count"]) == [6, 6]
    # world transform = parent TRS * child matrix (column-major in Primitive.transform): T(1,2,3) * Ry(90) * S(2) * T(0.5,0,0)
    ry = np.array([[0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]], np.float64)
    trs = np.eye(4); trs[:3, 3] = [1, 2, 3]
    world = trs @ ry @ np.diag([2.0, 2.0, 2.0, 1.0]) @ child.astype(np.float64)
    for k in range(2):
        np.testing.assert_allclose(np.asarray(p["transform"][k], np.float64).reshape(4, 4).T, world, atol=1e-6)
    # no material: factors 1, no textures, opaque
    m = p["material"]
    assert np.all(m["base_color"] == 1.0) and np.all(m["base_color_texture"] == -1) and np.all(m["alpha_mask"] == 0)


def test_jpeg_textures_load_through_the_vendored_stb_image(tmp_path, scene):
    """Sponza / Bistro ship JPEG + PNG mixes (reference: stbi_load, scene_loader.cpp:284-290). With the reference's vendored stb_image.h on the
    include path at build time the loader decodes JPEG (baseline and progressive), external file and embedded buffer view alike."""
    if not host_api.has_stb_image():
        pytest.skip("libvhr_host.so was built without the reference's vendored stb_image.h (own PNG decoder only)")
    import io
    from PIL import Image
    path = gltf_export.export(scene, tmp_path / "jpeg.gltf")
    doc = json.load(open(path))
    want = {}
    for k, image in enumerate(doc["images"]):
        src = Image.open(tmp_path / image["uri"]).convert("RGB").resize((64, 48))
        buf = io.BytesIO()
        src.save(buf, format="JPEG", quality=92, progressive=bool(k & 1))
        fn = f"jpeg_tex{k}.jpg"
        (tmp_path / fn).write_bytes(buf.getvalue())
        image["uri"], image["mimeType"] = fn, "image/jpeg"
        want[k] = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"), np.int32)      # libjpeg's decode of the same bytes
    json.dump(doc, open(path, "w"))
    got = host_api.parse_gltf(path)
    assert len(got["textures"]) == len(scene.textures)
    matched = 0
    for t in got["textures"]:
        assert t.rgba.shape == (48, 64, 4) and (t.rgba[..., 3] == 255).all()
        # the two IDCT / upsampling implementations differ by a few codes: find the source image this texture decodes
        errs = [float(np.abs(t.rgba[..., :3].astype(np.int32) - w).mean()) for w in want.values()]
        assert min(errs) < 2.0, errs
        matched += 1
    assert matched == len(scene.textures)


def _same_parse(a, b):
    for key in ("vertices", "indices", "primitives", "prims_per_mesh", "camera", "light"):
        assert np.array_equal(np.ascontiguousarray(a[key]).view(np.uint8), np.ascontiguousarray(b[key]).view(np.uint8)), key
    assert len(a["textures"]) == len(b["textures"])
    for ta, tb in zip(a["textures"], b["textures"]):
        assert np.array_equal(ta.rgba, tb.rgba) and ta.format == tb.format and tuple(ta.sampler) == tuple(tb.sampler)


@pytest.mark.parametrize("name,embed", [("scene.gltf", False), ("embedded.gltf", True), ("scene.glb", False)])
def test_cgltf_and_the_own_reader_agree_byte_for_byte(tmp_path, scene, monkeypatch, name, embed):
    """The two parsers behind SceneLoader::ParseScene hand back identical bytes: vertices, indices, primitives (transforms from
    cgltf_node_transform_world against the loader's own TRS product), camera, light, decoded textures, formats and samplers — on the
    exported procedural scene in its three containers and on the hand-built file with interleaved views and a node hierarchy."""
    if not host_api.has_cgltf():
        pytest.skip("libvhr_host.so was built without the reference's vendored cgltf.h")
    paths = [gltf_export.export(scene, tmp_path / name, embed=embed)]
    if not embed and name.endswith(".gltf"):
        paths += [_hand_built_gltf(tmp_path, t)[0] for t in (5121, 5125)]
    for path in paths:
        monkeypatch.setenv("VHR_GLTF_PARSER", "cgltf")
        a = host_api.parse_gltf(path)
        monkeypatch.setenv("VHR_GLTF_PARSER", "own")
        b = host_api.parse_gltf(path)
        _same_parse(a, b)
