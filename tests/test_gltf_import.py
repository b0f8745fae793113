"""glTF ingestion: src/common/importer/gltf.rs + src/pathtracer/importer/gltf.rs restated in
pathtracer_rs_b200/host/importer_gltf.cpp.  A small document exercising every branch is generated here (as .gltf with an
external .bin, with data URIs, and as .glb) and compared with what the reference's code does with it.  CPU only."""
import base64
import json
import struct

import numpy as np
import pytest


def _quat(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    return [*(a * np.sin(angle / 2)), float(np.cos(angle / 2))]


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _trs(t, q, s):
    m = np.eye(4)
    m[:3, :3] = _rot(q) @ np.diag(s)
    m[:3, 3] = t
    return m


QUAD_POS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=np.float32)
QUAD_N = np.tile(np.array([0, 0, 1], dtype=np.float32), (4, 1))
QUAD_UV = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32)
QUAD_TAN = np.tile(np.array([1, 0, 0, 1], dtype=np.float32), (4, 1))
QUAD_IDX = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint16)
TRI_POS = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0]], dtype=np.float32)
TRI_IDX = np.array([0, 1, 2], dtype=np.uint32)

ROOT_T, ROOT_Q, ROOT_S = [1.0, 2.0, -3.0], _quat([0, 1, 0], 0.7), [2.0, 2.0, 2.0]
CHILD_T = [0.5, 0.0, 0.0]
MATRIX_NODE = _trs([0.0, -1.0, 0.0], _quat([1, 0, 0], -np.pi / 2), [3.0, 1.0, 3.0])
CAM_T, CAM_Q = [0.0, 1.0, 8.0], _quat([0, 1, 0], 0.0)


def build_document(host, tmp_path, embed=False):
    """Returns (gltf json dict, binary blob).  Images are PNG files written next to the document (or data URIs)."""
    rng = np.random.default_rng(11)
    base = rng.integers(1, 256, (8, 8, 4), dtype=np.uint8)       # RGBA: alpha drives the MASK texture
    mr = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)         # G = roughness, B = metallic
    nrm = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)
    emis = np.zeros((8, 8, 3), dtype=np.uint8)
    emis[1, 6, :] = 200                                           # one bright interior texel at (u ~ 0.8, v ~ 0.2): only the quad's first triangle reaches it
    imgs = {"base.png": base, "mr.png": mr, "nrm.png": nrm, "emis.png": emis}
    for name, px in imgs.items():
        host.save_png(str(tmp_path / name), px)
    chunks, views = [], []

    def view(arr, stride=None):
        raw = arr.tobytes()
        off = sum(len(c) for c in chunks)
        pad = (-len(raw)) % 4
        chunks.append(raw + b"\0" * pad)
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(raw)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        return len(views) - 1

    acc = []

    def accessor(v, ctype, count, kind, **kw):
        acc.append({"bufferView": v, "componentType": ctype, "count": count, "type": kind, **kw})
        return len(acc) - 1

    a_qpos = accessor(view(QUAD_POS), 5126, 4, "VEC3")
    a_qn = accessor(view(QUAD_N), 5126, 4, "VEC3")
    a_quv = accessor(view(QUAD_UV), 5126, 4, "VEC2")
    a_qtan = accessor(view(QUAD_TAN), 5126, 4, "VEC4")
    a_qidx = accessor(view(QUAD_IDX), 5123, 6, "SCALAR")
    inter = np.zeros((3, 5), dtype=np.float32)  # interleaved positions with a 20-byte stride
    inter[:, :3] = TRI_POS
    a_tpos = accessor(view(inter, stride=20), 5126, 3, "VEC3")
    a_tidx = accessor(view(TRI_IDX), 5125, 3, "SCALAR")
    uv16 = (QUAD_UV * 65535).astype(np.uint16)
    a_quv16 = accessor(view(uv16), 5123, 4, "VEC2", normalized=True)

    def img(name):
        if embed:
            return {"uri": "data:image/png;base64," + base64.b64encode((tmp_path / name).read_bytes()).decode()}
        return {"uri": name}

    doc = {
        "asset": {"version": "2.0"},
        "scenes": [{"nodes": [0, 3, 4, 5]}],
        "nodes": [
            {"translation": ROOT_T, "rotation": ROOT_Q, "scale": ROOT_S, "children": [1, 2]},
            {"mesh": 0, "translation": CHILD_T},
            {"mesh": 1, "extensions": {"KHR_lights_punctual": {"light": 0}}},
            {"mesh": 2, "matrix": [float(x) for x in MATRIX_NODE.T.reshape(-1)]},
            {"camera": 0, "translation": CAM_T, "rotation": CAM_Q, "extensions": {"KHR_lights_punctual": {"light": 1}}},
            {"translation": [0.0, 5.0, 0.0], "extensions": {"KHR_lights_punctual": {"light": 2}}},
        ],
        "cameras": [{"type": "perspective", "perspective": {"yfov": 0.6, "znear": 0.05, "zfar": 200.0, "aspectRatio": 1.5}}],
        "meshes": [
            {"primitives": [{"attributes": {"POSITION": a_qpos, "NORMAL": a_qn, "TEXCOORD_0": a_quv, "TANGENT": a_qtan}, "indices": a_qidx, "material": 0},
                            {"attributes": {"POSITION": a_tpos}, "indices": a_tidx, "material": 1}]},
            {"primitives": [{"attributes": {"POSITION": a_qpos, "TEXCOORD_0": a_quv16}, "indices": a_qidx, "material": 4},
                            {"attributes": {"POSITION": a_tpos}, "indices": a_tidx, "material": 2}]},
            {"primitives": [{"attributes": {"POSITION": a_qpos}, "indices": a_qidx, "material": 3},
                            {"attributes": {"POSITION": a_tpos}, "indices": a_tidx, "material": 5},
                            {"attributes": {"POSITION": a_tpos}, "indices": a_tidx},
                            {"attributes": {"POSITION": a_qpos, "TEXCOORD_0": a_quv}, "indices": a_qidx, "material": 6}]},
        ],
        "materials": [
            {"pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.5, 0.25, 1.0], "baseColorTexture": {"index": 0}, "metallicFactor": 0.5, "roughnessFactor": 0.75,
                                      "metallicRoughnessTexture": {"index": 1}},
             "normalTexture": {"index": 2, "scale": 0.5}, "alphaMode": "MASK"},
            {"extensions": {"KHR_materials_transmission": {"transmissionFactor": 1.0}, "KHR_materials_ior": {"ior": 1.45}}},
            {"pbrMetallicRoughness": {"baseColorFactor": [0.6, 0.6, 0.6, 0.5]}, "alphaMode": "BLEND"},
            {"pbrMetallicRoughness": {"metallicFactor": 1.0, "roughnessFactor": 0.0}},
            {"emissiveFactor": [0.5, 0.9, 0.1], "emissiveTexture": {"index": 3}},
            {},
            {"emissiveFactor": [0.25, 0.0, 0.0]},
        ],
        "textures": [{"source": 0, "sampler": 0}, {"source": 1}, {"source": 2, "sampler": 1}, {"source": 3}],
        "samplers": [{"wrapS": 33071, "wrapT": 33071}, {"wrapS": 33648, "wrapT": 33648}],
        "images": [img("base.png"), img("mr.png"), img("nrm.png"), img("emis.png")],
        "extensions": {"KHR_lights_punctual": {"lights": [
            {"type": "point", "color": [0.5, 1.0, 1.0], "intensity": 40.0},
            {"type": "directional", "color": [1.0, 0.2, 0.2], "intensity": 3.0},
            {"type": "spot", "intensity": 7.0, "spot": {"innerConeAngle": 0.1, "outerConeAngle": 0.5}}]}},
        "accessors": acc, "bufferViews": views,
    }
    blob = b"".join(chunks)
    doc["buffers"] = [{"byteLength": len(blob)}]
    return doc, blob, imgs


def write_gltf(host, tmp_path, kind):
    doc, blob, imgs = build_document(host, tmp_path, embed=kind != "gltf")
    if kind == "gltf":
        doc["buffers"][0]["uri"] = "scene.bin"
        (tmp_path / "scene.bin").write_bytes(blob)
        (tmp_path / "scene.gltf").write_text(json.dumps(doc))
        return str(tmp_path / "scene.gltf"), imgs
    if kind == "embedded":
        doc["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
        (tmp_path / "embedded.gltf").write_text(json.dumps(doc, indent=1))
        return str(tmp_path / "embedded.gltf"), imgs
    js = json.dumps(doc).encode()
    js += b" " * ((-len(js)) % 4)
    blob += b"\0" * ((-len(blob)) % 4)
    body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
    (tmp_path / "scene.glb").write_bytes(b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body)
    return str(tmp_path / "scene.glb"), imgs


def _tris(pos, idx, m):
    p = (np.c_[pos.astype(np.float64), np.ones(len(pos))] @ m.T)[:, :3]
    return p[idx.reshape(-1, 3).astype(np.int64)]


def _tex(flat, i):
    return flat.desc.contents.textures[i]


def _level0(flat, tex):
    d = flat.desc.contents
    mm = d.mipmaps[tex.mip]
    n = mm.width[0] * mm.height[0] * mm.channels
    return np.ctypeslib.as_array(d.texels, shape=(d.n_texels,))[mm.level_offset[0]: mm.level_offset[0] + n].reshape(mm.height[0], mm.width[0], mm.channels)


def _igc(v):
    v = np.asarray(v, dtype=np.float32)
    return np.where(v <= 0.04045, v / np.float32(12.92), ((v + np.float32(0.055)) / np.float32(1.055)) ** np.float32(2.4))


@pytest.mark.parametrize("kind", ["gltf", "embedded", "glb"])
def test_gltf_scene(host, tmp_path, kind):
    path, imgs = write_gltf(host, tmp_path, kind)
    flat, cam = host.import_scene(path, res=(300, 200))
    d = flat.desc.contents
    # ---- geometry: node transforms t * r * s, children under parents, matrix nodes via decompose / recompose
    root = _trs(ROOT_T, ROOT_Q, ROOT_S)
    child = root @ _trs(CHILD_T, [0, 0, 0, 1], [1, 1, 1])
    want = np.concatenate([_tris(QUAD_POS, QUAD_IDX, child), _tris(TRI_POS, TRI_IDX, child),
                           _tris(QUAD_POS, QUAD_IDX, root), _tris(TRI_POS, TRI_IDX, root),
                           _tris(QUAD_POS, QUAD_IDX, MATRIX_NODE), _tris(TRI_POS, TRI_IDX, MATRIX_NODE), _tris(TRI_POS, TRI_IDX, MATRIX_NODE),
                           _tris(QUAD_POS, QUAD_IDX, MATRIX_NODE)])
    got = flat.prim_vertices().astype(np.float64)
    assert got.shape == want.shape == (12, 3, 3)
    key = lambda t: np.lexsort(np.round(t.reshape(len(t), -1), 4).T[::-1])  # noqa: E731
    assert np.allclose(got[key(got)], want[key(want)], atol=2e-5)
    # ---- materials: [default matte] + one per document material
    mats = [d.materials[i] for i in range(d.n_materials)]
    assert [m.type for m in mats] == [host.MAT_MATTE, host.MAT_DISNEY, host.MAT_GLASS, host.MAT_GLASS, host.MAT_MIRROR, host.MAT_DISNEY,
                                      host.MAT_DISNEY, host.MAT_DISNEY]
    m0 = mats[1]
    cf = _igc([0.8, 0.5, 0.25])  # Spectrum::from_slice_4(base_color_factor, gamma = true)
    color = _tex(flat, m0.tex[0])
    assert color.type == host.TEX_IMAGE
    assert np.allclose(_level0(flat, color), cf * _igc(imgs["base.png"][..., :3].astype(np.float32) / np.float32(255)), rtol=3e-6, atol=1e-7)
    assert d.mipmaps[color.mip].wrap == host.WRAP_CLAMP
    metallic, rough = _tex(flat, m0.tex[1]), _tex(flat, m0.tex[3])
    assert np.allclose(_level0(flat, metallic)[..., 0], np.float32(0.5) * (imgs["mr.png"][..., 2].astype(np.float32) / np.float32(255)))
    assert np.allclose(_level0(flat, rough)[..., 0], np.float32(0.75) * (imgs["mr.png"][..., 1].astype(np.float32) / np.float32(255)))
    assert _tex(flat, m0.tex[2]).v1[0] == 1.5  # default glTF ior
    nm = _tex(flat, m0.normal_map)
    n0 = imgs["nrm.png"].astype(np.float32) / np.float32(127.5) - np.float32(1)
    assert np.allclose(_level0(flat, nm), n0 * np.array([0.5, 0.5, 1.0], dtype=np.float32), atol=1e-6)
    assert d.mipmaps[nm.mip].wrap == host.WRAP_REPEAT  # mirrored repeat -> repeat
    glass = mats[2]
    assert _tex(flat, glass.tex[2]).v1[0] == np.float32(1.45) and list(_tex(flat, glass.tex[1]).v1) == [1.0, 1.0, 1.0]
    blend = mats[3]
    assert _tex(flat, blend.tex[2]).v1[0] == np.float32(1.33)
    assert np.allclose(list(_tex(flat, blend.tex[1]).v1), 1.0 - 0.5 * _igc([0.6, 0.6, 0.6]), rtol=1e-6)
    plain = mats[6]  # `{}`: metallic 1, roughness 1, white
    assert _tex(flat, plain.tex[1]).v1[0] == 1.0 and _tex(flat, plain.tex[3]).v1[0] == 1.0 and list(_tex(flat, plain.tex[0]).v1) == [1.0, 1.0, 1.0]
    # ---- alpha mask on the MASK primitive's mesh only
    alpha = [d.meshes[i].alpha_tex for i in range(d.n_meshes)]
    assert sum(a >= 0 for a in alpha) == 1 and alpha[0] >= 0
    assert np.allclose(_level0(flat, _tex(flat, alpha[0]))[..., 0], imgs["base.png"][..., 3].astype(np.float32) / np.float32(255))
    # ---- lights: area lights and point / spot lights in traversal order, then the deferred directional light
    lights = [d.lights[i] for i in range(d.n_lights)]
    types = [l.type for l in lights]
    # emissive-texture quad: only the triangle whose 10 x 10 sample points reach the bright texel emits; constant-ke quad: both
    assert types == [host.LIGHT_AREA, host.LIGHT_POINT, host.LIGHT_AREA, host.LIGHT_AREA, host.LIGHT_POINT, host.LIGHT_DIRECTIONAL]
    assert np.allclose(list(lights[1].color), [20.0, 20.0, 20.0])  # intensity * color[0] on all channels
    assert np.allclose(list(lights[1].pos), ROOT_T, atol=1e-6)
    ke_const = _tex(flat, lights[2].ke_tex)
    assert list(ke_const.v1) == [2.5, 2.5, 2.5]  # 10 x emissive_factor[0]
    assert list(lights[4].color) == [7.0, 7.0, 7.0] and np.allclose(list(lights[4].pos), [0, 5, 0])
    assert np.allclose(list(lights[5].color), [3.0, 3.0, 3.0]) and np.allclose(list(lights[5].pos), [0, 0, -1], atol=1e-6)
    ke_img = _tex(flat, lights[0].ke_tex)
    assert ke_img.type == host.TEX_IMAGE and np.allclose(_level0(flat, ke_img)[1, 6], np.float32(5.0) * _igc(np.float32(200) / np.float32(255)), rtol=3e-6)
    # ---- camera: the first camera found; aspect from -r, yfov / znear / zfar from the file
    assert (cam.width, cam.height) == (300, 200)
    assert np.allclose(list(cam.trans), CAM_T) and np.allclose(list(cam.rot), [0, 0, 0, 1], atol=1e-6)
    f = 1.0 / np.tan(0.3)
    assert np.allclose(list(cam.persp)[:2], [f / 1.5, f], rtol=1e-5)
    assert np.allclose(list(cam.persp)[2:], [(200.0 + 0.05) / (0.05 - 200.0), 2 * 200.0 * 0.05 / (0.05 - 200.0)], rtol=1e-5)


def test_gltf_default_camera_and_default_lights(host, tmp_path):
    doc, blob, _ = build_document(host, tmp_path, embed=True)
    doc["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
    doc["nodes"][4].pop("camera")
    doc.pop("cameras")
    (tmp_path / "nocam.gltf").write_text(json.dumps(doc))
    sky = host.synth_sky(32, 16, seed=2)
    host.save_hdr(str(tmp_path / "sky.hdr"), sky)
    flat, cam = host.import_scene(str(tmp_path / "nocam.gltf"), res=(400, 200), default_lights=True, sunsky_hdr=str(tmp_path / "sky.hdr"))
    d = flat.desc.contents
    bmin, bmax = flat.world_bound()
    assert np.allclose(list(cam.trans), bmax)  # get_default_camera: look from world_bound.p_max at the origin
    assert np.isclose(cam.persp[1], 1.0 / np.tan(0.5 * (np.pi / 2) * (200 / 400)), rtol=1e-5)
    assert d.lights[d.n_lights - 1].type == host.LIGHT_INFINITE and d.n_infinite_lights == 1
    l2w = np.array(d.envs[0].light_to_world).reshape(4, 4)
    assert np.allclose(l2w[:3, :3], [[1, 0, 0], [0, 0, 1], [0, -1, 0]], atol=1e-6)  # rotation by -pi/2 about x: z-up map, y-up scene


def test_gltf_errors(host, tmp_path):
    (tmp_path / "bad.gltf").write_text("{ not json")
    with pytest.raises(RuntimeError, match="JSON"):
        host.import_scene(str(tmp_path / "bad.gltf"))
    doc, blob, _ = build_document(host, tmp_path, embed=True)
    doc["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
    doc["samplers"][0]["wrapT"] = 10497
    (tmp_path / "wrap.gltf").write_text(json.dumps(doc))
    with pytest.raises(RuntimeError, match="wrapS != wrapT"):
        host.import_scene(str(tmp_path / "wrap.gltf"))
    doc["samplers"][0]["wrapT"] = 33071
    doc["meshes"][0]["primitives"][0].pop("indices")
    (tmp_path / "noidx.gltf").write_text(json.dumps(doc))
    with pytest.raises(RuntimeError, match="without indices"):
        host.import_scene(str(tmp_path / "noidx.gltf"))
