"""JPEG textures (pathtracer_rs_b200/host/jpeg_decode.cpp): what image::open hands the importers for .jpg files
(src/pathtracer/importer/gltf.rs:39-96, src/pathtracer/importer/mitsuba.rs:104-117).  The decoder follows the IJG
arithmetic, so it is checked bit for bit against libjpeg-turbo: committed files + pixels (tests/golden/jpeg.npz,
made by tests/golden/make_jpeg_golden.py) and, when Pillow is importable, a sweep over sizes / subsampling /
progressive / quality / restart intervals."""
import base64
import io
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg.npz")


def _igc(v):
    v = np.asarray(v, dtype=np.float32)
    return np.where(v <= 0.04045, v / np.float32(12.92), ((v + np.float32(0.055)) / np.float32(1.055)) ** np.float32(2.4))


def test_golden_files_decode_to_libjpeg_pixels(host):
    g = np.load(GOLDEN)
    names = sorted(k[:-5] for k in g.files if k.endswith("_file"))
    assert len(names) == 6
    for n in names:
        got = host.decode_image(g[n + "_file"].tobytes())
        assert got.shape == g[n + "_pixels"].shape, n
        assert np.array_equal(got, g[n + "_pixels"]), n


def test_sweep_against_pillow(host):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    n = 0
    for (w, h) in [(64, 48), (37, 29), (16, 16), (8, 8), (1, 1), (5, 3), (2, 7), (4, 7), (130, 71)]:
        for mode in ("smooth", "noise"):
            y, x = np.mgrid[0:h, 0:w]
            px = np.stack([128 + 100 * np.sin(x / 7.0) * np.cos(y / 5.0), 128 + 90 * np.sin((x + y) / 11.0), (x * 3 + y * 2) % 256], -1)
            px = np.clip(px + rng.normal(0, 12, px.shape), 0, 255).astype(np.uint8) if mode == "smooth" else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            for grey in (False, True):
                for sub in ((0, 1, 2) if not grey else (0,)):
                    for prog in (False, True):
                        for q, rst in ((50, 0), (90, 2), (100, 0)):
                            kw = dict(quality=q, progressive=prog)
                            if not grey:
                                kw["subsampling"] = sub
                            if rst:
                                kw["restart_marker_blocks"] = rst
                            b = io.BytesIO()
                            try:
                                Image.fromarray(px[..., 0] if grey else px).save(b, "JPEG", **kw)
                            except OSError:  # Pillow's own output buffer is too small for some tiny progressive files
                                continue
                            ref = np.asarray(Image.open(io.BytesIO(b.getvalue())))
                            got = host.decode_image(b.getvalue())
                            assert np.array_equal(got, ref if ref.ndim == 3 else ref[..., None]), (w, h, mode, grey, sub, prog, q, rst)
                            n += 1
    assert n > 400


def test_unsupported_and_broken_files_raise(host):
    g = np.load(GOLDEN)
    data = g["baseline_420_file"].tobytes()
    with pytest.raises(RuntimeError):
        host.decode_image(b"\xff\xd8\xff\xdb\x00\x03")  # truncated segment
    with pytest.raises(RuntimeError):
        host.decode_image(data[:2] + b"\xff\xd9")  # no frame
    sof = data.index(b"\xff\xc0")
    with pytest.raises(RuntimeError):  # 12-bit precision
        host.decode_image(data[:sof + 4] + b"\x0c" + data[sof + 5:])
    with pytest.raises(RuntimeError):  # arithmetic-coded frame marker
        host.decode_image(data[:sof + 1] + b"\xc9" + data[sof + 2:])
    cut = host.decode_image(data[: len(data) - 40])  # entropy data cut short: what was decoded is returned (libjpeg warns and does the same)
    assert cut.shape == g["baseline_420_pixels"].shape
    Image = pytest.importorskip("PIL.Image")
    b = io.BytesIO()
    Image.fromarray(np.zeros((8, 8, 4), dtype=np.uint8), "CMYK").save(b, "JPEG")
    with pytest.raises(RuntimeError):
        host.decode_image(b.getvalue())


def test_corrupted_files_never_crash(host):
    """Every byte of a file is attacker-controlled (textures come from scene files): flipped bytes either still decode
    to an image of the declared size or raise — tables, markers and entropy data are all hit by the mutations."""
    g = np.load(GOLDEN)
    rng = np.random.default_rng(11)
    outcomes = {"ok": 0, "error": 0}
    for name in ("baseline_420", "progressive_420", "baseline_422_restart", "grey"):
        data = bytearray(g[name + "_file"].tobytes())
        for _ in range(150):
            m = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                m[int(rng.integers(2, len(m)))] = int(rng.integers(0, 256))
            try:
                img = host.decode_image(bytes(m))
                assert img.ndim == 3 and img.size > 0
                outcomes["ok"] += 1
            except RuntimeError:
                outcomes["error"] += 1
    assert outcomes["ok"] > 50 and outcomes["error"] > 50, outcomes


XML = """<scene version="0.5.0">
  <sensor type="perspective"><float name="fov" value="40"/><transform name="toWorld"><matrix value="-1 0 0 0 0 1 0 0 0 0 -1 4 0 0 0 1"/></transform>
    <film type="ldrfilm"><integer name="width" value="64"/><integer name="height" value="48"/></film></sensor>
  <bsdf type="diffuse" id="image"><texture type="bitmap"><string name="filename" value="tex.jpg"/></texture></bsdf>
  <shape type="rectangle"><transform name="toWorld"><matrix value="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1"/></transform><ref id="image"/></shape>
</scene>"""


def test_mitsuba_bitmap_texture_from_a_jpeg(host, tmp_path):
    g = np.load(GOLDEN)
    (tmp_path / "tex.jpg").write_bytes(g["progressive_420_file"].tobytes())
    (tmp_path / "s.xml").write_text(XML)
    flat, cam = host.import_scene(str(tmp_path / "s.xml"))
    d = flat.desc.contents
    mats = [d.materials[i] for i in range(d.n_materials)]
    tex = d.textures[mats[0].tex[0]]
    assert tex.type == host.TEX_IMAGE
    mm = d.mipmaps[tex.mip]
    px = g["progressive_420_pixels"]
    # MIPMap::new resamples to powers of two (texture.rs:285-295), so compare through the host's own PNG path: the same
    # pixels saved as PNG must give the same pyramid
    host.save_png(str(tmp_path / "tex.png"), px)
    (tmp_path / "p.xml").write_text(XML.replace("tex.jpg", "tex.png"))
    flat2, _ = host.import_scene(str(tmp_path / "p.xml"))
    d2 = flat2.desc.contents
    assert d.n_texels == d2.n_texels and mm.width[0] == d2.mipmaps[tex.mip].width[0]
    assert np.array_equal(np.ctypeslib.as_array(d.texels, shape=(d.n_texels,)), np.ctypeslib.as_array(d2.texels, shape=(d2.n_texels,)))


def test_gltf_base_color_texture_from_a_jpeg(host, tmp_path):
    g = np.load(GOLDEN)
    px = g["baseline_422_restart_pixels"]
    uri = "data:image/jpeg;base64," + base64.b64encode(g["baseline_422_restart_file"].tobytes()).decode()
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    uv = np.array([[0, 0], [1, 0], [0, 1]], dtype=np.float32)
    idx = np.array([0, 1, 2, 0], dtype=np.uint16)  # 3 indices + padding
    blob = pos.tobytes() + uv.tobytes() + idx.tobytes()
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0}]}],
        "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
        "textures": [{"source": 0}], "images": [{"uri": uri}],
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24},
                        {"buffer": 0, "byteOffset": 60, "byteLength": 6}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
                      {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"},
                      {"bufferView": 2, "componentType": 5123, "count": 3, "type": "SCALAR"}],
    }
    (tmp_path / "t.gltf").write_text(json.dumps(doc))
    flat, cam = host.import_scene(str(tmp_path / "t.gltf"))
    d = flat.desc.contents
    assert any(d.textures[i].type == host.TEX_IMAGE for i in range(d.n_textures)), "the JPEG base-colour texture was not imported"
    # the same pixels as a PNG image must give the same scene, texel for texel
    host.save_png(str(tmp_path / "t.png"), px)
    doc["images"] = [{"uri": "data:image/png;base64," + base64.b64encode((tmp_path / "t.png").read_bytes()).decode()}]
    (tmp_path / "p.gltf").write_text(json.dumps(doc))
    flat2, _ = host.import_scene(str(tmp_path / "p.gltf"))
    d2 = flat2.desc.contents
    assert d.n_texels == d2.n_texels and d.n_textures == d2.n_textures
    assert np.array_equal(np.ctypeslib.as_array(d.texels, shape=(d.n_texels,)), np.ctypeslib.as_array(d2.texels, shape=(d2.n_texels,)))
