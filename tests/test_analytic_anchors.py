"""Analytic anchors for the whole integrator, independent of any reading of the reference: a convex body inside a
uniform white environment ("white furnace").  Every path that hits the body leaves it after one bounce and sees
radiance 1, so a Lambertian cube of albedo rho must render as rho, a perfect mirror as 1, a pane of glass as 1
(Fresnel reflection + transmission lose nothing, the radiance scaling of entering and leaving cancels) and the
background as 1 —
whatever the split between light sampling, BSDF sampling, MIS weights and Russian roulette
(src/pathtracer/integrator.rs:23-139, 401-499), and whatever the env-map importance sampling does
(light.rs:402-461, sampling.rs:128-230).  Run on the CPU oracle here and on the CUDA path under -m gpu."""
import numpy as np
import pytest

XML = """<scene version="0.5.0">
  <sensor type="perspective"><float name="fov" value="40"/>
    <transform name="toWorld"><matrix value="-1 0 0 0 0 1 0 0.4 0 0 -1 4 0 0 0 1"/></transform>
    <film type="ldrfilm"><integer name="width" value="48"/><integer name="height" value="48"/></film></sensor>
  {bsdf}
  <shape type="cube"><transform name="toWorld"><matrix value="{matrix}"/></transform><ref id="m"/></shape>
  <emitter type="envmap"><string name="filename" value="white.hdr"/><transform name="toWorld"><matrix value="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1"/></transform></emitter>
</scene>"""
MATTE = '<bsdf type="diffuse" id="m"><rgb name="reflectance" value="0.25, 0.5, 0.75"/></bsdf>'
MIRROR = '<bsdf type="conductor" id="m"><string name="material" value="none"/></bsdf>'
GLASS = '<bsdf type="dielectric" id="m"><float name="intIOR" value="1.5"/><float name="extIOR" value="1"/></bsdf>'
CUBE = "0.6 0 0.35 0 0 0.7 0 0 -0.35 0 0.6 0 0 0 0 1"  # rotated about y, convex: one bounce and out
SLAB = "0.9 0 0 0 0 0.9 0 0.3 0 0 0.04 0 0 0 0 1"  # a thin pane facing the camera: in through one face, out through the other
CASES = [(MATTE, CUBE, (0.25, 0.5, 0.75), 8), (MIRROR, CUBE, (1.0, 1.0, 1.0), 8), (GLASS, SLAB, (1.0, 1.0, 1.0), 30)]
IDS = ["matte", "mirror", "glass_pane"]


def _scene(host, tmp_path, bsdf, matrix):
    host.save_hdr(str(tmp_path / "white.hdr"), np.ones((8, 16, 3), dtype=np.float32))
    (tmp_path / "f.xml").write_text(XML.format(bsdf=bsdf, matrix=matrix))
    return host.import_scene(str(tmp_path / "f.xml"), res=(48, 48))


def _interior_mask(hit):
    """pixels whose whole 5 x 5 filter footprint is on the same side (all hit / all miss)"""
    h = hit.astype(np.int32)
    pad = np.pad(h, 2, mode="edge")
    s = sum(pad[dy:dy + h.shape[0], dx:dx + h.shape[1]] for dy in range(5) for dx in range(5))
    return s == 25, s == 0


def _check(host, cam, flat, rgb, intersect, expect_on_body):
    rays = host.coherent_rays(cam, 48)  # pixel-centre rays, scanline order
    hit = (intersect(rays)["prim"] >= 0).reshape(48, 48)
    body, sky = _interior_mask(hit)
    assert body.sum() > 150 and sky.sum() > 300
    assert np.allclose(rgb[sky], 1.0, rtol=2e-3)  # camera rays that miss read the map directly
    on = rgb[body]
    assert np.allclose(on.mean(axis=0), expect_on_body, rtol=1e-2), on.mean(axis=0)
    assert np.allclose(on, np.broadcast_to(expect_on_body, on.shape), rtol=0.12)


@pytest.mark.parametrize("bsdf,matrix,expect,depth", CASES, ids=IDS)
def test_white_furnace_oracle(host, oracle, tmp_path, bsdf, matrix, expect, depth):
    flat, cam = _scene(host, tmp_path, bsdf, matrix)
    params = host.default_render_params(spp=64, max_depth=depth)
    film, _ = oracle.render(flat, cam, params)
    rgb = film[..., :3] / film[..., 3:]
    _check(host, cam, flat, rgb, lambda r: oracle.intersect(flat, r)[0], np.array(expect, dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("bsdf,matrix,expect,depth", CASES, ids=IDS)
def test_white_furnace_gpu(gpu, host, tmp_path, bsdf, matrix, expect, depth):
    flat, cam = _scene(host, tmp_path, bsdf, matrix)
    scene = gpu.RenderScene(flat)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(64), max_depth=depth)
    film = gpu.Film(cam.width, cam.height)
    integ.render(cam, scene, film)
    a = film.download()
    _check(host, cam, flat, a[..., :3] / a[..., 3:], scene.intersect, np.array(expect, dtype=np.float32))
    scene.close()


# ---- direct lighting from a rectangular diffuse emitter: closed-form irradiance ------------------------------------------
# A horizontal one-sided emitter of radiance Le at height h above a Lambertian floor of albedo rho: with max_depth 1 the
# outgoing radiance of a floor point is rho * Le * F, F the point-to-rectangle form factor for parallel planes
# (the classic corner formula, summed over the four corners).  Pins DiffuseAreaLight / Triangle::sample / pdf_at_point /
# the MIS combination (light.rs:262-288, shape.rs:62-72, 541-578, integrator.rs:23-139) against an independent answer.
AREA_XML = """<scene version="0.5.0">
  <sensor type="perspective"><float name="fov" value="40"/>
    <transform name="toWorld"><matrix value="-1 0 0 0 0 0.8660254 -0.5 2 0 -0.5 -0.8660254 4.5 0 0 0 1"/></transform>
    <film type="ldrfilm"><integer name="width" value="48"/><integer name="height" value="48"/></film></sensor>
  <bsdf type="diffuse" id="floor"><rgb name="reflectance" value="0.5, 0.6, 0.7"/></bsdf>
  <bsdf type="diffuse" id="black"><rgb name="reflectance" value="0, 0, 0"/></bsdf>
  <shape type="rectangle"><transform name="toWorld"><matrix value="4 0 0 0 0 0 4 0 0 -4 0 0 0 0 0 1"/></transform><ref id="floor"/></shape>
  <shape type="rectangle"><transform name="toWorld"><matrix value="0.75 0 0 0.2 0 0 -1 3 0 0.5 0 0.4 0 0 0 1"/></transform><ref id="black"/>
    <emitter type="area"><rgb name="radiance" value="10, 5, 2"/></emitter></shape>
</scene>"""


def _form_factor(px, pz, x1, x2, z1, z2, h):
    def g(x, z):
        a, b = np.sqrt(x * x + h * h), np.sqrt(z * z + h * h)
        return (x / a * np.arctan(z / a) + z / b * np.arctan(x / b)) / (2 * np.pi)
    return g(x2 - px, z2 - pz) - g(x1 - px, z2 - pz) - g(x2 - px, z1 - pz) + g(x1 - px, z1 - pz)


def _check_area(host, cam, rgb, intersect):
    rays = host.coherent_rays(cam, 48)
    hits = intersect(rays)
    on_floor = (hits["prim"] >= 0).reshape(48, 48)
    p = (rays["o"] + rays["d"] * hits["t"][:, None]).reshape(48, 48, 3).astype(np.float64)
    on_floor &= np.abs(p[..., 1]) < 1e-4  # the floor, not the emitter
    inner, _ = _interior_mask(on_floor)
    assert inner.sum() > 1200
    f = _form_factor(p[..., 0], p[..., 2], 0.2 - 0.75, 0.2 + 0.75, 0.4 - 0.5, 0.4 + 0.5, 3.0)
    want = f[..., None] * np.array([0.5 * 10, 0.6 * 5, 0.7 * 2])
    got = rgb[inner].astype(np.float64)
    assert want[inner].max() > 0.15  # the lit region is in view
    assert np.allclose(got.mean(axis=0), want[inner].mean(axis=0), rtol=5e-3)
    assert np.allclose(got, want[inner], rtol=0.05, atol=2e-3)


def test_rectangle_light_closed_form_oracle(host, oracle, tmp_path):
    (tmp_path / "a.xml").write_text(AREA_XML)
    flat, cam = host.import_scene(str(tmp_path / "a.xml"), res=(48, 48))
    params = host.default_render_params(spp=256, max_depth=1)
    film, _ = oracle.render(flat, cam, params)
    _check_area(host, cam, film[..., :3] / film[..., 3:], lambda r: oracle.intersect(flat, r)[0])


@pytest.mark.gpu
def test_rectangle_light_closed_form_gpu(gpu, host, tmp_path):
    (tmp_path / "a.xml").write_text(AREA_XML)
    flat, cam = host.import_scene(str(tmp_path / "a.xml"), res=(48, 48))
    scene = gpu.RenderScene(flat)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(256), max_depth=1)
    film = gpu.Film(cam.width, cam.height)
    integ.render(cam, scene, film)
    a = film.download()
    _check_area(host, cam, a[..., :3] / a[..., 3:], scene.intersect)
    scene.close()
