"""Golden vectors (tests/golden/, produced by tests/golden/make_golden.py from the oracle).
CPU part: the oracle and the host-side scene code still reproduce them bit for bit.
GPU part: the CUDA path reproduces them through the C ABI without needing the oracle at run time."""
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(HERE, name))


def _sobol_cases(host):
    g = _load("sobol.npz")
    for tag, res, spp in (("c1", (512, 512), 16), ("c5", (3840, 2160), 1024)):
        cam = host.look_at_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, *res)
        yield cam, host.default_render_params(spp=spp), g[f"{tag}_px"], g[f"{tag}_sm"], g[f"{tag}_bits"], g[f"{tag}_index"]


def test_oracle_reproduces_sobol_golden(host, oracle):
    for cam, params, px, sm, bits, index in _sobol_cases(host):
        v, idx = oracle.sobol_samples(cam, params, px, sm, np.arange(48, dtype=np.int32))
        assert np.array_equal(v.view(np.uint32), bits) and np.array_equal(idx, index)


def test_oracle_reproduces_cornell_golden(host, oracle):
    flat, cam = host.make_scene(host.SCENE_CORNELL, res=(32, 32))
    g = _load("cornell_hits.npz")
    assert np.array_equal(flat.nodes().view(np.uint8), g["nodes"])  # host BVH build unchanged
    rays = g["rays"].view(host.RAY_DTYPE)
    hits, ctr = oracle.intersect(flat, rays)
    occ, ctr_p = oracle.intersect_p(flat, rays)
    assert np.array_equal(hits.view(np.uint8), g["hits"]) and np.array_equal(occ, g["occluded"])
    assert [*ctr, *ctr_p] == g["counters"].tolist()
    r = _load("cornell_render.npz")
    params = host.default_render_params(spp=8, max_depth=15)
    assert np.array_equal(oracle.path_radiance(flat, cam, params, r["px"], r["sm"]), r["radiance"])
    film, st = oracle.render(flat, cam, params, n_threads=1)
    assert np.array_equal(film, r["film"])
    assert [st[k] for k in ("camera_paths", "extension_rays", "shadow_rays", "mis_rays")] == r["stats"].tolist()
    film8, _ = oracle.render(flat, cam, params, n_threads=8)  # thread count must not change the result
    assert np.array_equal(film8, r["film"])


@pytest.mark.gpu
def test_gpu_reproduces_sobol_golden(gpu, host):
    for cam, params, px, sm, bits, index in _sobol_cases(host):
        v, idx = gpu.sobol_samples(cam, params, px, sm, np.arange(48, dtype=np.int32))
        assert np.array_equal(v.view(np.uint32), bits) and np.array_equal(idx, index)


@pytest.mark.gpu
def test_gpu_reproduces_cornell_golden(gpu, host):
    flat, cam = host.make_scene(host.SCENE_CORNELL, res=(32, 32))
    scene = gpu.RenderScene(flat)
    g = _load("cornell_hits.npz")
    rays = g["rays"].view(host.RAY_DTYPE)
    assert np.array_equal(scene.intersect(rays).view(np.uint8), g["hits"])  # ids, t and barycentrics bit for bit
    assert np.array_equal(scene.intersect(rays)["prim"], g["hits"].view(host.HIT_DTYPE)["prim"])
    assert np.array_equal(scene.intersect_p(rays), g["occluded"])
    r = _load("cornell_render.npz")
    params = host.default_render_params(spp=8, max_depth=15)
    rad = scene.path_radiance(cam, params, r["px"], r["sm"])
    close = np.isclose(rad, r["radiance"], rtol=1e-3, atol=1e-5).all(axis=1)
    assert close.mean() > 0.98
    film = gpu.Film(32, 32)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(8), max_depth=15)
    st = integ.render(cam, scene, film)
    out = film.download()
    ref = r["film"]
    assert np.allclose(out[..., 3], ref[..., 3], rtol=1e-5)
    img, rimg = out[..., :3] / out[..., 3:], ref[..., :3] / ref[..., 3:]
    assert float(np.mean((img - rimg) ** 2 / (rimg ** 2 + 1e-2))) < 1e-3
    assert st["camera_paths"] == int(r["stats"][0])
    scene.close()
