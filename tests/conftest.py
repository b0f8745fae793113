import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The native libraries must exist; build them if a fresh checkout has none."""
    need = [os.path.join(ROOT, "pathtracer_rs_b200", "lib", "libptrs_b200.so"),
            os.path.join(ROOT, "pathtracer_rs_b200", "lib", "libptrs_host.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def host():
    import pathtracer_rs_b200.host as h

    return h


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def gpu():
    import pathtracer_rs_b200.gpu as g

    if g.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    g.set_device(0)
    return g


@pytest.fixture(scope="session")
def cornell(host):
    return host.make_scene(host.SCENE_CORNELL, res=(96, 96))


@pytest.fixture(scope="session")
def cornell_env(host):
    """BASELINE configs[1]'s scene: Cornell box + the reference's own environment map (data/abandoned_tank_farm_04_1k.hdr,
    committed as a fixture), 2048 x 1024 Distribution2D per light.rs:375-387."""
    assert os.path.exists(host.TANK_FARM_HDR)
    return host.make_scene(host.SCENE_CORNELL_ENV, seed=1, res=(96, 96), env_hdr=host.TANK_FARM_HDR)


@pytest.fixture(scope="session")
def cornell_sky(host):
    """The same box under the synthetic sky the larger procedural scenes (C3, C5) use."""
    return host.make_scene(host.SCENE_CORNELL_ENV, seed=1, res=(96, 96))


@pytest.fixture(scope="session")
def field_small(host):
    """C3-shaped scene at ~60k triangles: glass / substrate / metal / Disney / matte + env + area lights."""
    return host.make_scene(host.SCENE_MATERIAL_FIELD, seed=1, n_tris=60000, res=(96, 64))


@pytest.fixture(scope="session")
def terrain_small(host):
    return host.make_scene(host.SCENE_TERRAIN, seed=1, n_tris=200000, res=(128, 128))


@pytest.fixture(scope="session")
def atrium_small(host):
    return host.make_scene(host.SCENE_ATRIUM, seed=1, n_tris=40000, res=(96, 54))
