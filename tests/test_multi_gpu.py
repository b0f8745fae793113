"""Several GPUs behind the C ABI (include/ptrs_b200.h: ptrs_multi_*, ptrs_comm_*, ptrs_film_reduce).

On a one-GPU box these run with one device (the whole path — per-device thread, stream, film, NCCL communicator of
one rank — is still exercised); with `gpurun --gpus 2` (or more) the same tests shard the sample numbers over two
devices and reduce the films over NVLink.  The world-size-2 CPU test of the decomposition is tests/test_distributed_cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_devices(gpu):
    return min(2, gpu.device_count())


def test_multi_render_equals_single_device_render(gpu, host, cornell_env):
    flat, cam = cornell_env
    n = _n_devices(gpu)
    params = host.default_render_params(spp=16, max_depth=8)
    multi = gpu.MultiScene(flat, n)
    film, stats, ms = multi.render(cam, params)
    assert len(stats) == n and ms > 0
    scene = gpu.RenderScene(flat)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(16), max_depth=8)
    single = gpu.Film(cam.width, cam.height)
    st = integ.render(cam, scene, single)
    want = single.download()
    assert sum(s["camera_paths"] for s in stats) == st["camera_paths"]
    assert [s["camera_paths"] for s in stats] == [st["camera_paths"] // n] * n  # 16 spp deal evenly
    # same terms, summed in another order (atomics within a device, the reduce across devices)
    assert np.allclose(film[..., 3], want[..., 3], rtol=2e-5)
    assert np.allclose(film, want, rtol=2e-4, atol=1e-5 * float(want[..., :3].max()))
    assert np.allclose(multi.to_channel_updates(cam.width, cam.height), single.to_channel_updates(), rtol=2e-4, atol=1e-5)
    # the caller's own (stride, phase) selection composes with the per-device deal
    p2 = host.default_render_params(spp=16, max_depth=8)
    p2.sample_stride, p2.sample_phase = 2, 1
    half, stats2, _ = multi.render(cam, p2)
    ref = gpu.Film(cam.width, cam.height)
    integ.render(cam, scene, ref, sample_stride=(2, 1))
    assert sum(s["camera_paths"] for s in stats2) == st["camera_paths"] // 2
    assert np.allclose(half, ref.download(), rtol=2e-4, atol=1e-5 * float(want[..., :3].max()))
    # a second render at another resolution re-creates the films
    flat2, cam2 = host.make_scene(host.SCENE_CORNELL, res=(40, 24))
    del flat2
    film3, _, _ = multi.render(cam2, params)
    assert film3.shape == (24, 40, 4) and (film3[..., 3] > 0).all()
    multi.close()
    scene.close()


def test_multi_create_rejects_bad_device_lists(gpu, host, cornell):
    flat, _ = cornell
    for n, devs in ((gpu.device_count() + 1, None), (2, [0, 0]), (1, [gpu.device_count()]), (0, None)):
        with pytest.raises(gpu.PtrsError) as e:
            gpu.MultiScene(flat, n, devices=devs)
        assert e.value.code == -1


def test_comm_of_one_rank_reduces_in_place(gpu, host, cornell):
    """ptrs_comm_unique_id / ptrs_comm_init_rank / ptrs_film_reduce with a world of one: NCCL is loaded and called, and the
    film is unchanged (the sum over one rank)."""
    flat, cam = cornell
    scene = gpu.RenderScene(flat)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(4), max_depth=5)
    film = gpu.Film(cam.width, cam.height)
    integ.render(cam, scene, film)
    before = film.download()
    uid = gpu.Comm.unique_id()
    assert len(uid) == 128
    comm = gpu.Comm(uid, 1, 0)
    comm.reduce_film(film, root=0)
    assert np.array_equal(film.download(), before)
    with pytest.raises(gpu.PtrsError):
        comm.reduce_film(film, root=1)
    comm.close()
    scene.close()
