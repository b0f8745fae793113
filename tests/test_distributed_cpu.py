"""World-size-2 check of the multi-rank decomposition on CPU (gloo): each rank renders its shard of the
Sobol sample numbers (here with the CPU oracle standing in for the device), the films are summed with the
product's reduce_film(), and the result equals the single-rank render up to float summation order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pathtracer_rs_b200.host as host
    from oracle import oracle
    from pathtracer_rs_b200.dist import reduce_film, sample_shard

    flat, cam = host.make_scene(host.SCENE_CORNELL, res=(24, 24))
    params = host.default_render_params(spp=8, max_depth=5)
    params.sample_stride, params.sample_phase = sample_shard(rank, world)
    film, st = oracle.render(flat, cam, params, n_threads=1)
    t = torch.from_numpy(film)
    # the communicator id travels from rank 0 to every rank through the process group (what bench.py does under torchrun)
    from pathtracer_rs_b200.dist import exchange_comm_id

    uid = exchange_comm_id(lambda: bytes(range(128)), rank, world)
    assert uid == bytes(range(128))
    reduce_film(t, dst=0)
    paths = torch.tensor([st["camera_paths"]], dtype=torch.int64)
    dist.reduce(paths, dst=0)
    if rank == 0:
        np.savez(out_path, film=t.numpy(), paths=paths.numpy())
    dist.destroy_process_group()


def test_two_rank_sample_sharding_sums_to_the_full_render(tmp_path, host, oracle):
    from pathtracer_rs_b200.dist import sample_shard, shard_sample_counts

    assert [sample_shard(r, 4) for r in range(4)] == [(4, 0), (4, 1), (4, 2), (4, 3)]
    assert shard_sample_counts(64, 8) == [8] * 8 and sum(shard_sample_counts(10, 4)) == 10
    out = str(tmp_path / "film.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    flat, cam = host.make_scene(host.SCENE_CORNELL, res=(24, 24))
    params = host.default_render_params(spp=8, max_depth=5)
    full, st = oracle.render(flat, cam, params, n_threads=1)
    assert int(got["paths"][0]) == st["camera_paths"]
    assert np.allclose(got["film"], full, rtol=1e-5, atol=1e-6)


def test_strong_scaling_plan_and_id_exchange_through_a_store():
    from pathtracer_rs_b200.dist import exchange_comm_id, strong_scaling_plan

    for spp, world in ((128, 8), (128, 3), (5, 8)):
        plan = strong_scaling_plan(spp, world)
        assert sum(n for _, _, n in plan) == spp and max(n for _, _, n in plan) - min(n for _, _, n in plan) <= 1
        covered = sorted(s for stride, phase, _ in plan for s in range(phase, spp, stride))
        assert covered == list(range(spp))  # every sample number exactly once

    class Store(dict):
        def set(self, k, v):
            self[k] = v

        def get(self, k):
            return self[k]

    store = Store()
    made = []

    def make():
        made.append(1)
        return b"\x07" * 128

    assert exchange_comm_id(make, 0, 4, store) == b"\x07" * 128
    assert all(exchange_comm_id(make, r, 4, store) == b"\x07" * 128 for r in (1, 2, 3))
    assert len(made) == 1  # only rank 0 asks NCCL for an id
    assert exchange_comm_id(make, 0, 1) == b"\x07" * 128  # a world of one needs no channel


def test_reduce_film_is_a_noop_without_a_process_group():
    from pathtracer_rs_b200.dist import reduce_film

    t = torch.ones(4, 4, 4)
    assert reduce_film(t) is t and float(t.sum()) == 64.0
