#!/usr/bin/env python
"""Regenerates the golden vectors in this directory from the CPU oracle.

The reference is a Rust crate that cannot be built or imported in this image (no rustc/cargo), so these
are NOT outputs of the reference binary: they freeze the oracle's restatement (oracle/*.hpp) so that later
edits to the oracle or to the host-side scene code cannot drift unnoticed, and they give the GPU tests
fixtures that do not depend on the oracle being built.   Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import pathtracer_rs_b200.host as host  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    # Sobol: two of BASELINE's resolutions
    out = {}
    for tag, res, spp in (("c1", (512, 512), 16), ("c5", (3840, 2160), 1024)):
        cam = host.look_at_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, *res)
        params = host.default_render_params(spp=spp)
        px = np.stack([rng.integers(-2, res[0] + 2, 512), rng.integers(-2, res[1] + 2, 512)], 1).astype(np.int32)
        sm = rng.integers(0, spp, 512).astype(np.int32)
        dims = np.arange(0, 48, dtype=np.int32)
        v, idx = oracle.sobol_samples(cam, params, px, sm, dims)
        out.update({f"{tag}_px": px, f"{tag}_sm": sm, f"{tag}_bits": v.view(np.uint32), f"{tag}_index": idx})
    np.savez_compressed(os.path.join(HERE, "sobol.npz"), **out)

    flat, cam = host.make_scene(host.SCENE_CORNELL, res=(32, 32))
    bmin, bmax = flat.world_bound()
    rays = np.concatenate([host.coherent_rays(cam, 48), host.incoherent_rays(bmin, bmax, 42, 3000)])
    hits, ctr = oracle.intersect(flat, rays)
    occ, ctr_p = oracle.intersect_p(flat, rays)
    np.savez_compressed(os.path.join(HERE, "cornell_hits.npz"), rays=rays.view(np.uint8), hits=hits.view(np.uint8), occluded=occ,
                        counters=np.array([*ctr, *ctr_p], dtype=np.uint64), nodes=flat.nodes().view(np.uint8))

    params = host.default_render_params(spp=8, max_depth=15)
    px = np.stack([rng.integers(-2, 34, 512), rng.integers(-2, 34, 512)], 1).astype(np.int32)
    sm = rng.integers(0, 8, 512).astype(np.int32)
    rad = oracle.path_radiance(flat, cam, params, px, sm)
    film, st = oracle.render(flat, cam, params, n_threads=1)
    np.savez_compressed(os.path.join(HERE, "cornell_render.npz"), px=px, sm=sm, radiance=rad, film=film,
                        stats=np.array([st[k] for k in ("camera_paths", "extension_rays", "shadow_rays", "mis_rays")], dtype=np.uint64))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
