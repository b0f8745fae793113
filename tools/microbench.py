#!/usr/bin/env python
"""BVH traversal microbenchmark (BASELINE configs[3]) for tuning and profiling.
usage: python tools/microbench.py [--tris N] [--side S] [--rays coherent|incoherent] [--any] [--iters K]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pathtracer_rs_b200.gpu as gpu  # noqa: E402
import pathtracer_rs_b200.host as host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--rays", default="incoherent")
    ap.add_argument("--any", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--scene", type=int, default=host.SCENE_TERRAIN)
    ap.add_argument("--all", action="store_true", help="run coherent/incoherent x closest/any in one process")
    ap.add_argument("--device-bvh", action="store_true", help="build the BVH on the GPU (ptrs_scene_create_device_bvh) instead of the host SAH build")
    a = ap.parse_args()
    gpu.set_device(0)
    flat, cam = host.make_scene(a.scene, seed=1, n_tris=a.tris, res=(a.side, a.side))
    scene = gpu.RenderScene(flat, device_bvh=a.device_bvh)
    n_nodes, build_ms = scene.bvh_info()
    print(f"tris={flat.n_prims} device nodes={n_nodes} host SAH build {flat.bvh_seconds * 1e3:.0f} ms" + (f", device build {build_ms:.2f} ms" if a.device_bvh else ""))
    if a.device_bvh:  # second build: allocations now come from the warmed pool
        scene.close()
        scene = gpu.RenderScene(flat, device_bvh=True)
        print(f"device build (warm) {scene.bvh_info()[1]:.2f} ms")
    bmin, bmax = flat.world_bound()
    if a.all:
        for kind in ("coherent", "incoherent"):
            for any_hit in (False, True):
                a.rays, a.any = kind, any_hit
                run(a, flat, cam, scene, bmin, bmax)
        return
    run(a, flat, cam, scene, bmin, bmax)


def run(a, flat, cam, scene, bmin, bmax):
    rays = host.coherent_rays(cam, a.side) if a.rays == "coherent" else host.incoherent_rays(bmin, bmax, 42, a.side * a.side)
    n = rays.shape[0]
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(n * (1 if a.any else 20), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    nodes, tris = scene.intersect_counted_device(d_rays.data_ptr(), n, d_out.data_ptr(), any_hit=a.any, stream=stream)
    alg = 32 * nodes + 36 * tris + n * (28 + (1 if a.any else 20))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if a.any:
            scene.intersect_p_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream)
        else:
            scene.intersect_device(d_rays.data_ptr(), n, d_out.data_ptr(), stream)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms)) * 1e-3
    print(f"{a.rays} {'any' if a.any else 'closest'} tris={flat.n_prims} rays={n}: {n / t / 1e6:.0f} Mrays/s  {t * 1e3:.2f} ms  "
          f"nodes/ray {nodes / n:.1f} tris/ray {tris / n:.2f}  alg {alg / t / 1e9:.0f} GB/s ({alg / t / 1e9 / 6536:.3f} of 6536)")


if __name__ == "__main__":
    main()
