#!/usr/bin/env python
"""Device BVH build timing: python tools/bvh_build_time.py [--tris N] [--reps K]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracer_rs_b200.gpu as gpu  # noqa: E402
import pathtracer_rs_b200.host as host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--scene", type=int, default=host.SCENE_TERRAIN)
a = ap.parse_args()
gpu.set_device(0)
flat, cam = host.make_scene(a.scene, seed=1, n_tris=a.tris, res=(256, 256))
print(f"tris={flat.n_prims} host SAH build {flat.bvh_seconds * 1e3:.0f} ms, host-tree depth {flat.bvh_depth}")
for _ in range(a.reps):
    s = gpu.RenderScene(flat, device_bvh=True)
    print("device build: nodes %d, %.2f ms" % s.bvh_info())
    s.close()
