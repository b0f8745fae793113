#!/usr/bin/env python
"""One render of a BASELINE config through the C ABI; prints stage times. For profiling / tuning.
usage: python tools/render_once.py [--scene 1] [--res 1024 1024] [--spp 16] [--depth 15] [--tris N] [--reps 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracer_rs_b200.gpu as gpu  # noqa: E402
import pathtracer_rs_b200.host as host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=host.SCENE_CORNELL_ENV)
    ap.add_argument("--res", type=int, nargs=2, default=[1024, 1024])
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--depth", type=int, default=15)
    ap.add_argument("--tris", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--batch", type=int, default=0, help="paths per wavefront batch (0 = library default)")
    ap.add_argument("--device-bvh", action="store_true", help="tree built by the library (ptrs_scene_create_device_bvh)")
    a = ap.parse_args()
    gpu.set_device(0)
    flat, cam = host.make_scene(a.scene, seed=1, n_tris=a.tris, res=tuple(a.res))
    scene = gpu.RenderScene(flat, device_bvh=a.device_bvh)
    if a.device_bvh:
        print("device tree: %d nodes, built in %.2f ms" % scene.bvh_info(), end="  ")
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(a.spp), max_depth=a.depth)
    integ.params.paths_per_batch = a.batch
    film = gpu.Film(cam.width, cam.height)
    for _ in range(a.reps):
        film.clear()
        st = integ.render(cam, scene, film)
    rays = st["extension_rays"] + st["shadow_rays"] + st["mis_rays"]
    print(f"paths {st['camera_paths']} rays {rays}  {st['camera_paths'] / st['ms_total'] / 1e3:.1f} Mpaths/s {rays / st['ms_total'] / 1e3:.0f} Mrays/s  "
          f"total {st['ms_total']:.1f} ms: gen {st['ms_generate']:.1f} extend {st['ms_extend']:.1f} shade {st['ms_shade']:.1f} "
          f"connect {st['ms_shadow']:.1f} acc {st['ms_accumulate']:.1f}  launches {st['launches']}")
    if a.out:
        from PIL import Image

        Image.fromarray(film.to_rgba_image()[..., :3]).save(a.out)


if __name__ == "__main__":
    main()
