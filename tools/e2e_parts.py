import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pathtracer_rs_b200.gpu as gpu, pathtracer_rs_b200.host as host
gpu.set_device(0)
flat, cam = host.make_scene(host.SCENE_CORNELL_ENV, seed=1, res=(1024, 1024))
integ = gpu.PathIntegrator(gpu.SamplerBuilder(16), max_depth=15)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc = gpu.RenderScene(flat); t1 = time.perf_counter()
    f = gpu.Film(1024, 1024); t2 = time.perf_counter()
    integ.render(cam, sc, f); t3 = time.perf_counter()
    img = f.download(); t4 = time.perf_counter()
    sc.close(); t5 = time.perf_counter()
    del f; t6 = time.perf_counter()
    torch.cuda.synchronize(); t7 = time.perf_counter()
    print("create %.2f film %.2f render %.2f download %.2f close %.2f delfilm %.2f sync %.2f" % tuple(1e3 * x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t7-t6)))
