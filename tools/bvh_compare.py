import numpy as np, sys
sys.path.insert(0, "/root/repo")
import pathtracer_rs_b200.gpu as gpu, pathtracer_rs_b200.host as host
gpu.set_device(0)
for kind, nt, res in ((host.SCENE_CORNELL, 0, (96,96)), (host.SCENE_MATERIAL_FIELD, 60000, (96,64)), (host.SCENE_TERRAIN, 200000, (128,128)), (host.SCENE_ATRIUM, 40000, (96,54))):
    flat, cam = host.make_scene(kind, seed=1, n_tris=nt, res=res)
    ref, dev = gpu.RenderScene(flat), gpu.RenderScene(flat, device_bvh=True)
    bmin, bmax = flat.world_bound()
    for name, rays in (("coh", host.coherent_rays(cam, 192)), ("inc", host.incoherent_rays(bmin, bmax, 7, 60000))):
        a, b = ref.intersect(rays), dev.intersect(rays)
        hit = (a["prim"] >= 0) & (b["prim"] >= 0)
        dt = a["t"][hit] != b["t"][hit]
        rel = np.abs(a["t"][hit][dt] - b["t"][hit][dt]) / a["t"][hit][dt]
        print(kind, name, "hitmask equal", np.array_equal(a["prim"] >= 0, b["prim"] >= 0), "t differs", dt.sum(), "of", hit.sum(), "max rel", rel.max() if rel.size else 0,
              "prim differs", (a["prim"] != b["prim"]).sum())
