#!/usr/bin/env python
"""bench.py — the reference's headline metric (camera samples/s and Mrays/s of the path integrator) on
BASELINE.json's configs[1]: Cornell box + environment map, 1024x1024, 64 spp, max_depth 15.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c2|c3|c5] [--spp S]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one full render of the workload (67.6 M camera paths) accumulated into a cleared film.
  value     paths/s with scene and film resident in HBM (CUDA events on the launching stream)
  e2e       the same through the C ABI with HOST buffers: ptrs_scene_create from host arrays, render,
            ptrs_film_download — host<->device copies inside the timed region
  roofline  extend kernel (closest-hit BVH traversal, the kernel BASELINE.json's "% of L2/HBM roofline" is about):
            algorithmic bytes (32 B x nodes tested + 36 B x triangles tested + 28 B ray + 20 B hit)
            / its summed launch time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json;
            roofline.stages adds the connect and shade stages the same way (shade is the largest share on C2);
            roofline.probes are the memory system's own figures measured in the same process — streaming reads and
            random 64-byte gathers from an L2-resident and from an HBM-sized buffer — and roofline.frac_of_l2_gather
            is the extend kernel against the one that matches its access pattern
  cpu_baseline  the C++ oracle (a restatement of the reference's rayon integrator; the Rust crate cannot
            be built in this image) on a bounded, strided sample of the same workload's 16x16 tiles
Multi-GPU (weak scaling): rank g renders Sobol sample numbers {s : s mod N == g} of a 64*N-spp render of
the same image, then one NCCL reduce of the film to rank 0 — the only collective on the path.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = {"name": "cornell+envmap 1024x1024 64spp depth15 (BASELINE configs[1])", "res": (1024, 1024), "spp": 64, "max_depth": 15,
            "scene": "SCENE_CORNELL_ENV", "n_tris": 0}
# The other BASELINE configs, selectable with --workload for the tables in DESIGN.md §6 (the default and
# the driver's line stay configs[1]).  --spp overrides the per-GPU sample count (C5's 1024 spp is run in
# full only when sharded; a reduced-spp run says so in config.workload).
WORKLOADS = {
    "c1": {"name": "cornell 512x512 16spp depth15 (BASELINE configs[0])", "res": (512, 512), "spp": 16, "max_depth": 15,
           "scene": "SCENE_CORNELL", "n_tris": 0},
    "c2": WORKLOAD,
    "c3": {"name": "1M-triangle material field (glass/substrate/metal/Disney/matte + env + area lights) 1920x1080 256spp depth15 (BASELINE configs[2])",
           "res": (1920, 1080), "spp": 256, "max_depth": 15, "scene": "SCENE_MATERIAL_FIELD", "n_tris": 1000000},
    "c5": {"name": "262k-triangle atrium 3840x2160 1024spp depth15 (BASELINE configs[4])", "res": (3840, 2160), "spp": 1024, "max_depth": 15,
           "scene": "SCENE_ATRIUM", "n_tris": 262144},
}
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def _run_nvml(self):
        """The same readings through NVML every 20 ms (an nvidia-smi process takes ~100 ms per sample, too coarse for a
        timed region of half a second).  False if NVML is not usable: the caller falls back to nvidia-smi."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.02)
        return True

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def host_threads():
    """Threads the CPU arm uses: every core this process may run on.  Passed explicitly because torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently make the reference arm single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_plan(oracle, flat, cam, params, target_s=12.0):
    """Pick a tile stride so that the oracle renders a uniform subset of the workload's 16x16 tiles in
    roughly target_s seconds on this host."""
    n_tiles = oracle.tile_count(cam, params)
    probe_stride = max(1, n_tiles // 64)
    t0 = time.perf_counter()
    _, st = oracle.render(flat, cam, params, n_threads=host_threads(), tile_stride=probe_stride)
    dt = time.perf_counter() - t0
    per_tile = dt / max(1, (n_tiles + probe_stride - 1) // probe_stride)
    want = max(16, int(target_s / max(per_tile, 1e-9)))
    return n_tiles, max(1, n_tiles // want)


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (C++ oracle, tile-parallel like
    integrator.rs:617-637, all host threads) on this arm's config; each step = a bounded tile sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pathtracer_rs_b200.host as host
    from oracle import oracle

    w = dict(WORKLOADS[args.workload])
    if args.spp:
        w["name"] += f" [run at {args.spp} spp per GPU]"
        w["spp"] = args.spp
    flat, cam = host.make_scene(getattr(host, w["scene"]), seed=1, n_tris=w["n_tris"], res=w["res"])
    params = host.default_render_params(spp=w["spp"], max_depth=w["max_depth"])
    n_tiles, stride = cpu_sample_plan(oracle, flat, cam, params, target_s=8.0)
    cores = host_threads()
    for _ in range(args.warmup):
        oracle.render(flat, cam, params, n_threads=cores, tile_stride=stride * 8)
    times, paths, rays = [], 0, 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        _, st = oracle.render(flat, cam, params, n_threads=cores, tile_stride=stride)
        times.append(time.perf_counter() - t0)
        paths = st["camera_paths"]
        rays = st["extension_rays"] + st["shadow_rays"] + st["mis_rays"]
    t = float(np.mean(times))
    value = paths / t
    sample = f"every {stride}th of the {n_tiles} 16x16 tiles, all {w['spp']} spp ({paths} camera paths per step)"
    line = {"impl": "reference", "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "mrays_per_s": rays / t / 1e6,
            "config": {"workload": w["name"], "resolution": list(w["res"]), "spp": w["spp"], "max_depth": w["max_depth"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = C++ restatement of the reference's rayon path integrator (oracle/); the Rust crate itself cannot be built in this image"}
    emit(line)


def bvh_microbench(gpu, host, torch, peak_gbs):
    """BASELINE configs[3]: fixed-ray intersection microbenchmark on a synthetic 10 M-triangle mesh,
    2^24 coherent and 2^24 incoherent rays; HBM-bound for incoherent rays."""
    n_tris = int(os.environ.get("PTRS_BENCH_BVH_TRIS", "10000000"))
    n_side = int(os.environ.get("PTRS_BENCH_BVH_SIDE", "4096"))
    flat, cam = host.make_scene(host.SCENE_TERRAIN, seed=1, n_tris=n_tris, res=(n_side, n_side))
    scene = gpu.RenderScene(flat)
    bmin, bmax = flat.world_bound()
    out = {"n_tris": flat.n_prims, "n_nodes": flat.n_nodes, "bvh_build_s": flat.bvh_seconds, "scene_mb": scene.device_bytes / 1e6}
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in ("coherent", "incoherent"):
        rays = host.coherent_rays(cam, n_side) if name == "coherent" else host.incoherent_rays(bmin, bmax, 42, n_side * n_side)
        n = rays.shape[0]
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
        d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
        d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")
        for any_hit in (False, True):
            d_out = d_occ if any_hit else d_hits
            nodes, tris = scene.intersect_counted_device(d_rays.data_ptr(), n, d_out.data_ptr(), any_hit=any_hit, stream=stream)
            alg_bytes = 32 * nodes + 36 * tris + n * (28 + (1 if any_hit else 20))
            ms = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if any_hit:
                    scene.intersect_p_device(d_rays.data_ptr(), n, d_occ.data_ptr(), stream)
                else:
                    scene.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
                e1.record()
                e1.synchronize()
                if it >= 3:
                    ms.append(e0.elapsed_time(e1))
            t = float(np.mean(ms)) * 1e-3
            key = f"{name}_{'any' if any_hit else 'closest'}"
            out[key] = {"mrays_per_s": n / t / 1e6, "ms": t * 1e3, "nodes_per_ray": nodes / n, "tris_per_ray": tris / n,
                        "achieved_gbs": alg_bytes / t / 1e9, "frac": alg_bytes / t / 1e9 / peak_gbs}
        del d_rays, d_hits, d_occ
    scene.close()
    # the same mesh with the BVH built on the device (ptrs_scene_create_device_bvh, csrc/k_bvh.cu)
    dev = {}
    for tag in ("build_ms_cold", "build_ms_warm"):
        scene = gpu.RenderScene(flat, device_bvh=True)
        dev["n_nodes"], dev[tag] = scene.bvh_info()
        if tag == "build_ms_cold":
            scene.close()
    dev["speedup_vs_host_sah"] = flat.bvh_seconds * 1e3 / dev["build_ms_warm"]
    rays = host.incoherent_rays(bmin, bmax, 42, n_side * n_side)
    n = rays.shape[0]
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.empty(n * 20, dtype=torch.uint8, device="cuda")
    nodes, tris = scene.intersect_counted_device(d_rays.data_ptr(), n, d_hits.data_ptr(), any_hit=False, stream=stream)
    ms = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scene.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
        e1.record()
        e1.synchronize()
        if it >= 3:
            ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms)) * 1e-3
    alg_bytes = 32 * nodes + 36 * tris + n * 48
    dev["incoherent_closest"] = {"mrays_per_s": n / t / 1e6, "ms": t * 1e3, "nodes_per_ray": nodes / n, "tris_per_ray": tris / n,
                                 "achieved_gbs": alg_bytes / t / 1e9, "frac": alg_bytes / t / 1e9 / peak_gbs}
    out["device_bvh"] = dev
    scene.close()
    try:  # the memory system's answer to this tree's access pattern: random 64-byte gathers over a buffer of the tree's size
        tree_bytes = int(out["n_nodes"]) * 32 + int(out["n_tris"]) * 48
        out["hbm_gather64_gbs_tree_sized"] = gpu.gather_bandwidth(tree_bytes, 256)
    except Exception:
        pass
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import pathtracer_rs_b200.gpu as gpu
    import pathtracer_rs_b200.host as host
    from pathtracer_rs_b200._abi import PtrsRenderParams
    from pathtracer_rs_b200.dist import reduce_film, sample_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the rendering hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    gpu.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    peak_gbs, peak_src = measured_peaks()

    w = dict(WORKLOADS[args.workload])
    if args.spp:
        w["name"] += f" [run at {args.spp} spp per GPU]"
        w["spp"] = args.spp
    W, H = w["res"]
    flat, cam = host.make_scene(getattr(host, w["scene"]), seed=1, n_tris=w["n_tris"], res=w["res"])
    scene = gpu.RenderScene(flat)
    # weak scaling: the image is rendered at spp * N with rank g taking sample numbers s = g (mod N)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(w["spp"] * n_gpus), max_depth=w["max_depth"])
    integ.preprocess(scene)
    film_t = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    film = gpu.Film(W, H, device_ptr=film_t.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        film_t.zero_()
        st = integ.render(cam, scene, film, stream=stream, sample_stride=sample_shard(rank, n_gpus))
        reduce_film(film_t, dst=0)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one untimed pass with visit counters on: algorithmic bytes of the (deterministic) step
    scene.set_stats_mode(True)
    st_count = step()
    scene.set_stats_mode(False)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    step_ms, ext_ms, stats = [], [], None
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stats = step()
        e1.record()
        barrier()
        step_ms.append(e0.elapsed_time(e1))
        ext_ms.append(stats["ms_extend"])
    clocks = sampler.summary() if sampler else None
    t_local = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)  # max over ranks
    total_ms = float(t_local.item())
    ms_per_step = total_ms / args.steps
    paths_step = stats["camera_paths"] * n_gpus
    rays_step = (stats["extension_rays"] + stats["shadow_rays"] + stats["mis_rays"]) * n_gpus
    value = paths_step / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers -----------------------------------------------
    host_film = np.empty((H, W, 4), dtype=np.float32)
    e2e_ms = []
    for it in range(max(2, min(args.steps, 3)) + 1):
        barrier()
        t0 = time.perf_counter()
        sc2 = gpu.RenderScene(flat)  # ptrs_scene_create: H2D of the whole flattened scene
        f2 = gpu.Film(W, H)
        t1 = time.perf_counter()
        integ.render(cam, sc2, f2, sample_stride=sample_shard(rank, n_gpus))
        t2 = time.perf_counter()
        if world > 1:
            reduce_film(torch.as_tensor(_CudaArray(f2.device_ptr, (H, W, 4)), device="cuda"), dst=0)
            torch.cuda.synchronize()
        if rank == 0:
            gpu._check(gpu.lib().ptrs_film_download(f2._h, host_film.ctypes.data_as(C.POINTER(C.c_float))))
        t3 = time.perf_counter()
        sc2.close()
        del f2
        barrier()
        if it > 0:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
            e2e_parts = {"scene_create_ms": (t1 - t0) * 1e3, "render_ms": (t2 - t1) * 1e3, "film_reduce_download_ms": (t3 - t2) * 1e3,
                         "destroy_ms": (time.perf_counter() - t3) * 1e3}
    e2e_local = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_local, op=dist.ReduceOp.MAX)
    e2e_value = paths_step / (float(e2e_local.item()) * 1e-3)
    h2d = int(flat.host_bytes) + C.sizeof(type(cam)) + C.sizeof(PtrsRenderParams)
    d2h = H * W * 16

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (extend) ---------------------------------------------------
    ext_rays = st_count["extension_rays"]
    alg_bytes = 32 * st_count["nodes_tested"] + 36 * st_count["tris_tested"] + ext_rays * (28 + 20)
    ext_s = float(np.mean(ext_ms)) * 1e-3
    achieved = alg_bytes / ext_s / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "extend_dram_bytes.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "extend_kernel<false> (closest-hit BVH traversal)", "achieved": achieved, "peak": peak_gbs,
                "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_step": alg_bytes, "launches_per_step": stats["extend_launches"],
                "avg_launch_ms": float(np.mean(ext_ms)) / max(1, stats["extend_launches"]),
                "bytes_per_ray": alg_bytes / ext_rays, "nodes_per_ray": st_count["nodes_tested"] / ext_rays,
                "tris_per_ray": st_count["tris_tested"] / ext_rays, "share_of_step": float(np.mean(ext_ms)) / (sum(step_ms) / len(step_ms)),
                "note": f"scene ({flat.n_prims} triangles, {flat.n_nodes} nodes) is cache resident (L1/L2): the HBM-bound case is bvh_microbench"}

    # the other two stages next to it, so the line shows where the rest of the step goes: the connect stage by the
    # same traversal byte count (ms_shadow also contains connect_resolve), the shade stage by its record traffic
    # (84 B in: queue entry, hit, 64 B path slot; up to 176 B out: slot, 96 B direct-lighting record, queue entries)
    step_mean = sum(step_ms) / len(step_ms)
    nee_rays = st_count["shadow_rays"] + st_count["mis_rays"]
    nee_bytes = 32 * st_count["nee_nodes_tested"] + 36 * st_count["nee_tris_tested"] + 32 * nee_rays + 1 * st_count["shadow_rays"] + 16 * st_count["mis_rays"]
    shade_bytes = (84 + 176) * ext_rays
    stages = {
        "extend": {"ms": float(np.mean(ext_ms)), "share_of_step": float(np.mean(ext_ms)) / step_mean, "achieved_gbs": achieved, "frac": achieved / peak_gbs},
        "connect+resolve": {"ms": stats["ms_shadow"], "share_of_step": stats["ms_shadow"] / step_mean,
                            "achieved_gbs": nee_bytes / (stats["ms_shadow"] * 1e-3) / 1e9, "frac": nee_bytes / (stats["ms_shadow"] * 1e-3) / 1e9 / peak_gbs,
                            "nodes_per_ray": st_count["nee_nodes_tested"] / max(1, nee_rays)},
        "shade (all materials + miss)": {"ms": stats["ms_shade"], "share_of_step": stats["ms_shade"] / step_mean,
                                         "achieved_gbs": shade_bytes / (stats["ms_shade"] * 1e-3) / 1e9, "frac": shade_bytes / (stats["ms_shade"] * 1e-3) / 1e9 / peak_gbs,
                                         "note": "latency bound (dependent scene / table loads at 16 warps per SM), not a bandwidth kernel; upper-bound bytes"},
    }
    roofline["stages"] = stages
    # What the memory system gives THIS access pattern, measured here and now (MEASURED_PEAKS.json has a copy
    # bandwidth only): streaming 256-bit reads and independent random 64-byte gathers (one sibling pair of nodes),
    # from an L2-resident buffer and from one far larger than L2 (ptrs_read_bandwidth / ptrs_gather_bandwidth).
    try:
        probes = {"l2_stream_read_gbs_64MiB": gpu.read_bandwidth(64 << 20, 30), "l2_gather64_gbs_64MiB": gpu.gather_bandwidth(64 << 20, 512),
                  "hbm_stream_read_gbs_4GiB": gpu.read_bandwidth(4 << 30, 2), "hbm_gather64_gbs_1GiB": gpu.gather_bandwidth(1 << 30, 256)}
        roofline["probes"] = probes
        roofline["frac_of_l2_gather"] = achieved / probes["l2_gather64_gbs_64MiB"]
    except Exception as e:  # diagnostic only
        roofline["probes"] = {"error": str(e)}

    # ---- CPU baseline: the oracle on a bounded sample of the same workload ----------------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        from oracle import oracle

        p1 = host.default_render_params(spp=w["spp"], max_depth=w["max_depth"])
        n_tiles, stride = cpu_sample_plan(oracle, flat, cam, p1, target_s=12.0)
        t0 = time.perf_counter()
        _, ost = oracle.render(flat, cam, p1, n_threads=host_threads(), tile_stride=stride)
        dt = time.perf_counter() - t0
        cpu = {"value": ost["camera_paths"] / dt, "unit": "samples/s", "cores": host_threads(), "kind": "port",
               "sample": f"every {stride}th of the {n_tiles} 16x16 tiles, all {w['spp']} spp ({ost['camera_paths']} camera paths, {dt:.1f} s)",
               "mrays_per_s": (ost["extension_rays"] + ost["shadow_rays"] + ost["mis_rays"]) / dt / 1e6}

    micro = None
    if n_gpus == 1 and not args.no_bvh_microbench:
        del film, film_t
        scene.close()
        torch.cuda.empty_cache()
        micro = bvh_microbench(gpu, host, torch, peak_gbs)

    line = {"metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mrays_per_s": rays_step / (ms_per_step * 1e-3) / 1e6,
            "config": {"workload": w["name"], "resolution": [W, H], "spp_per_gpu": w["spp"], "max_depth": w["max_depth"],
                       "camera_paths_per_step": paths_step, "rays_per_step": rays_step, "sharding": f"sample index mod {n_gpus}, film reduced with NCCL",
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(e2e_local.item()), "includes": "ptrs_scene_create from host arrays + render + ptrs_film_download + ptrs_scene_destroy",
                    "parts_last_step": e2e_parts},
            "gpu_launches": int(stats["launches"]) * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "stage_ms": {k: stats[k] for k in ("ms_generate", "ms_extend", "ms_shade", "ms_shadow", "ms_accumulate", "ms_total")},
            "rays": {k: stats[k] for k in ("extension_rays", "shadow_rays", "mis_rays")}, "bvh_microbench": micro}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap a device pointer owned by the library."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner) write to
    file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
    private duplicate of the original stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


_OUT = None


def emit(line):
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bvh-microbench", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
