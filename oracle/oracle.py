"""ctypes face of the CPU oracle (oracle/_build/liboracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from pathtracer_rs_b200._abi import PtrsCamera, PtrsHit, PtrsLobeDesc, PtrsRay, PtrsRenderParams, PtrsSceneDesc  # noqa: E402
from pathtracer_rs_b200.host import HIT_DTYPE, RAY_DTYPE  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle`")
        L = C.CDLL(path)
        i32p, u64p, fp = C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
        camp, rpp, descp = C.POINTER(PtrsCamera), C.POINTER(PtrsRenderParams), C.POINTER(PtrsSceneDesc)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_init.argtypes = [C.c_char_p]
        L.oracle_sobol_interval_to_index.restype = C.c_uint64
        L.oracle_sobol_interval_to_index.argtypes = [C.c_uint32, C.c_uint64, C.c_int32, C.c_int32]
        L.oracle_sobol_sample.restype = C.c_float
        L.oracle_sobol_sample.argtypes = [C.c_int64, C.c_uint32, C.c_uint64]
        L.oracle_sobol_samples.argtypes = [camp, rpp, i32p, i32p, C.c_size_t, i32p, C.c_size_t, fp, u64p]
        L.oracle_generate_rays.argtypes = [camp, rpp, i32p, i32p, C.c_size_t, C.POINTER(PtrsRay), fp, fp]
        L.oracle_intersect.argtypes = [descp, C.POINTER(PtrsRay), C.c_size_t, C.POINTER(PtrsHit), u64p, C.c_int]
        L.oracle_intersect_p.argtypes = [descp, C.POINTER(PtrsRay), C.c_size_t, C.POINTER(C.c_uint8), u64p, C.c_int]
        L.oracle_path_radiance.argtypes = [descp, camp, rpp, i32p, i32p, C.c_size_t, fp, C.c_int]
        L.oracle_render.argtypes = [descp, camp, rpp, fp, C.c_int, C.c_int64, C.c_int64, C.c_int64, u64p]
        L.oracle_tile_count.restype = C.c_int64
        L.oracle_tile_count.argtypes = [camp, rpp]
        L.oracle_film_resolve.argtypes = [fp, C.c_int, C.c_int, fp]
        L.oracle_gamma.restype = C.c_float
        L.oracle_gamma.argtypes = [C.c_uint32]
        L.oracle_log2_int.restype = C.c_uint32
        L.oracle_log2_int.argtypes = [C.c_uint64]
        L.oracle_solve_2x2.argtypes = [fp, fp, fp]
        L.oracle_next_float_up.restype = C.c_float
        L.oracle_next_float_up.argtypes = [C.c_float]
        L.oracle_next_float_down.restype = C.c_float
        L.oracle_next_float_down.argtypes = [C.c_float]
        L.oracle_cantor_pairing.restype = C.c_uint64
        L.oracle_cantor_pairing.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_fr_dielectric.restype = C.c_float
        L.oracle_fr_dielectric.argtypes = [C.c_float] * 3
        L.oracle_cosine_sample_hemisphere.argtypes = [C.c_float, C.c_float, fp]
        L.oracle_bxdf_eval.argtypes = [C.c_int, fp, fp, fp, fp]
        L.oracle_bxdf_sample.argtypes = [C.c_int, fp, fp, C.c_float, C.c_float, fp]
        lobep = C.POINTER(PtrsLobeDesc)
        L.oracle_lobe_eval.argtypes = [lobep, fp, fp, C.c_size_t, fp]
        L.oracle_lobe_sample.argtypes = [lobep, fp, fp, C.c_size_t, fp]
        L.oracle_light_sample.argtypes = [descp, C.c_int, fp, fp, fp, C.c_size_t, fp]
        L.oracle_light_pdf.argtypes = [descp, C.c_int, fp, fp, fp, C.c_size_t, fp]
        tables = os.path.join(_ROOT, "pathtracer_rs_b200", "data", "sobol_tables.bin")
        if L.oracle_init(tables.encode()) != 0:
            raise RuntimeError(L.oracle_last_error().decode())
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _pix(pixels, samples):
    px = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1, 2)
    sm = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1)
    assert px.shape[0] == sm.shape[0]
    return px, sm


def sobol_samples(cam, params, pixels, samples, dims):
    px, sm = _pix(pixels, samples)
    dm = np.ascontiguousarray(dims, dtype=np.int32)
    out = np.empty((px.shape[0], dm.shape[0]), dtype=np.float32)
    idx = np.empty(px.shape[0], dtype=np.uint64)
    lib().oracle_sobol_samples(C.byref(cam), C.byref(params), _p(px, C.c_int32), _p(sm, C.c_int32), px.shape[0],
                               _p(dm, C.c_int32), dm.shape[0], _p(out, C.c_float), _p(idx, C.c_uint64))
    return out, idx


def generate_rays(cam, params, pixels, samples):
    px, sm = _pix(pixels, samples)
    n = px.shape[0]
    rays = np.empty(n, dtype=RAY_DTYPE)
    pf = np.empty((n, 2), dtype=np.float32)
    rxry = np.empty((n, 6), dtype=np.float32)
    lib().oracle_generate_rays(C.byref(cam), C.byref(params), _p(px, C.c_int32), _p(sm, C.c_int32), n,
                               _p(rays, PtrsRay), _p(pf, C.c_float), _p(rxry, C.c_float))
    return rays, pf, rxry


def intersect(scene, rays, n_threads=0):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
    ctr = np.zeros(2, dtype=np.uint64)
    lib().oracle_intersect(scene.desc, _p(rays, PtrsRay), rays.shape[0], _p(hits, PtrsHit), _p(ctr, C.c_uint64), n_threads)
    return hits, (int(ctr[0]), int(ctr[1]))


def intersect_p(scene, rays, n_threads=0):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    occ = np.empty(rays.shape[0], dtype=np.uint8)
    ctr = np.zeros(2, dtype=np.uint64)
    lib().oracle_intersect_p(scene.desc, _p(rays, PtrsRay), rays.shape[0], _p(occ, C.c_uint8), _p(ctr, C.c_uint64), n_threads)
    return occ, (int(ctr[0]), int(ctr[1]))


def path_radiance(scene, cam, params, pixels, samples, n_threads=0):
    px, sm = _pix(pixels, samples)
    out = np.empty((px.shape[0], 3), dtype=np.float32)
    rc = lib().oracle_path_radiance(scene.desc, C.byref(cam), C.byref(params), _p(px, C.c_int32), _p(sm, C.c_int32),
                                    px.shape[0], _p(out, C.c_float), n_threads)
    if rc != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return out


def lobe_eval(lobe, wo, wi):
    """BxDF::f / pdf, one row per (wo, wi): f rgb, pdf (mirrors ptrs_bxdf_eval)."""
    o, w = np.ascontiguousarray(wo, dtype=np.float32), np.ascontiguousarray(wi, dtype=np.float32)
    out = np.empty((o.shape[0], 4), dtype=np.float32)
    lib().oracle_lobe_eval(C.byref(lobe), _p(o, C.c_float), _p(w, C.c_float), o.shape[0], _p(out, C.c_float))
    return out


def lobe_sample(lobe, wo, u):
    """BxDF::sample_f: wi xyz, f rgb, pdf, sampled type (mirrors ptrs_bxdf_sample)."""
    o, uu = np.ascontiguousarray(wo, dtype=np.float32), np.ascontiguousarray(u, dtype=np.float32)
    out = np.empty((o.shape[0], 8), dtype=np.float32)
    lib().oracle_lobe_sample(C.byref(lobe), _p(o, C.c_float), _p(uu, C.c_float), o.shape[0], _p(out, C.c_float))
    return out


def light_sample(scene, light, ref_p, ref_n, u):
    p, nn, uu = (np.ascontiguousarray(a, dtype=np.float32) for a in (ref_p, ref_n, u))
    out = np.empty((p.shape[0], 16), dtype=np.float32)
    lib().oracle_light_sample(scene.desc, light, _p(p, C.c_float), _p(nn, C.c_float), _p(uu, C.c_float), p.shape[0], _p(out, C.c_float))
    return out


def light_pdf(scene, light, ref_p, ref_n, wi):
    p, nn, w = (np.ascontiguousarray(a, dtype=np.float32) for a in (ref_p, ref_n, wi))
    out = np.empty(p.shape[0], dtype=np.float32)
    lib().oracle_light_pdf(scene.desc, light, _p(p, C.c_float), _p(nn, C.c_float), _p(w, C.c_float), p.shape[0], _p(out, C.c_float))
    return out


def tile_count(cam, params):
    return lib().oracle_tile_count(C.byref(cam), C.byref(params))


def render(scene, cam, params, n_threads=0, tile_begin=0, tile_end=0, tile_stride=1, film=None):
    """Returns (film_rgbw (H, W, 4) raw sums, stats dict)."""
    if film is None:
        film = np.zeros((cam.height, cam.width, 4), dtype=np.float32)
    st = np.zeros(6, dtype=np.uint64)
    rc = lib().oracle_render(scene.desc, C.byref(cam), C.byref(params), _p(film, C.c_float), n_threads, tile_begin, tile_end,
                             tile_stride, _p(st, C.c_uint64))
    if rc != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    keys = ["camera_paths", "extension_rays", "shadow_rays", "mis_rays", "nodes_tested", "tris_tested"]
    return film, dict(zip(keys, (int(x) for x in st)))


def resolve(film_rgbw):
    h, w, _ = film_rgbw.shape
    out = np.empty((h, w, 3), dtype=np.float32)
    f = np.ascontiguousarray(film_rgbw, dtype=np.float32)
    lib().oracle_film_resolve(_p(f, C.c_float), w, h, _p(out, C.c_float))
    return out
