"""Host-side scene preparation (libptrs_host.so): what the reference's Rust host does before
`PathIntegrator::render` — build meshes, materials, lights, the SAH BVH — ending in a PtrsSceneDesc.
No GPU code here."""
import ctypes as C
import os

import numpy as np

from ._abi import *  # noqa: F401,F403
from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCENE_CORNELL, SCENE_CORNELL_ENV, SCENE_MATERIAL_FIELD, SCENE_TERRAIN, SCENE_ATRIUM = range(5)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "lib", "libptrs_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        vp, i32, u32, u64, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
        fp = C.POINTER(C.c_float)
        L.ptrs_host_last_error.restype = C.c_char_p
        L.ptrs_host_builder_new.restype = vp
        L.ptrs_host_builder_free.argtypes = [vp]
        L.ptrs_host_add_constant_texture.argtypes = [vp, i32, f32, f32, f32]
        L.ptrs_host_add_checker_texture.argtypes = [vp, i32, fp, fp, f32, f32, f32, f32]
        L.ptrs_host_add_image_texture.argtypes = [vp, i32, fp, i32, i32, i32, f32, f32, f32, f32]
        L.ptrs_host_add_material.argtypes = [vp, C.POINTER(PtrsMaterial)]
        L.ptrs_host_set_normal_map.argtypes = [vp, i32, i32]
        L.ptrs_host_add_mesh.argtypes = [vp, fp, u32, fp, fp, fp, C.POINTER(u32), u32, fp, i32, i32, i32]
        L.ptrs_host_add_shape.argtypes = [vp, i32, fp, i32, i32]
        L.ptrs_host_add_point_light.argtypes = [vp, fp, fp]
        L.ptrs_host_add_directional_light.argtypes = [vp, fp, fp, fp]
        L.ptrs_host_add_infinite_light.argtypes = [vp, fp, fp, i32, i32]
        L.ptrs_host_mitsuba_env_light_to_world.argtypes = [fp]
        L.ptrs_host_finalize.restype = vp
        L.ptrs_host_finalize.argtypes = [vp, i32, i32]
        L.ptrs_host_scene_free.argtypes = [vp]
        L.ptrs_host_scene_desc.restype = C.POINTER(PtrsSceneDesc)
        L.ptrs_host_scene_desc.argtypes = [vp]
        L.ptrs_host_scene_bytes.restype = u64
        L.ptrs_host_scene_bytes.argtypes = [vp]
        L.ptrs_host_scene_desc_device_tables.restype = C.POINTER(PtrsSceneDesc)
        L.ptrs_host_scene_desc_device_tables.argtypes = [vp]
        L.ptrs_host_scene_bytes_device_tables.restype = u64
        L.ptrs_host_scene_bytes_device_tables.argtypes = [vp]
        L.ptrs_host_scene_bvh_depth.argtypes = [vp]
        L.ptrs_host_scene_bvh_seconds.restype = C.c_double
        L.ptrs_host_scene_bvh_seconds.argtypes = [vp]
        L.ptrs_host_make_scene.restype = vp
        L.ptrs_host_make_scene.argtypes = [i32, u64, u64, i32, i32, C.POINTER(PtrsCamera), i32]
        L.ptrs_host_make_scene_env.restype = vp
        L.ptrs_host_make_scene_env.argtypes = [i32, u64, u64, i32, i32, C.POINTER(PtrsCamera), i32, C.c_char_p]
        L.ptrs_host_synth_sky.argtypes = [i32, i32, u64, fp]
        L.ptrs_host_make_camera.argtypes = [fp, fp, f32, f32, f32, f32, i32, i32, C.POINTER(PtrsCamera)]
        L.ptrs_host_mitsuba_camera.argtypes = [fp, f32, i32, i32, i32, i32, C.POINTER(PtrsCamera)]
        L.ptrs_host_look_at_camera.argtypes = [fp, fp, fp, f32, i32, i32, C.POINTER(PtrsCamera)]
        L.ptrs_host_default_render_params.argtypes = [C.POINTER(PtrsRenderParams)]
        L.ptrs_host_gaussian_filter_table.argtypes = [f32, f32, fp]
        L.ptrs_host_coherent_rays.argtypes = [C.POINTER(PtrsCamera), i32, C.POINTER(PtrsRay)]
        L.ptrs_host_incoherent_rays.argtypes = [fp, fp, u64, u64, C.POINTER(PtrsRay)]
        L.ptrs_host_import_scene.restype = vp
        L.ptrs_host_import_scene.argtypes = [C.c_char_p, i32, i32, i32, C.c_char_p, C.POINTER(PtrsCamera), i32]
        ip = C.POINTER(C.c_int)
        L.ptrs_host_load_hdr.argtypes = [C.c_char_p, ip, ip, fp]
        L.ptrs_host_save_hdr.argtypes = [C.c_char_p, fp, i32, i32]
        L.ptrs_host_load_png.argtypes = [C.c_char_p, ip, ip, ip, C.POINTER(C.c_uint8)]
        L.ptrs_host_save_png.argtypes = [C.c_char_p, C.POINTER(C.c_uint8), i32, i32, i32]
        L.ptrs_host_decode_image.argtypes = [C.POINTER(C.c_uint8), C.c_size_t, ip, ip, ip, C.POINTER(C.c_uint8)]
        L.ptrs_host_snake_case.argtypes = [C.c_char_p, C.c_char_p, i32]
        u8p = C.POINTER(C.c_uint8)
        L.ptrs_host_tev_create_image.restype = C.c_int64
        L.ptrs_host_tev_create_image.argtypes = [i32, i32, C.c_char_p, u8p, C.c_int64]
        L.ptrs_host_tev_update_image.restype = C.c_int64
        L.ptrs_host_tev_update_image.argtypes = [fp, i32, i32, C.c_char_p, u8p, C.c_int64]
        L.ptrs_host_build_bvh.argtypes = [fp, u32, i32, i32, C.POINTER(PtrsBvhNode), u32, C.POINTER(u32), C.POINTER(u32)]
        _LIB = L
    return _LIB


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("t_max", np.float32)])
HIT_DTYPE = np.dtype([("prim", np.int32), ("t", np.float32), ("b0", np.float32), ("b1", np.float32), ("b2", np.float32)])
NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("bmax", np.float32, 3), ("offset", np.uint32),
                       ("n_prims", np.uint16), ("axis", np.uint8), ("pad", np.uint8)])
assert RAY_DTYPE.itemsize == 28 and HIT_DTYPE.itemsize == 20 and NODE_DTYPE.itemsize == 32


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class FlatScene:
    """Owns a finalized host scene; `.desc` is the PtrsSceneDesc* to hand to ptrs_scene_create."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("host scene build failed: " + lib().ptrs_host_last_error().decode(errors="replace"))
        self._h = handle
        self.desc = lib().ptrs_host_scene_desc(handle)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().ptrs_host_scene_free(self._h)
            except TypeError:  # interpreter shutdown: module globals are already gone
                pass
            self._h = None

    @property
    def n_prims(self):
        return self.desc.contents.n_prims

    @property
    def n_nodes(self):
        return self.desc.contents.n_nodes

    @property
    def n_lights(self):
        return self.desc.contents.n_lights

    @property
    def host_bytes(self):
        return lib().ptrs_host_scene_bytes(self._h)

    @property
    def desc_device_tables(self):
        """The same scene without the tables the device library builds itself: level-0-only MIP pyramids and env lights
        without Distribution2D arrays (ptrs_scene_create completes them on the device)."""
        return lib().ptrs_host_scene_desc_device_tables(self._h)

    @property
    def host_bytes_device_tables(self):
        return lib().ptrs_host_scene_bytes_device_tables(self._h)

    @property
    def bvh_depth(self):
        return lib().ptrs_host_scene_bvh_depth(self._h)

    @property
    def bvh_seconds(self):
        return lib().ptrs_host_scene_bvh_seconds(self._h)

    def nodes(self):
        d = self.desc.contents
        return np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_uint8)), shape=(d.n_nodes * 32,)).view(NODE_DTYPE)

    def world_bound(self):
        n = self.nodes()[0]
        return np.array(n["bmin"]), np.array(n["bmax"])

    def prim_vertices(self):
        d = self.desc.contents
        idx = np.ctypeslib.as_array(d.prim_vertex, shape=(d.n_prims, 3))
        pos = np.ctypeslib.as_array(d.pos, shape=(d.n_verts, 3))
        return pos[idx]  # (n_prims, 3, 3)


# The reference's environment map (data/abandoned_tank_farm_04_1k.hdr, what `<emitter type="sunsky"/>` loads:
# pathtracer/importer/mitsuba.rs:400-418), kept as a test / bench fixture.
TANK_FARM_HDR = os.path.join(os.path.dirname(_HERE), "tests", "golden", "abandoned_tank_farm_04_1k.hdr")


def make_scene(kind, seed=1, n_tris=0, res=(512, 512), n_threads=0, env_hdr=None):
    """Ready-made scene + camera. kind: SCENE_* constant.  env_hdr (SCENE_CORNELL_ENV only): Radiance .hdr file to
    use as the environment map instead of the synthetic sky."""
    cam = PtrsCamera()
    h = lib().ptrs_host_make_scene_env(kind, seed, n_tris, res[0], res[1], C.byref(cam), n_threads, os.fsencode(env_hdr) if env_hdr else None)
    return FlatScene(h), cam


def import_scene(path, res=(640, 480), default_lights=False, sunsky_hdr=None, n_threads=0):
    """common::importer::import (src/common/importer/mod.rs:6-25): Mitsuba .xml or glTF .gltf/.glb -> (FlatScene, camera).
    `res` is the CLI's -r WxH (default 640x480, common/mod.rs:14)."""
    cam = PtrsCamera()
    h = lib().ptrs_host_import_scene(os.fsencode(path), res[0], res[1], 1 if default_lights else 0,
                                     os.fsencode(sunsky_hdr) if sunsky_hdr else None, C.byref(cam), n_threads)
    return FlatScene(h), cam


def _img_check(rc):
    if rc != 0:
        raise RuntimeError(lib().ptrs_host_last_error().decode(errors="replace"))


def load_hdr(path):
    """Radiance RGBE -> (h, w, 3) float32, as image::hdr::HdrDecoder::read_image_hdr gives it to light.rs:331-346."""
    w, h = C.c_int(0), C.c_int(0)
    _img_check(lib().ptrs_host_load_hdr(os.fsencode(path), C.byref(w), C.byref(h), None))
    out = np.empty((h.value, w.value, 3), dtype=np.float32)
    _img_check(lib().ptrs_host_load_hdr(os.fsencode(path), C.byref(w), C.byref(h), _fp(out)))
    return out


def save_hdr(path, rgb):
    img = _f32(rgb)
    _img_check(lib().ptrs_host_save_hdr(os.fsencode(path), _fp(img), img.shape[1], img.shape[0]))


def load_png(path):
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    _img_check(lib().ptrs_host_load_png(os.fsencode(path), C.byref(w), C.byref(h), C.byref(c), None))
    out = np.empty((h.value, w.value, c.value), dtype=np.uint8)
    _img_check(lib().ptrs_host_load_png(os.fsencode(path), C.byref(w), C.byref(h), C.byref(c), out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out


def decode_image(data):
    """PNG or JPEG bytes -> (h, w, c) uint8, what image::open gives the importers."""
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    _img_check(lib().ptrs_host_decode_image(buf, len(data), C.byref(w), C.byref(h), C.byref(c), None))
    out = np.empty((h.value, w.value, c.value), dtype=np.uint8)
    _img_check(lib().ptrs_host_decode_image(buf, len(data), C.byref(w), C.byref(h), C.byref(c), out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out


def save_png(path, pixels):
    """film.to_rgba_image().save(path) (src/headless.rs:222, 231): (h, w, 1..4) uint8 -> PNG."""
    img = np.ascontiguousarray(pixels, dtype=np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    _img_check(lib().ptrs_host_save_png(os.fsencode(path), img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[1], img.shape[0], img.shape[2]))


def tev_create_image(width, height, name="render"):
    """TevControlCreateImage::new_message (src/headless.rs:84-100) as bytes."""
    n = lib().ptrs_host_tev_create_image(width, height, name.encode(), None, 0)
    buf = np.empty(n, dtype=np.uint8)
    lib().ptrs_host_tev_create_image(width, height, name.encode(), buf.ctypes.data_as(C.POINTER(C.c_uint8)), n)
    return buf.tobytes()


def tev_update_image(channels, name="render"):
    """TevControlUpdateImage::new_message (src/headless.rs:127-163): channels = (H, W, 3) resolved film ->
    the concatenated 100 x 100-tile messages."""
    img = _f32(channels)
    h, w, _ = img.shape
    planar = np.ascontiguousarray(np.moveaxis(img, 2, 0))
    n = lib().ptrs_host_tev_update_image(_fp(planar), w, h, name.encode(), None, 0)
    buf = np.empty(n, dtype=np.uint8)
    lib().ptrs_host_tev_update_image(_fp(planar), w, h, name.encode(), buf.ctypes.data_as(C.POINTER(C.c_uint8)), n)
    return buf.tobytes()


def snake_case(name):
    buf = C.create_string_buffer(4 * len(name) + 8)
    n = lib().ptrs_host_snake_case(name.encode(), buf, len(buf))
    return buf.value[:n].decode()


def default_render_params(spp=1, max_depth=15):
    p = PtrsRenderParams()
    lib().ptrs_host_default_render_params(C.byref(p))
    p.spp = spp
    p.max_depth = max_depth
    return p


def look_at_camera(eye, target, up, fovy_deg, width, height):
    cam = PtrsCamera()
    e, t, u = (_f32(v) for v in (eye, target, up))
    lib().ptrs_host_look_at_camera(_fp(e), _fp(t), _fp(u), fovy_deg, width, height, C.byref(cam))
    return cam


def mitsuba_camera(sensor_to_world, fov_deg, film_w, film_h, res_w, res_h):
    cam = PtrsCamera()
    m = _f32(sensor_to_world, (16,))
    lib().ptrs_host_mitsuba_camera(_fp(m), fov_deg, film_w, film_h, res_w, res_h, C.byref(cam))
    return cam


def coherent_rays(cam, side):
    rays = np.empty(side * side, dtype=RAY_DTYPE)
    lib().ptrs_host_coherent_rays(C.byref(cam), side, rays.ctypes.data_as(C.POINTER(PtrsRay)))
    return rays


def incoherent_rays(bmin, bmax, seed, n):
    rays = np.empty(n, dtype=RAY_DTYPE)
    a, b = _f32(bmin), _f32(bmax)
    lib().ptrs_host_incoherent_rays(_fp(a), _fp(b), seed, n, rays.ctypes.data_as(C.POINTER(PtrsRay)))
    return rays


def synth_sky(w=1024, h=512, seed=1):
    out = np.empty((h, w, 3), dtype=np.float32)
    lib().ptrs_host_synth_sky(w, h, seed, _fp(out))
    return out


def build_bvh(bounds, max_prims=4, n_threads=1):
    """bounds: (n, 6) float32 [min, max]. Returns (nodes structured array, prim_order)."""
    b = _f32(bounds, (-1, 6))
    n = b.shape[0]
    nodes = np.empty(2 * n + 1, dtype=NODE_DTYPE)
    order = np.empty(n, dtype=np.uint32)
    cnt = C.c_uint32(0)
    rc = lib().ptrs_host_build_bvh(_fp(b), n, max_prims, n_threads, nodes.ctypes.data_as(C.POINTER(PtrsBvhNode)),
                                   nodes.shape[0], C.byref(cnt), order.ctypes.data_as(C.POINTER(C.c_uint32)))
    if rc != 0:
        raise RuntimeError(lib().ptrs_host_last_error().decode(errors="replace"))
    return nodes[: cnt.value].copy(), order


class SceneBuilder:
    """Python face of ptrs_host::SceneBuilder (importer-side assembly, src/pathtracer/importer/*.rs)."""

    def __init__(self):
        self._b = lib().ptrs_host_builder_new()

    def __del__(self):
        if getattr(self, "_b", None):
            lib().ptrs_host_builder_free(self._b)
            self._b = None

    def constant_texture(self, value):
        v = np.atleast_1d(np.asarray(value, dtype=np.float32))
        if v.size == 1:
            return lib().ptrs_host_add_constant_texture(self._b, 1, float(v[0]), 0.0, 0.0)
        return lib().ptrs_host_add_constant_texture(self._b, 3, float(v[0]), float(v[1]), float(v[2]))

    def checker_texture(self, v1, v2, su=1.0, sv=1.0, du=0.0, dv=0.0, channels=3):
        a, b = _f32(np.resize(v1, 3)), _f32(np.resize(v2, 3))
        return lib().ptrs_host_add_checker_texture(self._b, channels, _fp(a), _fp(b), su, sv, du, dv)

    def image_texture(self, image, wrap=_abi.WRAP_REPEAT, su=1.0, sv=1.0, du=0.0, dv=0.0):
        img = _f32(image)
        if img.ndim == 2:
            img = img[:, :, None]
        h, w, c = img.shape
        r = lib().ptrs_host_add_image_texture(self._b, c, _fp(img), w, h, wrap, su, sv, du, dv)
        if r < 0:
            raise RuntimeError(lib().ptrs_host_last_error().decode(errors="replace"))
        return r

    def material(self, mtype, tex=(), remap_roughness=False, normal_map=-1):
        m = PtrsMaterial()
        m.type = mtype
        m.normal_map = normal_map
        for i in range(5):
            m.tex[i] = tex[i] if i < len(tex) else -1
        m.remap_roughness = 1 if remap_roughness else 0
        return lib().ptrs_host_add_material(self._b, C.byref(m))

    def mesh(self, pos, indices, normal=None, tangent=None, uv=None, transform=None, material=0, alpha_tex=-1, ke_tex=-1):
        p = _f32(pos, (-1, 3))
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        n, t, u = _f32(normal, (-1, 3)), _f32(tangent, (-1, 3)), _f32(uv, (-1, 2))
        x = _f32(transform, (16,))
        r = lib().ptrs_host_add_mesh(self._b, _fp(p), p.shape[0], _fp(n), _fp(t), _fp(u),
                                     idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.shape[0], _fp(x), material, alpha_tex, ke_tex)
        if r < 0:
            raise RuntimeError(lib().ptrs_host_last_error().decode(errors="replace"))
        return r

    def shape(self, kind, transform, material, ke_tex=-1):
        """kind: 'rectangle' | 'cube' (Mitsuba shapes, common/importer/mitsuba.rs:20-58)."""
        x = _f32(transform, (16,))
        return lib().ptrs_host_add_shape(self._b, {"rectangle": 0, "cube": 1}[kind], _fp(x), material, ke_tex)

    def point_light(self, transform, intensity):
        x, i = _f32(transform, (16,)), _f32(intensity, (3,))
        return lib().ptrs_host_add_point_light(self._b, _fp(x), _fp(i))

    def directional_light(self, transform, radiance, w_light):
        x, l, w = _f32(transform, (16,)), _f32(radiance, (3,)), _f32(w_light, (3,))
        return lib().ptrs_host_add_directional_light(self._b, _fp(x), _fp(l), _fp(w))

    def infinite_light(self, transform, rgb_image):
        x, img = _f32(transform, (16,)), _f32(rgb_image)
        h, w, _ = img.shape
        r = lib().ptrs_host_add_infinite_light(self._b, _fp(x), _fp(img), w, h)
        if r < 0:
            raise RuntimeError(lib().ptrs_host_last_error().decode(errors="replace"))
        return r

    def finalize(self, max_prims_in_node=4, n_threads=0):
        return FlatScene(lib().ptrs_host_finalize(self._b, max_prims_in_node, n_threads))


def mitsuba_env_light_to_world():
    m = np.empty(16, dtype=np.float32)
    lib().ptrs_host_mitsuba_env_light_to_world(_fp(m))
    return m.reshape(4, 4)
