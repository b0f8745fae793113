"""ctypes mirrors of the POD structs in include/ptrs_b200.h (keep in lock-step with the header)."""
import ctypes as C

ABI_VERSION = 1
FILTER_TABLE_WIDTH = 16
MAX_MIP_LEVELS = 16


class PtrsRay(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("d", C.c_float * 3), ("t_max", C.c_float)]


class PtrsHit(C.Structure):
    _fields_ = [("prim", C.c_int32), ("t", C.c_float), ("b0", C.c_float), ("b1", C.c_float), ("b2", C.c_float)]


class PtrsBvhNode(C.Structure):
    _fields_ = [("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3), ("offset", C.c_uint32),
                ("n_prims", C.c_uint16), ("axis", C.c_uint8), ("pad", C.c_uint8)]


class PtrsMesh(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("alpha_tex", C.c_int32)]


class PtrsTexture(C.Structure):
    _fields_ = [("type", C.c_int32), ("channels", C.c_int32), ("v1", C.c_float * 3), ("v2", C.c_float * 3),
                ("su", C.c_float), ("sv", C.c_float), ("du", C.c_float), ("dv", C.c_float),
                ("mip", C.c_int32), ("pad", C.c_int32)]


class PtrsMipMap(C.Structure):
    _fields_ = [("channels", C.c_int32), ("wrap", C.c_int32), ("n_levels", C.c_int32),
                ("width", C.c_int32 * MAX_MIP_LEVELS), ("height", C.c_int32 * MAX_MIP_LEVELS),
                ("level_offset", C.c_uint64 * MAX_MIP_LEVELS)]


class PtrsMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("normal_map", C.c_int32), ("tex", C.c_int32 * 5), ("remap_roughness", C.c_int32)]


class PtrsLight(C.Structure):
    _fields_ = [("type", C.c_int32), ("prim", C.c_int32), ("ke_tex", C.c_int32), ("env", C.c_int32),
                ("pos", C.c_float * 3), ("color", C.c_float * 3), ("area", C.c_float), ("world_radius", C.c_float),
                ("world_center", C.c_float * 3), ("pad", C.c_float)]


class PtrsEnvLight(C.Structure):
    _fields_ = [("light_to_world", C.c_float * 16), ("world_to_light", C.c_float * 16), ("mip", C.c_int32),
                ("nu", C.c_int32), ("nv", C.c_int32), ("pad", C.c_int32),
                ("cond_func", C.POINTER(C.c_float)), ("cond_cdf", C.POINTER(C.c_float)),
                ("cond_func_int", C.POINTER(C.c_float)), ("marg_func", C.POINTER(C.c_float)),
                ("marg_cdf", C.POINTER(C.c_float)), ("marg_func_int", C.c_float), ("pad2", C.c_float)]


class PtrsSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_nodes", C.c_uint32), ("nodes", C.POINTER(PtrsBvhNode)),
                ("n_prims", C.c_uint32), ("prim_vertex", C.POINTER(C.c_uint32)), ("prim_mesh", C.POINTER(C.c_int32)),
                ("prim_material", C.POINTER(C.c_int32)), ("prim_area_light", C.POINTER(C.c_int32)),
                ("n_verts", C.c_uint32), ("pos", C.POINTER(C.c_float)), ("normal", C.POINTER(C.c_float)),
                ("tangent", C.POINTER(C.c_float)), ("uv", C.POINTER(C.c_float)),
                ("n_meshes", C.c_uint32), ("meshes", C.POINTER(PtrsMesh)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(PtrsMaterial)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(PtrsTexture)),
                ("n_mipmaps", C.c_uint32), ("mipmaps", C.POINTER(PtrsMipMap)),
                ("n_texels", C.c_uint64), ("texels", C.POINTER(C.c_float)),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(PtrsLight)),
                ("n_infinite_lights", C.c_uint32), ("infinite_lights", C.POINTER(C.c_int32)),
                ("n_envs", C.c_uint32), ("envs", C.POINTER(PtrsEnvLight))]


class PtrsCamera(C.Structure):
    _fields_ = [("rot", C.c_float * 4), ("trans", C.c_float * 3), ("pad0", C.c_float),
                ("raster_to_screen", C.c_float * 16), ("persp", C.c_float * 4),
                ("dx_camera", C.c_float * 3), ("dy_camera", C.c_float * 3), ("width", C.c_int32), ("height", C.c_int32)]


class PtrsRenderParams(C.Structure):
    _fields_ = [("spp", C.c_int32), ("max_depth", C.c_int32), ("rr_threshold", C.c_float), ("rr_start_depth", C.c_int32),
                ("rr_enable", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("sample_stride", C.c_int32), ("sample_phase", C.c_int32), ("filter_radius", C.c_float * 2),
                ("filter_table", C.c_float * (FILTER_TABLE_WIDTH * FILTER_TABLE_WIDTH)),
                ("paths_per_batch", C.c_int32), ("flags", C.c_int32)]


class PtrsStats(C.Structure):
    _fields_ = [("camera_paths", C.c_uint64), ("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("mis_rays", C.c_uint64), ("nodes_tested", C.c_uint64), ("tris_tested", C.c_uint64),
                ("nee_nodes_tested", C.c_uint64), ("nee_tris_tested", C.c_uint64),
                ("ms_generate", C.c_float), ("ms_extend", C.c_float), ("ms_shade", C.c_float), ("ms_shadow", C.c_float),
                ("ms_accumulate", C.c_float), ("ms_total", C.c_float), ("launches", C.c_uint32), ("batches", C.c_uint32),
                ("extend_launches", C.c_uint32), ("connect_launches", C.c_uint32),
                ("ms_connect_trace", C.c_float), ("ms_resolve", C.c_float)]


class PtrsLobeDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("fresnel", C.c_int32), ("r", C.c_float * 3), ("t", C.c_float * 3), ("fa", C.c_float * 3),
                ("fb", C.c_float * 3), ("eta_a", C.c_float), ("eta_b", C.c_float), ("alpha_x", C.c_float), ("alpha_y", C.c_float),
                ("disney_g", C.c_int32), ("pad", C.c_int32)]


RENDER_EXACT_SHADING = 1
(LOBE_LAMBERTIAN, LOBE_SPECULAR_REFLECTION, LOBE_SPECULAR_TRANSMISSION, LOBE_FRESNEL_SPECULAR, LOBE_MICROFACET_REFLECTION,
 LOBE_MICROFACET_TRANSMISSION, LOBE_FRESNEL_BLEND, LOBE_DISNEY_DIFFUSE) = range(8)
FRESNEL_DIELECTRIC, FRESNEL_CONDUCTOR, FRESNEL_DISNEY, FRESNEL_NOOP = range(4)

# material / texture / light / wrap enums
MAT_MATTE, MAT_MIRROR, MAT_GLASS, MAT_METAL, MAT_SUBSTRATE, MAT_DISNEY = range(6)
TEX_CONSTANT, TEX_CHECKER, TEX_IMAGE = range(3)
WRAP_REPEAT, WRAP_BLACK, WRAP_CLAMP = range(3)
LIGHT_POINT, LIGHT_DIRECTIONAL, LIGHT_AREA, LIGHT_INFINITE = range(4)
MESH_HAS_NORMAL, MESH_HAS_TANGENT, MESH_HAS_UV = 1, 2, 4
