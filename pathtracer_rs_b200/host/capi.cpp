// C exports of the host-side scene preparation code, for the Python harness (ctypes) and any other
// FFI consumer.  Nothing here touches the GPU; the output is a PtrsSceneDesc for ptrs_scene_create.
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

#include "image_io.hpp"
#include "importers.hpp"
#include "procedural.hpp"
#include "scene_builder.hpp"
#include "tev.hpp"

using namespace ptrs_host;

namespace {
thread_local std::string g_err;
struct SceneBox {
  FlatScene fs;
  PtrsSceneDesc desc;
  PtrsSceneDesc slim;
};
template <class F>
int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
M4 to_m4(const float* m) {
  M4 r;
  if (m) std::memcpy(r.m, m, 64);
  else r = M4::identity();
  return r;
}
}  // namespace

extern "C" {

const char* ptrs_host_last_error() { return g_err.c_str(); }

void* ptrs_host_builder_new() { return new SceneBuilder(); }
void ptrs_host_builder_free(void* b) { delete (SceneBuilder*)b; }

int ptrs_host_add_constant_texture(void* b, int channels, float x, float y, float z) {
  return ((SceneBuilder*)b)->add_constant_texture(channels, x, y, z);
}
int ptrs_host_add_checker_texture(void* b, int channels, const float* v1, const float* v2, float su, float sv, float du, float dv) {
  return ((SceneBuilder*)b)->add_checker_texture(channels, v1, v2, su, sv, du, dv);
}
int ptrs_host_add_image_texture(void* b, int channels, const float* img, int w, int h, int wrap, float su, float sv, float du, float dv) {
  int id = -1;
  if (guard([&] { id = ((SceneBuilder*)b)->add_image_texture(channels, img, w, h, wrap, su, sv, du, dv); })) return -1;
  return id;
}
int ptrs_host_add_material(void* b, const PtrsMaterial* m) { return ((SceneBuilder*)b)->add_material(*m); }
void ptrs_host_set_normal_map(void* b, int mat, int tex) { ((SceneBuilder*)b)->set_normal_map(mat, tex); }

int ptrs_host_add_mesh(void* b, const float* pos, uint32_t n_verts, const float* normal, const float* tangent,
                       const float* uv, const uint32_t* indices, uint32_t n_tris, const float* xform16,
                       int material, int alpha_tex, int ke_tex) {
  int id = -1;
  if (guard([&] {
        MeshInput m;
        m.pos.assign(pos, pos + 3 * (size_t)n_verts);
        if (normal) m.normal.assign(normal, normal + 3 * (size_t)n_verts);
        if (tangent) m.tangent.assign(tangent, tangent + 3 * (size_t)n_verts);
        if (uv) m.uv.assign(uv, uv + 2 * (size_t)n_verts);
        m.indices.assign(indices, indices + 3 * (size_t)n_tris);
        m.obj_to_world = to_m4(xform16);
        m.material = material;
        m.alpha_tex = alpha_tex;
        m.ke_tex = ke_tex;
        id = ((SceneBuilder*)b)->add_mesh(m);
      }))
    return -1;
  return id;
}
// shape: 0 rectangle, 1 cube (Mitsuba <shape type=...>, common/importer/mitsuba.rs:20-58)
int ptrs_host_add_shape(void* b, int shape, const float* xform16, int material, int ke_tex) {
  int id = -1;
  if (guard([&] {
        MeshInput m = shape == 0 ? gen_rectangle() : gen_cube();
        m.obj_to_world = to_m4(xform16);
        m.material = material;
        m.ke_tex = ke_tex;
        id = ((SceneBuilder*)b)->add_mesh(m);
      }))
    return -1;
  return id;
}
int ptrs_host_add_point_light(void* b, const float* xform16, const float* intensity) {
  return ((SceneBuilder*)b)->add_point_light(to_m4(xform16), intensity);
}
int ptrs_host_add_directional_light(void* b, const float* xform16, const float* l, const float* w) {
  return ((SceneBuilder*)b)->add_directional_light(to_m4(xform16), l, w);
}
int ptrs_host_add_infinite_light(void* b, const float* xform16, const float* rgb, int w, int h) {
  int id = -1;
  if (guard([&] { id = ((SceneBuilder*)b)->add_infinite_light(to_m4(xform16), rgb, w, h); })) return -1;
  return id;
}
void ptrs_host_mitsuba_env_light_to_world(float* out16) {
  M4 m = mitsuba_env_light_to_world();
  std::memcpy(out16, m.m, 64);
}

void* ptrs_host_finalize(void* b, int max_prims_in_node, int n_threads) {
  SceneBox* box = nullptr;
  if (guard([&] {
        box = new SceneBox();
        box->fs = ((SceneBuilder*)b)->finalize(max_prims_in_node, n_threads);
        box->desc = box->fs.desc();
      })) {
    delete box;
    return nullptr;
  }
  return box;
}
void ptrs_host_scene_free(void* s) { delete (SceneBox*)s; }
const PtrsSceneDesc* ptrs_host_scene_desc(void* s) { return &((SceneBox*)s)->desc; }
// the same scene without the tables the device library builds itself (level-0-only pyramids, no Distribution2D arrays)
const PtrsSceneDesc* ptrs_host_scene_desc_device_tables(void* s) {
  SceneBox* b = (SceneBox*)s;
  b->slim = b->fs.desc_device_tables();
  return &b->slim;
}
uint64_t ptrs_host_scene_bytes_device_tables(void* s) { return ((SceneBox*)s)->fs.host_bytes_device_tables(); }
uint64_t ptrs_host_scene_bytes(void* s) { return ((SceneBox*)s)->fs.host_bytes(); }
int ptrs_host_scene_bvh_depth(void* s) { return ((SceneBox*)s)->fs.bvh_max_depth; }
double ptrs_host_scene_bvh_seconds(void* s) { return ((SceneBox*)s)->fs.bvh_build_seconds; }

// ---- ready-made scenes ------------------------------------------------------------------------
// kind: 0 cornell, 1 cornell + synthetic sky, 2 material field (C3), 3 terrain (C4), 4 atrium (C5)
// env_hdr (kind 1 only; may be NULL): Radiance .hdr file used as the environment map — the reference's
// `<emitter type="sunsky"/>` maps to data/abandoned_tank_farm_04_1k.hdr (pathtracer/importer/mitsuba.rs:400-418)
void* ptrs_host_make_scene_env(int kind, uint64_t seed, uint64_t n_tris, int res_w, int res_h, PtrsCamera* cam,
                               int n_threads, const char* env_hdr) {
  SceneBox* box = nullptr;
  if (guard([&] {
        SceneBuilder b;
        PtrsCamera c{};
        switch (kind) {
          case 0:
            build_cornell(b, nullptr, 0, 0);
            c = cornell_camera(res_w, res_h);
            break;
          case 1: {
            if (env_hdr && *env_hdr) {
              const ImageF32 img = load_hdr(env_hdr);
              build_cornell(b, img.data.data(), img.width, img.height);
            } else {
              std::vector<float> sky = synth_sky(1024, 512, seed);
              build_cornell(b, sky.data(), 1024, 512);
            }
            c = cornell_camera(res_w, res_h);
            break;
          }
          case 2: build_material_field(b, seed, n_tris, &c, res_w, res_h); break;
          case 3: build_terrain(b, seed, n_tris, &c, res_w, res_h); break;
          case 4: build_atrium(b, seed, n_tris, &c, res_w, res_h); break;
          default: throw std::runtime_error("unknown scene kind");
        }
        if (cam) *cam = c;
        box = new SceneBox();
        box->fs = b.finalize(4, n_threads);
        box->desc = box->fs.desc();
      })) {
    delete box;
    return nullptr;
  }
  return box;
}

void* ptrs_host_make_scene(int kind, uint64_t seed, uint64_t n_tris, int res_w, int res_h, PtrsCamera* cam, int n_threads) {
  return ptrs_host_make_scene_env(kind, seed, n_tris, res_w, res_h, cam, n_threads, nullptr);
}

// ---- scene files (importers.hpp) and image files (image_io.hpp) ----------------------------------
// common::importer::import(log, path, resolution, default_lights), src/common/importer/mod.rs:6-25
void* ptrs_host_import_scene(const char* path, int res_w, int res_h, int default_lights, const char* sunsky_hdr, PtrsCamera* cam,
                             int n_threads) {
  SceneBox* box = nullptr;
  if (guard([&] {
        SceneBuilder b;
        ImportOptions opt;
        opt.res_w = res_w;
        opt.res_h = res_h;
        opt.default_lights = default_lights != 0;
        if (sunsky_hdr) opt.sunsky_hdr = sunsky_hdr;
        const PtrsCamera c = import_scene(path, opt, b);
        if (cam) *cam = c;
        box = new SceneBox();
        box->fs = b.finalize(4, n_threads);
        box->desc = box->fs.desc();
      })) {
    delete box;
    return nullptr;
  }
  return box;
}
// two-call pattern: out == NULL returns the shape only
int ptrs_host_load_hdr(const char* path, int* w, int* h, float* out_rgb) {
  return guard([&] {
    const ImageF32 img = load_hdr(path);
    *w = img.width;
    *h = img.height;
    if (out_rgb) std::memcpy(out_rgb, img.data.data(), img.data.size() * 4);
  });
}
int ptrs_host_save_hdr(const char* path, const float* rgb, int w, int h) {
  return guard([&] { save_hdr(path, rgb, w, h); });
}
int ptrs_host_load_png(const char* path, int* w, int* h, int* channels, uint8_t* out) {
  return guard([&] {
    const ImageU8 img = load_png(path);
    *w = img.width;
    *h = img.height;
    *channels = img.channels;
    if (out) std::memcpy(out, img.data.data(), img.data.size());
  });
}
// PNG or JPEG bytes -> 8-bit pixels; call with out == NULL first for the size
int ptrs_host_decode_image(const uint8_t* bytes, size_t n, int* w, int* h, int* channels, uint8_t* out) {
  return guard([&] {
    const ImageU8 img = decode_image(bytes, n);
    *w = img.width;
    *h = img.height;
    *channels = img.channels;
    if (out) std::memcpy(out, img.data.data(), img.data.size());
  });
}
int ptrs_host_save_png(const char* path, const uint8_t* pixels, int w, int h, int channels) {
  return guard([&] { save_png(path, pixels, w, h, channels); });
}
int ptrs_host_snake_case(const char* in, char* out, int cap) {
  const std::string s = snake_case(in);
  if ((int)s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// tev wire messages (tev.hpp); returns the byte count, copies when `out` has room
int64_t ptrs_host_tev_create_image(int w, int h, const char* name, uint8_t* out, int64_t cap) {
  const std::vector<uint8_t> m = tev_create_image(w, h, name);
  if (out && (int64_t)m.size() <= cap) std::memcpy(out, m.data(), m.size());
  return (int64_t)m.size();
}
// all UpdateImage messages of one film, concatenated in sending order
int64_t ptrs_host_tev_update_image(const float* rgb_planar, int w, int h, const char* name, uint8_t* out, int64_t cap) {
  const float* ch[3] = {rgb_planar, rgb_planar + (size_t)w * h, rgb_planar + 2 * (size_t)w * h};
  int64_t total = 0;
  for (const auto& m : tev_update_image(ch, w, h, name)) {
    if (out && total + (int64_t)m.size() <= cap) std::memcpy(out + total, m.data(), m.size());
    total += (int64_t)m.size();
  }
  return total;
}

void ptrs_host_synth_sky(int w, int h, uint64_t seed, float* out) {
  std::vector<float> s = synth_sky(w, h, seed);
  std::memcpy(out, s.data(), s.size() * 4);
}

// ---- camera / params --------------------------------------------------------------------------
void ptrs_host_make_camera(const float* rot_quat, const float* trans, float aspect, float fovy, float znear,
                           float zfar, int w, int h, PtrsCamera* out) {
  *out = make_camera(rot_quat, trans, aspect, fovy, znear, zfar, w, h);
}
void ptrs_host_mitsuba_camera(const float* sensor16, float fov_deg, int film_w, int film_h, int res_w, int res_h,
                              PtrsCamera* out) {
  *out = mitsuba_camera(to_m4(sensor16), fov_deg, film_w, film_h, res_w, res_h);
}
void ptrs_host_look_at_camera(const float* eye, const float* target, const float* up, float fovy_deg, int w, int h,
                              PtrsCamera* out) {
  *out = look_at_camera(v3(eye[0], eye[1], eye[2]), v3(target[0], target[1], target[2]), v3(up[0], up[1], up[2]),
                        fovy_deg, w, h);
}
void ptrs_host_default_render_params(PtrsRenderParams* p) { default_render_params(p); }
void ptrs_host_gaussian_filter_table(float alpha, float radius, float* table256) {
  gaussian_filter_table(alpha, radius, table256);
}

// ---- ray sets ---------------------------------------------------------------------------------
void ptrs_host_coherent_rays(const PtrsCamera* cam, int side, PtrsRay* out) { coherent_rays(*cam, side, out); }
void ptrs_host_incoherent_rays(const float* mn, const float* mx, uint64_t seed, uint64_t n, PtrsRay* out) {
  incoherent_rays(mn, mx, seed, n, out);
}

// BVH build alone (tests): bounds = n * 6 floats (min, max)
int ptrs_host_build_bvh(const float* bounds6, uint32_t n, int max_prims, int n_threads, PtrsBvhNode* nodes_out,
                        uint32_t nodes_cap, uint32_t* n_nodes, uint32_t* prim_order_out) {
  return guard([&] {
    std::vector<Bounds3> b(n);
    std::memcpy(b.data(), bounds6, (size_t)n * 24);
    BvhBuildResult r = build_bvh(b, max_prims, n_threads);
    *n_nodes = (uint32_t)r.nodes.size();
    if (r.nodes.size() > nodes_cap) throw std::runtime_error("node capacity too small");
    std::memcpy(nodes_out, r.nodes.data(), r.nodes.size() * sizeof(PtrsBvhNode));
    std::memcpy(prim_order_out, r.prim_order.data(), (size_t)n * 4);
  });
}

}  // extern "C"
