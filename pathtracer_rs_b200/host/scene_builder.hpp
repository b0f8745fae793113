// Host-side scene assembly: what the reference's importers do after parsing
// (src/pathtracer/importer/mitsuba.rs:198-429, importer/gltf.rs:378-584) — transform meshes to world
// space, create one GeometricPrimitive per triangle and one DiffuseAreaLight per emissive triangle,
// build MIP pyramids and the env-map Distribution2D, run BVH::new(.., 4) — ending in the flat
// PtrsSceneDesc the device library consumes.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/ptrs_b200.h"
#include "bvh_build.hpp"
#include "vecmath.hpp"

namespace ptrs_host {

// MIPMap<T> (src/pathtracer/texture.rs:238-465), T = f32 (channels 1) or Spectrum/Vector3 (3).
struct HostMipMap {
  int channels = 3;
  int wrap = PTRS_WRAP_REPEAT;
  std::vector<int> width, height;
  std::vector<std::vector<float>> levels;  // row-major, `channels` floats per texel

  // MIPMap::new (texture.rs:279-405) incl. the Lanczos resample to powers of two.
  static HostMipMap build(const float* image, int width, int height, int channels, int wrap);
  void texel(int level, int s, int t, float* out) const;          // texture.rs:245-273
  void triangle(int level, float s, float t, float* out) const;   // texture.rs:413-429
  void lookup_width(float s, float t, float width, float* out) const;  // texture.rs:447-464
};

// Distribution1D / Distribution2D (src/pathtracer/sampling.rs:128-230) in flat arrays.
struct HostDistribution2D {
  int nu = 0, nv = 0;
  std::vector<float> cond_func, cond_cdf, cond_func_int, marg_func, marg_cdf;
  float marg_func_int = 0.0f;
  static HostDistribution2D build(const float* func, int nu, int nv);
};

struct MeshInput {
  std::vector<float> pos;      // 3 * nv, object space
  std::vector<float> normal;   // 3 * nv or empty
  std::vector<float> tangent;  // 3 * nv or empty
  std::vector<float> uv;       // 2 * nv or empty
  std::vector<uint32_t> indices;  // 3 * nt
  M4 obj_to_world = M4::identity();
  int material = 0;
  int alpha_tex = -1;
  int ke_tex = -1;             // >= 0: every triangle becomes a DiffuseAreaLight with this ke ...
  std::vector<uint8_t> tri_emits;  // ... unless this per-triangle mask (glTF emissive textures) says otherwise
};

struct FlatScene {
  std::vector<PtrsBvhNode> nodes;
  std::vector<uint32_t> prim_vertex;
  std::vector<int32_t> prim_mesh, prim_material, prim_area_light;
  std::vector<float> pos, normal, tangent, uv;
  std::vector<PtrsMesh> meshes;
  std::vector<PtrsMaterial> materials;
  std::vector<PtrsTexture> textures;
  std::vector<PtrsMipMap> mipmaps;
  std::vector<float> texels;
  std::vector<PtrsLight> lights;
  std::vector<int32_t> infinite_lights;
  std::vector<PtrsEnvLight> envs;
  std::vector<HostDistribution2D> env_dists;  // storage behind envs[i] pointers
  int bvh_max_depth = 0;
  double bvh_build_seconds = 0.0;

  PtrsSceneDesc desc() const;  // pointers into this object; valid while it is alive and unmoved
  // The same scene with the tables the device library can build itself left out (include/ptrs_b200.h): every MIP
  // pyramid carries level 0 only and the env lights carry no Distribution2D arrays.
  PtrsSceneDesc desc_device_tables() const;
  uint64_t host_bytes() const;
  uint64_t host_bytes_device_tables() const;
  // storage behind desc_device_tables()
  mutable std::vector<PtrsMipMap> slim_mipmaps;
  mutable std::vector<float> slim_texels;
  mutable std::vector<PtrsEnvLight> slim_envs;
};

class SceneBuilder {
 public:
  int add_constant_texture(int channels, float a, float b = 0.f, float c = 0.f);
  int add_checker_texture(int channels, const float v1[3], const float v2[3], float su, float sv,
                          float du, float dv);
  // image: row-major, `channels` floats per texel, already converted like ImageTexture::new does
  int add_image_texture(int channels, const float* image, int width, int height, int wrap, float su,
                        float sv, float du, float dv);
  int add_material(const PtrsMaterial& m);
  // convenience constructors mirroring importer/mitsuba.rs:84-181
  int add_matte(int kd_tex);
  int add_mirror();
  int add_glass(int kr_tex, int kt_tex, int index_tex);
  int add_metal(int eta_tex, int k_tex, int r_tex, int urough_tex, int vrough_tex, bool remap);
  int add_substrate(int kd_tex, int ks_tex, int nu_tex, int nv_tex, bool remap);
  int add_disney(int color_tex, int metallic_tex, int eta_tex, int roughness_tex);
  void set_normal_map(int material, int normal_tex);

  int add_mesh(const MeshInput& mesh);  // returns mesh id; appends area lights in triangle order
  int add_point_light(const M4& light_to_world, const float intensity[3]);
  int add_directional_light(const M4& light_to_world, const float l[3], const float w_light[3]);
  // InfiniteAreaLight::new (light.rs:349-399): texels = l * image (Spectrum), lat-long map
  int add_infinite_light(const M4& light_to_world, const float* rgb, int width, int height);

  size_t triangle_count() const { return tri_vertex_.size() / 3; }
  // host copy of an image texture's pyramid (importers evaluate emission maps with it), or null
  const HostMipMap* image_texture_mip(int texture) const;
  FlatScene finalize(int max_prims_in_node = 4, int n_threads = 0);

 private:
  int push_mip(const HostMipMap& mm);
  std::vector<float> pos_, normal_, tangent_, uv_;
  bool any_normal_ = false, any_tangent_ = false, any_uv_ = false;
  std::vector<uint32_t> tri_vertex_;
  std::vector<int32_t> tri_mesh_, tri_material_, tri_light_;
  std::vector<PtrsMesh> meshes_;
  std::vector<PtrsMaterial> materials_;
  std::vector<PtrsTexture> textures_;
  std::vector<PtrsMipMap> mipmaps_;
  std::vector<float> texels_;
  std::vector<PtrsLight> lights_;       // light.prim holds the INPUT triangle index until finalize
  std::vector<int32_t> infinite_lights_;
  std::vector<PtrsEnvLight> envs_;
  std::vector<HostDistribution2D> env_dists_;
  std::vector<HostMipMap> env_mips_;
  std::map<int, HostMipMap> tex_mips_;  // image texture id -> pyramid
};

// Camera::new (src/common/mod.rs:33-62) from an isometry + Perspective3::new(aspect, fovy, n, f).
PtrsCamera make_camera(const float rot_quat_ijkw[4], const float trans[3], float aspect, float fovy,
                       float znear, float zfar, int width, int height);
// rotation matrix (row-major 3x3, assumed orthonormal) -> unit quaternion (i, j, k, w)
void quat_from_matrix(const float r[9], float q[4]);
// get_camera (src/common/importer/mitsuba.rs:685-710): sensor toWorld (row-major 4x4), fov degrees,
// film width/height from the XML, render resolution from the CLI.
PtrsCamera mitsuba_camera(const M4& sensor_to_world, float fov_deg, int film_w, int film_h, int res_w,
                          int res_h);

// Film::new's filter table for the Gaussian(alpha, radius) filter (film.rs:135-144, filter.rs:61-89)
void gaussian_filter_table(float alpha, float radius, float table[256]);
void default_render_params(PtrsRenderParams* p);

}  // namespace ptrs_host
