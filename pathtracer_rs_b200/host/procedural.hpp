// Seeded synthetic scenes of the shapes named in BASELINE.json's configs, plus the fixed ray sets
// of the intersection microbenchmark.  Everything is deterministic in (kind, seed, size).
#pragma once
#include <cstdint>
#include <vector>

#include "scene_builder.hpp"

namespace ptrs_host {

// genmesh generators as the Mitsuba importer uses them (src/common/importer/mitsuba.rs:20-79):
// positions + per-vertex normals, no UVs.
MeshInput gen_rectangle();  // Plane::new(): [-1,1]^2 in z = 0, normal +z, 2 triangles
MeshInput gen_cube();       // Cube::new(): [-1,1]^3, 24 vertices with face normals, 12 triangles
MeshInput gen_sphere_uv(int u, int v, bool with_uv);  // unit sphere, (u x v) lat-long tessellation

// data/cornell-box.xml transcribed as data: 8 matte materials, 5 wall rectangles, 2 cubes and the
// emissive rectangle (radiance 17, 12, 4) in the file's shape order.  with_env adds an
// InfiniteAreaLight with the importer's env_light_to_world (importer/mitsuba.rs:365-372) over the
// given lat-long RGB image (what `<emitter type="sunsky"/>` maps to, :400-418).
void build_cornell(SceneBuilder& b, const float* env_rgb, int env_w, int env_h);
PtrsCamera cornell_camera(int res_w, int res_h);
M4 mitsuba_env_light_to_world();

// 1024x512-style procedural HDR sky (sun disc + horizon gradient + seeded clouds), stands in for
// data/abandoned_tank_farm_04_1k.hdr which does not travel with the repo.
std::vector<float> synth_sky(int w, int h, uint64_t seed);

// C3: ~n_tris triangles of tessellated spheres / boxes on a ground plane, materials cycled over
// glass / substrate / metal / Disney / matte, emissive quads and the synthetic sky.
void build_material_field(SceneBuilder& b, uint64_t seed, size_t n_tris, PtrsCamera* cam, int res_w, int res_h);
// C4: ~n_tris-triangle displaced terrain with scattered spheres inside the unit cube (all matte).
void build_terrain(SceneBuilder& b, uint64_t seed, size_t n_tris, PtrsCamera* cam, int res_w, int res_h);
// C5: "atrium" of columns, arches and draped cloth grids, directional + env light.
void build_atrium(SceneBuilder& b, uint64_t seed, size_t n_tris, PtrsCamera* cam, int res_w, int res_h);

// Ray sets for the microbenchmark.
// coherent: the camera's pixel-centre rays in scanline order (side x side of them)
void coherent_rays(const PtrsCamera& cam, int side, PtrsRay* out);
// incoherent: origins uniform in [mn, mx], directions uniform on the sphere, t_max = inf
void incoherent_rays(const float mn[3], const float mx[3], uint64_t seed, size_t n, PtrsRay* out);

PtrsCamera look_at_camera(V3 eye, V3 target, V3 up, float fovy_deg, int res_w, int res_h);

}  // namespace ptrs_host
