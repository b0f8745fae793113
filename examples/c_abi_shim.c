/* A plain-C caller of libptrs_b200.so that fills PtrsSceneDesc the way the Rust shim does
 * (integration/rust/b200.rs FlatScene::from_render_scene + tables.rs): the BVH node array verbatim, one entry per
 * BVH-ordered primitive in the four primitive arrays, mesh-major vertex pools, one DiffuseAreaLight per emissive
 * triangle — with no help from the C++ host library.  It doubles as the check that include/ptrs_b200.h is valid C.
 *
 * Scene: a floor quad (2 triangles, matte) under a small emissive quad (2 triangles, black matte, Le = 10), tree = one
 * interior node over two leaves.  Exit code 0 = every call behaved; messages on stderr otherwise.
 *   make -C examples && ./examples/c_abi_shim */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ptrs_b200.h"

#define CHECK(call)                                                                    \
  do {                                                                                 \
    int32_t rc__ = (call);                                                             \
    if (rc__ != PTRS_OK) {                                                             \
      fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc__, ptrs_last_error());         \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)
#define EXPECT(cond)                                        \
  do {                                                      \
    if (!(cond)) {                                          \
      fprintf(stderr, "expectation failed: %s\n", #cond); \
      return 1;                                             \
    }                                                       \
  } while (0)

static void node_bounds(PtrsBvhNode* n, const float* pos, const uint32_t* idx, int first, int count) {
  int t, k, a;
  for (a = 0; a < 3; ++a) {
    n->bounds_min[a] = 1e30f;
    n->bounds_max[a] = -1e30f;
  }
  for (t = first; t < first + count; ++t)
    for (k = 0; k < 3; ++k)
      for (a = 0; a < 3; ++a) {
        const float v = pos[3 * idx[3 * t + k] + a];
        if (v < n->bounds_min[a]) n->bounds_min[a] = v;
        if (v > n->bounds_max[a]) n->bounds_max[a] = v;
      }
}

int main(void) {
  /* mesh 0: floor [-2, 2]^2 at y = 0, normal +y; mesh 1: emitter [-0.5, 0.5]^2 at y = 2, facing down */
  const float pos[8 * 3] = {-2, 0, -2, 2, 0, -2, 2, 0, 2, -2, 0, 2, -0.5f, 2, -0.5f, 0.5f, 2, -0.5f, 0.5f, 2, 0.5f, -0.5f, 2, 0.5f};
  const uint32_t prim_vertex[4 * 3] = {0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7}; /* BVH order: floor, floor, light, light */
  const int32_t prim_mesh[4] = {0, 0, 1, 1}, prim_material[4] = {0, 0, 1, 1}, prim_area_light[4] = {-1, -1, 0, 1};
  PtrsMesh meshes[2];
  PtrsMaterial materials[2];
  PtrsTexture textures[3];
  PtrsLight lights[2];
  PtrsBvhNode nodes[3];
  PtrsSceneDesc d;
  PtrsScene* scene = NULL;
  PtrsFilm* film = NULL;
  PtrsMultiScene* multi = NULL;
  PtrsCamera cam;
  PtrsRenderParams params;
  PtrsStats stats;
  PtrsRay rays[3];
  PtrsHit hits[3];
  uint8_t occ[3];
  float* rgbw;
  float* rgbw2;
  int i, k, n_dev = 0;
  double sum = 0.0, sum2 = 0.0;

  memset(meshes, 0, sizeof meshes);
  meshes[0].alpha_tex = meshes[1].alpha_tex = -1;
  memset(textures, 0, sizeof textures);
  for (i = 0; i < 3; ++i) {
    textures[i].type = PTRS_TEX_CONSTANT;
    textures[i].channels = 3;
    textures[i].su = textures[i].sv = 1.f;
    textures[i].mip = -1;
  }
  textures[0].v1[0] = textures[0].v1[1] = textures[0].v1[2] = 0.5f;  /* floor kd */
  textures[2].v1[0] = textures[2].v1[1] = textures[2].v1[2] = 10.f;  /* emitter ke (textures[1] = black kd) */
  memset(materials, 0, sizeof materials);
  for (i = 0; i < 2; ++i) {
    materials[i].type = PTRS_MAT_MATTE;
    materials[i].normal_map = -1;
    for (k = 0; k < 5; ++k) materials[i].tex[k] = -1;
    materials[i].tex[0] = i; /* kd */
  }
  memset(lights, 0, sizeof lights);
  for (i = 0; i < 2; ++i) { /* one DiffuseAreaLight per emissive triangle (importer/mitsuba.rs:334-362) */
    lights[i].type = PTRS_LIGHT_AREA;
    lights[i].prim = 2 + i;
    lights[i].ke_tex = 2;
    lights[i].env = -1;
    lights[i].area = 0.5f; /* Triangle::area() of half a unit square */
  }
  memset(nodes, 0, sizeof nodes);
  node_bounds(&nodes[0], pos, prim_vertex, 0, 4);
  nodes[0].offset = 2; /* second child; the first is nodes[1] */
  nodes[0].axis = 1;
  node_bounds(&nodes[1], pos, prim_vertex, 0, 2);
  nodes[1].offset = 0;
  nodes[1].n_prims = 2;
  node_bounds(&nodes[2], pos, prim_vertex, 2, 2);
  nodes[2].offset = 2;
  nodes[2].n_prims = 2;

  memset(&d, 0, sizeof d);
  d.abi_version = PTRS_ABI_VERSION;
  d.n_nodes = 3;
  d.nodes = nodes;
  d.n_prims = 4;
  d.prim_vertex = prim_vertex;
  d.prim_mesh = prim_mesh;
  d.prim_material = prim_material;
  d.prim_area_light = prim_area_light;
  d.n_verts = 8;
  d.pos = pos;
  d.n_meshes = 2;
  d.meshes = meshes;
  d.n_materials = 2;
  d.materials = materials;
  d.n_textures = 3;
  d.textures = textures;
  d.n_lights = 2;
  d.lights = lights;

  EXPECT(ptrs_abi_version() == PTRS_ABI_VERSION);
  CHECK(ptrs_device_count(&n_dev));
  EXPECT(n_dev >= 1);
  CHECK(ptrs_scene_create(&d, &scene));

  /* RenderScene::intersect / intersect_p: down onto the floor past the emitter's edge, up into the emitter, off to the side */
  memset(rays, 0, sizeof rays);
  rays[0].o[0] = 1.f; rays[0].o[1] = 1.f; rays[0].d[1] = -1.f; rays[0].t_max = INFINITY;
  rays[1].o[1] = 1.f; rays[1].d[1] = 1.f; rays[1].t_max = INFINITY;
  rays[2].o[1] = 1.f; rays[2].d[0] = 1.f; rays[2].t_max = INFINITY;
  CHECK(ptrs_intersect(scene, rays, 3, hits));
  CHECK(ptrs_intersect_p(scene, rays, 3, occ));
  EXPECT(hits[0].prim == 0 || hits[0].prim == 1);
  EXPECT(fabsf(hits[0].t - 1.f) < 1e-6f);
  EXPECT(hits[1].prim == 2 || hits[1].prim == 3);
  EXPECT(hits[2].prim == -1);
  EXPECT(occ[0] == 1 && occ[1] == 1 && occ[2] == 0);

  /* PathIntegrator::render: camera at (0, 1, 5) looking down -z, 32 x 24, 16 spp */
  memset(&cam, 0, sizeof cam);
  cam.rot[3] = 1.f; /* identity rotation */
  cam.trans[1] = 1.f;
  cam.trans[2] = 5.f;
  cam.width = 32;
  cam.height = 24;
  { /* Camera::new (common/mod.rs:33-62) for Perspective3::new(4/3, 60 deg, 0.01, 1000) */
    const float aspect = 32.f / 24.f, fovy = 1.0471976f, zn = 0.01f, zf = 1000.f;
    const float m11 = 1.f / tanf(fovy / 2.f), m00 = m11 / aspect;
    float* m = cam.raster_to_screen;
    m[0] = 2.f / 32.f; m[3] = -1.f; m[5] = -2.f / 24.f; m[7] = 1.f; m[10] = 1.f; m[15] = 1.f;
    cam.persp[0] = m00; cam.persp[1] = m11; cam.persp[2] = (zf + zn) / (zn - zf); cam.persp[3] = 2.f * zf * zn / (zn - zf);
    /* raster_to_camera applied to the unit raster steps: unproject_point of (+-1 pixel, z = 0) */
    {
      const float kk = cam.persp[3] / cam.persp[2];
      cam.dx_camera[0] = (2.f / 32.f) * kk / m00;
      cam.dy_camera[1] = (-2.f / 24.f) * kk / m11;
    }
  }
  CHECK(ptrs_render_params_default(&params));
  params.spp = 16;
  params.max_depth = 5;
  CHECK(ptrs_film_create(cam.width, cam.height, &film));
  CHECK(ptrs_render(scene, &cam, &params, film, NULL));
  CHECK(ptrs_stats(scene, &stats));
  EXPECT(stats.camera_paths == (uint64_t)(32 + 4) * (24 + 4) * 16);
  EXPECT(stats.shadow_rays > 0 && stats.launches > 0);
  rgbw = (float*)malloc(sizeof(float) * 32 * 24 * 4);
  rgbw2 = (float*)malloc(sizeof(float) * 32 * 24 * 4);
  CHECK(ptrs_film_download(film, rgbw));
  for (i = 0; i < 32 * 24; ++i) {
    EXPECT(rgbw[4 * i + 3] > 0.f && rgbw[4 * i] >= 0.f && rgbw[4 * i] == rgbw[4 * i]);
    sum += rgbw[4 * i] / rgbw[4 * i + 3];
  }
  EXPECT(sum > 1.0); /* the lit floor is in view */

  /* the same through the multi-device entry point (one device here: no NCCL needed), and a film mismatch is refused */
  CHECK(ptrs_multi_create(&d, 1, NULL, 0, &multi));
  CHECK(ptrs_multi_render(multi, &cam, &params, rgbw2, NULL, NULL));
  for (i = 0; i < 32 * 24; ++i) sum2 += rgbw2[4 * i] / rgbw2[4 * i + 3];
  EXPECT(fabs(sum - sum2) <= 1e-3 * sum);
  cam.width = 33;
  EXPECT(ptrs_render(scene, &cam, &params, film, NULL) == PTRS_ERR_INVALID_ARGUMENT);
  EXPECT(strlen(ptrs_last_error()) > 0);

  CHECK(ptrs_multi_destroy(multi));
  CHECK(ptrs_film_destroy(film));
  CHECK(ptrs_scene_destroy(scene));
  free(rgbw);
  free(rgbw2);
  printf("c_abi_shim ok: %llu camera paths, %llu shadow rays, mean red %.4f\n", (unsigned long long)stats.camera_paths,
         (unsigned long long)stats.shadow_rays, sum / (32 * 24));
  return 0;
}
