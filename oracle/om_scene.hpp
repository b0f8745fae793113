// ORACLE — TEST INFRASTRUCTURE ONLY.  Interaction records (src/pathtracer/interaction.rs) and a
// read-only view of the flat scene (include/ptrs_b200.h) that plays the role of RenderScene.
#pragma once
#include "../include/ptrs_b200.h"
#include "om_math.hpp"

namespace oracle {

struct Ray {  // src/common/ray.rs:2-6
  Vec3 o, d;
  float t_max;
};
struct RayDifferential {  // ray.rs:8-36
  Ray ray;
  bool has_differentials = false;
  Vec3 rx_origin{0, 0, 0}, ry_origin{0, 0, 0}, rx_direction{0, 0, 0}, ry_direction{0, 0, 0};
  static RayDifferential from_ray(const Ray& r) {
    RayDifferential rd;
    rd.ray = r;
    return rd;
  }
  void scale_differentials(float s) {  // ray.rs:30-35
    rx_origin = ray.o + (rx_origin - ray.o) * s;
    ry_origin = ray.o + (ry_origin - ray.o) * s;
    rx_direction = ray.d + (rx_direction - ray.d) * s;
    ry_direction = ray.d + (ry_direction - ray.d) * s;
  }
};

constexpr float SHADOW_EPSILON = 0.0001f;  // interaction.rs:29

struct Interaction {  // interaction.rs:9-27
  Vec3 p{0, 0, 0}, p_error{0, 0, 0}, wo{0, 0, 0}, n{0, 0, 0};
  Ray spawn_ray(Vec3 d) const {  // interaction.rs:32-39
    Vec3 o = offset_ray_origin(p, p_error, n, d);
    return Ray{o, d, std::numeric_limits<float>::infinity()};
  }
  Ray spawn_ray_to_it(const Interaction& it2) const {  // interaction.rs:50-59
    Vec3 origin = offset_ray_origin(p, p_error, n, it2.p - p);
    Vec3 target = offset_ray_origin(it2.p, it2.p_error, it2.n, origin - it2.p);
    Vec3 d = target - origin;
    return Ray{origin, d, 1.0f - SHADOW_EPSILON};
  }
};

struct Shading {  // interaction.rs:62-69
  Vec3 n{0, 0, 0}, dpdu{0, 0, 0}, dpdv{0, 0, 0}, dndu{0, 0, 0}, dndv{0, 0, 0};
};

struct SurfaceInteraction {  // SurfaceMediumInteraction, interaction.rs:83-101 (bsdf lives beside it)
  Interaction general;
  Vec2 uv{0, 0};
  Vec3 dpdu{0, 0, 0}, dpdv{0, 0, 0}, dndu{0, 0, 0}, dndv{0, 0, 0};
  Shading shading;
  int32_t primitive = -1;  // index into the BVH-ordered primitive array
  Vec3 dpdx{0, 0, 0}, dpdy{0, 0, 0};
  float dudx = 0, dvdx = 0, dudy = 0, dvdy = 0;

  // SurfaceMediumInteraction::new, interaction.rs:133-172
  static SurfaceInteraction make(Vec3 p, Vec3 p_error, Vec2 uv, Vec3 wo, Vec3 dpdu, Vec3 dpdv) {
    SurfaceInteraction s;
    Vec3 n = normalize(cross(dpdu, dpdv));
    s.shading.n = n;
    s.shading.dpdu = dpdu;
    s.shading.dpdv = dpdv;
    s.general.p = p;
    s.general.p_error = p_error;
    s.general.wo = wo;
    s.general.n = n;
    s.uv = uv;
    s.dpdu = dpdu;
    s.dpdv = dpdv;
    return s;
  }
  // interaction.rs:194-214
  void set_shading_geometry(Vec3 dpdus, Vec3 dpdvs, Vec3 dndus, Vec3 dndvs, bool orientation_is_authoritative) {
    shading.n = normalize(cross(dpdus, dpdvs));
    if (orientation_is_authoritative) general.n = face_forward(general.n, shading.n);
    else shading.n = face_forward(shading.n, general.n);
    shading.dpdu = dpdus;
    shading.dpdv = dpdvs;
    shading.dndu = dndus;
    shading.dndv = dndvs;
  }
  // interaction.rs:216-281
  bool compute_differentials(const RayDifferential& ray) {
    if (!ray.has_differentials) return false;
    const Vec3 n = general.n, p = general.p;
    float d = dot(n, p);
    float tx = -(dot(n, ray.rx_origin) - d) / dot(n, ray.rx_direction);
    if (std::isinf(tx) || tx != tx) return false;
    Vec3 px = ray.rx_origin + tx * ray.rx_direction;
    float ty = -(dot(n, ray.ry_origin) - d) / dot(n, ray.ry_direction);
    if (std::isinf(ty) || ty != ty) return false;
    Vec3 py = ray.ry_origin + ty * ray.ry_direction;
    dpdx = px - p;
    dpdy = py - p;
    int dim[2];
    // sic: the reference compares n.x with n.y twice (interaction.rs:241)
    if (std::fabs(n.x) > std::fabs(n.y) && std::fabs(n.x) > std::fabs(n.y)) {
      dim[0] = 1;
      dim[1] = 2;
    } else if (std::fabs(n.y) > std::fabs(n.z)) {
      dim[0] = 0;
      dim[1] = 2;
    } else {
      dim[0] = 0;
      dim[1] = 1;
    }
    const float a[2][2] = {{dpdu[dim[0]], dpdv[dim[0]]}, {dpdu[dim[1]], dpdv[dim[1]]}};
    const float bx[2] = {px[dim[0]] - p[dim[0]], px[dim[1]] - p[dim[1]]};
    const float by[2] = {py[dim[0]] - p[dim[0]], py[dim[1]] - p[dim[1]]};
    if (!solve_linear_system_2x2(a, bx, &dudx, &dvdx)) dudx = dvdx = 0.0f;
    if (!solve_linear_system_2x2(a, by, &dudy, &dvdy)) dudy = dvdy = 0.0f;
    return true;
  }
};

struct Scene {
  const PtrsSceneDesc* d;
  Vec3 pos(uint32_t v) const { return {d->pos[3 * v], d->pos[3 * v + 1], d->pos[3 * v + 2]}; }
  Vec3 normal(uint32_t v) const { return {d->normal[3 * v], d->normal[3 * v + 1], d->normal[3 * v + 2]}; }
  Vec3 tangent(uint32_t v) const { return {d->tangent[3 * v], d->tangent[3 * v + 1], d->tangent[3 * v + 2]}; }
  Vec2 uv(uint32_t v) const { return {d->uv[2 * v], d->uv[2 * v + 1]}; }
  const PtrsMesh& mesh_of(int prim) const { return d->meshes[d->prim_mesh[prim]]; }
};

}  // namespace oracle
