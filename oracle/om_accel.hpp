// ORACLE — TEST INFRASTRUCTURE ONLY.  Accelerator and shape:
//   Bounds3::intersect_p_precomp      src/common/bounds.rs:190-232
//   Triangle::{intersect, intersect_p, sample, pdf_at_point, area}  src/pathtracer/shape.rs
//   GeometricPrimitive::intersect     src/pathtracer/primitive.rs:41-51
//   BVH::{intersect, intersect_p}     src/pathtracer/accelerator.rs:359-475
#pragma once
#include "om_texture.hpp"

namespace oracle {

struct TraversalCounters {
  uint64_t nodes_tested = 0, tris_tested = 0;
};

inline bool bounds_intersect_p_precomp(const PtrsBvhNode& nd, const Ray& r, Vec3 inv_dir, const bool dir_is_neg[3]) {
  const float* lo = nd.bounds_min;
  const float* hi = nd.bounds_max;
  float t_min = ((dir_is_neg[0] ? hi : lo)[0] - r.o.x) * inv_dir.x;
  float t_max = ((dir_is_neg[0] ? lo : hi)[0] - r.o.x) * inv_dir.x;
  float ty_min = ((dir_is_neg[1] ? hi : lo)[1] - r.o.y) * inv_dir.y;
  float ty_max = ((dir_is_neg[1] ? lo : hi)[1] - r.o.y) * inv_dir.y;
  const float g = 1.0f + 2.0f * gamma(3);
  t_max *= g;
  ty_max *= g;
  if (t_min > ty_max || ty_min > t_max) return false;
  if (ty_min > t_min) t_min = ty_min;
  if (ty_max < t_max) t_max = ty_max;
  float tz_min = ((dir_is_neg[2] ? hi : lo)[2] - r.o.z) * inv_dir.z;
  float tz_max = ((dir_is_neg[2] ? lo : hi)[2] - r.o.z) * inv_dir.z;
  tz_max *= g;
  if (t_min > tz_max || tz_min > t_max) return false;
  if (tz_min > t_min) t_min = tz_min;
  if (tz_max < t_max) t_max = tz_max;
  return (t_min < r.t_max) && (t_max > 0.0f);
}

struct TriVerts { Vec3 p0, p1, p2; uint32_t i0, i1, i2; };
inline TriVerts tri_verts(const Scene& sc, int prim) {
  TriVerts t;
  t.i0 = sc.d->prim_vertex[3 * prim];
  t.i1 = sc.d->prim_vertex[3 * prim + 1];
  t.i2 = sc.d->prim_vertex[3 * prim + 2];
  t.p0 = sc.pos(t.i0);
  t.p1 = sc.pos(t.i1);
  t.p2 = sc.pos(t.i2);
  return t;
}
inline void tri_uvs(const Scene& sc, int prim, const TriVerts& tv, Vec2 uv[3]) {  // shape.rs:34-48
  if (sc.mesh_of(prim).flags & PTRS_MESH_HAS_UV) {
    uv[0] = sc.uv(tv.i0);
    uv[1] = sc.uv(tv.i1);
    uv[2] = sc.uv(tv.i2);
  } else {
    uv[0] = Vec2{0.f, 0.f};
    uv[1] = Vec2{1.f, 0.f};
    uv[2] = Vec2{1.f, 1.f};
  }
}

// The watertight test shared by intersect and intersect_p (shape.rs:85-185 == :368-468).
// Returns false if rejected; otherwise b0,b1,b2,t.
inline bool tri_core(const TriVerts& tv, const Ray& r, float* b0, float* b1, float* b2, float* t_out) {
  Vec3 p0t = tv.p0 - r.o, p1t = tv.p1 - r.o, p2t = tv.p2 - r.o;
  int kz = max_dimension(vabs(r.d));
  int kx = kz + 1;
  if (kx == 3) kx = 0;
  int ky = kx + 1;
  if (ky == 3) ky = 0;
  Vec3 d = permute(r.d, kx, ky, kz);
  p0t = permute(p0t, kx, ky, kz);
  p1t = permute(p1t, kx, ky, kz);
  p2t = permute(p2t, kx, ky, kz);
  float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
  p0t.x += sx * p0t.z;
  p0t.y += sy * p0t.z;
  p1t.x += sx * p1t.z;
  p1t.y += sy * p1t.z;
  p2t.x += sx * p2t.z;
  p2t.y += sy * p2t.z;
  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
    e0 = (float)(p2typ1tx - p2txp1ty);
    double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
    e1 = (float)(p0typ2tx - p0txp2ty);
    double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
    e2 = (float)(p1typ0tx - p1txp0ty);
  }
  if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
  float det = e0 + e1 + e2;
  if (det == 0.0f) return false;
  p0t.z *= sz;
  p1t.z *= sz;
  p2t.z *= sz;
  float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < r.t_max * det)) return false;
  else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > r.t_max * det)) return false;
  float inv_det = 1.0f / det;
  *b0 = e0 * inv_det;
  *b1 = e1 * inv_det;
  *b2 = e2 * inv_det;
  float t = t_scaled * inv_det;
  float max_z_t = rmax(rmax(std::fabs(p0t.z), std::fabs(p1t.z)), std::fabs(p2t.z));
  float delta_z = gamma(3) * max_z_t;
  float max_x_t = rmax(rmax(std::fabs(p0t.x), std::fabs(p1t.x)), std::fabs(p2t.x));
  float max_y_t = rmax(rmax(std::fabs(p0t.y), std::fabs(p1t.y)), std::fabs(p2t.y));
  float delta_x = gamma(5) * (max_x_t + max_z_t);
  float delta_y = gamma(5) * (max_y_t + max_z_t);
  float delta_e = 2.0f * (gamma(2) * max_x_t * max_y_t + delta_y * max_x_t + delta_x * max_y_t);
  float max_e = rmax(rmax(std::fabs(e0), std::fabs(e1)), std::fabs(e2));
  float delta_t = 3.0f * (gamma(3) * max_e * max_z_t + delta_e * max_z_t + delta_z * max_e) * std::fabs(inv_det);
  if (t <= delta_t) return false;
  *t_out = t;
  return true;
}

// dpdu / dpdv block (shape.rs:187-215 == :472-500); false = degenerate triangle
inline bool tri_partials(const TriVerts& tv, const Vec2 uv[3], Vec3* dpdu, Vec3* dpdv) {
  *dpdu = V(0, 0, 0);
  *dpdv = V(0, 0, 0);
  float duv02[2] = {uv[0].x - uv[2].x, uv[0].y - uv[2].y}, duv12[2] = {uv[1].x - uv[2].x, uv[1].y - uv[2].y};
  Vec3 dp02 = tv.p0 - tv.p2, dp12 = tv.p1 - tv.p2;
  float determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
  bool degenerate_uv = std::fabs(determinant) < 1e-8f;
  if (!degenerate_uv) {
    float invdet = 1.0f / determinant;
    *dpdu = (duv12[1] * dp02 - duv02[1] * dp12) * invdet;
    *dpdv = (-duv12[0] * dp02 + duv02[0] * dp12) * invdet;
  }
  if (degenerate_uv || norm_squared(cross(*dpdu, *dpdv)) == 0.0f) {
    Vec3 ng = cross(tv.p2 - tv.p0, tv.p1 - tv.p0);
    if (norm_squared(ng) == 0.0f) return false;
    coordinate_system(normalize(ng), dpdu, dpdv);
  }
  return true;
}

// Triangle::intersect, shape.rs:74-360.  `barys` (optional) receives b0,b1,b2.
inline bool triangle_intersect(const Scene& sc, int prim, const Ray& r, float* t_hit, SurfaceInteraction* isect, float* barys = nullptr) {
  const TriVerts tv = tri_verts(sc, prim);
  float b0, b1, b2, t;
  if (!tri_core(tv, r, &b0, &b1, &b2, &t)) return false;
  Vec2 uv[3];
  tri_uvs(sc, prim, tv, uv);
  Vec3 dpdu, dpdv;
  if (!tri_partials(tv, uv, &dpdu, &dpdv)) return false;
  const Vec3 p0 = tv.p0, p1 = tv.p1, p2 = tv.p2;
  float x_abs_sum = std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x) + std::fabs(b2 * p2.x);
  float y_abs_sum = std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y) + std::fabs(b2 * p2.y);
  float z_abs_sum = std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z) + std::fabs(b2 * p2.z);
  Vec3 p_error = gamma(7) * V(x_abs_sum, y_abs_sum, z_abs_sum);
  Vec3 p_hit = b0 * p0 + b1 * p1 + b2 * p2;
  Vec2 uv_hit{b0 * uv[0].x + b1 * uv[1].x + b2 * uv[2].x, b0 * uv[0].y + b1 * uv[1].y + b2 * uv[2].y};
  const PtrsMesh& mesh = sc.mesh_of(prim);
  if (mesh.alpha_tex >= 0) {  // shape.rs:228-244
    SurfaceInteraction local = SurfaceInteraction::make(p_hit, V(0, 0, 0), uv_hit, -r.d, dpdu, dpdv);
    if (tex_f32(sc, mesh.alpha_tex, local) == 0.0f) return false;
  }
  *isect = SurfaceInteraction::make(p_hit, p_error, uv_hit, -r.d, dpdu, dpdv);
  Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
  isect->general.n = normalize(cross(dp02, dp12));
  isect->shading.n = isect->general.n;
  // reverse_orientation ^ transform_swaps_handedness is always false (triangles_from_mesh(.., false))
  const bool has_n = mesh.flags & PTRS_MESH_HAS_NORMAL, has_s = mesh.flags & PTRS_MESH_HAS_TANGENT;
  if (has_n || has_s) {
    Vec3 ns;
    if (has_n) {
      ns = b0 * sc.normal(tv.i0) + b1 * sc.normal(tv.i1) + b2 * sc.normal(tv.i2);
      if (norm_squared(ns) > 0.0f) ns = normalize(ns);
      else ns = isect->general.n;
    } else {
      ns = isect->general.n;
    }
    Vec3 ss;
    if (has_s) {
      ss = b0 * sc.tangent(tv.i0) + b1 * sc.tangent(tv.i1) + b2 * sc.tangent(tv.i2);
      if (norm_squared(ss) > 0.0f) ss = normalize(ss);
      else ss = normalize(isect->dpdu);
    } else {
      ss = normalize(isect->dpdu);
    }
    Vec3 ts = cross(ss, ns);
    if (norm_squared(ts) > 0.0f) {
      ts = normalize(ts);
      ss = cross(ts, ns);
    } else {
      coordinate_system(ns, &ss, &ts);
    }
    Vec3 dndu = V(0, 0, 0), dndv = V(0, 0, 0);
    if (has_n) {
      float duv02[2] = {uv[0].x - uv[2].x, uv[0].y - uv[2].y}, duv12[2] = {uv[1].x - uv[2].x, uv[1].y - uv[2].y};
      Vec3 n0 = sc.normal(tv.i0), n1 = sc.normal(tv.i1), n2 = sc.normal(tv.i2);
      Vec3 dn1 = n0 - n2, dn2 = n1 - n2;
      float determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
      bool degenerate_uv = std::fabs(determinant) < 1e-8f;
      if (degenerate_uv) {
        Vec3 dn = cross(n2 - n0, n1 - n0);
        if (norm_squared(dn) == 0.0f) {
          dndu = V(0, 0, 0);
          dndv = V(0, 0, 0);
        } else {
          coordinate_system(dn, &dndu, &dndv);
        }
      } else {
        float inv_det = 1.0f / determinant;
        dndu = (duv12[1] * dn1 - duv02[1] * dn2) * inv_det;
        dndv = (-duv12[0] * dn1 + duv02[0] * dn2) * inv_det;
      }
    }
    isect->set_shading_geometry(ss, ts, dndu, dndv, true);
  }
  *t_hit = t;
  if (barys) {
    barys[0] = b0;
    barys[1] = b1;
    barys[2] = b2;
  }
  return true;
}

// Triangle::intersect_p, shape.rs:362-524
inline bool triangle_intersect_p(const Scene& sc, int prim, const Ray& r) {
  const TriVerts tv = tri_verts(sc, prim);
  float b0, b1, b2, t;
  if (!tri_core(tv, r, &b0, &b1, &b2, &t)) return false;
  const PtrsMesh& mesh = sc.mesh_of(prim);
  if (mesh.alpha_tex >= 0) {
    Vec2 uv[3];
    tri_uvs(sc, prim, tv, uv);
    Vec3 dpdu, dpdv;
    if (!tri_partials(tv, uv, &dpdu, &dpdv)) return false;
    Vec3 p_hit = b0 * tv.p0 + b1 * tv.p1 + b2 * tv.p2;
    Vec2 uv_hit{b0 * uv[0].x + b1 * uv[1].x + b2 * uv[2].x, b0 * uv[0].y + b1 * uv[1].y + b2 * uv[2].y};
    SurfaceInteraction local = SurfaceInteraction::make(p_hit, V(0, 0, 0), uv_hit, -r.d, dpdu, dpdv);
    if (tex_f32(sc, mesh.alpha_tex, local) == 0.0f) return false;
  }
  return true;
}

inline float triangle_area(const Scene& sc, int prim) {  // shape.rs:533-539
  const TriVerts tv = tri_verts(sc, prim);
  return 0.5f * norm(cross(tv.p1 - tv.p0, tv.p2 - tv.p0));
}

// Triangle::sample, shape.rs:541-578 (uniform_sample_triangle :14-17)
inline SurfaceInteraction triangle_sample(const Scene& sc, int prim, Vec2 u) {
  float su0 = std::sqrt(u.x);
  float b[2] = {1.0f - su0, u.y * su0};
  const TriVerts tv = tri_verts(sc, prim);
  SurfaceInteraction si;
  Interaction it;
  it.p = (b[0] * tv.p0) + (b[1] * tv.p1) + (1.0f - b[0] - b[1]) * tv.p2;
  it.n = normalize(cross(tv.p1 - tv.p0, tv.p2 - tv.p0));
  if (sc.mesh_of(prim).flags & PTRS_MESH_HAS_NORMAL) {
    Vec3 ns = (b[0] * sc.normal(tv.i0)) + (b[1] * sc.normal(tv.i1)) + (1.0f - b[0] - b[1]) * sc.normal(tv.i2);
    it.n = face_forward(it.n, ns);
  }
  Vec3 p_abs_sum = vabs(b[0] * tv.p0) + vabs(b[1] * tv.p1) + vabs((1.0f - b[0] - b[1]) * tv.p2);
  it.p_error = gamma(6) * p_abs_sum;
  Vec2 uv[3];
  tri_uvs(sc, prim, tv, uv);
  float w2 = 1.0f - b[0] - b[1];
  si.general = it;
  si.uv = Vec2{b[0] * uv[0].x + b[1] * uv[1].x + w2 * uv[2].x, b[0] * uv[0].y + b[1] * uv[1].y + w2 * uv[2].y};
  return si;
}

// Triangle::pdf_at_point, shape.rs:62-72
inline float triangle_pdf_at_point(const Scene& sc, int prim, const Interaction& reference, Vec3 wi, float area) {
  Ray ray = reference.spawn_ray(wi);
  float t_hit = 0.0f;
  SurfaceInteraction isect_light;
  if (!triangle_intersect(sc, prim, ray, &t_hit, &isect_light)) return 0.0f;
  return norm_squared(reference.p - isect_light.general.p) / (std::fabs(dot(isect_light.general.n, -wi)) * area);
}

// BVH::intersect, accelerator.rs:359-417 (+ GeometricPrimitive::intersect, primitive.rs:41-51)
inline bool bvh_intersect(const Scene& sc, Ray* r, SurfaceInteraction* isect, float* barys = nullptr, TraversalCounters* ctr = nullptr) {
  if (sc.d->n_nodes == 0) return false;
  bool hit = false;
  Vec3 inv_dir = V(1.0f / r->d.x, 1.0f / r->d.y, 1.0f / r->d.z);
  const bool dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
  size_t to_visit_offset = 0, curr = 0;
  size_t nodes_to_visit[64];
  for (;;) {
    const PtrsBvhNode& node = sc.d->nodes[curr];
    if (ctr) ctr->nodes_tested++;
    if (bounds_intersect_p_precomp(node, *r, inv_dir, dir_is_neg)) {
      if (node.n_prims > 0) {
        for (uint32_t i = 0; i < node.n_prims; ++i) {
          const int prim = (int)(node.offset + i);
          float t_hit = 0.0f;
          if (ctr) ctr->tris_tested++;
          if (triangle_intersect(sc, prim, *r, &t_hit, isect, barys)) {
            r->t_max = t_hit;
            isect->primitive = prim;
            hit = true;
          }
        }
        if (to_visit_offset == 0) break;
        curr = nodes_to_visit[--to_visit_offset];
      } else {
        if (dir_is_neg[node.axis]) {
          nodes_to_visit[to_visit_offset++] = curr + 1;
          curr = node.offset;
        } else {
          nodes_to_visit[to_visit_offset++] = node.offset;
          curr = curr + 1;
        }
      }
    } else {
      if (to_visit_offset == 0) break;
      curr = nodes_to_visit[--to_visit_offset];
    }
  }
  return hit;
}

// BVH::intersect_p, accelerator.rs:419-475
inline bool bvh_intersect_p(const Scene& sc, const Ray& r, TraversalCounters* ctr = nullptr) {
  if (sc.d->n_nodes == 0) return false;
  Vec3 inv_dir = V(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
  const bool dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
  size_t to_visit_offset = 0, curr = 0;
  size_t nodes_to_visit[64];
  for (;;) {
    const PtrsBvhNode& node = sc.d->nodes[curr];
    if (ctr) ctr->nodes_tested++;
    if (bounds_intersect_p_precomp(node, r, inv_dir, dir_is_neg)) {
      if (node.n_prims > 0) {
        for (uint32_t i = 0; i < node.n_prims; ++i) {
          if (ctr) ctr->tris_tested++;
          if (triangle_intersect_p(sc, (int)(node.offset + i), r)) return true;
        }
        if (to_visit_offset == 0) break;
        curr = nodes_to_visit[--to_visit_offset];
      } else {
        if (dir_is_neg[node.axis]) {
          nodes_to_visit[to_visit_offset++] = curr + 1;
          curr = node.offset;
        } else {
          nodes_to_visit[to_visit_offset++] = node.offset;
          curr = curr + 1;
        }
      }
    } else {
      if (to_visit_offset == 0) break;
      curr = nodes_to_visit[--to_visit_offset];
    }
  }
  return false;
}

}  // namespace oracle
