// ORACLE — TEST INFRASTRUCTURE ONLY.  Camera ray generation (src/pathtracer/mod.rs:59-81), the path
// integrator (src/pathtracer/integrator.rs:392-503, 536-642) and the film (src/common/film.rs).
#pragma once
#include <omp.h>

#include <vector>

#include "om_shading.hpp"

namespace oracle {

// UnitQuaternion * Vector3 (nalgebra): t = 2 (q.v x v); v' = t * w + q.v x t + v
inline Vec3 quat_rotate(const float q[4], Vec3 v) {
  Vec3 qv = V(q[0], q[1], q[2]);
  Vec3 t = cross(qv, v) * 2.0f;
  Vec3 c = cross(qv, t);
  return t * q[3] + c + v;
}

// Camera::generate_ray_differential, pathtracer/mod.rs:59-81
inline RayDifferential generate_ray_differential(const PtrsCamera& cam, Vec2 p_film) {
  const float* m = cam.raster_to_screen;
  // raster_to_screen * Point3(x, y, 0): Affine3 (no normaliser)
  float sx = (m[0] * p_film.x + m[1] * p_film.y) + m[2] * 0.0f + m[3];
  float sy = (m[4] * p_film.x + m[5] * p_film.y) + m[6] * 0.0f + m[7];
  float sz = (m[8] * p_film.x + m[9] * p_film.y) + m[10] * 0.0f + m[11];
  // Perspective3::unproject_point
  float inverse_denom = cam.persp[3] / (sz + cam.persp[2]);
  Vec3 p_camera = V(sx * inverse_denom / cam.persp[0], sy * inverse_denom / cam.persp[1], -inverse_denom);
  Vec3 world_orig = V(cam.trans[0], cam.trans[1], cam.trans[2]);  // cam_to_world * origin
  Vec3 world_dir = quat_rotate(cam.rot, p_camera);
  Vec3 dxc = V(cam.dx_camera[0], cam.dx_camera[1], cam.dx_camera[2]), dyc = V(cam.dy_camera[0], cam.dy_camera[1], cam.dy_camera[2]);
  Vec3 rx_world_dir = quat_rotate(cam.rot, p_camera + dxc);
  Vec3 ry_world_dir = quat_rotate(cam.rot, p_camera + dyc);
  RayDifferential rd;
  rd.ray = Ray{world_orig, normalize(world_dir), std::numeric_limits<float>::infinity()};
  rd.has_differentials = true;
  rd.rx_origin = world_orig;
  rd.ry_origin = world_orig;
  rd.rx_direction = normalize(rx_world_dir);
  rd.ry_direction = normalize(ry_world_dir);
  return rd;
}

struct IntegratorParams {
  int32_t max_depth = 15;
  float rr_threshold = 1.0f;
  int32_t rr_start_depth = 3;
  bool rr_enable = true;
};

// PathIntegrator::li, integrator.rs:392-503
inline Spectrum path_li(const Scene& sc, const IntegratorParams& P, const RayDifferential& ray_in, SobolSampler* sampler, RayCounters* rc) {
  Spectrum l = S(0.0f), beta = S(1.0f);
  RayDifferential ray = ray_in;
  bool specular_bounce = false;
  int32_t bounces = 0;
  float eta_scale = 1.0f;
  for (;;) {
    SurfaceInteraction isect;
    if (rc) rc->extension++;
    bool found_intersection = bvh_intersect(sc, &ray.ray, &isect, nullptr, rc ? &rc->trav : nullptr);
    if (bounces == 0 || specular_bounce) {
      if (found_intersection) {
        l += beta * isect_le(sc, isect, -ray.ray.d);
      } else {
        for (uint32_t i = 0; i < sc.d->n_infinite_lights; ++i) l += beta * light_le(sc, sc.d->lights[sc.d->infinite_lights[i]], ray.ray);
      }
    }
    if (!found_intersection || bounces >= P.max_depth) break;
    // SurfaceMediumInteraction::compute_scattering_functions, interaction.rs:283-295
    if (!isect.compute_differentials(ray)) {
      isect.dudx = isect.dvdx = isect.dudy = isect.dvdy = 0.0f;
      isect.dpdx = V(0, 0, 0);
      isect.dpdy = V(0, 0, 0);
    }
    BSDF bsdf;
    if (!compute_scattering_functions(sc, &isect, &bsdf)) {
      ray = RayDifferential::from_ray(isect.general.spawn_ray(ray.ray.d));
      bounces -= 1;  // sic: the loop increment is skipped by `continue` (integrator.rs:434-439)
      continue;
    }
    if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) {
      Spectrum ld = beta * uniform_sample_one_light(sc, isect, bsdf, sampler, rc);
      l += ld;
    }
    Vec3 wo = -ray.ray.d, wi = V(0, 0, 0);
    float pdf = 0.0f;
    uint32_t flags = 0;
    Spectrum f = bsdf.sample_f(wo, &wi, sampler->get_2d(), &pdf, BSDF_ALL, &flags);
    if (is_black(f) || pdf == 0.0f) break;
    beta *= f * std::fabs(dot(wi, isect.shading.n)) / pdf;
    specular_bounce = (flags & BSDF_SPECULAR) != 0;
    if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
      float eta = bsdf.eta;
      eta_scale *= dot(wo, isect.general.n) > 0.0f ? eta * eta : 1.0f / (eta * eta);
    }
    ray = RayDifferential::from_ray(isect.general.spawn_ray(wi));
    if (P.rr_enable) {
      Spectrum rr_beta = beta * eta_scale;
      if (max_component(rr_beta) < P.rr_threshold && bounces > P.rr_start_depth) {
        float q = rmax(0.05f, 1.0f - max_component(rr_beta));
        if (sampler->get_1d() < q) break;
        beta /= 1.0f - q;
      }
    }
    bounces += 1;
  }
  return l;
}

// Film, film.rs ---------------------------------------------------------------------------------
struct Bounds2i { int32_t x0, y0, x1, y1; };

inline Bounds2i film_sample_bounds(int width, int height, const float radius[2]) {  // film.rs:174-185
  return Bounds2i{f2i(std::floor(0.5f - radius[0])), f2i(std::floor(0.5f - radius[1])),
                  f2i(std::ceil((float)width - 0.5f + radius[0])), f2i(std::ceil((float)height - 0.5f + radius[1]))};
}

struct FilmTile {  // film.rs:23-111
  Bounds2i pb;
  std::vector<float> px;  // rgbw per pixel
  float radius[2], inv_radius[2];
  const float* table;
  void add_sample(Vec2 p_film, Spectrum l) {  // film.rs:60-106
    float dx = p_film.x - 0.5f, dy = p_film.y - 0.5f;
    int p0x = f2i(std::ceil(dx - radius[0])), p0y = f2i(std::ceil(dy - radius[1]));
    int p1x = f2i(std::floor(dx + radius[0]) + 1.0f), p1y = f2i(std::floor(dy + radius[1]) + 1.0f);
    p0x = p0x > pb.x0 ? p0x : pb.x0;
    p0y = p0y > pb.y0 ? p0y : pb.y0;
    p1x = p1x < pb.x1 ? p1x : pb.x1;
    p1y = p1y < pb.y1 ? p1y : pb.y1;
    int ifx[32], ify[32];
    for (int x = p0x; x < p1x; ++x) {
      float fx = std::fabs(((float)x - dx) * inv_radius[0] * 16.0f);
      int v = f2i(std::floor(fx));
      ifx[x - p0x] = v < 15 ? v : 15;
    }
    for (int y = p0y; y < p1y; ++y) {
      float fy = std::fabs(((float)y - dy) * inv_radius[1] * 16.0f);
      int v = f2i(std::floor(fy));
      ify[y - p0y] = v < 15 ? v : 15;
    }
    const int w = pb.x1 - pb.x0;
    for (int y = p0y; y < p1y; ++y)
      for (int x = p0x; x < p1x; ++x) {
        float fw = table[ify[y - p0y] * 16 + ifx[x - p0x]];
        float* p = &px[((size_t)(y - pb.y0) * w + (x - pb.x0)) * 4];
        p[0] += l.r * fw;
        p[1] += l.g * fw;
        p[2] += l.b * fw;
        p[3] += fw;
      }
  }
};

struct RenderStats {
  uint64_t camera_paths = 0, extension = 0, shadow = 0, mis = 0, nodes = 0, tris = 0;
};

// PathIntegrator::render, integrator.rs:536-642.  film_rgbw: W*H*4 floats, accumulated into.
// tile_begin/tile_end restrict the run to a range of tiles (bounded CPU-baseline samples);
// tiles are merged in tile order, one fixed instance of the reference's nondeterministic merge.
inline RenderStats render(const SobolTables& T, const PtrsSceneDesc* desc, const PtrsCamera& cam, const PtrsRenderParams& rp,
                          float* film_rgbw, int n_threads, int64_t tile_begin, int64_t tile_end, int64_t tile_stride) {
  Scene sc{desc};
  IntegratorParams P;
  P.max_depth = rp.max_depth;
  P.rr_threshold = rp.rr_threshold;
  P.rr_start_depth = rp.rr_start_depth;
  P.rr_enable = rp.rr_enable != 0;
  const Bounds2i sb = film_sample_bounds(cam.width, cam.height, rp.filter_radius);
  const int32_t sbv[4] = {sb.x0, sb.y0, sb.x1, sb.y1};
  SobolSampler proto;
  proto.configure(&T, (size_t)rp.spp, sbv);
  const int spp = (int)proto.samples_per_pixel;
  const int s_begin = rp.sample_begin > 0 ? rp.sample_begin : 0;
  const int s_end = rp.sample_end > 0 ? (rp.sample_end < spp ? rp.sample_end : spp) : spp;
  const int stride = rp.sample_stride > 0 ? rp.sample_stride : 1, phase = rp.sample_phase;
  const int TILE = 16;
  const int ntx = (sb.x1 - sb.x0 + TILE - 1) / TILE, nty = (sb.y1 - sb.y0 + TILE - 1) / TILE;
  const int64_t n_tiles = (int64_t)ntx * nty;
  if (tile_end <= 0 || tile_end > n_tiles) tile_end = n_tiles;
  if (tile_begin < 0) tile_begin = 0;
  if (tile_stride < 1) tile_stride = 1;
  const int64_t n_run = tile_end > tile_begin ? (tile_end - tile_begin + tile_stride - 1) / tile_stride : 0;
  std::vector<FilmTile> tiles((size_t)n_run);
  const float scale = 1.0f / std::sqrt((float)spp);
  RenderStats total;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel num_threads(n_threads)
  {
    RenderStats st;
#pragma omp for schedule(dynamic, 1)
    for (int64_t tk = 0; tk < n_run; ++tk) {
      const int64_t ti = tile_begin + tk * tile_stride;
      // render_tile_vec order: (0..num_tiles.x).cartesian_product(0..num_tiles.y) -> x outer
      const int tx = (int)(ti / nty), ty = (int)(ti % nty);
      SobolSampler sampler = proto;
      const int x0 = sb.x0 + tx * TILE, x1 = std::min(x0 + TILE, sb.x1);
      const int y0 = sb.y0 + ty * TILE, y1 = std::min(y0 + TILE, sb.y1);
      FilmTile& ft = tiles[(size_t)tk];
      // Film::get_film_tile, film.rs:193-211
      Bounds2i b{f2i(std::ceil((float)x0 - 0.5f - rp.filter_radius[0])), f2i(std::ceil((float)y0 - 0.5f - rp.filter_radius[1])),
                 f2i(std::floor((float)x1 - 0.5f + rp.filter_radius[0])) + 1, f2i(std::floor((float)y1 - 0.5f + rp.filter_radius[1])) + 1};
      b.x0 = std::max(b.x0, 0);
      b.y0 = std::max(b.y0, 0);
      b.x1 = std::min(b.x1, cam.width);
      b.y1 = std::min(b.y1, cam.height);
      ft.pb = b;
      ft.px.assign((size_t)std::max(0, b.x1 - b.x0) * std::max(0, b.y1 - b.y0) * 4, 0.0f);
      ft.radius[0] = rp.filter_radius[0];
      ft.radius[1] = rp.filter_radius[1];
      ft.inv_radius[0] = 1.f / rp.filter_radius[0];
      ft.inv_radius[1] = 1.f / rp.filter_radius[1];
      ft.table = rp.filter_table;
      for (int x = x0; x < x1; ++x)
        for (int y = y0; y < y1; ++y) {
          sampler.start_pixel(x, y);
          for (int s = s_begin; s < s_end; ++s) {
            if (s % stride != phase) continue;
            sampler.set_sample((size_t)s);
            Vec2 p_film = sampler.get_camera_sample();
            RayDifferential ray = generate_ray_differential(cam, p_film);
            ray.scale_differentials(scale);
            RayCounters rc;
            Spectrum l = path_li(sc, P, ray, &sampler, &rc);
            st.camera_paths++;
            st.extension += rc.extension;
            st.shadow += rc.shadow;
            st.mis += rc.mis;
            st.nodes += rc.trav.nodes_tested;
            st.tris += rc.trav.tris_tested;
            ft.add_sample(p_film, l);
          }
        }
    }
#pragma omp critical
    {
      total.camera_paths += st.camera_paths;
      total.extension += st.extension;
      total.shadow += st.shadow;
      total.mis += st.mis;
      total.nodes += st.nodes;
      total.tris += st.tris;
    }
  }
  for (const FilmTile& ft : tiles) {  // Film::merge_film_tile, film.rs:213-228
    const int w = ft.pb.x1 - ft.pb.x0;
    for (int x = ft.pb.x0; x < ft.pb.x1; ++x)
      for (int y = ft.pb.y0; y < ft.pb.y1; ++y) {
        const float* p = &ft.px[((size_t)(y - ft.pb.y0) * w + (x - ft.pb.x0)) * 4];
        float* q = &film_rgbw[((size_t)y * cam.width + x) * 4];
        q[0] += p[0];
        q[1] += p[1];
        q[2] += p[2];
        q[3] += p[3];
      }
  }
  return total;
}

}  // namespace oracle
