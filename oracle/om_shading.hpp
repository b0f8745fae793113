// ORACLE — TEST INFRASTRUCTURE ONLY.  Materials (src/pathtracer/material/*.rs), lights
// (src/pathtracer/light.rs, sampling.rs:128-230) and the integrator's direct-lighting estimator
// (src/pathtracer/integrator.rs:23-217).
#pragma once
#include "om_accel.hpp"
#include "om_bxdf.hpp"
#include "om_sobol.hpp"

namespace oracle {

inline float sqr(float x) { return x * x; }
inline float schlick_r0_from_eta(float eta) { return sqr(eta - 1.0f) / sqr(eta + 1.0f); }  // material/mod.rs:100-102

// normal_mapping, material/mod.rs:39-79
inline void normal_mapping(const Scene& sc, int tex, SurfaceInteraction* si) {
  const Vec3 c0 = si->shading.dpdu, c1 = si->shading.dpdv, c2 = si->shading.n;
  Vec3 texture_n = normalize(tex_vec3(sc, tex, *si));
  // tbn * texture_n, column-major gemv: (c0 * x + c1 * y) + c2 * z
  Vec3 ns = normalize(V(c0.x * texture_n.x + c1.x * texture_n.y + c2.x * texture_n.z,
                        c0.y * texture_n.x + c1.y * texture_n.y + c2.y * texture_n.z,
                        c0.z * texture_n.x + c1.z * texture_n.y + c2.z * texture_n.z));
  Vec3 ss = si->shading.dpdu;
  Vec3 ts = cross(ss, ns);
  if (norm_squared(ts) > 0.0f) {
    ts = normalize(ts);
    ss = cross(ts, ns);
  } else {
    coordinate_system(ns, &ss, &ts);
  }
  si->shading.n = ns;
  si->shading.dpdu = ss;
  si->shading.dpdv = ts;
}

// Material::compute_scattering_functions; returns false when the material leaves si.bsdf = None
inline bool compute_scattering_functions(const Scene& sc, SurfaceInteraction* si, BSDF* out) {
  const PtrsMaterial& m = sc.d->materials[sc.d->prim_material[si->primitive]];
  if (m.normal_map >= 0) normal_mapping(sc, m.normal_map, si);  // NormalMaterial, mod.rs:130-135
  switch (m.type) {
    case PTRS_MAT_MATTE: {  // mod.rs:155-167
      BSDF bsdf = BSDF::make(*si, 1.0f);
      BxDF b;
      b.kind = BxDF::Lambertian;
      b.r = tex_spectrum(sc, m.tex[0], *si);
      bsdf.add(b);
      *out = bsdf;
      return true;
    }
    case PTRS_MAT_MIRROR: {  // mod.rs:180-195
      BSDF bsdf = BSDF::make(*si, 1.0f);
      BxDF b;
      b.kind = BxDF::SpecularReflection;
      b.r = S(1.0f);
      b.fresnel.kind = Fresnel::NoOp;
      bsdf.add(b);
      *out = bsdf;
      return true;
    }
    case PTRS_MAT_GLASS: {  // mod.rs:216-255
      float eta = tex_f32(sc, m.tex[2], *si);
      Spectrum r = tex_spectrum(sc, m.tex[0], *si), t = tex_spectrum(sc, m.tex[1], *si);
      BSDF bsdf = BSDF::make(*si, eta);
      if (is_black(r) && is_black(t)) return false;
      BxDF b;
      b.kind = BxDF::FresnelSpecular;
      b.r = r;
      b.t = t;
      b.eta_a = 1.0f;
      b.eta_b = eta;
      bsdf.add(b);
      *out = bsdf;
      return true;
    }
    case PTRS_MAT_METAL: {  // metal.rs:49-93
      BSDF bsdf = BSDF::make(*si, 1.0f);
      float u_rough = tex_f32(sc, m.tex[3], *si), v_rough = tex_f32(sc, m.tex[4], *si);
      if (m.remap_roughness) {
        u_rough = roughness_to_alpha(u_rough);
        v_rough = roughness_to_alpha(v_rough);
      }
      BxDF b;
      b.kind = BxDF::MicrofacetReflection;
      b.r = tex_spectrum(sc, m.tex[2], *si);
      b.dist = Distribution::make(u_rough, v_rough, false);
      b.fresnel.kind = Fresnel::Conductor;
      b.fresnel.c_eta_i = S(1.f);
      b.fresnel.c_eta_t = tex_spectrum(sc, m.tex[0], *si);
      b.fresnel.c_k = tex_spectrum(sc, m.tex[1], *si);
      bsdf.add(b);
      *out = bsdf;
      return true;
    }
    case PTRS_MAT_SUBSTRATE: {  // substrate.rs:42-68
      BSDF bsdf = BSDF::make(*si, 1.0f);
      Spectrum d = tex_spectrum(sc, m.tex[0], *si), s = tex_spectrum(sc, m.tex[1], *si);
      float rough_u = tex_f32(sc, m.tex[2], *si), rough_v = tex_f32(sc, m.tex[3], *si);
      if (!is_black(d) || is_black(s)) {  // sic (substrate.rs:55)
        if (m.remap_roughness) {
          rough_u = roughness_to_alpha(rough_u);
          rough_v = roughness_to_alpha(rough_v);
        }
        BxDF b;
        b.kind = BxDF::FresnelBlend;
        b.r = d;
        b.t = s;
        b.dist = Distribution::make(rough_u, rough_v, false);
        bsdf.add(b);
      }
      *out = bsdf;
      return true;
    }
    default: {  // PTRS_MAT_DISNEY, disney.rs:172-263
      BSDF bsdf = BSDF::make(*si, 1.0f);
      Spectrum c = tex_spectrum(sc, m.tex[0], *si);
      float metallic_weight = tex_f32(sc, m.tex[1], *si);
      float e = tex_f32(sc, m.tex[2], *si);
      float strans = 0.0f;
      float diffuse_weight = (1.0f - metallic_weight) * (1.0f - strans);
      float rough = tex_f32(sc, m.tex[3], *si);
      float lum = lum_y(c);
      Spectrum c_tint = lum > 0.0f ? c / lum : S(1.0f);
      if (diffuse_weight > 0.0f) {
        BxDF b;
        b.kind = BxDF::DisneyDiffuse;
        b.r = diffuse_weight * c;
        bsdf.add(b);
      }
      float aspect = 1.0f;
      float ax = rmax(0.001f, sqr(rough) / aspect), ay = rmax(0.001f, sqr(rough) * aspect);
      float spec_tint = 0.0f;
      Spectrum c_spec_0 = lerp(schlick_r0_from_eta(e) * lerp(S(1.f), c_tint, spec_tint), c, metallic_weight);
      BxDF b;
      b.kind = BxDF::MicrofacetReflection;
      b.r = S(1.f);
      b.dist = Distribution::make(ax, ay, true);
      b.fresnel.kind = Fresnel::Disney;
      b.fresnel.r0 = c_spec_0;
      b.fresnel.metallic = metallic_weight;
      b.fresnel.d_eta = e;
      bsdf.add(b);
      *out = bsdf;
      return true;
    }
  }
}

// ---- lights --------------------------------------------------------------------------------------
struct VisibilityTester {  // light.rs:33-42
  Interaction p0, p1;
};
inline bool is_delta_light(const PtrsLight& l) { return l.type == PTRS_LIGHT_POINT || l.type == PTRS_LIGHT_DIRECTIONAL; }

// Projective3 * Vector3 on a row-major 4x4 whose bottom row is (0,0,0,1)
inline Vec3 xform_vec(const float* m, Vec3 v) {
  return V(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

// Distribution1D::sample_continuous, sampling.rs:164-182
inline float dist1d_sample_continuous(const float* func, const float* cdf, float func_int, int n, float u, float* pdf, size_t* off) {
  size_t offset = find_interval((size_t)n + 1, [&](size_t i) { return cdf[i] <= u; });
  if (off) *off = offset;
  float du = u - cdf[offset];
  if ((cdf[offset + 1] - cdf[offset]) > 0.0f) du /= cdf[offset + 1] - cdf[offset];
  *pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
  return ((float)offset + du) / (float)n;
}
// Distribution2D::sample_continuous / pdf, sampling.rs:211-229
inline Vec2 dist2d_sample_continuous(const PtrsEnvLight& e, Vec2 u, float* pdf) {
  float pdfs[2] = {0, 0};
  size_t v = 1;
  float d1 = dist1d_sample_continuous(e.marg_func, e.marg_cdf, e.marg_func_int, e.nv, u.y, &pdfs[1], &v);
  float d0 = dist1d_sample_continuous(e.cond_func + v * e.nu, e.cond_cdf + v * (e.nu + 1), e.cond_func_int[v], e.nu, u.x, &pdfs[0], nullptr);
  *pdf = pdfs[0] * pdfs[1];
  return Vec2{d0, d1};
}
inline float dist2d_pdf(const PtrsEnvLight& e, Vec2 p) {
  uint64_t iu = f2usize(p.x * (float)e.nu), iv = f2usize(p.y * (float)e.nv);
  if (iu > (uint64_t)e.nu - 1) iu = e.nu - 1;
  if (iv > (uint64_t)e.nv - 1) iv = e.nv - 1;
  return e.cond_func[iv * e.nu + iu] / e.marg_func_int;
}

inline Spectrum env_lookup(const Scene& sc, const PtrsEnvLight& e, Vec2 st) {
  float o[3];
  mip_lookup_width(sc.d, sc.d->mipmaps[e.mip], st, 0.0f, o);
  return S(o[0], o[1], o[2]);
}

// DiffuseAreaLight::l, light.rs:252-258
inline Spectrum area_light_l(const Scene& sc, const PtrsLight& l, const SurfaceInteraction& inter, Vec3 w) {
  return dot(inter.general.n, w) > 0.0f ? tex_spectrum(sc, l.ke_tex, inter) : S(0.0f);
}
// SurfaceMediumInteraction::le, interaction.rs:297-303
inline Spectrum isect_le(const Scene& sc, const SurfaceInteraction& isect, Vec3 w) {
  int lid = sc.d->prim_area_light[isect.primitive];
  return lid >= 0 ? area_light_l(sc, sc.d->lights[lid], isect, w) : S(0.0f);
}
// Light::le: InfiniteAreaLight, light.rs:488-498; everything else returns black (light.rs:45-47)
inline Spectrum light_le(const Scene& sc, const PtrsLight& l, const Ray& r) {
  if (l.type != PTRS_LIGHT_INFINITE) return S(0.0f);
  const PtrsEnvLight& e = sc.d->envs[l.env];
  Vec3 w = normalize(xform_vec(e.world_to_light, r.d));
  Vec2 st{spherical_phi(w) * INV_2_PI, spherical_theta(w) * FRAC_1_PI};
  return env_lookup(sc, e, st);
}

// Light::sample_li.  *has_vis mirrors the Option<VisibilityTester> (None => the reference panics)
inline Spectrum light_sample_li(const Scene& sc, const PtrsLight& l, const Interaction& reference, Vec2 u, Vec3* wi, float* pdf,
                                VisibilityTester* vis, bool* has_vis) {
  *has_vis = true;
  switch (l.type) {
    case PTRS_LIGHT_POINT: {  // light.rs:97-116
      Vec3 p_light = V(l.pos[0], l.pos[1], l.pos[2]);
      *wi = normalize(p_light - reference.p);
      *pdf = 1.0f;
      vis->p0 = reference;
      vis->p1 = Interaction();
      vis->p1.p = p_light;
      return S(l.color[0], l.color[1], l.color[2]) / norm_squared(p_light - reference.p);
    }
    case PTRS_LIGHT_DIRECTIONAL: {  // light.rs:176-196
      Vec3 w_light = V(l.pos[0], l.pos[1], l.pos[2]);
      *wi = w_light;
      *pdf = 1.0f;
      vis->p0 = reference;
      vis->p1 = Interaction();
      vis->p1.p = reference.p + w_light * (2.0f * l.world_radius);
      return S(l.color[0], l.color[1], l.color[2]);
    }
    case PTRS_LIGHT_AREA: {  // light.rs:262-280
      SurfaceInteraction p_shape = triangle_sample(sc, l.prim, u);
      p_shape.primitive = l.prim;
      *wi = normalize(p_shape.general.p - reference.p);
      *pdf = triangle_pdf_at_point(sc, l.prim, reference, *wi, l.area);
      vis->p0 = reference;
      vis->p1 = p_shape.general;
      return area_light_l(sc, l, p_shape, -*wi);
    }
    default: {  // PTRS_LIGHT_INFINITE, light.rs:402-441
      const PtrsEnvLight& e = sc.d->envs[l.env];
      float map_pdf = 0.0f;
      Vec2 uv = dist2d_sample_continuous(e, u, &map_pdf);
      if (map_pdf == 0.0f) {
        *has_vis = false;
        return S(0.0f);
      }
      float theta = uv.y * PI, phi = uv.x * 2.0f * PI;
      float cos_theta_ = std::cos(theta), sin_theta_ = std::sin(theta);
      float sin_phi_ = std::sin(phi), cos_phi_ = std::cos(phi);
      *wi = xform_vec(e.light_to_world, V(sin_theta_ * cos_phi_, sin_theta_ * sin_phi_, cos_theta_));
      if (sin_theta_ == 0.0f) *pdf = 0.0f;
      else *pdf = map_pdf / (2.0f * PI * PI * sin_theta_);
      vis->p0 = reference;
      vis->p1 = Interaction();
      vis->p1.p = reference.p + *wi * (2.0f * l.world_radius);
      return env_lookup(sc, e, uv);
    }
  }
}

// Light::pdf_li
inline float light_pdf_li(const Scene& sc, const PtrsLight& l, const Interaction& reference, Vec3 w) {
  switch (l.type) {
    case PTRS_LIGHT_AREA: return triangle_pdf_at_point(sc, l.prim, reference, w, l.area);  // light.rs:286-288
    case PTRS_LIGHT_INFINITE: {                                                           // light.rs:447-461
      const PtrsEnvLight& e = sc.d->envs[l.env];
      Vec3 wi = xform_vec(e.world_to_light, w);
      float theta = spherical_theta(wi), phi = spherical_phi(wi);
      float sin_theta_ = std::sin(theta);
      if (sin_theta_ == 0.0f) return 0.0f;
      return dist2d_pdf(e, Vec2{phi * INV_2_PI, theta * FRAC_1_PI}) / (2.0f * PI * PI * sin_theta_);
    }
    default: return 0.0f;  // light.rs:122-124, 202-204
  }
}

struct RayCounters {
  uint64_t extension = 0, shadow = 0, mis = 0;
  TraversalCounters trav;
};

// estimate_direct, integrator.rs:23-139 (handle_media = false, specular = false)
inline Spectrum estimate_direct(const Scene& sc, const SurfaceInteraction& it, const BSDF& bsdf, Vec2 u_scattering, int light_idx,
                                Vec2 u_light, RayCounters* rc) {
  const PtrsLight& light = sc.d->lights[light_idx];
  const uint32_t bsdf_flags = BSDF_ALL & ~BSDF_SPECULAR;
  Spectrum ld = S(0.0f);
  Vec3 wi = V(0, 0, 0);
  float light_pdf = 0.0f, scattering_pdf = 0.0f;
  VisibilityTester vis;
  bool has_vis = false;
  Spectrum li = light_sample_li(sc, light, it.general, u_light, &wi, &light_pdf, &vis, &has_vis);
  if (!has_vis) throw std::runtime_error("estimate_direct: visibility.unwrap() on None (integrator.rs:51)");
  if (light_pdf > 0.0f && !is_black(li)) {
    Spectrum f = bsdf.f(it.general.wo, wi, bsdf_flags) * std::fabs(dot(wi, it.shading.n));
    scattering_pdf = bsdf.pdf(it.general.wo, wi, bsdf_flags);
    if (!is_black(f)) {
      if (rc) rc->shadow++;
      if (bvh_intersect_p(sc, vis.p0.spawn_ray_to_it(vis.p1), rc ? &rc->trav : nullptr)) li = S(0.0f);
      if (!is_black(li)) {
        if (is_delta_light(light)) {
          ld += f * li / light_pdf;
        } else {
          float weight = power_heuristic(1, light_pdf, 1, scattering_pdf);
          ld += f * li * weight / light_pdf;
        }
      }
    }
  }
  if (!is_delta_light(light)) {
    uint32_t sampled_type = BSDF_ALL;
    Spectrum f = bsdf.sample_f(it.general.wo, &wi, u_scattering, &scattering_pdf, bsdf_flags, &sampled_type);
    f *= std::fabs(dot(wi, it.shading.n));
    bool sampled_specular = (sampled_type & BSDF_SPECULAR) == BSDF_SPECULAR;
    if (!is_black(f) && scattering_pdf > 0.0f) {
      float weight = 1.0f;
      if (!sampled_specular) {
        light_pdf = light_pdf_li(sc, light, it.general, wi);
        if (light_pdf == 0.0f) return ld;
        weight = power_heuristic(1, scattering_pdf, 1, light_pdf);
      }
      SurfaceInteraction light_isect;
      Ray ray = it.general.spawn_ray(wi);
      if (rc) rc->mis++;
      bool found = bvh_intersect(sc, &ray, &light_isect, nullptr, rc ? &rc->trav : nullptr);
      Spectrum li2 = S(0.0f);
      if (found) {
        if (sc.d->prim_area_light[light_isect.primitive] == light_idx) li2 = isect_le(sc, light_isect, -wi);
      } else {
        li2 = light_le(sc, light, ray);
      }
      if (!is_black(li2)) ld += f * li2 * S(1.0f) * weight / scattering_pdf;
    }
  }
  return ld;
}

// uniform_sample_one_light, integrator.rs:192-217
inline Spectrum uniform_sample_one_light(const Scene& sc, const SurfaceInteraction& it, const BSDF& bsdf, SobolSampler* sampler,
                                         RayCounters* rc) {
  const size_t num_lights = sc.d->n_lights;
  if (num_lights == 0) return S(0.0f);
  Vec2 u_light = sampler->get_2d();
  Vec2 u_scattering = sampler->get_2d();
  uint64_t li = f2usize(std::floor(sampler->get_1d() * (float)num_lights));
  size_t light_idx = li < num_lights - 1 ? (size_t)li : num_lights - 1;
  return (float)num_lights * estimate_direct(sc, it, bsdf, u_scattering, (int)light_idx, u_light, rc);
}

}  // namespace oracle
