// ORACLE — TEST INFRASTRUCTURE ONLY.  BxDFs and the BSDF container:
//   src/pathtracer/bxdf/mod.rs, bxdf/fresnel.rs, bxdf/microfacet.rs, material/disney.rs:57-170,
//   src/pathtracer/sampling.rs:96-126, src/pathtracer/bsdf.rs
#pragma once
#include "om_scene.hpp"

namespace oracle {

// bxdf/mod.rs:11-67 ------------------------------------------------------------------------------
inline float cos_theta(Vec3 w) { return w.z; }
inline float cos_2_theta(Vec3 w) { return w.z * w.z; }
inline float abs_cos_theta(Vec3 w) { return std::fabs(w.z); }
inline float sin_2_theta(Vec3 w) { return rmax(0.0f, 1.0f - cos_2_theta(w)); }
inline float sin_theta(Vec3 w) { return std::sqrt(sin_2_theta(w)); }
inline float tan_2_theta(Vec3 w) { return sin_2_theta(w) / cos_2_theta(w); }
inline float tan_theta(Vec3 w) { return sin_theta(w) / cos_theta(w); }
inline float cos_phi(Vec3 w) {
  float st = sin_theta(w);
  return st == 0.0f ? 1.0f : rclamp(w.x / st, -1.0f, 1.0f);
}
inline float sin_phi(Vec3 w) {
  float st = sin_theta(w);
  return st == 0.0f ? 1.0f : rclamp(w.y / st, -1.0f, 1.0f);  // sic: 1.0 (bxdf/mod.rs:49-56)
}
inline float cos_2_phi(Vec3 w) { return cos_phi(w) * cos_phi(w); }
inline float sin_2_phi(Vec3 w) { return sin_phi(w) * sin_phi(w); }
inline bool same_hemisphere(Vec3 w, Vec3 wp) { return w.z * wp.z > 0.0f; }
inline Vec3 reflect(Vec3 wo, Vec3 n) { return -wo + 2.f * dot(wo, n) * n; }  // bxdf/mod.rs:69-71
inline bool refract(Vec3 wi, Vec3 n, float eta, Vec3* wt) {                  // bxdf/mod.rs:73-89
  float cos_theta_i = dot(n, wi);
  float sin_2_theta_i = rmax(0.0f, 1.0f - cos_theta_i * cos_theta_i);
  float sin_2_theta_t = eta * eta * sin_2_theta_i;
  if (sin_2_theta_t > 1.0f) return false;
  float cos_theta_t = std::sqrt(1.0f - sin_2_theta_t);
  *wt = eta * -wi + (eta * cos_theta_i - cos_theta_t) * n;
  return true;
}

enum : uint32_t {  // bxdf/mod.rs:91-101
  BSDF_REFLECTION = 1 << 0,
  BSDF_TRANSMISSION = 1 << 1,
  BSDF_DIFFUSE = 1 << 2,
  BSDF_GLOSSY = 1 << 3,
  BSDF_SPECULAR = 1 << 4,
  BSDF_ALL = 31
};

// sampling.rs:96-126 ------------------------------------------------------------------------------
inline Vec2 concentric_sample_disk(Vec2 u) {
  float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
  if (ox == 0.0f && oy == 0.0f) return Vec2{0.0f, 0.0f};
  float theta, r;
  if (std::fabs(ox) > std::fabs(oy)) {
    r = ox;
    theta = FRAC_PI_4 * (oy / ox);
  } else {
    r = oy;
    theta = FRAC_PI_2 - FRAC_PI_4 * (ox / oy);
  }
  return Vec2{r * std::cos(theta), r * std::sin(theta)};
}
inline Vec3 cosine_sample_hemisphere(Vec2 u) {
  Vec2 d = concentric_sample_disk(u);
  float z = std::sqrt(rmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
  return V(d.x, d.y, z);
}

// fresnel.rs:21-64 --------------------------------------------------------------------------------
inline float fr_dielectric(float cos_theta_i, float eta_i, float eta_t) {
  cos_theta_i = rclamp(cos_theta_i, -1.0f, 1.0f);
  bool entering = cos_theta_i > 0.0f;
  if (!entering) {
    float tmp = eta_i;
    eta_i = eta_t;
    eta_t = tmp;
    cos_theta_i = std::fabs(cos_theta_i);
  }
  float sin_theta_i = std::sqrt(rmax(0.0f, 1.0f - cos_theta_i * cos_theta_i));
  float sin_theta_t = eta_i / eta_t * sin_theta_i;
  if (sin_theta_t >= 1.0f) return 1.0f;
  float cos_theta_t = std::sqrt(rmax(0.0f, 1.0f - sin_theta_t * sin_theta_t));
  float r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
  float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
  return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
inline Spectrum fr_conductor(float cos_theta_i, Spectrum eta_i, Spectrum eta_t, Spectrum k) {
  cos_theta_i = rclamp(cos_theta_i, -1.f, 1.f);
  Spectrum eta = eta_t / eta_i, etak = k / eta_i;
  float cos_theta_i2 = cos_theta_i * cos_theta_i;
  float sin_theta_i2 = 1.f - cos_theta_i2;
  Spectrum eta2 = eta * eta, etak2 = etak * etak;
  Spectrum t0 = eta2 - etak2 - S(sin_theta_i2);
  Spectrum a2_plus_b2 = ssqrt(t0 * t0 + 4.f * eta2 * etak2);
  Spectrum t1 = a2_plus_b2 + S(cos_theta_i2);
  Spectrum a = ssqrt(0.5f * (a2_plus_b2 + t0));
  Spectrum t2 = 2.f * cos_theta_i * a;
  Spectrum rs = (t1 - t2) / (t1 + t2);
  Spectrum t3 = cos_theta_i2 * a2_plus_b2 + S(sin_theta_i2 * sin_theta_i2);
  Spectrum t4 = t2 * sin_theta_i2;
  Spectrum rp = rs * (t3 - t4) / (t3 + t4);
  return 0.5f * (rp + rs);
}

// disney.rs:57-68
inline float schlick_weight(float cos_theta) {
  float m = rclamp(1.0f - cos_theta, 0.0f, 1.0f);
  return (m * m) * (m * m) * m;
}
inline Spectrum fr_schlick_spectrum(Spectrum r0, float cos_theta) { return lerp(r0, S(1.f), schlick_weight(cos_theta)); }

struct Fresnel {  // fresnel.rs:12-19
  enum Kind { Dielectric, Conductor, Disney, NoOp } kind = NoOp;
  float eta_i = 1, eta_t = 1;                 // Dielectric
  Spectrum c_eta_i{1, 1, 1}, c_eta_t{1, 1, 1}, c_k{0, 0, 0};  // Conductor
  Spectrum r0{0, 0, 0};                       // Disney
  float metallic = 0, d_eta = 1;
  Spectrum evaluate(float cos_i) const {
    switch (kind) {
      case Dielectric: return S(fr_dielectric(cos_i, eta_i, eta_t));                      // fresnel.rs:77-81
      case Conductor: return fr_conductor(std::fabs(cos_i), c_eta_i, c_eta_t, c_k);       // fresnel.rs:95-99
      case Disney:                                                                         // disney.rs:128-136
        return lerp(S(fr_dielectric(cos_i, 1.f, d_eta)), fr_schlick_spectrum(r0, cos_i), metallic);
      default: return S(1.0f);                                                             // fresnel.rs:103-107
    }
  }
};

// microfacet.rs:14-174 + disney.rs:138-170 --------------------------------------------------------
inline void trowbridge_reitz_sample_11(float cos_theta, float u1, float u2, float* slope_x, float* slope_y) {
  if (cos_theta > 0.9999f) {
    float r = std::sqrt(u1 / (1.f - u1));
    float phi = 6.28318530718f * u2;
    *slope_x = r * std::cos(phi);
    *slope_y = r * std::sin(phi);
    return;
  }
  float sin_theta = std::sqrt(rmax(0.0f, 1.f - cos_theta * cos_theta));
  float tan_theta = sin_theta / cos_theta;
  float alpha = 1.f / tan_theta;
  float g1 = 2.f / (1.f + std::sqrt(1.f + 1.f / (alpha * alpha)));
  float a = 2.f * u1 / g1 - 1.f;
  float tmp = 1.f / (a * a - 1.f);
  if (tmp > 1e10f) tmp = 1e10f;
  float b = tan_theta;
  float d = std::sqrt(rmax(0.0f, b * b * tmp * tmp - (a * a - b * b) * tmp));
  float slope_x_1 = b * tmp - d, slope_x_2 = b * tmp + d;
  *slope_x = (a < 0.f || slope_x_2 > (1.f / tan_theta)) ? slope_x_1 : slope_x_2;
  float s;
  if (u2 > 0.5f) {
    s = 1.f;
    u2 = 2.f * (u2 - 0.5f);
  } else {
    s = -1.f;
    u2 = 2.f * (0.5f - u2);
  }
  float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) /
            (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  *slope_y = s * z * std::sqrt(1.f + *slope_x * *slope_x);
}
inline Vec3 trowbridge_reitz_sample(Vec3 wi, float alpha_x, float alpha_y, float u1, float u2) {
  Vec3 wi_stretched = normalize(V(alpha_x * wi.x, alpha_y * wi.y, wi.z));
  float slope_x = 0.0f, slope_y = 0.0f;
  trowbridge_reitz_sample_11(cos_theta(wi_stretched), u1, u2, &slope_x, &slope_y);
  float tmp = cos_phi(wi_stretched) * slope_x - sin_phi(wi_stretched) * slope_y;
  slope_y = sin_phi(wi_stretched) * slope_x + cos_phi(wi_stretched) * slope_y;
  slope_x = tmp;
  slope_x = alpha_x * slope_x;
  slope_y = alpha_y * slope_y;
  return normalize(V(-slope_x, -slope_y, 1.f));
}
inline float roughness_to_alpha(float roughness) {  // microfacet.rs:119-128
  roughness = rmax(roughness, 1e-3f);
  float x = std::log(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
struct Distribution {  // TrowbridgeReitzDistribution / DisneyMicrofacetDistribution
  float alpha_x = 0.001f, alpha_y = 0.001f;
  bool disney = false;  // separable G (disney.rs:160-162)
  static Distribution make(float ax, float ay, bool disney) {
    Distribution d;
    d.alpha_x = rmax(ax, 0.001f);
    d.alpha_y = rmax(ay, 0.001f);
    d.disney = disney;
    return d;
  }
  float d(Vec3 wh) const {
    float t2 = tan_2_theta(wh);
    if (std::isinf(t2)) return 0.0f;
    float cos_4_theta = cos_2_theta(wh) * cos_2_theta(wh);
    float e = (cos_2_phi(wh) / (alpha_x * alpha_x) + sin_2_phi(wh) / (alpha_y * alpha_y)) * t2;
    return 1.0f / (PI * alpha_x * alpha_y * cos_4_theta * (1.0f + e) * (1.0f + e));
  }
  float lambda(Vec3 w) const {
    float abs_tan_theta = std::fabs(tan_theta(w));
    if (std::isinf(abs_tan_theta)) return 0.0f;
    float alpha = std::sqrt((cos_2_phi(w) * alpha_x * alpha_x) + (sin_2_phi(w) * alpha_y * alpha_y));
    float a2t2 = (alpha * abs_tan_theta) * (alpha * abs_tan_theta);
    return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
  }
  float g1(Vec3 w) const { return 1.0f / (1.0f + lambda(w)); }
  float g(Vec3 wo, Vec3 wi) const {
    if (disney) return g1(wo) * g1(wi);
    return 1.0f / (1.0f + lambda(wo) + lambda(wi));
  }
  Vec3 sample_wh(Vec3 wo, Vec2 u) const {
    bool flip = wo.z < 0.f;
    Vec3 w = flip ? -wo : wo;
    Vec3 wh = trowbridge_reitz_sample(w, alpha_x, alpha_y, u.x, u.y);
    return flip ? -wh : wh;
  }
  float pdf(Vec3 wo, Vec3 wh) const { return d(wh) * g1(wo) * std::fabs(dot(wo, wh)) / abs_cos_theta(wo); }
};

// BxDF enum, bxdf/mod.rs:184-193 -------------------------------------------------------------------
struct BxDF {
  enum Kind { Lambertian, SpecularReflection, SpecularTransmission, FresnelSpecular, MicrofacetReflection,
              MicrofacetTransmission, FresnelBlend, DisneyDiffuse } kind = Lambertian;
  Spectrum r{0, 0, 0}, t{0, 0, 0};  // r also = rd (FresnelBlend); t = rs (FresnelBlend)
  float eta_a = 1, eta_b = 1;
  Fresnel fresnel;
  Distribution dist;

  uint32_t get_type() const {
    switch (kind) {
      case Lambertian: case DisneyDiffuse: return BSDF_REFLECTION | BSDF_DIFFUSE;
      case SpecularReflection: return BSDF_REFLECTION | BSDF_SPECULAR;
      case SpecularTransmission: return BSDF_TRANSMISSION | BSDF_SPECULAR;
      case FresnelSpecular: return BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
      case MicrofacetReflection: case FresnelBlend: return BSDF_REFLECTION | BSDF_GLOSSY;
      default: return BSDF_TRANSMISSION | BSDF_GLOSSY;
    }
  }
  bool matches_flags(uint32_t t_) const { return (get_type() & t_) == get_type(); }  // bxdf/mod.rs:167-169

  Spectrum schlick_fresnel(float cos_theta_) const {  // microfacet.rs:401-404 (rs = t)
    auto pow5 = [](float v) { return (v * v) * (v * v) * v; };
    return t + pow5(1.0f - cos_theta_) * (S(1.f) - t);
  }

  Spectrum f(Vec3 wo, Vec3 wi) const {
    switch (kind) {
      case Lambertian: return r * FRAC_1_PI;  // bxdf/mod.rs:206-208
      case DisneyDiffuse: {                   // disney.rs:80-88
        float fo = schlick_weight(abs_cos_theta(wo)), fi = schlick_weight(abs_cos_theta(wi));
        return r * FRAC_1_PI * (1.f - fo / 2.f) * (1.f - fi / 2.f);
      }
      case MicrofacetReflection: {            // microfacet.rs:197-212
        float cos_theta_o = abs_cos_theta(wo), cos_theta_i = abs_cos_theta(wi);
        Vec3 wh = wi + wo;
        if (cos_theta_i == 0.f || cos_theta_o == 0.f) return S(0.f);
        if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) return S(0.f);
        wh = normalize(wh);
        Spectrum fr = fresnel.evaluate(dot(wi, wh));
        return r * dist.d(wh) * dist.g(wo, wi) * fr / (4.0f * cos_theta_i * cos_theta_o);
      }
      case MicrofacetTransmission: {          // microfacet.rs:285-329
        if (same_hemisphere(wo, wi)) return S(0.f);
        float cos_theta_o = abs_cos_theta(wo), cos_theta_i = abs_cos_theta(wi);
        if (cos_theta_i == 0.f || cos_theta_o == 0.f) return S(0.f);
        float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
        Vec3 wh = normalize(wo + wi * eta);
        if (wh.z < 0.0f) wh = -wh;
        if (dot(wo, wh) * dot(wi, wh) > 0.f) return S(0.f);
        Spectrum fr = S(fr_dielectric(dot(wo, wh), eta_a, eta_b));
        float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        float factor = 1.0f / eta;  // TransportMode::Radiance
        return (S(1.f) - fr) * t *
               (dist.d(wh) * dist.g(wo, wi) * eta * eta * std::fabs(dot(wi, wh)) * std::fabs(dot(wo, wh)) * factor * factor /
                (cos_theta_i * cos_theta_o * sqrt_denom * sqrt_denom));
      }
      case FresnelBlend: {                    // microfacet.rs:408-427 (rd = r, rs = t)
        auto pow5 = [](float v) { return (v * v) * (v * v) * v; };
        Spectrum diffuse = (28.f / (23.f * PI)) * r * (S(1.f) - t) * (1.f - pow5(1.f - 0.5f * abs_cos_theta(wi))) *
                           (1.f - pow5(1.f - 0.5f * abs_cos_theta(wo)));
        Vec3 wh = wi + wo;
        if (is_zero(wh)) return S(0.f);
        wh = normalize(wh);
        Spectrum specular = dist.d(wh) / (4.f * std::fabs(dot(wi, wh)) * rmax(abs_cos_theta(wi), abs_cos_theta(wo))) *
                            schlick_fresnel(dot(wi, wh));
        return diffuse + specular;
      }
      default: return S(0.0f);  // specular lobes: fresnel.rs:121-123, 173-175, 240-242
    }
  }

  float pdf(Vec3 wo, Vec3 wi) const {
    switch (kind) {
      case Lambertian: case DisneyDiffuse:  // default, bxdf/mod.rs:173-179
        return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * FRAC_1_PI : 0.0f;
      case MicrofacetReflection: {          // microfacet.rs:245-251
        if (!same_hemisphere(wo, wi)) return 0.f;
        Vec3 wh = normalize(wo + wi);
        return dist.pdf(wo, wh) / (4.f * dot(wo, wh));
      }
      case MicrofacetTransmission: {        // microfacet.rs:363-383 (sic: rejects the OPPOSITE hemisphere)
        if (!same_hemisphere(wo, wi)) return 0.f;
        float eta = cos_theta(wo) > 0.f ? eta_a / eta_b : eta_b / eta_a;
        Vec3 wh = normalize(wo + wi * eta);
        if (dot(wo, wh) * dot(wi, wh) > 0.f) return 0.f;
        float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        float dwh_dwi = std::fabs((eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
        return dist.pdf(wo, wh) * dwh_dwi;
      }
      case FresnelBlend: {                  // microfacet.rs:460-469
        if (!same_hemisphere(wo, wi)) return 0.f;
        Vec3 wh = normalize(wo + wi);
        float pdf_wh = dist.pdf(wo, wh);
        return 0.5f * (abs_cos_theta(wi) * FRAC_1_PI + pdf_wh / (4.f * dot(wo, wh)));
      }
      default: return 0.0f;
    }
  }

  // returns f; *sampled_type is overwritten only by FresnelSpecular (fresnel.rs:254-288)
  Spectrum sample_f(Vec3 wo, Vec3* wi, Vec2 u, float* pdf_, uint32_t* sampled_type) const {
    switch (kind) {
      case Lambertian: case DisneyDiffuse: {  // default, bxdf/mod.rs:106-121
        *wi = cosine_sample_hemisphere(u);
        if (wo.z < 0.0f) wi->z *= -1.0f;
        *pdf_ = pdf(wo, *wi);
        return f(wo, *wi);
      }
      case SpecularReflection: {  // fresnel.rs:129-140
        *wi = V(-wo.x, -wo.y, wo.z);
        *pdf_ = 1.0f;
        return fresnel.evaluate(cos_theta(*wi)) * r / abs_cos_theta(*wi);
      }
      case SpecularTransmission: {  // fresnel.rs:181-208
        bool entering = cos_theta(wo) > 0.0f;
        float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
        if (!refract(wo, face_forward(V(0.f, 0.f, 1.f), wo), eta_i / eta_t, wi)) return S(0.0f);
        *pdf_ = 1.0f;
        Spectrum ft = t * (S(1.0f) - S(fr_dielectric(cos_theta(*wi), eta_a, eta_b)));
        ft *= (eta_i * eta_i) / (eta_t * eta_t);
        return ft / abs_cos_theta(*wi);
      }
      case FresnelSpecular: {  // fresnel.rs:244-288
        float fr = fr_dielectric(cos_theta(wo), eta_a, eta_b);
        if (u.x < fr) {
          *wi = V(-wo.x, -wo.y, wo.z);
          if (sampled_type) *sampled_type = BSDF_REFLECTION | BSDF_SPECULAR;
          *pdf_ = fr;
          return fr * r / abs_cos_theta(*wi);
        } else {
          bool entering = cos_theta(wo) > 0.0f;
          float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
          if (!refract(wo, face_forward(V(0.f, 0.f, 1.f), wo), eta_i / eta_t, wi)) return S(0.0f);
          Spectrum ft = t * (S(1.0f) - S(fr));
          ft *= (eta_i * eta_i) / (eta_t * eta_t);
          if (sampled_type) *sampled_type = BSDF_TRANSMISSION | BSDF_SPECULAR;
          *pdf_ = 1.0f - fr;
          return ft / abs_cos_theta(*wi);
        }
      }
      case MicrofacetReflection: {  // microfacet.rs:218-243
        if (wo.z == 0.f) return S(0.f);
        Vec3 wh = dist.sample_wh(wo, u);
        if (dot(wo, wh) < 0.f) return S(0.f);
        *wi = reflect(wo, wh);
        if (!same_hemisphere(wo, *wi)) return S(0.f);
        *pdf_ = dist.pdf(wo, wh) / (4.f * dot(wo, wh));
        return f(wo, *wi);
      }
      case MicrofacetTransmission: {  // microfacet.rs:335-361
        if (wo.z == 0.f) return S(0.f);
        Vec3 wh = dist.sample_wh(wo, u);
        if (dot(wo, wh) < 0.f) return S(0.f);
        float eta = cos_theta(wo) > 0.f ? eta_a / eta_b : eta_b / eta_a;
        if (!refract(wo, wh, eta, wi)) return S(0.f);
        *pdf_ = pdf(wo, *wi);
        return f(wo, *wi);
      }
      default: {  // FresnelBlend, microfacet.rs:433-458
        Vec2 uu = u;
        if (uu.x < 0.5f) {
          uu.x = rmin(2.f * uu.x, ONE_MINUS_EPSILON);
          *wi = cosine_sample_hemisphere(uu);
          if (wo.z < 0.f) wi->z *= -1.f;
        } else {
          uu.x = rmin(2.f * (uu.x - 0.5f), ONE_MINUS_EPSILON);
          Vec3 wh = dist.sample_wh(wo, uu);
          *wi = reflect(wo, wh);
          if (!same_hemisphere(wo, *wi)) return S(0.f);
        }
        *pdf_ = pdf(wo, *wi);
        return f(wo, *wi);
      }
    }
  }
};

// BSDF, bsdf.rs:8-222 ------------------------------------------------------------------------------
struct BSDF {
  float eta = 1.0f;
  Vec3 ns, ng, ss, ts;
  int n_bxdfs = 0;
  BxDF bxdfs[8];

  static BSDF make(const SurfaceInteraction& si, float eta) {  // bsdf.rs:20-34
    BSDF b;
    b.eta = eta;
    b.ns = si.shading.n;
    b.ss = normalize(si.shading.dpdu);
    b.ng = si.general.n;
    b.ts = cross(b.ns, b.ss);
    return b;
  }
  void add(const BxDF& x) { bxdfs[n_bxdfs++] = x; }
  int num_components(uint32_t flags) const {
    int n = 0;
    for (int i = 0; i < n_bxdfs; ++i) n += bxdfs[i].matches_flags(flags) ? 1 : 0;
    return n;
  }
  Vec3 world_to_local(Vec3 v) const { return V(dot(v, ss), dot(v, ts), dot(v, ns)); }
  Vec3 local_to_world(Vec3 v) const {
    return V(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
  }

  // bsdf.rs:66-148
  Spectrum sample_f(Vec3 wo_world, Vec3* wi_world, Vec2 u, float* pdf, uint32_t bxdf_type, uint32_t* sampled_type) const {
    int matching_comps = num_components(bxdf_type);
    if (matching_comps == 0) {
      *pdf = 0.0f;
      if (sampled_type) *sampled_type = 0;
      return S(0.0f);
    }
    uint64_t c64 = f2usize(std::floor(u.x * (float)matching_comps));
    int comp = (int)(c64 < (uint64_t)(matching_comps - 1) ? c64 : (uint64_t)(matching_comps - 1));
    const BxDF* bxdf = nullptr;
    int count = comp;
    for (int i = 0; i < n_bxdfs; ++i)
      if (bxdfs[i].matches_flags(bxdf_type)) {
        if (count == 0) {
          bxdf = &bxdfs[i];
          break;
        }
        count -= 1;
      }
    Vec2 u_remapped{(u.x * (float)matching_comps) - (float)comp, u.y};
    Vec3 wi = V(0, 0, 0);
    Vec3 wo = world_to_local(wo_world);
    *pdf = 0.0f;
    if (sampled_type) *sampled_type = bxdf->get_type();
    Spectrum f = bxdf->sample_f(wo, &wi, u_remapped, pdf, sampled_type);
    if (*pdf == 0.0f) {
      if (sampled_type) *sampled_type = 0;
      return S(0.0f);
    }
    *wi_world = local_to_world(wi);
    const bool spec = bxdf->get_type() & BSDF_SPECULAR;
    if (!spec && matching_comps > 1)
      for (int i = 0; i < n_bxdfs; ++i)
        if (&bxdfs[i] != bxdf && bxdfs[i].matches_flags(bxdf_type)) *pdf += bxdfs[i].pdf(wo, wi);
    if (matching_comps > 1) *pdf /= (float)matching_comps;
    if (!spec && matching_comps > 1) {
      bool reflect_ = dot(*wi_world, ng) * dot(wo_world, ng) > 0.0f;
      f = S(0.0f);
      for (int i = 0; i < n_bxdfs; ++i)
        if (bxdfs[i].matches_flags(bxdf_type) &&
            ((reflect_ && (bxdfs[i].get_type() & BSDF_REFLECTION)) || (!reflect_ && (bxdfs[i].get_type() & BSDF_TRANSMISSION))))
          f += bxdfs[i].f(wo, wi);
    }
    return f;
  }

  // bsdf.rs:150-187
  Spectrum f(Vec3 wo_w, Vec3 wi_w, uint32_t flags) const {
    Vec3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
    if (wo.z == 0.0f) return S(0.0f);
    bool reflect_ = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
    Spectrum f = S(0.0f);
    for (int i = 0; i < n_bxdfs; ++i)
      if (bxdfs[i].matches_flags(flags) &&
          ((reflect_ && (bxdfs[i].get_type() & BSDF_REFLECTION)) || (!reflect_ && (bxdfs[i].get_type() & BSDF_TRANSMISSION))))
        f += bxdfs[i].f(wo, wi);
    return f;
  }

  // bsdf.rs:189-222
  float pdf(Vec3 wo_world, Vec3 wi_world, uint32_t flags) const {
    if (n_bxdfs == 0) return 0.0f;
    Vec3 wo = world_to_local(wo_world), wi = world_to_local(wi_world);
    if (wo.z == 0.0f) return 0.0f;
    float pdf = 0.0f;
    int matching_comps = 0;
    for (int i = 0; i < n_bxdfs; ++i)
      if (bxdfs[i].matches_flags(flags)) {
        matching_comps += 1;
        pdf += bxdfs[i].pdf(wo, wi);
      }
    return matching_comps > 0 ? pdf / (float)matching_comps : 0.0f;
  }
};

}  // namespace oracle
