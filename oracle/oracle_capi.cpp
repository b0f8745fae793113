// ORACLE — TEST INFRASTRUCTURE ONLY.  C exports for the Python test / bench harness (ctypes).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this.
#include <cstring>
#include <string>

#include "om_integrator.hpp"

using namespace oracle;

namespace {
SobolTables g_tables;
bool g_init = false;
thread_local std::string g_err;

SobolSampler make_sampler(const PtrsCamera* cam, const PtrsRenderParams* rp) {
  Bounds2i sb = film_sample_bounds(cam->width, cam->height, rp->filter_radius);
  const int32_t sbv[4] = {sb.x0, sb.y0, sb.x1, sb.y1};
  SobolSampler s;
  s.configure(&g_tables, (size_t)rp->spp, sbv);
  return s;
}
}  // namespace

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

int oracle_init(const char* sobol_tables_path) {
  try {
    g_tables = SobolTables::load(sobol_tables_path);
    g_init = true;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// raw lowdiscrepancy.rs functions
uint64_t oracle_sobol_interval_to_index(uint32_t m, uint64_t frame, int32_t px, int32_t py) {
  return sobol_interval_to_index(g_tables, m, frame, px, py);
}
float oracle_sobol_sample(int64_t index, uint32_t dim, uint64_t scramble) { return sobol_sample(g_tables, index, dim, scramble); }

int oracle_sobol_samples(const PtrsCamera* cam, const PtrsRenderParams* rp, const int32_t* pixels_xy, const int32_t* sample_nums,
                         size_t n, const int32_t* dims, size_t n_dims, float* out, uint64_t* out_index) {
  SobolSampler s = make_sampler(cam, rp);
  for (size_t i = 0; i < n; ++i) {
    s.start_pixel(pixels_xy[2 * i], pixels_xy[2 * i + 1]);
    s.set_sample((size_t)sample_nums[i]);
    if (out_index) out_index[i] = (uint64_t)s.interval_sample_index;
    for (size_t k = 0; k < n_dims; ++k) out[i * n_dims + k] = s.sample_dimension(s.interval_sample_index, (size_t)dims[k]);
  }
  return 0;
}

int oracle_generate_rays(const PtrsCamera* cam, const PtrsRenderParams* rp, const int32_t* pixels_xy, const int32_t* sample_nums,
                         size_t n, PtrsRay* rays, float* p_film, float* rxry_dir) {
  SobolSampler s = make_sampler(cam, rp);
  const float scale = 1.0f / std::sqrt((float)s.samples_per_pixel);
  for (size_t i = 0; i < n; ++i) {
    s.start_pixel(pixels_xy[2 * i], pixels_xy[2 * i + 1]);
    s.set_sample((size_t)sample_nums[i]);
    Vec2 pf = s.get_camera_sample();
    RayDifferential rd = generate_ray_differential(*cam, pf);
    rd.scale_differentials(scale);
    std::memcpy(rays[i].o, &rd.ray.o, 12);
    std::memcpy(rays[i].d, &rd.ray.d, 12);
    rays[i].t_max = rd.ray.t_max;
    if (p_film) {
      p_film[2 * i] = pf.x;
      p_film[2 * i + 1] = pf.y;
    }
    if (rxry_dir) {
      std::memcpy(rxry_dir + 6 * i, &rd.rx_direction, 12);
      std::memcpy(rxry_dir + 6 * i + 3, &rd.ry_direction, 12);
    }
  }
  return 0;
}

int oracle_intersect(const PtrsSceneDesc* desc, const PtrsRay* rays, size_t n, PtrsHit* hits, uint64_t* counters2, int n_threads) {
  Scene sc{desc};
  uint64_t nodes = 0, tris = 0;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(n_threads) reduction(+ : nodes, tris)
  for (size_t i = 0; i < n; ++i) {
    Ray r{V(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].t_max};
    SurfaceInteraction isect;
    float b[3] = {0, 0, 0};
    TraversalCounters c;
    bool hit = bvh_intersect(sc, &r, &isect, b, &c);
    hits[i].prim = hit ? isect.primitive : -1;
    hits[i].t = hit ? r.t_max : 0.0f;
    hits[i].b0 = hit ? b[0] : 0.0f;
    hits[i].b1 = hit ? b[1] : 0.0f;
    hits[i].b2 = hit ? b[2] : 0.0f;
    nodes += c.nodes_tested;
    tris += c.tris_tested;
  }
  if (counters2) {
    counters2[0] = nodes;
    counters2[1] = tris;
  }
  return 0;
}

int oracle_intersect_p(const PtrsSceneDesc* desc, const PtrsRay* rays, size_t n, uint8_t* occluded, uint64_t* counters2, int n_threads) {
  Scene sc{desc};
  uint64_t nodes = 0, tris = 0;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(n_threads) reduction(+ : nodes, tris)
  for (size_t i = 0; i < n; ++i) {
    Ray r{V(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].t_max};
    TraversalCounters c;
    occluded[i] = bvh_intersect_p(sc, r, &c) ? 1 : 0;
    nodes += c.nodes_tested;
    tris += c.tris_tested;
  }
  if (counters2) {
    counters2[0] = nodes;
    counters2[1] = tris;
  }
  return 0;
}

// per-path radiance for chosen (pixel, sample) pairs: the li() of integrator.rs:579
int oracle_path_radiance(const PtrsSceneDesc* desc, const PtrsCamera* cam, const PtrsRenderParams* rp, const int32_t* pixels_xy,
                         const int32_t* sample_nums, size_t n, float* out_rgb, int n_threads) {
  try {
    Scene sc{desc};
    IntegratorParams P;
    P.max_depth = rp->max_depth;
    P.rr_threshold = rp->rr_threshold;
    P.rr_start_depth = rp->rr_start_depth;
    P.rr_enable = rp->rr_enable != 0;
    SobolSampler proto = make_sampler(cam, rp);
    const float scale = 1.0f / std::sqrt((float)proto.samples_per_pixel);
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (size_t i = 0; i < n; ++i) {
      SobolSampler s = proto;
      s.start_pixel(pixels_xy[2 * i], pixels_xy[2 * i + 1]);
      s.set_sample((size_t)sample_nums[i]);
      Vec2 pf = s.get_camera_sample();
      RayDifferential rd = generate_ray_differential(*cam, pf);
      rd.scale_differentials(scale);
      Spectrum l = path_li(sc, P, rd, &s, nullptr);
      out_rgb[3 * i] = l.r;
      out_rgb[3 * i + 1] = l.g;
      out_rgb[3 * i + 2] = l.b;
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// stats8: camera_paths, extension, shadow, mis, nodes, tris
int oracle_render(const PtrsSceneDesc* desc, const PtrsCamera* cam, const PtrsRenderParams* rp, float* film_rgbw, int n_threads,
                  int64_t tile_begin, int64_t tile_end, int64_t tile_stride, uint64_t* stats6) {
  try {
    RenderStats st = render(g_tables, desc, *cam, *rp, film_rgbw, n_threads, tile_begin, tile_end, tile_stride);
    if (stats6) {
      stats6[0] = st.camera_paths;
      stats6[1] = st.extension;
      stats6[2] = st.shadow;
      stats6[3] = st.mis;
      stats6[4] = st.nodes;
      stats6[5] = st.tris;
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int64_t oracle_tile_count(const PtrsCamera* cam, const PtrsRenderParams* rp) {
  Bounds2i sb = film_sample_bounds(cam->width, cam->height, rp->filter_radius);
  return (int64_t)((sb.x1 - sb.x0 + 15) / 16) * ((sb.y1 - sb.y0 + 15) / 16);
}

// Film::to_channel_updates / to_rgba_image, film.rs:230-271
void oracle_film_resolve(const float* rgbw, int w, int h, float* rgb) {
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    float inv_wt = 1.f / rgbw[4 * i + 3];
    rgb[3 * i] = rgbw[4 * i] * inv_wt;
    rgb[3 * i + 1] = rgbw[4 * i + 1] * inv_wt;
    rgb[3 * i + 2] = rgbw[4 * i + 2] * inv_wt;
  }
}

// ---- single-function probes for known-answer tests ----------------------------------------------
float oracle_gamma(uint32_t n) { return gamma(n); }
uint32_t oracle_log2_int(uint64_t i) { return log2_int(i); }
int oracle_solve_2x2(const float* a4, const float* b2, float* x2) {
  const float a[2][2] = {{a4[0], a4[1]}, {a4[2], a4[3]}};
  return solve_linear_system_2x2(a, b2, &x2[0], &x2[1]) ? 1 : 0;
}
float oracle_next_float_up(float v) { return next_float_up(v); }
float oracle_next_float_down(float v) { return next_float_down(v); }
uint64_t oracle_cantor_pairing(uint64_t x, uint64_t y) { return cantor_pairing(x, y); }
float oracle_fr_dielectric(float c, float ei, float et) { return fr_dielectric(c, ei, et); }
void oracle_cosine_sample_hemisphere(float u0, float u1, float* out3) {
  Vec3 v = cosine_sample_hemisphere(Vec2{u0, u1});
  std::memcpy(out3, &v, 12);
}
// lobe: 0 lambert, 1 microfacet-reflection(conductor), 2 fresnel-blend, 3 disney-diffuse, 4 microfacet-reflection(disney)
// p: r[3], t[3], ax, ay, eta[3], k[3], metallic, d_eta
static BxDF probe_lobe(int lobe, const float* p) {
  BxDF b;
  b.r = S(p[0], p[1], p[2]);
  b.t = S(p[3], p[4], p[5]);
  switch (lobe) {
    case 0: b.kind = BxDF::Lambertian; break;
    case 1:
      b.kind = BxDF::MicrofacetReflection;
      b.dist = Distribution::make(p[6], p[7], false);
      b.fresnel.kind = Fresnel::Conductor;
      b.fresnel.c_eta_i = S(1.f);
      b.fresnel.c_eta_t = S(p[8], p[9], p[10]);
      b.fresnel.c_k = S(p[11], p[12], p[13]);
      break;
    case 2:
      b.kind = BxDF::FresnelBlend;
      b.dist = Distribution::make(p[6], p[7], false);
      break;
    case 3: b.kind = BxDF::DisneyDiffuse; break;
    default:
      b.kind = BxDF::MicrofacetReflection;
      b.dist = Distribution::make(p[6], p[7], true);
      b.fresnel.kind = Fresnel::Disney;
      b.fresnel.r0 = S(p[3], p[4], p[5]);
      b.fresnel.metallic = p[14];
      b.fresnel.d_eta = p[15];
      break;
  }
  return b;
}
// out: f[3], pdf
void oracle_bxdf_eval(int lobe, const float* p, const float* wo3, const float* wi3, float* out4) {
  BxDF b = probe_lobe(lobe, p);
  Vec3 wo = V(wo3[0], wo3[1], wo3[2]), wi = V(wi3[0], wi3[1], wi3[2]);
  Spectrum f = b.f(wo, wi);
  out4[0] = f.r;
  out4[1] = f.g;
  out4[2] = f.b;
  out4[3] = b.pdf(wo, wi);
}
// out: wi[3], f[3], pdf
void oracle_bxdf_sample(int lobe, const float* p, const float* wo3, float u0, float u1, float* out7) {
  BxDF b = probe_lobe(lobe, p);
  Vec3 wo = V(wo3[0], wo3[1], wo3[2]), wi = V(0, 0, 0);
  float pdf = 0.f;
  Spectrum f = b.sample_f(wo, &wi, Vec2{u0, u1}, &pdf, nullptr);
  out7[0] = wi.x;
  out7[1] = wi.y;
  out7[2] = wi.z;
  out7[3] = f.r;
  out7[4] = f.g;
  out7[5] = f.b;
  out7[6] = pdf;
}

// ---- lobe / light probes mirroring ptrs_bxdf_eval / ptrs_bxdf_sample / ptrs_light_sample / ptrs_light_pdf --------
// PtrsLobeDesc -> the BxDF enum of bxdf/mod.rs:184-193 with the Fresnel / distribution objects the materials give it
static BxDF lobe_from_desc(const PtrsLobeDesc& d) {
  BxDF b;
  b.kind = (BxDF::Kind)d.kind;
  b.r = S(d.r[0], d.r[1], d.r[2]);
  b.t = S(d.t[0], d.t[1], d.t[2]);
  b.eta_a = d.eta_a;
  b.eta_b = d.eta_b;
  b.dist = Distribution::make(d.alpha_x, d.alpha_y, d.disney_g != 0);
  switch (d.fresnel) {
    case PTRS_FRESNEL_DIELECTRIC:
      b.fresnel.kind = Fresnel::Dielectric;
      b.fresnel.eta_i = d.eta_a;
      b.fresnel.eta_t = d.eta_b;
      break;
    case PTRS_FRESNEL_CONDUCTOR:
      b.fresnel.kind = Fresnel::Conductor;
      b.fresnel.c_eta_i = S(1.f);
      b.fresnel.c_eta_t = S(d.fa[0], d.fa[1], d.fa[2]);
      b.fresnel.c_k = S(d.fb[0], d.fb[1], d.fb[2]);
      break;
    case PTRS_FRESNEL_DISNEY:
      b.fresnel.kind = Fresnel::Disney;
      b.fresnel.r0 = S(d.fa[0], d.fa[1], d.fa[2]);
      b.fresnel.metallic = d.fb[0];
      b.fresnel.d_eta = d.fb[1];
      break;
    default: b.fresnel.kind = Fresnel::NoOp; break;
  }
  return b;
}
void oracle_lobe_eval(const PtrsLobeDesc* lobe, const float* wo, const float* wi, size_t n, float* out) {
  const BxDF b = lobe_from_desc(*lobe);
  for (size_t i = 0; i < n; ++i) {
    const Vec3 o = V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), w = V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
    const Spectrum f = b.f(o, w);
    out[4 * i] = f.r;
    out[4 * i + 1] = f.g;
    out[4 * i + 2] = f.b;
    out[4 * i + 3] = b.pdf(o, w);
  }
}
void oracle_lobe_sample(const PtrsLobeDesc* lobe, const float* wo, const float* u, size_t n, float* out) {
  const BxDF b = lobe_from_desc(*lobe);
  for (size_t i = 0; i < n; ++i) {
    const Vec3 o = V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    Vec3 w = V(0, 0, 0);
    float pdf = 0.f;
    uint32_t sampled = b.get_type();
    const Spectrum f = b.sample_f(o, &w, Vec2{u[2 * i], u[2 * i + 1]}, &pdf, &sampled);
    float* q = out + 8 * i;
    q[0] = w.x;
    q[1] = w.y;
    q[2] = w.z;
    q[3] = f.r;
    q[4] = f.g;
    q[5] = f.b;
    q[6] = pdf;
    q[7] = (float)sampled;
  }
}
// Light::sample_li + VisibilityTester's segment (light.rs, interaction.rs:50-59), reference p_error = 0
void oracle_light_sample(const PtrsSceneDesc* desc, int light, const float* ref_p, const float* ref_n, const float* u, size_t n, float* out) {
  Scene sc{desc};
  const PtrsLight& l = desc->lights[light];
  for (size_t i = 0; i < n; ++i) {
    Interaction ref;
    ref.p = V(ref_p[3 * i], ref_p[3 * i + 1], ref_p[3 * i + 2]);
    ref.n = V(ref_n[3 * i], ref_n[3 * i + 1], ref_n[3 * i + 2]);
    Vec3 wi = V(0, 0, 0);
    float pdf = 0.f;
    VisibilityTester vis;
    bool has_vis = true;
    const Spectrum li = light_sample_li(sc, l, ref, Vec2{u[2 * i], u[2 * i + 1]}, &wi, &pdf, &vis, &has_vis);
    float* q = out + 16 * i;
    for (int k = 0; k < 16; ++k) q[k] = 0.f;
    if (!has_vis) continue;  // the reference's caller would panic here (integrator.rs:51); the device reports zeros
    const Ray seg = vis.p0.spawn_ray_to_it(vis.p1);
    q[0] = li.r;
    q[1] = li.g;
    q[2] = li.b;
    q[3] = wi.x;
    q[4] = wi.y;
    q[5] = wi.z;
    q[6] = pdf;
    q[7] = seg.o.x;
    q[8] = seg.o.y;
    q[9] = seg.o.z;
    q[10] = seg.d.x;
    q[11] = seg.d.y;
    q[12] = seg.d.z;
  }
}
void oracle_light_pdf(const PtrsSceneDesc* desc, int light, const float* ref_p, const float* ref_n, const float* wi, size_t n, float* out) {
  Scene sc{desc};
  const PtrsLight& l = desc->lights[light];
  for (size_t i = 0; i < n; ++i) {
    Interaction ref;
    ref.p = V(ref_p[3 * i], ref_p[3 * i + 1], ref_p[3 * i + 2]);
    ref.n = V(ref_n[3 * i], ref_n[3 * i + 1], ref_n[3 * i + 2]);
    out[i] = light_pdf_li(sc, l, ref, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]));
  }
}

}  // extern "C"
