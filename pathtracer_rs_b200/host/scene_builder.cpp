#include "scene_builder.hpp"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace ptrs_host {

// ------------------------------------------------------------------------------------------------
// small helpers restating src/common/math.rs
// ------------------------------------------------------------------------------------------------
static inline int abs_mod_i(int a, int b) {  // math.rs:237-244
  int r = a - (a / b) * b;
  return r < 0 ? r + b : r;
}
static inline int round_up_pow2(int v) {  // math.rs:203-215
  v -= 1;
  v |= v >> 1;
  v |= v >> 2;
  v |= v >> 4;
  v |= v >> 8;
  v |= v >> 16;
  return v + 1;
}
static inline int log2_int(unsigned v) {  // math.rs:246-248
  int r = 0;
  while (v >>= 1) ++r;
  return r;
}
static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

bool invert(const M4& a, M4* out) {
  double m[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      m[i][j] = a.at(i, j);
      m[i][j + 4] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (std::fabs(m[r][c]) > std::fabs(m[piv][c])) piv = r;
    if (m[piv][c] == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 8; ++j) std::swap(m[piv][j], m[c][j]);
    double inv = 1.0 / m[c][c];
    for (int j = 0; j < 8; ++j) m[c][j] *= inv;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        double f = m[r][c];
        if (f != 0.0)
          for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out->at(i, j) = (float)m[i][j + 4];
  return true;
}

// ------------------------------------------------------------------------------------------------
// MIPMap (texture.rs:194-465)
// ------------------------------------------------------------------------------------------------
static float lanczos(float x, float tau) {  // texture.rs:199-211
  x = std::fabs(x);
  if (x < 1e-5f) return 1.f;
  if (x > 1.f) return 0.f;
  x *= 3.14159265358979323846f;
  float s = std::sin(x * tau) / (x * tau);
  float l = std::sin(x) / x;
  return s * l;
}
struct ResampleWeight {
  size_t first_texel;
  float weight[4];
};
static std::vector<ResampleWeight> resample_weights(int old_res, int new_res) {  // texture.rs:213-236
  std::vector<ResampleWeight> wt((size_t)new_res);
  const float filter_width = 2.f;
  for (int i = 0; i < new_res; ++i) {
    float center = ((float)i + 0.5f) * (float)old_res / (float)new_res;
    float first_texel = std::floor((center - filter_width) + 0.5f);
    float sum = 0.f;
    for (int j = 0; j < 4; ++j) {
      float pos = first_texel + (float)j + 0.5f;
      wt[i].weight[j] = lanczos((pos - center) / filter_width, 2.0f);
    }
    sum = wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3];
    float inv = 1.f / sum;
    for (int j = 0; j < 4; ++j) wt[i].weight[j] *= inv;
    wt[i].first_texel = first_texel > 0.f ? (size_t)first_texel : 0;  // `as usize` saturates at 0
  }
  return wt;
}

HostMipMap HostMipMap::build(const float* image, int w, int h, int channels, int wrap) {
  HostMipMap mm;
  mm.channels = channels;
  mm.wrap = wrap;
  const int C = channels;
  std::vector<float> base;
  int rw = w, rh = h;
  if (!is_pow2(w) || !is_pow2(h)) {
    const int pw = round_up_pow2(w), ph = round_up_pow2(h);
    std::vector<float> res((size_t)pw * ph * C, 0.f);
    auto sw = resample_weights(w, pw);
    for (int t = 0; t < h; ++t)
      for (int s = 0; s < pw; ++s)
        for (int j = 0; j < 4; ++j) {
          size_t orig_s = sw[s].first_texel + (size_t)j;
          if (wrap == PTRS_WRAP_REPEAT) orig_s = orig_s % (size_t)w;
          else if (wrap == PTRS_WRAP_CLAMP) orig_s = std::min(orig_s, (size_t)w - 1);
          if (orig_s > 0 && orig_s < (size_t)w)  // sic: texel 0 is skipped (texture.rs:320)
            for (int c = 0; c < C; ++c)
              res[((size_t)t * pw + s) * C + c] += image[((size_t)t * w + orig_s) * C + c] * sw[s].weight[j];
        }
    auto tw = resample_weights(h, ph);
    std::vector<float> work((size_t)ph * C);
    for (int s = 0; s < pw; ++s) {
      std::fill(work.begin(), work.end(), 0.f);
      for (int t = 0; t < ph; ++t)
        for (int j = 0; j < 4; ++j) {
          size_t off = tw[t].first_texel + (size_t)j;
          if (wrap == PTRS_WRAP_REPEAT) off = off % (size_t)h;
          else if (wrap == PTRS_WRAP_CLAMP) off = std::min(off, (size_t)h - 1);
          if (off < (size_t)h)
            for (int c = 0; c < C; ++c)
              work[(size_t)t * C + c] += res[(off * pw + s) * C + c] * tw[t].weight[j];
        }
      for (int t = 0; t < ph; ++t)
        for (int c = 0; c < C; ++c) res[((size_t)t * pw + s) * C + c] = work[(size_t)t * C + c];
    }
    base.swap(res);
    rw = pw;
    rh = ph;
  } else {
    base.assign(image, image + (size_t)w * h * C);
  }
  const int n_levels = 1 + log2_int((unsigned)std::max(rw, rh));
  if (n_levels > PTRS_MAX_MIP_LEVELS) throw std::runtime_error("mip pyramid too deep");
  mm.width.push_back(rw);
  mm.height.push_back(rh);
  mm.levels.push_back(std::move(base));
  for (int i = 1; i < n_levels; ++i) {
    const int sres = std::max(1, mm.width[i - 1] / 2), tres = std::max(1, mm.height[i - 1] / 2);
    std::vector<float> lvl((size_t)sres * tres * C);
    float a[3], b[3], c4[3], d[3];
    for (int t = 0; t < tres; ++t)
      for (int s = 0; s < sres; ++s) {
        mm.texel(i - 1, 2 * s, 2 * t, a);
        mm.texel(i - 1, 2 * s + 1, 2 * t, b);
        mm.texel(i - 1, 2 * s, 2 * t + 1, c4);
        mm.texel(i - 1, 2 * s + 1, 2 * t + 1, d);
        for (int c = 0; c < C; ++c) lvl[((size_t)t * sres + s) * C + c] = (((a[c] + b[c]) + c4[c]) + d[c]) * 0.25f;
      }
    mm.width.push_back(sres);
    mm.height.push_back(tres);
    mm.levels.push_back(std::move(lvl));
  }
  return mm;
}

void HostMipMap::texel(int level, int s, int t, float* out) const {
  const int W = width[level], H = height[level];
  if (wrap == PTRS_WRAP_REPEAT) {
    s = abs_mod_i(s, W);
    t = abs_mod_i(t, H);
  } else if (wrap == PTRS_WRAP_BLACK) {
    if (s < 0 || s >= W || t < 0 || t >= H) {
      for (int c = 0; c < channels; ++c) out[c] = 0.f;
      return;
    }
  } else {
    s = std::min(std::max(s, 0), W - 1);
    t = std::min(std::max(t, 0), H - 1);
  }
  const float* p = &levels[level][((size_t)t * W + s) * channels];
  for (int c = 0; c < channels; ++c) out[c] = p[c];
}

void HostMipMap::triangle(int level, float s_, float t_, float* out) const {
  level = std::min(std::max(level, 0), (int)levels.size() - 1);
  float s = s_ * (float)width[level] - 0.5f;
  float t = t_ * (float)height[level] - 0.5f;
  float s0f = std::floor(s), t0f = std::floor(t);
  float ds = s - s0f, dt = t - t0f;
  int s0 = (int)s0f, t0 = (int)t0f;
  float a[3], b[3], c[3], d[3];
  texel(level, s0, t0, a);
  texel(level, s0, t0 + 1, b);
  texel(level, s0 + 1, t0, c);
  texel(level, s0 + 1, t0 + 1, d);
  for (int k = 0; k < channels; ++k)
    out[k] = ((a[k] * (1.0f - ds) * (1.0f - dt) + b[k] * (1.0f - ds) * dt) + c[k] * ds * (1.0f - dt)) +
             d[k] * ds * dt;
}

void HostMipMap::lookup_width(float s, float t, float w, float* out) const {
  const int n = (int)levels.size();
  float level = (float)n - 1.0f + std::log2(std::fmax(w, 1e-8f));
  if (level < 0.0f) {
    triangle(0, s, t, out);
  } else if (level >= (float)(n - 1)) {
    triangle(n - 1, s, t, out);
  } else {
    float il = std::floor(level);
    float delta = level - il;
    float a[3], b[3];
    triangle((int)il, s, t, a);
    triangle((int)il + 1, s, t, b);
    for (int k = 0; k < channels; ++k) out[k] = a[k] * (1.0f - delta) + b[k] * delta;  // math::lerp
  }
}

// ------------------------------------------------------------------------------------------------
// Distribution2D (sampling.rs:128-209)
// ------------------------------------------------------------------------------------------------
static float build_1d(const float* f, int n, float* cdf /* n + 1 */) {
  cdf[0] = 0.f;
  for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (float)n;
  float func_int = cdf[n];
  if (func_int == 0.0f) {
    for (int i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
  } else {
    for (int i = 1; i < n + 1; ++i) cdf[i] /= func_int;
  }
  return func_int;
}

HostDistribution2D HostDistribution2D::build(const float* func, int nu, int nv) {
  HostDistribution2D d;
  d.nu = nu;
  d.nv = nv;
  d.cond_func.assign(func, func + (size_t)nu * nv);
  d.cond_cdf.resize((size_t)(nu + 1) * nv);
  d.cond_func_int.resize(nv);
  for (int v = 0; v < nv; ++v)
    d.cond_func_int[v] = build_1d(func + (size_t)v * nu, nu, &d.cond_cdf[(size_t)v * (nu + 1)]);
  d.marg_func = d.cond_func_int;
  d.marg_cdf.resize(nv + 1);
  d.marg_func_int = build_1d(d.marg_func.data(), nv, d.marg_cdf.data());
  return d;
}

// ------------------------------------------------------------------------------------------------
// SceneBuilder
// ------------------------------------------------------------------------------------------------
int SceneBuilder::add_constant_texture(int channels, float a, float b, float c) {
  PtrsTexture t{};
  t.type = PTRS_TEX_CONSTANT;
  t.channels = channels;
  t.v1[0] = a;
  t.v1[1] = channels == 1 ? a : b;
  t.v1[2] = channels == 1 ? a : c;
  t.su = t.sv = 1.f;
  t.mip = -1;
  textures_.push_back(t);
  return (int)textures_.size() - 1;
}
int SceneBuilder::add_checker_texture(int channels, const float v1[3], const float v2[3], float su,
                                      float sv, float du, float dv) {
  PtrsTexture t{};
  t.type = PTRS_TEX_CHECKER;
  t.channels = channels;
  std::memcpy(t.v1, v1, 12);
  std::memcpy(t.v2, v2, 12);
  t.su = su;
  t.sv = sv;
  t.du = du;
  t.dv = dv;
  t.mip = -1;
  textures_.push_back(t);
  return (int)textures_.size() - 1;
}
int SceneBuilder::push_mip(const HostMipMap& mm) {
  PtrsMipMap m{};
  m.channels = mm.channels;
  m.wrap = mm.wrap;
  m.n_levels = (int)mm.levels.size();
  for (int l = 0; l < m.n_levels; ++l) {
    m.width[l] = mm.width[l];
    m.height[l] = mm.height[l];
    m.level_offset[l] = texels_.size();
    texels_.insert(texels_.end(), mm.levels[l].begin(), mm.levels[l].end());
  }
  mipmaps_.push_back(m);
  return (int)mipmaps_.size() - 1;
}
int SceneBuilder::add_image_texture(int channels, const float* image, int width, int height, int wrap,
                                    float su, float sv, float du, float dv) {
  HostMipMap mm = HostMipMap::build(image, width, height, channels, wrap);
  PtrsTexture t{};
  t.type = PTRS_TEX_IMAGE;
  t.channels = channels;
  t.su = su;
  t.sv = sv;
  t.du = du;
  t.dv = dv;
  t.mip = push_mip(mm);
  textures_.push_back(t);
  tex_mips_.emplace((int)textures_.size() - 1, std::move(mm));
  return (int)textures_.size() - 1;
}
const HostMipMap* SceneBuilder::image_texture_mip(int texture) const {
  auto it = tex_mips_.find(texture);
  return it == tex_mips_.end() ? nullptr : &it->second;
}
int SceneBuilder::add_material(const PtrsMaterial& m) {
  materials_.push_back(m);
  return (int)materials_.size() - 1;
}
static PtrsMaterial mat(int type, int t0 = -1, int t1 = -1, int t2 = -1, int t3 = -1, int t4 = -1,
                        bool remap = false) {
  PtrsMaterial m{};
  m.type = type;
  m.normal_map = -1;
  m.tex[0] = t0;
  m.tex[1] = t1;
  m.tex[2] = t2;
  m.tex[3] = t3;
  m.tex[4] = t4;
  m.remap_roughness = remap ? 1 : 0;
  return m;
}
int SceneBuilder::add_matte(int kd) { return add_material(mat(PTRS_MAT_MATTE, kd)); }
int SceneBuilder::add_mirror() { return add_material(mat(PTRS_MAT_MIRROR)); }
int SceneBuilder::add_glass(int kr, int kt, int index) { return add_material(mat(PTRS_MAT_GLASS, kr, kt, index)); }
int SceneBuilder::add_metal(int eta, int k, int r, int ur, int vr, bool remap) {
  return add_material(mat(PTRS_MAT_METAL, eta, k, r, ur, vr, remap));
}
int SceneBuilder::add_substrate(int kd, int ks, int nu, int nv, bool remap) {
  return add_material(mat(PTRS_MAT_SUBSTRATE, kd, ks, nu, nv, -1, remap));
}
int SceneBuilder::add_disney(int color, int metallic, int eta, int roughness) {
  return add_material(mat(PTRS_MAT_DISNEY, color, metallic, eta, roughness));
}
void SceneBuilder::set_normal_map(int material, int normal_tex) { materials_.at(material).normal_map = normal_tex; }

int SceneBuilder::add_mesh(const MeshInput& mesh) {
  const size_t nv = mesh.pos.size() / 3, nt = mesh.indices.size() / 3;
  const uint32_t base = (uint32_t)(pos_.size() / 3);
  const bool hn = !mesh.normal.empty(), ht = !mesh.tangent.empty(), hu = !mesh.uv.empty();
  if ((hn && mesh.normal.size() != nv * 3) || (ht && mesh.tangent.size() != nv * 3) ||
      (hu && mesh.uv.size() != nv * 2))
    throw std::runtime_error("mesh attribute size mismatch");
  // TriangleMesh::new_with_transform (shape.rs:592-623): positions as points, normals and tangents
  // as plain vectors through the FORWARD matrix, not re-normalised.
  for (size_t i = 0; i < nv; ++i) {
    V3 p = xform_point(mesh.obj_to_world, v3(mesh.pos[3 * i], mesh.pos[3 * i + 1], mesh.pos[3 * i + 2]));
    pos_.push_back(p.x);
    pos_.push_back(p.y);
    pos_.push_back(p.z);
  }
  auto grow = [&](std::vector<float>& v, size_t per) { v.resize((pos_.size() / 3) * per, 0.f); };
  if (hn) any_normal_ = true;
  if (ht) any_tangent_ = true;
  if (hu) any_uv_ = true;
  grow(normal_, 3);
  grow(tangent_, 3);
  grow(uv_, 2);
  for (size_t i = 0; i < nv; ++i) {
    if (hn) {
      V3 n = xform_vector(mesh.obj_to_world, v3(mesh.normal[3 * i], mesh.normal[3 * i + 1], mesh.normal[3 * i + 2]));
      std::memcpy(&normal_[(base + i) * 3], &n, 12);
    }
    if (ht) {
      V3 s = xform_vector(mesh.obj_to_world, v3(mesh.tangent[3 * i], mesh.tangent[3 * i + 1], mesh.tangent[3 * i + 2]));
      std::memcpy(&tangent_[(base + i) * 3], &s, 12);
    }
    if (hu) {
      uv_[(base + i) * 2] = mesh.uv[2 * i];
      uv_[(base + i) * 2 + 1] = mesh.uv[2 * i + 1];
    }
  }
  PtrsMesh m{};
  m.flags = (hn ? PTRS_MESH_HAS_NORMAL : 0) | (ht ? PTRS_MESH_HAS_TANGENT : 0) | (hu ? PTRS_MESH_HAS_UV : 0);
  m.alpha_tex = mesh.alpha_tex;
  meshes_.push_back(m);
  const int mesh_id = (int)meshes_.size() - 1;
  for (size_t t = 0; t < nt; ++t) {
    const uint32_t tri = (uint32_t)(tri_vertex_.size() / 3);
    for (int k = 0; k < 3; ++k) {
      uint32_t idx = mesh.indices[3 * t + k];
      if (idx >= nv) throw std::runtime_error("mesh index out of range");
      tri_vertex_.push_back(base + idx);
    }
    tri_mesh_.push_back(mesh_id);
    tri_material_.push_back(mesh.material);
    int light = -1;
    if (!mesh.tri_emits.empty() && mesh.tri_emits.size() != nt) throw std::runtime_error("tri_emits size mismatch");
    if (mesh.ke_tex >= 0 && (mesh.tri_emits.empty() || mesh.tri_emits[t])) {  // importer/mitsuba.rs:309-323: one DiffuseAreaLight per triangle
      PtrsLight l{};
      l.type = PTRS_LIGHT_AREA;
      l.prim = (int32_t)tri;
      l.ke_tex = mesh.ke_tex;
      l.env = -1;
      lights_.push_back(l);
      light = (int)lights_.size() - 1;
    }
    tri_light_.push_back(light);
  }
  return mesh_id;
}

int SceneBuilder::add_point_light(const M4& l2w, const float intensity[3]) {
  PtrsLight l{};
  l.type = PTRS_LIGHT_POINT;
  l.prim = l.ke_tex = l.env = -1;
  V3 p = xform_point(l2w, v3(0, 0, 0));  // light.rs:94
  std::memcpy(l.pos, &p, 12);
  std::memcpy(l.color, intensity, 12);
  lights_.push_back(l);
  return (int)lights_.size() - 1;
}
int SceneBuilder::add_directional_light(const M4& l2w, const float lrgb[3], const float w[3]) {
  PtrsLight l{};
  l.type = PTRS_LIGHT_DIRECTIONAL;
  l.prim = l.ke_tex = l.env = -1;
  V3 d = normalize(xform_vector(l2w, v3(w[0], w[1], w[2])));  // light.rs:167
  std::memcpy(l.pos, &d, 12);
  std::memcpy(l.color, lrgb, 12);
  lights_.push_back(l);
  return (int)lights_.size() - 1;
}
int SceneBuilder::add_infinite_light(const M4& l2w, const float* rgb, int w, int h) {
  HostMipMap mm = HostMipMap::build(rgb, w, h, 3, PTRS_WRAP_REPEAT);
  // light.rs:372-387: the sampling distribution is tabulated at TWICE the map resolution
  const int width = 2 * w, height = 2 * h;  // 2 * texels.ncols() / nrows() of the ORIGINAL image
  const float f_width = 0.5f / (float)std::min(width, height);
  std::vector<float> img((size_t)width * height);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < height; ++v) {
    float vp = ((float)v + 0.5f) / (float)height;
    float sin_theta = std::sin(3.14159265358979323846f * vp);
    for (int u = 0; u < width; ++u) {
      float up = ((float)u + 0.5f) / (float)width;
      float c[3];
      mm.lookup_width(up, vp, f_width, c);
      float y = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f;  // spectrum.rs:112-115
      img[(size_t)v * width + u] = sin_theta * y;
    }
  }
  env_dists_.push_back(HostDistribution2D::build(img.data(), width, height));
  PtrsEnvLight e{};
  std::memcpy(e.light_to_world, l2w.m, 64);
  M4 inv;
  if (!invert(l2w, &inv)) throw std::runtime_error("singular light_to_world");
  std::memcpy(e.world_to_light, inv.m, 64);
  e.mip = push_mip(mm);
  e.nu = width;
  e.nv = height;
  envs_.push_back(e);
  PtrsLight l{};
  l.type = PTRS_LIGHT_INFINITE;
  l.prim = l.ke_tex = -1;
  l.env = (int)envs_.size() - 1;
  lights_.push_back(l);
  const int id = (int)lights_.size() - 1;
  infinite_lights_.push_back(id);  // pushed into BOTH lists (importer/mitsuba.rs:397-398)
  return id;
}

FlatScene SceneBuilder::finalize(int max_prims_in_node, int n_threads) {
  FlatScene fs;
  const size_t nt = tri_vertex_.size() / 3;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  std::vector<Bounds3> bounds(nt);
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (size_t t = 0; t < nt; ++t) {  // Triangle::world_bound, shape.rs:526-531
    Bounds3 b;
    for (int k = 0; k < 3; ++k) {
      float a = pos_[3 * tri_vertex_[3 * t] + k], c = pos_[3 * tri_vertex_[3 * t + 1] + k],
            d = pos_[3 * tri_vertex_[3 * t + 2] + k];
      b.mn[k] = std::fmin(std::fmin(a, c), d);
      b.mx[k] = std::fmax(std::fmax(a, c), d);
    }
    bounds[t] = b;
  }
  auto t0 = std::chrono::steady_clock::now();
  BvhBuildResult bvh = build_bvh(bounds, max_prims_in_node, n_threads);
  fs.bvh_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fs.bvh_max_depth = bvh.max_depth;
  fs.nodes = std::move(bvh.nodes);
  fs.prim_vertex.resize(nt * 3);
  fs.prim_mesh.resize(nt);
  fs.prim_material.resize(nt);
  fs.prim_area_light.resize(nt);
  std::vector<uint32_t> inverse(nt);
  for (size_t i = 0; i < nt; ++i) {
    const uint32_t src = bvh.prim_order[i];
    inverse[src] = (uint32_t)i;
    for (int k = 0; k < 3; ++k) fs.prim_vertex[3 * i + k] = tri_vertex_[3 * src + k];
    fs.prim_mesh[i] = tri_mesh_[src];
    fs.prim_material[i] = tri_material_[src];
    fs.prim_area_light[i] = tri_light_[src];
  }
  fs.pos = pos_;
  if (any_normal_) fs.normal = normal_;
  if (any_tangent_) fs.tangent = tangent_;
  if (any_uv_) fs.uv = uv_;
  fs.meshes = meshes_;
  fs.materials = materials_;
  fs.textures = textures_;
  fs.mipmaps = mipmaps_;
  fs.texels = texels_;
  fs.lights = lights_;
  fs.infinite_lights = infinite_lights_;
  fs.envs = envs_;
  fs.env_dists = env_dists_;

  // Light::preprocess(world_bound) (light.rs:220-222, 470-472; bounds.rs:126-134)
  float center[3] = {0, 0, 0}, radius = 0.f;
  if (!fs.nodes.empty()) {
    const PtrsBvhNode& r = fs.nodes[0];
    bool inside = true;
    for (int k = 0; k < 3; ++k) {
      center[k] = (r.bounds_min[k] + r.bounds_max[k]) * 0.5f;
      inside = inside && center[k] >= r.bounds_min[k] && center[k] <= r.bounds_max[k];
    }
    if (inside) {
      V3 d = v3(center[0] - r.bounds_max[0], center[1] - r.bounds_max[1], center[2] - r.bounds_max[2]);
      radius = norm(d);
    }
  }
  for (PtrsLight& l : fs.lights) {
    if (l.type == PTRS_LIGHT_AREA) {
      const uint32_t src = (uint32_t)l.prim;
      l.prim = (int32_t)inverse[src];
      V3 p0 = v3(pos_[3 * tri_vertex_[3 * src]], pos_[3 * tri_vertex_[3 * src] + 1], pos_[3 * tri_vertex_[3 * src] + 2]);
      V3 p1 = v3(pos_[3 * tri_vertex_[3 * src + 1]], pos_[3 * tri_vertex_[3 * src + 1] + 1], pos_[3 * tri_vertex_[3 * src + 1] + 2]);
      V3 p2 = v3(pos_[3 * tri_vertex_[3 * src + 2]], pos_[3 * tri_vertex_[3 * src + 2] + 1], pos_[3 * tri_vertex_[3 * src + 2] + 2]);
      l.area = 0.5f * norm(cross(p1 - p0, p2 - p0));  // shape.rs:533-539
    } else if (l.type == PTRS_LIGHT_DIRECTIONAL || l.type == PTRS_LIGHT_INFINITE) {
      std::memcpy(l.world_center, center, 12);
      l.world_radius = radius;
    }
  }
  return fs;
}

PtrsSceneDesc FlatScene::desc() const {
  PtrsSceneDesc d{};
  d.abi_version = PTRS_ABI_VERSION;
  d.n_nodes = (uint32_t)nodes.size();
  d.nodes = nodes.data();
  d.n_prims = (uint32_t)prim_mesh.size();
  d.prim_vertex = prim_vertex.data();
  d.prim_mesh = prim_mesh.data();
  d.prim_material = prim_material.data();
  d.prim_area_light = prim_area_light.data();
  d.n_verts = (uint32_t)(pos.size() / 3);
  d.pos = pos.data();
  d.normal = normal.empty() ? nullptr : normal.data();
  d.tangent = tangent.empty() ? nullptr : tangent.data();
  d.uv = uv.empty() ? nullptr : uv.data();
  d.n_meshes = (uint32_t)meshes.size();
  d.meshes = meshes.data();
  d.n_materials = (uint32_t)materials.size();
  d.materials = materials.data();
  d.n_textures = (uint32_t)textures.size();
  d.textures = textures.data();
  d.n_mipmaps = (uint32_t)mipmaps.size();
  d.mipmaps = mipmaps.data();
  d.n_texels = texels.size();
  d.texels = texels.data();
  d.n_lights = (uint32_t)lights.size();
  d.lights = lights.data();
  d.n_infinite_lights = (uint32_t)infinite_lights.size();
  d.infinite_lights = infinite_lights.data();
  d.n_envs = (uint32_t)envs.size();
  // env pointers are patched lazily: envs is mutable storage owned by this object
  FlatScene* self = const_cast<FlatScene*>(this);
  for (size_t i = 0; i < envs.size(); ++i) {
    const HostDistribution2D& hd = env_dists[i];
    PtrsEnvLight& e = self->envs[i];
    e.cond_func = hd.cond_func.data();
    e.cond_cdf = hd.cond_cdf.data();
    e.cond_func_int = hd.cond_func_int.data();
    e.marg_func = hd.marg_func.data();
    e.marg_cdf = hd.marg_cdf.data();
    e.marg_func_int = hd.marg_func_int;
  }
  d.envs = envs.data();
  return d;
}

PtrsSceneDesc FlatScene::desc_device_tables() const {
  PtrsSceneDesc d = desc();
  if (slim_mipmaps.size() != mipmaps.size()) {
    slim_mipmaps.clear();
    slim_texels.clear();
    for (const PtrsMipMap& m : mipmaps) {
      PtrsMipMap o{};
      o.channels = m.channels;
      o.wrap = m.wrap;
      o.n_levels = 1;
      o.width[0] = m.width[0];
      o.height[0] = m.height[0];
      o.level_offset[0] = slim_texels.size();
      const size_t n = (size_t)m.width[0] * m.height[0] * m.channels;
      slim_texels.insert(slim_texels.end(), texels.begin() + m.level_offset[0], texels.begin() + m.level_offset[0] + n);
      slim_mipmaps.push_back(o);
    }
  }
  slim_envs = envs;
  for (PtrsEnvLight& e : slim_envs) {
    e.cond_func = e.cond_cdf = e.cond_func_int = e.marg_func = e.marg_cdf = nullptr;
    e.marg_func_int = 0.f;
  }
  d.mipmaps = slim_mipmaps.data();
  d.n_texels = slim_texels.size();
  d.texels = slim_texels.data();
  d.envs = slim_envs.data();
  return d;
}

uint64_t FlatScene::host_bytes_device_tables() const {
  uint64_t b = host_bytes();
  desc_device_tables();
  b -= (texels.size() - slim_texels.size()) * 4;
  for (const auto& d : env_dists)
    b -= (d.cond_func.size() + d.cond_cdf.size() + d.cond_func_int.size() + d.marg_func.size() + d.marg_cdf.size()) * 4;
  return b;
}

uint64_t FlatScene::host_bytes() const {
  uint64_t b = nodes.size() * sizeof(PtrsBvhNode) + prim_vertex.size() * 4 + prim_mesh.size() * 12 +
               (pos.size() + normal.size() + tangent.size() + uv.size() + texels.size()) * 4 +
               meshes.size() * sizeof(PtrsMesh) + materials.size() * sizeof(PtrsMaterial) +
               textures.size() * sizeof(PtrsTexture) + mipmaps.size() * sizeof(PtrsMipMap) +
               lights.size() * sizeof(PtrsLight) + envs.size() * sizeof(PtrsEnvLight);
  for (const auto& d : env_dists)
    b += (d.cond_func.size() + d.cond_cdf.size() + d.cond_func_int.size() + d.marg_func.size() + d.marg_cdf.size()) * 4;
  return b;
}

// ------------------------------------------------------------------------------------------------
// Camera / film parameters
// ------------------------------------------------------------------------------------------------
void quat_from_matrix(const float r[9], float q[4]) {
  const float m00 = r[0], m01 = r[1], m02 = r[2], m10 = r[3], m11 = r[4], m12 = r[5], m20 = r[6],
              m21 = r[7], m22 = r[8];
  float tr = m00 + m11 + m22, x, y, z, w;
  if (tr > 0.f) {
    float s = std::sqrt(tr + 1.0f) * 2.f;
    w = 0.25f * s;
    x = (m21 - m12) / s;
    y = (m02 - m20) / s;
    z = (m10 - m01) / s;
  } else if (m00 > m11 && m00 > m22) {
    float s = std::sqrt(1.0f + m00 - m11 - m22) * 2.f;
    w = (m21 - m12) / s;
    x = 0.25f * s;
    y = (m01 + m10) / s;
    z = (m02 + m20) / s;
  } else if (m11 > m22) {
    float s = std::sqrt(1.0f + m11 - m00 - m22) * 2.f;
    w = (m02 - m20) / s;
    x = (m01 + m10) / s;
    y = 0.25f * s;
    z = (m12 + m21) / s;
  } else {
    float s = std::sqrt(1.0f + m22 - m00 - m11) * 2.f;
    w = (m10 - m01) / s;
    x = (m02 + m20) / s;
    y = (m12 + m21) / s;
    z = 0.25f * s;
  }
  float n = std::sqrt(x * x + y * y + z * z + w * w);
  q[0] = x / n;
  q[1] = y / n;
  q[2] = z / n;
  q[3] = w / n;
}

PtrsCamera make_camera(const float q[4], const float trans[3], float aspect, float fovy, float znear,
                       float zfar, int width, int height) {
  PtrsCamera c{};
  std::memcpy(c.rot, q, 16);
  std::memcpy(c.trans, trans, 12);
  c.width = width;
  c.height = height;
  // Perspective3::new (nalgebra 0.32): m11 = 1/tan(fovy/2); m00 = m11/aspect;
  // m22 = (far+near)/(near-far); m23 = 2*far*near/(near-far)
  const float m11 = 1.0f / std::tan(fovy / 2.0f);
  const float m00 = m11 / aspect;
  const float m22 = (zfar + znear) / (znear - zfar);
  const float m23 = zfar * znear * 2.0f / (znear - zfar);
  c.persp[0] = m00;
  c.persp[1] = m11;
  c.persp[2] = m22;
  c.persp[3] = m23;
  // screen_to_raster = scaling(res.x, res.y, 1) * scaling(1/2, -1/2, 1) * translation(1, -1, 0)
  const float rx = (float)width, ry = (float)height;
  M4 s2r = M4::identity();
  s2r.at(0, 0) = rx * (1.0f / 2.0f);
  s2r.at(1, 1) = ry * (1.0f / -2.0f);
  s2r.at(0, 3) = s2r.at(0, 0) * 1.0f;
  s2r.at(1, 3) = s2r.at(1, 1) * -1.0f;
  M4 r2s;
  invert(s2r, &r2s);
  std::memcpy(c.raster_to_screen, r2s.m, 64);
  // raster_to_camera = cam_to_screen.to_projective().inverse() * raster_to_screen
  M4 persp{};
  persp.at(0, 0) = m00;
  persp.at(1, 1) = m11;
  persp.at(2, 2) = m22;
  persp.at(2, 3) = m23;
  persp.at(3, 2) = -1.0f;
  M4 pinv;
  invert(persp, &pinv);
  M4 r2c = pinv * r2s;
  V3 o = xform_point(r2c, v3(0, 0, 0));
  V3 dx = xform_point(r2c, v3(1, 0, 0)) - o;
  V3 dy = xform_point(r2c, v3(0, 1, 0)) - o;
  std::memcpy(c.dx_camera, &dx, 12);
  std::memcpy(c.dy_camera, &dy, 12);
  return c;
}

PtrsCamera mitsuba_camera(const M4& sensor, float fov_deg, int film_w, int film_h, int res_w, int res_h) {
  const float fov = fov_deg * (3.14159265358979323846f / 180.0f);
  // Rotation3::new((0, -pi, 0)): axis-angle, angle pi about -Y  (right-to-left-hand fix-up)
  const float ang = -3.14159265358979323846f;
  const float cs = std::cos(ang), sn = std::sin(ang);
  M4 rot = M4::identity();
  rot.at(0, 0) = cs;
  rot.at(0, 2) = sn;
  rot.at(2, 0) = -sn;
  rot.at(2, 2) = cs;
  M4 m = sensor * rot;
  float r[9] = {m.at(0, 0), m.at(0, 1), m.at(0, 2), m.at(1, 0), m.at(1, 1), m.at(1, 2), m.at(2, 0), m.at(2, 1), m.at(2, 2)};
  // remove any (uniform) scale like the Similarity3 -> Isometry3 step does
  float sc = std::cbrt(std::fabs(r[0] * (r[4] * r[8] - r[5] * r[7]) - r[1] * (r[3] * r[8] - r[5] * r[6]) +
                                 r[2] * (r[3] * r[7] - r[4] * r[6])));
  if (sc > 0.f)
    for (float& v : r) v /= sc;
  float q[4];
  quat_from_matrix(r, q);
  float t[3] = {m.at(0, 3), m.at(1, 3), m.at(2, 3)};
  return make_camera(q, t, (float)res_w / (float)res_h, fov * ((float)film_h / (float)film_w), 0.01f,
                     10000.0f, res_w, res_h);
}

void gaussian_filter_table(float alpha, float radius, float table[256]) {
  const float expv = std::exp(-alpha * radius * radius);
  auto g = [&](float d) { return std::fmax(0.0f, std::exp(-alpha * d * d) - expv); };
  int off = 0;
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 16; ++x) {
      float px = ((float)x + 0.5f) * radius / 16.0f;
      float py = ((float)y + 0.5f) * radius / 16.0f;
      table[off++] = g(px) * g(py);
    }
}

void default_render_params(PtrsRenderParams* p) {
  std::memset(p, 0, sizeof(*p));
  p->spp = 1;          // main.rs default
  p->max_depth = 15;   // main.rs default
  p->rr_threshold = 1.0f;
  p->rr_start_depth = 3;
  p->rr_enable = 1;
  p->sample_begin = 0;
  p->sample_end = 0;
  p->sample_stride = 1;
  p->sample_phase = 0;
  p->filter_radius[0] = p->filter_radius[1] = 2.0f;
  gaussian_filter_table(2.0f, 2.0f, p->filter_table);
}

}  // namespace ptrs_host
