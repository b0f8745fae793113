#include "procedural.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace ptrs_host {

static const float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------------------------------------------
// genmesh 0.6.2 generators (crate not vendored with the reference; vertex and triangle order
// follow the crate's documented Plane / Cube generators; Quad(x,y,z,w) triangulates to (x,y,z),(x,z,w))
// ------------------------------------------------------------------------------------------------
MeshInput gen_rectangle() {
  MeshInput m;
  const float p[4][2] = {{-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
  for (auto& v : p) {
    m.pos.insert(m.pos.end(), {v[0], v[1], 0.f});
    m.normal.insert(m.normal.end(), {0.f, 0.f, 1.f});
  }
  m.indices = {0, 1, 3, 0, 3, 2};  // Quad(0, 1, 3, 2)
  return m;
}

MeshInput gen_cube() {
  MeshInput m;
  static const int faces[6][4] = {{6, 7, 5, 4}, {0, 1, 3, 2}, {3, 7, 6, 2}, {4, 5, 1, 0}, {5, 7, 3, 1}, {0, 2, 6, 4}};
  static const float normals[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  for (int f = 0; f < 6; ++f) {
    for (int k = 0; k < 4; ++k) {
      int idx = faces[f][k];
      m.pos.insert(m.pos.end(), {(idx & 4) ? 1.f : -1.f, (idx & 2) ? 1.f : -1.f, (idx & 1) ? 1.f : -1.f});
      m.normal.insert(m.normal.end(), {normals[f][0], normals[f][1], normals[f][2]});
    }
    uint32_t b = (uint32_t)f * 4;
    m.indices.insert(m.indices.end(), {b, b + 1, b + 2, b, b + 2, b + 3});
  }
  return m;
}

MeshInput gen_sphere_uv(int nu, int nv, bool with_uv) {
  MeshInput m;
  for (int j = 0; j <= nv; ++j) {
    float th = kPi * (float)j / (float)nv;
    for (int i = 0; i <= nu; ++i) {
      float ph = 2.f * kPi * (float)i / (float)nu;
      float x = std::sin(th) * std::cos(ph), y = std::cos(th), z = std::sin(th) * std::sin(ph);
      m.pos.insert(m.pos.end(), {x, y, z});
      m.normal.insert(m.normal.end(), {x, y, z});
      if (with_uv) m.uv.insert(m.uv.end(), {(float)i / (float)nu, (float)j / (float)nv});
    }
  }
  for (int j = 0; j < nv; ++j)
    for (int i = 0; i < nu; ++i) {
      uint32_t a = (uint32_t)(j * (nu + 1) + i), b = a + 1, c = a + (uint32_t)nu + 1, d = c + 1;
      if (j != 0) m.indices.insert(m.indices.end(), {a, b, c});
      if (j != nv - 1) m.indices.insert(m.indices.end(), {b, d, c});
    }
  return m;
}

static MeshInput gen_grid(int nx, int ny, bool with_uv, bool with_normal,
                          V3 (*f)(float, float, const void*), const void* ctx) {
  MeshInput m;
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i <= nx; ++i) {
      float u = (float)i / (float)nx, v = (float)j / (float)ny;
      V3 p = f(u, v, ctx);
      m.pos.insert(m.pos.end(), {p.x, p.y, p.z});
      if (with_uv) m.uv.insert(m.uv.end(), {u, v});
      if (with_normal) {
        const float e = 0.25f / (float)std::max(nx, ny);
        V3 du = f(u + e, v, ctx) - f(u - e, v, ctx), dv = f(u, v + e, ctx) - f(u, v - e, ctx);
        V3 n = normalize(cross(dv, du));
        m.normal.insert(m.normal.end(), {n.x, n.y, n.z});
      }
    }
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      uint32_t a = (uint32_t)(j * (nx + 1) + i), b = a + 1, c = a + (uint32_t)nx + 1, d = c + 1;
      m.indices.insert(m.indices.end(), {a, c, b, b, c, d});
    }
  return m;
}

static M4 trs(V3 t, V3 s, float yaw = 0.f) {
  M4 m = M4::identity();
  float c = std::cos(yaw), sn = std::sin(yaw);
  m.at(0, 0) = c * s.x;
  m.at(0, 2) = sn * s.z;
  m.at(2, 0) = -sn * s.x;
  m.at(2, 2) = c * s.z;
  m.at(1, 1) = s.y;
  m.at(0, 3) = t.x;
  m.at(1, 3) = t.y;
  m.at(2, 3) = t.z;
  return m;
}

// ------------------------------------------------------------------------------------------------
// Cornell box (values of data/cornell-box.xml)
// ------------------------------------------------------------------------------------------------
static M4 m4(std::initializer_list<float> v) {
  M4 m{};
  int i = 0;
  for (float x : v) m.m[i++] = x;
  return m;
}

M4 mitsuba_env_light_to_world() {
  // Matrix4::from_euler_angles(-pi/2, -pi/2, 0).append_nonuniform_scaling((1, 1, -1))
  const float roll = -kPi / 2.f, pitch = -kPi / 2.f, yaw = 0.f;
  float sr = std::sin(roll), cr = std::cos(roll), sp = std::sin(pitch), cp = std::cos(pitch),
        sy = std::sin(yaw), cy = std::cos(yaw);
  M4 m = M4::identity();
  m.at(0, 0) = cy * cp;
  m.at(0, 1) = cy * sp * sr - sy * cr;
  m.at(0, 2) = cy * sp * cr + sy * sr;
  m.at(1, 0) = sy * cp;
  m.at(1, 1) = sy * sp * sr + cy * cr;
  m.at(1, 2) = sy * sp * cr - cy * sr;
  m.at(2, 0) = -sp;
  m.at(2, 1) = cp * sr;
  m.at(2, 2) = cp * cr;
  for (int c = 0; c < 4; ++c) m.at(2, c) *= -1.0f;
  return m;
}

void build_cornell(SceneBuilder& b, const float* env_rgb, int env_w, int env_h) {
  const float refl[8][3] = {{0.63f, 0.065f, 0.05f},  {0.14f, 0.45f, 0.091f}, {0.725f, 0.71f, 0.68f},
                            {0.725f, 0.71f, 0.68f},  {0.725f, 0.71f, 0.68f}, {0.725f, 0.71f, 0.68f},
                            {0.725f, 0.71f, 0.68f},  {0.f, 0.f, 0.f}};
  enum { LeftWall, RightWall, Floor, Ceiling, BackWall, ShortBox, TallBox, Light };
  int mats[8];
  for (int i = 0; i < 8; ++i) mats[i] = b.add_matte(b.add_constant_texture(3, refl[i][0], refl[i][1], refl[i][2]));
  struct Shape { bool cube; M4 xf; int mat; bool emit; };
  const Shape shapes[] = {
      {false, m4({-4.37114e-008f, 1, 4.37114e-008f, 0, 0, -8.74228e-008f, 2, 0, 1, 4.37114e-008f, 1.91069e-015f, 0, 0, 0, 0, 1}), Floor, false},
      {false, m4({-1, 7.64274e-015f, -1.74846e-007f, 0, 8.74228e-008f, 8.74228e-008f, -2, 2, 0, -1, -4.37114e-008f, 0, 0, 0, 0, 1}), Ceiling, false},
      {false, m4({1.91069e-015f, 1, 1.31134e-007f, 0, 1, 3.82137e-015f, -8.74228e-008f, 1, -4.37114e-008f, 1.31134e-007f, -2, -1, 0, 0, 0, 1}), BackWall, false},
      {false, m4({4.37114e-008f, -1.74846e-007f, 2, 1, 1, 3.82137e-015f, -8.74228e-008f, 1, 3.82137e-015f, 1, 2.18557e-007f, 0, 0, 0, 0, 1}), RightWall, false},
      {false, m4({-4.37114e-008f, 8.74228e-008f, -2, -1, 1, 3.82137e-015f, -8.74228e-008f, 1, 0, -1, -4.37114e-008f, 0, 0, 0, 0, 1}), LeftWall, false},
      {true, m4({0.0851643f, 0.289542f, 1.31134e-008f, 0.328631f, 3.72265e-009f, 1.26563e-008f, -0.3f, 0.3f, -0.284951f, 0.0865363f, 5.73206e-016f, 0.374592f, 0, 0, 0, 1}), ShortBox, false},
      {true, m4({0.286776f, 0.098229f, -2.29282e-015f, -0.335439f, -4.36233e-009f, 1.23382e-008f, -0.6f, 0.6f, -0.0997984f, 0.282266f, 2.62268e-008f, -0.291415f, 0, 0, 0, 1}), TallBox, false},
      {false, m4({0.235f, -1.66103e-008f, -7.80685e-009f, -0.005f, -2.05444e-008f, 3.90343e-009f, -0.0893f, 1.98f, 2.05444e-008f, 0.19f, 8.30516e-009f, -0.03f, 0, 0, 0, 1}), Light, true},
  };
  for (const Shape& s : shapes) {
    MeshInput m = s.cube ? gen_cube() : gen_rectangle();
    m.obj_to_world = s.xf;
    m.material = mats[s.mat];
    if (s.emit) m.ke_tex = b.add_constant_texture(3, 17.f, 12.f, 4.f);
    b.add_mesh(m);
  }
  if (env_rgb) b.add_infinite_light(mitsuba_env_light_to_world(), env_rgb, env_w, env_h);
}

PtrsCamera cornell_camera(int res_w, int res_h) {
  M4 sensor = m4({-1, 0, 0, 0, 0, 1, 0, 1, 0, 0, -1, 6.8f, 0, 0, 0, 1});
  return mitsuba_camera(sensor, 19.5f, 1024, 1024, res_w, res_h);
}

PtrsCamera look_at_camera(V3 eye, V3 target, V3 up, float fovy_deg, int res_w, int res_h) {
  V3 f = normalize(target - eye);
  V3 r = normalize(cross(f, up));
  V3 u = cross(r, f);
  float rot[9] = {r.x, u.x, -f.x, r.y, u.y, -f.y, r.z, u.z, -f.z};
  float q[4];
  quat_from_matrix(rot, q);
  float t[3] = {eye.x, eye.y, eye.z};
  return make_camera(q, t, (float)res_w / (float)res_h, fovy_deg * (kPi / 180.f), 0.01f, 10000.f, res_w, res_h);
}

// ------------------------------------------------------------------------------------------------
// Synthetic HDR sky
// ------------------------------------------------------------------------------------------------
static float value_noise(float x, float y, uint32_t seed) {
  auto h = [&](int xi, int yi) {
    uint32_t n = (uint32_t)xi * 374761393u + (uint32_t)yi * 668265263u + seed * 2246822519u;
    n = (n ^ (n >> 13)) * 1274126177u;
    n ^= n >> 16;
    return (float)(n & 0xffffff) * (1.0f / 16777216.0f);
  };
  int x0 = (int)std::floor(x), y0 = (int)std::floor(y);
  float fx = x - (float)x0, fy = y - (float)y0;
  fx = fx * fx * (3.f - 2.f * fx);
  fy = fy * fy * (3.f - 2.f * fy);
  float a = h(x0, y0), b = h(x0 + 1, y0), c = h(x0, y0 + 1), d = h(x0 + 1, y0 + 1);
  return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy;
}

std::vector<float> synth_sky(int w, int h, uint64_t seed) {
  std::vector<float> img((size_t)w * h * 3);
  const float sun_theta = 0.30f * kPi, sun_phi = 0.6f * kPi;
  const V3 sun = v3(std::sin(sun_theta) * std::cos(sun_phi), std::sin(sun_theta) * std::sin(sun_phi), std::cos(sun_theta));
  for (int y = 0; y < h; ++y) {
    float th = kPi * ((float)y + 0.5f) / (float)h;
    for (int x = 0; x < w; ++x) {
      float ph = 2.f * kPi * ((float)x + 0.5f) / (float)w;
      V3 d = v3(std::sin(th) * std::cos(ph), std::sin(th) * std::sin(ph), std::cos(th));
      float up = d.z;
      float r, g, b;
      if (up > 0.f) {
        float t = std::pow(1.f - up, 3.f);
        r = 0.25f + 0.9f * t;
        g = 0.45f + 0.75f * t;
        b = 0.95f + 0.2f * t;
        float n = 0.f, amp = 0.5f, fr = 6.f;
        for (int o = 0; o < 4; ++o) {
          n += amp * value_noise(ph * fr, th * fr * 2.f, (uint32_t)seed + (uint32_t)o);
          amp *= 0.5f;
          fr *= 2.f;
        }
        float cloud = std::max(0.f, n - 0.45f) * 3.f;
        r += cloud;
        g += cloud;
        b += cloud;
      } else {
        float t = std::min(1.f, -up * 4.f);
        r = 0.18f * (1 - t) + 0.05f * t + 0.2f * (1 - t);
        g = 0.16f * (1 - t) + 0.05f * t + 0.2f * (1 - t);
        b = 0.12f * (1 - t) + 0.04f * t + 0.2f * (1 - t);
      }
      float cs = dot(d, sun);
      if (cs > 0.9994f) {  // ~2 degree disc
        r += 4000.f;
        g += 3600.f;
        b += 3000.f;
      } else if (cs > 0.f) {
        float gl = std::pow(cs, 256.f) * 12.f + std::pow(cs, 16.f) * 0.8f;
        r += gl;
        g += gl * 0.9f;
        b += gl * 0.7f;
      }
      float* p = &img[((size_t)y * w + x) * 3];
      p[0] = r;
      p[1] = g;
      p[2] = b;
    }
  }
  return img;
}

// ------------------------------------------------------------------------------------------------
// C3: material field
// ------------------------------------------------------------------------------------------------
void build_material_field(SceneBuilder& b, uint64_t seed, size_t n_tris, PtrsCamera* cam, int res_w, int res_h) {
  Pcg32 rng(seed, 3);
  // tessellation: each object ~2k triangles (32 x 32 lat-long sphere = 1984)
  const int su = 32, sv = 32;
  const size_t per_obj = (size_t)2 * su * (sv - 1);
  int g = (int)std::ceil(std::sqrt((double)n_tris / (double)per_obj));
  g = std::max(g, 2);
  const float spacing = 2.6f;
  const float half = 0.5f * spacing * (float)g;
  // ground: checker matte
  float c0[3] = {0.8f, 0.8f, 0.8f}, c1[3] = {0.2f, 0.25f, 0.3f};
  int ground_mat = b.add_matte(b.add_checker_texture(3, c0, c1, (float)g, (float)g, 0.f, 0.f));
  {
    MeshInput m = gen_rectangle();
    m.uv = {0, 0, 1, 0, 0, 1, 1, 1};
    M4 xf = M4::identity();  // rectangle in XY -> XZ plane, normal +Y
    xf.at(0, 0) = half * 1.5f;
    xf.at(1, 1) = 0.f;
    xf.at(1, 2) = 1.f;
    xf.at(2, 1) = -half * 1.5f;
    xf.at(2, 2) = 0.f;
    m.obj_to_world = xf;
    m.material = ground_mat;
    b.add_mesh(m);
  }
  int white = b.add_constant_texture(3, 1.f, 1.f, 1.f);
  const float metals[2][2][3] = {{{0.2004f, 0.9240f, 1.1022f}, {3.9129f, 2.4528f, 2.1421f}},   // Cu
                                 {{0.1431f, 0.3749f, 1.4424f}, {3.9831f, 2.3857f, 1.6032f}}};  // Au
  size_t made = 2;
  int obj = 0;
  for (int gz = 0; gz < g && made < n_tris; ++gz)
    for (int gx = 0; gx < g && made < n_tris; ++gx, ++obj) {
      const int kind = obj % 5;
      float hue[3] = {0.2f + 0.7f * rng.next_f32(), 0.2f + 0.7f * rng.next_f32(), 0.2f + 0.7f * rng.next_f32()};
      float alpha = 0.01f + 0.29f * rng.next_f32();
      int matid;
      if (kind == 0) {
        matid = b.add_glass(white, white, b.add_constant_texture(1, 1.5f));
      } else if (kind == 1) {
        int a = b.add_constant_texture(1, alpha);
        matid = b.add_substrate(b.add_constant_texture(3, hue[0], hue[1], hue[2]),
                                b.add_constant_texture(3, 0.04f, 0.04f, 0.04f), a, a, false);
      } else if (kind == 2) {
        const auto& mt = metals[obj / 5 % 2];
        int a = b.add_constant_texture(1, alpha);
        matid = b.add_metal(b.add_constant_texture(3, mt[0][0], mt[0][1], mt[0][2]),
                            b.add_constant_texture(3, mt[1][0], mt[1][1], mt[1][2]), white, a, a, false);
      } else if (kind == 3) {
        float metallic = (obj / 5) % 2 ? 1.f : 0.f;
        float rough = 0.1f + 0.7f * rng.next_f32();
        matid = b.add_disney(b.add_constant_texture(3, hue[0], hue[1], hue[2]), b.add_constant_texture(1, metallic),
                             b.add_constant_texture(1, 1.5f), b.add_constant_texture(1, rough));
      } else {
        matid = b.add_matte(b.add_constant_texture(3, hue[0], hue[1], hue[2]));
      }
      float radius = 0.6f + 0.5f * rng.next_f32();
      V3 c = v3(-half + spacing * ((float)gx + 0.5f) + 0.3f * (rng.next_f32() - 0.5f), radius,
                -half + spacing * ((float)gz + 0.5f) + 0.3f * (rng.next_f32() - 0.5f));
      MeshInput m = gen_sphere_uv(su, sv, true);
      m.obj_to_world = trs(c, v3(radius, radius * (0.7f + 0.6f * rng.next_f32()), radius), rng.next_f32() * 6.28f);
      m.material = matid;
      b.add_mesh(m);
      made += m.indices.size() / 3;
    }
  // a few emissive quads above the field
  int black = b.add_matte(b.add_constant_texture(3, 0.f, 0.f, 0.f));
  for (int i = 0; i < 4; ++i) {
    MeshInput m = gen_rectangle();
    M4 xf = M4::identity();  // face down (-Y)
    xf.at(0, 0) = 1.5f;
    xf.at(1, 1) = 0.f;
    xf.at(1, 2) = -1.f;
    xf.at(2, 1) = 1.5f;
    xf.at(2, 2) = 0.f;
    xf.at(0, 3) = (i % 2 ? 0.5f : -0.5f) * half;
    xf.at(1, 3) = 6.f;
    xf.at(2, 3) = (i / 2 ? 0.5f : -0.5f) * half;
    m.obj_to_world = xf;
    m.material = black;
    m.ke_tex = b.add_constant_texture(3, 30.f, 26.f, 20.f);
    b.add_mesh(m);
  }
  std::vector<float> sky = synth_sky(1024, 512, seed);
  b.add_infinite_light(mitsuba_env_light_to_world(), sky.data(), 1024, 512);
  if (cam) *cam = look_at_camera(v3(0.f, half * 0.55f, half * 1.35f), v3(0.f, 0.5f, 0.f), v3(0, 1, 0), 40.f, res_w, res_h);
}

// ------------------------------------------------------------------------------------------------
// C4: terrain in the unit cube
// ------------------------------------------------------------------------------------------------
struct TerrainCtx { uint32_t seed; };
static V3 terrain_fn(float u, float v, const void* ctx) {
  const TerrainCtx* t = (const TerrainCtx*)ctx;
  float hgt = 0.f, amp = 0.18f, fr = 3.f;
  for (int o = 0; o < 6; ++o) {
    hgt += amp * value_noise(u * fr + 17.f, v * fr + 5.f, t->seed + (uint32_t)o);
    amp *= 0.5f;
    fr *= 2.03f;
  }
  return v3(u, hgt, v);
}

void build_terrain(SceneBuilder& b, uint64_t seed, size_t n_tris, PtrsCamera* cam, int res_w, int res_h) {
  Pcg32 rng(seed, 4);
  int grey = b.add_matte(b.add_constant_texture(3, 0.6f, 0.55f, 0.5f));
  int red = b.add_matte(b.add_constant_texture(3, 0.7f, 0.3f, 0.25f));
  // ~84 % of the budget in the terrain grid, the rest in scattered spheres
  int n = (int)std::floor(std::sqrt(0.84 * (double)n_tris / 2.0));
  n = std::max(n, 4);
  TerrainCtx ctx{(uint32_t)seed};
  {
    MeshInput m = gen_grid(n, n, false, false, terrain_fn, &ctx);
    m.material = grey;
    b.add_mesh(m);
  }
  size_t made = (size_t)2 * n * n;
  const int su = 32, sv = 32;
  while (made + (size_t)2 * su * (sv - 1) <= n_tris) {
    float r = 0.004f + 0.02f * rng.next_f32();
    V3 c = v3(0.05f + 0.9f * rng.next_f32(), 0.f, 0.05f + 0.9f * rng.next_f32());
    c.y = terrain_fn(c.x, c.z, &ctx).y + r * (0.5f + 8.f * rng.next_f32() * rng.next_f32());
    MeshInput m = gen_sphere_uv(su, sv, false);
    m.normal.clear();
    m.obj_to_world = trs(c, v3(r, r, r));
    m.material = red;
    b.add_mesh(m);
    made += m.indices.size() / 3;
  }
  float li[3] = {3.f, 3.f, 3.f};
  M4 lt = M4::identity();
  lt.at(0, 3) = 0.5f;
  lt.at(1, 3) = 1.5f;
  lt.at(2, 3) = 0.5f;
  b.add_point_light(lt, li);
  if (cam) *cam = look_at_camera(v3(0.5f, 0.9f, 1.6f), v3(0.5f, 0.15f, 0.5f), v3(0, 1, 0), 45.f, res_w, res_h);
}

// ------------------------------------------------------------------------------------------------
// C5: atrium
// ------------------------------------------------------------------------------------------------
struct ClothCtx { V3 origin; float w, h, sag, phase; };
static V3 cloth_fn(float u, float v, const void* ctx) {
  const ClothCtx* c = (const ClothCtx*)ctx;
  float sag = c->sag * std::sin(kPi * u) * (0.3f + 0.7f * v);
  float wave = 0.05f * std::sin(u * 18.f + c->phase) * v;
  return v3(c->origin.x + c->w * u, c->origin.y - c->h * v + 0.02f * std::sin(u * 9.f), c->origin.z + sag + wave);
}
struct ColumnCtx { V3 base; float r, h; };
static V3 column_fn(float u, float v, const void* ctx) {
  const ColumnCtx* c = (const ColumnCtx*)ctx;
  float flute = 1.f + 0.04f * std::cos(u * 2.f * kPi * 16.f);
  float taper = 1.f - 0.12f * v + ((v < 0.06f || v > 0.94f) ? 0.25f : 0.f);
  float a = 2.f * kPi * u;
  return v3(c->base.x + c->r * flute * taper * std::cos(a), c->base.y + c->h * v, c->base.z + c->r * flute * taper * std::sin(a));
}
struct ArchCtx { V3 a; float span, rise, depth; bool along_x; };
static V3 arch_fn(float u, float v, const void* ctx) {
  const ArchCtx* c = (const ArchCtx*)ctx;
  float ang = kPi * u;
  float x = 0.5f * c->span * (1.f - std::cos(ang)), y = c->rise * std::sin(ang), d = c->depth * (v - 0.5f);
  return c->along_x ? v3(c->a.x + x, c->a.y + y, c->a.z + d) : v3(c->a.x + d, c->a.y + y, c->a.z + x);
}

void build_atrium(SceneBuilder& b, uint64_t seed, size_t n_tris, PtrsCamera* cam, int res_w, int res_h) {
  Pcg32 rng(seed, 5);
  const float L = 24.f, W = 10.f, H = 9.f;  // hall along x
  float c0[3] = {0.75f, 0.72f, 0.65f}, c1[3] = {0.35f, 0.3f, 0.28f};
  int floor_mat = b.add_matte(b.add_checker_texture(3, c0, c1, 24.f, 10.f, 0.f, 0.f));
  // procedural image texture (exercises the MIPMap path): 256 x 256 brick-like pattern
  std::vector<float> tex(256 * 256 * 3);
  for (int y = 0; y < 256; ++y)
    for (int x = 0; x < 256; ++x) {
      int row = y / 32, bx = (x + (row % 2) * 32) % 64;
      bool mortar = (y % 32) < 3 || bx < 3;
      float n = 0.15f * value_noise((float)x * 0.11f, (float)y * 0.11f, (uint32_t)seed);
      float* p = &tex[((size_t)y * 256 + x) * 3];
      p[0] = mortar ? 0.6f : 0.55f + n;
      p[1] = mortar ? 0.58f : 0.28f + n;
      p[2] = mortar ? 0.55f : 0.2f + n;
    }
  int wall_mat = b.add_matte(b.add_image_texture(3, tex.data(), 256, 256, PTRS_WRAP_REPEAT, 6.f, 3.f, 0.f, 0.f));
  int stone = b.add_matte(b.add_constant_texture(3, 0.7f, 0.68f, 0.62f));
  auto quad = [&](V3 o, V3 ex, V3 ey, int mat, int sub) {
    struct Q { V3 o, ex, ey; } q{o, ex, ey};
    MeshInput m = gen_grid(sub, sub, true, false,
                           [](float u, float v, const void* c) { const Q* q = (const Q*)c; return q->o + q->ex * u + q->ey * v; }, &q);
    m.material = mat;
    b.add_mesh(m);
    return m.indices.size() / 3;
  };
  size_t made = 0;
  made += quad(v3(-L / 2, 0, -W / 2), v3(0, 0, W), v3(L, 0, 0), floor_mat, 8);        // floor (normal +y)
  made += quad(v3(-L / 2, 0, -W / 2), v3(L, 0, 0), v3(0, H, 0), wall_mat, 8);        // back wall
  made += quad(v3(-L / 2, 0, W / 2), v3(0, H, 0), v3(L, 0, 0), wall_mat, 8);         // front wall
  made += quad(v3(-L / 2, 0, -W / 2), v3(0, H, 0), v3(0, 0, W), wall_mat, 8);        // left end
  made += quad(v3(L / 2, 0, -W / 2), v3(0, 0, W), v3(0, H, 0), wall_mat, 8);         // right end
  // open roof: only a frame of beams -> the sky lights the hall
  // budget split: columns 45 %, arches 15 %, cloth 35 %, objects 5 %
  const int n_cols = 20;
  int cu = 64, cv = std::max(4, (int)(0.45 * (double)n_tris / (n_cols * 2.0 * cu)));
  for (int i = 0; i < n_cols; ++i) {
    ColumnCtx c{v3(-L / 2 + 1.2f + (L - 2.4f) * (float)(i / 2) / (float)(n_cols / 2 - 1), 0.f, (i % 2 ? 1.f : -1.f) * (W / 2 - 1.6f)), 0.35f, 6.f};
    MeshInput m = gen_grid(cu, cv, true, true, column_fn, &c);
    m.material = stone;
    b.add_mesh(m);
    made += m.indices.size() / 3;
  }
  const int n_arch = 18;
  int au = std::max(8, (int)std::sqrt(0.15 * (double)n_tris / (n_arch * 2.0) * 4.0)), av = std::max(2, au / 4);
  for (int i = 0; i < n_arch; ++i) {
    float x0 = -L / 2 + 1.2f + (L - 2.4f) * (float)(i / 2) / (float)(n_cols / 2 - 1);
    ArchCtx a{v3(x0, 6.f, (i % 2 ? 1.f : -1.f) * (W / 2 - 1.6f)), (L - 2.4f) / (float)(n_cols / 2 - 1), 1.1f, 0.7f, true};
    MeshInput m = gen_grid(au, av, true, true, arch_fn, &a);
    m.material = stone;
    b.add_mesh(m);
    made += m.indices.size() / 3;
  }
  const int n_cloth = 6;
  int cl = std::max(8, (int)std::sqrt(0.35 * (double)n_tris / (n_cloth * 2.0)));
  for (int i = 0; i < n_cloth; ++i) {
    float hue[3] = {0.3f + 0.6f * rng.next_f32(), 0.2f + 0.5f * rng.next_f32(), 0.2f + 0.5f * rng.next_f32()};
    int mat = b.add_matte(b.add_constant_texture(3, hue[0], hue[1], hue[2]));
    ClothCtx c{v3(-L / 2 + 2.f + 3.6f * (float)i, 7.5f, (i % 2 ? 1.f : -1.f) * 1.2f), 2.6f, 4.5f, 0.8f, rng.next_f32() * 6.f};
    MeshInput m = gen_grid(cl, cl, true, true, cloth_fn, &c);
    m.material = mat;
    b.add_mesh(m);
    made += m.indices.size() / 3;
  }
  int white = b.add_constant_texture(3, 1.f, 1.f, 1.f);
  int glass = b.add_glass(white, white, b.add_constant_texture(1, 1.5f));
  int a = b.add_constant_texture(1, 0.05f);
  int gold = b.add_metal(b.add_constant_texture(3, 0.1431f, 0.3749f, 1.4424f), b.add_constant_texture(3, 3.9831f, 2.3857f, 1.6032f), white, a, a, false);
  int k = 0;
  while (made + 1984 <= n_tris && k < 64) {
    MeshInput m = gen_sphere_uv(32, 32, true);
    float r = 0.35f + 0.25f * rng.next_f32();
    m.obj_to_world = trs(v3(-L / 2 + 2.f + (L - 4.f) * rng.next_f32(), r, -1.5f + 3.f * rng.next_f32()), v3(r, r, r));
    m.material = (k++ % 2) ? glass : gold;
    b.add_mesh(m);
    made += m.indices.size() / 3;
  }
  float sun_l[3] = {6.f, 5.6f, 5.f}, sun_w[3] = {0.3f, 1.f, 0.25f};
  b.add_directional_light(M4::identity(), sun_l, sun_w);
  std::vector<float> sky = synth_sky(1024, 512, seed);
  for (float& v : sky) v = std::min(v, 40.f);  // the directional light stands in for the sun disc
  b.add_infinite_light(mitsuba_env_light_to_world(), sky.data(), 1024, 512);
  if (cam) *cam = look_at_camera(v3(-L / 2 + 1.5f, 2.2f, 0.6f), v3(L / 2, 3.0f, -0.4f), v3(0, 1, 0), 60.f, res_w, res_h);
}

// ------------------------------------------------------------------------------------------------
// Ray sets
// ------------------------------------------------------------------------------------------------
static inline V3 quat_rotate(const float q[4], V3 v) {
  V3 qv = v3(q[0], q[1], q[2]);
  V3 t = cross(qv, v) * 2.0f;
  V3 c = cross(qv, t);
  return t * q[3] + c + v;
}

void coherent_rays(const PtrsCamera& cam, int side, PtrsRay* out) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < side; ++y)
    for (int x = 0; x < side; ++x) {
      float fx = ((float)x + 0.5f) * (float)cam.width / (float)side, fy = ((float)y + 0.5f) * (float)cam.height / (float)side;
      const float* r = cam.raster_to_screen;
      float sx = r[0] * fx + r[1] * fy + r[3], sy = r[4] * fx + r[5] * fy + r[7], sz = r[11];
      float k = cam.persp[3] / (sz + cam.persp[2]);
      V3 pc = v3(sx * k / cam.persp[0], sy * k / cam.persp[1], -k);
      V3 d = normalize(quat_rotate(cam.rot, pc));
      PtrsRay& o = out[(size_t)y * side + x];
      std::memcpy(o.o, cam.trans, 12);
      std::memcpy(o.d, &d, 12);
      o.t_max = INFINITY;
    }
}

void incoherent_rays(const float mn[3], const float mx[3], uint64_t seed, size_t n, PtrsRay* out) {
  const size_t chunk = 1 << 16;
  const size_t n_chunks = (n + chunk - 1) / chunk;
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < n_chunks; ++c) {
    Pcg32 rng(seed, 1000 + c);
    for (size_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i) {
      for (int k = 0; k < 3; ++k) out[i].o[k] = mn[k] + (mx[k] - mn[k]) * rng.next_f32();
      float z = 1.f - 2.f * rng.next_f32();
      float r = std::sqrt(std::max(0.f, 1.f - z * z));
      float ph = 2.f * kPi * rng.next_f32();
      out[i].d[0] = r * std::cos(ph);
      out[i].d[1] = r * std::sin(ph);
      out[i].d[2] = z;
      out[i].t_max = INFINITY;
    }
  }
}

}  // namespace ptrs_host
