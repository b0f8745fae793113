// ORACLE — TEST INFRASTRUCTURE ONLY.  Texture evaluation: src/pathtracer/texture.rs
// (ConstantTexture :15-29, UVMap :31-54, CheckerTexture :56-89, ImageTexture :91-192,
//  MIPMap::{texel, triangle, lookup, lookup_width} :245-273, :407-464).
#pragma once
#include "om_scene.hpp"

namespace oracle {

struct Tex3 { float v[3]; };

inline void mip_texel(const PtrsSceneDesc* d, const PtrsMipMap& mm, int level, int s, int t, float* out) {
  const int W = mm.width[level], H = mm.height[level], C = mm.channels;
  if (mm.wrap == PTRS_WRAP_REPEAT) {
    s = abs_mod(s, W);
    t = abs_mod(t, H);
  } else if (mm.wrap == PTRS_WRAP_BLACK) {
    if (s < 0 || s >= W || t < 0 || t >= H) {
      for (int c = 0; c < C; ++c) out[c] = 0.f;
      return;
    }
  } else {
    s = s < 0 ? 0 : (s > W - 1 ? W - 1 : s);
    t = t < 0 ? 0 : (t > H - 1 ? H - 1 : t);
  }
  const float* p = d->texels + mm.level_offset[level] + ((size_t)t * W + s) * C;
  for (int c = 0; c < C; ++c) out[c] = p[c];
}

inline void mip_triangle(const PtrsSceneDesc* d, const PtrsMipMap& mm, int level, Vec2 st, float* out) {
  level = level < 0 ? 0 : (level > mm.n_levels - 1 ? mm.n_levels - 1 : level);
  float s = st.x * (float)mm.width[level] - 0.5f;
  float t = st.y * (float)mm.height[level] - 0.5f;
  float s0f = std::floor(s), t0f = std::floor(t);
  float ds = s - s0f, dt = t - t0f;
  int s0 = f2i(s0f), t0 = f2i(t0f);
  float a[3], b[3], c[3], e[3];
  mip_texel(d, mm, level, s0, t0, a);
  mip_texel(d, mm, level, s0, t0 + 1, b);
  mip_texel(d, mm, level, s0 + 1, t0, c);
  mip_texel(d, mm, level, s0 + 1, t0 + 1, e);
  for (int k = 0; k < mm.channels; ++k)
    out[k] = ((a[k] * (1.0f - ds) * (1.0f - dt) + b[k] * (1.0f - ds) * dt) + c[k] * ds * (1.0f - dt)) + e[k] * ds * dt;
}

inline void mip_lookup_width(const PtrsSceneDesc* d, const PtrsMipMap& mm, Vec2 st, float width, float* out) {
  const int n = mm.n_levels;
  float level = (float)n - 1.0f + std::log2(rmax(width, 1e-8f));
  if (level < 0.0f) {
    mip_triangle(d, mm, 0, st, out);
  } else if (level >= (float)(n - 1)) {
    mip_triangle(d, mm, n - 1, st, out);
  } else {
    float il = std::floor(level);
    float delta = level - il;
    float a[3], b[3];
    mip_triangle(d, mm, (int)f2usize(il), st, a);
    mip_triangle(d, mm, (int)f2usize(il) + 1, st, b);
    for (int k = 0; k < mm.channels; ++k) out[k] = a[k] * (1.0f - delta) + b[k] * delta;
  }
}

// Texture<T>::evaluate for any texture id; result in out[0..channels)
inline void tex_eval(const Scene& sc, int tex_id, const SurfaceInteraction& it, float* out) {
  const PtrsTexture& t = sc.d->textures[tex_id];
  switch (t.type) {
    case PTRS_TEX_CONSTANT:
      out[0] = t.v1[0]; out[1] = t.v1[1]; out[2] = t.v1[2];
      return;
    case PTRS_TEX_CHECKER: {
      float s = t.su * it.uv.x + t.du, tt = t.sv * it.uv.y + t.dv;  // UVMap::map
      float s_idx = s - std::floor(s), t_idx = tt - std::floor(tt);
      const float* v = ((s_idx <= 0.5f && t_idx <= 0.5f) || (s_idx >= 0.5f && t_idx >= 0.5f)) ? t.v2 : t.v1;
      out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
      return;
    }
    default: {
      float dsdx = t.su * it.dudx, dtdx = t.sv * it.dvdx, dsdy = t.su * it.dudy, dtdy = t.sv * it.dvdy;
      Vec2 st{t.su * it.uv.x + t.du, t.sv * it.uv.y + t.dv};
      float width = rmax(rmax(std::fabs(dsdx), std::fabs(dtdx)), rmax(std::fabs(dsdy), std::fabs(dtdy)));  // MIPMap::lookup
      out[1] = out[2] = 0.f;
      mip_lookup_width(sc.d, sc.d->mipmaps[t.mip], st, width, out);
      return;
    }
  }
}
inline float tex_f32(const Scene& sc, int id, const SurfaceInteraction& it) {
  float o[3];
  tex_eval(sc, id, it, o);
  return o[0];
}
inline Spectrum tex_spectrum(const Scene& sc, int id, const SurfaceInteraction& it) {
  float o[3];
  tex_eval(sc, id, it, o);
  return S(o[0], o[1], o[2]);
}
inline Vec3 tex_vec3(const Scene& sc, int id, const SurfaceInteraction& it) {
  float o[3];
  tex_eval(sc, id, it, o);
  return V(o[0], o[1], o[2]);
}

}  // namespace oracle
