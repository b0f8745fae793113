// Small float vector helpers for the host-side scene preparation code.
// All arithmetic is plain IEEE f32 in source order (the library is built with -ffp-contract=off),
// matching how rustc/LLVM compiles the reference (no FMA contraction).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace ptrs_host {

struct V3 {
  float x, y, z;
  float& operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / norm(a); }
inline V3 vmin(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }

// Row-major 4x4.
struct M4 {
  float m[16];
  float& at(int r, int c) { return m[r * 4 + c]; }
  float at(int r, int c) const { return m[r * 4 + c]; }
  static M4 identity() {
    M4 r{};
    for (int i = 0; i < 4; ++i) r.m[i * 5] = 1.0f;
    return r;
  }
};
inline M4 operator*(const M4& a, const M4& b) {
  M4 r{};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.0f;
      for (int k = 0; k < 4; ++k) s += a.at(i, k) * b.at(k, j);
      r.at(i, j) = s;
    }
  return r;
}
// Projective3 * Point3 (homogeneous divide) and * Vector3 (no translation): nalgebra semantics.
inline V3 xform_point(const M4& t, V3 p) {
  float x = t.at(0, 0) * p.x + t.at(0, 1) * p.y + t.at(0, 2) * p.z + t.at(0, 3);
  float y = t.at(1, 0) * p.x + t.at(1, 1) * p.y + t.at(1, 2) * p.z + t.at(1, 3);
  float z = t.at(2, 0) * p.x + t.at(2, 1) * p.y + t.at(2, 2) * p.z + t.at(2, 3);
  float w = t.at(3, 0) * p.x + t.at(3, 1) * p.y + t.at(3, 2) * p.z + t.at(3, 3);
  if (w != 0.0f && w != 1.0f) return {x / w, y / w, z / w};
  return {x, y, z};
}
inline V3 xform_vector(const M4& t, V3 v) {
  return {t.at(0, 0) * v.x + t.at(0, 1) * v.y + t.at(0, 2) * v.z,
          t.at(1, 0) * v.x + t.at(1, 1) * v.y + t.at(1, 2) * v.z,
          t.at(2, 0) * v.x + t.at(2, 1) * v.y + t.at(2, 2) * v.z};
}
bool invert(const M4& a, M4* out);  // Gauss-Jordan with partial pivoting, in double

// PCG32 (O'Neill), used by every seeded procedural generator.
struct Pcg32 {
  uint64_t state, inc;
  explicit Pcg32(uint64_t seed, uint64_t seq = 1) {
    state = 0u;
    inc = (seq << 1u) | 1u;
    next_u32();
    state += seed;
    next_u32();
  }
  uint32_t next_u32() {
    uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
  }
  float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }  // [0,1)
};

}  // namespace ptrs_host
