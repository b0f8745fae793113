// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Parity unpinned: the reference holds
// no golden vectors for this path and cannot be compiled here (no Rust toolchain).
//
// Vector math with nalgebra 0.32 / nalgebra-glm 0.18 evaluation order (crates not vendored with the
// reference; semantics from their documented behaviour) and the leaf helpers of
// src/common/math.rs.  Built with -O2 -ffp-contract=off: rustc never contracts a*b+c into an FMA.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace oracle {

struct Vec3 {
  float x, y, z;
  float& operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
};
inline Vec3 V(float x, float y, float z) { return Vec3{x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3& operator+=(Vec3& a, Vec3 b) { a = a + b; return a; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // (x + y) + z
inline Vec3 cross(Vec3 a, Vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float norm_squared(Vec3 a) { return dot(a, a); }
inline float norm(Vec3 a) { return std::sqrt(norm_squared(a)); }
inline Vec3 normalize(Vec3 a) { float n = norm(a); return {a.x / n, a.y / n, a.z / n}; }
inline Vec3 vabs(Vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline bool is_zero(Vec3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

struct Vec2 { float x, y; float operator[](int i) const { return i ? y : x; } };

// Rust numeric semantics
inline float rmax(float a, float b) { return std::fmax(a, b); }  // f32::max
inline float rmin(float a, float b) { return std::fmin(a, b); }  // f32::min
inline float rclamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // f32::clamp
inline int32_t f2i(float f) {  // `as i32`: saturating, NaN -> 0
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}
inline uint64_t f2usize(float f) {  // `as usize`
  if (!(f > 0.0f)) return 0;
  if (f >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)f;
}

// src/common/math.rs
constexpr float MACHINE_EPSILON = 1.1920929e-7f * 0.5f;           // math.rs:3
constexpr float INV_2_PI = 0.31830987f * 0.5f;                    // math.rs:4 (FRAC_1_PI * 0.5)
constexpr float ONE_MINUS_EPSILON = 0x1.fffffep-1f;               // math.rs:5
constexpr int32_t HALF_MAX_I_32 = INT32_MAX / 2;                  // math.rs:6
constexpr float PI = 3.14159274f, FRAC_1_PI = 0.318309873f, FRAC_PI_2 = 1.57079637f, FRAC_PI_4 = 0.785398185f;

inline float gamma(uint32_t n) {  // math.rs:8-10
  return ((float)n * MACHINE_EPSILON) / (1.0f - (float)n * MACHINE_EPSILON);
}
inline int max_dimension(Vec3 v) {  // math.rs:12-26 (ties go to the later axis)
  if (v.x > v.y) return v.x > v.z ? 0 : 2;
  return v.y > v.z ? 1 : 2;
}
inline Vec3 permute(Vec3 p, int x, int y, int z) { return {p[x], p[y], p[z]}; }  // math.rs:28-35
inline Vec3 face_forward(Vec3 n, Vec3 v) { return dot(n, v) < 0.0f ? -n : n; }   // math.rs:37-46
inline void coordinate_system(Vec3 v1, Vec3* v2, Vec3* v3) {                     // math.rs:48-61
  if (std::fabs(v1.x) > std::fabs(v1.y))
    *v2 = V(-v1.z, 0.0f, v1.x) / std::sqrt(v1.x * v1.x + v1.z * v1.z);
  else
    *v2 = V(0.0f, v1.z, -v1.y) / std::sqrt(v1.y * v1.y + v1.z * v1.z);
  *v3 = cross(v1, *v2);
}
inline uint32_t float_to_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bits_to_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float next_float_up(float v) {  // math.rs:71-88
  if (std::isinf(v) && v > 0.f) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = float_to_bits(v);
  if (v >= 0.0f) ui += 1; else ui -= 1;
  return bits_to_float(ui);
}
// math.rs:90-105.  QUIRK (reproduced, not fixed): the reference has the increments the other way
// round from pbrt (`if v > 0.0 { ui += 1 } else { ui -= 1 }`), so this moves UP by one ulp for
// either sign and turns +-0 into the NaN 0x7fffffff.
inline float next_float_down(float v) {
  if (std::isinf(v) && v < 0.0f) return v;
  if (v == 0.0f) v = -0.0f;
  uint32_t ui = float_to_bits(v);
  if (v > 0.0f) ui += 1; else ui -= 1;
  return bits_to_float(ui);
}
inline Vec3 offset_ray_origin(Vec3 p, Vec3 p_error, Vec3 n, Vec3 w) {  // math.rs:107-131
  float d = dot(vabs(n), p_error);
  Vec3 offset = d * n;
  if (dot(w, n) < 0.0f) offset = -offset;
  Vec3 po = p + offset;
  for (int i = 0; i < 3; ++i) {
    if (offset[i] > 0.0f) po[i] = next_float_up(po[i]);
    else if (offset[i] < 0.0f) po[i] = next_float_down(po[i]);
  }
  return po;
}
inline float gamma_correct(float v) {  // math.rs:133-139
  if (v <= 0.0031308f) return 12.92f * v;
  return 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f;
}
inline bool solve_linear_system_2x2(const float a[2][2], const float b[2], float* x0, float* x1) {  // math.rs:149-165
  float det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
  if (std::fabs(det) < 1e-10f) return false;
  float r0 = (a[1][1] * b[0] - a[0][1] * b[1]) / det;
  float r1 = (a[0][0] * b[1] - a[1][0] * b[0]) / det;
  if (r0 != r0 || r1 != r1) return false;
  *x0 = r0;
  *x1 = r1;
  return true;
}
inline float power_heuristic(int nf, float f_pdf, int ng, float g_pdf) {  // math.rs:167-171
  float f = (float)nf * f_pdf, g = (float)ng * g_pdf;
  return (f * f) / (f * f + g * g);
}
inline float spherical_theta(Vec3 v) { return std::acos(rclamp(v.z, -1.0f, 1.0f)); }  // math.rs:173-175
inline float spherical_phi(Vec3 v) {                                                  // math.rs:177-184
  float p = std::atan2(v.y, v.x);
  return p < 0.0f ? p + 2.0f * PI : p;
}
template <class P>
inline size_t find_interval(size_t size, P pred) {  // math.rs:186-201
  size_t first = 0, len = size;
  while (len > 0) {
    size_t half = len >> 1, middle = first + half;
    if (pred(middle)) {
      first = middle + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  size_t r = first - 1;  // wraps like release-mode usize when first == 0
  size_t hi = size - 2;
  return r > hi ? hi : r;  // clamp(0, size - 2)
}
inline int64_t round_up_pow_2_i64(int64_t v) {  // math.rs:217-230
  v -= 1;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v |= v >> 32;
  return v + 1;
}
inline int32_t round_up_pow_2_i32(int32_t v) {  // math.rs:203-215
  v -= 1;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}
inline int abs_mod(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }  // math.rs:237-244
inline uint32_t log2_int(uint64_t i) { return 63u - (uint32_t)__builtin_clzll(i); }       // math.rs:246-248 (i > 0)
inline float lerp(float x, float y, float a) { return x * (1.0f - a) + y * a; }           // math.rs:250-254
inline uint64_t cantor_pairing(uint64_t x, uint64_t y) { return (x + y) * (x + y + 1) / 2 + y; }  // math.rs:256-258

// RGBSpectrum (src/common/spectrum.rs:6-248): 3 x f32 with component-wise arithmetic.
struct Spectrum {
  float r, g, b;
  float operator[](int i) const { return (&r)[i]; }
};
inline Spectrum S(float c) { return {c, c, c}; }
inline Spectrum S(float r, float g, float b) { return {r, g, b}; }
inline Spectrum operator+(Spectrum a, Spectrum b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
inline Spectrum operator-(Spectrum a, Spectrum b) { return {a.r - b.r, a.g - b.g, a.b - b.b}; }
inline Spectrum operator*(Spectrum a, Spectrum b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
inline Spectrum operator/(Spectrum a, Spectrum b) { return {a.r / b.r, a.g / b.g, a.b / b.b}; }
inline Spectrum operator*(Spectrum a, float s) { return {a.r * s, a.g * s, a.b * s}; }
inline Spectrum operator*(float s, Spectrum a) { return {a.r * s, a.g * s, a.b * s}; }
inline Spectrum operator/(Spectrum a, float s) { return {a.r / s, a.g / s, a.b / s}; }
inline Spectrum& operator+=(Spectrum& a, Spectrum b) { a = a + b; return a; }
inline Spectrum& operator*=(Spectrum& a, Spectrum b) { a = a * b; return a; }
inline Spectrum& operator*=(Spectrum& a, float s) { a = a * s; return a; }
inline Spectrum& operator/=(Spectrum& a, float s) { a = a / s; return a; }
inline bool is_black(Spectrum s) { return s.r == 0.f && s.g == 0.f && s.b == 0.f; }  // spectrum.rs:104-106
inline float lum_y(Spectrum s) { return s.r * 0.212671f + s.g * 0.715160f + s.b * 0.072169f; }  // :112-115
inline float max_component(Spectrum s) { return rmax(rmax(s.r, s.g), s.b); }  // :117-119
inline Spectrum ssqrt(Spectrum s) { return {std::sqrt(s.r), std::sqrt(s.g), std::sqrt(s.b)}; }
inline Spectrum lerp(Spectrum x, Spectrum y, float a) { return x * (1.0f - a) + y * a; }

}  // namespace oracle
