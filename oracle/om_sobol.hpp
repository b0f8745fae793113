// ORACLE — TEST INFRASTRUCTURE ONLY.  Global Sobol sampler:
// src/pathtracer/sampler/sobol.rs:35-193, src/pathtracer/lowdiscrepancy.rs:9-57.
// Tables come from pathtracer_rs_b200/data/sobol_tables.bin (tools/gen_sobol_tables.py derives them
// from the Joe-Kuo direction numbers and checks them against src/pathtracer/sobolmatrices.rs).
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "om_math.hpp"

namespace oracle {

struct SobolTables {
  uint32_t n_dims = 0, n_cols = 0;
  std::vector<uint32_t> matrices;              // SOBOL_MATRICES_32
  std::vector<std::vector<uint64_t>> vdc;      // VD_C_SOBOL_MATRICES[m - 1]
  std::vector<std::vector<uint64_t>> vdc_inv;  // VD_C_SOBOL_MATRICES_INV[m - 1]

  static SobolTables load(const std::string& path) {
    SobolTables t;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    char magic[4];
    uint32_t hdr[4];
    if (std::fread(magic, 1, 4, f) != 4 || std::memcmp(magic, "SOBL", 4) || std::fread(hdr, 4, 4, f) != 4)
      throw std::runtime_error("bad sobol table header");
    t.n_dims = hdr[0];
    t.n_cols = hdr[1];
    t.matrices.resize((size_t)hdr[0] * hdr[1]);
    if (std::fread(t.matrices.data(), 4, t.matrices.size(), f) != t.matrices.size()) throw std::runtime_error("short read");
    auto rd = [&](std::vector<std::vector<uint64_t>>& dst, uint32_t n) {
      for (uint32_t i = 0; i < n; ++i) {
        uint32_t len;
        if (std::fread(&len, 4, 1, f) != 1) throw std::runtime_error("short read");
        std::vector<uint64_t> v(len);
        if (std::fread(v.data(), 8, len, f) != len) throw std::runtime_error("short read");
        dst.push_back(std::move(v));
      }
    };
    rd(t.vdc, hdr[2]);
    rd(t.vdc_inv, hdr[3]);
    std::fclose(f);
    return t;
  }
};

constexpr float INV_1_2_32 = 0x1.p-32f;  // lowdiscrepancy.rs:7

// lowdiscrepancy.rs:9-39
inline uint64_t sobol_interval_to_index(const SobolTables& T, uint32_t m, uint64_t frame, int32_t px, int32_t py) {
  if (m == 0) return 0;
  const uint32_t m2 = m << 1;
  uint64_t index = frame << m2;
  uint64_t delta = 0;
  for (int c = 0; frame != 0; frame >>= 1, ++c)
    if (frame & 1) delta ^= T.vdc[m - 1][c];
  uint64_t b = ((uint64_t)(((uint32_t)px) << m) | (uint64_t)(int64_t)py) ^ delta;
  for (int c = 0; b != 0; b >>= 1, ++c)
    if (b & 1) index ^= T.vdc_inv[m - 1][c];
  return index;
}

// lowdiscrepancy.rs:42-57
inline float sobol_sample(const SobolTables& T, int64_t a, size_t dimension, uint64_t scramble) {
  uint32_t v = (uint32_t)scramble;
  size_t i = dimension * T.n_cols;
  while (a != 0) {
    if (a & 1) v ^= T.matrices[i];
    a >>= 1;  // arithmetic shift of an i64, like Rust
    i += 1;
  }
  return rmin(ONE_MINUS_EPSILON, (float)v * INV_1_2_32);
}

constexpr size_t ARRAY_START_DIM = 5;  // sobol.rs:11

// SobolSampler with no sample arrays requested (the integrator never requests any).
struct SobolSampler {
  const SobolTables* T = nullptr;
  // SobolSamplerBuilder::new, sobol.rs:35-62
  size_t samples_per_pixel = 1;
  int32_t bounds_min[2] = {0, 0}, bounds_max[2] = {0, 0};
  int32_t resolution = 1;
  uint32_t log_2_resolution = 0;
  // per-pixel / per-sample state
  int32_t current_pixel[2] = {0, 0};
  size_t current_pixel_sample_index = 0;
  size_t dimension = 0;
  int64_t interval_sample_index = 0;
  size_t array_end_dim = 0;
  uint64_t current_scramble_index = 0;

  void configure(const SobolTables* tables, size_t spp, const int32_t sample_bounds[4]) {
    T = tables;
    samples_per_pixel = (size_t)round_up_pow_2_i64((int64_t)spp);
    bounds_min[0] = sample_bounds[0];
    bounds_min[1] = sample_bounds[1];
    bounds_max[0] = sample_bounds[2];
    bounds_max[1] = sample_bounds[3];
    int32_t dx = bounds_max[0] - bounds_min[0], dy = bounds_max[1] - bounds_min[1];
    resolution = round_up_pow_2_i32(dx > dy ? dx : dy);
    log_2_resolution = log2_int((uint64_t)resolution);
  }
  int64_t get_index_for_sample(uint64_t sample_num) const {  // sobol.rs:169-175
    return (int64_t)sobol_interval_to_index(*T, log_2_resolution, sample_num, current_pixel[0] - bounds_min[0],
                                            current_pixel[1] - bounds_min[1]);
  }
  void start_pixel(int32_t x, int32_t y) {  // sobol.rs:81-114
    current_pixel[0] = x;
    current_pixel[1] = y;
    current_pixel_sample_index = 0;
    current_scramble_index = cantor_pairing((uint64_t)(int64_t)(x + HALF_MAX_I_32), (uint64_t)(int64_t)(y + HALF_MAX_I_32));
    dimension = 0;
    interval_sample_index = get_index_for_sample(0);
    array_end_dim = ARRAY_START_DIM;
  }
  // extension for sharded rendering: position on sample `s` of the current pixel directly
  void set_sample(size_t s) {
    current_pixel_sample_index = s;
    dimension = 0;
    interval_sample_index = get_index_for_sample((uint64_t)s);
  }
  bool start_next_sample() {  // sobol.rs:122-127 + CoreSampler::start_next_sample
    dimension = 0;
    interval_sample_index = get_index_for_sample((uint64_t)(current_pixel_sample_index + 1));
    current_pixel_sample_index += 1;
    return current_pixel_sample_index < samples_per_pixel;
  }
  float sample_dimension(int64_t index, size_t dim) const {  // sobol.rs:177-193
    if (dim > 1024) throw std::runtime_error("sobol sampler can only sample up to 1024 dimensions");
    float s = sobol_sample(*T, index, dim, current_scramble_index);
    if (dim == 0 || dim == 1) {
      s = s * (float)resolution + (float)bounds_min[dim];
      s = rclamp(s - (float)current_pixel[dim], 0.f, ONE_MINUS_EPSILON);
    }
    return s;
  }
  float get_1d() {  // sobol.rs:129-137
    if (dimension >= ARRAY_START_DIM && dimension < array_end_dim) dimension = array_end_dim;
    float s = sample_dimension(interval_sample_index, dimension);
    dimension += 1;
    return s;
  }
  Vec2 get_2d() {  // sobol.rs:139-151
    if (dimension + 1 >= ARRAY_START_DIM && dimension < array_end_dim) dimension = array_end_dim;
    Vec2 s{sample_dimension(interval_sample_index, dimension), sample_dimension(interval_sample_index, dimension + 1)};
    dimension += 2;
    return s;
  }
  Vec2 get_camera_sample() {  // sobol.rs:116-120
    Vec2 u = get_2d();
    return Vec2{(float)current_pixel[0] + u.x, (float)current_pixel[1] + u.y};
  }
};

}  // namespace oracle
