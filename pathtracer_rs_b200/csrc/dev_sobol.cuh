// Global Sobol sampler on the device: src/pathtracer/sampler/sobol.rs:81-193,
// src/pathtracer/lowdiscrepancy.rs:9-57.  Integer XOR networks; bit-exact with the CPU path.
#pragma once
#include "dev_math.cuh"

namespace ptrs {

// VD_C_SOBOL_MATRICES[m-1] / VD_C_SOBOL_MATRICES_INV[m-1] for the render's log2 resolution m
// (at most 52 u64 each), uploaded per render into constant memory.
struct SobolConfig {
  uint64_t vdc[52];
  uint64_t vdc_inv[52];
  int32_t bounds_min[2];
  int32_t resolution;
  uint32_t log2_resolution;
  uint32_t n_vdc, n_vdc_inv;
  int32_t spp;  // rounded up to a power of two
  int32_t pad;
};

#define PT_SOBOL_COLS 52
#define PT_ARRAY_START_DIM 5

// lowdiscrepancy.rs:9-39
PT_DEV uint64_t sobol_interval_to_index(const SobolConfig& c, uint64_t frame, int32_t px, int32_t py) {
  const uint32_t m = c.log2_resolution;
  if (m == 0) return 0;
  uint64_t index = frame << (m << 1);
  uint64_t delta = 0;
  for (int k = 0; frame != 0; frame >>= 1, ++k)
    if (frame & 1) delta ^= c.vdc[k];
  uint64_t b = ((uint64_t)(((uint32_t)px) << m) | (uint64_t)(int64_t)py) ^ delta;
  for (int k = 0; b != 0; b >>= 1, ++k)
    if (b & 1) index ^= c.vdc_inv[k];
  return index;
}

// lowdiscrepancy.rs:42-57
PT_DEV float sobol_sample(const uint32_t* __restrict__ matrices, uint64_t index, uint32_t dimension, uint32_t scramble) {
  uint32_t v = scramble;
  const uint32_t* col = matrices + dimension * PT_SOBOL_COLS;
  uint32_t lo = (uint32_t)index, hi = (uint32_t)(index >> 32);
  while (lo) {
    int k = __ffs(lo) - 1;
    v ^= __ldg(col + k);
    lo &= lo - 1;
  }
  while (hi) {
    int k = __ffs(hi) - 1;
    v ^= __ldg(col + 32 + k);
    hi &= hi - 1;
  }
  return fminf(PT_ONE_MINUS_EPSILON, (float)v * 0x1.p-32f);
}

// Per-path sampler state: the reference's SobolSampler minus everything that is constant per render.
// A bounce draws up to 8 consecutive dimensions (NEE 2+2+1, BSDF 2, roulette 1; 9 with the 4 -> 5 skip
// of the first bounce), so shade fills a WINDOW of 9 dimensions in ONE pass over the set bits of the
// index: each bit costs 9 independent loads from a bit-major copy of the table (36 contiguous bytes)
// instead of 9 separate dependent-latency bit loops.  Values are identical to sobol_sample().
#define PT_SOBOL_WINDOW 9
struct PathSampler {
  uint64_t index;     // interval_sample_index
  uint32_t scramble;  // current_scramble_index as u32
  uint32_t dimension;
  int32_t px, py;
  uint32_t win_base;              // first dimension held in win[], 0xffffffff = no window
  uint32_t win[PT_SOBOL_WINDOW];  // scramble ^ (xor of matrix columns), i.e. the u32 before the 2^-32 scale
};

// mt = SOBOL_MATRICES_32 transposed to [52 bits][1024 dimensions].
// The bit loop is WARP-UNIFORM: it runs over the union of the set bits of the converged lanes' indices and
// each lane masks the column in or out.  Lanes of a shade launch sit at (nearly) the same dimension, so every
// load is a broadcast of one 36-byte row segment — one L1 wavefront instead of one per distinct bit.
PT_DEV void sobol_window_fill(const uint32_t* __restrict__ mt, PathSampler& s, uint32_t base) {
  if (base > 1024u - PT_SOBOL_WINDOW) base = 1024u - PT_SOBOL_WINDOW;
  s.win_base = base;
#pragma unroll
  for (int j = 0; j < PT_SOBOL_WINDOW; ++j) s.win[j] = s.scramble;
  const uint32_t lo = (uint32_t)s.index, hi = (uint32_t)(s.index >> 32);
  const uint32_t grp = __activemask();
  uint32_t all_lo = __reduce_or_sync(grp, lo), all_hi = __reduce_or_sync(grp, hi);
  while (all_lo) {
    const uint32_t b = (uint32_t)(__ffs(all_lo) - 1);
    const uint32_t* row = mt + b * 1024u + base;
    const uint32_t m = 0u - ((lo >> b) & 1u);
#pragma unroll
    for (int j = 0; j < PT_SOBOL_WINDOW; ++j) s.win[j] ^= __ldg(row + j) & m;
    all_lo &= all_lo - 1;
  }
  while (all_hi) {
    const uint32_t b = (uint32_t)(__ffs(all_hi) - 1);
    const uint32_t* row = mt + (32u + b) * 1024u + base;
    const uint32_t m = 0u - ((hi >> b) & 1u);
#pragma unroll
    for (int j = 0; j < PT_SOBOL_WINDOW; ++j) s.win[j] ^= __ldg(row + j) & m;
    all_hi &= all_hi - 1;
  }
}

PT_DEV uint32_t pixel_scramble(int32_t x, int32_t y) {  // sobol.rs:83-86 (+ `scramble as u32`)
  return (uint32_t)cantor_pairing((uint64_t)(int64_t)(x + PT_HALF_MAX_I32), (uint64_t)(int64_t)(y + PT_HALF_MAX_I32));
}

PT_DEV float sample_dimension(const SobolConfig& c, const uint32_t* __restrict__ matrices, const PathSampler& s, uint32_t dim) {
  float v;
  const uint32_t rel = dim - s.win_base;
  if (s.win_base != 0xffffffffu && rel < PT_SOBOL_WINDOW) {
    uint32_t raw = s.win[0];
#pragma unroll
    for (int j = 1; j < PT_SOBOL_WINDOW; ++j) raw = rel == j ? s.win[j] : raw;  // register select, no local memory
    v = fminf(PT_ONE_MINUS_EPSILON, (float)raw * 0x1.p-32f);
  } else {
    v = sobol_sample(matrices, s.index, dim, s.scramble);  // sobol.rs:177-193
  }
  if (dim == 0 || dim == 1) {
    int32_t pmin = dim == 0 ? c.bounds_min[0] : c.bounds_min[1];
    int32_t pix = dim == 0 ? s.px : s.py;
    v = v * (float)c.resolution + (float)pmin;
    v = rclamp(v - (float)pix, 0.f, PT_ONE_MINUS_EPSILON);
  }
  return v;
}
PT_DEV float get_1d(const SobolConfig& c, const uint32_t* __restrict__ m, PathSampler& s) {  // sobol.rs:129-137 (array range empty)
  float v = sample_dimension(c, m, s, s.dimension);
  s.dimension += 1;
  return v;
}
PT_DEV V2 get_2d(const SobolConfig& c, const uint32_t* __restrict__ m, PathSampler& s) {  // sobol.rs:139-151
  if (s.dimension + 1 >= PT_ARRAY_START_DIM && s.dimension < PT_ARRAY_START_DIM) s.dimension = PT_ARRAY_START_DIM;
  V2 v{sample_dimension(c, m, s, s.dimension), sample_dimension(c, m, s, s.dimension + 1)};
  s.dimension += 2;
  return v;
}

}  // namespace ptrs
