// Table construction on the device: what the reference's importers build on the host right
// before the path runs —
//   MIPMap::new's box-filtered pyramid                     src/pathtracer/texture.rs:345-405
//   InfiniteAreaLight::new's sampling density              src/pathtracer/light.rs:372-387
//   Distribution1D::new / Distribution2D::new              src/pathtracer/sampling.rs:133-162, 185-209
// — as kernels, so that a scene description may carry level 0 of an image only and no Distribution2D arrays (for
// the 1k environment map of BASELINE configs[1] that is 10.7 MB less pyramid and 16.8 MB less distribution to
// assemble on the host and to copy).  Built with the exact units' flags (no FMA contraction, IEEE division): every
// value is bit-identical to host/scene_builder.cpp's, which is what tests/test_gpu_parity.py checks.  The two places
// where the host path goes through libm — sin(pi v) per row of the density and log2 of the filter width — are
// evaluated by the library's host code and handed to the kernels, so they are the same glibc values.
#include <algorithm>

#include "launch.hpp"
#include "wavefront.cuh"

namespace ptrs {

namespace {

// level i from level i - 1: mean of the 2 x 2 block, addressed through the wrap mode (texture.rs:386-402)
__global__ void __launch_bounds__(256) mip_level_kernel(const float* __restrict__ prev, int pw, int ph, int channels, int wrap, float* __restrict__ out, int sres,
                                                        int tres) {
  const uint32_t n = (uint32_t)sres * (uint32_t)tres;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = (int)(i % (uint32_t)sres), t = (int)(i / (uint32_t)sres);
    auto texel = [&](int ss, int tt, int c) -> float {
      if (wrap == PTRS_WRAP_REPEAT) {
        ss = abs_mod(ss, pw);
        tt = abs_mod(tt, ph);
      } else if (wrap == PTRS_WRAP_BLACK) {
        if (ss < 0 || ss >= pw || tt < 0 || tt >= ph) return 0.f;
      } else {
        ss = min(max(ss, 0), pw - 1);
        tt = min(max(tt, 0), ph - 1);
      }
      return prev[((size_t)tt * pw + ss) * channels + c];
    };
    for (int c = 0; c < channels; ++c)
      out[(size_t)i * channels + c] = (((texel(2 * s, 2 * t, c) + texel(2 * s + 1, 2 * t, c)) + texel(2 * s, 2 * t + 1, c)) + texel(2 * s + 1, 2 * t + 1, c)) * 0.25f;
  }
}

// func[v][u] = sin(pi (v + 0.5) / nv) * y(MIPMap::lookup_width((u + 0.5) / nu, (v + 0.5) / nv, filter width))   light.rs:375-387
// mode: 0 = finest level only, 1 = coarsest level only, 2 = lerp(level il, il + 1, delta)   (texture.rs:447-464)
__global__ void __launch_bounds__(256) env_density_kernel(const __grid_constant__ DevScene sc, int mip, int nu, int nv, const float* __restrict__ row_sin, int mode,
                                                          int il, float delta, float* __restrict__ func) {
  const PtrsMipMap& mm = sc.mipmaps[mip];
  const uint32_t n = (uint32_t)nu * (uint32_t)nv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int u = (int)(i % (uint32_t)nu), v = (int)(i / (uint32_t)nu);
    const float up = ((float)u + 0.5f) / (float)nu, vp = ((float)v + 0.5f) / (float)nv;
    float c[3] = {0.f, 0.f, 0.f};
    if (mode == 0) {
      mip_triangle(sc, mm, 0, up, vp, c);
    } else if (mode == 1) {
      mip_triangle(sc, mm, mm.n_levels - 1, up, vp, c);
    } else {
      float a[3], b[3];
      mip_triangle(sc, mm, il, up, vp, a);
      mip_triangle(sc, mm, il + 1, up, vp, b);
      for (int k = 0; k < 3; ++k) c[k] = a[k] * (1.0f - delta) + b[k] * delta;
    }
    const float y = c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f;  // spectrum.rs:112-115
    func[i] = row_sin[v] * y;
  }
}

// Distribution1D::new for `rows` independent rows of n entries: the running sum is sequential in the reference
// (sampling.rs:139-145), and float addition does not associate, so one thread walks one row — 1024 rows of 2048
// entries for the 1k map.  cdf: rows x (n + 1); func_int: rows.
__global__ void __launch_bounds__(128) row_cdf_kernel(const float* __restrict__ func, int n, int rows, float* __restrict__ cdf, float* __restrict__ func_int) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* f = func + (size_t)r * n;
  float* c = cdf + (size_t)r * (n + 1);
  c[0] = 0.f;
  float acc = 0.f;
  for (int i = 1; i < n + 1; ++i) {
    acc = acc + f[i - 1] / (float)n;
    c[i] = acc;
  }
  const float fi = acc;
  func_int[r] = fi;
  if (fi == 0.0f) {
    for (int i = 1; i < n + 1; ++i) c[i] = (float)i / (float)n;
  } else {
    for (int i = 1; i < n + 1; ++i) c[i] = c[i] / fi;
  }
}

}  // namespace

void launch_mip_level(cudaStream_t st, const float* prev, int pw, int ph, int channels, int wrap, float* out, int sres, int tres) {
  const uint32_t n = (uint32_t)sres * (uint32_t)tres;
  const int grid = (int)std::min<uint32_t>((n + 255u) / 256u, 148u * 8u);
  mip_level_kernel<<<grid, 256, 0, st>>>(prev, pw, ph, channels, wrap, out, sres, tres);
}
void launch_env_density(cudaStream_t st, const DevScene& sc, int mip, int nu, int nv, const float* row_sin, int mode, int il, float delta, float* func) {
  const uint32_t n = (uint32_t)nu * (uint32_t)nv;
  const int grid = (int)std::min<uint32_t>((n + 255u) / 256u, 148u * 16u);
  env_density_kernel<<<grid, 256, 0, st>>>(sc, mip, nu, nv, row_sin, mode, il, delta, func);
}
void launch_row_cdf(cudaStream_t st, const float* func, int n, int rows, float* cdf, float* func_int) {
  row_cdf_kernel<<<(rows + 127) / 128, 128, 0, st>>>(func, n, rows, cdf, func_int);
}

}  // namespace ptrs
