// Function-level parity probes behind ptrs_bxdf_eval / ptrs_bxdf_sample / ptrs_light_sample / ptrs_light_pdf: the
// device's BxDF and light code evaluated one call at a time, for lobe-by-lobe and light-by-light comparison with the
// CPU path (bxdf/mod.rs, bxdf/fresnel.rs, bxdf/microfacet.rs, material/disney.rs, light.rs).  Built twice
// (-DPT_PROBE_EXACT=0 with the shade kernels' flags, =1 with the exact units' flags), like the shade kernels.
#include "launch.hpp"
#include "wavefront.cuh"

#ifndef PT_PROBE_EXACT
#define PT_PROBE_EXACT 0
#endif

namespace ptrs {

PT_DEV Lobe lobe_from_desc(const PtrsLobeDesc& d) {
  Lobe l;
  l.kind = d.kind;
  l.fresnel = d.fresnel;
  l.r = sp(d.r[0], d.r[1], d.r[2]);
  l.t = sp(d.t[0], d.t[1], d.t[2]);
  l.fa = sp(d.fa[0], d.fa[1], d.fa[2]);
  l.fb = sp(d.fb[0], d.fb[1], d.fb[2]);
  l.eta_a = d.eta_a;
  l.eta_b = d.eta_b;
  l.alpha_x = fmaxf(d.alpha_x, 0.001f);  // TrowbridgeReitzDistribution::new, microfacet.rs:113-116
  l.alpha_y = fmaxf(d.alpha_y, 0.001f);
  l.disney_g = d.disney_g;
  return l;
}

template <bool EXACT>
__global__ void bxdf_eval_probe_kernel(const __grid_constant__ PtrsLobeDesc desc, const float* __restrict__ wo, const float* __restrict__ wi, uint32_t n,
                                       float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Lobe l = lobe_from_desc(desc);
  const V3 o = mk3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), w = mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
  const Spec f = lobe_f(l, o, w);
  out[4 * i] = f.r;
  out[4 * i + 1] = f.g;
  out[4 * i + 2] = f.b;
  out[4 * i + 3] = lobe_pdf(l, o, w);
}

template <bool EXACT>
__global__ void bxdf_sample_probe_kernel(const __grid_constant__ PtrsLobeDesc desc, const float* __restrict__ wo, const float* __restrict__ u, uint32_t n,
                                         float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Lobe l = lobe_from_desc(desc);
  const V3 o = mk3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
  V3 w = mk3(0, 0, 0);
  float pdf = 0.f;
  uint32_t sampled = lobe_type(l);
  const Spec f = lobe_sample_f(l, o, &w, V2{u[2 * i], u[2 * i + 1]}, &pdf, &sampled);
  float* q = out + 8 * (size_t)i;
  q[0] = w.x;
  q[1] = w.y;
  q[2] = w.z;
  q[3] = f.r;
  q[4] = f.g;
  q[5] = f.b;
  q[6] = pdf;
  q[7] = (float)sampled;
}

template <bool EXACT>
__global__ void light_sample_probe_kernel(const __grid_constant__ DevScene sc, int light, const float* __restrict__ ref_p, const float* __restrict__ ref_n,
                                          const float* __restrict__ u, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Inter ref;
  ref.p = mk3(ref_p[3 * i], ref_p[3 * i + 1], ref_p[3 * i + 2]);
  ref.n = mk3(ref_n[3 * i], ref_n[3 * i + 1], ref_n[3 * i + 2]);
  ref.p_error = mk3(0, 0, 0);
  LightSample ls;
  light_sample_li(sc, sc.lights[light], ref, V2{u[2 * i], u[2 * i + 1]}, &ls);
  V3 so = mk3(0, 0, 0), sd = mk3(0, 0, 0);
  spawn_ray_to_it(ref, ls.p1, &so, &sd);
  float* q = out + 16 * (size_t)i;
  q[0] = ls.li.r;
  q[1] = ls.li.g;
  q[2] = ls.li.b;
  q[3] = ls.wi.x;
  q[4] = ls.wi.y;
  q[5] = ls.wi.z;
  q[6] = ls.pdf;
  q[7] = so.x;
  q[8] = so.y;
  q[9] = so.z;
  q[10] = sd.x;
  q[11] = sd.y;
  q[12] = sd.z;
  q[13] = q[14] = q[15] = 0.f;
}

template <bool EXACT>
__global__ void light_pdf_probe_kernel(const __grid_constant__ DevScene sc, int light, const float* __restrict__ ref_p, const float* __restrict__ ref_n,
                                       const float* __restrict__ wi, uint32_t n, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Inter ref;
  ref.p = mk3(ref_p[3 * i], ref_p[3 * i + 1], ref_p[3 * i + 2]);
  ref.n = mk3(ref_n[3 * i], ref_n[3 * i + 1], ref_n[3 * i + 2]);
  ref.p_error = mk3(0, 0, 0);
  out[i] = light_pdf_li(sc, sc.lights[light], ref, mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]));
}

#define PT_CAT2(a, b) a##b
#define PT_CAT(a, b) PT_CAT2(a, b)
#if PT_PROBE_EXACT
#define PT_PROBE_FN(name) PT_CAT(name, _exact)
#else
#define PT_PROBE_FN(name) PT_CAT(name, _fast)
#endif
constexpr bool kExact = PT_PROBE_EXACT != 0;

void PT_PROBE_FN(launch_bxdf_eval_probe)(cudaStream_t st, const PtrsLobeDesc& d, const float* wo, const float* wi, uint32_t n, float* out) {
  bxdf_eval_probe_kernel<kExact><<<(n + 127) / 128, 128, 0, st>>>(d, wo, wi, n, out);
}
void PT_PROBE_FN(launch_bxdf_sample_probe)(cudaStream_t st, const PtrsLobeDesc& d, const float* wo, const float* u, uint32_t n, float* out) {
  bxdf_sample_probe_kernel<kExact><<<(n + 127) / 128, 128, 0, st>>>(d, wo, u, n, out);
}
void PT_PROBE_FN(launch_light_sample_probe)(cudaStream_t st, const DevScene& sc, int light, const float* p, const float* nn, const float* u, uint32_t n,
                                            float* out) {
  light_sample_probe_kernel<kExact><<<(n + 127) / 128, 128, 0, st>>>(sc, light, p, nn, u, n, out);
}
void PT_PROBE_FN(launch_light_pdf_probe)(cudaStream_t st, const DevScene& sc, int light, const float* p, const float* nn, const float* wi, uint32_t n,
                                         float* out) {
  light_pdf_probe_kernel<kExact><<<(n + 127) / 128, 128, 0, st>>>(sc, light, p, nn, wi, n, out);
}

}  // namespace ptrs
