// extend / connect / standalone intersect kernels (BVH traversal lives only in this unit)
#include "launch.hpp"
#include "wavefront.cuh"

namespace ptrs {

// ---- extend ----------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128) extend_kernel(DevScene sc, PathArrays P, const int* __restrict__ q_ext, int* __restrict__ q_class,
                                                      uint32_t cap, RoundCounters* ctr, GlobalCounters* g) {
  const uint32_t n = ctr->n_ext;
  const int lane = threadIdx.x & 31;
  uint32_t c_nodes = 0, c_tris = 0;
  for (;;) {
    const uint32_t base = warp_fetch32(&ctr->t_ext);
    if (base >= n) break;
    const uint32_t i = base + lane;
    int cls = -1;
    int p = 0;
    if (i < n) {
      p = q_ext[i];
      const float4 o4 = P.ray_o[p], d4 = P.ray_d[p];
      DevHit hit;
      traverse<false, COUNT>(sc, mk3(o4), mk3(d4), CUDART_INF_F, &hit, &c_nodes, &c_tris);
      P.hit_prim[p] = hit.prim;
      P.hit_tb[p] = make_float4(hit.t, hit.b0, hit.b1, hit.b2);
      if (hit.prim < 0) cls = PT_CLASS_MISS;
      else cls = sc.materials[__float_as_int(__ldg(&sc.tri_verts[3 * (size_t)hit.prim].w))].type;
    }
#pragma unroll
    for (int c = 0; c < PT_N_CLASSES; ++c) warp_push(cls == c, (uint32_t)p, q_class + (size_t)c * cap, &ctr->n_class[c]);
  }
  if (COUNT) {
    warp_sum_add(c_nodes, &g->nodes_tested);
    warp_sum_add(c_tris, &g->tris_tested);
  }
}

// ---- connect ---------------------------------------------------------------------------------------
// second half of estimate_direct: shadow test (integrator.rs:66-78), MIS ray (integrator.rs:113-135)
template <bool COUNT>
__global__ void __launch_bounds__(128) connect_kernel(DevScene sc, PathArrays P, const int* __restrict__ q_nee, RoundCounters* ctr,
                                                       GlobalCounters* g) {
  const uint32_t n = ctr->n_nee;
  const int lane = threadIdx.x & 31;
  uint32_t c_nodes = 0, c_tris = 0;
  for (;;) {
    const uint32_t base = warp_fetch32(&ctr->t_nee);
    if (base >= n) break;
    const uint32_t i = base + lane;
    bool did_shadow = false, did_mis = false;
    if (i < n) {
      const int p = q_nee[i];
      const float4 n0 = P.nee0[p], n1 = P.nee1[p], n2 = P.nee2[p], n3 = P.nee3[p];
      const uint32_t nf = __float_as_uint(n3.w);
      const int light_idx = (int)(nf & 0x3fffffffu);
      Spec ld = sp(0.0f);
      if (nf & PT_NEE_SHADOW) {
        did_shadow = true;
        DevHit h;
        const bool occluded = traverse<true, COUNT>(sc, mk3(n0), mk3(n1), 1.0f - PT_SHADOW_EPSILON, &h, &c_nodes, &c_tris);
        if (!occluded) ld = ld + sp(n0.w, n1.w, n2.w);
      }
      if (nf & PT_NEE_MIS) {
        did_mis = true;
        const float4 n4 = P.nee4[p];
        const float scat_pdf = P.nee5[p].w;
        const V3 md = mk3(n3);
        DevHit h;
        const bool found = traverse<false, COUNT>(sc, mk3(n2), md, CUDART_INF_F, &h, &c_nodes, &c_tris);
        Spec li = sp(0.0f);
        if (found) {
          const int hl = __float_as_int(__ldg(&sc.tri_verts[3 * (size_t)h.prim + 1].w));
          if (hl == light_idx) {
            SurfInter si;
            reconstruct_hit(sc, h.prim, h.b0, h.b1, h.b2, md, &si);
            li = area_le(sc, hl, si, -md);
          }
        } else if (sc.lights[light_idx].type == PTRS_LIGHT_INFINITE) {
          li = env_le(sc, sc.lights[light_idx], md);
        }
        if (!is_black(li)) ld = ld + sp(n4.x, n4.y, n4.z) * li * sp(1.0f) * n4.w / scat_pdf;
      }
      const float4 b5 = P.nee5[p];
      const float4 l4 = P.L[p];
      Spec L = sp(l4.x, l4.y, l4.z) + sp(b5.x, b5.y, b5.z) * ((float)sc.n_lights * ld);
      P.L[p] = make_float4(L.r, L.g, L.b, 0.f);
    }
    warp_count(did_shadow, &g->shadow_rays);
    warp_count(did_mis, &g->mis_rays);
  }
  if (COUNT) {
    warp_sum_add(c_nodes, &g->nodes_tested);
    warp_sum_add(c_tris, &g->tris_tested);
  }
}

// ---- standalone traversal kernels (ptrs_intersect*, the BVH microbenchmark) --------------------------
template <bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128) intersect_kernel(DevScene sc, const PtrsRay* __restrict__ rays, uint32_t n, PtrsHit* __restrict__ hits,
                                                         uint8_t* __restrict__ occluded, uint32_t* ticket, GlobalCounters* g) {
  const int lane = threadIdx.x & 31;
  uint32_t c_nodes = 0, c_tris = 0;
  for (;;) {
    const uint32_t base = warp_fetch32(ticket);
    if (base >= n) break;
    const uint32_t i = base + lane;
    if (i >= n) continue;
    const float* r = (const float*)(rays + i);
    const V3 o = mk3(__ldg(r), __ldg(r + 1), __ldg(r + 2)), d = mk3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
    const float t_max = __ldg(r + 6);
    DevHit h;
    const bool found = traverse<ANY_HIT, COUNT>(sc, o, d, t_max, &h, &c_nodes, &c_tris);
    if (ANY_HIT) {
      occluded[i] = found ? 1 : 0;
    } else {
      PtrsHit out;
      out.prim = found ? h.prim : -1;
      out.t = found ? h.t : 0.f;
      out.b0 = h.b0;
      out.b1 = h.b1;
      out.b2 = h.b2;
      hits[i] = out;
    }
  }
  if (COUNT) {
    warp_sum_add(c_nodes, &g->nodes_tested);
    warp_sum_add(c_tris, &g->tris_tested);
  }
}


// ---- launchers -----------------------------------------------------------------------------------------
void launch_extend(cudaStream_t st, int sm, bool count, const DevScene& sc, const PathArrays& P, const int* q_ext, int* q_class, uint32_t cap,
                   RoundCounters* ctr, GlobalCounters* g) {
  static int grid[2] = {0, 0};
  if (!grid[0]) {
    grid[0] = persistent_grid(extend_kernel<false>, 128, sm);
    grid[1] = persistent_grid(extend_kernel<true>, 128, sm);
  }
  if (count) extend_kernel<true><<<grid[1], 128, 0, st>>>(sc, P, q_ext, q_class, cap, ctr, g);
  else extend_kernel<false><<<grid[0], 128, 0, st>>>(sc, P, q_ext, q_class, cap, ctr, g);
}
void launch_connect(cudaStream_t st, int sm, bool count, const DevScene& sc, const PathArrays& P, const int* q_nee, RoundCounters* ctr,
                    GlobalCounters* g) {
  static int grid[2] = {0, 0};
  if (!grid[0]) {
    grid[0] = persistent_grid(connect_kernel<false>, 128, sm);
    grid[1] = persistent_grid(connect_kernel<true>, 128, sm);
  }
  if (count) connect_kernel<true><<<grid[1], 128, 0, st>>>(sc, P, q_nee, ctr, g);
  else connect_kernel<false><<<grid[0], 128, 0, st>>>(sc, P, q_nee, ctr, g);
}
void launch_intersect(cudaStream_t st, int sm, bool any_hit, bool count, const DevScene& sc, const PtrsRay* rays, uint32_t n, PtrsHit* hits,
                      uint8_t* occluded, uint32_t* ticket, GlobalCounters* g) {
  static int grid[4] = {0, 0, 0, 0};
  if (!grid[0]) {
    grid[0] = persistent_grid(intersect_kernel<false, false>, 128, sm);
    grid[1] = persistent_grid(intersect_kernel<false, true>, 128, sm);
    grid[2] = persistent_grid(intersect_kernel<true, false>, 128, sm);
    grid[3] = persistent_grid(intersect_kernel<true, true>, 128, sm);
  }
  if (!any_hit && !count) intersect_kernel<false, false><<<grid[0], 128, 0, st>>>(sc, rays, n, hits, occluded, ticket, g);
  else if (!any_hit) intersect_kernel<false, true><<<grid[1], 128, 0, st>>>(sc, rays, n, hits, occluded, ticket, g);
  else if (!count) intersect_kernel<true, false><<<grid[2], 128, 0, st>>>(sc, rays, n, hits, occluded, ticket, g);
  else intersect_kernel<true, true><<<grid[3], 128, 0, st>>>(sc, rays, n, hits, occluded, ticket, g);
}

}  // namespace ptrs
