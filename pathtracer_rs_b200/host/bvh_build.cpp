#include "bvh_build.hpp"

#include <omp.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace ptrs_host {
namespace {

// Bounds3::empty(): +-f32::MAX, not infinity (src/common/bounds.rs:79-87).
inline Bounds3 empty_bounds() {
  return Bounds3{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
}
// na::RealField::min/max on f32 (bounds.rs:51-65) == f32::min/max: NaN-ignoring like fminf/fmaxf.
inline Bounds3 bunion(const Bounds3& a, const Bounds3& b) {
  Bounds3 r;
  for (int k = 0; k < 3; ++k) {
    r.mn[k] = std::fmin(a.mn[k], b.mn[k]);
    r.mx[k] = std::fmax(a.mx[k], b.mx[k]);
  }
  return r;
}
inline Bounds3 bunion_p(const Bounds3& a, const float p[3]) {
  Bounds3 r;
  for (int k = 0; k < 3; ++k) {
    r.mn[k] = std::fmin(a.mn[k], p[k]);
    r.mx[k] = std::fmax(a.mx[k], p[k]);
  }
  return r;
}
// bounds.rs:112-115
inline float surface_area(const Bounds3& b) {
  float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
  return 2.0f * (dx * dy + dx * dz + dy * dz);
}
// bounds.rs:93-95: diagonal().imax() -> index of the FIRST maximal component (nalgebra imax).
inline int maximum_extent(const Bounds3& b) {
  float d[3] = {b.mx[0] - b.mn[0], b.mx[1] - b.mn[1], b.mx[2] - b.mn[2]};
  int best = 0;
  for (int k = 1; k < 3; ++k)
    if (d[k] > d[best]) best = k;
  return best;
}
// bounds.rs:97-110, one component.
inline float offset_dim(const Bounds3& b, const float p[3], int dim) {
  float o = p[dim] - b.mn[dim];
  if (b.mx[dim] > b.mn[dim]) o /= b.mx[dim] - b.mn[dim];
  return o;
}

struct PrimInfo {  // BVHPrimitiveInfo, accelerator.rs:7-21
  uint32_t prim_num;
  float centroid[3];
  Bounds3 bounds;
};

constexpr int N_BUCKETS = 12;

inline int bucket_of(const Bounds3& cb, const PrimInfo& pi, int dim) {
  // `as usize` saturates negatives/NaN to 0 in Rust; offsets are >= 0 here.
  float f = (float)N_BUCKETS * offset_dim(cb, pi.centroid, dim);
  if (!(f > 0.0f)) return 0;                               // negatives and NaN
  if (f >= (float)N_BUCKETS) return N_BUCKETS - 1;           // also +inf: (int)inf would be undefined here
  return (int)f;
}

struct Builder {
  PrimInfo* info;
  int max_prims;
  size_t task_cutoff;  // sub-ranges at least this large spawn OpenMP tasks

  // Builds [start, end) and appends its nodes in pre-order to `out`, with interior offsets
  // relative to out's own index space (== absolute when out is the root vector).
  void build(size_t start, size_t end, std::vector<PtrsBvhNode>& out, int depth, int* max_depth) {
    if (depth > *max_depth) *max_depth = depth;
    const size_t my = out.size();
    out.emplace_back();

    Bounds3 bounds = empty_bounds();
    for (size_t i = start; i < end; ++i) bounds = bunion(bounds, info[i].bounds);
    const size_t n = end - start;

    auto make_leaf = [&]() {
      PtrsBvhNode& nd = out[my];
      std::memcpy(nd.bounds_min, bounds.mn, 12);
      std::memcpy(nd.bounds_max, bounds.mx, 12);
      nd.offset = (uint32_t)start;  // ordered_prims.len() at this point == start (DFS order)
      nd.n_prims = (uint16_t)n;
      nd.axis = 0;
      nd.pad = 0;
    };

    if (n == 1) {
      make_leaf();
      return;
    }
    Bounds3 cb = empty_bounds();
    for (size_t i = start; i < end; ++i) cb = bunion_p(cb, info[i].centroid);
    const int dim = maximum_extent(cb);
    size_t mid;
    if (cb.mx[dim] == cb.mn[dim]) {
      if (n <= 65535) {  // n_prims is 16 bits wide (LinearBVHNode, accelerator.rs:89-95)
        make_leaf();
        return;
      }
      mid = start + n / 2;  // more coincident primitives than a leaf can count: halve the range
    } else if (n <= 2) {
      // select_nth_unstable_by(mid - start) on two elements with distinct keys == ascending order
      mid = (start + end) / 2;
      if (info[start + 1].centroid[dim] < info[start].centroid[dim])
        std::swap(info[start], info[start + 1]);
    } else {
      size_t count[N_BUCKETS] = {0};
      Bounds3 bb[N_BUCKETS];
      for (int b = 0; b < N_BUCKETS; ++b) bb[b] = empty_bounds();
      for (size_t i = start; i < end; ++i) {
        int b = bucket_of(cb, info[i], dim);
        count[b] += 1;
        bb[b] = bunion(bb[b], info[i].bounds);
      }
      float cost[N_BUCKETS - 1];
      const float inv_total_sa_den = surface_area(bounds);
      for (int i = 0; i < N_BUCKETS - 1; ++i) {
        Bounds3 b0 = empty_bounds(), b1 = empty_bounds();
        size_t c0 = 0, c1 = 0;
        for (int j = 0; j <= i; ++j) {
          b0 = bunion(b0, bb[j]);
          c0 += count[j];
        }
        for (int j = i + 1; j < N_BUCKETS; ++j) {
          b1 = bunion(b1, bb[j]);
          c1 += count[j];
        }
        cost[i] = 1.0f + ((float)c0 * surface_area(b0) + (float)c1 * surface_area(b1)) / inv_total_sa_den;
      }
      float min_cost = cost[0];
      int split = 0;
      for (int i = 1; i < N_BUCKETS - 1; ++i)
        if (cost[i] < min_cost) {
          min_cost = cost[i];
          split = i;
        }
      const float leaf_cost = (float)n;
      if (n > (size_t)max_prims || min_cost < leaf_cost) {
        // Iterator::partition_in_place: repeatedly swap the first `false` with the last `true`.
        size_t first = start, last = end;
        auto pred = [&](const PrimInfo& pi) { return bucket_of(cb, pi, dim) <= split; };
        while (true) {
          while (first != last && pred(info[first])) ++first;
          if (first == last) break;
          --last;
          while (first != last && !pred(info[last])) --last;
          if (first == last) break;
          std::swap(info[first], info[last]);
          ++first;
        }
        mid = first;
      } else {
        make_leaf();
        return;
      }
    }

    // With finite bounds both sides are non-empty.  With non-finite ones (where the reference indexes its buckets out of
    // range and panics, or recurses without end) the range is halved instead, so the build always terminates with a
    // valid tree.
    if (mid == start || mid == end) mid = start + n / 2;
    int d0 = depth, d1 = depth;
    if (n >= task_cutoff) {
      std::vector<PtrsBvhNode> left, right;
#pragma omp task shared(left, d0) firstprivate(start, mid, depth)
      build(start, mid, left, depth + 1, &d0);
#pragma omp task shared(right, d1) firstprivate(mid, end, depth)
      build(mid, end, right, depth + 1, &d1);
#pragma omp taskwait
      // stitch: left subtree lands at my + 1, right subtree at my + 1 + left.size()
      const uint32_t lbase = (uint32_t)(my + 1);
      const uint32_t rbase = (uint32_t)(my + 1 + left.size());
      out.reserve(out.size() + left.size() + right.size());
      for (PtrsBvhNode nd : left) {
        if (nd.n_prims == 0) nd.offset += lbase;
        out.push_back(nd);
      }
      for (PtrsBvhNode nd : right) {
        if (nd.n_prims == 0) nd.offset += rbase;
        out.push_back(nd);
      }
      PtrsBvhNode& me = out[my];
      me.offset = rbase;
    } else {
      build(start, mid, out, depth + 1, &d0);
      const uint32_t second = (uint32_t)out.size();
      build(mid, end, out, depth + 1, &d1);
      out[my].offset = second;
    }
    if (d0 > *max_depth) *max_depth = d0;
    if (d1 > *max_depth) *max_depth = d1;
    PtrsBvhNode& me = out[my];
    std::memcpy(me.bounds_min, bounds.mn, 12);  // == union of the children's bounds
    std::memcpy(me.bounds_max, bounds.mx, 12);
    me.n_prims = 0;
    me.axis = (uint8_t)dim;
    me.pad = 0;
  }
};

}  // namespace

BvhBuildResult build_bvh(const std::vector<Bounds3>& bounds, int max_prims_in_node, int n_threads) {
  BvhBuildResult res;
  const size_t n = bounds.size();
  if (n == 0) return res;
  std::vector<PrimInfo> info(n);
  for (size_t i = 0; i < n; ++i) {
    info[i].prim_num = (uint32_t)i;
    info[i].bounds = bounds[i];
    for (int k = 0; k < 3; ++k)  // accelerator.rs:17: p_min + 0.5 * (p_max - p_min)
      info[i].centroid[k] = bounds[i].mn[k] + 0.5f * (bounds[i].mx[k] - bounds[i].mn[k]);
  }
  Builder b{info.data(), max_prims_in_node, n_threads > 1 ? std::max<size_t>(n / 256, 4096) : (size_t)-1};
  res.nodes.reserve(2 * n / std::max(1, max_prims_in_node / 2) + 16);
  int max_depth = 1;
  if (n_threads > 1) {
#pragma omp parallel num_threads(n_threads)
#pragma omp single
    b.build(0, n, res.nodes, 1, &max_depth);
  } else {
    b.build(0, n, res.nodes, 1, &max_depth);
  }
  res.max_depth = max_depth;
  res.prim_order.resize(n);
  for (size_t i = 0; i < n; ++i) res.prim_order[i] = info[i].prim_num;
  return res;
}

}  // namespace ptrs_host
