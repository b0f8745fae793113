// Host-side SAH BVH build producing the reference's LinearBVHNode array.
// Restates BVH::new / recursive_build / flatten_bvh_tree (src/pathtracer/accelerator.rs:103-346).
// This is the "reference-built BVH" of the north star: it runs on the host (as the Rust importer
// does, with max_prims_in_node = 4: importer/mitsuba.rs:361, importer/gltf.rs:547) and its output is
// handed to the device library through PtrsSceneDesc.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/ptrs_b200.h"

namespace ptrs_host {

struct Bounds3 {
  float mn[3], mx[3];
};

struct BvhBuildResult {
  std::vector<PtrsBvhNode> nodes;   // DFS pre-order, first child implicit at idx + 1
  std::vector<uint32_t> prim_order; // ordered_prims[i] = input primitive prim_order[i]
  int max_depth = 0;
};

// bounds[i] = world_bound of input primitive i (Triangle::world_bound, shape.rs:526-531).
// n_threads <= 1 builds serially; otherwise the top of the tree is split into OpenMP tasks whose
// sub-arrays are stitched back in DFS order (the result is identical to the serial build).
BvhBuildResult build_bvh(const std::vector<Bounds3>& bounds, int max_prims_in_node, int n_threads);

}  // namespace ptrs_host
