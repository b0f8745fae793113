// Ray reordering between bounces: the extend queue of a round is sorted by where its rays start (and roughly where
// they point) before it is traced.
//
// The reference walks one path at a time; there is nothing to reorder (integrator.rs:392-503).  In a wavefront the
// rays of a bounce arrive in the order the previous round's shade kernels happened to append them, so after the first
// bounce a warp's 32 rays start all over the scene: their traversals share no nodes (L1 / L2 misses) and sit in
// different phases (idle lanes).  Sorting the queue by a Morton key of the ray origin — direction octant and dominant
// axis in the low bits — puts rays that start next to each other into the same warp; the shadow / MIS rays and the
// next round's queue inherit that order because shade appends in queue order.  Every path's arithmetic is untouched:
// only the ORDER in which independent paths are processed changes, so radiance per path is bit-identical and the film
// differs by float summation order only (as it already does between two runs: atomics).
//
// The queue length lives in device memory (no host round trip between bounces), so the host sorts an upper estimate
// `m` of it: keys beyond the real length carry bit 29 and sort to the end; if the estimate was short, entries
// [m, n) stay where they were.  Keys are at most 29 bits.  The sort is cub::DeviceRadixSort (library code, like the
// sort of the device BVH build).
#include <cub/device/device_radix_sort.cuh>

#include "launch.hpp"
#include "wavefront.cuh"

namespace ptrs {

namespace {

__device__ __forceinline__ uint32_t spread8(uint32_t x) {  // 8 bits -> every third bit
  x &= 0xffu;
  x = (x | (x << 8)) & 0x00f00fu;
  x = (x | (x << 4)) & 0x0c30c3u;
  x = (x | (x << 2)) & 0x249249u;
  return x;
}

__global__ void __launch_bounds__(256) queue_keys_kernel(const PathSlot* __restrict__ slot, const int* __restrict__ q, const uint32_t* __restrict__ n_ptr,
                                                         uint32_t m, float3 lo, float3 scale, uint32_t* __restrict__ keys, int* __restrict__ q_sorted) {
  const uint32_t n = *n_ptr;
  const uint32_t upto = n > m ? n : m;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < upto; i += gridDim.x * blockDim.x) {
    if (i >= m) {  // beyond the part that is sorted: stays in place
      q_sorted[i] = q[i];
      continue;
    }
    uint32_t key = 1u << 29;  // padding: after every real key
    if (i < n) {
      const PathRay r = ld256(&slot[q[i]].r);
      const uint32_t cx = (uint32_t)fminf(fmaxf((r.ox - lo.x) * scale.x, 0.f), 255.f);
      const uint32_t cy = (uint32_t)fminf(fmaxf((r.oy - lo.y) * scale.y, 0.f), 255.f);
      const uint32_t cz = (uint32_t)fminf(fmaxf((r.oz - lo.z) * scale.z, 0.f), 255.f);
      const uint32_t oct = (r.dx < 0.f ? 1u : 0u) | (r.dy < 0.f ? 2u : 0u) | (r.dz < 0.f ? 4u : 0u);
      const float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
      const uint32_t dom = ax > ay ? (ax > az ? 0u : 2u) : (ay > az ? 1u : 2u);
      key = (((spread8(cx) << 2) | (spread8(cy) << 1) | spread8(cz)) << 5) | (oct << 2) | dom;
    }
    keys[i] = key;
  }
}

}  // namespace

size_t queue_sort_temp_bytes(uint32_t cap) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int*)nullptr, (int*)nullptr, (int)cap, 0, 30, (cudaStream_t)0);
  return bytes;
}

// q[0 .. *n_ptr) -> q_sorted, the first min(m, *n_ptr) entries ordered by ray key.  keys_a / keys_b: m words each.
int sort_queue(cudaStream_t st, int sm, const PathSlot* slot, const int* q, const uint32_t* n_ptr, uint32_t m, const float world_bound[6], uint32_t* keys_a,
               uint32_t* keys_b, int* q_sorted, void* temp, size_t temp_bytes, int begin_bit) {
  if (m == 0) return (int)cudaSuccess;
  float3 lo = make_float3(world_bound[0], world_bound[1], world_bound[2]);
  float3 scale;
  const float ex = world_bound[3] - world_bound[0], ey = world_bound[4] - world_bound[1], ez = world_bound[5] - world_bound[2];
  scale.x = ex > 0.f ? 256.f / ex : 0.f;
  scale.y = ey > 0.f ? 256.f / ey : 0.f;
  scale.z = ez > 0.f ? 256.f / ez : 0.f;
  queue_keys_kernel<<<sm * 8, 256, 0, st>>>(slot, q, n_ptr, m, lo, scale, keys_a, q_sorted);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint32_t*)keys_a, keys_b, q, q_sorted, (int)m, begin_bit, 30, st);
  return (int)e;
}

}  // namespace ptrs
