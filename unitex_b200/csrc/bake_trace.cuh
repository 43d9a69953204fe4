// Closest-hit traversal of the packed LBVH (48 B nodes: aabb[6], left, right, prim), shared by the standalone
// intersect kernel and the fused UV-bake texel kernel.  Order and quirks follow the reference's bvh_hit
// (TextureTools/texturetools/raytracing/rt_aprmis/bvhworkers/intersect_test2.slang:63-146): push left, push right, pop
// right first; slab test against the running closest t; Moller-Trumbore without a t-range test; the reported triangle is
// the LAST accepted leaf.  Translation units including this are built with -fmad=false.
#pragma once
#include <cuda_runtime.h>

namespace utx {

struct RayHit {
  int any, tid;
  float t, u, v;
};
__device__ __forceinline__ float dot3f(float ax, float ay, float az, float bx, float by, float bz) {
  return (ax * bx + ay * by) + az * bz;
}
// slab test with the per-ray reciprocal direction hoisted out of the traversal loop: the reference recomputes
// `1.0 / ray_d_i` (with 1e-6 substituted for 0) at every node (intersect_test2.slang:18-21); the quotient is the same
// IEEE value every time, so computing it once per ray is bit-identical and removes three divisions per node visit.
__device__ __forceinline__ bool aabb_hit_dev(const float* o, const float* inv, float tmin, float tmax, const float* bb) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t0 = (bb[i] - o[i]) * inv[i], t1 = (bb[3 + i] - o[i]) * inv[i];
    if (inv[i] < 0.0f) { const float t = t1; t1 = t0; t0 = t; }
    tmin = t0 > tmin ? t0 : tmin;
    tmax = t1 < tmax ? t1 : tmax;
    if (tmax < tmin) return false;
  }
  return true;
}
// `d` must already be normalised exactly like the reference does (d / |d|).  `tris`: optional packed triangle vertices
// ([F][3] float4, written by bvh_build behind the nodes): one contiguous 48-byte read per leaf instead of an index triple
// plus three scattered vertex reads.
__device__ __forceinline__ RayHit bvh_trace(const void* __restrict__ nodes_v, const float* __restrict__ vert, const int* __restrict__ tri,
                            const float* o, const float* d, const float4* __restrict__ tris = nullptr) {
  const float4* nodes = static_cast<const float4*>(nodes_v);
  float inv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float di = d[i];
    if (di == 0.0f) di = 0.000001f;
    inv[i] = 1.0f / di;
  }
  int stack[64];
  int count = 0;
  stack[count++] = 0;
  float closest = 1e9f;
  RayHit h;
  h.any = 0; h.tid = -1; h.t = 0.f; h.u = 0.f; h.v = 0.f;
  while (count > 0) {
    const int n = stack[--count];
    const float4 q0 = __ldg(nodes + static_cast<size_t>(n) * 3), q1 = __ldg(nodes + static_cast<size_t>(n) * 3 + 1);
    const float bb[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
    if (!aabb_hit_dev(o, inv, 0.0f, closest, bb)) continue;
    const int l = __float_as_int(q1.z), r = __float_as_int(q1.w);
    if (l != 0 && r != 0) {
      if (count + 2 <= 64) {
        stack[count++] = l;
        stack[count++] = r;
      }
    } else if (l == 0 && r == 0) {
      const int p = __float_as_int(__ldg(nodes + static_cast<size_t>(n) * 3 + 2).x);
      float a[3], b[3], c[3];
      if (tris) {
        const float4 ta = __ldg(tris + static_cast<size_t>(p) * 3), tb = __ldg(tris + static_cast<size_t>(p) * 3 + 1),
                     tc = __ldg(tris + static_cast<size_t>(p) * 3 + 2);
        a[0] = ta.x; a[1] = ta.y; a[2] = ta.z; b[0] = tb.x; b[1] = tb.y; b[2] = tb.z; c[0] = tc.x; c[1] = tc.y; c[2] = tc.z;
      } else {
        const float *pa = vert + static_cast<size_t>(tri[p * 3]) * 3, *pb = vert + static_cast<size_t>(tri[p * 3 + 1]) * 3,
                    *pc = vert + static_cast<size_t>(tri[p * 3 + 2]) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = pa[k]; b[k] = pb[k]; c[k] = pc[k]; }
      }
      const float e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
      const float e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
      const float px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
      const float det = dot3f(e1x, e1y, e1z, px, py, pz);
      const float eps = 1e-9f;
      if (det > -eps && det < eps) continue;
      const float idet = 1.0f / det;
      const float tx = o[0] - a[0], ty = o[1] - a[1], tz = o[2] - a[2];
      const float u = dot3f(tx, ty, tz, px, py, pz) * idet;
      if (u < 0 || u > 1) continue;
      const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
      const float v = dot3f(d[0], d[1], d[2], qx, qy, qz) * idet;
      if (v < 0 || u + v > 1) continue;
      const float t = dot3f(e2x, e2y, e2z, qx, qy, qz) * idet;   // no t-range test (reference quirk)
      closest = t < closest ? t : closest;
      h.any = 1; h.tid = p; h.t = closest; h.u = u; h.v = v;   // last accepted leaf wins (reference quirk)
    }
  }
  return h;
}

// Exact nearest neighbour of q among the points of a point-LBVH: fp32 squared distance ((dx^2+dy^2)+dz^2), lowest point id on
// ties (`ids` maps the tree's point index to the caller's id).  Boxes are pruned only when their (slightly deflated)
// distance bound exceeds the best distance, so rounding can never drop the true nearest point.
__device__ __forceinline__ int nn_trace(const void* __restrict__ nodes_v, const float* __restrict__ pts,
                                        const int* __restrict__ ids, const float* q, float* best_d2_out) {
  const float4* nodes = static_cast<const float4*>(nodes_v);
  int stack[64];
  int count = 0;
  stack[count++] = 0;
  float best = INFINITY;
  int best_id = -1;
  while (count > 0) {
    const int n = stack[--count];
    const float4 q0 = __ldg(nodes + static_cast<size_t>(n) * 3), q1 = __ldg(nodes + static_cast<size_t>(n) * 3 + 1);
    const float bx = fmaxf(fmaxf(q0.x - q[0], q[0] - q0.w), 0.f), by = fmaxf(fmaxf(q0.y - q[1], q[1] - q1.x), 0.f),
                bz = fmaxf(fmaxf(q0.z - q[2], q[2] - q1.y), 0.f);
    const float bd = ((bx * bx + by * by) + bz * bz) * 0.999999f;
    if (bd > best) continue;
    const int l = __float_as_int(q1.z), r = __float_as_int(q1.w);
    if (l == 0 && r == 0) {
      const int pidx = __float_as_int(__ldg(nodes + static_cast<size_t>(n) * 3 + 2).x);
      const float dx = pts[static_cast<size_t>(pidx) * 3] - q[0], dy = pts[static_cast<size_t>(pidx) * 3 + 1] - q[1],
                  dz = pts[static_cast<size_t>(pidx) * 3 + 2] - q[2];
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      const int id = ids ? ids[pidx] : pidx;
      if (d2 < best || (d2 == best && id < best_id)) { best = d2; best_id = id; }
    } else if (count + 2 <= 64) {
      // visit the nearer child first: push the farther one below it
      const float4 a0 = __ldg(nodes + static_cast<size_t>(l) * 3), a1 = __ldg(nodes + static_cast<size_t>(l) * 3 + 1);
      const float ax = fmaxf(fmaxf(a0.x - q[0], q[0] - a0.w), 0.f), ay = fmaxf(fmaxf(a0.y - q[1], q[1] - a1.x), 0.f),
                  az = fmaxf(fmaxf(a0.z - q[2], q[2] - a1.y), 0.f);
      const float4 c0 = __ldg(nodes + static_cast<size_t>(r) * 3), c1 = __ldg(nodes + static_cast<size_t>(r) * 3 + 1);
      const float cx = fmaxf(fmaxf(c0.x - q[0], q[0] - c0.w), 0.f), cy = fmaxf(fmaxf(c0.y - q[1], q[1] - c1.x), 0.f),
                  cz = fmaxf(fmaxf(c0.z - q[2], q[2] - c1.y), 0.f);
      const float dl = (ax * ax + ay * ay) + az * az, dr = (cx * cx + cy * cy) + cz * cz;
      if (dl <= dr) { stack[count++] = r; stack[count++] = l; }
      else { stack[count++] = l; stack[count++] = r; }
    }
  }
  if (best_d2_out) *best_d2_out = best;
  return best_id;
}

}  // namespace utx
