// Closest-hit traversal of the LBVH, shared by the standalone intersect kernel and the fused UV-bake texel kernel.
// Order and quirks follow the reference's bvh_hit
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
// Entry / exit parameters of the slab test without the running closest-t: te = max(0, near_x, near_y, near_z),
// tx = min(far_x, far_y, far_z), built with the reference's own compare-and-select chain (a NaN operand is skipped, exactly
// as in aabb_hit_dev).  The reference's test `tmax < tmin` after the last axis, with tmax seeded by the closest t, is then
//   fail  <=>  min(closest, tx) < te
// (its per-axis early exits cannot fail earlier than the final comparison: tmin only grows, tmax only shrinks).
__device__ __forceinline__ void slab_params(const float* o, const float* inv, const float* bb, float& te, float& tx) {
  te = 0.0f;
  tx = INFINITY;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t0 = (bb[i] - o[i]) * inv[i], t1 = (bb[3 + i] - o[i]) * inv[i];
    if (inv[i] < 0.0f) { const float t = t1; t1 = t0; t0 = t; }
    te = t0 > te ? t0 : te;
    tx = t1 < tx ? t1 : tx;
  }
}
__device__ __forceinline__ bool slab_pass(float te, float tx, float closest) {
  const float m = tx < closest ? tx : closest;
  return !(m < te);
}

// Closest-hit query on the traversal layout bvh_build writes behind the reference-layout nodes ("wide" nodes, 64 B per
// INTERNAL node: both children's boxes + both child references, leaves folded into their parents as ~prim):
//   W[0..1]            root box (6 floats), root reference, pad
//   W[2 + 4 n ..]      internal node n:  L.min xyz, L.max xyz, R.min xyz, R.max xyz, refL, refR, pad, pad
// One 64 B fetch per visited internal node yields both children's slab tests; a child that already fails is never pushed,
// one that passes is pushed with its entry distance te and re-checked against the closest t of the moment it is popped.
// This visits exactly the leaves the reference's loop visits, in the same order, with the same closest t at each visit:
// the reference tests a node's own box when it pops it (closest = value at pop time); a box fails iff
// min(closest_pop, tx) < te; closest only shrinks, so "fails at push time" implies "fails at pop time", and for the rest
// the pop-time comparison `closest < te` completes the identical predicate (tx >= te is already known).  Push order
// (left, then right; right popped first), Moller-Trumbore without a t-range test and "last accepted leaf wins" are the
// reference's (intersect_test2.slang:63-146).  `d` must already be normalised exactly like the reference does (d / |d|).
__device__ __forceinline__ RayHit bvh_trace(const float4* __restrict__ W, const float* __restrict__ vert, const int* __restrict__ tri,
                                            const float* o, const float* d) {
  float inv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float di = d[i];
    if (di == 0.0f) di = 0.000001f;      // intersect_test2.slang:18-21; the quotient is the same IEEE value at every node
    inv[i] = 1.0f / di;
  }
  int sref[64];
  float ste[64];
  int count = 0;
  float closest = 1e9f;
  RayHit h;
  h.any = 0; h.tid = -1; h.t = 0.f; h.u = 0.f; h.v = 0.f;
  {
    const float4 r0 = __ldg(W), r1 = __ldg(W + 1);
    const float bb[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
    float te, tx;
    slab_params(o, inv, bb, te, tx);
    if (slab_pass(te, tx, closest)) { sref[0] = __float_as_int(r1.z); ste[0] = te; count = 1; }
  }
  while (count > 0) {
    --count;
    const int ref = sref[count];
    if (closest < ste[count]) continue;
    if (ref >= 0) {
      const float4* nd = W + 2 + static_cast<size_t>(ref) * 4;
      const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
      const float bl[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y}, br[6] = {q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
      float tel, txl, ter, txr;
      slab_params(o, inv, bl, tel, txl);
      slab_params(o, inv, br, ter, txr);
      if (count + 2 <= 64) {
        if (slab_pass(tel, txl, closest)) { sref[count] = __float_as_int(q3.x); ste[count] = tel; ++count; }
        if (slab_pass(ter, txr, closest)) { sref[count] = __float_as_int(q3.y); ste[count] = ter; ++count; }
      }
    } else {
      const int p = ~ref;
      const float *pa = vert + static_cast<size_t>(tri[p * 3]) * 3, *pb = vert + static_cast<size_t>(tri[p * 3 + 1]) * 3,
                  *pc = vert + static_cast<size_t>(tri[p * 3 + 2]) * 3;
      float a[3], b[3], c[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { a[k] = pa[k]; b[k] = pb[k]; c[k] = pc[k]; }
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

// Exact k nearest neighbours (k <= KNN_MAX) of q among the points of a point-LBVH, ascending by (fp32 squared distance,
// point id) -- the order `knn(src, dst, k)` of the reference returns (pcd/knn/__init__.py:85-95, sorted by distance; ties by
// lowest id here, torch_kdtree@86961f7d [ext] leaves them unspecified).  bd / bi: caller's arrays of k entries; entries
// beyond the number of points found keep (INFINITY, -1).  Same conservative pruning as nn_trace, against the current
// k-th best distance, so the result does not depend on the visit order.
constexpr int KNN_MAX = 32;
__device__ __forceinline__ void knn_trace(const void* __restrict__ nodes_v, const float* __restrict__ pts,
                                          const int* __restrict__ ids, const float* q, int k, float* bd, int* bi) {
  const float4* nodes = static_cast<const float4*>(nodes_v);
  for (int j = 0; j < k; ++j) { bd[j] = INFINITY; bi[j] = -1; }
  int stack[64];
  int count = 0;
  stack[count++] = 0;
  while (count > 0) {
    const int n = stack[--count];
    const float4 q0 = __ldg(nodes + static_cast<size_t>(n) * 3), q1 = __ldg(nodes + static_cast<size_t>(n) * 3 + 1);
    const float bx = fmaxf(fmaxf(q0.x - q[0], q[0] - q0.w), 0.f), by = fmaxf(fmaxf(q0.y - q[1], q[1] - q1.x), 0.f),
                bz = fmaxf(fmaxf(q0.z - q[2], q[2] - q1.y), 0.f);
    const float bound = ((bx * bx + by * by) + bz * bz) * 0.999999f;
    if (bound > bd[k - 1]) continue;
    const int l = __float_as_int(q1.z), r = __float_as_int(q1.w);
    if (l == 0 && r == 0) {
      const int pidx = __float_as_int(__ldg(nodes + static_cast<size_t>(n) * 3 + 2).x);
      const float dx = pts[static_cast<size_t>(pidx) * 3] - q[0], dy = pts[static_cast<size_t>(pidx) * 3 + 1] - q[1],
                  dz = pts[static_cast<size_t>(pidx) * 3 + 2] - q[2];
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      const int id = ids ? ids[pidx] : pidx;
      const bool better = d2 < bd[k - 1] || (d2 == bd[k - 1] && (bi[k - 1] < 0 || id < bi[k - 1]));
      if (better) {
        int j = k - 1;
        while (j > 0 && (bd[j - 1] > d2 || (bd[j - 1] == d2 && (bi[j - 1] < 0 || bi[j - 1] > id)))) {
          bd[j] = bd[j - 1];
          bi[j] = bi[j - 1];
          --j;
        }
        bd[j] = d2;
        bi[j] = id;
      }
    } else if (count + 2 <= 64) {
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
}

}  // namespace utx
