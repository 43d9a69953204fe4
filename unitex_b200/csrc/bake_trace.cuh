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

// Moller-Trumbore exactly as the reference evaluates it (intersect_test2.slang:26-61): 1e-9 determinant epsilon, u, v and u + v
// range tests, NO t-range test (quirk a); every product / sum separately rounded (-fmad=false).
__device__ __forceinline__ bool triangle_hit_dev(const float* __restrict__ vert, const int* __restrict__ tri, int p, const float* o,
                                                 const float* d, float& t_out, float& u_out, float& v_out) {
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
  if (det > -eps && det < eps) return false;
  const float idet = 1.0f / det;
  const float tx = o[0] - a[0], ty = o[1] - a[1], tz = o[2] - a[2];
  const float u = dot3f(tx, ty, tz, px, py, pz) * idet;
  if (u < 0 || u > 1) return false;
  const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
  const float v = dot3f(d[0], d[1], d[2], qx, qy, qz) * idet;
  if (v < 0 || u + v > 1) return false;
  t_out = dot3f(e2x, e2y, e2z, qx, qy, qz) * idet;   // no t-range test (reference quirk)
  u_out = u;
  v_out = v;
  return true;
}

// Traversal layout ("wide" records) written behind the reference-layout nodes by bvh_build / point_bvh_build: one 128 B
// record per INTERNAL binary node n holding the boxes and references of its (up to) four grandchildren -- a child that is a leaf
// stays as one entry -- in the order [left part, right part]:
//   W[0..7]              header: root box (6 floats), root reference, pad            (128 B, so records are 128 B aligned)
//   W[8 + 8 n ..]        24 floats = 4 boxes (min xyz, max xyz), 4 int references, pad
// reference >= 0: internal node index, < 0: ~prim of a leaf, WIDE_EMPTY: unused slot (its box is inverted and fails every test).
constexpr int WIDE_EMPTY = static_cast<int>(0x80000000u);
__host__ __device__ __forceinline__ size_t wide_offset_bytes(int F) {      // from the start of the nodes buffer
  return (static_cast<size_t>(2 * F - 1) * 48 + 127) / 128 * 128;
}
struct WideRec {
  float bb[24];
  int ref[4];
};
__device__ __forceinline__ WideRec wide_load(const float4* __restrict__ W, int n) {
  const float4* nd = W + 8 + static_cast<size_t>(n) * 8;
  WideRec r;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float4 q = __ldg(nd + k);
    r.bb[4 * k] = q.x; r.bb[4 * k + 1] = q.y; r.bb[4 * k + 2] = q.z; r.bb[4 * k + 3] = q.w;
  }
  const float4 q = __ldg(nd + 6);
  r.ref[0] = __float_as_int(q.x); r.ref[1] = __float_as_int(q.y); r.ref[2] = __float_as_int(q.z); r.ref[3] = __float_as_int(q.w);
  return r;
}

// Closest-hit query.  One 128 B fetch per visited record yields four slab tests: two levels of the reference's binary walk per
// dependent memory round trip (the walk is latency-bound: ~150 dependent visits per ray, profiles/r01_summary.md).  An entry
// that already fails is never pushed, one that passes is pushed with its entry distance te and re-checked against the closest t
// of the moment it is popped.
// This visits exactly the leaves the reference's loop (intersect_test2.slang:63-146) visits, in the same order, with the same
// closest t at each visit:
//  * the reference tests a node's own box when it pops it (closest = value at pop time); a box fails iff
//    min(closest_pop, tx) < te; closest only shrinks, so "fails at push time" implies "fails at pop time", and for the rest the
//    pop-time comparison `closest < te` completes the identical predicate (tx >= te is already known);
//  * skipping the intermediate child's own test is exact: a grandchild's box lies inside the child's box (refit takes exact
//    min / max), the slab arithmetic is monotone in the box coordinates, so te_grandchild >= te_child and
//    tx_grandchild <= tx_child: whenever the child's box fails, both grandchildren fail on their own;
//  * order: the reference pushes left, right and pops right first, i.e. a depth-first walk that takes the right subtree
//    before the left one; pushing the entries [LL, LR, RL, RR] in that order and popping from the top is the same walk.
// Moller-Trumbore without a t-range test and "last accepted leaf wins" are the reference's quirks.  `d` must already be
// normalised exactly like the reference does (d / |d|).
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
  // The entry that would be popped next is kept in registers (cur / cur_te) instead of being pushed and popped straight
  // away: of a record's passing entries all but the last go to the stack, the last becomes `cur`.  Same order, ~half the
  // local-memory traffic (the stack was 41 % of the L1 wavefronts of this L1-bound kernel, profiles/r01_bake_ray_nn.ncu-rep).
  bool have = false;
  int ref = 0;
  float cur_te = 0.f;
  if (count > 0) { have = true; ref = sref[0]; cur_te = ste[0]; count = 0; }
  for (;;) {
    if (!have) {
      if (count == 0) break;
      --count;
      ref = sref[count];
      cur_te = ste[count];
    }
    have = false;
    if (closest < cur_te) continue;
    if (ref >= 0) {
      const WideRec r = wide_load(W, ref);
      int pend_ref = 0;
      float pend_te = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float te, tx;
        slab_params(o, inv, r.bb + 6 * k, te, tx);
        if (r.ref[k] != WIDE_EMPTY && slab_pass(te, tx, closest)) {
          if (have && count < 64) { sref[count] = pend_ref; ste[count] = pend_te; ++count; }
          pend_ref = r.ref[k]; pend_te = te; have = true;
        }
      }
      ref = pend_ref;
      cur_te = pend_te;
    } else {
      const int p = ~ref;
      float t, u, v;
      if (!triangle_hit_dev(vert, tri, p, o, d, t, u, v)) continue;
      closest = t < closest ? t : closest;
      h.any = 1; h.tid = p; h.t = closest; h.u = u; h.v = v;   // last accepted leaf wins (reference quirk)
    }
  }
  return h;
}

// ---------------------------------------------------------------------------------------------- point search tree
// The exact nearest-neighbour structure of the bake (torch_kdtree's role, pcd/knn/__init__.py:36-39,93-95): points sorted by
// Morton code, cut into clusters of PT_CLUSTER consecutive points, and a Karras LBVH over the CLUSTER boxes (same hierarchy /
// refit kernels as the triangle tree).  A leaf is scored by reading its points from one contiguous 128 B run of `spts`
// (x, y, z, caller's id) -- an eighth of the nodes to build and to walk compared with one leaf per point.
//   nodes  [2 n_c - 1] x 48 B (box, left, right, prim = cluster), absent when n_c == 1
//   spts   [n] float4, Morton order
// Results do not depend on the tree: fp32 squared distance ((dx^2+dy^2)+dz^2), ties by lowest id; a box is pruned only when
// its slightly deflated distance bound exceeds the current worst kept distance, so rounding can never drop a true neighbour.
#ifndef UTX_PT_CLUSTER
#define UTX_PT_CLUSTER 8      // measured on the bench bake / the reference's teaser mesh: 8 -> 8.6 / 11.0 ms, 16 -> 9.4 / 11.4, 32 -> 10.6 / 12.5
#endif
constexpr int PT_CLUSTER = UTX_PT_CLUSTER;
struct PointTree {
  const float4* wide;   // traversal records of the cluster tree (layout as above; leaf reference = ~cluster), unused when n_c == 1
  const float4* spts;
  int n_c, n;
};
__device__ __forceinline__ float box_dist2(const float* bb, const float* q) {
  const float ax = fmaxf(fmaxf(bb[0] - q[0], q[0] - bb[3]), 0.f), ay = fmaxf(fmaxf(bb[1] - q[1], q[1] - bb[4]), 0.f),
              az = fmaxf(fmaxf(bb[2] - q[2], q[2] - bb[5]), 0.f);
  return (ax * ax + ay * ay) + az * az;
}
// Generic walk: `worst()` = current pruning distance, `offer(d2, id)` = score one point.  Entries carry their bound and are
// re-checked against worst() when popped; of a record's entries the nearest is pushed last (popped first).
template <typename Worst, typename Offer>
__device__ __forceinline__ void point_tree_walk(const PointTree& pt, const float* q, Worst worst, Offer offer,
                                                const float* order_q = nullptr) {
  auto leaf = [&](int c) {
    const int j0 = c * PT_CLUSTER, j1 = min(j0 + PT_CLUSTER, pt.n);
    for (int j = j0; j < j1; ++j) {
      const float4 p = __ldg(pt.spts + j);
      const float dx = p.x - q[0], dy = p.y - q[1], dz = p.z - q[2];
      offer((dx * dx + dy * dy) + dz * dz, __float_as_int(p.w));
    }
  };
  if (pt.n_c <= 1) {
    if (pt.n > 0) leaf(0);
    return;
  }
  // `order_q` (default: q itself) decides which entry of a record is visited first; a warp of neighbouring queries passes ONE
  // common point so that its lanes walk the tree in the same order and stay converged (the result does not depend on the
  // order).  The entry to visit next stays in registers, the others go to the local-memory stack.
  const float* oq = order_q ? order_q : q;
  int sn[64];
  float sb[64];
  int count = 0;
  bool have = true;
  int ref = 0;
  float bound = 0.f;
  for (;;) {
    if (!have) {
      if (count == 0) break;
      --count;
      ref = sn[count];
      bound = sb[count];
    }
    have = false;
    if (bound > worst()) continue;
    if (ref < 0) { leaf(~ref); continue; }
    const WideRec r = wide_load(pt.wide, ref);
    float bd[4];
    int nearest = 0;
    float nkey = INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool used = r.ref[k] != WIDE_EMPTY;                       // (an inverted box is not "far" for box_dist2)
      bd[k] = used ? box_dist2(r.bb + 6 * k, q) * 0.999999f : INFINITY;
      const float key = used ? (order_q ? box_dist2(r.bb + 6 * k, oq) : bd[k]) : INFINITY;
      if (key < nkey) { nkey = key; nearest = k; }
    }
    const float w = worst();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (r.ref[k] == WIDE_EMPTY || !(bd[k] <= w)) continue;
      if (k == nearest) { have = true; ref = r.ref[k]; bound = bd[k]; }
      else if (count < 64) { sn[count] = r.ref[k]; sb[count] = bd[k]; ++count; }
    }
  }
}
// exact 1-NN: returns the id (-1 when the tree is empty), *best_d2_out = squared distance
__device__ __forceinline__ int nn_trace(const PointTree& pt, const float* q, float* best_d2_out, const float* order_q = nullptr) {
  float best = INFINITY;
  int best_id = -1;
  point_tree_walk(pt, q, [&]() { return best; },
                  [&](float d2, int id) {
                    if (d2 < best || (d2 == best && id < best_id) || best_id < 0) { best = d2; best_id = id; }
                  }, order_q);
  if (best_d2_out) *best_d2_out = best;
  return best_id;
}
// exact k-NN (k <= KNN_MAX), ascending by (squared distance, id) -- the order `knn(src, dst, k)` of the reference returns
// (pcd/knn/__init__.py:85-95, sorted by distance; ties by lowest id here, torch_kdtree@86961f7d [ext] leaves them unspecified).
// bd / bi: caller's arrays of k entries; entries beyond the number of points keep (INFINITY, -1).
constexpr int KNN_MAX = 32;
__device__ __forceinline__ void knn_trace(const PointTree& pt, const float* q, int k, float* bd, int* bi) {
  for (int j = 0; j < k; ++j) { bd[j] = INFINITY; bi[j] = -1; }
  point_tree_walk(pt, q, [&]() { return bd[k - 1]; },
                  [&](float d2, int id) {
                    const bool better = d2 < bd[k - 1] || (d2 == bd[k - 1] && (bi[k - 1] < 0 || id < bi[k - 1]));
                    if (!better) return;
                    int j = k - 1;
                    while (j > 0 && (bd[j - 1] > d2 || (bd[j - 1] == d2 && (bi[j - 1] < 0 || bi[j - 1] > id)))) {
                      bd[j] = bd[j - 1];
                      bi[j] = bi[j - 1];
                      --j;
                    }
                    bd[j] = d2;
                    bi[j] = id;
                  });
}

}  // namespace utx
