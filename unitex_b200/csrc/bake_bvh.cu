// LBVH build + closest-hit query: the reference's ray-tracer plug-in (RayTracing.intersects_closest / update_raw,
// TextureTools/texturetools/raytracing/__init__.py:12-80, default backend rt_aprmis/__init__.py:10-86, kernels
// rt_aprmis/bvhworkers/*.slang) rebuilt for B200.
//
// Same tree as the reference, built differently:
//   * per-triangle AABB + scene bounds in one pass (atomics on order-preserving ints instead of 6 torch reductions)
//   * 30-bit Morton codes (identical arithmetic), sorted with a multi-block device radix sort (the reference runs a
//     single-workgroup sort on ONE SM -- its slowest stage); stable, so ties keep element order exactly like its LSD sort
//   * Karras-2012 hierarchy, identical delta / range / split rules incl. the duplicate-code tie-break on sorted index
//   * ONE bottom-up refit pass with per-node arrival counters instead of one launch per tree level + host sync
//   * nodes packed in 48 B (aabb[6], left, right, prim) instead of two strided arrays (export / 1-NN search), plus a
//     traversal layout with both children's boxes in the parent (64 B per internal node, leaves folded in): one fetch and
//     two slab tests per visited node, children that fail are never pushed (bake_trace.cuh proves the visit order is kept)
// The traversal keeps the reference's order and quirks (intersect_test2.slang:63-146), because they decide which
// triangle id a ray reports and `rays_tid == tid_2d` is the bake's visibility test (renderer_inverse.py:321-323).
// Built with -fmad=false (see bake_raster.cu).
#include <cub/device/device_radix_sort.cuh>

#include "bake_trace.cuh"
#include "common.h"
#include "kernels.h"

namespace utx {
static size_t al256(size_t v) { return (v + 255) / 256 * 256; }
namespace {

struct __align__(16) Node {
  float bb[6];
  int left, right, prim;
  int pad[3];
};
static_assert(sizeof(Node) == 48, "node must be 48 bytes");

__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void bounds_init_kernel(int* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = 0x7fffffff;
  else if (threadIdx.x < 6) bounds[threadIdx.x] = static_cast<int>(0x80000000u);
}

// scene AABB: warp shuffle, then one atomic per BLOCK per component (one per warp put 6 x F/32 atomics on six addresses --
// 0.47 ms of the 3.5 M-point tree build, profiles/r01_bake_launches_final.csv)
__device__ __forceinline__ void block_bounds(const float* mn, const float* mx, bool valid, int* __restrict__ bounds) {
  __shared__ float red[6][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = valid ? mn[a] : INFINITY, hi = valid ? mx[a] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { red[a][warp] = lo; red[3 + a][warp] = hi; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int a = threadIdx.x;
    float v = red[a][0];
    const int nw = blockDim.x >> 5;
    for (int w = 1; w < nw; ++w) v = a < 3 ? fminf(v, red[a][w]) : fmaxf(v, red[a][w]);
    if (a < 3) atomicMin(bounds + a, f2ord(v));
    else atomicMax(bounds + a, f2ord(v));
  }
}

// get_elements.slang:1-40
__global__ void __launch_bounds__(256) elements_kernel(const float* __restrict__ vert, const int* __restrict__ tri, int F,
                                                       float* __restrict__ eab, int* __restrict__ bounds) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  float mn[3] = {1e9f, 1e9f, 1e9f}, mx[3] = {-1e9f, -1e9f, -1e9f};
  if (f < F) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* v = vert + static_cast<size_t>(tri[f * 3 + k]) * 3;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = fminf(mn[a], v[a]);
        mx[a] = fmaxf(mx[a], v[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float lo = fminf(mn[a], mx[a]), hi = fmaxf(mn[a], mx[a]);
      eab[static_cast<size_t>(f) * 6 + a] = lo;
      eab[static_cast<size_t>(f) * 6 + 3 + a] = hi;
      mn[a] = lo;
      mx[a] = hi;
    }
  }
  block_bounds(mn, mx, f < F, bounds);
}

// points as degenerate boxes (nearest-neighbour tree over the visible texels, bake_uv.cu)
__global__ void __launch_bounds__(256) point_elements_kernel(const float* __restrict__ pts, int n, float* __restrict__ eab,
                                                             int* __restrict__ bounds) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  float p[3] = {0.f, 0.f, 0.f};
  if (f < n) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p[a] = pts[static_cast<size_t>(f) * 3 + a];
      eab[static_cast<size_t>(f) * 6 + a] = p[a];
      eab[static_cast<size_t>(f) * 6 + 3 + a] = p[a];
    }
  }
  block_bounds(p, p, f < n, bounds);
}

__device__ __forceinline__ unsigned expand_bits(unsigned v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
// lbvh_morton_codes.slang:24-79
__global__ void __launch_bounds__(256) morton_kernel(const float* __restrict__ eab, const int* __restrict__ bounds, int F,
                                                     unsigned* __restrict__ codes, unsigned* __restrict__ elem) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float m[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = eab[static_cast<size_t>(f) * 6 + a], hi = eab[static_cast<size_t>(f) * 6 + 3 + a];
    const float center = lo + 0.5f * (hi - lo);
    const float gmin = ord2f(bounds[a]), gmax = ord2f(bounds[3 + a]);
    float x = (center - gmin) / (gmax - gmin);
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    m[a] = x;
  }
  codes[f] = expand_bits(static_cast<unsigned>(m[0])) * 4 + expand_bits(static_cast<unsigned>(m[1])) * 2 +
             expand_bits(static_cast<unsigned>(m[2]));
  elem[f] = static_cast<unsigned>(f);
}

__device__ __forceinline__ int find_msb(unsigned v) { return v == 0 ? -1 : 31 - __clz(v); }
__device__ __forceinline__ int delta(int i, unsigned code_i, int j, int n, const unsigned* __restrict__ codes) {
  if (j < 0 || j > n - 1) return -1;
  const unsigned code_j = codes[j];
  if (code_i == code_j) return 32 + 31 - find_msb(static_cast<unsigned>(i) ^ static_cast<unsigned>(j));
  return 31 - find_msb(code_i ^ code_j);
}

// lbvh_hierarchy.slang:109-244
__global__ void __launch_bounds__(256) hierarchy_kernel(int F, const unsigned* __restrict__ codes,
                                                        const unsigned* __restrict__ elem, const float* __restrict__ eab,
                                                        Node* __restrict__ nodes, int* __restrict__ parent,
                                                        int* __restrict__ arrivals) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= F) return;
  const int LEAF = F - 1;
  {
    const int e = static_cast<int>(elem[g]);
    Node n;
#pragma unroll
    for (int a = 0; a < 6; ++a) n.bb[a] = eab[static_cast<size_t>(e) * 6 + a];
    n.left = 0; n.right = 0; n.prim = e;
    n.pad[0] = n.pad[1] = n.pad[2] = 0;
    nodes[LEAF + g] = n;
  }
  if (g == 0) parent[0] = 0;
  if (g >= F - 1) return;
  arrivals[g] = 0;
  const unsigned code = codes[g];
  const int dL = delta(g, code, g - 1, F, codes), dR = delta(g, code, g + 1, F, codes);
  const int d = (dR >= dL) ? 1 : -1;
  const int dmin = min(dL, dR);
  int lmax = 2;
  while (delta(g, code, g + lmax * d, F, codes) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t > 0; t >>= 1)
    if (delta(g, code, g + (l + t) * d, F, codes) > dmin) l += t;
  const int j = g + l * d;
  const int first = min(g, j), last = max(g, j);
  const unsigned fcode = codes[first];
  const int common = delta(first, fcode, last, F, codes);
  int split = first, stride = last - first;
  do {
    stride = (stride + 1) >> 1;
    const int ns = split + stride;
    if (ns < last && delta(first, fcode, ns, F, codes) > common) split = ns;
  } while (stride > 1);
  const int ca = (split == first) ? LEAF + split : split;
  const int cb = (split + 1 == last) ? LEAF + split + 1 : split + 1;
  Node n;
#pragma unroll
  for (int a = 0; a < 3; ++a) { n.bb[a] = 1e9f; n.bb[3 + a] = -1e9f; }
  n.left = ca; n.right = cb; n.prim = 0;
  n.pad[0] = n.pad[1] = n.pad[2] = 0;
  nodes[g] = n;
  parent[ca] = g;
  parent[cb] = g;
}

// lbvh_bounding_boxes.slang:149-389 collapsed into one pass: the second thread to arrive at a node unions its children.
__global__ void __launch_bounds__(256) refit_kernel(int F, Node* __restrict__ nodes, const int* __restrict__ parent,
                                                    int* __restrict__ arrivals) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= F || F < 2) return;
  int n = parent[F - 1 + g];
  for (;;) {
    __threadfence();
    if (atomicAdd(arrivals + n, 1) == 0) return;
    __threadfence();
    const int ca = nodes[n].left, cb = nodes[n].right;
    volatile const float* A = nodes[ca].bb;
    volatile const float* B = nodes[cb].bb;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      nodes[n].bb[a] = fminf(A[a], B[a]);
      nodes[n].bb[3 + a] = fmaxf(A[3 + a], B[3 + a]);
    }
    if (n == 0) return;
    n = parent[n];
  }
}

// traversal layout behind the reference-layout nodes (bake_trace.cuh): 128 B header (root box, root reference) + 128 B per
// internal node holding the boxes and references of its up to four grandchildren, [left part, right part]
__global__ void __launch_bounds__(256) pack_wide_kernel(const Node* __restrict__ nodes, int F, float4* __restrict__ W) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int LEAF = F - 1;
  if (n == 0) {
    const Node r = nodes[0];
    const int ref = (F == 1) ? ~r.prim : 0;
    W[0] = make_float4(r.bb[0], r.bb[1], r.bb[2], r.bb[3]);
    W[1] = make_float4(r.bb[4], r.bb[5], __int_as_float(ref), 0.f);
  }
  if (n >= F - 1) return;
  int ent[4];
  int cnt = 0;
  const int ch[2] = {nodes[n].left, nodes[n].right};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (ch[s] >= LEAF) ent[cnt++] = ch[s];
    else { ent[cnt++] = nodes[ch[s]].left; ent[cnt++] = nodes[ch[s]].right; }
  }
  float v[28];
  for (int k = 0; k < 4; ++k) {
    if (k < cnt) {
      const Node c = nodes[ent[k]];
      for (int a = 0; a < 6; ++a) v[6 * k + a] = c.bb[a];
      v[24 + k] = __int_as_float(ent[k] >= LEAF ? ~c.prim : ent[k]);
    } else {
      for (int a = 0; a < 3; ++a) { v[6 * k + a] = INFINITY; v[6 * k + 3 + a] = -INFINITY; }
      v[24 + k] = __int_as_float(WIDE_EMPTY);
    }
  }
  float4* o = W + 8 + static_cast<size_t>(n) * 8;
  for (int k = 0; k < 7; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  o[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) export_kernel(const Node* __restrict__ nodes, int n, int* __restrict__ info,
                                                     float* __restrict__ aabb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Node nd = nodes[i];
  info[i * 3] = nd.left; info[i * 3 + 1] = nd.right; info[i * 3 + 2] = nd.prim;
#pragma unroll
  for (int a = 0; a < 6; ++a) aabb[static_cast<size_t>(i) * 6 + a] = nd.bb[a];
}

}  // namespace

namespace {
__global__ void __launch_bounds__(128) intersect_kernel(const void* __restrict__ nodes, const float* __restrict__ vert,
                                                        const int* __restrict__ tri, int F, const float* __restrict__ rays_o,
                                                        const float* __restrict__ rays_d, long long N,
                                                        unsigned char* __restrict__ hit, int* __restrict__ tid,
                                                        float* __restrict__ pos, float* __restrict__ uv) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float o[3] = {rays_o[r * 3], rays_o[r * 3 + 1], rays_o[r * 3 + 2]};
  float d[3] = {rays_d[r * 3], rays_d[r * 3 + 1], rays_d[r * 3 + 2]};
  const float len = sqrtf(dot3f(d[0], d[1], d[2], d[0], d[1], d[2]));
  d[0] = d[0] / len; d[1] = d[1] / len; d[2] = d[2] / len;
  const RayHit h = bvh_trace(reinterpret_cast<const float4*>(static_cast<const uint8_t*>(nodes) + wide_offset_bytes(F)), vert, tri, o, d);
  hit[r] = static_cast<unsigned char>(h.any);
  tid[r] = h.any ? h.tid : -1;
  pos[r * 3] = h.any ? o[0] + h.t * d[0] : 0.f;
  pos[r * 3 + 1] = h.any ? o[1] + h.t * d[1] : 0.f;
  pos[r * 3 + 2] = h.any ? o[2] + h.t * d[2] : 0.f;
  uv[r * 2] = h.u;
  uv[r * 2 + 1] = h.v;
}
}  // namespace


// reference-layout nodes (48 B each) + the traversal layout (128 B header + 128 B per internal node), 128 B aligned
size_t bvh_nodes_bytes(int F) { return wide_offset_bytes(F) + 128 + static_cast<size_t>(F) * 128; }

size_t bvh_workspace_bytes(int F) {
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<unsigned*>(nullptr), static_cast<unsigned*>(nullptr),
                                  static_cast<unsigned*>(nullptr), static_cast<unsigned*>(nullptr), F);
  return al256(static_cast<size_t>(F) * 24) + 4 * al256(static_cast<size_t>(F) * 4) + al256(static_cast<size_t>(2 * F) * 4) +
         al256(static_cast<size_t>(F) * 4) + 256 + al256(cub_bytes);
}

static int build_tree(const float* vert, const int* tri, const float* pts, int F, void* nodes_out, void* workspace,
                      size_t ws_bytes, cudaStream_t stream) {
  UTX_CHECK(F >= 2, "bvh_build: need at least 2 elements");
  UTX_CHECK(ws_bytes >= bvh_workspace_bytes(F), "bvh_build: workspace too small");
  UTX_CHECK((reinterpret_cast<uintptr_t>(nodes_out) & 15) == 0, "bvh_build: nodes must be 16B aligned");
  uint8_t* p = static_cast<uint8_t*>(workspace);
  float* eab = reinterpret_cast<float*>(p); p += al256(static_cast<size_t>(F) * 24);
  unsigned* codes = reinterpret_cast<unsigned*>(p); p += al256(static_cast<size_t>(F) * 4);
  unsigned* codes2 = reinterpret_cast<unsigned*>(p); p += al256(static_cast<size_t>(F) * 4);
  unsigned* elem = reinterpret_cast<unsigned*>(p); p += al256(static_cast<size_t>(F) * 4);
  unsigned* elem2 = reinterpret_cast<unsigned*>(p); p += al256(static_cast<size_t>(F) * 4);
  int* parent = reinterpret_cast<int*>(p); p += al256(static_cast<size_t>(2 * F) * 4);
  int* arrivals = reinterpret_cast<int*>(p); p += al256(static_cast<size_t>(F) * 4);
  int* bounds = reinterpret_cast<int*>(p); p += 256;
  size_t cub_bytes = ws_bytes - static_cast<size_t>(p - static_cast<uint8_t*>(workspace));
  const unsigned grid = (F + 255) / 256;
  bounds_init_kernel<<<1, 32, 0, stream>>>(bounds);
  if (pts) point_elements_kernel<<<grid, 256, 0, stream>>>(pts, F, eab, bounds);
  else elements_kernel<<<grid, 256, 0, stream>>>(vert, tri, F, eab, bounds);
  morton_kernel<<<grid, 256, 0, stream>>>(eab, bounds, F, codes, elem);
  UTX_CUDA(cub::DeviceRadixSort::SortPairs(p, cub_bytes, codes, codes2, elem, elem2, F, 0, 32, stream));
  Node* nodes = static_cast<Node*>(nodes_out);
  hierarchy_kernel<<<grid, 256, 0, stream>>>(F, codes2, elem2, eab, nodes, parent, arrivals);
  refit_kernel<<<grid, 256, 0, stream>>>(F, nodes, parent, arrivals);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int bvh_build(const float* vert, int V, const int* tri, int F, void* nodes_out, void* workspace, size_t ws_bytes,
              cudaStream_t stream) {
  (void)V;
  UTX_TRY(build_tree(vert, tri, nullptr, F, nodes_out, workspace, ws_bytes, stream));
  pack_wide_kernel<<<(F + 255) / 256, 256, 0, stream>>>(static_cast<const Node*>(nodes_out), F,
                                                       reinterpret_cast<float4*>(static_cast<uint8_t*>(nodes_out) + wide_offset_bytes(F)));
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- point search tree
namespace {
// one thread per cluster of PT_CLUSTER Morton-consecutive points: gathers them into `spts` (x, y, z, id), their box becomes the
// cluster's element box, the first point's code its Morton code (non-decreasing along the clusters; the hierarchy kernel breaks
// ties on the index exactly as for duplicate triangle codes)
__global__ void __launch_bounds__(256) cluster_kernel(const float* __restrict__ pts, const int* __restrict__ ids, int n, int n_c,
                                                      const unsigned* __restrict__ codes_sorted,
                                                      const unsigned* __restrict__ elem_sorted, float4* __restrict__ spts,
                                                      float* __restrict__ ceab, unsigned* __restrict__ ccodes,
                                                      unsigned* __restrict__ celem) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_c) return;
  const int j0 = c * PT_CLUSTER, j1 = min(j0 + PT_CLUSTER, n);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int j = j0; j < j1; ++j) {
    const int e = static_cast<int>(elem_sorted[j]);
    const float x = pts[static_cast<size_t>(e) * 3], y = pts[static_cast<size_t>(e) * 3 + 1], z = pts[static_cast<size_t>(e) * 3 + 2];
    spts[j] = make_float4(x, y, z, __int_as_float(ids ? ids[e] : e));
    lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z);
    hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ceab[static_cast<size_t>(c) * 6 + a] = lo[a];
    ceab[static_cast<size_t>(c) * 6 + 3 + a] = hi[a];
  }
  ccodes[c] = codes_sorted[j0];
  celem[c] = static_cast<unsigned>(c);
}
}  // namespace

PointTree point_tree_view(const void* nodes, int n) {
  PointTree pt;
  pt.n = n;
  pt.n_c = (n + PT_CLUSTER - 1) / PT_CLUSTER;
  const uint8_t* base = static_cast<const uint8_t*>(nodes);
  size_t off = 0;
  pt.wide = nullptr;
  if (pt.n_c >= 2) {
    off = wide_offset_bytes(pt.n_c);
    pt.wide = reinterpret_cast<const float4*>(base + off);
    off += 128 + static_cast<size_t>(pt.n_c) * 128;
  }
  pt.spts = reinterpret_cast<const float4*>(base + off);
  return pt;
}

// Search tree over n >= 1 points (`ids`: optional caller ids, default = point index).  nodes_out: bvh_nodes_bytes(max(n, 2))
// bytes, workspace: bvh_workspace_bytes(max(n, 2)).  Read it back with point_tree_view(nodes_out, n).
int point_bvh_build(const float* pts, const int* ids, int n, void* nodes_out, void* workspace, size_t ws_bytes,
                    cudaStream_t stream) {
  UTX_CHECK(n >= 1, "point_bvh_build: empty point set");
  UTX_CHECK(ws_bytes >= bvh_workspace_bytes(n < 2 ? 2 : n), "point_bvh_build: workspace too small");
  UTX_CHECK((reinterpret_cast<uintptr_t>(nodes_out) & 15) == 0, "point_bvh_build: nodes must be 16B aligned");
  const PointTree pt = point_tree_view(nodes_out, n);
  uint8_t* p = static_cast<uint8_t*>(workspace);
  const size_t F = n < 2 ? 2 : n;                       // same carve-up as build_tree (the sizes were computed for it)
  float* eab = reinterpret_cast<float*>(p); p += al256(F * 24);
  unsigned* codes = reinterpret_cast<unsigned*>(p); p += al256(F * 4);
  unsigned* codes2 = reinterpret_cast<unsigned*>(p); p += al256(F * 4);
  unsigned* elem = reinterpret_cast<unsigned*>(p); p += al256(F * 4);
  unsigned* elem2 = reinterpret_cast<unsigned*>(p); p += al256(F * 4);
  int* parent = reinterpret_cast<int*>(p); p += al256(2 * F * 4);
  int* arrivals = reinterpret_cast<int*>(p); p += al256(F * 4);
  int* bounds = reinterpret_cast<int*>(p); p += 256;
  size_t cub_bytes = ws_bytes - static_cast<size_t>(p - static_cast<uint8_t*>(workspace));
  const unsigned grid = (n + 255) / 256;
  bounds_init_kernel<<<1, 32, 0, stream>>>(bounds);
  point_elements_kernel<<<grid, 256, 0, stream>>>(pts, n, eab, bounds);
  morton_kernel<<<grid, 256, 0, stream>>>(eab, bounds, n, codes, elem);
  UTX_CUDA(cub::DeviceRadixSort::SortPairs(p, cub_bytes, codes, codes2, elem, elem2, n, 0, 32, stream));
  // the unsorted code / element arrays and the per-point boxes are dead now: the clusters' go there
  const unsigned cgrid = (pt.n_c + 255) / 256;
  cluster_kernel<<<cgrid, 256, 0, stream>>>(pts, ids, n, pt.n_c, codes2, elem2, const_cast<float4*>(pt.spts), eab, codes, elem);
  if (pt.n_c >= 2) {
    Node* nodes = static_cast<Node*>(nodes_out);
    hierarchy_kernel<<<cgrid, 256, 0, stream>>>(pt.n_c, codes, elem, eab, nodes, parent, arrivals);
    refit_kernel<<<cgrid, 256, 0, stream>>>(pt.n_c, nodes, parent, arrivals);
    pack_wide_kernel<<<cgrid, 256, 0, stream>>>(nodes, pt.n_c, const_cast<float4*>(pt.wide));
  }
  UTX_CUDA(cudaGetLastError());
  return 0;
}

namespace {
__global__ void __launch_bounds__(128) knn1_kernel(const PointTree pt, const float* __restrict__ dst, long long M,
                                                   long long* __restrict__ index, float* __restrict__ score) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float q[3] = {dst[i * 3], dst[i * 3 + 1], dst[i * 3 + 2]};
  float d2 = 0.f;
  index[i] = nn_trace(pt, q, &d2);
  score[i] = sqrtf(d2);
}
// one thread per query; the k candidates live in the thread's local arrays (k <= KNN_MAX)
__global__ void __launch_bounds__(128) knn_kernel(const PointTree pt, const float* __restrict__ dst, long long M, int k,
                                                  long long* __restrict__ index, float* __restrict__ score) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float q[3] = {dst[i * 3], dst[i * 3 + 1], dst[i * 3 + 2]};
  float bd[KNN_MAX];
  int bi[KNN_MAX];
  knn_trace(pt, q, k, bd, bi);
  for (int j = 0; j < k; ++j) {
    index[i * k + j] = bi[j];
    score[i * k + j] = sqrtf(bd[j]);
  }
}
}  // namespace

// knn(src, dst, 1) of pcd/knn/__init__.py:104-114: index int64 [M], score = distance [M]; lowest index on ties
int knn1(const float* src, int n_src, const float* dst, long long M, long long* index, float* score, void* nodes, void* workspace,
         size_t ws_bytes, cudaStream_t stream) {
  UTX_CHECK(n_src >= 1, "knn: empty source set");
  UTX_TRY(point_bvh_build(src, nullptr, n_src, nodes, workspace, ws_bytes, stream));
  if (M == 0) return 0;
  knn1_kernel<<<static_cast<unsigned>((M + 127) / 128), 128, 0, stream>>>(point_tree_view(nodes, n_src), dst, M, index, score);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// knn(src, dst, k) of pcd/knn/__init__.py:104-114 for 1 <= k <= min(32, n_src): index int64 [M,k], score = distance [M,k],
// each row ascending by (distance, index)
int knn(const float* src, int n_src, const float* dst, long long M, int k, long long* index, float* score, void* nodes,
        void* workspace, size_t ws_bytes, cudaStream_t stream) {
  UTX_CHECK(n_src >= 1, "knn: empty source set");
  UTX_CHECK(k >= 1 && k <= KNN_MAX, "knn: k must be in 1..32");
  UTX_CHECK(k <= n_src, "knn: k exceeds the number of source points");
  if (k == 1) return knn1(src, n_src, dst, M, index, score, nodes, workspace, ws_bytes, stream);
  UTX_TRY(point_bvh_build(src, nullptr, n_src, nodes, workspace, ws_bytes, stream));
  if (M == 0) return 0;
  knn_kernel<<<static_cast<unsigned>((M + 127) / 128), 128, 0, stream>>>(point_tree_view(nodes, n_src), dst, M, k, index, score);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int bvh_export(const void* nodes, int F, int* info, float* aabb, cudaStream_t stream) {
  const int n = 2 * F - 1;
  export_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const Node*>(nodes), n, info, aabb);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int bvh_intersect(const void* nodes, const float* vert, const int* tri, int F, const float* rays_o, const float* rays_d,
                  long long N, unsigned char* hit, int* tid, float* pos, float* uv, cudaStream_t stream) {
  if (N == 0) return 0;
  intersect_kernel<<<static_cast<unsigned>((N + 127) / 128), 128, 0, stream>>>(nodes, vert, tri, F, rays_o, rays_d, N, hit, tid,
                                                                               pos, uv);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace utx
