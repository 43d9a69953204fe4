// Fused UV bake: NVDiffRendererInverse.uv_to_pcd + bake_mv_to_uv_reproject_blur
// (TextureTools/texturetools/render/nvdiffrast/renderer_inverse.py:243-365, 574-633) without the per-view
// [n, H2D, W2D, *] rays / ndc / colour tensors and the ~40 masked_select / masked_scatter compactions of the reference.
//
//   texel pass      per covered texel: position + face normal from the UV raster, then per view: ray/normal angle test and
//                   projected bilinear fetch of alpha (texel_prep_kernel, which also appends the texel to the ray list of every
//                   view it faces); then one thread per listed ray: orthographic ray, LBVH closest hit, `tid == raster tid`
//                   (ray_kernel) -> one visibility bit and one alpha bit per view                      (:277-325, :343)
//   repair          the "misjudgment repair" convolutions as exact integer stencils on the 6-bit planes: k=3 ORs in a
//                   texel when any 8-neighbour is visible, k=5 when >= 6 of the 16 ring texels are      (:329-339)
//   compose         and-with-coverage/alpha, first-visible-view-wins in priority order, winning colour re-fetched (:591-603)
//   seam mask       3x3 boundary of every view's claim = "a 3x3 neighbour has a different owner", dilated 3x3, kept where
//                   the 7x7 erosion of the chart mask holds                                              (:435-444, :603-605)
//   nn fill         invisible covered texels take the colour of the 3-D nearest visible texel (exact 1-NN through the clustered
//                   point tree of bake_trace.cuh, lowest index on ties; queries compacted into a tile-ordered list) (:606-615)
//   lens blur       7x7 effective kernel of the 5-component complex separable blur, evaluated ONLY on seam texels
//                   (image/lens_blur.py:260-280; the reference blurs the whole atlas, then keeps seam texels) (:621-624)
//   pull-push       alpha-weighted 2x2 pyramid + bilinear up-fill of texels outside the charts (texture/stitching/mip.py:51-96)
// HBM-bound by contract for the streaming stages: one thread per texel, row-major so a warp reads/writes 32
// consecutive texels; the two tree walks (rays, nearest neighbour) run over compacted, tile-ordered work lists.
// Built with -fmad=false.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bake_trace.cuh"
#include "common.h"
#include "kernels.h"

namespace utx {
namespace {

constexpr int MAXV = 8;
#ifndef UTX_WALK_MIN_BLOCKS
#define UTX_WALK_MIN_BLOCKS 1      // resident CTAs per SM the two tree-walk kernels are compiled for (register cap).  Measured on
                                   // the bench bake: 1 (46-47 registers, 10 CTAs) 8.54 ms, 12 (40 registers, spills) 9.05, 16 (32) 10.25:
                                   // more resident warps thrash the L1 the walks live in
#endif

struct Views {
  int n;
  float mat[MAXV][16];   // P @ W2C, row-major
  float dir[MAXV][3];    // orthographic: -c2w[:3, 2]; perspective: the camera position c2w[:3, 3]
  int perspective;
  int priority[MAXV];
};

__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf((x * x + y * y) + z * z); }

// grid_sample(mode=bilinear, padding_mode=zeros, align_corners=False) of a [H, W, C<=4] image
template <int C>
__device__ __forceinline__ void bilinear(const float* __restrict__ img, int H, int W, float gx, float gy, float* out) {
  const float ix = ((gx + 1.0f) * W - 1.0f) / 2.0f, iy = ((gy + 1.0f) * H - 1.0f) / 2.0f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix, wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
  const float w[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};   // nw, ne, sw, se
  const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H) {
      const float* p = img + (static_cast<size_t>(ys[k]) * W + xs[k]) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) out[c] = out[c] + p[c] * w[k];
    }
  }
}

// ndc (x, y) of a texel in view `v`: barycentric interpolation of the per-vertex ndc like dr.interpolate(vertices_ndc) (:288)
__device__ __forceinline__ void texel_ndc(const float* m, const float* p0, const float* p1, const float* p2, float u, float v,
                                          float* gx, float* gy) {
  float nx[3], ny[3];
  const float* ps[3] = {p0, p1, p2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* p = ps[k];
    const float cx = ((m[0] * p[0] + m[1] * p[1]) + m[2] * p[2]) + m[3];
    const float cy = ((m[4] * p[0] + m[5] * p[1]) + m[6] * p[2]) + m[7];
    const float cw = ((m[12] * p[0] + m[13] * p[1]) + m[14] * p[2]) + m[15];
    nx[k] = cx / cw;
    ny[k] = cy / cw;
  }
  const float w = (1.0f - u) - v;
  *gx = (u * nx[0] + v * nx[1]) + w * nx[2];
  *gy = (u * ny[0] + v * ny[1]) + w * ny[2];
}

// ---- texel pass, in two kernels so that the ray tracing runs on DENSE, coherent warps ------------------------------
// texel_prep_kernel: position + face normal from the UV raster, then per view the angle test and the projected bilinear fetch of
// alpha; a texel that faces view i (the angle test) appends itself to view i's ray list.  Threads map to texels in 16x8 tiles (one 8x4 sub-tile per warp) and every block appends its rays
// with ONE atomicAdd per view, so a list is a sequence of per-tile runs: 32 consecutive entries are rays of one direction
// through neighbouring texels.  ray_kernel walks the lists, one thread per ray, and ORs the view's bit into raw_vis.
// Only the order in which rays run depends on the atomics, the result does not (checked bit for bit against the one-kernel
// form of r01, where 15 of 32 lanes were active on average -- profiles/r01_texel_kernel.metrics.csv -- because lanes not facing
// the view idled through every trace).
__device__ __forceinline__ int tile_texel(int H2, int W2, int block, int thread) {
  if (W2 < 16 || H2 < 8) return block * 128 + thread;
  const int tiles_x = W2 >> 4, ty = block / tiles_x, tx = block - ty * tiles_x;
  const int warp = thread >> 5, lane = thread & 31;
  const int x = (tx << 4) + ((warp & 1) << 3) + (lane & 7), y = (ty << 3) + ((warp >> 1) << 2) + (lane >> 3);
  return y * W2 + x;
}
__global__ void __launch_bounds__(128) texel_prep_kernel(const float4* __restrict__ rast, int H2, int W2,
                                                         const float* __restrict__ vert, const int* __restrict__ tri,
                                                         const Views vw, const float* __restrict__ images, int H, int W,
                                                         float cos_thresh, unsigned char* __restrict__ raw_vis,
                                                         unsigned char* __restrict__ alpha_ok, float* __restrict__ pos_out,
                                                         int* __restrict__ lists, int* __restrict__ counts) {
  __shared__ int wcount[MAXV][4];
  __shared__ int base[MAXV];
  const int T = H2 * W2;
  const int t = tile_texel(H2, W2, blockIdx.x, threadIdx.x);
  unsigned face = 0, aok = 0;
  if (t < T) {
    const float4 r = rast[t];
    const int f = static_cast<int>(r.w) - 1;
    float pos[3] = {0.f, 0.f, 0.f};
    if (f >= 0) {
      const float *p0 = vert + static_cast<size_t>(tri[f * 3]) * 3, *p1 = vert + static_cast<size_t>(tri[f * 3 + 1]) * 3,
                  *p2 = vert + static_cast<size_t>(tri[f * 3 + 2]) * 3;
      const float u = r.x, v = r.y, w = (1.0f - u) - v;
#pragma unroll
      for (int a = 0; a < 3; ++a) pos[a] = (u * p0[a] + v * p1[a]) + w * p2[a];
      const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
      float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
      const float nl = fmaxf(norm3(n[0], n[1], n[2]), 1e-12f);
      n[0] = n[0] / nl; n[1] = n[1] / nl; n[2] = n[2] / nl;
      const float nn = fmaxf(norm3(n[0], n[1], n[2]), 1e-8f);
      for (int i = 0; i < vw.n; ++i) {
        // ray direction (:279-285): the view direction (orthographic) or texel position - camera position (perspective)
        const float dr[3] = {vw.perspective ? pos[0] - vw.dir[i][0] : vw.dir[i][0], vw.perspective ? pos[1] - vw.dir[i][1] : vw.dir[i][1],
                             vw.perspective ? pos[2] - vw.dir[i][2] : vw.dir[i][2]};
        const float dl = fmaxf(norm3(dr[0], dr[1], dr[2]), 1e-12f);
        const float d[3] = {dr[0] / dl, dr[1] / dl, dr[2] / dl};
        const float dn = fmaxf(norm3(d[0], d[1], d[2]), 1e-8f);
        const float cosv = ((d[0] / dn) * (n[0] / nn) + (d[1] / dn) * (n[1] / nn)) + (d[2] / dn) * (n[2] / nn);
        float gx, gy;
        texel_ndc(vw.mat[i], p0, p1, p2, u, v, &gx, &gy);
        float rgba[4];
        bilinear<4>(images + static_cast<size_t>(i) * H * W * 4, H, W, gx, gy, rgba);
        if (rgba[3] > 0.999f) aok |= 1u << i;
        if (cosv < cos_thresh) face |= 1u << i;
      }
    }
    raw_vis[t] = 0;
    alpha_ok[t] = static_cast<unsigned char>(aok);
    pos_out[t * 3] = pos[0]; pos_out[t * 3 + 1] = pos[1]; pos_out[t * 3 + 2] = pos[2];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned ballots[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    ballots[i] = __ballot_sync(0xffffffffu, (face >> i) & 1u);
    if (lane == 0) wcount[i][warp] = __popc(ballots[i]);
  }
  __syncthreads();
  if (threadIdx.x < vw.n) {
    const int i = threadIdx.x;
    const int total = ((wcount[i][0] + wcount[i][1]) + wcount[i][2]) + wcount[i][3];
    base[i] = total ? atomicAdd(counts + i, total) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if ((face >> i) & 1u) {
      int off = base[i] + __popc(ballots[i] & ((1u << lane) - 1u));
      for (int w = 0; w < warp; ++w) off += wcount[i][w];
      lists[static_cast<size_t>(i) * T + off] = t;
    }
  }
}
__global__ void __launch_bounds__(128, UTX_WALK_MIN_BLOCKS) ray_kernel(const int* __restrict__ lists, const int* __restrict__ counts, int T,
                                                  const float4* __restrict__ rast, const float* __restrict__ pos_in,
                                                  const float* __restrict__ vert, const int* __restrict__ tri,
                                                  const float4* __restrict__ wide, const Views vw, unsigned* raw_vis_words) {
  const int view = blockIdx.y;
  // grid-stride over the view's list: the leaf-grid path hands this kernel a short list of unknown length (the rays of crowded
  // cells) and launches a small grid instead of one thread per possible ray
  for (int i = blockIdx.x * blockDim.x + threadIdx.x, n_list = counts[view]; i < n_list; i += gridDim.x * blockDim.x) {
  const int t = lists[static_cast<size_t>(view) * T + i];
  const int f = static_cast<int>(rast[t].w) - 1;
  const float pos[3] = {pos_in[static_cast<size_t>(t) * 3], pos_in[static_cast<size_t>(t) * 3 + 1], pos_in[static_cast<size_t>(t) * 3 + 2]};
  const float k2s3 = 3.4641016151377544f;   // float32(2 * sqrt(3)), renderer_inverse.py:284
  const float* vd = vw.dir[view];
  const float dr[3] = {vw.perspective ? pos[0] - vd[0] : vd[0], vw.perspective ? pos[1] - vd[1] : vd[1], vw.perspective ? pos[2] - vd[2] : vd[2]};
  const float o[3] = {vw.perspective ? vd[0] : pos[0] - k2s3 * dr[0], vw.perspective ? vd[1] : pos[1] - k2s3 * dr[1],
                      vw.perspective ? vd[2] : pos[2] - k2s3 * dr[2]};       // :279-284
  const float dl = fmaxf(norm3(dr[0], dr[1], dr[2]), 1e-12f);
  float d[3] = {dr[0] / dl, dr[1] / dl, dr[2] / dl};                    // F.normalize (:285)
  const float len = norm3(d[0], d[1], d[2]);                            // the tracer normalises again (intersect_test2.slang:283)
  d[0] = d[0] / len; d[1] = d[1] / len; d[2] = d[2] / len;
  const RayHit h = bvh_trace(wide, vert, tri, o, d);
  if (h.any && h.tid == f) atomicOr(raw_vis_words + (t >> 2), (1u << view) << ((t & 3) * 8));
  }
}

// ------------------------------------------------------------------------------------------------ leaf grids for parallel rays
// What the reference's tracer computes, restated (proof below, checked bit for bit at teaser_robot scale by
// tests/test_gpu_teaser.py against the oracle's literal tree walk):
//     scan the LEAVES in DESCENDING order of their position g in the Morton-sorted list; a leaf is "visited" iff its OWN box
//     passes the slab test against the running closest t; a visited leaf whose triangle is hit sets closest = min(closest, t)
//     and becomes the reported triangle (last accepted leaf wins).
//  * order: an internal node of the Karras tree covers a contiguous range [first, last] of the sorted leaves, its left child
//    [first, split], its right child [split + 1, last]; the reference pushes left, right and pops right first
//    (intersect_test2.slang:104-121), so leaves are reached in strictly descending g;
//  * ancestors do not matter: boxes nest (the refit takes exact min / max), the slab arithmetic is monotone in the box
//    coordinates, so te_ancestor <= te_leaf and tx_ancestor >= tx_leaf; an ancestor is popped BEFORE the leaf, when closest is
//    at least as large; hence "the leaf's own test passes" implies "every ancestor's test passed", and the converse is the
//    leaf's own test.  (The 64-entry stack of the reference never overflows: its depth is at most the tree depth + 1.)
// So the hierarchy is only an index for finding the leaves whose boxes a ray can pass -- and for the bake's rays, PARALLEL to a
// coordinate axis (the six box views the reference bake is hard-wired to, renderer_inverse.py:171,256), a much better index
// exists: a 2-D grid over the two other axes.  Per view, every leaf is entered into the cells its box rectangle (grown by
// eps: the 1e-6 the reference substitutes for a zero direction component tilts the slab test by <= 1e-6 * t) overlaps, each
// cell's list is sorted by descending g, and a ray walks ONE cell's list: ~15 box tests instead of ~300 tree nodes, the same
// exact slab and Moller-Trumbore arithmetic.  teaser_robot: ray stage 4.28 ms (tree walk) -> 2.2 ms (grid build 0.8 + walk 1.4),
// bake 9.49 -> 7.60 ms; two-sphere mesh 8.55 -> 6.28 ms; identical output checksums (profiles/r02_summary.md).
struct GridCfg {
  int n, G, F;
  int axis[MAXV];
};
struct GridDev {   // written by grid_setup_kernel from the root box
  float lo[3], scale[3], eps;
};
__device__ __forceinline__ int grid_cell(float x, float lo, float scale, int G) {   // monotone non-decreasing in x
  const float f = floorf((x - lo) * scale);
  return f < 0.f ? 0 : (f >= static_cast<float>(G) ? G - 1 : static_cast<int>(f));
}
__global__ void grid_setup_kernel(const float4* __restrict__ nodes4, int G, GridDev* gd) {
  const float4 q0 = nodes4[0], q1 = nodes4[1];                       // root = node 0
  const float lo[3] = {q0.x, q0.y, q0.z}, hi[3] = {q0.w, q1.x, q1.y};
  float ext = 0.f;
  for (int a = 0; a < 3; ++a) ext = fmaxf(ext, hi[a] - lo[a]);
  // rays start 2 sqrt(3) before the texel (renderer_inverse.py:284): every t of interest is below ext + 8
  gd->eps = 4e-6f * (ext + 8.0f);
  for (int a = 0; a < 3; ++a) {
    const float pad = fmaxf((hi[a] - lo[a]) * 1e-3f, 8.0f * gd->eps);
    gd->lo[a] = lo[a] - pad;
    gd->scale[a] = static_cast<float>(G) / fmaxf((hi[a] - lo[a]) + 2.0f * pad, 1e-20f);
  }
}
// mode 0: count the leaves per cell; mode 1: write them (slot = start + cursor++)
template <int kMode>
__global__ void __launch_bounds__(256) grid_scatter_kernel(const float4* __restrict__ nodes4, const GridCfg gc, const GridDev* __restrict__ gdp,
                                                           int* __restrict__ count, const int* __restrict__ start, int* __restrict__ list,
                                                           long long cap) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(gc.n) * gc.F) return;
  // threads walk the leaves from the LAST one down: the atomics then fill every cell in roughly descending g, which is the
  // order grid_sort_kernel wants -- its insertion sort is linear on nearly sorted lists and quadratic on reversed ones
  // (filled in ascending order it took 2.7 ms on teaser_robot, more than the tree walk it replaces saved)
  const int view = static_cast<int>(idx / gc.F), g = gc.F - 1 - static_cast<int>(idx - static_cast<long long>(view) * gc.F);
  const GridDev gd = *gdp;
  const float4* nd = nodes4 + 3 * static_cast<size_t>(gc.F - 1 + g);
  const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1);
  const float lo[3] = {q0.x, q0.y, q0.z}, hi[3] = {q0.w, q1.x, q1.y};
  const int a = gc.axis[view], b = (a + 1) % 3, c = (a + 2) % 3, G = gc.G;
  const int b0 = grid_cell(lo[b] - gd.eps, gd.lo[b], gd.scale[b], G), b1 = grid_cell(hi[b] + gd.eps, gd.lo[b], gd.scale[b], G);
  const int c0 = grid_cell(lo[c] - gd.eps, gd.lo[c], gd.scale[c], G), c1 = grid_cell(hi[c] + gd.eps, gd.lo[c], gd.scale[c], G);
  for (int cc = c0; cc <= c1; ++cc)
    for (int cb = b0; cb <= b1; ++cb) {
      const int cell = (view * G + cc) * G + cb;
      const int k = atomicAdd(count + cell, 1);
      if (kMode == 1) {
        const long long slot = static_cast<long long>(start[cell]) + k;
        if (slot < cap) list[slot] = g;
      }
    }
}
// A cell whose list is longer than this is left to the tree walk: where the surface is tangent to the view direction (the limb
// of a sphere) thousands of leaves pile up in one cell; one thread sorting such a list set the time of the whole sort kernel
// (1.8 ms on teaser_robot, 10 ms on the two-sphere mesh), and a ray scanning it gains nothing over the hierarchy.
constexpr int GRID_LMAX = 256;
// One WARP per cell: the list as the atomics filled it (arbitrary order) -> descending g, by rank counting: an element's place
// is the number of larger elements of its cell (the g of one cell are distinct).  Every lane reads the same list entries
// (broadcast loads), there is no serial chain -- a one-thread-per-cell insertion sort in global memory took 1.9 ms on
// teaser_robot because a single unsorted list of 256 entries is ~16 k dependent round trips.
__global__ void __launch_bounds__(256) grid_sort_kernel(const int* __restrict__ start, int ncells, const int* __restrict__ in,
                                                        int* __restrict__ out) {
  const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (cell >= ncells) return;
  const int k0 = start[cell], n = start[cell + 1] - k0;
  if (n == 0 || n > GRID_LMAX) return;
  for (int e = lane; e < n; e += 32) {
    const int g = in[k0 + e];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += in[k0 + j] > g;
    out[k0 + rank] = g;
  }
}
__global__ void __launch_bounds__(128) ray_grid_kernel(const int* __restrict__ lists, const int* __restrict__ counts, int T,
                                                       const float4* __restrict__ rast, const float* __restrict__ pos_in,
                                                       const float* __restrict__ vert, const int* __restrict__ tri,
                                                       const float4* __restrict__ nodes4, const GridCfg gc, const GridDev* __restrict__ gdp,
                                                       const int* __restrict__ start, const int* __restrict__ glist, const Views vw,
                                                       unsigned* raw_vis_words, int* __restrict__ lists2, int* __restrict__ counts2) {
  const int view = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[view]) return;
  const int t = lists[static_cast<size_t>(view) * T + i];
  const int f = static_cast<int>(rast[t].w) - 1;
  const float pos[3] = {pos_in[static_cast<size_t>(t) * 3], pos_in[static_cast<size_t>(t) * 3 + 1], pos_in[static_cast<size_t>(t) * 3 + 2]};
  const float k2s3 = 3.4641016151377544f;   // float32(2 * sqrt(3)), renderer_inverse.py:284
  const float* dr = vw.dir[view];
  const float o[3] = {pos[0] - k2s3 * dr[0], pos[1] - k2s3 * dr[1], pos[2] - k2s3 * dr[2]};      // :279-284 (orthographic)
  const float dl = fmaxf(norm3(dr[0], dr[1], dr[2]), 1e-12f);
  float d[3] = {dr[0] / dl, dr[1] / dl, dr[2] / dl};                    // F.normalize (:285)
  const float len = norm3(d[0], d[1], d[2]);                            // the tracer normalises again (intersect_test2.slang:283)
  d[0] = d[0] / len; d[1] = d[1] / len; d[2] = d[2] / len;
  float inv[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float di = d[k];
    if (di == 0.0f) di = 0.000001f;      // intersect_test2.slang:18-21
    inv[k] = 1.0f / di;
  }
  const GridDev gd = *gdp;
  const int a = gc.axis[view], b = (a + 1) % 3, c = (a + 2) % 3, G = gc.G;
  const int cell = (view * G + grid_cell(o[c], gd.lo[c], gd.scale[c], G)) * G + grid_cell(o[b], gd.lo[b], gd.scale[b], G);
  const int k0 = start[cell], k1 = start[cell + 1];
  if (k1 - k0 > GRID_LMAX) {           // crowded cell (unsorted): this ray goes to the tree walk's list
    lists2[static_cast<size_t>(view) * T + atomicAdd(counts2 + view, 1)] = t;
    return;
  }
  float closest = 1e9f;
  int htid = -1;
  // one candidate: its own box against the running closest t, then (rarely) the triangle; strictly in list order
  auto visit = [&](const float4* nd, const float4 q0, const float4 q1) {
    const float bb[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
    float te, tx;
    slab_params(o, inv, bb, te, tx);
    if (!slab_pass(te, tx, closest)) return;
    const int p = __float_as_int(__ldg(nd + 2).x);
    float th, u, v;
    if (!triangle_hit_dev(vert, tri, p, o, d, th, u, v)) return;
    closest = th < closest ? th : closest;
    htid = p;                                                            // last accepted leaf wins (reference quirk)
  };
  int k = k0;
  for (; k + 4 <= k1; k += 4) {                                           // four candidates' boxes in flight at once
    const float4* nd[4];
    float4 q0[4], q1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) nd[j] = nodes4 + 3 * static_cast<size_t>(gc.F - 1 + __ldg(glist + k + j));
#pragma unroll
    for (int j = 0; j < 4; ++j) { q0[j] = __ldg(nd[j]); q1[j] = __ldg(nd[j] + 1); }
#pragma unroll
    for (int j = 0; j < 4; ++j) visit(nd[j], q0[j], q1[j]);
  }
  for (; k < k1; ++k) {
    const float4* nd = nodes4 + 3 * static_cast<size_t>(gc.F - 1 + __ldg(glist + k));
    visit(nd, __ldg(nd), __ldg(nd + 1));
  }
  if (htid >= 0 && htid == f) atomicOr(raw_vis_words + (t >> 2), (1u << view) << ((t & 3) * 8));
}

// k = 3: conv >= 3  <=>  at least one of the 8 ring texels set (9 r - c >= 3, c <= 1)
__global__ void __launch_bounds__(256) repair3_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                      int H, int W) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W) return;
  const int y = t / W, x = t % W;
  unsigned acc = in[t];
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      if (dy == 0 && dx == 0) continue;
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc |= in[yy * W + xx];
    }
  out[t] = static_cast<unsigned char>(acc);
}
// k = 5: 25 r - c >= 135 with c <= 9  <=>  r >= 6 of the 16 ring texels
__global__ void __launch_bounds__(256) repair5_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                      int H, int W, int n_views) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W) return;
  const int y = t / W, x = t % W;
  int cnt[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) cnt[i] = 0;
  for (int dy = -2; dy <= 2; ++dy)
    for (int dx = -2; dx <= 2; ++dx) {
      if (abs(dy) != 2 && abs(dx) != 2) continue;
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const unsigned b = in[yy * W + xx];
#pragma unroll
      for (int i = 0; i < MAXV; ++i) cnt[i] += (b >> i) & 1u;
    }
  unsigned acc = in[t];
  for (int i = 0; i < n_views; ++i)
    if (cnt[i] >= 6) acc |= 1u << i;
  out[t] = static_cast<unsigned char>(acc);
}

__global__ void __launch_bounds__(128) compose_kernel(const float4* __restrict__ rast, int T, const float* __restrict__ vert,
                                                      const int* __restrict__ tri, const Views vw,
                                                      const float* __restrict__ images, int H, int W,
                                                      const unsigned char* __restrict__ vis_rep,
                                                      const unsigned char* __restrict__ alpha_ok,
                                                      unsigned char* __restrict__ mask2d, unsigned char* __restrict__ mask_vis,
                                                      signed char* __restrict__ owner, float* __restrict__ color) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float4 r = rast[t];
  const int f = static_cast<int>(r.w) - 1;
  const unsigned bits = f >= 0 ? (vis_rep[t] & alpha_ok[t]) : 0u;
  mask2d[t] = f >= 0;
  for (int i = 0; i < vw.n; ++i) mask_vis[static_cast<size_t>(i) * T + t] = (bits >> i) & 1u;
  int own = -1;
  for (int k = 0; k < vw.n; ++k) {
    const int i = vw.priority[k];
    if ((bits >> i) & 1u) { own = i; break; }
  }
  owner[t] = static_cast<signed char>(own);
  float c[3] = {0.f, 0.f, 0.f};
  if (own >= 0) {
    const float *p0 = vert + static_cast<size_t>(tri[f * 3]) * 3, *p1 = vert + static_cast<size_t>(tri[f * 3 + 1]) * 3,
                *p2 = vert + static_cast<size_t>(tri[f * 3 + 2]) * 3;
    float gx, gy, rgba[4];
    texel_ndc(vw.mat[own], p0, p1, p2, r.x, r.y, &gx, &gy);
    bilinear<4>(images + static_cast<size_t>(own) * H * W * 4, H, W, gx, gy, rgba);
    c[0] = rgba[0]; c[1] = rgba[1]; c[2] = rgba[2];
  }
  color[t * 3] = c[0]; color[t * 3 + 1] = c[1]; color[t * 3 + 2] = c[2];
}

__global__ void __launch_bounds__(256) seam0_kernel(const signed char* __restrict__ owner, unsigned char* __restrict__ b0,
                                                    int H, int W) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W) return;
  const int y = t / W, x = t % W;
  const int o = owner[t];
  int diff = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) diff |= owner[yy * W + xx] != o;
    }
  b0[t] = static_cast<unsigned char>(diff);
}
__global__ void __launch_bounds__(256) seam1_kernel(const unsigned char* __restrict__ b0,
                                                    const unsigned char* __restrict__ mask2d,
                                                    unsigned char* __restrict__ seam, int H, int W) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W) return;
  const int y = t / W, x = t % W;
  int any = 0, all = 1;
  for (int dy = -3; dy <= 3; ++dy)
    for (int dx = -3; dx <= 3; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      if (abs(dy) <= 1 && abs(dx) <= 1) any |= b0[yy * W + xx];
      all &= mask2d[yy * W + xx];
    }
  seam[t] = static_cast<unsigned char>(any && all);
}

// ------------------------------------------------------------------------------------------------ exact 1-NN fill
// Visible texels are compacted in index order (flags -> exclusive scan -> scatter), an LBVH is built over their 3-D
// positions (same builder as the triangle tree, bake_bvh.cu) and every covered-but-invisible texel walks it for its exact
// nearest neighbour.  Cost is logarithmic in the number of visible texels and independent of how far the hidden region is
// from the nearest visible one (the first version used a uniform grid whose ring search was 88 % of the bake).
// flags the owned texels (the source points) and appends the covered-but-unowned ones (the queries) to `qlist`, one atomicAdd
// per block: the query kernels then run on dense warps of neighbouring texels instead of the ~10 % of a full-atlas launch
__global__ void __launch_bounds__(256) nn_flag_kernel(const unsigned char* __restrict__ mask2d,
                                                      const signed char* __restrict__ owner, int T, int* __restrict__ flags,
                                                      int* __restrict__ nn_index, int* __restrict__ qlist,
                                                      int* __restrict__ qcount, int H2, int W2) {
  __shared__ int wcount[8];
  __shared__ int base;
  // 16x8 texel tiles, one 8x4 sub-tile per warp (as in texel_prep_kernel): 32 consecutive queries are spatial neighbours
  const int t = tile_texel(H2, W2, blockIdx.x * 2 + (threadIdx.x >> 7), threadIdx.x & 127);
  bool query = false;
  if (t < T) {
    const int o = owner[t];
    flags[t] = o >= 0;
    if (nn_index) nn_index[t] = -1;
    query = mask2d[t] && o < 0;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned b = __ballot_sync(0xffffffffu, query);
  if (lane == 0) wcount[warp] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < 8; ++w) total += wcount[w];
    base = total ? atomicAdd(qcount, total) : 0;
  }
  __syncthreads();
  if (query) {
    int off = base + __popc(b & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) off += wcount[w];
    qlist[off] = t;
  }
}
__global__ void __launch_bounds__(256) nn_compact_kernel(const signed char* __restrict__ owner, const int* __restrict__ offs,
                                                         const float* __restrict__ pos, int T, int* __restrict__ ids,
                                                         float* __restrict__ pts) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T || owner[t] < 0) return;
  const int k = offs[t];
  ids[k] = t;
  pts[static_cast<size_t>(k) * 3] = pos[static_cast<size_t>(t) * 3];
  pts[static_cast<size_t>(k) * 3 + 1] = pos[static_cast<size_t>(t) * 3 + 1];
  pts[static_cast<size_t>(k) * 3 + 2] = pos[static_cast<size_t>(t) * 3 + 2];
}
__device__ __forceinline__ void nn_query_run(const int* __restrict__ qlist, int n_q, const float* __restrict__ pos, const PointTree& pt,
                                             const float* color_in, float* color_out, int* __restrict__ nn_index, int W2, int i) {
  if (i >= n_q) return;
  const int t = qlist[i];
  const float q[3] = {pos[static_cast<size_t>(t) * 3], pos[static_cast<size_t>(t) * 3 + 1], pos[static_cast<size_t>(t) * 3 + 2]};
  // When the warp's queries are neighbours (texels within 16 rows / columns of the first live lane's) its lanes order their
  // walks by ONE common point (that lane's query) and so stay converged: 2.86 -> 2.52 ms on the bench bake.  Applied to every
  // warp it doubles the time (5.4 ms): lanes far from the common point descend into the wrong subtree first.
  const unsigned live = __activemask();
  const int lead = __ffs(live) - 1;
  const float oq[3] = {__shfl_sync(live, q[0], lead), __shfl_sync(live, q[1], lead), __shfl_sync(live, q[2], lead)};
  const int tl = __shfl_sync(live, t, lead);
  const int dy = t / W2 - tl / W2, dx = t % W2 - tl % W2;
  const bool near_lead = dy > -16 && dy < 16 && dx > -16 && dx < 16;
  const bool common = __all_sync(live, near_lead);
  const int best = nn_trace(pt, q, nullptr, common ? oq : nullptr);
  if (nn_index) nn_index[t] = best;
  if (best >= 0) {
    color_out[t * 3] = color_in[static_cast<size_t>(best) * 3];
    color_out[t * 3 + 1] = color_in[static_cast<size_t>(best) * 3 + 1];
    color_out[t * 3 + 2] = color_in[static_cast<size_t>(best) * 3 + 2];
  }
}
// Warp-cooperative walk: the 32 lanes of a warp hold 32 NEIGHBOURING queries (one 8x4 texel patch of the query list) and
// descend the point tree TOGETHER -- one warp-uniform stack, every record / leaf load a broadcast, every lane scoring every
// visited leaf against its own query.
// Why this is exact: nearest-neighbour search has no order dependence -- scoring a point a lane "did not need" can only
// confirm its current best, so a lane may visit any superset of the nodes its own pruned walk would visit.  The only pruning
// left is warp-wide: an entry is skipped when EVERY lane's (deflated) box bound exceeds that lane's current best -- at push
// time lane by lane (ballot), at pop time through the conservative scalars  min over lanes of the bound  vs  max over lanes
// of the best.  Ties go to the lowest id as in nn_trace.
// Why it is fast: the per-lane walks of r01 (nn_query_run below, kept for warps whose queries are NOT neighbours) ran at
// 12.7 of 32 lanes with every 16-byte record load touching its own line and 4.3 long-scoreboard stalls per issue
// (profiles/r01_bake_ray_nn.metrics.csv, r01_nn_query_persist.ncu-rep): neighbouring queries need almost the same nodes, but
// lanes that prune differently fall out of step and never re-converge.  Here the union of the needed nodes is walked once with
// all lanes busy.  (A per-lane refill variant -- a lane takes a new query as soon as its walk ends -- was measured SLOWER in
// r01, 2.9 vs 2.5 ms: lanes then hold unrelated queries; it is in the history of this file.)
__device__ __forceinline__ void nn_query_run_coop(const float q[3], const PointTree& pt, int* s_ref, float* s_bound,
                                                  float& best_out, int& best_id_out) {
  constexpr unsigned FULL = 0xffffffffu;
  float best = INFINITY;
  int best_id = -1;
  float wbest = INFINITY;          // max over lanes of `best` (warp-uniform)
  int count = 0;
  bool have = true;
  int ref = pt.n_c <= 1 ? ~0 : 0;  // a tree of one cluster is its only leaf
  float bound = 0.f;
  for (;;) {
    if (!have) {
      if (count == 0) break;
      --count;
      ref = s_ref[count];
      bound = s_bound[count];
    }
    have = false;
    if (bound > wbest) continue;                                        // min bound over lanes > max best over lanes: nobody needs it
    if (ref < 0) {
      const int c = ~ref;
      const int j0 = c * PT_CLUSTER, j1 = min(j0 + PT_CLUSTER, pt.n);
      auto score = [&](const float4 p) {
        const float dx = p.x - q[0], dy = p.y - q[1], dz = p.z - q[2];
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        const int id = __float_as_int(p.w);
        if (d2 < best || (d2 == best && id < best_id) || best_id < 0) { best = d2; best_id = id; }
      };
      if (j1 - j0 == PT_CLUSTER) {                                         // full leaf: all its loads in flight at once
        float4 pbuf[PT_CLUSTER];
#pragma unroll
        for (int j = 0; j < PT_CLUSTER; ++j) pbuf[j] = __ldg(pt.spts + j0 + j);   // broadcast loads
#pragma unroll
        for (int j = 0; j < PT_CLUSTER; ++j) score(pbuf[j]);
      } else {
        for (int j = j0; j < j1; ++j) score(__ldg(pt.spts + j));
      }
      wbest = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(best)));   // non-negative floats order like their bits
      continue;
    }
    const WideRec r = wide_load(pt.wide, ref);                           // broadcast load
    // per entry: the smallest bound over the lanes, or INFINITY when the slot is unused or no lane needs it (all warp-uniform)
    float mb[4];
    int rf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool used = r.ref[k] != WIDE_EMPTY;                          // warp-uniform
      const float bd = used ? box_dist2(r.bb + 6 * k, q) * 0.999999f : INFINITY;
      const bool any = __ballot_sync(FULL, bd <= best) != 0u;
      const float m = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(bd)));
      mb[k] = (used && any) ? m : INFINITY;
      rf[k] = r.ref[k];
    }
    // 4-element sorting network, DESCENDING by bound (registers only): the nearest entry ends up in slot 3 and is visited next,
    // the others are pushed farthest first (popped nearest first); INFINITY entries sort to the front and are skipped
#define UTX_CSWAP(a, b)                                                            \
    if (mb[a] < mb[b]) { const float tf = mb[a]; mb[a] = mb[b]; mb[b] = tf; const int ti = rf[a]; rf[a] = rf[b]; rf[b] = ti; }
    UTX_CSWAP(0, 1) UTX_CSWAP(2, 3) UTX_CSWAP(0, 2) UTX_CSWAP(1, 3) UTX_CSWAP(1, 2)
#undef UTX_CSWAP
    if (mb[3] == INFINITY) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (mb[j] < INFINITY && count < 64) { s_ref[count] = rf[j]; s_bound[count] = mb[j]; ++count; }
    have = true;
    ref = rf[3];
    bound = mb[3];
  }
  best_out = best;
  best_id_out = best_id;
}

#ifdef UTX_NN_DEBUG
// diagnosis only (scripts/nn_visits.py): points scored per query instead of the neighbour index
__global__ void __launch_bounds__(128) nn_count_kernel(const int* __restrict__ qlist, int n_q, const float* __restrict__ pos,
                                                       const PointTree pt, int* __restrict__ count_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q) return;
  const int t = qlist[i];
  const float q[3] = {pos[static_cast<size_t>(t) * 3], pos[static_cast<size_t>(t) * 3 + 1], pos[static_cast<size_t>(t) * 3 + 2]};
  float best = INFINITY;
  int best_id = -1, scored = 0;
  point_tree_walk(pt, q, [&]() { return best; },
                  [&](float d2, int id) {
                    ++scored;
                    if (d2 < best || (d2 == best && id < best_id) || best_id < 0) { best = d2; best_id = id; }
                  });
  count_out[t] = scored;
}
#endif
// Persistent warps: every warp takes the next run of 32 consecutive list entries from a global counter until the list is
// exhausted.  Walk lengths are very uneven (a query deep inside an occluded region scans the whole rim of its empty ball), and with
// one fixed block of queries per CTA the kernel ran at 29 % of its resident warps (profiles/r01_bake_ray_nn.metrics.csv): a CTA's
// slot was held until its slowest warp finished, and the last wave left most SMs idle.  (Measured: no change, 2.49 ms -- the
// list is only 2.3 runs per resident warp, so the tail of one run remains; kept because it is never worse.)
template <bool kCoop>
__global__ void __launch_bounds__(128, UTX_WALK_MIN_BLOCKS) nn_query_kernel(const int* __restrict__ qlist, int n_q, const float* __restrict__ pos,
                                                       const PointTree pt, const float* color_in, float* color_out,
                                                       int* __restrict__ nn_index, int W2, int* __restrict__ next_run, int always_coop) {
  __shared__ int s_ref[4][64];
  __shared__ float s_bound[4][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (;;) {
    int run = 0;
    if (lane == 0) run = atomicAdd(next_run, 1);
    run = __shfl_sync(0xffffffffu, run, 0);
    if (run * 32 >= n_q) return;
    if (!kCoop) {
      nn_query_run(qlist, n_q, pos, pt, color_in, color_out, nn_index, W2, run * 32 + lane);
      continue;
    }
    // `always_coop` (default): every run walks cooperatively -- a run that straddles charts holds a few groups of neighbours and
    // pays the union of their walks, still far less than 32 divergent ones.  Otherwise (A/B) only runs whose texels lie within
    // 16 rows / columns of the first one do, the rest fall back to the per-lane walks
    const int i = run * 32 + lane;
    const bool valid = i < n_q;
    const int t = qlist[valid ? i : run * 32];                          // tail lanes duplicate the first query (harmless)
    const int tl = __shfl_sync(0xffffffffu, t, 0);
    const int dy = t / W2 - tl / W2, dx = t % W2 - tl % W2;
    const bool common = __all_sync(0xffffffffu, dy > -16 && dy < 16 && dx > -16 && dx < 16) || always_coop;
    if (!common) {
      if (valid) nn_query_run(qlist, n_q, pos, pt, color_in, color_out, nn_index, W2, i);
      __syncwarp();
      continue;
    }
    const float q[3] = {pos[static_cast<size_t>(t) * 3], pos[static_cast<size_t>(t) * 3 + 1], pos[static_cast<size_t>(t) * 3 + 2]};
    float best;
    int best_id;
    nn_query_run_coop(q, pt, s_ref[warp], s_bound[warp], best, best_id);
    __syncwarp();
    if (valid) {
      if (nn_index) nn_index[t] = best_id;
      if (best_id >= 0) {
        color_out[t * 3] = color_in[static_cast<size_t>(best_id) * 3];
        color_out[t * 3 + 1] = color_in[static_cast<size_t>(best_id) * 3 + 1];
        color_out[t * 3 + 2] = color_in[static_cast<size_t>(best_id) * 3 + 2];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ lens blur on seams
__global__ void __launch_bounds__(256) lens_blur_kernel(const float* __restrict__ color_in, const unsigned char* __restrict__ seam,
                                                        const float* __restrict__ k2d /*[49]*/, float gamma, int H, int W,
                                                        float* __restrict__ color_out, int reflect) {
  __shared__ float ks[49];
  if (threadIdx.x < 49) ks[threadIdx.x] = k2d[threadIdx.x];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W) return;
  float c[3] = {color_in[t * 3], color_in[t * 3 + 1], color_in[t * 3 + 2]};
  if (seam[t]) {
    const int y = t / W, x = t % W;
    float acc[3] = {0.f, 0.f, 0.f};
    // border: zeros (the lens blur's conv2d padding) or mirrored without the edge texel (torchvision gaussian_blur pads 'reflect')
    for (int dy = -3; dy <= 3; ++dy) {
      int yy = y + dy;
      if (reflect) yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
      if (yy < 0 || yy >= H) continue;
      for (int dx = -3; dx <= 3; ++dx) {
        int xx = x + dx;
        if (reflect) xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
        if (xx < 0 || xx >= W) continue;
        const float kw = ks[(dy + 3) * 7 + (dx + 3)];
        const float* p = color_in + (static_cast<size_t>(yy) * W + xx) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          // the lens blur's exposure gamma is 5 (lens_blur.py:162): x^5 as three multiplications instead of 147 powf per seam texel
          const float q = p[a] * p[a];
          acc[a] = acc[a] + kw * (gamma == 1.0f ? p[a] : (gamma == 5.0f ? q * q * p[a] : powf(p[a], gamma)));
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
      c[a] = gamma == 1.0f ? acc[a] : fminf(fmaxf(powf(fmaxf(acc[a], 0.f), 1.0f / gamma), 0.f), 1.f);
  }
  color_out[t * 3] = c[0]; color_out[t * 3 + 1] = c[1]; color_out[t * 3 + 2] = c[2];
}

// ------------------------------------------------------------------------------------------------ pull-push
__global__ void __launch_bounds__(256) pp_init_kernel(const float* __restrict__ color, const unsigned char* __restrict__ mask,
                                                      int T, float* __restrict__ c0, unsigned char* __restrict__ m0) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const bool m = mask[t] != 0;
  m0[t] = m;
#pragma unroll
  for (int a = 0; a < 3; ++a) c0[t * 3 + a] = m ? color[t * 3 + a] : 0.f;
}
__global__ void __launch_bounds__(256) pp_down_kernel(const float* __restrict__ c, const unsigned char* __restrict__ m, int H,
                                                      int W, float* __restrict__ cd, unsigned char* __restrict__ md) {
  const int Hd = H / 2, Wd = W / 2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Hd * Wd) return;
  const int y = t / Wd, x = t % Wd;
  const int i00 = (2 * y) * W + 2 * x, i01 = i00 + 1, i10 = i00 + W, i11 = i10 + 1;
  const float a = (((m[i00] ? 1.f : 0.f) + (m[i01] ? 1.f : 0.f)) + (m[i10] ? 1.f : 0.f) + (m[i11] ? 1.f : 0.f)) * 0.25f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = (((c[i00 * 3 + k] + c[i01 * 3 + k]) + c[i10 * 3 + k]) + c[i11 * 3 + k]) * 0.25f;
    if (a > 0.f && a < 1.f) v = v / a;
    cd[t * 3 + k] = v;
  }
  md[t] = a > 0.f;
}
__global__ void __launch_bounds__(256) pp_up_kernel(float* __restrict__ c, const unsigned char* __restrict__ m, int H, int W,
                                                    const float* __restrict__ cd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W || m[t]) return;
  const int Hd = H / 2, Wd = W / 2;
  const int y = t / W, x = t % W;
  const int yn = y >> 1, xn = x >> 1;
  const int yf = min(max(yn + ((y & 1) ? 1 : -1), 0), Hd - 1), xf = min(max(xn + ((x & 1) ? 1 : -1), 0), Wd - 1);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float v = ((0.5625f * cd[(yn * Wd + xn) * 3 + k] + 0.1875f * cd[(yn * Wd + xf) * 3 + k]) +
                     0.1875f * cd[(yf * Wd + xn) * 3 + k]) + 0.0625f * cd[(yf * Wd + xf) * 3 + k];
    c[t * 3 + k] = v;
  }
}

__global__ void __launch_bounds__(256) transform_points_kernel(const float* __restrict__ vert, int V,
                                                               const float* __restrict__ mats, int n,
                                                               float4* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n) * V) return;
  const int b = static_cast<int>(i / V), v = static_cast<int>(i % V);
  const float* m = mats + b * 16;
  const float* p = vert + static_cast<size_t>(v) * 3;
  float o[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) o[r] = ((m[r * 4] * p[0] + m[r * 4 + 1] * p[1]) + m[r * 4 + 2] * p[2]) + m[r * 4 + 3];
  out[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------ k-NN colour fills
// (bake_mv_to_uv_kdtree, renderer_inverse.py:367-433).  `want`: texels with owner == want are recoloured (want >= 0), or
// every covered texel (want == -2), or the covered-but-unowned ones (want == -1).  The colour is the mean of the k nearest
// source points' colours, summed in ascending (distance, id) order: `colors[index, :].mean(dim=-2)` (:421, :431).
__global__ void __launch_bounds__(128) knn_mean_kernel(const unsigned char* __restrict__ mask2d,
                                                       const signed char* __restrict__ owner, int want,
                                                       const float* __restrict__ pos, int T, const PointTree pt, int k,
                                                       const float* src_col, float* dst_col, int* __restrict__ nn_index,
                                                       const int* __restrict__ list = nullptr, int n_list = 0) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (list) {                                   // the fill stage's compacted query list (its flag kernel reset nn_index)
    if (t >= n_list) return;
    t = list[t];
  } else {
    if (t >= T) return;
    if (nn_index) nn_index[t] = -1;
    if (!mask2d[t] || pt.n == 0) return;
    const int o = owner[t];
    if (want >= 0 ? o != want : (want == -1 && o >= 0)) return;
  }
  const float q[3] = {pos[static_cast<size_t>(t) * 3], pos[static_cast<size_t>(t) * 3 + 1], pos[static_cast<size_t>(t) * 3 + 2]};
  float bd[KNN_MAX];
  int bi[KNN_MAX];
  const int kk = k < pt.n ? k : pt.n;
  knn_trace(pt, q, kk, bd, bi);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < kk; ++j) {
    const float* c = src_col + static_cast<size_t>(bi[j]) * 3;
    acc[0] = acc[0] + c[0]; acc[1] = acc[1] + c[1]; acc[2] = acc[2] + c[2];
  }
  const float kf = static_cast<float>(kk);
  dst_col[t * 3] = acc[0] / kf; dst_col[t * 3 + 1] = acc[1] / kf; dst_col[t * 3 + 2] = acc[2] / kf;
  if (nn_index) nn_index[t] = bi[0];
}

// per-view pixel clouds of mv_to_pcd (:227-231): a pixel belongs to its view's cloud when its visible alpha is set
__global__ void __launch_bounds__(256) pix_flag_kernel(const float4* __restrict__ rgba, long long N, int* __restrict__ flags) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < N) flags[i] = rgba[i].w > 0.5f;
}
__global__ void __launch_bounds__(256) pix_compact_kernel(const float4* __restrict__ rgba, const float* __restrict__ pix_pos,
                                                          const int* __restrict__ flags, const int* __restrict__ offs,
                                                          long long N, float* __restrict__ pts, float* __restrict__ cols) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N || !flags[i]) return;
  const size_t k = static_cast<size_t>(offs[i]) * 3;
  const float4 c = rgba[i];
  pts[k] = pix_pos[i * 3]; pts[k + 1] = pix_pos[i * 3 + 1]; pts[k + 2] = pix_pos[i * 3 + 2];
  cols[k] = c.x; cols[k + 1] = c.y; cols[k + 2] = c.z;
}
inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

// fixed carve-up of the caller's workspace for an H2 x W2 atlas; the staged entry points find each other's results here
struct BakeWs {
  unsigned char *raw, *aok, *rep3, *rep5, *b0, *seam;
  float *pos, *col_a, *col_b;
  signed char* owner;
  int *ids, *flags, *offs;
  float* pts;
  void* scan_tmp;
  size_t scan_bytes;
  int* counters;   // [0..MAXV) ray-list lengths, [MAXV] query-list length
  uint8_t* rest;   // nearest-neighbour tree + builder scratch during the fill, the pull-push pyramid afterwards
};
BakeWs carve(void* workspace, int T) {
  BakeWs w{};
  uint8_t* p = static_cast<uint8_t*>(workspace);
  auto take = [&](size_t b) { uint8_t* q = p; p += al(b); return q; };
  w.raw = take(T); w.aok = take(T); w.rep3 = take(T); w.rep5 = take(T); w.b0 = take(T); w.seam = take(T);
  w.pos = reinterpret_cast<float*>(take(static_cast<size_t>(T) * 12));
  w.col_a = reinterpret_cast<float*>(take(static_cast<size_t>(T) * 12));
  w.col_b = reinterpret_cast<float*>(take(static_cast<size_t>(T) * 12));
  w.owner = reinterpret_cast<signed char*>(take(static_cast<size_t>(T) * 4));
  w.ids = reinterpret_cast<int*>(take(static_cast<size_t>(T) * 4));
  w.flags = reinterpret_cast<int*>(take(static_cast<size_t>(T) * 4));
  w.offs = reinterpret_cast<int*>(take(static_cast<size_t>(T) * 4 + 4));
  w.pts = reinterpret_cast<float*>(take(static_cast<size_t>(T) * 12));
  cub::DeviceScan::ExclusiveSum(nullptr, w.scan_bytes, w.flags, w.offs, T);
  w.scan_tmp = take(w.scan_bytes);
  w.counters = reinterpret_cast<int*>(take(256));
  w.rest = p;
  return w;
}

int check_atlas(int H2, int W2, size_t ws_bytes, const char* who) {
  UTX_CHECK(H2 >= 8 && W2 >= 8 && (H2 & (H2 - 1)) == 0 && (W2 & (W2 - 1)) == 0, "uv_bake: atlas must be a power of two >= 8");
  UTX_CHECK(ws_bytes >= uv_bake_workspace_bytes(H2, W2), "uv_bake: workspace too small");
  (void)who;
  return 0;
}

// flags -> exclusive scan -> total on the host (the tree builder needs the point count: one sync)
int count_flags(const int* flags, int* offs, long long N, void* scan_tmp, size_t scan_bytes, int* total, cudaStream_t stream,
                const int* extra_dev = nullptr, int* extra = nullptr) {
  UTX_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flags, offs, static_cast<int>(N), stream));
  int last_off = 0, last_flag = 0;
  UTX_CUDA(cudaMemcpyAsync(&last_off, offs + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
  UTX_CUDA(cudaMemcpyAsync(&last_flag, flags + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
  if (extra_dev) UTX_CUDA(cudaMemcpyAsync(extra, extra_dev, 4, cudaMemcpyDeviceToHost, stream));
  UTX_CUDA(cudaStreamSynchronize(stream));
  *total = last_off + last_flag;
  return 0;
}


// Builds the per-view leaf grids in the free part of the workspace (behind the ray lists) and traces the listed rays through
// them.  *done stays false when the views are not axis-parallel orthographic ones, the mesh is tiny, or the lists do not fit.
int trace_with_leaf_grids(const void* nodes, int F, int n_views, const float* view_dirs_host, int perspective, const int* lists, int T,
                          const float4* rast, const float* vert, const int* tri, const Views& vw, const BakeWs& w, void* workspace,
                          size_t ws_bytes, unsigned g128, cudaStream_t stream, bool* done) {
  *done = false;
  static const int mode = std::getenv("UTX_RAY_IMPL") ? std::atoi(std::getenv("UTX_RAY_IMPL")) : 0;   // 1: always the tree walk (A/B)
  if (mode == 1 || perspective || F < 256) return 0;
  GridCfg gc{};
  gc.n = n_views;
  gc.F = F;
  for (int i = 0; i < n_views; ++i) {
    const double x = view_dirs_host[i * 3], y = view_dirs_host[i * 3 + 1], z = view_dirs_host[i * 3 + 2];
    const double l = std::sqrt(x * x + y * y + z * z);
    if (!(l > 0)) return 0;
    const double dn[3] = {std::fabs(x / l), std::fabs(y / l), std::fabs(z / l)};
    int a = dn[0] > dn[1] ? (dn[0] > dn[2] ? 0 : 2) : (dn[1] > dn[2] ? 1 : 2);
    if (dn[(a + 1) % 3] > 5e-7 || dn[(a + 2) % 3] > 5e-7) return 0;      // not parallel to an axis: eps would not cover the tilt
    gc.axis[i] = a;
  }
  int G = 32;
  while (G < 1024 && static_cast<long long>(G) * G * 2 < F) G *= 2;       // ~2 leaves per cell and view on average
  static const int g_override = std::getenv("UTX_RAY_GRID") ? std::atoi(std::getenv("UTX_RAY_GRID")) : 0;   // tuning knob
  if (g_override >= 8 && g_override <= 4096) G = g_override;
  gc.G = G;
  const long long ncells = static_cast<long long>(n_views) * G * G;
  // workspace behind the ray lists [n_views][T]: second ray lists | GridDev | count [ncells] | start [ncells + 1] | scan temp | list [cap]
  uint8_t* p = w.rest + al(static_cast<size_t>(n_views) * T * 4);
  int* lists2 = reinterpret_cast<int*>(p);
  p += al(static_cast<size_t>(n_views) * T * 4);
  int* counts2 = w.counters + 16;      // (zeroed with the other counters before texel_prep)
  uint8_t* end = static_cast<uint8_t*>(workspace) + ws_bytes;
  auto take = [&](size_t b) { uint8_t* q = p; p += al(b); return q; };
  GridDev* gd = reinterpret_cast<GridDev*>(take(sizeof(GridDev)));
  int* count = reinterpret_cast<int*>(take(static_cast<size_t>(ncells + 1) * 4));
  int* start = reinterpret_cast<int*>(take(static_cast<size_t>(ncells + 1) * 4));
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, count, start, static_cast<int>(ncells + 1));
  void* scan_tmp = take(scan_bytes);
  if (p >= end) return 0;
  const long long cap = (end - p) / 8;                       // two lists: as filled, and sorted
  int* gfill = reinterpret_cast<int*>(p);
  int* glist = gfill + cap;
  const float4* nodes4 = reinterpret_cast<const float4*>(nodes);
  const long long work = static_cast<long long>(n_views) * F;
  const unsigned gw = static_cast<unsigned>((work + 255) / 256);
  grid_setup_kernel<<<1, 1, 0, stream>>>(nodes4, G, gd);
  UTX_CUDA(cudaMemsetAsync(count, 0, static_cast<size_t>(ncells + 1) * 4, stream));
  grid_scatter_kernel<0><<<gw, 256, 0, stream>>>(nodes4, gc, gd, count, nullptr, nullptr, 0);
  UTX_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, count, start, static_cast<int>(ncells + 1), stream));
  int total = 0;
  UTX_CUDA(cudaMemcpyAsync(&total, start + ncells, 4, cudaMemcpyDeviceToHost, stream));
  UTX_CUDA(cudaStreamSynchronize(stream));
  if (total < 0 || total > cap) return 0;      // does not fit (or the 32-bit scan overflowed): the tree walk
  UTX_CUDA(cudaMemsetAsync(count, 0, static_cast<size_t>(ncells) * 4, stream));
  grid_scatter_kernel<1><<<gw, 256, 0, stream>>>(nodes4, gc, gd, count, start, gfill, cap);
  grid_sort_kernel<<<static_cast<unsigned>((ncells * 32 + 255) / 256), 256, 0, stream>>>(start, static_cast<int>(ncells), gfill, glist);
  ray_grid_kernel<<<dim3(g128, n_views), 128, 0, stream>>>(lists, w.counters, T, rast, w.pos, vert, tri, nodes4, gc, gd, start, glist, vw,
                                                           reinterpret_cast<unsigned*>(w.raw), lists2, counts2);
  // the rays of crowded cells through the hierarchy (blocks beyond the list's length exit at once)
  const float4* wide = reinterpret_cast<const float4*>(static_cast<const uint8_t*>(nodes) + wide_offset_bytes(F));
  ray_kernel<<<dim3(static_cast<unsigned>(2 * num_sms()), n_views), 128, 0, stream>>>(lists2, counts2, T, rast, w.pos, vert, tri, wide, vw,
                                                                                   reinterpret_cast<unsigned*>(w.raw));
  UTX_CUDA(cudaGetLastError());
  if (std::getenv("UTX_RAY_DEBUG")) {      // diagnosis: rays per view, and how many of them went to the tree walk
    int c[24];
    UTX_CUDA(cudaMemcpyAsync(c, w.counters, sizeof(c), cudaMemcpyDeviceToHost, stream));
    UTX_CUDA(cudaStreamSynchronize(stream));
    long r = 0, r2 = 0;
    for (int i = 0; i < n_views; ++i) { r += c[i]; r2 += c[16 + i]; }
    std::fprintf(stderr, "utx: leaf grids G=%d pairs=%d rays=%ld to_tree_walk=%ld\n", G, total, r, r2);
  }
  *done = true;
  return 0;
}

}  // namespace

int transform_points(const float* vert, int V, const float* mats, int n, float* out, cudaStream_t stream) {
  const long long tot = static_cast<long long>(n) * V;
  if (tot == 0) return 0;
  transform_points_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, stream>>>(vert, V, mats, n,
                                                                                        reinterpret_cast<float4*>(out));
  UTX_CUDA(cudaGetLastError());
  return 0;
}

size_t uv_bake_workspace_bytes(int H2, int W2) {
  const size_t T = static_cast<size_t>(H2) * W2;
  size_t pyr_c = 0, pyr_m = 0;
  for (int h = H2, w = W2, l = 0; l < 16 && h >= 1 && w >= 1; ++l, h /= 2, w /= 2) {
    pyr_c += al(static_cast<size_t>(h) * w * 12);
    pyr_m += al(static_cast<size_t>(h) * w);
  }
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<int*>(nullptr), static_cast<int*>(nullptr), static_cast<int>(T));
  return 6 * al(T) + al(T * 12) * 4 + al(T * 4) * 4 + al(scan_bytes) + al(bvh_nodes_bytes(static_cast<int>(T))) +
         al(bvh_workspace_bytes(static_cast<int>(T))) + pyr_c + pyr_m + 8192 + 256;
}

void uv_bake_layout(int H2, int W2, size_t* off_owner, size_t* off_pos, size_t* off_color, size_t* off_seam) {
  uint8_t* base = reinterpret_cast<uint8_t*>(static_cast<uintptr_t>(256));
  const BakeWs w = carve(base, H2 * W2);
  *off_owner = reinterpret_cast<uint8_t*>(w.owner) - base;
  *off_pos = reinterpret_cast<uint8_t*>(w.pos) - base;
  *off_color = reinterpret_cast<uint8_t*>(w.col_a) - base;
  *off_seam = w.seam - base;
}

// Stage 1 -- uv_to_pcd (:243-365) + the priority composite and seam mask of bake_mv_to_uv_reproject_blur (:591-605).
// Leaves in the workspace: owner i8 [T] (winning view or -1), pos fp32 [T,3], colour fp32 [T,3] (the owner's reprojected
// colour, 0 elsewhere), seam u8 [T].
int uv_bake_visibility(const float* vert, int V, const int* tri, int F, const void* nodes, const float* rast2d, int H2, int W2,
                       int n_views, const float* view_mats_host, const float* view_dirs_host, int perspective, const int* priority_host,
                       const float* images_rgba, int H, int W, float cos_thresh, unsigned char* mask2d,
                       unsigned char* mask_vis, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  (void)V;
  UTX_CHECK(n_views >= 1 && n_views <= MAXV, "uv_bake: 1..8 views");
  UTX_TRY(check_atlas(H2, W2, ws_bytes, "uv_bake_visibility"));
  const int T = H2 * W2;
  Views vw;
  vw.n = n_views;
  vw.perspective = perspective != 0;
  for (int i = 0; i < n_views; ++i) {
    for (int k = 0; k < 16; ++k) vw.mat[i][k] = view_mats_host[i * 16 + k];
    for (int k = 0; k < 3; ++k) vw.dir[i][k] = view_dirs_host[i * 3 + k];
    vw.priority[i] = priority_host[i];
  }
  const BakeWs w = carve(workspace, T);
  const unsigned g128 = (T + 127) / 128, g256 = (T + 255) / 256;
  const float4* rast = reinterpret_cast<const float4*>(rast2d);
  const float4* wide = reinterpret_cast<const float4*>(static_cast<const uint8_t*>(nodes) + wide_offset_bytes(F));
  // ray lists [n_views][T] live in the region the nearest-neighbour tree / pull-push pyramid use later (>= 160 T bytes >= 32 T)
  int* lists = reinterpret_cast<int*>(w.rest);
  UTX_CUDA(cudaMemsetAsync(w.counters, 0, 256, stream));
  texel_prep_kernel<<<g128, 128, 0, stream>>>(rast, H2, W2, vert, tri, vw, images_rgba, H, W, cos_thresh, w.raw, w.aok, w.pos,
                                              lists, w.counters);
  bool traced = false;
  UTX_TRY(trace_with_leaf_grids(nodes, F, n_views, view_dirs_host, perspective, lists, T, rast, vert, tri, vw, w, workspace, ws_bytes,
                                g128, stream, &traced));
  if (!traced)       // perspective or oblique views, tiny meshes, or a grid that would not fit: the tree walk
    ray_kernel<<<dim3(g128, n_views), 128, 0, stream>>>(lists, w.counters, T, rast, w.pos, vert, tri, wide, vw,
                                                        reinterpret_cast<unsigned*>(w.raw));
  repair3_kernel<<<g256, 256, 0, stream>>>(w.raw, w.rep3, H2, W2);
  repair5_kernel<<<g256, 256, 0, stream>>>(w.rep3, w.rep5, H2, W2, n_views);
  compose_kernel<<<g128, 128, 0, stream>>>(rast, T, vert, tri, vw, images_rgba, H, W, w.rep5, w.aok, mask2d, mask_vis, w.owner, w.col_a);
  seam0_kernel<<<g256, 256, 0, stream>>>(w.owner, w.b0, H2, W2);
  seam1_kernel<<<g256, 256, 0, stream>>>(w.b0, mask2d, w.seam, H2, W2);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// Stage 2 -- covered-but-unowned texels take the mean colour of their k nearest owned texels (3-D distance, exact):
// k = 1 is the fill of bake_mv_to_uv_reproject_blur (:606-615), k = 32 the `order_mean` tail of bake_mv_to_uv_kdtree (:427-432).
int uv_bake_fill(const unsigned char* mask2d, int H2, int W2, int k, int* nn_index_out, void* workspace, size_t ws_bytes,
                 cudaStream_t stream) {
  UTX_TRY(check_atlas(H2, W2, ws_bytes, "uv_bake_fill"));
  UTX_CHECK(k >= 1 && k <= KNN_MAX, "uv_bake_fill: k must be in 1..32");
  const int T = H2 * W2;
  const BakeWs w = carve(workspace, T);
  const unsigned g256 = (T + 255) / 256;
  // query list: T ints over the four visibility byte planes (raw | aok | rep3 | rep5), dead after stage 1
  int* qlist = reinterpret_cast<int*>(w.raw);
  int* qcount = w.counters + MAXV;
  UTX_CUDA(cudaMemsetAsync(qcount, 0, 4, stream));
  nn_flag_kernel<<<g256, 256, 0, stream>>>(mask2d, w.owner, T, w.flags, nn_index_out, qlist, qcount, H2, W2);
  int n_pts = 0, n_q = 0;
  UTX_TRY(count_flags(w.flags, w.offs, T, w.scan_tmp, w.scan_bytes, &n_pts, stream, qcount, &n_q));
  if (n_q == 0 || n_pts == 0) return 0;
  nn_compact_kernel<<<g256, 256, 0, stream>>>(w.owner, w.offs, w.pos, T, w.ids, w.pts);
  void* nn_nodes = w.rest;
  const size_t wsb = bvh_workspace_bytes(n_pts < 2 ? 2 : n_pts);
  void* nn_ws = w.rest + al(bvh_nodes_bytes(n_pts < 2 ? 2 : n_pts));
  UTX_TRY(point_bvh_build(w.pts, w.ids, n_pts, nn_nodes, nn_ws, wsb, stream));
  const PointTree pt = point_tree_view(nn_nodes, n_pts);
  const unsigned gq = (n_q + 127) / 128;
  if (k == 1)
  {
    int* next_run = w.counters + MAXV + 1;
    UTX_CUDA(cudaMemsetAsync(next_run, 0, 4, stream));
    const unsigned persistent = std::min<unsigned>(gq, static_cast<unsigned>(num_sms()) * 16u);   // 16 CTAs of 128 threads per SM
    // UTX_NN_IMPL (A/B knob; identical output): 0 (default) every run of 32 queries walks cooperatively; 1 cooperative only for
    // runs inside one 16-texel neighbourhood, per-lane walks otherwise; 2 the per-lane walks of r01.  teaser_robot, ncu
    // (profiles/r02_nn_variants.md): 2 -> 4.30 ms, 12.1 of 32 lanes, 1.81 G warp instructions, 539 M L1 wavefronts;
    // 1 -> 3.66 ms, 23.7 lanes; 0 -> 3.03 ms, 32.0 lanes, 1.41 G instructions, 317 M wavefronts.
    static const int impl = std::getenv("UTX_NN_IMPL") ? std::atoi(std::getenv("UTX_NN_IMPL")) : 0;
    if (impl == 2)
      nn_query_kernel<false><<<persistent, 128, 0, stream>>>(qlist, n_q, w.pos, pt, w.col_a, w.col_a, nn_index_out, W2, next_run, 0);
    else
      nn_query_kernel<true><<<persistent, 128, 0, stream>>>(qlist, n_q, w.pos, pt, w.col_a, w.col_a, nn_index_out, W2, next_run, impl != 1);
  }
  else
    knn_mean_kernel<<<gq, 128, 0, stream>>>(mask2d, w.owner, -1, w.pos, T, pt, k, w.col_a, w.col_a, nn_index_out, qlist, n_q);
#ifdef UTX_NN_DEBUG
  if (k == 1 && nn_index_out && std::getenv("UTX_NN_COUNT")) nn_count_kernel<<<gq, 128, 0, stream>>>(qlist, n_q, w.pos, pt, nn_index_out);
#endif
  UTX_CUDA(cudaGetLastError());
  return 0;
}

size_t uv_bake_views_workspace_bytes(int n_views, int H, int W) {
  const size_t N = static_cast<size_t>(n_views) * H * W;
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<int*>(nullptr), static_cast<int*>(nullptr), static_cast<int>(N));
  return al(N * 4) + al(N * 4 + 4) + 2 * al(N * 12) + al(scan_bytes) + al(bvh_nodes_bytes(static_cast<int>(N))) +
         al(bvh_workspace_bytes(static_cast<int>(N))) + 4096;
}

// Stage 2' -- the visible part of bake_mv_to_uv_kdtree: texels are coloured from the multi-view POINT CLOUD instead of the
// bilinear reprojection.  merge = 0 (`order_mean`, :406-417): every texel owned by view i takes the mean colour of its k nearest
// points of view i's pixel cloud.  merge = 1 (`mean`, :385-389): every covered texel takes the mean of its k nearest points of
// the union cloud (the caller skips the fill stage).  pix_pos [n,H,W,3]: vertex positions interpolated at the
// pixels (dr.interpolate, :188); images_rgba [n,H,W,4]: colour + visible alpha, pixels with alpha set form the clouds (:227-231).
int uv_bake_views_knn(const float* pix_pos, const float* images_rgba, int n_views, int H, int W, int k, int merge,
                      const unsigned char* mask2d, int H2, int W2, void* workspace, size_t ws_bytes, void* scratch,
                      size_t scratch_bytes, cudaStream_t stream) {
  UTX_TRY(check_atlas(H2, W2, ws_bytes, "uv_bake_views_knn"));
  UTX_CHECK(n_views >= 1 && n_views <= MAXV, "uv_bake_views_knn: 1..8 views");
  UTX_CHECK(k >= 1 && k <= KNN_MAX, "uv_bake_views_knn: k must be in 1..32");
  UTX_CHECK(scratch_bytes >= uv_bake_views_workspace_bytes(n_views, H, W), "uv_bake_views_knn: scratch too small");
  const int T = H2 * W2;
  const long long HW = static_cast<long long>(H) * W, N = HW * n_views;
  UTX_CHECK(N < (1LL << 31), "uv_bake_views_knn: too many pixels");
  const BakeWs w = carve(workspace, T);
  uint8_t* p = static_cast<uint8_t*>(scratch);
  auto take = [&](size_t b) { uint8_t* q = p; p += al(b); return q; };
  int* flags = reinterpret_cast<int*>(take(N * 4));
  int* offs = reinterpret_cast<int*>(take(N * 4 + 4));
  float* pts = reinterpret_cast<float*>(take(N * 12));
  float* cols = reinterpret_cast<float*>(take(N * 12));
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, offs, static_cast<int>(N));
  void* scan_tmp = take(scan_bytes);
  void* nodes = take(bvh_nodes_bytes(static_cast<int>(N)));
  const size_t tree_ws_bytes = bvh_workspace_bytes(static_cast<int>(N));
  void* tree_ws = take(tree_ws_bytes);
  const float4* rgba = reinterpret_cast<const float4*>(images_rgba);
  const unsigned gN = static_cast<unsigned>((N + 255) / 256);
  pix_flag_kernel<<<gN, 256, 0, stream>>>(rgba, N, flags);
  int total = 0;
  UTX_TRY(count_flags(flags, offs, N, scan_tmp, scan_bytes, &total, stream));
  pix_compact_kernel<<<gN, 256, 0, stream>>>(rgba, pix_pos, flags, offs, N, pts, cols);
  int start[MAXV + 1];
  start[0] = 0;
  start[n_views] = total;
  for (int v = 1; v < n_views; ++v)
    UTX_CUDA(cudaMemcpyAsync(&start[v], offs + v * HW, 4, cudaMemcpyDeviceToHost, stream));
  UTX_CUDA(cudaStreamSynchronize(stream));
  const unsigned g128 = (T + 127) / 128;
  if (merge) {
    UTX_CHECK(total >= 1, "uv_bake_views_knn: the views hold no visible pixel");
    UTX_TRY(point_bvh_build(pts, nullptr, total, nodes, tree_ws, tree_ws_bytes, stream));
    knn_mean_kernel<<<g128, 128, 0, stream>>>(mask2d, w.owner, -2, w.pos, T, point_tree_view(nodes, total), k, cols, w.col_a, nullptr);
  } else {
    for (int v = 0; v < n_views; ++v) {
      const int cnt = start[v + 1] - start[v];
      if (cnt == 0) continue;                              // an empty view owns no texel (ownership needs alpha > 0.999)
      const float* vp = pts + static_cast<size_t>(start[v]) * 3;
      UTX_TRY(point_bvh_build(vp, nullptr, cnt, nodes, tree_ws, tree_ws_bytes, stream));
      knn_mean_kernel<<<g128, 128, 0, stream>>>(mask2d, w.owner, v, w.pos, T, point_tree_view(nodes, cnt), k,
                                                cols + static_cast<size_t>(start[v]) * 3, w.col_a, nullptr);
    }
  }
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// Stage 3 -- seam blur (reproject only, :617-624) and pull-push into the texels outside the charts (:627, :423).
int uv_bake_finish(const unsigned char* mask2d, int H2, int W2, int blur, const float* blur_k2d, float blur_gamma,
                   float* color_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  UTX_TRY(check_atlas(H2, W2, ws_bytes, "uv_bake_finish"));
  const int T = H2 * W2;
  const BakeWs w = carve(workspace, T);
  const unsigned g256 = (T + 255) / 256;
  const float* src = w.col_a;
  if (blur) {
    lens_blur_kernel<<<g256, 256, 0, stream>>>(w.col_a, w.seam, blur_k2d, blur_gamma, H2, W2, w.col_b, blur == 2);
    src = w.col_b;
  }
  int levels = 0;
  for (int s = (H2 < W2 ? H2 : W2); s > 1; s >>= 1) ++levels;
  levels = levels - 2 > 0 ? levels - 2 : 0;
  if (levels == 0) {
    UTX_CUDA(cudaMemcpyAsync(color_out, src, static_cast<size_t>(T) * 12, cudaMemcpyDeviceToDevice, stream));
  } else {
    uint8_t* p = w.rest;
    auto take = [&](size_t b) { uint8_t* q = p; p += al(b); return q; };
    std::vector<float*> pc(levels + 1);
    std::vector<unsigned char*> pm(levels + 1);
    pc[0] = color_out;
    pm[0] = take(T);
    pp_init_kernel<<<g256, 256, 0, stream>>>(src, mask2d, T, pc[0], pm[0]);
    int h = H2, ww = W2;
    for (int l = 1; l <= levels; ++l) {
      pc[l] = reinterpret_cast<float*>(take(static_cast<size_t>(h / 2) * (ww / 2) * 12));
      pm[l] = take(static_cast<size_t>(h / 2) * (ww / 2));
      pp_down_kernel<<<((h / 2) * (ww / 2) + 255) / 256, 256, 0, stream>>>(pc[l - 1], pm[l - 1], h, ww, pc[l], pm[l]);
      h /= 2; ww /= 2;
    }
    for (int l = levels; l >= 1; --l) {
      const int hh = H2 >> (l - 1), wl = W2 >> (l - 1);
      pp_up_kernel<<<(hh * wl + 255) / 256, 256, 0, stream>>>(pc[l - 1], pm[l - 1], hh, wl, pc[l]);
    }
    UTX_CHECK(static_cast<size_t>(p - static_cast<uint8_t*>(workspace)) <= ws_bytes, "uv_bake: workspace overrun");
  }
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// The three stages back to back = NVDiffRendererInverse.infer(method='reproject', reproject_method='lens') (:635-726)
int uv_bake(const float* vert, int V, const int* tri, int F, const void* nodes, const float* rast2d, int H2, int W2,
            int n_views, const float* view_mats_host, const float* view_dirs_host, int perspective, const int* priority_host,
            const float* images_rgba, int H, int W, float cos_thresh, const float* blur_k2d, float blur_gamma,
            const float* grid_lo_host, float grid_extent, unsigned char* mask2d, unsigned char* mask_vis, float* color_out,
            int* nn_index_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  (void)grid_lo_host; (void)grid_extent;
  UTX_TRY(uv_bake_visibility(vert, V, tri, F, nodes, rast2d, H2, W2, n_views, view_mats_host, view_dirs_host, perspective, priority_host,
                             images_rgba, H, W, cos_thresh, mask2d, mask_vis, workspace, ws_bytes, stream));
  UTX_TRY(uv_bake_fill(mask2d, H2, W2, 1, nn_index_out, workspace, ws_bytes, stream));
  return uv_bake_finish(mask2d, H2, W2, 1, blur_k2d, blur_gamma, color_out, workspace, ws_bytes, stream);
}

}  // namespace utx
