// Triangle rasteriser + attribute interpolation for the bake (replaces nvdiffrast's dr.rasterize / dr.interpolate as
// called at TextureTools/texturetools/render/nvdiffrast/renderer_inverse.py:183,188,273,277,288 and
// renderer_base.py:142,173,191).  Output convention is nvdiffrast's: [B,H,W,4] = (u, v, z/w, triangle_id + 1),
// 0 = background, u/v = weights of triangle vertices 0 and 1, row 0 at y_clip = -1.
//
// Coverage is decided in integers (8 sub-pixel bits, pixel centres, top-left rule), depth by an order-preserving
// 64-bit key (depth bits << 32 | triangle id) and atomicMin, so the result does not depend on thread scheduling:
// nearest z/w wins, ties go to the lowest triangle id.  Small triangles (<= 8 pixel centres in the box) are drawn by the
// thread that set them up, the rest by persistent warps whose lanes stride over the bounding box; a resolve pass then
// recomputes (u, v, z/w) of the winner per pixel with fully coalesced float4 stores.  Built with -fmad=false: every fp32 op is separately rounded, so
// the CPU oracle reproduces ids AND barycentrics bit for bit.
#include "common.h"
#include "kernels.h"

namespace utx {
namespace {

constexpr int SUBPIX = 256;

__device__ __forceinline__ long long edge_fn(long long ax, long long ay, long long bx, long long by, long long px,
                                             long long py) {
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}
__device__ __forceinline__ bool tie_ok(long long ax, long long ay, long long bx, long long by) {
  const long long dx = bx - ax, dy = by - ay;
  return (dy < 0) || (dy == 0 && dx > 0);
}
__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ int snap(float ndc, int size) {
  const float p = (ndc * 0.5f + 0.5f) * static_cast<float>(size);
  return static_cast<int>(floorf(p * static_cast<float>(SUBPIX) + 0.5f));
}

struct TriSetup {
  int X[3], Y[3];
  float zn[3], w[3];
  long long area;
  bool ok;
};
__device__ __forceinline__ TriSetup setup_tri(const float* __restrict__ P, const int* __restrict__ tri, int f, int H,
                                              int W) {
  TriSetup t;
  t.ok = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 p = *reinterpret_cast<const float4*>(P + static_cast<size_t>(tri[f * 3 + k]) * 4);
    if (!(p.w > 0.0f)) t.ok = false;
    t.X[k] = snap(p.x / p.w, W);
    t.Y[k] = snap(p.y / p.w, H);
    t.zn[k] = p.z / p.w;
    t.w[k] = p.w;
  }
  t.area = edge_fn(t.X[0], t.Y[0], t.X[1], t.Y[1], t.X[2], t.Y[2]);
  if (t.area == 0) t.ok = false;
  return t;
}

__global__ void __launch_bounds__(256) raster_clear_kernel(unsigned long long* zbuf, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) zbuf[i] = ~0ull;
}

// Pass 1, one THREAD per (view, triangle): set-up; triangles whose bounding box holds <= kSmallBox pixel centres (the common
// case of the 6 x 512^2 view pass: ~0.5 pixel per triangle) are drawn right here, the others are appended to a work list.
// Pass 2, persistent WARPS over that list: lanes stride over the bounding box.
constexpr int kSmallBox = 8;

struct TriDraw {
  TriSetup t;
  long long sgn;
  int x0, y0, bw;
  long long npix;
  bool t0, t1, t2;
  float fa;
};
__device__ __forceinline__ bool prepare_draw(const float* P, const int* tri, int f, int H, int W, TriDraw& d) {
  d.t = setup_tri(P, tri, f, H, W);
  if (!d.t.ok) return false;
  const TriSetup& t = d.t;
  d.sgn = t.area > 0 ? 1 : -1;
  const int minx = min(t.X[0], min(t.X[1], t.X[2])), maxx = max(t.X[0], max(t.X[1], t.X[2]));
  const int miny = min(t.Y[0], min(t.Y[1], t.Y[2])), maxy = max(t.Y[0], max(t.Y[1], t.Y[2]));
  int x0 = (minx - SUBPIX / 2 + SUBPIX - 1) >> 8, x1 = (maxx - SUBPIX / 2) >> 8;
  int y0 = (miny - SUBPIX / 2 + SUBPIX - 1) >> 8, y1 = (maxy - SUBPIX / 2) >> 8;
  x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, W - 1); y1 = min(y1, H - 1);
  if (x1 < x0 || y1 < y0) return false;
  const long long sgn = d.sgn;
  d.t0 = sgn > 0 ? tie_ok(t.X[1], t.Y[1], t.X[2], t.Y[2]) : tie_ok(t.X[2], t.Y[2], t.X[1], t.Y[1]);
  d.t1 = sgn > 0 ? tie_ok(t.X[2], t.Y[2], t.X[0], t.Y[0]) : tie_ok(t.X[0], t.Y[0], t.X[2], t.Y[2]);
  d.t2 = sgn > 0 ? tie_ok(t.X[0], t.Y[0], t.X[1], t.Y[1]) : tie_ok(t.X[1], t.Y[1], t.X[0], t.Y[0]);
  d.fa = static_cast<float>(t.area * sgn);
  d.x0 = x0; d.y0 = y0; d.bw = x1 - x0 + 1;
  d.npix = static_cast<long long>(d.bw) * (y1 - y0 + 1);
  return true;
}
__device__ __forceinline__ void draw_pixels(const TriDraw& d, int f, int W, unsigned long long* zb, int first, int step) {
  const TriSetup& t = d.t;
  for (long long i = first; i < d.npix; i += step) {
    const int x = d.x0 + static_cast<int>(i % d.bw), y = d.y0 + static_cast<int>(i / d.bw);
    const long long px = static_cast<long long>(x) * SUBPIX + SUBPIX / 2, py = static_cast<long long>(y) * SUBPIX + SUBPIX / 2;
    const long long e0 = edge_fn(t.X[1], t.Y[1], t.X[2], t.Y[2], px, py) * d.sgn;
    const long long e1 = edge_fn(t.X[2], t.Y[2], t.X[0], t.Y[0], px, py) * d.sgn;
    const long long e2 = edge_fn(t.X[0], t.Y[0], t.X[1], t.Y[1], px, py) * d.sgn;
    if (e0 < 0 || e1 < 0 || e2 < 0) continue;
    if ((e0 == 0 && !d.t0) || (e1 == 0 && !d.t1) || (e2 == 0 && !d.t2)) continue;
    const float u = static_cast<float>(e0) / d.fa, v = static_cast<float>(e1) / d.fa, w2 = static_cast<float>(e2) / d.fa;
    const float zw = (u * t.zn[0] + v * t.zn[1]) + w2 * t.zn[2];
    const unsigned long long key = (static_cast<unsigned long long>(ordered_bits(zw)) << 32) | static_cast<unsigned>(f);
    atomicMin(zb + static_cast<size_t>(y) * W + x, key);
  }
}

__global__ void __launch_bounds__(256) raster_small_kernel(const float* __restrict__ pos, int pos_batched, int V,
                                                           const int* __restrict__ tri, int F, int B, int H, int W,
                                                           unsigned long long* __restrict__ zbuf, unsigned* __restrict__ list,
                                                           unsigned* __restrict__ count) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<long long>(F) * B) return;
  const int b = static_cast<int>(g / F), f = static_cast<int>(g % F);
  const float* P = pos + (pos_batched ? static_cast<size_t>(b) * V * 4 : 0);
  TriDraw d;
  if (!prepare_draw(P, tri, f, H, W, d)) return;
  if (d.npix > kSmallBox) {
    list[atomicAdd(count, 1u)] = static_cast<unsigned>(g);
    return;
  }
  draw_pixels(d, f, W, zbuf + static_cast<size_t>(b) * H * W, 0, 1);
}

__global__ void __launch_bounds__(256) raster_large_kernel(const float* __restrict__ pos, int pos_batched, int V,
                                                           const int* __restrict__ tri, int F, int H, int W,
                                                           unsigned long long* __restrict__ zbuf,
                                                           const unsigned* __restrict__ list,
                                                           const unsigned* __restrict__ count) {
  const unsigned n = *count;
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const unsigned g = list[i];
    const int b = static_cast<int>(g / F), f = static_cast<int>(g % F);
    const float* P = pos + (pos_batched ? static_cast<size_t>(b) * V * 4 : 0);
    TriDraw d;
    if (!prepare_draw(P, tri, f, H, W, d)) continue;
    draw_pixels(d, f, W, zbuf + static_cast<size_t>(b) * H * W, lane, 32);
  }
}

__global__ void __launch_bounds__(256) raster_resolve_kernel(const float* __restrict__ pos, int pos_batched, int V,
                                                             const int* __restrict__ tri, int B, int H, int W,
                                                             const unsigned long long* __restrict__ zbuf,
                                                             float4* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n = static_cast<size_t>(B) * H * W;
  if (i >= n) return;
  const unsigned long long key = zbuf[i];
  if (key == ~0ull) {
    out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int b = static_cast<int>(i / (static_cast<size_t>(H) * W));
  const int rem = static_cast<int>(i % (static_cast<size_t>(H) * W));
  const int y = rem / W, x = rem % W;
  const int f = static_cast<int>(static_cast<unsigned>(key));
  const float* P = pos + (pos_batched ? static_cast<size_t>(b) * V * 4 : 0);
  const TriSetup t = setup_tri(P, tri, f, H, W);
  const long long sgn = t.area > 0 ? 1 : -1;
  const long long px = static_cast<long long>(x) * SUBPIX + SUBPIX / 2, py = static_cast<long long>(y) * SUBPIX + SUBPIX / 2;
  const float fa = static_cast<float>(t.area * sgn);
  const float u = static_cast<float>(edge_fn(t.X[1], t.Y[1], t.X[2], t.Y[2], px, py) * sgn) / fa;
  const float v = static_cast<float>(edge_fn(t.X[2], t.Y[2], t.X[0], t.Y[0], px, py) * sgn) / fa;
  const float w2 = static_cast<float>(edge_fn(t.X[0], t.Y[0], t.X[1], t.Y[1], px, py) * sgn) / fa;
  // (u, v) are perspective-correct like nvdiffrast's: screen-space weights divided by the clip w of their vertex and
  // renormalised.  Skipped when all three w are exactly 1 (orthographic views, the UV-space raster): results there are unchanged.
  float uu = u, vv = v;
  if (!(t.w[0] == 1.0f && t.w[1] == 1.0f && t.w[2] == 1.0f)) {
    const float a0 = u / t.w[0], a1 = v / t.w[1], a2 = w2 / t.w[2];
    const float sum = (a0 + a1) + a2;
    uu = a0 / sum;
    vv = a1 / sum;
  }
  out[i] = make_float4(uu, vv, (u * t.zn[0] + v * t.zn[1]) + w2 * t.zn[2], static_cast<float>(f + 1));
}

__global__ void __launch_bounds__(256) interpolate_kernel(const float* __restrict__ attr, int attr_batched, int V, int C,
                                                          const float4* __restrict__ rast, const int* __restrict__ tri,
                                                          int B, size_t HW, float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * HW) return;
  const float4 r = rast[i];
  float* o = out + i * C;
  const int f = static_cast<int>(r.w) - 1;
  if (f < 0) {
    for (int c = 0; c < C; ++c) o[c] = 0.f;
    return;
  }
  const int b = static_cast<int>(i / HW);
  const float* A = attr + (attr_batched ? static_cast<size_t>(b) * V * C : 0);
  const float *a0 = A + static_cast<size_t>(tri[f * 3]) * C, *a1 = A + static_cast<size_t>(tri[f * 3 + 1]) * C,
              *a2 = A + static_cast<size_t>(tri[f * 3 + 2]) * C;
  const float u = r.x, v = r.y, w = (1.0f - u) - v;
  for (int c = 0; c < C; ++c) o[c] = (u * a0[c] + v * a1[c]) + w * a2[c];
}

}  // namespace

size_t rasterize_workspace_bytes(int B, int H, int W, int F) {
  return static_cast<size_t>(B) * H * W * 8 + static_cast<size_t>(B) * F * 4 + 256;   // depth/id keys + large-triangle list + counter
}

int rasterize(const float* pos, int pos_batched, int V, const int* tri, int F, int B, int H, int W, float* rast_out,
              void* workspace, cudaStream_t stream) {
  UTX_CHECK(B > 0 && H > 0 && W > 0 && H <= 8192 && W <= 8192, "rasterize: bad viewport");
  UTX_CHECK((reinterpret_cast<uintptr_t>(pos) & 15) == 0 && (reinterpret_cast<uintptr_t>(rast_out) & 15) == 0,
            "rasterize: pos/rast_out must be 16B aligned");
  unsigned long long* zbuf = static_cast<unsigned long long*>(workspace);
  const size_t n = static_cast<size_t>(B) * H * W;
  raster_clear_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(zbuf, n);
  if (F > 0) {
    const long long tris = static_cast<long long>(F) * B;
    UTX_CHECK(tris < (1ll << 32), "rasterize: too many (view, triangle) pairs");
    unsigned* list = reinterpret_cast<unsigned*>(static_cast<uint8_t*>(workspace) + n * 8);
    unsigned* count = list + tris;
    UTX_CUDA(cudaMemsetAsync(count, 0, 4, stream));
    raster_small_kernel<<<static_cast<unsigned>((tris + 255) / 256), 256, 0, stream>>>(pos, pos_batched, V, tri, F, B, H, W, zbuf,
                                                                                      list, count);
    raster_large_kernel<<<num_sms() * 8, 256, 0, stream>>>(pos, pos_batched, V, tri, F, H, W, zbuf, list, count);
  }
  raster_resolve_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      pos, pos_batched, V, tri, B, H, W, zbuf, reinterpret_cast<float4*>(rast_out));
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int interpolate(const float* attr, int attr_batched, int V, int C, const float* rast, const int* tri, int B, int H, int W,
                float* out, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(B) * H * W;
  if (n == 0) return 0;
  interpolate_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      attr, attr_batched, V, C, reinterpret_cast<const float4*>(rast), tri, B, static_cast<size_t>(H) * W, out);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace utx
