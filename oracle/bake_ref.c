/* Oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): plain-C restatement of the geometric half of the UV bake.
 *
 *   ora_rasterize / ora_interpolate   nvdiffrast dr.rasterize / dr.interpolate as used at
 *                                     TextureTools/texturetools/render/nvdiffrast/renderer_inverse.py:183,188,273,277,288
 *                                     [ext nvdiffrast@729261dc, source absent: PARITY UNPINNED].  The tie rules the
 *                                     product and this oracle share are OURS and are stated in DESIGN.md: 8 sub-pixel
 *                                     bits, pixel-centre sampling, top-left fill rule, nearest z/w wins, ties -> lowest
 *                                     triangle id, (u,v) = weights of vertices 0 and 1, output (u, v, z/w, id+1).
 *   ora_lbvh_build                    rt_aprmis/bvhhelpers.py:20-83 + bvhworkers/get_elements.slang:1-40,
 *                                     lbvh_morton_codes.slang:24-79, lbvh_single_radixsort.slang (stable LSD sort),
 *                                     lbvh_hierarchy.slang:40-244, lbvh_bounding_boxes.slang:149-389 (exact unions)
 *   ora_intersect                     bvhworkers/intersect_test2.slang:14-146,269-309 INCLUDING its quirks: no t-range
 *                                     test in triangle_hit, hit_tid overwritten by every accepted hit (last accepted
 *                                     leaf in push-left/push-right/pop-right order wins), 1e-6 for zero direction
 *                                     components, 1e-9 determinant epsilon.
 *
 * Compile with -ffp-contract=off: every fp32 operation is a separately rounded IEEE operation, which is what the CUDA
 * side is held to as well (bake kernels are built with -fmad=false), so triangle ids are comparable bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ rasterizer */
#define SUBPIX 256

static int64_t edge_fn(int64_t ax, int64_t ay, int64_t bx, int64_t by, int64_t px, int64_t py) {
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}
static int tie_ok(int64_t ax, int64_t ay, int64_t bx, int64_t by) { /* top-left rule for interior e > 0 */
  int64_t dx = bx - ax, dy = by - ay;
  return (dy < 0) || (dy == 0 && dx > 0);
}
static uint32_t ordered_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static int32_t snap(float ndc, int size) {
  float p = (ndc * 0.5f + 0.5f) * (float)size;
  return (int32_t)floorf(p * (float)SUBPIX + 0.5f);
}

/* pos: [B or 1][V][4] clip coords; tri: [F][3]; out: [B][H][W][4] */
void ora_rasterize(const float* pos, int pos_batched, int V, const int32_t* tri, int F, int B, int H, int W, float* out) {
  uint64_t* zbuf = (uint64_t*)malloc((size_t)H * W * sizeof(uint64_t));
  for (int b = 0; b < B; ++b) {
    const float* P = pos + (pos_batched ? (size_t)b * V * 4 : 0);
    for (size_t i = 0; i < (size_t)H * W; ++i) zbuf[i] = ~0ull;
    for (int f = 0; f < F; ++f) {
      int32_t X[3], Y[3];
      float zn[3];
      int bad = 0;
      for (int k = 0; k < 3; ++k) {
        const float* p = P + (size_t)tri[f * 3 + k] * 4;
        if (!(p[3] > 0.0f)) { bad = 1; break; }
        X[k] = snap(p[0] / p[3], W);
        Y[k] = snap(p[1] / p[3], H);
        zn[k] = p[2] / p[3];
      }
      if (bad) continue;
      int64_t area = edge_fn(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
      if (area == 0) continue;
      int64_t sgn = area > 0 ? 1 : -1;
      int32_t minx = X[0] < X[1] ? X[0] : X[1]; if (X[2] < minx) minx = X[2];
      int32_t maxx = X[0] > X[1] ? X[0] : X[1]; if (X[2] > maxx) maxx = X[2];
      int32_t miny = Y[0] < Y[1] ? Y[0] : Y[1]; if (Y[2] < miny) miny = Y[2];
      int32_t maxy = Y[0] > Y[1] ? Y[0] : Y[1]; if (Y[2] > maxy) maxy = Y[2];
      int x0 = (minx - SUBPIX / 2 + SUBPIX - 1) >> 8, x1 = (maxx - SUBPIX / 2) >> 8;   /* pixel centres inside bbox */
      int y0 = (miny - SUBPIX / 2 + SUBPIX - 1) >> 8, y1 = (maxy - SUBPIX / 2) >> 8;
      if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > W - 1) x1 = W - 1; if (y1 > H - 1) y1 = H - 1;
      /* sign-normalised edge values E_k >= 0 inside; E0/E1/E2 weigh vertices 0/1/2.  The oriented edge opposite vertex 0
         is v1->v2 for positive area and v2->v1 otherwise (same for the others): that is what the tie rule looks at. */
      int t0 = sgn > 0 ? tie_ok(X[1], Y[1], X[2], Y[2]) : tie_ok(X[2], Y[2], X[1], Y[1]);
      int t1 = sgn > 0 ? tie_ok(X[2], Y[2], X[0], Y[0]) : tie_ok(X[0], Y[0], X[2], Y[2]);
      int t2 = sgn > 0 ? tie_ok(X[0], Y[0], X[1], Y[1]) : tie_ok(X[1], Y[1], X[0], Y[0]);
      float fa = (float)(area * sgn);
      for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
          int64_t px = (int64_t)x * SUBPIX + SUBPIX / 2, py = (int64_t)y * SUBPIX + SUBPIX / 2;
          int64_t e0 = edge_fn(X[1], Y[1], X[2], Y[2], px, py) * sgn;
          int64_t e1 = edge_fn(X[2], Y[2], X[0], Y[0], px, py) * sgn;
          int64_t e2 = edge_fn(X[0], Y[0], X[1], Y[1], px, py) * sgn;
          if (e0 < 0 || e1 < 0 || e2 < 0) continue;
          if ((e0 == 0 && !t0) || (e1 == 0 && !t1) || (e2 == 0 && !t2)) continue;
          float u = (float)e0 / fa, v = (float)e1 / fa, w2 = (float)e2 / fa;
          float zw = (u * zn[0] + v * zn[1]) + w2 * zn[2];
          uint64_t key = ((uint64_t)ordered_bits(zw) << 32) | (uint32_t)f;
          size_t idx = (size_t)y * W + x;
          if (key < zbuf[idx]) zbuf[idx] = key;
        }
    }
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float* o = out + (((size_t)b * H + y) * W + x) * 4;
        uint64_t key = zbuf[(size_t)y * W + x];
        if (key == ~0ull) { o[0] = o[1] = o[2] = o[3] = 0.0f; continue; }
        int f = (int)(uint32_t)key;
        int32_t X[3], Y[3];
        float zn[3], cw[3];
        for (int k = 0; k < 3; ++k) {
          const float* p = P + (size_t)tri[f * 3 + k] * 4;
          X[k] = snap(p[0] / p[3], W);
          Y[k] = snap(p[1] / p[3], H);
          zn[k] = p[2] / p[3];
          cw[k] = p[3];
        }
        int64_t area = edge_fn(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
        int64_t sgn = area > 0 ? 1 : -1;
        int64_t px = (int64_t)x * SUBPIX + SUBPIX / 2, py = (int64_t)y * SUBPIX + SUBPIX / 2;
        float fa = (float)(area * sgn);
        float u = (float)(edge_fn(X[1], Y[1], X[2], Y[2], px, py) * sgn) / fa;   /* weight of vertex 0 */
        float v = (float)(edge_fn(X[2], Y[2], X[0], Y[0], px, py) * sgn) / fa;   /* weight of vertex 1 */
        float w2 = (float)(edge_fn(X[0], Y[0], X[1], Y[1], px, py) * sgn) / fa;
        o[0] = u; o[1] = v;
        if (!(cw[0] == 1.0f && cw[1] == 1.0f && cw[2] == 1.0f)) {   /* perspective-correct weights (nvdiffrast [ext]); w == 1: unchanged */
          float a0 = u / cw[0], a1 = v / cw[1], a2 = w2 / cw[2];
          float sum = (a0 + a1) + a2;
          o[0] = a0 / sum; o[1] = a1 / sum;
        }
        o[2] = (u * zn[0] + v * zn[1]) + w2 * zn[2];
        o[3] = (float)(f + 1);
      }
  }
  free(zbuf);
}

/* attr: [B or 1][V][C]; rast: [B][H][W][4]; out [B][H][W][C] = u a0 + v a1 + (1-u-v) a2, 0 on background */
void ora_interpolate(const float* attr, int attr_batched, int V, int C, const float* rast, const int32_t* tri, int B, int H,
                     int W, float* out) {
  for (int b = 0; b < B; ++b)
    for (size_t p = 0; p < (size_t)H * W; ++p) {
      const float* r = rast + ((size_t)b * H * W + p) * 4;
      float* o = out + ((size_t)b * H * W + p) * C;
      int f = (int)r[3] - 1;
      if (f < 0) { for (int c = 0; c < C; ++c) o[c] = 0.0f; continue; }
      const float* A = attr + (attr_batched ? (size_t)b * V * C : 0);
      const float *a0 = A + (size_t)tri[f * 3] * C, *a1 = A + (size_t)tri[f * 3 + 1] * C, *a2 = A + (size_t)tri[f * 3 + 2] * C;
      float u = r[0], v = r[1], w = (1.0f - u) - v;
      for (int c = 0; c < C; ++c) o[c] = (u * a0[c] + v * a1[c]) + w * a2[c];
    }
}

/* ------------------------------------------------------------------------------------------------ LBVH build */
static uint32_t expand_bits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static uint32_t morton3d(float x, float y, float z) {
  x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
  y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
  z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
  return expand_bits((uint32_t)x) * 4 + expand_bits((uint32_t)y) * 2 + expand_bits((uint32_t)z);
}
static int find_msb(uint32_t v) {
  if (v == 0) return -1;
  int m = 31;
  while (!((v >> m) & 1)) m--;
  return m;
}
static int delta(int i, uint32_t code_i, int j, int n, const uint32_t* codes) {
  if (j < 0 || j > n - 1) return -1;
  uint32_t code_j = codes[j];
  if (code_i == code_j) return 32 + 31 - find_msb((uint32_t)i ^ (uint32_t)j);
  return 31 - find_msb(code_i ^ code_j);
}

/* info: [2F-1][3] (left,right,prim); aabb: [2F-1][6]; also returns sorted (code, element) for inspection */
void ora_lbvh_build(const float* vert, int V, const int32_t* tri, int F, int32_t* info, float* aabb, int32_t* sorted_out) {
  (void)V;
  float* eab = (float*)malloc((size_t)F * 6 * sizeof(float));
  float gmin[3] = {INFINITY, INFINITY, INFINITY}, gmax[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int f = 0; f < F; ++f) {
    float mn[3] = {1e9f, 1e9f, 1e9f}, mx[3] = {-1e9f, -1e9f, -1e9f};
    for (int k = 0; k < 3; ++k) {
      const float* v = vert + (size_t)tri[f * 3 + k] * 3;
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], v[a]); mx[a] = fmaxf(mx[a], v[a]); }
    }
    for (int a = 0; a < 3; ++a) {
      eab[f * 6 + a] = fminf(mn[a], mx[a]);
      eab[f * 6 + 3 + a] = fmaxf(mn[a], mx[a]);
      gmin[a] = fminf(gmin[a], eab[f * 6 + a]);
      gmax[a] = fmaxf(gmax[a], eab[f * 6 + 3 + a]);
    }
  }
  uint32_t* codes = (uint32_t*)malloc((size_t)F * 4);
  uint32_t* elem = (uint32_t*)malloc((size_t)F * 4);
  for (int f = 0; f < F; ++f) {
    float m[3];
    for (int a = 0; a < 3; ++a) {
      float lo = eab[f * 6 + a], hi = eab[f * 6 + 3 + a];
      float center = lo + 0.5f * (hi - lo);
      m[a] = (center - gmin[a]) / (gmax[a] - gmin[a]);
    }
    codes[f] = morton3d(m[0], m[1], m[2]);
    elem[f] = (uint32_t)f;
  }
  /* stable LSD radix sort, 4 passes x 8 bits (lbvh_single_radixsort.slang) */
  uint32_t* c2 = (uint32_t*)malloc((size_t)F * 4);
  uint32_t* e2 = (uint32_t*)malloc((size_t)F * 4);
  for (int pass = 0; pass < 4; ++pass) {
    size_t hist[257];
    memset(hist, 0, sizeof(hist));
    for (int i = 0; i < F; ++i) hist[((codes[i] >> (8 * pass)) & 255) + 1]++;
    for (int i = 0; i < 256; ++i) hist[i + 1] += hist[i];
    for (int i = 0; i < F; ++i) {
      size_t d = hist[(codes[i] >> (8 * pass)) & 255]++;
      c2[d] = codes[i]; e2[d] = elem[i];
    }
    uint32_t* t = codes; codes = c2; c2 = t;
    t = elem; elem = e2; e2 = t;
  }
  if (sorted_out)
    for (int i = 0; i < F; ++i) { sorted_out[2 * i] = (int32_t)codes[i]; sorted_out[2 * i + 1] = (int32_t)elem[i]; }
  const int LEAF = F - 1;
  int32_t* parent = (int32_t*)malloc((size_t)(2 * F - 1) * 4);
  for (int g = 0; g < F; ++g) {
    int e = (int)elem[g];
    info[(LEAF + g) * 3 + 0] = 0; info[(LEAF + g) * 3 + 1] = 0; info[(LEAF + g) * 3 + 2] = e;
    memcpy(aabb + (size_t)(LEAF + g) * 6, eab + (size_t)e * 6, 24);
  }
  parent[0] = 0;
  for (int g = 0; g < F - 1; ++g) {
    /* determineRange */
    uint32_t code = codes[g];
    int dL = delta(g, code, g - 1, F, codes), dR = delta(g, code, g + 1, F, codes);
    int d = (dR >= dL) ? 1 : -1;
    int dmin = dL < dR ? dL : dR;
    int lmax = 2;
    while (delta(g, code, g + lmax * d, F, codes) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1)
      if (delta(g, code, g + (l + t) * d, F, codes) > dmin) l += t;
    int j = g + l * d;
    int first = g < j ? g : j, last = g > j ? g : j;
    /* findSplit */
    uint32_t fcode = codes[first];
    int common = delta(first, fcode, last, F, codes);
    int split = first, stride = last - first;
    do {
      stride = (stride + 1) >> 1;
      int ns = split + stride;
      if (ns < last && delta(first, fcode, ns, F, codes) > common) split = ns;
    } while (stride > 1);
    int ca = (split == first) ? LEAF + split : split;
    int cb = (split + 1 == last) ? LEAF + split + 1 : split + 1;
    info[g * 3 + 0] = ca; info[g * 3 + 1] = cb; info[g * 3 + 2] = 0;
    parent[ca] = g; parent[cb] = g;
  }
  /* exact bottom-up unions (the per-level refit of lbvh_bounding_boxes.slang converges to exactly this) */
  int* done = (int*)calloc((size_t)(F > 1 ? F - 1 : 1), sizeof(int));
  for (int g = 0; g < F - 1; ++g)
    for (int a = 0; a < 3; ++a) { aabb[(size_t)g * 6 + a] = 1e9f; aabb[(size_t)g * 6 + 3 + a] = -1e9f; }
  if (F > 1)
    for (int g = 0; g < F; ++g) {
      int n = parent[LEAF + g];
      for (;;) {
        if (++done[n] < 2) break;   /* second arrival: both children final */
        int ca = info[n * 3], cb = info[n * 3 + 1];
        for (int a = 0; a < 3; ++a) {
          aabb[(size_t)n * 6 + a] = fminf(aabb[(size_t)ca * 6 + a], aabb[(size_t)cb * 6 + a]);
          aabb[(size_t)n * 6 + 3 + a] = fmaxf(aabb[(size_t)ca * 6 + 3 + a], aabb[(size_t)cb * 6 + 3 + a]);
        }
        if (n == 0) break;
        n = parent[n];
      }
    }
  free(done); free(parent); free(c2); free(e2); free(codes); free(elem); free(eab);
}

/* ------------------------------------------------------------------------------------------------ intersect */
typedef struct { float x, y, z; } v3;
static v3 sub3(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static v3 cross3(v3 a, v3 b) { v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; return r; }
static float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

static int aabb_hit(v3 o, v3 d, float tmin, float tmax, const float* bb) {
  const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  for (int i = 0; i < 3; ++i) {
    float di = dd[i];
    if (di == 0.0f) di = 0.000001f;
    float inv = 1.0f / di;
    float t0 = (bb[i] - oo[i]) * inv, t1 = (bb[3 + i] - oo[i]) * inv;
    if (inv < 0.0f) { float t = t1; t1 = t0; t0 = t; }
    tmin = t0 > tmin ? t0 : tmin;
    tmax = t1 < tmax ? t1 : tmax;
    if (tmax < tmin) return 0;
  }
  return 1;
}
static int triangle_hit(v3 o, v3 d, v3 v0, v3 v1, v3 v2, float* t_hit, float* u_out, float* v_out) {
  const float eps = 1e-9f;
  v3 E1 = sub3(v1, v0), E2 = sub3(v2, v0);
  v3 P = cross3(d, E2);
  float det = dot3(E1, P);
  if (det > -eps && det < eps) return 0;
  float inv = 1.0f / det;
  v3 T = sub3(o, v0);
  float u = dot3(T, P) * inv;
  if (u < 0 || u > 1) return 0;
  v3 Q = cross3(T, E1);
  float v = dot3(d, Q) * inv;
  if (v < 0 || u + v > 1) return 0;
  *t_hit = dot3(E2, Q) * inv;   /* NOTE: no t-range test (reference quirk a) */
  *u_out = u; *v_out = v;
  return 1;
}

void ora_intersect(const float* vert, const int32_t* tri, const int32_t* info, const float* aabb, const float* rays_o,
                   const float* rays_d, int64_t N, uint8_t* hit, int32_t* tid, float* pos, float* uv) {
  for (int64_t r = 0; r < N; ++r) {
    v3 o = {rays_o[r * 3], rays_o[r * 3 + 1], rays_o[r * 3 + 2]};
    v3 d = {rays_d[r * 3], rays_d[r * 3 + 1], rays_d[r * 3 + 2]};
    float len = sqrtf(dot3(d, d));
    d.x = d.x / len; d.y = d.y / len; d.z = d.z / len;
    int stack[64], count = 0;
    stack[count++] = 0;
    float closest = 1e9f, hit_t = 0.f, hu = 0.f, hv = 0.f;
    int any = 0, htid = -1;
    while (count > 0) {
      int n = stack[--count];
      if (!aabb_hit(o, d, 0.0f, closest, aabb + (size_t)n * 6)) continue;
      int l = info[n * 3], rr = info[n * 3 + 1];
      if (l != 0 && rr != 0) {
        if (count + 2 <= 64) { stack[count++] = l; stack[count++] = rr; }
      } else if (l == 0 && rr == 0) {
        int p = info[n * 3 + 2];
        const float *a = vert + (size_t)tri[p * 3] * 3, *b = vert + (size_t)tri[p * 3 + 1] * 3, *c = vert + (size_t)tri[p * 3 + 2] * 3;
        v3 v0 = {a[0], a[1], a[2]}, v1 = {b[0], b[1], b[2]}, v2 = {c[0], c[1], c[2]};
        float t, u, v;
        if (triangle_hit(o, d, v0, v1, v2, &t, &u, &v)) {
          closest = t < closest ? t : closest;
          any = 1; htid = p; hit_t = closest; hu = u; hv = v;   /* quirk b: id of the LAST accepted leaf */
        }
      }
    }
    hit[r] = (uint8_t)any;
    tid[r] = any ? htid : -1;
    if (any) {
      pos[r * 3] = o.x + hit_t * d.x; pos[r * 3 + 1] = o.y + hit_t * d.y; pos[r * 3 + 2] = o.z + hit_t * d.z;
      uv[r * 2] = hu; uv[r * 2 + 1] = hv;
    } else {
      pos[r * 3] = pos[r * 3 + 1] = pos[r * 3 + 2] = 0.f; uv[r * 2] = uv[r * 2 + 1] = 0.f;
    }
  }
}

/* The same query WITHOUT the hierarchy: scan the leaves in descending order of their position g in the Morton-sorted list and
 * test each leaf's own box against the running closest t.  Identical results to ora_intersect (tests/test_oracle_bake.py):
 * the right-first depth-first walk reaches leaves in descending g, and because boxes nest and closest only shrinks, a leaf's own
 * test passing implies every ancestor's test passed.  This is the statement the product's leaf-grid ray kernel rests on. */
void ora_intersect_leafscan(const float* vert, const int32_t* tri, const int32_t* info, const float* aabb, int F, const float* rays_o,
                            const float* rays_d, int64_t N, uint8_t* hit, int32_t* tid, float* pos, float* uv) {
  const int LEAF = F - 1;
  for (int64_t r = 0; r < N; ++r) {
    v3 o = {rays_o[r * 3], rays_o[r * 3 + 1], rays_o[r * 3 + 2]};
    v3 d = {rays_d[r * 3], rays_d[r * 3 + 1], rays_d[r * 3 + 2]};
    float len = sqrtf(dot3(d, d));
    d.x = d.x / len; d.y = d.y / len; d.z = d.z / len;
    float closest = 1e9f, hit_t = 0.f, hu = 0.f, hv = 0.f;
    int any = 0, htid = -1;
    for (int g = F - 1; g >= 0; --g) {
      const int n = LEAF + g;
      if (!aabb_hit(o, d, 0.0f, closest, aabb + (size_t)n * 6)) continue;
      int p = info[n * 3 + 2];
      const float *a = vert + (size_t)tri[p * 3] * 3, *b = vert + (size_t)tri[p * 3 + 1] * 3, *c = vert + (size_t)tri[p * 3 + 2] * 3;
      v3 v0 = {a[0], a[1], a[2]}, v1 = {b[0], b[1], b[2]}, v2 = {c[0], c[1], c[2]};
      float t, u, v;
      if (triangle_hit(o, d, v0, v1, v2, &t, &u, &v)) {
        closest = t < closest ? t : closest;
        any = 1; htid = p; hit_t = closest; hu = u; hv = v;
      }
    }
    hit[r] = (uint8_t)any;
    tid[r] = any ? htid : -1;
    if (any) {
      pos[r * 3] = o.x + hit_t * d.x; pos[r * 3 + 1] = o.y + hit_t * d.y; pos[r * 3 + 2] = o.z + hit_t * d.z;
      uv[r * 2] = hu; uv[r * 2 + 1] = hv;
    } else {
      pos[r * 3] = pos[r * 3 + 1] = pos[r * 3 + 2] = 0.f; uv[r * 2] = uv[r * 2 + 1] = 0.f;
    }
  }
}

