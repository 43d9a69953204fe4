// HBM-bound pieces of the FLUX VAE (AutoencoderKL [ext diffusers]; reference call sites
// flux_piplines/texturing/pipeline.py:226-238 encode, :688-692 decode).  Activations are NHWC bf16 so that every 3x3
// convolution is one tcgen05 GEMM (gemm_sm100.cu) over an im2col operand:  [N*Ho*Wo, 9*Cin] x [Cout, 9*Cin]^T.
//   im2col3x3        gather with zero padding; optional nearest-2x upsample folded into the gather (Upsample2D) and
//                    stride-2 / asymmetric (0,1,0,1) padding (Downsample2D)
//   groupnorm        GroupNorm(32, eps 1e-6, affine) [+ SiLU]: fp64-accumulated statistics, one normalise pass
//   softmax_rows     fp32 scores -> bf16 probabilities (single-head 512-dim attention of the mid block)
//   transpose        [R,C] -> [C,R] bf16 (V^T operand for P @ V on the K-major GEMM)
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {
namespace {

__global__ void __launch_bounds__(256) im2col_kernel(const bf16* __restrict__ x, int N, int Hin, int Win, int C, int up,
                                                     int stride, int pad, int Ho, int Wo, int Kpad,
                                                     bf16* __restrict__ out) {
  // one thread per (output pixel, 8-wide K chunk)
  const int kchunks = Kpad / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(N) * Ho * Wo * kchunks;
  if (i >= total) return;
  const int kc = static_cast<int>(i % kchunks);
  const long long pix = i / kchunks;
  const int xo = static_cast<int>(pix % Wo), yo = static_cast<int>((pix / Wo) % Ho), n = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
  const int Hs = Hin * up, Ws = Win * up;
  uint4 val = make_uint4(0, 0, 0, 0);
  const int k0 = kc * 8;
  if ((C & 7) == 0) {
    if (k0 < 9 * C) {
      const int tap = k0 / C, c = k0 % C;
      const int ys = yo * stride + tap / 3 - pad, xs = xo * stride + tap % 3 - pad;
      if (ys >= 0 && ys < Hs && xs >= 0 && xs < Ws)
        val = *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * Hin + ys / up) * Win + xs / up) * C + c);
    }
  } else {
    unsigned short e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      e[j] = 0;
      if (k < 9 * C) {
        const int tap = k / C, c = k % C;
        const int ys = yo * stride + tap / 3 - pad, xs = xo * stride + tap % 3 - pad;
        if (ys >= 0 && ys < Hs && xs >= 0 && xs < Ws)
          e[j] = reinterpret_cast<const unsigned short*>(x)[((static_cast<long long>(n) * Hin + ys / up) * Win + xs / up) * C + c];
      }
    }
    val = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
  }
  *reinterpret_cast<uint4*>(out + pix * Kpad + k0) = val;
}

// GroupNorm statistics, NHWC: grid (pixel chunks, N), blockDim = a multiple of the C/8 vectors of one pixel, so a thread
// always owns the same 8 channels and a warp reads whole pixels contiguously (16-byte vectors).  Deterministic: per-thread
// fp32 partials are folded per channel in a fixed order, per group in double, and every block writes its own slot of
// `partial` [N][chunks][G][2]; gn_coef_kernel adds the chunks in order.  HBM-bound: the tensor is read once.
// Rows per block: a function of the shape only (the partials are added in chunk order, so the result does not depend on the
// device), sized for ~1024 blocks per image.  With a fixed 2048 rows the 128^2 and 256^2 layers of the decoder ran on 8 and 32
// blocks: 86 us per launch for 16 MB, 17 launches per decode (profiles/r02_vae_launches.csv).
__host__ __device__ inline int gn_rows_per_block(int HW) {
  const int r = (HW + 1023) / 1024;
  return r < 64 ? 64 : (r + 63) / 64 * 64;
}
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, int HW, int C, int G,
                                                       double* __restrict__ partial) {
  extern __shared__ float sm[];                       // [blockDim][16] partials, then [2][C] channel sums
  const int n = blockIdx.y, nchunks = gridDim.x;
  const int nv = C >> 3;                              // vectors per pixel
  const int lanes = blockDim.x / nv;                  // pixels in flight per block step
  const int cv = threadIdx.x % nv;
  const int rpb = gn_rows_per_block(HW);
  const int r0 = blockIdx.x * rpb, r1 = min(r0 + rpb, HW);
  const uint4* xv = reinterpret_cast<const uint4*>(x + static_cast<long long>(n) * HW * C);
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  // four independent 16-byte loads in flight per thread (with one, the kernel ran at ~1.2 TB/s: profiles/r02_vae_launches.csv,
  // 187 us per launch against 93 us for the apply pass that moves twice the bytes); blockDim is a multiple of nv, so all of a
  // thread's vectors hold the same 8 channels
  const long long vend = static_cast<long long>(r1) * nv, bd = blockDim.x;
  long long v = static_cast<long long>(r0) * nv + threadIdx.x;
  auto acc = [&](const uint4 raw) {
    const float f[8] = {bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y), bf16lo(raw.z), bf16hi(raw.z), bf16lo(raw.w), bf16hi(raw.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] = fmaf(f[j], f[j], ss[j]); }
  };
  for (; v + 3 * bd < vend; v += 4 * bd) {
    const uint4 q0 = xv[v], q1 = xv[v + bd], q2 = xv[v + 2 * bd], q3 = xv[v + 3 * bd];
    acc(q0); acc(q1); acc(q2); acc(q3);
  }
  for (; v < vend; v += bd) acc(xv[v]);
  float* part = sm + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) { part[j] = s[j]; part[8 + j] = ss[j]; }
  __syncthreads();
  float* ch = sm + blockDim.x * 16;                   // [2][C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int v = c >> 3, j = c & 7;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < lanes; ++k) { a += sm[(v + k * nv) * 16 + j]; b += sm[(v + k * nv) * 16 + 8 + j]; }
    ch[c] = a; ch[C + c] = b;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0, b = 0;
    for (int c = 0; c < cpg; ++c) { a += ch[g * cpg + c]; b += ch[C + g * cpg + c]; }
    double* o = partial + ((static_cast<long long>(n) * nchunks + blockIdx.x) * G + g) * 2;
    o[0] = a; o[1] = b;
  }
  (void)cv;
}

// per (image, channel) affine of the normalisation: y = x * a + b with a = rstd * gamma, b = beta - mean * rstd * gamma.
// One WARP per (image, group): lanes take the chunks k = lane, lane + 32, ... in order, then a fixed shuffle tree -- the same
// summation order on every run (deterministic), and 32-way parallel (one thread per channel walking all chunks serially took
// 38 us per launch at 1024^2, 1.1 ms per decode).
__global__ void __launch_bounds__(256) gn_coef_kernel(const double* __restrict__ partial, int nchunks, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, int HW, int C, int G, int N,
                                                      float* __restrict__ coef /*[N][2][C]*/) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= N * G) return;
  const int n = w / G, g = w - n * G;
  const int cpg = C / G;
  double sa = 0, sb = 0;
  for (int k = lane; k < nchunks; k += 32) {
    const double* o = partial + ((static_cast<long long>(n) * nchunks + k) * G + g) * 2;
    sa += o[0]; sb += o[1];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, off);
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
  }
  const double cnt = static_cast<double>(HW) * cpg;
  const double m = sa / cnt;
  const double var = sb / cnt - m * m;
  const float rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-6f);
  for (int j = lane; j < cpg; j += 32) {
    const int c = g * cpg + j;
    const float a = rstd * gamma[c];
    coef[(static_cast<long long>(n) * 2) * C + c] = a;
    coef[(static_cast<long long>(n) * 2 + 1) * C + c] = beta[c] - static_cast<float>(m) * a;
  }
}

// y = [silu](x * a + b), HBM-bound (read + write once).  grid (blocks, N); blockDim is a multiple of the C/8 vectors of a pixel,
// so a thread keeps the same 8 channels through its grid-stride loop: its 16 coefficients are loaded once, there is no division
// in the loop, and four 16-byte loads are in flight per thread.  (One vector per step with a 64-bit div / mod and four
// coefficient loads each ran at 3.1 TB/s: 171 us for the 268 MB layers, profiles/r02_vae_launches.csv.)
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C,
                                                       const float* __restrict__ coef, int silu, long long per_img) {
  const int nv = C >> 3;
  const int n = blockIdx.y;
  const int c0 = (threadIdx.x % nv) * 8;
  const float* ca = coef + static_cast<long long>(n) * 2 * C + c0;
  const float4 a0 = *reinterpret_cast<const float4*>(ca), a1 = *reinterpret_cast<const float4*>(ca + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(ca + C), b1 = *reinterpret_cast<const float4*>(ca + C + 4);
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const uint4* xv = reinterpret_cast<const uint4*>(x) + n * per_img;
  uint4* yv = reinterpret_cast<uint4*>(y) + n * per_img;
  auto f = [&](const uint4 raw) {
    float v[8] = {bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y), bf16lo(raw.z), bf16hi(raw.z), bf16lo(raw.w), bf16hi(raw.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(v[j], av[j], bv[j]);
      if (silu) t = t / (1.0f + __expf(-t));
      v[j] = t;
    }
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  };
  const long long bd = static_cast<long long>(gridDim.x) * blockDim.x;      // a multiple of nv
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * bd < per_img; i += 4 * bd) {
    const uint4 q0 = xv[i], q1 = xv[i + bd], q2 = xv[i + 2 * bd], q3 = xv[i + 3 * bd];
    yv[i] = f(q0); yv[i + bd] = f(q1); yv[i + 2 * bd] = f(q2); yv[i + 3 * bd] = f(q3);
  }
  for (; i < per_img; i += bd) yv[i] = f(xv[i]);
}

// nearest-neighbour 2x upsampling, NHWC (Upsample2D [ext] before its 3x3 convolution): one 16-byte vector per thread
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int nv,
                                                         long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nv);
    long long p = i / nv;
    const int xo = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int yo = static_cast<int>(p % (2 * H));
    const long long n = p / (2 * H);
    y[i] = x[((n * H + (yo >> 1)) * W + (xo >> 1)) * nv + v];
  }
}

// one block per row: softmax over N fp32 scores -> bf16
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long lds, bf16* __restrict__ P, long ldp,
                                                           int N) {
  const float* s = S + static_cast<long long>(blockIdx.x) * lds;
  bf16* p = P + static_cast<long long>(blockIdx.x) * ldp;
  __shared__ float red[8];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, s[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sum += __expf(s[i] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < N; i += blockDim.x) p[i] = __float2bfloat16(__expf(s[i] - mx) * inv);
}

// The same for rows of up to 16 x 1024 scores (the mid block at 1024^2: 16 384) held in registers: ONE read of the row -- the
// three-pass kernel above re-read it for the sum and again for the output and computed every exponential twice (1.9 ms per
// decode at 0.85 TB/s, profiles/r02_vae_launches.csv) -- 16 B loads, 8 B stores.  Needs N, lds, ldp % 4 == 0 and 16 B-aligned S.
constexpr int SOFTMAX_NV = 16;
__global__ void __launch_bounds__(256, 2) softmax_rows_reg_kernel(const float* __restrict__ S, long lds, bf16* __restrict__ P,
                                                                  long ldp, int N) {
  const float4* s = reinterpret_cast<const float4*>(S + static_cast<long long>(blockIdx.x) * lds);
  uint2* p = reinterpret_cast<uint2*>(P + static_cast<long long>(blockIdx.x) * ldp);
  const int n4 = N >> 2;
  __shared__ float red[8];
  float4 v[SOFTMAX_NV];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < SOFTMAX_NV; ++j) {
    const int i = threadIdx.x + j * 256;
    v[j] = i < n4 ? s[i] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    mx = fmaxf(fmaxf(mx, fmaxf(v[j].x, v[j].y)), fmaxf(v[j].z, v[j].w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < SOFTMAX_NV; ++j) {
    v[j].x = __expf(v[j].x - mx); v[j].y = __expf(v[j].y - mx); v[j].z = __expf(v[j].z - mx); v[j].w = __expf(v[j].w - mx);
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < SOFTMAX_NV; ++j) {
    const int i = threadIdx.x + j * 256;
    if (i < n4) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[j].x * inv, v[j].y * inv), hi = __floats2bfloat162_rn(v[j].z * inv, v[j].w * inv);
      p[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
}

__global__ void __launch_bounds__(256) transpose_kernel(const bf16* __restrict__ x, long ldx, bf16* __restrict__ y, long ldy,
                                                        int R, int Cc) {
  __shared__ bf16 tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (by + j < R && bx + tx < Cc) tile[j][tx] = x[static_cast<long long>(by + j) * ldx + bx + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (bx + j < Cc && by + tx < R) y[static_cast<long long>(bx + j) * ldy + by + tx] = tile[tx][j];
}

__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, float v, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// layout changes at the two ends of the VAE (the reference's tensors are NCHW; the kernels in between run NHWC).  Small
// tensors (a 3-channel image, a 16-channel latent): one thread per element, reads or writes coalesced along W.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const bf16* __restrict__ x, int C, int HW, bf16* __restrict__ y,
                                                           long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // index into y: ((n*HW + p)*C + c)
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long np = i / C;
  const long long p = np % HW, n = np / HW;
  y[i] = x[(n * C + c) * HW + p];
}
template <typename OutT>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16* __restrict__ x, long ldx, int C, int HW, int clamp_from,
                                                           float lo, float hi, OutT* __restrict__ y, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // index into y: ((n*C + c)*HW + p)
  if (i >= total) return;
  const long long p = i % HW;
  const long long nc = i / HW;
  const int c = static_cast<int>(nc % C);
  const long long n = nc / C;
  float v = __bfloat162float(x[(n * HW + p) * ldx + c]);
  if (c >= clamp_from) v = fminf(fmaxf(v, lo), hi);
  if constexpr (sizeof(OutT) == 2) y[i] = __float2bfloat16(v);
  else y[i] = v;
}

}  // namespace

int fill_f32(float* p, float v, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  fill_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(p, v, n);
  UTX_CUDA(cudaGetLastError());
  return 0;
}
// x [N, C, H, W] -> y [N*H*W, C]
int nchw_to_nhwc_bf16(const bf16* x, int N, int C, int H, int W, bf16* y, cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * C * H * W;
  if (total == 0) return 0;
  nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, C, H * W, y, total);
  UTX_CUDA(cudaGetLastError());
  return 0;
}
// x [N*H*W, ldx] (first C columns) -> y [N, C, H, W]
int nhwc_to_nchw_bf16(const bf16* x, long ldx, int N, int C, int H, int W, bf16* y, cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * C * H * W;
  if (total == 0) return 0;
  nhwc_to_nchw_kernel<bf16><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, ldx, C, H * W, C, 0.f, 0.f, y, total);
  UTX_CUDA(cudaGetLastError());
  return 0;
}
// encoder head: x [N*H*W, ldx] (first C2 = 2*latent columns: mean | logvar) -> y [N, C2, H, W] fp32, logvar clamped to [lo, hi]
// (DiagonalGaussianDistribution [ext]: torch.clamp(logvar, -30, 20))
int moments_to_nchw_f32(const bf16* x, long ldx, int N, int C2, int H, int W, float lo, float hi, float* y, cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * C2 * H * W;
  if (total == 0) return 0;
  nhwc_to_nchw_kernel<float><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, ldx, C2, H * W, C2 / 2, lo, hi, y, total);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int im2col3x3(const bf16* x, int N, int Hin, int Win, int C, int up, int stride, int pad, int Ho, int Wo, int Kpad, bf16* out,
              cudaStream_t stream) {
  UTX_CHECK(Kpad % 8 == 0 && Kpad >= 9 * C, "im2col3x3: Kpad must be a multiple of 8 and >= 9*C");
  UTX_CHECK(up == 1 || up == 2, "im2col3x3: up must be 1 or 2");
  const long long total = static_cast<long long>(N) * Ho * Wo * (Kpad / 8);
  if (total == 0) return 0;
  im2col_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, N, Hin, Win, C, up, stride, pad, Ho, Wo,
                                                                                Kpad, out);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

size_t groupnorm_workspace_bytes(int N, int HW, int C, int G) {
  const size_t nchunks = (static_cast<size_t>(HW) + gn_rows_per_block(HW) - 1) / gn_rows_per_block(HW);
  return static_cast<size_t>(N) * nchunks * G * 2 * sizeof(double) + static_cast<size_t>(N) * 2 * C * sizeof(float);
}

int groupnorm_nhwc(const bf16* x, bf16* y, int N, int HW, int C, int G, const float* gamma, const float* beta, int silu,
                   double* stats_ws, cudaStream_t stream) {
  UTX_CHECK(C % G == 0 && C % 8 == 0 && C <= 2048, "groupnorm: C must be a multiple of G and of 8, <= 2048");
  if (N == 0 || HW == 0) return 0;
  const int nchunks = (HW + gn_rows_per_block(HW) - 1) / gn_rows_per_block(HW);
  // workspace (groupnorm_workspace_bytes): per-block moments [N][chunks][G][2] doubles, then [N][2][C] fp32 coefficients
  float* coef = reinterpret_cast<float*>(stats_ws + static_cast<size_t>(N) * nchunks * G * 2);
  const int nv = C / 8;
  const int threads = (256 / nv) * nv;
  dim3 grid(nchunks, N);
  gn_stats_kernel<<<grid, threads, (threads * 16 + 2 * C) * sizeof(float), stream>>>(x, HW, C, G, stats_ws);
  gn_coef_kernel<<<(N * G * 32 + 255) / 256, 256, 0, stream>>>(stats_ws, nchunks, gamma, beta, HW, C, G, N, coef);
  const long long per_img = static_cast<long long>(HW) * nv;
  long long blocks = (per_img + 4 * threads - 1) / (4 * threads);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  gn_apply_kernel<<<dim3(static_cast<unsigned>(blocks), N), threads, 0, stream>>>(x, y, C, coef, silu, per_img);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int upsample2x_nhwc(const bf16* x, int N, int H, int W, int C, bf16* y, cudaStream_t stream) {
  UTX_CHECK(C % 8 == 0, "upsample2x: C must be a multiple of 8");
  const long long total = static_cast<long long>(N) * 4 * H * W * (C / 8);
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  upsample2x_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), H, W,
                                                                       C / 8, total);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int softmax_rows(const float* S, long lds, bf16* P, long ldp, int M, int N, cudaStream_t stream) {
  if (M == 0) return 0;
  const bool vec = N % 4 == 0 && lds % 4 == 0 && ldp % 4 == 0 && N <= SOFTMAX_NV * 1024 && (reinterpret_cast<uintptr_t>(S) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(P) & 7) == 0;
  if (vec) softmax_rows_reg_kernel<<<M, 256, 0, stream>>>(S, lds, P, ldp, N);
  else softmax_rows_kernel<<<M, 256, 0, stream>>>(S, lds, P, ldp, N);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int transpose_bf16(const bf16* x, long ldx, bf16* y, long ldy, int R, int Cc, cudaStream_t stream) {
  if (R == 0 || Cc == 0) return 0;
  dim3 grid((Cc + 31) / 32, (R + 31) / 32);
  transpose_kernel<<<grid, 256, 0, stream>>>(x, ldx, y, ldy, R, Cc);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace utx
