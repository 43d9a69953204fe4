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

// statistics: grid (chunks, groups, N); each block reduces `rows_per_block` pixels of one group
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, int HW, int C, int G, int rows_per_block,
                                                       double* __restrict__ stats /*[N,G,2]*/) {
  const int g = blockIdx.y, n = blockIdx.z;
  const int cpg = C / G;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, HW);
  float s = 0.f, ss = 0.f;
  const long long base = static_cast<long long>(n) * HW * C + g * cpg;
  const int per_row = cpg;   // contiguous channels of the group within a pixel
  for (long long idx = static_cast<long long>(r0) * per_row + threadIdx.x; idx < static_cast<long long>(r1) * per_row; idx += blockDim.x) {
    const int r = static_cast<int>(idx / per_row), c = static_cast<int>(idx % per_row);
    const float v = __bfloat162float(x[base + static_cast<long long>(r) * C + c]);
    s += v;
    ss = fmaf(v, v, ss);
  }
  __shared__ float sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
    atomicAdd(stats + (static_cast<long long>(n) * G + g) * 2, a);
    atomicAdd(stats + (static_cast<long long>(n) * G + g) * 2 + 1, b);
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int HW, int C, int G,
                                                       const double* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int silu, long long total8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const long long e0 = i * 8;
  const int c0 = static_cast<int>(e0 % C);
  const int n = static_cast<int>(e0 / (static_cast<long long>(HW) * C));
  const int cpg = C / G;
  const uint4 raw = *reinterpret_cast<const uint4*>(x + e0);
  float f[8] = {bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y), bf16lo(raw.z), bf16hi(raw.z), bf16lo(raw.w), bf16hi(raw.w)};
  const double cnt = static_cast<double>(HW) * cpg;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j, g = c / cpg;
    const double m = stats[(static_cast<long long>(n) * G + g) * 2] / cnt;
    const double var = stats[(static_cast<long long>(n) * G + g) * 2 + 1] / cnt - m * m;
    const float rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-6f);
    float v = (f[j] - static_cast<float>(m)) * rstd * gamma[c] + beta[c];
    if (silu) v = v / (1.0f + __expf(-v));
    f[j] = v;
  }
  *reinterpret_cast<uint4*>(y + e0) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
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

}  // namespace

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

int groupnorm_nhwc(const bf16* x, bf16* y, int N, int HW, int C, int G, const float* gamma, const float* beta, int silu,
                   double* stats_ws, cudaStream_t stream) {
  UTX_CHECK(C % G == 0 && C % 8 == 0, "groupnorm: C must be a multiple of G and of 8");
  UTX_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * N * G * 2, stream));
  const int rows_per_block = 4096;
  dim3 grid((HW + rows_per_block - 1) / rows_per_block, G, N);
  gn_stats_kernel<<<grid, 256, 0, stream>>>(x, HW, C, G, rows_per_block, stats_ws);
  const long long total8 = static_cast<long long>(N) * HW * C / 8;
  gn_apply_kernel<<<static_cast<unsigned>((total8 + 255) / 256), 256, 0, stream>>>(x, y, HW, C, G, stats_ws, gamma, beta, silu,
                                                                                  total8);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int softmax_rows(const float* S, long lds, bf16* P, long ldp, int M, int N, cudaStream_t stream) {
  if (M == 0) return 0;
  softmax_rows_kernel<<<M, 256, 0, stream>>>(S, lds, P, ldp, N);
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
