// HBM-bound kernels of the DiT step: adaLN LayerNorm+modulate, per-head RMSNorm+RoPE, modulation GEMV,
// timestep sinusoid, RoPE table, Euler update, LoRA merge.  All: 16-byte vector accesses, one warp per row
// (or per output), fp32 math, one rounding to bf16 at the store.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// ----------------------------------------------------------------------------- LayerNorm + modulate
// AdaLayerNormZero / ZeroSingle / Continuous [ext diffusers]: LN(eps 1e-6, no affine) * (1 + scale) + shift.
// One warp per row; the row (NV x 256 elements) lives in registers between the statistics and the output pass.
// Register cap: left alone, ptxas hoists the modulation-vector loads of the whole output pass and the D = 3072 instance needs 179
// registers -- ONE resident block of 8 rows per SM (37.9 us per launch = 3.2 TB/s, profiles/r01_launches_final.csv).  Capped at
// 128 (two resident blocks, 16 rows = 96 KB in flight per SM; 85 for three blocks spills the row).
template <int NV>
__global__ void __launch_bounds__(256, (NV <= 6 ? 3 : NV <= 12 ? 2 : 1)) ln_modulate_kernel(const bf16* __restrict__ x, long ldx, bf16* __restrict__ y,
                                                          long ldy, int rows, int D, int rows0,
                                                          const float* __restrict__ shift0,
                                                          const float* __restrict__ scale0,
                                                          const float* __restrict__ shift1,
                                                          const float* __restrict__ scale1) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long>(row) * ldx);
  uint4 raw[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) raw[i] = xr[i * 32 + lane];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(raw[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += f[j];
  }
  const float mean = warp_sum(sum) / D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(raw[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - mean;
      var = fmaf(d, d, var);
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / D + 1e-6f);
  const float* sh = row < rows0 ? shift0 : shift1;
  const float* sc = row < rows0 ? scale0 : scale1;
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<long>(row) * ldy);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    float f[8];
    unpack8(raw[i], f);
    const float4 s0 = *reinterpret_cast<const float4*>(sc + col), s1 = *reinterpret_cast<const float4*>(sc + col + 4);
    const float4 h0 = *reinterpret_cast<const float4*>(sh + col), h1 = *reinterpret_cast<const float4*>(sh + col + 4);
    const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf((f[j] - mean) * rstd, 1.0f + scv[j], shv[j]);
    yr[i * 32 + lane] = pack8(f);
  }
}

// ----------------------------------------------------------------------------- per-head RMSNorm + RoPE (in place)
// attention_processor.py:56-59,73-76 (norm_q/norm_k/norm_added_*) and :85-87 (apply_rotary_emb).
// One warp per (token, q|k, pair of heads): 256 contiguous elements, one 16-byte load/store per lane; each half-warp
// owns one head (RMS reduction over 16 lanes), each lane four rotation pairs.
__global__ void __launch_bounds__(256) rmsnorm_rope_kernel(bf16* __restrict__ qkv, long ld, int S, int H, int rows0,
                                                           const bf16* __restrict__ wq0, const bf16* __restrict__ wk0,
                                                           const bf16* __restrict__ wq1, const bf16* __restrict__ wk1,
                                                           const float* __restrict__ cos_t,
                                                           const float* __restrict__ sin_t) {
  const long gw = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int HP = H >> 1;
  const long total = static_cast<long>(S) * HP * 2;
  if (gw >= total) return;
  const int lane = threadIdx.x & 31;
  const int which = static_cast<int>(gw % 2);          // 0 = q, 1 = k
  const int hp = static_cast<int>((gw / 2) % HP);
  const int tok = static_cast<int>(gw / (2L * HP));
  const int e0 = (lane & 15) * 8;                      // element offset inside the head
  bf16* p = qkv + static_cast<long>(tok) * ld + static_cast<long>(which) * H * 128 + hp * 256 + lane * 8;
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  float f[8];
  unpack8(raw, f);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rs = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
  const bf16* w = tok < rows0 ? (which ? wk0 : wq0) : (which ? wk1 : wq1);
  float wv[8];
  unpack8(*reinterpret_cast<const uint4*>(w + e0), wv);
  const float* ct = cos_t + static_cast<long>(tok) * 128 + e0;
  const float* st = sin_t + static_cast<long>(tok) * 128 + e0;
  const float4 c0 = *reinterpret_cast<const float4*>(ct), c1 = *reinterpret_cast<const float4*>(ct + 4);
  const float4 s0 = *reinterpret_cast<const float4*>(st), s1 = *reinterpret_cast<const float4*>(st + 4);
  const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = f[j] * rs * wv[j];
  float o[8];
  // out[2i] = x[2i] cos - x[2i+1] sin ; out[2i+1] = x[2i+1] cos + x[2i] sin
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    o[j] = f[j] * cv[j] - f[j + 1] * sv[j];
    o[j + 1] = f[j + 1] * cv[j + 1] + f[j] * sv[j + 1];
  }
  *reinterpret_cast<uint4*>(p) = pack8(o);
}

// ----------------------------------------------------------------------------- GEMV  y = W f(x) + b
// All adaLN modulation vectors of a step (one [N_mod, D] weight) and the time/guidance/pooled MLPs.
// One warp per output row, weights streamed once with 16B loads; HBM-bound by construction (2 bytes / MAC).
__global__ void __launch_bounds__(256) gemv_kernel(const bf16* __restrict__ W, const bf16* __restrict__ b,
                                                   const float* __restrict__ x, float* __restrict__ y, int N, int K,
                                                   int silu_in, int accumulate) {
  extern __shared__ float xs[];
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const float v = x[i];
    xs[i] = silu_in ? v / (1.0f + __expf(-v)) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += gridDim.x * wpb) {
    const uint4* wr = reinterpret_cast<const uint4*>(W + static_cast<long>(n) * K);
    float acc = 0.f;
    for (int i = lane; i < K / 8; i += 32) {
      const uint4 u = wr[i];
      float f[8];
      unpack8(u, f);
      const float4 x0 = *reinterpret_cast<const float4*>(xs + i * 8), x1 = *reinterpret_cast<const float4*>(xs + i * 8 + 4);
      acc = fmaf(f[0], x0.x, acc); acc = fmaf(f[1], x0.y, acc); acc = fmaf(f[2], x0.z, acc); acc = fmaf(f[3], x0.w, acc);
      acc = fmaf(f[4], x1.x, acc); acc = fmaf(f[5], x1.y, acc); acc = fmaf(f[6], x1.z, acc); acc = fmaf(f[7], x1.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (b) acc += __bfloat162float(b[n]);
      y[n] = accumulate ? y[n] + acc : acc;
    }
  }
}

// ----------------------------------------------------------------------------- timestep / guidance sinusoid
// get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0) [ext]: [cos(t f_i) | sin(t f_i)].
__global__ void time_sinusoid_kernel(float t, float g, float* __restrict__ out) {
  const int i = threadIdx.x;   // 0..127
  const float freq = expf(-9.210340371976184f * static_cast<float>(i) / 128.0f);
  out[i] = cosf(t * freq);
  out[128 + i] = sinf(t * freq);
  out[256 + i] = cosf(g * freq);
  out[384 + i] = sinf(g * freq);
}

// the same with the two scalars read from device memory (CUDA-graph replay of the denoise step: the graph is captured once,
// the per-step scalars are written by set_step_scalars just before every launch)
__global__ void time_sinusoid_dev_kernel(const float* __restrict__ tg, float* __restrict__ out) {
  const int i = threadIdx.x;   // 0..127
  const float t = tg[0], g = tg[1];
  const float freq = expf(-9.210340371976184f * static_cast<float>(i) / 128.0f);
  out[i] = cosf(t * freq);
  out[128 + i] = sinf(t * freq);
  out[256 + i] = cosf(g * freq);
  out[384 + i] = sinf(g * freq);
}
__global__ void set_step_scalars_kernel(float* __restrict__ p, float t, float g, float dsigma) {
  p[0] = t;
  p[1] = g;
  p[2] = dsigma;
}

// ----------------------------------------------------------------------------- RoPE table (FluxPosEmbed [ext])
__global__ void rope_table_kernel(const float* __restrict__ ids, int S, float* __restrict__ cos_t,
                                  float* __restrict__ sin_t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (token, pair) with 64 pairs per token
  if (idx >= S * 64) return;
  const int tok = idx >> 6, pr = idx & 63;
  int axis, i, d;
  if (pr < 8) { axis = 0; i = pr; d = 16; }
  else if (pr < 36) { axis = 1; i = pr - 8; d = 56; }
  else { axis = 2; i = pr - 36; d = 56; }
  const double freq = 1.0 / pow(10000.0, static_cast<double>(2 * i) / d);
  const double ang = static_cast<double>(ids[tok * 3 + axis]) * freq;
  const float c = static_cast<float>(cos(ang)), s = static_cast<float>(sin(ang));
  cos_t[tok * 128 + 2 * pr] = c;
  cos_t[tok * 128 + 2 * pr + 1] = c;
  sin_t[tok * 128 + 2 * pr] = s;
  sin_t[tok * 128 + 2 * pr + 1] = s;
}

// ----------------------------------------------------------------------------- Euler update (a6)
// FlowMatchEulerDiscreteScheduler.step [ext]: `sample.float() + (sigma' - sigma) * model_output`, cast back to bf16.
// The sigmas are 0-dim fp32 tensors there, so torch forms the product in model_output's dtype: the increment is rounded
// to bf16 BEFORE the fp32 add.  Mirrored (pinned by tests/golden/ref_flux_call.npz, generated by the reference's own
// __call__): torch casts BOTH operands of that product to bf16 (the 0-dim sigma difference included), multiplies in fp32
// and rounds: bf16(dsigma) -> mul -> bf16 RN -> fp32 add -> bf16 RN, no fused multiply-add.
__global__ void euler_kernel(bf16* __restrict__ lat, const bf16* __restrict__ v, long n8, float dsigma,
                             const float* __restrict__ dsigma_dev) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  if (dsigma_dev) dsigma = *dsigma_dev;     // graph replay: the step's sigma difference lives in device memory
  float a[8], b[8];
  unpack8(reinterpret_cast<const uint4*>(lat)[i], a);
  unpack8(reinterpret_cast<const uint4*>(v)[i], b);
  const float d = __bfloat162float(__float2bfloat16_rn(dsigma));
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = __fadd_rn(a[j], __bfloat162float(__float2bfloat16_rn(__fmul_rn(d, b[j]))));
  reinterpret_cast<uint4*>(lat)[i] = pack8(a);
}

// ----------------------------------------------------------------------------- LoRA merge (a5)
// W[o,i] += scale * sum_r B[o,r] A[r,i].  32x32 output tile per block, rank walked in chunks of 32 through smem.
__global__ void __launch_bounds__(256) lora_merge_kernel(bf16* __restrict__ W, long ldw, const float* __restrict__ A,
                                                         const float* __restrict__ B, int out_f, int in_f, int rank,
                                                         float scale) {
  __shared__ float sA[32][33], sB[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = 0; r0 < rank; r0 += 32) {
    for (int k = ty; k < 32; k += 8) {
      const int r = r0 + k;
      sA[k][tx] = (r < rank && i0 + tx < in_f) ? A[static_cast<long>(r) * in_f + i0 + tx] : 0.f;
      const int o = o0 + k;
      sB[k][tx] = (o < out_f && r0 + tx < rank) ? B[static_cast<long>(o) * rank + r0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ol = ty + q * 8;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc[q] = fmaf(sB[ol][k], sA[k][tx], acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int o = o0 + ty + q * 8, i = i0 + tx;
    if (o < out_f && i < in_f) {
      bf16* w = W + static_cast<long>(o) * ldw + i;
      *w = __float2bfloat16(__bfloat162float(*w) + scale * acc[q]);
    }
  }
}

}  // namespace

int ln_modulate(const bf16* x, long ldx, bf16* y, long ldy, int rows, int D, int rows0, const float* shift0,
                const float* scale0, const float* shift1, const float* scale1, cudaStream_t stream) {
  UTX_CHECK(D % 256 == 0 && D <= 256 * 16, "ln_modulate: D must be a multiple of 256, <= 4096");
  UTX_CHECK(ldx % 8 == 0 && ldy % 8 == 0, "ln_modulate: leading dims must be multiples of 8");
  if (rows == 0) return 0;
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb), block(32 * wpb);
#define LN_CASE(NV)                                                                                       \
  case NV:                                                                                                \
    ln_modulate_kernel<NV><<<grid, block, 0, stream>>>(x, ldx, y, ldy, rows, D, rows0, shift0, scale0,    \
                                                       shift1, scale1);                                   \
    break;
  switch (D / 256) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(6) LN_CASE(8) LN_CASE(12) LN_CASE(16)
    default:
      UTX_CHECK(false, "ln_modulate: unsupported D/256");
  }
#undef LN_CASE
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int rmsnorm_rope(bf16* qkv, long ld_qkv, int S, int H, int rows0, const bf16* wq0, const bf16* wk0, const bf16* wq1,
                 const bf16* wk1, const float* cos_t, const float* sin_t, cudaStream_t stream) {
  if (S == 0) return 0;
  UTX_CHECK(ld_qkv % 8 == 0 && H % 2 == 0, "rmsnorm_rope: ld must be a multiple of 8 and H even");
  const long warps = static_cast<long>(S) * H;
  const int wpb = 8;
  rmsnorm_rope_kernel<<<static_cast<unsigned>((warps + wpb - 1) / wpb), 32 * wpb, 0, stream>>>(
      qkv, ld_qkv, S, H, rows0, wq0, wk0, wq1, wk1, cos_t, sin_t);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int gemv_bf16(const bf16* W, const bf16* b, const float* x, float* y, int N, int K, int silu_in, int accumulate,
              cudaStream_t stream) {
  UTX_CHECK(K % 8 == 0 && K * 4 <= 48 * 1024, "gemv: K must be a multiple of 8 and <= 12288");
  if (N == 0) return 0;
  const int wpb = 8;
  int grid = (N + wpb - 1) / wpb;
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  gemv_kernel<<<grid, 32 * wpb, K * sizeof(float), stream>>>(W, b, x, y, N, K, silu_in, accumulate);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// sequence-parallel attention tail: recv [P][rows][w] (peer p's heads for this rank's rows) -> cat[row, p * w + j]
__global__ void __launch_bounds__(256) sp_unpack_kernel(const uint4* __restrict__ recv, bf16* __restrict__ cat, long ld_cat, int rows,
                                                        int w8, long total) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % w8);
    const long pr = i / w8;
    const int row = static_cast<int>(pr % rows), p = static_cast<int>(pr / rows);
    *reinterpret_cast<uint4*>(cat + row * ld_cat + (static_cast<long>(p) * w8 + j) * 8) = recv[i];
  }
}
int sp_unpack_heads(const bf16* recv, bf16* cat, long ld_cat, int rows, int w, int npeers, cudaStream_t stream) {
  UTX_CHECK(w % 8 == 0 && ld_cat % 8 == 0, "sp_unpack_heads: widths must be multiples of 8");
  const long total = static_cast<long>(npeers) * rows * (w / 8);
  if (total == 0) return 0;
  long blocks = (total + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  sp_unpack_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const uint4*>(recv), cat, ld_cat, rows, w / 8, total);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int time_sinusoid(float t_scaled, float g_scaled, float* out512, cudaStream_t stream) {
  time_sinusoid_kernel<<<1, 128, 0, stream>>>(t_scaled, g_scaled, out512);
  UTX_CUDA(cudaGetLastError());
  return 0;
}
int time_sinusoid_dev(const float* tg_dev, float* out512, cudaStream_t stream) {
  time_sinusoid_dev_kernel<<<1, 128, 0, stream>>>(tg_dev, out512);
  UTX_CUDA(cudaGetLastError());
  return 0;
}
int set_step_scalars(float* dev3, float t_scaled, float g_scaled, float dsigma, cudaStream_t stream) {
  set_step_scalars_kernel<<<1, 1, 0, stream>>>(dev3, t_scaled, g_scaled, dsigma);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int rope_table(const float* ids, int S, float* cos_t, float* sin_t, cudaStream_t stream) {
  if (S == 0) return 0;
  rope_table_kernel<<<(S * 64 + 255) / 256, 256, 0, stream>>>(ids, S, cos_t, sin_t);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int euler_update(bf16* latents, const bf16* v, int rows, int cols, float dsigma, cudaStream_t stream, const float* dsigma_dev) {
  const long n = static_cast<long>(rows) * cols;
  UTX_CHECK(n % 8 == 0, "euler_update: element count must be a multiple of 8");
  if (n == 0) return 0;
  const long n8 = n / 8;
  euler_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(latents, v, n8, dsigma, dsigma_dev);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

int lora_merge(bf16* W, long ldw, const float* A, const float* B, int out_f, int in_f, int rank, float scale,
               cudaStream_t stream) {
  if (out_f == 0 || in_f == 0) return 0;
  dim3 grid((in_f + 31) / 32, (out_f + 31) / 32);
  lora_merge_kernel<<<grid, 256, 0, stream>>>(W, ldw, A, B, out_f, in_f, rank, scale);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace utx
