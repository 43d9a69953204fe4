// Internal C++ launch API shared by the kernels, the FLUX engine and the C ABI (capi.cu).
// Every launcher returns 0 on success (non-zero + utx::get_error() otherwise) and only enqueues
// work on `stream`.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace utx {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ tcgen05 GEMM  C = epi(A @ W^T + bias)
enum GemmEpilogue : int {
  EPI_BIAS = 0,        // C = acc + bias
  EPI_BIAS_GELU = 1,   // C = gelu_tanh(acc + bias) for columns >= gelu_col_start, acc + bias below
  EPI_GATE_RES = 2,    // C = res + gate[n] * (acc + bias)        (res may alias C)
};
struct GemmProblem {
  const bf16* A;   // [M, K] row-major, row stride lda
  long lda;
  const bf16* W;   // [N, K] row-major (nn.Linear weight), row stride ldw
  long ldw;
  int M;
  bf16* C;         // [M, N] row stride ldc
  long ldc;
  const bf16* bias;   // [N] or nullptr
  const float* gate;  // [N] fp32 (EPI_GATE_RES)
  const bf16* res;    // [M, N] row stride ldres (EPI_GATE_RES)
  long ldres;
  // optional column split: columns >= split_col go to C2[:, col - split_col] (split_col % 256 == 0)
  int split_col;      // 0 = no split
  bf16* C2;
  long ldc2;
};
struct GemmArgs {
  int N, K;
  int epi;
  int gelu_col_start;
  int nprob;             // 1 or 2 problems sharing N, K and the epilogue (txt + img streams)
  GemmProblem prob[2];
};
int gemm_bf16_tn(const GemmArgs& args, cudaStream_t stream);

// ------------------------------------------------------------------ fused joint attention (tcgen05 flash attention)
// qkv: [S, 3*H*128] bf16 (q | k | v, head-major inside each third), already RMS-normed + RoPE'd.
// out: [S, ld_out] bf16, head h written at columns [h*128, h*128+128).  softmax(q k^T / sqrt(128)) v, no mask.
int attention_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream);

// ------------------------------------------------------------------ HBM-bound elementwise / reduction kernels
// y[r,:] = LayerNorm(x[r,:], eps=1e-6, no affine) * (1 + scale) + shift; rows < rows0 use (shift0,scale0), others (shift1,scale1)
int ln_modulate(const bf16* x, long ldx, bf16* y, long ldy, int rows, int D, int rows0, const float* shift0,
                const float* scale0, const float* shift1, const float* scale1, cudaStream_t stream);
// in-place per-head RMSNorm(eps 1e-6, weight) + RoPE on the q and k thirds of qkv [S, 3*H*128].
// rows < rows0 use (wq0, wk0) [txt stream], the rest (wq1, wk1).  cos/sin: [S,128] fp32.
int rmsnorm_rope(bf16* qkv, long ld_qkv, int S, int H, int rows0, const bf16* wq0, const bf16* wk0, const bf16* wq1,
                 const bf16* wk1, const float* cos_t, const float* sin_t, cudaStream_t stream);
// y[n] = sum_k W[n,k] * f(x[k]) + b[n]   (f = silu if silu_in);  W bf16 [N,K], x fp32 [K], y fp32 [N]; y += if accumulate
int gemv_bf16(const bf16* W, const bf16* b, const float* x, float* y, int N, int K, int silu_in, int accumulate,
              cudaStream_t stream);
// out[0:256] = sinusoid(1000 * t) ; out[256:512] = sinusoid(1000 * g)   (flip_sin_to_cos, fp32)
int time_sinusoid(float t_scaled, float g_scaled, float* out512, cudaStream_t stream);
// cos/sin table [S,128] fp32 from ids [S,3] fp32 with axes (16,56,56), theta 1e4, angles in fp64
int rope_table(const float* ids, int S, float* cos_t, float* sin_t, cudaStream_t stream);
// latents[r,c] = bf16( float(latents[r,c]) + dsigma * float(v[r,c]) ) for r < rows
int euler_update(bf16* latents, const bf16* v, int rows, int cols, float dsigma, cudaStream_t stream);
// W[o,i] += scale * sum_r B[o,r] * A[r,i]   (fp32 math, bf16 storage; LoRA merge)
int lora_merge(bf16* W, long ldw, const float* A, const float* B, int out_f, int in_f, int rank, float scale,
               cudaStream_t stream);

}  // namespace utx
