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
  EPI_BIAS_F32 = 3,    // C (float*, ldc in floats) = out_scale * (acc + bias)
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
  // fused per-head RMSNorm + RoPE on the q|k columns (GemmArgs::qk_cols > 0): this problem's norm weights [128] and the
  // token index of its row 0 in the cos/sin tables
  const bf16* wq;
  const bf16* wk;
  int row_offset;
  // sequence-parallel head scatter of the q | k | v columns into the all-to-all send layout (gemm_epilogue.cuh); sc_hl = 0: off
  int sc_hl, sc_rows, sc_row_base, sc_D;
  bf16* sc_peer[8];   // direct mode: every rank's attention input mapped over NVLink (nullptr: send layout in C)
};
struct GemmArgs {
  int N, K;
  int epi;
  int gelu_col_start;
  float out_scale;       // EPI_BIAS_F32 only
  int qk_cols;           // 0 = off; 2*D: columns [0,D) are q heads, [D,2D) k heads (128 per head) -> RMSNorm + RoPE in the epilogue
  const float* cos_t;    // [S,128] fp32 tables for qk_cols > 0
  const float* sin_t;
  int nprob;             // 1 or 2 problems sharing N, K and the epilogue (txt + img streams)
  // implicit-GEMM 3x3 convolution (stride 1, pad 1): prob[0].A is the NHWC activation [conv_n, conv_h, conv_w, conv_c],
  // M = conv_n*conv_h*conv_w output pixels, K = 9*conv_c with W arranged [Cout][ky][kx][Cin]; 0 = plain GEMM
  int conv_n, conv_h, conv_w, conv_c;
  GemmProblem prob[2];
};
int gemm_bf16_tn(const GemmArgs& args, cudaStream_t stream);

// ------------------------------------------------------------------ fused joint attention (tcgen05 flash attention)
// qkv: [S, 3*H*128] bf16 (q | k | v, head-major inside each third), already RMS-normed + RoPE'd.
// out: [S, ld_out] bf16, head h written at columns [h*128, h*128+128).  softmax(q k^T / sqrt(128)) v, no mask.
// scatter != nullptr (sequence-parallel direct mode): output row r goes to rank r / rows_per_rank's buffer over NVLink,
// base[owner][(r % rows_per_rank) * ld + col0 + head * 128 ...] instead of `out`
struct AttnScatter {
  bf16* base[8];
  int rows_per_rank;
  long ld;
  int col0;
};
int attention_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream,
                   const AttnScatter* scatter = nullptr);

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
// the same with (t, g) read from device memory tg_dev[0..1]; set_step_scalars writes (t, g, dsigma) to dev3[0..2] on `stream`
int time_sinusoid_dev(const float* tg_dev, float* out512, cudaStream_t stream);
int set_step_scalars(float* dev3, float t_scaled, float g_scaled, float dsigma, cudaStream_t stream);
// cos/sin table [S,128] fp32 from ids [S,3] fp32 with axes (16,56,56), theta 1e4, angles in fp64
int rope_table(const float* ids, int S, float* cos_t, float* sin_t, cudaStream_t stream);
// latents[r,c] = bf16( float(latents[r,c]) + dsigma * float(v[r,c]) ) for r < rows
// dsigma_dev != nullptr: the sigma difference is read from device memory instead (CUDA-graph replay)
int euler_update(bf16* latents, const bf16* v, int rows, int cols, float dsigma, cudaStream_t stream,
                 const float* dsigma_dev = nullptr);
// W[o,i] += scale * sum_r B[o,r] * A[r,i]   (fp32 math, bf16 storage; LoRA merge)
int lora_merge(bf16* W, long ldw, const float* A, const float* B, int out_f, int in_f, int rank, float scale,
               cudaStream_t stream);


// ------------------------------------------------------------------ bake: rasterise / interpolate / LBVH / fused UV bake
size_t rasterize_workspace_bytes(int B, int H, int W, int F);
// pos [B or 1, V, 4] clip space, tri [F,3] -> rast [B,H,W,4] = (u, v, z/w, id+1)
int rasterize(const float* pos, int pos_batched, int V, const int* tri, int F, int B, int H, int W, float* rast_out,
              void* workspace, cudaStream_t stream);
int interpolate(const float* attr, int attr_batched, int V, int C, const float* rast, const int* tri, int B, int H, int W,
                float* out, cudaStream_t stream);
// mv_to_pcd(filt_gradient_points=True) (bake_filter.cu): attrs [n,H,W,6] = interpolated (position, vertex normal), view_dirs [n,3]
// device (ray direction, or camera position when perspective) -> mask_vis u8 [n,H,W]
int mv_visibility_filter(const float* attrs, const float* rast, const float* face_normals, const float* view_dirs, int perspective,
                         int n, int H, int W, float grad_thr, float cos_thr, unsigned char* mask_vis, cudaStream_t stream);
// kdtree_method='mvpaint' blend of a [M,k] neighbour table (bake_filter.cu) -> out [M,3]
int mvpaint_blend(const float* score, const long long* index, long long M, int k, const float* cloud_c, const float* cloud_n,
                  const float* tex_n, float* out, cudaStream_t stream);
// out[b, v, :] = mats[b] (row-major 4x4) @ [vert[v], 1]
int transform_points(const float* vert, int V, const float* mats, int n, float* out, cudaStream_t stream);
size_t bvh_nodes_bytes(int F);
size_t bvh_workspace_bytes(int F);
int bvh_build(const float* vert, int V, const int* tri, int F, void* nodes_out, void* workspace, size_t ws_bytes,
              cudaStream_t stream);
struct PointTree;   // bake_trace.cuh
int point_bvh_build(const float* pts, const int* ids, int n, void* nodes_out, void* workspace, size_t ws_bytes,
                    cudaStream_t stream);
PointTree point_tree_view(const void* nodes, int n);
int knn1(const float* src, int n_src, const float* dst, long long M, long long* index, float* score, void* nodes, void* workspace,
         size_t ws_bytes, cudaStream_t stream);
int bvh_export(const void* nodes, int F, int* info, float* aabb, cudaStream_t stream);
int bvh_intersect(const void* nodes, const float* vert, const int* tri, int F, const float* rays_o, const float* rays_d,
                  long long N, unsigned char* hit, int* tid, float* pos, float* uv, cudaStream_t stream);
int knn(const float* src, int n_src, const float* dst, long long M, int k, long long* index, float* score, void* nodes,
        void* workspace, size_t ws_bytes, cudaStream_t stream);
size_t uv_bake_workspace_bytes(int H2, int W2);
// staged form of uv_bake over one workspace (bake_uv.cu): visibility -> [views_knn] -> fill -> finish
void uv_bake_layout(int H2, int W2, size_t* off_owner, size_t* off_pos, size_t* off_color, size_t* off_seam);
int uv_bake_visibility(const float* vert, int V, const int* tri, int F, const void* nodes, const float* rast2d, int H2, int W2,
                       int n_views, const float* view_mats_host, const float* view_dirs_host, int perspective, const int* priority_host,
                       const float* images_rgba, int H, int W, float cos_thresh, unsigned char* mask2d,
                       unsigned char* mask_vis, void* workspace, size_t ws_bytes, cudaStream_t stream);
int uv_bake_fill(const unsigned char* mask2d, int H2, int W2, int k, int* nn_index_out, void* workspace, size_t ws_bytes,
                 cudaStream_t stream);
size_t uv_bake_views_workspace_bytes(int n_views, int H, int W);
int uv_bake_views_knn(const float* pix_pos, const float* images_rgba, int n_views, int H, int W, int k, int merge,
                      const unsigned char* mask2d, int H2, int W2, void* workspace, size_t ws_bytes, void* scratch,
                      size_t scratch_bytes, cudaStream_t stream);
int uv_bake_finish(const unsigned char* mask2d, int H2, int W2, int blur, const float* blur_k2d, float blur_gamma,
                   float* color_out, void* workspace, size_t ws_bytes, cudaStream_t stream);
int uv_bake(const float* vert, int V, const int* tri, int F, const void* nodes, const float* rast2d, int H2, int W2,
            int n_views, const float* view_mats_host, const float* view_dirs_host, int perspective, const int* priority_host,
            const float* images_rgba, int H, int W, float cos_thresh, const float* blur_k2d, float blur_gamma,
            const float* grid_lo_host, float grid_extent, unsigned char* mask2d, unsigned char* mask_vis, float* color_out,
            int* nn_index_out, void* workspace, size_t ws_bytes, cudaStream_t stream);


// ------------------------------------------------------------------ VAE (NHWC bf16; convs = im2col + gemm_bf16_tn)
// y[N*H*W, Cout] = conv3x3(x NHWC [N,H,W,C], w [Cout][3][3][C]) + bias, or gate * (conv + bias) + res: no im2col buffer, the A tiles are TMA boxes of
// the activation shifted by the tap, zero padding comes from TMA's out-of-bounds fill.  C % 64 == 0; W % 128 == 0 or
// (128 % W == 0 and H % (128 / W) == 0)
int conv3x3_nhwc(const bf16* x, int N, int H, int W, int C, const bf16* w, const bf16* bias, int Cout, bf16* y, long ldy,
                 const float* gate, const bf16* res, long ldres, cudaStream_t stream);
int im2col3x3(const bf16* x, int N, int Hin, int Win, int C, int up, int stride, int pad, int Ho, int Wo, int Kpad, bf16* out,
              cudaStream_t stream);
int upsample2x_nhwc(const bf16* x, int N, int H, int W, int C, bf16* y, cudaStream_t stream);
size_t groupnorm_workspace_bytes(int N, int HW, int C, int G);
int groupnorm_nhwc(const bf16* x, bf16* y, int N, int HW, int C, int G, const float* gamma, const float* beta, int silu,
                   double* stats_ws, cudaStream_t stream);
int softmax_rows(const float* S, long lds, bf16* P, long ldp, int M, int N, cudaStream_t stream);
int transpose_bf16(const bf16* x, long ldx, bf16* y, long ldy, int R, int Cc, cudaStream_t stream);

}  // namespace utx
