// extern "C" surface for the building-block launchers (include/unitex_b200.h).  No torch types cross this boundary.
#include "../../include/unitex_b200.h"
#include "common.h"
#include "kernels.h"

using namespace utx;

extern "C" {

const char* utx_last_error(void) { return get_error(); }
int utx_version(void) { return 100; }

int utx_gemm_bf16(const void* A, long lda, const void* W, long ldw, const void* bias, void* C, long ldc, int M, int N,
                  int K, int epi, const float* gate, const void* res, long ldres, void* stream) {
  UTX_CHECK(M >= 0 && N > 0 && K > 0, "utx_gemm_bf16: bad shape");
  if (M == 0) return 0;   // empty batch (torch hands out null data pointers for empty tensors): nothing to launch
  UTX_CHECK(A && W && C, "utx_gemm_bf16: null pointer");
  GemmArgs a{};
  a.N = N; a.K = K; a.epi = epi; a.gelu_col_start = 0; a.nprob = 1;
  a.prob[0] = GemmProblem{static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, M, static_cast<bf16*>(C),
                          ldc, static_cast<const bf16*>(bias), gate, static_cast<const bf16*>(res), ldres, 0, nullptr, 0};
  return gemm_bf16_tn(a, static_cast<cudaStream_t>(stream));
}

int utx_gemm_bf16_qkv(const void* A, long lda, const void* W, const void* bias, void* C, long ldc, int M, int heads, int K,
                      const void* wq, const void* wk, const float* cos_t, const float* sin_t, int row_offset, void* stream) {
  UTX_CHECK(A && W && C && wq && wk && cos_t && sin_t, "utx_gemm_bf16_qkv: null pointer");
  GemmArgs a{};
  a.N = 3 * heads * 128; a.K = K; a.epi = EPI_BIAS; a.nprob = 1;
  a.qk_cols = 2 * heads * 128; a.cos_t = cos_t; a.sin_t = sin_t;
  a.prob[0] = GemmProblem{static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), K, M, static_cast<bf16*>(C), ldc,
                          static_cast<const bf16*>(bias), nullptr, nullptr, 0, 0, nullptr, 0, static_cast<const bf16*>(wq),
                          static_cast<const bf16*>(wk), row_offset};
  return gemm_bf16_tn(a, static_cast<cudaStream_t>(stream));
}

int utx_gemm_bf16_grouped2(const void* A0, long lda0, const void* W0, const void* bias0, void* C0, long ldc0, int M0,
                           const void* A1, long lda1, const void* W1, const void* bias1, void* C1, long ldc1, int M1,
                           int N, int K, int epi, const float* gate0, const float* gate1, void* stream) {
  UTX_CHECK(A0 && W0 && C0 && A1 && W1 && C1, "utx_gemm_bf16_grouped2: null pointer");
  GemmArgs a{};
  a.N = N; a.K = K; a.epi = epi; a.gelu_col_start = 0; a.nprob = 2;
  // EPI_GATE_RES in grouped form is residual-in-place (res == C), as the engine uses it
  a.prob[0] = GemmProblem{static_cast<const bf16*>(A0), lda0, static_cast<const bf16*>(W0), K, M0,
                          static_cast<bf16*>(C0), ldc0, static_cast<const bf16*>(bias0), gate0,
                          epi == EPI_GATE_RES ? static_cast<const bf16*>(C0) : nullptr, ldc0, 0, nullptr, 0};
  a.prob[1] = GemmProblem{static_cast<const bf16*>(A1), lda1, static_cast<const bf16*>(W1), K, M1,
                          static_cast<bf16*>(C1), ldc1, static_cast<const bf16*>(bias1), gate1,
                          epi == EPI_GATE_RES ? static_cast<const bf16*>(C1) : nullptr, ldc1, 0, nullptr, 0};
  return gemm_bf16_tn(a, static_cast<cudaStream_t>(stream));
}

int utx_attention_bf16(const void* qkv, long ld_qkv, void* out, long ld_out, int S, int H, void* stream) {
  UTX_CHECK(qkv && out, "utx_attention_bf16: null pointer");
  return attention_bf16(static_cast<const bf16*>(qkv), ld_qkv, static_cast<bf16*>(out), ld_out, S, H,
                        static_cast<cudaStream_t>(stream));
}

int utx_ln_modulate(const void* x, long ldx, void* y, long ldy, int rows, int D, int rows0, const float* shift0,
                    const float* scale0, const float* shift1, const float* scale1, void* stream) {
  UTX_CHECK(x && y && shift1 && scale1, "utx_ln_modulate: null pointer");
  return ln_modulate(static_cast<const bf16*>(x), ldx, static_cast<bf16*>(y), ldy, rows, D, rows0, shift0, scale0,
                     shift1, scale1, static_cast<cudaStream_t>(stream));
}

int utx_rmsnorm_rope(void* qkv, long ld_qkv, int S, int H, int rows0, const void* wq0, const void* wk0,
                     const void* wq1, const void* wk1, const float* cos_t, const float* sin_t, void* stream) {
  UTX_CHECK(qkv && wq1 && wk1 && cos_t && sin_t, "utx_rmsnorm_rope: null pointer");
  return rmsnorm_rope(static_cast<bf16*>(qkv), ld_qkv, S, H, rows0, static_cast<const bf16*>(wq0),
                      static_cast<const bf16*>(wk0), static_cast<const bf16*>(wq1), static_cast<const bf16*>(wk1),
                      cos_t, sin_t, static_cast<cudaStream_t>(stream));
}

int utx_gemv_bf16(const void* W, const void* b, const float* x, float* y, int N, int K, int silu_in, int accumulate,
                  void* stream) {
  UTX_CHECK(W && x && y, "utx_gemv_bf16: null pointer");
  return gemv_bf16(static_cast<const bf16*>(W), static_cast<const bf16*>(b), x, y, N, K, silu_in, accumulate,
                   static_cast<cudaStream_t>(stream));
}

int utx_rope_table(const float* ids, int S, float* cos_t, float* sin_t, void* stream) {
  UTX_CHECK(ids && cos_t && sin_t, "utx_rope_table: null pointer");
  return rope_table(ids, S, cos_t, sin_t, static_cast<cudaStream_t>(stream));
}

int utx_euler_update(void* latents, const void* v, int rows, int cols, float dsigma, void* stream) {
  UTX_CHECK(latents && v, "utx_euler_update: null pointer");
  return euler_update(static_cast<bf16*>(latents), static_cast<const bf16*>(v), rows, cols, dsigma,
                      static_cast<cudaStream_t>(stream));
}

int utx_lora_merge(void* W, long ldw, const float* A, const float* B, int out_features, int in_features, int rank,
                   float scale, void* stream) {
  UTX_CHECK(W && A && B, "utx_lora_merge: null pointer");
  return lora_merge(static_cast<bf16*>(W), ldw, A, B, out_features, in_features, rank, scale,
                    static_cast<cudaStream_t>(stream));
}

size_t utx_rasterize_workspace_bytes(int B, int H, int W, int F) { return rasterize_workspace_bytes(B, H, W, F); }
int utx_rasterize(const float* pos, int pos_batched, int V, const int32_t* tri, int F, int B, int H, int W,
                  float* rast_out, void* workspace, void* stream) {
  UTX_CHECK(pos && tri && rast_out && workspace, "utx_rasterize: null pointer");
  return rasterize(pos, pos_batched, V, tri, F, B, H, W, rast_out, workspace, static_cast<cudaStream_t>(stream));
}
int utx_interpolate(const float* attr, int attr_batched, int V, int C, const float* rast, const int32_t* tri, int B, int H,
                    int W, float* out, void* stream) {
  UTX_CHECK(attr && rast && tri && out, "utx_interpolate: null pointer");
  return interpolate(attr, attr_batched, V, C, rast, tri, B, H, W, out, static_cast<cudaStream_t>(stream));
}
int utx_mv_visibility_filter(const float* attrs, const float* rast, const float* face_normals, const float* view_dirs,
                             int perspective, int n, int H, int W, float grad_thr, float cos_thr, unsigned char* mask_vis,
                             void* stream) {
  UTX_CHECK(attrs && rast && face_normals && view_dirs && mask_vis, "utx_mv_visibility_filter: null pointer");
  return mv_visibility_filter(attrs, rast, face_normals, view_dirs, perspective, n, H, W, grad_thr, cos_thr, mask_vis,
                              static_cast<cudaStream_t>(stream));
}
int utx_mvpaint_blend(const float* score, const long long* index, long long M, int k, const float* cloud_c,
                      const float* cloud_n, const float* tex_n, float* out, void* stream) {
  UTX_CHECK(score && index && cloud_c && cloud_n && tex_n && out, "utx_mvpaint_blend: null pointer");
  return mvpaint_blend(score, index, M, k, cloud_c, cloud_n, tex_n, out, static_cast<cudaStream_t>(stream));
}
int utx_transform_points(const float* vert, int V, const float* mats, int n, float* out, void* stream) {
  UTX_CHECK(vert && mats && out, "utx_transform_points: null pointer");
  return transform_points(vert, V, mats, n, out, static_cast<cudaStream_t>(stream));
}
size_t utx_bvh_nodes_bytes(int F) { return bvh_nodes_bytes(F); }
size_t utx_bvh_workspace_bytes(int F) { return bvh_workspace_bytes(F); }
int utx_bvh_build(const float* vert, int V, const int32_t* tri, int F, void* nodes, void* workspace, size_t workspace_bytes,
                  void* stream) {
  UTX_CHECK(vert && tri && nodes && workspace, "utx_bvh_build: null pointer");
  return bvh_build(vert, V, tri, F, nodes, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
int utx_bvh_export(const void* nodes, int F, int32_t* info, float* aabb, void* stream) {
  UTX_CHECK(nodes && info && aabb, "utx_bvh_export: null pointer");
  return bvh_export(nodes, F, info, aabb, static_cast<cudaStream_t>(stream));
}
int utx_bvh_intersect(const void* nodes, const float* vert, const int32_t* tri, int F, const float* rays_o, const float* rays_d,
                      long long N, unsigned char* hit, int32_t* tri_idx, float* loc, float* uv, void* stream) {
  UTX_CHECK(nodes && vert && tri && rays_o && rays_d && hit && tri_idx && loc && uv, "utx_bvh_intersect: null pointer");
  return bvh_intersect(nodes, vert, tri, F, rays_o, rays_d, N, hit, tri_idx, loc, uv, static_cast<cudaStream_t>(stream));
}
int utx_knn1(const float* src, int n_src, const float* dst, long long M, long long* index, float* score, void* nodes,
             void* workspace, size_t workspace_bytes, void* stream) {
  UTX_CHECK(src && dst && index && score && nodes && workspace, "utx_knn1: null pointer");
  return knn1(src, n_src, dst, M, index, score, nodes, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
int utx_knn(const float* src, int n_src, const float* dst, long long M, int k, long long* index, float* score, void* nodes,
            void* workspace, size_t workspace_bytes, void* stream) {
  UTX_CHECK(src && dst && index && score && nodes && workspace, "utx_knn: null pointer");
  return knn(src, n_src, dst, M, k, index, score, nodes, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
size_t utx_uv_bake_workspace_bytes(int H2, int W2) { return uv_bake_workspace_bytes(H2, W2); }
int utx_uv_bake_layout(int H2, int W2, size_t* off_owner, size_t* off_pos, size_t* off_color, size_t* off_seam) {
  UTX_CHECK(off_owner && off_pos && off_color && off_seam, "utx_uv_bake_layout: null pointer");
  UTX_CHECK(H2 >= 8 && W2 >= 8, "utx_uv_bake_layout: atlas too small");
  uv_bake_layout(H2, W2, off_owner, off_pos, off_color, off_seam);
  return 0;
}
int utx_uv_bake_visibility(const float* vert, int V, const int32_t* tri, int F, const void* nodes, const float* rast2d, int H2,
                           int W2, int n_views, const float* view_mats, const float* view_dirs, int perspective, const int32_t* priority,
                           const float* images_rgba, int H, int W, float cos_thresh, unsigned char* mask2d,
                           unsigned char* mask_vis, void* workspace, size_t workspace_bytes, void* stream) {
  UTX_CHECK(vert && tri && nodes && rast2d && view_mats && view_dirs && priority && images_rgba && mask2d && mask_vis && workspace,
            "utx_uv_bake_visibility: null pointer");
  return uv_bake_visibility(vert, V, tri, F, nodes, rast2d, H2, W2, n_views, view_mats, view_dirs, perspective, priority, images_rgba, H, W,
                            cos_thresh, mask2d, mask_vis, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
int utx_uv_bake_fill(const unsigned char* mask2d, int H2, int W2, int k, int32_t* nn_index, void* workspace,
                     size_t workspace_bytes, void* stream) {
  UTX_CHECK(mask2d && workspace, "utx_uv_bake_fill: null pointer");
  return uv_bake_fill(mask2d, H2, W2, k, nn_index, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
size_t utx_uv_bake_views_workspace_bytes(int n_views, int H, int W) { return uv_bake_views_workspace_bytes(n_views, H, W); }
int utx_uv_bake_views_knn(const float* pix_pos, const float* images_rgba, int n_views, int H, int W, int k, int merge,
                          const unsigned char* mask2d, int H2, int W2, void* workspace, size_t workspace_bytes, void* scratch,
                          size_t scratch_bytes, void* stream) {
  UTX_CHECK(pix_pos && images_rgba && mask2d && workspace && scratch, "utx_uv_bake_views_knn: null pointer");
  return uv_bake_views_knn(pix_pos, images_rgba, n_views, H, W, k, merge, mask2d, H2, W2, workspace, workspace_bytes, scratch,
                           scratch_bytes, static_cast<cudaStream_t>(stream));
}
int utx_uv_bake_finish(const unsigned char* mask2d, int H2, int W2, int blur, const float* blur_k2d, float blur_gamma,
                       float* color, void* workspace, size_t workspace_bytes, void* stream) {
  UTX_CHECK(mask2d && color && workspace && (!blur || blur_k2d), "utx_uv_bake_finish: null pointer");
  return uv_bake_finish(mask2d, H2, W2, blur, blur_k2d, blur_gamma, color, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream));
}
int utx_uv_bake(const float* vert, int V, const int32_t* tri, int F, const void* nodes, const float* rast2d, int H2, int W2,
                int n_views, const float* view_mats, const float* view_dirs, int perspective, const int32_t* priority,
                const float* images_rgba, int H, int W, float cos_thresh, const float* blur_k2d, float blur_gamma,
                const float* grid_lo, float grid_extent, unsigned char* mask2d, unsigned char* mask_vis, float* color,
                int32_t* nn_index, void* workspace, size_t workspace_bytes, void* stream) {
  UTX_CHECK(vert && tri && nodes && rast2d && view_mats && view_dirs && priority && images_rgba && blur_k2d && grid_lo &&
                mask2d && mask_vis && color && workspace,
            "utx_uv_bake: null pointer");
  return uv_bake(vert, V, tri, F, nodes, rast2d, H2, W2, n_views, view_mats, view_dirs, perspective, priority, images_rgba, H, W,
                 cos_thresh, blur_k2d, blur_gamma, grid_lo, grid_extent, mask2d, mask_vis, color, nn_index, workspace,
                 workspace_bytes, static_cast<cudaStream_t>(stream));
}

int utx_conv3x3_nhwc(const void* x, int N, int H, int W, int C, const void* w, const void* bias, int Cout, void* y, long ldy,
                     const float* gate, const void* res, long ldres, void* stream) {
  UTX_CHECK(x && w && y, "utx_conv3x3_nhwc: null pointer");
  return conv3x3_nhwc(static_cast<const bf16*>(x), N, H, W, C, static_cast<const bf16*>(w), static_cast<const bf16*>(bias), Cout,
                      static_cast<bf16*>(y), ldy, gate, static_cast<const bf16*>(res), ldres, static_cast<cudaStream_t>(stream));
}
int utx_im2col3x3(const void* x, int N, int Hin, int Win, int C, int up, int stride, int pad, int Ho, int Wo, int Kpad,
                  void* out, void* stream) {
  UTX_CHECK(x && out, "utx_im2col3x3: null pointer");
  return im2col3x3(static_cast<const bf16*>(x), N, Hin, Win, C, up, stride, pad, Ho, Wo, Kpad, static_cast<bf16*>(out),
                   static_cast<cudaStream_t>(stream));
}
int utx_upsample2x_nhwc(const void* x, int N, int H, int W, int C, void* y, void* stream) {
  UTX_CHECK(x && y, "utx_upsample2x_nhwc: null pointer");
  return upsample2x_nhwc(static_cast<const bf16*>(x), N, H, W, C, static_cast<bf16*>(y), static_cast<cudaStream_t>(stream));
}
size_t utx_groupnorm_workspace_bytes(int N, int HW, int C, int G) { return groupnorm_workspace_bytes(N, HW, C, G); }
int utx_groupnorm_nhwc(const void* x, void* y, int N, int HW, int C, int G, const float* gamma, const float* beta, int silu,
                       void* stats_ws, void* stream) {
  UTX_CHECK(x && y && gamma && beta && stats_ws, "utx_groupnorm_nhwc: null pointer");
  return groupnorm_nhwc(static_cast<const bf16*>(x), static_cast<bf16*>(y), N, HW, C, G, gamma, beta, silu,
                        static_cast<double*>(stats_ws), static_cast<cudaStream_t>(stream));
}
int utx_gemm_bf16_f32out(const void* A, long lda, const void* W, long ldw, const void* bias, float* C, long ldc, int M, int N,
                         int K, float scale, void* stream) {
  UTX_CHECK(A && W && C, "utx_gemm_bf16_f32out: null pointer");
  GemmArgs a{};
  a.N = N; a.K = K; a.epi = EPI_BIAS_F32; a.gelu_col_start = 0; a.out_scale = scale; a.nprob = 1;
  a.prob[0] = GemmProblem{static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, M, reinterpret_cast<bf16*>(C),
                          ldc, static_cast<const bf16*>(bias), nullptr, nullptr, 0, 0, nullptr, 0};
  return gemm_bf16_tn(a, static_cast<cudaStream_t>(stream));
}
int utx_softmax_rows(const float* S, long lds, void* P, long ldp, int M, int N, void* stream) {
  UTX_CHECK(S && P, "utx_softmax_rows: null pointer");
  return softmax_rows(S, lds, static_cast<bf16*>(P), ldp, M, N, static_cast<cudaStream_t>(stream));
}
int utx_transpose_bf16(const void* x, long ldx, void* y, long ldy, int R, int C, void* stream) {
  UTX_CHECK(x && y, "utx_transpose_bf16: null pointer");
  return transpose_bf16(static_cast<const bf16*>(x), ldx, static_cast<bf16*>(y), ldy, R, C, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
