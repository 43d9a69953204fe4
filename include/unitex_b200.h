/* unitex_b200 -- C ABI of the B200-native UniTEX hot path (libunitex_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  All data pointers are
 * DEVICE pointers unless a comment says "host"; bf16 tensors are row-major with the stated leading dimension in
 * elements; every call only enqueues work on `stream` (a cudaStream_t passed as void*).  Outputs are caller
 * allocated and borrowed until the stream is synchronised.  Return value: 0 = ok, non-zero = error, message in
 * utx_last_error() (thread-local).  No global state; one utx_flux handle per device.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the UniTEX repository root).
 */
#ifndef UNITEX_B200_H
#define UNITEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* utx_last_error(void);
int utx_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * FLUX.1-dev MM-DiT transformer.  Replaces `self.transformer(...)` in PBRFluxPipeline.__call__
 * (flux_piplines/texturing/pipeline.py:646-656; flux_piplines/delight/pipeline.py identical), i.e. diffusers'
 * FluxTransformer2DModel.forward, and the Euler loop around it (:634-681).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct utx_flux utx_flux;

typedef struct utx_flux_config {
  int in_channels;            /* 64 */
  int num_layers;             /* 19 double-stream blocks */
  int num_single_layers;      /* 38 single-stream blocks */
  int num_heads;              /* 24 */
  int head_dim;               /* 128 (required) */
  int joint_attention_dim;    /* 4096 */
  int pooled_projection_dim;  /* 768 */
  int guidance_embeds;        /* 1 */
  int mlp_ratio;              /* 4 */
} utx_flux_config;

/* Weights are bf16 nn.Linear layout [out, in] (+ bias [out]); LoRA adapters already merged (utx_lora_merge). */
typedef struct utx_double_block {
  const void *w_qkv_img, *b_qkv_img;   /* [3D, D]  attn.to_q | to_k | to_v stacked on rows            */
  const void *w_qkv_txt, *b_qkv_txt;   /* [3D, D]  attn.add_q_proj | add_k_proj | add_v_proj          */
  const void *rms_q_img, *rms_k_img;   /* [128]    attn.norm_q / norm_k weights                        */
  const void *rms_q_txt, *rms_k_txt;   /* [128]    attn.norm_added_q / norm_added_k                    */
  const void *w_out_img, *b_out_img;   /* [D, D]   attn.to_out.0                                       */
  const void *w_out_txt, *b_out_txt;   /* [D, D]   attn.to_add_out                                     */
  const void *w_ff1_img, *b_ff1_img;   /* [4D, D]  ff.net.0.proj                                       */
  const void *w_ff2_img, *b_ff2_img;   /* [D, 4D]  ff.net.2                                            */
  const void *w_ff1_txt, *b_ff1_txt;   /* [4D, D]  ff_context.net.0.proj                               */
  const void *w_ff2_txt, *b_ff2_txt;   /* [D, 4D]  ff_context.net.2                                    */
} utx_double_block;

typedef struct utx_single_block {
  const void *w_qkvmlp, *b_qkvmlp;     /* [3D + 4D, D]  attn.to_q | to_k | to_v | proj_mlp             */
  const void *rms_q, *rms_k;           /* [128]                                                        */
  const void *w_out, *b_out;           /* [D, 5D]  proj_out over cat[attn, mlp]                        */
} utx_single_block;

typedef struct utx_flux_weights {
  const void *w_x_embed, *b_x_embed;       /* [D, in_channels]       x_embedder                        */
  const void *w_ctx_embed, *b_ctx_embed;   /* [D, joint_attention]   context_embedder                  */
  const void *w_t1, *b_t1, *w_t2, *b_t2;   /* time_text_embed.timestep_embedder.linear_1/2             */
  const void *w_g1, *b_g1, *w_g2, *b_g2;   /* time_text_embed.guidance_embedder.linear_1/2 (or NULL)   */
  const void *w_p1, *b_p1, *w_p2, *b_p2;   /* time_text_embed.text_embedder.linear_1/2                 */
  /* every adaLN linear stacked on rows, in execution order: per double block [norm1.linear (6D) |
   * norm1_context.linear (6D)], per single block [norm.linear (3D)], then norm_out.linear (2D). */
  const void *w_mod, *b_mod;               /* [(12 L + 3 Ls + 2) D, D]                                 */
  const utx_double_block* double_blocks;   /* host array [num_layers]                                  */
  const utx_single_block* single_blocks;   /* host array [num_single_layers]                           */
  const void *w_proj_out, *b_proj_out;     /* [in_channels, D]       proj_out                          */
} utx_flux_weights;

int utx_flux_create(const utx_flux_config* cfg, utx_flux** out);
void utx_flux_destroy(utx_flux* h);
/* copies the pointer tables; the device memory stays owned by the caller */
int utx_flux_set_weights(utx_flux* h, const utx_flux_weights* w);
size_t utx_flux_workspace_bytes(const utx_flux* h, int s_txt, int s_img);
/* Per-call constants: RoPE table from ids [s_txt + s_img, 3] fp32 (txt rows first -- torch.cat((txt_ids, img_ids)),
 * pipeline.py:652-653), context embedding of enc [s_txt, joint_attention_dim] bf16 (the reference passes zeros,
 * pipeline.py:538-543) and the pooled projection [pooled_projection_dim] fp32. */
int utx_flux_prepare(utx_flux* h, void* workspace, size_t workspace_bytes, const float* ids, const void* enc,
                     const float* pooled, int s_txt, int s_img, void* stream);
/* One transformer forward: latents [s_img, in_channels] bf16 -> v_out [s_img, in_channels] bf16.
 * `timestep` is t/1000 and `guidance` the raw scale, exactly as the reference passes them (:648-649); both are
 * multiplied by 1000 in bf16 inside, like FluxTransformer2DModel.forward does on a bf16 model. */
int utx_flux_forward(utx_flux* h, const void* latents, float timestep, float guidance, void* v_out, void* stream);
/* The hot loop (:634-681): n_steps x { forward; latents[:s_noise] += (sigma[i+1]-sigma[i]) * v (fp32, stored bf16) }.
 * Rows >= s_noise are the clean condition tokens and are never written, which is what the per-step re-imposition
 * (:644-645) amounts to.  sigmas: HOST fp32 [n_steps + 1]. */
int utx_flux_denoise(utx_flux* h, void* latents, int s_noise, const float* sigmas, int n_steps, float guidance,
                     void* stream);
/* utx_flux_denoise runs a step (forward + Euler update) as ONE CUDA-graph launch from the second step on a given
 * (latents, s_noise) onwards: the ~240 kernel launches of a step are captured once, the step's scalars (timestep, guidance,
 * sigma difference) live in device memory and are rewritten before every launch.  Bit-identical to the eager path
 * (UTX_FLUX_GRAPH=0 disables it; profiling mode runs eagerly).  Number of steps that ran as graph launches so far: */
long utx_flux_graph_replays(const utx_flux* h);

/* Sequence-parallel ("Ulysses") mode: ONE grid's token sequence split over the ranks of `comm` (SURVEY 8e "if one grid must
 * span GPUs"; the reference has no counterpart, it runs one grid on one GPU).  Every rank keeps the full merged weights and owns
 * S / nranks consecutive rows of the [txt | img] sequence for every Linear / LayerNorm; around each joint attention
 * (attention_processor.py:81-91) an all-to-all turns "my rows, all heads" into "all rows, my heads" and back:
 * the QKV GEMM's epilogue writes q | k | v straight into the per-peer send layout, the attention runs over the whole sequence
 * for num_heads / nranks heads.  Every rank passes the SAME ids / enc / latents to utx_flux_prepare / forward / denoise and ends
 * with the SAME v / latents (v is all-gathered), bit-identical to the single-GPU result.  Requires nranks | num_heads and
 * nranks | (s_txt + s_img).  comm = NULL switches back; call utx_flux_prepare again after changing the mode. */
typedef struct utx_comm utx_comm;
int utx_flux_set_sequence_parallel(utx_flux* h, utx_comm* comm);

/* Instrumentation for bench.py: kernel launches issued by the engine per category, and (after
 * utx_flux_profile(h, 1)) the CUDA-event time of each category on the launching stream.  launches/ms: [4]
 * indexed by UTX_PROF_*; reading synchronises on the recorded events. */
enum { UTX_PROF_GEMM = 0, UTX_PROF_ATTN = 1, UTX_PROF_ELEM = 2, UTX_PROF_OTHER = 3, UTX_PROF_NCAT = 4 };
int utx_flux_profile(utx_flux* h, int enable);
int utx_flux_profile_read(utx_flux* h, long* launches, float* ms, int reset);

/* W[out,in] (bf16, row stride ldw) += scale * B[out,rank] @ A[rank,in] (fp32).  Replaces peft's un-merged
 * LoRA evaluation selected by pipeline.set_adapters (pipeline.py:108-118,245,263). */
int utx_lora_merge(void* W, long ldw, const float* A, const float* B, int out_features, int in_features, int rank,
                   float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Building blocks (exported for the parity tests; the engine above calls the same launchers).
 * ------------------------------------------------------------------------------------------------------------------ */
/* C[M,N] = epi(A[M,K] @ W[N,K]^T + bias).  epi: 0 bias, 1 bias+GELU(tanh), 2 res + gate*(.)  */
int utx_gemm_bf16(const void* A, long lda, const void* W, long ldw, const void* bias, void* C, long ldc, int M, int N,
                  int K, int epi, const float* gate, const void* res, long ldres, void* stream);
/* QKV projection with the joint attention's per-head RMSNorm(eps 1e-6, weight) + RoPE fused into the epilogue for the
 * q and k thirds (attention_processor.py:43-59,85-87): C[M, 3*Hd*128]; wq/wk [128] bf16; cos/sin [>= row_offset+M, 128] fp32 */
int utx_gemm_bf16_qkv(const void* A, long lda, const void* W, const void* bias, void* C, long ldc, int M, int heads, int K,
                      const void* wq, const void* wk, const float* cos_t, const float* sin_t, int row_offset, void* stream);
/* two problems sharing N, K and the epilogue in one launch (txt + img streams) */
int utx_gemm_bf16_grouped2(const void* A0, long lda0, const void* W0, const void* bias0, void* C0, long ldc0, int M0,
                           const void* A1, long lda1, const void* W1, const void* bias1, void* C1, long ldc1, int M1,
                           int N, int K, int epi, const float* gate0, const float* gate1, void* stream);
/* softmax(q k^T / sqrt(128)) v over qkv [S, 3*H*128] (attention_processor.py:89-91) */
int utx_attention_bf16(const void* qkv, long ld_qkv, void* out, long ld_out, int S, int H, void* stream);
/* LayerNorm(eps 1e-6, no affine) * (1 + scale) + shift per stream: rows [0, rows0) use (shift0, scale0), the rest (shift1, scale1).
 * AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous of diffusers' FluxTransformer2DModel [ext], reached from
 * `self.transformer(...)` at flux_piplines/texturing/pipeline.py:646-656. */
int utx_ln_modulate(const void* x, long ldx, void* y, long ldy, int rows, int D, int rows0, const float* shift0,
                    const float* scale0, const float* shift1, const float* scale1, void* stream);
/* in place on the q and k thirds of qkv: per-head RMSNorm(eps 1e-6) x weight, then the rotary embedding
 * (attention_processor.py:53-56 norm_q / norm_k, :86-87 apply_rotary_emb); rows [0, rows0) take (wq0, wk0), the rest (wq1, wk1) */
int utx_rmsnorm_rope(void* qkv, long ld_qkv, int S, int H, int rows0, const void* wq0, const void* wk0,
                     const void* wq1, const void* wk1, const float* cos_t, const float* sin_t, void* stream);
/* y[N] (+)= W[N,K] @ (silu_in ? silu(x) : x) + b: the timestep / guidance / pooled-text embedders and every block's modulation
 * Linear(SiLU(temb)) [ext: CombinedTimestepGuidanceTextProjEmbeddings, AdaLayerNormZero.linear], one evaluation per step (:646-656) */
int utx_gemv_bf16(const void* W, const void* b, const float* x, float* y, int N, int K, int silu_in, int accumulate,
                  void* stream);
/* cos / sin [S,128] of FluxPosEmbed(theta 10000, axes_dim (16, 56, 56)) [ext] over the ids the sampler builds at
 * flux_piplines/texturing/pipeline.py:303-393 (text ids | latent ids | condition ids with their offsets) */
int utx_rope_table(const float* ids, int S, float* cos_t, float* sin_t, void* stream);
/* FlowMatchEulerDiscreteScheduler.step [ext] as called at flux_piplines/texturing/pipeline.py:660:
 * latents = bf16(fp32(latents) + bf16(dsigma * v)) on the first `rows` rows (the noise rows; the condition rows stay) */
int utx_euler_update(void* latents, const void* v, int rows, int cols, float dsigma, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Bake: rasterise -> back-project -> UV bake.  Replaces nvdiffrast, the Slang LBVH ray tracer and the torch tail of
 * NVDiffRendererInverse (TextureTools/texturetools/render/nvdiffrast/renderer_inverse.py).
 * ------------------------------------------------------------------------------------------------------------------ */
/* dr.rasterize (renderer_inverse.py:183,273; renderer_base.py:142): pos [B or 1, V, 4] clip space fp32, tri [F,3] int32
 * -> rast [B,H,W,4] = (u, v, z/w, triangle_id + 1), 0 = background.  workspace: utx_rasterize_workspace_bytes. */
size_t utx_rasterize_workspace_bytes(int B, int H, int W, int F);
int utx_rasterize(const float* pos, int pos_batched, int V, const int32_t* tri, int F, int B, int H, int W,
                  float* rast_out, void* workspace, void* stream);
/* dr.interpolate (renderer_inverse.py:188,277,288): out[b,y,x,:] = u a0 + v a1 + (1-u-v) a2 */
int utx_interpolate(const float* attr, int attr_batched, int V, int C, const float* rast, const int32_t* tri, int B, int H,
                    int W, float* out, void* stream);
/* The visible-pixel mask of mv_to_pcd(filt_gradient_points=True) (renderer_inverse.py:186-214): attrs [n,H,W,6] =
 * dr.interpolate of (vertex position | vertex normal) (:188), rast [n,H,W,4], face_normals [F,3] (PBRMesh.normals), view_dirs
 * DEVICE [n,3] = -c2w[:3,2] (perspective = 0) or the camera positions c2w[:3,3] (perspective = 1).  mask_vis u8 [n,H,W] = covered
 * & cos(ray, face normal) < cos_thr & |torch.gradient(attrs)| < grad_thr on all 31 pixels x - 15 .. x + 15 of the row (the
 * reference's MaxPool2d(31, 1, 15) on an [n,H,W,1] tensor erodes along x only; kept).  H, W >= 2 like torch.gradient. */
int utx_mv_visibility_filter(const float* attrs, const float* rast, const float* face_normals, const float* view_dirs,
                             int perspective, int n, int H, int W, float grad_thr, float cos_thr, unsigned char* mask_vis,
                             void* stream);
/* kdtree_method='mvpaint' (renderer_inverse.py:390-399): score / index [M,k] from utx_knn over the union pixel cloud, cloud_c /
 * cloud_n [N,3] colour and face normal per cloud point, tex_n [M,3] face normal per texel -> out [M,3] = sum c w / sum w with
 * w = normalize(1 / score, p = 1) x cosine_similarity(cloud_n[index], tex_n); non-finite results -> 0. */
int utx_mvpaint_blend(const float* score, const long long* index, long long M, int k, const float* cloud_c,
                      const float* cloud_n, const float* tex_n, float* out, void* stream);
/* vertices_homo @ (P @ W2C)^T (renderer_inverse.py:178,263): out [n, V, 4] */
int utx_transform_points(const float* vert, int V, const float* mats, int n, float* out, void* stream);
/* RayTracing(vertices, faces) / update_raw (raytracing/__init__.py:12-80; rt_aprmis/bvhhelpers.py:20-83): builds the
 * same LBVH the reference builds into `nodes` (utx_bvh_nodes_bytes(F) bytes: the reference-layout nodes, 48 B each, followed by the
 * traversal layout the intersect / bake kernels walk). */
size_t utx_bvh_nodes_bytes(int F);
size_t utx_bvh_workspace_bytes(int F);
int utx_bvh_build(const float* vert, int V, const int32_t* tri, int F, void* nodes, void* workspace, size_t workspace_bytes,
                  void* stream);
/* reference node layout for inspection: info [2F-1, 3] (left, right, prim), aabb [2F-1, 6] */
int utx_bvh_export(const void* nodes, int F, int32_t* info, float* aabb, void* stream);
/* intersects_closest (rt_aprmis/__init__.py:36-86): hit u8 [N], tri_idx i32 [N] (-1 = miss), loc [N,3], uv [N,2] */
int utx_bvh_intersect(const void* nodes, const float* vert, const int32_t* tri, int F, const float* rays_o, const float* rays_d,
                      long long N, unsigned char* hit, int32_t* tri_idx, float* loc, float* uv, void* stream);
/* knn(src, dst, k=1) (pcd/knn/__init__.py:104-114, default backend torch_kdtree): exact nearest source point of every
 * dst point, lowest index on ties.  index int64 [M], score fp32 [M] = Euclidean distance.  nodes: utx_bvh_nodes_bytes(n_src),
 * workspace: utx_bvh_workspace_bytes(n_src). */
int utx_knn1(const float* src, int n_src, const float* dst, long long M, long long* index, float* score, void* nodes,
             void* workspace, size_t workspace_bytes, void* stream);
/* knn(src, dst, k) for 1 <= k <= min(32, n_src) (pcd/knn/__init__.py:104-114; k = 8+1 / 32 in renderer_inverse.py:377,382,
 * 429,464): index int64 [M,k], score fp32 [M,k] = Euclidean distance, each row ascending by (distance, index).  Buffers as
 * for utx_knn1. */
int utx_knn(const float* src, int n_src, const float* dst, long long M, int k, long long* index, float* score, void* nodes,
            void* workspace, size_t workspace_bytes, void* stream);
/* uv_to_pcd + bake_mv_to_uv_reproject_blur (renderer_inverse.py:243-365,574-633) fused.  rast2d: UV raster [H2,W2,4];
 * view_mats/view_dirs/priority/grid_lo: HOST arrays ([n,16] P@W2C row-major, [n,3] = -c2w[:3,2] for orthographic views (perspective = 0) or the
 * camera positions c2w[:3,3] (perspective = 1; renderer_inverse.py:279-284), [n], [3]);
 * images_rgba: device [n,H,W,4] = view colour + visible alpha; blur_k2d: device [49].  Outputs: mask2d u8 [H2*W2],
 * mask_vis u8 [n, H2*W2], color [H2*W2, 3] fp32, nn_index i32 [H2*W2] or NULL. */
size_t utx_uv_bake_workspace_bytes(int H2, int W2);
int utx_uv_bake(const float* vert, int V, const int32_t* tri, int F, const void* nodes, const float* rast2d, int H2, int W2,
                int n_views, const float* view_mats, const float* view_dirs, int perspective, const int32_t* priority,
                const float* images_rgba, int H, int W, float cos_thresh, const float* blur_k2d, float blur_gamma,
                const float* grid_lo, float grid_extent, unsigned char* mask2d, unsigned char* mask_vis, float* color,
                int32_t* nn_index, void* workspace, size_t workspace_bytes, void* stream);

/* Staged form of utx_uv_bake over the SAME workspace, for the bake variants that differ between the stages
 * (`method='kdtree'`, renderer_inverse.py:367-433; `*_inpainting=True` with a registered query field, :93-103,:427-432,:609-614):
 *   utx_uv_bake_visibility   uv_to_pcd (:243-365) + priority composite + seam mask (:591-605): writes mask2d, mask_vis and
 *                            leaves owner i8 [T] (winning view, -1 = none), pos fp32 [T,3], colour fp32 [T,3], seam u8 [T] in
 *                            the workspace at the byte offsets utx_uv_bake_layout reports
 *   utx_uv_bake_views_knn    kdtree variant, visible part: merge = 0 `order_mean` -- texels owned by view i take the mean colour
 *                            of their k nearest points of view i's pixel cloud (:406-417); merge = 1 `mean` -- every covered texel
 *                            from the union cloud (:385-389).  pix_pos [n,H,W,3] = vertex positions interpolated at the pixels
 *                            (:188), images_rgba as for utx_uv_bake (alpha = mask_visiable); scratch:
 *                            utx_uv_bake_views_workspace_bytes(n, H, W)
 *   utx_uv_bake_fill         covered-but-unowned texels <- mean colour of their k nearest owned texels; k = 1 reproject
 *                            (:606-615), k = 32 kdtree (:427-431).  A caller with a query field writes those colours itself
 *                            (through the layout offsets) and skips this stage.
 *   utx_uv_bake_finish       blur != 0: lens blur on the seam texels (:617-624); then pull-push (:627, :423) -> color [T,3]
 * utx_uv_bake == visibility; fill(k = 1); finish(blur = 1). */
int utx_uv_bake_layout(int H2, int W2, size_t* off_owner, size_t* off_pos, size_t* off_color, size_t* off_seam);
int utx_uv_bake_visibility(const float* vert, int V, const int32_t* tri, int F, const void* nodes, const float* rast2d, int H2,
                           int W2, int n_views, const float* view_mats, const float* view_dirs, int perspective, const int32_t* priority,
                           const float* images_rgba, int H, int W, float cos_thresh, unsigned char* mask2d,
                           unsigned char* mask_vis, void* workspace, size_t workspace_bytes, void* stream);
size_t utx_uv_bake_views_workspace_bytes(int n_views, int H, int W);
int utx_uv_bake_views_knn(const float* pix_pos, const float* images_rgba, int n_views, int H, int W, int k, int merge,
                          const unsigned char* mask2d, int H2, int W2, void* workspace, size_t workspace_bytes, void* scratch,
                          size_t scratch_bytes, void* stream);
int utx_uv_bake_fill(const unsigned char* mask2d, int H2, int W2, int k, int32_t* nn_index, void* workspace,
                     size_t workspace_bytes, void* stream);
/* blur: 0 none, 1 the 7x7 kernel with a zero border (lens blur, gamma 5), 2 with a mirrored border (gaussian, gamma 1) */
int utx_uv_bake_finish(const unsigned char* mask2d, int H2, int W2, int blur, const float* blur_k2d, float blur_gamma,
                       float* color, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * FLUX VAE building blocks (AutoencoderKL; flux_piplines/texturing/pipeline.py:226-238 encode, :688-692 decode).
 * Activations NHWC bf16; a 3x3 conv = utx_im2col3x3 + utx_gemm_bf16 with W arranged [Cout, ky, kx, Cin].
 * ------------------------------------------------------------------------------------------------------------------ */
/* out [N*Ho*Wo, Kpad] (zero padded beyond 9*C); up = 1|2 nearest upsample folded in; pad = top/left padding */
/* Implicit-GEMM 3x3 convolution, stride 1, zero padding 1 (the ResnetBlock2D / mid-block convolutions of AutoencoderKL [ext]):
 * y[N*H*W, Cout] = conv(x) + bias, or gate[c] * (conv(x) + bias) + res when gate / res are given (fused residual add).  No
 * im2col buffer: the A tiles are TMA boxes of the activation shifted by the tap.  Cin % 64 == 0, Cout % 8 == 0,
 * W % 128 == 0 or (128 % W == 0 and H % (128 / W) == 0); other convolutions go through utx_im2col3x3 + utx_gemm_bf16. */
int utx_conv3x3_nhwc(const void* x, int N, int H, int W, int C, const void* w, const void* bias, int Cout, void* y, long ldy,
                     const float* gate, const void* res, long ldres, void* stream);
int utx_im2col3x3(const void* x, int N, int Hin, int Win, int C, int up, int stride, int pad, int Ho, int Wo, int Kpad,
                  void* out, void* stream);
/* nearest-neighbour 2x upsampling NHWC [N,H,W,C] -> [N,2H,2W,C] (Upsample2D [ext], ahead of its implicit-GEMM convolution) */
int utx_upsample2x_nhwc(const void* x, int N, int H, int W, int C, void* y, void* stream);
/* GroupNorm(G, eps 1e-6, affine fp32) [+ SiLU], deterministic; stats_ws: utx_groupnorm_workspace_bytes(N, HW, C, G) bytes, 8B aligned */
size_t utx_groupnorm_workspace_bytes(int N, int HW, int C, int G);
int utx_groupnorm_nhwc(const void* x, void* y, int N, int HW, int C, int G, const float* gamma, const float* beta, int silu,
                       void* stats_ws, void* stream);
/* C fp32 [M,N] = scale * (A @ W^T + bias) */
int utx_gemm_bf16_f32out(const void* A, long lda, const void* W, long ldw, const void* bias, float* C, long ldc, int M, int N,
                         int K, float scale, void* stream);
/* P bf16 [M,N] = softmax over each row of the fp32 scores S (the mid block's single-head attention, Attention [ext]) */
int utx_softmax_rows(const float* S, long lds, void* P, long ldp, int M, int N, void* stream);
/* y [C,R] = x [R,C]^T (V^T operand of P @ V on the K-major GEMM) */
int utx_transpose_bf16(const void* x, long ldx, void* y, long ldy, int R, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * FLUX VAE, whole passes.  Replaces `self.vae.encode(image).latent_dist` and `self.vae.decode(latents)` of the reference
 * sampler (flux_piplines/texturing/pipeline.py:226-238 and :688-692; diffusers AutoencoderKL [ext]).  Tensors at the boundary
 * are NCHW like the reference's; activations inside are NHWC bf16.  Weight layout (packed once by the caller):
 *   3x3 conv    bf16 [ceil8(Cout)][ceil64(9*Cin)], K ordered (ky, kx, cin), zero padded; bias bf16 [ceil8(Cout)]
 *   1x1 conv / Linear   bf16 [Cout][Cin], bias bf16 [Cout]
 *   GroupNorm   fp32 weight / bias [C]
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct utx_vae utx_vae;

typedef struct utx_vae_config {
  int in_channels;             /* 3 */
  int latent_channels;         /* 16 */
  int num_blocks;              /* 4 */
  int block_out_channels[8];   /* 128, 256, 512, 512 */
  int layers_per_block;        /* 2 (the decoder runs layers_per_block + 1 resnets per up block) */
  int norm_num_groups;         /* 32 */
} utx_vae_config;

typedef struct utx_vae_resnet {   /* ResnetBlock2D */
  const float *gn1_w, *gn1_b;
  const void *conv1_w, *conv1_b;
  const float *gn2_w, *gn2_b;
  const void *conv2_w, *conv2_b;
  const void *short_w, *short_b;   /* conv_shortcut (1x1) or NULL when cin == cout */
  int cin, cout;
} utx_vae_resnet;

typedef struct utx_vae_attn {     /* mid_block.attentions.0: one head of C channels */
  const float *gn_w, *gn_b;
  const void *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo;
} utx_vae_attn;

typedef struct utx_vae_mid {
  utx_vae_resnet res0;
  utx_vae_attn attn;
  utx_vae_resnet res1;
} utx_vae_mid;

typedef struct utx_vae_weights {
  const void *enc_conv_in_w, *enc_conv_in_b;
  const utx_vae_resnet* enc_res;            /* host array [num_blocks * layers_per_block], block-major            */
  const void* const* enc_down_w;            /* host arrays [num_blocks - 1]: down_blocks.i.downsamplers.0.conv    */
  const void* const* enc_down_b;
  utx_vae_mid enc_mid;
  const float *enc_norm_w, *enc_norm_b;     /* conv_norm_out */
  const void *enc_conv_out_w, *enc_conv_out_b;
  const void *dec_conv_in_w, *dec_conv_in_b;
  utx_vae_mid dec_mid;
  const utx_vae_resnet* dec_res;            /* host array [num_blocks * (layers_per_block + 1)]                   */
  const void* const* dec_up_w;              /* host arrays [num_blocks - 1]: up_blocks.i.upsamplers.0.conv        */
  const void* const* dec_up_b;
  const float *dec_norm_w, *dec_norm_b;
  const void *dec_conv_out_w, *dec_conv_out_b;
} utx_vae_weights;

int utx_vae_create(const utx_vae_config* cfg, utx_vae** out);
void utx_vae_destroy(utx_vae* h);
/* copies the pointer tables; the device memory stays owned by the caller */
int utx_vae_set_weights(utx_vae* h, const utx_vae_weights* w);
/* decode != 0: (H, W) is the latent size; else the image size.  256B-aligned workspace of this many bytes. */
size_t utx_vae_workspace_bytes(utx_vae* h, int N, int H, int W, int decode);
/* z [N, latent_channels, h, w] bf16 (already latents / scaling_factor + shift_factor, :689) -> img [N, in_channels, 8h, 8w] bf16 */
int utx_vae_decode(utx_vae* h, const void* z, int N, int H, int W, void* img, void* workspace, size_t workspace_bytes,
                   void* stream);
/* img [N, in_channels, H, W] bf16 in [-1, 1] -> moments [N, 2 * latent_channels, H/8, W/8] fp32 = mean | logvar, logvar clamped
 * to [-30, 20] (DiagonalGaussianDistribution [ext]); the caller draws the sample (:234) */
int utx_vae_encode(utx_vae* h, const void* img, int N, int H, int W, float* moments, void* workspace, size_t workspace_bytes,
                   void* stream);
/* launcher calls issued by the engine since the last reset (bench instrumentation) */
long utx_vae_launches(utx_vae* h, int reset);

/* ------------------------------------------------------------------------------------------------------------------
 * The path's one collective (north_star): after the denoise loop and the VAE decode every rank contributes its
 * finished view tile(s) and receives every rank's, before UV projection.  The reference has no counterpart (it runs the
 * assets one after another on one GPU: run.py:5-10, pipeline.py:231-291).  One ncclAllGather over NVLink / NVSwitch.
 * Bootstrap like NCCL's own: rank 0 makes a 128-byte id, the host shares it out of band, every rank joins; the communicator
 * is bound to the device that is current at utx_comm_init.
 * ------------------------------------------------------------------------------------------------------------------ */
int utx_comm_unique_id(void* id128 /* host, 128 bytes out */);
int utx_comm_init(utx_comm** out, const void* id128 /* host */, int nranks, int rank);
void utx_comm_destroy(utx_comm* c);
/* send[p * bytes_per_peer ..] goes to rank p, recv[p * bytes_per_peer ..] comes from rank p (grouped ncclSend / ncclRecv) */
int utx_comm_alltoall(utx_comm* c, const void* send, void* recv, size_t bytes_per_peer, void* stream);
/* out [nranks * bytes_per_rank] (rank-major) <- every rank's tile [bytes_per_rank]; uneven shards are padded by the caller */
int utx_allgather_tiles(utx_comm* c, const void* tile, void* out, size_t bytes_per_rank, void* stream);

/* Peer memory for the fused compute + exchange kernels of the sequence-parallel mode: a device buffer allocated by one rank
 * and mapped by the others (CUDA IPC, NVLink P2P inside one box).  alloc (zero-filled) / export a 64-byte handle on the owner;
 * import / close on the peers; the host shares the handles out of band (torch.distributed.all_gather_object, MPI, ...). */
int utx_peer_alloc(void** ptr, size_t bytes);
void utx_peer_free(void* ptr);
int utx_peer_export(void* ptr, void* handle64 /* host, 64 bytes out */);
int utx_peer_import(const void* handle64 /* host */, void** ptr);
void utx_peer_close(void* ptr);
/* Bytes of the exchange region utx_flux_set_sp_peers expects for this sequence (0 = not in sequence-parallel mode). */
size_t utx_flux_sp_region_bytes(const utx_flux* h, int s_txt, int s_img);
/* Direct mode of utx_flux_set_sequence_parallel: regions[r] = rank r's exchange region as mapped into THIS process (host array
 * [nranks]; regions[rank] is this rank's own utx_peer_alloc'ed buffer).  With it the two all-to-alls around each attention
 * disappear: the QKV GEMM's epilogue stores every head's q | k | v rows straight into the owning rank's attention input over
 * NVLink, the attention kernel's epilogue stores its output rows straight into the token owner's activation buffer, and a
 * flag barrier over peer memory (one tiny kernel) replaces each collective -- the transfers overlap the tiles that are still
 * being computed.  regions = NULL returns to the NCCL all-to-all data path.  Call utx_flux_prepare again afterwards. */
int utx_flux_set_sp_peers(utx_flux* h, void* const* regions, size_t region_bytes);

#ifdef __cplusplus
}
#endif
#endif /* UNITEX_B200_H */
