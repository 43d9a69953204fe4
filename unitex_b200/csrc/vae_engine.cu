// FLUX AutoencoderKL encode / decode orchestrated on the host in C++ over the sm_100a kernels (implicit-GEMM 3x3 convolutions
// on the tcgen05 GEMM, GroupNorm+SiLU, chunked single-head mid attention).  Replaces `self.vae.encode(...)` /
// `self.vae.decode(...)` of the reference sampler (flux_piplines/texturing/pipeline.py:226-238, :688-692; diffusers
// AutoencoderKL [ext], FLUX config: latent 16, blocks (128,256,512,512), 2 layers per block, GroupNorm 32, one mid-block
// attention head of 512 channels, no quant convs).  Round 1 drove the same kernels from Python (~160 ctypes calls and as many
// torch allocations per decode); this file is the C ABI `utx_vae_*` of SURVEY 8b.
//
// Memory: one caller-provided workspace -- four activation slots of the largest NHWC tensor of the pass, an im2col buffer
// (only conv_in and the encoder's stride-2 convolutions use it), the attention scratch and the GroupNorm statistics.  Sizes
// come from a dry run of the same traversal (`Run::dry`), so the layout can never disagree with the execution.
//
// Mid attention (HW x HW scores, one 512-wide head): scores are produced, normalised and consumed in ROW CHUNKS that bound the
// scratch at 384 MB (fp32 scores + bf16 probabilities of 4096 rows at 1024^2) instead of materialising the whole 16 384^2
// fp32 matrix (1.07 GB + 0.5 GB at 1024^2; 2.4 + 1.2 GB at the reference's 512 x 3072 strip) as round 1 did.  Chunks are kept
// LARGE on purpose: the P @ V product of a chunk is an [R, 512] x K = HW GEMM with only R/256 x 4 output tiles, and with
// L2-sized chunks of 512 rows it ran on 4 of the 74 CTA pairs (82 us per chunk, 5.2 ms per decode for 0.55 TFLOP:
// profiles/r02_vae_launches.csv); the extra HBM traffic of a chunk that does not fit L2 is ~0.5 ms per decode.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/unitex_b200.h"
#include "common.h"
#include "kernels.h"

using namespace utx;

struct utx_vae {
  utx_vae_config cfg;
  utx_vae_weights w;
  std::vector<utx_vae_resnet> enc_res, dec_res;
  std::vector<const void*> enc_down_w, enc_down_b, dec_up_w, dec_up_b;
  bool has_weights = false;
  long launches = 0;
};

namespace utx {
// vae_ops.cu
int fill_f32(float* p, float v, long long n, cudaStream_t stream);
int nchw_to_nhwc_bf16(const bf16* x, int N, int C, int H, int W, bf16* y, cudaStream_t stream);
int nhwc_to_nchw_bf16(const bf16* x, long ldx, int N, int C, int H, int W, bf16* y, cudaStream_t stream);
int moments_to_nchw_f32(const bf16* x, long ldx, int N, int C2, int H, int W, float lo, float hi, float* y, cudaStream_t stream);
}  // namespace utx

namespace {

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
inline int ceil_to(int v, int m) { return (v + m - 1) / m * m; }
constexpr size_t kAttnChunkBytes = 384u << 20;   // fp32 scores + bf16 probabilities of one row chunk

struct Layout {
  size_t act = 0;       // bytes of one activation slot
  size_t col = 0;       // im2col buffer
  size_t attn = 0;      // q, k, v, v^T, o, score chunk, probability chunk
  size_t stats = 0;     // GroupNorm workspace
  size_t ones = 4096 * 4;   // fp32 ones (gate of the fused residual epilogue), >= the widest channel count
  size_t total() const { return 4 * align_up(act) + align_up(col) + align_up(attn) + align_up(stats) + align_up(ones); }
};

// One pass over the network.  dry = true only records the sizes the pass needs.
struct Run {
  utx_vae* h;
  bool dry;
  cudaStream_t st;
  Layout L;
  // live buffers (dry == false)
  bf16* slot[4] = {nullptr, nullptr, nullptr, nullptr};
  bool used[4] = {false, false, false, false};
  bf16* col = nullptr;
  uint8_t* attn = nullptr;
  double* stats = nullptr;
  float* ones = nullptr;
  int err = 0;

  bf16* acquire(size_t elems) {
    if (dry) {
      L.act = std::max(L.act, elems * 2);
      return nullptr;
    }
    for (int i = 0; i < 4; ++i)
      if (!used[i]) {
        used[i] = true;
        return slot[i];
      }
    set_error("vae: activation slots exhausted");
    err = 1;
    return slot[0];
  }
  void release(const bf16* p) {
    if (dry) return;
    for (int i = 0; i < 4; ++i)
      if (slot[i] == p) used[i] = false;
  }
  void chk(int r) {
    if (r != 0 && err == 0) err = r;
    if (!dry) h->launches++;
  }

  // GroupNorm(G, eps 1e-6, affine) [+ SiLU] into a fresh slot
  bf16* gn(const bf16* x, int N, int HW, int C, const float* gamma, const float* beta, bool silu) {
    bf16* y = acquire(static_cast<size_t>(N) * HW * C);
    L.stats = std::max(L.stats, groupnorm_workspace_bytes(N, HW, C, h->cfg.norm_num_groups));
    if (!dry && !err) chk(groupnorm_nhwc(x, y, N, HW, C, h->cfg.norm_num_groups, gamma, beta, silu ? 1 : 0, stats, st));
    return y;
  }

  // 3x3 convolution.  up = 2: nearest 2x upsampling first (Upsample2D); stride 2 / pad 0: Downsample2D's asymmetric (0,1,0,1)
  // padding.  res != nullptr: y = res + conv(x) written over res (residual fused in the GEMM epilogue).  Returns y, sets Ho / Wo.
  bf16* conv3(const bf16* x, int N, int H, int W, int Cin, const void* w, const void* b, int Cout, int up, int stride, int pad,
              bf16* res, int* Ho_out, int* Wo_out) {
    const int Cop = ceil_to(Cout, 8), Kp = ceil_to(9 * Cin, 64);
    int Hs = H * up, Ws = W * up;
    const int Ho = stride == 1 ? Hs : (Hs + 1 - 3) / 2 + 1, Wo = stride == 1 ? Ws : (Ws + 1 - 3) / 2 + 1;
    *Ho_out = Ho;
    *Wo_out = Wo;
    const bool implicit_ok = stride == 1 && pad == 1 && Cin % 64 == 0 && Kp == 9 * Cin &&
                             (Ws % 128 == 0 || (128 % Ws == 0 && Hs % (128 / Ws) == 0));
    const bf16* xin = x;
    bf16* xu = nullptr;
    if (implicit_ok && up == 2) {   // materialise the upsampled tensor (4x the input, not the 36x of an im2col buffer)
      xu = acquire(static_cast<size_t>(N) * Hs * Ws * Cin);
      if (!dry && !err) chk(upsample2x_nhwc(x, N, H, W, Cin, xu, st));
      xin = xu;
    }
    bf16* y = res ? res : acquire(static_cast<size_t>(N) * Ho * Wo * Cop);
    L.ones = std::max(L.ones, static_cast<size_t>(Cop) * 4);
    if (implicit_ok) {
      if (!dry && !err)
        chk(conv3x3_nhwc(xin, N, Hs, Ws, Cin, static_cast<const bf16*>(w), static_cast<const bf16*>(b), Cop, y, Cop,
                         res ? ones : nullptr, res, res ? Cop : 0, st));
    } else {
      L.col = std::max(L.col, static_cast<size_t>(N) * Ho * Wo * Kp * 2);
      if (!dry && !err) {
        chk(im2col3x3(x, N, H, W, Cin, up, stride, pad, Ho, Wo, Kp, col, st));
        chk(gemm(col, Kp, w, Kp, b, y, Cop, N * Ho * Wo, Cop, Kp, res));
      }
    }
    if (xu) release(xu);
    return y;
  }

  int gemm(const bf16* A, long lda, const void* W, long ldw, const void* bias, bf16* C, long ldc, int M, int Nn, int K,
           const bf16* res) {
    GemmArgs a{};
    a.N = Nn; a.K = K; a.epi = res ? EPI_GATE_RES : EPI_BIAS; a.nprob = 1;
    a.prob[0] = GemmProblem{A, lda, static_cast<const bf16*>(W), ldw, M, C, ldc, static_cast<const bf16*>(bias), res ? ones : nullptr,
                            res, ldc, 0, nullptr, 0, nullptr, nullptr, 0};
    return gemm_bf16_tn(a, st);
  }

  // ResnetBlock2D [ext]: x + conv2(silu(gn2(conv1(silu(gn1(x)))))), 1x1 conv_shortcut when the channel count changes.
  // Consumes x (its slot is released or becomes the output).
  bf16* resnet(bf16* x, int N, int H, int W, const utx_vae_resnet& r) {
    const int Cin = r.cin, Cout = r.cout;
    int ho, wo;
    bf16* h1 = gn(x, N, H * W, Cin, r.gn1_w, r.gn1_b, true);
    bf16* h2 = conv3(h1, N, H, W, Cin, r.conv1_w, r.conv1_b, Cout, 1, 1, 1, nullptr, &ho, &wo);
    release(h1);
    bf16* h3 = gn(h2, N, H * W, Cout, r.gn2_w, r.gn2_b, true);
    release(h2);
    bf16* sc = x;
    if (r.short_w) {   // 1x1 convolution == Linear over channels
      sc = acquire(static_cast<size_t>(N) * H * W * Cout);
      if (!dry && !err) chk(gemm(x, Cin, r.short_w, Cin, r.short_b, sc, Cout, N * H * W, Cout, Cin, nullptr));
      release(x);
    }
    bf16* y = conv3(h3, N, H, W, Cout, r.conv2_w, r.conv2_b, Cout, 1, 1, 1, sc, &ho, &wo);
    release(h3);
    return y;
  }

  // Attention block of the mid block [ext Attention(heads = 1, dim_head = C, residual_connection)]: x + to_out(softmax(q k^T /
  // sqrt(C)) v) with q, k, v = Linear(GroupNorm(x)).  In place on x.
  void attention(bf16* x, int N, int HW, int C, const utx_vae_attn& a) {
    // scratch: q, k, v [HW, C], v^T [C, HW], o [HW, C]; score chunk fp32 [R, HW], probability chunk bf16 [R, HW]
    int R = static_cast<int>(kAttnChunkBytes / (static_cast<size_t>(HW) * 6));
    R = std::max(128, R / 128 * 128);
    R = std::min(R, ceil_to(HW, 128));
    const size_t mat = align_up(static_cast<size_t>(HW) * C * 2);
    const size_t need = 5 * mat + align_up(static_cast<size_t>(R) * HW * 4) + align_up(static_cast<size_t>(R) * HW * 2);
    L.attn = std::max(L.attn, need);
    for (int n = 0; n < N; ++n) {
      bf16* xs = x + static_cast<size_t>(n) * HW * C;
      bf16* hn = gn(xs, 1, HW, C, a.gn_w, a.gn_b, false);
      if (!dry && !err) {
        bf16* q = reinterpret_cast<bf16*>(attn);
        bf16* k = reinterpret_cast<bf16*>(attn + mat);
        bf16* v = reinterpret_cast<bf16*>(attn + 2 * mat);
        bf16* vt = reinterpret_cast<bf16*>(attn + 3 * mat);
        bf16* o = reinterpret_cast<bf16*>(attn + 4 * mat);
        float* S = reinterpret_cast<float*>(attn + 5 * mat);
        bf16* P = reinterpret_cast<bf16*>(attn + 5 * mat + align_up(static_cast<size_t>(R) * HW * 4));
        chk(gemm(hn, C, a.wq, C, a.bq, q, C, HW, C, C, nullptr));
        chk(gemm(hn, C, a.wk, C, a.bk, k, C, HW, C, C, nullptr));
        chk(gemm(hn, C, a.wv, C, a.bv, v, C, HW, C, C, nullptr));
        chk(transpose_bf16(v, C, vt, HW, HW, C, st));
        const float scale = 1.0f / sqrtf(static_cast<float>(C));
        for (int r0 = 0; r0 < HW && !err; r0 += R) {
          const int rows = std::min(R, HW - r0);
          GemmArgs g{};
          g.N = HW; g.K = C; g.epi = EPI_BIAS_F32; g.out_scale = scale; g.nprob = 1;
          g.prob[0] = GemmProblem{q + static_cast<size_t>(r0) * C, C, k, C, rows, reinterpret_cast<bf16*>(S), HW, nullptr, nullptr,
                                  nullptr, 0, 0, nullptr, 0, nullptr, nullptr, 0};
          chk(gemm_bf16_tn(g, st));
          chk(softmax_rows(S, HW, P, HW, rows, HW, st));
          chk(gemm(P, HW, vt, HW, nullptr, o + static_cast<size_t>(r0) * C, C, rows, C, HW, nullptr));
        }
        chk(gemm(o, C, a.wo, C, a.bo, xs, C, HW, C, C, xs));   // residual in place
      }
      release(hn);
    }
  }

  bf16* mid(bf16* x, int N, int H, int W, int C, const utx_vae_mid& m) {
    x = resnet(x, N, H, W, m.res0);
    attention(x, N, H * W, C, m.attn);
    return resnet(x, N, H, W, m.res1);
  }
};

int bind(Run& r, void* workspace, size_t workspace_bytes) {
  const Layout& L = r.L;
  UTX_CHECK(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "utx_vae: workspace must be 256B aligned");
  UTX_CHECK(workspace_bytes >= L.total(), "utx_vae: workspace too small");
  uint8_t* b = static_cast<uint8_t*>(workspace);
  for (int i = 0; i < 4; ++i) {
    r.slot[i] = reinterpret_cast<bf16*>(b);
    b += align_up(L.act);
  }
  r.col = reinterpret_cast<bf16*>(b);
  b += align_up(L.col);
  r.attn = b;
  b += align_up(L.attn);
  r.stats = reinterpret_cast<double*>(b);
  b += align_up(L.stats);
  r.ones = reinterpret_cast<float*>(b);
  UTX_TRY(fill_f32(r.ones, 1.0f, static_cast<long long>(L.ones / 4), r.st));
  return 0;
}

// z [N, Cz, h, w] NCHW bf16 -> img [N, 3, 8h, 8w] NCHW bf16
int decode_pass(Run& r, const bf16* z, int N, int H, int W, bf16* img) {
  utx_vae* h = r.h;
  const utx_vae_config& c = h->cfg;
  const int nb = c.num_blocks, Cz = c.latent_channels;
  int Ctop = c.block_out_channels[nb - 1];
  bf16* x0 = r.acquire(static_cast<size_t>(N) * H * W * Cz);
  if (!r.dry) r.chk(nchw_to_nhwc_bf16(z, N, Cz, H, W, x0, r.st));
  int ho, wo;
  bf16* x = r.conv3(x0, N, H, W, Cz, h->w.dec_conv_in_w, h->w.dec_conv_in_b, Ctop, 1, 1, 1, nullptr, &ho, &wo);
  r.release(x0);
  x = r.mid(x, N, H, W, Ctop, h->w.dec_mid);
  int C = Ctop;
  for (int i = 0; i < nb; ++i) {
    for (int j = 0; j < c.layers_per_block + 1; ++j) {
      const utx_vae_resnet& rs = h->dec_res[i * (c.layers_per_block + 1) + j];
      x = r.resnet(x, N, H, W, rs);
      C = rs.cout;
    }
    if (i < nb - 1) {
      bf16* y = r.conv3(x, N, H, W, C, h->dec_up_w[i], h->dec_up_b[i], C, 2, 1, 1, nullptr, &ho, &wo);
      r.release(x);
      x = y;
      H = ho;
      W = wo;
    }
  }
  bf16* xn = r.gn(x, N, H * W, C, h->w.dec_norm_w, h->w.dec_norm_b, true);
  r.release(x);
  bf16* y = r.conv3(xn, N, H, W, C, h->w.dec_conv_out_w, h->w.dec_conv_out_b, c.in_channels, 1, 1, 1, nullptr, &ho, &wo);
  r.release(xn);
  if (!r.dry && !r.err) r.chk(nhwc_to_nchw_bf16(y, ceil_to(c.in_channels, 8), N, c.in_channels, H, W, img, r.st));
  r.release(y);
  return r.err;
}

// img [N, 3, H, W] NCHW bf16 in [-1, 1] -> moments [N, 2 Cz, H/8, W/8] NCHW fp32 (mean | logvar clamped to [-30, 20])
int encode_pass(Run& r, const bf16* img, int N, int H, int W, float* moments) {
  utx_vae* h = r.h;
  const utx_vae_config& c = h->cfg;
  const int nb = c.num_blocks;
  bf16* x0 = r.acquire(static_cast<size_t>(N) * H * W * c.in_channels);
  if (!r.dry) r.chk(nchw_to_nhwc_bf16(img, N, c.in_channels, H, W, x0, r.st));
  int ho, wo;
  int C = c.block_out_channels[0];
  bf16* x = r.conv3(x0, N, H, W, c.in_channels, h->w.enc_conv_in_w, h->w.enc_conv_in_b, C, 1, 1, 1, nullptr, &ho, &wo);
  r.release(x0);
  for (int i = 0; i < nb; ++i) {
    for (int j = 0; j < c.layers_per_block; ++j) {
      const utx_vae_resnet& rs = h->enc_res[i * c.layers_per_block + j];
      x = r.resnet(x, N, H, W, rs);
      C = rs.cout;
    }
    if (i < nb - 1) {
      bf16* y = r.conv3(x, N, H, W, C, h->enc_down_w[i], h->enc_down_b[i], C, 1, 2, 0, nullptr, &ho, &wo);
      r.release(x);
      x = y;
      H = ho;
      W = wo;
    }
  }
  x = r.mid(x, N, H, W, C, h->w.enc_mid);
  bf16* xn = r.gn(x, N, H * W, C, h->w.enc_norm_w, h->w.enc_norm_b, true);
  r.release(x);
  bf16* m = r.conv3(xn, N, H, W, C, h->w.enc_conv_out_w, h->w.enc_conv_out_b, 2 * c.latent_channels, 1, 1, 1, nullptr, &ho, &wo);
  r.release(xn);
  if (!r.dry && !r.err)
    r.chk(moments_to_nchw_f32(m, ceil_to(2 * c.latent_channels, 8), N, 2 * c.latent_channels, H, W, -30.0f, 20.0f, moments, r.st));
  r.release(m);
  return r.err;
}

int check_dims(const utx_vae* h, int N, int H, int W, int decode) {
  UTX_CHECK(h && h->has_weights, "utx_vae: weights not set");
  UTX_CHECK(N > 0 && H > 0 && W > 0, "utx_vae: empty input");
  const int f = 1 << (h->cfg.num_blocks - 1);
  if (!decode) UTX_CHECK(H % f == 0 && W % f == 0, "utx_vae_encode: image size must be a multiple of the down-sampling factor");
  return 0;
}

}  // namespace

extern "C" {

int utx_vae_create(const utx_vae_config* cfg, utx_vae** out) {
  UTX_CHECK(cfg && out, "utx_vae_create: null argument");
  UTX_CHECK(cfg->num_blocks >= 1 && cfg->num_blocks <= 8, "utx_vae_create: 1..8 blocks");
  UTX_CHECK(cfg->layers_per_block >= 1 && cfg->norm_num_groups > 0, "utx_vae_create: bad layer / group counts");
  UTX_CHECK(cfg->in_channels > 0 && cfg->in_channels <= 8 && cfg->latent_channels % 8 == 0, "utx_vae_create: in_channels <= 8, latent_channels % 8 == 0");
  for (int i = 0; i < cfg->num_blocks; ++i)
    UTX_CHECK(cfg->block_out_channels[i] % 64 == 0 && cfg->block_out_channels[i] % cfg->norm_num_groups == 0,
              "utx_vae_create: block_out_channels must be multiples of 64 and of norm_num_groups");
  utx_vae* h = new utx_vae();
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void utx_vae_destroy(utx_vae* h) { delete h; }

int utx_vae_set_weights(utx_vae* h, const utx_vae_weights* w) {
  UTX_CHECK(h && w, "utx_vae_set_weights: null argument");
  const utx_vae_config& c = h->cfg;
  UTX_CHECK(w->enc_res && w->dec_res && (c.num_blocks == 1 || (w->enc_down_w && w->enc_down_b && w->dec_up_w && w->dec_up_b)),
            "utx_vae_set_weights: null block table");
  h->w = *w;
  h->enc_res.assign(w->enc_res, w->enc_res + c.num_blocks * c.layers_per_block);
  h->dec_res.assign(w->dec_res, w->dec_res + c.num_blocks * (c.layers_per_block + 1));
  h->enc_down_w.assign(w->enc_down_w, w->enc_down_w + (c.num_blocks - 1));
  h->enc_down_b.assign(w->enc_down_b, w->enc_down_b + (c.num_blocks - 1));
  h->dec_up_w.assign(w->dec_up_w, w->dec_up_w + (c.num_blocks - 1));
  h->dec_up_b.assign(w->dec_up_b, w->dec_up_b + (c.num_blocks - 1));
  h->has_weights = true;
  return 0;
}

size_t utx_vae_workspace_bytes(utx_vae* h, int N, int H, int W, int decode) {
  if (check_dims(h, N, H, W, decode) != 0) return 0;
  Run r{h, true, nullptr};
  if (decode) decode_pass(r, nullptr, N, H, W, nullptr);
  else encode_pass(r, nullptr, N, H, W, nullptr);
  return r.L.total();
}

int utx_vae_decode(utx_vae* h, const void* z, int N, int H, int W, void* img, void* workspace, size_t workspace_bytes,
                   void* stream) {
  UTX_TRY(check_dims(h, N, H, W, 1));
  UTX_CHECK(z && img, "utx_vae_decode: null pointer");
  Run dry{h, true, nullptr};
  decode_pass(dry, nullptr, N, H, W, nullptr);
  Run r{h, false, static_cast<cudaStream_t>(stream)};
  r.L = dry.L;
  UTX_TRY(bind(r, workspace, workspace_bytes));
  return decode_pass(r, static_cast<const bf16*>(z), N, H, W, static_cast<bf16*>(img));
}

int utx_vae_encode(utx_vae* h, const void* img, int N, int H, int W, float* moments, void* workspace, size_t workspace_bytes,
                   void* stream) {
  UTX_TRY(check_dims(h, N, H, W, 0));
  UTX_CHECK(img && moments, "utx_vae_encode: null pointer");
  Run dry{h, true, nullptr};
  encode_pass(dry, nullptr, N, H, W, nullptr);
  Run r{h, false, static_cast<cudaStream_t>(stream)};
  r.L = dry.L;
  UTX_TRY(bind(r, workspace, workspace_bytes));
  return encode_pass(r, static_cast<const bf16*>(img), N, H, W, moments);
}

long utx_vae_launches(utx_vae* h, int reset) {
  if (!h) return 0;
  const long n = h->launches;
  if (reset) h->launches = 0;
  return n;
}

}  // extern "C"
