// FLUX.1-dev MM-DiT forward + flow-match Euler loop, orchestrated on the host in C++ over the sm_100a kernels.
// Replaces diffusers' FluxTransformer2DModel.forward as called at flux_piplines/texturing/pipeline.py:646-656 and the
// loop :634-681.  Token layout in the residual buffer: [txt rows | img rows] from the start, so the double-stream
// blocks address the two streams as row ranges of one buffer (grouped GEMM launches) and the torch.cat before the
// single-stream blocks costs nothing.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/unitex_b200.h"
#include "common.h"
#include "kernels.h"

using namespace utx;

namespace utx {   // comm.cu
int comm_nranks(const utx_comm* c);
int comm_rank(const utx_comm* c);
int comm_alltoall(utx_comm* c, const void* send, void* recv, size_t bytes_per_peer, cudaStream_t stream);
int comm_allgather(utx_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream);
int peer_barrier(unsigned* const* flags_dev, unsigned* local_flags, int rank, int nranks, unsigned epoch, cudaStream_t stream);
// elementwise.cu: cat[row, p * w + j] = recv[p][row][j]  (the received attention heads of peer p into their columns)
int sp_unpack_heads(const bf16* recv, bf16* cat, long ld_cat, int rows, int w, int npeers, cudaStream_t stream);
}  // namespace utx

struct utx_flux {
  utx_flux_config cfg;
  utx_flux_weights w;
  std::vector<utx_double_block> dbl;
  std::vector<utx_single_block> sgl;
  bool has_weights = false;
  // sequence-parallel ("Ulysses") mode: ONE grid's tokens split over the ranks of sp_comm (nullptr: off, everything local)
  utx_comm* sp_comm = nullptr;
  int sp_n = 1, sp_rank = 0;
  // per-call state (set by prepare).  Rows this rank owns: global rows [r0, r0 + S_loc) of the [txt | img] sequence, of which
  // the first st_loc are text rows and the other si_loc image rows (image tokens img0 ...); without sequence parallelism
  // r0 = 0, S_loc = s_txt + s_img, st_loc = s_txt, si_loc = s_img.
  int r0 = 0, S_loc = 0, st_loc = 0, si_loc = 0, img0 = 0;
  bf16 *qkv_all = nullptr, *attn_all = nullptr, *attn_recv = nullptr, *v_loc = nullptr, *v_all = nullptr;
  // direct mode: every rank's exchange region (cat | qkv_all | flags) mapped into this process; the epilogues of the QKV GEMM
  // and of the attention kernel store into the peers' regions, a flag barrier replaces each all-to-all
  std::vector<uint8_t*> sp_regions;
  size_t sp_region_bytes = 0;
  unsigned** sp_flags_dev = nullptr;   // device array [sp_n]: every rank's flag array
  unsigned sp_epoch = 0;
  size_t sp_off_cat = 0, sp_off_qkv = 0, sp_off_flags = 0;
  int s_txt = 0, s_img = 0;
  bf16 *x = nullptr, *xn = nullptr, *qkv = nullptr, *cat = nullptr, *ctx0 = nullptr, *v_tmp = nullptr;
  float *cos_t = nullptr, *sin_t = nullptr, *mod = nullptr, *temb = nullptr, *sincos = nullptr, *hvec = nullptr,
        *pooled = nullptr;
  bool prepared = false;
  // instrumentation: launches per category and (when profiling) CUDA-event time per category
  bool profile = false;
  long launches[UTX_PROF_NCAT] = {0, 0, 0, 0};
  float prof_ms[UTX_PROF_NCAT] = {0, 0, 0, 0};
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_used;   // (category, index of start event); stop = start + 1
  size_t ev_next = 0;
  // CUDA graph of one denoise step (forward + Euler update), captured once per (latents, s_noise) and replayed every step:
  // the step's scalars (t, guidance, sigma difference) live in device memory (step_params) and are written by a one-thread
  // kernel just before each launch, so ONE graph serves all 28 steps.  Capture runs on a stream the engine owns (the legacy
  // default stream cannot capture); the instantiated graph is launched on the caller's stream.
  struct StepGraph {
    void* latents = nullptr;
    int s_noise = 0;
    int seen = 0;                 // eager calls seen with this key (the graph is built on the second one)
    cudaGraphExec_t exec = nullptr;
    long launches[UTX_PROF_NCAT] = {0, 0, 0, 0};
  };
  std::vector<StepGraph> graphs;
  cudaStream_t capture_stream = nullptr;
  float* step_params = nullptr;   // device [4]: t_eff, g_eff, dsigma
  int use_graph = -1;             // -1 = not read yet (UTX_FLUX_GRAPH, default on)
  long graph_replays = 0;
};

namespace {

// Runs one launcher under a category: counts the launch and, when profiling, brackets it with events on `st`.
template <class F>
int run_cat(utx_flux* h, int cat, cudaStream_t st, F&& f) {
  h->launches[cat]++;
  if (!h->profile) return f();
  if (h->ev_next + 2 > h->ev_pool.size()) {
    for (int i = 0; i < 256; ++i) {
      cudaEvent_t e;
      UTX_CUDA(cudaEventCreate(&e));
      h->ev_pool.push_back(e);
    }
  }
  const size_t i0 = h->ev_next;
  h->ev_next += 2;
  UTX_CUDA(cudaEventRecord(h->ev_pool[i0], st));
  int r = f();
  UTX_CUDA(cudaEventRecord(h->ev_pool[i0 + 1], st));
  h->ev_used.emplace_back(cat, static_cast<int>(i0));
  return r;
}
#define CAT_GEMM(expr) UTX_TRY(run_cat(h, UTX_PROF_GEMM, st, [&] { return (expr); }))
#define CAT_ATTN(expr) UTX_TRY(run_cat(h, UTX_PROF_ATTN, st, [&] { return (expr); }))
#define CAT_ELEM(expr) UTX_TRY(run_cat(h, UTX_PROF_ELEM, st, [&] { return (expr); }))
#define CAT_OTHER(expr) UTX_TRY(run_cat(h, UTX_PROF_OTHER, st, [&] { return (expr); }))

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
inline int D_of(const utx_flux_config& c) { return c.num_heads * c.head_dim; }
inline long n_mod_rows(const utx_flux_config& c) {
  return static_cast<long>(12 * c.num_layers + 3 * c.num_single_layers + 2) * D_of(c);
}
// round-to-nearest-even fp32 -> bf16 -> fp32, on the host (mirrors torch's .to(torch.bfloat16))
inline float bf16_round(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return f;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  std::memcpy(&f, &u, 4);
  return f;
}

struct WsLayout {
  size_t x, xn, qkv, cat, ctx0, v_tmp, cos_t, sin_t, mod, temb, sincos, hvec, pooled, step_params;
  size_t qkv_all, attn_all, attn_recv, v_loc, v_all;   // sequence-parallel mode only
  size_t total;
};
// rows per rank: S / P (P = 1 without sequence parallelism)
WsLayout ws_layout(const utx_flux_config& c, int s_txt, int s_img, int P) {
  const size_t S = static_cast<size_t>(s_txt) + s_img, D = D_of(c);
  const size_t Sl = S / P;
  WsLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  L.x = take(Sl * D * 2);
  L.xn = take(Sl * D * 2);
  L.qkv = take(Sl * 3 * D * 2);
  L.cat = take(Sl * (1 + c.mlp_ratio) * D * 2);
  L.ctx0 = take(std::min<size_t>(s_txt, Sl) * D * 2);
  L.v_tmp = take(static_cast<size_t>(s_img) * c.in_channels * 2);
  L.cos_t = take(S * 128 * 4);
  L.sin_t = take(S * 128 * 4);
  L.mod = take(n_mod_rows(c) * 4);
  L.temb = take(D * 4);
  L.sincos = take(512 * 4);
  L.hvec = take(D * 4);
  L.pooled = take(static_cast<size_t>(c.pooled_projection_dim) * 4);
  L.step_params = take(16);
  if (P > 1) {
    L.qkv_all = take(Sl * 3 * D * 2);       // [S, 3 * (H/P) * 128]: every token, this rank's heads
    L.attn_all = take(Sl * D * 2);          // [S, (H/P) * 128]
    L.attn_recv = take(Sl * D * 2);         // [P][S_loc][(H/P) * 128]
    L.v_loc = take(Sl * c.in_channels * 2);
    L.v_all = take(S * c.in_channels * 2);
  }
  L.total = off;
  return L;
}

// direct mode: the q | k | v scatter of `p` targets the peers' attention inputs; rows are then global (r0 + local row)
void set_peers(const utx_flux* h, GemmProblem& p, int local_row_base) {
  if (h->sp_regions.empty()) return;
  for (int r = 0; r < h->sp_n; ++r) p.sc_peer[r] = reinterpret_cast<bf16*>(h->sp_regions[r] + h->sp_off_qkv);
  p.sc_row_base = h->r0 + local_row_base;
}

// exchange region of the direct mode: [cat (S_loc x 5D) | qkv_all (S x 3 * (H/P) * 128) | flags]
struct RegionLayout {
  size_t cat, qkv_all, flags, total;
};
RegionLayout region_layout(const utx_flux_config& c, int s_txt, int s_img, int P) {
  const size_t S = static_cast<size_t>(s_txt) + s_img, D = D_of(c), Sl = S / P;
  RegionLayout R{};
  size_t off = 0;
  R.cat = off; off += align_up(Sl * (1 + c.mlp_ratio) * D * 2);
  R.qkv_all = off; off += align_up(Sl * 3 * D * 2);
  R.flags = off; off += align_up(64 * sizeof(unsigned));
  R.total = off;
  return R;
}

int gemm1(const bf16* A, long lda, const bf16* W, long ldw, const bf16* bias, bf16* C, long ldc, int M, int N, int K,
          int epi, const float* gate, const bf16* res, long ldres, cudaStream_t st, int gelu_start = 0,
          int split_col = 0, bf16* C2 = nullptr, long ldc2 = 0, int qk_cols = 0, const bf16* wq = nullptr,
          const bf16* wk = nullptr, const float* cos_t = nullptr, const float* sin_t = nullptr, int row_offset = 0,
          int sc_hl = 0, int sc_rows = 0, int sc_D = 0, const utx_flux* peers_of = nullptr) {
  GemmArgs a{};
  a.N = N; a.K = K; a.epi = epi; a.gelu_col_start = gelu_start; a.nprob = 1;
  a.qk_cols = qk_cols; a.cos_t = cos_t; a.sin_t = sin_t;
  a.prob[0] = GemmProblem{A, lda, W, ldw, M, C, ldc, bias, gate, res, ldres, split_col, C2, ldc2, wq, wk, row_offset,
                          sc_hl, sc_rows, 0, sc_D};
  if (peers_of) set_peers(peers_of, a.prob[0], 0);
  return gemm_bf16_tn(a, st);
}

// this rank's txt rows [0, st_loc) with the *_txt weights and img rows [st_loc, S_loc) with the *_img weights, one launch.
// scatter: the q | k | v columns go to the all-to-all send layout (sequence-parallel mode), C is then the send buffer.
int gemm_streams(const utx_flux* h, const bf16* A, long lda, const void* W_txt, const void* b_txt, const void* W_img,
                 const void* b_img, long ldw, bf16* C, long ldc, int N, int K, int epi, const float* gate_txt,
                 const float* gate_img, const bf16* res, long ldres, cudaStream_t st, int qk_cols = 0,
                 const void* wq_txt = nullptr, const void* wk_txt = nullptr, const void* wq_img = nullptr,
                 const void* wk_img = nullptr, bool scatter = false) {
  GemmArgs a{};
  a.N = N; a.K = K; a.epi = epi; a.gelu_col_start = 0; a.nprob = 2;
  a.qk_cols = qk_cols; a.cos_t = h->cos_t; a.sin_t = h->sin_t;
  const long st_rows = h->st_loc;
  const int hl = scatter ? h->cfg.num_heads / h->sp_n : 0, D = D_of(h->cfg);
  a.prob[0] = GemmProblem{A, lda, static_cast<const bf16*>(W_txt), ldw, h->st_loc, C, ldc,
                          static_cast<const bf16*>(b_txt), gate_txt, res, ldres, 0, nullptr, 0,
                          static_cast<const bf16*>(wq_txt), static_cast<const bf16*>(wk_txt), h->r0, hl, h->S_loc, 0, D};
  a.prob[1] = GemmProblem{A + st_rows * lda, lda, static_cast<const bf16*>(W_img), ldw, h->si_loc,
                          scatter ? C : C + st_rows * ldc, ldc,
                          static_cast<const bf16*>(b_img), gate_img, res ? res + st_rows * ldres : nullptr, ldres, 0,
                          nullptr, 0, static_cast<const bf16*>(wq_img), static_cast<const bf16*>(wk_img), h->r0 + h->st_loc,
                          hl, h->S_loc, h->st_loc, D};
  if (scatter) {
    set_peers(h, a.prob[0], 0);
    set_peers(h, a.prob[1], h->st_loc);
  }
  return gemm_bf16_tn(a, st);
}

// joint attention of one block.  Local: qkv [S, 3D] -> cat[:, :D].  Sequence-parallel: the QKV GEMM left every peer's heads in
// the send layout; all-to-all -> this rank holds ALL tokens of ITS heads -> attention -> all-to-all back -> the heads of all
// peers for THIS rank's tokens, unpacked into cat[:, :D].
int attention_block(utx_flux* h, cudaStream_t st) {
  const int D = D_of(h->cfg), H = h->cfg.num_heads;
  const long ldc5 = 5L * D;
  if (h->sp_n == 1) {
    CAT_ATTN(attention_bf16(h->qkv, 3L * D, h->cat, ldc5, h->S_loc, H, st));
    return 0;
  }
  const int P = h->sp_n, Hl = H / P, S = h->S_loc * P;
  const size_t w = static_cast<size_t>(Hl) * 128;
  if (!h->sp_regions.empty()) {
    // direct mode: the QKV GEMM has stored this rank's rows of every head into the head owner's qkv_all; wait until every peer
    // has done the same here, run the attention of MY heads over the whole sequence with an epilogue that stores each output row
    // into the row owner's cat[:, my head columns], and wait again before the out-projection reads the local cat
    unsigned* local_flags = reinterpret_cast<unsigned*>(h->sp_regions[h->sp_rank] + h->sp_off_flags);
    CAT_OTHER(peer_barrier(h->sp_flags_dev, local_flags, h->sp_rank, P, ++h->sp_epoch, st));
    AttnScatter sc{};
    for (int r = 0; r < P; ++r) sc.base[r] = reinterpret_cast<bf16*>(h->sp_regions[r] + h->sp_off_cat);
    sc.rows_per_rank = h->S_loc;
    sc.ld = ldc5;
    sc.col0 = h->sp_rank * static_cast<int>(w);
    CAT_ATTN(attention_bf16(h->qkv_all, 3L * w, nullptr, ldc5, S, Hl, st, &sc));
    CAT_OTHER(peer_barrier(h->sp_flags_dev, local_flags, h->sp_rank, P, ++h->sp_epoch, st));
    return 0;
  }
  CAT_OTHER(comm_alltoall(h->sp_comm, h->qkv, h->qkv_all, static_cast<size_t>(h->S_loc) * 3 * w * 2, st));
  CAT_ATTN(attention_bf16(h->qkv_all, 3L * w, h->attn_all, static_cast<long>(w), S, Hl, st));
  CAT_OTHER(comm_alltoall(h->sp_comm, h->attn_all, h->attn_recv, static_cast<size_t>(h->S_loc) * w * 2, st));
  CAT_ELEM(sp_unpack_heads(h->attn_recv, h->cat, ldc5, h->S_loc, static_cast<int>(w), P, st));
  return 0;
}

void drop_graphs(utx_flux* h) {
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

int forward_impl(utx_flux* h, const void* latents, float t_eff, float g_eff, const float* tg_dev, void* v_out, cudaStream_t st);

}  // namespace

extern "C" {

int utx_flux_create(const utx_flux_config* cfg, utx_flux** out) {
  UTX_CHECK(cfg && out, "utx_flux_create: null argument");
  UTX_CHECK(cfg->head_dim == 128, "utx_flux_create: head_dim must be 128");
  UTX_CHECK(cfg->num_heads > 0 && (cfg->num_heads * 128) % 256 == 0, "utx_flux_create: num_heads must be even");
  UTX_CHECK(cfg->in_channels % 64 == 0 && cfg->joint_attention_dim % 64 == 0 && cfg->pooled_projection_dim % 8 == 0,
            "utx_flux_create: in_channels/joint_attention_dim must be multiples of 64");
  UTX_CHECK(cfg->mlp_ratio == 4, "utx_flux_create: mlp_ratio must be 4");
  utx_flux* h = new utx_flux();
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void utx_flux_destroy(utx_flux* h) {
  if (!h) return;
  drop_graphs(h);
  if (h->sp_flags_dev) cudaFree(h->sp_flags_dev);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  delete h;
}

int utx_flux_set_weights(utx_flux* h, const utx_flux_weights* w) {
  UTX_CHECK(h && w && w->double_blocks && w->single_blocks, "utx_flux_set_weights: null argument");
  h->w = *w;
  h->dbl.assign(w->double_blocks, w->double_blocks + h->cfg.num_layers);
  h->sgl.assign(w->single_blocks, w->single_blocks + h->cfg.num_single_layers);
  h->w.double_blocks = h->dbl.data();
  h->w.single_blocks = h->sgl.data();
  h->has_weights = true;
  drop_graphs(h);                 // captured launches hold the old weight pointers
  return 0;
}

size_t utx_flux_workspace_bytes(const utx_flux* h, int s_txt, int s_img) {
  if (!h) return 0;
  if ((s_txt + s_img) % h->sp_n != 0) return 0;
  return ws_layout(h->cfg, s_txt, s_img, h->sp_n).total;
}

int utx_flux_prepare(utx_flux* h, void* workspace, size_t workspace_bytes, const float* ids, const void* enc,
                     const float* pooled, int s_txt, int s_img, void* stream) {
  UTX_CHECK(h && h->has_weights, "utx_flux_prepare: weights not set");
  UTX_CHECK(workspace && ids && enc && pooled, "utx_flux_prepare: null argument");
  UTX_CHECK(s_txt >= 0 && s_img > 0, "utx_flux_prepare: bad sequence lengths");
  UTX_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "utx_flux_prepare: workspace must be 256B aligned");
  const int P = h->sp_n;
  UTX_CHECK((s_txt + s_img) % P == 0, "utx_flux_prepare: the sequence length must be a multiple of the sequence-parallel world size");
  const WsLayout L = ws_layout(h->cfg, s_txt, s_img, P);
  UTX_CHECK(workspace_bytes >= L.total, "utx_flux_prepare: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* b = static_cast<uint8_t*>(workspace);
  h->s_txt = s_txt; h->s_img = s_img;
  h->S_loc = (s_txt + s_img) / P;
  h->r0 = h->sp_rank * h->S_loc;
  h->st_loc = std::min(std::max(s_txt - h->r0, 0), h->S_loc);
  h->si_loc = h->S_loc - h->st_loc;
  h->img0 = std::max(h->r0 - s_txt, 0);
  h->x = reinterpret_cast<bf16*>(b + L.x); h->xn = reinterpret_cast<bf16*>(b + L.xn);
  h->qkv = reinterpret_cast<bf16*>(b + L.qkv); h->cat = reinterpret_cast<bf16*>(b + L.cat);
  h->ctx0 = reinterpret_cast<bf16*>(b + L.ctx0); h->v_tmp = reinterpret_cast<bf16*>(b + L.v_tmp);
  h->cos_t = reinterpret_cast<float*>(b + L.cos_t); h->sin_t = reinterpret_cast<float*>(b + L.sin_t);
  h->mod = reinterpret_cast<float*>(b + L.mod); h->temb = reinterpret_cast<float*>(b + L.temb);
  h->sincos = reinterpret_cast<float*>(b + L.sincos); h->hvec = reinterpret_cast<float*>(b + L.hvec);
  h->pooled = reinterpret_cast<float*>(b + L.pooled);
  h->step_params = reinterpret_cast<float*>(b + L.step_params);
  if (P > 1) {
    h->qkv_all = reinterpret_cast<bf16*>(b + L.qkv_all); h->attn_all = reinterpret_cast<bf16*>(b + L.attn_all);
    h->attn_recv = reinterpret_cast<bf16*>(b + L.attn_recv); h->v_loc = reinterpret_cast<bf16*>(b + L.v_loc);
    h->v_all = reinterpret_cast<bf16*>(b + L.v_all);
  }
  if (!h->sp_regions.empty()) {
    const RegionLayout R = region_layout(h->cfg, s_txt, s_img, P);
    UTX_CHECK(h->sp_region_bytes >= R.total, "utx_flux_prepare: the peer exchange region is too small for this sequence");
    h->sp_off_cat = R.cat; h->sp_off_qkv = R.qkv_all; h->sp_off_flags = R.flags;
    uint8_t* mine = h->sp_regions[h->sp_rank];
    h->cat = reinterpret_cast<bf16*>(mine + R.cat);
    h->qkv_all = reinterpret_cast<bf16*>(mine + R.qkv_all);
    std::vector<unsigned*> fl(P);
    for (int r = 0; r < P; ++r) fl[r] = reinterpret_cast<unsigned*>(h->sp_regions[r] + R.flags);
    if (!h->sp_flags_dev) UTX_CUDA(cudaMalloc(&h->sp_flags_dev, 32 * sizeof(unsigned*)));
    UTX_CUDA(cudaMemcpyAsync(h->sp_flags_dev, fl.data(), P * sizeof(unsigned*), cudaMemcpyHostToDevice, st));
    UTX_CUDA(cudaStreamSynchronize(st));          // fl is a stack object
  }
  drop_graphs(h);                 // buffers, sequence lengths or RoPE table may have changed
  const int D = D_of(h->cfg);
  UTX_TRY(rope_table(ids, s_txt + s_img, h->cos_t, h->sin_t, st));     // every rank keeps the whole table: rows are global tokens
  if (h->st_loc > 0)
    UTX_TRY(gemm1(static_cast<const bf16*>(enc) + static_cast<long>(h->r0) * h->cfg.joint_attention_dim, h->cfg.joint_attention_dim,
                  static_cast<const bf16*>(h->w.w_ctx_embed), h->cfg.joint_attention_dim,
                  static_cast<const bf16*>(h->w.b_ctx_embed), h->ctx0, D, h->st_loc, D, h->cfg.joint_attention_dim, EPI_BIAS,
                  nullptr, nullptr, 0, st));
  UTX_CUDA(cudaMemcpyAsync(h->pooled, pooled, sizeof(float) * h->cfg.pooled_projection_dim, cudaMemcpyDeviceToDevice,
                           st));
  h->prepared = true;
  return 0;
}

// timestep.to(bf16) * 1000 and guidance.to(bf16) * 1000, each product rounded to bf16 [ext transformer forward]
static inline float scaled_bf16(float v) { return bf16_round(bf16_round(v) * 1000.0f); }

int utx_flux_forward(utx_flux* h, const void* latents, float timestep, float guidance, void* v_out, void* stream) {
  UTX_CHECK(h && h->prepared, "utx_flux_forward: call utx_flux_prepare first");
  UTX_CHECK(latents && v_out, "utx_flux_forward: null argument");
  return forward_impl(h, latents, scaled_bf16(timestep), scaled_bf16(guidance), nullptr, v_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace {
// tg_dev != nullptr: the sinusoid reads (t_eff, g_eff) from device memory (graph capture / replay), else from the arguments
int forward_impl(utx_flux* h, const void* latents, float t_eff, float g_eff, const float* tg_dev, void* v_out, cudaStream_t st) {
  const utx_flux_config& c = h->cfg;
  const utx_flux_weights& w = h->w;
  const int D = D_of(c), M4 = 4 * D;
  // rows of THIS rank (everything when not sequence-parallel): st text rows, then si image rows = image tokens img0 ...
  const int st_n = h->st_loc, si_n = h->si_loc, S = h->S_loc;
  const bool sp = h->sp_n > 1;
  const int sc_hl = sp ? c.num_heads / h->sp_n : 0;
  const long ldc5 = 5L * D;

  if (tg_dev) CAT_ELEM(time_sinusoid_dev(tg_dev, h->sincos, st));
  else CAT_ELEM(time_sinusoid(t_eff, g_eff, h->sincos, st));
  auto B = [](const void* p) { return static_cast<const bf16*>(p); };
  // temb = MLP_t(sin) + MLP_g(sin) + MLP_p(pooled)
  CAT_ELEM(gemv_bf16(B(w.w_t1), B(w.b_t1), h->sincos, h->hvec, D, 256, 0, 0, st));
  CAT_ELEM(gemv_bf16(B(w.w_t2), B(w.b_t2), h->hvec, h->temb, D, D, 1, 0, st));
  if (c.guidance_embeds) {
    CAT_ELEM(gemv_bf16(B(w.w_g1), B(w.b_g1), h->sincos + 256, h->hvec, D, 256, 0, 0, st));
    CAT_ELEM(gemv_bf16(B(w.w_g2), B(w.b_g2), h->hvec, h->temb, D, D, 1, 1, st));
  }
  CAT_ELEM(gemv_bf16(B(w.w_p1), B(w.b_p1), h->pooled, h->hvec, D, c.pooled_projection_dim, 0, 0, st));
  CAT_ELEM(gemv_bf16(B(w.w_p2), B(w.b_p2), h->hvec, h->temb, D, D, 1, 1, st));
  // every adaLN modulation vector of the step in one HBM-bound pass (6.5 GB of weights at the real size).  Sequence-parallel:
  // the rows are independent, so each rank computes its 1/P of them and the vector (4 MB) is all-gathered in place
  const long n_mod = n_mod_rows(c);
  if (sp && n_mod % h->sp_n == 0) {
    const long per = n_mod / h->sp_n, r_off = per * h->sp_rank;
    CAT_ELEM(gemv_bf16(B(w.w_mod) + r_off * D, B(w.b_mod) + r_off, h->temb, h->mod + r_off, static_cast<int>(per), D, 1, 0, st));
    CAT_OTHER(comm_allgather(h->sp_comm, h->mod + r_off, h->mod, static_cast<size_t>(per) * 4, st));
  } else {
    CAT_ELEM(gemv_bf16(B(w.w_mod), B(w.b_mod), h->temb, h->mod, static_cast<int>(n_mod), D, 1, 0, st));
  }

  // embedders: x = [ctx0 | x_embedder(latents)] (this rank's rows of it)
  if (st_n > 0)
    UTX_CUDA(cudaMemcpyAsync(h->x, h->ctx0, static_cast<size_t>(st_n) * D * 2, cudaMemcpyDeviceToDevice, st));
  if (si_n > 0)
    CAT_GEMM(gemm1(B(latents) + static_cast<long>(h->img0) * c.in_channels, c.in_channels, B(w.w_x_embed), c.in_channels,
                  B(w.b_x_embed), h->x + static_cast<long>(st_n) * D, D, si_n, D, c.in_channels, EPI_BIAS, nullptr, nullptr, 0, st));

  const float* mod = h->mod;
  for (int i = 0; i < c.num_layers; ++i) {
    const utx_double_block& b = h->dbl[i];
    const float* mi = mod + static_cast<long>(i) * 12 * D;   // img: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const float* mt = mi + 6L * D;                           // txt: same order
    CAT_ELEM(ln_modulate(h->x, D, h->xn, D, S, D, st_n, mt, mt + D, mi, mi + D, st));
    // q|k|v projection of both streams; per-head RMSNorm + RoPE of q and k fused into the epilogue
    CAT_GEMM(gemm_streams(h, h->xn, D, b.w_qkv_txt, b.b_qkv_txt, b.w_qkv_img, b.b_qkv_img, D, h->qkv, 3L * D, 3 * D, D,
                         EPI_BIAS, nullptr, nullptr, nullptr, 0, st, 2 * D, b.rms_q_txt, b.rms_k_txt, b.rms_q_img,
                         b.rms_k_img, sp));
    UTX_TRY(attention_block(h, st));
    CAT_GEMM(gemm_streams(h, h->cat, ldc5, b.w_out_txt, b.b_out_txt, b.w_out_img, b.b_out_img, D, h->x, D, D, D,
                         EPI_GATE_RES, mt + 2L * D, mi + 2L * D, h->x, D, st));
    CAT_ELEM(ln_modulate(h->x, D, h->xn, D, S, D, st_n, mt + 3L * D, mt + 4L * D, mi + 3L * D, mi + 4L * D, st));
    CAT_GEMM(gemm_streams(h, h->xn, D, b.w_ff1_txt, b.b_ff1_txt, b.w_ff1_img, b.b_ff1_img, D, h->cat + D, ldc5, M4, D,
                         EPI_BIAS_GELU, nullptr, nullptr, nullptr, 0, st));
    CAT_GEMM(gemm_streams(h, h->cat + D, ldc5, b.w_ff2_txt, b.b_ff2_txt, b.w_ff2_img, b.b_ff2_img, M4, h->x, D, D, M4,
                         EPI_GATE_RES, mt + 5L * D, mi + 5L * D, h->x, D, st));
  }
  const float* mod_s = mod + static_cast<long>(c.num_layers) * 12 * D;
  for (int i = 0; i < c.num_single_layers; ++i) {
    const utx_single_block& b = h->sgl[i];
    const float* ms = mod_s + static_cast<long>(i) * 3 * D;   // shift, scale, gate
    CAT_ELEM(ln_modulate(h->x, D, h->xn, D, S, D, 0, ms, ms + D, ms, ms + D, st));
    // one GEMM for to_q|to_k|to_v|proj_mlp: q,k,v -> qkv buffer (or the all-to-all send layout), GELU(mlp) -> cat[:, D:]
    // (RMSNorm + RoPE of q and k fused into the same epilogue)
    CAT_GEMM(gemm1(h->xn, D, B(b.w_qkvmlp), D, B(b.b_qkvmlp), h->qkv, 3L * D, S, 7 * D, D, EPI_BIAS_GELU, nullptr,
                  nullptr, 0, st, 3 * D, 3 * D, h->cat + D, ldc5, 2 * D, B(b.rms_q), B(b.rms_k), h->cos_t, h->sin_t, h->r0,
                  sc_hl, S, D, sp ? h : nullptr));
    UTX_TRY(attention_block(h, st));
    CAT_GEMM(gemm1(h->cat, ldc5, B(b.w_out), ldc5, B(b.b_out), h->x, D, S, D, 5 * D, EPI_GATE_RES, ms + 2L * D, h->x, D,
                  st));
  }
  // AdaLayerNormContinuous (chunk order: scale, shift) + proj_out on the img rows
  const float* mf = mod_s + static_cast<long>(c.num_single_layers) * 3 * D;
  bf16* ximg = h->x + static_cast<long>(st_n) * D;
  bf16* xnimg = h->xn + static_cast<long>(st_n) * D;
  if (si_n > 0) CAT_ELEM(ln_modulate(ximg, D, xnimg, D, si_n, D, 0, mf + D, mf, mf + D, mf, st));
  if (!sp) {
    CAT_GEMM(gemm1(xnimg, D, B(w.w_proj_out), D, B(w.b_proj_out), static_cast<bf16*>(v_out), c.in_channels, si_n,
                  c.in_channels, D, EPI_BIAS, nullptr, nullptr, 0, st));
    return 0;
  }
  // sequence-parallel: v for this rank's image rows, all-gathered (rows in global token order) so that EVERY rank ends the
  // forward with the whole v and applies the same Euler update to its copy of the latents
  if (si_n > 0)
    CAT_GEMM(gemm1(xnimg, D, B(w.w_proj_out), D, B(w.b_proj_out), h->v_loc + static_cast<long>(st_n) * c.in_channels, c.in_channels,
                  si_n, c.in_channels, D, EPI_BIAS, nullptr, nullptr, 0, st));
  CAT_OTHER(comm_allgather(h->sp_comm, h->v_loc, h->v_all, static_cast<size_t>(S) * c.in_channels * 2, st));
  UTX_CUDA(cudaMemcpyAsync(v_out, h->v_all + static_cast<long>(h->s_txt) * c.in_channels,
                           static_cast<size_t>(h->s_img) * c.in_channels * 2, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// One denoise step through a captured graph.  Returns 0 and sets *done when the step ran as a graph launch; leaves *done false
// when the caller should run the step eagerly (graphs disabled, profiling on, first call with this key, capture failed).
int step_via_graph(utx_flux* h, void* latents, int s_noise, float t_eff, float g_eff, float dsigma, cudaStream_t st, bool* done) {
  *done = false;
  if (h->use_graph < 0) {
    const char* e = std::getenv("UTX_FLUX_GRAPH");
    h->use_graph = (e && e[0] == '0') ? 0 : 1;
  }
  if (!h->use_graph || h->profile || h->sp_n > 1) return 0;   // (collectives stay outside graphs)
  utx_flux::StepGraph* g = nullptr;
  for (auto& c : h->graphs)
    if (c.latents == latents && c.s_noise == s_noise) g = &c;
  if (!g) {
    if (h->graphs.size() >= 8) drop_graphs(h);
    h->graphs.emplace_back();
    g = &h->graphs.back();
    g->latents = latents;
    g->s_noise = s_noise;
  }
  if (!g->exec) {
    if (g->seen++ == 0) return 0;          // first step with this key runs eagerly (lazy kernel attributes, module load)
    if (g->seen > 2) return 0;             // a capture already failed for this key: stay eager
    if (!h->capture_stream) UTX_CUDA(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
    long before[UTX_PROF_NCAT];
    for (int i = 0; i < UTX_PROF_NCAT; ++i) before[i] = h->launches[i];
    cudaStream_t cs = h->capture_stream;
    UTX_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int r = forward_impl(h, latents, 0.f, 0.f, h->step_params, h->v_tmp, cs);
    if (r == 0) r = run_cat(h, UTX_PROF_ELEM, cs, [&] {
      return euler_update(static_cast<bf16*>(latents), h->v_tmp, s_noise, h->cfg.in_channels, 0.f, cs, h->step_params + 2);
    });
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    for (int i = 0; i < UTX_PROF_NCAT; ++i) {
      g->launches[i] = h->launches[i] - before[i];
      h->launches[i] = before[i];          // nothing ran yet
    }
    if (r != 0 || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return 0;                            // eager fallback; g->seen > 2 from now on
    }
    const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      g->exec = nullptr;
      cudaGetLastError();
      return 0;
    }
  }
  UTX_TRY(set_step_scalars(h->step_params, t_eff, g_eff, dsigma, st));
  UTX_CUDA(cudaGraphLaunch(g->exec, st));
  for (int i = 0; i < UTX_PROF_NCAT; ++i) h->launches[i] += g->launches[i];
  h->launches[UTX_PROF_OTHER] += 1;        // set_step_scalars
  h->graph_replays++;
  *done = true;
  return 0;
}
}  // namespace

extern "C" {

int utx_flux_denoise(utx_flux* h, void* latents, int s_noise, const float* sigmas, int n_steps, float guidance,
                     void* stream) {
  UTX_CHECK(h && h->prepared, "utx_flux_denoise: call utx_flux_prepare first");
  UTX_CHECK(latents && sigmas && n_steps >= 0, "utx_flux_denoise: bad argument");
  UTX_CHECK(s_noise > 0 && s_noise <= h->s_img, "utx_flux_denoise: s_noise out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n_steps; ++i) {
    // t = 1000 sigma (fp32) -> .to(bf16) (:643) -> / 1000 in bf16 (:648)
    const float t_in = bf16_round(bf16_round(sigmas[i] * 1000.0f) / 1000.0f);
    bool done = false;
    UTX_TRY(step_via_graph(h, latents, s_noise, scaled_bf16(t_in), scaled_bf16(guidance), sigmas[i + 1] - sigmas[i], st, &done));
    if (done) continue;
    UTX_TRY(utx_flux_forward(h, latents, t_in, guidance, h->v_tmp, stream));
    CAT_ELEM(euler_update(static_cast<bf16*>(latents), h->v_tmp, s_noise, h->cfg.in_channels, sigmas[i + 1] - sigmas[i],
                          st));
  }
  return 0;
}

int utx_flux_set_sequence_parallel(utx_flux* h, utx_comm* comm) {
  UTX_CHECK(h, "utx_flux_set_sequence_parallel: null handle");
  const int n = comm ? comm_nranks(comm) : 1;
  UTX_CHECK(n >= 1 && h->cfg.num_heads % n == 0, "utx_flux_set_sequence_parallel: the world size must divide the number of heads");
  h->sp_comm = n > 1 ? comm : nullptr;
  h->sp_n = n;
  h->sp_rank = n > 1 ? comm_rank(comm) : 0;
  h->sp_regions.clear();
  h->prepared = false;            // the workspace layout depends on the mode: prepare again
  drop_graphs(h);
  return 0;
}

size_t utx_flux_sp_region_bytes(const utx_flux* h, int s_txt, int s_img) {
  if (!h || h->sp_n <= 1 || (s_txt + s_img) % h->sp_n != 0) return 0;
  return region_layout(h->cfg, s_txt, s_img, h->sp_n).total;
}

int utx_flux_set_sp_peers(utx_flux* h, void* const* regions, size_t region_bytes) {
  UTX_CHECK(h, "utx_flux_set_sp_peers: null handle");
  h->sp_regions.clear();
  h->sp_region_bytes = 0;
  h->prepared = false;
  drop_graphs(h);
  if (!regions) return 0;
  UTX_CHECK(h->sp_n > 1 && h->sp_n <= 8, "utx_flux_set_sp_peers: needs sequence-parallel mode with 2..8 ranks");
  for (int r = 0; r < h->sp_n; ++r) {
    UTX_CHECK(regions[r] && (reinterpret_cast<uintptr_t>(regions[r]) & 255) == 0, "utx_flux_set_sp_peers: null / unaligned region");
    h->sp_regions.push_back(static_cast<uint8_t*>(regions[r]));
  }
  h->sp_region_bytes = region_bytes;
  return 0;
}

long utx_flux_graph_replays(const utx_flux* h) { return h ? h->graph_replays : 0; }

int utx_flux_profile(utx_flux* h, int enable) {
  UTX_CHECK(h, "utx_flux_profile: null handle");
  h->profile = enable != 0;
  return 0;
}

int utx_flux_profile_read(utx_flux* h, long* launches, float* ms, int reset) {
  UTX_CHECK(h && launches && ms, "utx_flux_profile_read: null argument");
  for (auto& u : h->ev_used) {
    float t = 0.f;
    UTX_CUDA(cudaEventSynchronize(h->ev_pool[u.second + 1]));
    UTX_CUDA(cudaEventElapsedTime(&t, h->ev_pool[u.second], h->ev_pool[u.second + 1]));
    h->prof_ms[u.first] += t;
  }
  h->ev_used.clear();
  h->ev_next = 0;
  for (int i = 0; i < UTX_PROF_NCAT; ++i) {
    launches[i] = h->launches[i];
    ms[i] = h->prof_ms[i];
    if (reset) {
      h->launches[i] = 0;
      h->prof_ms[i] = 0.f;
    }
  }
  return 0;
}

}  // extern "C"
