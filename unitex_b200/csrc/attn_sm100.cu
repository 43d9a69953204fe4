// Joint (txt+img) bidirectional attention for the MM-DiT blocks, flash-style, on tcgen05 + TMEM + TMA.
//   out = softmax(q k^T / sqrt(128)) v      per head, no mask         (reference math:
//   flux_piplines/texturing/attention_processor.py:89-91; q/k arrive RMS-normed and RoPE'd, :61-87)
//
// One CTA = 128 query rows of one head.  Warp roles:
//   warp 0      TMA producer: Q once, then K_j / V_j 128-row tiles through two 2-deep mbarrier rings
//   warp 1      MMA issuer (one thread): S_j = Q K_j^T (SS, both K-major) into a double-buffered TMEM tile,
//               then O += P_j V_j (A = P from smem K-major, B = V MN-major straight from the [kv, d] TMA tile)
//   warp 2      TMEM allocator
//   warps 4..7  softmax: one thread per query row (tcgen05.ld 32x32b: lane == row), online softmax in the log2
//               domain with lazy rescaling (O/l are only rescaled when the running max grows by > 8, so the
//               TMEM round trip of O is rare), P rounded to bf16 into 128B-swizzled smem for the PV MMA.
// S_{j+1} is issued before softmax_j finishes, so QK^T of the next tile and PV of the previous one overlap the
// exponentials of the current one.
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {
namespace {

constexpr int HD = 128;    // head dim (FLUX attention_head_dim)
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int kThreads = 256;
constexpr int TILE_BYTES = 128 * HD * 2;   // 32 KB: [2 d-halves][128 rows][64] bf16, 128B swizzle
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int OFF_Q = 0;
constexpr int OFF_P = TILE_BYTES;
constexpr int OFF_K = 2 * TILE_BYTES;
constexpr int OFF_V = 4 * TILE_BYTES;
constexpr int OFF_BAR = 6 * TILE_BYTES;
constexpr int SMEM_TOTAL = OFF_BAR + 32 * 8 + 1024;
constexpr float kRescaleThreshold = 8.0f;

enum Bar { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, P_FULL = 13,
           PV_DONE = 14, O_FULL = 15, NUM_BARS = 16 };

__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, long ld_out, int S, int H,
                 float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int D = H * HD;
  const int n_kv = (S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) prefetch_tmap(&tm_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool by_softmax_warps = (i == S_EMPTY || i == S_EMPTY + 1 || i == P_FULL);
      mbar_init(&bar[i], by_softmax_warps ? 4 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // 2 x 128 fp32 columns
  const uint32_t tmem_O = tmem_base + 256;    // 128 fp32 columns

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    mbar_arrive_expect_tx(&bar[Q_FULL], TILE_BYTES);
    tma_load_2d(smem + OFF_Q, &tm_qkv, &bar[Q_FULL], head * HD, q0);
    tma_load_2d(smem + OFF_Q + HALF_BYTES, &tm_qkv, &bar[Q_FULL], head * HD + 64, q0);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&bar[K_EMPTY + s], ph ^ 1);
      mbar_arrive_expect_tx(&bar[K_FULL + s], TILE_BYTES);
      uint8_t* kd = smem + OFF_K + s * TILE_BYTES;
      tma_load_2d(kd, &tm_qkv, &bar[K_FULL + s], D + head * HD, j * BKV);
      tma_load_2d(kd + HALF_BYTES, &tm_qkv, &bar[K_FULL + s], D + head * HD + 64, j * BKV);
      mbar_wait(&bar[V_EMPTY + s], ph ^ 1);
      mbar_arrive_expect_tx(&bar[V_FULL + s], TILE_BYTES);
      uint8_t* vd = smem + OFF_V + s * TILE_BYTES;
      tma_load_2d(vd, &tm_qkv, &bar[V_FULL + s], 2 * D + head * HD, j * BKV);
      tma_load_2d(vd + HALF_BYTES, &tm_qkv, &bar[V_FULL + s], 2 * D + head * HD + 64, j * BKV);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);    // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);     // P (K-major) x V (MN-major)
    const uint32_t q_addr = smem_u32(smem + OFF_Q);
    const uint32_t p_addr = smem_u32(smem + OFF_P);
    auto issue_S = [&](int j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&bar[K_FULL + s], ph);
      mbar_wait(&bar[S_EMPTY + s], ph ^ 1);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(smem + OFF_K + s * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
        umma_ss(tmem_S + s * BKV, make_sdesc(q_addr + off, 16, 1024), make_sdesc(k_addr + off, 16, 1024), idesc_s,
                kk != 0);
      }
      umma_commit(&bar[K_EMPTY + s]);
      umma_commit(&bar[S_FULL + s]);
    };
    mbar_wait(&bar[Q_FULL], 0);
    issue_S(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_S(j + 1);
      const int s = j & 1;
      mbar_wait(&bar[V_FULL + s], (j >> 1) & 1);
      mbar_wait(&bar[P_FULL], j & 1);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(smem + OFF_V + s * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk) {
        // A: P[128 q, 16 kv] K-major.  B: V[16 kv, 128 d] MN-major: 64-wide d groups LBO apart, 8-row kv groups SBO apart
        const uint64_t da = make_sdesc(p_addr + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
        const uint64_t db = make_sdesc(v_addr + kk * 16 * 128, HALF_BYTES, 1024);
        umma_ss(tmem_O, da, db, idesc_o, (j | kk) != 0);
      }
      umma_commit(&bar[V_EMPTY + s]);
      umma_commit(&bar[PV_DONE]);
      if (j == n_kv - 1) umma_commit(&bar[O_FULL]);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax: thread == query row
    const int ew = warp - 4;
    const int r = ew * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(ew * 32) << 16;
    uint8_t* prow = smem + OFF_P + r * 128;
    float m_used = -INFINITY;   // running max (log2 domain) currently folded into O and l
    float l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      mbar_wait(&bar[S_FULL + s], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_S + lane_off + s * BKV + c * 32, sv[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[S_EMPTY + s]);

      const int kv_valid = S - j * BKV;   // >= 128 except on a ragged last tile
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(sv[c][i]);
          if (kv_valid < BKV && c * 32 + i >= kv_valid) {
            x = -INFINITY;
            sv[c][i] = __float_as_uint(x);
          }
          mx = fmaxf(mx, x);
        }
      const float m_new = mx * scale_log2;
      const bool upd = m_new > m_used + kRescaleThreshold;
      const float m_next = upd ? m_new : m_used;
      const float alpha = upd ? ex2_approx(m_used - m_next) : 1.0f;
      float lsum = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c][i]), scale_log2, -m_next));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c][i + 1]), scale_log2, -m_next));
          lsum += p0 + p1;
          pk[c * 16 + i / 2] = pack_bf16x2(p0, p1);
        }
      l = l * alpha + lsum;
      m_used = m_next;

      if (j > 0) {
        mbar_wait(&bar[PV_DONE], (j - 1) & 1);   // PV_{j-1} retired: P buffer free, O consistent
        tc_fence_after();
        if (__any_sync(0xffffffffu, upd)) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t ov[32];
            tmem_ld32(tmem_O + lane_off + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(tmem_O + lane_off + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      // P row -> smem, K-major 128B-swizzled: 16B chunk ch of row r lands at chunk (ch ^ (r & 7))
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        const int half = ch >> 3, cc = ch & 7;
        uint4 val = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(prow + half * HALF_BYTES + ((cc ^ (r & 7)) << 4)) = val;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[P_FULL]);
    }
    // ---------------------------------------------------------------- epilogue: O / l -> bf16 -> global
    mbar_wait(&bar[O_FULL], 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q0 + r;
    bf16* orow = out + static_cast<long>(row) * ld_out + head * HD;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t ov[32];
      tmem_ld32(tmem_O + lane_off + c * 32, ov);
      tmem_ld_wait();
      if (row < S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int attention2_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream);
int attention3_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream);

int attention_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream) {
  // UTX_ATTN_IMPL=1: one query tile per CTA, P through smem (this file).  =2 (default): two-tile ping-pong, P in TMEM
  // (attn2_sm100.cu).  =3: one query tile, S and P double-buffered in TMEM (attn3_sm100.cu).  All stay built so the
  // parity tests can run each.
  const char* impl = getenv("UTX_ATTN_IMPL");
  if (impl != nullptr && impl[0] == '3') return attention3_bf16(qkv, ld_qkv, out, ld_out, S, H, stream);
  if (impl == nullptr || impl[0] != '1') return attention2_bf16(qkv, ld_qkv, out, ld_out, S, H, stream);
  UTX_CHECK(S > 0 && H > 0, "attention: empty problem");
  UTX_CHECK(ld_qkv >= 3L * H * HD && ld_out % 8 == 0, "attention: bad leading dimensions");
  CUtensorMap tm;
  UTX_TRY(make_tmap_2d_bf16(&tm, qkv, S, 3L * H * HD, ld_qkv, 128, 64));
  static bool attr_set = false;
  if (!attr_set) {
    UTX_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((S + BQ - 1) / BQ, H);
  attention_kernel<<<grid, kThreads, SMEM_TOTAL, stream>>>(tm, out, ld_out, S, H, scale_log2);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace utx
