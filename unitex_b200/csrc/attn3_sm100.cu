// Joint attention v3: one 128-row query tile per CTA, S and P double-buffered in TMEM so that nothing aliases.
//   out = softmax(q k^T / sqrt(128)) v   (flux_piplines/texturing/attention_processor.py:89-91)
//
// Why (profiles/r01_summary.md, scripts/attn_trace.cu): in v2 P overwrites its own S tile, so S(j+1) cannot be issued
// before PV(j) has consumed P(j): every query tile runs the serial chain  softmax (~1900 clk) -> PV (512) -> QK^T (512) ->
// softmax ...  and two tiles ping-ponging only half hide it (period 3500 clk per 2 x 1024 clk of MMA work, tensor pipe 53 %).
// Here the 512 TMEM columns hold  S0 | S1 | P0 | P1 | O  (128 + 128 + 64 + 64 + 128):
//   warp 0 / 3    TMA producers: Q once and the K_j ring (3 deep) / the V_j ring (2 deep)
//   warp 1        MMA issuer (whole warp, one elected lane):  S(j+1) = Q K_{j+1}^T is issued BEFORE  O += P(j) V_j, i.e.
//                 while the softmax of tile j is still running; the softmax never waits for the tensor pipe in steady state
//   warp 2        TMEM allocator
//   warps 4-11    softmax, TWO threads per query row (warps 4-7 keys 0-63, warps 8-11 keys 64-127 of the tile;
//                 warp % 4 = TMEM lane quarter): two softmax warps per SM sub-partition hide each other's latencies.  The
//                 halves of a row exchange their maxima through shared memory behind a 256-thread named barrier; row sums
//                 are merged once, before the epilogue.
// P = exp2(S*scale - m) is rounded to bf16 (two per 32-bit TMEM column) and consumed from TMEM as the A operand of the PV MMA
// (TS mode).  O / l are rescaled lazily (running max grows by > 8), after waiting for the previous PV to retire.
#include <cstdlib>
#include <type_traits>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NSK = 3, NSV = 2;            // K / V ring depths (K_{j+1} is consumed ~1.5 tiles ahead of V_j)
constexpr int kThreads = 384;
constexpr int TILE_BYTES = 128 * HD * 2;
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int OFF_Q = 0;
constexpr int OFF_K = TILE_BYTES;
constexpr int OFF_V = OFF_K + NSK * TILE_BYTES;
constexpr int OFF_BAR = OFF_V + NSV * TILE_BYTES;
constexpr int OFF_XCH = OFF_BAR + 32 * 8;  // float [3][2 halves][128 rows]: row-max ping-pong + row sums
constexpr int SMEM_TOTAL = OFF_XCH + 3 * 2 * 128 * 4 + 1024;
constexpr float kRescaleThreshold = 8.0f;
constexpr int kPolyOf8 = 2;                // exponential pairs per 8 on the FMA pipe (polynomial) instead of MUFU
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384;   // TMEM column offsets

enum Bar { Q_FULL = 0, K_FULL = 1, K_EMPTY = 4, V_FULL = 7, V_EMPTY = 9, S_FULL = 11, S_EMPTY = 13, P_FULL = 15, P_EMPTY = 17,
           O_FULL = 19, NUM_BARS = 20 };

__global__ void __launch_bounds__(kThreads, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, long ld_out, int S, int H,
                  float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + NUM_BARS);

  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int D = H * HD;
  const int n_kv = (S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) prefetch_tmap(&tm_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_BARS; ++i) {
      const bool by_softmax = (i >= S_EMPTY && i < S_EMPTY + 2) || (i >= P_FULL && i < P_FULL + 2);
      mbar_init(&bar[i], by_softmax ? 8 : 1);   // one arrival per softmax warp / one TMA or tcgen05.commit arrival
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer for Q and K (V has its own thread so
      // that a full V ring never holds back the K tile the next QK^T needs)
      mbar_arrive_expect_tx(&bar[Q_FULL], TILE_BYTES);
      tma_load_2d(smem + OFF_Q, &tm_qkv, &bar[Q_FULL], head * HD, q0);
      tma_load_2d(smem + OFF_Q + HALF_BYTES, &tm_qkv, &bar[Q_FULL], head * HD + 64, q0);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&bar[K_EMPTY + s], ph ^ 1);
        mbar_arrive_expect_tx(&bar[K_FULL + s], TILE_BYTES);
        uint8_t* kd = smem + OFF_K + s * TILE_BYTES;
        tma_load_2d(kd, &tm_qkv, &bar[K_FULL + s], D + head * HD, j * BKV);
        tma_load_2d(kd + HALF_BYTES, &tm_qkv, &bar[K_FULL + s], D + head * HD + 64, j * BKV);
        if (++s == NSK) { s = 0; ph ^= 1; }
      }
    } else if (warp == 3 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer for V
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&bar[V_EMPTY + s], ph ^ 1);
        mbar_arrive_expect_tx(&bar[V_FULL + s], TILE_BYTES);
        uint8_t* vd = smem + OFF_V + s * TILE_BYTES;
        tma_load_2d(vd, &tm_qkv, &bar[V_FULL + s], 2 * D + head * HD, j * BKV);
        tma_load_2d(vd + HALF_BYTES, &tm_qkv, &bar[V_FULL + s], 2 * D + head * HD + 64, j * BKV);
        if (++s == NSV) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer: warp-uniform loop, one elected lane issues
      const bool leader = elect_one();
      const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);   // Q (smem, K-major) x K (smem, K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);    // P (TMEM)          x V (smem, MN-major)
      auto bar_a = [&](int b) { return sb + OFF_BAR + b * 8; };
      auto commit = [&](int b) {
        if (leader) umma_commit_a(bar_a(b));
        __syncwarp();
      };
      auto issue_S = [&](int j, int ks) {   // S[j & 1] = Q K_j^T, K_j in ring stage ks
        const uint32_t q_addr = sb + OFF_Q, k_addr = sb + OFF_K + ks * TILE_BYTES;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
            umma_ss(tmem_u + TM_S + (j & 1) * 128, make_sdesc(q_addr + off, 16, 1024), make_sdesc(k_addr + off, 16, 1024), idesc_s,
                    kk != 0);
          }
        }
        __syncwarp();
      };
      auto issue_PV = [&](int j, int vs) {  // O += P[j & 1] V_j, V_j in ring stage vs
        const uint32_t v_addr = sb + OFF_V + vs * TILE_BYTES;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)
            umma_ts(tmem_u + TM_O, tmem_u + TM_P + (j & 1) * 64 + kk * 8, make_sdesc(v_addr + kk * 16 * 128, HALF_BYTES, 1024), idesc_o,
                    (j | kk) != 0);
        }
        __syncwarp();
      };
      mbar_wait_a(bar_a(Q_FULL), 0);
      mbar_wait_a(bar_a(K_FULL), 0);
      tc_fence_after();
      issue_S(0, 0);
      commit(S_FULL);
      commit(K_EMPTY);
      int ks = 1, vs = 0;            // ring stage of K_{j+1} / V_j
      uint32_t kph = 0, vph = 0;     // their phase parities
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          const int b = (j + 1) & 1;
          mbar_wait_a(bar_a(K_FULL + ks), kph);
          if (j >= 1) mbar_wait_a(bar_a(S_EMPTY + b), (((j + 1) >> 1) - 1) & 1);   // softmax has read S(j-1) out of this buffer
          tc_fence_after();
          issue_S(j + 1, ks);
          commit(S_FULL + b);
          commit(K_EMPTY + ks);
          if (++ks == NSK) { ks = 0; kph ^= 1; }
        }
        mbar_wait_a(bar_a(V_FULL + vs), vph);
        mbar_wait_a(bar_a(P_FULL + (j & 1)), (j >> 1) & 1);
        tc_fence_after();
        issue_PV(j, vs);
        commit(P_EMPTY + (j & 1));     // also "PV(j) retired" for the softmax's O rescale
        commit(V_EMPTY + vs);
        if (++vs == NSV) { vs = 0; vph ^= 1; }
      }
      commit(O_FULL);
    }
  } else {
    // ---------------------------------------------------------------- softmax: two threads per query row
    const int ew = warp & 3;                       // TMEM lane quarter this warp may touch
    const int hlf = (warp - 4) >> 2;               // which 64 key columns of the tile
    const int r = ew * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + TM_S + hlf * 64;     // + (j & 1) * 128
    const uint32_t tP = tmem_base + lane_off + TM_P + hlf * 32;     // + (j & 1) * 64
    const uint32_t tO = tmem_base + lane_off + TM_O + hlf * 64;
    float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
    auto xslot = [&](int buf, int h) { return xch + (buf * 2 + h) * 128 + r; };
    auto pair_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };   // the 8 softmax warps
    float m_used = -INFINITY, l = 0.f;
    auto tile = [&](int j, auto ragged_c) {
      constexpr bool RAGGED = decltype(ragged_c)::value;
      const int b = j & 1;
      mbar_wait(&bar[S_FULL + b], (j >> 1) & 1);
      tc_fence_after();
      const int kv_valid = S - j * BKV - hlf * 64;          // valid keys among my 64 columns
      uint32_t sv[2][32];
      tmem_ld32(tS + b * 128, sv[0]);
      tmem_ld32(tS + b * 128 + 32, sv[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[S_EMPTY + b]);       // the MMA warp may overwrite this S buffer with S(j+2)
      float mx8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mx8[k] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!RAGGED || c * 32 + i < kv_valid) mx8[c * 4 + ((i >> 1) & 3)] = fmaxf(mx8[c * 4 + ((i >> 1) & 3)], __uint_as_float(sv[c][i]));
      float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      *xslot(b, hlf) = mx;                                   // merge with the other half of the row
      pair_sync();
      mx = fmaxf(mx, *xslot(b, hlf ^ 1));
      const float m_new = mx * scale_log2;
      const bool upd = m_new > m_used + kRescaleThreshold;
      const float m_next = upd ? m_new : m_used;
      const float alpha = upd ? ex2_approx(m_used - m_next) : 1.0f;
      m_used = m_next;
      if (j > 0 && __any_sync(0xffffffffu, upd)) {
        mbar_wait(&bar[P_EMPTY + (b ^ 1)], ((j - 1) >> 1) & 1);   // PV(j-1) has retired: O is consistent and idle until P(j) exists
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t ov[16];
          tmem_ld16(tO + c * 16, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tO + c * 16, ov);
        }
      }
      uint64_t lsum2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
      const uint64_t sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_next, -m_next);
      if (j >= 2) mbar_wait(&bar[P_EMPTY + b], ((j >> 1) - 1) & 1);   // PV(j-2) has consumed this P buffer
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t x2 = fma2(pack2(__uint_as_float(sv[c][2 * i]), __uint_as_float(sv[c][2 * i + 1])), sc2, nm2);
          uint64_t p2;
          if ((i & 7) < kPolyOf8) {
            p2 = exp2_poly2(x2);
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            p2 = pack2(ex2_approx(x0), ex2_approx(x1));
          }
          if (RAGGED) {
            float p0, p1;
            unpack2(p2, p0, p1);
            if (c * 32 + 2 * i >= kv_valid) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= kv_valid) p1 = 0.f;
            p2 = pack2(p0, p1);
          }
          lsum2[i & 1] = add2(lsum2[i & 1], p2);
          float p0, p1;
          unpack2(p2, p0, p1);
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tP + b * 64 + c * 16, pk);
      }
      float ls0, ls1;
      unpack2(add2(lsum2[0], lsum2[1]), ls0, ls1);
      l = l * alpha + (ls0 + ls1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[P_FULL + b]);
    };
    const int n_full = S / BKV;
    for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
    if (n_full < n_kv) tile(n_full, std::true_type{});
    // ---------------------------------------------------------------- epilogue: merge the halves' row sums, write my 64 columns
    *xslot(2, hlf) = l;
    pair_sync();
    l += *xslot(2, hlf ^ 1);
    mbar_wait(&bar[O_FULL], 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q0 + r;
    bf16* orow = out + static_cast<long>(row) * ld_out + head * HD + hlf * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld32(tO + c * 32, ov);
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

int attention3_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream) {
  UTX_CHECK(S > 0 && H > 0, "attention: empty problem");
  UTX_CHECK(ld_qkv >= 3L * H * HD && ld_out % 8 == 0, "attention: bad leading dimensions");
  CUtensorMap tm;
  UTX_TRY(make_tmap_2d_bf16(&tm, qkv, S, 3L * H * HD, ld_qkv, 128, 64));
  static bool attr_set = false;
  if (!attr_set) {
    UTX_CUDA(cudaFuncSetAttribute(attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((S + BQ - 1) / BQ, H);
  attention3_kernel<<<grid, kThreads, SMEM_TOTAL, stream>>>(tm, out, ld_out, S, H, scale_log2);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace utx
