// Joint attention v2: two 128-row query tiles per CTA ping-ponging on the tensor pipe, P kept in TMEM.
//   out = softmax(q k^T / sqrt(128)) v   (flux_piplines/texturing/attention_processor.py:89-91)
//
// Why (profiles/r01_summary.md): v1 runs one softmax warp per SM sub-partition, so MUFU sits at 33 % and the tensor pipe
// at 31 %.  Here a CTA owns 256 query rows of one head:
//   warp 0        TMA producer: Q_A, Q_B once; K_j / V_j tiles through two 2-deep mbarrier rings
//   warp 1        MMA issuer (one thread).  Per kv tile j:  O_A += P_A V_j ; S_A = Q_A K_{j+1}^T ; O_B += P_B V_j ;
//                 S_B = Q_B K_{j+1}^T  -- so while warpgroup A does softmax(S_A) the pipe runs tile B's MMAs and vice versa
//   warp 2        TMEM allocator (all 512 columns: S_A | S_B | O_A | O_B, 128 fp32 columns each)
//   warps 4-7     softmax warpgroup A (thread == query row), warps 8-11 warpgroup B
// P = exp2(S*scale - m) is rounded to bf16 and written back over the first 64 columns of its own S tile
// (tcgen05.st, two bf16 per 32-bit column), and the PV product reads it from there as the A operand (TS-mode MMA): no smem
// round trip for P, half the smem operand traffic of the PV MMAs.  S is read from TMEM twice (row max with four loads in
// flight, then exponentials with the next 32-column chunk prefetched) so the softmax threads stay under 168 registers.
// O / l are rescaled lazily (running max grows by > 8).
// `tcgen05.commit` of S_X(j+1) retires every earlier MMA of the issuing thread, so "S_X(j+1) ready" also means
// "PV_X(j) done": no separate barrier guards the O rescale or the P overwrite.
#include <cstdlib>
#include <type_traits>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;       // rows per softmax warpgroup
constexpr int BKV = 128;
constexpr int kThreads = 384;
constexpr int TILE_BYTES = 128 * HD * 2;
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int OFF_Q = 0;                   // 2 tiles
constexpr int OFF_K = 2 * TILE_BYTES;      // 2 stages
constexpr int OFF_V = 4 * TILE_BYTES;      // 2 stages
constexpr int OFF_BAR = 6 * TILE_BYTES;
constexpr int SMEM_TOTAL = OFF_BAR + 32 * 8 + 1024;
constexpr float kRescaleThreshold = 8.0f;
constexpr int kRegsSoftmax = 208, kRegsOther = 88;   // 256 x 208 + 128 x 88 = 64512 = 384 x 168

// scripts/attn_trace.cu compiles this file with -DUTX_ATTN_TRACE to record SM-clock stamps of one CTA's roles
#ifdef UTX_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;   // [3 roles][n_kv][8 points]
#define UTX_TR(role, j, pt)                                                                                   \
  do {                                                                                                        \
    if (blockIdx.x == 1 && blockIdx.y == 0 && g_attn_trace) g_attn_trace[((role) * 128 + (j)) * 8 + (pt)] = clock64(); \
  } while (0)
#else
#define UTX_TR(role, j, pt) do { } while (0)
#endif

enum Bar { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, P_FULL = 11, O_FULL = 13, PH_FULL = 14, NUM_BARS = 16 };

// kPolyOf8: exponential pairs per 8 evaluated by the FMA-pipe polynomial instead of MUFU
template <int kPolyOf8>
__global__ void __launch_bounds__(kThreads, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, long ld_out, int S, int H,
                  float scale_log2, const AttnScatter sc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + NUM_BARS);

  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int head = blockIdx.y;
  const int D = H * HD;
  const int n_kv = (S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) prefetch_tmap(&tm_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NUM_BARS; ++i)
      mbar_init(&bar[i], (i == P_FULL || i == P_FULL + 1 || i == PH_FULL || i == PH_FULL + 1) ? 4 : 1);   // 4 softmax warps each
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // register re-balancing: the softmax warpgroups keep a whole 128-column S row per thread
  // (ptxas only honours the new budget for code nested under the branch that executes the setmaxnreg)
  if (warp < 4) {
  setmaxnreg_dec<kRegsOther>();
  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    mbar_arrive_expect_tx(&bar[Q_FULL], 2 * TILE_BYTES);
    for (int x = 0; x < 2; ++x) {
      tma_load_2d(smem + OFF_Q + x * TILE_BYTES, &tm_qkv, &bar[Q_FULL], head * HD, q0 + x * BQ);
      tma_load_2d(smem + OFF_Q + x * TILE_BYTES + HALF_BYTES, &tm_qkv, &bar[Q_FULL], head * HD + 64, q0 + x * BQ);
    }
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
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: the whole warp runs the loop (waits and
    // descriptor arithmetic stay warp-uniform), one elected lane issues.  With a single divergent thread owning the loop
    // every UTCHMMA cost ~19 SASS instructions (R2UR + elect loop) = ~91 clk against the 64 clk the 128x128x16 MMA takes:
    // the tensor pipe was issue-starved (scripts/attn_trace.cu timeline, profiles/r01_summary.md).
    const bool leader = elect_one();
    const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);   // warp-uniform shared-window base
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform copy: TMEM operands stay in uniform registers
    constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);   // Q (smem, K-major) x K (smem, K-major)
    constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0, 1);    // P (TMEM)          x V (smem, MN-major)
    auto issue_S = [&](int x, int j) {   // S_x = Q_x K_j^T
      const uint32_t q_addr = sb + OFF_Q + x * TILE_BYTES;
      const uint32_t k_addr = sb + OFF_K + (j & 1) * TILE_BYTES;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_ss(tmem_u + x * 128, make_sdesc(q_addr + off, 16, 1024), make_sdesc(k_addr + off, 16, 1024), idesc_s, kk != 0);
        }
        umma_commit_a(sb + OFF_BAR + (S_FULL + x) * 8);
      }
      __syncwarp();
    };
    auto issue_PV = [&](int x, int j, int half) {  // O_x += P_x[:, 64*half : 64*half+64] V_j[64*half : ...], P_x bf16 over S_x
      const uint32_t v_addr = sb + OFF_V + (j & 1) * TILE_BYTES;
      if (leader) {
#pragma unroll
        for (int kk = 4 * half; kk < 4 * half + 4; ++kk)
          umma_ts(tmem_u + 256 + x * 128, tmem_u + x * 128 + kk * 8, make_sdesc(v_addr + kk * 16 * 128, HALF_BYTES, 1024),
                  idesc_o, (j | kk) != 0);
      }
      __syncwarp();
    };
    auto commit = [&](int b) {
      if (leader) umma_commit_a(sb + OFF_BAR + (b) * 8);
      __syncwarp();
    };
    mbar_wait_a(sb + OFF_BAR + (Q_FULL) * 8, 0);
    mbar_wait_a(sb + OFF_BAR + (K_FULL) * 8, 0);
    tc_fence_after();
    issue_S(0, 0);
    issue_S(1, 0);
    commit(K_EMPTY);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1, sn = (j + 1) & 1;
      const bool more = j + 1 < n_kv;
      if (leader) UTX_TR(2, j, 0);
      mbar_wait_a(sb + OFF_BAR + (V_FULL + s) * 8, (j >> 1) & 1);
      if (leader) UTX_TR(2, j, 1);
      mbar_wait_a(sb + OFF_BAR + (PH_FULL) * 8, j & 1);
      if (leader) UTX_TR(2, j, 2);
      tc_fence_after();
      issue_PV(0, j, 0);
      mbar_wait_a(sb + OFF_BAR + (P_FULL) * 8, j & 1);
      tc_fence_after();
      issue_PV(0, j, 1);
      if (more) {
        mbar_wait_a(sb + OFF_BAR + (K_FULL + sn) * 8, ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_S(0, j + 1);
      }
      if (leader) UTX_TR(2, j, 3);
      mbar_wait_a(sb + OFF_BAR + (PH_FULL + 1) * 8, j & 1);
      if (leader) UTX_TR(2, j, 4);
      tc_fence_after();
      issue_PV(1, j, 0);
      mbar_wait_a(sb + OFF_BAR + (P_FULL + 1) * 8, j & 1);
      tc_fence_after();
      issue_PV(1, j, 1);
      commit(V_EMPTY + s);
      if (more) {
        issue_S(1, j + 1);
        commit(K_EMPTY + sn);
      }
      if (leader) UTX_TR(2, j, 5);
    }
    commit(O_FULL);
  }
  } else {
    setmaxnreg_inc<kRegsSoftmax>();
    // ---------------------------------------------------------------- softmax warpgroups: thread == query row
    const int x = (warp - 4) >> 2;                 // 0 = tile A, 1 = tile B
    const int ew = (warp - 4) & 3;                 // == warp % 4: the TMEM lane quarter this warp may touch
    const int r = ew * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + x * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + x * 128;
    float m_used = -INFINITY, l = 0.f;
    // one KV tile of the online softmax; RAGGED only for a last tile with fewer than 128 keys
    auto tile = [&](int j, auto ragged_c) {
      constexpr bool RAGGED = decltype(ragged_c)::value;
      if (lane == 0 && ew == 0) UTX_TR(x, j, 0);
      mbar_wait(&bar[S_FULL + x], j & 1);
      if (lane == 0 && ew == 0) UTX_TR(x, j, 1);
      tc_fence_after();
      const int kv_valid = S - j * BKV;
      // the whole 128-column S row lives in registers (setmaxnreg gives the softmax warpgroups 232): one TMEM read per tile
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, sv[c]);
      tmem_ld_wait();
      if (lane == 0 && ew == 0) UTX_TR(x, j, 2);
      float mx8[8];      // eight independent chains: four left every FMNMX3 waiting on its predecessor
#pragma unroll
      for (int k = 0; k < 8; ++k) mx8[k] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!RAGGED || c * 32 + i < kv_valid) mx8[c * 2 + ((i >> 1) & 1)] = fmaxf(mx8[c * 2 + ((i >> 1) & 1)], __uint_as_float(sv[c][i]));
      const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      const float m_new = mx * scale_log2;
      const bool upd = m_new > m_used + kRescaleThreshold;
      const float m_next = upd ? m_new : m_used;
      const float alpha = upd ? ex2_approx(m_used - m_next) : 1.0f;
      m_used = m_next;
      if (j > 0 && __any_sync(0xffffffffu, upd)) {     // PV_x(j-1) has retired (see header): O_x is consistent
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t ov[16];
          tmem_ld16(tO + c * 16, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tO + c * 16, ov);
        }
      }
      if (lane == 0 && ew == 0) UTX_TR(x, j, 3);
      // exponentials in place: packed fp32x2 scale/sum, kPolyOf8 of every 8 pairs on the FMA pipe instead of MUFU;
      // P (bf16) is written back over S in four 16-column stores; the first half is announced early so the PV MMAs
      // of keys 0-63 start while keys 64-127 are still being exponentiated
      uint64_t lsum2 = pack2(0.f, 0.f);
      const uint64_t sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_next, -m_next);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
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
          lsum2 = add2(lsum2, p2);
          float p0, p1;
          unpack2(p2, p0, p1);
          pk[i] = pack_bf16x2(p0, p1);
        }
        if (c == 2) {   // stores of chunks 0-1 (keys 0-63) have had this chunk's arithmetic to land
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[PH_FULL + x]);
        }
        tmem_st16(tS + c * 16, pk);
      }
      float ls0, ls1;
      unpack2(lsum2, ls0, ls1);
      l = l * alpha + (ls0 + ls1);
      if (lane == 0 && ew == 0) UTX_TR(x, j, 4);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[P_FULL + x]);
      if (lane == 0 && ew == 0) UTX_TR(x, j, 5);
    };
    const int n_full = S / BKV;
    for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
    if (n_full < n_kv) tile(n_full, std::true_type{});
    // ---------------------------------------------------------------- epilogue
    mbar_wait(&bar[O_FULL], 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q0 + x * BQ + r;
    bf16* orow = out + static_cast<long>(row) * ld_out + head * HD;
    if (sc.rows_per_rank > 0) {          // sequence-parallel direct mode: the row's owner receives it over NVLink
      const int owner = row / sc.rows_per_rank;
      bf16* base = sc.base[0];
#pragma unroll
      for (int q = 1; q < 8; ++q) base = owner == q ? sc.base[q] : base;
      orow = base + static_cast<long>(row - owner * sc.rows_per_rank) * sc.ld + sc.col0 + head * HD;
    }
    // direct (peer) mode: a row's 256 bytes go out as ONE contiguous burst per 16 lanes instead of 16-byte pieces of 32 rows per
    // store instruction -- staged through this tile's Q buffer, which no MMA reads any more (O_FULL covers every earlier MMA)
    const bool staged = sc.rows_per_rank > 0;
    uint8_t* stage = smem + OFF_Q + x * TILE_BYTES + ew * 8192;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t ov[32];
      tmem_ld32(tO + c * 32, ov);
      tmem_ld_wait();
      if (row < S || staged) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
          if (staged) *reinterpret_cast<uint4*>(stage + lane * 256 + (((c * 4 + g) ^ (lane & 15)) << 4)) = o;
          else *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o;
        }
      }
    }
    if (staged) {
      __syncwarp();
      const int row0 = q0 + x * BQ + ew * 32;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int idx = i * 32 + lane, rr = idx >> 4, kc = idx & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(stage + rr * 256 + ((kc ^ (rr & 15)) << 4));
        const int grow = row0 + rr;
        if (grow < S) {
          const int owner = grow / sc.rows_per_rank;
          bf16* base = sc.base[0];
#pragma unroll
          for (int q = 1; q < 8; ++q) base = owner == q ? sc.base[q] : base;
          *reinterpret_cast<uint4*>(base + static_cast<long>(grow - owner * sc.rows_per_rank) * sc.ld + sc.col0 + head * HD + kc * 8) = v;
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

int attention2_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream,
                    const AttnScatter* scatter) {
  UTX_CHECK(S > 0 && H > 0, "attention: empty problem");
  UTX_CHECK(ld_qkv >= 3L * H * HD && ld_out % 8 == 0, "attention: bad leading dimensions");
  CUtensorMap tm;
  UTX_TRY(make_tmap_2d_bf16(&tm, qkv, S, 3L * H * HD, ld_qkv, 128, 64));
  static int poly = -1;
  if (poly < 0) {
    const char* e = std::getenv("UTX_ATTN_POLY");   // tuning knob; default measured best at S = 9728
    poly = e ? std::atoi(e) : 2;
    if (poly < 0 || poly > 3) poly = 2;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    UTX_CUDA(cudaFuncSetAttribute(attention2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    UTX_CUDA(cudaFuncSetAttribute(attention2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    UTX_CUDA(cudaFuncSetAttribute(attention2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    UTX_CUDA(cudaFuncSetAttribute(attention2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  AttnScatter sc{};
  if (scatter) sc = *scatter;
  dim3 grid((S + 2 * BQ - 1) / (2 * BQ), H);
  switch (poly) {
    case 0: attention2_kernel<0><<<grid, kThreads, SMEM_TOTAL, stream>>>(tm, out, ld_out, S, H, scale_log2, sc); break;
    case 1: attention2_kernel<1><<<grid, kThreads, SMEM_TOTAL, stream>>>(tm, out, ld_out, S, H, scale_log2, sc); break;
    case 3: attention2_kernel<3><<<grid, kThreads, SMEM_TOTAL, stream>>>(tm, out, ld_out, S, H, scale_log2, sc); break;
    default: attention2_kernel<2><<<grid, kThreads, SMEM_TOTAL, stream>>>(tm, out, ld_out, S, H, scale_log2, sc); break;
  }
  UTX_CUDA(cudaGetLastError());
  return 0;
}

// The engine's attention entry point (kernels.h).  Round 1 kept two more kernels behind UTX_ATTN_IMPL (v1: one query tile per
// CTA with P through shared memory; v3: S and P double-buffered in TMEM); both lost to this one in situ and were removed
// from the product library (git history: attn_sm100.cu, attn3_sm100.cu; measurements in profiles/r01_summary.md).
int attention_bf16(const bf16* qkv, long ld_qkv, bf16* out, long ld_out, int S, int H, cudaStream_t stream,
                   const AttnScatter* scatter) {
  return attention2_bf16(qkv, ld_qkv, out, ld_out, S, H, stream, scatter);
}

}  // namespace utx
