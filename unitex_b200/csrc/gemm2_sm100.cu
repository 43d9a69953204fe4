// 2-CTA (cta_group::2) variant of the bf16 GEMM:  C = epilogue(A[M,K] @ W[N,K]^T + bias), 256 x 256 tiles per CTA PAIR.
//
// Why (profiles/r01_summary.md): the 1-CTA kernel streams 48 KB of operands per 512 MMA-cycles into smem AND reads them
// back out for the MMA (96 + 96 B/clk against a 128 B/clk shared-memory port), which pins the tensor pipe at ~70 % active.
// A pair of CTAs on one TPC shares the weight tile: each CTA stages its 128 A rows and HALF of the 256 W rows (32 KB per
// stage), the leader issues tcgen05.mma.cta_group::2 (M = 256 across the two SMs) and each SM reads 64 B/clk.
//   every CTA : warp 0 TMA producer (its A rows, its half of W; complete_tx on the LEADER's full barrier),
//               warp 2 TMEM allocator (cta_group::2), warps 4-7 epilogue for its own 128 accumulator rows
//   leader    : warp 1 MMA issuer; its commits are multicast to both CTAs' empty / tmem-full barriers;
//               both CTAs' epilogue warps arrive remotely on the leader's tmem-empty barrier.
// Epilogues, grouping of two problems and the band-swizzled tile walk are those of gemm_sm100.cu.
#include <cstdlib>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {
namespace {

constexpr int BM = 128;        // rows per CTA (256 per pair)
constexpr int BK = 64;
constexpr int GROUP_M_DEFAULT = 8;   // bands of 8 pair-row-blocks (2048 rows); UTX_GEMM_GROUP_M overrides (tuning knob)
constexpr int kThreads = 256;
constexpr int A_BYTES = BM * BK * 2;
// BN = tile columns (each CTA stages BN/2 rows of W): 256 for every DiT Linear; 128 for the Cout = 128 convolutions of the VAE's
// 1024^2 stage, where the 1-CTA 128 x 128 tile needs 128 B/clk of shared-memory fill per MMA clock (its whole port) and the
// pair tile 96 B/clk
template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 256 ? 6 : 8;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_STAGE_OFF = BAR_OFF + 256;            // 4 epilogue warps x 8 KB: staging of scattered stores (gemm_epilogue.cuh)
  static constexpr int SMEM_TOTAL = EPI_STAGE_OFF + 4 * 8192 + 1024;
  static_assert((2 * STAGES + 4) * 8 + 16 <= 256, "barrier block");
  static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
};

struct DevParams {
  int K, tiles_n, nprob, total_tiles, group_m;
  int conv_cblk, conv_w, conv_hw;   // implicit 3x3 convolution (see gemm_sm100.cu): channel blocks per tap (0 = plain GEMM), width, pixels per image
  EpiParams e;
  EpiProblem prob[2];
};

struct TileCoord { int pi, m_blk, n_blk; };

__device__ __forceinline__ TileCoord decode_tile(const DevParams& p, int t) {
  TileCoord tc;
  tc.pi = 0;
  const int t0 = p.prob[0].tiles_m * p.tiles_n;
  if (p.nprob > 1 && t >= t0) { tc.pi = 1; t -= t0; }
  const int tiles_m = p.prob[tc.pi].tiles_m;
  const int GROUP_M = p.group_m;
  const int band_sz = GROUP_M * p.tiles_n;
  const int band = t / band_sz;
  const int r = t - band * band_sz;
  const int rows = min(GROUP_M, tiles_m - band * GROUP_M);
  tc.m_blk = band * GROUP_M + r % rows;
  tc.n_blk = r / rows;
  return tc;
}

template <int BN, bool kScatter>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                     const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, const DevParams p) {
  constexpr int STAGES = Cfg<BN>::STAGES, STAGE_BYTES = Cfg<BN>::STAGE_BYTES, BAR_OFF = Cfg<BN>::BAR_OFF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nk = p.K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0); prefetch_tmap(&tmB0);
    if (p.nprob > 1) { prefetch_tmap(&tmA1); prefetch_tmap(&tmB1); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 2 * BN);
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int s = 0;
    uint32_t ph = 0;
    for (int t = pair; t < p.total_tiles; t += npairs) {
      const TileCoord tc = decode_tile(p, t);
      const CUtensorMap* ta = tc.pi ? &tmA1 : &tmA0;
      const CUtensorMap* tb = tc.pi ? &tmB1 : &tmB0;
      const int row_a = tc.m_blk * (2 * BM) + static_cast<int>(rank) * BM;
      const int row_b = tc.n_blk * BN + static_cast<int>(rank) * (BN / 2);
      // implicit convolution: this CTA's 128 output pixels start at (img, y0, x0); tap (ky, kx) reads the same activation box
      // shifted by (ky - 1, kx - 1), TMA's out-of-bounds zero fill is the padding ring (and the tail past the last image)
      int img = 0, y0 = 0, x0 = 0;
      if (p.conv_cblk > 0) {
        img = row_a / p.conv_hw;
        const int rem = row_a - img * p.conv_hw;
        y0 = rem / p.conv_w;
        x0 = rem - y0 * p.conv_w;
      }
      int tap = 0, cb = 0;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * STAGE_BYTES);   // bytes of BOTH CTAs land on the leader's barrier
        uint8_t* st = smem + s * STAGE_BYTES;
        if (p.conv_cblk > 0) {
          const int ky = tap / 3, kx = tap - 3 * ky;
          tma_load_4d_2sm(st, ta, &full[s], cb * BK, x0 + kx - 1, y0 + ky - 1, img);
          if (++cb == p.conv_cblk) { cb = 0; ++tap; }
        } else {
          tma_load_2d_2sm(st, ta, &full[s], kb * BK, row_a);
        }
        tma_load_2d_2sm(st + A_BYTES, tb, &full[s], kb * BK, row_b);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA): warp-uniform loop, one elected
    // lane issues (a loop owned by a single divergent thread costs ~19 SASS instructions per tcgen05.mma, attn2_sm100.cu)
    const bool leader = elect_one();
    const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t bar0 = sb + BAR_OFF;                                   // full[STAGES] | empty[STAGES] | tfull[2] | tempty[2]
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int t = pair; t < p.total_tiles; t += npairs) {
      mbar_wait_a(bar0 + (2 * STAGES + 2 + as) * 8, aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * BN;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait_a(bar0 + s * 8, ph);
        tc_fence_after();
        const uint32_t a_addr = sb + s * STAGE_BYTES;
        const uint32_t b_addr = a_addr + A_BYTES;
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss_2sm(d_tmem, make_sdesc(a_addr + k * 32, 16, 1024), make_sdesc(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0);
          umma_commit_2sm_a(bar0 + (STAGES + s) * 8);
          if (kb == nk - 1) umma_commit_2sm_a(bar0 + (2 * STAGES + as) * 8);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs, own 128 accumulator rows)
    const int ew = warp - 4;
    int as = 0;
    uint32_t aph = 0;
    for (int t = pair; t < p.total_tiles; t += npairs) {
      const TileCoord tc = decode_tile(p, t);
      const EpiProblem& pr = p.prob[tc.pi];
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const int row = tc.m_blk * (2 * BM) + static_cast<int>(rank) * BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      epilogue_tile<BN, kScatter>(p.e, pr, taddr, row, tc.n_blk * BN, smem + Cfg<BN>::EPI_STAGE_OFF + ew * 8192);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty[as], 0);   // the leader's MMA thread waits for both CTAs' epilogues
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();          // neither CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

template <int BN>
int launch2(const GemmArgs& a, cudaStream_t stream) {
  DevParams p{};
  p.K = a.K;
  p.tiles_n = a.N / BN;
  p.nprob = a.nprob;
  static int group_m = 0;
  if (group_m == 0) {
    const char* e = std::getenv("UTX_GEMM_GROUP_M");
    group_m = e ? std::atoi(e) : GROUP_M_DEFAULT;
    if (group_m < 1) group_m = GROUP_M_DEFAULT;
  }
  p.group_m = group_m;
  p.conv_cblk = a.conv_c / BK;
  p.conv_w = a.conv_w;
  p.conv_hw = a.conv_h * a.conv_w;
  p.e = EpiParams{a.N, a.epi, a.gelu_col_start, a.out_scale, a.qk_cols, a.cos_t, a.sin_t};
  CUtensorMap tm[4];
  int total = 0;
  for (int i = 0; i < a.nprob; ++i) {
    const GemmProblem& g = a.prob[i];
    EpiProblem& d = p.prob[i];
    d = EpiProblem{g.M, (g.M + 2 * BM - 1) / (2 * BM), g.C, g.ldc, g.bias, g.gate, g.res, g.ldres, g.split_col, g.C2, g.ldc2,
                   g.wq, g.wk, g.row_offset, g.sc_hl, g.sc_rows, g.sc_row_base, g.sc_D,
                   {g.sc_peer[0], g.sc_peer[1], g.sc_peer[2], g.sc_peer[3], g.sc_peer[4], g.sc_peer[5], g.sc_peer[6], g.sc_peer[7]}};
    total += d.tiles_m * p.tiles_n;
    if (a.conv_c > 0) {
      const int bw = a.conv_w < BM ? a.conv_w : BM;
      UTX_TRY(make_tmap_nhwc_bf16(&tm[2 * i], g.A, a.conv_n, a.conv_h, a.conv_w, a.conv_c, bw, BM / bw));
    } else {
      UTX_TRY(make_tmap_2d_bf16(&tm[2 * i], g.A, g.M, a.K, g.lda, BM, BK));
    }
    UTX_TRY(make_tmap_2d_bf16(&tm[2 * i + 1], g.W, a.N, a.K, g.ldw, BN / 2, BK));
  }
  if (a.nprob == 1) { tm[2] = tm[0]; tm[3] = tm[1]; }
  p.total_tiles = total;
  if (total == 0) return 0;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    UTX_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_TOTAL));
    UTX_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_TOTAL));
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = total < max_pairs ? total : max_pairs;
  if (a.prob[0].sc_hl != 0)      // sequence-parallel scatter of q | k | v: its own instantiation (staged stores)
    gemm2_bf16_tn_kernel<BN, true><<<2 * pairs, kThreads, Cfg<BN>::SMEM_TOTAL, stream>>>(tm[0], tm[1], tm[2], tm[3], p);
  else
    gemm2_bf16_tn_kernel<BN, false><<<2 * pairs, kThreads, Cfg<BN>::SMEM_TOTAL, stream>>>(tm[0], tm[1], tm[2], tm[3], p);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// returns -1 when the shape does not fit this kernel (caller falls back to the 1-CTA kernel)
int gemm2_bf16_tn(const GemmArgs& a, cudaStream_t stream) {
  if (a.N % 256 == 0) {
    // few, long tiles (a [4096, 512] x K = 16384 product has 32 tiles of 256 x 256 for 74 CTA pairs): halve the tile width
    long tiles = 0;
    for (int i = 0; i < a.nprob; ++i) tiles += static_cast<long>((a.prob[i].M + 2 * BM - 1) / (2 * BM)) * (a.N / 256);
    if (tiles < num_sms() / 2 && a.qk_cols == 0 && a.prob[0].split_col == 0) return launch2<128>(a, stream);
    return launch2<256>(a, stream);
  }
  if (a.N % 128 == 0 && a.qk_cols == 0) return launch2<128>(a, stream);
  return -1;
}

}  // namespace utx
