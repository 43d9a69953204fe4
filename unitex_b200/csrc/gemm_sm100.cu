// Persistent warp-specialised bf16 GEMM for sm_100a:  C = epilogue(A[M,K] @ W[N,K]^T + bias).
//
//   warp 0      TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer    (one thread, tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16 per instruction)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4..7  epilogue      (tcgen05.ld 32x32b -> bias / GELU-tanh / gate*x+residual -> bf16 -> global)
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.  Up to two
// problems that share (N, K, epilogue) are scheduled in one launch (the txt and img streams of an MM-DiT double block),
// so the 512-token txt stream does not cost a mostly idle wave of its own.  Tiles are walked in bands of GROUP_M
// row-blocks (row-block fastest) so that concurrently resident CTAs share one weight column-block out of L2 and the
// activation band stays L2 resident.
//
// Every Linear of the FLUX DiT (reference call site flux_piplines/texturing/pipeline.py:646) runs
// through this kernel; LoRA adapters are merged into W beforehand (a5).
#include <cstdlib>

#include "common.h"
#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace utx {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GROUP_M = 16;
constexpr int kThreads = 256;

struct DevParams {
  int K, tiles_n, nprob, total_tiles;
  int conv_cblk, conv_w, conv_hw;   // implicit 3x3 convolution: channel blocks per tap (0 = plain GEMM), image width, pixels per image
  EpiParams e;
  EpiProblem prob[2];
};

struct TileCoord {
  int pi, m_blk, n_blk;
};

__device__ __forceinline__ TileCoord decode_tile(const DevParams& p, int t) {
  TileCoord tc;
  tc.pi = 0;
  const int t0 = p.prob[0].tiles_m * p.tiles_n;
  if (p.nprob > 1 && t >= t0) {
    tc.pi = 1;
    t -= t0;
  }
  const int tiles_m = p.prob[tc.pi].tiles_m;
  const int band_sz = GROUP_M * p.tiles_n;
  const int band = t / band_sz;
  const int r = t - band * band_sz;
  const int rows = min(GROUP_M, tiles_m - band * GROUP_M);
  tc.m_blk = band * GROUP_M + r % rows;
  tc.n_blk = r / rows;
  return tc;
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;   // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const DevParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  const int nk = p.K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB0);
    if (p.nprob > 1) {
      prefetch_tmap(&tmA1);
      prefetch_tmap(&tmB1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      const CUtensorMap* ta = tc.pi ? &tmA1 : &tmA0;
      const CUtensorMap* tb = tc.pi ? &tmB1 : &tmB0;
      // implicit convolution: the tile's 128 output pixels start at (img, y0, x0); tap (ky, kx) reads the same box of the
      // activation shifted by (ky - 1, kx - 1), out-of-range rows / columns arrive as zeros (= the padding ring)
      int img = 0, y0 = 0, x0 = 0;
      if (p.conv_cblk > 0) {
        const int p0 = tc.m_blk * BM;
        img = p0 / p.conv_hw;
        const int rem = p0 - img * p.conv_hw;
        y0 = rem / p.conv_w;
        x0 = rem - y0 * p.conv_w;
      }
      int tap = 0, cb = 0;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], L::STAGE_BYTES);
        uint8_t* st = smem + s * L::STAGE_BYTES;
        if (p.conv_cblk > 0) {
          const int ky = tap / 3, kx = tap - 3 * ky;
          tma_load_4d(st, ta, &full[s], cb * BK, x0 + kx - 1, y0 + ky - 1, img);
          if (++cb == p.conv_cblk) { cb = 0; ++tap; }
        } else {
          tma_load_2d(st, ta, &full[s], kb * BK, tc.m_blk * BM);
        }
        tma_load_2d(st + L::A_BYTES, tb, &full[s], kb * BK, tc.n_blk * BN);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: the whole warp runs the loop (waits and
    // descriptor arithmetic stay warp-uniform), one elected lane issues -- a loop owned by a single divergent thread costs
    // ~19 SASS instructions per tcgen05.mma (see attn2_sm100.cu)
    const bool leader = elect_one();
    const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t bar0 = __shfl_sync(0xffffffffu, smem_u32(full), 0);   // full[STAGES] | empty[STAGES] | tfull[2] | tempty[2]
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait_a(bar0 + (2 * STAGES + 2 + as) * 8, aph ^ 1);   // epilogue drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * BN;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait_a(bar0 + s * 8, ph);
        tc_fence_after();
        const uint32_t a_addr = sb + s * L::STAGE_BYTES;
        const uint32_t b_addr = a_addr + L::A_BYTES;
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_sdesc(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_sdesc(b_addr + k * 32, 16, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit_a(bar0 + (STAGES + s) * 8);                                 // smem stage reusable once these MMAs retire
          if (kb == nk - 1) umma_commit_a(bar0 + (2 * STAGES + as) * 8);          // accumulator complete
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (warp w owns TMEM lanes 32*(w%4)..)
    const int ew = warp - 4;
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      const EpiProblem& pr = p.prob[tc.pi];
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const int row = tc.m_blk * BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
      epilogue_tile<BN>(p.e, pr, taddr, row, tc.n_blk * BN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, int STAGES>
int launch(const GemmArgs& a, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  DevParams p{};
  p.K = a.K;
  p.tiles_n = (a.N + BN - 1) / BN;
  p.nprob = a.nprob;
  p.conv_cblk = a.conv_c / BK;
  p.conv_w = a.conv_w;
  p.conv_hw = a.conv_h * a.conv_w;
  p.e = EpiParams{a.N, a.epi, a.gelu_col_start, a.out_scale, a.qk_cols, a.cos_t, a.sin_t};
  CUtensorMap tm[4];
  int total = 0;
  for (int i = 0; i < a.nprob; ++i) {
    const GemmProblem& g = a.prob[i];
    EpiProblem& d = p.prob[i];
    d = EpiProblem{g.M, (g.M + BM - 1) / BM, g.C, g.ldc, g.bias, g.gate, g.res, g.ldres, g.split_col, g.C2, g.ldc2, g.wq, g.wk,
                   g.row_offset, g.sc_hl, g.sc_rows, g.sc_row_base, g.sc_D,
                   {g.sc_peer[0], g.sc_peer[1], g.sc_peer[2], g.sc_peer[3], g.sc_peer[4], g.sc_peer[5], g.sc_peer[6], g.sc_peer[7]}};
    total += d.tiles_m * p.tiles_n;
    if (a.conv_c > 0) {
      const int bw = a.conv_w < BM ? a.conv_w : BM;
      UTX_TRY(make_tmap_nhwc_bf16(&tm[2 * i], g.A, a.conv_n, a.conv_h, a.conv_w, a.conv_c, bw, BM / bw));
    } else {
      UTX_TRY(make_tmap_2d_bf16(&tm[2 * i], g.A, g.M, a.K, g.lda, BM, BK));
    }
    UTX_TRY(make_tmap_2d_bf16(&tm[2 * i + 1], g.W, a.N, a.K, g.ldw, BN, BK));
  }
  if (a.nprob == 1) {
    tm[2] = tm[0];
    tm[3] = tm[1];
  }
  p.total_tiles = total;
  if (total == 0) return 0;
  auto kern = gemm_bf16_tn_kernel<BN, STAGES>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    UTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  }
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, kThreads, L::TOTAL, stream>>>(tm[0], tm[1], tm[2], tm[3], p);
  UTX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int gemm2_bf16_tn(const GemmArgs& a, cudaStream_t stream);   // 2-CTA pairs, gemm2_sm100.cu; -1 = shape not supported

int gemm_bf16_tn(const GemmArgs& a_in, cudaStream_t stream) {
  UTX_CHECK(a_in.nprob == 1 || a_in.nprob == 2, "gemm: nprob must be 1 or 2");
  // empty problems (no text tokens, zero-row calls) are dropped here: a TMA descriptor cannot describe a 0-row tensor
  GemmArgs a = a_in;
  if (a.nprob == 2 && a.prob[1].M == 0) a.nprob = 1;
  if (a.nprob == 2 && a.prob[0].M == 0) { a.prob[0] = a.prob[1]; a.nprob = 1; }
  if (a.nprob == 1 && a.prob[0].M == 0) return 0;
  UTX_CHECK(a.K % BK == 0 && a.K > 0, "gemm: K must be a positive multiple of 64");
  UTX_CHECK(a.N % 8 == 0 && a.N > 0, "gemm: N must be a positive multiple of 8");
  for (int i = 0; i < a.nprob; ++i) {
    const GemmProblem& g = a.prob[i];
    UTX_CHECK(g.M >= 0, "gemm: negative M");
    UTX_CHECK(g.ldc % 8 == 0 && (g.res == nullptr || g.ldres % 8 == 0), "gemm: ldc/ldres must be multiples of 8");
    UTX_CHECK(a.epi != EPI_BIAS_F32 || g.split_col == 0, "gemm: fp32 output cannot be combined with a column split");
    UTX_CHECK(a.qk_cols == 0 || (g.wq && g.wk), "gemm: qk_cols needs RMSNorm weights");
    UTX_CHECK(a.epi != EPI_GATE_RES || (g.gate && g.res), "gemm: EPI_GATE_RES needs gate and res");
    UTX_CHECK(g.split_col == 0 || (g.split_col % 256 == 0 && g.C2 && g.ldc2 % 8 == 0), "gemm: bad column split");
  }
  // Default: cta_group::2 kernel (256x256 tiles per CTA pair, gemm2_sm100.cu) whenever N % 256 == 0 -- 2-3 % faster at the
  // DiT shapes (profiles/r01_microbench_gemm.json).  UTX_GEMM_IMPL=1 forces the 1-CTA kernel of this file.
  const char* impl = getenv("UTX_GEMM_IMPL");
  if ((impl == nullptr || impl[0] != '1') && a.N % 128 == 0) {   // plain and implicit-convolution problems alike; N % 256 != 0 -> 256 x 128 pair tiles
    const int r = gemm2_bf16_tn(a, stream);
    if (r >= 0) return r;
  }
  UTX_CHECK(a.qk_cols == 0 || (a.qk_cols % 256 == 0 && a.N % 128 == 0 && a.cos_t && a.sin_t && a.epi != EPI_BIAS_F32 &&
                               a.epi != EPI_GATE_RES),
            "gemm: bad fused q/k norm-rope configuration");
  UTX_CHECK(a.prob[0].sc_hl == 0, "gemm: the sequence-parallel q|k|v scatter is only built into the 2-CTA kernel (N % 128 == 0, UTX_GEMM_IMPL unset)");
  if (a.N % 256 == 0) return launch<256, 4>(a, stream);
  if (a.N % 128 == 0) return launch<128, 6>(a, stream);
  return launch<64, 8>(a, stream);
}

int conv3x3_nhwc(const bf16* x, int N, int H, int W, int C, const bf16* w, const bf16* bias, int Cout, bf16* y, long ldy,
                 const float* gate, const bf16* res, long ldres, cudaStream_t stream) {
  UTX_CHECK(N > 0 && H > 0 && W > 0, "conv3x3: empty input");
  UTX_CHECK(C % 64 == 0 && C > 0, "conv3x3: Cin must be a multiple of 64");
  UTX_CHECK(W % BM == 0 || (BM % W == 0 && H % (BM / W) == 0), "conv3x3: image width must tile 128-pixel row blocks");
  UTX_CHECK((res == nullptr) == (gate == nullptr), "conv3x3: the residual epilogue needs both gate and res");
  GemmArgs a{};
  a.N = Cout; a.K = 9 * C; a.epi = res ? EPI_GATE_RES : EPI_BIAS; a.nprob = 1;
  a.conv_n = N; a.conv_h = H; a.conv_w = W; a.conv_c = C;
  a.prob[0] = GemmProblem{x, static_cast<long>(C), w, 9L * C, N * H * W, y, ldy, bias, gate, res, ldres, 0, nullptr, 0,
                          nullptr, nullptr, 0};
  return gemm_bf16_tn(a, stream);
}

}  // namespace utx
