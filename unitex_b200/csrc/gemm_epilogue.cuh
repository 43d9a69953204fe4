// Shared epilogue of the two tcgen05 GEMM kernels (gemm_sm100.cu: 1 CTA per 128 x BN tile; gemm2_sm100.cu: CTA pairs).
// One thread owns one accumulator row of the tile (tcgen05.ld 32x32b: lane == row), so row-wise reductions are thread-local:
// besides bias / GELU-tanh / gate*x+residual / fp32 output, columns below `qk_cols` (the q and k thirds of a QKV
// projection) get the per-head RMSNorm(eps 1e-6, weight) + RoPE of the joint attention
// (flux_piplines/texturing/attention_processor.py:56-59,73-76,85-87) fused in, which removes a 240 MB read-modify-write
// pass over the qkv buffer per block.
#pragma once
#include "kernels.h"
#include "ptx.cuh"

namespace utx {

struct EpiProblem {
  int M;
  int tiles_m;
  bf16* C;
  long ldc;
  const bf16* bias;
  const float* gate;
  const bf16* res;
  long ldres;
  int split_col;
  bf16* C2;
  long ldc2;
  const bf16* wq;      // [128] RMSNorm weights of this problem's q / k heads (qk_cols > 0)
  const bf16* wk;
  int row_offset;      // token index of row 0 (img stream of a grouped launch starts at s_txt)
  // Sequence-parallel head scatter (sc_hl = 0: off).  Columns [0, 3*sc_D) are q | k | v with sc_D = heads * 128; head h belongs
  // to peer p = h / sc_hl and the element lands in the all-to-all SEND layout [peer][sc_rows][3][sc_hl * 128]:
  //   C[((p * sc_rows + sc_row_base + row) * 3 + third) * (sc_hl * 128) + (h % sc_hl) * 128 + c]
  // so that every peer's share is one contiguous chunk and, once received, reads as a [S, 3 * sc_hl * 128] qkv matrix.
  int sc_hl, sc_rows, sc_row_base, sc_D;
  // direct mode (sc_peer[0] != nullptr): sc_peer[p] = rank p's attention input [S, 3 * sc_hl * 128] mapped over NVLink; the
  // element goes straight there, at global row sc_row_base + row:  sc_peer[p][((sc_row_base + row) * 3 + third) * w + ...]
  bf16* sc_peer[8];
};
struct EpiParams {
  int N, epi, gelu_col_start;
  float out_scale;
  int qk_cols;         // 0 = off; else 2*D: columns [0, D) are q heads, [D, 2D) k heads, 128 columns per head
  const float* cos_t;  // [S, 128] fp32
  const float* sin_t;
};

template <bool kScatter = true>
__device__ __forceinline__ bf16* epi_dst(const EpiProblem& pr, bf16* crow, int col_shift, int row, int col) {
  if (!kScatter || pr.sc_hl == 0 || col >= 3 * pr.sc_D) return crow + (col - col_shift);
  const int third = col / pr.sc_D, within = col - third * pr.sc_D;
  const int h = within >> 7, c = within & 127;
  const int peer = h / pr.sc_hl, hl = h - peer * pr.sc_hl;
  const long w = static_cast<long>(pr.sc_hl) * 128;
  if (pr.sc_peer[0] != nullptr) {
    bf16* base = pr.sc_peer[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) base = peer == q ? pr.sc_peer[q] : base;      // (no dynamic indexing of a kernel parameter)
    return base + ((static_cast<long>(pr.sc_row_base) + row) * 3 + third) * w + hl * 128 + c;
  }
  return pr.C + ((static_cast<long>(peer) * pr.sc_rows + pr.sc_row_base + row) * 3 + third) * w + hl * 128 + c;
}

// Staging of scattered stores (sequence-parallel q | k | v): with one accumulator row per thread a store instruction writes 32
// rows x 16 B -- 32 half-filled sectors at a 2-6 KB stride, which over NVLink (direct mode: the rows go to a PEER's memory) held
// the QKV GEMM back by 57 % (profiles/r02_summary.md).  Instead a warp parks 128 columns of its 32 rows in 8 KB of shared
// memory (16-byte chunks XOR-swizzled by the row, conflict-free both ways) and writes them out 2 rows x 256 contiguous bytes per
// instruction.
__device__ __forceinline__ void epi_stage16(uint8_t* stage, int lane, int chunk /* 0..15 */, const uint4& v) {
  *reinterpret_cast<uint4*>(stage + lane * 256 + ((chunk ^ (lane & 15)) << 4)) = v;
}
__device__ __forceinline__ void epi_flush128(const EpiProblem& pr, const uint8_t* stage, int row0, int col0, int lane) {
  __syncwarp();
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int idx = i * 32 + lane, r = idx >> 4, kc = idx & 15;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 256 + ((kc ^ (r & 15)) << 4));
    const int row = row0 + r;
    if (row < pr.M) *reinterpret_cast<uint4*>(epi_dst(pr, nullptr, 0, row, col0 + kc * 8)) = v;
  }
  __syncwarp();
}

// stage != nullptr: the packed result is parked in the warp's staging buffer at 16-byte chunk `stage_chunk` instead of stored
template <bool kScatter = false>
__device__ __forceinline__ void epi_store8(const EpiParams& p, const EpiProblem& pr, float (&f)[8], int row, int col,
                                           bf16* crow, int col_shift, const bf16* rrow, uint8_t* stage = nullptr,
                                           int stage_chunk = 0) {
  if (p.epi == EPI_BIAS_GELU) {
    if (col >= p.gelu_col_start) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = gelu_tanh(f[j]);
    }
  } else if (p.epi == EPI_GATE_RES) {
    const float4 g0 = *reinterpret_cast<const float4*>(pr.gate + col);
    const float4 g1 = *reinterpret_cast<const float4*>(pr.gate + col + 4);
    const uint4 r = *reinterpret_cast<const uint4*>(rrow + col);
    f[0] = fmaf(g0.x, f[0], bf16lo(r.x)); f[1] = fmaf(g0.y, f[1], bf16hi(r.x));
    f[2] = fmaf(g0.z, f[2], bf16lo(r.y)); f[3] = fmaf(g0.w, f[3], bf16hi(r.y));
    f[4] = fmaf(g1.x, f[4], bf16lo(r.z)); f[5] = fmaf(g1.y, f[5], bf16hi(r.z));
    f[6] = fmaf(g1.z, f[6], bf16lo(r.w)); f[7] = fmaf(g1.w, f[7], bf16hi(r.w));
  } else if (p.epi == EPI_BIAS_F32) {   // C is float*, ldc in floats (attention scores of the VAE mid block)
    float* c32 = reinterpret_cast<float*>(pr.C) + static_cast<long>(row) * pr.ldc + col;
    *reinterpret_cast<float4*>(c32) = make_float4(f[0] * p.out_scale, f[1] * p.out_scale, f[2] * p.out_scale, f[3] * p.out_scale);
    *reinterpret_cast<float4*>(c32 + 4) = make_float4(f[4] * p.out_scale, f[5] * p.out_scale, f[6] * p.out_scale, f[7] * p.out_scale);
    return;
  }
  const uint4 packed = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  if (kScatter && stage) epi_stage16(stage, threadIdx.x & 31, stage_chunk, packed);
  else *reinterpret_cast<uint4*>(epi_dst<kScatter>(pr, crow, col_shift, row, col)) = packed;
}

__device__ __forceinline__ void epi_add_bias8(const bf16* bias, int col, float (&f)[8]) {
  if (bias) {
    const uint4 b = *reinterpret_cast<const uint4*>(bias + col);
    f[0] += bf16lo(b.x); f[1] += bf16hi(b.x); f[2] += bf16lo(b.y); f[3] += bf16hi(b.y);
    f[4] += bf16lo(b.z); f[5] += bf16hi(b.z); f[6] += bf16lo(b.w); f[7] += bf16hi(b.w);
  }
}

// taddr: TMEM address of this warp's lane quarter at the tile's first accumulator column.  All 32 lanes must call this.
// kScatter: the instantiation the sequence-parallel launches use (pr.sc_hl != 0).  stage: 8 KB of shared memory owned by the
// calling warp, used when this tile's columns are scattered.  The plain instantiation carries none of that code: with the
// staging decided at run time the single-block q|k|v|mlp GEMM lost 4-5 % (register pressure in the fused norm-RoPE path).
template <int BN, bool kScatter = false>
__device__ __forceinline__ void epilogue_tile(const EpiParams& p, const EpiProblem& pr, uint32_t taddr, int row, int n0,
                                              uint8_t* stage = nullptr) {
  const bool row_ok = row < pr.M;
  const int lane = threadIdx.x & 31;
  const bool staged = kScatter && stage != nullptr && pr.sc_hl != 0 && n0 < 3 * pr.sc_D;   // tile-uniform: tiles do not straddle 3 * sc_D
  const int row0 = row - lane;
  bf16* crow;
  int col_shift = 0;
  if (pr.split_col > 0 && n0 >= pr.split_col) {
    crow = pr.C2 + static_cast<long>(row) * pr.ldc2;
    col_shift = pr.split_col;
  } else {
    crow = pr.C + static_cast<long>(row) * pr.ldc;
  }
  const bf16* rrow = pr.res ? pr.res + static_cast<long>(row) * pr.ldres : nullptr;
  if (BN >= 128 && p.qk_cols > 0 && n0 < p.qk_cols) {
    // ---- q / k heads: RMSNorm over the head's 128 columns, weight, RoPE; one head = 4 chunks of 32 accumulator columns
    const int D = p.qk_cols >> 1;
#pragma unroll 1
    for (int hc = 0; hc < BN / 128; ++hc) {
      const int col0 = n0 + hc * 128;
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(taddr + hc * 128 + c * 32, v[c]);
      tmem_ld_wait();
      if (!row_ok && !staged) continue;          // (staged: every lane takes part in the warp-wide flush; rows >= M are not written)
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c][g * 8 + j]);
          epi_add_bias8(pr.bias, col0 + c * 32 + g * 8, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            ss = fmaf(f[j], f[j], ss);
            v[c][g * 8 + j] = __float_as_uint(f[j]);
          }
        }
      const float rs = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
      const bf16* w = col0 >= D ? pr.wk : pr.wq;
      const long tok = static_cast<long>(row) + pr.row_offset;
      const float* ct = p.cos_t + tok * 128;
      const float* st = p.sin_t + tok * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int e = c * 32 + g * 8;   // element offset inside the head
          const uint4 wr = *reinterpret_cast<const uint4*>(w + e);
          const float wv[8] = {bf16lo(wr.x), bf16hi(wr.x), bf16lo(wr.y), bf16hi(wr.y), bf16lo(wr.z), bf16hi(wr.z), bf16lo(wr.w), bf16hi(wr.w)};
          const float4 c0 = *reinterpret_cast<const float4*>(ct + e), c1 = *reinterpret_cast<const float4*>(ct + e + 4);
          const float4 s0 = *reinterpret_cast<const float4*>(st + e), s1 = *reinterpret_cast<const float4*>(st + e + 4);
          const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          float x[8], o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[c][g * 8 + j]) * rs * wv[j];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            o[j] = x[j] * cv[j] - x[j + 1] * sv[j];
            o[j + 1] = x[j + 1] * cv[j + 1] + x[j] * sv[j + 1];
          }
          const uint4 packed = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          if (staged) epi_stage16(stage, lane, e >> 3, packed);
          else *reinterpret_cast<uint4*>(epi_dst<kScatter>(pr, crow, col_shift, row, col0 + e)) = packed;
        }
      if (staged) epi_flush128(pr, stage, row0, col0, lane);
    }
    return;
  }
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld_wait();
    const int col0 = n0 + c * 32;
    if (row_ok || staged) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = col0 + g * 8;
        if (col < p.N) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
          epi_add_bias8(pr.bias, col, f);
          epi_store8<kScatter>(p, pr, f, row, col, crow, col_shift, rrow, staged ? stage : nullptr, (c & 3) * 4 + g);
        }
      }
    }
    if (staged && (c & 3) == 3) epi_flush128(pr, stage, row0, n0 + (c - 3) * 32, lane);   // 128 columns = one head's v (or q / k) block
  }
}

}  // namespace utx
