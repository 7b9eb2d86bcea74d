// Epilogue shared by the single-CTA (gemm.cu) and CTA-pair (gemm2.cu) GEMM kernels: argument block, the aux-aware tile
// order and the per-tile TMEM -> registers -> global pass (bias / column scale / multiply-by-aux / residual / QuickGELU and its
// derivative / fp32 and fp16 hi|lo outputs).  8 epilogue warps: warp w owns TMEM lanes [32 (w % 4), +32) and column half
// (w - 4) / 4 of the tile, in 16-column chunks, software-pipelined one chunk deep.
#pragma once
#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_CW = 16;  // epilogue chunk width (columns)

struct EpiParams {
  const float* bias;
  const float* residual;
  const __half* aux16;
  int aux_rows;
  int ld_aux;
  __half* out_aux16;
  int ld_out_aux;
  float* out_f32;
  int ld_out;
  __half* out_f16;
  int ld_out16;
  int out_f16_splits;
  int act;
  int scale_cols;
  float scale;
  int wide;  // every row of every side input / output is 32-byte aligned: 256-bit global accesses
  // aux-aware tile order (SEMABS_ACT_MUL_AUX16 with aux_rows < M): rows r, r + aux_rows, r + 2 aux_rows ... multiply by the
  // same aux row, so their tiles are visited back to back and the aux tile is fetched from HBM once instead of once per
  // repeat (ncu round 1: 1.34 GB read for 0.34 GB of operands on the fc2 dgrad)
  int raster_rows;    // aux_rows, or 0 = plain row-major tile order
  int raster_groups;  // ceil(aux_rows / 128)
  int raster_reps;    // M / aux_rows
};

// virtual tile index -> (m_blk, n_blk); false = this virtual index maps to no tile (skipped by every role alike)
__device__ __forceinline__ bool tile_coords(const EpiParams& ep, int t, int num_m, int num_n, int tile_rows, int& m_blk, int& n_blk) {
  n_blk = t % num_n;
  const int u = t / num_n;
  if (ep.raster_rows == 0) {
    m_blk = u;
    return true;
  }
  const int p = u % ep.raster_reps, g = u / ep.raster_reps;
  m_blk = g + int((long long)p * ep.raster_rows / tile_rows);
  const int next = (p + 1 == ep.raster_reps) ? num_m : int((long long)(p + 1) * ep.raster_rows / tile_rows);
  return m_blk < next;
}

// tmem_acc: TMEM address (lane 0) of this tile's accumulator; row0 / col_tile0: first output row of THIS CTA's 128 rows / first
// column of the tile.  Called by warps 4..11.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const EpiParams& ep, uint32_t tmem_acc, int row0, int col_tile0, int M, int N,
                                                   int warp, int lane) {
  const int q = warp & 3, chalf = (warp - 4) >> 2;
  constexpr int CW = GEMM_CW;
  constexpr int HALF = BN / 2;       // columns per epilogue warp
  constexpr int NC = HALF / CW;      // chunks per warp and tile
  const int row = row0 + q * 32 + lane;
  const bool row_ok = row < M;
  const __half* aux_row = nullptr;
  if (ep.act == SEMABS_ACT_MUL_AUX16 && row_ok) aux_row = ep.aux16 + size_t(row % ep.aux_rows) * ep.ld_aux;
  // Software-pipelined over 16-column chunks: the TMEM load and the global side inputs (aux / residual) of chunk
  // c+1 are in flight while chunk c is converted and stored.
  const int colbase = col_tile0 + chalf * HALF;
  const uint32_t t_addr = tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(chalf * HALF);
  const float* res_row = (ep.residual && row_ok) ? ep.residual + size_t(row) * ep.ld_out : nullptr;
  const bool wide = ep.wide != 0;
  auto load_side = [&](int c, uint32_t(&au)[CW / 2], uint32_t(&rs)[CW]) {
    const int col0 = colbase + c * CW;
    if (col0 < N) {
      if (aux_row) ld_row_words<CW / 2>(aux_row + col0, au, wide);
      if (res_row) ld_row_words<CW>(res_row + col0, rs, wide);
    }
  };
  auto process = [&](int c, const uint32_t(&r)[CW], const uint32_t(&au)[CW / 2], const uint32_t(&rs)[CW]) {
    const int col0 = colbase + c * CW;
    if (row_ok && col0 < N) {
      float v[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.bias) {
#pragma unroll
        for (int j = 0; j < CW; j += 4) {
          float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
          v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
        }
      }
      if (col0 < ep.scale_cols) {
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (col0 + j < ep.scale_cols) v[j] *= ep.scale;
      }
      if (aux_row) {
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&au[j]));
          v[2 * j] *= f.x, v[2 * j + 1] *= f.y;
        }
      }
      if (res_row) {
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] += __uint_as_float(rs[j]);
      }
      if (ep.out_f32) {
        st_row_words<CW>(ep.out_f32 + size_t(row) * ep.ld_out + col0, reinterpret_cast<const uint32_t*>(v), wide);
      }
      if (ep.out_f16) {
        __align__(16) __half2 h[CW / 2];
        if (ep.act == SEMABS_ACT_QUICKGELU) {
          // one sigmoid gives both the activation and (for the backward sweep) its derivative
#pragma unroll
          for (int j = 0; j < CW; j += 2) {
            // (fast sigmoid: with expf + IEEE division the 16 sigmoids of a chunk were 60 % of the kernel's instructions and
            // the epilogue warps, two per scheduler, could not keep up with the K = 1024 main loop: tensor pipe 56 %)
            const float s0 = sigmoidf_fast(1.702f * v[j]), s1 = sigmoidf_fast(1.702f * v[j + 1]);
            h[j >> 1] = __floats2half2_rn(s0 + 1.702f * v[j] * s0 * (1.0f - s0), s1 + 1.702f * v[j + 1] * s1 * (1.0f - s1));
            v[j] *= s0, v[j + 1] *= s1;
          }
          if (ep.out_aux16) {
            st_row_words<CW / 2>(ep.out_aux16 + size_t(row) * ep.ld_out_aux + col0, reinterpret_cast<const uint32_t*>(h), wide);
          }
        }
        __half* o = ep.out_f16 + size_t(row) * ep.ld_out16 + col0;
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        st_row_words<CW / 2>(o, reinterpret_cast<const uint32_t*>(h), wide);
        if (ep.out_f16_splits == 2) {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) {
            float2 f = __half22float2(h[j]);
            h[j] = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
          }
          st_row_words<CW / 2>(o + N, reinterpret_cast<const uint32_t*>(h), wide);
        }
      }
    }
  };
  uint32_t r0[CW], r1[CW];
  uint32_t a0[CW / 2], a1[CW / 2];
  uint32_t s0[CW], s1[CW];
  tmem_ld_32x32b_x16(t_addr, r0);
  load_side(0, a0, s0);
#pragma unroll 1
  for (int c = 0; c < NC; c += 2) {
    tc_wait_ld();
    if (c + 1 < NC) {
      tmem_ld_32x32b_x16(t_addr + uint32_t((c + 1) * CW), r1);
      load_side(c + 1, a1, s1);
    }
    process(c, r0, a0, s0);
    if (c + 1 < NC) {
      tc_wait_ld();
      if (c + 2 < NC) {
        tmem_ld_32x32b_x16(t_addr + uint32_t((c + 2) * CW), r0);
        load_side(c + 2, a0, s0);
      }
      process(c + 1, r1, a1, s1);
    }
  }
}

}  // namespace sb
