// ViT attention on the 5th-generation tensor cores (tcgen05 + TMEM), one CTA per (sequence, head), T <= 272 keys.
//
// Forward (reference: MultiheadAttention inside ResidualAttentionBlock.attention, CLIP/clip/model_explainability.py:
// 206-236 and auxiliary.multi_head_attention_forward, CLIP/clip/auxiliary.py:306-345 — scaled QK^T, softmax, AV):
//
//   TMA  : Q tile [128 x 64], K and V [272 x 64] of the head, fp16 hi (+ lo) parts, 128B-swizzled
//   MMA 1: S[128 x 272] = Q K^T           (SS, K-major operands, fp32 in TMEM; hi*hi + lo*hi + hi*lo)
//   SIMT : one thread per query row reads its S row from TMEM (lane = row: max / sum need no shuffles), writes the
//          normalised probabilities to HBM (fp16, one full 32 B sector per store) and back into TMEM — in place over
//          S — as packed fp16 hi / lo A operands
//   MMA 2: O[128 x 64] = P V              (TS: A from TMEM, B = V as an MN-major smem operand; 3 split terms)
//   SIMT : O row -> HBM (fp32 and fp16 hi|lo for the out-projection GEMM)
//
// The mma.sync kernels in vit_attn.cu keep the score strip in registers (255 regs, 6 warps / SM) and were latency
// bound at ~1/4 of their own tensor rate; here the strip lives in TMEM and a 128-row tile costs ~1.6 K tensor cycles
// per product.
#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int TC_HD = 64;                          // head dim = one 128-byte swizzle row of fp16
constexpr int TC_BOX_ROWS = 136;                   // TMA box rows; two boxes = 272 rows
constexpr int TC_BOX_BYTES = TC_BOX_ROWS * 128;    // 17408 = 17 swizzle atoms
constexpr int TC_KV_BYTES = 2 * TC_BOX_BYTES;      // 272 x 64 fp16
constexpr int TC_MAX_T = 272;
constexpr int TC_COL_S = 0;                        // TMEM columns: S / P strip
constexpr int TC_COL_O = 288;                      // O (forward), dQ / dK (backward)
constexpr int TC_COL_O2 = 352;                     // dV (backward, column pass)
constexpr int TC_THREADS = 256;

// smem descriptors. K-major SW128 (row = 128 B): SBO = 1024 (8-row atom), K step of 16 elements = +32 B.
// MN-major SW128 (the smem row holds 64 consecutive M/N elements of ONE k): SBO = 1024 = next group of 8 k rows,
// LBO = next block of 64 M/N elements (unused when N == 64), K step of 16 = 16 rows = +2048 B.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return make_smem_desc(saddr, 16, 1024, SW_128B); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo) { return make_smem_desc(saddr, lbo, 1024, SW_128B); }

__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void split_pack(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------
// Self-test of the two operand forms the attention kernels add on top of the GEMM's: A read from TMEM (written with
// tcgen05.st as packed fp16 pairs) and an MN-major B tile.  D[128 x 64] = A[128 x Kd] * Bm[Kd x 64].
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) ts_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ Bm,
                                                                     float* __restrict__ D, int Kd, int lbo, int sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 256 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // B tile, 128B-swizzled rows: 16-byte chunk c of row k lands in chunk c ^ (k & 7)
  for (int idx = threadIdx.x; idx < Kd * 8; idx += blockDim.x) {
    const int k = idx >> 3, c = idx & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(Bm + size_t(k) * 64 + c * 8);
    *reinterpret_cast<uint4*>(smem + k * 128 + ((c ^ (k & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 4) {
    const int q = warp & 3, m = q * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    for (int c = 0; c < Kd / 16; ++c) {
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = *reinterpret_cast<const uint32_t*>(A + size_t(m) * Kd + c * 16 + 2 * e);
      tmem_st_32x32b_x8(t_row + uint32_t(c * 8), w);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = make_idesc_f16(128, 64, false, true);
    const uint64_t db0 = make_smem_desc(smem_u32(smem), uint32_t(lbo), uint32_t(sbo), SW_128B);
    for (int s = 0; s < Kd / 16; ++s)
      umma_f16_ts_elect(tmem_base + 256, tmem_base + uint32_t(s * 8), db0 + uint64_t(s) * (2048 >> 4), idesc, s > 0, leader);
    umma_commit_elect(bar, leader);
  }
  if (warp >= 4) {
    const int q = warp & 3, m = q * 32 + lane;
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t r[32];
    for (int c = 0; c < 2; ++c) {
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(256 + c * 32), r);
      tc_wait_ld();
      for (int e = 0; e < 32; ++e) D[size_t(m) * 64 + c * 32 + e] = __uint_as_float(r[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
struct AttnFwdTcArgs {
  __half* probs16;  // [B*H, T, ldp] or null
  int ldp;
  float* o32;       // [B*T, d] or null
  __half* o16;      // [B*T, o_splits*d] or null
  int o_splits;
  int T, H, d, causal;
  int in_splits;    // qkv16 rows are [hi (3d) | lo (3d)] when 2
};

struct TcSmem {
  static constexpr int Q_HI = 0;
  static constexpr int Q_LO = Q_HI + TC_BOX_BYTES;
  static constexpr int K_HI = Q_LO + TC_BOX_BYTES;
  static constexpr int K_LO = K_HI + TC_KV_BYTES;
  static constexpr int V_HI = K_LO + TC_KV_BYTES;
  static constexpr int V_LO = V_HI + TC_KV_BYTES;
  static constexpr int BARS = V_LO + TC_KV_BYTES;
  static constexpr int TOTAL = BARS + 128 + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, AttnFwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcSmem::BARS);
  uint64_t *bar_kv = bars, *bar_q = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4, *bar_free = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int T = a.T, d = a.d;
  const int row0 = b * T;
  const int n_mt = (T + 127) / 128;
  const int ncol = (T + 15) & ~15;
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;
  const bool split = a.in_splits == 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(bar_kv, 1), mbar_init(bar_q, 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, 128), mbar_init(bar_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA + MMA issue =====
    const uint32_t leader = elect_one() ? 1u : 0u;
    const int colQ = h * TC_HD, colK = d + h * TC_HD, colV = 2 * d + h * TC_HD, lo_off = 3 * d;
    if (leader) {
      mbar_arrive_expect_tx(bar_kv, (split ? 4u : 2u) * TC_KV_BYTES);
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_2d(smem + TcSmem::K_HI + bx * TC_BOX_BYTES, &tm, bar_kv, colK, row0 + bx * TC_BOX_ROWS);
        tma_load_2d(smem + TcSmem::V_HI + bx * TC_BOX_BYTES, &tm, bar_kv, colV, row0 + bx * TC_BOX_ROWS);
        if (split) {
          tma_load_2d(smem + TcSmem::K_LO + bx * TC_BOX_BYTES, &tm, bar_kv, lo_off + colK, row0 + bx * TC_BOX_ROWS);
          tma_load_2d(smem + TcSmem::V_LO + bx * TC_BOX_BYTES, &tm, bar_kv, lo_off + colV, row0 + bx * TC_BOX_ROWS);
        }
      }
    }
    const uint32_t sbase = smem_u32(smem);
    const uint64_t dQh = desc_kmajor(sbase + TcSmem::Q_HI), dQl = desc_kmajor(sbase + TcSmem::Q_LO);
    const uint64_t dKh = desc_kmajor(sbase + TcSmem::K_HI), dKl = desc_kmajor(sbase + TcSmem::K_LO);
    const uint64_t dVh = desc_mnmajor(sbase + TcSmem::V_HI, 16), dVl = desc_mnmajor(sbase + TcSmem::V_LO, 16);
    const uint32_t idesc_s1 = make_idesc_f16(128, n1), idesc_s2 = make_idesc_f16(128, n2 ? n2 : 16);
    constexpr uint32_t idesc_pv = make_idesc_f16(128, TC_HD, false, true);
    const uint32_t tS = tmem_base + TC_COL_S, tO = tmem_base + TC_COL_O;
    for (int mt = 0; mt < n_mt; ++mt) {
      if (leader) {
        mbar_arrive_expect_tx(bar_q, (split ? 2u : 1u) * TC_BOX_BYTES);
        tma_load_2d(smem + TcSmem::Q_HI, &tm, bar_q, colQ, row0 + mt * 128);
        if (split) tma_load_2d(smem + TcSmem::Q_LO, &tm, bar_q, lo_off + colQ, row0 + mt * 128);
      }
      mbar_wait(bar_q, mt & 1);
      if (mt == 0) mbar_wait(bar_kv, 0);
      if (mt > 0) mbar_wait(bar_free, (mt - 1) & 1);
      tc_fence_after();
      // S = Q K^T : hi*hi (+ lo*hi + hi*lo)
      const int terms = split ? 3 : 1;
      for (int term = 0; term < terms; ++term) {
        const uint64_t da = term == 1 ? dQl : dQh, db = term == 2 ? dKl : dKh;
#pragma unroll
        for (int k = 0; k < TC_HD / 16; ++k) {
          const uint32_t acc = (term | k) != 0;
          umma_f16_elect(tS, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_s1, acc, leader);
          if (n2) umma_f16_elect(tS + 256, da + uint64_t(2 * k), db + uint64_t(2 * k) + uint64_t((256 * 128) >> 4), idesc_s2, acc, leader);
        }
      }
      umma_commit_elect(bar_s, leader);
      mbar_wait(bar_p, mt & 1);
      tc_fence_after();
      // O = P V : A = packed fp16 P in TMEM (k-step s: hi at column 32 (s/2) + 8 (s%2), lo 16 columns further)
      for (int s = 0; s < ncol / 16; ++s) {
        const uint32_t a_hi = tS + uint32_t(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
        const uint64_t kadv = uint64_t(s) * (2048 >> 4);
        umma_f16_ts_elect(tO, a_hi, dVh + kadv, idesc_pv, s > 0, leader);
        if (split) {
          umma_f16_ts_elect(tO, a_lo, dVh + kadv, idesc_pv, 1, leader);
          umma_f16_ts_elect(tO, a_hi, dVl + kadv, idesc_pv, 1, leader);
        }
      }
      umma_commit_elect(bar_o, leader);
    }
  } else if (warp >= 4) {
    // ===== softmax + epilogue: thread = query row =====
    const int q = warp & 3;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    for (int mt = 0; mt < n_mt; ++mt) {
      const int i = mt * 128 + q * 32 + lane;
      const bool valid = i < T;
      const int jmax = a.causal ? (i < T ? i + 1 : T) : T;  // keys [0, jmax) take part
      mbar_wait(bar_s, mt & 1);
      tc_fence_after();
      // pass 1: row maximum
      float mx = -INFINITY;
      for (int c = 0; c < ncol; c += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c), r);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (c + e < jmax) mx = fmaxf(mx, __uint_as_float(r[e]));
      }
      // pass 2: e = exp(s - max) kept in place, row sum
      const float mb = mx * LOG2E;
      float sum = 0.f;
      for (int c = 0; c < ncol; c += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c), r);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float ev = (c + e < jmax) ? fast_exp2(fmaf(__uint_as_float(r[e]), LOG2E, -mb)) : 0.f;
          sum += ev;
          r[e] = __float_as_uint(ev);
        }
        tmem_st_32x32b_x16(t_row + uint32_t(TC_COL_S + c), r);
      }
      tc_wait_st();
      // pass 3: normalise, write probabilities to HBM, pack fp16 hi / lo A operands over the strip (32 columns at a
      // time: the packed words of a 32-column chunk land inside the same 32 columns)
      const float inv = 1.0f / sum;
      __half* prow = (a.probs16 && valid) ? a.probs16 + (size_t(bh) * T + i) * a.ldp : nullptr;
      for (int c = 0; c < ncol; c += 32) {
        const int w = (ncol - c >= 32) ? 32 : 16;
        uint32_t r[32];
        tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c), *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        if (w == 32) tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c + 16), *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tc_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = __uint_as_float(r[2 * e]) * inv, p1 = __uint_as_float(r[2 * e + 1]) * inv;
          split_pack(p0, p1, hi[e], lo[e]);
        }
        if (prow) {
          if (c < a.ldp) st_global_256(prow + c, hi);
          if (w == 32 && c + 16 < a.ldp) st_global_256(prow + c + 16, hi + 8);
        }
        // k-step s = c/16 (+1): hi -> columns c + 8 (s%2), lo -> 16 further
        tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c), hi);
        tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c + 16), lo);
        if (w == 32) {
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c + 8), hi + 8);
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c + 24), lo + 8);
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
      // epilogue
      mbar_wait(bar_o, mt & 1);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O), *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O + 32), *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_free);
      if (valid) {
        const size_t row = size_t(row0) + i;
        if (a.o32) {
          float* dst = a.o32 + row * d + h * TC_HD;
#pragma unroll
          for (int e = 0; e < 64; e += 8) st_global_256(dst + e, o + e);
        }
        if (a.o16) {
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) split_pack(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1]), hi[e], lo[e]);
          __half* dst = a.o16 + row * size_t(a.o_splits) * d + h * TC_HD;
#pragma unroll
          for (int e = 0; e < 32; e += 8) st_global_256(dst + 2 * e, hi + e);
          if (a.o_splits == 2) {
#pragma unroll
            for (int e = 0; e < 32; e += 8) st_global_256(dst + d + 2 * e, lo + e);
          }
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

// ---------------------------------------------------------------------------------------------------------
// backward + relevance (ClipGradcam.interpret, CLIP/clip/clip_gradcam.py:90-126; the autograd graph of
// auxiliary.multi_head_attention_forward), one CTA per (label, sequence, head):
//   G = dO V^T ; delta_i = dO_i . O_i ; dS = A o (G - delta) ; w_j = (1/H) sum_i r_i relu?(G o A)_ij
//   dQ = scale dS K ; dK = dS^T Q ; dV = A^T dO
// Row pass (thread = query row i):  MMA G[128 x 272] -> dS packed in place -> TS MMA dQ = dS K (K as MN-major B).
// Column pass (thread = key j):     MMA G^T = V dO^T -> relevance column sums are thread-local -> dS^T and A^T packed
//                                   in place -> TS MMAs dK = dS^T Q, dV = A^T dO (Q, dO as MN-major B).
// ---------------------------------------------------------------------------------------------------------
struct AttnBwdTcArgs {
  const __half* probs16;  // [B*H, T, ldp]
  int ldp;
  const float* o32;       // [B*T, d]
  const __half* dO16;     // [P*B*T, ld_do]
  int ld_do;
  float* delta;           // [P*B*H, T]
  const float* r;         // [P*B, T]
  float* wpart;           // [P*B*H, T]
  __half* dqkv16;         // [P*B*T, splits*3d]
  int P, B, T, H, d, splits;
  float scale;
  int positive_only, need_dqkv;
};

struct RowSmem {
  static constexpr int DO = 0;
  static constexpr int V = DO + TC_BOX_BYTES;
  static constexpr int K = V + TC_KV_BYTES;
  static constexpr int BARS = K + TC_KV_BYTES;
  static constexpr int TOTAL = BARS + 128 + 1024;
};

__device__ __forceinline__ void store_row64_f16(__half* dst, int lo_off, int splits, const uint32_t (&o)[64], float scale) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) split_pack(__uint_as_float(o[2 * e]) * scale, __uint_as_float(o[2 * e + 1]) * scale, hi[e], lo[e]);
#pragma unroll
  for (int e = 0; e < 32; e += 8) st_global_256(dst + 2 * e, hi + e);
  if (splits == 2) {
#pragma unroll
    for (int e = 0; e < 32; e += 8) st_global_256(dst + lo_off + 2 * e, lo + e);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_row_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RowSmem::BARS);
  uint64_t *bar_kv = bars, *bar_q = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4, *bar_free = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x % a.P, bh = blockIdx.x / a.P, b = bh / a.H, h = bh % a.H;
  const int pb = p * a.B + b;
  const int T = a.T, d = a.d;
  const int n_mt = (T + 127) / 128;
  const int ncol = (T + 15) & ~15;
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_kv, 1), mbar_init(bar_q, 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, 128), mbar_init(bar_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    if (leader) {
      mbar_arrive_expect_tx(bar_kv, 2u * TC_KV_BYTES);
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_2d(smem + RowSmem::V + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, 2 * d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
        tma_load_2d(smem + RowSmem::K + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
      }
    }
    const uint32_t sbase = smem_u32(smem);
    const uint64_t dDO = desc_kmajor(sbase + RowSmem::DO), dV = desc_kmajor(sbase + RowSmem::V);
    const uint64_t dK = desc_mnmajor(sbase + RowSmem::K, 16);
    const uint32_t idesc_s1 = make_idesc_f16(128, n1), idesc_s2 = make_idesc_f16(128, n2 ? n2 : 16);
    constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);
    const uint32_t tS = tmem_base + TC_COL_S, tO = tmem_base + TC_COL_O;
    for (int mt = 0; mt < n_mt; ++mt) {
      if (leader) {
        mbar_arrive_expect_tx(bar_q, TC_BOX_BYTES);
        tma_load_2d(smem + RowSmem::DO, &tm_do, bar_q, h * TC_HD, pb * T + mt * 128);
      }
      mbar_wait(bar_q, mt & 1);
      if (mt == 0) mbar_wait(bar_kv, 0);
      if (mt > 0) mbar_wait(bar_free, (mt - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < TC_HD / 16; ++k) {
        umma_f16_elect(tS, dDO + uint64_t(2 * k), dV + uint64_t(2 * k), idesc_s1, k != 0, leader);
        if (n2) umma_f16_elect(tS + 256, dDO + uint64_t(2 * k), dV + uint64_t(2 * k) + uint64_t((256 * 128) >> 4), idesc_s2, k != 0, leader);
      }
      umma_commit_elect(bar_s, leader);
      mbar_wait(bar_p, mt & 1);
      tc_fence_after();
      if (a.need_dqkv) {
        for (int s = 0; s < ncol / 16; ++s)
          umma_f16_ts_elect(tO, tS + uint32_t(16 * s), dK + uint64_t(s) * (2048 >> 4), idesc_o, s > 0, leader);
      }
      umma_commit_elect(bar_o, leader);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    for (int mt = 0; mt < n_mt; ++mt) {
      const int i = mt * 128 + q * 32 + lane;
      const bool valid = i < T;
      // probability row and delta_i = dO_i . O_i, fetched while the G MMA runs
      uint32_t arow[TC_MAX_T / 16][8];
      const __half* prow = a.probs16 + (size_t(bh) * T + (valid ? i : 0)) * a.ldp;
#pragma unroll
      for (int c = 0; c < TC_MAX_T / 16; ++c) {
        if (c * 16 < ncol) {
          if (valid) {
            ld_global_256(prow + c * 16, arow[c]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) arow[c][e] = 0u;
          }
        }
      }
      float delta = 0.f;
      if (valid) {
        const float* orow = a.o32 + (size_t(b) * T + i) * d + h * TC_HD;
        const __half* grow = a.dO16 + (size_t(pb) * T + i) * a.ld_do + h * TC_HD;
#pragma unroll
        for (int e = 0; e < 64; e += 16) {
          uint32_t ov[16], gv[8];
          ld_global_256(orow + e, ov);
          ld_global_256(orow + e + 8, ov + 8);
          ld_global_256(grow + e, gv);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&gv[k]));
            delta = fmaf(g2.x, __uint_as_float(ov[2 * k]), delta);
            delta = fmaf(g2.y, __uint_as_float(ov[2 * k + 1]), delta);
          }
        }
        a.delta[(size_t(pb) * a.H + h) * T + i] = delta;
      }
      mbar_wait(bar_s, mt & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < TC_MAX_T / 16; ++c) {
        if (c * 16 < ncol) {
          uint32_t g[16], w[8];
          tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c * 16), g);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&arow[c][e]));
            w[e] = pack_h2(a2.x * (__uint_as_float(g[2 * e]) - delta), a2.y * (__uint_as_float(g[2 * e + 1]) - delta));
          }
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c * 16), w);
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, mt & 1);
      tc_fence_after();
      if (a.need_dqkv) {
        uint32_t o[64];
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O), *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O + 32), *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_free);
        if (valid) store_row64_f16(a.dqkv16 + (size_t(pb) * T + i) * size_t(a.splits) * 3 * d + h * TC_HD, 3 * d, a.splits, o, a.scale);
      } else {
        tc_fence_before();
        mbar_arrive(bar_free);
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

struct ColSmem {
  static constexpr int V = 0;                          // key tile (A operand of G^T)
  static constexpr int DO = V + TC_BOX_BYTES;          // all query rows: B of G^T (K-major) and of dV (MN-major)
  static constexpr int Q = DO + TC_KV_BYTES;           // all query rows: B of dK (MN-major)
  static constexpr int PR = Q + TC_KV_BYTES;           // probabilities [272 i][128 j]: 2 column blocks x 2 row boxes
  static constexpr int DR = PR + 4 * TC_BOX_BYTES;     // float2 {delta_i, r_i}
  static constexpr int BARS = DR + TC_MAX_T * 8;
  static constexpr int TOTAL = BARS + 128 + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_col_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_pr, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ColSmem::BARS);
  uint64_t *bar_kv = bars, *bar_q = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4, *bar_free = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float2* s_dr = reinterpret_cast<float2*>(smem + ColSmem::DR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x % a.P, bh = blockIdx.x / a.P, b = bh / a.H, h = bh % a.H;
  const int pb = p * a.B + b;
  const int T = a.T, d = a.d;
  const int n_mt = (T + 127) / 128;
  const int ncol = (T + 15) & ~15;  // query columns of G^T
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_pr);
    mbar_init(bar_kv, 1), mbar_init(bar_q, 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, 128), mbar_init(bar_free, 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < TC_MAX_T; i += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (i < T) {
      v.x = a.need_dqkv ? a.delta[(size_t(pb) * a.H + h) * T + i] : 0.f;
      v.y = a.r[size_t(pb) * T + i];
    }
    s_dr[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    if (leader) {
      mbar_arrive_expect_tx(bar_kv, 2u * TC_KV_BYTES);
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_2d(smem + ColSmem::DO + bx * TC_BOX_BYTES, &tm_do, bar_kv, h * TC_HD, pb * T + bx * TC_BOX_ROWS);
        tma_load_2d(smem + ColSmem::Q + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, h * TC_HD, b * T + bx * TC_BOX_ROWS);
      }
    }
    const uint32_t sbase = smem_u32(smem);
    const uint64_t dVt = desc_kmajor(sbase + ColSmem::V), dDOk = desc_kmajor(sbase + ColSmem::DO);
    const uint64_t dDOm = desc_mnmajor(sbase + ColSmem::DO, 16), dQm = desc_mnmajor(sbase + ColSmem::Q, 16);
    const uint32_t idesc_s1 = make_idesc_f16(128, n1), idesc_s2 = make_idesc_f16(128, n2 ? n2 : 16);
    constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);
    const uint32_t tS = tmem_base + TC_COL_S, tK = tmem_base + TC_COL_O, tV = tmem_base + TC_COL_O2;
    for (int mt = 0; mt < n_mt; ++mt) {
      if (mt > 0) mbar_wait(bar_free, (mt - 1) & 1);  // previous tile's threads are done with the probability tile
      if (leader) {
        mbar_arrive_expect_tx(bar_q, 5u * TC_BOX_BYTES);
        tma_load_2d(smem + ColSmem::V, &tm_qkv, bar_q, 2 * d + h * TC_HD, b * T + mt * 128);
        for (int cb = 0; cb < 2; ++cb)
          for (int bx = 0; bx < 2; ++bx)
            tma_load_2d(smem + ColSmem::PR + (cb * 2 + bx) * TC_BOX_BYTES, &tm_pr, bar_q, mt * 128 + cb * 64, bh * T + bx * TC_BOX_ROWS);
      }
      mbar_wait(bar_q, mt & 1);
      if (mt == 0) mbar_wait(bar_kv, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < TC_HD / 16; ++k) {
        umma_f16_elect(tS, dVt + uint64_t(2 * k), dDOk + uint64_t(2 * k), idesc_s1, k != 0, leader);
        if (n2) umma_f16_elect(tS + 256, dVt + uint64_t(2 * k), dDOk + uint64_t(2 * k) + uint64_t((256 * 128) >> 4), idesc_s2, k != 0, leader);
      }
      umma_commit_elect(bar_s, leader);
      mbar_wait(bar_p, mt & 1);
      tc_fence_after();
      if (a.need_dqkv) {
        for (int s = 0; s < ncol / 16; ++s) {
          const uint64_t kadv = uint64_t(s) * (2048 >> 4);
          umma_f16_ts_elect(tK, tS + uint32_t(16 * s), dQm + kadv, idesc_o, s > 0, leader);
          umma_f16_ts_elect(tV, tS + uint32_t(16 * s + 8), dDOm + kadv, idesc_o, s > 0, leader);
        }
      }
      umma_commit_elect(bar_o, leader);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    const int jj = q * 32 + lane;  // key inside the tile
    // probability tile: column block jj/64, 16-byte chunk (jj%64)/8 (xor-swizzled with the row), element jj%8
    const uint8_t* ptile = smem + ColSmem::PR + (jj >> 6) * (2 * TC_BOX_BYTES) + (jj & 7) * 2;
    const int chunk = (jj & 63) >> 3;
    const float invH = 1.0f / a.H;
    for (int mt = 0; mt < n_mt; ++mt) {
      const int j = mt * 128 + jj;
      const bool valid = j < T;
      mbar_wait(bar_q, mt & 1);  // probability tile landed
      mbar_wait(bar_s, mt & 1);
      tc_fence_after();
      float w = 0.f;
      for (int c = 0; c < ncol; c += 16) {
        uint32_t g[16], ds[8], aw[8];
        tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c), g);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float av[2], dv[2];
          uint32_t araw[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = c + e + u;
            const int bx = i >= TC_BOX_ROWS, ii = i - bx * TC_BOX_ROWS;
            const uint16_t raw = (i < T) ? *reinterpret_cast<const uint16_t*>(ptile + bx * TC_BOX_BYTES + ii * 128 + ((chunk ^ (ii & 7)) << 4)) : uint16_t(0);
            araw[u] = raw;
            av[u] = __half2float(__ushort_as_half(raw));
            const float2 dr = s_dr[i];
            const float gv = __uint_as_float(g[e + u]);
            float x = gv * av[u];
            if (a.positive_only) x = fmaxf(x, 0.f);
            w = fmaf(dr.y, x, w);
            dv[u] = av[u] * (gv - dr.x);
          }
          ds[e >> 1] = pack_h2(dv[0], dv[1]);
          aw[e >> 1] = araw[0] | (araw[1] << 16);
        }
        tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c), ds);
        tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c + 8), aw);
      }
      if (valid) a.wpart[(size_t(pb) * a.H + h) * T + j] = w * invH;
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, mt & 1);
      tc_fence_after();
      if (a.need_dqkv) {
        uint32_t o[64];
        __half* drow = a.dqkv16 + (size_t(pb) * T + (valid ? j : 0)) * size_t(a.splits) * 3 * d + h * TC_HD;
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O), *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O + 32), *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tc_wait_ld();
        if (valid) store_row64_f16(drow + d, 3 * d, a.splits, o, 1.0f);
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O2), *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O2 + 32), *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tc_wait_ld();
        if (valid) store_row64_f16(drow + 2 * d, 3 * d, a.splits, o, 1.0f);
      }
      tc_fence_before();
      mbar_arrive(bar_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_tile_tmap(CUtensorMap* tm, const void* base, long long rows, long long cols, long long pitch) {
  uint64_t dims[2] = {uint64_t(cols), uint64_t(rows)};
  uint64_t str[1] = {uint64_t(pitch) * 2};
  uint32_t box[2] = {TC_HD, TC_BOX_ROWS};
  return make_tmap_f16(tm, base, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int make_qkv_tmap(CUtensorMap* tm, const void* base, long long rows, long long cols) {
  uint64_t dims[2] = {uint64_t(cols), uint64_t(rows)};
  uint64_t str[1] = {uint64_t(cols) * 2};
  uint32_t box[2] = {TC_HD, TC_BOX_ROWS};
  return make_tmap_f16(tm, base, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_selftest_ts_mma(const void* A16, const void* B16, float* D, int32_t Kd, int32_t lbo, int32_t sbo,
                                      void* stream) {
  SB_REQUIRE(A16 && B16 && D && Kd > 0 && Kd % 16 == 0 && Kd <= 256, "semabs_selftest_ts_mma: bad arguments");
  const int smem = 256 * 128 + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(ts_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  ts_selftest_kernel<<<1, TC_THREADS, smem, (cudaStream_t)stream>>>((const __half*)A16, (const __half*)B16, D, Kd, lbo, sbo);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_attn_fwd_tc(const void* qkv16, int32_t in_splits, void* probs16, int32_t ld_p16, float* o32, void* o16,
                                  int32_t o_splits, int32_t B, int32_t T, int32_t H, int32_t causal, void* stream) {
  SB_REQUIRE(qkv16 && (o32 || o16) && B > 0 && T > 0 && H > 0, "semabs_attn_fwd_tc: bad arguments");
  SB_REQUIRE(T <= TC_MAX_T, "semabs_attn_fwd_tc: T=%d exceeds the %d-key strip", T, TC_MAX_T);
  SB_REQUIRE(in_splits == 1 || in_splits == 2, "semabs_attn_fwd_tc: in_splits must be 1 or 2");
  SB_REQUIRE(!probs16 || (ld_p16 >= T && ld_p16 % 16 == 0), "semabs_attn_fwd_tc: bad probs16 pitch %d", ld_p16);
  const int d = H * TC_HD;
  CUtensorMap tm;
  if (int rc = make_qkv_tmap(&tm, qkv16, (long long)B * T, (long long)in_splits * 3 * d)) return rc;
  AttnFwdTcArgs a{(__half*)probs16, ld_p16, o32, (__half*)o16, o_splits, T, H, d, causal, in_splits};
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::TOTAL));
    configured = true;
  }
  attn_fwd_tc_kernel<<<B * H, TC_THREADS, TcSmem::TOTAL, (cudaStream_t)stream>>>(tm, a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_attn_bwd_tc(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                                  const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart,
                                  void* dqkv16, int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits,
                                  int32_t positive_only, int32_t need_dqkv, void* stream) {
  SB_REQUIRE(qkv16 && probs16 && o32 && dO16 && r && delta_ws && wpart, "semabs_attn_bwd_tc: null pointer");
  SB_REQUIRE(!need_dqkv || dqkv16, "semabs_attn_bwd_tc: dqkv16 missing");
  SB_REQUIRE(P > 0 && B > 0 && T > 0 && H > 0 && T <= TC_MAX_T, "semabs_attn_bwd_tc: bad shape (T <= %d)", TC_MAX_T);
  SB_REQUIRE(ld_p16 % 16 == 0 && ld_p16 >= ((T + 15) / 16) * 16, "semabs_attn_bwd_tc: bad probs16 pitch %d", ld_p16);
  const int d = H * TC_HD;
  SB_REQUIRE(ld_qkv >= 3 * d && ld_qkv % 8 == 0 && ld_do >= d && ld_do % 16 == 0, "semabs_attn_bwd_tc: bad pitch");
  SB_REQUIRE(splits == 1 || splits == 2, "semabs_attn_bwd_tc: splits must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tm_qkv, tm_do, tm_pr;
  if (int rc = make_tile_tmap(&tm_qkv, qkv16, (long long)B * T, 3LL * d, ld_qkv)) return rc;
  if (int rc = make_tile_tmap(&tm_do, dO16, (long long)P * B * T, d, ld_do)) return rc;
  if (int rc = make_tile_tmap(&tm_pr, probs16, (long long)B * H * T, ld_p16, ld_p16)) return rc;
  AttnBwdTcArgs a{};
  a.probs16 = (const __half*)probs16, a.ldp = ld_p16, a.o32 = o32, a.dO16 = (const __half*)dO16, a.ld_do = ld_do;
  a.delta = delta_ws, a.r = r, a.wpart = wpart, a.dqkv16 = (__half*)dqkv16;
  a.P = P, a.B = B, a.T = T, a.H = H, a.d = d, a.splits = splits, a.scale = 0.125f;
  a.positive_only = positive_only, a.need_dqkv = need_dqkv;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_row_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RowSmem::TOTAL));
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_col_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ColSmem::TOTAL));
    configured = true;
  }
  const int grid = P * B * H;
  if (need_dqkv) {  // the row pass produces dQ and delta; the relevance-only last step needs neither
    attn_bwd_row_tc_kernel<<<grid, TC_THREADS, RowSmem::TOTAL, st>>>(tm_qkv, tm_do, a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  attn_bwd_col_tc_kernel<<<grid, TC_THREADS, ColSmem::TOTAL, st>>>(tm_qkv, tm_do, tm_pr, a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
