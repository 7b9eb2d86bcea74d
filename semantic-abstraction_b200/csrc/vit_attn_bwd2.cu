// Attention backward + relevance on tcgen05, second generation (same math and TMEM / shared-memory layouts as the row / column
// passes of vit_attn_tc.cu — see the derivation there; reference: ClipGradcam.interpret, CLIP/clip/clip_gradcam.py:90-126 over
// the autograd graph of auxiliary.multi_head_attention_forward, CLIP/clip/auxiliary.py:306-345).
//
// What the first generation measured (profiles/r02_attn_bwd_stalls.txt, ncu source-level stall sampling at the bench shape
// 95 tiles x 16 labels): tensor pipe 13 % busy; per (unit, label) item ~9.4 k cycles in the row pass of which 31 % was the
// delta_i = dO_i . O_i phase (dependent global loads + a CTA barrier), 22 % waiting for MMAs that the SAME warp had to issue
// after finishing its own element-wise share, 30 % the latency-bound TMEM ld -> compute -> st chain; the column pass spent
// 55 % of its issue slots in the element-wise loop (16 two-byte shared loads + bounds selects per 16 elements, repeated for
// every one of the 16 labels although the probability tile is the same), and the one-row / one-key SIMT tail cost 19 % of
// the attention backward on its own.  Changes:
//   * delta is computed once by a bandwidth kernel (one thread per (row, head): 128 B of dO, 256 B of O) — no per-item
//     global-latency phase, no CTA barrier in the passes, and the column pass / tail read the same array;
//   * a dedicated control warp (warp 8) issues every TMA and MMA: the tensor pipe no longer waits for warp 0's SIMT share,
//     and the next item's G MMA starts the moment the previous product has drained;
//   * row pass: the element-wise loop loads two 16-column chunks per tcgen05.wait::ld (half the exposed TMEM latency), the
//     per-item scalar (delta) is prefetched one item ahead;
//   * column pass: each thread keeps its key's probability column (A^T, 136 values = 68 registers) across the 16 labels of a
//     unit, like the row pass keeps its row: the 272 LDS.U16 + selects per item are paid once per unit;
//     {delta_i, r_i} of the next item are fetched into registers before the element-wise loop and stored to shared after it;
//   * tail: one CTA per (sequence, head) stages Q / K / V once and walks the labels with dO staged per label; every phase
//     is a block-wide pass over padded (conflict-free) shared rows instead of one latency-bound warp per (label, head).
#include "vit_attn_tc.cuh"

namespace sb {

constexpr int TC_BWD_THREADS = 288;  // 8 element-wise warps (2 per TMEM lane quadrant) + 1 control warp
constexpr int TC_SIMT = 256;
constexpr int NCH_HALF = (TC_MAX_T / 16 + 1) / 2;  // 9 chunks of 16 strip columns per thread at most

// ---------------------------------------------------------------------------------------------------------
// delta[pb, h, i] = sum_c dO[pb, i, h*64 + c] * O[b, i, h*64 + c]
// ---------------------------------------------------------------------------------------------------------
// block = 16 consecutive (b, i) rows x H heads, looping over the P labels: a thread keeps its 64 O values in registers and
// reads 128 B of dO per label (whole contiguous rows per warp); each label's results are transposed through shared memory so
// that a head's 16 deltas leave as one 64-byte segment.  (First version: one thread per (label, row, head) re-reading O
// for every label and storing 4 bytes at a stride of T floats — 0.36 ms, L2-bandwidth bound on the 16x re-read of O.)
constexpr int DELTA_ROWS = 16;
__global__ void __launch_bounds__(512) attn_delta_kernel(AttnBwdTcArgs a) {
  __shared__ float s[2][DELTA_ROWS][33];
  const int H = a.H, tid = threadIdx.x;
  const int rows_total = a.B * a.T;
  const int r = tid / H, h = tid - r * H;
  const int row = blockIdx.x * DELTA_ROWS + r;  // (b, i)
  const bool ok = row < rows_total;
  const int i = ok ? row % a.T : 0, b = ok ? row / a.T : 0;
  uint32_t ov[64];
  if (ok) {
    const float* o = a.o32 + (size_t(b) * a.T + i) * a.d + h * TC_HD;
#pragma unroll
    for (int e = 0; e < 8; ++e) ld_global_256(o + 8 * e, ov + 8 * e);
  }
  const int h2 = tid / DELTA_ROWS, r2 = tid % DELTA_ROWS;
  const int row2 = blockIdx.x * DELTA_ROWS + r2;
  const int i2 = row2 % a.T, b2 = row2 / a.T;
  uint32_t gv[32];
  auto load_g = [&](int p) {
    const __half* g = a.dO16 + (size_t(p * a.B + b) * a.T + i) * a.ld_do + h * TC_HD;
#pragma unroll
    for (int e = 0; e < 4; ++e) ld_global_256(g + 16 * e, gv + 8 * e);
  };
  if (ok) load_g(0);
  for (int p = 0; p < a.P; ++p) {
    float acc0 = 0.f, acc1 = 0.f;
    if (ok) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&gv[k]));
        acc0 = fmaf(g2.x, __uint_as_float(ov[2 * k]), acc0);
        acc1 = fmaf(g2.y, __uint_as_float(ov[2 * k + 1]), acc1);
      }
      if (p + 1 < a.P) load_g(p + 1);  // next label's row in flight across the transpose
      s[p & 1][r][h] = acc0 + acc1;
    }
    __syncthreads();  // (two buffers: the writes of label p+1 cannot overtake the reads of label p by more than one barrier)
    if (row2 < rows_total) a.delta[(size_t(p * a.B + b2) * H + h2) * a.T + i2] = s[p & 1][r2][h2];
  }
}

// ---------------------------------------------------------------------------------------------------------
// row pass: thread = query row i.  G = dO V^T -> dS = A o (G - delta) packed in place -> dQ = scale dS K
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_BWD_THREADS, 1)
attn_bwd_row_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RowSmem::BARS);
  uint64_t *bar_kv = bars, *bar_do = bars + 1 /* [2] */, *bar_s = bars + 3, *bar_p = bars + 4, *bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, d = a.d, P = a.P;
  const int ncol = (T + 15) & ~15;
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;
  const int nch = ncol / 16, c_split = (nch + 1) / 2;
  const int n_units = a.B * a.H * a.n_full;
  const int n_my = (n_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int N = n_my * P;  // items of this CTA: unit-major, label fastest

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_kv, 1), mbar_init(&bar_do[0], 1), mbar_init(&bar_do[1], 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, TC_SIMT);
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
  const uint32_t tS = tmem_base + TC_COL_S, tO = tmem_base + TC_COL_O;

  auto item = [&](int n, int& bh, int& mt, int& p) {
    const int u = int(blockIdx.x) + (n / P) * int(gridDim.x);
    p = n % P, bh = u / a.n_full, mt = u % a.n_full;
  };

  if (warp == 8) {
    // ===== control warp: TMA + MMA issue (convergent; only the elected lane's instructions take effect) =====
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t sbase = smem_u32(smem);
    const uint64_t dV = desc_kmajor(sbase + RowSmem::V), dK = desc_mnmajor(sbase + RowSmem::K, 16);
    const uint32_t idesc_s1 = make_idesc_f16(128, n1), idesc_s2 = make_idesc_f16(128, n2 ? n2 : 16);
    constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);
    auto load_do = [&](int n) {
      int bh, mt, p;
      item(n, bh, mt, p);
      const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
      mbar_arrive_expect_tx(&bar_do[n & 1], TC_BOX_BYTES);
      tma_load_2d(smem + RowSmem::DO + (n & 1) * TC_BOX_BYTES, &tm_do, &bar_do[n & 1], h * TC_HD, pb * T + mt * 128);
    };
    auto load_kv = [&](int n) {  // n = first item of a unit
      int bh, mt, p;
      item(n, bh, mt, p);
      const int b = bh / a.H, h = bh % a.H;
      mbar_arrive_expect_tx(bar_kv, 2u * TC_KV_BYTES);
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_2d(smem + RowSmem::V + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, 2 * d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
        tma_load_2d(smem + RowSmem::K + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
      }
    };
    auto issue_g = [&](int n) {
      if (n % P == 0) mbar_wait(bar_kv, (n / P) & 1);
      mbar_wait(&bar_do[n & 1], (n >> 1) & 1);
      tc_fence_after();
      const uint64_t dDO = desc_kmajor(sbase + RowSmem::DO + (n & 1) * TC_BOX_BYTES);
#pragma unroll
      for (int k = 0; k < TC_HD / 16; ++k) {
        umma_f16_elect(tS, dDO + uint64_t(2 * k), dV + uint64_t(2 * k), idesc_s1, k != 0, leader);
        if (n2) umma_f16_elect(tS + 256, dDO + uint64_t(2 * k), dV + uint64_t(2 * k) + uint64_t((256 * 128) >> 4), idesc_s2, k != 0, leader);
      }
      umma_commit_elect(bar_s, leader);
    };
    if (N > 0) {
      if (leader) {
        load_kv(0);
        load_do(0);
        if (N > 1) load_do(1);
      }
      __syncwarp();
      issue_g(0);
    }
    for (int n = 0; n < N; ++n) {
      mbar_wait(bar_p, n & 1);  // every row's dS is packed in the strip (and the epilogue of item n-1 has drained dQ)
      tc_fence_after();
      if (leader && n + 2 < N) load_do(n + 2);  // G(n) — long complete — was the last reader of dO stage n&1
      __syncwarp();
      for (int s = 0; s < nch; ++s)
        umma_f16_ts_elect(tO, tS + uint32_t(16 * s), dK + uint64_t(s) * (2048 >> 4), idesc_o, s > 0, leader);
      umma_commit_elect(bar_o, leader);
      if (n + 1 < N) {
        // the strip (A operand of dQ(n)) and, at a unit boundary, K / V are free only once dQ(n) has completed
        mbar_wait(bar_o, n & 1);
        tc_fence_after();
        if ((n + 1) % P == 0 && leader) load_kv(n + 1);
        __syncwarp();
        issue_g(n + 1);  // overlaps the element-wise warps' epilogue of item n
      }
    }
  } else {
    // ===== element-wise warps =====
    const int q = warp & 3, half = warp >> 2, rr = q * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    const int c0 = half ? c_split : 0, c1 = half ? nch : c_split;
    uint32_t arow[NCH_HALF][8];  // this thread's half of its probability row (kept across the P labels of a unit)
    float dnext = 0.f;
    auto delta_of = [&](int n) -> float {
      int bh, mt, p;
      item(n, bh, mt, p);
      const int i = mt * 128 + rr;
      if (i >= T) return 0.f;
      const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
      return a.delta[(size_t(pb) * a.H + h) * T + i];
    };
    if (N > 0) dnext = delta_of(0);
    for (int n = 0; n < N; ++n) {
      int bh, mt, p;
      item(n, bh, mt, p);
      const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
      const int i = mt * 128 + rr;
      const bool valid = i < T;
      if (p == 0) {
        const __half* prow = a.probs16 + (size_t(bh) * T + (valid ? i : 0)) * a.ldp;
#pragma unroll
        for (int cc = 0; cc < NCH_HALF; ++cc) {
          if (c0 + cc < c1) {
            if (valid) {
              ld_global_256(prow + (c0 + cc) * 16, arow[cc]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) arow[cc][e] = 0u;
            }
          }
        }
      }
      const float delta = dnext;
      if (n + 1 < N) dnext = delta_of(n + 1);  // in flight during this item's element-wise phase

      mbar_wait(bar_s, n & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < NCH_HALF; cc += 2) {
        const bool v0 = c0 + cc < c1, v1 = (cc + 1 < NCH_HALF) && (c0 + cc + 1 < c1);  // warp-uniform
        uint32_t g0[16], g1[16], w[8];
        if (v0) tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + (c0 + cc) * 16), g0);
        if (v1) tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + (c0 + cc + 1) * 16), g1);
        tc_wait_ld();
        if (v0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&arow[cc][e]));
            w[e] = pack_h2(a2.x * (__uint_as_float(g0[2 * e]) - delta), a2.y * (__uint_as_float(g0[2 * e + 1]) - delta));
          }
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + (c0 + cc) * 16), w);
        }
        if (v1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&arow[(cc + 1) % NCH_HALF][e]));
            w[e] = pack_h2(a2.x * (__uint_as_float(g1[2 * e]) - delta), a2.y * (__uint_as_float(g1[2 * e + 1]) - delta));
          }
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + (c0 + cc + 1) * 16), w);
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, n & 1);
      tc_fence_after();
      {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O + 32 * half), o);
        tc_wait_ld();
        if (valid)
          store_row_f16(a.dqkv16 + (size_t(pb) * T + i) * size_t(a.splits) * 3 * d + h * TC_HD + 32 * half, 3 * d, a.splits, o, 32, a.scale);
      }
      tc_fence_before();
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
// column pass: thread = key j.  G^T = V dO^T -> relevance column sums, dS^T and A^T packed in place -> dK = dS^T Q, dV = A^T dO
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_BWD_THREADS, 1)
attn_bwd_col_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_pr, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ColSmem::BARS);
  uint64_t *bar_kv = bars, *bar_do = bars + 1 /* [2] */, *bar_s = bars + 3, *bar_p = bars + 4, *bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float2* s_dr = reinterpret_cast<float2*>(smem + ColSmem::DR);
  float* s_w = reinterpret_cast<float*>(smem + ColSmem::W);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, d = a.d, P = a.P;
  const int ncol = (T + 15) & ~15;  // query columns of G^T
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;
  const int nch = ncol / 16, c_split = (nch + 1) / 2;
  const int n_units = a.B * a.H * a.n_full;
  const int n_my = (n_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int N = n_my * P;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_pr);
    mbar_init(bar_kv, 1), mbar_init(&bar_do[0], 1), mbar_init(&bar_do[1], 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, TC_SIMT);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  auto item = [&](int n, int& bh, int& mt, int& p) {
    const int u = int(blockIdx.x) + (n / P) * int(gridDim.x);
    p = n % P, bh = u / a.n_full, mt = u % a.n_full;
  };
  // {delta_i, r_i} of item n for the (at most two) query rows this thread stages: i = tid and tid + 256
  auto fetch_dr = [&](int n, float2& v0, float2& v1) {
    int bh, mt, p;
    item(n, bh, mt, p);
    const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
    const int i0 = int(threadIdx.x), i1 = int(threadIdx.x) + TC_SIMT;
    v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
    if (i0 < T) {
      v0.x = a.need_dqkv ? a.delta[(size_t(pb) * a.H + h) * T + i0] : 0.f;
      v0.y = a.r[size_t(pb) * T + i0];
    }
    if (i1 < T) {
      v1.x = a.need_dqkv ? a.delta[(size_t(pb) * a.H + h) * T + i1] : 0.f;
      v1.y = a.r[size_t(pb) * T + i1];
    }
  };
  auto store_dr = [&](int n, const float2& v0, const float2& v1) {
    s_dr[(n & 1) * TC_MAX_T + int(threadIdx.x)] = v0;
    if (int(threadIdx.x) + TC_SIMT < TC_MAX_T) s_dr[(n & 1) * TC_MAX_T + int(threadIdx.x) + TC_SIMT] = v1;
  };
  if (warp < 8 && N > 0) {
    float2 v0, v1;
    fetch_dr(0, v0, v1);
    store_dr(0, v0, v1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base + TC_COL_S, tK = tmem_base + TC_COL_O, tV = tmem_base + TC_COL_O2;

  if (warp == 8) {
    // ===== control warp =====
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t sbase = smem_u32(smem);
    const uint64_t dVt = desc_kmajor(sbase + ColSmem::V), dQm = desc_mnmajor(sbase + ColSmem::Q, 16);
    const uint32_t idesc_s1 = make_idesc_f16(128, n1), idesc_s2 = make_idesc_f16(128, n2 ? n2 : 16);
    constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);
    auto load_do = [&](int n) {
      int bh, mt, p;
      item(n, bh, mt, p);
      const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
      mbar_arrive_expect_tx(&bar_do[n & 1], TC_KV_BYTES);
      for (int bx = 0; bx < 2; ++bx)
        tma_load_2d(smem + ColSmem::DO + (n & 1) * TC_KV_BYTES + bx * TC_BOX_BYTES, &tm_do, &bar_do[n & 1], h * TC_HD,
                    pb * T + bx * TC_BOX_ROWS);
    };
    auto load_unit = [&](int n) {
      int bh, mt, p;
      item(n, bh, mt, p);
      const int b = bh / a.H, h = bh % a.H;
      mbar_arrive_expect_tx(bar_kv, 7u * TC_BOX_BYTES);
      for (int bx = 0; bx < 2; ++bx)
        tma_load_2d(smem + ColSmem::Q + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, h * TC_HD, b * T + bx * TC_BOX_ROWS);
      tma_load_2d(smem + ColSmem::V, &tm_qkv, bar_kv, 2 * d + h * TC_HD, b * T + mt * 128);
      for (int cb = 0; cb < 2; ++cb)
        for (int bx = 0; bx < 2; ++bx)
          tma_load_2d(smem + ColSmem::PR + (cb * 2 + bx) * TC_BOX_BYTES, &tm_pr, bar_kv, mt * 128 + cb * 64, bh * T + bx * TC_BOX_ROWS);
    };
    auto issue_g = [&](int n) {
      if (n % P == 0) mbar_wait(bar_kv, (n / P) & 1);
      mbar_wait(&bar_do[n & 1], (n >> 1) & 1);
      tc_fence_after();
      const uint64_t dDOk = desc_kmajor(sbase + ColSmem::DO + (n & 1) * TC_KV_BYTES);
#pragma unroll
      for (int k = 0; k < TC_HD / 16; ++k) {
        umma_f16_elect(tS, dVt + uint64_t(2 * k), dDOk + uint64_t(2 * k), idesc_s1, k != 0, leader);
        if (n2) umma_f16_elect(tS + 256, dVt + uint64_t(2 * k), dDOk + uint64_t(2 * k) + uint64_t((256 * 128) >> 4), idesc_s2, k != 0, leader);
      }
      umma_commit_elect(bar_s, leader);
    };
    if (N > 0) {
      if (leader) {
        load_unit(0);
        load_do(0);
        if (N > 1) load_do(1);
      }
      __syncwarp();
      issue_g(0);
    }
    for (int n = 0; n < N; ++n) {
      mbar_wait(bar_p, n & 1);
      tc_fence_after();
      if (a.need_dqkv) {
        const uint64_t dDOm = desc_mnmajor(sbase + ColSmem::DO + (n & 1) * TC_KV_BYTES, 16);
        for (int s = 0; s < nch; ++s) {
          const uint64_t kadv = uint64_t(s) * (2048 >> 4);
          umma_f16_ts_elect(tK, tS + uint32_t(16 * s), dQm + kadv, idesc_o, s > 0, leader);
          umma_f16_ts_elect(tV, tS + uint32_t(16 * s + 8), dDOm + kadv, idesc_o, s > 0, leader);
        }
      }
      umma_commit_elect(bar_o, leader);
      mbar_wait(bar_o, n & 1);
      tc_fence_after();
      if (leader && n + 2 < N) load_do(n + 2);  // stage n&1 was last read by dV(n)
      if (n + 1 < N) {
        // unit boundary: Q / V / the probability tile are free (dK(n) done; every thread copied its probability column
        // into registers before arriving at bar_p of the unit's first item)
        if ((n + 1) % P == 0 && leader) load_unit(n + 1);
        __syncwarp();
        issue_g(n + 1);
      }
    }
  } else {
    // ===== element-wise warps =====
    const int q = warp & 3, half = warp >> 2, jj = q * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    const int c0 = half ? c_split : 0, c1 = half ? nch : c_split;
    // probability tile addressing: element (i, jj) sits at  i*128 + (((jj%64)/8 ^ (i%8)) << 4) + (jj%8)*2  inside column
    // block jj/64 (the two 136-row TMA boxes of a column block are contiguous and 136 % 8 == 0)
    const uint8_t* pcol = smem + ColSmem::PR + (jj >> 6) * (2 * TC_BOX_BYTES) + (jj & 7) * 2;
    const float invH = 1.0f / a.H;
    uint32_t acol[NCH_HALF][8];  // A[i, j] for this thread's key j and its half of the query rows i, packed pairs
    for (int n = 0; n < N; ++n) {
      int bh, mt, p;
      item(n, bh, mt, p);
      const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
      const int j = mt * 128 + jj;
      const bool valid = j < T;
      float2 nx0, nx1;
      if (n + 1 < N) fetch_dr(n + 1, nx0, nx1);  // global loads in flight during the element-wise phase
      if (p == 0) {
        mbar_wait(bar_kv, (n / P) & 1);  // the tile is read with ordinary loads: every thread acquires the TMA writes
#pragma unroll
        for (int cc = 0; cc < NCH_HALF; ++cc) {
          if (c0 + cc < c1) {
            const uint8_t* prow = pcol + (c0 + cc) * 2048;
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int k0 = e & 7, k1 = (e + 1) & 7;
              uint32_t lo = *reinterpret_cast<const uint16_t*>(prow + (e >> 3) * 1024 + k0 * 128 + ((((jj & 63) >> 3) ^ k0) << 4));
              uint32_t hi = *reinterpret_cast<const uint16_t*>(prow + ((e + 1) >> 3) * 1024 + k1 * 128 + ((((jj & 63) >> 3) ^ k1) << 4));
              const int i0 = (c0 + cc) * 16 + e;
              if (i0 >= T) lo = 0u;      // rows past T belong to the next head
              if (i0 + 1 >= T) hi = 0u;
              acol[cc][e >> 1] = lo | (hi << 16);
            }
          }
        }
      }
      mbar_wait(bar_s, n & 1);
      tc_fence_after();
      const float2* dr = s_dr + (n & 1) * TC_MAX_T;
      float w = 0.f;
#pragma unroll
      for (int cc = 0; cc < NCH_HALF; ++cc) {
        if (c0 + cc < c1) {
          const int c = c0 + cc;
          uint32_t g[16], ds[8];
          tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c * 16), g);
          const float4* dr4 = reinterpret_cast<const float4*>(dr + c * 16);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float2 av = __half22float2(*reinterpret_cast<const __half2*>(&acol[cc][e >> 1]));
            const float4 d4 = dr4[e >> 1];  // {delta_i, r_i, delta_i+1, r_i+1}
            const float g0 = __uint_as_float(g[e]), g1 = __uint_as_float(g[e + 1]);
            float x0 = g0 * av.x, x1 = g1 * av.y;
            if (a.positive_only) x0 = fmaxf(x0, 0.f), x1 = fmaxf(x1, 0.f);
            w = fmaf(d4.y, x0, w);
            w = fmaf(d4.w, x1, w);
            ds[e >> 1] = pack_h2(av.x * (g0 - d4.x), av.y * (g1 - d4.z));
          }
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c * 16), ds);
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c * 16 + 8), acol[cc]);
        }
      }
      s_w[(n & 1) * 256 + half * 128 + jj] = w;
      if (n + 1 < N) store_dr(n + 1, nx0, nx1);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, n & 1);
      mbar_wait(bar_p, n & 1);  // (complete by now) acquire the other threads' s_w / s_dr stores
      tc_fence_after();
      if (half == 0 && valid)
        a.wpart[(size_t(pb) * a.H + h) * T + j] = (s_w[(n & 1) * 256 + jj] + s_w[(n & 1) * 256 + 128 + jj]) * invH;
      if (a.need_dqkv) {
        // 64 accumulator columns in two 32-column loads (the thread's 68 probability registers stay live across items)
        const uint32_t col = half ? TC_COL_O2 : TC_COL_O;
        __half* orow = a.dqkv16 + (size_t(pb) * T + (valid ? j : 0)) * size_t(a.splits) * 3 * d + (half ? 2 * d : d) + h * TC_HD;
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(t_row + col + 32 * part, o);
          tc_wait_ld();
          if (valid) store_row_f16(orow + 32 * part, 3 * d, a.splits, o, 32, 1.0f);
        }
      }
      tc_fence_before();
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
// tail: the single query row / key x0 = 128 * n_full past the last full tile (ViT-L/14: row / key 256 of 257).
// One CTA per (sequence, head); Q, K, V staged once (padded rows: 16-byte row reads by 32 lanes are conflict-free at a
// 144-byte pitch), dO staged per label.  Per label:
//   column part  g_i = dO_i . V_x0 (thread = query i) -> relevance sum, ds_i = A[i,x0] (g_i - delta_i)
//                dK_x0 = sum_i ds_i Q_i, dV_x0 = sum_i A[i,x0] dO_i           (thread = channel pair x row segment)
//   row part     g_j = dO_x0 . V_j (thread = key j) -> ds_j = A[x0,j] (g_j - delta_x0);  dQ_x0 = scale sum_j ds_j K_j
// ---------------------------------------------------------------------------------------------------------
constexpr int TAIL_PITCH = 72;                       // halfs per staged row (64 + 8 pad = 144 B)
constexpr int TAIL_ROWS = 264;                       // >= T (<= 257 + ...) rounded up
constexpr int TAIL_TILE_BYTES = TAIL_ROWS * TAIL_PITCH * 2;
// Two groups of 8 warps per CTA, each on its own labels (p = group, group + 2, ...) with its own dO tile, scratch and named
// barrier: the per-label work is a chain of five short barrier-separated phases, latency-bound with two warps per scheduler
// (7.2 k cycles per label for ~2.8 k cycles of instructions); the second group fills the other half of every stall and the
// next label's dO lands while the group finishes the phases that no longer read the tile.
constexpr int TAIL_GROUPS = 2;
constexpr int TAIL_THREADS = TAIL_GROUPS * TC_SIMT;
struct TailSmem {
  static constexpr int Q = 0, K = Q + TAIL_TILE_BYTES, V = K + TAIL_TILE_BYTES, DO = V + TAIL_TILE_BYTES;  // DO: one tile per group
  static constexpr int DS = DO + TAIL_GROUPS * TAIL_TILE_BYTES;  // per group float [2][TAIL_ROWS]: ds and a (column part) / ds (row part)
  static constexpr int RED = DS + TAIL_GROUPS * 2 * TAIL_ROWS * 4;   // per group float [8 segments][3][64] partial sums
  static constexpr int WRED = RED + TAIL_GROUPS * 8 * 3 * 64 * 4;    // per group float [8] warp partials of the relevance sum
  static constexpr int TOTAL = WRED + TAIL_GROUPS * 32;
};
static_assert(TailSmem::TOTAL + 16 <= 227 * 1024, "tail kernel shared memory");
__device__ __forceinline__ void tail_group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(TC_SIMT) : "memory"); }

// rows [0, T) x 64 halfs -> padded shared rows, 16 bytes per cp.async (no register round trip: every copy of the tile is
// in flight at once; the first version staged through registers and spent 41 % of the kernel waiting on those loads)
__device__ __forceinline__ void tail_stage_async(__half* dst, const __half* src, int ld, int T, int tid, int nthreads = TC_SIMT) {
  for (int idx = tid; idx < T * 8; idx += nthreads) {
    const int r = idx >> 3, c = idx & 7;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + r * TAIL_PITCH + c * 8)),
                 "l"(src + size_t(r) * ld + c * 8)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ float tail_dot64(const __half* x, const __half* y) {  // two staged rows (16-byte aligned)
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 ux = *reinterpret_cast<const uint4*>(x + c * 8), uy = *reinterpret_cast<const uint4*>(y + c * 8);
    const __half2* hx = reinterpret_cast<const __half2*>(&ux);
    const __half2* hy = reinterpret_cast<const __half2*>(&uy);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fx = __half22float2(hx[e]), fy = __half22float2(hy[e]);
      acc0 = fmaf(fx.x, fy.x, acc0), acc1 = fmaf(fx.y, fy.y, acc1);
    }
  }
  return acc0 + acc1;
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) attn_bwd_tail2_kernel(AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 15) & ~uintptr_t(15));
  const int grp = threadIdx.x / TC_SIMT;                   // label group of this thread
  const int tid = threadIdx.x % TC_SIMT, warp = tid >> 5, lane = tid & 31;  // indices inside the group
  __half* sQ = reinterpret_cast<__half*>(smem + TailSmem::Q);
  __half* sK = reinterpret_cast<__half*>(smem + TailSmem::K);
  __half* sV = reinterpret_cast<__half*>(smem + TailSmem::V);
  __half* sG = reinterpret_cast<__half*>(smem + TailSmem::DO + grp * TAIL_TILE_BYTES);
  float* s_ds = reinterpret_cast<float*>(smem + TailSmem::DS) + grp * 2 * TAIL_ROWS;
  float* s_av = s_ds + TAIL_ROWS;
  float* s_red = reinterpret_cast<float*>(smem + TailSmem::RED) + grp * 8 * 192;
  float* s_wred = reinterpret_cast<float*>(smem + TailSmem::WRED) + grp * 8;
  const int T = a.T, d = a.d;
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int x0 = a.n_full * 128;  // n_tail == 1
  const __half* qkv = a.qkv16 + size_t(b) * T * a.ldq + h * TC_HD;
  const __half* Ab = a.probs16 + size_t(bh) * T * a.ldp;
  const size_t ld = size_t(a.splits) * 3 * d;
  auto dO_of = [&](int p) { return a.dO16 + size_t(p * a.B + b) * T * a.ld_do + h * TC_HD; };
  tail_stage_async(sQ, qkv, a.ldq, T, threadIdx.x, TAIL_THREADS);
  tail_stage_async(sK, qkv + d, a.ldq, T, threadIdx.x, TAIL_THREADS);
  tail_stage_async(sV, qkv + 2 * d, a.ldq, T, threadIdx.x, TAIL_THREADS);
  if (grp < a.P) tail_stage_async(sG, dO_of(grp), a.ld_do, T, tid);
  // probability column x0 (rows i) and row x0 (columns j) of this head: per-thread registers, rows tid and tid + 256
  const int r1 = tid + TC_SIMT;
  const float acol0 = tid < T ? __half2float(Ab[size_t(tid) * a.ldp + x0]) : 0.f;
  const float acol1 = r1 < T ? __half2float(Ab[size_t(r1) * a.ldp + x0]) : 0.f;
  const float arow0 = tid < T ? __half2float(Ab[size_t(x0) * a.ldp + tid]) : 0.f;
  const float arow1 = r1 < T ? __half2float(Ab[size_t(x0) * a.ldp + r1]) : 0.f;
  const int cp = tid & 31, seg = tid >> 5;  // reduction layout: channel pair (2 cp, 2 cp + 1) x 8 row segments
  const int rows_per_seg = (T + 7) / 8;
  const int i_beg = seg * rows_per_seg, i_end = min(T, (seg + 1) * rows_per_seg);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();  // Q / K / V (staged by all threads) and each group's first dO tile are visible
  for (int p = grp; p < a.P; p += TAIL_GROUPS) {
    const int pb = p * a.B + b;
    const float* dl = a.delta + (size_t(pb) * a.H + h) * T;
    const float* rp = a.r + size_t(pb) * T;
    // per-label scalars: in flight while the staged tile lands
    const float rv0 = tid < T ? rp[tid] : 0.f, rv1 = r1 < T ? rp[r1] : 0.f;
    float dv0 = 0.f, dv1 = 0.f, dx0 = 0.f;
    if (a.need_dqkv) {
      dv0 = tid < T ? dl[tid] : 0.f, dv1 = r1 < T ? dl[r1] : 0.f, dx0 = dl[x0];
    }
    if (p != grp) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      tail_group_sync(grp);  // label p's dO visible; the group's scratch of label p - 2 is consumed
    }
    const bool more = p + TAIL_GROUPS < a.P;
    // ---- column part, scalars: thread = query row(s) tid, tid + 256
    float wsum = 0.f;
    {
      const __half* v0 = sV + x0 * TAIL_PITCH;
      if (tid < T) {
        const float g = tail_dot64(sG + tid * TAIL_PITCH, v0);
        float x = g * acol0;
        if (a.positive_only) x = fmaxf(x, 0.f);
        wsum = rv0 * x;
        s_ds[tid] = acol0 * (g - dv0), s_av[tid] = acol0;
      }
      if (r1 < T) {
        const float g = tail_dot64(sG + r1 * TAIL_PITCH, v0);
        float x = g * acol1;
        if (a.positive_only) x = fmaxf(x, 0.f);
        wsum = fmaf(rv1, x, wsum);
        s_ds[r1] = acol1 * (g - dv1), s_av[r1] = acol1;
      }
    }
    wsum = warp_sum(wsum);
    if (lane == 0) s_wred[warp] = wsum;
    tail_group_sync(grp);
    if (tid == 0) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += s_wred[k];
      a.wpart[(size_t(pb) * a.H + h) * T + x0] = s / a.H;
    }
    if (!a.need_dqkv) {  // (block-uniform) the tile is free: every thread of the group passed the barrier after its last read
      if (more) tail_stage_async(sG, dO_of(p + TAIL_GROUPS), a.ld_do, T, tid);
      continue;
    }
    // ---- column part, vectors: dK_x0 = sum_i ds_i Q_i ; dV_x0 = sum_i a_i dO_i
    {
      float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
#pragma unroll 4
      for (int i = i_beg; i < i_end; ++i) {
        const float ds = s_ds[i], av = s_av[i];
        const float2 fq = __half22float2(*reinterpret_cast<const __half2*>(sQ + i * TAIL_PITCH + 2 * cp));
        const float2 fg = __half22float2(*reinterpret_cast<const __half2*>(sG + i * TAIL_PITCH + 2 * cp));
        k0 = fmaf(ds, fq.x, k0), k1 = fmaf(ds, fq.y, k1);
        v0 = fmaf(av, fg.x, v0), v1 = fmaf(av, fg.y, v1);
      }
      float* rd = s_red + seg * 192;
      rd[2 * cp] = k0, rd[2 * cp + 1] = k1, rd[64 + 2 * cp] = v0, rd[64 + 2 * cp + 1] = v1;
    }
    tail_group_sync(grp);  // s_ds / s_av consumed; s_red (dK, dV) written
    // ---- row part, scalars: thread = key(s) j = tid, tid + 256
    {
      const __half* g0 = sG + x0 * TAIL_PITCH;
      if (tid < T) s_ds[tid] = arow0 * (tail_dot64(g0, sV + tid * TAIL_PITCH) - dx0);
      if (r1 < T) s_ds[r1] = arow1 * (tail_dot64(g0, sV + r1 * TAIL_PITCH) - dx0);
    }
    tail_group_sync(grp);
    if (more) tail_stage_async(sG, dO_of(p + TAIL_GROUPS), a.ld_do, T, tid);  // nobody reads the tile again: lands during the rest
    {
      float q0 = 0.f, q1 = 0.f;
#pragma unroll 4
      for (int j = i_beg; j < i_end; ++j) {
        const float ds = s_ds[j];
        const float2 fk = __half22float2(*reinterpret_cast<const __half2*>(sK + j * TAIL_PITCH + 2 * cp));
        q0 = fmaf(ds, fk.x, q0), q1 = fmaf(ds, fk.y, q1);
      }
      float* rd = s_red + seg * 192;
      rd[128 + 2 * cp] = q0, rd[128 + 2 * cp + 1] = q1;
    }
    tail_group_sync(grp);
    if (tid < 96) {  // 3 vectors (dK, dV, dQ) x 32 channel pairs
      const int vec = tid >> 5, c2 = tid & 31;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int sg = 0; sg < 8; ++sg) s0 += s_red[sg * 192 + vec * 64 + 2 * c2], s1 += s_red[sg * 192 + vec * 64 + 2 * c2 + 1];
      __half* orow = a.dqkv16 + (size_t(pb) * T + x0) * ld + h * TC_HD;
      const int col = (vec == 0 ? d : vec == 1 ? 2 * d : 0) + 2 * c2;
      const float sc = vec == 2 ? a.scale : 1.0f;
      uint32_t hi, lo;
      split_pack(s0 * sc, s1 * sc, hi, lo);
      *reinterpret_cast<uint32_t*>(orow + col) = hi;
      if (a.splits == 2) *reinterpret_cast<uint32_t*>(orow + 3 * d + col) = lo;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// launch helpers shared with the third generation (vit_attn_bwd3.cu reuses the delta and tail kernels)
int launch_attn_delta(const AttnBwdTcArgs& a, cudaStream_t st) {
  SB_REQUIRE(a.H <= 32, "attention backward: at most 32 heads");
  const int rows = a.B * a.T;
  attn_delta_kernel<<<(rows + DELTA_ROWS - 1) / DELTA_ROWS, DELTA_ROWS * a.H, 0, st>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_attn_tail2(const AttnBwdTcArgs& a, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tail2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TailSmem::TOTAL + 16));
    configured = true;
  }
  attn_bwd_tail2_kernel<<<a.B * a.H, TAIL_THREADS, TailSmem::TOTAL + 16, st>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sb

using namespace sb;

// Second-generation attention backward: same contract as semabs_attn_bwd_tc (include/semabs_b200.h); T <= 272 and
// T % 128 in {0, 1} (the ViT geometries: 50 -> one tile, 257 -> two tiles + one tail row; other remainders take the
// first-generation kernels through semabs_attn_bwd_tc).
extern "C" int semabs_attn_bwd_tc2(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                                   const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart,
                                   void* dqkv16, int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits,
                                   int32_t positive_only, int32_t need_dqkv, void* stream) {
  SB_REQUIRE(qkv16 && probs16 && o32 && dO16 && r && delta_ws && wpart, "semabs_attn_bwd_tc2: null pointer");
  SB_REQUIRE(!need_dqkv || dqkv16, "semabs_attn_bwd_tc2: dqkv16 missing");
  SB_REQUIRE(P > 0 && B > 0 && T > 0 && H > 0 && T <= TC_MAX_T, "semabs_attn_bwd_tc2: bad shape (T <= %d)", TC_MAX_T);
  SB_REQUIRE(ld_p16 % 16 == 0 && ld_p16 >= ((T + 15) / 16) * 16, "semabs_attn_bwd_tc2: bad probs16 pitch %d", ld_p16);
  const int d = H * TC_HD;
  SB_REQUIRE(ld_qkv >= 3 * d && ld_qkv % 8 == 0 && ld_do >= d && ld_do % 16 == 0, "semabs_attn_bwd_tc2: bad pitch");
  SB_REQUIRE(splits == 1 || splits == 2, "semabs_attn_bwd_tc2: splits must be 1 or 2");
  const int rem = T % 128;
  SB_REQUIRE(T <= 128 || rem <= 1, "semabs_attn_bwd_tc2: T %% 128 must be 0 or 1 above one tile (got T=%d)", T);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tm_qkv, tm_do, tm_pr;
  if (int rc = make_tile_tmap(&tm_qkv, qkv16, (long long)B * T, 3LL * d, ld_qkv)) return rc;
  if (int rc = make_tile_tmap(&tm_do, dO16, (long long)P * B * T, d, ld_do)) return rc;
  if (int rc = make_tile_tmap(&tm_pr, probs16, (long long)B * H * T, ld_p16, ld_p16)) return rc;
  AttnBwdTcArgs a{};
  a.qkv16 = (const __half*)qkv16, a.ldq = ld_qkv;
  a.probs16 = (const __half*)probs16, a.ldp = ld_p16, a.o32 = o32, a.dO16 = (const __half*)dO16, a.ld_do = ld_do;
  a.delta = delta_ws, a.r = r, a.wpart = wpart, a.dqkv16 = (__half*)dqkv16;
  a.P = P, a.B = B, a.T = T, a.H = H, a.d = d, a.splits = splits, a.scale = 0.125f;
  a.positive_only = positive_only, a.need_dqkv = need_dqkv;
  a.n_tail = (T > 128 && rem == 1) ? 1 : 0;
  a.n_full = a.n_tail ? T / 128 : (T + 127) / 128;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_row_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RowSmem::TOTAL));
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_col_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ColSmem::TOTAL));
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tail2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TailSmem::TOTAL + 16));
    configured = true;
  }
  const int n_units = B * H * a.n_full;
  const int grid = n_units < num_sms() ? n_units : num_sms();
  if (need_dqkv) {  // delta feeds dS in both passes and the tail; the relevance-only last step needs neither delta nor dQ
    SB_REQUIRE(H <= 32, "semabs_attn_bwd_tc2: at most 32 heads");
    const int rows = B * T;
    attn_delta_kernel<<<(rows + DELTA_ROWS - 1) / DELTA_ROWS, DELTA_ROWS * H, 0, st>>>(a);
    SB_CHECK_CUDA(cudaGetLastError());
    attn_bwd_row_tc2_kernel<<<grid, TC_BWD_THREADS, RowSmem::TOTAL, st>>>(tm_qkv, tm_do, a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  attn_bwd_col_tc2_kernel<<<grid, TC_BWD_THREADS, ColSmem::TOTAL, st>>>(tm_qkv, tm_do, tm_pr, a);
  SB_CHECK_CUDA(cudaGetLastError());
  if (a.n_tail) {
    attn_bwd_tail2_kernel<<<B * H, TAIL_THREADS, TailSmem::TOTAL + 16, st>>>(a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
