// Attention backward + relevance on tcgen05, third generation: the row / column passes of vit_attn_bwd2.cu (same math, same
// TMEM strip layout, same operand descriptors — see the derivations in vit_attn_tc.cu) re-scheduled as a two-chunk software
// pipeline.  Reference: ClipGradcam.interpret, CLIP/clip/clip_gradcam.py:90-126 over the autograd graph of
// auxiliary.multi_head_attention_forward, CLIP/clip/auxiliary.py:306-345.
//
// What the second generation measured (profiles/r02_ncu_full_attn_bwd2.txt): tensor pipe 17-19 % busy, ~4.9 k cycles per
// (unit, label) item in the row pass against ~1.1 k cycles of MMA work, because the four phases of an item
//     G = dO V^T  ->  element-wise dS  ->  dQ = dS K  ->  epilogue
// run back to back: the single strip of 272 TMEM columns is the output of the first MMA and the A operand of the second, so
// the tensor pipe idles during the element-wise phase and the element-wise warps idle during both MMAs.
//
// Here the strip is cut into two column chunks (ViT-L/14: 144 + 128 of the 272 columns) with their own barriers:
//   * element-wise warps 0-3 own chunk 0, warps 4-7 chunk 1 (the same column split as before, so the register-resident
//     probability rows / columns are unchanged) and never wait for each other;
//   * the control warp issues  dX(n, c) ; G(n+1, c)  as soon as chunk c of item n has been packed: tcgen05.mma instructions
//     execute in issue order, so G(n+1, c) may overwrite the chunk that dX(n, c) reads without a round trip through a barrier;
//     while one half of the warps runs its element-wise phase the tensor pipe works on the other chunk;
//   * row pass: two dQ accumulators (the epilogue of item n-1 runs after the element-wise phase of item n), K / V double
//     buffered over units and three dO stages, so nothing on the tensor side waits at a unit boundary;
//   * column pass: dK + dV accumulators are single (272 + 2 x 128 columns do not fit TMEM); chunk-0 warps drain dK of item
//     n-1 after their element-wise phase of item n, chunk-1 warps drain dV of item n before theirs of item n+1;
//     {delta_i, r_i} staging and the relevance partial sums are per chunk owner.
// delta and the one-row / one-key tail are the second generation's kernels (launch helpers in vit_attn_bwd2.cu).
//
// Measured on B200 while tuning (tools/ubench/*.cu, tools/attn_trace.py = clock64 time stamps of CTA 0 through
// semabs_debug_attn_trace, ncu source-level stall samples under profiles/r02_ncu_full_attn_bwd3.txt):
//   * one thread issues a tcgen05.mma every max(N / 2, ~36 + N / 4) cycles when A comes from shared memory (N = 64: 48, 144: 72,
//     256: 128) and every N / 2 cycles when A comes from TMEM (N = 64: 32); the instruction does not return before the pipe
//     accepts it, so the issuing warp's own bookkeeping between two MMAs is tensor idle time (compile-time chunk geometry took
//     the row pass's control warp from ~70 to ~45 cycles per MMA); a TMA issue costs its thread ~500 cycles;
//   * tcgen05.ld x16 + wait::ld is a 23-cycle round trip on an idle SM (3.4 KB / clk / SM with 16 warps), but ~120 cycles
//     while the tensor pipe is busy; steps of 8 or 32 columns were both slower than 16;
//   * the element-wise warps are the limit: 72 probability registers per thread + 16 strip values + the 168-register cap of a
//     9-warp CTA leave ptxas no room to run loads ahead (it folds a second prefetch buffer into the first), so each warp is a
//     serial chain at ~0.4 instructions / cycle, two warps per scheduler.  Four pipeline chunks instead of two, test_wait spin
//     loops instead of try_wait, and a separate TMA-issue warp were each built and measured: 0.60 / 1.89 ms against 0.56 /
//     1.76 ms (row / column) for this version — no gain, reverted.  What would move it: probabilities in
//     TMEM instead of registers (row pass: 272 + 136 + 64 columns), 16 element-wise warps, dV as a label-batched N = 256 GEMM.
//   * a tail without resident Q / K / V tiles (two CTAs per SM: dO stages + the ds of all labels in shared memory, Q / K rows
//     read from global in the final reductions) was built too: 0.60 + 0.35 ms against 0.79 ms for the second generation's
//     tail kernel — the global-latency-bound reductions cost more than the occupancy gained; reverted.
#include "vit_attn_tc.cuh"

namespace sb {

int launch_attn_delta(const AttnBwdTcArgs& a, cudaStream_t st);  // vit_attn_bwd2.cu
int launch_attn_tail2(const AttnBwdTcArgs& a, cudaStream_t st);  // vit_attn_bwd2.cu

constexpr int T3_THREADS = 288;  // 8 element-wise warps (2 per TMEM lane quadrant: chunk 0 / chunk 1) + 1 control warp
constexpr int T3_SIMT = 256;
constexpr int T3_NCH = (TC_MAX_T / 16 + 1) / 2;  // 9 chunks of 16 strip columns per thread at most

// dS pair = A o (G - delta) as the fp16 MMA operand: the difference is formed in fp32, rounded to fp16 and multiplied by the
// (fp16) probabilities with one packed multiply — 4 instructions per pair instead of 7 (two unpacks, two subtracts, two
// multiplies, one pack); the element-wise warps are issue-bound, not the tensor pipe.  One extra fp16 rounding of a factor.
__device__ __forceinline__ uint32_t ds_pair(uint32_t a_pair, float x0, float x1) {
  const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&a_pair), __floats2half2_rn(x0, x1));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// Shared-memory load the compiler may schedule freely (not volatile: a volatile one is issued right in front of its first use
// and its latency shows up 72 times per item).  Ordering against the barrier that publishes the data comes from the address:
// the caller passes it through order_after_wait() once the wait has returned.
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t order_after_wait(uint32_t x) {
  asm volatile("" : "+r"(x)::"memory");
  return x;
}

// Element-wise phase of the column pass over NC 16-query chunks of one thread's (= one key's) strip range; {delta_i, r_i} come
// from shared memory (warp-uniform addresses).  NC > 0: compile-time chunk count (ViT-L/14: 9 and 8), NC == 0: run-time count.
// Measured alternatives at the bench shape (1.62 ms as written): steps of 8 queries 1.83 ms, 32 columns per TMEM round trip
// 1.84 ms, a prefetching second 16-register buffer 1.62 ms — ptxas folds the two buffers into one: the 72 probability
// registers of a thread leave no room under the 168-register cap of a 9-warp CTA, which is what bounds this phase.
template <int NC>
__device__ __forceinline__ void col_elementwise(uint32_t t_strip, uint32_t dr_addr, const uint32_t (&acol)[T3_NCH][8], int nc,
                                                float clamp_lo, float& w0, float& w1) {
  constexpr int MAXC = NC ? NC : T3_NCH;
#pragma unroll
  for (int cc = 0; cc < MAXC; ++cc) {
    if (NC == 0 && cc >= nc) break;
    uint32_t g[16], out[8];
    tmem_ld_32x32b_x16(t_strip + uint32_t(cc * 16), g);
    tc_wait_ld();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t ap = acol[cc][e];
      const float2 av = __half22float2(*reinterpret_cast<const __half2*>(&ap));
      const float4 d4 = lds128(dr_addr + uint32_t(cc * 16 + 2 * e) * 8);  // {delta_i, r_i, delta_i+1, r_i+1}
      const float g0 = __uint_as_float(g[2 * e]), g1 = __uint_as_float(g[2 * e + 1]);
      // relevance term r_i max(g a, 0) = r_i (a max(g, 0)) since a >= 0
      w0 = fmaf(d4.y, av.x * fmaxf(g0, clamp_lo), w0);
      w1 = fmaf(d4.w, av.y * fmaxf(g1, clamp_lo), w1);
      out[e] = ds_pair(ap, g0 - d4.x, g1 - d4.z);
    }
    tmem_st_32x32b_x8(t_strip + uint32_t(cc * 16), out);
    tmem_st_32x32b_x8(t_strip + uint32_t(cc * 16 + 8), acol[cc]);
  }
}

struct Row3Smem {
  static constexpr int KVBUF = 2 * TC_KV_BYTES;          // K then V of one unit
  static constexpr int KV = 0;                           // 2 buffers (unit parity)
  static constexpr int DO = 2 * KVBUF;                   // 3 stages of one 136-row box
  static constexpr int BARS = DO + 3 * TC_BOX_BYTES;
  static constexpr int TOTAL = BARS + 256 + 1024;
};
static_assert(Row3Smem::TOTAL <= 232448, "row pass shared memory");

// ---------------------------------------------------------------------------------------------------------
// row pass: thread = query row i.  G = dO V^T -> dS = A o (G - delta) packed in place -> dQ = scale dS K
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T3_THREADS, 1)
attn_bwd_row_tc3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Row3Smem::BARS);
  uint64_t *bar_kv = bars /* [2] */, *bar_do = bars + 2 /* [3] */, *bar_s = bars + 5 /* [2] */, *bar_p = bars + 7 /* [2] */;
  uint64_t *bar_o = bars + 9 /* [2] */, *bar_e = bars + 11 /* [2] */;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, d = a.d, P = a.P;
  const int ncol = (T + 15) & ~15;
  const int nch = ncol / 16, c_split = (nch + 1) / 2;
  const int nA = c_split * 16, nB = ncol - nA;  // strip columns of chunk 0 / chunk 1
  const int n_units = a.B * a.H * a.n_full;
  const int n_my = (n_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int N = n_my * P;  // items of this CTA: unit-major, label fastest

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_kv[i], 1), mbar_init(&bar_s[i], 1), mbar_init(&bar_o[i], 1);
      mbar_init(&bar_p[i], T3_SIMT / 2), mbar_init(&bar_e[i], T3_SIMT);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&bar_do[i], 1);
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
  const uint32_t tS = tmem_base + TC_COL_S;

  // unit ul of this CTA -> (sequence b, head h, 128-row tile mt); the divisions run once per unit, not per item
  auto unit_coord = [&](int ul, int& bh, int& b, int& h, int& mt) {
    const int u = int(blockIdx.x) + ul * int(gridDim.x);
    bh = u / a.n_full, mt = u - bh * a.n_full;
    b = bh / a.H, h = bh - b * a.H;
  };

  if (warp == 8) {
    // ===== control warp: TMA + MMA issue (convergent; only the elected lane's instructions take effect) =====
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc_gA = make_idesc_f16(128, nA), idesc_gB = make_idesc_f16(128, nB ? nB : 16);
    constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);
    // dO loader position (runs two items ahead of the main loop)
    int l_n = 0, l_p = 0, l_ul = 0, l_b = 0, l_h = 0, l_mt = 0, l_bh = 0;
    unit_coord(0, l_bh, l_b, l_h, l_mt);
    auto load_do = [&]() {
      if (leader) {
        const int st = l_n % 3;
        mbar_arrive_expect_tx(&bar_do[st], TC_BOX_BYTES);
        tma_load_2d(smem + Row3Smem::DO + st * TC_BOX_BYTES, &tm_do, &bar_do[st], l_h * TC_HD, (l_p * a.B + l_b) * T + l_mt * 128);
      }
      ++l_n;
      if (++l_p == P) {
        l_p = 0, ++l_ul;
        if (l_ul < n_my) unit_coord(l_ul, l_bh, l_b, l_h, l_mt);
      }
    };
    auto load_kv = [&](int ul) {
      int bh, b, h, mt;
      unit_coord(ul, bh, b, h, mt);
      uint8_t* buf = smem + Row3Smem::KV + (ul & 1) * Row3Smem::KVBUF;
      mbar_arrive_expect_tx(&bar_kv[ul & 1], 2u * TC_KV_BYTES);
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_2d(buf + bx * TC_BOX_BYTES, &tm_qkv, &bar_kv[ul & 1], d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
        tma_load_2d(buf + TC_KV_BYTES + bx * TC_BOX_BYTES, &tm_qkv, &bar_kv[ul & 1], 2 * d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
      }
    };
    // G issue position: item g_n = label g_p of unit g_ul (both chunks of an item, then the next item)
    int g_n = 0, g_p = 0, g_ul = 0;
    auto issue_g = [&](int c) {
      const int st = g_n % 3;
      if (c == 0) {
        if (g_p == 0) mbar_wait(&bar_kv[g_ul & 1], (g_ul >> 1) & 1);
        mbar_wait(&bar_do[st], (g_n / 3) & 1);
        tc_fence_after();
      }
      if (c == 0 || nB) {
        const uint64_t dDO = desc_kmajor(sbase + Row3Smem::DO + st * TC_BOX_BYTES);
        const uint64_t dV = desc_kmajor(sbase + Row3Smem::KV + (g_ul & 1) * Row3Smem::KVBUF + TC_KV_BYTES + (c ? nA * 128 : 0));
#pragma unroll
        for (int k = 0; k < TC_HD / 16; ++k)
          umma_f16_elect(tS + uint32_t(c ? nA : 0), dDO + uint64_t(2 * k), dV + uint64_t(2 * k), c ? idesc_gB : idesc_gA, k != 0, leader);
      }
      umma_commit_elect(&bar_s[c], leader);
      if (c == 1) {
        ++g_n;
        if (++g_p == P) g_p = 0, ++g_ul;
      }
    };
    auto issue_q = [&](int n, int ul, int c) {
      const uint64_t dK = desc_mnmajor(sbase + Row3Smem::KV + (ul & 1) * Row3Smem::KVBUF, 16);
      const uint32_t tO = tmem_base + uint32_t((n & 1) ? TC_COL_O2 : TC_COL_O);
      const int s0 = c ? c_split : 0, s1 = c ? nch : c_split;
      for (int s = s0; s < s1; ++s) umma_f16_ts_elect(tO, tS + uint32_t(16 * s), dK + uint64_t(s) * (2048 >> 4), idesc_o, s > 0, leader);
      if (c == 1) umma_commit_elect(&bar_o[n & 1], leader);
    };
    if (N > 0) {
      if (leader) {
        load_kv(0);
        if (n_my > 1) load_kv(1);
      }
      load_do();
      if (N > 1) load_do();
      __syncwarp();
      issue_g(0);
      issue_g(1);
    }
    int m_p = 0, m_ul = 0;  // main position: item n = label m_p of unit m_ul
    for (int n = 0; n < N; ++n) {
      // dO stage (n + 2) % 3 held item n - 1, whose G chunks were both consumed (bar_p waits of the previous iteration)
      if (n + 2 < N) load_do();
      __syncwarp();
      mbar_wait(&bar_p[0], n & 1);  // chunk 0 of the strip holds packed dS(n)
      tc_trace(a, 0, 0, n);
      if (n >= 2) mbar_wait(&bar_e[n & 1], ((n - 2) >> 1) & 1);  // accumulator n & 1 drained by the epilogue of item n - 2
      tc_fence_after();
      tc_trace(a, 0, 1, n);
      issue_q(n, m_ul, 0);
      if (n + 1 < N) issue_g(0);  // in issue order behind dQ(n, 0): may overwrite chunk 0
      tc_trace(a, 0, 2, n);
      mbar_wait(&bar_p[1], n & 1);
      tc_fence_after();
      tc_trace(a, 0, 3, n);
      issue_q(n, m_ul, 1);
      if (n + 1 < N) issue_g(1);
      tc_trace(a, 0, 4, n);
      if (m_p == 0 && n > 0) {
        // unit m_ul has started; the previous unit's K / V buffer is free once dQ(n - 1) has completed (long ago)
        mbar_wait(&bar_o[(n - 1) & 1], ((n - 1) >> 1) & 1);
        if (leader && m_ul + 1 < n_my) load_kv(m_ul + 1);
        __syncwarp();
      }
      if (++m_p == P) m_p = 0, ++m_ul;
    }
  } else {
    // ===== element-wise warps: half 0 owns strip chunk 0, half 1 chunk 1 =====
    const int q = warp & 3, half = warp >> 2, rr = q * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    const int c0 = half ? c_split : 0, c1 = half ? nch : c_split;
    const size_t S3 = size_t(a.splits) * 3 * d;
    const size_t d_stride = size_t(a.B) * a.H * T;   // delta: label p -> p + 1
    const size_t o_stride = size_t(a.B) * T * S3;    // dqkv16 rows: label p -> p + 1
    uint32_t arow[T3_NCH][8];  // this thread's half of its probability row (kept across the P labels of a unit)
    auto epilogue = [&](int m, __half* orow, bool valid) {
      mbar_wait(&bar_o[m & 1], (m >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_row + uint32_t(((m & 1) ? TC_COL_O2 : TC_COL_O) + 32 * half), o);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&bar_e[m & 1]);
      if (valid) store_row_f16(orow, 3 * d, a.splits, o, 32, a.scale);
    };
    // per-unit state, advanced by pointer additions per label (no per-item multiplies, nothing carried for the next unit:
    // the register file is full of probabilities and a spill reload on this path stalls the whole element-wise warp)
    bool valid = false, prev_valid = false;
    const float* dptr = nullptr;
    __half *optr = nullptr, *prev_o = nullptr;
    int p = 0, ul = 0;
    float dnext = 0.f;
    for (int n = 0; n < N; ++n) {
      if (p == 0) {
        int bh, b, h, mt;
        unit_coord(ul, bh, b, h, mt);
        const int i = mt * 128 + rr;
        valid = i < T;
        const int ic = valid ? i : 0;
        const __half* prow = a.probs16 + (size_t(bh) * T + ic) * a.ldp;
        dptr = a.delta + (size_t(b) * a.H + h) * T + ic;
        optr = a.dqkv16 + (size_t(b) * T + ic) * S3 + h * TC_HD + 32 * half;
        dnext = valid ? *dptr : 0.f;
#pragma unroll
        for (int cc = 0; cc < T3_NCH; ++cc) {
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
      if (p + 1 < P) {  // next label's scalar in flight during this item's element-wise phase
        dptr += d_stride;
        dnext = valid ? *dptr : 0.f;
      }

      if (q == 0) tc_trace(a, 0, 5 + 5 * half, n);  // loop top
      mbar_wait(&bar_s[half], n & 1);
      tc_fence_after();
      if (q == 0) tc_trace(a, 0, 6 + 5 * half, n);  // G chunk seen
      // TMEM loads run one 16-column chunk ahead of the arithmetic (two register buffers)
      uint32_t gbuf[2][16];
      tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c0 * 16), gbuf[0]);
      tc_wait_ld();
#pragma unroll
      for (int cc = 0; cc < T3_NCH; ++cc) {
        if (c0 + cc < c1) {
          const int c = c0 + cc;
          const bool has_next = (cc + 1 < T3_NCH) && (c + 1 < c1);  // warp-uniform
          uint32_t dtok = __float_as_uint(delta);  // ordering token: this chunk's arithmetic stays behind the prefetch
          if (has_next) tmem_ld_32x32b_x16_tok(t_row + uint32_t(TC_COL_S + (c + 1) * 16), gbuf[(cc + 1) & 1], dtok);
          const float dl = __uint_as_float(dtok);
          const uint32_t(&g)[16] = gbuf[cc & 1];
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = ds_pair(arow[cc][e], __uint_as_float(g[2 * e]) - dl, __uint_as_float(g[2 * e + 1]) - dl);
          tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c * 16), w);
          if (has_next) tc_wait_ld();
        }
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_p[half]);
      if (q == 0) tc_trace(a, 0, 7 + 5 * half, n);  // chunk packed, arrived
      if (n >= 1) epilogue(n - 1, prev_o, prev_valid);  // its dQ has had the whole element-wise phase of item n to complete
      if (q == 0) tc_trace(a, 0, 8 + 5 * half, n);  // epilogue of item n - 1 done
      prev_o = optr, prev_valid = valid;
      optr += o_stride;
      if (++p == P) p = 0, ++ul;
    }
    if (N > 0) epilogue(N - 1, prev_o, prev_valid);
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
__global__ void __launch_bounds__(T3_THREADS, 1)
attn_bwd_col_tc3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_pr, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ColSmem::BARS);
  uint64_t *bar_kv = bars, *bar_do = bars + 1 /* [2] */, *bar_s = bars + 3 /* [2] */, *bar_p = bars + 5 /* [2] */;
  uint64_t *bar_o = bars + 7, *bar_e = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float2* s_dr = reinterpret_cast<float2*>(smem + ColSmem::DR);  // [2 items][272 query rows] {delta_i, r_i}
  float* s_w = reinterpret_cast<float*>(smem + ColSmem::W);      // [2 items][2 halves][128 keys]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, d = a.d, P = a.P;
  const int ncol = (T + 15) & ~15;  // query columns of G^T
  const int nch = ncol / 16, c_split = (nch + 1) / 2;
  const int nA = c_split * 16, nB = ncol - nA;
  const int n_units = a.B * a.H * a.n_full;
  const int n_my = (n_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int N = n_my * P;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_pr);
    mbar_init(bar_kv, 1), mbar_init(&bar_do[0], 1), mbar_init(&bar_do[1], 1), mbar_init(bar_o, 1), mbar_init(bar_e, T3_SIMT);
    for (int i = 0; i < 2; ++i) mbar_init(&bar_s[i], 1), mbar_init(&bar_p[i], T3_SIMT / 2);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  auto unit_coord = [&](int ul, int& bh, int& b, int& h, int& mt) {
    const int u = int(blockIdx.x) + ul * int(gridDim.x);
    bh = u / a.n_full, mt = u - bh * a.n_full;
    b = bh / a.H, h = bh - b * a.H;
  };
  // {delta_i, r_i} of an item, staged by the owner of the chunk the query row i belongs to: chunk 0 rows t and t + 128
  // (< nA), chunk 1 row nA + t, for t = thread index within its half
  const int half_t = int(threadIdx.x) & 127, my_half = (int(threadIdx.x) >> 7) & 1;
  const int row_a = my_half ? nA + half_t : half_t;
  const int row_b = my_half ? ncol : half_t + 128;  // second row of a chunk-0 thread (valid when < nA)
  const bool ok_a = row_a < T && (my_half || row_a < nA), ok_b = row_b < nA && row_b < T;
  const size_t d_stride = size_t(a.B) * a.H * T;  // delta / wpart: label p -> p + 1
  const size_t r_stride = size_t(a.B) * T;        // r: label p -> p + 1
  auto fetch_dr = [&](const float* dbase, const float* rbase, float2& v0, float2& v1) {
    v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
    if (ok_a) v0 = make_float2(a.need_dqkv ? dbase[row_a] : 0.f, rbase[row_a]);
    if (ok_b) v1 = make_float2(a.need_dqkv ? dbase[row_b] : 0.f, rbase[row_b]);
  };
  auto store_dr = [&](int n, const float2& v0, const float2& v1) {
    if (my_half ? row_a < ncol : row_a < nA) s_dr[(n & 1) * TC_MAX_T + row_a] = v0;  // own chunk only: the halves run apart
    if (row_b < nA) s_dr[(n & 1) * TC_MAX_T + row_b] = v1;
  };
  if (warp < 8 && N > 0) {
    int bh, b, h, mt;
    unit_coord(0, bh, b, h, mt);
    float2 v0, v1;
    fetch_dr(a.delta + (size_t(b) * a.H + h) * T, a.r + size_t(b) * T, v0, v1);
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
    const uint32_t idesc_gA = make_idesc_f16(128, nA), idesc_gB = make_idesc_f16(128, nB ? nB : 16);
    constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);
    // dO loader position (two items ahead of the main loop)
    int l_n = 0, l_p = 0, l_ul = 0, l_b = 0, l_h = 0, l_mt = 0, l_bh = 0;
    unit_coord(0, l_bh, l_b, l_h, l_mt);
    auto load_do = [&]() {
      if (leader) {
        const int st = l_n & 1;
        mbar_arrive_expect_tx(&bar_do[st], TC_KV_BYTES);
        for (int bx = 0; bx < 2; ++bx)
          tma_load_2d(smem + ColSmem::DO + st * TC_KV_BYTES + bx * TC_BOX_BYTES, &tm_do, &bar_do[st], l_h * TC_HD,
                      (l_p * a.B + l_b) * T + bx * TC_BOX_ROWS);
      }
      ++l_n;
      if (++l_p == P) {
        l_p = 0, ++l_ul;
        if (l_ul < n_my) unit_coord(l_ul, l_bh, l_b, l_h, l_mt);
      }
    };
    auto load_unit = [&](int ul) {
      int bh, b, h, mt;
      unit_coord(ul, bh, b, h, mt);
      mbar_arrive_expect_tx(bar_kv, 7u * TC_BOX_BYTES);
      for (int bx = 0; bx < 2; ++bx)
        tma_load_2d(smem + ColSmem::Q + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, h * TC_HD, b * T + bx * TC_BOX_ROWS);
      tma_load_2d(smem + ColSmem::V, &tm_qkv, bar_kv, 2 * d + h * TC_HD, b * T + mt * 128);
      for (int cb = 0; cb < 2; ++cb)
        for (int bx = 0; bx < 2; ++bx)
          tma_load_2d(smem + ColSmem::PR + (cb * 2 + bx) * TC_BOX_BYTES, &tm_pr, bar_kv, mt * 128 + cb * 64, bh * T + bx * TC_BOX_ROWS);
    };
    int g_n = 0, g_p = 0, g_ul = 0;  // G issue position
    auto issue_g = [&](int c) {
      if (c == 0) {
        if (g_p == 0) mbar_wait(bar_kv, g_ul & 1);
        mbar_wait(&bar_do[g_n & 1], (g_n >> 1) & 1);
        tc_fence_after();
      }
      if (c == 0 || nB) {
        const uint64_t dDOk = desc_kmajor(sbase + ColSmem::DO + (g_n & 1) * TC_KV_BYTES + (c ? nA * 128 : 0));
#pragma unroll
        for (int k = 0; k < TC_HD / 16; ++k)
          umma_f16_elect(tS + uint32_t(c ? nA : 0), dVt + uint64_t(2 * k), dDOk + uint64_t(2 * k), c ? idesc_gB : idesc_gA, k != 0, leader);
      }
      umma_commit_elect(&bar_s[c], leader);
      if (c == 1) {
        ++g_n;
        if (++g_p == P) g_p = 0, ++g_ul;
      }
    };
    auto issue_kv = [&](int n, int c) {
      if (a.need_dqkv) {
        const uint64_t dDOm = desc_mnmajor(sbase + ColSmem::DO + (n & 1) * TC_KV_BYTES, 16);
        const int s0 = c ? c_split : 0, s1 = c ? nch : c_split;
        for (int s = s0; s < s1; ++s) {
          const uint64_t kadv = uint64_t(s) * (2048 >> 4);
          umma_f16_ts_elect(tK, tS + uint32_t(16 * s), dQm + kadv, idesc_o, s > 0, leader);
          umma_f16_ts_elect(tV, tS + uint32_t(16 * s + 8), dDOm + kadv, idesc_o, s > 0, leader);
        }
      }
      if (c == 1) umma_commit_elect(bar_o, leader);
    };
    if (N > 0) {
      if (leader) load_unit(0);
      load_do();
      if (N > 1) load_do();
      __syncwarp();
      issue_g(0);
      issue_g(1);
    }
    int m_p = 0, m_ul = 0;
    for (int n = 0; n < N; ++n) {
      const bool more = n + 1 < N, boundary = m_p + 1 == P;
      mbar_wait(&bar_p[0], n & 1);
      tc_trace(a, 1, 0, n);
      if (n >= 1) mbar_wait(bar_e, (n - 1) & 1);  // dK / dV accumulators drained by the epilogues of item n - 1
      tc_fence_after();
      tc_trace(a, 1, 1, n);
      issue_kv(n, 0);
      if (more && !boundary) issue_g(0);
      tc_trace(a, 1, 2, n);
      mbar_wait(&bar_p[1], n & 1);
      tc_fence_after();
      tc_trace(a, 1, 3, n);
      issue_kv(n, 1);
      if (more && !boundary) issue_g(1);
      tc_trace(a, 1, 4, n);
      // dO stage n & 1 (B operand of dV(n)) and, at a unit boundary, Q / V / the probability tile are free once item n's
      // MMAs have completed; the next item's G chunks are already queued behind them
      mbar_wait(bar_o, n & 1);
      tc_fence_after();
      tc_trace(a, 1, 15, n);
      if (n + 2 < N) load_do();
      if (more && boundary) {
        if (leader) load_unit(m_ul + 1);
        __syncwarp();
        issue_g(0);
        issue_g(1);
      }
      __syncwarp();
      if (++m_p == P) m_p = 0, ++m_ul;
    }
  } else {
    // ===== element-wise warps: half 0 owns query chunk 0 and drains dK, half 1 owns chunk 1 and drains dV =====
    const int q = warp & 3, half = warp >> 2, jj = q * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    const int c0 = half ? c_split : 0, c1 = half ? nch : c_split;
    // probability tile addressing: element (i, jj) sits at  i*128 + (((jj%64)/8 ^ (i%8)) << 4) + (jj%8)*2  inside column
    // block jj/64 (the two 136-row TMA boxes of a column block are contiguous and 136 % 8 == 0)
    const uint8_t* pcol = smem + ColSmem::PR + (jj >> 6) * (2 * TC_BOX_BYTES) + (jj & 7) * 2;
    const float invH = 1.0f / a.H;
    const float clamp_lo = a.positive_only ? 0.f : -INFINITY;
    const size_t S3 = size_t(a.splits) * 3 * d;
    const size_t o_stride = size_t(a.B) * T * S3;
    uint32_t acol[T3_NCH][8];  // A[i, j] for this thread's key j and its chunk of the query rows i, packed pairs
    // per-unit state advanced by pointer additions per label; the epilogue of the previous item (chunk-0 owners) gets its
    // two pointers and the validity flag from copies taken before the advance
    // Output pointers describe the item whose epilogue comes next: the current item for chunk-1 owners, the previous one
    // for chunk-0 owners (they are set / advanced after that epilogue).
    bool valid = false;
    const float *dbase = nullptr, *rbase = nullptr;  // delta / r of the NEXT item to stage, query row 0
    float* wptr = nullptr;                            // wpart of this key
    __half* optr = nullptr;                           // dK (half 0) / dV (half 1) row of this key
    auto set_out = [&](int u_local) {
      int bh, b, h, mt;
      unit_coord(u_local, bh, b, h, mt);
      const int j = mt * 128 + jj;
      valid = j < T;
      const int jc = valid ? j : 0;
      wptr = a.wpart + (size_t(b) * a.H + h) * T + jc;
      optr = a.dqkv16 + (size_t(b) * T + jc) * S3 + (half ? 2 * d : d) + h * TC_HD;
    };
    auto epilogue = [&](int m, __half* orow, float* wrow, bool ok) {
      mbar_wait(bar_o, m & 1);  // item m's MMAs are complete (and, transitively, both halves' s_w partial sums are visible)
      tc_fence_after();
      float wsum = 0.f;
      if (half == 0) wsum = (s_w[(m & 1) * 256 + jj] + s_w[(m & 1) * 256 + 128 + jj]) * invH;
      if (a.need_dqkv) {
        // two 32-column loads one after the other: the thread's 72 probability registers stay live across items
        const uint32_t col = half ? TC_COL_O2 : TC_COL_O;
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_row + col, o);
        tc_wait_ld();
        if (ok) store_row_f16(orow, 3 * d, a.splits, o, 32, 1.0f);
        tmem_ld_32x32b_x32(t_row + col + 32, o);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_e);
        if (ok) store_row_f16(orow + 32, 3 * d, a.splits, o, 32, 1.0f);
      } else {
        mbar_arrive(bar_e);
      }
      if (half == 0 && ok) *wrow = wsum;
    };
    int p = 0, ul = 0;
    for (int n = 0; n < N; ++n) {
      if (p == 0) {
        int bh, b, h, mt;
        unit_coord(ul, bh, b, h, mt);
        dbase = a.delta + (size_t(b) * a.H + h) * T;
        rbase = a.r + size_t(b) * T;
      }
      float2 nx0 = make_float2(0.f, 0.f), nx1 = nx0;
      bool have_next = false;
      if (p + 1 < P) {  // {delta, r} of the next label: global loads in flight during the element-wise phase
        dbase += d_stride, rbase += r_stride;
        fetch_dr(dbase, rbase, nx0, nx1);
        have_next = true;
      } else if (n + 1 < N) {  // first label of the next unit (once per unit: the divisions are off the per-item path)
        int bh, b, h, mt;
        unit_coord(ul + 1, bh, b, h, mt);
        fetch_dr(a.delta + (size_t(b) * a.H + h) * T, a.r + size_t(b) * T, nx0, nx1);
        have_next = true;
      }
      if (p == 0) {
        mbar_wait(bar_kv, ul & 1);  // the tile is read with ordinary loads: every thread acquires the TMA writes
#pragma unroll
        for (int cc = 0; cc < T3_NCH; ++cc) {
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
      if (q == 0) tc_trace(a, 1, 5 + 5 * half, n);  // loop top (after the next item's scalar fetch / unit set-up)
      mbar_wait(&bar_s[half], n & 1);
      tc_fence_after();
      if (q == 0) tc_trace(a, 1, 6 + 5 * half, n);  // G^T chunk seen
      const uint32_t dr_addr = order_after_wait(smem_u32(s_dr + (n & 1) * TC_MAX_T));
      float w0 = 0.f, w1 = 0.f;
      {
        const uint32_t t_strip = t_row + uint32_t(TC_COL_S + c0 * 16), dr0 = dr_addr + uint32_t(c0 * 16) * 8;
        if (nch == 17) {  // ViT-L/14 (T = 257): 9 + 8 chunks
          if (half == 0) col_elementwise<9>(t_strip, dr0, acol, 9, clamp_lo, w0, w1);
          else col_elementwise<8>(t_strip, dr0, acol, 8, clamp_lo, w0, w1);
        } else {
          col_elementwise<0>(t_strip, dr0, acol, c1 - c0, clamp_lo, w0, w1);
        }
      }
      s_w[(n & 1) * 256 + half * 128 + jj] = w0 + w1;
      if (have_next) store_dr(n + 1, nx0, nx1);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_p[half]);
      if (q == 0) tc_trace(a, 1, 7 + 5 * half, n);  // chunk packed, arrived
      // chunk-1 owners drain dV(n) now (their next chunk is queued behind dK / dV(n, 1) anyway); chunk-0 owners drain
      // dK(n - 1), which completed during their element-wise phase of item n
      if (half == 1) {
        if (p == 0) set_out(ul);
        epilogue(n, optr, wptr, valid);
        optr += o_stride, wptr += d_stride;
      } else {
        if (n >= 1) {  // the pointers still describe item n - 1
          epilogue(n - 1, optr, wptr, valid);
          optr += o_stride, wptr += d_stride;
        }
        if (p == 0) set_out(ul);
      }
      if (q == 0) tc_trace(a, 1, 8 + 5 * half, n);  // epilogue done
      if (++p == P) p = 0, ++ul;
    }
    if (N > 0 && half == 0) epilogue(N - 1, optr, wptr, valid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sb

using namespace sb;

static long long* g_attn_trace = nullptr;
// Debug aid: device buffer of 2 x 16 x 64 int64 that the NEXT semabs_attn_bwd_tc3 launches fill with clock64 time stamps of
// CTA 0's pipeline events (kernel 0 = row pass, 1 = column pass; event slots: 0-4 control warp, 5-8 / 10-13 first warp of
// chunk 0 / chunk 1 owners, 15 column-pass MMAs complete); null switches it off.  tools/attn_trace.py prints the timeline.
extern "C" int semabs_debug_attn_trace(long long* device_buf) {
  g_attn_trace = device_buf;
  return 0;
}

// Third-generation attention backward: same contract as semabs_attn_bwd_tc2 (include/semabs_b200.h).
extern "C" int semabs_attn_bwd_tc3(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                                   const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart,
                                   void* dqkv16, int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits,
                                   int32_t positive_only, int32_t need_dqkv, void* stream) {
  // one label per unit: the row pass's K / V look-ahead (next unit loaded one unit ahead into the other buffer) would have to
  // be issued before the current unit has released that buffer — the second generation handles the (rare: ragged label
  // chunks) case
  if (P == 1)
    return semabs_attn_bwd_tc2(qkv16, ld_qkv, probs16, ld_p16, o32, dO16, ld_do, r, delta_ws, wpart, dqkv16, P, B, T, H, splits,
                               positive_only, need_dqkv, stream);
  SB_REQUIRE(qkv16 && probs16 && o32 && dO16 && r && delta_ws && wpart, "semabs_attn_bwd_tc3: null pointer");
  SB_REQUIRE(!need_dqkv || dqkv16, "semabs_attn_bwd_tc3: dqkv16 missing");
  SB_REQUIRE(P > 0 && B > 0 && T > 0 && H > 0 && T <= TC_MAX_T, "semabs_attn_bwd_tc3: bad shape (T <= %d)", TC_MAX_T);
  SB_REQUIRE(ld_p16 % 16 == 0 && ld_p16 >= ((T + 15) / 16) * 16, "semabs_attn_bwd_tc3: bad probs16 pitch %d", ld_p16);
  const int d = H * TC_HD;
  SB_REQUIRE(ld_qkv >= 3 * d && ld_qkv % 8 == 0 && ld_do >= d && ld_do % 16 == 0, "semabs_attn_bwd_tc3: bad pitch");
  SB_REQUIRE(splits == 1 || splits == 2, "semabs_attn_bwd_tc3: splits must be 1 or 2");
  const int rem = T % 128;
  SB_REQUIRE(T <= 128 || rem <= 1, "semabs_attn_bwd_tc3: T %% 128 must be 0 or 1 above one tile (got T=%d)", T);
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
  a.trace = g_attn_trace;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_row_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Row3Smem::TOTAL));
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_col_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ColSmem::TOTAL));
    configured = true;
  }
  const int n_units = B * H * a.n_full;
  const int grid = n_units < num_sms() ? n_units : num_sms();
  if (need_dqkv) {  // delta feeds dS in both passes and the tail; the relevance-only last step needs neither delta nor dQ
    if (int rc = launch_attn_delta(a, st)) return rc;
    attn_bwd_row_tc3_kernel<<<grid, T3_THREADS, Row3Smem::TOTAL, st>>>(tm_qkv, tm_do, a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  attn_bwd_col_tc3_kernel<<<grid, T3_THREADS, ColSmem::TOTAL, st>>>(tm_qkv, tm_do, tm_pr, a);
  SB_CHECK_CUDA(cudaGetLastError());
  if (a.n_tail)
    if (int rc = launch_attn_tail2(a, st)) return rc;
  return 0;
}
