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
#include "vit_attn_tc.cuh"

namespace sb {

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
  const __half* qkv16;  // the same matrix the tensor map describes: the SIMT tail rows read their query row directly
  int ldq;
  long long* trace;     // debug (semabs_debug_attn_fwd_trace): clock64 stamps of one late CTA; null in production
};
__device__ __forceinline__ void fwd_trace(const AttnFwdTcArgs& a, int slot, int lane) {
  if (a.trace && blockIdx.x == gridDim.x / 2 && lane == 0) a.trace[slot] = clock64();
}

struct TcSmem {
  static constexpr int Q_HI = 0;
  static constexpr int Q_LO = Q_HI + TC_BOX_BYTES;
  static constexpr int K_HI = Q_LO + TC_BOX_BYTES;
  static constexpr int K_LO = K_HI + TC_KV_BYTES;
  static constexpr int V_HI = K_LO + TC_KV_BYTES;
  static constexpr int V_LO = V_HI + TC_KV_BYTES;
  static constexpr int BARS = V_LO + TC_KV_BYTES;
  static constexpr int TAIL = BARS + 128;  // float [8 + 3 x 64]: maxima, sums and partial outputs of the three tail warps
  static constexpr int TOTAL = TAIL + 1024 + 1024;
};

// The query row past the last full 128-row tile (T = 257 = 2 x 128 + 1: as a third MMA tile the class-token-plus-256-patch
// sequence's last row cost as much as 128 rows, a third of the kernel) on the three warps that were idle anyway: warp w takes
// the keys [96 w, 96 w + 96), the three combine maximum, sum and partial output through shared memory.  K / V are read from the
// TMA-written tiles (128-byte swizzle: the 16-byte chunk c of row r sits at chunk c ^ (r & 7); two boxes of 136 rows) with
// explicit shared-space loads (the generic form ran ~10x slower here).  Timeline: semabs_debug_attn_fwd_trace.
__device__ __forceinline__ uint32_t tc_kv_row(uint32_t base, int j) {
  return base + uint32_t((j / TC_BOX_ROWS) * TC_BOX_BYTES + (j % TC_BOX_ROWS) * 128);
}
__device__ __forceinline__ uint4 lds128u(uint32_t saddr) {
  uint4 v;
  asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void tail_warps_sync() { asm volatile("bar.sync 1, 96;" ::: "memory"); }
// 8 fp16 hi (+ lo) values of one swizzled 16-byte chunk -> fp32
__device__ __forceinline__ void tc_chunk_f32(uint32_t hi_row, uint32_t lo_row, int pos, bool split, float (&x)[8]) {
  const uint4 u = lds128u(hi_row + (pos << 4));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[t]));
    x[2 * t] = f.x, x[2 * t + 1] = f.y;
  }
  if (split) {
    const uint4 ul = lds128u(lo_row + (pos << 4));
    const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wl[t]));
      x[2 * t] += f.x, x[2 * t + 1] += f.y;
    }
  }
}
constexpr int TC_TAIL_GROUPS = 3;  // 32-key groups per tail warp: 3 warps x 3 x 32 = 288 >= TC_MAX_T
__device__ __noinline__ void attn_fwd_tail_row(uint32_t smem, float* scratch, const AttnFwdTcArgs& a, int bh, int row0, int h, int i,
                                               int w, int lane, int k_hi, int k_lo, int v_hi, int v_lo) {
  const int T = a.T, d = a.d;
  const bool split = a.in_splits == 2;
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr int G = TC_TAIL_GROUPS;
  const __half* qrow = a.qkv16 + size_t(row0 + i) * a.ldq + h * TC_HD;
  float q[TC_HD];
#pragma unroll
  for (int e = 0; e < TC_HD / 2; ++e) {
    float2 x = __half22float2(reinterpret_cast<const __half2*>(qrow)[e]);
    if (split) {
      const float2 y = __half22float2(reinterpret_cast<const __half2*>(qrow + 3 * d)[e]);
      x.x += y.x, x.y += y.y;
    }
    q[2 * e] = x.x, q[2 * e + 1] = x.y;
  }
  const int jmax = a.causal ? (i + 1 < T ? i + 1 : T) : T;
  float p[G];
  float mx = -INFINITY;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int j = (w * G + g) * 32 + lane;
    p[g] = -INFINITY;
    if (j < jmax) {
      const uint32_t kh = tc_kv_row(smem + k_hi, j), kl = tc_kv_row(smem + k_lo, j);
      const int sw = (j % TC_BOX_ROWS) & 7;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};  // four chains: a single one is 64 dependent FMAs per key
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float x[8];
        tc_chunk_f32(kh, kl, ch ^ sw, split, x);
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] = fmaf(q[8 * ch + 2 * t], x[2 * t], acc[t]), acc[t] = fmaf(q[8 * ch + 2 * t + 1], x[2 * t + 1], acc[t]);
      }
      p[g] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      mx = fmaxf(mx, p[g]);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) scratch[w] = mx;
  tail_warps_sync();
  mx = fmaxf(fmaxf(scratch[0], scratch[1]), scratch[2]);
  const float mb = mx * LOG2E;
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    p[g] = ((w * G + g) * 32 + lane < jmax) ? fast_exp2(fmaf(p[g], LOG2E, -mb)) : 0.f;
    sum += p[g];
  }
  sum = warp_sum(sum);
  if (lane == 0) scratch[4 + w] = sum;
  tail_warps_sync();
  const float inv = 1.0f / ((scratch[4] + scratch[5]) + scratch[6]);
  __half* prow = a.probs16 ? a.probs16 + (size_t(bh) * T + i) * a.ldp : nullptr;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    p[g] *= inv;
    const int j = (w * G + g) * 32 + lane;
    if (prow && j < a.ldp) prow[j] = __float2half_rn(p[g]);
    if (!split) p[g] = __half2float(__float2half_rn(p[g]));  // the MMA path feeds P V with fp16 probabilities when the operands are single
  }
  // partial O_i = sum over this warp's keys of p_j V_j: every lane sums its own keys over all 64 channels (64 independent
  // chains, conflict-free 16-byte row reads), then a reduce-scatter over the warp leaves channels 2 lane, 2 lane + 1 in `lane`
  float v[TC_HD];
#pragma unroll
  for (int e = 0; e < TC_HD; ++e) v[e] = 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int j = (w * G + g) * 32 + lane;
    if (j < jmax) {
      const uint32_t vh = tc_kv_row(smem + v_hi, j), vl = tc_kv_row(smem + v_lo, j);
      const int sw = (j % TC_BOX_ROWS) & 7;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float x[8];
        tc_chunk_f32(vh, vl, ch ^ sw, split, x);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[8 * ch + t] = fmaf(p[g], x[t], v[8 * ch + t]);
      }
    }
  }
#pragma unroll
  for (int m = 16, n = TC_HD; m >= 1; m >>= 1, n >>= 1) {
    const bool up = (lane & m) != 0;  // this lane keeps the upper half of its n values, its partner the lower half
#pragma unroll
    for (int e = 0; e < n / 2; ++e) {
      const float send = up ? v[e] : v[e + n / 2];
      const float keep = up ? v[e + n / 2] : v[e];
      v[e] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  *reinterpret_cast<float2*>(scratch + 8 + w * TC_HD + 2 * lane) = make_float2(v[0], v[1]);
  tail_warps_sync();
  if (w != 0) return;
  const float2 p0 = *reinterpret_cast<const float2*>(scratch + 8 + 2 * lane);
  const float2 p1 = *reinterpret_cast<const float2*>(scratch + 8 + TC_HD + 2 * lane);
  const float2 p2 = *reinterpret_cast<const float2*>(scratch + 8 + 2 * TC_HD + 2 * lane);
  const float o0 = (p0.x + p1.x) + p2.x, o1 = (p0.y + p1.y) + p2.y;
  const size_t row = size_t(row0) + i;
  if (a.o32) *reinterpret_cast<float2*>(a.o32 + row * d + h * TC_HD + 2 * lane) = make_float2(o0, o1);
  if (a.o16) {
    uint32_t hi, lo;
    split_pack(o0, o1, hi, lo);
    __half* dst = a.o16 + row * size_t(a.o_splits) * d + h * TC_HD + 2 * lane;
    *reinterpret_cast<uint32_t*>(dst) = hi;
    if (a.o_splits == 2) *reinterpret_cast<uint32_t*>(dst + d) = lo;
  }
}

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
  // a single row beyond the last full 128-row tile goes to the otherwise idle warps 1..3
  const int n_tail = (T > 128 && (T & 127) == 1) ? 1 : 0;
  const int n_mt = n_tail ? T / 128 : (T + 127) / 128;
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
  if (warp == 4) fwd_trace(a, 0, lane);
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
  } else if (warp <= 3) {
    if (n_tail) {
      fwd_trace(a, 32 + 4 * (warp - 1), lane);
      mbar_wait(bar_kv, 0);  // K / V of the whole sequence are in shared memory
      fwd_trace(a, 33 + 4 * (warp - 1), lane);
      attn_fwd_tail_row(smem_u32(smem), reinterpret_cast<float*>(smem + TcSmem::TAIL), a, bh, row0, h, n_mt * 128, warp - 1, lane,
                        TcSmem::K_HI, TcSmem::K_LO, TcSmem::V_HI, TcSmem::V_LO);
      fwd_trace(a, 34 + 4 * (warp - 1), lane);
    }
  } else {
    // ===== softmax + epilogue: thread = query row =====
    const int q = warp & 3;
    const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    for (int mt = 0; mt < n_mt; ++mt) {
      const int i = mt * 128 + q * 32 + lane;
      const bool valid = i < T;
      const int jmax = a.causal ? (i < T ? i + 1 : T) : T;  // keys [0, jmax) take part
      if (warp == 4) fwd_trace(a, 1 + 8 * mt, lane);
      mbar_wait(bar_s, mt & 1);
      tc_fence_after();
      if (warp == 4) fwd_trace(a, 2 + 8 * mt, lane);
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
      if (warp == 4) fwd_trace(a, 3 + 8 * mt, lane);
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
      if (warp == 4) fwd_trace(a, 4 + 8 * mt, lane);
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
      if (warp == 4) fwd_trace(a, 5 + 8 * mt, lane);
      // epilogue
      mbar_wait(bar_o, mt & 1);
      tc_fence_after();
      if (warp == 4) fwd_trace(a, 6 + 8 * mt, lane);
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
      if (warp == 4) fwd_trace(a, 7 + 8 * mt, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
    fwd_trace(a, 31, lane);
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward + relevance (ClipGradcam.interpret, CLIP/clip/clip_gradcam.py:90-126; the autograd graph of
// auxiliary.multi_head_attention_forward):
//   G = dO V^T ; delta_i = dO_i . O_i ; dS = A o (G - delta) ; w_j = (1/H) sum_i r_i relu?(G o A)_ij
//   dQ = scale dS K ; dK = dS^T Q ; dV = A^T dO
// Row pass (thread = query row i):  MMA G[128 x 272] -> dS packed in place -> TS MMA dQ = dS K (K as MN-major B).
// Column pass (thread = key j):     MMA G^T = V dO^T -> relevance column sums are thread-local -> dS^T and A^T packed
//                                   in place -> TS MMAs dK = dS^T Q, dV = A^T dO (Q, dO as MN-major B).
// Both are persistent: a CTA owns (sequence, head, 128-row tile) units and walks the P label cotangents of each unit,
// so K / V / Q / the probability tile (and, in the row pass, each thread's probability row, in registers) are fetched
// once per unit while the per-label dO tiles stream through a 2-stage TMA ring.  All 8 warps share the elementwise
// phase (warp w and w+4 split the strip columns of TMEM lane quadrant w%4); warp 0 also issues TMA and MMAs.
// T = 128 k + t with t <= 8 (ViT-L/14: 257 = 2*128 + 1) leaves t rows / keys to attn_bwd_tail_kernel (SIMT) instead of
// paying a whole 128-row tile for them.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_row_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RowSmem::BARS);
  uint64_t *bar_kv = bars, *bar_do = bars + 1 /* [2] */, *bar_s = bars + 3, *bar_p = bars + 4, *bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* s_dpart = reinterpret_cast<float*>(smem + RowSmem::DPART);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2, rr = q * 32 + lane;
  const int T = a.T, d = a.d, P = a.P;
  const int ncol = (T + 15) & ~15;
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;
  const int nch = ncol / 16, c_split = (nch + 1) / 2;
  const int c0 = half ? c_split : 0, c1 = half ? nch : c_split;
  const int n_units = a.B * a.H * a.n_full;
  const int n_my = (n_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int N = n_my * P;  // items of this CTA: unit-major, label fastest

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_kv, 1), mbar_init(&bar_do[0], 1), mbar_init(&bar_do[1], 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, TC_THREADS);
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
  const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
  const uint32_t tS = tmem_base + TC_COL_S, tO = tmem_base + TC_COL_O;

  auto item = [&](int n, int& bh, int& mt, int& p) {
    const int u = int(blockIdx.x) + (n / P) * int(gridDim.x);
    p = n % P, bh = u / a.n_full, mt = u % a.n_full;
  };
  uint32_t leader = 0;
  if (warp == 0) leader = elect_one() ? 1u : 0u;
  const uint32_t sbase = smem_u32(smem);
  const uint64_t dV = desc_kmajor(sbase + RowSmem::V), dK = desc_mnmajor(sbase + RowSmem::K, 16);
  const uint32_t idesc_s1 = make_idesc_f16(128, n1), idesc_s2 = make_idesc_f16(128, n2 ? n2 : 16);
  constexpr uint32_t idesc_o = make_idesc_f16(128, TC_HD, false, true);

  auto load_do = [&](int n) {  // elected lane of warp 0
    int bh, mt, p;
    item(n, bh, mt, p);
    const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
    mbar_arrive_expect_tx(&bar_do[n & 1], TC_BOX_BYTES);
    tma_load_2d(smem + RowSmem::DO + (n & 1) * TC_BOX_BYTES, &tm_do, &bar_do[n & 1], h * TC_HD, pb * T + mt * 128);
  };
  auto load_kv = [&](int n) {  // elected lane of warp 0; n = first item of a unit
    int bh, mt, p;
    item(n, bh, mt, p);
    const int b = bh / a.H, h = bh % a.H;
    mbar_arrive_expect_tx(bar_kv, 2u * TC_KV_BYTES);
    for (int bx = 0; bx < 2; ++bx) {
      tma_load_2d(smem + RowSmem::V + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, 2 * d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
      tma_load_2d(smem + RowSmem::K + bx * TC_BOX_BYTES, &tm_qkv, bar_kv, d + h * TC_HD, b * T + bx * TC_BOX_ROWS);
    }
  };
  auto issue_g = [&](int n) {  // warp 0, convergent
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

  if (warp == 0 && N > 0) {
    if (leader) {
      load_kv(0);
      load_do(0);
      if (N > 1) load_do(1);
    }
    __syncwarp();
    issue_g(0);
  }

  uint32_t arow[(TC_MAX_T / 16 + 1) / 2][8];  // this thread's half of its probability row (kept across the P labels)
  for (int n = 0; n < N; ++n) {
    int bh, mt, p;
    item(n, bh, mt, p);
    const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
    const int i = mt * 128 + rr;
    const bool valid = i < T;
    if (p == 0) {
      const __half* prow = a.probs16 + (size_t(bh) * T + (valid ? i : 0)) * a.ldp;
#pragma unroll
      for (int cc = 0; cc < (TC_MAX_T / 16 + 1) / 2; ++cc) {
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
    // delta_i = dO_i . O_i : each half sums 32 of the 64 head channels
    {
      float part = 0.f;
      if (valid) {
        const float* orow = a.o32 + (size_t(b) * T + i) * d + h * TC_HD + 32 * half;
        const __half* grow = a.dO16 + (size_t(pb) * T + i) * a.ld_do + h * TC_HD + 32 * half;
        uint32_t ov[32], gv[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) ld_global_256(orow + 8 * e, ov + 8 * e);
        ld_global_256(grow, gv);
        ld_global_256(grow + 16, gv + 8);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&gv[k]));
          part = fmaf(g2.x, __uint_as_float(ov[2 * k]), part);
          part = fmaf(g2.y, __uint_as_float(ov[2 * k + 1]), part);
        }
      }
      s_dpart[half * 128 + rr] = part;
      // rows left to the tail kernel: their delta is needed by the column pass too
      if (mt == 0 && int(threadIdx.x) < a.n_tail) {
        const int it = a.n_full * 128 + int(threadIdx.x);
        const float* orow = a.o32 + (size_t(b) * T + it) * d + h * TC_HD;
        const __half2* grow = reinterpret_cast<const __half2*>(a.dO16 + (size_t(pb) * T + it) * a.ld_do + h * TC_HD);
        float dl = 0.f;
        for (int k = 0; k < 32; ++k) {
          const float2 g2 = __half22float2(grow[k]);
          dl = fmaf(g2.x, orow[2 * k], dl), dl = fmaf(g2.y, orow[2 * k + 1], dl);
        }
        a.delta[(size_t(pb) * a.H + h) * T + it] = dl;
      }
    }
    __syncthreads();
    const float delta = s_dpart[rr] + s_dpart[128 + rr];
    if (half == 0 && valid) a.delta[(size_t(pb) * a.H + h) * T + i] = delta;

    mbar_wait(bar_s, n & 1);
    tc_fence_after();
    if (leader && n + 2 < N) load_do(n + 2);  // G(n) has consumed stage n&1
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < (TC_MAX_T / 16 + 1) / 2; ++cc) {
      if (c0 + cc < c1) {
        uint32_t g[16], w[8];
        tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + (c0 + cc) * 16), g);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&arow[cc][e]));
          w[e] = pack_h2(a2.x * (__uint_as_float(g[2 * e]) - delta), a2.y * (__uint_as_float(g[2 * e + 1]) - delta));
        }
        tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + (c0 + cc) * 16), w);
      }
    }
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(bar_p);
    if (warp == 0) {
      mbar_wait(bar_p, n & 1);
      tc_fence_after();
      for (int s = 0; s < nch; ++s)
        umma_f16_ts_elect(tO, tS + uint32_t(16 * s), dK + uint64_t(s) * (2048 >> 4), idesc_o, s > 0, leader);
      umma_commit_elect(bar_o, leader);
    }
    mbar_wait(bar_o, n & 1);
    tc_fence_after();
    if (warp == 0 && n + 1 < N) {
      if ((n + 1) % P == 0 && leader) load_kv(n + 1);  // dQ(n) was the last reader of this unit's K / V
      __syncwarp();
      issue_g(n + 1);  // overlaps the epilogue below (different TMEM columns)
    }
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_row + uint32_t(TC_COL_O + 32 * half), o);
      tc_wait_ld();
      if (valid)
        store_row_f16(a.dqkv16 + (size_t(pb) * T + i) * size_t(a.splits) * 3 * d + h * TC_HD + 32 * half, 3 * d, a.splits, o, 32, a.scale);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_col_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_pr, AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ColSmem::BARS);
  uint64_t *bar_kv = bars, *bar_do = bars + 1 /* [2] */, *bar_s = bars + 3, *bar_p = bars + 4, *bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float2* s_dr = reinterpret_cast<float2*>(smem + ColSmem::DR);
  float* s_w = reinterpret_cast<float*>(smem + ColSmem::W);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2, jj = q * 32 + lane;
  const int T = a.T, d = a.d, P = a.P;
  const int ncol = (T + 15) & ~15;  // query columns of G^T
  const int n1 = ncol < 256 ? ncol : 256, n2 = ncol - n1;
  const int nch = ncol / 16, c_split = (nch + 1) / 2;
  const int c0 = half ? c_split : 0, c1 = half ? nch : c_split;
  const int n_units = a.B * a.H * a.n_full;
  const int n_my = (n_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int N = n_my * P;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_pr);
    mbar_init(bar_kv, 1), mbar_init(&bar_do[0], 1), mbar_init(&bar_do[1], 1), mbar_init(bar_s, 1), mbar_init(bar_o, 1);
    mbar_init(bar_p, TC_THREADS);
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
  auto load_dr = [&](int n) {  // all threads: {delta_i, r_i} of item n into buffer n&1
    int bh, mt, p;
    item(n, bh, mt, p);
    const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
    for (int i = threadIdx.x; i < TC_MAX_T; i += TC_THREADS) {
      float2 v = make_float2(0.f, 0.f);
      if (i < T) {
        v.x = a.need_dqkv ? a.delta[(size_t(pb) * a.H + h) * T + i] : 0.f;
        v.y = a.r[size_t(pb) * T + i];
      }
      s_dr[(n & 1) * TC_MAX_T + i] = v;
    }
  };
  if (N > 0) load_dr(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16);
  const uint32_t tS = tmem_base + TC_COL_S, tK = tmem_base + TC_COL_O, tV = tmem_base + TC_COL_O2;

  uint32_t leader = 0;
  if (warp == 0) leader = elect_one() ? 1u : 0u;
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

  if (warp == 0 && N > 0) {
    if (leader) {
      load_unit(0);
      load_do(0);
      if (N > 1) load_do(1);
    }
    __syncwarp();
    issue_g(0);
  }

  // probability tile addressing: element (i, jj) sits at  i*128 + (((jj%64)/8 ^ (i%8)) << 4) + (jj%8)*2  inside column block
  // jj/64 (the two 136-row TMA boxes of a column block are contiguous and 136 % 8 == 0, so i needs no box arithmetic)
  const uint8_t* pcol = smem + ColSmem::PR + (jj >> 6) * (2 * TC_BOX_BYTES) + (jj & 7) * 2;
  uint32_t poff[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) poff[k] = uint32_t(k * 128 + (((((jj & 63) >> 3) ^ k)) << 4));
  const float invH = 1.0f / a.H;

  for (int n = 0; n < N; ++n) {
    int bh, mt, p;
    item(n, bh, mt, p);
    const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
    const int j = mt * 128 + jj;
    const bool valid = j < T;
    if (n + 1 < N) load_dr(n + 1);
    if (p == 0) mbar_wait(bar_kv, (n / P) & 1);  // the probability tile is read with ordinary loads: every thread acquires it
    mbar_wait(bar_s, n & 1);
    tc_fence_after();
    const float2* dr = s_dr + (n & 1) * TC_MAX_T;
    float w = 0.f;
    for (int c = c0; c < c1; ++c) {
      uint32_t g[16], ds[8], aw[8];
      tmem_ld_32x32b_x16(t_row + uint32_t(TC_COL_S + c * 16), g);
      const uint8_t* prow = pcol + c * 2048;
      const float4* dr4 = reinterpret_cast<const float4*>(dr + c * 16);
      const bool full = c * 16 + 16 <= T;
      uint32_t raw[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        raw[e] = *reinterpret_cast<const uint16_t*>(prow + (e >> 3) * 1024 + poff[e & 7]);
        if (!full && c * 16 + e >= T) raw[e] = 0u;  // rows past T belong to the next head
      }
      tc_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        const uint32_t word = raw[e] | (raw[e + 1] << 16);
        const float2 av = __half22float2(*reinterpret_cast<const __half2*>(&word));
        const float4 d4 = dr4[e >> 1];  // {delta_i, r_i, delta_i+1, r_i+1}
        const float g0 = __uint_as_float(g[e]), g1 = __uint_as_float(g[e + 1]);
        float x0 = g0 * av.x, x1 = g1 * av.y;
        if (a.positive_only) x0 = fmaxf(x0, 0.f), x1 = fmaxf(x1, 0.f);
        w = fmaf(d4.y, x0, w);
        w = fmaf(d4.w, x1, w);
        ds[e >> 1] = pack_h2(av.x * (g0 - d4.x), av.y * (g1 - d4.z));
        aw[e >> 1] = word;
      }
      tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c * 16), ds);
      tmem_st_32x32b_x8(t_row + uint32_t(TC_COL_S + c * 16 + 8), aw);
    }
    s_w[(n & 1) * 256 + half * 128 + jj] = w;
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_p, n & 1);
    if (warp == 0) {
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
    }
    mbar_wait(bar_o, n & 1);
    tc_fence_after();
    if (warp == 0) {
      if (leader && n + 2 < N) load_do(n + 2);  // stage n&1 was last read by dV(n)
      if (n + 1 < N) {
        if ((n + 1) % P == 0 && leader) load_unit(n + 1);
        __syncwarp();
        issue_g(n + 1);
      }
    }
    if (half == 0 && valid)
      a.wpart[(size_t(pb) * a.H + h) * T + j] = (s_w[(n & 1) * 256 + jj] + s_w[(n & 1) * 256 + 128 + jj]) * invH;
    if (a.need_dqkv) {
      uint32_t o[64];
      const uint32_t col = half ? TC_COL_O2 : TC_COL_O;
      tmem_ld_32x32b_x32(t_row + col, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
      tmem_ld_32x32b_x32(t_row + col + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
      tc_wait_ld();
      if (valid)
        store_row_f16(a.dqkv16 + (size_t(pb) * T + j) * size_t(a.splits) * 3 * d + (half ? 2 * d : d) + h * TC_HD, 3 * d, a.splits,
                      o, 64, 1.0f);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------
// SIMT tail: the n_tail (<= 8) query rows / keys past the last full 128-row tile.  One warp per (label, sequence, head,
// tail index): the column part (key j0: relevance, dK_j0, dV_j0) runs with lanes over queries, the row part (query i0:
// dQ_i0) with lanes over keys.  delta was written by the row pass.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_row64(const __half* p, __half2 (&v)[32]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + 8 * e * 2);
    *reinterpret_cast<uint4*>(&v[8 * e]) = u;
    const uint4 u2 = *reinterpret_cast<const uint4*>(p + 8 * e * 2 + 8);
    *reinterpret_cast<uint4*>(&v[8 * e + 4]) = u2;
  }
}
__device__ __forceinline__ float dot64(const __half2 (&x)[32], const __half2 (&y)[32]) {
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const float2 fx = __half22float2(x[e]), fy = __half22float2(y[e]);
    acc = fmaf(fx.x, fy.x, acc), acc = fmaf(fx.y, fy.y, acc);
  }
  return acc;
}
// sums acc[0..63] over the warp; lane l returns elements 2l, 2l+1
__device__ __forceinline__ float2 warp_reduce64(float (&acc)[64], int lane) {
  float2 mine = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 64; ++e) {
    const float s = warp_sum(acc[e]);
    if ((e >> 1) == lane) {
      if (e & 1) mine.y = s; else mine.x = s;
    }
  }
  return mine;
}
__device__ __forceinline__ void store_pair_split(__half* row, int col, int lo_off, int splits, float x, float y) {
  uint32_t hi, lo;
  split_pack(x, y, hi, lo);
  *reinterpret_cast<uint32_t*>(row + col) = hi;
  if (splits == 2) *reinterpret_cast<uint32_t*>(row + lo_off + col) = lo;
}

__global__ void __launch_bounds__(128) attn_bwd_tail_kernel(AttnBwdTcArgs a) {
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= a.P * a.B * a.H * a.n_tail) return;
  const int t = gw % a.n_tail, rest = gw / a.n_tail, p = rest % a.P, bh = rest / a.P;
  const int b = bh / a.H, h = bh % a.H, pb = p * a.B + b;
  const int T = a.T, d = a.d;
  const int x0 = a.n_full * 128 + t;  // the tail row / key
  const __half* qkv = a.qkv16 + size_t(b) * T * a.ldq + h * TC_HD;
  const __half* dOb = a.dO16 + size_t(pb) * T * a.ld_do + h * TC_HD;
  const __half* Ab = a.probs16 + size_t(bh) * T * a.ldp;
  const float* dl = a.delta + (size_t(pb) * a.H + h) * T;
  const size_t ld = size_t(a.splits) * 3 * d;
  constexpr int JMAX = (TC_MAX_T + 31) / 32;
  // Two phases per part: (A) lanes over queries / keys — one 64-long dot product per lane gives the per-row scalars
  // (relevance term, dS, A); (B) lanes over the 64 head channels (one half2 each) — the weighted row sums dK, dV, dQ are
  // accumulated with the scalars broadcast by shuffles, rows read as whole 128-byte segments.  (The first version kept
  // 3 x 64 accumulators per lane and reduced them with 3 x 64 warp sums: 950 us per launch at 9 % occupancy.)
  // ---- column part: key x0
  {
    __half2 v0[32];
    load_row64(qkv + size_t(x0) * a.ldq + 2 * d, v0);
    float dsv[JMAX], avv[JMAX];
    float wsum = 0.f;
#pragma unroll
    for (int c = 0; c < JMAX; ++c) {
      const int i = c * 32 + lane;
      dsv[c] = 0.f, avv[c] = 0.f;
      if (i < T) {
        __half2 gi[32];
        load_row64(dOb + size_t(i) * a.ld_do, gi);
        const float g = dot64(gi, v0);
        const float av = __half2float(Ab[size_t(i) * a.ldp + x0]);
        float x = g * av;
        if (a.positive_only) x = fmaxf(x, 0.f);
        wsum = fmaf(a.r[size_t(pb) * T + i], x, wsum);
        dsv[c] = av * (g - dl[i]), avv[c] = av;
      }
    }
    wsum = warp_sum(wsum);
    if (lane == 0) a.wpart[(size_t(pb) * a.H + h) * T + x0] = wsum / a.H;
    if (a.need_dqkv) {
      float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
#pragma unroll
      for (int c = 0; c < JMAX; ++c) {
        const int n = min(32, T - c * 32);  // warp-uniform
        for (int ii = 0; ii < n; ++ii) {
          const int i = c * 32 + ii;
          const float ds = __shfl_sync(0xffffffffu, dsv[c], ii), av = __shfl_sync(0xffffffffu, avv[c], ii);
          const float2 fq = __half22float2(reinterpret_cast<const __half2*>(qkv + size_t(i) * a.ldq)[lane]);
          const float2 fg = __half22float2(reinterpret_cast<const __half2*>(dOb + size_t(i) * a.ld_do)[lane]);
          dk0 = fmaf(ds, fq.x, dk0), dk1 = fmaf(ds, fq.y, dk1);
          dv0 = fmaf(av, fg.x, dv0), dv1 = fmaf(av, fg.y, dv1);
        }
      }
      __half* orow = a.dqkv16 + (size_t(pb) * T + x0) * ld + h * TC_HD;
      store_pair_split(orow, d + 2 * lane, 3 * d, a.splits, dk0, dk1);
      store_pair_split(orow, 2 * d + 2 * lane, 3 * d, a.splits, dv0, dv1);
    }
  }
  // ---- row part: query x0
  if (a.need_dqkv) {
    __half2 g0[32];
    load_row64(dOb + size_t(x0) * a.ld_do, g0);
    const float delta0 = dl[x0];
    float dsv[JMAX];
#pragma unroll
    for (int c = 0; c < JMAX; ++c) {
      const int j = c * 32 + lane;
      dsv[c] = 0.f;
      if (j < T) {
        __half2 vj[32];
        load_row64(qkv + size_t(j) * a.ldq + 2 * d, vj);
        const float g = dot64(g0, vj);
        const float av = __half2float(Ab[size_t(x0) * a.ldp + j]);
        dsv[c] = av * (g - delta0);
      }
    }
    float dq0 = 0.f, dq1 = 0.f;
#pragma unroll
    for (int c = 0; c < JMAX; ++c) {
      const int n = min(32, T - c * 32);
      for (int jj = 0; jj < n; ++jj) {
        const float ds = __shfl_sync(0xffffffffu, dsv[c], jj);
        const float2 fk = __half22float2(reinterpret_cast<const __half2*>(qkv + size_t(c * 32 + jj) * a.ldq + d)[lane]);
        dq0 = fmaf(ds, fk.x, dq0), dq1 = fmaf(ds, fk.y, dq1);
      }
    }
    __half* orow = a.dqkv16 + (size_t(pb) * T + x0) * ld + h * TC_HD;
    store_pair_split(orow, 2 * lane, 3 * d, a.splits, dq0 * a.scale, dq1 * a.scale);
  }
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

static long long* g_attn_fwd_trace = nullptr;
extern "C" int semabs_debug_attn_fwd_trace(long long* device_buf) {
  g_attn_fwd_trace = device_buf;
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
  AttnFwdTcArgs a{(__half*)probs16, ld_p16, o32, (__half*)o16, o_splits, T, H, d, causal, in_splits, (const __half*)qkv16, in_splits * 3 * d, g_attn_fwd_trace};
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
  a.qkv16 = (const __half*)qkv16, a.ldq = ld_qkv;
  a.probs16 = (const __half*)probs16, a.ldp = ld_p16, a.o32 = o32, a.dO16 = (const __half*)dO16, a.ld_do = ld_do;
  a.delta = delta_ws, a.r = r, a.wpart = wpart, a.dqkv16 = (__half*)dqkv16;
  a.P = P, a.B = B, a.T = T, a.H = H, a.d = d, a.splits = splits, a.scale = 0.125f;
  a.positive_only = positive_only, a.need_dqkv = need_dqkv;
  const int rem = T % 128;
  a.n_tail = (T > 128 && rem >= 1 && rem <= TC_TAIL_MAX) ? rem : 0;
  a.n_full = a.n_tail ? T / 128 : (T + 127) / 128;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_row_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RowSmem::TOTAL));
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_col_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ColSmem::TOTAL));
    configured = true;
  }
  const int n_units = B * H * a.n_full;
  const int grid = n_units < num_sms() ? n_units : num_sms();
  if (need_dqkv) {  // the row pass produces dQ and delta; the relevance-only last step needs neither
    attn_bwd_row_tc_kernel<<<grid, TC_THREADS, RowSmem::TOTAL, st>>>(tm_qkv, tm_do, a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  attn_bwd_col_tc_kernel<<<grid, TC_THREADS, ColSmem::TOTAL, st>>>(tm_qkv, tm_do, tm_pr, a);
  SB_CHECK_CUDA(cudaGetLastError());
  if (a.n_tail) {
    const int warps = P * B * H * a.n_tail;
    attn_bwd_tail_kernel<<<(warps + 3) / 4, 128, 0, st>>>(a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
