// Shared declarations of the tcgen05 attention kernels (vit_attn_tc.cu: forward + first backward; vit_attn_bwd2.cu: the
// pipelined backward): TMEM column map, shared-memory descriptors of the two operand forms, argument blocks, smem layouts.
#pragma once
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

constexpr int TC_TAIL_MAX = 8;

struct AttnBwdTcArgs {
  const __half* qkv16;    // [B*T, ldq]
  int ldq;
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
  int n_full;             // 128-row MMA tiles per (sequence, head)
  int n_tail;             // rows / keys left to the SIMT tail kernel
  long long* trace;       // debug: clock64 time stamps of CTA 0's pipeline events (null in production), see semabs_debug_attn_trace
};
constexpr int TC_TRACE_ITEMS = 64;   // items of CTA 0 recorded
constexpr int TC_TRACE_EVENTS = 16;  // event slots per kernel
__device__ __forceinline__ void tc_trace(const AttnBwdTcArgs& a, int kernel, int event, int n) {
  if (a.trace && blockIdx.x == 0 && n < TC_TRACE_ITEMS && (threadIdx.x & 31) == 0)
    a.trace[(kernel * TC_TRACE_EVENTS + event) * TC_TRACE_ITEMS + n] = clock64();
}

__device__ __forceinline__ void store_row_f16(__half* dst, int lo_off, int splits, const uint32_t* o, int n, float scale) {
  // n fp32 values (multiple of 16) -> fp16 hi (and lo at +lo_off elements)
  if (splits != 2) {  // single-precision rows: no residual to form
#pragma unroll
    for (int e0 = 0; e0 < n; e0 += 16) {
      uint32_t hi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) hi[e] = pack_h2(__uint_as_float(o[e0 + 2 * e]) * scale, __uint_as_float(o[e0 + 2 * e + 1]) * scale);
      st_global_256(dst + e0, hi);
    }
    return;
  }
#pragma unroll
  for (int e0 = 0; e0 < n; e0 += 16) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
      split_pack(__uint_as_float(o[e0 + 2 * e]) * scale, __uint_as_float(o[e0 + 2 * e + 1]) * scale, hi[e], lo[e]);
    st_global_256(dst + e0, hi);
    if (splits == 2) st_global_256(dst + lo_off + e0, lo);
  }
}

struct RowSmem {
  static constexpr int K = 0;
  static constexpr int V = K + TC_KV_BYTES;
  static constexpr int DO = V + TC_KV_BYTES;            // 2 stages
  static constexpr int DPART = DO + 2 * TC_BOX_BYTES;   // float [2][128] delta partial sums
  static constexpr int BARS = DPART + 1024;
  static constexpr int TOTAL = BARS + 128 + 1024;
};

struct ColSmem {
  static constexpr int Q = 0;                           // all query rows: B of dK (MN-major)
  static constexpr int V = Q + TC_KV_BYTES;             // key tile (A operand of G^T)
  static constexpr int PR = V + TC_BOX_BYTES;           // probabilities [272 i][128 j]: 2 column blocks x 2 row boxes
  static constexpr int DO = PR + 4 * TC_BOX_BYTES;      // 2 stages of all query rows: B of G^T (K-major) / dV (MN-major)
  static constexpr int DR = DO + 2 * TC_KV_BYTES;       // float2 {delta_i, r_i} [2][272]
  static constexpr int W = DR + 2 * TC_MAX_T * 8;       // float [2 buffers][2 halves][128] relevance partial sums
  static constexpr int BARS = W + 2048;
  static constexpr int TOTAL = BARS + 128 + 1024;
};

static inline int make_tile_tmap(CUtensorMap* tm, const void* base, long long rows, long long cols, long long pitch) {
  uint64_t dims[2] = {uint64_t(cols), uint64_t(rows)};
  uint64_t str[1] = {uint64_t(pitch) * 2};
  uint32_t box[2] = {TC_HD, TC_BOX_ROWS};
  return make_tmap_f16(tm, base, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static inline int make_qkv_tmap(CUtensorMap* tm, const void* base, long long rows, long long cols) {
  uint64_t dims[2] = {uint64_t(cols), uint64_t(rows)};
  uint64_t str[1] = {uint64_t(cols) * 2};
  uint32_t box[2] = {TC_HD, TC_BOX_ROWS};
  return make_tmap_f16(tm, base, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sb
