// Persistent, warp-specialised TN GEMM for sm_100a: TMA (128B-swizzled tiles) -> smem ring -> tcgen05.mma with
// fp32 accumulators in TMEM (double-buffered) -> tcgen05.ld epilogue with fused bias / QuickGELU / residual.
//
//   C[M,N] = epi( A[M, splits*K] * B[N,K]^T )           A, B fp16 K-major
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
// warps 4..11 = epilogue: warp w owns TMEM lanes [32*(w%4), +32) and the column half (w-4)/4 of the tile, in 16-column chunks.
// (Round 1 ran 4 epilogue warps over 32-column chunks: with K = 1024 a 128x256 tile is only 8192 tensor cycles long and the
// one-row-per-thread epilogue — TMEM load, side-input loads, convert, store for 256 columns — did not always fit: ncu showed
// the tensor pipe at 64 % (aux-multiplying dgrad) / 77 % (out-proj dgrad) on those shapes against 84-88 % for K >= 3072.)
// Replaces F.linear call sites of the reference ViT (see include/semabs_b200.h for file:line).
#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 fp16 = 128 bytes = one swizzle-128B row
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_CW = 16;  // epilogue chunk width (columns)

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  // NOTE: the epilogue deliberately does NOT stage through shared memory: with SS-mode 128x128x16 MMAs the tensor
  // core already consumes the full 128 B/clk of shared-memory bandwidth, and an smem-transposed epilogue (tried in
  // round 1: -40 % on K=1024 shapes) steals it. Each thread owns one accumulator row and moves whole 32 B sectors.
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KiB alignment
};

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
__device__ __forceinline__ bool tile_coords(const EpiParams& ep, int t, int num_m, int num_n, int& m_blk, int& n_blk) {
  n_blk = t % num_n;
  const int u = t / num_n;
  if (ep.raster_rows == 0) {
    m_blk = u;
    return true;
  }
  const int p = u % ep.raster_reps, g = u / ep.raster_reps;
  m_blk = g + int((long long)p * ep.raster_rows / 128);
  const int next = (p + 1 == ep.raster_reps) ? num_m : int((long long)(p + 1) * ep.raster_rows / 128);
  return m_blk < next;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                   int kblocks_total, int kblocks_wrap_b, EpiParams ep) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = (ep.raster_rows ? ep.raster_groups * ep.raster_reps : num_m) * num_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        if (!tile_coords(ep, t, num_m, num_n, m_blk, n_blk)) continue;
        for (int kb = 0; kb < kblocks_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(sA, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sB, &tmB, &full_bar[stage], (kb % kblocks_wrap_b) * GEMM_BK, n_blk * BN);
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp runs the loop, the elected lane issues (see umma_f16_elect) =====
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN);
      // descriptor = loop-invariant high word | (smem address >> 4)
      const uint64_t desc0 = make_smem_desc(0, 16, 1024, SW_128B) + (smem_u32(smem) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        {
          int m_blk, n_blk;
          if (!tile_coords(ep, t, num_m, num_n, m_blk, n_blk)) continue;
        }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = desc0 + uint32_t(stage) * uint32_t(S::STAGE_BYTES >> 4);
          const uint64_t db = da + uint32_t(S::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128B swizzle row: +2 in the 16-byte address field
            umma_f16_elect(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0, leader);
          }
          umma_commit_elect(&empty_bar[stage], leader);  // smem slot reusable once these MMAs retire
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_elect(&tmem_full[acc], leader);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3, chalf = (warp - 4) >> 2;
    constexpr int CW = GEMM_CW;
    constexpr int HALF = BN / 2;       // columns per epilogue warp
    constexpr int NC = HALF / CW;      // chunks per warp and tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      if (!tile_coords(ep, t, num_m, num_n, m_blk, n_blk)) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < M;
      const __half* aux_row = nullptr;
      if (ep.act == SEMABS_ACT_MUL_AUX16 && row_ok) aux_row = ep.aux16 + size_t(row % ep.aux_rows) * ep.ld_aux;
      // Software-pipelined over 16-column chunks: the TMEM load and the global side inputs (aux / residual) of chunk
      // c+1 are in flight while chunk c is converted and stored.
      const int colbase = n_blk * BN + chalf * HALF;
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + chalf * HALF);
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
                const float s0 = sigmoidf_precise(1.702f * v[j]), s1 = sigmoidf_precise(1.702f * v[j + 1]);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int kblocks_total,
                       int kblocks_wrap_b, const EpiParams& ep, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int num_tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  gemm_f16_tn_kernel<BN><<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tmA, tmB, M, N, kblocks_total, kblocks_wrap_b, ep);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sb

extern "C" int semabs_gemm_f16(const void* A, int32_t lda, const void* B, int32_t ldb, int32_t M, int32_t N,
                               int32_t K, int32_t a_splits, const semabs_gemm_epilogue* e, void* stream) {
  using namespace sb;
  SB_REQUIRE(A && B && e, "semabs_gemm_f16: null pointer");
  SB_REQUIRE(M > 0 && N > 0 && K > 0, "semabs_gemm_f16: bad shape M=%d N=%d K=%d", M, N, K);
  SB_REQUIRE(a_splits == 1 || a_splits == 2, "semabs_gemm_f16: a_splits must be 1 or 2");
  SB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "semabs_gemm_f16: K/lda/ldb must be multiples of 8");
  SB_REQUIRE(N % 32 == 0, "semabs_gemm_f16: N must be a multiple of 32 (got %d)", N);
  SB_REQUIRE(a_splits == 1 || K % GEMM_BK == 0, "semabs_gemm_f16: split-A needs K %% 64 == 0");
  SB_REQUIRE(e->out_f32 || e->out_f16, "semabs_gemm_f16: no output buffer");
  SB_REQUIRE(e->act != SEMABS_ACT_MUL_AUX16 || (e->aux16 && e->aux_rows > 0), "semabs_gemm_f16: aux16 missing");
  SB_REQUIRE(e->ld_out % 4 == 0 && e->ld_out16 % 8 == 0 && e->ld_aux % 8 == 0 && e->ld_out_aux % 8 == 0,
             "semabs_gemm_f16: bad output pitch");

  const int kblocks = (K + GEMM_BK - 1) / GEMM_BK;
  // 128x256 tiles when the problem is big enough to fill the SMs: one MMA then reads A (4 KB) + B (8 KB) per 128
  // tensor cycles = 96 B/clk of shared-memory bandwidth instead of the 128 B/clk (the full SMEM rate) of 128x128
  const long long tiles256 = (long long)((M + GEMM_BM - 1) / GEMM_BM) * (N / 256);
  const int BN = (N % 256 == 0 && tiles256 >= 2LL * num_sms()) ? 256 : (N % 128 == 0) ? 128 : (N % 64 == 0 ? 64 : 32);

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {uint64_t(a_splits) * uint64_t(K), uint64_t(M)};
    uint64_t str[1] = {uint64_t(lda) * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    if (int rc = make_tmap_f16(&tmA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    uint64_t dims[2] = {uint64_t(K), uint64_t(N)};
    uint64_t str[1] = {uint64_t(ldb) * 2};
    uint32_t box[2] = {GEMM_BK, uint32_t(BN)};
    if (int rc = make_tmap_f16(&tmB, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  EpiParams ep;
  ep.bias = e->bias;
  ep.residual = e->residual;
  ep.aux16 = reinterpret_cast<const __half*>(e->aux16);
  ep.aux_rows = e->aux_rows;
  ep.ld_aux = e->ld_aux;
  ep.out_aux16 = reinterpret_cast<__half*>(e->out_aux16);
  ep.ld_out_aux = e->ld_out_aux;
  ep.out_f32 = e->out_f32;
  ep.ld_out = e->ld_out;
  ep.out_f16 = reinterpret_cast<__half*>(e->out_f16);
  ep.ld_out16 = e->ld_out16;
  ep.out_f16_splits = e->out_f16_splits;
  ep.act = e->act;
  ep.scale_cols = e->scale_cols;
  ep.scale = e->scale;
  {
    auto ok32 = [](const void* ptr, long long pitch_bytes) { return !ptr || ((reinterpret_cast<uintptr_t>(ptr) | uintptr_t(pitch_bytes)) & 31) == 0; };
    ep.wide = ok32(ep.residual, 4LL * ep.ld_out) && ok32(ep.out_f32, 4LL * ep.ld_out) && ok32(ep.out_f16, 2LL * ep.ld_out16) &&
              ok32(ep.aux16, 2LL * ep.ld_aux) && ok32(ep.out_aux16, 2LL * ep.ld_out_aux);
  }
  ep.raster_rows = ep.raster_groups = ep.raster_reps = 0;
  if (ep.act == SEMABS_ACT_MUL_AUX16 && ep.aux_rows >= GEMM_BM && ep.aux_rows < M && M % ep.aux_rows == 0) {
    ep.raster_rows = ep.aux_rows;
    ep.raster_groups = (ep.aux_rows + GEMM_BM - 1) / GEMM_BM;
    ep.raster_reps = M / ep.aux_rows;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int kb_total = kblocks * a_splits;
  if (BN == 256) return launch_gemm<256>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  if (BN == 128) return launch_gemm<128>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  if (BN == 64) return launch_gemm<64>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  return launch_gemm<32>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
}
