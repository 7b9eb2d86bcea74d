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
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 fp16 = 128 bytes = one swizzle-128B row
constexpr int GEMM_THREADS = 384;

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
        if (!tile_coords(ep, t, num_m, num_n, GEMM_BM, m_blk, n_blk)) continue;
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
          if (!tile_coords(ep, t, num_m, num_n, GEMM_BM, m_blk, n_blk)) continue;
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
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      if (!tile_coords(ep, t, num_m, num_n, GEMM_BM, m_blk, n_blk)) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<BN>(ep, tmem_base + uint32_t(acc * BN), m_blk * GEMM_BM, n_blk * BN, M, N, warp, lane);
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

int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int kblocks_total, int kblocks_wrap_b,
                     const EpiParams& ep, cudaStream_t stream);  // gemm2.cu
static int g_gemm_pair = 1;

}  // namespace sb

// 1 (default): shapes that qualify for 128 x 256 tiles run on CTA pairs (gemm2.cu: 256 x 256 per pair); 0: single-CTA kernels
// only (A/B measurements, cross-check in the tests)
extern "C" int semabs_set_gemm_pair(int32_t enable) {
  sb::g_gemm_pair = enable ? 1 : 0;
  return 0;
}

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
  const bool pair = BN == 256 && g_gemm_pair;
  const int tile_rows = pair ? 2 * GEMM_BM : GEMM_BM;
  ep.raster_rows = ep.raster_groups = ep.raster_reps = 0;
  if (ep.act == SEMABS_ACT_MUL_AUX16 && ep.aux_rows >= tile_rows && ep.aux_rows < M && M % ep.aux_rows == 0) {
    ep.raster_rows = ep.aux_rows;
    ep.raster_groups = (ep.aux_rows + tile_rows - 1) / tile_rows;
    ep.raster_reps = M / ep.aux_rows;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int kb_total = kblocks * a_splits;
  if (pair) {
    // B rows per CTA = 128: the pair kernel's B box is [64 x 128]
    uint64_t dims[2] = {uint64_t(K), uint64_t(N)};
    uint64_t str[1] = {uint64_t(ldb) * 2};
    uint32_t box[2] = {GEMM_BK, 128};
    if (int rc = make_tmap_f16(&tmB, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    return launch_gemm_pair(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  }
  if (BN == 256) return launch_gemm<256>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  if (BN == 128) return launch_gemm<128>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  if (BN == 64) return launch_gemm<64>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
  return launch_gemm<32>(tmA, tmB, M, N, kb_total, kblocks, ep, st);
}
