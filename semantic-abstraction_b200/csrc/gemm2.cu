// CTA-pair (cta_group::2) variant of the persistent TN GEMM of gemm.cu for the large shapes of the relevancy sweep:
//
//   one 256 x 256 output tile per CTA PAIR: each CTA loads its own 128 rows of A and its own 128 of the 256 B rows (N) per
//   64-wide K block, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads A / B from both CTAs'
//   shared memory and accumulates 128 x 256 fp32 into EACH CTA's tensor memory; both CTAs run the gemm.cu epilogue on their
//   own 128 rows.
//
// Why: ncu on the 128 x 256 single-CTA kernel (profiles/r02_ncu_full_vit_bwd.txt) shows the tensor pipe at 64-88 % with
// every operand byte arriving from L2: a 128 x 256 tile pulls (128 + 256) x 64 x 2 B = 48 KB per 512 tensor cycles = 96
// B/clk/SM through the L2 -> SM path and the same through shared memory; the pair halves the B traffic per CTA: 32 KB per
// 512 cycles = 64 B/clk/SM, and the freed shared memory holds 6 pipeline stages instead of 4.
//
// Protocol (per stage s; "leader" = cluster rank 0):
//   full[s]   lives in the leader, count 2: the leader's producer arrives with expect_tx(bytes of BOTH CTAs), the peer's
//             producer arrives remotely; both CTAs' TMA loads credit their bytes to the leader's barrier (cta_group::2 TMA);
//   empty[s]  one per CTA, count 1: the leader's tcgen05.commit multicasts the arrive to both CTAs;
//   tmem_full[a]  one per CTA, count 1, multicast commit;  tmem_empty[a] lives in the leader, count 2 x 8 epilogue warps
//             (the peer's warps arrive remotely).
#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int G2_BM = 128;            // rows per CTA (256 per pair)
constexpr int G2_BN = 256;            // tile columns (128 B rows loaded per CTA)
constexpr int G2_BK = 64;
constexpr int G2_THREADS = 384;
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;        // 16 KB
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;  // 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + 256 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_f16_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                        int kblocks_total, int kblocks_wrap_b, EpiParams ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (M + 2 * G2_BM - 1) / (2 * G2_BM);   // 256-row tile rows
  const int num_n = N / G2_BN;
  const int num_tiles = (ep.raster_rows ? ep.raster_groups * ep.raster_reps : num_m) * num_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocation of both CTAs visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int m_blk, n_blk;
        if (!tile_coords(ep, t, num_m, num_n, 2 * G2_BM, m_blk, n_blk)) continue;
        const int row_a = m_blk * 2 * G2_BM + int(rank) * G2_BM;
        const int row_b = n_blk * G2_BN + int(rank) * (G2_BN / 2);
        for (int kb = 0; kb < kblocks_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * G2_STAGE_BYTES;
          uint8_t* sB = sA + G2_A_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * G2_STAGE_BYTES);
          else mbar_arrive_cluster(full_leader);
          tma_load_2d_pair(sA, &tmA, full_leader, kb * G2_BK, row_a);
          tma_load_2d_pair(sB, &tmB, full_leader, (kb % kblocks_wrap_b) * G2_BK, row_b);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only (whole warp convergent, elected lane issues) =====
    if (is_leader) {
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_f16(2 * G2_BM, G2_BN);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024, SW_128B) + (smem_u32(smem) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        {
          int m_blk, n_blk;
          if (!tile_coords(ep, t, num_m, num_n, 2 * G2_BM, m_blk, n_blk)) continue;
        }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BN;
        for (int kb = 0; kb < kblocks_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = desc0 + uint32_t(stage) * uint32_t(G2_STAGE_BYTES >> 4);
          const uint64_t db = da + uint32_t(G2_A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k)
            umma_f16_pair_elect(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0, leader);
          umma_commit_pair_elect(&empty_bar[stage], 3u, leader);  // the slot is reusable in BOTH CTAs once these MMAs retire
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair_elect(&tmem_full[acc], 3u, leader);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, own 128 rows) =====
    const uint32_t empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0), empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int m_blk, n_blk;
      if (!tile_coords(ep, t, num_m, num_n, 2 * G2_BM, m_blk, n_blk)) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<G2_BN>(ep, tmem_base + uint32_t(acc * G2_BN), m_blk * 2 * G2_BM + int(rank) * G2_BM, n_blk * G2_BN, M, N,
                                warp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? empty_leader1 : empty_leader0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still address this CTA's shared / tensor memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int kblocks_total, int kblocks_wrap_b,
                     const EpiParams& ep, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16_tn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    configured = true;
  }
  const int num_tiles = ((M + 2 * G2_BM - 1) / (2 * G2_BM)) * (N / G2_BN);
  int pairs = num_sms() / 2;
  if (num_tiles < pairs) pairs = num_tiles;
  gemm_f16_tn_pair_kernel<<<2 * pairs, G2_THREADS, G2_SMEM, stream>>>(tmA, tmB, M, N, kblocks_total, kblocks_wrap_b, ep);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sb
