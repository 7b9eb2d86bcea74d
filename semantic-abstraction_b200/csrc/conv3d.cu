// Implicit-GEMM 3-D convolution for the voxel UNet on tcgen05 (reference: unet3d.py:16-17 conv3d 3x3x3 pad 1 no
// bias, :428-440 ConvTranspose3d k3 s2 p1 called with output_size, :578 final 1x1x1 conv).
//
// Activations are channels-last fp16 [N, D, H, W, a_splits*C] (a_splits = 2: hi | lo split of the fp32 value).
// One output tile = 128 voxels (a TMA box over (W,H,D,N)) x BN output channels.  For every tap (dz,dy,dx) of a
// tap list the producer warp issues ONE 5-D TMA load of the box shifted by the tap — out-of-bounds voxels are
// zero-filled by TMA, which IS the conv's zero padding — plus the matching [BN x KB] weight slice; the MMA warp
// accumulates all taps x channel blocks x precision passes into one TMEM accumulator (M = 128, N = BN, K = 16).
// Epilogue (4 warps, one voxel per thread): bias, residual add (ExtResNetBlock `out += residual`,
// unet3d.py:254 / decoder skip sum :395-396), ReLU, fp32 + optional fp16 stores, and per-(sample, group)
// sum / sum-of-squares for the NEXT GroupNorm kept in registers across tiles and flushed with fp64 atomics.
// Transposed convolution = 8 output-parity classes, each an ordinary tap list over the input grid whose
// results are scattered with stride 2.
#include <stdlib.h>

#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int CONV_THREADS = 256;
constexpr int MAX_TAPS = 27;
constexpr int MAX_PASSES = 3;

struct ConvParams {
  // problem
  int N, D, H, W;            // input grid (= iteration space)
  int C_in, C_out;           // per-split input channels (multiple of 16), output channels (multiple of 16)
  int ntaps;
  int8_t tap[MAX_TAPS][3];   // (dz, dy, dx) input offset of each tap
  int16_t tap_w[MAX_TAPS];   // weight slice index of each tap
  int w_slices;              // number of weight slices per weight split (27 for 3x3x3)
  int npass;
  int8_t pass_a[MAX_PASSES]; // activation split used by each precision pass
  int8_t pass_w[MAX_PASSES]; // weight split used by each precision pass
  // tiling
  int bw, bh, bd, bn;        // TMA box = tile of 128 voxels
  int tiles_w, tiles_h, tiles_d, tiles_n, n_tiles_out;  // n_tiles_out = C_out / BN
  // output mapping: out voxel = (n, z*os + oz, y*os + oy, x*os + ox) in a grid of (Do, Ho, Wo)
  int os, oz, oy, ox, Do, Ho, Wo;
  // epilogue
  const float* bias;         // [C_out] or null
  const float* residual;     // same layout as out32, or null
  int relu;
  float* out32;              // [N, Do, Ho, Wo, C_out] or null
  __half* out16;             // [N, Do, Ho, Wo, o16_splits*C_out] or null
  int o16_splits;
  double* stats;             // [N, G, 2] (sum, sumsq) or null
  int groups;                // G of the consumer GroupNorm
  int stages;                // depth of the operand ring actually used (<= ConvCfg::STAGES; fewer = several CTAs per SM)
};

// FUSED (precise mode): the three precision terms x_hi W_hi + x_hi W_lo + x_lo W_hi run as TWO stages per (tap, K block)
// instead of three: stage 0 = x_hi against [W_hi ; W_lo] stacked along N (one MMA of width 2 BN, accumulator columns
// [0,BN) and [BN,2BN)), stage 1 = x_lo against W_hi (width BN, columns [0,BN)); the epilogue adds the two column halves.
// The activation box — whose TMA cost (box rows, not bytes) bounds this kernel — is fetched twice per tap, not 3 times.
template <int BN, int KB, bool FUSED>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * KB * 2;
  static constexpr int B_BYTES_RAW = BN * KB * 2;
  static constexpr int B_BYTES = ((FUSED ? 2 : 1) * B_BYTES_RAW + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue transpose staging (coalesced path, BN <= 64): 4 warps x 32 voxel rows x (BN + 4) floats
  static constexpr int TR_PITCH = BN + 4;
  static constexpr int TR_BYTES = BN <= 64 ? 4 * 32 * TR_PITCH * 4 : 0;
  static constexpr int STAGES_RAW = (192 * 1024 - TR_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 12 ? 12 : STAGES_RAW;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + TR_BYTES + 512 + 1024;
  static constexpr uint64_t SWZ = (KB == 64) ? SW_128B : (KB == 32 ? SW_64B : SW_32B);
  static constexpr uint32_t SBO = 8 * KB * 2;  // 8 rows of one swizzle atom
  static constexpr int ACC_COLS = (FUSED ? 2 : 1) * BN;  // columns of one accumulator buffer
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
};

template <int BN, int KB, bool FUSED>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, KB, FUSED>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STG = p.stages;
  float* tr_base = reinterpret_cast<float*>(smem + STG * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STG * Cfg::STAGE_BYTES + Cfg::TR_BYTES);
  uint64_t* empty_bar = full_bar + STG;
  uint64_t* tmem_full = empty_bar + STG;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spatial_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  const int num_tiles = spatial_tiles * p.n_tiles_out;
  const int kblocks = p.C_in / KB;
  const int ksteps = (FUSED ? 2 : p.npass) * p.ntaps * kblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STG; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile id -> (n-channel block fastest, then w, h, d, n): consecutive CTAs share the activation box in L2
  auto decode = [&](int t, int& nb, int& x0, int& y0, int& z0, int& n0) {
    nb = t % p.n_tiles_out;
    int s = t / p.n_tiles_out;
    x0 = (s % p.tiles_w) * p.bw;
    s /= p.tiles_w;
    y0 = (s % p.tiles_h) * p.bh;
    s /= p.tiles_h;
    z0 = (s % p.tiles_d) * p.bd;
    s /= p.tiles_d;
    n0 = s * p.bn;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int nb, x0, y0, z0, n0;
        decode(t, nb, x0, y0, z0, n0);
        if (FUSED) {
          for (int tp = 0; tp < p.ntaps; ++tp) {
            const int dz = p.tap[tp][0], dy = p.tap[tp][1], dx = p.tap[tp][2];
            const int wk_hi = p.tap_w[tp] * p.C_in, wk_lo = (p.w_slices + p.tap_w[tp]) * p.C_in;
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + (h == 0 ? 2 : 1) * Cfg::B_BYTES_RAW);
                tma_load_5d(sA, &tmA, &full_bar[stage], h * p.C_in + kb * KB, x0 + dx, y0 + dy, z0 + dz, n0);
                tma_load_2d(sB, &tmB, &full_bar[stage], wk_hi + kb * KB, nb * BN);
                if (h == 0) tma_load_2d(sB + Cfg::B_BYTES_RAW, &tmB, &full_bar[stage], wk_lo + kb * KB, nb * BN);
                if (++stage == STG) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
        } else {
          for (int ps = 0; ps < p.npass; ++ps) {
            const int a_off = p.pass_a[ps] * p.C_in;
            const int w_off = p.pass_w[ps] * p.w_slices;
            for (int tp = 0; tp < p.ntaps; ++tp) {
              const int dz = p.tap[tp][0], dy = p.tap[tp][1], dx = p.tap[tp][2];
              const int wk = (w_off + p.tap_w[tp]) * p.C_in;
              for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES_RAW);
                tma_load_5d(sA, &tmA, &full_bar[stage], a_off + kb * KB, x0 + dx, y0 + dy, z0 + dz, n0);
                tma_load_2d(sB, &tmB, &full_bar[stage], wk + kb * KB, nb * BN);
                if (++stage == STG) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the loop, the elected lane issues (see umma_f16_elect in ptx.cuh)
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_f16(128, BN);
      constexpr uint32_t idesc_wide = make_idesc_f16(128, FUSED ? 2 * BN : BN);  // x_hi * [W_hi ; W_lo]
      const uint64_t desc0 = make_smem_desc(0, 16, Cfg::SBO, Cfg::SWZ) + (smem_u32(smem) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = desc0 + uint32_t(stage) * uint32_t(Cfg::STAGE_BYTES >> 4);
          const uint64_t db = da + uint32_t(Cfg::A_BYTES >> 4);
          // FUSED: even stages are the wide ones (they also initialise all 2 BN columns at ks == 0)
          const uint32_t id = (FUSED && (ks & 1) == 0) ? idesc_wide : idesc;
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            umma_f16_elect(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), id, (ks | k) != 0, leader);
          umma_commit_elect(&empty_bar[stage], leader);
          if (++stage == STG) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_elect(&tmem_full[acc], leader);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = voxel
    const int lw = r % p.bw, lh = (r / p.bw) % p.bh, ld = (r / (p.bw * p.bh)) % p.bd, ln = r / (p.bw * p.bh * p.bd);
    // GroupNorm statistics of the consumer, accumulated across tiles while the sample index is unchanged
    constexpr int MAXG = 8;
    float gs[MAXG], gq[MAXG];
#pragma unroll
    for (int i = 0; i < MAXG; ++i) gs[i] = gq[i] = 0.f;
    int stat_n = -1, stat_nb = -1;
    const int cpg = p.stats ? p.C_out / p.groups : 1;            // channels per group
    const int gpt = p.stats ? (BN >= cpg ? BN / cpg : 1) : 0;    // groups covered by one BN-wide tile
    auto flush = [&]() {
      if (stat_n < 0) return;
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        if (i < gpt) {
          const float s = warp_sum(gs[i]), s2 = warp_sum(gq[i]);
          if (lane == 0) {
            const int gidx = (BN >= cpg) ? stat_nb * gpt + i : (stat_nb * BN) / cpg;
            double* dst = p.stats + (size_t(stat_n) * p.groups + gidx) * 2;
            atomicAdd(dst, double(s));
            atomicAdd(dst + 1, double(s2));
          }
        }
        gs[i] = gq[i] = 0.f;
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    // Coalesced epilogue (BN <= 64, unit output stride, 32 lanes = 32 consecutive x of one grid row): the accumulator rows (+ bias)
    // go through a per-warp shared-memory transpose, then residual add, ReLU, statistics and the stores run with 4 BN bytes of
    // every voxel row contiguous across BN / 4 lanes.  Row-per-thread float4 accesses touch 32 lines per instruction: measured
    // ~5 us per 128 voxels x 32 channels in every kernel that used them — the 64-channel 64^3 convolutions spent 13.5 us per tile
    // with 7 us of MMAs.
    if (BN <= 64 && p.os == 1 && (p.bw % 32) == 0 && (!p.stats || (cpg % 4) == 0)) {
      constexpr int Q = (BN <= 64 ? BN : 64) / 4;  // float4 columns per voxel row of this tile (4, 8 or 16: divides 32)
      constexpr int VPI = 32 / Q;                  // voxel rows covered by one warp instruction
      constexpr int PITCH = Cfg::TR_PITCH;
      float* tr = tr_base + (warp - 4) * 32 * PITCH;
      const int my_col = lane % Q, my_sub = lane / Q;
      float cs = 0.f, cq = 0.f;
      int my_g = 0;
      auto flush_c = [&]() {
        if (stat_n < 0) return;
        __syncwarp();
        if (lane < 16) tr[lane] = 0.f;
        __syncwarp();
        atomicAdd(&tr[2 * my_g], cs);
        atomicAdd(&tr[2 * my_g + 1], cq);
        __syncwarp();
        if (lane < 2 * p.groups && (tr[lane] != 0.f)) atomicAdd(p.stats + size_t(stat_n) * p.groups * 2 + lane, double(tr[lane]));
        __syncwarp();
        cs = cq = 0.f;
      };
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int nb, x0, y0, z0, n0;
        decode(t, nb, x0, y0, z0, n0);
        const int n = n0 + ln, z = z0 + ld, y = y0 + lh, x = x0 + lw;
        const int n_w = __shfl_sync(0xffffffffu, n, 0);
        const bool ok = n_w < p.N;  // warp-uniform: the 32 lanes share (n, z, y)
        if (p.stats && (n_w != stat_n || nb != stat_nb)) {
          flush_c();
          stat_n = ok ? n_w : -1, stat_nb = nb;
          my_g = (nb * BN + 4 * my_col) / cpg;
        }
        // first voxel row of the warp: lanes are consecutive x
        const size_t row_w = ((size_t(n_w) * p.Do + __shfl_sync(0xffffffffu, z, 0)) * p.Ho + __shfl_sync(0xffffffffu, y, 0)) * p.Wo +
                             __shfl_sync(0xffffffffu, x, 0);
        float4 rq[Q];
        if (p.residual && ok) {
          const float* src = p.residual + (row_w + my_sub) * p.C_out + nb * BN + 4 * my_col;
#pragma unroll
          for (int k = 0; k < Q; ++k) rq[k] = *reinterpret_cast<const float4*>(src + size_t(k * VPI) * p.C_out);
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        float* mine = tr + lane * PITCH;
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t rr[16], r2[16];
          tmem_ld_32x32b_x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * Cfg::ACC_COLS + c * 16), rr);
          if (FUSED) tmem_ld_32x32b_x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * Cfg::ACC_COLS + BN + c * 16), r2);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v;
            v.x = FUSED ? __uint_as_float(rr[j]) + __uint_as_float(r2[j]) : __uint_as_float(rr[j]);
            v.y = FUSED ? __uint_as_float(rr[j + 1]) + __uint_as_float(r2[j + 1]) : __uint_as_float(rr[j + 1]);
            v.z = FUSED ? __uint_as_float(rr[j + 2]) + __uint_as_float(r2[j + 2]) : __uint_as_float(rr[j + 2]);
            v.w = FUSED ? __uint_as_float(rr[j + 3]) + __uint_as_float(r2[j + 3]) : __uint_as_float(rr[j + 3]);
            *reinterpret_cast<float4*>(mine + c * 16 + j) = v;
          }
        }
        // the accumulator has been read: hand the TMEM buffer back before the global-memory phase
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        if (ok) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + nb * BN + 4 * my_col);
          const size_t col = size_t(nb) * BN + 4 * my_col;
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            const int vrow = k * VPI + my_sub;  // voxel row within the warp
            float4 v = *reinterpret_cast<const float4*>(tr + vrow * PITCH + 4 * my_col);
            v.x += b4.x, v.y += b4.y, v.z += b4.z, v.w += b4.w;
            if (p.residual) v.x += rq[k].x, v.y += rq[k].y, v.z += rq[k].z, v.w += rq[k].w;
            if (p.relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
            const size_t orow = row_w + vrow;
            if (p.out32) *reinterpret_cast<float4*>(p.out32 + orow * p.C_out + col) = v;
            if (p.out16) {
              __half* o = p.out16 + orow * size_t(p.o16_splits) * p.C_out + col;
              const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
              uint2 u;
              u.x = *reinterpret_cast<const uint32_t*>(&h0), u.y = *reinterpret_cast<const uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(o) = u;
              if (p.o16_splits == 2) {
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
                u.x = *reinterpret_cast<const uint32_t*>(&l0), u.y = *reinterpret_cast<const uint32_t*>(&l1);
                *reinterpret_cast<uint2*>(o + p.C_out) = u;
              }
            }
            cs += (v.x + v.y) + (v.z + v.w);
            cq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
          }
        }
        __syncwarp();  // the staging rows are rewritten by the next tile
      }
      if (p.stats) flush_c();
    } else
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int nb, x0, y0, z0, n0;
      decode(t, nb, x0, y0, z0, n0);
      const int n = n0 + ln, z = z0 + ld, y = y0 + lh, x = x0 + lw;
      const bool ok = n < p.N;  // W/H/D are multiples of the box, only the sample axis can overhang
      // warp-uniform sample index: rows-per-sample is a multiple of 32 (or the whole tile is one sample)
      const int n_w = __shfl_sync(0xffffffffu, n, 0);
      if (p.stats && (n_w != stat_n || nb != stat_nb)) {
        flush();
        stat_n = n_w < p.N ? n_w : -1;
        stat_nb = nb;
      }
      const size_t ovox = ((size_t(n) * p.Do + (z * p.os + p.oz)) * p.Ho + (y * p.os + p.oy)) * p.Wo + (x * p.os + p.ox);
      // the residual (decoder skip / block input) does not depend on the accumulator: chunk 0 is requested BEFORE waiting for
      // the MMAs and chunk c+1 while chunk c is processed.  (ncu, round 2: the transposed convolutions into the 128^3 level ran at
      // 1 TB/s with the tensor pipe 2-17 % busy — one exposed global-load round trip per 16-channel chunk of every tile.)
      float4 rs_cur[4], rs_nxt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) rs_cur[j] = rs_nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* rs_row = (p.residual && ok) ? p.residual + ovox * p.C_out + nb * BN : nullptr;
      if (rs_row) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rs_cur[j] = *reinterpret_cast<const float4*>(rs_row + 4 * j);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        if (rs_row && c + 1 < BN / 16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rs_nxt[j] = *reinterpret_cast<const float4*>(rs_row + (c + 1) * 16 + 4 * j);
        }
        uint32_t rr[16];
        tmem_ld_32x32b_x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * Cfg::ACC_COLS + c * 16), rr);
        uint32_t r2[16];
        if (FUSED) tmem_ld_32x32b_x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * Cfg::ACC_COLS + BN + c * 16), r2);
        tc_wait_ld();
        if (ok) {
          const int col0 = nb * BN + c * 16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = FUSED ? __uint_as_float(rr[j]) + __uint_as_float(r2[j]) : __uint_as_float(rr[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + j);
              v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
            }
          }
          if (p.residual) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = rs_cur[j >> 2];
              v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.out32) {
            float* o = p.out32 + ovox * p.C_out + col0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (p.out16) {
            __half* o = p.out16 + ovox * size_t(p.o16_splits) * p.C_out + col0;
            __align__(16) __half2 hh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) hh[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<const uint4*>(hh)[0];
            reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<const uint4*>(hh)[1];
            if (p.o16_splits == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f = __half22float2(hh[j]);
                hh[j] = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
              }
              reinterpret_cast<uint4*>(o + p.C_out)[0] = reinterpret_cast<const uint4*>(hh)[0];
              reinterpret_cast<uint4*>(o + p.C_out)[1] = reinterpret_cast<const uint4*>(hh)[1];
            }
          }
          if (p.stats) {
            // 16 consecutive channels -> 16/cpg partial group sums, added to the accumulators of the groups
            // [first, first + per) of this BN-wide tile (static register indices, compare instead of divide)
            float ps[8], pq[8];
            group_sums16(v, cpg, ps, pq);
            const int per = cpg >= 16 ? 1 : 16 / cpg;
            const int first = (BN >= cpg) ? (c * 16) / cpg : 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (k < per) {
#pragma unroll
                for (int i = 0; i < MAXG; ++i)
                  if (i == first + k) gs[i] += ps[k], gq[i] += pq[k];
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rs_cur[j] = rs_nxt[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.stats && stat_n >= 0) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int KB, bool FUSED>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t st) {
  using Cfg = ConvCfg<BN, KB, FUSED>;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(conv3d_igemm_kernel<BN, KB, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL));
    configured = true;
  }
  const int num_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * p.n_tiles_out;
  // Short tap lists over big grids with narrow tiles (the parity classes of the transposed convolution into the 128^3 level)
  // are not tensor-bound (ncu: 5 us per 128-voxel tile whatever the tap count, 2-18 % tensor pipe, ~2 TB/s of strided 128-byte
  // rows): a 4-stage ring leaves room for two CTAs per SM (shared memory and 2 x TMEM_COLS <= 512), so two tiles' epilogues and
  // operand loads overlap: 275 -> 230 us per class.  (BN = 64, the 64^3 level: 80 -> 98 us, not taken.  What these launches
  // want is one pass over all eight classes — input read once, contiguous output rows)
  ConvParams q = p;
  int per_sm = 1;
  q.stages = Cfg::STAGES;
  if (BN <= 32 && p.ntaps <= 8 && 2 * Cfg::TMEM_COLS <= 512 && num_tiles >= 4 * num_sms()) {
    for (int s = 4; s >= 2; --s) {
      if (s < Cfg::STAGES && s * Cfg::STAGE_BYTES + Cfg::TR_BYTES + 1536 <= 112 * 1024) {
        q.stages = s, per_sm = 2;
        break;
      }
    }
  }
  const int smem_bytes = q.stages * Cfg::STAGE_BYTES + Cfg::TR_BYTES + 512 + 1024;
  const int cap = per_sm * num_sms();
  const int grid = num_tiles < cap ? num_tiles : cap;
  conv3d_igemm_kernel<BN, KB, FUSED><<<grid, CONV_THREADS, smem_bytes, st>>>(tmA, tmB, q);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int KB>
static int dispatch_bn(int BN, bool fused, const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t st) {
  if (fused) {
    switch (BN) {
      case 16: return launch_conv<16, KB, true>(tmA, tmB, p, st);
      case 32: return launch_conv<32, KB, true>(tmA, tmB, p, st);
      case 64: return launch_conv<64, KB, true>(tmA, tmB, p, st);
      default: return launch_conv<128, KB, true>(tmA, tmB, p, st);
    }
  }
  switch (BN) {
    case 16: return launch_conv<16, KB, false>(tmA, tmB, p, st);
    case 32: return launch_conv<32, KB, false>(tmA, tmB, p, st);
    case 64: return launch_conv<64, KB, false>(tmA, tmB, p, st);
    default: return launch_conv<128, KB, false>(tmA, tmB, p, st);
  }
}

}  // namespace sb

using namespace sb;

// kind: 0 = 3x3x3 pad 1 (27 taps), 1 = 1x1x1, 2 = transposed k3 s2 p1 output-parity class `parity` (bit2=z,bit1=y,bit0=x),
// 3 = stride-2 3x3x3 pad 1 over the 2x grid (adjoint of kind 2), input-parity class `parity` per call
extern "C" int semabs_conv3d(const void* x16, int32_t a_splits, const void* w16, int32_t w_splits, int32_t kind,
                             int32_t parity, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C_in, int32_t C_out,
                             int32_t precise, const float* bias, const float* residual, int32_t relu, float* out32,
                             void* out16, int32_t o16_splits, double* stats, int32_t groups, void* stream) {
  SB_REQUIRE(x16 && w16 && (out32 || out16), "semabs_conv3d: null pointer");
  SB_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "semabs_conv3d: bad grid");
  SB_REQUIRE(C_in % 16 == 0 && (C_in <= 64 ? (C_in == 16 || C_in == 32 || C_in == 64) : C_in % 64 == 0),
             "semabs_conv3d: C_in=%d must be 16, 32 or a multiple of 64 (pad the channels)", C_in);
  SB_REQUIRE(C_out % 16 == 0, "semabs_conv3d: C_out=%d must be a multiple of 16", C_out);
  SB_REQUIRE(!precise || (a_splits == 2 && w_splits == 2), "semabs_conv3d: precise mode needs hi/lo activations and weights");
  SB_REQUIRE(!stats || (groups > 0 && groups <= 8 && C_out % groups == 0), "semabs_conv3d: bad GroupNorm groups");
  ConvParams p{};
  p.N = N, p.D = D, p.H = H, p.W = W, p.C_in = C_in, p.C_out = C_out;
  p.os = 1, p.oz = p.oy = p.ox = 0, p.Do = D, p.Ho = H, p.Wo = W;
  if (kind == 0) {
    p.ntaps = 27, p.w_slices = 27;
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          const int t = (kd * 3 + kh) * 3 + kw;
          p.tap[t][0] = kd - 1, p.tap[t][1] = kh - 1, p.tap[t][2] = kw - 1, p.tap_w[t] = t;
        }
  } else if (kind == 1) {
    p.ntaps = 1, p.w_slices = 1;
    p.tap[0][0] = p.tap[0][1] = p.tap[0][2] = 0, p.tap_w[0] = 0;
  } else if (kind == 2) {
    // out[2j + par] = sum over (k, delta): par 0 -> (k=1, +0); par 1 -> (k=2, +0), (k=0, +1)   (o = 2i - 1 + k)
    p.w_slices = 27;
    const int pz = (parity >> 2) & 1, py = (parity >> 1) & 1, px = parity & 1;
    const int kz[2][2] = {{1, -1}, {2, 0}}, dl[2][2] = {{0, 0}, {0, 1}};
    int t = 0;
    for (int a = 0; a < (pz ? 2 : 1); ++a)
      for (int b = 0; b < (py ? 2 : 1); ++b)
        for (int c = 0; c < (px ? 2 : 1); ++c) {
          p.tap[t][0] = dl[pz][a], p.tap[t][1] = dl[py][b], p.tap[t][2] = dl[px][c];
          p.tap_w[t] = (kz[pz][a] * 3 + kz[py][b]) * 3 + kz[px][c];
          ++t;
        }
    p.ntaps = t;
    p.os = 2, p.oz = pz, p.oy = py, p.ox = px, p.Do = 2 * D, p.Ho = 2 * H, p.Wo = 2 * W;
  } else if (kind == 3) {
    // adjoint of kind 2 (data gradient of the transposed conv = 3x3x3 conv, stride 2, padding 1, over the 2x grid):
    // out[i] = sum_k in[2i - 1 + k] W[k].  Input parity class `parity` of the (2D,2H,2W) grid per call:
    // parity 0 -> (k=1, j=i); parity 1 -> (k=0, j=i-1), (k=2, j=i) with in[2j + parity]
    p.w_slices = 27;
    const int pz = (parity >> 2) & 1, py = (parity >> 1) & 1, px = parity & 1;
    const int kk[2][2] = {{1, -1}, {0, 2}}, dl[2][2] = {{0, 0}, {-1, 0}};
    int t = 0;
    for (int a = 0; a < (pz ? 2 : 1); ++a)
      for (int b = 0; b < (py ? 2 : 1); ++b)
        for (int c = 0; c < (px ? 2 : 1); ++c) {
          p.tap[t][0] = dl[pz][a], p.tap[t][1] = dl[py][b], p.tap[t][2] = dl[px][c];
          p.tap_w[t] = (kk[pz][a] * 3 + kk[py][b]) * 3 + kk[px][c];
          ++t;
        }
    p.ntaps = t;
  } else {
    SB_REQUIRE(false, "semabs_conv3d: unknown kind %d", kind);
  }
  if (precise) {
    p.npass = 3;
    p.pass_a[0] = 0, p.pass_w[0] = 0;
    p.pass_a[1] = 1, p.pass_w[1] = 0;
    p.pass_a[2] = 0, p.pass_w[2] = 1;
  } else {
    p.npass = 1, p.pass_a[0] = 0, p.pass_w[0] = 0;
  }
  // 128-voxel box
  p.bw = W < 128 ? W : 128;
  p.bh = H < 128 / p.bw ? H : 128 / p.bw;
  p.bd = D < 128 / (p.bw * p.bh) ? D : 128 / (p.bw * p.bh);
  p.bn = 128 / (p.bw * p.bh * p.bd);
  SB_REQUIRE(p.bw * p.bh * p.bd * p.bn == 128 && W % p.bw == 0 && H % p.bh == 0 && D % p.bd == 0,
             "semabs_conv3d: grid %dx%dx%d cannot be tiled into 128-voxel boxes (power-of-two extents expected)", D, H, W);
  SB_REQUIRE(p.bn == 1 || (p.bw * p.bh * p.bd) % 32 == 0, "semabs_conv3d: grid too small (needs >= 32 voxels)");
  p.tiles_w = W / p.bw, p.tiles_h = H / p.bh, p.tiles_d = D / p.bd, p.tiles_n = (N + p.bn - 1) / p.bn;
  const int spatial_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  // widest N tile that still gives every SM work (deep levels have few voxels but many channels)
  int BN = 128;
  while (BN > 16 && (C_out % BN != 0 || spatial_tiles * (C_out / BN) < num_sms())) BN >>= 1;
  while (C_out % BN != 0) BN >>= 1;
  if (stats) {
    const int cpg = C_out / groups;
    SB_REQUIRE(cpg >= 2 && (cpg & (cpg - 1)) == 0, "semabs_conv3d: channels per group must be a power of two >= 2");
    // a BN tile must hold whole groups, or lie inside one group
    SB_REQUIRE(BN % cpg == 0 || cpg % BN == 0, "semabs_conv3d: BN=%d incompatible with %d channels per group", BN, cpg);
    SB_REQUIRE(BN < cpg || BN / cpg <= 8, "semabs_conv3d: too many groups per tile");
  }
  p.n_tiles_out = C_out / BN;
  p.bias = bias, p.residual = residual, p.relu = relu, p.out32 = out32, p.out16 = (__half*)out16;
  p.o16_splits = o16_splits, p.stats = stats, p.groups = groups;

  const int KB = C_in >= 64 ? 64 : C_in;
  const CUtensorMapSwizzle swz = KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB;
  {
    const uint64_t C = uint64_t(a_splits) * C_in;
    uint64_t dims[5] = {C, uint64_t(W), uint64_t(H), uint64_t(D), uint64_t(N)};
    uint64_t str[4] = {C * 2, C * 2 * W, C * 2 * W * H, C * 2 * W * H * D};
    const uint8_t* base = static_cast<const uint8_t*>(x16);
    if (kind == 3) {
      // strided view of one parity class of the [N, 2D, 2H, 2W, C] tensor
      const uint64_t W2 = 2 * uint64_t(W), H2 = 2 * uint64_t(H), D2 = 2 * uint64_t(D);
      str[0] = 2 * C * 2, str[1] = 2 * W2 * C * 2, str[2] = 2 * H2 * W2 * C * 2, str[3] = D2 * H2 * W2 * C * 2;
      base += ((uint64_t((parity >> 2) & 1) * H2 + uint64_t((parity >> 1) & 1)) * W2 + uint64_t(parity & 1)) * C * 2;
    }
    uint32_t box[5] = {uint32_t(KB), uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bd), uint32_t(p.bn)};
    if (int rc = make_tmap_f16(&tmA, base, 5, dims, str, box, swz)) return rc;
  }
  {
    const uint64_t K = uint64_t(w_splits) * p.w_slices * C_in;
    uint64_t dims[2] = {K, uint64_t(C_out)};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {uint32_t(KB), uint32_t(BN)};
    if (int rc = make_tmap_f16(&tmB, w16, 2, dims, str, box, swz)) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  static const bool fused_enabled = [] {
    const char* e = getenv("SEMABS_CONV_FUSED");  // debugging switch: 0 = the original three-stage precise schedule
    return !(e && e[0] == '0');
  }();
  const bool fused = precise != 0 && fused_enabled;
  if (KB == 64) return dispatch_bn<64>(BN, fused, tmA, tmB, p, st);
  if (KB == 32) return dispatch_bn<32>(BN, fused, tmA, tmB, p, st);
  return dispatch_bn<16>(BN, fused, tmA, tmB, p, st);
}
