// Transposed 3-D convolution (ConvTranspose3d k3 s2 p1 called with output_size = 2x, reference unet3d.py:428-440) into a
// 32-channel level, ALL EIGHT output-parity classes in one pass (the decoder's skip sum and bias in the epilogue, the next
// GroupNorm's statistics accumulated on the way: Upsampling.forward + summation joining, unet3d.py:385-396).
//
// Why: the per-class launches of conv3d.cu (8 x 230-280 us into the 128^3 level) are not tensor-bound — ncu: 5 us per
// 128-voxel tile whatever the tap count, 2-18 % tensor pipe, ~2 TB/s of 128-byte rows at a stride of two voxels, the input
// level read eight times.  One output voxel (2z + pz, 2y + py, 2x + px) is
//     sum over input shifts s <= p (componentwise, s in {0,1}^3) of  W[k = p + 1 - 2s] x[z + sz, y + sy, x + sx],
// so a tile of 128 INPUT voxels needs 8 shifted activation boxes, and shift s serves every class p >= s.  Here:
//   * accumulator = 8 classes x 32 channels = 256 TMEM columns (two buffers = all of TMEM), class c = pz*4 + py*2 + px at
//     columns [32c, 32c + 32);
//   * one pipeline stage = (precision pass, shift, 64-channel K block): the shifted box (TMA, out-of-bounds = zero) + the weight
//     slices of the classes using that shift stacked along N; the MMA warp issues one tcgen05.mma per run of consecutive
//     classes (shift 0: N = 256; 27 (class, tap) pairs in 14 instructions instead of 27 N = 32 ones — an N = 32 instruction costs
//     40 cycles, an N = 256 one 128);
//   * epilogue thread = input voxel: its 8 output voxels x 32 channels, the two px neighbours being adjacent 128-byte rows.
// Input read once, output and skip tensor touched once, contiguously per (z, y) output row pair.
#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int CT_THREADS = 256;
constexpr int CT_KB = 64;                    // input channels per stage (one 128-byte swizzle row)
constexpr int CT_CO = 32;                    // output channels (one class = 32 accumulator columns)
constexpr int CT_A_BYTES = 128 * CT_KB * 2;  // 16 KB
constexpr int CT_SLICE = CT_CO * CT_KB * 2;  // 4 KB: one class's weight slice of a K block
constexpr int CT_STAGE = CT_A_BYTES + 8 * CT_SLICE;
constexpr int CT_STAGES = 3;
constexpr int CT_TR_PITCH = 68;                               // floats per staged row pair (64 + 4: conflict-free 16-byte accesses)
constexpr int CT_TR_BYTES = 4 * 32 * CT_TR_PITCH * 4;         // transpose staging of the 4 epilogue warps
constexpr int CT_SMEM = CT_STAGES * CT_STAGE + CT_TR_BYTES + 512 + 1024;

struct ConvTParams {
  int N, D, H, W;  // input grid
  int C_in, C_out;
  int n_cb;        // C_out / 32 channel blocks: tile = (128 input voxels, one block of 32 output channels)
  int npass;
  int8_t pass_a[3], pass_w[3];
  int bw, bh, bd, bn;
  int tiles_w, tiles_h, tiles_d, tiles_n;
  const float* bias;      // [32] or null
  const float* residual;  // [N, 2D, 2H, 2W, 32] or null (the encoder feature of the output level)
  float* out32;           // [N, 2D, 2H, 2W, C_out] or null
  __half* out_planar;     // [N][2 C_out / 8 chunks: hi then lo][8 D H W voxels][8] fp16 = the halo convolution's operand layout, or null
  double* stats;          // [N, G, 2] or null
  int groups;
};

__global__ void __launch_bounds__(CT_THREADS, 1)
convt_allparity_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ ConvTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tr_base = reinterpret_cast<float*>(smem + CT_STAGES * CT_STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CT_STAGES * CT_STAGE + CT_TR_BYTES);
  uint64_t* empty_bar = full_bar + CT_STAGES;
  uint64_t* tmem_full = empty_bar + CT_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * p.n_cb;
  const int kblocks = p.C_in / CT_KB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CT_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tmem_full[a], 1), mbar_init(&tmem_empty[a], 4);
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

  // tile id -> (channel block fastest: neighbouring CTAs share the activation boxes in L2, then w, h, d, n)
  auto decode = [&](int t, int& cb, int& x0, int& y0, int& z0, int& n0) {
    cb = t % p.n_cb;
    int s = t / p.n_cb;
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
        int cb, x0, y0, z0, n0;
        decode(t, cb, x0, y0, z0, n0);
        for (int ps = 0; ps < p.npass; ++ps) {
          const int a_off = p.pass_a[ps] * p.C_in, w_off = p.pass_w[ps] * 27;
          for (int s = 0; s < 8; ++s) {
            const int sz = (s >> 2) & 1, sy = (s >> 1) & 1, sx = s & 1;
            const int ncls = 8 >> (sz + sy + sx);  // classes p >= s
            for (int kb = 0; kb < kblocks; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sA = smem + stage * CT_STAGE;
              uint8_t* sB = sA + CT_A_BYTES;
              mbar_arrive_expect_tx(&full_bar[stage], CT_A_BYTES + ncls * CT_SLICE);
              tma_load_5d(sA, &tmA, &full_bar[stage], a_off + kb * CT_KB, x0 + sx, y0 + sy, z0 + sz, n0);
              int idx = 0;
              for (int c = 0; c < 8; ++c) {
                if ((c & s) != s) continue;
                // tap of class (pz, py, px) under shift (sz, sy, sx): k = p + 1 - 2 s per axis
                const int kz = ((c >> 2) & 1) + 1 - 2 * sz, ky = ((c >> 1) & 1) + 1 - 2 * sy, kx = (c & 1) + 1 - 2 * sx;
                tma_load_2d(sB + idx * CT_SLICE, &tmB, &full_bar[stage], (w_off + (kz * 3 + ky) * 3 + kx) * p.C_in + kb * CT_KB, cb * CT_CO);
                ++idx;
              }
              if (++stage == CT_STAGES) stage = 0, phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint64_t desc0 = make_smem_desc(0, 16, 1024, SW_128B) + (smem_u32(smem) >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      bool first = true;
      for (int ps = 0; ps < p.npass; ++ps) {
        for (int s = 0; s < 8; ++s) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = desc0 + uint32_t(stage) * uint32_t(CT_STAGE >> 4);
            const uint64_t db0 = da + uint32_t(CT_A_BYTES >> 4);
            // runs of consecutive classes among {c : c & s == s}; their slices are stacked in class order
            int pos = 0, c = 0;
            while (c < 8) {
              if ((c & s) != s) {
                ++c;
                continue;
              }
              int len = 1;
              while (c + len < 8 && ((c + len) & s) == s) ++len;
              const uint32_t idesc = make_idesc_f16(128, CT_CO * len);
              const uint64_t db = db0 + uint32_t(pos) * uint32_t(CT_SLICE >> 4);
#pragma unroll
              for (int k = 0; k < CT_KB / 16; ++k)
                umma_f16_elect(d_tmem + uint32_t(c * CT_CO), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, !(first && k == 0), leader);
              pos += len, c += len;
            }
            first = false;  // the very first stage is shift 0: one N = 256 run initialises every column
            umma_commit_elect(&empty_bar[stage], leader);
            if (++stage == CT_STAGES) stage = 0, phase ^= 1;
          }
        }
      }
      umma_commit_elect(&tmem_full[acc], leader);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = input voxel
    const int lw = r % p.bw, lh = (r / p.bw) % p.bh, ld = (r / (p.bw * p.bh)) % p.bd, ln = r / (p.bw * p.bh * p.bd);
    const int Do = 2 * p.D, Ho = 2 * p.H, Wo = 2 * p.W;
    constexpr int MAXG = 8;
    float gs[MAXG], gq[MAXG];
#pragma unroll
    for (int i = 0; i < MAXG; ++i) gs[i] = gq[i] = 0.f;
    int stat_n = -1, stat_cb = -1;
    const int cpg = p.stats ? p.C_out / p.groups : 16;  // channels per group (power of two >= 2)
    auto flush = [&]() {
      if (stat_n < 0) return;
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        if (i < p.groups) {
          const float s = warp_sum(gs[i]), s2 = warp_sum(gq[i]);
          if (lane == 0) {
            double* dst = p.stats + (size_t(stat_n) * p.groups + i) * 2;
            atomicAdd(dst, double(s));
            atomicAdd(dst + 1, double(s2));
          }
        }
        gs[i] = gq[i] = 0.f;
      }
    };
    float bias_r[CT_CO];
    int bias_cb = -1;
    auto load_bias = [&](int cb) {
      if (cb == bias_cb) return;
      bias_cb = cb;
#pragma unroll
      for (int j = 0; j < CT_CO; ++j) bias_r[j] = p.bias ? p.bias[cb * CT_CO + j] : 0.f;
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    // Coalesced path (32 lanes = 32 consecutive x of one input row, i.e. box width % 32 == 0): the two px classes of a thread
    // are adjacent 128-byte output rows, so a warp's (pz, py) pair of classes is 8 KB of contiguous output.  The accumulator
    // rows (+ bias) go through a per-warp shared-memory transpose; skip sum, statistics and the stores then run on 512 contiguous
    // bytes per instruction.  (Row-per-thread float4 accesses touch 32 lines per instruction: measured 5 us per class and tile,
    // the same in every kernel that uses them, against 0.6 us of MMAs.)
    const bool coalesced = (p.bw % 32) == 0;
    float* tr = tr_base + (warp - 4) * 32 * CT_TR_PITCH;
    const int my_col = lane & 15;                               // float4 column of this lane in the transposed domain
    int my_g = 0;                                               // its GroupNorm group in the current channel block
    float cs = 0.f, cq = 0.f;                                   // statistics of this lane's group (coalesced path)
    auto flush_coalesced = [&]() {
      if (stat_n < 0) return;
      // lanes with the same group: reduce through shared memory (rare: once per sample and warp)
      __syncwarp();
      if (lane < 16) tr[lane] = 0.f;
      __syncwarp();
      atomicAdd(&tr[2 * my_g], cs);
      atomicAdd(&tr[2 * my_g + 1], cq);
      __syncwarp();
      if (lane < 2 * p.groups) atomicAdd(p.stats + size_t(stat_n) * p.groups * 2 + lane, double(tr[lane]));
      __syncwarp();
      cs = cq = 0.f;
    };
    if (coalesced) {
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int cb, x0, y0, z0, n0;
        decode(t, cb, x0, y0, z0, n0);
        load_bias(cb);
        const int n = n0 + ln, z = z0 + ld, y = y0 + lh, x = x0 + lw;
        const int n_w = __shfl_sync(0xffffffffu, n, 0), z_w = __shfl_sync(0xffffffffu, z, 0), y_w = __shfl_sync(0xffffffffu, y, 0);
        const int x_w = __shfl_sync(0xffffffffu, x, 0);
        const bool ok = n_w < p.N;  // warp-uniform
        if (p.stats && (n_w != stat_n || cb != stat_cb)) {
          flush_coalesced();
          stat_n = ok ? n_w : -1, stat_cb = cb;
          my_g = (cb * CT_CO + 4 * (my_col & 7)) / cpg;  // (cpg is a multiple of 4 on this path)
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int cp2 = 0; cp2 < 4; ++cp2) {  // (pz, py) pairs; both px classes at once
          const int pz = cp2 >> 1, py = cp2 & 1;
          // the warp's 8 KB of output: rows 2 x_w .. 2 x_w + 63 of output line (2z + pz, 2y + py)
          const size_t row0 = ((size_t(n_w) * Do + (2 * z_w + pz)) * Ho + (2 * y_w + py)) * Wo + 2 * x_w;
          float4 rq[16];
          if (p.residual && ok) {
            // float4 (k, lane) of the warp's 64 output rows: row (k * 32 + lane) / 8, 16-byte column (k * 32 + lane) % 8 of this channel block
            const float* src = p.residual + row0 * p.C_out + cb * CT_CO + 4 * (lane & 7);
#pragma unroll
            for (int k = 0; k < 16; ++k) rq[k] = *reinterpret_cast<const float4*>(src + size_t(4 * k + (lane >> 3)) * p.C_out);
          }
          uint32_t rr[64];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 256 + (2 * cp2) * CT_CO), *reinterpret_cast<uint32_t(*)[32]>(&rr[0]));
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 256 + (2 * cp2 + 1) * CT_CO), *reinterpret_cast<uint32_t(*)[32]>(&rr[32]));
          tc_wait_ld();
          float* mine = tr + lane * CT_TR_PITCH;
#pragma unroll
          for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(mine + j) = make_float4(__uint_as_float(rr[j]) + bias_r[j & 31], __uint_as_float(rr[j + 1]) + bias_r[(j + 1) & 31],
                                                               __uint_as_float(rr[j + 2]) + bias_r[(j + 2) & 31], __uint_as_float(rr[j + 3]) + bias_r[(j + 3) & 31]);
          __syncwarp();
          if (ok) {
            float* dst = p.out32 ? p.out32 + row0 * p.C_out + cb * CT_CO + 4 * (lane & 7) : nullptr;
            // chunk-planar hi | lo copy (the operand of a following halo convolution with folded GroupNorm): an even / odd lane
            // pair holds the 8 channels of one 16-byte chunk; the even lane stores the hi halves, the odd lane the lo halves
            const size_t s_out = size_t(Do) * Ho * Wo;
            __half* pl = nullptr;
            if (p.out_planar) {
              const int chunk = cb * (CT_CO / 8) + ((lane & 7) >> 1) + ((lane & 1) ? p.C_out / 8 : 0);
              pl = p.out_planar + (size_t(n_w) * (2 * p.C_out / 8) + chunk) * s_out * 8 + (row0 - size_t(n_w) * s_out) * 8;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float4 v = *reinterpret_cast<const float4*>(tr + (2 * k + (lane >> 4)) * CT_TR_PITCH + 4 * my_col);
              if (p.residual) v.x += rq[k].x, v.y += rq[k].y, v.z += rq[k].z, v.w += rq[k].w;
              if (dst) *reinterpret_cast<float4*>(dst + size_t(4 * k + (lane >> 3)) * p.C_out) = v;
              if (p.out_planar) {  // (warp-uniform)
                float4 o;
                o.x = __shfl_xor_sync(0xffffffffu, v.x, 1), o.y = __shfl_xor_sync(0xffffffffu, v.y, 1);
                o.z = __shfl_xor_sync(0xffffffffu, v.z, 1), o.w = __shfl_xor_sync(0xffffffffu, v.w, 1);
                const float4 f0 = (lane & 1) ? o : v, f1 = (lane & 1) ? v : o;  // channels 8c .. 8c+3, 8c+4 .. 8c+7
                __align__(16) __half2 h[4];
                h[0] = __floats2half2_rn(f0.x, f0.y), h[1] = __floats2half2_rn(f0.z, f0.w);
                h[2] = __floats2half2_rn(f1.x, f1.y), h[3] = __floats2half2_rn(f1.z, f1.w);
                if (lane & 1) {
                  const float2 a0 = __half22float2(h[0]), a1 = __half22float2(h[1]), a2 = __half22float2(h[2]), a3 = __half22float2(h[3]);
                  h[0] = __floats2half2_rn(f0.x - a0.x, f0.y - a0.y), h[1] = __floats2half2_rn(f0.z - a1.x, f0.w - a1.y);
                  h[2] = __floats2half2_rn(f1.x - a2.x, f1.y - a2.y), h[3] = __floats2half2_rn(f1.z - a3.x, f1.w - a3.y);
                }
                *reinterpret_cast<uint4*>(pl + size_t(4 * k + (lane >> 3)) * 8) = *reinterpret_cast<const uint4*>(h);
              }
              cs += (v.x + v.y) + (v.z + v.w);
              cq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (p.stats) flush_coalesced();
    } else
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int cb, x0, y0, z0, n0;
      decode(t, cb, x0, y0, z0, n0);
      load_bias(cb);
      const int n = n0 + ln, z = z0 + ld, y = y0 + lh, x = x0 + lw;
      const bool ok = n < p.N;
      const int n_w = __shfl_sync(0xffffffffu, n, 0);  // warp-uniform (rows per sample are a multiple of 32, or one sample per tile)
      if (p.stats && n_w != stat_n) {
        flush();
        stat_n = n_w < p.N ? n_w : -1;
      }
      auto ovox = [&](int c) -> size_t {
        return ((size_t(n) * Do + (2 * z + ((c >> 2) & 1))) * Ho + (2 * y + ((c >> 1) & 1))) * Wo + (2 * x + (c & 1));
      };
      // the skip rows do not depend on the accumulator: class 0's row is requested before the MMAs are awaited, class c + 1's
      // while class c is processed
      float4 rs_cur[8], rs_nxt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) rs_cur[j] = rs_nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.residual && ok) {
        const float* row = p.residual + ovox(0) * p.C_out + cb * CT_CO;
#pragma unroll
        for (int j = 0; j < 8; ++j) rs_cur[j] = *reinterpret_cast<const float4*>(row + 4 * j);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        if (p.residual && ok && c + 1 < 8) {
          const float* row = p.residual + ovox(c + 1) * p.C_out + cb * CT_CO;
#pragma unroll
          for (int j = 0; j < 8; ++j) rs_nxt[j] = *reinterpret_cast<const float4*>(row + 4 * j);
        }
        uint32_t rr[32];
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 256 + c * CT_CO), rr);
        tc_wait_ld();
        if (ok) {
          float v[CT_CO];
#pragma unroll
          for (int j = 0; j < CT_CO; j += 4) {
            const float4 b = rs_cur[j >> 2];
            v[j] = __uint_as_float(rr[j]) + bias_r[j] + b.x, v[j + 1] = __uint_as_float(rr[j + 1]) + bias_r[j + 1] + b.y;
            v[j + 2] = __uint_as_float(rr[j + 2]) + bias_r[j + 2] + b.z, v[j + 3] = __uint_as_float(rr[j + 3]) + bias_r[j + 3] + b.w;
          }
          float* o = p.out32 + ovox(c) * p.C_out + cb * CT_CO;
#pragma unroll
          for (int j = 0; j < CT_CO; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.stats) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float ps[8], pq[8];
              group_sums16(*reinterpret_cast<const float(*)[16]>(&v[16 * half]), cpg, ps, pq);
              const int per = cpg >= 16 ? 1 : 16 / cpg;             // groups inside 16 channels
              const int first = (cb * CT_CO + 16 * half) / cpg;    // first group of this half
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
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) rs_cur[j] = rs_nxt[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.stats && !coalesced) flush();
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

extern "C" int semabs_conv_transpose3d_s2(const void* x16, int32_t a_splits, const void* w16, int32_t w_splits, int32_t N, int32_t D,
                                          int32_t H, int32_t W, int32_t C_in, int32_t C_out, int32_t precise, const float* bias,
                                          const float* residual, float* out32, void* out_planar, double* stats, int32_t groups,
                                          void* stream) {
  SB_REQUIRE(x16 && w16 && (out32 || out_planar), "semabs_conv_transpose3d_s2: null pointer");
  SB_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "semabs_conv_transpose3d_s2: bad grid");
  SB_REQUIRE(C_out % CT_CO == 0 && C_in % CT_KB == 0, "semabs_conv_transpose3d_s2: C_in %% 64 == 0 and C_out %% 32 == 0 only (got %d -> %d)", C_in, C_out);
  SB_REQUIRE(!precise || (a_splits == 2 && w_splits == 2), "semabs_conv_transpose3d_s2: precise mode needs hi/lo activations and weights");
  SB_REQUIRE(!stats || (groups > 0 && groups <= 8 && C_out % groups == 0 && (C_out / groups) % 4 == 0 && ((C_out / groups) & (C_out / groups - 1)) == 0),
             "semabs_conv_transpose3d_s2: channels per GroupNorm group must be a power of two >= 4");
  ConvTParams p{};
  p.N = N, p.D = D, p.H = H, p.W = W, p.C_in = C_in, p.C_out = C_out, p.n_cb = C_out / CT_CO;
  if (precise) {
    p.npass = 3;
    p.pass_a[0] = 0, p.pass_w[0] = 0, p.pass_a[1] = 1, p.pass_w[1] = 0, p.pass_a[2] = 0, p.pass_w[2] = 1;
  } else {
    p.npass = 1, p.pass_a[0] = 0, p.pass_w[0] = 0;
  }
  p.bw = W < 128 ? W : 128;
  p.bh = H < 128 / p.bw ? H : 128 / p.bw;
  p.bd = D < 128 / (p.bw * p.bh) ? D : 128 / (p.bw * p.bh);
  p.bn = 128 / (p.bw * p.bh * p.bd);
  SB_REQUIRE(p.bw * p.bh * p.bd * p.bn == 128 && W % p.bw == 0 && H % p.bh == 0 && D % p.bd == 0,
             "semabs_conv_transpose3d_s2: grid %dx%dx%d cannot be tiled into 128-voxel boxes", D, H, W);
  SB_REQUIRE(p.bn == 1 || (p.bw * p.bh * p.bd) % 32 == 0, "semabs_conv_transpose3d_s2: grid too small (needs >= 32 voxels)");
  p.tiles_w = W / p.bw, p.tiles_h = H / p.bh, p.tiles_d = D / p.bd, p.tiles_n = (N + p.bn - 1) / p.bn;
  p.bias = bias, p.residual = residual, p.out32 = out32, p.out_planar = (__half*)out_planar, p.stats = stats, p.groups = groups;
  SB_REQUIRE(!out_planar || (p.bw % 32 == 0 && C_out % 8 == 0), "semabs_conv_transpose3d_s2: the planar output needs a grid width that is a multiple of 32");
  SB_REQUIRE(out32 || p.bw % 32 == 0, "semabs_conv_transpose3d_s2: out32 missing");
  CUtensorMap tmA, tmB;
  {
    const uint64_t C = uint64_t(a_splits) * C_in;
    uint64_t dims[5] = {C, uint64_t(W), uint64_t(H), uint64_t(D), uint64_t(N)};
    uint64_t str[4] = {C * 2, C * 2 * W, C * 2 * W * H, C * 2 * W * H * D};
    uint32_t box[5] = {uint32_t(CT_KB), uint32_t(p.bw), uint32_t(p.bh), uint32_t(p.bd), uint32_t(p.bn)};
    if (int rc = make_tmap_f16(&tmA, x16, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    const uint64_t K = uint64_t(w_splits) * 27 * C_in;
    uint64_t dims[2] = {K, uint64_t(C_out)};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {uint32_t(CT_KB), uint32_t(CT_CO)};
    if (int rc = make_tmap_f16(&tmB, w16, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(convt_allparity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    configured = true;
  }
  const int num_tiles = p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * p.n_cb;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  convt_allparity_kernel<<<grid, CT_THREADS, CT_SMEM, (cudaStream_t)stream>>>(tmA, tmB, p);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
