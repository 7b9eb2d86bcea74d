// Halo-resident 3x3x3 convolution for the full-resolution UNet level (W = 128, C in {16, 32}) on tcgen05.
//
// conv3d.cu re-loads the activation box once per tap (27 x) and per precision pass (3 x): at 128^3 that makes the
// level-0 convolutions L2-bandwidth bound (measured 5.5 TB/s L2->SM, 9 ms per conv at N=4, C=32).  Here each input
// voxel row is loaded ONCE per output row that needs it in (y) and reused along z through a sliding window:
//
//   * activations are "chunk-planar" fp16: [n][chunk][z][y][x][8 channels] (chunk = 8 channels = 16 B; hi chunks
//     then lo chunks), so a TMA box (8, 130, 3) is one z-plane of a chunk: 3 y-rows x (128 + 2 halo) voxels, with
//     out-of-bounds voxels zero-filled (= the conv's zero padding);
//   * in shared memory a plane is [chunk][3 rows][130 voxels][16 B]: for one 16 B chunk the voxels of a row are
//     contiguous at 16 B pitch, which is exactly the NO-SWIZZLE K-major UMMA operand layout (8-row core matrices of
//     128 B, SBO = 128 B between row groups, LBO = chunk pitch between the two 8-channel halves of K = 16).  Every
//     tap (dy, dx) is therefore just a different descriptor START ADDRESS into the same plane — no data movement;
//   * a CTA walks a column of output rows along z for fixed (n, y): ring of 4 planes (z-1, z, z+1 + one in flight),
//     one new plane per output row => each activation byte crosses L2->SM 3x (once per y neighbour) instead of 81x;
//   * weights stream per tap through a small ring as pre-packed core-matrix images (cp.async.bulk).
// Warp roles / TMEM double buffering / epilogue are those of conv3d.cu.
#include "../../include/semabs_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace sb {

constexpr int HALO_THREADS = 256;
constexpr int HALO_W = 128;
constexpr int HALO_XP = HALO_W + 2;                 // voxels per row incl. halo
constexpr int HALO_CHUNK_DATA = 3 * HALO_XP * 16;   // one chunk of one plane: 3 rows x 130 voxels x 16 B = 6240
constexpr int HALO_CHUNK_BYTES = (HALO_CHUNK_DATA + 127) / 128 * 128;  // smem pitch (TMA destinations 128 B aligned)
constexpr int HALO_PLANES = 4;
constexpr int HALO_WSTAGES = 6;

struct HaloParams {
  int N, D, H;
  int C_in, C_out;
  int nchunks;            // a_splits * C_in / 8
  int npass;
  int8_t pass_a[3], pass_w[3];
  int w_splits;
  int zseg, nseg;         // output rows per work item along z, segments per column
  int plane_bytes;        // nchunks * HALO_CHUNK_BYTES
  int wtap_bytes;         // w_splits * C_in * C_out * 2
  const float* residual;
  int relu;
  float* out32;
  __half* out16;
  int o16_splits;
  double* stats;
  int groups;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int BN>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __half* __restrict__ wimg,
                   const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* planes = smem;                                         // HALO_PLANES x plane_bytes
  uint8_t* wring = planes + HALO_PLANES * p.plane_bytes;          // HALO_WSTAGES x wtap_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + HALO_WSTAGES * p.wtap_bytes);
  uint64_t* plane_full = bars;
  uint64_t* plane_empty = plane_full + HALO_PLANES;
  uint64_t* w_full = plane_empty + HALO_PLANES;
  uint64_t* w_empty = w_full + HALO_WSTAGES;
  uint64_t* tmem_full = w_empty + HALO_WSTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.N * p.H * p.nseg;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HALO_PLANES; ++s) mbar_init(&plane_full[s], 1), mbar_init(&plane_empty[s], 1);
    for (int s = 0; s < HALO_WSTAGES; ++s) mbar_init(&w_full[s], 1), mbar_init(&w_empty[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tmem_full[a], 1), mbar_init(&tmem_empty[a], 4);
    fence_barrier_init();
  }
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& n, int& y, int& z0) {
    const int zs = item % p.nseg;
    y = (item / p.nseg) % p.H;
    n = item / (p.nseg * p.H);
    z0 = zs * p.zseg;
  };

  if (warp == 0) {
    // ===== producer: activation planes =====
    if (lane == 0) {
      uint32_t pc = 0;  // running plane counter -> ring slot / phase
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int n, y, z0;
        decode(item, n, y, z0);
        for (int k = 0; k < p.zseg + 2; ++k, ++pc) {
          const int slot = pc % HALO_PLANES;
          const uint32_t phase = (pc / HALO_PLANES) & 1;
          mbar_wait(&plane_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&plane_full[slot], p.nchunks * HALO_CHUNK_DATA);
          uint8_t* dst = planes + slot * p.plane_bytes;
          for (int c = 0; c < p.nchunks; ++c)
            tma_load_5d(dst + c * HALO_CHUNK_BYTES, &tmA, &plane_full[slot], 0, -1, y - 1, z0 - 1 + k, n * p.nchunks + c);
        }
      }
    }
  } else if (warp == 3) {
    // ===== producer: weight taps =====
    if (lane == 0) {
      uint32_t wc = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int i = 0; i < p.zseg; ++i) {
          for (int tp = 0; tp < 27; ++tp, ++wc) {
            const int st = wc % HALO_WSTAGES;
            const uint32_t phase = (wc / HALO_WSTAGES) & 1;
            mbar_wait(&w_empty[st], phase ^ 1);
            mbar_arrive_expect_tx(&w_full[st], p.wtap_bytes);
            bulk_load(wring + st * p.wtap_bytes, reinterpret_cast<const uint8_t*>(wimg) + size_t(tp) * p.wtap_bytes,
                      p.wtap_bytes, &w_full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, BN);
      const int ksteps = p.C_in / 16;
      const uint32_t wsplit_bytes = uint32_t(p.C_in) * p.C_out * 2;
      const uint32_t wk_bytes = uint32_t(p.C_out / 8) * 256;  // one K=16 block of B: [C_out/8 groups][2 k-chunks][128 B]
      uint32_t pc = 0, wc = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        // planes pc+0 (z0-1) and pc+1 (z0) must have landed before the first row; plane pc+i+2 before dz=+1 of row i
        for (int k = 0; k < 2; ++k) mbar_wait(&plane_full[(pc + k) % HALO_PLANES], ((pc + k) / HALO_PLANES) & 1);
        for (int i = 0; i < p.zseg; ++i) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          bool first = true;
          for (int dz = 0; dz < 3; ++dz) {
            const uint32_t pidx = pc + i + dz;
            if (dz == 2) mbar_wait(&plane_full[pidx % HALO_PLANES], (pidx / HALO_PLANES) & 1);
            tc_fence_after();
            const uint32_t plane_addr = smem_u32(planes + (pidx % HALO_PLANES) * p.plane_bytes);
            for (int dy = 0; dy < 3; ++dy) {
              for (int dx = 0; dx < 3; ++dx, ++wc) {
                const int st = wc % HALO_WSTAGES;
                mbar_wait(&w_full[st], (wc / HALO_WSTAGES) & 1);
                tc_fence_after();
                const uint32_t w_addr = smem_u32(wring + st * p.wtap_bytes);
                const uint32_t tap_off = uint32_t(dy * HALO_XP + dx) * 16u;
                for (int ps = 0; ps < p.npass; ++ps) {
                  const uint32_t a_chunk0 = uint32_t(p.pass_a[ps]) * (p.C_in / 8);
                  const uint32_t wb = w_addr + uint32_t(p.pass_w[ps]) * wsplit_bytes;
                  for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t a_addr = plane_addr + (a_chunk0 + 2 * ks) * HALO_CHUNK_BYTES + tap_off;
                    // A: rows = voxels at 16 B pitch (8-row core matrices of 128 B), K halves HALO_CHUNK_BYTES apart
                    const uint64_t da = make_smem_desc(a_addr, HALO_CHUNK_BYTES, 128, SW_NONE);
                    // B: [C_out/8 groups (256 B)][2 k-chunks (128 B)][8 rows x 16 B]
                    const uint64_t db = make_smem_desc(wb + ks * wk_bytes, 128, 256, SW_NONE);
                    umma_f16(d_tmem, da, db, idesc, first ? 0u : 1u);
                    first = false;
                  }
                }
                umma_commit(&w_empty[st]);
              }
            }
          }
          umma_commit(&tmem_full[acc]);
          umma_commit(&plane_empty[(pc + i) % HALO_PLANES]);  // plane z-1 is not needed by later rows
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        // the last two planes of the column are not reused by this CTA
        umma_commit(&plane_empty[(pc + p.zseg) % HALO_PLANES]);
        umma_commit(&plane_empty[(pc + p.zseg + 1) % HALO_PLANES]);
        pc += p.zseg + 2;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = voxel x of the output row =====
    const int q = warp & 3;
    const int x = q * 32 + lane;
    constexpr int MAXG = 8;
    float gs[MAXG], gq[MAXG];
#pragma unroll
    for (int i = 0; i < MAXG; ++i) gs[i] = gq[i] = 0.f;
    int stat_n = -1;
    const int cpg = p.stats ? p.C_out / p.groups : 1;
    const int ngroups = p.stats ? p.groups : 0;
    auto flush = [&]() {
      if (stat_n < 0) return;
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        if (i < ngroups) {
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
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int n, y, z0;
      decode(item, n, y, z0);
      if (p.stats && n != stat_n) {
        flush();
        stat_n = n;
      }
      for (int i = 0; i < p.zseg; ++i) {
        const int z = z0 + i;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const size_t ovox = ((size_t(n) * p.D + z) * p.H + y) * HALO_W + x;
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t rr[16];
          tmem_ld_32x32b_x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + c * 16), rr);
          tc_wait_ld();
          const int col0 = c * 16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
          if (p.residual) {
            const float* rs = p.residual + ovox * p.C_out + col0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(rs + j);
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
            // channels per group is a power of two in {2,4,8,16,32}: groups inside this 16-channel chunk
            const int per = cpg >= 16 ? 1 : 16 / cpg;
#pragma unroll
            for (int i = 0; i < MAXG; ++i) {
              const int gfirst = (c * 16) / cpg;  // first group touched by this chunk
              const int lg = i - gfirst;
              if (lg >= 0 && lg < per) {
                float s = 0.f, s2 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (cpg >= 16 || j / cpg == lg) s += v[j], s2 += v[j] * v[j];
                gs[i] += s, gq[i] += s2;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    if (p.stats) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_conv3d_halo(const void* x16_planar, int32_t a_splits, const void* w_img, int32_t w_splits,
                                  int32_t N, int32_t D, int32_t H, int32_t W, int32_t C_in, int32_t C_out,
                                  int32_t precise, const float* residual, int32_t relu, float* out32, void* out16,
                                  int32_t o16_splits, double* stats, int32_t groups, void* stream) {
  SB_REQUIRE(x16_planar && w_img && (out32 || out16), "semabs_conv3d_halo: null pointer");
  SB_REQUIRE(W == HALO_W, "semabs_conv3d_halo: W must be %d (got %d)", HALO_W, W);
  SB_REQUIRE((C_in == 16 || C_in == 32) && (C_out == 16 || C_out == 32), "semabs_conv3d_halo: C_in/C_out must be 16 or 32");
  SB_REQUIRE(!precise || (a_splits == 2 && w_splits == 2), "semabs_conv3d_halo: precise mode needs hi/lo operands");
  SB_REQUIRE(!stats || (groups >= 1 && groups <= 8 && C_out % groups == 0 && (C_out / groups) >= 2),
             "semabs_conv3d_halo: bad GroupNorm groups");
  SB_REQUIRE(N > 0 && D > 0 && H > 0, "semabs_conv3d_halo: bad grid");
  HaloParams p{};
  p.N = N, p.D = D, p.H = H, p.C_in = C_in, p.C_out = C_out;
  p.nchunks = a_splits * C_in / 8;
  p.w_splits = w_splits;
  if (precise) {
    p.npass = 3;
    p.pass_a[0] = 0, p.pass_w[0] = 0, p.pass_a[1] = 1, p.pass_w[1] = 0, p.pass_a[2] = 0, p.pass_w[2] = 1;
  } else {
    p.npass = 1, p.pass_a[0] = 0, p.pass_w[0] = 0;
  }
  // work item = (n, y, z segment); enough items for ~7 rounds over the SMs, segments not shorter than 16 rows
  int zseg = D;
  while (zseg > 16 && (long long)N * H * (D / zseg) < 6LL * num_sms() && zseg % 2 == 0) zseg /= 2;
  p.zseg = zseg, p.nseg = D / zseg;
  SB_REQUIRE(p.zseg * p.nseg == D, "semabs_conv3d_halo: D=%d not divisible into z segments", D);
  p.plane_bytes = p.nchunks * HALO_CHUNK_BYTES;
  p.wtap_bytes = w_splits * C_in * C_out * 2;
  p.residual = residual, p.relu = relu, p.out32 = out32, p.out16 = (__half*)out16, p.o16_splits = o16_splits;
  p.stats = stats, p.groups = groups;
  const size_t smem = size_t(HALO_PLANES) * p.plane_bytes + size_t(HALO_WSTAGES) * p.wtap_bytes + 512 + 128;
  SB_REQUIRE(smem <= 227 * 1024, "semabs_conv3d_halo: %zu bytes of shared memory needed", smem);

  CUtensorMap tmA;
  {
    // [n*nchunks][z][y][x][8] fp16
    uint64_t dims[5] = {8, uint64_t(W), uint64_t(H), uint64_t(D), uint64_t(N) * p.nchunks};
    uint64_t str[4] = {16, 16ull * W, 16ull * W * H, 16ull * W * H * D};
    uint32_t box[5] = {8, uint32_t(HALO_XP), 3, 1, 1};
    if (int rc = make_tmap_f16(&tmA, x16_planar, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int items = N * H * p.nseg;
  const int grid = items < num_sms() ? items : num_sms();
  if (C_out == 32) {
    static bool cfg = false;
    if (!cfg) {
      SB_CHECK_CUDA(cudaFuncSetAttribute(conv3d_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      cfg = true;
    }
    conv3d_halo_kernel<32><<<grid, HALO_THREADS, smem, st>>>(tmA, (const __half*)w_img, p);
  } else {
    static bool cfg = false;
    if (!cfg) {
      SB_CHECK_CUDA(cudaFuncSetAttribute(conv3d_halo_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      cfg = true;
    }
    conv3d_halo_kernel<16><<<grid, HALO_THREADS, smem, st>>>(tmA, (const __half*)w_img, p);
  }
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
