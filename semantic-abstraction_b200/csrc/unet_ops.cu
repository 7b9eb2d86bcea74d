// Bandwidth-bound stages of the residual 3-D UNet (reference unet3d.py): GroupNorm apply (statistics come from the
// producer's epilogue, see conv3d.cu), 2x2x2 max-pool with fused statistics for the next GroupNorm
// (Encoder.forward, unet3d.py:313-317), and the NCDHW <-> channels-last conversions at the module boundary.
// All kernels: grid (x, N); a thread owns one channel quad (float4) so its GroupNorm group is loop-invariant and
// statistics are reduced registers -> shared (fp32) -> global (fp64 atomics, a few per CTA).
#include "../../include/semabs_b200.h"
#include <cuda_fp16.h>
#include "common.cuh"

namespace sb {

constexpr int EW_THREADS = 256;

struct QuadStats {
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;  // channels (0,1) and (2,3) of the quad
  __device__ __forceinline__ void add(const float4& v) {
    s0 += v.x + v.y, q0 += v.x * v.x + v.y * v.y;
    s1 += v.z + v.w, q1 += v.z * v.z + v.w * v.w;
  }
};

// reduce the per-thread quad statistics of a CTA into stats[n, g, 2]
__device__ __forceinline__ void flush_quad_stats(const QuadStats& qs, int cq, int cpg, int groups, double* stats_n) {
  __shared__ float sm[16];
  if (threadIdx.x < 16) sm[threadIdx.x] = 0.f;
  __syncthreads();
  const int g0 = (4 * cq) / cpg, g1 = (4 * cq + 2) / cpg;
  // warp-level pre-reduction when the whole warp maps to one group
  atomicAdd(&sm[2 * g0], qs.s0);
  atomicAdd(&sm[2 * g0 + 1], qs.q0);
  atomicAdd(&sm[2 * g1], qs.s1);
  atomicAdd(&sm[2 * g1 + 1], qs.q1);
  __syncthreads();
  if (threadIdx.x < 2 * groups) atomicAdd(stats_n + threadIdx.x, double(sm[threadIdx.x]));
}

// ---- NCDHW fp32 -> channels-last fp32 (channel-padded with zeros) + statistics -----------------------
__global__ void __launch_bounds__(EW_THREADS)
ncdhw_to_ndhwc_kernel(const float* __restrict__ x, float* __restrict__ y, long long S, int C, int Cpad, int groups,
                      double* __restrict__ stats) {
  const int n = blockIdx.y;
  const int qpc = Cpad / 4;               // quads per voxel
  const int vpb = EW_THREADS / qpc;       // voxels per CTA iteration
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc;
  const float* xn = x + size_t(n) * C * S;
  float* yn = y + size_t(n) * S * Cpad;
  QuadStats qs;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    float4 o;
    const int c = 4 * cq;
    o.x = c + 0 < C ? xn[size_t(c + 0) * S + v] : 0.f;
    o.y = c + 1 < C ? xn[size_t(c + 1) * S + v] : 0.f;
    o.z = c + 2 < C ? xn[size_t(c + 2) * S + v] : 0.f;
    o.w = c + 3 < C ? xn[size_t(c + 3) * S + v] : 0.f;
    *reinterpret_cast<float4*>(yn + size_t(v) * Cpad + c) = o;
    qs.add(o);
  }
  if (stats) {
    const int cpg = groups == 1 ? Cpad : C / groups;
    flush_quad_stats(qs, cq, cpg < 2 ? 2 : cpg, groups, stats + size_t(n) * groups * 2);
  }
}

// ---- NCDHW fp32 -> chunk-planar fp16 hi | lo [N][2 Cpad / 8][S][8] + statistics ------------------------------
// The operand layout of the halo convolution, written straight from the module input (inference with the first GroupNorm folded
// into conv1): no channels-last fp32 copy, no GroupNorm-apply pass.  Thread = (voxel, 8-channel chunk), voxels fastest across
// the CTA so that both the plane reads and the 16-byte chunk writes are contiguous.
__global__ void __launch_bounds__(EW_THREADS)
ncdhw_to_planar_kernel(const float* __restrict__ x, __half* __restrict__ y, long long S, int C, int Cpad, int groups,
                       double* __restrict__ stats) {
  __shared__ float sm[16];
  if (threadIdx.x < 16) sm[threadIdx.x] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int qpc = Cpad / 8;            // chunks per voxel
  const int vpb = EW_THREADS / qpc;    // voxels per CTA iteration
  const int vl = threadIdx.x % vpb, cq = threadIdx.x / vpb, c = 8 * cq;
  const float* xn = x + size_t(n) * C * S;
  __half* hi = y + (size_t(n) * 2 * qpc + cq) * S * 8;
  __half* lo = y + (size_t(n) * 2 * qpc + qpc + cq) * S * 8;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;  // channels c .. c+3 and c+4 .. c+7
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = c + j < C ? xn[size_t(c + j) * S + v] : 0.f;
    __align__(16) __half2 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
      const float2 b = __half22float2(h[j]);
      l[j] = __floats2half2_rn(f[2 * j] - b.x, f[2 * j + 1] - b.y);
    }
    *reinterpret_cast<uint4*>(hi + v * 8) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + v * 8) = *reinterpret_cast<const uint4*>(l);
    s0 += (f[0] + f[1]) + (f[2] + f[3]), q0 += (f[0] * f[0] + f[1] * f[1]) + (f[2] * f[2] + f[3] * f[3]);
    s1 += (f[4] + f[5]) + (f[6] + f[7]), q1 += (f[4] * f[4] + f[5] * f[5]) + (f[6] * f[6] + f[7] * f[7]);
  }
  if (stats) {
    const int cpg = groups == 1 ? Cpad : C / groups;  // (a multiple of 4: the host checks)
    const int g0 = min(c / cpg, groups - 1), g1 = min((c + 4) / cpg, groups - 1);
    atomicAdd(&sm[2 * g0], s0), atomicAdd(&sm[2 * g0 + 1], q0);
    atomicAdd(&sm[2 * g1], s1), atomicAdd(&sm[2 * g1 + 1], q1);
    __syncthreads();
    if (threadIdx.x < 2 * groups) atomicAdd(stats + size_t(n) * groups * 2 + threadIdx.x, double(sm[threadIdx.x]));
  }
}

// ---- channels-last fp32 -> NCDHW fp32 -------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
ndhwc_to_ncdhw_kernel(const float* __restrict__ x, float* __restrict__ y, long long S, int C) {
  // tile transpose through shared memory: 32 voxels x 32 channels
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long v0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  for (int r = ty; r < 32; r += 8) {
    const long long v = v0 + r;
    tile[r][tx] = (v < S && c0 + tx < C) ? x[(size_t(n) * S + v) * C + c0 + tx] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long v = v0 + tx;
    if (c < C && v < S) y[(size_t(n) * C + c) * S + v] = tile[tx][r];
  }
}

// ---- final 1x1x1 convolution fused with the channels-last -> NCDHW conversion (unet3d.py:565, 619) -------------
// x16 [N, S, splits * CIN] fp16 rows [hi | lo] (the last decoder block's output as its epilogue wrote it), w [C_out, CIN] fp32,
// y [N, C_out, S] fp32.  Thread = voxel: its input row in registers (hi + lo = the fp32 value to ~22 bits), weights broadcast
// from shared memory, one coalesced 128-byte store per warp and output channel.  Replaces an implicit-GEMM launch that wrote
// fp32 channels-last (1 GB at 4 x 128^3 x 32) plus a transpose pass that read it back: 3 GB -> 1.5 GB of DRAM traffic, and the
// products run in fp32 FMAs instead of three fp16 MMAs.
template <int CIN>
__global__ void __launch_bounds__(256) final_conv1x1_ncdhw_kernel(const __half* __restrict__ x16, int splits, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, float* __restrict__ y, long long S,
                                                                  int C_out) {
  extern __shared__ float s_w[];  // [C_out][CIN] then bias [C_out]
  for (int i = threadIdx.x; i < C_out * CIN; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < C_out; i += blockDim.x) s_w[C_out * CIN + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  // persistent blocks (the weights are staged once per block), two voxels per thread and step: every broadcast weight load
  // feeds two FMAs.  v and v + 256 so that the plane stores of both stay coalesced.
  for (long long v0 = (long long)blockIdx.x * 512 + threadIdx.x; v0 < S; v0 += (long long)gridDim.x * 512) {
    const long long v1 = v0 + 256;
    const bool two = v1 < S;
    float x[2][CIN];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const __half* row = x16 + (size_t(n) * S + (u == 0 || two ? v0 + 256 * u : v0)) * size_t(splits) * CIN;
#pragma unroll
      for (int c = 0; c < CIN; c += 8) {
        const uint4 uh = *reinterpret_cast<const uint4*>(row + c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&uh);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          x[u][c + 2 * e] = f.x, x[u][c + 2 * e + 1] = f.y;
        }
        if (splits == 2) {
          const uint4 ul = *reinterpret_cast<const uint4*>(row + CIN + c);
          const __half2* l2 = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(l2[e]);
            x[u][c + 2 * e] += f.x, x[u][c + 2 * e + 1] += f.y;
          }
        }
      }
    }
    float* yn = y + size_t(n) * C_out * S + v0;
    for (int o = 0; o < C_out; ++o) {
      const float4* wo = reinterpret_cast<const float4*>(s_w + o * CIN);
      const float bo = s_w[C_out * CIN + o];
      float a0 = bo, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = bo, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
      for (int c = 0; c < CIN / 4; ++c) {
        const float4 w4 = wo[c];
        a0 = fmaf(w4.x, x[0][4 * c], a0), a1 = fmaf(w4.y, x[0][4 * c + 1], a1);
        a2 = fmaf(w4.z, x[0][4 * c + 2], a2), a3 = fmaf(w4.w, x[0][4 * c + 3], a3);
        b0 = fmaf(w4.x, x[1][4 * c], b0), b1 = fmaf(w4.y, x[1][4 * c + 1], b1);
        b2 = fmaf(w4.z, x[1][4 * c + 2], b2), b3 = fmaf(w4.w, x[1][4 * c + 3], b3);
      }
      yn[size_t(o) * S] = (a0 + a1) + (a2 + a3);
      if (two) yn[size_t(o) * S + 256] = (b0 + b1) + (b2 + b3);
    }
  }
}

// ---- GroupNorm folded into the consuming 3x3x3 convolution (32 -> 32 channels, inference, see conv3d_halo2.cu) -------------
// GroupNorm(x)[c] = a_c x_c + b_c per (sample, channel) with a = gamma rstd, b = beta - mean a, so
//   conv_W(GroupNorm(x)) = conv_{W a}(x) + sum over the taps that fall inside the grid of (W b).
// One launch writes, for every sample, the halo kernel's resident weight image of W a (ops.pack_halo_weights layout, hi | lo
// rows) and the 27-entry border-class bias table.  Statistics arithmetic = gn_apply_kernel's (fp64 mean / variance, eps 1e-5).
constexpr int FOLD_C = 32;
constexpr int FOLD_IMG_HALFS = 2 * 27 * 2 * 4 * 2 * 8 * 8;  // [half][tap][kb][row group][kc][row][elem]
__global__ void __launch_bounds__(256) fold_groupnorm_halo_kernel(const float* __restrict__ w /*[32][32][27]*/, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const double* __restrict__ stats,
                                                                  int stats_stride, int groups, double inv_count, __half* __restrict__ img,
                                                                  float* __restrict__ bias_cls /*[N][27][32]*/) {
  __shared__ float s_a[FOLD_C], s_b[FOLD_C];
  __shared__ float s_wb[FOLD_C * 27];  // sum_ci W[co][ci][tap] b[ci]
  const int n = blockIdx.x, cpg = FOLD_C / groups;
  if (threadIdx.x < FOLD_C) {
    const int c = threadIdx.x, g = c / cpg;
    const double mean = stats[size_t(n) * stats_stride + 2 * g] * inv_count;
    double var = stats[size_t(n) * stats_stride + 2 * g + 1] * inv_count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = float(1.0 / sqrt(var + 1e-5));
    const float a = rstd * gamma[c];
    s_a[c] = a, s_b[c] = beta[c] - float(mean) * a;
  }
  __syncthreads();
  __half* im = img + size_t(n) * FOLD_IMG_HALFS;
  // weight image: the (co, ci, tap) index space in gridDim.y slices; the bias table is block (n, 0)'s
  const int total = FOLD_C * FOLD_C * 27, per = (total + gridDim.y - 1) / gridDim.y;
  const int lo_idx = blockIdx.y * per, hi_idx = min(total, lo_idx + per);
  for (int idx = lo_idx + threadIdx.x; idx < hi_idx; idx += blockDim.x) {
    const int tap = idx % 27, ci = (idx / 27) % FOLD_C, co = idx / (27 * FOLD_C);
    const float v = w[idx] * s_a[ci];
    const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
    const int h = co >> 4, r16 = co & 15, kb = ci >> 4, kc = (ci & 15) >> 3, e = ci & 7;
    const size_t base = (size_t(h) * 27 + tap) * 2 + kb;
    // rows 0-15 of a half: W_hi of its 16 output channels, rows 16-31: W_lo
    im[((((base * 4 + (r16 >> 3)) * 2 + kc) * 8) + (r16 & 7)) * 8 + e] = hi;
    im[((((base * 4 + 2 + (r16 >> 3)) * 2 + kc) * 8) + (r16 & 7)) * 8 + e] = lo;
  }
  if (blockIdx.y != 0) return;
  for (int idx = threadIdx.x; idx < FOLD_C * 27; idx += blockDim.x) {
    const int tap = idx % 27, co = idx / 27;
    float acc = 0.f;
    for (int ci = 0; ci < FOLD_C; ++ci) acc = fmaf(w[(co * FOLD_C + ci) * 27 + tap], s_b[ci], acc);
    s_wb[idx] = acc;
  }
  __syncthreads();
  // border classes per axis: 0 = first voxel (taps 1, 2 inside), 1 = interior (all), 2 = last voxel (taps 0, 1)
  for (int idx = threadIdx.x; idx < 27 * FOLD_C; idx += blockDim.x) {
    const int co = idx % FOLD_C, cls = idx / FOLD_C;
    const int cz = cls / 9, cy = (cls / 3) % 3, cx = cls % 3;
    float acc = 0.f;
    for (int tz = 0; tz < 3; ++tz) {
      if ((cz == 0 && tz == 0) || (cz == 2 && tz == 2)) continue;
      for (int ty = 0; ty < 3; ++ty) {
        if ((cy == 0 && ty == 0) || (cy == 2 && ty == 2)) continue;
        for (int tx = 0; tx < 3; ++tx) {
          if ((cx == 0 && tx == 0) || (cx == 2 && tx == 2)) continue;
          acc += s_wb[co * 27 + (tz * 3 + ty) * 3 + tx];
        }
      }
    }
    bias_cls[(size_t(n) * 27 + cls) * FOLD_C + co] = acc;
  }
}

// ---- GroupNorm apply: raw fp32 channels-last + (sum, sumsq) -> normalised fp16 (hi | lo) ------------------
// torch.nn.GroupNorm semantics (unet3d.py:78-83): biased variance over (C/G) x D x H x W, eps 1e-5, affine.
__global__ void __launch_bounds__(EW_THREADS)
gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, __half* __restrict__ y, long long S, int C, int groups, int cpg,
                double inv_count, int splits, int planar) {
  __shared__ float s_scale[1024], s_shift[1024];
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = min(c / cpg, groups - 1);
    const double mean = stats[(size_t(n) * groups + g) * 2] * inv_count;
    double var = stats[(size_t(n) * groups + g) * 2 + 1] * inv_count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = float(1.0 / sqrt(var + 1e-5));
    const float sc = rstd * gamma[c];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - float(mean) * sc;
  }
  __syncthreads();
  const int qpc = C / 4, vpb = EW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float4 sc = make_float4(s_scale[c], s_scale[c + 1], s_scale[c + 2], s_scale[c + 3]);
  const float4 sh = make_float4(s_shift[c], s_shift[c + 1], s_shift[c + 2], s_shift[c + 3]);
  const float* xn = x + size_t(n) * S * C;
  __half* yn = y + size_t(n) * S * splits * C;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    const float4 a = *reinterpret_cast<const float4*>(xn + size_t(v) * C + c);
    const float o0 = fmaf(a.x, sc.x, sh.x), o1 = fmaf(a.y, sc.y, sh.y), o2 = fmaf(a.z, sc.z, sh.z), o3 = fmaf(a.w, sc.w, sh.w);
    const __half2 h0 = __floats2half2_rn(o0, o1), h1 = __floats2half2_rn(o2, o3);
    // channels-last [v][splits*C], or chunk-planar [chunk][v][8] (chunk = 8 channels; hi chunks then lo chunks) for
    // the halo-resident level-0 convolution (conv3d_halo.cu)
    __half* dst = planar ? yn + (size_t(c >> 3) * S + v) * 8 + (c & 7) : yn + size_t(v) * splits * C + c;
    *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    if (splits == 2) {
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      const __half2 l0 = __floats2half2_rn(o0 - f0.x, o1 - f0.y), l1 = __floats2half2_rn(o2 - f1.x, o3 - f1.y);
      __half* dlo = planar ? dst + size_t(C >> 3) * S * 8 : dst + C;
      *reinterpret_cast<uint2*>(dlo) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
  }
}

// ---- MaxPool3d(2) + statistics of the pooled tensor --------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
maxpool2_kernel(const float* __restrict__ x, float* __restrict__ y, int D, int H, int W, int C, int groups,
                double* __restrict__ stats) {
  const int n = blockIdx.y;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long So = (long long)Do * Ho * Wo;
  const int qpc = C / 4, vpb = EW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float* xn = x + size_t(n) * D * H * W * C;
  float* yn = y + size_t(n) * So * C;
  QuadStats qs;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < So; v += (long long)gridDim.x * vpb) {
    const int xo = int(v % Wo), yo = int((v / Wo) % Ho), zo = int(v / ((long long)Wo * Ho));
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t src = ((size_t(2 * zo + dz) * H + (2 * yo + dy)) * W + (2 * xo + dx)) * C + c;
          const float4 a = *reinterpret_cast<const float4*>(xn + src);
          m.x = fmaxf(m.x, a.x), m.y = fmaxf(m.y, a.y), m.z = fmaxf(m.z, a.z), m.w = fmaxf(m.w, a.w);
        }
    *reinterpret_cast<float4*>(yn + size_t(v) * C + c) = m;
    qs.add(m);
  }
  if (stats) flush_quad_stats(qs, cq, C / groups, groups, stats + size_t(n) * groups * 2);
}

static int ew_grid(long long S, int vpb) {
  long long need = (S + vpb - 1) / vpb;
  long long cap = (long long)num_sms() * 8;
  return int(need < cap ? need : cap);
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_ncdhw_to_ndhwc(const float* x, float* y, int32_t N, int64_t S, int32_t C, int32_t Cpad,
                                     int32_t groups, double* stats, void* stream) {
  SB_REQUIRE(x && y && N > 0 && S > 0 && C > 0 && Cpad >= C && Cpad % 4 == 0 && (EW_THREADS % (Cpad / 4)) == 0,
             "semabs_ncdhw_to_ndhwc: bad arguments (Cpad=%d)", Cpad);
  SB_REQUIRE(!stats || (groups >= 1 && groups <= 8 && (groups == 1 || (C == Cpad && C % groups == 0 && (C / groups) % 2 == 0))),
             "semabs_ncdhw_to_ndhwc: unsupported GroupNorm grouping C=%d groups=%d", C, groups);
  const int vpb = EW_THREADS / (Cpad / 4);
  dim3 grid(ew_grid(S, vpb), N);
  ncdhw_to_ndhwc_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, y, S, C, Cpad, groups, stats);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_ncdhw_to_planar(const float* x, void* y16, int32_t N, int64_t S, int32_t C, int32_t Cpad, int32_t groups,
                                      double* stats, void* stream) {
  SB_REQUIRE(x && y16 && N > 0 && S > 0 && C > 0 && Cpad >= C, "semabs_ncdhw_to_planar: bad arguments");
  SB_REQUIRE(Cpad % 8 == 0 && EW_THREADS % (Cpad / 8) == 0 && Cpad <= 64, "semabs_ncdhw_to_planar: unsupported channel count %d", Cpad);
  SB_REQUIRE(!stats || (groups >= 1 && groups <= 8 && (groups == 1 || (C % groups == 0 && (C / groups) % 4 == 0))),
             "semabs_ncdhw_to_planar: unsupported GroupNorm grouping");
  const int vpb = EW_THREADS / (Cpad / 8);
  dim3 grid(ew_grid(S, vpb), N);
  ncdhw_to_planar_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, (__half*)y16, S, C, Cpad, groups, stats);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_ndhwc_to_ncdhw(const float* x, float* y, int32_t N, int64_t S, int32_t C, void* stream) {
  SB_REQUIRE(x && y && N > 0 && S > 0 && C > 0, "semabs_ndhwc_to_ncdhw: bad arguments");
  dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, N);
  ndhwc_to_ncdhw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, S, C);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_final_conv1x1_ncdhw(const void* x16, int32_t splits, const float* w, const float* bias, float* y, int32_t N,
                                          int64_t S, int32_t C_in, int32_t C_out, void* stream) {
  SB_REQUIRE(x16 && w && y && N > 0 && S > 0, "semabs_final_conv1x1_ncdhw: bad arguments");
  SB_REQUIRE(splits == 1 || splits == 2, "semabs_final_conv1x1_ncdhw: splits must be 1 or 2");
  SB_REQUIRE((C_in == 16 || C_in == 32 || C_in == 64) && C_out >= 1 && C_out <= 256,
             "semabs_final_conv1x1_ncdhw: unsupported channels %d -> %d", C_in, C_out);
  const long long want = (S + 511) / 512, cap = 8LL * num_sms();
  dim3 grid((unsigned)(want < cap ? want : cap), N);
  const size_t sm = size_t(C_out) * (C_in + 1) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const __half* x = (const __half*)x16;
  if (C_in == 16) final_conv1x1_ncdhw_kernel<16><<<grid, 256, sm, st>>>(x, splits, w, bias, y, S, C_out);
  else if (C_in == 32) final_conv1x1_ncdhw_kernel<32><<<grid, 256, sm, st>>>(x, splits, w, bias, y, S, C_out);
  else {
    static bool configured = false;
    if (!configured) {
      SB_CHECK_CUDA(cudaFuncSetAttribute(final_conv1x1_ncdhw_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 65 * 4));
      configured = true;
    }
    final_conv1x1_ncdhw_kernel<64><<<grid, 256, sm, st>>>(x, splits, w, bias, y, S, C_out);
  }
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_fold_groupnorm_halo(const float* w, const float* gamma, const float* beta, const double* stats,
                                          int32_t stats_stride, int32_t N, int64_t S, int32_t groups, void* w_img, float* bias_cls,
                                          void* stream) {
  SB_REQUIRE(w && gamma && beta && stats && w_img && bias_cls && N > 0 && S > 0, "semabs_fold_groupnorm_halo: null pointer");
  SB_REQUIRE(groups >= 1 && groups <= 8 && FOLD_C % groups == 0 && stats_stride >= 2 * groups,
             "semabs_fold_groupnorm_halo: unsupported grouping %d / stride %d", groups, stats_stride);
  const double inv_count = 1.0 / (double(S) * double(FOLD_C / groups));
  fold_groupnorm_halo_kernel<<<dim3(N, 16), 256, 0, (cudaStream_t)stream>>>(w, gamma, beta, stats, stats_stride, groups, inv_count, (__half*)w_img,
                                                                    bias_cls);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta,
                                      void* y16, int32_t N, int64_t S, int32_t C, int32_t C_real, int32_t groups,
                                      int32_t splits, int32_t planar, void* stream) {
  SB_REQUIRE(x && stats && gamma && beta && y16 && N > 0 && S > 0, "semabs_groupnorm_apply: null pointer");
  SB_REQUIRE(C % 4 == 0 && C <= 1024 && (EW_THREADS % (C / 4)) == 0 && groups >= 1 && groups <= 8 && C_real <= C,
             "semabs_groupnorm_apply: unsupported channel count %d", C);
  const int cpg = groups == 1 ? C : C_real / groups;
  const double inv_count = 1.0 / (double(S) * double(groups == 1 ? C_real : cpg));
  const int vpb = EW_THREADS / (C / 4);
  dim3 grid(ew_grid(S, vpb), N);
  SB_REQUIRE(!planar || C % 8 == 0, "semabs_groupnorm_apply: planar output needs C %% 8 == 0");
  gn_apply_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, stats, gamma, beta, (__half*)y16, S, C, groups, cpg,
                                                                 inv_count, splits, planar);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_maxpool3d_2(const float* x, float* y, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                                  int32_t groups, double* stats, void* stream) {
  SB_REQUIRE(x && y && N > 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "semabs_maxpool3d_2: bad arguments");
  SB_REQUIRE(C % 4 == 0 && (EW_THREADS % (C / 4)) == 0, "semabs_maxpool3d_2: unsupported channel count %d", C);
  SB_REQUIRE(!stats || (groups >= 1 && groups <= 8 && C % groups == 0 && (C / groups) % 2 == 0),
             "semabs_maxpool3d_2: unsupported GroupNorm grouping");
  const int vpb = EW_THREADS / (C / 4);
  const long long So = (long long)(D / 2) * (H / 2) * (W / 2);
  dim3 grid(ew_grid(So, vpb), N);
  maxpool2_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(x, y, D, H, W, C, groups, stats);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
