// Backward of the residual 3-D UNet stages (training step, reference utils.py:404-422 `loss.backward()` through
// unet3d.py's create_conv 'gcr' units, ExtResNetBlock, MaxPool3d, ConvTranspose3d and the final 1x1x1 conv).
//
// Gradient tensors that feed tensor-core MMAs are fp16 with a per-tensor power-of-two scale chosen on the device from
// the tensor's |max| (fp32 gradients of a mean BCE over ~10^7 points sit far below fp16's normal range):
//   * fp32 gradient tensors are TRUE scale unless they are the raw output of a data-gradient convolution, in which case
//     a device float `scale` travels with them (true = stored / scale); every kernel here takes that pointer (or null);
//   * `semabs_unet_bwd_pack` turns an fp32 gradient into the fp16 operands (ReLU mask applied) and writes their scale.
// Data gradients of the convolutions reuse the forward implicit-GEMM kernels with adjoint weight packs (conv3d.cu kinds
// 0/1/3).  Weight gradients: `conv3d_wgrad_kernel` below — a split-K reduction over the voxels with both operands read
// from zero-padded channels-last fp16 volumes [N, D+2, H+2, W+2, C], so that a tap is a flat row offset and the padding
// ring supplies the zeros of the convolution's border.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

constexpr int BW_THREADS = 256;

__device__ __forceinline__ float scale_from_amax(float amax) {
  // power of two that brings the largest magnitude into [2^12, 2^13)
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e = 12 - ilogbf(amax);
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return ldexpf(1.f, e);
}

__device__ __forceinline__ void block_amax_flush(float m, unsigned int* slot) {
  m = warp_max(m);
  __shared__ float sm[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    v = warp_max(v);
    if (lane == 0 && v > 0.f) atomicMax(slot, __float_as_uint(v));  // non-negative floats order like their bit patterns
  }
}

static int bw_grid(long long S, int vpb) {
  long long need = (S + vpb - 1) / vpb;
  long long cap = (long long)num_sms() * 8;
  return int(need < cap ? need : cap);
}

// ---- |max| of an fp32 tensor -----------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS) absmax_kernel(const float4* __restrict__ x, long long n4, unsigned int* slot) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = x[i];
    m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))), m);
  }
  block_amax_flush(m, slot);
}

// ---- fp32 gradient (or activation) -> fp16 MMA operands ---------------------------------------------------
// g [N,S,C] fp32; optional ReLU mask (mask > 0); f = scale_from_amax(*amax) (1 when amax == null).
//   pad : zero-padded channels-last [N, D+2, H+2, W+2, Cp] (interior written only); with parity != 0 the grid (D,H,W)
//         is split into its 8 parity sub-grids, stored as 8 consecutive padded volumes [8][N, D/2+2, H/2+2, W/2+2, Cp]
//   op  : unpadded operand of the forward conv kernels: channels-last [N,S,splits*C] (op_layout 1) or chunk-planar
//         [N][splits*C/8][S][8] (op_layout 2, conv3d_halo.cu); splits == 2 appends the fp16 remainder (hi | lo)
__global__ void __launch_bounds__(BW_THREADS)
bwd_pack_kernel(const float* __restrict__ g, const float* __restrict__ g_scale, const unsigned int* __restrict__ amax,
                const float* __restrict__ mask, int D, int H, int W, int C, __half* __restrict__ pad, int Cp, int parity,
                __half* __restrict__ op, int op_layout, int op_splits, float* __restrict__ scale_out, int N) {
  const int n = blockIdx.y;
  const long long S = (long long)D * H * W;
  const float f = amax ? scale_from_amax(__uint_as_float(*amax)) : 1.f;
  if (scale_out && blockIdx.x == 0 && n == 0 && threadIdx.x == 0) *scale_out = (g_scale ? *g_scale : 1.f) * f;
  const int qpc = C / 4, vpb = BW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float* gn = g + size_t(n) * S * C;
  const float* mn = mask ? mask + size_t(n) * S * C : nullptr;
  const int PD = (parity ? D / 2 : D) + 2, PH = (parity ? H / 2 : H) + 2, PW = (parity ? W / 2 : W) + 2;
  const size_t vol = size_t(N) * PD * PH * PW;  // voxels of one padded (parity) volume over all samples
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    float4 a = *reinterpret_cast<const float4*>(gn + size_t(v) * C + c);
    a.x *= f, a.y *= f, a.z *= f, a.w *= f;
    if (mn) {
      const float4 m = *reinterpret_cast<const float4*>(mn + size_t(v) * C + c);
      a.x = m.x > 0.f ? a.x : 0.f, a.y = m.y > 0.f ? a.y : 0.f, a.z = m.z > 0.f ? a.z : 0.f, a.w = m.w > 0.f ? a.w : 0.f;
    }
    const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
    const uint2 packed = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    if (pad) {
      const int x = int(v % W), y = int((v / W) % H), z = int(v / ((long long)W * H));
      size_t p;
      if (parity) {
        const int q = ((z & 1) << 2) | ((y & 1) << 1) | (x & 1);
        p = size_t(q) * vol + ((size_t(n) * PD + (z >> 1) + 1) * PH + (y >> 1) + 1) * PW + (x >> 1) + 1;
      } else {
        p = ((size_t(n) * PD + z + 1) * PH + y + 1) * PW + x + 1;
      }
      *reinterpret_cast<uint2*>(pad + p * Cp + c) = packed;
    }
    if (op) {
      // same layouts as gn_apply_kernel (unet_ops.cu): [v][splits*C] or chunk-planar, hi chunks then lo chunks
      __half* on = op + size_t(n) * S * op_splits * C;
      __half* dst = op_layout == 2 ? on + (size_t(c >> 3) * S + v) * 8 + (c & 7) : on + size_t(v) * op_splits * C + c;
      *reinterpret_cast<uint2*>(dst) = packed;
      if (op_splits == 2) {
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(a.x - f0.x, a.y - f0.y), l1 = __floats2half2_rn(a.z - f1.x, a.w - f1.y);
        __half* dlo = op_layout == 2 ? dst + size_t(C >> 3) * S * 8 : dst + C;
        *reinterpret_cast<uint2*>(dlo) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
    }
  }
}

// ---- GroupNorm apply into the padded operand layout (recomputes the forward's normalised activation) -----------
__global__ void __launch_bounds__(BW_THREADS)
gn_apply_padded_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, __half* __restrict__ pad, int D, int H, int W, int C, int groups,
                       int cpg, double inv_count, int Cp) {
  __shared__ float s_scale[1024], s_shift[1024];
  const int n = blockIdx.y;
  const long long S = (long long)D * H * W;
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
  const int qpc = C / 4, vpb = BW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float4 sc = make_float4(s_scale[c], s_scale[c + 1], s_scale[c + 2], s_scale[c + 3]);
  const float4 sh = make_float4(s_shift[c], s_shift[c + 1], s_shift[c + 2], s_shift[c + 3]);
  const float* xn = x + size_t(n) * S * C;
  const int PD = D + 2, PH = H + 2, PW = W + 2;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    const float4 a = *reinterpret_cast<const float4*>(xn + size_t(v) * C + c);
    const __half2 h0 = __floats2half2_rn(fmaf(a.x, sc.x, sh.x), fmaf(a.y, sc.y, sh.y));
    const __half2 h1 = __floats2half2_rn(fmaf(a.z, sc.z, sh.z), fmaf(a.w, sc.w, sh.w));
    const int xx = int(v % W), y = int((v / W) % H), z = int(v / ((long long)W * H));
    const size_t p = ((size_t(n) * PD + z + 1) * PH + y + 1) * PW + xx + 1;
    *reinterpret_cast<uint2*>(pad + p * Cp + c) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  }
}

// ---- GroupNorm backward, pass 1: per (sample, channel) sums of dy and dy * x ------------------------------------
__global__ void __launch_bounds__(BW_THREADS)
gn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, long long S, int C, double* __restrict__ sums) {
  const int n = blockIdx.y;
  const int qpc = C / 4, vpb = BW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float* dn = dy + size_t(n) * S * C;
  const float* xn = x ? x + size_t(n) * S * C : nullptr;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    const float4 d = *reinterpret_cast<const float4*>(dn + size_t(v) * C + c);
    s.x += d.x, s.y += d.y, s.z += d.z, s.w += d.w;
    if (xn) {
      const float4 a = *reinterpret_cast<const float4*>(xn + size_t(v) * C + c);
      q.x = fmaf(d.x, a.x, q.x), q.y = fmaf(d.y, a.y, q.y), q.z = fmaf(d.z, a.z, q.z), q.w = fmaf(d.w, a.w, q.w);
    }
  }
  // threads with the same channel quad sit qpc apart: reduce through shared memory, then one fp64 atomic per value
  __shared__ float4 sm_s[BW_THREADS], sm_q[BW_THREADS];
  sm_s[threadIdx.x] = s, sm_q[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < qpc) {
    float4 ts = sm_s[threadIdx.x], tq = sm_q[threadIdx.x];
    for (int j = 1; j < vpb; ++j) {
      const float4 a = sm_s[threadIdx.x + j * qpc], b = sm_q[threadIdx.x + j * qpc];
      ts.x += a.x, ts.y += a.y, ts.z += a.z, ts.w += a.w;
      tq.x += b.x, tq.y += b.y, tq.z += b.z, tq.w += b.w;
    }
    double* o = sums + (size_t(n) * C + c) * 2;
    atomicAdd(o + 0, double(ts.x)), atomicAdd(o + 1, double(tq.x));
    atomicAdd(o + 2, double(ts.y)), atomicAdd(o + 3, double(tq.y));
    atomicAdd(o + 4, double(ts.z)), atomicAdd(o + 5, double(tq.z));
    atomicAdd(o + 6, double(ts.w)), atomicAdd(o + 7, double(tq.w));
  }
}

// ---- GroupNorm backward, pass 2 ------------------------------------------------------------------------------
// dx = rstd * (gamma*dy - mean_g(gamma*dy) - xhat * mean_g(gamma*dy*xhat))   [torch.nn.GroupNorm backward]
//      (+ add * [add_mask > 0] / add_scale)  (+ previous content of dx when accumulate)
__global__ void __launch_bounds__(BW_THREADS)
gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ dy_scale, const float* __restrict__ x,
                    const double* __restrict__ stats, const float* __restrict__ gamma, const double* __restrict__ sums,
                    long long S, int C, int C_real, int groups, int cpg, double inv_count, const float* __restrict__ add,
                    const float* __restrict__ add_scale, const float* __restrict__ add_mask, float* __restrict__ dx,
                    int accumulate, unsigned int* amax) {
  __shared__ float s_a[1024], s_b[1024], s_c[1024];  // dx = s_a[c]*dy + s_b[c]*x + s_c[c]
  __shared__ double g_m[8], g_r[8], g_k1[8], g_k2[8];
  const int n = blockIdx.y;
  const float inv_s = dy_scale ? 1.f / *dy_scale : 1.f;
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    const double mean = stats[(size_t(n) * groups + g) * 2] * inv_count;
    double var = stats[(size_t(n) * groups + g) * 2 + 1] * inv_count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const double rstd = 1.0 / sqrt(var + 1e-5);
    const int c0 = g * cpg, c1 = groups == 1 ? C_real : c0 + cpg;
    double k1 = 0.0, k2 = 0.0;  // sum_c gamma * sum(dy), sum_c gamma * sum(dy * xhat)
    for (int c = c0; c < c1; ++c) {
      const double sd = sums[(size_t(n) * C + c) * 2], sdx = sums[(size_t(n) * C + c) * 2 + 1];
      k1 += double(gamma[c]) * sd;
      k2 += double(gamma[c]) * (sdx - mean * sd) * rstd;
    }
    g_m[g] = mean, g_r[g] = rstd, g_k1[g] = k1 * inv_count, g_k2[g] = k2 * inv_count;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (c < C_real) {
      const int g = min(c / cpg, groups - 1);
      const double r = g_r[g], m = g_m[g];
      // dx = r*gamma*dy - r*k1 - (x-m)*r * r*k2
      s_a[c] = float(r * double(gamma[c]) * inv_s);
      s_b[c] = float(-r * r * g_k2[g] * inv_s);
      s_c[c] = float((-r * g_k1[g] + m * r * r * g_k2[g]) * inv_s);
    } else {
      s_a[c] = s_b[c] = s_c[c] = 0.f;
    }
  }
  __syncthreads();
  const float inv_add = add_scale ? 1.f / *add_scale : 1.f;
  const int qpc = C / 4, vpb = BW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float4 ka = make_float4(s_a[c], s_a[c + 1], s_a[c + 2], s_a[c + 3]);
  const float4 kb = make_float4(s_b[c], s_b[c + 1], s_b[c + 2], s_b[c + 3]);
  const float4 kc = make_float4(s_c[c], s_c[c + 1], s_c[c + 2], s_c[c + 3]);
  const size_t base = size_t(n) * S * C;
  float m = 0.f;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    const size_t o = base + size_t(v) * C + c;
    const float4 d = *reinterpret_cast<const float4*>(dy + o);
    const float4 a = *reinterpret_cast<const float4*>(x + o);
    float4 r;
    r.x = fmaf(ka.x, d.x, fmaf(kb.x, a.x, kc.x));
    r.y = fmaf(ka.y, d.y, fmaf(kb.y, a.y, kc.y));
    r.z = fmaf(ka.z, d.z, fmaf(kb.z, a.z, kc.z));
    r.w = fmaf(ka.w, d.w, fmaf(kb.w, a.w, kc.w));
    if (add) {
      float4 e = *reinterpret_cast<const float4*>(add + o);
      if (add_mask) {
        const float4 k = *reinterpret_cast<const float4*>(add_mask + o);
        e.x = k.x > 0.f ? e.x : 0.f, e.y = k.y > 0.f ? e.y : 0.f, e.z = k.z > 0.f ? e.z : 0.f, e.w = k.w > 0.f ? e.w : 0.f;
      }
      r.x = fmaf(e.x, inv_add, r.x), r.y = fmaf(e.y, inv_add, r.y), r.z = fmaf(e.z, inv_add, r.z), r.w = fmaf(e.w, inv_add, r.w);
    }
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(dx + o);
      r.x += p.x, r.y += p.y, r.z += p.z, r.w += p.w;
    }
    *reinterpret_cast<float4*>(dx + o) = r;
    m = fmaxf(fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))), m);
  }
  if (amax) block_amax_flush(m, amax);
}

// dgamma[c] (+)= sum_n rstd * (sum(dy x) - mean * sum(dy)) / scale;  dbeta[c] (+)= sum_n sum(dy) / scale.
// stats == null: bias gradient only (dbeta <- column sums), used for the conv biases.
__global__ void gn_param_grads_kernel(const double* __restrict__ sums, const double* __restrict__ stats,
                                      const float* __restrict__ scale, int N, int C, int C_real, int groups, int cpg,
                                      double inv_count, float* dgamma, float* dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C_real) return;
  const double inv_s = scale ? 1.0 / double(*scale) : 1.0;
  double dg = 0.0, db = 0.0;
  for (int n = 0; n < N; ++n) {
    const double sd = sums[(size_t(n) * C + c) * 2], sdx = sums[(size_t(n) * C + c) * 2 + 1];
    db += sd;
    if (stats) {
      const int g = min(c / cpg, groups - 1);
      const double mean = stats[(size_t(n) * groups + g) * 2] * inv_count;
      double var = stats[(size_t(n) * groups + g) * 2 + 1] * inv_count - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      dg += (sdx - mean * sd) / sqrt(var + 1e-5);
    }
  }
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + float(dg * inv_s);
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + float(db * inv_s);
}

// ---- MaxPool3d(2) backward: the gradient goes to the first maximum in (z,y,x) scan order (torch's arg-max rule) ---
__global__ void __launch_bounds__(BW_THREADS)
maxpool2_bwd_kernel(const float* __restrict__ g, const float* __restrict__ g_scale, const float* __restrict__ x, int D,
                    int H, int W, int C, float* __restrict__ dx, int accumulate, unsigned int* amax) {
  const int n = blockIdx.y;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long So = (long long)Do * Ho * Wo;
  const float inv_s = g_scale ? 1.f / *g_scale : 1.f;
  const int qpc = C / 4, vpb = BW_THREADS / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc, c = 4 * cq;
  const float* xn = x + size_t(n) * D * H * W * C;
  float* dn = dx + size_t(n) * D * H * W * C;
  const float* gn = g + size_t(n) * So * C;
  float mx = 0.f;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < So; v += (long long)gridDim.x * vpb) {
    const int xo = int(v % Wo), yo = int((v / Wo) % Ho), zo = int(v / ((long long)Wo * Ho));
    float4 gv = *reinterpret_cast<const float4*>(gn + size_t(v) * C + c);
    gv.x *= inv_s, gv.y *= inv_s, gv.z *= inv_s, gv.w *= inv_s;
    float4 a[8];
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int ix = 0, iy = 0, iz = 0, iw = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t src = ((size_t(2 * zo + (k >> 2)) * H + (2 * yo + ((k >> 1) & 1))) * W + (2 * xo + (k & 1))) * C + c;
      a[k] = *reinterpret_cast<const float4*>(xn + src);
      if (a[k].x > m.x) m.x = a[k].x, ix = k;
      if (a[k].y > m.y) m.y = a[k].y, iy = k;
      if (a[k].z > m.z) m.z = a[k].z, iz = k;
      if (a[k].w > m.w) m.w = a[k].w, iw = k;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const size_t dst = ((size_t(2 * zo + (k >> 2)) * H + (2 * yo + ((k >> 1) & 1))) * W + (2 * xo + (k & 1))) * C + c;
      float4 r = make_float4(ix == k ? gv.x : 0.f, iy == k ? gv.y : 0.f, iz == k ? gv.z : 0.f, iw == k ? gv.w : 0.f);
      if (accumulate) {
        const float4 p = *reinterpret_cast<const float4*>(dn + dst);
        r.x += p.x, r.y += p.y, r.z += p.z, r.w += p.w;
      }
      *reinterpret_cast<float4*>(dn + dst) = r;
      mx = fmaxf(fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))), mx);
    }
  }
  if (amax) block_amax_flush(mx, amax);
}

// =================================================================================================================
// Weight gradient: out[slot][a][b] = sum_p A[p][a] * B[p + off(slot)][b]  over the flat padded voxel index p
// =================================================================================================================
constexpr int WG_WARPS = 9;            // one (dz,dy) row segment of the 3x3x3 stencil per warp, <= 3 taps (dx) each
constexpr int WG_THREADS = WG_WARPS * 32;
constexpr int WG_KB = 64;              // voxels per pipeline stage
constexpr int WG_STAGES = 3;
constexpr int WG_SEGROWS = WG_KB + 2;  // a segment holds the rows [p0 + seg_off, p0 + seg_off + KB + 2)

struct WgradParams {
  const __half* A;   // [*, lda] padded channels-last, zero outside the interior (and in the guard rows)
  const __half* B;   // [*, ldb]
  int lda, ldb;      // channel strides (elements)
  int Ca, Cb;        // padded channel counts covered by the CTA grid (multiples of MA / NB)
  long long nvox;    // rows of A to reduce over (multiple of WG_KB, guard rows included)
  int nchunks, nsplit;
  int nseg;
  long long seg_off[WG_WARPS];
  int seg_ntaps[WG_WARPS];
  int seg_sh[WG_WARPS][3];
  int seg_slot[WG_WARPS][3];
  int nslots;
  float* partial;    // [nsplit][nslots][Ca][Cb]
};

__device__ __forceinline__ void mma_16816_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void cp_async_16(uint32_t saddr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gptr) : "memory");
}

// rows of CH halves (CH = 16 or 32); the 16-byte chunks of a row are XOR-swizzled so that 8 consecutive rows of one
// logical chunk land in 8 different bank groups (conflict-free transposed ldmatrix)
template <int CH>
__device__ __forceinline__ uint32_t row_chunk_off(int row, int chunk) {
  constexpr int CPR = CH / 8;                     // chunks per row: 2 or 4
  constexpr int SH = CPR == 4 ? 1 : 2;
  return uint32_t(row) * (CH * 2) + uint32_t((chunk ^ ((row >> SH) & (CPR - 1))) * 16);
}

// 16 x 16 channel tiles need 96 registers and 63 KB of shared memory: two CTAs per SM (ncu: one CTA kept the mma.sync pipe
// 25 % busy at 14 % warp occupancy — issue / latency bound, not L2 bound)
template <int MA, int NB>
__global__ void __launch_bounds__(WG_THREADS, (MA * NB <= 256) ? 2 : 1) conv3d_wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int MT = MA / 16, NT = NB / 8;
  constexpr int A_BYTES = WG_KB * MA * 2;
  constexpr int SEG_BYTES = WG_SEGROWS * NB * 2;
  constexpr int STAGE_BYTES = A_BYTES + WG_WARPS * SEG_BYTES;
  extern __shared__ __align__(128) uint8_t wg_smem[];
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(wg_smem));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_blocks = p.Ca / MA, b_blocks = p.Cb / NB;
  const int pair = blockIdx.x % (a_blocks * b_blocks);
  const int split = blockIdx.x / (a_blocks * b_blocks);
  const int ablk = pair / b_blocks, bblk = pair % b_blocks;
  const __half* Ag = p.A + ablk * MA;
  const __half* Bg = p.B + bblk * NB;

  float acc[3][MT][NT][4];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][mt][nt][e] = 0.f;

  const int my_ntaps = warp < p.nseg ? p.seg_ntaps[warp] : 0;
  const int sh0 = p.seg_sh[warp][0], sh1 = p.seg_sh[warp][1], sh2 = p.seg_sh[warp][2];

  auto issue = [&](int chunk, int stage) {
    const long long p0 = (long long)chunk * WG_KB;
    const uint32_t sA = sbase + stage * STAGE_BYTES;
    constexpr int ACH = MA / 8, BCH = NB / 8;
    for (int i = threadIdx.x; i < WG_KB * ACH; i += WG_THREADS) {
      const int row = i / ACH, ch = i % ACH;
      cp_async_16(sA + row_chunk_off<MA>(row, ch), Ag + (p0 + row) * p.lda + ch * 8);
    }
    const int per_seg = WG_SEGROWS * BCH;
    for (int i = threadIdx.x; i < p.nseg * per_seg; i += WG_THREADS) {
      const int s = i / per_seg, r = i % per_seg;
      const int row = r / BCH, ch = r % BCH;
      cp_async_16(sA + A_BYTES + s * SEG_BYTES + row_chunk_off<NB>(row, ch),
                  Bg + (p0 + p.seg_off[s] + row) * (long long)p.ldb + ch * 8);
    }
  };

  // chunks of this split: split, split + nsplit, ...
  const int my_chunks = p.nchunks > split ? (p.nchunks - split + p.nsplit - 1) / p.nsplit : 0;
#pragma unroll
  for (int s = 0; s < WG_STAGES - 1; ++s) {
    if (s < my_chunks) issue(split + s * p.nsplit, s);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  for (int it = 0; it < my_chunks; ++it) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(WG_STAGES - 2) : "memory");
    __syncthreads();
    {
      const int nx = it + WG_STAGES - 1;
      if (nx < my_chunks) issue(split + nx * p.nsplit, nx % WG_STAGES);
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    if (my_ntaps > 0) {
      const uint32_t sA = sbase + (it % WG_STAGES) * STAGE_BYTES;
      const uint32_t sB = sA + A_BYTES + warp * SEG_BYTES;
#pragma unroll
      for (int ks = 0; ks < WG_KB / 16; ++ks) {
        uint32_t fa[MT][4];
        {
          const int row = ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) ldsm_x4_trans(fa[mt], sA + row_chunk_off<MA>(row, mt * 2 + ((lane >> 3) & 1)));
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < my_ntaps) {
            const int sh = j == 0 ? sh0 : (j == 1 ? sh1 : sh2);
            const int row = ks * 16 + sh + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
              uint32_t fb[4];
              ldsm_x4_trans(fb, sB + row_chunk_off<NB>(row, np * 2 + (lane >> 4)));
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                mma_16816_f16(acc[j][mt][2 * np], fa[mt], fb[0], fb[1]);
                mma_16816_f16(acc[j][mt][2 * np + 1], fa[mt], fb[2], fb[3]);
              }
            }
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  // partial[split][slot][a][b]
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j < my_ntaps) {
      const int slot = p.seg_slot[warp][j];
      float* out = p.partial + ((size_t(split) * p.nslots + slot) * p.Ca + ablk * MA) * p.Cb + bblk * NB;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int m = mt * 16 + g, n = nt * 8 + 2 * t;
          *reinterpret_cast<float2*>(out + size_t(m) * p.Cb + n) = make_float2(acc[j][mt][nt][0], acc[j][mt][nt][1]);
          *reinterpret_cast<float2*>(out + size_t(m + 8) * p.Cb + n) = make_float2(acc[j][mt][nt][2], acc[j][mt][nt][3]);
        }
    }
  }
}

// grad[(a * Cb_real + b) * KT + slot_k[slot]] (+)= sum_split partial[split][slot][a][b] / scale
// grid (ceil(Ca_real*Cb_real / 256), nslots): consecutive threads read consecutive b (coalesced) and add the splits in
// a fixed order (deterministic)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int nslots, int Ca, int Cb, int Ca_real, int Cb_real,
                    int KT, const int* __restrict__ slot_k, const float* __restrict__ scale, float* __restrict__ grad,
                    int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Ca_real * Cb_real) return;
  const int a = int(i / Cb_real), b = int(i % Cb_real);
  const int s = blockIdx.y;
  const float inv_s = scale ? 1.f / *scale : 1.f;
  const float* src = partial + (size_t(s) * Ca + a) * Cb + b;
  const size_t stride = size_t(nslots) * Ca * Cb;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
  int k = 0;
  for (; k + 4 <= nsplit; k += 4) {
    v0 += src[size_t(k) * stride], v1 += src[size_t(k + 1) * stride];
    v2 += src[size_t(k + 2) * stride], v3 += src[size_t(k + 3) * stride];
  }
  for (; k < nsplit; ++k) v0 += src[size_t(k) * stride];
  float* o = grad + (size_t(a) * Cb_real + b) * KT + slot_k[s];
  *o = (accumulate ? *o : 0.f) + ((v0 + v1) + (v2 + v3)) * inv_s;
}

template <int MA, int NB>
static int launch_wgrad(const WgradParams& p, int grid, cudaStream_t st) {
  constexpr int STAGE_BYTES = WG_KB * MA * 2 + WG_WARPS * WG_SEGROWS * NB * 2;
  constexpr int SMEM = WG_STAGES * STAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(conv3d_wgrad_kernel<MA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  conv3d_wgrad_kernel<MA, NB><<<grid, WG_THREADS, SMEM, st>>>(p);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_absmax_f32(const float* x, int64_t n, void* amax_slot, void* stream) {
  SB_REQUIRE(x && amax_slot && n > 0 && n % 4 == 0, "semabs_absmax_f32: bad arguments");
  const long long n4 = n / 4;
  long long blocks = (n4 + BW_THREADS - 1) / BW_THREADS;
  const long long cap = (long long)num_sms() * 8;
  absmax_kernel<<<int(blocks < cap ? blocks : cap), BW_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), n4, (unsigned int*)amax_slot);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static bool quad_ok(int C) { return C % 4 == 0 && C >= 4 && C <= 1024 && (BW_THREADS % (C / 4)) == 0; }

extern "C" int semabs_unet_bwd_pack(const float* g, const float* g_scale, const void* amax, const float* mask, int32_t N,
                                    int32_t D, int32_t H, int32_t W, int32_t C, void* pad16, int32_t Cp, int32_t parity,
                                    void* op16, int32_t op_layout, int32_t op_splits, float* scale_out, void* stream) {
  SB_REQUIRE(g && (pad16 || op16) && N > 0 && D > 0 && H > 0 && W > 0, "semabs_unet_bwd_pack: bad arguments");
  SB_REQUIRE(quad_ok(C) && (!pad16 || Cp >= C), "semabs_unet_bwd_pack: unsupported channel count %d", C);
  SB_REQUIRE(!parity || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "semabs_unet_bwd_pack: parity split needs even extents");
  SB_REQUIRE(!op16 || ((op_layout == 1 || (op_layout == 2 && C % 8 == 0)) && (op_splits == 1 || op_splits == 2)),
             "semabs_unet_bwd_pack: bad operand layout");
  const int vpb = BW_THREADS / (C / 4);
  dim3 grid(bw_grid((long long)D * H * W, vpb), N);
  bwd_pack_kernel<<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(g, g_scale, (const unsigned int*)amax, mask, D, H, W, C,
                                                                 (__half*)pad16, Cp, parity, (__half*)op16, op_layout,
                                                                 op_splits, scale_out, N);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_groupnorm_apply_padded(const float* x, const double* stats, const float* gamma, const float* beta,
                                             void* pad16, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                                             int32_t C_real, int32_t groups, int32_t Cp, void* stream) {
  SB_REQUIRE(x && stats && gamma && beta && pad16 && N > 0, "semabs_groupnorm_apply_padded: null pointer");
  SB_REQUIRE(quad_ok(C) && groups >= 1 && groups <= 8 && C_real <= C && Cp >= C,
             "semabs_groupnorm_apply_padded: unsupported channel count %d", C);
  const long long S = (long long)D * H * W;
  const int cpg = groups == 1 ? C : C_real / groups;
  const double inv_count = 1.0 / (double(S) * double(groups == 1 ? C_real : cpg));
  const int vpb = BW_THREADS / (C / 4);
  dim3 grid(bw_grid(S, vpb), N);
  gn_apply_padded_kernel<<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(x, stats, gamma, beta, (__half*)pad16, D, H, W, C,
                                                                        groups, cpg, inv_count, Cp);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_groupnorm_bwd_reduce(const float* dy, const float* x, int32_t N, int64_t S, int32_t C, double* sums,
                                           void* stream) {
  SB_REQUIRE(dy && sums && N > 0 && S > 0 && quad_ok(C), "semabs_groupnorm_bwd_reduce: bad arguments (C=%d)", C);
  const int vpb = BW_THREADS / (C / 4);
  long long need = (S + vpb - 1) / vpb;
  const long long cap = (long long)num_sms() * 4;
  dim3 grid(int(need < cap ? need : cap), N);
  gn_bwd_reduce_kernel<<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(dy, x, S, C, sums);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_groupnorm_bwd_apply(const float* dy, const float* dy_scale, const float* x, const double* stats,
                                          const float* gamma, const double* sums, int32_t N, int64_t S, int32_t C,
                                          int32_t C_real, int32_t groups, const float* add, const float* add_scale,
                                          const float* add_mask, float* dx, int32_t accumulate, void* amax_slot,
                                          void* stream) {
  SB_REQUIRE(dy && x && stats && gamma && sums && dx && N > 0 && S > 0, "semabs_groupnorm_bwd_apply: null pointer");
  SB_REQUIRE(quad_ok(C) && groups >= 1 && groups <= 8 && C_real <= C, "semabs_groupnorm_bwd_apply: unsupported C=%d", C);
  const int cpg = groups == 1 ? C : C_real / groups;
  const double inv_count = 1.0 / (double(S) * double(groups == 1 ? C_real : cpg));
  const int vpb = BW_THREADS / (C / 4);
  dim3 grid(bw_grid(S, vpb), N);
  gn_bwd_apply_kernel<<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(dy, dy_scale, x, stats, gamma, sums, S, C, C_real,
                                                                     groups, cpg, inv_count, add, add_scale, add_mask, dx,
                                                                     accumulate, (unsigned int*)amax_slot);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_groupnorm_param_grads(const double* sums, const double* stats, const float* scale, int32_t N,
                                            int64_t S, int32_t C, int32_t C_real, int32_t groups, float* dgamma,
                                            float* dbeta, int32_t accumulate, void* stream) {
  SB_REQUIRE(sums && (dgamma || dbeta) && N > 0 && C_real > 0 && C_real <= C, "semabs_groupnorm_param_grads: bad arguments");
  SB_REQUIRE(!dgamma || stats, "semabs_groupnorm_param_grads: dgamma needs the forward statistics");
  const int g = groups < 1 ? 1 : groups;
  const int cpg = g == 1 ? C : C_real / g;
  const double inv_count = 1.0 / (double(S) * double(g == 1 ? C_real : cpg));
  gn_param_grads_kernel<<<(C_real + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, stats, scale, N, C, C_real, g, cpg,
                                                                               inv_count, dgamma, dbeta, accumulate);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_maxpool3d_2_bwd(const float* g, const float* g_scale, const float* x, int32_t N, int32_t D, int32_t H,
                                      int32_t W, int32_t C, float* dx, int32_t accumulate, void* amax_slot, void* stream) {
  SB_REQUIRE(g && x && dx && N > 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "semabs_maxpool3d_2_bwd: bad arguments");
  SB_REQUIRE(quad_ok(C), "semabs_maxpool3d_2_bwd: unsupported channel count %d", C);
  const int vpb = BW_THREADS / (C / 4);
  const long long So = (long long)(D / 2) * (H / 2) * (W / 2);
  dim3 grid(bw_grid(So, vpb), N);
  maxpool2_bwd_kernel<<<grid, BW_THREADS, 0, (cudaStream_t)stream>>>(g, g_scale, x, D, H, W, C, dx, accumulate,
                                                                     (unsigned int*)amax_slot);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_conv3d_wgrad(const void* A16, int32_t lda, int32_t Ca, int32_t Ca_real, const void* B16, int32_t ldb,
                                   int32_t Cb, int32_t Cb_real, int64_t nvox, int32_t nseg, const int64_t* seg_off,
                                   const int32_t* seg_ntaps, const int32_t* seg_sh, const int32_t* seg_slot,
                                   int32_t nslots, const int32_t* slot_k_dev, int32_t KT, void* workspace,
                                   int64_t workspace_bytes, const float* scale, float* grad, int32_t accumulate,
                                   void* stream) {
  SB_REQUIRE(A16 && B16 && workspace && grad && seg_off && seg_ntaps && seg_sh && seg_slot && slot_k_dev,
             "semabs_conv3d_wgrad: null pointer");
  SB_REQUIRE(nseg >= 1 && nseg <= WG_WARPS && nslots >= 1 && nslots <= 27 && nvox > 0 && nvox % WG_KB == 0,
             "semabs_conv3d_wgrad: bad segment table / voxel count (nvox must be a multiple of %d)", WG_KB);
  SB_REQUIRE(Ca % 16 == 0 && Cb % 16 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= Ca && ldb >= Cb && Ca_real <= Ca &&
                 Cb_real <= Cb,
             "semabs_conv3d_wgrad: channel counts must be multiples of 16 (Ca=%d Cb=%d)", Ca, Cb);
  WgradParams p{};
  p.A = (const __half*)A16, p.B = (const __half*)B16, p.lda = lda, p.ldb = ldb, p.Ca = Ca, p.Cb = Cb;
  p.nvox = nvox, p.nchunks = int(nvox / WG_KB), p.nseg = nseg, p.nslots = nslots;
  for (int s = 0; s < WG_WARPS; ++s) {
    p.seg_off[s] = s < nseg ? seg_off[s] : 0;
    p.seg_ntaps[s] = s < nseg ? seg_ntaps[s] : 0;
    SB_REQUIRE(p.seg_ntaps[s] >= 0 && p.seg_ntaps[s] <= 3, "semabs_conv3d_wgrad: a segment holds at most 3 taps");
    for (int j = 0; j < 3; ++j) {
      p.seg_sh[s][j] = s < nseg ? seg_sh[s * 3 + j] : 0;
      p.seg_slot[s][j] = s < nseg ? seg_slot[s * 3 + j] : 0;
      SB_REQUIRE(p.seg_sh[s][j] >= 0 && p.seg_sh[s][j] <= 2 && p.seg_slot[s][j] >= 0 && p.seg_slot[s][j] < nslots,
                 "semabs_conv3d_wgrad: bad tap shift / slot");
    }
  }
  const int MA = Ca % 32 == 0 ? 32 : 16, NB = Cb % 32 == 0 ? 32 : 16;
  const int pairs = (Ca / MA) * (Cb / NB);
  const size_t per_split = size_t(nslots) * Ca * Cb * sizeof(float);
  const int cta_slots = num_sms() * ((MA * NB <= 256) ? 2 : 1);
  int nsplit = pairs >= cta_slots ? 1 : cta_slots / pairs;
  if (nsplit > p.nchunks) nsplit = p.nchunks;
  if (size_t(nsplit) * per_split > size_t(workspace_bytes)) nsplit = int(size_t(workspace_bytes) / per_split);
  SB_REQUIRE(nsplit >= 1, "semabs_conv3d_wgrad: workspace of %lld bytes is too small (%zu per split)",
             (long long)workspace_bytes, per_split);
  p.nsplit = nsplit;
  p.partial = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = pairs * nsplit;
  int rc;
  if (MA == 32 && NB == 32) rc = launch_wgrad<32, 32>(p, grid, st);
  else if (MA == 32) rc = launch_wgrad<32, 16>(p, grid, st);
  else if (NB == 32) rc = launch_wgrad<16, 32>(p, grid, st);
  else rc = launch_wgrad<16, 16>(p, grid, st);
  if (rc) return rc;
  const long long outs = (long long)Ca_real * Cb_real;
  wgrad_reduce_kernel<<<dim3(unsigned((outs + 255) / 256), nslots), 256, 0, st>>>(p.partial, nsplit, nslots, Ca, Cb, Ca_real,
                                                                                  Cb_real, KT, slot_k_dev, scale, grad,
                                                                                  accumulate);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
