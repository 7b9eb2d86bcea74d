// Multi-head attention of the CLIP transformers: forward that materialises the softmax probabilities (the
// reference keeps them through a hook, CLIP/clip/auxiliary.py:307-337) and the hand-written backward that yields,
// for P stacked label cotangents, the per-head relevance term  sum_i r_i * relu(dA ⊙ A)[i, :]  of
// ClipGradcam.interpret (clip_gradcam.py:90-126) together with dQ/dK/dV for the next block down.
//
// Forward: fp32 SIMT (4 % of the FLOPs; exact softmax for parity).  Backward: mma.sync m16n8k16 fp16 tiles with
// fp32 accumulation, two passes (row owner -> dQ, column owner -> dK, dV, relevance) so that nothing needs atomics
// and results are run-to-run deterministic.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

constexpr int HD = 64;  // head dim of every CLIP transformer (width / heads == 64)

// =========================================================================================================
// forward
// =========================================================================================================
constexpr int KV_STRIDE = 68;  // floats; float4-aligned rows, conflict-free for quarter-warp float4 reads

template <int NCHUNK>
__global__ void __launch_bounds__(256, 1)
attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ probs, float* __restrict__ o32,
                __half* __restrict__ o16, int T, int H, int d, int causal, int splits) {
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = Ks + size_t(T) * KV_STRIDE;
  float* Ps = Vs + size_t(T) * KV_STRIDE;  // [8 warps][Tpad]
  const int Tpad = NCHUNK * 32;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + size_t(b) * T * 3 * d + h * HD;

  for (int idx = threadIdx.x; idx < T * (HD / 4); idx += blockDim.x) {
    const int j = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
    const float4 kk = *reinterpret_cast<const float4*>(base + size_t(j) * 3 * d + d + c);
    const float4 vv = *reinterpret_cast<const float4*>(base + size_t(j) * 3 * d + 2 * d + c);
    *reinterpret_cast<float4*>(Ks + j * KV_STRIDE + c) = kk;
    *reinterpret_cast<float4*>(Vs + j * KV_STRIDE + c) = vv;
  }
  __syncthreads();

  float* prow = Ps + warp * Tpad;
  for (int i = warp; i < T; i += 8) {
    float q[HD];
    const float4* qp = reinterpret_cast<const float4*>(base + size_t(i) * 3 * d);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 t4 = qp[c];
      q[4 * c] = t4.x, q[4 * c + 1] = t4.y, q[4 * c + 2] = t4.z, q[4 * c + 3] = t4.w;
    }
    float s[NCHUNK];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      const int j = c * 32 + lane;
      float acc = -INFINITY;
      if (j < T && !(causal && j > i)) {
        acc = 0.f;
        const float4* kp = reinterpret_cast<const float4*>(Ks + j * KV_STRIDE);
#pragma unroll
        for (int e = 0; e < HD / 4; ++e) {
          const float4 k4 = kp[e];
          acc = fmaf(q[4 * e], k4.x, acc);
          acc = fmaf(q[4 * e + 1], k4.y, acc);
          acc = fmaf(q[4 * e + 2], k4.z, acc);
          acc = fmaf(q[4 * e + 3], k4.w, acc);
        }
      }
      s[c] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      s[c] = (s[c] == -INFINITY) ? 0.f : expf(s[c] - mx);
      sum += s[c];
    }
    sum = warp_sum(sum);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      const int j = c * 32 + lane;
      const float a = s[c] / sum;
      prow[j] = a;
      if (probs && j < T) probs[(size_t(blockIdx.x) * T + i) * T + j] = a;
    }
    __syncwarp();
    // O[i, :] = sum_j a_j V[j, :]; lane owns columns 2*lane, 2*lane+1
    float o0 = 0.f, o1 = 0.f;
    const int jend = causal ? (i + 1) : T;
    int j = 0;
    for (; j + 4 <= jend; j += 4) {
      const float4 a4 = *reinterpret_cast<const float4*>(prow + j);
      const float2 v0 = *reinterpret_cast<const float2*>(Vs + (j + 0) * KV_STRIDE + 2 * lane);
      const float2 v1 = *reinterpret_cast<const float2*>(Vs + (j + 1) * KV_STRIDE + 2 * lane);
      const float2 v2 = *reinterpret_cast<const float2*>(Vs + (j + 2) * KV_STRIDE + 2 * lane);
      const float2 v3 = *reinterpret_cast<const float2*>(Vs + (j + 3) * KV_STRIDE + 2 * lane);
      o0 = fmaf(a4.x, v0.x, o0), o1 = fmaf(a4.x, v0.y, o1);
      o0 = fmaf(a4.y, v1.x, o0), o1 = fmaf(a4.y, v1.y, o1);
      o0 = fmaf(a4.z, v2.x, o0), o1 = fmaf(a4.z, v2.y, o1);
      o0 = fmaf(a4.w, v3.x, o0), o1 = fmaf(a4.w, v3.y, o1);
    }
    for (; j < jend; ++j) {
      const float a = prow[j];
      const float2 v = *reinterpret_cast<const float2*>(Vs + j * KV_STRIDE + 2 * lane);
      o0 = fmaf(a, v.x, o0), o1 = fmaf(a, v.y, o1);
    }
    const size_t row = size_t(b) * T + i;
    const int col = h * HD + 2 * lane;
    if (o32) *reinterpret_cast<float2*>(o32 + row * d + col) = make_float2(o0, o1);
    if (o16) {
      const __half2 hi = __floats2half2_rn(o0, o1);
      *reinterpret_cast<__half2*>(o16 + row * size_t(splits) * d + col) = hi;
      if (splits == 2) {
        const float2 f = __half22float2(hi);
        *reinterpret_cast<__half2*>(o16 + row * size_t(splits) * d + d + col) = __floats2half2_rn(o0 - f.x, o1 - f.y);
      }
    }
    __syncwarp();
  }
}

// =========================================================================================================
// backward
// =========================================================================================================
// delta[pbh, i] = sum_d dO[pb, i, h, d] * O[b, i, h, d]  ( == sum_j dA_ij A_ij, the softmax-backward row term )
__global__ void attn_bwd_delta_kernel(const __half* __restrict__ dO16, int ld_do, const float* __restrict__ o32,
                                      float* __restrict__ delta, int PB, int B, int T, int H, int d) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= PB * T * H) return;
  const int h = gw % H, i = (gw / H) % T, pb = gw / (H * T);
  const int b = pb % B;
  const __half2 g = *reinterpret_cast<const __half2*>(dO16 + (size_t(pb) * T + i) * ld_do + h * HD + 2 * lane);
  const float2 o = *reinterpret_cast<const float2*>(o32 + (size_t(b) * T + i) * d + h * HD + 2 * lane);
  const float2 gf = __half22float2(g);
  float s = gf.x * o.x + gf.y * o.y;
  s = warp_sum(s);
  if (lane == 0) delta[(size_t(pb) * H + h) * T + i] = s;
}

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int TS = 72;   // half stride of 64-wide fp16 smem tiles (conflict-free 32-bit fragment loads)
constexpr int AS = 68;   // float stride of the fp32 probability tile

struct AttnBwdArgs {
  const float* qkv;      // [B*T, 3d] fp32 (q already scaled)
  const float* probs;    // [B*H, T, T] fp32
  const __half* dO16;    // [P*B*T, ld_do] fp16 (hi part used)
  int ld_do;
  const float* delta;    // [P*B*H, T]
  const float* r;        // [P*B, T] current rollout row vector
  float* wpart;          // [P*B*H, T] out: per-head relevance contribution
  __half* dqkv16;        // [P*B*T, splits*3d] out
  int P, B, T, H, d, splits;
  float scale;           // hd^-0.5 (chain rule through q *= scale)
  int positive_only;
  int need_dqkv;         // 0 for the lowest rollout block (only relevance is needed)
};

__device__ __forceinline__ void store_h2_split(__half* base, size_t row_off, int col, int width, int splits, float x,
                                               float y) {
  const __half2 hi = __floats2half2_rn(x, y);
  *reinterpret_cast<__half2*>(base + row_off + col) = hi;
  if (splits == 2) {
    const float2 f = __half22float2(hi);
    *reinterpret_cast<__half2*>(base + row_off + width + col) = __floats2half2_rn(x - f.x, y - f.y);
  }
}

// ---- pass 1: row owner. CTA = 64 query rows (4 warps x 16), loops over key blocks; dQ = (A ⊙ (dA - delta)) K
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnBwdArgs a) {
  __shared__ __align__(16) __half Vs[64 * TS];  // [key][d]
  __shared__ __align__(16) __half Kt[64 * TS];  // [d][key]
  const int p = blockIdx.x % a.P, iblk = blockIdx.x / a.P;
  const int b = blockIdx.y / a.H, h = blockIdx.y % a.H;
  const int pb = p * a.B + b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int T = a.T, d = a.d;
  const int i0 = iblk * 64 + warp * 16;
  const int ia = i0 + g, ib = i0 + g + 8;

  // dO fragments (A operand, 16 rows x 64)
  uint32_t fo[4][4];
  {
    const __half* ra = a.dO16 + (size_t(pb) * T + ia) * a.ld_do + h * HD;
    const __half* rb = a.dO16 + (size_t(pb) * T + ib) * a.ld_do + h * HD;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      fo[kt][0] = ia < T ? *reinterpret_cast<const uint32_t*>(ra + kt * 16 + 2 * t) : 0u;
      fo[kt][1] = ib < T ? *reinterpret_cast<const uint32_t*>(rb + kt * 16 + 2 * t) : 0u;
      fo[kt][2] = ia < T ? *reinterpret_cast<const uint32_t*>(ra + kt * 16 + 2 * t + 8) : 0u;
      fo[kt][3] = ib < T ? *reinterpret_cast<const uint32_t*>(rb + kt * 16 + 2 * t + 8) : 0u;
    }
  }
  const float* dl = a.delta + (size_t(pb) * a.H + h) * T;
  const float da = ia < T ? dl[ia] : 0.f, db = ib < T ? dl[ib] : 0.f;
  const float* Arow_a = a.probs + (size_t(blockIdx.y) * T + (ia < T ? ia : 0)) * T;
  const float* Arow_b = a.probs + (size_t(blockIdx.y) * T + (ib < T ? ib : 0)) * T;

  float dq[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;

  const float* kvbase = a.qkv + size_t(b) * T * 3 * d + h * HD;
  for (int j0 = 0; j0 < T; j0 += 64) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 32; idx += blockDim.x) {
      const int j = idx >> 5, c = (idx & 31) * 2;
      float2 kk = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
      if (j0 + j < T) {
        kk = *reinterpret_cast<const float2*>(kvbase + size_t(j0 + j) * 3 * d + d + c);
        vv = *reinterpret_cast<const float2*>(kvbase + size_t(j0 + j) * 3 * d + 2 * d + c);
      }
      *reinterpret_cast<__half2*>(Vs + j * TS + c) = __floats2half2_rn(vv.x, vv.y);
      Kt[c * TS + j] = __float2half_rn(kk.x);
      Kt[(c + 1) * TS + j] = __float2half_rn(kk.y);
    }
    __syncthreads();

    float gacc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      gacc[n][0] = gacc[n][1] = gacc[n][2] = gacc[n][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(Vs + (n * 8 + g) * TS + kt * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(Vs + (n * 8 + g) * TS + kt * 16 + 2 * t + 8);
        mma_16816(gacc[n], fo[kt], b0, b1);
      }
    }
    // dS = A ⊙ (G - delta_i)
    uint32_t fs[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int j = j0 + n * 8 + 2 * t;
      const float a00 = (ia < T && j < T) ? Arow_a[j] : 0.f;
      const float a01 = (ia < T && j + 1 < T) ? Arow_a[j + 1] : 0.f;
      const float a10 = (ib < T && j < T) ? Arow_b[j] : 0.f;
      const float a11 = (ib < T && j + 1 < T) ? Arow_b[j + 1] : 0.f;
      const float s00 = a00 * (gacc[n][0] - da), s01 = a01 * (gacc[n][1] - da);
      const float s10 = a10 * (gacc[n][2] - db), s11 = a11 * (gacc[n][3] - db);
      fs[n >> 1][(n & 1) * 2 + 0] = pack_h2(s00, s01);
      fs[n >> 1][(n & 1) * 2 + 1] = pack_h2(s10, s11);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(Kt + (n * 8 + g) * TS + kt * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(Kt + (n * 8 + g) * TS + kt * 16 + 2 * t + 8);
        mma_16816(dq[n], fs[kt], b0, b1);
      }
    }
  }
  const size_t ld = size_t(a.splits) * 3 * d;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = h * HD + n * 8 + 2 * t;
    if (ia < T) store_h2_split(a.dqkv16, (size_t(pb) * T + ia) * ld, col, 3 * d, a.splits, dq[n][0] * a.scale, dq[n][1] * a.scale);
    if (ib < T) store_h2_split(a.dqkv16, (size_t(pb) * T + ib) * ld, col, 3 * d, a.splits, dq[n][2] * a.scale, dq[n][3] * a.scale);
  }
}

// ---- pass 2: column owner. CTA = 64 key rows, loops over query blocks; relevance, dK = dSᵀ Q, dV = Aᵀ dO
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(AttnBwdArgs a) {
  __shared__ __align__(16) __half dOs[64 * TS];  // [query][d]
  __shared__ __align__(16) __half dOt[64 * TS];  // [d][query]
  __shared__ __align__(16) __half Qt[64 * TS];   // [d][query]
  __shared__ __align__(16) float As[64 * AS];    // [query][key]
  __shared__ float Dl[64], Rw[64];
  const int p = blockIdx.x % a.P, jblk = blockIdx.x / a.P;
  const int b = blockIdx.y / a.H, h = blockIdx.y % a.H;
  const int pb = p * a.B + b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int T = a.T, d = a.d;
  const int jbase = jblk * 64;
  const int ja = jbase + warp * 16 + g, jb = ja + 8;
  const float* kvbase = a.qkv + size_t(b) * T * 3 * d + h * HD;

  // V_j fragments (A operand of Gᵀ = V dOᵀ)
  uint32_t fv[4][4];
  {
    const float* ra = kvbase + size_t(ja < T ? ja : 0) * 3 * d + 2 * d;
    const float* rb = kvbase + size_t(jb < T ? jb : 0) * 3 * d + 2 * d;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const float2 x0 = *reinterpret_cast<const float2*>(ra + kt * 16 + 2 * t);
      const float2 x1 = *reinterpret_cast<const float2*>(rb + kt * 16 + 2 * t);
      const float2 x2 = *reinterpret_cast<const float2*>(ra + kt * 16 + 2 * t + 8);
      const float2 x3 = *reinterpret_cast<const float2*>(rb + kt * 16 + 2 * t + 8);
      fv[kt][0] = ja < T ? pack_h2(x0.x, x0.y) : 0u;
      fv[kt][1] = jb < T ? pack_h2(x1.x, x1.y) : 0u;
      fv[kt][2] = ja < T ? pack_h2(x2.x, x2.y) : 0u;
      fv[kt][3] = jb < T ? pack_h2(x3.x, x3.y) : 0u;
    }
  }
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  float wa = 0.f, wb = 0.f;  // relevance partial sums for rows ja / jb

  const float* dl = a.delta + (size_t(pb) * a.H + h) * T;
  const float* rr = a.r + size_t(pb) * T;
  const float* Abase = a.probs + size_t(blockIdx.y) * T * T;

  for (int i0 = 0; i0 < T; i0 += 64) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 32; idx += blockDim.x) {
      const int i = idx >> 5, c = (idx & 31) * 2;
      __half2 go = __floats2half2_rn(0.f, 0.f);
      float2 qq = make_float2(0.f, 0.f);
      if (i0 + i < T) {
        go = *reinterpret_cast<const __half2*>(a.dO16 + (size_t(pb) * T + i0 + i) * a.ld_do + h * HD + c);
        qq = *reinterpret_cast<const float2*>(kvbase + size_t(i0 + i) * 3 * d + c);
      }
      *reinterpret_cast<__half2*>(dOs + i * TS + c) = go;
      dOt[c * TS + i] = __low2half(go);
      dOt[(c + 1) * TS + i] = __high2half(go);
      Qt[c * TS + i] = __float2half_rn(qq.x);
      Qt[(c + 1) * TS + i] = __float2half_rn(qq.y);
    }
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
      const int i = idx >> 6, j = idx & 63;
      As[i * AS + j] = (i0 + i < T && jbase + j < T) ? Abase[size_t(i0 + i) * T + jbase + j] : 0.f;
    }
    if (threadIdx.x < 64) {
      const int i = i0 + threadIdx.x;
      Dl[threadIdx.x] = i < T ? dl[i] : 0.f;
      Rw[threadIdx.x] = i < T ? rr[i] : 0.f;
    }
    __syncthreads();

    // Gᵀ tile [16 keys x 64 queries]
    float gacc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      gacc[n][0] = gacc[n][1] = gacc[n][2] = gacc[n][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(dOs + (n * 8 + g) * TS + kt * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(dOs + (n * 8 + g) * TS + kt * 16 + 2 * t + 8);
        mma_16816(gacc[n], fv[kt], b0, b1);
      }
    }
    uint32_t fs[4][4], fa[4][4];
    const int wj = warp * 16 + g;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int i = n * 8 + 2 * t;  // local query index of c0 / c2 ; +1 for c1 / c3
      const float a00 = As[i * AS + wj], a01 = As[(i + 1) * AS + wj];
      const float a10 = As[i * AS + wj + 8], a11 = As[(i + 1) * AS + wj + 8];
      const float d0 = Dl[i], d1 = Dl[i + 1], r0 = Rw[i], r1 = Rw[i + 1];
      float x00 = gacc[n][0] * a00, x01 = gacc[n][1] * a01, x10 = gacc[n][2] * a10, x11 = gacc[n][3] * a11;
      if (a.positive_only) x00 = fmaxf(x00, 0.f), x01 = fmaxf(x01, 0.f), x10 = fmaxf(x10, 0.f), x11 = fmaxf(x11, 0.f);
      wa += r0 * x00 + r1 * x01;
      wb += r0 * x10 + r1 * x11;
      fs[n >> 1][(n & 1) * 2 + 0] = pack_h2(a00 * (gacc[n][0] - d0), a01 * (gacc[n][1] - d1));
      fs[n >> 1][(n & 1) * 2 + 1] = pack_h2(a10 * (gacc[n][2] - d0), a11 * (gacc[n][3] - d1));
      fa[n >> 1][(n & 1) * 2 + 0] = pack_h2(a00, a01);
      fa[n >> 1][(n & 1) * 2 + 1] = pack_h2(a10, a11);
    }
    if (a.need_dqkv) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
          const uint32_t q0 = *reinterpret_cast<const uint32_t*>(Qt + (n * 8 + g) * TS + kt * 16 + 2 * t);
          const uint32_t q1 = *reinterpret_cast<const uint32_t*>(Qt + (n * 8 + g) * TS + kt * 16 + 2 * t + 8);
          mma_16816(dk[n], fs[kt], q0, q1);
          const uint32_t o0 = *reinterpret_cast<const uint32_t*>(dOt + (n * 8 + g) * TS + kt * 16 + 2 * t);
          const uint32_t o1 = *reinterpret_cast<const uint32_t*>(dOt + (n * 8 + g) * TS + kt * 16 + 2 * t + 8);
          mma_16816(dv[n], fa[kt], o0, o1);
        }
      }
    }
  }
  // relevance: reduce over the 4 lanes of a quad (they hold different query columns of the same key row)
  wa += __shfl_xor_sync(0xffffffffu, wa, 1);
  wa += __shfl_xor_sync(0xffffffffu, wa, 2);
  wb += __shfl_xor_sync(0xffffffffu, wb, 1);
  wb += __shfl_xor_sync(0xffffffffu, wb, 2);
  if (t == 0) {
    float* wp = a.wpart + (size_t(pb) * a.H + h) * T;
    const float invH = 1.0f / a.H;
    if (ja < T) wp[ja] = wa * invH;
    if (jb < T) wp[jb] = wb * invH;
  }
  if (a.need_dqkv) {
    const size_t ld = size_t(a.splits) * 3 * d;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = h * HD + n * 8 + 2 * t;
      if (ja < T) {
        store_h2_split(a.dqkv16, (size_t(pb) * T + ja) * ld, d + col, 3 * d, a.splits, dk[n][0], dk[n][1]);
        store_h2_split(a.dqkv16, (size_t(pb) * T + ja) * ld, 2 * d + col, 3 * d, a.splits, dv[n][0], dv[n][1]);
      }
      if (jb < T) {
        store_h2_split(a.dqkv16, (size_t(pb) * T + jb) * ld, d + col, 3 * d, a.splits, dk[n][2], dk[n][3]);
        store_h2_split(a.dqkv16, (size_t(pb) * T + jb) * ld, 2 * d + col, 3 * d, a.splits, dv[n][2], dv[n][3]);
      }
    }
  }
}

template <int NCHUNK>
static int launch_attn_fwd(const float* qkv, float* probs, float* o32, __half* o16, int B, int T, int H, int d,
                           int causal, int splits, cudaStream_t st) {
  const size_t smem = (size_t(2) * T * KV_STRIDE + size_t(8) * NCHUNK * 32) * sizeof(float);
  SB_REQUIRE(smem <= 227 * 1024, "attention forward: T=%d does not fit in shared memory", T);
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NCHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  attn_fwd_kernel<NCHUNK><<<B * H, 256, smem, st>>>(qkv, probs, o32, o16, T, H, d, causal, splits);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_attn_fwd(const float* qkv, float* probs, float* o32, void* o16, int32_t B, int32_t T, int32_t H,
                               int32_t causal, int32_t splits, void* stream) {
  SB_REQUIRE(qkv && (o32 || o16) && B > 0 && T > 0 && H > 0, "semabs_attn_fwd: bad arguments");
  const int d = H * HD;
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunk = (T + 31) / 32;
  if (nchunk <= 2) return launch_attn_fwd<2>(qkv, probs, o32, (__half*)o16, B, T, H, d, causal, splits, st);
  if (nchunk <= 3) return launch_attn_fwd<3>(qkv, probs, o32, (__half*)o16, B, T, H, d, causal, splits, st);
  if (nchunk <= 9) return launch_attn_fwd<9>(qkv, probs, o32, (__half*)o16, B, T, H, d, causal, splits, st);
  SB_REQUIRE(nchunk <= 13, "semabs_attn_fwd: T=%d > 416 tokens is not supported", T);
  return launch_attn_fwd<13>(qkv, probs, o32, (__half*)o16, B, T, H, d, causal, splits, st);
}

extern "C" int semabs_attn_bwd(const float* qkv, const float* probs, const float* o32, const void* dO16, int32_t ld_do,
                               const float* r, float* delta_ws, float* wpart, void* dqkv16, int32_t P, int32_t B,
                               int32_t T, int32_t H, int32_t splits, int32_t positive_only, int32_t need_dqkv,
                               void* stream) {
  SB_REQUIRE(qkv && probs && o32 && dO16 && r && delta_ws && wpart, "semabs_attn_bwd: null pointer");
  SB_REQUIRE(!need_dqkv || dqkv16, "semabs_attn_bwd: dqkv16 missing");
  SB_REQUIRE(P > 0 && B > 0 && T > 0 && H > 0 && B * H <= 65535, "semabs_attn_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = H * HD;
  const int PB = P * B;
  {
    const long long warps = (long long)PB * T * H;
    attn_bwd_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>((const __half*)dO16, ld_do, o32, delta_ws, PB, B,
                                                                       T, H, d);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  AttnBwdArgs a{};
  a.qkv = qkv, a.probs = probs, a.dO16 = (const __half*)dO16, a.ld_do = ld_do, a.delta = delta_ws, a.r = r;
  a.wpart = wpart, a.dqkv16 = (__half*)dqkv16, a.P = P, a.B = B, a.T = T, a.H = H, a.d = d, a.splits = splits;
  a.scale = 0.125f, a.positive_only = positive_only, a.need_dqkv = need_dqkv;
  const int nblk = (T + 63) / 64;
  dim3 grid(P * nblk, B * H);
  if (need_dqkv) {
    attn_bwd_dq_kernel<<<grid, 128, 0, st>>>(a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  attn_bwd_dkv_kernel<<<grid, 128, 0, st>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
