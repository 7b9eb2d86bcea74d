// Non-GEMM stages of the CLIP ViT / text transformer forward and of the hand-written backward that replaces
// torch.autograd.grad in ClipGradcam.interpret (reference: CLIP/clip/clip_gradcam.py:70-132).
// Everything here is bandwidth-bound row work: one warp (or CTA) per row, warp-shuffle reductions, fp32 math,
// fp16 (hi | lo split) copies emitted for the tcgen05 GEMM that consumes the row next.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

__device__ __forceinline__ void store_split(__half* row16, int col, int width, int splits, float v) {
  __half h = __float2half_rn(v);
  row16[col] = h;
  if (splits == 2) row16[width + col] = __float2half_rn(v - __half2float(h));
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 32 floats of shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ---------------------------------------------------------------------------------------------------------
// Patch embedding operand: conv1 has stride == kernel (model_explainability.py:304-310) so im2col is a pure
// gather. out[b*g*g + gy*g + gx, (c*p + i)*p + j] = tiles[b, c, gy*p + i, gx*p + j]   (zero padded to Kpad)
// ---------------------------------------------------------------------------------------------------------
__global__ void im2col_patch_kernel(const float* __restrict__ tiles, __half* __restrict__ out, int B, int R, int p,
                                    int g, int Kp, int Kpad, int splits) {
  const int row = blockIdx.x;  // b*g*g + gy*g + gx
  const int b = row / (g * g), gy = (row / g) % g, gx = row % g;
  __half* orow = out + size_t(row) * splits * Kpad;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < Kp) {
      const int c = k / (p * p), i = (k / p) % p, j = k % p;
      v = tiles[((size_t(b) * 3 + c) * R + gy * p + i) * R + gx * p + j];
    }
    store_split(orow, k, Kpad, splits, v);
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm forward (reference LayerNorm computes in fp32, eps 1e-5: model_explainability.py:188-194).
// Optional fused "token assembly" used for ln_pre: row (b,t) = (t == 0 ? cls : patch[b, t-1]) + pos[t]
// (VisionTransformer.forward, model_explainability.py:324-344).  One CTA per row.
// ---------------------------------------------------------------------------------------------------------
struct LnFwdArgs {
  const float* x;        // rows at x + row * x_stride (plain mode)
  long long x_stride;
  const float* patch;    // token-assembly mode if != nullptr: [B*(T-1), d]
  const float* cls;      // [d]
  const float* pos;      // [T, d]
  int T;
  const float* gamma;
  const float* beta;
  float* y32;            // optional fp32 output [M, d]
  __half* y16;           // optional fp16 output [M, splits*d]
  float* xsum;           // optional: pre-LN row (token-assembly mode) [M, d]
  float* mean;           // optional [M]
  float* rstd;           // optional [M]
  int M, d, splits;
};

__global__ void __launch_bounds__(256) layernorm_fwd_kernel(LnFwdArgs a) {
  __shared__ float red[32];
  extern __shared__ float rowbuf[];  // d floats
  const int row = blockIdx.x;
  if (a.patch) {
    const int b = row / a.T, t = row % a.T;
    const float* src = (t == 0) ? a.cls : a.patch + (size_t(b) * (a.T - 1) + (t - 1)) * a.d;
    const float* pos = a.pos + size_t(t) * a.d;
    for (int i = threadIdx.x; i < a.d; i += blockDim.x) rowbuf[i] = src[i] + pos[i];
  } else {
    const float* src = a.x + size_t(row) * a.x_stride;
    for (int i = threadIdx.x; i < a.d; i += blockDim.x) rowbuf[i] = src[i];
  }
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) s += rowbuf[i];
  const float mean = block_sum(s, red) / a.d;
  float v = 0.f;
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
    float c = rowbuf[i] - mean;
    v += c * c;
  }
  const float var = block_sum(v, red) / a.d;
  const float rstd = rsqrtf(var + 1e-5f);
  if (threadIdx.x == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
    const float y = (rowbuf[i] - mean) * rstd * a.gamma[i] + a.beta[i];
    if (a.y32) a.y32[size_t(row) * a.d + i] = y;
    if (a.y16) store_split(a.y16 + size_t(row) * a.splits * a.d, i, a.d, a.splits, y);
    if (a.xsum) a.xsum[size_t(row) * a.d + i] = rowbuf[i];
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward w.r.t. its input, for P stacked cotangents that share one forward row:
//   g = dy * gamma ; dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ dres]
// forward row of output row r is (r % x_rows).  Output rows may be strided (used to scatter the CLS-token
// gradient of ln_post into the [P*B, T, d] token grid).
// ---------------------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const float* dy;       // [M, d] fp32, or null when dy16 is given
  const __half* dy16;    // [M, d] fp16 (the dgrad GEMM's fp16 output: 2 instead of 4 bytes per element each way)
  const float* dres;     // optional [M, d] (row stride = out_stride)
  const float* x;        // forward input rows, x + (r % x_rows) * x_stride
  long long x_stride;
  int x_rows;
  const float* mean;     // [x_rows]
  const float* rstd;     // [x_rows]
  const float* gamma;
  float* dx32;           // [M rows, stride out_stride]
  long long out_stride;
  __half* dx16;          // optional, [M rows, stride out16_stride], hi | lo split at +d
  long long out16_stride;
  int M, d, splits;
};

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(LnBwdArgs a) {
  __shared__ float red[32];
  extern __shared__ float buf[];  // 2*d floats: g, xhat
  float* g = buf;
  float* xh = buf + a.d;
  const int row = blockIdx.x;
  const int fr = row % a.x_rows;
  const float mean = a.mean[fr], rstd = a.rstd[fr];
  const float* x = a.x + size_t(fr) * a.x_stride;
  const float* dy = a.dy ? a.dy + size_t(row) * a.d : nullptr;
  const __half* dyh = a.dy16 ? a.dy16 + size_t(row) * a.d : nullptr;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
    const float gi = (dy ? dy[i] : __half2float(dyh[i])) * a.gamma[i];
    const float xi = (x[i] - mean) * rstd;
    g[i] = gi;
    xh[i] = xi;
    s1 += gi;
    s2 += gi * xi;
  }
  const float m1 = block_sum(s1, red) / a.d;
  const float m2 = block_sum(s2, red) / a.d;
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
    float v = rstd * (g[i] - m1 - xh[i] * m2);
    if (a.dres) v += a.dres[size_t(row) * a.out_stride + i];
    a.dx32[size_t(row) * a.out_stride + i] = v;
    if (a.dx16) store_split(a.dx16 + size_t(row) * a.out16_stride, i, a.d, a.splits, v);
  }
}


// ---------------------------------------------------------------------------------------------------------
// Warp-per-row LayerNorm forward / backward for d = 128 * NV (512, 768, 1024): float4 per lane, the row lives in
// registers, reductions are pure warp shuffles (no block barrier).  Same math as the generic kernels above.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split4(__half* row16, int col, int width, int splits, const float4& v) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(row16 + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  if (splits == 2) {
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    *reinterpret_cast<uint2*>(row16 + width + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}

template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_warp_kernel(LnFwdArgs a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= a.M) return;
  const float* src = a.x + size_t(row) * a.x_stride;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k] = *reinterpret_cast<const float4*>(src + (k * 32 + lane) * 4);
    s += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  const float mean = warp_sum(s) / a.d;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float c0 = v[k].x - mean, c1 = v[k].y - mean, c2 = v[k].z - mean, c3 = v[k].w - mean;
    q += c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3;
  }
  const float rstd = rsqrtf(warp_sum(q) / a.d + 1e-5f);
  if (lane == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (k * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(a.gamma + col), b = *reinterpret_cast<const float4*>(a.beta + col);
    float4 y;
    y.x = (v[k].x - mean) * rstd * g.x + b.x, y.y = (v[k].y - mean) * rstd * g.y + b.y;
    y.z = (v[k].z - mean) * rstd * g.z + b.z, y.w = (v[k].w - mean) * rstd * g.w + b.w;
    if (a.y32) *reinterpret_cast<float4*>(a.y32 + size_t(row) * a.d + col) = y;
    if (a.y16) store_split4(a.y16 + size_t(row) * a.splits * a.d, col, a.d, a.splits, y);
  }
}

template <int NV>
__global__ void __launch_bounds__(128, 4) layernorm_bwd_warp_kernel(LnBwdArgs a) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= a.M) return;
  const int fr = row % a.x_rows;
  const float mean = a.mean[fr], rstd = a.rstd[fr];
  const float* x = a.x + size_t(fr) * a.x_stride;
  const float* dy = a.dy ? a.dy + size_t(row) * a.d : nullptr;
  const __half* dyh = a.dy16 ? a.dy16 + size_t(row) * a.d : nullptr;
  float4 g[NV], xh[NV];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (k * 32 + lane) * 4;
    float4 d4;
    if (dy) {
      d4 = *reinterpret_cast<const float4*>(dy + col);
    } else {
      const uint2 u = *reinterpret_cast<const uint2*>(dyh + col);
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      d4 = make_float4(f0.x, f0.y, f1.x, f1.y);
    }
    const float4 w4 = *reinterpret_cast<const float4*>(a.gamma + col);
    const float4 x4 = *reinterpret_cast<const float4*>(x + col);
    g[k] = make_float4(d4.x * w4.x, d4.y * w4.y, d4.z * w4.z, d4.w * w4.w);
    xh[k] = make_float4((x4.x - mean) * rstd, (x4.y - mean) * rstd, (x4.z - mean) * rstd, (x4.w - mean) * rstd);
    s1 += g[k].x + g[k].y + g[k].z + g[k].w;
    s2 += g[k].x * xh[k].x + g[k].y * xh[k].y + g[k].z * xh[k].z + g[k].w * xh[k].w;
  }
  // the residual-stream gradient is fetched here, with the other loads still in flight: read inside the store loop
  // below, each of its NV loads waited behind the previous store (possible aliasing) = NV serial DRAM round trips
  float4 rs[NV];
  if (a.dres) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
      rs[k] = *reinterpret_cast<const float4*>(a.dres + size_t(row) * a.out_stride + (k * 32 + lane) * 4);
  }
  const float m1 = warp_sum(s1) / a.d, m2 = warp_sum(s2) / a.d;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (k * 32 + lane) * 4;
    float4 v;
    v.x = rstd * (g[k].x - m1 - xh[k].x * m2), v.y = rstd * (g[k].y - m1 - xh[k].y * m2);
    v.z = rstd * (g[k].z - m1 - xh[k].z * m2), v.w = rstd * (g[k].w - m1 - xh[k].w * m2);
    if (a.dres) v.x += rs[k].x, v.y += rs[k].y, v.z += rs[k].z, v.w += rs[k].w;
    *reinterpret_cast<float4*>(a.dx32 + size_t(row) * a.out_stride + col) = v;
    if (a.dx16) store_split4(a.dx16 + size_t(row) * a.out16_stride, col, a.d, a.splits, v);
  }
}

template <typename Args, typename K4, typename K6, typename K8>
static bool launch_warp_ln(const Args& a, int d, cudaStream_t st, K4 k4, K6 k6, K8 k8, int rows_per_cta = 8) {
  const int grid = (a.M + rows_per_cta - 1) / rows_per_cta, threads = rows_per_cta * 32;
  if (d == 512) k4<<<grid, threads, 0, st>>>(a);
  else if (d == 768) k6<<<grid, threads, 0, st>>>(a);
  else if (d == 1024) k8<<<grid, threads, 0, st>>>(a);
  else return false;
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// CLIP logits and the backward seed (ClipGradcam.forward, clip_gradcam.py:58-68):
//   fhat = f / |f| ; logit[b,p] = 100 * fhat . W[:,p]
//   d logit[b,p] / d f = 100 * (W_p - fhat (fhat . W_p)) / |f|      -> seed16[(p*B + b), :]
// one warp per (p, b)
// ---------------------------------------------------------------------------------------------------------
__global__ void clip_logit_seed_kernel(const float* __restrict__ f, const float* __restrict__ W /*[E,P]*/,
                                       float* __restrict__ logits /*[B,P]*/, __half* __restrict__ seed16, int B, int P,
                                       int E, int splits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P * B) return;
  const int p = warp / B, b = warp % B;
  const float* fr = f + size_t(b) * E;
  float nn = 0.f, dot = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float v = fr[e];
    nn += v * v;
    dot += v * W[size_t(e) * P + p];
  }
  nn = warp_sum(nn);
  dot = warp_sum(dot);
  const float norm = sqrtf(nn);
  const float inv = 1.0f / norm;
  const float c = dot * inv;  // fhat . W_p
  if (lane == 0 && logits) logits[size_t(b) * P + p] = 100.0f * c;
  if (seed16) {
    __half* srow = seed16 + size_t(warp) * splits * E;
    for (int e = lane; e < E; e += 32) {
      const float v = 100.0f * (W[size_t(e) * P + p] - fr[e] * inv * c) * inv;
      store_split(srow, e, E, splits, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Text side (CLIP.encode_text, model_explainability.py:468-482; zeroshot_classifier, clip_gradcam.py:12-27)
// ---------------------------------------------------------------------------------------------------------
__global__ void token_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ table,
                                   const float* __restrict__ pos, float* __restrict__ x, int n_rows, int ctx, int d) {
  const int row = blockIdx.x;
  const int t = row % ctx;
  const float* e = table + size_t(tokens[row]) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) x[size_t(row) * d + i] = e[i] + pos[size_t(t) * d + i];
}

// W[e, c] = mean_t( feat[c*nt + t, e] / |feat[c*nt + t, :]| ) ; one CTA per class
__global__ void zeroshot_weights_kernel(const float* __restrict__ feat, float* __restrict__ W, int n_classes, int nt,
                                        int E) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  for (int e0 = 0; e0 < E; e0 += blockDim.x) {
    // accumulate over templates; E is small so recomputing the norms per chunk is fine
    float acc = 0.f;
    for (int t = 0; t < nt; ++t) {
      const float* fr = feat + (size_t(c) * nt + t) * E;
      float s = 0.f;
      for (int e = threadIdx.x; e < E; e += blockDim.x) s += fr[e] * fr[e];
      const float norm = sqrtf(block_sum(s, red));
      const int e = e0 + threadIdx.x;
      if (e < E) acc += fr[e] / norm;
    }
    const int e = e0 + threadIdx.x;
    if (e < E) W[size_t(e) * n_classes + c] = acc / nt;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Relevance rollout, row-vector form of R <- R + cam R restricted to row 0 (clip_gradcam.py:124-127):
//   r[pb, j] += sum_h wpart[pb, h, j]     (wpart already holds sum_i r_i * mean-weighted relu(G*A)[i,j])
// ---------------------------------------------------------------------------------------------------------
__global__ void rollout_update_kernel(float* __restrict__ r, const float* __restrict__ wpart, int PB, int H, int T) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= PB * T) return;
  const int pb = idx / T, j = idx % T;
  float s = 0.f;
  for (int h = 0; h < H; ++h) s += wpart[(size_t(pb) * H + h) * T + j];
  r[idx] += s;
}

__global__ void rollout_init_kernel(float* __restrict__ r, int PB, int T) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < PB * T) r[idx] = (idx % T == 0) ? 1.f : 0.f;
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_vit_im2col(const float* tiles, void* out16, int32_t B, int32_t R, int32_t patch, int32_t Kpad,
                                 int32_t splits, void* stream) {
  SB_REQUIRE(tiles && out16 && B > 0 && R % patch == 0, "semabs_vit_im2col: bad arguments");
  const int g = R / patch, Kp = 3 * patch * patch;
  SB_REQUIRE(Kpad >= Kp && Kpad % 8 == 0, "semabs_vit_im2col: Kpad %d too small for %d", Kpad, Kp);
  im2col_patch_kernel<<<B * g * g, 128, 0, (cudaStream_t)stream>>>(tiles, (__half*)out16, B, R, patch, g, Kp, Kpad,
                                                                   splits);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_layernorm_fwd(const float* x, int64_t x_stride, const float* gamma, const float* beta,
                                    float* y32, void* y16, float* mean, float* rstd, int32_t M, int32_t d,
                                    int32_t splits, void* stream) {
  SB_REQUIRE(x && gamma && beta && M > 0 && d > 0 && (y32 || y16), "semabs_layernorm_fwd: bad arguments");
  LnFwdArgs a{};
  a.x = x, a.x_stride = x_stride, a.gamma = gamma, a.beta = beta, a.y32 = y32, a.y16 = (__half*)y16;
  a.mean = mean, a.rstd = rstd, a.M = M, a.d = d, a.splits = splits;
  const bool aligned = (x_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (!(aligned && launch_warp_ln(a, d, (cudaStream_t)stream, layernorm_fwd_warp_kernel<4>, layernorm_fwd_warp_kernel<6>,
                                  layernorm_fwd_warp_kernel<8>)))
    layernorm_fwd_kernel<<<M, 256, d * sizeof(float), (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_vit_embed_lnpre(const float* patch, const float* cls, const float* pos, const float* gamma,
                                      const float* beta, float* x_out, int32_t B, int32_t T, int32_t d, void* stream) {
  SB_REQUIRE(patch && cls && pos && gamma && beta && x_out && B > 0 && T > 1, "semabs_vit_embed_lnpre: bad arguments");
  LnFwdArgs a{};
  a.patch = patch, a.cls = cls, a.pos = pos, a.T = T, a.gamma = gamma, a.beta = beta, a.y32 = x_out;
  a.M = B * T, a.d = d, a.splits = 1;
  layernorm_fwd_kernel<<<B * T, 256, d * sizeof(float), (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int layernorm_bwd_impl(const float* dy, const void* dy16, const float* dres, const float* x, int64_t x_stride,
                              int32_t x_rows, const float* mean, const float* rstd, const float* gamma,
                              float* dx32, int64_t out_stride, void* dx16, int64_t out16_stride, int32_t M,
                              int32_t d, int32_t splits, void* stream) {
  SB_REQUIRE((dy || dy16) && x && mean && rstd && gamma && dx32 && M > 0 && x_rows > 0, "semabs_layernorm_bwd: bad arguments");
  LnBwdArgs a{};
  a.dy = dy, a.dy16 = (const __half*)dy16, a.dres = dres, a.x = x, a.x_stride = x_stride, a.x_rows = x_rows, a.mean = mean, a.rstd = rstd;
  a.gamma = gamma, a.dx32 = dx32, a.out_stride = out_stride, a.dx16 = (__half*)dx16, a.out16_stride = out16_stride;
  a.M = M, a.d = d, a.splits = splits;
  const bool aligned = (x_stride % 4 == 0) && (out_stride % 4 == 0) && (out16_stride % 4 == 0);
  if (!(aligned && launch_warp_ln(a, d, (cudaStream_t)stream, layernorm_bwd_warp_kernel<4>, layernorm_bwd_warp_kernel<6>,
                                  layernorm_bwd_warp_kernel<8>, 4)))
    layernorm_bwd_kernel<<<M, 256, 2 * d * sizeof(float), (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_layernorm_bwd(const float* dy, const float* dres, const float* x, int64_t x_stride,
                                    int32_t x_rows, const float* mean, const float* rstd, const float* gamma,
                                    float* dx32, int64_t out_stride, void* dx16, int64_t out16_stride, int32_t M,
                                    int32_t d, int32_t splits, void* stream) {
  return layernorm_bwd_impl(dy, nullptr, dres, x, x_stride, x_rows, mean, rstd, gamma, dx32, out_stride, dx16, out16_stride, M, d,
                            splits, stream);
}

// same operator with the cotangent given as fp16 rows [M, d] (what the dgrad GEMM emits with out_f16)
extern "C" int semabs_layernorm_bwd_h(const void* dy16, const float* dres, const float* x, int64_t x_stride,
                                      int32_t x_rows, const float* mean, const float* rstd, const float* gamma,
                                      float* dx32, int64_t out_stride, void* dx16, int64_t out16_stride, int32_t M,
                                      int32_t d, int32_t splits, void* stream) {
  return layernorm_bwd_impl(nullptr, dy16, dres, x, x_stride, x_rows, mean, rstd, gamma, dx32, out_stride, dx16, out16_stride, M, d,
                            splits, stream);
}

extern "C" int semabs_clip_logit_seed(const float* f, const float* W, float* logits, void* seed16, int32_t B,
                                      int32_t P, int32_t E, int32_t splits, void* stream) {
  SB_REQUIRE(f && W && (logits || seed16) && B > 0 && P > 0 && E > 0, "semabs_clip_logit_seed: bad arguments");
  const int warps = P * B;
  clip_logit_seed_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(f, W, logits, (__half*)seed16, B, P, E,
                                                                            splits);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_text_embed(const int32_t* tokens, const float* table, const float* pos, float* x,
                                 int32_t n_texts, int32_t ctx, int32_t d, void* stream) {
  SB_REQUIRE(tokens && table && pos && x && n_texts > 0, "semabs_text_embed: bad arguments");
  token_embed_kernel<<<n_texts * ctx, 128, 0, (cudaStream_t)stream>>>(tokens, table, pos, x, n_texts * ctx, ctx, d);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_zeroshot_weights(const float* feat, float* W, int32_t n_classes, int32_t n_templates, int32_t E,
                                       void* stream) {
  SB_REQUIRE(feat && W && n_classes > 0 && n_templates > 0, "semabs_zeroshot_weights: bad arguments");
  zeroshot_weights_kernel<<<n_classes, 256, 0, (cudaStream_t)stream>>>(feat, W, n_classes, n_templates, E);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_rollout_init(float* r, int32_t PB, int32_t T, void* stream) {
  SB_REQUIRE(r && PB > 0 && T > 0, "semabs_rollout_init: bad arguments");
  rollout_init_kernel<<<(PB * T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(r, PB, T);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_rollout_update(float* r, const float* wpart, int32_t PB, int32_t H, int32_t T, void* stream) {
  SB_REQUIRE(r && wpart && PB > 0 && H > 0 && T > 0, "semabs_rollout_update: bad arguments");
  rollout_update_kernel<<<(PB * T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(r, wpart, PB, H, T);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
