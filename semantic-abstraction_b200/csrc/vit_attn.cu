// STATUS (round 2): of this file the product path uses `semabs_attn_fwd` for the text tower (T = 77, causal) and
// `attn_bwd_cls_kernel` (class-token-only backward of the top block).  The mma.sync backward passes below are the fall-back for
// T > 272 and the cross-check of the tcgen05 generations (tests/test_vit_kernels_gpu.py impl "mma"); the image tower runs on
// vit_attn_tc.cu (forward) and vit_attn_bwd3.cu (backward).
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
attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ probs, __half* __restrict__ probs16, int ld_p16,
                float* __restrict__ o32, __half* __restrict__ o16, int T, int H, int d, int causal, int splits) {
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
      if (probs16 && j < ld_p16) probs16[(size_t(blockIdx.x) * T + i) * ld_p16 + j] = __float2half_rn(a);
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
// All MMA operands are fp16 produced upstream: qkv16 [B*T, 3d] (QKV GEMM epilogue), probs16 [B*H, T, ldp]
// (attention forward, rows zero-padded to ldp >= 272), dO16 (out-proj dgrad GEMM epilogue).  One CTA owns one
// (label p, tile b, head h); its 6 warps split the 17 16-row MMA tiles of the 257 tokens.  Two passes so that
// nothing needs atomics (run-to-run deterministic):
//   row owner    : delta_i = dO_i . O_i, dA = dO V^T, dS = A ⊙ (dA - delta), dQ = scale dS K
//   column owner : dA^T = V dO^T, relevance_j = sum_i r_i relu(dA ⊙ A)_ij / H, dK = dS^T Q, dV = A^T dO
// The whole head's K,V (resp. Q,dO) stay in shared memory; A is streamed from L2 (p is the fastest grid index, so
// the P CTAs that share a (b,h) pair run together).
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_row))));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }

constexpr int TS = 72;        // half stride of [token][64] tiles: 144 B rows, 16B-aligned, conflict-free fragment loads
constexpr int TR = 272;       // 17 MMA row tiles
constexpr int BWD_WARPS = 6;
constexpr int STRIP = 16;     // half stride of the per-warp [64][16] probability strip (two buffers per warp); the two
                              // 16-byte halves of a row are XOR-swizzled with bit 2 of the row => conflict-free ldmatrix

struct AttnBwdArgs {
  const __half* qkv16;   // [B*T, ldq >= 3d] (q already scaled)
  int ldq;
  const __half* probs16; // [B*H, T, ldp]
  int ldp;
  const float* o32;      // [B*T, d] forward attention output (pre out-proj)
  const __half* dO16;    // [P*B*T, ld_do]
  int ld_do;
  float* delta;          // [P*B*H, T] written by the row pass, read by the column pass
  const float* r;        // [P*B, T] current rollout row vector
  float* wpart;          // [P*B*H, T] out: per-head relevance contribution
  __half* dqkv16;        // [P*B*T, splits*3d] out
  int P, B, T, H, d, splits;
  float scale;
  int positive_only;
  int need_dqkv;
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

// rows [0,T) of a [T, 64] fp16 slice (row pitch `pitch` halfs) -> smem [TR][TS], rows >= T zeroed
__device__ __forceinline__ void load_head_tile(__half* dst, const __half* src, size_t pitch, int T) {
  for (int idx = threadIdx.x; idx < TR * 8; idx += blockDim.x) {
    const int row = idx >> 3, c = (idx & 7) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < T) v = *reinterpret_cast<const uint4*>(src + size_t(row) * pitch + c);
    *reinterpret_cast<uint4*>(dst + row * TS + c) = v;
  }
}

__global__ void __launch_bounds__(BWD_WARPS * 32, 2) attn_bwd_row_kernel(AttnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __half* Ks = reinterpret_cast<__half*>(smraw);
  __half* Vs = Ks + TR * TS;
  const int p = blockIdx.x, bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H;
  const int pb = p * a.B + b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int T = a.T, d = a.d;
  const __half* kv = a.qkv16 + size_t(b) * T * a.ldq + h * HD;
  load_head_tile(Ks, kv + d, size_t(a.ldq), T);
  load_head_tile(Vs, kv + 2 * d, size_t(a.ldq), T);
  __syncthreads();
  const int ntiles_total = (T + 7) / 8;  // key n-tiles that contain at least one valid key

  for (int mt = warp; mt * 16 < T; mt += BWD_WARPS) {
    const int ia = mt * 16 + g, ib = ia + 8;
    const bool va = ia < T, vb = ib < T;
    uint32_t fo[4][4];
    float da = 0.f, db = 0.f;
    {
      const __half* ra = a.dO16 + (size_t(pb) * T + (va ? ia : 0)) * a.ld_do + h * HD;
      const __half* rb = a.dO16 + (size_t(pb) * T + (vb ? ib : 0)) * a.ld_do + h * HD;
      const float* oa = a.o32 + (size_t(b) * T + (va ? ia : 0)) * d + h * HD;
      const float* ob = a.o32 + (size_t(b) * T + (vb ? ib : 0)) * d + h * HD;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int col = kt * 16 + 2 * t + 8 * hf;
          const uint32_t xa = va ? *reinterpret_cast<const uint32_t*>(ra + col) : 0u;
          const uint32_t xb = vb ? *reinterpret_cast<const uint32_t*>(rb + col) : 0u;
          fo[kt][2 * hf] = xa, fo[kt][2 * hf + 1] = xb;
          const float2 fa = unpack_h2(xa), fb = unpack_h2(xb);
          const float2 pa = *reinterpret_cast<const float2*>(oa + col), pb2 = *reinterpret_cast<const float2*>(ob + col);
          da += fa.x * pa.x + fa.y * pa.y;
          db += fb.x * pb2.x + fb.y * pb2.y;
        }
      }
      da += __shfl_xor_sync(0xffffffffu, da, 1), da += __shfl_xor_sync(0xffffffffu, da, 2);
      db += __shfl_xor_sync(0xffffffffu, db, 1), db += __shfl_xor_sync(0xffffffffu, db, 2);
      if (t == 0) {
        float* dl = a.delta + (size_t(pb) * a.H + h) * T;
        if (va) dl[ia] = da;
        if (vb) dl[ib] = db;
      }
    }
    if (!a.need_dqkv) continue;
    const __half* Arow_a = a.probs16 + (size_t(bh) * T + (va ? ia : 0)) * a.ldp;
    const __half* Arow_b = a.probs16 + (size_t(bh) * T + (vb ? ib : 0)) * a.ldp;
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;

    // probabilities of this warp's two rows are streamed from L2 one 64-key chunk AHEAD of their use (the loads
    // were the top stall of the first version: a dependent L2 round trip per 8-key tile)
    uint32_t cur_a[8], cur_b[8];
    auto load_probs = [&](int jc, uint32_t (&xa)[8], uint32_t (&xb)[8]) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int key = jc * 64 + n * 8 + 2 * t;
        const bool in = key < a.ldp;  // rows are zero-padded up to ldp
        xa[n] = (va && in) ? *reinterpret_cast<const uint32_t*>(Arow_a + key) : 0u;
        xb[n] = (vb && in) ? *reinterpret_cast<const uint32_t*>(Arow_b + key) : 0u;
      }
    };
    load_probs(0, cur_a, cur_b);
    for (int jc = 0; jc * 8 < ntiles_total; ++jc) {
      const int nt = min(8, ntiles_total - jc * 8);
      uint32_t na[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, nb[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if ((jc + 1) * 8 < ntiles_total) load_probs(jc + 1, na, nb);
      uint32_t fs[4][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float gacc[4] = {0.f, 0.f, 0.f, 0.f};
        float2 aa = make_float2(0.f, 0.f), ab = make_float2(0.f, 0.f);
        if (n < nt) {
          const int key = jc * 64 + n * 8;
#pragma unroll
          for (int kt = 0; kt < 4; ++kt) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(Vs + (key + g) * TS + kt * 16 + 2 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(Vs + (key + g) * TS + kt * 16 + 2 * t + 8);
            mma_16816(gacc, fo[kt], b0, b1);
          }
          aa = unpack_h2(cur_a[n]);
          ab = unpack_h2(cur_b[n]);
        }
        fs[n >> 1][(n & 1) * 2 + 0] = pack_h2(aa.x * (gacc[0] - da), aa.y * (gacc[1] - da));
        fs[n >> 1][(n & 1) * 2 + 1] = pack_h2(ab.x * (gacc[2] - db), ab.y * (gacc[3] - db));
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) cur_a[n] = na[n], cur_b[n] = nb[n];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        if (2 * kt < nt) {
          const int krow = jc * 64 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
          for (int nd2 = 0; nd2 < 4; ++nd2) {
            uint32_t bb[4];
            ldmatrix_x4_trans(bb, Ks + krow * TS + nd2 * 16 + (lane >> 4) * 8);
            mma_16816(dq[2 * nd2], fs[kt], bb[0], bb[1]);
            mma_16816(dq[2 * nd2 + 1], fs[kt], bb[2], bb[3]);
          }
        }
      }
    }
    const size_t ld = size_t(a.splits) * 3 * d;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = h * HD + n * 8 + 2 * t;
      if (va) store_h2_split(a.dqkv16, (size_t(pb) * T + ia) * ld, col, 3 * d, a.splits, dq[n][0] * a.scale, dq[n][1] * a.scale);
      if (vb) store_h2_split(a.dqkv16, (size_t(pb) * T + ib) * ld, col, 3 * d, a.splits, dq[n][2] * a.scale, dq[n][3] * a.scale);
    }
  }
}

__global__ void __launch_bounds__(BWD_WARPS * 32, 2) attn_bwd_col_kernel(AttnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __half* Qs = reinterpret_cast<__half*>(smraw);
  __half* dOs = Qs + TR * TS;
  float* Dl = reinterpret_cast<float*>(dOs + TR * TS);
  float* Rw = Dl + TR;
  __half* strips = reinterpret_cast<__half*>(Rw + TR);
  const int p = blockIdx.x, bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H;
  const int pb = p * a.B + b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int T = a.T, d = a.d;
  const __half* qv = a.qkv16 + size_t(b) * T * a.ldq + h * HD;
  load_head_tile(Qs, qv, size_t(a.ldq), T);
  load_head_tile(dOs, a.dO16 + size_t(pb) * T * a.ld_do + h * HD, size_t(a.ld_do), T);
  for (int i = threadIdx.x; i < TR; i += blockDim.x) {
    Dl[i] = (i < T && a.need_dqkv) ? a.delta[(size_t(pb) * a.H + h) * T + i] : 0.f;
    Rw[i] = i < T ? a.r[size_t(pb) * T + i] : 0.f;
  }
  __syncthreads();
  __half* strip_buf = strips + warp * 2 * 64 * STRIP;
  const int ntiles_total = (T + 7) / 8;
  const __half* Abase = a.probs16 + size_t(bh) * T * a.ldp;

  for (int mt = warp; mt * 16 < T; mt += BWD_WARPS) {
    const int j0 = mt * 16, ja = j0 + g, jb = ja + 8;
    const bool va = ja < T, vb = jb < T;
    uint32_t fv[4][4];
    {
      const __half* ra = qv + size_t(va ? ja : 0) * a.ldq + 2 * d;
      const __half* rb = qv + size_t(vb ? jb : 0) * a.ldq + 2 * d;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        fv[kt][0] = va ? *reinterpret_cast<const uint32_t*>(ra + kt * 16 + 2 * t) : 0u;
        fv[kt][1] = vb ? *reinterpret_cast<const uint32_t*>(rb + kt * 16 + 2 * t) : 0u;
        fv[kt][2] = va ? *reinterpret_cast<const uint32_t*>(ra + kt * 16 + 2 * t + 8) : 0u;
        fv[kt][3] = vb ? *reinterpret_cast<const uint32_t*>(rb + kt * 16 + 2 * t + 8) : 0u;
      }
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
      dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
    }
    float wa = 0.f, wb = 0.f;

    // [64 queries][16 keys] strips of A are double-buffered with cp.async: chunk ic+1 streams from L2 while chunk ic
    // is multiplied (zero-fill for query rows >= T)
    auto stage_strip = [&](int ic, __half* dstbuf) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = lane + 32 * k, row = c >> 1, hf = c & 1;
        const bool ok = ic * 64 + row < T;
        const __half* src = ok ? Abase + size_t(ic * 64 + row) * a.ldp + j0 + hf * 8 : Abase;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                         static_cast<uint32_t>(__cvta_generic_to_shared(dstbuf + row * STRIP + (hf ^ ((row >> 2) & 1)) * 8))),
                     "l"(src), "r"(ok ? 16u : 0u)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncwarp();
    stage_strip(0, strip_buf);
    for (int ic = 0; ic * 8 < ntiles_total; ++ic) {
      const int nt = min(8, ntiles_total - ic * 8);
      const int ibase = ic * 64;
      __half* strip = strip_buf + (ic & 1) * 64 * STRIP;
      if ((ic + 1) * 8 < ntiles_total) {
        stage_strip(ic + 1, strip_buf + ((ic + 1) & 1) * 64 * STRIP);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      uint32_t fs[4][4], fa[4][4];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        // A^T fragments (rows = keys, k = queries) through a transposed ldmatrix of the [query][key] strip
        const int srow = kt * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
        ldmatrix_x4_trans(fa[kt], strip + srow * STRIP + ((((lane >> 3) & 1)) ^ ((srow >> 2) & 1)) * 8);
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float gacc[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < nt) {
#pragma unroll
          for (int kt = 0; kt < 4; ++kt) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(dOs + (ibase + n * 8 + g) * TS + kt * 16 + 2 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(dOs + (ibase + n * 8 + g) * TS + kt * 16 + 2 * t + 8);
            mma_16816(gacc, fv[kt], b0, b1);
          }
        }
        // C layout of this n-tile: c0,c1 = (key ja; queries i, i+1), c2,c3 = (key jb; queries i, i+1)
        const int i = ibase + n * 8 + 2 * t;
        const float2 aa = unpack_h2(fa[n >> 1][(n & 1) * 2 + 0]);  // A[i, ja], A[i+1, ja]
        const float2 ab = unpack_h2(fa[n >> 1][(n & 1) * 2 + 1]);  // A[i, jb], A[i+1, jb]
        const float d0 = Dl[i], d1 = Dl[i + 1], r0 = Rw[i], r1 = Rw[i + 1];
        float x00 = gacc[0] * aa.x, x01 = gacc[1] * aa.y, x10 = gacc[2] * ab.x, x11 = gacc[3] * ab.y;
        if (a.positive_only) x00 = fmaxf(x00, 0.f), x01 = fmaxf(x01, 0.f), x10 = fmaxf(x10, 0.f), x11 = fmaxf(x11, 0.f);
        wa += r0 * x00 + r1 * x01;
        wb += r0 * x10 + r1 * x11;
        fs[n >> 1][(n & 1) * 2 + 0] = pack_h2(aa.x * (gacc[0] - d0), aa.y * (gacc[1] - d1));
        fs[n >> 1][(n & 1) * 2 + 1] = pack_h2(ab.x * (gacc[2] - d0), ab.y * (gacc[3] - d1));
      }
      if (a.need_dqkv) {
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
          if (2 * kt < nt) {
            const int krow = ibase + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int nd2 = 0; nd2 < 4; ++nd2) {
              uint32_t bq[4], bo[4];
              ldmatrix_x4_trans(bq, Qs + krow * TS + nd2 * 16 + (lane >> 4) * 8);
              ldmatrix_x4_trans(bo, dOs + krow * TS + nd2 * 16 + (lane >> 4) * 8);
              mma_16816(dk[2 * nd2], fs[kt], bq[0], bq[1]);
              mma_16816(dk[2 * nd2 + 1], fs[kt], bq[2], bq[3]);
              mma_16816(dv[2 * nd2], fa[kt], bo[0], bo[1]);
              mma_16816(dv[2 * nd2 + 1], fa[kt], bo[2], bo[3]);
            }
          }
        }
      }
    }
    wa += __shfl_xor_sync(0xffffffffu, wa, 1), wa += __shfl_xor_sync(0xffffffffu, wa, 2);
    wb += __shfl_xor_sync(0xffffffffu, wb, 1), wb += __shfl_xor_sync(0xffffffffu, wb, 2);
    if (t == 0) {
      float* wp = a.wpart + (size_t(pb) * a.H + h) * T;
      const float invH = 1.0f / a.H;
      if (va) wp[ja] = wa * invH;
      if (vb) wp[jb] = wb * invH;
    }
    if (a.need_dqkv) {
      const size_t ld = size_t(a.splits) * 3 * d;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int col = h * HD + n * 8 + 2 * t;
        if (va) {
          store_h2_split(a.dqkv16, (size_t(pb) * T + ja) * ld, d + col, 3 * d, a.splits, dk[n][0], dk[n][1]);
          store_h2_split(a.dqkv16, (size_t(pb) * T + ja) * ld, 2 * d + col, 3 * d, a.splits, dv[n][0], dv[n][1]);
        }
        if (vb) {
          store_h2_split(a.dqkv16, (size_t(pb) * T + jb) * ld, d + col, 3 * d, a.splits, dk[n][2], dk[n][3]);
          store_h2_split(a.dqkv16, (size_t(pb) * T + jb) * ld, 2 * d + col, 3 * d, a.splits, dv[n][2], dv[n][3]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward on mma.sync with 3-term fp16 splits (hi*hi + hi*lo + lo*hi, fp32 accumulate ≈ fp32 accuracy):
// one CTA per (tile, head); K and V of the head live in shared memory as hi/lo fp16 pairs; each warp owns 16-row
// query tiles, keeps the whole score strip [16 x T] in registers, does an exact two-pass softmax and multiplies
// by V.  Used for T <= 272 (every CLIP ViT-B/32, ViT-L/14 and text case); the SIMT kernel above is the fallback.
// ---------------------------------------------------------------------------------------------------------
constexpr int FWD_WARPS = 6;   // (9 warps would cap registers at 224/thread and spill the score strip)
constexpr int NT_MAX = TR / 8;  // 34 key n-tiles

struct AttnFwdArgs {
  const float* qkv;
  float* probs;
  __half* probs16;
  int ldp;
  float* o32;
  __half* o16;
  int T, H, d, causal, splits;
};

__device__ __forceinline__ void split_h2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(FWD_WARPS * 32, 1) attn_fwd_mma_kernel(AttnFwdArgs a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __half* Kh = reinterpret_cast<__half*>(smraw);
  __half* Kl = Kh + TR * TS;
  __half* Vh = Kl + TR * TS;
  __half* Vl = Vh + TR * TS;
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int T = a.T, d = a.d;
  const float* base = a.qkv + size_t(b) * T * 3 * d + h * HD;
  for (int idx = threadIdx.x; idx < TR * 32; idx += blockDim.x) {
    const int row = idx >> 5, c = (idx & 31) * 2;
    float2 kk = make_float2(0.f, 0.f), vv = make_float2(0.f, 0.f);
    if (row < T) {
      kk = *reinterpret_cast<const float2*>(base + size_t(row) * 3 * d + d + c);
      vv = *reinterpret_cast<const float2*>(base + size_t(row) * 3 * d + 2 * d + c);
    }
    uint32_t hi, lo;
    split_h2(kk.x, kk.y, hi, lo);
    *reinterpret_cast<uint32_t*>(Kh + row * TS + c) = hi, *reinterpret_cast<uint32_t*>(Kl + row * TS + c) = lo;
    split_h2(vv.x, vv.y, hi, lo);
    *reinterpret_cast<uint32_t*>(Vh + row * TS + c) = hi, *reinterpret_cast<uint32_t*>(Vl + row * TS + c) = lo;
  }
  __syncthreads();
  const int ntiles = (T + 7) / 8;

  for (int mt = warp; mt * 16 < T; mt += FWD_WARPS) {
    const int ia = mt * 16 + g, ib = ia + 8;
    const bool va = ia < T, vb = ib < T;
    float s[NT_MAX][4];
    {
      uint32_t qh[4][4], ql[4][4];
      const float* ra = base + size_t(va ? ia : 0) * 3 * d;
      const float* rb = base + size_t(vb ? ib : 0) * 3 * d;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int col = kt * 16 + 2 * t + 8 * hf;
          const float2 xa = va ? *reinterpret_cast<const float2*>(ra + col) : make_float2(0.f, 0.f);
          const float2 xb = vb ? *reinterpret_cast<const float2*>(rb + col) : make_float2(0.f, 0.f);
          split_h2(xa.x, xa.y, qh[kt][2 * hf], ql[kt][2 * hf]);
          split_h2(xb.x, xb.y, qh[kt][2 * hf + 1], ql[kt][2 * hf + 1]);
        }
#pragma unroll
      for (int n = 0; n < NT_MAX; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        if (n < ntiles) {
#pragma unroll
          for (int kt = 0; kt < 4; ++kt) {
            const int off = (n * 8 + g) * TS + kt * 16 + 2 * t;
            const uint32_t h0 = *reinterpret_cast<const uint32_t*>(Kh + off), h1 = *reinterpret_cast<const uint32_t*>(Kh + off + 8);
            const uint32_t l0 = *reinterpret_cast<const uint32_t*>(Kl + off), l1 = *reinterpret_cast<const uint32_t*>(Kl + off + 8);
            mma_16816(s[n], qh[kt], h0, h1);
            mma_16816(s[n], qh[kt], l0, l1);
            mma_16816(s[n], ql[kt], h0, h1);
          }
        }
      }
    }
    // mask + exact softmax over the row (rows ia: c0,c1 ; ib: c2,c3)
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = n * 8 + 2 * t + e;
        const bool dead = j >= T;
        if (dead || (a.causal && j > ia)) s[n][e] = -INFINITY;
        if (dead || (a.causal && j > ib)) s[n][2 + e] = -INFINITY;
      }
      ma = fmaxf(ma, fmaxf(s[n][0], s[n][1]));
      mb = fmaxf(mb, fmaxf(s[n][2], s[n][3]));
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)), ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)), mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float sa = 0.f, sb2 = 0.f;
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] = (s[n][e] == -INFINITY) ? 0.f : expf(s[n][e] - ma);
        s[n][2 + e] = (s[n][2 + e] == -INFINITY) ? 0.f : expf(s[n][2 + e] - mb);
      }
      sa += s[n][0] + s[n][1];
      sb2 += s[n][2] + s[n][3];
    }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1), sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb2 += __shfl_xor_sync(0xffffffffu, sb2, 1), sb2 += __shfl_xor_sync(0xffffffffu, sb2, 2);
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
      s[n][0] /= sa, s[n][1] /= sa, s[n][2] /= sb2, s[n][3] /= sb2;
      const int j = n * 8 + 2 * t;
      if (a.probs16 && j < a.ldp) {
        if (va) *reinterpret_cast<__half2*>(a.probs16 + (size_t(bh) * T + ia) * a.ldp + j) = __floats2half2_rn(s[n][0], s[n][1]);
        if (vb) *reinterpret_cast<__half2*>(a.probs16 + (size_t(bh) * T + ib) * a.ldp + j) = __floats2half2_rn(s[n][2], s[n][3]);
      }
      if (a.probs) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (va && j + e < T) a.probs[(size_t(bh) * T + ia) * T + j + e] = s[n][e];
          if (vb && j + e < T) a.probs[(size_t(bh) * T + ib) * T + j + e] = s[n][2 + e];
        }
      }
    }
    // O = A V
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NT_MAX / 2; ++kt) {
      if (2 * kt < ntiles) {
        uint32_t ah[4], al[4];
        split_h2(s[2 * kt][0], s[2 * kt][1], ah[0], al[0]);
        split_h2(s[2 * kt][2], s[2 * kt][3], ah[1], al[1]);
        split_h2(s[2 * kt + 1][0], s[2 * kt + 1][1], ah[2], al[2]);
        split_h2(s[2 * kt + 1][2], s[2 * kt + 1][3], ah[3], al[3]);
        const int krow = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
          uint32_t bhv[4], blv[4];
          ldmatrix_x4_trans(bhv, Vh + krow * TS + nd2 * 16 + (lane >> 4) * 8);
          ldmatrix_x4_trans(blv, Vl + krow * TS + nd2 * 16 + (lane >> 4) * 8);
          mma_16816(o[2 * nd2], ah, bhv[0], bhv[1]);
          mma_16816(o[2 * nd2], ah, blv[0], blv[1]);
          mma_16816(o[2 * nd2], al, bhv[0], bhv[1]);
          mma_16816(o[2 * nd2 + 1], ah, bhv[2], bhv[3]);
          mma_16816(o[2 * nd2 + 1], ah, blv[2], blv[3]);
          mma_16816(o[2 * nd2 + 1], al, bhv[2], bhv[3]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = h * HD + n * 8 + 2 * t;
      if (va) {
        const size_t row = size_t(b) * T + ia;
        if (a.o32) *reinterpret_cast<float2*>(a.o32 + row * d + col) = make_float2(o[n][0], o[n][1]);
        if (a.o16) store_h2_split(a.o16, row * size_t(a.splits) * d, col, d, a.splits, o[n][0], o[n][1]);
      }
      if (vb) {
        const size_t row = size_t(b) * T + ib;
        if (a.o32) *reinterpret_cast<float2*>(a.o32 + row * d + col) = make_float2(o[n][2], o[n][3]);
        if (a.o16) store_h2_split(a.o16, row * size_t(a.splits) * d, col, d, a.splits, o[n][2], o[n][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the LAST block: the cotangent of the block output is non-zero at the class token only
// (logits depend on x[:, 0] alone, model_explainability.py:349), so dO has a single non-zero query row and the whole
// attention backward collapses to vector work: one warp per (label, tile, head), lanes over keys.
//   dA_0j = dO_0 . V_j ; w_j = r_0 relu(dA_0j A_0j) / H ; dS_0j = A_0j (dA_0j - sum_j dA_0j A_0j)
//   dQ_0 = scale sum_j dS_0j K_j ; dK_j = dS_0j Q_0 ; dV_j = A_0j dO_0 ; dQ_i = 0 for i > 0
// ---------------------------------------------------------------------------------------------------------
struct AttnClsArgs {
  const __half* qkv16;   // [B*T, ldq >= 3d]
  int ldq;
  const __half* probs16; // [B*H, T, ldp]
  int ldp;
  const __half* dO16;    // [P*B, ld_do]  (class-token rows only)
  int ld_do;
  const float* r;        // [P*B, T]
  float* wpart;          // [P*B*H, T]
  __half* dqkv16;        // [P*B*T, splits*3d]
  int P, B, T, H, d, splits;
  float scale;
  int positive_only, need_dqkv;
};

// (Two CTAs per SM: the kernel is a latency-bound walk over the 257 keys of a (label, sequence, head) per warp — round 2 ncu
// showed 2.2 ms per launch with 8 warps per SM; Q_0 is no longer replicated in every lane, which was 64 of ~170 registers.)
__global__ void __launch_bounds__(256, 2) attn_bwd_cls_kernel(AttnClsArgs a) {
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= a.P * a.B * a.H) return;
  const int h = gw % a.H, pb = gw / a.H, b = pb % a.B;
  const int T = a.T, d = a.d;
  const __half* base = a.qkv16 + size_t(b) * T * a.ldq + h * HD;
  // dO_0 (64 values) replicated in registers for the dot products; of Q_0 a lane only needs its own channel pair
  float dO[HD];
  float q0a, q0b, doa, dob;
  {
    const __half2* g2 = reinterpret_cast<const __half2*>(a.dO16 + size_t(pb) * a.ld_do + h * HD);
    const __half2* q2 = reinterpret_cast<const __half2*>(base);
#pragma unroll
    for (int e = 0; e < HD / 2; ++e) {
      const float2 x = __half22float2(g2[e]);
      dO[2 * e] = x.x, dO[2 * e + 1] = x.y;
    }
    const float2 y = __half22float2(q2[lane]), x = __half22float2(g2[lane]);
    q0a = y.x, q0b = y.y, doa = x.x, dob = x.y;
  }
  const __half* Arow = a.probs16 + size_t(b * a.H + h) * T * a.ldp;  // query row 0
  const float r0 = a.r[size_t(pb) * T];
  constexpr int JMAX = (TR + 31) / 32;  // 9 keys per lane
  float G[JMAX], A[JMAX];
  float dsum = 0.f;
#pragma unroll
  for (int c = 0; c < JMAX; ++c) {
    const int j = c * 32 + lane;
    G[c] = 0.f, A[c] = 0.f;
    if (j < T) {
      // every lane reads its own key row: 16-byte loads, so a warp load costs 32 L1 wavefronts per 8 channels, not per 2
      const uint4* v4 = reinterpret_cast<const uint4*>(base + size_t(j) * a.ldq + 2 * d);
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < HD / 8; ++e) {
        const uint4 u = __ldg(v4 + e);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
          acc = fmaf(dO[8 * e + 2 * q], v.x, acc), acc = fmaf(dO[8 * e + 2 * q + 1], v.y, acc);
        }
      }
      G[c] = acc;
      A[c] = __half2float(Arow[j]);
      dsum += acc * A[c];
    }
  }
  dsum = warp_sum(dsum);
  float* wp = a.wpart + size_t(gw) * T;
  const size_t ld = size_t(a.splits) * 3 * d;
  const float invH = 1.0f / a.H;
  float ds[JMAX];
#pragma unroll
  for (int c = 0; c < JMAX; ++c) {
    const int j = c * 32 + lane;
    ds[c] = A[c] * (G[c] - dsum);
    if (j < T) {
      float x = G[c] * A[c];
      if (a.positive_only) x = fmaxf(x, 0.f);
      wp[j] = r0 * x * invH;
    }
  }
  if (a.need_dqkv) {
    // One key row per step, the warp's 32 lanes across the 64 head channels (one half2 each): every store below is a
    // whole 128-byte row segment.  (The first version let each lane write its own rows 4 bytes at a time: 32 partial
    // sectors per warp store, 5.8 ms per launch for 2.4 GB of output.)
    float dq0 = 0.f, dq1 = 0.f;
#pragma unroll
    for (int c = 0; c < JMAX; ++c) {
      const int jn = min(32, T - c * 32);  // warp-uniform
      for (int jj = 0; jj < jn; ++jj) {
        const int j = c * 32 + jj;
        const float dsj = __shfl_sync(0xffffffffu, ds[c], jj), aj = __shfl_sync(0xffffffffu, A[c], jj);
        const float2 k = __half22float2(reinterpret_cast<const __half2*>(base + size_t(j) * a.ldq + d)[lane]);
        dq0 = fmaf(dsj, k.x, dq0), dq1 = fmaf(dsj, k.y, dq1);
        __half* orow = a.dqkv16 + (size_t(pb) * T + j) * ld + h * HD;
        store_h2_split(orow, 0, d + 2 * lane, 3 * d, a.splits, dsj * q0a, dsj * q0b);
        store_h2_split(orow, 0, 2 * d + 2 * lane, 3 * d, a.splits, aj * doa, aj * dob);
        if (j > 0) store_h2_split(orow, 0, 2 * lane, 3 * d, a.splits, 0.f, 0.f);
      }
    }
    // dQ_0 = scale * sum_j dS_0j K_j: lane e already holds elements 2e, 2e+1
    __half* orow = a.dqkv16 + (size_t(pb) * T) * ld + h * HD;
    store_h2_split(orow, 0, 2 * lane, 3 * d, a.splits, dq0 * a.scale, dq1 * a.scale);
  }
}

template <int NCHUNK>
static int launch_attn_fwd(const float* qkv, float* probs, __half* probs16, int ld_p16, float* o32, __half* o16, int B,
                           int T, int H, int d, int causal, int splits, cudaStream_t st) {
  const size_t smem = (size_t(2) * T * KV_STRIDE + size_t(8) * NCHUNK * 32) * sizeof(float);
  SB_REQUIRE(smem <= 227 * 1024, "attention forward: T=%d does not fit in shared memory", T);
  SB_REQUIRE(!probs16 || ld_p16 <= NCHUNK * 32, "attention forward: probs16 pitch %d too wide", ld_p16);
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NCHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  attn_fwd_kernel<NCHUNK><<<B * H, 256, smem, st>>>(qkv, probs, probs16, ld_p16, o32, o16, T, H, d, causal, splits);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_attn_fwd(const float* qkv, float* probs, void* probs16, int32_t ld_p16, float* o32, void* o16,
                               int32_t B, int32_t T, int32_t H, int32_t causal, int32_t splits, void* stream) {
  SB_REQUIRE(qkv && (o32 || o16) && B > 0 && T > 0 && H > 0, "semabs_attn_fwd: bad arguments");
  SB_REQUIRE(!probs16 || (ld_p16 >= T && ld_p16 % 8 == 0), "semabs_attn_fwd: bad probs16 pitch %d", ld_p16);
  const int d = H * HD;
  cudaStream_t st = (cudaStream_t)stream;
  __half* p16 = (__half*)probs16;
  if (T <= TR && (!probs16 || ld_p16 <= TR)) {
    AttnFwdArgs a{qkv, probs, p16, ld_p16, o32, (__half*)o16, T, H, d, causal, splits};
    const size_t smem = size_t(4) * TR * TS * sizeof(__half);
    static bool configured = false;
    if (!configured) {
      SB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    attn_fwd_mma_kernel<<<B * H, FWD_WARPS * 32, smem, st>>>(a);
    SB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int nchunk = (T + 31) / 32;
  if (nchunk <= 2) return launch_attn_fwd<2>(qkv, probs, p16, ld_p16, o32, (__half*)o16, B, T, H, d, causal, splits, st);
  if (nchunk <= 3) return launch_attn_fwd<3>(qkv, probs, p16, ld_p16, o32, (__half*)o16, B, T, H, d, causal, splits, st);
  if (nchunk <= 9) return launch_attn_fwd<9>(qkv, probs, p16, ld_p16, o32, (__half*)o16, B, T, H, d, causal, splits, st);
  SB_REQUIRE(nchunk <= 13, "semabs_attn_fwd: T=%d > 416 tokens is not supported", T);
  return launch_attn_fwd<13>(qkv, probs, p16, ld_p16, o32, (__half*)o16, B, T, H, d, causal, splits, st);
}

extern "C" int semabs_attn_bwd(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                               const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart,
                               void* dqkv16, int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits,
                               int32_t positive_only, int32_t need_dqkv, void* stream) {
  SB_REQUIRE(qkv16 && probs16 && o32 && dO16 && r && delta_ws && wpart, "semabs_attn_bwd: null pointer");
  SB_REQUIRE(!need_dqkv || dqkv16, "semabs_attn_bwd: dqkv16 missing");
  SB_REQUIRE(P > 0 && B > 0 && T > 0 && H > 0 && B * H <= 65535, "semabs_attn_bwd: bad shape");
  SB_REQUIRE(T <= TR, "semabs_attn_bwd: T=%d > %d tokens is not supported", T, TR);
  SB_REQUIRE(ld_p16 % 8 == 0 && ld_p16 >= ((T + 15) / 16) * 16, "semabs_attn_bwd: probs16 pitch %d must be a multiple of 8 and >= %d",
             ld_p16, ((T + 15) / 16) * 16);
  cudaStream_t st = (cudaStream_t)stream;
  SB_REQUIRE(ld_qkv >= 3 * H * HD && ld_qkv % 8 == 0, "semabs_attn_bwd: bad qkv16 pitch %d", ld_qkv);
  AttnBwdArgs a{};
  a.ldq = ld_qkv;
  a.qkv16 = (const __half*)qkv16, a.probs16 = (const __half*)probs16, a.ldp = ld_p16, a.o32 = o32;
  a.dO16 = (const __half*)dO16, a.ld_do = ld_do, a.delta = delta_ws, a.r = r, a.wpart = wpart;
  a.dqkv16 = (__half*)dqkv16, a.P = P, a.B = B, a.T = T, a.H = H, a.d = H * HD, a.splits = splits;
  a.scale = 0.125f, a.positive_only = positive_only, a.need_dqkv = need_dqkv;
  const size_t smem_row = size_t(2) * TR * TS * sizeof(__half);
  const size_t smem_col = smem_row + size_t(2) * TR * sizeof(float) + size_t(BWD_WARPS) * 2 * 64 * STRIP * sizeof(__half);
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_row));
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
    configured = true;
  }
  dim3 grid(P, B * H);
  if (need_dqkv) {
    attn_bwd_row_kernel<<<grid, BWD_WARPS * 32, smem_row, st>>>(a);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  attn_bwd_col_kernel<<<grid, BWD_WARPS * 32, smem_col, st>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_attn_bwd_cls(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const void* dO16_cls,
                                   int32_t ld_do, const float* r, float* wpart, void* dqkv16, int32_t P, int32_t B,
                                   int32_t T, int32_t H, int32_t splits, int32_t positive_only, int32_t need_dqkv,
                                   void* stream) {
  SB_REQUIRE(qkv16 && probs16 && dO16_cls && r && wpart, "semabs_attn_bwd_cls: null pointer");
  SB_REQUIRE(!need_dqkv || dqkv16, "semabs_attn_bwd_cls: dqkv16 missing");
  SB_REQUIRE(P > 0 && B > 0 && T > 0 && T <= TR && H > 0, "semabs_attn_bwd_cls: bad shape");
  SB_REQUIRE(ld_qkv >= 3 * H * HD && ld_qkv % 8 == 0, "semabs_attn_bwd_cls: bad qkv16 pitch %d", ld_qkv);
  AttnClsArgs a{};
  a.ldq = ld_qkv;
  a.qkv16 = (const __half*)qkv16, a.probs16 = (const __half*)probs16, a.ldp = ld_p16, a.dO16 = (const __half*)dO16_cls;
  a.ld_do = ld_do, a.r = r, a.wpart = wpart, a.dqkv16 = (__half*)dqkv16, a.P = P, a.B = B, a.T = T, a.H = H;
  a.d = H * HD, a.splits = splits, a.scale = 0.125f, a.positive_only = positive_only, a.need_dqkv = need_dqkv;
  const int warps = P * B * H;
  attn_bwd_cls_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
