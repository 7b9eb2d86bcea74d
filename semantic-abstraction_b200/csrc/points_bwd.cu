// Backward of the point <-> voxel stages of SemAbs3D / SemAbsVOOL (training step; reference net.py:204-256 implicit
// decoder, :300-309 cosine pointing head, :185-201 scatter-mean voxelisation, :358-367 point MLP under loss.backward()).
//
// Two stages per module:
//   1. a per-item kernel (warp per query / per 4 points, same mapping as the forward kernels in points.cu) recomputes
//      the forward activations, forms the layer deltas, routes the feature gradient (trilinear scatter-add into the
//      volume gradient / gather from it) and writes [inputs | activations | deltas] rows to a scratch matrix;
//   2. weight gradients are outer-product reductions over those rows (`semabs_outer_reduce_f32`, fp32 SIMT register
//      tiles, split over the items, atomics at the end); bias gradients are column sums (semabs_groupnorm_bwd_reduce).
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

__device__ __forceinline__ float leaky_b(float x) { return x > 0.f ? x : 0.01f * x; }
__device__ __forceinline__ float leaky_grad_from_out(float y) { return y > 0.f ? 1.f : 0.01f; }  // sign(y) == sign(x)

struct GridSpecB {
  float neg_lc[3];
  float scale[3];
  int shape[3];
};
__device__ __forceinline__ float grid_coord_b(float p, const GridSpecB& g, int a) { return (p + g.neg_lc[a]) * g.scale[a]; }

// ---------------------------------------------------------------------------------------------------------------
// implicit decoder backward, stage 1 (one warp per query)
// ---------------------------------------------------------------------------------------------------------------
struct DecodeBwdArgs {
  const float* vol0;
  const float* vol1;
  int C0, nvol;
  const float* query;
  int N, nq;
  GridSpecB grid;
  int concat_xyz;
  const float *w1t, *w1, *b1, *w2t, *w2, *b2;  // w1t [Cin][Hs], w1 [Hs][Cin], w2t [Hs][out], w2 [out][Hs]
  int Hs, out_dim;
  const float* emb;
  float temperature;
  const float* dout;   // [N,nq,out_dim] or [N,nq] with emb
  float* dvol0;        // [N,X,Y,Z,C0], zero-initialised by the caller; accumulated with atomics
  float* dvol1;
  float* demb;         // [N,out_dim] (zero-initialised) or null
  float* scratch;      // [N*nq][ld]: in @0, h @off_h, d_o @off_do, d_pre @off_dp
  int ld, off_h, off_do, off_dp;
};

__global__ void __launch_bounds__(256) sample_decode_bwd_kernel(DecodeBwdArgs a) {
  extern __shared__ float sm[];
  const int Cf = a.C0 * a.nvol;
  const int Cin = Cf + (a.concat_xyz ? 3 : 0);
  float* w1t = sm;
  float* w1 = w1t + Cin * a.Hs;
  float* w2t = w1 + Cin * a.Hs;
  float* w2 = w2t + a.Hs * a.out_dim;
  float* bb = w2 + a.Hs * a.out_dim;
  for (int i = threadIdx.x; i < Cin * a.Hs; i += blockDim.x) w1t[i] = a.w1t[i], w1[i] = a.w1[i];
  for (int i = threadIdx.x; i < a.Hs * a.out_dim; i += blockDim.x) w2t[i] = a.w2t[i], w2[i] = a.w2[i];
  for (int i = threadIdx.x; i < a.Hs; i += blockDim.x) bb[i] = a.b1[i];
  for (int i = threadIdx.x; i < a.out_dim; i += blockDim.x) bb[a.Hs + i] = a.b2[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int X = a.grid.shape[0], Y = a.grid.shape[1], Z = a.grid.shape[2];
  const long long total = (long long)a.N * a.nq;
  for (long long qi = (long long)blockIdx.x * 8 + warp; qi < total; qi += (long long)gridDim.x * 8) {
    const int n = int(qi / a.nq);
    const float* qp = a.query + qi * 3;
    float gn[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      float v = grid_coord_b(qp[ax], a.grid, ax);
      v = fminf(fmaxf(v, 0.f), float(a.grid.shape[ax] - 1));
      v = v / float(a.grid.shape[ax]);
      gn[ax] = 2.0f * v - 1.0f;
    }
    float fz = ((gn[0] + 1.f) / 2.f) * float(Z - 1), fy = ((gn[1] + 1.f) / 2.f) * float(Y - 1), fx = ((gn[2] + 1.f) / 2.f) * float(X - 1);
    fz = fminf(fmaxf(fz, 0.f), float(Z - 1)), fy = fminf(fmaxf(fy, 0.f), float(Y - 1)), fx = fminf(fmaxf(fx, 0.f), float(X - 1));
    const float z0f = floorf(fz), y0f = floorf(fy), x0f = floorf(fx);
    const int z0 = int(z0f), y0 = int(y0f), x0 = int(x0f);
    const float tz = fz - z0f, ty = fy - y0f, tx = fx - x0f;
    // ---- forward recompute (identical arithmetic to sample_decode_kernel) ----
    float f[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = lane + 32 * k;
      if (c < Cf) {
        const float* vol = (c < a.C0 ? a.vol0 : a.vol1) + size_t(n) * X * Y * Z * a.C0 + (c < a.C0 ? c : c - a.C0);
        float acc = 0.f;
#pragma unroll
        for (int corner = 0; corner < 8; ++corner) {
          const int dz = corner & 1, dy = (corner >> 1) & 1, dx = corner >> 2;
          const int zz = z0 + dz, yy = y0 + dy, xx = x0 + dx;
          const float w = (dz ? tz : 1.f - tz) * (dy ? ty : 1.f - ty) * (dx ? tx : 1.f - tx);
          if (zz < Z && yy < Y && xx < X) acc += w * vol[((size_t(xx) * Y + yy) * Z + zz) * a.C0];
        }
        f[k] = acc;
      }
    }
    float h[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (lane + 32 * k < a.Hs) h[k] = bb[lane + 32 * k];
    for (int c = 0; c < Cf; ++c) {
      const float fc = __shfl_sync(0xffffffffu, c < 32 ? f[0] : f[1], c & 31);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.Hs) h[k] = fmaf(w1t[c * a.Hs + lane + 32 * k], fc, h[k]);
    }
    if (a.concat_xyz) {
#pragma unroll
      for (int ax = 0; ax < 3; ++ax)
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (lane + 32 * k < a.Hs) h[k] = fmaf(w1t[(Cf + ax) * a.Hs + lane + 32 * k], gn[ax], h[k]);
    }
    h[0] = leaky_b(h[0]), h[1] = leaky_b(h[1]);
    float o[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (lane + 32 * k < a.out_dim) o[k] = bb[a.Hs + lane + 32 * k];
    for (int j = 0; j < a.Hs; ++j) {
      const float hj = __shfl_sync(0xffffffffu, j < 32 ? h[0] : h[1], j & 31);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.out_dim) o[k] = fmaf(w2t[j * a.out_dim + lane + 32 * k], hj, o[k]);
    }
    // ---- output delta ----
    float d_o[2] = {0.f, 0.f};
    if (a.emb) {
      float e[2] = {0.f, 0.f};
      float dot = 0.f, no = 0.f, ne = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.out_dim) {
          e[k] = a.emb[size_t(n) * a.out_dim + lane + 32 * k];
          dot += o[k] * e[k], no += o[k] * o[k], ne += e[k] * e[k];
        }
      dot = warp_sum(dot), no = warp_sum(no), ne = warp_sum(ne);
      const float sno = sqrtf(no), sne = sqrtf(ne);
      const float nof = fmaxf(sno, 1e-8f), nef = fmaxf(sne, 1e-8f);
      const float g = a.dout[qi] / a.temperature;
      const float inv = 1.f / (nof * nef);
      const float ko = sno > 1e-8f ? dot * inv / (nof * nof) : 0.f;  // d/d o of the clamped norm vanishes below eps
      const float ke = sne > 1e-8f ? dot * inv / (nef * nef) : 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.out_dim) {
          d_o[k] = g * (e[k] * inv - ko * o[k]);
          if (a.demb) atomicAdd(a.demb + size_t(n) * a.out_dim + lane + 32 * k, g * (o[k] * inv - ke * e[k]));
        }
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.out_dim) d_o[k] = a.dout[qi * a.out_dim + lane + 32 * k];
    }
    // ---- hidden delta: d_pre[j] = leaky'(h_j) * sum_k W2[k][j] d_o[k] ----
    float d_pre[2] = {0.f, 0.f};
    for (int k = 0; k < a.out_dim; ++k) {
      const float dk = __shfl_sync(0xffffffffu, k < 32 ? d_o[0] : d_o[1], k & 31);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
        if (lane + 32 * jj < a.Hs) d_pre[jj] = fmaf(w2[k * a.Hs + lane + 32 * jj], dk, d_pre[jj]);
    }
    d_pre[0] *= leaky_grad_from_out(h[0]), d_pre[1] *= leaky_grad_from_out(h[1]);
    // ---- feature delta: d_f[c] = sum_j W1[j][c] d_pre[j]; trilinear scatter-add ----
    float d_f[2] = {0.f, 0.f};
    for (int j = 0; j < a.Hs; ++j) {
      const float dj = __shfl_sync(0xffffffffu, j < 32 ? d_pre[0] : d_pre[1], j & 31);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < Cf) d_f[k] = fmaf(w1[j * Cin + lane + 32 * k], dj, d_f[k]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = lane + 32 * k;
      if (c < Cf) {
        float* dv = (c < a.C0 ? a.dvol0 : a.dvol1);
        if (dv) {
          dv += size_t(n) * X * Y * Z * a.C0 + (c < a.C0 ? c : c - a.C0);
#pragma unroll
          for (int corner = 0; corner < 8; ++corner) {
            const int dz = corner & 1, dy = (corner >> 1) & 1, dx = corner >> 2;
            const int zz = z0 + dz, yy = y0 + dy, xx = x0 + dx;
            const float w = (dz ? tz : 1.f - tz) * (dy ? ty : 1.f - ty) * (dx ? tx : 1.f - tx);
            if (zz < Z && yy < Y && xx < X && w != 0.f) atomicAdd(dv + ((size_t(xx) * Y + yy) * Z + zz) * a.C0, w * d_f[k]);
          }
        }
      }
    }
    // ---- scratch row for the weight-gradient reductions ----
    float* row = a.scratch + size_t(qi) * a.ld;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k;
      if (i < Cf) row[i] = f[k];
      if (i < a.Hs) row[a.off_h + i] = h[k], row[a.off_dp + i] = d_pre[k];
      if (i < a.out_dim) row[a.off_do + i] = d_o[k];
    }
    if (a.concat_xyz && lane < 3) row[Cf + lane] = gn[lane];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// point MLP + scatter-mean backward, stage 1 (one warp per 4 points)
// ---------------------------------------------------------------------------------------------------------------
constexpr int PB_PTS = 4;
constexpr int PB_WARPS = 8;

struct PointBwdArgs {
  const float* xyz;
  const float* feat;
  int xyz_div, N, npts, F, hidden, C;
  const float *w1t, *b1, *w2t, *w2, *b2, *w3;  // w1t [3+F][Hd], w2t [Hd][Hd] (in-major), w2 [Hd][Hd] (out-major), w3 [C][Hd]
  GridSpecB grid;
  const float* dvol;   // [N, S, Cpad] gradient of the voxelised volume
  const float* cnt;    // [N, S] points per voxel (from the forward)
  int Cpad;
  float* scratch;      // [N*npts][ld]: in @0 (8), h1 @8, h2 @8+Hd, d3 @8+2Hd (C padded to 4), d2, d1
  int ld, off_d3, off_d2, off_d1;
};

__global__ void __launch_bounds__(PB_WARPS * 32) point_mlp_bwd_kernel(PointBwdArgs a) {
  extern __shared__ float sm[];
  const int in_dim = 3 + a.F, Hd = a.hidden, C = a.C;
  float* w1t = sm;
  float* w2t = w1t + in_dim * Hd;
  float* w2 = w2t + Hd * Hd;
  float* w3 = w2 + Hd * Hd;
  float* bb = w3 + C * Hd;                       // b1[Hd] b2[Hd]
  float* wbuf = bb + 2 * Hd;                     // per warp: h1 [Hd][4], h2 [Hd][4], d2 [Hd][4], d3 [C][4]
  for (int i = threadIdx.x; i < in_dim * Hd; i += blockDim.x) w1t[i] = a.w1t[i];
  for (int i = threadIdx.x; i < Hd * Hd; i += blockDim.x) w2t[i] = a.w2t[i], w2[i] = a.w2[i];
  for (int i = threadIdx.x; i < C * Hd; i += blockDim.x) w3[i] = a.w3[i];
  for (int i = threadIdx.x; i < Hd; i += blockDim.x) bb[i] = a.b1[i], bb[Hd + i] = a.b2[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* h1 = wbuf + size_t(warp) * (3 * Hd + C) * PB_PTS;
  float* h2 = h1 + Hd * PB_PTS;
  float* d2 = h2 + Hd * PB_PTS;
  float* d3 = d2 + Hd * PB_PTS;
  const long long total = (long long)a.N * a.npts;
  const long long groups = (total + PB_PTS - 1) / PB_PTS;
  const long long S = (long long)a.grid.shape[0] * a.grid.shape[1] * a.grid.shape[2];
  const int KH = Hd / 32;
  for (long long gidx = (long long)blockIdx.x * PB_WARPS + warp; gidx < groups; gidx += (long long)gridDim.x * PB_WARPS) {
    const long long p0 = gidx * PB_PTS;
    float in[PB_PTS][8];
    long long vox[PB_PTS];
    int nn[PB_PTS];
#pragma unroll
    for (int q = 0; q < PB_PTS; ++q) {
      const long long pt = p0 + q;
      const bool ok = pt < total;
      const int n = ok ? int(pt / a.npts) : 0;
      const int i = ok ? int(pt % a.npts) : 0;
      const float* xp = a.xyz + (size_t(n / a.xyz_div) * a.npts + i) * 3;
      const float* fp = a.feat + (size_t(n) * a.npts + i) * a.F;
      in[q][0] = xp[0], in[q][1] = xp[1], in[q][2] = xp[2];
#pragma unroll
      for (int f = 0; f < 5; ++f) in[q][3 + f] = f < a.F ? fp[f] : 0.f;
      long long flat = 0;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        long long v = (long long)grid_coord_b(in[q][ax], a.grid, ax);
        v = v < 0 ? 0 : (v > a.grid.shape[ax] - 1 ? a.grid.shape[ax] - 1 : v);
        flat = flat * a.grid.shape[ax] + v;
      }
      vox[q] = ok ? flat : -1;
      nn[q] = n;
    }
    // forward recompute: h1, h2 (same arithmetic as point_mlp_scatter_kernel)
    for (int k = 0; k < KH; ++k) {
      const int o = lane + 32 * k;
      float acc[PB_PTS];
#pragma unroll
      for (int q = 0; q < PB_PTS; ++q) acc[q] = bb[o];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < in_dim) {
          const float w = w1t[i * Hd + o];
#pragma unroll
          for (int q = 0; q < PB_PTS; ++q) acc[q] = fmaf(w, in[q][i], acc[q]);
        }
      }
      *reinterpret_cast<float4*>(h1 + o * PB_PTS) = make_float4(leaky_b(acc[0]), leaky_b(acc[1]), leaky_b(acc[2]), leaky_b(acc[3]));
    }
    __syncwarp();
    for (int k = 0; k < KH; ++k) {
      const int o = lane + 32 * k;
      float acc[PB_PTS];
#pragma unroll
      for (int q = 0; q < PB_PTS; ++q) acc[q] = bb[Hd + o];
      for (int i = 0; i < Hd; ++i) {
        const float w = w2t[i * Hd + o];
        const float4 hh = *reinterpret_cast<const float4*>(h1 + i * PB_PTS);
        acc[0] = fmaf(w, hh.x, acc[0]), acc[1] = fmaf(w, hh.y, acc[1]), acc[2] = fmaf(w, hh.z, acc[2]), acc[3] = fmaf(w, hh.w, acc[3]);
      }
      *reinterpret_cast<float4*>(h2 + o * PB_PTS) = make_float4(leaky_b(acc[0]), leaky_b(acc[1]), leaky_b(acc[2]), leaky_b(acc[3]));
    }
    // d3[q][o] = dvol[n, vox, o] / cnt[n, vox]   (scatter-mean backward; empty voxels never own a point)
    for (int o = lane; o < C; o += 32) {
      float v[PB_PTS];
#pragma unroll
      for (int q = 0; q < PB_PTS; ++q) {
        v[q] = 0.f;
        if (vox[q] >= 0) {
          const size_t cell = size_t(nn[q]) * S + vox[q];
          v[q] = a.dvol[cell * a.Cpad + o] / a.cnt[cell];
        }
      }
      *reinterpret_cast<float4*>(d3 + o * PB_PTS) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncwarp();
    // d2[q][j] = leaky'(h2) * sum_o W3[o][j] d3[q][o]
    for (int k = 0; k < KH; ++k) {
      const int j = lane + 32 * k;
      float acc[PB_PTS] = {0.f, 0.f, 0.f, 0.f};
      for (int o = 0; o < C; ++o) {
        const float w = w3[o * Hd + j];
        const float4 d = *reinterpret_cast<const float4*>(d3 + o * PB_PTS);
        acc[0] = fmaf(w, d.x, acc[0]), acc[1] = fmaf(w, d.y, acc[1]), acc[2] = fmaf(w, d.z, acc[2]), acc[3] = fmaf(w, d.w, acc[3]);
      }
      const float4 hh = *reinterpret_cast<const float4*>(h2 + j * PB_PTS);
      *reinterpret_cast<float4*>(d2 + j * PB_PTS) =
          make_float4(acc[0] * leaky_grad_from_out(hh.x), acc[1] * leaky_grad_from_out(hh.y),
                      acc[2] * leaky_grad_from_out(hh.z), acc[3] * leaky_grad_from_out(hh.w));
    }
    __syncwarp();
    // d1[q][i] = leaky'(h1) * sum_j W2[j][i] d2[q][j]; rows out
    for (int k = 0; k < KH; ++k) {
      const int i = lane + 32 * k;
      float acc[PB_PTS] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < Hd; ++j) {
        const float w = w2[j * Hd + i];
        const float4 d = *reinterpret_cast<const float4*>(d2 + j * PB_PTS);
        acc[0] = fmaf(w, d.x, acc[0]), acc[1] = fmaf(w, d.y, acc[1]), acc[2] = fmaf(w, d.z, acc[2]), acc[3] = fmaf(w, d.w, acc[3]);
      }
      const float4 hh1 = *reinterpret_cast<const float4*>(h1 + i * PB_PTS);
      const float4 hh2 = *reinterpret_cast<const float4*>(h2 + i * PB_PTS);
      const float4 dd2 = *reinterpret_cast<const float4*>(d2 + i * PB_PTS);
      const float d1v[PB_PTS] = {acc[0] * leaky_grad_from_out(hh1.x), acc[1] * leaky_grad_from_out(hh1.y),
                                 acc[2] * leaky_grad_from_out(hh1.z), acc[3] * leaky_grad_from_out(hh1.w)};
      const float h1v[PB_PTS] = {hh1.x, hh1.y, hh1.z, hh1.w}, h2v[PB_PTS] = {hh2.x, hh2.y, hh2.z, hh2.w};
      const float d2v[PB_PTS] = {dd2.x, dd2.y, dd2.z, dd2.w};
#pragma unroll
      for (int q = 0; q < PB_PTS; ++q)
        if (vox[q] >= 0) {
          float* row = a.scratch + size_t(p0 + q) * a.ld;
          row[8 + i] = h1v[q], row[8 + Hd + i] = h2v[q], row[a.off_d2 + i] = d2v[q], row[a.off_d1 + i] = d1v[q];
        }
    }
    for (int o = lane; o < C; o += 32) {
      const float4 d = *reinterpret_cast<const float4*>(d3 + o * PB_PTS);
      const float dv[PB_PTS] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int q = 0; q < PB_PTS; ++q)
        if (vox[q] >= 0) a.scratch[size_t(p0 + q) * a.ld + a.off_d3 + o] = dv[q];
    }
    if (lane < 8) {
#pragma unroll
      for (int q = 0; q < PB_PTS; ++q)
        if (vox[q] >= 0) {
          float v = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == lane) v = in[q][i];
          a.scratch[size_t(p0 + q) * a.ld + lane] = v;
        }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// out[r][c] += scale * sum_p A[p][r] * B[p][c]     (A [P][lda], B [P][ldb] fp32 row-major; 64x64 output tile per CTA,
// 4x4 per thread, the items split over gridDim.y; fp32 atomics join the splits)
// ---------------------------------------------------------------------------------------------------------------
constexpr int OR_T = 64, OR_K = 32;

__global__ void __launch_bounds__(256)
outer_reduce_kernel(const float* __restrict__ A, int lda, int R, const float* __restrict__ B, int ldb, int Cc, long long P,
                    float scale, float* __restrict__ out, int ld_out) {
  __shared__ __align__(16) float As[OR_K][OR_T], Bs[OR_K][OR_T];
  const int tiles_c = (Cc + OR_T - 1) / OR_T;
  const int r0 = (blockIdx.x / tiles_c) * OR_T, c0 = (blockIdx.x % tiles_c) * OR_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long chunks = (P + OR_K - 1) / OR_K;
  for (long long ch = blockIdx.y; ch < chunks; ch += gridDim.y) {
    const long long p0 = ch * OR_K;
    // 32 rows x 64 columns per operand: 2048 elements each, 8 per thread
    for (int i = threadIdx.x; i < OR_K * OR_T; i += 256) {
      const int pr = i / OR_T, cc = i % OR_T;
      const long long p = p0 + pr;
      As[pr][cc] = (p < P && r0 + cc < R) ? A[size_t(p) * lda + r0 + cc] : 0.f;
      Bs[pr][cc] = (p < P && c0 + cc < Cc) ? B[size_t(p) * ldb + c0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < OR_K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bq[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bq[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (r < R && c < Cc) atomicAdd(out + size_t(r) * ld_out + c, acc[i][j] * scale);
    }
}

static void fill_grid_b(GridSpecB& g, const float* neg_lc, const float* scale, const int32_t* shape) {
  for (int i = 0; i < 3; ++i) g.neg_lc[i] = neg_lc[i], g.scale[i] = scale[i], g.shape[i] = shape[i];
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_outer_reduce_f32(const float* A, int32_t lda, int32_t R, const float* B, int32_t ldb, int32_t Cc,
                                       int64_t P, float scale, float* out, int32_t ld_out, void* stream) {
  SB_REQUIRE(A && B && out && R > 0 && Cc > 0 && P > 0 && lda >= R && ldb >= Cc && ld_out >= Cc,
             "semabs_outer_reduce_f32: bad arguments");
  const int tiles = ((R + OR_T - 1) / OR_T) * ((Cc + OR_T - 1) / OR_T);
  const long long chunks = (P + OR_K - 1) / OR_K;
  long long splits = (long long)num_sms() * 4 / tiles;
  if (splits < 1) splits = 1;
  if (splits > chunks) splits = chunks;
  outer_reduce_kernel<<<dim3(tiles, (unsigned)splits), 256, 0, (cudaStream_t)stream>>>(A, lda, R, B, ldb, Cc, P, scale, out,
                                                                                      ld_out);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_sample_decode_bwd(const float* vol0, const float* vol1, int32_t C0, const float* query, int32_t N,
                                        int32_t nq, const float* neg_lc, const float* scale, const int32_t* shape,
                                        int32_t concat_xyz, const float* w1t, const float* w1, const float* b1,
                                        const float* w2t, const float* w2, const float* b2, int32_t Hs, int32_t out_dim,
                                        const float* emb, float temperature, const float* dout, float* dvol0,
                                        float* dvol1, float* demb, float* scratch, int32_t ld, int32_t off_h,
                                        int32_t off_do, int32_t off_dp, void* stream) {
  SB_REQUIRE(vol0 && query && w1t && w1 && b1 && w2t && w2 && b2 && dout && scratch && neg_lc && scale && shape,
             "semabs_sample_decode_bwd: null pointer");
  const int nvol = vol1 ? 2 : 1;
  SB_REQUIRE(C0 * nvol <= 64 && Hs <= 64 && out_dim <= 64 && Hs >= 1 && out_dim >= 1,
             "semabs_sample_decode_bwd: sizes above 64 are not supported");
  const int Cin = C0 * nvol + (concat_xyz ? 3 : 0);
  SB_REQUIRE(off_h >= Cin && off_do >= off_h + Hs && off_dp >= off_do + out_dim && ld >= off_dp + Hs,
             "semabs_sample_decode_bwd: scratch row layout too small");
  DecodeBwdArgs a{};
  a.vol0 = vol0, a.vol1 = vol1, a.C0 = C0, a.nvol = nvol, a.query = query, a.N = N, a.nq = nq;
  fill_grid_b(a.grid, neg_lc, scale, shape);
  a.concat_xyz = concat_xyz, a.w1t = w1t, a.w1 = w1, a.b1 = b1, a.w2t = w2t, a.w2 = w2, a.b2 = b2, a.Hs = Hs, a.out_dim = out_dim;
  a.emb = emb, a.temperature = temperature, a.dout = dout, a.dvol0 = dvol0, a.dvol1 = dvol1, a.demb = demb;
  a.scratch = scratch, a.ld = ld, a.off_h = off_h, a.off_do = off_do, a.off_dp = off_dp;
  const size_t smem = (size_t(2) * Cin * Hs + size_t(2) * Hs * out_dim + Hs + out_dim) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(sample_decode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  const long long total = (long long)N * nq;
  long long blocks = (total + 7) / 8;
  if (blocks > (long long)num_sms() * 4) blocks = (long long)num_sms() * 4;
  sample_decode_bwd_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_points_to_voxels_bwd(const float* xyz, int32_t xyz_div, const float* feat, int32_t N, int32_t npts,
                                           int32_t F, int32_t hidden, int32_t C, const float* w1t, const float* b1,
                                           const float* w2t, const float* w2, const float* b2, const float* w3,
                                           const float* neg_lc, const float* scale, const int32_t* shape,
                                           const float* dvol, const float* cnt, int32_t Cpad, float* scratch, int32_t ld,
                                           int32_t off_d3, int32_t off_d2, int32_t off_d1, void* stream) {
  SB_REQUIRE(xyz && feat && w1t && b1 && w2t && w2 && b2 && w3 && dvol && cnt && scratch && neg_lc && scale && shape,
             "semabs_points_to_voxels_bwd: null pointer");
  SB_REQUIRE(N > 0 && npts > 0 && F >= 1 && F <= 5 && xyz_div >= 1, "semabs_points_to_voxels_bwd: bad shape");
  SB_REQUIRE(hidden % 32 == 0 && hidden >= 32 && hidden <= 128 && C >= 1 && C <= 64 && Cpad >= C,
             "semabs_points_to_voxels_bwd: unsupported MLP size (hidden=%d, C=%d)", hidden, C);
  SB_REQUIRE(off_d3 >= 8 + 2 * hidden && off_d2 >= off_d3 + C && off_d1 >= off_d2 + hidden && ld >= off_d1 + hidden,
             "semabs_points_to_voxels_bwd: scratch row layout too small");
  PointBwdArgs a{};
  a.xyz = xyz, a.feat = feat, a.xyz_div = xyz_div, a.N = N, a.npts = npts, a.F = F, a.hidden = hidden, a.C = C;
  a.w1t = w1t, a.b1 = b1, a.w2t = w2t, a.w2 = w2, a.b2 = b2, a.w3 = w3;
  fill_grid_b(a.grid, neg_lc, scale, shape);
  a.dvol = dvol, a.cnt = cnt, a.Cpad = Cpad, a.scratch = scratch, a.ld = ld, a.off_d3 = off_d3, a.off_d2 = off_d2, a.off_d1 = off_d1;
  const size_t smem = (size_t(3 + F) * hidden + size_t(2) * hidden * hidden + size_t(C) * hidden + 2 * hidden +
                       size_t(PB_WARPS) * (3 * hidden + C) * PB_PTS) * sizeof(float);
  SB_REQUIRE(smem <= 227 * 1024, "semabs_points_to_voxels_bwd: MLP does not fit in shared memory");
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(point_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const long long groups_of_pts = ((long long)N * npts + PB_PTS - 1) / PB_PTS;
  long long blocks = (groups_of_pts + PB_WARPS - 1) / PB_WARPS;
  if (blocks > num_sms()) blocks = num_sms();
  point_mlp_bwd_kernel<<<(unsigned)blocks, PB_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
