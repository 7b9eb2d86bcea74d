// Point <-> voxel stages of SemAbs3D / SemAbsVOOL (reference net.py): the per-point feature MLP fused with the
// scatter-MEAN voxelisation (pts_feat_extractor net.py:358-367,395-404 + VirtualGrid.scatter_points :185-201), and the
// implicit decoder = trilinear gather at query points fused with its 2-layer MLP and, for VOOL, the cosine
// similarity pointing head (ImplicitVolumetricDecoder.forward :215-256, PointingAttention.cosine_sim :300-309).
// HBM-bound random access: volumes are channels-last so one voxel's channels are one 64-128 B line; a warp
// handles a point / query with lanes over channels, weights live in shared memory (transposed, conflict-free).
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }  // nn.LeakyReLU() default slope

struct GridSpec {
  float neg_lc[3];   // -lower_corner
  float scale[3];    // (shape - 1) / (uc - lc), computed in fp32 on the host exactly like net.py:95-98
  int shape[3];      // (X, Y, Z)
};

// VirtualGrid.get_points_grid_idxs (net.py:84-113)
__device__ __forceinline__ float grid_coord(float p, const GridSpec& g, int a) { return (p + g.neg_lc[a]) * g.scale[a]; }

// ---------------------------------------------------------------------------------------------------------
// point MLP + scatter (sum and count); 4 points per warp pass so every weight read feeds 4 FMAs
// ---------------------------------------------------------------------------------------------------------
constexpr int PTS_PER_WARP = 4;
constexpr int MLP_WARPS = 8;

struct PointMlpArgs {
  const float* xyz;     // [n_xyz_batches, npts, 3]
  const float* feat;    // [N, npts, F]
  int xyz_div;          // sample n reads xyz batch n / xyz_div (SemAbs3D repeats xyz over patches, net.py:386-390)
  int N, npts, F;
  int use_mlp, hidden, C;             // C = output channels (= F when use_mlp == 0)
  const float *w1t, *b1, *w2t, *b2, *w3t, *b3;  // transposed weights: w1t [3+F][hidden], w2t [hidden][hidden], w3t [hidden][C]
  GridSpec grid;
  float* vol_sum;       // [N, X*Y*Z, Cpad] (zero-initialised)
  float* vol_cnt;       // [N, X*Y*Z]      (zero-initialised)
  int Cpad;
};

__global__ void __launch_bounds__(MLP_WARPS * 32) point_mlp_scatter_kernel(PointMlpArgs a) {
  extern __shared__ float sm[];
  const int in_dim = 3 + a.F, Hd = a.hidden, C = a.C;
  float* w1t = sm;                       // [in_dim][Hd]
  float* w2t = w1t + in_dim * Hd;        // [Hd][Hd]
  float* w3t = w2t + Hd * Hd;            // [Hd][C]
  float* bb = w3t + Hd * C;              // b1[Hd] b2[Hd] b3[C]
  float* hbuf = bb + 2 * Hd + C;         // per warp: 2 x [Hd][4]
  if (a.use_mlp) {
    for (int i = threadIdx.x; i < in_dim * Hd; i += blockDim.x) w1t[i] = a.w1t[i];
    for (int i = threadIdx.x; i < Hd * Hd; i += blockDim.x) w2t[i] = a.w2t[i];
    for (int i = threadIdx.x; i < Hd * C; i += blockDim.x) w3t[i] = a.w3t[i];
    for (int i = threadIdx.x; i < Hd; i += blockDim.x) bb[i] = a.b1[i], bb[Hd + i] = a.b2[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) bb[2 * Hd + i] = a.b3[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* h1 = hbuf + warp * 2 * Hd * PTS_PER_WARP;
  float* h2 = h1 + Hd * PTS_PER_WARP;
  const long long total = (long long)a.N * a.npts;
  const long long groups = (total + PTS_PER_WARP - 1) / PTS_PER_WARP;
  const int KH = Hd / 32;  // hidden units per lane (<= 8)
  for (long long gidx = (long long)blockIdx.x * MLP_WARPS + warp; gidx < groups; gidx += (long long)gridDim.x * MLP_WARPS) {
    const long long p0 = gidx * PTS_PER_WARP;
    float in[PTS_PER_WARP][8];
    long long vox[PTS_PER_WARP];
    int nn[PTS_PER_WARP];
#pragma unroll
    for (int q = 0; q < PTS_PER_WARP; ++q) {
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
        long long v = (long long)grid_coord(in[q][ax], a.grid, ax);  // .to(int64): truncation toward zero
        v = v < 0 ? 0 : (v > a.grid.shape[ax] - 1 ? a.grid.shape[ax] - 1 : v);
        flat = flat * a.grid.shape[ax] + v;
      }
      vox[q] = ok ? flat : -1;
      nn[q] = n;
    }
    const long long S = (long long)a.grid.shape[0] * a.grid.shape[1] * a.grid.shape[2];
    if (!a.use_mlp) {
#pragma unroll
      for (int q = 0; q < PTS_PER_WARP; ++q)
        if (vox[q] >= 0) {
#pragma unroll
          for (int f = 0; f < 5; ++f)
            if (lane == f && f < a.F) atomicAdd(a.vol_sum + (size_t(nn[q]) * S + vox[q]) * a.Cpad + f, in[q][3 + f]);
          if (lane == 0) atomicAdd(a.vol_cnt + size_t(nn[q]) * S + vox[q], 1.0f);
        }
      continue;
    }
    // layer 1: lane owns hidden units lane + 32k
    for (int k = 0; k < KH; ++k) {
      const int o = lane + 32 * k;
      float acc[PTS_PER_WARP];
#pragma unroll
      for (int q = 0; q < PTS_PER_WARP; ++q) acc[q] = bb[o];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < in_dim) {
          const float w = w1t[i * Hd + o];
#pragma unroll
          for (int q = 0; q < PTS_PER_WARP; ++q) acc[q] = fmaf(w, in[q][i], acc[q]);
        }
      }
      *reinterpret_cast<float4*>(h1 + o * PTS_PER_WARP) = make_float4(leaky(acc[0]), leaky(acc[1]), leaky(acc[2]), leaky(acc[3]));
    }
    __syncwarp();
    // layer 2
    for (int k = 0; k < KH; ++k) {
      const int o = lane + 32 * k;
      float acc[PTS_PER_WARP];
#pragma unroll
      for (int q = 0; q < PTS_PER_WARP; ++q) acc[q] = bb[Hd + o];
      for (int i = 0; i < Hd; ++i) {
        const float w = w2t[i * Hd + o];
        const float4 h = *reinterpret_cast<const float4*>(h1 + i * PTS_PER_WARP);
        acc[0] = fmaf(w, h.x, acc[0]), acc[1] = fmaf(w, h.y, acc[1]), acc[2] = fmaf(w, h.z, acc[2]), acc[3] = fmaf(w, h.w, acc[3]);
      }
      *reinterpret_cast<float4*>(h2 + o * PTS_PER_WARP) = make_float4(leaky(acc[0]), leaky(acc[1]), leaky(acc[2]), leaky(acc[3]));
    }
    __syncwarp();
    // layer 3 + scatter
    for (int o = lane; o < C; o += 32) {
      float acc[PTS_PER_WARP];
#pragma unroll
      for (int q = 0; q < PTS_PER_WARP; ++q) acc[q] = bb[2 * Hd + o];
      for (int i = 0; i < Hd; ++i) {
        const float w = w3t[i * C + o];
        const float4 h = *reinterpret_cast<const float4*>(h2 + i * PTS_PER_WARP);
        acc[0] = fmaf(w, h.x, acc[0]), acc[1] = fmaf(w, h.y, acc[1]), acc[2] = fmaf(w, h.z, acc[2]), acc[3] = fmaf(w, h.w, acc[3]);
      }
#pragma unroll
      for (int q = 0; q < PTS_PER_WARP; ++q)
        if (vox[q] >= 0) atomicAdd(a.vol_sum + (size_t(nn[q]) * S + vox[q]) * a.Cpad + o, acc[q]);
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < PTS_PER_WARP; ++q)
        if (vox[q] >= 0) atomicAdd(a.vol_cnt + size_t(nn[q]) * S + vox[q], 1.0f);
    }
    __syncwarp();
  }
}

// vol = sum / count (0 where empty) in place, + statistics for the UNet's first GroupNorm
__global__ void __launch_bounds__(256) scatter_finalize_kernel(float* __restrict__ vol, const float* __restrict__ cnt,
                                                               long long S, int Cpad, int C, int groups,
                                                               double* __restrict__ stats) {
  __shared__ float smst[16];
  if (threadIdx.x < 16) smst[threadIdx.x] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int qpc = Cpad / 4, vpb = 256 / qpc;
  const int cq = threadIdx.x % qpc, vl = threadIdx.x / qpc;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (long long v = (long long)blockIdx.x * vpb + vl; v < S; v += (long long)gridDim.x * vpb) {
    const float c = cnt[size_t(n) * S + v];
    if (c > 0.f) {
      float4* p = reinterpret_cast<float4*>(vol + (size_t(n) * S + v) * Cpad + 4 * cq);
      float4 a = *p;
      a.x /= c, a.y /= c, a.z /= c, a.w /= c;
      *p = a;
      s0 += a.x + a.y, q0 += a.x * a.x + a.y * a.y, s1 += a.z + a.w, q1 += a.z * a.z + a.w * a.w;
    }
  }
  if (stats) {
    const int cpg = groups == 1 ? Cpad : C / groups;
    // padded channels (>= C) hold zeros: clamp their group index instead of indexing past smst (C < Cpad, groups > 1)
    const int g0 = min((4 * cq) / cpg, groups - 1), g1 = min((4 * cq + 2) / cpg, groups - 1);
    atomicAdd(&smst[2 * g0], s0), atomicAdd(&smst[2 * g0 + 1], q0);
    atomicAdd(&smst[2 * g1], s1), atomicAdd(&smst[2 * g1 + 1], q1);
    __syncthreads();
    if (threadIdx.x < 2 * groups) atomicAdd(stats + size_t(n) * groups * 2 + threadIdx.x, double(smst[threadIdx.x]));
  }
}

// ---------------------------------------------------------------------------------------------------------
// implicit decoder: trilinear gather (grid_sample bilinear / border / align_corners=True with the reference's
// (x,y,z)->(W,H,D) argument order) + Linear(Cin,Hs) LeakyReLU Linear(Hs,out) [+ cosine similarity head]
// one warp per query
// ---------------------------------------------------------------------------------------------------------
struct DecodeArgs {
  const float* vol0;     // [N, X,Y,Z, C0] channels-last
  const float* vol1;     // optional second volume (VOOL concatenates target | reference features, net.py:556)
  int C0;                // channels per volume
  int nvol;
  const float* query;    // [N, nq, 3]
  int N, nq;
  GridSpec grid;
  int concat_xyz;
  const float *w1t, *b1, *w2t, *b2;  // w1t [Cin][Hs], w2t [Hs][out]
  int Hs, out_dim;
  const float* emb;      // optional [N, out_dim]: output = cos_sim(mlp_out, emb[n]) / temperature
  float temperature;
  float* out;            // [N, nq, out_dim] or [N, nq] with emb
};

__global__ void __launch_bounds__(256) sample_decode_kernel(DecodeArgs a) {
  extern __shared__ float sm[];
  const int Cf = a.C0 * a.nvol;
  const int Cin = Cf + (a.concat_xyz ? 3 : 0);
  float* w1t = sm;                     // [Cin][Hs]
  float* w2t = w1t + Cin * a.Hs;       // [Hs][out]
  float* bb = w2t + a.Hs * a.out_dim;  // b1[Hs], b2[out]
  for (int i = threadIdx.x; i < Cin * a.Hs; i += blockDim.x) w1t[i] = a.w1t[i];
  for (int i = threadIdx.x; i < a.Hs * a.out_dim; i += blockDim.x) w2t[i] = a.w2t[i];
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
      float v = grid_coord(qp[ax], a.grid, ax);
      v = fminf(fmaxf(v, 0.f), float(a.grid.shape[ax] - 1));   // clamp (net.py:107-110)
      v = v / float(a.grid.shape[ax]);                         // divided by shape, not shape-1 (net.py:221-222)
      gn[ax] = 2.0f * v - 1.0f;
    }
    // grid_sample: grid[...,0] -> last volume axis (Z), [...,1] -> Y, [...,2] -> X; align_corners=True, border
    float fz = ((gn[0] + 1.f) / 2.f) * float(Z - 1), fy = ((gn[1] + 1.f) / 2.f) * float(Y - 1), fx = ((gn[2] + 1.f) / 2.f) * float(X - 1);
    fz = fminf(fmaxf(fz, 0.f), float(Z - 1)), fy = fminf(fmaxf(fy, 0.f), float(Y - 1)), fx = fminf(fmaxf(fx, 0.f), float(X - 1));
    const float z0f = floorf(fz), y0f = floorf(fy), x0f = floorf(fx);
    const int z0 = int(z0f), y0 = int(y0f), x0 = int(x0f);
    const float tz = fz - z0f, ty = fy - y0f, tx = fx - x0f;
    // features: lane c owns channels c, c+32 (Cf <= 64)
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
    // hidden layer: lane owns units lane, lane+32 (Hs <= 64)
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
    h[0] = leaky(h[0]), h[1] = leaky(h[1]);
    // output layer: lane owns outputs lane, lane+32 (out_dim <= 64)
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
    if (a.emb) {
      // torch.cosine_similarity(key, query, dim=-1) / temperature (net.py:300-309), eps = 1e-8
      float dot = 0.f, no = 0.f, ne = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.out_dim) {
          const float e = a.emb[size_t(n) * a.out_dim + lane + 32 * k];
          dot += o[k] * e, no += o[k] * o[k], ne += e * e;
        }
      dot = warp_sum(dot), no = warp_sum(no), ne = warp_sum(ne);
      if (lane == 0) a.out[qi] = dot / (fmaxf(sqrtf(no), 1e-8f) * fmaxf(sqrtf(ne), 1e-8f)) / a.temperature;
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < a.out_dim) a.out[qi * a.out_dim + lane + 32 * k] = o[k];
    }
  }
}

static void fill_grid(GridSpec& g, const float* neg_lc, const float* scale, const int32_t* shape) {
  for (int i = 0; i < 3; ++i) g.neg_lc[i] = neg_lc[i], g.scale[i] = scale[i], g.shape[i] = shape[i];
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_points_to_voxels(const float* xyz, int32_t xyz_div, const float* feat, int32_t N, int32_t npts,
                                       int32_t F, int32_t use_mlp, int32_t hidden, int32_t C, const float* w1t,
                                       const float* b1, const float* w2t, const float* b2, const float* w3t,
                                       const float* b3, const float* neg_lc, const float* scale, const int32_t* shape,
                                       float* vol, float* cnt, int32_t Cpad, int32_t groups, double* stats,
                                       void* stream) {
  SB_REQUIRE(xyz && feat && neg_lc && scale && shape && vol && cnt, "semabs_points_to_voxels: null pointer");
  SB_REQUIRE(N > 0 && npts > 0 && F >= 1 && F <= 5 && xyz_div >= 1, "semabs_points_to_voxels: bad shape");
  SB_REQUIRE(Cpad % 4 == 0 && Cpad >= C && (256 % (Cpad / 4)) == 0, "semabs_points_to_voxels: bad Cpad %d", Cpad);
  cudaStream_t st = (cudaStream_t)stream;
  PointMlpArgs a{};
  a.xyz = xyz, a.feat = feat, a.xyz_div = xyz_div, a.N = N, a.npts = npts, a.F = F, a.use_mlp = use_mlp;
  a.hidden = use_mlp ? hidden : 32, a.C = use_mlp ? C : F;
  a.w1t = w1t, a.b1 = b1, a.w2t = w2t, a.b2 = b2, a.w3t = w3t, a.b3 = b3;
  fill_grid(a.grid, neg_lc, scale, shape);
  a.vol_sum = vol, a.vol_cnt = cnt, a.Cpad = Cpad;
  size_t smem = 0;
  if (use_mlp) {
    SB_REQUIRE(w1t && b1 && w2t && b2 && w3t && b3, "semabs_points_to_voxels: MLP weights missing");
    SB_REQUIRE(hidden % 32 == 0 && hidden >= 32 && hidden <= 256 && C >= 1 && C <= 256, "semabs_points_to_voxels: unsupported MLP size");
    smem = (size_t(3 + F) * hidden + size_t(hidden) * hidden + size_t(hidden) * C + 2 * hidden + C +
            size_t(MLP_WARPS) * 2 * hidden * PTS_PER_WARP) * sizeof(float);
    SB_REQUIRE(smem <= 227 * 1024, "semabs_points_to_voxels: MLP does not fit in shared memory");
  } else {
    SB_REQUIRE(F <= 32, "semabs_points_to_voxels: too many raw features");
  }
  static bool configured = false;
  if (!configured) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(point_mlp_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const long long S = (long long)shape[0] * shape[1] * shape[2];
  SB_CHECK_CUDA(cudaMemsetAsync(vol, 0, size_t(N) * S * Cpad * sizeof(float), st));
  SB_CHECK_CUDA(cudaMemsetAsync(cnt, 0, size_t(N) * S * sizeof(float), st));
  const long long groups_of_pts = ((long long)N * npts + PTS_PER_WARP - 1) / PTS_PER_WARP;
  long long blocks = (groups_of_pts + MLP_WARPS - 1) / MLP_WARPS;
  if (blocks > num_sms()) blocks = num_sms();
  point_mlp_scatter_kernel<<<(unsigned)blocks, MLP_WARPS * 32, smem, st>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  const int vpb = 256 / (Cpad / 4);
  long long fb = (S + vpb - 1) / vpb;
  if (fb > (long long)num_sms() * 8) fb = (long long)num_sms() * 8;
  scatter_finalize_kernel<<<dim3((unsigned)fb, N), 256, 0, st>>>(vol, cnt, S, Cpad, a.C, groups, stats);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_sample_decode(const float* vol0, const float* vol1, int32_t C0, const float* query, int32_t N,
                                    int32_t nq, const float* neg_lc, const float* scale, const int32_t* shape,
                                    int32_t concat_xyz, const float* w1t, const float* b1, const float* w2t,
                                    const float* b2, int32_t Hs, int32_t out_dim, const float* emb, float temperature,
                                    float* out, void* stream) {
  SB_REQUIRE(vol0 && query && w1t && b1 && w2t && b2 && out && neg_lc && scale && shape, "semabs_sample_decode: null pointer");
  const int nvol = vol1 ? 2 : 1;
  SB_REQUIRE(C0 * nvol <= 64 && Hs <= 64 && out_dim <= 64 && Hs >= 1 && out_dim >= 1,
             "semabs_sample_decode: sizes above 64 are not supported (C=%d x%d, hidden=%d, out=%d)", C0, nvol, Hs, out_dim);
  DecodeArgs a{};
  a.vol0 = vol0, a.vol1 = vol1, a.C0 = C0, a.nvol = nvol, a.query = query, a.N = N, a.nq = nq;
  fill_grid(a.grid, neg_lc, scale, shape);
  a.concat_xyz = concat_xyz, a.w1t = w1t, a.b1 = b1, a.w2t = w2t, a.b2 = b2, a.Hs = Hs, a.out_dim = out_dim;
  a.emb = emb, a.temperature = temperature, a.out = out;
  const int Cin = C0 * nvol + (concat_xyz ? 3 : 0);
  const size_t smem = (size_t(Cin) * Hs + size_t(Hs) * out_dim + Hs + out_dim) * sizeof(float);
  const long long total = (long long)N * nq;
  long long blocks = (total + 7) / 8;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  sample_decode_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(a);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
