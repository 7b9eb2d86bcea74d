// Evaluation metrics of the training / evaluation loop on the device:
//   utils.prediction_analysis (utils.py:338-380): per (scene, patch) confusion counts over the non-ignored points
//     -> iou, precision, recall, false_negative, false_positive — the reference loops over (b, p) in Python;
//   utils.voxelize_points (utils.py:617-665): predictions / labels / ignore flags scatter-MAXed onto a coarse grid
//     (VirtualGrid(reduce_method="max") -> torch_scatter.scatter(reduce="max"), empty voxels 0), then the same counts.
// The three scattered features only take the values {0,1} (prediction, ignore) or {-1,+1} (label), so the scatter-max
// collapses to four flags per voxel, set with atomicOr:  any prediction, any positive label, occupied, any ignored
// (voxelised label == 0 <=> no point fell into the voxel <=> "missing label", utils.py:645).
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

struct GridSpecM {
  float neg_lc[3];
  float scale[3];
  int shape[3];
};

__device__ __forceinline__ void block_add_counts(const unsigned long long (&c)[7], unsigned long long* out) {
  __shared__ unsigned long long sm[7];
  if (threadIdx.x < 7) sm[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    unsigned long long v = c[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[i], v);
  }
  __syncthreads();
  if (threadIdx.x < 7 && sm[threadIdx.x]) atomicAdd(out + threadIdx.x, sm[threadIdx.x]);
}

// counts[n][7] = tp, predicted positives, label positives, union, fn, fp, kept
__global__ void __launch_bounds__(256)
confusion_points_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ label,
                        const uint8_t* __restrict__ ignore, long long npts, unsigned long long* __restrict__ counts) {
  const int n = blockIdx.y;
  const uint8_t *p = pred + size_t(n) * npts, *l = label + size_t(n) * npts, *g = ignore + size_t(n) * npts;
  unsigned long long c[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (long long)gridDim.x * blockDim.x) {
    if (g[i]) continue;
    const bool pp = p[i] != 0, ll = l[i] != 0;
    c[0] += pp && ll, c[1] += pp, c[2] += ll, c[3] += pp || ll, c[4] += ll && !pp, c[5] += pp && !ll, c[6] += 1;
  }
  block_add_counts(c, counts + size_t(n) * 7);
}

// flags[n][voxel] |= 1 (prediction) | 2 (positive label) | 4 (occupied) | 8 (ignored point)
__global__ void __launch_bounds__(256)
voxel_flags_kernel(const float* __restrict__ xyz, int xyz_div, const uint8_t* __restrict__ pred,
                   const uint8_t* __restrict__ label, const uint8_t* __restrict__ ignore, long long npts, GridSpecM grid,
                   unsigned int* __restrict__ flags) {
  const int n = blockIdx.y;
  const long long V = (long long)grid.shape[0] * grid.shape[1] * grid.shape[2];
  const float* x = xyz + size_t(n / xyz_div) * npts * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (long long)gridDim.x * blockDim.x) {
    long long flat = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      long long v = (long long)((x[i * 3 + ax] + grid.neg_lc[ax]) * grid.scale[ax]);  // VirtualGrid.get_points_grid_idxs
      v = v < 0 ? 0 : (v > grid.shape[ax] - 1 ? grid.shape[ax] - 1 : v);
      flat = flat * grid.shape[ax] + v;
    }
    const size_t k = size_t(n) * npts + i;
    const unsigned int f = (pred[k] ? 1u : 0u) | (label[k] ? 2u : 0u) | 4u | (ignore[k] ? 8u : 0u);
    atomicOr(flags + size_t(n) * V + flat, f);
  }
}

__global__ void __launch_bounds__(256)
confusion_voxels_kernel(const unsigned int* __restrict__ flags, long long V, unsigned long long* __restrict__ counts,
                        uint8_t* __restrict__ vox_pred, uint8_t* __restrict__ vox_label, uint8_t* __restrict__ vox_ignore) {
  const int n = blockIdx.y;
  unsigned long long c[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int f = flags[size_t(n) * V + i];
    const bool pp = f & 1u, ll = f & 2u, ig = (f & 8u) || !(f & 4u);
    if (vox_pred) vox_pred[size_t(n) * V + i] = pp, vox_label[size_t(n) * V + i] = ll, vox_ignore[size_t(n) * V + i] = ig;
    if (ig) continue;
    c[0] += pp && ll, c[1] += pp, c[2] += ll, c[3] += pp || ll, c[4] += ll && !pp, c[5] += pp && !ll, c[6] += 1;
  }
  block_add_counts(c, counts + size_t(n) * 7);
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_confusion_counts(const uint8_t* pred, const uint8_t* label, const uint8_t* ignore, int32_t N,
                                       int64_t npts, uint64_t* counts, void* stream) {
  SB_REQUIRE(pred && label && ignore && counts && N > 0 && npts > 0, "semabs_confusion_counts: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SB_CHECK_CUDA(cudaMemsetAsync(counts, 0, size_t(N) * 7 * sizeof(uint64_t), st));
  long long blocks = (npts + 255) / 256;
  const long long cap = (long long)num_sms() * 8 / N + 1;
  if (blocks > cap) blocks = cap;
  confusion_points_kernel<<<dim3((unsigned)blocks, N), 256, 0, st>>>(pred, label, ignore, npts,
                                                                    reinterpret_cast<unsigned long long*>(counts));
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_voxelized_confusion_counts(const float* xyz, int32_t xyz_div, const uint8_t* pred,
                                                 const uint8_t* label, const uint8_t* ignore, int32_t N, int64_t npts,
                                                 const float* neg_lc, const float* scale, const int32_t* shape,
                                                 uint32_t* flags_ws, uint64_t* counts, uint8_t* vox_pred,
                                                 uint8_t* vox_label, uint8_t* vox_ignore, void* stream) {
  SB_REQUIRE(xyz && pred && label && ignore && neg_lc && scale && shape && flags_ws && counts && N > 0 && npts > 0 &&
                 xyz_div >= 1,
             "semabs_voxelized_confusion_counts: bad arguments");
  SB_REQUIRE((vox_pred != nullptr) == (vox_label != nullptr) && (vox_pred != nullptr) == (vox_ignore != nullptr),
             "semabs_voxelized_confusion_counts: pass all three voxel outputs or none");
  GridSpecM g{};
  for (int i = 0; i < 3; ++i) g.neg_lc[i] = neg_lc[i], g.scale[i] = scale[i], g.shape[i] = shape[i];
  const long long V = (long long)shape[0] * shape[1] * shape[2];
  cudaStream_t st = (cudaStream_t)stream;
  SB_CHECK_CUDA(cudaMemsetAsync(flags_ws, 0, size_t(N) * V * sizeof(uint32_t), st));
  SB_CHECK_CUDA(cudaMemsetAsync(counts, 0, size_t(N) * 7 * sizeof(uint64_t), st));
  long long blocks = (npts + 255) / 256;
  const long long cap = (long long)num_sms() * 8 / N + 1;
  if (blocks > cap) blocks = cap;
  voxel_flags_kernel<<<dim3((unsigned)blocks, N), 256, 0, st>>>(xyz, xyz_div, pred, label, ignore, npts, g, flags_ws);
  SB_CHECK_CUDA(cudaGetLastError());
  long long vb = (V + 255) / 256;
  if (vb > cap) vb = cap;
  confusion_voxels_kernel<<<dim3((unsigned)vb, N), 256, 0, st>>>(flags_ws, V, reinterpret_cast<unsigned long long*>(counts),
                                                                vox_pred, vox_label, vox_ignore);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
