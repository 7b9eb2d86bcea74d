// Relevancy store, device half.  Writer side = generate_relevancy.generate_saliency_helper
// (reference generate_relevancy.py:95-111): the [P, H, W] maps are resized to the storage resolution with
// F.interpolate(mode="nearest-exact") and the mean map over the labels is appended as row P.  Reader side =
// SceneUnderstandDataset.load_patches (reference dataset.py:817-872): the selected rows minus the stored mean map,
// bilinearly (align_corners=False) up-sampled back to the image size (the x50 gain of dataset.py:1049-1054 is the `gain`
// argument).  Both are pure bandwidth passes: one thread per output pixel, coalesced along x.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

// ATen nearest_exact_idx: min(floorf((dst + 0.5) * scale), in - 1), scale = float(in) / out (UpSample.h)
__device__ __forceinline__ int nearest_exact_src(int dst, float scale, int in_size) {
  const int s = int(floorf((float(dst) + 0.5f) * scale));
  return s < in_size - 1 ? s : in_size - 1;
}

// maps [P, H, W] -> out [P + 1, SH, SW]; thread = one (y, x) of the storage grid, loops over the labels (mean in a register)
__global__ void __launch_bounds__(256) store_pack_kernel(const float* __restrict__ maps, int P, int H, int W, int SH, int SW,
                                                         float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= SH * SW) return;
  const int x = idx % SW, y = idx / SW;
  const int sy = nearest_exact_src(y, float(H) / float(SH), H), sx = nearest_exact_src(x, float(W) / float(SW), W);
  const float* src = maps + size_t(sy) * W + sx;
  float sum = 0.f;
  for (int p = 0; p < P; ++p) {
    const float v = src[size_t(p) * H * W];
    out[size_t(p) * SH * SW + idx] = v;
    sum += v;  // same left-to-right order as torch's mean(dim=0) over a [P, ...] tensor reduces one output element
  }
  out[size_t(P) * SH * SW + idx] = sum / float(P);
}

// stored [N, SH, SW]; out[k] = gain * bilinear(stored[rows[k]] - stored[mean_row]) at [H, W]
__global__ void __launch_bounds__(256) store_unpack_kernel(const float* __restrict__ stored, const int* __restrict__ rows, int K,
                                                           int mean_row, int SH, int SW, int H, int W, float gain,
                                                           float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)K * H * W) return;
  const int x = int(idx % W), y = int((idx / W) % H), k = int(idx / ((long long)W * H));
  // ATen area_pixel_compute_source_index(align_corners = false)
  float fy = (float(SH) / float(H)) * (float(y) + 0.5f) - 0.5f, fx = (float(SW) / float(W)) * (float(x) + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy, fx = fx < 0.f ? 0.f : fx;
  const int y0 = int(fy), x0 = int(fx);
  const int y1 = y0 + (y0 < SH - 1 ? 1 : 0), x1 = x0 + (x0 < SW - 1 ? 1 : 0);
  const float wy1 = fy - float(y0), wx1 = fx - float(x0), wy0 = 1.f - wy1, wx0 = 1.f - wx1;
  const float* m = stored + size_t(rows[k]) * SH * SW;
  float v00 = m[y0 * SW + x0], v01 = m[y0 * SW + x1], v10 = m[y1 * SW + x0], v11 = m[y1 * SW + x1];
  if (mean_row >= 0) {  // the reference subtracts the mean map at storage resolution, before the up-sampling
    const float* mm = stored + size_t(mean_row) * SH * SW;
    v00 -= mm[y0 * SW + x0], v01 -= mm[y0 * SW + x1], v10 -= mm[y1 * SW + x0], v11 -= mm[y1 * SW + x1];
  }
  out[idx] = gain * (wy0 * (wx0 * v00 + wx1 * v01) + wy1 * (wx0 * v10 + wx1 * v11));
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_relevancy_store_pack(const float* maps, int32_t P, int32_t H, int32_t W, int32_t SH, int32_t SW,
                                           float* out, void* stream) {
  SB_REQUIRE(maps && out && P > 0 && H > 0 && W > 0 && SH > 0 && SW > 0, "semabs_relevancy_store_pack: bad arguments");
  store_pack_kernel<<<(SH * SW + 255) / 256, 256, 0, (cudaStream_t)stream>>>(maps, P, H, W, SH, SW, out);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_relevancy_store_unpack(const float* stored, const int32_t* rows, int32_t K, int32_t mean_row, int32_t SH,
                                             int32_t SW, int32_t H, int32_t W, float gain, float* out, void* stream) {
  SB_REQUIRE(stored && rows && out && K > 0 && H > 0 && W > 0 && SH > 0 && SW > 0, "semabs_relevancy_store_unpack: bad arguments");
  const long long n = (long long)K * H * W;
  store_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(stored, rows, K, mean_row, SH, SW, H, W, gain, out);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
