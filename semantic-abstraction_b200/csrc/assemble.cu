// Tile-pyramid assembly of ClipWrapper.get_clip_saliency_convolve (reference: CLIP/clip/__init__.py:205-236):
// every tile's g x g relevance is bilinearly up-sampled to the tile size (F.interpolate, align_corners=False),
// overlap-added into one fp16 accumulator per tile size IN TILE-CREATION ORDER (the reference's accumulators are
// .half() even on CPU, __init__.py:149-153,227-229), divided by the coverage count (initialised to 1e-5,
// __init__.py:249-253) and averaged over tile sizes.  One thread per output pixel walks the tiles that cover it
// in the reference's order, so the fp16 rounding sequence — and therefore the arg-max pixel — is reproduced.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& w0, float& w1) {
  // ATen area_pixel_compute_source_index(align_corners=false): max(0, scale*(dst+0.5)-0.5)
  float src = scale * (float(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = int(src);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  w1 = src - float(i0);
  w0 = 1.0f - w1;
}

// rel: [P, n_tiles, g, g] fp32; tile_desc: int32 [n_tiles, 3] = (row0, col0, size) in creation order;
// size_order: the distinct tile sizes in the order the reference sums them (cropping_augmentations order).
__global__ void tile_assemble_kernel(const float* __restrict__ rel, const int* __restrict__ tile_desc, int n_tiles,
                                     const int* __restrict__ size_order, int n_sizes, int g, int H, int W, int P,
                                     float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P * H * W) return;
  const int x = int(idx % W), y = int((idx / W) % H), p = int(idx / ((long long)W * H));
  float total = 0.f;
  for (int s = 0; s < n_sizes; ++s) {
    const int size = size_order[s];
    const float scale = float(g) / float(size);
    __half acc = __float2half_rn(0.f);
    float cnt = 1e-5f;
    for (int tIdx = 0; tIdx < n_tiles; ++tIdx) {
      if (tile_desc[3 * tIdx + 2] != size) continue;
      const int r0 = tile_desc[3 * tIdx], c0 = tile_desc[3 * tIdx + 1];
      const int ly = y - r0, lx = x - c0;
      if (ly < 0 || lx < 0 || ly >= size || lx >= size) continue;
      int y0, y1, x0, x1;
      float wy0, wy1, wx0, wx1;
      bilinear_src(ly, scale, g, y0, y1, wy0, wy1);
      bilinear_src(lx, scale, g, x0, x1, wx0, wx1);
      const float* m = rel + (size_t(p) * n_tiles + tIdx) * g * g;
      const float top = wx0 * m[y0 * g + x0] + wx1 * m[y0 * g + x1];
      const float bot = wx0 * m[y1 * g + x0] + wx1 * m[y1 * g + x1];
      const float v = wy0 * top + wy1 * bot;
      acc = __float2half_rn(__half2float(acc) + v);
      cnt += 1.0f;
    }
    total += __half2float(acc) / cnt;
  }
  out[idx] = total / float(n_sizes);
}

// horizontal-flip test-time augmentation (__init__.py:170-204): (rel + flip_x(rel_of_flipped_tiles)) / 2
__global__ void flip_average_kernel(float* __restrict__ rel, const float* __restrict__ rel_flipped, long long n_maps,
                                    int g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_maps * g * g) return;
  const int x = int(idx % g);
  const long long rowbase = idx - x;
  rel[idx] = (rel[idx] + rel_flipped[rowbase + (g - 1 - x)]) / 2.0f;
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_tile_assemble(const float* rel, const int32_t* tile_desc, int32_t n_tiles,
                                    const int32_t* size_order, int32_t n_sizes, int32_t g, int32_t H, int32_t W,
                                    int32_t P, float* out, void* stream) {
  SB_REQUIRE(rel && tile_desc && size_order && out, "semabs_tile_assemble: null pointer");
  SB_REQUIRE(n_tiles > 0 && n_sizes > 0 && g > 0 && H > 0 && W > 0 && P > 0, "semabs_tile_assemble: bad shape");
  const long long n = (long long)P * H * W;
  tile_assemble_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rel, tile_desc, n_tiles,
                                                                                       size_order, n_sizes, g, H, W, P,
                                                                                       out);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_flip_average(float* rel, const float* rel_flipped, int64_t n_maps, int32_t g, void* stream) {
  SB_REQUIRE(rel && rel_flipped && n_maps > 0 && g > 0, "semabs_flip_average: bad arguments");
  const long long n = n_maps * g * g;
  flip_average_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rel, rel_flipped, n_maps, g);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
