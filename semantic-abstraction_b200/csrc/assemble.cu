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

// ---------------------------------------------------------------------------------------------------------------
// Tile preprocessing on the device: crop -> Pillow-exact bicubic resize to 224x224 -> /255 -> normalise
// (reference: `_transform`, CLIP/clip/clip_explainability.py:98-108, applied per tile in create_tiles,
// CLIP/clip/__init__.py:257-281 — the reference's self-declared bottleneck, :275).
// Pillow's 8-bit resampler (ImagingResampleHorizontal/Vertical_8bpc) is integer arithmetic: per output coordinate a
// window [xmin, xmin+xmax) of 22-bit fixed-point coefficients, accumulator initialised to 2^21, arithmetic shift by
// 22, clip to [0,255]; horizontal pass first, its uint8 result feeds the vertical pass.  The coefficient tables are
// built on the host exactly like precompute_coeffs / normalize_coeffs_8bpc (wrapper.pillow_bicubic_coeffs); this kernel
// evaluates both passes on the fly per output pixel (window^2 taps, <= 49 at 336 -> 224), which reproduces Pillow bit
// for bit, so the tiles — and everything downstream — are identical to the host path's.
// ---------------------------------------------------------------------------------------------------------------
namespace sb {

__global__ void __launch_bounds__(256)
tile_preprocess_kernel(const uint8_t* __restrict__ images, int H, int W, const int32_t* __restrict__ tiles,
                       const int32_t* __restrict__ coef, const int32_t* __restrict__ bounds, int kmax, int R,
                       float mean0, float mean1, float mean2, float std0, float std1, float std2,
                       float* __restrict__ out) {
  const int t = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= R * R) return;
  const int yy = pix / R, xx = pix % R;
  const int img = tiles[t * 5 + 0], row0 = tiles[t * 5 + 1], col0 = tiles[t * 5 + 2], sid = tiles[t * 5 + 4];
  const int32_t* kx = coef + (size_t(sid) * R + xx) * kmax;
  const int32_t* ky = coef + (size_t(sid) * R + yy) * kmax;
  const int xmin = bounds[(sid * R + xx) * 2], xcnt = bounds[(sid * R + xx) * 2 + 1];
  const int ymin = bounds[(sid * R + yy) * 2], ycnt = bounds[(sid * R + yy) * 2 + 1];
  const uint8_t* base = images + (size_t(img) * H + row0) * W * 3 + size_t(col0) * 3;
  int acc0 = 1 << 21, acc1 = 1 << 21, acc2 = 1 << 21;
  for (int y = 0; y < ycnt; ++y) {
    const uint8_t* rowp = base + size_t(ymin + y) * W * 3 + size_t(xmin) * 3;
    int h0 = 1 << 21, h1 = 1 << 21, h2 = 1 << 21;
    for (int x = 0; x < xcnt; ++x) {
      const int k = kx[x];
      h0 += int(rowp[3 * x]) * k, h1 += int(rowp[3 * x + 1]) * k, h2 += int(rowp[3 * x + 2]) * k;
    }
    h0 >>= 22, h1 >>= 22, h2 >>= 22;
    h0 = min(max(h0, 0), 255), h1 = min(max(h1, 0), 255), h2 = min(max(h2, 0), 255);
    const int k = ky[y];
    acc0 += h0 * k, acc1 += h1 * k, acc2 += h2 * k;
  }
  acc0 >>= 22, acc1 >>= 22, acc2 >>= 22;
  acc0 = min(max(acc0, 0), 255), acc1 = min(max(acc1, 0), 255), acc2 = min(max(acc2, 0), 255);
  // ToTensor (/255) then Normalize ((v - mean) / std), IEEE fp32 like the host path
  float* o = out + (size_t(t) * 3 * R + yy) * R + xx;
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(float(acc0), 255.f), mean0), std0);
  o[size_t(R) * R] = __fdiv_rn(__fsub_rn(__fdiv_rn(float(acc1), 255.f), mean1), std1);
  o[size_t(2) * R * R] = __fdiv_rn(__fsub_rn(__fdiv_rn(float(acc2), 255.f), mean2), std2);
}

}  // namespace sb

extern "C" int semabs_tile_preprocess(const uint8_t* images, int32_t n_images, int32_t H, int32_t W, const int32_t* tiles,
                                      int32_t n_tiles, const int32_t* coef, const int32_t* bounds, int32_t n_sizes,
                                      int32_t kmax, int32_t R, const float* mean3, const float* std3, float* out,
                                      void* stream) {
  SB_REQUIRE(images && tiles && coef && bounds && mean3 && std3 && out, "semabs_tile_preprocess: null pointer");
  SB_REQUIRE(n_images > 0 && H > 0 && W > 0 && n_tiles > 0 && n_sizes > 0 && kmax > 0 && R > 0,
             "semabs_tile_preprocess: bad arguments");
  dim3 grid((R * R + 255) / 256, n_tiles);
  sb::tile_preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, H, W, tiles, coef, bounds, kmax, R, mean3[0],
                                                                    mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
