// ColorJitter on the device for the reference's test-time augmentation copies (CLIP/clip/__init__.py:55-57, 246-247:
// torchvision.transforms.ColorJitter(brightness=0.6, contrast=0.6, saturation=0.6, hue=0.1) applied to the image, 5 times
// in the "ours" saliency config).  The four operations follow torchvision's tensor implementation on uint8 images
// (transforms/_functional_tensor.py: _blend, rgb_to_grayscale, adjust_hue via _rgb2hsv / _hsv2rgb, convert_image_dtype),
// float32 arithmetic in the same operation order with no fused multiply-adds, so results are bit-identical to
// torchvision's CUDA / CPU tensor path for brightness, saturation and hue; contrast uses the exactly rounded mean of the
// grey image (integer sum) where torch reduces in float32 — at most one LSB apart on a handful of pixels.
// Images are HWC uint8 (what semabs_tile_preprocess consumes); one thread per pixel.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

__device__ __forceinline__ uint8_t to_u8_trunc(float v) {  // clamp(0, 255).to(uint8)
  v = fminf(fmaxf(v, 0.f), 255.f);
  return uint8_t(int(v));
}
__device__ __forceinline__ float grey_u8(float r, float g, float b) {  // (0.2989 r + 0.587 g + 0.114 b).to(uint8)
  const float l = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
  return float(uint8_t(int(l)));
}
__device__ __forceinline__ float blend(float ratio, float inv_ratio, float x, float y) {
  return __fadd_rn(__fmul_rn(ratio, x), __fmul_rn(inv_ratio, y));
}

__global__ void __launch_bounds__(256) jitter_grey_sum_kernel(const uint8_t* __restrict__ img, long long npix,
                                                              unsigned long long* __restrict__ sum) {
  unsigned long long local = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x)
    local += (unsigned long long)grey_u8(float(img[3 * i]), float(img[3 * i + 1]), float(img[3 * i + 2]));
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(sum, local);
}

// op: 0 brightness, 1 contrast (other = grey mean), 2 saturation (other = own grey), 3 hue
__global__ void __launch_bounds__(256) jitter_op_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long npix,
                                                        int op, float factor, float inv_factor,
                                                        const unsigned long long* __restrict__ grey_sum) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float r = float(in[3 * i]), g = float(in[3 * i + 1]), b = float(in[3 * i + 2]);
  float o0, o1, o2;
  if (op == 3) {
    // convert_image_dtype(uint8 -> float32): x / 255
    const float fr = __fdiv_rn(r, 255.f), fg = __fdiv_rn(g, 255.f), fb = __fdiv_rn(b, 255.f);
    const float maxc = fmaxf(fr, fmaxf(fg, fb)), minc = fminf(fr, fminf(fg, fb));
    const bool eqc = maxc == minc;
    const float cr = __fsub_rn(maxc, minc);
    const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
    const float div = eqc ? 1.f : cr;
    const float rc = __fdiv_rn(__fsub_rn(maxc, fr), div), gc = __fdiv_rn(__fsub_rn(maxc, fg), div), bc = __fdiv_rn(__fsub_rn(maxc, fb), div);
    const float hr = (maxc == fr) ? __fsub_rn(bc, gc) : 0.f;
    const float hg = (maxc == fg && maxc != fr) ? __fsub_rn(__fadd_rn(2.f, rc), bc) : 0.f;
    const float hb = (maxc != fg && maxc != fr) ? __fsub_rn(__fadd_rn(4.f, gc), rc) : 0.f;
    float h = __fadd_rn(__fadd_rn(hr, hg), hb);
    h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
    // h = (h + hue_factor) % 1.0  (python-style remainder: result in [0, 1))
    h = __fadd_rn(h, factor);
    h = __fsub_rn(h, floorf(h));
    const float h6 = __fmul_rn(h, 6.f);
    const float fl = floorf(h6);
    const float f = __fsub_rn(h6, fl);
    int sect = int(fl) % 6;
    if (sect < 0) sect += 6;
    const float v = maxc;
    const float p = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, s)), 0.f), 1.f);
    const float q = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, f))), 0.f), 1.f);
    const float t = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f)))), 0.f), 1.f);
    float R, G, B;
    switch (sect) {
      case 0: R = v, G = t, B = p; break;
      case 1: R = q, G = v, B = p; break;
      case 2: R = p, G = v, B = t; break;
      case 3: R = p, G = q, B = v; break;
      case 4: R = t, G = p, B = v; break;
      default: R = v, G = p, B = q; break;
    }
    // convert_image_dtype(float32 -> uint8): x * (255 + 1 - 1e-3), truncated
    constexpr float K = 255.999f;
    out[3 * i] = uint8_t(int(__fmul_rn(R, K))), out[3 * i + 1] = uint8_t(int(__fmul_rn(G, K))), out[3 * i + 2] = uint8_t(int(__fmul_rn(B, K)));
    return;
  }
  if (op == 0) {
    o0 = blend(factor, inv_factor, r, 0.f), o1 = blend(factor, inv_factor, g, 0.f), o2 = blend(factor, inv_factor, b, 0.f);
  } else if (op == 1) {
    const float mean = float(double(*grey_sum) / double(npix));
    o0 = blend(factor, inv_factor, r, mean), o1 = blend(factor, inv_factor, g, mean), o2 = blend(factor, inv_factor, b, mean);
  } else {
    const float l = grey_u8(r, g, b);
    o0 = blend(factor, inv_factor, r, l), o1 = blend(factor, inv_factor, g, l), o2 = blend(factor, inv_factor, b, l);
  }
  out[3 * i] = to_u8_trunc(o0), out[3 * i + 1] = to_u8_trunc(o1), out[3 * i + 2] = to_u8_trunc(o2);
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_color_jitter_op(const uint8_t* in, uint8_t* out, int64_t npix, int32_t op, float factor, void* scratch8,
                                      void* stream) {
  SB_REQUIRE(in && out && npix > 0 && op >= 0 && op <= 3, "semabs_color_jitter_op: bad arguments");
  SB_REQUIRE(op != 1 || scratch8, "semabs_color_jitter_op: contrast needs an 8-byte device scratch word");
  cudaStream_t st = (cudaStream_t)stream;
  if (op == 1) {
    SB_CHECK_CUDA(cudaMemsetAsync(scratch8, 0, 8, st));
    const int blocks = int((npix + 255) / 256 < 1184 ? (npix + 255) / 256 : 1184);
    jitter_grey_sum_kernel<<<blocks, 256, 0, st>>>(in, npix, (unsigned long long*)scratch8);
    SB_CHECK_CUDA(cudaGetLastError());
  }
  // torch evaluates `ratio * img1 + (1.0 - ratio) * img2` with the python doubles ratio and (1.0 - ratio) rounded to float32
  const float inv = float(1.0 - double(factor));
  jitter_op_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(in, out, npix, op, factor, inv, (const unsigned long long*)scratch8);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
