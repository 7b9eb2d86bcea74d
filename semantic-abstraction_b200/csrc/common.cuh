// Host-side helpers shared by every translation unit of libsemabs_b200.so: status/last-error plumbing for the
// C ABI, the lazily resolved cuTensorMapEncodeTiled entry point (no link-time dependency on libcuda) and
// small device math helpers used by several kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace sb {

// ---- C-ABI status ----------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int num_sms();

#define SB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      sb::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

#define SB_REQUIRE(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      sb::set_last_error(__VA_ARGS__); \
      return 2;                       \
    }                                 \
  } while (0)

// ---- TMA descriptors -------------------------------------------------------------------------------
// rank-R tiled tensor map over fp16 data. dims/box innermost first; strides_bytes[i] = byte stride of dim i+1.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, CUtensorMapSwizzle swizzle);

// ---- device math -----------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.0f / (1.0f + expf(-x)); }
// ex2.approx + rcp.approx (2 MUFU + 3 FP32 instructions; relative error ~1e-6 over the QuickGELU input range, against ~25
// instructions of expf + IEEE division): for epilogues whose warps are issue / latency bound (gemm_epilogue.cuh)
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// QuickGELU (reference: CLIP/clip/model_explainability.py:197-199): x * sigmoid(1.702 x)
__device__ __forceinline__ float quick_gelu(float x) { return x * sigmoidf_precise(1.702f * x); }
__device__ __forceinline__ float quick_gelu_grad(float x) {
  float s = sigmoidf_precise(1.702f * x);
  return s + 1.702f * x * s * (1.0f - s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum and sum of squares of 16 consecutive channels, folded into 16/CPG GroupNorm groups (CPG channels per group, a
// power of two; CPG >= 16 -> one group).  Compile-time CPG: the first version divided by a run-time cpg per element and
// made the conv epilogue — not the MMAs — the critical path (+1.3 ms per level-0 conv).
template <int CPG>
__device__ __forceinline__ void group_sums16_t(const float (&v)[16], float (&ps)[8], float (&pq)[8]) {
  constexpr int NG = CPG >= 16 ? 1 : 16 / CPG;
  constexpr int W = CPG >= 16 ? 16 : CPG;
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) s += v[k * W + j], q += v[k * W + j] * v[k * W + j];
    ps[k] = s, pq[k] = q;
  }
}
__device__ __forceinline__ void group_sums16(const float (&v)[16], int cpg, float (&ps)[8], float (&pq)[8]) {
  switch (cpg) {
    case 2: group_sums16_t<2>(v, ps, pq); break;
    case 4: group_sums16_t<4>(v, ps, pq); break;
    case 8: group_sums16_t<8>(v, ps, pq); break;
    default: group_sums16_t<16>(v, ps, pq); break;
  }
}

}  // namespace sb
