/*
 * libsemabs_b200.so — C ABI of the B200-native hot path of real-stanford/semantic-abstraction.
 *
 * The reference is 100 % Python and has no FFI boundary of its own: the drop-in boundary is its
 * Python API (CLIP/clip/__init__.py:103 `ClipWrapper.get_clip_saliency`, net.py:383 `SemAbs3D.forward`,
 * unet3d.py:596 `Abstract3DUNet.forward`).  The Python mirror of that API (package `semantic-abstraction_b200`)
 * binds the entry points below through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success; on failure a non-zero status and `semabs_last_error()` describes it
 *     (thread-local, valid until the next failing call on the same thread). No exceptions cross the boundary.
 *   - all pointers are DEVICE pointers owned by the caller unless a parameter is documented as host memory;
 *     `stream` is a cudaStream_t passed as void* (NULL = default stream). Calls are asynchronous on `stream`.
 *   - fp16 buffers are IEEE binary16 ("half"), fp32 buffers are IEEE binary32. Shapes are row-major.
 *   - nothing in here touches torch types.
 */
#ifndef SEMABS_B200_H_
#define SEMABS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMABS_ABI_VERSION 1

const char* semabs_last_error(void);
int semabs_abi_version(void);
int semabs_device_sync(void);

/* ------------------------------------------------------------------------------------------------------------
 * Dense GEMM on tcgen05 (TMA-fed, TMEM accumulators):  C[M,N] = epilogue( A[M, a_splits*K] * B[N,K]^T )
 *
 * Replaces every torch.nn.functional.linear on the ViT path of the reference:
 *   in-proj  CLIP/clip/auxiliary.py:129, out-proj :340, MLP c_fc / c_proj CLIP/clip/model_explainability.py:210-217,
 *   patch embedding conv (stride == kernel ⇒ GEMM) model_explainability.py:304-310, visual proj :352,
 *   and their hand-written dgrad counterparts (the reference gets those from torch.autograd, clip_gradcam.py:90-97).
 *
 * A and B are fp16, K-major (row-major [rows, K]).  `a_splits` ∈ {1,2}: with 2, A holds a hi/lo fp16 split of an
 * fp32 activation ([M, 2K], hi in columns [0,K), lo in [K,2K)) and the kernel accumulates A_hi·Bᵀ + A_lo·Bᵀ
 * (CLIP weights are fp16-exact — convert_weights, model_explainability.py:501-527 — so this is ≈fp32 accurate).
 * K must be a multiple of 8 (16-byte rows); lda/ldb in elements, multiples of 8.
 * ---------------------------------------------------------------------------------------------------------- */
enum {
  SEMABS_ACT_NONE = 0,
  SEMABS_ACT_QUICKGELU = 1,      /* out_f32 = pre-activation v, out_f16 = v * sigmoid(1.702 v)           */
  SEMABS_ACT_QUICKGELU_GRAD = 2  /* v *= d/du quickgelu(u), u = aux[(row % aux_rows), col]                */
};

typedef struct semabs_gemm_epilogue {
  const float* bias;      /* [N] or NULL                                                                  */
  const float* residual;  /* [M, ld_out] fp32 added after bias/activation, or NULL                        */
  const float* aux;       /* [aux_rows, ld_aux] fp32, used by SEMABS_ACT_QUICKGELU_GRAD                   */
  int32_t aux_rows;
  int32_t ld_aux;
  float* out_f32;         /* [M, ld_out] or NULL                                                          */
  int32_t ld_out;
  void* out_f16;          /* [M, ld_out16] fp16 or NULL; with out_f16_splits == 2 the lo part goes to      */
  int32_t ld_out16;       /*   columns [N, 2N)                                                            */
  int32_t out_f16_splits; /* 1 or 2                                                                       */
  int32_t act;            /* SEMABS_ACT_*                                                                 */
  int32_t scale_cols;     /* columns [0, scale_cols) are multiplied by `scale` after the bias (q *= hd^-.5,*/
  float scale;            /*   auxiliary.py:207)                                                          */
} semabs_gemm_epilogue;

int semabs_gemm_f16(const void* A, int32_t lda, const void* B, int32_t ldb, int32_t M, int32_t N, int32_t K,
                    int32_t a_splits, const semabs_gemm_epilogue* ep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMABS_B200_H_ */
