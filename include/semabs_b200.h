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
  SEMABS_ACT_QUICKGELU = 1,      /* out_f32 = pre-activation v, out_f16 = v * sigmoid(1.702 v),          */
                                 /* out_aux16 = d/dv quickgelu(v) (fp16, kept for the backward sweep)    */
  SEMABS_ACT_MUL_AUX16 = 2       /* v *= aux16[(row % aux_rows), col]  (dgrad through QuickGELU)         */
};

typedef struct semabs_gemm_epilogue {
  const float* bias;      /* [N] or NULL                                                                  */
  const float* residual;  /* [M, ld_out] fp32 added after bias/activation, or NULL                        */
  const void* aux16;      /* [aux_rows, ld_aux] fp16, used by SEMABS_ACT_MUL_AUX16                        */
  int32_t aux_rows;
  int32_t ld_aux;
  void* out_aux16;        /* [M, ld_out_aux] fp16 or NULL, written by SEMABS_ACT_QUICKGELU                 */
  int32_t ld_out_aux;
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
/* Shapes that qualify for 128 x 256 tiles (N % 256 == 0 and at least two tiles per SM) run on CTA pairs by default
 * (gemm2.cu: tcgen05.mma.cta_group::2, one 256 x 256 tile per pair); enable = 0 keeps them on the single-CTA kernel
 * (A/B measurements and the cross-check in tests/test_gemm_gpu.py).  Process-wide switch, not thread-safe. */
int semabs_set_gemm_pair(int32_t enable);


/* ------------------------------------------------------------------------------------------------------------
 * CLIP ViT / text-transformer stages other than the GEMMs (vit_ops.cu, vit_attn.cu).
 * fp16 outputs named *16 are [rows, splits*width]: hi part in columns [0,width), lo part (splits == 2) after it.
 * ---------------------------------------------------------------------------------------------------------- */

/* Patch-embedding operand (conv1 with stride == kernel, CLIP/clip/model_explainability.py:304-310,325):
 * tiles [B,3,R,R] fp32 -> out16 [B*(R/patch)^2, splits*Kpad], column (c*patch + i)*patch + j, zero padded. */
int semabs_vit_im2col(const float* tiles, void* out16, int32_t B, int32_t R, int32_t patch, int32_t Kpad,
                      int32_t splits, void* stream);

/* LayerNorm in fp32, eps 1e-5 (model_explainability.py:188-194). Row m is read at x + m*x_stride.
 * Any of y32 [M,d], y16 [M,splits*d], mean [M], rstd [M] may be NULL (at least one of y32 / y16 is required). */
int semabs_layernorm_fwd(const float* x, int64_t x_stride, const float* gamma, const float* beta, float* y32,
                         void* y16, float* mean, float* rstd, int32_t M, int32_t d, int32_t splits, void* stream);

/* Token assembly + ln_pre (VisionTransformer.forward, model_explainability.py:326-344):
 * x_out[b,t,:] = LN((t == 0 ? cls : patch[b,t-1,:]) + pos[t,:]).  `pos` is the table the reference would add
 * (already passed through its interpolate_positional_emb quirk, auxiliary.py:24-38, when T != 50). */
int semabs_vit_embed_lnpre(const float* patch, const float* cls, const float* pos, const float* gamma,
                           const float* beta, float* x_out, int32_t B, int32_t T, int32_t d, void* stream);

/* LayerNorm input-gradient for M stacked cotangent rows that share x_rows forward rows (row m uses forward row
 * m % x_rows): dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) + dres, g = dy*gamma.  Output rows may be strided
 * (out_stride / out16_stride in elements); dres (optional) uses out_stride.  Replaces the LayerNorm node of the
 * torch.autograd.grad call at CLIP/clip/clip_gradcam.py:90-97. */
int semabs_layernorm_bwd(const float* dy, const float* dres, const float* x, int64_t x_stride, int32_t x_rows,
                         const float* mean, const float* rstd, const float* gamma, float* dx32, int64_t out_stride,
                         void* dx16, int64_t out16_stride, int32_t M, int32_t d, int32_t splits, void* stream);
/* Same operator with the cotangent dy given as fp16 rows [M, d] (the dgrad GEMM's out_f16): the LayerNorm backward is
 * HBM-bound at 14 bytes per element; an fp16 dy makes it 12 and halves the producing GEMM's output write. */
int semabs_layernorm_bwd_h(const void* dy16, const float* dres, const float* x, int64_t x_stride, int32_t x_rows,
                           const float* mean, const float* rstd, const float* gamma, float* dx32, int64_t out_stride,
                           void* dx16, int64_t out16_stride, int32_t M, int32_t d, int32_t splits, void* stream);

/* Multi-head self-attention forward, head dim 64 (multi_head_attention_forward, CLIP/clip/auxiliary.py:260-337).
 * qkv [B*T, 3d] fp32 with q pre-scaled (auxiliary.py:207); softmax(q k^T) — what the reference stores through its
 * hook (:334) — is written to probs [B*H,T,T] fp32 (optional) and/or probs16 [B*H,T,ld_p16] fp16 (optional, rows
 * zero-padded up to ld_p16; the backward kernels read this one); o32 (optional) [B*T,d]; o16 (optional)
 * [B*T,splits*d]. causal != 0 applies the text transformer's mask (model_explainability.py:452-458). T <= 416. */
int semabs_attn_fwd(const float* qkv, float* probs, void* probs16, int32_t ld_p16, float* o32, void* o16, int32_t B,
                    int32_t T, int32_t H, int32_t causal, int32_t splits, void* stream);

/* Same operator on tcgen05 / TMEM (vit_attn_tc.cu) for T <= 272: qkv16 [B*T, in_splits*3d] fp16 — the QKV GEMM's
 * out_f16 with out_f16_splits = in_splits, i.e. rows [hi(3d) | lo(3d)], q pre-scaled — instead of fp32 qkv.  With
 * in_splits = 2 both products run as hi*hi + lo*hi + hi*lo (fp32-grade).  probs16 pitch must be a multiple of 16.
 * When T % 128 == 1 (ViT-L/14: 257 tokens) the last query row is computed by three SIMT warps from the K / V tiles in
 * shared memory instead of a third 128-row MMA tile; its results obey the same tolerances. */
int semabs_attn_fwd_tc(const void* qkv16, int32_t in_splits, void* probs16, int32_t ld_p16, float* o32, void* o16,
                       int32_t o_splits, int32_t B, int32_t T, int32_t H, int32_t causal, void* stream);

/* semabs_attn_bwd on tcgen05 / TMEM for T <= 272 (same arguments and results; ld_p16 and ld_do multiples of 16):
 * row pass G = dO V^T -> dS in TMEM -> dQ = dS K ; column pass G^T = V dO^T -> relevance column sums, dS^T and A^T in
 * TMEM -> dK = dS^T Q, dV = A^T dO.  With need_dqkv == 0 only wpart is produced (one launch). */
int semabs_attn_bwd_tc(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                       const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart, void* dqkv16,
                       int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits, int32_t positive_only,
                       int32_t need_dqkv, void* stream);

/* Second generation of semabs_attn_bwd_tc (vit_attn_bwd2.cu; same arguments and results, T <= 272 with T % 128 in {0, 1} above
 * one tile — the ViT geometries): delta_i = dO_i . O_i from a separate bandwidth kernel into delta_ws, a dedicated control
 * warp for TMA / MMA issue, probability rows / columns held in registers across the P labels of a unit, block-wide tail.
 * The product path (ClipEngine) calls this one; the first generation stays as its cross-check. */
int semabs_attn_bwd_tc2(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                        const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart, void* dqkv16,
                        int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits, int32_t positive_only,
                        int32_t need_dqkv, void* stream);

/* Third generation (vit_attn_bwd3.cu; same arguments, results and shape rules as semabs_attn_bwd_tc2): the strip of each
 * (unit, label) item is cut into two column chunks that flow through  G -> element-wise -> dQ | dK, dV  as a software pipeline
 * (tcgen05.mma executes in issue order, so the next item's G chunk is queued right behind the product that reads the chunk),
 * K / V double-buffered over units in the row pass.  The product path (ClipEngine) calls this one. */
int semabs_attn_bwd_tc3(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32,
                        const void* dO16, int32_t ld_do, const float* r, float* delta_ws, float* wpart, void* dqkv16,
                        int32_t P, int32_t B, int32_t T, int32_t H, int32_t splits, int32_t positive_only,
                        int32_t need_dqkv, void* stream);

/* Debug aid of semabs_attn_bwd_tc3: device buffer of 2 x 16 x 64 int64 that the next launches fill with clock64 time stamps of
 * CTA 0's pipeline events (tools/attn_trace.py prints the timeline); null = off (default). */
int semabs_debug_attn_trace(long long* device_buf);
/* Same for semabs_attn_fwd_tc: 64 int64 slots, clock64 stamps of the CTA in the middle of the grid (slot 0 = start, 1 + 8 t ..
 * 7 + 8 t = softmax warp 4 in MMA tile t: before / after the wait for S, after the max pass, after the exp pass, after the
 * normalise-and-pack pass, after the wait for O, after the stores; 31 = TMEM freed; 32 + 4 w .. 34 + 4 w = SIMT tail warp w:
 * start, K / V in shared memory, row done). */
int semabs_debug_attn_fwd_trace(long long* device_buf);

/* Known-answer hook for the two tcgen05 operand forms the attention kernels add to the GEMM's (A operand in TMEM,
 * MN-major B in shared memory): D[128,64] = A16[128,Kd] * B16[Kd,64], Kd % 16 == 0, Kd <= 256; lbo / sbo are the
 * descriptor byte offsets under test.  Test infrastructure for tests/test_vit_kernels_gpu.py. */
int semabs_selftest_ts_mma(const void* A16, const void* B16, float* D, int32_t Kd, int32_t lbo, int32_t sbo, void* stream);

/* Attention backward for P stacked cotangents + the relevance term of ClipGradcam.interpret
 * (clip_gradcam.py:90-126).  qkv16 [B*T,ld_qkv >= 3d] fp16 (q pre-scaled; the first 3d columns of each row), probs16 [B*H,T,ld_p16] fp16 with
 * ld_p16 >= 16*ceil(T/16), o32 [B*T,d] forward attention output, dO16 [P*B*T, ld_do] fp16 = gradient w.r.t. the
 * pre-out-proj attention output.
 *   dA = dO V^T ;  wpart[pb,h,j] = (1/H) sum_i r[pb,i] * relu?(dA ⊙ A)[i,j]   (relu iff positive_only)
 *   dS = A ⊙ (dA - rowsum(dA ⊙ A)) ; dQ = scale dS K ; dK = dS^T Q ; dV = A^T dO -> dqkv16 [P*B*T, splits*3d]
 * delta_ws is a [P*B*H*T] fp32 workspace.  need_dqkv == 0 computes the relevance term only.  T <= 272. */
int semabs_attn_bwd(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const float* o32, const void* dO16,
                    int32_t ld_do, const float* r, float* delta_ws, float* wpart, void* dqkv16, int32_t P, int32_t B,
                    int32_t T, int32_t H, int32_t splits, int32_t positive_only, int32_t need_dqkv, void* stream);

/* Attention backward of the LAST transformer block, where only the class-token row of dO is non-zero (the logits
 * read x[:,0] only, model_explainability.py:349): same outputs as semabs_attn_bwd (wpart, dqkv16 for all T rows)
 * from dO16_cls [P*B, ld_do] = class-token rows of the cotangent. */
int semabs_attn_bwd_cls(const void* qkv16, int32_t ld_qkv, const void* probs16, int32_t ld_p16, const void* dO16_cls, int32_t ld_do,
                        const float* r, float* wpart, void* dqkv16, int32_t P, int32_t B, int32_t T, int32_t H,
                        int32_t splits, int32_t positive_only, int32_t need_dqkv, void* stream);

/* logits[b,p] = 100 * f_b/|f_b| . W[:,p] (ClipGradcam.forward, clip_gradcam.py:58-68) and the cotangent seed
 * d logits[b,p] / d f_b -> seed16 [P*B, splits*E] (row p*B + b). Either output may be NULL. W is [E,P] fp32. */
int semabs_clip_logit_seed(const float* f, const float* W, float* logits, void* seed16, int32_t B, int32_t P,
                           int32_t E, int32_t splits, void* stream);

/* Rollout restricted to row 0 of R (clip_gradcam.py:81-84,124-127): r <- e_0, then per block r += sum_h wpart. */
int semabs_rollout_init(float* r, int32_t PB, int32_t T, void* stream);
int semabs_rollout_update(float* r, const float* wpart, int32_t PB, int32_t H, int32_t T, void* stream);

/* Text side: x[n,t,:] = token_embedding[tokens[n,t]] + positional_embedding[t] (CLIP.encode_text,
 * model_explainability.py:468-471); W[e,c] = mean_t feat[c*nt+t, e]/|feat[c*nt+t,:]| (zeroshot_classifier,
 * clip_gradcam.py:12-27). tokens is int32 on the device. */
int semabs_text_embed(const int32_t* tokens, const float* table, const float* pos, float* x, int32_t n_texts,
                      int32_t ctx, int32_t d, void* stream);
int semabs_zeroshot_weights(const float* feat, float* W, int32_t n_classes, int32_t n_templates, int32_t E,
                            void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tile-pyramid assembly (assemble.cu) — ClipWrapper.get_clip_saliency_convolve, CLIP/clip/__init__.py:205-236.
 * rel [P, n_tiles, g, g] fp32; tile_desc int32 [n_tiles,3] = (row0, col0, size) in the reference's tile-creation
 * order (__init__.py:257-273); size_order = distinct tile sizes in cropping_augmentations order. Per size the
 * bilinear (align_corners=False) up-sampled tiles are added in order into an fp16 accumulator (the reference's
 * .half() buffers, :149-153,:227-229), divided by the coverage count (init 1e-5, :249-253); sizes are averaged.
 * out [P,H,W] fp32.  semabs_flip_average: rel = (rel + flip_x(rel_flipped)) / 2 (:170-204), [n_maps, g, g].
 * ---------------------------------------------------------------------------------------------------------- */
int semabs_tile_assemble(const float* rel, const int32_t* tile_desc, int32_t n_tiles, const int32_t* size_order,
                         int32_t n_sizes, int32_t g, int32_t H, int32_t W, int32_t P, float* out, void* stream);
int semabs_flip_average(float* rel, const float* rel_flipped, int64_t n_maps, int32_t g, void* stream);

/* ColorJitter on the device (jitter.cu) — the reference's test-time augmentation copies (CLIP/clip/__init__.py:55-57,246-247).
 * One operation of torchvision's tensor implementation on a uint8 HWC image: op 0 brightness, 1 contrast, 2 saturation,
 * 3 hue, `factor` as drawn by ColorJitter.get_params.  in / out [npix, 3] uint8 (may alias for ops 0-2 only when equal);
 * scratch8: 8 bytes of device memory (contrast's grey-image sum). */
int semabs_color_jitter_op(const uint8_t* in, uint8_t* out, int64_t npix, int32_t op, float factor, void* scratch8, void* stream);

/* Relevancy store, device half (store.cu) — SURVEY.md §8 f4.  Writer (generate_relevancy.py:95-111): maps [P,H,W] fp32 ->
 * out [P+1,SH,SW] fp32 = F.interpolate(mode="nearest-exact") to the storage grid, row P = the mean map over the labels.
 * Reader (dataset.py:817-872): out[k] [H,W] = gain * bilinear_align_corners_false(stored[rows[k]] - stored[mean_row]) (mean_row
 * < 0: nothing subtracted); rows int32 [K] on the device. */
int semabs_relevancy_store_pack(const float* maps, int32_t P, int32_t H, int32_t W, int32_t SH, int32_t SW, float* out,
                                void* stream);
int semabs_relevancy_store_unpack(const float* stored, const int32_t* rows, int32_t K, int32_t mean_row, int32_t SH, int32_t SW,
                                  int32_t H, int32_t W, float gain, float* out, void* stream);

/* Tile preprocessing on the device (assemble.cu) — the reference's `_transform` per tile (clip_explainability.py:98-108,
 * create_tiles __init__.py:257-281): crop a square tile out of a uint8 HWC image, Pillow-exact bicubic resize to RxR
 * (Pillow's 8-bit two-pass fixed-point resampler, bit for bit), /255, (v-mean)/std -> out [n_tiles,3,R,R] fp32.
 * images [n_images,H,W,3] uint8 (device); tiles int32 [n_tiles,5] = (image, row0, col0, size, size_id) (device);
 * coef int32 [n_sizes,R,kmax], bounds int32 [n_sizes,R,2] = (first input index, window length) (device): Pillow's
 * precompute_coeffs / normalize_coeffs_8bpc tables for size -> R (built by the host wrapper); mean3/std3 HOST. */
int semabs_tile_preprocess(const uint8_t* images, int32_t n_images, int32_t H, int32_t W, const int32_t* tiles,
                           int32_t n_tiles, const int32_t* coef, const int32_t* bounds, int32_t n_sizes, int32_t kmax,
                           int32_t R, const float* mean3, const float* std3, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Residual 3-D UNet stages (conv3d.cu, unet_ops.cu) — reference unet3d.py.
 * Internal activation layout is channels-last: raw fp32 [N,D,H,W,C] and, as MMA operand, fp16
 * [N,D,H,W,a_splits*C] (hi | lo). GroupNorm statistics buffers are fp64 [N,G,2] = (sum, sum of squares) and must
 * be zeroed by the caller before the producing kernel runs.
 * ---------------------------------------------------------------------------------------------------------- */

/* Implicit-GEMM convolution on tcgen05, TMA im2col by taps.
 *   kind 0: nn.Conv3d 3x3x3, padding 1, no bias (unet3d.py:16-17,55-61); weights w16 [C_out, w_splits*27*C_in],
 *           slice t = (kd*3+kh)*3+kw, i.e. conv.weight.permute(0,2,3,4,1)
 *   kind 1: 1x1x1 conv (final_conv, unet3d.py:578); w16 [C_out, w_splits*C_in]
 *   kind 2: nn.ConvTranspose3d k3 s2 p1 with output_size = 2x input (Upsampling, unet3d.py:428-440), one output
 *           parity class per call (parity bit2 = z, bit1 = y, bit0 = x; 8 calls cover the output);
 *           w16 [C_out, w_splits*27*C_in] = upsample.weight.permute(1,2,3,4,0)
 *   kind 3: adjoint of kind 2 = 3x3x3 conv, stride 2, padding 1 over x16 [N,2D,2H,2W,C_in] -> (D,H,W) (data gradient
 *           of the transposed conv); one INPUT parity class per call, the 8 calls are chained through
 *           residual == out32; w16 [C_out, w_splits*27*C_in] = upsample.weight.permute(0,2,3,4,1)
 * precise != 0 (needs a_splits == w_splits == 2) accumulates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (≈ fp32 accuracy).
 * Epilogue: + bias[C_out] (opt) + residual (opt, fp32, same layout as out32), ReLU (opt); writes out32
 * [N,Do,Ho,Wo,C_out] and/or out16 [.., o16_splits*C_out]; accumulates `stats` of what it wrote for a GroupNorm
 * with `groups` groups over C_out (opt). C_in must be 16, 32 or a multiple of 64; C_out a multiple of 16;
 * D,H,W powers of two with D*H*W >= 32. */
int semabs_conv3d(const void* x16, int32_t a_splits, const void* w16, int32_t w_splits, int32_t kind, int32_t parity,
                  int32_t N, int32_t D, int32_t H, int32_t W, int32_t C_in, int32_t C_out, int32_t precise,
                  const float* bias, const float* residual, int32_t relu, float* out32, void* out16,
                  int32_t o16_splits, double* stats, int32_t groups, void* stream);

/* Module-boundary layout conversions: x [N,C,S] fp32 (NCDHW, S = D*H*W) <-> y [N,S,Cpad] channels-last
 * (channels >= C zero-filled). The forward direction also accumulates the statistics for the first GroupNorm
 * (groups == 1 when C < num_groups, unet3d.py:72-73). */
int semabs_ncdhw_to_ndhwc(const float* x, float* y, int32_t N, int64_t S, int32_t C, int32_t Cpad, int32_t groups,
                          double* stats, void* stream);
int semabs_ndhwc_to_ncdhw(const float* x, float* y, int32_t N, int64_t S, int32_t C, void* stream);
/* Module input straight into the halo convolution's operand layout: x [N,C,S] fp32 -> y16 [N][2 Cpad / 8][S][8] fp16 (hi chunks,
 * then lo chunks; channels >= C zero) + the first GroupNorm's statistics (inference path with that GroupNorm folded into conv1). */
int semabs_ncdhw_to_planar(const float* x, void* y16, int32_t N, int64_t S, int32_t C, int32_t Cpad, int32_t groups, double* stats,
                           void* stream);
/* ConvTranspose3d(k = 3, s = 2, p = 1, output_size = 2x) + bias + skip sum (Upsampling.forward + summation joining, unet3d.py:
 * 385-396, 428-440) with all eight output-parity classes in ONE launch (conv3d_convt.cu): x16 [N,D,H,W,a_splits*C_in] fp16, w16 as
 * for semabs_conv3d kind 2 ([C_out, w_splits*27*C_in]), residual / out32 [N,2D,2H,2W,C_out] fp32, stats [N,groups,2] of the output
 * (optional; channels per group a power of two >= 4); out_planar (optional, W % 32 == 0): the same values as chunk-planar fp16 hi | lo
 * [N][2 C_out / 8][8 D H W][8] = the operand layout of semabs_conv3d_halo(_fused); one of out32 / out_planar may be null.
 * C_out % 32 == 0 and C_in % 64 == 0; other shapes: 8 x kind 2. */
int semabs_conv_transpose3d_s2(const void* x16, int32_t a_splits, const void* w16, int32_t w_splits, int32_t N, int32_t D, int32_t H,
                               int32_t W, int32_t C_in, int32_t C_out, int32_t precise, const float* bias, const float* residual,
                               float* out32, void* out_planar, double* stats, int32_t groups, void* stream);
/* final_conv (nn.Conv3d(f_maps[0], out_channels, 1), unet3d.py:565 / 619) fused with the conversion back to NCDHW (inference
 * path): x16 [N,S,splits*C_in] fp16 rows [hi | lo], w [C_out,C_in] fp32, bias [C_out] or null -> y [N,C_out,S] fp32; fp32 FMAs.
 * C_in in {16, 32, 64}. */
int semabs_final_conv1x1_ncdhw(const void* x16, int32_t splits, const float* w, const float* bias, float* y, int32_t N, int64_t S,
                               int32_t C_in, int32_t C_out, void* stream);

/* nn.GroupNorm apply (unet3d.py:78-83; biased variance, eps 1e-5, affine): x raw fp32 [N,S,C] + stats ->
 * y16 [N,S,splits*C] (planar == 0) or the chunk-planar layout [N][splits*C/8][S][8] consumed by
 * semabs_conv3d_halo (planar != 0). gamma/beta are [C] (zero for padded channels >= C_real). */
int semabs_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* y16,
                           int32_t N, int64_t S, int32_t C, int32_t C_real, int32_t groups, int32_t splits,
                           int32_t planar, void* stream);

/* Halo-resident 3x3x3 conv (padding 1, no bias) for the full-resolution level: W == 128, C_in, C_out in {16, 32}.
 * x16_planar: [N][a_splits*C_in/8][D][H][W][8] fp16 (see semabs_groupnorm_apply). w_img: per tap t = (kd*3+kh)*3+kw
 * and weight split s the UMMA no-swizzle core-matrix image [C_in/16][C_out/8][2][8][8]
 * (= W[co = g*8 + r][ci = kb*16 + kc*8 + e][tap]), taps outermost. Same epilogue options as semabs_conv3d.
 * Every activation row is loaded from L2 once per y-neighbour (3x) instead of once per tap and pass (81x). */
int semabs_conv3d_halo(const void* x16_planar, int32_t a_splits, const void* w_img, int32_t w_splits, int32_t N,
                       int32_t D, int32_t H, int32_t W, int32_t C_in, int32_t C_out, int32_t precise,
                       const float* residual, int32_t relu, float* out32, void* out16, int32_t o16_splits,
                       double* stats, int32_t groups, void* stream);
/* C_out = 32 shapes of semabs_conv3d_halo run on CTA pairs by default (conv3d_halo2.cu: tcgen05.mma.cta_group::2, each CTA
 * keeps its own output row's planes and one 16-channel half of the weights, so every activation row is read from shared
 * memory once for all 32 channels); enable = 0 forces the single-CTA kernel.  Process-wide switch, not thread-safe. */
int semabs_set_halo_pair(int32_t enable);
/* One sample, C_out = 32, CTA-pair kernel, with the GroupNorm of the CONSUMED tensor folded into the convolution (inference path
 * of ResidualUNet3D at the 128-wide level): w_img = pack of W * (gamma * rstd) of this sample, bias_cls [27][32] = the shift term
 * sum_taps W (beta - mean * gamma * rstd) for each of the 3 x 3 x 3 border classes (which taps fall inside the grid); the operand
 * x16_planar is then the RAW tensor as chunk-planar hi | lo fp16, written by the producing convolution through out_planar.
 * res_planar: residual read from such a tensor.  bias_cls / residual / res_planar / out32 / out16 / out_planar / stats optional. */
int semabs_conv3d_halo_fused(const void* x16_planar, int32_t a_splits, const void* w_img, int32_t w_splits, int32_t D, int32_t H,
                             int32_t C_in, int32_t precise, const float* bias_cls, const float* residual, const void* res_planar,
                             int32_t relu, float* out32, void* out16, int32_t o16_splits, void* out_planar, double* stats,
                             int32_t groups, void* stream);
/* The fold itself, for every sample in one launch: w [32][32][27] fp32 (conv.weight of a 32 -> 32 channel convolution), gamma /
 * beta [32] and stats (per sample stats_stride doubles: [groups][2] = sum, sum of squares over S voxels x 32 / groups channels) ->
 * w_img [N][55296] fp16 (ops.pack_halo_weights layout of W * gamma * rstd, hi | lo rows) and bias_cls [N][27][32] fp32. */
int semabs_fold_groupnorm_halo(const float* w, const float* gamma, const float* beta, const double* stats, int32_t stats_stride,
                               int32_t N, int64_t S, int32_t groups, void* w_img, float* bias_cls, void* stream);
/* Debug aid of the pair kernel: copies and clears its barrier time-out records (up to 64 x 8 int32); returns their number. */
int semabs_debug_halo_pair_dump(int32_t* out512);


/* nn.MaxPool3d(2) (Encoder.forward, unet3d.py:298,313-317) + statistics of the pooled tensor. */
int semabs_maxpool3d_2(const float* x, float* y, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                       int32_t groups, double* stats, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward of the UNet stages (unet_bwd.cu) — what `loss.backward()` (utils.py:404-411) does through
 * unet3d.py's 'gcr' units (GroupNorm -> Conv3d -> ReLU, :20-95), ExtResNetBlock (:243-259), MaxPool3d (:298),
 * ConvTranspose3d + skip sum (:385-396,428-440) and final_conv (:578).
 * Scaling protocol: an fp32 gradient tensor is true scale unless it is the raw output of a data-gradient conv run
 * on scaled fp16 operands; then a device float `scale` accompanies it (true = stored / scale). `amax` slots are
 * device uint32 holding the bit pattern of a non-negative float, zeroed by the caller, updated with atomicMax.
 * Data gradients: semabs_conv3d / semabs_conv3d_halo with adjoint weight packs (kind 0 with flipped taps and
 * swapped channel roles, kind 1 transposed, kind 3 for the transposed conv).
 * ---------------------------------------------------------------------------------------------------------- */

/* amax = max(amax, max|x|), n % 4 == 0. */
int semabs_absmax_f32(const float* x, int64_t n, void* amax_slot, void* stream);

/* fp32 [N,D,H,W,C] -> fp16 MMA operands, value * f * [mask > 0], f = 2^(12 - ilogb(*amax)) (1 if amax == NULL);
 * *scale_out = (*g_scale or 1) * f.  pad16 (optional): zero-padded channels-last [N, D+2, H+2, W+2, Cp], interior
 * only is written (ring and guard rows must already be zero); parity != 0: the 8 parity sub-grids as 8 consecutive
 * padded volumes [8][N, D/2+2, H/2+2, W/2+2, Cp] (sub-grid q = (z&1)<<2 | (y&1)<<1 | (x&1)).  op16 (optional):
 * op_layout 1 = channels-last [N,S,op_splits*C], 2 = chunk-planar [N][op_splits*C/8][S][8] (semabs_conv3d_halo
 * operand); op_splits == 2 stores hi | lo (lo = fp16 of the rounding remainder) for the 3-pass precise convolutions. */
int semabs_unet_bwd_pack(const float* g, const float* g_scale, const void* amax, const float* mask, int32_t N, int32_t D,
                         int32_t H, int32_t W, int32_t C, void* pad16, int32_t Cp, int32_t parity, void* op16,
                         int32_t op_layout, int32_t op_splits, float* scale_out, void* stream);

/* semabs_groupnorm_apply into the padded operand layout [N, D+2, H+2, W+2, Cp] (single fp16 split). */
int semabs_groupnorm_apply_padded(const float* x, const double* stats, const float* gamma, const float* beta, void* pad16,
                                  int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t C_real, int32_t groups,
                                  int32_t Cp, void* stream);

/* Weight gradient as a split-K reduction over flat padded voxels (mma.sync m16n8k16, fp32 accumulation):
 *   grad[(a*Cb_real + b)*KT + slot_k[slot]] (+)= (1 / *scale) * sum_p A16[p][a] * B16[p + off(slot)][b]
 * A16 / B16: padded channels-last fp16 with channel strides lda / ldb; p runs over [0, nvox) (multiple of 64; rows of
 * A outside the interior must be zero, B must be finite wherever A is non-zero... and readable for every offset).
 * Taps are grouped in <= 9 segments of <= 3 taps: tap j of segment s has off = seg_off[s] + seg_sh[3s+j]
 * (seg_sh in {0,1,2}) and output slot seg_slot[3s+j] < nslots.  seg_* are HOST arrays, slot_k a DEVICE int32 array.
 * workspace holds the per-split partial sums ([nsplit][nslots][Ca][Cb] fp32; nsplit is fitted to workspace_bytes).
 * Conv3d weight [Co,Ci,27]: A = scaled output gradient, B = normalised input, off = tap offset; ConvTranspose3d
 * weight [Ci,Co,27]: A = input, B = parity-split output gradient. */
int semabs_conv3d_wgrad(const void* A16, int32_t lda, int32_t Ca, int32_t Ca_real, const void* B16, int32_t ldb,
                        int32_t Cb, int32_t Cb_real, int64_t nvox, int32_t nseg, const int64_t* seg_off,
                        const int32_t* seg_ntaps, const int32_t* seg_sh, const int32_t* seg_slot, int32_t nslots,
                        const int32_t* slot_k_dev, int32_t KT, void* workspace, int64_t workspace_bytes,
                        const float* scale, float* grad, int32_t accumulate, void* stream);

/* GroupNorm backward (torch.nn.GroupNorm semantics). reduce: sums[N,C,2] += (sum_v dy, sum_v dy*x) (x may be NULL:
 * column sums only, for bias gradients); sums zeroed by the caller.  apply:
 *   dx = rstd*(gamma*dy - mean_g(gamma*dy) - xhat*mean_g(gamma*dy*xhat)) / *dy_scale
 *        + add * [add_mask > 0] / *add_scale  (+ dx when accumulate),   amax slot updated with max|dx|.
 * param_grads: dgamma[c] (+)= sum_n rstd*(sum dy x - mean sum dy) / *scale, dbeta[c] (+)= sum_n sum dy / *scale
 * (stats == NULL, dgamma == NULL: bias gradient of a convolution). */
int semabs_groupnorm_bwd_reduce(const float* dy, const float* x, int32_t N, int64_t S, int32_t C, double* sums,
                                void* stream);
int semabs_groupnorm_bwd_apply(const float* dy, const float* dy_scale, const float* x, const double* stats,
                               const float* gamma, const double* sums, int32_t N, int64_t S, int32_t C, int32_t C_real,
                               int32_t groups, const float* add, const float* add_scale, const float* add_mask,
                               float* dx, int32_t accumulate, void* amax_slot, void* stream);
int semabs_groupnorm_param_grads(const double* sums, const double* stats, const float* scale, int32_t N, int64_t S,
                                 int32_t C, int32_t C_real, int32_t groups, float* dgamma, float* dbeta,
                                 int32_t accumulate, void* stream);

/* MaxPool3d(2) backward: g [N,D/2,H/2,W/2,C] (/ *g_scale) routed to the first maximum (z,y,x scan order) of each
 * 2x2x2 cell of x [N,D,H,W,C]; dx written (or added to, accumulate != 0); amax slot updated with max|dx|. */
int semabs_maxpool3d_2_bwd(const float* g, const float* g_scale, const float* x, int32_t N, int32_t D, int32_t H,
                           int32_t W, int32_t C, float* dx, int32_t accumulate, void* amax_slot, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Point <-> voxel stages of SemAbs3D / SemAbsVOOL (points.cu) — reference net.py.
 * Grid description (HOST pointers, 3 values each): neg_lc = -lower_corner, scale = (shape-1)/(uc-lc) computed
 * in fp32 like VirtualGrid.get_points_grid_idxs (net.py:84-113), shape = (X,Y,Z).
 * ---------------------------------------------------------------------------------------------------------- */

/* pts_feat_extractor (Linear(3+F,h) LeakyReLU Linear(h,h) LeakyReLU Linear(h,C); net.py:358-367,395-404) fused with
 * VirtualGrid.scatter_points (net.py:185-201) using the reduce method the reference really applies: MEAN, empty
 * voxels 0.  xyz [N/xyz_div, npts, 3] (sample n reads batch n/xyz_div: SemAbs3D repeats xyz over patches,
 * net.py:386-390), feat [N, npts, F].  Weights are TRANSPOSED ([in][out]) fp32 device arrays.  use_mlp == 0
 * scatters the raw features.  Output vol [N, X*Y*Z, Cpad] channels-last fp32 (this call zeroes it and `cnt`
 * [N, X*Y*Z]); `stats` (optional, zeroed by caller) receives the statistics of the UNet's first GroupNorm. */
int semabs_points_to_voxels(const float* xyz, int32_t xyz_div, const float* feat, int32_t N, int32_t npts, int32_t F,
                            int32_t use_mlp, int32_t hidden, int32_t C, const float* w1t, const float* b1,
                            const float* w2t, const float* b2, const float* w3t, const float* b3,
                            const float* neg_lc, const float* scale, const int32_t* shape, float* vol, float* cnt,
                            int32_t Cpad, int32_t groups, double* stats, void* stream);

/* ImplicitVolumetricDecoder.forward (net.py:215-256): trilinear feature fetch (grid_sample bilinear / border /
 * align_corners=True fed (x,y,z) in (W,H,D) order, indices normalised by shape instead of shape-1 — both quirks
 * reproduced) + Linear(Cin,Hs) LeakyReLU Linear(Hs,out_dim). vol0/vol1 channels-last [N,X,Y,Z,C0] (vol1 optional:
 * the VOOL sampler reads target | reference volumes, net.py:556). query [N,nq,3]. With emb [N,out_dim] the output
 * is cosine_similarity(mlp_out, emb[n]) / temperature -> out [N,nq] (PointingAttention.cosine_sim, net.py:300-309);
 * otherwise out [N,nq,out_dim]. */
int semabs_sample_decode(const float* vol0, const float* vol1, int32_t C0, const float* query, int32_t N, int32_t nq,
                         const float* neg_lc, const float* scale, const int32_t* shape, int32_t concat_xyz,
                         const float* w1t, const float* b1, const float* w2t, const float* b2, int32_t Hs,
                         int32_t out_dim, const float* emb, float temperature, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward of the point <-> voxel stages (points_bwd.cu) — net.py:204-256, :300-309, :185-201, :358-367 under
 * loss.backward().  Stage 1 kernels recompute the forward, route feature gradients and write one scratch row per item;
 * weight gradients are semabs_outer_reduce_f32 over those rows, bias gradients column sums.
 * ---------------------------------------------------------------------------------------------------------- */

/* out[r*ld_out + c] += scale * sum_p A[p*lda + r] * B[p*ldb + c]   (r < R, c < Cc, p < P; fp32; out is accumulated
 * into with atomics, so the caller zero-initialises it). nn.Linear weight gradient: A = output deltas, B = inputs. */
int semabs_outer_reduce_f32(const float* A, int32_t lda, int32_t R, const float* B, int32_t ldb, int32_t Cc, int64_t P,
                            float scale, float* out, int32_t ld_out, void* stream);

/* Backward of semabs_sample_decode. dout: gradient of its output ([N,nq,out_dim], or [N,nq] with emb).
 * w1 / w2 are the nn.Linear weights in their own layout ([Hs][Cin], [out_dim][Hs]), w1t / w2t their transposes.
 * dvol0 / dvol1 (zero-initialised by the caller, may be NULL): trilinear scatter-add of the feature gradient;
 * demb [N,out_dim] (zero-initialised, optional): gradient of the cosine head's query embedding.
 * scratch [N*nq][ld] receives per query: inputs (features | normalised xyz) at 0, hidden activations at off_h,
 * output deltas at off_do, hidden deltas at off_dp. */
int semabs_sample_decode_bwd(const float* vol0, const float* vol1, int32_t C0, const float* query, int32_t N, int32_t nq,
                             const float* neg_lc, const float* scale, const int32_t* shape, int32_t concat_xyz,
                             const float* w1t, const float* w1, const float* b1, const float* w2t, const float* w2,
                             const float* b2, int32_t Hs, int32_t out_dim, const float* emb, float temperature,
                             const float* dout, float* dvol0, float* dvol1, float* demb, float* scratch, int32_t ld,
                             int32_t off_h, int32_t off_do, int32_t off_dp, void* stream);

/* Backward of semabs_points_to_voxels (use_mlp != 0): dvol [N,S,Cpad] is the gradient of the voxelised volume, cnt the
 * per-voxel point counts the forward produced (scatter-MEAN: every point receives dvol[voxel] / cnt[voxel]).
 * w2t [hidden][hidden] is in-major (as in the forward), w2 / w3 are the nn.Linear weights ([out][in]).
 * scratch [N*npts][ld] receives per point: inputs (xyz | features, 8 floats) at 0, h1 at 8, h2 at 8+hidden, the three
 * layer deltas at off_d3 (C values), off_d2, off_d1 (hidden values each). */
int semabs_points_to_voxels_bwd(const float* xyz, int32_t xyz_div, const float* feat, int32_t N, int32_t npts, int32_t F,
                                int32_t hidden, int32_t C, const float* w1t, const float* b1, const float* w2t,
                                const float* w2, const float* b2, const float* w3, const float* neg_lc,
                                const float* scale, const int32_t* shape, const float* dvol, const float* cnt,
                                int32_t Cpad, float* scratch, int32_t ld, int32_t off_d3, int32_t off_d2, int32_t off_d1,
                                void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser side of the training step (train.cu) — reference utils.loop (utils.py:404-422).
 * ---------------------------------------------------------------------------------------------------------- */

/* Masked, weighted binary_cross_entropy_with_logits with mean reduction over the kept points + its gradient +
 * accuracy (train_ovssc.get_losses, train_ovssc.py:133-150; train_vool.py:163-185): kept = !ignore[i];
 * loss_acc[0] = sum_kept w*bce / #kept, loss_acc[1] = mean_kept((logit > 0) == label);
 * dlogits (optional) = w*(sigmoid(x) - y)/#kept on kept points, 0 elsewhere. weight / ignore may be NULL.
 * acc_ws: 3 doubles of workspace. */
int semabs_bce_with_logits(const float* logits, const float* labels, const float* weight, const uint8_t* ignore,
                           int64_t n, double* acc_ws, float* loss_acc, float* dlogits, void* stream);

/* Chunk table shared by the three calls below: a device array of n_chunks records
 *   { float* p; const float* g; float* m; float* v; int32_t n; int32_t tensor; }   (semabs_lamb_chunk_bytes() bytes)
 * covering every parameter tensor that has a gradient (<= 65536 elements per chunk, tensor = its index). */
int32_t semabs_lamb_chunk_bytes(void);

/* out[0] = sum over all chunks of |g|^2 (the squared total_norm of torch.nn.utils.clip_grad_norm_, utils.py:415). */
int semabs_grad_sumsq(const void* chunks, int32_t n_chunks, double* out, void* stream);
/* g *= min(1, max_grad_norm / (sqrt(grad_sumsq) + 1e-6)) in place (clip_grad_norm_). */
int semabs_clip_grads(const void* chunks, int32_t n_chunks, const double* grad_sumsq, float max_grad_norm,
                      void* stream);

/* One LAMB step for the whole model (arm/optim/lamb.py:59-127): m,v moments without bias correction,
 * adam_step = m/(sqrt(v)+eps) + wd*p, trust = clamp(|p|,0,10)/|adam_step| (1 if either is 0, or adam_mode),
 * p -= lr*trust*adam_step. If grad_sumsq != NULL the clip coefficient above is applied to g on the fly
 * (gradients are left untouched). norms_ws: 2*n_tensors doubles. */
int semabs_lamb_step(const void* chunks, int32_t n_chunks, int32_t n_tensors, double* norms_ws,
                     const double* grad_sumsq, float max_grad_norm, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int32_t adam_mode, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Evaluation metrics on the device (metrics.cu) — utils.prediction_analysis (utils.py:338-380) and
 * utils.voxelize_points (utils.py:617-665) as used by train_ovssc.get_detailed_stats (train_ovssc.py:20-78).
 * pred / label / ignore: uint8 {0,1} [N, npts] (N = scenes x patches). counts: uint64 [N,7] =
 * (true positives, predicted positives, label positives, union, false negatives, false positives, kept elements), from
 * which iou = tp/union, precision = tp/pred_pos, recall = tp/label_pos, false_negative = fn/kept, false_positive = fp/kept.
 * The voxelised variant first scatter-MAXes the three point features onto `shape` voxels (xyz [N/xyz_div, npts, 3];
 * grid description as for semabs_points_to_voxels), ignores voxels that hold an ignored point or no point at all, and
 * optionally returns the voxelised booleans [N, X*Y*Z]. flags_ws: uint32 [N, X*Y*Z] workspace.
 * ---------------------------------------------------------------------------------------------------------- */
int semabs_confusion_counts(const uint8_t* pred, const uint8_t* label, const uint8_t* ignore, int32_t N, int64_t npts,
                            uint64_t* counts, void* stream);
int semabs_voxelized_confusion_counts(const float* xyz, int32_t xyz_div, const uint8_t* pred, const uint8_t* label,
                                      const uint8_t* ignore, int32_t N, int64_t npts, const float* neg_lc,
                                      const float* scale, const int32_t* shape, uint32_t* flags_ws, uint64_t* counts,
                                      uint8_t* vox_pred, uint8_t* vox_label, uint8_t* vox_ignore, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMABS_B200_H_ */
