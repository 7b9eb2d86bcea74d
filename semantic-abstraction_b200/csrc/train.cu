// Optimiser-side kernels of the training step (reference: utils.loop train branch utils.py:404-422):
//   * masked / weighted BCE-with-logits loss + gradient + accuracy (train_ovssc.get_losses, train_ovssc.py:133-150;
//     train_vool.get_losses, train_vool.py:163-185)
//   * global gradient norm for clip_grad_norm_ (utils.py:415) and
//   * multi-tensor LAMB (arm/optim/lamb.py:59-127: no bias correction, weight norm clamped to [0, 10], trust ratio 1
//     when either norm is 0) with the clip coefficient folded into the gradient read.
// The reference walks its 121-134 parameter tensors in a Python loop with ~12 ATen launches each; here one chunk
// table drives three launches for the whole model.  All HBM-bound; float4 where alignment allows.
#include "../../include/semabs_b200.h"
#include "common.cuh"

namespace sb {

// -------------------------------------------------------------------------------------------------------------
// BCE with logits
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bce_logits(float x, float y) { return fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))); }

// pass 1: acc[0] += sum w*l over kept, acc[1] += #kept, acc[2] += #correct ((x > 0) == y) over kept
__global__ void __launch_bounds__(256) bce_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ w, const uint8_t* __restrict__ ignore,
                                                         long long n, double* __restrict__ acc) {
  float s = 0.f, c = 0.f, k = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (ignore && ignore[i]) continue;
    const float xi = x[i], yi = y[i];
    s += (w ? w[i] : 1.f) * bce_logits(xi, yi);
    k += 1.f;
    c += ((xi > 0.f) == (yi > 0.5f)) ? 1.f : 0.f;
  }
  __shared__ float red[3][8];
  s = warp_sum(s), k = warp_sum(k), c = warp_sum(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = s, red[1][warp] = k, red[2][warp] = c;
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(acc + threadIdx.x, t);
  }
}

// pass 2: loss = acc[0]/acc[1]; dx = w (sigmoid(x) - y) / acc[1] on kept points, 0 elsewhere
__global__ void __launch_bounds__(256) bce_finish_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ w, const uint8_t* __restrict__ ignore,
                                                         long long n, const double* __restrict__ acc,
                                                         float* __restrict__ loss_acc, float* __restrict__ dx) {
  const double kept = acc[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    loss_acc[0] = kept > 0 ? float(acc[0] / kept) : nanf("");  // mean over an empty selection is NaN in torch too
    loss_acc[1] = kept > 0 ? float(acc[2] / kept) : nanf("");
  }
  if (!dx) return;
  const float inv = kept > 0 ? float(1.0 / kept) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = 0.f;
    if (!(ignore && ignore[i])) g = (w ? w[i] : 1.f) * (sigmoidf_precise(x[i]) - y[i]) * inv;
    dx[i] = g;
  }
}

// -------------------------------------------------------------------------------------------------------------
// multi-tensor LAMB
// -------------------------------------------------------------------------------------------------------------
struct LambChunk {
  float* p;          // parameter chunk
  const float* g;    // gradient chunk
  float* m;          // exp_avg chunk
  float* v;          // exp_avg_sq chunk
  int32_t n;         // elements in this chunk
  int32_t tensor;    // tensor index (norm slots)
};

// norms[2t] += |p|^2 , norms[2t+1] += |adam_step|^2 after updating m, v ; gnorm_sq (optional) = sum |g|^2 of the model
__global__ void __launch_bounds__(256) lamb_moments_kernel(const LambChunk* __restrict__ chunks, double* __restrict__ norms,
                                                           const double* __restrict__ gnorm_sq, float max_grad_norm,
                                                           float beta1, float beta2, float eps, float weight_decay) {
  const LambChunk c = chunks[blockIdx.x];
  float coef = 1.f;
  if (gnorm_sq) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = float(sqrt(*gnorm_sq));
    coef = fminf(max_grad_norm / (total + 1e-6f), 1.0f);
  }
  float sp = 0.f, sa = 0.f;
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
    const float g = c.g[i] * coef, p = c.p[i];
    const float m = c.m[i] * beta1 + (1.f - beta1) * g;
    const float v = c.v[i] * beta2 + (1.f - beta2) * g * g;
    c.m[i] = m, c.v[i] = v;
    const float a = m / (sqrtf(v) + eps) + weight_decay * p;
    sp += p * p, sa += a * a;
  }
  __shared__ float red[2][8];
  sp = warp_sum(sp), sa = warp_sum(sa);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = sp, red[1][warp] = sa;
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(norms + 2 * c.tensor + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(256) lamb_apply_kernel(const LambChunk* __restrict__ chunks, const double* __restrict__ norms,
                                                         float lr, float eps, float weight_decay, int adam_mode) {
  const LambChunk c = chunks[blockIdx.x];
  const float wn = fminf(fmaxf(float(sqrt(norms[2 * c.tensor])), 0.f), 10.f);
  const float an = float(sqrt(norms[2 * c.tensor + 1]));
  float trust = (wn == 0.f || an == 0.f) ? 1.f : wn / an;
  if (adam_mode) trust = 1.f;
  const float step = lr * trust;
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
    const float p = c.p[i];
    const float a = c.m[i] / (sqrtf(c.v[i]) + eps) + weight_decay * p;
    c.p[i] = p - step * a;
  }
}

__global__ void __launch_bounds__(256) sumsq_chunks_kernel(const LambChunk* __restrict__ chunks, double* __restrict__ out) {
  const LambChunk c = chunks[blockIdx.x];
  float s = 0.f;
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) s += c.g[i] * c.g[i];
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

// in-place gradient scaling of torch.nn.utils.clip_grad_norm_ (for callers that want the clipped gradients themselves;
// semabs_lamb_step folds the same coefficient into its gradient read instead)
__global__ void __launch_bounds__(256) clip_scale_kernel(const LambChunk* __restrict__ chunks, const double* __restrict__ gnorm_sq,
                                                         float max_grad_norm) {
  const LambChunk c = chunks[blockIdx.x];
  const float coef = fminf(max_grad_norm / (float(sqrt(*gnorm_sq)) + 1e-6f), 1.0f);
  if (coef >= 1.0f) return;
  float* g = const_cast<float*>(c.g);
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) g[i] *= coef;
}

}  // namespace sb

using namespace sb;

extern "C" int semabs_bce_with_logits(const float* logits, const float* labels, const float* weight,
                                      const uint8_t* ignore, int64_t n, double* acc_ws, float* loss_acc, float* dlogits,
                                      void* stream) {
  SB_REQUIRE(logits && labels && acc_ws && loss_acc && n > 0, "semabs_bce_with_logits: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SB_CHECK_CUDA(cudaMemsetAsync(acc_ws, 0, 3 * sizeof(double), st));
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  bce_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, labels, weight, ignore, n, acc_ws);
  SB_CHECK_CUDA(cudaGetLastError());
  bce_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, labels, weight, ignore, n, acc_ws, loss_acc, dlogits);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_grad_sumsq(const void* chunks, int32_t n_chunks, double* out, void* stream) {
  SB_REQUIRE(chunks && out && n_chunks > 0, "semabs_grad_sumsq: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
  sumsq_chunks_kernel<<<n_chunks, 256, 0, st>>>((const LambChunk*)chunks, out);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int semabs_lamb_step(const void* chunks, int32_t n_chunks, int32_t n_tensors, double* norms_ws,
                                const double* grad_sumsq, float max_grad_norm, float lr, float beta1, float beta2,
                                float eps, float weight_decay, int32_t adam_mode, void* stream) {
  SB_REQUIRE(chunks && norms_ws && n_chunks > 0 && n_tensors > 0, "semabs_lamb_step: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SB_CHECK_CUDA(cudaMemsetAsync(norms_ws, 0, size_t(2) * n_tensors * sizeof(double), st));
  lamb_moments_kernel<<<n_chunks, 256, 0, st>>>((const LambChunk*)chunks, norms_ws, grad_sumsq, max_grad_norm, beta1, beta2,
                                                eps, weight_decay);
  SB_CHECK_CUDA(cudaGetLastError());
  lamb_apply_kernel<<<n_chunks, 256, 0, st>>>((const LambChunk*)chunks, norms_ws, lr, eps, weight_decay, adam_mode);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t semabs_lamb_chunk_bytes(void) { return (int32_t)sizeof(LambChunk); }

extern "C" int semabs_clip_grads(const void* chunks, int32_t n_chunks, const double* grad_sumsq, float max_grad_norm,
                                 void* stream) {
  SB_REQUIRE(chunks && grad_sumsq && n_chunks > 0, "semabs_clip_grads: bad arguments");
  clip_scale_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const LambChunk*)chunks, grad_sumsq, max_grad_norm);
  SB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
