// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>  // 0: ld x16 + wait each, 1: ld x32 + wait each, 2: 4 x (ld x32) then wait, 3: st x8 + wait, 4: ld x16, wait, 16 FMAs, st x8 (chain)
__global__ void k(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 32) & 255) + (warp >> 2) * 0;  // all warps of a quadrant read the same columns (bandwidth test)
    if (MODE == 0 || MODE == 4) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(base + col));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (MODE == 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * 1.0001f + 1.0f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + col), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
      } else {
        acc += __uint_as_float(r[0]) + __uint_as_float(r[15]);
      }
    } else if (MODE == 1 || MODE == 2) {
      const int reps = MODE == 2 ? 4 : 1;
      for (int q = 0; q < reps; ++q) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + ((col + q * 32) & 255)));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
    } else if (MODE == 3) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + col), "r"(r[0]), "r"(r[1]),
                   "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc + __uint_as_float(r[3]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}
template <int MODE>
void run(const char* name, int warps, double bytes_per_iter_per_warp) {
  long long* d;
  float* s;
  cudaMalloc(&d, 8 * 148);
  cudaMalloc(&s, 4);
  const int iters = 20000;
  k<MODE><<<148, warps * 32>>>(100, d, s);
  k<MODE><<<148, warps * 32>>>(iters, d, s);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  double cyc = double(h[0]) / iters;
  printf("%-34s warps %2d: %7.1f cycles / iteration / warp, %7.1f B/clk/SM  (%s)\n", name, warps, cyc, bytes_per_iter_per_warp * warps / cyc,
         cudaGetErrorString(e));
  cudaFree(d), cudaFree(s);
}
int main() {
  for (int w : {1, 4, 8, 16}) run<0>("ld x16 + wait", w, 32 * 16 * 4.0);
  for (int w : {1, 4, 8, 16}) run<1>("ld x32 + wait", w, 32 * 32 * 4.0);
  for (int w : {1, 4, 8, 16}) run<2>("4 x ld x32, one wait", w, 4 * 32 * 32 * 4.0);
  for (int w : {1, 4, 8, 16}) run<3>("st x8 + wait", w, 32 * 8 * 4.0);
  for (int w : {1, 4, 8, 16}) run<4>("ld x16, wait, 16 fma, st x8 chain", w, 32 * 16 * 4.0);
  return 0;
}
